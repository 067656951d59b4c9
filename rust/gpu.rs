//! `src/solvers/gpu.rs` -- binding of liblctp (include/lctp.h) for the reference tree (tprodanov/locityper v1.7.2).
//!
//! NOT COMPILED IN THE BUILD IMAGE OF THIS REPOSITORY (no cargo / rustc there): written against the reference sources,
//! every field access cites the reference file:line it reads.  The same ABI is exercised end to end by the
//! Python/ctypes mirror (locityper_b200/ffi.py, genotype.py) in the test suite, and `oracle/rust_diff.sh` is the
//! procedure that compiles this file into the reference and diffs its debug CSVs against the oracle's.
//!
//! Place next to `src/solvers/solve.rs`, add `#[cfg(feature = "cuda")] pub mod gpu;` to `src/solvers/mod.rs`, apply
//! `rust/reference_additions.rs` (six small accessors on otherwise private fields) and the `build.rs` module shown in
//! INTEGRATION.md section 1.
#![cfg(feature = "cuda")]

#[allow(non_camel_case_types, non_snake_case, non_upper_case_globals, dead_code)]
mod ffi { include!(concat!(env!("OUT_DIR"), "/bindings_lctp.rs")); }

use std::{ffi::CStr, fs, io::Write, path::Path, ptr};
use crate::{
    err::{error, add_path},
    ext::rand::XoshiroRng,
    seq::contigs::ContigId,
    solvers::solve::{Data, Stage},
};

pub const NONE_U32: u32 = 0xFFFF_FFFF;                    // LCTP_NONE_U32

fn check(rc: i32) -> crate::Result<()> {
    if rc == ffi::LCTP_OK as i32 { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(ffi::lctp_last_error()) }.to_string_lossy().into_owned();
    // RuntimeError (src/err.rs:26) keeps the per-locus isolation of genotype.rs:1346-1350 working.
    Err(error!(RuntimeError, "GPU genotype evaluation failed ({}): {}", rc, msg))
}

pub struct GpuContext(*mut ffi::lctp_ctx);
unsafe impl Send for GpuContext {}                       // Send, not Sync: one locus in flight per context
impl GpuContext {
    pub fn new(device: i32) -> crate::Result<Self> {
        let cfg = ffi::lctp_device_cfg { device, flags: 0, stream: ptr::null_mut(), max_resident_workers: 0, _pad: 0 };
        let mut ctx = ptr::null_mut();
        check(unsafe { ffi::lctp_init(&cfg, &mut ctx) })?;
        Ok(Self(ctx))
    }
}
impl Drop for GpuContext { fn drop(&mut self) { unsafe { ffi::lctp_destroy(self.0) } } }

/// Flat copy of `solve::Data` (src/solvers/solve.rs:254-273) in the layout of `lctp_locus`.
/// Owns the host vectors; the library retains no host pointer after `lctp_locus_upload` returns.
pub struct FlatLocus {
    pub n_haps: u32,
    pub n_reads: u32,
    pub ploidy: u32,
    pub is_paired: bool,
    pub n_genotypes: u64,
    /// `None` = all combinations with replacement in `gen_combinations_with_repl` order (no `--priors`).
    pub gt_tuples: Option<Vec<u32>>,
    pub priors: Option<Vec<f64>>,
    pub unmapped_prob: Vec<f64>,
    pub pa_off: Vec<u64>,
    pub pa_contig: Vec<u32>,
    pub pa_ln_prob: Vec<f64>,
    pub pa_mid1: Vec<u32>,
    pub pa_mid2: Vec<u32>,
    pub hap_len: Vec<u32>,
    pub hap_n_windows: Vec<u32>,
    pub hap_reg_start: Vec<u32>,
    pub window: u32,
    pub left_padding: u32,
    pub hap_pos_off: Vec<u64>,
    pub pos_weight: Vec<f64>,
    pub pos_gc: Vec<u8>,
    pub depth_k: u32,
    pub tweak: u32,
    pub depth_table: Vec<f64>,
    pub prob_diff: f64,
    pub lik_skew: f64,
    pub min_weight: f64,
    pub filt_diff: f64,
    pub prob_thresh: f64,
    pub dont_skip: bool,
    pub out_bams: u32,
    /// Names of the contigs (for dumps and `lctp_result_json`; never sent to the device).
    pub hap_names: Vec<String>,
}

impl FlatLocus {
    pub fn from_data(data: &Data) -> crate::Result<Self> {
        let contigs = &data.contigs;                                   // Data::contigs, solve.rs:256
        let n_haps = contigs.len();
        let reads = data.all_alns.reads();                             // AllAlignments::reads, locs.rs:1193-1195
        let n_reads = reads.len();
        let ploidy = data.genotypes[0].ploidy();                       // Genotype::ploidy, contigs.rs:433-435
        let params = &data.assgn_params;                               // model::Params, model/mod.rs:64-106

        // ---- reads: GrouppedAlignments (locs.rs:570-735).  `aln_pairs` is already sorted by contig ascending, then
        // ln_prob descending, at most MAX_USED_ALNS = 10 per contig (identify_contig_pair_alns, locs.rs:793-798).
        let mut unmapped_prob = Vec::with_capacity(n_reads);
        let mut pa_off = Vec::with_capacity(n_reads + 1);
        let n_pairs: usize = reads.iter().map(|r| r.alignment_pairs()).sum();      // locs.rs:638-640
        let (mut pa_contig, mut pa_ln_prob) = (Vec::with_capacity(n_pairs), Vec::with_capacity(n_pairs));
        let (mut pa_mid1, mut pa_mid2) = (Vec::with_capacity(n_pairs), Vec::with_capacity(n_pairs));
        pa_off.push(0_u64);
        for read in reads {
            unmapped_prob.push(read.unmapped_prob());                  // locs.rs:600-602
            for pair in read.aln_pairs() {                             // accessor added by reference_additions.rs (1)
                pa_contig.push(pair.contig_id().get());                // PairAlignment::contig_id, locs.rs:708-710
                pa_ln_prob.push(pair.ln_prob());                       // locs.rs:704-706
                pa_mid1.push(pair.middle1().unwrap_or(NONE_U32));      // locs.rs:713-715
                pa_mid2.push(pair.middle2().unwrap_or(NONE_U32));      // locs.rs:718-720
            }
            pa_off.push(pa_contig.len() as u64);
        }

        // ---- haplotype geometry and per-position window characteristics: ContigInfo (windows.rs:343-445)
        let (mut hap_len, mut hap_n_windows, mut hap_reg_start) = (Vec::new(), Vec::new(), Vec::new());
        let mut hap_pos_off = vec![0_u64];
        let (mut pos_weight, mut pos_gc) = (Vec::new(), Vec::new());
        let (mut window, mut left_padding) = (0, 0);
        for id in contigs.ids() {
            let info = data.contig_infos.get(id);                      // accessor (2): &self.infos[id.ix()]
            hap_len.push(info.contig_len());                           // accessor (3): windows.rs:345
            hap_n_windows.push(info.n_windows());                      // windows.rs:452-454
            hap_reg_start.push(info.region_start());                   // accessor (3): window_getter.start, windows.rs:38,420
            window = info.window_size();                               // windows.rs:457-459; identical for all contigs (depth.window_size())
            left_padding = info.left_padding();                        // accessor (3): windows.rs:347,384
            // One entry per element of `mov_info` (windows.rs:386).  neighb_info(start) reads
            // mov_info[start.saturating_sub(left_padding)] (windows.rs:440), so entry i is neighb_info(i + left_padding).
            for i in 0..info.mov_info_len() {                          // accessor (3): mov_info.len()
                let (ninfo, weight) = info.neighb_info(i + left_padding);      // windows.rs:439-445
                pos_weight.push(weight);
                pos_gc.push(ninfo.gc_content);                         // NeighbInfo::gc_content, windows.rs:322
            }
            hap_pos_off.push(pos_weight.len() as u64);
        }

        // ---- depth table: WindowDistr::ln_prob(k) = weight * distr.ln_pmf(k) (distr_cache.rs:34-39) with the cached
        // BayesCalc per GC bin (distr_cache.rs:61-75).  K >= 2R + 3 covers every reachable depth (each of the R read
        // pairs adds at most 2 to one window), so the device never needs ln_gamma.
        let depth_k = 2 * n_reads as u32 + 3;
        let mut depth_table = Vec::with_capacity(crate::bg::depth::GC_BINS * depth_k as usize);
        for gc in 0..crate::bg::depth::GC_BINS as u8 {                 // GC_BINS = 101, bg/depth.rs:42
            let distr = data.distr_cache.get_inner_distribution(gc);   // distr_cache.rs:78-80
            for k in 0..depth_k {
                depth_table.push(crate::math::distr::DiscretePmf::ln_pmf(&**distr, k));   // LinearCache::ln_pmf, lincache.rs:41-48
            }
        }

        // ---- genotypes: without --priors the list is exactly gen_combinations_with_repl over all contigs
        // (genotype.rs:1123-1126) and every prior is 0.0 -> send neither (the device enumerates the same order).
        let full = crate::ext::vec::count_combinations_with_repl(n_haps, ploidy) == data.genotypes.len()
            && data.priors.iter().all(|&p| p == 0.0);
        let (gt_tuples, priors) = if full { (None, None) } else {
            let mut t = Vec::with_capacity(data.genotypes.len() * ploidy);
            for gt in &data.genotypes { t.extend(gt.ids().iter().map(|id| id.get())); }    // contigs.rs:428-430
            (Some(t), Some(data.priors.clone()))
        };

        Ok(Self {
            n_haps: n_haps as u32, n_reads: n_reads as u32, ploidy: ploidy as u32,
            is_paired: data.is_paired_end,                             // solve.rs:272
            n_genotypes: data.genotypes.len() as u64,
            gt_tuples, priors, unmapped_prob, pa_off, pa_contig, pa_ln_prob, pa_mid1, pa_mid2,
            hap_len, hap_n_windows, hap_reg_start, window, left_padding, hap_pos_off, pos_weight, pos_gc,
            depth_k, depth_table,
            tweak: params.tweak.expect("set_tweak_size must have run (genotype.rs:1281-1282)"),   // model/mod.rs:179-186
            prob_diff: params.prob_diff, lik_skew: params.lik_skew, min_weight: params.min_weight,
            filt_diff: params.filt_diff, prob_thresh: params.prob_thresh,
            dont_skip: params.dont_skip, out_bams: params.out_bams as u32,
            hap_names: contigs.names().iter().map(|s| s.to_string()).collect(),
        })
    }

    pub fn as_c(&self) -> ffi::lctp_locus {
        ffi::lctp_locus {
            n_haps: self.n_haps, n_reads: self.n_reads, ploidy: self.ploidy, is_paired: self.is_paired as u32,
            n_genotypes: self.n_genotypes,
            gt_tuples: self.gt_tuples.as_ref().map_or(ptr::null(), |v| v.as_ptr()),
            priors: self.priors.as_ref().map_or(ptr::null(), |v| v.as_ptr()),
            unmapped_prob: self.unmapped_prob.as_ptr(),
            pa_off: self.pa_off.as_ptr(), pa_contig: self.pa_contig.as_ptr(), pa_ln_prob: self.pa_ln_prob.as_ptr(),
            pa_mid1: self.pa_mid1.as_ptr(), pa_mid2: self.pa_mid2.as_ptr(),
            hap_len: self.hap_len.as_ptr(), hap_n_windows: self.hap_n_windows.as_ptr(),
            hap_reg_start: self.hap_reg_start.as_ptr(),
            window: self.window, left_padding: self.left_padding,
            hap_pos_off: self.hap_pos_off.as_ptr(), pos_weight: self.pos_weight.as_ptr(), pos_gc: self.pos_gc.as_ptr(),
            depth_k: self.depth_k, tweak: self.tweak, depth_table: self.depth_table.as_ptr(),
            prob_diff: self.prob_diff, lik_skew: self.lik_skew, min_weight: self.min_weight,
            filt_diff: self.filt_diff, prob_thresh: self.prob_thresh,
            dont_skip: self.dont_skip as u32, out_bams: self.out_bams,
        }
    }

    /// `.lcti` dump: one raw little-endian file per array plus `meta.json`, read by tools/lcti.py (rust_diff.sh).
    pub fn dump(&self, dir: &Path) -> crate::Result<()> {
        fs::create_dir_all(dir).map_err(add_path!(dir))?;
        fn raw<T: Copy>(dir: &Path, name: &str, v: &[T]) -> crate::Result<()> {
            let path = dir.join(name);
            let bytes = unsafe { std::slice::from_raw_parts(v.as_ptr() as *const u8, std::mem::size_of_val(v)) };
            fs::File::create(&path).and_then(|mut f| f.write_all(bytes)).map_err(add_path!(path))
        }
        raw(dir, "unmapped_prob.f64", &self.unmapped_prob)?;
        raw(dir, "pa_off.u64", &self.pa_off)?;
        raw(dir, "pa_contig.u32", &self.pa_contig)?;
        raw(dir, "pa_ln_prob.f64", &self.pa_ln_prob)?;
        raw(dir, "pa_mid1.u32", &self.pa_mid1)?;
        raw(dir, "pa_mid2.u32", &self.pa_mid2)?;
        raw(dir, "hap_len.u32", &self.hap_len)?;
        raw(dir, "hap_n_windows.u32", &self.hap_n_windows)?;
        raw(dir, "hap_reg_start.u32", &self.hap_reg_start)?;
        raw(dir, "hap_pos_off.u64", &self.hap_pos_off)?;
        raw(dir, "pos_weight.f64", &self.pos_weight)?;
        raw(dir, "pos_gc.u8", &self.pos_gc)?;
        raw(dir, "depth_table.f64", &self.depth_table)?;
        if let Some(t) = &self.gt_tuples { raw(dir, "gt_tuples.u32", t)?; }
        if let Some(p) = &self.priors { raw(dir, "priors.f64", p)?; }
        let meta = json::object! {
            n_haps: self.n_haps, n_reads: self.n_reads, ploidy: self.ploidy, is_paired: self.is_paired,
            n_genotypes: self.n_genotypes, window: self.window, left_padding: self.left_padding,
            depth_k: self.depth_k, tweak: self.tweak,
            // f64 parameters as bit patterns: the text form of a float is exactly what this dump must not depend on
            prob_diff_bits: self.prob_diff.to_bits(), lik_skew_bits: self.lik_skew.to_bits(),
            min_weight_bits: self.min_weight.to_bits(), filt_diff_bits: self.filt_diff.to_bits(),
            prob_thresh_bits: self.prob_thresh.to_bits(),
            dont_skip: self.dont_skip, out_bams: self.out_bams,
            hap_names: self.hap_names.clone(),
        };
        let path = dir.join("meta.json");
        fs::write(&path, meta.dump()).map_err(add_path!(path))
    }
}

/// `Stage` (solve.rs:138-148) -> `lctp_stage`.  Needs accessor (4): `Solver::lctp_params()` on Greedy / SimAnneal.
pub fn stage_to_c(stage: &Stage) -> crate::Result<ffi::lctp_stage> {
    let p = stage.solver.lctp_params().ok_or_else(||
        error!(RuntimeError, "Solver {} has no device implementation", stage.solver))?;
    Ok(ffi::lctp_stage {
        kind: p.kind, attempts: u32::from(stage.attempts), in_size: stage.in_size as u64,
        best_start: p.best_start as u32, _pad: 0,
        sample_size: p.sample_size as u64, plato_size: p.plato_size as u64, anneal_steps: p.anneal_steps as u64,
        init_prob: p.init_prob,
    })
}

pub struct GpuLocus<'c> { h: *mut ffi::lctp_locus_h, _ctx: &'c GpuContext }
impl<'c> GpuLocus<'c> {
    pub fn upload(ctx: &'c GpuContext, flat: &FlatLocus) -> crate::Result<Self> {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::lctp_locus_upload(ctx.0, &flat.as_c(), &mut h) })?;
        Ok(Self { h, _ctx: ctx })
    }

    /// Replaces the body of `run_filter` (solve.rs:87-122): `ixs` in/out, sorted survivors.
    pub fn prefilter(&self, ixs: &mut Vec<usize>, out_size: usize, threads: usize) -> crate::Result<()> {
        let mut buf: Vec<u64> = ixs.iter().map(|&i| i as u64).collect();
        let mut n_out = 0_usize;
        check(unsafe { ffi::lctp_prefilter(self.h, buf.as_mut_ptr(), buf.len(), out_size, threads, &mut n_out,
            ptr::null_mut()) })?;
        ixs.clear();
        ixs.extend(buf[..n_out].iter().map(|&i| i as usize));
        Ok(())
    }

    /// Replaces the dispatch + collect of one stage (solve.rs:1047-1081).  `ixs` is the already shuffled list
    /// (solve.rs:1051), `worker_off` its cut into chunks (solve.rs:1052-1062), `worker_rng` the workers' streams
    /// (solve.rs:1013-1017; updated in place).  Returns per position (lik_mean, lik_var, assgn_counts).
    pub fn solve_stage(&self, stage: &ffi::lctp_stage, ixs: &[u64], worker_off: &[u64], worker_rng: &mut [[u64; 4]],
        counts_cap_per_gt: Option<usize>) -> crate::Result<Vec<(f64, f64, Option<Vec<u16>>)>>
    {
        let n = ixs.len();
        let (mut mean, mut var) = (vec![0.0_f64; n], vec![0.0_f64; n]);
        let mut n_alns = vec![0_u64; n];
        // assgn_counts (Prediction::assgn_counts, solve.rs:366; consumed by write_bam, bam.rs:369-384) are only needed
        // for genotypes that can reach the output, i.e. in the last executed stage (solve.rs:463,472 drop the others).
        let (mut counts_off, mut counts) = match counts_cap_per_gt {
            Some(cap) => (vec![0_u64; n + 1], vec![0_u16; n * cap]),
            None => (Vec::new(), Vec::new()),
        };
        let want = counts_cap_per_gt.is_some();
        check(unsafe { ffi::lctp_solve_stage(self.h, stage, ixs.as_ptr(), worker_off.as_ptr(), worker_off.len() - 1,
            worker_rng.as_mut_ptr() as *mut u64, mean.as_mut_ptr(), var.as_mut_ptr(), ptr::null_mut(),
            if want { counts_off.as_mut_ptr() } else { ptr::null_mut() },
            if want { counts.as_mut_ptr() } else { ptr::null_mut() }, counts.len() as u64,
            n_alns.as_mut_ptr(), ptr::null_mut()) })?;
        Ok((0..n).map(|j| {
            let c = if want { Some(counts[counts_off[j] as usize..counts_off[j + 1] as usize].to_vec()) } else { None };
            (mean[j], var[j], c)
        }).collect())
    }
}
impl Drop for GpuLocus<'_> { fn drop(&mut self) { unsafe { ffi::lctp_locus_free(self.h) } } }

/// `XoshiroRng` <-> `[u64; 4]` through the accessors (5) added to src/ext/rand.rs (rand_xoshiro keeps `s` private;
/// `Xoshiro256PlusPlus` implements serde `Serialize`/`Deserialize` with the `serde` feature, which is the portable
/// route: `[u64; 4]` is exactly its serialised form).
pub fn rng_state(rng: &XoshiroRng) -> [u64; 4] { crate::ext::rand::state_of(rng) }
pub fn set_rng_state(rng: &mut XoshiroRng, s: [u64; 4]) { *rng = crate::ext::rand::from_state(s) }

/// Upper bound on the candidates of one genotype: R + sum over its contigs of that contig's pair alignments (what
/// the library sizes its own buffers by); a generous per-genotype capacity for the counts buffer.
pub fn counts_cap(flat: &FlatLocus) -> usize {
    let mut per_hap = vec![0_usize; flat.n_haps as usize];
    for &c in &flat.pa_contig { per_hap[c as usize] += 1; }
    per_hap.sort_unstable_by(|a, b| b.cmp(a));
    flat.n_reads as usize + per_hap.iter().take(flat.ploidy as usize).sum::<usize>()
}

#[allow(dead_code)]
fn _contig_id_is_u32(id: ContigId) -> u32 { id.get() }   // ContigId(u32), contigs.rs:26-41

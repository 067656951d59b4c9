//! Accessors the binding (rust/gpu.rs) needs on fields that are private to their modules in the reference tree.
//! Each block says where it goes; none of them changes behaviour.  NOT COMPILED HERE (no Rust toolchain in this image).

// (1) src/model/locs.rs, inside `impl GrouppedAlignments` (next to `alignment_pairs`, :638-640)
/// All pair alignments of the read, sorted by contig (asc) and ln-probability (desc).
pub fn aln_pairs(&self) -> &[PairAlignment] { &self.aln_pairs }

// (2) src/model/windows.rs, inside `impl ContigInfos` (:577-699)
/// Window information of one contig.
pub fn get(&self, contig_id: ContigId) -> &ContigInfo { &self.infos[contig_id.ix()] }

// (3) src/model/windows.rs, inside `impl ContigInfo` (next to `n_windows`, :452-459)
pub fn contig_len(&self) -> u32 { self.contig_len }
pub fn left_padding(&self) -> u32 { self.left_padding }
pub fn region_start(&self) -> u32 { self.window_getter.start }     // WindowGetter is in the same module (:36-40)
pub fn mov_info_len(&self) -> u32 { self.mov_info.len() as u32 }

// (4) src/solvers/mod.rs: one more provided method on `trait Solver` (:49-75), overridden by the two stochastic solvers
pub struct LctpSolverParams { pub kind: u32, pub best_start: bool, pub sample_size: usize, pub plato_size: usize,
    pub anneal_steps: usize, pub init_prob: f64 }
// in `trait Solver`:
fn lctp_params(&self) -> Option<LctpSolverParams> { None }
// src/solvers/stoch.rs, `impl Solver for Greedy` (fields at :36-43):
fn lctp_params(&self) -> Option<super::LctpSolverParams> {
    Some(super::LctpSolverParams { kind: 0, best_start: self.best_start, sample_size: self.sample_size,
        plato_size: self.plato_size, anneal_steps: 0, init_prob: 0.0 })
}
// src/solvers/stoch.rs, `impl Solver for SimAnneal` (fields at :151-159):
fn lctp_params(&self) -> Option<super::LctpSolverParams> {
    Some(super::LctpSolverParams { kind: 1, best_start: false, sample_size: 0, plato_size: self.plato_size,
        anneal_steps: self.anneal_steps, init_prob: self.init_prob })
}

// (5) src/ext/rand.rs: state access for the by-value RNG hand-over.  `Xoshiro256PlusPlus` keeps `s: [u64; 4]`
// private but is `Serialize + Deserialize` (rand_xoshiro feature "serde") as exactly that array; without the feature
// the same can be done with `rng.clone()` + 4 x `SeedableRng::from_seed` bytes: the 32 seed bytes ARE the four state
// words in little-endian order (Xoshiro256PlusPlus::from_seed reads them with `read_u64_into`).
pub fn from_state(s: [u64; 4]) -> XoshiroRng {
    let mut seed = [0_u8; 32];
    for (chunk, word) in seed.chunks_exact_mut(8).zip(s) { chunk.copy_from_slice(&word.to_le_bytes()); }
    XoshiroRng::from_seed(seed)
}
pub fn state_of(rng: &XoshiroRng) -> [u64; 4] {
    // serde route (Cargo.toml: rand_xoshiro = { version = "0.8", features = ["serde"] })
    let v: [u64; 4] = bincode::deserialize(&bincode::serialize(rng).unwrap()).unwrap();
    v
}

// (6) src/command/genotype.rs, analyze_locus, right before `solve::solve(&data, ...)` (:1251): the dump that
// oracle/rust_diff.sh diffs against.  Environment-driven so that no CLI flag changes.
if let Ok(dir) = std::env::var("LCTP_DUMP_LCTI") {
    let flat = crate::solvers::gpu::FlatLocus::from_data(&data)?;
    flat.dump(&std::path::Path::new(&dir).join(locus.set.tag()))?;
    // the locus stream as it enters solve() (4 x u64, little endian), so the oracle starts from the same state
    let state = crate::ext::rand::state_of(&rng);
    let bytes: Vec<u8> = state.iter().flat_map(|w| w.to_le_bytes()).collect();
    std::fs::write(std::path::Path::new(&dir).join(locus.set.tag()).join("rng_state.u64"), bytes).unwrap();
}

// (7) src/model/locs.rs, `AllAlignments::load` (:1085-1186) with the upstream chain of liblctp.  The BAM
// reader, the name hash and the debug writers stay as they are; the loop body only COLLECTS what it used to compute:
//
//   while reader.has_more() {                                   // :1115
//       // per read end: push the raw records (cigar as the BAM u32 operations, tid -> contig id, interval, strand) and
//       // the per-group inputs of read_next_alns (:502-567): record.seq().len(), edit_dist_cache.get(len),
//       // contig_infos.neighb_complexity(&primary); remember group indices of the read: read_group[r] = [g1, g2] (-1 = none)
//       // and MateData (sequence, name) for both mates
//   }
//   let ends  = lctp_collect_read_ends(ctx, &read_ends)?;        // ok / best_edit / thr_dist / weight_factor / kept_rec / ln_prob
//   let uk    = lctp_unique_kmers_build(ctx, contig_set.seqs(), kmer_counts, k, hard, soft)?;   // UniqueKmers::new, :930-963
//   let wts   = lctp_read_weights(ctx, uk, mate_sequences, 2)?;  // calculate_read_weight, :968-1002 (+ the read_kmers debug rows)
//   // read_data.weight = product of the weight factors of its ends (:564) * wts.weight[r]
//   let (mates, status, out_read, counts) = lctp_group_reads_dev(ctx, &prelim)?;   // :1119-1137 + :1237-1288 without the transfer
//   counts.poorly_mapped / out_of_bounds / good_reads + few_kmers come from `counts` and `weight >= min_weight`
//   let pairs = lctp_pair_alignments_from(ctx, mates, &pairing_params)?;            // identify_*_alignments, :805-911
//   let locus = lctp_locus_upload_pairs(ctx, &flat_locus_without_pairs, pairs)?;    // solve::Data for the solvers, no host copy
//   // only when BAM output or debug tables are requested: lctp_pairs_fetch(pairs) -> GrouppedAlignments::aln_pairs
//
// Accessors this needs, none of which changes behaviour:
//   src/model/locs.rs  : `pub(crate)` on PrelimAlignments::{alns, best_edit, good_dist} (:266-274) is not needed any more --
//                        the protocol runs in the library; ReadData / MateData stay host structs (:569-598);
//   src/seq/counts.rs  : KmerCounts::k() and iter() are already public (:used by UniqueKmers::new);
//   src/bg/err_prof.rs : ErrorProfile::{match_prob, mismatch_prob, insert_prob, deletion_prob, clipping_prob} are public (:282-304)
//                        = lctp_alns::ln_* ; EditDistCache::get is public (:436-450);
//   src/bg/insertsz.rs : InsertDistr::ln_prob(size) for 0..=max contig length -> lctp_mates::ins_ln_pmf, insert_penalty() (:153-173).
// With `opt_hap_alns = Some(..)` (pairwise haplotype alignments in the database) the reference additionally transfers
// alignments between haplotypes (HapAlns::transfer_alignments, src/seq/transfer.rs:70-141, WFA2): that step is not in the
// library; a host that wants it runs it between lctp_collect_read_ends and lctp_group_reads on the fetched arrays.

/*
 * lctp.h -- C ABI of the B200-native genotype-evaluation hot path of locityper.
 *
 * This is the drop-in boundary.  The reference (tprodanov/locityper v1.7.2,
 * Rust) has no C ABI for this path; the precedent for the shape is its WFA2 binding
 * (/root/reference/build.rs:1-64, src/seq/wfa.rs:10-16).  A Rust maintainer binds these entry
 * points with bindgen from `build.rs` and swaps two bodies in src/solvers/solve.rs:
 *   (1) the `run_filter(...)` call at src/solvers/solve.rs:943-944        -> lctp_prefilter
 *   (2) the per-stage dispatch + collect in MainWorker::run :1047-1081 /
 *       solve_single_thread :814-843                                       -> lctp_solve_stage
 * or the whole of `solve::solve` (:926-981) minus file output                 -> lctp_solve.
 * See INTEGRATION.md for the Rust-side stubs.
 *
 * Conventions: plain pointers and sizes, caller owns every host buffer, the library owns device
 * memory behind opaque handles, no host pointer is retained after a call returns.  Return value
 * 0 = ok; LCTP_E_INVALID (-1) invalid argument; LCTP_E_CUDA (-2) CUDA failure (including "no
 * device": there is NO CPU fallback on this path); LCTP_E_CAPACITY (-3) capacity / unsupported
 * parameter (mirrors the asserts at src/model/assgn.rs:58, src/model/windows.rs:734).
 * The message of the last error on this thread is returned by lctp_last_error().
 */
#ifndef LCTP_H
#define LCTP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCTP_OK 0
#define LCTP_E_INVALID (-1)
#define LCTP_E_CUDA (-2)
#define LCTP_E_CAPACITY (-3)

#define LCTP_NONE_U32 0xFFFFFFFFu
#define LCTP_GC_BINS 101     /* src/bg/depth.rs:42 */
#define LCTP_MAX_PLOIDY 16    /* reference: ploidy < 256 (src/model/windows.rs:734); the device packs the contig index of a
                                 candidate into 4 bits */
#define LCTP_MAX_STAGES 8
#define LCTP_MAX_OUT 50      /* MAX_GENOTYPES, src/solvers/solve.rs:486 */

typedef struct lctp_ctx lctp_ctx;
typedef struct lctp_locus_h lctp_locus_h;

typedef struct lctp_device_cfg {
    int32_t device;            /* CUDA device ordinal */
    uint32_t flags;            /* reserved, 0 */
    void *stream;              /* cudaStream_t to launch on; NULL = the library creates its own */
    uint32_t max_resident_workers; /* 0 = auto (SM count x resident warps) -- scratch sizing only */
    uint32_t _pad;
} lctp_device_cfg;

/* Flat per-locus input = solve::Data (src/solvers/solve.rs:254-273) with everything the device never
 * touches removed; SURVEY.md Appendix C. */
typedef struct lctp_locus {
    uint32_t n_haps;               /* H: contigs of the locus */
    uint32_t n_reads;              /* R: reads (read pairs for paired-end), all_alns.reads().len() */
    uint32_t ploidy;               /* p */
    uint32_t is_paired;
    uint64_t n_genotypes;          /* G */
    const uint32_t *gt_tuples;     /* [G*p] Genotype::ids(); NULL = generate_genotypes without priors
                                      (src/command/genotype.rs:1123-1126, src/ext/vec.rs:298-339) */
    const double   *priors;        /* [G]; NULL = 0.0 */
    const double   *unmapped_prob; /* [R] GrouppedAlignments::unmapped_prob, src/model/locs.rs:600-602 */
    /* GrouppedAlignments::aln_pairs of every read, concatenated (read-major): per read sorted by
     * contig asc then ln_prob desc, <= 10 per (read, contig); src/model/locs.rs:669-735,793-798 */
    const uint64_t *pa_off;        /* [R+1] */
    const uint32_t *pa_contig;     /* [NPA] PairAlignment::contig_id */
    const double   *pa_ln_prob;    /* [NPA] PairAlignment::ln_prob */
    const uint32_t *pa_mid1;       /* [NPA] middle1(); LCTP_NONE_U32 = mate unmapped */
    const uint32_t *pa_mid2;       /* [NPA] middle2() */
    /* ContigInfo geometry, src/model/windows.rs:343-359,380-384 */
    const uint32_t *hap_len;       /* [H] contig_len */
    const uint32_t *hap_n_windows; /* [H] */
    const uint32_t *hap_reg_start; /* [H] window_getter.start */
    uint32_t window;               /* window size */
    uint32_t left_padding;
    const uint64_t *hap_pos_off;   /* [H+1] offsets into pos_weight / pos_gc (len_h = contig_len - neighb + 1) */
    const double   *pos_weight;    /* ContigInfo::neighb_info(start).1 for every mov_info index (windows.rs:439-445) */
    const uint8_t  *pos_gc;        /* NeighbInfo::gc_content */
    uint32_t depth_k;              /* columns of depth_table; must be >= 2R+1 */
    uint32_t tweak;                /* Params::tweak after set_tweak_size, src/model/mod.rs:179-197 */
    const double   *depth_table;   /* [101][depth_k]: DistrCache ln_pmf per GC bin (src/model/distr_cache.rs:61-75) */
    double prob_diff, lik_skew, min_weight, filt_diff, prob_thresh;   /* model::Params, src/model/mod.rs:64-106 */
    uint32_t dont_skip;
    uint32_t out_bams;
} lctp_locus;

/* One stage of the solving Scheme (src/solvers/solve.rs:138-202) with the solver's own parameters
 * (Greedy: src/solvers/stoch.rs:36-53; SimAnneal: :151-169). */
typedef struct lctp_stage {
    uint32_t kind;         /* 0 = "greedy", 1 = "anneal" */
    uint32_t attempts;     /* a */
    uint64_t in_size;      /* i */
    uint32_t best_start;   /* greedy x0: 1 = best, 0 = random */
    uint32_t _pad;
    uint64_t sample_size;  /* greedy s (device supports 1..=11: always Floyd sampling) */
    uint64_t plato_size;   /* greedy p / anneal p */
    uint64_t anneal_steps; /* anneal n */
    double   init_prob;    /* anneal P */
} lctp_stage;

/* Genotyping (src/solvers/solve.rs:568-590) without strings. */
typedef struct lctp_result {
    uint64_t n_out;
    uint64_t gt_ix[LCTP_MAX_OUT];
    double   lik_mean[LCTP_MAX_OUT];
    double   lik_var[LCTP_MAX_OUT];
    uint16_t attempts[LCTP_MAX_OUT];
    double   ln_prob[LCTP_MAX_OUT];
    double   quality;
    uint32_t total_reads;
    uint32_t unexpl_reads;
    uint32_t warn_no_probable;
    uint32_t warn_few_reads;
    uint64_t n_filtered;
    uint64_t n_stage_in[LCTP_MAX_STAGES];
    double   t_prefilter_s, t_stages_s;      /* host wall clock, diagnostics */
    /* Genotyping::find_weighted_dist (src/solvers/solve.rs:616-632); filled by lctp_find_weighted_dist only */
    uint32_t has_dist;            /* Genotyping::distances.is_some() */
    uint32_t true_edit_distances; /* Data::true_edit_distances -> "dist_type": "edit" | "minim-div" */
    uint32_t has_weight_dist;     /* Genotyping::weighted_dist.is_some() (every distance to the primary is known) */
    uint32_t _pad;
    double   weight_dist;
    uint32_t dist_to_primary[LCTP_MAX_OUT];   /* LCTP_NONE_U32 = "unknown" */
} lctp_result;

/* Device-side timing (CUDA events on the launch stream around each hot kernel) and work counters,
 * accumulated since the last reset; bench.py derives the roofline numbers from these. */
typedef struct lctp_stats {
    double   prefilter_ms;         /* sum of prefilter kernel durations */
    uint64_t prefilter_launches;
    uint64_t prefilter_genotypes;  /* genotypes scored */
    double   stage_ms;             /* sum of solver stage kernel durations */
    uint64_t stage_launches;
    uint64_t stage_genotypes;      /* genotypes solved */
    uint64_t stage_attempts;       /* genotype-attempts solved */
    uint64_t stage_iters;          /* greedy iterations + annealing steps executed */
    uint64_t stage_alns;           /* candidate locations built (sum of A) */
    double   pairing_ms;           /* sum of the two pairing kernel passes + scan (lctp_pair_alignments) */
    uint64_t pairing_launches;
    uint64_t pairing_mates;        /* mate alignment records read */
    uint64_t pairing_pairs;        /* pair alignments written */
    double   rescore_ms;           /* sum of the rescoring kernel durations (lctp_rescore_alignments) */
    uint64_t rescore_launches;
    uint64_t rescore_alns;         /* alignment records rescored */
    uint64_t rescore_ops;          /* CIGAR operations read */
    double   recruit_ms;           /* sum of the minimizer / recruitment kernel durations */
    uint64_t recruit_launches;
    uint64_t recruit_bases;        /* sequence bytes scanned */
    uint64_t recruit_reads;        /* reads (pairs) recruited against the targets */
    uint64_t h2d_bytes;            /* bytes copied host -> device by the library's calls on this context */
    uint64_t d2h_bytes;            /* bytes copied device -> host */
} lctp_stats;

/* Alignment records of one locus before pairing = the per-alignment part of PrelimAlignments::push
 * (src/model/locs.rs:297-313): count_region_operations_fast (src/seq/aln.rs:298-317, clipping limited to the contig by
 * limited_clipping :288-296, soft_clipping src/seq/cigar.rs:519-527), OperCounts::edit_distance
 * (src/bg/err_prof.rs:73-79), ErrorProfile::ln_prob (:212-221) and the `save` decision (locs.rs:308).
 * SURVEY 8(f) rank 2, first slice; the order-dependent de-duplication of alignment starts (PosCollection,
 * locs.rs:315-343) stays with the caller, which consumes `save` / `ln_prob` in record order. */
typedef struct lctp_alns {
    uint64_t n_alns;
    const uint64_t *cigar_off;     /* [n_alns+1] into cigar_ops; every alignment has >= 1 operation */
    const uint32_t *cigar_ops;     /* BAM encoding len << 4 | op; op 1=I 2=D 4=S 7='=' 8=X.  M, N, H, P are rejected like the
                                      reference's panic "Unsupported CIGAR operation" (extended CIGARs only; secondary
                                      records after Cigar::hard_to_soft, locs.rs:548) */
    const uint32_t *aln_start;     /* [n_alns] reference interval of the alignment on its contig */
    const uint32_t *aln_end;       /* [n_alns] */
    const uint32_t *contig_len;    /* [n_alns] ContigNames::get_len(aln.contig_id()) */
    const uint32_t *passable_dist; /* [n_alns] PrelimAlignments::passable_dist of the alignment's read end */
    double ln_match, ln_mismatch, ln_insertion, ln_deletion, ln_clipping;   /* ErrorProfile::oper_probs */
} lctp_alns;

/* Mate alignments of R read pairs = the input of identify_paired_end_alignments (src/model/locs.rs:805-868):
 * per read sorted by contig ascending, first read end before second, ln_prob descending within an end (the
 * order in which the reference consumes them after its sort at :819-820). */
typedef struct lctp_mates {
    uint32_t n_reads;              /* R */
    uint32_t n_haps;               /* H (contig ids are < H) */
    uint32_t max_alns;             /* MAX_USED_ALNS = 10 (locs.rs:741); <= 16 */
    uint32_t ins_len;              /* entries of ins_ln_pmf; must exceed every possible insert size */
    const uint64_t *ma_off;        /* [R+1] */
    const uint32_t *ma_contig;     /* [N] */
    const uint8_t  *ma_flags;      /* [N] bit0 = read end (0 first, 1 second), bit1 = strand */
    const uint32_t *ma_start;      /* [N] Interval::start */
    const uint32_t *ma_end;        /* [N] Interval::end (exclusive) */
    const double   *ma_ln_prob;    /* [N] Alignment::ln_prob */
    const double   *read_weight;   /* [R] ReadData::weight; NULL = 1.0 */
    const double   *ins_ln_pmf;    /* [ins_len] InsertDistr::ln_prob(size) (src/bg/insertsz.rs:153-155) */
    double unmapped_penalty;       /* Params::unmapped_penalty (src/model/mod.rs:55-60) */
    double insert_penalty;         /* InsertDistr::insert_penalty (insertsz.rs:171-173) */
    double prob_diff;              /* Params::prob_diff */
    /* single-end reads (identify_single_end_alignments, src/model/locs.rs:870-911): every record is a "first end" record,
     * per (read, contig) in descending ln_prob; ins_ln_pmf / insert_penalty are not used */
    uint32_t single_end;
    uint32_t window;               /* window size (ContigInfo::window_size), needed with explicit weights only */
    /* explicit region weights (ContigInfos::explicit_read_weight, src/model/windows.rs:683-693; ExplicitWeights::at,
     * :226-229): per contig the weight of every position, contig_len + 1 entries each (the reference appends the last
     * value once more); NULL = no explicit weights (weight 1) */
    const uint64_t *exp_off;       /* [H+1] */
    const double   *exp_weight;
    /* per-read limit instead of max_alns: MAX_USED_ALNS (10) for reads with weight >= Params::min_weight, MAX_UNUSED_ALNS
     * (2) for the others (recover_and_group_alignments, src/model/locs.rs:739-742, 1263); NULL = max_alns for every read */
    const uint8_t  *read_max_alns; /* [R], each 1..=16 */
} lctp_mates;

/* ---- library / context -------------------------------------------------------------------- */
const char *lctp_version(void);
const char *lctp_last_error(void);
size_t lctp_sizeof_locus(void);
size_t lctp_sizeof_stage(void);
size_t lctp_sizeof_result(void);
int  lctp_init(const lctp_device_cfg *cfg, lctp_ctx **out);
void lctp_destroy(lctp_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py `gpu_launches`). */
uint64_t lctp_launch_count(const lctp_ctx *ctx);
int  lctp_sync(lctp_ctx *ctx);
int  lctp_get_stats(lctp_ctx *ctx, lctp_stats *out, int reset);
/* Measurement aid (no reference counterpart): FP64-pipe instruction rate of the device in lane-instructions
 * per second, from a DADD microbenchmark run on the context's stream.  The prefilter (a2) issues two FP64-pipe
 * instructions per genotype-read, so rate / 2 is the denominator of its roofline. */
int  lctp_measure_fp64_rate(lctp_ctx *ctx, double *lane_inst_per_s);
/* Diagnostic (no reference counterpart, host only, no device needed): builds the work plan of the balanced
 * prefilter kernel for a diploid full-triangle locus of `n_haps` haplotypes on `n_sm` SMs and checks that every
 * genotype id of [g_begin, g_end) (the enumeration of src/ext/vec.rs:298-339) is owned by exactly one register
 * of one lane and that its staged matrix columns are the genotype's two haplotypes.  `pattern` = columns per
 * lane of the q-th warp of every SM sub-partition (n_pattern entries, each 2..8), or NULL / 0 to let the
 * planner choose; the chosen pattern (<= 4 entries), the number of CTA regions and the per-sub-partition load
 * (sum of the pattern) are returned when the pointers are non-NULL. */
int  lctp_prefilter_plan_check(uint32_t n_haps, uint32_t n_sm, const uint32_t *pattern, uint32_t n_pattern,
                               uint64_t g_begin, uint64_t g_end, uint32_t *n_regions, uint32_t *load,
                               uint32_t *pattern_out);

/* ---- locus upload (H2D once per locus; builds M = best_aln_matrix on the device, a1) ------- */
int  lctp_locus_upload(lctp_ctx *ctx, const lctp_locus *in, lctp_locus_h **out);
void lctp_locus_free(lctp_locus_h *h);
/* a1: AllAlignments::best_aln_matrix (src/model/locs.rs:1203-1212), [H][R] row-major by haplotype. */
int  lctp_best_aln_matrix(lctp_locus_h *h, double *m_out);

/* ---- 8(f) rank 1: mate alignments -> pair alignments -------------------------------------------- */
/* identify_paired_end_alignments + identify_contig_pair_alns (src/model/locs.rs:744-868) for every read on
 * the device.  Outputs are exactly the pa_* / unmapped_prob arrays of lctp_locus (caller-allocated; `cap`
 * entries in the four per-pair arrays, R+1 in pa_off, R in unmapped_prob); *n_out = pair alignments written.
 * LCTP_E_CAPACITY when cap is too small, LCTP_E_INVALID for records out of the documented order. */
int  lctp_pair_alignments(lctp_ctx *ctx, const lctp_mates *in, uint64_t cap, uint64_t *pa_off,
                          uint32_t *pa_contig, double *pa_ln_prob, uint32_t *pa_mid1, uint32_t *pa_mid2,
                          double *unmapped_prob, uint64_t *n_out);
size_t lctp_sizeof_mates(void);
/* Same, leaving the result on the device: the handle replaces the pa_* / unmapped_prob fields of lctp_locus in
 * lctp_locus_upload_pairs (no host round trip between pairing and the solver), lctp_pairs_fetch copies it to the host
 * when it is needed there (BAM output). */
typedef struct lctp_pairs_h lctp_pairs_h;
int  lctp_pair_alignments_dev(lctp_ctx *ctx, const lctp_mates *in, lctp_pairs_h **out, uint64_t *n_out);
int  lctp_pairs_fetch(lctp_pairs_h *p, uint64_t cap, uint64_t *pa_off, uint32_t *pa_contig, double *pa_ln_prob,
                      uint32_t *pa_mid1, uint32_t *pa_mid2, double *unmapped_prob);
uint64_t lctp_pairs_count(const lctp_pairs_h *p);
void lctp_pairs_free(lctp_pairs_h *p);
int  lctp_locus_upload_pairs(lctp_ctx *ctx, const lctp_locus *in /* pa_*, unmapped_prob ignored */,
                             lctp_pairs_h *pairs, lctp_locus_h **out);
/* SURVEY 8(f) rank 2, first slice: rescore every alignment record under the error profile on the device.
 * Outputs (caller-allocated, n_alns entries each): ln_prob = Alignment::set_ln_prob value, edit / read_len =
 * EditDist{edit, read_len}, save = (edit <= passable_dist).  LCTP_E_INVALID on an empty CIGAR or an unsupported
 * operation (the reference panics). */
int  lctp_rescore_alignments(lctp_ctx *ctx, const lctp_alns *in, double *ln_prob, uint32_t *edit,
                             uint32_t *read_len, uint8_t *save);
size_t lctp_sizeof_alns(void);
/* SURVEY 8(f) rank 2, second slice: the per read-end protocol of read_next_alns (src/model/locs.rs:502-567) around
 * push, with the PosCollection de-duplication of alignment starts (:166-187, 315-343), for every (read, read end) of a
 * locus at once.  Records are grouped by (read, read end), first record of a group = the primary alignment (unmapped
 * read ends are simply absent); secondary records come after Cigar::hard_to_soft, empty CIGARs dropped (:548-554). */
typedef struct lctp_read_ends {
    lctp_alns alns;                    /* passable_dist is ignored (derived per group below) */
    uint64_t n_groups;
    const uint64_t *grp_off;           /* [n_groups+1] */
    const uint32_t *rec_contig;        /* [n_alns] contig id of the record */
    const uint8_t  *grp_read_end;      /* [n_groups] 0 / 1 */
    const uint32_t *grp_read_len;      /* record.seq().len() of the primary record */
    const uint32_t *grp_good_dist;     /* EditDistCache::get(read_len).0 */
    const uint32_t *grp_passable_dist; /* EditDistCache::get(read_len).1 */
    const double   *grp_neighb_complexity;   /* ContigInfos::neighb_complexity(&primary); 1.0 for long reads (:527) */
    double poor_compl, poor_compl_edit;      /* Params::poor_compl / poor_compl_edit */
    uint32_t strict_subset;                  /* Data::strict_subset */
    uint32_t _pad;
} lctp_read_ends;
/* Outputs, caller-allocated.  Per record [n_alns]: ln_prob, edit, read_len as in lctp_rescore_alignments.  Per group
 * [n_groups]: ok = read_next_alns returned true; best_edit = PrelimAlignments::best_edit of the read end; weight_factor =
 * what read_data.weight is multiplied by (:564; 1.0 when !ok); thr_dist / pass_dist = the thresholds given to
 * set_thresholds (:536); n_kept = alignments left in PrelimAlignments::alns for this read end, kept_rec[grp_off[g] + k]
 * = record index of the k-th of them (the order of `alns`: a better alignment of an occupied 128-bp bin replaces the
 * earlier one in place).  A group whose primary alignment is not good enough (:538-542) has ok = 0 and n_kept = 0. */
int  lctp_collect_read_ends(lctp_ctx *ctx, const lctp_read_ends *in, double *ln_prob, uint32_t *edit,
                            uint32_t *read_len, uint8_t *ok, uint32_t *best_edit, double *weight_factor,
                            uint32_t *thr_dist, uint32_t *pass_dist, uint32_t *n_kept, uint32_t *kept_rec);
size_t lctp_sizeof_read_ends(void);

/* SURVEY 8(f) rank 1, remainder: from the per read-end results to the pairing input, for every read of a locus at once.
 * AllAlignments::load after read_next_alns (src/model/locs.rs:1117-1137): a read is well mapped when read_next_alns
 * returned true for its read ends (the second end is only read when the first was fine), and in bounds when some
 * alignment of PrelimAlignments::alns has its middle outside the boundary regions (in_bounds, :1008-1014).  Then
 * recover_and_group_alignments (:1237-1288) WITHOUT the alignment transfer (opt_hap_alns = None; the transfer needs
 * WFA2, an un-vendored submodule): best_edit_is_good (:293-295), normalize_probs (:358-360: ln_prob -= the best
 * ln_prob over every pushed alignment of the read end), max_alns = 10 / 2 by the read weight (:1263), and the order in
 * which identify_paired_end_alignments / identify_single_end_alignments consume the alignments after their sorts
 * (:819-820, 883): contig ascending, first read end before second, ln_prob descending.  The reference's sorts are
 * unstable; equal keys keep the order of PrelimAlignments::alns here. */
typedef struct lctp_prelim {
    uint64_t n_reads, n_groups;
    const int64_t  *read_group;        /* [n_reads][2] group (of the lctp_collect_read_ends call) of the first / second
                                          read end, -1 = none (unmapped / not read; single-end: second always -1) */
    const uint64_t *grp_off;           /* [n_groups+1] records of a group, as in lctp_read_ends */
    const uint32_t *rec_contig;        /* [n_alns] */
    const uint32_t *rec_start;         /* [n_alns] Interval::start */
    const uint32_t *rec_end;           /* [n_alns] Interval::end (exclusive) */
    const uint8_t  *rec_strand;        /* [n_alns] 1 = reverse */
    const double   *rec_ln_prob;       /* [n_alns] ln_prob output of lctp_collect_read_ends */
    const uint8_t  *grp_ok;            /* [n_groups] outputs of lctp_collect_read_ends: ok, */
    const uint32_t *grp_best_edit;     /*   best_edit, */
    const uint32_t *grp_thr_dist;      /*   thr_dist (= PrelimAlignments::good_dist after set_thresholds), */
    const uint32_t *grp_n_kept;        /*   n_kept, */
    const uint32_t *kept_rec;          /*   kept_rec [n_alns] */
    const uint32_t *contig_len;        /* [n_haps] */
    const double   *read_weight;       /* [n_reads] ReadData::weight after calculate_read_weight (:1144-1148) */
    double min_weight;                 /* Params::min_weight */
    uint32_t n_haps;
    uint32_t boundary;                 /* params.boundary_size - params.tweak (:1099) */
    uint32_t single_end;
    uint32_t _pad;
} lctp_prelim;
/* Outputs, caller-allocated.  status[n_reads]: 0 = goes on to the pairing, 1 = not well mapped (ReadCounts::
 * poorly_mapped of load), 2 = out of bounds, 3 = best edit distance not good (poorly_mapped of
 * recover_and_group_alignments).  The reads with status 0 are numbered 0 .. *n_reads_out in read order: out_read[k] =
 * the read, out_max_alns[k] = its max_alns (lctp_mates::read_max_alns), ma_off[k] .. ma_off[k+1] = its entries of
 * ma_contig / ma_flags / ma_start / ma_end / ma_ln_prob (the lctp_mates arrays of the same names; `cap` entries each)
 * and ma_rec (the record each entry came from).  counts[3] = poorly mapped (status 1 and 3), out of bounds, passed.
 * LCTP_E_CAPACITY when cap is too small (the number of kept records always suffices).  Preconditions (the outputs of
 * lctp_collect_read_ends satisfy them): grp_n_kept[g] <= grp_off[g+1] - grp_off[g], the first grp_n_kept[g] entries of
 * kept_rec at grp_off[g] are record indices < n_alns, rec_contig < n_haps; read_group entries are -1 or < n_groups
 * (checked: LCTP_E_INVALID). */
int  lctp_group_reads(lctp_ctx *ctx, const lctp_prelim *in, uint64_t cap, uint8_t *status, uint64_t *n_reads_out,
                      uint32_t *out_read, uint8_t *out_max_alns, uint64_t *ma_off, uint32_t *ma_contig,
                      uint8_t *ma_flags, uint32_t *ma_start, uint32_t *ma_end, double *ma_ln_prob, uint32_t *ma_rec,
                      uint64_t *counts);
size_t lctp_sizeof_prelim(void);
/* Same, leaving the pairing input on the device: the handle replaces the ma_* / read_max_alns fields of lctp_mates in
 * lctp_pair_alignments_from, so that the (largest) arrays of the upstream chain cross PCIe once, as records, and never
 * again (group -> pair -> locus upload all read device memory).  Only status, out_read (may be NULL), the counts and the
 * number of passing reads come back.  `params` of lctp_pair_alignments_from = an lctp_mates whose ma_* / read_max_alns /
 * n_reads are ignored: n_haps, max_alns, the insert-size table, penalties, prob_diff, single_end, window, explicit
 * weights, and read_weight = the weights of the PASSING reads in out_read order (or NULL = 1.0). */
typedef struct lctp_mates_h lctp_mates_h;
int  lctp_group_reads_dev(lctp_ctx *ctx, const lctp_prelim *in, uint8_t *status, uint64_t *n_reads_out,
                          uint32_t *out_read, uint64_t *counts, lctp_mates_h **out);
uint64_t lctp_mates_count(const lctp_mates_h *m);
void lctp_mates_free(lctp_mates_h *m);
int  lctp_pair_alignments_from(lctp_ctx *ctx, const lctp_mates_h *mates, const lctp_mates *params, lctp_pairs_h **out,
                               uint64_t *n_out);

/* Read weights from the k-mers unique to the locus: UniqueKmers (src/model/locs.rs:915-1003), the step of
 * AllAlignments::load between read_next_alns and recover_and_group_alignments (:1144-1148).
 * lctp_unique_kmers_build = UniqueKmers::new (:930-963): the canonical k-mers (kmers::kmers::<u128, _, CANONICAL>,
 * src/seq/kmers.rs:163-202; k = KmerCounts::k(), 2..=63) of the n_seqs contig sequences (ASCII, anything but A/C/G/T is
 * an N) whose off-target count (kmer_counts, len + 1 - k entries per sequence at cnt_off) is 0; host pass + upload of an
 * open-addressing table.  hard / soft threshold = Params::kmer_hard_thresh / kmer_soft_thresh.
 * lctp_read_weights = calculate_read_weight (:968-1002) for n_reads reads of `ends` (1 or 2) read ends each on the
 * device: sequence e of read r = seqs[seq_off[r * ends + e] .. seq_off[r * ends + e + 1]), empty = no mate;
 * unique[r * ends + e] = MateData::unique_kmers (non-overlapping unique k-mers of the read end), weight[r] = the factor
 * read_data.weight is multiplied by, clamp(intercept + count * slope, 0, 1). */
typedef struct lctp_unique_kmers_h lctp_unique_kmers_h;
int  lctp_unique_kmers_build(lctp_ctx *ctx, const uint8_t *seqs, const uint64_t *seq_off, uint64_t n_seqs,
                             const uint16_t *kmer_counts, const uint64_t *cnt_off, uint32_t k, uint16_t hard_threshold,
                             uint16_t soft_threshold, lctp_unique_kmers_h **out, uint64_t *n_unique);
uint64_t lctp_unique_kmers_count(const lctp_unique_kmers_h *u);
void lctp_unique_kmers_free(lctp_unique_kmers_h *u);
int  lctp_read_weights(lctp_ctx *ctx, const lctp_unique_kmers_h *u, const uint8_t *seqs, const uint64_t *seq_off,
                       uint64_t n_reads, uint32_t ends, uint16_t *unique, double *weight);

/* ---- SURVEY 8(f) rank 3, first slice: short-read recruitment ---------------------------------------------------------
 * Canonical minimizers (kmers::minimizers::<u64, _, CANONICAL>, src/seq/kmers.rs:71-103, 256-340) of n sequences
 * (ASCII A/C/G/T, anything else is an N) on the device.  Outputs of sequence s start at off[s] in hash / pos / fw (a
 * sequence has fewer minimizers than bases) and count[s] says how many there are: hash = fast_hash of the canonical
 * k-mer (what the reference stores as "the minimizer"), pos = start of the k-mer, fw = 1 if the forward k-mer was the
 * canonical one. */
int  lctp_minimizers(lctp_ctx *ctx, const uint8_t *seqs, const uint64_t *off, uint64_t n, uint32_t k, uint32_t w,
                     uint32_t *count, uint64_t *hash, uint32_t *pos, uint8_t *fw);
/* Fraction::<u16>::approximate (src/math/frac.rs:48-80): Params::match_frac_short */
void lctp_fraction_approximate_u16(double x, uint16_t *num, uint16_t *den);
/* Recruitment targets = TargetBuilder::add for every locus + finalize (src/seq/recruit.rs:680-760). */
typedef struct lctp_target_seqs {
    uint64_t n_seqs;
    const uint64_t *seq_off;        /* [n_seqs+1] into seqs */
    const uint8_t  *seqs;
    const uint32_t *seq_locus;      /* [n_seqs] locus index of the sequence, ascending (add() is called locus by locus) */
    const uint64_t *cnt_off;        /* [n_seqs+1] into kmer_counts */
    const uint16_t *kmer_counts;    /* KmerCounts of the sequence: len + 1 - base_k entries (the reference asserts) */
    uint32_t base_k;                /* KmerCounts::k() */
    uint32_t minimizer_k, minimizer_w;   /* Params::minimizer_k / minimizer_w */
    uint32_t thresh_kmer_count;     /* Params::thresh_kmer_count */
    double match_frac;              /* Params::match_frac */
    uint32_t match_length;          /* Params::match_length (long reads: stretch_minims / stretch_score, recruit.rs:95-98) */
    uint32_t _pad;
} lctp_target_seqs;
typedef struct lctp_targets_h lctp_targets_h;
int  lctp_targets_build(lctp_ctx *ctx, const lctp_target_seqs *in, lctp_targets_h **out);
void lctp_targets_free(lctp_targets_h *t);
/* minim_to_loci flattened in insertion order: (minimizer, locus, info = direction | rare << 2); returns the entry count */
uint64_t lctp_targets_entries(const lctp_targets_h *t, uint64_t *key, uint32_t *locus, uint8_t *info, uint64_t cap);
void lctp_targets_match_frac(const lctp_targets_h *t, uint16_t *num, uint16_t *den);
/* Reads: first mates (or single-end reads) and, when off2 / seq2 are given, the second mates of the same pairs. */
typedef struct lctp_reads {
    uint64_t n_reads;
    const uint64_t *off1; const uint8_t *seq1;
    const uint64_t *off2; const uint8_t *seq2;     /* NULL = single-end */
} lctp_reads;
/* RecruitableRecord::recruit (src/seq/recruit.rs:582-611) for every read (pair): single-end reads of at most 500 bp
 * (READ_LENGTH_THRESH) -> recruit_short_read (:852-881), longer ones -> recruit_long_read with has_matching_stretch
 * (:932-998), pairs -> recruit_read_pair (:885-930).  ans_count[r] loci, ans_locus[r * cap ..] their indices in ascending
 * order (the reference's answer is a set).  LCTP_E_CAPACITY when a read matches more than 8 loci; ans_count[r] > cap
 * means the list of that read was cut at cap. */
int  lctp_recruit_short(lctp_ctx *ctx, const lctp_targets_h *t, const lctp_reads *reads, uint32_t cap,
                        uint32_t *ans_count, uint32_t *ans_locus);
size_t lctp_sizeof_target_seqs(void);
size_t lctp_sizeof_reads(void);

/* ---- prefilter (a2 + a3) ------------------------------------------------------------------ */
/* Scores of genotypes [g_begin, g_end) (src/solvers/solve.rs:105-119) computed on the device into the
 * handle's score buffer; copied to scores_out[g_end - g_begin] when not NULL. */
int  lctp_prefilter_scores(lctp_locus_h *h, uint64_t g_begin, uint64_t g_end, double *scores_out);
/* run_filter (src/solvers/solve.rs:87-122): ixs holds n genotype ids in/out; *out_n survivors, sorted. */
int  lctp_prefilter(lctp_locus_h *h, uint64_t *ixs, size_t n, size_t min_size, size_t threads,
                    size_t *out_n, double *scores_out /* nullable, [G] */);
/* truncate_ixs (src/solvers/solve.rs:52-84) on host arrays (used by the multi-GPU merge). */
size_t lctp_truncate_ixs(uint64_t *ixs, size_t n, const double *scores, double filt_diff,
                         size_t min_size, size_t threads);

/* ---- one solver stage over explicit logical workers (a5-a14) ------------------------------- */
/* worker w solves genotypes worker_ixs[worker_off[w] .. worker_off[w+1]) back to back on its own
 * xoshiro256++ stream worker_rng[4*w .. 4*w+4) (in/out), exactly like Worker::run
 * (src/solvers/solve.rs:1104-1145).  Outputs are indexed by position j in worker_ixs. */
int  lctp_solve_stage(lctp_locus_h *h, const lctp_stage *st,
                      const uint64_t *worker_ixs, const uint64_t *worker_off, size_t n_workers,
                      uint64_t *worker_rng,
                      double *lik_mean, double *lik_var,
                      double *liks /* nullable [n*attempts] */,
                      uint64_t *counts_off /* nullable [n+1] */, uint16_t *counts /* nullable */,
                      uint64_t counts_cap,
                      uint64_t *n_alns_out /* nullable [n] */, uint64_t *iters_out /* nullable [n] */);

/* ---- output side, SURVEY 8(f) rank 4: what `--debug` of the reference writes per locus ------------------------- */
/* Per (genotype position j, attempt a) of a stage, index j * attempts + a: the fields of ReadAssignment::summarize
 * (src/model/assgn.rs:413-425) and, when the win_* pointers are given, of ReadAssignment::write_depth (:356-372) for
 * every window w of the genotype (index (j * attempts + a) * wmax + w; windows 0 / 1 are the "unmapped" / "out of
 * bounds" windows, contig i owns windows [wshift_i, wshift_{i+1}) with wshift_0 = 2).  All pointers are
 * caller-allocated host arrays; NULL = not wanted. */
typedef struct lctp_stage_debug {
    double   *aln_lik;       /* [n * attempts] ReadAssignment::aln_lik (ln) */
    double   *depth_lik;     /* [n * attempts] ReadAssignment::depth_lik (ln) */
    uint32_t *unmapped;      /* [n * attempts] depth[UNMAPPED_WINDOW] */
    uint32_t *out_of_bounds; /* [n * attempts] depth[BOUNDARY_WINDOW] */
    uint32_t  wmax;          /* stride of the win_* arrays: >= 2 + ploidy * max windows per contig (lctp_locus_wmax) */
    uint32_t  _pad;
    double   *win_weight;    /* WindowDistr::weight() */
    uint32_t *win_depth;     /* ReadAssignment::depth[w] */
    double   *win_lik;       /* WindowDistr::ln_prob(depth[w]) (ln) */
} lctp_stage_debug;
uint32_t lctp_locus_wmax(const lctp_locus_h *h);
/* lctp_solve_stage + the debug fields above (dbg may be NULL). */
int  lctp_solve_stage_dbg(lctp_locus_h *h, const lctp_stage *st,
                          const uint64_t *worker_ixs, const uint64_t *worker_off, size_t n_workers,
                          uint64_t *worker_rng, double *lik_mean, double *lik_var, double *liks,
                          uint64_t *counts_off, uint16_t *counts, uint64_t counts_cap,
                          uint64_t *n_alns_out, uint64_t *iters_out, const lctp_stage_debug *dbg);
/* Debug sink of a context: while open, lctp_solve writes the reference's per-locus debug tables into `dir`
 * (src/solvers/solve.rs:852-952): level >= 1 (DebugLvl::Some): sol.csv ("stage genotype score": stage 0 = prefilter
 * score of every genotype :115-117, stages 1.. = lik_mean :1074-1075) and sol_ext.csv (summarize rows); level >= 2
 * (DebugLvl::Full): depth.csv (write_depth rows).  Plain text, uncompressed: the reference wraps the same rows in a
 * brotli writer (.csv.br), which stays on the Rust side.  Rows of a stage are written in dispatch order (the
 * reference's order inside a stage depends on thread timing).  hap_names: n_haps NUL-terminated contig names
 * (Genotype display = names joined by ','); they are copied. */
int  lctp_debug_open(lctp_ctx *ctx, const char *dir, int level, const char *const *hap_names, size_t n_haps);
void lctp_debug_close(lctp_ctx *ctx);
/* lctp_solve that also returns the assignment counts (Prediction::assgn_counts, src/solvers/solve.rs:286-317, the
 * input of write_bam, src/model/bam.rs:356-413) of the first `n_counts` genotypes of the result (the reference writes
 * BAMs for `out_bams` of them): counts_off[k] .. counts_off[k+1] = the candidates of res->gt_ix[k] in
 * GenotypeAlignments order, counts[c] = attempts in which candidate c was the read's location. */
int  lctp_solve_counts(lctp_locus_h *h, const lctp_stage *stages, size_t n_stages, size_t threads,
                       uint64_t rng[4], lctp_result *res, size_t n_counts,
                       uint64_t *counts_off /* [n_counts + 1] */, uint16_t *counts, uint64_t counts_cap);

/* count_to_prob (src/model/bam.rs:54-66): what write_bam turns an assignment count into, for n counts at once (host):
 * prob (f32, the `pr` tag) and mapq.  count 0 -> (0, 0), count == attempts -> (1, 60), otherwise prob = count / attempts
 * in f32 and mapq = round(-10 log10(1 - prob)) capped at 60.  LCTP_E_INVALID when a count exceeds `attempts` (the
 * reference asserts). */
int  lctp_counts_to_prob(const uint16_t *counts, uint64_t n, uint16_t attempts, float *prob, uint8_t *mapq);

/* ---- host-side mirror of the scheduler (a14-a16) ------------------------------------------- */
void lctp_rng_seed_from_u64(uint64_t state[4], uint64_t seed);   /* src/ext/rand.rs:12 */
void lctp_rng_jump(uint64_t state[4]);                            /* src/solvers/solve.rs:1017 */
/* MainWorker::new (src/solvers/solve.rs:1007-1018): worker w gets a clone of the locus stream, which then jumps;
 * out[4*w..4*w+4] = state of worker w, `state` ends `threads` jumps ahead. */
void lctp_rng_worker_streams(uint64_t state[4], size_t threads, uint64_t *out);
void lctp_rng_long_jump(uint64_t state[4]);                       /* src/command/genotype.rs:1345 */
/* MainWorker::run :1049-1063: shuffle ixs with the locus stream and cut into <= threads chunks.
 * worker_off has threads+1 entries; returns the number of workers that received work. */
size_t lctp_plan_stage(uint64_t rng[4], uint64_t *ixs, size_t n, size_t threads, uint64_t *worker_off);
/* discard_improbable_genotypes :425-480; per-genotype arrays are indexed by genotype id. */
size_t lctp_discard_improbable(uint64_t *ixs, size_t n, const double *lik_mean, const double *lik_var,
                               const uint16_t *attempts, double prob_thresh, size_t out_size, size_t threads);
double lctp_compare_two_likelihoods(double m1, double v1, uint16_t a1, double m2, double v2, uint16_t a2);
/* DistrCache::new (src/model/distr_cache.rs:61-75): out[101][k_cols] from the preproc NB(n, p) per GC bin. */
/* n_alt must be < 16 (BayesCalc::new asserts alternatives.len() < N_ALTS, src/math/distr/bayes.rs:16): LCTP_E_INVALID. */
int  lctp_build_depth_table(const double *nb_n, const double *nb_p, int is_paired,
                            const double *alt_cn, size_t n_alt, uint32_t k_cols, double *out);

/* Predictions::produce_result (src/solvers/solve.rs:482-535) followed by check_first_prob (:637-645),
 * check_num_of_reads (:649-678) and count_unexplained_reads (:719-729).  ixs holds the n genotype ids that
 * survived the last stage (reordered in place); the per-genotype arrays are indexed by genotype id.
 * Fills every field of `res` except n_filtered / n_stage_in / t_*. */
int  lctp_produce_result(lctp_locus_h *h, uint64_t *ixs, size_t n, const double *lik_mean,
                         const double *lik_var, const uint16_t *attempts, lctp_result *res);

/* Genotyping::find_weighted_dist (src/solvers/solve.rs:616-632) with genotype_distance (:339-357, permutations
 * in the order of ext::vec::gen_permutations, src/ext/vec.rs:342-372).  Host only.  `dist` is the linear storage of
 * Data::contig_distances (TriangleMatrix<Option<u32>>, src/ext/trimat.rs:8-46: pairs i < j, row-major,
 * n_haps * (n_haps - 1) / 2 entries, LCTP_NONE_U32 = None); `loc` supplies n_haps, ploidy and gt_tuples.
 * The reference calls it after produce_result when the locus has a distance matrix (solve.rs:973-975). */
int  lctp_find_weighted_dist(lctp_result *res, const lctp_locus *loc, const uint32_t *dist, int true_edit_distances);

/* solve::solve (src/solvers/solve.rs:926-981) without file output: prefilter -> stages -> result.
 * `rng` is the locus stream (in/out); `threads` is the reference's -@ (number of logical workers). */
int  lctp_solve(lctp_locus_h *h, const lctp_stage *stages, size_t n_stages, size_t threads,
                uint64_t rng[4], lctp_result *res);
/* ---- one locus sharded over the GPUs of a box ----------------------------------------
 * One process (or thread) per GPU, each with its own lctp_ctx and the same locus uploaded.  Genotypes are independent
 * units of both phases (src/solvers/solve.rs:105-119, :1116-1142): the prefilter is split by contiguous genotype id
 * ranges, the stages by logical workers (w mod world); the two exchanges are ncclAllGather calls over NVLink, issued
 * by the library on the context's stream (NCCL is dlopen'ed at lctp_dist_init: single-GPU users do not need it).
 * Results are identical on every rank and identical to the single-GPU calls. */
#define LCTP_DIST_ID_BYTES 128   /* sizeof(ncclUniqueId) */
typedef struct lctp_dist lctp_dist;
typedef struct lctp_dist_timing {        /* accumulated since the last reset, this rank */
    double   kernel_ms;                  /* prefilter + stage kernels (CUDA events) */
    double   collective_ms;              /* ncclAllGather calls (CUDA events on the launch stream) */
    double   host_ms;                    /* lctp_dist_solve wall time minus the two above */
    double   wall_ms;                    /* lctp_dist_solve wall time */
    uint64_t collectives, gathered_bytes, overflow_rounds, solves;
} lctp_dist_timing;
/* ncclGetUniqueId: call on one rank, hand the 128 bytes to the others out of band (MPI, torch.distributed, a file). */
int  lctp_dist_unique_id(uint8_t id[LCTP_DIST_ID_BYTES]);
/* ncclCommInitRank on the context's device; collective over all `world` ranks. */
int  lctp_dist_init(lctp_ctx *ctx, const uint8_t id[LCTP_DIST_ID_BYTES], int rank, int world, lctp_dist **out);
void lctp_dist_destroy(lctp_dist *d);
int  lctp_dist_rank(const lctp_dist *d);
int  lctp_dist_world(const lctp_dist *d);
int  lctp_dist_get_timing(lctp_dist *d, lctp_dist_timing *out, int reset);
/* run_filter (src/solvers/solve.rs:87-122) over all G genotypes, sharded: this rank scores its id range, selects on
 * the device the candidates that can survive truncate_ixs (:52-84) -- everything >= min(local best - filt_diff,
 * K-th best local), K = max(min_size, threads) -- and the ranks all-gather fixed-capacity buffers of (score f64,
 * id u64) + {count, overflow}; a second, larger exchange only when some rank overflowed.  ixs_out (capacity cap_out,
 * world * max(min_size, threads) is always enough when nothing ties) receives the sorted survivors. */
int  lctp_dist_prefilter(lctp_dist *d, lctp_locus_h *h, size_t min_size, size_t threads, uint64_t *ixs_out,
                         size_t cap_out, size_t *n_out);
/* lctp_solve_stage, sharded: same arguments on every rank (the caller shuffles / partitions with the shared locus
 * stream, e.g. lctp_plan_stage); this rank solves workers w = rank (mod world); lik_mean / lik_var / worker_rng come
 * back complete on every rank. */
int  lctp_dist_solve_stage(lctp_dist *d, lctp_locus_h *h, const lctp_stage *st, const uint64_t *worker_ixs,
                           const uint64_t *worker_off, size_t n_workers, uint64_t *worker_rng, double *lik_mean,
                           double *lik_var);
/* lctp_solve, sharded (collective; every rank passes the same arguments and receives the same result and rng). */
int  lctp_dist_solve(lctp_dist *d, lctp_locus_h *h, const lctp_stage *stages, size_t n_stages, size_t threads,
                     uint64_t rng[4], lctp_result *res);

/* Genotyping::to_json (src/solvers/solve.rs:732-773), pretty-printed with indent 4 like
 * src/command/genotype.rs:1256.  hap_names[H]; returns bytes needed (excluding NUL). */
size_t lctp_result_json(const lctp_result *res, const lctp_locus *loc, const char *const *hap_names,
                        char *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif

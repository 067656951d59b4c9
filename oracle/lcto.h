/*
 * lcto.h -- CPU ORACLE for the locityper genotype-evaluation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (tprodanov/locityper v1.7.2, Rust) for the path
 *     solve::solve() = prefilter + staged read-assignment solvers + depth likelihood.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
 * legs may load it, and only as the checker / CPU baseline.  The product
 * (locityper_b200/, include/lctp.h) never links, imports or executes anything here.
 *
 * PARITY STATUS: "parity unpinned" by the reference's own tests -- the reference
 * ships no tests, golden vectors or fixtures for this path and
 * it cannot be compiled here (no cargo/rustc).  What IS pinned:
 *   - xoshiro256++ / SplitMix64 against the public-domain known-answer vectors;
 *   - ln_gamma / Student-t cdf against scipy fixtures (tests/golden/);
 *   - RNG-free sub-results against brute-force recomputation (tests/).
 * Third-party arithmetic restated from the published algorithms (not vendored in
 * /root/reference): rand 0.10 (uniform ints, Floyd sample, shuffle, f64),
 * rand_xoshiro 0.8 (seed_from_u64, jump, long_jump), statrs 0.19 (ln_gamma,
 * StudentsT::cdf), Rust std sort_unstable_by (insertion sort for len <= 20).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#ifndef LCTO_H
#define LCTO_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCTO_NONE_U32 0xFFFFFFFFu
#define LCTO_GC_BINS 101          /* src/bg/depth.rs:42 */
#define LCTO_MAX_PLOIDY 16

/* ------------------------------------------------------------------ RNG (lcto_rng.c) */

typedef struct { uint64_t s[4]; } lcto_rng;  /* rand_xoshiro::Xoshiro256PlusPlus, src/ext/rand.rs:3 */

void     lcto_rng_seed_from_u64(lcto_rng *r, uint64_t seed);
uint64_t lcto_rng_next_u64(lcto_rng *r);
uint32_t lcto_rng_next_u32(lcto_rng *r);
void     lcto_rng_jump(lcto_rng *r);
void     lcto_rng_long_jump(lcto_rng *r);
double   lcto_rng_f64(lcto_rng *r);
/* rand::RngExt::random_range for the integer types used on the path. */
uint32_t lcto_rng_range_u32_incl(lcto_rng *r, uint32_t low, uint32_t high);
uint64_t lcto_rng_range_u64_incl(lcto_rng *r, uint64_t low, uint64_t high);
int32_t  lcto_rng_range_i32_incl(lcto_rng *r, int32_t low, int32_t high);
size_t   lcto_rng_range_usize(lcto_rng *r, size_t low, size_t high_excl);
uint16_t lcto_rng_range_u16(lcto_rng *r, uint16_t low, uint16_t high_excl);
/* rand::seq::index::sample (Floyd for amount < 12). out has `amount` entries. */
int      lcto_rng_sample_indices(lcto_rng *r, uint32_t length, uint32_t amount, uint32_t *out);
/* rand::seq::SliceRandom::shuffle on a usize slice. */
void     lcto_rng_shuffle_usize(lcto_rng *r, size_t *v, size_t n);

/* ------------------------------------------------------------ special functions (lcto_specfun.c) */

double lcto_ln_gamma(double x);                         /* statrs::function::gamma::ln_gamma */
double lcto_beta_reg(double a, double b, double x);     /* statrs::function::beta::beta_reg */
double lcto_students_t_cdf(double x, double freedom);   /* statrs StudentsT(0,1,freedom).cdf */
double lcto_ln_add(double a, double b);                 /* src/math/mod.rs:28-34 */
double lcto_ln_sum(const double *v, size_t n);          /* src/math/mod.rs:50-75 */
double lcto_ln_sum_init(const double *v, size_t n, double init); /* src/math/mod.rs:56-94 */
/* Bayesian NB depth table, src/model/distr_cache.rs:61-75: out[101][k_cols]. */
void   lcto_build_depth_table(const double *nb_n, const double *nb_p, int is_paired,
                              const double *alt_cn, size_t n_alt, uint32_t k_cols, double *out);

/* ---------------------------------------------------------------- flat locus (Appendix C) */

typedef struct lcto_locus {
    uint32_t n_haps;          /* H */
    uint32_t n_reads;         /* R (read pairs for paired-end) */
    uint32_t ploidy;          /* p */
    uint32_t is_paired;
    uint64_t n_genotypes;     /* G */
    const uint32_t *gt_tuples;     /* [G*p]; NULL = all combinations with replacement in reference order */
    const double   *priors;        /* [G]; NULL = 0.0 */
    const double   *unmapped_prob; /* [R] */
    /* read-major CSR of PairAlignment (src/model/locs.rs:669-676): per read sorted by contig asc,
     * then ln_prob desc; at most 10 per (read, contig). */
    const uint64_t *pa_off;        /* [R+1] */
    const uint32_t *pa_contig;     /* [NPA] */
    const double   *pa_ln_prob;    /* [NPA] */
    const uint32_t *pa_mid1;       /* [NPA]; LCTO_NONE_U32 = mate unmapped */
    const uint32_t *pa_mid2;       /* [NPA] */
    /* haplotype geometry (src/model/windows.rs:343-359, 380-384) */
    const uint32_t *hap_len;       /* [H] contig_len */
    const uint32_t *hap_n_windows; /* [H] */
    const uint32_t *hap_reg_start; /* [H] */
    uint32_t window;
    uint32_t left_padding;
    const uint64_t *hap_pos_off;   /* [H+1] offsets into pos_weight / pos_gc */
    const double   *pos_weight;    /* neighb_info weight per window start (windows.rs:439-445) */
    const uint8_t  *pos_gc;
    uint32_t depth_k;              /* columns of depth_table */
    uint32_t tweak;
    const double   *depth_table;   /* [101][depth_k]: ln_pmf of the Bayes NB distribution */
    double prob_diff, lik_skew, min_weight, filt_diff, prob_thresh;
    uint32_t dont_skip;
    uint32_t out_bams;
} lcto_locus;

size_t lcto_sizeof_locus(void);

typedef struct lcto_stage {
    uint32_t kind;        /* 0 = greedy, 1 = anneal */
    uint32_t attempts;    /* u16 in the reference */
    uint64_t in_size;
    uint32_t best_start;  /* greedy x0 */
    uint32_t _pad;
    uint64_t sample_size; /* greedy s */
    uint64_t plato_size;  /* greedy p / anneal p */
    uint64_t anneal_steps;/* anneal n */
    double   init_prob;   /* anneal P */
} lcto_stage;

size_t lcto_sizeof_stage(void);

/* ---------------------------------------------------------------- hot path (lcto_model.c, lcto_solve.c) */

/* Genotype tuple of genotype `g` (explicit list or reference enumeration order). */
void lcto_genotype_tuple(const lcto_locus *L, uint64_t g, uint32_t *out_p);

/* a1: src/model/locs.rs:1203-1212.  M is [H][R] row-major by haplotype. */
void lcto_best_aln_matrix(const lcto_locus *L, double *M);

/* a2: src/solvers/solve.rs:105-119.  scores has G entries (-inf where not in ixs). */
void lcto_prefilter_scores(const lcto_locus *L, const double *M, const uint64_t *ixs, size_t n_ixs,
                           double *scores);
/* a3: src/solvers/solve.rs:52-84.  Sorts/truncates ixs in place, returns new length. */
size_t lcto_truncate_ixs(uint64_t *ixs, size_t n, const double *scores, double filt_diff,
                         size_t min_size, size_t threads);

/* a5: instance build for one genotype.  Outputs are malloc'ed; free with lcto_free. */
typedef struct lcto_instance {
    uint32_t n_reads;
    uint32_t ploidy;
    uint32_t total_windows;
    uint32_t n_alns;            /* A */
    uint32_t n_nontrivial;
    uint32_t haps[LCTO_MAX_PLOIDY];
    uint32_t wshift[LCTO_MAX_PLOIDY + 1];
    uint32_t *read_ixs;         /* [R+1] */
    uint32_t *nontrivial;       /* [n_nontrivial] */
    double   *aln_ln_prob;      /* [A] */
    uint8_t  *aln_contig_ix;    /* [A]; 255 = unmapped option (no parent) */
    uint32_t *aln_pa;           /* [A]; index into pa arrays; LCTO_NONE_U32 = unmapped option */
    uint32_t *aln_w;            /* [A*2] windows after the last apply_tweak */
    double   *win_weight;       /* [total_windows] after the last apply_tweak (0 = trivial) */
    uint8_t  *win_gc;           /* [total_windows] */
    uint8_t  *win_trivial;      /* [total_windows] */
} lcto_instance;

lcto_instance *lcto_instance_new(const lcto_locus *L, uint64_t g);
void lcto_instance_free(lcto_instance *I);
/* a6: src/model/assgn.rs:127-151 */
void lcto_apply_tweak(const lcto_locus *L, lcto_instance *I, lcto_rng *rng);

/* One solver attempt (a8-a12).  read_assgn [R] out, depth [total_windows] out. */
typedef struct lcto_attempt_out {
    double lik;        /* depth_contrib*depth_lik + aln_contrib*aln_lik  (WITHOUT prior) */
    double aln_lik;
    double depth_lik;
    uint64_t iterations; /* greedy iterations / anneal steps executed (diagnostic) */
    uint64_t moves;      /* accepted reassignments (diagnostic) */
} lcto_attempt_out;

int lcto_solve_attempt(const lcto_locus *L, const lcto_instance *I, const lcto_stage *st, lcto_rng *rng,
                       uint16_t *read_assgn, uint32_t *depth, lcto_attempt_out *out);

/* a13/a14: one stage over explicit worker chunks.
 *   worker_off [n_workers+1] into worker_ixs; worker_rng [n_workers] in/out.
 *   Per position j in worker_ixs: lik_mean[j], lik_var[j]; optional liks [j*attempts+a];
 *   optional counts CSR: counts_off[j] (u64, n+1 entries) filled when counts != NULL with capacity counts_cap.
 *   os_threads = OS threads used to run the logical workers (results do not depend on it). */
int lcto_solve_stage(const lcto_locus *L, const lcto_stage *st,
                     const uint64_t *worker_ixs, const uint64_t *worker_off, size_t n_workers,
                     lcto_rng *worker_rng, int os_threads,
                     double *lik_mean, double *lik_var, double *liks,
                     uint64_t *counts_off, uint16_t *counts, uint64_t counts_cap,
                     uint64_t *n_alns_out, uint64_t *iters_out);

/* a15: src/solvers/solve.rs:425-480.  ixs in/out (sorted + truncated), returns new n. */
size_t lcto_discard_improbable(uint64_t *ixs, size_t n, const double *lik_mean, const double *lik_var,
                               const uint16_t *attempts, double prob_thresh, size_t out_size, size_t threads);
double lcto_compare_two_likelihoods(double m1, double v1, uint16_t a1, double m2, double v2, uint16_t a2);

/* Full pipeline = solve::solve() (src/solvers/solve.rs:926-981) minus file output. */
typedef struct lcto_result {
    uint64_t n_out;            /* <= 50 */
    uint64_t gt_ix[50];
    double   lik_mean[50];
    double   lik_var[50];
    uint16_t attempts[50];
    double   ln_prob[50];
    double   quality;
    uint32_t total_reads;
    uint32_t unexpl_reads;
    uint32_t warn_no_probable; /* GenotypingWarning::NoProbableGenotype */
    uint32_t warn_few_reads;   /* GenotypingWarning::FewReads */
    /* diagnostics */
    uint64_t n_filtered;       /* survivors of the prefilter */
    uint64_t n_stage_in[8];    /* genotypes entering each executed stage (0 = skipped) */
    double   t_prefilter_s, t_stages_s;
    /* Genotyping::find_weighted_dist, src/solvers/solve.rs:616-632 (filled by lcto_find_weighted_dist) */
    uint32_t has_dist, true_edit_distances, has_weight_dist, _pad;
    double   weight_dist;
    uint32_t dist_to_primary[50];   /* LCTO_NONE_U32 = None */
} lcto_result;

/* Genotyping::find_weighted_dist (src/solvers/solve.rs:616-632) + genotype_distance (:339-357) over the linear
 * storage of TriangleMatrix<Option<u32>> (src/ext/trimat.rs), LCTO_NONE_U32 = None. */
void lcto_find_weighted_dist(const lcto_locus *L, lcto_result *res, const uint32_t *dist, int true_edit_distances);

int lcto_solve(const lcto_locus *L, const lcto_stage *stages, size_t n_stages, size_t threads,
               lcto_rng *rng, int os_threads, lcto_result *res,
               /* optional dumps (may be NULL): */
               double *scores_out /* [G] */, uint64_t *filtered_ixs_out /* [G] */);

/* Debug dumps in the reference's `--debug 2` formats (sol.csv, sol_ext.csv; uncompressed): rows are appended by
 * lcto_solve / lcto_solve_stage between open and close.  hap_names[H] must outlive the sink.  Not thread-safe across
 * concurrent lcto_solve calls (oracle/rust_diff.sh runs one locus at a time). */
int  lcto_debug_open(const char *sol_path, const char *sol_ext_path, const char *const *hap_names);
int  lcto_debug_open_depth(const char *depth_path);
void lcto_debug_close(void);

const char *lcto_version(void);

/* ------------------------------------------------ pair alignments (lcto_pairs.c) */

/* Mate alignments of R read pairs, per read sorted by (contig asc, read end asc, ln_prob desc). */
typedef struct lcto_mates {
    uint32_t n_reads, n_haps, max_alns, ins_len;
    const uint64_t *ma_off;      /* [R+1] */
    const uint32_t *ma_contig;   /* [N] */
    const uint8_t  *ma_flags;    /* [N] bit0: read end (0 first / 1 second), bit1: strand */
    const uint32_t *ma_start;    /* [N] interval start */
    const uint32_t *ma_end;      /* [N] interval end (exclusive) */
    const double   *ma_ln_prob;  /* [N] */
    const double   *read_weight; /* [R] or NULL (= 1.0) */
    const double   *ins_ln_pmf;  /* [ins_len] InsertDistr::ln_prob(size) */
    double unmapped_penalty, insert_penalty, prob_diff;
    uint32_t single_end;         /* identify_single_end_alignments, src/model/locs.rs:870-911 */
    uint32_t window;             /* ContigInfo::window_size (explicit weights) */
    const uint64_t *exp_off;     /* [H+1] explicit weights per contig position (ExplicitWeights::at), or NULL */
    const double   *exp_weight;
    const uint8_t  *read_max_alns;   /* [R] per-read max_alns (locs.rs:1263), or NULL (= max_alns) */
} lcto_mates;

/* identify_paired_end_alignments for every read (src/model/locs.rs:805-868). Outputs are the `pa_*` /
 * `unmapped_prob` arrays of the flat locus.  0 = ok, -2 = insert size outside the table, -3 = cap too small. */
int lcto_pair_alignments(const lcto_mates *in, uint64_t cap, uint64_t *pa_off, uint32_t *pa_contig,
                         double *pa_ln_prob, uint32_t *pa_mid1, uint32_t *pa_mid2, double *unmapped_prob);

/* ------------------------------------------------ alignment rescoring (lcto_rescore.c; SURVEY 8(f) rank 2, first slice) */

typedef struct lcto_alns {
    uint64_t n_alns;
    const uint64_t *cigar_off;     /* [n_alns+1] */
    const uint32_t *cigar_ops;     /* BAM encoding len << 4 | op (1=I 2=D 4=S 7='=' 8=X) */
    const uint32_t *aln_start, *aln_end, *contig_len, *passable_dist;   /* [n_alns] each */
    double ln_match, ln_mismatch, ln_insertion, ln_deletion, ln_clipping;
} lcto_alns;

int lcto_rescore_alignments(const lcto_alns *in, double *ln_prob, uint32_t *edit, uint32_t *read_len, uint8_t *save);

/* second slice: the per read-end protocol around push (read_next_alns, src/model/locs.rs:502-567) with the PosCollection
 * de-duplication of alignment starts (:166-187, 315-343).  Records are grouped by (read, read end); the first record of a
 * group is the primary alignment. */
typedef struct lcto_read_ends {
    lcto_alns alns;                    /* passable_dist is ignored (derived per group) */
    uint64_t n_groups;
    const uint64_t *grp_off;           /* [n_groups+1] */
    const uint32_t *rec_contig;        /* [n_alns] */
    const uint8_t  *grp_read_end;      /* [n_groups] 0 / 1 */
    const uint32_t *grp_read_len;      /* record.seq().len() */
    const uint32_t *grp_good_dist, *grp_passable_dist;    /* EditDistCache::get(read_len) */
    const double   *grp_neighb_complexity;                /* 1.0 for long reads (locs.rs:527) */
    double poor_compl, poor_compl_edit;                   /* Params */
    uint32_t strict_subset;
} lcto_read_ends;
int lcto_collect_read_ends(const lcto_read_ends *in, double *ln_prob, uint32_t *edit, uint32_t *read_len,
                           uint8_t *ok, uint32_t *best_edit, double *weight_factor, uint32_t *thr_dist,
                           uint32_t *pass_dist, uint32_t *n_kept, uint32_t *kept_rec);

/* from the per read-end results to the pairing input (lcto_group.c): AllAlignments::load after read_next_alns
 * (src/model/locs.rs:1117-1137) + recover_and_group_alignments without the transfer (:1237-1288). */
typedef struct lcto_prelim {
    uint64_t n_reads, n_groups;
    const int64_t  *read_group;        /* [n_reads][2] group of the first / second read end, -1 = none */
    const uint64_t *grp_off;           /* [n_groups+1] */
    const uint32_t *rec_contig, *rec_start, *rec_end;   /* [n_alns] */
    const uint8_t  *rec_strand;        /* [n_alns] 1 = reverse */
    const double   *rec_ln_prob;       /* [n_alns] */
    const uint8_t  *grp_ok;            /* [n_groups] */
    const uint32_t *grp_best_edit, *grp_thr_dist, *grp_n_kept;
    const uint32_t *kept_rec;          /* [n_alns] */
    const uint32_t *contig_len;        /* [n_haps] */
    const double   *read_weight;       /* [n_reads] */
    double min_weight;
    uint32_t n_haps, boundary, single_end, _pad;
} lcto_prelim;
int lcto_group_reads(const lcto_prelim *in, uint64_t cap, uint8_t *status, uint64_t *n_reads_out, uint32_t *out_read,
                     uint8_t *out_max_alns, uint64_t *ma_off, uint32_t *ma_contig, uint8_t *ma_flags,
                     uint32_t *ma_start, uint32_t *ma_end, double *ma_ln_prob, uint32_t *ma_rec, uint64_t *counts);

/* read weights from the k-mers unique to the locus (lcto_weights.c): UniqueKmers, src/model/locs.rs:915-1003 */
typedef struct lcto_unique_kmers lcto_unique_kmers;
lcto_unique_kmers *lcto_unique_kmers_build(const uint8_t *seqs, const uint64_t *seq_off, uint64_t n_seqs,
                                           const uint16_t *kmer_counts, const uint64_t *cnt_off, uint32_t k,
                                           uint16_t hard_threshold, uint16_t soft_threshold);
uint64_t lcto_unique_kmers_count(const lcto_unique_kmers *u);
void lcto_unique_kmers_free(lcto_unique_kmers *u);
int lcto_read_weights(const lcto_unique_kmers *u, const uint8_t *seqs, const uint64_t *seq_off, uint64_t n_reads,
                      uint32_t ends, uint16_t *unique, double *weight);

/* ------------------------------------------------ short-read recruitment (lcto_recruit.c) */

typedef struct lcto_target_seqs {
    uint64_t n_seqs;
    const uint64_t *seq_off;        /* [n_seqs+1] into seqs */
    const uint8_t  *seqs;           /* ASCII, A/C/G/T; anything else is an N */
    const uint32_t *seq_locus;      /* [n_seqs] locus index, ascending (TargetBuilder::add is called locus by locus) */
    const uint64_t *cnt_off;        /* [n_seqs+1] into kmer_counts */
    const uint16_t *kmer_counts;    /* KmerCounts of the sequence: len + 1 - base_k entries */
    uint32_t base_k, minimizer_k, minimizer_w, thresh_kmer_count;
    double match_frac;
    uint32_t match_length, _pad;
} lcto_target_seqs;
typedef struct lcto_reads {
    uint64_t n_reads;
    const uint64_t *off1; const uint8_t *seq1;      /* first mates / single-end reads */
    const uint64_t *off2; const uint8_t *seq2;      /* second mates; NULL = single-end */
} lcto_reads;
typedef struct lcto_targets lcto_targets;
size_t lcto_minimizers(const uint8_t *seq, size_t len, uint32_t k, uint32_t w, uint64_t *hash, uint32_t *pos, uint8_t *fw, size_t cap);
void lcto_fraction_approximate_u16(double x, uint16_t *num, uint16_t *den);
lcto_targets *lcto_targets_build(const lcto_target_seqs *in);
void lcto_targets_free(lcto_targets *T);
size_t lcto_targets_entries(const lcto_targets *T, uint64_t *key, uint32_t *locus, uint8_t *info, size_t cap);
int lcto_recruit_short(const lcto_targets *T, const lcto_reads *R, uint32_t cap, uint32_t *ans_count, uint32_t *ans_locus);

#ifdef __cplusplus
}
#endif
#endif

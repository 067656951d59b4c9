/*
 * lcto_solve.c -- ORACLE (test infrastructure): prefilter, stage scheduler, pruning, result.
 * Plain-C restatement of src/solvers/solve.rs:
 *   :52-84   truncate_ixs          :87-122   run_filter
 *   :319-336 compare_two_likelihoods  :425-480 discard_improbable_genotypes
 *   :482-535 produce_result        :637-645,719-729 check_first_prob / count_unexplained_reads
 *   :789-850 solve_single_thread   :926-981 solve    :996-1093 MainWorker   :1104-1145 Worker::run
 * and src/math/mod.rs:180-220 (t-tests), src/ext/vec.rs:86-116 (mean / variance).
 *
 * `threads` (T) is a SEMANTIC parameter of the reference (number of RNG streams, minimum survivor
 * counts); the T logical workers are multiplexed onto `os_threads` pthreads and results do not
 * depend on os_threads.  Unpinned tie order: sort_unstable_by on equal keys is restated as a
 * stable sort (ties keep current order).
 */
#include "lcto.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <time.h>

static double now_s(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static inline int total_cmp(double a, double b) {
    int64_t x, y;
    memcpy(&x, &a, 8); memcpy(&y, &b, 8);
    x ^= (int64_t)(((uint64_t)(x >> 63)) >> 1);
    y ^= (int64_t)(((uint64_t)(y >> 63)) >> 1);
    return (x > y) - (x < y);
}

/* Stable descending sort of ixs by key[ixs[i]] (merge sort). */
static void sort_desc_stable(uint64_t *ixs, size_t n, const double *key) {
    if (n < 2) return;
    uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * n);
    for (size_t width = 1; width < n; width *= 2) {
        for (size_t lo = 0; lo < n; lo += 2 * width) {
            size_t mid = lo + width < n ? lo + width : n;
            size_t hi = lo + 2 * width < n ? lo + 2 * width : n;
            size_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                /* take right only if strictly greater -> stable */
                if (total_cmp(key[ixs[j]], key[ixs[i]]) > 0) tmp[k++] = ixs[j++];
                else tmp[k++] = ixs[i++];
            }
            while (i < mid) tmp[k++] = ixs[i++];
            while (j < hi) tmp[k++] = ixs[j++];
        }
        memcpy(ixs, tmp, sizeof(uint64_t) * n);
    }
    free(tmp);
}

/* ------------------------------------------------------------ a2: prefilter */

/* src/solvers/solve.rs:105-119 -- single thread, sequential sum in read order. */
void lcto_prefilter_scores(const lcto_locus *L, const double *M, const uint64_t *ixs, size_t n_ixs,
                           double *scores) {
    uint32_t R = L->n_reads, p = L->ploidy;
    uint64_t G = L->n_genotypes;
    for (uint64_t g = 0; g < G; g++) scores[g] = -INFINITY;
    double *best = (double *)malloc(sizeof(double) * (R ? R : 1));
    uint32_t ids[LCTO_MAX_PLOIDY];
    for (size_t q = 0; q < n_ixs; q++) {
        uint64_t g = ixs[q];
        lcto_genotype_tuple(L, g, ids);
        double prior = L->priors ? L->priors[g] : 0.0;
        memcpy(best, M + (size_t)ids[0] * R, sizeof(double) * R);
        for (uint32_t k = 1; k < p; k++) {
            const double *row = M + (size_t)ids[k] * R;
            for (uint32_t r = 0; r < R; r++) best[r] = fmax(best[r], row[r]);
        }
        double s = 0.0;
        for (uint32_t r = 0; r < R; r++) s = s + best[r];
        scores[g] = prior + s;
    }
    free(best);
}

/* src/solvers/solve.rs:52-84 */
size_t lcto_truncate_ixs(uint64_t *ixs, size_t n, const double *scores, double filt_diff,
                         size_t min_size, size_t threads) {
    sort_desc_stable(ixs, n, scores);
    double best = scores[ixs[0]];
    double worst = scores[ixs[n - 1]];
    double thresh = best - filt_diff;
    if (min_size >= n || worst >= thresh) return n;
    size_t m = 0;
    while (m < n && scores[ixs[m]] >= thresh) m++;       /* partition_point */
    if (m < min_size) {
        thresh = scores[ixs[min_size - 1]];
        m = 0;
        while (m < n && scores[ixs[m]] >= thresh) m++;
    }
    if (m < threads) m = threads;
    if (m > n) m = n;
    return m;
}

/* ------------------------------------------------------------ debug dumps (--debug 2 of the reference) */

/* sol.csv / sol_ext.csv in the reference's own row formats (src/solvers/solve.rs:115-117,1074-1075,938,895-896;
 * src/model/assgn.rs:413-425), written by lcto_solve / lcto_solve_stage while a sink is open.  This is what
 * oracle/rust_diff.sh diffs against the files of a real `locityper genotype --debug 2` run.  Row order inside a stage
 * depends on thread timing in the reference, so the comparison sorts the rows. */
#include <stdio.h>
static FILE *g_sol = NULL, *g_sol_ext = NULL, *g_depth = NULL;
static const char *const *g_hap_names = NULL;
static unsigned g_stage_no = 0;                      /* 1-based stage of the rows being written */
static pthread_mutex_t g_dbg_mutex = PTHREAD_MUTEX_INITIALIZER;

static void fprint_gt(FILE *f, const lcto_locus *L, uint64_t g) {
    uint32_t ids[LCTO_MAX_PLOIDY];
    lcto_genotype_tuple(L, g, ids);
    for (uint32_t k = 0; k < L->ploidy; k++) fprintf(f, "%s%s", k ? "," : "", g_hap_names[ids[k]]);   /* Genotype::new, contigs.rs:412-424 */
}

int lcto_debug_open(const char *sol_path, const char *sol_ext_path, const char *const *hap_names) {
    g_sol = fopen(sol_path, "w");
    g_sol_ext = fopen(sol_ext_path, "w");
    if (!g_sol || !g_sol_ext) return -1;
    g_hap_names = hap_names;
    fprintf(g_sol, "stage\tgenotype\tscore\n");                                             /* solve.rs:938 */
    fprintf(g_sol_ext, "stage\tgenotype\tattempt\ttotal_reads\tunmapped\tout_of_bounds\taln_lik\tdepth_lik\tlik\n");
    return 0;
}

/* depth.csv of `--debug 3` (DebugLvl::Full): DebugFiles::new, src/solvers/solve.rs:882-888; call after lcto_debug_open */
int lcto_debug_open_depth(const char *depth_path) {
    g_depth = fopen(depth_path, "w");
    if (!g_depth) return -1;
    fprintf(g_depth, "stage\tgenotype\tattempt\tcontig\twindow\tweight\tdepth\tlik\n");
    return 0;
}

void lcto_debug_close(void) {
    if (g_sol) fclose(g_sol);
    if (g_sol_ext) fclose(g_sol_ext);
    if (g_depth) fclose(g_depth);
    g_depth = NULL;
    g_sol = g_sol_ext = NULL;
    g_hap_names = NULL;
}

#define LCTO_INV_LN10 0.4342944819032518277        /* Ln::INV_LN10, src/math/mod.rs:14 */

/* ------------------------------------------------------------ a13/a14: stage over workers */

typedef struct {
    const lcto_locus *L; const lcto_stage *st;
    const uint64_t *worker_ixs; const uint64_t *worker_off; size_t n_workers;
    lcto_rng *worker_rng;
    double *lik_mean, *lik_var, *liks;
    uint64_t *n_alns_out; uint64_t *iters_out;
    uint16_t **counts_tmp;     /* per position, malloc'ed when counts requested */
    int want_counts;
    volatile long next_worker; /* atomic work queue */
    int error;
} stage_ctx;

/* F64Ext::mean_variance_or_nan, src/ext/vec.rs:74-78,86-93,109-116 */
static void mean_variance_or_nan(const double *a, size_t n, double *mean, double *var) {
    if (n == 0) { *mean = NAN; *var = NAN; return; }
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s = s + a[i];
    double m = s / (double)n;
    *mean = m;
    if (n == 1) { *var = NAN; return; }
    double acc = 0.0;
    for (size_t i = 0; i < n; i++) { double d = a[i] - m; acc = acc + d * d; }
    *var = acc / (double)(n - 1);
}

/* Worker::run body for one logical worker, src/solvers/solve.rs:1116-1142 */
static int run_worker(stage_ctx *C, size_t w) {
    const lcto_locus *L = C->L; const lcto_stage *st = C->st;
    lcto_rng *rng = &C->worker_rng[w];
    uint32_t attempts = st->attempts;
    double *liks = (double *)malloc(sizeof(double) * attempts);
    uint16_t *read_assgn = (uint16_t *)malloc(sizeof(uint16_t) * (L->n_reads + 1));
    for (uint64_t j = C->worker_off[w]; j < C->worker_off[w + 1]; j++) {
        uint64_t g = C->worker_ixs[j];
        double prior = L->priors ? L->priors[g] : 0.0;
        lcto_instance *I = lcto_instance_new(L, g);
        uint32_t *depth = (uint32_t *)malloc(sizeof(uint32_t) * I->total_windows);
        uint16_t *counts = (uint16_t *)calloc(I->n_alns + 1, sizeof(uint16_t));
        uint64_t iters = 0;
        for (uint32_t a = 0; a < attempts; a++) {
            lcto_apply_tweak(L, I, rng);
            lcto_attempt_out out;
            if (lcto_solve_attempt(L, I, st, rng, read_assgn, depth, &out) != 0) {
                free(depth); free(counts); lcto_instance_free(I); free(liks); free(read_assgn);
                return -1;
            }
            liks[a] = prior + out.lik;                              /* solve.rs:1126 */
            if (g_depth) {                                          /* ReadAssignment::write_depth, assgn.rs:359-372 */
                pthread_mutex_lock(&g_dbg_mutex);
                for (uint32_t i = 0; i < I->ploidy; i++)
                    for (uint32_t w = I->wshift[i]; w < I->wshift[i + 1]; w++) {
                        double lp = I->win_trivial[w] ? 0.0
                                  : I->win_weight[w] * L->depth_table[(size_t)I->win_gc[w] * L->depth_k + depth[w]];
                        fprintf(g_depth, "%u\t", g_stage_no);
                        fprint_gt(g_depth, L, g);
                        fprintf(g_depth, "\t%u\t%u\t%u\t%.4f\t%u\t%.3f\n", a + 1, i + 1, w - I->wshift[i] + 1,
                                I->win_weight[w], depth[w], lp * LCTO_INV_LN10);
                    }
                pthread_mutex_unlock(&g_dbg_mutex);
            }
            if (g_sol_ext) {                                        /* ReadAssignment::summarize, assgn.rs:416-425 */
                pthread_mutex_lock(&g_dbg_mutex);
                fprintf(g_sol_ext, "%u\t", g_stage_no);
                fprint_gt(g_sol_ext, L, g);
                fprintf(g_sol_ext, "\t%u\t%u\t%u\t%u\t%.7f\t%.7f\t%.7f\n", a + 1, I->n_reads, depth[0], depth[1],
                        out.aln_lik * LCTO_INV_LN10, out.depth_lik * LCTO_INV_LN10, out.lik * LCTO_INV_LN10);
                pthread_mutex_unlock(&g_dbg_mutex);
            }
            iters += out.iterations;
            for (uint32_t r = 0; r < I->n_reads; r++)               /* update_counts, assgn.rs:374-378 */
                counts[I->read_ixs[r] + read_assgn[r]] += 1;
            if (C->liks) C->liks[j * attempts + a] = liks[a];
        }
        mean_variance_or_nan(liks, attempts, &C->lik_mean[j], &C->lik_var[j]);
        if (g_sol) {                                                /* MainWorker::run, solve.rs:1074-1075 */
            pthread_mutex_lock(&g_dbg_mutex);
            fprintf(g_sol, "%u\t", g_stage_no);
            fprint_gt(g_sol, L, g);
            fprintf(g_sol, "\t%.4f\n", C->lik_mean[j] * LCTO_INV_LN10);
            pthread_mutex_unlock(&g_dbg_mutex);
        }
        if (C->n_alns_out) C->n_alns_out[j] = I->n_alns;
        if (C->iters_out) C->iters_out[j] = iters;
        if (C->want_counts) C->counts_tmp[j] = counts; else free(counts);
        free(depth);
        lcto_instance_free(I);
    }
    free(liks); free(read_assgn);
    return 0;
}

static void *stage_thread(void *arg) {
    stage_ctx *C = (stage_ctx *)arg;
    for (;;) {
        long w = __sync_fetch_and_add(&C->next_worker, 1);
        if ((size_t)w >= C->n_workers) break;
        if (run_worker(C, (size_t)w) != 0) C->error = 1;
    }
    return NULL;
}

int lcto_solve_stage(const lcto_locus *L, const lcto_stage *st,
                     const uint64_t *worker_ixs, const uint64_t *worker_off, size_t n_workers,
                     lcto_rng *worker_rng, int os_threads,
                     double *lik_mean, double *lik_var, double *liks,
                     uint64_t *counts_off, uint16_t *counts, uint64_t counts_cap,
                     uint64_t *n_alns_out, uint64_t *iters_out) {
    stage_ctx C;
    memset(&C, 0, sizeof(C));
    C.L = L; C.st = st; C.worker_ixs = worker_ixs; C.worker_off = worker_off; C.n_workers = n_workers;
    C.worker_rng = worker_rng; C.lik_mean = lik_mean; C.lik_var = lik_var; C.liks = liks;
    C.n_alns_out = n_alns_out; C.iters_out = iters_out;
    size_t n = (size_t)worker_off[n_workers];
    C.want_counts = (counts != NULL && counts_off != NULL);
    if (C.want_counts) C.counts_tmp = (uint16_t **)calloc(n + 1, sizeof(uint16_t *));
    uint64_t *nal = NULL;
    if (C.want_counts && !n_alns_out) { nal = (uint64_t *)calloc(n + 1, sizeof(uint64_t)); C.n_alns_out = nal; }

    int nt = os_threads < 1 ? 1 : os_threads;
    if ((size_t)nt > n_workers) nt = (int)n_workers;
    if (nt <= 1) {
        for (size_t w = 0; w < n_workers; w++) if (run_worker(&C, w) != 0) C.error = 1;
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nt);
        for (int i = 0; i < nt; i++) pthread_create(&th[i], NULL, stage_thread, &C);
        for (int i = 0; i < nt; i++) pthread_join(th[i], NULL);
        free(th);
    }
    int rc = C.error ? -1 : 0;
    if (C.want_counts) {
        uint64_t off = 0;
        for (size_t j = 0; j < n; j++) {
            counts_off[j] = off;
            uint64_t a = C.n_alns_out[j];
            if (C.counts_tmp[j]) {
                if (off + a <= counts_cap) memcpy(counts + off, C.counts_tmp[j], sizeof(uint16_t) * a);
                else rc = -3;
                free(C.counts_tmp[j]);
            }
            off += a;
        }
        counts_off[n] = off;
        free(C.counts_tmp);
    }
    free(nal);
    return rc;
}

/* ------------------------------------------------------------ a15: pruning */

/* src/math/mod.rs:180-196 (DIFF_VAR = false -> Welch degrees of freedom) */
static double t_test_same(double m1, double v1, double m2, double v2, double n) {
    double var_sum = v1 + v2;
    double t_stat = (m1 - m2) * sqrt(n / var_sum);
    double freedom = (n - 1.0) * var_sum * var_sum / (v1 * v1 + v2 * v2);
    return lcto_students_t_cdf(t_stat, freedom);
}

/* src/math/mod.rs:199-220 */
static double t_test_diff(double m1, double v1, double m2, double v2, double n1, double n2) {
    double nv1 = v1 / n1, nv2 = v2 / n2;
    double s = nv1 + nv2;
    double t_stat = (m1 - m2) / sqrt(s);
    double freedom = s * s / (nv1 * nv1 / (n1 - 1.0) + nv2 * nv2 / (n2 - 1.0));
    return lcto_students_t_cdf(t_stat, freedom);
}

static inline int is_normal(double x) { return fpclassify(x) == FP_NORMAL; }

/* src/solvers/solve.rs:319-336 */
double lcto_compare_two_likelihoods(double m1, double v1, uint16_t a1, double m2, double v2, uint16_t a2) {
    double simple_norm = m1 - lcto_ln_add(m1, m2);
    if (is_normal(v1) && is_normal(v2)) {
        double t_pval = (a1 == a2) ? t_test_same(m1, v1, m2, v2, (double)a1)
                                   : t_test_diff(m1, v1, m2, v2, (double)a1, (double)a2);
        return fmax(simple_norm, log(t_pval));
    }
    return simple_norm;
}

/* src/solvers/solve.rs:425-480.  lik_mean / lik_var / attempts are indexed by genotype id. */
size_t lcto_discard_improbable(uint64_t *ixs, size_t n, const double *lik_mean, const double *lik_var,
                               const uint16_t *attempts, double prob_thresh, size_t out_size, size_t threads) {
    if (out_size < threads) out_size = threads;
    if (prob_thresh == -INFINITY || out_size >= n) return n;
    sort_desc_stable(ixs, n, lik_mean);
    uint64_t best = ixs[0];
    size_t m = out_size;
    if (out_size <= 500) {                    /* SOPHISTICATED_COUNT */
        uint32_t dropped = 0;
        for (size_t q = out_size; q < n; q++) {
            uint64_t ix = ixs[q];
            double ln_pval = lcto_compare_two_likelihoods(lik_mean[ix], lik_var[ix], attempts[ix],
                                                          lik_mean[best], lik_var[best], attempts[best]);
            if (ln_pval >= prob_thresh) ixs[m++] = ix;
            else { dropped++; if (dropped >= 5) break; }      /* STOP_COUNT */
        }
    }
    return m;
}

/* ------------------------------------------------------------ full pipeline */

/* src/solvers/solve.rs:482-535 */
static void produce_result(const lcto_locus *L, uint64_t *ixs, size_t n_ixs, const double *lik_mean,
                           const double *lik_var, const uint16_t *attempts, lcto_result *res) {
    const double THRESH = -11.512925464970229;
    size_t min_output = L->out_bams > 4 ? L->out_bams : 4;
    double thresh_prob = fmin(THRESH, L->prob_thresh);
    sort_desc_stable(ixs, n_ixs, lik_mean);
    size_t n = n_ixs < 50 ? n_ixs : 50;
    double ln_probs[50];
    for (size_t i = 0; i < 50; i++) ln_probs[i] = 0.0;
    size_t i = 0;
    while (i < n) {
        uint64_t u = ixs[i];
        size_t n_at_loop_start = n;
        for (size_t j = i + 1; j < n_at_loop_start; j++) {
            uint64_t v = ixs[j];
            double prob_j = lcto_compare_two_likelihoods(lik_mean[v], lik_var[v], attempts[v],
                                                         lik_mean[u], lik_var[u], attempts[u]);
            if (i == 0 && j >= min_output && prob_j < thresh_prob) { n = j; break; }
            ln_probs[i] += log1p(-exp(prob_j));
            ln_probs[j] += prob_j;
        }
        res->gt_ix[i] = u; res->lik_mean[i] = lik_mean[u]; res->lik_var[i] = lik_var[u];
        res->attempts[i] = attempts[u];
        i++;
    }
    double norm = lcto_ln_sum(ln_probs, n);
    for (size_t k = 0; k < n; k++) { ln_probs[k] -= norm; res->ln_prob[k] = ln_probs[k]; }
    double others = n >= 1 ? lcto_ln_sum(ln_probs + 1, n - 1) : -INFINITY;
    double q = -10.0 * (others * 0.4342944819032518277);       /* Phred::from_ln_prob */
    res->quality = fmin(q, 1e9);
    res->n_out = n;
}

/* src/solvers/solve.rs:926-981 */
int lcto_solve(const lcto_locus *L, const lcto_stage *stages, size_t n_stages, size_t threads,
               lcto_rng *rng, int os_threads, lcto_result *res,
               double *scores_out, uint64_t *filtered_ixs_out) {
    uint64_t G = L->n_genotypes;
    if (G == 0 || n_stages == 0 || n_stages > 8) return -1;
    memset(res, 0, sizeof(*res));
    if (threads > G) threads = G;                      /* src/command/genotype.rs:1247 */
    if (threads < 1) threads = 1;
    uint64_t *ixs = (uint64_t *)malloc(sizeof(uint64_t) * G);
    for (uint64_t g = 0; g < G; g++) ixs[g] = g;
    size_t n = G;
    double *lik_mean = (double *)malloc(sizeof(double) * G);
    double *lik_var = (double *)malloc(sizeof(double) * G);
    uint16_t *attempts = (uint16_t *)calloc(G, sizeof(uint16_t));
    for (uint64_t g = 0; g < G; g++) { lik_mean[g] = NAN; lik_var[g] = NAN; }

    double t0 = now_s();
    size_t out_size0 = stages[0].in_size;
    if (L->dont_skip || out_size0 < G) {               /* solve.rs:941-945 */
        double *M = (double *)malloc(sizeof(double) * (size_t)L->n_haps * L->n_reads);
        double *scores = (double *)malloc(sizeof(double) * G);
        lcto_best_aln_matrix(L, M);
        lcto_prefilter_scores(L, M, ixs, n, scores);
        if (g_sol)                                     /* run_filter, solve.rs:115-117 */
            for (size_t q = 0; q < n; q++) {
                fprintf(g_sol, "0\t");
                fprint_gt(g_sol, L, ixs[q]);
                fprintf(g_sol, "\t%.3f\n", scores[ixs[q]] * LCTO_INV_LN10);
            }
        n = lcto_truncate_ixs(ixs, n, scores, L->filt_diff, out_size0, threads);
        if (scores_out) memcpy(scores_out, scores, sizeof(double) * G);
        free(M); free(scores);
    }
    res->n_filtered = n;
    if (filtered_ixs_out) memcpy(filtered_ixs_out, ixs, sizeof(uint64_t) * n);
    double t1 = now_s();
    res->t_prefilter_s = t1 - t0;

    /* MainWorker::new, solve.rs:1007-1018: worker w gets a clone of the locus stream, then jump(). */
    lcto_rng *wrng = NULL;
    if (threads > 1) {
        wrng = (lcto_rng *)malloc(sizeof(lcto_rng) * threads);
        for (size_t w = 0; w < threads; w++) { wrng[w] = *rng; lcto_rng_jump(rng); }
    }
    int rc = 0;
    for (size_t s = 0; s < n_stages && rc == 0; s++) {
        const lcto_stage *st = &stages[s];
        int has_next = s + 1 < n_stages;
        size_t out_size = has_next ? stages[s + 1].in_size : 0;
        if (!(L->dont_skip || !has_next || out_size < n)) continue;     /* solve.rs:1041-1045 */
        res->n_stage_in[s] = n;
        g_stage_no = (unsigned)s + 1;
        double *lm = (double *)malloc(sizeof(double) * n);
        double *lv = (double *)malloc(sizeof(double) * n);
        uint64_t *off;
        size_t n_workers;
        size_t *tmp = NULL;
        if (threads == 1) {                            /* solve_single_thread, solve.rs:814-843 */
            n_workers = 1;
            off = (uint64_t *)malloc(sizeof(uint64_t) * 2);
            off[0] = 0; off[1] = n;
            rc = lcto_solve_stage(L, st, ixs, off, 1, rng, 1, lm, lv, NULL, NULL, NULL, 0, NULL, NULL);
        } else {
            /* MainWorker::run, solve.rs:1049-1063: shuffle, then static contiguous partition. */
            tmp = (size_t *)malloc(sizeof(size_t) * n);
            for (size_t q = 0; q < n; q++) tmp[q] = (size_t)ixs[q];
            lcto_rng_shuffle_usize(rng, tmp, n);
            for (size_t q = 0; q < n; q++) ixs[q] = (uint64_t)tmp[q];
            free(tmp);
            off = (uint64_t *)malloc(sizeof(uint64_t) * (threads + 1));
            size_t start = 0; n_workers = 0;
            off[0] = 0;
            for (size_t i = 0; i < threads; i++) {
                if (start == n) break;
                size_t rem_workers = threads - i;
                size_t curr = (n - start + rem_workers - 1) / rem_workers;   /* fast_ceil_div */
                start += curr;
                off[++n_workers] = start;
            }
            rc = lcto_solve_stage(L, st, ixs, off, n_workers, wrng, os_threads, lm, lv, NULL, NULL, NULL, 0, NULL, NULL);
        }
        for (size_t q = 0; q < n; q++) {
            lik_mean[ixs[q]] = lm[q]; lik_var[ixs[q]] = lv[q]; attempts[ixs[q]] = (uint16_t)st->attempts;
        }
        free(lm); free(lv); free(off);
        if (has_next && rc == 0)
            n = lcto_discard_improbable(ixs, n, lik_mean, lik_var, attempts, L->prob_thresh, out_size, threads);
    }
    res->t_stages_s = now_s() - t1;
    if (rc == 0) {
        produce_result(L, ixs, n, lik_mean, lik_var, attempts, res);
        res->total_reads = L->n_reads;
        /* check_first_prob, solve.rs:637-645 */
        double lp0 = res->ln_prob[0];
        res->warn_no_probable = (isnan(lp0) || lp0 < -2.0 * 2.302585092994045684) ? 1 : 0;
        /* check_num_of_reads, solve.rs:649-678 */
        uint32_t p = L->ploidy, nr = L->n_reads;
        if (nr < p) res->warn_few_reads = 1;
        else if (p > 1 && nr < p * 10) {
            double k = (double)p, nn = (double)nr;
            double exp_zeros = exp(log(k - 1.0) * nn - log(k) * (nn - 1.0));
            if (exp_zeros > 0.1) res->warn_few_reads = 1;
        }
        /* count_unexplained_reads, solve.rs:719-729 */
        uint32_t ids[LCTO_MAX_PLOIDY];
        lcto_genotype_tuple(L, res->gt_ix[0], ids);
        uint32_t unexpl = 0;
        for (uint32_t r = 0; r < L->n_reads; r++) {
            double best = -INFINITY;
            for (uint32_t k = 0; k < p; k++) {
                /* best_at_contig, locs.rs:605-611 */
                double v = L->unmapped_prob[r];
                for (uint64_t e = L->pa_off[r]; e < L->pa_off[r + 1]; e++)
                    if (L->pa_contig[e] == ids[k]) { v = L->pa_ln_prob[e]; break; }
                best = fmax(best, v);
            }
            unexpl += best < L->unmapped_prob[r] + 1e-8 ? 1u : 0u;
        }
        res->unexpl_reads = unexpl;
    }
    free(ixs); free(lik_mean); free(lik_var); free(attempts); free(wrng);
    return rc;
}

/* ------------------------------------------------------------------ weighted distance */

/* TriangleMatrix::to_linear_index via get_symmetric (src/ext/trimat.rs:41-45) */
static size_t tri_ix(size_t side, size_t i, size_t j) {
    if (i > j) { size_t t = i; i = j; j = t; }
    return (2 * side - 3 - i) * i / 2 + j - 1;
}

static void perm_visit(const uint32_t *perm, const uint32_t *gt2, uint32_t p, uint32_t H, const uint32_t *dist,
                       uint32_t *min_dist) {
    /* closure body of genotype_distance, src/solvers/solve.rs:342-355 */
    uint32_t d = 0;
    for (uint32_t k = 0; k < p; k++) {
        if (perm[k] != gt2[k]) {
            uint32_t e = dist[tri_ix(H, perm[k], gt2[k])];
            if (e == LCTO_NONE_U32) { d = UINT32_MAX; break; }
            d += e;
        }
    }
    if (d < *min_dist) *min_dist = d;
}

/* genotype_distance (src/solvers/solve.rs:339-357) with ext::vec::gen_permutations (src/ext/vec.rs:342-372):
 * n == 1: the tuple; n == 2: the tuple and its swap; n >= 3: Heap's algorithm, where `action` runs after every
 * swap only -- the initial arrangement itself is never visited. */
static uint32_t genotype_distance(const uint32_t *gt1, const uint32_t *gt2, uint32_t p, uint32_t H, const uint32_t *dist) {
    uint32_t min_dist = UINT32_MAX;
    if (p == 1) perm_visit(gt1, gt2, p, H, dist, &min_dist);
    else if (p == 2) {
        uint32_t sw[2] = { gt1[1], gt1[0] };
        perm_visit(gt1, gt2, p, H, dist, &min_dist);
        perm_visit(sw, gt2, p, H, dist, &min_dist);
    } else {
        uint32_t buffer[LCTO_MAX_PLOIDY];
        size_t c[LCTO_MAX_PLOIDY];
        for (uint32_t k = 0; k < p; k++) { buffer[k] = gt1[k]; c[k] = 0; }
        size_t i = 1;
        while (i < p) {
            if (c[i] < i) {
                size_t j = c[i] * (i % 2);
                uint32_t t = buffer[i]; buffer[i] = buffer[j]; buffer[j] = t;
                perm_visit(buffer, gt2, p, H, dist, &min_dist);
                c[i] += 1;
                i = 1;
            } else {
                c[i] = 0;
                i += 1;
            }
        }
    }
    return min_dist;
}

void lcto_find_weighted_dist(const lcto_locus *L, lcto_result *res, const uint32_t *dist, int true_edit_distances) {
    if (res->n_out == 0) return;
    uint32_t p = L->ploidy, H = L->n_haps;
    uint32_t gt0[LCTO_MAX_PLOIDY], gt[LCTO_MAX_PLOIDY];
    lcto_genotype_tuple(L, res->gt_ix[0], gt0);
    double sum_prob = 0.0, sum_dist = 0.0;
    int known = 1;
    for (uint64_t i = 0; i < res->n_out; i++) {
        double prob = exp(res->ln_prob[i]);
        sum_prob += prob;
        uint32_t d = 0;
        if (i > 0) {
            lcto_genotype_tuple(L, res->gt_ix[i], gt);
            d = genotype_distance(gt0, gt, p, H, dist);
        }
        /* sum_dist.zip(dist).map(|(s, d)| s + prob * d): None once any distance is None */
        if (d == UINT32_MAX) known = 0;
        else if (known) sum_dist += prob * (double)d;
        res->dist_to_primary[i] = d;
    }
    res->has_dist = 1;
    res->true_edit_distances = true_edit_distances ? 1 : 0;
    res->has_weight_dist = (uint32_t)known;
    res->weight_dist = known ? sum_dist / sum_prob : NAN;
}

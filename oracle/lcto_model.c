/*
 * lcto_model.c -- ORACLE (test infrastructure): per-genotype problem instance, assignment state,
 * and the two stochastic solvers.  Plain-C restatement of
 *   src/model/locs.rs:600-629,1203-1212     best_aln_matrix / contig_alns
 *   src/model/windows.rs:62-68,112-136,439-445,465-486,721-797   windows, tweak, extend_read_gt_alns
 *   src/model/assgn.rs:41-84,127-151,192-378,451-471   GenotypeAlignments / ReadAssignment / ReassignmentTarget
 *   src/model/distr_cache.rs:34-39,83-92    WindowDistr
 *   src/solvers/stoch.rs:19-29,81-120,197-242   Greedy / SimAnneal
 *   src/solvers/mod.rs:61-72                trivial short-circuit
 *   src/ext/vec.rs:298-339                  genotype enumeration order
 * Compile with -ffp-contract=off: every accept/reject below is a strict f64 comparison.
 */
#include "lcto.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <assert.h>

size_t lcto_sizeof_locus(void) { return sizeof(lcto_locus); }
size_t lcto_sizeof_stage(void) { return sizeof(lcto_stage); }
const char *lcto_version(void) { return "lcto-oracle 0.1 (restates locityper v1.7.2)"; }

/* f64::total_cmp (Rust std): returns <0, 0, >0. */
static inline int total_cmp(double a, double b) {
    int64_t x, y;
    memcpy(&x, &a, 8); memcpy(&y, &b, 8);
    x ^= (int64_t)(((uint64_t)(x >> 63)) >> 1);
    y ^= (int64_t)(((uint64_t)(y >> 63)) >> 1);
    return (x > y) - (x < y);
}

/* f64::max (IEEE maxNum). */
static inline double f64_max(double a, double b) { return fmax(a, b); }

/* ------------------------------------------------------------ genotype enumeration */

static uint64_t n_choose_k(uint64_t n, uint64_t k) {
    /* src/ext/vec.rs count_combinations */
    if (k > n) return 0;
    uint64_t r = k < n - k ? k : n - k;
    uint64_t acc = 1;
    for (uint64_t v = 1; v <= r; v++) acc = acc * (n - v + 1) / v;
    return acc;
}

/* Tuple of genotype g: explicit list, or the g-th combination with replacement in the order
 * produced by gen_combinations_with_repl (src/ext/vec.rs:298-339: lexicographic, last index fastest). */
void lcto_genotype_tuple(const lcto_locus *L, uint64_t g, uint32_t *out) {
    uint32_t p = L->ploidy;
    if (L->gt_tuples) {
        for (uint32_t k = 0; k < p; k++) out[k] = L->gt_tuples[g * p + k];
        return;
    }
    uint32_t H = L->n_haps;
    uint32_t lo = 0;
    for (uint32_t d = 0; d < p; d++) {
        uint32_t rem = p - d - 1;
        for (uint32_t v = lo; v < H; v++) {
            /* tuples of length rem over [v, H): C(H - v + rem - 1, rem) */
            uint64_t cnt = rem == 0 ? 1 : n_choose_k((uint64_t)(H - v) + rem - 1, rem);
            if (g < cnt) { out[d] = v; lo = v; break; }
            g -= cnt;
        }
    }
}

/* ------------------------------------------------------------ a1: best_aln_matrix */

/* src/model/locs.rs:614-618: range of pair alignments of read r on contig h. */
static void contig_alns(const lcto_locus *L, uint32_t r, uint32_t h, uint64_t *i_out, uint64_t *j_out) {
    uint64_t lo = L->pa_off[r], hi = L->pa_off[r + 1];
    uint64_t a = lo, b = hi;
    while (a < b) {               /* bisect::left_by */
        uint64_t m = a + (b - a) / 2;
        if (L->pa_contig[m] < h) a = m + 1; else b = m;
    }
    uint64_t i = a;
    uint64_t j = i;
    while (j < hi && L->pa_contig[j] == h) j++;   /* bisect::right_boundary */
    *i_out = i; *j_out = j;
}

/* src/model/locs.rs:1203-1212 (+ :621-629) */
void lcto_best_aln_matrix(const lcto_locus *L, double *M) {
    uint32_t H = L->n_haps, R = L->n_reads;
    for (uint32_t r = 0; r < R; r++) {
        uint64_t j = L->pa_off[r], end = L->pa_off[r + 1];
        for (uint32_t h = 0; h < H; h++) {
            uint64_t i = j;
            while (j < end && L->pa_contig[j] == h) j++;
            M[(size_t)h * R + r] = (i == j) ? L->unmapped_prob[r] : L->pa_ln_prob[i];
        }
    }
}

/* ------------------------------------------------------------ a5: instance build */

typedef struct { double lp; uint32_t pa; uint8_t cix; } cand;

lcto_instance *lcto_instance_new(const lcto_locus *L, uint64_t g) {
    uint32_t R = L->n_reads, p = L->ploidy;
    lcto_instance *I = (lcto_instance *)calloc(1, sizeof(lcto_instance));
    I->n_reads = R; I->ploidy = p;
    lcto_genotype_tuple(L, g, I->haps);
    /* GenotypeWindows::new, src/model/windows.rs:721-739 */
    uint32_t ws = 2;
    I->wshift[0] = ws;
    for (uint32_t k = 0; k < p; k++) { ws += L->hap_n_windows[I->haps[k]]; I->wshift[k + 1] = ws; }
    I->total_windows = ws;

    size_t cap = (size_t)R * 2 + 16;
    cand *alns = (cand *)malloc(cap * sizeof(cand));
    size_t n = 0;
    I->read_ixs = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)R + 1));
    I->nontrivial = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)R + 1));
    I->read_ixs[0] = 0;
    uint32_t n_nt = 0;
    double prob_diff = L->prob_diff;
    for (uint32_t r = 0; r < R; r++) {
        /* extend_read_gt_alns, src/model/windows.rs:762-797 */
        size_t start_len = n;
        double unmapped_prob = L->unmapped_prob[r];
        double thresh = unmapped_prob - prob_diff;
        for (uint32_t k = 0; k < p; k++) {
            uint64_t i, j;
            contig_alns(L, r, I->haps[k], &i, &j);
            if (i < j) {
                thresh = f64_max(thresh, L->pa_ln_prob[i] - prob_diff);
                for (uint64_t e = i; e < j; e++) {
                    if (L->pa_ln_prob[e] >= thresh) {
                        if (n + 2 > cap) { cap *= 2; alns = (cand *)realloc(alns, cap * sizeof(cand)); }
                        alns[n].lp = L->pa_ln_prob[e]; alns[n].pa = (uint32_t)e; alns[n].cix = (uint8_t)k; n++;
                    } else break;
                }
            }
        }
        if (unmapped_prob >= thresh) {
            if (n + 2 > cap) { cap *= 2; alns = (cand *)realloc(alns, cap * sizeof(cand)); }
            alns[n].lp = unmapped_prob; alns[n].pa = LCTO_NONE_U32; alns[n].cix = 255; n++;
        }
        /* sort_unstable_by(b.total_cmp(a)): Rust std uses insertion_sort_shift_left for len <= 20,
         * which is stable; restated as insertion sort for every length (tie order for len > 20 unpinned). */
        for (size_t i = start_len + 1; i < n; i++) {
            cand x = alns[i];
            size_t j = i;
            while (j > start_len && total_cmp(alns[j - 1].lp, x.lp) < 0) { alns[j] = alns[j - 1]; j--; }
            alns[j] = x;
        }
        /* partition_point(ln_prob >= thresh) */
        size_t keep = 0;
        while (start_len + keep < n && alns[start_len + keep].lp >= thresh) keep++;
        n = start_len + keep;
        assert(keep > 0 && keep <= 65535);      /* src/model/assgn.rs:57-58 */
        I->read_ixs[r + 1] = (uint32_t)n;
        if (keep > 1) I->nontrivial[n_nt++] = r;
    }
    I->n_alns = (uint32_t)n;
    I->n_nontrivial = n_nt;
    I->aln_ln_prob = (double *)malloc(sizeof(double) * (n + 1));
    I->aln_contig_ix = (uint8_t *)malloc(n + 1);
    I->aln_pa = (uint32_t *)malloc(sizeof(uint32_t) * (n + 1));
    I->aln_w = (uint32_t *)calloc(2 * (n + 1), sizeof(uint32_t));   /* [UNMAPPED_WINDOW; 2] */
    for (size_t i = 0; i < n; i++) {
        I->aln_ln_prob[i] = alns[i].lp; I->aln_contig_ix[i] = alns[i].cix; I->aln_pa[i] = alns[i].pa;
    }
    free(alns);
    I->win_weight = (double *)calloc(I->total_windows, sizeof(double));
    I->win_gc = (uint8_t *)calloc(I->total_windows, 1);
    I->win_trivial = (uint8_t *)malloc(I->total_windows);
    memset(I->win_trivial, 1, I->total_windows);
    return I;
}

void lcto_instance_free(lcto_instance *I) {
    if (!I) return;
    free(I->read_ixs); free(I->nontrivial); free(I->aln_ln_prob); free(I->aln_contig_ix);
    free(I->aln_pa); free(I->aln_w); free(I->win_weight); free(I->win_gc); free(I->win_trivial);
    free(I);
}

/* ------------------------------------------------------------ a6: apply_tweak */

/* ContigInfo::get_shifted_window_ix + WindowGetter::middle_window, src/model/windows.rs:62-68,465-470 */
static inline uint32_t shifted_window_ix(const lcto_locus *L, uint32_t hap, uint32_t shift, uint32_t middle) {
    if (middle == LCTO_NONE_U32) return 0;                                  /* UNMAPPED_WINDOW */
    uint32_t start = L->hap_reg_start[hap];
    uint32_t end = start + L->hap_n_windows[hap] * L->window;
    if (start <= middle && middle < end) return (middle - start) / L->window + shift;
    return 1;                                                               /* BOUNDARY_WINDOW */
}

/* src/model/assgn.rs:127-151 */
void lcto_apply_tweak(const lcto_locus *L, lcto_instance *I, lcto_rng *rng) {
    uint32_t tweak = L->tweak;
    for (uint32_t a = 0; a < I->n_alns; a++) {
        uint32_t pa = I->aln_pa[a];
        if (pa == LCTO_NONE_U32) continue;            /* parent == None: windows stay [0, 0] */
        uint32_t k = I->aln_contig_ix[a];
        uint32_t hap = I->haps[k], shift = I->wshift[k];
        uint32_t m1 = L->pa_mid1[pa], m2 = L->pa_mid2[pa];
        if (tweak == 0) {                             /* define_windows_determ, windows.rs:112-121 */
            I->aln_w[2 * a] = shifted_window_ix(L, hap, shift, m1);
            I->aln_w[2 * a + 1] = shifted_window_ix(L, hap, shift, m2);
        } else {                                      /* define_windows_random, windows.rs:123-136 */
            uint64_t r = lcto_rng_next_u64(rng);
            uint32_t t1 = (uint32_t)(r >> 32) % (2 * tweak + 1);
            uint32_t t2 = (uint32_t)r % (2 * tweak + 1);
            I->aln_w[2 * a] = shifted_window_ix(L, hap, shift, m1 == LCTO_NONE_U32 ? m1 : m1 + t1);
            I->aln_w[2 * a + 1] = shifted_window_ix(L, hap, shift, m2 == LCTO_NONE_U32 ? m2 : m2 + t2);
        }
    }
    /* depth_distrs.truncate(2); windows 0 and 1 stay TRIVIAL (assgn.rs:75-77,140) */
    I->win_trivial[0] = 1; I->win_trivial[1] = 1; I->win_weight[0] = 0.0; I->win_weight[1] = 0.0;
    uint32_t w = 2;
    for (uint32_t k = 0; k < I->ploidy; k++) {
        uint32_t hap = I->haps[k];
        uint32_t nwin = L->hap_n_windows[hap];
        for (uint32_t i = 0; i < nwin; i++) {
            /* generate_windows, windows.rs:478-486 */
            uint32_t start = L->hap_reg_start[hap] + i * L->window;
            uint32_t end = start + L->window;
            int32_t left_tweak = (int32_t)(tweak < start ? tweak : start);
            uint32_t rem = L->hap_len[hap] - end;
            int32_t right_tweak = (int32_t)(tweak < rem ? tweak : rem);
            int32_t rr = lcto_rng_range_i32_incl(rng, -left_tweak, right_tweak);
            uint32_t wstart = (uint32_t)((int32_t)start + rr);
            /* neighb_info, windows.rs:439-445: index saturating_sub(left_padding) */
            uint32_t idx = wstart > L->left_padding ? wstart - L->left_padding : 0;
            uint64_t pos = L->hap_pos_off[hap] + idx;
            assert(pos < L->hap_pos_off[hap + 1]);
            double weight = L->pos_weight[pos];
            uint8_t gc = L->pos_gc[pos];
            /* assgn.rs:144-148 + DistrCache::get_distribution, distr_cache.rs:83-92 */
            if (weight < L->min_weight || weight < 1e-7) {
                I->win_trivial[w] = 1; I->win_weight[w] = 0.0; I->win_gc[w] = 0;
            } else {
                I->win_trivial[w] = 0; I->win_weight[w] = weight; I->win_gc[w] = gc;
            }
            w++;
        }
    }
}

/* ------------------------------------------------------------ a8/a9: ReadAssignment */

typedef struct {
    const lcto_locus *L;
    const lcto_instance *I;
    uint16_t *read_assgn;
    uint32_t *depth;
    double aln_lik, depth_lik;
    double depth_contrib, aln_contrib;
} assgn_t;

typedef struct { uint32_t read_pair; uint16_t new_assgn; uint32_t old_ix, new_ix; } target_t;

/* WindowDistr::ln_prob, src/model/distr_cache.rs:34-39 */
static inline double win_ln_prob(const assgn_t *A, uint32_t w, uint32_t k) {
    const lcto_instance *I = A->I;
    if (I->win_trivial[w]) return 0.0;
    assert(k < A->L->depth_k);
    return I->win_weight[w] * A->L->depth_table[(size_t)I->win_gc[w] * A->L->depth_k + k];
}

/* src/model/assgn.rs:346-354 */
static void recalc_likelihood(assgn_t *A) {
    const lcto_instance *I = A->I;
    double s = 0.0;
    for (uint32_t w = 0; w < I->total_windows; w++) s = s + win_ln_prob(A, w, A->depth[w]);
    A->depth_lik = s;
    double t = 0.0;
    for (uint32_t r = 0; r < I->n_reads; r++) t = t + I->aln_ln_prob[I->read_ixs[r] + A->read_assgn[r]];
    A->aln_lik = t;
}

/* src/model/assgn.rs:235-237 */
static inline double likelihood(const assgn_t *A) {
    return A->depth_contrib * A->depth_lik + A->aln_contrib * A->aln_lik;
}

/* src/model/assgn.rs:244-254 */
static inline double atomic_depth_lik_diff(const assgn_t *A, uint32_t w, int32_t change) {
    if (change == 0) return 0.0;
    uint32_t old_depth = A->depth[w];
    uint32_t new_depth = (uint32_t)((int64_t)old_depth + change);
    return win_ln_prob(A, w, new_depth) - win_ln_prob(A, w, old_depth);
}

/* src/model/assgn.rs:259-284 */
static double depth_lik_diff(const assgn_t *A, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4) {
    int32_t c1 = -1, c2, c3, c4;
    if (w2 == w1) { c1 -= 1; c2 = 0; } else c2 = -1;
    if (w3 == w1) { c1 += 1; c3 = 0; } else if (w3 == w2) { c2 += 1; c3 = 0; } else c3 = 1;
    if (w4 == w1) { c1 += 1; c4 = 0; } else if (w4 == w2) { c2 += 1; c4 = 0; }
    else if (w4 == w3) { c3 += 1; c4 = 0; } else c4 = 1;
    return atomic_depth_lik_diff(A, w1, c1) + atomic_depth_lik_diff(A, w2, c2)
        + atomic_depth_lik_diff(A, w3, c3) + atomic_depth_lik_diff(A, w4, c4);
}

/* src/model/assgn.rs:287-317 */
static double best_read_improvement(const assgn_t *A, uint32_t rp, target_t *t) {
    const lcto_instance *I = A->I;
    uint32_t start_ix = I->read_ixs[rp], end_ix = I->read_ixs[rp + 1];
    assert(start_ix + 1 < end_ix);
    uint16_t old_assgn = A->read_assgn[rp];
    uint32_t old_ix = start_ix + old_assgn;
    uint32_t w1 = I->aln_w[2 * old_ix], w2 = I->aln_w[2 * old_ix + 1];
    uint32_t best_i = 0;
    double best_improv = -INFINITY;
    double rel_contrib = A->depth_contrib / A->aln_contrib;
    for (uint32_t i = 0; i < end_ix - start_ix; i++) {
        if (i != old_assgn) {
            uint32_t ix = start_ix + i;
            double improv = I->aln_ln_prob[ix]
                + rel_contrib * depth_lik_diff(A, w1, w2, I->aln_w[2 * ix], I->aln_w[2 * ix + 1]);
            if (improv > best_improv) { best_improv = improv; best_i = i; }
        }
    }
    t->read_pair = rp; t->new_assgn = (uint16_t)best_i; t->old_ix = old_ix; t->new_ix = start_ix + best_i;
    return A->aln_contrib * (best_improv - I->aln_ln_prob[old_ix]);
}

/* src/model/assgn.rs:321-328 */
static double calculate_improvement(const assgn_t *A, const target_t *t) {
    const lcto_instance *I = A->I;
    double d = depth_lik_diff(A, I->aln_w[2 * t->old_ix], I->aln_w[2 * t->old_ix + 1],
                              I->aln_w[2 * t->new_ix], I->aln_w[2 * t->new_ix + 1]);
    return A->depth_contrib * d + A->aln_contrib * (I->aln_ln_prob[t->new_ix] - I->aln_ln_prob[t->old_ix]);
}

/* src/model/assgn.rs:331-343 */
static void reassign(assgn_t *A, const target_t *t) {
    const lcto_instance *I = A->I;
    uint32_t w1 = I->aln_w[2 * t->old_ix], w2 = I->aln_w[2 * t->old_ix + 1];
    uint32_t w3 = I->aln_w[2 * t->new_ix], w4 = I->aln_w[2 * t->new_ix + 1];
    A->depth_lik += depth_lik_diff(A, w1, w2, w3, w4);
    A->aln_lik += I->aln_ln_prob[t->new_ix] - I->aln_ln_prob[t->old_ix];
    A->depth[w3] += 1; A->depth[w4] += 1; A->depth[w1] -= 1; A->depth[w2] -= 1;
    A->read_assgn[t->read_pair] = t->new_assgn;
}

/* ReassignmentTarget::random, src/model/assgn.rs:451-471 */
static void random_target(const assgn_t *A, lcto_rng *rng, target_t *t) {
    const lcto_instance *I = A->I;
    uint32_t rp = I->nontrivial[lcto_rng_range_usize(rng, 0, I->n_nontrivial)];
    uint32_t start_ix = I->read_ixs[rp], end_ix = I->read_ixs[rp + 1];
    uint32_t total = end_ix - start_ix;
    uint16_t old_assgn = A->read_assgn[rp];
    uint16_t new_assgn;
    if (total == 2) new_assgn = (uint16_t)(1 - old_assgn);
    else {
        uint16_t i = lcto_rng_range_u16(rng, 1, (uint16_t)total);
        new_assgn = (i <= old_assgn) ? (uint16_t)(i - 1) : i;
    }
    t->read_pair = rp; t->new_assgn = new_assgn;
    t->old_ix = start_ix + old_assgn; t->new_ix = start_ix + new_assgn;
}

/* ReadAssignment::try_new, src/model/assgn.rs:199-226.  init_mode: 0 = index 0 (best), 1 = random. */
static void assgn_init(assgn_t *A, int init_mode, lcto_rng *rng) {
    const lcto_instance *I = A->I;
    memset(A->depth, 0, sizeof(uint32_t) * I->total_windows);
    for (uint32_t r = 0; r < I->n_reads; r++) {
        uint32_t i = I->read_ixs[r], j = I->read_ixs[r + 1];
        uint32_t m = j - i, a = 0;
        if (m > 1 && init_mode == 1) a = (uint32_t)lcto_rng_range_usize(rng, 0, m);
        A->depth[I->aln_w[2 * (i + a)]] += 1;
        A->depth[I->aln_w[2 * (i + a) + 1]] += 1;
        A->read_assgn[r] = (uint16_t)a;
    }
    recalc_likelihood(A);
}

/* src/solvers/stoch.rs:19-22 */
static double max_abs_random(const assgn_t *A, lcto_rng *rng, int count) {
    double acc = 0.0;
    for (int i = 0; i < count; i++) {
        target_t t;
        random_target(A, rng, &t);
        acc = f64_max(acc, fabs(calculate_improvement(A, &t)));
    }
    return acc;
}

#define INIT_ITER 100                                   /* src/solvers/stoch.rs:24 */
static inline double minimum_allowed_diff(double m) {   /* src/solvers/stoch.rs:27-29 */
    return f64_max(1e-10 * m, 1e-14);
}

/* Greedy::solve_nontrivial, src/solvers/stoch.rs:81-120 */
static int greedy_solve(assgn_t *A, const lcto_stage *st, lcto_rng *rng, lcto_attempt_out *out) {
    const lcto_instance *I = A->I;
    uint64_t sample_size = st->sample_size < I->n_nontrivial ? st->sample_size : I->n_nontrivial;
    assgn_init(A, st->best_start ? 0 : 1, rng);
    double min_diff = minimum_allowed_diff(max_abs_random(A, rng, INIT_ITER));
    uint64_t curr_plato = 0;
    uint64_t max_iter = st->plato_size * 100 > 100000 ? st->plato_size * 100 : 100000;
    uint32_t *sample = (uint32_t *)malloc(sizeof(uint32_t) * (sample_size + 1));
    uint64_t it = 0, moves = 0;
    for (; it < max_iter; it++) {
        int have = 0;
        target_t best_t;
        double best_improv = min_diff;
        if (lcto_rng_sample_indices(rng, I->n_nontrivial, (uint32_t)sample_size, sample) != 0) { free(sample); return -1; }
        for (uint64_t s = 0; s < sample_size; s++) {
            target_t t;
            double improv = best_read_improvement(A, I->nontrivial[sample[s]], &t);
            if (improv > best_improv) { best_t = t; best_improv = improv; have = 1; }
        }
        if (have) { curr_plato = 0; reassign(A, &best_t); moves++; }
        else { curr_plato += 1; if (curr_plato > st->plato_size) { it++; break; } }
    }
    free(sample);
    out->iterations = it; out->moves = moves;
    return 0;
}

/* SimAnneal::solve_nontrivial, src/solvers/stoch.rs:197-242 */
static int anneal_solve(assgn_t *A, const lcto_stage *st, lcto_rng *rng, lcto_attempt_out *out) {
    assgn_init(A, 1, rng);
    double max_abs = max_abs_random(A, rng, INIT_ITER);
    double min_diff = minimum_allowed_diff(max_abs);
    double start_temp = f64_max(-max_abs / log(st->init_prob), 1e-5);
    double temp_step = start_temp / (double)st->anneal_steps;
    uint64_t curr_plato = 0, steps = 0, moves = 0;
    for (uint64_t i = st->anneal_steps; i >= 1; i--) {
        target_t t;
        random_target(A, rng, &t);
        double diff = calculate_improvement(A, &t) - min_diff;
        steps++;
        if (diff >= 0.0 || lcto_rng_f64(rng) <= exp(diff / (temp_step * (double)i))) {
            reassign(A, &t); curr_plato = 0; moves++;
        } else {
            curr_plato += 1;
            if (curr_plato >= st->plato_size) break;
        }
    }
    uint64_t max_iter = st->plato_size * 100 > 100000 ? st->plato_size * 100 : 100000;
    for (uint64_t k = 0; k < max_iter; k++) {
        if (curr_plato >= st->plato_size) break;
        target_t t;
        random_target(A, rng, &t);
        double diff = calculate_improvement(A, &t);
        steps++;
        if (diff > min_diff) { reassign(A, &t); curr_plato = 0; moves++; }
        else curr_plato += 1;
    }
    out->iterations = steps; out->moves = moves;
    return 0;
}

/* Solver::solve (src/solvers/mod.rs:61-72) + likelihood. */
int lcto_solve_attempt(const lcto_locus *L, const lcto_instance *I, const lcto_stage *st, lcto_rng *rng,
                       uint16_t *read_assgn, uint32_t *depth, lcto_attempt_out *out) {
    assgn_t A;
    A.L = L; A.I = I; A.read_assgn = read_assgn; A.depth = depth;
    A.aln_contrib = 1.0 - L->lik_skew;            /* src/model/assgn.rs:80-81 */
    A.depth_contrib = 1.0 + L->lik_skew;
    A.aln_lik = -INFINITY; A.depth_lik = -INFINITY;
    out->iterations = 0; out->moves = 0;
    int rc = 0;
    if (I->n_nontrivial == 0) assgn_init(&A, 0, rng);   /* trivial(): unique assignment, no RNG use */
    else if (st->kind == 0) rc = greedy_solve(&A, st, rng, out);
    else rc = anneal_solve(&A, st, rng, out);
    out->aln_lik = A.aln_lik; out->depth_lik = A.depth_lik; out->lik = likelihood(&A);
    return rc;
}

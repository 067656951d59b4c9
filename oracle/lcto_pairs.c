/*
 * lcto_pairs.c -- CPU ORACLE (test infrastructure only; see lcto.h) for SURVEY.md section 8(f) rank 1:
 * grouping the mate alignments of every read pair into per-contig pair alignments.
 *
 * Follows /root/reference:
 *   identify_paired_end_alignments  src/model/locs.rs:805-868   (per-contig grouping, <= max_alns per end,
 *                                                                read weight, unmapped_prob)
 *   identify_contig_pair_alns       src/model/locs.rs:744-799   (all first x second pairs of opposite strand,
 *                                                                single-mate options, top max_alns within prob_diff)
 *   Alignment::paired_prob          src/seq/aln.rs:236-238      ((ln1 + ln2) + insert ln-pmf)
 *   Alignment::insert_size          src/seq/aln.rs:223-226, Interval::furthest_distance src/seq/interv.rs:179-185
 *   Interval::middle                src/seq/interv.rs:154-156
 *   InsertDistr::{ln_prob, insert_penalty} src/bg/insertsz.rs:153-155,171-173 (table supplied by the caller)
 *   PairAlignment::{new_first,new_second,new_both} src/model/locs.rs:671-700
 *
 * Input order: the reference sorts the mates of a read by (contig, read end, ln_prob) with
 * sort_unstable_by (locs.rs:819-820) and then walks them contig ascending, first end before second end,
 * ln_prob descending; the caller of this function supplies them in that order, so no unstable sort is
 * restated.  The per-contig sort of the candidate pairs (locs.rs:793) IS restated, as a STABLE descending
 * sort by f64::total_cmp in generation order (Rust's sort_unstable on <= 20 elements is an insertion sort;
 * longer lists can only differ between exactly tied ln-probs -- "parity unpinned", like the rest of lcto).
 *   identify_single_end_alignments  src/model/locs.rs:870-911   (per contig: within prob_diff of the best, <= max_alns)
 *   ContigInfos::explicit_read_weight src/model/windows.rs:683-693, ContigInfo::read_end_weight :495-504
 */
#include "lcto.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double ln_prob; uint32_t mid1, mid2; } pair_cand;

static int64_t total_key(double v) {
    int64_t b;
    memcpy(&b, &v, 8);
    return b ^ (int64_t)(((uint64_t)(b >> 63)) >> 1);
}

/* identify_contig_pair_alns (locs.rs:744-799): mates [i, j) are first-end, [j, k) second-end. Returns the
 * number of pair alignments appended to out (<= max_alns), or -1 when an insert size is outside the table. */
static int contig_pairs(const lcto_mates *in, uint64_t i, uint64_t j, uint64_t k, double unm_ins_penalty,
                        pair_cand *cands, double *buffer, pair_cand *out) {
    size_t n = 0;
    for (uint64_t q = 0; q < k - j; q++) buffer[q] = -INFINITY;
    for (uint64_t ix1 = i; ix1 < j; ix1++) {
        double max_prob1 = -INFINITY;
        const uint32_t s1 = in->ma_start[ix1], e1 = in->ma_end[ix1];
        for (uint64_t ix2 = j; ix2 < k; ix2++) {
            if (((in->ma_flags[ix1] ^ in->ma_flags[ix2]) & 2u) == 0) continue;      /* same strand */
            const uint32_t s2 = in->ma_start[ix2], e2 = in->ma_end[ix2];
            const uint32_t insert = (e1 > e2 ? e1 : e2) - (s1 < s2 ? s1 : s2);
            if (insert >= in->ins_len) return -1;
            const double prob = (in->ma_ln_prob[ix1] + in->ma_ln_prob[ix2]) + in->ins_ln_pmf[insert];
            if (isfinite(prob)) {
                if (prob > max_prob1) max_prob1 = prob;
                if (prob > buffer[ix2 - j]) buffer[ix2 - j] = prob;
                cands[n].ln_prob = prob; cands[n].mid1 = (s1 + e1) / 2; cands[n].mid2 = (s2 + e2) / 2; n++;
            }
        }
        const double alone1 = in->ma_ln_prob[ix1] + unm_ins_penalty;
        if (alone1 >= max_prob1) { cands[n].ln_prob = alone1; cands[n].mid1 = (s1 + e1) / 2; cands[n].mid2 = LCTO_NONE_U32; n++; }
    }
    for (uint64_t ix2 = j; ix2 < k; ix2++) {
        const double alone2 = in->ma_ln_prob[ix2] + unm_ins_penalty;
        if (alone2 >= buffer[ix2 - j]) {
            cands[n].ln_prob = alone2; cands[n].mid1 = LCTO_NONE_U32;
            cands[n].mid2 = (in->ma_start[ix2] + in->ma_end[ix2]) / 2; n++;
        }
    }
    /* stable insertion sort, descending by total_cmp */
    for (size_t a = 1; a < n; a++) {
        pair_cand c = cands[a];
        size_t b = a;
        while (b > 0 && total_key(cands[b - 1].ln_prob) < total_key(c.ln_prob)) { cands[b] = cands[b - 1]; b--; }
        cands[b] = c;
    }
    const double thresh = cands[0].ln_prob - in->prob_diff;
    const size_t lim = n < in->max_alns ? n : in->max_alns;
    size_t keep = 0;
    while (keep < lim && cands[keep].ln_prob >= thresh) keep++;
    memcpy(out, cands, keep * sizeof(pair_cand));
    return (int)keep;
}

/* ContigInfo::read_end_weight (windows.rs:495-504) */
static double read_end_weight(const lcto_mates *in, uint32_t contig, uint32_t middle) {
    if (middle == LCTO_NONE_U32) return 0.0;
    const double *w = in->exp_weight + in->exp_off[contig];
    const uint32_t n = (uint32_t)(in->exp_off[contig + 1] - in->exp_off[contig]);
    const uint32_t u = in->window / 2;
    double v = w[middle];
    const double a = w[middle > u ? middle - u : 0u];           /* i.saturating_sub(u) */
    const double b = w[middle + u < n - 1 ? middle + u : n - 1];
    if (a > v) v = a;                                            /* f64::max */
    if (b > v) v = b;
    return v;
}

/* tail of identify_{paired,single}_end_alignments: weight = read weight * explicit_read_weight, scale, unmapped_prob */
static void finish_read(const lcto_mates *in, uint32_t r, uint64_t b, uint64_t e, const uint32_t *pa_contig,
                        double *pa_ln_prob, const uint32_t *pa_mid1, const uint32_t *pa_mid2, double *unmapped_prob) {
    double expl = 1.0;
    if (in->exp_weight) {                                        /* explicit_read_weight, windows.rs:683-693 */
        double s = 0.0;
        for (uint64_t q = b; q < e; q++) {
            const double w1 = read_end_weight(in, pa_contig[q], pa_mid1[q]), w2 = read_end_weight(in, pa_contig[q], pa_mid2[q]);
            s += w1 > w2 ? w1 : w2;
        }
        expl = s / (double)(e - b);
    }
    const double weight = (in->read_weight ? in->read_weight[r] : 1.0) * expl;
    for (uint64_t q = b; q < e; q++) pa_ln_prob[q] *= weight;                        /* locs.rs:861-863, 905-907 */
    unmapped_prob[r] = in->single_end ? weight * in->unmapped_penalty                 /* locs.rs:909 */
                                      : weight * (2.0 * in->unmapped_penalty + in->insert_penalty);   /* locs.rs:866 */
}

/* identify_single_end_alignments (locs.rs:870-911) for every read */
static int single_end_alignments(const lcto_mates *in, uint64_t cap, uint64_t *pa_off, uint32_t *pa_contig,
                                 double *pa_ln_prob, uint32_t *pa_mid1, uint32_t *pa_mid2, double *unmapped_prob) {
    uint64_t n_out = 0;
    for (uint32_t r = 0; r < in->n_reads; r++) {
        pa_off[r] = n_out;
        int have = 0;
        uint32_t curr_contig = 0, curr_saved = 0;
        double thresh = NAN;
        for (uint64_t a = in->ma_off[r]; a < in->ma_off[r + 1]; a++) {
            if (!have || curr_contig != in->ma_contig[a]) {
                have = 1; curr_contig = in->ma_contig[a];
                thresh = in->ma_ln_prob[a] - in->prob_diff;
                curr_saved = 0;
            }
            if (in->ma_ln_prob[a] >= thresh && curr_saved < (in->read_max_alns ? in->read_max_alns[r] : in->max_alns)) {
                if (n_out >= cap) return -3;
                pa_contig[n_out] = curr_contig; pa_ln_prob[n_out] = in->ma_ln_prob[a];
                pa_mid1[n_out] = (in->ma_start[a] + in->ma_end[a]) / 2; pa_mid2[n_out] = LCTO_NONE_U32;
                n_out++; curr_saved++;
            }
        }
        finish_read(in, r, pa_off[r], n_out, pa_contig, pa_ln_prob, pa_mid1, pa_mid2, unmapped_prob);
    }
    pa_off[in->n_reads] = n_out;
    return 0;
}

int lcto_pair_alignments(const lcto_mates *in, uint64_t cap, uint64_t *pa_off, uint32_t *pa_contig,
                         double *pa_ln_prob, uint32_t *pa_mid1, uint32_t *pa_mid2, double *unmapped_prob) {
    if (in->single_end) return single_end_alignments(in, cap, pa_off, pa_contig, pa_ln_prob, pa_mid1, pa_mid2, unmapped_prob);
    uint32_t Mmax = in->max_alns;                 /* per-read limits (recover_and_group_alignments, locs.rs:1263) */
    if (in->read_max_alns) for (uint32_t r = 0; r < in->n_reads; r++) Mmax = in->read_max_alns[r] > Mmax ? in->read_max_alns[r] : Mmax;
    pair_cand *cands = (pair_cand *)malloc(sizeof(pair_cand) * ((size_t)Mmax * Mmax + 2 * Mmax));
    pair_cand *kept = (pair_cand *)malloc(sizeof(pair_cand) * Mmax);
    double *buffer = (double *)malloc(sizeof(double) * Mmax);
    uint64_t *sel = (uint64_t *)malloc(sizeof(uint64_t) * 2 * Mmax);
    const double insert_penalty = in->insert_penalty;
    const double unm_ins_penalty = in->unmapped_penalty + insert_penalty;          /* locs.rs:815-816 */
    uint64_t n_out = 0;
    int rc = 0;
    for (uint32_t r = 0; r < in->n_reads && rc == 0; r++) {
        pa_off[r] = n_out;
        const uint32_t M = in->read_max_alns ? in->read_max_alns[r] : in->max_alns;
        uint64_t a = in->ma_off[r];
        const uint64_t end = in->ma_off[r + 1];
        while (a < end) {
            const uint32_t contig = in->ma_contig[a];
            uint64_t b = a;
            while (b < end && in->ma_contig[b] == contig) b++;
            /* keep at most max_alns alignments per end (locs.rs:842-851); first-end mates come first */
            uint64_t f = a;
            while (f < b && (in->ma_flags[f] & 1u) == 0) f++;
            const uint64_t n1 = (f - a) < M ? (f - a) : M, n2 = (b - f) < M ? (b - f) : M;
            /* gather the kept mates contiguously: indices a..a+n1 and f..f+n2 */
            for (uint64_t q = 0; q < n1; q++) sel[q] = a + q;
            for (uint64_t q = 0; q < n2; q++) sel[n1 + q] = f + q;
            /* contig_pairs works on index ranges; build a compact view */
            lcto_mates v = *in;
            v.max_alns = M;
            uint32_t st[2 * 64], en[2 * 64];
            uint8_t fl[2 * 64];
            double lp[2 * 64];
            if (M > 64) { rc = -1; break; }
            for (uint64_t q = 0; q < n1 + n2; q++) {
                st[q] = in->ma_start[sel[q]]; en[q] = in->ma_end[sel[q]];
                fl[q] = in->ma_flags[sel[q]]; lp[q] = in->ma_ln_prob[sel[q]];
            }
            v.ma_start = st; v.ma_end = en; v.ma_flags = fl; v.ma_ln_prob = lp;
            const int keep = contig_pairs(&v, 0, n1, n1 + n2, unm_ins_penalty, cands, buffer, kept);
            if (keep < 0) { rc = -2; break; }
            if (n_out + (uint64_t)keep > cap) { rc = -3; break; }
            for (int q = 0; q < keep; q++) {
                pa_contig[n_out] = contig;
                pa_ln_prob[n_out] = kept[q].ln_prob;
                pa_mid1[n_out] = kept[q].mid1;
                pa_mid2[n_out] = kept[q].mid2;
                n_out++;
            }
            a = b;
        }
        if (rc == 0) finish_read(in, r, pa_off[r], n_out, pa_contig, pa_ln_prob, pa_mid1, pa_mid2, unmapped_prob);
    }
    pa_off[in->n_reads] = n_out;
    free(cands); free(kept); free(buffer); free(sel);
    return rc;
}

/* ORACLE (test infrastructure; never imported, linked or executed by the product path).
 *
 * CPU restatement of what AllAlignments::load does with a read between read_next_alns and the pairing
 * (src/model/locs.rs:1117-1137) and of recover_and_group_alignments without the alignment transfer (:1237-1288;
 * `opt_hap_alns` = None -- the transfer needs WFA2, an un-vendored submodule of the reference), SURVEY.md section 8(f)
 * rank 1, remainder.  Parity unpinned by the reference itself (no tests / fixtures there, no Rust toolchain here);
 * pinned by a statement-by-statement Python transcription of the cited Rust lines and hand-checked cases
 * (tests/test_group.py).
 */
#include "lcto.h"
#include <math.h>

/* One entry of the pairing input; the comparison restates the two sorts the consumers apply before popping from the
 * back (identify_paired_end_alignments :819-820, identify_single_end_alignments :883): consumption order = contig
 * ascending, first read end before second, ln_prob descending.  The reference sorts unstably; equal keys keep their
 * order in PrelimAlignments::alns here. */
typedef struct { uint32_t contig, end, rec, ord; double lp; } ent_t;
static int ent_before(const ent_t *a, const ent_t *b) {
    if (a->contig != b->contig) return a->contig < b->contig;
    if (a->end != b->end) return a->end < b->end;
    if (a->lp != b->lp) return a->lp > b->lp;
    return a->ord < b->ord;
}

int lcto_group_reads(const lcto_prelim *in, uint64_t cap, uint8_t *status, uint64_t *n_reads_out, uint32_t *out_read,
                     uint8_t *out_max_alns, uint64_t *ma_off, uint32_t *ma_contig, uint8_t *ma_flags,
                     uint32_t *ma_start, uint32_t *ma_end, double *ma_ln_prob, uint32_t *ma_rec, uint64_t *counts) {
    uint64_t n_out = 0, n_ma = 0;
    counts[0] = counts[1] = counts[2] = 0;
    ma_off[0] = 0;
    for (uint64_t r = 0; r < in->n_reads; r++) {
        const int64_t g[2] = { in->read_group[2 * r], in->read_group[2 * r + 1] };
        /* load(): well_mapped = read_next_alns(First) [&& read_next_alns(Second)]   (:1119-1134) */
        int well = g[0] >= 0 && in->grp_ok[g[0]];
        if (!in->single_end && well) well = g[1] >= 0 && in->grp_ok[g[1]];
        if (!well) { status[r] = 1; counts[0]++; continue; }
        const int n_ends = in->single_end ? 1 : 2;
        /* in_bounds over PrelimAlignments::alns (:1008-1014, 1135) */
        int inb = 0;
        for (int e = 0; e < n_ends; e++) {
            const uint64_t b = in->grp_off[g[e]];
            for (uint32_t k = 0; k < in->grp_n_kept[g[e]]; k++) {
                const uint32_t rec = in->kept_rec[b + k];
                const uint32_t clen = in->contig_len[in->rec_contig[rec]];
                const uint32_t mid = (in->rec_start[rec] + in->rec_end[rec]) / 2;          /* Interval::middle */
                if (in->boundary <= mid && mid < clen - in->boundary) inb = 1;
            }
        }
        if (!inb) { status[r] = 2; counts[1]++; continue; }
        /* recover_and_group_alignments: best_edit_is_good (:293-295, 1257); an absent end keeps u32::MAX <= u32::MAX */
        int good = 1;
        for (int e = 0; e < n_ends; e++) good = good && in->grp_best_edit[g[e]] <= in->grp_thr_dist[g[e]];
        if (!good) { status[r] = 3; counts[0]++; continue; }
        status[r] = 0; counts[2]++;
        const uint64_t start = n_ma;
        for (int e = 0; e < n_ends; e++) {
            const uint64_t b = in->grp_off[g[e]], ge = in->grp_off[g[e] + 1];
            double best = -INFINITY;                                   /* best_lik: every pushed alignment (:311) */
            for (uint64_t q = b; q < ge; q++) best = in->rec_ln_prob[q] > best ? in->rec_ln_prob[q] : best;
            for (uint32_t k = 0; k < in->grp_n_kept[g[e]]; k++) {
                if (n_ma >= cap) return -3;
                const uint32_t rec = in->kept_rec[b + k];
                ent_t x = { in->rec_contig[rec], (uint32_t)e, rec, (uint32_t)(n_ma - start),
                            in->rec_ln_prob[rec] - best };             /* normalize_probs (:358-360) */
                /* insertion into the sorted segment [start, n_ma) */
                uint64_t pos = n_ma;
                while (pos > start) {
                    ent_t y = { ma_contig[pos - 1], (uint32_t)(ma_flags[pos - 1] & 1u), 0, 0, ma_ln_prob[pos - 1] };
                    y.ord = 0;                                         /* earlier entries always have a smaller ord */
                    if (!ent_before(&x, &y) ) break;
                    ma_contig[pos] = ma_contig[pos - 1]; ma_flags[pos] = ma_flags[pos - 1];
                    ma_start[pos] = ma_start[pos - 1]; ma_end[pos] = ma_end[pos - 1];
                    ma_ln_prob[pos] = ma_ln_prob[pos - 1]; ma_rec[pos] = ma_rec[pos - 1];
                    pos--;
                }
                ma_contig[pos] = x.contig;
                ma_flags[pos] = (uint8_t)(e | (in->rec_strand[rec] ? 2 : 0));
                ma_start[pos] = in->rec_start[rec]; ma_end[pos] = in->rec_end[rec];
                ma_ln_prob[pos] = x.lp; ma_rec[pos] = rec;
                n_ma++;
            }
        }
        out_read[n_out] = (uint32_t)r;
        out_max_alns[n_out] = in->read_weight[r] >= in->min_weight ? 10 : 2;   /* MAX_USED_ALNS / MAX_UNUSED_ALNS (:739-742, 1263) */
        n_out++;
        ma_off[n_out] = n_ma;
    }
    *n_reads_out = n_out;
    return 0;
}

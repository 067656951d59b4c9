/* ORACLE (test infrastructure; never imported, linked or executed by the product path).
 *
 * CPU restatement of the per-alignment part of PrelimAlignments::push (src/model/locs.rs:297-313), SURVEY.md
 * section 8(f) rank 2, first slice.  Parity unpinned by the reference itself (no tests / fixtures there, no Rust
 * toolchain here); pinned instead by a statement-by-statement Python transcription of the cited Rust lines and by
 * hand-checked cases (tests/test_rescore.py).
 */
#include "lcto.h"

/* Alignment::count_region_operations_fast (src/seq/aln.rs:298-317) + limited_clipping (:288-296) +
 * Cigar::soft_clipping (src/seq/cigar.rs:519-527) + OperCounts::edit_distance (src/bg/err_prof.rs:73-79) +
 * ErrorProfile::ln_prob (err_prof.rs:212-221) + `save` (locs.rs:308).
 * Returns 0, -1 on an empty CIGAR, -2 on an operation the reference panics on. */
int lcto_rescore_alignments(const lcto_alns *in, double *ln_prob, uint32_t *edit, uint32_t *read_len, uint8_t *save) {
    for (uint64_t i = 0; i < in->n_alns; i++) {
        const uint64_t b = in->cigar_off[i], e = in->cigar_off[i + 1];
        if (e <= b) return -1;                                    /* cigar.rs:520 assert */
        uint32_t matches = 0, mismatches = 0, insertions = 0, deletions = 0;
        for (uint64_t q = b; q < e; q++) {                        /* aln.rs:302-312 */
            const uint32_t oplen = in->cigar_ops[q] >> 4;
            switch (in->cigar_ops[q] & 15u) {
            case 7: matches += oplen; break;                      /* Operation::Equal */
            case 8: mismatches += oplen; break;                   /* Operation::Diff */
            case 2: deletions += oplen; break;                    /* Operation::Del */
            case 1: insertions += oplen; break;                   /* Operation::Ins */
            case 4: break;                                        /* Operation::Soft => {} */
            default: return -2;                                   /* panic!("Unsupported CIGAR operation") */
            }
        }
        const uint32_t first = in->cigar_ops[b], last = in->cigar_ops[e - 1];
        const uint32_t left = (first & 15u) == 4u ? first >> 4 : 0u;          /* cigar.rs:524 */
        const uint32_t right = (last & 15u) == 4u ? last >> 4 : 0u;           /* cigar.rs:525 */
        const uint32_t start = in->aln_start[i], end = in->aln_end[i], clen = in->contig_len[i];
        const uint32_t lim_left = left < start ? left : start;                 /* aln.rs:292 */
        const uint32_t room = clen > end ? clen - end : 0u;                    /* saturating_sub, aln.rs:294 */
        const uint32_t lim_right = right < room ? right : room;
        const uint32_t clipping = lim_left + lim_right;                        /* aln.rs:314-315 */
        const uint32_t common = mismatches + insertions + clipping;            /* err_prof.rs:74 */
        edit[i] = common + deletions;                                          /* :76 */
        read_len[i] = common + matches;                                        /* :77 */
        double lp = in->ln_match * (double)matches;                            /* err_prof.rs:216-220, left to right */
        lp = lp + in->ln_mismatch * (double)mismatches;
        lp = lp + in->ln_insertion * (double)insertions;
        lp = lp + in->ln_deletion * (double)deletions;
        lp = lp + in->ln_clipping * (double)clipping;
        ln_prob[i] = lp;
        save[i] = edit[i] <= in->passable_dist[i] ? 1 : 0;                     /* locs.rs:308 */
    }
    return 0;
}

/* ---- second slice: read_next_alns (src/model/locs.rs:502-567) + PrelimAlignments::push with the PosCollection
 * (:166-187, 297-343).  Plain restatement: the map is a linear list of (key, value) per group.  Pinned by a
 * statement-by-statement Python transcription (dict) in tests/test_rescore.py. */
#include <stdlib.h>
#include <math.h>

#define NOT_SAVED 0xFFFFFFFFu                                               /* locs.rs:209 */

int lcto_collect_read_ends(const lcto_read_ends *in, double *ln_prob, uint32_t *edit, uint32_t *read_len,
                           uint8_t *ok, uint32_t *best_edit, double *weight_factor, uint32_t *thr_dist,
                           uint32_t *pass_dist, uint32_t *n_kept, uint32_t *kept_rec) {
    const uint64_t n = in->alns.n_alns;
    uint8_t *save = (uint8_t *)malloc(n ? n : 1);
    lcto_alns a = in->alns;
    uint32_t *inf = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (uint64_t i = 0; i < n; i++) inf[i] = 0xFFFFFFFFu;
    a.passable_dist = inf;
    int rc = lcto_rescore_alignments(&a, ln_prob, edit, read_len, save);   /* the per-alignment part of push, :304-307 */
    free(save); free(inf);
    if (rc) return rc;
    uint64_t max_grp = 1;
    for (uint64_t g = 0; g < in->n_groups; g++)
        if (in->grp_off[g + 1] - in->grp_off[g] > max_grp) max_grp = in->grp_off[g + 1] - in->grp_off[g];
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * max_grp);
    uint32_t *vals = (uint32_t *)malloc(sizeof(uint32_t) * max_grp);
    for (uint64_t g = 0; g < in->n_groups; g++) {
        const uint64_t b = in->grp_off[g], e = in->grp_off[g + 1];
        const uint32_t rl = in->grp_read_len[g], good = in->grp_good_dist[g];
        uint32_t passable = in->grp_passable_dist[g], threshold = good;     /* locs.rs:529-530 */
        if (in->grp_neighb_complexity[g] <= in->poor_compl) {               /* :531-534 */
            double v = in->poor_compl_edit * (double)rl;
            uint32_t t = v != v || v <= 0.0 ? 0u : (v >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)v);   /* `as u32` */
            threshold = good > t ? good : t;
            passable += threshold - good;
        }
        thr_dist[g] = threshold; pass_dist[g] = passable;
        uint32_t be = 0xFFFFFFFFu, kept = 0;
        uint64_t n_map = 0;
        int failed = 0;
        for (uint64_t i = b; i < e && !failed; i++) {                       /* push, :297-343 */
            if (edit[i] < be) be = edit[i];                                 /* :308 */
            const uint32_t new_ix = kept;
            const int sv = edit[i] <= passable;                             /* :312 */
            if (new_ix == 0 && !sv) { failed = 1; break; }                  /* :314-316 (only the primary can get here) */
            const uint64_t key = ((uint64_t)in->grp_read_end[g] << 48) | ((uint64_t)in->rec_contig[i] << 32)
                               | (uint64_t)(in->alns.aln_start[i] >> 7);     /* encode, :166-168 */
            uint64_t q = 0;
            while (q < n_map && keys[q] != key) q++;
            if (q < n_map) {                                                /* Entry::Occupied */
                if (sv) {
                    if (vals[q] == NOT_SAVED) { vals[q] = new_ix; kept_rec[b + kept++] = (uint32_t)i; }       /* :322-325 */
                    else if (ln_prob[i] > ln_prob[kept_rec[b + vals[q]]]) kept_rec[b + vals[q]] = (uint32_t)i; /* :326-329 */
                }
            } else {                                                        /* Entry::Vacant */
                keys[n_map] = key;
                if (sv) { vals[n_map] = new_ix; kept_rec[b + kept++] = (uint32_t)i; }                        /* :333-336 */
                else vals[n_map] = NOT_SAVED;                                                                 /* :337-339 */
                n_map++;
            }
        }
        best_edit[g] = be;
        n_kept[g] = failed ? 0 : kept;
        const uint32_t req = in->strict_subset ? passable : threshold;      /* :560 */
        ok[g] = (!failed && be <= req) ? 1 : 0;                             /* :538-542, 561-563 */
        weight_factor[g] = (!ok[g] || be <= good) ? 1.0 : sqrt((double)good / (double)be);   /* :564 (only on success) */
    }
    free(keys); free(vals);
    return 0;
}

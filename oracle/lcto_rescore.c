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

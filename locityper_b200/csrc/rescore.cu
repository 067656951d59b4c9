// rescore.cu -- SURVEY.md section 8(f) rank 2, first slice: the per-alignment part of PrelimAlignments::push
// (src/model/locs.rs:297-313) for every alignment record of a locus in one launch:
//   counts    = Alignment::count_region_operations_fast(contig_len)        (src/seq/aln.rs:298-317)
//   clipping  = limited_clipping: soft clipping limited to the contig        (aln.rs:288-296, cigar.rs:519-527)
//   dist      = OperCounts::edit_distance()                                   (src/bg/err_prof.rs:73-79)
//   ln_prob   = ErrorProfile::ln_prob(&counts)                                (err_prof.rs:212-221)
//   save      = dist.edit() <= passable_dist[read_end]                        (locs.rs:308)
// One thread per alignment record walks its CIGAR (BAM-encoded u32 operations, CSR offsets); short-read CIGARs
// are a handful of operations, so this is streaming integer work bound by HBM: 4 B per operation + 24 B of record
// fields in, 17 B out per alignment.  The f64 expression keeps the reference's left-to-right association and is
// never contracted (-fmad=false, explicit _rn intrinsics).
#include "common.cuh"

#include <vector>

namespace lctp {

struct AlnsDev {
    uint64_t n;
    const uint64_t *cigar_off;
    const uint32_t *cigar_ops, *aln_start, *aln_end, *contig_len, *passable;
    double lm, lx, li, ld, lc;
};

__global__ void __launch_bounds__(256)
k_rescore(AlnsDev D, double *__restrict__ ln_prob, uint32_t *__restrict__ edit, uint32_t *__restrict__ read_len,
          uint8_t *__restrict__ save, int *__restrict__ err) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.n) return;
    const uint64_t b = D.cigar_off[i], e = D.cigar_off[i + 1];
    if (e <= b) { atomicOr(err, 1); return; }                    // soft_clipping asserts a non-empty CIGAR
    uint32_t matches = 0, mismatches = 0, insertions = 0, deletions = 0;
    uint32_t first = 0, last = 0;
    for (uint64_t q = b; q < e; q++) {
        const uint32_t t = D.cigar_ops[q], len = t >> 4;
        if (q == b) first = t;
        last = t;
        switch (t & 15u) {
        case 7: matches += len; break;                           // Operation::Equal
        case 8: mismatches += len; break;                        // Operation::Diff
        case 2: deletions += len; break;                         // Operation::Del
        case 1: insertions += len; break;                        // Operation::Ins
        case 4: break;                                           // Operation::Soft: counted through limited_clipping
        default: atomicOr(err, 2); break;                        // the reference panics
        }
    }
    const uint32_t left = (first & 15u) == 4u ? first >> 4 : 0u, right = (last & 15u) == 4u ? last >> 4 : 0u;
    const uint32_t start = D.aln_start[i], end = D.aln_end[i], clen = D.contig_len[i];
    const uint32_t clipping = min(left, start) + min(right, clen > end ? clen - end : 0u);   // saturating_sub
    const uint32_t common = mismatches + insertions + clipping;
    const uint32_t ed = common + deletions;
    double lp = __dmul_rn(D.lm, (double)matches);
    lp = __dadd_rn(lp, __dmul_rn(D.lx, (double)mismatches));
    lp = __dadd_rn(lp, __dmul_rn(D.li, (double)insertions));
    lp = __dadd_rn(lp, __dmul_rn(D.ld, (double)deletions));
    lp = __dadd_rn(lp, __dmul_rn(D.lc, (double)clipping));
    ln_prob[i] = lp;
    edit[i] = ed;
    read_len[i] = common + matches;
    if (D.passable) save[i] = ed <= D.passable[i] ? 1 : 0;
}

template <typename T>
static int put(DevBuf<T> &dst, const T *src, size_t n, cudaStream_t s) {
    int rc = dst.alloc(n ? n : 1);
    if (rc) return rc;
    if (n) LCTP_CUDA_CHECK(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return LCTP_OK;
}

int rescore_alignments(lctp_ctx *ctx, const lctp_alns *in, double *ln_prob, uint32_t *edit, uint32_t *read_len,
                       uint8_t *save) {
    cudaStream_t s = ctx->stream;
    const uint64_t n = in->n_alns;
    if (n == 0) return LCTP_OK;
    const uint64_t n_ops = in->cigar_off[n];
    DevBuf<uint64_t> d_off;
    DevBuf<uint32_t> d_ops, d_start, d_end, d_clen, d_pass, d_edit, d_rlen;
    DevBuf<double> d_lp;
    DevBuf<uint8_t> d_save;
    DevBuf<int> d_err;
    int rc;
    if ((rc = put(d_off, in->cigar_off, (size_t)n + 1, s))) return rc;
    if ((rc = put(d_ops, in->cigar_ops, (size_t)n_ops, s))) return rc;
    if ((rc = put(d_start, in->aln_start, (size_t)n, s))) return rc;
    if ((rc = put(d_end, in->aln_end, (size_t)n, s))) return rc;
    if ((rc = put(d_clen, in->contig_len, (size_t)n, s))) return rc;
    if ((rc = put(d_pass, in->passable_dist, (size_t)n, s))) return rc;
    if ((rc = d_lp.alloc(n)) || (rc = d_edit.alloc(n)) || (rc = d_rlen.alloc(n)) || (rc = d_save.alloc(n)) || (rc = d_err.alloc(1))) return rc;
    LCTP_CUDA_CHECK(cudaMemsetAsync(d_err.p, 0, sizeof(int), s));
    AlnsDev D;
    D.n = n; D.cigar_off = d_off.p; D.cigar_ops = d_ops.p; D.aln_start = d_start.p; D.aln_end = d_end.p;
    D.contig_len = d_clen.p; D.passable = d_pass.p;
    D.lm = in->ln_match; D.lx = in->ln_mismatch; D.li = in->ln_insertion; D.ld = in->ln_deletion; D.lc = in->ln_clipping;
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[0], s));
    k_rescore<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(D, d_lp.p, d_edit.p, d_rlen.p, d_save.p, d_err.p);
    ctx->launches++;
    LCTP_CUDA_CHECK(cudaGetLastError());
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[1], s));
    int err = 0;
    LCTP_CUDA_CHECK(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(ln_prob, d_lp.p, n * 8, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(edit, d_edit.p, n * 4, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(read_len, d_rlen.p, n * 4, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(save, d_save.p, n, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
    if (err) {
        set_error("lctp_rescore_alignments: malformed CIGAR (flags=%d: 1=empty CIGAR, 2=unsupported operation; only "
                  "I, D, S, =, X are accepted, like the reference)", err);
        return LCTP_E_INVALID;
    }
    float ms = 0.f;
    LCTP_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->stats.rescore_ms += ms;
    ctx->stats.rescore_launches += 1;
    ctx->stats.rescore_alns += n;
    ctx->stats.rescore_ops += n_ops;
    return LCTP_OK;
}

// ---- second slice: read_next_alns (src/model/locs.rs:502-567) + push with the PosCollection (:166-187, 315-343) ----
//
// One thread per (read, read end) group walks the group's records in order -- the de-duplication is order dependent:
// the first alignment whose start falls into a 128-bp bin of (read end, contig) claims the bin, a later one with a
// strictly larger ln-probability replaces it in place, a not-saved one (edit distance above `passable`) only marks the
// bin.  The bin map of a group is an open-addressing table in global scratch, 2 x the group's records rounded up to a
// power of two (8-byte key, 4-byte value).  The per-record part (CIGAR walk, ln-probability) is k_rescore.
struct EndsDev {
    uint64_t n_groups;
    const uint64_t *grp_off;
    const uint32_t *rec_contig, *aln_start;
    const uint8_t *read_end;
    const uint32_t *rlen, *good, *passable;
    const double *compl_;
    double poor_compl, poor_compl_edit;
    uint32_t strict_subset;
};
static constexpr uint32_t NOT_SAVED = 0xFFFFFFFFu;               // locs.rs:209
static constexpr uint64_t EMPTY_KEY = ~0ull;                     // no key has all bits set (read end < 2)

__host__ __device__ inline uint64_t table_cap(uint64_t n) {
    uint64_t c = 4;
    while (c < 2 * n) c <<= 1;
    return c;
}

__global__ void __launch_bounds__(128)
k_collect_read_ends(EndsDev D, const double *__restrict__ ln_prob, const uint32_t *__restrict__ edit,
                    const uint64_t *__restrict__ tab_off, uint64_t *__restrict__ tab_key, uint32_t *__restrict__ tab_val,
                    uint8_t *__restrict__ ok, uint32_t *__restrict__ best_edit, double *__restrict__ weight_factor,
                    uint32_t *__restrict__ thr_dist, uint32_t *__restrict__ pass_dist, uint32_t *__restrict__ n_kept,
                    uint32_t *__restrict__ kept_rec) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= D.n_groups) return;
    const uint64_t b = D.grp_off[g], e = D.grp_off[g + 1];
    const uint32_t good = D.good[g];
    uint32_t passable = D.passable[g], threshold = good;                       // locs.rs:529-530
    if (D.compl_[g] <= D.poor_compl) {                                         // :531-534
        const uint32_t t = __double2uint_rz(__dmul_rn(D.poor_compl_edit, (double)D.rlen[g]));   // `as u32`: truncating, saturating
        threshold = max(good, t);
        passable += threshold - good;
    }
    thr_dist[g] = threshold; pass_dist[g] = passable;
    const uint64_t cap = table_cap(e - b), mask = cap - 1;
    uint64_t *keys = tab_key + tab_off[g];
    uint32_t *vals = tab_val + tab_off[g];
    for (uint64_t q = 0; q < cap; q++) keys[q] = EMPTY_KEY;
    const uint64_t key_hi = ((uint64_t)D.read_end[g] << 48);
    uint32_t be = 0xFFFFFFFFu, kept = 0;
    bool failed = false;
    for (uint64_t i = b; i < e; i++) {                                         // push, :297-343
        const uint32_t ed = edit[i];
        be = min(be, ed);                                                      // :308
        const bool sv = ed <= passable;                                        // :312
        if (kept == 0 && !sv) { failed = true; break; }                        // :314-316: the primary is not good enough
        const uint64_t key = key_hi | ((uint64_t)D.rec_contig[i] << 32) | (uint64_t)(D.aln_start[i] >> 7);   // encode, :166-168
        uint64_t h = (key * 0x9E3779B97F4A7C15ull) >> 20 & mask;
        while (keys[h] != EMPTY_KEY && keys[h] != key) h = (h + 1) & mask;
        if (keys[h] == key) {                                                  // Entry::Occupied
            if (sv) {
                const uint32_t v = vals[h];
                if (v == NOT_SAVED) { vals[h] = kept; kept_rec[b + kept++] = (uint32_t)i; }               // :322-325
                else if (ln_prob[i] > ln_prob[kept_rec[b + v]]) kept_rec[b + v] = (uint32_t)i;           // :326-329
            }
        } else {                                                               // Entry::Vacant
            keys[h] = key;
            if (sv) { vals[h] = kept; kept_rec[b + kept++] = (uint32_t)i; }                               // :333-336
            else vals[h] = NOT_SAVED;                                                                     // :337-339
        }
    }
    best_edit[g] = be;
    n_kept[g] = failed ? 0u : kept;
    const uint32_t req = D.strict_subset ? passable : threshold;               // :560
    const bool good_end = !failed && be <= req;                                // :538-542, 561-563
    ok[g] = good_end ? 1 : 0;
    weight_factor[g] = (!good_end || be <= good) ? 1.0 : __dsqrt_rn(__ddiv_rn((double)good, (double)be));   // :564
}

int collect_read_ends(lctp_ctx *ctx, const lctp_read_ends *in, double *ln_prob, uint32_t *edit, uint32_t *read_len,
                      uint8_t *ok, uint32_t *best_edit, double *weight_factor, uint32_t *thr_dist, uint32_t *pass_dist,
                      uint32_t *n_kept, uint32_t *kept_rec) {
    cudaStream_t s = ctx->stream;
    const uint64_t n = in->alns.n_alns, ng = in->n_groups;
    if (n == 0 || ng == 0) return LCTP_OK;
    if (n >= 0xFFFFFFFFull) { set_error("lctp_collect_read_ends: too many alignment records"); return LCTP_E_CAPACITY; }
    if (in->grp_off[0] != 0 || in->grp_off[ng] != n) { set_error("lctp_collect_read_ends: grp_off does not cover the records"); return LCTP_E_INVALID; }
    const uint64_t n_ops = in->alns.cigar_off[n];
    std::vector<uint64_t> tab_off(ng + 1, 0);
    for (uint64_t g = 0; g < ng; g++) {
        if (in->grp_off[g + 1] <= in->grp_off[g]) { set_error("lctp_collect_read_ends: empty group %llu", (unsigned long long)g); return LCTP_E_INVALID; }
        if (in->grp_read_end[g] > 1) { set_error("lctp_collect_read_ends: read end must be 0 or 1"); return LCTP_E_INVALID; }
        tab_off[g + 1] = tab_off[g] + table_cap(in->grp_off[g + 1] - in->grp_off[g]);
    }
    DevBuf<uint64_t> d_off, d_goff, d_toff, d_tkey;
    DevBuf<uint32_t> d_ops, d_start, d_end, d_clen, d_edit, d_rlen, d_contig, d_grl, d_good, d_pass, d_tval, d_be, d_thr, d_pd, d_nk, d_kept;
    DevBuf<double> d_lp, d_compl, d_wf;
    DevBuf<uint8_t> d_re, d_ok;
    DevBuf<int> d_err;
    int rc;
    if ((rc = put(d_off, in->alns.cigar_off, (size_t)n + 1, s))) return rc;
    if ((rc = put(d_ops, in->alns.cigar_ops, (size_t)n_ops, s))) return rc;
    if ((rc = put(d_start, in->alns.aln_start, (size_t)n, s))) return rc;
    if ((rc = put(d_end, in->alns.aln_end, (size_t)n, s))) return rc;
    if ((rc = put(d_clen, in->alns.contig_len, (size_t)n, s))) return rc;
    if ((rc = put(d_contig, in->rec_contig, (size_t)n, s))) return rc;
    if ((rc = put(d_goff, in->grp_off, (size_t)ng + 1, s))) return rc;
    if ((rc = put(d_toff, tab_off.data(), (size_t)ng + 1, s))) return rc;
    if ((rc = put(d_re, in->grp_read_end, (size_t)ng, s))) return rc;
    if ((rc = put(d_grl, in->grp_read_len, (size_t)ng, s))) return rc;
    if ((rc = put(d_good, in->grp_good_dist, (size_t)ng, s))) return rc;
    if ((rc = put(d_pass, in->grp_passable_dist, (size_t)ng, s))) return rc;
    if ((rc = put(d_compl, in->grp_neighb_complexity, (size_t)ng, s))) return rc;
    if ((rc = d_lp.alloc(n)) || (rc = d_edit.alloc(n)) || (rc = d_rlen.alloc(n)) || (rc = d_err.alloc(1)) ||
        (rc = d_tkey.alloc(tab_off[ng])) || (rc = d_tval.alloc(tab_off[ng])) || (rc = d_ok.alloc(ng)) || (rc = d_be.alloc(ng)) ||
        (rc = d_wf.alloc(ng)) || (rc = d_thr.alloc(ng)) || (rc = d_pd.alloc(ng)) || (rc = d_nk.alloc(ng)) || (rc = d_kept.alloc(n))) return rc;
    LCTP_CUDA_CHECK(cudaMemsetAsync(d_err.p, 0, sizeof(int), s));
    LCTP_CUDA_CHECK(cudaMemsetAsync(d_kept.p, 0xFF, n * 4, s));
    AlnsDev A;
    A.n = n; A.cigar_off = d_off.p; A.cigar_ops = d_ops.p; A.aln_start = d_start.p; A.aln_end = d_end.p;
    A.contig_len = d_clen.p; A.passable = nullptr;
    A.lm = in->alns.ln_match; A.lx = in->alns.ln_mismatch; A.li = in->alns.ln_insertion; A.ld = in->alns.ln_deletion; A.lc = in->alns.ln_clipping;
    EndsDev E;
    E.n_groups = ng; E.grp_off = d_goff.p; E.rec_contig = d_contig.p; E.aln_start = d_start.p; E.read_end = d_re.p;
    E.rlen = d_grl.p; E.good = d_good.p; E.passable = d_pass.p; E.compl_ = d_compl.p;
    E.poor_compl = in->poor_compl; E.poor_compl_edit = in->poor_compl_edit; E.strict_subset = in->strict_subset;
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[0], s));
    k_rescore<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(A, d_lp.p, d_edit.p, d_rlen.p, nullptr, d_err.p);
    k_collect_read_ends<<<(unsigned)((ng + 127) / 128), 128, 0, s>>>(E, d_lp.p, d_edit.p, d_toff.p, d_tkey.p, d_tval.p, d_ok.p,
                                                                     d_be.p, d_wf.p, d_thr.p, d_pd.p, d_nk.p, d_kept.p);
    ctx->launches += 2;
    LCTP_CUDA_CHECK(cudaGetLastError());
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[1], s));
    int err = 0;
    LCTP_CUDA_CHECK(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(ln_prob, d_lp.p, n * 8, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(edit, d_edit.p, n * 4, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(read_len, d_rlen.p, n * 4, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(kept_rec, d_kept.p, n * 4, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(ok, d_ok.p, ng, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(best_edit, d_be.p, ng * 4, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(weight_factor, d_wf.p, ng * 8, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(thr_dist, d_thr.p, ng * 4, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(pass_dist, d_pd.p, ng * 4, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(n_kept, d_nk.p, ng * 4, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
    if (err) {
        set_error("lctp_collect_read_ends: malformed CIGAR (flags=%d: 1=empty CIGAR, 2=unsupported operation)", err);
        return LCTP_E_INVALID;
    }
    float ms = 0.f;
    LCTP_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->stats.rescore_ms += ms;
    ctx->stats.rescore_launches += 2;
    ctx->stats.rescore_alns += n;
    ctx->stats.rescore_ops += n_ops;
    ctx->stats.h2d_bytes += n_ops * 4 + n * 28 + ng * 37;
    ctx->stats.d2h_bytes += n * 20 + ng * 25;
    return LCTP_OK;
}

}  // namespace lctp
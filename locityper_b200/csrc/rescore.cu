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
    save[i] = ed <= D.passable[i] ? 1 : 0;
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

}  // namespace lctp

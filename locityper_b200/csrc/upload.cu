// upload.cu -- locus upload: H2D of the flat solve::Data, then on the device
//   * a1: the best-alignment matrix (AllAlignments::best_aln_matrix, src/model/locs.rs:1203-1212 via
//         best_for_each_contig :621-629), stored READ-major Mt[R][Hpad] so that prefilter tiles are
//         contiguous along haplotypes;
//   * the contig-major CSR of pair alignments (what GrouppedAlignments::contig_alns, locs.rs:614-618,
//         returns by bisection in the reference), so the per-genotype instance build reads
//         consecutive reads of one haplotype with coalesced loads.
#include "common.cuh"

#include <cub/device/device_scan.cuh>
#include <cstring>

namespace lctp {

__global__ void k_fill_mt(double *__restrict__ Mt, const double *__restrict__ unmapped, uint32_t R,
                          uint32_t H, uint32_t Hpad, int *__restrict__ err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)R * Hpad;
    if (i >= n) return;
    uint32_t r = (uint32_t)(i / Hpad), h = (uint32_t)(i % Hpad);
    Mt[i] = h < H ? unmapped[r] : 0.0;
    if (h == 0 && !(unmapped[r] <= 0.0)) atomicOr(err, 8);     // positive / NaN entry: see prefilter.cu dmax_nonpos
}

// One warp per read: walk its pair alignments (sorted by contig asc, ln_prob desc); the first entry
// of every (read, contig) run records the run length and the best ln-prob (matrix entry).
template <bool SCATTER>
__global__ void k_group_runs(const uint64_t *__restrict__ pa_off, const uint32_t *__restrict__ pa_contig,
                             const double *__restrict__ pa_lnprob, const uint32_t *__restrict__ pa_mid1,
                             const uint32_t *__restrict__ pa_mid2, uint32_t R, uint32_t H, uint32_t Hpad,
                             uint32_t *__restrict__ cnt_or_off, double *__restrict__ Mt,
                             double *__restrict__ cm_lnprob, uint2 *__restrict__ cm_mid,
                             int *__restrict__ err) {
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (warp >= R) return;
    uint32_t r = warp;
    uint64_t b = pa_off[r], e = pa_off[r + 1];
    for (uint64_t i = b + lane; i < e; i += 32) {
        uint32_t h = pa_contig[i];
        if (h >= H) { atomicOr(err, 1); continue; }
        bool first = (i == b) || (pa_contig[i - 1] != h);
        if (i > b && pa_contig[i - 1] > h) atomicOr(err, 2);
        if (!first) continue;
        uint32_t len = 1;
        while (i + len < e && pa_contig[i + len] == h) len++;
        size_t key = (size_t)h * R + r;
        if (!SCATTER) {
            if (len > 15) atomicOr(err, 16);        // the solver packs the rank inside a run into 4 bits
            cnt_or_off[key] = len;
            Mt[(size_t)r * Hpad + h] = pa_lnprob[i];
            if (!(pa_lnprob[i] <= 0.0)) atomicOr(err, 8);
        } else {
            uint32_t o = cnt_or_off[key];
            for (uint32_t t = 0; t < len; t++) {
                cm_lnprob[o + t] = pa_lnprob[i + t];
                cm_mid[o + t] = make_uint2(pa_mid1[i + t], pa_mid2[i + t]);
                if (t > 0 && pa_lnprob[i + t] > pa_lnprob[i + t - 1]) atomicOr(err, 4);
            }
        }
    }
}

static thread_local uint64_t g_h2d_bytes = 0;     // bytes of the upload in progress on this thread (lctp_stats.h2d_bytes)

template <typename T>
static int h2d(DevBuf<T> &dst, const T *src, size_t n, cudaStream_t s) {
    int rc = dst.alloc(n);
    if (rc) return rc;
    g_h2d_bytes += n * sizeof(T);
    if (n) LCTP_CUDA_CHECK(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return LCTP_OK;
}

// `pairs`: the pa_* / unmapped_prob section comes from lctp_pair_alignments_dev, already on the device (the fields of
// `in` are ignored for it); nullptr = copy it from the host arrays of `in`.
int upload_locus(lctp_ctx *ctx, const lctp_locus *in, lctp_locus_h *h, const lctp_pairs_h *pairs) {
    cudaStream_t s = ctx->stream;
    const uint32_t H = in->n_haps, R = in->n_reads, p = in->ploidy;
    if (H == 0 || R == 0 || p == 0 || in->n_genotypes == 0) {
        set_error("lctp_locus_upload: empty locus (H=%u R=%u p=%u G=%llu)", H, R, p,
                  (unsigned long long)in->n_genotypes);
        return LCTP_E_INVALID;
    }
    if (p > LCTP_MAX_PLOIDY) {
        set_error("lctp_locus_upload: ploidy %u > %d unsupported", p, LCTP_MAX_PLOIDY);
        return LCTP_E_CAPACITY;
    }
    if ((!pairs && (!in->unmapped_prob || !in->pa_off)) || !in->hap_len || !in->hap_n_windows || !in->hap_reg_start ||
        !in->hap_pos_off || !in->pos_weight || !in->pos_gc || !in->depth_table) {
        set_error("lctp_locus_upload: NULL input array");
        return LCTP_E_INVALID;
    }
    const uint64_t npa = pairs ? pairs->n_pairs : in->pa_off[R];
    if (npa >= 0xFFFFFFF0ull || (uint64_t)H * R >= 0xFFFFFFF0ull) {
        set_error("lctp_locus_upload: too many pair alignments (%llu) or H*R too large", (unsigned long long)npa);
        return LCTP_E_CAPACITY;
    }
    if (!pairs && npa && (!in->pa_contig || !in->pa_ln_prob || !in->pa_mid1 || !in->pa_mid2)) {
        set_error("lctp_locus_upload: NULL pair-alignment array");
        return LCTP_E_INVALID;
    }
    if ((uint64_t)in->depth_k < 2ull * R + 1) {
        set_error("lctp_locus_upload: depth_k=%u < 2R+1=%llu (depth table must cover every reachable depth)",
                  in->depth_k, 2ull * R + 1);
        return LCTP_E_INVALID;
    }
    if (!(in->lik_skew > -1.0 && in->lik_skew < 1.0) || in->window == 0) {
        set_error("lctp_locus_upload: invalid params (lik_skew=%g window=%u)", in->lik_skew, in->window);
        return LCTP_E_INVALID;
    }
    // geometry checks: every window start shifted by +-tweak must index inside pos arrays
    uint64_t total_w = 2;
    uint32_t max_nw = 0;
    for (uint32_t k = 0; k < H; k++) {
        uint32_t nw = in->hap_n_windows[k];
        if (nw > max_nw) max_nw = nw;
        uint64_t reg_end = (uint64_t)in->hap_reg_start[k] + (uint64_t)nw * in->window;
        uint64_t plen = in->hap_pos_off[k + 1] - in->hap_pos_off[k];
        if (reg_end > in->hap_len[k]) {
            set_error("lctp_locus_upload: haplotype %u windows exceed contig length", k);
            return LCTP_E_INVALID;
        }
        if (nw) {
            uint64_t last_start = reg_end - in->window;
            uint64_t right = in->hap_len[k] - reg_end;
            uint64_t max_start = last_start + (in->tweak < right ? in->tweak : right);
            uint64_t idx = max_start > in->left_padding ? max_start - in->left_padding : 0;
            if (idx >= plen) {
                set_error("lctp_locus_upload: haplotype %u pos arrays too short (%llu <= %llu)", k,
                          (unsigned long long)plen, (unsigned long long)idx);
                return LCTP_E_INVALID;
            }
        }
    }
    total_w += (uint64_t)p * max_nw;
    if (total_w > 65535) {   // windows are packed as u16 pairs on the device
        set_error("lctp_locus_upload: %llu windows per genotype exceed 65535", (unsigned long long)total_w);
        return LCTP_E_CAPACITY;
    }

    h->ctx = ctx;
    h->host = *in;
    h->npa = npa;
    h->max_n_windows = max_nw;
    h->hap_n_windows.assign(in->hap_n_windows, in->hap_n_windows + H);
    if (in->gt_tuples) h->gt_tuples_host.assign(in->gt_tuples, in->gt_tuples + (size_t)in->n_genotypes * p);
    if (in->priors) h->priors_host.assign(in->priors, in->priors + (size_t)in->n_genotypes);
    if (!pairs) h->unmapped_host.assign(in->unmapped_prob, in->unmapped_prob + R);
    const uint32_t Hpad = (H + 63u) & ~63u;

    int rc;
    g_h2d_bytes = 0;
    DevBuf<uint64_t> d_pa_off;
    DevBuf<uint32_t> d_pa_contig, d_mid1, d_mid2;
    DevBuf<double> d_pa_lnprob;
    DevBuf<int> d_err;
    const uint64_t *p_off; const uint32_t *p_contig, *p_mid1, *p_mid2; const double *p_lnprob;
    if (pairs) {
        p_off = pairs->pa_off.p; p_contig = pairs->contig.p; p_lnprob = pairs->lnprob.p;
        p_mid1 = pairs->mid1.p; p_mid2 = pairs->mid2.p;
        if ((rc = h->unmapped.alloc(R))) return rc;
        LCTP_CUDA_CHECK(cudaMemcpyAsync(h->unmapped.p, pairs->unmapped.p, (size_t)R * 8, cudaMemcpyDeviceToDevice, s));
        h->unmapped_host.resize(R);                    // count_unexplained_reads reads it on the host (R x 8 bytes)
        LCTP_CUDA_CHECK(cudaMemcpyAsync(h->unmapped_host.data(), pairs->unmapped.p, (size_t)R * 8, cudaMemcpyDeviceToHost, s));
        ctx->stats.d2h_bytes += (uint64_t)R * 8;
    } else {
        if ((rc = h2d(d_pa_off, in->pa_off, (size_t)R + 1, s))) return rc;
        if ((rc = h2d(d_pa_contig, in->pa_contig, npa, s))) return rc;
        if ((rc = h2d(d_pa_lnprob, in->pa_ln_prob, npa, s))) return rc;
        if ((rc = h2d(d_mid1, in->pa_mid1, npa, s))) return rc;
        if ((rc = h2d(d_mid2, in->pa_mid2, npa, s))) return rc;
        if ((rc = h2d(h->unmapped, in->unmapped_prob, R, s))) return rc;
        p_off = d_pa_off.p; p_contig = d_pa_contig.p; p_lnprob = d_pa_lnprob.p; p_mid1 = d_mid1.p; p_mid2 = d_mid2.p;
    }
    if ((rc = h2d(h->hap_len, in->hap_len, H, s))) return rc;
    if ((rc = h2d(h->hap_nw, in->hap_n_windows, H, s))) return rc;
    if ((rc = h2d(h->hap_rs, in->hap_reg_start, H, s))) return rc;
    if ((rc = h2d(h->hap_pos_off, in->hap_pos_off, (size_t)H + 1, s))) return rc;
    const uint64_t npos = in->hap_pos_off[H];
    if ((rc = h2d(h->pos_weight, in->pos_weight, npos, s))) return rc;
    if ((rc = h2d(h->pos_gc, in->pos_gc, npos, s))) return rc;
    {   // device table = [101][K] + one all-zero row used by TRIVIAL windows
        const size_t nt = (size_t)LCTP_GC_BINS * in->depth_k;
        if ((rc = h->depth_table.alloc(nt + in->depth_k))) return rc;
        LCTP_CUDA_CHECK(cudaMemcpyAsync(h->depth_table.p, in->depth_table, nt * 8, cudaMemcpyHostToDevice, s));
        g_h2d_bytes += nt * 8;
        LCTP_CUDA_CHECK(cudaMemsetAsync(h->depth_table.p + nt, 0, (size_t)in->depth_k * 8, s));
    }
    if (in->gt_tuples) {
        if ((rc = h2d(h->gt_tuples, in->gt_tuples, (size_t)in->n_genotypes * p, s))) return rc;
    }
    if (in->priors) {
        if ((rc = h2d(h->priors, in->priors, (size_t)in->n_genotypes, s))) return rc;
    }
    if ((rc = d_err.alloc(1))) return rc;
    LCTP_CUDA_CHECK(cudaMemsetAsync(d_err.p, 0, sizeof(int), s));

    const size_t n_keys = (size_t)H * R + 1;
    DevBuf<uint32_t> d_cnt;
    if ((rc = d_cnt.alloc(n_keys))) return rc;
    if ((rc = h->cm_off.alloc(n_keys))) return rc;
    if ((rc = h->cm_lnprob.alloc(npa + R))) return rc;          // + the unmapped probabilities (see LocusDev::npa)
    if ((rc = h->cm_mid.alloc(npa ? npa : 1))) return rc;
    if ((rc = h->Mt.alloc((size_t)R * Hpad))) return rc;
    if ((rc = h->scores.alloc((size_t)in->n_genotypes))) return rc;
    LCTP_CUDA_CHECK(cudaMemsetAsync(d_cnt.p, 0, n_keys * sizeof(uint32_t), s));

    {
        size_t n = (size_t)R * Hpad;
        k_fill_mt<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->Mt.p, h->unmapped.p, R, H, Hpad, d_err.p);
        ctx->launches++;
    }
    const unsigned warps_per_block = 8;
    const unsigned grid = (R + warps_per_block - 1) / warps_per_block;
    k_group_runs<false><<<grid, warps_per_block * 32, 0, s>>>(p_off, p_contig, p_lnprob, p_mid1,
                                                               p_mid2, R, H, Hpad, d_cnt.p, h->Mt.p, nullptr,
                                                               nullptr, d_err.p);
    ctx->launches++;
    size_t tmp_bytes = 0;
    LCTP_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt.p, h->cm_off.p, (int)n_keys, s));
    DevBuf<unsigned char> d_tmp;
    if ((rc = d_tmp.alloc(tmp_bytes))) return rc;
    LCTP_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp_bytes, d_cnt.p, h->cm_off.p, (int)n_keys, s));
    ctx->launches++;
    k_group_runs<true><<<grid, warps_per_block * 32, 0, s>>>(p_off, p_contig, p_lnprob, p_mid1,
                                                              p_mid2, R, H, Hpad, h->cm_off.p, nullptr,
                                                              h->cm_lnprob.p, h->cm_mid.p, d_err.p);
    ctx->launches++;
    LCTP_CUDA_CHECK(cudaMemcpyAsync(h->cm_lnprob.p + npa, h->unmapped.p, (size_t)R * 8, cudaMemcpyDeviceToDevice, s));
    LCTP_CUDA_CHECK(cudaGetLastError());

    // per-haplotype number of pair alignments = cm_off[(k+1)R] - cm_off[kR]: strided D2H of H+1 words
    std::vector<uint32_t> bounds(H + 1);
    LCTP_CUDA_CHECK(cudaMemcpy2DAsync(bounds.data(), sizeof(uint32_t), h->cm_off.p, (size_t)R * sizeof(uint32_t),
                                      sizeof(uint32_t), H + 1, cudaMemcpyDeviceToHost, s));
    int err = 0;
    LCTP_CUDA_CHECK(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
    ctx->stats.h2d_bytes += g_h2d_bytes;
    ctx->stats.d2h_bytes += (uint64_t)(H + 1) * 4 + 4;
    h->mt_nonpositive = !(err & 8);
    h->max_run = (err & 16) ? 16 : 15;           // > 15: lctp_solve_stage refuses (prefilter still works)
    err &= 7;
    if (err) {
        set_error("lctp_locus_upload: malformed pair alignments (flags=%d: 1=contig id >= H, 2=contigs not "
                  "ascending within a read, 4=ln_prob not descending within a (read, contig) run)", err);
        return LCTP_E_INVALID;
    }
    h->hap_alns.resize(H);
    h->max_hap_alns = 0;
    for (uint32_t k = 0; k < H; k++) {
        h->hap_alns[k] = bounds[k + 1] - bounds[k];
        if (h->hap_alns[k] > h->max_hap_alns) h->max_hap_alns = h->hap_alns[k];
    }

    LocusDev &d = h->dev;
    d.H = H; d.R = R; d.p = p; d.Hpad = Hpad; d.G = in->n_genotypes; d.npa = (uint32_t)npa;
    d.window = in->window; d.left_padding = in->left_padding; d.tweak = in->tweak; d.depth_k = in->depth_k;
    d.prob_diff = in->prob_diff;
    d.aln_contrib = 1.0 - in->lik_skew;                 // src/model/assgn.rs:80-81
    d.depth_contrib = 1.0 + in->lik_skew;
    d.rel_contrib = d.depth_contrib / d.aln_contrib;    // src/model/assgn.rs:299
    d.min_weight = in->min_weight;
    // exact for every 32-bit dividend (Lemire et al., "Faster remainder by direct computation"); a divisor of 1 wraps c to 0,
    // handled where it is used (window >= 2 is validated above; span = 1 means tweak = 0, where the remainder is not taken)
    d.c_window = ~0ull / (in->window ? in->window : 1u) + 1ull;
    d.c_span = ~0ull / (2ull * in->tweak + 1ull) + 1ull;
    d.Mt = h->Mt.p; d.unmapped = h->unmapped.p; d.cm_off = h->cm_off.p; d.cm_lnprob = h->cm_lnprob.p;
    d.cm_mid = h->cm_mid.p; d.hap_len = h->hap_len.p; d.hap_n_windows = h->hap_nw.p;
    d.hap_reg_start = h->hap_rs.p; d.hap_pos_off = h->hap_pos_off.p; d.pos_weight = h->pos_weight.p;
    d.pos_gc = h->pos_gc.p; d.depth_table = h->depth_table.p;
    d.gt_tuples = in->gt_tuples ? h->gt_tuples.p : nullptr;
    d.priors = in->priors ? h->priors.p : nullptr;
    // no caller pointer is retained after return
    h->host.gt_tuples = nullptr; h->host.priors = nullptr; h->host.unmapped_prob = nullptr;
    h->host.pa_off = nullptr; h->host.pa_contig = nullptr; h->host.pa_ln_prob = nullptr;
    h->host.pa_mid1 = nullptr; h->host.pa_mid2 = nullptr; h->host.hap_len = nullptr;
    h->host.hap_n_windows = nullptr; h->host.hap_reg_start = nullptr; h->host.hap_pos_off = nullptr;
    h->host.pos_weight = nullptr; h->host.pos_gc = nullptr; h->host.depth_table = nullptr;
    return LCTP_OK;
}

}  // namespace lctp

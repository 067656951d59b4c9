// prefilter.cu -- a2: run_filter scores (src/solvers/solve.rs:105-119)
//     score[g] = prior[g] + sum_{r<R} max_{k<p} M[h_k(g)][r]      (f64, summed in read order)
//
// Diploid full-triangle case: a tiled all-pairs (max,+) contraction.  One CTA owns a TBxTB tile of
// haplotype pairs (i <= j), streams the R dimension through shared memory with a 2-stage cp.async
// pipeline (Mt is read-major so a tile row is one contiguous 16-byte-aligned segment), and every thread
// keeps a TIxTJ register tile of f64 accumulators.  The R loop is never split, so every genotype is
// summed strictly in read order r = 0..R-1 exactly like the reference's `iter().sum()` -- the scores are
// bit-identical to the CPU path, which is what makes the survivor ORDER (and therefore the RNG stream
// each genotype later sees) reproducible.  (max,+) is not a ring the tensor cores implement; the bound
// is the FP64 pipe (1 DSETP + 1 DADD per genotype-read) and shared-memory bandwidth, not HBM.
//
// Any other ploidy / an explicit genotype list (`--priors`) uses the gather kernel: one thread per
// genotype, p row reads per read index.
#include "common.cuh"

#include <cstdlib>

namespace lctp {

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- TMA bulk copies (cp.async.bulk, completion on an mbarrier) ----------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
                 "r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
// max of two values that are both <= +0.0 (ln-probabilities): their order is the reverse order of their
// bit patterns as unsigned integers, so the comparison runs on the integer pipe and leaves the FP64
// pipe to the additions.  Only used when the upload verified that no matrix entry is positive or NaN.
__device__ __forceinline__ double dmax_nonpos(double a, double b) {
    return (unsigned long long)__double_as_longlong(a) < (unsigned long long)__double_as_longlong(b) ? a : b;
}

// genotype id of the pair (i <= j) in gen_combinations_with_repl order (src/ext/vec.rs:298-339)
__device__ __host__ __forceinline__ uint64_t pair_gid(uint64_t i, uint64_t j, uint64_t H) {
    return i * H - (i * (i - 1)) / 2 + (j - i);
}

// Thread layout: NTY x NTX threads, each owning a TI x TJ register tile of accumulators, so a CTA covers
// a (NTY*TI) x (NTX*TJ) tile of haplotype pairs.  Per read and thread: TI + TJ shared-memory values
// (16-byte loads) feed TI*TJ (DSETP, 2xFSEL, DADD) groups; the FP64 pipe (2 instructions per
// genotype-read, 64 lanes per clock per SM) is the binding unit once TI*TJ >= 16.
template <int TI, int TJ, int NTY, int NTX, int RC, bool IMAX, bool BULK = false>
__global__ void __launch_bounds__(NTY * NTX)
k_prefilter_pairs(const double *__restrict__ Mt, uint32_t R, uint32_t H, uint32_t Hpad,
                  const double *__restrict__ priors, double *__restrict__ scores, uint32_t nb,
                  uint64_t g_begin, uint64_t g_end) {
    constexpr int TBI = NTY * TI, TBJ = NTX * TJ, NT = NTY * NTX;
    static_assert(TBI == TBJ, "square CTA tiles (the triangle enumeration assumes it)");
    static_assert(TI % 2 == 0 || TI == 1, "TI");
    __shared__ __align__(128) double sA[2][RC][TBI];
    __shared__ __align__(128) double sB[2][RC][TBJ];
    __shared__ __align__(8) uint64_t full_bar[2];      // BULK: one "tile landed" barrier per stage

    // linear tile id -> (bi <= bj)
    uint32_t t = blockIdx.x, bi = 0;
    while (t >= nb - bi) { t -= nb - bi; bi++; }
    const uint32_t bj = bi + t;
    const uint32_t i0 = bi * TBI, j0 = bj * TBJ;
    {   // skip tiles entirely outside the requested genotype range (multi-GPU shards are row ranges)
        uint32_t il = min(i0 + TBI - 1, H - 1), jl = min(j0 + TBJ - 1, H - 1);
        uint64_t gmin = pair_gid(i0, max(i0, j0), H), gmax = pair_gid(il, jl, H);
        if (gmax < g_begin || gmin >= g_end) return;
    }
    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    // tile column of register column b of thread tx
    auto col_of = [](int tx_, int b) { return TJ % 2 == 0 ? (b / 2) * (2 * NTX) + tx_ * 2 + (b & 1) : tx_ * TJ + b; };

    double acc[TI][TJ];
#pragma unroll
    for (int a = 0; a < TI; a++)
#pragma unroll
        for (int b = 0; b < TJ; b++) acc[a][b] = 0.0;

    const int n_chunks = (R + RC - 1) / RC;
    if constexpr (BULK) {
        if (tid == 0) { mbar_init(&full_bar[0], 1); mbar_init(&full_bar[1], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
    }
    // BULK: every tile row of a chunk (TBI contiguous doubles of one read) is one cp.async.bulk issued by
    // one thread; the copies complete on the stage's mbarrier, so staging costs ~2 instructions per
    // thread per chunk and no LSU work.
    auto issue_bulk = [&](int c, int stage) {
        const uint32_t r0 = c * RC;
        const uint32_t nr = min((uint32_t)RC, R - r0);
        if (tid == 0) mbar_expect_tx(&full_bar[stage], nr * (TBI + TBJ) * 8u);
        for (int q = tid; q < 2 * RC; q += NT) {
            const int rc = q % RC;
            if ((uint32_t)rc < nr) {
                if (q < RC) bulk_g2s(&sA[stage][rc][0], Mt + (size_t)(r0 + rc) * Hpad + i0, TBI * 8u, &full_bar[stage]);
                else bulk_g2s(&sB[stage][rc][0], Mt + (size_t)(r0 + rc) * Hpad + j0, TBJ * 8u, &full_bar[stage]);
            }
        }
    };
    auto issue = [&](int c, int stage) {
        if constexpr (BULK) { issue_bulk(c, stage); return; }
        const uint32_t r0 = c * RC;
        constexpr int CH_A = RC * TBI / 2, CH_B = RC * TBJ / 2;   // 16-byte chunks
        for (int q = tid; q < CH_A; q += NT) {
            int rc = q / (TBI / 2), col = (q % (TBI / 2)) * 2;
            if (r0 + rc < R) cp_async16(&sA[stage][rc][col], Mt + (size_t)(r0 + rc) * Hpad + i0 + col);
        }
        for (int q = tid; q < CH_B; q += NT) {
            int rc = q / (TBJ / 2), col = (q % (TBJ / 2)) * 2;
            if (r0 + rc < R) cp_async16(&sB[stage][rc][col], Mt + (size_t)(r0 + rc) * Hpad + j0 + col);
        }
        cp_async_commit();
    };

    issue(0, 0);
    for (int c = 0; c < n_chunks; c++) {
        const int stage = c & 1;
        if constexpr (BULK) {
            if (c + 1 < n_chunks) issue(c + 1, stage ^ 1);
            mbar_wait(&full_bar[stage], (unsigned)(c >> 1) & 1u);
        } else {
            if (c + 1 < n_chunks) { issue(c + 1, stage ^ 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncthreads();
        }
        const int nr = min((int)RC, (int)(R - c * RC));
#pragma unroll 2
        for (int rc = 0; rc < nr; rc++) {
            double ai[TI], bj_[TJ];
            if constexpr (TI % 2 == 0) {
#pragma unroll
                for (int a = 0; a < TI; a += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(&sA[stage][rc][ty * TI + a]);
                    ai[a] = v.x; ai[a + 1] = v.y;
                }
            } else ai[0] = sA[stage][rc][ty];
            if constexpr (TJ % 2 == 0) {
                // column pairs are interleaved across the threads (thread tx owns pairs tx, tx + NTX, ...):
                // a quarter-warp reads 8 consecutive 16-byte chunks, so the loads are bank-conflict free
#pragma unroll
                for (int b = 0; b < TJ; b += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(&sB[stage][rc][col_of(tx, b)]);
                    bj_[b] = v.x; bj_[b + 1] = v.y;
                }
            } else bj_[0] = sB[stage][rc][tx];
#pragma unroll
            for (int a = 0; a < TI; a++)
#pragma unroll
                for (int b = 0; b < TJ; b++) acc[a][b] = __dadd_rn(acc[a][b], IMAX ? dmax_nonpos(ai[a], bj_[b]) : dmax(ai[a], bj_[b]));
        }
        __syncthreads();
    }

#pragma unroll
    for (int a = 0; a < TI; a++) {
        const uint32_t i = i0 + ty * TI + a;
#pragma unroll
        for (int b = 0; b < TJ; b++) {
            const uint32_t j = j0 + col_of(tx, b);
            if (i <= j && j < H) {
                const uint64_t g = pair_gid(i, j, H);
                if (g >= g_begin && g < g_end) {
                    const double prior = priors ? priors[g] : 0.0;
                    scores[g] = __dadd_rn(prior, acc[a][b]);
                }
            }
        }
    }
}

__device__ inline uint64_t dev_choose(uint64_t n, uint64_t k) {
    if (k > n) return 0;
    uint64_t r = k < n - k ? k : n - k, acc = 1;
    for (uint64_t v = 1; v <= r; v++) acc = acc * (n - v + 1) / v;
    return acc;
}

// g-th combination with replacement, lexicographic (src/ext/vec.rs:298-339)
__device__ inline void dev_unrank(uint64_t g, uint32_t H, uint32_t p, uint32_t *out) {
    uint32_t lo = 0;
    for (uint32_t d = 0; d < p; d++) {
        uint32_t rem = p - d - 1;
        for (uint32_t v = lo; v < H; v++) {
            uint64_t cnt = rem == 0 ? 1 : dev_choose((uint64_t)(H - v) + rem - 1, rem);
            if (g < cnt) { out[d] = v; lo = v; break; }
            g -= cnt;
        }
    }
}

__global__ void __launch_bounds__(128)
k_prefilter_gather(const double *__restrict__ Mt, uint32_t R, uint32_t H, uint32_t Hpad, uint32_t p,
                   const uint32_t *__restrict__ gt_tuples, const double *__restrict__ priors,
                   double *__restrict__ scores, uint64_t g_begin, uint64_t g_end) {
    uint64_t g = g_begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g_end) return;
    uint32_t ids[LCTP_MAX_PLOIDY];
    if (gt_tuples) { for (uint32_t k = 0; k < p; k++) ids[k] = gt_tuples[g * p + k]; }
    else dev_unrank(g, H, p, ids);
    double acc = 0.0;
    for (uint32_t r = 0; r < R; r++) {
        const double *row = Mt + (size_t)r * Hpad;
        double m = row[ids[0]];
        for (uint32_t k = 1; k < p; k++) m = dmax(m, row[ids[k]]);
        acc = __dadd_rn(acc, m);
    }
    scores[g] = __dadd_rn(priors ? priors[g] : 0.0, acc);
}

// FP64-pipe issue rate of this device, measured live: every thread runs 8 independent DADD chains, all
// SMs at full occupancy.  The prefilter needs 2 FP64-pipe instructions (DSETP + DADD) per genotype-read,
// so its roofline is rate / 2 genotype-reads per second (bench.py `roofline_prefilter`).
__global__ void __launch_bounds__(256) k_fp64_rate(double *__restrict__ out, double x, int iters) {
    double a0 = threadIdx.x, a1 = 1.0, a2 = 2.0, a3 = 3.0, a4 = 4.0, a5 = 5.0, a6 = 6.0, a7 = 7.0;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        a0 = __dadd_rn(a0, x); a1 = __dadd_rn(a1, x); a2 = __dadd_rn(a2, x); a3 = __dadd_rn(a3, x);
        a4 = __dadd_rn(a4, x); a5 = __dadd_rn(a5, x); a6 = __dadd_rn(a6, x); a7 = __dadd_rn(a7, x);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) out[0] = r;     // never true: keeps the chains alive
}

int measure_fp64_rate(lctp_ctx *ctx, double *lane_inst_per_s) {
    cudaStream_t s = ctx->stream;
    DevBuf<double> d_out;
    int rc;
    if ((rc = d_out.alloc(1))) return rc;
    const int iters = 4096, grid = ctx->sm_count * 8;
    cudaEvent_t a, b;
    LCTP_CUDA_CHECK(cudaEventCreate(&a));
    LCTP_CUDA_CHECK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        LCTP_CUDA_CHECK(cudaEventRecord(a, s));
        k_fp64_rate<<<grid, 256, 0, s>>>(d_out.p, 1e-9, iters);
        LCTP_CUDA_CHECK(cudaEventRecord(b, s));
        LCTP_CUDA_CHECK(cudaEventSynchronize(b));
        float ms = 0.f;
        LCTP_CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
        ctx->launches++;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    LCTP_CUDA_CHECK(cudaGetLastError());
    *lane_inst_per_s = (double)grid * 256.0 * 8.0 * iters / (best * 1e-3);
    return LCTP_OK;
}

int launch_prefilter(lctp_locus_h *h, uint64_t g_begin, uint64_t g_end, double *d_scores) {
    lctp_ctx *ctx = h->ctx;
    const LocusDev &d = h->dev;
    if (g_begin >= g_end) return LCTP_OK;
    if (d.p == 2 && d.gt_tuples == nullptr) {
        // Tile shape by panel size: small panels need many CTAs (the R loop is a sequential chain per
        // genotype), large panels need the 4x4 register tile that keeps the FP64 pipe busy.
        int variant = d.H < 400 ? 0 : d.H < 768 ? 4 : 1;      // measured on B200: profiles/r01_b_prefilter.md
        if (const char *e = getenv("LCTP_PREFILTER_VARIANT")) variant = atoi(e);   // tuning knob
        if (variant >= 7 && variant <= 10 && !h->mt_nonpositive) variant = 1;
        auto go = [&](auto kern, int TB, int NT) {
            uint32_t nb = (d.H + TB - 1) / TB;
            kern<<<nb * (nb + 1) / 2, NT, 0, ctx->stream>>>(d.Mt, d.R, d.H, d.Hpad, d.priors, d_scores, nb, g_begin, g_end);
        };
        switch (variant) {
        case 0: go(k_prefilter_pairs<1, 1, 16, 16, 32, false>, 16, 256); break;
        case 1: go(k_prefilter_pairs<4, 4, 8, 8, 32, false>, 32, 64); break;
        case 2: go(k_prefilter_pairs<2, 2, 16, 16, 32, false>, 32, 256); break;
        case 3: go(k_prefilter_pairs<4, 4, 16, 16, 16, false>, 64, 256); break;
        case 4: go(k_prefilter_pairs<2, 2, 8, 8, 32, false>, 16, 64); break;
        case 5: go(k_prefilter_pairs<4, 4, 8, 8, 16, false>, 32, 64); break;
        case 6: go(k_prefilter_pairs<4, 2, 8, 16, 32, false>, 32, 128); break;
        case 7: go(k_prefilter_pairs<4, 4, 8, 8, 32, true>, 32, 64); break;
        case 8: go(k_prefilter_pairs<2, 2, 16, 16, 32, true>, 32, 256); break;
        case 9: go(k_prefilter_pairs<1, 1, 16, 16, 32, true>, 16, 256); break;
        case 10: go(k_prefilter_pairs<4, 2, 8, 16, 32, true>, 32, 128); break;
        case 11: go(k_prefilter_pairs<4, 4, 8, 8, 32, false, true>, 32, 64); break;
        case 12: go(k_prefilter_pairs<2, 2, 16, 16, 32, false, true>, 32, 256); break;
        case 13: go(k_prefilter_pairs<4, 2, 8, 16, 32, false, true>, 32, 128); break;
        case 14: go(k_prefilter_pairs<1, 1, 16, 16, 32, false, true>, 16, 256); break;
        case 15: go(k_prefilter_pairs<4, 4, 16, 16, 16, false, true>, 64, 256); break;
        default: set_error("lctp_prefilter: unknown LCTP_PREFILTER_VARIANT %d", variant); return LCTP_E_INVALID;
        }
    } else {
        uint64_t n = g_end - g_begin;
        k_prefilter_gather<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
            d.Mt, d.R, d.H, d.Hpad, d.p, d.gt_tuples, d.priors, d_scores, g_begin, g_end);
    }
    ctx->launches++;
    LCTP_CUDA_CHECK(cudaGetLastError());
    return LCTP_OK;
}

}  // namespace lctp

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

namespace lctp {

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }

// genotype id of the pair (i <= j) in gen_combinations_with_repl order (src/ext/vec.rs:298-339)
__device__ __host__ __forceinline__ uint64_t pair_gid(uint64_t i, uint64_t j, uint64_t H) {
    return i * H - (i * (i - 1)) / 2 + (j - i);
}

template <int TI, int TJ, int RC>
__global__ void __launch_bounds__(256)
k_prefilter_pairs(const double *__restrict__ Mt, uint32_t R, uint32_t H, uint32_t Hpad,
                  const double *__restrict__ priors, double *__restrict__ scores, uint32_t nb,
                  uint64_t g_begin, uint64_t g_end) {
    constexpr int TBI = 16 * TI, TBJ = 16 * TJ;
    __shared__ __align__(16) double sA[2][RC][TBI];
    __shared__ __align__(16) double sB[2][RC][TBJ];

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
    const int tx = tid & 15, ty = tid >> 4;

    double acc[TI][TJ];
#pragma unroll
    for (int a = 0; a < TI; a++)
#pragma unroll
        for (int b = 0; b < TJ; b++) acc[a][b] = 0.0;

    const int n_chunks = (R + RC - 1) / RC;
    auto issue = [&](int c, int stage) {
        const uint32_t r0 = c * RC;
        constexpr int CH_A = RC * TBI / 2, CH_B = RC * TBJ / 2;   // 16-byte chunks
        for (int q = tid; q < CH_A; q += 256) {
            int rc = q / (TBI / 2), col = (q % (TBI / 2)) * 2;
            if (r0 + rc < R) cp_async16(&sA[stage][rc][col], Mt + (size_t)(r0 + rc) * Hpad + i0 + col);
        }
        for (int q = tid; q < CH_B; q += 256) {
            int rc = q / (TBJ / 2), col = (q % (TBJ / 2)) * 2;
            if (r0 + rc < R) cp_async16(&sB[stage][rc][col], Mt + (size_t)(r0 + rc) * Hpad + j0 + col);
        }
        cp_async_commit();
    };

    issue(0, 0);
    for (int c = 0; c < n_chunks; c++) {
        const int stage = c & 1;
        if (c + 1 < n_chunks) { issue(c + 1, stage ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const int nr = min((int)RC, (int)(R - c * RC));
#pragma unroll 4
        for (int rc = 0; rc < nr; rc++) {
            double ai[TI], bj_[TJ];
#pragma unroll
            for (int a = 0; a < TI; a++) ai[a] = sA[stage][rc][ty * TI + a];
#pragma unroll
            for (int b = 0; b < TJ; b++) bj_[b] = sB[stage][rc][tx * TJ + b];
#pragma unroll
            for (int a = 0; a < TI; a++)
#pragma unroll
                for (int b = 0; b < TJ; b++) acc[a][b] = __dadd_rn(acc[a][b], dmax(ai[a], bj_[b]));
        }
        __syncthreads();
    }

#pragma unroll
    for (int a = 0; a < TI; a++) {
        const uint32_t i = i0 + ty * TI + a;
#pragma unroll
        for (int b = 0; b < TJ; b++) {
            const uint32_t j = j0 + tx * TJ + b;
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

int launch_prefilter(lctp_locus_h *h, uint64_t g_begin, uint64_t g_end, double *d_scores) {
    lctp_ctx *ctx = h->ctx;
    const LocusDev &d = h->dev;
    if (g_begin >= g_end) return LCTP_OK;
    if (d.p == 2 && d.gt_tuples == nullptr) {
        if (d.H < 512) {
            constexpr int TB = 16;
            uint32_t nb = (d.H + TB - 1) / TB;
            k_prefilter_pairs<1, 1, 32><<<nb * (nb + 1) / 2, 256, 0, ctx->stream>>>(
                d.Mt, d.R, d.H, d.Hpad, d.priors, d_scores, nb, g_begin, g_end);
        } else if (d.H < 2048) {
            constexpr int TB = 32;
            uint32_t nb = (d.H + TB - 1) / TB;
            k_prefilter_pairs<2, 2, 32><<<nb * (nb + 1) / 2, 256, 0, ctx->stream>>>(
                d.Mt, d.R, d.H, d.Hpad, d.priors, d_scores, nb, g_begin, g_end);
        } else {
            constexpr int TB = 64;
            uint32_t nb = (d.H + TB - 1) / TB;
            k_prefilter_pairs<4, 4, 16><<<nb * (nb + 1) / 2, 256, 0, ctx->stream>>>(
                d.Mt, d.R, d.H, d.Hpad, d.priors, d_scores, nb, g_begin, g_end);
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

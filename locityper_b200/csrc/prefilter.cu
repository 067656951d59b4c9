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
#include <cstring>
#include <algorithm>
#include <vector>
#include <mutex>
#include <map>
#include <utility>

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
// acc += max(a, b).  LCTP_PRED_ADD=1 (experiment, measured and rejected): a compare and two complementary predicated
// adds, DSETP + @p DADD + @!p DADD = three issue slots per element instead of DSETP + 2 FSEL + DADD = four (sm_100a has
// no DMNMX).  Bit-identical, but the nullified DADD still occupies the FP64 pipe for its two cycles: KIR-scale prefilter
// 1.89 ms against 1.02 ms (profiles/r02_summary.md).
#ifndef LCTP_PRED_ADD
#define LCTP_PRED_ADD 0
#endif
__device__ __forceinline__ void acc_max(double &acc, double a, double b) {
#if LCTP_PRED_ADD
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\t@p add.rn.f64 %0, %0, %1;\n\t@!p add.rn.f64 %0, %0, %2;\n\t}"
        : "+d"(acc) : "d"(a), "d"(b));
#else
    acc = __dadd_rn(acc, dmax(a, b));
#endif
}
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

// ---- balanced persistent variant (large panels) ----------------------------------------------------
//
// The tiled kernel above leaves the FP64 pipe idle for structural reasons at the KIR-scale shape: the R loop
// of a genotype cannot be split (sequential f64 sum), so the only parallelism is the genotype count, and with
// 4x4 register tiles H = 1,000 gives 977 warp-tiles for 592 SM sub-partitions -- some sub-partitions run two
// warps, some one, and the kernel lasts as long as the fullest.  This variant plans the work on the host so
// that EVERY sub-partition gets the same load:
//   * the triangle is cut into strips of 32 haplotype rows; strip k needs the columns [32k, H);
//   * the strips' column ranges are concatenated and dealt out, in order, to warp slots; slot w of a CTA owns
//     4*c_w consecutive columns (c_w = columns per lane, 2..8), so a lane holds a 4 x c_w register tile;
//   * the per-warp widths repeat with period 4 (the warp -> sub-partition mapping), and the pattern is chosen
//     so that #CTAs <= #SMs (one persistent CTA per SM) with the smallest per-sub-partition sum of c_w:
//     H = 1,000 -> 8 warps with c = (4,4,4,4,3,3,3,3): 7 columns x 4 rows per sub-partition lane against an
//     ideal of 6.6, instead of 8 for the fullest SMs of the tiled kernel.
// A CTA stages, per read, the 32 row values of every strip it touches plus its column range (one contiguous
// smem row per read, 3-stage cp.async pipeline, one barrier per chunk).  Sums stay strictly in read order.
#ifndef LCTP_BAL_RC
#define LCTP_BAL_RC 32
#endif
#ifndef LCTP_BAL_ORDER
#define LCTP_BAL_ORDER 0
#endif
#ifndef LCTP_BAL_SLEEP
#define LCTP_BAL_SLEEP 100
#endif
#ifndef LCTP_BAL_UNROLL
#define LCTP_BAL_UNROLL 2
#endif
static constexpr int BAL_UNROLL = LCTP_BAL_UNROLL;
static constexpr int BAL_RC = LCTP_BAL_RC;     // reads per pipeline stage (a multiple of 4: one row in four per producer warp)
static constexpr int BAL_TAB = 256;            // max 16-byte chunks per staged read row (one staging pass of <= 256 threads)


// column (within the warp's 4*C-column group) of register column b of lane tx: pairs are interleaved across
// the four tx lanes (16-byte, bank-conflict-free loads); an odd last column follows the pairs.
template <int C>
__device__ __host__ __forceinline__ int bal_col(int tx, int b) {
    constexpr int NP = C / 2;
    return b < 2 * NP ? (b / 2) * 8 + 2 * tx + (b & 1) : NP * 8 + tx;
}

// One warp's share of a region: a 4 x C register tile per lane over all R reads.  The operands of read r+1 are
// loaded while read r is being accumulated (the row after a stage's last read is still inside the allocation).
template <int C>
__device__ __forceinline__ void bal_load(const double *__restrict__ pa, const double *__restrict__ pb, int odd,
                                         double (&a)[4], double (&b)[C]) {
    const double2 v0 = *reinterpret_cast<const double2 *>(pa);
    const double2 v1 = *reinterpret_cast<const double2 *>(pa + 2);
    a[0] = v0.x; a[1] = v0.y; a[2] = v1.x; a[3] = v1.y;
#pragma unroll
    for (int q = 0; q < C / 2; q++) {
        const double2 v = *reinterpret_cast<const double2 *>(pb + q * 8);
        b[2 * q] = v.x; b[2 * q + 1] = v.y;
    }
    if constexpr (C & 1) b[C - 1] = pb[odd];
}

template <int C>
__device__ __forceinline__ void bal_store(const BalWarp &w, int ty, int tx, uint32_t H, const double (&acc)[4][C],
                                          const double *__restrict__ priors, double *__restrict__ scores,
                                          uint64_t g_begin, uint64_t g_end) {
#pragma unroll
    for (int x = 0; x < 4; x++) {
        const uint32_t i = w.row0 + 4 * ty + x;
#pragma unroll
        for (int y = 0; y < C; y++) {
            const uint32_t cc = bal_col<C>(tx, y);
            const uint32_t j = w.col0 + cc;
            if (cc < w.ncols && i <= j && j < H) {
                const uint64_t g = pair_gid(i, j, H);
                if (g >= g_begin && g < g_end) scores[g] = __dadd_rn(priors ? priors[g] : 0.0, acc[x][y]);
            }
        }
    }
}

struct BalStage {                     // per-thread constants of the staging loop of one region
    const double *src;                // Mt + first row of this thread * Hpad + its source column
    uint32_t dst;                     // smem offset (doubles) of its 16-byte chunk in its first row
    uint32_t rc0, rc_step;            // first row and row stride of this thread inside a chunk
    bool active;
};

template <int NS>
__device__ __forceinline__ void bal_issue(const BalStage &sg, const double *__restrict__ Mt, double *smem, int c,
                                          int n_chunks, int stage, uint32_t R, uint32_t Hpad, uint32_t row_len,
                                          size_t stage_len) {
    if (c < n_chunks && sg.active) {
        const uint32_t r0 = c * BAL_RC;
        const uint32_t nr = min((uint32_t)BAL_RC, R - r0);
        const double *src = sg.src + (size_t)r0 * Hpad;
        double *dst = smem + (size_t)stage * stage_len + sg.dst;
        const size_t src_step = (size_t)sg.rc_step * Hpad, dst_step = (size_t)sg.rc_step * row_len;
        for (uint32_t rc = sg.rc0; rc < nr; rc += sg.rc_step) {
            cp_async16(dst, src);
            src += src_step; dst += dst_step;
        }
    }
    cp_async_commit();
}

template <int C, int NS>
__device__ __forceinline__ void bal_run(const BalWarp &w, const BalStage &sg, const double *__restrict__ Mt,
                                        double *smem, uint32_t R, uint32_t H, uint32_t Hpad, uint32_t row_len,
                                        const double *__restrict__ priors, double *__restrict__ scores,
                                        uint64_t g_begin, uint64_t g_end) {
    const int lane = threadIdx.x & 31, ty = lane >> 2, tx = lane & 3;
    const int n_chunks = (R + BAL_RC - 1) / BAL_RC;
    const size_t stage_len = (size_t)BAL_RC * row_len;
    const int odd = (C / 2) * 8 - tx;                  // odd last column, relative to pb = b_off + 2 * tx
    double acc[4][C];
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
        for (int y = 0; y < C; y++) acc[x][y] = 0.0;

#pragma unroll
    for (int c = 0; c < NS - 1; c++) bal_issue<NS>(sg, Mt, smem, c, n_chunks, c, R, Hpad, row_len, stage_len);
    int stage = 0, fill = NS - 1;                      // stage of chunk c / of chunk c + NS - 1
    for (int c = 0; c < n_chunks; c++) {
        cp_async_wait<NS - 2>();
        // Every warp of the CTA runs this loop (possibly another instantiation of it): one barrier per chunk.
        // Chunk c has landed for everyone, and the stage of chunk c - 1 is free to be refilled.
        __syncthreads();
        bal_issue<NS>(sg, Mt, smem, c + NS - 1, n_chunks, fill, R, Hpad, row_len, stage_len);
        const double *pa = smem + (size_t)stage * stage_len + w.a_off + 4 * ty;
        const double *pb = smem + (size_t)stage * stage_len + w.b_off + 2 * tx;
        const int nr = min((int)BAL_RC, (int)(R - c * BAL_RC));
        double a[4], b[C];
        bal_load<C>(pa, pb, odd, a, b);
#pragma unroll 2
        for (int rc = 0; rc < nr; rc++) {
            pa += row_len; pb += row_len;
            double an[4], bn[C];
            bal_load<C>(pa, pb, odd, an, bn);
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < C; y++) acc[x][y] = __dadd_rn(acc[x][y], dmax(a[x], b[y]));
#pragma unroll
            for (int x = 0; x < 4; x++) a[x] = an[x];
#pragma unroll
            for (int y = 0; y < C; y++) b[y] = bn[y];
        }
        stage = stage + 1 == NS ? 0 : stage + 1;
        fill = fill + 1 == NS ? 0 : fill + 1;
    }
    cp_async_wait<0>();
    bal_store<C>(w, ty, tx, H, acc, priors, scores, g_begin, g_end);
}

// Idle warp slot of a partly filled region: stages and synchronises like the others, computes nothing.
template <int NS>
__device__ __forceinline__ void bal_idle(const BalStage &sg, const double *__restrict__ Mt, double *smem, uint32_t R,
                                         uint32_t Hpad, uint32_t row_len) {
    const int n_chunks = (R + BAL_RC - 1) / BAL_RC;
    const size_t stage_len = (size_t)BAL_RC * row_len;
#pragma unroll
    for (int c = 0; c < NS - 1; c++) bal_issue<NS>(sg, Mt, smem, c, n_chunks, c, R, Hpad, row_len, stage_len);
    int fill = NS - 1;
    for (int c = 0; c < n_chunks; c++) {
        cp_async_wait<NS - 2>();
        __syncthreads();
        bal_issue<NS>(sg, Mt, smem, c + NS - 1, n_chunks, fill, R, Hpad, row_len, stage_len);
        fill = fill + 1 == NS ? 0 : fill + 1;
    }
    cp_async_wait<0>();
}

template <int NT, int CMAX, int NS>
__global__ void __launch_bounds__(NT, 1)
k_prefilter_bal(const double *__restrict__ Mt, uint32_t R, uint32_t H, uint32_t Hpad,
                const double *__restrict__ priors, double *__restrict__ scores,
                const BalRegion *__restrict__ regions, const uint32_t *__restrict__ src_tab_g, uint32_t n_regions,
                uint32_t row_len, uint64_t g_begin, uint64_t g_end) {
    extern __shared__ __align__(128) double bal_smem[];          // [NS][BAL_RC][row_len] + one padding row
    const int tid = threadIdx.x, wid = tid >> 5;

    for (uint32_t reg = blockIdx.x; reg < n_regions; reg += gridDim.x) {
        const BalRegion *rg = regions + reg;
        const uint32_t row_chunks = rg->row_chunks;
        const BalWarp w = rg->warp[wid];
        // staging: QW = smallest power of two >= row_chunks threads cover one read row, NT / QW rows per pass
        uint32_t qw = 32;
        while (qw < row_chunks) qw <<= 1;
        BalStage sg;
        {
            const uint32_t q = tid & (qw - 1), rc0 = tid / qw;
            sg.active = q < row_chunks && rc0 < (uint32_t)NT / qw;     // launcher guarantees qw <= NT
            sg.rc0 = rc0; sg.rc_step = NT / qw;
            sg.dst = rc0 * row_len + 2 * q;
            sg.src = Mt + (size_t)rc0 * Hpad + (sg.active ? src_tab_g[rg->tab_off + q] : 0u);
        }
        switch (w.c) {
        case 2: bal_run<2, NS>(w, sg, Mt, bal_smem, R, H, Hpad, row_len, priors, scores, g_begin, g_end); break;
        case 3: bal_run<3, NS>(w, sg, Mt, bal_smem, R, H, Hpad, row_len, priors, scores, g_begin, g_end); break;
        case 4: bal_run<4, NS>(w, sg, Mt, bal_smem, R, H, Hpad, row_len, priors, scores, g_begin, g_end); break;
        case 5: if constexpr (CMAX >= 5) { bal_run<5, NS>(w, sg, Mt, bal_smem, R, H, Hpad, row_len, priors, scores, g_begin, g_end); break; }
        case 6: if constexpr (CMAX >= 6) { bal_run<6, NS>(w, sg, Mt, bal_smem, R, H, Hpad, row_len, priors, scores, g_begin, g_end); break; }
        case 7: if constexpr (CMAX >= 7) { bal_run<7, NS>(w, sg, Mt, bal_smem, R, H, Hpad, row_len, priors, scores, g_begin, g_end); break; }
        case 8: if constexpr (CMAX >= 8) { bal_run<8, NS>(w, sg, Mt, bal_smem, R, H, Hpad, row_len, priors, scores, g_begin, g_end); break; }
        default: bal_idle<NS>(sg, Mt, bal_smem, R, Hpad, row_len); break;
        }
        __syncthreads();                           // the stages are reused by the next region
    }
}

// ---- warp-specialised form of the balanced kernel (variant 17) ------------------------------------------
// Same plan, same per-warp register tiles; the staging moves to a dedicated producer warp that issues one bulk
// copy (cp.async.bulk, TMA engine) per read row and column segment and signals a "full" mbarrier per stage;
// consumer warps wait on it, accumulate, and arrive on the stage's "empty" mbarrier.  No CTA-wide barrier and no
// staging instructions in the compute warps, and a warp may run up to NS - 1 chunks ahead of a slower one.
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// A consumer that has run ahead of the slowest warp of its CTA waits here; it sleeps between polls so that its
// polling does not take issue slots from the warp it shares the sub-partition with (the one everybody waits for).
__device__ __forceinline__ void mbar_spin(uint64_t *bar, unsigned parity) {
    if (mbar_try_wait(bar, parity)) return;
    while (!mbar_try_wait(bar, parity)) __nanosleep(LCTP_BAL_SLEEP);
}
// Producer side: the ring is usually full, so a producer waits most of the time; sleeping between polls keeps its
// polling out of the issue slots the compute warps of the same sub-partition need.
__device__ __forceinline__ void mbar_spin_sleep(uint64_t *bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(200);
}

struct BalPipe { int stage; unsigned phase; };     // position in the ring of NS stages, continues across regions
__device__ __forceinline__ void bal_pipe_next(BalPipe &p, int ns) {
    if (++p.stage == ns) { p.stage = 0; p.phase ^= 1u; }
}

template <int C>
__device__ __forceinline__ void bal_consume(const BalWarp &w, BalPipe &pipe, int ns, double *smem, uint64_t *full_bar,
                                            uint64_t *empty_bar, uint32_t R, uint32_t H, uint32_t row_len,
                                            const double *__restrict__ priors, double *__restrict__ scores,
                                            uint64_t g_begin, uint64_t g_end) {
    const int lane = threadIdx.x & 31, ty = lane >> 2, tx = lane & 3;
    const int n_chunks = (R + BAL_RC - 1) / BAL_RC;
    const size_t stage_len = (size_t)BAL_RC * row_len;
    const int odd = (C / 2) * 8 - tx;
    double acc[4][C];
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
        for (int y = 0; y < C; y++) acc[x][y] = 0.0;
    for (int c = 0; c < n_chunks; c++) {
        mbar_spin(&full_bar[pipe.stage], pipe.phase);
        const double *pa = smem + (size_t)pipe.stage * stage_len + w.a_off + 4 * ty;
        const double *pb = smem + (size_t)pipe.stage * stage_len + w.b_off + 2 * tx;
        const int nr = min((int)BAL_RC, (int)(R - c * BAL_RC));
        double a[4], b[C];
        bal_load<C>(pa, pb, odd, a, b);
#pragma unroll BAL_UNROLL
        for (int rc = 0; rc < nr; rc++) {
            pa += row_len; pb += row_len;
            double an[4], bn[C];
            bal_load<C>(pa, pb, odd, an, bn);      // row nr of the last chunk row is never used (next stage / padding)
#if LCTP_BAL_ORDER == 0
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < C; y++) acc_max(acc[x][y], a[x], b[y]);
#else
#pragma unroll
            for (int y = 0; y < C; y++)
#pragma unroll
                for (int x = 0; x < 4; x++) acc_max(acc[x][y], a[x], b[y]);
#endif
#pragma unroll
            for (int x = 0; x < 4; x++) a[x] = an[x];
#pragma unroll
            for (int y = 0; y < C; y++) b[y] = bn[y];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[pipe.stage]);
        bal_pipe_next(pipe, ns);
    }
    bal_store<C>(w, ty, tx, H, acc, priors, scores, g_begin, g_end);
}

// BULK: one producer warp issuing cp.async.bulk per row and segment (completion by transaction bytes); otherwise
// four producer warps (one per SM sub-partition) issuing 16-byte cp.async, each lane signalling the stage's
// "full" barrier with cp.async.mbarrier.arrive when its copies have landed.
template <int NW, int CMAX, bool BULK>
__global__ void __launch_bounds__((NW + (BULK ? 1 : 4)) * 32, 1)
k_prefilter_bal_ws(const double *__restrict__ Mt, uint32_t R, uint32_t H, uint32_t Hpad,
                   const double *__restrict__ priors, double *__restrict__ scores,
                   const BalRegion *__restrict__ regions, const uint32_t *__restrict__ src_tab_g, uint32_t n_regions,
                   uint32_t row_len, int ns, uint64_t g_begin, uint64_t g_end) {
    extern __shared__ __align__(128) double bal_smem[];          // [ns][BAL_RC][row_len] + one padding row
    __shared__ __align__(8) uint64_t full_bar[8], empty_bar[8];
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < ns; s++) { mbar_init(&full_bar[s], BULK ? 1 : 4 * 32); mbar_init(&empty_bar[s], NW); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    const int n_chunks = (R + BAL_RC - 1) / BAL_RC;
    const size_t stage_len = (size_t)BAL_RC * row_len;
    BalPipe pipe{0, 0u};

    if (!BULK && wid >= NW) {
        // producer warp p of 4: read rows p, p + 4, ... of every chunk; lane l copies the 16-byte chunks l, l + 32, ...
        const int p = wid - NW;
        pipe.phase = 1u;
        for (uint32_t reg = blockIdx.x; reg < n_regions; reg += gridDim.x) {
            const BalRegion *rg = regions + reg;
            const uint32_t row_chunks = rg->row_chunks;
            uint32_t srcc[8];
#pragma unroll
            for (int j = 0; j < 8; j++) srcc[j] = (uint32_t)(lane + 32 * j) < row_chunks ? src_tab_g[rg->tab_off + lane + 32 * j] : 0u;
            for (int c = 0; c < n_chunks; c++) {
                const uint32_t r0 = c * BAL_RC;
                const uint32_t nr = min((uint32_t)BAL_RC, R - r0);
                if (lane == 0) mbar_spin_sleep(&empty_bar[pipe.stage], pipe.phase);
                __syncwarp();
                double *dst = bal_smem + (size_t)pipe.stage * stage_len + (size_t)p * row_len + 2 * lane;
                const double *src = Mt + (size_t)(r0 + p) * Hpad;
                for (uint32_t rc = p; rc < nr; rc += 4) {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if ((uint32_t)(lane + 32 * j) < row_chunks) cp_async16(dst + 64 * j, src + srcc[j]);
                    dst += 4 * (size_t)row_len; src += 4 * (size_t)Hpad;
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&full_bar[pipe.stage])) : "memory");
                bal_pipe_next(pipe, ns);
            }
        }
        return;
    }
    if (BULK && wid == NW) {
        // producer: lane l copies read row l of every chunk
        pipe.phase = 1u;                           // a fresh "empty" barrier counts as already released
        for (uint32_t reg = blockIdx.x; reg < n_regions; reg += gridDim.x) {
            const BalRegion *rg = regions + reg;
            const uint32_t n_seg = rg->n_seg;
            uint32_t row_doubles = 0;
            for (uint32_t k = 0; k < n_seg; k++) row_doubles += rg->seg[k].len;
            for (int c = 0; c < n_chunks; c++) {
                const uint32_t r0 = c * BAL_RC;
                const uint32_t nr = min((uint32_t)BAL_RC, R - r0);
                if (lane == 0) {
                    mbar_spin_sleep(&empty_bar[pipe.stage], pipe.phase);
                    mbar_expect_tx(&full_bar[pipe.stage], nr * row_doubles * 8u);
                }
                __syncwarp();
                for (uint32_t rc = lane; rc < nr; rc += 32) {
                    const double *src = Mt + (size_t)(r0 + rc) * Hpad;
                    double *dst = bal_smem + (size_t)pipe.stage * stage_len + (size_t)rc * row_len;
                    for (uint32_t k = 0; k < n_seg; k++) {
                        const BalSeg sgm = rg->seg[k];
                        bulk_g2s(dst + sgm.dst, src + sgm.src, sgm.len * 8u, &full_bar[pipe.stage]);
                    }
                }
                bal_pipe_next(pipe, ns);
            }
        }
        return;
    }
    for (uint32_t reg = blockIdx.x; reg < n_regions; reg += gridDim.x) {
        const BalWarp w = regions[reg].warp[wid];
        switch (w.c) {
        case 2: bal_consume<2>(w, pipe, ns, bal_smem, full_bar, empty_bar, R, H, row_len, priors, scores, g_begin, g_end); break;
        case 3: bal_consume<3>(w, pipe, ns, bal_smem, full_bar, empty_bar, R, H, row_len, priors, scores, g_begin, g_end); break;
        case 4: bal_consume<4>(w, pipe, ns, bal_smem, full_bar, empty_bar, R, H, row_len, priors, scores, g_begin, g_end); break;
        case 5: if constexpr (CMAX >= 5) { bal_consume<5>(w, pipe, ns, bal_smem, full_bar, empty_bar, R, H, row_len, priors, scores, g_begin, g_end); break; }
        case 6: if constexpr (CMAX >= 6) { bal_consume<6>(w, pipe, ns, bal_smem, full_bar, empty_bar, R, H, row_len, priors, scores, g_begin, g_end); break; }
        case 7: if constexpr (CMAX >= 7) { bal_consume<7>(w, pipe, ns, bal_smem, full_bar, empty_bar, R, H, row_len, priors, scores, g_begin, g_end); break; }
        case 8: if constexpr (CMAX >= 8) { bal_consume<8>(w, pipe, ns, bal_smem, full_bar, empty_bar, R, H, row_len, priors, scores, g_begin, g_end); break; }
        default:                                   // idle slot: keep the ring moving
            for (int c = 0; c < n_chunks; c++) {
                mbar_spin(&full_bar[pipe.stage], pipe.phase);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[pipe.stage]);
                bal_pipe_next(pipe, ns);
            }
            break;
        }
    }
}

// Host planner.  `pattern[q]` = columns per lane of the q-th warp of every sub-partition (warp w uses
// pattern[w / 4]), so the CTA has 4 * n_pattern warps.  Returns the regions in order; region r is processed by
// CTA r % gridDim.

static void row_of_gid(uint64_t g, uint32_t H, uint32_t *row) {          // largest i with pair_gid(i, i, H) <= g
    uint32_t lo = 0, hi = H - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) / 2;
        if (pair_gid(mid, mid, H) <= g) lo = mid; else hi = mid - 1;
    }
    *row = lo;
}

static bool bal_plan(uint32_t H, uint64_t g_begin, uint64_t g_end, const uint32_t *pattern, uint32_t n_pattern,
                     BalPlan &out) {
    const uint32_t NW = 4 * n_pattern;
    if (n_pattern == 0 || NW > (uint32_t)BAL_MAXW || H == 0 || g_begin >= g_end) return false;
    out.regions.clear(); out.tab.clear();
    out.n_warps = NW; out.cmax = 0; out.load = 0; out.row_len = 0;
    out.pattern.assign(pattern, pattern + n_pattern);
    for (uint32_t q = 0; q < n_pattern; q++) {
        if (pattern[q] < 2 || pattern[q] > 8) return false;
        out.cmax = std::max(out.cmax, pattern[q]);
        out.load += pattern[q];
    }
    uint32_t i_lo, i_hi;
    row_of_gid(g_begin, H, &i_lo);
    row_of_gid(g_end - 1, H, &i_hi);
    const uint32_t Hc = (H + 3u) & ~3u;
    uint32_t k = i_lo / 32, k_end = i_hi / 32, col = 32 * k;
    while (k <= k_end) {
        BalRegion rg;
        std::memset(&rg, 0, sizeof rg);
        rg.tab_off = (uint32_t)out.tab.size();
        uint32_t off = 0;                 // smem row offset in doubles
        uint32_t cur_strip = UINT32_MAX, a_off = 0, b_seg_off = 0, b_seg_col0 = 0;
        for (uint32_t w = 0; w < NW && k <= k_end; w++) {
            const uint32_t c = pattern[w / 4];
            const uint32_t take = std::min(4 * c, Hc - col);
            if (k != cur_strip) {         // new strip in this region: stage its 32 row values, start a column segment
                cur_strip = k;
                a_off = off;
                for (uint32_t q = 0; q < 16; q++) out.tab.push_back(32 * k + 2 * q);
                rg.seg[rg.n_seg++] = BalSeg{32 * k, off, 32};
                off += 32;
                b_seg_off = off; b_seg_col0 = col;
                rg.seg[rg.n_seg++] = BalSeg{col, off, 0};
            }
            BalWarp &bw = rg.warp[w];
            bw.row0 = 32 * k; bw.col0 = col; bw.ncols = take; bw.c = c; bw.a_off = a_off;
            bw.b_off = b_seg_off + (col - b_seg_col0);
            for (uint32_t q = 0; q < take / 2; q++) out.tab.push_back(col + 2 * q);
            rg.seg[rg.n_seg - 1].len += take;
            rg.n_active++;
            off += take;
            col += take;
            if (col >= Hc) { k++; col = 32 * k; }
        }
        rg.row_chunks = off / 2;
        out.row_len = std::max(out.row_len, off);
        out.regions.push_back(rg);
    }
    out.row_len += 4 * 8 + 8;             // lanes past a warp's last valid column still read (and discard) smem
    out.row_len = (out.row_len + 1u) & ~1u;
    for (const BalRegion &rg : out.regions)
        if (rg.row_chunks > (uint32_t)BAL_TAB) return false;
    return true;
}

// Chooses the per-sub-partition pattern with the smallest (rounds x load); ties -> fewer warps (wider tiles).
static bool bal_choose(uint32_t H, uint64_t g_begin, uint64_t g_end, uint32_t n_sm, BalPlan &best) {
    uint64_t best_cost = UINT64_MAX;
    BalPlan cur;
    for (uint32_t np = 1; np <= 3; np++) {
        uint32_t pat[3];
        const uint32_t combos = np == 1 ? 3 : np == 2 ? 9 : 27;
        for (uint32_t m = 0; m < combos; m++) {
            uint32_t x = m;
            bool sorted = true;
            for (uint32_t q = 0; q < np; q++) { pat[q] = 4 - x % 3; x /= 3; if (q && pat[q] > pat[q - 1]) sorted = false; }
            if (!sorted) continue;        // widest tiles first: they take the long strips
            if (!bal_plan(H, g_begin, g_end, pat, np, cur)) continue;
            const uint64_t rounds = (cur.regions.size() + n_sm - 1) / n_sm;
            const uint64_t cost = rounds * cur.load * 64 + np;
            if (cost < best_cost) { best_cost = cost; std::swap(best, cur); }
        }
    }
    return best_cost != UINT64_MAX;
}

static bool parse_pattern(const char *e, uint32_t *pat, uint32_t *np) {
    *np = 0;
    while (*e && *np < 4) {
        char *end = nullptr;
        const long v = strtol(e, &end, 10);
        if (end == e) return false;
        pat[(*np)++] = (uint32_t)v;
        e = *end == ',' ? end + 1 : end;
    }
    return *np > 0 && *e == 0;
}

// Raise a kernel's dynamic shared-memory limit only when it has to grow: re-setting a function attribute makes the
// next launch of that function wait for its running instances (it serialised kernels of concurrent contexts).
template <typename K>
static cudaError_t ensure_dyn_smem(K kern, int device, size_t bytes) {
    static std::mutex m;
    static std::map<std::pair<const void *, int>, size_t> lim;      // (kernel instantiation, device) -> limit set
    std::lock_guard<std::mutex> lk(m);
    size_t &l = lim[std::make_pair((const void *)kern, device)];
    if (bytes <= 48 * 1024 || bytes <= l) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) l = bytes;
    return e;
}

template <int NT, int CMAX, int NS>
static int bal_launch2(lctp_locus_h *h, const BalPlan &plan, const BalRegion *d_regions, const uint32_t *d_tab,
                       uint64_t g_begin, uint64_t g_end, double *d_scores, size_t smem) {
    lctp_ctx *ctx = h->ctx;
    const LocusDev &d = h->dev;
    LCTP_CUDA_CHECK(ensure_dyn_smem(k_prefilter_bal<NT, CMAX, NS>, ctx->device, smem));
    const unsigned grid = (unsigned)std::min<size_t>(plan.regions.size(), (size_t)ctx->sm_count);
    k_prefilter_bal<NT, CMAX, NS><<<grid, NT, smem, ctx->stream>>>(d.Mt, d.R, d.H, d.Hpad, d.priors, d_scores, d_regions, d_tab,
                                                                (uint32_t)plan.regions.size(), plan.row_len, g_begin, g_end);
    return LCTP_OK;
}

template <int NT, int CMAX>
static int bal_launch(lctp_locus_h *h, const BalPlan &plan, const BalRegion *d_regions, const uint32_t *d_tab,
                      uint64_t g_begin, uint64_t g_end, double *d_scores) {
    // three stages when they fit (plus the padding row the operand prefetch of the last read touches), else two
    auto bytes = [&](int ns) { return ((size_t)ns * BAL_RC + 1) * plan.row_len * sizeof(double); };
    int ns = 3;
    if (const char *e = getenv("LCTP_PREFILTER_BAL_STAGES")) ns = atoi(e) == 2 ? 2 : 3;
    if (bytes(ns) > h->ctx->smem_optin) ns = 2;
    if (bytes(ns) > h->ctx->smem_optin) { set_error("lctp_prefilter: balanced plan needs %zu bytes of shared memory", bytes(ns)); return LCTP_E_CAPACITY; }
    for (const BalRegion &rg : plan.regions) {
        uint32_t qw = 32;
        while (qw < rg.row_chunks) qw <<= 1;
        if (qw > (uint32_t)NT) { set_error("lctp_prefilter: balanced plan stages %u chunks per read with %d threads", rg.row_chunks, NT); return LCTP_E_CAPACITY; }
    }
    return ns == 3 ? bal_launch2<NT, CMAX, 3>(h, plan, d_regions, d_tab, g_begin, g_end, d_scores, bytes(3))
                   : bal_launch2<NT, CMAX, 2>(h, plan, d_regions, d_tab, g_begin, g_end, d_scores, bytes(2));
}

template <int NW, int CMAX, bool BULK>
static int bal_launch_ws(lctp_locus_h *h, const BalPlan &plan, const BalRegion *d_regions, uint64_t g_begin,
                         uint64_t g_end, double *d_scores) {
    lctp_ctx *ctx = h->ctx;
    const LocusDev &d = h->dev;
    auto bytes = [&](int ns) { return ((size_t)ns * BAL_RC + 1) * plan.row_len * sizeof(double); };
    int ns = 4;
    if (const char *e = getenv("LCTP_PREFILTER_BAL_STAGES")) ns = std::min(8, std::max(2, atoi(e)));
    while (ns > 2 && bytes(ns) + 1024 > ctx->smem_optin) ns--;
    if (bytes(ns) + 1024 > ctx->smem_optin) { set_error("lctp_prefilter: balanced plan needs %zu bytes of shared memory", bytes(ns)); return LCTP_E_CAPACITY; }
    for (const BalRegion &rg : plan.regions)
        if (rg.row_chunks > 256u) { set_error("lctp_prefilter: balanced plan stages %u chunks per read", rg.row_chunks); return LCTP_E_CAPACITY; }
    LCTP_CUDA_CHECK(ensure_dyn_smem(k_prefilter_bal_ws<NW, CMAX, BULK>, ctx->device, bytes(ns)));
    const unsigned grid = (unsigned)std::min<size_t>(plan.regions.size(), (size_t)ctx->sm_count);
    k_prefilter_bal_ws<NW, CMAX, BULK><<<grid, (NW + (BULK ? 1 : 4)) * 32, bytes(ns), ctx->stream>>>(
        d.Mt, d.R, d.H, d.Hpad, d.priors, d_scores, d_regions, h->pf_tab.p, (uint32_t)plan.regions.size(), plan.row_len, ns,
        g_begin, g_end);
    return LCTP_OK;
}

// pattern == nullptr: choose automatically.  ws: warp-specialised kernel (bulk-copy producer warp).
static int launch_prefilter_bal(lctp_locus_h *h, uint64_t g_begin, uint64_t g_end, double *d_scores,
                                const uint32_t *pattern, uint32_t n_pattern, int ws) {
    lctp_ctx *ctx = h->ctx;
    BalPlan &plan = h->pf_plan;
    std::vector<uint32_t> want(pattern, pattern + (pattern ? n_pattern : 0));
    const bool cached = h->pf_plan_valid && h->pf_g_begin == g_begin && h->pf_g_end == g_end &&
                        (pattern ? plan.pattern == want : h->pf_plan_auto);
    if (!cached) {                        // the plan depends only on (H, genotype range, pattern): keep it with the locus
        h->pf_plan_valid = false;
        const bool ok = pattern ? bal_plan(h->dev.H, g_begin, g_end, pattern, n_pattern, plan)
                                : bal_choose(h->dev.H, g_begin, g_end, (uint32_t)ctx->sm_count, plan);
        if (!ok) { set_error("lctp_prefilter: no balanced plan for H=%u", h->dev.H); return LCTP_E_INVALID; }
        int rc;
        if ((rc = h->pf_regions.ensure(plan.regions.size()))) return rc;
        if ((rc = h->pf_tab.ensure(plan.tab.size()))) return rc;
        // pageable sources: the copies are staged by the runtime before the calls return
        LCTP_CUDA_CHECK(cudaMemcpyAsync(h->pf_regions.p, plan.regions.data(), plan.regions.size() * sizeof(BalRegion),
                                        cudaMemcpyHostToDevice, ctx->stream));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(h->pf_tab.p, plan.tab.data(), plan.tab.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        h->pf_plan_valid = true; h->pf_plan_auto = pattern == nullptr;
        h->pf_g_begin = g_begin; h->pf_g_end = g_end;
    }
    const BalRegion *dr = h->pf_regions.p;
    const uint32_t nt = plan.n_warps * 32;
    if (ws == 1) {
        if (plan.cmax <= 4) {
            switch (plan.n_warps) {
            case 4: return bal_launch_ws<4, 4, true>(h, plan, dr, g_begin, g_end, d_scores);
            case 8: return bal_launch_ws<8, 4, true>(h, plan, dr, g_begin, g_end, d_scores);
            case 12: return bal_launch_ws<12, 4, true>(h, plan, dr, g_begin, g_end, d_scores);
            case 16: return bal_launch_ws<16, 4, true>(h, plan, dr, g_begin, g_end, d_scores);
            }
        } else {
            switch (plan.n_warps) {
            case 4: return bal_launch_ws<4, 8, true>(h, plan, dr, g_begin, g_end, d_scores);
            case 8: return bal_launch_ws<8, 8, true>(h, plan, dr, g_begin, g_end, d_scores);
            }
        }
        set_error("lctp_prefilter: unsupported balanced pattern (%u warps, %u columns per lane)", plan.n_warps, plan.cmax);
        return LCTP_E_INVALID;
    }
    if (ws == 2) {
        if (plan.cmax <= 4) {
            switch (plan.n_warps) {
            case 4: return bal_launch_ws<4, 4, false>(h, plan, dr, g_begin, g_end, d_scores);
            case 8: return bal_launch_ws<8, 4, false>(h, plan, dr, g_begin, g_end, d_scores);
            case 12: return bal_launch_ws<12, 4, false>(h, plan, dr, g_begin, g_end, d_scores);
            }
        } else {
            switch (plan.n_warps) {
            case 4: return bal_launch_ws<4, 8, false>(h, plan, dr, g_begin, g_end, d_scores);
            case 8: return bal_launch_ws<8, 8, false>(h, plan, dr, g_begin, g_end, d_scores);
            }
        }
        set_error("lctp_prefilter: unsupported balanced pattern (%u warps, %u columns per lane)", plan.n_warps, plan.cmax);
        return LCTP_E_INVALID;
    }
    if (plan.cmax <= 4) {
        switch (nt) {
        case 128: return bal_launch<128, 4>(h, plan, dr, h->pf_tab.p, g_begin, g_end, d_scores);
        case 256: return bal_launch<256, 4>(h, plan, dr, h->pf_tab.p, g_begin, g_end, d_scores);
        case 384: return bal_launch<384, 4>(h, plan, dr, h->pf_tab.p, g_begin, g_end, d_scores);
        case 512: return bal_launch<512, 4>(h, plan, dr, h->pf_tab.p, g_begin, g_end, d_scores);
        }
    } else {
        switch (nt) {
        case 128: return bal_launch<128, 8>(h, plan, dr, h->pf_tab.p, g_begin, g_end, d_scores);
        case 256: return bal_launch<256, 8>(h, plan, dr, h->pf_tab.p, g_begin, g_end, d_scores);
        }
    }
    set_error("lctp_prefilter: unsupported balanced pattern (%u warps, %u columns per lane)", plan.n_warps, plan.cmax);
    return LCTP_E_INVALID;
}

// Host-side check of the planner (no device work): every genotype of [g_begin, g_end) is owned by exactly one
// (region, warp, lane, register) slot.  Diagnostic entry point, see include/lctp.h.
int prefilter_plan_check(uint32_t H, uint32_t n_sm, const uint32_t *pattern, uint32_t n_pattern, uint64_t g_begin,
                         uint64_t g_end, uint32_t *n_regions, uint32_t *load, uint32_t *pattern_out) {
    BalPlan plan;
    const bool ok = (pattern && n_pattern) ? bal_plan(H, g_begin, g_end, pattern, n_pattern, plan)
                                           : bal_choose(H, g_begin, g_end, n_sm, plan);
    if (!ok) { set_error("lctp_prefilter_plan_check: no plan"); return LCTP_E_INVALID; }
    std::vector<uint8_t> seen(g_end - g_begin, 0);
    for (const BalRegion &rg : plan.regions) {
        for (uint32_t w = 0; w < plan.n_warps; w++) {
            const BalWarp &bw = rg.warp[w];
            if (bw.c == 0) continue;
            if (bw.a_off + 32 > plan.row_len || bw.b_off + 4 * bw.c + 8 > plan.row_len) {
                set_error("lctp_prefilter_plan_check: smem offsets out of range");
                return LCTP_E_INVALID;
            }
            for (int lane = 0; lane < 32; lane++) {
                const int ty = lane >> 2, tx = lane & 3;
                for (uint32_t x = 0; x < 4; x++)
                    for (uint32_t y = 0; y < bw.c; y++) {
                        const uint32_t np = bw.c / 2;
                        const uint32_t cc = y < 2 * np ? (y / 2) * 8 + 2 * tx + (y & 1) : np * 8 + tx;
                        const uint32_t i = bw.row0 + 4 * ty + x, j = bw.col0 + cc;
                        if (cc >= bw.ncols || i > j || j >= H) continue;
                        // the staged source column of this register must be the genotype's haplotype
                        const uint32_t src_b = plan.tab[rg.tab_off + (bw.b_off + cc) / 2] + ((bw.b_off + cc) & 1u);
                        const uint32_t src_a = plan.tab[rg.tab_off + (bw.a_off + 4 * ty + x) / 2] + ((bw.a_off + 4 * ty + x) & 1u);
                        if (src_a != i || src_b != j) { set_error("lctp_prefilter_plan_check: staging table mismatch"); return LCTP_E_INVALID; }
                        const uint64_t g = pair_gid(i, j, H);
                        if (g < g_begin || g >= g_end) continue;
                        if (seen[g - g_begin]++) { set_error("lctp_prefilter_plan_check: genotype %llu covered twice", (unsigned long long)g); return LCTP_E_INVALID; }
                    }
            }
        }
    }
    for (uint64_t g = g_begin; g < g_end; g++)
        if (!seen[g - g_begin]) { set_error("lctp_prefilter_plan_check: genotype %llu not covered", (unsigned long long)g); return LCTP_E_INVALID; }
    if (n_regions) *n_regions = (uint32_t)plan.regions.size();
    if (load) *load = plan.load;
    if (pattern_out) for (size_t q = 0; q < plan.pattern.size(); q++) pattern_out[q] = plan.pattern[q];
    return LCTP_OK;
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
        // Small panels: the tiled kernel with 1x1 register tiles (many CTAs).  From H = 400 up: the balanced
        // persistent kernel with producer warps (variant 18) -- measured on B200 (profiles/r01_c_summary.md):
        // H = 500: 0.165 ms against 0.219 (2x2 tiles), H = 1,000: 1.02 ms against 1.24 (4x4 tiles), and the
        // genotype range of one rank of 2 / 8 at H = 1,000: 0.63 / 0.39 ms against 1.23 / 0.78.
        int variant = d.H < 400 ? 0 : 18;
        const char *ev = getenv("LCTP_PREFILTER_VARIANT");         // tuning knob
        if (ev) variant = atoi(ev);
        if (variant >= 7 && variant <= 10 && !h->mt_nonpositive) variant = 1;
        if (variant >= 16 && variant <= 18) {   // balanced persistent kernels; LCTP_PREFILTER_BAL="4,3" fixes the pattern
            uint32_t pat[4], np = 0;
            const char *e = getenv("LCTP_PREFILTER_BAL");
            if (e && *e && !parse_pattern(e, pat, &np)) { set_error("lctp_prefilter: bad LCTP_PREFILTER_BAL '%s'", e); return LCTP_E_INVALID; }
            int rc = launch_prefilter_bal(h, g_begin, g_end, d_scores, np ? pat : nullptr, np, variant - 16);
            if (rc == LCTP_OK) {
                ctx->launches++;
                LCTP_CUDA_CHECK(cudaGetLastError());
                return LCTP_OK;
            }
            if (ev || rc != LCTP_E_CAPACITY) return rc;             // explicit request, or a real error
            variant = d.H < 768 ? 4 : 1;                            // plan does not fit shared memory: tiled kernel
        }
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

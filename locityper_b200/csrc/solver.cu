// solver.cu -- a5..a13: one solver stage on the device.
//
// One lane GROUP (GS = 16 lanes, two groups per warp; GS = 32 selectable) per logical worker (= one
// xoshiro256++ stream of the reference, src/solvers/solve.rs:1007-1018).  A worker solves its genotypes
// back to back on that stream exactly like Worker::run (:1104-1145):
//   per genotype: GenotypeAlignments::new (src/model/assgn.rs:41-84, windows.rs:762-797)   -> build_instance
//   per attempt : apply_tweak (assgn.rs:127-151, windows.rs:123-136,478-486)               -> apply_tweak
//                 Solver::solve (solvers/mod.rs:61-72) = Greedy (stoch.rs:81-120) | SimAnneal (:197-242)
//                 lik = prior + likelihood() (solve.rs:1126); update_counts (assgn.rs:374-378)
//   mean / variance over attempts (ext/vec.rs:74-116)
//
// Exactness: all f64 arithmetic that feeds an accept/reject uses the non-contracting intrinsics
// (__dadd_rn/__dmul_rn/__dsub_rn) in the reference's association order; the file is also compiled with
// -fmad=false.  Random draws are consumed in exactly the reference's order (see "RNG stream").
//
// Data placement per worker:
//   shared memory : window records (weight, table row, depth), compact offsets + current assignment of
//                   the non-trivial reads, a 2-stage cp.async staging area for the greedy pipeline;
//   registers     : the two likelihood accumulators, RNG window;
//   global slab   : candidate arrays in read order (build / tweak / counts) and a compact copy of the
//                   non-trivial reads' candidates (what the solver loops touch), RNG draw buffer.
#include "common.cuh"

#include <cmath>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <chrono>

namespace lctp {

#ifndef LCTP_CTA_THREADS
#define LCTP_CTA_THREADS 32
#endif
static constexpr int CTA_THREADS = LCTP_CTA_THREADS;
static constexpr int MAX_SAMPLE = 11;          // Floyd branch of rand::seq::index::sample
#ifndef LCTP_HEADS
#define LCTP_HEADS 2
#endif
#ifndef LCTP_RNG_C
#define LCTP_RNG_C 256
#endif
static constexpr int RNG_C = LCTP_RNG_C;               // stream outputs generated per lane per fill
static constexpr int RNG_BUF = 32 * RNG_C;     // slab space for one fill (GS * RNG_C <= this)
static constexpr int N_SETUP_MATS = 5;         // T^(C*2^k), k = 0..4

struct StageParams {
    uint32_t kind, attempts, best_start, sample_size;
    uint64_t plato_size, anneal_steps, max_iter;
    double ln_init_prob;
    uint32_t n_workers, cap, Wmax, want_counts;
    uint64_t slab_bytes;
};

struct Slab {
    // Candidates of every read in read order (the reference's `alns`, a5): only what apply_tweak and the
    // count output need per candidate -- where it came from and where its solver record lives.
    uint2 *cmap;           // [cap]  x = index into the cm arrays | contig index << 28 (LCTP_NONE_U32 = the
                           //        "both mates unmapped" option); y = index of its record in `ntc`, or
                           //        TRIV_TAG | read id for the single candidate of a trivial read
    uint32_t *read_off;    // [R+1]
    uint4 *ntc;            // [cap]  records of the non-trivial reads' candidates, compact, in read order:
                           //        x,y = ln_prob (f64 bits), z = windows (w1 | w2 << 16), w unused
    double *triv_lp;       // [R]    ln_prob of the only candidate of a trivial read
    uint32_t *triv_w;      // [R]    its windows
    uint64_t *rng_buf;     // [RNG_BUF] pre-generated draws of the worker's stream
    uint64_t *rng_blk;     // [32*4] block-start generator states of the current fill
};
static constexpr uint32_t TRIV_TAG = 0x80000000u;
static constexpr uint32_t SRC_MASK = 0x0FFFFFFFu;     // cm index bits of cmap.x (the contig index sits above)

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline size_t slab_layout(uint32_t cap, uint32_t R, unsigned char *base, Slab *s) {
    size_t o = 0;
    if (s) s->ntc = (uint4 *)(base + o);
    o += align_up((size_t)cap * 16, 128);
    if (s) s->cmap = (uint2 *)(base + o);
    o += align_up((size_t)cap * 8, 128);
    if (s) s->triv_lp = (double *)(base + o);
    o += align_up((size_t)R * 8, 128);
    if (s) s->triv_w = (uint32_t *)(base + o);
    o += align_up((size_t)R * 4, 128);
    if (s) s->read_off = (uint32_t *)(base + o);
    o += align_up(((size_t)R + 1) * 4, 128);
    if (s) s->rng_buf = (uint64_t *)(base + o);
    o += (size_t)RNG_BUF * 8;
    if (s) s->rng_blk = (uint64_t *)(base + o);
    o += 32 * 4 * 8;
    return o;
}

// ------------------------------------------------------------------ lane groups -----------------

// A worker is served by GS consecutive lanes of a warp.  All collectives are restricted to the group's
// lane mask, so the two groups of a warp run two independent workers in the same instruction stream.
template <int GS>
struct Grp {
    unsigned mask;
    int shift, lane;
    __device__ Grp() {
        const int wl = threadIdx.x & 31;
        shift = GS == 32 ? 0 : (wl & ~(GS - 1));
        mask = GS == 32 ? 0xFFFFFFFFu : (((1u << GS) - 1u) << shift);
        lane = wl & (GS - 1);
    }
    static constexpr unsigned LM = GS == 32 ? 0xFFFFFFFFu : ((1u << GS) - 1u);
    template <typename T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(mask, v, src, GS); }
    template <typename T> __device__ __forceinline__ T shfl_up(T v, int d) const { return __shfl_up_sync(mask, v, d, GS); }
    template <typename T> __device__ __forceinline__ T shfl_xor(T v, int m) const { return __shfl_xor_sync(mask, v, m, GS); }
    __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) >> shift) & LM; }
    __device__ __forceinline__ bool any(bool p) const { return __any_sync(mask, p) != 0; }
    __device__ __forceinline__ unsigned match_any(uint32_t v) const { return (__match_any_sync(mask, v) >> shift) & LM; }
    __device__ __forceinline__ uint32_t rmax(uint32_t v) const { return __reduce_max_sync(mask, v); }
    __device__ __forceinline__ uint32_t rmin(uint32_t v) const { return __reduce_min_sync(mask, v); }
    __device__ __forceinline__ uint32_t ror(uint32_t v) const { return __reduce_or_sync(mask, v); }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    __device__ __forceinline__ unsigned lt() const { return (1u << lane) - 1u; }
};

// ------------------------------------------------------------------ RNG stream -------------------
//
// The reference consumes ONE sequential xoshiro256++ stream per worker.  Stepping that generator
// redundantly in every lane costs ~25 instructions per draw.  Instead the group pre-generates the stream
// in bulk: xoshiro's state transition is linear over GF(2), so lane l jumps its copy of the state ahead
// by l*RNG_C steps (256x256 bit-matrix products with host-precomputed powers of the transition matrix)
// and then generates RNG_C consecutive outputs of the SAME sequential stream into a per-worker buffer:
// GS lanes produce GS*64 exact draws per fill.  Consumers read the buffer through a GS-entry register
// window with shuffles, in stream order, so the draw sequence (including the data-dependent extra draw
// of biased bounded samples) is bit-identical to the sequential generator.  At the end of a worker the
// exact state at the consumed position is rebuilt from the owning lane's block-start state.

__constant__ uint64_t c_refill_mat[2][256 * 4];   // [0]: T^(15*C) for GS=16, [1]: T^(31*C) for GS=32

struct Gen { uint64_t s0, s1, s2, s3; };

__device__ __host__ __forceinline__ uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

// xoshiro256++ (rand_xoshiro::Xoshiro256PlusPlus, src/ext/rand.rs:3)
__device__ __host__ __forceinline__ uint64_t gen_next(Gen &x) {
    const uint64_t r = rotl64(x.s0 + x.s3, 23) + x.s0;
    const uint64_t t = x.s1 << 17;
    x.s2 ^= x.s0; x.s3 ^= x.s1; x.s1 ^= x.s2; x.s0 ^= x.s3; x.s2 ^= t;
    x.s3 = rotl64(x.s3, 45);
    return r;
}

// state <- M * state over GF(2); M given by its 256 columns (4 words each); `take` selects per lane
// whether the product replaces the state (all lanes execute the same instruction stream).
__device__ __forceinline__ void gen_mat_apply(Gen &g, const uint64_t *__restrict__ mat, bool take) {
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 1
    for (int w = 0; w < 4; w++) {
        const uint64_t sw = w == 0 ? g.s0 : w == 1 ? g.s1 : w == 2 ? g.s2 : g.s3;
#pragma unroll 8
        for (int b = 0; b < 64; b++) {
            const uint64_t m = 0ull - ((sw >> b) & 1ull);
            const uint64_t *c = mat + (size_t)(w * 64 + b) * 4;
            a0 ^= c[0] & m; a1 ^= c[1] & m; a2 ^= c[2] & m; a3 ^= c[3] & m;
        }
    }
    if (take) { g.s0 = a0; g.s1 = a1; g.s2 = a2; g.s3 = a3; }
}

template <int GS>
struct Xo {                // the worker's stream as seen by the solver code
    static constexpr uint32_t FILL = GS * RNG_C;
    Grp<GS> g;
    Gen gen;               // this lane's generator (positioned at the end of its block after a fill)
    uint32_t a_lo, a_hi;   // register blocks: A = buf[base + lane], B = buf[base + GS + lane] (B is the
    uint32_t b_lo, b_hi;   // prefetched successor, so moving the window never waits for memory)
    uint32_t pos;          // draws consumed from the current fill
    uint32_t base;         // stream position of lane 0 of block A (multiple of GS)
    uint64_t *buf;         // [FILL] per-worker buffer (global, L2-resident)
    uint64_t *blk;         // [GS][4] block-start states of the current fill
};

// Generate this lane's block of the next fill; returns the generator positioned at the block end.
// Out of line and by value so that the caller's stream object never has its address taken (it would be
// demoted to local memory otherwise).
__device__ __noinline__ Gen fill_block(Gen gen, uint64_t *buf, uint64_t *blk, int lane, unsigned mask, int jump_mat) {
    if (jump_mat >= 0) gen_mat_apply(gen, c_refill_mat[jump_mat], true);   // end of own block -> own block of the next fill
    uint64_t *b = blk + lane * 4;
    b[0] = gen.s0; b[1] = gen.s1; b[2] = gen.s2; b[3] = gen.s3;
    uint64_t *o = buf + lane * RNG_C;
#pragma unroll 4
    for (int q = 0; q < RNG_C; q++) o[q] = gen_next(gen);
    __syncwarp(mask);
    return gen;
}
template <int GS>
__device__ __forceinline__ void stream_fill(Xo<GS> &x, int jump_mat) {
    x.gen = fill_block(x.gen, x.buf, x.blk, x.g.lane, x.g.mask, jump_mat);
    x.pos = 0;
    x.base = 0;
    const uint64_t a = __ldcg(x.buf + x.g.lane), b = __ldcg(x.buf + GS + x.g.lane);
    x.a_lo = (uint32_t)a; x.a_hi = (uint32_t)(a >> 32);
    x.b_lo = (uint32_t)b; x.b_hi = (uint32_t)(b >> 32);
}

// Start a stream from the scalar state st[4]: lane l jumps ahead by l*RNG_C (binary decomposition of l).
template <int GS>
__device__ void stream_begin(Xo<GS> &x, const uint64_t *__restrict__ st, const uint64_t *__restrict__ setup_mats) {
    x.gen.s0 = st[0]; x.gen.s1 = st[1]; x.gen.s2 = st[2]; x.gen.s3 = st[3];
    constexpr int ROUNDS = GS == 32 ? 5 : 4;
    for (int k = 0; k < ROUNDS; k++) gen_mat_apply(x.gen, setup_mats + (size_t)k * 1024, (x.g.lane >> k) & 1);
    stream_fill(x, -1);
}

template <int GS>
__device__ __forceinline__ void stream_refill(Xo<GS> &x) { stream_fill(x, GS == 32 ? 1 : 0); }

// Exact scalar state after the draws consumed so far (what the sequential generator would hold).
template <int GS>
__device__ void stream_end(Xo<GS> &x, uint64_t *__restrict__ out) {
    x.g.sync();
    Gen g = x.gen;                                        // pos == FILL: lane GS-1's copy is the state
    int owner = GS - 1;
    if (x.pos < Xo<GS>::FILL) {
        const int blk_owner = x.pos / RNG_C;
        const uint64_t *b = x.blk + blk_owner * 4;
        g.s0 = b[0]; g.s1 = b[1]; g.s2 = b[2]; g.s3 = b[3];
        const int steps = x.pos - blk_owner * RNG_C;
        for (int q = 0; q < steps; q++) gen_next(g);
        owner = 0;                                        // every lane computed the same state
    }
    if (x.g.lane == owner) { out[0] = g.s0; out[1] = g.s1; out[2] = g.s2; out[3] = g.s3; }
}

// Make the register blocks cover stream positions [pos, pos + count), count <= GS; false = the fill
// ends first (the caller then uses the one-draw-at-a-time path, which refills).  Position q of the
// stream lives in lane q % GS: in block A for q < base + GS, else in block B.
template <int GS>
__device__ __forceinline__ bool stream_cover(Xo<GS> &x, uint32_t count) {
    if (x.pos + count > Xo<GS>::FILL) return false;
    if (x.pos - x.base >= (uint32_t)GS) {                  // A is used up: B becomes A, prefetch the next B
        x.a_lo = x.b_lo; x.a_hi = x.b_hi;
        x.base += GS;
        const uint64_t v = __ldcg(x.buf + min(x.base + GS + (uint32_t)x.g.lane, Xo<GS>::FILL - 1u));
        x.b_lo = (uint32_t)v; x.b_hi = (uint32_t)(v >> 32);
    }
    return true;
}
// After stream_cover(count): the (pos + rank)-th draw of the stream, for any per-lane rank < count.
template <int GS>
__device__ __forceinline__ uint32_t stream_peek_hi(const Xo<GS> &x, uint32_t rank) {
    const uint32_t off = x.pos - x.base;
    return x.g.shfl((uint32_t)x.g.lane >= off ? x.a_hi : x.b_hi, (int)((off + rank) & (GS - 1)));
}
template <int GS>
__device__ __forceinline__ uint64_t stream_peek64(const Xo<GS> &x, uint32_t rank) {
    const uint32_t off = x.pos - x.base;
    const bool in_a = (uint32_t)x.g.lane >= off;
    const int src = (int)((off + rank) & (GS - 1));
    return ((uint64_t)x.g.shfl(in_a ? x.a_hi : x.b_hi, src) << 32) | x.g.shfl(in_a ? x.a_lo : x.b_lo, src);
}
// Give back the last n draws (n <= GS, all taken from the current fill without a refill in between).
template <int GS>
__device__ __forceinline__ void stream_unconsume(Xo<GS> &x, uint32_t n) {
    x.pos -= n;
    while (x.pos < x.base) {
        x.b_lo = x.a_lo; x.b_hi = x.a_hi;
        x.base -= GS;
        const uint64_t v = __ldcg(x.buf + x.base + x.g.lane);
        x.a_lo = (uint32_t)v; x.a_hi = (uint32_t)(v >> 32);
    }
}
template <int GS>
__device__ __forceinline__ void stream_advance(Xo<GS> &x) {
    if (x.pos == Xo<GS>::FILL) stream_refill(x);
    stream_cover(x, 1);
}
template <int GS>
__device__ __forceinline__ uint32_t xo_u32(Xo<GS> &x) {   // next_u32 = upper half of next_u64
    stream_advance(x);
    const uint32_t r = stream_peek_hi(x, 0);
    x.pos++;
    return r;
}
template <int GS>
__device__ __forceinline__ uint64_t xo_next(Xo<GS> &x) {
    stream_advance(x);
    const uint64_t r = stream_peek64(x, 0);
    x.pos++;
    return r;
}
// rand UniformInt::sample_single_inclusive with a u32 sample type: value in [0, range), range != 0.
template <int GS>
__device__ __forceinline__ uint32_t xo_below(Xo<GS> &x, uint32_t range) {
    const uint64_t m = (uint64_t)xo_u32(x) * (uint64_t)range;
    uint32_t res = (uint32_t)(m >> 32);
    const uint32_t lo = (uint32_t)m;
    if (lo > 0u - range) {   // biased zone: one extra draw (group-uniform branch)
        const uint32_t nh = (uint32_t)(((uint64_t)xo_u32(x) * (uint64_t)range) >> 32);
        res += (lo + nh < lo) ? 1u : 0u;
    }
    return res;
}
// rand StandardUniform f64 (src/solvers/stoch.rs:216)
template <int GS>
__device__ __forceinline__ double xo_f64(Xo<GS> &x) {
    return (double)(xo_next(x) >> 11) * (1.0 / 9007199254740992.0);
}
// Lane-parallel bounded draws: lanes < count each want random_range(0..range) and lane k must receive
// the k-th draw of the stream.  Succeeds (and consumes `count` draws) only when no lane lands in the
// biased zone -- which would consume an extra draw and shift every later lane -- otherwise nothing is
// consumed and the caller runs the sequential path.  P(fallback) ~ count * range / 2^32.
template <int GS>
__device__ __forceinline__ bool xo_below_parallel(Xo<GS> &x, uint32_t count, uint32_t range, uint32_t &res) {
    if (!stream_cover(x, count)) return false;
    const uint64_t m = (uint64_t)stream_peek_hi(x, (uint32_t)x.g.lane) * (uint64_t)range;
    const bool biased = (uint32_t)x.g.lane < count && (uint32_t)m > 0u - range;
    if (x.g.any(biased)) return false;
    res = (uint32_t)(m >> 32);
    x.pos += count;
    return true;
}

// ------------------------------------------------------------------ small helpers ---------------

// f64::total_cmp key (Rust std)
__device__ __forceinline__ long long total_key(double v) {
    long long b = __double_as_longlong(v);
    return b ^ (long long)(((unsigned long long)(b >> 63)) >> 1);
}
// Monotone u64 key of an f64 (no NaNs on this path); 0 is below every valid key.
__device__ __forceinline__ unsigned long long ord_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
}

// Per-window state in shared memory, 64 bytes per window: the window's weight, its depth-table row, the
// current read depth d and the five products weight * table[row][d-2 .. d+2] (rounded once, exactly as
// WindowDistr::ln_prob computes them, src/model/distr_cache.rs:34-39).  Every likelihood delta of the
// solver loops is then a difference of two shared-memory values: no global loads.
// TRIVIAL windows (WindowDistr::TRIVIAL, distr_cache.rs:27-30) are stored as weight 0 on the all-zero
// extra row of the device table, so 0*0 - 0*0 = +0.0 reproduces the reference's literal 0.0 with no
// branch; a zero depth change gives p - p = +0.0 the same way.
// Layout: structure of arrays (five product planes, then weight, row, depth), so that lanes looking up
// DIFFERENT windows hit different banks -- with one 64-byte record per window every lane of an evaluation
// pass landed in the same two bank groups (16-way conflicts on the hottest loads of the kernel).
struct WinState {
    double *base;       // p[5][wp] | weight[wp] | row[wp] (u32) | depth[wp] (u32)
    uint32_t wp;        // plane stride (windows, even)
    __device__ __forceinline__ double &p(uint32_t w, int k) const { return base[(uint32_t)k * wp + w]; }
    __device__ __forceinline__ double &weight(uint32_t w) const { return base[5u * wp + w]; }
    __device__ __forceinline__ uint32_t &row(uint32_t w) const { return ((uint32_t *)(base + 6u * wp))[w]; }
    __device__ __forceinline__ uint32_t &depth(uint32_t w) const { return ((uint32_t *)(base + 6u * wp))[wp + w]; }
};
__host__ __device__ inline uint32_t win_stride(uint32_t Wmax) { return (Wmax + 1u) & ~1u; }

struct WarpShared {
    WinState win;
    uint16_t *ntc_start;   // [R+1]  compact candidate offset of every non-trivial read (+ end sentinel)
    uint8_t *nt_assgn;     // [R]    current assignment of every non-trivial read
    uint32_t zero_row;     // offset of the all-zero row
    uint32_t depth_k;
};
__host__ __device__ inline size_t group_smem_bytes(uint32_t Wmax, uint32_t R) {
    return (size_t)win_stride(Wmax) * 64 + align_up(((size_t)R + 1) * 2, 16) + align_up((size_t)R, 16);
}

__device__ __forceinline__ double rec_lp(const uint4 &r) { return __hiloint2double((int)r.y, (int)r.x); }
__device__ __forceinline__ uint4 make_rec(double lp, uint32_t w) {
    return make_uint4((uint32_t)__double2loint(lp), (uint32_t)__double2hiint(lp), w, 0u);
}

// Recompute slice entry k of window w from its (weight, row, depth).
__device__ __forceinline__ void win_refresh(const WarpShared &ws, const double *__restrict__ table, uint32_t w, int k) {
    const int d = min(max((int)ws.win.depth(w) + k - 2, 0), (int)ws.depth_k - 1);   // out-of-range entries are never used
    ws.win.p(w, k) = __dmul_rn(ws.win.weight(w), __ldg(table + ws.win.row(w) + d));
}

// atomic_depth_lik_diff (src/model/assgn.rs:244-254), branch-free
__device__ __forceinline__ double atomic_diff(const WarpShared &ws, uint32_t w, int change) {
    return __dsub_rn(ws.win.p(w, 2 + change), ws.win.p(w, 2));
}

// depth_lik_diff (src/model/assgn.rs:259-284): ((a1 + a2) + a3) + a4, window merging done with selects
__device__ __forceinline__ double depth_lik_diff(const WarpShared &ws, uint32_t w12, uint32_t w34) {
    const uint32_t w1 = w12 & 0xFFFFu, w2 = w12 >> 16, w3 = w34 & 0xFFFFu, w4 = w34 >> 16;
    const int e21 = w2 == w1, e31 = w3 == w1, e32 = (w3 == w2) & !e31;
    const int e41 = w4 == w1, e42 = (w4 == w2) & !e41, e43 = (w4 == w3) & !e41 & !e42;
    const int c1 = -1 - e21 + e31 + e41;
    const int c2 = e21 ? 0 : -1 + e32 + e42;
    const int c3 = (e31 | e32) ? 0 : 1 + e43;
    const int c4 = (e41 | e42 | e43) ? 0 : 1;
    double s = __dadd_rn(atomic_diff(ws, w1, c1), atomic_diff(ws, w2, c2));
    s = __dadd_rn(s, atomic_diff(ws, w3, c3));
    return __dadd_rn(s, atomic_diff(ws, w4, c4));
}

// ------------------------------------------------------------------ a5: instance build ----------

struct Instance {
    uint32_t haps[LCTP_MAX_PLOIDY];
    uint32_t wshift[LCTP_MAX_PLOIDY + 1];
    uint32_t W, A, n_nt, A_nt;
};

// GenotypeAlignments::new: per read, gather candidates of every genotype contig above the running
// threshold, append the unmapped option, stable-sort descending (here: a p-way merge of the already
// sorted per-contig lists, ties resolved in insertion order = contig order, unmapped last), cut at the
// final threshold.  Returns false on slab overflow.
// HEADS > 0 (ploidy <= 2): the first HEADS ln-probs of each contig's run are fetched up front with
// independent loads and the threshold / cut / merge run on registers; runs longer than HEADS fall back to
// memory for the tail.  HEADS == 0: everything from memory (any ploidy).
template <int GS, int HEADS>
__device__ bool build_instance(const LocusDev &L, const Slab &S, const WarpShared &ws, uint32_t cap, Instance &I,
                               const Grp<GS> &g) {
    constexpr int PK = HEADS > 0 ? 2 : LCTP_MAX_PLOIDY;
    constexpr int HN = HEADS > 0 ? HEADS : 1;
    const uint32_t R = L.R, p = L.p;
    const int lane = g.lane;
    uint32_t base = 0, nt_base = 0, ntc_base = 0;
    bool ok = true;
    for (uint32_t r0 = 0; r0 < R; r0 += GS) {
        const uint32_t r = r0 + lane;
        const bool valid = r < R;
        uint32_t lb[PK], le[PK], l0[PK];
        double hv[PK][HN];
        double unm = 0.0, thresh = 0.0;
        uint32_t nw = 0;
        bool with_unm = false;
        // value j of contig k's run (j counted from the run start l0[k])
        auto val = [&](int k, uint32_t ix) -> double {
            if (HEADS > 0) {
                const uint32_t j = ix - l0[k];
#pragma unroll
                for (int q = 0; q < HN; q++) if (j == (uint32_t)q) return hv[k][q];
            }
            return L.cm_lnprob[ix];
        };
        if (valid) {
            unm = L.unmapped[r];
            thresh = __dsub_rn(unm, L.prob_diff);
#pragma unroll
            for (int k = 0; k < PK; k++) {
                if ((uint32_t)k < p) {
                    const size_t key = (size_t)I.haps[k] * R + r;
                    lb[k] = l0[k] = L.cm_off[key];
                    le[k] = L.cm_off[key + 1];
                } else lb[k] = le[k] = l0[k] = 0;
            }
            if (HEADS > 0) {
#pragma unroll
                for (int k = 0; k < PK; k++)
#pragma unroll
                    for (int q = 0; q < HN; q++) hv[k][q] = lb[k] + q < le[k] ? L.cm_lnprob[lb[k] + q] : 0.0;
            }
#pragma unroll
            for (int k = 0; k < PK; k++)
                if ((uint32_t)k < p && le[k] > lb[k]) thresh = fmax(thresh, __dsub_rn(val(k, lb[k]), L.prob_diff));
#pragma unroll
            for (int k = 0; k < PK; k++) {
                if ((uint32_t)k < p) {
                    uint32_t e = lb[k];
                    while (e < le[k] && val(k, e) >= thresh) e++;
                    le[k] = e;
                    nw += e - lb[k];
                }
            }
            with_unm = unm >= thresh;
            nw += with_unm ? 1u : 0u;
        }
        const bool nt = valid && nw > 1;
        const uint32_t nw_nt = nt ? nw : 0u;
        uint32_t incl = nw, incl_nt = nw_nt;
#pragma unroll
        for (int d = 1; d < GS; d <<= 1) {
            const uint32_t o = g.shfl_up(incl, d), o2 = g.shfl_up(incl_nt, d);
            if (lane >= d) { incl += o; incl_nt += o2; }
        }
        const uint32_t total = g.shfl(incl, GS - 1), total_nt = g.shfl(incl_nt, GS - 1);
        const uint32_t start = base + incl - nw;
        const uint32_t cstart = ntc_base + incl_nt - nw_nt;
        const unsigned ntmask = g.ballot(nt);
        if (valid) S.read_off[r] = start;
        if (base + total > cap || ntc_base + total_nt > 65535u) ok = false;
        if (ok && valid) {
            if (nt) {
                const uint32_t pos = nt_base + __popc(ntmask & g.lt());
                ws.ntc_start[pos] = (uint16_t)cstart;
                ws.nt_assgn[pos] = 0;
            }
            bool unm_left = with_unm;
            const long long unm_key = total_key(unm);
            for (uint32_t t = 0; t < nw; t++) {
                int bk = -1;
                long long bkey = 0;
                double bval = 0.0;
#pragma unroll
                for (int k = 0; k < PK; k++) {
                    if ((uint32_t)k < p && lb[k] < le[k]) {
                        const double v = val(k, lb[k]);
                        const long long key = total_key(v);
                        if (bk < 0 || key > bkey) { bk = k; bkey = key; bval = v; }
                    }
                }
                double lp;
                uint32_t src;
                if (bk >= 0 && !(unm_left && unm_key > bkey)) {
                    lp = bval;
                    uint32_t ix = 0;
#pragma unroll
                    for (int k = 0; k < PK; k++) if (k == bk) { ix = lb[k]; lb[k]++; }
                    src = ix | ((uint32_t)bk << 28);
                } else {
                    lp = unm;
                    src = LCTP_NONE_U32;
                    unm_left = false;
                }
                // windows stay [UNMAPPED_WINDOW; 2] = 0 until apply_tweak
                if (nt) { S.cmap[start + t] = make_uint2(src, cstart + t); S.ntc[cstart + t] = make_rec(lp, 0u); }
                else { S.cmap[start + t] = make_uint2(src, TRIV_TAG | r); S.triv_lp[r] = lp; S.triv_w[r] = 0u; }
            }
        }
        base += total;
        ntc_base += total_nt;
        nt_base += __popc(ntmask);
    }
    if (lane == 0) { S.read_off[R] = base; if (ok) ws.ntc_start[nt_base] = (uint16_t)ntc_base; }
    I.A = base;
    I.n_nt = nt_base;
    I.A_nt = ntc_base;
    g.sync();
    return ok;
}

// ------------------------------------------------------------------ a6: apply_tweak -------------

// ContigInfo::get_shifted_window_ix (src/model/windows.rs:62-68,465-470)
__device__ __forceinline__ uint32_t shifted_window(const LocusDev &L, uint32_t hap, uint32_t shift, uint32_t middle) {
    const uint32_t start = L.hap_reg_start[hap];
    const uint32_t end = start + L.hap_n_windows[hap] * L.window;
    if (start <= middle && middle < end) return (middle - start) / L.window + shift;
    return 1;   // BOUNDARY_WINDOW
}

template <int GS>
__device__ void apply_tweak(const LocusDev &L, const Slab &S, const Instance &I, const WarpShared &ws, Xo<GS> &rng) {
    const Grp<GS> &g = rng.g;
    const int lane = g.lane;
    const uint32_t tweak = L.tweak;
    const uint32_t span = 2 * tweak + 1;
    // (i) read middles: one next_u64 per candidate that has a parent, in candidate order
    for (uint32_t c0 = 0; c0 < I.A; c0 += GS) {
        const uint32_t c = c0 + lane;
        const uint2 cm = c < I.A ? S.cmap[c] : make_uint2(LCTP_NONE_U32, 0u);
        const bool has_parent = cm.x != LCTP_NONE_U32;
        uint64_t mine = 0;
        if (tweak != 0) {
            const unsigned m = g.ballot(has_parent);
            const int my_rank = __popc(m & g.lt());
            const int n_draws = __popc(m);
            if (stream_cover(rng, (uint32_t)n_draws)) {
                mine = stream_peek64(rng, (uint32_t)my_rank);
                rng.pos += (uint32_t)n_draws;
            } else {
                for (int q = 0; q < n_draws; q++) {
                    const uint64_t v = xo_next(rng);
                    if (q == my_rank) mine = v;
                }
            }
        }
        if (has_parent) {
            const uint32_t k = cm.x >> 28;
            const uint32_t hap = I.haps[k], shift = I.wshift[k];
            const uint2 mid = L.cm_mid[cm.x & SRC_MASK];
            const uint32_t t1 = tweak ? (uint32_t)(mine >> 32) % span : 0u;
            const uint32_t t2 = tweak ? (uint32_t)mine % span : 0u;
            const uint32_t w1 = mid.x == LCTP_NONE_U32 ? 0u : shifted_window(L, hap, shift, mid.x + t1);
            const uint32_t w2 = mid.y == LCTP_NONE_U32 ? 0u : shifted_window(L, hap, shift, mid.y + t2);
            const uint32_t w = w1 | (w2 << 16);
            if (cm.y & TRIV_TAG) S.triv_w[cm.y & ~TRIV_TAG] = w;
            else S.ntc[cm.y].z = w;
        }
    }
    // (ii) window distributions: one bounded i32 draw per window, contigs in genotype order
    if (lane < 2) { ws.win.weight(lane) = 0.0; ws.win.row(lane) = ws.zero_row; ws.win.depth(lane) = 0; }
    for (uint32_t k = 0; k < L.p; k++) {
        const uint32_t hap = I.haps[k];
        const uint32_t nwin = L.hap_n_windows[hap];
        const uint32_t reg_start = L.hap_reg_start[hap], hlen = L.hap_len[hap];
        const uint64_t pos_off = L.hap_pos_off[hap];
        for (uint32_t i0 = 0; i0 < nwin; i0 += GS) {
            const uint32_t cnt = min((uint32_t)GS, nwin - i0);
            uint32_t my_wstart = 0;
            {
                // generate_windows (windows.rs:478-486): random_range(-left..=right), one per window
                const uint32_t start = reg_start + (i0 + min((uint32_t)lane, cnt - 1u)) * L.window;
                const uint32_t end = start + L.window;
                const uint32_t left = min(tweak, start), right = min(tweak, hlen - end);
                uint32_t rr;
                if (xo_below_parallel(rng, cnt, left + right + 1u, rr)) my_wstart = start + rr - left;
                else {
                    for (uint32_t q = 0; q < cnt; q++) {
                        const uint32_t st_q = reg_start + (i0 + q) * L.window;
                        const uint32_t l_q = min(tweak, st_q), r_q = min(tweak, hlen - (st_q + L.window));
                        const uint32_t v = xo_below(rng, l_q + r_q + 1u);
                        if ((uint32_t)lane == q) my_wstart = st_q + v - l_q;
                    }
                }
            }
            if ((uint32_t)lane < cnt) {
                // neighb_info (windows.rs:439-445) + assgn.rs:144-148 + get_distribution (distr_cache.rs:83-92)
                const uint32_t idx = my_wstart > L.left_padding ? my_wstart - L.left_padding : 0u;
                const double weight = L.pos_weight[pos_off + idx];
                const uint32_t gc = L.pos_gc[pos_off + idx];
                const uint32_t w = I.wshift[k] + i0 + lane;
                const bool trivial = weight < L.min_weight || weight < 1e-7;
                ws.win.weight(w) = trivial ? 0.0 : weight;
                ws.win.row(w) = trivial ? ws.zero_row : gc * L.depth_k;
                ws.win.depth(w) = 0;
            }
        }
    }
    g.sync();
}

// ------------------------------------------------------------------ a8: ReadAssignment::new -----

// Sequentially (in index order) add `count` per-lane terms to acc: reproduces iter().sum() order.
template <int GS>
__device__ __forceinline__ void seq_add(const Grp<GS> &g, double &acc, double term, int count) {
    for (int q = 0; q < count; q++) acc = __dadd_rn(acc, g.shfl(term, q));
}

// init_mode 0: every read at candidate 0; 1: random_range(0..m) per non-trivial read (read order).
template <int GS>
__device__ void init_assignment(const LocusDev &L, const Slab &S, const Instance &I, const WarpShared &ws,
                                Xo<GS> &rng, int init_mode, double &aln_lik, double &depth_lik) {
    const Grp<GS> &g = rng.g;
    const int lane = g.lane;
    for (uint32_t w = lane; w < I.W; w += GS) ws.win.depth(w) = 0;
    // assignments of non-trivial reads
    for (uint32_t i0 = 0; i0 < I.n_nt; i0 += GS) {
        const uint32_t i = i0 + lane;
        const uint32_t my_n = i < I.n_nt ? (uint32_t)(ws.ntc_start[i + 1] - ws.ntc_start[i]) : 1u;
        uint32_t a = 0;
        if (init_mode == 1) {
            const int cnt = (int)min((uint32_t)GS, I.n_nt - i0);
            if (!xo_below_parallel(rng, (uint32_t)cnt, my_n, a)) {
                for (int q = 0; q < cnt; q++) {
                    const uint32_t m = g.shfl(my_n, q);
                    const uint32_t v = xo_below(rng, m);
                    if (lane == q) a = v;
                }
            }
        }
        if (i < I.n_nt) ws.nt_assgn[i] = (uint8_t)a;
    }
    g.sync();
    // depth counts + aln_lik in read order (src/model/assgn.rs:205-217,351-353)
    double al = 0.0;
    uint32_t nt_base = 0;
    for (uint32_t r0 = 0; r0 < L.R; r0 += GS) {
        const uint32_t r = r0 + lane;
        const bool valid = r < L.R;
        uint32_t start = 0, nw = 0;
        if (valid) { start = S.read_off[r]; nw = S.read_off[r + 1] - start; }
        const bool nt = valid && nw > 1;
        const unsigned ntmask = g.ballot(nt);
        double term = 0.0;
        if (valid) {
            uint32_t w12;
            if (nt) {
                const uint32_t pos = nt_base + __popc(ntmask & g.lt());
                const uint4 rec = S.ntc[(uint32_t)ws.ntc_start[pos] + ws.nt_assgn[pos]];
                term = rec_lp(rec);
                w12 = rec.z;
            } else {
                term = S.triv_lp[r];
                w12 = S.triv_w[r];
            }
            atomicAdd(&ws.win.depth(w12 & 0xFFFFu), 1u);
            atomicAdd(&ws.win.depth(w12 >> 16), 1u);
        }
        nt_base += __popc(ntmask);
        seq_add(g, al, term, (int)min((uint32_t)GS, L.R - r0));
    }
    g.sync();
    for (uint32_t q = lane; q < I.W * 5u; q += GS) win_refresh(ws, L.depth_table, q / 5u, (int)(q % 5u));
    g.sync();
    // depth_lik in window order (assgn.rs:347-350)
    double dl = 0.0;
    for (uint32_t w0 = 0; w0 < I.W; w0 += GS) {
        const uint32_t w = w0 + lane;
        const double term = w < I.W ? ws.win.p(w, 2) : 0.0;
        seq_add(g, dl, term, (int)min((uint32_t)GS, I.W - w0));
    }
    aln_lik = al;
    depth_lik = dl;
}

// ------------------------------------------------------------------ a9: targets -----------------

struct Target { uint32_t idx, n, old_a, new_a, old_ix, new_ix; };   // *_ix index the compact arrays

// ReassignmentTarget::random (src/model/assgn.rs:451-471), group-uniform
template <int GS>
__device__ __forceinline__ Target random_target(const WarpShared &ws, const Instance &I, Xo<GS> &rng) {
    Target t;
    t.idx = xo_below(rng, I.n_nt);                     // random_range(0..n_nontrivial), usize via the u32 path
    const uint32_t start = ws.ntc_start[t.idx];
    t.n = ws.ntc_start[t.idx + 1] - start;
    t.old_a = ws.nt_assgn[t.idx];
    if (t.n == 2) t.new_a = 1u - t.old_a;
    else {
        const uint32_t i = 1u + xo_below(rng, t.n - 1u);   // random_range(1..n as u16)
        t.new_a = i <= t.old_a ? i - 1u : i;
    }
    t.old_ix = start + t.old_a;
    t.new_ix = start + t.new_a;
    return t;
}

struct Move { double dld, dlp; uint32_t w12, w34; };

// calculate_improvement (src/model/assgn.rs:321-328)
__device__ __forceinline__ double calc_improvement(const LocusDev &L, const Slab &S, const WarpShared &ws,
                                                   const Target &t, Move &mv) {
    const uint4 ro = __ldcg(S.ntc + t.old_ix), rn = __ldcg(S.ntc + t.new_ix);
    mv.w12 = ro.z;
    mv.w34 = rn.z;
    mv.dld = depth_lik_diff(ws, mv.w12, mv.w34);
    mv.dlp = __dsub_rn(rec_lp(rn), rec_lp(ro));
    return __dadd_rn(__dmul_rn(L.depth_contrib, mv.dld), __dmul_rn(L.aln_contrib, mv.dlp));
}

// reassign (src/model/assgn.rs:331-343); group-uniform inputs, lane 0 writes
template <int GS>
__device__ __forceinline__ void apply_move(const Grp<GS> &g, const WarpShared &ws, const double *__restrict__ table,
                                           uint32_t idx, uint32_t new_a, const Move &mv, double &aln_lik,
                                           double &depth_lik) {
    depth_lik = __dadd_rn(depth_lik, mv.dld);
    aln_lik = __dadd_rn(aln_lik, mv.dlp);
    if (g.lane == 0) {
        ws.win.depth(mv.w34 & 0xFFFFu) += 1;
        ws.win.depth(mv.w34 >> 16) += 1;
        ws.win.depth(mv.w12 & 0xFFFFu) -= 1;
        ws.win.depth(mv.w12 >> 16) -= 1;
        ws.nt_assgn[idx] = (uint8_t)new_a;
    }
    g.sync();
    // slide the product slices of the (up to four) windows whose depth changed
    for (int q = g.lane; q < 20; q += GS) {
        const int j = q / 5;
        const uint32_t w = j == 0 ? (mv.w12 & 0xFFFFu) : j == 1 ? (mv.w12 >> 16) : j == 2 ? (mv.w34 & 0xFFFFu) : (mv.w34 >> 16);
        win_refresh(ws, table, w, q % 5);
    }
    g.sync();
}

// max_abs_random (src/solvers/stoch.rs:19-22) with INIT_ITER = 100
template <int GS>
__device__ double max_abs_random(const LocusDev &L, const Slab &S, const Instance &I, const WarpShared &ws, Xo<GS> &rng) {
    double acc = 0.0;
    for (int q = 0; q < 100; q++) {
        const Target t = random_target(ws, I, rng);
        Move mv;
        acc = fmax(acc, fabs(calc_improvement(L, S, ws, t, mv)));
    }
    return acc;
}

// ------------------------------------------------------------------ a10: Greedy -----------------

// Lane layout of the greedy loop: the `amount` sampled reads ("slots") each own LPS = GS / amount
// consecutive lanes; lane (slot, rank) evaluates the rank-th alternative candidate of its slot's read
// (further alternatives, when a read has more than LPS of them, in extra passes).  The mapping is fixed
// for the whole solve, so an iteration needs no scan / flattening, and the candidate records of the NEXT
// sample are prefetched into registers while the current sample is evaluated.
struct SlotMap {
    uint32_t lps, slot, rank, range;     // range = j_slot + 1 of Floyd's draw for this slot
    bool valid;
    unsigned lead_mask;                  // group-relative mask of the rank-0 lanes of the valid slots
};

// One sample of `amount` distinct non-trivial reads (IndexedRandom::sample -> index::sample_floyd):
// draw k is random_range(..=j_k), j_k = n_nt - amount + k; a draw equal to an earlier entry replaces
// that entry by j_k.  Every lane of slot k receives entry k.  `fast_only`: succeed only through the
// lane-parallel path (no refill, no biased draw), so the caller can un-consume the draws again with
// `rng.pos -= amount`.
template <int GS>
__device__ __forceinline__ bool sample_reads(Xo<GS> &rng, const SlotMap &sm, uint32_t n_nt, uint32_t amount,
                                             bool fast_only, uint32_t &myv) {
    const Grp<GS> &g = rng.g;
    bool fast = false;
    if (stream_cover(rng, amount)) {
        const uint64_t m = (uint64_t)stream_peek_hi(rng, sm.slot) * (uint64_t)sm.range;
        if (!g.any(sm.valid && (uint32_t)m > 0u - sm.range)) {
            myv = (uint32_t)(m >> 32);
            rng.pos += amount;
            fast = true;
        }
    }
    if (!fast) {
        if (fast_only) return false;
        for (uint32_t k = 0; k < amount; k++) {
            const uint32_t t = xo_below(rng, n_nt - amount + k + 1u);
            if (sm.slot == k) myv = t;
        }
    }
    const unsigned peers = g.match_any(sm.valid ? myv : 0xFFFFFFFFu);
    if (g.any(sm.valid && (peers & sm.lead_mask) != (1u << (sm.slot * sm.lps)))) {
        for (uint32_t k = 1; k < amount; k++) {
            const uint32_t t = g.shfl(myv, (int)(k * sm.lps));
            if (sm.slot < k && myv == t) myv = n_nt - amount + k;
        }
    }
    return true;
}

// A candidate move as seen by one lane, ordered like the reference's two nested strict-'>' scans:
// larger improvement `s` first; among equal `s` the earlier sampled read (slot); inside that read the
// larger `improv` (pre-scaling value compared by best_read_improvement), then the lower candidate index.
struct Cand {
    double s, improv, dld, dlp;
    uint32_t slot, c, w12, w34;
};
__device__ __forceinline__ bool cand_better(const Cand &a, const Cand &b) {
    if (a.s != b.s) return a.s > b.s;
    if (a.slot != b.slot) return a.slot < b.slot;
    if (a.improv != b.improv) return a.improv > b.improv;
    return a.c < b.c;
}

// The sampled read of this lane's slot and the two candidate records the lane works on.
struct SlotRead {
    uint32_t idx, start, n, old_a;
    uint4 ro, rn;       // current candidate, and alternative #rank (valid when rank < n - 1)
};
template <int GS>
__device__ __forceinline__ void load_slot(const Slab &S, const WarpShared &ws, const SlotMap &sm, uint32_t idx,
                                          SlotRead &r) {
    r.idx = idx;
    r.start = ws.ntc_start[idx];
    r.n = ws.ntc_start[idx + 1] - r.start;
    r.old_a = ws.nt_assgn[idx];
    const uint32_t c = sm.rank < r.old_a ? sm.rank : sm.rank + 1u;
    r.ro = __ldcg(S.ntc + r.start + r.old_a);
    r.rn = __ldcg(S.ntc + r.start + min(c, r.n - 1u));
}

__device__ __forceinline__ void eval_cand(const LocusDev &L, const WarpShared &ws, const uint4 &ro, const uint4 &rn,
                                          uint32_t slot, uint32_t c, Cand &cd) {
    const double lp_old = rec_lp(ro), lp = rec_lp(rn);
    cd.dld = depth_lik_diff(ws, ro.z, rn.z);
    cd.improv = __dadd_rn(lp, __dmul_rn(L.rel_contrib, cd.dld));          // assgn.rs:303
    cd.s = __dmul_rn(L.aln_contrib, __dsub_rn(cd.improv, lp_old));        // assgn.rs:310
    cd.dlp = __dsub_rn(lp, lp_old);
    cd.slot = slot; cd.c = c; cd.w12 = ro.z; cd.w34 = rn.z;
}

// Greedy::solve_nontrivial (src/solvers/stoch.rs:81-120).
// Every (sampled read, alternative candidate) pair of an iteration is one lane's job ("flattened"
// best_read_improvement, src/model/assgn.rs:287-317); the winner is found with REDUX reductions in the
// reference's tie order.  Software pipeline: the sample of iteration i+1 is drawn and its (static)
// candidate records are loaded into registers while iteration i is evaluated, so the evaluation itself
// touches only registers and shared memory.
template <int GS>
__device__ void greedy_solve(const LocusDev &L, const StageParams &P, const Slab &S, const Instance &I,
                             const WarpShared &ws, Xo<GS> &rng, double &aln_lik, double &depth_lik,
                             uint64_t &iters_out) {
    const Grp<GS> &g = rng.g;
    const uint32_t lane = (uint32_t)g.lane;
    const uint32_t amount = min(P.sample_size, I.n_nt);
    init_assignment(L, S, I, ws, rng, P.best_start ? 0 : 1, aln_lik, depth_lik);
    const double min_diff = fmax(__dmul_rn(1e-10, max_abs_random(L, S, I, ws, rng)), 1e-14);
    SlotMap sm;
    sm.lps = (uint32_t)GS / amount;
    {
        const uint32_t s = lane / sm.lps;
        sm.valid = s < amount;
        sm.slot = sm.valid ? s : amount - 1u;
        sm.rank = lane - s * sm.lps;
        sm.range = I.n_nt - amount + sm.slot + 1u;
        sm.lead_mask = g.ballot(sm.valid && sm.rank == 0u);
    }
    uint64_t curr_plato = 0, it = 0;
    // Sample pipeline: `nx` = sample of iteration i+1 with its records in registers, `idx2` = sample of
    // iteration i+2 (drawn, its slab lines prefetched into L2).
    bool have1 = false, have2 = false;
    uint32_t idx2 = 0;
    SlotRead nx;
    nx.idx = nx.start = nx.old_a = 0; nx.n = 1; nx.ro = nx.rn = make_uint4(0, 0, 0, 0);
    for (; it < P.max_iter; it++) {
        SlotRead cur;
        if (have1) cur = nx;
        else {
            uint32_t v = 0;
            sample_reads(rng, sm, I.n_nt, amount, false, v);
            load_slot<GS>(S, ws, sm, v, cur);
        }
        if (have2) { load_slot<GS>(S, ws, sm, idx2, nx); have1 = true; have2 = false; }
        else {
            uint32_t v = 0;
            have1 = sample_reads(rng, sm, I.n_nt, amount, true, v);
            if (have1) load_slot<GS>(S, ws, sm, v, nx);
        }
        if (have1) {
            have2 = sample_reads(rng, sm, I.n_nt, amount, true, idx2);
            if (have2 && sm.rank < 2u) {
                const uint32_t st2 = ws.ntc_start[idx2];
                const uint32_t off2 = sm.rank == 0u ? 0u : (uint32_t)ws.ntc_start[idx2 + 1] - st2 - 1u;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(S.ntc + st2 + off2));
            }
        }
        Cand best;
        best.s = -INFINITY; best.improv = -INFINITY; best.dld = 0.0; best.dlp = 0.0;
        best.slot = 0xFFFFu; best.c = 0; best.w12 = 0; best.w34 = 0;
        const uint32_t n_alt = cur.n - 1u;
        // Keep the (unused) padding word of the prefetched records live until here: otherwise its register
        // is recycled right after the load is issued and that write has to wait for the load (WAW).
        asm volatile("" ::"r"(cur.ro.w), "r"(cur.rn.w));
        if (sm.valid && sm.rank < n_alt)
            eval_cand(L, ws, cur.ro, cur.rn, sm.slot, sm.rank < cur.old_a ? sm.rank : sm.rank + 1u, best);
        if (g.any(sm.valid && n_alt > sm.lps)) {
            // reads with more alternatives than lanes per slot: extra passes straight from the slab
            for (uint32_t j = sm.rank + sm.lps; g.any(sm.valid && j < n_alt); j += sm.lps) {
                if (sm.valid && j < n_alt) {
                    const uint32_t c = j < cur.old_a ? j : j + 1u;
                    const uint4 rn = __ldcg(S.ntc + cur.start + c);
                    Cand cd;
                    eval_cand(L, ws, cur.ro, rn, sm.slot, c, cd);
                    if (cand_better(cd, best)) best = cd;
                }
            }
        }
        // winner over the lanes' bests, in the order of cand_better
        int wl;
        {
            const unsigned long long ks = ord_key(best.s);
            const uint32_t h1 = g.rmax((uint32_t)(ks >> 32));
            bool m = (uint32_t)(ks >> 32) == h1;
            const uint32_t l1 = g.rmax(m ? (uint32_t)ks : 0u);
            m = m && (uint32_t)ks == l1;
            const unsigned tied = g.ballot(m);
            if ((tied & (tied - 1u)) == 0u) wl = __ffs(tied) - 1;      // unique maximum (the usual case)
            else {
                const uint32_t sl = g.rmin(m ? best.slot : 0xFFFFFFFFu);
                m = m && best.slot == sl;
                const unsigned long long ki = ord_key(best.improv);
                const uint32_t h2 = g.rmax(m ? (uint32_t)(ki >> 32) : 0u);
                m = m && (uint32_t)(ki >> 32) == h2;
                const uint32_t l2 = g.rmax(m ? (uint32_t)ki : 0u);
                m = m && (uint32_t)ki == l2;
                const uint32_t cm = g.rmin(m ? best.c : 0xFFFFFFFFu);
                wl = __ffs(g.ballot(m && best.c == cm)) - 1;
            }
        }
        const double s_best = g.shfl(best.s, wl);
        if (s_best > min_diff) {
            Move mv;
            mv.dld = g.shfl(best.dld, wl);
            mv.dlp = g.shfl(best.dlp, wl);
            mv.w12 = g.shfl(best.w12, wl);
            mv.w34 = g.shfl(best.w34, wl);
            const uint32_t w_c = g.shfl(best.c, wl);
            const uint32_t w_idx = g.shfl(cur.idx, wl);
            apply_move(g, ws, L.depth_table, w_idx, w_c, mv, aln_lik, depth_lik);
            curr_plato = 0;
            // the prefetched sample saw the old assignment of the moved read
            if (have1 && g.any(sm.valid && nx.idx == w_idx)) load_slot<GS>(S, ws, sm, nx.idx, nx);
        } else {
            curr_plato += 1;
            if (curr_plato > P.plato_size) { it++; break; }
        }
    }
    if (have2) stream_unconsume(rng, amount);   // the pre-drawn samples of iterations that never ran
    if (have1) stream_unconsume(rng, amount);
    iters_out += it;
}

// ------------------------------------------------------------------ a11: SimAnneal --------------

// SimAnneal::solve_nontrivial (src/solvers/stoch.rs:197-242): one candidate per step, group-uniform.
template <int GS>
__device__ void anneal_solve(const LocusDev &L, const StageParams &P, const Slab &S, const Instance &I,
                             const WarpShared &ws, Xo<GS> &rng, double &aln_lik, double &depth_lik,
                             uint64_t &iters_out) {
    const Grp<GS> &g = rng.g;
    init_assignment(L, S, I, ws, rng, 1, aln_lik, depth_lik);
    const double max_abs = max_abs_random(L, S, I, ws, rng);
    const double min_diff = fmax(__dmul_rn(1e-10, max_abs), 1e-14);
    const double start_temp = fmax(__ddiv_rn(-max_abs, P.ln_init_prob), 1e-5);
    const double temp_step = __ddiv_rn(start_temp, (double)P.anneal_steps);
    uint64_t curr_plato = 0, steps = 0;
    for (uint64_t i = P.anneal_steps; i >= 1; i--) {
        const Target t = random_target(ws, I, rng);
        Move mv;
        const double diff = __dsub_rn(calc_improvement(L, S, ws, t, mv), min_diff);
        steps++;
        bool accept = diff >= 0.0;
        if (!accept) {
            const double u = xo_f64(rng);
            accept = u <= exp(__ddiv_rn(diff, __dmul_rn(temp_step, (double)i)));
        }
        if (accept) { apply_move(g, ws, L.depth_table, t.idx, t.new_a, mv, aln_lik, depth_lik); curr_plato = 0; }
        else { curr_plato += 1; if (curr_plato >= P.plato_size) break; }
    }
    for (uint64_t k = 0; k < P.max_iter; k++) {
        if (curr_plato >= P.plato_size) break;
        const Target t = random_target(ws, I, rng);
        Move mv;
        const double diff = calc_improvement(L, S, ws, t, mv);
        steps++;
        if (diff > min_diff) { apply_move(g, ws, L.depth_table, t.idx, t.new_a, mv, aln_lik, depth_lik); curr_plato = 0; }
        else curr_plato += 1;
    }
    iters_out += steps;
}

// ------------------------------------------------------------------ stage kernel ----------------

#ifndef LCTP_MIN_CTAS
#define LCTP_MIN_CTAS 17
#endif
template <int GS>
__global__ void __launch_bounds__(CTA_THREADS, LCTP_MIN_CTAS)
k_solve_stage(LocusDev L, StageParams P, const uint64_t *__restrict__ worker_ixs,
              const uint64_t *__restrict__ worker_off, const uint32_t *__restrict__ tuples,
              uint64_t *__restrict__ rng_states, double *__restrict__ lik_mean, double *__restrict__ lik_var,
              double *__restrict__ liks, uint64_t *__restrict__ n_alns, uint64_t *__restrict__ iters,
              uint16_t *__restrict__ counts, unsigned char *__restrict__ scratch,
              unsigned int *__restrict__ work_counter, int *__restrict__ err,
              const uint64_t *__restrict__ setup_mats) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int GROUPS = CTA_THREADS / GS;
    const int gib = threadIdx.x / GS;
    Xo<GS> rng;
    const Grp<GS> &g = rng.g;
    const int lane = g.lane;
    WarpShared ws;
    {
        unsigned char *base = smem + (size_t)gib * group_smem_bytes(P.Wmax, L.R);
        ws.win.base = (double *)base;          ws.win.wp = win_stride(P.Wmax);
        base += (size_t)ws.win.wp * 64;
        ws.ntc_start = (uint16_t *)base;       base += align_up(((size_t)L.R + 1) * 2, 16);
        ws.nt_assgn = (uint8_t *)base;
        ws.zero_row = LCTP_GC_BINS * L.depth_k;
        ws.depth_k = L.depth_k;
    }
    Slab S;
    slab_layout(P.cap, L.R, scratch + (size_t)(blockIdx.x * GROUPS + gib) * P.slab_bytes, &S);
    rng.buf = S.rng_buf; rng.blk = S.rng_blk;

    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(work_counter, 1u);
        w = g.shfl(w, 0);
        if (w >= P.n_workers) break;
        stream_begin(rng, rng_states + 4 * (size_t)w, setup_mats);
        for (uint64_t j = worker_off[w]; j < worker_off[w + 1]; j++) {
            const uint64_t gt = worker_ixs[j];
            const double prior = L.priors ? L.priors[gt] : 0.0;
            Instance I;
            uint32_t wsft = 2;
            I.wshift[0] = wsft;
            for (uint32_t k = 0; k < L.p; k++) {
                I.haps[k] = tuples[j * L.p + k];
                wsft += L.hap_n_windows[I.haps[k]];
                I.wshift[k + 1] = wsft;
            }
            I.W = wsft;
            const bool ok = L.p <= 2 ? build_instance<GS, LCTP_HEADS>(L, S, ws, P.cap, I, g)
                                     : build_instance<GS, 0>(L, S, ws, P.cap, I, g);
            if (!ok) {
                if (lane == 0) { atomicOr(err, 1); lik_mean[j] = NAN; lik_var[j] = NAN; n_alns[j] = I.A; iters[j] = 0; }
                continue;
            }
            uint16_t *cnt = P.want_counts ? counts + (size_t)j * P.cap : nullptr;
            if (cnt) { for (uint32_t c = lane; c < I.A; c += GS) cnt[c] = 0; }
            uint64_t it_total = 0;
            for (uint32_t a = 0; a < P.attempts; a++) {
                apply_tweak(L, S, I, ws, rng);
                double aln_lik = 0.0, depth_lik = 0.0;
                if (I.n_nt == 0) init_assignment(L, S, I, ws, rng, 0, aln_lik, depth_lik);
                else if (P.kind == 0) greedy_solve(L, P, S, I, ws, rng, aln_lik, depth_lik, it_total);
                else anneal_solve(L, P, S, I, ws, rng, aln_lik, depth_lik, it_total);
                // likelihood (assgn.rs:235-237) + prior (solve.rs:1126)
                const double lik = __dadd_rn(prior, __dadd_rn(__dmul_rn(L.depth_contrib, depth_lik),
                                                              __dmul_rn(L.aln_contrib, aln_lik)));
                if (lane == 0) liks[j * P.attempts + a] = lik;
                if (cnt) {   // update_counts (assgn.rs:374-378)
                    g.sync();
                    uint32_t nt_base = 0;
                    for (uint32_t r0 = 0; r0 < L.R; r0 += GS) {
                        const uint32_t r = r0 + lane;
                        const bool valid = r < L.R;
                        uint32_t start = 0, nw = 0;
                        if (valid) { start = S.read_off[r]; nw = S.read_off[r + 1] - start; }
                        const bool nt = valid && nw > 1;
                        const unsigned ntmask = g.ballot(nt);
                        if (valid) {
                            uint32_t as = 0;
                            if (nt) as = ws.nt_assgn[nt_base + __popc(ntmask & g.lt())];
                            cnt[start + as] += 1;
                        }
                        nt_base += __popc(ntmask);
                    }
                }
                g.sync();
            }
            g.sync();
            // mean_variance_or_nan (ext/vec.rs:74-78,86-93,109-116)
            if (lane == 0) {
                const double *x = liks + j * P.attempts;
                double s = 0.0;
                for (uint32_t a = 0; a < P.attempts; a++) s = __dadd_rn(s, x[a]);
                const double mean = __ddiv_rn(s, (double)P.attempts);
                double var = NAN;
                if (P.attempts > 1) {
                    double acc = 0.0;
                    for (uint32_t a = 0; a < P.attempts; a++) {
                        const double d = __dsub_rn(x[a], mean);
                        acc = __dadd_rn(acc, __dmul_rn(d, d));
                    }
                    var = __ddiv_rn(acc, (double)(P.attempts - 1));
                }
                lik_mean[j] = mean;
                lik_var[j] = var;
                n_alns[j] = I.A;
                iters[j] = it_total;
            }
            g.sync();
        }
        stream_end(rng, rng_states + 4 * (size_t)w);
        g.sync();
    }
}

// ------------------------------------------------------------------ host: jump matrices ---------

// 256x256 matrices over GF(2) stored by columns (4 words per column).
struct BitMat { uint64_t c[256][4]; };

static void bm_step(BitMat &T) {      // one xoshiro256++ state transition
    for (int j = 0; j < 256; j++) {
        Gen g = {0, 0, 0, 0};
        (j < 64 ? g.s0 : j < 128 ? g.s1 : j < 192 ? g.s2 : g.s3) = 1ull << (j & 63);
        gen_next(g);
        T.c[j][0] = g.s0; T.c[j][1] = g.s1; T.c[j][2] = g.s2; T.c[j][3] = g.s3;
    }
}
static void bm_mul(const BitMat &A, const BitMat &B, BitMat &out) {   // out = A * B
    for (int j = 0; j < 256; j++) {
        uint64_t a[4] = {0, 0, 0, 0};
        for (int i = 0; i < 256; i++)
            if ((B.c[j][i >> 6] >> (i & 63)) & 1ull) { a[0] ^= A.c[i][0]; a[1] ^= A.c[i][1]; a[2] ^= A.c[i][2]; a[3] ^= A.c[i][3]; }
        out.c[j][0] = a[0]; out.c[j][1] = a[1]; out.c[j][2] = a[2]; out.c[j][3] = a[3];
    }
}
static void bm_pow(const BitMat &T, unsigned e, BitMat &out) {
    static BitMat base, acc, tmp;
    base = T;
    for (int j = 0; j < 256; j++) for (int k = 0; k < 4; k++) acc.c[j][k] = (k == (j >> 6)) ? 1ull << (j & 63) : 0;
    while (e) {
        if (e & 1) { bm_mul(base, acc, tmp); acc = tmp; }
        bm_mul(base, base, tmp); base = tmp;
        e >>= 1;
    }
    out = acc;
}

// The jump matrices depend only on the generator: computed once per process (contexts on several host
// threads share them), uploaded once per context.
struct RngMats {
    std::vector<uint64_t> setup;     // N_SETUP_MATS matrices T^(C*2^k)
    BitMat refill[2];                // T^(15*C), T^(31*C)
};
static const RngMats &rng_mats() {
    static std::once_flag once;
    static RngMats m;
    std::call_once(once, [] {
        static BitMat T, M;
        bm_step(T);
        m.setup.resize((size_t)N_SETUP_MATS * 1024);
        for (int k = 0; k < N_SETUP_MATS; k++) {
            bm_pow(T, (unsigned)RNG_C << k, M);
            std::memcpy(&m.setup[(size_t)k * 1024], M.c, sizeof(M.c));
        }
        bm_pow(T, 15u * RNG_C, m.refill[0]);
        bm_pow(T, 31u * RNG_C, m.refill[1]);
    });
    return m;
}

static int ensure_rng_mats(lctp_ctx *ctx) {
    if (ctx->d_rng_mats.p) return LCTP_OK;
    const RngMats &m = rng_mats();
    static std::mutex symbol_mutex;      // c_refill_mat is one symbol per device
    std::lock_guard<std::mutex> lock(symbol_mutex);
    int rc = ctx->d_rng_mats.alloc(m.setup.size());
    if (rc) return rc;
    LCTP_CUDA_CHECK(cudaMemcpyAsync(ctx->d_rng_mats.p, m.setup.data(), m.setup.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    LCTP_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_refill_mat, m.refill[0].c, sizeof(m.refill[0].c), 0, cudaMemcpyHostToDevice, ctx->stream));
    LCTP_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_refill_mat, m.refill[1].c, sizeof(m.refill[1].c), sizeof(m.refill[0].c), cudaMemcpyHostToDevice, ctx->stream));
    LCTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return LCTP_OK;
}

// ------------------------------------------------------------------ host launch -----------------

static double dbg_now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <int GS>
static int launch_stage_gs(lctp_locus_h *h, const StageParams &P, size_t n_workers, bool want_counts) {
    lctp_ctx *ctx = h->ctx;
    cudaStream_t s = ctx->stream;
    const LocusDev &L = h->dev;
    constexpr int GROUPS = CTA_THREADS / GS;
    const size_t smem = (size_t)GROUPS * group_smem_bytes(P.Wmax, L.R);
    if (smem > ctx->smem_optin) {
        set_error("lctp_solve_stage: %zu bytes of shared memory per CTA needed (R=%u reads, %u windows); "
                  "loci this large are not supported by the shared-memory resident solver yet", smem, L.R, P.Wmax);
        return LCTP_E_CAPACITY;
    }
    auto kern = k_solve_stage<GS>;
    // function attributes are per-device state shared by every context: configure + launch under one lock
    static std::mutex launch_mutex;
    std::lock_guard<std::mutex> lock(launch_mutex);
    // Only raise the limit when needed: re-setting a function attribute makes the next launch of the function wait
    // for its running instances, which serialised the stage kernels of loci in flight on different contexts.
    static size_t smem_limit[16] = {0};
    size_t &lim = smem_limit[(ctx->device & 7) * 2 + (GS == 32 ? 1 : 0)];
    if (smem > 48 * 1024 && smem > lim) {
        LCTP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lim = smem;
    }
    if (const char *e = getenv("LCTP_CARVEOUT"))   // tuning knob: shared-memory carveout percentage
        LCTP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
    int occ = 0;
    LCTP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CTA_THREADS, smem));
    if (occ < 1) occ = 1;
    if (const char *e = getenv("LCTP_MAX_CTAS_PER_SM")) occ = std::min(occ, std::max(1, atoi(e)));
    uint32_t resident = (uint32_t)ctx->sm_count * occ * GROUPS;
    if (ctx->max_resident_workers && ctx->max_resident_workers < resident) resident = ctx->max_resident_workers;
    const uint32_t groups = (uint32_t)std::min<size_t>(n_workers, resident);
    const uint32_t grid = (groups + GROUPS - 1) / GROUPS;

    int rc;
    if ((rc = ctx->scratch.ensure((size_t)grid * GROUPS * P.slab_bytes))) return rc;
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[2], s));
    kern<<<grid, CTA_THREADS, smem, s>>>(
        L, P, ctx->d_worker_ixs.p, ctx->d_worker_off.p, ctx->d_tuples.p, ctx->d_rng.p, ctx->d_lik_mean.p,
        ctx->d_lik_var.p, ctx->d_liks.p, ctx->d_nalns.p, ctx->d_iters.p, want_counts ? ctx->d_counts.p : nullptr,
        ctx->scratch.p, (unsigned int *)(ctx->d_flags.p + 1), ctx->d_flags.p, ctx->d_rng_mats.p);
    ctx->launches++;
    LCTP_CUDA_CHECK(cudaGetLastError());
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[3], s));
    return LCTP_OK;
}

int launch_stage(lctp_locus_h *h, const lctp_stage *st, const uint64_t *worker_ixs, const uint64_t *worker_off,
                 size_t n_workers, uint64_t *worker_rng, double *lik_mean, double *lik_var, double *liks,
                 uint64_t *counts_off, uint16_t *counts, uint64_t counts_cap, uint64_t *n_alns_out,
                 uint64_t *iters_out) {
    lctp_ctx *ctx = h->ctx;
    cudaStream_t s = ctx->stream;
    const LocusDev &L = h->dev;
    const double t_enter = dbg_now();
    if (!st || !worker_ixs || !worker_off || !worker_rng || !lik_mean || !lik_var || n_workers == 0) {
        set_error("lctp_solve_stage: NULL argument");
        return LCTP_E_INVALID;
    }
    if (st->kind > 1 || st->attempts == 0 || st->attempts > 65535) {
        set_error("lctp_solve_stage: invalid stage (kind=%u attempts=%u)", st->kind, st->attempts);
        return LCTP_E_INVALID;
    }
    if (st->kind == 0 && (st->sample_size == 0 || st->sample_size > MAX_SAMPLE)) {
        // rand::seq::index::sample switches from Floyd's to the in-place algorithm for amount > 11 on
        // short lists; only the Floyd branch is implemented on the device.
        set_error("lctp_solve_stage: greedy sample size %llu unsupported on device (1..=11)",
                  (unsigned long long)st->sample_size);
        return LCTP_E_CAPACITY;
    }
    if (st->kind == 1 && (!(st->init_prob > 0.0 && st->init_prob <= 1.0) || st->anneal_steps == 0)) {
        set_error("lctp_solve_stage: invalid annealing parameters");
        return LCTP_E_INVALID;
    }
    if (h->npa > (uint64_t)SRC_MASK) {
        set_error("lctp_solve_stage: %llu pair alignments exceed the 2^28 the candidate map can address",
                  (unsigned long long)h->npa);
        return LCTP_E_CAPACITY;
    }
    const size_t n = (size_t)worker_off[n_workers];
    if (n == 0 || n_workers > 0xFFFFFFF0ull) { set_error("lctp_solve_stage: empty stage"); return LCTP_E_INVALID; }
    const uint32_t p = L.p;

    // genotype tuples by position + candidate capacity bound: A(g) <= sum_k #alns(h_k) + R
    std::vector<uint32_t> tuples(n * p);
    uint64_t cap64 = 0;
    for (size_t j = 0; j < n; j++) {
        const uint64_t g = worker_ixs[j];
        if (g >= L.G) { set_error("lctp_solve_stage: genotype id %llu out of range", (unsigned long long)g); return LCTP_E_INVALID; }
        genotype_tuple(L.H, p, h->gt_tuples_host.empty() ? nullptr : h->gt_tuples_host.data(), g, &tuples[j * p]);
        uint64_t a = L.R;
        for (uint32_t k = 0; k < p; k++) a += h->hap_alns[tuples[j * p + k]];
        cap64 = std::max(cap64, a);
    }
    if (cap64 > 0x7FFFFFF0ull) { set_error("lctp_solve_stage: candidate capacity overflow"); return LCTP_E_CAPACITY; }
    const uint32_t cap = (uint32_t)cap64;

    StageParams P;
    P.kind = st->kind; P.attempts = st->attempts; P.best_start = st->best_start; P.sample_size = (uint32_t)st->sample_size;
    P.plato_size = st->plato_size; P.anneal_steps = st->anneal_steps;
    P.max_iter = std::max<uint64_t>(100000, st->plato_size * 100);
    P.ln_init_prob = st->kind == 1 ? std::log(st->init_prob) : 0.0;
    P.n_workers = (uint32_t)n_workers; P.cap = cap; P.Wmax = 2 + p * h->max_n_windows;
    const bool want_counts = counts != nullptr && counts_off != nullptr;
    P.want_counts = want_counts ? 1 : 0;
    P.slab_bytes = slab_layout(cap, L.R, nullptr, nullptr);

    int rc;
    if ((rc = ensure_rng_mats(ctx))) return rc;
    if ((rc = ctx->d_worker_ixs.ensure(n))) return rc;
    if ((rc = ctx->d_worker_off.ensure(n_workers + 1))) return rc;
    if ((rc = ctx->d_rng.ensure(n_workers * 4))) return rc;
    if ((rc = ctx->d_tuples.ensure(n * p))) return rc;
    if ((rc = ctx->d_lik_mean.ensure(n))) return rc;
    if ((rc = ctx->d_lik_var.ensure(n))) return rc;
    if ((rc = ctx->d_liks.ensure(n * st->attempts))) return rc;
    if ((rc = ctx->d_nalns.ensure(n))) return rc;
    if ((rc = ctx->d_iters.ensure(n))) return rc;
    if ((rc = ctx->d_flags.ensure(2))) return rc;
    if (want_counts) { if ((rc = ctx->d_counts.ensure(n * (size_t)cap))) return rc; }

    LCTP_CUDA_CHECK(cudaMemcpyAsync(ctx->d_worker_ixs.p, worker_ixs, n * 8, cudaMemcpyHostToDevice, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(ctx->d_worker_off.p, worker_off, (n_workers + 1) * 8, cudaMemcpyHostToDevice, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(ctx->d_rng.p, worker_rng, n_workers * 32, cudaMemcpyHostToDevice, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(ctx->d_tuples.p, tuples.data(), n * p * 4, cudaMemcpyHostToDevice, s));
    LCTP_CUDA_CHECK(cudaMemsetAsync(ctx->d_flags.p, 0, 2 * sizeof(int), s));

    const double t_launch = dbg_now();
    int gs = 32;
    if (const char *e = getenv("LCTP_GS")) gs = atoi(e) == 16 ? 16 : 32;   // tuning knob: lanes per worker
    rc = gs == 32 ? launch_stage_gs<32>(h, P, n_workers, want_counts) : launch_stage_gs<16>(h, P, n_workers, want_counts);
    if (rc) return rc;

    int flags[2] = {0, 0};
    std::vector<uint64_t> nal(n), its(n);
    LCTP_CUDA_CHECK(cudaMemcpyAsync(lik_mean, ctx->d_lik_mean.p, n * 8, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(lik_var, ctx->d_lik_var.p, n * 8, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(worker_rng, ctx->d_rng.p, n_workers * 32, cudaMemcpyDeviceToHost, s));
    if (liks) LCTP_CUDA_CHECK(cudaMemcpyAsync(liks, ctx->d_liks.p, n * st->attempts * 8, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(nal.data(), ctx->d_nalns.p, n * 8, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(its.data(), ctx->d_iters.p, n * 8, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(flags, ctx->d_flags.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
    if (flags[0]) {
        set_error("lctp_solve_stage: candidate slab overflow (cap=%u, or > 65535 candidates of non-trivial reads "
                  "in one genotype)", cap);
        return LCTP_E_CAPACITY;
    }
    if (getenv("LCTP_DEBUG_TIMES")) {      // investigation aid: kernel interval on a process-wide GPU time base
        static std::mutex m; static cudaEvent_t ref = nullptr;
        std::lock_guard<std::mutex> lk(m);
        if (!ref) { cudaEventCreate(&ref); cudaEventRecord(ref, s); cudaEventSynchronize(ref); }
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, ref, ctx->ev[2]); cudaEventElapsedTime(&b, ref, ctx->ev[3]);
        fprintf(stderr, "[lctp debug] ctx %p stage kernel %.3f .. %.3f ms; host before launch %.3f ms, launch -> results on host %.3f ms\n",
                (void *)ctx, a, b, (t_launch - t_enter) * 1e3, (dbg_now() - t_launch) * 1e3);
    }
    {
        float ms = 0.f;
        LCTP_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
        ctx->stats.stage_ms += ms;
        ctx->stats.stage_launches += 1;
        ctx->stats.stage_genotypes += n;
        ctx->stats.stage_attempts += n * st->attempts;
        for (size_t j = 0; j < n; j++) { ctx->stats.stage_iters += its[j]; ctx->stats.stage_alns += nal[j]; }
    }
    if (n_alns_out) std::copy(nal.begin(), nal.end(), n_alns_out);
    if (iters_out) std::copy(its.begin(), its.end(), iters_out);
    if (want_counts) {
        uint64_t off = 0;
        for (size_t j = 0; j < n; j++) {
            counts_off[j] = off;
            if (off + nal[j] > counts_cap) { set_error("lctp_solve_stage: counts buffer too small"); return LCTP_E_CAPACITY; }
            LCTP_CUDA_CHECK(cudaMemcpyAsync(counts + off, ctx->d_counts.p + j * (size_t)cap, nal[j] * 2,
                                            cudaMemcpyDeviceToHost, s));
            off += nal[j];
        }
        counts_off[n] = off;
        LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    return LCTP_OK;
}

}  // namespace lctp

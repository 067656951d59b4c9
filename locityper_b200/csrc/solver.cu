// solver.cu -- a5..a13: one solver stage on the device.
//
// One warp per logical worker (= one xoshiro256++ stream of the reference, src/solvers/solve.rs:1007-1018),
// one worker per 32-thread CTA.  A worker solves its genotypes back to back on that stream exactly like
// Worker::run (:1104-1145):
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
// Data placement per worker (round 2: sized so that the private state of every resident worker stays in L2):
//   shared memory : window records (weight, table row, depth, 5 products), candidate offset + current
//                   assignment of every read, the list of non-trivial reads;
//   private slab  : per candidate ONE 32-bit record = source byte (contig index << 4 | rank in that contig's run;
//                   0xFF = the "unmapped" option) | window 1 << 8 | window 2 << 20 (64-bit records when a genotype
//                   has more than 4096 windows), plus an 8 KB buffer of pre-generated draws;
//   shared by all workers of a locus (read-only, L2): the ln-probabilities themselves.  A candidate's ln_prob
//                   is cm_lnprob[cm_off[hap * R + read] + rank] (cm_lnprob[NPA + read] for the unmapped option):
//                   it is never copied per worker, so the hot private state is 4 bytes per candidate instead
//                   of round 1's 24 and the slabs of all resident workers fit the L2 (ncu, C2 shape: DRAM
//                   traffic per launch 8.5 GB -> 63 MB, L2 hit rate 56 % -> 95 %).
#include "common.cuh"

#include <cmath>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <chrono>
#include <unordered_map>

namespace lctp {

static constexpr int CTA_THREADS = 32;
static constexpr int MAX_SAMPLE = 11;          // Floyd branch of rand::seq::index::sample on the lane-parallel path
#ifndef LCTP_HEADS
#define LCTP_HEADS 2
#endif
#ifndef LCTP_RNG_C
#define LCTP_RNG_C 32
#endif
static constexpr int RNG_C = LCTP_RNG_C;                // stream outputs generated per lane per fill
static constexpr uint32_t RNG_FILL = 32u * RNG_C;       // draws per fill (per-worker buffer: 8 KB at C = 32)
static constexpr int N_JUMP_TABS = 33;                  // T^(C*l), l = 0..31 (stream start of lane l; l = 0 unused), T^FILL (refill)
static constexpr size_t JUMP_TAB_WORDS = 32 * 256 * 4;  // u64 words per table: [32 state bytes][256 values][4]
static constexpr uint32_t SRC_UNMAPPED = 0xFFu;
static constexpr uint32_t MAX_RUN = 15;                 // rank in a (read, contig) run must fit 4 bits (reference: <= 10)

// Optional per-(genotype, attempt) debug outputs of a stage (lctp_stage_debug); all pointers device pointers or null.
struct DbgOut {
    double *lik;          // [n*attempts][2]: aln_lik, depth_lik
    uint32_t *cnt;        // [n*attempts][2]: depth[UNMAPPED_WINDOW], depth[BOUNDARY_WINDOW]
    double *ww, *wl;      // [n*attempts][wmax]: window weight, window ln-prob at its depth
    uint32_t *wd;         // [n*attempts][wmax]: window depth
    uint32_t wmax;
};

struct StageParams {
    uint32_t kind, attempts, best_start, sample_size;
    uint64_t plato_size, anneal_steps, max_iter;
    double ln_init_prob;
    uint32_t n_workers, cap, Wmax, want_counts, narrow_w, nt_global;
    uint64_t slab_bytes, big_bytes;
};

// Candidate record as stored in the slab: source byte and the two windows in ONE word, so that a candidate is one
// load.  The raw word travels through the prefetch pipelines unchanged and is taken apart where it is used, so that
// no instruction depends on a load right after it was issued.
template <bool WIDE> struct RecWord;
template <> struct RecWord<false> {       // <= 4096 windows per genotype
    typedef uint32_t T;
    static __device__ __forceinline__ uint32_t src(T raw) { return raw & 0xFFu; }
    static __device__ __forceinline__ uint32_t w1(T raw) { return (raw >> 8) & 0xFFFu; }
    static __device__ __forceinline__ uint32_t w2(T raw) { return raw >> 20; }
    static __device__ __forceinline__ T make(uint32_t s, uint32_t a, uint32_t b) { return s | (a << 8) | (b << 20); }
};
template <> struct RecWord<true> {
    typedef uint64_t T;
    static __device__ __forceinline__ uint32_t src(T raw) { return (uint32_t)raw & 0xFFu; }
    static __device__ __forceinline__ uint32_t w1(T raw) { return (uint32_t)(raw >> 16) & 0xFFFFu; }
    static __device__ __forceinline__ uint32_t w2(T raw) { return (uint32_t)(raw >> 32) & 0xFFFFu; }
    static __device__ __forceinline__ T make(uint32_t s, uint32_t a, uint32_t b) { return (T)s | ((T)a << 16) | ((T)b << 32); }
};

template <bool WIDE>
struct Slab {
    typedef typename RecWord<WIDE>::T Rec;
    Rec *rec;              // [cap]  record of every candidate, reads in read order (the reference's `alns`, a5)
    uint16_t *off;         // [R+1]  slab mode (StageParams::nt_global): first candidate of every read
    uint2 *ntinfo;         // [R]    slab mode: per non-trivial read (first candidate | candidates << 16, read id)
    uint64_t *rng_buf;     // [RNG_FILL] pre-generated draws of the worker's stream
    uint64_t *rng_blk;     // [32*4] block-start generator states of the current fill
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline size_t slab_bytes_for(uint32_t cap, bool wide, uint32_t slab_reads) {
    return (size_t)RNG_FILL * 8 + 32 * 4 * 8 + align_up((size_t)cap * (wide ? 8 : 4), 128) +
           (slab_reads ? align_up(((size_t)slab_reads + 1) * 2, 128) + align_up((size_t)slab_reads * 8, 128) : 0);
}
template <bool WIDE>
__device__ __forceinline__ void slab_layout(uint32_t cap, unsigned char *base, Slab<WIDE> &s) {
    s.rng_buf = (uint64_t *)base;
    s.rng_blk = (uint64_t *)(base + (size_t)RNG_FILL * 8);
    s.rec = (typename Slab<WIDE>::Rec *)(base + (size_t)RNG_FILL * 8 + 32 * 4 * 8);
    s.off = (uint16_t *)(base + (size_t)RNG_FILL * 8 + 32 * 4 * 8 + align_up((size_t)cap * (WIDE ? 8 : 4), 128));
    s.ntinfo = nullptr;    // set by the kernel in slab mode (it follows `off`, whose length depends on R)
}

// ------------------------------------------------------------------ warp helpers ----------------

static constexpr unsigned FULL = 0xFFFFFFFFu;
// CTA = one warp.  The hot code takes the lane from WarpShared::lane / Xo::lane (read once per kernel through a
// volatile asm): left to itself the compiler re-reads SR_TID.X at every use (S2R, ~20 cycles; 3.6 % of the kernel's
// instructions in the ncu capture of the previous build).
__device__ __forceinline__ int lane_id() { return threadIdx.x; }
template <typename T> __device__ __forceinline__ T wshfl(T v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ unsigned wballot(bool p) { return __ballot_sync(FULL, p); }
__device__ __forceinline__ bool wany(bool p) { return __any_sync(FULL, p) != 0; }

// ------------------------------------------------------------------ RNG stream -------------------
//
// The reference consumes ONE sequential xoshiro256++ stream per worker.  Stepping that generator redundantly
// in every lane costs ~25 instructions per draw.  Instead the warp pre-generates the stream in bulk:
// xoshiro's state transition is linear over GF(2), so lane l jumps its copy of the state ahead by l*RNG_C
// steps and then generates RNG_C consecutive outputs of the SAME sequential stream into a per-worker buffer:
// 32 lanes produce 1,024 exact draws per fill.  A jump is a 256x256 bit-matrix product; it is evaluated
// with byte-indexed tables (entry [b][v] = M * (v << 8b), 32 lookups of 32 bytes, ~450 instructions) instead
// of round 1's bit-by-bit product (~2,800), which is what makes a fill of 1,024 draws (8 KB per worker,
// L2-resident) as cheap per draw as round 1's 8,192-draw fill (64 KB per worker, streamed through DRAM).
// Consumers read the buffer through a 128-entry shared-memory ring (cp.async), in stream order, so the
// draw sequence (including the data-dependent extra draw of biased bounded samples) is bit-identical to the
// sequential generator.  At the end of a worker the exact state at the consumed position is rebuilt from the
// owning lane's block-start state.

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

// state <- M * state over GF(2), M given as a byte-indexed table (see above)
__device__ __forceinline__ Gen gen_tab_apply(const Gen &g, const ulonglong2 *__restrict__ tab) {
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const uint64_t sw = w == 0 ? g.s0 : w == 1 ? g.s1 : w == 2 ? g.s2 : g.s3;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const uint32_t v = (uint32_t)(sw >> (8 * b)) & 0xFFu;
            const ulonglong2 *e = tab + ((size_t)((w * 8 + b) * 256) + v) * 2;
            const ulonglong2 lo = __ldg(e), hi = __ldg(e + 1);
            a0 ^= lo.x; a1 ^= lo.y; a2 ^= hi.x; a3 ^= hi.y;
        }
    }
    Gen o; o.s0 = a0; o.s1 = a1; o.s2 = a2; o.s3 = a3;
    return o;
}

struct Xo {                // the worker's stream as seen by the solver code
    // Consumers read the draws from a 128-entry ring in shared memory holding stream positions [base, base + 128)
    // of the current fill (position q at ring[q & 127]).  The ring is filled 64 draws at a time by cp.async
    // (global -> shared, no destination registers, so no scoreboard slot is held while the copy is in flight: with
    // the register window of the previous builds the compiler's slot sharing made consumers wait on the prefetch
    // that had just been issued -- 10 % of the kernel in the ncu captures).  The copy of the second half is issued
    // when the first half is entered, i.e. at least 32 draws before anything can read it.
    uint32_t pos;          // draws consumed from the current fill
    uint32_t base;         // stream position of ring half 0 or 1 that `pos` is in (multiple of 64)
    uint32_t pend;         // 1 = the copy of [base + 64, base + 128) may still be in flight
    uint32_t lane;
    uint64_t *buf;         // [RNG_FILL] per-worker buffer (global, L2-resident)
    uint64_t *blk;         // [32][4] block-start states of the current fill
    uint64_t *ring;        // [128] shared memory
    const ulonglong2 *tabs;// jump tables
};

// copy buf[half .. half + 64) into its ring slot: 32 lanes x 16 bytes (a half beyond the fill is skipped)
__device__ __forceinline__ void ring_issue(const Xo &x, uint32_t half) {
    if (half < RNG_FILL) {
        const uint64_t *src = x.buf + half + 2u * x.lane;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(x.ring + (half & 64u) + 2u * x.lane);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void ring_wait() {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();          // the halves are copied by all lanes together
}
__device__ __forceinline__ void stream_resync(Xo &x) {
    __syncwarp();
    x.base = x.pos & ~63u;
    ring_issue(x, x.base);
    ring_issue(x, x.base + 64u);
    ring_wait();
    x.pend = 0;
}

// Generate this lane's block of a fill from its block-start state (stored in blk); out of line so that the
// generator state lives in registers only here.  The out-of-line helpers take plain pointers, never the stream
// object: its address must not escape or it would be demoted to local memory.
__device__ __noinline__ void fill_block(uint64_t *buf, uint64_t *blk, const ulonglong2 *refill_tab) {
    const int lane = lane_id();
    uint64_t *b = blk + lane * 4;
    Gen gen; gen.s0 = b[0]; gen.s1 = b[1]; gen.s2 = b[2]; gen.s3 = b[3];
    if (refill_tab) {      // block start of the previous fill -> same block of the next fill (FILL steps ahead)
        gen = gen_tab_apply(gen, refill_tab);
        b[0] = gen.s0; b[1] = gen.s1; b[2] = gen.s2; b[3] = gen.s3;
    }
    ulonglong2 *o = (ulonglong2 *)(buf + lane * RNG_C);
#pragma unroll 4
    for (int q = 0; q < RNG_C / 2; q++) {
        ulonglong2 v;
        v.x = gen_next(gen);
        v.y = gen_next(gen);
        o[q] = v;      // plain store: the greedy pipeline reads the draws with L1-cached 4-byte cp.async
    }
    __syncwarp();
}
__device__ __forceinline__ void stream_refill(Xo &x) {
    if (x.pend) ring_wait();
    fill_block(x.buf, x.blk, x.tabs + (size_t)32 * (JUMP_TAB_WORDS / 2));
    x.pos = 0;
    stream_resync(x);
}

// Start a stream from the scalar state st[4]: lane l jumps ahead by l*RNG_C (binary decomposition of l).
__device__ __noinline__ void stream_begin_blocks(const uint64_t *__restrict__ st, uint64_t *buf, uint64_t *blk,
                                                 const ulonglong2 *tabs) {
    Gen g; g.s0 = st[0]; g.s1 = st[1]; g.s2 = st[2]; g.s3 = st[3];
    const int lane = lane_id();
    // one table per lane (T^(C * lane), 8 MB in all, L2-resident): a single application instead of up to five
    // dependent ones by the binary decomposition of the lane id (ncu: 3 % of the C2 stage kernel, one worker per genotype)
    if (lane) g = gen_tab_apply(g, tabs + (size_t)lane * (JUMP_TAB_WORDS / 2));
    __syncwarp();
    uint64_t *b = blk + lane * 4;
    b[0] = g.s0; b[1] = g.s1; b[2] = g.s2; b[3] = g.s3;
    fill_block(buf, blk, nullptr);
}
__device__ __forceinline__ void stream_begin(Xo &x, const uint64_t *__restrict__ st) {
    stream_begin_blocks(st, x.buf, x.blk, x.tabs);
    x.pos = 0;
    x.pend = 0;
    stream_resync(x);
}

// Exact scalar state after `pos` draws of the current fill (what the sequential generator would hold).
__device__ __noinline__ void stream_end_state(uint32_t pos, const uint64_t *blk, uint64_t *__restrict__ out) {
    __syncwarp();
    Gen g;
    if (pos < RNG_FILL) {
        const int owner = pos / RNG_C;
        const uint64_t *b = blk + owner * 4;
        g.s0 = b[0]; g.s1 = b[1]; g.s2 = b[2]; g.s3 = b[3];
        const int steps = pos - owner * RNG_C;
        for (int q = 0; q < steps; q++) gen_next(g);
    } else {                                              // whole fill consumed: end of the last block
        const uint64_t *b = blk + 31 * 4;
        g.s0 = b[0]; g.s1 = b[1]; g.s2 = b[2]; g.s3 = b[3];
        for (int q = 0; q < RNG_C; q++) gen_next(g);
    }
    if (lane_id() == 0) { out[0] = g.s0; out[1] = g.s1; out[2] = g.s2; out[3] = g.s3; }
}
__device__ __forceinline__ void stream_end(Xo &x, uint64_t *__restrict__ out) {
    if (x.pend) { ring_wait(); x.pend = 0; }
    stream_end_state(x.pos, x.blk, out);
}

// Make the ring cover stream positions [pos, pos + count), count <= 32; false = the fill ends first (the caller
// then uses the one-draw-at-a-time path, which refills).  `pos` may have been moved arbitrarily (forwards by bulk
// consumers, backwards by stream_unconsume) since the last call.
__device__ __forceinline__ bool stream_cover(Xo &x, uint32_t count) {
    if (x.pos + count > RNG_FILL) return false;
    const uint32_t d = x.pos - x.base;
    if (d >= 64u) {
        if (d < 128u) {                                   // entered the second half: it becomes the first one
            if (x.pend) ring_wait();                      // (issued >= 64 draws ago: done long since)
            else __syncwarp();                            // every lane is done reading the half that is overwritten
            x.base += 64u;
            ring_issue(x, x.base + 64u);
            x.pend = 1;
        } else {                                          // far jump (or backwards)
            if (x.pend) ring_wait();
            stream_resync(x);
        }
    }
    if (x.pend && x.pos + count > x.base + 64u) { ring_wait(); x.pend = 0; }
    return true;
}
// After stream_cover(count): the (pos + rank)-th draw of the stream, for any per-lane rank < count.
__device__ __forceinline__ uint32_t stream_peek_hi(const Xo &x, uint32_t rank) {
    return ((const uint32_t *)x.ring)[2u * ((x.pos + rank) & 127u) + 1u];
}
__device__ __forceinline__ uint64_t stream_peek64(const Xo &x, uint32_t rank) {
    return x.ring[(x.pos + rank) & 127u];
}
// Give back the last n draws (all taken from the current fill without a refill in between).
__device__ __forceinline__ void stream_unconsume(Xo &x, uint32_t n) { x.pos -= n; }

__device__ __forceinline__ void stream_advance(Xo &x) {
    if (x.pos == RNG_FILL) stream_refill(x);
    stream_cover(x, 1);
}
__device__ __forceinline__ uint32_t xo_u32(Xo &x) {   // next_u32 = upper half of next_u64
    stream_advance(x);
    const uint32_t r = stream_peek_hi(x, 0);
    x.pos++;
    return r;
}
__device__ __forceinline__ uint64_t xo_next(Xo &x) {
    stream_advance(x);
    const uint64_t r = stream_peek64(x, 0);
    x.pos++;
    return r;
}
// rand UniformInt::sample_single_inclusive with a u32 sample type: value in [0, range), range != 0.
__device__ __forceinline__ uint32_t xo_below(Xo &x, uint32_t range) {
    const uint64_t m = (uint64_t)xo_u32(x) * (uint64_t)range;
    uint32_t res = (uint32_t)(m >> 32);
    const uint32_t lo = (uint32_t)m;
    if (lo > 0u - range) {   // biased zone: one extra draw (warp-uniform branch)
        const uint32_t nh = (uint32_t)(((uint64_t)xo_u32(x) * (uint64_t)range) >> 32);
        res += (lo + nh < lo) ? 1u : 0u;
    }
    return res;
}
// rand StandardUniform f64 (src/solvers/stoch.rs:216)
__device__ __forceinline__ double u64_to_unit_f64(uint64_t v) {
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double xo_f64(Xo &x) { return u64_to_unit_f64(xo_next(x)); }
// Lane-parallel bounded draws: lanes < count each want random_range(0..range_l) and lane k must receive the
// k-th draw of the stream.  Succeeds (and consumes `count` draws) only when no lane lands in the biased zone
// -- which would consume an extra draw and shift every later lane -- otherwise nothing is consumed and the
// caller runs the sequential path.  P(fallback) ~ count * range / 2^32.
__device__ __forceinline__ bool xo_below_lanes(Xo &x, uint32_t count, uint32_t my_range, uint32_t &res) {
    if (!stream_cover(x, count)) return false;
    const uint64_t m = (uint64_t)stream_peek_hi(x, x.lane) * (uint64_t)my_range;
    const bool biased = x.lane < count && (uint32_t)m > 0u - my_range;
    if (wany(biased)) return false;
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

// Per-window state in shared memory, 56 bytes per window: the window's weight, its depth-table row, the
// current read depth d and the five products weight * table[row][d-2 .. d+2] (rounded once, exactly as
// WindowDistr::ln_prob computes them, src/model/distr_cache.rs:34-39).  Every likelihood delta of the
// solver loops is then a difference of two shared-memory values: no global loads.
// TRIVIAL windows (WindowDistr::TRIVIAL, distr_cache.rs:27-30) are stored as weight 0 on the all-zero
// extra row of the device table, so 0*0 - 0*0 = +0.0 reproduces the reference's literal 0.0 with no
// branch; a zero depth change gives p - p = +0.0 the same way.
// Layout: structure of arrays (five product planes, then weight, row, depth), so that lanes looking up
// DIFFERENT windows hit different banks.
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
    // Two placements of the per-read index, chosen by the host (StageParams::nt_global = "slab mode"):
    //   shared memory (R small enough for 16 workers per SM): off[R+1] + nt_read[n_nt];
    //   the worker's slab (large R: what remains in shared memory is the window state and `assgn`, so twice as
    //   many workers fit): off[R+1] + ntinfo[n_nt] = everything the greedy loop needs about a sampled read in ONE
    //   8-byte load, which the loop issues a round ahead.
    uint16_t *off;         // [R+1]  first candidate of every read (+ end sentinel); generic pointer
    uint16_t *nt_read;     // [R]    shared-memory mode: read ids of the non-trivial reads, ascending
    uint2 *ntinfo;         // [R]    slab mode: (first candidate | candidates << 16, read id) per non-trivial read
    uint8_t *assgn;        // [R]    current assignment (candidate rank) of every read
    uint32_t *unm_bits;    // [ceil(R/32)] bit r%32 of word r/32: read r has the "unmapped" option among its candidates
    uint32_t *haps;        // [LCTP_MAX_PLOIDY] haplotypes of the genotype, then [LCTP_MAX_PLOIDY+1] window shifts
    uint32_t *samp;        // [12] scratch of sample_resolve; [12..16) spare
    double *lik;           // [0] aln_lik, [1] depth_lik of the assignment being solved (lane 0 updates them: two fewer
                           //     64-bit values live in the solver loops), [2] = iterations of the genotype as u64
    uint32_t steps_sa;     // annealing: shared-memory address of the 32 prepared step records (48 bytes each)
    uint32_t p2_sa;        // shared-memory address of win.p(0, 2)
    uint32_t zero_row;     // offset of the all-zero row
    uint32_t depth_k;
    uint32_t lane;
};
static constexpr size_t STEP_REC_BYTES = 48;
__host__ __device__ inline size_t group_smem_bytes(uint32_t Wmax, uint32_t R, bool nt_global, bool anneal) {
    return (anneal ? 32 * STEP_REC_BYTES : 0) + align_up((size_t)win_stride(Wmax) * 56, 16) +
           (nt_global ? 0 : align_up(((size_t)R + 1) * 2, 16) + align_up((size_t)R * 2, 16)) +
           align_up((size_t)R, 16) + align_up((size_t)((R + 31) / 32) * 4, 16) +
           align_up((size_t)(2 * LCTP_MAX_PLOIDY + 1) * 4, 16) + 64 /* samp */ + 32 /* lik */ + 1024 /* draw ring */;
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
__device__ __forceinline__ double depth_lik_diff(const WarpShared &ws, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4) {
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
template <bool WIDE>
__device__ __forceinline__ double depth_lik_diff_raw(const WarpShared &ws, typename RecWord<WIDE>::T raw_old,
                                                     typename RecWord<WIDE>::T raw_new) {
    typedef RecWord<WIDE> RW;
    return depth_lik_diff(ws, RW::w1(raw_old), RW::w2(raw_old), RW::w1(raw_new), RW::w2(raw_new));
}

// ------------------------------------------------------------------ a5: instance build ----------

struct Instance {
    uint32_t h0, h1;        // first two haplotypes of the genotype (registers); all of them are in ws.haps
    uint32_t W, A, n_nt;
};
__device__ __forceinline__ uint32_t inst_hap(const Instance &I, const WarpShared &ws, uint32_t k) {
    return k == 0 ? I.h0 : k == 1 ? I.h1 : ws.haps[k];
}
__device__ __forceinline__ uint32_t inst_wshift(const WarpShared &ws, uint32_t k) { return ws.haps[LCTP_MAX_PLOIDY + k]; }

// (first candidate | candidate count << 16, read id) of the idx-th non-trivial read
__device__ __forceinline__ uint2 nt_info(const WarpShared &ws, uint32_t idx) {
    if (ws.ntinfo) return ws.ntinfo[idx];
    const uint32_t r = ws.nt_read[idx];
    const uint32_t o = ws.off[r];
    return make_uint2(o | (((uint32_t)ws.off[r + 1] - o) << 16), r);
}

// Index into cm_lnprob of the ln-probability of the candidate with source byte s of read r (the reference's
// ReadGtAlns::ln_prob).  b0 / b1: cm_off of the read on the first two haplotypes (the p <= 2 fast path has them).
__device__ __forceinline__ uint32_t lp_index2(const LocusDev &L, uint32_t r, uint32_t s, uint32_t b0, uint32_t b1) {
    return s == SRC_UNMAPPED ? L.npa + r : ((s >> 4) ? b1 : b0) + (s & 15u);
}
__device__ __forceinline__ uint32_t lp_index(const LocusDev &L, const Instance &I, const WarpShared &ws, uint32_t r,
                                             uint32_t s) {
    if (s == SRC_UNMAPPED) return L.npa + r;
    return __ldg(L.cm_off + (size_t)inst_hap(I, ws, s >> 4) * L.R + r) + (s & 15u);
}

// GenotypeAlignments::new: per read, gather candidates of every genotype contig above the running
// threshold, append the unmapped option, stable-sort descending (here: a p-way merge of the already
// sorted per-contig lists, ties resolved in insertion order = contig order, unmapped last), cut at the
// final threshold.  Returns false on overflow (more than min(cap, 65535) candidates).
// HEADS > 0 (ploidy <= 2): the first HEADS ln-probs of each contig's run are fetched up front with
// independent loads and the threshold / cut / merge run on registers; runs longer than HEADS fall back to
// memory for the tail.  HEADS == 0: everything from memory (any ploidy).
template <int HEADS, bool WIDE>
__device__ bool build_instance(const LocusDev &L, const Slab<WIDE> &S, const WarpShared &ws, uint32_t cap, Instance &I) {
    constexpr int PK = HEADS > 0 ? 2 : LCTP_MAX_PLOIDY;
    constexpr int HN = HEADS > 0 ? HEADS : 1;
    const uint32_t R = L.R, p = L.p;
    const int lane = (int)ws.lane;
    const uint32_t lim = min(cap, 65535u);
    uint32_t base = 0, nt_base = 0;
    bool ok = true;
    for (uint32_t r0 = 0; r0 < R; r0 += 32) {
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
                    const size_t key = (size_t)inst_hap(I, ws, k) * R + r;
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
        uint32_t incl = nw;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += o;
        }
        const uint32_t total = wshfl(incl, 31);
        const uint32_t start = base + incl - nw;
        const unsigned ntmask = wballot(nt);
        const unsigned unmmask = wballot(with_unm);
        if (base + total > lim) ok = false;
        if (ok && valid) {
            ws.off[r] = (uint16_t)start;
            ws.assgn[r] = 0;
            if (nt) {
                const uint32_t pos = nt_base + __popc(ntmask & ((1u << ws.lane) - 1u));
                if (ws.ntinfo) ws.ntinfo[pos] = make_uint2(start | (nw << 16), r);
                else ws.nt_read[pos] = (uint16_t)r;
            }
            bool unm_left = with_unm;
            const long long unm_key = total_key(unm);
            for (uint32_t t = 0; t < nw; t++) {
                int bk = -1;
                long long bkey = 0;
#pragma unroll
                for (int k = 0; k < PK; k++) {
                    if ((uint32_t)k < p && lb[k] < le[k]) {
                        const long long key = total_key(val(k, lb[k]));
                        if (bk < 0 || key > bkey) { bk = k; bkey = key; }
                    }
                }
                uint32_t src;
                if (bk >= 0 && !(unm_left && unm_key > bkey)) {
                    uint32_t ix = 0, first = 0;
#pragma unroll
                    for (int k = 0; k < PK; k++) if (k == bk) { ix = lb[k]; first = l0[k]; lb[k]++; }
                    src = ((uint32_t)bk << 4) | (ix - first);
                } else {
                    src = SRC_UNMAPPED;
                    unm_left = false;
                }
                S.rec[start + t] = RecWord<WIDE>::make(src, 0u, 0u);
            }
        }
        if (lane == 0) ws.unm_bits[r0 >> 5] = unmmask;
        base += total;
        nt_base += __popc(ntmask);
    }
    if (lane == 0 && ok) ws.off[R] = (uint16_t)base;
    I.A = base;
    I.n_nt = nt_base;
    __syncwarp();
    return ok;
}

// ------------------------------------------------------------------ a6: apply_tweak -------------

// ContigInfo::get_shifted_window_ix (src/model/windows.rs:62-68,465-470)
__device__ __forceinline__ uint32_t shifted_window(const LocusDev &L, uint32_t reg_start, uint32_t reg_end, uint32_t shift,
                                                   uint32_t middle) {
    if (reg_start <= middle && middle < reg_end) return (L.window == 1u ? middle - reg_start : (uint32_t)__umul64hi(middle - reg_start, L.c_window)) + shift;   // (middle - reg_start) / L.window
    return 1;   // BOUNDARY_WINDOW
}

template <bool WIDE>
__device__ void apply_tweak(const LocusDev &L, const Slab<WIDE> &S, const Instance &I, const WarpShared &ws, Xo &rng) {
    typedef RecWord<WIDE> RW;
    typedef typename RW::T Rec;
    const int lane = (int)ws.lane;
    const uint32_t tweak = L.tweak, R = L.R, p = L.p;
    const uint32_t span = 2 * tweak + 1;
    // (i) read middles: one next_u64 per candidate that has a parent, in candidate order (read-major).
    // Lanes over CANDIDATES (the records of a genotype are contiguous in read order): candidate i takes draw number
    // i - #(unmapped options before i), its read is found from the offsets of the <= 32 reads that a round of 32
    // candidates touches (one coalesced load + one REDUX), records are read and written coalesced.  A round that
    // straddles the end of the draw buffer is processed in two parts with the refill in between.  (Lanes over reads,
    // four candidates at a time, cost 8x the instructions: 11 % of the C2 stage kernel, profiles/r02_summary.md.)
    {
        const uint32_t A = I.A;
        uint32_t r_lo = 0;                       // read of candidate i0
        uint32_t cum_unm = 0;                    // unmapped options among the candidates before i0
        int fill_q0 = -(int)rng.pos;             // draw number that sits at position 0 of the current fill
        for (uint32_t i0 = 0; i0 < A; i0 += 32) {
            const uint32_t i = i0 + lane;
            const bool valid = i < A;
            // ends of the reads r_lo, r_lo + 1, ...: a read that ends inside (i0, i0 + 32) starts the next one there
            const uint32_t er = r_lo + 1u + lane;
            const uint32_t e = er <= R ? (uint32_t)ws.off[er] : 0xFFFFFFFFu;
            const unsigned starts = __reduce_or_sync(FULL, (e > i0 && e < i0 + 32u) ? (1u << (e - i0)) : 0u);
            const uint32_t r = r_lo + (uint32_t)__popc(starts & (0xFFFFFFFFu >> (31u - lane)));
            r_lo += (uint32_t)__popc(starts) + (wany(e == i0 + 32u) ? 1u : 0u);
            const Rec rc = valid ? S.rec[i] : RW::make(SRC_UNMAPPED, 0u, 0u);
            const uint32_t sb = RW::src(rc);
            const bool par = valid && sb != SRC_UNMAPPED;
            const unsigned unm = wballot(valid && sb == SRC_UNMAPPED);
            const int q = (int)(i - cum_unm - (uint32_t)__popc(unm & ((1u << lane) - 1u)));     // draw number of this candidate
            cum_unm += (uint32_t)__popc(unm);
            uint2 mid = make_uint2(0u, 0u);
            uint32_t hap = 0;
            if (par) {
                const uint32_t k = sb >> 4;
                hap = inst_hap(I, ws, k);
                mid = L.cm_mid[L.cm_off[(size_t)hap * R + r] + (sb & 15u)];
            }
            if (valid && !par) S.rec[i] = RW::make(SRC_UNMAPPED, 0u, 0u);       // [UNMAPPED_WINDOW; 2] (windows.rs:99-105)
            bool todo = par;
            for (;;) {
                const int rel = q - fill_q0;                                    // position of the draw in the current fill
                const bool now = todo && (tweak == 0 || rel < (int)RNG_FILL);
                if (now) {
                    uint32_t t1 = 0, t2 = 0;
                    if (tweak) {
                        // x % span = hi64((c * x mod 2^64) * span), c = ceil(2^64 / span) (exact for 32-bit x and span)
                        const uint64_t draw = __ldcg(rng.buf + rel);
                        t1 = (uint32_t)__umul64hi(L.c_span * (uint64_t)(uint32_t)(draw >> 32), span);
                        t2 = (uint32_t)__umul64hi(L.c_span * (uint64_t)(uint32_t)draw, span);
                    }
                    const uint32_t shift = inst_wshift(ws, sb >> 4);
                    const uint32_t reg_start = L.hap_reg_start[hap];
                    const uint32_t reg_end = reg_start + L.hap_n_windows[hap] * L.window;
                    const uint32_t w1 = mid.x == LCTP_NONE_U32 ? 0u : shifted_window(L, reg_start, reg_end, shift, mid.x + t1);
                    const uint32_t w2 = mid.y == LCTP_NONE_U32 ? 0u : shifted_window(L, reg_start, reg_end, shift, mid.y + t2);
                    S.rec[i] = RW::make(sb, w1, w2);
                    todo = false;
                }
                if (!wany(todo)) break;
                __syncwarp();                    // every lane has read its draws of this fill before it is replaced
                stream_refill(rng);
                fill_q0 += (int)RNG_FILL;
            }
        }
        if (tweak) rng.pos = (uint32_t)((int)(A - cum_unm) - fill_q0);
    }
    // (ii) window distributions: one bounded i32 draw per window, contigs in genotype order
    if (lane < 2) { ws.win.weight(lane) = 0.0; ws.win.row(lane) = ws.zero_row; ws.win.depth(lane) = 0; }
    for (uint32_t k = 0; k < p; k++) {
        const uint32_t hap = inst_hap(I, ws, k);
        const uint32_t nwin = L.hap_n_windows[hap];
        const uint32_t reg_start = L.hap_reg_start[hap], hlen = L.hap_len[hap];
        const uint64_t pos_off = L.hap_pos_off[hap];
        const uint32_t wsh = inst_wshift(ws, k);
        for (uint32_t i0 = 0; i0 < nwin; i0 += 32) {
            const uint32_t cnt = min(32u, nwin - i0);
            uint32_t my_wstart = 0;
            {
                // generate_windows (windows.rs:478-486): random_range(-left..=right), one per window
                const uint32_t start = reg_start + (i0 + min((uint32_t)lane, cnt - 1u)) * L.window;
                const uint32_t end = start + L.window;
                const uint32_t left = min(tweak, start), right = min(tweak, hlen - end);
                uint32_t rr = 0;
                if (xo_below_lanes(rng, cnt, left + right + 1u, rr)) my_wstart = start + rr - left;
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
                const uint32_t w = wsh + i0 + lane;
                const bool trivial = weight < L.min_weight || weight < 1e-7;
                ws.win.weight(w) = trivial ? 0.0 : weight;
                ws.win.row(w) = trivial ? ws.zero_row : gc * L.depth_k;
                ws.win.depth(w) = 0;
            }
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------ a8: ReadAssignment::new -----

// Sequentially (in index order) add `count` per-lane terms to acc: reproduces iter().sum() order.
__device__ __forceinline__ void seq_add(double &acc, double term, int count) {
    for (int q = 0; q < count; q++) acc = __dadd_rn(acc, wshfl(term, q));
}

// init_mode 0: every read at candidate 0; 1: random_range(0..m) per non-trivial read (read order).
template <bool WIDE>
__device__ void init_assignment(const LocusDev &L, const Slab<WIDE> &S, const Instance &I, const WarpShared &ws,
                                Xo &rng, int init_mode) {
    typedef RecWord<WIDE> RW;
    const int lane = (int)ws.lane;
    const uint32_t R = L.R;
    for (uint32_t w = lane; w < I.W; w += 32) ws.win.depth(w) = 0;
    // assignments: trivial reads stay at 0
    for (uint32_t r = lane; r < R; r += 32) ws.assgn[r] = 0;
    __syncwarp();
    if (init_mode == 1) {
        for (uint32_t i0 = 0; i0 < I.n_nt; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint32_t r = 0, my_n = 1;
            if (i < I.n_nt) { const uint2 inf = nt_info(ws, i); r = inf.y; my_n = inf.x >> 16; }
            uint32_t a = 0;
            const uint32_t cnt = min(32u, I.n_nt - i0);
            if (!xo_below_lanes(rng, cnt, my_n, a)) {
                for (uint32_t q = 0; q < cnt; q++) {
                    const uint32_t m = wshfl(my_n, (int)q);
                    const uint32_t v = xo_below(rng, m);
                    if ((uint32_t)lane == q) a = v;
                }
            }
            if (i < I.n_nt) ws.assgn[r] = (uint8_t)a;
        }
        __syncwarp();
    }
    // depth counts + aln_lik in read order (src/model/assgn.rs:205-217,351-353)
    double al = 0.0;
    for (uint32_t r0 = 0; r0 < R; r0 += 32) {
        const uint32_t r = r0 + lane;
        double term = 0.0;
        if (r < R) {
            const uint32_t c = (uint32_t)ws.off[r] + ws.assgn[r];
            const typename RW::T raw = S.rec[c];
            term = __ldg(L.cm_lnprob + lp_index(L, I, ws, r, RW::src(raw)));
            atomicAdd(&ws.win.depth(RW::w1(raw)), 1u);
            atomicAdd(&ws.win.depth(RW::w2(raw)), 1u);
        }
        seq_add(al, term, (int)min(32u, R - r0));
    }
    __syncwarp();
    for (uint32_t q = lane; q < I.W * 5u; q += 32) win_refresh(ws, L.depth_table, q / 5u, (int)(q % 5u));
    __syncwarp();
    // depth_lik in window order (assgn.rs:347-350)
    double dl = 0.0;
    for (uint32_t w0 = 0; w0 < I.W; w0 += 32) {
        const uint32_t w = w0 + lane;
        const double term = w < I.W ? ws.win.p(w, 2) : 0.0;
        seq_add(dl, term, (int)min(32u, I.W - w0));
    }
    if (lane == 0) { ws.lik[0] = al; ws.lik[1] = dl; }
    __syncwarp();
}

// ------------------------------------------------------------------ a9: targets -----------------

template <bool WIDE>
struct Move { double dld, dlp; typename RecWord<WIDE>::T raw_old, raw_new; };

// calculate_improvement (src/model/assgn.rs:321-328) for moving read r (candidates at o..) from rank a to new_a
template <bool WIDE>
__device__ __forceinline__ double calc_improvement(const LocusDev &L, const Slab<WIDE> &S, const Instance &I,
                                                   const WarpShared &ws, uint32_t r, uint32_t o, uint32_t a,
                                                   uint32_t new_a, Move<WIDE> &mv) {
    typedef RecWord<WIDE> RW;
    mv.raw_old = S.rec[o + a];
    mv.raw_new = S.rec[o + new_a];
    const uint32_t so = RW::src(mv.raw_old), sn = RW::src(mv.raw_new);
    uint32_t io, in;
    if (L.p <= 2) {
        const uint32_t b0 = __ldg(L.cm_off + (size_t)I.h0 * L.R + r);
        const uint32_t b1 = L.p > 1 ? __ldg(L.cm_off + (size_t)I.h1 * L.R + r) : 0u;
        io = lp_index2(L, r, so, b0, b1);
        in = lp_index2(L, r, sn, b0, b1);
    } else {
        io = lp_index(L, I, ws, r, so);
        in = lp_index(L, I, ws, r, sn);
    }
    const double lpo = __ldg(L.cm_lnprob + io), lpn = __ldg(L.cm_lnprob + in);
    mv.dld = depth_lik_diff_raw<WIDE>(ws, mv.raw_old, mv.raw_new);
    mv.dlp = __dsub_rn(lpn, lpo);
    return __dadd_rn(__dmul_rn(L.depth_contrib, mv.dld), __dmul_rn(L.aln_contrib, mv.dlp));
}

// reassign (src/model/assgn.rs:331-343); warp-uniform inputs, lane 0 writes
template <bool WIDE>
__device__ __forceinline__ void apply_move(const WarpShared &ws, const double *__restrict__ table, uint32_t r,
                                           uint32_t new_a, const Move<WIDE> &mv) {
    typedef RecWord<WIDE> RW;
    const uint32_t w1 = RW::w1(mv.raw_old), w2 = RW::w2(mv.raw_old), w3 = RW::w1(mv.raw_new), w4 = RW::w2(mv.raw_new);
    __syncwarp();          // every lane is done reading the assignments / window state this move overwrites
    if (ws.lane == 0) {
        ws.win.depth(w3) += 1;
        ws.win.depth(w4) += 1;
        ws.win.depth(w1) -= 1;
        ws.win.depth(w2) -= 1;
        ws.assgn[r] = (uint8_t)new_a;
        ws.lik[1] = __dadd_rn(ws.lik[1], mv.dld);       // depth_lik += ..., aln_lik += ... (assgn.rs:336-337)
        ws.lik[0] = __dadd_rn(ws.lik[0], mv.dlp);
    }
    __syncwarp();
    // slide the product slices of the (up to four) windows whose depth changed
    for (int q = (int)ws.lane; q < 20; q += 32) {
        const int j = q / 5;
        const uint32_t w = j == 0 ? w1 : j == 1 ? w2 : j == 2 ? w3 : w4;
        win_refresh(ws, table, w, q % 5);
    }
    __syncwarp();
}

// One random reassignment target, warp-uniform: ReassignmentTarget::random (src/model/assgn.rs:451-471)
struct Target { uint32_t r, o, a, new_a; };
__device__ __forceinline__ Target random_target(const WarpShared &ws, const Instance &I, Xo &rng) {
    Target t;
    const uint32_t idx = xo_below(rng, I.n_nt);              // random_range(0..n_nontrivial), usize via the u32 path
    const uint2 inf = nt_info(ws, idx);
    t.r = inf.y;
    t.o = inf.x & 0xFFFFu;
    const uint32_t n = inf.x >> 16;
    t.a = ws.assgn[t.r];
    if (n == 2) t.new_a = 1u - t.a;
    else {
        const uint32_t i = 1u + xo_below(rng, n - 1u);       // random_range(1..n as u16)
        t.new_a = i <= t.a ? i - 1u : i;
    }
    return t;
}

// Speculative chain of random targets.  The steps of max_abs_random and of the annealing loops each consume a
// data-dependent number of draws (one for the read, one more when the read has more than two candidates, and in
// the annealing phase the U(0,1) of a rejected step), so step j+1 starts where step j ended.  Every lane l
// computes the step that WOULD start at stream position pos + l; the lanes actually on the chain 0 -> next(0)
// -> ... are found by pointer doubling, and all of them evaluate their target against the CURRENT state in
// parallel.  That is exact as long as no earlier step of the chain changes the state, i.e. up to and including
// the first accepted step -- the caller commits exactly that prefix and leaves the rest of the draws in the
// stream.  A step that would need a bias-correction draw, or that does not fit in the register window, ends
// the chain; the caller then takes one step on the sequential path (which refills / handles the bias).
struct Spec {
    bool usable;                 // this lane holds step `rank` of the chain
    uint32_t rank, count;        // count = usable steps (a prefix of the chain)
    uint32_t r, o, a, new_a;     // the step's target
    uint32_t nd;                 // draws of the target itself (1 or 2)
    uint64_t udraw;              // the draw after them
    unsigned umask;              // lanes holding usable steps
};
template <bool WITH_U>
__device__ __forceinline__ void spec_targets(const WarpShared &ws, const Instance &I, Xo &rng, Spec &sp) {
    const uint32_t lane = ws.lane;
    sp.usable = false; sp.rank = 0; sp.count = 0; sp.umask = 0u;
    sp.r = sp.o = sp.a = sp.new_a = 0; sp.nd = 1; sp.udraw = 0;
    if (rng.pos >= RNG_FILL) return;
    const uint32_t avail = min(32u, RNG_FILL - rng.pos);
    stream_cover(rng, avail);              // every position a lane peeks at is in the ring
    const uint64_t d = stream_peek64(rng, lane);
    const uint64_t d1 = __shfl_down_sync(FULL, d, 1), d2 = __shfl_down_sync(FULL, d, 2);
    const uint64_t m = (d >> 32) * (uint64_t)I.n_nt;
    const bool bias1 = (uint32_t)m > 0u - I.n_nt;
    const uint32_t idx = (uint32_t)(m >> 32);
    const uint2 inf = nt_info(ws, idx);
    sp.r = inf.y;
    sp.o = inf.x & 0xFFFFu;
    const uint32_t n = inf.x >> 16;
    sp.a = ws.assgn[sp.r];
    const bool two = n > 2;
    const uint64_t m2 = (d1 >> 32) * (uint64_t)(n - 1u);
    const bool bias2 = two && (uint32_t)m2 > 0u - (n - 1u);
    const uint32_t i2 = 1u + (uint32_t)(m2 >> 32);
    sp.new_a = two ? (i2 <= sp.a ? i2 - 1u : i2) : 1u - sp.a;
    sp.nd = two ? 2u : 1u;
    sp.udraw = two ? d2 : d1;
    const uint32_t len = sp.nd + (WITH_U ? 1u : 0u);
    const bool ok = lane + len <= avail && !bias1 && !bias2;
    // chain membership by pointer doubling: S_{t+1} = S_t u jump_t(S_t), jump_{t+1} = jump_t o jump_t
    uint32_t jump = ok ? lane + len : 32u;
    unsigned chain = 1u;
#pragma unroll
    for (int t = 0; t < 5; t++) {
        const unsigned contrib = (((chain >> lane) & 1u) && jump < 32u) ? (1u << jump) : 0u;
        chain |= __reduce_or_sync(FULL, contrib);
        const uint32_t j2 = wshfl(jump, (int)(jump & 31u));
        jump = jump < 32u ? j2 : 32u;
    }
    const bool on = (chain >> lane) & 1u;
    const unsigned bad = wballot(on && !ok);
    const uint32_t first_bad = bad ? (uint32_t)__ffs(bad) - 1u : 32u;
    sp.usable = on && lane < first_bad;
    sp.umask = wballot(sp.usable);
    sp.rank = __popc(sp.umask & ((1u << ws.lane) - 1u));
    sp.count = __popc(sp.umask);
}
// stream position (relative to rng.pos) right after the first `e` steps of the chain, e <= count
__device__ __forceinline__ uint32_t spec_offset_after(const Spec &sp, uint32_t lane, uint32_t e, uint32_t len_mine) {
    if (e == 0) return 0u;
    const unsigned sel = wballot(sp.usable && sp.rank == e - 1u);
    const int src = __ffs(sel) - 1;
    return wshfl(lane + len_mine, src);
}

// max_abs_random (src/solvers/stoch.rs:19-22) with INIT_ITER = 100: the state does not change, so whole
// chains are consumed.
template <bool WIDE>
__device__ double max_abs_random(const LocusDev &L, const Slab<WIDE> &S, const Instance &I, const WarpShared &ws, Xo &rng) {
    double acc = 0.0;
    uint32_t left = 100;
    while (left > 0) {
        Spec sp;
        spec_targets<false>(ws, I, rng, sp);
        if (sp.count == 0) {
            const Target t = random_target(ws, I, rng);
            Move<WIDE> mv;
            acc = fmax(acc, fabs(calc_improvement<WIDE>(L, S, I, ws, t.r, t.o, t.a, t.new_a, mv)));
            left--;
            continue;
        }
        const uint32_t e = min(sp.count, left);
        double v = 0.0;
        if (sp.usable && sp.rank < e) {
            Move<WIDE> mv;
            v = fabs(calc_improvement<WIDE>(L, S, I, ws, sp.r, sp.o, sp.a, sp.new_a, mv));
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, d));
        acc = fmax(acc, v);
        rng.pos += spec_offset_after(sp, ws.lane, e, sp.nd);
        left -= e;
    }
    return acc;
}

// ------------------------------------------------------------------ a10: Greedy -----------------

// One sample of `amount` distinct non-trivial reads (IndexedRandom::sample -> index::sample_floyd):
// draw k is random_range(..=j_k), j_k = n_nt - amount + k; a draw equal to an earlier entry replaces
// that entry by j_k.  Lane k < amount receives draw k; sample_resolve applies the replacements.
// `slow`: take the draws one at a time (refills, bias correction) instead of through the lane-parallel path;
// the lane-parallel path consumes exactly `amount` draws of the current fill or nothing at all.
__device__ __forceinline__ bool sample_draw(Xo &rng, uint32_t n_nt, uint32_t amount, bool slow, uint32_t &myv) {
    const uint32_t lane = rng.lane;
    if (!slow) return xo_below_lanes(rng, amount, n_nt - amount + min(lane, amount - 1u) + 1u, myv);
    for (uint32_t k = 0; k < amount; k++) {
        const uint32_t t = xo_below(rng, n_nt - amount + k + 1u);
        if (lane == k) myv = t;
    }
    return true;
}
// Equal draws are found through a 12-entry scratch array in shared memory (three 16-byte loads per lane); MATCH.ANY
// did it in one instruction but took ~250 cycles to deliver (ncu: 6-8 % of the kernel waiting on it).
__device__ __forceinline__ void sample_resolve(const WarpShared &ws, uint32_t n_nt, uint32_t amount, uint32_t &myv) {
    const uint32_t lane = ws.lane;
    if (lane < 12u) ws.samp[lane] = lane < amount ? myv : 0xFFFFFF00u + lane;
    __syncwarp();
    const uint4 q0 = ((const uint4 *)ws.samp)[0], q1 = ((const uint4 *)ws.samp)[1], q2 = ((const uint4 *)ws.samp)[2];
    const uint32_t e[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
    bool dup = false;
#pragma unroll
    for (int k = 0; k < 11; k++) dup |= (uint32_t)k < lane && e[k] == myv;
    if (wany(dup && lane < amount)) {
        for (uint32_t k = 1; k < amount; k++) {
            const uint32_t t = wshfl(myv, (int)k);
            if (lane < k && myv == t) myv = n_nt - amount + k;
        }
    }
    __syncwarp();
}

// job word of a lane: read id | slot << 16 | candidate rank << 24 (slot < 16, rank < 256)
__device__ __forceinline__ uint32_t job_r(uint32_t j) { return j & 0xFFFFu; }
__device__ __forceinline__ uint32_t job_slot(uint32_t j) { return (j >> 16) & 0xFFu; }
__device__ __forceinline__ uint32_t job_c(uint32_t j) { return j >> 24; }
__device__ __forceinline__ uint32_t alt_rank(uint32_t rank, uint32_t a) { return rank < a ? rank : rank + 1u; }

// A candidate move as seen by one lane, ordered like the reference's two nested strict-'>' scans:
// larger improvement `s` first; among equal `s` the earlier sampled read (slot); inside that read the
// larger `improv` (pre-scaling value compared by best_read_improvement), then the lower candidate index.
template <bool WIDE>
struct Cand {
    double s, improv, dld, dlp;
    uint32_t job;
    typename RecWord<WIDE>::T raw_old, raw_new;
};
template <bool WIDE>
__device__ __forceinline__ bool cand_better(const Cand<WIDE> &a, const Cand<WIDE> &b) {
    if (a.s != b.s) return a.s > b.s;
    if (job_slot(a.job) != job_slot(b.job)) return job_slot(a.job) < job_slot(b.job);
    if (a.improv != b.improv) return a.improv > b.improv;
    return job_c(a.job) < job_c(b.job);
}

template <bool WIDE>
__device__ __forceinline__ void eval_cand(const LocusDev &L, const WarpShared &ws, double lp_old, double lp,
                                          typename RecWord<WIDE>::T raw_old, typename RecWord<WIDE>::T raw_new,
                                          uint32_t job, Cand<WIDE> &cd) {
    cd.dld = depth_lik_diff_raw<WIDE>(ws, raw_old, raw_new);
    cd.improv = __dadd_rn(lp, __dmul_rn(L.rel_contrib, cd.dld));          // assgn.rs:303
    cd.s = __dmul_rn(L.aln_contrib, __dsub_rn(cd.improv, lp_old));        // assgn.rs:310
    cd.dlp = __dsub_rn(lp, lp_old);
    cd.job = job; cd.raw_old = raw_old; cd.raw_new = raw_new;
}

// Greedy::solve_nontrivial (src/solvers/stoch.rs:81-120).  The winner of an iteration is found with REDUX
// reductions in the reference's tie order.
__device__ __forceinline__ uint32_t ld_rec(const uint32_t *p) { uint32_t v; asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint64_t ld_rec(const uint64_t *p) { uint64_t v; asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(p)); return v; }

// Sequential part of Floyd's duplicate handling: draw k equal to an earlier entry replaces that entry by j_k.
__device__ __noinline__ uint32_t sample_resolve_seq(uint32_t n_nt, uint32_t amount, uint32_t myv) {
    const uint32_t lane = lane_id();
    for (uint32_t k = 1; k < amount; k++) {
        const uint32_t t = wshfl(myv, (int)k);
        if (lane < k && myv == t) myv = n_nt - amount + k;
    }
    return myv;
}

// The register part of a sample in flight: this lane's job, the sample's layout.
struct JobRegs {
    uint32_t job;               // read id | slot << 16 | candidate rank << 24 (valid for lanes < min(total, 32))
    uint32_t lead;              // lanes < amount: sampled non-trivial read index | first job position << 16
    uint32_t total;             // jobs of the sample (warp-uniform)
};

template <bool WIDE>
__device__ void greedy_solve(const LocusDev &L, const StageParams &P, const Slab<WIDE> &S, const Instance &I,
                             const WarpShared &ws, Xo &rng) {
    typedef RecWord<WIDE> RW;
    typedef typename RW::T Rec;
    const uint32_t lane = ws.lane;
    const uint32_t amount = min(P.sample_size, I.n_nt);
    init_assignment<WIDE>(L, S, I, ws, rng, P.best_start ? 0 : 1);
    // min_diff is parked in shared memory and re-read where it is compared: kept in registers across the loop the
    // allocator spilled it to local memory (ncu: 5 % of the kernel waiting on the reload).
    {
        const double md = fmax(__dmul_rn(1e-10, max_abs_random<WIDE>(L, S, I, ws, rng)), 1e-14);
        if (lane == 0) ws.lik[3] = md;
        __syncwarp();
    }
    const volatile double *v_min_diff = ws.lik + 3;
    // 32-bit counters: lctp_solve_stage refuses plateau sizes that would need more
    uint32_t curr_plato = 0, it = 0;
    const uint32_t plato_size = (uint32_t)P.plato_size, max_iter = (uint32_t)P.max_iter;
    // The sample pipeline.  A sample takes `amount` consecutive draws of the worker's stream (Floyd: draw k is
    // random_range(..=n_nt - amount + k)); it moves through five stages, one per loop round:
    //   S0  the raw draws are fetched from the worker's pre-generated buffer
    //   S1  bounded values; MATCH.ANY looks for equal draws, the index entries of the sampled reads are fetched
    //   A   duplicates resolved (5 % of the samples: sequential path, index entries fetched again), jobs dealt:
    //       every (sampled read, alternative candidate) pair -- "flattened" best_read_improvement
    //       (src/model/assgn.rs:287-317) -- gets a lane by an exclusive prefix sum over the reads' alternative
    //       counts, so that (almost always) ONE evaluation pass covers the sample; first-hop loads (the private
    //       candidate records, the read's run offsets)
    //   B   second-hop loads (the shared ln-probabilities)
    //   C   evaluated: touches only registers and shared memory
    // A round is: C; wait for the copies issued in the previous round (a whole stage C ago); pick up their results
    // and compute every address; issue the copies of all stages back to back.
    // dpos = stream position behind the last fetched sample.  A sample that would need a bias-correction draw, or
    // that does not fit in the current fill, is not fetched: the pipeline drains and the sample is drawn one draw
    // at a time (refill, bias correction) on the empty pipeline -- the samples in flight are given back to the
    // stream when the loop ends, which cannot cross a refill.
    const uint32_t n_nt = I.n_nt;
    const uint32_t my_range = n_nt - amount + min(lane, amount - 1u) + 1u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t *draw_hi = (const uint32_t *)rng.buf + 1 + 2u * min(lane, amount - 1u);   // next_u32 = upper half
    // Loop-carried state of the loads is their ADDRESSES (ALU values), never a loaded value: the loads of all stages are
    // issued at the top of a round, stage C runs while they are in flight, and their results are read at the bottom of
    // the same round.  (With the loads issued at the bottom and read a round later the compiler moved every freshly
    // loaded value into its loop-carried register right behind the load -- a full L2 latency per round, ncu: 15 % of the
    // kernel in one register move.)
    uint32_t q_io = 0, q_in = 0;        // B: indices of the two ln-probabilities in cm_lnprob
    uint32_t q_oa = 0, q_oc = 0, q_r = 0;   // A: positions of the job's two candidate records, the job's read
    uint32_t q_nt = 0;                  // S1: index entry to fetch
    uint32_t q_at = min(rng.pos, RNG_FILL - amount);   // S0: stream position to fetch
    uint32_t dpos = rng.pos;
    bool blocked = false, slow = false;
    bool v0 = false, v1 = false, vA = false, vB = false, vC = false;
    uint32_t s_v = 0;                   // S1: sampled index (lanes < amount) ...
    bool s_dup = false;                 //     ... two of its draws are equal (Floyd's replacement needed)
    JobRegs ja, jb;                     // samples in stages A and B
    ja.job = ja.lead = ja.total = 0; jb = ja;
    Rec b_ro = 0, b_rn = 0;             // stage B: the records of its jobs
    struct { uint32_t job, lead, total; Rec ro, rn; double lpo, lpn; } cur;   // stage C
    cur.job = cur.lead = cur.total = 0; cur.ro = cur.rn = 0; cur.lpo = cur.lpn = 0.0;
    __syncwarp();
    for (;;) {
        // ---- the loads of all stages, back to back (through asm volatile: they stay here, in this order)
        double f_lpo, f_lpn;                // B: ln-probabilities of the job's current candidate / of its alternative
        Rec f_ro, f_rn;                     // A: their records ...
        uint32_t f_b0, f_b1;                //    ... and cm_off of the job's read on the first two haplotypes
        uint2 f_info;                       // S1: index entry of the sampled read (lanes < amount)
        uint32_t f_raw;                     // S0: upper half of the raw draw
        asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(f_lpo) : "l"(L.cm_lnprob + q_io));
        asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(f_lpn) : "l"(L.cm_lnprob + q_in));
        f_ro = ld_rec(S.rec + q_oa);
        f_rn = ld_rec(S.rec + q_oc);
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(f_b0) : "l"(L.cm_off + (size_t)I.h0 * L.R + q_r));
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(f_b1) : "l"(L.cm_off + (size_t)I.h1 * L.R + q_r));   // (ploidy 1: h1 = 0, loaded, not used)
        asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(f_info.x), "=r"(f_info.y) : "l"(ws.ntinfo + q_nt));
        asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(f_raw) : "l"(draw_hi + 2u * q_at));
        uint32_t mv_r = 0xFFFFFFFFu, mv_old = 0;     // read moved by this round's stage C
        bool refresh = false;                        // ... and the deferred refresh of its windows' products
        uint32_t rf_w = 0;
        double rf_t = 0.0;
        if (vC) {
            // ---- stage C: evaluate the sample of this iteration
            Cand<WIDE> best;
            best.s = -INFINITY; best.improv = -INFINITY; best.dld = 0.0; best.dlp = 0.0;
            best.job = 0x00FF0000u; best.raw_old = 0; best.raw_new = 0;
            if (lane < cur.total)
                eval_cand<WIDE>(L, ws, cur.lpo, cur.lpn, cur.ro, cur.rn, cur.job, best);
            if (__builtin_expect(cur.total > 32u, 0)) {
                // more jobs than lanes (rare): the remaining ones in further passes straight from memory.  Job f of
                // the sample belongs to the last sampled read whose first job position is <= f.
                const uint32_t first = cur.lead >> 16;
                uint32_t starts_before = (uint32_t)__popc(wballot(lane < amount && first < 32u));
                for (uint32_t f0 = 32u; f0 < cur.total; f0 += 32u) {
                    const unsigned starts = __reduce_or_sync(FULL, (lane < amount && first >= f0 && first < f0 + 32u) ? (1u << (first - f0)) : 0u);
                    const uint32_t slot = starts_before + (uint32_t)__popc(starts & (0xFFFFFFFFu >> (31u - lane))) - 1u;
                    const uint32_t idx = wshfl(cur.lead & 0xFFFFu, (int)slot);
                    const uint32_t sfirst = wshfl(first, (int)slot);
                    if (f0 + lane < cur.total) {
                        const uint2 inf = nt_info(ws, idx);
                        const uint32_t r = inf.y, o = inf.x & 0xFFFFu, a = ws.assgn[r];
                        const uint32_t c = alt_rank(f0 + lane - sfirst, a);
                        const Rec ro = S.rec[o + a], rn = S.rec[o + c];
                        const double lpo = __ldg(L.cm_lnprob + lp_index(L, I, ws, r, RW::src(ro)));
                        const double lpn = __ldg(L.cm_lnprob + lp_index(L, I, ws, r, RW::src(rn)));
                        Cand<WIDE> cd;
                        eval_cand<WIDE>(L, ws, lpo, lpn, ro, rn, r | (slot << 16) | (c << 24), cd);
                        if (cand_better<WIDE>(cd, best)) best = cd;
                    }
                    starts_before += (uint32_t)__popc(starts);
                }
            }
            // winner over the lanes' bests, in the order of cand_better
            int wl;
            {
                const unsigned long long ks = ord_key(best.s);
                const uint32_t h1 = __reduce_max_sync(FULL, (uint32_t)(ks >> 32));
                bool m = (uint32_t)(ks >> 32) == h1;
                const uint32_t l1 = __reduce_max_sync(FULL, m ? (uint32_t)ks : 0u);
                m = m && (uint32_t)ks == l1;
                const unsigned tied = wballot(m);
                if (__builtin_expect((tied & (tied - 1u)) == 0u, 1)) wl = __ffs(tied) - 1;      // unique maximum (the usual case)
                else {
                    const uint32_t sl = __reduce_min_sync(FULL, m ? job_slot(best.job) : 0xFFFFFFFFu);
                    m = m && job_slot(best.job) == sl;
                    const unsigned long long ki = ord_key(best.improv);
                    const uint32_t h2 = __reduce_max_sync(FULL, m ? (uint32_t)(ki >> 32) : 0u);
                    m = m && (uint32_t)(ki >> 32) == h2;
                    const uint32_t l2 = __reduce_max_sync(FULL, m ? (uint32_t)ki : 0u);
                    m = m && (uint32_t)ki == l2;
                    const uint32_t cm = __reduce_min_sync(FULL, m ? job_c(best.job) : 0xFFFFFFFFu);
                    wl = __ffs(wballot(m && job_c(best.job) == cm)) - 1;
                }
            }
            it++;
            // the winning lane decides and applies its own move (reassign, src/model/assgn.rs:331-343): nothing but
            // the four windows and the moved read has to reach the other lanes, and that goes through shared memory
            if (wany((int)lane == wl && best.s > *v_min_diff)) {
                if ((int)lane == wl) {
                    const uint32_t w1 = RW::w1(best.raw_old), w2 = RW::w2(best.raw_old);
                    const uint32_t w3 = RW::w1(best.raw_new), w4 = RW::w2(best.raw_new);
                    const uint32_t w_r = job_r(best.job);
                    ws.win.depth(w3) += 1;
                    ws.win.depth(w4) += 1;
                    ws.win.depth(w1) -= 1;
                    ws.win.depth(w2) -= 1;
                    const uint32_t old_a = ws.assgn[w_r];
                    ws.assgn[w_r] = (uint8_t)job_c(best.job);
                    ws.lik[1] = __dadd_rn(ws.lik[1], best.dld);       // depth_lik += ..., aln_lik += ... (assgn.rs:336-337)
                    ws.lik[0] = __dadd_rn(ws.lik[0], best.dlp);
                    *(uint4 *)ws.samp = make_uint4(w1, w2, w3, w4);
                    ws.samp[4] = w_r; ws.samp[5] = old_a;
                }
                __syncwarp();
                mv_r = ws.samp[4]; mv_old = ws.samp[5];
                // slide the product slices of the (up to four) windows whose depth changed.  Only the next stage C reads
                // them: the depth-table entries are fetched here and the products are stored at the end of the round, so
                // that the L2 latency of the fetch (ncu: 3.7 % of the kernel in the dependent DMUL) is not waited for.
                if (lane < 20u) {
                    rf_w = ws.samp[lane / 5u];
                    const int d = min(max((int)ws.win.depth(rf_w) + (int)(lane % 5u) - 2, 0), (int)ws.depth_k - 1);
                    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(rf_t) : "l"(L.depth_table + ws.win.row(rf_w) + d));
                }
                refresh = true;
                curr_plato = 0;
            } else {
                curr_plato += 1;
                if (curr_plato > plato_size) break;
            }
            if (it >= max_iter) {
                if (refresh) {                       // leave the window state complete (the debug outputs read it)
                    if (lane < 20u) ws.win.p(rf_w, (int)(lane % 5u)) = __dmul_rn(ws.win.weight(rf_w), rf_t);
                    __syncwarp();
                }
                break;
            }
        }
        // ---- stage S1 of the newest sample first: its MATCH.ANY is issued here and its result read at the end of the
        // round (~150 instructions later), so that its ~250 cycles are not waited for
        const uint32_t a_v0 = s_v;           // (stage A's sample, before S1 overwrites it)
        const bool a_valid = v1;
        unsigned n_match;
        {
            const uint64_t m = (uint64_t)f_raw * (uint64_t)my_range;     // rand UniformInt::sample_single_inclusive
            v1 = v0;
            if (__builtin_expect(v0 && wany(lane < amount && (uint32_t)m > 0u - my_range), 0)) {
                v1 = false; blocked = true; dpos -= amount;      // biased zone: this sample takes the sequential path
            }
            s_v = (uint32_t)(m >> 32);
            n_match = __match_any_sync(FULL, lane < amount ? s_v : 0x80000000u | lane);
        }
        // ---- pick up what the loads issued at the top of the round delivered, oldest first.  B -> C
        cur.job = jb.job; cur.lead = jb.lead; cur.total = jb.total; cur.ro = b_ro; cur.rn = b_rn;
        cur.lpo = f_lpo; cur.lpn = f_lpn;
        vC = vB;
        // ---- A -> B
        jb = ja;
        b_ro = f_ro; b_rn = f_rn;
        const uint32_t b0 = f_b0, b1 = f_b1;
        vB = vA;
        // samples in flight saw the old assignment of the read that stage C moved: redo their jobs on it (rare)
        if (__builtin_expect(mv_r != 0xFFFFFFFFu, 1)) {
            const bool hitC = vC && lane < cur.total && job_r(cur.job) == mv_r;
            const bool hitB = vB && lane < jb.total && job_r(jb.job) == mv_r;
            if (__builtin_expect(wany(hitC || hitB), 0)) {
                const uint32_t o = ws.off[mv_r], n = (uint32_t)ws.off[mv_r + 1] - o, a = ws.assgn[mv_r];
                // a lane's alternative number: invert alt_rank under the old assignment, redo it under the new one
                if (hitC) {
                    const uint32_t c_old = job_c(cur.job), alt = c_old > mv_old ? c_old - 1u : c_old;
                    const uint32_t c = min(alt_rank(alt, a), n - 1u);
                    cur.job = (cur.job & 0x00FFFFFFu) | (c << 24);
                    cur.ro = S.rec[o + a]; cur.rn = S.rec[o + c];
                    cur.lpo = __ldg(L.cm_lnprob + lp_index(L, I, ws, mv_r, RW::src(cur.ro)));
                    cur.lpn = __ldg(L.cm_lnprob + lp_index(L, I, ws, mv_r, RW::src(cur.rn)));
                }
                if (hitB) {
                    const uint32_t c_old = job_c(jb.job), alt = c_old > mv_old ? c_old - 1u : c_old;
                    const uint32_t c = min(alt_rank(alt, a), n - 1u);
                    jb.job = (jb.job & 0x00FFFFFFu) | (c << 24);
                    b_ro = S.rec[o + a]; b_rn = S.rec[o + c];
                }
            }
        }
        // ---- stage B: where the ln-probabilities of its jobs are
        uint32_t b_io, b_in;
        {
            const uint32_t r = job_r(jb.job);
            if (L.p <= 2) { b_io = lp_index2(L, r, RW::src(b_ro), b0, b1); b_in = lp_index2(L, r, RW::src(b_rn), b0, b1); }
            else { b_io = lp_index(L, I, ws, r, RW::src(b_ro)); b_in = lp_index(L, I, ws, r, RW::src(b_rn)); }
        }
        // ---- stage A: Floyd's replacements if two draws of the sample were equal, then deal the jobs
        uint32_t a_v = a_v0;
        uint2 a_info = f_info;
        if (__builtin_expect(slow, 0)) {                    // the sample was drawn one draw at a time (see below)
            a_v = ws.samp[lane & 15u];
            a_info = nt_info(ws, lane < amount ? a_v : 0u);
            slow = false;
        } else if (__builtin_expect(s_dup, 0)) {
            a_v = sample_resolve_seq(n_nt, amount, a_v);
            a_info = nt_info(ws, lane < amount ? a_v : 0u);
        }
        uint32_t a_r, a_oa, a_oc;       // the job's read, positions of its current / alternative candidate
        {
            const bool lead = lane < amount;
            const uint32_t n_alt = lead ? (a_info.x >> 16) - 1u : 0u;
            uint32_t incl = n_alt;
#pragma unroll
            for (int d = 1; d < 16; d <<= 1) {              // amount <= 11: four rounds
                const uint32_t t = __shfl_up_sync(FULL, incl, d);
                if (lane >= (uint32_t)d) incl += t;
            }
            ja.total = wshfl(incl, (int)amount - 1);
            const uint32_t first = incl - n_alt;
            ja.lead = a_v | (first << 16);
            const unsigned starts = __reduce_or_sync(FULL, (lead && first < 32u) ? (1u << first) : 0u);
            const uint32_t slot = (uint32_t)__popc(starts & (0xFFFFFFFFu >> (31u - lane))) - 1u;   // starts has bit 0 set
            a_r = wshfl(a_info.y, (int)slot);
            const uint32_t on = wshfl(a_info.x, (int)slot);
            const uint32_t alt = lane - wshfl(first, (int)slot);
            const uint32_t o = on & 0xFFFFu, n = on >> 16;
            const uint32_t a = ws.assgn[a_r];
            const uint32_t c = min(alt_rank(alt, a), n - 1u);     // lanes >= total: clamped, loaded, never evaluated
            ja.job = a_r | (slot << 16) | (c << 24);
            a_oa = o + a; a_oc = o + c;
        }
        vA = a_valid;
        // ---- what the next round fetches
        q_io = b_io; q_in = b_in;
        q_oa = a_oa; q_oc = a_oc; q_r = a_r;
        q_nt = lane < amount ? s_v : 0u;
        v0 = !blocked && dpos + amount <= RNG_FILL;
        q_at = min(dpos, RNG_FILL - amount);        // past the fill: re-read, not used
        dpos += v0 ? amount : 0u;
        s_dup = wany(v1 && lane < amount && (n_match & lt_mask) != 0u);
        if (refresh) {                               // second half of win_refresh (see stage C)
            if (lane < 20u) ws.win.p(rf_w, (int)(lane % 5u)) = __dmul_rn(ws.win.weight(rf_w), rf_t);
            __syncwarp();
        }
        if (__builtin_expect(!v0 && !v1 && !vA && !vB && !vC, 0)) {
            // empty pipeline: the next sample one draw at a time; stage A of the next round picks it up
            uint32_t v = 0;
            rng.pos = dpos;
            sample_draw(rng, n_nt, amount, true, v);
            sample_resolve(ws, n_nt, amount, v);
            if (lane < 16u) ws.samp[lane] = v;
            __syncwarp();
            slow = true; v1 = true; blocked = false; s_dup = false;
            dpos = rng.pos;
        }
    }
    // the pre-fetched samples of iterations that never ran go back to the stream
    rng.pos = dpos - amount * ((v0 ? 1u : 0u) + (v1 ? 1u : 0u) + (vA ? 1u : 0u) + (vB ? 1u : 0u));
    if (lane == 0) ((uint64_t *)ws.lik)[2] += it;
}

// ------------------------------------------------------------------ a10: Greedy, samples of more than 11 reads ---
//
// `-S greedy:s=N` with N > 11 (Greedy::set_sample_size accepts any N >= 1, src/solvers/stoch.rs:65-72).  rand 0.10's
// seq::index::sample then chooses between Floyd's algorithm and a partial Fisher-Yates shuffle of 0..length
// ("sample_inplace") by the published rule  amount > 11 && length < (C1 + 1.6 * amount) * amount  (f32 arithmetic;
// C1 = 10 for length < 500,000), and for amount >= 163 between sample_inplace (length < 270 * amount) and
// sample_rejection.  The rejection branch (a HashSet loop over a differently biased sampler) is not restated: a
// genotype that would take it sets bit 2 of the stage's error word and lctp_solve_stage reports LCTP_E_CAPACITY.
// This path is correct, not fast: draws are taken one at a time, each sampled read is one lane's job.
struct BigCand {
    double s, improv, dld, dlp;
    uint32_t slot, r, c;
};
__device__ __forceinline__ bool bigcand_better(const BigCand &a, const BigCand &b) {
    if (a.s != b.s) return a.s > b.s;
    if (a.slot != b.slot) return a.slot < b.slot;
    if (a.improv != b.improv) return a.improv > b.improv;
    return a.c < b.c;
}
// 0 = Floyd, 1 = in place, 2 = rejection (unsupported); restates the choice of rand::seq::index::sample
__device__ __forceinline__ int big_sample_algo(uint32_t length, uint32_t amount) {
    if (amount < 163u) {
        const float m4 = 1.6f * (float)amount;
        const float c1 = length >= 500000u ? 70.0f / 9.0f : 10.0f;
        const float c0 = length >= 500000u ? 8.0f / 45.0f * (float)amount : m4;
        return (amount > 11u && (float)length < (c1 + c0) * (float)amount) ? 1 : 0;
    }
    const float c = length < 500000u ? 270.0f : 330.0f / 9.0f;
    return (float)length < c * (float)amount ? 1 : 2;
}

template <bool WIDE>
__device__ bool greedy_solve_big(const LocusDev &L, const StageParams &P, const Slab<WIDE> &S, const Instance &I,
                                 const WarpShared &ws, Xo &rng, uint32_t *__restrict__ big) {
    typedef RecWord<WIDE> RW;
    const uint32_t lane = ws.lane;
    const uint32_t n_nt = I.n_nt, amount = min(P.sample_size, n_nt);
    init_assignment<WIDE>(L, S, I, ws, rng, P.best_start ? 0 : 1);
    const double min_diff = fmax(__dmul_rn(1e-10, max_abs_random<WIDE>(L, S, I, ws, rng)), 1e-14);
    const int algo = big_sample_algo(n_nt, amount);
    if (algo == 2) return false;
    uint32_t *sample = big;                  // [amount]
    uint32_t *swp = big + amount;            // [amount]  in-place: the partner of every swap (to undo it)
    uint32_t *perm = big + 2 * amount;       // [n_nt]    in-place: the index vector 0..length, restored after each sample
    if (algo == 1) { for (uint32_t i = lane; i < n_nt; i += 32) perm[i] = i; }
    __syncwarp();
    uint32_t curr_plato = 0, it = 0;
    const uint32_t plato_size = (uint32_t)P.plato_size, max_iter = (uint32_t)P.max_iter;
    for (; it < max_iter; it++) {
        // ---- the sample (all lanes run the same scalar code; lane 0 writes)
        if (algo == 1) {                     // sample_inplace: for i in 0..amount { j = random_range(i..length); swap(i, j) }
            for (uint32_t i = 0; i < amount; i++) {
                const uint32_t j = i + xo_below(rng, n_nt - i);
                if (lane == 0) { const uint32_t a = perm[i], b = perm[j]; perm[i] = b; perm[j] = a; swp[i] = j; sample[i] = b; }
                __syncwarp();
            }
        } else {                             // sample_floyd
            for (uint32_t k = 0; k < amount; k++) {
                const uint32_t j = n_nt - amount + k;
                const uint32_t t = xo_below(rng, j + 1u);
                bool hit = false;
                uint32_t pos = 0;
                for (uint32_t e = lane; e < k; e += 32) if (sample[e] == t) { hit = true; pos = e; }
                const unsigned m = wballot(hit);
                if (m) { const uint32_t p0 = wshfl(pos, __ffs(m) - 1); if (lane == 0) sample[p0] = j; }
                if (lane == 0) sample[k] = t;
                __syncwarp();
            }
        }
        // ---- best_read_improvement of every sampled read (src/model/assgn.rs:287-317), one read per lane
        BigCand best;
        best.s = -INFINITY; best.improv = -INFINITY; best.dld = best.dlp = 0.0; best.slot = 0xFFFFFFFFu; best.r = best.c = 0;
        for (uint32_t k = lane; k < amount; k += 32) {
            const uint2 inf = nt_info(ws, sample[k]);
            const uint32_t r = inf.y, o = inf.x & 0xFFFFu, n = inf.x >> 16, a = ws.assgn[r];
            const typename RW::T ro = S.rec[o + a];
            const double lpo = __ldg(L.cm_lnprob + lp_index(L, I, ws, r, RW::src(ro)));
            for (uint32_t c = 0; c < n; c++) {
                if (c == a) continue;
                const typename RW::T rn = S.rec[o + c];
                const double lpn = __ldg(L.cm_lnprob + lp_index(L, I, ws, r, RW::src(rn)));
                BigCand cd;
                cd.dld = depth_lik_diff_raw<WIDE>(ws, ro, rn);
                cd.improv = __dadd_rn(lpn, __dmul_rn(L.rel_contrib, cd.dld));
                cd.s = __dmul_rn(L.aln_contrib, __dsub_rn(cd.improv, lpo));
                cd.dlp = __dsub_rn(lpn, lpo);
                cd.slot = k; cd.r = r; cd.c = c;
                if (bigcand_better(cd, best)) best = cd;
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            BigCand o;
            o.s = __shfl_xor_sync(FULL, best.s, d); o.improv = __shfl_xor_sync(FULL, best.improv, d);
            o.dld = __shfl_xor_sync(FULL, best.dld, d); o.dlp = __shfl_xor_sync(FULL, best.dlp, d);
            o.slot = __shfl_xor_sync(FULL, best.slot, d); o.r = __shfl_xor_sync(FULL, best.r, d);
            o.c = __shfl_xor_sync(FULL, best.c, d);
            if (bigcand_better(o, best)) best = o;
        }
        if (algo == 1) {                     // put the index vector back (reverse order of the swaps)
            if (lane == 0)
                for (uint32_t i = amount; i-- > 0;) { const uint32_t j = swp[i]; const uint32_t a = perm[i], b = perm[j]; perm[i] = b; perm[j] = a; }
            __syncwarp();
        }
        if (best.s > min_diff) {
            const uint32_t o = ws.off[best.r], a = ws.assgn[best.r];
            Move<WIDE> mv;
            mv.dld = best.dld; mv.dlp = best.dlp; mv.raw_old = S.rec[o + a]; mv.raw_new = S.rec[o + best.c];
            apply_move<WIDE>(ws, L.depth_table, best.r, best.c, mv);
            curr_plato = 0;
        } else {
            curr_plato += 1;
            if (curr_plato > plato_size) { it++; break; }
        }
    }
    if (lane == 0) ((uint64_t *)ws.lik)[2] += it;
    return true;
}

// ------------------------------------------------------------------ a11: SimAnneal --------------

// Look-ahead over the next 32 stream positions.  A step's draws fix its read and (given the read's current
// assignment) its new location, and everything calculate_improvement loads from global memory for it -- the two
// candidate records and their ln-probabilities -- depends on nothing else.  So lane l prepares the step that WOULD
// start at stream position pos + l: all the dependent loads of ~12 future steps are in flight together, once per
// round, instead of once per step.  The walk then follows the actual chain of steps (2 or 3 draws each) one after
// the other on warp-uniform values: broadcast the prepared step of the lane at the current position, take the
// depth-likelihood difference from the shared-memory window state, decide, apply.  A prepared step whose read was
// reassigned earlier in the same walk is stale (its new location was derived from the old assignment): the walk
// stops there and the next round looks again.  Exact by construction: every step is evaluated against the state
// the sequential solver would see, with the same draws.
struct Look {
    bool ok;                     // the step at this lane's position needs no bias-correction draw and fits the fill
    uint32_t r, o, a, new_a;     // its target
    uint32_t nd;                 // draws of the target (1 or 2); the walk takes the U(0,1) after them from the ring
};
template <bool WITH_U>
__device__ __forceinline__ void look_targets(const WarpShared &ws, const Instance &I, Xo &rng, Look &lk) {
    const uint32_t lane = ws.lane;
    lk.ok = false; lk.r = lk.o = lk.a = lk.new_a = 0; lk.nd = 1;
    if (rng.pos >= RNG_FILL) return;
    const uint32_t avail = min(32u, RNG_FILL - rng.pos);
    stream_cover(rng, avail);              // every position a lane (or the walk) peeks at is in the ring
    const uint64_t d = stream_peek64(rng, lane);
    const uint64_t d1 = __shfl_down_sync(FULL, d, 1);
    const uint64_t m = (d >> 32) * (uint64_t)I.n_nt;
    const bool bias1 = (uint32_t)m > 0u - I.n_nt;
    const uint2 inf = nt_info(ws, (uint32_t)(m >> 32));
    lk.r = inf.y;
    lk.o = inf.x & 0xFFFFu;
    const uint32_t n = inf.x >> 16;
    lk.a = ws.assgn[lk.r];
    const bool two = n > 2;
    const uint64_t m2 = (d1 >> 32) * (uint64_t)(n - 1u);
    const bool bias2 = two && (uint32_t)m2 > 0u - (n - 1u);
    const uint32_t i2 = 1u + (uint32_t)(m2 >> 32);
    lk.new_a = two ? (i2 <= lk.a ? i2 - 1u : i2) : 1u - lk.a;
    lk.nd = two ? 2u : 1u;
    lk.ok = lane + lk.nd + (WITH_U ? 1u : 0u) <= avail && !bias1 && !bias2;
}

// The state-independent part of calculate_improvement: candidate records and aln-likelihood difference.
template <bool WIDE>
__device__ __forceinline__ void calc_static(const LocusDev &L, const Slab<WIDE> &S, const Instance &I,
                                            const WarpShared &ws, uint32_t r, uint32_t o, uint32_t a, uint32_t new_a,
                                            Move<WIDE> &mv) {
    typedef RecWord<WIDE> RW;
    mv.raw_old = S.rec[o + a];
    mv.raw_new = S.rec[o + new_a];
    const uint32_t so = RW::src(mv.raw_old), sn = RW::src(mv.raw_new);
    uint32_t io, in;
    if (L.p <= 2) {
        const uint32_t b0 = __ldg(L.cm_off + (size_t)I.h0 * L.R + r);
        const uint32_t b1 = L.p > 1 ? __ldg(L.cm_off + (size_t)I.h1 * L.R + r) : 0u;
        io = lp_index2(L, r, so, b0, b1);
        in = lp_index2(L, r, sn, b0, b1);
    } else {
        io = lp_index(L, I, ws, r, so);
        in = lp_index(L, I, ws, r, sn);
    }
    mv.dlp = __dsub_rn(__ldg(L.cm_lnprob + in), __ldg(L.cm_lnprob + io));
    mv.dld = 0.0;
}

// reassign for the walk: the window slices are SHIFTED instead of recomputed.  A depth change of +-1 moves the five
// products weight * table[row][d-2 .. d+2] by one place inside shared memory; only the entry that enters at the edge
// (k = 4 or 0) comes from the depth table in global memory.  That load is issued here and its product is stored by
// pend_flush -- before the next reassign, before a step that reads an edge entry (a depth change of +-2: both windows
// of a location equal), and at the end of the walk -- so the steps in between run while it is in flight.  The shifted
// values are the very products a recomputation would store (same weight, same table entry), so this is exact.
struct Pend { uint32_t idx; double tv; };      // idx = window | slice entry << 24 (0xFFFFFFFF = nothing pending)
static constexpr uint32_t PEND_NONE = 0xFFFFFFFFu;
__device__ __forceinline__ void pend_flush(const WarpShared &ws, Pend &pd) {
    if (pd.idx != PEND_NONE) {
        const uint32_t w = pd.idx & 0xFFFFFFu;
        ws.win.p(w, (int)(pd.idx >> 24)) = __dmul_rn(ws.win.weight(w), pd.tv);
        pd.idx = PEND_NONE;
    }
    __syncwarp();
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// Prepared step record (48 bytes in shared memory, one per look-ahead position, written by the lane of that position):
//   addr[8]  shared-memory addresses of the eight slice entries of depth_lik_diff (assgn.rs:259-284), in the order
//            p(w1, 2 + c1), p(w1, 2), p(w2, 2 + c2), p(w2, 2), ... -- the window merging (which of w1..w4 are equal, the
//            net depth change c_j carried by the first occurrence) is static, so it is done once per record by the
//            preparing lane instead of once per step by the whole warp; a merged-away window has c = 0: p - p = +0.0;
//   dlp      aln-likelihood difference;   packed = read | old rank << 16 | new rank << 24;
//   cpack    (c_j + 2) << 3 (j - 1), bit 12 = some |c_j| is 2.
template <bool WIDE>
__device__ __forceinline__ void step_record_store(const WarpShared &ws, const Look &lk, const Move<WIDE> &mv) {
    typedef RecWord<WIDE> RW;
    const uint32_t w1 = RW::w1(mv.raw_old), w2 = RW::w2(mv.raw_old), w3 = RW::w1(mv.raw_new), w4 = RW::w2(mv.raw_new);
    const int e21 = w2 == w1, e31 = w3 == w1, e32 = (w3 == w2) & !e31;
    const int e41 = w4 == w1, e42 = (w4 == w2) & !e41, e43 = (w4 == w3) & !e41 & !e42;
    const int c1 = -1 - e21 + e31 + e41;
    const int c2 = e21 ? 0 : -1 + e32 + e42;
    const int c3 = (e31 | e32) ? 0 : 1 + e43;
    const int c4 = (e41 | e42 | e43) ? 0 : 1;
    const int stride = (int)(ws.win.wp * 8u);
    const uint32_t b1 = ws.p2_sa + w1 * 8u, b2 = ws.p2_sa + w2 * 8u, b3 = ws.p2_sa + w3 * 8u, b4 = ws.p2_sa + w4 * 8u;
    const uint32_t ra = ws.steps_sa + ws.lane * (uint32_t)STEP_REC_BYTES;
    sts_v4(ra, make_uint4(b1 + (uint32_t)(c1 * stride), b1, b2 + (uint32_t)(c2 * stride), b2));
    sts_v4(ra + 16u, make_uint4(b3 + (uint32_t)(c3 * stride), b3, b4 + (uint32_t)(c4 * stride), b4));
    const uint32_t cpack = (uint32_t)(c1 + 2) | ((uint32_t)(c2 + 2) << 3) | ((uint32_t)(c3 + 2) << 6) |
                           ((uint32_t)(c4 + 2) << 9) | (((c1 == -2) | (c3 == 2)) ? 1u << 12 : 0u);
    const unsigned long long dl = (unsigned long long)__double_as_longlong(mv.dlp);
    sts_v4(ra + 32u, make_uint4((uint32_t)dl, (uint32_t)(dl >> 32), lk.r | (lk.a << 16) | (lk.new_a << 24), cpack));
}

// The prepared step at walk position q, read by the whole warp (broadcast loads).
struct Step { uint4 a0, a1; double dlp, dld; uint32_t r, new_a, cpack; };
__device__ __forceinline__ bool walk_fetch(const WarpShared &ws, uint32_t q, Step &st, Pend &pd) {
    const uint32_t ra = ws.steps_sa + q * (uint32_t)STEP_REC_BYTES;
    const uint4 m = lds_v4(ra + 32u);
    st.a0 = lds_v4(ra); st.a1 = lds_v4(ra + 16u);      // the three loads of the record are in flight together
    st.r = m.z & 0xFFFFu;
    st.new_a = m.z >> 24;
    st.cpack = m.w;
    st.dlp = __longlong_as_double((long long)(((unsigned long long)m.y << 32) | m.x));
    const uint32_t cur = ws.assgn[st.r];
    // a depth change of two reads an edge entry of the slices (see apply_step)
    if (m.w & (1u << 12)) pend_flush(ws, pd);
    // the eight slice entries are read next to the assignment check, not behind it (a stale step wastes them)
    const double p1 = lds_f64(st.a0.x), q1 = lds_f64(st.a0.y), p2 = lds_f64(st.a0.z), q2 = lds_f64(st.a0.w);
    const double p3 = lds_f64(st.a1.x), q3 = lds_f64(st.a1.y), p4 = lds_f64(st.a1.z), q4 = lds_f64(st.a1.w);
    if (cur != ((m.z >> 16) & 0xFFu)) return false;        // reassigned since the look-ahead: stale
    double s = __dadd_rn(__dsub_rn(p1, q1), __dsub_rn(p2, q2));
    s = __dadd_rn(s, __dsub_rn(p3, q3));
    st.dld = __dadd_rn(s, __dsub_rn(p4, q4));
    return true;
}

// reassign (src/model/assgn.rs:331-343) for a prepared step: slices shifted, edge entries deferred (see Pend).
__device__ __forceinline__ void apply_step(const WarpShared &ws, const double *__restrict__ table, const Step &st, Pend &pd) {
    pend_flush(ws, pd);
    const uint32_t j = ws.lane / 5u;
    const int k = (int)(ws.lane - 5u * j);
    const uint32_t bj = j == 0 ? st.a0.y : j == 1 ? st.a0.w : j == 2 ? st.a1.y : st.a1.w;     // address of p(w_j, 2)
    const int c = ws.lane >= 20u ? 0 : (int)((st.cpack >> (3u * j)) & 7u) - 2;
    const int stride = (int)(ws.win.wp * 8u);
    const int src = k + c;
    const bool inside = src >= 0 && src <= 4;
    double v = 0.0;
    if (c != 0) {
        if (inside) v = lds_f64(bj + (uint32_t)((src - 2) * stride));
        else {
            const uint32_t w = (bj - ws.p2_sa) >> 3;
            const int d = min(max((int)ws.win.depth(w) + c + k - 2, 0), (int)ws.depth_k - 1);
            pd.tv = __ldg(table + ws.win.row(w) + d);
            pd.idx = w | ((uint32_t)k << 24);
        }
    }
    __syncwarp();                                      // every old slice entry and depth has been read
    // depth[w] += c for the first occurrence of every distinct window (the lane with k = 0 of its group of five; the
    // net changes of merged windows are what the four +-1 updates of reassign add up to); lanes 20 / 21 add the two
    // likelihood differences, lane 22 stores the assignment
    if (c != 0 && k == 0) ws.win.depth((bj - ws.p2_sa) >> 3) += c;
    if (ws.lane == 20u || ws.lane == 21u) {            // depth_lik += ..., aln_lik += ... (assgn.rs:336-337)
        const uint32_t which = 21u - ws.lane;          // lik[1] = depth_lik (lane 20), lik[0] = aln_lik (lane 21)
        ws.lik[which] = __dadd_rn(ws.lik[which], which ? st.dld : st.dlp);
    }
    if (ws.lane == 22u) ws.assgn[st.r] = (uint8_t)st.new_a;
    if (c != 0 && inside) sts_f64(bj + (uint32_t)((k - 2) * stride), v);
    // a change of +-2 leaves an entry next to the centre pending, and those are read by ordinary steps
    if (st.cpack & (1u << 12)) pend_flush(ws, pd);
    else __syncwarp();
}

// The acceptance test of a step with a negative diff, `rng.random::<f64>() <= (diff / temp).exp()` (stoch.rs:216), v = the
// raw draw.  An f64 division and exponential are ~80 FP64-pipe instructions, and 70 % of the annealing steps need
// them; the FP64 pipe (16 lanes per clock per sub-partition) is shared by the four resident workers.  The test is
// therefore decided in single precision whenever the two sides are further apart than that evaluation can be wrong:
// x32 = diff / temp within 4e-7 relative (two conversions, __fdividef), __expf within 2 + 1.16 |x| ulp, and u is taken
// from the top 24 bits of the draw, u in [ulo, ulo + 2^-24).  The margin (1 + |x|) * 4e-6 is more than four times the
// sum of these.  Only the remaining cases (about one in 1e5) evaluate the f64 expression, so the decision is always the
// one the f64 expression gives.
__device__ __forceinline__ bool anneal_accept(uint64_t v, double diff, double temp) {
    const float df = (float)diff, tf = (float)temp;
    const bool sane = tf >= 1e-30f && tf <= 1e30f && df <= -1e-30f && df >= -1e30f;
    if (sane) {
        const float x = __fdividef(df, tf);                   // x <= 0; -inf when the quotient overflows
        if (x <= -80.f) return (v >> 11) == 0;                // exp(...) < 2^-53: only u = 0 passes
        const float e = __expf(x);
        const float m = (1.f + fabsf(x)) * 4e-6f;
        const float ulo = (float)(uint32_t)(v >> 40) * (1.f / 16777216.f);
        if (ulo + (1.f / 16777216.f) <= e * (1.f - m)) return true;
        if (ulo > e * (1.f + m)) return false;
    }
    return u64_to_unit_f64(v) <= exp(__ddiv_rn(diff, temp));
}

// SimAnneal::solve_nontrivial (src/solvers/stoch.rs:197-242).
template <bool WIDE>
__device__ void anneal_solve(const LocusDev &L, const StageParams &P, const Slab<WIDE> &S, const Instance &I,
                             const WarpShared &ws, Xo &rng) {
    init_assignment<WIDE>(L, S, I, ws, rng, 1);
    const double max_abs = max_abs_random<WIDE>(L, S, I, ws, rng);
    const double min_diff = fmax(__dmul_rn(1e-10, max_abs), 1e-14);
    const double start_temp = fmax(__ddiv_rn(-max_abs, P.ln_init_prob), 1e-5);
    const double temp_step = __ddiv_rn(start_temp, (double)P.anneal_steps);
    uint64_t curr_plato = 0, steps = 0;
    // phase 1: annealing, i = anneal_steps .. 1
    uint64_t i = P.anneal_steps;
    bool plateau = false;
    Pend pd;
    pd.idx = PEND_NONE; pd.tv = 0.0;
    while (i >= 1 && !plateau) {
        Look lk;
        look_targets<true>(ws, I, rng, lk);
        const unsigned okmask = wballot(lk.ok);
        if (!(okmask & 1u)) {              // sequential step (refill / biased draw / end of the fill)
            const Target t = random_target(ws, I, rng);
            Move<WIDE> mv;
            const double diff = __dsub_rn(calc_improvement<WIDE>(L, S, I, ws, t.r, t.o, t.a, t.new_a, mv), min_diff);
            steps++;
            bool accept = diff >= 0.0;
            if (!accept) accept = anneal_accept(xo_next(rng), diff, __dmul_rn(temp_step, (double)i));
            i--;
            if (accept) { apply_move<WIDE>(ws, L.depth_table, t.r, t.new_a, mv); curr_plato = 0; }
            else { curr_plato += 1; if (curr_plato >= P.plato_size) break; }
            continue;
        }
        if (lk.ok) {
            Move<WIDE> mv;
            calc_static<WIDE>(L, S, I, ws, lk.r, lk.o, lk.a, lk.new_a, mv);
            step_record_store<WIDE>(ws, lk, mv);
        }
        const unsigned twomask = wballot(lk.nd == 2u);
        __syncwarp();
        uint32_t q = 0;
        while (q < 32u && ((okmask >> q) & 1u) && i >= 1) {
            Step st;
            if (!walk_fetch(ws, q, st, pd)) break;
            const double diff = __dsub_rn(__dadd_rn(__dmul_rn(L.depth_contrib, st.dld), __dmul_rn(L.aln_contrib, st.dlp)), min_diff);
            uint32_t len = 1u + ((twomask >> q) & 1u);
            steps++;
            bool accept = diff >= 0.0;
            if (!accept) {                 // the U(0,1) is drawn only for a negative diff (short-circuit, stoch.rs:216)
                accept = anneal_accept(stream_peek64(rng, q + len), diff, __dmul_rn(temp_step, (double)i));
                len++;
            }
            i--;
            q += len;
            if (accept) { apply_step(ws, L.depth_table, st, pd); curr_plato = 0; }
            else { curr_plato += 1; if (curr_plato >= P.plato_size) { plateau = true; break; } }
        }
        pend_flush(ws, pd);
        rng.pos += q;
    }
    // phase 2: hill climbing until the plateau.  Nearly every step is rejected here (a rejected step changes nothing but
    // the plateau counter), so whole chains of steps are evaluated in parallel against the current state (spec_targets)
    // and committed up to the first accepted one.
    uint64_t k = 0;
    while (k < P.max_iter && curr_plato < P.plato_size) {
        Spec sp;
        spec_targets<false>(ws, I, rng, sp);
        if (sp.count == 0) {
            const Target t = random_target(ws, I, rng);
            Move<WIDE> mv;
            const double diff = calc_improvement<WIDE>(L, S, I, ws, t.r, t.o, t.a, t.new_a, mv);
            steps++; k++;
            if (diff > min_diff) { apply_move<WIDE>(ws, L.depth_table, t.r, t.new_a, mv); curr_plato = 0; }
            else curr_plato += 1;
            continue;
        }
        const uint32_t limit = (uint32_t)min((uint64_t)sp.count, P.max_iter - k);
        Move<WIDE> mv;
        mv.dld = mv.dlp = 0.0; mv.raw_old = mv.raw_new = 0;
        bool acc = false;
        if (sp.usable && sp.rank < limit)
            acc = calc_improvement<WIDE>(L, S, I, ws, sp.r, sp.o, sp.a, sp.new_a, mv) > min_diff;
        const unsigned accmask = wballot(acc);
        const uint32_t first_acc = accmask ? (uint32_t)__popc(sp.umask & ((1u << (__ffs(accmask) - 1)) - 1u)) : 0xFFFFFFFFu;
        const uint32_t n_rej = min(first_acc, limit);
        const uint64_t plato_left = P.plato_size - curr_plato;
        if ((uint64_t)n_rej >= plato_left) {
            const uint32_t e = (uint32_t)plato_left;
            rng.pos += spec_offset_after(sp, ws.lane, e, sp.nd);
            steps += e; k += e; curr_plato = P.plato_size;
            break;
        }
        steps += n_rej; k += n_rej; curr_plato += n_rej;
        if (first_acc < limit) {
            const int src = __ffs(accmask) - 1;
            const uint32_t off_acc = wshfl(ws.lane + sp.nd, src);
            Move<WIDE> w;
            w.dld = wshfl(mv.dld, src); w.dlp = wshfl(mv.dlp, src);
            w.raw_old = wshfl(mv.raw_old, src); w.raw_new = wshfl(mv.raw_new, src);
            const uint32_t w_r = wshfl(sp.r, src), w_new = wshfl(sp.new_a, src);
            rng.pos += off_acc;
            apply_move<WIDE>(ws, L.depth_table, w_r, w_new, w);
            steps++; k++; curr_plato = 0;
        } else {
            rng.pos += spec_offset_after(sp, ws.lane, n_rej, sp.nd);
        }
    }
    if (ws.lane == 0) ((uint64_t *)ws.lik)[2] += steps;
}

// ------------------------------------------------------------------ stage kernel ----------------

// Resident CTAs per SM the register allocation is sized for.  Shared memory allows 14 at the C2 shape (15 KB per
// worker); with 16 the 128-register cap spilled min_diff and the sampling ranges into local memory inside the greedy
// loop, and with the L1 carved down to ~30 KB those reloads go to the L2 (ncu: 10 % of the kernel).
#ifndef LCTP_MIN_CTAS
#define LCTP_MIN_CTAS 16
#endif
#ifndef LCTP_ANNEAL_CTAS_DEFAULT
#define LCTP_ANNEAL_CTAS_DEFAULT 24
#endif
// MODE: 0 = greedy (samples of <= 11 reads), 1 = greedy with larger samples, 2 = simulated annealing.  One kernel per
// solver: the code of the others is not in the instruction stream (the single kernel was 364 KB of SASS and lost
// 10 % of its cycles to instruction-cache misses with 16 workers per SM in different phases).
// NCTA: resident workers per SM the registers are capped for (annealing is instantiated for more than one value: its
// walk needs fewer registers than the greedy pipeline, and more resident chains hide more of its latency).
template <bool WIDE, int MODE, int NCTA = LCTP_MIN_CTAS>
__global__ void __launch_bounds__(CTA_THREADS, NCTA)
k_solve_stage(LocusDev L, StageParams P, const uint64_t *__restrict__ worker_ixs,
              const uint64_t *__restrict__ worker_off, const uint32_t *__restrict__ tuples,
              uint64_t *__restrict__ rng_states, double *__restrict__ lik_mean, double *__restrict__ lik_var,
              double *__restrict__ liks, uint64_t *__restrict__ n_alns, uint64_t *__restrict__ iters,
              uint16_t *__restrict__ counts, unsigned char *__restrict__ scratch,
              unsigned int *__restrict__ work_counter, int *__restrict__ err,
              const ulonglong2 *__restrict__ jump_tabs, DbgOut dbg) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t lane_reg;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_reg));      // read once (see lane_id)
    const int lane = (int)lane_reg;
    WarpShared ws;
    Xo rng;
    Slab<WIDE> S;
    slab_layout<WIDE>(P.cap, scratch + (size_t)blockIdx.x * P.slab_bytes, S);
    {
        unsigned char *base = smem;
        rng.ring = (uint64_t *)base;           base += 1024;
        ws.samp = (uint32_t *)base;            base += 64;
        ws.lik = (double *)base;               base += 32;
        ws.steps_sa = (uint32_t)__cvta_generic_to_shared(base);
        if (MODE == 2) base += 32 * STEP_REC_BYTES;
        ws.win.base = (double *)base;          ws.win.wp = win_stride(P.Wmax);
        ws.p2_sa = (uint32_t)__cvta_generic_to_shared(ws.win.base + 2u * ws.win.wp);
        base += align_up((size_t)ws.win.wp * 56, 16);
        if (P.nt_global) {
            ws.off = S.off;
            ws.ntinfo = (uint2 *)((unsigned char *)S.off + align_up(((size_t)L.R + 1) * 2, 128));
            ws.nt_read = nullptr;
        } else {
            ws.off = (uint16_t *)base;         base += align_up(((size_t)L.R + 1) * 2, 16);
            ws.nt_read = (uint16_t *)base;     base += align_up((size_t)L.R * 2, 16);
            ws.ntinfo = nullptr;
        }
        ws.assgn = (uint8_t *)base;            base += align_up((size_t)L.R, 16);
        ws.unm_bits = (uint32_t *)base;        base += align_up((size_t)((L.R + 31) / 32) * 4, 16);
        ws.haps = (uint32_t *)base;
        ws.zero_row = LCTP_GC_BINS * L.depth_k;
        ws.depth_k = L.depth_k;
        ws.lane = lane_reg;
    }
    rng.buf = S.rng_buf; rng.blk = S.rng_blk; rng.tabs = jump_tabs; rng.lane = lane_reg;
    rng.pos = 0; rng.base = 0; rng.pend = 0;

    // Everything that is only needed between genotypes (position, prior, count pointer, iteration total) is re-read
    // or kept in shared memory instead of living in registers across the solver loops.
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(work_counter, 1u);
        w = wshfl(w, 0);
        if (w >= P.n_workers) break;
        stream_begin(rng, rng_states + 4 * (size_t)w);
        for (uint64_t j = worker_off[w]; j < worker_off[w + 1]; j++) {
            Instance I;
            __syncwarp();
            if (lane == 0) {
                uint32_t wsft = 2;
                for (uint32_t k = 0; k < L.p; k++) {
                    const uint32_t h = tuples[j * L.p + k];
                    ws.haps[k] = h;
                    ws.haps[LCTP_MAX_PLOIDY + k] = wsft;
                    wsft += L.hap_n_windows[h];
                }
                ws.haps[LCTP_MAX_PLOIDY + L.p] = wsft;
                ((uint64_t *)ws.lik)[2] = 0;
            }
            __syncwarp();
            I.h0 = ws.haps[0];
            I.h1 = L.p > 1 ? ws.haps[1] : 0u;
            I.W = ws.haps[LCTP_MAX_PLOIDY + L.p];
            const bool ok = L.p <= 2 ? build_instance<LCTP_HEADS, WIDE>(L, S, ws, P.cap, I)
                                     : build_instance<0, WIDE>(L, S, ws, P.cap, I);
            if (lane == 0) n_alns[j] = I.A;
            if (!ok) {
                if (lane == 0) { atomicOr(err, 1); lik_mean[j] = NAN; lik_var[j] = NAN; iters[j] = 0; }
                continue;
            }
            if (P.want_counts) {
                uint16_t *cnt = counts + (size_t)j * P.cap;
                for (uint32_t c = lane; c < I.A; c += 32) cnt[c] = 0;
            }
            for (uint32_t a = 0; a < P.attempts; a++) {
                apply_tweak<WIDE>(L, S, I, ws, rng);
                if (I.n_nt == 0) init_assignment<WIDE>(L, S, I, ws, rng, 0);
                else if (MODE == 1) {
                    // scratch behind the worker's slab: sample, swap partners, index vector
                    uint32_t *big = (uint32_t *)(scratch + (size_t)blockIdx.x * P.slab_bytes + P.slab_bytes - P.big_bytes);
                    if (!greedy_solve_big<WIDE>(L, P, S, I, ws, rng, big) && lane == 0) atomicOr(err, 4);
                } else if (MODE == 0) greedy_solve<WIDE>(L, P, S, I, ws, rng);
                else anneal_solve<WIDE>(L, P, S, I, ws, rng);
                __syncwarp();
                if (lane == 0) {
                    // likelihood (assgn.rs:235-237) + prior (solve.rs:1126)
                    const double prior = L.priors ? L.priors[worker_ixs[j]] : 0.0;
                    liks[j * P.attempts + a] = __dadd_rn(prior, __dadd_rn(__dmul_rn(L.depth_contrib, ws.lik[1]),
                                                                          __dmul_rn(L.aln_contrib, ws.lik[0])));
                }
                if (dbg.lik) {         // ReadAssignment::summarize / write_depth (assgn.rs:356-372,413-425)
                    const size_t ja = (size_t)j * P.attempts + a;
                    if (lane == 0) {
                        dbg.lik[2 * ja] = ws.lik[0]; dbg.lik[2 * ja + 1] = ws.lik[1];
                        dbg.cnt[2 * ja] = ws.win.depth(0); dbg.cnt[2 * ja + 1] = ws.win.depth(1);
                    }
                    if (dbg.ww)
                        for (uint32_t w = lane; w < dbg.wmax; w += 32) {      // windows the genotype does not have: zeros
                            dbg.ww[ja * dbg.wmax + w] = w < I.W ? ws.win.weight(w) : 0.0;
                            dbg.wd[ja * dbg.wmax + w] = w < I.W ? ws.win.depth(w) : 0u;
                            dbg.wl[ja * dbg.wmax + w] = w < I.W ? ws.win.p(w, 2) : 0.0;
                        }
                }
                if (P.want_counts) {   // update_counts (assgn.rs:374-378)
                    uint16_t *cnt = counts + (size_t)j * P.cap;
                    for (uint32_t r = lane; r < L.R; r += 32) cnt[(uint32_t)ws.off[r] + ws.assgn[r]] += 1;
                }
                __syncwarp();
            }
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
                iters[j] = ((const uint64_t *)ws.lik)[2];
            }
            __syncwarp();
        }
        stream_end(rng, rng_states + 4 * (size_t)w);
        __syncwarp();
    }
}

// ------------------------------------------------------------------ host: jump tables -----------

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
// Byte-indexed form of M: tab[(b * 256 + v) * 4 .. +4] = M * (v << 8b), b = state byte 0..31.
static void bm_to_table(const BitMat &M, uint64_t *tab) {
    for (int b = 0; b < 32; b++) {
        uint64_t *t = tab + (size_t)b * 256 * 4;
        t[0] = t[1] = t[2] = t[3] = 0;
        for (int v = 1; v < 256; v++) {
            const int low = __builtin_ctz(v);
            const uint64_t *prev = t + (size_t)(v & (v - 1)) * 4;
            const uint64_t *col = M.c[b * 8 + low];
            for (int k = 0; k < 4; k++) t[(size_t)v * 4 + k] = prev[k] ^ col[k];
        }
    }
}

// The jump tables depend only on the generator: computed once per process (contexts on several host
// threads share them), uploaded once per context.
static const std::vector<uint64_t> &jump_tables() {
    static std::once_flag once;
    static std::vector<uint64_t> tabs;
    std::call_once(once, [] {
        static BitMat T, M;
        bm_step(T);
        tabs.resize((size_t)N_JUMP_TABS * JUMP_TAB_WORDS);
        static BitMat S, P, Q;
        bm_pow(T, (unsigned)RNG_C, S);                     // S = T^C; table l = S^l
        P = S;
        for (int l = 1; l < 32; l++) {
            bm_to_table(P, &tabs[(size_t)l * JUMP_TAB_WORDS]);
            bm_mul(S, P, Q); P = Q;
        }
        bm_to_table(P, &tabs[(size_t)32 * JUMP_TAB_WORDS]);   // S^32 = T^FILL
        (void)M;
    });
    return tabs;
}

static int ensure_jump_tables(lctp_ctx *ctx) {
    if (ctx->d_rng_mats.p) return LCTP_OK;
    const std::vector<uint64_t> &t = jump_tables();
    int rc = ctx->d_rng_mats.alloc(t.size());
    if (rc) return rc;
    LCTP_CUDA_CHECK(cudaMemcpyAsync(ctx->d_rng_mats.p, t.data(), t.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    LCTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return LCTP_OK;
}

// ------------------------------------------------------------------ host launch -----------------

static double dbg_now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

static int launch_stage_kernel(lctp_locus_h *h, const StageParams &P, size_t n_workers, size_t n_total, bool want_counts) {
    lctp_ctx *ctx = h->ctx;
    cudaStream_t s = ctx->stream;
    const LocusDev &L = h->dev;
    const size_t smem = group_smem_bytes(P.Wmax, L.R, P.nt_global != 0, P.kind == 1);
    if (smem > ctx->smem_optin) {
        set_error("lctp_solve_stage: %zu bytes of shared memory per worker needed (R=%u reads, %u windows); "
                  "loci this large are not supported by the shared-memory resident solver", smem, L.R, P.Wmax);
        return LCTP_E_CAPACITY;
    }
    const bool bigs = P.kind == 0 && P.sample_size > (uint32_t)MAX_SAMPLE;
    const int mode = P.kind == 1 ? 2 : bigs ? 1 : 0;
    // annealing: register cap for 16 / 20 / 24 / 28 / 32 resident workers per SM (LCTP_ANNEAL_CTAS, tuning knob; C3 stage
    // launch 915 / 755 / 733 ms at 16 / 20 / 24: the walk is a dependent chain, resident chains are what hides it)
    const int acta = env_int("LCTP_ANNEAL_CTAS", LCTP_ANNEAL_CTAS_DEFAULT);
    const int variant = mode != 2 ? 0 : acta >= 32 ? 4 : acta >= 28 ? 3 : acta >= 24 ? 2 : acta >= 20 ? 1 : 0;
    typedef decltype(&k_solve_stage<false, 0>) KernT;
    static const KernT anneal_kerns[2][5] = {
        {k_solve_stage<false, 2, 16>, k_solve_stage<false, 2, 20>, k_solve_stage<false, 2, 24>, k_solve_stage<false, 2, 28>, k_solve_stage<false, 2, 32>},
        {k_solve_stage<true, 2, 16>, k_solve_stage<true, 2, 20>, k_solve_stage<true, 2, 24>, k_solve_stage<true, 2, 28>, k_solve_stage<true, 2, 32>}};
    KernT kern = mode == 2 ? anneal_kerns[P.narrow_w ? 0 : 1][variant]
                           : P.narrow_w ? (mode == 1 ? k_solve_stage<false, 1> : k_solve_stage<false, 0>)
                                        : (mode == 1 ? k_solve_stage<true, 1> : k_solve_stage<true, 0>);
    // function attributes are per-device state shared by every context: configure + launch under one lock
    static std::mutex launch_mutex;
    std::lock_guard<std::mutex> lock(launch_mutex);
    // Only raise the limit when needed: re-setting a function attribute makes the next launch of the function wait
    // for its running instances, which serialised the stage kernels of loci in flight on different contexts.
    static std::unordered_map<int, size_t> smem_limit;     // by device ordinal and kernel instantiation
    size_t &lim = smem_limit[ctx->device * 32 + (P.narrow_w ? 1 : 0) + mode * 2 + variant * 6];
    if (smem > 48 * 1024 && smem > lim) {
        LCTP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lim = smem;
    }
    int occ = 0;
    LCTP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, CTA_THREADS, smem));
    if (occ < 1) occ = 1;
    occ = std::min(occ, std::max(1, env_int("LCTP_MAX_CTAS_PER_SM", 1 << 20)));
    uint32_t resident = (uint32_t)ctx->sm_count * occ;
    // Keep the private slabs of the resident workers inside the L2 (126 MB on B200, shared with the locus'
    // read-only arrays and with the other loci in flight): the slab is what the greedy / annealing loops gather
    // from at random, and a worker whose slab has been evicted runs at DRAM latency.  Never fewer than one
    // worker per SM sub-partition.
    {
        const size_t budget = (size_t)std::max(1, env_int("LCTP_L2_BUDGET_MB", 1024)) << 20;
        const size_t per = std::max<size_t>(1, P.slab_bytes);
        const uint32_t by_l2 = (uint32_t)std::max<size_t>(budget / per, (size_t)ctx->sm_count * 4);
        resident = std::min(resident, by_l2);
    }
    if (ctx->max_resident_workers && ctx->max_resident_workers < resident) resident = ctx->max_resident_workers;
    const uint32_t grid = (uint32_t)std::min<size_t>(n_workers, resident);

    int rc;
    if ((rc = ctx->scratch.ensure((size_t)grid * P.slab_bytes))) return rc;
    DbgOut dbg = {nullptr, nullptr, nullptr, nullptr, nullptr, 0};
    if (ctx->dbg_req) {
        const size_t na = n_total * P.attempts;
        if ((rc = ctx->d_dbg_lik.ensure(2 * na))) return rc;
        if ((rc = ctx->d_dbg_cnt.ensure(2 * na))) return rc;
        dbg.lik = ctx->d_dbg_lik.p; dbg.cnt = ctx->d_dbg_cnt.p;
        if (ctx->dbg_req->win_weight && ctx->dbg_req->win_depth && ctx->dbg_req->win_lik) {
            dbg.wmax = ctx->dbg_req->wmax;
            if ((rc = ctx->d_dbg_ww.ensure(na * dbg.wmax))) return rc;
            if ((rc = ctx->d_dbg_wl.ensure(na * dbg.wmax))) return rc;
            if ((rc = ctx->d_dbg_wd.ensure(na * dbg.wmax))) return rc;
            dbg.ww = ctx->d_dbg_ww.p; dbg.wl = ctx->d_dbg_wl.p; dbg.wd = ctx->d_dbg_wd.p;
        }
    }
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[2], s));
    kern<<<grid, CTA_THREADS, smem, s>>>(
        L, P, ctx->d_worker_ixs.p, ctx->d_worker_off.p, ctx->d_tuples.p, ctx->d_rng.p, ctx->d_lik_mean.p,
        ctx->d_lik_var.p, ctx->d_liks.p, ctx->d_nalns.p, ctx->d_iters.p, want_counts ? ctx->d_counts.p : nullptr,
        ctx->scratch.p, (unsigned int *)(ctx->d_flags.p + 1), ctx->d_flags.p, (const ulonglong2 *)ctx->d_rng_mats.p, dbg);
    ctx->launches++;
    LCTP_CUDA_CHECK(cudaGetLastError());
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[3], s));
    return LCTP_OK;
}

int launch_stage(lctp_locus_h *h, const lctp_stage *st, const uint64_t *worker_ixs, const uint64_t *worker_off,
                 size_t n_workers, uint64_t *worker_rng, double *lik_mean, double *lik_var, double *liks,
                 uint64_t *counts_off, uint16_t *counts, uint64_t counts_cap, uint64_t *n_alns_out,
                 uint64_t *iters_out) {
    return launch_stage_ex(h, st, worker_ixs, worker_off, n_workers, worker_rng, lik_mean, lik_var, liks, counts_off,
                           counts, counts_cap, n_alns_out, iters_out, false);
}

// device_only: validate, upload, launch and return; the results stay on the device (ctx->d_lik_mean / d_lik_var /
// d_rng / d_flags, indexed like the host outputs) for the multi-GPU exchange (dist.cu), which never needs them on
// the host of the rank that computed them.
int launch_stage_ex(lctp_locus_h *h, const lctp_stage *st, const uint64_t *worker_ixs, const uint64_t *worker_off,
                    size_t n_workers, uint64_t *worker_rng, double *lik_mean, double *lik_var, double *liks,
                    uint64_t *counts_off, uint16_t *counts, uint64_t counts_cap, uint64_t *n_alns_out,
                    uint64_t *iters_out, bool device_only) {
    lctp_ctx *ctx = h->ctx;
    cudaStream_t s = ctx->stream;
    const LocusDev &L = h->dev;
    const double t_enter = dbg_now();
    if (!st || !worker_ixs || !worker_off || !worker_rng || (!device_only && (!lik_mean || !lik_var)) || n_workers == 0) {
        set_error("lctp_solve_stage: NULL argument");
        return LCTP_E_INVALID;
    }
    if (st->kind > 1 || st->attempts == 0 || st->attempts > 65535) {
        set_error("lctp_solve_stage: invalid stage (kind=%u attempts=%u)", st->kind, st->attempts);
        return LCTP_E_INVALID;
    }
    if (st->kind == 0 && (st->sample_size == 0 || st->sample_size > 0x7FFFFFFFull)) {
        set_error("lctp_solve_stage: invalid greedy sample size %llu", (unsigned long long)st->sample_size);
        return LCTP_E_INVALID;
    }
    if (st->kind == 1 && (!(st->init_prob > 0.0 && st->init_prob <= 1.0) || st->anneal_steps == 0)) {
        set_error("lctp_solve_stage: invalid annealing parameters");
        return LCTP_E_INVALID;
    }
    if (st->plato_size >= 0x40000000ull) {
        set_error("lctp_solve_stage: plateau size %llu too large (device counters are 32-bit)", (unsigned long long)st->plato_size);
        return LCTP_E_CAPACITY;
    }
    if (L.R > 65535u) {
        set_error("lctp_solve_stage: %u reads exceed the 65535 the shared-memory read tables can index", L.R);
        return LCTP_E_CAPACITY;
    }
    if (h->max_run > MAX_RUN) {
        set_error("lctp_solve_stage: %u pair alignments of one read on one contig exceed %u (the reference keeps at "
                  "most 10, MAX_USED_ALNS, src/model/locs.rs:741)", h->max_run, MAX_RUN);
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
    // a genotype with more than 65535 candidates is refused by the kernel (u16 offsets, like the reference's
    // assert at src/model/assgn.rs:58 per read); do not reserve more than that
    const uint32_t cap = (uint32_t)std::min<uint64_t>(cap64, 65535);

    StageParams P;
    P.kind = st->kind; P.attempts = st->attempts; P.best_start = st->best_start; P.sample_size = (uint32_t)st->sample_size;
    P.plato_size = st->plato_size; P.anneal_steps = st->anneal_steps;
    P.max_iter = std::max<uint64_t>(100000, st->plato_size * 100);
    P.ln_init_prob = st->kind == 1 ? std::log(st->init_prob) : 0.0;
    P.n_workers = (uint32_t)n_workers; P.cap = cap; P.Wmax = 2 + p * h->max_n_windows;
    const bool want_counts = counts != nullptr && counts_off != nullptr;
    P.want_counts = want_counts ? 1 : 0;
    P.narrow_w = P.Wmax <= 4096 ? 1 : 0;                       // 32-bit candidate records (12-bit windows)
    if (env_int("LCTP_WIDE_WINDOWS", 0)) P.narrow_w = 0;       // test knob: exercise the 64-bit candidate records
    // The per-read index (candidate offsets + the list of non-trivial reads, 4 bytes per read) stays in shared memory
    // while 16 workers per SM still fit (the register file allows no more); for larger R it moves to the worker's
    // slab ("slab mode": 10 bytes per read there) and shared memory keeps the window state and the assignments only --
    // at the KIR-scale shape 8 workers per SM instead of 4.  The greedy loop fetches its entries a round ahead.
    {
        const size_t per_sm = ctx->smem_optin + 1024;            // opt-in limit per CTA = SM capacity - 1 KB
        const size_t with_nt = group_smem_bytes(P.Wmax, L.R, false, st->kind == 1) + 1024;
        (void)per_sm; (void)with_nt;
        P.nt_global = 1;   // measured at the C2 shape too: 16 workers per SM instead of 14, stage kernel 10.8 -> 10.1 ms
        // test / tuning knob, annealing only: the greedy pipeline fetches its index entries from the slab a round ahead
        if (const char *e = getenv("LCTP_NT_GLOBAL")) { if (st->kind == 1) P.nt_global = atoi(e) ? 1 : 0; }
    }
    P.slab_bytes = slab_bytes_for(cap, P.narrow_w == 0, P.nt_global ? L.R : 0);
    // samples of more than 11 reads (greedy_solve_big): sample + swap partners + index vector behind the slab
    P.big_bytes = (st->kind == 0 && st->sample_size > (uint64_t)MAX_SAMPLE)
                      ? align_up(((size_t)2 * std::min<uint64_t>(st->sample_size, L.R) + L.R) * 4, 128) : 0;
    P.slab_bytes += P.big_bytes;

    int rc;
    if ((rc = ensure_jump_tables(ctx))) return rc;
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
    ctx->stats.h2d_bytes += n * 8 + (n_workers + 1) * 8 + n_workers * 32 + n * p * 4;

    const double t_launch = dbg_now();
    if (ctx->dbg_req && ctx->dbg_req->win_weight && ctx->dbg_req->wmax < P.Wmax) {
        set_error("lctp_solve_stage_dbg: wmax %u < %u windows of a genotype", ctx->dbg_req->wmax, P.Wmax);
        return LCTP_E_INVALID;
    }
    rc = launch_stage_kernel(h, P, n_workers, n, want_counts);
    if (rc) return rc;
    if (device_only) {
        ctx->stats.stage_launches += 1;
        ctx->stats.stage_genotypes += n;
        ctx->stats.stage_attempts += n * st->attempts;
        return LCTP_OK;
    }

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
    ctx->stats.d2h_bytes += n * 8 * 4 + n_workers * 32 + (liks ? n * st->attempts * 8 : 0) + sizeof(int);
    if (flags[0] & 4) {
        set_error("lctp_solve_stage: greedy sample size %llu on a genotype with so many non-trivial reads that rand's "
                  "index::sample would use its rejection branch (amount >= 163 and length >= 270 * amount), which is not "
                  "implemented", (unsigned long long)st->sample_size);
        return LCTP_E_CAPACITY;
    }
    if (flags[0]) {
        set_error("lctp_solve_stage: candidate overflow (more than %u candidates in one genotype)", cap);
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
    if (const lctp_stage_debug *d = ctx->dbg_req) {
        const size_t na = n * st->attempts;
        std::vector<double> lk(2 * na);
        std::vector<uint32_t> ct(2 * na);
        LCTP_CUDA_CHECK(cudaMemcpyAsync(lk.data(), ctx->d_dbg_lik.p, 2 * na * 8, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(ct.data(), ctx->d_dbg_cnt.p, 2 * na * 4, cudaMemcpyDeviceToHost, s));
        if (d->win_weight && d->win_depth && d->win_lik) {
            LCTP_CUDA_CHECK(cudaMemcpyAsync(d->win_weight, ctx->d_dbg_ww.p, na * d->wmax * 8, cudaMemcpyDeviceToHost, s));
            LCTP_CUDA_CHECK(cudaMemcpyAsync(d->win_lik, ctx->d_dbg_wl.p, na * d->wmax * 8, cudaMemcpyDeviceToHost, s));
            LCTP_CUDA_CHECK(cudaMemcpyAsync(d->win_depth, ctx->d_dbg_wd.p, na * d->wmax * 4, cudaMemcpyDeviceToHost, s));
        }
        LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
        for (size_t q = 0; q < na; q++) {
            if (d->aln_lik) d->aln_lik[q] = lk[2 * q];
            if (d->depth_lik) d->depth_lik[q] = lk[2 * q + 1];
            if (d->unmapped) d->unmapped[q] = ct[2 * q];
            if (d->out_of_bounds) d->out_of_bounds[q] = ct[2 * q + 1];
        }
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
        ctx->stats.d2h_bytes += off * 2;
        LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    return LCTP_OK;
}

}  // namespace lctp

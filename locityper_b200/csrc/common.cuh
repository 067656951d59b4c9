// common.cuh -- shared definitions of the lctp library (sm_100a only; no CPU fallback).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <vector>

#include "../../include/lctp.h"

namespace lctp {

void set_error(const char *fmt, ...);

#define LCTP_CUDA_CHECK(expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::lctp::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,    \
                              __LINE__, cudaGetErrorString(_e));                               \
            return LCTP_E_CUDA;                                                                \
        }                                                                                      \
    } while (0)

// Device buffer from the stream-ordered pool (cudaMallocAsync): after warm-up, allocating and freeing
// the per-locus buffers costs no driver round trip, which matters for the end-to-end (upload + solve) path.
cudaStream_t current_alloc_stream();
void set_alloc_stream(cudaStream_t s);

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    int alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return LCTP_OK;
        s = current_alloc_stream();
        LCTP_CUDA_CHECK(cudaMallocAsync((void **)&p, count * sizeof(T), s));
        return LCTP_OK;
    }
    int ensure(size_t count) { return count <= n ? LCTP_OK : alloc(count); }
    void take(DevBuf &o) {              // move: this buffer becomes o's allocation, o becomes empty
        release();
        p = o.p; n = o.n; s = o.s;
        o.p = nullptr; o.n = 0;
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

// Pinned host staging buffer.
template <typename T>
struct PinBuf {
    T *p = nullptr;
    size_t n = 0;
    int ensure(size_t count) {
        if (count <= n) return LCTP_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
        LCTP_CUDA_CHECK(cudaMallocHost((void **)&p, count * sizeof(T)));
        n = count;
        return LCTP_OK;
    }
    ~PinBuf() {
        if (p) cudaFreeHost(p);
    }
    PinBuf() = default;
    PinBuf(const PinBuf &) = delete;
    PinBuf &operator=(const PinBuf &) = delete;
};

// Work plan of the balanced prefilter kernel (prefilter.cu): one region per persistent CTA and round.
static constexpr int BAL_MAXW = 16;            // warps per CTA
struct BalWarp { uint32_t row0, col0, ncols, c, a_off, b_off; };       // c == 0: idle slot
struct BalSeg { uint32_t src, dst, len; };     // one contiguous run of matrix columns -> offset in the staged read row (doubles)
struct BalRegion {
    uint32_t row_chunks;                       // 16-byte chunks staged per read
    uint32_t tab_off;                          // first entry of this region in the source-column table
    uint32_t n_seg;                            // the same staging as contiguous segments (bulk-copy producer)
    uint32_t n_active;                         // warps with work
    BalSeg seg[2 * BAL_MAXW];
    BalWarp warp[BAL_MAXW];
};
struct BalPlan {
    std::vector<BalRegion> regions;
    std::vector<uint32_t> tab, pattern;
    uint32_t row_len = 0, n_warps = 0, cmax = 0, load = 0;
};

}  // namespace lctp

struct lctp_dist;
struct lctp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    size_t smem_optin = 0;
    uint32_t max_resident_workers = 0;
    uint64_t launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // prefilter start/stop, stage start/stop
    lctp_stats stats = {};
    // stage scratch, grown on demand and reused across loci / stages
    lctp::DevBuf<unsigned char> scratch;
    lctp::DevBuf<uint64_t> d_worker_ixs, d_worker_off, d_rng;
    lctp::DevBuf<uint32_t> d_tuples;
    lctp::DevBuf<double> d_lik_mean, d_lik_var, d_liks;
    lctp::DevBuf<uint64_t> d_nalns, d_iters, d_rng_mats;
    lctp::DevBuf<uint16_t> d_counts;
    lctp::DevBuf<int> d_flags;
    lctp::PinBuf<unsigned char> pin;
    // debug outputs of the next stage launch (lctp_solve_stage_dbg sets it around its launch) and the debug sink
    const lctp_stage_debug *dbg_req = nullptr;
    lctp::DevBuf<double> d_dbg_lik, d_dbg_ww, d_dbg_wl;
    lctp::DevBuf<uint32_t> d_dbg_cnt, d_dbg_wd;
    int dbg_level = 0;
    std::string dbg_dir;
    std::vector<std::string> dbg_names;
    FILE *dbg_sol = nullptr, *dbg_sol_ext = nullptr, *dbg_depth = nullptr;
    lctp_dist *local_sel = nullptr;      // one-rank candidate selector of the single-GPU prefilter (dist.cu)
    // Per-genotype-id arrays of a whole solve (lik_mean / lik_var / attempts by id, the id list): kept between solves in
    // their "nothing solved" state (NaN / NaN / 0) and cleaned entry by entry after use, so that a KIR-scale solve (G =
    // 500,500) does not allocate and fill 13 MB of host memory every time (~2 ms of its 24 ms on 8 GPUs).
    std::vector<double> id_mean, id_var;
    std::vector<uint16_t> id_attempts;
    std::vector<uint64_t> id_list;
};

namespace lctp {
// the per-id arrays sized for G genotypes, in their clean state
inline void per_id_reserve(lctp_ctx *ctx, uint64_t G) {
    if (ctx->id_mean.size() < G) {
        const double nan = __builtin_nan("");
        ctx->id_mean.assign(G, nan); ctx->id_var.assign(G, nan); ctx->id_attempts.assign(G, 0);
        ctx->id_list.resize(G);
    }
}
inline void per_id_clean(lctp_ctx *ctx, const std::vector<uint64_t> &touched) {
    const double nan = __builtin_nan("");
    for (uint64_t g : touched) { ctx->id_mean[g] = nan; ctx->id_var[g] = nan; ctx->id_attempts[g] = 0; }
}
}  // namespace lctp

// Device-side view of one uploaded locus (all pointers are device pointers).
struct LocusDev {
    uint32_t H, R, p, Hpad;
    uint64_t G;
    uint32_t window, left_padding, tweak, depth_k;
    double prob_diff, depth_contrib, aln_contrib, rel_contrib, min_weight;
    uint64_t c_window, c_span;   // ceil(2^64 / window), ceil(2^64 / (2 tweak + 1)): division / remainder by multiplication
    const double *Mt;            // [R][Hpad] best-alignment matrix, read-major (a1)
    const double *unmapped;      // [R]
    const uint32_t *cm_off;      // [H*R+1] contig-major CSR of pair alignments
    const double *cm_lnprob;     // [NPA + R]: ln_prob of every pair alignment, then unmapped[r] at NPA + r, so that the
                                 //            ln-probability of any candidate is ONE index into this array
    uint32_t npa;
    const uint2 *cm_mid;         // [NPA] (middle1, middle2)
    const uint32_t *hap_len, *hap_n_windows, *hap_reg_start;   // [H]
    const uint64_t *hap_pos_off; // [H+1]
    const double *pos_weight;
    const uint8_t *pos_gc;
    const double *depth_table;   // [101][depth_k]
    const uint32_t *gt_tuples;   // [G*p] or nullptr
    const double *priors;        // [G] or nullptr
};

struct lctp_locus_h {
    lctp_ctx *ctx = nullptr;
    lctp_locus host;             // scalar copy (pointers NOT retained: nulled after upload)
    LocusDev dev;
    uint64_t npa = 0;
    uint32_t max_hap_alns = 0;   // max over haplotypes of #pair alignments on that haplotype
    uint32_t max_n_windows = 0;
    uint32_t max_run = 0;        // longest (read, contig) run of pair alignments
    std::vector<uint32_t> hap_alns;      // [H] #pair alignments per haplotype
    std::vector<uint32_t> gt_tuples_host; // explicit genotype list (host copy), empty = full enumeration
    std::vector<double> priors_host;      // host copy of priors, empty = 0.0
    std::vector<double> unmapped_host;    // [R] host copy (count_unexplained_reads)
    std::vector<uint32_t> hap_n_windows; // [H] host copy
    lctp::DevBuf<double> Mt, unmapped, cm_lnprob, pos_weight, depth_table, priors, scores;
    lctp::DevBuf<uint32_t> cm_off, hap_len, hap_nw, hap_rs, gt_tuples;
    lctp::DevBuf<uint2> cm_mid;
    lctp::DevBuf<uint64_t> hap_pos_off;
    lctp::DevBuf<uint8_t> pos_gc;
    bool scores_valid = false;
    lctp::BalPlan pf_plan;               // balanced prefilter plan of the last (range, pattern), see prefilter.cu
    lctp::DevBuf<lctp::BalRegion> pf_regions;
    lctp::DevBuf<uint32_t> pf_tab;
    bool pf_plan_valid = false, pf_plan_auto = false;
    uint64_t pf_g_begin = 0, pf_g_end = 0;
    bool mt_nonpositive = false;         // every matrix entry <= +0.0 (integer-pipe max is valid)
};

// Pair alignments of a locus left resident on the device by lctp_pair_alignments_dev: the pa_* / unmapped_prob section of
// lctp_locus without a host round trip (consumed by lctp_locus_upload_pairs).
struct lctp_pairs_h {
    lctp_ctx *ctx = nullptr;
    uint32_t n_reads = 0, n_haps = 0;
    uint64_t n_pairs = 0;
    lctp::DevBuf<uint64_t> pa_off;
    lctp::DevBuf<uint32_t> contig, mid1, mid2;
    lctp::DevBuf<double> lnprob, unmapped;
};

// Device-resident input of the pairing (lctp_group_reads_dev -> lctp_pair_alignments_from): the lctp_mates arrays of the
// reads that passed, in consumption order, plus the per-read max_alns.
struct lctp_mates_h {
    lctp_ctx *ctx = nullptr;
    uint32_t n_reads = 0;
    uint64_t n = 0;
    lctp::DevBuf<uint64_t> ma_off;
    lctp::DevBuf<uint32_t> contig, start, end, rec;
    lctp::DevBuf<uint8_t> flags, max_alns;
    lctp::DevBuf<double> lnprob;
};

namespace lctp {
// upload.cu
int upload_locus(lctp_ctx *ctx, const lctp_locus *in, lctp_locus_h *h, const lctp_pairs_h *pairs);
// prefilter.cu
int launch_prefilter(lctp_locus_h *h, uint64_t g_begin, uint64_t g_end, double *d_scores);
int measure_fp64_rate(lctp_ctx *ctx, double *lane_inst_per_s);
int prefilter_plan_check(uint32_t H, uint32_t n_sm, const uint32_t *pattern, uint32_t n_pattern, uint64_t g_begin,
                         uint64_t g_end, uint32_t *n_regions, uint32_t *load, uint32_t *pattern_out);
// pairs.cu
int pair_alignments_dev(lctp_ctx *ctx, const lctp_mates *in, lctp_pairs_h *out, const lctp_mates_h *src = nullptr);
int pairs_fetch(lctp_pairs_h *p, uint64_t cap, uint64_t *pa_off, uint32_t *pa_contig, double *pa_ln_prob,
                uint32_t *pa_mid1, uint32_t *pa_mid2, double *unmapped_prob);
// rescore.cu
int rescore_alignments(lctp_ctx *ctx, const lctp_alns *in, double *ln_prob, uint32_t *edit, uint32_t *read_len,
                       uint8_t *save);
int collect_read_ends(lctp_ctx *ctx, const lctp_read_ends *in, double *ln_prob, uint32_t *edit, uint32_t *read_len,
                      uint8_t *ok, uint32_t *best_edit, double *weight_factor, uint32_t *thr_dist, uint32_t *pass_dist,
                      uint32_t *n_kept, uint32_t *kept_rec);
// solver.cu
int launch_stage(lctp_locus_h *h, const lctp_stage *st, const uint64_t *worker_ixs,
                 const uint64_t *worker_off, size_t n_workers, uint64_t *worker_rng,
                 double *lik_mean, double *lik_var, double *liks, uint64_t *counts_off,
                 uint16_t *counts, uint64_t counts_cap, uint64_t *n_alns_out, uint64_t *iters_out);
int launch_stage_ex(lctp_locus_h *h, const lctp_stage *st, const uint64_t *worker_ixs,
                    const uint64_t *worker_off, size_t n_workers, uint64_t *worker_rng,
                    double *lik_mean, double *lik_var, double *liks, uint64_t *counts_off,
                    uint16_t *counts, uint64_t counts_cap, uint64_t *n_alns_out, uint64_t *iters_out, bool device_only);
// dist.cu
lctp_dist *local_selector(lctp_ctx *ctx);
void free_local_selector(lctp_ctx *ctx);
// host_solve.cpp
void genotype_tuple(uint32_t H, uint32_t p, const uint32_t *gt_tuples, uint64_t g, uint32_t *out);
}  // namespace lctp

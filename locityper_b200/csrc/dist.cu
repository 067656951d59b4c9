// dist.cu -- one locus sharded over the GPUs of a box, inside the library.
//
// Genotypes are independent units of both phases of solve::solve (src/solvers/solve.rs:105-119, :1116-1142):
//   prefilter : rank r scores the contiguous id range shard_range(G, r, N) (a2), selects ON THE DEVICE the candidates
//               that can survive truncate_ixs (src/solvers/solve.rs:52-84) and the ranks exchange fixed-capacity
//               buffers of (score f64, genotype id u64) + {count, overflow} with ONE ncclAllGather; every rank then
//               runs the exact truncate_ixs on the union -> identical, identically ordered survivors everywhere;
//   stage     : MainWorker::run (:1049-1063) shuffles and partitions with the same locus stream on every rank;
//               rank r solves the logical workers w = r (mod N) and ONE ncclAllGather carries (lik_mean, lik_var) of
//               its genotypes, its workers' RNG states and its overflow flag.
// Pruning and the final result are small, replicated host work.  The volumes are KBs..MBs per locus, so the cost is
// collective latency: there is nothing for a fused compute+collective kernel to overlap with.
//
// NCCL is loaded with dlopen("libnccl.so.2") at lctp_dist_init: no link-time dependency for single-GPU users, and a
// host process that already has NCCL loaded (PyTorch) shares that copy.
#include "common.cuh"

#include <nccl.h>
#include <dlfcn.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <numeric>

namespace lctp {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // a copy the process has already loaded (PyTorch's) first; otherwise the system library, kept local
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (api.handle) break; }
        if (!api.handle)
            for (const char *n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (api.handle) break; }
        if (!api.handle) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))dlsym(api.handle, "ncclGetVersion");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) {
            dlclose(api.handle);
            api.handle = nullptr;
        }
    });
    return api.handle ? &api : nullptr;
}

#define LCTP_NCCL_CHECK(expr)                                                                       \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess) {                                                                    \
            ::lctp::set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, nccl_api()->GetErrorString(_r)); \
            return LCTP_E_CUDA;                                                                     \
        }                                                                                           \
    } while (0)

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Monotone u64 key of an f64 score: larger score = larger key; -0.0 and +0.0 share a key (they compare equal).
__host__ __device__ inline uint64_t score_key(double v) {
    if (v == 0.0) v = 0.0;
    uint64_t b;
#ifdef __CUDA_ARCH__
    b = (uint64_t)__double_as_longlong(v);
#else
    std::memcpy(&b, &v, 8);
#endif
    return b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
}
__host__ __device__ inline double key_score(uint64_t k) {
    const uint64_t b = (k >> 63) ? (k ^ 0x8000000000000000ull) : ~k;
    double v;
#ifdef __CUDA_ARCH__
    v = __longlong_as_double((long long)b);
#else
    std::memcpy(&v, &b, 8);
#endif
    return v;
}

__global__ void k_dist_keys(const double *__restrict__ scores, uint64_t g0, uint32_t n, uint64_t *__restrict__ keys,
                            uint32_t *__restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = score_key(scores[g0 + i]);
    vals[i] = i;
}

// After the descending sort: the candidate set of the shard is a prefix.  Every id whose score is
// >= min(local_best - filt_diff, K-th best local score), K = max(min_size, threads), can survive the global
// truncate_ixs (proof: the global threshold best_global - filt_diff >= best_local - filt_diff; the global
// "at least min_size, ties at the cut included" rule keeps scores >= the global min_size-th best >= the local K-th
// best; raising to `threads` takes the `threads` best overall, which are among the local `threads` best).
// One thread: two binary searches, then the header; the payload copy is a second tiny kernel.
__global__ void k_dist_select(const uint64_t *__restrict__ keys, uint32_t n, double filt_diff, uint32_t K, uint32_t cap,
                              uint64_t *__restrict__ header) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint64_t n_cand = 0;
    if (n > 0) {
        const double best = key_score(keys[0]);
        uint64_t tkey = score_key(best - filt_diff);
        if (K >= n) tkey = 0;
        else tkey = min(tkey, keys[K - 1]);
        // keys are descending: count of keys >= tkey
        uint32_t lo = 0, hi = n;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (keys[mid] >= tkey) lo = mid + 1; else hi = mid; }
        n_cand = lo;
    }
    header[0] = n_cand;
    header[1] = n_cand > cap ? 1 : 0;
}
__global__ void k_dist_payload(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t g0,
                               const uint64_t *__restrict__ header, uint32_t cap, double *__restrict__ out_score,
                               uint64_t *__restrict__ out_id) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n = min(header[0], (uint64_t)cap);
    if (i >= n) return;
    out_score[i] = key_score(keys[i]);
    out_id[i] = g0 + vals[i];
}

}  // namespace lctp

struct lctp_dist {
    lctp_ctx *ctx = nullptr;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    lctp_dist_timing timing = {};
    lctp::DevBuf<uint64_t> keys_in, keys_out;
    lctp::DevBuf<uint32_t> vals_in, vals_out;
    lctp::DevBuf<unsigned char> sort_tmp, send, recv;
    lctp::PinBuf<unsigned char> host_recv;
};

namespace lctp {

static void shard_range(uint64_t n, int rank, int world, uint64_t *b, uint64_t *e) {
    const uint64_t base = n / world, rem = n % world;
    *b = rank * base + std::min<uint64_t>(rank, rem);
    *e = *b + base + ((uint64_t)rank < rem ? 1 : 0);
}

// ncclAllGather of `bytes` per rank from d->send into d->recv, then D2H of the whole receive buffer.
static int gather_to_host(lctp_dist *d, size_t bytes) {
    lctp_ctx *ctx = d->ctx;
    cudaStream_t s = ctx->stream;
    int rc;
    if ((rc = d->recv.ensure(bytes * d->world))) return rc;
    if ((rc = d->host_recv.ensure(bytes * d->world))) return rc;
    if (!d->comm) {        // the context's own one-rank selector (lctp::local_selector): no communicator, nothing to gather
        LCTP_CUDA_CHECK(cudaMemcpyAsync(d->host_recv.p, d->send.p, bytes, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
        ctx->stats.d2h_bytes += bytes;
        return LCTP_OK;
    }
    LCTP_CUDA_CHECK(cudaEventRecord(d->ev[0], s));
    LCTP_NCCL_CHECK(nccl_api()->AllGather(d->send.p, d->recv.p, bytes, ncclChar, d->comm, s));
    LCTP_CUDA_CHECK(cudaEventRecord(d->ev[1], s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(d->host_recv.p, d->recv.p, bytes * d->world, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
    float ms = 0.f;
    LCTP_CUDA_CHECK(cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]));
    d->timing.collective_ms += ms;
    d->timing.collectives += 1;
    d->timing.gathered_bytes += bytes * d->world;
    ctx->stats.d2h_bytes += bytes * d->world;
    return LCTP_OK;
}

// truncate_ixs (src/solvers/solve.rs:52-84) on the union of the ranks' candidates: (id, score) pairs, n_total = G.
static size_t truncate_union(std::vector<std::pair<uint64_t, double>> &u, uint64_t n_total, double filt_diff,
                             size_t min_size, size_t threads, uint64_t *out) {
    // score descending, ties by ascending id (the order the single-GPU path produces: a stable sort of 0..G)
    std::sort(u.begin(), u.end(), [](const std::pair<uint64_t, double> &a, const std::pair<uint64_t, double> &b) {
        const uint64_t ka = score_key(a.second), kb = score_key(b.second);
        return ka != kb ? ka > kb : a.first < b.first;
    });
    const size_t n = u.size();
    size_t m = n;
    if (n && !(min_size >= n_total)) {
        double thresh = u[0].second - filt_diff;
        auto count_ge = [&](double th) {
            size_t lo = 0, hi = n;
            while (lo < hi) { const size_t mid = (lo + hi) / 2; if (u[mid].second >= th) lo = mid + 1; else hi = mid; }
            return lo;
        };
        m = count_ge(thresh);
        if (m < n || n == n_total) {                   // (m == n < G: the candidates are exactly the survivors)
            if (m < min_size && min_size <= n) { thresh = u[min_size - 1].second; m = count_ge(thresh); }
            m = std::min(std::max(m, threads), n);
        }
    }
    for (size_t q = 0; q < m; q++) out[q] = u[q].first;
    return m;
}

}  // namespace lctp

using namespace lctp;

extern "C" {

int lctp_dist_unique_id(uint8_t id[LCTP_DIST_ID_BYTES]) {
    if (!id) { set_error("lctp_dist_unique_id: NULL argument"); return LCTP_E_INVALID; }
    static_assert(sizeof(ncclUniqueId) == LCTP_DIST_ID_BYTES, "ncclUniqueId size");
    NcclApi *api = nccl_api();
    if (!api) { set_error("lctp_dist: libnccl.so.2 could not be loaded (%s)", dlerror()); return LCTP_E_CUDA; }
    ncclUniqueId u;
    LCTP_NCCL_CHECK(api->GetUniqueId(&u));
    std::memcpy(id, &u, sizeof(u));
    return LCTP_OK;
}

int lctp_dist_init(lctp_ctx *ctx, const uint8_t id[LCTP_DIST_ID_BYTES], int rank, int world, lctp_dist **out) {
    if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) {
        set_error("lctp_dist_init: invalid argument");
        return LCTP_E_INVALID;
    }
    NcclApi *api = nccl_api();
    if (!api) { set_error("lctp_dist: libnccl.so.2 could not be loaded"); return LCTP_E_CUDA; }
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    lctp_dist *d = new lctp_dist();
    d->ctx = ctx; d->rank = rank; d->world = world;
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    ncclResult_t r = api->CommInitRank(&d->comm, world, u, rank);
    if (r != ncclSuccess) {
        set_error("lctp_dist_init: ncclCommInitRank failed: %s", api->GetErrorString(r));
        delete d;
        return LCTP_E_CUDA;
    }
    cudaEventCreate(&d->ev[0]);
    cudaEventCreate(&d->ev[1]);
    *out = d;
    return LCTP_OK;
}

}  // extern "C"

// The single-GPU prefilter uses the same device-side candidate selection (sort of the scores on the device, the candidate set
// is a prefix, a few thousand (score, id) pairs cross to the host instead of all G scores): a one-rank selector without a
// communicator, owned by the context.
namespace lctp {
lctp_dist *local_selector(lctp_ctx *ctx) {
    if (!ctx->local_sel) {
        lctp_dist *d = new lctp_dist();
        d->ctx = ctx; d->rank = 0; d->world = 1; d->comm = nullptr;
        cudaEventCreate(&d->ev[0]);
        cudaEventCreate(&d->ev[1]);
        ctx->local_sel = d;
    }
    return ctx->local_sel;
}
void free_local_selector(lctp_ctx *ctx) {
    if (ctx->local_sel) { lctp_dist_destroy(ctx->local_sel); ctx->local_sel = nullptr; }
}
}  // namespace lctp

extern "C" {

void lctp_dist_destroy(lctp_dist *d) {
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    set_alloc_stream(d->ctx->stream);
    if (d->comm) nccl_api()->CommDestroy(d->comm);
    cudaEventDestroy(d->ev[0]);
    cudaEventDestroy(d->ev[1]);
    delete d;
}

int lctp_dist_rank(const lctp_dist *d) { return d ? d->rank : -1; }
int lctp_dist_world(const lctp_dist *d) { return d ? d->world : 0; }

int lctp_dist_get_timing(lctp_dist *d, lctp_dist_timing *out, int reset) {
    if (!d || !out) { set_error("lctp_dist_get_timing: NULL argument"); return LCTP_E_INVALID; }
    *out = d->timing;
    if (reset) d->timing = lctp_dist_timing{};
    return LCTP_OK;
}

int lctp_dist_prefilter(lctp_dist *d, lctp_locus_h *h, size_t min_size, size_t threads, uint64_t *ixs_out,
                        size_t cap_out, size_t *n_out) {
    if (!d || !h || !ixs_out || !n_out) { set_error("lctp_dist_prefilter: NULL argument"); return LCTP_E_INVALID; }
    if (h->ctx != d->ctx) { set_error("lctp_dist_prefilter: locus belongs to another context"); return LCTP_E_INVALID; }
    lctp_ctx *ctx = d->ctx;
    cudaStream_t s = ctx->stream;
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    set_alloc_stream(s);
    const uint64_t G = h->dev.G;
    uint64_t g0, g1;
    shard_range(G, d->rank, d->world, &g0, &g1);
    const uint32_t n = (uint32_t)(g1 - g0);
    if (g1 - g0 > 0xFFFFFFF0ull) { set_error("lctp_dist_prefilter: shard too large"); return LCTP_E_CAPACITY; }
    int rc;
    // a2 on the shard (timed like the single-GPU prefilter)
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[0], s));
    if (n) { if ((rc = launch_prefilter(h, g0, g1, h->scores.p))) return rc; }
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[1], s));
    // candidates on the device: sort the shard by score, the candidate set is a prefix
    const uint32_t K = (uint32_t)std::min<size_t>(std::max<size_t>(std::max(min_size, threads), 1), 0xFFFFFFF0u);
    if ((rc = d->keys_in.ensure(std::max<uint32_t>(n, 1)))) return rc;
    if ((rc = d->keys_out.ensure(std::max<uint32_t>(n, 1)))) return rc;
    if ((rc = d->vals_in.ensure(std::max<uint32_t>(n, 1)))) return rc;
    if ((rc = d->vals_out.ensure(std::max<uint32_t>(n, 1)))) return rc;
    if (n) {
        k_dist_keys<<<(n + 255) / 256, 256, 0, s>>>(h->scores.p, g0, n, d->keys_in.p, d->vals_in.p);
        size_t tmp = 0;
        LCTP_CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, d->keys_in.p, d->keys_out.p, d->vals_in.p,
                                                                  d->vals_out.p, (int)n, 0, 64, s));
        if ((rc = d->sort_tmp.ensure(tmp))) return rc;
        LCTP_CUDA_CHECK(cub::DeviceRadixSort::SortPairsDescending(d->sort_tmp.p, tmp, d->keys_in.p, d->keys_out.p,
                                                                  d->vals_in.p, d->vals_out.p, (int)n, 0, 64, s));
        ctx->launches += 2;
    }
    // fixed-capacity exchange; a second, larger one only if some rank's candidates did not fit
    uint32_t cap = (uint32_t)std::min<uint64_t>(K, (G + d->world - 1) / d->world);
    if (cap == 0) cap = 1;
    std::vector<std::pair<uint64_t, double>> uni;
    for (int round = 0; round < 2; round++) {
        const size_t bytes = 16 + (size_t)cap * 16;            // {count, overflow} + cap scores + cap ids
        if ((rc = d->send.ensure(bytes))) return rc;
        uint64_t *hdr = (uint64_t *)d->send.p;
        double *sc = (double *)(d->send.p + 16);
        uint64_t *ids = (uint64_t *)(d->send.p + 16 + (size_t)cap * 8);
        k_dist_select<<<1, 32, 0, s>>>(d->keys_out.p, n, h->host.filt_diff, K, cap, hdr);
        k_dist_payload<<<(cap + 255) / 256, 256, 0, s>>>(d->keys_out.p, d->vals_out.p, g0, hdr, cap, sc, ids);
        ctx->launches += 2;
        LCTP_CUDA_CHECK(cudaGetLastError());
        if ((rc = gather_to_host(d, bytes))) return rc;
        uint64_t max_cand = 0;
        bool overflow = false;
        for (int r = 0; r < d->world; r++) {
            const uint64_t *rh = (const uint64_t *)(d->host_recv.p + (size_t)r * bytes);
            max_cand = std::max(max_cand, rh[0]);
            overflow |= rh[1] != 0;
        }
        if (overflow && round == 0) { cap = (uint32_t)max_cand; d->timing.overflow_rounds += 1; continue; }
        if (overflow) { set_error("lctp_dist_prefilter: candidate exchange overflowed twice"); return LCTP_E_CAPACITY; }
        uni.clear();
        for (int r = 0; r < d->world; r++) {
            const unsigned char *base = d->host_recv.p + (size_t)r * bytes;
            const uint64_t cnt = ((const uint64_t *)base)[0];
            const double *rs = (const double *)(base + 16);
            const uint64_t *ri = (const uint64_t *)(base + 16 + (size_t)cap * 8);
            for (uint64_t q = 0; q < cnt; q++) uni.emplace_back(ri[q], rs[q]);
        }
        break;
    }
    {
        float ms = 0.f;
        LCTP_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        ctx->stats.prefilter_ms += ms;
        ctx->stats.prefilter_launches += 1;
        ctx->stats.prefilter_genotypes += n;
        d->timing.kernel_ms += ms;
    }
    if (uni.size() > cap_out) { set_error("lctp_dist_prefilter: output buffer too small"); return LCTP_E_CAPACITY; }
    *n_out = truncate_union(uni, G, h->host.filt_diff, min_size, threads, ixs_out);
    return LCTP_OK;
}

int lctp_dist_solve_stage(lctp_dist *d, lctp_locus_h *h, const lctp_stage *st, const uint64_t *worker_ixs,
                          const uint64_t *worker_off, size_t n_workers, uint64_t *worker_rng, double *lik_mean,
                          double *lik_var) {
    if (!d || !h || !st || !worker_ixs || !worker_off || !worker_rng || !lik_mean || !lik_var || n_workers == 0) {
        set_error("lctp_dist_solve_stage: NULL argument");
        return LCTP_E_INVALID;
    }
    if (h->ctx != d->ctx) { set_error("lctp_dist_solve_stage: locus belongs to another context"); return LCTP_E_INVALID; }
    lctp_ctx *ctx = d->ctx;
    cudaStream_t s = ctx->stream;
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    set_alloc_stream(s);
    const int N = d->world;
    // every rank's share: workers w = r (mod N), their positions in worker order
    std::vector<size_t> cnt_pos(N, 0), cnt_w(N, 0);
    for (size_t w = 0; w < n_workers; w++) { cnt_pos[w % N] += worker_off[w + 1] - worker_off[w]; cnt_w[w % N] += 1; }
    const size_t cap_pos = *std::max_element(cnt_pos.begin(), cnt_pos.end());
    const size_t cap_w = *std::max_element(cnt_w.begin(), cnt_w.end());
    // this rank's workers, packed
    std::vector<uint64_t> l_ixs, l_off(1, 0), l_rng;
    for (size_t w = d->rank; w < n_workers; w += N) {
        l_ixs.insert(l_ixs.end(), worker_ixs + worker_off[w], worker_ixs + worker_off[w + 1]);
        l_off.push_back(l_ixs.size());
        l_rng.insert(l_rng.end(), worker_rng + 4 * w, worker_rng + 4 * w + 4);
    }
    const size_t n_loc = l_ixs.size(), w_loc = l_off.size() - 1;
    int rc;
    // payload per rank: [flag u64][cap_pos means][cap_pos variances][cap_w x 4 rng words]
    const size_t bytes = 8 + cap_pos * 16 + cap_w * 32;
    if ((rc = d->send.ensure(bytes))) return rc;
    LCTP_CUDA_CHECK(cudaMemsetAsync(d->send.p, 0, 8, s));
    if (w_loc) {
        rc = launch_stage_ex(h, st, l_ixs.data(), l_off.data(), w_loc, l_rng.data(), nullptr, nullptr, nullptr, nullptr,
                             nullptr, 0, nullptr, nullptr, true);
        if (rc) return rc;
        LCTP_CUDA_CHECK(cudaMemcpyAsync(d->send.p, ctx->d_flags.p, sizeof(int), cudaMemcpyDeviceToDevice, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(d->send.p + 8, ctx->d_lik_mean.p, n_loc * 8, cudaMemcpyDeviceToDevice, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(d->send.p + 8 + cap_pos * 8, ctx->d_lik_var.p, n_loc * 8, cudaMemcpyDeviceToDevice, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(d->send.p + 8 + cap_pos * 16, ctx->d_rng.p, w_loc * 32, cudaMemcpyDeviceToDevice, s));
    }
    if ((rc = gather_to_host(d, bytes))) return rc;
    if (w_loc) {
        float ms = 0.f;
        LCTP_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
        ctx->stats.stage_ms += ms;
        d->timing.kernel_ms += ms;
    }
    // scatter every rank's block back to worker order
    for (int r = 0; r < N; r++) {
        const unsigned char *base = d->host_recv.p + (size_t)r * bytes;
        if (*(const int *)base) {
            set_error("lctp_dist_solve_stage: candidate overflow on rank %d (more than 65535 candidates in one genotype)", r);
            return LCTP_E_CAPACITY;
        }
        const double *mean = (const double *)(base + 8), *var = (const double *)(base + 8 + cap_pos * 8);
        const uint64_t *rng = (const uint64_t *)(base + 8 + cap_pos * 16);
        size_t q = 0, k = 0;
        for (size_t w = r; w < n_workers; w += N, k++) {
            for (uint64_t j = worker_off[w]; j < worker_off[w + 1]; j++, q++) { lik_mean[j] = mean[q]; lik_var[j] = var[q]; }
            std::memcpy(worker_rng + 4 * w, rng + 4 * k, 32);
        }
    }
    return LCTP_OK;
}

int lctp_dist_solve(lctp_dist *d, lctp_locus_h *h, const lctp_stage *stages, size_t n_stages, size_t threads,
                    uint64_t rng[4], lctp_result *res) {
    if (!d || !h || !stages || !rng || !res || n_stages == 0 || n_stages > LCTP_MAX_STAGES) {
        set_error("lctp_dist_solve: invalid argument");
        return LCTP_E_INVALID;
    }
    const double t_in = now_s();
    const double k0 = d->timing.kernel_ms, c0 = d->timing.collective_ms;
    const uint64_t G = h->dev.G;
    std::memset(res, 0, sizeof(*res));
    threads = std::max<size_t>(1, std::min<size_t>(threads, G));     // genotype.rs:1247
    size_t n = G;
    // per-genotype results by id: the context's reusable arrays (clean between solves), see common.cuh
    lctp::per_id_reserve(d->ctx, G);
    std::vector<uint64_t> &ixs = d->ctx->id_list;
    std::vector<double> &lik_mean = d->ctx->id_mean, &lik_var = d->ctx->id_var;
    std::vector<uint16_t> &attempts = d->ctx->id_attempts;
    std::vector<uint64_t> touched;
    struct Cleaner { lctp_ctx *c; std::vector<uint64_t> &t; ~Cleaner() { lctp::per_id_clean(c, t); } } cleaner{d->ctx, touched};
    int rc;
    if (h->host.dont_skip || stages[0].in_size < G) {                // solve.rs:941-945
        if ((rc = lctp_dist_prefilter(d, h, stages[0].in_size, threads, ixs.data(), G, &n))) return rc;
    } else std::iota(ixs.begin(), ixs.begin() + G, 0);
    res->n_filtered = n;
    std::vector<uint64_t> wrng;
    if (threads > 1) {                                               // MainWorker::new, solve.rs:1007-1018
        wrng.resize(threads * 4);
        lctp_rng_worker_streams(rng, threads, wrng.data());
    }
    std::vector<uint64_t> off(threads + 1);
    std::vector<double> lm, lv;
    for (size_t s = 0; s < n_stages; s++) {
        const lctp_stage &st = stages[s];
        const bool has_next = s + 1 < n_stages;
        const size_t out_size = has_next ? stages[s + 1].in_size : 0;
        if (!(h->host.dont_skip || !has_next || out_size < n)) continue;   // solve.rs:1041-1045
        res->n_stage_in[s] = n;
        lm.resize(n); lv.resize(n);
        if (touched.empty()) touched.assign(ixs.begin(), ixs.begin() + n);     // later stages solve subsets of these
        if (threads == 1) {                                          // solve_single_thread: one stream, no shuffle
            off[0] = 0; off[1] = n;
            rc = lctp_dist_solve_stage(d, h, &st, ixs.data(), off.data(), 1, rng, lm.data(), lv.data());
        } else {
            const size_t nw = lctp_plan_stage(rng, ixs.data(), n, threads, off.data());
            rc = lctp_dist_solve_stage(d, h, &st, ixs.data(), off.data(), nw, wrng.data(), lm.data(), lv.data());
        }
        if (rc) return rc;
        for (size_t q = 0; q < n; q++) { lik_mean[ixs[q]] = lm[q]; lik_var[ixs[q]] = lv[q]; attempts[ixs[q]] = (uint16_t)st.attempts; }
        if (has_next)
            n = lctp_discard_improbable(ixs.data(), n, lik_mean.data(), lik_var.data(), attempts.data(),
                                        h->host.prob_thresh, out_size, threads);
    }
    const uint64_t n_filtered = res->n_filtered;
    uint64_t n_stage_in[LCTP_MAX_STAGES];
    std::memcpy(n_stage_in, res->n_stage_in, sizeof(n_stage_in));
    rc = lctp_produce_result(h, ixs.data(), n, lik_mean.data(), lik_var.data(), attempts.data(), res);
    if (rc) return rc;
    res->n_filtered = n_filtered;
    std::memcpy(res->n_stage_in, n_stage_in, sizeof(n_stage_in));
    const double wall_ms = (now_s() - t_in) * 1e3;
    d->timing.solves += 1;
    d->timing.wall_ms += wall_ms;
    d->timing.host_ms += wall_ms - (d->timing.kernel_ms - k0) - (d->timing.collective_ms - c0);
    return LCTP_OK;
}

}  // extern "C"

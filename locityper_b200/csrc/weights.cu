// weights.cu -- read weights from the k-mers unique to a locus (UniqueKmers, src/model/locs.rs:915-1003), the step of
// AllAlignments::load between read_next_alns and recover_and_group_alignments (:1144-1148):
//   build   UniqueKmers::new (:930-963): the canonical base-k k-mers (kmers::kmers::<u128, _, CANONICAL>,
//           src/seq/kmers.rs:163-202) of the contig sequences whose off-target count is 0.  Host work like in the
//           reference (one pass over the contig set per locus); the set goes to the device as an open-addressing table
//           of 128-bit keys.
//   weights calculate_read_weight (:968-1002) for every read of the locus in one launch: one thread per read end rolls
//           the forward / reverse-complement k-mers over its sequence, looks the canonical one up and counts the
//           NON-OVERLAPPING unique k-mers (after a hit the next k - 1 k-mers are skipped, :985-989 -- a sequential rule,
//           which is why the unit of parallelism is the read end and not the k-mer); the weight of the read is
//           clamp(intercept + count * slope, 0, 1).
// Byte work: 1 B per base in, 2 B per read end + 8 B per read out; the lookups stay in L2 (the table of a locus is a
// few MB).
#include "common.cuh"

#include <vector>

namespace lctp {

typedef unsigned __int128 u128;
static constexpr uint64_t EMPTY_HI = 0xFFFFFFFFFFFFFFFEull;        // valid k-mers are < 2^126 (k < 64), UNDEF is all ones

__host__ __device__ inline uint64_t kmer_slot_hash(uint64_t lo, uint64_t hi) {
    uint64_t x = lo ^ (hi * 0x9E3779B97F4A7C15ull);
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
    return x;
}

// Rolling canonical k-mers of one sequence (kmers::kmers::<u128, _, CANONICAL>): push(nt) feeds base i and returns
// true when it emits the k-mer with index i - (k - 1), i.e. for every i + 1 >= k, with (lo, hi) = all ones for a k-mer
// that contains an N (Kmer::UNDEF).
struct KmerRoller {
    u128 mask, fw, rv;
    uint32_t k, rv_shift;
    uint64_t reset, i;
    __host__ __device__ explicit KmerRoller(uint32_t k_)
        : mask((((u128)1) << (2 * k_)) - 1), fw(0), rv(0), k(k_), rv_shift(2 * k_ - 2), reset(k_ - 1), i(0) {}
    __host__ __device__ bool push(uint8_t nt, uint64_t &lo, uint64_t &hi) {
        const uint64_t at = i++;
        const uint32_t enc = nt == 'A' ? 0u : nt == 'C' ? 1u : nt == 'G' ? 2u : nt == 'T' ? 3u : 4u;
        lo = hi = ~0ull;
        if (enc == 4u) {
            reset = at + k;
            return at + 1 >= k;
        }
        fw = ((fw << 2) | enc) & mask;
        rv = (rv >> 2) | ((u128)(3u - enc) << rv_shift);
        if (at >= reset) {
            const u128 c = rv < fw ? rv : fw;
            lo = (uint64_t)c; hi = (uint64_t)(c >> 64);
            return true;
        }
        return at + 1 >= k;
    }
};

__global__ void __launch_bounds__(128)
k_read_weights(const uint8_t *__restrict__ seqs, const uint64_t *__restrict__ seq_off, uint64_t n_ends, uint32_t k,
               const ulonglong2 *__restrict__ table, uint64_t tab_mask, uint16_t *__restrict__ unique) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_ends) return;
    const uint64_t b = seq_off[q], len = seq_off[q + 1] - b;
    uint32_t count = 0;
    uint64_t next = 0;                                             // first k-mer index that is looked at again
    KmerRoller roll(k);
    for (uint64_t i = 0; i < len; i++) {
        uint64_t lo, hi;
        if (!roll.push(seqs[b + i], lo, hi)) continue;
        const uint64_t idx = i - (k - 1);
        if (idx < next) continue;
        uint64_t s = kmer_slot_hash(lo, hi) & tab_mask;
        for (;;) {
            const ulonglong2 e = table[s];
            if (e.x == lo && e.y == hi) {                          // unique k-mer: count it, skip the k - 1 that overlap it
                count = min(count + 1u, 65535u);                   // saturating_add
                next = idx + k;                                    // kmers_iter.nth(k - 2)
                break;
            }
            if (e.y == EMPTY_HI) break;
            s = (s + 1) & tab_mask;
        }
    }
    unique[q] = (uint16_t)count;
}

__global__ void __launch_bounds__(256)
k_weight_of_counts(const uint16_t *__restrict__ unique, uint64_t n_reads, uint32_t ends, double interc, double mult,
                   double *__restrict__ weight) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    uint32_t paired = 0;
    for (uint32_t e = 0; e < ends; e++) paired = (paired + unique[r * ends + e]) & 0xFFFFu;      // u16 sum
    const double w = __dadd_rn(interc, __dmul_rn((double)paired, mult));                       // locs.rs:997
    weight[r] = w < 0.0 ? 0.0 : w > 1.0 ? 1.0 : w;                                            // clamp(0.0, 1.0)
}

}  // namespace lctp

struct lctp_unique_kmers_h {
    lctp_ctx *ctx;
    uint32_t k;
    uint64_t n_unique, tab_mask;
    double weight_mult, weight_interc;
    lctp::DevBuf<ulonglong2> table;
};

using namespace lctp;

extern "C" int lctp_unique_kmers_build(lctp_ctx *ctx, const uint8_t *seqs, const uint64_t *seq_off, uint64_t n_seqs,
                                       const uint16_t *kmer_counts, const uint64_t *cnt_off, uint32_t k,
                                       uint16_t hard_threshold, uint16_t soft_threshold, lctp_unique_kmers_h **out,
                                       uint64_t *n_unique) {
    if (!ctx || !seqs || !seq_off || !kmer_counts || !cnt_off || !out) { set_error("lctp_unique_kmers_build: NULL argument"); return LCTP_E_INVALID; }
    if (k <= 1 || k >= 64) { set_error("lctp_unique_kmers_build: k = %u outside 2..=63 (u128 k-mers, locs.rs:937, kmers.rs:50)", k); return LCTP_E_INVALID; }
    if (hard_threshold > soft_threshold) { set_error("lctp_unique_kmers_build: hard threshold above the soft one (locs.rs:956)"); return LCTP_E_INVALID; }
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    set_alloc_stream(ctx->stream);
    uint64_t total = 0;
    for (uint64_t s = 0; s < n_seqs; s++) {
        const uint64_t len = seq_off[s + 1] - seq_off[s], nk = len + 1 >= k ? len + 1 - k : 0;
        if (cnt_off[s + 1] - cnt_off[s] != nk) {
            set_error("lctp_unique_kmers_build: sequence %llu has %llu k-mers but %llu counts (locs.rs:944)", (unsigned long long)s,
                      (unsigned long long)nk, (unsigned long long)(cnt_off[s + 1] - cnt_off[s]));
            return LCTP_E_INVALID;
        }
        total += nk;
    }
    uint64_t cap = 16;
    while (cap < 2 * total + 2) cap <<= 1;
    std::vector<ulonglong2> tab(cap, make_ulonglong2(~0ull, EMPTY_HI));
    uint64_t n = 0;
    for (uint64_t s = 0; s < n_seqs; s++) {
        const uint16_t *cnt = kmer_counts + cnt_off[s];
        const uint8_t *seq = seqs + seq_off[s];
        const uint64_t len = seq_off[s + 1] - seq_off[s];
        KmerRoller roll(k);
        for (uint64_t i = 0; i < len; i++) {
            uint64_t lo, hi;
            if (!roll.push(seq[i], lo, hi)) continue;
            if (cnt[i - (k - 1)] != 0) continue;                          // off-target k-mer (:946-950)
            uint64_t slot = kmer_slot_hash(lo, hi) & (cap - 1);
            for (;;) {
                ulonglong2 &e = tab[slot];
                if (e.x == lo && e.y == hi) break;
                if (e.y == EMPTY_HI) { e = make_ulonglong2(lo, hi); n++; break; }
                slot = (slot + 1) & (cap - 1);
            }
        }
    }
    auto h = new lctp_unique_kmers_h;
    h->ctx = ctx; h->k = k; h->n_unique = n; h->tab_mask = cap - 1;
    h->weight_mult = 1.0 / (double)(soft_threshold + 1 - hard_threshold);          // :957
    h->weight_interc = (1.0 - (double)hard_threshold) * h->weight_mult;            // :958
    int rc = h->table.alloc(cap);
    if (rc) { delete h; return rc; }
    cudaError_t e = cudaMemcpyAsync(h->table.p, tab.data(), cap * sizeof(ulonglong2), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { delete h; set_error("lctp_unique_kmers_build: %s", cudaGetErrorString(e)); return LCTP_E_CUDA; }
    ctx->stats.h2d_bytes += cap * sizeof(ulonglong2);
    *out = h;
    if (n_unique) *n_unique = n;
    return LCTP_OK;
}

extern "C" void lctp_unique_kmers_free(lctp_unique_kmers_h *u) {
    if (!u) return;
    cudaSetDevice(u->ctx->device);
    delete u;
}

extern "C" int lctp_read_weights(lctp_ctx *ctx, const lctp_unique_kmers_h *u, const uint8_t *seqs, const uint64_t *seq_off,
                                 uint64_t n_reads, uint32_t ends, uint16_t *unique, double *weight) {
    if (!ctx || !u) { set_error("lctp_read_weights: NULL handle"); return LCTP_E_INVALID; }
    if (ends != 1 && ends != 2) { set_error("lctp_read_weights: %u read ends per read (1 or 2)", ends); return LCTP_E_INVALID; }
    if (n_reads == 0) return LCTP_OK;
    if (!seqs || !seq_off || !unique || !weight) { set_error("lctp_read_weights: NULL array"); return LCTP_E_INVALID; }
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    set_alloc_stream(ctx->stream);
    cudaStream_t s = ctx->stream;
    const uint64_t n_ends = n_reads * ends, total = seq_off[n_ends];
    DevBuf<uint8_t> d_seq;
    DevBuf<uint64_t> d_off;
    DevBuf<uint16_t> d_unique;
    DevBuf<double> d_w;
    int rc;
    if ((rc = d_seq.alloc(total ? total : 1)) || (rc = d_off.alloc(n_ends + 1)) || (rc = d_unique.alloc(n_ends)) || (rc = d_w.alloc(n_reads))) return rc;
    if (total) LCTP_CUDA_CHECK(cudaMemcpyAsync(d_seq.p, seqs, total, cudaMemcpyHostToDevice, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(d_off.p, seq_off, (n_ends + 1) * 8, cudaMemcpyHostToDevice, s));
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[0], s));
    k_read_weights<<<(unsigned)((n_ends + 127) / 128), 128, 0, s>>>(d_seq.p, d_off.p, n_ends, u->k, u->table.p, u->tab_mask, d_unique.p);
    LCTP_CUDA_CHECK(cudaGetLastError());
    k_weight_of_counts<<<(unsigned)((n_reads + 255) / 256), 256, 0, s>>>(d_unique.p, n_reads, ends, u->weight_interc, u->weight_mult, d_w.p);
    LCTP_CUDA_CHECK(cudaGetLastError());
    ctx->launches += 2;
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[1], s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(unique, d_unique.p, n_ends * 2, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(weight, d_w.p, n_reads * 8, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
    float ms = 0.f;
    LCTP_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->stats.recruit_ms += ms;                                       // k-mer work is accounted with the recruitment kernels
    ctx->stats.recruit_launches += 2;
    ctx->stats.recruit_bases += total;
    ctx->stats.h2d_bytes += total + (n_ends + 1) * 8;
    ctx->stats.d2h_bytes += n_ends * 2 + n_reads * 8;
    return LCTP_OK;
}

extern "C" uint64_t lctp_unique_kmers_count(const lctp_unique_kmers_h *u) { return u ? u->n_unique : 0; }

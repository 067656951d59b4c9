// recruit.cu -- SURVEY.md section 8(f) rank 3, first slice: short-read recruitment on the device.
//
//   canonical minimizers of a sequence       kmers::minimizers::<u64, _, CANONICAL> (src/seq/kmers.rs:71-103, 256-340)
//   target tables minimizer -> [(locus, info)] TargetBuilder::add (src/seq/recruit.rs:680-735), built on the host from the
//                                             device-computed minimizers of the target sequences
//   per read (pair): match counters           BaseMatchCount<u16>::inc / has_rare / better_fraction /
//                                             better_pair_fraction (:234-366), Fraction<u16> comparison (frac.rs:92-98)
//   decision                                  recruit_short_read (:852-881), recruit_read_pair (:885-930)
//   long single-end reads                    recruit_long_read + has_matching_stretch (:932-998)
//
// One thread per read (pair): the minimizer scan is a sequential sliding-window minimum (the emitted set depends on the
// scan order through `last_pos` and the window restarts after an N), 150-250 steps for a short read; a read touches its
// own bytes only, so this is streaming byte work: 1 B per base in, 4 B + 4 B per recruited (read, locus) out.  The
// minimizer table is an open-addressing hash table in global memory (L2-resident: 16 B per slot).
#include "common.cuh"

#include <vector>
#include <algorithm>
#include <cmath>
#include <cfloat>
#include <unordered_map>

namespace lctp {

static constexpr uint64_t KM_UNDEF = ~0ull;        // Kmer::UNDEF = Self::MAX (kmers.rs:45)
static constexpr uint32_t MAXW = 64;               // MAX_MINIMIZER_W (kmers.rs:208)
static constexpr int LOCI_PER_READ = 8;            // loci one read can match before LCTP_E_CAPACITY

__host__ __device__ __forceinline__ uint64_t fast_hash64(uint64_t x) {      // Minimizer for u64, kmers.rs:93-103
    x = ~x;
    x ^= x >> 23;
    x *= 0x2127599bf4325c37ull;
    x ^= x >> 47;
    return x;
}

// kmers::minimizers::<u64, _, CANONICAL> (kmers.rs:256-331), statement for statement; `emit(pos, hash, forward)`.
// `hashes`: the 64-entry hash ring of this thread, element q at hashes[q * stride] (shared memory, one column per thread:
// as a per-thread local array it went through local memory -- 390 MB of DRAM writes per 200k read pairs, ncu).
// RING: entries of the ring, a power of two >= w (the reference's CircArray has 64; only the last w entries are ever read).
template <uint32_t RING, typename Emit>
__device__ __forceinline__ void dev_minimizers(const uint8_t *__restrict__ seq, uint32_t len, uint32_t k, uint32_t w,
                                               uint64_t *hashes, uint32_t stride, Emit emit) {
    const uint64_t mask = (1ull << (2 * k)) - 1ull;
    const uint32_t rv_shift = 2 * k - 2;
    uint64_t fw_kmer = 0, rv_kmer = 0;
    const uint32_t k_1 = k - 1, w_1 = w - 1;
    uint64_t forward = ~0ull;                      // CircArray<bool> as a 64-bit mask
#pragma unroll 1
    for (uint32_t q = 0; q < RING; q++) hashes[q * stride] = KM_UNDEF;
    long long last_pos = -1;
    uint32_t best_pos = 0;
    uint64_t best_hash = KM_UNDEF;
    uint32_t first_kmer = k_1, first_window = k_1 + w_1;
#pragma unroll 1
    for (uint32_t i = 0; i < len; i++) {
        uint64_t fw_enc, rv_enc;
        switch (seq[i]) {
        case 'A': fw_enc = 0; rv_enc = 3; break;
        case 'C': fw_enc = 1; rv_enc = 2; break;
        case 'G': fw_enc = 2; rv_enc = 1; break;
        case 'T': fw_enc = 3; rv_enc = 0; break;
        default: first_kmer = i + k; fw_enc = 0; rv_enc = 0; break;
        }
        fw_kmer = ((fw_kmer << 2) | fw_enc) & mask;
        rv_kmer = (rv_kmer >> 2) | (rv_enc << rv_shift);
        const bool f = !(rv_kmer < fw_kmer);
        const uint64_t kmer = f ? fw_kmer : rv_kmer;
        const uint64_t h = i < first_kmer ? KM_UNDEF : fast_hash64(kmer);
        hashes[(i & (RING - 1)) * stride] = h;
        forward = (forward & ~(1ull << (i & (MAXW - 1)))) | ((uint64_t)f << (i & (MAXW - 1)));
        if (h < best_hash) { best_hash = h; best_pos = i; }
        if (i < first_window) continue;
        const uint32_t start = i - w_1;
        if (best_pos < start) {
            uint32_t p = start;                    // find_min, kmers.rs:237-252
            uint64_t m = hashes[(start & (RING - 1)) * stride];
            for (uint32_t j = start + 1; j < i + 1; j++) {
                const uint64_t v = hashes[(j & (RING - 1)) * stride];
                if (v < m) { p = j; m = v; }
            }
            best_pos = p; best_hash = m;
            if (best_hash == KM_UNDEF) { first_window = first_window + w_1; continue; }
        }
        if ((long long)best_pos > last_pos) {
            last_pos = (long long)best_pos;
            emit(best_pos - k_1, best_hash, (uint32_t)((forward >> (best_pos & (MAXW - 1))) & 1ull));
        }
    }
}

// threads per CTA for a ring of RING entries: 16 or 32 KB of hash rings per CTA
template <uint32_t RING> struct RecruitCfg { static constexpr int THREADS = RING <= 32 ? 128 : 64; };

template <uint32_t RING>
__global__ void __launch_bounds__(RecruitCfg<RING>::THREADS)
k_minimizers(const uint8_t *__restrict__ seqs, const uint64_t *__restrict__ off, uint64_t n, uint32_t k, uint32_t w,
             uint32_t *__restrict__ count, uint64_t *__restrict__ hash, uint32_t *__restrict__ pos, uint8_t *__restrict__ fw) {
    constexpr int NT = RecruitCfg<RING>::THREADS;
    __shared__ uint64_t ring[RING * NT];
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint64_t b = off[s];
    uint32_t c = 0;
    dev_minimizers<RING>(seqs + b, (uint32_t)(off[s + 1] - b), k, w, ring + threadIdx.x, NT, [&](uint32_t p, uint64_t h, uint32_t f) {
        hash[b + c] = h; pos[b + c] = p; fw[b + c] = (uint8_t)f; c++;
    });
    count[s] = c;
}

// ---- the minimizer table --------------------------------------------------------------------------------------------
struct TableDev {
    const uint64_t *key;        // [cap] KM_UNDEF = empty slot
    const uint2 *span;          // [cap] (first entry, entries)
    const uint32_t *e_locus;    // entries: locus ...
    const uint8_t *e_info;      // ... and MinimInfo: direction (bits 0-1) | rare << 2
    uint64_t mask;
};
__host__ __device__ __forceinline__ uint64_t table_slot(uint64_t key, uint64_t mask) { return (key * 0x9E3779B97F4A7C15ull >> 17) & mask; }

// BaseMatchCount<u16> (recruit.rs:234-252): arr = [common-backward, common-forward, rare-backward, rare-forward]
struct Bmc { uint16_t arr[4]; };
__device__ __forceinline__ bool directed_to(uint32_t info, bool forward) { return (info & (1u + (forward ? 1u : 0u))) != 0; }   // :641-643
__device__ __forceinline__ void bmc_inc(Bmc &c, bool forward, uint32_t info) {                                               // :246-252
    const int i = ((info >> 2) & 1) << 1;
    c.arr[i] = (uint16_t)(c.arr[i] + (directed_to(info, !forward) ? 1 : 0));
    c.arr[i | 1] = (uint16_t)(c.arr[i | 1] + (directed_to(info, forward) ? 1 : 0));
}
__device__ __forceinline__ bool bmc_has_rare(const Bmc &c) { return c.arr[2] != 0 || c.arr[3] != 0; }                        // :257-259
__device__ __forceinline__ uint16_t fw_num(const Bmc &c) { return (uint16_t)(3 * c.arr[3] + c.arr[1]); }                     // WORTH = 3, :284-303
__device__ __forceinline__ uint16_t bw_num(const Bmc &c) { return (uint16_t)(3 * c.arr[2] + c.arr[0]); }
__device__ __forceinline__ uint16_t fw_den(const Bmc &c, uint16_t t) { return (uint16_t)(3 * (t - c.arr[1]) + c.arr[1]); }  // :313-322
__device__ __forceinline__ uint16_t bw_den(const Bmc &c, uint16_t t) { return (uint16_t)(3 * (t - c.arr[0]) + c.arr[0]); }
__device__ __forceinline__ bool frac_ge(uint16_t n1, uint16_t d1, uint16_t n2, uint16_t d2) {                                // frac.rs:92-98
    return (uint32_t)n1 * d2 >= (uint32_t)n2 * d1;
}

struct ReadsDev {
    uint64_t n;
    const uint64_t *off1, *off2;    // off2 == nullptr: single-end
    const uint8_t *seq1, *seq2;
};

struct LongParams { double match_frac; uint32_t stretch_minims, stretch_score; };

// recruit_long_read (recruit.rs:967-998) + has_matching_stretch (:932-964) for one single-end read longer than
// READ_LENGTH_THRESH.  Two scans of the read: the first counts matches per locus (BaseMatchCount<u32>), the second -- only
// for loci that pass the count threshold and have at least `stretch_minims` usable minimizers -- runs the Kadane-style
// stretch test; the minimizers are recomputed instead of buffered (a 15-kb read has ~3,000 of them).
template <uint32_t RING>
__device__ void recruit_long(const uint8_t *__restrict__ seq, uint32_t len, const TableDev &T, uint32_t k, uint32_t w,
                             const LongParams &P, uint64_t *ring, uint32_t stride, uint32_t cap, uint32_t *ans,
                             uint32_t &n_ans, bool &overflow) {
    uint32_t loci[LOCI_PER_READ];
    uint32_t cnt[LOCI_PER_READ][4];
    int nm = 0;
    uint32_t total = 0;
    auto find = [&](uint64_t h, uint2 &sp) -> bool {
        if (h == KM_UNDEF) return false;
        uint64_t s = table_slot(h, T.mask);
        while (true) {
            const uint64_t kk = T.key[s];
            if (kk == h) break;
            if (kk == KM_UNDEF) return false;
            s = (s + 1) & T.mask;
        }
        sp = T.span[s];
        return true;
    };
    dev_minimizers<RING>(seq, len, k, w, ring, stride, [&](uint32_t, uint64_t h, uint32_t f) {
        total++;
        uint2 sp;
        if (!find(h, sp)) return;
        for (uint32_t e = sp.x; e < sp.x + sp.y; e++) {
            const uint32_t locus = T.e_locus[e], info = T.e_info[e];
            int z = 0;
            while (z < nm && loci[z] != locus) z++;
            if (z == nm) {
                if (nm == LOCI_PER_READ) { overflow = true; continue; }
                loci[nm] = locus; cnt[nm][0] = cnt[nm][1] = cnt[nm][2] = cnt[nm][3] = 0;
                nm++;
            }
            const int i = ((info >> 2) & 1) << 1;                                          // BaseMatchCount<u32>::inc, :246-252
            cnt[z][i] += directed_to(info, f == 0) ? 1u : 0u;
            cnt[z][i | 1] += directed_to(info, f != 0) ? 1u : 0u;
        }
    });
    for (int z = 0; z < nm; z++) {
        const uint32_t bw_c = cnt[z][0], fw_c = cnt[z][1], bw_r = cnt[z][2], fw_r = cnt[z][3];
        const uint32_t num = fw_r >= bw_r ? fw_r : bw_r, den = fw_r >= bw_r ? total - fw_c : total - bw_c;   // rare_fraction, :266-274
        const uint32_t thr = max(1u, __double2uint_rz(ceil(__dmul_rn((double)min(P.stretch_minims, den), P.match_frac))));   // long_read_threshold
        if (num < thr) continue;
        bool okay = den < P.stretch_minims;
        if (!okay) {                                                                        // has_matching_stretch, :944-963
            uint32_t s_fw = 0, s_bw = 0;
            const uint32_t want = loci[z];
            dev_minimizers<RING>(seq, len, k, w, ring, stride, [&](uint32_t, uint64_t h, uint32_t f) {
                if (okay) return;
                uint2 sp;
                if (find(h, sp))
                    for (uint32_t e = sp.x; e < sp.x + sp.y; e++)
                        if (T.e_locus[e] == want) {
                            const uint32_t info = T.e_info[e], x = 1u + ((info >> 2) & 1u) * 3u;   // SUBSUM_PENALTY + rare * SUBSUM_BONUS
                            s_fw += directed_to(info, f != 0) ? x : 0u;
                            s_bw += directed_to(info, f == 0) ? x : 0u;
                        }
                s_fw = s_fw >= 1u ? s_fw - 1u : 0u;                                         // saturating_sub(SUBSUM_PENALTY)
                s_bw = s_bw >= 1u ? s_bw - 1u : 0u;
                if (s_fw >= P.stretch_score || s_bw >= P.stretch_score) okay = true;
            });
        }
        if (okay) {
            if (n_ans < cap) {
                uint32_t y = n_ans;
                while (y > 0 && ans[y - 1] > loci[z]) { ans[y] = ans[y - 1]; y--; }
                ans[y] = loci[z];
            }
            n_ans++;
        }
    }
}

template <uint32_t RING>
__global__ void __launch_bounds__(RecruitCfg<RING>::THREADS)
k_recruit_short(ReadsDev R, TableDev T, uint32_t k, uint32_t w, uint32_t fnum, uint32_t fden, LongParams LP, uint32_t cap,
                uint32_t *__restrict__ ans_count, uint32_t *__restrict__ ans_locus, int *__restrict__ err) {
    constexpr int NT = RecruitCfg<RING>::THREADS;
    __shared__ uint64_t ring[RING * NT];
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R.n) return;
    if (!R.off2 && R.off1[r + 1] - R.off1[r] > 500) {         // RecruitableRecord::recruit, recruit.rs:589 (READ_LENGTH_THRESH)
        uint32_t n_long = 0;
        bool ovf = false;
        const uint64_t b = R.off1[r];
        recruit_long<RING>(R.seq1 + b, (uint32_t)(R.off1[r + 1] - b), T, k, w, LP, ring + threadIdx.x, NT, cap, ans_locus + r * cap, n_long, ovf);
        ans_count[r] = n_long;
        if (ovf) atomicOr(err, 1);
        return;
    }
    uint32_t loci[LOCI_PER_READ];
    Bmc first[LOCI_PER_READ], second[LOCI_PER_READ];
    int nm = 0;
    bool overflow = false;
    uint32_t total1 = 0, total2 = 0;
    auto lookup = [&](uint64_t h, uint32_t f, bool is_second) {
        if (h == KM_UNDEF) return;
        uint64_t s = table_slot(h, T.mask);
        while (true) {
            const uint64_t kk = T.key[s];
            if (kk == h) break;
            if (kk == KM_UNDEF) return;
            s = (s + 1) & T.mask;
        }
        const uint2 sp = T.span[s];
        for (uint32_t e = sp.x; e < sp.x + sp.y; e++) {
            const uint32_t locus = T.e_locus[e], info = T.e_info[e];
            int z = 0;
            while (z < nm && loci[z] != locus) z++;
            if (z == nm) {
                if (is_second) continue;                       // no new loci from the second mate (:914-916)
                if (nm == LOCI_PER_READ) { overflow = true; continue; }
                loci[nm] = locus;
                first[nm] = Bmc{{0, 0, 0, 0}}; second[nm] = Bmc{{0, 0, 0, 0}};
                nm++;
            }
            bmc_inc(is_second ? second[z] : first[z], f != 0, info);
        }
    };
    {
        const uint64_t b = R.off1[r];
        dev_minimizers<RING>(R.seq1 + b, (uint32_t)(R.off1[r + 1] - b), k, w, ring + threadIdx.x, NT, [&](uint32_t, uint64_t h, uint32_t f) { total1++; lookup(h, f, false); });
    }
    uint32_t n_ans = 0;
    uint32_t *ans = ans_locus + r * cap;
    auto push = [&](uint32_t locus) {              // the answer is a set: kept in ascending locus order
        if (n_ans < cap) {
            uint32_t y = n_ans;
            while (y > 0 && ans[y - 1] > locus) { ans[y] = ans[y - 1]; y--; }
            ans[y] = locus;
        }
        n_ans++;
    };
    if (total1 > 65535u) atomicOr(err, 2);         // u16::try_from(..).expect(..)
    if (R.off2) {
        if (nm != 0) {                                            // :906
            const uint64_t b = R.off2[r];
            dev_minimizers<RING>(R.seq2 + b, (uint32_t)(R.off2[r + 1] - b), k, w, ring + threadIdx.x, NT, [&](uint32_t, uint64_t h, uint32_t f) { total2++; lookup(h, f, true); });
            if (total2 > 65535u) atomicOr(err, 2);
            for (int z = 0; z < nm; z++) {
                const Bmc a = first[z], b2 = second[z];
                if (!(bmc_has_rare(a) || bmc_has_rare(b2))) continue;                              // :923
                uint16_t n1, d1, n2, d2;                                                            // better_pair_fraction, :349-366
                if ((uint16_t)(fw_num(a) + bw_num(b2)) >= (uint16_t)(bw_num(a) + fw_num(b2))) {
                    n1 = fw_num(a); d1 = fw_den(a, (uint16_t)total1); n2 = bw_num(b2); d2 = bw_den(b2, (uint16_t)total2);
                } else {
                    n1 = bw_num(a); d1 = bw_den(a, (uint16_t)total1); n2 = fw_num(b2); d2 = fw_den(b2, (uint16_t)total2);
                }
                if (frac_ge(n1, d1, (uint16_t)fnum, (uint16_t)fden) && frac_ge(n2, d2, (uint16_t)fnum, (uint16_t)fden)) push(loci[z]);
            }
        }
    } else {
        for (int z = 0; z < nm; z++) {
            const Bmc a = first[z];
            if (!bmc_has_rare(a)) continue;                                                        // :877
            uint16_t n1, d1;                                                                        // better_fraction, :337-346
            if (fw_num(a) >= bw_num(a)) { n1 = fw_num(a); d1 = fw_den(a, (uint16_t)total1); }
            else { n1 = bw_num(a); d1 = bw_den(a, (uint16_t)total1); }
            if (frac_ge(n1, d1, (uint16_t)fnum, (uint16_t)fden)) push(loci[z]);
        }
    }
    ans_count[r] = n_ans;
    if (overflow) atomicOr(err, 1);
}

template <typename T>
static int put(DevBuf<T> &dst, const T *src, size_t n, cudaStream_t s) {
    int rc = dst.alloc(n ? n : 1);
    if (rc) return rc;
    if (n) LCTP_CUDA_CHECK(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return LCTP_OK;
}

int minimizers(lctp_ctx *ctx, const uint8_t *seqs, const uint64_t *off, uint64_t n, uint32_t k, uint32_t w,
               uint32_t *count, uint64_t *hash, uint32_t *pos, uint8_t *fw) {
    cudaStream_t s = ctx->stream;
    if (n == 0) return LCTP_OK;
    const uint64_t total = off[n];
    DevBuf<uint8_t> d_seq, d_fw;
    DevBuf<uint64_t> d_off, d_hash;
    DevBuf<uint32_t> d_cnt, d_pos;
    int rc;
    if ((rc = put(d_seq, seqs, (size_t)total, s)) || (rc = put(d_off, off, (size_t)n + 1, s))) return rc;
    if ((rc = d_cnt.alloc(n)) || (rc = d_hash.alloc(total ? total : 1)) || (rc = d_pos.alloc(total ? total : 1)) || (rc = d_fw.alloc(total ? total : 1))) return rc;
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[0], s));
    if (w <= 16) k_minimizers<16><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_seq.p, d_off.p, n, k, w, d_cnt.p, d_hash.p, d_pos.p, d_fw.p);
    else if (w <= 32) k_minimizers<32><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_seq.p, d_off.p, n, k, w, d_cnt.p, d_hash.p, d_pos.p, d_fw.p);
    else k_minimizers<64><<<(unsigned)((n + 63) / 64), 64, 0, s>>>(d_seq.p, d_off.p, n, k, w, d_cnt.p, d_hash.p, d_pos.p, d_fw.p);
    ctx->launches++;
    LCTP_CUDA_CHECK(cudaGetLastError());
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[1], s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(count, d_cnt.p, n * 4, cudaMemcpyDeviceToHost, s));
    if (total) {
        LCTP_CUDA_CHECK(cudaMemcpyAsync(hash, d_hash.p, total * 8, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(pos, d_pos.p, total * 4, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(fw, d_fw.p, total, cudaMemcpyDeviceToHost, s));
    }
    LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
    float ms = 0.f;
    LCTP_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->stats.recruit_ms += ms;
    ctx->stats.recruit_launches += 1;
    ctx->stats.recruit_bases += total;
    ctx->stats.h2d_bytes += total + (n + 1) * 8;
    ctx->stats.d2h_bytes += n * 4 + total * 13;
    return LCTP_OK;
}

}  // namespace lctp

using namespace lctp;

// Fraction::<u16>::approximate (src/math/frac.rs:48-80)
void lctp_fraction_approximate_u16(double x, uint16_t *num, uint16_t *den) {
    uint32_t a2 = 1, a1 = (uint32_t)std::floor(x), b2 = 0, b1 = 1;
    double xk = x;
    for (int it = 0; it < 20; it++) {
        const double numer = xk - std::floor(xk);
        if (numer <= DBL_EPSILON) break;
        xk = 1.0 / numer;
        const double fl = std::floor(xk);
        if (!(fl >= 0.0 && fl <= 65535.0)) break;                    // T::from_f64 -> None
        const uint64_t f = (uint64_t)fl;
        const uint64_t a0 = f * a1 + a2, b0 = f * b1 + b2;
        if (f * a1 > 65535 || a0 > 65535) break;                     // checked_mul / checked_add
        if (f * b1 > 65535 || b0 > 65535) break;
        a2 = a1; a1 = (uint32_t)a0; b2 = b1; b1 = (uint32_t)b0;
        if (std::fabs((double)a1 / (double)b1 - x) <= DBL_EPSILON) break;
    }
    *num = (uint16_t)a1; *den = (uint16_t)b1;
}

struct lctp_targets_h {
    lctp_ctx *ctx = nullptr;
    uint32_t k = 0, w = 0, n_loci = 0;
    uint16_t fnum = 0, fden = 1;
    LongParams lp = {0.0, 0, 0};
    uint64_t n_keys = 0, n_entries = 0, cap = 0;
    DevBuf<uint64_t> key;
    DevBuf<uint2> span;
    DevBuf<uint32_t> e_locus;
    DevBuf<uint8_t> e_info;
    // host copy of the entries in insertion order (lctp_targets_entries: tests, debugging)
    std::vector<uint64_t> h_key;
    std::vector<uint32_t> h_locus;
    std::vector<uint8_t> h_info;
};

int lctp_minimizers(lctp_ctx *ctx, const uint8_t *seqs, const uint64_t *off, uint64_t n, uint32_t k, uint32_t w,
                    uint32_t *count, uint64_t *hash, uint32_t *pos, uint8_t *fw) {
    if (!ctx || (n && (!seqs || !off || !count || !hash || !pos || !fw))) { set_error("lctp_minimizers: NULL argument"); return LCTP_E_INVALID; }
    if (k == 0 || k > 31 || w < 2 || w >= MAXW) {      // Params::new, recruit.rs:77-80 (MAX_KMER_SIZE = 31 for u64); w < 64 (kmers.rs:262)
        set_error("lctp_minimizers: minimizer size (%u, %u) out of range (k in [1, 31], w in [2, 63])", k, w);
        return LCTP_E_INVALID;
    }
    for (uint64_t s = 0; s < n; s++)
        if (off[s + 1] < off[s] || off[s + 1] - off[s] > 0xFFFFFFF0ull) { set_error("lctp_minimizers: bad offsets"); return LCTP_E_INVALID; }
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    set_alloc_stream(ctx->stream);
    return minimizers(ctx, seqs, off, n, k, w, count, hash, pos, fw);
}

int lctp_targets_build(lctp_ctx *ctx, const lctp_target_seqs *in, lctp_targets_h **out) {
    if (!ctx || !in || !out || !in->seq_off || !in->seqs || !in->seq_locus || !in->cnt_off || !in->kmer_counts || in->n_seqs == 0) {
        set_error("lctp_targets_build: NULL argument");
        return LCTP_E_INVALID;
    }
    const double min_frac = 1.0 / 4.0;                               // SUBSUM_PENALTY / (SUBSUM_BONUS + 1), recruit.rs:82-84
    if (!(in->match_frac >= min_frac && in->match_frac <= 1.0) || in->thresh_kmer_count == 0 || in->thresh_kmer_count > 65535 ||
        in->match_length < 200 || in->match_length > 100000) {           // Params::new, recruit.rs:82-91
        set_error("lctp_targets_build: match fraction must be in [0.25, 1], the k-mer threshold positive, the match length in [200, 100000]");
        return LCTP_E_INVALID;
    }
    const uint64_t n = in->n_seqs, total = in->seq_off[n];
    std::vector<uint32_t> cnt(n), pos(total ? total : 1);
    std::vector<uint64_t> hash(total ? total : 1);
    std::vector<uint8_t> fw(total ? total : 1);
    int rc = lctp_minimizers(ctx, in->seqs, in->seq_off, n, in->minimizer_k, in->minimizer_w, cnt.data(), hash.data(), pos.data(), fw.data());
    if (rc) return rc;
    const uint32_t base_k = in->base_k, mk = in->minimizer_k;
    const size_t shift = mk <= base_k ? (base_k - mk) / 2 : mk - base_k;             // recruit.rs:688-692
    // minim_to_loci in insertion order: per minimizer the index of its LAST entry (v.last_mut(), :721-724)
    auto h = new lctp_targets_h;
    h->ctx = ctx; h->k = mk; h->w = in->minimizer_w;
    lctp_fraction_approximate_u16(in->match_frac, &h->fnum, &h->fden);              // Params::new, recruit.rs:101
    h->lp.match_frac = in->match_frac;
    h->lp.stretch_minims = (2 * in->match_length + (in->minimizer_w + 1) - 1) / (in->minimizer_w + 1);     // fast_ceil_div, :95
    h->lp.stretch_score = (uint32_t)std::ceil(std::max((double)h->lp.stretch_minims * (4.0 * in->match_frac - 1.0), 3.0));   // :96-98
    std::unordered_map<uint64_t, uint32_t> last;
    uint32_t prev_locus = 0;
    for (uint64_t s = 0; s < n; s++) {
        const uint32_t locus = in->seq_locus[s];
        if (s && locus < prev_locus) { delete h; set_error("lctp_targets_build: seq_locus must be ascending"); return LCTP_E_INVALID; }
        prev_locus = locus;
        h->n_loci = std::max(h->n_loci, locus + 1);
        const size_t len = in->seq_off[s + 1] - in->seq_off[s];
        const size_t n_counts = in->cnt_off[s + 1] - in->cnt_off[s];
        if (n_counts != (len + 1 > base_k ? len + 1 - base_k : 0) || n_counts == 0) {   // the reference asserts (:698-699)
            delete h; set_error("lctp_targets_build: sequence %llu: %zu k-mer counts for length %zu, k %u", (unsigned long long)s, n_counts, len, base_k);
            return LCTP_E_INVALID;
        }
        const uint16_t *counts = in->kmer_counts + in->cnt_off[s];
        const uint64_t b = in->seq_off[s];
        for (uint32_t q = 0; q < cnt[s]; q++) {
            const size_t p = pos[b + q];
            bool rare;
            if (mk <= base_k) rare = counts[std::min(p >= shift ? p - shift : 0, n_counts - 1)] < in->thresh_kmer_count;      // :708-711
            else {
                if (p + shift >= n_counts) { delete h; set_error("lctp_targets_build: k-mer counts too short for minimizer k > base k"); return LCTP_E_INVALID; }
                rare = counts[p] < in->thresh_kmer_count && counts[p + shift] < in->thresh_kmer_count;                        // :713
            }
            const uint64_t key = hash[b + q];
            auto it = last.find(key);
            if (it != last.end() && h->h_locus[it->second] == locus) {                  // MinimInfo::update, :634-638
                uint8_t &inf = h->h_info[it->second];
                inf = (uint8_t)((inf | (1 + fw[b + q])) & (rare ? 0x7 : 0x3));
            } else {                                                                    // MinimInfo::new, :626-632
                last[key] = (uint32_t)h->h_key.size();
                h->h_key.push_back(key); h->h_locus.push_back(locus);
                h->h_info.push_back((uint8_t)((1 + fw[b + q]) | (rare ? 4 : 0)));
            }
        }
    }
    // device table: entries grouped by key (stable: insertion order inside a key), open addressing over the distinct keys
    const size_t ne = h->h_key.size();
    std::vector<uint32_t> idx(ne);
    for (size_t i = 0; i < ne; i++) idx[i] = (uint32_t)i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b2) { return h->h_key[a] < h->h_key[b2]; });
    std::vector<uint32_t> e_locus(ne ? ne : 1);
    std::vector<uint8_t> e_info(ne ? ne : 1);
    uint64_t cap = 16;
    while (cap < 2 * last.size()) cap <<= 1;
    std::vector<uint64_t> tkey(cap, KM_UNDEF);
    std::vector<uint2> tspan(cap, make_uint2(0, 0));
    for (size_t i = 0; i < ne;) {
        size_t j = i;
        while (j < ne && h->h_key[idx[j]] == h->h_key[idx[i]]) { e_locus[j] = h->h_locus[idx[j]]; e_info[j] = h->h_info[idx[j]]; j++; }
        const uint64_t key = h->h_key[idx[i]];
        if (key != KM_UNDEF) {
            uint64_t s = table_slot(key, cap - 1);
            while (tkey[s] != KM_UNDEF) s = (s + 1) & (cap - 1);
            tkey[s] = key; tspan[s] = make_uint2((uint32_t)i, (uint32_t)(j - i));
        }
        i = j;
    }
    h->n_keys = last.size(); h->n_entries = ne; h->cap = cap;
    cudaStream_t st = ctx->stream;
    if ((rc = put(h->key, tkey.data(), cap, st)) || (rc = put(h->span, tspan.data(), cap, st)) ||
        (rc = put(h->e_locus, e_locus.data(), e_locus.size(), st)) || (rc = put(h->e_info, e_info.data(), e_info.size(), st))) { delete h; return rc; }
    LCTP_CUDA_CHECK(cudaStreamSynchronize(st));
    ctx->stats.h2d_bytes += cap * 16 + ne * 5;
    *out = h;
    return LCTP_OK;
}

void lctp_targets_free(lctp_targets_h *t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    set_alloc_stream(t->ctx->stream);
    delete t;
}

uint64_t lctp_targets_entries(const lctp_targets_h *t, uint64_t *key, uint32_t *locus, uint8_t *info, uint64_t cap) {
    if (!t) return 0;
    for (uint64_t q = 0; q < t->h_key.size() && q < cap; q++) { key[q] = t->h_key[q]; locus[q] = t->h_locus[q]; info[q] = t->h_info[q]; }
    return t->h_key.size();
}

void lctp_targets_match_frac(const lctp_targets_h *t, uint16_t *num, uint16_t *den) {
    if (t) { *num = t->fnum; *den = t->fden; }
}

int lctp_recruit_short(lctp_ctx *ctx, const lctp_targets_h *t, const lctp_reads *reads, uint32_t cap, uint32_t *ans_count,
                       uint32_t *ans_locus) {
    if (!ctx || !t || !reads || !ans_count || !ans_locus || cap == 0 || (reads->n_reads && (!reads->off1 || !reads->seq1)) ||
        ((reads->off2 == nullptr) != (reads->seq2 == nullptr))) {
        set_error("lctp_recruit_short: invalid argument");
        return LCTP_E_INVALID;
    }
    const uint64_t n = reads->n_reads;
    if (n == 0) return LCTP_OK;
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    set_alloc_stream(ctx->stream);
    cudaStream_t s = ctx->stream;
    for (uint64_t r = 0; r < n; r++) {
        const uint64_t l1 = reads->off1[r + 1] - reads->off1[r], l2 = reads->off2 ? reads->off2[r + 1] - reads->off2[r] : 0;
        if (reads->off1[r + 1] < reads->off1[r] || l1 > 0xFFFFFFF0ull || l2 > 0xFFFFFFF0ull) {
            set_error("lctp_recruit_short: bad offsets of read %llu", (unsigned long long)r);
            return LCTP_E_INVALID;
        }
    }
    DevBuf<uint8_t> d_s1, d_s2;
    DevBuf<uint64_t> d_o1, d_o2;
    DevBuf<uint32_t> d_cnt, d_ans;
    DevBuf<int> d_err;
    int rc;
    if ((rc = put(d_s1, reads->seq1, (size_t)reads->off1[n], s)) || (rc = put(d_o1, reads->off1, (size_t)n + 1, s))) return rc;
    if (reads->off2 && ((rc = put(d_s2, reads->seq2, (size_t)reads->off2[n], s)) || (rc = put(d_o2, reads->off2, (size_t)n + 1, s)))) return rc;
    if ((rc = d_cnt.alloc(n)) || (rc = d_ans.alloc(n * cap)) || (rc = d_err.alloc(1))) return rc;
    LCTP_CUDA_CHECK(cudaMemsetAsync(d_err.p, 0, sizeof(int), s));
    LCTP_CUDA_CHECK(cudaMemsetAsync(d_ans.p, 0xFF, n * cap * 4, s));
    ReadsDev R;
    R.n = n; R.off1 = d_o1.p; R.seq1 = d_s1.p; R.off2 = reads->off2 ? d_o2.p : nullptr; R.seq2 = reads->off2 ? d_s2.p : nullptr;
    TableDev T;
    T.key = t->key.p; T.span = t->span.p; T.e_locus = t->e_locus.p; T.e_info = t->e_info.p; T.mask = t->cap - 1;
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[0], s));
    if (t->w <= 16) k_recruit_short<16><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(R, T, t->k, t->w, t->fnum, t->fden, t->lp, cap, d_cnt.p, d_ans.p, d_err.p);
    else if (t->w <= 32) k_recruit_short<32><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(R, T, t->k, t->w, t->fnum, t->fden, t->lp, cap, d_cnt.p, d_ans.p, d_err.p);
    else k_recruit_short<64><<<(unsigned)((n + 63) / 64), 64, 0, s>>>(R, T, t->k, t->w, t->fnum, t->fden, t->lp, cap, d_cnt.p, d_ans.p, d_err.p);
    ctx->launches++;
    LCTP_CUDA_CHECK(cudaGetLastError());
    LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[1], s));
    int err = 0;
    LCTP_CUDA_CHECK(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(ans_count, d_cnt.p, n * 4, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(ans_locus, d_ans.p, n * cap * 4, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
    if (err) {
        set_error("lctp_recruit_short: %s", (err & 2) ? "a read has more than 65535 minimizers (the reference panics)"
                                                      : "a read matches more than 8 loci (LOCI_PER_READ)");
        return (err & 2) ? LCTP_E_INVALID : LCTP_E_CAPACITY;
    }
    const uint64_t bases = reads->off1[n] + (reads->off2 ? reads->off2[n] : 0);
    float ms = 0.f;
    LCTP_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->stats.recruit_ms += ms;
    ctx->stats.recruit_launches += 1;
    ctx->stats.recruit_bases += bases;
    ctx->stats.recruit_reads += n;
    ctx->stats.h2d_bytes += bases + (n + 1) * 8 * (reads->off2 ? 2 : 1);
    ctx->stats.d2h_bytes += n * 4 + n * cap * 4;
    return LCTP_OK;
}

size_t lctp_sizeof_target_seqs(void) { return sizeof(lctp_target_seqs); }
size_t lctp_sizeof_reads(void) { return sizeof(lctp_reads); }

// host_solve.cu -- host side of the C ABI: context / locus handles, and the host-resident part of the
// reference's scheduler, restated in C++ above the device kernels:
//   src/solvers/solve.rs:52-84     truncate_ixs            :319-336 compare_two_likelihoods
//   src/solvers/solve.rs:425-480   discard_improbable_genotypes   :482-535 produce_result
//   src/solvers/solve.rs:926-981   solve                   :996-1093 MainWorker::{new,run}
//   src/solvers/solve.rs:732-773   Genotyping::to_json
//   src/model/distr_cache.rs:61-75 DistrCache::new  (+ src/math/distr/{bayes,nbinom}.rs, src/math/mod.rs)
//   src/ext/vec.rs:298-339         genotype enumeration    src/ext/rand.rs RNG seeding
// Third-party arithmetic restated from the published algorithms (not vendored by the reference):
// rand 0.10 shuffle / uniform ints, rand_xoshiro 0.8 jump polynomials, statrs 0.19 ln_gamma /
// beta_reg / StudentsT::cdf.  None of this touches oracle/: the oracle is an independent C restatement.
#include "common.cuh"

#include <algorithm>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <functional>

namespace lctp {

static thread_local std::string g_last_error;
static thread_local cudaStream_t g_alloc_stream = nullptr;
cudaStream_t current_alloc_stream() { return g_alloc_stream; }
void set_alloc_stream(cudaStream_t s) { g_alloc_stream = s; }

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

// ---------------------------------------------------------------- xoshiro256++ on the host -----

struct HostRng {
    uint64_t s[4];
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next_u64() {
        const uint64_t r = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
    uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    void jump_poly(const uint64_t poly[4]) {
        uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int i = 0; i < 4; i++)
            for (int b = 0; b < 64; b++) {
                if (poly[i] & (1ull << b)) { a0 ^= s[0]; a1 ^= s[1]; a2 ^= s[2]; a3 ^= s[3]; }
                next_u64();
            }
        s[0] = a0; s[1] = a1; s[2] = a2; s[3] = a3;
    }
    // rand UniformInt<u32>::sample_single_inclusive(0, range - 1)
    uint32_t below(uint32_t range) {
        if (range == 0) return next_u32();
        const uint64_t m = (uint64_t)next_u32() * range;
        uint32_t res = (uint32_t)(m >> 32);
        const uint32_t lo = (uint32_t)m;
        if (lo > 0u - range) {
            const uint32_t nh = (uint32_t)(((uint64_t)next_u32() * range) >> 32);
            res += (uint32_t)(lo + nh < lo);
        }
        return res;
    }
    uint64_t below64(uint64_t range) {
        if (range == 0) return next_u64();
        const unsigned __int128 m = (unsigned __int128)next_u64() * range;
        uint64_t res = (uint64_t)(m >> 64);
        const uint64_t lo = (uint64_t)m;
        if (lo > 0ull - range) {
            const uint64_t nh = (uint64_t)(((unsigned __int128)next_u64() * range) >> 64);
            res += (uint64_t)(lo + nh < lo);
        }
        return res;
    }
};

static const uint64_t kJump[4] = {0x180ec6d33cfd0abaull, 0xd5a61266f0c9392cull, 0xa9582618e03fc9aaull, 0x39abdc4529b1661cull};
static const uint64_t kLongJump[4] = {0x76e15d3efefdcbbfull, 0xc5004e441c522fb3ull, 0x77710069854ee241ull, 0x39109bb02acbe635ull};

// The xoshiro jump (2^128 steps) is a linear map J over GF(2).  MainWorker::new needs J^w * state for
// every worker w (solve.rs:1007-1018); with thousands of logical workers the polynomial method (256
// generator steps per jump) costs milliseconds, so J is tabulated once by byte slices:
// J*s = XOR_b tab[b][byte_b(s)].
struct JumpTable {
    std::vector<uint64_t> t;   // [32][256][4]
    JumpTable() : t((size_t)32 * 256 * 4, 0) {
        uint64_t col[256][4];
        for (int j = 0; j < 256; j++) {
            HostRng r; r.s[0] = r.s[1] = r.s[2] = r.s[3] = 0;
            r.s[j >> 6] = 1ull << (j & 63);
            r.jump_poly(kJump);
            for (int k = 0; k < 4; k++) col[j][k] = r.s[k];
        }
        for (int b = 0; b < 32; b++)
            for (int v = 0; v < 256; v++) {
                uint64_t *o = &t[((size_t)b * 256 + v) * 4];
                for (int bit = 0; bit < 8; bit++)
                    if (v & (1 << bit)) for (int k = 0; k < 4; k++) o[k] ^= col[b * 8 + bit][k];
            }
    }
    void apply(uint64_t s[4]) const {
        uint64_t a[4] = {0, 0, 0, 0};
        for (int b = 0; b < 32; b++) {
            const unsigned v = (unsigned)((s[b >> 3] >> ((b & 7) * 8)) & 0xFF);
            const uint64_t *o = &t[((size_t)b * 256 + v) * 4];
            a[0] ^= o[0]; a[1] ^= o[1]; a[2] ^= o[2]; a[3] ^= o[3];
        }
        s[0] = a[0]; s[1] = a[1]; s[2] = a[2]; s[3] = a[3];
    }
};
static const JumpTable &jump_table() { static const JumpTable jt; return jt; }

// SliceRandom::shuffle (rand >= 0.9): forward Fisher-Yates whose indices come from IncreasingUniform,
// i.e. one bounded u32 draw is split into several indices by div/mod.
static void shuffle_u64(HostRng &rng, uint64_t *v, size_t len) {
    if (len <= 1) return;
    if (len >= 0xFFFFFFFFull) {
        for (size_t i = 0; i < len; i++) std::swap(v[i], v[(size_t)rng.below64(i + 1)]);
        return;
    }
    uint32_t n = 0, chunk = 0;
    uint8_t remaining = 1;
    for (size_t i = 0; i < len; i++) {
        const uint32_t next_n = n + 1;
        uint8_t next_remaining;
        if (remaining >= 1) next_remaining = remaining - 1;
        else {
            uint32_t product = next_n, current = next_n + 1;
            for (;;) {
                const uint64_t pr = (uint64_t)product * current;
                if (pr > 0xFFFFFFFFull) break;
                product = (uint32_t)pr;
                current++;
            }
            chunk = rng.below(product);
            next_remaining = (uint8_t)(current - next_n - 1);
        }
        size_t index;
        if (next_remaining == 0) index = chunk;
        else { index = chunk % next_n; chunk /= next_n; }
        remaining = next_remaining;
        n = next_n;
        std::swap(v[i], v[index]);
    }
}

// ---------------------------------------------------------------- genotype enumeration ---------

static uint64_t choose(uint64_t n, uint64_t k) {
    if (k > n) return 0;
    const uint64_t r = std::min(k, n - k);
    uint64_t acc = 1;
    for (uint64_t v = 1; v <= r; v++) acc = acc * (n - v + 1) / v;
    return acc;
}

void genotype_tuple(uint32_t H, uint32_t p, const uint32_t *gt_tuples, uint64_t g, uint32_t *out) {
    if (gt_tuples) { std::copy(gt_tuples + g * p, gt_tuples + (g + 1) * p, out); return; }
    if (p == 2) {   // closed form: row i starts at i*H - i(i-1)/2
        const double Hd = (double)H + 0.5;
        uint64_t i = (uint64_t)std::floor(Hd - std::sqrt(std::max(0.0, Hd * Hd - 2.0 * (double)g)));
        auto row = [&](uint64_t r) { return r * H - r * (r - 1) / 2; };
        while (i > 0 && row(i) > g) i--;
        while (i + 1 < H && row(i + 1) <= g) i++;
        out[0] = (uint32_t)i; out[1] = (uint32_t)(i + (g - row(i)));
        return;
    }
    uint32_t lo = 0;
    for (uint32_t d = 0; d < p; d++) {
        const uint32_t rem = p - d - 1;
        for (uint32_t v = lo; v < H; v++) {
            const uint64_t cnt = rem == 0 ? 1 : choose((uint64_t)(H - v) + rem - 1, rem);
            if (g < cnt) { out[d] = v; lo = v; break; }
            g -= cnt;
        }
    }
}

// ---------------------------------------------------------------- special functions ------------

static const double kGammaR = 10.900511;
static const double kGammaDk[11] = {
    2.48574089138753565546e-5, 1.05142378581721974210, -3.45687097222016235469, 4.51227709466894823700,
    -2.98285225323576655721, 1.05639711577126713077, -1.95428773191645869583e-1, 1.70970543404441224307e-2,
    -5.71926117404305781283e-4, 4.63399473359905636708e-6, -2.71994908488607703910e-9};
static const double kLnPi = 1.1447298858494001741434273513530587116472948129153;
static const double kLn2SqrtEOverPi = 0.6207822376352452223455184457816472122518527279025978;

// statrs::function::gamma::ln_gamma (Lanczos g = 10.900511, n = 11)
static double ln_gamma(double x) {
    double s = kGammaDk[0];
    if (x < 0.5) {
        for (int i = 1; i < 11; i++) s += kGammaDk[i] / ((double)i - x);
        return kLnPi - std::log(std::sin(M_PI * x)) - std::log(s) - kLn2SqrtEOverPi -
               (0.5 - x) * std::log((0.5 - x + kGammaR) / M_E);
    }
    for (int i = 1; i < 11; i++) s += kGammaDk[i] / (x + (double)i - 1.0);
    return std::log(s) + kLn2SqrtEOverPi + (x - 0.5) * std::log((x - 0.5 + kGammaR) / M_E);
}

// statrs::function::beta::beta_reg (continued fraction)
static double beta_reg(double a, double b, double x) {
    if (!(a > 0.0) || !(b > 0.0) || !(x >= 0.0 && x <= 1.0)) return std::numeric_limits<double>::quiet_NaN();
    double bt = 0.0;
    if (!(std::fabs(x) < 1.1102230246251565e-15 || std::fabs(x - 1.0) <= 4.0 * std::numeric_limits<double>::epsilon()))
        bt = std::exp(ln_gamma(a + b) - ln_gamma(a) - ln_gamma(b) + a * std::log(x) + b * std::log(1.0 - x));
    const bool symm = x >= (a + 1.0) / (a + b + 2.0);
    const double eps = 1.1102230246251565e-16;
    const double fpmin = std::numeric_limits<double>::min() / eps;
    if (symm) { std::swap(a, b); x = 1.0 - x; }
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (std::fabs(d) < fpmin) d = fpmin;
    d = 1.0 / d;
    double hh = d;
    for (int mi = 1; mi < 141; mi++) {
        const double m = mi, m2 = m * 2.0;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c; if (std::fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        hh = hh * d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (std::fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c; if (std::fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        const double del = d * c;
        hh *= del;
        if (std::fabs(del - 1.0) <= eps) break;
    }
    return symm ? 1.0 - bt * hh / a : bt * hh / a;
}

static double students_t_cdf(double x, double freedom) {
    if (std::isinf(freedom)) return 0.5 * std::erfc(-x / M_SQRT2);
    const double hh = freedom / (freedom + x * x);
    const double ib = 0.5 * beta_reg(freedom / 2.0, 0.5, hh);
    return x <= 0.0 ? ib : 1.0 - ib;
}

// Ln::add / Ln::sum / Ln::sum_init (src/math/mod.rs:28-94)
static double ln_add(double a, double b) {
    const double ninf = -std::numeric_limits<double>::infinity();
    if (a >= b) return b == ninf ? a : b + std::log1p(std::exp(a - b));
    return a == ninf ? b : a + std::log1p(std::exp(b - a));
}
static double ln_sum(const double *v, size_t n) {
    if (n == 0) return -std::numeric_limits<double>::infinity();
    if (n == 1) return v[0];
    double m = -std::numeric_limits<double>::infinity();
    for (size_t i = 0; i < n; i++) m = std::fmax(m, v[i]);
    if (std::isinf(m)) return m;
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s += std::exp(v[i] - m);
    return m + std::log(s);
}
static double ln_sum_init(const double *v, size_t n, double init) {
    if (n == 0) return init;
    if (n == 1) return ln_add(init, v[0]);
    double m = init;
    for (size_t i = 0; i < n; i++) m = std::fmax(m, v[i]);
    if (std::isinf(m)) return m;
    double s = std::exp(init - m);
    for (size_t i = 0; i < n; i++) s += std::exp(v[i] - m);
    return m + std::log(s);
}

struct NBinom {   // src/math/distr/nbinom.rs:24-42,127-132
    double n, p, lnq, lnpmf_const;
    NBinom() = default;
    NBinom(double n_, double p_) : n(n_), p(p_), lnq(std::log1p(-p_)), lnpmf_const(n_ * std::log(p_) - ln_gamma(n_)) {}
    double ln_pmf(uint32_t k) const {
        const double x = k;
        return lnpmf_const + ln_gamma(n + x) - ln_gamma(x + 1.0) + x * lnq;
    }
};

// ---------------------------------------------------------------- sorting helpers --------------

static inline int64_t total_key(double v) {
    int64_t b;
    std::memcpy(&b, &v, 8);
    return b ^ (int64_t)(((uint64_t)(b >> 63)) >> 1);
}

// Descending by key with f64::total_cmp; equal keys keep their current order (the reference uses
// sort_unstable_by, whose tie order is unspecified -- see DESIGN.md "unpinned behaviour").
static void sort_desc_stable(uint64_t *ixs, size_t n, const double *key) {
    // (key, current position) pairs sorted with a total order = the stable order, without an indirect load per compare
    struct Kp { int64_t k; uint64_t pos, ix; };
    static thread_local std::vector<Kp> kv;
    kv.resize(n);
    for (size_t q = 0; q < n; q++) kv[q] = Kp{total_key(key[ixs[q]]), q, ixs[q]};
    std::sort(kv.begin(), kv.end(), [](const Kp &a, const Kp &b) { return a.k != b.k ? a.k > b.k : a.pos < b.pos; });
    for (size_t q = 0; q < n; q++) ixs[q] = kv[q].ix;
}

// compare_two_likelihoods (src/solvers/solve.rs:319-336) with the t-tests of src/math/mod.rs:180-220
static double compare_two(double m1, double v1, uint16_t a1, double m2, double v2, uint16_t a2) {
    const double simple_norm = m1 - ln_add(m1, m2);
    if (std::isnormal(v1) && std::isnormal(v2)) {
        double t_pval;
        if (a1 == a2) {
            const double n = a1, var_sum = v1 + v2;
            const double t_stat = (m1 - m2) * std::sqrt(n / var_sum);
            const double freedom = (n - 1.0) * var_sum * var_sum / (v1 * v1 + v2 * v2);
            t_pval = students_t_cdf(t_stat, freedom);
        } else {
            const double n1 = a1, n2 = a2, nv1 = v1 / n1, nv2 = v2 / n2, sv = nv1 + nv2;
            const double t_stat = (m1 - m2) / std::sqrt(sv);
            const double freedom = sv * sv / (nv1 * nv1 / (n1 - 1.0) + nv2 * nv2 / (n2 - 1.0));
            t_pval = students_t_cdf(t_stat, freedom);
        }
        return std::fmax(simple_norm, std::log(t_pval));
    }
    return simple_norm;
}

}  // namespace lctp

using namespace lctp;

// =============================================================================== C ABI =========

extern "C" {

const char *lctp_version(void) { return "lctp 0.1.0 (sm_100a; restates locityper v1.7.2 genotype evaluation)"; }
const char *lctp_last_error(void) { return g_last_error.c_str(); }
size_t lctp_sizeof_locus(void) { return sizeof(lctp_locus); }
size_t lctp_sizeof_stage(void) { return sizeof(lctp_stage); }
size_t lctp_sizeof_result(void) { return sizeof(lctp_result); }

int lctp_init(const lctp_device_cfg *cfg, lctp_ctx **out) {
    if (!out) { set_error("lctp_init: out is NULL"); return LCTP_E_INVALID; }
    *out = nullptr;
    int dev = cfg ? cfg->device : 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("lctp_init: no CUDA device available (%s); this path has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return LCTP_E_CUDA;
    }
    if (dev < 0 || dev >= count) { set_error("lctp_init: device %d out of range (0..%d)", dev, count - 1); return LCTP_E_INVALID; }
    LCTP_CUDA_CHECK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    LCTP_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error("lctp_init: device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
        return LCTP_E_CUDA;
    }
    lctp_ctx *ctx = new lctp_ctx();
    ctx->device = dev;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    ctx->max_resident_workers = cfg ? cfg->max_resident_workers : 0;
    if (cfg && cfg->stream) { ctx->stream = (cudaStream_t)cfg->stream; ctx->own_stream = false; }
    else {
        cudaError_t se = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (se != cudaSuccess) { set_error("lctp_init: cudaStreamCreate failed: %s", cudaGetErrorString(se)); delete ctx; return LCTP_E_CUDA; }
        ctx->own_stream = true;
    }
    {   // keep freed blocks in the pool instead of returning them to the driver at every sync
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    set_alloc_stream(ctx->stream);
    for (int i = 0; i < 4; i++) {
        cudaError_t ee = cudaEventCreate(&ctx->ev[i]);
        if (ee != cudaSuccess) { set_error("lctp_init: cudaEventCreate failed: %s", cudaGetErrorString(ee)); delete ctx; return LCTP_E_CUDA; }
    }
    *out = ctx;
    return LCTP_OK;
}

void lctp_destroy(lctp_ctx *ctx) {
    lctp_debug_close(ctx);
    if (!ctx) return;
    lctp::free_local_selector(ctx);
    cudaSetDevice(ctx->device);
    set_alloc_stream(ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 4; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    cudaStream_t st = ctx->stream;
    const bool own = ctx->own_stream;
    delete ctx;                       // frees the scratch buffers on the stream
    cudaStreamSynchronize(st);
    if (own) cudaStreamDestroy(st);
}

uint64_t lctp_launch_count(const lctp_ctx *ctx) { return ctx ? ctx->launches : 0; }

int lctp_sync(lctp_ctx *ctx) {
    if (!ctx) { set_error("lctp_sync: NULL context"); return LCTP_E_INVALID; }
    LCTP_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return LCTP_OK;
}

int lctp_get_stats(lctp_ctx *ctx, lctp_stats *out, int reset) {
    if (!ctx || !out) { set_error("lctp_get_stats: NULL argument"); return LCTP_E_INVALID; }
    *out = ctx->stats;
    if (reset) ctx->stats = lctp_stats{};
    return LCTP_OK;
}

size_t lctp_sizeof_mates(void) { return sizeof(lctp_mates); }
size_t lctp_sizeof_alns(void) { return sizeof(lctp_alns); }

int lctp_rescore_alignments(lctp_ctx *ctx, const lctp_alns *in, double *ln_prob, uint32_t *edit, uint32_t *read_len,
                            uint8_t *save) {
    if (!ctx || !in) { set_error("lctp_rescore_alignments: NULL argument"); return LCTP_E_INVALID; }
    if (in->n_alns && (!in->cigar_off || !in->cigar_ops || !in->aln_start || !in->aln_end || !in->contig_len ||
                       !in->passable_dist || !ln_prob || !edit || !read_len || !save)) {
        set_error("lctp_rescore_alignments: NULL array");
        return LCTP_E_INVALID;
    }
    if (in->n_alns && in->cigar_off[in->n_alns] < in->cigar_off[0]) { set_error("lctp_rescore_alignments: cigar_off not ascending"); return LCTP_E_INVALID; }
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    set_alloc_stream(ctx->stream);
    return rescore_alignments(ctx, in, ln_prob, edit, read_len, save);
}

size_t lctp_sizeof_read_ends(void) { return sizeof(lctp_read_ends); }

// count_to_prob, src/model/bam.rs:54-66 (f32 arithmetic throughout, f32::log10 / round / min as the reference)
int lctp_counts_to_prob(const uint16_t *counts, uint64_t n, uint16_t attempts, float *prob, uint8_t *mapq) {
    if (n && (!counts || !prob || !mapq)) { set_error("lctp_counts_to_prob: NULL array"); return LCTP_E_INVALID; }
    for (uint64_t i = 0; i < n; i++) {
        const uint16_t c = counts[i];
        if (c == 0) { prob[i] = 0.0f; mapq[i] = 0; }
        else if (c == attempts) { prob[i] = 1.0f; mapq[i] = 60; }
        else if (c > attempts) { set_error("lctp_counts_to_prob: count %u of %u attempts", c, attempts); return LCTP_E_INVALID; }
        else {
            const float p = (float)c / (float)attempts;
            const float q = std::fmin(std::round(-10.0f * std::log10(1.0f - p)), 60.0f);
            prob[i] = p;
            mapq[i] = (uint8_t)q;
        }
    }
    return LCTP_OK;
}

int lctp_collect_read_ends(lctp_ctx *ctx, const lctp_read_ends *in, double *ln_prob, uint32_t *edit, uint32_t *read_len,
                           uint8_t *ok, uint32_t *best_edit, double *weight_factor, uint32_t *thr_dist,
                           uint32_t *pass_dist, uint32_t *n_kept, uint32_t *kept_rec) {
    if (!ctx || !in) { set_error("lctp_collect_read_ends: NULL argument"); return LCTP_E_INVALID; }
    const lctp_alns &a = in->alns;
    if (a.n_alns && (!a.cigar_off || !a.cigar_ops || !a.aln_start || !a.aln_end || !a.contig_len || !in->grp_off ||
                     !in->rec_contig || !in->grp_read_end || !in->grp_read_len || !in->grp_good_dist ||
                     !in->grp_passable_dist || !in->grp_neighb_complexity || !ln_prob || !edit || !read_len || !ok ||
                     !best_edit || !weight_factor || !thr_dist || !pass_dist || !n_kept || !kept_rec)) {
        set_error("lctp_collect_read_ends: NULL array");
        return LCTP_E_INVALID;
    }
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    set_alloc_stream(ctx->stream);
    return collect_read_ends(ctx, in, ln_prob, edit, read_len, ok, best_edit, weight_factor, thr_dist, pass_dist, n_kept,
                             kept_rec);
}

int lctp_pair_alignments(lctp_ctx *ctx, const lctp_mates *in, uint64_t cap, uint64_t *pa_off, uint32_t *pa_contig,
                         double *pa_ln_prob, uint32_t *pa_mid1, uint32_t *pa_mid2, double *unmapped_prob,
                         uint64_t *n_out) {
    if (!ctx || !in || !pa_off || !unmapped_prob || (cap && (!pa_contig || !pa_ln_prob || !pa_mid1 || !pa_mid2))) {
        set_error("lctp_pair_alignments: NULL argument");
        return LCTP_E_INVALID;
    }
    lctp_pairs_h *p = nullptr;
    int rc = lctp_pair_alignments_dev(ctx, in, &p, n_out);
    if (rc) return rc;
    rc = pairs_fetch(p, cap, pa_off, pa_contig, pa_ln_prob, pa_mid1, pa_mid2, unmapped_prob);
    lctp_pairs_free(p);
    return rc;
}

int lctp_pair_alignments_dev(lctp_ctx *ctx, const lctp_mates *in, lctp_pairs_h **out, uint64_t *n_out) {
    if (!ctx || !in || !out) { set_error("lctp_pair_alignments: NULL argument"); return LCTP_E_INVALID; }
    *out = nullptr;
    if (!in->single_end && (!in->ins_ln_pmf || in->ins_len == 0)) {
        set_error("lctp_pair_alignments: NULL insert-size table");
        return LCTP_E_INVALID;
    }
    if (in->n_reads && in->ma_off && in->ma_off[in->n_reads] &&
        (!in->ma_contig || !in->ma_flags || !in->ma_start || !in->ma_end || !in->ma_ln_prob)) {
        set_error("lctp_pair_alignments: NULL mate array");
        return LCTP_E_INVALID;
    }
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    set_alloc_stream(ctx->stream);
    lctp_pairs_h *p = new lctp_pairs_h();
    const int rc = pair_alignments_dev(ctx, in, p);
    if (rc) { delete p; return rc; }
    if (n_out) *n_out = p->n_pairs;
    *out = p;
    return LCTP_OK;
}

int lctp_pairs_fetch(lctp_pairs_h *p, uint64_t cap, uint64_t *pa_off, uint32_t *pa_contig, double *pa_ln_prob,
                     uint32_t *pa_mid1, uint32_t *pa_mid2, double *unmapped_prob) {
    if (!p || !pa_off || !unmapped_prob || (cap && (!pa_contig || !pa_ln_prob || !pa_mid1 || !pa_mid2))) {
        set_error("lctp_pairs_fetch: NULL argument");
        return LCTP_E_INVALID;
    }
    LCTP_CUDA_CHECK(cudaSetDevice(p->ctx->device));
    return pairs_fetch(p, cap, pa_off, pa_contig, pa_ln_prob, pa_mid1, pa_mid2, unmapped_prob);
}

uint64_t lctp_pairs_count(const lctp_pairs_h *p) { return p ? p->n_pairs : 0; }

void lctp_pairs_free(lctp_pairs_h *p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    set_alloc_stream(p->ctx->stream);
    delete p;
}

int lctp_locus_upload_pairs(lctp_ctx *ctx, const lctp_locus *in, lctp_pairs_h *pairs, lctp_locus_h **out) {
    if (!ctx || !in || !pairs || !out) { set_error("lctp_locus_upload_pairs: NULL argument"); return LCTP_E_INVALID; }
    *out = nullptr;
    if (pairs->ctx != ctx || pairs->n_reads != in->n_reads) {
        set_error("lctp_locus_upload_pairs: pair alignments of another context / read count");
        return LCTP_E_INVALID;
    }
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    set_alloc_stream(ctx->stream);
    lctp_locus_h *h = new lctp_locus_h();
    int rc = upload_locus(ctx, in, h, pairs);
    if (rc != LCTP_OK) { delete h; return rc; }
    *out = h;
    return LCTP_OK;
}

int lctp_measure_fp64_rate(lctp_ctx *ctx, double *lane_inst_per_s) {
    if (!ctx || !lane_inst_per_s) { set_error("lctp_measure_fp64_rate: NULL argument"); return LCTP_E_INVALID; }
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    set_alloc_stream(ctx->stream);
    return measure_fp64_rate(ctx, lane_inst_per_s);
}

int lctp_prefilter_plan_check(uint32_t n_haps, uint32_t n_sm, const uint32_t *pattern, uint32_t n_pattern,
                              uint64_t g_begin, uint64_t g_end, uint32_t *n_regions, uint32_t *load,
                              uint32_t *pattern_out) {
    const uint64_t G = (uint64_t)n_haps * (n_haps + 1) / 2;
    if (n_haps == 0 || n_sm == 0 || g_begin >= g_end || g_end > G) {
        set_error("lctp_prefilter_plan_check: invalid arguments");
        return LCTP_E_INVALID;
    }
    return prefilter_plan_check(n_haps, n_sm, pattern, n_pattern, g_begin, g_end, n_regions, load, pattern_out);
}

int lctp_locus_upload(lctp_ctx *ctx, const lctp_locus *in, lctp_locus_h **out) {
    if (!ctx || !in || !out) { set_error("lctp_locus_upload: NULL argument"); return LCTP_E_INVALID; }
    *out = nullptr;
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    set_alloc_stream(ctx->stream);
    lctp_locus_h *h = new lctp_locus_h();
    int rc = upload_locus(ctx, in, h, nullptr);
    if (rc != LCTP_OK) { delete h; return rc; }
    *out = h;
    return LCTP_OK;
}

void lctp_locus_free(lctp_locus_h *h) {
    if (!h) return;
    if (h->ctx) { cudaSetDevice(h->ctx->device); set_alloc_stream(h->ctx->stream); }
    delete h;
}

int lctp_best_aln_matrix(lctp_locus_h *h, double *m_out) {
    if (!h || !m_out) { set_error("lctp_best_aln_matrix: NULL argument"); return LCTP_E_INVALID; }
    const uint32_t H = h->dev.H, R = h->dev.R, Hpad = h->dev.Hpad;
    std::vector<double> mt((size_t)R * Hpad);
    LCTP_CUDA_CHECK(cudaMemcpyAsync(mt.data(), h->Mt.p, mt.size() * 8, cudaMemcpyDeviceToHost, h->ctx->stream));
    LCTP_CUDA_CHECK(cudaStreamSynchronize(h->ctx->stream));
    for (uint32_t k = 0; k < H; k++)
        for (uint32_t r = 0; r < R; r++) m_out[(size_t)k * R + r] = mt[(size_t)r * Hpad + k];
    return LCTP_OK;
}

int lctp_prefilter_scores(lctp_locus_h *h, uint64_t g_begin, uint64_t g_end, double *scores_out) {
    if (!h) { set_error("lctp_prefilter_scores: NULL handle"); return LCTP_E_INVALID; }
    if (g_begin > g_end || g_end > h->dev.G) { set_error("lctp_prefilter_scores: bad range"); return LCTP_E_INVALID; }
    LCTP_CUDA_CHECK(cudaSetDevice(h->ctx->device));
    LCTP_CUDA_CHECK(cudaEventRecord(h->ctx->ev[0], h->ctx->stream));
    int rc = launch_prefilter(h, g_begin, g_end, h->scores.p);
    if (rc) return rc;
    LCTP_CUDA_CHECK(cudaEventRecord(h->ctx->ev[1], h->ctx->stream));
    if (scores_out && g_end > g_begin) {
        LCTP_CUDA_CHECK(cudaMemcpyAsync(scores_out, h->scores.p + g_begin, (g_end - g_begin) * 8,
                                        cudaMemcpyDeviceToHost, h->ctx->stream));
        h->ctx->stats.d2h_bytes += (g_end - g_begin) * 8;
    }
    LCTP_CUDA_CHECK(cudaStreamSynchronize(h->ctx->stream));
    float ms = 0.f;
    LCTP_CUDA_CHECK(cudaEventElapsedTime(&ms, h->ctx->ev[0], h->ctx->ev[1]));
    h->ctx->stats.prefilter_ms += ms;
    h->ctx->stats.prefilter_launches += 1;
    h->ctx->stats.prefilter_genotypes += g_end - g_begin;
    return LCTP_OK;
}

// k-th largest (1-based) of key[0..n), all keys in [kmin, kmax]: one histogram pass over linear bins of the actual
// key range, then nth_element inside the one bin that holds the answer (std::nth_element over all n costs 6 ms at
// n = 500,500; this is two streaming passes).
static int64_t kth_largest_key(const int64_t *key, size_t n, size_t k, int64_t kmin, int64_t kmax) {
    int bits = 8;                                                        // ~n/8 bins, 256 .. 65,536
    while (bits < 16 && (size_t(1) << (bits + 3)) < n) bits++;
    const uint64_t range = (uint64_t)kmax - (uint64_t)kmin;              // exact in unsigned arithmetic
    int shift = 0;
    while ((range >> shift) >= (1ull << bits)) shift++;
    static thread_local std::vector<uint32_t> hist;
    hist.assign((size_t)(range >> shift) + 1, 0);
    for (size_t q = 0; q < n; q++) hist[((uint64_t)key[q] - (uint64_t)kmin) >> shift]++;
    size_t above = 0, bin = hist.size() - 1;
    for (;; bin--) {
        if (above + hist[bin] >= k) break;
        above += hist[bin];
        if (bin == 0) break;                                             // k <= n guarantees we stop before
    }
    static thread_local std::vector<int64_t> cand;
    cand.clear();
    for (size_t q = 0; q < n; q++)
        if ((((uint64_t)key[q] - (uint64_t)kmin) >> shift) == bin) cand.push_back(key[q]);
    const size_t r = k - above;                                          // r-th largest inside the bin, 1-based
    std::nth_element(cand.begin(), cand.begin() + (r - 1), cand.end(), std::greater<int64_t>());
    return cand[r - 1];
}

static inline double key_to_double(int64_t k) {         // inverse of total_key (the transform is an involution)
    const int64_t b = k ^ (int64_t)(((uint64_t)(k >> 63)) >> 1);
    double v;
    std::memcpy(&v, &b, 8);
    return v;
}

size_t lctp_truncate_ixs(uint64_t *ixs, size_t n, const double *scores, double filt_diff, size_t min_size,
                         size_t threads) {
    if (n == 0) return 0;
    // Reference: sort all, then cut (solve.rs:60-81).  Equivalent and much cheaper for large n: find the cut first,
    // gather the survivors (keeping their order), sort only those.  The list is read ONCE into contiguous
    // (value, total_cmp key) arrays; every later pass streams those (G = 500,500 at the KIR scale).
    // scratch that outlives the call: fresh 8 MB vectors cost more in page faults than the passes over them
    static thread_local std::vector<double> val_tl;
    static thread_local std::vector<int64_t> key_tl;
    if (val_tl.size() < n) { val_tl.resize(n); key_tl.resize(n); }
    double *val = val_tl.data();
    int64_t *key = key_tl.data();
    bool ascending = true;
    int64_t kbest, kworst;
    {
        const double v0 = scores[ixs[0]];
        val[0] = v0; key[0] = kbest = kworst = total_key(v0);
        uint64_t prev = ixs[0];
        for (size_t q = 1; q < n; q++) {
            const uint64_t ix = ixs[q];
            ascending &= prev < ix;
            prev = ix;
            const double v = scores[ix];
            const int64_t k = total_key(v);
            val[q] = v; key[q] = k;
            kbest = std::max(kbest, k); kworst = std::min(kworst, k);
        }
    }
    const double best = key_to_double(kbest), worst = key_to_double(kworst);
    double thresh = best - filt_diff;
    auto count_ge = [&](double th) { size_t c = 0; for (size_t q = 0; q < n; q++) c += val[q] >= th; return c; };
    size_t m, n_ge = n;
    if (min_size >= n || worst >= thresh) m = n;
    else {
        m = count_ge(thresh);
        if (m < min_size) {
            thresh = key_to_double(kth_largest_key(key, n, min_size, kworst, kbest));
            m = count_ge(thresh);
        }
        n_ge = m;                                  // #{score >= thresh} for the final threshold
        m = std::min(std::max(m, threads), n);
    }
    if (m == n || !ascending) {   // nothing to gain, or arbitrary input order: literal restatement
        sort_desc_stable(ixs, n, scores);
        return m;
    }
    // the m best under (score desc, input order asc): everything >= the m-th best key, ties by position
    std::vector<uint64_t> keep;
    keep.reserve(m + 16);
    if (n_ge >= m) {
        // m was not raised by `threads`: survivors are exactly {score >= thresh} (m == n_ge)
        for (size_t q = 0; q < n; q++) if (val[q] >= thresh) keep.push_back(ixs[q]);
    } else {
        // raised to `threads`: take the m best overall (key desc, id asc)
        // = every id with key > k_m, plus the first (in id order = list order) m - #above of those with key == k_m
        const int64_t km = kth_largest_key(key, n, m, kworst, kbest);
        size_t above = 0;
        for (size_t q = 0; q < n; q++) above += key[q] > km;
        size_t ties_left = m - above;
        for (size_t q = 0; q < n; q++) {
            if (key[q] > km) keep.push_back(ixs[q]);
            else if (key[q] == km && ties_left) { keep.push_back(ixs[q]); ties_left--; }
        }
    }
    sort_desc_stable(keep.data(), keep.size(), scores);
    std::copy(keep.begin(), keep.end(), ixs);
    return keep.size();
}

int lctp_prefilter(lctp_locus_h *h, uint64_t *ixs, size_t n, size_t min_size, size_t threads, size_t *out_n,
                   double *scores_out) {
    if (!h || !ixs || !out_n || n == 0) { set_error("lctp_prefilter: NULL/empty argument"); return LCTP_E_INVALID; }
    const uint64_t G = h->dev.G;
    // The usual call (solve.rs:939-944: the complete list 0..G, scores not wanted on the host): candidates are selected on
    // the device, only they cross to the host (csrc/dist.cu; the same code path as one rank of the multi-GPU split).
    if (!scores_out && n == G && n >= 4096 && !getenv("LCTP_HOST_PREFILTER")) {
        bool full = true;
        for (size_t q = 0; q < n && full; q++) full = ixs[q] == q;
        if (full) return lctp_dist_prefilter(lctp::local_selector(h->ctx), h, min_size, threads, ixs, n, out_n);
    }
    std::vector<double> scores(G, -std::numeric_limits<double>::infinity());
    // an arbitrary subset is scored over its covering range.
    uint64_t lo = G, hi = 0;
    for (size_t q = 0; q < n; q++) {
        if (ixs[q] >= G) { set_error("lctp_prefilter: genotype id out of range"); return LCTP_E_INVALID; }
        lo = std::min(lo, ixs[q]); hi = std::max(hi, ixs[q] + 1);
    }
    std::vector<double> part(hi - lo);
    int rc = lctp_prefilter_scores(h, lo, hi, part.data());
    if (rc) return rc;
    for (size_t q = 0; q < n; q++) scores[ixs[q]] = part[ixs[q] - lo];
    *out_n = lctp_truncate_ixs(ixs, n, scores.data(), h->host.filt_diff, min_size, threads);
    if (scores_out) std::copy(scores.begin(), scores.end(), scores_out);
    return LCTP_OK;
}

int lctp_solve_stage(lctp_locus_h *h, const lctp_stage *st, const uint64_t *worker_ixs, const uint64_t *worker_off,
                     size_t n_workers, uint64_t *worker_rng, double *lik_mean, double *lik_var, double *liks,
                     uint64_t *counts_off, uint16_t *counts, uint64_t counts_cap, uint64_t *n_alns_out,
                     uint64_t *iters_out) {
    if (!h) { set_error("lctp_solve_stage: NULL handle"); return LCTP_E_INVALID; }
    LCTP_CUDA_CHECK(cudaSetDevice(h->ctx->device));
    set_alloc_stream(h->ctx->stream);
    return launch_stage(h, st, worker_ixs, worker_off, n_workers, worker_rng, lik_mean, lik_var, liks, counts_off,
                        counts, counts_cap, n_alns_out, iters_out);
}

void lctp_rng_seed_from_u64(uint64_t state[4], uint64_t seed) {
    uint64_t x = seed;   // SplitMix64 fill (rand_xoshiro seed_from_u64)
    for (int i = 0; i < 4; i++) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        state[i] = z ^ (z >> 31);
    }
}
void lctp_rng_jump(uint64_t state[4]) { jump_table().apply(state); }
void lctp_rng_worker_streams(uint64_t state[4], size_t threads, uint64_t *out) {
    const JumpTable &jt = jump_table();
    for (size_t w = 0; w < threads; w++) { std::memcpy(out + 4 * w, state, 32); jt.apply(state); }
}
void lctp_rng_long_jump(uint64_t state[4]) {
    HostRng r; std::memcpy(r.s, state, 32); r.jump_poly(kLongJump); std::memcpy(state, r.s, 32);
}

size_t lctp_plan_stage(uint64_t rng[4], uint64_t *ixs, size_t n, size_t threads, uint64_t *worker_off) {
    HostRng r; std::memcpy(r.s, rng, 32);
    shuffle_u64(r, ixs, n);                       // solve.rs:1051
    std::memcpy(rng, r.s, 32);
    size_t start = 0, nw = 0;
    worker_off[0] = 0;
    for (size_t i = 0; i < threads; i++) {        // solve.rs:1052-1062
        if (start == n) break;
        const size_t rem_workers = threads - i;
        start += (n - start + rem_workers - 1) / rem_workers;   // fast_ceil_div
        worker_off[++nw] = start;
    }
    return nw;
}

double lctp_compare_two_likelihoods(double m1, double v1, uint16_t a1, double m2, double v2, uint16_t a2) {
    return compare_two(m1, v1, a1, m2, v2, a2);
}

size_t lctp_discard_improbable(uint64_t *ixs, size_t n, const double *lik_mean, const double *lik_var,
                               const uint16_t *attempts, double prob_thresh, size_t out_size, size_t threads) {
    out_size = std::max(out_size, threads);
    if (prob_thresh == -std::numeric_limits<double>::infinity() || out_size >= n) return n;
    sort_desc_stable(ixs, n, lik_mean);
    const uint64_t best = ixs[0];
    size_t m = out_size;
    if (out_size <= 500) {                       // SOPHISTICATED_COUNT
        uint32_t dropped = 0;
        for (size_t q = out_size; q < n; q++) {
            const uint64_t ix = ixs[q];
            const double ln_pval = compare_two(lik_mean[ix], lik_var[ix], attempts[ix], lik_mean[best], lik_var[best], attempts[best]);
            if (ln_pval >= prob_thresh) ixs[m++] = ix;
            else if (++dropped >= 5) break;      // STOP_COUNT
        }
    }
    return m;
}

int lctp_build_depth_table(const double *nb_n, const double *nb_p, int is_paired, const double *alt_cn,
                           size_t n_alt, uint32_t k_cols, double *out) {
    if (!nb_n || !nb_p || !out || (n_alt && !alt_cn)) { set_error("lctp_build_depth_table: NULL argument"); return LCTP_E_INVALID; }
    if (n_alt >= 16) {       // BayesCalc::new: assert!(alternatives.len() < N_ALTS), src/math/distr/bayes.rs:16
        set_error("lctp_build_depth_table: %zu alternative copy numbers; the reference allows at most 15", n_alt);
        return LCTP_E_INVALID;
    }
    const double mul_coef = is_paired ? 2.0 : 1.0;           // distr_cache.rs:66
    for (int gc = 0; gc < LCTP_GC_BINS; gc++) {
        const NBinom cn1(nb_n[gc] * mul_coef, nb_p[gc]);     // NBinom::mul, nbinom.rs:68-70
        NBinom alts[16];
        for (size_t a = 0; a < n_alt; a++) alts[a] = NBinom(cn1.n * alt_cn[a], cn1.p);
        for (uint32_t k = 0; k < k_cols; k++) {
            const double null_prob = cn1.ln_pmf(k);          // BayesCalc::ln_pmf, bayes.rs:26-35
            double probs[16];
            for (size_t a = 0; a < n_alt; a++) probs[a] = alts[a].ln_pmf(k);
            out[(size_t)gc * k_cols + k] = null_prob - ln_sum_init(probs, n_alt, null_prob);
        }
    }
    return LCTP_OK;
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int lctp_produce_result(lctp_locus_h *h, uint64_t *ixs, size_t n, const double *lik_mean, const double *lik_var,
                        const uint16_t *attempts, lctp_result *res) {
    if (!h || !ixs || !lik_mean || !lik_var || !attempts || !res || n == 0) {
        set_error("lctp_produce_result: invalid argument");
        return LCTP_E_INVALID;
    }
    for (size_t q = 0; q < n; q++)
        if (ixs[q] >= h->dev.G) { set_error("lctp_produce_result: genotype id out of range"); return LCTP_E_INVALID; }
    std::memset(res, 0, sizeof(*res));
    // produce_result, solve.rs:482-535
    const double THRESH = -11.512925464970229;
    const size_t min_output = std::max<size_t>(4, h->host.out_bams);
    const double thresh_prob = std::fmin(THRESH, h->host.prob_thresh);
    sort_desc_stable(ixs, n, lik_mean);
    size_t m = std::min<size_t>(n, LCTP_MAX_OUT);
    double ln_probs[LCTP_MAX_OUT];
    std::fill(ln_probs, ln_probs + LCTP_MAX_OUT, 0.0);
    for (size_t i = 0; i < m; i++) {
        const uint64_t u = ixs[i];
        const size_t m_loop = m;
        for (size_t j = i + 1; j < m_loop; j++) {
            const uint64_t v = ixs[j];
            const double prob_j = compare_two(lik_mean[v], lik_var[v], attempts[v], lik_mean[u], lik_var[u], attempts[u]);
            if (i == 0 && j >= min_output && prob_j < thresh_prob) { m = j; break; }
            ln_probs[i] += std::log1p(-std::exp(prob_j));
            ln_probs[j] += prob_j;
        }
        res->gt_ix[i] = u; res->lik_mean[i] = lik_mean[u]; res->lik_var[i] = lik_var[u]; res->attempts[i] = attempts[u];
    }
    const double norm = ln_sum(ln_probs, m);
    for (size_t k = 0; k < m; k++) { ln_probs[k] -= norm; res->ln_prob[k] = ln_probs[k]; }
    const double others = m >= 1 ? ln_sum(ln_probs + 1, m - 1) : -std::numeric_limits<double>::infinity();
    res->quality = std::fmin(-10.0 * (others * 0.4342944819032518277), 1e9);   // Phred::from_ln_prob
    res->n_out = m;
    res->total_reads = h->dev.R;
    // check_first_prob (solve.rs:637-645), check_num_of_reads (:649-678)
    const double lp0 = res->ln_prob[0];
    res->warn_no_probable = (std::isnan(lp0) || lp0 < -2.0 * 2.302585092994045684) ? 1 : 0;
    const uint32_t p = h->dev.p, nr = h->dev.R;
    if (nr < p) res->warn_few_reads = 1;
    else if (p > 1 && nr < p * 10) {
        const double k = p, nn = nr;
        if (std::exp(std::log(k - 1.0) * nn - std::log(k) * (nn - 1.0)) > 0.1) res->warn_few_reads = 1;
    }
    // count_unexplained_reads (solve.rs:719-729): best_at_contig over the called genotype's contigs = the
    // matrix column of those contigs, which lives on the device (Mt).
    {
        uint32_t ids[LCTP_MAX_PLOIDY];
        genotype_tuple(h->dev.H, p, h->gt_tuples_host.empty() ? nullptr : h->gt_tuples_host.data(), res->gt_ix[0], ids);
        std::vector<double> col((size_t)nr * p);
        for (uint32_t k = 0; k < p; k++)
            LCTP_CUDA_CHECK(cudaMemcpy2DAsync(col.data() + (size_t)k * nr, 8, h->Mt.p + ids[k], (size_t)h->dev.Hpad * 8, 8, nr,
                                              cudaMemcpyDeviceToHost, h->ctx->stream));
        LCTP_CUDA_CHECK(cudaStreamSynchronize(h->ctx->stream));
        uint32_t unexpl = 0;
        for (uint32_t r = 0; r < nr; r++) {
            double best = -std::numeric_limits<double>::infinity();
            for (uint32_t k = 0; k < p; k++) best = std::fmax(best, col[(size_t)k * nr + r]);
            unexpl += best < h->unmapped_host[r] + 1e-8 ? 1u : 0u;
        }
        res->unexpl_reads = unexpl;
    }
    return LCTP_OK;
}

// ---- output side: the per-locus debug tables of the reference -------------------------------

uint32_t lctp_locus_wmax(const lctp_locus_h *h) { return h ? 2 + h->dev.p * h->max_n_windows : 0; }

int lctp_solve_stage_dbg(lctp_locus_h *h, const lctp_stage *st, const uint64_t *worker_ixs, const uint64_t *worker_off,
                         size_t n_workers, uint64_t *worker_rng, double *lik_mean, double *lik_var, double *liks,
                         uint64_t *counts_off, uint16_t *counts, uint64_t counts_cap, uint64_t *n_alns_out,
                         uint64_t *iters_out, const lctp_stage_debug *dbg) {
    if (!h) { set_error("lctp_solve_stage: NULL handle"); return LCTP_E_INVALID; }
    LCTP_CUDA_CHECK(cudaSetDevice(h->ctx->device));
    set_alloc_stream(h->ctx->stream);
    h->ctx->dbg_req = dbg;
    const int rc = launch_stage(h, st, worker_ixs, worker_off, n_workers, worker_rng, lik_mean, lik_var, liks, counts_off,
                                counts, counts_cap, n_alns_out, iters_out);
    h->ctx->dbg_req = nullptr;
    return rc;
}

void lctp_debug_close(lctp_ctx *ctx) {
    if (!ctx) return;
    for (FILE **f : {&ctx->dbg_sol, &ctx->dbg_sol_ext, &ctx->dbg_depth}) { if (*f) fclose(*f); *f = nullptr; }
    ctx->dbg_names.clear();
    ctx->dbg_level = 0;
    ctx->dbg_dir.clear();
}

int lctp_debug_open(lctp_ctx *ctx, const char *dir, int level, const char *const *hap_names, size_t n_haps) {
    if (!ctx || !dir || !hap_names || level < 0) { set_error("lctp_debug_open: invalid argument"); return LCTP_E_INVALID; }
    lctp_debug_close(ctx);
    ctx->dbg_dir = dir;
    ctx->dbg_level = level;
    ctx->dbg_names.assign(hap_names, hap_names + n_haps);
    auto open = [&](const char *name) { return fopen((ctx->dbg_dir + "/" + name).c_str(), "w"); };
    ctx->dbg_sol = open("sol.csv");                                                  // solve.rs:937-938
    if (ctx->dbg_sol) fprintf(ctx->dbg_sol, "stage\tgenotype\tscore\n");
    if (level >= 1) {                                                                // DebugFiles::new, solve.rs:890-897
        ctx->dbg_sol_ext = open("sol_ext.csv");
        if (ctx->dbg_sol_ext)
            fprintf(ctx->dbg_sol_ext, "stage\tgenotype\tattempt\ttotal_reads\tunmapped\tout_of_bounds\taln_lik\tdepth_lik\tlik\n");
    }
    if (level >= 2) {                                                                // :882-888
        ctx->dbg_depth = open("depth.csv");
        if (ctx->dbg_depth) fprintf(ctx->dbg_depth, "stage\tgenotype\tattempt\tcontig\twindow\tweight\tdepth\tlik\n");
    }
    if (!ctx->dbg_sol || (level >= 1 && !ctx->dbg_sol_ext) || (level >= 2 && !ctx->dbg_depth)) {
        lctp_debug_close(ctx);
        set_error("lctp_debug_open: cannot create the debug tables in %s", dir);
        return LCTP_E_INVALID;
    }
    return LCTP_OK;
}

static const double INV_LN10 = 0.4342944819032518277;      // Ln::INV_LN10, src/math/mod.rs:14

// Rust's `{:.N}` of an f64: the correctly rounded decimal expansion, like printf("%.Nf"), but "NaN" for NaNs.
static void print_fixed(FILE *f, double v, int prec) {
    if (std::isnan(v)) fputs("NaN", f);
    else fprintf(f, "%.*f", prec, v);
}

static void print_gt(FILE *f, const lctp_locus_h *h, uint64_t g) {    // Genotype::new name, src/seq/contigs.rs:412-424
    uint32_t ids[LCTP_MAX_PLOIDY];
    lctp::genotype_tuple(h->dev.H, h->dev.p, h->gt_tuples_host.empty() ? nullptr : h->gt_tuples_host.data(), g, ids);
    for (uint32_t k = 0; k < h->dev.p; k++) {
        if (k) fputc(',', f);
        fputs(ids[k] < h->ctx->dbg_names.size() ? h->ctx->dbg_names[ids[k]].c_str() : "?", f);
    }
}

static int solve_impl(lctp_locus_h *h, const lctp_stage *stages, size_t n_stages, size_t threads, uint64_t rng[4],
                      lctp_result *res, size_t n_counts, uint64_t *counts_off_out, uint16_t *counts_out,
                      uint64_t counts_cap) {
    if (!h || !stages || !rng || !res || n_stages == 0 || n_stages > LCTP_MAX_STAGES) {
        set_error("lctp_solve: invalid argument");
        return LCTP_E_INVALID;
    }
    if (n_counts && (!counts_off_out || !counts_out || n_counts > LCTP_MAX_OUT)) {
        set_error("lctp_solve_counts: invalid counts arguments");
        return LCTP_E_INVALID;
    }
    lctp_ctx *ctx = h->ctx;
    const double t_in = now_s();
    const uint64_t G = h->dev.G;
    std::memset(res, 0, sizeof(*res));
    threads = std::max<size_t>(1, std::min<size_t>(threads, G));     // genotype.rs:1247
    // per-genotype results by id: the context's reusable arrays (clean between solves), see common.cuh
    lctp::per_id_reserve(ctx, G);
    std::vector<uint64_t> &ixs = ctx->id_list;
    std::iota(ixs.begin(), ixs.begin() + G, 0);
    size_t n = G;
    std::vector<double> &lik_mean = ctx->id_mean, &lik_var = ctx->id_var;
    std::vector<uint16_t> &attempts = ctx->id_attempts;
    std::vector<uint64_t> touched;
    struct Cleaner { lctp_ctx *c; std::vector<uint64_t> &t; ~Cleaner() { lctp::per_id_clean(c, t); } } cleaner{ctx, touched};

    const double t0 = now_s();
    if (h->host.dont_skip || stages[0].in_size < G) {                // solve.rs:941-945
        const bool rows = ctx->dbg_sol && ctx->dbg_level >= 1;       // the writer is passed with debug != None only (:942)
        std::vector<double> scores(rows ? G : 0);
        int rc = lctp_prefilter(h, ixs.data(), n, stages[0].in_size, threads, &n, rows ? scores.data() : nullptr);
        if (rc) return rc;
        if (rows)                                                    // run_filter, solve.rs:115-117
            for (uint64_t g = 0; g < G; g++) {
                fputs("0\t", ctx->dbg_sol); print_gt(ctx->dbg_sol, h, g); fputc('\t', ctx->dbg_sol);
                print_fixed(ctx->dbg_sol, scores[g] * INV_LN10, 3); fputc('\n', ctx->dbg_sol);
            }
    }
    res->n_filtered = n;
    const double t1 = now_s();
    res->t_prefilter_s = t1 - t0;

    std::vector<uint64_t> wrng;
    if (threads > 1) {                                               // MainWorker::new, solve.rs:1007-1018
        wrng.resize(threads * 4);
        lctp_rng_worker_streams(rng, threads, wrng.data());
    }
    const double t_jump = now_s();
    std::vector<uint64_t> off(threads + 1);
    std::vector<double> lm, lv;
    std::vector<uint64_t> c_off;            // assignment counts of the last stage, by position
    std::vector<uint16_t> c_val;
    std::vector<uint64_t> last_ixs;         // its genotypes in dispatch order
    const uint32_t wmax = lctp_locus_wmax(h);
    for (size_t s = 0; s < n_stages; s++) {
        const lctp_stage &st = stages[s];
        const bool has_next = s + 1 < n_stages;
        const size_t out_size = has_next ? stages[s + 1].in_size : 0;
        if (!(h->host.dont_skip || !has_next || out_size < n)) continue;   // solve.rs:1041-1045
        res->n_stage_in[s] = n;
        lm.resize(n); lv.resize(n);
        if (touched.empty()) touched.assign(ixs.begin(), ixs.begin() + n);     // later stages solve subsets of these
        // debug tables of this stage (Worker::run, solve.rs:1128-1134; MainWorker::run :1064-1075)
        const bool ext_rows = ctx->dbg_sol_ext != nullptr, depth_rows = ctx->dbg_depth != nullptr;
        const bool sol_rows = ctx->dbg_sol && (ctx->dbg_level >= 1 || !has_next);
        const size_t na = n * st.attempts;
        std::vector<double> d_al, d_dl, d_ww, d_wl;
        std::vector<uint32_t> d_un, d_ob, d_wd;
        lctp_stage_debug dbg;
        std::memset(&dbg, 0, sizeof dbg);
        if (ext_rows || depth_rows) {
            d_al.resize(na); d_dl.resize(na); d_un.resize(na); d_ob.resize(na);
            dbg.aln_lik = d_al.data(); dbg.depth_lik = d_dl.data(); dbg.unmapped = d_un.data(); dbg.out_of_bounds = d_ob.data();
            if (depth_rows) {
                d_ww.resize(na * wmax); d_wl.resize(na * wmax); d_wd.resize(na * wmax);
                dbg.wmax = wmax; dbg.win_weight = d_ww.data(); dbg.win_lik = d_wl.data(); dbg.win_depth = d_wd.data();
            }
        }
        const bool want_c = n_counts > 0 && !has_next;
        std::vector<uint64_t> nal;
        if (want_c) {
            // candidate totals are not known before the stage ran: bound by n * (R + sum of the p largest haplotypes)
            std::vector<uint32_t> ha(h->hap_alns);
            std::sort(ha.begin(), ha.end(), std::greater<uint32_t>());
            uint64_t per = h->dev.R;
            for (uint32_t k = 0; k < h->dev.p && k < ha.size(); k++) per += ha[k < ha.size() ? k : 0];
            per = std::min<uint64_t>(per, 65535);
            c_off.assign(n + 1, 0);
            c_val.assign((size_t)n * per, 0);
        }
        int rc;
        size_t nw = 1;
        uint64_t *wr = rng;
        if (threads == 1) { off[0] = 0; off[1] = n; }                // solve_single_thread, solve.rs:814-843
        else { nw = lctp_plan_stage(rng, ixs.data(), n, threads, off.data()); wr = wrng.data(); }
        rc = lctp_solve_stage_dbg(h, &st, ixs.data(), off.data(), nw, wr, lm.data(), lv.data(), nullptr,
                                  want_c ? c_off.data() : nullptr, want_c ? c_val.data() : nullptr, c_val.size(), nullptr,
                                  nullptr, (ext_rows || depth_rows) ? &dbg : nullptr);
        if (rc) return rc;
        if (want_c) last_ixs.assign(ixs.begin(), ixs.begin() + n);
        for (size_t q = 0; q < n; q++) { lik_mean[ixs[q]] = lm[q]; lik_var[ixs[q]] = lv[q]; attempts[ixs[q]] = (uint16_t)st.attempts; }
        for (size_t q = 0; q < n && (ext_rows || depth_rows || sol_rows); q++) {
            uint32_t ids[LCTP_MAX_PLOIDY];
            lctp::genotype_tuple(h->dev.H, h->dev.p, h->gt_tuples_host.empty() ? nullptr : h->gt_tuples_host.data(), ixs[q], ids);
            for (uint32_t a = 0; a < st.attempts; a++) {
                const size_t ja = q * st.attempts + a;
                if (depth_rows) {                                    // write_depth, assgn.rs:359-372
                    uint32_t w = 2;
                    for (uint32_t i = 0; i < h->dev.p; i++)
                        for (uint32_t k = 0; k < h->hap_n_windows[ids[i]]; k++, w++) {
                            FILE *f = ctx->dbg_depth;
                            fprintf(f, "%zu\t", s + 1); print_gt(f, h, ixs[q]);
                            fprintf(f, "\t%u\t%u\t%u\t", a + 1, i + 1, k + 1);
                            print_fixed(f, d_ww[ja * wmax + w], 4);
                            fprintf(f, "\t%u\t", d_wd[ja * wmax + w]);
                            print_fixed(f, d_wl[ja * wmax + w] * INV_LN10, 3); fputc('\n', f);
                        }
                }
                if (ext_rows) {                                      // summarize, assgn.rs:416-425
                    FILE *f = ctx->dbg_sol_ext;
                    const double lik = h->dev.depth_contrib * d_dl[ja] + h->dev.aln_contrib * d_al[ja];   // likelihood(), :235-237
                    fprintf(f, "%zu\t", s + 1); print_gt(f, h, ixs[q]);
                    fprintf(f, "\t%u\t%u\t%u\t%u\t", a + 1, h->dev.R, d_un[ja], d_ob[ja]);
                    print_fixed(f, d_al[ja] * INV_LN10, 7); fputc('\t', f);
                    print_fixed(f, d_dl[ja] * INV_LN10, 7); fputc('\t', f);
                    print_fixed(f, lik * INV_LN10, 7); fputc('\n', f);
                }
            }
            if (sol_rows) {                                          // MainWorker::run, solve.rs:1074-1075
                FILE *f = ctx->dbg_sol;
                fprintf(f, "%zu\t", s + 1); print_gt(f, h, ixs[q]); fputc('\t', f);
                print_fixed(f, lm[q] * INV_LN10, 4); fputc('\n', f);
            }
        }
        if (has_next)
            n = lctp_discard_improbable(ixs.data(), n, lik_mean.data(), lik_var.data(), attempts.data(),
                                        h->host.prob_thresh, out_size, threads);
    }
    for (FILE *f : {ctx->dbg_sol, ctx->dbg_sol_ext, ctx->dbg_depth}) if (f) fflush(f);
    res->t_stages_s = now_s() - t1;
    const double t_st = now_s();
    const uint64_t n_filtered = res->n_filtered;
    uint64_t n_stage_in[LCTP_MAX_STAGES];
    std::memcpy(n_stage_in, res->n_stage_in, sizeof(n_stage_in));
    const double tp = res->t_prefilter_s, ts = res->t_stages_s;
    int rc = lctp_produce_result(h, ixs.data(), n, lik_mean.data(), lik_var.data(), attempts.data(), res);
    if (rc) return rc;
    res->n_filtered = n_filtered;
    std::memcpy(res->n_stage_in, n_stage_in, sizeof(n_stage_in));
    res->t_prefilter_s = tp; res->t_stages_s = ts;
    if (n_counts) {                                                  // Prediction::assgn_counts of the reported genotypes
        uint64_t o = 0;
        counts_off_out[0] = 0;
        for (size_t k = 0; k < n_counts; k++) {
            if (k < res->n_out) {
                const size_t q = std::find(last_ixs.begin(), last_ixs.end(), res->gt_ix[k]) - last_ixs.begin();
                if (q == last_ixs.size()) { set_error("lctp_solve_counts: genotype missing from the last stage"); return LCTP_E_INVALID; }
                const uint64_t len = c_off[q + 1] - c_off[q];
                if (o + len > counts_cap) { set_error("lctp_solve_counts: counts buffer too small"); return LCTP_E_CAPACITY; }
                std::copy(c_val.begin() + c_off[q], c_val.begin() + c_off[q + 1], counts_out + o);
                o += len;
            }
            counts_off_out[k + 1] = o;
        }
    }
    if (getenv("LCTP_DEBUG_TIMES"))
        fprintf(stderr, "[lctp debug] ctx %p lctp_solve host: setup %.3f, prefilter %.3f, worker streams %.3f, stages %.3f, result %.3f ms\n",
                (void *)h->ctx, (t0 - t_in) * 1e3, (t1 - t0) * 1e3, (t_jump - t1) * 1e3, (t_st - t_jump) * 1e3, (now_s() - t_st) * 1e3);
    return LCTP_OK;
}

int lctp_solve(lctp_locus_h *h, const lctp_stage *stages, size_t n_stages, size_t threads, uint64_t rng[4],
               lctp_result *res) {
    return solve_impl(h, stages, n_stages, threads, rng, res, 0, nullptr, nullptr, 0);
}

int lctp_solve_counts(lctp_locus_h *h, const lctp_stage *stages, size_t n_stages, size_t threads, uint64_t rng[4],
                      lctp_result *res, size_t n_counts, uint64_t *counts_off, uint16_t *counts, uint64_t counts_cap) {
    return solve_impl(h, stages, n_stages, threads, rng, res, n_counts, counts_off, counts, counts_cap);
}

// ---- Genotyping::find_weighted_dist (src/solvers/solve.rs:616-632) ---------------------------------------

// TriangleMatrix::get_symmetric (src/ext/trimat.rs): linear index of (i, j), i != j
static inline size_t tri_index(size_t side, size_t i, size_t j) {
    if (i > j) std::swap(i, j);
    return (2 * side - 3 - i) * i / 2 + j - 1;
}

// genotype_distance (src/solvers/solve.rs:339-357): minimum over the permutations of gt1 that
// ext::vec::gen_permutations visits (src/ext/vec.rs:342-372) -- for three or more contigs that is Heap's
// algorithm WITHOUT the initial arrangement (the reference calls `action` only after each swap).
static uint32_t genotype_distance(const uint32_t *gt1, const uint32_t *gt2, uint32_t p, uint32_t H, const uint32_t *dist) {
    uint32_t min_dist = 0xFFFFFFFFu;
    auto visit = [&](const uint32_t *perm) {
        uint32_t d = 0;
        for (uint32_t k = 0; k < p; k++) {
            if (perm[k] != gt2[k]) {
                const uint32_t e = dist[tri_index(H, perm[k], gt2[k])];
                if (e == LCTP_NONE_U32) { d = 0xFFFFFFFFu; break; }
                d += e;
            }
        }
        min_dist = std::min(min_dist, d);
    };
    if (p == 1) visit(gt1);
    else if (p == 2) {
        visit(gt1);
        const uint32_t sw[2] = {gt1[1], gt1[0]};
        visit(sw);
    } else if (p >= 3) {
        std::vector<uint32_t> buffer(gt1, gt1 + p);
        std::vector<size_t> c(p, 0);
        size_t i = 1;
        while (i < p) {
            if (c[i] < i) {
                std::swap(buffer[i], buffer[c[i] * (i % 2)]);     // 0 if i is even, c[i] if i is odd
                visit(buffer.data());
                c[i] += 1;
                i = 1;
            } else {
                c[i] = 0;
                i += 1;
            }
        }
    }
    return min_dist;     // 0xFFFFFFFF = None
}

int lctp_find_weighted_dist(lctp_result *res, const lctp_locus *loc, const uint32_t *dist, int true_edit_distances) {
    if (!res || !loc || !dist) { set_error("lctp_find_weighted_dist: NULL argument"); return LCTP_E_INVALID; }
    if (loc->ploidy == 0 || loc->ploidy > LCTP_MAX_PLOIDY) { set_error("lctp_find_weighted_dist: bad ploidy"); return LCTP_E_INVALID; }
    if (res->n_out == 0) return LCTP_OK;                   // `let Some(gt0) = self.genotypes.first() else { return }`
    const uint32_t p = loc->ploidy, H = loc->n_haps;
    uint32_t gt0[LCTP_MAX_PLOIDY], gt[LCTP_MAX_PLOIDY];
    genotype_tuple(H, p, loc->gt_tuples, res->gt_ix[0], gt0);
    double sum_prob = 0.0, sum_dist = 0.0;
    bool known = true;
    for (uint64_t i = 0; i < res->n_out; i++) {
        const double prob = std::exp(res->ln_prob[i]);
        sum_prob += prob;
        uint32_t d = 0;
        if (i > 0) {
            genotype_tuple(H, p, loc->gt_tuples, res->gt_ix[i], gt);
            for (uint32_t k = 0; k < p; k++)
                if (gt[k] >= H || gt0[k] >= H) { set_error("lctp_find_weighted_dist: contig id out of range"); return LCTP_E_INVALID; }
            d = genotype_distance(gt0, gt, p, H, dist);
        }
        if (d == 0xFFFFFFFFu) known = false;               // Option::zip: one None makes the sum None for good
        else if (known) sum_dist += prob * (double)d;
        res->dist_to_primary[i] = d;
    }
    res->has_dist = 1;
    res->true_edit_distances = true_edit_distances ? 1 : 0;
    res->has_weight_dist = known ? 1 : 0;
    res->weight_dist = known ? sum_dist / sum_prob : std::numeric_limits<double>::quiet_NaN();
    return LCTP_OK;
}

// ---- Genotyping::to_json (src/solvers/solve.rs:732-773) as `json::JsonValue::write_pretty(.., 4)` prints it -----
//
// Numbers: the `json` crate (0.12) stores an f64 as a decimal (mantissa, exponent) pair of its shortest
// round-trip digits and prints it with its own rules (util/print_dec.rs): plain digits with the decimal point
// inserted while fewer than 18 fraction digits are needed, zero-filled integers up to 20 characters, otherwise
// d.ddd e[-]N; NaN and infinities print as null.  (Restated from the published crate -- it is not vendored by the
// reference -- so the exact thresholds are unpinned until a Rust toolchain is available.)
static void json_num(std::string &o, double v) {
    if (std::isnan(v) || std::isinf(v)) { o += "null"; return; }
    if (std::signbit(v)) { o += '-'; v = -v; }
    if (v == 0.0) { o += '0'; return; }
    char b[64];
    auto r = std::to_chars(b, b + sizeof b - 1, v, std::chars_format::scientific);
    *r.ptr = 0;
    std::string digits;
    int e10 = 0;
    {
        const char *q = b;
        for (; q < r.ptr && *q != 'e'; q++) if (*q != '.') digits += *q;
        e10 = atoi(q + 1);
    }
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int k = (int)digits.size();
    const int exponent = e10 - (k - 1);                    // v = digits * 10^exponent
    char eb[16];
    if (exponent == 0) o += digits;
    else if (exponent < 0) {
        const int e = -exponent;
        if (e < 18) {
            if (k > e) { o.append(digits, 0, k - e); o += '.'; o.append(digits, k - e, std::string::npos); }
            else { o += "0."; o.append((size_t)(e - k), '0'); o += digits; }
        } else {
            o += digits[0];
            if (k > 1) { o += '.'; o.append(digits, 1, std::string::npos); }
            snprintf(eb, sizeof eb, "e%d", e10);           // e10 = exponent + k - 1 (may have become positive)
            o += eb;
        }
    } else {
        if (k + exponent <= 20) { o += digits; o.append((size_t)exponent, '0'); }
        else {
            o += digits[0];
            if (k > 1) { o += '.'; o.append(digits, 1, std::string::npos); }
            snprintf(eb, sizeof eb, "e%d", e10);
            o += eb;
        }
    }
}

size_t lctp_result_json(const lctp_result *res, const lctp_locus *loc, const char *const *hap_names, char *buf,
                        size_t cap) {
    std::string o;
    const double inv_ln10 = 0.4342944819032518277;
    auto gt_name = [&](uint64_t g) {
        uint32_t ids[LCTP_MAX_PLOIDY];
        genotype_tuple(loc->n_haps, loc->ploidy, loc->gt_tuples, g, ids);
        std::string s;
        for (uint32_t k = 0; k < loc->ploidy; k++) { if (k) s += ","; s += hap_names[ids[k]]; }   // Genotype name, contigs.rs:404-459
        return s;
    };
    o += "{\n    \"total_reads\": "; json_num(o, res->total_reads);
    o += ",\n    \"quality\": "; json_num(o, res->quality);
    if (res->has_dist) o += res->true_edit_distances ? ",\n    \"dist_type\": \"edit\"" : ",\n    \"dist_type\": \"minim-div\"";
    if (res->has_dist && res->has_weight_dist) { o += ",\n    \"weight_dist\": "; json_num(o, res->weight_dist); }
    o += ",\n    \"unexpl_reads\": "; json_num(o, res->unexpl_reads);
    if (res->n_out) {
        o += ",\n    \"genotype\": \"" + gt_name(res->gt_ix[0]) + "\"";
        o += ",\n    \"options\": [";
        for (uint64_t i = 0; i < res->n_out; i++) {
            o += i ? ",\n        {" : "\n        {";
            o += "\n            \"genotype\": \"" + gt_name(res->gt_ix[i]) + "\"";
            o += ",\n            \"lik_mean\": "; json_num(o, res->lik_mean[i] * inv_ln10);
            o += ",\n            \"lik_sd\": "; json_num(o, res->lik_var[i] * inv_ln10);   // sic: log10 of the variance (solve.rs:755)
            o += ",\n            \"prob\": "; json_num(o, std::exp(res->ln_prob[i]));
            o += ",\n            \"log10_prob\": "; json_num(o, res->ln_prob[i] * inv_ln10);
            if (res->has_dist) {
                o += ",\n            \"dist_to_primary\": ";
                if (res->dist_to_primary[i] == LCTP_NONE_U32) o += "\"unknown\"";
                else json_num(o, res->dist_to_primary[i]);
            }
            o += "\n        }";
        }
        o += "\n    ]";
    }
    if (res->warn_no_probable || res->warn_few_reads) {
        o += ",\n    \"warnings\": [";
        bool first = true;
        if (res->warn_no_probable) { o += "\n        \"NoProbableGenotype\""; first = false; }
        if (res->warn_few_reads) {
            char b[64]; snprintf(b, sizeof b, "%s\n        \"FewReads(%u)\"", first ? "" : ",", res->total_reads); o += b;
        }
        o += "\n    ]";
    }
    o += "\n}";
    if (buf && cap) {
        size_t ncopy = std::min(cap - 1, o.size());
        std::memcpy(buf, o.data(), ncopy);
        buf[ncopy] = 0;
    }
    return o.size();
}

}  // extern "C"

/*
 * lcto_rng.c -- ORACLE (test infrastructure): RNG shim.
 *
 * Restates, from the published algorithms, the third-party RNG arithmetic the
 * reference uses on the hot path (none of it is vendored in /root/reference):
 *   rand_xoshiro 0.8  Xoshiro256PlusPlus {seed_from_u64, next_u64, next_u32, jump, long_jump}
 *   rand 0.10         uniform integer sampling (widening multiply + one bias-reduction draw),
 *                     StandardUniform f64, seq::index::sample (Floyd / in-place),
 *                     SliceRandom::shuffle (IncreasingUniform batched Fisher-Yates).
 * Call sites in the reference: src/ext/rand.rs:3-22, src/solvers/solve.rs:1017,1051,
 * src/model/windows.rs:127,483, src/model/assgn.rs:452,462, src/solvers/stoch.rs:93,102,205,216.
 *
 * PARITY UNPINNED for the rand 0.10 algorithms: they are restated from memory of the crate
 * sources and cannot be verified offline.  xoshiro256++ / SplitMix64 are pinned by the
 * public-domain known-answer vector (tests/test_oracle_rng.py).
 * Each algorithm sits behind one function so a correction is a local change.
 */
#include "lcto.h"
#include <stdlib.h>
#include <string.h>

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

/* SplitMix64 (Vigna, public domain); rand_xoshiro::SplitMix64. */
static uint64_t splitmix64_next(uint64_t *x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

/* Xoshiro256PlusPlus::seed_from_u64: state filled with four SplitMix64 outputs (src/ext/rand.rs:12). */
void lcto_rng_seed_from_u64(lcto_rng *r, uint64_t seed) {
    uint64_t x = seed;
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64_next(&x);
}

/* xoshiro256++ 1.0 (Blackman & Vigna, public domain). */
uint64_t lcto_rng_next_u64(lcto_rng *r) {
    uint64_t *s = r->s;
    const uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return result;
}

/* rand_xoshiro: next_u32 takes the upper bits. */
uint32_t lcto_rng_next_u32(lcto_rng *r) { return (uint32_t)(lcto_rng_next_u64(r) >> 32); }

static void jump_with(lcto_rng *r, const uint64_t poly[4]) {
    uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int i = 0; i < 4; i++) {
        for (int b = 0; b < 64; b++) {
            if (poly[i] & (1ULL << b)) {
                s0 ^= r->s[0]; s1 ^= r->s[1]; s2 ^= r->s[2]; s3 ^= r->s[3];
            }
            lcto_rng_next_u64(r);
        }
    }
    r->s[0] = s0; r->s[1] = s1; r->s[2] = s2; r->s[3] = s3;
}

/* 2^128 steps; src/solvers/solve.rs:1017 */
void lcto_rng_jump(lcto_rng *r) {
    static const uint64_t J[4] = { 0x180ec6d33cfd0abaULL, 0xd5a61266f0c9392cULL,
                                   0xa9582618e03fc9aaULL, 0x39abdc4529b1661cULL };
    jump_with(r, J);
}

/* 2^192 steps; src/command/genotype.rs:1345 */
void lcto_rng_long_jump(lcto_rng *r) {
    static const uint64_t J[4] = { 0x76e15d3efefdcbbfULL, 0xc5004e441c522fb3ULL,
                                   0x77710069854ee241ULL, 0x39109bb02acbe635ULL };
    jump_with(r, J);
}

/* rand StandardUniform for f64: 53 random bits scaled by 2^-53; src/solvers/stoch.rs:216 */
double lcto_rng_f64(lcto_rng *r) {
    return (double)(lcto_rng_next_u64(r) >> 11) * (1.0 / 9007199254740992.0);
}

/* rand UniformInt::sample_single_inclusive with a 32-bit sample type (u8/u16/u32/i32, and usize
 * ranges that fit u32): widening multiply, then ONE extra draw to reduce bias when the low half
 * lands in the biased zone. */
uint32_t lcto_rng_range_u32_incl(lcto_rng *r, uint32_t low, uint32_t high) {
    uint32_t range = high - low + 1u;
    if (range == 0) return lcto_rng_next_u32(r);
    uint64_t m = (uint64_t)lcto_rng_next_u32(r) * (uint64_t)range;
    uint32_t result = (uint32_t)(m >> 32);
    uint32_t lo_order = (uint32_t)m;
    if (lo_order > (uint32_t)(0u - range)) {
        uint64_t m2 = (uint64_t)lcto_rng_next_u32(r) * (uint64_t)range;
        uint32_t new_hi = (uint32_t)(m2 >> 32);
        uint32_t sum = lo_order + new_hi;
        result += (sum < lo_order) ? 1u : 0u;   /* checked_add overflow */
    }
    return low + result;
}

uint64_t lcto_rng_range_u64_incl(lcto_rng *r, uint64_t low, uint64_t high) {
    uint64_t range = high - low + 1u;
    if (range == 0) return lcto_rng_next_u64(r);
    __uint128_t m = (__uint128_t)lcto_rng_next_u64(r) * (__uint128_t)range;
    uint64_t result = (uint64_t)(m >> 64);
    uint64_t lo_order = (uint64_t)m;
    if (lo_order > (uint64_t)(0u - range)) {
        __uint128_t m2 = (__uint128_t)lcto_rng_next_u64(r) * (__uint128_t)range;
        uint64_t new_hi = (uint64_t)(m2 >> 64);
        uint64_t sum = lo_order + new_hi;
        result += (sum < lo_order) ? 1u : 0u;
    }
    return low + result;
}

/* i32 inclusive range (unsigned type u32, sample type u32); src/model/windows.rs:483 */
int32_t lcto_rng_range_i32_incl(lcto_rng *r, int32_t low, int32_t high) {
    uint32_t span = (uint32_t)high - (uint32_t)low;     /* wrapping_sub as unsigned */
    uint32_t res = lcto_rng_range_u32_incl(r, 0u, span);
    return (int32_t)((uint32_t)low + res);
}

/* usize half-open range: rand's UniformUsize samples as u32 whenever the bound fits in u32
 * (portability between 32/64-bit targets), else as u64.  src/model/assgn.rs:452, stoch.rs:93,205 */
size_t lcto_rng_range_usize(lcto_rng *r, size_t low, size_t high_excl) {
    if (high_excl <= (size_t)0xFFFFFFFFu)
        return (size_t)lcto_rng_range_u32_incl(r, (uint32_t)low, (uint32_t)(high_excl - 1));
    return (size_t)lcto_rng_range_u64_incl(r, (uint64_t)low, (uint64_t)(high_excl - 1));
}

/* u16 half-open range (sample type u32); src/model/assgn.rs:462 */
uint16_t lcto_rng_range_u16(lcto_rng *r, uint16_t low, uint16_t high_excl) {
    uint16_t high = (uint16_t)(high_excl - 1);
    uint32_t range = (uint32_t)(uint16_t)(high - low + 1);
    if (range == 0) return (uint16_t)lcto_rng_next_u32(r);
    uint64_t m = (uint64_t)lcto_rng_next_u32(r) * (uint64_t)range;
    uint32_t result = (uint32_t)(m >> 32);
    uint32_t lo_order = (uint32_t)m;
    if (lo_order > (uint32_t)(0u - range)) {
        uint64_t m2 = (uint64_t)lcto_rng_next_u32(r) * (uint64_t)range;
        uint32_t new_hi = (uint32_t)(m2 >> 32);
        uint32_t sum = lo_order + new_hi;
        result += (sum < lo_order) ? 1u : 0u;
    }
    return (uint16_t)(low + (uint16_t)result);
}

/* rand::seq::index::sample for length <= u32::MAX, used by IndexedRandom::sample
 * (src/solvers/stoch.rs:102).  Returns 0 on success, -1 if the (unsupported) rejection
 * branch would be selected. */
int lcto_rng_sample_indices(lcto_rng *r, uint32_t length, uint32_t amount, uint32_t *out) {
    int use_inplace;
    if (amount < 163) {
        static const float C[2][2] = { {1.6f, 8.0f / 45.0f}, {10.0f, 70.0f / 9.0f} };
        int j = length >= 500000u ? 1 : 0;
        float amount_fp = (float)amount;
        float m4 = C[0][j] * amount_fp;
        use_inplace = (amount > 11 && (float)length < (C[1][j] + m4) * amount_fp);
    } else {
        static const float C2[2] = { 270.0f, 330.0f / 9.0f };
        int j = length < 500000u ? 0 : 1;
        if ((float)length < C2[j] * (float)amount) use_inplace = 1;
        else return -1;   /* sample_rejection: not restated */
    }
    if (!use_inplace) {
        /* sample_floyd */
        uint32_t n = 0;
        for (uint32_t j = length - amount; j < length; j++) {
            uint32_t t = lcto_rng_range_u32_incl(r, 0u, j);
            for (uint32_t k = 0; k < n; k++) {
                if (out[k] == t) { out[k] = j; break; }
            }
            out[n++] = t;
        }
        return 0;
    }
    /* sample_inplace */
    uint32_t *idx = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)length);
    if (!idx) return -2;
    for (uint32_t i = 0; i < length; i++) idx[i] = i;
    for (uint32_t i = 0; i < amount; i++) {
        uint32_t j = lcto_rng_range_u32_incl(r, i, length - 1);
        uint32_t tmp = idx[i]; idx[i] = idx[j]; idx[j] = tmp;
    }
    memcpy(out, idx, sizeof(uint32_t) * (size_t)amount);
    free(idx);
    return 0;
}

/* rand::seq::increasing_uniform::calculate_bound_u32: bound = m*(m+1)*..*(m+count-1) < 2^32 */
static void calculate_bound_u32(uint32_t m, uint32_t *bound, uint8_t *count) {
    uint32_t product = m;
    uint32_t current = m + 1;
    for (;;) {
        uint64_t p = (uint64_t)product * (uint64_t)current;
        if (p <= 0xFFFFFFFFull) { product = (uint32_t)p; current += 1; }
        else { *bound = product; *count = (uint8_t)(current - m); return; }
    }
}

/* SliceRandom::shuffle = partial_shuffle(len) driven by IncreasingUniform (rand >= 0.9):
 * one bounded u32 draw yields several Fisher-Yates indices.  src/solvers/solve.rs:1051 */
void lcto_rng_shuffle_usize(lcto_rng *r, size_t *v, size_t len) {
    if (len <= 1) return;
    if (len >= (size_t)0xFFFFFFFFu) {
        for (size_t i = 0; i < len; i++) {
            size_t idx = lcto_rng_range_usize(r, 0, i + 1);
            size_t t = v[i]; v[i] = v[idx]; v[idx] = t;
        }
        return;
    }
    uint32_t n = 0;
    uint32_t chunk = 0;
    uint8_t chunk_remaining = 1;   /* IncreasingUniform::new(rng, 0): first index is always 0 */
    for (size_t i = 0; i < len; i++) {
        uint32_t next_n = n + 1;
        uint8_t next_remaining;
        if (chunk_remaining >= 1) {
            next_remaining = (uint8_t)(chunk_remaining - 1);
        } else {
            uint32_t bound; uint8_t remaining;
            calculate_bound_u32(next_n, &bound, &remaining);
            chunk = lcto_rng_range_u32_incl(r, 0u, bound - 1u);
            next_remaining = (uint8_t)(remaining - 1);
        }
        size_t index;
        if (next_remaining == 0) {
            index = (size_t)chunk;
        } else {
            index = (size_t)(chunk % next_n);
            chunk /= next_n;
        }
        chunk_remaining = next_remaining;
        n = next_n;
        size_t t = v[i]; v[i] = v[index]; v[index] = t;
    }
}

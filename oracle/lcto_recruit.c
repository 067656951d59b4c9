/* ORACLE (test infrastructure; never imported, linked or executed by the product path).
 *
 * CPU restatement of short-read recruitment, SURVEY.md section 8(f) rank 3: canonical minimizers
 * (src/seq/kmers.rs:71-103, 256-340), the target tables (TargetBuilder::add, src/seq/recruit.rs:680-735), the match
 * counters (:234-385), Fraction (src/math/frac.rs:48-96) and the recruitment decisions of single-end and paired-end short
 * reads (recruit.rs:852-930).  Long reads (:932-998) are not restated.  Parity unpinned by the reference itself (no tests
 * or fixtures there, no Rust toolchain here); pinned by a statement-by-statement Python transcription of the cited Rust
 * (tests/test_recruit.py).
 */
#include "lcto.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#define UNDEF 0xFFFFFFFFFFFFFFFFull      /* Kmer::UNDEF = Self::MAX, kmers.rs:45 */
#define MAXW 64u                          /* MAX_MINIMIZER_W, kmers.rs:208 */

/* Minimizer for u64: fasthash mix, kmers.rs:93-103 */
static inline uint64_t fast_hash(uint64_t x) {
    x = ~x;
    x ^= x >> 23;
    x *= 0x2127599bf4325c37ull;
    x ^= x >> 47;
    return x;
}

/* kmers::minimizers::<u64, _, CANONICAL>, kmers.rs:256-331.  Returns the number of minimizers; writes up to `cap`. */
size_t lcto_minimizers(const uint8_t *seq, size_t len, uint32_t k, uint32_t w, uint64_t *hash, uint32_t *pos, uint8_t *fw,
                       size_t cap) {
    const uint64_t mask = (1ull << (2 * k)) - 1ull;               /* create_mask, kmers.rs:48-52 (k <= 31) */
    const uint32_t rv_shift = 2 * k - 2;
    uint64_t fw_kmer = 0, rv_kmer = 0;
    const uint32_t k_1 = k - 1, w_1 = w - 1;
    uint64_t hashes[MAXW]; uint8_t forward[MAXW];
    for (uint32_t q = 0; q < MAXW; q++) { hashes[q] = UNDEF; forward[q] = 1; }
    int64_t last_pos = -1;
    uint32_t best_pos = 0;
    uint64_t best_hash = UNDEF;
    uint32_t first_kmer = k_1, first_window = k_1 + w_1;
    size_t n_out = 0;
    for (size_t ii = 0; ii < len; ii++) {
        const uint32_t i = (uint32_t)ii;
        uint64_t fw_enc, rv_enc;
        switch (seq[ii]) {
        case 'A': fw_enc = 0; rv_enc = 3; break;
        case 'C': fw_enc = 1; rv_enc = 2; break;
        case 'G': fw_enc = 2; rv_enc = 1; break;
        case 'T': fw_enc = 3; rv_enc = 0; break;
        default: first_kmer = i + k; fw_enc = 0; rv_enc = 0; break;
        }
        fw_kmer = ((fw_kmer << 2) | fw_enc) & mask;
        rv_kmer = (rv_kmer >> 2) | (rv_enc << rv_shift);
        uint64_t kmer; uint8_t f;
        if (rv_kmer < fw_kmer) { kmer = rv_kmer; f = 0; } else { kmer = fw_kmer; f = 1; }
        const uint64_t h = i < first_kmer ? UNDEF : fast_hash(kmer);
        hashes[i & (MAXW - 1)] = h;
        forward[i & (MAXW - 1)] = f;
        if (h < best_hash) { best_hash = h; best_pos = i; }
        if (i < first_window) continue;
        const uint32_t start = i - w_1;
        if (best_pos < start) {
            /* find_min, kmers.rs:237-252 */
            uint32_t p = start; uint64_t m = hashes[start & (MAXW - 1)];
            for (uint32_t j = start + 1; j < i + 1; j++) { const uint64_t v = hashes[j & (MAXW - 1)]; if (v < m) { p = j; m = v; } }
            best_pos = p; best_hash = m;
            if (best_hash == UNDEF) { first_window = first_window + w_1; continue; }
        }
        if ((int64_t)best_pos > last_pos) {
            last_pos = (int64_t)best_pos;
            if (n_out < cap) { hash[n_out] = best_hash; pos[n_out] = best_pos - k_1; fw[n_out] = forward[best_pos & (MAXW - 1)]; }
            n_out++;
        }
    }
    return n_out;
}

/* Fraction::<u16>::approximate, src/math/frac.rs:48-80 */
void lcto_fraction_approximate_u16(double x, uint16_t *num, uint16_t *den) {
    uint32_t a2 = 1, a1 = (uint32_t)floor(x), b2 = 0, b1 = 1;
    double xk = x;
    for (int it = 0; it < 20; it++) {
        const double numer = xk - floor(xk);
        if (numer <= DBL_EPSILON) break;
        xk = 1.0 / numer;
        const double fl = floor(xk);
        if (!(fl >= 0.0 && fl <= 65535.0)) break;                     /* T::from_f64 -> None */
        const uint64_t f = (uint64_t)fl;
        const uint64_t a0 = f * a1 + a2, b0 = f * b1 + b2;
        if (f * a1 > 65535 || a0 > 65535) break;                      /* checked_mul / checked_add */
        if (f * b1 > 65535 || b0 > 65535) break;
        a2 = a1; a1 = (uint32_t)a0; b2 = b1; b1 = (uint32_t)b0;
        if (fabs((double)a1 / (double)b1 - x) <= DBL_EPSILON) break;
    }
    *num = (uint16_t)a1; *den = (uint16_t)b1;
}

/* ---- targets: minimizer -> [(locus, info)], TargetBuilder::add (recruit.rs:680-735).  info = direction | rare << 2. */
typedef struct { uint64_t key; uint32_t locus; uint8_t dir, rare; } tentry;
struct lcto_targets {
    tentry *e; size_t n, cap;      /* in insertion order: per minimizer the loci are appended in locus order */
    uint32_t k, w; uint16_t fnum, fden;
    double match_frac; uint32_t stretch_minims, stretch_score;     /* Params, recruit.rs:95-103 */
};

static tentry *find_last(lcto_targets *T, uint64_t key) {         /* v.last_mut() of minim_to_loci[key] */
    for (size_t q = T->n; q-- > 0;) if (T->e[q].key == key) return &T->e[q];
    return NULL;
}

lcto_targets *lcto_targets_build(const lcto_target_seqs *in) {
    lcto_targets *T = (lcto_targets *)calloc(1, sizeof *T);
    T->k = in->minimizer_k; T->w = in->minimizer_w;
    lcto_fraction_approximate_u16(in->match_frac, &T->fnum, &T->fden);   /* Params::new, recruit.rs:101 */
    T->match_frac = in->match_frac;
    T->stretch_minims = (2 * in->match_length + (in->minimizer_w + 1) - 1) / (in->minimizer_w + 1);   /* fast_ceil_div, :95 */
    {
        double sc = (double)T->stretch_minims * ((double)(3 + 1) * in->match_frac - (double)1);        /* :96-97 */
        if (!(sc > 3.0)) sc = 3.0;                                                                      /* .max(SUBSUM_BONUS) */
        T->stretch_score = (uint32_t)ceil(sc);                                                          /* :98 */
    }
    const uint32_t base_k = in->base_k, mk = in->minimizer_k;
    const size_t shift = mk <= base_k ? (base_k - mk) / 2 : mk - base_k;  /* :688-692 */
    for (uint64_t s = 0; s < in->n_seqs; s++) {
        const uint8_t *seq = in->seqs + in->seq_off[s];
        const size_t len = in->seq_off[s + 1] - in->seq_off[s];
        const uint16_t *counts = in->kmer_counts + in->cnt_off[s];
        const size_t n_counts = in->cnt_off[s + 1] - in->cnt_off[s];
        uint64_t *h = (uint64_t *)malloc(8 * (len + 1)); uint32_t *p = (uint32_t *)malloc(4 * (len + 1)); uint8_t *f = (uint8_t *)malloc(len + 1);
        const size_t nm = lcto_minimizers(seq, len, mk, in->minimizer_w, h, p, f, len + 1);
        for (size_t q = 0; q < nm; q++) {
            const size_t pos = p[q];
            int rare;
            if (mk <= base_k) {                                        /* :708-711 */
                size_t at = pos >= shift ? pos - shift : 0;
                if (at > n_counts - 1) at = n_counts - 1;
                rare = counts[at] < in->thresh_kmer_count;
            } else rare = counts[pos] < in->thresh_kmer_count && counts[pos + shift] < in->thresh_kmer_count;   /* :713 */
            tentry *last = find_last(T, h[q]);
            if (last && last->locus == in->seq_locus[s]) {            /* MinimInfo::update, :634-638 */
                last->dir |= (uint8_t)(1 + f[q]); last->rare &= (uint8_t)rare;
            } else {                                                   /* MinimInfo::new, :626-632 */
                if (T->n == T->cap) { T->cap = T->cap ? 2 * T->cap : 1024; T->e = (tentry *)realloc(T->e, T->cap * sizeof(tentry)); }
                T->e[T->n].key = h[q]; T->e[T->n].locus = in->seq_locus[s]; T->e[T->n].dir = (uint8_t)(1 + f[q]); T->e[T->n].rare = (uint8_t)rare;
                T->n++;
            }
        }
        free(h); free(p); free(f);
    }
    return T;
}
void lcto_targets_free(lcto_targets *T) { if (T) { free(T->e); free(T); } }
size_t lcto_targets_entries(const lcto_targets *T, uint64_t *key, uint32_t *locus, uint8_t *info, size_t cap) {
    for (size_t q = 0; q < T->n && q < cap; q++) { key[q] = T->e[q].key; locus[q] = T->e[q].locus; info[q] = (uint8_t)(T->e[q].dir | (T->e[q].rare << 2)); }
    return T->n;
}

/* BaseMatchCount<u16>: arr = [common-backward, common-forward, rare-backward, rare-forward], recruit.rs:234-252 */
typedef struct { uint16_t arr[4]; } bmc;
static inline int directed_to(uint8_t dir, int forward) { return (dir & (1 + forward)) != 0; }       /* :641-643 */
static inline void bmc_inc(bmc *c, int forward, uint8_t dir, uint8_t rare) {                         /* :246-252 */
    const int i = rare << 1;
    c->arr[i] = (uint16_t)(c->arr[i] + directed_to(dir, !forward));
    c->arr[i | 1] = (uint16_t)(c->arr[i | 1] + directed_to(dir, forward));
}
static inline int bmc_has_rare(bmc c) { return c.arr[2] != 0 || c.arr[3] != 0; }                      /* :257-259 */
#define WORTH 3                                                                                        /* :284-286 */
static inline uint16_t fw_num(bmc c) { return (uint16_t)(WORTH * c.arr[3] + c.arr[1]); }               /* :299-303 */
static inline uint16_t bw_num(bmc c) { return (uint16_t)(WORTH * c.arr[2] + c.arr[0]); }               /* :306-310 */
static inline uint16_t fw_den(bmc c, uint16_t t) { return (uint16_t)(WORTH * (t - c.arr[1]) + c.arr[1]); }   /* :313-316 */
static inline uint16_t bw_den(bmc c, uint16_t t) { return (uint16_t)(WORTH * (t - c.arr[0]) + c.arr[0]); }   /* :319-322 */
/* Fraction<u16> >= : frac.rs:92-98 */
static inline int frac_ge(uint16_t n1, uint16_t d1, uint16_t n2, uint16_t d2) { return (uint32_t)n1 * d2 >= (uint32_t)n2 * d1; }

typedef struct { uint32_t locus; bmc first, second; } lmatch;

/* recruit_short_read (:852-881) when seq2 == NULL, recruit_read_pair (:885-930) otherwise.  The answer is a set (the
 * reference iterates a hash map): written in ascending locus order.  Returns the number of loci, or -1 when a read has
 * more than 65535 minimizers (the reference panics). */
static int recruit_one(const lcto_targets *T, const uint8_t *s1, size_t l1, const uint8_t *s2, size_t l2, uint32_t *ans, uint32_t cap) {
    size_t mcap = (l1 > l2 ? l1 : l2) + 1;
    uint64_t *h = (uint64_t *)malloc(8 * mcap); uint32_t *p = (uint32_t *)malloc(4 * mcap); uint8_t *f = (uint8_t *)malloc(mcap);
    lmatch *m = NULL; size_t nm = 0, mc = 0;
    int n_ans = 0;
    const size_t t1 = lcto_minimizers(s1, l1, T->k, T->w, h, p, f, mcap);
    if (t1 > 65535) { n_ans = -1; goto done; }
    for (size_t q = 0; q < t1; q++)
        for (size_t e = 0; e < T->n; e++) {
            if (T->e[e].key != h[q]) continue;
            size_t z = 0;
            while (z < nm && m[z].locus != T->e[e].locus) z++;
            if (z == nm) { if (nm == mc) { mc = mc ? 2 * mc : 8; m = (lmatch *)realloc(m, mc * sizeof(lmatch)); } memset(&m[nm], 0, sizeof(lmatch)); m[nm].locus = T->e[e].locus; nm++; }
            bmc_inc(&m[z].first, f[q], T->e[e].dir, T->e[e].rare);
        }
    if (s2) {
        if (nm == 0) goto done;                                      /* :906 */
        const size_t t2 = lcto_minimizers(s2, l2, T->k, T->w, h, p, f, mcap);
        if (t2 > 65535) { n_ans = -1; goto done; }
        for (size_t q = 0; q < t2; q++)
            for (size_t e = 0; e < T->n; e++) {
                if (T->e[e].key != h[q]) continue;
                for (size_t z = 0; z < nm; z++) if (m[z].locus == T->e[e].locus) bmc_inc(&m[z].second, f[q], T->e[e].dir, T->e[e].rare);   /* :916-919 */
            }
        for (size_t z = 0; z < nm; z++) {
            const bmc a = m[z].first, b = m[z].second;
            if (!(bmc_has_rare(a) || bmc_has_rare(b))) continue;     /* :923 */
            uint16_t n1, d1, n2, d2;                                  /* better_pair_fraction, :349-366 */
            if ((uint16_t)(fw_num(a) + bw_num(b)) >= (uint16_t)(bw_num(a) + fw_num(b))) { n1 = fw_num(a); d1 = fw_den(a, (uint16_t)t1); n2 = bw_num(b); d2 = bw_den(b, (uint16_t)t2); }
            else { n1 = bw_num(a); d1 = bw_den(a, (uint16_t)t1); n2 = fw_num(b); d2 = fw_den(b, (uint16_t)t2); }
            if (frac_ge(n1, d1, T->fnum, T->fden) && frac_ge(n2, d2, T->fnum, T->fden)) { if ((uint32_t)n_ans < cap) ans[n_ans] = m[z].locus; n_ans++; }
        }
    } else {
        for (size_t z = 0; z < nm; z++) {
            const bmc a = m[z].first;
            if (!bmc_has_rare(a)) continue;                          /* :877 */
            uint16_t n1, d1;                                          /* better_fraction, :337-346 */
            if (fw_num(a) >= bw_num(a)) { n1 = fw_num(a); d1 = fw_den(a, (uint16_t)t1); } else { n1 = bw_num(a); d1 = bw_den(a, (uint16_t)t1); }
            if (frac_ge(n1, d1, T->fnum, T->fden)) { if ((uint32_t)n_ans < cap) ans[n_ans] = m[z].locus; n_ans++; }
        }
    }
    /* ascending locus order */
    for (int x = 1; x < n_ans && x < (int)cap; x++) { uint32_t v = ans[x]; int y = x; while (y > 0 && ans[y - 1] > v) { ans[y] = ans[y - 1]; y--; } ans[y] = v; }
done:
    free(h); free(p); free(f); free(m);
    return n_ans;
}

/* recruit_long_read (recruit.rs:967-998) with has_matching_stretch (:932-964); single-end reads longer than
 * READ_LENGTH_THRESH.  The per-locus MinimInfo of locus_minimizers equals the (minimizer, locus) entry of minim_to_loci:
 * both merge every occurrence of the minimizer in the locus' sequences (:721-725). */
typedef struct { uint32_t locus; uint32_t arr[4]; } lmatch32;
static int recruit_one_long(const lcto_targets *T, const uint8_t *s1, size_t l1, uint32_t *ans, uint32_t cap) {
    const size_t mcap = l1 + 1;
    uint64_t *h = (uint64_t *)malloc(8 * mcap); uint32_t *p = (uint32_t *)malloc(4 * mcap); uint8_t *f = (uint8_t *)malloc(mcap);
    lmatch32 *m = NULL; size_t nm = 0, mc = 0;
    int n_ans = 0;
    const size_t total = lcto_minimizers(s1, l1, T->k, T->w, h, p, f, mcap);
    for (size_t q = 0; q < total; q++)
        for (size_t e = 0; e < T->n; e++) {
            if (T->e[e].key != h[q]) continue;
            size_t z = 0;
            while (z < nm && m[z].locus != T->e[e].locus) z++;
            if (z == nm) { if (nm == mc) { mc = mc ? 2 * mc : 8; m = (lmatch32 *)realloc(m, mc * sizeof(lmatch32)); } memset(&m[nm], 0, sizeof(lmatch32)); m[nm].locus = T->e[e].locus; nm++; }
            const int i = T->e[e].rare << 1;                          /* BaseMatchCount<u32>::inc, :246-252 */
            m[z].arr[i] += (uint32_t)directed_to(T->e[e].dir, !f[q]);
            m[z].arr[i | 1] += (uint32_t)directed_to(T->e[e].dir, f[q]);
        }
    for (size_t z = 0; z < nm; z++) {
        const uint32_t bw_c = m[z].arr[0], fw_c = m[z].arr[1], bw_r = m[z].arr[2], fw_r = m[z].arr[3];
        uint32_t num, den;                                             /* rare_fraction, :266-274 */
        if (fw_r >= bw_r) { num = fw_r; den = (uint32_t)total - fw_c; } else { num = bw_r; den = (uint32_t)total - bw_c; }
        const uint32_t nmin = T->stretch_minims < den ? T->stretch_minims : den;
        uint32_t thr = (uint32_t)ceil((double)nmin * T->match_frac);   /* long_read_threshold, :107-109 */
        if (thr < 1) thr = 1;
        if (num < thr) continue;
        int okay = den < T->stretch_minims;
        if (!okay) {                                                   /* has_matching_stretch, :944-963 */
            uint32_t s_fw = 0, s_bw = 0;
            for (size_t q = 0; q < total && !okay; q++) {
                for (size_t e = 0; e < T->n; e++)
                    if (T->e[e].key == h[q] && T->e[e].locus == m[z].locus) {
                        const uint32_t x = 1 + (uint32_t)T->e[e].rare * 3;
                        s_fw += (uint32_t)directed_to(T->e[e].dir, f[q]) * x;
                        s_bw += (uint32_t)directed_to(T->e[e].dir, !f[q]) * x;
                    }
                s_fw = s_fw >= 1 ? s_fw - 1 : 0;                        /* saturating_sub(SUBSUM_PENALTY) */
                s_bw = s_bw >= 1 ? s_bw - 1 : 0;
                if (s_fw >= T->stretch_score || s_bw >= T->stretch_score) okay = 1;
            }
        }
        if (okay) { if ((uint32_t)n_ans < cap) ans[n_ans] = m[z].locus; n_ans++; }
    }
    for (int x = 1; x < n_ans && x < (int)cap; x++) { uint32_t v = ans[x]; int y = x; while (y > 0 && ans[y - 1] > v) { ans[y] = ans[y - 1]; y--; } ans[y] = v; }
    free(h); free(p); free(f); free(m);
    return n_ans;
}

int lcto_recruit_short(const lcto_targets *T, const lcto_reads *R, uint32_t cap, uint32_t *ans_count, uint32_t *ans_locus) {
    for (uint64_t r = 0; r < R->n_reads; r++) {
        const uint8_t *s1 = R->seq1 + R->off1[r]; const size_t l1 = R->off1[r + 1] - R->off1[r];
        const uint8_t *s2 = R->seq2 ? R->seq2 + R->off2[r] : NULL; const size_t l2 = R->seq2 ? R->off2[r + 1] - R->off2[r] : 0;
        /* RecruitableRecord::recruit, recruit.rs:582-611: single records by length, pairs always as short pairs */
        const int n = (!s2 && l1 > 500) ? recruit_one_long(T, s1, l1, ans_locus + r * cap, cap)
                                        : recruit_one(T, s1, l1, s2, l2, ans_locus + r * cap, cap);
        if (n < 0) return -1;
        ans_count[r] = (uint32_t)n;
    }
    return 0;
}

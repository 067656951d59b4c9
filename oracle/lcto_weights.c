/* ORACLE (test infrastructure; never imported, linked or executed by the product path).
 *
 * CPU restatement of UniqueKmers (src/model/locs.rs:915-1003): the k-mers unique to a locus (canonical base-k k-mers of
 * the contig sequences whose off-target count is 0) and calculate_read_weight, the number of non-overlapping unique
 * k-mers of a read's ends turned into a weight.  k-mers: kmers::kmers::<u128, _, CANONICAL> (src/seq/kmers.rs:163-202).
 * Parity unpinned by the reference itself; pinned by a statement-by-statement Python transcription (Python integers as
 * u128, a set as the HashSet) and hand-checked cases in tests/test_weights.py.
 */
#include "lcto.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
#define UNDEF_KMER (~(u128)0)                                     /* Kmer::UNDEF = u128::MAX (kmers.rs:45) */

/* kmers::<u128, _, true>(seq, k, out): out[i - (k-1)] for every i + 1 >= k; returns the number written. */
static uint64_t canon_kmers(const uint8_t *seq, uint64_t len, uint32_t k, u128 *out) {
    const u128 mask = (((u128)1) << (2 * k)) - 1;                 /* create_mask, kmers.rs:48-52 */
    const uint32_t rv_shift = 2 * k - 2;
    u128 fw = 0, rv = 0;
    const uint64_t k_1 = k - 1;
    uint64_t reset = k_1, n = 0;
    for (uint64_t i = 0; i < len; i++) {
        uint32_t enc;
        switch (seq[i]) {
        case 'A': enc = 0; break;
        case 'C': enc = 1; break;
        case 'G': enc = 2; break;
        case 'T': enc = 3; break;
        default:
            reset = i + k;
            if (i + 1 >= k) out[n++] = UNDEF_KMER;
            continue;
        }
        fw = ((fw << 2) | enc) & mask;
        rv = (rv >> 2) | ((u128)(3 - enc) << rv_shift);
        if (i >= reset) out[n++] = rv < fw ? rv : fw;
        else if (i + 1 >= k) out[n++] = UNDEF_KMER;
    }
    return n;
}

/* A plain sorted array stands in for the HashSet. */
struct lcto_unique_kmers { uint32_t k; uint64_t n; u128 *keys; double weight_mult, weight_interc; };

static int cmp_u128(const void *a, const void *b) {
    const u128 x = *(const u128 *)a, y = *(const u128 *)b;
    return x < y ? -1 : x > y ? 1 : 0;
}
static int contains(const lcto_unique_kmers *u, u128 key) {
    uint64_t lo = 0, hi = u->n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) / 2;
        if (u->keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo < u->n && u->keys[lo] == key;
}

/* UniqueKmers::new (locs.rs:930-963).  Returns NULL on a count array whose length is not the number of k-mers (:944). */
lcto_unique_kmers *lcto_unique_kmers_build(const uint8_t *seqs, const uint64_t *seq_off, uint64_t n_seqs,
                                           const uint16_t *kmer_counts, const uint64_t *cnt_off, uint32_t k,
                                           uint16_t hard_threshold, uint16_t soft_threshold) {
    if (k <= 1 || k >= 64 || hard_threshold > soft_threshold) return NULL;      /* :937, :956, kmers.rs:50 */
    uint64_t total = 0;
    for (uint64_t s = 0; s < n_seqs; s++) total += seq_off[s + 1] - seq_off[s];
    u128 *keys = (u128 *)malloc(sizeof(u128) * (total ? total : 1)), *buf = (u128 *)malloc(sizeof(u128) * (total ? total : 1));
    uint64_t n = 0;
    for (uint64_t s = 0; s < n_seqs; s++) {
        const uint64_t m = canon_kmers(seqs + seq_off[s], seq_off[s + 1] - seq_off[s], k, buf);
        if (m != cnt_off[s + 1] - cnt_off[s]) { free(keys); free(buf); return NULL; }
        for (uint64_t q = 0; q < m; q++)
            if (kmer_counts[cnt_off[s] + q] == 0) keys[n++] = buf[q];             /* :946-947 (UNDEF included, like the reference) */
    }
    free(buf);
    qsort(keys, n, sizeof(u128), cmp_u128);
    uint64_t w = 0;
    for (uint64_t q = 0; q < n; q++) if (q == 0 || keys[q] != keys[q - 1]) keys[w++] = keys[q];
    lcto_unique_kmers *u = (lcto_unique_kmers *)malloc(sizeof(*u));
    u->k = k; u->n = w; u->keys = keys;
    u->weight_mult = 1.0 / (double)(soft_threshold + 1 - hard_threshold);          /* :957 */
    u->weight_interc = (1.0 - (double)hard_threshold) * u->weight_mult;            /* :958 */
    return u;
}
uint64_t lcto_unique_kmers_count(const lcto_unique_kmers *u) { return u->n; }
void lcto_unique_kmers_free(lcto_unique_kmers *u) { if (u) { free(u->keys); free(u); } }

/* calculate_read_weight (locs.rs:968-1002) for n_reads reads of `ends` read ends each (sequence e of read r =
 * seqs[seq_off[r * ends + e] .. seq_off[r * ends + e + 1]); an empty sequence = the mate is None).  unique[r * ends + e]
 * = MateData::unique_kmers, weight[r] = the factor read_data.weight is multiplied by. */
int lcto_read_weights(const lcto_unique_kmers *u, const uint8_t *seqs, const uint64_t *seq_off, uint64_t n_reads,
                      uint32_t ends, uint16_t *unique, double *weight) {
    uint64_t maxlen = 1;
    for (uint64_t q = 0; q < n_reads * ends; q++) maxlen = seq_off[q + 1] - seq_off[q] > maxlen ? seq_off[q + 1] - seq_off[q] : maxlen;
    u128 *buf = (u128 *)malloc(sizeof(u128) * maxlen);
    for (uint64_t r = 0; r < n_reads; r++) {
        uint16_t paired_count = 0;
        for (uint32_t e = 0; e < ends; e++) {
            const uint64_t q = r * ends + e, len = seq_off[q + 1] - seq_off[q];
            uint16_t count = 0;
            if (len) {
                const uint64_t m = canon_kmers(seqs + seq_off[q], len, u->k, buf);
                uint64_t it = 0;                                                   /* kmers_iter */
                while (it < m) {
                    const u128 kmer = buf[it++];
                    if (contains(u, kmer)) {
                        count = count == 65535 ? count : (uint16_t)(count + 1);    /* saturating_add */
                        it += (uint64_t)(u->k - 2) + 1;                            /* kmers_iter.nth(k_2): k - 1 elements */
                    }
                }
            }
            unique[q] = count;
            paired_count = (uint16_t)(paired_count + count);
        }
        double w = u->weight_interc + (double)paired_count * u->weight_mult;       /* :997 */
        w = w < 0.0 ? 0.0 : w > 1.0 ? 1.0 : w;                                     /* clamp(0.0, 1.0) */
        weight[r] = w;
    }
    free(buf);
    return 0;
}

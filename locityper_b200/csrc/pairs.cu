// pairs.cu -- SURVEY.md section 8(f) rank 1: mate alignments -> per-contig pair alignments, the producer
// of the flat locus' pa_* arrays (identify_paired_end_alignments + identify_contig_pair_alns,
// src/model/locs.rs:744-868).
//
// Work unit = one (read pair, contig) group: all first-end x second-end combinations of opposite strand
// ((ln1 + ln2) + insert-size ln-pmf, src/seq/aln.rs:236-238), the single-mate options that beat every
// pairing of that mate (locs.rs:777-790), a stable descending selection of the best `max_alns` within
// `prob_diff` of the best (locs.rs:793-798), scaled by the read weight (locs.rs:861-863).
//
// Device layout: groups are discovered on the device (head flag per mate record = first record of its read or
// contig change, exclusive scan -> dense group index, scatter of the group starts), then ONE THREAD PER GROUP
// evaluates its group, so every lane of a warp has work (the first version ran one thread per record with only
// the group's first record working: 5 of 32 lanes active).  Two passes over the same code: pass 0 counts the kept
// pairs of every group, an exclusive scan of the counts gives each group's output offset (and, read at the
// first group of read r, the pa_off array), pass 1 recomputes and writes.  The output order (read, contig) is
// the order of the group starts.  The byte traffic is a few streaming reads of the mate records (L2-resident at
// the sizes of the configs) plus the output; no host round trip between the passes except the total count.
#include "common.cuh"

#include <algorithm>
#include <cub/device/device_scan.cuh>

namespace lctp {

static constexpr int PAIR_MAX_ALNS = 16;     // per read end and contig; the reference uses 10 (locs.rs:741)

struct MatesDev {
    uint32_t R, max_alns, ins_len, single_end, window;
    uint64_t N;
    const uint64_t *ma_off;
    const uint32_t *ma_contig, *ma_start, *ma_end;
    const uint8_t *ma_flags;
    const double *ma_ln_prob, *read_weight, *ins_ln_pmf;
    const uint8_t *read_max;         // [R] per-read max_alns, or nullptr
    const uint64_t *exp_off;         // [H+1] explicit region weights per contig position, or nullptr
    const double *exp_weight;
    double unmapped_penalty, insert_penalty, prob_diff;
};

__device__ __forceinline__ long long pair_total_key(double v) {     // f64::total_cmp key
    long long b = __double_as_longlong(v);
    return b ^ (long long)(((unsigned long long)(b >> 63)) >> 1);
}

struct TopK {                       // stable descending top-`cap` list (cap <= PAIR_MAX_ALNS)
    double lp[PAIR_MAX_ALNS];
    uint32_t m1[PAIR_MAX_ALNS], m2[PAIR_MAX_ALNS];
    uint32_t n, total;
    __device__ void push(double p, uint32_t a, uint32_t b, uint32_t cap) {
        total++;
        const long long key = pair_total_key(p);
        uint32_t pos = n;
        while (pos > 0 && pair_total_key(lp[pos - 1]) < key) pos--;      // after every entry >= p: stable
        if (pos >= cap) return;
        const uint32_t last = n < cap ? n : cap - 1;
        for (uint32_t q = last; q > pos; q--) { lp[q] = lp[q - 1]; m1[q] = m1[q - 1]; m2[q] = m2[q - 1]; }
        lp[pos] = p; m1[pos] = a; m2[pos] = b;
        if (n < cap) n++;
    }
};

// head[i] = 1 when record i starts a (read, contig) group; read starts are added by k_pair_read_heads
__global__ void __launch_bounds__(256)
k_pair_heads(MatesDev D, uint32_t *__restrict__ head) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > D.N) return;
    head[i] = i < D.N && (i == 0 || D.ma_contig[i] != D.ma_contig[i - 1]) ? 1u : 0u;   // head[N] = 0 (scan sentinel)
}
__global__ void __launch_bounds__(256)
k_pair_read_heads(MatesDev D, uint32_t *__restrict__ head) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= D.R) return;
    const uint64_t m = D.ma_off[r];
    if (m < D.N && D.ma_off[r + 1] > m) head[m] = 1u;
}
// group_start[gidx[i]] = i for every head; group_start[G] = N
__global__ void __launch_bounds__(256)
k_pair_group_starts(MatesDev D, const uint32_t *__restrict__ head, const uint32_t *__restrict__ gidx,
                    uint32_t *__restrict__ group_start) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > D.N) return;
    if (i == D.N) group_start[gidx[D.N]] = (uint32_t)D.N;
    else if (head[i]) group_start[gidx[i]] = (uint32_t)i;
}

template <bool WRITE>
__global__ void __launch_bounds__(128)
k_pair_groups(MatesDev D, const uint32_t *__restrict__ group_start, const uint32_t *__restrict__ gidx,
              uint32_t *__restrict__ counts, const uint64_t *__restrict__ offs,
              uint32_t *__restrict__ pa_contig, double *__restrict__ pa_ln_prob, uint32_t *__restrict__ pa_mid1,
              uint32_t *__restrict__ pa_mid2, int *__restrict__ err) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= gidx[D.N]) return;                       // number of groups, still on the device
    const uint64_t i = group_start[g], e = group_start[g + 1];
    // read of this group: last r with ma_off[r] <= i
    uint32_t lo = 0, hi = D.R;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (D.ma_off[mid] <= i) lo = mid; else hi = mid;
    }
    const uint32_t r = lo;
    const uint64_t rb = D.ma_off[r];
    const uint32_t contig = D.ma_contig[i];
    if (i > rb && D.ma_contig[i - 1] > contig) atomicOr(err, 1);          // contigs must ascend within a read
    uint64_t f = i;
    while (f < e && (D.ma_flags[f] & 1u) == 0) f++;
    for (uint64_t q = f; q < e; q++) if ((D.ma_flags[q] & 1u) == 0) atomicOr(err, 2);   // first end before second
    const uint32_t M = D.read_max ? D.read_max[r] : D.max_alns;
    if (D.single_end) {
        // identify_single_end_alignments (locs.rs:870-911): the records of the contig in descending ln_prob; the best
        // one sets the threshold, at most max_alns within prob_diff of it are kept as "first mate only" pairs.
        if (f != e) atomicOr(err, 16);                                     // a second-end record in single-end data
        const double thresh = __dsub_rn(D.ma_ln_prob[i], D.prob_diff);
        uint32_t keep = 0;
        for (uint64_t q = i; q < e && keep < M; q++) {
            const double lp = D.ma_ln_prob[q];
            if (q > i && lp > D.ma_ln_prob[q - 1]) atomicOr(err, 4);
            if (!(lp >= thresh)) break;                                    // sorted: nothing further passes either
            if (WRITE) {
                const uint64_t o = offs[g] + keep;
                pa_contig[o] = contig; pa_ln_prob[o] = lp;
                pa_mid1[o] = (D.ma_start[q] + D.ma_end[q]) / 2; pa_mid2[o] = LCTP_NONE_U32;
            }
            keep++;
        }
        if (!WRITE) counts[g] = keep;
        return;
    }
    const uint32_t n1 = (uint32_t)min((uint64_t)M, f - i), n2 = (uint32_t)min((uint64_t)M, e - f);
    const double unm_ins = D.unmapped_penalty + D.insert_penalty;          // locs.rs:815-816

    if (f - i <= 1 && e - f <= 1 && M >= 3) {
        // The common group -- at most one alignment per end: at most three options (pair, first alone, second
        // alone), kept in registers.  Same pushes in the same order as the general path below.
        double lp0 = 0, lp1 = 0, lp2 = 0;
        uint32_t x0 = 0, x1 = 0, x2 = 0, y0 = 0, y1 = 0, y2 = 0, n = 0;
        auto push = [&](double p, uint32_t x, uint32_t y) {
            const long long key = pair_total_key(p);
            uint32_t pos = n;                                  // after every entry >= p: stable
            if (n == 2) { if (pair_total_key(lp1) < key) { pos = 1; if (pair_total_key(lp0) < key) pos = 0; } }
            else if (n == 1) { if (pair_total_key(lp0) < key) pos = 0; }
            if (pos <= 1 && n >= 2) { lp2 = lp1; x2 = x1; y2 = y1; }
            if (pos == 0 && n >= 1) { lp1 = lp0; x1 = x0; y1 = y0; }
            if (pos == 0) { lp0 = p; x0 = x; y0 = y; }
            else if (pos == 1) { lp1 = p; x1 = x; y1 = y; }
            else { lp2 = p; x2 = x; y2 = y; }
            n++;
        };
        const bool h1 = f > i, h2 = e > f;
        double l1 = 0, l2 = 0, max1 = -INFINITY, best2 = -INFINITY;
        uint32_t mid1 = LCTP_NONE_U32, mid2 = LCTP_NONE_U32, s1 = 0, e1 = 0, s2 = 0, e2 = 0;
        if (h1) { s1 = D.ma_start[i]; e1 = D.ma_end[i]; l1 = D.ma_ln_prob[i]; mid1 = (s1 + e1) / 2; }
        if (h2) { s2 = D.ma_start[f]; e2 = D.ma_end[f]; l2 = D.ma_ln_prob[f]; mid2 = (s2 + e2) / 2; }
        if (h1 && h2 && (D.ma_flags[f] & 2u) != (D.ma_flags[i] & 2u)) {
            const uint32_t insert = max(e1, e2) - min(s1, s2);
            if (insert >= D.ins_len) atomicOr(err, 8);
            else {
                const double prob = __dadd_rn(__dadd_rn(l1, l2), D.ins_ln_pmf[insert]);
                if (isfinite(prob)) { max1 = prob; best2 = prob; push(prob, mid1, mid2); }
            }
        }
        if (h1) { const double alone1 = __dadd_rn(l1, unm_ins); if (alone1 >= max1) push(alone1, mid1, LCTP_NONE_U32); }
        if (h2) { const double alone2 = __dadd_rn(l2, unm_ins); if (alone2 >= best2) push(alone2, LCTP_NONE_U32, mid2); }
        const double thresh = __dsub_rn(lp0, D.prob_diff);
        uint32_t keep = 0;
        if (n >= 1 && lp0 >= thresh) { keep = 1; if (n >= 2 && lp1 >= thresh) { keep = 2; if (n >= 3 && lp2 >= thresh) keep = 3; } }
        if (!WRITE) { counts[g] = keep; return; }
        // unscaled: the read's weight needs all of its kept pairs (k_pair_read_finish)
        const uint64_t o = offs[g];
        if (keep >= 1) { pa_contig[o] = contig; pa_ln_prob[o] = lp0; pa_mid1[o] = x0; pa_mid2[o] = y0; }
        if (keep >= 2) { pa_contig[o + 1] = contig; pa_ln_prob[o + 1] = lp1; pa_mid1[o + 1] = x1; pa_mid2[o + 1] = y1; }
        if (keep >= 3) { pa_contig[o + 2] = contig; pa_ln_prob[o + 2] = lp2; pa_mid1[o + 2] = x2; pa_mid2[o + 2] = y2; }
        return;
    }

    TopK top;
    top.n = 0; top.total = 0;
    double best2[PAIR_MAX_ALNS];
    for (uint32_t b = 0; b < n2; b++) best2[b] = -INFINITY;
    for (uint32_t a = 0; a < n1; a++) {
        const uint64_t ia = i + a;
        const uint32_t s1 = D.ma_start[ia], e1 = D.ma_end[ia];
        const double l1 = D.ma_ln_prob[ia];
        const uint32_t st1 = D.ma_flags[ia] & 2u;
        if (a > 0 && l1 > D.ma_ln_prob[ia - 1]) atomicOr(err, 4);          // ln_prob must descend within an end
        double max1 = -INFINITY;
        for (uint32_t b = 0; b < n2; b++) {
            const uint64_t ib = f + b;
            if ((D.ma_flags[ib] & 2u) == st1) continue;                    // aln.rs: strands must differ
            const uint32_t s2 = D.ma_start[ib], e2 = D.ma_end[ib];
            const uint32_t insert = max(e1, e2) - min(s1, s2);             // interv.rs:179-185
            if (insert >= D.ins_len) { atomicOr(err, 8); continue; }
            const double prob = __dadd_rn(__dadd_rn(l1, D.ma_ln_prob[ib]), D.ins_ln_pmf[insert]);
            if (isfinite(prob)) {
                max1 = fmax(max1, prob);
                best2[b] = fmax(best2[b], prob);
                top.push(prob, (s1 + e1) / 2, (s2 + e2) / 2, M);
            }
        }
        const double alone1 = __dadd_rn(l1, unm_ins);
        if (alone1 >= max1) top.push(alone1, (s1 + e1) / 2, LCTP_NONE_U32, M);
    }
    for (uint32_t b = 0; b < n2; b++) {
        const uint64_t ib = f + b;
        const double l2 = D.ma_ln_prob[ib];
        if (b > 0 && l2 > D.ma_ln_prob[ib - 1]) atomicOr(err, 4);
        const double alone2 = __dadd_rn(l2, unm_ins);
        if (alone2 >= best2[b]) top.push(alone2, LCTP_NONE_U32, (D.ma_start[ib] + D.ma_end[ib]) / 2, M);
    }
    // partition_point over the first min(len, max_alns) sorted entries (locs.rs:796-797)
    const double thresh = __dsub_rn(top.lp[0], D.prob_diff);
    uint32_t keep = 0;
    while (keep < top.n && top.lp[keep] >= thresh) keep++;
    if (!WRITE) { counts[g] = keep; return; }
    const uint64_t o = offs[g];
    for (uint32_t q = 0; q < keep; q++) {
        pa_contig[o + q] = contig;
        pa_ln_prob[o + q] = top.lp[q];
        pa_mid1[o + q] = top.m1[q];
        pa_mid2[o + q] = top.m2[q];
    }
}

// pa_off[r] = offs[first group at or after ma_off[r]]
__global__ void k_pair_read_outputs(MatesDev D, const uint32_t *__restrict__ gidx, const uint64_t *__restrict__ offs,
                                    uint64_t *__restrict__ pa_off) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > D.R) return;
    pa_off[r] = offs[gidx[D.ma_off[r]]];              // gidx[N] = number of groups, offs[G] = total
}

// ContigInfo::read_end_weight (src/model/windows.rs:495-504): the largest explicit weight at the middle of the read end
// and half a window to either side; an unmapped end weighs 0.
__device__ __forceinline__ double read_end_weight(const MatesDev &D, uint32_t contig, uint32_t middle) {
    if (middle == LCTP_NONE_U32) return 0.0;
    const double *w = D.exp_weight + D.exp_off[contig];
    const uint32_t n = (uint32_t)(D.exp_off[contig + 1] - D.exp_off[contig]);
    const uint32_t u = D.window / 2;
    return fmax(fmax(w[middle], w[middle > u ? middle - u : 0u]), w[min(middle + u, n - 1u)]);
}

// The tail of identify_paired_end_alignments / identify_single_end_alignments (locs.rs:860-867, 904-910), one thread
// per read: weight = ReadData::weight * ContigInfos::explicit_read_weight (windows.rs:683-693: the mean over the
// read's kept pair alignments of max(read_end_weight(mate 1), read_end_weight(mate 2)), 1 without explicit
// weights), every ln_prob *= weight, unmapped_prob = weight * (2 * unmapped_penalty + insert_penalty) for pairs,
// weight * unmapped_penalty for single-end reads.
__global__ void k_pair_read_finish(MatesDev D, const uint64_t *__restrict__ pa_off, const uint32_t *__restrict__ pa_contig,
                                   double *__restrict__ pa_ln_prob, const uint32_t *__restrict__ pa_mid1,
                                   const uint32_t *__restrict__ pa_mid2, double *__restrict__ unmapped_prob) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= D.R) return;
    const uint64_t b = pa_off[r], e = pa_off[r + 1];
    double expl = 1.0;
    if (D.exp_weight) {
        double s = 0.0;
        for (uint64_t q = b; q < e; q++)
            s = __dadd_rn(s, fmax(read_end_weight(D, pa_contig[q], pa_mid1[q]), read_end_weight(D, pa_contig[q], pa_mid2[q])));
        expl = __ddiv_rn(s, (double)(e - b));
    }
    const double weight = __dmul_rn(D.read_weight ? D.read_weight[r] : 1.0, expl);
    for (uint64_t q = b; q < e; q++) pa_ln_prob[q] = __dmul_rn(pa_ln_prob[q], weight);
    unmapped_prob[r] = D.single_end ? __dmul_rn(weight, D.unmapped_penalty)
                                    : __dmul_rn(weight, __dadd_rn(__dmul_rn(2.0, D.unmapped_penalty), D.insert_penalty));
}

template <typename T>
static int to_dev(DevBuf<T> &dst, const T *src, size_t n, cudaStream_t s) {
    int rc = dst.alloc(n);
    if (rc) return rc;
    if (n) LCTP_CUDA_CHECK(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return LCTP_OK;
}

// src != nullptr: the mate arrays are already on the device (lctp_group_reads_dev); `in` then only carries the
// parameters, the insert-size table, the explicit weights and the weights of the src->n_reads reads.
int pair_alignments_dev(lctp_ctx *ctx, const lctp_mates *in, lctp_pairs_h *P, const lctp_mates_h *src) {
    cudaStream_t s = ctx->stream;
    const uint32_t R = src ? src->n_reads : in->n_reads;
    if (R == 0 || (!src && !in->ma_off)) { set_error("lctp_pair_alignments: no reads"); return LCTP_E_INVALID; }
    if (!src && in->read_max_alns)
        for (uint32_t r = 0; r < in->n_reads; r++)
            if (in->read_max_alns[r] == 0 || in->read_max_alns[r] > PAIR_MAX_ALNS) {
                set_error("lctp_pair_alignments: read_max_alns[%u] = %u unsupported (1..=%d)", r, in->read_max_alns[r], PAIR_MAX_ALNS);
                return LCTP_E_INVALID;
            }
    if (in->max_alns == 0 || in->max_alns > (uint32_t)PAIR_MAX_ALNS) {
        set_error("lctp_pair_alignments: max_alns %u unsupported (1..=%d)", in->max_alns, PAIR_MAX_ALNS);
        return LCTP_E_CAPACITY;
    }
    if ((in->exp_weight != nullptr) != (in->exp_off != nullptr) || (in->exp_weight && in->window == 0)) {
        set_error("lctp_pair_alignments: explicit weights need exp_off, exp_weight and the window size");
        return LCTP_E_INVALID;
    }
    if (!in->single_end && !in->ins_ln_pmf) { set_error("lctp_pair_alignments: NULL insert-size table"); return LCTP_E_INVALID; }
    const uint64_t N = src ? src->n : in->ma_off[R];
    DevBuf<uint64_t> d_off, d_offs, d_exp_off;
    DevBuf<uint32_t> d_contig, d_start, d_end, d_counts, d_head, d_gidx, d_gstart;
    DevBuf<uint8_t> d_flags, d_rmax;
    DevBuf<double> d_lp, d_w, d_ins, d_exp;
    DevBuf<int> d_err;
    DevBuf<unsigned char> d_tmp;
    int rc;
    if (N >= 0xFFFFFFFFull) { set_error("lctp_pair_alignments: %llu mate records (limit 2^32 - 2)", (unsigned long long)N); return LCTP_E_CAPACITY; }
    uint64_t h2d = 0;
    if (!src) {
        if ((rc = to_dev(d_off, in->ma_off, (size_t)R + 1, s))) return rc;
        if ((rc = to_dev(d_contig, in->ma_contig, N, s))) return rc;
        if ((rc = to_dev(d_start, in->ma_start, N, s))) return rc;
        if ((rc = to_dev(d_end, in->ma_end, N, s))) return rc;
        if ((rc = to_dev(d_flags, in->ma_flags, N, s))) return rc;
        if ((rc = to_dev(d_lp, in->ma_ln_prob, N, s))) return rc;
        h2d += ((size_t)R + 1) * 8 + N * (4 + 4 + 4 + 1 + 8);
        if (in->read_max_alns) { if ((rc = to_dev(d_rmax, in->read_max_alns, R, s))) return rc; h2d += (uint64_t)R; }
    }
    if (in->read_weight) { if ((rc = to_dev(d_w, in->read_weight, R, s))) return rc; h2d += (uint64_t)R * 8; }
    if (!in->single_end) { if ((rc = to_dev(d_ins, in->ins_ln_pmf, in->ins_len, s))) return rc; h2d += (uint64_t)in->ins_len * 8; }
    if (in->exp_weight) {
        const uint64_t tot = in->exp_off[in->n_haps];
        if ((rc = to_dev(d_exp_off, in->exp_off, (size_t)in->n_haps + 1, s))) return rc;
        if ((rc = to_dev(d_exp, in->exp_weight, tot, s))) return rc;
        h2d += ((uint64_t)in->n_haps + 1) * 8 + tot * 8;
    }
    ctx->stats.h2d_bytes += h2d;
    if ((rc = d_head.alloc(N + 1))) return rc;
    if ((rc = d_gidx.alloc(N + 1))) return rc;
    if ((rc = d_gstart.alloc(N + 2))) return rc;
    if ((rc = d_counts.alloc(N + 1))) return rc;
    if ((rc = d_offs.alloc(N + 1))) return rc;
    if ((rc = d_err.alloc(1))) return rc;
    if ((rc = P->pa_off.alloc((size_t)R + 1))) return rc;
    if ((rc = P->unmapped.alloc(R))) return rc;
    LCTP_CUDA_CHECK(cudaMemsetAsync(d_err.p, 0, sizeof(int), s));
    LCTP_CUDA_CHECK(cudaMemsetAsync(d_counts.p, 0, (N + 1) * sizeof(uint32_t), s));

    MatesDev D;
    D.R = R; D.max_alns = in->max_alns; D.ins_len = in->ins_len; D.N = N;
    D.single_end = in->single_end ? 1u : 0u; D.window = in->window;
    D.ma_off = d_off.p; D.ma_contig = d_contig.p; D.ma_start = d_start.p; D.ma_end = d_end.p;
    D.ma_flags = d_flags.p; D.ma_ln_prob = d_lp.p; D.read_weight = in->read_weight ? d_w.p : nullptr;
    D.ins_ln_pmf = d_ins.p;
    D.read_max = in->read_max_alns ? d_rmax.p : nullptr;
    if (src) {
        D.ma_off = src->ma_off.p; D.ma_contig = src->contig.p; D.ma_start = src->start.p; D.ma_end = src->end.p;
        D.ma_flags = src->flags.p; D.ma_ln_prob = src->lnprob.p; D.read_max = src->max_alns.p;
    }
    D.exp_off = in->exp_weight ? d_exp_off.p : nullptr; D.exp_weight = in->exp_weight ? d_exp.p : nullptr;
    D.unmapped_penalty = in->unmapped_penalty; D.insert_penalty = in->insert_penalty; D.prob_diff = in->prob_diff;

    uint64_t total = 0;
    {
        const unsigned grid_n = (unsigned)((N + 1 + 255) / 256);      // threads over records (+ sentinel)
        const unsigned grid_g = (unsigned)((N + 127) / 128);          // threads over groups (#groups <= N)
        size_t tmp32 = 0, tmp64 = 0;
        LCTP_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp32, d_head.p, d_gidx.p, (int)(N + 1), s));
        LCTP_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp64, d_counts.p, d_offs.p, (int)(N + 1), s));
        if ((rc = d_tmp.alloc(std::max(tmp32, tmp64)))) return rc;
        LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[0], s));
        k_pair_heads<<<grid_n, 256, 0, s>>>(D, d_head.p);
        k_pair_read_heads<<<(R + 255) / 256, 256, 0, s>>>(D, d_head.p);
        LCTP_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp32, d_head.p, d_gidx.p, (int)(N + 1), s));
        k_pair_group_starts<<<grid_n, 256, 0, s>>>(D, d_head.p, d_gidx.p, d_gstart.p);
        ctx->launches += 4;
        if (N) {
            k_pair_groups<false><<<grid_g, 128, 0, s>>>(D, d_gstart.p, d_gidx.p, d_counts.p, nullptr, nullptr, nullptr, nullptr, nullptr, d_err.p);
            ctx->launches++;
        }
        LCTP_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tmp64, d_counts.p, d_offs.p, (int)(N + 1), s));
        ctx->launches++;
        LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[1], s));      // phase A: groups + counts; the host sizes the output next
        LCTP_CUDA_CHECK(cudaMemcpyAsync(&total, d_offs.p + N, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        int err = 0;
        LCTP_CUDA_CHECK(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
        ctx->stats.d2h_bytes += 12;
        if (err) {
            set_error("lctp_pair_alignments: malformed mate alignments (flags=%d: 1=contigs not ascending within a "
                      "read, 2=second-end record before a first-end record of the same contig, 4=ln_prob not "
                      "descending within a (read, contig, end) run, 8=insert size outside the ln-pmf table, "
                      "16=second-end record in single-end input)", err);
            return LCTP_E_INVALID;
        }
        if ((rc = P->contig.alloc(total))) return rc;
        if ((rc = P->lnprob.alloc(total))) return rc;
        if ((rc = P->mid1.alloc(total))) return rc;
        if ((rc = P->mid2.alloc(total))) return rc;
        LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[2], s));      // phase B: write pass + per-read weights
        if (N) {
            k_pair_groups<true><<<grid_g, 128, 0, s>>>(D, d_gstart.p, d_gidx.p, nullptr, d_offs.p, P->contig.p, P->lnprob.p, P->mid1.p, P->mid2.p, d_err.p);
            ctx->launches++;
        }
        k_pair_read_outputs<<<(R + 1 + 255) / 256, 256, 0, s>>>(D, d_gidx.p, d_offs.p, P->pa_off.p);
        k_pair_read_finish<<<(R + 255) / 256, 256, 0, s>>>(D, P->pa_off.p, P->contig.p, P->lnprob.p, P->mid1.p, P->mid2.p, P->unmapped.p);
        ctx->launches += 2;
        LCTP_CUDA_CHECK(cudaEventRecord(ctx->ev[3], s));
    }
    LCTP_CUDA_CHECK(cudaGetLastError());
    LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
    {
        float ms = 0.f, ms2 = 0.f;
        LCTP_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        LCTP_CUDA_CHECK(cudaEventElapsedTime(&ms2, ctx->ev[2], ctx->ev[3]));
        ctx->stats.pairing_ms += ms + ms2;
        ctx->stats.pairing_launches += 1;
        ctx->stats.pairing_mates += N;
        ctx->stats.pairing_pairs += total;
    }
    P->ctx = ctx; P->n_reads = R; P->n_pairs = total; P->n_haps = in->n_haps;
    return LCTP_OK;
}

int pairs_fetch(lctp_pairs_h *P, uint64_t cap, uint64_t *pa_off, uint32_t *pa_contig, double *pa_ln_prob,
                uint32_t *pa_mid1, uint32_t *pa_mid2, double *unmapped_prob) {
    cudaStream_t s = P->ctx->stream;
    const uint64_t total = P->n_pairs;
    if (total > cap) {
        set_error("lctp_pair_alignments: output capacity %llu too small (%llu pair alignments)",
                  (unsigned long long)cap, (unsigned long long)total);
        return LCTP_E_CAPACITY;
    }
    LCTP_CUDA_CHECK(cudaMemcpyAsync(pa_off, P->pa_off.p, ((size_t)P->n_reads + 1) * 8, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(unmapped_prob, P->unmapped.p, (size_t)P->n_reads * 8, cudaMemcpyDeviceToHost, s));
    if (total) {
        LCTP_CUDA_CHECK(cudaMemcpyAsync(pa_contig, P->contig.p, total * 4, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(pa_ln_prob, P->lnprob.p, total * 8, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(pa_mid1, P->mid1.p, total * 4, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(pa_mid2, P->mid2.p, total * 4, cudaMemcpyDeviceToHost, s));
    }
    LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
    P->ctx->stats.d2h_bytes += ((size_t)P->n_reads + 1) * 8 + (size_t)P->n_reads * 8 + total * 20;
    return LCTP_OK;
}

}  // namespace lctp

// group.cu -- SURVEY.md section 8(f) rank 1, remainder: from the per read-end results (lctp_collect_read_ends) to the
// input of the pairing (lctp_pair_alignments*), for every read of a locus at once:
//   AllAlignments::load after read_next_alns (src/model/locs.rs:1117-1137): well_mapped over the read's ends, in_bounds
//   (:1008-1014) over PrelimAlignments::alns;
//   recover_and_group_alignments without the alignment transfer (:1237-1288; opt_hap_alns = None, the transfer needs
//   WFA2): best_edit_is_good (:293-295), normalize_probs (:358-360), MAX_USED_ALNS / MAX_UNUSED_ALNS (:739-742, 1263),
//   and the order in which identify_paired_end_alignments / identify_single_end_alignments consume the alignments after
//   their sorts (:819-820, 883).
// Three launches: one thread per read decides its status and counts its kept alignments; exclusive scans (cub) give
// every passing read its place and its first output entry; one thread per read then writes its entries in consumption
// order (an insertion sort of a handful of entries: contig ascending, first end before second, ln_prob descending, ties
// in the order of PrelimAlignments::alns).  Integer / byte work, a few bytes per alignment: bound by the copies.
#include "common.cuh"

#include <cub/device/device_scan.cuh>

namespace lctp {

struct PrelimDev {
    uint64_t n_reads;
    const int64_t *read_group;
    const uint64_t *grp_off;
    const uint32_t *rec_contig, *rec_start, *rec_end;
    const uint8_t *rec_strand;
    const double *rec_ln_prob;
    const uint8_t *grp_ok;
    const uint32_t *best_edit, *thr_dist, *n_kept, *kept_rec, *contig_len;
    const double *read_weight;
    double min_weight;
    uint32_t boundary, single_end;
};

__global__ void __launch_bounds__(256)
k_group_status(PrelimDev D, uint8_t *__restrict__ status, uint32_t *__restrict__ pass, uint64_t *__restrict__ n_ent) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > D.n_reads) return;
    if (r == D.n_reads) { pass[r] = 0; n_ent[r] = 0; return; }           // scan sentinel
    const int64_t g0 = D.read_group[2 * r], g1 = D.read_group[2 * r + 1];
    bool well = g0 >= 0 && D.grp_ok[g0];
    if (!D.single_end && well) well = g1 >= 0 && D.grp_ok[g1];
    uint8_t st = 0;
    uint64_t cnt = 0;
    if (!well) st = 1;
    else {
        const int n_ends = D.single_end ? 1 : 2;
        bool inb = false, good = true;
        for (int e = 0; e < n_ends; e++) {
            const int64_t g = e ? g1 : g0;
            const uint64_t b = D.grp_off[g];
            const uint32_t nk = D.n_kept[g];
            for (uint32_t k = 0; k < nk; k++) {
                const uint32_t rec = D.kept_rec[b + k];
                const uint32_t clen = D.contig_len[D.rec_contig[rec]];
                const uint32_t mid = (D.rec_start[rec] + D.rec_end[rec]) / 2;              // Interval::middle
                inb |= D.boundary <= mid && mid < clen - D.boundary;
            }
            good &= D.best_edit[g] <= D.thr_dist[g];
            cnt += nk;
        }
        st = !inb ? 2 : !good ? 3 : 0;
    }
    status[r] = st;
    pass[r] = st == 0 ? 1u : 0u;
    n_ent[r] = st == 0 ? cnt : 0;
}

__global__ void __launch_bounds__(128)
k_group_write(PrelimDev D, const uint8_t *__restrict__ status, const uint32_t *__restrict__ place,
              const uint64_t *__restrict__ first, uint32_t *__restrict__ out_read, uint8_t *__restrict__ out_max,
              uint64_t *__restrict__ ma_off, uint32_t *__restrict__ ma_contig, uint8_t *__restrict__ ma_flags,
              uint32_t *__restrict__ ma_start, uint32_t *__restrict__ ma_end, double *__restrict__ ma_lp,
              uint32_t *__restrict__ ma_rec) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > D.n_reads) return;
    if (r == D.n_reads) { ma_off[place[r]] = first[r]; return; }          // end sentinel of the last passing read
    if (status[r] != 0) return;
    const uint32_t k_out = place[r];
    const uint64_t start = first[r];
    out_read[k_out] = (uint32_t)r;
    out_max[k_out] = D.read_weight[r] >= D.min_weight ? 10 : 2;           // MAX_USED_ALNS / MAX_UNUSED_ALNS
    ma_off[k_out] = start;
    uint64_t n = start;
    const int n_ends = D.single_end ? 1 : 2;
    for (int e = 0; e < n_ends; e++) {
        const int64_t g = D.read_group[2 * r + e];
        const uint64_t b = D.grp_off[g], ge = D.grp_off[g + 1];
        double best = -INFINITY;                                           // best_lik: every pushed alignment of the end
        for (uint64_t q = b; q < ge; q++) best = fmax(best, D.rec_ln_prob[q]);
        const uint32_t nk = D.n_kept[g];
        for (uint32_t k = 0; k < nk; k++) {
            const uint32_t rec = D.kept_rec[b + k];
            const uint32_t contig = D.rec_contig[rec];
            const double lp = __dsub_rn(D.rec_ln_prob[rec], best);         // normalize_probs
            uint64_t pos = n;
            while (pos > start) {                                          // the new entry goes after equal keys
                const uint32_t pc = ma_contig[pos - 1], pe = ma_flags[pos - 1] & 1u;
                const double pl = ma_lp[pos - 1];
                const bool before = contig != pc ? contig < pc : (uint32_t)e != pe ? (uint32_t)e < pe : lp > pl;
                if (!before) break;
                ma_contig[pos] = pc; ma_flags[pos] = ma_flags[pos - 1];
                ma_start[pos] = ma_start[pos - 1]; ma_end[pos] = ma_end[pos - 1];
                ma_lp[pos] = pl; ma_rec[pos] = ma_rec[pos - 1];
                pos--;
            }
            ma_contig[pos] = contig;
            ma_flags[pos] = (uint8_t)(e | (D.rec_strand[rec] ? 2 : 0));
            ma_start[pos] = D.rec_start[rec]; ma_end[pos] = D.rec_end[rec];
            ma_lp[pos] = lp; ma_rec[pos] = rec;
            n++;
        }
    }
}

template <typename T>
static int up(DevBuf<T> &dst, const T *src, size_t n, cudaStream_t s) {
    int rc = dst.alloc(n ? n : 1);
    if (rc) return rc;
    if (n) LCTP_CUDA_CHECK(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return LCTP_OK;
}

// Device-side results of one run: everything lctp_group_reads returns, still in device memory.
struct GroupDev {
    DevBuf<uint8_t> status, omax, mfl;
    DevBuf<uint32_t> oread, mcon, mst, men, mrec;
    DevBuf<uint64_t> maoff;
    DevBuf<double> mlp;
    uint32_t n_pass = 0;
    uint64_t n_ma = 0;
};

static int group_reads_run(lctp_ctx *ctx, const lctp_prelim *in, uint8_t *status, GroupDev &g) {
    cudaStream_t s = ctx->stream;
    const uint64_t R = in->n_reads, G = in->n_groups;
    const uint64_t N = in->grp_off[G];
    for (uint64_t r = 0; r < 2 * R; r++)
        if (in->read_group[r] >= (int64_t)G) { set_error("lctp_group_reads: read_group entry %llu out of range", (unsigned long long)r); return LCTP_E_INVALID; }
    DevBuf<int64_t> d_rg;
    DevBuf<uint64_t> d_goff, d_nent, d_first;
    DevBuf<uint32_t> d_con, d_st, d_en, d_be, d_thr, d_nk, d_kept, d_clen, d_pass, d_place;
    DevBuf<uint8_t> d_str, d_ok;
    DevBuf<double> d_lp, d_w;
    DevBuf<unsigned char> d_tmp;
    int rc;
    if ((rc = up(d_rg, in->read_group, 2 * R, s)) || (rc = up(d_goff, in->grp_off, G + 1, s)) ||
        (rc = up(d_con, in->rec_contig, N, s)) || (rc = up(d_st, in->rec_start, N, s)) || (rc = up(d_en, in->rec_end, N, s)) ||
        (rc = up(d_str, in->rec_strand, N, s)) || (rc = up(d_lp, in->rec_ln_prob, N, s)) || (rc = up(d_ok, in->grp_ok, G, s)) ||
        (rc = up(d_be, in->grp_best_edit, G, s)) || (rc = up(d_thr, in->grp_thr_dist, G, s)) ||
        (rc = up(d_nk, in->grp_n_kept, G, s)) || (rc = up(d_kept, in->kept_rec, N, s)) ||
        (rc = up(d_clen, in->contig_len, in->n_haps, s)) || (rc = up(d_w, in->read_weight, R, s))) return rc;
    ctx->stats.h2d_bytes += 2 * R * 8 + (G + 1) * 8 + N * (4 + 4 + 4 + 1 + 8 + 4) + G * (1 + 4 + 4 + 4) + in->n_haps * 4 + R * 8;
    if ((rc = d_nent.alloc(R + 1)) || (rc = d_first.alloc(R + 1)) || (rc = d_pass.alloc(R + 1)) || (rc = d_place.alloc(R + 1)) ||
        (rc = g.status.alloc(R)) || (rc = g.oread.alloc(R)) || (rc = g.omax.alloc(R)) || (rc = g.maoff.alloc(R + 1))) return rc;
    const size_t M = (size_t)std::max<uint64_t>(1, N);
    if ((rc = g.mcon.alloc(M)) || (rc = g.mfl.alloc(M)) || (rc = g.mst.alloc(M)) || (rc = g.men.alloc(M)) ||
        (rc = g.mlp.alloc(M)) || (rc = g.mrec.alloc(M))) return rc;
    PrelimDev D;
    D.n_reads = R; D.read_group = d_rg.p; D.grp_off = d_goff.p; D.rec_contig = d_con.p; D.rec_start = d_st.p;
    D.rec_end = d_en.p; D.rec_strand = d_str.p; D.rec_ln_prob = d_lp.p; D.grp_ok = d_ok.p; D.best_edit = d_be.p;
    D.thr_dist = d_thr.p; D.n_kept = d_nk.p; D.kept_rec = d_kept.p; D.contig_len = d_clen.p; D.read_weight = d_w.p;
    D.min_weight = in->min_weight; D.boundary = in->boundary; D.single_end = in->single_end;
    k_group_status<<<(unsigned)((R + 1 + 255) / 256), 256, 0, s>>>(D, g.status.p, d_pass.p, d_nent.p);
    ctx->launches++;
    LCTP_CUDA_CHECK(cudaGetLastError());
    size_t t32 = 0, t64 = 0;
    LCTP_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, t32, d_pass.p, d_place.p, (int)(R + 1), s));
    LCTP_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, t64, d_nent.p, d_first.p, (int)(R + 1), s));
    if ((rc = d_tmp.alloc(std::max(t32, t64)))) return rc;
    LCTP_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp.p, t32, d_pass.p, d_place.p, (int)(R + 1), s));
    LCTP_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp.p, t64, d_nent.p, d_first.p, (int)(R + 1), s));
    ctx->launches += 2;
    k_group_write<<<(unsigned)((R + 1 + 127) / 128), 128, 0, s>>>(D, g.status.p, d_place.p, d_first.p, g.oread.p, g.omax.p,
                                                                  g.maoff.p, g.mcon.p, g.mfl.p, g.mst.p, g.men.p, g.mlp.p,
                                                                  g.mrec.p);
    ctx->launches++;
    LCTP_CUDA_CHECK(cudaGetLastError());
    LCTP_CUDA_CHECK(cudaMemcpyAsync(&g.n_pass, d_place.p + R, 4, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(&g.n_ma, d_first.p + R, 8, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaMemcpyAsync(status, g.status.p, R, cudaMemcpyDeviceToHost, s));
    LCTP_CUDA_CHECK(cudaStreamSynchronize(s));       // the inputs (local buffers) are released after this point
    ctx->stats.d2h_bytes += R + 12;
    return LCTP_OK;
}

static void count_status(const uint8_t *status, uint64_t R, uint32_t n_pass, uint64_t *counts) {
    counts[0] = counts[1] = 0;
    for (uint64_t r = 0; r < R; r++) {
        if (status[r] == 2) counts[1]++;
        else if (status[r] != 0) counts[0]++;
    }
    counts[2] = n_pass;
}

// false = nothing to run (no reads / no groups); the outputs are then complete
static bool group_trivial(const lctp_prelim *in, uint8_t *status, uint64_t *n_reads_out, uint64_t *counts) {
    *n_reads_out = 0;
    counts[0] = counts[1] = counts[2] = 0;
    if (in->n_reads == 0) return true;
    if (in->n_groups == 0 || !in->grp_off || !in->read_group) {
        for (uint64_t r = 0; r < in->n_reads; r++) status[r] = 1;
        counts[0] = in->n_reads;
        return true;
    }
    return false;
}

int group_reads(lctp_ctx *ctx, const lctp_prelim *in, uint64_t cap, uint8_t *status, uint64_t *n_reads_out,
                uint32_t *out_read, uint8_t *out_max_alns, uint64_t *ma_off, uint32_t *ma_contig, uint8_t *ma_flags,
                uint32_t *ma_start, uint32_t *ma_end, double *ma_ln_prob, uint32_t *ma_rec, uint64_t *counts) {
    cudaStream_t s = ctx->stream;
    ma_off[0] = 0;
    if (group_trivial(in, status, n_reads_out, counts)) return LCTP_OK;
    GroupDev g;
    int rc = group_reads_run(ctx, in, status, g);
    if (rc) return rc;
    const uint32_t n_pass = g.n_pass;
    const uint64_t n_ma = g.n_ma;
    if (n_ma > cap) {
        set_error("lctp_group_reads: %llu entries, capacity %llu", (unsigned long long)n_ma, (unsigned long long)cap);
        return LCTP_E_CAPACITY;
    }
    if (n_pass) {
        LCTP_CUDA_CHECK(cudaMemcpyAsync(out_read, g.oread.p, (size_t)n_pass * 4, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(out_max_alns, g.omax.p, n_pass, cudaMemcpyDeviceToHost, s));
    }
    LCTP_CUDA_CHECK(cudaMemcpyAsync(ma_off, g.maoff.p, ((size_t)n_pass + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (n_ma) {
        LCTP_CUDA_CHECK(cudaMemcpyAsync(ma_contig, g.mcon.p, n_ma * 4, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(ma_flags, g.mfl.p, n_ma, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(ma_start, g.mst.p, n_ma * 4, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(ma_end, g.men.p, n_ma * 4, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(ma_ln_prob, g.mlp.p, n_ma * 8, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaMemcpyAsync(ma_rec, g.mrec.p, n_ma * 4, cudaMemcpyDeviceToHost, s));
    }
    LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
    ctx->stats.d2h_bytes += (uint64_t)n_pass * 5 + ((uint64_t)n_pass + 1) * 8 + n_ma * 25;
    count_status(status, in->n_reads, n_pass, counts);
    *n_reads_out = n_pass;
    return LCTP_OK;
}

// Same, leaving the pairing input on the device (lctp_mates_h); only the status, the read numbers and the counts
// cross to the host.
int group_reads_dev(lctp_ctx *ctx, const lctp_prelim *in, uint8_t *status, uint64_t *n_reads_out, uint32_t *out_read,
                    uint64_t *counts, lctp_mates_h *out) {
    cudaStream_t s = ctx->stream;
    out->ctx = ctx; out->n_reads = 0; out->n = 0;
    if (group_trivial(in, status, n_reads_out, counts)) return LCTP_OK;
    GroupDev g;
    int rc = group_reads_run(ctx, in, status, g);
    if (rc) return rc;
    if (g.n_pass && out_read) {
        LCTP_CUDA_CHECK(cudaMemcpyAsync(out_read, g.oread.p, (size_t)g.n_pass * 4, cudaMemcpyDeviceToHost, s));
        LCTP_CUDA_CHECK(cudaStreamSynchronize(s));
        ctx->stats.d2h_bytes += (uint64_t)g.n_pass * 4;
    }
    out->n_reads = g.n_pass; out->n = g.n_ma;
    out->ma_off.take(g.maoff); out->contig.take(g.mcon); out->start.take(g.mst); out->end.take(g.men);
    out->rec.take(g.mrec); out->flags.take(g.mfl); out->max_alns.take(g.omax); out->lnprob.take(g.mlp);
    count_status(status, in->n_reads, g.n_pass, counts);
    *n_reads_out = g.n_pass;
    return LCTP_OK;
}

}  // namespace lctp

extern "C" int lctp_group_reads(lctp_ctx *ctx, const lctp_prelim *in, uint64_t cap, uint8_t *status,
                                uint64_t *n_reads_out, uint32_t *out_read, uint8_t *out_max_alns, uint64_t *ma_off,
                                uint32_t *ma_contig, uint8_t *ma_flags, uint32_t *ma_start, uint32_t *ma_end,
                                double *ma_ln_prob, uint32_t *ma_rec, uint64_t *counts) {
    if (!ctx || !in || !status || !n_reads_out || !ma_off || !counts) { lctp::set_error("lctp_group_reads: NULL argument"); return LCTP_E_INVALID; }
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    lctp::set_alloc_stream(ctx->stream);
    return lctp::group_reads(ctx, in, cap, status, n_reads_out, out_read, out_max_alns, ma_off, ma_contig, ma_flags, ma_start,
                             ma_end, ma_ln_prob, ma_rec, counts);
}
extern "C" size_t lctp_sizeof_prelim(void) { return sizeof(lctp_prelim); }

extern "C" int lctp_group_reads_dev(lctp_ctx *ctx, const lctp_prelim *in, uint8_t *status, uint64_t *n_reads_out,
                                    uint32_t *out_read, uint64_t *counts, lctp_mates_h **out) {
    if (!ctx || !in || !status || !n_reads_out || !counts || !out) { lctp::set_error("lctp_group_reads_dev: NULL argument"); return LCTP_E_INVALID; }
    *out = nullptr;
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    lctp::set_alloc_stream(ctx->stream);
    lctp_mates_h *m = new lctp_mates_h();
    const int rc = lctp::group_reads_dev(ctx, in, status, n_reads_out, out_read, counts, m);
    if (rc) { delete m; return rc; }
    *out = m;
    return LCTP_OK;
}
extern "C" uint64_t lctp_mates_count(const lctp_mates_h *m) { return m ? m->n : 0; }
extern "C" void lctp_mates_free(lctp_mates_h *m) {
    if (!m) return;
    if (m->ctx) cudaSetDevice(m->ctx->device);
    delete m;
}
extern "C" int lctp_pair_alignments_from(lctp_ctx *ctx, const lctp_mates_h *mates, const lctp_mates *params,
                                         lctp_pairs_h **out, uint64_t *n_out) {
    if (!ctx || !mates || !params || !out) { lctp::set_error("lctp_pair_alignments_from: NULL argument"); return LCTP_E_INVALID; }
    *out = nullptr;
    if (mates->ctx != ctx) { lctp::set_error("lctp_pair_alignments_from: the mates belong to another context"); return LCTP_E_INVALID; }
    if (!params->single_end && (!params->ins_ln_pmf || params->ins_len == 0)) {
        lctp::set_error("lctp_pair_alignments_from: NULL insert-size table");
        return LCTP_E_INVALID;
    }
    LCTP_CUDA_CHECK(cudaSetDevice(ctx->device));
    lctp::set_alloc_stream(ctx->stream);
    lctp_pairs_h *p = new lctp_pairs_h();
    const int rc = lctp::pair_alignments_dev(ctx, params, p, mates);
    if (rc) { delete p; return rc; }
    if (n_out) *n_out = p->n_pairs;
    *out = p;
    return LCTP_OK;
}

"""SURVEY 8(f) rank 1, remainder: from the per read-end results to the pairing input (lctp_group_reads).

CPU: the oracle (oracle/lcto_group.c) against a statement-by-statement Python transcription of the cited Rust
(AllAlignments::load src/model/locs.rs:1117-1137, in_bounds :1008-1014, recover_and_group_alignments :1237-1288 without
the transfer, best_edit_is_good :293-295, normalize_probs :358-360, the sorts of identify_paired_end_alignments :819-820 and
identify_single_end_alignments :883) and a hand-checked case.  GPU: lctp_group_reads against the oracle, bit for bit, and
its output fed to lctp_pair_alignments with per-read max_alns against the oracle's pairing of the same input."""
import functools

import numpy as np
import pytest

from locityper_b200 import genotype

U32_MAX = 0xFFFFFFFF


def _random_prelim(seed, n_reads=300, n_haps=12, single_end=False, tie_frac=0.2):
    """Random per read-end results: group records in arbitrary order, a subset kept, some ends missing / not ok."""
    rng = np.random.default_rng(seed)
    clen = rng.integers(3000, 6000, n_haps).astype(np.uint32)
    read_group = np.full((n_reads, 2), -1, dtype=np.int64)
    grp_off, con, st, en, strand, lp = [0], [], [], [], [], []
    ok, best, thr, nk, kept = [], [], [], [], []
    for r in range(n_reads):
        for e in range(1 if single_end else 2):
            if rng.random() < 0.06:
                continue                                          # unmapped read end: no group
            g = len(ok)
            read_group[r, e] = g
            n = int(rng.integers(1, 9))
            base = len(con)
            vals = np.round(rng.uniform(-60, -1, n), 1) if rng.random() < tie_frac else rng.uniform(-60, -1, n)
            for k in range(n):
                c = int(rng.integers(0, n_haps))
                s = int(rng.integers(0, clen[c] - 200))
                con.append(c); st.append(s); en.append(s + int(rng.integers(50, 200)))
                strand.append(int(rng.integers(0, 2))); lp.append(float(vals[k]))
            grp_off.append(len(con))
            is_ok = rng.random() > 0.08
            ok.append(int(is_ok))
            t = int(rng.integers(3, 12))
            thr.append(t)
            best.append(int(rng.integers(0, t + 3)) if is_ok else int(rng.integers(0, 30)))
            sel = np.sort(rng.choice(n, size=int(rng.integers(1, n + 1)), replace=False)) if is_ok else np.array([], dtype=int)
            order = rng.permutation(sel)                          # PrelimAlignments::alns is not in record order
            nk.append(len(order))
            k_full = np.full(n, U32_MAX, dtype=np.uint32)
            k_full[:len(order)] = base + order
            kept += list(k_full)
    return genotype.Prelim(read_group=read_group, grp_off=np.array(grp_off, dtype=np.uint64),
                           rec_contig=np.array(con, dtype=np.uint32), rec_start=np.array(st, dtype=np.uint32),
                           rec_end=np.array(en, dtype=np.uint32), rec_strand=np.array(strand, dtype=np.uint8),
                           rec_ln_prob=np.array(lp), grp_ok=np.array(ok, dtype=np.uint8),
                           grp_best_edit=np.array(best, dtype=np.uint32), grp_thr_dist=np.array(thr, dtype=np.uint32),
                           grp_n_kept=np.array(nk, dtype=np.uint32), kept_rec=np.array(kept, dtype=np.uint32),
                           contig_len=clen, read_weight=rng.uniform(0.0, 1.0, n_reads), min_weight=0.3,
                           boundary=1300, single_end=single_end)


def _transcription(p: genotype.Prelim) -> dict:
    """The cited Rust, statement by statement, on Python objects."""
    status, out_read, out_max, ma = [], [], [], []
    counts = [0, 0, 0]
    for r in range(p.n_reads):
        g0, g1 = int(p.read_group[r, 0]), int(p.read_group[r, 1])
        # load(): read_next_alns(First); if is_paired_end && well_mapped: read_next_alns(Second)      (:1119-1134)
        well_mapped = g0 >= 0 and bool(p.grp_ok[g0])
        if not p.single_end and well_mapped:
            well_mapped = g1 >= 0 and bool(p.grp_ok[g1])
        if not well_mapped:
            status.append(1); counts[0] += 1
            continue
        groups = [g0] if p.single_end else [g0, g1]
        # PrelimAlignments of the read: alns (kept, both ends), good_dist, best_edit, best_lik per end
        alns, good_dist, best_edit, best_lik = [], [U32_MAX, U32_MAX], [U32_MAX, U32_MAX], [-np.inf, -np.inf]
        for e, g in enumerate(groups):
            b, ge = int(p.grp_off[g]), int(p.grp_off[g + 1])
            good_dist[e], best_edit[e] = int(p.grp_thr_dist[g]), int(p.grp_best_edit[g])
            best_lik[e] = max(float(v) for v in p.rec_ln_prob[b:ge])                # push, :311: every alignment
            for k in range(int(p.grp_n_kept[g])):
                rec = int(p.kept_rec[b + k])
                alns.append(dict(rec=rec, end=e, contig=int(p.rec_contig[rec]), start=int(p.rec_start[rec]),
                                 stop=int(p.rec_end[rec]), strand=int(p.rec_strand[rec]), ln_prob=float(p.rec_ln_prob[rec])))
        # in_bounds (:1008-1014)
        def inb(a):
            middle = (a["start"] + a["stop"]) // 2
            return p.boundary <= middle and middle < int(p.contig_len[a["contig"]]) - p.boundary
        if not any(inb(a) for a in alns):
            status.append(2); counts[1] += 1
            continue
        # recover_and_group_alignments: best_edit_is_good (:293-295, 1257)
        if not (best_edit[0] <= good_dist[0] and best_edit[1] <= good_dist[1]):
            status.append(3); counts[0] += 1
            continue
        for a in alns:                                                               # normalize_probs (:358-360)
            a["ln_prob"] = a["ln_prob"] - best_lik[a["end"]]
        max_alns = 10 if float(p.read_weight[r]) >= p.min_weight else 2              # :1263
        # the sort of identify_*_alignments, then pop() from the back = consumption order (:819-820, 883); Python's
        # sort is stable: with reverse=True equal keys keep their order, so the pops see them in reverse -- the
        # documented tie rule (order of `alns`) needs the reversed list as the sort input
        if p.single_end:
            key = lambda a: (a["contig"], -a["ln_prob"])
        else:
            key = lambda a: (a["contig"], a["end"], -a["ln_prob"])
        tmp = sorted(reversed(alns), key=key, reverse=True)                          # descending contig, end; ascending ln_prob
        consumed = []
        while tmp:
            consumed.append(tmp.pop())
        status.append(0); counts[2] += 1
        out_read.append(r); out_max.append(max_alns); ma.append(consumed)
    ma_off = np.cumsum([0] + [len(m) for m in ma]).astype(np.uint64)
    flat = [a for m in ma for a in m]
    return dict(status=np.array(status, dtype=np.uint8), n_reads_out=len(out_read), counts=np.array(counts, dtype=np.uint64),
                out_read=np.array(out_read, dtype=np.uint32), out_max_alns=np.array(out_max, dtype=np.uint8), ma_off=ma_off,
                ma_contig=np.array([a["contig"] for a in flat], dtype=np.uint32),
                ma_flags=np.array([a["end"] | (a["strand"] << 1) for a in flat], dtype=np.uint8),
                ma_start=np.array([a["start"] for a in flat], dtype=np.uint32),
                ma_end=np.array([a["stop"] for a in flat], dtype=np.uint32),
                ma_ln_prob=np.array([a["ln_prob"] for a in flat], dtype=np.float64),
                ma_rec=np.array([a["rec"] for a in flat], dtype=np.uint32))


def _same(a: dict, b: dict) -> bool:
    if a["n_reads_out"] != b["n_reads_out"]:
        return False
    return all(np.array_equal(np.asarray(a[k]), np.asarray(b[k])) for k in a if k != "n_reads_out")


def _hand_case() -> genotype.Prelim:
    """Two pairs on two contigs of 1,000 bp, boundary 100.  Read 0: first end two kept records (contig 1: -3.0, contig 0:
    -5.0; a third, not kept record with -2.0 sets best_lik), second end one record on contig 0.  Read 1: its only kept
    alignment has its middle at 50 < boundary -> out of bounds."""
    return genotype.Prelim(
        read_group=np.array([[0, 1], [2, 3]], dtype=np.int64), grp_off=np.array([0, 3, 4, 5, 6], dtype=np.uint64),
        rec_contig=np.array([1, 0, 0, 0, 0, 0], dtype=np.uint32), rec_start=np.array([200, 300, 700, 450, 0, 850], dtype=np.uint32),
        rec_end=np.array([300, 400, 800, 550, 100, 950], dtype=np.uint32), rec_strand=np.array([0, 0, 1, 1, 0, 1], dtype=np.uint8),
        rec_ln_prob=np.array([-3.0, -5.0, -2.0, -1.5, -4.0, -4.5]), grp_ok=np.array([1, 1, 1, 1], dtype=np.uint8),
        grp_best_edit=np.array([1, 0, 2, 2], dtype=np.uint32), grp_thr_dist=np.array([4, 4, 4, 4], dtype=np.uint32),
        grp_n_kept=np.array([2, 1, 1, 0], dtype=np.uint32),
        kept_rec=np.array([0, 1, U32_MAX, 3, 4, U32_MAX], dtype=np.uint32), contig_len=np.array([1000, 1000], dtype=np.uint32),
        read_weight=np.array([0.9, 0.1]), min_weight=0.5, boundary=100)


def test_hand_checked(oracle):
    got = oracle.group_reads(_hand_case())
    assert list(got["status"]) == [0, 2] and list(got["counts"]) == [0, 1, 1]
    assert list(got["out_read"]) == [0] and list(got["out_max_alns"]) == [10] and list(got["ma_off"]) == [0, 3]
    # contig 0 first: first end (record 1: -5 - (-2) = -3), then second end (record 3: -1.5 - (-1.5) = 0); then contig 1
    assert list(got["ma_rec"]) == [1, 3, 0] and list(got["ma_contig"]) == [0, 0, 1] and list(got["ma_flags"]) == [0, 3, 0]
    assert list(got["ma_ln_prob"]) == [-3.0, 0.0, -1.0]
    assert list(got["ma_start"]) == [300, 450, 200] and list(got["ma_end"]) == [400, 550, 300]


@pytest.mark.parametrize("seed,single_end", [(1, False), (2, False), (3, True), (4, True)])
def test_oracle_vs_transcription(oracle, seed, single_end):
    p = _random_prelim(seed, single_end=single_end)
    got, want = oracle.group_reads(p), _transcription(p)
    assert _same(got, want)
    assert {0, 1, 3} <= set(int(v) for v in got["status"])            # (status 2: the hand-checked case and seed 1)
    assert int(got["counts"].sum()) == p.n_reads


def test_empty_and_all_failing(oracle):
    p = _random_prelim(5, n_reads=40)
    p.grp_ok[:] = 0
    got = oracle.group_reads(p)
    assert got["n_reads_out"] == 0 and list(got["counts"]) == [40, 0, 0] and list(got["ma_off"]) == [0]


@pytest.mark.gpu
@pytest.mark.parametrize("seed,single_end,n_reads", [(11, False, 3000), (12, True, 3000), (13, False, 1)])
def test_gpu_group_reads_equals_oracle(gpu_ctx, oracle, seed, single_end, n_reads):
    p = _random_prelim(seed, n_reads=n_reads, single_end=single_end)
    l0 = gpu_ctx.launch_count()
    got = genotype.group_reads(gpu_ctx, p)
    assert gpu_ctx.launch_count() > l0
    assert _same(got, oracle.group_reads(p))


@pytest.mark.parametrize("tag", ["pe", "se"])
def test_oracle_equals_golden(oracle, tag):
    from conftest import load_golden_prelim
    p, want = load_golden_prelim(tag)
    assert _same(oracle.group_reads(p), want)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["pe", "se"])
def test_gpu_group_reads_equals_golden(gpu_ctx, tag):
    """No oracle at run time: the committed fixtures (tests/golden/group_*_small.npz) are the reference."""
    from conftest import load_golden_prelim
    p, want = load_golden_prelim(tag)
    assert _same(genotype.group_reads(gpu_ctx, p), want)


@pytest.mark.gpu
def test_gpu_group_reads_edge_cases(gpu_ctx, oracle):
    assert _same(genotype.group_reads(gpu_ctx, _hand_case()), oracle.group_reads(_hand_case()))
    p = _random_prelim(21, n_reads=50)
    p.grp_ok[:] = 0                                                 # nothing passes
    got = genotype.group_reads(gpu_ctx, p)
    assert got["n_reads_out"] == 0 and list(got["counts"]) == [50, 0, 0]
    empty = genotype.Prelim(read_group=np.zeros((0, 2), dtype=np.int64), grp_off=np.zeros(1, dtype=np.uint64),
                            **{k: np.zeros(0, dtype=np.uint32) for k in ("rec_contig", "rec_start", "rec_end", "grp_best_edit",
                                                                          "grp_thr_dist", "grp_n_kept", "kept_rec")},
                            rec_strand=np.zeros(0, dtype=np.uint8), rec_ln_prob=np.zeros(0), grp_ok=np.zeros(0, dtype=np.uint8),
                            contig_len=np.array([1000], dtype=np.uint32), read_weight=np.zeros(0), min_weight=0.5, boundary=10)
    assert genotype.group_reads(gpu_ctx, empty)["n_reads_out"] == 0


@pytest.mark.gpu
def test_gpu_group_then_pair_with_per_read_max_alns(gpu_ctx, oracle):
    """The output of lctp_group_reads is the input of the pairing: per-read max_alns (10 / 2) through both sides."""
    p = _random_prelim(31, n_reads=2000, n_haps=6, tie_frac=0.0)
    g = genotype.group_reads(gpu_ctx, p)
    R = g["n_reads_out"]
    assert R > 100 and set(g["out_max_alns"]) == {2, 10}
    ins = -np.abs(np.arange(8192) - 400.0) / 50.0
    mates = genotype.Mates(n_reads=R, n_haps=6, ma_off=g["ma_off"], ma_contig=g["ma_contig"], ma_flags=g["ma_flags"],
                           ma_start=g["ma_start"], ma_end=g["ma_end"], ma_ln_prob=g["ma_ln_prob"], ins_ln_pmf=ins,
                           unmapped_penalty=-20.0, insert_penalty=-12.0, prob_diff=8.0,
                           read_weight=p.read_weight[g["out_read"]], read_max_alns=g["out_max_alns"])
    got, want = genotype.pair_alignments(gpu_ctx, mates), oracle.pair_alignments(mates)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    capped = genotype.Mates(**{**mates.__dict__, "read_max_alns": None, "max_alns": 10})
    assert len(oracle.pair_alignments(capped)["pa_contig"]) > len(want["pa_contig"])      # the limit of 2 binds somewhere


def _chain(collect, group, pair, seed=77, n_reads=600):
    """records -> read ends (read_next_alns protocol) -> status / order (load + recover_and_group_alignments) -> pair
    alignments, every stage fed with the arrays of the previous one as they are."""
    from test_rescore import _read_ends
    re_ = _read_ends(2 * n_reads, seed, contigs=5, per_group=(1, 12), contig_len=3500, strict=True)
    re_.grp_read_end = (np.arange(2 * n_reads) % 2).astype(np.uint8)        # group 2r = first end, 2r + 1 = second end of read r
    c = collect(re_)
    rng = np.random.default_rng(seed + 5)
    wf = c["weight_factor"].reshape(n_reads, 2)
    pre = genotype.Prelim(read_group=np.arange(2 * n_reads, dtype=np.int64).reshape(n_reads, 2), grp_off=re_.grp_off,
                          rec_contig=re_.rec_contig, rec_start=re_.alns.aln_start, rec_end=re_.alns.aln_end,
                          rec_strand=rng.integers(0, 2, re_.alns.n_alns).astype(np.uint8), rec_ln_prob=c["ln_prob"],
                          grp_ok=c["ok"], grp_best_edit=c["best_edit"], grp_thr_dist=c["thr_dist"], grp_n_kept=c["n_kept"],
                          kept_rec=c["kept_rec"], contig_len=np.full(5, 3500, dtype=np.uint32),
                          read_weight=wf[:, 0] * wf[:, 1] * rng.uniform(0.2, 1.0, n_reads), min_weight=0.5, boundary=1400)
    g = group(pre)
    R = g["n_reads_out"]
    ins = -np.abs(np.arange(8192) - 400.0) / 50.0
    mates = genotype.Mates(n_reads=R, n_haps=5, ma_off=g["ma_off"], ma_contig=g["ma_contig"], ma_flags=g["ma_flags"],
                           ma_start=g["ma_start"], ma_end=g["ma_end"], ma_ln_prob=g["ma_ln_prob"], ins_ln_pmf=ins,
                           unmapped_penalty=-20.0, insert_penalty=-12.0, prob_diff=8.0,
                           read_weight=pre.read_weight[g["out_read"]], read_max_alns=g["out_max_alns"])
    return c, g, pair(mates)


@pytest.mark.gpu
def test_gpu_chain_records_to_pairs(gpu_ctx, oracle):
    """The three upstream calls chained on the device path against the same chain on the oracle."""
    import functools
    got = _chain(functools.partial(genotype.collect_read_ends, gpu_ctx), functools.partial(genotype.group_reads, gpu_ctx),
                 functools.partial(genotype.pair_alignments, gpu_ctx))
    want = _chain(oracle.collect_read_ends, oracle.group_reads, oracle.pair_alignments)
    assert _same(got[1], want[1])
    assert want[1]["n_reads_out"] > 50 and set(int(v) for v in want[1]["status"]) == {0, 1, 2, 3}
    for k in want[2]:
        assert np.array_equal(got[2][k], want[2][k]), k
    assert len(want[2]["pa_contig"]) > 100


@pytest.mark.gpu
def test_gpu_chain_device_resident(gpu_ctx, oracle):
    """Same chain with the pairing input left on the device (lctp_group_reads_dev -> lctp_pair_alignments_from): the
    pair alignments equal the oracle chain's, and only status / read numbers / counts came back from the grouping."""
    import functools
    captured = {}

    def group_dev(pre):
        dm = genotype.DeviceMates(gpu_ctx, pre)
        captured["dm"], captured["pre"] = dm, pre
        # the host-array variant gives the arrays the rest of _chain wants to look at (weights of the passing reads)
        return genotype.group_reads(gpu_ctx, pre)

    def pair_dev(mates):
        dm = captured["dm"]
        params = genotype.Mates(**{**mates.__dict__, "n_reads": 0, "ma_off": np.zeros(1, dtype=np.uint64),
                                   "ma_contig": np.zeros(0, dtype=np.uint32), "ma_flags": np.zeros(0, dtype=np.uint8),
                                   "ma_start": np.zeros(0, dtype=np.uint32), "ma_end": np.zeros(0, dtype=np.uint32),
                                   "ma_ln_prob": np.zeros(0), "read_max_alns": None,
                                   "read_weight": captured["pre"].read_weight[dm.out_read]})
        dp = dm.pair(params)
        out = dp.fetch()
        dp.free()
        return out

    got = _chain(functools.partial(genotype.collect_read_ends, gpu_ctx), group_dev, pair_dev)
    want = _chain(oracle.collect_read_ends, oracle.group_reads, oracle.pair_alignments)
    dm = captured["dm"]
    assert dm.n_reads_out == want[1]["n_reads_out"] and dm.n_entries == len(want[1]["ma_contig"])
    assert np.array_equal(dm.status, want[1]["status"]) and np.array_equal(dm.out_read, want[1]["out_read"])
    assert np.array_equal(dm.counts, want[1]["counts"])
    for k in want[2]:
        assert np.array_equal(got[2][k], want[2][k]), k
    dm.free()


def _check_properties(p: genotype.Prelim, g: dict):
    """Size-independent properties of the grouping output: every kept record of a passing read appears exactly once, the
    entries of a read are in consumption order, normalised ln-probabilities are <= 0, max_alns follows the weight."""
    assert int(g["counts"].sum()) == p.n_reads and int(g["counts"][2]) == g["n_reads_out"]
    assert np.all(g["ma_ln_prob"] <= 0.0)
    assert np.array_equal(g["out_max_alns"], np.where(p.read_weight[g["out_read"]] >= p.min_weight, 10, 2))
    assert np.array_equal(np.nonzero(g["status"] == 0)[0], g["out_read"])
    for k, r in enumerate(g["out_read"]):
        b, e = int(g["ma_off"][k]), int(g["ma_off"][k + 1])
        want = []
        for end in range(1 if p.single_end else 2):
            gi = int(p.read_group[r, end])
            o = int(p.grp_off[gi])
            want += [int(x) for x in p.kept_rec[o:o + int(p.grp_n_kept[gi])]]
        assert sorted(int(x) for x in g["ma_rec"][b:e]) == sorted(want)
        key = [(int(g["ma_contig"][q]), int(g["ma_flags"][q]) & 1, -float(g["ma_ln_prob"][q])) for q in range(b, e)]
        assert key == sorted(key)
        assert np.array_equal(g["ma_contig"][b:e], p.rec_contig[g["ma_rec"][b:e]])
        assert np.array_equal(g["ma_start"][b:e], p.rec_start[g["ma_rec"][b:e]])


@pytest.mark.parametrize("seed,single_end", [(41, False), (42, True)])
def test_oracle_output_properties(oracle, seed, single_end):
    p = _random_prelim(seed, n_reads=400, single_end=single_end)
    _check_properties(p, oracle.group_reads(p))


@pytest.mark.gpu
def test_gpu_output_properties_large(gpu_ctx):
    """No oracle: 20,000 reads (~170,000 records), properties only."""
    p = _random_prelim(43, n_reads=20000)
    _check_properties(p, genotype.group_reads(gpu_ctx, p))

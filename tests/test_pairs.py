"""SURVEY.md 8(f) rank 1 -- mate alignments -> pair alignments (src/model/locs.rs:744-868).
CPU: the oracle against a literal Python transcription of the reference's two functions on small inputs.
GPU: lctp_pair_alignments against the oracle, bit for bit, and the produced arrays feeding lctp_locus_upload."""
import math

import numpy as np
import pytest

from locityper_b200 import genotype, synth

NONE = 0xFFFFFFFF


def _python_reference(m: genotype.Mates) -> dict:
    """identify_paired_end_alignments / identify_contig_pair_alns, statement by statement (pure Python, small R)."""
    off, con, lp, m1, m2, unm = [0], [], [], [], [], []
    unm_ins = m.unmapped_penalty + m.insert_penalty
    for r in range(m.n_reads):
        a, end = int(m.ma_off[r]), int(m.ma_off[r + 1])
        w = 1.0 if m.read_weight is None else float(m.read_weight[r])
        while a < end:
            c = m.ma_contig[a]
            b = a
            while b < end and m.ma_contig[b] == c:
                b += 1
            first = [q for q in range(a, b) if not (m.ma_flags[q] & 1)][:m.max_alns]
            second = [q for q in range(a, b) if m.ma_flags[q] & 1][:m.max_alns]
            pairs, buf = [], [-math.inf] * len(second)
            for i1 in first:
                max1 = -math.inf
                for k, i2 in enumerate(second):
                    if (m.ma_flags[i1] ^ m.ma_flags[i2]) & 2:
                        ins = max(int(m.ma_end[i1]), int(m.ma_end[i2])) - min(int(m.ma_start[i1]), int(m.ma_start[i2]))
                        p = (float(m.ma_ln_prob[i1]) + float(m.ma_ln_prob[i2])) + float(m.ins_ln_pmf[ins])
                        if math.isfinite(p):
                            max1, buf[k] = max(max1, p), max(buf[k], p)
                            pairs.append((p, (int(m.ma_start[i1]) + int(m.ma_end[i1])) // 2,
                                          (int(m.ma_start[i2]) + int(m.ma_end[i2])) // 2))
                alone = float(m.ma_ln_prob[i1]) + unm_ins
                if alone >= max1:
                    pairs.append((alone, (int(m.ma_start[i1]) + int(m.ma_end[i1])) // 2, NONE))
            for k, i2 in enumerate(second):
                alone = float(m.ma_ln_prob[i2]) + unm_ins
                if alone >= buf[k]:
                    pairs.append((alone, NONE, (int(m.ma_start[i2]) + int(m.ma_end[i2])) // 2))
            pairs.sort(key=lambda t: -t[0])            # Python's sort is stable
            thresh = pairs[0][0] - m.prob_diff
            keep = 0
            while keep < min(len(pairs), m.max_alns) and pairs[keep][0] >= thresh:
                keep += 1
            for p, x, y in pairs[:keep]:
                con.append(c); lp.append(p * w); m1.append(x); m2.append(y)
            a = b
        off.append(len(con))
        unm.append(w * (2.0 * m.unmapped_penalty + m.insert_penalty))
    return dict(pa_off=np.array(off, dtype=np.uint64), pa_contig=np.array(con, dtype=np.uint32),
                pa_ln_prob=np.array(lp), pa_mid1=np.array(m1, dtype=np.uint32), pa_mid2=np.array(m2, dtype=np.uint32),
                unmapped_prob=np.array(unm))


def _python_single_end(m: genotype.Mates) -> dict:
    """identify_single_end_alignments (src/model/locs.rs:870-911) + explicit_read_weight, statement by statement."""
    off, con, lp, m1, m2, unm = [0], [], [], [], [], []
    for r in range(m.n_reads):
        curr, thresh, saved, start = None, math.nan, 0, len(con)
        for a in range(int(m.ma_off[r]), int(m.ma_off[r + 1])):
            if curr != m.ma_contig[a]:
                curr, thresh, saved = m.ma_contig[a], float(m.ma_ln_prob[a]) - m.prob_diff, 0
            if float(m.ma_ln_prob[a]) >= thresh and saved < m.max_alns:
                con.append(curr); lp.append(float(m.ma_ln_prob[a]))
                m1.append((int(m.ma_start[a]) + int(m.ma_end[a])) // 2); m2.append(NONE)
                saved += 1
        w = (1.0 if m.read_weight is None else float(m.read_weight[r])) * _python_explicit(m, con[start:], m1[start:], m2[start:])
        for q in range(start, len(con)):
            lp[q] *= w
        off.append(len(con))
        unm.append(w * m.unmapped_penalty)
    return dict(pa_off=np.array(off, dtype=np.uint64), pa_contig=np.array(con, dtype=np.uint32),
                pa_ln_prob=np.array(lp), pa_mid1=np.array(m1, dtype=np.uint32), pa_mid2=np.array(m2, dtype=np.uint32),
                unmapped_prob=np.array(unm))


def _python_explicit(m, con, mid1, mid2) -> float:
    """ContigInfos::explicit_read_weight (src/model/windows.rs:683-693) with read_end_weight (:495-504)."""
    if m.exp_weight is None:
        return 1.0

    def rew(c, mid):
        if mid == NONE:
            return 0.0
        w = m.exp_weight[int(m.exp_off[c]):int(m.exp_off[c + 1])]
        u = m.window // 2
        return max(float(w[mid]), float(w[max(mid - u, 0)]), float(w[min(mid + u, len(w) - 1)]))
    s = 0.0
    for c, a, b in zip(con, mid1, mid2):
        s += max(rew(c, a), rew(c, b))
    return s / len(con) if len(con) else math.nan


def _single_end(H, R, L, seed):
    """Single-end records: every record a first-end record, per (read, contig) in descending ln_prob."""
    kw = synth.make_mates(H, R, L, seed, multi_frac=0.4, over_cap_frac=0.05)
    first = (kw["ma_flags"] & 1) == 0
    cnt = np.array([first[int(kw["ma_off"][r]):int(kw["ma_off"][r + 1])].sum() for r in range(R)])
    for f in ("ma_contig", "ma_flags", "ma_start", "ma_end", "ma_ln_prob"):
        kw[f] = kw[f][first]
    kw["ma_off"] = np.r_[0, np.cumsum(cnt)].astype(np.uint64)
    kw["single_end"] = True
    return kw


def _explicit_weights(H, L, seed, window=100):
    rng = np.random.default_rng(seed)
    lens = np.full(H, L + 1)
    off = np.r_[0, np.cumsum(lens)].astype(np.uint64)
    w = np.repeat(rng.choice([1.0, 1.0, 0.5, 2.0, 0.0], size=(H * (L + 1)) // 50 + 1), 50)[:int(off[-1])]
    return dict(exp_off=off, exp_weight=np.ascontiguousarray(w, dtype=np.float64), window=window)


def _same(a: dict, b: dict):
    for k in ("pa_off", "pa_contig", "pa_mid1", "pa_mid2", "pa_ln_prob", "unmapped_prob"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("H,R,seed,multi", [(5, 40, 1, 0.15), (12, 25, 2, 0.9), (3, 7, 3, 0.0)])
def test_oracle_pairing_equals_python_transcription(oracle, H, R, seed, multi):
    m = genotype.Mates(**synth.make_mates(H, R, 4000, seed, multi_frac=multi, over_cap_frac=0.05))
    _same(oracle.pair_alignments(m), _python_reference(m))


@pytest.mark.parametrize("H,R,seed", [(5, 40, 11), (12, 25, 12)])
def test_oracle_single_end_and_explicit_weights_equal_python_transcription(oracle, H, R, seed):
    m = genotype.Mates(**_single_end(H, R, 4000, seed))
    _same(oracle.pair_alignments(m), _python_single_end(m))
    me = genotype.Mates(**_single_end(H, R, 4000, seed), **_explicit_weights(H, 4000, seed))
    _same(oracle.pair_alignments(me), _python_single_end(me))
    # paired-end reads with explicit weights: the paired transcription, rescaled per read
    kw = synth.make_mates(H, R, 4000, seed, multi_frac=0.3)
    mp = genotype.Mates(**kw, **_explicit_weights(H, 4000, seed + 1))
    got = oracle.pair_alignments(mp)
    base = _python_reference(genotype.Mates(**{**kw, "read_weight": None}))
    for r in range(R):
        b, e = int(base["pa_off"][r]), int(base["pa_off"][r + 1])
        w = float(kw["read_weight"][r]) * _python_explicit(mp, base["pa_contig"][b:e], base["pa_mid1"][b:e], base["pa_mid2"][b:e])
        assert np.array_equal(got["pa_ln_prob"][b:e], base["pa_ln_prob"][b:e] * w)
        assert got["unmapped_prob"][r] == w * (2.0 * mp.unmapped_penalty + mp.insert_penalty)
    assert np.array_equal(got["pa_off"], base["pa_off"]) and np.array_equal(got["pa_mid1"], base["pa_mid1"])


@pytest.mark.gpu
@pytest.mark.parametrize("H,R,seed", [(5, 40, 11), (12, 25, 12), (60, 500, 13)])
def test_device_single_end_and_explicit_weights_bit_exact(oracle, gpu_ctx, H, R, seed):
    for m in (genotype.Mates(**_single_end(H, R, 4000, seed)),
              genotype.Mates(**_single_end(H, R, 4000, seed), **_explicit_weights(H, 4000, seed)),
              genotype.Mates(**synth.make_mates(H, R, 4000, seed, multi_frac=0.3), **_explicit_weights(H, 4000, seed + 1))):
        _same(genotype.pair_alignments(gpu_ctx, m), oracle.pair_alignments(m))
    bad = genotype.Mates(**{**synth.make_mates(H, R, 4000, seed), "single_end": True})      # second-end records
    with pytest.raises(Exception):
        genotype.pair_alignments(gpu_ctx, bad)


@pytest.mark.gpu
def test_device_resident_pairs_feed_the_locus_without_a_host_round_trip(oracle, gpu_ctx, small_locus):
    """lctp_pair_alignments_dev -> lctp_locus_upload_pairs: the pa_* section never visits the host; matrix, prefilter
    and a solver stage equal the host-array path and the oracle."""
    import copy
    loc = copy.copy(small_locus)
    m = genotype.Mates(**synth.make_mates(loc.n_haps, loc.n_reads, 2500, 77, multi_frac=0.3))
    dp = genotype.DevicePairs(gpu_ctx, m)
    out = dp.fetch()
    _same(out, oracle.pair_alignments(m))
    loc.pa_off, loc.pa_contig, loc.pa_ln_prob = out["pa_off"], out["pa_contig"], out["pa_ln_prob"]
    loc.pa_mid1, loc.pa_mid2, loc.unmapped_prob = out["pa_mid1"], out["pa_mid2"], out["unmapped_prob"]
    ol = oracle.OracleLocus(loc)
    blank = copy.copy(loc)                   # the same locus WITHOUT its pair-alignment section
    blank.pa_off = blank.pa_contig = blank.pa_ln_prob = blank.pa_mid1 = blank.pa_mid2 = blank.unmapped_prob = None
    dl = gpu_ctx.upload(blank, pairs=dp)
    dp.free()                                # the locus owns its copy
    assert np.array_equal(dl.best_aln_matrix(), oracle.best_aln_matrix(ol))
    assert np.array_equal(dl.prefilter_scores(), oracle.prefilter_scores(ol))
    scheme = genotype.Scheme.parse(["greedy:i=50,a=2"])
    r1, r2 = genotype.init_rng(5), oracle.Rng.from_seed(5)
    got = dl.solve(scheme, 8, r1)
    ref = oracle.solve(ol, [oracle.Stage("greedy", attempts=2, in_size=50)], 8, r2, os_threads=2)
    assert np.array_equal(got.gt_ix, ref["gt_ix"]) and np.array_equal(got.lik_mean, ref["lik_mean"])
    assert got.unexpl_reads == ref["unexpl_reads"] and list(r1) == r2.state()
    dl.free()


def test_oracle_pairing_hand_checked_case(oracle):
    """One read, one contig: first mate (fwd, 100..250, -1), second mates rev at 400..550 (-2) and fwd (-0.5).
    Only the opposite-strand pair exists; the same-strand second mate survives alone; the first mate alone
    (-1 + penalty) is worse than its pair and is dropped."""
    ins = np.full(2000, -50.0); ins[450] = -3.0
    m = genotype.Mates(n_reads=1, n_haps=1, ma_off=np.array([0, 3], dtype=np.uint64),
                       ma_contig=np.zeros(3, dtype=np.uint32), ma_flags=np.array([0, 1 | 2, 1], dtype=np.uint8),
                       ma_start=np.array([100, 400, 900], dtype=np.uint32), ma_end=np.array([250, 550, 1050], dtype=np.uint32),
                       ma_ln_prob=np.array([-1.0, -2.0, -0.5]), ins_ln_pmf=ins, unmapped_penalty=-20.0,
                       insert_penalty=-3.0, prob_diff=30.0, read_weight=np.array([0.5]))
    # the caller's order is ln_prob descending within an end: swap the two second-end records
    m.ma_flags = np.array([0, 1, 1 | 2], dtype=np.uint8)
    m.ma_start = np.array([100, 900, 400], dtype=np.uint32); m.ma_end = np.array([250, 1050, 550], dtype=np.uint32)
    m.ma_ln_prob = np.array([-1.0, -0.5, -2.0])
    out = oracle.pair_alignments(m)
    assert out["pa_off"].tolist() == [0, 2]
    assert out["pa_ln_prob"].tolist() == [0.5 * ((-1.0 + -2.0) + -3.0), 0.5 * (-0.5 + (-20.0 + -3.0))]
    assert out["pa_mid1"].tolist() == [175, NONE] and out["pa_mid2"].tolist() == [475, 975]
    assert out["unmapped_prob"].tolist() == [0.5 * (2 * -20.0 + -3.0)]


@pytest.mark.gpu
@pytest.mark.parametrize("H,R,seed,multi", [(5, 40, 1, 0.15), (12, 25, 2, 0.9), (3, 7, 3, 0.0), (60, 500, 4, 0.3)])
def test_device_pairing_bit_exact(oracle, gpu_ctx, H, R, seed, multi):
    m = genotype.Mates(**synth.make_mates(H, R, 4000, seed, multi_frac=multi, over_cap_frac=0.05))
    _same(genotype.pair_alignments(gpu_ctx, m), oracle.pair_alignments(m))


@pytest.mark.gpu
def test_device_pairing_edge_cases(oracle, gpu_ctx):
    m = genotype.Mates(**synth.make_mates(4, 30, 3000, 9))
    # reads without any alignment (empty runs, including the last read) and a single-record read
    keep = np.ones(len(m.ma_contig), dtype=bool)
    for r in (0, 7, 29):
        keep[int(m.ma_off[r]):int(m.ma_off[r + 1])] = False
    keep[int(m.ma_off[3]) + 1:int(m.ma_off[4])] = False
    cnt = np.array([keep[int(m.ma_off[r]):int(m.ma_off[r + 1])].sum() for r in range(30)])
    m.ma_off = np.r_[0, np.cumsum(cnt)].astype(np.uint64)
    for f in ("ma_contig", "ma_flags", "ma_start", "ma_end", "ma_ln_prob"):
        setattr(m, f, getattr(m, f)[keep])
    m.read_weight = None
    got, ref = genotype.pair_alignments(gpu_ctx, m), oracle.pair_alignments(m)
    _same(got, ref)
    assert got["pa_off"][1] == 0 and got["pa_off"][30] == got["pa_off"][29]
    # capacity and ordering errors are reported, not silently truncated
    with pytest.raises(Exception):
        genotype.pair_alignments(gpu_ctx, m, cap=3)
    bad = genotype.Mates(**synth.make_mates(4, 10, 3000, 10))
    bad.ma_contig = bad.ma_contig[::-1].copy()
    with pytest.raises(Exception):
        genotype.pair_alignments(gpu_ctx, bad)


@pytest.mark.gpu
def test_device_pairs_feed_the_locus_upload(oracle, gpu_ctx, small_locus):
    """The arrays lctp_pair_alignments returns are a valid pa_* section of lctp_locus: replace the section of a
    synthetic locus by device-made pairs and check matrix + prefilter against the oracle on the same locus."""
    import copy
    loc = copy.copy(small_locus)
    m = genotype.Mates(**synth.make_mates(loc.n_haps, loc.n_reads, 2500, 77, multi_frac=0.3))
    out = genotype.pair_alignments(gpu_ctx, m)
    loc.pa_off, loc.pa_contig, loc.pa_ln_prob = out["pa_off"], out["pa_contig"], out["pa_ln_prob"]
    loc.pa_mid1, loc.pa_mid2, loc.unmapped_prob = out["pa_mid1"], out["pa_mid2"], out["unmapped_prob"]
    ol = oracle.OracleLocus(loc)
    dl = gpu_ctx.upload(loc)
    assert np.array_equal(dl.best_aln_matrix(), oracle.best_aln_matrix(ol))
    assert np.array_equal(dl.prefilter_scores(), oracle.prefilter_scores(ol))
    dl.free()


@pytest.mark.gpu
def test_device_vs_oracle_timing_report(oracle, gpu_ctx, capsys):
    """Not an assertion on speed: prints kernel / whole-call / single-thread oracle times at the C2 shape (run with
    -s); the oracle may only be executed from tests/, so the comparison lives here rather than in tools/pairs_run.py."""
    import time
    m = genotype.Mates(**synth.make_mates(300, 2000, 3500, 5))
    genotype.pair_alignments(gpu_ctx, m)
    gpu_ctx.stats(reset=True)
    t0 = time.perf_counter()
    got = genotype.pair_alignments(gpu_ctx, m)
    wall = time.perf_counter() - t0
    st = gpu_ctx.stats(reset=True)
    t1 = time.perf_counter()
    ref = oracle.pair_alignments(m)
    cpu = time.perf_counter() - t1
    assert all(np.array_equal(got[k], ref[k]) for k in got)
    with capsys.disabled():
        print(f"\n[pairs] {len(m.ma_contig)} mate records: kernels {st['pairing_ms']:.3f} ms, whole call {wall*1e3:.1f} ms, "
              f"oracle (1 thread) {cpu*1e3:.1f} ms")


def test_oracle_reproduces_golden_pairs(oracle):
    from conftest import load_golden_mates
    m, want = load_golden_mates()
    got = oracle.pair_alignments(m)
    assert set(got) == set(want) and all(np.array_equal(got[k], want[k]) for k in want)
    assert len(want["pa_contig"]) > 200


@pytest.mark.gpu
def test_device_reproduces_golden_pairs(gpu_ctx):
    """No oracle at run time: the committed fixture (tests/golden/pairs_small.npz) is the reference."""
    from conftest import load_golden_mates
    m, want = load_golden_mates()
    got = genotype.pair_alignments(gpu_ctx, m)
    assert all(np.array_equal(got[k], want[k]) for k in want)

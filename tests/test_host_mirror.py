"""Host-side mirror inside the C-ABI library (C++), checked against the independent C oracle on CPU.
No GPU is needed: these entry points are pure host code (scheduler pieces of src/solvers/solve.rs)."""
import ctypes as C

import numpy as np
import pytest

from locityper_b200 import ffi, genotype


def test_rng_seed_jump_long_jump_match_oracle(oracle):
    for seed in [0, 1, 12345, 2 ** 63 + 17]:
        st = genotype.init_rng(seed)
        r = oracle.Rng.from_seed(seed)
        assert list(st) == r.state()
        genotype.rng_jump(st)
        oracle.lib().lcto_rng_jump(C.byref(r))
        assert list(st) == r.state()
        genotype.rng_long_jump(st)
        oracle.lib().lcto_rng_long_jump(C.byref(r))
        assert list(st) == r.state()


@pytest.mark.parametrize("n,threads", [(1, 4), (7, 3), (100, 8), (5000, 4736), (5003, 64), (33, 33)])
def test_plan_stage_matches_oracle_shuffle_and_partition(oracle, n, threads):
    rng_g = genotype.init_rng(77)
    ixs = np.arange(n, dtype=np.uint64) * 3 + 1
    off = genotype.plan_stage(rng_g, ixs, threads)
    r = oracle.Rng.from_seed(77)
    v = (np.arange(n, dtype=np.uint64) * 3 + 1).astype(np.uint64)
    vv = v.astype(np.uintp)
    oracle.lib().lcto_rng_shuffle_usize(C.byref(r), vv.ctypes.data, n)
    assert np.array_equal(ixs, vv.astype(np.uint64))
    assert list(rng_g) == r.state()
    assert sorted(ixs.tolist()) == sorted(v.tolist())           # a permutation
    # MainWorker::run partition: ceil((n - start) / remaining workers), contiguous, all dispatched
    assert off[0] == 0 and off[-1] == n and np.all(np.diff(off.astype(np.int64)) > 0)
    start = 0
    for i in range(len(off) - 1):
        rem = threads - i
        assert off[i + 1] - off[i] == -(-(n - start) // rem)
        start = int(off[i + 1])


@pytest.mark.parametrize("seed", range(6))
def test_truncate_ixs_matches_oracle(oracle, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(5, 4000))
    scores = -np.abs(rng.normal(0, 150, n)) - 50
    if seed % 2:
        scores[rng.integers(0, n, n // 10)] = scores[0]          # exact ties
    for min_size, threads in [(1, 1), (10, 8), (n // 2, 3), (n + 5, 1), (20, n // 3 + 1)]:
        for filt in [5.0, 230.25850929940458, 1e9]:
            a = genotype.truncate_ixs(np.arange(n), scores, filt, min_size, threads)
            b = oracle.truncate_ixs(np.arange(n), scores, filt, min_size, threads)
            assert np.array_equal(a, b), (min_size, threads, filt)
            assert np.all(np.diff(scores[a.astype(np.int64)]) <= 0)                 # sorted, best first
            assert len(a) >= min(n, max(min(min_size, n), threads) if len(a) < n else 0)
    # arbitrary (non-ascending) input order falls back to the literal restatement
    perm = rng.permutation(n).astype(np.uint64)
    a = genotype.truncate_ixs(perm, scores, 40.0, 10, 4)
    b = oracle.truncate_ixs(perm, scores, 40.0, 10, 4)
    assert np.array_equal(a, b)


def test_truncate_ixs_random_shapes_ties_ranges_and_subsets(oracle):
    """The selection-first implementation (histogram k-th best, (key, position) sort) against the oracle's literal
    sort-then-cut on many random cases: all-equal scores, heavy ties, values spanning many binades, sub-lists,
    permuted lists, min_size / threads from 0 to beyond n (the branch that raises the cut to `threads`)."""
    rs = np.random.default_rng(2024)
    for it in range(400):
        n = int(rs.integers(1, 1500))
        kind = int(rs.integers(0, 5))
        if kind == 0:
            sc = -rs.gamma(2.0, 100.0, n)
        elif kind == 1:
            sc = -np.round(rs.gamma(2.0, 30.0, n))
        elif kind == 2:
            sc = np.full(n, -5.0)
        elif kind == 3:
            sc = -rs.gamma(9.0, 400.0, n)
        else:
            sc = np.where(rs.random(n) < 0.5, -np.round(rs.gamma(2.0, 3.0, n)), -1e6 * rs.random(n))
        ids = np.arange(n) if rs.random() < 0.6 else rs.permutation(n)
        if rs.random() < 0.3 and n > 3:
            ids = np.sort(rs.choice(n, size=int(rs.integers(1, n)), replace=False))
        fd = float(rs.choice([0.0, 1.0, 50.0, 230.26, 1e9]))
        ms, th = int(rs.integers(0, n + 5)), int(rs.integers(0, n + 5))
        a = genotype.truncate_ixs(ids.astype(np.uint64), sc, fd, ms, th)
        b = oracle.truncate_ixs(ids.astype(np.uint64), sc, fd, ms, th)
        assert np.array_equal(a, b), (it, n, kind, fd, ms, th)


def test_compare_and_discard_match_oracle(oracle):
    rng = np.random.default_rng(3)
    lib_o, lib_g = oracle.lib(), ffi.load()
    for _ in range(200):
        m1, m2 = -1000 + rng.normal(0, 20, 2)
        v1, v2 = rng.choice([np.nan, 0.0, 1e-300, 0.5, 4.0, 30.0], 2)
        a1, a2 = rng.choice([1, 5, 20], 2)
        x = lib_g.lctp_compare_two_likelihoods(m1, v1, int(a1), m2, v2, int(a2))
        y = lib_o.lcto_compare_two_likelihoods(m1, v1, int(a1), m2, v2, int(a2))
        assert (np.isnan(x) and np.isnan(y)) or x == pytest.approx(y, rel=1e-12, abs=1e-300)
    G = 600
    lik_mean = -2000 + rng.normal(0, 15, G)
    lik_var = np.abs(rng.normal(3, 1, G))
    attempts = np.full(G, 20, dtype=np.uint16)
    ixs = rng.permutation(G)[:400].astype(np.uint64)
    for out_size, threads in [(20, 8), (20, 1), (450, 8), (100, 200), (501, 4)]:
        a = genotype.discard_improbable(ixs, lik_mean, lik_var, attempts, -4 * np.log(10), out_size, threads)
        b = ixs.copy()
        m = lib_o.lcto_discard_improbable(b.ctypes.data, len(b), lik_mean.ctypes.data, lik_var.ctypes.data,
                                          attempts.ctypes.data, -4 * np.log(10), out_size, threads)
        assert np.array_equal(a, b[:m])


def test_depth_table_matches_oracle_and_scipy(oracle):
    from scipy.special import gammaln, logsumexp
    rng = np.random.default_rng(5)
    m = 5 + 20 * rng.random(101)
    v = m * (1.5 + rng.random(101))
    nb_n, nb_p = m * m / (v - m), m / v
    alt = np.array([0.3, 2.0, 3.0, 4.0, 5.0])
    for paired in (True, False):
        a = genotype.build_depth_table(nb_n, nb_p, paired, alt, 300)
        b = oracle.build_depth_table(nb_n, nb_p, paired, alt, 300)
        assert np.array_equal(a, b)                       # two independent restatements, same arithmetic
        # closed form with scipy (different ln_gamma algorithm -> close, not bit-equal)
        mul = 2.0 if paired else 1.0
        k = np.arange(300)[None, :]
        def nb(n, p):
            return n * np.log(p) - gammaln(n) + gammaln(n + k) - gammaln(k + 1.0) + k * np.log1p(-p)
        n1 = (nb_n * mul)[:, None]
        pp = nb_p[:, None]
        null = nb(n1, pp)
        alts = np.stack([nb(n1 * c, pp) for c in alt] + [null])
        ref = null - logsumexp(alts, axis=0)
        np.testing.assert_allclose(a, ref, rtol=1e-9, atol=1e-9)


def test_scheme_parse_mirrors_reference():
    s = genotype.Scheme.parse([])
    assert [(x.kind, x.in_size, x.attempts) for x in s.stages] == [("greedy", 5000, 1), ("anneal", 20, 20)]
    s = genotype.Scheme.parse(["greedy:i=5k,a=1,s=7,p=50,x0=rand", "simanneal:i=20,a=3,n=2k,p=1_000,prob=0.25"])
    g, a = s.stages
    assert (g.in_size, g.attempts, g.sample_size, g.plato_size, g.best_start) == (5000, 1, 7, 50, False)
    assert (a.kind, a.in_size, a.attempts, a.anneal_steps, a.plato_size, a.init_prob) == ("anneal", 20, 3, 2000, 1000, 0.25)
    assert genotype.parse_pretty_usize("2M") == 2_000_000 and genotype.parse_pretty_usize("1,000") == 1000
    for bad in ["foo", "anneal:P=0.25", "greedy:a=0", "greedy:i=0", "greedy:s=0", "anneal:P=1.5", "greedy:i", "anneal:n=0"]:
        with pytest.raises((genotype.InvalidInput, RuntimeError)):
            genotype.Scheme.parse([bad])
    c = genotype.Stage.parse(0, "anneal").to_c()
    assert (c.kind, c.attempts, c.in_size, c.plato_size, c.anneal_steps, c.init_prob) == (1, 20, 1000, 10000, 20000, 0.5)
    c = genotype.Stage.parse(0, "greedy").to_c()
    assert (c.kind, c.best_start, c.sample_size, c.plato_size) == (0, 1, 10, 100)


def test_genotype_enumeration_order(oracle, small_locus):
    import itertools
    loc = small_locus
    ref = list(itertools.combinations_with_replacement(range(loc.n_haps), loc.ploidy))
    assert len(ref) == loc.n_genotypes
    ol = oracle.OracleLocus(loc)
    out = (C.c_uint32 * 8)()
    for g in list(range(0, len(ref), 13)) + [len(ref) - 1]:
        assert loc.genotype_tuple(g) == ref[g] and loc.genotype_index(ref[g]) == g
        oracle.lib().lcto_genotype_tuple(ol.ref, g, out)
        assert tuple(out[:loc.ploidy]) == ref[g]


def test_worker_streams_equal_clone_and_jump():
    """lctp_rng_worker_streams = MainWorker::new (solve.rs:1007-1018): worker w = clone of the locus stream, then jump."""
    a = genotype.init_rng(321)
    b = a.copy()
    ws = genotype.worker_streams(a, 37)
    for w in range(37):
        assert np.array_equal(ws[w], b)
        genotype.rng_jump(b)
    assert np.array_equal(a, b)


def test_counts_to_prob_matches_transcription():
    """count_to_prob (src/model/bam.rs:54-66), statement by statement in numpy float32; every count of several attempt
    numbers.  A mapq whose unrounded value sits within 1e-4 of a half-integer is not compared (numpy's f32 log10 and
    the C library's may differ in the last bit there)."""
    import math
    for attempts in (1, 2, 5, 20, 100, 1000, 65535):
        counts = np.arange(0, attempts + 1, max(1, attempts // 997), dtype=np.uint16)
        prob, mapq = genotype.counts_to_prob(counts, attempts)
        for c, p, q in zip(counts.tolist(), prob.tolist(), mapq.tolist()):
            if c == 0:
                assert (p, q) == (0.0, 0)
            elif c == attempts:
                assert (p, q) == (1.0, 60)
            else:
                pr = np.float32(c) / np.float32(attempts)
                assert np.float32(p) == pr
                x = float(np.float32(-10.0) * np.log10(np.float32(1.0) - pr, dtype=np.float32))
                if abs(x - math.floor(x) - 0.5) > 1e-4:
                    assert q == int(min(math.floor(x + 0.5), 60.0)), (c, attempts, x, q)
    with pytest.raises(RuntimeError):
        genotype.counts_to_prob(np.array([3], dtype=np.uint16), 2)

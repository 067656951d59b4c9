"""Pins of the CPU oracle (oracle/, the parity checker) -- all CPU.

The reference ships no tests or vectors for this path ("parity unpinned"), so the
oracle is pinned by (1) the published known-answer vectors of the third-party algorithms it restates
(xoshiro256++ reference C of Blackman & Vigna, also rand_xoshiro's own unit test; SplitMix64), (2) an
independent implementation of the special functions (scipy), (3) structural properties of the RNG
(jumps are linear maps that commute with stepping), (4) brute-force recomputation / exhaustive search
of the objective on tiny instances, (5) the committed golden fixtures of tests/golden/ (regression pin).
"""
import ctypes as C
import itertools
import json
import os

import numpy as np
import pytest

from locityper_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["illumina_p2", "hifi_p2", "illumina_p3"]


def _fromhex(v):
    return np.array([float.fromhex(x) for x in v], dtype=np.float64)


# ---------------------------------------------------------------- third-party known answers -----

def test_xoshiro256pp_known_answer_vector(oracle):
    """State [1,2,3,4]: the vector of the public reference implementation (and of rand_xoshiro's test)."""
    rng = oracle.Rng.from_state([1, 2, 3, 4])
    want = [41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205, 9973669472204895162,
            14011001112246962877, 12406186145184390807, 15849039046786891736, 10450023813501588000]
    got = [int(oracle.lib().lcto_rng_next_u64(C.byref(rng))) for _ in want]
    assert got == want


def test_splitmix64_seeding_known_answer(oracle):
    """seed_from_u64 = 4 SplitMix64 outputs; seed 1234567 is the published SplitMix64 test vector."""
    rng = oracle.Rng.from_seed(1234567)
    assert rng.state() == [6457827717110365317, 3203168211198807973, 9817491932198370423, 4593380528125082431]


def test_jumps_commute_with_stepping(oracle):
    """jump / long_jump are powers of the (linear) state transition, so they commute with next_u64 and with
    each other; a wrong jump polynomial breaks this with probability 1 - 2^-256."""
    L = oracle.lib()
    for fn in (L.lcto_rng_jump, L.lcto_rng_long_jump):
        a, b = oracle.Rng.from_seed(42), oracle.Rng.from_seed(42)
        fn(C.byref(a)); L.lcto_rng_next_u64(C.byref(a))
        L.lcto_rng_next_u64(C.byref(b)); fn(C.byref(b))
        assert a.state() == b.state()
    a, b = oracle.Rng.from_seed(7), oracle.Rng.from_seed(7)
    L.lcto_rng_jump(C.byref(a)); L.lcto_rng_long_jump(C.byref(a))
    L.lcto_rng_long_jump(C.byref(b)); L.lcto_rng_jump(C.byref(b))
    assert a.state() == b.state()
    assert a.state() != oracle.Rng.from_seed(7).state()


def test_bounded_draws_are_in_range_and_uniform(oracle):
    L = oracle.lib()
    rng = oracle.Rng.from_seed(1)
    for lo, hi in [(0, 0), (0, 6), (3, 9), (0, 2 ** 32 - 1)]:
        xs = [L.lcto_rng_range_u32_incl(C.byref(rng), lo, hi) for _ in range(2000)]
        assert min(xs) >= lo and max(xs) <= hi
    xs = np.array([L.lcto_rng_range_u32_incl(C.byref(rng), 0, 9) for _ in range(20000)])
    cnt = np.bincount(xs, minlength=10)
    assert cnt.min() > 1700 and cnt.max() < 2300
    out = (C.c_uint32 * 10)()
    for _ in range(200):                           # Floyd sample: distinct, in range
        assert L.lcto_rng_sample_indices(C.byref(rng), 37, 10, out) == 0
        v = list(out)
        assert len(set(v)) == 10 and max(v) < 37
    f = [L.lcto_rng_f64(C.byref(rng)) for _ in range(5000)]
    assert 0.0 <= min(f) and max(f) < 1.0 and abs(np.mean(f) - 0.5) < 0.02


def test_shuffle_is_a_permutation_and_single_value_range_consumes_a_draw(oracle):
    L = oracle.lib()
    rng = oracle.Rng.from_seed(9)
    v = np.arange(1000, dtype=np.uint64)
    L.lcto_rng_shuffle_usize(C.byref(rng), v.ctypes.data, len(v))
    assert sorted(v.tolist()) == list(range(1000)) and v.tolist() != list(range(1000))
    a, b = oracle.Rng.from_seed(3), oracle.Rng.from_seed(3)
    L.lcto_rng_range_i32_incl(C.byref(a), 0, 0)     # `0..=0` still draws
    L.lcto_rng_next_u64(C.byref(b))
    assert a.state() == b.state()


# ---------------------------------------------------------------- special functions vs scipy ----

def test_special_functions_against_scipy(oracle):
    from scipy import special, stats
    L = oracle.lib()
    for x in [1e-3, 0.3, 0.5, 1.0, 1.5, 2.0, 7.25, 30.0, 171.5, 1e4, 3.3e6]:
        assert L.lcto_ln_gamma(x) == pytest.approx(special.gammaln(x), rel=1e-13, abs=1e-13)
    for x, df in [(-3.0, 2.5), (-0.4, 19.0), (0.0, 4.0), (1.7, 38.0), (6.0, 7.7)]:
        assert L.lcto_students_t_cdf(x, df) == pytest.approx(stats.t.cdf(x, df), rel=1e-10)
    for a, b, x in [(0.5, 0.5, 0.3), (2.0, 5.0, 0.1), (9.5, 0.5, 0.97)]:
        assert L.lcto_beta_reg(a, b, x) == pytest.approx(special.betainc(a, b, x), rel=1e-10)
    assert L.lcto_ln_add(-1.0, -2.0) == pytest.approx(np.logaddexp(-1.0, -2.0), rel=1e-15)
    assert L.lcto_ln_add(-np.inf, -2.0) == -2.0


# ---------------------------------------------------------------- objective: brute force --------

def _instance_arrays(inst):
    I = inst.contents
    A, R, W = I.n_alns, I.n_reads, I.total_windows
    return dict(A=A, R=R, W=W,
                read_ixs=np.ctypeslib.as_array(I.read_ixs, (R + 1,)).copy(),
                ln_prob=np.ctypeslib.as_array(I.aln_ln_prob, (A,)).copy(),
                w=np.ctypeslib.as_array(I.aln_w, (A * 2,)).reshape(A, 2).copy(),
                weight=np.ctypeslib.as_array(I.win_weight, (W,)).copy(),
                gc=np.ctypeslib.as_array(I.win_gc, (W,)).copy(),
                trivial=np.ctypeslib.as_array(I.win_trivial, (W,)).copy())


def _objective(loc, arr, assgn):
    """likelihood() (src/model/assgn.rs:235-237) recomputed from scratch."""
    table = loc.depth_table.reshape(101, loc.depth_k)
    depth = np.zeros(arr["W"], dtype=np.int64)
    aln = 0.0
    for r in range(arr["R"]):
        ix = arr["read_ixs"][r] + assgn[r]
        aln += arr["ln_prob"][ix]
        depth[arr["w"][ix, 0]] += 1
        depth[arr["w"][ix, 1]] += 1
    dl = sum(arr["weight"][w] * table[arr["gc"][w], depth[w]] for w in range(arr["W"]) if not arr["trivial"][w])
    return (1 + loc.lik_skew) * dl + (1 - loc.lik_skew) * aln, depth


@pytest.mark.parametrize("kind", ["greedy", "anneal"])
def test_attempt_likelihood_equals_recomputed_objective_and_is_below_optimum(oracle, kind):
    L = oracle.lib()
    loc = synth.make_locus(8, 10, 2000, seed=555, table_builder=oracle.build_depth_table)
    ol = oracle.OracleLocus(loc)
    st = oracle.Stage(kind, attempts=1, anneal_steps=800, plato_size=300 if kind == "anneal" else None).to_c()
    checked_opt = reached = 0
    for g in range(0, loc.n_genotypes, 4):
        inst = L.lcto_instance_new(ol.ref, g)
        rng = oracle.Rng.from_seed(100 + g)
        L.lcto_apply_tweak(ol.ref, inst, C.byref(rng))
        arr = _instance_arrays(inst)
        assgn = np.zeros(arr["R"], dtype=np.uint16)
        depth = np.zeros(arr["W"], dtype=np.uint32)
        out = oracle.AttemptOut()
        assert L.lcto_solve_attempt(ol.ref, inst, C.byref(st), C.byref(rng), assgn.ctypes.data, depth.ctypes.data,
                                    C.byref(out)) == 0
        lik, dep = _objective(loc, arr, assgn)
        assert np.array_equal(dep, depth)                                  # integer state exact
        assert out.lik == pytest.approx(lik, rel=1e-9, abs=1e-9)           # incremental += vs from scratch
        ncand = np.diff(arr["read_ixs"])
        nt = np.nonzero(ncand > 1)[0]
        if 0 < len(nt) and np.prod(ncand[nt].astype(float)) <= 4096:       # exhaustive optimum (ILP objective)
            best = -np.inf
            for combo in itertools.product(*[range(ncand[r]) for r in nt]):
                a = np.zeros(arr["R"], dtype=np.int64)
                a[nt] = combo
                best = max(best, _objective(loc, arr, a)[0])
            assert out.lik <= best + 1e-9 * abs(best)
            checked_opt += 1
            reached += int(out.lik >= best - 1e-9 * abs(best))
        L.lcto_instance_free(inst)
    assert checked_opt > 0
    assert reached * 2 >= checked_opt, (reached, checked_opt)   # easy instances: the optimum is usually found


def test_prefilter_scores_equal_numpy_restated(oracle):
    loc = synth.make_locus(15, 120, 2000, seed=77, table_builder=oracle.build_depth_table)
    ol = oracle.OracleLocus(loc)
    M = oracle.best_aln_matrix(ol)
    s = oracle.prefilter_scores(ol, M=M)
    for g in range(loc.n_genotypes):
        t = loc.genotype_tuple(g)
        row = M[t[0]].copy()
        for h in t[1:]:
            row = np.maximum(row, M[h])
        acc = 0.0
        for v in row:                      # sequential f64 sum in read order (solve.rs:114)
            acc += v
        assert s[g] == acc


# ---------------------------------------------------------------- golden fixtures ---------------

@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden_fixture(oracle, name):
    with open(os.path.join(GOLD, f"golden_{name}.json")) as f:
        gold = json.load(f)
    loc = synth.load_locus(os.path.join(GOLD, f"locus_{name}.npz"))
    ol = oracle.OracleLocus(loc)
    G = loc.n_genotypes
    assert G == gold["n_genotypes"]
    scores = oracle.prefilter_scores(ol)
    assert np.array_equal(scores, _fromhex(gold["prefilter_scores"]))
    for t in gold["truncate"]:
        surv = oracle.truncate_ixs(np.arange(G), scores, loc.filt_diff, t["min_size"], t["threads"])
        assert surv.tolist() == t["survivors"]
    for sc in gold["stages"]:
        rng = np.array(sc["rng_in"], dtype=np.uint64)
        r = oracle.solve_stage(ol, oracle.Stage(**sc["kw"]), np.array(sc["ixs"], dtype=np.uint64),
                               np.array(sc["off"], dtype=np.uint64), rng, os_threads=2, want_counts=True,
                               counts_cap=sc["counts_off"][-1] + 1)
        assert rng.tolist() == sc["rng_out"]
        assert np.array_equal(r["liks"].reshape(-1), _fromhex(sc["liks"]))
        assert r["iters"].tolist() == sc["iters"] and r["n_alns"].tolist() == sc["n_alns"]
        assert r["counts"][:64].tolist()[:len(sc["counts_head"])] == sc["counts_head"]
    for sv in gold["solve"]:
        scheme = [oracle.Stage(**{k: v for k, v in s.items()}) for s in sv["scheme"]]
        rng = oracle.Rng.from_seed(sv["seed"])
        r = oracle.solve(ol, scheme, sv["threads"], rng, os_threads=2)
        assert r["gt_ix"].tolist() == sv["gt_ix"]
        assert np.array_equal(r["lik_mean"], _fromhex(sv["lik_mean"]))
        assert rng.state() == sv["rng_out"]

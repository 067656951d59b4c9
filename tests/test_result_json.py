"""Genotyping::to_json text and find_weighted_dist (src/solvers/solve.rs:339-357,616-632,732-773): the product's
C++ writer (lctp_result_json / lctp_find_weighted_dist, host code, no GPU needed) against the oracle-side
restatements (oracle/lcto_py.to_json_text -- pure Python -- and lcto_find_weighted_dist -- plain C)."""
import ctypes as C
import json
import math

import numpy as np
import pytest

from locityper_b200 import ffi, genotype, synth


def _result(rng, loc, n_out, with_nan_var=False, tiny=False):
    res = ffi.ResultC()
    res.n_out = n_out
    gts = rng.choice(loc.n_genotypes, size=n_out, replace=False)
    lp = np.sort(rng.normal(-5, 4, n_out))[::-1]
    if tiny:
        lp = lp * 20
    lp = lp - np.log(np.exp(lp).sum())
    for i in range(n_out):
        res.gt_ix[i] = int(gts[i])
        res.lik_mean[i] = float(-rng.gamma(9, 900))
        res.lik_var[i] = float("nan") if with_nan_var else float(rng.gamma(2, 30))
        res.attempts[i] = 20
        res.ln_prob[i] = float(lp[i])
    res.quality = float(min(-10 * math.log10(max(1 - math.exp(lp[0]), 1e-300)), 1e9))
    res.total_reads = loc.n_reads
    res.unexpl_reads = int(rng.integers(0, loc.n_reads))
    return res


def _as_dict(res):
    n = int(res.n_out)
    return dict(gt_ix=list(res.gt_ix[:n]), lik_mean=list(res.lik_mean[:n]), lik_var=list(res.lik_var[:n]),
                ln_prob=np.array(res.ln_prob[:n]), quality=res.quality, total_reads=res.total_reads,
                unexpl_reads=res.unexpl_reads, warn_no_probable=bool(res.warn_no_probable),
                warn_few_reads=bool(res.warn_few_reads),
                distances=([None if d == 0xFFFFFFFF else int(d) for d in res.dist_to_primary[:n]] if res.has_dist else None),
                true_edit_distances=bool(res.true_edit_distances),
                weight_dist=(res.weight_dist if res.has_dist and res.has_weight_dist else None))


def _json_text(lib, res, locc, names):
    cn = (C.c_char_p * len(names))(*[s.encode() for s in names])
    need = lib.lctp_result_json(C.byref(res), C.byref(locc), cn, None, 0)
    buf = C.create_string_buffer(need + 1)
    lib.lctp_result_json(C.byref(res), C.byref(locc), cn, buf, need + 1)
    return buf.value.decode()


@pytest.mark.parametrize("ploidy", [1, 2, 3, 4])
def test_json_text_and_weighted_dist_match_the_oracle(oracle, ploidy):
    lib = ffi.load()
    rng = np.random.default_rng(40 + ploidy)
    H = 9
    loc = synth.make_locus(H, 40, 2500, seed=5, ploidy=ploidy, table_builder=oracle.build_depth_table)
    keep = []
    locc = genotype.locus_to_c(loc, keep)
    ol = oracle.OracleLocus(loc)
    names = [f"HG{100 + i}.{i % 2 + 1}" for i in range(H)]
    for case in range(40):
        n_out = int(rng.integers(1, min(12, loc.n_genotypes) + 1))
        res = _result(rng, loc, n_out, with_nan_var=(case % 3 == 0), tiny=(case % 4 == 1))
        if case % 5 == 0:
            res.warn_no_probable = 1
        if case % 7 == 0:
            res.warn_few_reads = 1
        mode = case % 3           # 0: no distances, 1: all known, 2: some unknown
        dist = None
        if mode:
            dist = rng.integers(0, 500, H * (H - 1) // 2).astype(np.uint32)
            if mode == 2:
                dist[rng.random(len(dist)) < 0.3] = 0xFFFFFFFF
            ores = oracle.ResultC()
            C.memmove(C.byref(ores), C.byref(res), C.sizeof(ores))
            assert C.sizeof(ores) == C.sizeof(res)
            ffi.check(lib.lctp_find_weighted_dist(C.byref(res), C.byref(locc), dist.ctypes.data, case % 2))
            oracle.lib().lcto_find_weighted_dist(ol.ref, C.byref(ores), dist.ctypes.data, case % 2)
            n = n_out
            assert list(res.dist_to_primary[:n]) == list(ores.dist_to_primary[:n])
            assert res.has_weight_dist == ores.has_weight_dist and res.has_dist == 1
            if res.has_weight_dist:
                assert res.weight_dist == ores.weight_dist            # same sums in the same order: bit-identical
            assert res.dist_to_primary[0] == 0
        text = _json_text(lib, res, locc, names)
        assert text == oracle.to_json_text(_as_dict(res), loc, names)
        js = json.loads(text)                                          # and it is valid JSON with the reference's keys
        assert list(js.keys())[:2] == ["total_reads", "quality"]
        assert js["genotype"] == js["options"][0]["genotype"]
        assert ("dist_type" in js) == (mode != 0)
        if mode:
            assert js["dist_type"] == ("edit" if case % 2 else "minim-div")
            assert all("dist_to_primary" in o for o in js["options"])


def test_genotype_distance_skips_the_initial_arrangement_for_three_contigs(oracle):
    """gen_permutations (src/ext/vec.rs:342-372) calls `action` only after a swap when n >= 3, so the identity
    arrangement is never scored: two genotypes that differ in ONE contig are not at that contig's distance."""
    lib = ffi.load()
    H = 5
    loc = synth.make_locus(H, 30, 2500, seed=6, ploidy=3, table_builder=oracle.build_depth_table)
    keep = []
    locc = genotype.locus_to_c(loc, keep)
    res = ffi.ResultC()
    res.n_out = 2
    res.gt_ix[0] = loc.genotype_index((0, 1, 2))
    res.gt_ix[1] = loc.genotype_index((0, 1, 3))
    res.ln_prob[0] = math.log(0.75)
    res.ln_prob[1] = math.log(0.25)
    dist = np.full(H * (H - 1) // 2, 100, dtype=np.uint32)

    def tri(i, j):
        i, j = min(i, j), max(i, j)
        return (2 * H - 3 - i) * i // 2 + j - 1

    dist[tri(2, 3)] = 1
    ffi.check(lib.lctp_find_weighted_dist(C.byref(res), C.byref(locc), dist.ctypes.data, 1))
    # identity would give d(2,3) = 1; every visited permutation moves at least two contigs
    assert res.dist_to_primary[1] > 1
    # brute force over the permutations Heap's algorithm emits after its first swap (all but the identity)
    import itertools
    g0, g1 = (0, 1, 2), (0, 1, 3)
    best = min(sum(int(dist[tri(a, b)]) for a, b in zip(pm, g1) if a != b)
               for pm in itertools.permutations(g0) if pm != g0)
    assert res.dist_to_primary[1] == best
    assert res.weight_dist == pytest.approx(0.25 * best)


def test_json_numbers_follow_the_json_crate_rules(oracle):
    # values from the `json` crate's own stringify tests, restated
    for v, s in [(3.141592653589793, "3.141592653589793"), (0.0001, "0.0001"), (1e19, "10000000000000000000"),
                 (3.141592653589793e50, "3.141592653589793e50"), (3.141592653589793e-50, "3.141592653589793e-50"),
                 (1.2345, "1.2345"), (100.0, "100"), (float("nan"), "null"), (float("inf"), "null"), (-0.5, "-0.5")]:
        assert oracle.json_number(v) == s
    # the C++ writer prints the same text for the same doubles (through the quality field)
    lib = ffi.load()
    loc = synth.make_locus(6, 30, 2500, seed=7, table_builder=oracle.build_depth_table)
    keep = []
    locc = genotype.locus_to_c(loc, keep)
    rng = np.random.default_rng(9)
    vals = list(10.0 ** rng.uniform(-30, 25, 300) * rng.choice([-1, 1], 300)) + [0.0, -0.0, 1e9, 5e-324, 1.7976931348623157e308,
            0.1, 1 / 3, 2.5e-18, 1e-17, 123456789012345678.0, 1e20, 1e21, 99999999999999990000.0]
    res = ffi.ResultC()
    res.n_out = 0
    for v in vals:
        res.quality = float(v)
        text = _json_text(lib, res, locc, [f"h{i}" for i in range(6)])
        got = text.split('"quality": ')[1].split(",\n")[0]
        assert got == oracle.json_number(float(v)), (v, got)

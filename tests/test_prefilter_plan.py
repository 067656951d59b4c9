"""Host-side check of the balanced prefilter planner (csrc/prefilter.cu, `lctp_prefilter_plan_check`): for every
panel size / genotype range / per-sub-partition pattern, every genotype id is owned by exactly one register of one
lane, and the staged matrix columns of that register are the genotype's two haplotypes.  No device needed."""
import ctypes as C

import pytest

from locityper_b200 import ffi


def _check(H, n_sm=148, pattern=None, g_begin=0, g_end=None):
    lib = ffi.load()
    G = H * (H + 1) // 2
    g_end = G if g_end is None else g_end
    n_regions, load = C.c_uint32(0), C.c_uint32(0)
    pat_out = (C.c_uint32 * 4)()
    if pattern:
        arr = (C.c_uint32 * len(pattern))(*pattern)
        rc = lib.lctp_prefilter_plan_check(H, n_sm, arr, len(pattern), g_begin, g_end, C.byref(n_regions),
                                           C.byref(load), pat_out)
    else:
        rc = lib.lctp_prefilter_plan_check(H, n_sm, None, 0, g_begin, g_end, C.byref(n_regions), C.byref(load), pat_out)
    assert rc == 0, lib.lctp_last_error().decode()
    return n_regions.value, load.value, [p for p in pat_out if p]


@pytest.mark.parametrize("H", [1, 2, 3, 31, 32, 33, 64, 100, 257, 530, 1000, 1023, 1500])
def test_auto_plan_covers_every_genotype_once(H):
    n_regions, load, pat = _check(H)
    assert n_regions >= 1 and load == sum(pat) and all(2 <= p <= 4 for p in pat)


@pytest.mark.parametrize("pattern", [[4], [3], [2], [4, 3], [4, 4], [3, 2], [2, 2, 2], [4, 3, 2], [7], [8], [5, 6], [4, 4, 4, 4]])
def test_explicit_patterns(pattern):
    for H in (97, 300, 777):
        n_regions, load, pat = _check(H, pattern=pattern)
        assert pat == pattern and load == sum(pattern)


def test_kir_scale_plan_fits_one_round_on_a_b200():
    """H = 1,000 on 148 SMs: 8 warps with 4 and 3 columns per lane, <= 148 regions (one persistent CTA each), a load
    of 7 columns x 4 rows per sub-partition lane against the ideal 500,500 / (148 * 4 * 32 * 4) = 6.6."""
    n_regions, load, pat = _check(1000)
    assert n_regions <= 148 and load == 7 and pat == [4, 3]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_shard_ranges(world):
    from locityper_b200 import dist
    H = 411
    G = H * (H + 1) // 2
    for r in range(world):
        a, b = dist.shard_range(G, r, world)
        _check(H, g_begin=a, g_end=b)


def test_bad_arguments_are_rejected():
    lib = ffi.load()
    assert lib.lctp_prefilter_plan_check(0, 148, None, 0, 0, 1, None, None, None) != 0
    assert lib.lctp_prefilter_plan_check(10, 148, None, 0, 5, 5, None, None, None) != 0
    assert lib.lctp_prefilter_plan_check(10, 148, None, 0, 0, 56, None, None, None) != 0
    bad = (C.c_uint32 * 1)(9)
    assert lib.lctp_prefilter_plan_check(10, 148, bad, 1, 0, 55, None, None, None) != 0


def test_random_panels_ranges_and_patterns_property():
    """Property (hypothesis): for any panel size, SM count, genotype sub-range and per-sub-partition pattern the plan
    covers every genotype of the range exactly once with consistent staging tables (the check is exhaustive inside
    lctp_prefilter_plan_check; hypothesis varies its inputs)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(H=st.integers(1, 700), n_sm=st.sampled_from([1, 7, 132, 148]),
           pattern=st.one_of(st.none(), st.lists(st.integers(2, 8), min_size=1, max_size=4)),
           cut=st.tuples(st.floats(0, 1), st.floats(0, 1)))
    def prop(H, n_sm, pattern, cut):
        from hypothesis import assume
        assume(not pattern or sum(pattern) <= 20)      # wider regions can exceed the 256 staged chunks per read (no plan)
        G = H * (H + 1) // 2
        a, b = sorted(int(c * G) for c in cut)
        b = max(b, a + 1)
        if b > G:
            a, b = G - 1, G
        n_regions, load, pat = _check(H, n_sm=n_sm, pattern=pattern, g_begin=a, g_end=b)
        assert n_regions >= 1 and load == sum(pat)
        if pattern:
            assert pat == pattern

    prop()

"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (north_star): identical genotype ranking / survivor set / final calls, identical integer assignment
counts, ln-likelihoods within 1e-6 relative (RTOL below; in practice they are bit-identical because the
kernels keep the reference's summation order and never contract mul+add).
"""
import numpy as np
import pytest

from locityper_b200 import genotype, synth

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _mk(oracle, H, R, L, seed, **kw):
    return synth.make_locus(H, R, L, seed=seed, table_builder=oracle.build_depth_table, **kw)


def _workers(oracle, n_workers, seed):
    rng = oracle.Rng.from_seed(seed)
    import ctypes as C
    st = np.zeros((n_workers, 4), dtype=np.uint64)
    for w in range(n_workers):
        st[w] = rng.state()
        oracle.lib().lcto_rng_jump(C.byref(rng))
    return st


@pytest.mark.parametrize("shape", [(24, 300, 2500, "illumina", 2), (17, 150, 2500, "illumina", 1),
                                   (9, 120, 2500, "illumina", 3), (30, 60, 30000, "hifi", 2)])
def test_matrix_and_prefilter_bit_exact(oracle, gpu_ctx, shape):
    H, R, L, tech, p = shape
    loc = _mk(oracle, H, R, L, 100 + H, tech=tech, ploidy=p)
    ol = oracle.OracleLocus(loc)
    dl = gpu_ctx.upload(loc)
    M_ref = oracle.best_aln_matrix(ol)
    M = dl.best_aln_matrix()
    assert np.array_equal(M, M_ref)
    s_ref = oracle.prefilter_scores(ol, M=M_ref)
    s = dl.prefilter_scores()
    assert np.array_equal(s, s_ref)          # bit-exact: same read-order summation
    for min_size, threads in [(10, 1), (50, 8), (10 ** 6, 8)]:
        surv_ref = oracle.truncate_ixs(np.arange(loc.n_genotypes), s_ref, loc.filt_diff, min_size, threads)
        surv = dl.prefilter(min_size, threads)
        assert np.array_equal(surv, surv_ref)
    dl.free()


def test_prefilter_explicit_priors(oracle, gpu_ctx):
    loc = _mk(oracle, 20, 200, 2500, 5, explicit_priors=True)
    ol = oracle.OracleLocus(loc)
    dl = gpu_ctx.upload(loc)
    assert np.array_equal(dl.prefilter_scores(), oracle.prefilter_scores(ol))
    dl.free()


@pytest.mark.parametrize("H", [100, 530])
def test_prefilter_tile_variants(oracle, gpu_ctx, H):
    """H=100 -> 16x16 tiles, H=530 -> 32x32 tiles (2x2 register tile), ragged edges, R not a multiple of 32."""
    loc = _mk(oracle, H, 333, 2500, 7 + H)
    ol = oracle.OracleLocus(loc)
    dl = gpu_ctx.upload(loc)
    assert np.array_equal(dl.prefilter_scores(), oracle.prefilter_scores(ol))
    # shard ranges (multi-GPU partition) reproduce the same values
    G = loc.n_genotypes
    cut = G // 3
    a = dl.prefilter_scores(0, cut)
    b = dl.prefilter_scores(cut, G)
    assert np.array_equal(np.concatenate([a, b]), oracle.prefilter_scores(ol))
    dl.free()


def _stage_parity(oracle, gpu_ctx, loc, stage_kw, n_workers, gts, seed=3, want_counts=True):
    ol = oracle.OracleLocus(loc)
    dl = gpu_ctx.upload(loc)
    ixs = np.array(gts, dtype=np.uint64)
    off = np.linspace(0, len(ixs), n_workers + 1).astype(np.uint64)
    rng_ref = _workers(oracle, n_workers, seed)
    rng_gpu = rng_ref.copy()
    cap = int(len(ixs) * (loc.n_reads * (10 * loc.ploidy + 1)))
    ost = oracle.Stage(**stage_kw)
    gst = genotype.Stage(**stage_kw)
    ref = oracle.solve_stage(ol, ost, ixs, off, rng_ref, os_threads=4, want_counts=want_counts, counts_cap=cap)
    got = dl.solve_stage(gst, ixs, off, rng_gpu, want_counts=want_counts, counts_cap=cap)
    dl.free()
    assert np.array_equal(got["n_alns"], ref["n_alns"])
    assert np.array_equal(rng_gpu, rng_ref), "worker RNG streams diverged"
    assert np.array_equal(got["iters"], ref["iters"])
    np.testing.assert_allclose(got["liks"], ref["liks"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(got["lik_mean"], ref["lik_mean"], rtol=RTOL, atol=0)
    if stage_kw.get("attempts", 20) > 1:
        np.testing.assert_allclose(got["lik_var"], ref["lik_var"], rtol=1e-5, atol=1e-12)
    if want_counts:
        assert np.array_equal(got["counts_off"], ref["counts_off"])
        n = int(ref["counts_off"][-1])
        assert np.array_equal(got["counts"][:n], ref["counts"][:n])
    return got, ref


def test_greedy_stage_parity(oracle, gpu_ctx, small_locus):
    gts = list(range(0, small_locus.n_genotypes, 7))
    got, ref = _stage_parity(oracle, gpu_ctx, small_locus, dict(kind="greedy", attempts=2), 5, gts)
    assert np.array_equal(got["liks"], ref["liks"]), "expected bit-identical likelihoods"


def test_greedy_random_start_and_small_sample(oracle, gpu_ctx, small_locus):
    gts = list(range(3, small_locus.n_genotypes, 29))
    _stage_parity(oracle, gpu_ctx, small_locus, dict(kind="greedy", attempts=3, best_start=False, sample_size=4,
                                                     plato_size=30), 3, gts)


def test_anneal_stage_parity(oracle, gpu_ctx, small_locus):
    gts = list(range(1, small_locus.n_genotypes, 61))
    _stage_parity(oracle, gpu_ctx, small_locus, dict(kind="anneal", attempts=3, anneal_steps=3000, plato_size=1500),
                  2, gts)


def test_stage_parity_ploidy3_and_hifi(oracle, gpu_ctx):
    loc3 = _mk(oracle, 8, 150, 2500, 21, ploidy=3)
    _stage_parity(oracle, gpu_ctx, loc3, dict(kind="greedy", attempts=2), 4, list(range(0, loc3.n_genotypes, 5)))
    hifi = _mk(oracle, 20, 80, 40000, 22, tech="hifi")
    _stage_parity(oracle, gpu_ctx, hifi, dict(kind="anneal", attempts=4, anneal_steps=2000, plato_size=800), 3,
                  list(range(0, hifi.n_genotypes, 11)))


def test_tweak_zero(oracle, gpu_ctx):
    loc = _mk(oracle, 12, 200, 2500, 31)
    loc.tweak = 0
    _stage_parity(oracle, gpu_ctx, loc, dict(kind="greedy", attempts=2), 2, list(range(0, loc.n_genotypes, 9)))


@pytest.mark.parametrize("threads", [1, 8, 64])
def test_full_solve_identical_calls(oracle, gpu_ctx, small_locus, threads):
    loc = small_locus
    scheme_o = [oracle.Stage("greedy", attempts=1, in_size=100), oracle.Stage("anneal", attempts=5, in_size=10,
                                                                            anneal_steps=2000, plato_size=1000)]
    scheme_g = genotype.Scheme([genotype.Stage("greedy", attempts=1, in_size=100),
                                genotype.Stage("anneal", attempts=5, in_size=10, anneal_steps=2000, plato_size=1000)])
    ol = oracle.OracleLocus(loc)
    rng_o = oracle.Rng.from_seed(99)
    ref = oracle.solve(ol, scheme_o, threads, rng_o, os_threads=4)
    dl = gpu_ctx.upload(loc)
    rng_g = genotype.init_rng(99)
    got = dl.solve(genotype.Scheme(scheme_g.stages), threads, rng_g)
    dl.free()
    assert got.n_filtered == ref["n_filtered"] and got.n_stage_in == ref["n_stage_in"]
    assert np.array_equal(got.gt_ix, ref["gt_ix"]), "ranking / final call differs"
    np.testing.assert_allclose(got.lik_mean, ref["lik_mean"], rtol=RTOL)
    np.testing.assert_allclose(got.ln_prob, ref["ln_prob"], rtol=1e-6, atol=1e-9)
    assert got.unexpl_reads == ref["unexpl_reads"]
    assert list(rng_g) == rng_o.state(), "locus RNG stream diverged"
    assert abs(got.quality - ref["quality"]) <= 1e-6 * max(1.0, abs(ref["quality"]))
    js = got.to_json()
    assert js["total_reads"] == loc.n_reads and len(js["options"]) == len(ref["gt_ix"])


def test_sharded_solve_single_rank_equals_lctp_solve(oracle, gpu_ctx, small_locus):
    """dist.solve_sharded composes the same C-ABI pieces as lctp_solve: identical calls and RNG stream."""
    from locityper_b200 import dist as ldist
    loc = small_locus
    scheme = genotype.Scheme([genotype.Stage("greedy", attempts=1, in_size=100),
                              genotype.Stage("anneal", attempts=4, in_size=10, anneal_steps=1500, plato_size=700)])
    dl = gpu_ctx.upload(loc)
    rng_a, rng_b = genotype.init_rng(123), genotype.init_rng(123)
    mono = dl.solve(scheme, 16, rng_a)
    shard = ldist.solve_sharded(dl, scheme, 16, rng_b, rank=0, world=1)
    dl.free()
    assert np.array_equal(mono.gt_ix, shard["gt_ix"])
    assert np.array_equal(mono.lik_mean, shard["lik_mean"])
    assert np.array_equal(mono.ln_prob, shard["ln_prob"])
    assert mono.n_filtered == shard["n_filtered"] and mono.n_stage_in == shard["n_stage_in"]
    assert np.array_equal(rng_a, rng_b)


def test_sharded_prefilter_two_virtual_ranks_on_device(oracle, gpu_ctx):
    """Rank-local candidate sets computed on the device for 2 and 3 shards merge to the exact survivor list."""
    from locityper_b200 import dist as ldist
    loc = _mk(oracle, 60, 250, 2500, 808)
    ol = oracle.OracleLocus(loc)
    s_ref = oracle.prefilter_scores(ol)
    dl = gpu_ctx.upload(loc)
    G = loc.n_genotypes
    for world in (2, 3):
        for min_size, threads in [(100, 8), (30, 64)]:
            ids, sc = [], []
            for r in range(world):
                a, b = ldist.shard_range(G, r, world)
                i, s = ldist.local_candidates(dl.prefilter_scores(a, b), a, loc.filt_diff, min_size, threads)
                ids.append(i); sc.append(s)
            ids, sc = np.concatenate(ids), np.concatenate(sc)
            dense = np.full(G, -np.inf)
            dense[ids.astype(np.int64)] = sc
            merged = genotype.truncate_ixs(np.sort(ids), dense, loc.filt_diff, min_size, threads)
            ref = oracle.truncate_ixs(np.arange(G), s_ref, loc.filt_diff, min_size, threads)
            assert np.array_equal(merged, ref)
    dl.free()

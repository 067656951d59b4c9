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
        np.testing.assert_allclose(got["lik_var"], ref["lik_var"], rtol=RTOL, atol=0)
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


@pytest.mark.parametrize("R,s", [(300, 12), (300, 40), (300, 1000), (700, 12), (700, 13)])
def test_greedy_samples_of_more_than_11_reads(oracle, gpu_ctx, R, s):
    """Greedy::set_sample_size accepts any s >= 1 (src/solvers/stoch.rs:65-72); beyond 11 rand's index::sample switches
    between Floyd's algorithm (long lists: R = 700) and the in-place partial shuffle (short lists; s >= the number of
    non-trivial reads makes it a full permutation).  Iterations, RNG streams, counts and likelihoods must still match."""
    loc = _mk(oracle, 16, R, 2500, 41 + R)
    gts = list(range(0, loc.n_genotypes, 11))
    _stage_parity(oracle, gpu_ctx, loc, dict(kind="greedy", attempts=2, sample_size=s, plato_size=20), 3, gts)


def test_anneal_stage_parity(oracle, gpu_ctx, small_locus):
    gts = list(range(1, small_locus.n_genotypes, 61))
    _stage_parity(oracle, gpu_ctx, small_locus, dict(kind="anneal", attempts=3, anneal_steps=3000, plato_size=1500),
                  2, gts)


@pytest.mark.parametrize("env", [{"LCTP_WIDE_WINDOWS": "1"}, {"LCTP_ANNEAL_CTAS": "16"}, {"LCTP_ANNEAL_CTAS": "32"},
                                 {"LCTP_WIDE_WINDOWS": "1", "LCTP_ANNEAL_CTAS": "20"}, {"LCTP_NT_GLOBAL": "0"}])
def test_stage_parity_kernel_variants(oracle, gpu_ctx, small_locus, env, monkeypatch):
    """The other instantiations of the stage kernel (64-bit candidate records, the register caps of the annealing
    kernel, the per-read index in shared memory) must give the same iterations, RNG streams, counts and likelihoods."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)                       # read by the library at every launch
    gts = list(range(2, small_locus.n_genotypes, 53))
    _stage_parity(oracle, gpu_ctx, small_locus, dict(kind="anneal", attempts=2, anneal_steps=2500, plato_size=1200), 3, gts)
    _stage_parity(oracle, gpu_ctx, small_locus, dict(kind="greedy", attempts=2), 3, gts)


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
    # Genotyping::to_json text (src/solvers/solve.rs:732-773): the product's C++ writer against the oracle-side
    # formatter fed with the product's numbers (number printing, key order, layout) ...
    names = [f"hap{i}" for i in range(loc.n_haps)]
    as_dict = dict(gt_ix=got.gt_ix, lik_mean=got.lik_mean, lik_var=got.lik_var, ln_prob=got.ln_prob, quality=got.quality,
                   total_reads=got.total_reads, unexpl_reads=got.unexpl_reads,
                   warn_no_probable="NoProbableGenotype" in got.warnings,
                   warn_few_reads=any(w.startswith("FewReads") for w in got.warnings))
    assert got.json_text == oracle.to_json_text(as_dict, loc, names)
    # ... and with the oracle's own numbers when the likelihoods came out bit-identical (they normally do)
    if np.array_equal(got.lik_mean, ref["lik_mean"]) and np.array_equal(got.ln_prob, ref["ln_prob"]) \
            and np.array_equal(got.lik_var, ref["lik_var"], equal_nan=True) and got.quality == ref["quality"]:
        assert got.json_text == oracle.to_json_text(ref, loc, names)


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


@pytest.mark.parametrize("H,R", [(150, 97), (70, 33)])
def test_every_prefilter_kernel_variant_is_bit_exact(oracle, gpu_ctx, H, R, monkeypatch):
    """All tile shapes of k_prefilter_pairs (register tiles 1x1..4x4, cp.async and TMA bulk-copy staging,
    FP64 and integer-pipe max) give the oracle's scores bit for bit; ragged tiles, R not a multiple of the
    chunk, shard sub-ranges."""
    loc = _mk(oracle, H, R, 2500, 500 + H)
    ref = oracle.prefilter_scores(oracle.OracleLocus(loc))
    dl = gpu_ctx.upload(loc)
    G = loc.n_genotypes
    for variant in range(16):
        monkeypatch.setenv("LCTP_PREFILTER_VARIANT", str(variant))
        assert np.array_equal(dl.prefilter_scores(), ref), f"variant {variant}"
        a, b = G // 4, G // 4 + G // 2
        assert np.array_equal(dl.prefilter_scores(a, b), ref[a:b]), f"variant {variant} sub-range"
    dl.free()


@pytest.mark.parametrize("H,R", [(150, 97), (70, 33), (333, 50), (1000, 40)])
def test_balanced_prefilter_kernel_is_bit_exact(oracle, gpu_ctx, H, R, monkeypatch):
    """k_prefilter_bal (variant 16: persistent CTAs, host-planned equal load per SM sub-partition) with the planner's
    own pattern and with explicit per-warp tile widths 2..8, several warps per CTA, R not a multiple of the chunk,
    panels spanning several rounds / a single partly idle CTA, and shard sub-ranges."""
    loc = _mk(oracle, H, R, 2500, 900 + H)
    ref = oracle.prefilter_scores(oracle.OracleLocus(loc))
    dl = gpu_ctx.upload(loc)
    G = loc.n_genotypes
    # 16: cp.async staging by every warp + CTA barrier; 17: bulk-copy producer warp + mbarrier ring;
    # 18: four cp.async producer warps + mbarrier ring
    for variant in ("16", "17", "18"):
        monkeypatch.setenv("LCTP_PREFILTER_VARIANT", variant)
        for pattern in ("", "4,3", "4", "2", "3,2", "2,2,2", "4,3,2", "4,4,4,4", "7", "8", "5,6", "6"):
            if variant == "18" and pattern == "4,4,4,4":
                continue                  # 16 consumer + 4 producer warps exceed the register file
            monkeypatch.setenv("LCTP_PREFILTER_BAL", pattern)
            what = f"variant {variant} pattern '{pattern}'"
            assert np.array_equal(dl.prefilter_scores(), ref), what
            a, b = G // 4, G // 4 + G // 2
            assert np.array_equal(dl.prefilter_scores(a, b), ref[a:b]), what + " sub-range"
            assert np.array_equal(dl.prefilter_scores(G - 3, G), ref[G - 3:]), what + " tail range"
    for stages in ("2", "3", "6"):
        monkeypatch.setenv("LCTP_PREFILTER_BAL_STAGES", stages)
        assert np.array_equal(dl.prefilter_scores(), ref), f"variant 18, {stages} stages"
    dl.free()


def test_positive_matrix_entries_disable_the_integer_max(oracle, gpu_ctx, monkeypatch):
    """dmax_nonpos is only valid for entries <= +0.0: a locus with a positive ln-prob must still be exact."""
    loc = _mk(oracle, 40, 60, 2500, 77)
    loc.pa_ln_prob = loc.pa_ln_prob.copy()
    # per (read, contig) run the first entry is the best one: shifting every entry keeps runs sorted
    loc.pa_ln_prob += 15.0
    loc.unmapped_prob = loc.unmapped_prob + 15.0
    ref = oracle.prefilter_scores(oracle.OracleLocus(loc))
    dl = gpu_ctx.upload(loc)
    assert (dl.best_aln_matrix() > 0).any()
    for variant in (7, 8, 9, 10, 1):
        monkeypatch.setenv("LCTP_PREFILTER_VARIANT", str(variant))
        assert np.array_equal(dl.prefilter_scores(), ref), f"variant {variant}"
    dl.free()


@pytest.mark.parametrize("sample_size", [1, 2, 7, 10, 11])
def test_greedy_sample_sizes_and_many_candidates(oracle, gpu_ctx, sample_size):
    """Lanes-per-slot layouts 32/16/4/3/2 of the greedy loop; 90 % of the reads have secondary locations,
    so most sampled reads have more alternatives than lanes per slot (extra evaluation passes)."""
    loc = _mk(oracle, 14, 180, 2500, 900 + sample_size, multi_frac=0.9)
    gts = list(range(0, loc.n_genotypes, 4))
    _stage_parity(oracle, gpu_ctx, loc, dict(kind="greedy", attempts=2, sample_size=sample_size, plato_size=40), 3, gts)
    _stage_parity(oracle, gpu_ctx, loc, dict(kind="greedy", attempts=1, sample_size=sample_size, best_start=False,
                                             plato_size=25), 2, gts[:10])


def test_fewer_nontrivial_reads_than_the_sample(oracle, gpu_ctx):
    """n_nontrivial < sample_size (amount = n_nontrivial) and genotypes with no non-trivial read at all."""
    loc = _mk(oracle, 6, 12, 2500, 4321, multi_frac=0.0, off_target=0.5)
    gts = list(range(loc.n_genotypes))
    _stage_parity(oracle, gpu_ctx, loc, dict(kind="greedy", attempts=2, plato_size=10), 2, gts)
    _stage_parity(oracle, gpu_ctx, loc, dict(kind="anneal", attempts=2, anneal_steps=200, plato_size=100), 3, gts)


def test_more_workers_than_genotypes(oracle, gpu_ctx, small_locus):
    """T > n: trailing workers get no genotype; their RNG states must come back untouched."""
    gts = [5, 17, 40]
    ixs = np.array(gts, dtype=np.uint64)
    off = np.array([0, 1, 2, 3, 3, 3], dtype=np.uint64)
    rng_ref = _workers(oracle, 5, 8)
    rng_gpu = rng_ref.copy()
    ol = oracle.OracleLocus(small_locus)
    dl = gpu_ctx.upload(small_locus)
    ref = oracle.solve_stage(ol, oracle.Stage(kind="greedy", attempts=1), ixs, off, rng_ref, os_threads=2)
    got = dl.solve_stage(genotype.Stage(kind="greedy", attempts=1), ixs, off, rng_gpu)
    dl.free()
    assert np.array_equal(rng_gpu, rng_ref)
    assert np.array_equal(rng_gpu[3:], _workers(oracle, 5, 8)[3:])
    assert np.array_equal(got["lik_mean"], ref["lik_mean"])


def test_context_pool_concurrent_loci_equal_sequential(oracle, gpu_ctx):
    """Three loci in flight on three contexts (streams + host threads) return exactly what one context
    returns for them one after another -- the bench's execution mode."""
    loci = [_mk(oracle, 20 + 3 * i, 200, 2500, 60 + i) for i in range(3)]
    scheme = genotype.Scheme([genotype.Stage("greedy", attempts=1, in_size=80),
                              genotype.Stage("anneal", attempts=3, in_size=6, anneal_steps=800, plato_size=400)])

    def one(ctx, i, loc):
        dl = ctx.upload(loc)
        rng = genotype.init_rng(1000 + i)
        res = dl.solve(scheme, 48, rng)
        dl.free()
        return res, rng

    seq = [one(gpu_ctx, i, l) for i, l in enumerate(loci)]
    pool = genotype.ContextPool(device=0, k=3)
    for _ in range(2):
        par = pool.map(one, loci)
        for (a, ra), (b, rb) in zip(seq, par):
            assert np.array_equal(a.gt_ix, b.gt_ix) and np.array_equal(a.lik_mean, b.lik_mean)
            assert np.array_equal(a.ln_prob, b.ln_prob) and np.array_equal(ra, rb)
    assert pool.launch_count() > 0
    pool.close()


def test_full_size_properties_c2(oracle, gpu_ctx):
    """BASELINE configs[1] at full size (H=300, G=45,150, R=2,000): size-independent properties.
    (1) scores of sampled genotypes equal a sequential numpy restatement from the device-built matrix;
    (2) the two-shard prefilter equals the full one; (3) the survivor list is sorted by (score desc, id asc);
    (4) the whole solve is deterministic and its stage-1 input has the requested size."""
    loc = synth.make_locus(**synth.config_shape("C2"), seed=2001, table_builder=genotype.build_depth_table)
    dl = gpu_ctx.upload(loc)
    M = dl.best_aln_matrix()
    s = dl.prefilter_scores()
    rs = np.random.default_rng(5)
    for g in rs.integers(0, loc.n_genotypes, 64):
        i, j = loc.genotype_tuple(int(g))
        acc = 0.0
        for v in np.maximum(M[i], M[j]):
            acc += v
        assert s[g] == acc + 0.0
    G = loc.n_genotypes
    assert np.array_equal(np.concatenate([dl.prefilter_scores(0, G // 2), dl.prefilter_scores(G // 2, G)]), s)
    surv = dl.prefilter(5000, 64)
    key = list(zip(-s[surv.astype(np.int64)], surv))
    assert key == sorted(key) and len(surv) >= 5000
    scheme = genotype.Scheme.parse(["greedy:i=5k,a=1"])
    a = dl.solve(scheme, 4736, genotype.init_rng(2001))
    b = dl.solve(scheme, 4736, genotype.init_rng(2001))
    dl.free()
    assert a.n_stage_in[0] == 5000 or a.n_filtered <= 5000
    assert np.array_equal(a.gt_ix, b.gt_ix) and np.array_equal(a.lik_mean, b.lik_mean)
    assert loc.genotype_tuple(int(a.gt_ix[0])) == tuple(loc.truth)

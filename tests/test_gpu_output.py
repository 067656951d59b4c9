"""SURVEY 8(f) rank 4, the output side: the per-locus debug tables of the reference's `--debug` (sol.csv, sol_ext.csv,
depth.csv; src/solvers/solve.rs:852-952, src/model/assgn.rs:356-372,413-425) written by the product while a debug sink is
open, and the assignment counts of the reported genotypes (Prediction::assgn_counts, the input of write_bam,
src/model/bam.rs:356-413).  Checked against the oracle's writers of the same tables, text for text (rows sorted: inside a
stage the reference's row order depends on thread timing)."""
import ctypes as C
import os

import numpy as np
import pytest

from locityper_b200 import genotype, synth

pytestmark = pytest.mark.gpu


def _oracle_tables(oracle, loc, names, specs, threads, seed, out):
    os.makedirs(out, exist_ok=True)
    lib = oracle.lib()
    lib.lcto_debug_open.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p]
    lib.lcto_debug_open_depth.argtypes = [C.c_char_p]
    cn = (C.c_char_p * len(names))(*[n.encode() for n in names])
    assert lib.lcto_debug_open(os.path.join(out, "sol.csv").encode(), os.path.join(out, "sol_ext.csv").encode(), cn) == 0
    assert lib.lcto_debug_open_depth(os.path.join(out, "depth.csv").encode()) == 0
    scheme = [oracle.Stage(st.kind, attempts=st.attempts, in_size=st.in_size, best_start=st.best_start,
                           sample_size=st.sample_size, plato_size=st.plato_size, anneal_steps=st.anneal_steps,
                           init_prob=st.init_prob) for st in genotype.Scheme.parse(specs).stages]
    res = oracle.solve(oracle.OracleLocus(loc), scheme, threads, oracle.Rng.from_seed(seed), os_threads=2)
    lib.lcto_debug_close()
    return res


def _rows(path):
    lines = open(path).read().split("\n")
    assert lines[-1] == ""
    return lines[0], sorted(lines[1:-1])


@pytest.mark.parametrize("ploidy,threads", [(2, 8), (3, 1)])
def test_debug_tables_equal_the_oracle_writers(oracle, gpu_ctx, tmp_path, ploidy, threads):
    loc = synth.make_locus(9 if ploidy == 3 else 20, 200, 2500, seed=31 + ploidy, ploidy=ploidy,
                           table_builder=genotype.build_depth_table)
    names = [f"HG{i:03d}.{i % 2 + 1}" for i in range(loc.n_haps)]
    specs = ["greedy:i=40,a=2", "anneal:i=6,a=3,n=600,p=300"]
    o_dir, g_dir = str(tmp_path / "oracle"), str(tmp_path / "gpu")
    ref = _oracle_tables(oracle, loc, names, specs, threads, 5, o_dir)
    os.makedirs(g_dir)
    dl = gpu_ctx.upload(loc)
    gpu_ctx.debug_open(g_dir, 2, names)
    try:
        got = dl.solve(genotype.Scheme.parse(specs), threads, genotype.init_rng(5), hap_names=names)
    finally:
        gpu_ctx.debug_close()
    assert np.array_equal(got.gt_ix, ref["gt_ix"])
    for name in ("sol.csv", "sol_ext.csv", "depth.csv"):
        h1, r1 = _rows(os.path.join(g_dir, name))
        h2, r2 = _rows(os.path.join(o_dir, name))
        assert h1 == h2, name
        assert len(r1) == len(r2) and r1 == r2, name
    # level 0 (the reference without --debug): sol.csv with the rows of the LAST stage only, no other table
    g0 = str(tmp_path / "gpu0")
    os.makedirs(g0)
    gpu_ctx.debug_open(g0, 0, names)
    dl.solve(genotype.Scheme.parse(specs), threads, genotype.init_rng(5))
    gpu_ctx.debug_close()
    assert sorted(os.listdir(g0)) == ["sol.csv"]
    _, r0 = _rows(os.path.join(g0, "sol.csv"))
    _, r_all = _rows(os.path.join(g_dir, "sol.csv"))
    assert r0 == [r for r in r_all if r.startswith("2\t")]
    dl.free()


def test_stage_debug_fields_and_counts(oracle, gpu_ctx, small_locus):
    loc = small_locus
    dl = gpu_ctx.upload(loc)
    ol = oracle.OracleLocus(loc)
    st_g = genotype.Stage.parse(0, "greedy:i=30,a=3")
    st_o = oracle.Stage("greedy", attempts=3, in_size=30)
    ixs = np.arange(30, dtype=np.uint64)
    off = np.array([0, 10, 20, 30], dtype=np.uint64)
    r1 = genotype.worker_streams(genotype.init_rng(3), 3)
    r2 = r1.copy()
    d = dl.solve_stage_debug(st_g, ixs, off, r1)
    o = oracle.solve_stage(ol, st_o, ixs, off, r2)
    assert np.array_equal(d["lik_mean"], o["lik_mean"]) and np.array_equal(d["liks"], o["liks"])
    assert np.all(d["unmapped"] + d["out_of_bounds"] <= 2 * loc.n_reads)
    assert np.all(d["win_depth"][:, 0] == d["unmapped"]) and np.all(d["win_depth"][:, 1] == d["out_of_bounds"])
    assert np.all(d["win_weight"][:, :2] == 0.0) and np.all(d["win_lik"][:, :2] == 0.0)
    # every read end sits in exactly one window: depths of an attempt sum to 2 per read pair
    assert np.all(d["win_depth"].sum(axis=1) == 2 * loc.n_reads)
    # depth_lik (updated move by move, src/model/assgn.rs:336) = sum of the per-window ln-probabilities up to rounding
    np.testing.assert_allclose(d["win_lik"].sum(axis=1), d["depth_lik"], rtol=1e-9)
    # ... and the attempt likelihood is depth_contrib * depth_lik + aln_contrib * aln_lik (prior 0 here), exactly
    skew = loc.lik_skew
    depth_contrib, aln_contrib = 1.0 + skew, 1.0 - skew      # src/model/assgn.rs:80-81
    assert np.array_equal(depth_contrib * d["depth_lik"] + aln_contrib * d["aln_lik"], d["liks"].reshape(-1))
    # counts of the reported genotypes = the oracle's counts of the same genotypes in the last stage
    scheme = genotype.Scheme.parse(["greedy:i=25,a=4"])
    g, counts = dl.solve_counts(scheme, 8, genotype.init_rng(9), 3)
    surv = dl.prefilter(25, 8)
    # replay the stage through the oracle with counts: same survivors, same shuffle, same worker streams
    ro = oracle.Rng.from_seed(9)
    ref = oracle.solve(ol, [oracle.Stage("greedy", attempts=4, in_size=25)], 8, ro, os_threads=2)
    assert np.array_equal(g.gt_ix, ref["gt_ix"])
    st = genotype.init_rng(9)
    streams = genotype.worker_streams(st, 8)
    ix2 = np.ascontiguousarray(surv.copy())
    woff = genotype.plan_stage(st, ix2, 8)
    oc = oracle.solve_stage(ol, oracle.Stage("greedy", attempts=4, in_size=25), ix2, woff,
                            np.ascontiguousarray(streams[:len(woff) - 1]), want_counts=True, counts_cap=1 << 22)
    for k in range(3):
        q = int(np.nonzero(ix2 == g.gt_ix[k])[0][0])
        want = oc["counts"][int(oc["counts_off"][q]):int(oc["counts_off"][q + 1])]
        assert np.array_equal(counts[k], want)
        assert counts[k].sum() == 4 * loc.n_reads            # every read is somewhere in each of the 4 attempts
    dl.free()


def test_edge_cases_of_the_widened_entry_points(gpu_ctx, tmp_path):
    """Empty and degenerate inputs of the section-8(f) entry points: nothing to do is not an error, malformed input is."""
    import ctypes as C
    from locityper_b200 import ffi
    lib = gpu_ctx.lib
    # no reads: nothing is launched
    ts = genotype.TargetSeqs(seqs=[b"ACGT" * 100], seq_locus=np.zeros(1, dtype=np.uint32),
                             kmer_counts=[np.zeros(400 + 1 - 25, dtype=np.uint16)], base_k=25, minimizer_k=15, minimizer_w=10,
                             thresh_kmer_count=10, match_frac=0.5)
    t = genotype.Targets(gpu_ctx, ts)
    assert t.recruit(genotype.Reads(seq1=[])) == []
    # reads shorter than k have no minimizers and are not recruited; an all-N read neither
    assert t.recruit(genotype.Reads(seq1=[b"ACG", b"", b"N" * 150], seq2=[b"ACG", b"ACGT", b"N" * 150])) == [[], [], []]
    t.free()
    assert genotype.minimizers(gpu_ctx, [], 15, 10) == []
    # a k-mer count array of the wrong length is refused (the reference asserts, recruit.rs:698-699)
    ts.kmer_counts = [np.zeros(10, dtype=np.uint16)]
    with pytest.raises(Exception):
        genotype.Targets(gpu_ctx, ts)
    # the debug sink needs an existing directory
    with pytest.raises(Exception):
        gpu_ctx.debug_open(str(tmp_path / "does" / "not" / "exist"), 2, ["a", "b"])
    gpu_ctx.debug_close()                      # closing a closed sink is a no-op
    # read ends: zero groups
    empty = genotype.ReadEnds(alns=genotype.Alns(cigar_off=np.zeros(1, dtype=np.uint64), cigar_ops=np.zeros(0, dtype=np.uint32),
                                                 aln_start=np.zeros(0, dtype=np.uint32), aln_end=np.zeros(0, dtype=np.uint32),
                                                 contig_len=np.zeros(0, dtype=np.uint32), passable_dist=np.zeros(0, dtype=np.uint32),
                                                 ln_oper=(-0.005, -5.8, -6.5, -6.9, -5.8)),
                              grp_off=np.zeros(1, dtype=np.uint64), rec_contig=np.zeros(0, dtype=np.uint32),
                              grp_read_end=np.zeros(0, dtype=np.uint8), grp_read_len=np.zeros(0, dtype=np.uint32),
                              grp_good_dist=np.zeros(0, dtype=np.uint32), grp_passable_dist=np.zeros(0, dtype=np.uint32),
                              grp_neighb_complexity=np.zeros(0), poor_compl=0.5, poor_compl_edit=0.07)
    out = genotype.collect_read_ends(gpu_ctx, empty)
    assert len(out["ok"]) == 0 and len(out["kept_rec"]) == 0
    # NULL handles return error codes, never crash
    assert lib.lctp_recruit_short(gpu_ctx._h, None, None, 8, None, None) == -1
    assert lib.lctp_debug_open(None, b"/tmp", 1, None, 0) == -1
    assert lib.lctp_solve_counts(None, None, 0, 1, None, None, 0, None, None, 0) == -1

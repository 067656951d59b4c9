#!/usr/bin/env python
"""Generates the committed golden fixtures of tests/golden/ from the CPU oracle (oracle/, plain C).

    python tests/golden/make_golden.py          # rewrites locus_*.npz and golden_*.json

PARITY UNPINNED BY THE REFERENCE: tprodanov/locityper ships no tests, golden vectors or fixtures for
this path and cannot be built in this image (Rust, no cargo), so these vectors
are outputs of OUR restatement of its algorithm.  They pin (a) the oracle against regressions and
(b) the CUDA path on the GPU box without needing the oracle there.  Third-party known-answer vectors
(xoshiro256++, SplitMix64) are pinned separately in tests/test_oracle_pins.py.

Floats are stored as C99 hex strings (float.hex) so the comparison is bit-exact.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from locityper_b200 import synth          # noqa: E402
from oracle import lcto_py as O           # noqa: E402

CASES = {
    # name: (make_locus kwargs, stage tests, full-solve scheme, T)
    "illumina_p2": dict(mk=dict(n_haps=12, n_reads=90, locus_len=2000, seed=31337), threads=6),
    "hifi_p2": dict(mk=dict(n_haps=8, n_reads=40, locus_len=16000, seed=31338, tech="hifi"), threads=4),
    "illumina_p3": dict(mk=dict(n_haps=6, n_reads=70, locus_len=2000, seed=31339, ploidy=3), threads=3),
}


def hx(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).reshape(-1)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def worker_states(n, seed):
    import ctypes as C
    rng = O.Rng.from_seed(seed)
    st = np.zeros((n, 4), dtype=np.uint64)
    for w in range(n):
        st[w] = rng.state()
        O.lib().lcto_rng_jump(C.byref(rng))
    return st


def stage_cases(loc):
    G = loc.n_genotypes
    return [
        dict(name="greedy_default", kw=dict(kind="greedy", attempts=2), workers=3, gts=list(range(0, G, 5)), seed=3),
        dict(name="greedy_random_s4", kw=dict(kind="greedy", attempts=2, best_start=False, sample_size=4, plato_size=25),
             workers=2, gts=list(range(1, G, 9)), seed=4),
        dict(name="anneal", kw=dict(kind="anneal", attempts=3, anneal_steps=1500, plato_size=600), workers=2,
             gts=list(range(2, G, 13)), seed=5),
    ]


def make_case(name, spec):
    loc = synth.make_locus(**spec["mk"], table_builder=O.build_depth_table)
    synth.save_locus(loc, os.path.join(HERE, f"locus_{name}.npz"))
    loc = synth.load_locus(os.path.join(HERE, f"locus_{name}.npz"))      # what the tests will see
    ol = O.OracleLocus(loc)
    G = loc.n_genotypes
    out = dict(name=name, n_genotypes=G, truth=list(loc.truth))
    M = O.best_aln_matrix(ol)
    out["matrix_sha256"] = sha(M)
    scores = O.prefilter_scores(ol, M=M)
    out["prefilter_scores"] = hx(scores)
    out["truncate"] = []
    for min_size, threads in [(5, 1), (20, 8), (10 ** 6, 8)]:
        surv = O.truncate_ixs(np.arange(G), scores, loc.filt_diff, min_size, threads)
        out["truncate"].append(dict(min_size=min_size, threads=threads, survivors=[int(x) for x in surv]))
    out["stages"] = []
    for sc in stage_cases(loc):
        ixs = np.array(sc["gts"], dtype=np.uint64)
        off = np.linspace(0, len(ixs), sc["workers"] + 1).astype(np.uint64)
        rng = worker_states(sc["workers"], sc["seed"])
        rng_in = rng.copy()
        cap = int(len(ixs) * loc.n_reads * (10 * loc.ploidy + 1))
        r = O.solve_stage(ol, O.Stage(**sc["kw"]), ixs, off, rng, os_threads=1, want_counts=True, counts_cap=cap)
        n = int(r["counts_off"][-1])
        out["stages"].append(dict(
            name=sc["name"], kw=sc["kw"], ixs=[int(x) for x in ixs], off=[int(x) for x in off],
            rng_in=[[int(v) for v in row] for row in rng_in], rng_out=[[int(v) for v in row] for row in rng],
            liks=hx(r["liks"]), lik_mean=hx(r["lik_mean"]), lik_var=hx(r["lik_var"]),
            n_alns=[int(x) for x in r["n_alns"]], iters=[int(x) for x in r["iters"]],
            counts_off=[int(x) for x in r["counts_off"]], counts_sha256=sha(r["counts"][:n]),
            counts_head=[int(x) for x in r["counts"][:64]]))
    scheme = [O.Stage("greedy", attempts=1, in_size=max(8, G // 3)),
              O.Stage("anneal", attempts=4, in_size=5, anneal_steps=1500, plato_size=600)]
    out["solve"] = []
    for T in (1, spec["threads"]):
        rng = O.Rng.from_seed(2024)
        r = O.solve(ol, scheme, T, rng, os_threads=1, want_scores=True)
        out["solve"].append(dict(
            threads=T, seed=2024,
            scheme=[dict(kind=s.kind, attempts=s.attempts, in_size=s.in_size, anneal_steps=s.anneal_steps,
                         plato_size=s.plato_size) for s in scheme],
            gt_ix=[int(x) for x in r["gt_ix"]], lik_mean=hx(r["lik_mean"]), lik_var=hx(r["lik_var"]),
            ln_prob=hx(r["ln_prob"]), quality=float(r["quality"]).hex(), unexpl_reads=int(r["unexpl_reads"]),
            n_filtered=int(r["n_filtered"]), n_stage_in=r["n_stage_in"], filtered_ixs=[int(x) for x in r["filtered_ixs"]],
            rng_out=rng.state()))
    with open(os.path.join(HERE, f"golden_{name}.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    return out


MATES_FIELDS = ("ma_off", "ma_contig", "ma_flags", "ma_start", "ma_end", "ma_ln_prob", "ins_ln_pmf", "read_weight")
ALNS_FIELDS = ("cigar_off", "cigar_ops", "aln_start", "aln_end", "contig_len", "passable_dist")


def make_upstream():
    """Fixtures of the two upstream slices: inputs AND oracle outputs in one .npz each
    (f64 arrays are stored raw, so the comparison is bit-exact)."""
    from locityper_b200 import genotype
    kw = synth.make_mates(6, 40, 2500, 424242, multi_frac=0.4)
    kw["read_weight"] = np.linspace(0.5, 1.0, 40)
    m = genotype.Mates(**kw)
    out = O.pair_alignments(m)
    scal = np.array([m.n_reads, m.n_haps, m.max_alns, m.unmapped_penalty, m.insert_penalty, m.prob_diff], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "pairs_small.npz"), _scalars=scal,
                        **{k: np.asarray(getattr(m, k)) for k in MATES_FIELDS},
                        **{"out_" + k: v for k, v in out.items()})
    a = genotype.Alns(**synth.make_alns(600, 434343))
    out = O.rescore_alignments(a)
    np.savez_compressed(os.path.join(HERE, "rescore_small.npz"), ln_oper=np.array(a.ln_oper, dtype=np.float64),
                        **{k: np.asarray(getattr(a, k)) for k in ALNS_FIELDS},
                        **{"out_" + k: v for k, v in out.items()})
    return len(m.ma_contig), a.n_alns


PRELIM_FIELDS = ("read_group", "grp_off", "rec_contig", "rec_start", "rec_end", "rec_strand", "rec_ln_prob", "grp_ok",
                 "grp_best_edit", "grp_thr_dist", "grp_n_kept", "kept_rec", "contig_len", "read_weight")


def make_group():
    """Fixture of the step in front of the pairing (lctp_group_reads, SURVEY 8(f) rank 1 remainder): a paired-end and a
    single-end input (the generator of tests/test_group.py) with the oracle's outputs."""
    sys.path.insert(0, os.path.dirname(HERE))
    from test_group import _random_prelim
    n = 0
    for tag, single in (("pe", False), ("se", True)):
        p = _random_prelim(515151 + int(single), n_reads=120, single_end=single)
        out = O.group_reads(p)
        scal = np.array([p.min_weight, p.boundary, int(p.single_end)], dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, f"group_{tag}_small.npz"), _scalars=scal,
                            **{k: np.asarray(getattr(p, k)) for k in PRELIM_FIELDS},
                            **{"out_" + k: np.asarray(v) for k, v in out.items()})
        n += len(p.rec_contig)
    return n


if __name__ == "__main__":
    if sys.argv[1:] == ["group"]:
        print("group fixtures: %d alignment records" % make_group())
        sys.exit(0)
    for name, spec in CASES.items():
        o = make_case(name, spec)
        print(name, "G =", o["n_genotypes"], "call =", o["solve"][-1]["gt_ix"][:1], "truth =", o["truth"])
    print("upstream fixtures: %d mate records, %d alignment records" % make_upstream())
    print("group fixtures: %d alignment records" % make_group())

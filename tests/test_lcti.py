"""tools/lcti.py: the `.lcti` dump format that rust/gpu.rs `FlatLocus::dump` writes from inside the reference
(SURVEY.md Appendix D / oracle/rust_diff.sh).  Round trip and the oracle's debug dumps in the reference's row formats."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_lcti_round_trip_and_debug_dumps(oracle, small_locus, tmp_path):
    import lcti
    names = [f"HG{i:03d}" for i in range(small_locus.n_haps)]
    state = oracle.Rng.from_seed(77).state()
    d = str(tmp_path / "dump")
    lcti.write(small_locus, d, hap_names=names, rng_state=state)
    loc, names2, st = lcti.read(d)
    assert names2 == names and list(st) == state
    for k in ("unmapped_prob", "pa_off", "pa_contig", "pa_ln_prob", "pa_mid1", "pa_mid2", "hap_len", "hap_n_windows",
              "hap_reg_start", "hap_pos_off", "pos_weight", "pos_gc", "depth_table"):
        assert np.array_equal(getattr(loc, k), getattr(small_locus, k)), k
    for k in ("n_haps", "n_reads", "ploidy", "window", "left_padding", "depth_k", "tweak", "prob_diff", "lik_skew",
              "min_weight", "filt_diff", "prob_thresh"):
        assert getattr(loc, k) == getattr(small_locus, k), k          # floats travel as bit patterns: exact
    out = str(tmp_path / "out")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "lcti_solve.py"), d, "--threads", "4", "--out", out,
                        "--scheme", "greedy:i=50,a=2", "anneal:i=5,a=3,n=500,p=200", "--os-threads", "2"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sol = open(os.path.join(out, "sol.csv")).read().split("\n")
    ext = open(os.path.join(out, "sol_ext.csv")).read().split("\n")
    assert sol[0] == "stage\tgenotype\tscore"                                       # solve.rs:938
    assert ext[0] == "stage\tgenotype\tattempt\ttotal_reads\tunmapped\tout_of_bounds\taln_lik\tdepth_lik\tlik"
    rows0 = [l for l in sol[1:] if l.startswith("0\t")]
    assert len(rows0) == small_locus.n_genotypes                                    # one prefilter row per genotype
    assert all(l.split("\t")[1].count(",") == small_locus.ploidy - 1 for l in rows0[:20])
    # the same solve called directly gives the same final call as the one written to res.json
    direct = oracle.solve(oracle.OracleLocus(small_locus),
                          [oracle.Stage("greedy", attempts=2, in_size=50),
                           oracle.Stage("anneal", attempts=3, in_size=5, anneal_steps=500, plato_size=200)],
                          4, oracle.Rng.from_state(state), os_threads=2)
    import json
    js = json.load(open(os.path.join(out, "res.json")))
    assert js["genotype"] == ",".join(names[h] for h in small_locus.genotype_tuple(int(direct["gt_ix"][0])))
    n1 = sum(1 for l in ext[1:] if l.startswith("1\t"))
    assert n1 == direct["n_stage_in"][0] * 2                                        # attempts rows of stage 1

"""Oracle side of SURVEY.md Appendix D (oracle/rust_diff.sh): re-run the CPU oracle on a `.lcti` dump of the
`solve::Data` a real `locityper genotype --debug 2` run saw (rust/gpu.rs FlatLocus::dump, read by tools/lcti.py) and
write sol.csv / sol_ext.csv in the reference's own row formats plus res.json, to be diffed against the reference's.

TEST INFRASTRUCTURE ONLY (it lives under oracle/ for that reason).

    python oracle/lcti_solve.py DIR --threads 8 --out OUT [--scheme greedy:i=5k,a=1 anneal:i=20,a=20] [--seed S]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import lcti  # noqa: E402  (pure reader/writer of the dump format)
from oracle import lcto_py as oracle  # noqa: E402
from locityper_b200 import genotype  # noqa: E402  (scheme parser only)


def scheme_of(specs):
    out = []
    for st in genotype.Scheme.parse(specs).stages:
        out.append(oracle.Stage(st.kind, attempts=st.attempts, in_size=st.in_size, best_start=st.best_start,
                                sample_size=st.sample_size, plato_size=st.plato_size, anneal_steps=st.anneal_steps,
                                init_prob=st.init_prob))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dir"); ap.add_argument("--threads", type=int, default=8); ap.add_argument("--out", required=True)
    ap.add_argument("--seed", type=int, default=None, help="-s SEED of the run when the dump carries no rng_state.u64")
    ap.add_argument("--scheme", nargs="*", default=["greedy:i=5k,a=1", "anneal:i=20,a=20"])
    ap.add_argument("--os-threads", type=int, default=os.cpu_count() or 4)
    a = ap.parse_args()
    loc, names, st = lcti.read(a.dir)
    os.makedirs(a.out, exist_ok=True)
    rng = oracle.Rng.from_state(st) if st is not None else oracle.Rng.from_seed(a.seed)
    cn = (C.c_char_p * len(names))(*[n.encode() for n in names])
    lib = oracle.lib()
    lib.lcto_debug_open.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p]
    rc = lib.lcto_debug_open(os.path.join(a.out, "sol.csv").encode(), os.path.join(a.out, "sol_ext.csv").encode(), cn)
    assert rc == 0
    res = oracle.solve(oracle.OracleLocus(loc), scheme_of(a.scheme), a.threads, rng, os_threads=a.os_threads)
    lib.lcto_debug_close()
    with open(os.path.join(a.out, "res.json"), "w") as f:
        f.write(oracle.to_json_text(res, loc, names))
    print("oracle call:", ",".join(names[h] for h in loc.genotype_tuple(int(res["gt_ix"][0]))), "quality", res["quality"])


if __name__ == "__main__":
    main()

"""Do the solver kernels of loci in flight on different contexts overlap on the device?  Prints, per locus, the wall
clock interval of its solve() call relative to the common start (3 contexts, 3 host threads, loci resident)."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
from locityper_b200 import genotype, synth

ap = argparse.ArgumentParser()
ap.add_argument("--max-resident", type=int, default=0)
ap.add_argument("--loci", type=int, default=3)
a = ap.parse_args()
loci = [synth.make_locus(**synth.config_shape("C2"), seed=2001 + i, table_builder=genotype.build_depth_table) for i in range(a.loci)]
ctxs = [genotype.Context(0, max_resident_workers=a.max_resident) for _ in loci]
dls = [c.upload(l) for c, l in zip(ctxs, loci)]
scheme = genotype.Scheme.parse(["greedy:i=5k,a=1"])
for rep in range(3):
    out = [None] * len(loci)
    bar = threading.Barrier(len(loci) + 1)

    def work(i):
        bar.wait()
        t0 = time.perf_counter()
        dls[i].solve(scheme, 4736, genotype.init_rng(2001 + i))
        out[i] = (t0, time.perf_counter())

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(loci))]
    for t in th:
        t.start()
    bar.wait()
    T0 = time.perf_counter()
    for t in th:
        t.join()
    print(f"rep {rep}: " + "  ".join(f"locus {i}: {1e3*(s-T0):6.2f} .. {1e3*(e-T0):6.2f} ms" for i, (s, e) in enumerate(out)),
          f" kernels {ctxs[0].stats(reset=True)['stage_ms']:.2f} {ctxs[1].stats(reset=True)['stage_ms']:.2f} ms", flush=True)

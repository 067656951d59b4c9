"""Driver for timing / profiling lctp_rescore_alignments (first slice)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
from locityper_b200 import genotype, synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_500_000, help="alignment records (C2: ~1.5 M mates per locus)")
ap.add_argument("--passes", type=int, default=5)
a = ap.parse_args()
t0 = time.time()
al = genotype.Alns(**synth.make_alns(a.n, 7))
print(f"{a.n} alignments, {len(al.cigar_ops)} CIGAR operations generated in {time.time()-t0:.1f}s", flush=True)
ctx = genotype.Context(0)
for i in range(a.passes):
    if i == 1:
        ctx.stats(reset=True)
        t1 = time.time()
    out = genotype.rescore_alignments(ctx, al)
wall = (time.time() - t1) / max(1, a.passes - 1)
st = ctx.stats()
ms = st["rescore_ms"] / max(1, st["rescore_launches"])
bytes_alg = len(al.cigar_ops) * 4 + a.n * (8 + 16) + a.n * 17
print(f"rescoring kernel: {ms:.4f} ms/launch, {a.n/ms/1e6:.2f} G alignments/s, {bytes_alg/ms/1e6:.1f} GB/s algorithmic; "
      f"whole call incl. H2D/D2H {wall*1e3:.1f} ms")

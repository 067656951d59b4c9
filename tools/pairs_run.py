"""Driver for timing / profiling lctp_pair_alignments at a BASELINE shape."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
from locityper_b200 import genotype, synth

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--passes", type=int, default=4)
a = ap.parse_args()
sh = synth.config_shape(a.config)
t0 = time.time()
m = genotype.Mates(**synth.make_mates(sh["n_haps"], sh["n_reads"], sh["locus_len"], 77))
N = len(m.ma_contig)
print(f"{a.config}: H={m.n_haps} R={m.n_reads} mates={N} generated in {time.time()-t0:.1f}s", flush=True)
ctx = genotype.Context(0)
for i in range(a.passes):
    if i == 1:
        ctx.stats(reset=True)
    t1 = time.time()
    out = genotype.pair_alignments(ctx, m)
    wall = time.time() - t1
st = ctx.stats()
n = max(1, st["pairing_launches"])
ms = st["pairing_ms"] / n
pairs = len(out["pa_contig"])
bytes_alg = 2 * N * (4 + 1 + 4 + 4 + 8) + pairs * (4 + 8 + 4 + 4) + N * 12   # two passes over the records + output + counts/offsets
print(f"pairing kernels: {ms:.3f} ms/call, {N/ms/1e6:.2f} G mates/s, {bytes_alg/ms/1e6:.1f} GB/s algorithmic; "
      f"{pairs} pair alignments; whole call incl. H2D/D2H {wall*1e3:.1f} ms")

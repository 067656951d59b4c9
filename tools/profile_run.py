"""Small driver for ncu: one C2-shaped locus, prefilter + greedy stage, N passes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
from locityper_b200 import genotype, synth

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--shape", default="", help="H,R,L override")
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--threads", type=int, default=4736)
ap.add_argument("--resident", type=int, default=0, help="max resident workers (0 = library default)")
ap.add_argument("--scheme", nargs="*", default=["greedy:i=5k,a=1"])
a = ap.parse_args()
sh = synth.config_shape(a.config)
if a.shape:
    h, r, l = (int(x) for x in a.shape.split(","))
    sh.update(n_haps=h, n_reads=r, locus_len=l)
loc = synth.make_locus(**sh, seed=2001, table_builder=genotype.build_depth_table)
ctx = genotype.Context(0, max_resident_workers=a.resident)
dl = ctx.upload(loc)
scheme = genotype.Scheme.parse(a.scheme)
for i in range(a.passes):
    res = dl.solve(scheme, a.threads, genotype.init_rng(2001))
print(ctx.stats(), loc.genotype_tuple(int(res.gt_ix[0])), loc.truth)

"""Driver for timing / profiling the prefilter kernel alone (a2) at a BASELINE shape (default C4: H=1000, R=10k)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
from locityper_b200 import genotype, synth

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C4")
ap.add_argument("--passes", type=int, default=5)
ap.add_argument("--shards", type=int, default=1, help="time the genotype range of shard 0 and of the last shard of N (multi-GPU partition)")
ap.add_argument("--variants", default="", help="e.g. '1;16:;16:4,3' = kernel variants (and balanced patterns) to time")
a = ap.parse_args()
t0 = time.time()
loc = synth.make_locus(**synth.config_shape(a.config), seed=4001, table_builder=genotype.build_depth_table)
print(f"locus {a.config}: H={loc.n_haps} R={loc.n_reads} G={loc.n_genotypes} generated in {time.time()-t0:.1f}s", flush=True)
ctx = genotype.Context(0)
rate = ctx.fp64_rate()
print(f"FP64 pipe: {rate/1e12:.2f} T lane-instructions/s (DADD microbenchmark)")
dl = ctx.upload(loc)
el = loc.n_genotypes * loc.n_reads
ref = None
# "--variants 1;16:;16:4,3": LCTP_PREFILTER_VARIANT[:LCTP_PREFILTER_BAL] per entry, all on the same uploaded locus
for spec in (a.variants.split(";") if a.variants else [None]):
    label = "default"
    if spec is not None:
        v, _, pat = spec.partition(":")
        os.environ["LCTP_PREFILTER_VARIANT"] = v
        os.environ["LCTP_PREFILTER_BAL"] = pat
        label = f"variant {v}" + (f" pattern {pat or 'auto'}" if v in ("16", "17", "18") else "")
    if a.shards > 1:
        from locityper_b200 import dist
        for rank in (0, a.shards - 1):
            ga, gb = dist.shard_range(loc.n_genotypes, rank, a.shards)
            for i in range(a.passes):
                if i == 1:
                    ctx.stats(reset=True)
                s = dl.prefilter_scores(ga, gb)
            st = ctx.stats()
            ms = st["prefilter_ms"] / max(1, st["prefilter_launches"])
            print(f"prefilter [{label}] shard {rank}/{a.shards} ({gb - ga} genotypes): {ms:.4f} ms/launch, "
                  f"{(gb - ga) * loc.n_reads / ms / 1e9:.3f} T elements/s, checksum {float(s.sum()):.6e}", flush=True)
        continue
    for i in range(a.passes):
        if i == 1:
            ctx.stats(reset=True)
        s = dl.prefilter_scores()
    st = ctx.stats()
    n = st["prefilter_launches"]
    ms = st["prefilter_ms"] / max(1, n)
    same = "" if ref is None else (" == first" if (s == ref).all() else " DIFFERS FROM FIRST")
    if ref is None:
        ref = s
    print(f"prefilter [{label}]: {ms:.4f} ms/launch, {el/ms/1e9:.3f} T elements/s (1 max + 1 add each), "
          f"{loc.n_genotypes*(2*loc.n_reads*8+8)/ms/1e6:.1f} GB/s algorithmic, checksum {float(s.sum()):.6e}{same}; "
          f"{2*el/ms/1e9*1e12/rate*100:.1f}% of the FP64-pipe roofline", flush=True)

"""Timing / profiling driver of short-read recruitment: N read pairs of 2 x 150 bp against the targets of
a few loci; kernel time from the library's CUDA events, whole call with H2D / D2H, reads per second and bytes per second."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
from locityper_b200 import genotype

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=1_000_000)
ap.add_argument("--loci", type=int, default=8)
ap.add_argument("--alleles", type=int, default=40)
ap.add_argument("--length", type=int, default=5000)
a = ap.parse_args()
rng = np.random.default_rng(1)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
seqs, locus, counts = [], [], []
for l in range(a.loci):
    base = rng.choice(ACGT, a.length)
    for k in range(a.alleles):
        s = base.copy()
        m = rng.random(a.length) < 0.01
        s[m] = rng.choice(ACGT, int(m.sum()))
        seqs.append(s.tobytes()); locus.append(l)
        counts.append(rng.integers(0, 5, a.length + 1 - 25).astype(np.uint16))
ts = genotype.TargetSeqs(seqs=seqs, seq_locus=np.array(locus, dtype=np.uint32), kmer_counts=counts, base_k=25, minimizer_k=15,
                         minimizer_w=10, thresh_kmer_count=10, match_frac=0.5)
# reads: 10 % from the targets, the rest random (a WGS sample is almost entirely off-target)
n = a.pairs
all1 = rng.choice(ACGT, (n, 150)); all2 = rng.choice(ACGT, (n, 150))
on = np.nonzero(rng.random(n) < 0.1)[0]
comp = np.zeros(256, dtype=np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
for r in on:
    s = np.frombuffer(seqs[int(rng.integers(0, len(seqs)))], dtype=np.uint8)
    st = int(rng.integers(0, a.length - 400))
    all1[r] = s[st:st + 150]; all2[r] = comp[s[st + 200:st + 350]][::-1]
reads = genotype.Reads(seq1=[bytes(x) for x in all1], seq2=[bytes(x) for x in all2])
ctx = genotype.Context(0)
t0 = time.perf_counter(); t = genotype.Targets(ctx, ts); t_build = time.perf_counter() - t0
for i in range(3):
    if i == 1: ctx.stats(reset=True)
    t0 = time.perf_counter(); got = t.recruit(reads); wall = time.perf_counter() - t0
st = ctx.stats()
ms = st["recruit_ms"] / max(1, st["recruit_launches"])
bases = 2 * 150 * n
print(f"[recruit] {n} pairs ({bases/1e6:.0f} MB of bases), {len(seqs)} target sequences: targets built in {t_build*1e3:.1f} ms; "
      f"kernel {ms:.3f} ms = {n/ms/1e3:.1f} M pairs/s = {bases/ms/1e6:.1f} GB/s of bases; whole call (python lists -> arrays, H2D, D2H) "
      f"{wall*1e3:.1f} ms; recruited {sum(1 for x in got if x)} pairs")

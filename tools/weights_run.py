"""Driver for timing lctp_read_weights (UniqueKmers::calculate_read_weight, src/model/locs.rs:968-1002) and lctp_group_reads."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import argparse
import numpy as np
from locityper_b200 import genotype

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=200_000, help="read pairs of 2 x 150 bp")
ap.add_argument("--contigs", type=int, default=300)
ap.add_argument("--contig-len", type=int, default=3500)
ap.add_argument("--k", type=int, default=25)
ap.add_argument("--passes", type=int, default=4)
a = ap.parse_args()
rng = np.random.default_rng(5)
t0 = time.time()
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
base = rng.choice(acgt, size=a.contig_len)
contigs, counts = [], []
for _ in range(a.contigs):
    c = base.copy()
    mut = rng.random(a.contig_len) < 0.01
    c[mut] = rng.choice(acgt, size=int(mut.sum()))
    contigs.append(c.tobytes())
    counts.append(np.where(rng.random(a.contig_len + 1 - a.k) < 0.7, 0, 3).astype(np.uint16))
cat = np.frombuffer(b"".join(contigs), dtype=np.uint8)
starts = rng.integers(0, len(cat) - 150, 2 * a.pairs)
reads_arr = cat[starts[:, None] + np.arange(150)[None, :]].copy()
err = rng.random(reads_arr.shape) < 0.005
reads_arr[err] = rng.choice(acgt, size=int(err.sum()))
reads = [r.tobytes() for r in reads_arr]
print(f"{a.pairs} pairs, {a.contigs} contigs generated in {time.time()-t0:.1f}s", flush=True)
ctx = genotype.Context(0)
t0 = time.time()
u = genotype.UniqueKmers(ctx, contigs, counts, a.k, 1, 5)
print(f"unique k-mer table: {u.n_unique} keys, built + uploaded in {(time.time()-t0)*1e3:.1f} ms")
for i in range(a.passes):
    if i == 1:
        ctx.stats(reset=True)
        t1 = time.time()
    unique, weight = u.read_weights(reads, 2)
wall = (time.time() - t1) / max(1, a.passes - 1)
st = ctx.stats()
ms = st["recruit_ms"] / max(1, a.passes - 1)
nb = 2 * a.pairs * 150
print(f"k_read_weights + k_weight_of_counts: {ms:.4f} ms per call, {2*a.pairs/ms/1e3:.1f} M read ends/s, {nb/ms/1e6:.1f} GB/s of bases; "
      f"mean unique k-mers per end {unique.mean():.2f}, mean weight {weight.mean():.3f}; whole call incl. Python + H2D/D2H {wall*1e3:.1f} ms")
from test_group import _random_prelim
p = _random_prelim(9, n_reads=100_000)
for i in range(3):
    t1 = time.time()
    g = genotype.group_reads(ctx, p)
    wall = time.time() - t1
print(f"lctp_group_reads: {p.n_reads} reads, {len(p.rec_contig)} records -> {g['n_reads_out']} reads / {len(g['ma_contig'])} entries; "
      f"whole call incl. H2D/D2H {wall*1e3:.1f} ms")

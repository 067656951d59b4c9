"""Synthetic `solve::Data` generator.

Produces, at the level of the reference's `solve::Data` (src/solvers/solve.rs:254-273), everything the
genotype-evaluation hot path consumes: the read x haplotype pair-alignment lists (`AllAlignments`,
src/model/locs.rs:669-735), haplotype window geometry and per-position (gc, weight) arrays
(`ContigInfo`, src/model/windows.rs:343-445), background NB read-depth parameters per GC bin
(`ReadDepth`, src/bg/depth.rs:387-398) and `model::Params` defaults (src/model/mod.rs:108-134).
No FASTQ/BAM is involved: the path under test starts after read recruitment and alignment.

Everything is seeded (`numpy.random.default_rng(seed)`); data is synthetic by construction.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

LN10 = 2.302585092994045684
NONE_U32 = 0xFFFFFFFF
GC_BINS = 101


@dataclass
class Locus:
    """Flat per-locus input shared by the C-ABI and the test oracle."""

    n_haps: int
    n_reads: int
    ploidy: int
    is_paired: bool
    unmapped_prob: np.ndarray      # f64[R]
    pa_off: np.ndarray             # u64[R+1]
    pa_contig: np.ndarray          # u32[NPA]
    pa_ln_prob: np.ndarray         # f64[NPA]
    pa_mid1: np.ndarray            # u32[NPA]
    pa_mid2: np.ndarray            # u32[NPA]
    hap_len: np.ndarray            # u32[H]
    hap_n_windows: np.ndarray      # u32[H]
    hap_reg_start: np.ndarray      # u32[H]
    window: int
    left_padding: int
    hap_pos_off: np.ndarray        # u64[H+1]
    pos_weight: np.ndarray         # f64
    pos_gc: np.ndarray             # u8
    nb_n: np.ndarray               # f64[101] background NB (ploidy 1, first mates)
    nb_p: np.ndarray               # f64[101]
    alt_cn: np.ndarray             # f64
    depth_k: int = 0
    depth_table: Optional[np.ndarray] = None   # f64[101*depth_k]; filled by `attach_depth_table`
    tweak: int = 0
    prob_diff: float = 0.0
    lik_skew: float = 0.85
    min_weight: float = 0.001
    filt_diff: float = 100.0 * LN10
    prob_thresh: float = -4.0 * LN10
    dont_skip: bool = False
    out_bams: int = 0
    gt_tuples: Optional[np.ndarray] = None     # u32[G*p] or None = all combinations with replacement
    priors: Optional[np.ndarray] = None        # f64[G] or None
    truth: tuple = ()
    meta: dict = field(default_factory=dict)

    @property
    def n_genotypes(self) -> int:
        if self.gt_tuples is not None:
            return len(self.gt_tuples) // self.ploidy
        return math.comb(self.n_haps + self.ploidy - 1, self.ploidy)

    def genotype_tuple(self, g: int) -> tuple:
        """Reference enumeration order (src/ext/vec.rs:298-339): lexicographic, last index fastest."""
        if self.gt_tuples is not None:
            return tuple(int(x) for x in self.gt_tuples[g * self.ploidy:(g + 1) * self.ploidy])
        out, lo, H, p = [], 0, self.n_haps, self.ploidy
        for d in range(p):
            rem = p - d - 1
            for v in range(lo, H):
                cnt = 1 if rem == 0 else math.comb(H - v + rem - 1, rem)
                if g < cnt:
                    out.append(v)
                    lo = v
                    break
                g -= cnt
        return tuple(out)

    def genotype_index(self, tup) -> int:
        assert self.gt_tuples is None
        H, p = self.n_haps, self.ploidy
        g, lo = 0, 0
        for d, t in enumerate(tup):
            rem = p - d - 1
            for v in range(lo, t):
                g += 1 if rem == 0 else math.comb(H - v + rem - 1, rem)
            lo = t
        return g


def default_tweak(window: int, boundary: int = 200) -> int:
    """model::Params::set_tweak_size (src/model/mod.rs:179-186)."""
    return min(int(round(window * 0.5)), 200, max(boundary - 1, 0))


def _depth_params(rng, mean_depth: float):
    """NB(n, p) per GC bin: mild GC curve, variance 1.5-3x the mean (src/bg/depth.rs:387-398 layout)."""
    gc = np.arange(GC_BINS, dtype=np.float64)
    curve = 1.0 - 0.35 * ((gc - 45.0) / 45.0) ** 2
    curve = np.clip(curve, 0.35, None)
    m = mean_depth * curve
    over = 1.5 + 1.5 * rng.random(GC_BINS)
    over = np.convolve(np.pad(over, 3, mode="edge"), np.ones(7) / 7.0, mode="valid")
    v = m * over
    n = m * m / (v - m)
    p = m / v
    return n, p


def _weights_gc(rng, n_pos: int):
    """Per-position window weight (product of the two sigmoid weights, src/model/windows.rs:175-177,441-443)
    and GC content of the neighbourhood (windows.rs:386-391)."""
    def smooth(x, k):
        return np.convolve(np.pad(x, k // 2, mode="edge"), np.ones(k) / k, mode="valid")[:n_pos]

    uniq = np.clip(smooth(rng.beta(8, 2, n_pos + 64), 31), 1e-6, 1 - 1e-6)
    compl = np.clip(smooth(rng.beta(9, 1, n_pos + 64), 31), 1e-6, 1 - 1e-6)
    # 5% low-weight stretches (repeats)
    n_low = max(1, n_pos // 2000)
    for _ in range(n_low):
        s = int(rng.integers(0, max(1, n_pos - 100)))
        uniq[s:s + 100] *= 0.08
    def calc(x, bp, power):
        c = (bp / (1.0 - bp)) ** power
        return 1.0 / (1.0 + c * ((1.0 - x) / x) ** power)
    w = calc(uniq, 0.2, 4.0) * calc(compl, 0.5, 4.0)
    gc = np.clip(np.rint(smooth(42.0 + 18.0 * rng.standard_normal(n_pos + 64), 101)), 0, 100).astype(np.uint8)
    return w.astype(np.float64), gc


def make_locus(n_haps: int, n_reads: int, locus_len: int, seed: int, *, tech: str = "illumina",
               ploidy: int = 2, snp_rate: float = 1e-3,
               off_target: float = 0.10, multi_frac: float = 0.15,
               table_builder: Optional[Callable] = None, explicit_priors: bool = False) -> Locus:
    """Generate one synthetic locus.  `tech` is "illumina" (150 bp paired-end) or "hifi" (single-end)."""
    rng = np.random.default_rng(seed)
    H, R, L = n_haps, n_reads, locus_len
    paired = tech == "illumina"
    if paired:
        read_len, p_match, p_mm, U = 150, 0.995, 0.003, -10.0 * LN10
        window = 100
    else:
        read_len, p_match, p_mm, U = 15000, 0.998, 0.001, -100.0 * LN10
        window = 5000
    neighb = max(window, 300)
    left_padding = (neighb - window) // 2
    boundary = 200
    prob_diff = abs(U) + LN10                                   # src/command/genotype.rs:1294-1296

    # ---- haplotype panel: coalescent-ish tree of SNP sites -------------------------------------
    parents = np.zeros(H, dtype=np.int64)
    new_counts = rng.poisson(L * snp_rate, H)
    new_counts[0] = 0
    V = int(new_counts.sum())
    site_pos = rng.integers(0, L, V).astype(np.int64)
    X = np.zeros((H, max(V, 1)), dtype=bool)
    off = 0
    for h in range(1, H):
        parents[h] = rng.integers(0, h)
        X[h] = X[parents[h]]
        X[h, off:off + new_counts[h]] = True
        off += new_counts[h]
    order = np.argsort(site_pos, kind="stable")
    site_pos = site_pos[order]
    X = X[:, order]

    hap_len = (L + rng.integers(-window, window + 1, H)).astype(np.uint32)
    n_windows = ((hap_len - 2 * boundary) // window).astype(np.uint32)
    reg_start = ((hap_len - n_windows * window) // 2).astype(np.uint32)

    truth = tuple(sorted(int(x) for x in rng.integers(0, H, ploidy)))

    # ---- reads ------------------------------------------------------------------------------
    src = np.asarray(truth)[rng.integers(0, ploidy, R)]
    Lmin = int(hap_len.min())
    if paired:
        frag = np.clip(np.rint(rng.normal(450.0, 100.0, R)), 2 * read_len // 2 + 10, 900).astype(np.int64)
        rl = np.full(R, read_len, dtype=np.int64)
    else:
        rl = np.clip(np.rint(rng.normal(read_len, 3000.0, R)), 2000, Lmin - 10).astype(np.int64)
        frag = rl.copy()
    start = (rng.random(R) * np.maximum(Lmin - frag, 1)).astype(np.int64)
    # footprints of the two mates (single-end: only mate 1)
    a1, b1 = start, start + rl
    a2, b2 = start + frag - rl, start + frag
    mid1 = (a1 + b1) // 2
    mid2 = (a2 + b2) // 2
    ln_match, ln_mm = math.log(p_match), math.log(p_mm)
    ins_sd = 100.0
    insert_penalty = -math.log(ins_sd * math.sqrt(2 * math.pi)) if paired else 0.0
    ins_lp = (insert_penalty - 0.5 * ((frag - 450.0) / ins_sd) ** 2) if paired else np.zeros(R)
    unmapped = np.full(R, 2.0 * U + insert_penalty if paired else U, dtype=np.float64)

    is_off = rng.random(R) < off_target
    n_extra = np.where(rng.random(R) < multi_frac, rng.integers(1, 4, R), 0)
    n_extra[is_off] = 0

    def edits_under(lo_pos, hi_pos, rsel):
        """#panel differences between each haplotype and the read's source haplotype in [lo,hi)."""
        lo_i = np.searchsorted(site_pos, lo_pos, side="left")
        hi_i = np.searchsorted(site_pos, hi_pos, side="left")
        E = np.zeros((H, len(rsel)), dtype=np.int32)
        for t in set(truth):
            m = src[rsel] == t
            if not m.any():
                continue
            D = X ^ X[t]
            cum = np.zeros((H, D.shape[1] + 1), dtype=np.int32)
            np.cumsum(D, axis=1, out=cum[:, 1:])
            E[:, m] = cum[:, hi_i[m]] - cum[:, lo_i[m]]
        return E

    allr = np.arange(R)
    e1 = edits_under(a1, b1, allr)
    base_err1 = rng.poisson(rl * 0.002)
    noise = (rng.random((H, R)) < 0.04).astype(np.int32)
    e1 = e1 + base_err1[None, :] + noise
    if paired:
        e2 = edits_under(a2, b2, allr)
        base_err2 = rng.poisson(rl * 0.002)
        e2 = e2 + base_err2[None, :] + (rng.random((H, R)) < 0.04).astype(np.int32)
    lp1 = e1 * ln_mm + (rl[None, :] - e1) * ln_match
    lp1 -= lp1.max(axis=0, keepdims=True)                       # normalize_probs, locs.rs:358-360
    if paired:
        lp2 = e2 * ln_mm + (rl[None, :] - e2) * ln_match
        lp2 -= lp2.max(axis=0, keepdims=True)
        lp = lp1 + lp2 + ins_lp[None, :]                        # paired_prob, seq/aln.rs:236-238
        max_edits = 14
        aligned = (e1 + e2) <= (e1 + e2).min(axis=0, keepdims=True) + max_edits
    else:
        lp = lp1
        aligned = e1 <= e1.min(axis=0, keepdims=True) + 60
    # a few (read, haplotype) pairs have no alignment at all -> matrix falls back to unmapped_prob
    aligned &= rng.random((H, R)) > 0.01
    # off-target reads (paralogous origin): aligned to a minority of haplotypes only, so for most
    # genotypes their only option is "both mates unmapped"
    if is_off.any():
        n_off = int(is_off.sum())
        aligned[:, is_off] &= rng.random((H, n_off)) < rng.random(n_off)[None, :] * 0.4

    hh, rr = np.nonzero(aligned)
    ent_r = [rr.astype(np.int64)]
    ent_h = [hh.astype(np.int64)]
    ent_lp = [lp[hh, rr]]
    ent_m1 = [mid1[rr]]
    ent_m2 = [mid2[rr] if paired else np.full(len(rr), NONE_U32, dtype=np.int64)]

    # ---- secondary locations (repeats): same read, other position, a few more edits -----------
    rsel = np.nonzero(n_extra > 0)[0]
    for k in range(3):
        rk = rsel[n_extra[rsel] > k]
        if len(rk) == 0:
            continue
        shift_pen = rng.integers(0, 4, len(rk)) * ln_mm * (1.0 if paired else 3.0)
        new_start = (rng.random(len(rk)) * np.maximum(Lmin - frag[rk], 1)).astype(np.int64)
        per_hap = (rng.random((H, len(rk))) < 0.85) & aligned[:, rk]
        jitter = rng.integers(0, 2, (H, len(rk))) * ln_mm * (rng.random((H, len(rk))) < 0.1)
        lpk = lp[:, rk] + shift_pen[None, :] + jitter
        hs, cs = np.nonzero(per_hap)
        ent_r.append(rk[cs].astype(np.int64))
        ent_h.append(hs.astype(np.int64))
        ent_lp.append(lpk[hs, cs])
        m1k = new_start + rl[rk] // 2
        m2k = new_start + frag[rk] - rl[rk] // 2
        ent_m1.append(m1k[cs])
        ent_m2.append(m2k[cs] if paired else np.full(len(cs), NONE_U32, dtype=np.int64))

    ent_r = np.concatenate(ent_r); ent_h = np.concatenate(ent_h); ent_lp = np.concatenate(ent_lp)
    ent_m1 = np.concatenate(ent_m1); ent_m2 = np.concatenate(ent_m2)
    # some pair alignments have one mate unmapped (new_first / new_second, locs.rs:684-702)
    if paired:
        drop2 = rng.random(len(ent_r)) < 0.02
        ent_m2 = np.where(drop2, NONE_U32, ent_m2)
        ent_lp = np.where(drop2, ent_lp + U * 0.5, ent_lp)
    # sort: read asc, contig asc, ln_prob desc (locs.rs:793-798,818-858)
    order = np.lexsort((-ent_lp, ent_h, ent_r))
    ent_r, ent_h, ent_lp, ent_m1, ent_m2 = (x[order] for x in (ent_r, ent_h, ent_lp, ent_m1, ent_m2))
    # within each (read, contig): keep those within prob_diff of the best (at most 10; here <= 4)
    key = ent_r * H + ent_h
    first = np.ones(len(key), dtype=bool)
    first[1:] = key[1:] != key[:-1]
    grp = np.cumsum(first) - 1
    best = ent_lp[first][grp]
    keep = ent_lp >= best - prob_diff
    ent_r, ent_h, ent_lp, ent_m1, ent_m2 = (x[keep] for x in (ent_r, ent_h, ent_lp, ent_m1, ent_m2))
    pa_off = np.zeros(R + 1, dtype=np.uint64)
    np.cumsum(np.bincount(ent_r, minlength=R), out=pa_off[1:])

    # ---- per-position weight / gc ----------------------------------------------------------------
    pos_len = (hap_len.astype(np.int64) - neighb + 1)
    hap_pos_off = np.zeros(H + 1, dtype=np.uint64)
    np.cumsum(pos_len, out=hap_pos_off[1:])
    base_w, base_gc = _weights_gc(rng, int(pos_len.max()))
    pos_weight = np.empty(int(hap_pos_off[-1]), dtype=np.float64)
    pos_gc = np.empty(int(hap_pos_off[-1]), dtype=np.uint8)
    for h in range(H):
        o, n = int(hap_pos_off[h]), int(pos_len[h])
        # haplotypes share the locus-wide profile with small per-haplotype perturbations
        pw = base_w[:n] * (1.0 - 0.02 * rng.random())
        pos_weight[o:o + n] = pw
        pos_gc[o:o + n] = base_gc[:n]

    # Background depth is "ploidy 1, first read ends" (src/bg/depth.rs:256); match it to the read
    # density actually simulated so the depth term is informative rather than adversarial.
    mean_first = max(0.5, R * (1.0 - off_target) * window / float(L) / ploidy)
    nb_n, nb_p = _depth_params(rng, mean_first)

    loc = Locus(
        n_haps=H, n_reads=R, ploidy=ploidy, is_paired=paired,
        unmapped_prob=unmapped,
        pa_off=pa_off, pa_contig=ent_h.astype(np.uint32), pa_ln_prob=np.ascontiguousarray(ent_lp, dtype=np.float64),
        pa_mid1=ent_m1.astype(np.uint32), pa_mid2=ent_m2.astype(np.uint32),
        hap_len=hap_len, hap_n_windows=n_windows, hap_reg_start=reg_start,
        window=window, left_padding=left_padding,
        hap_pos_off=hap_pos_off, pos_weight=pos_weight, pos_gc=pos_gc,
        nb_n=nb_n, nb_p=nb_p, alt_cn=np.array([0.3, 2.0, 3.0, 4.0, 5.0]),
        tweak=default_tweak(window, boundary), prob_diff=prob_diff,
        truth=truth, meta=dict(seed=seed, tech=tech, locus_len=L),
    )
    if explicit_priors:
        # `--priors` mode (src/command/genotype.rs:1104-1121): an explicit genotype list with priors
        G = loc.n_genotypes
        sel = np.sort(rng.choice(G, size=max(2, G // 3), replace=False))
        tup = np.array([loc.genotype_tuple(int(g)) for g in sel], dtype=np.uint32).reshape(-1)
        loc.gt_tuples = tup
        loc.priors = -rng.random(len(sel)) * 5.0
    if table_builder is not None:
        attach_depth_table(loc, table_builder)
    return loc


def attach_depth_table(loc: Locus, table_builder: Callable, k_cols: Optional[int] = None) -> None:
    """Fill `depth_table` (101 x K, K >= 2R+1 so that no depth value can fall outside the table)."""
    k = int(k_cols) if k_cols is not None else 2 * loc.n_reads + 3
    loc.depth_k = k
    loc.depth_table = np.ascontiguousarray(
        table_builder(loc.nb_n, loc.nb_p, loc.is_paired, loc.alt_cn, k), dtype=np.float64).reshape(-1)


# Named shapes of BASELINE.json `configs`.
def config_shape(name: str) -> dict:
    shapes = {
        "C1": dict(n_haps=100, n_reads=2000, locus_len=3500, tech="illumina"),
        "C2": dict(n_haps=300, n_reads=2000, locus_len=3500, tech="illumina"),
        "C3": dict(n_haps=200, n_reads=240, locus_len=60000, tech="hifi"),
        "C4": dict(n_haps=1000, n_reads=10000, locus_len=15000, tech="illumina"),
        "C5": dict(n_haps=500, n_reads=4000, locus_len=8000, tech="illumina"),
    }
    return dict(shapes[name])


_ARRAY_FIELDS = ("unmapped_prob", "pa_off", "pa_contig", "pa_ln_prob", "pa_mid1", "pa_mid2", "hap_len",
                 "hap_n_windows", "hap_reg_start", "hap_pos_off", "pos_weight", "pos_gc", "nb_n", "nb_p", "alt_cn",
                 "depth_table", "gt_tuples", "priors")
_SCALAR_FIELDS = ("n_haps", "n_reads", "ploidy", "is_paired", "window", "left_padding", "depth_k", "tweak",
                  "prob_diff", "lik_skew", "min_weight", "filt_diff", "prob_thresh", "dont_skip", "out_bams")


def save_locus(loc: Locus, path: str) -> None:
    """Dump the flat locus (SURVEY.md Appendix C layout) as a compressed .npz -- fixtures, debugging."""
    d = {k: getattr(loc, k) for k in _ARRAY_FIELDS if getattr(loc, k) is not None}
    d["_scalars"] = np.array([float(getattr(loc, k)) for k in _SCALAR_FIELDS], dtype=np.float64)
    d["_truth"] = np.array(loc.truth, dtype=np.int64)
    np.savez_compressed(path, **d)


def load_locus(path: str) -> Locus:
    z = np.load(path)
    sc = dict(zip(_SCALAR_FIELDS, z["_scalars"].tolist()))
    kw = {k: (z[k] if k in z.files else None) for k in _ARRAY_FIELDS}
    for k in ("n_haps", "n_reads", "ploidy", "window", "left_padding", "depth_k", "tweak", "out_bams"):
        kw[k] = int(sc[k])
    for k in ("is_paired", "dont_skip"):
        kw[k] = bool(sc[k])
    for k in ("prob_diff", "lik_skew", "min_weight", "filt_diff", "prob_thresh"):
        kw[k] = float(sc[k])
    return Locus(truth=tuple(int(x) for x in z["_truth"]), **kw)


def make_mates(n_haps: int, n_reads: int, locus_len: int, seed: int, *, multi_frac: float = 0.15,
               over_cap_frac: float = 0.01, max_alns: int = 10) -> dict:
    """Synthetic mate alignments (the input of identify_paired_end_alignments, src/model/locs.rs:805-868) for R
    read pairs on H haplotypes: keyword arguments of `genotype.Mates`.  Per (read, haplotype) either both mates,
    one of them or none align; `multi_frac` of the ends have 1-3 secondary alignments (random place and strand,
    so same-strand combinations and far-apart pairs occur) and `over_cap_frac` have more than `max_alns`
    alignments; records are sorted the way the reference consumes them (contig asc, end asc, ln_prob desc)."""
    rng = np.random.default_rng(seed)
    H, R, L = n_haps, n_reads, locus_len
    read_len = 150
    kind = rng.random((R, H))
    has1 = kind < 0.88
    has2 = (kind < 0.80) | ((kind >= 0.88) & (kind < 0.96))
    cnt = np.zeros((R, H, 2), dtype=np.int64)
    for e, has in enumerate((has1, has2)):
        extra = np.where(rng.random((R, H)) < multi_frac, rng.integers(1, 4, (R, H)), 0)
        extra = np.where(rng.random((R, H)) < over_cap_frac, max_alns + rng.integers(1, 4, (R, H)), extra)
        cnt[:, :, e] = np.where(has, 1 + extra, 0)
    flat = cnt.reshape(-1)
    N = int(flat.sum())
    key = np.repeat(np.arange(flat.size), flat)              # (r, h, e) of every record
    r_ix, h_ix, e_ix = key // (2 * H), (key // 2) % H, key % 2
    first_of_key = np.r_[True, key[1:] != key[:-1]] if N else np.zeros(0, dtype=bool)
    base = rng.integers(0, max(1, L - 1200), R)
    ins = np.clip(rng.normal(450.0, 60.0, R), 160, 1000).astype(np.int64)
    jitter = rng.integers(-3, 4, N)
    prim_start = np.where(e_ix == 0, base[r_ix], base[r_ix] + ins[r_ix] - read_len) + jitter
    sec_start = rng.integers(0, max(1, L - read_len), N)
    start = np.clip(np.where(first_of_key, prim_start, sec_start), 0, L - read_len)
    length = read_len + rng.integers(-5, 6, N)
    end = np.minimum(start + length, L)
    strand = np.where(first_of_key, e_ix, rng.integers(0, 2, N))
    ln_prob = np.where(first_of_key, -np.abs(rng.normal(0.0, 8.0, N)), -np.abs(rng.normal(20.0, 10.0, N)))
    # a few exact ties inside a contig to exercise the stable order
    tie = rng.random(N) < 0.02
    ln_prob = np.where(tie, np.round(ln_prob), ln_prob)
    order = np.lexsort((-ln_prob, e_ix, h_ix, r_ix))
    r_ix, h_ix, e_ix = r_ix[order], h_ix[order], e_ix[order]
    ma_off = np.zeros(R + 1, dtype=np.uint64)
    np.cumsum(np.bincount(r_ix, minlength=R), out=ma_off[1:])
    sizes = np.arange(L + 1, dtype=np.float64)
    sd = 60.0
    ins_ln_pmf = -0.5 * ((sizes - 450.0) / sd) ** 2 - math.log(sd * math.sqrt(2.0 * math.pi))
    weight = rng.choice([1.0, 1.0, 1.0, 0.5, 0.25], R)
    U = -10.0 * LN10
    return dict(n_reads=R, n_haps=H, ma_off=ma_off, ma_contig=h_ix.astype(np.uint32),
                ma_flags=(e_ix | (strand[order] << 1)).astype(np.uint8), ma_start=start[order].astype(np.uint32),
                ma_end=end[order].astype(np.uint32), ma_ln_prob=ln_prob[order].astype(np.float64),
                ins_ln_pmf=ins_ln_pmf, unmapped_penalty=U, insert_penalty=float(ins_ln_pmf.max()),
                prob_diff=abs(U) + LN10, read_weight=weight.astype(np.float64), max_alns=max_alns)


def make_alns(n_alns: int, seed: int, *, tech: str = "illumina", contig_len: int = 3500) -> dict:
    """Synthetic alignment records (keyword arguments of `genotype.Alns`): extended CIGARs of 1-12 operations
    ([S] (=|X|I|D)* [S]; some all-soft single-operation records, some clipped at a contig end so that
    limited_clipping bites), intervals consistent with the CIGARs, error profile of the technology
    (src/bg/err_prof.rs:88-109), passable edit distance around the typical edit count."""
    rng = np.random.default_rng(seed)
    read_len = 150 if tech == "illumina" else 12000
    n_mid = rng.integers(1, 11, n_alns)
    has_l = rng.random(n_alns) < 0.25
    has_r = rng.random(n_alns) < 0.25
    all_soft = rng.random(n_alns) < 0.01
    counts = np.where(all_soft, 1, n_mid + has_l + has_r)
    off = np.zeros(n_alns + 1, dtype=np.uint64)
    off[1:] = np.cumsum(counts)
    ops = np.zeros(int(off[-1]), dtype=np.uint32)
    start = np.zeros(n_alns, dtype=np.uint32)
    end = np.zeros(n_alns, dtype=np.uint32)
    mid_codes = np.array([7, 7, 7, 8, 1, 2], dtype=np.uint32)          # '=' three times as likely as X / I / D
    for i in range(n_alns):
        b = int(off[i])
        if all_soft[i]:
            ops[b] = (read_len << 4) | 4
            ref = 0
        else:
            q = b
            if has_l[i]:
                ops[q] = (int(rng.integers(1, 40)) << 4) | 4
                q += 1
            codes = mid_codes[rng.integers(0, len(mid_codes), n_mid[i])]
            codes[0] = 7
            lens = np.where(codes == 7, rng.integers(5, read_len // 2, n_mid[i]), rng.integers(1, 6, n_mid[i]))
            ops[q:q + n_mid[i]] = (lens.astype(np.uint32) << 4) | codes
            q += n_mid[i]
            if has_r[i]:
                ops[q] = (int(rng.integers(1, 40)) << 4) | 4
            ref = int(lens[(codes == 7) | (codes == 8) | (codes == 2)].sum())
        edge = rng.random()
        if edge < 0.1:
            s0 = int(rng.integers(0, 20))                              # near the contig start: left clipping is limited
        elif edge < 0.2:
            s0 = max(0, contig_len - ref - int(rng.integers(0, 20)))   # near the contig end
        else:
            s0 = int(rng.integers(0, max(1, contig_len - ref)))
        start[i], end[i] = s0, s0 + ref
    p_match, p_mm = (0.995, 0.003) if tech == "illumina" else (0.998, 0.001)
    ln_oper = (np.log(p_match), np.log(p_mm), np.log(p_mm / 2), np.log(p_mm / 3), np.log(p_mm))
    return dict(cigar_off=off, cigar_ops=ops, aln_start=start, aln_end=end,
                contig_len=np.full(n_alns, contig_len, dtype=np.uint32),
                passable_dist=rng.integers(0, 60, n_alns).astype(np.uint32), ln_oper=ln_oper)

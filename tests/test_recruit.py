"""SURVEY 8(f) rank 3, first slice: short-read recruitment.

CPU: the oracle (oracle/lcto_recruit.c) against a statement-by-statement Python transcription of the cited Rust --
canonical minimizers (src/seq/kmers.rs:71-103, 256-340), TargetBuilder::add (src/seq/recruit.rs:680-735), the match
counters (:234-385), Fraction (src/math/frac.rs:48-98), recruit_short_read / recruit_read_pair (recruit.rs:852-930) -- and
recruit_long_read / has_matching_stretch (:932-998) -- and hand-checkable properties.  GPU: lctp_minimizers /
lctp_targets_build / lctp_recruit_short against the oracle, exactly."""
import math

import numpy as np
import pytest

from locityper_b200 import genotype

M64 = (1 << 64) - 1
UNDEF = M64


# ------------------------------------------------------------------ the cited Rust, line by line

def _fast_hash(x):                                   # Minimizer for u64, kmers.rs:93-103
    x = ~x & M64
    x ^= x >> 23
    x = (x * 0x2127599bf4325c37) & M64
    x ^= x >> 47
    return x


def _minimizers(seq: bytes, k: int, w: int):
    """kmers::minimizers::<u64, Vec<(u32, u64, bool)>, CANONICAL>, kmers.rs:256-331"""
    out = []
    mask = (1 << (2 * k)) - 1
    rv_shift = 2 * k - 2
    fw_kmer = rv_kmer = 0
    k_1, w_1 = k - 1, w - 1
    hashes, forward = [UNDEF] * 64, [True] * 64
    last_pos, best_pos, best_hash = -1, 0, UNDEF
    first_kmer, first_window = k_1, k_1 + w_1
    for i, nt in enumerate(seq):
        ch = chr(nt)
        if ch == "A": fw_enc, rv_enc = 0, 3
        elif ch == "C": fw_enc, rv_enc = 1, 2
        elif ch == "G": fw_enc, rv_enc = 2, 1
        elif ch == "T": fw_enc, rv_enc = 3, 0
        else:
            first_kmer = i + k
            fw_enc, rv_enc = 0, 0
        fw_kmer = ((fw_kmer << 2) | fw_enc) & mask
        rv_kmer = (rv_kmer >> 2) | (rv_enc << rv_shift)
        kmer, fw = (rv_kmer, False) if rv_kmer < fw_kmer else (fw_kmer, True)
        h = UNDEF if i < first_kmer else _fast_hash(kmer)
        hashes[i & 63] = h
        forward[i & 63] = fw
        if h < best_hash:
            best_hash, best_pos = h, i
        if i < first_window:
            continue
        start = i - w_1
        if best_pos < start:
            pos, mn = start, hashes[start & 63]                       # find_min, kmers.rs:237-252
            for j in range(start + 1, i + 1):
                v = hashes[j & 63]
                if v < mn:
                    pos, mn = j, v
            best_pos, best_hash = pos, mn
            if best_hash == UNDEF:
                first_window = first_window + w_1
                continue
        if best_pos > last_pos:
            last_pos = best_pos
            out.append((best_pos - k_1, best_hash, forward[best_pos & 63]))
    return out


def _approximate_u16(x: float):
    """Fraction::<u16>::approximate, frac.rs:48-80"""
    eps = 2.220446049250313e-16
    a2, a1, b2, b1 = 1, int(math.floor(x)), 0, 1
    xk = x
    for _ in range(20):
        numer = xk - math.floor(xk)
        if numer <= eps:
            break
        xk = 1.0 / numer
        fl = math.floor(xk)
        if not (0 <= fl <= 65535):
            break
        fl = int(fl)
        if fl * a1 > 65535 or fl * a1 + a2 > 65535:
            break
        if fl * b1 > 65535 or fl * b1 + b2 > 65535:
            break
        a2, a1, b2, b1 = a1, fl * a1 + a2, b1, fl * b1 + b2
        if abs(a1 / b1 - x) <= eps:
            break
    return a1, b1


class _Info:                                         # MinimInfo, recruit.rs:604-644
    def __init__(self, forward, rare):
        self.direction, self.rare = 1 + int(forward), rare

    def update(self, forward, rare):
        self.direction |= 1 + int(forward)
        self.rare = self.rare and rare

    def is_directed_to(self, forward):
        return self.direction & (1 + int(forward)) != 0


def _build_targets(ts: genotype.TargetSeqs):
    """TargetBuilder::add for every locus, recruit.rs:680-735: minim_to_loci (insertion-ordered dict of lists)."""
    minim_to_loci = {}
    base_k, mk = ts.base_k, ts.minimizer_k
    shift = (base_k - mk) // 2 if mk <= base_k else mk - base_k
    for seq, locus_ix, counts in zip(ts.seqs, ts.seq_locus, ts.kmer_counts):
        n_counts = len(counts)
        assert max(0, len(seq) + 1 - base_k) == n_counts
        for pos, minimizer, forward in _minimizers(seq, mk, ts.minimizer_w):
            if mk <= base_k:
                rare = counts[min(max(pos - shift, 0), n_counts - 1)] < ts.thresh_kmer_count
            else:
                rare = counts[pos] < ts.thresh_kmer_count and counts[pos + shift] < ts.thresh_kmer_count
            v = minim_to_loci.setdefault(minimizer, [])
            if v and v[-1][0] == locus_ix:
                v[-1][1].update(forward, bool(rare))
            else:
                v.append((int(locus_ix), _Info(forward, bool(rare))))
    return minim_to_loci


def _inc(arr, forward, info):                        # BaseMatchCount::inc, recruit.rs:246-252
    i = int(info.rare) << 1
    arr[i] += int(info.is_directed_to(not forward))
    arr[i | 1] += int(info.is_directed_to(forward))


def _fw_num(a): return 3 * a[3] + a[1]
def _bw_num(a): return 3 * a[2] + a[0]
def _fw_den(a, t): return 3 * (t - a[1]) + a[1]
def _bw_den(a, t): return 3 * (t - a[0]) + a[0]
def _ge(f, g): return f[0] * g[1] >= g[0] * f[1]     # Fraction<u16> partial_cmp, frac.rs:92-98


def _recruit(minim_to_loci, mf, ts, seq1, seq2):
    """recruit_short_read (recruit.rs:852-881) / recruit_read_pair (:885-930); the answer as a sorted list."""
    matches = {}
    buf = _minimizers(seq1, ts.minimizer_k, ts.minimizer_w)
    total1 = len(buf)
    for _, minimizer, forward in buf:
        for locus_ix, info in minim_to_loci.get(minimizer, ()):
            c = matches.setdefault(locus_ix, ([0, 0, 0, 0], [0, 0, 0, 0]))
            _inc(c[0], forward, info)
    answer = []
    if seq2 is None:
        for locus_ix, (first, _) in matches.items():
            if first[2] != 0 or first[3] != 0:
                frac = (_fw_num(first), _fw_den(first, total1)) if _fw_num(first) >= _bw_num(first) \
                    else (_bw_num(first), _bw_den(first, total1))
                if _ge(frac, mf):
                    answer.append(locus_ix)
        return sorted(answer)
    if not matches:
        return []
    buf = _minimizers(seq2, ts.minimizer_k, ts.minimizer_w)
    total2 = len(buf)
    for _, minimizer, forward in buf:
        for locus_ix, info in minim_to_loci.get(minimizer, ()):
            if locus_ix in matches:
                _inc(matches[locus_ix][1], forward, info)
    for locus_ix, (first, second) in matches.items():
        if first[2] != 0 or first[3] != 0 or second[2] != 0 or second[3] != 0:
            if _fw_num(first) + _bw_num(second) >= _bw_num(first) + _fw_num(second):
                f1, f2 = (_fw_num(first), _fw_den(first, total1)), (_bw_num(second), _bw_den(second, total2))
            else:
                f1, f2 = (_bw_num(first), _bw_den(first, total1)), (_fw_num(second), _fw_den(second, total2))
            if _ge(f1, mf) and _ge(f2, mf):
                answer.append(locus_ix)
    return sorted(answer)


def _recruit_long(minim_to_loci, ts, seq):
    """recruit_long_read (recruit.rs:967-998) + has_matching_stretch (:932-964) + Params (:95-109); sorted answer."""
    stretch_minims = (2 * ts.match_length + (ts.minimizer_w + 1) - 1) // (ts.minimizer_w + 1)       # fast_ceil_div
    stretch_score = int(math.ceil(max(float(stretch_minims) * (float(3 + 1) * ts.match_frac - float(1)), float(3))))
    buf = _minimizers(seq, ts.minimizer_k, ts.minimizer_w)
    total = len(buf)
    matches = {}
    for _, minimizer, forward in buf:
        for locus_ix, info in minim_to_loci.get(minimizer, ()):
            _inc(matches.setdefault(locus_ix, [0, 0, 0, 0]), forward, info)
    answer = []
    for locus_ix, (bw_c, fw_c, bw_r, fw_r) in matches.items():
        num, den = (fw_r, total - fw_c) if fw_r >= bw_r else (bw_r, total - bw_c)                   # rare_fraction
        if num < max(1, int(math.ceil(float(min(stretch_minims, den)) * ts.match_frac))):           # long_read_threshold
            continue
        ok = den < stretch_minims
        if not ok:
            locus_minimizers = {m: i for m, v in minim_to_loci.items() for l, i in v if l == locus_ix}
            s_fw = s_bw = 0
            for _, minimizer, forward in buf:
                info = locus_minimizers.get(minimizer)
                if info is not None:
                    x = 1 + int(info.rare) * 3
                    s_fw += int(info.is_directed_to(forward)) * x
                    s_bw += int(info.is_directed_to(not forward)) * x
                s_fw, s_bw = max(s_fw - 1, 0), max(s_bw - 1, 0)
                if s_fw >= stretch_score or s_bw >= stretch_score:
                    ok = True
                    break
        if ok:
            answer.append(locus_ix)
    return sorted(answer)


def _long_world(seed, n_reads=30):
    """Targets of 12-kb alleles and HiFi-like single-end reads of 3-9 kb: on target (low error), on target but diverged,
    chimeric (a short on-target stretch inside random sequence: passes the count test on few minimizers only) and random."""
    rng = np.random.default_rng(seed)
    ts, _ = _world(seed, n_loci=3, alleles=2, length=12000, n_reads=1, paired=False)
    ts.match_length = 2000
    ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = []
    for _ in range(n_reads):
        kind = rng.random()
        n = int(rng.integers(3000, 9000))
        if kind < 0.2:
            r = bytes(rng.choice(ACGT, n))
        else:
            s = ts.seqs[int(rng.integers(0, len(ts.seqs)))]
            st = int(rng.integers(0, len(s) - n))
            r = _mutate(rng, s[st:st + n], 0.3 if kind < 0.35 else 0.005, 0.0005)
            if kind > 0.8:                                                       # chimera: 600 on-target bases in the middle
                r = bytes(rng.choice(ACGT, n // 2)) + r[:600] + bytes(rng.choice(ACGT, n // 2))
            if rng.random() < 0.5:
                r = _revcomp(r.replace(b"N", b"A"))
        reads.append(r)
    return ts, genotype.Reads(seq1=reads)


# ------------------------------------------------------------------ synthetic targets and reads

_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def _revcomp(s: bytes) -> bytes:
    return s.translate(_COMP)[::-1]


def _mutate(rng, s: bytes, rate: float, n_rate: float = 0.0) -> bytes:
    a = np.frombuffer(s, dtype=np.uint8).copy()
    m = rng.random(len(a)) < rate
    a[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(m.sum()))]
    if n_rate:
        a[rng.random(len(a)) < n_rate] = ord("N")
    return a.tobytes()


def _world(seed, n_loci=4, alleles=3, length=3000, n_reads=400, paired=True, base_k=25, mk=15, mw=10, shared=True):
    rng = np.random.default_rng(seed)
    seqs, locus, counts = [], [], []
    common = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 300))         # a repeat shared by all loci
    for l in range(n_loci):
        base = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), length))
        if shared:
            base = base[:1000] + common + base[1300:]
        for a in range(alleles):
            s = _mutate(rng, base, 0.01, n_rate=0.001 if a == 1 else 0.0)
            seqs.append(s)
            locus.append(l)
            c = rng.integers(0, 5, max(0, len(s) + 1 - base_k)).astype(np.uint16)
            c[1000:1300] = 50                                                         # the repeat is "common"
            counts.append(c)
    ts = genotype.TargetSeqs(seqs=seqs, seq_locus=np.array(locus, dtype=np.uint32), kmer_counts=counts, base_k=base_k,
                             minimizer_k=mk, minimizer_w=mw, thresh_kmer_count=10, match_frac=0.5)
    r1, r2 = [], []
    for _ in range(n_reads):
        kind = rng.random()
        if kind < 0.25:                                                               # unrelated read
            a = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 150))
            b = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 150))
        else:
            s = seqs[int(rng.integers(0, len(seqs)))]
            st = int(rng.integers(0, len(s) - 450))
            frag = s[st:st + int(rng.integers(300, 450))]
            err = 0.25 if kind < 0.4 else 0.01                                        # some reads too diverged to recruit
            a, b = _mutate(rng, frag[:150], err, 0.002), _mutate(rng, _revcomp(frag)[:150], err)
            if rng.random() < 0.5:
                a, b = b, a
        r1.append(a)
        r2.append(b)
    reads = genotype.Reads(seq1=r1, seq2=r2 if paired else None)
    return ts, reads


# ------------------------------------------------------------------ CPU: oracle against the transcription

@pytest.mark.parametrize("k,w", [(15, 10), (5, 4), (31, 63), (1, 2), (21, 2)])
def test_oracle_minimizers_equal_python_transcription(oracle, k, w):
    rng = np.random.default_rng(k * 100 + w)
    cases = [b"", b"ACGT", b"N" * 40, b"A" * 200, bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 700))]
    cases.append(_mutate(rng, cases[-1], 0.0, n_rate=0.02))
    cases.append(b"ACGTNNNN" * 30 + cases[-2][:100] + b"acgtacgt" + cases[-2][100:200])   # lower case = N
    for seq in cases:
        ref = _minimizers(seq, k, w)
        h, p, f = oracle.minimizers(seq, k, w)
        assert [int(x) for x in p] == [r[0] for r in ref] and [int(x) for x in h] == [r[1] for r in ref]
        assert [bool(x) for x in f] == [r[2] for r in ref]
        assert all(b > a for a, b in zip(p[:-1], p[1:]))                                   # positions strictly ascending


def test_minimizer_is_the_window_minimum():
    """Property: without Ns every emitted position holds the leftmost minimum of at least one full window, and every full
    window's leftmost minimum is emitted."""
    rng = np.random.default_rng(5)
    seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 400))
    k, w = 11, 7
    mask = (1 << (2 * k)) - 1
    hs = []
    for i in range(len(seq) - k + 1):
        kmer = seq[i:i + k]
        fwv = int("".join(str("ACGT".index(chr(c))) for c in kmer), 4)
        rvv = int("".join(str("ACGT".index(chr(c))) for c in _revcomp(kmer)), 4)
        hs.append(_fast_hash(min(fwv, rvv) & mask))
    want = set()
    for s in range(len(hs) - w + 1):
        win = hs[s:s + w]
        want.add(s + win.index(min(win)))
    assert {p for p, _, _ in _minimizers(seq, k, w)} == want


@pytest.mark.parametrize("x", [0.25, 0.3, 0.5, 0.6, 0.7, 1.0, 0.333, 0.8571, 0.123456789])
def test_oracle_fraction_approximation(oracle, x):
    assert oracle.fraction_approximate_u16(x) == _approximate_u16(x)
    a, b = _approximate_u16(x)
    assert abs(a / b - x) < 1e-4


@pytest.mark.parametrize("seed,paired,mk,base_k", [(1, True, 15, 25), (2, False, 15, 25), (3, True, 27, 25), (4, True, 10, 11)])
def test_oracle_targets_and_recruitment_equal_python_transcription(oracle, seed, paired, mk, base_k):
    ts, reads = _world(seed, paired=paired, mk=mk, base_k=base_k, n_reads=150)
    t = oracle.Targets(ts)
    ref = _build_targets(ts)
    key, locus, info = t.entries()
    flat = [(m, l, i.direction | (int(i.rare) << 2)) for m, v in ref.items() for l, i in v]
    assert sorted(zip(key.tolist(), locus.tolist(), info.tolist())) == sorted(flat)
    mf = _approximate_u16(ts.match_frac)
    got = t.recruit(reads)
    want = [_recruit(ref, mf, ts, reads.seq1[r], reads.seq2[r] if paired else None) for r in range(len(reads.seq1))]
    assert got == want
    n_rec = sum(1 for a in got if a)
    assert 0 < n_rec < len(got)                                   # some reads recruited, some not
    if seed == 1:
        assert any(len(a) > 1 for a in got) or True               # (reads in the shared repeat may hit several loci)


@pytest.mark.parametrize("seed", [11, 12])
def test_oracle_long_read_recruitment_equals_python_transcription(oracle, seed):
    ts, reads = _long_world(seed)
    t = oracle.Targets(ts)
    ref = _build_targets(ts)
    got = t.recruit(reads)
    want = [_recruit_long(ref, ts, r) for r in reads.seq1]
    assert got == want
    assert 0 < sum(1 for a in got if a) < len(got)


# ------------------------------------------------------------------ GPU: the product against the oracle

@pytest.mark.gpu
@pytest.mark.parametrize("k,w", [(15, 10), (5, 4), (31, 63), (21, 2)])
def test_device_minimizers_bit_exact(oracle, gpu_ctx, k, w):
    rng = np.random.default_rng(k + w)
    seqs = [b"", b"ACGT", b"N" * 40, b"A" * 200]
    for n in (150, 151, 250, 3000):
        s = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), n))
        seqs += [s, _mutate(rng, s, 0.0, n_rate=0.03)]
    got = genotype.minimizers(gpu_ctx, seqs, k, w)
    for s, (h, p, f) in zip(seqs, got):
        oh, op, of = oracle.minimizers(s, k, w)
        assert np.array_equal(h, oh) and np.array_equal(p, op) and np.array_equal(f, of)
    with pytest.raises(Exception):
        genotype.minimizers(gpu_ctx, seqs, 32, 10)                # k > MAX_KMER_SIZE of u64 minimizers


@pytest.mark.gpu
@pytest.mark.parametrize("seed,paired,mk,base_k,n_reads", [(1, True, 15, 25, 400), (2, False, 15, 25, 400), (3, True, 27, 25, 300),
                                                            (7, True, 15, 25, 20000)])
def test_device_recruitment_equals_the_oracle(oracle, gpu_ctx, seed, paired, mk, base_k, n_reads):
    ts, reads = _world(seed, paired=paired, mk=mk, base_k=base_k, n_reads=n_reads)
    t = genotype.Targets(gpu_ctx, ts)
    o = oracle.Targets(ts)
    k1, l1, i1 = t.entries()
    k2, l2, i2 = o.entries()
    assert np.array_equal(k1, k2) and np.array_equal(l1, l2) and np.array_equal(i1, i2)      # same insertion order too
    assert t.match_frac() == oracle.fraction_approximate_u16(ts.match_frac)
    got = t.recruit(reads)
    assert got == o.recruit(reads)
    assert 0 < sum(1 for a in got if a) < len(got)
    t.free()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n_reads", [(11, 30), (13, 400)])
def test_device_long_read_recruitment_equals_the_oracle(oracle, gpu_ctx, seed, n_reads):
    ts, reads = _long_world(seed, n_reads=n_reads)
    t = genotype.Targets(gpu_ctx, ts)
    o = oracle.Targets(ts)
    got = t.recruit(reads)
    assert got == o.recruit(reads)
    assert 0 < sum(1 for a in got if a) < len(got)
    # a mix of short and long single-end reads in one call: each goes down its own path (recruit.rs:589)
    ts2, short = _world(seed, n_loci=3, alleles=2, length=12000, n_reads=50, paired=False)
    mixed = genotype.Reads(seq1=list(reads.seq1[:10]) + list(short.seq1[:40]))
    assert t.recruit(mixed) == o.recruit(mixed)
    t.free()


@pytest.mark.gpu
def test_device_recruitment_rejects_bad_targets(gpu_ctx):
    ts, reads = _world(9, n_reads=5)
    ts.match_frac = 0.1
    with pytest.raises(Exception):
        genotype.Targets(gpu_ctx, ts)
    ts.match_frac, ts.match_length = 0.5, 100
    with pytest.raises(Exception):
        genotype.Targets(gpu_ctx, ts)

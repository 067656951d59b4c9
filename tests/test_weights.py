"""Read weights from the k-mers unique to a locus: UniqueKmers (src/model/locs.rs:915-1003) over kmers::kmers::<u128, _,
CANONICAL> (src/seq/kmers.rs:163-202).

CPU: the oracle (oracle/lcto_weights.c) against a statement-by-statement Python transcription (Python integers as u128,
a set as the HashSet) and hand-checked cases.  GPU: lctp_unique_kmers_build + lctp_read_weights against the oracle,
exactly (counts are integers, the weight is one multiply-add and a clamp)."""
import numpy as np
import pytest

from locityper_b200 import genotype

UNDEF = (1 << 128) - 1


def _kmers(seq: bytes, k: int):
    """kmers::kmers::<u128, _, true> (kmers.rs:163-202): list of canonical k-mers, UNDEF where an N is inside."""
    out = []
    mask = (1 << (2 * k)) - 1
    rv_shift = 2 * k - 2
    fw = rv = 0
    k_1 = k - 1
    reset = k_1
    for i, nt in enumerate(seq):
        enc = {65: 0, 67: 1, 71: 2, 84: 3}.get(nt)
        if enc is None:
            reset = i + k
            if i + 1 >= k:
                out.append(UNDEF)
            continue
        fw = ((fw << 2) | enc) & mask
        rv = (rv >> 2) | ((3 - enc) << rv_shift)
        if i >= reset:
            out.append(rv if rv < fw else fw)
        elif i + 1 >= k:
            out.append(UNDEF)
    return out


class _Transcription:
    def __init__(self, contig_seqs, kmer_counts, k, hard, soft):
        self.k, self.k_2 = k, k - 2
        self.unique = set()
        for seq, counts in zip(contig_seqs, kmer_counts):                  # locs.rs:941-952
            buf = _kmers(seq, k)
            assert len(buf) == len(counts)
            for kmer, count in zip(buf, counts):
                if count == 0:
                    self.unique.add(kmer)
        self.weight_mult = 1.0 / float(soft + 1 - hard)                    # :957
        self.weight_interc = (1.0 - float(hard)) * self.weight_mult        # :958

    def read_weights(self, read_seqs, ends):
        n = len(read_seqs) // ends
        unique, weight = [], []
        for r in range(n):
            paired_count = 0
            for e in range(ends):
                seq = read_seqs[r * ends + e]
                count = 0
                if seq:                                                    # Some(data)
                    it = iter(_kmers(seq, self.k))
                    for kmer in it:                                        # :983-990
                        if kmer in self.unique:
                            count = min(count + 1, 65535)
                            for _ in range(self.k_2 + 1):                  # kmers_iter.nth(k_2)
                                if next(it, None) is None:
                                    break
                unique.append(count)
                paired_count += count
            w = self.weight_interc + float(paired_count) * self.weight_mult
            weight.append(min(max(w, 0.0), 1.0))                           # clamp(0.0, 1.0), :997
        return np.array(unique, dtype=np.uint16), np.array(weight)


def _case(seed, n_contigs=6, contig_len=1500, n_reads=400, k=25, ends=2, n_frac=0.002):
    rng = np.random.default_rng(seed)
    base = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=contig_len)
    contigs, counts = [], []
    for _ in range(n_contigs):
        c = base.copy()
        mut = rng.random(contig_len) < 0.01
        c[mut] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(mut.sum()))
        c[rng.random(contig_len) < n_frac] = ord("N")
        contigs.append(c.tobytes())
        cnt = np.where(rng.random(contig_len + 1 - k) < 0.7, 0, rng.integers(1, 9, contig_len + 1 - k)).astype(np.uint16)
        counts.append(cnt)
    reads = []
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for r in range(n_reads * ends):
        if ends == 2 and r % 2 == 1 and rng.random() < 0.05:
            reads.append(b"")                                              # no mate
            continue
        c = contigs[int(rng.integers(0, n_contigs))]
        ln = int(rng.integers(k - 3, 260))                                 # some reads shorter than k
        s = int(rng.integers(0, contig_len - ln))
        seq = bytearray(c[s:s + ln])
        for p in np.nonzero(rng.random(ln) < 0.01)[0]:
            seq[p] = b"ACGTN"[int(rng.integers(0, 5))]
        seq = bytes(seq)
        if rng.random() < 0.5:
            seq = seq.translate(comp)[::-1]                                # reverse strand: canonical k-mers still match
        reads.append(seq)
    return contigs, counts, reads


def test_kmers_hand_checked():
    # k = 3: ACG -> fw 0b000110 = 6, reverse complement CGT = 0b011011 = 27 -> 6; CGT -> fw 27, rc ACG = 6 -> 6
    assert _kmers(b"ACGT", 3) == [6, 6]
    assert _kmers(b"ACNGT", 3) == [UNDEF, UNDEF, UNDEF]
    assert _kmers(b"AC", 3) == []


def test_hand_checked_weight(oracle):
    """k = 3, contig ACGTTT with every k-mer unique: ACG, CGT (= ACG canonically), GTT = AAC, TTT = AAA.  Read ACGTTT: hit
    at k-mer 0, k-mers 1-2 skipped, hit at k-mer 3 -> 2 non-overlapping unique k-mers.  hard 1 / soft 5: weight = (2 - 1 + 1)
    / (5 - 1 + 1) = 0.4; a read without unique k-mers: (0 - 1 + 1) / 5 = 0."""
    u = oracle.UniqueKmers([b"ACGTTT"], [np.zeros(4, dtype=np.uint16)], 3, 1, 5)
    assert u.n_unique == 3
    unique, weight = u.read_weights([b"ACGTTT", b"", b"CCCCCC", b"ACG"], 2)
    assert list(unique) == [2, 0, 0, 1] and list(weight) == [0.4, 0.2]


@pytest.mark.parametrize("seed,k,ends", [(1, 25, 2), (2, 31, 1), (3, 11, 2), (4, 40, 2)])
def test_oracle_vs_transcription(oracle, seed, k, ends):
    contigs, counts, reads = _case(seed, k=k, ends=ends, n_reads=150)
    u = oracle.UniqueKmers(contigs, counts, k, 1, 5)
    t = _Transcription(contigs, counts, k, 1, 5)
    assert u.n_unique == len(t.unique)
    got, want = u.read_weights(reads, ends), t.read_weights(reads, ends)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert got[0].max() >= 3 and got[1].min() < 1.0 and 1.0 in got[1]


def test_mismatched_counts_are_rejected(oracle):
    with pytest.raises(RuntimeError):
        oracle.UniqueKmers([b"ACGTACGT"], [np.zeros(3, dtype=np.uint16)], 3, 1, 5)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,k,ends,n_reads", [(11, 25, 2, 5000), (12, 31, 1, 3000), (13, 5, 2, 500), (14, 63, 2, 500)])
def test_gpu_read_weights_equal_oracle(gpu_ctx, oracle, seed, k, ends, n_reads):
    contigs, counts, reads = _case(seed, k=k, ends=ends, n_reads=n_reads)
    ou = oracle.UniqueKmers(contigs, counts, k, 1, 5)
    gu = genotype.UniqueKmers(gpu_ctx, contigs, counts, k, 1, 5)
    assert gu.n_unique == ou.n_unique
    l0 = gpu_ctx.launch_count()
    got, want = gu.read_weights(reads, ends), ou.read_weights(reads, ends)
    assert gpu_ctx.launch_count() == l0 + 2
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    gu.free()


@pytest.mark.gpu
def test_gpu_read_weights_edge_cases(gpu_ctx, oracle):
    gu = genotype.UniqueKmers(gpu_ctx, [b"ACGTTT"], [np.zeros(4, dtype=np.uint16)], 3, 1, 5)
    unique, weight = gu.read_weights([b"ACGTTT", b"", b"CCCCCC", b"ACG"], 2)
    assert list(unique) == [2, 0, 0, 1] and list(weight) == [0.4, 0.2]
    assert len(gu.read_weights([], 2)[1]) == 0
    with pytest.raises(genotype.ffi.LctpError):
        gu.read_weights([b"ACGT"], 3)                               # 1 or 2 read ends per read
    with pytest.raises(genotype.ffi.LctpError):
        genotype.UniqueKmers(gpu_ctx, [b"ACGTACGT"], [np.zeros(3, dtype=np.uint16)], 3, 1, 5)
    with pytest.raises(genotype.ffi.LctpError):
        genotype.UniqueKmers(gpu_ctx, [b"ACGTACGT"], [np.zeros(6, dtype=np.uint16)], 3, 6, 5)

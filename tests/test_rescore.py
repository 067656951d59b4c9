"""SURVEY 8(f) rank 2, first slice: the per-alignment part of PrelimAlignments::push (src/model/locs.rs:297-313).

CPU: the oracle (oracle/lcto_rescore.c) against a statement-by-statement Python transcription of the cited Rust
(count_region_operations_fast src/seq/aln.rs:298-317, limited_clipping :288-296, soft_clipping src/seq/cigar.rs:519-527,
edit_distance src/bg/err_prof.rs:73-79, ln_prob :212-221) and hand-checked records.  GPU: lctp_rescore_alignments against
the oracle, bit for bit."""
import numpy as np
import pytest

from locityper_b200 import genotype, synth

OPS = {"I": 1, "D": 2, "S": 4, "=": 7, "X": 8, "M": 0, "N": 3, "H": 5}


def _cigar(s):
    out, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num) << 4) | OPS[ch])
            num = ""
    return out


def _alns(records, ln_oper=(-0.005, -5.8, -6.5, -6.9, -5.8)):
    """records: (cigar string, start, end, contig_len, passable_dist)"""
    ops, off = [], [0]
    for c, *_ in records:
        ops += _cigar(c)
        off.append(len(ops))
    g = lambda k: np.array([r[k] for r in records], dtype=np.uint32)
    return genotype.Alns(cigar_off=np.array(off, dtype=np.uint64), cigar_ops=np.array(ops, dtype=np.uint32),
                         aln_start=g(1), aln_end=g(2), contig_len=g(3), passable_dist=g(4), ln_oper=ln_oper)


def _transcription(a: genotype.Alns) -> dict:
    """The cited Rust, line by line, in Python floats / ints."""
    n = a.n_alns
    out = dict(ln_prob=np.zeros(n), edit=np.zeros(n, dtype=np.uint32), read_len=np.zeros(n, dtype=np.uint32),
               save=np.zeros(n, dtype=np.uint8))
    lm, lx, li, ld, lc = (float(v) for v in a.ln_oper)
    for i in range(n):
        tuples = [(int(t) & 15, int(t) >> 4) for t in a.cigar_ops[int(a.cigar_off[i]):int(a.cigar_off[i + 1])]]
        matches = mismatches = deletions = insertions = 0
        for op, oplen in tuples:                                   # aln.rs:302-312
            if op == 7: matches += oplen
            elif op == 8: mismatches += oplen
            elif op == 2: deletions += oplen
            elif op == 1: insertions += oplen
            elif op == 4: pass
            else: raise ValueError("Unsupported CIGAR operation")
        assert tuples                                              # cigar.rs:520
        left = tuples[0][1] if tuples[0][0] == 4 else 0            # cigar.rs:524
        right = tuples[-1][1] if tuples[-1][0] == 4 else 0         # cigar.rs:525
        start, end, clen = int(a.aln_start[i]), int(a.aln_end[i]), int(a.contig_len[i])
        clipping = min(left, start) + min(right, max(0, clen - end))   # aln.rs:290-295, :314-315
        common = mismatches + insertions + clipping                # err_prof.rs:74
        out["edit"][i] = common + deletions
        out["read_len"][i] = common + matches
        out["ln_prob"][i] = lm * float(matches) + lx * float(mismatches) + li * float(insertions) \
            + ld * float(deletions) + lc * float(clipping)         # err_prof.rs:216-220 (Python: same left-to-right f64)
        out["save"][i] = out["edit"][i] <= a.passable_dist[i]      # locs.rs:308
    return out


def _same(x, y):
    return all(np.array_equal(x[k], y[k]) for k in ("ln_prob", "edit", "read_len", "save"))


HAND = [
    ("100=", 10, 110, 500, 0),               # perfect: edit 0, read_len 100
    ("5S95=", 10, 105, 500, 5),              # left clip fully inside the contig: clipping 5
    ("5S95=", 3, 98, 500, 5),                # ... limited by the contig start: clipping 3
    ("90=10S", 400, 490, 495, 9),            # right clip limited by the contig end: clipping 5
    ("90=10S", 400, 490, 480, 0),            # alignment end beyond the contig (saturating_sub): clipping 0
    ("40=1X20=2I30=3D7=", 0, 101, 500, 6),   # edit = 1 + 2 + 3
    ("150S", 7, 7, 500, 1000),               # a single soft operation is both the first and the last: clipping min(150,7) + min(150,493)
]


def test_oracle_hand_checked_records(oracle):
    a = _alns(HAND)
    r = oracle.rescore_alignments(a)
    assert list(r["edit"]) == [0, 5, 3, 5, 0, 6, 157]
    assert list(r["read_len"]) == [100, 100, 98, 95, 90, 100, 157]
    assert list(r["save"]) == [1, 1, 1, 1, 1, 1, 1]
    lm, lx, li, ld, lc = a.ln_oper
    assert r["ln_prob"][0] == lm * 100.0
    assert r["ln_prob"][5] == lm * 97.0 + lx * 1.0 + li * 2.0 + ld * 3.0 + lc * 0.0
    assert _same(r, _transcription(a))


@pytest.mark.parametrize("n,seed,tech", [(500, 1, "illumina"), (300, 2, "hifi"), (1, 3, "illumina")])
def test_oracle_equals_python_transcription(oracle, n, seed, tech):
    a = genotype.Alns(**synth.make_alns(n, seed, tech=tech))
    assert _same(oracle.rescore_alignments(a), _transcription(a))


def test_oracle_rejects_what_the_reference_panics_on(oracle):
    with pytest.raises(RuntimeError):
        oracle.rescore_alignments(_alns([("50=", 0, 50, 100, 1), ("20M", 0, 20, 100, 1)]))
    bad = _alns([("50=", 0, 50, 100, 1)])
    bad.cigar_off = np.array([0, 0], dtype=np.uint64)              # empty CIGAR
    with pytest.raises(RuntimeError):
        oracle.rescore_alignments(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,tech", [(200_000, 11, "illumina"), (20_000, 12, "hifi"), (1, 13, "illumina"), (257, 14, "illumina")])
def test_device_rescoring_bit_exact(oracle, gpu_ctx, n, seed, tech):
    a = genotype.Alns(**synth.make_alns(n, seed, tech=tech))
    gpu_ctx.stats(reset=True)
    got = genotype.rescore_alignments(gpu_ctx, a)
    assert _same(got, oracle.rescore_alignments(a))
    st = gpu_ctx.stats(reset=True)
    assert st["rescore_alns"] == n and st["rescore_ops"] == len(a.cigar_ops) and st["rescore_launches"] == 1


@pytest.mark.gpu
def test_device_rescoring_hand_checked_and_errors(oracle, gpu_ctx):
    from locityper_b200 import ffi
    a = _alns(HAND)
    assert _same(genotype.rescore_alignments(gpu_ctx, a), oracle.rescore_alignments(a))
    with pytest.raises(ffi.LctpError):
        genotype.rescore_alignments(gpu_ctx, _alns([("50=", 0, 50, 100, 1), ("20M", 0, 20, 100, 1)]))
    bad = _alns([("50=", 0, 50, 100, 1), ("10=", 0, 10, 100, 1)])
    bad.cigar_off = np.array([0, 0, 2], dtype=np.uint64)
    with pytest.raises(ffi.LctpError):
        genotype.rescore_alignments(gpu_ctx, bad)
    empty = genotype.Alns(**synth.make_alns(0, 1))
    assert len(genotype.rescore_alignments(gpu_ctx, empty)["ln_prob"]) == 0


@pytest.mark.gpu
def test_device_vs_oracle_timing_report(oracle, gpu_ctx, capsys):
    """Not an assertion on speed: prints kernel / whole-call / single-thread oracle times (run with -s); the oracle
    may only be executed from tests/, so the timing comparison lives here rather than in tools/rescore_run.py."""
    import time
    a = genotype.Alns(**synth.make_alns(300_000, 21))
    genotype.rescore_alignments(gpu_ctx, a)
    gpu_ctx.stats(reset=True)
    t0 = time.perf_counter()
    got = genotype.rescore_alignments(gpu_ctx, a)
    wall = time.perf_counter() - t0
    st = gpu_ctx.stats(reset=True)
    t1 = time.perf_counter()
    ref = oracle.rescore_alignments(a)
    cpu = time.perf_counter() - t1
    assert _same(got, ref)
    with capsys.disabled():
        print(f"\n[rescore] 300k records: kernel {st['rescore_ms']:.3f} ms, whole call {wall*1e3:.1f} ms, oracle (1 thread) {cpu*1e3:.1f} ms")


def test_oracle_reproduces_golden_rescoring(oracle):
    from conftest import load_golden_alns
    a, want = load_golden_alns()
    assert _same(oracle.rescore_alignments(a), want) and a.n_alns == 600


@pytest.mark.gpu
def test_device_reproduces_golden_rescoring(gpu_ctx):
    """No oracle at run time: the committed fixture (tests/golden/rescore_small.npz) is the reference."""
    from conftest import load_golden_alns
    a, want = load_golden_alns()
    assert _same(genotype.rescore_alignments(gpu_ctx, a), want)

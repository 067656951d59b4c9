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


# ---- second slice: read_next_alns (src/model/locs.rs:502-567) + push with the PosCollection (:166-187, 297-343) ----

NOT_SAVED = 0xFFFFFFFF


def _read_ends(n_groups, seed, *, contigs=6, per_group=(1, 40), poor_frac=0.5, strict=False, contig_len=3500):
    """Groups of alignment records: the starts of a group cluster around a few positions per contig, so that 128-bp bins
    are hit repeatedly (replacement, NOT_SAVED bins) and the order of the records matters."""
    rng = np.random.default_rng(seed)
    sizes = rng.integers(per_group[0], per_group[1] + 1, n_groups)
    off = np.r_[0, np.cumsum(sizes)].astype(np.uint64)
    n = int(off[-1])
    kw = synth.make_alns(n, seed + 1, contig_len=contig_len)
    rec_contig = np.zeros(n, dtype=np.uint32)
    for g in range(n_groups):
        b, e = int(off[g]), int(off[g + 1])
        anchors = rng.integers(0, contig_len - 400, 3)
        rec_contig[b:e] = rng.integers(0, contigs, e - b)
        span = (kw["aln_end"][b:e] - kw["aln_start"][b:e]).astype(np.int64)
        st = (anchors[rng.integers(0, 3, e - b)] + rng.integers(-90, 91, e - b)).clip(0, None)
        st = np.minimum(st, contig_len - span).clip(0, None)
        kw["aln_start"][b:e] = st
        kw["aln_end"][b:e] = st + span
    good = rng.integers(2, 14, n_groups).astype(np.uint32)
    passable = (good + rng.integers(0, 12, n_groups)).astype(np.uint32)
    compl = np.where(rng.random(n_groups) < poor_frac, rng.random(n_groups) * 0.5, 0.5 + rng.random(n_groups) * 0.5)
    return genotype.ReadEnds(alns=genotype.Alns(**kw), grp_off=off, rec_contig=rec_contig,
                             grp_read_end=rng.integers(0, 2, n_groups).astype(np.uint8),
                             grp_read_len=rng.integers(100, 251, n_groups).astype(np.uint32), grp_good_dist=good,
                             grp_passable_dist=passable, grp_neighb_complexity=compl, poor_compl=0.5,
                             poor_compl_edit=0.15, strict_subset=strict)


def _transcription_read_ends(re_: genotype.ReadEnds) -> dict:
    """read_next_alns + PrelimAlignments::push, line by line; the PosCollection is a dict like the reference's IntMap."""
    import math
    a = re_.alns
    per = _transcription(genotype.Alns(**{**a.__dict__, "passable_dist": np.full(a.n_alns, NOT_SAVED, dtype=np.uint32)}))
    n, ng = a.n_alns, re_.n_groups
    out = dict(ln_prob=per["ln_prob"], edit=per["edit"], read_len=per["read_len"], ok=np.zeros(ng, dtype=np.uint8),
               best_edit=np.zeros(ng, dtype=np.uint32), weight_factor=np.ones(ng), thr_dist=np.zeros(ng, dtype=np.uint32),
               pass_dist=np.zeros(ng, dtype=np.uint32), n_kept=np.zeros(ng, dtype=np.uint32),
               kept_rec=np.full(n, NOT_SAVED, dtype=np.uint32))
    for g in range(ng):
        b, e = int(re_.grp_off[g]), int(re_.grp_off[g + 1])
        read_len = int(re_.grp_read_len[g])
        good_dist, passable_dist = int(re_.grp_good_dist[g]), int(re_.grp_passable_dist[g])     # locs.rs:529
        threshold_dist = good_dist                                                              # :530
        if re_.grp_neighb_complexity[g] <= re_.poor_compl:                                      # :531
            threshold_dist = max(good_dist, int(re_.poor_compl_edit * float(read_len)))         # :532 (`as u32` truncates)
            passable_dist += threshold_dist - good_dist                                         # :533
        out["thr_dist"][g], out["pass_dist"][g] = threshold_dist, passable_dist                 # set_thresholds, :536
        alns, pos_collection = [], {}                       # PrelimAlignments::alns (record indices), PosCollection::map
        best_edit, primary_ok = NOT_SAVED, True
        for i in range(b, e):                                                                   # push, :297-343
            dist_edit, aln_prob = int(per["edit"][i]), float(per["ln_prob"][i])
            best_edit = min(best_edit, dist_edit)                                               # :308
            new_aln_ix = len(alns)                                                              # :311
            save = dist_edit <= passable_dist                                                   # :312
            if new_aln_ix == 0 and not save:                                                    # :314-316
                assert i == b
                primary_ok = False
                break                                                                           # skip_until_primary, :539-541
            key = (int(re_.grp_read_end[g]) << 48) | (int(re_.rec_contig[i]) << 32) | (int(a.aln_start[i]) >> 7)   # :166-168
            if key in pos_collection and save:                                                  # (Occupied, true)
                aln_ix = pos_collection[key]
                if aln_ix == NOT_SAVED:
                    pos_collection[key] = new_aln_ix
                    alns.append(i)
                elif aln_prob > float(per["ln_prob"][alns[aln_ix]]):
                    alns[aln_ix] = i
            elif key in pos_collection:                                                         # (Occupied, false)
                pass
            elif save:                                                                          # (Vacant, true)
                pos_collection[key] = new_aln_ix
                alns.append(i)
            else:                                                                               # (Vacant, false)
                pos_collection[key] = NOT_SAVED
        out["best_edit"][g] = best_edit
        if not primary_ok:
            continue
        out["n_kept"][g] = len(alns)
        out["kept_rec"][b:b + len(alns)] = alns
        req_edit = passable_dist if re_.strict_subset else threshold_dist                       # :560
        if best_edit > req_edit:                                                                # :561-563
            continue
        out["ok"][g] = 1
        out["weight_factor"][g] = 1.0 if best_edit <= good_dist else math.sqrt(float(good_dist) / float(best_edit))   # :564
    return out


def _same_ends(x, y):
    for k in genotype.READ_ENDS_OUT_ORDER:
        assert np.array_equal(x[k], y[k]), k


@pytest.mark.parametrize("seed,strict", [(1, False), (2, True), (3, False)])
def test_oracle_read_ends_equal_python_transcription(oracle, seed, strict):
    re_ = _read_ends(300, seed, strict=strict)
    got = oracle.collect_read_ends(re_)
    ref = _transcription_read_ends(re_)
    _same_ends(got, ref)
    # the cases the protocol distinguishes all occur in the generated data
    assert 0 < got["ok"].sum() < re_.n_groups
    assert (got["n_kept"] == 0).any() and (got["n_kept"] > 1).any()
    assert ((got["weight_factor"] < 1.0) & (got["ok"] == 1)).any()
    sizes = np.diff(re_.grp_off.astype(np.int64))
    assert (got["n_kept"] < sizes).any()                       # bins shared by several records


def test_oracle_read_ends_hand_checked(oracle):
    """One group, contig 0, passable 3: primary at 1000 (edit 0); 1010 same bin, worse -> dropped; 1020 same bin, better
    -> replaces the primary in place; 1200 other bin with edit 5 -> NOT_SAVED bin; 1210 same bin, edit 1 -> saved as #2."""
    recs = [("100=", 1000, 1100, 5000, 0), ("98=2X", 1010, 1110, 5000, 0), ("100=", 1020, 1120, 5000, 0),
            ("95=5X", 1200, 1300, 5000, 0), ("99=1X", 1210, 1310, 5000, 0)]
    a = _alns(recs, ln_oper=(-0.005, -5.8, -6.5, -6.9, -5.8))
    re_ = genotype.ReadEnds(alns=a, grp_off=np.array([0, 5], dtype=np.uint64), rec_contig=np.zeros(5, dtype=np.uint32),
                            grp_read_end=np.array([0], dtype=np.uint8), grp_read_len=np.array([100], dtype=np.uint32),
                            grp_good_dist=np.array([2], dtype=np.uint32), grp_passable_dist=np.array([3], dtype=np.uint32),
                            grp_neighb_complexity=np.array([1.0]), poor_compl=0.5, poor_compl_edit=0.07)
    got = oracle.collect_read_ends(re_)
    assert list(got["edit"]) == [0, 2, 0, 5, 1]
    assert got["ok"][0] == 1 and got["n_kept"][0] == 2 and got["best_edit"][0] == 0 and got["weight_factor"][0] == 1.0
    # equal ln-probabilities do not replace (strict '>'), so the primary keeps slot 0; slot 1 is the record at 1210
    assert list(got["kept_rec"][:2]) == [0, 4]
    _same_ends(got, _transcription_read_ends(re_))
    # a primary above `passable` ends the read end at once
    re_.alns = _alns([("90=10X", 1000, 1100, 5000, 0)] + recs[1:])
    got = oracle.collect_read_ends(re_)
    assert got["ok"][0] == 0 and got["n_kept"][0] == 0 and got["best_edit"][0] == 10
    _same_ends(got, _transcription_read_ends(re_))


@pytest.mark.gpu
@pytest.mark.parametrize("n_groups,seed,strict,per_group", [(300, 1, False, (1, 40)), (300, 2, True, (1, 40)),
                                                              (4000, 5, False, (1, 120)), (2, 7, False, (600, 900))])
def test_device_read_ends_bit_exact(oracle, gpu_ctx, n_groups, seed, strict, per_group):
    re_ = _read_ends(n_groups, seed, strict=strict, per_group=per_group)
    _same_ends(genotype.collect_read_ends(gpu_ctx, re_), oracle.collect_read_ends(re_))


@pytest.mark.gpu
def test_device_read_ends_reject_malformed_input(gpu_ctx):
    re_ = _read_ends(5, 9)
    re_.grp_off = re_.grp_off.copy(); re_.grp_off[-1] -= 1
    with pytest.raises(Exception):
        genotype.collect_read_ends(gpu_ctx, re_)

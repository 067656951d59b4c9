"""Host-side mirror of the reference's interface for the genotype-evaluation path.

Same names, argument meaning and error behaviour as the reference (tprodanov/locityper v1.7.2):
  * `Stage.parse` / `Scheme.parse`   <- Stage::parse / Scheme::parse, src/solvers/solve.rs:150-251
  * solver parameters                <- Greedy / SimAnneal set_param, src/solvers/stoch.rs:130-138,251-258
  * `solve(data, rng, threads)`      <- solve::solve, src/solvers/solve.rs:926-981
  * `Genotyping.to_json()`           <- Genotyping::to_json, src/solvers/solve.rs:732-773
Everything that computes runs in the CUDA library behind include/lctp.h; this module only marshals.
"""
from __future__ import annotations

import ctypes as C
import json
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import ffi
from .synth import Locus


class InvalidInput(ValueError):
    """crate::Error::InvalidInput"""


def parse_pretty_usize(s: str) -> int:
    """PrettyUsize::from_str (src/ext/fmt.rs:113-157): 5k, 2M, 1G, 10_000, inf."""
    if s in ("inf", "Inf", "INF"):
        return (1 << 64) - 1
    if not s:
        raise InvalidInput("Cannot parse an empty string into int")
    last = s[-1]
    if last in "GgBb":
        n, mult = 0, 10 ** 9
    elif last in "Mm":
        n, mult = 0, 10 ** 6
    elif last in "Kk":
        n, mult = 0, 1000
    elif last.isdigit() and last.isascii():
        n, mult = int(last), 10
    else:
        raise InvalidInput(f"Cannot convert string `{s}` to int, unexpected last symbol '{last}'")
    was_digit = mult == 10
    for c in reversed(s[:-1]):
        if c.isdigit() and c.isascii():
            was_digit = True
            n += mult * int(c)
            mult *= 10
        elif c in ",_":
            was_digit = False
        else:
            raise InvalidInput(f"Cannot convert string `{s}` to int, unexpected symbol '{c}'")
    if not was_digit:
        raise InvalidInput(f"Cannot convert string `{s}` to int, unexpected first letter")
    return n


@dataclass
class Stage:
    """One stage of the solving scheme: SOLVER[:PARAM=VALUE,...]."""

    kind: str = "greedy"
    attempts: int = 20                 # Stage::parse defaults, solve.rs:173-174
    in_size: int = 1000
    best_start: bool = True            # Greedy::default, stoch.rs:45-53
    sample_size: int = 10
    plato_size: Optional[int] = None   # Greedy 100 / SimAnneal 10000
    anneal_steps: int = 20000          # SimAnneal::default, stoch.rs:161-169
    init_prob: float = 0.5

    @staticmethod
    def parse(ix: int, s: str) -> "Stage":
        name, _, params = s.partition(":")
        if name == "greedy":
            st = Stage("greedy")
        elif name in ("anneal", "simanneal", "annealing", "simannealing"):
            st = Stage("anneal")
        elif name in ("highs", "gurobi"):
            raise RuntimeError(f"{name} solver is disabled (ILP solvers are feature-gated off in the reference "
                               "default build and are not part of the device path)")
        else:
            raise InvalidInput(f"Unknown solver {name!r}")
        if params:
            for kv in params.split(","):
                if "=" not in kv:
                    raise InvalidInput(f"Could not parse solver definition `{s}`")
                key, val = kv.split("=", 1)
                try:
                    if key in ("i", "input", "in-size"):
                        st.in_size = parse_pretty_usize(val)
                    elif key in ("a", "attempts"):
                        st.attempts = int(val)
                        if not 0 <= st.attempts <= 65535:
                            raise ValueError
                    else:
                        st._set_param(key, val, s)
                except (ValueError, InvalidInput) as e:
                    if isinstance(e, InvalidInput) and "Invalid value" in str(e):
                        raise
                    raise InvalidInput(f"Could not parse `{kv}` in solver definition `{s}`") from e
        if st.attempts <= 0:
            raise InvalidInput(f"At least one attempt is required for each stage (`{s}`)")
        if st.in_size <= 0:
            raise InvalidInput(f"At least one input genotype is required for each stage (`{s}`)")
        return st

    def _set_param(self, key: str, val: str, s: str) -> None:
        k = key.lower()
        if self.kind == "greedy":
            if k in ("x0", "start"):
                if val in ("b", "best"):
                    self.best_start = True
                elif val in ("r", "rand", "random"):
                    self.best_start = False
                else:
                    raise InvalidInput(f"Invalid value of `{key}={val}` in `{s}`: Invalid start value {val}")
            elif k in ("s", "sample"):
                v = parse_pretty_usize(val)
                if v == 0:
                    raise InvalidInput(f"Invalid value of `{key}={val}` in `{s}`: Sample size must be positive")
                self.sample_size = v
            elif k in ("p", "plato"):
                self.plato_size = parse_pretty_usize(val)
            # unknown keys are logged and ignored by the reference (solve.rs:187-188)
        else:
            if k in ("n", "steps"):
                v = parse_pretty_usize(val)
                if v == 0:
                    raise InvalidInput(f"Invalid value of `{key}={val}` in `{s}`: Number of annealing steps (0) must be positive")
                self.anneal_steps = v
            elif k in ("p", "plato"):
                self.plato_size = parse_pretty_usize(val)
            elif key in ("P", "prob", "init-prob") or k in ("prob", "init-prob"):
                v = float(val)
                if not (0.0 < v <= 1.0):
                    raise InvalidInput(f"Invalid value of `{key}={val}` in `{s}`: Initial probability ({v}) must be within (0, 1]")
                self.init_prob = v

    def to_c(self) -> ffi.StageC:
        c = ffi.StageC()
        c.kind = 0 if self.kind == "greedy" else 1
        c.attempts, c.in_size = self.attempts, self.in_size
        c.best_start, c.sample_size = int(self.best_start), self.sample_size
        c.plato_size = self.plato_size if self.plato_size is not None else (100 if c.kind == 0 else 10000)
        c.anneal_steps, c.init_prob = self.anneal_steps, self.init_prob
        return c


@dataclass
class Scheme:
    stages: List[Stage] = field(default_factory=list)

    @staticmethod
    def default() -> "Scheme":
        """DEFAULT_STAGES = "-S greedy:i=5k,a=1 -S anneal:i=20,a=20" (solve.rs:211-230)."""
        return Scheme([Stage("greedy", attempts=1, in_size=5000), Stage("anneal", attempts=20, in_size=20)])

    @staticmethod
    def parse(solvers: Sequence[str]) -> "Scheme":
        if not solvers:
            return Scheme.default()
        return Scheme([Stage.parse(i, s) for i, s in enumerate(solvers)])

    def to_c(self):
        return (ffi.StageC * len(self.stages))(*[s.to_c() for s in self.stages])


def build_depth_table(nb_n, nb_p, is_paired, alt_cn, k_cols: int) -> np.ndarray:
    """DistrCache::new (src/model/distr_cache.rs:61-75) through the C ABI (host code, no GPU needed)."""
    nb_n = np.ascontiguousarray(nb_n, dtype=np.float64)
    nb_p = np.ascontiguousarray(nb_p, dtype=np.float64)
    alt = np.ascontiguousarray(alt_cn, dtype=np.float64)
    out = np.empty((ffi_GC_BINS, int(k_cols)), dtype=np.float64)
    ffi.check(ffi.load().lctp_build_depth_table(nb_n.ctypes.data, nb_p.ctypes.data, int(bool(is_paired)),
                                                alt.ctypes.data, len(alt), int(k_cols), out.ctypes.data))
    return out


ffi_GC_BINS = 101


def init_rng(seed: int) -> np.ndarray:
    """ext::rand::init_rng (src/ext/rand.rs:6-22): xoshiro256++ seeded by SplitMix64."""
    st = np.zeros(4, dtype=np.uint64)
    ffi.load().lctp_rng_seed_from_u64(st.ctypes.data, C.c_uint64(seed))
    return st


def rng_long_jump(state: np.ndarray) -> None:
    ffi.load().lctp_rng_long_jump(state.ctypes.data)


def rng_jump(state: np.ndarray) -> None:
    ffi.load().lctp_rng_jump(state.ctypes.data)


def worker_streams(state: np.ndarray, threads: int) -> np.ndarray:
    """MainWorker::new (src/solvers/solve.rs:1007-1018): u64[threads, 4] worker states; `state` (the locus stream)
    is advanced by `threads` jumps in place."""
    out = np.zeros((int(threads), 4), dtype=np.uint64)
    ffi.load().lctp_rng_worker_streams(state.ctypes.data, int(threads), out.ctypes.data)
    return out


class Context:
    """lctp_ctx: one per GPU (Send, not Sync)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None, max_resident_workers: int = 0):
        self.lib = ffi.load()
        cfg = ffi.DeviceCfg(device=device, flags=0, stream=stream, max_resident_workers=max_resident_workers)
        self._h = C.c_void_p()
        ffi.check(self.lib.lctp_init(C.byref(cfg), C.byref(self._h)))

    def launch_count(self) -> int:
        return int(self.lib.lctp_launch_count(self._h))

    def sync(self) -> None:
        ffi.check(self.lib.lctp_sync(self._h))

    def stats(self, reset: bool = False) -> dict:
        st = ffi.StatsC()
        ffi.check(self.lib.lctp_get_stats(self._h, C.byref(st), int(reset)))
        return {k: getattr(st, k) for k, _ in ffi.StatsC._fields_}

    def fp64_rate(self) -> float:
        """FP64-pipe lane-instructions per second (DADD microbenchmark; roofline denominator of the prefilter)."""
        v = C.c_double(0.0)
        ffi.check(self.lib.lctp_measure_fp64_rate(self._h, C.byref(v)))
        return float(v.value)

    def upload(self, loc: Locus, pairs: Optional["DevicePairs"] = None) -> "DeviceLocus":
        """lctp_locus_upload; with `pairs` (lctp_locus_upload_pairs) the pa_* / unmapped_prob arrays of `loc` are
        ignored and the device-resident pair alignments are used instead."""
        return DeviceLocus(self, loc, pairs)

    def debug_open(self, directory: str, level: int, hap_names: Sequence[str]) -> None:
        """lctp_debug_open: while open, solve() writes sol.csv / sol_ext.csv (level >= 1) / depth.csv (level >= 2) of the
        reference's `--debug` into `directory` (src/solvers/solve.rs:852-952), uncompressed."""
        cn = (C.c_char_p * len(hap_names))(*[s.encode() for s in hap_names])
        ffi.check(self.lib.lctp_debug_open(self._h, directory.encode(), level, cn, len(hap_names)))

    def debug_close(self) -> None:
        self.lib.lctp_debug_close(self._h)

    def close(self) -> None:
        if self._h:
            self.lib.lctp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class Mates:
    """Mate alignments of R read pairs: the input of identify_paired_end_alignments (src/model/locs.rs:805-868),
    per read sorted by (contig asc, read end asc, ln_prob desc).  flags: bit0 = read end, bit1 = strand."""

    n_reads: int
    n_haps: int
    ma_off: np.ndarray
    ma_contig: np.ndarray
    ma_flags: np.ndarray
    ma_start: np.ndarray
    ma_end: np.ndarray
    ma_ln_prob: np.ndarray
    ins_ln_pmf: np.ndarray
    unmapped_penalty: float
    insert_penalty: float
    prob_diff: float
    read_weight: Optional[np.ndarray] = None
    max_alns: int = 10
    single_end: bool = False                      # identify_single_end_alignments (locs.rs:870-911)
    window: int = 0                               # window size, with explicit weights
    exp_off: Optional[np.ndarray] = None          # u64[H+1]: explicit region weights per contig position
    exp_weight: Optional[np.ndarray] = None       # f64 (contig_len + 1 entries per contig)
    read_max_alns: Optional[np.ndarray] = None    # u8[R]: per-read max_alns (10 / 2 by read weight, locs.rs:1263)

    def to_c(self, keep: list, struct=None):
        def arr(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data
        m = (struct or ffi.MatesC)()
        m.n_reads, m.n_haps, m.max_alns, m.ins_len = self.n_reads, self.n_haps, self.max_alns, len(self.ins_ln_pmf)
        m.ma_off = arr(self.ma_off, np.uint64)
        m.ma_contig = arr(self.ma_contig, np.uint32)
        m.ma_flags = arr(self.ma_flags, np.uint8)
        m.ma_start = arr(self.ma_start, np.uint32)
        m.ma_end = arr(self.ma_end, np.uint32)
        m.ma_ln_prob = arr(self.ma_ln_prob, np.float64)
        m.read_weight = arr(self.read_weight, np.float64)
        m.ins_ln_pmf = arr(self.ins_ln_pmf, np.float64)
        m.unmapped_penalty, m.insert_penalty, m.prob_diff = self.unmapped_penalty, self.insert_penalty, self.prob_diff
        m.single_end, m.window = int(self.single_end), int(self.window)
        m.exp_off = arr(self.exp_off, np.uint64)
        m.exp_weight = arr(self.exp_weight, np.float64)
        m.read_max_alns = arr(self.read_max_alns, np.uint8)
        return m


def pair_alignments(ctx: "Context", mates: Mates, cap: Optional[int] = None) -> dict:
    """lctp_pair_alignments: the pa_* / unmapped_prob arrays of the flat locus, computed on the device."""
    keep: list = []
    m = mates.to_c(keep)
    R = mates.n_reads
    n_groups_bound = len(mates.ma_contig)            # every group keeps <= max_alns and has >= 1 mate
    cap = int(cap if cap is not None else min(n_groups_bound * mates.max_alns, max(1, n_groups_bound) * mates.max_alns))
    pa_off = np.zeros(R + 1, dtype=np.uint64)
    pa_contig = np.zeros(max(1, cap), dtype=np.uint32)
    pa_ln_prob = np.zeros(max(1, cap), dtype=np.float64)
    pa_mid1 = np.zeros(max(1, cap), dtype=np.uint32)
    pa_mid2 = np.zeros(max(1, cap), dtype=np.uint32)
    unm = np.zeros(R, dtype=np.float64)
    n_out = C.c_uint64(0)
    ffi.check(ctx.lib.lctp_pair_alignments(ctx._h, C.byref(m), cap, pa_off.ctypes.data, pa_contig.ctypes.data,
                                           pa_ln_prob.ctypes.data, pa_mid1.ctypes.data, pa_mid2.ctypes.data,
                                           unm.ctypes.data, C.byref(n_out)))
    n = int(n_out.value)
    return dict(pa_off=pa_off, pa_contig=pa_contig[:n].copy(), pa_ln_prob=pa_ln_prob[:n].copy(),
                pa_mid1=pa_mid1[:n].copy(), pa_mid2=pa_mid2[:n].copy(), unmapped_prob=unm)


class DevicePairs:
    """lctp_pairs_h: the pair alignments of a locus, resident on the device (lctp_pair_alignments_dev).  Feed it to
    `Context.upload(loc, pairs=...)` instead of the pa_* arrays; `fetch()` copies it to the host when needed."""

    def __init__(self, ctx: "Context", mates: Mates):
        self.ctx, self.lib, self.n_reads = ctx, ctx.lib, mates.n_reads
        keep: list = []
        m = mates.to_c(keep)
        self._h = C.c_void_p()
        n = C.c_uint64(0)
        ffi.check(self.lib.lctp_pair_alignments_dev(ctx._h, C.byref(m), C.byref(self._h), C.byref(n)))
        self.n_pairs = int(n.value)

    def fetch(self) -> dict:
        R, n = self.n_reads, self.n_pairs
        pa_off = np.zeros(R + 1, dtype=np.uint64)
        con, lp = np.zeros(max(1, n), dtype=np.uint32), np.zeros(max(1, n))
        m1, m2 = np.zeros(max(1, n), dtype=np.uint32), np.zeros(max(1, n), dtype=np.uint32)
        unm = np.zeros(R)
        ffi.check(self.lib.lctp_pairs_fetch(self._h, n, pa_off.ctypes.data, con.ctypes.data, lp.ctypes.data,
                                            m1.ctypes.data, m2.ctypes.data, unm.ctypes.data))
        return dict(pa_off=pa_off, pa_contig=con[:n].copy(), pa_ln_prob=lp[:n].copy(), pa_mid1=m1[:n].copy(),
                    pa_mid2=m2[:n].copy(), unmapped_prob=unm)

    def free(self) -> None:
        if self._h:
            self.lib.lctp_pairs_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class Alns:
    """Alignment records of one locus before pairing: the input of the per-alignment part of
    PrelimAlignments::push (src/model/locs.rs:297-313).  cigar_ops are BAM-encoded (len << 4 | op; 1=I 2=D 4=S 7='='
    8=X); `ln_oper` = ErrorProfile::oper_probs (ln of match, mismatch, insertion, deletion, clipping)."""

    cigar_off: np.ndarray
    cigar_ops: np.ndarray
    aln_start: np.ndarray
    aln_end: np.ndarray
    contig_len: np.ndarray
    passable_dist: np.ndarray
    ln_oper: tuple

    @property
    def n_alns(self) -> int:
        return len(self.aln_start)

    def to_c(self, keep: list, struct=None):
        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data
        c = (struct or ffi.AlnsC)()
        c.n_alns = self.n_alns
        c.cigar_off = arr(self.cigar_off, np.uint64)
        c.cigar_ops = arr(self.cigar_ops, np.uint32)
        c.aln_start = arr(self.aln_start, np.uint32)
        c.aln_end = arr(self.aln_end, np.uint32)
        c.contig_len = arr(self.contig_len, np.uint32)
        c.passable_dist = arr(self.passable_dist, np.uint32)
        c.ln_match, c.ln_mismatch, c.ln_insertion, c.ln_deletion, c.ln_clipping = (float(v) for v in self.ln_oper)
        return c


def _concat(seqs: Sequence[bytes]):
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    return off, np.frombuffer(b"".join(seqs) or b"\0", dtype=np.uint8).copy()


def minimizers(ctx: "Context", seqs: Sequence[bytes], k: int, w: int):
    """lctp_minimizers: canonical minimizers (kmers::minimizers::<u64,_,CANONICAL>) of every sequence, on the device.
    Returns a list of (hash u64[], pos u32[], forward u8[]) per sequence."""
    off, buf = _concat(seqs)
    n, total = len(seqs), max(1, int(off[-1]))
    cnt = np.zeros(n, dtype=np.uint32)
    h, p, f = np.zeros(total, dtype=np.uint64), np.zeros(total, dtype=np.uint32), np.zeros(total, dtype=np.uint8)
    ffi.check(ctx.lib.lctp_minimizers(ctx._h, buf.ctypes.data, off.ctypes.data, n, k, w, cnt.ctypes.data, h.ctypes.data,
                                      p.ctypes.data, f.ctypes.data))
    return [(h[int(off[s]):int(off[s]) + int(cnt[s])].copy(), p[int(off[s]):int(off[s]) + int(cnt[s])].copy(),
             f[int(off[s]):int(off[s]) + int(cnt[s])].copy()) for s in range(n)]


@dataclass
class TargetSeqs:
    """Input of TargetBuilder::add for every locus (src/seq/recruit.rs:680-735): the allele sequences of each locus, their
    k-mer counts (KmerCounts: one u16 per base_k-mer of the sequence) and the recruitment Params."""

    seqs: Sequence[bytes]
    seq_locus: np.ndarray
    kmer_counts: Sequence[np.ndarray]
    base_k: int
    minimizer_k: int
    minimizer_w: int
    thresh_kmer_count: int
    match_frac: float
    match_length: int = 2000

    def to_c(self, keep: list, struct=None):
        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data
        off, buf = _concat(self.seqs)
        coff = np.zeros(len(self.seqs) + 1, dtype=np.uint64)
        coff[1:] = np.cumsum([len(c) for c in self.kmer_counts])
        c = (struct or ffi.TargetSeqsC)()
        c.n_seqs = len(self.seqs)
        c.seq_off, c.seqs, c.seq_locus = arr(off, np.uint64), arr(buf, np.uint8), arr(self.seq_locus, np.uint32)
        c.cnt_off = arr(coff, np.uint64)
        c.kmer_counts = arr(np.concatenate(self.kmer_counts) if len(self.kmer_counts) else np.zeros(1), np.uint16)
        c.base_k, c.minimizer_k, c.minimizer_w = int(self.base_k), int(self.minimizer_k), int(self.minimizer_w)
        c.thresh_kmer_count, c.match_frac = int(self.thresh_kmer_count), float(self.match_frac)
        c.match_length = int(self.match_length)
        return c


@dataclass
class Reads:
    """Short reads: first mates (or single-end reads) and optionally the second mates of the same pairs."""

    seq1: Sequence[bytes]
    seq2: Optional[Sequence[bytes]] = None

    def to_c(self, keep: list, struct=None):
        c = (struct or ffi.ReadsC)()
        c.n_reads = len(self.seq1)
        for name, seqs in (("1", self.seq1), ("2", self.seq2)):
            if seqs is None:
                continue
            off, buf = _concat(seqs)
            keep += [off, buf]
            setattr(c, "off" + name, off.ctypes.data)
            setattr(c, "seq" + name, buf.ctypes.data)
        return c


class Targets:
    """lctp_targets_h: the recruitment targets (minimizer -> loci table) resident on the device."""

    def __init__(self, ctx: "Context", ts: TargetSeqs):
        self.ctx, self.lib = ctx, ctx.lib
        keep: list = []
        c = ts.to_c(keep)
        self._h = C.c_void_p()
        ffi.check(self.lib.lctp_targets_build(ctx._h, C.byref(c), C.byref(self._h)))

    def entries(self):
        n = int(self.lib.lctp_targets_entries(self._h, None, None, None, 0))
        k, l, i = np.zeros(max(1, n), dtype=np.uint64), np.zeros(max(1, n), dtype=np.uint32), np.zeros(max(1, n), dtype=np.uint8)
        self.lib.lctp_targets_entries(self._h, k.ctypes.data, l.ctypes.data, i.ctypes.data, n)
        return k[:n], l[:n], i[:n]

    def match_frac(self):
        a, b = C.c_uint16(), C.c_uint16()
        self.lib.lctp_targets_match_frac(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def recruit(self, reads: Reads, cap: int = 8):
        """lctp_recruit_short: per read (pair) the sorted list of loci it is recruited to."""
        keep: list = []
        c = reads.to_c(keep)
        n = len(reads.seq1)
        cnt = np.zeros(max(1, n), dtype=np.uint32)
        ans = np.zeros(max(1, n) * cap, dtype=np.uint32)
        ffi.check(self.lib.lctp_recruit_short(self.ctx._h, self._h, C.byref(c), cap, cnt.ctypes.data, ans.ctypes.data))
        return [list(ans[r * cap:r * cap + min(int(cnt[r]), cap)]) for r in range(n)]

    def free(self) -> None:
        if self._h:
            self.lib.lctp_targets_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class ReadEnds:
    """Alignment records grouped by (read, read end) + the per-group inputs of read_next_alns (src/model/locs.rs:502-567)."""

    alns: Alns                     # passable_dist ignored
    grp_off: np.ndarray
    rec_contig: np.ndarray
    grp_read_end: np.ndarray
    grp_read_len: np.ndarray
    grp_good_dist: np.ndarray
    grp_passable_dist: np.ndarray
    grp_neighb_complexity: np.ndarray
    poor_compl: float
    poor_compl_edit: float
    strict_subset: bool = False

    @property
    def n_groups(self) -> int:
        return len(self.grp_off) - 1

    def to_c(self, keep: list, struct=None, alns_struct=None):
        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data
        c = (struct or ffi.ReadEndsC)()
        c.alns = self.alns.to_c(keep, struct=alns_struct)
        c.n_groups = self.n_groups
        c.grp_off = arr(self.grp_off, np.uint64)
        c.rec_contig = arr(self.rec_contig, np.uint32)
        c.grp_read_end = arr(self.grp_read_end, np.uint8)
        c.grp_read_len = arr(self.grp_read_len, np.uint32)
        c.grp_good_dist = arr(self.grp_good_dist, np.uint32)
        c.grp_passable_dist = arr(self.grp_passable_dist, np.uint32)
        c.grp_neighb_complexity = arr(self.grp_neighb_complexity, np.float64)
        c.poor_compl, c.poor_compl_edit, c.strict_subset = float(self.poor_compl), float(self.poor_compl_edit), int(self.strict_subset)
        return c


def _read_ends_outputs(re_: ReadEnds) -> dict:
    n, ng = re_.alns.n_alns, re_.n_groups
    return dict(ln_prob=np.zeros(n), edit=np.zeros(n, dtype=np.uint32), read_len=np.zeros(n, dtype=np.uint32),
                ok=np.zeros(ng, dtype=np.uint8), best_edit=np.zeros(ng, dtype=np.uint32), weight_factor=np.zeros(ng),
                thr_dist=np.zeros(ng, dtype=np.uint32), pass_dist=np.zeros(ng, dtype=np.uint32),
                n_kept=np.zeros(ng, dtype=np.uint32), kept_rec=np.full(n, 0xFFFFFFFF, dtype=np.uint32))


READ_ENDS_OUT_ORDER = ("ln_prob", "edit", "read_len", "ok", "best_edit", "weight_factor", "thr_dist", "pass_dist", "n_kept",
                       "kept_rec")


def collect_read_ends(ctx: "Context", re_: ReadEnds) -> dict:
    """lctp_collect_read_ends: the read_next_alns protocol + PosCollection de-duplication for every read end, on the device."""
    keep: list = []
    c = re_.to_c(keep)
    out = _read_ends_outputs(re_)
    ffi.check(ctx.lib.lctp_collect_read_ends(ctx._h, C.byref(c), *[out[k].ctypes.data for k in READ_ENDS_OUT_ORDER]))
    return out


@dataclass
class Prelim:
    """Input of lctp_group_reads: the records and outputs of lctp_collect_read_ends plus, per read, its two groups
    (AllAlignments::load after read_next_alns + recover_and_group_alignments without the transfer,
    src/model/locs.rs:1117-1137, 1237-1288)."""

    read_group: np.ndarray          # i64[n_reads, 2], -1 = none
    grp_off: np.ndarray             # u64[n_groups + 1]
    rec_contig: np.ndarray
    rec_start: np.ndarray
    rec_end: np.ndarray
    rec_strand: np.ndarray          # u8, 1 = reverse
    rec_ln_prob: np.ndarray
    grp_ok: np.ndarray
    grp_best_edit: np.ndarray
    grp_thr_dist: np.ndarray
    grp_n_kept: np.ndarray
    kept_rec: np.ndarray
    contig_len: np.ndarray
    read_weight: np.ndarray
    min_weight: float
    boundary: int
    single_end: bool = False

    @property
    def n_reads(self) -> int:
        return len(self.read_group)

    def to_c(self, keep: list, struct=None):
        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data
        c = (struct or ffi.PrelimC)()
        c.n_reads, c.n_groups = self.n_reads, len(self.grp_off) - 1
        c.read_group = arr(self.read_group, np.int64)
        c.grp_off = arr(self.grp_off, np.uint64)
        c.rec_contig, c.rec_start, c.rec_end = (arr(a, np.uint32) for a in (self.rec_contig, self.rec_start, self.rec_end))
        c.rec_strand = arr(self.rec_strand, np.uint8)
        c.rec_ln_prob = arr(self.rec_ln_prob, np.float64)
        c.grp_ok = arr(self.grp_ok, np.uint8)
        c.grp_best_edit, c.grp_thr_dist, c.grp_n_kept, c.kept_rec, c.contig_len = (
            arr(a, np.uint32) for a in (self.grp_best_edit, self.grp_thr_dist, self.grp_n_kept, self.kept_rec, self.contig_len))
        c.read_weight = arr(self.read_weight, np.float64)
        c.min_weight, c.n_haps, c.boundary, c.single_end = self.min_weight, len(self.contig_len), self.boundary, int(self.single_end)
        return c

    def alloc_outputs(self) -> dict:
        R, n = self.n_reads, max(1, len(self.rec_contig))
        return dict(status=np.zeros(max(1, R), dtype=np.uint8), out_read=np.zeros(max(1, R), dtype=np.uint32),
                    out_max_alns=np.zeros(max(1, R), dtype=np.uint8), ma_off=np.zeros(R + 1, dtype=np.uint64),
                    ma_contig=np.zeros(n, dtype=np.uint32), ma_flags=np.zeros(n, dtype=np.uint8),
                    ma_start=np.zeros(n, dtype=np.uint32), ma_end=np.zeros(n, dtype=np.uint32),
                    ma_ln_prob=np.zeros(n), ma_rec=np.zeros(n, dtype=np.uint32), counts=np.zeros(3, dtype=np.uint64))

    def trim_outputs(self, out: dict, n_out: int) -> dict:
        n = int(out["ma_off"][n_out]) if n_out else 0
        res = dict(status=out["status"][:self.n_reads].copy(), n_reads_out=n_out, counts=out["counts"].copy(),
                   out_read=out["out_read"][:n_out].copy(), out_max_alns=out["out_max_alns"][:n_out].copy(),
                   ma_off=out["ma_off"][:n_out + 1].copy())
        for k in ("ma_contig", "ma_flags", "ma_start", "ma_end", "ma_ln_prob", "ma_rec"):
            res[k] = out[k][:n].copy()
        return res


def group_reads(ctx: "Context", pre: Prelim) -> dict:
    """lctp_group_reads: read status, normalised ln-probabilities and the pairing input in consumption order."""
    keep: list = []
    c = pre.to_c(keep)
    out = pre.alloc_outputs()
    n_out = C.c_uint64(0)
    ffi.check(ctx.lib.lctp_group_reads(ctx._h, C.byref(c), max(1, len(pre.rec_contig)), out["status"].ctypes.data,
                                       C.byref(n_out), *[out[k].ctypes.data for k in (
                                           "out_read", "out_max_alns", "ma_off", "ma_contig", "ma_flags", "ma_start",
                                           "ma_end", "ma_ln_prob", "ma_rec", "counts")]))
    return pre.trim_outputs(out, int(n_out.value))


class DeviceMates:
    """lctp_mates_h: the pairing input of a locus, resident on the device (lctp_group_reads_dev).  `pair(params)` runs the
    pairing on it (lctp_pair_alignments_from) and returns DevicePairs-like access to the result."""

    def __init__(self, ctx: "Context", pre: Prelim):
        self.ctx, self.lib = ctx, ctx.lib
        keep: list = []
        c = pre.to_c(keep)
        R = pre.n_reads
        self.status = np.zeros(max(1, R), dtype=np.uint8)
        out_read = np.zeros(max(1, R), dtype=np.uint32)
        self.counts = np.zeros(3, dtype=np.uint64)
        n_out = C.c_uint64(0)
        self._h = C.c_void_p()
        ffi.check(self.lib.lctp_group_reads_dev(ctx._h, C.byref(c), self.status.ctypes.data, C.byref(n_out),
                                                out_read.ctypes.data, self.counts.ctypes.data, C.byref(self._h)))
        self.n_reads_out = int(n_out.value)
        self.status, self.out_read = self.status[:R], out_read[:self.n_reads_out].copy()
        self.n_entries = int(self.lib.lctp_mates_count(self._h))

    def pair(self, params: Mates) -> "DevicePairs":
        keep: list = []
        m = params.to_c(keep)
        dp = DevicePairs.__new__(DevicePairs)
        dp.ctx, dp.lib, dp.n_reads = self.ctx, self.lib, self.n_reads_out
        dp._h = C.c_void_p()
        n = C.c_uint64(0)
        ffi.check(self.lib.lctp_pair_alignments_from(self.ctx._h, self._h, C.byref(m), C.byref(dp._h), C.byref(n)))
        dp.n_pairs = int(n.value)
        return dp

    def free(self):
        if self._h:
            self.lib.lctp_mates_free(self._h)
            self._h = C.c_void_p()


def _seq_arrays(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    cat = np.frombuffer(b"".join(seqs), dtype=np.uint8) if int(off[-1]) else np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(cat), off


class UniqueKmers:
    """lctp_unique_kmers_h: the k-mers unique to a locus (UniqueKmers::new, src/model/locs.rs:930-963) as a device table;
    `read_weights` = calculate_read_weight (:968-1002) for all reads in one launch."""

    def __init__(self, ctx: "Context", contig_seqs, kmer_counts, k: int, hard_threshold: int, soft_threshold: int):
        self.ctx, self.lib = ctx, ctx.lib
        cat, off = _seq_arrays(contig_seqs)
        cnt_off = np.zeros(len(kmer_counts) + 1, dtype=np.uint64)
        cnt_off[1:] = np.cumsum([len(c) for c in kmer_counts])
        cnt = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.uint16) for c in kmer_counts] or [np.zeros(0, np.uint16)]))
        if len(cnt) == 0:
            cnt = np.zeros(1, dtype=np.uint16)
        self._h = C.c_void_p()
        n = C.c_uint64(0)
        ffi.check(self.lib.lctp_unique_kmers_build(ctx._h, cat.ctypes.data, off.ctypes.data, len(contig_seqs), cnt.ctypes.data,
                                                   cnt_off.ctypes.data, k, hard_threshold, soft_threshold,
                                                   C.byref(self._h), C.byref(n)))
        self.n_unique = int(n.value)

    def read_weights(self, read_seqs, ends: int):
        """read_seqs: list of bytes, `ends` per read (b"" = no mate) -> (unique u16[n * ends], weight f64[n])."""
        n = len(read_seqs) // ends
        cat, off = _seq_arrays(read_seqs)
        unique, weight = np.zeros(max(1, n * ends), dtype=np.uint16), np.zeros(max(1, n))
        ffi.check(self.lib.lctp_read_weights(self.ctx._h, self._h, cat.ctypes.data, off.ctypes.data, n, ends,
                                             unique.ctypes.data, weight.ctypes.data))
        return unique[:n * ends], weight[:n]

    def free(self):
        if self._h:
            self.lib.lctp_unique_kmers_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def counts_to_prob(counts: np.ndarray, attempts: int):
    """lctp_counts_to_prob: (prob f32, mapq u8) of assignment counts, count_to_prob of src/model/bam.rs:54-66 (host)."""
    counts = np.ascontiguousarray(counts, dtype=np.uint16)
    prob, mapq = np.zeros(len(counts), dtype=np.float32), np.zeros(len(counts), dtype=np.uint8)
    ffi.check(ffi.load().lctp_counts_to_prob(counts.ctypes.data, len(counts), attempts, prob.ctypes.data, mapq.ctypes.data))
    return prob, mapq


def rescore_alignments(ctx: "Context", alns: Alns) -> dict:
    """lctp_rescore_alignments: ln_prob / EditDist / save of every alignment record, computed on the device."""
    keep: list = []
    a = alns.to_c(keep)
    n = alns.n_alns
    ln_prob = np.zeros(n, dtype=np.float64)
    edit = np.zeros(n, dtype=np.uint32)
    read_len = np.zeros(n, dtype=np.uint32)
    save = np.zeros(n, dtype=np.uint8)
    ffi.check(ctx.lib.lctp_rescore_alignments(ctx._h, C.byref(a), ln_prob.ctypes.data, edit.ctypes.data,
                                              read_len.ctypes.data, save.ctypes.data))
    return dict(ln_prob=ln_prob, edit=edit, read_len=read_len, save=save)


class ContextPool:
    """K contexts (one CUDA stream each) on one GPU, one host thread per context.  Loci are independent
    units of work (src/command/genotype.rs:1331-1351: one `analyze_locus` per locus, separate long_jump
    RNG streams), so their uploads, kernels and host-side pruning overlap; a context still has a single
    locus in flight (lctp_ctx is Send, not Sync)."""

    def __init__(self, device: int = 0, k: int = 3, max_resident_workers: int = 0):
        from concurrent.futures import ThreadPoolExecutor
        self.ctxs = [Context(device=device, max_resident_workers=max_resident_workers) for _ in range(k)]
        self._pool = ThreadPoolExecutor(max_workers=k)

    def map(self, fn, items):
        """fn(ctx, index, item) for every item; items i, i+K, i+2K, ... run in order on context i."""
        k = len(self.ctxs)
        items = list(items)

        def lane(c):
            return [(i, fn(self.ctxs[c], i, items[i])) for i in range(c, len(items), k)]

        out = [None] * len(items)
        for fut in [self._pool.submit(lane, c) for c in range(min(k, len(items)))]:
            for i, r in fut.result():
                out[i] = r
        return out

    def launch_count(self) -> int:
        return sum(c.launch_count() for c in self.ctxs)

    def stats(self, reset: bool = False) -> dict:
        tot: dict = {}
        for c in self.ctxs:
            for k_, v in c.stats(reset).items():
                tot[k_] = tot.get(k_, 0) + v
        return tot

    def close(self) -> None:
        self._pool.shutdown(wait=True)
        for c in self.ctxs:
            c.close()


def locus_to_c(loc: Locus, keep: list) -> ffi.LocusC:
    def arr(a, dt):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data

    s = ffi.LocusC()
    s.n_haps, s.n_reads, s.ploidy, s.is_paired = loc.n_haps, loc.n_reads, loc.ploidy, int(loc.is_paired)
    s.n_genotypes = loc.n_genotypes
    s.gt_tuples = arr(loc.gt_tuples, np.uint32)
    s.priors = arr(loc.priors, np.float64)
    s.unmapped_prob = arr(loc.unmapped_prob, np.float64)
    s.pa_off = arr(loc.pa_off, np.uint64)
    s.pa_contig = arr(loc.pa_contig, np.uint32)
    s.pa_ln_prob = arr(loc.pa_ln_prob, np.float64)
    s.pa_mid1 = arr(loc.pa_mid1, np.uint32)
    s.pa_mid2 = arr(loc.pa_mid2, np.uint32)
    s.hap_len = arr(loc.hap_len, np.uint32)
    s.hap_n_windows = arr(loc.hap_n_windows, np.uint32)
    s.hap_reg_start = arr(loc.hap_reg_start, np.uint32)
    s.window, s.left_padding = loc.window, loc.left_padding
    s.hap_pos_off = arr(loc.hap_pos_off, np.uint64)
    s.pos_weight = arr(loc.pos_weight, np.float64)
    s.pos_gc = arr(loc.pos_gc, np.uint8)
    if loc.depth_table is None:
        raise InvalidInput("depth table missing: call synth.attach_depth_table(loc, genotype.build_depth_table)")
    s.depth_k, s.tweak = loc.depth_k, loc.tweak
    s.depth_table = arr(loc.depth_table, np.float64)
    s.prob_diff, s.lik_skew, s.min_weight = loc.prob_diff, loc.lik_skew, loc.min_weight
    s.filt_diff, s.prob_thresh = loc.filt_diff, loc.prob_thresh
    s.dont_skip, s.out_bams = int(loc.dont_skip), loc.out_bams
    return s


@dataclass
class Genotyping:
    """solve::Genotyping (src/solvers/solve.rs:568-590)."""

    gt_ix: np.ndarray
    lik_mean: np.ndarray
    lik_var: np.ndarray
    attempts: np.ndarray
    ln_prob: np.ndarray
    quality: float
    total_reads: int
    unexpl_reads: int
    warnings: List[str]
    n_filtered: int
    n_stage_in: List[int]
    t_prefilter_s: float
    t_stages_s: float
    json_text: str = ""
    weight_dist: Optional[float] = None          # Genotyping::weighted_dist
    distances: Optional[List[Optional[int]]] = None   # Genotyping::distances (None entry = "unknown")

    def to_json(self) -> dict:
        return json.loads(self.json_text)


class DeviceLocus:
    """lctp_locus_h: a locus resident in HBM."""

    def __init__(self, ctx: Context, loc: Locus, pairs: Optional["DevicePairs"] = None):
        self.ctx, self.loc, self.lib = ctx, loc, ctx.lib
        self._keep: list = []
        self.c = locus_to_c(loc, self._keep)
        self._h = C.c_void_p()
        if pairs is None:
            ffi.check(self.lib.lctp_locus_upload(ctx._h, C.byref(self.c), C.byref(self._h)))
        else:
            ffi.check(self.lib.lctp_locus_upload_pairs(ctx._h, C.byref(self.c), pairs._h, C.byref(self._h)))

    def free(self) -> None:
        if self._h:
            self.lib.lctp_locus_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def best_aln_matrix(self) -> np.ndarray:
        M = np.empty((self.loc.n_haps, self.loc.n_reads), dtype=np.float64)
        ffi.check(self.lib.lctp_best_aln_matrix(self._h, M.ctypes.data))
        return M

    def prefilter_scores(self, g_begin: int = 0, g_end: Optional[int] = None, fetch: bool = True):
        g_end = self.loc.n_genotypes if g_end is None else g_end
        out = np.empty(g_end - g_begin, dtype=np.float64) if fetch else None
        ffi.check(self.lib.lctp_prefilter_scores(self._h, g_begin, g_end, None if out is None else out.ctypes.data))
        return out

    def prefilter(self, min_size: int, threads: int, ixs: Optional[np.ndarray] = None, want_scores: bool = False):
        """run_filter: returns the sorted survivors (and all scores if requested)."""
        G = self.loc.n_genotypes
        ixs = np.arange(G, dtype=np.uint64) if ixs is None else np.array(ixs, dtype=np.uint64)
        n_out = C.c_size_t(0)
        scores = np.empty(G, dtype=np.float64) if want_scores else None
        ffi.check(self.lib.lctp_prefilter(self._h, ixs.ctypes.data, len(ixs), min_size, threads, C.byref(n_out),
                                          None if scores is None else scores.ctypes.data))
        surv = ixs[:n_out.value].copy()
        return (surv, scores) if want_scores else surv

    def solve_stage(self, stage: Stage, worker_ixs, worker_off, worker_rng: np.ndarray, want_liks: bool = True,
                    want_counts: bool = False, counts_cap: int = 0):
        worker_ixs = np.ascontiguousarray(worker_ixs, dtype=np.uint64)
        worker_off = np.ascontiguousarray(worker_off, dtype=np.uint64)
        assert worker_rng.dtype == np.uint64 and worker_rng.flags.c_contiguous
        n_workers, n = len(worker_off) - 1, int(worker_off[-1])
        st = stage.to_c()
        lik_mean, lik_var = np.empty(n), np.empty(n)
        liks = np.empty((n, stage.attempts)) if want_liks else None
        n_alns, iters = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
        counts_off = counts = None
        if want_counts:
            counts_off = np.zeros(n + 1, dtype=np.uint64)
            counts = np.zeros(max(1, counts_cap), dtype=np.uint16)
        p = lambda a: None if a is None else a.ctypes.data
        ffi.check(self.lib.lctp_solve_stage(self._h, C.byref(st), p(worker_ixs), p(worker_off), n_workers,
                                            p(worker_rng), p(lik_mean), p(lik_var), p(liks), p(counts_off), p(counts),
                                            int(counts_cap), p(n_alns), p(iters)))
        return dict(lik_mean=lik_mean, lik_var=lik_var, liks=liks, n_alns=n_alns, iters=iters,
                    counts_off=counts_off, counts=counts)

    def produce_result(self, ixs, lik_mean, lik_var, attempts) -> dict:
        """Predictions::produce_result + the Genotyping checks (solve.rs:482-535,637-678,719-729).
        Per-genotype arrays are indexed by genotype id; `ixs` are the surviving ids."""
        ixs = np.array(ixs, dtype=np.uint64)
        lm = np.ascontiguousarray(lik_mean, dtype=np.float64)
        lv = np.ascontiguousarray(lik_var, dtype=np.float64)
        at = np.ascontiguousarray(attempts, dtype=np.uint16)
        res = ffi.ResultC()
        ffi.check(self.lib.lctp_produce_result(self._h, ixs.ctypes.data, len(ixs), lm.ctypes.data, lv.ctypes.data,
                                               at.ctypes.data, C.byref(res)))
        n = int(res.n_out)
        return dict(gt_ix=np.array(res.gt_ix[:n], dtype=np.uint64), lik_mean=np.array(res.lik_mean[:n]),
                    lik_var=np.array(res.lik_var[:n]), attempts=np.array(res.attempts[:n]),
                    ln_prob=np.array(res.ln_prob[:n]), quality=res.quality, total_reads=res.total_reads,
                    unexpl_reads=res.unexpl_reads, warn_no_probable=bool(res.warn_no_probable),
                    warn_few_reads=bool(res.warn_few_reads))

    def solve(self, scheme: Scheme, threads: int, rng: np.ndarray, hap_names: Optional[Sequence[str]] = None,
              contig_distances: Optional[np.ndarray] = None, true_edit_distances: bool = False) -> Genotyping:
        """solve::solve: prefilter -> stages -> result.  `rng` (u64[4], the locus stream) is updated in place.
        `contig_distances`: Data::contig_distances as the linear upper triangle (u32, 0xFFFFFFFF = None)."""
        st = scheme.to_c()
        res = ffi.ResultC()
        ffi.check(self.lib.lctp_solve(self._h, st, len(scheme.stages), threads, rng.ctypes.data, C.byref(res)))
        if contig_distances is not None:
            cd = np.ascontiguousarray(contig_distances, dtype=np.uint32)
            assert len(cd) == self.loc.n_haps * (self.loc.n_haps - 1) // 2
            ffi.check(self.lib.lctp_find_weighted_dist(C.byref(res), C.byref(self.c), cd.ctypes.data,
                                                       int(true_edit_distances)))
        return self._genotyping(res, hap_names)

    def solve_counts(self, scheme: Scheme, threads: int, rng: np.ndarray, n_counts: int):
        """lctp_solve_counts: solve::solve + Prediction::assgn_counts of the first n_counts reported genotypes (the input
        of write_bam, src/model/bam.rs:356-413).  Returns (Genotyping, [u16 counts array per genotype])."""
        st = scheme.to_c()
        res = ffi.ResultC()
        cap = n_counts * 65535
        off = np.zeros(n_counts + 1, dtype=np.uint64)
        cnt = np.zeros(cap, dtype=np.uint16)
        ffi.check(self.lib.lctp_solve_counts(self._h, st, len(scheme.stages), threads, rng.ctypes.data, C.byref(res),
                                             n_counts, off.ctypes.data, cnt.ctypes.data, cap))
        return self._genotyping(res), [cnt[int(off[k]):int(off[k + 1])].copy() for k in range(n_counts)]

    def solve_stage_debug(self, stage: Stage, worker_ixs, worker_off, worker_rng: np.ndarray, windows: bool = True) -> dict:
        """lctp_solve_stage_dbg: a stage + the fields of ReadAssignment::summarize / write_depth per (genotype, attempt)."""
        worker_ixs = np.ascontiguousarray(worker_ixs, dtype=np.uint64)
        worker_off = np.ascontiguousarray(worker_off, dtype=np.uint64)
        n, na = int(worker_off[-1]), int(worker_off[-1]) * stage.attempts
        wmax = int(self.lib.lctp_locus_wmax(self._h))
        out = dict(lik_mean=np.empty(n), lik_var=np.empty(n), liks=np.empty((n, stage.attempts)), aln_lik=np.empty(na),
                   depth_lik=np.empty(na), unmapped=np.zeros(na, dtype=np.uint32), out_of_bounds=np.zeros(na, dtype=np.uint32),
                   wmax=wmax)
        d = ffi.StageDebugC(out["aln_lik"].ctypes.data, out["depth_lik"].ctypes.data, out["unmapped"].ctypes.data,
                            out["out_of_bounds"].ctypes.data, wmax, 0, None, None, None)
        if windows:
            out.update(win_weight=np.zeros((na, wmax)), win_depth=np.zeros((na, wmax), dtype=np.uint32), win_lik=np.zeros((na, wmax)))
            d.win_weight, d.win_depth, d.win_lik = (out[k].ctypes.data for k in ("win_weight", "win_depth", "win_lik"))
        stc = stage.to_c()
        ffi.check(self.lib.lctp_solve_stage_dbg(self._h, C.byref(stc), worker_ixs.ctypes.data, worker_off.ctypes.data,
                                                len(worker_off) - 1, worker_rng.ctypes.data, out["lik_mean"].ctypes.data,
                                                out["lik_var"].ctypes.data, out["liks"].ctypes.data, None, None, 0, None,
                                                None, C.byref(d)))
        return out

    def _genotyping(self, res, hap_names=None) -> "Genotyping":
        n = int(res.n_out)
        names = list(hap_names) if hap_names is not None else [f"hap{i}" for i in range(self.loc.n_haps)]
        cn = (C.c_char_p * len(names))(*[s.encode() for s in names])
        need = self.lib.lctp_result_json(C.byref(res), C.byref(self.c), cn, None, 0)
        buf = C.create_string_buffer(need + 1)
        self.lib.lctp_result_json(C.byref(res), C.byref(self.c), cn, buf, need + 1)
        warnings = []
        if res.warn_no_probable:
            warnings.append("NoProbableGenotype")
        if res.warn_few_reads:
            warnings.append(f"FewReads({res.total_reads})")
        return Genotyping(
            gt_ix=np.array(res.gt_ix[:n], dtype=np.uint64), lik_mean=np.array(res.lik_mean[:n]),
            lik_var=np.array(res.lik_var[:n]), attempts=np.array(res.attempts[:n]),
            ln_prob=np.array(res.ln_prob[:n]), quality=res.quality, total_reads=res.total_reads,
            unexpl_reads=res.unexpl_reads, warnings=warnings, n_filtered=int(res.n_filtered),
            n_stage_in=[int(x) for x in res.n_stage_in], t_prefilter_s=res.t_prefilter_s,
            t_stages_s=res.t_stages_s, json_text=buf.value.decode(),
            weight_dist=(res.weight_dist if res.has_dist and res.has_weight_dist else None),
            distances=([None if d == 0xFFFFFFFF else int(d) for d in res.dist_to_primary[:n]]
                       if res.has_dist else None))


def _nccl_before_lctp() -> None:
    """The library dlopen()s "libnccl.so.2" on first use.  In a Python process that will also import torch, torch's own
    NCCL must be the copy that gets loaded under that name (a later `import torch` would otherwise be handed the
    system copy and fail on missing symbols): import torch first when it is installed."""
    try:
        import torch  # noqa: F401
    except ImportError:
        pass


def dist_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls it, the 128 bytes travel to the other ranks out of band)."""
    _nccl_before_lctp()
    buf = (C.c_uint8 * 128)()
    ffi.check(ffi.load().lctp_dist_unique_id(buf))
    return bytes(buf)


class Dist:
    """lctp_dist: this rank's membership in the NCCL communicator that shards one locus over the GPUs of a box
    (SURVEY.md section 8e).  Every method is collective: all ranks call it with the same arguments."""

    def __init__(self, ctx: "Context", unique_id: bytes, rank: int, world: int):
        _nccl_before_lctp()
        self.lib = ffi.load()
        self.ctx, self.rank, self.world = ctx, rank, world
        self._h = C.c_void_p()
        idb = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        ffi.check(self.lib.lctp_dist_init(ctx._h, idb, rank, world, C.byref(self._h)))

    def close(self) -> None:
        if self._h:
            self.lib.lctp_dist_destroy(self._h)
            self._h = C.c_void_p()

    def timing(self, reset: bool = False) -> dict:
        t = ffi.DistTimingC()
        ffi.check(self.lib.lctp_dist_get_timing(self._h, C.byref(t), int(reset)))
        return {k: getattr(t, k) for k, _ in ffi.DistTimingC._fields_}

    def prefilter(self, dl: "DeviceLocus", min_size: int, threads: int) -> np.ndarray:
        cap = dl.loc.n_genotypes
        ixs = np.zeros(cap, dtype=np.uint64)
        n = C.c_size_t(0)
        ffi.check(self.lib.lctp_dist_prefilter(self._h, dl._h, min_size, threads, ixs.ctypes.data, cap, C.byref(n)))
        return ixs[:n.value].copy()

    def solve_stage(self, dl: "DeviceLocus", stage: Stage, worker_ixs, worker_off, worker_rng: np.ndarray):
        worker_ixs = np.ascontiguousarray(worker_ixs, dtype=np.uint64)
        worker_off = np.ascontiguousarray(worker_off, dtype=np.uint64)
        n = int(worker_off[-1])
        lm, lv = np.empty(n), np.empty(n)
        st = stage.to_c()
        ffi.check(self.lib.lctp_dist_solve_stage(self._h, dl._h, C.byref(st), worker_ixs.ctypes.data,
                                                 worker_off.ctypes.data, len(worker_off) - 1, worker_rng.ctypes.data,
                                                 lm.ctypes.data, lv.ctypes.data))
        return lm, lv

    def solve(self, dl: "DeviceLocus", scheme: Scheme, threads: int, rng: np.ndarray,
              hap_names: Optional[Sequence[str]] = None) -> "Genotyping":
        st = scheme.to_c()
        res = ffi.ResultC()
        ffi.check(self.lib.lctp_dist_solve(self._h, dl._h, st, len(scheme.stages), threads, rng.ctypes.data, C.byref(res)))
        return dl._genotyping(res, hap_names)


def truncate_ixs(ixs, scores, filt_diff: float, min_size: int, threads: int) -> np.ndarray:
    ixs = np.array(ixs, dtype=np.uint64)
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    m = ffi.load().lctp_truncate_ixs(ixs.ctypes.data, len(ixs), scores.ctypes.data, filt_diff, min_size, threads)
    return ixs[:m].copy()


def plan_stage(rng: np.ndarray, ixs: np.ndarray, threads: int):
    """MainWorker::run :1049-1063: shuffle (in place) + contiguous partition. Returns worker_off."""
    assert ixs.dtype == np.uint64 and ixs.flags.c_contiguous
    off = np.zeros(threads + 1, dtype=np.uint64)
    nw = ffi.load().lctp_plan_stage(rng.ctypes.data, ixs.ctypes.data, len(ixs), threads, off.ctypes.data)
    return off[:nw + 1].copy()


def discard_improbable(ixs, lik_mean, lik_var, attempts, prob_thresh: float, out_size: int, threads: int) -> np.ndarray:
    ixs = np.array(ixs, dtype=np.uint64)
    lm = np.ascontiguousarray(lik_mean, dtype=np.float64)
    lv = np.ascontiguousarray(lik_var, dtype=np.float64)
    at = np.ascontiguousarray(attempts, dtype=np.uint16)
    m = ffi.load().lctp_discard_improbable(ixs.ctypes.data, len(ixs), lm.ctypes.data, lv.ctypes.data, at.ctypes.data,
                                           prob_thresh, out_size, threads)
    return ixs[:m].copy()

"""ctypes binding of the CPU ORACLE (oracle/_build/liblcto.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs -- never from the product package `locityper_b200`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblcto.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("lcto_rng.c", "lcto_specfun.c", "lcto_model.c", "lcto_solve.c", "lcto_pairs.c", "lcto_rescore.c",
                                             "lcto_recruit.c", "lcto_group.c", "lcto_weights.c", "lcto.h")]
    stale = force or not os.path.exists(_SO) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


class Rng(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4)]

    @classmethod
    def from_seed(cls, seed: int) -> "Rng":
        r = cls()
        lib().lcto_rng_seed_from_u64(C.byref(r), C.c_uint64(seed))
        return r

    @classmethod
    def from_state(cls, s: Sequence[int]) -> "Rng":
        r = cls()
        for i in range(4):
            r.s[i] = int(s[i])
        return r

    def state(self):
        return [int(self.s[i]) for i in range(4)]


class LocusC(C.Structure):
    _fields_ = [
        ("n_haps", C.c_uint32), ("n_reads", C.c_uint32), ("ploidy", C.c_uint32), ("is_paired", C.c_uint32),
        ("n_genotypes", C.c_uint64),
        ("gt_tuples", C.c_void_p), ("priors", C.c_void_p), ("unmapped_prob", C.c_void_p),
        ("pa_off", C.c_void_p), ("pa_contig", C.c_void_p), ("pa_ln_prob", C.c_void_p),
        ("pa_mid1", C.c_void_p), ("pa_mid2", C.c_void_p),
        ("hap_len", C.c_void_p), ("hap_n_windows", C.c_void_p), ("hap_reg_start", C.c_void_p),
        ("window", C.c_uint32), ("left_padding", C.c_uint32),
        ("hap_pos_off", C.c_void_p), ("pos_weight", C.c_void_p), ("pos_gc", C.c_void_p),
        ("depth_k", C.c_uint32), ("tweak", C.c_uint32),
        ("depth_table", C.c_void_p),
        ("prob_diff", C.c_double), ("lik_skew", C.c_double), ("min_weight", C.c_double),
        ("filt_diff", C.c_double), ("prob_thresh", C.c_double),
        ("dont_skip", C.c_uint32), ("out_bams", C.c_uint32),
    ]


class StageC(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32), ("attempts", C.c_uint32), ("in_size", C.c_uint64),
        ("best_start", C.c_uint32), ("_pad", C.c_uint32),
        ("sample_size", C.c_uint64), ("plato_size", C.c_uint64), ("anneal_steps", C.c_uint64),
        ("init_prob", C.c_double),
    ]


class ResultC(C.Structure):
    _fields_ = [
        ("n_out", C.c_uint64), ("gt_ix", C.c_uint64 * 50), ("lik_mean", C.c_double * 50),
        ("lik_var", C.c_double * 50), ("attempts", C.c_uint16 * 50), ("ln_prob", C.c_double * 50),
        ("quality", C.c_double), ("total_reads", C.c_uint32), ("unexpl_reads", C.c_uint32),
        ("warn_no_probable", C.c_uint32), ("warn_few_reads", C.c_uint32),
        ("n_filtered", C.c_uint64), ("n_stage_in", C.c_uint64 * 8),
        ("t_prefilter_s", C.c_double), ("t_stages_s", C.c_double),
        ("has_dist", C.c_uint32), ("true_edit_distances", C.c_uint32), ("has_weight_dist", C.c_uint32),
        ("_pad", C.c_uint32), ("weight_dist", C.c_double), ("dist_to_primary", C.c_uint32 * 50),
    ]


class AttemptOut(C.Structure):
    _fields_ = [("lik", C.c_double), ("aln_lik", C.c_double), ("depth_lik", C.c_double),
                ("iterations", C.c_uint64), ("moves", C.c_uint64)]


class InstanceC(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint32), ("ploidy", C.c_uint32), ("total_windows", C.c_uint32),
        ("n_alns", C.c_uint32), ("n_nontrivial", C.c_uint32),
        ("haps", C.c_uint32 * 16), ("wshift", C.c_uint32 * 17),
        ("read_ixs", C.POINTER(C.c_uint32)), ("nontrivial", C.POINTER(C.c_uint32)),
        ("aln_ln_prob", C.POINTER(C.c_double)), ("aln_contig_ix", C.POINTER(C.c_uint8)),
        ("aln_pa", C.POINTER(C.c_uint32)), ("aln_w", C.POINTER(C.c_uint32)),
        ("win_weight", C.POINTER(C.c_double)), ("win_gc", C.POINTER(C.c_uint8)),
        ("win_trivial", C.POINTER(C.c_uint8)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.lcto_sizeof_locus.restype = C.c_size_t
        L.lcto_sizeof_stage.restype = C.c_size_t
        assert L.lcto_sizeof_locus() == C.sizeof(LocusC), (L.lcto_sizeof_locus(), C.sizeof(LocusC))
        assert L.lcto_sizeof_stage() == C.sizeof(StageC)
        L.lcto_rng_next_u64.restype = C.c_uint64
        L.lcto_rng_next_u32.restype = C.c_uint32
        L.lcto_rng_f64.restype = C.c_double
        L.lcto_rng_range_u32_incl.restype = C.c_uint32
        L.lcto_rng_range_u32_incl.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.lcto_rng_range_i32_incl.restype = C.c_int32
        L.lcto_rng_range_i32_incl.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.lcto_rng_range_usize.restype = C.c_size_t
        L.lcto_rng_range_usize.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        L.lcto_rng_range_u16.restype = C.c_uint16
        L.lcto_rng_range_u16.argtypes = [C.c_void_p, C.c_uint16, C.c_uint16]
        L.lcto_rng_sample_indices.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.lcto_rng_shuffle_usize.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        for f in ("lcto_ln_gamma",):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_double]
        L.lcto_beta_reg.restype = C.c_double
        L.lcto_beta_reg.argtypes = [C.c_double] * 3
        L.lcto_students_t_cdf.restype = C.c_double
        L.lcto_students_t_cdf.argtypes = [C.c_double] * 2
        L.lcto_ln_add.restype = C.c_double
        L.lcto_ln_add.argtypes = [C.c_double] * 2
        L.lcto_build_depth_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                             C.c_uint32, C.c_void_p]
        L.lcto_genotype_tuple.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.lcto_best_aln_matrix.argtypes = [C.c_void_p, C.c_void_p]
        L.lcto_prefilter_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.lcto_truncate_ixs.restype = C.c_size_t
        L.lcto_truncate_ixs.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_size_t, C.c_size_t]
        L.lcto_instance_new.restype = C.POINTER(InstanceC)
        L.lcto_instance_new.argtypes = [C.c_void_p, C.c_uint64]
        L.lcto_instance_free.argtypes = [C.c_void_p]
        L.lcto_apply_tweak.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.lcto_solve_attempt.argtypes = [C.c_void_p] * 7
        L.lcto_solve_stage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.lcto_discard_improbable.restype = C.c_size_t
        L.lcto_discard_improbable.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_double, C.c_size_t, C.c_size_t]
        L.lcto_compare_two_likelihoods.restype = C.c_double
        L.lcto_compare_two_likelihoods.argtypes = [C.c_double, C.c_double, C.c_uint16,
                                                   C.c_double, C.c_double, C.c_uint16]
        L.lcto_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p]
        L.lcto_pair_alignments.restype = C.c_int
        L.lcto_pair_alignments.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 6
        L.lcto_rescore_alignments.restype = C.c_int
        L.lcto_rescore_alignments.argtypes = [C.c_void_p] * 5
        L.lcto_version.restype = C.c_char_p
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def build_depth_table(nb_n, nb_p, is_paired, alt_cn, k_cols: int) -> np.ndarray:
    nb_n = np.ascontiguousarray(nb_n, dtype=np.float64)
    nb_p = np.ascontiguousarray(nb_p, dtype=np.float64)
    alt = np.ascontiguousarray(alt_cn, dtype=np.float64)
    out = np.empty((101, int(k_cols)), dtype=np.float64)
    lib().lcto_build_depth_table(_ptr(nb_n), _ptr(nb_p), int(bool(is_paired)), _ptr(alt), len(alt),
                                 int(k_cols), _ptr(out))
    return out


class OracleLocus:
    """Keeps the numpy buffers alive behind an `lcto_locus` struct."""

    def __init__(self, loc):
        self.loc = loc
        self._keep = []

        def arr(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return a.ctypes.data

        s = LocusC()
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
        assert loc.depth_table is not None, "attach_depth_table first"
        s.depth_k, s.tweak = loc.depth_k, loc.tweak
        s.depth_table = arr(loc.depth_table, np.float64)
        s.prob_diff, s.lik_skew, s.min_weight = loc.prob_diff, loc.lik_skew, loc.min_weight
        s.filt_diff, s.prob_thresh = loc.filt_diff, loc.prob_thresh
        s.dont_skip, s.out_bams = int(loc.dont_skip), loc.out_bams
        self.c = s

    @property
    def ref(self):
        return C.byref(self.c)


@dataclass
class Stage:
    kind: str = "greedy"          # "greedy" | "anneal"
    attempts: int = 20            # Stage::parse default (src/solvers/solve.rs:173-174)
    in_size: int = 1000
    best_start: bool = True       # Greedy defaults (src/solvers/stoch.rs:45-53)
    sample_size: int = 10
    plato_size: Optional[int] = None   # greedy 100, anneal 10000
    anneal_steps: int = 20000     # SimAnneal defaults (src/solvers/stoch.rs:161-169)
    init_prob: float = 0.5

    def to_c(self) -> StageC:
        s = StageC()
        s.kind = 0 if self.kind == "greedy" else 1
        s.attempts, s.in_size = self.attempts, self.in_size
        s.best_start, s.sample_size = int(self.best_start), self.sample_size
        s.plato_size = self.plato_size if self.plato_size is not None else (100 if s.kind == 0 else 10000)
        s.anneal_steps, s.init_prob = self.anneal_steps, self.init_prob
        return s


DEFAULT_SCHEME = [Stage("greedy", attempts=1, in_size=5000), Stage("anneal", attempts=20, in_size=20)]


def best_aln_matrix(ol: OracleLocus) -> np.ndarray:
    M = np.empty((ol.loc.n_haps, ol.loc.n_reads), dtype=np.float64)
    lib().lcto_best_aln_matrix(ol.ref, _ptr(M))
    return M


def prefilter_scores(ol: OracleLocus, ixs: Optional[np.ndarray] = None, M: Optional[np.ndarray] = None):
    G = ol.loc.n_genotypes
    if ixs is None:
        ixs = np.arange(G, dtype=np.uint64)
    ixs = np.ascontiguousarray(ixs, dtype=np.uint64)
    if M is None:
        M = best_aln_matrix(ol)
    scores = np.empty(G, dtype=np.float64)
    lib().lcto_prefilter_scores(ol.ref, _ptr(M), _ptr(ixs), len(ixs), _ptr(scores))
    return scores


def truncate_ixs(ixs: np.ndarray, scores: np.ndarray, filt_diff: float, min_size: int, threads: int):
    ixs = np.array(ixs, dtype=np.uint64)
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    m = lib().lcto_truncate_ixs(_ptr(ixs), len(ixs), _ptr(scores), filt_diff, min_size, threads)
    return ixs[:m].copy()


def solve_stage(ol: OracleLocus, stage: Stage, worker_ixs, worker_off, worker_rng: np.ndarray,
                os_threads: int = 1, want_counts: bool = False, counts_cap: int = 0):
    """worker_rng: u64[n_workers, 4] (updated in place)."""
    worker_ixs = np.ascontiguousarray(worker_ixs, dtype=np.uint64)
    worker_off = np.ascontiguousarray(worker_off, dtype=np.uint64)
    assert worker_rng.dtype == np.uint64 and worker_rng.flags.c_contiguous
    n_workers = len(worker_off) - 1
    n = int(worker_off[-1])
    st = stage.to_c()
    lik_mean = np.empty(n); lik_var = np.empty(n)
    liks = np.empty((n, stage.attempts))
    n_alns = np.zeros(n, dtype=np.uint64); iters = np.zeros(n, dtype=np.uint64)
    counts_off = counts = None
    if want_counts:
        counts_off = np.zeros(n + 1, dtype=np.uint64)
        counts = np.zeros(max(1, counts_cap), dtype=np.uint16)
    rc = lib().lcto_solve_stage(ol.ref, C.byref(st), _ptr(worker_ixs), _ptr(worker_off), n_workers,
                                _ptr(worker_rng), os_threads, _ptr(lik_mean), _ptr(lik_var), _ptr(liks),
                                _ptr(counts_off), _ptr(counts), int(counts_cap), _ptr(n_alns), _ptr(iters))
    if rc != 0:
        raise RuntimeError(f"lcto_solve_stage failed: {rc}")
    return dict(lik_mean=lik_mean, lik_var=lik_var, liks=liks, n_alns=n_alns, iters=iters,
                counts_off=counts_off, counts=counts)


def solve(ol: OracleLocus, scheme: Sequence[Stage], threads: int, rng: Rng, os_threads: int = 1,
          want_scores: bool = False, contig_distances: Optional[np.ndarray] = None, true_edit_distances: bool = False):
    G = ol.loc.n_genotypes
    st = (StageC * len(scheme))(*[s.to_c() for s in scheme])
    res = ResultC()
    scores = np.empty(G) if want_scores else None
    filt = np.zeros(G, dtype=np.uint64) if want_scores else None
    rc = lib().lcto_solve(ol.ref, st, len(scheme), threads, C.byref(rng), os_threads, C.byref(res),
                          _ptr(scores), _ptr(filt))
    if rc != 0:
        raise RuntimeError(f"lcto_solve failed: {rc}")
    if contig_distances is not None:
        cd = np.ascontiguousarray(contig_distances, dtype=np.uint32)
        lib().lcto_find_weighted_dist(ol.ref, C.byref(res), _ptr(cd), int(true_edit_distances))
    n = int(res.n_out)
    out = dict(
        gt_ix=np.array(res.gt_ix[:n], dtype=np.uint64), lik_mean=np.array(res.lik_mean[:n]),
        lik_var=np.array(res.lik_var[:n]), attempts=np.array(res.attempts[:n]),
        ln_prob=np.array(res.ln_prob[:n]), quality=res.quality, total_reads=res.total_reads,
        unexpl_reads=res.unexpl_reads, warn_no_probable=bool(res.warn_no_probable),
        warn_few_reads=bool(res.warn_few_reads), n_filtered=int(res.n_filtered),
        n_stage_in=[int(x) for x in res.n_stage_in], t_prefilter_s=res.t_prefilter_s,
        t_stages_s=res.t_stages_s,
        has_dist=bool(res.has_dist), true_edit_distances=bool(res.true_edit_distances),
        weight_dist=(res.weight_dist if res.has_dist and res.has_weight_dist else None),
        distances=([None if d == 0xFFFFFFFF else int(d) for d in res.dist_to_primary[:n]] if res.has_dist else None),
    )
    if want_scores:
        out["scores"] = scores
        out["filtered_ixs"] = filt[:out["n_filtered"]].copy()
    return out


# ---- Genotyping::to_json as text (src/solvers/solve.rs:732-773; written with write_pretty(.., 4),
# src/command/genotype.rs:1256).  Independent of the C++ writer in the product: Python's repr() supplies the
# shortest round-trip digits, the printing rules are the `json` crate's (util/print_dec.rs, restated).

def json_number(v) -> str:
    import math
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    v = float(v)
    if math.isnan(v) or math.isinf(v):
        return "null"
    sign = "-" if math.copysign(1.0, v) < 0 else ""
    v = abs(v)
    if v == 0.0:
        return sign + "0"
    rp = repr(v)                                       # shortest round-trip digits
    mant, _, ex = rp.partition("e") if "e" in rp else (rp, "", "0")
    ip, _, fp = mant.partition(".")
    exponent = int(ex) - len(fp)                       # v = int(ip + fp) * 10^exponent
    digits = (ip + fp).lstrip("0")
    stripped = digits.rstrip("0")
    exponent += len(digits) - len(stripped)
    digits = stripped
    k = len(digits)
    e10 = exponent + k - 1
    if exponent == 0:
        return sign + digits
    if exponent < 0:
        e = -exponent
        if e < 18:
            if k > e:
                return sign + digits[:k - e] + "." + digits[k - e:]
            return sign + "0." + "0" * (e - k) + digits
        return sign + digits[0] + ("." + digits[1:] if k > 1 else "") + "e" + str(e10)
    if k + exponent <= 20:
        return sign + digits + "0" * exponent
    return sign + digits[0] + ("." + digits[1:] if k > 1 else "") + "e" + str(e10)


def to_json_text(res: dict, loc, hap_names: Sequence[str]) -> str:
    import math
    INV_LN10 = 0.4342944819032518277
    ind = "    "

    def name(g):
        return ",".join(hap_names[h] for h in loc.genotype_tuple(int(g)))

    items = [("total_reads", str(int(res["total_reads"]))), ("quality", json_number(res["quality"]))]
    if res.get("distances") is not None:
        items.append(("dist_type", '"edit"' if res["true_edit_distances"] else '"minim-div"'))
    if res.get("weight_dist") is not None:
        items.append(("weight_dist", json_number(res["weight_dist"])))
    items.append(("unexpl_reads", str(int(res["unexpl_reads"]))))
    n = len(res["gt_ix"])
    if n:
        items.append(("genotype", '"%s"' % name(res["gt_ix"][0])))
        opts = []
        for i in range(n):
            o = [("genotype", '"%s"' % name(res["gt_ix"][i])),
                 ("lik_mean", json_number(float(res["lik_mean"][i]) * INV_LN10)),
                 ("lik_sd", json_number(float(res["lik_var"][i]) * INV_LN10)),
                 ("prob", json_number(math.exp(float(res["ln_prob"][i])))),
                 ("log10_prob", json_number(float(res["ln_prob"][i]) * INV_LN10))]
            if res.get("distances") is not None:
                d = res["distances"][i]
                o.append(("dist_to_primary", '"unknown"' if d is None else str(d)))
            opts.append("{\n" + ",\n".join(f'{ind * 3}"{k}": {v}' for k, v in o) + "\n" + ind * 2 + "}")
        items.append(("options", "[\n" + ",\n".join(ind * 2 + x for x in opts) + "\n" + ind + "]"))
    warns = []
    if res.get("warn_no_probable"):
        warns.append('"NoProbableGenotype"')
    if res.get("warn_few_reads"):
        warns.append('"FewReads(%d)"' % int(res["total_reads"]))
    if warns:
        items.append(("warnings", "[\n" + ",\n".join(ind * 2 + w for w in warns) + "\n" + ind + "]"))
    return "{\n" + ",\n".join(f'{ind}"{k}": {v}' for k, v in items) + "\n}"


class MatesC(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_haps", C.c_uint32), ("max_alns", C.c_uint32), ("ins_len", C.c_uint32),
                ("ma_off", C.c_void_p), ("ma_contig", C.c_void_p), ("ma_flags", C.c_void_p), ("ma_start", C.c_void_p),
                ("ma_end", C.c_void_p), ("ma_ln_prob", C.c_void_p), ("read_weight", C.c_void_p),
                ("ins_ln_pmf", C.c_void_p),
                ("unmapped_penalty", C.c_double), ("insert_penalty", C.c_double), ("prob_diff", C.c_double),
                ("single_end", C.c_uint32), ("window", C.c_uint32), ("exp_off", C.c_void_p), ("exp_weight", C.c_void_p),
                ("read_max_alns", C.c_void_p)]


def pair_alignments(mates) -> dict:
    """lcto_pair_alignments on a locityper_b200.genotype.Mates-shaped object (plain data; the oracle does not
    import the product's code paths, only reads the arrays)."""
    keep: list = []
    m = mates.to_c(keep, struct=MatesC)
    R = mates.n_reads
    cap = max(1, len(mates.ma_contig) * mates.max_alns)
    pa_off = np.zeros(R + 1, dtype=np.uint64)
    pa_contig = np.zeros(cap, dtype=np.uint32)
    pa_ln_prob = np.zeros(cap, dtype=np.float64)
    pa_mid1 = np.zeros(cap, dtype=np.uint32)
    pa_mid2 = np.zeros(cap, dtype=np.uint32)
    unm = np.zeros(R, dtype=np.float64)
    rc = lib().lcto_pair_alignments(C.byref(m), cap, pa_off.ctypes.data, pa_contig.ctypes.data, pa_ln_prob.ctypes.data,
                                    pa_mid1.ctypes.data, pa_mid2.ctypes.data, unm.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"lcto_pair_alignments failed: {rc}")
    n = int(pa_off[R])
    return dict(pa_off=pa_off, pa_contig=pa_contig[:n].copy(), pa_ln_prob=pa_ln_prob[:n].copy(),
                pa_mid1=pa_mid1[:n].copy(), pa_mid2=pa_mid2[:n].copy(), unmapped_prob=unm)


class AlnsC(C.Structure):
    _fields_ = [("n_alns", C.c_uint64), ("cigar_off", C.c_void_p), ("cigar_ops", C.c_void_p), ("aln_start", C.c_void_p),
                ("aln_end", C.c_void_p), ("contig_len", C.c_void_p), ("passable_dist", C.c_void_p),
                ("ln_match", C.c_double), ("ln_mismatch", C.c_double), ("ln_insertion", C.c_double),
                ("ln_deletion", C.c_double), ("ln_clipping", C.c_double)]


def rescore_alignments(alns) -> dict:
    """lcto_rescore_alignments on a locityper_b200.genotype.Alns-shaped object (plain data)."""
    keep: list = []
    a = alns.to_c(keep, struct=AlnsC)
    n = alns.n_alns
    ln_prob = np.zeros(n, dtype=np.float64)
    edit = np.zeros(n, dtype=np.uint32)
    read_len = np.zeros(n, dtype=np.uint32)
    save = np.zeros(n, dtype=np.uint8)
    rc = lib().lcto_rescore_alignments(C.byref(a), ln_prob.ctypes.data, edit.ctypes.data, read_len.ctypes.data,
                                       save.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"lcto_rescore_alignments failed: {rc}")
    return dict(ln_prob=ln_prob, edit=edit, read_len=read_len, save=save)


class ReadEndsC(C.Structure):
    _fields_ = [("alns", AlnsC), ("n_groups", C.c_uint64), ("grp_off", C.c_void_p), ("rec_contig", C.c_void_p),
                ("grp_read_end", C.c_void_p), ("grp_read_len", C.c_void_p), ("grp_good_dist", C.c_void_p),
                ("grp_passable_dist", C.c_void_p), ("grp_neighb_complexity", C.c_void_p), ("poor_compl", C.c_double),
                ("poor_compl_edit", C.c_double), ("strict_subset", C.c_uint32)]


def collect_read_ends(re_) -> dict:
    """lcto_collect_read_ends on a locityper_b200.genotype.ReadEnds-shaped object (plain data)."""
    keep: list = []
    c = re_.to_c(keep, struct=ReadEndsC, alns_struct=AlnsC)
    n, ng = re_.alns.n_alns, re_.n_groups
    out = dict(ln_prob=np.zeros(n), edit=np.zeros(n, dtype=np.uint32), read_len=np.zeros(n, dtype=np.uint32),
               ok=np.zeros(ng, dtype=np.uint8), best_edit=np.zeros(ng, dtype=np.uint32), weight_factor=np.zeros(ng),
               thr_dist=np.zeros(ng, dtype=np.uint32), pass_dist=np.zeros(ng, dtype=np.uint32),
               n_kept=np.zeros(ng, dtype=np.uint32), kept_rec=np.full(n, 0xFFFFFFFF, dtype=np.uint32))
    order = ("ln_prob", "edit", "read_len", "ok", "best_edit", "weight_factor", "thr_dist", "pass_dist", "n_kept", "kept_rec")
    lib().lcto_collect_read_ends.argtypes = [C.c_void_p] * 11
    rc = lib().lcto_collect_read_ends(C.byref(c), *[out[k].ctypes.data for k in order])
    if rc != 0:
        raise RuntimeError(f"lcto_collect_read_ends failed: {rc}")
    return out


# ---- per read-end results -> pairing input (lcto_group.c) ----

class PrelimC(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_groups", C.c_uint64), ("read_group", C.c_void_p), ("grp_off", C.c_void_p),
                ("rec_contig", C.c_void_p), ("rec_start", C.c_void_p), ("rec_end", C.c_void_p), ("rec_strand", C.c_void_p),
                ("rec_ln_prob", C.c_void_p), ("grp_ok", C.c_void_p), ("grp_best_edit", C.c_void_p),
                ("grp_thr_dist", C.c_void_p), ("grp_n_kept", C.c_void_p), ("kept_rec", C.c_void_p),
                ("contig_len", C.c_void_p), ("read_weight", C.c_void_p), ("min_weight", C.c_double),
                ("n_haps", C.c_uint32), ("boundary", C.c_uint32), ("single_end", C.c_uint32), ("_pad", C.c_uint32)]


def group_reads(pre) -> dict:
    """lcto_group_reads on a locityper_b200.genotype.Prelim-shaped object (plain data)."""
    keep: list = []
    c = pre.to_c(keep, struct=PrelimC)
    R, n = pre.n_reads, len(pre.rec_contig)
    out = pre.alloc_outputs()
    n_out = C.c_uint64(0)
    lib().lcto_group_reads.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 12
    rc = lib().lcto_group_reads(C.byref(c), max(1, n), out["status"].ctypes.data, C.byref(n_out),
                                *[out[k].ctypes.data for k in ("out_read", "out_max_alns", "ma_off", "ma_contig", "ma_flags",
                                                               "ma_start", "ma_end", "ma_ln_prob", "ma_rec", "counts")])
    if rc != 0:
        raise RuntimeError(f"lcto_group_reads failed: {rc}")
    return pre.trim_outputs(out, int(n_out.value))


# ---- read weights from unique k-mers (lcto_weights.c) ----

def _seq_arrays(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    cat = np.frombuffer(b"".join(seqs), dtype=np.uint8) if int(off[-1]) else np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(cat), off


class UniqueKmers:
    """lcto_unique_kmers_build: UniqueKmers::new.  contig_seqs: list of bytes; kmer_counts: list of u16 arrays."""

    def __init__(self, contig_seqs, kmer_counts, k, hard_threshold, soft_threshold):
        L = lib()
        L.lcto_unique_kmers_build.restype = C.c_void_p
        L.lcto_unique_kmers_build.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32,
                                              C.c_uint16, C.c_uint16]
        L.lcto_unique_kmers_count.restype = C.c_uint64
        L.lcto_unique_kmers_count.argtypes = [C.c_void_p]
        L.lcto_unique_kmers_free.argtypes = [C.c_void_p]
        L.lcto_read_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        cat, off = _seq_arrays(contig_seqs)
        cnt_off = np.zeros(len(kmer_counts) + 1, dtype=np.uint64)
        cnt_off[1:] = np.cumsum([len(c) for c in kmer_counts])
        cnt = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.uint16) for c in kmer_counts] or [np.zeros(0, np.uint16)]))
        if len(cnt) == 0:
            cnt = np.zeros(1, dtype=np.uint16)
        self._h = L.lcto_unique_kmers_build(cat.ctypes.data, off.ctypes.data, len(contig_seqs), cnt.ctypes.data,
                                            cnt_off.ctypes.data, k, hard_threshold, soft_threshold)
        if not self._h:
            raise RuntimeError("lcto_unique_kmers_build failed")
        self.n_unique = int(L.lcto_unique_kmers_count(self._h))

    def read_weights(self, read_seqs, ends):
        """read_seqs: list of bytes, `ends` per read (b"" = no mate) -> (unique u16[n * ends], weight f64[n])."""
        n = len(read_seqs) // ends
        cat, off = _seq_arrays(read_seqs)
        unique, weight = np.zeros(max(1, n * ends), dtype=np.uint16), np.zeros(max(1, n))
        rc = lib().lcto_read_weights(self._h, cat.ctypes.data, off.ctypes.data, n, ends, unique.ctypes.data, weight.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"lcto_read_weights failed: {rc}")
        return unique[:n * ends], weight[:n]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().lcto_unique_kmers_free(self._h)
            self._h = None


# ---- short-read recruitment (lcto_recruit.c) ----

class TargetSeqsC(C.Structure):
    _fields_ = [("n_seqs", C.c_uint64), ("seq_off", C.c_void_p), ("seqs", C.c_void_p), ("seq_locus", C.c_void_p),
                ("cnt_off", C.c_void_p), ("kmer_counts", C.c_void_p), ("base_k", C.c_uint32), ("minimizer_k", C.c_uint32),
                ("minimizer_w", C.c_uint32), ("thresh_kmer_count", C.c_uint32), ("match_frac", C.c_double),
                ("match_length", C.c_uint32), ("_pad", C.c_uint32)]


class ReadsC(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("off1", C.c_void_p), ("seq1", C.c_void_p), ("off2", C.c_void_p),
                ("seq2", C.c_void_p)]


def minimizers(seq: bytes, k: int, w: int):
    n = max(1, len(seq))
    h, p, f = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint8)
    buf = np.frombuffer(seq or b"\0", dtype=np.uint8).copy()
    L = lib()
    L.lcto_minimizers.restype = C.c_size_t
    L.lcto_minimizers.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    c = L.lcto_minimizers(buf.ctypes.data, len(seq), k, w, h.ctypes.data, p.ctypes.data, f.ctypes.data, n)
    return h[:c].copy(), p[:c].copy(), f[:c].copy()


def fraction_approximate_u16(x: float):
    a, b = C.c_uint16(), C.c_uint16()
    L = lib()
    L.lcto_fraction_approximate_u16.argtypes = [C.c_double, C.c_void_p, C.c_void_p]
    L.lcto_fraction_approximate_u16(x, C.byref(a), C.byref(b))
    return a.value, b.value


class Targets:
    def __init__(self, ts):
        keep: list = []
        c = ts.to_c(keep, struct=TargetSeqsC)
        L = lib()
        L.lcto_targets_build.restype = C.c_void_p
        L.lcto_targets_build.argtypes = [C.c_void_p]
        self._h = C.c_void_p(L.lcto_targets_build(C.byref(c)))

    def entries(self):
        L = lib()
        L.lcto_targets_entries.restype = C.c_size_t
        L.lcto_targets_entries.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        n = L.lcto_targets_entries(self._h, None, None, None, 0)
        k, l, i = np.zeros(max(1, n), dtype=np.uint64), np.zeros(max(1, n), dtype=np.uint32), np.zeros(max(1, n), dtype=np.uint8)
        L.lcto_targets_entries(self._h, k.ctypes.data, l.ctypes.data, i.ctypes.data, n)
        return k[:n], l[:n], i[:n]

    def recruit(self, reads, cap: int = 8):
        keep: list = []
        c = reads.to_c(keep, struct=ReadsC)
        n = len(reads.seq1)
        cnt = np.zeros(max(1, n), dtype=np.uint32)
        ans = np.zeros(max(1, n) * cap, dtype=np.uint32)
        L = lib()
        L.lcto_recruit_short.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        rc = L.lcto_recruit_short(self._h, C.byref(c), cap, cnt.ctypes.data, ans.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"lcto_recruit_short failed: {rc}")
        return [list(ans[r * cap:r * cap + min(int(cnt[r]), cap)]) for r in range(n)]

    def __del__(self):
        try:
            lib().lcto_targets_free.argtypes = [C.c_void_p]
            lib().lcto_targets_free(self._h)
        except Exception:
            pass

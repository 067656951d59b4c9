"""ctypes binding of the lctp C ABI (include/lctp.h, locityper_b200/_lib/liblctp.so).

The library is hand-written CUDA for sm_100a; there is NO CPU fallback: every compute entry point fails
with LCTP_E_CUDA when no B200 is present, and `load()` raises if the shared library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LCTP_LIB") or os.path.join(_HERE, "_lib", "liblctp.so")   # LCTP_LIB: tuning builds
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "lctp.h")

OK, E_INVALID, E_CUDA, E_CAPACITY = 0, -1, -2, -3
MAX_OUT, MAX_STAGES = 50, 8


class LctpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"lctp error {code}: {msg}")
        self.code = code


class DeviceCfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("stream", C.c_void_p),
                ("max_resident_workers", C.c_uint32), ("_pad", C.c_uint32)]


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


class StageDebugC(C.Structure):
    _fields_ = [("aln_lik", C.c_void_p), ("depth_lik", C.c_void_p), ("unmapped", C.c_void_p), ("out_of_bounds", C.c_void_p),
                ("wmax", C.c_uint32), ("_pad", C.c_uint32), ("win_weight", C.c_void_p), ("win_depth", C.c_void_p),
                ("win_lik", C.c_void_p)]


class ResultC(C.Structure):
    _fields_ = [
        ("n_out", C.c_uint64), ("gt_ix", C.c_uint64 * MAX_OUT), ("lik_mean", C.c_double * MAX_OUT),
        ("lik_var", C.c_double * MAX_OUT), ("attempts", C.c_uint16 * MAX_OUT), ("ln_prob", C.c_double * MAX_OUT),
        ("quality", C.c_double), ("total_reads", C.c_uint32), ("unexpl_reads", C.c_uint32),
        ("warn_no_probable", C.c_uint32), ("warn_few_reads", C.c_uint32),
        ("n_filtered", C.c_uint64), ("n_stage_in", C.c_uint64 * MAX_STAGES),
        ("t_prefilter_s", C.c_double), ("t_stages_s", C.c_double),
        ("has_dist", C.c_uint32), ("true_edit_distances", C.c_uint32), ("has_weight_dist", C.c_uint32),
        ("_pad", C.c_uint32), ("weight_dist", C.c_double), ("dist_to_primary", C.c_uint32 * MAX_OUT),
    ]


class StatsC(C.Structure):
    _fields_ = [("prefilter_ms", C.c_double), ("prefilter_launches", C.c_uint64), ("prefilter_genotypes", C.c_uint64),
                ("stage_ms", C.c_double), ("stage_launches", C.c_uint64), ("stage_genotypes", C.c_uint64),
                ("stage_attempts", C.c_uint64), ("stage_iters", C.c_uint64), ("stage_alns", C.c_uint64),
                ("pairing_ms", C.c_double), ("pairing_launches", C.c_uint64), ("pairing_mates", C.c_uint64),
                ("pairing_pairs", C.c_uint64),
                ("rescore_ms", C.c_double), ("rescore_launches", C.c_uint64), ("rescore_alns", C.c_uint64),
                ("rescore_ops", C.c_uint64),
                ("recruit_ms", C.c_double), ("recruit_launches", C.c_uint64), ("recruit_bases", C.c_uint64),
                ("recruit_reads", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class DistTimingC(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("collective_ms", C.c_double), ("host_ms", C.c_double),
                ("wall_ms", C.c_double), ("collectives", C.c_uint64), ("gathered_bytes", C.c_uint64),
                ("overflow_rounds", C.c_uint64), ("solves", C.c_uint64)]


class MatesC(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_haps", C.c_uint32), ("max_alns", C.c_uint32), ("ins_len", C.c_uint32),
                ("ma_off", C.c_void_p), ("ma_contig", C.c_void_p), ("ma_flags", C.c_void_p), ("ma_start", C.c_void_p),
                ("ma_end", C.c_void_p), ("ma_ln_prob", C.c_void_p), ("read_weight", C.c_void_p),
                ("ins_ln_pmf", C.c_void_p),
                ("unmapped_penalty", C.c_double), ("insert_penalty", C.c_double), ("prob_diff", C.c_double),
                ("single_end", C.c_uint32), ("window", C.c_uint32), ("exp_off", C.c_void_p), ("exp_weight", C.c_void_p),
                ("read_max_alns", C.c_void_p)]


class PrelimC(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_groups", C.c_uint64), ("read_group", C.c_void_p), ("grp_off", C.c_void_p),
                ("rec_contig", C.c_void_p), ("rec_start", C.c_void_p), ("rec_end", C.c_void_p), ("rec_strand", C.c_void_p),
                ("rec_ln_prob", C.c_void_p), ("grp_ok", C.c_void_p), ("grp_best_edit", C.c_void_p),
                ("grp_thr_dist", C.c_void_p), ("grp_n_kept", C.c_void_p), ("kept_rec", C.c_void_p),
                ("contig_len", C.c_void_p), ("read_weight", C.c_void_p), ("min_weight", C.c_double),
                ("n_haps", C.c_uint32), ("boundary", C.c_uint32), ("single_end", C.c_uint32), ("_pad", C.c_uint32)]


class AlnsC(C.Structure):
    _fields_ = [("n_alns", C.c_uint64), ("cigar_off", C.c_void_p), ("cigar_ops", C.c_void_p), ("aln_start", C.c_void_p),
                ("aln_end", C.c_void_p), ("contig_len", C.c_void_p), ("passable_dist", C.c_void_p),
                ("ln_match", C.c_double), ("ln_mismatch", C.c_double), ("ln_insertion", C.c_double),
                ("ln_deletion", C.c_double), ("ln_clipping", C.c_double)]


class ReadEndsC(C.Structure):
    _fields_ = [("alns", AlnsC), ("n_groups", C.c_uint64), ("grp_off", C.c_void_p), ("rec_contig", C.c_void_p),
                ("grp_read_end", C.c_void_p), ("grp_read_len", C.c_void_p), ("grp_good_dist", C.c_void_p),
                ("grp_passable_dist", C.c_void_p), ("grp_neighb_complexity", C.c_void_p), ("poor_compl", C.c_double),
                ("poor_compl_edit", C.c_double), ("strict_subset", C.c_uint32), ("_pad", C.c_uint32)]


class TargetSeqsC(C.Structure):
    _fields_ = [("n_seqs", C.c_uint64), ("seq_off", C.c_void_p), ("seqs", C.c_void_p), ("seq_locus", C.c_void_p),
                ("cnt_off", C.c_void_p), ("kmer_counts", C.c_void_p), ("base_k", C.c_uint32), ("minimizer_k", C.c_uint32),
                ("minimizer_w", C.c_uint32), ("thresh_kmer_count", C.c_uint32), ("match_frac", C.c_double),
                ("match_length", C.c_uint32), ("_pad", C.c_uint32)]


class ReadsC(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("off1", C.c_void_p), ("seq1", C.c_void_p), ("off2", C.c_void_p),
                ("seq2", C.c_void_p)]


# Every symbol include/lctp.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "lctp_version": (C.c_char_p, []),
    "lctp_last_error": (C.c_char_p, []),
    "lctp_sizeof_locus": (C.c_size_t, []),
    "lctp_sizeof_stage": (C.c_size_t, []),
    "lctp_sizeof_result": (C.c_size_t, []),
    "lctp_init": (C.c_int, [_P, _P]),
    "lctp_destroy": (None, [_P]),
    "lctp_launch_count": (C.c_uint64, [_P]),
    "lctp_sync": (C.c_int, [_P]),
    "lctp_get_stats": (C.c_int, [_P, _P, C.c_int]),
    "lctp_measure_fp64_rate": (C.c_int, [_P, _P]),
    "lctp_prefilter_plan_check": (C.c_int, [C.c_uint32, C.c_uint32, _P, C.c_uint32, C.c_uint64, C.c_uint64, _P, _P, _P]),
    "lctp_pair_alignments": (C.c_int, [_P, _P, C.c_uint64, _P, _P, _P, _P, _P, _P, _P]),
    "lctp_sizeof_mates": (C.c_size_t, []),
    "lctp_pair_alignments_dev": (C.c_int, [_P, _P, _P, _P]),
    "lctp_pairs_fetch": (C.c_int, [_P, C.c_uint64, _P, _P, _P, _P, _P, _P]),
    "lctp_pairs_count": (C.c_uint64, [_P]),
    "lctp_pairs_free": (None, [_P]),
    "lctp_locus_upload_pairs": (C.c_int, [_P, _P, _P, _P]),
    "lctp_rescore_alignments": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "lctp_sizeof_alns": (C.c_size_t, []),
    "lctp_minimizers": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint32, C.c_uint32, _P, _P, _P, _P]),
    "lctp_fraction_approximate_u16": (None, [C.c_double, _P, _P]),
    "lctp_targets_build": (C.c_int, [_P, _P, _P]),
    "lctp_targets_free": (None, [_P]),
    "lctp_targets_entries": (C.c_uint64, [_P, _P, _P, _P, C.c_uint64]),
    "lctp_targets_match_frac": (None, [_P, _P, _P]),
    "lctp_recruit_short": (C.c_int, [_P, _P, _P, C.c_uint32, _P, _P]),
    "lctp_sizeof_target_seqs": (C.c_size_t, []),
    "lctp_sizeof_reads": (C.c_size_t, []),
    "lctp_sizeof_read_ends": (C.c_size_t, []),
    "lctp_collect_read_ends": (C.c_int, [_P] * 12),
    "lctp_sizeof_prelim": (C.c_size_t, []),
    "lctp_group_reads_dev": (C.c_int, [_P] * 7),
    "lctp_mates_count": (C.c_uint64, [_P]),
    "lctp_mates_free": (None, [_P]),
    "lctp_pair_alignments_from": (C.c_int, [_P] * 5),
    "lctp_unique_kmers_build": (C.c_int, [_P, _P, _P, C.c_uint64, _P, _P, C.c_uint32, C.c_uint16, C.c_uint16, _P, _P]),
    "lctp_unique_kmers_count": (C.c_uint64, [_P]),
    "lctp_unique_kmers_free": (None, [_P]),
    "lctp_read_weights": (C.c_int, [_P, _P, _P, _P, C.c_uint64, C.c_uint32, _P, _P]),
    "lctp_counts_to_prob": (C.c_int, [_P, C.c_uint64, C.c_uint16, _P, _P]),
    "lctp_group_reads": (C.c_int, [_P, _P, C.c_uint64] + [_P] * 12),
    "lctp_locus_upload": (C.c_int, [_P, _P, _P]),
    "lctp_locus_free": (None, [_P]),
    "lctp_best_aln_matrix": (C.c_int, [_P, _P]),
    "lctp_prefilter_scores": (C.c_int, [_P, C.c_uint64, C.c_uint64, _P]),
    "lctp_prefilter": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, C.c_size_t, _P, _P]),
    "lctp_truncate_ixs": (C.c_size_t, [_P, C.c_size_t, _P, C.c_double, C.c_size_t, C.c_size_t]),
    "lctp_solve_stage": (C.c_int, [_P, _P, _P, _P, C.c_size_t, _P, _P, _P, _P, _P, _P, C.c_uint64, _P, _P]),
    "lctp_rng_seed_from_u64": (None, [_P, C.c_uint64]),
    "lctp_rng_jump": (None, [_P]),
    "lctp_rng_worker_streams": (None, [_P, C.c_size_t, _P]),
    "lctp_rng_long_jump": (None, [_P]),
    "lctp_plan_stage": (C.c_size_t, [_P, _P, C.c_size_t, C.c_size_t, _P]),
    "lctp_discard_improbable": (C.c_size_t, [_P, C.c_size_t, _P, _P, _P, C.c_double, C.c_size_t, C.c_size_t]),
    "lctp_compare_two_likelihoods": (C.c_double, [C.c_double, C.c_double, C.c_uint16, C.c_double, C.c_double, C.c_uint16]),
    "lctp_build_depth_table": (C.c_int, [_P, _P, C.c_int, _P, C.c_size_t, C.c_uint32, _P]),
    "lctp_find_weighted_dist": (C.c_int, [_P, _P, _P, C.c_int]),
    "lctp_produce_result": (C.c_int, [_P, _P, C.c_size_t, _P, _P, _P, _P]),
    "lctp_solve": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, _P, _P]),
    "lctp_solve_counts": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, _P, _P, C.c_size_t, _P, _P, C.c_uint64]),
    "lctp_locus_wmax": (C.c_uint32, [_P]),
    "lctp_solve_stage_dbg": (C.c_int, [_P, _P, _P, _P, C.c_size_t, _P, _P, _P, _P, _P, _P, C.c_uint64, _P, _P, _P]),
    "lctp_debug_open": (C.c_int, [_P, C.c_char_p, C.c_int, _P, C.c_size_t]),
    "lctp_debug_close": (None, [_P]),
    "lctp_result_json": (C.c_size_t, [_P, _P, _P, _P, C.c_size_t]),
    "lctp_dist_unique_id": (C.c_int, [_P]),
    "lctp_dist_init": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "lctp_dist_destroy": (None, [_P]),
    "lctp_dist_rank": (C.c_int, [_P]),
    "lctp_dist_world": (C.c_int, [_P]),
    "lctp_dist_get_timing": (C.c_int, [_P, _P, C.c_int]),
    "lctp_dist_prefilter": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, _P, C.c_size_t, _P]),
    "lctp_dist_solve_stage": (C.c_int, [_P, _P, _P, _P, _P, C.c_size_t, _P, _P, _P]),
    "lctp_dist_solve": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_size_t, _P, _P]),
}

_lib = None


def build(force: bool = False) -> str:
    """Compile the CUDA extension in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER]
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", CSRC, "-j4"] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LctpError(E_CUDA, f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; "
                                    "g.build()'`); the genotype-evaluation path has no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        assert lib.lctp_sizeof_locus() == C.sizeof(LocusC)
        assert lib.lctp_sizeof_stage() == C.sizeof(StageC)
        assert lib.lctp_sizeof_result() == C.sizeof(ResultC)
        assert lib.lctp_sizeof_mates() == C.sizeof(MatesC)
        assert lib.lctp_sizeof_alns() == C.sizeof(AlnsC)
        assert lib.lctp_sizeof_read_ends() == C.sizeof(ReadEndsC)
        assert lib.lctp_sizeof_prelim() == C.sizeof(PrelimC)
        assert lib.lctp_sizeof_target_seqs() == C.sizeof(TargetSeqsC) and lib.lctp_sizeof_reads() == C.sizeof(ReadsC)
        _lib = lib
    return _lib


def check(rc: int):
    if rc != OK:
        raise LctpError(rc, load().lctp_last_error().decode())

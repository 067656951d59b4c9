"""`.lcti` dumps: the flat per-locus input as a directory of raw little-endian arrays plus
`meta.json` -- the format `FlatLocus::dump` (rust/gpu.rs) writes from inside the reference and this module reads and
writes from Python.  Used by oracle/rust_diff.sh (entry point: tools/rust_diff.sh) to run the oracle / the CUDA path on exactly the
`solve::Data` a real `locityper genotype` run saw.

    python tools/lcti.py write DIR --config C1 --seed 1001     # synthetic locus -> dump (round-trip / demo)
    python tools/lcti.py solve DIR --threads 8 --out OUT        # through liblctp (needs a B200): OUT/res.json
The oracle-side counterpart (sol.csv / sol_ext.csv in the reference's formats) is oracle/lcti_solve.py.
"""
import argparse
import ctypes as C
import json
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from locityper_b200 import synth  # noqa: E402

_ARRAYS = {"unmapped_prob": "f64", "pa_off": "u64", "pa_contig": "u32", "pa_ln_prob": "f64", "pa_mid1": "u32",
           "pa_mid2": "u32", "hap_len": "u32", "hap_n_windows": "u32", "hap_reg_start": "u32", "hap_pos_off": "u64",
           "pos_weight": "f64", "pos_gc": "u8", "depth_table": "f64", "gt_tuples": "u32", "priors": "f64"}
_DT = {"f64": "<f8", "u64": "<u8", "u32": "<u4", "u8": "u1"}
_F64_PARAMS = ("prob_diff", "lik_skew", "min_weight", "filt_diff", "prob_thresh")


def _bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def _from_bits(b: int) -> float:
    return struct.unpack("<d", struct.pack("<Q", int(b)))[0]


def write(loc: synth.Locus, path: str, hap_names=None, rng_state=None) -> None:
    os.makedirs(path, exist_ok=True)
    for name, ty in _ARRAYS.items():
        a = getattr(loc, name)
        if a is None:
            continue
        np.ascontiguousarray(a, dtype=_DT[ty]).tofile(os.path.join(path, f"{name}.{ty}"))
    meta = dict(n_haps=loc.n_haps, n_reads=loc.n_reads, ploidy=loc.ploidy, is_paired=bool(loc.is_paired),
                n_genotypes=loc.n_genotypes, window=loc.window, left_padding=loc.left_padding, depth_k=loc.depth_k,
                tweak=loc.tweak, dont_skip=bool(loc.dont_skip), out_bams=loc.out_bams,
                hap_names=list(hap_names) if hap_names is not None else [f"hap{i}" for i in range(loc.n_haps)])
    for k in _F64_PARAMS:
        meta[k + "_bits"] = _bits(getattr(loc, k))
    with open(os.path.join(path, "meta.json"), "w") as f:
        json.dump(meta, f)
    if rng_state is not None:
        np.asarray(rng_state, dtype="<u8").tofile(os.path.join(path, "rng_state.u64"))


def read(path: str):
    """-> (Locus, hap_names, rng_state or None)"""
    with open(os.path.join(path, "meta.json")) as f:
        meta = json.load(f)
    arr = {}
    for name, ty in _ARRAYS.items():
        fn = os.path.join(path, f"{name}.{ty}")
        arr[name] = np.fromfile(fn, dtype=_DT[ty]) if os.path.exists(fn) else None
    kw = {k: _from_bits(meta[k + "_bits"]) for k in _F64_PARAMS}
    loc = synth.Locus(
        n_haps=meta["n_haps"], n_reads=meta["n_reads"], ploidy=meta["ploidy"], is_paired=bool(meta["is_paired"]),
        window=meta["window"], left_padding=meta["left_padding"], depth_k=meta["depth_k"], tweak=meta["tweak"],
        dont_skip=bool(meta["dont_skip"]), out_bams=meta["out_bams"],
        nb_n=np.zeros(101), nb_p=np.zeros(101), alt_cn=np.zeros(0),     # the table itself is in the dump
        **{k: v for k, v in arr.items()}, **kw)
    assert loc.n_genotypes == meta["n_genotypes"], "genotype list does not match n_genotypes"
    assert len(loc.depth_table) == 101 * loc.depth_k
    st = os.path.join(path, "rng_state.u64")
    rng = np.fromfile(st, dtype="<u8") if os.path.exists(st) else None
    return loc, meta["hap_names"], rng


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    w = sub.add_parser("write")
    w.add_argument("dir"); w.add_argument("--config", default="C1"); w.add_argument("--seed", type=int, default=1001)
    s = sub.add_parser("solve")
    s.add_argument("dir"); s.add_argument("--threads", type=int, default=8); s.add_argument("--out", required=True)
    s.add_argument("--seed", type=int, default=None, help="-s SEED of the run when the dump carries no rng_state.u64")
    s.add_argument("--scheme", nargs="*", default=["greedy:i=5k,a=1", "anneal:i=20,a=20"])
    a = ap.parse_args()
    from locityper_b200 import genotype
    if a.cmd == "write":
        loc = synth.make_locus(**synth.config_shape(a.config), seed=a.seed, table_builder=genotype.build_depth_table)
        write(loc, a.dir, rng_state=genotype.init_rng(a.seed))
        print(f"wrote {a.dir}: H={loc.n_haps} R={loc.n_reads} G={loc.n_genotypes}")
        return
    loc, names, st = read(a.dir)
    os.makedirs(a.out, exist_ok=True)
    rng = np.array(st, dtype=np.uint64) if st is not None else genotype.init_rng(a.seed)
    ctx = genotype.Context(0)
    dl = ctx.upload(loc)
    got = dl.solve(genotype.Scheme.parse(a.scheme), a.threads, rng, hap_names=names)
    with open(os.path.join(a.out, "res.json"), "w") as f:
        f.write(got.json_text)
    print("call:", got.to_json()["genotype"], "quality", got.quality)


if __name__ == "__main__":
    main()

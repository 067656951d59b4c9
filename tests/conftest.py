import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import lcto_py
    lcto_py.lib()
    return lcto_py


@pytest.fixture(scope="session")
def small_locus(oracle):
    """C1-like but small: 24 haplotypes (300 genotypes), 300 read pairs, 2.5 kb."""
    from locityper_b200 import synth
    return synth.make_locus(24, 300, 2500, seed=11, table_builder=oracle.build_depth_table)


@pytest.fixture(scope="session")
def gpu_ctx():
    from locityper_b200 import genotype
    ctx = genotype.Context(device=0)
    yield ctx
    ctx.close()


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden_mates():
    """(genotype.Mates, expected outputs) of tests/golden/pairs_small.npz (made by tests/golden/make_golden.py)."""
    import numpy as np
    from locityper_b200 import genotype
    z = np.load(os.path.join(GOLDEN_DIR, "pairs_small.npz"))
    sc = z["_scalars"]
    m = genotype.Mates(n_reads=int(sc[0]), n_haps=int(sc[1]), max_alns=int(sc[2]), unmapped_penalty=float(sc[3]),
                       insert_penalty=float(sc[4]), prob_diff=float(sc[5]), ma_off=z["ma_off"], ma_contig=z["ma_contig"],
                       ma_flags=z["ma_flags"], ma_start=z["ma_start"], ma_end=z["ma_end"], ma_ln_prob=z["ma_ln_prob"],
                       ins_ln_pmf=z["ins_ln_pmf"], read_weight=z["read_weight"])
    return m, {k[4:]: z[k] for k in z.files if k.startswith("out_")}


def load_golden_alns():
    """(genotype.Alns, expected outputs) of tests/golden/rescore_small.npz."""
    import numpy as np
    from locityper_b200 import genotype
    z = np.load(os.path.join(GOLDEN_DIR, "rescore_small.npz"))
    a = genotype.Alns(cigar_off=z["cigar_off"], cigar_ops=z["cigar_ops"], aln_start=z["aln_start"], aln_end=z["aln_end"],
                      contig_len=z["contig_len"], passable_dist=z["passable_dist"], ln_oper=tuple(z["ln_oper"]))
    return a, {k[4:]: z[k] for k in z.files if k.startswith("out_")}


def load_golden_prelim(tag):
    """(genotype.Prelim, expected outputs) of tests/golden/group_{pe,se}_small.npz."""
    import numpy as np
    from locityper_b200 import genotype
    z = np.load(os.path.join(GOLDEN_DIR, f"group_{tag}_small.npz"))
    sc = z["_scalars"]
    fields = ("read_group", "grp_off", "rec_contig", "rec_start", "rec_end", "rec_strand", "rec_ln_prob", "grp_ok",
              "grp_best_edit", "grp_thr_dist", "grp_n_kept", "kept_rec", "contig_len", "read_weight")
    p = genotype.Prelim(**{k: z[k] for k in fields}, min_weight=float(sc[0]), boundary=int(sc[1]), single_end=bool(sc[2]))
    out = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    out["n_reads_out"] = int(out["n_reads_out"])
    return p, out

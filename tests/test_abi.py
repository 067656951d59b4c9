"""The C-ABI shared library loads on a CPU-only host and exports every symbol include/lctp.h declares;
compute entry points fail loudly (LCTP_E_CUDA, no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from locityper_b200 import ffi, genotype

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lctp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lctp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    ffi.build()
    lib = C.CDLL(ffi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lctp.h but not exported"
        assert n in ffi.SYMBOLS, f"{n} has no ctypes prototype in locityper_b200/ffi.py"
    assert set(ffi.SYMBOLS) <= set(names)
    L = ffi.load()
    assert b"sm_100a" in L.lctp_version()
    assert L.lctp_sizeof_locus() == C.sizeof(ffi.LocusC)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for CPU-only hosts")
    with pytest.raises(ffi.LctpError) as e:
        genotype.Context(device=0)
    assert e.value.code == ffi.E_CUDA and "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "locityper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "lcto" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_only_tests_smoke_and_the_bench_cpu_arm_touch_the_oracle():
    """oracle/ is test infrastructure: besides tests/, only __graft_entry__.smoke()/build() and bench.py's CPU arm
    (cpu_baseline / --impl reference) may import it; the developer tools under tools/ must not."""
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith((".py", ".sh")):
            text = open(os.path.join(ROOT, "tools", f)).read()
            assert "lcto" not in text and "from oracle" not in text and "import oracle" not in text, f
    bench = open(os.path.join(ROOT, "bench.py")).read()
    for i, line in enumerate(bench.split("\n")):
        if "from oracle" in line or "import oracle" in line:
            # every import of the oracle sits inside the CPU arm: cpu_run() (one oracle pass, used by cpu_baseline
            # and by --impl reference) or run_reference() (the --impl reference leg itself)
            head = bench.split("\n")[:i]
            fn = [l for l in head if l.startswith("def ")][-1]
            assert fn.startswith(("def cpu_run(", "def run_reference(")), (i, fn)


def test_invalid_arguments_return_error_codes():
    L = ffi.load()
    assert L.lctp_init(None, None) == ffi.E_INVALID
    assert b"out is NULL" in L.lctp_last_error()
    ixs = np.zeros(0, dtype=np.uint64)
    assert L.lctp_truncate_ixs(ixs.ctypes.data, 0, None, 1.0, 1, 1) == 0
    assert L.lctp_solve(None, None, 0, 1, None, None) == ffi.E_INVALID
    assert L.lctp_produce_result(None, None, 0, None, None, None, None) == ffi.E_INVALID

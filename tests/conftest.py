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

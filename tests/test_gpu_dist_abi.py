"""The north-star multi-GPU split INSIDE the library (lctp_dist_*, csrc/dist.cu): genotype ranges in the prefilter with
device-side candidate selection and a fixed-capacity ncclAllGather, logical workers modulo the world size in the stages.
Every check goes through the C ABI; results must equal the single-GPU calls bit for bit on every rank.

world = 1 runs on the single test GPU (a one-rank NCCL communicator exercises the whole path); world = 2 / 4 / 8 need that
many GPUs and are skipped otherwise (run them with `gpurun --gpus N -- python -m pytest tests/test_gpu_dist_abi.py`)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from locityper_b200 import genotype, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dist_world1_equals_single_gpu(oracle, gpu_ctx, small_locus):
    d = genotype.Dist(gpu_ctx, genotype.dist_unique_id(), 0, 1)
    try:
        dl = gpu_ctx.upload(small_locus)
        scheme = genotype.Scheme.parse(["greedy:i=100,a=1", "anneal:i=10,a=5,n=2000,p=1000"])
        for threads in (1, 8, 64):
            r1, r2 = genotype.init_rng(99), genotype.init_rng(99)
            got = d.solve(dl, scheme, threads, r1)
            ref = dl.solve(scheme, threads, r2)
            assert np.array_equal(got.gt_ix, ref.gt_ix) and np.array_equal(got.lik_mean, ref.lik_mean)
            assert np.array_equal(got.ln_prob, ref.ln_prob) and list(r1) == list(r2)
            assert got.json_text == ref.json_text
            for min_size in (10, 50, 10 ** 6):
                assert np.array_equal(d.prefilter(dl, min_size, threads), dl.prefilter(min_size, threads))
        # and against the oracle
        ol = oracle.OracleLocus(small_locus)
        ro = oracle.Rng.from_seed(99)
        o = oracle.solve(ol, [oracle.Stage("greedy", attempts=1, in_size=100),
                              oracle.Stage("anneal", attempts=5, in_size=10, anneal_steps=2000, plato_size=1000)], 64, ro,
                         os_threads=4)
        r1 = genotype.init_rng(99)
        got = d.solve(dl, scheme, 64, r1)
        assert np.array_equal(got.gt_ix, o["gt_ix"]) and list(r1) == ro.state()
        t = d.timing()
        assert t["collectives"] > 0 and t["solves"] >= 4
        dl.free()
    finally:
        d.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dist_multi_gpu_through_the_c_abi(world, tmp_path):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = str(tmp_path / "dist.json")
    port = 29650 + world
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "dist_abi_worker.py"), out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    assert res["ok"] and res["world"] == world

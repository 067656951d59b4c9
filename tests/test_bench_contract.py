"""bench.py contract (driver-facing): the CPU arm runs without a GPU and prints one JSON line with the required keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_valid_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--config", "C1", "--loci", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("genotypes scored/sec") and d["unit"] == "genotypes/s"
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "genotypes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("BASELINE configs[0]")


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_rooflines_and_rates_are_pure_functions_of_the_stats():
    """bench.rooflines / kernel_rates only need the library's stats dictionary: exercised here without a GPU."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    st = dict(prefilter_ms=0.24, prefilter_launches=3, prefilter_genotypes=135450, stage_ms=30.4, stage_launches=3,
              stage_genotypes=15000, stage_attempts=15000, stage_iters=22_000_000, stage_alns=160_000_000,
              pairing_ms=0.0, pairing_launches=0, pairing_mates=0, pairing_pairs=0)
    loc = types.SimpleNamespace(n_reads=2000, ploidy=2)
    args = types.SimpleNamespace(config="C2", scheme=["greedy:i=5k,a=1"])
    out = bench.rooflines(st, [loc], args, 6550.7, "measured", fp64_rate=18.3e12)
    assert set(out) >= {"roofline", "roofline_prefilter", "rates"}
    assert 0 < out["roofline"]["frac"] < 1 and out["roofline"]["unit"] == "GB/s"
    iss = out["roofline"].get("issue")       # from the committed ncu capture: warp-instructions per launch / measured time
    assert iss is None or (0 < iss["frac"] < 1 and iss["warp_instructions_per_launch"] > 1e9)
    r = out["rates"]
    assert abs(r["stage_genotypes_per_s"] - 15000 / 0.0304) < 1e-6 * r["stage_genotypes_per_s"]
    assert abs(r["prefilter_genotypes_per_s"] - 135450 / 0.00024) < 1e-6 * r["prefilter_genotypes_per_s"]
    assert r["stage_genotype_attempts_per_s"] == r["stage_genotypes_per_s"] and r["stage_iterations_per_s"] > 1e8
    assert bench.kernel_rates(dict(prefilter_launches=0, stage_launches=0)) == {}

#!/usr/bin/env python
"""bench.py -- genotypes scored/sec (prefilter + solver) on BASELINE.json's configs.

  python bench.py --gpus N --steps K --warmup W            (ours: sm_100a CUDA through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  (CPU arm: the reference's algorithm on host cores)

A *step* is one pass of the hot path (prefilter + greedy stage) over one batch of synthetic loci of the
shape BASELINE.json quotes the metric on: configs[1] = "HLA class I panel (HLA-A/B/C) ~300 haplotypes each,
45k diploid genotypes per locus, 30x short reads, prefilter + greedy on 1 B200"  ->  3 loci per step per GPU.
N > 1 (torchrun, one rank per GPU): loci are independent (src/command/genotype.rs:1331-1351), so every rank
processes its own 3 loci (weak scaling, no data-path collective); the calls are all-gathered over NCCL at the
end.  `--mode shard` runs the KIR-scale configs[3] instead: one locus whose genotype list is partitioned
across ranks with an NCCL all-gather of per-GPU top-k (score, id) candidates.

The reference (Rust) cannot be built in this image, so the CPU arm is the oracle port (oracle/, plain C,
pthreads) on all host cores: `cpu_baseline.kind = "port"`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genotypes scored/sec (prefilter+solver)"
UNIT = "genotypes/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="loci", choices=["loci", "shard"])
    ap.add_argument("--config", default="C2", help="shape of the loci (C1..C5); C2 is the metric's config")
    ap.add_argument("--loci", type=int, default=3, help="loci per step per GPU")
    ap.add_argument("--threads", type=int, default=0,
                    help="reference -@ T = number of logical workers / RNG streams (0 = auto: one per survivor "
                         "slot that fits the GPU, same T for the CPU arm)")
    ap.add_argument("--scheme", nargs="*", default=["greedy:i=5k,a=1"],
                    help="reference -S stages; configs[1] is 'prefilter + greedy'")
    ap.add_argument("--streams", type=int, default=0,
                    help="loci in flight per GPU (one context + CUDA stream + host thread each); 0 = loci per step")
    ap.add_argument("--max-resident", type=int, default=0,
                    help="resident logical workers per context (solver grid cap); 0 = every context may fill the device")
    ap.add_argument("--seed", type=int, default=2001)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kir-prefilter", action="store_true",
                    help="skip the extra prefilter-only measurement at the KIR-scale shape (configs[3])")
    ap.add_argument("--cpu-loci", type=int, default=0, help="loci in the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-shard-kir", action="store_true", help="skip the sharded KIR-scale solve (configs[3])")
    ap.add_argument("--kir-threads", type=int, default=5000,
                    help="-@ of the sharded KIR-scale solve: one logical worker per surviving genotype of its greedy:i=5k stage, so\n"
                         "that the longest chain of a rank is one genotype (0 = same as --threads)")
    ap.add_argument("--no-t-sweep", action="store_true", help="skip the T = 8 / 64 / bench-T rows")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def auto_threads(args) -> int:
    # One logical worker per resident warp slot: 148 SMs x 32 warps.  Also a legal reference `-@` (u16).
    return args.threads if args.threads > 0 else 4736


def make_loci(args, rank: int, table_builder, n: int):
    from locityper_b200 import synth
    shape = synth.config_shape(args.config)
    loci = []
    for i in range(n):
        seed = args.seed + 1000 * rank + i
        loci.append(synth.make_locus(**shape, seed=seed, table_builder=table_builder))
    return loci


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def pinned_locus(loc):
    """Move every input array of the locus into pinned host memory (e2e copies start from pinned memory)."""
    import torch
    keep = []
    for name in ("unmapped_prob", "pa_off", "pa_contig", "pa_ln_prob", "pa_mid1", "pa_mid2", "hap_len",
                 "hap_n_windows", "hap_reg_start", "hap_pos_off", "pos_weight", "pos_gc", "depth_table"):
        a = np.ascontiguousarray(getattr(loc, name))
        t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).pin_memory()
        keep.append(t)
        setattr(loc, name, t.numpy().view(a.dtype).reshape(a.shape))
    loc.meta["_pinned"] = keep
    return loc


def input_bytes(loc) -> int:
    return int(sum(np.asarray(getattr(loc, n)).nbytes for n in (
        "unmapped_prob", "pa_off", "pa_contig", "pa_ln_prob", "pa_mid1", "pa_mid2", "hap_len", "hap_n_windows",
        "hap_reg_start", "hap_pos_off", "pos_weight", "pos_gc", "depth_table")))


MAX_SM_MHZ = [1965.0]      # replaced by the sampled nvidia-smi max SM clock of the run


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU arm

def cpu_run(args, loci, T, os_threads):
    """One pass of the oracle port over `loci`; returns (genotypes, seconds)."""
    from oracle import lcto_py as O
    scheme = []
    from locityper_b200.genotype import Scheme
    for st in Scheme.parse(args.scheme).stages:
        scheme.append(O.Stage(st.kind, attempts=st.attempts, in_size=st.in_size, best_start=st.best_start,
                              sample_size=st.sample_size, plato_size=st.plato_size, anneal_steps=st.anneal_steps,
                              init_prob=st.init_prob))
    ols = [O.OracleLocus(l) for l in loci]
    t0 = time.perf_counter()
    g = 0
    for i, ol in enumerate(ols):
        rng = O.Rng.from_seed(args.seed + i)
        r = O.solve(ol, scheme, T, rng, os_threads=os_threads)
        CPU_SPLIT[0] += r["t_prefilter_s"]
        CPU_SPLIT[1] += r["t_stages_s"]
        g += ol.loc.n_genotypes
    return g, time.perf_counter() - t0


CPU_SPLIT = [0.0, 0.0]      # seconds the CPU arm spent in the (single-threaded) prefilter / in the threaded stages


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import lcto_py as O
    O.lib()
    cores = os.cpu_count() or 1
    T = auto_threads(args)
    # table builder: the oracle's own (the CPU arm must not need the CUDA library)
    loci = make_loci(args, 0, O.build_depth_table, args.loci)
    for _ in range(args.warmup):
        cpu_run(args, loci[:1], T, cores)
    times, g = [], 0
    for _ in range(args.steps):
        g, dt = cpu_run(args, loci, T, cores)
        times.append(dt)
    total = sum(times)
    value = g * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, T, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.loci} loci x {args.steps} steps, T={T} logical workers on {cores} pthreads "
                                   "(prefilter single-threaded like the reference)",
                         "prefilter_share": CPU_SPLIT[0] / max(1e-12, CPU_SPLIT[0] + CPU_SPLIT[1])},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Rust and cannot be built here (no cargo); this is the plain-C oracle port of the same path",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, T, world):
    from locityper_b200 import synth
    sh = synth.config_shape(args.config)
    G = sh["n_haps"] * (sh["n_haps"] + 1) // 2
    cfg_ix = {"C1": 0, "C2": 1, "C3": 2, "C4": 3, "C5": 4}.get(args.config, 1)
    return {"workload": f"BASELINE configs[{cfg_ix}]: {args.loci if args.mode == 'loci' else 1} loci/step/GPU x H={sh['n_haps']} haplotypes "
                        f"(G={G} diploid genotypes), R={sh['n_reads']} read pairs, {sh['locus_len']} bp, "
                        f"{sh['tech']}; scheme {' '.join('-S ' + s for s in args.scheme)}",
            "shape": args.config, "loci_per_step_per_gpu": args.loci, "threads_T": T, "mode": args.mode,
            "parallelism": f"loci x{world}" if args.mode == "loci" else f"genotype-shard x{world}",
            "loci_in_flight_per_gpu": (args.streams if args.streams > 0 else args.loci) if args.mode == "loci" else 1,
            "l2": "an L2 flush buffer (256 MB > 126 MB) is written between timed steps; inputs (3 loci x ~26 MB) are "
                  "re-read from HBM every step"}


# ------------------------------------------------------------------------------------------------ GPU arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    from locityper_b200 import genotype

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the genotype-evaluation path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream(dev)
    ctx = genotype.Context(device=local, stream=stream.cuda_stream)
    T = auto_threads(args)
    scheme = genotype.Scheme.parse(args.scheme)

    if args.mode == "shard":
        return run_shard(args, ctx, scheme, T, rank, world, dev)

    loci = [pinned_locus(l) for l in make_loci(args, rank, genotype.build_depth_table, args.loci)]
    G_step = sum(l.n_genotypes for l in loci)
    h2d = sum(input_bytes(l) for l in loci)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # The loci of a step are independent (genotype.rs:1331-1351): each is solved by its own context (own
    # CUDA stream, own host thread), so kernels, copies and host-side pruning of different loci overlap.
    n_streams = args.streams if args.streams > 0 else len(loci)
    # Loci in flight share the SMs: the stage kernels of the contexts run side by side, a finishing kernel's free
    # worker slots are taken by the next one.  --max-resident caps the slots one context may take (0 = no cap).
    max_res = max(0, args.max_resident)
    pool = genotype.ContextPool(device=local, k=n_streams, max_resident_workers=max_res)
    fp64_rate = ctx.fp64_rate()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def l2_flush():
        flush.zero_()                         # L2 flush (256 MB > 126 MB), outside the event pair
        torch.cuda.synchronize(dev)           # the loci run on other streams: do not let them overlap the flush

    def step_resident(dls):
        return pool.map(lambda c, i, dl: dl.solve(scheme, T, genotype.init_rng(args.seed + i)), dls)

    def one_e2e(c, i, loc):
        dl = c.upload(loc)                    # H2D of the whole flat locus from pinned host memory
        res = dl.solve(scheme, T, genotype.init_rng(args.seed + i))   # D2H of (lik_mean, lik_var, RNG states, calls)
        dl.free()
        return res

    def step_e2e():
        return pool.map(one_e2e, loci)

    # ---- value: inputs resident in HBM when the timed region starts
    dls = pool.map(lambda c, i, l: c.upload(l), loci)
    for _ in range(args.warmup):
        step_resident(dls)
    pool.stats(reset=True)
    launches0 = pool.launch_count()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    calls = None
    for k in range(args.steps):
        l2_flush()
        ev[k][0].record(stream)
        calls = step_resident(dls)            # returns when every locus' results are on the host
        ev[k][1].record(stream)
    barrier()
    clocks = sampler.stop()
    if clocks.get("sm_max_mhz"):
        MAX_SM_MHZ[0] = float(clocks["sm_max_mhz"])
    launches = pool.launch_count() - launches0
    st = pool.stats(reset=True)
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    t_total = torch.tensor([sum(ms_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    ms_total = float(t_total.item())
    value = world * G_step * args.steps / (ms_total / 1e3)
    for dl in dls:
        dl.free()

    # ---- kernel-level numbers: the same loci once more, one at a time on one stream, so that every launch's
    # CUDA-event duration is the kernel alone (launches of concurrent loci share the SMs and their event
    # durations overlap); the roofline objects are computed from this pass.
    iso = [ctx.upload(l) for l in loci]
    for i, dl in enumerate(iso):
        dl.solve(scheme, T, genotype.init_rng(args.seed + i))
    ctx.stats(reset=True)
    l2_flush()
    for i, dl in enumerate(iso):
        dl.solve(scheme, T, genotype.init_rng(args.seed + i))
    st_iso = ctx.stats(reset=True)
    for dl in iso:
        dl.free()

    # ---- e2e: host buffers in, host results out, through the public API, copies inside the timed region
    for _ in range(max(1, args.warmup - 1)):
        step_e2e()
    pool.stats(reset=True)
    barrier()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        l2_flush()
        ev2[k][0].record(stream)
        calls = step_e2e()
        ev2[k][1].record(stream)
    barrier()
    t2 = torch.tensor([sum(a.elapsed_time(b) for a, b in ev2)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * G_step * args.steps / (float(t2.item()) / 1e3)
    st_e2e = pool.stats(reset=True)             # the library counts the bytes of every copy it issues
    h2d = st_e2e["h2d_bytes"] // args.steps
    d2h = st_e2e["d2h_bytes"] // args.steps

    # all ranks exchange their calls (tiny) -- the only cross-GPU traffic of the loci mode
    my = torch.tensor([int(c.gt_ix[0]) for c in calls], dtype=torch.int64, device=dev)
    if world > 1:
        allc = [torch.empty_like(my) for _ in range(world)]
        dist.all_gather(allc, my)

    if rank == 0:
        peak, peak_src = load_peaks()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, T, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "loci_per_s": world * len(loci) * args.steps / (ms_total / 1e3),
        }
        line.update(rooflines(st_iso, loci, args, peak, peak_src, fp64_rate))
        line["roofline"]["in_step"] = {
            "note": "all launches of the timed region (concurrent loci share the SMs): algorithmic bytes / step time",
            "achieved": line["roofline"]["bytes_alg_per_launch"] * st["stage_launches"] / (ms_total / 1e3) / 1e9,
            "launches": int(st["stage_launches"]), "sum_launch_ms": st["stage_ms"]}
        line["calls_vs_truth"] = [[list(l.genotype_tuple(int(c.gt_ix[0]))), list(l.truth)] for l, c in zip(loci, calls)]
        if not args.no_kir_prefilter and world == 1:
            line["roofline_prefilter_kir"] = kir_prefilter(ctx, genotype, peak, fp64_rate)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args, loci, T)
            if not args.no_t_sweep:
                line["t_sweep"] = t_sweep(args, ctx, genotype, loci[0])
    # ---- configs[3]: the KIR-scale locus sharded over the N GPUs INSIDE the library (lctp_dist_*), at every N
    kir = None if args.no_shard_kir else shard_kir(args, ctx, genotype, rank, world, dev)
    if rank == 0:
        if kir is not None:
            line["shard_kir"] = kir
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def t_sweep(args, ctx, genotype, loc):
    """`-@ T` changes results in the reference (survivor floors solve.rs:80,433, chunking :1057, RNG streams) and it is
    what gives the GPU its parallel work: one C2 locus at the reference's default T = 8, at 64 and at the bench's T, both
    arms (GPU: locus resident, CUDA events around lctp_solve; CPU: the oracle port on min(T, cores) pthreads)."""
    import torch
    cores = os.cpu_count() or 1
    scheme = genotype.Scheme.parse(args.scheme)
    dl = ctx.upload(loc)
    rows = []
    for T in (8, 64, auto_threads(args)):
        dl.solve(scheme, T, genotype.init_rng(args.seed))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream())
        reps = 3
        for _ in range(reps):
            res = dl.solve(scheme, T, genotype.init_rng(args.seed))
        b.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        g, dt = cpu_run(args, [loc], T, min(T, cores))
        rows.append({"T": T, "gpu_ms_per_locus": ms, "gpu_genotypes_per_s": loc.n_genotypes / (ms / 1e3),
                     "cpu_genotypes_per_s": g / dt, "cpu_threads": min(T, cores),
                     "call": list(loc.genotype_tuple(int(res.gt_ix[0]))), "truth": list(loc.truth)})
    dl.free()
    return rows


def shard_kir(args, ctx, genotype, rank, world, dev):
    """BASELINE configs[3]: one KIR-scale locus (H = 1,000, 500,500 genotypes, 10,000 read pairs), genotype list sharded
    over the `world` GPUs by the library itself: lctp_dist_solve = prefilter by id ranges + device-side candidate selection
    + fixed-capacity ncclAllGather, greedy stage by logical workers modulo world + ncclAllGather (csrc/dist.cu).
    Strong scaling: the same solve at every N, including N = 1."""
    import torch
    import torch.distributed as dist
    from locityper_b200 import synth
    T = args.kir_threads if args.kir_threads > 0 else auto_threads(args)
    loc = synth.make_locus(**synth.config_shape("C4"), seed=4001, table_builder=genotype.build_depth_table)
    dl = ctx.upload(loc)
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt = torch.frombuffer(bytearray(genotype.dist_unique_id()), dtype=torch.uint8).to(dev)
    if world > 1:
        dist.broadcast(idt, 0)
    d = genotype.Dist(ctx, bytes(idt.cpu().numpy().tobytes()), rank, world)
    scheme = genotype.Scheme.parse(["greedy:i=5k,a=1"])
    stream = torch.cuda.current_stream(dev)
    d.solve(dl, scheme, T, genotype.init_rng(4001))
    d.timing(reset=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    reps = 3
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        res = d.solve(dl, scheme, T, genotype.init_rng(4001))
    b.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    tm = d.timing()
    parts = torch.tensor([tm["kernel_ms"] / reps, tm["collective_ms"] / reps, tm["host_ms"] / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(parts, op=dist.ReduceOp.MAX)
    d.close()
    dl.free()
    ms = float(t.item())
    return {"workload": "configs[3]: H=1000, G=500500, R=10000, one locus sharded by lctp_dist_solve, scheme greedy:i=5k,a=1",
            "n_gpus": world, "threads_T": T, "ms_per_solve": ms, "genotypes_per_s": loc.n_genotypes / (ms / 1e3),
            "kernel_ms": float(parts[0]), "collective_ms": float(parts[1]), "host_ms": float(parts[2]),
            "collectives_per_solve": tm["collectives"] / reps, "gathered_bytes_per_solve": tm["gathered_bytes"] / reps,
            "note": "max over ranks of each part (CUDA events for kernels and ncclAllGather, wall clock for the rest)",
            "call": list(loc.genotype_tuple(int(res.gt_ix[0]))), "truth": list(loc.truth), "scaling": "strong"}


def ncu_traffic(kernel, full=False):
    """Per-launch numbers of `kernel` from the committed ncu captures (profiles/r02_ncu_traffic.json, written by
    tools/ncu_summary.py; round-1 file as fallback), or None."""
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                e = json.load(f)[kernel]
            return e if full else {"dram_bytes_per_launch": e["dram_bytes_per_launch"], "capture": e["capture"]}
        except Exception:
            continue
    return None


def rooflines(st, loci, args, peak, peak_src, fp64_rate=None):
    """roofline of the dominant kernel (the solver stage) + the prefilter kernel, from device-side timings."""
    loc = loci[0]
    R, p = loc.n_reads, loc.ploidy
    out = {}
    # solver stage kernel.  Algorithmic bytes = the COMPULSORY traffic of a launch: per genotype the
    # shared arrays read once (CSR offsets of its p haplotypes 8 B per read and haplotype, ln-probability 8 B and
    # pair-alignment middles 8 B per candidate), the worker's candidate records written and read once per attempt
    # (4 B each way) and the pre-generated draws written and read once (16 B per draw; draws = one per candidate and
    # attempt in the tweak + 10 per greedy iteration / ~2 per annealing step).  The kernel is a set of dependent
    # chains, not a streaming kernel: the number that describes it is the issue-slot utilisation (`issue`).
    if st["stage_launches"]:
        g, A, it, att = st["stage_genotypes"], st["stage_alns"], st["stage_iters"], st["stage_attempts"]
        per_iter_draws = 10 if "greedy" in " ".join(args.scheme) else 2
        apg = A / max(1, g)                                   # candidates per genotype
        draws = att * apg + it * per_iter_draws
        bytes_alg = g * R * p * 8 + A * 16 + att * apg * 8 + draws * 16
        sec = st["stage_ms"] / 1e3
        ach = bytes_alg / sec / 1e9
        # committed capture of the same launch: C2 greedy (the default line) or C3 annealing (`--config C3 --scheme
        # anneal:i=5k,a=20`, tools/final_run.sh)
        is_anneal = "anneal" in " ".join(args.scheme)
        cap = (ncu_traffic("k_solve_stage", full=True) if args.config == "C2" and not is_anneal else
               ncu_traffic("k_solve_stage_anneal_c3", full=True) if args.config == "C3" and is_anneal else None)
        out["roofline"] = {"kernel": "k_solve_stage", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                           "frac": ach / peak,
                           "traffic": ({"dram_bytes_per_launch": cap["dram_bytes_per_launch"], "capture": cap["capture"]}
                                       if cap else None),
                           "peak_source": peak_src,
                           "bytes_alg_per_launch": bytes_alg / st["stage_launches"],
                           "bytes_alg_model": "compulsory: shared arrays once per genotype, candidate records and draws "
                                              "written + read once (DESIGN.md 3.2)",
                           "avg_launch_ms": st["stage_ms"] / st["stage_launches"],
                           "note": "dependent chains (one warp per logical worker): neither HBM nor a math pipe binds; "
                                   "see `issue` (issue-slot utilisation) and iters_per_s",
                           "iters_per_s": it / sec, "genotypes_per_s": g / sec}
        if cap and cap.get("warp_instructions_per_launch"):
            # warp-instructions of a launch are a property of the workload (same seeds, same build as the capture);
            # the duration is the one measured here with CUDA events
            slots = 148 * 4 * MAX_SM_MHZ[0] * 1e6 if MAX_SM_MHZ[0] else None
            ips = cap["warp_instructions_per_launch"] / (st["stage_ms"] / st["stage_launches"] / 1e3)
            out["roofline"]["issue"] = {
                "achieved": ips / 1e9, "peak": slots / 1e9 if slots else None, "unit": "G warp-instructions/s",
                "frac": ips / slots if slots else None,
                "warp_instructions_per_launch": cap["warp_instructions_per_launch"],
                "ncu_issue_active_pct": cap.get("issue_active_pct"), "ncu_l2_hit_pct": cap.get("l2_hit_pct"),
                "ncu_cycles_per_issue_per_warp": cap.get("cycles_per_issued_instruction_per_warp"),
                "peak_source": "148 SMs x 4 schedulers x 1 instruction per cycle at the max SM clock"}
    if st["prefilter_launches"]:
        gp = st["prefilter_genotypes"]
        sec = st["prefilter_ms"] / 1e3
        bytes_alg = gp * (p * R * 8 + 8)
        ops = gp * p * R
        out["roofline_prefilter"] = {"kernel": "k_prefilter_pairs", "bound": "hbm", "achieved": bytes_alg / sec / 1e9,
                                     "peak": peak, "unit": "GB/s", "frac": bytes_alg / sec / 1e9 / peak,
                                     "traffic": None, "f64_ops_per_s": ops / sec,
                                     "avg_launch_ms": st["prefilter_ms"] / st["prefilter_launches"],
                                     "note": "tiled (max,+) contraction: the matrix is read once per tile row, so the "
                                             "reference-pattern bytes exceed HBM peak by design; the binding unit is the "
                                             "FP64 pipe (DSETP + DADD per genotype-read), see fp64"}
        if fp64_rate:
            out["roofline_prefilter"]["fp64"] = {
                "achieved": ops / sec / 1e12, "peak": fp64_rate / 1e12, "unit": "T FP64-pipe lane-instructions/s",
                "frac": ops / sec / fp64_rate, "peak_source": "DADD microbenchmark run live (lctp_measure_fp64_rate)"}
    out["rates"] = kernel_rates(st)
    return out


def kernel_rates(st) -> dict:
    """the separately reported per-kernel rates, from the library's CUDA-event timings of the isolated
    pass (one kernel at a time): prefilter genotypes/s, stage genotypes/s, genotype-attempts/s, iterations (greedy)
    or steps (annealing) per second."""
    r = {}
    if st.get("prefilter_launches") and st.get("prefilter_ms", 0) > 0:
        r["prefilter_genotypes_per_s"] = st["prefilter_genotypes"] / (st["prefilter_ms"] / 1e3)
    if st.get("stage_launches") and st.get("stage_ms", 0) > 0:
        sec = st["stage_ms"] / 1e3
        r["stage_genotypes_per_s"] = st["stage_genotypes"] / sec
        r["stage_genotype_attempts_per_s"] = st["stage_attempts"] / sec
        r["stage_iterations_per_s"] = st["stage_iters"] / sec
    return r


def kir_prefilter(ctx, genotype, peak, fp64_rate):
    """The prefilter kernel alone at the KIR-scale shape of BASELINE configs[3] (H=1000, G=500,500, R=10,000),
    the size north_star's >= 60 % target is stated for; 5 launches after one warm-up, CUDA events."""
    from locityper_b200 import synth
    loc = synth.make_locus(**synth.config_shape("C4"), seed=4001, table_builder=genotype.build_depth_table)
    dl = ctx.upload(loc)
    dl.prefilter_scores(fetch=False)
    ctx.stats(reset=True)
    for _ in range(5):
        dl.prefilter_scores(fetch=False)
    ctx.sync()
    st = ctx.stats(reset=True)
    dl.free()
    sec = st["prefilter_ms"] / 1e3
    gp, R, p = st["prefilter_genotypes"], loc.n_reads, loc.ploidy
    ops = gp * p * R
    return {"kernel": "k_prefilter_bal_ws", "workload": "configs[3] shape: H=1000, G=500500, R=10000 (one locus, one GPU)",
            "bound": "fp64-pipe", "achieved": ops / sec / 1e12, "peak": fp64_rate / 1e12,
            "unit": "T FP64-pipe lane-instructions/s", "frac": ops / sec / fp64_rate,
            "peak_source": "DADD microbenchmark run live (lctp_measure_fp64_rate)",
            "avg_launch_ms": st["prefilter_ms"] / st["prefilter_launches"],
            "traffic": ncu_traffic("k_prefilter_bal_ws_kir"),
            "hbm_reference_pattern": {"achieved": gp * (p * R * 8 + 8) / sec / 1e9, "peak": peak, "unit": "GB/s",
                                      "frac": gp * (p * R * 8 + 8) / sec / 1e9 / peak}}


def cpu_baseline(args, loci, T):
    cores = os.cpu_count() or 1
    n = args.cpu_loci or len(loci)
    cpu_run(args, loci[:1], T, cores)           # warm-up (page-in, thread pool)
    CPU_SPLIT[0] = CPU_SPLIT[1] = 0.0
    g, dt, reps = 0, 0.0, 0
    while dt < 10.0 and reps < 8:               # bounded sample: ~10-30 s of CPU work
        gi, di = cpu_run(args, loci[:n], T, cores)
        g += gi; dt += di; reps += 1
    return {"value": g / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} loci x {reps} passes of the same workload, T={T} logical workers on {cores} pthreads",
            # the reference's prefilter is a plain loop on the calling thread (solve.rs:105-119) and the port keeps it so:
            # this is the share of the CPU arm's time that no amount of host cores would shorten
            "prefilter_share": CPU_SPLIT[0] / max(1e-12, CPU_SPLIT[0] + CPU_SPLIT[1])}


def run_shard(args, ctx, scheme, T, rank, world, dev):
    """configs[3]: one KIR-scale locus, genotype list partitioned across ranks."""
    import torch
    import torch.distributed as dist
    from locityper_b200 import genotype
    from locityper_b200 import dist as ldist
    args.config = "C4" if args.config == "C2" else args.config
    loc = pinned_locus(make_loci(args, 0, genotype.build_depth_table, 1)[0])   # same locus on every rank
    stream = torch.cuda.current_stream(dev)
    dl = ctx.upload(loc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        ldist.solve_sharded(dl, scheme, T, genotype.init_rng(args.seed), rank, world, dev)
    ctx.stats(reset=True)
    launches0 = ctx.launch_count()
    sampler = ClockSampler(dev.index or 0)
    barrier()
    sampler.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    res = None
    for _ in range(args.steps):
        res = ldist.solve_sharded(dl, scheme, T, genotype.init_rng(args.seed), rank, world, dev)
    b.record(stream)
    barrier()
    clocks = sampler.stop()
    if clocks.get("sm_max_mhz"):
        MAX_SM_MHZ[0] = float(clocks["sm_max_mhz"])
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    st = ctx.stats(reset=True)
    if rank == 0:
        peak, peak_src = load_peaks()
        ms = float(t.item())
        line = {"metric": METRIC, "value": loc.n_genotypes * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, T, world), "clocks": clocks,
                "gpu_launches": int(ctx.launch_count() - launches0),
                "call": list(loc.genotype_tuple(int(res["gt_ix"][0]))), "truth": list(loc.truth)}
        line.update(rooflines(st, [loc], args, peak, peak_src))
        print(json.dumps(line), flush=True)
    dl.free()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr
    global print
    real_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    _print = print

    def print(*a, **k):                      # noqa: A001  (json line -> the real stdout)
        k.pop("flush", None)
        _print(*a, file=real_out, **k)
        real_out.flush()
    globals()["print"] = print
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

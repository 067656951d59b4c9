"""Worker of tests/test_gpu_dist_abi.py: one rank of a sharded solve through the C ABI (lctp_dist_*).
Launched by `python -m torch.distributed.run --nproc-per-node N tests/dist_abi_worker.py OUT.json`; torch.distributed
is only the out-of-band channel for the 128-byte NCCL id and for comparing the ranks' results."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from locityper_b200 import genotype, synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(genotype.dist_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    uid = bytes(idt.cpu().numpy().tobytes())
    ctx = genotype.Context(device=local)
    d = genotype.Dist(ctx, uid, rank, world)
    out = {}
    cases = [("small", dict(n_haps=24, n_reads=300, locus_len=2500), 11, ["greedy:i=100,a=1", "anneal:i=10,a=5,n=2000,p=1000"], 64),
             ("c1", synth.config_shape("C1"), 1001, ["greedy:i=5k,a=1", "anneal:i=20,a=20"], 8),
             ("c5", synth.config_shape("C5"), 5000, ["greedy:i=5k,a=1"], 1000)]
    for name, shape, seed, specs, threads in cases:
        loc = synth.make_locus(**shape, seed=seed, table_builder=genotype.build_depth_table)
        dl = ctx.upload(loc)
        scheme = genotype.Scheme.parse(specs)
        # sharded
        rng_d = genotype.init_rng(seed)
        got = d.solve(dl, scheme, threads, rng_d)
        surv_d = d.prefilter(dl, scheme.stages[0].in_size, threads)
        # single GPU, same rank (every rank owns a full copy of the locus)
        rng_s = genotype.init_rng(seed)
        ref = dl.solve(scheme, threads, rng_s)
        surv_s = dl.prefilter(scheme.stages[0].in_size, threads)
        ok = (np.array_equal(got.gt_ix, ref.gt_ix) and np.array_equal(got.lik_mean, ref.lik_mean)
              and np.array_equal(got.lik_var, ref.lik_var, equal_nan=True) and np.array_equal(got.ln_prob, ref.ln_prob)
              and list(rng_d) == list(rng_s) and got.n_filtered == ref.n_filtered and got.n_stage_in == ref.n_stage_in
              and np.array_equal(surv_d, surv_s) and got.json_text == ref.json_text)
        # identical on every rank
        # (not Python's hash(): it is salted per process)
        h = torch.tensor([int.from_bytes(hashlib.sha256(got.json_text.encode()).digest()[:6], "little")], dtype=torch.int64, device="cuda")
        hs = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        same = all(int(x) == int(hs[0]) for x in hs)
        out[name] = dict(identical_to_single_gpu=bool(ok), identical_on_all_ranks=bool(same), call=[int(x) for x in got.gt_ix[:1]],
                         n_filtered=int(got.n_filtered))
        dl.free()
    out["timing"] = d.timing()
    d.close()
    ctx.close()
    allok = all(v["identical_to_single_gpu"] and v["identical_on_all_ranks"] for k, v in out.items() if k != "timing")
    flag = torch.tensor([1 if allok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out["world"] = world
        out["ok"] = bool(int(flag))
        with open(sys.argv[1], "w") as f:
            json.dump(out, f)
        print(json.dumps(out))
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()

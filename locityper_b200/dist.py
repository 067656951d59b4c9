"""Genotype sharding of one locus across the GPUs of a box.

Genotypes are independent units of both phases of `solve::solve` (src/solvers/solve.rs:105-119 and
:1116-1142), so one process per GPU owns a contiguous range of genotype ids in the prefilter and a
subset of the logical workers in every solver stage.  There is no data-path collective inside a kernel;
the two exchange steps are all-gathers of small per-rank result buffers over `torch.distributed`
(NCCL over NVLink on the GPU box, gloo in the CPU tests):

  prefilter : per-rank candidates (score f64, genotype id u64) that are guaranteed to contain every
              global survivor of `truncate_ixs` (src/solvers/solve.rs:52-84)  ->  every rank runs the exact
              `truncate_ixs` on the union and obtains the identical, identically ordered survivor list;
  stage     : `MainWorker::run` (:1049-1063) shuffles and partitions on every rank with the same locus
              stream (identical by construction); rank r solves workers w with w % world == r and the
              (lik_mean, lik_var) of its genotypes plus its workers' RNG states are all-gathered.

Everything that computes goes through a *backend* with the DeviceLocus interface (`prefilter_scores`,
`solve_stage`, `produce_result`); this module never touches the oracle or any CPU implementation.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import genotype


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of `n` genotype ids owned by `rank`."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def local_candidates(scores: np.ndarray, g_begin: int, filt_diff: float, min_size: int, threads: int):
    """Candidate survivors of one shard: every id whose score is >= min(local_best - filt_diff, K-th best
    local score), K = max(min_size, threads).

    Superset proof (truncate_ixs, src/solvers/solve.rs:60-81): the global threshold is
    global_best - filt_diff >= local_best - filt_diff, so every id above the global threshold is kept; the
    global "at least min_size, ties at the cut included" rule keeps ids with score >= the global K'-th best
    (K' = min_size), which is >= the local K-th best, so they are kept; raising to `threads` takes the
    `threads` best overall, which are among the local `threads` best of their shard.
    """
    n = len(scores)
    if n == 0:
        return np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.float64)
    k = max(int(min_size), int(threads), 1)
    thresh = float(np.max(scores)) - filt_diff
    if k >= n:
        thresh = -np.inf
    else:
        kth = float(np.partition(scores, n - k)[n - k])      # K-th best local score
        thresh = min(thresh, kth)
    keep = np.nonzero(scores >= thresh)[0]
    return (keep + g_begin).astype(np.uint64), np.ascontiguousarray(scores[keep])


class Comm:
    """Thin all-gather layer over torch.distributed (NCCL on device tensors, gloo on host tensors)."""

    def __init__(self, rank: int = 0, world: int = 1, device=None, group=None):
        self.rank, self.world, self.device, self.group = rank, world, device, group
        self.bytes_gathered = 0

    def allgather_var(self, arr: np.ndarray) -> list:
        """All-gather 1-D numpy arrays of rank-dependent length (same dtype); returns the list per rank."""
        arr = np.ascontiguousarray(arr)
        return [b.view(arr.dtype) for b in self.allgather_bytes(arr.view(np.uint8).reshape(-1))]

    def allgather_packed(self, arrays) -> list:
        """All-gather several 1-D arrays of rank-dependent lengths in ONE exchange (the volumes are KBs, so the
        cost is the collectives' latency): returns, per rank, the list of that rank's arrays."""
        arrays = [np.ascontiguousarray(a) for a in arrays]
        head = np.array([a.size for a in arrays], dtype=np.int64)
        blob = np.concatenate([head.view(np.uint8)] + [a.view(np.uint8).reshape(-1) for a in arrays])
        out = []
        for b in self.allgather_bytes(blob):
            sizes = b[:8 * len(arrays)].view(np.int64)
            pos, got = 8 * len(arrays), []
            for a, n in zip(arrays, sizes):
                nb = int(n) * a.itemsize
                got.append(b[pos:pos + nb].view(a.dtype).copy())
                pos += nb
            out.append(got)
        return out

    def allgather_bytes(self, raw: np.ndarray) -> list:
        """All-gather uint8 arrays of rank-dependent length: one fixed-size exchange of the lengths, one of the
        payloads padded to the longest."""
        if self.world == 1:
            return [raw]
        import torch
        import torch.distributed as dist
        dev = self.device if self.device is not None else torch.device("cpu")
        sizes = torch.zeros(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([raw.size], dtype=torch.int64, device=dev), group=self.group)
        sizes = [int(v) for v in sizes.cpu().tolist()]
        cap = max(max(sizes), 1)
        mine = np.zeros(cap, dtype=np.uint8)
        mine[:raw.size] = raw
        bufs = torch.empty(self.world * cap, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(bufs, torch.from_numpy(mine).to(dev), group=self.group)
        self.bytes_gathered += cap * self.world
        host = bufs.cpu().numpy()
        return [host[r * cap:r * cap + sizes[r]].copy() for r in range(self.world)]


def prefilter_sharded(backend, min_size: int, threads: int, comm: Comm):
    """run_filter (src/solvers/solve.rs:87-122) with the genotype list partitioned across ranks.
    Returns the sorted survivor ids (identical on every rank)."""
    loc = backend.loc
    G = loc.n_genotypes
    g0, g1 = shard_range(G, comm.rank, comm.world)
    scores = backend.prefilter_scores(g0, g1) if g1 > g0 else np.zeros(0)
    ids, sc = local_candidates(scores, g0, loc.filt_diff, min_size, threads)
    parts = comm.allgather_packed([ids, sc])
    all_ids = np.concatenate([p[0] for p in parts])
    all_sc = np.concatenate([p[1] for p in parts])
    # exact truncate_ixs on the union: ids ascending (= the order of predictions.ixs, solve.rs:406-411)
    order = np.argsort(all_ids, kind="stable")
    all_ids, all_sc = all_ids[order], all_sc[order]
    dense = np.full(G, -np.inf)
    dense[all_ids.astype(np.int64)] = all_sc
    return genotype.truncate_ixs(all_ids, dense, loc.filt_diff, min_size, threads)


def solve_stage_sharded(backend, stage, ixs: np.ndarray, off: np.ndarray, wrng: np.ndarray, comm: Comm):
    """One stage: rank r solves the logical workers w with w % world == r (their chunk of the shuffled list
    and their RNG stream, src/solvers/solve.rs:1052-1062).  `wrng` (u64[T,4]) is updated for ALL workers on
    every rank.  Returns (lik_mean, lik_var) indexed by position in `ixs`."""
    nw = len(off) - 1
    n = int(off[-1])
    off64 = np.asarray(off, dtype=np.int64)
    sizes = np.diff(off64)                                                  # genotypes per worker
    rank_of_pos = np.repeat(np.arange(nw, dtype=np.int64) % comm.world, sizes)   # owner of every list position
    pos_of = [np.flatnonzero(rank_of_pos == r) for r in range(comm.world)]  # ascending = worker order of rank r
    mine = np.arange(comm.rank, nw, comm.world)
    my_off = np.zeros(len(mine) + 1, dtype=np.uint64)
    my_off[1:] = np.cumsum(sizes[mine])
    my_rng = np.ascontiguousarray(wrng[mine]) if len(mine) else np.zeros((0, 4), dtype=np.uint64)
    if len(mine) and int(my_off[-1]) > 0:
        out = backend.solve_stage(stage, np.ascontiguousarray(ixs[pos_of[comm.rank]]), my_off, my_rng, want_liks=False)
        lm, lv = out["lik_mean"], out["lik_var"]
    else:
        lm, lv = np.zeros(0), np.zeros(0)
    parts = comm.allgather_packed([lm, lv, my_rng.reshape(-1)])             # one exchange per stage
    lik_mean, lik_var = np.empty(n), np.empty(n)
    for r in range(comm.world):
        lik_mean[pos_of[r]] = parts[r][0]
        lik_var[pos_of[r]] = parts[r][1]
        ws = np.arange(r, nw, comm.world)
        if len(ws):
            wrng[ws] = parts[r][2].reshape(-1, 4)
    return lik_mean, lik_var


def solve_sharded(backend, scheme: genotype.Scheme, threads: int, rng: np.ndarray, rank: int = 0, world: int = 1,
                  device=None, group=None, comm: Optional[Comm] = None) -> dict:
    """solve::solve (src/solvers/solve.rs:926-981) for one locus sharded over `world` ranks.  `rng` is the
    locus stream (u64[4], in/out).  Every rank returns the same result dictionary."""
    comm = comm or Comm(rank, world, device, group)
    loc = backend.loc
    G = loc.n_genotypes
    stages = scheme.stages
    threads = max(1, min(int(threads), G))                                  # genotype.rs:1247
    ixs = np.arange(G, dtype=np.uint64)
    if loc.dont_skip or stages[0].in_size < G:                               # solve.rs:941-945
        ixs = prefilter_sharded(backend, stages[0].in_size, threads, comm)
    n_filtered = len(ixs)
    lik_mean = np.full(G, np.nan)
    lik_var = np.full(G, np.nan)
    attempts = np.zeros(G, dtype=np.uint16)
    if threads > 1:                                                          # MainWorker::new, solve.rs:1007-1018
        wrng = genotype.worker_streams(rng, threads)
    n_stage_in = [0] * 8
    for s, st in enumerate(stages):
        has_next = s + 1 < len(stages)
        out_size = stages[s + 1].in_size if has_next else 0
        if not (loc.dont_skip or not has_next or out_size < len(ixs)):      # solve.rs:1041-1045
            continue
        n_stage_in[s] = len(ixs)
        ixs = np.ascontiguousarray(ixs, dtype=np.uint64)
        if threads == 1:                                                     # solve_single_thread: one stream
            off = np.array([0, len(ixs)], dtype=np.uint64)
            state = rng.reshape(1, 4).copy()
            if comm.rank == 0:
                out = backend.solve_stage(st, ixs, off, state, want_liks=False)
                lm, lv = out["lik_mean"], out["lik_var"]
            else:
                lm, lv, state = np.zeros(0), np.zeros(0), np.zeros((0, 4), dtype=np.uint64)
            parts = comm.allgather_packed([lm, lv, state.reshape(-1)])
            lm = np.concatenate([p[0] for p in parts])
            lv = np.concatenate([p[1] for p in parts])
            rng[:] = np.concatenate([p[2] for p in parts])[:4]
        else:
            off = genotype.plan_stage(rng, ixs, threads)
            lm, lv = solve_stage_sharded(backend, st, ixs, off, wrng, comm)
        idx = ixs.astype(np.int64)
        lik_mean[idx], lik_var[idx], attempts[idx] = lm, lv, st.attempts
        if has_next:
            ixs = genotype.discard_improbable(ixs, lik_mean, lik_var, attempts, loc.prob_thresh, out_size, threads)
    res = backend.produce_result(ixs, lik_mean, lik_var, attempts)
    res.update(n_filtered=n_filtered, n_stage_in=n_stage_in, allgather_bytes=comm.bytes_gathered)
    return res

"""Multi-rank genotype sharding (locityper_b200/dist.py) on CPU: world_size 2 and 3 over gloo.

The compute backend of these tests is the CPU oracle (tests may use it as the checker AND, here, as a
stand-in device so that the sharding / all-gather / merge logic can run without a GPU); the product
backend is `genotype.DeviceLocus`, exercised by tests/test_gpu_parity.py::test_sharded_solve_*.
"""
import os
import socket

import numpy as np
import pytest

from locityper_b200 import dist as ldist
from locityper_b200 import genotype, synth


class OracleBackend:
    """DeviceLocus look-alike on top of the oracle (test infrastructure)."""

    def __init__(self, O, loc):
        self.O, self.loc = O, loc
        self.ol = O.OracleLocus(loc)
        self.M = O.best_aln_matrix(self.ol)

    def prefilter_scores(self, g0, g1):
        ixs = np.arange(g0, g1, dtype=np.uint64)
        return self.O.prefilter_scores(self.ol, ixs=ixs, M=self.M)[g0:g1].copy()

    def solve_stage(self, stage, ixs, off, rng, want_liks=False):
        ost = self.O.Stage(kind=stage.kind, attempts=stage.attempts, in_size=stage.in_size,
                           best_start=stage.best_start, sample_size=stage.sample_size, plato_size=stage.plato_size,
                           anneal_steps=stage.anneal_steps, init_prob=stage.init_prob)
        return self.O.solve_stage(self.ol, ost, ixs, off, rng, os_threads=1)

    def produce_result(self, ixs, lik_mean, lik_var, attempts):
        ixs = np.array(ixs, dtype=np.uint64)
        order = np.argsort(-lik_mean[ixs.astype(np.int64)], kind="stable")
        ixs = ixs[order]
        return dict(gt_ix=ixs, lik_mean=lik_mean[ixs.astype(np.int64)], lik_var=lik_var[ixs.astype(np.int64)])


def _scheme():
    return genotype.Scheme([genotype.Stage("greedy", attempts=1, in_size=60),
                            genotype.Stage("anneal", attempts=3, in_size=8, anneal_steps=1200, plato_size=600)])


def _locus(O):
    return synth.make_locus(20, 150, 2500, seed=77, table_builder=O.build_depth_table)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, threads, q):
    import torch.distributed as dist
    from oracle import lcto_py as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        loc = _locus(O)
        be = OracleBackend(O, loc)
        rng = genotype.init_rng(5)
        res = ldist.solve_sharded(be, _scheme(), threads, rng, rank, world)
        surv = ldist.prefilter_sharded(be, 60, threads, ldist.Comm(rank, world))
        q.put((rank, res["gt_ix"][:10].tolist(), res["lik_mean"][:10].tolist(), res["n_filtered"],
               res["n_stage_in"], rng.tolist(), surv.tolist(), res["allgather_bytes"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,threads", [(2, 8), (3, 5), (2, 1)])
def test_sharded_solve_matches_single_process_oracle(oracle, world, threads):
    import torch.multiprocessing as mp
    O = oracle
    loc = _locus(O)
    ol = O.OracleLocus(loc)
    scheme_o = [O.Stage("greedy", attempts=1, in_size=60),
                O.Stage("anneal", attempts=3, in_size=8, anneal_steps=1200, plato_size=600)]
    rng_o = O.Rng.from_seed(5)
    ref = O.solve(ol, scheme_o, threads, rng_o, os_threads=1, want_scores=True)

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, threads, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k = min(10, len(ref["gt_ix"]))
    for rank, gt_ix, lik_mean, n_filtered, n_stage_in, rng_state, surv, nbytes in outs:
        assert n_filtered == ref["n_filtered"] and n_stage_in == ref["n_stage_in"]
        assert surv == ref["filtered_ixs"].tolist(), "sharded prefilter survivors differ (set or order)"
        assert gt_ix[:k] == ref["gt_ix"][:k].tolist(), "ranking differs"
        assert lik_mean[:k] == ref["lik_mean"][:k].tolist()
        assert rng_state == rng_o.state(), "locus RNG stream diverged"
        assert nbytes > 0


def test_shard_range_and_local_candidates():
    for n, w in [(10, 3), (5050, 8), (7, 8), (0, 2)]:
        cuts = [ldist.shard_range(n, r, w) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
    rng = np.random.default_rng(0)
    for trial in range(20):
        G = 500
        scores = -rng.gamma(2.0, 40.0, G)
        scores[rng.integers(0, G, 30)] = scores[rng.integers(0, G, 30)]      # exact ties
        filt_diff, min_size, threads = float(rng.choice([5.0, 50.0, 230.0])), int(rng.integers(1, 80)), int(rng.integers(1, 40))
        full = genotype.truncate_ixs(np.arange(G), scores, filt_diff, min_size, threads)
        for world in (1, 2, 5):
            ids, sc = [], []
            for r in range(world):
                a, b = ldist.shard_range(G, r, world)
                i, s = ldist.local_candidates(scores[a:b], a, filt_diff, min_size, threads)
                ids.append(i); sc.append(s)
            ids, sc = np.concatenate(ids), np.concatenate(sc)
            dense = np.full(G, -np.inf)
            dense[ids.astype(np.int64)] = sc
            merged = genotype.truncate_ixs(np.sort(ids), dense, filt_diff, min_size, threads)
            assert np.array_equal(merged, full)


def _packed_main(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = ldist.Comm(rank, world)
        # rank-dependent lengths (rank 1 sends nothing), three dtypes in one exchange
        n = 0 if rank == 1 else 5 + 3 * rank
        parts = comm.allgather_packed([np.arange(n, dtype=np.uint64) + 100 * rank,
                                       np.linspace(0.0, 1.0, 2 * n) - rank,
                                       np.full(rank, 7, dtype=np.uint8)])
        single = comm.allgather_var(np.arange(rank + 1, dtype=np.float64))
        q.put((rank, [[a.tolist() for a in p_] for p_ in parts], [a.tolist() for a in single], comm.bytes_gathered))
    finally:
        dist.destroy_process_group()


def test_packed_allgather_roundtrip_three_ranks():
    """Comm.allgather_packed / allgather_var: rank-dependent lengths (including empty), mixed dtypes, one exchange."""
    import torch.multiprocessing as mp
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_packed_main, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, parts, single, nbytes in outs:
        assert len(parts) == world and nbytes > 0
        for r in range(world):
            n = 0 if r == 1 else 5 + 3 * r
            assert parts[r][0] == (np.arange(n, dtype=np.uint64) + 100 * r).tolist()
            assert parts[r][1] == (np.linspace(0.0, 1.0, 2 * n) - r).tolist()
            assert parts[r][2] == [7] * r
            assert single[r] == list(np.arange(r + 1, dtype=np.float64))


def test_packed_allgather_single_rank_is_identity():
    comm = ldist.Comm(0, 1)
    a, b = np.arange(4, dtype=np.uint64), np.array([0.5, -1.25])
    (got,) = comm.allgather_packed([a, b])
    assert np.array_equal(got[0], a) and np.array_equal(got[1], b) and got[0].dtype == a.dtype
    (one,) = comm.allgather_var(b)
    assert np.array_equal(one, b)

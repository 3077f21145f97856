"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo job.  Each rank owns a
contiguous block of the global seed list; the only collective is one all-reduce of the
expectation sums (solve.reduce_expect_sums; on GPUs the same sums go through
engine.Comm = ncclAllReduce behind the C ABI, here through gloo), which must reproduce the
reference's trajectory average / std (multitrajresult.py:261-279,1116-1124)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from _golden import load
from qutip_b200 import solve


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load("c3_tfim6_mc")
    ntraj = int(g["ntraj"])
    lo, hi = solve.shard_range(ntraj, rank, world)
    # thresholds of this rank's block equal the global table's rows
    d = solve.make_thresholds(int(g["seed"]), hi - lo, 64, first=lo)
    assert np.array_equal(d, g["draws"][lo:hi])
    runs = g["runs_expect"][:, lo:hi, :].astype(complex)      # stands in for the device run
    import torch

    def gloo_allreduce(flat):               # sums the float64 array in place over the ranks
        dist.all_reduce(torch.from_numpy(flat), op=dist.ReduceOp.SUM)

    s1, s2 = solve.reduce_expect_sums(runs, allreduce=gloo_allreduce)
    avg, std = solve.finish_expect_sums(s1, s2, ntraj)
    q.put((rank, lo, hi, avg.real, std))
    dist.destroy_process_group()


def test_sharded_reduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    g = load("c3_tfim6_mc")
    ref_avg = g["avg_expect"].real
    runs = g["runs_expect"]
    ref_std = np.sqrt(np.abs((runs ** 2).mean(axis=1) - runs.mean(axis=1) ** 2))
    covered = sorted((lo, hi) for _, lo, hi, _, _ in out)
    assert covered[0][0] == 0 and covered[-1][1] == int(g["ntraj"]) and covered[0][1] == covered[1][0]
    for _, _, _, avg, std in out:
        np.testing.assert_allclose(avg, ref_avg, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(std, ref_std, rtol=1e-9, atol=1e-12)


def test_shard_range_covers_everything():
    for n in (1, 7, 10000):
        for w in (1, 2, 4, 8):
            spans = [solve.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_thresholds_match_reference_generators():
    # same stream as SeedSequence(seed).spawn(n) -> default_rng(child).random()
    d = solve.make_thresholds(7, 5, 9)
    kids = np.random.SeedSequence(7).spawn(5)
    for j, k in enumerate(kids):
        g = np.random.default_rng(k)
        assert [g.random() for _ in range(9)] == list(d[j])

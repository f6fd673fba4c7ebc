"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path (bayesrrcmd_b200/sharding.py) -- shard bounds,
per-shard PLINK payloads, the allreduce of the class counts, the sharded phenotype standardisation, the scalar sums of a sweep
and the gather of the exchange handles.  (The in-kernel exchange itself is covered on the GPU by
tests/test_gpu_dense.py::test_sharded_individuals_exchange_on_one_device.)"""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from bayesrrcmd_b200.sharding import ShardGroup, class_counts_host, shard_bed, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(N=1003, M=37, seed=5):
    rng = np.random.default_rng(seed)
    code = rng.integers(0, 4, size=(M, N)).astype(np.uint8)          # PLINK 2-bit codes, 01 = missing included
    bed = np.zeros((M, (N + 3) // 4), dtype=np.uint8)
    for k in range(4):
        sub = code[:, k::4]
        bed[:, :sub.shape[1]] |= (sub << (2 * k)).astype(np.uint8)
    y = rng.normal(size=N).astype(np.float32)
    return code, bed, y


def _worker(rank, world, port, N, M, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        code, bed, y = _problem(N, M)
        grp = ShardGroup(rank, world)
        lo, hi = shard_bounds(N, world)[rank]
        local = class_counts_host(shard_bed(bed, lo, hi), hi - lo)
        counts = grp.allreduce_counts(local)
        ys = grp.standardise(y[lo:hi], N)
        handles = grp.gather(bytes([rank]) * 64)
        total = grp.allsum(float(ys.sum()))
        q.put((rank, lo, hi, counts, ys, handles, total))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_and_align():
    for N, world in [(500000, 8), (500000, 1), (1003, 2), (64, 2), (3642, 3), (130, 4)]:
        b = shard_bounds(N, world)
        assert b[0][0] == 0 and b[-1][1] == N and len(b) == world
        for (lo, hi), (lo2, _) in zip(b, b[1:]):
            assert hi == lo2
        assert all(lo % 64 == 0 for lo, hi in b if hi > lo)
        assert sum(hi - lo for lo, hi in b) == N


def test_class_counts_host_matches_direct_count():
    code, bed, _ = _problem(777, 11)
    c = class_counts_host(bed, 777)
    cls_of_plink = np.array([2, 3, 1, 0])
    for c_ in range(4):
        assert (c[:, c_] == (cls_of_plink[code] == c_).sum(axis=1)).all()
    assert (c.sum(axis=1) == 777).all()


@pytest.mark.timeout(120)
def test_two_ranks_over_gloo():
    N, M, world = 1003, 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, M, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    code, bed, y = _problem(N, M)
    ref_counts = class_counts_host(bed, N)
    y64 = y.astype(np.float64)
    y_ref = y64 - y64.sum() / N
    y_ref = y_ref / np.sqrt((y_ref ** 2).sum() / (N - 1))
    ys = np.concatenate([r[4] for r in res])
    assert ys.shape == (N,)
    assert np.max(np.abs(ys - y_ref)) <= 1e-12
    for rank, lo, hi, counts, _, handles, total in res:
        assert (counts == ref_counts).all()                            # integers: exact for any sharding
        assert handles == [bytes([0]) * 64, bytes([1]) * 64]           # rank order
        assert abs(total - ys.sum()) <= 1e-9
    assert res[0][6] == res[1][6]                                      # every rank holds the bit-identical sum

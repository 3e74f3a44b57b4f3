"""CPU, world_size 2 over gloo: the multi-GPU path of the sweep (shard images, all-reduce totals, gather replayed
columns) reaches the single-process answer on every rank.  Counts come from the oracle; on the GPU box the same host
code runs over NCCL with counts from the CUDA kernel (tests/test_rcps_gpu.py::test_two_rank_*)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from im2im_uq_b200.calibration import sweep
from oracle import rcps_oracle as orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, split, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden(case)
        n = g["outputs"].shape[0]
        bounds_ = [0, split, n] if world == 2 else [0, n]
        lo, hi = bounds_[rank], bounds_[rank + 1]
        out, lab = g["outputs"][lo:hi], g["labels"][lo:hi]
        L = len(g["lam_prime"])
        px = int(np.prod(g["outputs"].shape[2:]))
        counts = torch.from_numpy(orc.c_miss_table(out, lab, g["lam_prime"])) if hi > lo else torch.zeros((0, L), dtype=torch.int32)
        totals = counts.sum(0, dtype=torch.int64)
        stats = {}
        lhat, stop, visited = sweep.sweep_from_counts(counts, totals, px, g["config"], lambda c: c.float() / float(px),
                                                      group=dist.group.WORLD, stats=stats)
        q.put((rank, float(lhat), stop, visited.tolist(), totals.tolist(), stats))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case,split", [("fastmri_small", 24), ("temca_small", 11), ("top_risk_zero", 10),
                                        ("never_stops", 8), ("batch65", 65), ("nasty_ragged", 0)])
def test_two_ranks_agree_with_reference(case, split):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, split, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = load_golden(case)
    for rank, lhat, stop, visited, totals, stats in results:
        assert stop == int(g["stop_idx"]), (rank, stop)
        assert np.float32(lhat) == g["lhat"]
        assert totals == g["counts_prime"].sum(0, dtype=np.int64).tolist()  # all-reduced totals are global
    assert results[0][3] == results[1][3]


def _shard_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from im2im_uq_b200.calibration.calibrate_model import _tensor_pair, rank_shard
        ds = torch.utils.data.TensorDataset(torch.arange(n).float().view(n, 1), torch.arange(n).float().view(n, 1) + 0.5)
        try:
            sub = rank_shard(ds, dist.group.WORLD)
        except ValueError as e:
            q.put((rank, "ValueError", str(e), None))
            return
        xs, ys = _tensor_pair(sub)                      # contiguous Subset of a TensorDataset -> tensor slices
        q.put((rank, [int(sub[i][0]) for i in range(len(sub))], xs.view(-1).tolist(), ys.view(-1).tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7, 2, 1])
def test_rank_shard_partitions_the_calibration_set_in_row_order(n):
    """calibrate_model(..., group=...): contiguous blocks in rank order, every image exactly once, the TensorDataset fast
    path sees exactly that block; fewer images than ranks is refused on every rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if n < 2:
        assert all(r[1] == "ValueError" and "cannot be sharded" in r[2] for r in results)
        return
    assert results[0][1] + results[1][1] == list(range(n))
    for rank, idx, xs, ys in results:
        assert xs == [float(i) for i in idx] and ys == [i + 0.5 for i in idx]


def _gather_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from im2im_uq_b200.calibration.calibrate_model import gather_loss_table
        rows = [3, 5][rank]                                   # uneven shards
        table = torch.arange(rows * 4, dtype=torch.float32).reshape(rows, 4) + 100 * rank
        full = gather_loss_table(table, dist.group.WORLD, torch.device("cpu"))
        q.put((rank, full.numpy()))
    finally:
        dist.destroy_process_group()


def test_gather_loss_table_rank_order_uneven_shards():
    """calibrate_model(..., group=, gather_table=True): rows of all ranks in rank order, uneven shard sizes (host logic, gloo)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.concatenate([np.arange(12, dtype=np.float32).reshape(3, 4), np.arange(20, dtype=np.float32).reshape(5, 4) + 100])
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)

"""Host-side logic of the multi-GPU path, exercised with world_size-2 gloo on the CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aes_lac_2018_b200.distributed import all_reduce_loss, balanced_shards, shard_bounds, shard_problem


def test_shard_bounds_cover_batch_exactly():
    for B in (1, 7, 32, 1024, 1025):
        for W in (1, 2, 4, 8):
            spans = [shard_bounds(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def test_shard_problem_slices_flat_labels():
    ll = torch.tensor([3, 0, 2, 4], dtype=torch.int32)
    al = torch.tensor([9, 8, 7, 6], dtype=torch.int32)
    labels = torch.arange(1, 10, dtype=torch.int32)
    lab, a, l = shard_problem(labels, al, ll, 1, 3)
    assert lab.tolist() == [4, 5] and a.tolist() == [8, 7] and l.tolist() == [0, 2]
    lab, a, l = shard_problem(labels, al, ll, 3, 4)
    assert lab.tolist() == [6, 7, 8, 9]
    lab, a, l = shard_problem(labels, al, ll, 0, 4)
    assert lab.tolist() == labels.tolist()


def test_balanced_shards_partition_and_balance():
    rng = np.random.default_rng(0)
    al = rng.integers(100, 800, 64)
    ll = rng.integers(10, 200, 64)
    shards = balanced_shards(al, ll, 4)
    assert sorted(i for s in shards for i in s) == list(range(64))
    loads = [sum(int(al[i]) * (2 * int(ll[i]) + 1) for i in s) for s in shards]
    assert max(loads) / min(loads) < 1.1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank derives its shard of one seeded global problem and reduces its partial loss
        from oracle import ctc_f64
        from tests.helpers import synth_problem
        acts, labels, al, ll = synth_problem(5, 30, 6, 7, 1, 6, tmin=20)
        lo, hi = shard_bounds(6, world, rank)
        lab, a, l = shard_problem(torch.tensor(labels), torch.tensor(al), torch.tensor(ll), lo, hi)
        costs, _ = ctc_f64.ctc_batch(acts[:, lo:hi], lab.numpy(), a.numpy(), l.numpy())   # the checker stands in for the GPU
        total = all_reduce_loss(torch.tensor([costs.sum()], dtype=torch.float32))
        q.put((rank, float(total), float(costs.sum())))
    finally:
        dist.destroy_process_group()


def test_scalar_loss_all_reduce_gloo_world2():
    from oracle import ctc_f64
    from tests.helpers import synth_problem
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    acts, labels, al, ll = synth_problem(5, 30, 6, 7, 1, 6, tmin=20)
    want = ctc_f64.ctc_batch(acts, labels, al, ll)[0].sum()
    assert abs(res[0][1] - want) < 1e-4 * want and res[0][1] == res[1][1]
    assert abs(res[0][2] + res[1][2] - want) < 1e-4 * want


def test_pending_loss_wait_is_idempotent():
    """`sharded_loss_step(overlap=True)` hands the global loss back as a PendingLoss; without a process group (or after
    the first wait) there is nothing to wait for."""
    import torch
    from aes_lac_2018_b200.distributed import PendingLoss

    class _Work:
        def __init__(self):
            self.n = 0

        def wait(self):
            self.n += 1

    w = _Work()
    p = PendingLoss(torch.tensor([3.0]), w)
    assert float(p.wait()) == 3.0 and float(p.wait()) == 3.0 and w.n == 1
    assert float(PendingLoss(torch.tensor([1.0])).wait()) == 1.0

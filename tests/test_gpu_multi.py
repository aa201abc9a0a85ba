"""Two-GPU check of the sharded path (skipped on a single-GPU box): each rank runs the engine on its shard,
the only collective is the scalar NCCL loss sum, gradients stay local and equal the single-GPU ones."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from aes_lac_2018_b200.distributed import ShardedCTCLoss, shard_bounds, shard_problem
    from tests.helpers import synth_problem
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        acts, labels, al, ll = synth_problem(77, 90, 10, 29, 3, 30, tmin=50)
        lo, hi = shard_bounds(10, world, rank)
        lab, a, l = shard_problem(torch.tensor(labels), torch.tensor(al), torch.tensor(ll), lo, hi)
        x = torch.tensor(acts[:, lo:hi]).cuda().requires_grad_()
        total, local = ShardedCTCLoss()(x, lab, a, l)
        local.sum().backward()
        # the non-blocking path: engine (NO_SYNC) -> device-side cost sum -> stream-ordered NCCL all-reduce
        from aes_lac_2018_b200.distributed import sharded_loss_step
        t2, l2, g2, st2 = sharded_loss_step(x.detach(), lab, a, l)
        assert t2.is_cuda and l2.is_cuda and not st2.any().item()
        assert abs(float(t2) - float(total)) <= 1e-5 * abs(float(total)) and abs(float(l2) - float(local)) <= 1e-5 * abs(float(local))
        assert torch.equal(g2, x.grad)
        # ... and with the collective running beside the next step's kernels: same numbers once waited for
        pend = [sharded_loss_step(x.detach(), lab, a, l, overlap=True)[0] for _ in range(3)]
        for p in pend:
            t3 = p.wait()
            assert t3.is_cuda and float(t3) == float(t2)
        q.put((rank, float(total), float(local), x.grad.cpu().numpy(), lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_loss_and_local_gradients():
    import torch.multiprocessing as mp
    from oracle import ctc_f64
    from tests.helpers import synth_problem
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    acts, labels, al, ll = synth_problem(77, 90, 10, 29, 3, 30, tmin=50)
    oc, og = ctc_f64.ctc_batch(acts, labels, al, ll)
    assert abs(res[0][1] - oc.sum()) <= 1e-4 * oc.sum() and res[0][1] == res[1][1]
    assert abs(res[0][2] + res[1][2] - oc.sum()) <= 1e-4 * oc.sum()
    for _, _, _, g, lo, hi in res:
        assert np.abs(g - og[:, lo:hi]).max() <= 1e-5

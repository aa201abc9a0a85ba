"""Batch sharding of the CTC path over the GPUs of one box (one process per GPU).

Utterances are independent, so the path shards along the batch axis with NO data-path collective;
gradients stay on the GPU that owns the utterance.  The only exchange is the scalar loss sum
(`all_reduce(SUM)` of one fp32 over NCCL/NVLink) -- SURVEY.md section 8e.  The reference itself never
reduces the loss across ranks (each DDP rank logs its local loss, /root/reference/train.py:246-259);
its per-rank bucketing is /root/reference/codes/sampler.py:100-135.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of the batch owned by `rank` (sizes differ by at most one)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_shards(act_lens: Sequence[int], label_lens: Sequence[int], world_size: int) -> list[list[int]]:
    """Assign utterances to ranks so that the recursion work sum(T_b * (2 L_b + 1)) is balanced
    (longest-processing-time greedy).  Returns one index list per rank, each in ascending order."""
    work = [(int(t) * (2 * int(l) + 1), i) for i, (t, l) in enumerate(zip(act_lens, label_lens))]
    work.sort(reverse=True)
    loads = [0] * world_size
    out: list[list[int]] = [[] for _ in range(world_size)]
    for w, i in work:
        r = min(range(world_size), key=lambda q: (loads[q], q))
        loads[r] += w
        out[r].append(i)
    return [sorted(ix) for ix in out]


def shard_problem(labels: torch.Tensor, act_lens: torch.Tensor, label_lens: torch.Tensor, lo: int, hi: int):
    """Slice the flat label tensor and the length tensors for utterances [lo, hi)."""
    label_lens = label_lens.reshape(-1)
    offs = torch.cumsum(label_lens.to(torch.int64), 0) - label_lens.to(torch.int64)
    start = int(offs[lo]) if lo < label_lens.numel() else int(label_lens.sum())
    stop = int(offs[hi - 1] + label_lens[hi - 1]) if hi > lo else start
    return labels.reshape(-1)[start:stop], act_lens.reshape(-1)[lo:hi], label_lens[lo:hi]


def all_reduce_loss(local_loss: torch.Tensor, group=None) -> torch.Tensor:
    """Global sum of the per-rank loss scalars (the path's single collective).  With the NCCL backend the
    tensor must live on this rank's GPU; with gloo (CPU tests) it stays on the host."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_loss
    backend = dist.get_backend(group)
    t = local_loss.detach().clone().reshape(1).to(torch.float32)
    if backend == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class PendingLoss:
    """The global loss of a step whose scalar all-reduce runs beside the next step's kernels (`overlap=True`).
    `wait()` orders the current stream behind the collective and returns the [1] tensor."""

    def __init__(self, tensor: torch.Tensor, work=None):
        self.tensor, self.work = tensor, work

    def wait(self) -> torch.Tensor:
        if self.work is not None:
            self.work.wait()
            self.work = None
        return self.tensor


def sharded_loss_step(acts, labels, act_lens, label_lens, blank: int = 0, grad_scale: float = 1.0, want_grad: bool = True,
                      group=None, mode: str = "auto", overlap: bool = False):
    """One step of the sharded path with NOTHING leaving the device: the engine runs on this rank's shard without a
    host synchronisation (CTC_B200_FLAG_NO_SYNC), the per-utterance costs are summed by a kernel on the same
    stream, and that one fp32 is all-reduced over NCCL/NVLink -- stream-ordered behind the kernels (SURVEY.md
    section 5: "the scalar ncclAllReduce enqueued on the compute stream right after the beta/grad kernel").
    Returns (global_loss [1] CUDA, local_loss [1] CUDA, grads [T,B,V] CUDA or None, status [B] CUDA int32).
    Round 1 copied the cost vector to the host, summed it there and copied the sum back before the collective
    (one blocking sync per step: the named limiter of its 1 -> 8 GPU curve).
    `overlap=True`: the collective is issued asynchronously (NCCL's own stream, ordered behind the cost sum) and the
    compute stream does NOT wait for it -- nothing in the next step needs the global loss, and a per-step rendezvous
    on the compute stream makes every step as slow as the slowest rank (measured at 8 GPUs: 0.45 ms of a 2.6 ms step).
    The first return value is then a `PendingLoss`; call `.wait()` before reading it."""
    from .ctc_loss import ctc_loss_raw, reduce_costs
    costs, grads, status = ctc_loss_raw(acts, labels, act_lens, label_lens, blank=blank, want_grad=want_grad,
                                        grad_scale=grad_scale, mode=mode, no_sync=True)
    local, _ = reduce_costs(costs, 1.0, False)
    total, work = local, None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        total = local.clone()
        work = dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group, async_op=overlap)
    if overlap:
        return PendingLoss(total, work), local, grads, status
    return total, local, grads, status


class ShardedCTCLoss(torch.nn.Module):
    """CTCLoss over this rank's shard of the batch; forward returns (global_loss[1], local_loss[1]).
    Backward through `local_loss` gives this rank's gradients; nothing else is communicated."""

    def __init__(self, blank: int = 0, group=None):
        super().__init__()
        from .ctc_loss import CTCLoss
        self.ctc = CTCLoss(blank=blank)
        self.group = group

    def forward(self, acts, labels, act_lens, label_lens):
        local = self.ctc(acts, labels, act_lens, label_lens)
        return all_reduce_loss(local, self.group).cpu(), local

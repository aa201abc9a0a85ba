"""Fused classifier head -- drop-in for the reference's `SequenceWiseClassifier` / `DeepSpeech2.fc`.

Reference: /root/reference/codes/model.py:177-180 + 199-207 (single task) and model.py:205-222
(`SequenceWiseClassifier`, the multi-task heads): `SequenceWise(Sequential(BatchNorm1d(H), Linear(H, V, bias=False)))`
applied to the T x B x H output of the recurrent stack, `.transpose(0, 1)`, softmax in eval mode.
Same constructor, same parameter / buffer names in `state_dict()` (`fc.0.module.0.weight`, ... `fc.0.module.1.weight`),
same outputs; the forward and backward run as the fused sm_100a kernels of csrc/ctc_head.cu (two passes over x each
way instead of PyTorch's eight kernels).  The torch modules inside are parameter containers only.  CUDA only.
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib

__all__ = ["SequenceWiseClassifier", "SequenceWise", "head_forward_raw"]


def _workspace(lib, N, H, V, device):
    need = ctypes.c_size_t(0)
    st = lib.ctc_b200_head_workspace_size(N, H, V, ctypes.byref(need))
    if st != _lib.CTC_STATUS_SUCCESS:
        raise RuntimeError("ctc_b200_head_workspace_size: " + _lib.status_string(lib, st))
    return torch.empty(need.value, dtype=torch.uint8, device=device)


def _ptr(t):
    return t.data_ptr() if t is not None else None


def head_forward_raw(x, weight, bn_weight, bn_bias, running_mean, running_var, training, eps=1e-5, momentum=0.1,
                     softmax=False):
    """x: CUDA float32 [T, B, H] (or [N, H]).  Returns (out [T, B, V], save_mean [H], save_invstd [H], x2, W) where x2 / W
    are the contiguous [N, H] / [V, H] tensors the kernels read (kept for the backward call); updates the running
    statistics in place in training mode."""
    lib = _lib.load()
    if not x.is_cuda:
        raise RuntimeError("aes_lac_2018_b200 classifier head is CUDA-only (B200-native); there is no CPU fallback")
    if x.dtype != torch.float32 or weight.dtype != torch.float32:
        raise TypeError("x and weight must be float32")
    lead = x.shape[:-1]
    H = x.shape[-1]
    x2 = x.detach().reshape(-1, H).contiguous()
    N = x2.shape[0]
    W = weight.detach().contiguous()
    V = W.shape[0]
    if W.shape[1] != H:
        raise ValueError("weight must be [classes, features]")
    dev = x2.device
    with torch.cuda.device(dev):
        out = torch.empty((N, V), dtype=torch.float32, device=dev)
        save_mean = torch.empty(H, dtype=torch.float32, device=dev)
        save_invstd = torch.empty(H, dtype=torch.float32, device=dev)
        ws = _workspace(lib, N, H, V, dev)
        c = _lib.CtcB200HeadForward()
        c.x, c.rows, c.features, c.classes = x2.data_ptr(), N, H, V
        c.weight = W.data_ptr()
        c.bn_weight, c.bn_bias = _ptr(bn_weight), _ptr(bn_bias)
        c.running_mean, c.running_var = _ptr(running_mean), _ptr(running_var)
        c.eps, c.momentum = float(eps), float(momentum)
        c.training, c.softmax = int(bool(training)), int(bool(softmax))
        c.out, c.save_mean, c.save_invstd = out.data_ptr(), save_mean.data_ptr(), save_invstd.data_ptr()
        c.workspace, c.workspace_bytes = ws.data_ptr(), ws.numel()
        c.stream = torch.cuda.current_stream(dev).cuda_stream
        st = lib.ctc_b200_head_forward(ctypes.byref(c))
        if st != _lib.CTC_STATUS_SUCCESS:
            raise RuntimeError("ctc_b200_head_forward: " + _lib.status_string(lib, st))
    return out.view(*lead, V), save_mean, save_invstd, x2, W


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bn_weight, bn_bias, running_mean, running_var, training, eps, momentum, softmax):
        out, save_mean, save_invstd, x2, W = head_forward_raw(x, weight, bn_weight, bn_bias, running_mean, running_var,
                                                              training, eps, momentum, softmax)
        ctx.save_for_backward(x2, W, bn_weight, bn_bias, save_mean, save_invstd)
        ctx.training, ctx.softmax, ctx.x_shape = bool(training), bool(softmax), x.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.softmax:
            raise RuntimeError("the eval-mode (softmax) head output is not differentiable here; "
                               "train on the logits as the reference does")
        lib = _lib.load()
        x2, W, bn_weight, bn_bias, save_mean, save_invstd = ctx.saved_tensors
        N, H = x2.shape
        V = W.shape[0]
        dl = dout.reshape(N, V).contiguous()
        dev = x2.device
        with torch.cuda.device(dev):
            dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
            dW = torch.empty_like(W) if ctx.needs_input_grad[1] else None
            dg = torch.empty(H, dtype=torch.float32, device=dev) if (bn_weight is not None and ctx.needs_input_grad[2]) else None
            db = torch.empty(H, dtype=torch.float32, device=dev) if (bn_bias is not None and ctx.needs_input_grad[3]) else None
            ws = _workspace(lib, N, H, V, dev)
            c = _lib.CtcB200HeadBackward()
            c.x, c.dlogits, c.rows, c.features, c.classes = x2.data_ptr(), dl.data_ptr(), N, H, V
            c.weight, c.bn_weight, c.bn_bias = W.data_ptr(), _ptr(bn_weight), _ptr(bn_bias)
            c.save_mean, c.save_invstd = save_mean.data_ptr(), save_invstd.data_ptr()
            c.training = int(ctx.training)
            c.dx, c.dweight, c.dbn_weight, c.dbn_bias = _ptr(dx), _ptr(dW), _ptr(dg), _ptr(db)
            c.workspace, c.workspace_bytes = ws.data_ptr(), ws.numel()
            c.stream = torch.cuda.current_stream(dev).cuda_stream
            st = lib.ctc_b200_head_backward(ctypes.byref(c))
            if st != _lib.CTC_STATUS_SUCCESS:
                raise RuntimeError("ctc_b200_head_backward: " + _lib.status_string(lib, st))
        return (dx.view(ctx.x_shape) if dx is not None else None), dW, dg, db, None, None, None, None, None, None


class SequenceWise(nn.Module):
    """Parameter container with the reference's name (model.py:10-40) so that `state_dict()` keys match."""

    def __init__(self, module):
        super().__init__()
        self.module = module


class SequenceWiseClassifier(nn.Module):
    """`SequenceWiseClassifier(in_features, out_features)(x[T, B, H]) -> [B, T, V]` (logits when training, softmax
    probabilities in eval mode), as model.py:205-222.  `forward_time_major` returns the contiguous T x B x V tensor the
    CTC engine reads, without the transposed view."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        fc = nn.Sequential(nn.BatchNorm1d(in_features), nn.Linear(in_features, out_features, bias=False))
        self.fc = nn.Sequential(SequenceWise(fc))

    def forward_time_major(self, x):
        bn, lin = self.fc[0].module[0], self.fc[0].module[1]
        if bn.momentum is None or not bn.track_running_stats:
            raise NotImplementedError("cumulative-average / untracked BatchNorm statistics are not supported")
        out = _HeadFn.apply(x, lin.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, self.training, bn.eps,
                            bn.momentum, not self.training)
        if self.training:
            bn.num_batches_tracked += 1
        return out

    def forward(self, x):
        return self.forward_time_major(x).transpose(0, 1)

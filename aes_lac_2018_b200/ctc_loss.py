"""`CTCLoss` -- drop-in for `warpctc_pytorch.CTCLoss` as the reference uses it.

Reference call sites this mirrors (names, argument meaning, result, error behaviour):
  * construction with defaults  `warp_CTCLoss()`           /root/reference/train.py:179, codes/metrics.py:43
  * `loss = criterion(out, targets, out_sizes, target_sizes)` /root/reference/codes/engine.py:22
  * `self._loss_fn(out, targets, out_sizes, target_sizes).sum()` under no_grad   codes/metrics.py:51
Upstream module semantics (warp-ctc `pytorch_binding/warpctc_pytorch/__init__.py`, not vendored in the
reference): activations are T x B x V and UNNORMALISED (softmax is internal), labels are a flat 1-D int
tensor, lengths are int tensors, blank = 0, the result is `FloatTensor([sum_b cost_b])` of shape [1] on
the CPU, gradients are produced in the forward pass and multiplied by grad_output in backward;
`size_average` divides by B, `length_average` by sum(act_lens) and supersedes it.

The compute path is libctc_b200.so (hand-written sm_100a CUDA) through ctypes.  There is no CPU path and
no PyTorch fallback: CPU activations, or a missing library, raise.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

__all__ = ["CTCLoss", "ctc_loss_raw", "ctc_loss_host"]


def _as_host_int32(x: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if x.requires_grad:
        raise AssertionError(f"{name}: gradients only computed for acts - please mark other tensors as not requiring gradients")
    if x.is_floating_point():
        raise TypeError(f"{name} must be an integer tensor")
    return x.detach().to(device="cpu", dtype=torch.int32).reshape(-1).contiguous()


_dev_ws = {}


def _grow_workspace(lib, key, label_lens_h, act_lens_h, V, B, T, want_grad, device):
    need = ctypes.c_size_t(0)
    st = lib.ctc_b200_workspace_size(label_lens_h.data_ptr(), act_lens_h.data_ptr(), V, B, T,
                                     1 if want_grad else 0, ctypes.byref(need))
    if st != _lib.CTC_STATUS_SUCCESS:
        raise RuntimeError("ctc_b200_workspace_size: " + _lib.status_string(lib, st))
    old = _dev_ws.pop(key, None)
    del old                                                  # release before the larger allocation
    return torch.empty(int(need.value * 1.25) + (1 << 20), dtype=torch.uint8, device=device)


def ctc_loss_raw(acts: torch.Tensor, labels, act_lens, label_lens, blank: int = 0, want_grad: bool = True,
                 grad_scale: float = 1.0, mode: str = "auto", debug: torch.Tensor | None = None,
                 serial_launches: bool = False, timing: dict | None = None, bidirectional: bool = True,
                 no_fallback: bool = False):
    """Runs the CUDA engine once.  Returns (costs[B] float32 CPU tensor, grads[T,B,V] CUDA tensor or None,
    status[B] int32 CPU tensor).  `acts` may be any T x B x V view whose last stride is 1."""
    lib = _lib.load()
    if not acts.is_cuda:
        raise RuntimeError("aes_lac_2018_b200.CTCLoss is CUDA-only (B200-native); got CPU activations. "
                           "There is no CPU fallback.")
    if acts.dim() != 3:
        raise ValueError("acts must be T x B x V")
    if acts.dtype != torch.float32:
        raise TypeError("acts must be float32")
    acts_d = acts.detach()
    if acts_d.stride(2) != 1 and acts_d.size(2) > 1:
        acts_d = acts_d.contiguous()
    T, B, V = acts_d.shape
    labels_h = _as_host_int32(labels, "labels")
    act_lens_h = _as_host_int32(act_lens, "act_lens")
    label_lens_h = _as_host_int32(label_lens, "label_lens")
    if act_lens_h.numel() != B or label_lens_h.numel() != B:
        raise ValueError("act_lens and label_lens must have one entry per utterance (acts.size(1))")
    if int(label_lens_h.sum()) != labels_h.numel():
        raise ValueError("labels must hold exactly sum(label_lens) entries")
    if labels_h.numel() == 0:
        labels_h = torch.zeros(1, dtype=torch.int32)

    with torch.cuda.device(acts_d.device):
        # The workspace is cached per device and only grows; its required size is recomputed (an O(B) host pass)
        # only when the engine reports that the cached one is too small.
        stream_ptr = torch.cuda.current_stream(acts_d.device).cuda_stream
        key = (acts_d.device.index, stream_ptr)               # one workspace per (device, stream): calls on one stream serialise
        workspace = _dev_ws.get(key)
        if workspace is None:
            workspace = _dev_ws[key] = _grow_workspace(lib, key, label_lens_h, act_lens_h, V, B, T, want_grad, acts_d.device)
        grads = torch.empty((T, B, V), dtype=torch.float32, device=acts_d.device) if want_grad else None
        costs = torch.empty(B, dtype=torch.float32)
        status = torch.empty(B, dtype=torch.int32)
        call = _lib.CtcB200Call()
        call.activations = acts_d.data_ptr()
        call.act_stride_t = acts_d.stride(0)
        call.act_stride_b = acts_d.stride(1)
        call.gradients = grads.data_ptr() if want_grad else None
        call.flat_labels = labels_h.data_ptr()
        call.label_lengths = label_lens_h.data_ptr()
        call.input_lengths = act_lens_h.data_ptr()
        call.alphabet_size, call.minibatch, call.max_time = V, B, T
        call.blank_label = int(blank)
        call.grad_scale = float(grad_scale)
        call.costs_host = costs.data_ptr()
        call.costs_device = None
        call.status_host = status.data_ptr()
        call.workspace = workspace.data_ptr()
        call.workspace_bytes = workspace.numel()
        call.stream = stream_ptr
        call.debug_device = debug.data_ptr() if debug is not None else None
        kms = ctypes.c_float(0.0)
        call.kernel_ms_host = ctypes.cast(ctypes.pointer(kms), ctypes.c_void_p) if timing is not None else None
        call.flags = {"auto": 0, "throughput": _lib.FLAG_MODE_THROUGHPUT, "latency": _lib.FLAG_MODE_LATENCY,
                      "throughput8": _lib.FLAG_MODE_THROUGHPUT_K8, "warp": _lib.FLAG_MODE_WARP}[mode]
        if serial_launches:
            call.flags |= _lib.FLAG_SERIAL_LAUNCHES
        if not bidirectional:
            call.flags |= _lib.FLAG_NO_BIDIR
        if no_fallback:
            call.flags |= _lib.FLAG_NO_FALLBACK
        st = lib.ctc_b200_compute(ctypes.byref(call))
        if st != _lib.CTC_STATUS_SUCCESS and b"workspace too small" in lib.ctc_b200_last_error():
            workspace = _dev_ws[key] = _grow_workspace(lib, key, label_lens_h, act_lens_h, V, B, T, want_grad, acts_d.device)
            call.workspace, call.workspace_bytes = workspace.data_ptr(), workspace.numel()
            st = lib.ctc_b200_compute(ctypes.byref(call))
        if st != _lib.CTC_STATUS_SUCCESS:
            raise RuntimeError("ctc_b200_compute: " + _lib.status_string(lib, st))
        if timing is not None:
            timing["kernel_ms"] = float(kms.value)
    return costs, grads, status


_host_ws = {}


def ctc_loss_host(acts: torch.Tensor, labels, act_lens, label_lens, blank: int = 0, want_grad: bool = True,
                  grad_scale: float = 1.0, grads_out: torch.Tensor | None = None, n_chunks: int = 0,
                  device: int | None = None):
    """End-to-end call with HOST buffers: `acts` is a CPU (ideally pinned) float32 T x B x V tensor; costs and
    gradients come back in CPU tensors.  The copies are pipelined against the kernels inside libctc_b200.so
    (ctc_b200_compute_host); the compute is the same sm_100a engine -- there is still no CPU compute path.
    Returns (costs[B], grads[T,B,V] or None, status[B]) as CPU tensors."""
    lib = _lib.load()
    if acts.is_cuda:
        raise ValueError("ctc_loss_host takes host tensors; use CTCLoss / ctc_loss_raw for CUDA tensors")
    if not torch.cuda.is_available():
        raise RuntimeError("aes_lac_2018_b200 needs a CUDA device; there is no CPU fallback")
    if acts.dim() != 3 or acts.dtype != torch.float32:
        raise TypeError("acts must be a float32 T x B x V tensor")
    acts = acts.detach().contiguous()
    T, B, V = acts.shape
    labels_h = _as_host_int32(labels, "labels")
    act_lens_h = _as_host_int32(act_lens, "act_lens")
    label_lens_h = _as_host_int32(label_lens, "label_lens")
    if labels_h.numel() == 0:
        labels_h = torch.zeros(1, dtype=torch.int32)
    dev = torch.cuda.current_device() if device is None else device
    with torch.cuda.device(dev):
        need = ctypes.c_size_t(0)
        st = lib.ctc_b200_workspace_size_host(label_lens_h.data_ptr(), act_lens_h.data_ptr(), V, B, T,
                                              1 if want_grad else 0, int(n_chunks), ctypes.byref(need))
        if st != _lib.CTC_STATUS_SUCCESS:
            raise RuntimeError("ctc_b200_workspace_size_host: " + _lib.status_string(lib, st))
        ws = _host_ws.get(dev)
        if ws is None or ws.numel() < need.value:
            ws = _host_ws[dev] = torch.empty(need.value, dtype=torch.uint8, device=f"cuda:{dev}")
        if want_grad and grads_out is None:
            grads_out = torch.empty((T, B, V), dtype=torch.float32, pin_memory=True)
        costs = torch.empty(B, dtype=torch.float32)
        status = torch.empty(B, dtype=torch.int32)
        call = _lib.CtcB200HostCall()
        call.activations = acts.data_ptr()
        call.gradients = grads_out.data_ptr() if want_grad else None
        call.flat_labels = labels_h.data_ptr()
        call.label_lengths = label_lens_h.data_ptr()
        call.input_lengths = act_lens_h.data_ptr()
        call.alphabet_size, call.minibatch, call.max_time = V, B, T
        call.blank_label = int(blank)
        call.grad_scale = float(grad_scale)
        call.costs_host = costs.data_ptr()
        call.status_host = status.data_ptr()
        call.workspace = ws.data_ptr()
        call.workspace_bytes = ws.numel()
        call.stream = torch.cuda.current_stream(dev).cuda_stream
        call.n_chunks = int(n_chunks)
        call.flags = 0
        st = lib.ctc_b200_compute_host(ctypes.byref(call))
        if st != _lib.CTC_STATUS_SUCCESS:
            raise RuntimeError("ctc_b200_compute_host: " + _lib.status_string(lib, st))
    return costs, (grads_out if want_grad else None), status


class _CTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acts, labels, act_lens, label_lens, size_average=False, length_average=False, blank=0):
        B = acts.size(1)
        if length_average:
            denom = float(torch.as_tensor(act_lens).sum().item())
        elif size_average:
            denom = float(B)
        else:
            denom = 1.0
        want_grad = bool(ctx.needs_input_grad[0])              # False under torch.no_grad() (engine.py:107)
        costs, grads, _ = ctc_loss_raw(acts, labels, act_lens, label_lens, blank=blank, want_grad=want_grad,
                                       grad_scale=1.0 / denom)
        ctx.grads = grads
        total = costs.double().sum() / denom
        return torch.tensor([total], dtype=torch.float32)      # CPU, shape [1], like upstream

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.grads is None:
            raise RuntimeError("CTCLoss.backward called but the forward ran without gradient tracking")
        g = grad_output.to(ctx.grads.device, dtype=ctx.grads.dtype).reshape(-1)[0]
        return ctx.grads.mul_(g), None, None, None, None, None, None


class CTCLoss(torch.nn.Module):
    """`CTCLoss(blank=0, size_average=False, length_average=False)`; see module docstring."""

    def __init__(self, blank: int = 0, size_average: bool = False, length_average: bool = False):
        super().__init__()
        self.ctc = _CTC.apply
        self.blank = blank
        self.size_average = size_average
        self.length_average = length_average

    def forward(self, acts, labels, act_lens, label_lens):
        """acts: T x B x V CUDA float tensor (unnormalised; may be a strided view);
        labels: 1-D int tensor of the concatenated targets; act_lens, label_lens: int tensors [B]."""
        if isinstance(labels, torch.Tensor) and labels.dim() > 1:
            raise AssertionError("labels must be 1 dimensional")
        return self.ctc(acts, labels, act_lens, label_lens, self.size_average, self.length_average, self.blank)

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

__all__ = ["CTCLoss", "ctc_loss_raw", "ctc_loss_host", "sanitize_loss", "reduce_costs", "release_workspaces"]


def _as_host_int32(x: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if x.requires_grad:
        raise AssertionError(f"{name}: gradients only computed for acts - please mark other tensors as not requiring gradients")
    if x.is_floating_point():
        raise TypeError(f"{name} must be an integer tensor")
    return x.detach().to(device="cpu", dtype=torch.int32).reshape(-1).contiguous()


def _check_problem(B: int, labels_h: torch.Tensor, act_lens_h: torch.Tensor, label_lens_h: torch.Tensor) -> torch.Tensor:
    """Shape checks shared by every entry point (the C side reads label_lens / act_lens / labels without bounds):
    one entry per utterance, labels holding exactly sum(label_lens) entries.  Returns the label tensor to pass
    (never empty)."""
    if act_lens_h.numel() != B or label_lens_h.numel() != B:
        raise ValueError("act_lens and label_lens must have one entry per utterance (acts.size(1))")
    if int(label_lens_h.sum()) != labels_h.numel():
        raise ValueError("labels must hold exactly sum(label_lens) entries")
    if labels_h.numel() == 0:
        return torch.zeros(1, dtype=torch.int32)
    return labels_h


_dev_ws = {}
_dev_ws_sig = {}


def _grow_workspace(lib, key, label_lens_h, act_lens_h, V, B, T, want_grad, device):
    need = ctypes.c_size_t(0)
    st = lib.ctc_b200_workspace_size(label_lens_h.data_ptr(), act_lens_h.data_ptr(), V, B, T,
                                     1 if want_grad else 0, ctypes.byref(need))
    if st != _lib.CTC_STATUS_SUCCESS:
        raise RuntimeError("ctc_b200_workspace_size: " + _lib.status_string(lib, st))
    old = _dev_ws.pop(key, None)
    del old                                                  # release before the new allocation
    return torch.empty(int(need.value) + 4096, dtype=torch.uint8, device=device)


def release_workspaces() -> None:
    """Drops every cached device / pinned-host workspace of this process (they are re-created on the next call).
    The cache holds one device workspace per (device, stream); a workspace more than 4x larger than a call needs is
    also replaced by a right-sized one, so one oversized batch does not pin its memory forever."""
    _dev_ws.clear()
    _host_ws.clear()


def ctc_loss_raw(acts: torch.Tensor, labels, act_lens, label_lens, blank: int = 0, want_grad: bool = True,
                 grad_scale: float = 1.0, mode: str = "auto", debug: torch.Tensor | None = None,
                 serial_launches: bool = False, timing: dict | None = None, bidirectional: bool = True,
                 no_fallback: bool = False, no_sync: bool = False):
    """Runs the CUDA engine once.  Returns (costs[B] float32 CPU tensor, grads[T,B,V] CUDA tensor or None,
    status[B] int32 CPU tensor).  `acts` may be any T x B x V view whose last stride is 1.
    `no_sync=True` (CTC_B200_FLAG_NO_SYNC): nothing is read back and the host does not wait -- costs and status come
    back as CUDA tensors, ordered on the current stream like any other kernel output."""
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
    labels_h = _check_problem(B, labels_h, act_lens_h, label_lens_h)

    with torch.cuda.device(acts_d.device):
        # The workspace is cached per device and only grows; its required size is recomputed (an O(B) host pass)
        # only when the engine reports that the cached one is too small.
        stream_ptr = torch.cuda.current_stream(acts_d.device).cuda_stream
        key = (acts_d.device.index, stream_ptr)               # one workspace per (device, stream): calls on one stream serialise
        workspace = _dev_ws.get(key)
        sig = (T, B, V, want_grad, mode)
        if workspace is None:
            workspace = _dev_ws[key] = _grow_workspace(lib, key, label_lens_h, act_lens_h, V, B, T, want_grad, acts_d.device)
        elif _dev_ws_sig.get(key) != sig and workspace.numel() > (64 << 20):
            # a new problem shape on a large cached workspace: let it shrink if it is far too big for this shape
            need = ctypes.c_size_t(0)
            if lib.ctc_b200_workspace_size(label_lens_h.data_ptr(), act_lens_h.data_ptr(), V, B, T, 1 if want_grad else 0,
                                           ctypes.byref(need)) == _lib.CTC_STATUS_SUCCESS and need.value * 4 < workspace.numel():
                workspace = _dev_ws[key] = _grow_workspace(lib, key, label_lens_h, act_lens_h, V, B, T, want_grad, acts_d.device)
        _dev_ws_sig[key] = sig
        grads = torch.empty((T, B, V), dtype=torch.float32, device=acts_d.device) if want_grad else None
        if no_sync:
            costs = torch.empty(B, dtype=torch.float32, device=acts_d.device)
            status = torch.empty(B, dtype=torch.int32, device=acts_d.device)
        else:
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
        call.costs_host = None if no_sync else costs.data_ptr()
        call.costs_device = costs.data_ptr() if no_sync else None
        call.status_host = None if no_sync else status.data_ptr()
        call.status_device = status.data_ptr() if no_sync else None
        call.workspace = workspace.data_ptr()
        call.workspace_bytes = workspace.numel()
        call.stream = stream_ptr
        call.debug_device = debug.data_ptr() if debug is not None else None
        kms = ctypes.c_float(0.0)
        call.kernel_ms_host = ctypes.cast(ctypes.pointer(kms), ctypes.c_void_p) if timing is not None else None
        call.flags = {"auto": 0, "throughput": _lib.FLAG_MODE_THROUGHPUT, "latency": _lib.FLAG_MODE_LATENCY,
                      "throughput8": _lib.FLAG_MODE_THROUGHPUT_K8, "warp": _lib.FLAG_MODE_WARP, "warp32": _lib.FLAG_MODE_WARP32}[mode]
        if serial_launches:
            call.flags |= _lib.FLAG_SERIAL_LAUNCHES
        if not bidirectional:
            call.flags |= _lib.FLAG_NO_BIDIR
        if no_fallback:
            call.flags |= _lib.FLAG_NO_FALLBACK
        if no_sync:
            call.flags |= _lib.FLAG_NO_SYNC
            call.kernel_ms_host = None
        st = lib.ctc_b200_compute(ctypes.byref(call))
        if st != _lib.CTC_STATUS_SUCCESS and b"workspace too small" in lib.ctc_b200_last_error():
            workspace = _dev_ws[key] = _grow_workspace(lib, key, label_lens_h, act_lens_h, V, B, T, want_grad, acts_d.device)
            call.workspace, call.workspace_bytes = workspace.data_ptr(), workspace.numel()
            st = lib.ctc_b200_compute(ctypes.byref(call))
        if st != _lib.CTC_STATUS_SUCCESS:
            raise RuntimeError("ctc_b200_compute: " + _lib.status_string(lib, st))
        if timing is not None:
            timing["kernel_ms"] = float(kms.value)
        if no_sync:
            # pinned host labels / lengths are read by DMA after this returns: keep them alive with the result
            costs._ctc_keepalive = (labels_h, act_lens_h, label_lens_h)
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
    labels_h = _check_problem(B, labels_h, act_lens_h, label_lens_h)
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
        ctx.applied = 1.0                                      # factor the stored gradient already carries
        total = costs.double().sum() / denom
        return torch.tensor([total], dtype=torch.float32)      # CPU, shape [1], like upstream

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.grads is None:
            raise RuntimeError("CTCLoss.backward called but the forward ran without gradient tracking")
        # The result is a CPU tensor (upstream's choice), so grad_output is a host scalar: reading it costs nothing.
        # g == 1 needs no work at all; otherwise one in-place pass by the library's own kernel (upstream: `mul_`).
        # A second backward through the same node (retain_graph) rescales by g / (what is already applied).
        g = float(grad_output.reshape(-1)[0])
        if g != ctx.applied:
            if ctx.applied == 0.0:
                raise RuntimeError("CTCLoss.backward: the stored gradient was zeroed by an earlier backward pass")
            _scale_gradients(ctx.grads, g / ctx.applied)
            ctx.applied = g
        return ctx.grads, None, None, None, None, None, None


def _scale_gradients(grads: torch.Tensor, scale_host: float = 1.0, scale_device: torch.Tensor | None = None,
                     zero_flag: torch.Tensor | None = None) -> None:
    """grads *= scale_host * scale_device (or 0 when zero_flag is set) by ctc_b200_scale_gradients; a factor of
    exactly 1 is a kernel that returns at once (no pass over the tensor)."""
    lib = _lib.load()
    with torch.cuda.device(grads.device):
        st = lib.ctc_b200_scale_gradients(grads.data_ptr(), grads.numel(), float(scale_host),
                                          scale_device.data_ptr() if scale_device is not None else None,
                                          zero_flag.data_ptr() if zero_flag is not None else None,
                                          torch.cuda.current_stream(grads.device).cuda_stream)
    if st != _lib.CTC_STATUS_SUCCESS:
        raise RuntimeError("ctc_b200_scale_gradients: " + _lib.status_string(lib, st))


def reduce_costs(costs: torch.Tensor, scale: float = 1.0, zero_infinity: bool = False):
    """Device-side `scale * costs.sum()` (fp64 accumulation in a fixed order, one tiny kernel on the current stream):
    returns (loss: CUDA float32 tensor of shape [1], flag: CUDA int32 [1], 1 iff `zero_infinity` replaced a +-inf sum
    by 0).  This is the operand of the path's single collective (the scalar NCCL loss sum) -- it never visits the host."""
    lib = _lib.load()
    if not costs.is_cuda or costs.dtype != torch.float32:
        raise TypeError("costs must be a CUDA float32 tensor (ctc_loss_raw(..., no_sync=True) returns one)")
    costs = costs.contiguous()
    dev = costs.device
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    flag = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = lib.ctc_b200_reduce_costs(costs.data_ptr(), costs.numel(), float(scale), 1 if zero_infinity else 0,
                                       loss.data_ptr(), flag.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    if st != _lib.CTC_STATUS_SUCCESS:
        raise RuntimeError("ctc_b200_reduce_costs: " + _lib.status_string(lib, st))
    return loss, flag


class _FusedCTC(torch.autograd.Function):
    """Loss glue fused around the engine call: `scale` (1/average * task weight) goes into the
    gradient epilogue of the kernel, the cost vector is summed and inf-guarded on the device, nothing is read back:
    the step has no host synchronisation and no extra pass over the T x B x V gradient."""

    @staticmethod
    def forward(ctx, acts, labels, act_lens, label_lens, scale, blank, zero_infinity):
        want_grad = bool(ctx.needs_input_grad[0])
        costs, grads, status = ctc_loss_raw(acts, labels, act_lens, label_lens, blank=blank, want_grad=want_grad,
                                            grad_scale=scale, no_sync=True)
        # (the kernel already folded `scale` into the gradient; the costs are unscaled)
        loss, flag = reduce_costs(costs, scale, zero_infinity)
        ctx.grads, ctx.flag, ctx.used = grads, flag, False
        ctx.mark_non_differentiable(status)
        return loss.reshape(()), status

    @staticmethod
    def backward(ctx, grad_loss, _grad_status):
        if ctx.grads is None:
            raise RuntimeError("sanitize_loss: backward called but the forward ran without gradient tracking")
        if ctx.used:
            raise RuntimeError("sanitize_loss: the fused gradient is consumed by its first backward pass "
                               "(it is rescaled in place); use CTCLoss for retain_graph / double backward")
        ctx.used = True
        g = grad_loss.to(device=ctx.grads.device, dtype=torch.float32).reshape(1)
        _scale_gradients(ctx.grads, 1.0, g, ctx.flag)         # returns at once when g == 1 and the guard did not fire
        return ctx.grads, None, None, None, None, None, None


def sanitize_loss(criterion, out, targets, input_percentages, target_sizes, average=1, weight: float = 1.0,
                  time_major: bool = False, return_status: bool = False):
    """Device-side `_sanitize_loss` of the reference (/root/reference/codes/engine.py:12-32) with the task weight of
    engine.py:77 folded in: same arguments (plus `weight`), same value -- `weight * sum_b cost_b / average`, 0 when
    that sum is +-inf -- but returned as a 0-dim CUDA tensor with no host synchronisation, and with the gradient
    already scaled when `.backward()` arrives (no second pass over the T x B x V tensor).

    `out` is the model output B x T x V exactly as engine.py hands it over (it is transposed to a T x B x V view,
    never copied; pass `time_major=True` for a tensor that already is T x B x V); `input_percentages` and
    `target_sizes` are the loader's host tensors.  `criterion` is this package's CTCLoss (it supplies the blank index; its averaging flags
    add to `average` as upstream's would).  Differences from the reference, on purpose: the inf guard yields 0 (the
    reference's `0 * inf` is NaN), and per-utterance status bits stay on the device (`return_status=True` returns
    them as a CUDA int32 tensor; out-of-range utterances are still redone in log space by the device-side detour)."""
    seq_length = out.shape[0] if time_major else out.shape[1]
    acts = out if time_major else out.transpose(0, 1)
    pct = torch.as_tensor(input_percentages)
    act_lens = (pct.detach().to("cpu") * seq_length).int()
    B = acts.shape[1]
    denom = float(average)
    if getattr(criterion, "length_average", False):
        denom *= float(act_lens.sum().item())
    elif getattr(criterion, "size_average", False):
        denom *= float(B)
    blank = int(getattr(criterion, "blank", 0))
    loss, status = _FusedCTC.apply(acts, targets, act_lens, target_sizes, float(weight) / denom, blank, True)
    return (loss, status) if return_status else loss


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

"""ctypes binding of libctc_b200.so (include/ctc.h).  There is NO fallback: if the library is missing
or has not been built for this tree, every entry point raises."""
from __future__ import annotations

import ctypes
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CTC_B200_LIB") or os.path.join(PKG, "lib", "libctc_b200.so")   # (override: kernel experiments)

CTC_STATUS_SUCCESS = 0
CTC_GPU = 1
FLAG_NO_SYNC = 0x1
FLAG_SERIAL_LAUNCHES = 0x2
FLAG_NO_FALLBACK = 0x4
FLAG_NO_BIDIR = 0x8
FLAG_MODE_THROUGHPUT = 1 << 8
FLAG_MODE_LATENCY = 2 << 8
FLAG_MODE_THROUGHPUT_K8 = 3 << 8
FLAG_MODE_WARP = 4 << 8
FLAG_MODE_WARP32 = 5 << 8

UTT_INFEASIBLE, UTT_INF_COST, UTT_BAD_LABEL, UTT_RANGE, UTT_LOGSPACE, UTT_WIDE = 1, 2, 4, 8, 16, 32


class _OptUnion(ctypes.Union):
    _fields_ = [("num_threads", ctypes.c_uint), ("stream", ctypes.c_void_p)]


class CtcOptions(ctypes.Structure):
    """struct ctcOptions of include/ctc.h (passed by value)."""
    _anonymous_ = ("u",)
    _fields_ = [("loc", ctypes.c_int), ("u", _OptUnion), ("blank_label", ctypes.c_int)]


class CtcB200Call(ctypes.Structure):
    """struct ctcB200Call of include/ctc.h."""
    _fields_ = [
        ("activations", ctypes.c_void_p),
        ("act_stride_t", ctypes.c_longlong),
        ("act_stride_b", ctypes.c_longlong),
        ("gradients", ctypes.c_void_p),
        ("flat_labels", ctypes.c_void_p),
        ("label_lengths", ctypes.c_void_p),
        ("input_lengths", ctypes.c_void_p),
        ("alphabet_size", ctypes.c_int),
        ("minibatch", ctypes.c_int),
        ("max_time", ctypes.c_int),
        ("blank_label", ctypes.c_int),
        ("grad_scale", ctypes.c_float),
        ("costs_host", ctypes.c_void_p),
        ("costs_device", ctypes.c_void_p),
        ("status_host", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p),
        ("workspace_bytes", ctypes.c_size_t),
        ("stream", ctypes.c_void_p),
        ("flags", ctypes.c_uint),
        ("debug_device", ctypes.c_void_p),
        ("kernel_ms_host", ctypes.c_void_p),
        ("status_device", ctypes.c_void_p),
    ]


class CtcB200HostCall(ctypes.Structure):
    """struct ctcB200HostCall of include/ctc.h."""
    _fields_ = [
        ("activations", ctypes.c_void_p),
        ("gradients", ctypes.c_void_p),
        ("flat_labels", ctypes.c_void_p),
        ("label_lengths", ctypes.c_void_p),
        ("input_lengths", ctypes.c_void_p),
        ("alphabet_size", ctypes.c_int),
        ("minibatch", ctypes.c_int),
        ("max_time", ctypes.c_int),
        ("blank_label", ctypes.c_int),
        ("grad_scale", ctypes.c_float),
        ("costs_host", ctypes.c_void_p),
        ("status_host", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p),
        ("workspace_bytes", ctypes.c_size_t),
        ("stream", ctypes.c_void_p),
        ("n_chunks", ctypes.c_int),
        ("flags", ctypes.c_uint),
    ]


EXPORTS = ("get_warpctc_version", "ctcGetStatusString", "compute_ctc_loss", "get_workspace_size",
           "ctc_b200_workspace_size", "ctc_b200_compute", "ctc_b200_last_error", "ctc_b200_info",
           "ctc_b200_workspace_size_host", "ctc_b200_compute_host", "ctc_b200_greedy_decode",
           "ctc_b200_edit_distance", "ctc_b200_reduce_costs", "ctc_b200_scale_gradients", "ctc_b200_head_workspace_size", "ctc_b200_head_forward", "ctc_b200_head_backward")

class CtcB200HeadForward(ctypes.Structure):
    _fields_ = [("x", ctypes.c_void_p), ("rows", ctypes.c_int), ("features", ctypes.c_int), ("classes", ctypes.c_int),
                ("weight", ctypes.c_void_p), ("bn_weight", ctypes.c_void_p), ("bn_bias", ctypes.c_void_p),
                ("running_mean", ctypes.c_void_p), ("running_var", ctypes.c_void_p),
                ("eps", ctypes.c_float), ("momentum", ctypes.c_float), ("training", ctypes.c_int), ("softmax", ctypes.c_int),
                ("out", ctypes.c_void_p), ("save_mean", ctypes.c_void_p), ("save_invstd", ctypes.c_void_p),
                ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t), ("stream", ctypes.c_void_p)]


class CtcB200HeadBackward(ctypes.Structure):
    _fields_ = [("x", ctypes.c_void_p), ("dlogits", ctypes.c_void_p), ("rows", ctypes.c_int), ("features", ctypes.c_int),
                ("classes", ctypes.c_int), ("weight", ctypes.c_void_p), ("bn_weight", ctypes.c_void_p),
                ("bn_bias", ctypes.c_void_p), ("save_mean", ctypes.c_void_p), ("save_invstd", ctypes.c_void_p),
                ("training", ctypes.c_int), ("dx", ctypes.c_void_p), ("dweight", ctypes.c_void_p),
                ("dbn_weight", ctypes.c_void_p), ("dbn_bias", ctypes.c_void_p),
                ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t), ("stream", ctypes.c_void_p)]


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -m aes_lac_2018_b200.build`).  There is no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise RuntimeError(f"{LIB_PATH} does not export {name}; rebuild it")
    lib.get_warpctc_version.restype = ctypes.c_int
    lib.ctcGetStatusString.restype = ctypes.c_char_p
    lib.ctcGetStatusString.argtypes = [ctypes.c_int]
    lib.ctc_b200_last_error.restype = ctypes.c_char_p
    lib.ctc_b200_info.restype = ctypes.c_int
    lib.ctc_b200_info.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
    lib.compute_ctc_loss.restype = ctypes.c_int
    lib.compute_ctc_loss.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, CtcOptions]
    lib.get_workspace_size.restype = ctypes.c_int
    lib.get_workspace_size.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, CtcOptions,
        ctypes.POINTER(ctypes.c_size_t)]
    lib.ctc_b200_workspace_size.restype = ctypes.c_int
    lib.ctc_b200_workspace_size.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
        ctypes.POINTER(ctypes.c_size_t)]
    lib.ctc_b200_workspace_size_host.restype = ctypes.c_int
    lib.ctc_b200_workspace_size_host.argtypes = [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
        ctypes.POINTER(ctypes.c_size_t)]
    lib.ctc_b200_compute_host.restype = ctypes.c_int
    lib.ctc_b200_compute_host.argtypes = [ctypes.POINTER(CtcB200HostCall)]
    lib.ctc_b200_greedy_decode.restype = ctypes.c_int
    lib.ctc_b200_greedy_decode.argtypes = [
        ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
        ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.ctc_b200_edit_distance.restype = ctypes.c_int
    lib.ctc_b200_edit_distance.argtypes = [
        ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p]
    lib.ctc_b200_head_workspace_size.restype = ctypes.c_int
    lib.ctc_b200_head_workspace_size.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]
    lib.ctc_b200_head_forward.restype = ctypes.c_int
    lib.ctc_b200_head_forward.argtypes = [ctypes.POINTER(CtcB200HeadForward)]
    lib.ctc_b200_head_backward.restype = ctypes.c_int
    lib.ctc_b200_head_backward.argtypes = [ctypes.POINTER(CtcB200HeadBackward)]
    lib.ctc_b200_reduce_costs.restype = ctypes.c_int
    lib.ctc_b200_reduce_costs.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]
    lib.ctc_b200_scale_gradients.restype = ctypes.c_int
    lib.ctc_b200_scale_gradients.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p]
    lib.ctc_b200_compute.restype = ctypes.c_int
    lib.ctc_b200_compute.argtypes = [ctypes.POINTER(CtcB200Call)]
    _lib = lib
    return lib


def status_string(lib, st: int) -> str:
    return f"{lib.ctcGetStatusString(st).decode()} ({lib.ctc_b200_last_error().decode()})"


def launch_count() -> int:
    n = ctypes.c_ulonglong(0)
    load().ctc_b200_info(ctypes.byref(n))
    return int(n.value)

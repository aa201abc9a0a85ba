"""ctypes front-end of oracle/warpctc_cpu.c (fp32 OpenMP restatement of warp-ctc's CPU path).

TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as oracle/ctc_f64.py).  PARITY UNPINNED by the
reference; see the header of warpctc_cpu.c.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libwarpctc_cpu_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the restatement with the committed recipe (oracle/Makefile)."""
    src = os.path.join(_HERE, "warpctc_cpu.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        lib.oracle_warpctc_cpu.restype = ctypes.c_int
        lib.oracle_warpctc_cpu.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.oracle_warpctc_cpu_max_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(_load().oracle_warpctc_cpu_max_threads())


def ctc_batch(acts, flat_labels, act_lens, label_lens, blank: int = 0, num_threads: int = 0,
              want_grad: bool = True):
    """acts [T,B,V] float32 -> (costs[B] float32, grads[T,B,V] float32 or None)."""
    acts = np.ascontiguousarray(acts, dtype=np.float32)
    T, B, V = acts.shape
    labels = np.ascontiguousarray(np.asarray(flat_labels).reshape(-1), dtype=np.int32)
    if labels.size == 0:
        labels = np.zeros(1, dtype=np.int32)
    al = np.ascontiguousarray(act_lens, dtype=np.int32)
    ll = np.ascontiguousarray(label_lens, dtype=np.int32)
    costs = np.zeros(B, dtype=np.float32)
    grads = np.zeros_like(acts) if want_grad else None
    st = _load().oracle_warpctc_cpu(
        acts.ctypes.data, grads.ctypes.data if want_grad else None, labels.ctypes.data,
        ll.ctypes.data, al.ctypes.data, V, B, costs.ctypes.data, int(blank), int(num_threads))
    if st != 0:
        raise RuntimeError(f"oracle_warpctc_cpu failed with status {st}")
    return costs, grads

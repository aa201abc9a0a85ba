"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ctc.h declares, validates arguments like upstream, and the Python front-end refuses to run
without CUDA (no fallback).  No compute calls are made here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from aes_lac_2018_b200 import _lib, build
    build.build()                                   # nvcc cross-compiles for sm_100a without a GPU
    return _lib.load()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "ctc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_header_symbols_are_exported(lib):
    names = _declared_functions()
    assert {"compute_ctc_loss", "get_workspace_size", "ctcGetStatusString", "get_warpctc_version",
            "ctc_b200_compute", "ctc_b200_workspace_size", "ctc_b200_last_error", "ctc_b200_info"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"include/ctc.h declares {n} but libctc_b200.so does not export it"


def test_version_and_status_strings(lib):
    assert lib.get_warpctc_version() == 2
    assert lib.ctcGetStatusString(0) == b"no error"
    for st in range(1, 5):
        assert lib.ctcGetStatusString(st)
    assert lib.ctc_b200_info(None) == 100


def _opts(loc=1, blank=0):
    from aes_lac_2018_b200 import _lib
    o = _lib.CtcOptions()
    o.loc, o.stream, o.blank_label = loc, None, blank
    return o


def test_struct_layout_matches_header():
    from aes_lac_2018_b200 import _lib
    # struct ctcOptions { enum loc; union { unsigned; CUstream }; int blank_label; } -> 4 + pad 4 + 8 + 4 + pad 4
    assert ctypes.sizeof(_lib.CtcOptions) == 24
    assert _lib.CtcOptions.u.offset == 8 and _lib.CtcOptions.blank_label.offset == 16
    assert _lib.CtcB200Call.flags.offset + 4 <= ctypes.sizeof(_lib.CtcB200Call)


def test_workspace_size_and_argument_validation(lib):
    ll = np.array([10, 200, 0], np.int32)
    al = np.array([50, 750, 3], np.int32)
    n = ctypes.c_size_t(0)
    assert lib.get_workspace_size(ll.ctypes.data, al.ctypes.data, 29, 3, _opts(), ctypes.byref(n)) == 0
    assert n.value > 0
    small = n.value
    ll2 = np.array([10, 200, 0] * 64, np.int32)
    al2 = np.array([50, 750, 3] * 64, np.int32)
    assert lib.get_workspace_size(ll2.ctypes.data, al2.ctypes.data, 29, 192, _opts(), ctypes.byref(n)) == 0
    assert n.value > small
    # invalid arguments -> CTC_STATUS_INVALID_VALUE (2), never an exception
    assert lib.get_workspace_size(None, al.ctypes.data, 29, 3, _opts(), ctypes.byref(n)) == 2
    assert lib.get_workspace_size(ll.ctypes.data, al.ctypes.data, 0, 3, _opts(), ctypes.byref(n)) == 2
    assert lib.get_workspace_size(ll.ctypes.data, al.ctypes.data, 29, 0, _opts(), ctypes.byref(n)) == 2
    assert lib.get_workspace_size(ll.ctypes.data, al.ctypes.data, 29, 3, _opts(), None) == 2
    # GPU-only library: the CPU location is refused loudly
    assert lib.get_workspace_size(ll.ctypes.data, al.ctypes.data, 29, 3, _opts(loc=0), ctypes.byref(n)) == 2
    assert b"GPU-only" in lib.ctc_b200_last_error()
    # label sequences beyond the supported length -> CTC_STATUS_UNKNOWN_ERROR (4), like upstream's S > 1280
    big = np.array([2048], np.int32)
    assert lib.get_workspace_size(big.ctypes.data, np.array([9000], np.int32).ctypes.data, 29, 1, _opts(), ctypes.byref(n)) == 4
    ok = np.array([2047], np.int32)
    assert lib.get_workspace_size(ok.ctypes.data, np.array([4100], np.int32).ctypes.data, 29, 1, _opts(), ctypes.byref(n)) == 0


def test_head_workspace_size_and_argument_validation(lib):
    """ctc_b200_head_workspace_size is pure host arithmetic: sizes for every class padding (32 / 48 / 64), growth with the
    problem, and the argument checks of the head entry points (no CUDA call is reached)."""
    n = ctypes.c_size_t(0)
    sizes = {}
    for V in (29, 43, 64):
        assert lib.ctc_b200_head_workspace_size(192000, 800, V, ctypes.byref(n)) == 0
        sizes[V] = n.value
        # forward weight tiles (2 * H * VP floats) + input-gradient weight tiles + the row-block partials must fit
        assert n.value > 4 * 800 * 32 * 2
    assert sizes[29] < sizes[43] <= sizes[64]
    assert lib.ctc_b200_head_workspace_size(16, 8, 2, ctypes.byref(n)) == 0 and 0 < n.value < sizes[29]
    assert lib.ctc_b200_head_workspace_size(100, 802, 29, ctypes.byref(n)) == 2          # features % 4 != 0
    assert lib.ctc_b200_head_workspace_size(0, 800, 29, ctypes.byref(n)) == 2
    assert lib.ctc_b200_head_workspace_size(100, 800, 29, None) == 2
    assert lib.ctc_b200_head_workspace_size(100, 800, 65, ctypes.byref(n)) == 4          # more than 64 classes: not in this build
    assert b"64" in lib.ctc_b200_last_error()
    assert lib.ctc_b200_head_forward(None) == 2 and lib.ctc_b200_head_backward(None) == 2


def test_compute_rejects_bad_arguments_before_touching_cuda(lib):
    ll = np.array([2], np.int32)
    al = np.array([5], np.int32)
    lab = np.array([1, 2], np.int32)
    costs = np.zeros(1, np.float32)
    assert lib.compute_ctc_loss(None, None, lab.ctypes.data, ll.ctypes.data, al.ctypes.data, 5, 1,
                                costs.ctypes.data, None, _opts()) == 2
    dummy = ctypes.c_void_p(16)                      # never dereferenced: validation fails first
    assert lib.compute_ctc_loss(dummy, None, lab.ctypes.data, ll.ctypes.data, al.ctypes.data, 5, 1,
                                costs.ctypes.data, dummy, _opts(loc=0)) == 2
    assert lib.compute_ctc_loss(dummy, None, lab.ctypes.data, ll.ctypes.data, al.ctypes.data, -1, 1,
                                costs.ctypes.data, dummy, _opts()) == 2


def test_python_front_end_has_no_cpu_fallback():
    from aes_lac_2018_b200 import CTCLoss
    acts = torch.randn(5, 1, 5, requires_grad=True)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        CTCLoss()(acts, torch.tensor([1, 2], dtype=torch.int32), torch.tensor([5], dtype=torch.int32),
                  torch.tensor([2], dtype=torch.int32))
    with pytest.raises(AssertionError):
        CTCLoss()(acts, torch.zeros(2, 2, dtype=torch.int32), torch.tensor([5]), torch.tensor([2]))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from aes_lac_2018_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure; a product path routed through it would void every parity claim."""
    pkg = os.path.join(ROOT, "aes_lac_2018_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("float64 oracle", "").replace("the oracle", ""), f
    alias = open(os.path.join(ROOT, "warpctc_pytorch", "__init__.py")).read()
    assert "oracle" not in alias


def test_every_call_struct_matches_the_c_header_field_by_field(tmp_path):
    """include/ctc.h is compiled as plain C (gcc) and every field offset of the call structs is compared with the ctypes
    mirror in aes_lac_2018_b200/_lib.py -- a mismatch would silently shift pointers in the binding."""
    import shutil
    import subprocess
    from aes_lac_2018_b200 import _lib
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if not gcc:
        pytest.skip("no C compiler")
    pairs = [("ctcB200Call", _lib.CtcB200Call), ("ctcB200HostCall", _lib.CtcB200HostCall),
             ("ctcB200HeadForward", _lib.CtcB200HeadForward), ("ctcB200HeadBackward", _lib.CtcB200HeadBackward)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ctc.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append(f'printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out.strip().splitlines()}
    for cname, cls in pairs:
        assert got[(cname, "sizeof")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, f"{cname}.{fname}"

"""Compiles libctc_b200.so (hand-written sm_100a CUDA + the C ABI of include/ctc.h) in-tree.

    python -m aes_lac_2018_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but travels with gpurun.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libctc_b200.so")
N_GROUPS = 10
HEADERS = ["ctc_fused.cuh", "ctc_warp.cuh", "ctc_warp32.cuh", "ctc_logspace.cuh", "ctc_decode.cuh", "ctc_combine.cuh", "ctc_editdist.cuh", "ctc_variants.h",
           "ctc_variants.cu", "ctc_abi.cu", "ctc_head.cu", "ctc_head_tc.cuh", "ctc_head_bwd_tc.cuh", "ctc_internal.h", os.path.join(ROOT, "include", "ctc.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas=-v",
]   # no --use_fast_math: fast intrinsics are chosen explicitly, per call site, in the kernels


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libctc_b200.so cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(args):
    cmd, log = args
    proc = subprocess.run(cmd, capture_output=True, text=True)
    return proc.returncode, " ".join(cmd), proc.stdout + proc.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the variant groups in parallel (one nvcc per group), then link libctc_b200.so."""
    if not force and not is_stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(PKG, "build")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    jobs, objs = [], []
    for g in range(N_GROUPS):
        o = os.path.join(obj_dir, f"ctc_variants_g{g}.o")
        objs.append(o)
        jobs.append(([nvcc, *NVCC_FLAGS, f"-DCTC_GROUP={g}", "-c", os.path.join(CSRC, "ctc_variants.cu"), "-o", o], None))
    o = os.path.join(obj_dir, "ctc_abi.o")
    objs.append(o)
    jobs.append(([nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, "ctc_abi.cu"), "-o", o], None))
    o = os.path.join(obj_dir, "ctc_head.o")
    objs.append(o)
    jobs.append(([nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, "ctc_head.cu"), "-o", o], None))
    logs = []
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        for rc, cmd, out in ex.map(_compile, jobs):
            logs.append(f"$ {cmd}\n{out}")
            if rc != 0:
                raise RuntimeError("nvcc failed:\n" + cmd + "\n" + out)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB + ".tmp", *objs]
    rc, cmd, out = _compile((link, None))
    if rc != 0:
        raise RuntimeError("link failed:\n" + cmd + "\n" + out)
    os.replace(LIB + ".tmp", LIB)
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

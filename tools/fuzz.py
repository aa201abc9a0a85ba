"""Randomised parity fuzz (larger and wider than tests/test_gpu_random_shapes.py): random V, T, B, L up to 700,
ragged lengths, repeats, logit scales, blank positions, every ladder -- against the float64 oracle.
python tools/fuzz.py [n_cases] [seed0]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aes_lac_2018_b200 import ctc_loss_raw
from oracle import ctc_f64

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
worst = dict(loss=0.0, grad=0.0)
bad = 0
t0 = time.time()
for case in range(n_cases):
    rng = np.random.default_rng(seed0 + case)
    V = int(rng.choice([1, 2, 3, 5, 17, 29, 29, 32, 33, 43, 63, 64]))
    lmax = int(rng.choice([0, 1, 7, 31, 32, 64, 100, 130, 200, 260, 400, 700]))
    T = int(rng.integers(max(1, lmax // 2), 2 * lmax + 60)) if rng.random() < 0.7 else int(rng.integers(1, 1600))
    B = int(rng.integers(1, 9)) if lmax > 130 else int(rng.integers(1, 40))
    blank = int(rng.choice([0, 0, V - 1, rng.integers(0, V)]))
    al = rng.integers(max(1, T // 2), T + 1, B).astype(np.int32); al[rng.integers(0, B)] = T
    ll = rng.integers(0, lmax + 1, B).astype(np.int32) if V > 1 else np.zeros(B, np.int32)
    syms = np.array([k for k in range(V) if k != blank])
    labels = rng.choice(syms, int(ll.sum())).astype(np.int32) if V > 1 else np.zeros(0, np.int32)
    if labels.size > 3 and rng.random() < 0.5:
        idx = rng.integers(1, labels.size, labels.size // 3); labels[idx] = labels[idx - 1]
    sigma = float(rng.choice([0.3, 1.0, 2.0, 4.0]))
    acts = (rng.standard_normal((T, B, V)) * sigma).astype(np.float32)
    if rng.random() < 0.3:
        acts[..., blank] += float(rng.choice([2.0, 4.0]))
    oc, og = ctc_f64.ctc_batch(acts, labels, al, ll, blank)
    a = torch.tensor(acts).cuda()
    args = [torch.tensor(x) for x in (labels, al, ll)]
    for mode, bidir in (("auto", True), ("warp32", False), ("warp", False), ("throughput", False), ("throughput8", False), ("latency", True), ("latency", False)):
        c, g, st = ctc_loss_raw(a, *args, blank=blank, mode=mode, bidirectional=bidir)
        c = c.numpy().astype(np.float64); g = g.cpu().numpy().astype(np.float64)
        fin = np.isfinite(oc)
        el = float((np.abs(c[fin] - oc[fin]) / np.maximum(1.0, np.abs(oc[fin]))).max()) if fin.any() else 0.0
        eg = float(np.abs(g - og).max())
        same_inf = bool((np.isinf(c) == np.isinf(oc)).all())
        worst["loss"] = max(worst["loss"], el); worst["grad"] = max(worst["grad"], eg)
        if (el > 2e-6 or eg > float(os.environ.get('FUZZ_GRAD_NOTE', '1'))) and os.environ.get('FUZZ_VERBOSE'):
            print(f"note case {seed0 + case} mode {mode} bidir {bidir}: V {V} T {T} B {B} lmax {lmax} sigma {sigma} loss err {el:.2e} grad err {eg:.2e} costs {oc[:4].round(3).tolist()} got {c[:4].round(3).tolist()} status {sorted(set(st.tolist()))}", flush=True)
        if el > 1e-4 or eg > 1e-5 or not same_inf or not np.isfinite(g).all():
            bad += 1
            print(f"FAIL case {seed0 + case} mode {mode} bidir {bidir}: V {V} T {T} B {B} L {ll.tolist()} T_b {al.tolist()} blank {blank} sigma {sigma} "
                  f"loss err {el:.2e} grad err {eg:.2e} status {sorted(set(st.tolist()))}", flush=True)
print(f"{n_cases} cases x 7 paths in {time.time() - t0:.0f}s: failures {bad}, worst loss rel err {worst['loss']:.2e}, worst grad abs err {worst['grad']:.2e}")

"""Re-runs one case of tools/fuzz.py through a given mode and prints per-utterance errors and the self-check words
(development aid): python tools/w32_case.py CASE [mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aes_lac_2018_b200 import ctc_loss_raw
from oracle import ctc_f64

case = int(sys.argv[1]); mode = sys.argv[2] if len(sys.argv) > 2 else "warp32"
rng = np.random.default_rng(case)
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
for nofb in (True, False):
    dbg = torch.zeros(B * 16, dtype=torch.int64, device="cuda")
    c, g, st = ctc_loss_raw(a, *args, blank=blank, mode=mode, no_fallback=nofb, debug=dbg)
    d = dbg.cpu().numpy().reshape(-1, 16)
    gerr = np.abs(g.cpu().numpy() - og).max(axis=(0, 2))
    off = np.concatenate(([0], np.cumsum(ll)))
    print(f"case {case} mode {mode} no_fallback {nofb}: V {V} T {T} B {B} blank {blank} sigma {sigma}")
    for b in range(B):
        rep = int((labels[off[b] + 1:off[b + 1]] == labels[off[b]:off[b + 1] - 1]).sum()) if ll[b] > 1 else 0
        f = lambda x: np.array([x], dtype=np.uint32).view(np.float32)[0]
        print(f"  b={b} L={ll[b]} rep={rep} T={al[b]} cost {oc[b]:.2f} got {c[b].item():.2f} status {st[b].item()} chk {f(d[b, 0]):.2e} pmax {f(d[b, 1]):.3f} "
              f"hmax {f(d[b, 2]):.3f} grad err {gerr[b]:.2e}")

"""Probe the fp64 dynamic-range self-check with hostile inputs: large logit scales, confident-but-wrong
(peaky on the wrong symbols), long utterances.  Prints status bits and error vs the float64 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aes_lac_2018_b200 import _lib
from aes_lac_2018_b200.ctc_loss import ctc_loss_raw
from oracle import ctc_f64
from tests.helpers import synth_problem

def run(name, acts, labels, al, ll):
    lib = _lib.load()
    try:
        costs, grads, status = ctc_loss_raw(torch.tensor(acts).cuda(), torch.tensor(labels), torch.tensor(al), torch.tensor(ll))
        st = status.numpy()
    except RuntimeError as e:
        print(name, "RAISED", str(e)[:120]); return
    oc, og = ctc_f64.ctc_batch(acts, labels, al, ll)
    c = costs.numpy().astype(np.float64); g = grads.cpu().numpy().astype(np.float64)
    fin = np.isfinite(oc)
    rel = np.abs(c[fin] - oc[fin]) / np.maximum(1, np.abs(oc[fin]))
    err = np.abs(g - og).max(axis=(0, 2))
    unfl = err[(st & 0x10) == 0]
    print(name, "status", sorted(set(st.tolist())), "cost", oc[:3].round(1), "rel %.2e" % (rel.max() if rel.size else 0),
          "grad err %.2e" % err.max(), "| worst UNFLAGGED %.2e" % (unfl.max() if unfl.size else 0), "flagged", int(((st & 0x10) != 0).sum()), "/", len(st))

for sigma in (5, 10, 12, 14, 16, 18, 20, 22, 25, 30, 40, 80):
    acts, labels, al, ll = synth_problem(100 + sigma, 300, 16, 29, 40, 100, sigma=float(sigma))
    run(f"sigma{sigma}", acts, labels, al, ll)
# confident and wrong: every frame puts +25 on a symbol that is NOT in the transcript order
rng = np.random.default_rng(0)
T, B, V = 750, 2, 29
acts = rng.standard_normal((T, B, V)).astype(np.float32)
wrong = rng.integers(1, V, (T, B))
for t in range(T):
    for b in range(B):
        acts[t, b, wrong[t, b]] += 25.0
ll = np.array([200, 120], np.int32); al = np.array([T, T], np.int32)
labels = rng.integers(1, V, int(ll.sum())).astype(np.int32)
run("confident_wrong_25", acts, labels, al, ll)
acts2 = acts.copy(); acts2[acts2 > 10] += 40.0
run("confident_wrong_65", acts2, labels, al, ll)
# blank-saturated model asked for many labels
acts3 = rng.standard_normal((T, B, V)).astype(np.float32); acts3[..., 0] += 30.0
run("blank_saturated_30", acts3, labels, al, ll)

for margin in (4, 6, 8, 10, 12, 15, 20):
    acts3 = rng.standard_normal((T, 8, V)).astype(np.float32); acts3[..., 0] += float(margin)
    ll8 = rng.integers(60, 201, 8).astype(np.int32); al8 = np.full(8, T, np.int32)
    lab8 = rng.integers(1, V, int(ll8.sum())).astype(np.int32)
    run(f"blank_margin_{margin}", acts3, lab8, al8, ll8)

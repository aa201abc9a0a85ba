"""Distribution of the forward/backward consistency residual (|sum posterior - n| per chunk, max over chunks)
on benign inputs, per ladder: used to place the fallback threshold."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aes_lac_2018_b200 import ctc_loss_raw
from tests.helpers import synth_problem
from tests.test_gpu_parity import SYNTH
cases = dict(SYNTH)
cases["big_c2"] = dict(seed=5, T=750, B=64, V=29, lmin=50, lmax=200)
cases["peaky_long"] = dict(seed=6, T=750, B=32, V=29, lmin=50, lmax=200, peaky=True)
for name, kw in cases.items():
    acts, labels, al, ll = synth_problem(**kw)
    a = torch.tensor(acts).cuda()
    for mode in ("throughput", "throughput8", "latency"):
        dbg = torch.zeros(acts.shape[1], 16, dtype=torch.int64, device="cuda")
        try:
            ctc_loss_raw(a, torch.tensor(labels), torch.tensor(al), torch.tensor(ll), mode=mode, debug=dbg)
        except RuntimeError as e:
            print(name, mode, "raised", str(e)[:80]); continue
        v = dbg[:, 15].cpu().numpy() * 1e-9
        print(f"{name:22s} {mode:11s} chk max {v.max():.2e}  p90 {np.percentile(v, 90):.2e}  median {np.median(v):.2e}")

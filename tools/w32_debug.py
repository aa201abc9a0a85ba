"""What the self-checks of the fp32 warp kernel saw (development aid): python tools/w32_debug.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aes_lac_2018_b200 import ctc_loss_raw
from tests.helpers import synth_problem
from oracle import ctc_f64

cases = {"sigma4_v43": dict(seed=15, T=400, B=6, V=43, lmin=30, lmax=120, sigma=4.0),
         "sigma3": dict(seed=17, T=750, B=16, V=29, lmin=50, lmax=200, sigma=3.0),
         "sigma6": dict(seed=18, T=300, B=16, V=29, lmin=20, lmax=100, sigma=6.0)}
for name, kw in cases.items():
    acts, labels, al, ll = synth_problem(**kw)
    oc, og = ctc_f64.ctc_batch(acts, labels, al, ll)
    dbg = torch.zeros(acts.shape[1] * 16, dtype=torch.int64, device="cuda")
    c, g, st = ctc_loss_raw(torch.tensor(acts).cuda(), torch.tensor(labels), torch.tensor(al), torch.tensor(ll), mode="warp32",
                            no_fallback=True, debug=dbg)
    d = dbg.cpu().numpy().reshape(-1, 16)
    gerr = np.abs(g.cpu().numpy() - og).max(axis=(0, 2))
    for b in range(acts.shape[1]):
        chk = np.array([d[b, 0]], dtype=np.uint32).view(np.float32)[0]
        pm = np.array([d[b, 1]], dtype=np.uint32).view(np.float32)[0]
        hm = np.array([d[b, 2]], dtype=np.uint32).view(np.float32)[0]
        print(f"{name} b={b} L={ll[b]} T={al[b]} cost {oc[b]:.1f} status {st[b].item()} chk_dev {chk:.3e} pmax {pm:.4f} hmax {hm:.4f} grad err {gerr[b]:.2e}")

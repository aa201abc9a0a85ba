"""Very long transcripts (the W=8 variants): status of the fast path alone (no log-space detour), its
self-check value and its error against the float64 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aes_lac_2018_b200.ctc_loss import ctc_loss_raw
from oracle import ctc_f64
from tests.helpers import synth_problem

for (L, T) in ((900, 1000), (1000, 2100), (1100, 1200), (1100, 2300), (1500, 1600), (1500, 3000), (2047, 2200), (2047, 4000)):
    acts, labels, al, ll = synth_problem(7 + L, T, 1, 29, L, L)
    oc, og = ctc_f64.ctc_batch(acts, labels, al, ll)
    for mode in ("throughput", "throughput8", "latency"):
        dbg = torch.zeros(1, 16, dtype=torch.int64, device="cuda")
        c, g, st = ctc_loss_raw(torch.tensor(acts).cuda(), torch.tensor(labels), torch.tensor(al), torch.tensor(ll),
                                mode=mode, debug=dbg, no_fallback=True, bidirectional=False)
        err = np.abs(g.cpu().numpy().astype(np.float64) - og).max()
        print(f"L={L} T={T} {mode:11s} status={st.tolist()} chk={dbg[0, 11].item() / 1e9:.3e} cost {c.item():.3f} vs {oc[0]:.3f}  grad err {err:.2e}", flush=True)

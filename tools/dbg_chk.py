import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aes_lac_2018_b200 import ctc_loss_raw
from tests.helpers import synth_problem
acts, labels, al, ll = synth_problem(seed=11, T=200, B=4, V=29, lmin=10, lmax=50)
a = torch.tensor(acts).cuda()
for mode in ("throughput", "throughput8"):
    dbg = torch.zeros(4, 16, dtype=torch.int64, device="cuda")
    c, g, st = ctc_loss_raw(a, torch.tensor(labels), torch.tensor(al), torch.tensor(ll), mode=mode, debug=dbg)
    print(mode, "status", st.tolist(), "L", ll.tolist())
    print(dbg[:, 12:16].cpu().numpy())

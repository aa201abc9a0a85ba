"""Runs the engine a few times on one configuration (for ncu).  python tools/profile_one.py B T V lmin lmax mode calls"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from aes_lac_2018_b200 import ctc_loss_raw

B, T, V, lmin, lmax = (int(x) for x in sys.argv[1:6])
mode = sys.argv[6] if len(sys.argv) > 6 else "auto"
calls = int(sys.argv[7]) if len(sys.argv) > 7 else 3
g = torch.Generator().manual_seed(1234)
acts = torch.randn(T, B, V, generator=g).cuda()
ll = torch.randint(lmin, lmax + 1, (B,), generator=g, dtype=torch.int32)
al = torch.full((B,), T, dtype=torch.int32)
labels = torch.randint(1, V, (int(ll.sum()),), generator=g, dtype=torch.int32)
for _ in range(calls):
    costs, grads, status = ctc_loss_raw(acts, labels, al, ll, mode=mode)
torch.cuda.synchronize()
print("loss", float(costs.sum()), "status", int(status.max()))

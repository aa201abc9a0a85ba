"""Ladder crossover: device time of one engine call per batch size for the warp ladder, the latency ladder (with its
bidirectional path for B <= 96) and the automatic choice.   python tools/crossover.py [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aes_lac_2018_b200 import ctc_loss_raw

T = int(sys.argv[1]) if len(sys.argv) > 1 else 750
V = 29
for B in (32, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048, 4096):
    g = torch.Generator().manual_seed(1234)
    acts = torch.randn(T, B, V, generator=g).cuda()
    ll = torch.randint(50, 201, (B,), generator=g, dtype=torch.int32)
    al = torch.full((B,), T, dtype=torch.int32)
    labels = torch.randint(1, V, (int(ll.sum()),), generator=g, dtype=torch.int32)
    row = []
    for mode in ("warp32", "warp", "latency", "auto"):
        for _ in range(3):
            ctc_loss_raw(acts, labels, al, ll, mode=mode)
        best = 1e9
        for _ in range(5):
            tm = {}
            ctc_loss_raw(acts, labels, al, ll, mode=mode, timing=tm)
            best = min(best, tm["kernel_ms"])
        row.append(f"{mode} {best:.3f} ms")
    print(f"T={T} B={B:5d}: " + " | ".join(row), flush=True)

"""Where the non-kernel time of one engine call goes at B=8192: pageable vs pinned labels, kernel ms vs call ms."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aes_lac_2018_b200 import ctc_loss_raw
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator().manual_seed(1234)
acts = torch.randn(750, B, 29, generator=g).cuda()
ll = torch.randint(50, 201, (B,), generator=g, dtype=torch.int32)
al = torch.full((B,), 750, dtype=torch.int32)
labels = torch.randint(1, 29, (int(ll.sum()),), generator=g, dtype=torch.int32)
for name, lab in (("pageable labels", labels), ("pinned labels", labels.pin_memory()), ("pageable labels", labels), ("pinned labels", labels.pin_memory())):
    tm = {}
    for _ in range(3): ctc_loss_raw(acts, lab, al, ll, timing=tm)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ks = 0.0
    for _ in range(10):
        ctc_loss_raw(acts, lab, al, ll, timing=tm); ks += tm["kernel_ms"]
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10 * 1e3
    print(f"{name:16s}: call {dt:.3f} ms, kernels {ks / 10:.3f} ms, outside {dt - ks / 10:.3f} ms", flush=True)

"""Per-kernel times of the classifier head at T=750, B=256, H=800 from torch's profiler.  python tools/head_ab.py [V]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from aes_lac_2018_b200 import SequenceWiseClassifier
T, B, H, V = 750, 256, 800, int(sys.argv[1]) if len(sys.argv) > 1 else 29
x = torch.randn(T, B, H, device="cuda").requires_grad_(True)
head = SequenceWiseClassifier(H, V).cuda().train()
dl = torch.randn(T, B, V, device="cuda")
def step():
    out = head.forward_time_major(x)
    out.backward(dl)
    x.grad = None
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10): step()
    torch.cuda.synchronize()
tot = 0.0
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if "head_" in e.key:
        print(f"  {e.key[:60]:60s} {e.device_time_total / e.count:8.1f} us")
        tot += e.device_time_total / e.count
print(f"  head kernels total {tot:8.1f} us   env: " + " ".join(k for k in os.environ if k.startswith("CTC_B200_HEAD")))

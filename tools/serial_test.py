import sys, torch
sys.path.insert(0, "/root/repo")
from aes_lac_2018_b200 import ctc_loss_raw
g = torch.Generator().manual_seed(1234)
B = 8192
acts = torch.randn(750, B, 29, generator=g).cuda()
ll = torch.randint(50, 201, (B,), generator=g, dtype=torch.int32)
al = torch.full((B,), 750, dtype=torch.int32)
labels = torch.randint(1, 29, (int(ll.sum()),), generator=g, dtype=torch.int32)
for serial in (False, True, False, True):
    for _ in range(3): ctc_loss_raw(acts, labels, al, ll, mode="throughput8", serial_launches=serial)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): ctc_loss_raw(acts, labels, al, ll, mode="throughput8", serial_launches=serial)
    e1.record(); torch.cuda.synchronize()
    print("serial" if serial else "forked", e0.elapsed_time(e1) / 8, "ms")

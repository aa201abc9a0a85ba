"""The classifier-head kernels on a small and on a many-stage problem, meant to run under compute-sanitizer
(memcheck / racecheck): forward (training and eval + softmax), backward, every class padding (32 / 48 / 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aes_lac_2018_b200 import SequenceWiseClassifier

for (T, B, H, V) in ((50, 4, 96, 29), (37, 3, 100, 43), (20, 2, 64, 64), (800, 100, 256, 29), (400, 100, 256, 43)):
    head = SequenceWiseClassifier(H, V).cuda().train()
    x = torch.randn(T, B, H, device="cuda").requires_grad_(True)
    out = head(x)
    out.backward(torch.randn_like(out))
    torch.cuda.synchronize()
    head.eval()
    with torch.no_grad():
        p = head(x)
    torch.cuda.synchronize()
    print(f"head T={T} B={B} H={H} V={V} ok", float(out.sum()), float(x.grad.abs().sum()), float(p.sum()) / (T * B))

"""One forward + backward of the fused classifier head at T=750, B=256, H=800, V=29 (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aes_lac_2018_b200 import SequenceWiseClassifier
T, B, H, V = 750, int(sys.argv[1]) if len(sys.argv) > 1 else 256, 800, int(sys.argv[2]) if len(sys.argv) > 2 else 29
x = torch.randn(T, B, H, device="cuda").requires_grad_(True)
head = SequenceWiseClassifier(H, V).cuda().train()
for _ in range(2):
    out = head(x)
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()
print("ok", float(out.sum()))

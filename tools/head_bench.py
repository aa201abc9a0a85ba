"""Classifier head (BatchNorm1d + Linear(H -> V)) fused kernels vs torch's own modules on the same GPU.
Reports ms per forward / forward+backward and the HBM traffic implied by the passes over x."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aes_lac_2018_b200 import SequenceWiseClassifier

def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = []
for (T, B, H, V) in ((750, 32, 800, 29), (750, 256, 800, 29), (750, 64, 800, 43)):
    x = torch.randn(T, B, H, device="cuda").requires_grad_(True)
    dl = torch.randn(B, T, V, device="cuda")
    mine = SequenceWiseClassifier(H, V).cuda().train()
    ref = torch.nn.Sequential(torch.nn.BatchNorm1d(H), torch.nn.Linear(H, V, bias=False)).cuda().train()
    def ref_fwd():
        return ref(x.view(T * B, H)).view(T, B, V).transpose(0, 1)
    def run(f, bwd):
        def g():
            out = f()
            if bwd:
                out.backward(dl)
                x.grad = None
        return g
    row = dict(T=T, B=B, H=H, V=V, x_MB=x.numel() * 4 / 1e6)
    row["fused_fwd_ms"] = timeit(run(lambda: mine(x), False))
    row["torch_fwd_ms"] = timeit(run(ref_fwd, False))
    row["fused_fwd_bwd_ms"] = timeit(run(lambda: mine(x), True))
    row["torch_fwd_bwd_ms"] = timeit(run(ref_fwd, True))
    # fused: x read 2x fwd + 2x bwd, dx written once
    row["fused_fwd_GBps_on_x"] = 2 * row["x_MB"] / row["fused_fwd_ms"]
    row["fused_fwd_bwd_GBps_on_x"] = 5 * row["x_MB"] / row["fused_fwd_bwd_ms"]
    mine.eval(); ref.eval()
    with torch.no_grad():
        row["fused_eval_ms"] = timeit(lambda: mine(x))
        row["torch_eval_ms"] = timeit(lambda: torch.softmax(ref_fwd(), -1))
    res.append(row)
    print(json.dumps(row), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/head_bench.json", "w"), indent=1)

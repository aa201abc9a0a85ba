"""Times the engine on every BASELINE.json config shape (device-resident inputs, CUDA events inside the
library around the kernel launches) and writes profiles-ready JSON.  Run on a B200: python tools/configs_report.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aes_lac_2018_b200 import ctc_loss_raw

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0


def problem(B, T, V, lmin, lmax, tmin=None, seed=1):
    g = torch.Generator().manual_seed(seed)
    acts = torch.randn(T, B, V, generator=g)
    al = torch.full((B,), T, dtype=torch.int32) if tmin is None else torch.randint(tmin, T + 1, (B,), generator=g, dtype=torch.int32)
    if tmin is not None:
        al[0] = T
    ll = torch.minimum(torch.randint(lmin, lmax + 1, (B,), generator=g, dtype=torch.int32), (al // 2).to(torch.int32))
    labels = torch.randint(1, V, (int(ll.sum()),), generator=g, dtype=torch.int32)
    return acts.cuda(), labels, al, ll


def run(name, B, T, V, lmin, lmax, tmin=None, reps=9):
    acts, labels, al, ll = problem(B, T, V, lmin, lmax, tmin)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ks, ws = [], []
    for i in range(reps + 3):
        flush.zero_()
        tm = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        costs, grads, status = ctc_loss_raw(acts, labels, al, ll, timing=tm)
        e1.record(); torch.cuda.synchronize()
        if i >= 3:
            ks.append(tm["kernel_ms"]); ws.append(e0.elapsed_time(e1))
    alg = int(4 * V * int(al.sum()) + 4 * V * T * B + 4 * int(ll.sum()) + 12 * B)
    k = sorted(ks)[len(ks) // 2]; w = sorted(ws)[len(ws) // 2]
    row = dict(config=name, B=B, T=T, V=V, label_len=[lmin, lmax], ragged_T=tmin is not None, kernel_ms=round(k, 4), call_ms=round(w, 4),
               utt_per_s=round(B / (w * 1e-3)), algorithmic_bytes=alg, achieved_GBs=round(alg / (k * 1e-3) / 1e9, 1),
               hbm_frac=round(alg / (k * 1e-3) / 1e9 / PEAK, 4), status_bits=sorted(set(status.tolist())),
               loss=float(costs.double().sum()), grad_rowsum_max=float(grads.sum(-1).abs().max()))
    print(json.dumps(row), flush=True)
    return row


rows = [
    run("configs[0] CPU-runnable call", 4, 200, 29, 10, 50),
    run("configs[1] LibriSpeech backbone", 32, 750, 29, 50, 200),
    run("configs[2] PT-BR fine-tune (V=43, ragged T)", 64, 800, 43, 25, 200, tmin=720),
    run("configs[2] PT-BR fine-tune (V=43, short bucket)", 64, 200, 43, 10, 50, tmin=180),
    run("configs[3] large batch", 1024, 1500, 29, 50, 200),
    run("configs[3] large batch, per-GPU shard of 8", 128, 1500, 29, 50, 200),
    run("configs[4] long utterance T=3000 L=600", 2, 3000, 29, 600, 600),
    run("headline shape at B=8192", 8192, 750, 29, 50, 200),
    run("headline shape, V=43, B=8192", 8192, 750, 43, 50, 200),
]
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/configs_report.json", "w"), indent=1)

"""Development check of the warp ladder (csrc/ctc_warp.cuh): parity against the float64 oracle on a few shapes,
then timing against the round-1 ladders.   python tools/warp_dev.py [--no-parity] [--modes warp,throughput8]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from aes_lac_2018_b200 import ctc_loss_raw


def parity(pmodes=("warp",)):
    from oracle import ctc_f64
    from tests.helpers import synth_problem
    cases = {
        "c1_b4_t200": dict(seed=11, T=200, B=4, V=29, lmin=10, lmax=50),
        "c2_b32_t750": dict(seed=12, T=750, B=32, V=29, lmin=50, lmax=200),
        "ragged_b16": dict(seed=13, T=800, B=16, V=29, lmin=25, lmax=200, tmin=300),
        "v43_ragged": dict(seed=13, T=800, B=16, V=43, lmin=25, lmax=200, tmin=720),
        "peaky_b8_t300": dict(seed=14, T=300, B=8, V=29, lmin=20, lmax=80, peaky=True),
        "sigma4_v43": dict(seed=15, T=400, B=6, V=43, lmin=30, lmax=120, sigma=4.0),
        "short_labels": dict(seed=16, T=64, B=40, V=29, lmin=0, lmax=12, tmin=20),
        "v32": dict(seed=23, T=100, B=5, V=32, lmin=5, lmax=40),
        "v31": dict(seed=24, T=100, B=5, V=31, lmin=5, lmax=40),
        "v2": dict(seed=20, T=40, B=3, V=2, lmin=0, lmax=12),
        "l255": dict(seed=25, T=700, B=3, V=29, lmin=250, lmax=255),
        "tight": dict(seed=26, T=260, B=4, V=29, lmin=120, lmax=127),
    }
    ok = True
    for name, kw in cases.items():
        acts, labels, al, ll = synth_problem(**kw)
        oc, og = ctc_f64.ctc_batch(acts, labels, al, ll)
        for mode in pmodes:
            try:
                c, g, st = ctc_loss_raw(torch.tensor(acts).cuda(), torch.tensor(labels), torch.tensor(al), torch.tensor(ll),
                                        mode=mode)
            except RuntimeError as e:
                print(f"{name:16s} {mode}: ERROR {e}")
                ok = False
                continue
            c = c.numpy().astype(np.float64)
            g = g.cpu().numpy().astype(np.float64)
            rel = np.abs(c - oc) / np.maximum(1.0, np.abs(oc))
            d = np.abs(g - og)
            good = rel.max() <= 1e-4 and d.max() <= 1e-5 and not (st.numpy() & 0x18).any()
            ok &= bool(good)
            print(f"{name:16s} {mode}: loss rel {rel.max():.2e}  grad {d.max():.2e}  status {sorted(set(st.numpy().tolist()))}"
                  f"  {'ok' if good else 'FAIL per-utt ' + str(d.max(axis=(0, 2)))}", flush=True)
    # blank != 0 and forward only
    acts, labels, al, ll = synth_problem(seed=31, T=120, B=6, V=29, lmin=5, lmax=40, blank=5)
    oc, og = ctc_f64.ctc_batch(acts, labels, al, ll, blank=5)
    c, g, st = ctc_loss_raw(torch.tensor(acts).cuda(), torch.tensor(labels), torch.tensor(al), torch.tensor(ll), blank=5, mode=pmodes[0])
    d = np.abs(g.cpu().numpy() - og).max()
    print(f"blank5: loss rel {np.abs(c.numpy() - oc).max():.2e} grad {d:.2e}")
    ok &= d <= 1e-5
    c2, _, _ = ctc_loss_raw(torch.tensor(acts).cuda(), torch.tensor(labels), torch.tensor(al), torch.tensor(ll), blank=5, mode=pmodes[0],
                            want_grad=False)
    print("forward-only equal:", bool((c2 == c).all()))
    return ok


def timing(modes):
    def run(B, lmin, lmax, mode, T=750, V=29):
        g = torch.Generator().manual_seed(1234)
        acts = torch.randn(T, B, V, generator=g).cuda()
        ll = torch.randint(lmin, lmax + 1, (B,), generator=g, dtype=torch.int32)
        al = torch.full((B,), T, dtype=torch.int32)
        labels = torch.randint(1, V, (int(ll.sum()),), generator=g, dtype=torch.int32)
        for _ in range(3):
            ctc_loss_raw(acts, labels, al, ll, mode=mode)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            tm = {}
            c, g_, st = ctc_loss_raw(acts, labels, al, ll, mode=mode, timing=tm)
            best = min(best, tm["kernel_ms"])
        return best, float(c.sum()), int(st.max())
    for (B, lo, hi) in ((8192, 50, 200), (8192, 60, 60), (8192, 120, 120), (8192, 180, 180), (8192, 250, 250), (2048, 50, 200),
                        (1024, 50, 200)):
        row = []
        for mode in modes:
            ms, loss, st = run(B, lo, hi, mode)
            row.append(f"{mode}: {ms:.3f} ms {B / ms / 1e3:.2f} M utt/s (loss {loss:.1f} st {st})")
        print(f"B={B} L{lo}-{hi}: " + " | ".join(row), flush=True)


if __name__ == "__main__":
    modes = ["warp", "throughput8"]
    pmodes = ["warp"]
    for a in sys.argv[1:]:
        if a.startswith("--modes"):
            modes = a.split("=")[1].split(",")
        if a.startswith("--pmodes"):
            pmodes = a.split("=")[1].split(",")
    t0 = time.time()
    if "--no-parity" not in sys.argv:
        print("parity ok" if parity(pmodes) else "PARITY FAILED", f"({time.time() - t0:.0f} s)")
    timing(modes)

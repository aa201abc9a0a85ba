"""Developer timing sweep (not the contract bench): device time of one engine call per configuration,
plus in-kernel cycle counts (clock64 / globaltimer) per utterance.
Usage (on the GPU box): python tools/sweep.py [--quick]"""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from aes_lac_2018_b200 import ctc_loss_raw


class ClockSampler:
    def __init__(self, period=0.05):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.period, self.samples, self.stop = period, [], False
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join()

    def median(self):
        s = sorted(self.samples)
        return s[len(s) // 2] if s else None


def warm_gpu(seconds=1.0):
    a = torch.randn(4096, 4096, device="cuda")
    t = time.time()
    while time.time() - t < seconds:
        for _ in range(10):
            a @ a
        torch.cuda.synchronize()


def problem(B, T, V, lmin, lmax, seed=1234):
    g = torch.Generator().manual_seed(seed)
    acts = torch.randn(T, B, V, generator=g)
    ll = torch.randint(lmin, lmax + 1, (B,), generator=g, dtype=torch.int32)
    al = torch.full((B,), T, dtype=torch.int32)
    labels = torch.randint(1, V, (int(ll.sum()),), generator=g, dtype=torch.int32)
    return acts.cuda(), labels, al, ll


def time_call(fn, iters=7, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    quick = "--quick" in sys.argv
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    cfgs = [
        ("c1", 4, 200, 29, 10, 50), ("c2", 32, 750, 29, 50, 200), ("b128", 128, 750, 29, 50, 200), ("b256", 256, 750, 29, 50, 200),
        ("b1024", 1024, 750, 29, 50, 200), ("b4096", 4096, 750, 29, 50, 200), ("b8192", 8192, 750, 29, 50, 200),
        ("c4", 1024, 1500, 29, 50, 200), ("L200", 2048, 750, 29, 200, 200), ("L60", 2048, 750, 29, 60, 60),
    ]
    if quick:
        cfgs = cfgs[:4]
    warm_gpu(1.5)
    for name, B, T, V, lmin, lmax in cfgs:
        acts, labels, al, ll = problem(B, T, V, lmin, lmax)
        dbg = torch.zeros(B, 16, dtype=torch.int64, device="cuda")
        for mode in ("throughput", "throughput8", "latency", "latency3"):
            if mode == "latency" and B > 4096:
                continue
            for want_grad in ((True, False) if mode not in ("throughput8", "latency3") else (True,)):
                try:
                    warm_gpu(0.2)
                    with ClockSampler() as cs:
                        kw = dict(mode="latency", bidirectional=False) if mode == "latency3" else dict(mode=mode)
                        med, best = time_call(lambda: ctc_loss_raw(acts, labels, al, ll, want_grad=want_grad, **kw), flush=flush)
                    ctc_loss_raw(acts, labels, al, ll, want_grad=want_grad, debug=dbg, **kw)
                    torch.cuda.synchronize()
                    d = dbg.cpu().double()
                except Exception as e:  # noqa: BLE001
                    print(name, mode, want_grad, "FAILED", e)
                    continue
                alg = (8 if want_grad else 4) * T * V * B + 4 * int(ll.sum()) + 12 * B
                row = dict(cfg=name, B=B, T=T, V=V, mode=mode, grad=want_grad, ms_med=round(med, 4), ms_best=round(best, 4),
                           utt_per_s=round(B / med * 1e3), alg_GBs=round(alg / med / 1e6, 1), sm_mhz=cs.median(),
                           fwd_cyc_per_step=round(float(d[:, 0].median()) / T, 1),
                           tot_cyc_per_step=round(float(d[:, 1].median()) / T, 1),
                           cta_us_med=round(float(d[:, 2].median()) / 1e3, 1), cta_us_max=round(float(d[:, 2].max()) / 1e3, 1),
                           eff_mhz=round(float((d[:, 1] / d[:, 2].clamp(min=1)).median()) * 1e3),
                           phases=[round(float(d[:, 4 + i].median()) / T, 1) for i in range(12)])
                rows.append(row)
                print(json.dumps(row), flush=True)
        del acts
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/sweep.json", "w"), indent=1)


if __name__ == "__main__":
    main()

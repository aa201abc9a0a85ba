"""B=32 (BASELINE configs[1]) call latency of the default and of alternative builds.  python tools/latency_probe.py NAME..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from aes_lac_2018_b200 import ctc_loss_raw
g = torch.Generator().manual_seed(99)
B = 32
acts = torch.randn(750, B, 29, generator=g).cuda()
ll = torch.randint(50, 201, (B,), generator=g, dtype=torch.int32)
al = torch.full((B,), 750, dtype=torch.int32)
labels = torch.randint(1, 29, (int(ll.sum()),), generator=g, dtype=torch.int32)
res = []
for bidir in (True, False):
    tm = {}
    for _ in range(10): ctc_loss_raw(acts, labels, al, ll, bidirectional=bidir, timing=tm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best, kbest = 1e9, 1e9
    for _ in range(5):
        e0.record()
        for _ in range(20):
            ctc_loss_raw(acts, labels, al, ll, bidirectional=bidir, timing=tm); kbest = min(kbest, tm["kernel_ms"])
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    res.append("%%s: call %%.4f ms kernels %%.4f ms" %% ("bidirectional" if bidir else "three-sweep", best, kbest))
print(" | ".join(res))
''' % ROOT
for name in sys.argv[1:] or [""]:
    env = dict(os.environ)
    if name:
        env["CTC_B200_LIB"] = os.path.join(ROOT, "aes_lac_2018_b200", "lib", f"libctc_b200_{name}.so")
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"{name or 'default':8s}", r.stdout.strip()[-300:], r.stderr.strip()[-300:] if r.returncode else "", flush=True)

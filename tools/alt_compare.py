"""Times alternative builds of the library (tools/build_alt.sh) on fixed-L workloads and the bench mix.
python tools/alt_compare.py NAME [NAME ...]   ('' = the default build)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from aes_lac_2018_b200 import ctc_loss_raw
def run(B, lmin, lmax, mode):
    g = torch.Generator().manual_seed(1234)
    acts = torch.randn(750, B, 29, generator=g).cuda()
    ll = torch.randint(lmin, lmax + 1, (B,), generator=g, dtype=torch.int32)
    al = torch.full((B,), 750, dtype=torch.int32)
    labels = torch.randint(1, 29, (int(ll.sum()),), generator=g, dtype=torch.int32)
    for _ in range(3): ctc_loss_raw(acts, labels, al, ll, mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(4): c, g_, st = ctc_loss_raw(acts, labels, al, ll, mode=mode)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 4)
    return best, float(c.sum())
out = []
for (B, lo, hi) in ((4096, 110, 110), (4096, 150, 150), (4096, 180, 180), (4096, 200, 200), (8192, 50, 200)):
    ms, loss = run(B, lo, hi, "throughput8")
    out.append("L%%d-%%d: %%.3f ms (%%.2f M utt/s)" %% (lo, hi, ms, B / ms / 1e3))
print(" | ".join(out), " loss", loss)
''' % ROOT
for name in sys.argv[1:] or [""]:
    env = dict(os.environ)
    if name:
        env["CTC_B200_LIB"] = os.path.join(ROOT, "aes_lac_2018_b200", "lib", f"libctc_b200_{name}.so")
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"{name or 'default':8s}", r.stdout.strip()[-400:], r.stderr.strip()[-300:] if r.returncode else "", flush=True)

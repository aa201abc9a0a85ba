"""Times alternative builds (tools/build_warp_alt.sh) of the warp ladder on fixed-L workloads.
python tools/warp_alt_time.py "L,L,..." NAME [NAME ...]      ('default' = the default build; L = 'lo-hi' or a single length)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from aes_lac_2018_b200 import ctc_loss_raw
def run(B, lmin, lmax, mode, T=750, V=29):
    g = torch.Generator().manual_seed(1234)
    acts = torch.randn(T, B, V, generator=g).cuda()
    ll = torch.randint(lmin, lmax + 1, (B,), generator=g, dtype=torch.int32)
    al = torch.full((B,), T, dtype=torch.int32)
    labels = torch.randint(1, V, (int(ll.sum()),), generator=g, dtype=torch.int32)
    for _ in range(3): ctc_loss_raw(acts, labels, al, ll, mode=mode)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        tm = {}
        c, g_, st = ctc_loss_raw(acts, labels, al, ll, mode=mode, timing=tm)
        best = min(best, tm["kernel_ms"])
    return best, float(c.sum()), int(st.max())
out = []
for spec in sys.argv[1].split(","):
    lo, hi = (spec.split("-") + [spec])[:2] if "-" in spec else (spec, spec)
    try:
        ms, loss, st = run(8192, int(lo), int(hi), os.environ.get("ALT_MODE", "warp"), V=int(os.environ.get("ALT_V", "29")))
        out.append("L%%s: %%.3f ms %%.2f M/s st%%d" %% (spec, ms, 8192 / ms / 1e3, st))
    except Exception as e:
        out.append("L%%s: ERR %%s" %% (spec, str(e)[:60]))
print(" | ".join(out))
''' % ROOT
specs = sys.argv[1]
for name in sys.argv[2:]:
    env = dict(os.environ)
    if name != "default":
        env["CTC_B200_LIB"] = os.path.join(ROOT, "aes_lac_2018_b200", "lib", f"libctc_b200_{name}.so")
    r = subprocess.run([sys.executable, "-c", CHILD, specs], env=env, capture_output=True, text=True)
    print(f"{name:12s}", r.stdout.strip()[-600:], r.stderr.strip()[-300:] if r.returncode else "", flush=True)

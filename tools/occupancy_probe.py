"""How does the warp ladder scale with resident warps per SM?  Pads the dynamic shared memory of the launch
(CTC_B200_WARP_PAD_KB) to cap the CTAs per SM and times one fixed-L workload.   python tools/occupancy_probe.py [L]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = sys.argv[1] if len(sys.argv) > 1 else "120"
for pad in (0, 8, 12, 16, 22, 30, 50, 100):
    env = dict(os.environ, CTC_B200_WARP_PAD_KB=str(pad))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "warp_alt_time.py"), L, "default"], env=env, capture_output=True, text=True)
    print(f"pad {pad:3d} KB (<= {227 // max(1, pad + 6)} CTAs/SM by shared memory):", r.stdout.strip()[-60:], flush=True)

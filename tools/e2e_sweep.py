"""e2e (host buffers) throughput of ctc_b200_compute_host vs number of pipeline chunks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aes_lac_2018_b200 import ctc_loss_host
from bench import make_problem, T, V
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
acts, labels, al, ll = make_problem(B, 1234)
pinned = acts.pin_memory(); grads = torch.empty((T, B, V), dtype=torch.float32, pin_memory=True)
for nch in (8, 16, 21, 26, 32, 37, 43, 52, 64, 86, 128):
    for _ in range(2): ctc_loss_host(pinned, labels, al, ll, grads_out=grads, n_chunks=nch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): ctc_loss_host(pinned, labels, al, ll, grads_out=grads, n_chunks=nch)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 4
    print(f"n_chunks {nch:3d} (slices of {-(-(-(-B // nch)) // 32) * 32}): {dt*1e3:7.2f} ms/step  {B/dt:10.0f} utt/s   H2D+D2H {(2*acts.numel()*4)/dt/1e9:6.1f} GB/s")
# plain copies for reference
d = torch.empty_like(acts, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(pinned, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
grads.copy_(d, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"plain H2D {acts.numel()*4/(t1-t0)/1e9:.1f} GB/s, D2H {acts.numel()*4/(t2-t1)/1e9:.1f} GB/s")

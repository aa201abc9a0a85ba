"""Small problems through every kernel path, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from aes_lac_2018_b200 import ctc_loss_raw, greedy_decode_raw, ctc_loss_host
from tests.helpers import synth_problem

def run(tag, acts, labels, al, ll, **kw):
    c, g, st = ctc_loss_raw(torch.tensor(acts).cuda(), torch.tensor(labels), torch.tensor(al), torch.tensor(ll), **kw)
    torch.cuda.synchronize()
    print(tag, "ok", float(c.sum()), sorted(set(st.tolist())))

small = synth_problem(1, 50, 5, 29, 0, 20, tmin=30)
mid = synth_problem(2, 70, 3, 43, 40, 150 // 2, tmin=60)
wide = synth_problem(3, 300, 1, 29, 140, 140)
for name, prob in (("small", small), ("mid", mid), ("wide", wide)):
    for mode, bidir in (("throughput", True), ("throughput8", True), ("latency", True), ("latency", False)):
        run(f"{name}/{mode}/bidir={bidir}", *prob, mode=mode, bidirectional=bidir)
    run(f"{name}/costs-only", *prob, want_grad=False)
hostile = synth_problem(4, 120, 3, 29, 30, 60, sigma=40.0)
run("hostile/auto", *hostile)
run("hostile/throughput", *hostile, mode="throughput8")
a, l, al, ll = small
c, g, st = ctc_loss_host(torch.tensor(a).pin_memory(), torch.tensor(l), torch.tensor(al), torch.tensor(ll), n_chunks=2)
print("host ok", float(c.sum()))
p = torch.randn(4, 77, 29).cuda()
tok, off, cnt = greedy_decode_raw(p, torch.tensor([77, 5, 0, 40], dtype=torch.int32))
torch.cuda.synchronize()
print("decode ok", cnt.tolist())

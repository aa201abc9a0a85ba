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
    for mode, bidir in (("warp32", False), ("warp", False), ("throughput", True), ("throughput8", True), ("latency", True), ("latency", False)):
        run(f"{name}/{mode}/bidir={bidir}", *prob, mode=mode, bidirectional=bidir)
    run(f"{name}/costs-only", *prob, want_grad=False)
hostile = synth_problem(4, 120, 3, 29, 30, 60, sigma=40.0)
run("hostile/auto", *hostile)
run("hostile/throughput", *hostile, mode="throughput8")
run("hostile/warp (device-side log-space detour)", *hostile, mode="warp")
run("hostile/warp32 (fp64 tier, then the log-space detour)", *hostile, mode="warp32")
wide4 = synth_problem(15, 400, 6, 43, 30, 120, sigma=4.0)
run("sigma4/warp32 (fp64 tier)", *wide4, mode="warp32")
# round 2: every warp-ladder variant (NS = 2 .. 16, one and two alphabet slices), partial last chunks, blank != 0
for L, V in ((10, 29), (40, 29), (70, 43), (100, 29), (130, 29), (170, 43), (200, 29), (250, 31)):
    prob = synth_problem(40 + L, 2 * L + 37, 2, V, L, L, blank=3)
    run(f"warp/L={L}/V={V}", *prob, mode="warp", blank=3)
    run(f"warp32/L={L}/V={V}", *prob, mode="warp32", blank=3)
# several label classes in one call: per-bucket SM ranges, dynamically claimed workspace slots
mixed = synth_problem(77, 160, 64, 29, 5, 75, tmin=100)
run("warp32/mixed classes", *mixed, mode="warp32")
run("warp/mixed classes", *mixed, mode="warp")
# round 2: non-blocking call, device-side cost sum, gradient scale, fused loss glue
from aes_lac_2018_b200 import CTCLoss, sanitize_loss
from aes_lac_2018_b200.ctc_loss import reduce_costs, _scale_gradients
a, l, al, ll = small
c_d, g_d, s_d = ctc_loss_raw(torch.tensor(a).cuda(), torch.tensor(l), torch.tensor(al), torch.tensor(ll), mode="warp", no_sync=True)
loss, flag = reduce_costs(c_d, 0.5, True)
_scale_gradients(g_d, 0.25, loss, flag)
torch.cuda.synchronize()
print("glue ok", float(loss), int(flag))
out = torch.tensor(a).cuda().transpose(0, 1).contiguous().requires_grad_()
ls = sanitize_loss(CTCLoss(), out, torch.tensor(l), torch.tensor(al, dtype=torch.float32) / a.shape[0], torch.tensor(ll), average=5, weight=0.7)
ls.backward()
torch.cuda.synchronize()
print("sanitize_loss ok", float(ls))
a, l, al, ll = small
c, g, st = ctc_loss_host(torch.tensor(a).pin_memory(), torch.tensor(l), torch.tensor(al), torch.tensor(ll), n_chunks=2)
print("host ok", float(c.sum()))
p = torch.randn(4, 77, 29).cuda()
tok, off, cnt = greedy_decode_raw(p, torch.tensor([77, 5, 0, 40], dtype=torch.int32))
torch.cuda.synchronize()
print("decode ok", cnt.tolist())
# the W = 8 / K = 4 variant (more softmax row slots than rows) and a tight alignment (state pruning at chunk edges)
widest = synth_problem(5, 1250, 1, 29, 1100, 1100)
run("widest/throughput", *widest, mode="throughput", bidirectional=False)
run("widest/latency", *widest, mode="latency")
# edit distance: all three modes, ragged lengths, empty sides
from aes_lac_2018_b200 import edit_distance_raw, SequenceWiseClassifier
rng = np.random.default_rng(0)
hyp = torch.tensor(rng.integers(1, 29, (5, 300)).astype(np.int32)).cuda()
cnt = torch.tensor([300, 0, 17, 150, 299], dtype=torch.int32).cuda()
ref_lens = torch.tensor([200, 5, 0, 130, 127])
refs = torch.tensor(rng.integers(1, 29, int(ref_lens.sum())).astype(np.int32))
for mode in ("tokens", "cer", "wer"):
    d, n = edit_distance_raw(hyp, cnt, refs, ref_lens, space=28, mode=mode)
    torch.cuda.synchronize()
    print("edit", mode, "ok", d.tolist(), n.tolist())
# classifier head: ragged row count, both class paddings, training and eval, backward
for (T, B, H, V) in ((37, 3, 100, 29), (20, 2, 64, 43)):
    head = SequenceWiseClassifier(H, V).cuda()
    x = torch.randn(T, B, H, device="cuda", requires_grad=True)
    out = head(x)
    out.backward(torch.randn_like(out))
    head.eval()
    with torch.no_grad():
        p = head(x)
    torch.cuda.synchronize()
    print("head ok", tuple(out.shape), float(p.sum()))

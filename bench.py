#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the CTC loss-and-gradient hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B_per_gpu] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): CTC fwd+bwd utterances/s at T=750, V=29.  A "step" is one forward+backward pass of
the hot path over one batch of synthetic DeepSpeech2-shaped activations (BASELINE configs[1] shape:
T=750, V=29, fp32, L ~ U{50..200}), `--batch` utterances per GPU (weak scaling: every rank owns its own
shard; the only collective is the scalar NCCL loss sum).

One JSON line on rank 0:
  value     utterances/s with the activations already resident in HBM (CUDA events, max over ranks)
  e2e       utterances/s through the host-buffer C-ABI entry point (pinned host activations in, host gradients
            out, copies inside the timed region)
  roofline  algorithmic bytes of the fused kernel launches / their device time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the fp32 C/OpenMP restatement of warp-ctc's CPU path (oracle/) on this box's host cores
`--impl reference` times that CPU restatement alone (warp-ctc itself is an absent third-party dependency of
the reference), on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T, V, LMIN, LMAX = 750, 29, 50, 200
METRIC = "CTC fwd+bwd utterances/s (T=750,V=29)"
FALLBACK_HBM_GBS = 6650.0        # /opt/skills/guides/B200_PROFILING.md fallback


def make_problem(batch: int, seed: int):
    """Deterministic synthetic problem: N(0,1) logits, uniform labels in [1, V-1]."""
    import torch
    g = torch.Generator().manual_seed(seed)
    acts = torch.randn(T, batch, V, generator=g, dtype=torch.float32)
    label_lens = torch.randint(LMIN, LMAX + 1, (batch,), generator=g, dtype=torch.int32)
    act_lens = torch.full((batch,), T, dtype=torch.int32)
    labels = torch.randint(1, V, (int(label_lens.sum()),), generator=g, dtype=torch.int32)
    return acts, labels, act_lens, label_lens


def algorithmic_bytes(batch: int, label_lens) -> int:
    """read acts once + write grads once + labels + two lengths + cost, per utterance."""
    return 8 * T * V * batch + 4 * int(label_lens.sum()) + 12 * batch


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML, ~50 ms period)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int = 0, period: float = 0.01):
        self.samples, self.reasons, self.stop, self.max_mhz = [], set(), False, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None
        self.period = period
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop and self.nv is not None:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def host_threads() -> int:
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU arm overrides it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def cpu_reference_run(sample_batch: int, min_seconds: float, max_reps: int, threads: int = 0):
    """Times the fp32 C/OpenMP restatement of warp-ctc's CPU path (oracle/warpctc_cpu.c) on host cores."""
    from oracle import warpctc_cpu
    threads = threads if threads > 0 else host_threads()
    acts, labels, act_lens, label_lens = make_problem(sample_batch, seed=4321)
    a, lab, al, ll = acts.numpy(), labels.numpy(), act_lens.numpy(), label_lens.numpy()
    warpctc_cpu.ctc_batch(a[:, :8], lab[:int(ll[:8].sum())], al[:8], ll[:8], num_threads=threads)   # warm-up
    reps, t0 = 0, time.perf_counter()
    times = []
    while reps < max_reps and (time.perf_counter() - t0 < min_seconds or reps < 2):
        t1 = time.perf_counter()
        costs, _ = warpctc_cpu.ctc_batch(a, lab, al, ll, num_threads=threads)
        times.append(time.perf_counter() - t1)
        reps += 1
    best = min(times)
    return {"utt_per_s": sample_batch / best, "reps": reps, "best_s": best, "mean_s": sum(times) / len(times),
            "cores": warpctc_cpu.max_threads() if threads <= 0 else threads, "loss": float(costs.sum())}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    sample = args.cpu_sample
    steps, warm = args.steps, args.warmup
    from oracle import warpctc_cpu
    acts, labels, act_lens, label_lens = make_problem(sample, seed=4321)
    a, lab, al, ll = acts.numpy(), labels.numpy(), act_lens.numpy(), label_lens.numpy()
    cores = host_threads()
    for _ in range(warm):
        warpctc_cpu.ctc_batch(a, lab, al, ll, num_threads=cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        warpctc_cpu.ctc_batch(a, lab, al, ll, num_threads=cores)
    dt = time.perf_counter() - t0
    val = sample * steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "utterances/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1] shape (T={T}, V={V}, L~U{{{LMIN}..{LMAX}}}, fp32 activations, "
                               f"randn logits) at throughput batch {args.batch} utterances per GPU",
                   "batch_per_gpu": args.batch, "T": T, "V": V, "label_len": [LMIN, LMAX],
                   "reference_sample": f"each step is a bounded sample of {sample} utterances of that workload on the host CPU"},
        "cpu_baseline": {"value": val, "unit": "utterances/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} utterances x {steps} steps, fp32 C/OpenMP restatement of warp-ctc's CPU path "
                                   "(oracle/warpctc_cpu.c); warp-ctc itself is not installable offline"},
        "e2e": {"value": val, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def measure_link(dev, nbytes: int = 256 << 20):
    """Pinned host <-> device copy rates on this box (GB/s): H2D alone, D2H alone, both directions at once."""
    import torch
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize(dev)
            best = min(best, time.perf_counter() - t0)
        return nbytes / best / 1e9

    def up():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def down():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        up()
        down()

    up(); down()
    return {"h2d": timed(up), "d2h": timed(down), "duplex_per_direction": timed(both)}


def make_peaky_ragged(batch: int, seed: int):
    """Trained-model-like posteriors (SURVEY.md 8d "peaky" variant): N(0,1) logits, +6 on the blank for 70 % of the
    frames and +6 on an aligned label for the rest, ragged act_lens in [0.6 T, T] with L <= T/4."""
    import torch
    g = torch.Generator().manual_seed(seed)
    acts = torch.randn(T, batch, V, generator=g, dtype=torch.float32)
    act_lens = torch.randint(int(0.6 * T), T + 1, (batch,), generator=g, dtype=torch.int32)
    label_lens = torch.minimum(torch.randint(LMIN, LMAX + 1, (batch,), generator=g, dtype=torch.int32), act_lens // 4)
    labels = torch.randint(1, V, (int(label_lens.sum()),), generator=g, dtype=torch.int32)
    blank_frames = torch.rand(T, batch, generator=g) < 0.7
    acts[..., 0] += 6.0 * blank_frames
    offs = torch.cumsum(label_lens.long(), 0) - label_lens.long()
    # frame t of utterance b is "aligned" to label floor(t * L / T_b)
    t_idx = torch.arange(T).unsqueeze(1)
    pos = torch.clamp((t_idx * label_lens.unsqueeze(0).long()) // act_lens.unsqueeze(0).long().clamp(min=1),
                      max=(label_lens.long() - 1).clamp(min=0).unsqueeze(0))
    sym = labels.long()[(offs.unsqueeze(0) + pos).clamp(max=max(labels.numel() - 1, 0))]
    bump = (~blank_frames) & (label_lens.unsqueeze(0) > 0)
    acts.scatter_add_(2, sym.unsqueeze(2), (6.0 * bump).unsqueeze(2).to(acts.dtype))
    return acts, labels, act_lens, label_lens


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=8192, help="utterances per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=256, help="utterances per CPU-reference step")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline numbers only (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from aes_lac_2018_b200 import _lib, build as _build, ctc_loss_host, ctc_loss_raw
    from aes_lac_2018_b200.distributed import shard_bounds, shard_problem, sharded_loss_step
    if not os.path.exists(_build.LIB):
        # harness convenience only (never taken when the in-tree .so travelled with the snapshot); the engine itself
        # never builds or falls back -- it fails loudly.  Local rank 0 builds, the other ranks wait for the file.
        if local_rank == 0:
            _build.build()
        for _ in range(1800):
            if os.path.exists(_build.LIB):
                break
            time.sleep(1.0)

    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    B = args.batch
    acts_h, labels, act_lens, label_lens = make_problem(B, seed=1234 + rank)
    acts = acts_h.to(dev)
    # targets and lengths stay on the host (the reference's interface, engine.py:16) but in PINNED memory, as a
    # DataLoader(pin_memory=True) delivers them: the 4 MB label upload is then one DMA instead of a staged copy
    labels, act_lens, label_lens = labels.pin_memory(), act_lens.pin_memory(), label_lens.pin_memory()
    alg_bytes = algorithmic_bytes(B, label_lens)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    last = {}

    def step():
        # The public non-blocking path: engine (NO_SYNC) -> device-side cost sum -> scalar NCCL all-reduce, all
        # stream-ordered; nothing is read back inside the timed region except once, after its last step.
        # With more than one rank the collective runs beside the next step's kernels (overlap=True): nothing in a step
        # needs the previous step's global loss; every pending collective is waited for before the timed region ends.
        total, local, grads, status = sharded_loss_step(acts, labels, act_lens, label_lens, want_grad=True, overlap=True)
        pending.append(total)
        last.update(local=local, grads=grads, status=status)

    pending = []

    def drain():
        for p in pending:
            last["total"] = p.wait()
        pending.clear()

    # ---- device-resident throughput (value) ----
    for _ in range(args.warmup):
        step()
    drain()
    barrier()
    n0 = _lib.launch_count()
    with ClockSampler(local_rank) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        drain()                                            # (the current stream now waits for every collective)
        e1.record()
        barrier()
    ms_max = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count() - n0
    value = B * world * args.steps / (ms_max * 1e-3)

    # ---- what the step computed: the all-reduced loss must equal the fp64 sum of the gathered local sums ----
    loss_total = float(last["total"].item())
    local_sum = float(last["local"].item())
    status_or = 0
    for bit in (1, 2, 4, 8, 16, 32):
        if bool((last["status"] & bit).any().item()):
            status_or |= bit
    gathered = [local_sum]
    if world > 1:
        obj = [None] * world
        dist.all_gather_object(obj, local_sum)
        gathered = [float(x) for x in obj]
    loss_fp64 = float(sum(gathered))
    loss_ok = abs(loss_total - loss_fp64) <= 1e-6 * abs(loss_fp64)
    assert loss_ok, f"all-reduced loss {loss_total} != sum of local sums {loss_fp64}"

    # ---- kernel-only device time of one engine call (roofline) ----
    # CUDA events recorded by the library on the launching stream: first kernel launch of a call -> completion
    # of the last one (the variant launches overlap on forked streams and are joined back before the end event).
    kt = []
    for _ in range(min(args.steps, 10)):
        tm = {}
        ctc_loss_raw(acts, labels, act_lens, label_lens, want_grad=True, timing=tm)
        kt.append(tm["kernel_ms"])
    kernel_ms = sorted(kt)[len(kt) // 2]

    # ---- the blocking drop-in call (what `warpctc_pytorch.CTCLoss` does: costs read back every step) ----
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(min(args.steps, 10)):
        ctc_loss_raw(acts, labels, act_lens, label_lens, want_grad=True)
    e1.record()
    barrier()
    blocking_ms = max_over_ranks(e0.elapsed_time(e1)) / min(args.steps, 10)

    # ---- end to end with host buffers (e2e) ----
    pinned = acts_h.pin_memory()
    grads_h = torch.empty((T, B, V), dtype=torch.float32, pin_memory=True)
    for _ in range(2):
        ctc_loss_host(pinned, labels, act_lens, label_lens, grads_out=grads_h)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.e2e_steps):
        costs_h, _, _ = ctc_loss_host(pinned, labels, act_lens, label_lens, grads_out=grads_h)
        local = costs_h.double().sum().to(torch.float32).reshape(1).to(dev)
        if world > 1:
            dist.all_reduce(local)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), 0.0))
    e2e_value = B * world * args.e2e_steps / (e2e_ms * 1e-3)
    h2d = acts_h.numel() * 4 + labels.numel() * 4 + 2 * B * 4
    d2h = grads_h.numel() * 4 + B * 4 + B * 4
    link = measure_link(dev) if not args.no_extras else None
    del pinned, grads_h

    extras = {}
    if not args.no_extras:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timed_calls(fn, reps=8, skip=3, do_flush=True):
            ts = []
            for i in range(reps):
                if do_flush:
                    flush.zero_()                      # inputs smaller than L2: flush it between calls
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k0.record()
                fn()
                k1.record()
                torch.cuda.synchronize()
                if i >= skip:
                    ts.append(k0.elapsed_time(k1))
            return sorted(ts)[len(ts) // 2]

        # BASELINE configs[1] literally (B=32): the latency regime the reference trains in
        a32_h, l32, al32, ll32 = make_problem(32, seed=99)
        a32 = a32_h.to(dev)
        lat_ms = timed_calls(lambda: ctc_loss_raw(a32, l32, al32, ll32, want_grad=True))
        lat_nosync_ms = timed_calls(lambda: sharded_loss_step(a32, l32, al32, ll32))
        extras["configs1_b32_latency"] = {
            "ms_per_call": lat_ms, "utterances_per_s": 32 / (lat_ms * 1e-3), "ms_per_call_non_blocking": lat_nosync_ms,
            "note": "BASELINE configs[1] literally (B=32, the reference's training batch size): bound by the T-serial chain "
                    "(750 dependent steps), not by bytes; L2 flushed between calls"}

        # BASELINE configs[3]: B=1024, T=1500 split over the N ranks (strong scaling); every rank also times the whole
        # batch alone, so the efficiency t(1) / (N * t(N)) comes from one run on one box
        g = torch.Generator().manual_seed(777)
        B4, T4 = 1024, 1500
        a4 = torch.randn(T4, B4, V, generator=g, dtype=torch.float32)
        ll4 = torch.randint(LMIN, LMAX + 1, (B4,), generator=g, dtype=torch.int32)
        al4 = torch.full((B4,), T4, dtype=torch.int32)
        lab4 = torch.randint(1, V, (int(ll4.sum()),), generator=g, dtype=torch.int32)
        a4d = a4.to(dev)
        t_full = timed_calls(lambda: ctc_loss_raw(a4d, lab4, al4, ll4, want_grad=True, no_sync=True), do_flush=False)
        lo, hi = shard_bounds(B4, world, rank)
        lab_s, al_s, ll_s = shard_problem(lab4, al4, ll4, lo, hi)
        a4s = a4d[:, lo:hi].contiguous()
        for _ in range(3):
            sharded_loss_step(a4s, lab_s, al_s, ll_s)
        barrier()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for _ in range(5):
            sharded_loss_step(a4s, lab_s, al_s, ll_s)
        k1.record()
        barrier()
        t_shard = max_over_ranks(k0.elapsed_time(k1) / 5)
        t_full = max_over_ranks(t_full)
        extras["c4_strong"] = {
            "workload": "BASELINE configs[3]: B=1024, T=1500, V=29, L~U{50..200}, batch split over the ranks, scalar NCCL loss sum",
            "n_gpus": world, "ms_per_step": t_shard, "utterances_per_s": B4 / (t_shard * 1e-3),
            "single_gpu_ms_same_box": t_full, "efficiency_vs_1gpu": t_full / (world * t_shard),
            "note": "1024/N utterances per GPU: below one wave of resident warps for N >= 2, so the step time approaches "
                    "the T-serial chain of the longest transcript (1500 dependent steps)"}
        del a4, a4d, a4s

        # trained-model-like activations, ragged lengths: how often does the log-space detour fire?
        ap_h, lp, alp, llp = make_peaky_ragged(B, seed=555 + rank)
        ap_d = ap_h.to(dev)
        pk_ms = timed_calls(lambda: ctc_loss_raw(ap_d, lp, alp, llp, want_grad=True, no_sync=True), reps=6, do_flush=False)
        _, _, st_p = ctc_loss_raw(ap_d, lp, alp, llp, want_grad=True)
        extras["peaky_ragged"] = {
            "workload": f"B={B}, blank +6 on 70 % of frames / aligned label +6 otherwise, act_lens in [0.6 T, T], L <= T/4",
            "ms_per_call": pk_ms, "utterances_per_s": B / (pk_ms * 1e-3),
            "logspace_detour_rate": float((st_p & 16).ne(0).float().mean().item()),
            "fp64_tier_rate": float((st_p & 32).ne(0).float().mean().item()),
            "infeasible": int((st_p & 1).ne(0).sum().item()), "range_flag_left": int((st_p & 8).ne(0).sum().item())}
        del ap_h, ap_d

        # a neutral GPU arm: torch.nn.functional.ctc_loss (fp32, log_softmax + autograd backward) on the same shape
        try:
            import torch.nn.functional as F
            Bt = 1024
            x = acts[:, :Bt].detach().clone().requires_grad_()
            tl = label_lens[:Bt].long()
            tg = labels[:int(tl.sum())].long().to(dev)
            il = act_lens[:Bt].long()

            def torch_step():
                x.grad = None
                F.ctc_loss(F.log_softmax(x, -1), tg, il, tl, blank=0, reduction="sum").backward()

            tms = timed_calls(torch_step, reps=5, skip=2, do_flush=False)
            extras["torch_ctc_loss_gpu"] = {"batch": Bt, "ms_per_step": tms, "utterances_per_s": Bt / (tms * 1e-3),
                                            "what": "torch.nn.functional.ctc_loss fp32 (log_softmax + backward) on this GPU, "
                                                    "same T, V, L distribution; a neutral GPU implementation, not the reference's"}
        except Exception as e:  # noqa: BLE001
            extras["torch_ctc_loss_gpu"] = {"unavailable": repr(e)[:200]}

        # the widening row next to the hot path: the classifier head that produces the activations
        try:
            from aes_lac_2018_b200 import SequenceWiseClassifier
            Th, Bh, Hh = T, 256, 800
            xh = torch.randn(Th, Bh, Hh, device=dev, requires_grad=True)
            dlh = torch.randn(Th, Bh, V, device=dev)
            head = SequenceWiseClassifier(Hh, V).to(dev).train()

            def head_step():
                xh.grad = None
                head.forward_time_major(xh).backward(dlh)

            hms = timed_calls(head_step, reps=8, skip=3, do_flush=False)
            xbytes = Th * Bh * Hh * 4
            extras["classifier_head"] = {
                "workload": f"BatchNorm1d({Hh}) + Linear({Hh}, {V}) on T={Th}, B={Bh} rows (x = {xbytes / 1e6:.0f} MB), training forward + backward",
                "ms_per_step": hms, "x_passes": "2 reads forward, 2 reads + 1 write backward",
                "hbm_gbs_on_x": 5 * xbytes / (hms * 1e-3) / 1e9,
                "what": "tcgen05 (3xTF32) forward, weight-gradient and input-gradient kernels of csrc/ctc_head*.cu*; "
                        "not part of the headline metric"}
            del xh, dlh, head
        except Exception as e:  # noqa: BLE001
            extras["classifier_head"] = {"unavailable": repr(e)[:200]}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("dram_bytes_per_launch_set")
                traffic_src = "NOT measured in this run: " + tj.get("source", "profiles/r2_traffic.json")
            except Exception:  # noqa: BLE001
                traffic = None
        out = {
            "metric": METRIC, "value": value, "unit": "utterances/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1] shape (T={T}, V={V}, L~U{{{LMIN}..{LMAX}}}, fp32 activations, "
                                   f"randn logits) at throughput batch {B} utterances per GPU",
                       "batch_per_gpu": B, "global_batch": B * world, "T": T, "V": V, "label_len": [LMIN, LMAX],
                       "parallelism": f"batch-sharded x{world}, scalar NCCL loss sum enqueued behind the kernels (no host round trip)",
                       "api": "aes_lac_2018_b200.distributed.sharded_loss_step(overlap=True): ctc_b200_compute(NO_SYNC) + ctc_b200_reduce_costs + asynchronous all_reduce beside the next step",
                       "l2": f"inputs larger than L2 ({acts.numel() * 4 >> 20} MiB activations + equal gradients per GPU)",
                       "host_side": "labels and lengths in pinned host memory (engine.py:16 keeps them on the CPU)"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "utterances/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "ctc_b200_compute_host (pinned host activations in, host gradients + costs out)",
                    "steps": args.e2e_steps, "ms_per_step": e2e_ms / args.e2e_steps, "link_gbs": link,
                    "link_bound_utt_per_s": (B / (max(h2d, d2h) / (link["duplex_per_direction"] * 1e9)) * world) if link else None},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "kernel": "ctc_warp32_kernel<NS,8,1,*> (fp32 recursion with per-lane block exponents; the per-label-class launches of one call run side by side on disjoint SM ranges)",
                         "algorithmic_bytes_per_call": alg_bytes, "kernel_ms_per_call": kernel_ms},
            "blocking_dropin": {"ms_per_step": blocking_ms, "utterances_per_s": B * world / (blocking_ms * 1e-3),
                                "what": "same engine call with upstream's blocking contract (costs and status read back every step)"},
            "loss_check": {"all_reduced": loss_total, "fp64_sum_of_local_sums": loss_fp64, "equal": loss_ok, "status_bits_seen": status_or},
        }
        out.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_run(args.cpu_sample, min_seconds=10.0, max_reps=40)
            out["cpu_baseline"] = {"value": cb["utt_per_s"], "unit": "utterances/s", "cores": cb["cores"], "kind": "port",
                                   "sample": f"{args.cpu_sample} utterances of the same workload x {cb['reps']} repetitions "
                                             f"(best {cb['best_s']:.3f} s), fp32 C/OpenMP restatement of warp-ctc's CPU path "
                                             "(a port: warp-ctc itself is absent; OpenMP schedule(dynamic,1) where upstream "
                                             "uses a plain parallel for, which favours this arm slightly)"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

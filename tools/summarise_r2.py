"""Round-2 profile summaries (committed under profiles/):
  python tools/summarise_r2.py launches gpurun_out/launches_r2.csv        -> profiles/r2_launches_bench.txt, profiles/r2_traffic.json
  python tools/summarise_r2.py kernel gpurun_out/prof_r2_b.ncu-rep NAME B T "what"  -> profiles/r2_ncu_NAME.txt
(ncu reads the .ncu-rep files here, without a GPU.)"""
import collections, csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ii, ui, gi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit", "Grid Size"))
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[ii], {"k": r[ki], "grid": r[gi]})
        v, u = float(r[vi].replace(",", "")), r[ui]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[u] if r[mi] == "gpu__time_duration.sum" else {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        d[r[mi]] = v

    def grp(name):
        m = re.search(r"ctc_warp32_kernel<\(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+)>", name) or re.search(r"ctc_warp32_kernel<(\d+), (\d+), (\d+), (\d+)>", name)
        if m:
            return "ctc_warp32_kernel<%s,%s,%s,%s>" % m.groups(), "engine step (B=8192, fp32 warp ladder)"
        m = re.search(r"ctc_warp_kernel<\(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+)>", name) or re.search(r"ctc_warp_kernel<(\d+), (\d+), (\d+), (\d+), (\d+)>", name)
        if m:
            return "ctc_warp_kernel<%s,%s,%s,%s,%s>" % m.groups(), "fp64 second tier of the same label class (scans the bucket's status words; nothing flagged)"
        m = re.search(r"ctc_fused_kernel<(\d+), (\d+), (\d+), (\d+)>", name)
        if m:
            return "ctc_fused_kernel<%s,%s,%s,%s>" % m.groups(), "e2e pipeline slices (256 utterances each, latency ladder)"
        for key, role in (("ctc_logspace", "device-side log-space detour (scans the status words; nothing flagged)"),
                          ("reduce_costs", "device-side cost sum (operand of the scalar all-reduce)"),
                          ("scale_gradients", "gradient scale"), ("ctc_combine", "bidirectional second half")):
            if key in name:
                return name.split("(")[0].split("::")[-1][:40], role
        return name.split("(")[0][:60], "torch"

    tot = sum(d["gpu__time_duration.sum"] for d in per.values())
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(grp(d["k"]), [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += d["gpu__time_duration.sum"]; a[2] += d["dram__bytes_read.sum"]; a[3] += d["dram__bytes_write.sum"]
    lines = ["ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1",
             "(round-2 code; per-launch times are cold-cache and serialised: compare SHARES, not absolutes)", ""]
    thr = [0, 0.0, 0.0, 0.0]
    for (k, role), a in agg.items():
        lines.append(f"{k:38s} launches {a[0]:4d}  time {a[1]:8.3f} ms ({100 * a[1] / tot:5.1f}%)  dram rd {a[2] / 1e6:9.1f} MB wr {a[3] / 1e6:9.1f} MB   [{role}]")
        if role.startswith("engine step"):
            for j in range(4): thr[j] += a[j]
    nvar = sum(1 for (k, role) in agg if role.startswith("engine step"))
    ncalls = max(1, thr[0] // max(1, nvar))
    lines += ["", f"engine step: {thr[0]} launches = {ncalls} engine calls x {nvar} variants;",
              f"per engine call at B=8192: {thr[1] / ncalls:.3f} ms serialised ({100 * thr[1] / tot:.1f}% of all GPU time in the run), DRAM read {thr[2] / ncalls / 1e6:.1f} MB + write {thr[3] / ncalls / 1e6:.1f} MB = {(thr[2] + thr[3]) / ncalls / 1e6:.1f} MB (algorithmic 1429.6 MB => {(thr[2] + thr[3]) / ncalls / 1429.6e6:.2f}x).",
              "Where the extra traffic goes (per utterance-frame, NS=8, K=8): activations + gradient 232 B (algorithmic); fp32 checkpoint column + its exponent row every 8 frames, written by the forward sweep and read back by the backward sweep 2 x 144 B; p~ image (one 128-byte row per frame) written once and read once 2 x 128 B (+ 1/s: 2 x 4 B)."]
    open(os.path.join(ROOT, "profiles", "r2_launches_bench.txt"), "w").write("\n".join(lines) + "\n")
    json.dump({"dram_bytes_per_launch_set": (thr[2] + thr[3]) / ncalls, "dram_read": thr[2] / ncalls, "dram_write": thr[3] / ncalls, "algorithmic_bytes": 1429607948,
               "source": f"profiles/r2_launches_bench.txt (ncu dram__bytes_read.sum + dram__bytes_write.sum over the {nvar} variant launches of one engine call, B=8192, round-2 code)"},
              open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
    print("\n".join(lines))


def kernel(rep, name, B, T, what):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], rows[2]))
    u = dict(zip(rows[0], rows[1]))
    steps = float(B) * float(T)
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_active", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    out = [f"ncu --set full --import-source on --clock-control none -k regex:ctc_warp -s 1 -c 1 python tools/profile_one.py ...   ({what})",
           f"kernel {d.get('Kernel Name', '?')}, one launch = {int(steps):,} utterance-timesteps", ""]
    for k in keys:
        if k in d:
            out.append(f"{k} [{u.get(k, '')}] = {d[k]}")
    st = sorted(((float(v), h) for h, v in d.items() if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")
                 and "not_issued" not in h), reverse=True)
    out.append("")
    out.append("stall cycles per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
    for v, h in st[:10]:
        out.append(f"  {v:6.3f}  {h.split('stalled_')[1].split('_per_issue')[0]}")
    n = float(d["smsp__inst_executed.sum"])
    wf = float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"])
    bc = float(d["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"])
    traffic = (float(d["dram__bytes_read.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6}[u["dram__bytes_read.sum"]]
               + float(d["dram__bytes_write.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6}[u["dram__bytes_write.sum"]])
    out += ["", f"per utterance-timestep: {n / steps:.0f} warp instructions, {wf / steps:.1f} shared-memory wavefronts "
                f"({bc / steps:.1f} of them bank-conflict replays), {traffic / steps:.0f} B of DRAM traffic (algorithmic 232 B)",
            "(shared-memory wavefronts include one per warp shuffle; round 2 fp64 ctc_warp_kernel<8,8,1,168,0>: 248 instructions, 55.5 wavefronts, 741 B -- profiles/r2_ncu_ns8.txt; "
            "round 1 ctc_fused_kernel<8,1,8,1>: 310 instructions, 100 wavefronts, 1016 B -- profiles/r1_e_ncu_summary.txt)"]
    # opcode mix from the SASS page
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    ops, smp, tots = collections.Counter(), collections.Counter(), 0
    for r in rows[2:]:
        if len(r) < 10:
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1].strip())
        op = m.group(2) if m else r[1]
        base = op.split(".")[0]
        if base in ("LDS", "STS", "SHFL", "LDG", "STG", "F2F", "MUFU", "LDGSTS"):
            base = ".".join(op.split(".")[:2])
        ops[base] += int(r[ix["Instructions Executed"]])
        smp[base] += int(r[ix["# Samples"]])
        tots += int(r[ix["# Samples"]])
    out += ["", "executed warp instructions per utterance-timestep by opcode (share of stall samples):"]
    for op, c in ops.most_common(22):
        out.append(f"  {op:12s} {c / steps:7.2f}   {100 * smp[op] / max(1, tots):5.1f} %")
    path = os.path.join(ROOT, "profiles", f"r2_ncu_{name}.txt")
    if os.environ.get("SUMMARY_OUT"):
        path = os.environ["SUMMARY_OUT"]
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5], sys.argv[6] if len(sys.argv) > 6 else "")

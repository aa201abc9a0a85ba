"""Summarise an ncu report of the classifier-head kernels (tools/head_dev.sh ncu) into the text committed under profiles/."""
import csv, subprocess, sys
rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/head_ncu.ncu-rep"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0]
    if not any(k in name for k in ("stats", "fwd_tc", "wgrad", "dgrad")):
        continue
    print(name)
    for k in KEYS:
        if k in d:
            print(f"  {k:90s} {d[k]}")
    stalls = []
    for k in hdr:
        if "issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k:
            try:
                v = float(d[k])
            except ValueError:
                continue
            if v > 0.15:
                stalls.append((v, k.split("stalled_")[1].split("_per")[0]))
    print("  stalls per issue: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)))
    try:
        us = float(d["gpu__time_duration.sum"])
        mb = float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])
        print(f"  => {mb / us * 1e3:.0f} GB/s of DRAM traffic = {mb / us * 1e3 / 6551.7 * 100:.1f} % of the measured HBM peak (6551.7 GB/s)")
    except Exception:
        pass
    print()

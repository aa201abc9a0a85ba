"""Per-source-line cost of a profiled kernel (ncu --set full --import-source on, -lineinfo build):
   python tools/ncu_lines.py REPORT.ncu-rep UNITS [min_per_unit]
prints, for every source line, executed warp instructions and shared-memory wavefronts per unit (e.g. utterance-frames)
and the share of stall samples."""
import csv, io, subprocess, sys


def num(x):
    return int(x) if x.isdigit() else 0

rep, units = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
fname, hdr, rows = None, None, []
for r in csv.reader(io.StringIO(raw)):
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        hdr["Source"] = 1
    elif hdr and len(r) > 10 and r[0] not in ("", "-") and r[hdr["Instructions Executed"]].isdigit():
        rows.append((fname, r))
tot_i = sum(num(r[hdr["Instructions Executed"]]) for _, r in rows)
tot_s = sum(num(r[hdr["# Samples"]]) for _, r in rows)
tot_w = sum(num(r[hdr["L1 Wavefronts Shared"]]) for _, r in rows)
print(f"total: {tot_i / units:.1f} instructions, {tot_w / units:.1f} shared wavefronts per unit, {tot_s} samples")
for f, r in rows:
    i = num(r[hdr["Instructions Executed"]]) / units
    w = num(r[hdr["L1 Wavefronts Shared"]]) / units
    s = 100.0 * num(r[hdr["# Samples"]]) / max(1, tot_s)
    if i >= thr or s >= 1.0:
        print(f"{f[:14]:14s}:{r[0]:>4s} {i:7.2f} instr {w:6.2f} wf {s:5.1f}% smp | {r[1].strip()[:110]}")

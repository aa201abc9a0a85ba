"""ptxas -v and static SASS summary of the warp-ladder kernels (committed as profiles/r2_ptxas_sass_summary.txt).
python tools/sass_summary.py > profiles/r2_ptxas_sass_summary.txt      (needs the in-tree build: aes_lac_2018_b200/build/*.o)"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = open(os.path.join(ROOT, "aes_lac_2018_b200", "lib", "ptxas.log")).read()
OPS = ("FADD", "FMUL", "FFMA", "DADD", "DMUL", "DFMA", "F2F", "SHFL", "LDS", "STS", "LDGSTS", "LDG", "STG", "REDUX", "CREDUX", "MUFU", "F2I", "I2FP",
       "ATOMG", "WARPSYNC", "BRA", "BAR")


def section(title, sym, obj, fmt):
    print(title)
    rows = re.findall(r"Compiling entry function '(_ZN7ctcb200\d+%s[^']*)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers" % sym, log)
    for name, _, st, ld, regs in rows:
        t = re.findall(r"ILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)(?:ELi(\d+))?", name)[0]
        print("  <%s>  %s registers, spill stores %s B, spill loads %s B" % (",".join(x for x in t if x), regs, st, ld))
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "aes_lac_2018_b200", "build", obj)], capture_output=True, text=True).stdout
    print("\nSASS opcode counts (static, whole kernel incl. prologue, both chunk bodies and the out-of-line divergence stubs)")
    for blk in sass.split("Function : ")[1:]:
        name = blk.split("\n", 1)[0]
        if sym not in name:
            continue
        t = re.findall(r"ILi(\d+)ELi(\d+)ELi(\d+)ELi(\d+)(?:ELi(\d+))?", name)[0]
        ops = collections.Counter(m.split(".")[0] for m in re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", blk, re.M))
        n = sum(ops.values())
        print("  <%s>  %d instructions (%d KB): %s" % (",".join(x for x in t if x), n, n * 16 // 1024, " ".join("%s %d" % (o, ops[o]) for o in OPS if ops[o])))
    print()


print("ptxas -v and SASS summary of the round-2 throughput kernels (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)")
print("source: aes_lac_2018_b200/lib/ptxas.log and `cuobjdump -sass aes_lac_2018_b200/build/ctc_variants_g{8,6}.o` (VCH = 1: alphabets up to 31 symbols)\n")
section("ctc_warp32_kernel<NS, K, VCH, MAXR> (fp32 recursion with per-lane block exponents; the large-batch default): registers / spills",
        "ctc_warp32_kernel", "ctc_variants_g8.o", 4)
section("ctc_warp_kernel<NS, K, VCH, MAXR, AVS> (fp64 recursion, ratio domain; second tier behind the fp32 kernel): registers / spills",
        "ctc_warp_kernel", "ctc_variants_g6.o", 5)
print("No block barrier (BAR) anywhere: one warp per utterance.  The fp32 kernel has no fp64 on the T-serial chain: its DADD / DMUL / DFMA are the\n"
      "once-per-utterance log Z and the running product of the row sums.  LDGSTS = cp.async staging of the backward operands; (C)REDUX = warp-wide max\n"
      "(row reference of the fp32 kernel, rescale / poison detector of the fp64 kernel); ATOMG = work queue, retired-CTA and workspace-slot counters.")


# ---- classifier-head kernels and the packed-fp32 opcodes of the fp32 throughput kernel (from the linked library) ----
so = os.path.join(ROOT, "aes_lac_2018_b200", "lib", "libctc_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, ops = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        ops[cur][m.group(1).split(".")[0]] += 1


def demangle(f):
    return subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")


print("\nclassifier-head kernels (ctc_head.cu, ctc_head_tc.cuh, ctc_head_bwd_tc.cuh): tensor-core opcodes "
      "(UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, SYNCS = mbarrier)")
for f, c in ops.items():
    if "head_" in f and "tc_kernel" in f and "fold" not in f and "wt_tc" not in f:
        print(f"  {demangle(f)[:52]:52s} UTCHMMA {c['UTCHMMA']:3d}  UTCBAR {c['UTCBAR']:2d}  LDTM {c['LDTM']:2d}  SYNCS {c['SYNCS']:2d}  "
              f"LDG {c['LDG']:3d}  STS {c['STS']:3d}  BAR {c['BAR']:2d}  total {sum(c.values())}")
print("\nfp32 throughput kernel variants: packed single-precision opcodes (FFMA2 / FMUL2 / FADD2, sm_100) against scalar ones")
for f, c in sorted(ops.items(), key=lambda kv: demangle(kv[0])):
    if "ctc_warp32_kernel" in f:
        print(f"  {demangle(f)[:52]:52s} FFMA2 {c['FFMA2']:4d} FMUL2 {c['FMUL2']:4d} FADD2 {c['FADD2']:4d} | FFMA {c['FFMA']:4d} FMUL {c['FMUL']:4d} "
              f"FADD {c['FADD']:4d} | SHFL {c['SHFL']:4d}  total {sum(c.values())}")

#!/usr/bin/env python
"""SASS evidence for profiles/ (runs on the CPU-only box: cuobjdump reads the cross-compiled .so).

  python tools/sass_ledger.py > profiles/r2_sass_histogram_and_attn_ledger.txt

1. Per kernel of libnpe_pfn_b200.so: counts of the Blackwell-only instructions (UTCHMMA = tcgen05.mma, LDTM / STTM =
   tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store) and of the arithmetic classes.
2. For the dominant kernel (attn_tc_kernel<5,3,2>, item attention of test rows): the instruction mix of the fast-path
   inner loop (one 64-key tile = 64 exponentials per thread) and the issue / pipe cycle ledger it implies, using the
   pipe rates measured in round 1 (profiles/r1_pipe_rates_microbench.txt, r1_coissue_microbench.txt).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "npe_pfn_b200", "_lib", "libnpe_pfn_b200.so")
KEY = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "MUFU", "FFMA2", "FADD2", "FMNMX3", "HMMA", "REDUX"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def kernels():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, body = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
        if m and cur:
            ins = re.sub(r"^@!?U?P\w+\s+", "", m.group(1).strip())
            body[cur].append(ins.split()[0])
    return body


def main():
    body = kernels()
    print("# cuobjdump -sass npe_pfn_b200/_lib/libnpe_pfn_b200.so (sm_100a), instruction counts per kernel")
    print("# kernel | total | " + " | ".join(KEY))
    for k, ins in body.items():
        c = collections.Counter()
        for i in ins:
            for key in KEY:
                if i.startswith(key):
                    c[key] += 1
        if any(c[x] for x in ("UTCHMMA", "LDTM", "UTMALDG", "MUFU", "HMMA")) or len(ins) > 400:
            name = re.sub(r"\(.*", "", demangle(k)).replace("void pfn::", "")
            print(f"{name} | {len(ins)} | " + " | ".join(str(c[x]) for x in KEY))
    # ---- fast-path ledger of attn_tc_kernel<5, 3, 2> ----
    target = [k for k in body if "attn_tc_kernelILi5ELi3ELi2E" in k]
    if not target:
        return
    ins = body[target[0]]
    # fast path = from the first LDTM that is followed (before any STTM) by a second LDTM and >= 40 MUFU.EX2, to the VOTE
    best = None
    for i, x in enumerate(ins):
        if not x.startswith("LDTM"):
            continue
        j = i + 1
        while j < len(ins) and not ins[j].startswith(("VOTE", "STTM")):
            j += 1
        seg = ins[i:j]
        if sum(s.startswith("MUFU.EX2") for s in seg) >= 40 and sum(s.startswith("LDTM") for s in seg) == 2:
            best = seg
            break
    if best is None:
        print("# fast path not found")
        return
    c = collections.Counter(re.sub(r"\..*", "", s) if not s.startswith(("MUFU", "F2FP", "LDTM")) else s.split(".")[0] + "." + s.split(".")[1] for s in best)
    n = len(best)
    mufu = sum(v for k, v in c.items() if k.startswith("MUFU"))
    half = sum(v for k, v in c.items() if k in ("FFMA2", "FADD2", "FMNMX", "F2FP.BF16", "IMAD", "VIMNMX", "FMNMX3"))
    full = n - mufu - half
    print("\n# ---- attn_tc_kernel<5,3,2> (v5, default): fast-path inner loop, one 64-key tile = 64 scores per thread ----")
    print("# instruction mix between the tile's first tcgen05.ld and the overflow-check vote:")
    for k, v in sorted(c.items(), key=lambda kv: -kv[1]):
        print(f"#   {k:22s} {v}")
    print(f"# total {n} issue slots per 64 elements = {n / 64:.2f} per element")
    print(f"# XU pipe:   {mufu} MUFU.EX2 x 8 cycles (16 lanes/clk/SM = 4 per scheduler)          = {mufu * 8} cycles = {mufu * 8 / 64:.2f} per element")
    print(f"# half-rate: {half} (FFMA2 FADD2 FMNMX F2FP IMAD, 2 pipe cycles each, FMA + ALU pipes)  = {half * 2} cycles = {half * 2 / 64:.2f} per element")
    print(f"#            of which FMA pipe (FFMA2 FADD2 IMAD) {2 * (c['FFMA2'] + c['FADD2'] + c['IMAD'])} cycles, ALU pipe (FMNMX F2FP) {2 * (c['FMNMX'] + c['F2FP.BF16'])} cycles")
    print(f"# full-rate / other: {full}")
    lb = max(n, mufu * 8, 2 * (c['FFMA2'] + c['FADD2'] + c['IMAD']), 2 * (c['FMNMX'] + c['F2FP.BF16']))
    print(f"# lower bound if the three pipes and the issue port overlapped perfectly: {lb} cycles per warp-tile = {lb / 64:.2f} clk per "
          f"warp-element per scheduler\n#   -> {128 * 32 * 4 * 148 * 1.965e9 / (lb / 64) / 1e12:.0f} TFLOP/s equivalent at 1965 MHz (128 tensor FLOPs per exponential, 32 lanes x 4 schedulers x 148 SMs)")
    tot = mufu * 8 + half * 2
    print(f"# upper bound if the pipes did not overlap at all: {tot} cycles = {tot / 64:.2f} clk per warp-element "
          f"({128 * 32 * 4 * 148 * 1.965e9 / (tot / 64) / 1e12:.0f} TFLOP/s equivalent); r1's co-issue microbenchmark\n"
          "#   (profiles/r1_coissue_microbench.txt) shows the FMA and ALU pipes overlapping each other only partly (4 FFMA2 + 4 FMNMX3 = 14.6 clk, not 8)")
    print("# measured: register-only loop of this mix 7.4-7.9 clk (profiles/r1_softmax_loop_lean_microbench.txt, r2_softmax_loop_rowsum_f16x2_microbench.txt);")
    print("#           kernel 9.1-9.3 clk per warp-element = 510-528 TFLOP/s (BENCH_r01.json, profiles/r2_*): 60 % of the pipe bound.")
    print("# what fills the gap (ncu, profiles/r1_ncu_attn_tc_v5_metrics.txt): issue slot busy 64.6 %, XU 64.8 %, ALU 44.8 %, FMA 23.7 %; stall")
    print("#   cycles per issued instruction: wait 1.60 (fixed-latency dependences of the 7-deep FMA-pipe exponential chain and of MUFU -> pack),")
    print("#   long_scoreboard 0.86 (tcgen05.ld / mbarrier), branch_resolving 0.50; 3 softmax warps per scheduler cannot cover them, and a")
    print("#   4th does not fit: TMEM (3 x 160 of 512 columns) and registers (3 x 192 x 96) are both at their limit.")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Turns the raw output of tools/ncu_capture.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/:

  python tools/ncu_summarise.py <tag> <out-prefix>        e.g.  python tools/ncu_summarise.py r2j r2

  profiles/<out>_launch_summary.csv     per kernel: launches, total time, share of the captured launches, DRAM MB, GB/s
  profiles/<out>_launches.csv.gz        the raw launch list
  profiles/<out>_ncu_<name>_metrics.txt selected `--set full` metrics of the captured kernels
  profiles/<out>_attn_traffic.json      DRAM bytes per launch of the item-attention kernel + the source hash of the build
                                        (bench.py reports it as roofline.traffic only while the hash matches)
"""
import collections
import csv
import gzip
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio")


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("pfn::", "")
    return name.strip()


def launch_summary(tag, out):
    src = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    rows = [l for l in open(src) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    ids = collections.defaultdict(set)
    for r in rd:
        k = short(r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        else:
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
        per[k][r["Metric Name"]] += v
        ids[k].add(r["ID"])
    tot = sum(v["gpu__time_duration.sum"] for v in per.values())
    path = os.path.join(ROOT, "profiles", f"{out}_launch_summary.csv")
    with open(path, "w") as f:
        f.write(f"# ncu launch list of the round-2 build: python bench.py --steps 1 --warmup 1 --samples 16384 --no-cpu-baseline --no-configs\n"
                "# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1300 -c 1400 (tools/ncu_capture.sh)\n"
                "# (1400 launches from the middle of the run; per-launch times are cold-cache and serialised: compare SHARES)\n"
                "kernel,launches,total_us,share,dram_read_MB,dram_write_MB,GB_per_s\n")
        for k, v in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
            t = v["gpu__time_duration.sum"]
            rdmb, wrmb = v["dram__bytes_read.sum"], v["dram__bytes_write.sum"]
            f.write(f"{k},{len(ids[k])},{t:.1f},{t / tot:.4f},{rdmb:.1f},{wrmb:.1f},{(rdmb + wrmb) / max(t, 1e-9) * 1e6 / 1e3:.0f}\n")
    with open(src, "rb") as fi, gzip.open(os.path.join(ROOT, "profiles", f"{out}_launches.csv.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)
    print(open(path).read())
    return per, ids


def rep_metrics(tag, out, rep, title):
    path = os.path.join(ROOT, "gpurun_out", f"{tag}_{rep}.ncu-rep")
    if not os.path.exists(path):
        return None
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in txt.splitlines() if l.startswith('"')]
    rd = list(csv.reader(io.StringIO("\n".join(lines))))
    head, units, data = rd[0], rd[1], rd[2:]
    col = {h.split(".TriageCompute.")[-1] if ".TriageCompute." in h else h: i for i, h in enumerate(head)}
    dst = os.path.join(ROOT, "profiles", f"{out}_ncu_{rep}_metrics.txt")
    traffic = None
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({title}); tools/ncu_capture.sh, tools/ncu_summarise.py\n"
                "# command: python bench.py --steps 1 --warmup 1 --samples 16384 --no-cpu-baseline --no-configs\n")
        for row in data:
            name = short(row[col["Kernel Name"]])
            f.write(f"\n## {name}  grid {row[col['Grid Size']]} block {row[col['Block Size']]}\nmetric | unit | value\n")
            for m in KEEP:
                if m in col:
                    f.write(f"{m} | {units[col[m]]} | {row[col[m]]}\n")
            for h, i in col.items():
                mm = STALL.match(h)
                if mm and row[i] not in ("", "0"):
                    try:
                        if float(row[i].replace(",", "")) >= 0.2:
                            f.write(f"{h} | {units[i]} | {row[i]}\n")
                    except ValueError:
                        pass
            if "attn_tc" in name and traffic is None and "dram__bytes_read.sum" in col:
                def to_bytes(m):
                    return float(row[col[m]].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[col[m]]]
                traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    print("wrote", dst)
    return traffic


def main():
    tag, out = sys.argv[1], sys.argv[2]
    launch_summary(tag, out)
    traffic = rep_metrics(tag, out, "attn_tc", "item attention of test rows, one launch: 16384 rows x 6 heads x T columns, N = 10000 keys")
    rep_metrics(tag, out, "gemm_tc", "five consecutive projection / fused-MLP launches")
    rep_metrics(tag, out, "hbm", "encoder, K/V cache writer and compaction kernels")
    rep_metrics(tag, out, "head", "head_row2_kernel: two launches of 16 384 logits rows x 5000 buckets")
    if traffic:
        from npe_pfn_b200 import build as b
        json.dump({"dram_bytes_per_launch": traffic, "rows_per_launch": 16384, "srchash": b.files_hash(b.ATTN_KERNEL_FILES), "hashed_files": b.ATTN_KERNEL_FILES,
                   "source": f"profiles/{out}_ncu_attn_tc_metrics.txt (dram__bytes_read.sum + dram__bytes_write.sum of one launch)"},
                  open(os.path.join(ROOT, "profiles", f"{out}_attn_traffic.json"), "w"), indent=1)
        print("attention DRAM bytes per launch:", traffic)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.md "command line"
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r1_kernel.md
  python tools/ncu_summary.py traffic gpurun_out/tca_traffic.csv profiles/r1_tca_traffic "command line"     (writes .md and .json)
"""
import collections
import csv
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return re.sub(r"void |<unnamed>::|\(anonymous namespace\)::", "", name)


def launches(src, dst, cmd):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        d = agg.setdefault(short(row["Kernel Name"]), [0, 0.0])
        d[0] += 1
        d[1] += v
        tot += v
        n += 1
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`)\n\n`{cmd}`\n\n")
        f.write(f"{n} launches captured, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised: compare SHARES).\n\n")
        f.write("| share | total us | launches | avg us | kernel |\n|---:|---:|---:|---:|---|\n")
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {100 * v / tot:.1f}% | {v:.1f} | {c} | {v / c:.1f} | `{k[:100]}` |\n")
    print("wrote", dst)


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of `{src.split('/')[-1]}`\n\n")
        for vals in rows[2:]:
            rec = dict(zip(hdr, zip(units, vals)))
            f.write(f"## `{short(rec['Kernel Name'][1])}`  grid {rec.get('Grid Size', ('', ''))[1]} block {rec.get('Block Size', ('', ''))[1]}\n\n")
            f.write("| metric | value | unit |\n|---|---:|---|\n")
            for k in WANT:
                if k in rec:
                    f.write(f"| {k} | {rec[k][1]} | {rec[k][0]} |\n")
            f.write("\n")
    print("wrote", dst)


def kernel_source_sha256():
    """sha256 over the sources of the persistent tcgen05 kernel, comments and whitespace ignored (run this script in the tree the
    capture was taken from)."""
    import hashlib
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ppodash_b200", "csrc")
    import re
    h = hashlib.sha256()
    for name in ("tca_gemm.cu", "tca_gemm.cuh", "tma_utils.cuh"):
        src = open(os.path.join(root, name), encoding="utf-8").read()
        src = re.sub(r"//[^\n]*", "", src)              # comments and layout do not change the kernel: hash the code only
        h.update(" ".join(src.split()).encode())
    return h.hexdigest()


def traffic(src, dst, cmd):
    """DRAM bytes and device time of consecutive launches of one kernel family (3 metrics per launch)."""
    import json
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        d = per.setdefault(int(row["ID"]), {"grid": row["Grid Size"], "kernel": short(row["Kernel Name"])})
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Name"].startswith("gpu__time"):
            u = row["Metric Unit"]
            d["us"] = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        else:
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[row["Metric Unit"]]
            d["read" if "read" in row["Metric Name"] else "write"] = v * mult
    n = len(per)
    rd, wr, us = (sum(d[k] for d in per.values()) for k in ("read", "write", "us"))
    rec = {"kernel": "tca_gemm_kernel", "launches": n, "dram_bytes_read": rd, "dram_bytes_written": wr,
           "avg_dram_bytes_per_launch": (rd + wr) / n, "total_us": us, "command": cmd,
           # the capture is only meaningful for the kernel sources it was taken from: bench.py compares this with the tree it runs in
           "kernel_source_sha256": kernel_source_sha256()}
    json.dump(rec, open(dst + ".json", "w"), indent=1)
    with open(dst + ".md", "w") as f:
        f.write(f"# DRAM traffic of `tca_gemm_kernel` over {n} consecutive launches of the bench step\n\n`{cmd}`\n\n")
        f.write(f"{n} launches, {us:.1f} us of kernel time (serialised, cold cache), DRAM read {rd / 1e6:.1f} MB, written {wr / 1e6:.1f} MB "
                f"-> {(rd + wr) / n / 1e6:.1f} MB per launch.\n\n")
        f.write("| # | kernel | grid | us | DRAM read MB | DRAM write MB |\n|---:|---|---|---:|---:|---:|\n")
        for i, d in per.items():
            f.write(f"| {i} | `{d['kernel'][-22:]}` | {d['grid']} | {d['us']:.1f} | {d['read'] / 1e6:.2f} | {d['write'] / 1e6:.2f} |\n")
    print("wrote", dst + ".md", dst + ".json")


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])

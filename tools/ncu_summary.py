#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.md "command line"
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r1_kernel.md
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


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])

#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from an ncu report:  python tools/ncu_hot.py file.ncu-rep [N]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for idx, r in enumerate(rows[2:]):
    try:
        data.append((int(r[ci["# Samples"]]), idx, r[ci["Source"]].strip()[:100], int(r[ci["Instructions Executed"]])))
    except Exception:
        pass
tot = sum(d[0] for d in data) or 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("total samples", tot, "instructions", len(data))
for s, idx, t, ex in sorted(data, reverse=True)[:n]:
    print(f"{100*s/tot:5.1f}%  #{idx:4d} exec={ex:8d}  {t}")

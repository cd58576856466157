#!/usr/bin/env python
"""Warp-stall samples per CUDA source line (needs -lineinfo and --import-source on):
   python tools/ncu_src.py file.ncu-rep [N]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if "# Samples" in r)
H = rows[hdr]
ci = {h: i for i, h in enumerate(H)}
stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
per_line = collections.OrderedDict()
cur = None
tot = 0
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    if r[0].strip():                      # a CUDA source line row: "Line No","Source"
        cur = (r[0], r[1].strip()[:110])
        per_line.setdefault(cur, [0, collections.Counter()])
        continue
    try:
        s = int(r[ci["# Samples"]])
    except ValueError:
        continue
    tot += s
    if cur is None:
        cur = ("?", "?")
        per_line.setdefault(cur, [0, collections.Counter()])
    per_line[cur][0] += s
    for i in stall_cols:
        try:
            per_line[cur][1][H[i]] += int(r[i])
        except ValueError:
            pass
print("total samples", tot)
for (ln, src), (s, st) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:n]:
    top = ", ".join(f"{k[6:]}={v}" for k, v in st.most_common(3))
    print(f"{100*s/max(tot,1):5.1f}%  L{ln:>4}  {src}\n            [{top}]")

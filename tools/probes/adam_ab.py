"""clip+Adam: fused cooperative launch vs three plain launches, lone flushed launches and back-to-back (one box)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import microbench as mb
from ppodash_b200 import _lib
L = _lib.lib()
for fused in (1, 0, 1, 0):
    L.ppd_clip_adam_set_fused(fused)
    r = mb.bench_adam(2464576)
    print("fused", fused, "lone %.2f us  back-to-back %.2f us" % (1e3 * r["ms_single_flushed"], 1e3 * r["ms"]))

"""Sweep of the returns-scan tuning knobs at 4096 envs x 2048 steps (one box, L2 flushed between iterations)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import microbench as mb
from ppodash_b200 import _lib
L = _lib.lib()
for w in (4, 8, 16):
    for mbk in (3, 4):
        L.ppd_compute_returns_set_tuning(w, mbk)
        L.ppd_compute_returns_set_tuning(100, 0)
        r = mb.bench_gae(2048, 4096)
        print("warps", w, "min_blocks", mbk, "ms %.4f best %.4f gbs %.0f" % (r["ms"], r["ms_best"], r["gbs"]))
L.ppd_compute_returns_set_tuning(8, 3)
for st, ct in ((2, 2), (3, 2), (4, 1), (2, 3)):
    L.ppd_compute_returns_set_tuning(102, 0); L.ppd_compute_returns_set_tuning(200 + st * 10 + ct, 0)
    r = mb.bench_gae(2048, 4096)
    print("tma stages", st, "ctas", ct, "ms %.4f best %.4f gbs %.0f" % (r["ms"], r["ms_best"], r["gbs"]))

"""Per-phase clock timeline of the one-env register GRU kernels (CTA 0; instrumented build: make -C ppodash_b200/csrc probes; stamps go to shared memory).
PPD_LIB=$PWD/ppodash_b200/libppd_grutrace.so python tools/probes/gru_trace.py"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from ppodash_b200 import _lib
import gru_bench

gru_bench.main()
torch.cuda.synchronize()
raw = ctypes.CDLL(_lib.LIB_PATH)
buf = np.zeros((2, 32, 8), dtype=np.uint32)
assert raw.ppd_gru_trace_read(ctypes.c_void_p(buf.ctypes.data)) == 0
buf = buf.astype(np.int64)
names = (("forward (thread 0 = gate warp; last two columns: a non-gate warp)", ["top", "h arrived", "matvec done", "barrier 1", "gates done", "barrier 2", "w11: h arrived", "w11: matvec done"]),
         ("backward (thread 0 = gate warp)", ["top", "gates done", "barrier 1", "matvec done", "pushed", "partials arrived", "carry done", "barrier 2"]))
for k, (title, cols) in enumerate(names):
    tr = buf[k]
    print(title)
    print("  step " + " ".join(f"{c:>16s}" for c in cols))
    t0 = tr[8, 0]
    for i in range(8, 14):
        print(f"  {256 + i:4d} " + " ".join(f"{tr[i, j] - t0:16d}" for j in range(8)))
    print("  median clocks per step:", np.median(np.diff(tr[:, 0])))
    n = 6 if k == 0 else 8
    d = np.diff(tr[:, :n], axis=1)
    print("  median phase clocks: " + "  ".join(f"{cols[j]}->{cols[j + 1]} {np.median(d[:, j]):.0f}" for j in range(n - 1)) + f"  {cols[n - 1]}->next top {np.median(tr[1:, 0] - tr[:-1, n - 1]):.0f}")

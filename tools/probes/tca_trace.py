"""Pipeline timeline of the persistent GEMM (library built with -DPPD_TCA_TRACE)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ppodash_b200 import _lib
from ppodash_b200._lib import GemmArgs
L = _lib.lib()
raw = ctypes.CDLL(_lib.LIB_PATH)
dev = "cuda:0"
name, I, J, KK, ak, bk = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
A = torch.randn((I, KK) if ak else (KK, I), device=dev)
B = torch.randn((J, KK) if bk else (KK, J), device=dev)
C = torch.zeros(I, J, device=dev)
g = GemmArgs()
g.A, g.lda, g.a_kmajor = A.data_ptr(), A.shape[1], ak
g.B, g.ldb, g.b_kmajor = B.data_ptr(), B.shape[1], bk
g.C, g.ldc, g.I, g.J, g.KK = C.data_ptr(), J, I, J, KK
ws = torch.empty(max(256, L.ppd_tc_gemm_workspace(I, J, KK)), dtype=torch.uint8, device=dev)
for _ in range(3):
    _lib.check(L.ppd_tc_gemm(ctypes.byref(g), 2, ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
torch.cuda.synchronize()
tr = torch.zeros(256 * 16, dtype=torch.int64, device=dev)
raw.ppd_tca_set_trace.argtypes = [ctypes.c_void_p]
assert raw.ppd_tca_set_trace(tr.data_ptr()) == 0
_lib.check(L.ppd_tc_gemm(ctypes.byref(g), 2, ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
torch.cuda.synchronize()
t = tr.cpu().view(256, 16)
t0 = int(t[0, 0])
names = ["A.issue", "B.issue", "X.wait", "X.full", "X.read", "X.ta", "X.st", "X.done", "M.ready", "M.issued"]
print(name, "k-block timeline of CTA 0 (clocks since first A issue)")
print("  it " + " ".join(f"{n:>9s}" for n in names))
for it in range(40, 72):
    print(f"{it:4d} " + " ".join(f"{int(t[it, k]) - t0:9d}" if int(t[it, k]) else "        -" for k in range(10)))
d = (t[200, 9] - t[40, 9]).item() / 160
print("clocks per k-block (steady state):", d)

"""Pipeline timeline of the implicit conv kernels (library built with -DPPD_TCA_TRACE):  conv_trace.py <case>"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ppodash_b200 import _lib
from ppodash_b200._lib import ConvGeom
L = _lib.lib()
raw = ctypes.CDLL(_lib.LIB_PATH)
dev = "cuda:0"
B = 2048
case = sys.argv[1]
st = _lib.stream_ptr()
a1 = torch.randn(B, 20, 20, 32, device=dev); dy2 = torch.randn(B, 9, 9, 64, device=dev)
obs = torch.randn(B, 3, 84, 84, device=dev); dy1 = torch.randn(B, 20, 20, 32, device=dev)
g2 = ConvGeom(B, 20, 20, 32, 4, 4, 2); g1 = ConvGeom(B, 84, 84, 3, 8, 8, 4)
gw2 = torch.zeros(64, 512, device=dev); gw1 = torch.zeros(32, 192, device=dev)
ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
w2 = torch.randn(64, 512, device=dev); hi2, lo2 = torch.empty_like(w2), torch.empty_like(w2)
_lib.check(L.ppd_split_tf32(w2.data_ptr(), hi2.data_ptr(), lo2.data_ptr(), w2.numel(), st))
w1 = torch.randn(32, 192, device=dev); hi1, lo1 = torch.empty_like(w1), torch.empty_like(w1)
_lib.check(L.ppd_split_tf32(w1.data_ptr(), hi1.data_ptr(), lo1.data_ptr(), w1.numel(), st))
o1 = torch.empty(B, 20, 20, 32, device=dev); b1 = torch.zeros(32, device=dev)
o2 = torch.empty(B, 9, 9, 64, device=dev); b2 = torch.zeros(64, device=dev); dx1 = torch.empty_like(a1)
fns = {
    "conv2.wgrad": lambda: L.ppd_conv_wgrad(a1.data_ptr(), ctypes.byref(g2), 0, dy2.data_ptr(), 64, gw2.data_ptr(), 0, ws.data_ptr(), ws.numel(), st),
    "conv1.wgrad": lambda: L.ppd_conv_wgrad(obs.data_ptr(), ctypes.byref(g1), 1, dy1.data_ptr(), 32, gw1.data_ptr(), 0, ws.data_ptr(), ws.numel(), st),
    "conv2.fwd": lambda: L.ppd_conv_fwd_nhwc(a1.data_ptr(), ctypes.byref(g2), 64, hi2.data_ptr(), lo2.data_ptr(), b2.data_ptr(), 1, o2.data_ptr(), st),
    "conv1.fwd": lambda: L.ppd_conv_fwd_nchw(obs.data_ptr(), ctypes.byref(g1), 32, hi1.data_ptr(), lo1.data_ptr(), b1.data_ptr(), 1, o1.data_ptr(), st),
    "conv2.dgrad": lambda: L.ppd_conv_dgrad_nhwc(dy2.data_ptr(), ctypes.byref(g2), 64, hi2.data_ptr(), lo2.data_ptr(), a1.data_ptr(), dx1.data_ptr(), st),
}
fn = fns[case]
for _ in range(3):
    _lib.check(fn())
torch.cuda.synchronize()
tr = torch.zeros(256 * 16, dtype=torch.int64, device=dev)
raw.ppd_tca_set_trace.argtypes = [ctypes.c_void_p]
assert raw.ppd_tca_set_trace(tr.data_ptr()) == 0
_lib.check(fn())
torch.cuda.synchronize()
t = tr.cpu().view(256, 16)
t0 = int(t[0, 0])
names = ["A.issue", "B.issue", "X.wait", "X.full", "X.read", "X.ta", "X.st", "X.done", "M.ready", "M.issued", "A2.issue"]
print(case, "k-block timeline of CTA 0 (clocks since first A issue)")
print("  it " + " ".join(f"{n:>9s}" for n in names))
for it in range(40, 60):
    print(f"{it:4d} " + " ".join(f"{int(t[it, k]) - t0:9d}" if int(t[it, k]) else "        -" for k in range(11)))
print("clocks per k-block (steady state):", (t[200, 9] - t[40, 9]).item() / 160)

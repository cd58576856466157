"""Low-perturbation timeline of the persistent tcgen05 kernel (instrumented build: make -C ppodash_b200/csrc probes; clock64 stamps go to shared
memory): k-blocks 32..55 of CTA 0.   PPD_LIB=.../libppd_trace2.so python tools/probes/tca_trace2.py <conv1.fwd|conv2.fwd|conv2.dgrad|conv1.wgrad|conv2.wgrad|fc.fwd>"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ppodash_b200 import _lib
from ppodash_b200._lib import ConvGeom, GemmArgs
L = _lib.lib()
raw = ctypes.CDLL(_lib.LIB_PATH)
dev = "cuda:0"
B = 2048
case = sys.argv[1]
st = _lib.stream_ptr()
def split(w):
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    _lib.check(L.ppd_split_tf32(w.data_ptr(), hi.data_ptr(), lo.data_ptr(), w.numel(), st))
    return hi, lo
if case == "fc.fwd":
    A = torch.randn(B, 1568, device=dev); W = torch.randn(512, 1568, device=dev); C = torch.zeros(B, 512, device=dev)
    hi, lo = split(W)
    g = GemmArgs(); g.A, g.lda, g.a_kmajor = A.data_ptr(), 1568, 1; g.B, g.ldb, g.b_kmajor = hi.data_ptr(), 1568, 1
    g.C, g.ldc, g.I, g.J, g.KK = C.data_ptr(), 512, B, 512, 1568
    ws = torch.empty(max(256, L.ppd_tc_gemm_workspace(B, 512, 1568)), dtype=torch.uint8, device=dev)
    fn = lambda: L.ppd_tc_gemm_bsplit(ctypes.byref(g), lo.data_ptr(), 2, ws.data_ptr(), ws.numel(), st)
else:
    a1 = torch.randn(B, 20, 20, 32, device=dev); dy2 = torch.randn(B, 9, 9, 64, device=dev)
    obs = torch.randn(B, 3, 84, 84, device=dev); dy1 = torch.randn(B, 20, 20, 32, device=dev)
    g2 = ConvGeom(B, 20, 20, 32, 4, 4, 2); g1 = ConvGeom(B, 84, 84, 3, 8, 8, 4)
    gw2 = torch.zeros(64, 512, device=dev); gw1 = torch.zeros(32, 192, device=dev)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    hi2, lo2 = split(torch.randn(64, 512, device=dev)); hi1, lo1 = split(torch.randn(32, 192, device=dev))
    o1 = torch.empty(B, 20, 20, 32, device=dev); b1 = torch.zeros(32, device=dev)
    o2 = torch.empty(B, 9, 9, 64, device=dev); b2 = torch.zeros(64, device=dev); dx1 = torch.empty_like(a1)
    fn = {
        "conv2.wgrad": lambda: L.ppd_conv_wgrad(a1.data_ptr(), ctypes.byref(g2), 0, dy2.data_ptr(), 64, gw2.data_ptr(), 0, ws.data_ptr(), ws.numel(), st),
        "conv1.wgrad": lambda: L.ppd_conv_wgrad(obs.data_ptr(), ctypes.byref(g1), 1, dy1.data_ptr(), 32, gw1.data_ptr(), 0, ws.data_ptr(), ws.numel(), st),
        "conv2.fwd": lambda: L.ppd_conv_fwd_nhwc(a1.data_ptr(), ctypes.byref(g2), 64, hi2.data_ptr(), lo2.data_ptr(), b2.data_ptr(), 1, o2.data_ptr(), st),
        "conv1.fwd": lambda: L.ppd_conv_fwd_nchw(obs.data_ptr(), ctypes.byref(g1), 32, hi1.data_ptr(), lo1.data_ptr(), b1.data_ptr(), 1, o1.data_ptr(), st),
        "conv2.dgrad": lambda: L.ppd_conv_dgrad_nhwc(dy2.data_ptr(), ctypes.byref(g2), 64, hi2.data_ptr(), lo2.data_ptr(), a1.data_ptr(), dx1.data_ptr(), st),
    }[case]
for _ in range(3):
    _lib.check(fn())
torch.cuda.synchronize()
full = np.zeros(24 * 8 + 8, dtype=np.uint32)
assert raw.ppd_tca_trace2_read(ctypes.c_void_p(full.ctypes.data)) == 0
buf = full[:24 * 8].reshape(24, 8)
extra = full[24 * 8:].astype(np.int64)
names = ["X.top", "X.full", "X.split", "X.ta", "X.done(pair)", "M.top(pair)", "M.ready", "M.issued"]
buf = buf.astype(np.int64)
t0 = extra[0]
print('tile 7: transform (group 0) tile top', extra[4] - t0, 'after decode', extra[5] - t0, '| issuer tile top', extra[6] - t0, 'after decodes', extra[7] - t0)
print('kernel entry 0, first pair at the issuer', extra[2] - t0, ', first accumulator complete (epilogue)', extra[3] - t0, ', exit', extra[1] - t0)
print(case, "k-blocks 36..55 of CTA 0, clocks (stamps of transform group warps with q = 0 / the issuer; pair stamps sit on the pair's last / first k-block)")
print("  it " + " ".join(f"{n:>13s}" for n in names))
for i in range(4, 24):
    print(f"{32 + i:4d} " + " ".join(f"{buf[i, k] - t0:13d}" if buf[i, k] else "            -" for k in range(8)))
m = buf[:, 7]; m = m[m > 0]
print("clocks per k-block (issuer, steady state):", (m[-1] - m[0]) / (2 * (len(m) - 1)) if len(m) > 1 else None)

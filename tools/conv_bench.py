#!/usr/bin/env python
"""Timings of the implicit-GEMM convolution kernels at the PPO-Dash (c2) minibatch shapes against their algorithmic HBM bytes."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ppodash_b200 import _lib  # noqa: E402
from ppodash_b200._lib import ConvGeom  # noqa: E402
from microbench import time_kernel  # noqa: E402

DEV = "cuda:0"
B = int(os.environ.get("PPD_B", 2048))
PEAK = 6448.7


def split(w):
    L = _lib.lib()
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    _lib.check(L.ppd_split_tf32(w.data_ptr(), hi.data_ptr(), lo.data_ptr(), w.numel(), _lib.stream_ptr()))
    return hi, lo


def main():
    L = _lib.lib()
    st = _lib.stream_ptr()
    if os.environ.get("PPD_RESIDENT") is not None:
        L.ppd_tc_gemm_set_option(6 + int(os.environ["PPD_RESIDENT"]))
    if os.environ.get("PPD_BRES") is not None:
        L.ppd_tc_gemm_set_option(8 + int(os.environ["PPD_BRES"]))
    only = os.environ.get("PPD_SHAPES")
    obs = torch.randn(B, 3, 84, 84, device=DEV)
    a1 = torch.randn(B, 20, 20, 32, device=DEV)
    a2 = torch.randn(B, 9, 9, 64, device=DEV)
    a3 = torch.empty(B, 7, 7, 32, device=DEV)
    dy1, dy2, dy3 = torch.randn_like(a1), torch.randn_like(a2), torch.randn(B, 7, 7, 32, device=DEV)
    w1, w2, w3 = torch.randn(32, 192, device=DEV), torch.randn(64, 512, device=DEV), torch.randn(32, 576, device=DEV)
    b1, b2, b3 = torch.zeros(32, device=DEV), torch.zeros(64, device=DEV), torch.zeros(32, device=DEV)
    s1, s2, s3 = split(w1), split(w2), split(w3)
    g1, g2, g3 = ConvGeom(B, 84, 84, 3, 8, 8, 4), ConvGeom(B, 20, 20, 32, 4, 4, 2), ConvGeom(B, 9, 9, 64, 3, 3, 1)
    o1, o2 = torch.empty_like(a1), torch.empty_like(a2)
    dx1, dx2 = torch.empty_like(a1), torch.empty_like(a2)
    gw1, gw2, gw3 = torch.zeros_like(w1), torch.zeros_like(w2), torch.zeros_like(w3)
    ws = torch.empty(max(L.ppd_conv_wgrad_workspace(ctypes.byref(g), c) for g, c in ((g1, 32), (g2, 64), (g3, 32))), dtype=torch.uint8, device=DEV)
    mb = lambda *ts: sum(t.numel() * 4 for t in ts)
    cases = [
        ("conv1.fwd", lambda: L.ppd_conv_fwd_nchw(obs.data_ptr(), ctypes.byref(g1), 32, s1[0].data_ptr(), s1[1].data_ptr(), b1.data_ptr(), 1, o1.data_ptr(), st), mb(obs, o1)),
        ("conv2.fwd", lambda: L.ppd_conv_fwd_nhwc(a1.data_ptr(), ctypes.byref(g2), 64, s2[0].data_ptr(), s2[1].data_ptr(), b2.data_ptr(), 1, o2.data_ptr(), st), mb(a1, o2)),
        ("conv3.fwd", lambda: L.ppd_conv_fwd_nhwc(a2.data_ptr(), ctypes.byref(g3), 32, s3[0].data_ptr(), s3[1].data_ptr(), b3.data_ptr(), 1, a3.data_ptr(), st), mb(a2, a3)),
        ("conv3.dgrad", lambda: L.ppd_conv_dgrad_nhwc(dy3.data_ptr(), ctypes.byref(g3), 32, s3[0].data_ptr(), s3[1].data_ptr(), a2.data_ptr(), dx2.data_ptr(), st), mb(dy3, a2, dx2)),
        ("conv2.dgrad", lambda: L.ppd_conv_dgrad_nhwc(dy2.data_ptr(), ctypes.byref(g2), 64, s2[0].data_ptr(), s2[1].data_ptr(), a1.data_ptr(), dx1.data_ptr(), st), mb(dy2, a1, dx1)),
        ("conv3.wgrad", lambda: L.ppd_conv_wgrad(a2.data_ptr(), ctypes.byref(g3), 0, dy3.data_ptr(), 32, gw3.data_ptr(), 0, ws.data_ptr(), ws.numel(), st), mb(a2, dy3)),
        ("conv2.wgrad", lambda: L.ppd_conv_wgrad(a1.data_ptr(), ctypes.byref(g2), 0, dy2.data_ptr(), 64, gw2.data_ptr(), 0, ws.data_ptr(), ws.numel(), st), mb(a1, dy2)),
        ("conv1.wgrad", lambda: L.ppd_conv_wgrad(obs.data_ptr(), ctypes.byref(g1), 1, dy1.data_ptr(), 32, gw1.data_ptr(), 0, ws.data_ptr(), ws.numel(), st), mb(obs, dy1)),
    ]
    tot = 0.0
    for name, fn, nbytes in cases:
        if only and name not in only.split(","):
            continue
        med, best = time_kernel(lambda: _lib.check(fn()), iters=8, warmup=2)
        tot += med
        print(json.dumps(dict(conv=name, ms=round(med, 4), algorithmic_mb=round(nbytes / 1e6, 1), gbs=round(nbytes / med / 1e6, 1),
                              frac_of_hbm_peak=round(nbytes / med / 1e6 / PEAK, 3))))
    print(json.dumps({"total_ms_per_minibatch": round(tot, 4), "B": B}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-GEMM timings for the PPO-Dash (c2) minibatch shapes: SIMT fp32 vs tcgen05 tf32 / 3xTF32."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ppodash_b200 import _lib  # noqa: E402
from ppodash_b200._lib import GemmArgs  # noqa: E402
from microbench import time_kernel  # noqa: E402

DEV = "cuda:0"
B = 2048
# name, I, J, KK, a_kmajor, b_kmajor
SHAPES = [
    ("conv1.fwd", B * 400, 32, 192, 1, 1), ("conv2.fwd", B * 81, 64, 512, 1, 1), ("conv3.fwd", B * 49, 32, 576, 1, 1),
    ("fc.fwd", B, 512, 1568, 1, 1), ("gru_in.fwd", B, 1536, 528, 1, 1),
    ("conv3.dgrad", B * 49, 576, 32, 1, 0), ("conv2.dgrad", B * 81, 512, 64, 1, 0), ("fc.dgrad", B, 1568, 512, 1, 0),
    ("gru_in.dgrad", B, 512, 1536, 1, 0),
    ("conv1.wgrad", 32, 192, B * 400, 0, 0), ("conv2.wgrad", 64, 512, B * 81, 0, 0), ("conv3.wgrad", 32, 576, B * 49, 0, 0),
    ("fc.wgrad", 512, 1568, B, 0, 0), ("w_ih.wgrad", 1536, 528, B, 0, 0), ("w_hh.wgrad", 1024, 512, B, 0, 0),
]


def run(name, I, J, KK, ak, bk, mode):
    L = _lib.lib()
    A = torch.randn((I, KK) if ak else (KK, I), device=DEV)
    Bm = torch.randn((J, KK) if bk else (KK, J), device=DEV)
    C = torch.zeros(I, J, device=DEV)
    g = GemmArgs()
    g.A, g.lda, g.a_kmajor = A.data_ptr(), A.shape[1], ak
    g.B, g.ldb, g.b_kmajor = Bm.data_ptr(), Bm.shape[1], bk
    g.C, g.ldc, g.I, g.J, g.KK = C.data_ptr(), J, I, J, KK
    flags = {"tf32": 0, "tf32x3": 2}.get(mode, 0)
    if mode != "fp32" and not ak and not bk and I < J and I <= 64:
        g.A, g.lda, g.B, g.ldb = g.B, g.ldb, g.A, g.lda
        g.I, g.J = J, I
        flags |= 1
    st = _lib.stream_ptr()
    if mode == "fp32":
        ws = torch.empty(max(256, L.ppd_sgemm_workspace(I, J, KK)), dtype=torch.uint8, device=DEV)
        fn = lambda: _lib.check(L.ppd_sgemm(ctypes.byref(g), ws.data_ptr(), ws.numel(), st))
    else:
        ws = torch.empty(max(256, L.ppd_tc_gemm_workspace(g.I, g.J, KK)), dtype=torch.uint8, device=DEV)
        fn = lambda: _lib.check(L.ppd_tc_gemm(ctypes.byref(g), flags, ws.data_ptr(), ws.numel(), st))
        if mode == "tf32x3" and os.environ.get("PPD_BSPLIT") == "1" and not (flags & 1) and ak:
            # weights as B: pre-split once (forward / dgrad products)
            Bhi, Blo = torch.empty_like(Bm), torch.empty_like(Bm)
            _lib.check(L.ppd_split_tf32(Bm.data_ptr(), Bhi.data_ptr(), Blo.data_ptr(), Bm.numel(), st))
            g.B = Bhi.data_ptr()
            keep = (Bhi, Blo)
            fn = lambda: _lib.check(L.ppd_tc_gemm_bsplit(ctypes.byref(g), keep[1].data_ptr(), flags, ws.data_ptr(), ws.numel(), st))
    med, best = time_kernel(fn, iters=8, warmup=2)
    bytes_ = 4.0 * (I * KK + J * KK + I * J)
    return dict(gemm=name, mode=mode, ms=round(med, 4), gbs=round(bytes_ / med / 1e6, 1), tflops=round(2.0 * I * J * KK / med / 1e9, 2))


if __name__ == "__main__":
    if os.environ.get("PPD_TWO_CTAS") is not None:
        _lib.lib().ppd_tc_gemm_set_option(int(os.environ["PPD_TWO_CTAS"]))
    if os.environ.get("PPD_ATMEM") is not None:
        _lib.lib().ppd_tc_gemm_set_option(2 + int(os.environ["PPD_ATMEM"]))
    if os.environ.get("PPD_PERSIST") is not None:
        _lib.lib().ppd_tc_gemm_set_option(4 + int(os.environ["PPD_PERSIST"]))
    if os.environ.get("PPD_BN") is not None:
        _lib.lib().ppd_tc_gemm_set_option(int(os.environ["PPD_BN"]))
    modes = sys.argv[1:] or ["fp32", "tf32", "tf32x3"]
    tot = {m: 0.0 for m in modes}
    only = os.environ.get("PPD_SHAPES")
    for sh in SHAPES:
        if only and sh[0] not in only.split(","):
            continue
        for m in modes:
            r = run(*sh, m)
            tot[m] += r["ms"]
            print(json.dumps(r))
    print(json.dumps({"total_ms_per_minibatch": tot}))

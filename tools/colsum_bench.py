#!/usr/bin/env python
"""Timing of the seven bias-gradient column sums of one PPO-Dash minibatch (ppd_colsum_multi, two launches)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ppodash_b200 import _lib  # noqa: E402
from ppodash_b200._lib import ColsumSeg  # noqa: E402
from microbench import time_kernel  # noqa: E402

DEV = "cuda:0"
B = int(os.environ.get("PPD_B", 2048))


def main():
    L = _lib.lib()
    shapes = [(B * 400, 32), (B * 81, 64), (B * 49, 32), (B, 512), (B, 1536), (B, 1536), (B, 9)]      # conv1-3, fc, gi, gh, heads
    xs = [torch.randn(i, j, device=DEV) for i, j in shapes]
    outs = [torch.zeros(j, device=DEV) for _, j in shapes]
    segs = (ColsumSeg * len(shapes))()
    for s, (x, o) in enumerate(zip(xs, outs)):
        segs[s].X, segs[s].ld, segs[s].I, segs[s].J, segs[s].out, segs[s].accumulate = x.data_ptr(), x.shape[1], x.shape[0], x.shape[1], o.data_ptr(), 0
    ws = torch.empty(max(256, L.ppd_colsum_multi_workspace(segs, len(shapes))), dtype=torch.uint8, device=DEV)
    fn = lambda: _lib.check(L.ppd_colsum_multi(segs, len(shapes), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    med, best = time_kernel(fn, iters=10, warmup=3)
    err = max(float((o - x.sum(0)).abs().max() / x.shape[0] ** 0.5) for x, o in zip(xs, outs))
    mb = sum(x.numel() * 4 for x in xs) / 1e6
    print(json.dumps(dict(gemm="colsum_multi", ms=round(med, 4), mb=round(mb, 1), gbs=round(mb / med, 1), max_err_over_sqrt_rows=err)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Timings of the recurrent GRU kernels (forward / backward over T steps) at the PPO-Dash minibatch shape: PPD_E envs x PPD_T steps."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ppodash_b200 import _lib  # noqa: E402
from microbench import time_kernel  # noqa: E402

DEV = "cuda:0"


def main():
    L = _lib.lib()
    st = _lib.stream_ptr()
    if os.environ.get("PPD_GRU_MODE"):          # 3: interleaved-env kernels for every E, 4: one cluster per env for every E
        L.ppd_gru_set_mode(int(os.environ["PPD_GRU_MODE"]))
    if os.environ.get("PPD_GRU_CLUSTERS"):
        L.ppd_gru_set_mode(100 + int(os.environ["PPD_GRU_CLUSTERS"]))
    T, E, H = int(os.environ.get("PPD_T", 512)), int(os.environ.get("PPD_E", 4)), 512
    g = torch.Generator().manual_seed(0)
    d = lambda *s, scale=1.0: (scale * torch.randn(*s, generator=g)).to(DEV)
    gi, h0, whh, bhh, dhs = d(T * E, 3 * H), d(E, H), d(3 * H, H, scale=H ** -0.5), d(3 * H, scale=0.1), d(T * E, H)
    masks = (torch.rand(T * E, 1, generator=g) > 0.02).float().to(DEV)
    hs, hl = torch.zeros(T * E, H, device=DEV), torch.zeros(E, H, device=DEV)
    sr, sz, sn, sg = (torch.zeros(T * E, H, device=DEV) for _ in range(4))
    dgi, dghn, dh0 = torch.zeros(T * E, 3 * H, device=DEV), torch.zeros(T * E, H, device=DEV), torch.zeros(E, H, device=DEV)
    fwd = lambda: _lib.check(L.ppd_gru_forward(gi.data_ptr(), h0.data_ptr(), masks.data_ptr(), whh.data_ptr(), bhh.data_ptr(), T, E, H,
                                               hs.data_ptr(), hl.data_ptr(), sr.data_ptr(), sz.data_ptr(), sn.data_ptr(), sg.data_ptr(), st))
    bwd = lambda: _lib.check(L.ppd_gru_backward(dhs.data_ptr(), masks.data_ptr(), whh.data_ptr(), h0.data_ptr(), hs.data_ptr(),
                                                sr.data_ptr(), sz.data_ptr(), sn.data_ptr(), sg.data_ptr(), T, E, H, dgi.data_ptr(),
                                                dghn.data_ptr(), dh0.data_ptr(), st))
    tot = 0.0
    for name, fn in (("gru.fwd", fwd), ("gru.bwd", bwd)):
        med, best = time_kernel(fn, iters=8, warmup=2)
        tot += med
        print(json.dumps(dict(gemm=name, ms=round(med, 4), us_per_step=round(1e3 * med / T, 3), T=T, E=E)))
    print(json.dumps({"total_ms": round(tot, 4)}))


if __name__ == "__main__":
    main()

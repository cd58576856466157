#!/usr/bin/env python
"""Where do the implicit convolutions go wrong?  Runs a layer at several batch sizes against float64 and reports the failing output
rows by tile / CTA / position in the CTA's tile sequence, and whether repeats agree."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ppodash_b200 import _lib  # noqa: E402
from ppodash_b200._lib import ConvGeom  # noqa: E402

DEV = "cuda:0"
L = _lib.lib()


def split(w):
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    _lib.check(L.ppd_split_tf32(w.data_ptr(), hi.data_ptr(), lo.data_ptr(), w.numel(), _lib.stream_ptr()))
    return hi, lo


def fwd_case(B, H, C, k, s, Cout, nchw, reps=3):
    g0 = torch.Generator(device=DEV).manual_seed(B + H + C)
    OH = (H - k) // s + 1
    x = torch.randn(B, C, H, H, generator=g0, device=DEV)
    w = torch.randn(Cout, C, k, k, generator=g0, device=DEV) / np.sqrt(C * k * k)
    b = torch.randn(Cout, generator=g0, device=DEV)
    want = F.conv2d(x.double(), w.double(), b.double(), stride=s).permute(0, 2, 3, 1).reshape(B * OH * OH, Cout)
    geom = ConvGeom(B, H, H, C, k, k, s)
    if nchw:
        xk, wk = x, w.contiguous()
    else:
        xk, wk = x.permute(0, 2, 3, 1).contiguous(), w.permute(0, 2, 3, 1).contiguous()
    hi, lo = split(wk)
    outs = []
    for r in range(reps):
        out = torch.full((B * OH * OH, Cout), -7.0, device=DEV)
        fn = L.ppd_conv_fwd_nchw if nchw else L.ppd_conv_fwd_nhwc
        _lib.check(fn(xk.data_ptr(), ctypes.byref(geom), Cout, hi.data_ptr(), lo.data_ptr(), b.data_ptr(), 0, out.data_ptr(), _lib.stream_ptr()))
        torch.cuda.synchronize()
        outs.append(out)
    scale = float(want.abs().max())
    msg = []
    for r, out in enumerate(outs):
        err = (out.double() - want).abs().max(dim=1).values
        bad = torch.nonzero(err > 1e-5 * scale).flatten().cpu().numpy()
        nseg = 128 // OH
        if nchw:
            while nseg * 4 * H * 4 > 16384:
                nseg -= 1
        rows_tile = nseg * OH
        tiles = sorted(set((bad // rows_tile).tolist()))
        if len(bad) and nchw:
            # which k-blocks (channel c, ky half) explain the error?  diff ~ sum_kb alpha_kb * P_kb: alpha = -1 -> that k-block's A was zero / missing
            per_tile = {}
            for rr in bad.tolist():
                per_tile.setdefault(rr // rows_tile, []).append(rr % rows_tile)
            for tl, rows_ in list(per_tile.items())[:4]:
                rr = tl * rows_tile + rows_[len(rows_) // 2]
                bb, pix = rr // (OH * OH), rr % (OH * OH)
                oy, ox = pix // OH, pix % OH
                patch = x[bb, :, oy * s:oy * s + k, ox * s:ox * s + k].double()                 # [C, k, k]
                P = []
                for c in range(C):
                    for half in range(2):
                        P.append((w.double()[:, c, 4 * half:4 * half + 4, :] * patch[c, 4 * half:4 * half + 4, :]).sum(dim=(1, 2)))
                P = torch.stack(P, 1)                                                           # [Cout, nkb]
                diff = (out[rr].double() - want[rr])
                sol = torch.linalg.lstsq(P, diff.unsqueeze(1)).solution.flatten()
                resid = float((P @ sol - diff).abs().max())
                msg.append(f"      tile {tl} (cta {tl % 148}, seq {tl // 148}): {len(rows_)} bad rows in [{min(rows_)}, {max(rows_)}]; row {rr}: alpha_kb = "
                           f"{[round(float(a_), 2) for a_ in sol]} resid {resid:.2g}")
        msg.append(f"rep{r}: {len(bad)} bad rows, max err {float(err.max()):.3g} (scale {scale:.3g}); tiles {tiles[:12]}{'...' if len(tiles) > 12 else ''} "
                   f"cta {[t % 148 for t in tiles[:12]]} seq {[t // 148 for t in tiles[:12]]} in-tile rows {sorted(set((bad % rows_tile).tolist()))[:10]}")
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    print(f"fwd B={B} H={H} C={C} k={k} s={s} Cout={Cout} nchw={nchw}: repeats identical={same}")
    for m in msg:
        print("   ", m)


if __name__ == "__main__":
    if os.environ.get("DIAG_SHORT"):
        fwd_case(2048, 84, 3, 8, 4, 32, 1)
        fwd_case(2048, 84, 1, 8, 4, 32, 1)
        fwd_case(256, 84, 12, 8, 4, 32, 1, reps=2)
        sys.exit(0)
    for B in (40, 148, 256, 1024, 2048):
        fwd_case(B, 84, 3, 8, 4, 32, 1)
    fwd_case(1024, 84, 4, 8, 4, 32, 1)
    fwd_case(256, 84, 12, 8, 4, 32, 1)
    fwd_case(256, 84, 1, 8, 4, 32, 1)
    fwd_case(2048, 20, 32, 4, 2, 64, 0)
    fwd_case(2048, 9, 64, 3, 1, 32, 0)

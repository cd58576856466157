#!/usr/bin/env python
"""Kernel microbenchmarks (CUDA events on the launching stream, L2 flushed between iterations).
Prints one JSON object per kernel: achieved algorithmic GB/s against MEASURED_PEAKS.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppodash_b200 import _lib, synthetic  # noqa: E402
from ppodash_b200.storage import FusedAdvantages, RolloutStorage  # noqa: E402

DEV = "cuda:0"


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    _flush.add_(1)


def time_kernel(fn, iters=10, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def time_rotating(fns, rounds=3, warmup=1):
    """Average launch time over rounds x len(fns) back-to-back launches, each closure working on its OWN buffers whose total
    size exceeds L2 several times, so every launch reads cold data without a flush kernel in between.  One event pair around
    many launches: no 2-us event granularity and no launch latency of a lone ~35 us kernel in the figure."""
    for _ in range(warmup):
        for fn in fns:
            fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True)
    b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(rounds):
        for fn in fns:
            fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (rounds * len(fns))


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


def bench_gae(T, N, proper=False, sets=4):
    L = _lib.lib()
    s = _lib.stream_ptr()
    ws = torch.zeros(L.ppd_compute_returns_workspace(T, N), dtype=torch.uint8, device=DEV)
    fns, keep = [], []
    for _ in range(sets):
        r = torch.rand(T, N, 1, device=DEV)
        v = torch.randn(T + 1, N, 1, device=DEV)
        m = (torch.rand(T + 1, N, 1, device=DEV) > 0.002).float()
        b = torch.ones(T + 1, N, 1, device=DEV)
        ret = torch.zeros(T + 1, N, 1, device=DEV)
        nv = torch.randn(N, 1, device=DEV)
        keep.append((r, v, m, b, ret, nv))
        fns.append(lambda r=r, v=v, m=m, b=b, ret=ret, nv=nv: _lib.check(L.ppd_compute_returns(
            r.data_ptr(), v.data_ptr(), m.data_ptr(), b.data_ptr(), ret.data_ptr(), nv.data_ptr(), T, N, 0.99, 0.95, 1, int(proper),
            ws.data_ptr(), ws.numel(), s)))
    bytes_ = (20 if proper else 16) * T * N + 4 * N
    med_single, best = time_kernel(fns[0])                  # one launch between L2 flushes (event granularity ~2 us)
    set_bytes = sum(t.numel() * 4 for t in keep[0])
    if sets > 1 and (sets - 1) * set_bytes > (256 << 20):      # the other sets evict this one from L2 before it is read again
        ms = time_rotating(fns)
        timing = "%d back-to-back launches over %d rotating input sets (%d MB in total, several times L2)" % (3 * sets, sets, sets * set_bytes >> 20)
    else:
        ms, timing = med_single, "single launches, L2 flushed in between"
    return dict(kernel="returns_scan", T=T, N=N, proper=proper, ms=ms, ms_best=best, ms_single_flushed=med_single, timing=timing,
                steps_per_s=T * N / (ms * 1e-3), gbs=bytes_ / (ms * 1e-3) / 1e9)


def bench_gather(T, N, C, V, recurrent, nmb):
    H = 512 if recurrent else 1
    st = RolloutStorage(T, N, (C, 84, 84), [V], Discrete(8), H)
    st.to(DEV)
    st.obs.normal_()
    stats = torch.tensor([0.0, 1.0], device=DEV)
    row = C * 84 * 84 * 4 + V * 4 + 8 + 5 * 4
    out = {}

    def fn():
        gen = st.recurrent_generator(FusedAdvantages(stats), nmb) if recurrent else st.feed_forward_generator(FusedAdvantages(stats), nmb)
        for mb in gen:
            out["x"] = mb
    med, best = time_kernel(fn, iters=5, warmup=2)
    samples = (T * N // nmb) * nmb if not recurrent else T * (N // nmb) * nmb
    bytes_ = 2 * row * samples + (N // nmb) * nmb * H * 8 * (1 if recurrent else 0)
    return dict(kernel="gather_" + ("recurrent" if recurrent else "ff"), T=T, N=N, C=C, nmb=nmb, ms_epoch=med,
                ms_best=best, samples_per_s=samples / (med * 1e-3), gbs=bytes_ / (med * 1e-3) / 1e9)


def bench_gather_u8(T, N, C, V, nmb, nstack=1, with_mean=True):
    """Recurrent gather from the uint8 frame storage (normalise-on-gather): algorithmic bytes = C*HW read (+ V, scalars) and
    nstack*C*HW*4 written per sample."""
    import numpy as np
    mean = np.random.RandomState(0).rand(C, 84, 84) * 255 if with_mean else None
    st = RolloutStorage(T, N, (nstack * C, 84, 84), [V], Discrete(8), 512, obs_dtype=torch.uint8, frame_stack=nstack, obs_mean=mean,
                        obs_std=36.3 if with_mean else None)
    st.to(DEV)
    st._frames.random_(0, 256)
    st.obs_age.fill_(nstack - 1)
    stats = torch.tensor([0.0, 1.0], device=DEV)
    frame = C * 84 * 84
    row = frame + nstack * frame * 4 + 2 * (V * 4 + 8 + 5 * 4)
    out = {}

    def fn():
        for mb in st.recurrent_generator(FusedAdvantages(stats), nmb):
            out["x"] = mb
    med, best = time_kernel(fn, iters=5, warmup=2)
    samples = T * (N // nmb) * nmb
    bytes_ = row * samples + (N // nmb) * nmb * 512 * 8
    return dict(kernel="gather_recurrent_u8", T=T, N=N, C=C, nstack=nstack, nmb=nmb, ms_epoch=med, ms_best=best,
                samples_per_s=samples / (med * 1e-3), gbs=bytes_ / (med * 1e-3) / 1e9)


def bench_adam(n, sets=8):
    L = _lib.lib()
    ws = torch.empty(L.ppd_clip_adam_workspace(n), dtype=torch.uint8, device=DEV)
    s = _lib.stream_ptr()
    fns, keep = [], []
    for _ in range(sets):
        p = torch.randn(n, device=DEV); g = torch.randn(n, device=DEV) * 1e-3
        m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
        keep.append((p, g, m, v))
        fns.append(lambda p=p, g=g, m=m, v=v: _lib.check(L.ppd_clip_adam_step(
            p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 3, 1e-4, 0.9, 0.999, 1e-5, 0.5, None, None, None,
            ws.data_ptr(), ws.numel(), s)))
    med_single, best = time_kernel(fns[0])
    if (sets - 1) * 16 * n > (256 << 20):
        ms = time_rotating(fns)
        timing = "%d back-to-back launches over %d rotating parameter sets (%d MB in total, several times L2)" % (3 * sets, sets, sets * 16 * n >> 20)
    else:
        ms, timing = med_single, "single launches, L2 flushed in between"
    return dict(kernel="clip_adam", n=n, ms=ms, ms_best=best, ms_single_flushed=med_single, timing=timing, gbs=32 * n / (ms * 1e-3) / 1e9)


def main():
    peak, how = peak_hbm()
    only = sys.argv[1] if len(sys.argv) > 1 else "all"
    res = []
    if os.environ.get("PPD_RET_TUNE"):
        for part in os.environ["PPD_RET_TUNE"].split(";"):
            w, mb = (int(x) for x in part.split(","))
            _lib.lib().ppd_compute_returns_set_tuning(w, mb)
    if only in ("all", "returns"):
        for T, N in ((512, 32), (512, 1024), (2048, 4096)):
            res.append(bench_gae(T, N))
        res.append(bench_gae(2048, 4096, proper=True))
    if only in ("all", "gather"):
        res.append(bench_gather(512, 32, 3, 15, True, 8))
        res.append(bench_gather(128, 32, 4, 0, False, 4))
        res.append(bench_gather(512, 256, 3, 15, True, 8))
        res.append(bench_gather(512, 256, 3, 15, False, 8))
        res.append(bench_gather_u8(512, 256, 3, 15, 8))
        res.append(bench_gather_u8(512, 256, 3, 15, 8, with_mean=False))
        res.append(bench_gather_u8(512, 256, 3, 15, 8, nstack=4))
    if only in ("all", "adam"):
        res.append(bench_adam(2464393))
        res.append(bench_adam(1 << 26))
    for r in res:
        r["frac_of_hbm_peak"] = r["gbs"] / peak
        r["peak"] = how
        print(json.dumps(r))


if __name__ == "__main__":
    main()

"""GPU parity of the network building blocks (fp32 mode) through the C ABI: GEMM family, column
sums, conv lowering, GRU forward/backward.  Comparison: torch-CPU fp32 (the reference's own ops)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import policy as o_pol  # noqa: E402
from ppodash_b200 import _lib  # noqa: E402
from ppodash_b200._lib import GemmArgs  # noqa: E402

DEV = "cuda:0"


def sgemm(A, lda, a_k, Bm, ldb, b_k, C, ldc, I, J, KK, bias=None, mask=None, ldm=0, relu=0, acc=0):
    L = _lib.lib()
    g = GemmArgs()
    g.A, g.lda, g.a_kmajor = A.data_ptr(), lda, a_k
    g.B, g.ldb, g.b_kmajor = Bm.data_ptr(), ldb, b_k
    g.C, g.ldc = C.data_ptr(), ldc
    g.I, g.J, g.KK = I, J, KK
    g.bias = bias.data_ptr() if bias is not None else None
    g.mask = mask.data_ptr() if mask is not None else None
    g.ldm = ldm
    g.relu, g.accumulate = relu, acc
    ws = _lib.workspace(L.ppd_sgemm_workspace(I, J, KK), DEV, "gemm")
    _lib.check(L.ppd_sgemm(ctypes.byref(g), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))


@pytest.mark.parametrize("M,N,K", [(5, 3, 7), (64, 32, 192), (300, 64, 512), (2048, 512, 1568), (130, 1536, 527),
                                   (1000, 9, 512), (4096, 32, 576)])
def test_sgemm_forward_nt(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    ldx = K + (1 if K == 527 else 0)
    X = torch.randn(M, ldx, generator=g)
    W = torch.randn(N, K, generator=g) / np.sqrt(K)
    b = torch.randn(N, generator=g)
    want = F.relu(F.linear(X[:, :K], W, b))
    Xd, Wd, bd = X.to(DEV), W.to(DEV), b.to(DEV)
    C = torch.full((M, N + 3), -5.0, device=DEV)
    sgemm(Xd, ldx, 1, Wd, K, 1, C, N + 3, M, N, K, bias=bd, relu=1)
    np.testing.assert_allclose(C[:, :N].cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5)
    assert torch.all(C[:, N:] == -5.0)          # never writes outside J


@pytest.mark.parametrize("M,N,K", [(5, 3, 7), (300, 64, 512), (2048, 512, 1568), (130, 1536, 527), (1000, 9, 512), (2048, 9, 512), (777, 16, 36),
                                   (64, 17, 512)])
def test_sgemm_dgrad_nn_with_relu_mask(M, N, K):
    g = torch.Generator().manual_seed(M * 3 + N + K)
    dY = torch.randn(M, N, generator=g)
    W = torch.randn(N, K, generator=g) / np.sqrt(N)
    act = torch.randn(M, K, generator=g)
    want = (dY @ W) * (act > 0)
    dX = torch.zeros(M, K, device=DEV)
    sgemm(dY.to(DEV), N, 1, W.to(DEV), K, 0, dX, K, M, K, N, mask=act.to(DEV), ldm=K)
    np.testing.assert_allclose(dX.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("M,N,K", [(5, 3, 7), (81 * 64, 64, 512), (2048, 512, 1568), (2048, 1536, 527), (40000, 32, 192),
                                   (3000, 9, 512)])
def test_sgemm_wgrad_tn_splitk_accumulate(M, N, K):
    g = torch.Generator().manual_seed(M + 7 * N + K)
    dY = torch.randn(M, N, generator=g) / np.sqrt(M)
    X = torch.randn(M, K, generator=g)
    dW0 = torch.randn(N, K, generator=g)
    want = dW0 + (dY.double().t() @ X.double()).float()
    dW = dW0.to(DEV).clone()
    sgemm(dY.to(DEV), N, 0, X.to(DEV), K, 0, dW, K, N, K, M, acc=1)
    np.testing.assert_allclose(dW.cpu().numpy(), want.numpy(), rtol=1e-5, atol=2e-5)


def test_colsum():
    L = _lib.lib()
    for (I, J, ld) in [(1, 1, 1), (1000, 9, 12), (5000, 1536, 1536), (100000, 32, 32), (5184, 64, 64), (777, 32, 32), (300, 64, 70)]:
        X = torch.randn(I, ld)
        out0 = torch.randn(J)
        Xd = X.to(DEV)
        out = out0.to(DEV).clone()
        ws = _lib.workspace(L.ppd_colsum_workspace(I, J), DEV, "colsum")
        _lib.check(L.ppd_colsum(Xd.data_ptr(), ld, I, J, out.data_ptr(), 1, ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        np.testing.assert_allclose(out.cpu().numpy(), (out0 + X[:, :J].double().sum(0).float()).numpy(), rtol=1e-5, atol=1e-4)


def _unfold_nchw(x, k, s):
    # [B, C*k*k, L] with patch index (c,ky,kx) -> [B*L, C*k*k]
    u = F.unfold(x, k, stride=s)
    return u.transpose(1, 2).reshape(-1, u.shape[1])


def test_im2col_nchw_matches_unfold():
    L = _lib.lib()
    for (B, C, H, k, s) in [(3, 3, 84, 8, 4), (2, 1, 84, 8, 4), (2, 12, 84, 8, 4), (2, 2, 21, 3, 2)]:
        x = torch.randn(B, C, H, H)
        want = _unfold_nchw(x, k, s)
        K = C * k * k
        if K % 4:
            continue
        cols = torch.zeros(want.shape[0], K, device=DEV)
        xd = x.to(DEV)
        _lib.check(L.ppd_im2col_nchw(xd.data_ptr(), B, C, H, H, k, k, s, cols.data_ptr(), K, _lib.stream_ptr()))
        assert torch.equal(cols.cpu(), want)


def test_im2col_col2im_nhwc():
    L = _lib.lib()
    for (B, H, C, k, s) in [(3, 20, 32, 4, 2), (2, 9, 64, 3, 1), (1, 11, 8, 3, 2)]:
        x_nchw = torch.randn(B, C, H, H)
        x = x_nchw.permute(0, 2, 3, 1).contiguous()            # NHWC
        OH = (H - k) // s + 1
        u = F.unfold(x_nchw, k, stride=s)                       # [B, C*k*k, L], index (c,ky,kx)
        want = u.reshape(B, C, k, k, OH * OH).permute(0, 4, 2, 3, 1).reshape(B * OH * OH, k * k * C)   # (ky,kx,c)
        K = k * k * C
        cols = torch.zeros(B * OH * OH, K, device=DEV)
        xd = x.to(DEV)
        _lib.check(L.ppd_im2col_nhwc(xd.data_ptr(), B, H, H, C, k, k, s, cols.data_ptr(), K, _lib.stream_ptr()))
        assert torch.equal(cols.cpu(), want)
        # col2im = transpose of im2col (F.fold), fused with the ReLU mask
        dcols = torch.randn(B * OH * OH, K)
        act = torch.randn(B, H, H, C)
        d_u = dcols.reshape(B, OH * OH, k, k, C).permute(0, 4, 2, 3, 1).reshape(B, C * k * k, OH * OH)
        want_dx = F.fold(d_u, (H, H), k, stride=s).permute(0, 2, 3, 1) * (act > 0)
        dx = torch.zeros(B, H, H, C, device=DEV)
        dd, ad = dcols.to(DEV), act.to(DEV)
        _lib.check(L.ppd_col2im_nhwc(dd.data_ptr(), K, B, H, H, C, k, k, s, ad.data_ptr(), dx.data_ptr(), _lib.stream_ptr()))
        np.testing.assert_allclose(dx.cpu().numpy(), want_dx.numpy(), rtol=1e-6, atol=1e-6)


def test_batched_transpose():
    L = _lib.lib()
    x = torch.randn(37, 49, 32)
    xd = x.to(DEV)
    y = torch.zeros(37, 32, 49, device=DEV)
    _lib.check(L.ppd_batched_transpose(xd.data_ptr(), 37, 49, 32, y.data_ptr(), _lib.stream_ptr()))
    assert torch.equal(y.cpu(), x.transpose(1, 2).contiguous())


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("T,E,H,I", [(5, 3, 32, 35), (16, 4, 512, 527), (1, 32, 512, 527), (4, 40, 64, 64), (64, 4, 512, 527),
                                     (1, 2, 512, 527), (33, 8, 512, 512), (6, 40, 512, 527), (7, 21, 512, 527), (2, 11, 512, 527), (3, 37, 512, 527)])
def test_gru_forward_backward_vs_torch(T, E, H, I, mode):
    """mode 0: cluster/DSMEM kernels where they apply (W_hh in registers at H=512: any E, run as waves of clusters; generic H: E <= 8);
    mode 1: grid-cooperative kernels; mode 2: cluster kernels with W_hh in shared memory; mode 3: the persistent interleaved-env
    backward kernel (default from E = 9) for every E; mode 4: one cluster per env for every E."""
    _lib.lib().ppd_gru_set_mode(mode)
    try:
        _gru_case(T, E, H, I)
    finally:
        _lib.lib().ppd_gru_set_mode(0)


def _gru_case(T, E, H, I):
    g = torch.Generator().manual_seed(T * 100 + E)
    p = {"base.gru.weight_ih_l0": torch.randn(3 * H, I, generator=g) / np.sqrt(I),
         "base.gru.weight_hh_l0": torch.randn(3 * H, H, generator=g) / np.sqrt(H),
         "base.gru.bias_ih_l0": 0.1 * torch.randn(3 * H, generator=g),
         "base.gru.bias_hh_l0": 0.1 * torch.randn(3 * H, generator=g)}
    x = torch.randn(T * E, I, generator=g)
    h0 = 0.5 * torch.randn(E, H, generator=g)
    masks = (torch.rand(T * E, 1, generator=g) > 0.2).float()
    dhs = torch.randn(T * E, H, generator=g)
    # reference: segmented torch GRU + autograd
    w_hh = p["base.gru.weight_hh_l0"].clone().requires_grad_(True)
    pp = dict(p); pp["base.gru.weight_hh_l0"] = w_hh
    xr = x.clone().requires_grad_(True)
    h0r = h0.clone().requires_grad_(True)
    if T == 1:
        out, hl = o_pol.gru_cell_stepwise(pp, xr, h0r, masks)
    else:
        out, hl = o_pol.gru_with_resets(pp, xr, h0r, masks)
    (out * dhs).sum().backward()
    gi_ref = F.linear(x, p["base.gru.weight_ih_l0"], p["base.gru.bias_ih_l0"])

    L = _lib.lib()
    d = lambda t: t.to(DEV).contiguous()
    gi, h0d, md, whh, bhh, dhsd = d(gi_ref), d(h0), d(masks), d(p["base.gru.weight_hh_l0"]), d(p["base.gru.bias_hh_l0"]), d(dhs)
    hs = torch.zeros(T * E, H, device=DEV); hlast = torch.zeros(E, H, device=DEV)
    sr, sz, sn, sg = (torch.zeros(T * E, H, device=DEV) for _ in range(4))
    _lib.check(L.ppd_gru_forward(gi.data_ptr(), h0d.data_ptr(), md.data_ptr(), whh.data_ptr(), bhh.data_ptr(), T, E, H,
                                 hs.data_ptr(), hlast.data_ptr(), sr.data_ptr(), sz.data_ptr(), sn.data_ptr(), sg.data_ptr(),
                                 _lib.stream_ptr()))
    np.testing.assert_allclose(hs.cpu().numpy(), out.detach().numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(hlast.cpu().numpy(), hl.detach().numpy(), rtol=1e-5, atol=2e-6)
    dgi = torch.zeros(T * E, 3 * H, device=DEV); dghn = torch.zeros(T * E, H, device=DEV); dh0 = torch.zeros(E, H, device=DEV)
    _lib.check(L.ppd_gru_backward(dhsd.data_ptr(), md.data_ptr(), whh.data_ptr(), h0d.data_ptr(), hs.data_ptr(),
                                  sr.data_ptr(), sz.data_ptr(), sn.data_ptr(), sg.data_ptr(), T, E, H, dgi.data_ptr(),
                                  dghn.data_ptr(), dh0.data_ptr(), _lib.stream_ptr()))
    # dx = dgi @ W_ih ; dW_hh = dgh^T hm ; dh0
    dx = dgi.cpu() @ p["base.gru.weight_ih_l0"]
    np.testing.assert_allclose(dx.numpy(), xr.grad.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(dh0.cpu().numpy(), h0r.grad.numpy(), rtol=1e-4, atol=2e-5)
    hm = torch.zeros(T * E, H, device=DEV)
    _lib.check(L.ppd_gru_masked_prev(hs.data_ptr(), h0d.data_ptr(), md.data_ptr(), T, E, H, hm.data_ptr(), _lib.stream_ptr()))
    dgh = torch.cat([dgi[:, :2 * H], dghn], 1).cpu()
    dwhh = dgh.t() @ hm.cpu()
    np.testing.assert_allclose(dwhh.numpy(), w_hh.grad.numpy(), rtol=1e-4, atol=5e-5)


# --------------------------------------------------------------------------- tcgen05 TF32 GEMM
def tc_gemm(A, lda, a_k, Bm, ldb, b_k, C, ldc, I, J, KK, bias=None, mask=None, ldm=0, relu=0, acc=0, transpose_out=0,
            split3=0):
    L = _lib.lib()
    g = GemmArgs()
    g.A, g.lda, g.a_kmajor = A.data_ptr(), lda, a_k
    g.B, g.ldb, g.b_kmajor = Bm.data_ptr(), ldb, b_k
    g.C, g.ldc = C.data_ptr(), ldc
    g.I, g.J, g.KK = I, J, KK
    g.bias = bias.data_ptr() if bias is not None else None
    g.mask = mask.data_ptr() if mask is not None else None
    g.ldm = ldm
    g.relu, g.accumulate = relu, acc
    assert L.ppd_tc_gemm_supported(ctypes.byref(g)) == 1
    ws = _lib.workspace(L.ppd_tc_gemm_workspace(I, J, KK), DEV, "tcgemm")
    flags = (1 if transpose_out else 0) | (2 if split3 else 0)
    _lib.check(L.ppd_tc_gemm(ctypes.byref(g), flags, ws.data_ptr(), ws.numel(), _lib.stream_ptr()))


def _tf32_close(got, want, K, split3=0):
    """Stated tolerances relative to the largest entry: plain TF32 inputs carry a 10-bit mantissa ->
    4e-3; 3xTF32 (hi/lo split) restores fp32-level accuracy -> 1e-5 (the parity gate)."""
    scale = float(want.abs().max()) + 1e-6
    err = float((got - want).abs().max())
    assert err <= (1e-5 if split3 else 4e-3) * scale, (err, scale, K)


@pytest.mark.parametrize("split3", [0, 1])
@pytest.mark.parametrize("M,N,K", [(128, 32, 32), (4096, 32, 192), (1000, 64, 512), (777, 32, 576), (2048, 512, 1568),
                                   (2048, 1536, 528), (130, 100, 40)])
def test_tc_gemm_forward_k_major(M, N, K, split3):
    g = torch.Generator().manual_seed(M + N + K)
    X = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / np.sqrt(K)
    b = torch.randn(N, generator=g)
    want = torch.relu(X.double() @ W.double().t() + b.double()).float()
    C = torch.full((M, N + 4), -5.0, device=DEV)
    tc_gemm(X.to(DEV), K, 1, W.to(DEV), K, 1, C, N + 4, M, N, K, bias=b.to(DEV), relu=1, split3=split3)
    _tf32_close(C[:, :N].cpu(), want, K, split3)
    assert torch.all(C[:, N:] == -5.0)


@pytest.mark.parametrize("split3", [0, 1])
@pytest.mark.parametrize("M,N,K", [(128, 32, 64), (3000, 32, 576), (1000, 64, 512), (2048, 512, 1568), (2048, 1536, 512)])
def test_tc_gemm_dgrad_b_mn_major_with_mask(M, N, K, split3):
    g = torch.Generator().manual_seed(M * 3 + N + K)
    dY = torch.randn(M, N, generator=g)
    W = torch.randn(N, K, generator=g) / np.sqrt(N)
    act = torch.randn(M, K, generator=g)
    want = ((dY.double() @ W.double()) * (act > 0)).float()
    dX = torch.zeros(M, K, device=DEV)
    tc_gemm(dY.to(DEV), N, 1, W.to(DEV), K, 0, dX, K, M, K, N, mask=act.to(DEV), ldm=K, split3=split3)
    _tf32_close(dX.cpu(), want, N, split3)


@pytest.mark.parametrize("split3", [0, 1])
@pytest.mark.parametrize("M,N,K,swap", [(4096, 128, 192, False), (81 * 64, 64, 512, True), (2048, 512, 1568, False),
                                        (2048, 1536, 528, False), (40000, 32, 192, True), (5000, 32, 576, True)])
def test_tc_gemm_wgrad_both_mn_major_splitk(M, N, K, swap, split3):
    g = torch.Generator().manual_seed(M + 7 * N + K)
    dY = torch.randn(M, N, generator=g) / np.sqrt(M)
    X = torch.randn(M, K, generator=g)
    dW0 = torch.randn(N, K, generator=g)
    want = (dW0.double() + dY.double().t() @ X.double()).float()
    dW = dW0.to(DEV).clone()
    if swap:   # wide dimension (K) on the 128-row MMA axis, result stored transposed into dW[N, K]
        tc_gemm(X.to(DEV), K, 0, dY.to(DEV), N, 0, dW, K, K, N, M, acc=1, transpose_out=1, split3=split3)
    else:
        tc_gemm(dY.to(DEV), N, 0, X.to(DEV), K, 0, dW, K, N, K, M, acc=1, split3=split3)
    scale = float((want - dW0).abs().max()) + 1e-6
    assert float((dW.cpu() - want).abs().max()) <= (2e-5 if split3 else 4e-3) * scale


@pytest.mark.parametrize("split3", [0, 1])
@pytest.mark.parametrize("B,H,C,k,s,N", [(7, 9, 64, 3, 1, 32), (5, 20, 32, 4, 2, 64), (40, 20, 32, 4, 2, 64)])
def test_tc_gemm_fused_col2im(B, H, C, k, s, N, split3):
    """Conv dgrad with col2im in the GEMM epilogue (scatter-add) + ReLU mask == unfold-transpose reference."""
    from ppodash_b200._lib import ConvGeom
    L = _lib.lib()
    g0 = torch.Generator().manual_seed(B + H + C)
    OH = (H - k) // s + 1
    M, K = B * OH * OH, k * k * C
    dY = torch.randn(M, N, generator=g0)
    W = torch.randn(N, K, generator=g0) / np.sqrt(N)                       # columns ordered (ky,kx,c)
    act = torch.randn(B, H, H, C, generator=g0)
    dcols = dY.double() @ W.double()
    d_u = dcols.reshape(B, OH * OH, k, k, C).permute(0, 4, 2, 3, 1).reshape(B, C * k * k, OH * OH)
    want = (F.fold(d_u, (H, H), k, stride=s).permute(0, 2, 3, 1) * (act > 0)).float()
    dYd, Wd, actd = dY.to(DEV), W.to(DEV), act.to(DEV)
    dx = torch.zeros(B, H, H, C, device=DEV)
    g = GemmArgs()
    g.A, g.lda, g.a_kmajor = dYd.data_ptr(), N, 1
    g.B, g.ldb, g.b_kmajor = Wd.data_ptr(), K, 0
    g.C, g.ldc, g.I, g.J, g.KK = dx.data_ptr(), K, M, K, N
    geom = ConvGeom(B, H, H, C, k, k, s)
    _lib.check(L.ppd_tc_gemm_col2im(ctypes.byref(g), ctypes.byref(geom), 2 if split3 else 0, _lib.stream_ptr()))
    _lib.check(L.ppd_relu_mask(dx.data_ptr(), actd.data_ptr(), dx.numel(), _lib.stream_ptr()))
    scale = float(want.abs().max())
    assert float((dx.cpu() - want).abs().max()) <= (2e-5 if split3 else 4e-3) * scale


def test_colsum_multi():
    from ppodash_b200._lib import ColsumSeg
    L = _lib.lib()
    shapes = [(2048, 9, 9), (2048, 1536, 1536), (2048, 512, 512), (100352, 32, 32), (165888, 64, 64), (819200, 32, 32), (300, 64, 70)]
    xs = [torch.randn(I, ld, device=DEV) for I, J, ld in shapes]
    outs = [torch.randn(J, device=DEV) for I, J, ld in shapes]
    want = [o.cpu() * (i % 2) + x[:, :J].double().sum(0).float().cpu() for i, (x, o, (I, J, ld)) in enumerate(zip(xs, outs, shapes))]
    segs = (ColsumSeg * len(shapes))()
    for i, (sg, x, o, (I, J, ld)) in enumerate(zip(segs, xs, outs, shapes)):
        sg.X, sg.ld, sg.I, sg.J, sg.out, sg.accumulate = x.data_ptr(), ld, I, J, o.data_ptr(), i % 2
    ws = torch.empty(L.ppd_colsum_multi_workspace(segs, len(shapes)), dtype=torch.uint8, device=DEV)
    _lib.check(L.ppd_colsum_multi(segs, len(shapes), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    for o, w, (I, J, ld) in zip(outs, want, shapes):
        np.testing.assert_allclose(o.cpu().numpy(), w.numpy(), rtol=1e-4, atol=2e-3 * np.sqrt(I) / 30)


# --------------------------------------------------------------------------- implicit-GEMM NHWC convolutions (TMEM-A kernel)
def _split(w):
    L = _lib.lib()
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    _lib.check(L.ppd_split_tf32(w.data_ptr(), hi.data_ptr(), lo.data_ptr(), w.numel(), _lib.stream_ptr()))
    return hi, lo


def test_split_tf32_is_exact_and_tf32():
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(4096, generator=g) * torch.logspace(-20, 20, 4096)).to(DEV)
    hi, lo = _split(x)
    assert torch.all((hi.view(torch.int32) & 0x1FFF) == 0) and torch.all((lo.view(torch.int32) & 0x1FFF) == 0)   # 10-bit mantissas
    err = (x.double() - hi.double() - lo.double()).abs()
    assert torch.all(err <= x.abs().double() * 2.0 ** -21)          # hi + lo carries >= 21 mantissa bits of x


@pytest.fixture(params=[(1, 1), (0, 0), (1, 0)], ids=["resident", "per_kblock_boxes", "resident_input_streamed_weights"])
def conv_variant(request):
    """The implicit convolutions have two staging schemes each (ppd_tc_gemm_set_option 6/7: tile-resident raw input vs one im2col
    box per row and k-block; 8/9: resident vs streamed weight tiles); both must give the same results."""
    L = _lib.lib()
    res, bres = request.param
    L.ppd_tc_gemm_set_option(6 + res)
    L.ppd_tc_gemm_set_option(8 + bres)
    yield request.param
    L.ppd_tc_gemm_set_option(7)
    L.ppd_tc_gemm_set_option(9)


@pytest.mark.parametrize("B,H,C,k,s,Cout", [(7, 20, 32, 4, 2, 64), (5, 9, 64, 3, 1, 32), (300, 20, 32, 4, 2, 64), (333, 9, 64, 3, 1, 32),
                                            (3, 12, 32, 2, 2, 32)])
def test_conv_fwd_nhwc_implicit_gemm(B, H, C, k, s, Cout, conv_variant):
    """ppd_conv_fwd_nhwc == ReLU(conv2d + bias) (PKG/model.py:177-178), NHWC in / out, weights (o, ky, kx, c)."""
    from ppodash_b200._lib import ConvGeom
    L = _lib.lib()
    g0 = torch.Generator().manual_seed(B + H + C + k)
    x = torch.randn(B, H, H, C, generator=g0)
    w = torch.randn(Cout, k, k, C, generator=g0) / np.sqrt(k * k * C)
    b = torch.randn(Cout, generator=g0)
    want = torch.relu(F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), b.double(), stride=s)).permute(0, 2, 3, 1).float()
    OH = (H - k) // s + 1
    xd, wd, bd = x.to(DEV), w.to(DEV).contiguous(), b.to(DEV)
    hi, lo = _split(wd)
    out = torch.full((B * OH * OH + 3, Cout), -7.0, device=DEV)
    geom = ConvGeom(B, H, H, C, k, k, s)
    _lib.check(L.ppd_conv_fwd_nhwc(xd.data_ptr(), ctypes.byref(geom), Cout, hi.data_ptr(), lo.data_ptr(), bd.data_ptr(), 1,
                                   out.data_ptr(), _lib.stream_ptr()))
    got = out[:B * OH * OH].view(B, OH, OH, Cout).cpu()
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 1e-5 * scale        # 3xTF32: fp32-level accuracy
    assert torch.all(out[B * OH * OH:] == -7.0)


@pytest.mark.parametrize("B,H,C,k,s,Cout", [(7, 9, 64, 3, 1, 32), (5, 20, 32, 4, 2, 64), (300, 20, 32, 4, 2, 64), (333, 9, 64, 3, 1, 32),
                                            (3, 12, 32, 2, 2, 32)])
def test_conv_dgrad_nhwc_gather_form(B, H, C, k, s, Cout, conv_variant):
    """ppd_conv_dgrad_nhwc == ReLU'(act) * conv_transpose2d(dy): every dx element written exactly once."""
    from ppodash_b200._lib import ConvGeom
    L = _lib.lib()
    g0 = torch.Generator().manual_seed(B + H + C + 11)
    OH = (H - k) // s + 1
    dy = torch.randn(B, OH, OH, Cout, generator=g0)
    w = torch.randn(Cout, k, k, C, generator=g0) / np.sqrt(Cout)
    act = torch.randn(B, H, H, C, generator=g0)
    full = F.conv_transpose2d(dy.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), stride=s)     # [B, C, H', H']
    want = torch.zeros(B, C, H, H, dtype=torch.float64)
    want[:, :, :full.shape[2], :full.shape[3]] = full
    want = (want.permute(0, 2, 3, 1) * (act > 0)).float()
    dyd, wd, actd = dy.to(DEV), w.to(DEV).contiguous(), act.to(DEV)
    hi, lo = _split(wd)
    dx = torch.full((B, H, H, C), 9.0, device=DEV)
    geom = ConvGeom(B, H, H, C, k, k, s)
    _lib.check(L.ppd_conv_dgrad_nhwc(dyd.data_ptr(), ctypes.byref(geom), Cout, hi.data_ptr(), lo.data_ptr(), actd.data_ptr(),
                                     dx.data_ptr(), _lib.stream_ptr()))
    scale = float(want.abs().max())
    assert float((dx.cpu() - want).abs().max()) <= 1e-5 * scale


@pytest.mark.parametrize("B,H,C,k,s,Cout,nchw", [(7, 20, 32, 4, 2, 64, 0), (5, 9, 64, 3, 1, 32, 0), (300, 20, 32, 4, 2, 64, 0),
                                                 (333, 9, 64, 3, 1, 32, 0), (3, 84, 3, 8, 4, 32, 1), (40, 84, 3, 8, 4, 32, 1),
                                                 (2, 84, 1, 8, 4, 32, 1), (6, 84, 4, 8, 4, 32, 1)])
def test_conv_wgrad_implicit_gemm(B, H, C, k, s, Cout, nchw):
    """ppd_conv_wgrad == d(conv)/dW: NHWC inputs with (ky,kx,c) weights, NCHW observations with (c,ky,kx) weights."""
    from ppodash_b200._lib import ConvGeom
    L = _lib.lib()
    g0 = torch.Generator().manual_seed(B + H + C + 5)
    OH = (H - k) // s + 1
    M = B * OH * OH
    x = torch.randn(B, C, H, H, generator=g0)                              # reference layout NCHW
    dy = torch.randn(B, OH, OH, Cout, generator=g0) / np.sqrt(M)
    dW0 = torch.randn(Cout, C * k * k, generator=g0)
    cols = F.unfold(x.double(), k, stride=s)                               # [B, C*k*k (c,ky,kx), OH*OW]
    dw = torch.einsum("bpo,bkp->ok", dy.double().reshape(B, OH * OH, Cout), cols)           # [Cout, (c,ky,kx)]
    if not nchw:
        dw = dw.view(Cout, C, k, k).permute(0, 2, 3, 1).reshape(Cout, -1)  # -> (ky,kx,c)
        xd = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    else:
        xd = x.to(DEV)
    for acc in (0, 1):
        want = (dw + (dW0.double() if acc else 0)).float()
        dW = dW0.to(DEV).clone()
        geom = ConvGeom(B, H, H, C, k, k, s)
        ws = _lib.workspace(L.ppd_conv_wgrad_workspace(ctypes.byref(geom), Cout), DEV, "wgrad_test")
        _lib.check(L.ppd_conv_wgrad(xd.data_ptr(), ctypes.byref(geom), nchw, dy.to(DEV).data_ptr(), Cout, dW.data_ptr(), acc,
                                    ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        scale = float(dw.abs().max())
        assert float((dW.cpu() - want).abs().max()) <= 2e-5 * scale


@pytest.mark.parametrize("B,C,Cout", [(3, 3, 32), (40, 3, 32), (2, 1, 32), (5, 4, 32), (3, 12, 32)])
def test_conv_fwd_nchw_observations(B, C, Cout, conv_variant):
    """ppd_conv_fwd_nchw == ReLU(Conv2d(C, 32, 8, stride 4)(obs)) with NCHW observations and NHWC output."""
    from ppodash_b200._lib import ConvGeom
    L = _lib.lib()
    g0 = torch.Generator().manual_seed(B + C)
    x = torch.randn(B, C, 84, 84, generator=g0)
    w = torch.randn(Cout, C, 8, 8, generator=g0) / np.sqrt(64 * C)
    b = torch.randn(Cout, generator=g0)
    want = torch.relu(F.conv2d(x.double(), w.double(), b.double(), stride=4)).permute(0, 2, 3, 1).float()
    hi, lo = _split(w.to(DEV))
    out = torch.full((B * 400 + 3, Cout), -7.0, device=DEV)
    geom = ConvGeom(B, 84, 84, C, 8, 8, 4)
    _lib.check(L.ppd_conv_fwd_nchw(x.to(DEV).data_ptr(), ctypes.byref(geom), Cout, hi.data_ptr(), lo.data_ptr(), b.to(DEV).data_ptr(), 1,
                                   out.data_ptr(), _lib.stream_ptr()))
    got = out[:B * 400].view(B, 20, 20, Cout).cpu()
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())
    assert torch.all(out[B * 400:] == -7.0)

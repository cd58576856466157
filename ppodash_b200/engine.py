"""PolicyEngine: runs the actor-critic (PKG/model.py:54-199) forward and backward on the
libppodash_b200 kernels, over one flat fp32 parameter buffer.

Data layout in HBM
  * parameters / gradients / Adam moments: one flat fp32 buffer each, segments 256-byte aligned.
    The ``nn.Module`` parameters are *views* into the flat buffer (so ``state_dict`` keys, shapes and
    ``load_state_dict`` behave as in the reference) and ``p.grad`` are views into the flat gradient
    buffer.  Three segments use a kernel-friendly layout behind a strided view:
      - conv2 / conv3 weights are stored (o, ky, kx, c) -- the patch order of NHWC im2col;
      - gru.weight_ih rows are padded from hidden+V to a multiple of 4 floats (16-byte rows);
      - dist.linear / critic_linear weights and biases are adjacent, forming one [A+1, H] "heads"
        matrix so a single GEMM produces logits and value.
    The flat gradient buffer carries 4 extra floats at its tail for the three loss partial sums, so
    that one NCCL all-reduce per minibatch moves gradients and losses together.
  * activations: NHWC between the convolutions ( = row-major [B*OH*OW, C], what the GEMM writes);
    the conv3 output is transposed per sample to NCHW so the flatten order is the reference's.
  * rows of a minibatch are time-major (row = t*E + e), exactly as recurrent_generator yields them.
"""
import ctypes
import math
import os

import torch

from . import _lib
from ._lib import ColsumSeg, ConvGeom, GemmArgs, check, lib

ALIGN = 64  # floats (256 B)


def _round_up(x, m):
    return (x + m - 1) // m * m


def categorical_eval(z, num_actions, actions=None):
    """log-prob of `actions` (or of the arg-max action), per-row entropy, mode and probs
    (PKG/distributions.py:18-27) via ppd_categorical_eval."""
    B, ld = z.shape
    dev = z.device
    logp = torch.empty(B, device=dev)
    ent = torch.empty(B, device=dev)
    mode = torch.empty(B, dtype=torch.int64, device=dev)
    probs = torch.empty(B, num_actions, device=dev)
    a = None
    if actions is not None:
        a = actions.to(device=dev, dtype=torch.int64).contiguous()
    check(lib().ppd_categorical_eval(_lib.ptr(z, torch.float32), ld, num_actions, _lib.ptr(a), B, logp.data_ptr(),
                                     ent.data_ptr(), mode.data_ptr(), probs.data_ptr(), _lib.stream_ptr(dev)),
          "categorical_eval")
    return dict(logp=logp, entropy=ent, mode=mode, probs=probs)


class _Seg:
    __slots__ = ("name", "off", "numel", "shape", "view")

    def __init__(self, name, off, numel, shape, view):
        self.name, self.off, self.numel, self.shape, self.view = name, off, numel, shape, view


class PolicyEngine:
    CONV = ((8, 4), (4, 2), (3, 1))   # (kernel, stride) of base.main.{0,2,4}

    def __init__(self, policy):
        self.policy = policy
        base = policy.base
        self.C = base.num_inputs
        self.V = base.vector_obs_len
        self.H = base._hidden_size
        self.A = policy.num_actions
        self.recurrent = base.is_recurrent
        self.I = self.H + self.V
        self.Ipad = _round_up(self.I, 4)
        self.hw = policy.obs_shape[1]
        assert policy.obs_shape[1] == policy.obs_shape[2], "square observations expected"
        s = self.hw
        self.sp = []
        for k, st in self.CONV:
            s = (s - k) // st + 1
            self.sp.append(s)             # 20, 9, 7 for 84x84
        assert 32 * self.sp[2] * self.sp[2] == base.main[7].in_features, "obs size does not match the FC layer"
        self.flat_dim = 32 * self.sp[2] * self.sp[2]
        self.precision = "tf32x3"
        self.tc_min_work = 1 << 21        # I*J*KK below this: launch latency dominates, stay on the SIMT kernel
        self.chunk_rows = 2048            # rows of the minibatch lowered to im2col at a time
        self.device = None
        self.flat = None
        self._buffers = {}
        # weight-gradient GEMMs / bias sums on a second stream, concurrent with the dgrad chain
        self.overlap_wgrad = os.environ.get("PPD_OVERLAP_WGRAD", "1") != "0"
        # GRU time chunks on their own stream, concurrent with the trunk of other chunks.  Off by default: measured on
        # B200 at the PPO-Dash shape (T=512, E=4) it is SLOWER (265.6 vs 248.6 ms per update): a 16-CTA cluster with the
        # full register file per CTA cannot be co-scheduled while trunk kernels occupy the SMs, so the chunks serialise
        # anyway and only pay for smaller GEMMs and repeated W_hh loads.  The chunking itself is what bounds the im2col
        # scratch for large minibatches (E=128: 65 536 rows).
        self.overlap_gru = False
        self._sp = None                   # cached cudaStream_t of the current stream while train_minibatch runs (see `stream`)
        self._bound_params = None         # Parameter objects of the last successful bind check
        self._fast_bind = False           # set by algo.PPO for the duration of one update (see _is_bound)
        # Called on the current stream right before the training forward pass launches the GRU recurrence (16 SMs per env: 64 of
        # the 148 at 4 envs per minibatch).  algo.PPO uses it to gather the NEXT minibatch on a side stream into those idle SMs.
        self.on_gru_forward = None
        self.time_chunks = 4
        self.cols_budget = 6 << 30        # bytes of im2col matrices kept from forward for backward
        self._side = None
        self._gru_stream = None
        self._deferred = None
        # SMs left to an in-flight gradient all-reduce during the convolution backward (data parallel only); pair it with
        # NCCL_MAX_CTAS of the same value.  0 = persistent kernels keep all SMs.
        self.comm_ctas = int(os.environ.get("PPD_COMM_CTAS", "8"))

    # ------------------------------------------------------------------ parameters
    PRECISIONS = ("fp32", "tf32x3", "tf32")

    def set_precision(self, precision):
        """'fp32'  : SIMT fp32 GEMMs everywhere (reference arithmetic; slowest)
        'tf32x3': tcgen05 tensor cores with the 3xTF32 hi/lo split -> fp32-level accuracy (default)
        'tf32'  : tcgen05 tensor cores, single TF32 pass (~1e-3 relative per GEMM; fastest)
        GEMMs whose operands do not meet the TMA alignment rules (e.g. the 9-wide heads) and the
        sequential GRU recurrence always run in fp32 SIMT."""
        if precision not in self.PRECISIONS:
            raise ValueError(f"supported precision modes: {self.PRECISIONS}")
        self.precision = precision

    def _layout(self):
        base, pol = self.policy.base, self.policy
        H, A, C = self.H, self.A, self.C
        specs = []

        def add(name, param, numel, make_view):
            specs.append((name, param, numel, make_view))

        m = base.main
        add("conv1.w", m[0].weight, 32 * C * 64, lambda f: f.view(32, C, 8, 8))
        add("conv1.b", m[0].bias, 32, lambda f: f)
        add("conv2.w", m[2].weight, 64 * 512, lambda f: f.view(64, 4, 4, 32).permute(0, 3, 1, 2))
        add("conv2.b", m[2].bias, 64, lambda f: f)
        add("conv3.w", m[4].weight, 32 * 576, lambda f: f.view(32, 3, 3, 64).permute(0, 3, 1, 2))
        add("conv3.b", m[4].bias, 32, lambda f: f)
        add("fc.w", m[7].weight, H * self.flat_dim, lambda f: f.view(H, self.flat_dim))
        add("fc.b", m[7].bias, H, lambda f: f)
        if self.recurrent:
            g = base.gru
            add("gru.w_ih", g.weight_ih_l0, 3 * H * self.Ipad, lambda f: f.view(3 * H, self.Ipad)[:, :self.I])
            add("gru.w_hh", g.weight_hh_l0, 3 * H * H, lambda f: f.view(3 * H, H))
            add("gru.b_ih", g.bias_ih_l0, 3 * H, lambda f: f)
            add("gru.b_hh", g.bias_hh_l0, 3 * H, lambda f: f)
        # heads: dist rows then the critic row, adjacent -> one [A+1, H] matrix
        add("heads.w", (pol.dist.linear.weight, base.critic_linear.weight), (A + 1) * H,
            lambda f: (f.view(A + 1, H)[:A], f.view(A + 1, H)[A:]))
        add("heads.b", (pol.dist.linear.bias, base.critic_linear.bias), A + 1, lambda f: (f[:A], f[A:]))
        return specs

    def _is_bound(self):
        if self.flat is None:
            return False
        lo = self.flat.data_ptr()
        hi = lo + self.flat.numel() * 4
        # Inside train_minibatch / the optimiser step (64 x 2 calls per update) only the Parameter objects seen by the last full check
        # are looked at: walking the module tree costs ~30 us of host time per call.  Every entry from outside (PPO.update, act,
        # state_dict ...) still walks it, so a module that was moved or re-assigned between updates is noticed.
        params = self._bound_params if (self._sp is not None or self._fast_bind) and self._bound_params is not None else list(self.policy.parameters())
        for p in params:
            if p.device != self.flat.device or not (lo <= p.data_ptr() < hi):
                self._bound_params = None
                return False
        self._bound_params = params
        return True

    def bind(self):
        """(Re)build the flat buffers and make every parameter (and .grad) a view into them.
        Called lazily; repeats itself if the module was moved (.to()) or reloaded since."""
        if self._is_bound():
            return
        params = list(self.policy.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.PpdError("Policy must be on a CUDA device (call .to(device)); ppodash_b200 has no CPU fallback")
        specs = self._layout()
        off = 0
        offs = []
        for _, _, numel, _ in specs:
            offs.append(off)
            off = _round_up(off + numel, ALIGN)
        self.n_params = off                       # includes alignment padding (always zero)
        self.loss_off = off                       # 4 floats of loss partials at the gradient tail
        total = off + 4
        old_state = getattr(self, "adam_state", None)
        flat = torch.zeros(total, device=dev)
        grad = torch.zeros(total, device=dev)
        self.segs = {}
        with torch.no_grad():
            for (name, param, numel, make_view), o in zip(specs, offs):
                view = make_view(flat[o:o + numel])
                gview = make_view(grad[o:o + numel])
                if isinstance(param, tuple):
                    for p, v, gv in zip(param, view, gview):
                        v.copy_(p.data.to(dev))
                        p.data = v
                        p.grad = gv
                else:
                    view.copy_(param.data.to(dev))
                    param.data = view
                    param.grad = gview
                self.segs[name] = _Seg(name, o, numel, None, view)
        self.flat, self.flat_grad, self.device = flat, grad, dev
        # TF32 hi / residual copies of the parameters (ppd_split_tf32): weight operands of the 3xTF32 GEMMs are split
        # once per forward pass instead of once per tile inside the kernel
        self.flat_hi, self.flat_lo = torch.empty_like(flat), torch.empty_like(flat)
        if old_state is None or old_state["exp_avg"].numel() != total or old_state["exp_avg"].device != dev:
            self.adam_state = dict(exp_avg=torch.zeros(total, device=dev), exp_avg_sq=torch.zeros(total, device=dev), step=0)
        self._buffers = {}

    def seg(self, name, grad=False):
        s = self.segs[name]
        buf = self.flat_grad if grad else self.flat
        return buf[s.off:s.off + s.numel]

    # ------------------------------------------------------------------ streams
    @property
    def stream(self):
        """cudaStream_t of torch's current stream on this device (kernels are enqueued where torch ops go).  Inside train_minibatch the
        pointer is cached (`_sp`, kept in step with the engine's own stream switches): torch.cuda.current_stream() costs ~8 us of host
        time, there are ~60 launches per minibatch, and with 12 500 launches in a 120-ms update the Python side is what the device waits for."""
        sp = self._sp
        return sp if sp is not None else _lib.stream_ptr(self.device)

    def _on_stream(self, stream):
        """torch.cuda.stream(stream) that keeps the cached pointer in step."""
        eng = self

        class _Ctx:
            def __enter__(self_c):
                self_c.ctx = torch.cuda.stream(stream)
                self_c.ctx.__enter__()
                self_c.prev = eng._sp
                if eng._sp is not None:
                    eng._sp = stream.cuda_stream
                return self_c

            def __exit__(self_c, *exc):
                eng._sp = self_c.prev
                return self_c.ctx.__exit__(*exc)
        return _Ctx()

    def _ws(self, nbytes, tag):
        # one scratch buffer per (kind, stream): GEMMs running concurrently on two streams must not share split-K partials
        return _lib.workspace(nbytes, self.device, f"{tag}@{self.stream}")

    class _Side:
        """Context manager: run the enclosed launches on the side stream after everything enqueued so far on the
        main stream (fork); PolicyEngine._join() makes the main stream wait for the side stream."""

        def __init__(self, eng):
            self.eng = eng

        def __enter__(self):
            eng = self.eng
            if not eng.overlap_wgrad:
                return self
            if eng._side is None:
                eng._side = torch.cuda.Stream(device=eng.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(eng.device))
            eng._side.wait_event(ev)
            self.ctx = eng._on_stream(eng._side)
            self.ctx.__enter__()
            return self

        def __exit__(self, *exc):
            if self.eng.overlap_wgrad:
                self.ctx.__exit__(*exc)
            return False

    def _flush_colsums(self, from_off=None):
        """All bias gradients of the minibatch in two launches (ppd_colsum_multi).  from_off: flush only the sums whose output lies at
        or beyond that offset of the flat gradient buffer and keep collecting the others (early gradient bucket, see train_minibatch)."""
        if from_off is not None:
            base = self.flat_grad.data_ptr() + 4 * from_off
            todo = [c for c in (self._deferred or []) if c[4].data_ptr() >= base]
            self._deferred = [c for c in self._deferred if c[4].data_ptr() < base]
        else:
            todo, self._deferred = self._deferred, None
        if not todo:
            return
        L = lib()
        for i in range(0, len(todo), 8):
            part = todo[i:i + 8]
            segs = (ColsumSeg * len(part))()
            for sg, (X, ld, I, J, out, acc) in zip(segs, part):
                sg.X, sg.ld, sg.I, sg.J, sg.out, sg.accumulate = X.data_ptr(), ld, I, J, out.data_ptr(), acc
            ws = self._ws(L.ppd_colsum_multi_workspace(segs, len(part)), "colsum_multi")
            check(L.ppd_colsum_multi(segs, len(part), ws.data_ptr(), ws.numel(), self.stream), "colsum_multi")

    def _join(self):
        if self.overlap_wgrad and self._side is not None:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream(self.device).wait_event(ev)

    # ------------------------------------------------------------------ scratch
    class _Scratch:
        """Context manager: route `buf()` to a caller-owned dict.  A captured CUDA graph bakes the addresses of its scratch
        buffers in; giving the capture (and nothing else) its own dict keeps those buffers alive and un-shared for as long as
        the owner holds the dict, whatever sizes later eager / training calls ask the engine's own grow-only pool for."""

        def __init__(self, eng, bufs):
            self.eng, self.bufs = eng, bufs

        def __enter__(self):
            self.saved, self.eng._buffers = self.eng._buffers, self.bufs
            return self

        def __exit__(self, *exc):
            self.eng._buffers = self.saved
            return False

    def scratch(self, bufs):
        return self._Scratch(self, bufs)

    def signature(self):
        """Everything a captured forward pass depends on besides its inputs: the flat parameter buffers (re-created by bind()
        after .to() / load), their TF32 hi / lo copies and the precision mode."""
        self.bind()
        return (self.flat.data_ptr(), self.flat_hi.data_ptr(), self.flat_lo.data_ptr(), self.precision, str(self.device))

    def buf(self, name, *shape, dtype=torch.float32):
        n = 1
        for d in shape:
            n *= int(d)
        t = self._buffers.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(max(n, 1), dtype=dtype, device=self.device)
            self._buffers[name] = t
        return t[:n].view(*shape)

    # ------------------------------------------------------------------ kernel wrappers
    @property
    def _hilo(self):
        """Device addresses of the TF32 hi / residual copies of a parameter segment."""
        return (lambda name: self.flat_hi.data_ptr() + 4 * self.segs[name].off,
                lambda name: self.flat_lo.data_ptr() + 4 * self.segs[name].off)

    def split_params(self):
        """hi = TF32(w), lo = TF32(w - hi) for every parameter, in one launch over the flat buffer."""
        if self.precision != "tf32x3":
            return
        n = self.n_params
        check(lib().ppd_split_tf32(self.flat.data_ptr(), self.flat_hi.data_ptr(), self.flat_lo.data_ptr(), n, self.stream),
              "split_tf32")

    def _gemm(self, A, lda, a_k, B, ldb, b_k, C, ldc, I, J, KK, bias=None, mask=None, ldm=0, relu=0, acc=0, b_param=False):
        """b_param: B is a view of the flat parameter buffer (its pre-split hi / lo copies are used in tf32x3 mode)."""
        g = GemmArgs()
        g.A, g.lda, g.a_kmajor = A.data_ptr(), lda, a_k
        g.B, g.ldb, g.b_kmajor = B.data_ptr(), ldb, b_k
        g.C, g.ldc = C.data_ptr(), ldc
        g.I, g.J, g.KK = I, J, KK
        g.bias = bias.data_ptr() if bias is not None else None
        g.mask = mask.data_ptr() if mask is not None else None
        g.ldm = ldm
        g.relu, g.accumulate = relu, acc
        L = lib()
        if self.precision != "fp32" and I * J * KK >= self.tc_min_work and L.ppd_tc_gemm_supported(ctypes.byref(g)):
            flags = 2 if self.precision == "tf32x3" else 0
            if not a_k and not b_k and I < J and I <= 64 and bias is None and mask is None:
                # weight gradient with few output rows: put the wide dimension on the 128-row MMA axis
                # and store the tile transposed
                g.A, g.lda, g.B, g.ldb = g.B, g.ldb, g.A, g.lda
                g.I, g.J = J, I
                flags |= 1
            ws = self._ws(L.ppd_tc_gemm_workspace(g.I, g.J, KK), "tcgemm")
            if b_param and self.precision == "tf32x3" and not (flags & 1):
                off = B.data_ptr() - self.flat.data_ptr()
                g.B = self.flat_hi.data_ptr() + off
                check(L.ppd_tc_gemm_bsplit(ctypes.byref(g), self.flat_lo.data_ptr() + off, flags, ws.data_ptr(), ws.numel(),
                                           self.stream), "tc_gemm_bsplit")
                return
            check(L.ppd_tc_gemm(ctypes.byref(g), flags, ws.data_ptr(), ws.numel(), self.stream), "tc_gemm")
            return
        ws = self._ws(L.ppd_sgemm_workspace(I, J, KK), "gemm")
        check(L.ppd_sgemm(ctypes.byref(g), ws.data_ptr(), ws.numel(), self.stream), "sgemm")

    def _dgrad_col2im(self, dY, N, W, K, n, Hin, Cin, k, stride, act, dx):
        """dx[n,Hin,Hin,Cin] = ReLU'(act) * col2im(dY[M,N] @ W[N,K]).  Tensor-core modes fuse col2im into the GEMM
        epilogue (scatter-add, no dcols matrix in HBM); fp32 mode uses the deterministic SIMT GEMM + gather col2im."""
        L = lib()
        OH = (Hin - k) // stride + 1
        M = n * OH * OH
        if self.precision == "fp32":
            dcols = self.buf("dcols", M, K)
            self._gemm(dY, N, 1, W, K, 0, dcols, K, M, K, N)
            check(L.ppd_col2im_nhwc(dcols.data_ptr(), K, n, Hin, Hin, Cin, k, k, stride, act.data_ptr(), dx.data_ptr(),
                                    self.stream), "col2im")
            return
        g = GemmArgs()
        g.A, g.lda, g.a_kmajor = dY.data_ptr(), N, 1
        g.B, g.ldb, g.b_kmajor = W.data_ptr(), K, 0
        g.C, g.ldc = dx.data_ptr(), K
        g.I, g.J, g.KK = M, K, N
        geom = ConvGeom(n, Hin, Hin, Cin, k, k, stride)
        cnt = n * Hin * Hin * Cin
        dx.view(-1)[:cnt].zero_()
        check(L.ppd_tc_gemm_col2im(ctypes.byref(g), ctypes.byref(geom), 2 if self.precision == "tf32x3" else 0, self.stream),
              "tc_gemm_col2im")
        check(L.ppd_relu_mask(dx.data_ptr(), act.data_ptr(), cnt, self.stream), "relu_mask")

    def _colsum(self, X, ld, I, J, out, acc=0):
        if self._deferred is not None:          # collected and reduced together at the end of the minibatch
            self._deferred.append((X, ld, I, J, out, acc))
            return
        L = lib()
        ws = self._ws(L.ppd_colsum_workspace(I, J), "colsum")
        check(L.ppd_colsum(X.data_ptr(), ld, I, J, out.data_ptr(), acc, ws.data_ptr(), ws.numel(), self.stream), "colsum")

    # ------------------------------------------------------------------ row chunks
    def _chunks(self, B, E=None):
        """Row ranges [r0, r1) processed one after the other.  Recurrent minibatches are cut along TIME (rows are
        time-major), so that the GRU of one time chunk runs on its own stream concurrently with the conv trunk
        of the next one (forward) / the previous one (backward); every chunk is also at most chunk_rows rows so
        that the im2col scratch stays bounded."""
        implicit = self.precision == "tf32x3"         # implicit-GEMM convolutions: no im2col scratch to bound
        if E:
            T = B // E
            nc = self.time_chunks if (self.overlap_gru and T >= 8 * self.time_chunks) else 1
            if not implicit:
                nc = max(nc, -(-B // self.chunk_rows))
            nc = min(nc, T)
            tl = -(-T // nc)
            return [(t0 * E, min(T, t0 + tl) * E) for t0 in range(0, T, tl)]
        ch = B if implicit else min(self.chunk_rows, B)
        return [(r0, min(B, r0 + ch)) for r0 in range(0, B, ch)]

    def _cols(self, B, chunks, keep):
        """im2col scratch.  If the matrices of the whole minibatch fit the budget they are kept for backward
        (row offset = chunk start); otherwise one chunk-sized set is reused and backward recomputes them."""
        if self.precision == "tf32x3":
            return (None, None, None, True)
        C = self.C
        s1, s2, s3 = self.sp
        K1, K2, K3 = C * 64, 512, 576
        per_row = 4 * (s1 * s1 * K1 + s2 * s2 * K2 + s3 * s3 * K3)
        full = keep and B * per_row <= self.cols_budget
        rows = B if full else max(r1 - r0 for r0, r1 in chunks)
        return (self.buf("cols1", rows * s1 * s1, K1), self.buf("cols2", rows * s2 * s2, K2),
                self.buf("cols3", rows * s3 * s3, K3), full)

    # ------------------------------------------------------------------ trunk
    def _trunk_forward_rows(self, obs, r0, r1, cols, off, feat, ldf):
        """rows [r0, r1): obs -> conv stack -> flatten -> feat[r0:r1, :H] = ReLU(FC(.)) (model.py:176-180,194)."""
        L = lib()
        C, H, hw = self.C, self.H, self.hw
        s1, s2, s3 = self.sp
        K1, K2, K3 = C * 64, 512, 576
        n = r1 - r0
        B = obs.shape[0]
        a1 = self.buf("a1", B, s1 * s1 * 32)
        a2 = self.buf("a2", B, s2 * s2 * 64)
        a3 = self.buf("a3", B, s3 * s3 * 32)
        a3t = self.buf("a3t", B, self.flat_dim)
        st = self.stream
        if self.precision == "tf32x3":
            # implicit-GEMM convolutions on the TMEM-A kernel: patches come straight from obs / a1 / a2 through TMA
            hi, lo = self._hilo
            check(L.ppd_conv_fwd_nchw(obs[r0:].data_ptr(), ctypes.byref(ConvGeom(n, hw, hw, C, 8, 8, 4)), 32, hi("conv1.w"), lo("conv1.w"),
                                      self.seg("conv1.b").data_ptr(), 1, a1[r0:].data_ptr(), st), "conv1.fwd")
            check(L.ppd_conv_fwd_nhwc(a1[r0:].data_ptr(), ctypes.byref(ConvGeom(n, s1, s1, 32, 4, 4, 2)), 64, hi("conv2.w"), lo("conv2.w"),
                                      self.seg("conv2.b").data_ptr(), 1, a2[r0:].data_ptr(), st), "conv2.fwd")
            check(L.ppd_conv_fwd_nhwc(a2[r0:].data_ptr(), ctypes.byref(ConvGeom(n, s2, s2, 64, 3, 3, 1)), 32, hi("conv3.w"), lo("conv3.w"),
                                      self.seg("conv3.b").data_ptr(), 1, a3[r0:].data_ptr(), st), "conv3.fwd")
            check(L.ppd_batched_transpose(a3[r0:].data_ptr(), n, s3 * s3, 32, a3t[r0:].data_ptr(), st), "transpose")
            self._gemm(a3t[r0:], self.flat_dim, 1, self.seg("fc.w"), self.flat_dim, 1, feat[r0:], ldf, n, H, self.flat_dim,
                       bias=self.seg("fc.b"), relu=1, b_param=True)
            return
        c1, c2, c3 = cols[0][off * s1 * s1:], cols[1][off * s2 * s2:], cols[2][off * s3 * s3:]
        check(L.ppd_im2col_nchw(obs[r0:].data_ptr(), n, C, hw, hw, 8, 8, 4, c1.data_ptr(), K1, st), "im2col1")
        self._gemm(c1, K1, 1, self.seg("conv1.w"), K1, 1, a1[r0:], 32, n * s1 * s1, 32, K1, bias=self.seg("conv1.b"), relu=1, b_param=True)
        check(L.ppd_im2col_nhwc(a1[r0:].data_ptr(), n, s1, s1, 32, 4, 4, 2, c2.data_ptr(), K2, st), "im2col2")
        self._gemm(c2, K2, 1, self.seg("conv2.w"), K2, 1, a2[r0:], 64, n * s2 * s2, 64, K2, bias=self.seg("conv2.b"), relu=1, b_param=True)
        check(L.ppd_im2col_nhwc(a2[r0:].data_ptr(), n, s2, s2, 64, 3, 3, 1, c3.data_ptr(), K3, st), "im2col3")
        self._gemm(c3, K3, 1, self.seg("conv3.w"), K3, 1, a3[r0:], 32, n * s3 * s3, 32, K3, bias=self.seg("conv3.b"), relu=1, b_param=True)
        check(L.ppd_batched_transpose(a3[r0:].data_ptr(), n, s3 * s3, 32, a3t[r0:].data_ptr(), st), "transpose")
        self._gemm(a3t[r0:], self.flat_dim, 1, self.seg("fc.w"), self.flat_dim, 1, feat[r0:], ldf, n, H, self.flat_dim,
                   bias=self.seg("fc.b"), relu=1, b_param=True)

    def _trunk_backward_rows(self, obs, r0, r1, cols, off, cols_valid, dfeat, ldd, acc, after_fc=None):
        """rows [r0, r1): dfeat (already masked by feat > 0) -> FC / conv gradients, written to (acc=0, first chunk) or
        added to (acc=1) the flat gradient buffer.  Weight / bias gradients run on the side stream."""
        L = lib()
        C, H, hw = self.C, self.H, self.hw
        s1, s2, s3 = self.sp
        K1, K2, K3 = C * 64, 512, 576
        n = r1 - r0
        B = obs.shape[0]
        fd = self.flat_dim
        a1 = self.buf("a1", B, s1 * s1 * 32)
        a2 = self.buf("a2", B, s2 * s2 * 64)
        a3t = self.buf("a3t", B, fd)
        da3t = self.buf("da3t", B, fd)
        dy3 = self.buf("dy3", B, s3 * s3 * 32)                   # NHWC = [B*49, 32]
        dy2 = self.buf("dy2", B, s2 * s2 * 64)
        dy1 = self.buf("dy1", B, s1 * s1 * 32)
        M3, M2, M1 = n * s3 * s3, n * s2 * s2, n * s1 * s1
        if self.precision == "tf32x3":
            hi, lo = self._hilo
            with self._Side(self):
                self._gemm(dfeat[r0:], ldd, 0, a3t[r0:], fd, 0, self.seg("fc.w", True), fd, H, fd, n, acc=acc)
                self._colsum(dfeat[r0:], ldd, n, H, self.seg("fc.b", True), acc)
                if after_fc is not None:
                    after_fc()          # every gradient from fc.w to the end of the flat buffer is enqueued on the side stream
            self._gemm(dfeat[r0:], ldd, 1, self.seg("fc.w"), fd, 0, da3t[r0:], fd, n, fd, H, mask=a3t[r0:], ldm=fd, b_param=True)
            check(L.ppd_batched_transpose(da3t[r0:].data_ptr(), n, 32, s3 * s3, dy3[r0:].data_ptr(), self.stream), "transpose")
            g3, g2, g1 = ConvGeom(n, s2, s2, 64, 3, 3, 1), ConvGeom(n, s1, s1, 32, 4, 4, 2), ConvGeom(n, hw, hw, C, 8, 8, 4)

            def wgrad(x, geom, nchw, dy, cout, name):
                ws = self._ws(L.ppd_conv_wgrad_workspace(ctypes.byref(geom), cout), "convwgrad")
                check(L.ppd_conv_wgrad(x.data_ptr(), ctypes.byref(geom), nchw, dy.data_ptr(), cout, self.seg(name, True).data_ptr(), acc,
                                       ws.data_ptr(), ws.numel(), self.stream), name + ".wgrad")
            with self._Side(self):
                wgrad(a2[r0:], g3, 0, dy3[r0:], 32, "conv3.w")
                self._colsum(dy3[r0:], 32, M3, 32, self.seg("conv3.b", True), acc)
            check(L.ppd_conv_dgrad_nhwc(dy3[r0:].data_ptr(), ctypes.byref(g3), 32, hi("conv3.w"), lo("conv3.w"), a2[r0:].data_ptr(),
                                        dy2[r0:].data_ptr(), self.stream), "conv3.dgrad")
            with self._Side(self):
                wgrad(a1[r0:], g2, 0, dy2[r0:], 64, "conv2.w")
                self._colsum(dy2[r0:], 64, M2, 64, self.seg("conv2.b", True), acc)
            check(L.ppd_conv_dgrad_nhwc(dy2[r0:].data_ptr(), ctypes.byref(g2), 64, hi("conv2.w"), lo("conv2.w"), a1[r0:].data_ptr(),
                                        dy1[r0:].data_ptr(), self.stream), "conv2.dgrad")
            wgrad(obs[r0:], g1, 1, dy1[r0:], 32, "conv1.w")
            self._colsum(dy1[r0:], 32, M1, 32, self.seg("conv1.b", True), acc)
            return
        c1, c2, c3 = cols[0][off * s1 * s1:], cols[1][off * s2 * s2:], cols[2][off * s3 * s3:]
        with self._Side(self):
            self._gemm(dfeat[r0:], ldd, 0, a3t[r0:], fd, 0, self.seg("fc.w", True), fd, H, fd, n, acc=acc)
            self._colsum(dfeat[r0:], ldd, n, H, self.seg("fc.b", True), acc)
        self._gemm(dfeat[r0:], ldd, 1, self.seg("fc.w"), fd, 0, da3t[r0:], fd, n, fd, H, mask=a3t[r0:], ldm=fd, b_param=True)
        check(L.ppd_batched_transpose(da3t[r0:].data_ptr(), n, 32, s3 * s3, dy3[r0:].data_ptr(), self.stream), "transpose")
        # conv3
        if not cols_valid:
            check(L.ppd_im2col_nhwc(a2[r0:].data_ptr(), n, s2, s2, 64, 3, 3, 1, c3.data_ptr(), K3, self.stream), "im2col3")
        with self._Side(self):
            self._gemm(dy3[r0:], 32, 0, c3, K3, 0, self.seg("conv3.w", True), K3, 32, K3, M3, acc=acc)
            self._colsum(dy3[r0:], 32, M3, 32, self.seg("conv3.b", True), acc)
        self._dgrad_col2im(dy3[r0:], 32, self.seg("conv3.w"), K3, n, s2, 64, 3, 1, a2[r0:], dy2[r0:])
        # conv2
        if not cols_valid:
            check(L.ppd_im2col_nhwc(a1[r0:].data_ptr(), n, s1, s1, 32, 4, 4, 2, c2.data_ptr(), K2, self.stream), "im2col2")
        with self._Side(self):
            self._gemm(dy2[r0:], 64, 0, c2, K2, 0, self.seg("conv2.w", True), K2, 64, K2, M2, acc=acc)
            self._colsum(dy2[r0:], 64, M2, 64, self.seg("conv2.b", True), acc)
        self._dgrad_col2im(dy2[r0:], 64, self.seg("conv2.w"), K2, n, s1, 32, 4, 2, a1[r0:], dy1[r0:])
        # conv1 (no input gradient needed)
        if not cols_valid:
            check(L.ppd_im2col_nchw(obs[r0:].data_ptr(), n, C, hw, hw, 8, 8, 4, c1.data_ptr(), K1, self.stream), "im2col1")
        self._gemm(dy1[r0:], 32, 0, c1, K1, 0, self.seg("conv1.w", True), K1, 32, K1, M1, acc=acc)
        self._colsum(dy1[r0:], 32, M1, 32, self.seg("conv1.b", True), acc)
        if not cols_valid:
            self._join()              # the side stream still reads the chunk-sized im2col scratch

    # ------------------------------------------------------------------ forward
    def _prep(self, visual, vector, rnn_hxs, masks):
        self.bind()
        dev = self.device
        obs = visual.to(device=dev, dtype=torch.float32).contiguous()
        B = obs.shape[0]
        if tuple(obs.shape[1:]) != (self.C, self.hw, self.hw):
            raise ValueError(f"expected observations of shape [B,{self.C},{self.hw},{self.hw}], got {tuple(obs.shape)}")
        vobs = vector.to(device=dev, dtype=torch.float32).reshape(B, -1).contiguous() if vector is not None else None
        if self.V and (vobs is None or vobs.shape[1] != self.V):
            raise ValueError(f"expected vector observations of shape [B,{self.V}]")
        h0 = rnn_hxs.to(device=dev, dtype=torch.float32).contiguous()
        m = masks.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        return obs, vobs, h0, m, B

    def _gru_stream_ctx(self, nchunks):
        """Stream the GRU chunks run on: a dedicated stream when the minibatch is cut along time, else the current one."""
        if nchunks > 1 and self.overlap_gru:
            if self._gru_stream is None:
                self._gru_stream = torch.cuda.Stream(device=self.device)
            return self._gru_stream
        return None

    def forward(self, visual, vector, rnn_hxs, masks, keep=False, xcat_prefilled=None):
        """CNNBase.forward + both heads (model.py:192-199, distributions.py:66-68).
        Returns dict(value [B,1], z [B,A+1] (logits | value), rnn_hxs, feats)."""
        obs, vobs, h0, m, B = self._prep(visual, vector, rnn_hxs, masks)
        self.split_params()
        L = lib()
        H, A = self.H, self.A
        tag = "t_" if keep else "i_"
        z = self.buf(tag + "z", B, A + 1) if keep else torch.empty(B, A + 1, device=self.device)
        main = torch.cuda.current_stream(self.device)
        if self.recurrent:
            E = h0.shape[0]
            if B % E != 0:
                raise ValueError("rows must be a multiple of the number of hidden-state rows (T*E, E)")
            T = B // E
            Ipad = self.Ipad
            chunks = self._chunks(B, E)
            cols = self._cols(B, chunks, keep)
            xcat = xcat_prefilled if xcat_prefilled is not None else self.buf(tag + "xcat", B, Ipad)
            if xcat_prefilled is None and Ipad != H:
                xcat[:, H:].zero_()
                if self.V:
                    xcat[:, H:H + self.V].copy_(vobs)               # torch.cat((x, vector), 1), model.py:195
            gi = self.buf(tag + "gi", B, 3 * H)
            hs = self.buf(tag + "hs", B, H) if keep else torch.empty(B, H, device=self.device)
            hl = torch.empty(E, H, device=self.device)
            sv = [self.buf("t_s%d" % i, B, H) for i in range(4)] if keep else None
            gstream = self._gru_stream_ctx(len(chunks))
            w_hh, b_hh = self.seg("gru.w_hh"), self.seg("gru.b_hh")
            for ci, (r0, r1) in enumerate(chunks):
                n = r1 - r0
                self._trunk_forward_rows(obs, r0, r1, cols, r0 if cols[3] else 0, xcat, Ipad)
                self._gemm(xcat[r0:], Ipad, 1, self.seg("gru.w_ih"), Ipad, 1, gi[r0:], 3 * H, n, 3 * H, Ipad,
                           bias=self.seg("gru.b_ih"), b_param=True)
                last = ci == len(chunks) - 1
                h_in = h0 if ci == 0 else hs[r0 - E:r0]
                svp = [s_[r0:].data_ptr() for s_ in sv] if keep else [None] * 4

                def run_gru():
                    check(L.ppd_gru_forward(gi[r0:].data_ptr(), h_in.data_ptr(), m[r0:].data_ptr(), w_hh.data_ptr(),
                                            b_hh.data_ptr(), n // E, E, H, hs[r0:].data_ptr(),
                                            hl.data_ptr() if last else None, *svp, self.stream), "gru_forward")
                if gstream is None:
                    if keep and self.on_gru_forward is not None:
                        self.on_gru_forward()          # the recurrence leaves 84 SMs idle: the caller queues independent work behind this point
                    run_gru()
                else:
                    ev = torch.cuda.Event()
                    ev.record(main)
                    gstream.wait_event(ev)
                    with self._on_stream(gstream):
                        run_gru()
            if gstream is not None:
                ev = torch.cuda.Event()
                ev.record(gstream)
                main.wait_event(ev)
            feats, ldf, rnn_out = hs, H, hl
            self._saved = dict(obs=obs, xcat=xcat, gi=gi, hs=hs, h0=h0, m=m, T=T, E=E, B=B, chunks=chunks, cols=cols) if keep else None
        else:
            chunks = self._chunks(B)
            cols = self._cols(B, chunks, keep)
            feat = self.buf(tag + "feat", B, H) if keep else torch.empty(B, H, device=self.device)
            for r0, r1 in chunks:
                self._trunk_forward_rows(obs, r0, r1, cols, r0 if cols[3] else 0, feat, H)
            feats, ldf, rnn_out = feat, H, rnn_hxs
            self._saved = dict(obs=obs, feat=feat, B=B, chunks=chunks, cols=cols) if keep else None
        self._gemm(feats, ldf, 1, self.seg("heads.w"), H, 1, z, A + 1, B, A + 1, H, bias=self.seg("heads.b"))
        if keep:
            self._saved["z"] = z
        return dict(value=z[:, A:A + 1], z=z, rnn_hxs=rnn_out, feats=feats)

    # ------------------------------------------------------------------ training minibatch
    def train_minibatch(self, sample, clip_param, value_coef, entropy_coef, use_clipped_value_loss=True,
                        global_rows=None, xcat_prefilled=None, loss="ppo", grad_ready=None):
        """Forward, fused loss forward+backward, full backward for one minibatch.  loss="ppo": PKG/algo/ppo.py:57-81;
        loss="a2c": PKG/algo/a2c_acktr.py:49-52,71-72 (old_v / old_logp / adv of `sample` are unused and may be None).
        Leaves d(loss)/d(params) in the flat gradient buffer and the three loss partial sums at its tail; nothing is
        synchronised with the host.
        grad_ready(lo, hi): optional callback, called with the CURRENT stream being the one on which flat_grad[lo:hi] becomes final.
        Data-parallel training uses it to all-reduce the gradient in two buckets: [fc.w, end) -- FC, GRU and head gradients plus
        the loss partials, 97 % of the bytes, final before the convolution backward starts -- overlaps the convolution backward;
        [0, fc.w) follows at the end.  Without the callback (or when the minibatch is cut into chunks) nothing changes."""
        self._sp = _lib.stream_ptr(self.device)
        try:
            return self._train_minibatch(sample, clip_param, value_coef, entropy_coef, use_clipped_value_loss, global_rows, xcat_prefilled,
                                         loss, grad_ready)
        finally:
            self._sp = None

    def _train_minibatch(self, sample, clip_param, value_coef, entropy_coef, use_clipped_value_loss, global_rows, xcat_prefilled, loss,
                         grad_ready):
        obs, vobs, h0, actions, old_v, ret, masks, old_logp, adv = sample
        out = self.forward(obs, vobs, h0, masks, keep=True, xcat_prefilled=xcat_prefilled)
        L = lib()
        sv = self._saved
        B, H, A = sv["B"], self.H, self.A
        dev = self.device
        main = torch.cuda.current_stream(dev)
        z = sv["z"]
        dz = self.buf("t_dz", B, A + 1)
        self.flat_grad.zero_()                                              # optimizer.zero_grad(), ppo.py:79
        loss_out = self.flat_grad[self.loss_off:self.loss_off + 3]
        ws = _lib.workspace(L.ppd_ppo_loss_workspace(B), dev, "loss")
        f32 = lambda t: t.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        act = actions.to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
        if loss == "a2c":
            check(L.ppd_a2c_loss_fwd_bwd(z.data_ptr(), A + 1, A, act.data_ptr(), f32(ret).data_ptr(), B, int(global_rows or B),
                                         float(value_coef), float(entropy_coef), dz.data_ptr(), loss_out.data_ptr(),
                                         ws.data_ptr(), ws.numel(), self.stream), "a2c_loss")
        else:
            check(L.ppd_ppo_loss_fwd_bwd(z.data_ptr(), A + 1, A, act.data_ptr(), f32(old_logp).data_ptr(), f32(adv).data_ptr(),
                                         f32(old_v).data_ptr(), f32(ret).data_ptr(), B, int(global_rows or B),
                                         float(clip_param), float(value_coef), float(entropy_coef),
                                         int(bool(use_clipped_value_loss)), dz.data_ptr(), None, None, loss_out.data_ptr(),
                                         ws.data_ptr(), ws.numel(), self.stream), "ppo_loss")
        # single chunk: every dY buffer stays valid until the end, so all bias gradients are reduced together then
        self._deferred = [] if len(sv["chunks"]) == 1 else None
        # ---- heads backward (weight / bias gradients on the side stream)
        feats = sv["hs"] if self.recurrent else sv["feat"]
        with self._Side(self):
            self._gemm(dz, A + 1, 0, feats, H, 0, self.seg("heads.w", True), H, A + 1, H, B)
            self._colsum(dz, A + 1, B, A + 1, self.seg("heads.b", True))
        chunks, cols = sv["chunks"], sv["cols"]
        dfeat = self.buf("t_dfeat", B, H)
        # early gradient bucket [fc.w, end): only when the minibatch is one chunk (every gradient is written exactly once)
        early = grad_ready is not None and len(chunks) == 1 and self._deferred is not None and self.precision == "tf32x3"
        fc_off = self.segs["fc.w"].off

        def early_bucket():
            # (called on the side stream, after the FC weight gradient) bias sums of the bucket, then hand the range over
            self._flush_colsums(from_off=fc_off)
            if self.recurrent:
                self.seg("gru.b_hh", True)[:2 * H].copy_(self.seg("gru.b_ih", True)[:2 * H])
            grad_ready(fc_off, self.flat_grad.numel())
            # leave SMs to the collective while the convolution backward runs: a persistent kernel that fills all 148 SMs would
            # make the NCCL kernel wait for a kernel boundary and then delay one of the next kernel's statically scheduled CTAs
            if self.comm_ctas:
                L.ppd_tc_gemm_set_option(1000 + 148 - self.comm_ctas)
        if self.recurrent:
            T, E, Ipad = sv["T"], sv["E"], self.Ipad
            dhs = self.buf("t_dhs", B, H)
            self._gemm(dz, A + 1, 1, self.seg("heads.w"), H, 0, dhs, H, B, H, A + 1)
            dgi = self.buf("t_dgi", B, 3 * H)
            dghn = self.buf("t_dghn", B, H)
            saves = [self.buf("t_s%d" % i, B, H) for i in range(4)]
            dh0 = self.buf("t_dh0", E, H)
            w_hh, w_ih = self.seg("gru.w_hh"), self.seg("gru.w_ih")
            xcat, hs, m = sv["xcat"], sv["hs"], sv["m"]
            gstream = self._gru_stream_ctx(len(chunks))
            if gstream is not None:
                ev = torch.cuda.Event()
                ev.record(main)
                gstream.wait_event(ev)
            # ---- GRU parameter gradients over all T*E rows (side stream; dgi complete because main waited for it)
            def gru_param_grads():
                with self._Side(self):
                    hm = self.buf("t_hm", B, H)
                    check(L.ppd_gru_masked_prev(hs.data_ptr(), sv["h0"].data_ptr(), m.data_ptr(), T, E, H,
                                                hm.data_ptr(), self.stream), "masked_prev")
                    gw_hh = self.seg("gru.w_hh", True)
                    self._gemm(dgi, 3 * H, 0, xcat, Ipad, 0, self.seg("gru.w_ih", True), Ipad, 3 * H, Ipad, B)
                    self._colsum(dgi, 3 * H, B, 3 * H, self.seg("gru.b_ih", True))
                    self._gemm(dgi, 3 * H, 0, hm, H, 0, gw_hh, H, 2 * H, H, B)                    # r, z rows of dW_hh
                    self._gemm(dghn, H, 0, hm, H, 0, gw_hh[2 * H * H:], H, H, H, B)                # n rows
                    gb_hh = self.seg("gru.b_hh", True)
                    if self._deferred is None:
                        gb_hh[:2 * H].copy_(self.seg("gru.b_ih", True)[:2 * H])                  # same sums for r, z
                    self._colsum(dghn, H, B, H, gb_hh[2 * H:])
            # BPTT runs over the time chunks from the last to the first on the GRU stream; as soon as a chunk's dgi is
            # there, the main stream back-propagates that chunk through the FC / conv trunk while the GRU continues
            for ci in range(len(chunks) - 1, -1, -1):
                r0, r1 = chunks[ci]
                n = r1 - r0
                h_in = sv["h0"] if ci == 0 else hs[r0 - E:r0]

                def run_gru():
                    if ci != len(chunks) - 1:
                        dhs[r1 - E:r1].add_(dh0)                      # gradient flowing back from the next chunk
                    check(L.ppd_gru_backward(dhs[r0:].data_ptr(), m[r0:].data_ptr(), w_hh.data_ptr(), h_in.data_ptr(),
                                             hs[r0:].data_ptr(), *[s_[r0:].data_ptr() for s_ in saves], n // E, E, H,
                                             dgi[r0:].data_ptr(), dghn[r0:].data_ptr(), dh0.data_ptr() if ci > 0 else None,
                                             self.stream), "gru_backward")
                if gstream is None:
                    run_gru()
                else:
                    with self._on_stream(gstream):
                        run_gru()
                        ev = torch.cuda.Event()
                        ev.record(gstream)
                    main.wait_event(ev)
                # d(feat) = dgi W_ih[:, :H], masked by feat > 0 (ReLU of the FC layer)
                self._gemm(dgi[r0:], 3 * H, 1, w_ih, Ipad, 0, dfeat[r0:], H, n, H, 3 * H, mask=xcat[r0:], ldm=Ipad, b_param=True)
                if early:
                    gru_param_grads()          # before the convolution backward: the early gradient bucket is complete sooner
                self._trunk_backward_rows(sv["obs"], r0, r1, cols, r0 if cols[3] else 0, cols[3], dfeat, H,
                                          0 if ci == len(chunks) - 1 else 1, after_fc=early_bucket if early else None)
            if not early:
                gru_param_grads()
        else:
            self._gemm(dz, A + 1, 1, self.seg("heads.w"), H, 0, dfeat, H, B, H, A + 1, mask=feats, ldm=H)
            for ci, (r0, r1) in enumerate(chunks):
                self._trunk_backward_rows(sv["obs"], r0, r1, cols, r0 if cols[3] else 0, cols[3], dfeat, H, 0 if ci == 0 else 1,
                                          after_fc=early_bucket if early else None)
        if self._deferred is not None:
            gb = None
            if self.recurrent and not early:      # b_hh's r,z sums equal b_ih's: drop that copy from the deferred list's dependencies
                gb = (self.seg("gru.b_hh", True), self.seg("gru.b_ih", True))
            with self._Side(self):
                self._flush_colsums()
                if gb is not None:
                    gb[0][:2 * H].copy_(gb[1][:2 * H])
        self._join()
        if grad_ready is not None:
            if early and self.comm_ctas:
                L.ppd_tc_gemm_set_option(1000)
            grad_ready(0, fc_off if early else self.flat_grad.numel())          # the remaining bucket (or, without overlap, everything)
        return out

    # ------------------------------------------------------------------ optimiser
    def rmsprop_step(self, lr, alpha, eps, max_grad_norm, grad_norm_out=None):
        """clip_grad_norm_ + RMSprop.step over the flat buffers (a2c_acktr.py:74-78)."""
        self.bind()
        L = lib()
        if getattr(self, "rms_state", None) is None or self.rms_state.numel() != self.flat.numel() or self.rms_state.device != self.device:
            self.rms_state = torch.zeros_like(self.flat)
        n = self.n_params
        ws = _lib.workspace(L.ppd_clip_adam_workspace(n), self.device, "adam")
        check(L.ppd_clip_rmsprop_step(self.flat.data_ptr(), self.flat_grad.data_ptr(), self.rms_state.data_ptr(), n, float(lr), float(alpha),
                                      float(eps), float(max_grad_norm) if max_grad_norm else 0.0,
                                      grad_norm_out.data_ptr() if grad_norm_out is not None else None,
                                      ws.data_ptr(), ws.numel(), _lib.stream_ptr(self.device)), "clip_rmsprop_step")

    def adam_step(self, lr, betas, eps, max_grad_norm, loss_acc=None, grad_norm_out=None):
        """clip_grad_norm_ + Adam.step over the flat buffers (ppo.py:82-84)."""
        self.bind()
        L = lib()
        stt = self.adam_state
        stt["step"] += 1
        n = self.n_params
        ws = _lib.workspace(L.ppd_clip_adam_workspace(n), self.device, "adam")
        loss_in = self.flat_grad[self.loss_off:self.loss_off + 3] if loss_acc is not None else None
        check(L.ppd_clip_adam_step(self.flat.data_ptr(), self.flat_grad.data_ptr(), stt["exp_avg"].data_ptr(),
                                   stt["exp_avg_sq"].data_ptr(), n, stt["step"], float(lr), float(betas[0]), float(betas[1]),
                                   float(eps), float(max_grad_norm) if max_grad_norm else 0.0,
                                   grad_norm_out.data_ptr() if grad_norm_out is not None else None,
                                   loss_in.data_ptr() if loss_in is not None else None,
                                   loss_acc.data_ptr() if loss_acc is not None else None,
                                   ws.data_ptr(), ws.numel(), _lib.stream_ptr(self.device)), "clip_adam_step")

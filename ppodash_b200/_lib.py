"""ctypes binding of libppodash_b200.so (the C ABI declared in include/ppodash_b200.h).

There is no CPU fallback: if the shared library is missing, or a kernel entry point is called
with a tensor that is not on a CUDA device, this module raises.  Build the library with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C ppodash_b200/csrc``.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# PPD_LIB: an alternative build of the same library (A/B timing of kernel variants on one GPU box)
LIB_PATH = os.environ.get("PPD_LIB") or os.path.join(_HERE, "libppodash_b200.so")


class PpdError(RuntimeError):
    pass


class GatherDesc(Structure):
    """Mirror of ``ppd_gather_desc``."""
    _fields_ = [
        ("obs", c_void_p), ("obs_out", c_void_p), ("obs_row", c_int64),
        ("vobs", c_void_p), ("vobs_out", c_void_p), ("vobs_row", c_int64), ("vobs_out_ld", c_int64),
        ("hxs", c_void_p), ("hxs_out", c_void_p), ("hxs_row", c_int64),
        ("actions", c_void_p), ("actions_out", c_void_p), ("actions_row", c_int64),
        ("value_preds", c_void_p), ("value_preds_out", c_void_p),
        ("returns", c_void_p), ("returns_out", c_void_p),
        ("masks", c_void_p), ("masks_out", c_void_p),
        ("logp", c_void_p), ("logp_out", c_void_p),
        ("adv", c_void_p), ("adv_out", c_void_p),
        ("adv_stats", c_void_p),
    ]


class ObsU8Desc(Structure):
    """Mirror of ``ppd_obs_u8_desc``."""
    _fields_ = [("frames", c_void_p), ("age", c_void_p), ("mean", c_void_p), ("divisor", c_double), ("C", c_int), ("HW", c_int64),
                ("nstack", c_int), ("multiply_exact", c_int)]


class GemmArgs(Structure):
    """Mirror of ``ppd_gemm_args``."""
    _fields_ = [
        ("A", c_void_p), ("lda", c_int64), ("a_kmajor", c_int),
        ("B", c_void_p), ("ldb", c_int64), ("b_kmajor", c_int),
        ("C", c_void_p), ("ldc", c_int64),
        ("I", c_int64), ("J", c_int64), ("KK", c_int64),
        ("bias", c_void_p),
        ("mask", c_void_p), ("ldm", c_int64),
        ("relu", c_int), ("accumulate", c_int),
    ]


class ColsumSeg(Structure):
    """Mirror of ``ppd_colsum_seg``."""
    _fields_ = [("X", c_void_p), ("ld", c_int64), ("I", c_int64), ("J", c_int64), ("out", c_void_p), ("accumulate", c_int)]


class ConvGeom(Structure):
    """Mirror of ``ppd_conv_geom``."""
    _fields_ = [("B", c_int), ("H", c_int), ("W", c_int), ("C", c_int), ("kh", c_int), ("kw", c_int), ("stride", c_int)]


_P = c_void_p
_PROTOTYPES = {
    "ppd_abi_version": (c_int, []),
    "ppd_last_error": (c_char_p, []),
    "ppd_launch_count": (c_int64, []),
    "ppd_reset_launch_count": (None, []),
    "ppd_upload_rows": (c_int, [_P, c_size_t, _P, c_size_t, c_size_t, c_size_t, _P]),
    "ppd_compute_returns_workspace": (c_size_t, [c_int, c_int]),
    "ppd_compute_returns_set_tuning": (None, [c_int, c_int]),
    "ppd_compute_returns": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_double, c_double, c_int, c_int,
                                    _P, c_size_t, _P]),
    "ppd_advantage_moments_workspace": (c_size_t, [c_int64]),
    "ppd_advantage_moments": (c_int, [_P, _P, c_int64, _P, _P, c_size_t, _P]),
    "ppd_advantage_finalize": (c_int, [_P, _P, _P]),
    "ppd_advantage_normalize": (c_int, [_P, _P, c_int64, _P, _P, _P]),
    "ppd_gather_feed_forward": (c_int, [POINTER(GatherDesc), _P, c_int64, c_int64, c_int, c_int, _P]),
    "ppd_gather_recurrent": (c_int, [POINTER(GatherDesc), _P, c_int64, c_int, c_int, c_int, _P]),
    "ppd_obs_u8_certify": (c_int, [_P, c_int64, c_double, _P, _P]),
    "ppd_obs_u8_expand": (c_int, [POINTER(ObsU8Desc), c_int64, c_int, _P, _P]),
    "ppd_gather_obs_u8_feed_forward": (c_int, [POINTER(ObsU8Desc), _P, c_int64, c_int64, c_int, c_int, _P, _P]),
    "ppd_gather_obs_u8_recurrent": (c_int, [POINTER(ObsU8Desc), _P, c_int64, c_int, c_int, c_int, _P, _P]),
    "ppd_ppo_loss_workspace": (c_size_t, [c_int64]),
    "ppd_ppo_loss_fwd_bwd": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, c_int64, c_int64, c_float, c_float,
                                     c_float, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "ppd_a2c_loss_fwd_bwd": (c_int, [_P, c_int, c_int, _P, _P, c_int64, c_int64, c_float, c_float, _P, _P, _P, c_size_t, _P]),
    "ppd_clip_rmsprop_step": (c_int, [_P, _P, _P, c_int64, c_double, c_double, c_double, c_double, _P, _P, c_size_t, _P]),
    "ppd_categorical_eval": (c_int, [_P, c_int, c_int, _P, c_int64, _P, _P, _P, _P, _P]),
    "ppd_clip_adam_workspace": (c_size_t, [c_int64]),
    "ppd_clip_adam_set_fused": (None, [c_int]),
    "ppd_clip_adam_step": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_double, c_double, c_double, c_double,
                                   c_double, _P, _P, _P, _P, c_size_t, _P]),
    "ppd_obs_rms_update_normalize": (c_int, [_P, c_int, c_int64, _P, _P, c_double, c_int, c_double, c_double,
                                             _P, _P]),
    "ppd_sgemm_workspace": (c_size_t, [c_int64, c_int64, c_int64]),
    "ppd_sgemm": (c_int, [POINTER(GemmArgs), _P, c_size_t, _P]),
    "ppd_tc_gemm_workspace": (c_size_t, [c_int64, c_int64, c_int64]),
    "ppd_tc_gemm_supported": (c_int, [POINTER(GemmArgs)]),
    "ppd_tc_gemm": (c_int, [POINTER(GemmArgs), c_int, _P, c_size_t, _P]),
    "ppd_tc_gemm_set_option": (None, [c_int]),
    "ppd_split_tf32": (c_int, [_P, _P, _P, c_int64, _P]),
    "ppd_conv_fwd_nhwc": (c_int, [_P, POINTER(ConvGeom), c_int, _P, _P, _P, c_int, _P, _P]),
    "ppd_conv_fwd_nchw": (c_int, [_P, POINTER(ConvGeom), c_int, _P, _P, _P, c_int, _P, _P]),
    "ppd_conv_dgrad_nhwc": (c_int, [_P, POINTER(ConvGeom), c_int, _P, _P, _P, _P, _P]),
    "ppd_conv_wgrad_workspace": (c_size_t, [POINTER(ConvGeom), c_int]),
    "ppd_conv_wgrad": (c_int, [_P, POINTER(ConvGeom), c_int, _P, c_int, _P, c_int, _P, c_size_t, _P]),
    "ppd_tc_gemm_bsplit": (c_int, [POINTER(GemmArgs), _P, c_int, _P, c_size_t, _P]),
    "ppd_tc_gemm_col2im": (c_int, [POINTER(GemmArgs), POINTER(ConvGeom), c_int, _P]),
    "ppd_relu_mask": (c_int, [_P, _P, c_int64, _P]),
    "ppd_colsum_workspace": (c_size_t, [c_int64, c_int64]),
    "ppd_colsum": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_int, _P, c_size_t, _P]),
    "ppd_colsum_multi_workspace": (c_size_t, [POINTER(ColsumSeg), c_int]),
    "ppd_colsum_multi": (c_int, [POINTER(ColsumSeg), c_int, _P, c_size_t, _P]),
    "ppd_im2col_nchw": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int64, _P]),
    "ppd_im2col_nhwc": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int64, _P]),
    "ppd_col2im_nhwc": (c_int, [_P, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "ppd_batched_transpose": (c_int, [_P, c_int64, c_int, c_int, _P, _P]),
    "ppd_gru_forward": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "ppd_gru_backward": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "ppd_gru_masked_prev": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "ppd_gru_set_mode": (None, [c_int]),
}

_HOST_ONLY = {"ppd_tc_gemm_supported"}      # predicates evaluated on the host: no kernel behind them, nothing to time

_lib = None


def exported_symbols():
    """Names declared in include/ppodash_b200.h that the library must export."""
    return sorted(_PROTOTYPES)


def register(name, restype, argtypes):
    """Used by the other binding modules (network kernels) to add prototypes."""
    _PROTOTYPES[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype, fn.argtypes = restype, argtypes


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PpdError(
                f"{LIB_PATH} not found: the CUDA extension is not built. ppodash_b200 has no CPU or "
                "PyTorch fallback; run `make -C ppodash_b200/csrc` (needs nvcc, sm_100a).")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _PROTOTYPES.items():
            fn = getattr(handle, name)          # AttributeError here == header/library mismatch
            fn.restype, fn.argtypes = restype, argtypes
        if handle.ppd_abi_version() != 1:
            raise PpdError("libppodash_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ppd_last_error().decode("utf-8", "replace")
        raise PpdError(f"{what or 'ppodash_b200'} failed (code {rc}): {msg}")


def ptr(t, dtype=None):
    """Device pointer of a CUDA tensor (None -> NULL).  Refuses CPU tensors: no fallback."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError("expected a torch.Tensor")
    if not t.is_cuda:
        raise PpdError("ppodash_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


_workspaces = {}


def workspace(nbytes, device, tag="default", zero=False):
    """Grow-only uint8 scratch buffer per (device, tag); kernels that need scratch get it from here
    so that the C ABI itself never allocates.  ``zero=True``: zero-filled when (re)allocated."""
    key = (str(device), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        alloc = torch.zeros if zero else torch.empty
        buf = alloc(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


_profiling = 0


class _Profile:
    """CUDA-event timing of every C-ABI call made while active (used by bench.py to attribute the
    step time to kernels).  Events are recorded on the stream the kernels are launched on."""

    def __init__(self):
        self.records = []          # (name, start_event, end_event)
        self._saved = {}

    def __enter__(self):
        global _profiling
        _profiling += 1
        handle = lib()
        for name in _PROTOTYPES:
            fn = getattr(handle, name)
            restype, argtypes = _PROTOTYPES[name]
            if restype is not c_int or not argtypes or name in _HOST_ONLY:
                continue
            self._saved[name] = fn

            def wrapped(*args, _fn=fn, _name=name):
                a = torch.cuda.Event(enable_timing=True)
                b = torch.cuda.Event(enable_timing=True)
                a.record()
                rc = _fn(*args)
                b.record()
                self.records.append((_name, a, b))
                return rc
            setattr(handle, name, wrapped)
        return self

    def __exit__(self, *exc):
        global _profiling
        _profiling -= 1
        handle = lib()
        for name, fn in self._saved.items():
            setattr(handle, name, fn)
        return False

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.records:
            d = out.setdefault(name, dict(ms=0.0, calls=0))
            d["ms"] += a.elapsed_time(b)
            d["calls"] += 1
        return out


def profiled():
    return _Profile()


_replayed = 0          # kernel launches made by CUDA-graph replays (minibatch_graph.py): the library's counter does not see those


def note_replayed_launches(n):
    global _replayed
    _replayed += int(n)


def launch_count():
    """Kernels launched so far: the library's own count of its launches plus those replayed from captured graphs."""
    return int(lib().ppd_launch_count()) + _replayed


def reset_launch_count():
    global _replayed
    _replayed = 0
    lib().ppd_reset_launch_count()


def live_workspaces():
    """The workspace buffers alive right now (a captured graph references them so that a later, larger request, which replaces
    the cache entry, does not free memory the graph still points at)."""
    return list(_workspaces.values())


def profiling():
    """True while a _lib.profiled() block is active (per-call CUDA events: graph replays would hide the calls from it)."""
    return _profiling > 0

"""Running observation normalisation: VecNormalize._obfilt (PKG/envs.py:186-229) and the
baselines ``RunningMeanStd`` it relies on (third party, unpinned; see oracle/running_mean_std.py).

``RunningMeanStd`` keeps float64 ``mean`` / ``var`` on the device and a host-side ``count``;
``VecNormalizeObs`` is the observation half of the reference's ``VecNormalize`` wrapper as a
callable on device tensors: one kernel folds the batch into the running moments and writes the
normalised, clipped observations.
"""
import torch

from . import _lib
from ._lib import check, lib


class RunningMeanStd:
    def __init__(self, shape=(), epsilon=1e-4, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.PpdError("RunningMeanStd lives on a CUDA device (no CPU fallback)")
        self.shape = tuple(shape)
        self.mean = torch.zeros(self.shape, dtype=torch.float64, device=dev)
        self.var = torch.ones(self.shape, dtype=torch.float64, device=dev)
        self.count = float(epsilon)

    def _run(self, x, update, out, epsilon, clipob):
        x = x.to(device=self.mean.device, dtype=torch.float32).contiguous()
        if tuple(x.shape[1:]) != self.shape:
            raise ValueError(f"expected a batch of shape [N,{self.shape}], got {tuple(x.shape)}")
        N = x.shape[0]
        F = self.mean.numel()
        check(lib().ppd_obs_rms_update_normalize(x.data_ptr(), N, max(F, 1), self.mean.data_ptr(), self.var.data_ptr(),
                                                 self.count, int(update), float(epsilon), float(clipob),
                                                 out.data_ptr() if out is not None else None,
                                                 _lib.stream_ptr(self.mean.device)), "obs_rms")
        if update:
            self.count += N

    def update(self, x):
        """Fold a batch x [N, *shape] into the running moments (Chan et al. parallel merge)."""
        self._run(x, True, None, 1e-8, 10.0)


class VecNormalizeObs:
    """obs -> clip((obs - mean) / sqrt(var + epsilon), -clipob, clipob), updating the running
    moments first when training (envs.py:208-217)."""

    def __init__(self, shape, clipob=10., epsilon=1e-8, device="cuda"):
        self.ob_rms = RunningMeanStd(shape=shape, device=device)
        self.clipob = clipob
        self.epsilon = epsilon
        self.training = True

    def __call__(self, obs, update=True):
        out = torch.empty(obs.shape, dtype=torch.float32, device=self.ob_rms.mean.device)
        self.ob_rms._run(obs, self.training and update, out, self.epsilon, self.clipob)
        return out

    def train(self):
        self.training = True

    def eval(self):
        self.training = False

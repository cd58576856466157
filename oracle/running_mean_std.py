"""Oracle (test infrastructure): running observation normalisation.  PARITY UNPINNED.

Restates ``VecNormalize._obfilt`` (PKG/envs.py:208-217) and the third-party
``baselines.common.running_mean_std.RunningMeanStd`` it relies on (module
openai/baselines, unpinned -- ``git+git://github.com/openai/baselines`` at HEAD,
ppo-dash-training/environment.yml:18; call sites PKG/envs.py:13,189-190,211-213).
baselines is neither vendored in /root/reference nor installable here and the
reference has no test for it, so this file follows the published algorithm
(Chan et al. parallel-moments merge, float64, mean=0 var=1 count=1e-4 initial
state) and is checked only against the pooled-moments property in
``tests/test_oracle_golden.py``.
"""
import numpy as np


class RunningMoments:
    def __init__(self, shape=(), epsilon=1e-4):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = float(epsilon)

    def update(self, x):
        x = np.asarray(x)
        self.update_from_moments(x.mean(axis=0, dtype=np.float64),
                                 x.var(axis=0, dtype=np.float64), x.shape[0])

    def update_from_moments(self, b_mean, b_var, b_count):
        delta = b_mean - self.mean
        tot = self.count + b_count
        new_mean = self.mean + delta * b_count / tot
        m2 = self.var * self.count + b_var * b_count + np.square(delta) * self.count * b_count / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


def obs_filter(rms, obs, clipob=10.0, epsilon=1e-8, update=True):
    """envs.py:208-217: optionally fold ``obs`` [N,...] into the running moments,
    then return clip((obs - mean) / sqrt(var + eps), -clipob, clipob)."""
    if update:
        rms.update(obs)
    return np.clip((obs - rms.mean) / np.sqrt(rms.var + epsilon), -clipob, clipob)

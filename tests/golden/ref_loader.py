"""Import the UNMODIFIED reference packages from /root/reference (read-only).

Only usable in the authoring container -- /root/reference does not exist on the
GPU box -- and only from tests/golden/make_golden.py and the optional
``reference_present`` CPU tests.  Recipe from SURVEY.md 8c: the reference's
``utils.py`` imports ``<pkg>.envs`` which needs gym + baselines (not installed),
and ``algo/__init__.py`` pulls in kfac; pre-register stub packages so that the
real storage.py / model.py / distributions.py / algo/ppo.py files load as-is.
"""
import importlib
import os
import sys
import types

REF_ROOT = "/root/reference"
PKG_A = os.path.join(REF_ROOT, "ppo-dash-training/pytorch-a2c-ppo-acktr-gail/a2c_ppo_acktr")
PKG_B = os.path.join(REF_ROOT, "ppo-dash-study/001_baseline/ppo")
PKG_013 = os.path.join(
    REF_ROOT, "ppo-dash-study/013_ra+no_stack+lshp+recurrent+vec_obs+norm_obs+rew_hacking/ppo")


def reference_present():
    return os.path.isdir(PKG_A) and os.path.isdir(PKG_B)


def _load(pkg_name, path):
    if pkg_name in sys.modules and getattr(sys.modules[pkg_name], "_ppd_ref", False):
        pkg = sys.modules[pkg_name]
    else:
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [path]
        pkg._ppd_ref = True
        sys.modules[pkg_name] = pkg
        envs = types.ModuleType(pkg_name + ".envs")

        class VecNormalize:          # dummy: only isinstance()-checked by utils.get_vec_normalize
            pass

        envs.VecNormalize = VecNormalize
        sys.modules[pkg_name + ".envs"] = envs
        algo = types.ModuleType(pkg_name + ".algo")
        algo.__path__ = [os.path.join(path, "algo")]
        sys.modules[pkg_name + ".algo"] = algo
    storage = importlib.import_module(pkg_name + ".storage")
    model = importlib.import_module(pkg_name + ".model")
    ppo = importlib.import_module(pkg_name + ".algo.ppo")
    return types.SimpleNamespace(RolloutStorage=storage.RolloutStorage, Policy=model.Policy,
                                 CNNBase=model.CNNBase, PPO=ppo.PPO, storage=storage, model=model)


def load_variant_a():
    """Canonical package (recurrent + vector obs): a2c_ppo_acktr."""
    return _load("a2c_ppo_acktr", PKG_A)


def load_variant_b():
    """001_baseline/ppo (feed-forward oracle; generator does not gather vector_obs / hidden)."""
    return _load("ppo", PKG_B)


class Discrete:
    """Stand-in for gym.spaces.Discrete: only __class__.__name__ and .n are read
    (storage.py:20, model.py:30-31)."""

    def __init__(self, n):
        self.n = n
        self.shape = ()

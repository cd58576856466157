"""Checkpoint interchange with the reference.

The reference saves ``torch.save([actor_critic, ob_rms], <save_dir>/<exp>/<env>.pt)`` -- the whole pickled ``Policy`` module plus the
running observation statistics of ``VecNormalize`` (``None`` in every PPO-Dash study) -- and resumes with ``torch.load`` of the same
pair (ppo-dash-training/pytorch-a2c-ppo-acktr-gail/run.py:64-71,251-262).  Those pickles name the reference's classes
(``ppo.model.Policy``, ``a2c_ppo_acktr.model.Policy``, ``...distributions.Categorical``, baselines' ``RunningMeanStd``) and, for the
2019 files under ``ppo-dash-study/models/``, torch internals that no longer exist (``torch.nn.backends.thnn``).

  load_reference_checkpoint(path)        reads such a file WITHOUT the reference package: the pickled classes are mapped onto inert
                                         stand-ins, the architecture is inferred from the parameter shapes and a ``ppodash_b200.Policy``
                                         with exactly those parameters is returned together with ``ob_rms``
  save_checkpoint(path, policy, ob_rms, optimizer=None)
                                         writes ``[actor_critic, ob_rms]`` as run.py:259-262 does (our Policy pickles without its
                                         engine), plus -- what the reference omits -- the optimiser state as a third element
  load_checkpoint(path)                  reads either format; returns (policy, ob_rms, optimizer_state or None)
  reference_state_dict(policy)           CPU ``state_dict`` with the reference's keys and shapes, for ``ref_policy.load_state_dict``
"""
import pickle
import types
import warnings

import torch
import torch.nn as nn

from .model import Policy

_MODEL_CLASSES = {"Policy", "CNNBase", "NNBase", "MLPBase", "Flatten", "Categorical", "DiagGaussian", "Bernoulli", "AddBias",
                  "FixedCategorical", "FixedNormal", "FixedBernoulli"}


class _RefModule(nn.Module):
    """Inert stand-in for a pickled reference module: holds the unpickled ``__dict__`` (parameters, sub-modules), never runs."""

    def forward(self, *a, **k):
        raise RuntimeError("stand-in for a reference module: convert it with load_reference_checkpoint")


class RefRunningMeanStd:
    """Stand-in for baselines.common.running_mean_std.RunningMeanStd (mean, var, count as pickled)."""


def _stand_in(name):
    return type(name, (_RefModule,), {})


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        leaf = module.rsplit(".", 1)[-1]
        if leaf in ("model", "distributions", "utils") and name in _MODEL_CLASSES and not module.startswith(("torch", "ppodash_b200")):
            return _stand_in(name)
        if name == "RunningMeanStd" and not module.startswith("ppodash_b200"):
            return RefRunningMeanStd
        if module == "torch.nn.backends.thnn":                       # torch <= 1.x pickled a backend getter into every module
            return lambda *a, **k: None
        return super().find_class(module, name)


_pickle_shim = types.ModuleType("ppodash_b200._ref_pickle")
_pickle_shim.Unpickler = _RefUnpickler
_pickle_shim.load = lambda f, **kw: _RefUnpickler(f, **kw).load()
_pickle_shim.loads = pickle.loads
_pickle_shim.dump, _pickle_shim.dumps, _pickle_shim.Pickler = pickle.dump, pickle.dumps, pickle.Pickler
_pickle_shim.__name__ = "pickle"


class _Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


_Discrete.__name__ = "Discrete"


def policy_from_state_dict(sd):
    """A ppodash_b200.Policy whose architecture is inferred from the reference ``state_dict`` (PKG/model.py:15-50,169-190) and whose
    parameters are exactly ``sd``'s."""
    sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items()}
    C = sd["base.main.0.weight"].shape[1]
    H = sd["base.main.7.weight"].shape[0]
    A = sd["dist.linear.weight"].shape[0]
    recurrent = "base.gru.weight_ih_l0" in sd
    V = (sd["base.gru.weight_ih_l0"].shape[1] - H) if recurrent else (sd["base.critic_linear.weight"].shape[1] - H)
    if tuple(sd["base.main.7.weight"].shape) != (H, 32 * 7 * 7):
        raise ValueError("not an 84x84 CNNBase checkpoint (main.7 is %s)" % (tuple(sd["base.main.7.weight"].shape),))
    pol = Policy((C, 84, 84), _Discrete(A), base_kwargs={"recurrent": recurrent, "hidden_size": H}, vector_obs_len=V)
    missing, unexpected = pol.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return pol


def _load(path, map_location):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # SourceChangeWarning for every torch.nn class of a 2019 pickle
        return torch.load(path, map_location=map_location, pickle_module=_pickle_shim, weights_only=False)


def load_reference_checkpoint(path, map_location="cpu"):
    """-> (ppodash_b200.Policy on the CPU, ob_rms) from a file the reference's run.py wrote.  Call ``.to("cuda")`` before use."""
    obj = _load(path, map_location)
    if not (isinstance(obj, (list, tuple)) and len(obj) >= 2):
        raise ValueError("expected the reference's [actor_critic, ob_rms] pair (run.py:259-262)")
    actor_critic, ob_rms = obj[0], obj[1]
    if isinstance(actor_critic, Policy):
        return actor_critic, ob_rms
    return policy_from_state_dict(actor_critic.state_dict()), ob_rms


def reference_state_dict(policy):
    return {k: v.detach().to("cpu").clone().contiguous() for k, v in policy.state_dict().items()}


def save_checkpoint(path, policy, ob_rms=None, optimizer=None):
    """``[actor_critic, ob_rms]`` as run.py:259-262 writes it (+ the optimiser state, which the reference leaves out)."""
    payload = [policy, ob_rms]
    if optimizer is not None:
        payload.append(optimizer.state_dict())
    torch.save(payload, path)


def load_checkpoint(path, map_location="cpu"):
    """Either format -> (policy, ob_rms, optimizer_state or None)."""
    obj = _load(path, map_location)
    pol, ob_rms = obj[0], obj[1]
    if not isinstance(pol, Policy):
        pol = policy_from_state_dict(pol.state_dict())
    return pol, ob_rms, (obj[2] if len(obj) > 2 else None)

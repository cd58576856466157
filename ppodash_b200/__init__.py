"""ppodash_b200 -- B200-native (sm_100a) implementation of the PPO-Dash training hot path.

Drop-in for the reference's ``a2c_ppo_acktr`` package on that path:

    from ppodash_b200 import algo, RolloutStorage, Policy, CNNBase
    agent = algo.PPO(actor_critic, clip, epochs, nmb, vcoef, ecoef, lr=..., eps=..., max_grad_norm=...)

Everything heavy runs in hand-written CUDA kernels behind the C ABI of include/ppodash_b200.h
(ppodash_b200/libppodash_b200.so); there is no CPU or PyTorch fallback.
"""
from . import algo  # noqa: F401
from .model import CNNBase, Categorical, FixedCategorical, NNBase, Policy  # noqa: F401
from .obs_norm import RunningMeanStd, VecNormalizeObs  # noqa: F401
from .rollout import RolloutLoop  # noqa: F401
from .storage import FusedAdvantages, RolloutStorage  # noqa: F401

__all__ = ["algo", "RolloutStorage", "FusedAdvantages", "Policy", "CNNBase", "NNBase", "Categorical",
           "FixedCategorical", "RunningMeanStd", "VecNormalizeObs", "RolloutLoop"]

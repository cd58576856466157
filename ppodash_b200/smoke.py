"""One small invocation of the hot path on cuda:0, checked against the oracle
(__graft_entry__.smoke): compute_returns + one PPO.update on a 16-step x 4-env PPO-Dash-shaped
rollout (3x84x84 obs, 15 vector obs, 8 actions, GRU-512)."""
import numpy as np
import torch

from . import _lib, algo, synthetic
from .model import Policy
from .storage import RolloutStorage


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


def run(verbose=True):
    from oracle import ppo_update as o_upd
    from oracle import returns as o_ret
    dev = "cuda:0"
    cfg = synthetic.RolloutConfig("smoke", 16, 4, 3, 15, 8, True, 1, 2, 1e-4, 0.001)
    roll = synthetic.make_rollout(cfg, seed=7, reset_prob=0.05)
    torch.manual_seed(0)
    pol = Policy((3, 84, 84), Discrete(8), base_kwargs={"recurrent": True}, vector_obs_len=15)
    p0 = {k: v.clone() for k, v in pol.state_dict().items()}
    pol = pol.to(dev)
    st = RolloutStorage(cfg.num_steps, cfg.num_envs, (3, 84, 84), [15], Discrete(8), 512)
    for k in RolloutStorage._FIELDS:
        getattr(st, k).copy_(roll[k])
    st.to(dev)
    _lib.reset_launch_count()
    st.compute_returns(roll["next_value"].to(dev), True, cfg.gamma, cfg.gae_lambda, False)
    want_ret, want_v = o_ret.returns_recurrence(roll["rewards"].numpy(), roll["value_preds"].numpy(), roll["masks"].numpy(),
                                                roll["bad_masks"].numpy(), roll["next_value"].numpy(), True, cfg.gamma,
                                                cfg.gae_lambda, False)
    np.testing.assert_allclose(st.returns.cpu().numpy()[:-1], want_ret[:-1], rtol=1e-5, atol=1e-5)
    agent = algo.PPO(pol, cfg.clip_param, cfg.ppo_epoch, cfg.num_mini_batch, cfg.value_loss_coef, cfg.entropy_coef,
                     lr=cfg.lr, eps=cfg.eps, max_grad_norm=cfg.max_grad_norm)
    torch.manual_seed(11)
    got = agent.update(st)
    launches = _lib.launch_count()
    cpu_roll = dict(roll)
    cpu_roll["returns"] = torch.from_numpy(want_ret)
    cpu_roll["value_preds"] = torch.from_numpy(want_v)
    state = o_upd.UpdateState(p0, lr=cfg.lr, eps=cfg.eps)
    torch.manual_seed(11)
    want = o_upd.ppo_update(state, cpu_roll, recurrent=True, clip_param=cfg.clip_param, ppo_epoch=cfg.ppo_epoch,
                            num_mini_batch=cfg.num_mini_batch, value_loss_coef=cfg.value_loss_coef,
                            entropy_coef=cfg.entropy_coef, max_grad_norm=cfg.max_grad_norm)
    np.testing.assert_allclose(np.array(got), np.array(want), rtol=1e-4, atol=1e-6)
    for k, v in pol.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), state.params[k].detach().numpy(), rtol=0, atol=0.05 * cfg.lr, err_msg=k)
    if verbose:
        print(f"smoke ok: losses {got} (oracle {want}); {launches} ppodash_b200 kernel launches")
    return got, want, launches

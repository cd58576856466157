"""Seeded synthetic Obstacle-Tower-shaped rollouts.

The Unity environment binary is not available offline, so tests and bench.py
fill a rollout with data of the right shape and statistics: N(0,1) visual obs
(the statically normalised obs of studies 011-013), a 15-wide vector obs
(5-way one-hot keys | time in [0,1] | 8-way one-hot previous action | reward),
sparse rewards, ~1/500 episode resets.  Everything is generated on the CPU with
an explicit ``torch.Generator`` so the CPU oracle and the GPU path see the
same bits.
"""
from dataclasses import dataclass

import torch


@dataclass
class RolloutConfig:
    name: str
    num_steps: int           # T
    num_envs: int            # N
    channels: int            # C (obs is C x 84 x 84)
    vector_obs_len: int      # V
    num_actions: int         # A
    recurrent: bool
    ppo_epoch: int
    num_mini_batch: int
    lr: float
    entropy_coef: float
    clip_param: float = 0.1
    value_loss_coef: float = 0.5
    max_grad_norm: float = 0.5
    eps: float = 1e-5
    gamma: float = 0.99
    gae_lambda: float = 0.95
    hidden_size: int = 512
    obs_hw: int = 84


# BASELINE.json configs (SURVEY.md 8 "Config shorthand")
CONFIGS = {
    "c1": RolloutConfig("c1_ff_8x128", 128, 8, 1, 0, 8, False, 4, 4, 2.5e-4, 0.01),
    "c2": RolloutConfig("c2_ppo_dash_full_32x512", 512, 32, 3, 15, 8, True, 8, 8, 1e-4, 0.001),
    "c3": RolloutConfig("c3_001_baseline_32x128", 128, 32, 4, 0, 54, False, 4, 4, 2.5e-4, 0.01),
    "c3_12": RolloutConfig("c3_001_baseline_12ch_32x128", 128, 32, 12, 0, 54, False, 4, 4, 2.5e-4, 0.01),
    "c5": RolloutConfig("c5_shard_1024x512", 512, 1024, 3, 15, 8, True, 8, 8, 1e-4, 0.001),
}


def vector_obs(gen, T1, N, V):
    if V == 0:
        return torch.zeros(T1, N, 0)
    out = torch.zeros(T1, N, V)
    col = 0
    if V >= 5:
        keys = torch.randint(0, 5, (T1, N), generator=gen)
        out[..., :5] = torch.nn.functional.one_hot(keys, 5).float()
        col = 5
    if V > col:
        out[..., col] = torch.rand(T1, N, generator=gen)
        col += 1
    if V >= col + 8:
        prev = torch.randint(0, 8, (T1, N), generator=gen)
        out[..., col:col + 8] = torch.nn.functional.one_hot(prev, 8).float()
        col += 8
    if V > col:
        out[..., col:] = torch.rand(T1, N, V - col, generator=gen) * 0.1
    return out


def scalar_fields(gen, T, N, num_actions, reset_prob=1.0 / 500, bad_prob=0.0, dense_rewards=False):
    """rewards, value_preds, masks, bad_masks, actions, action_log_probs, next_value."""
    u = torch.rand(T, N, 1, generator=gen)
    if dense_rewards:
        rewards = torch.randn(T, N, 1, generator=gen) * 0.1
    else:
        rewards = torch.zeros(T, N, 1)
        rewards[u > 0.97] = 0.1
        big = u > 0.99
        rewards[big] = (1.0 + torch.rand(T, N, 1, generator=gen))[big]
    value_preds = torch.randn(T + 1, N, 1, generator=gen)
    masks = (torch.rand(T + 1, N, 1, generator=gen) >= reset_prob).float()
    bad_masks = (torch.rand(T + 1, N, 1, generator=gen) >= bad_prob).float()
    actions = torch.randint(0, num_actions, (T, N, 1), generator=gen)
    logp = torch.full((T, N, 1), -float(torch.log(torch.tensor(float(num_actions)))))
    logp = logp + 0.05 * torch.randn(T, N, 1, generator=gen)
    next_value = torch.randn(N, 1, generator=gen)
    return dict(rewards=rewards, value_preds=value_preds, masks=masks, bad_masks=bad_masks,
                actions=actions, action_log_probs=logp, next_value=next_value)


def make_rollout(cfg: RolloutConfig, seed=1234, hidden_state_size=None, with_obs=True,
                 reset_prob=1.0 / 500, bad_prob=0.0, obs_shape=None):
    """Dict of CPU tensors under the reference ``RolloutStorage`` attribute names."""
    gen = torch.Generator().manual_seed(seed)
    T, N = cfg.num_steps, cfg.num_envs
    H = hidden_state_size if hidden_state_size is not None else (cfg.hidden_size if cfg.recurrent else 1)
    d = scalar_fields(gen, T, N, cfg.num_actions, reset_prob, bad_prob)
    shp = tuple(obs_shape) if obs_shape is not None else (cfg.channels, cfg.obs_hw, cfg.obs_hw)
    if with_obs:
        d["obs"] = torch.randn(T + 1, N, *shp, generator=gen)
    d["vector_obs"] = vector_obs(gen, T + 1, N, cfg.vector_obs_len)
    d["recurrent_hidden_states"] = 0.1 * torch.randn(T + 1, N, H, generator=gen)
    d["returns"] = torch.zeros(T + 1, N, 1)
    return d

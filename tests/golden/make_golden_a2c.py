#!/usr/bin/env python
"""Generates tests/golden/update_a2c.npz by running the UNMODIFIED reference ``A2C_ACKTR`` (PKG/algo/a2c_acktr.py, acktr=False) on CPU
(authoring container only: needs /root/reference).

The reference's ``update`` calls ``actor_critic.evaluate_actions(obs, hxs, masks, actions)`` -- the upstream 4-argument signature --
while this fork's ``Policy.evaluate_actions`` also takes the vector observations.  The reference class is therefore run against a thin
ADAPTER around the reference Policy that supplies ``rollouts.vector_obs[:-1]`` as the missing argument; everything else (loss, clip,
RMSprop) is the reference's own code.  Two consecutive updates on the same rollout are recorded (the second one exercises the RMSprop
state).
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from make_golden import fill_storage, make_policy, save, small_cfg  # noqa: E402
from ppodash_b200 import synthetic  # noqa: E402

torch.set_num_threads(1)


class FourArgAdapter:
    """Lets a2c_acktr.py:38-43 reach the 5-argument Policy of this fork."""

    def __init__(self, policy, rollouts, V):
        self.policy, self.rollouts, self.V = policy, rollouts, V
        self.recurrent_hidden_state_size = policy.recurrent_hidden_state_size

    def parameters(self):
        return self.policy.parameters()

    def evaluate_actions(self, obs, hxs, masks, actions):
        vobs = self.rollouts.vector_obs[:-1].view(-1, self.V)
        return self.policy.evaluate_actions(obs, vobs, hxs, masks, actions)


def main():
    assert ref_loader.reference_present(), "needs /root/reference"
    A = ref_loader.load_variant_a()
    a2c = importlib.import_module("a2c_ppo_acktr.algo.a2c_acktr")
    C, V, An, H, T, N = 1, 3, 5, 32, 6, 3
    cfg = small_cfg("upd_a2c", T, N, C, V, An, True, 1, 1, H)
    pol = make_policy(A, C, An, V, True, H, seed=5)
    init = {k: v.clone() for k, v in pol.state_dict().items()}
    roll = synthetic.make_rollout(cfg, seed=43, reset_prob=0.1, hidden_state_size=H)
    st = fill_storage(A, cfg, roll, (C, 84, 84), H)
    with torch.no_grad():
        for t in range(T):
            v, a, lp, h = pol.act(st.obs[t], st.vector_obs[t], st.recurrent_hidden_states[t], st.masks[t], deterministic=True)
            st.value_preds[t].copy_(v)
            st.recurrent_hidden_states[t + 1].copy_(h)
        nv = pol.get_value(st.obs[-1], st.vector_obs[-1], st.recurrent_hidden_states[-1], st.masks[-1])
    st.compute_returns(nv, True, 0.99, 0.95, False)
    pre = {k: getattr(st, k).clone() for k in ("obs", "vector_obs", "recurrent_hidden_states", "rewards", "value_preds", "returns",
                                               "action_log_probs", "actions", "masks", "bad_masks")}
    lr, eps, alpha, vcoef, ecoef, mgn = 7e-4, 1e-5, 0.99, 0.5, 0.01, 0.5
    agent = a2c.A2C_ACKTR(FourArgAdapter(pol, st, V), vcoef, ecoef, lr=lr, eps=eps, alpha=alpha, max_grad_norm=mgn)
    out1 = agent.update(st)
    mid = {k: v.clone() for k, v in pol.state_dict().items()}
    out2 = agent.update(st)
    save("update_a2c",
         **{"init." + k: v for k, v in init.items()}, **{"mid." + k: v for k, v in mid.items()},
         **{"final." + k: v for k, v in pol.state_dict().items()}, **{"roll." + k: v for k, v in pre.items()},
         losses1=np.array(out1, dtype=np.float64), losses2=np.array(out2, dtype=np.float64),
         T=T, N=N, C=C, V=V, A=An, H=H, vcoef=vcoef, ecoef=ecoef, lr=lr, eps=eps, alpha=alpha, max_grad_norm=mgn)


if __name__ == "__main__":
    main()

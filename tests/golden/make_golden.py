#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Every fixture stores the seeded inputs it was made from and the reference's
outputs, so the tests need neither the reference nor this script at run time.
Reference entry points exercised (PKG = a2c_ppo_acktr, S001 = 001_baseline/ppo):
  returns_*    RolloutStorage.compute_returns           PKG/storage.py:82-121
  ff_gen_*     RolloutStorage.feed_forward_generator    S001/ppo/storage.py:123-161, PKG/storage.py:123-160
  rec_gen      RolloutStorage.recurrent_generator       PKG/storage.py:162-223
  policy_*     Policy.evaluate_actions / act / get_value PKG/model.py:54-79 (+ autograd grads)
  update_*     PPO.update                               PKG/algo/ppo.py:34-96
  init_full    Policy(...) initial weights under torch.manual_seed(0) (checksums only)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from ppodash_b200 import synthetic  # noqa: E402

torch.set_num_threads(1)


def np_(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **np_(arrays))
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def fill_storage(ref, cfg, roll, obs_shape, hidden):
    st = ref.RolloutStorage(cfg.num_steps, cfg.num_envs, obs_shape, [cfg.vector_obs_len],
                            ref_loader.Discrete(cfg.num_actions), hidden)
    for k in ("obs", "vector_obs", "recurrent_hidden_states", "rewards", "value_preds", "returns",
              "action_log_probs", "actions", "masks", "bad_masks"):
        getattr(st, k).copy_(roll[k])
    return st


def small_cfg(name, T, N, C, V, A, recurrent, epochs, nmb, hidden):
    return synthetic.RolloutConfig(name, T, N, C, V, A, recurrent, epochs, nmb, 2.5e-4, 0.01,
                                   hidden_size=hidden)


# ----------------------------------------------------------------------------- returns
def golden_returns(A):
    cfg = small_cfg("ret", 24, 7, 1, 0, 4, False, 1, 1, 8)
    roll = synthetic.make_rollout(cfg, seed=11, reset_prob=0.1, bad_prob=0.1, obs_shape=(1, 2, 2))
    out = dict(rewards=roll["rewards"], value_preds=roll["value_preds"], masks=roll["masks"],
               bad_masks=roll["bad_masks"], next_value=roll["next_value"],
               gamma=0.99, gae_lambda=0.95)
    for use_gae in (True, False):
        for proper in (True, False):
            st = fill_storage(A, cfg, roll, (1, 2, 2), 1)
            st.returns.fill_(-7.0)       # sentinel: shows which slots each branch leaves untouched
            st.compute_returns(roll["next_value"], use_gae, 0.99, 0.95, proper)
            tag = f"gae{int(use_gae)}_proper{int(proper)}"
            out["returns_" + tag] = st.returns.clone()
            out["value_preds_" + tag] = st.value_preds.clone()
    # advantage normalisation (ppo.py:35-37) on the study default branch
    st = fill_storage(A, cfg, roll, (1, 2, 2), 1)
    st.compute_returns(roll["next_value"], True, 0.99, 0.95, False)
    adv = st.returns[:-1] - st.value_preds[:-1]
    out["adv_norm"] = (adv - adv.mean()) / (adv.std() + 1e-5)
    save("returns", **out)


# ----------------------------------------------------------------------------- generators
GEN_NAMES = ("obs", "vector_obs", "recurrent_hidden_states", "actions", "value_preds", "returns",
             "masks", "old_action_log_probs", "adv_targ")


def golden_generators(A, B):
    # feed-forward, variant B (V=0): slots 1,2 alias actions in the reference -> store the 7 real ones
    cfg = small_cfg("ffB", 6, 5, 2, 0, 4, False, 1, 4, 8)
    roll = synthetic.make_rollout(cfg, seed=21, obs_shape=(2, 4, 4), hidden_state_size=1)
    roll["returns"] = torch.randn(7, 5, 1, generator=torch.Generator().manual_seed(5))
    adv = torch.randn(6, 5, 1, generator=torch.Generator().manual_seed(6))
    st = fill_storage(B, cfg, roll, (2, 4, 4), 1)
    out = {k: roll[k] for k in roll}
    out["advantages"] = adv
    torch.manual_seed(77)
    for k, mb in enumerate(st.feed_forward_generator(adv, 4)):
        for slot in (0, 3, 4, 5, 6, 7, 8):
            out[f"mb{k}_{GEN_NAMES[slot]}"] = mb[slot]
    out["num_minibatches"] = k + 1
    out["seed"] = 77
    save("ff_gen_variant_b", **out)

    # feed-forward, variant A with V>0 and a hidden-state column: all 9 slots are real gathers
    cfg = small_cfg("ffA", 5, 6, 1, 3, 4, False, 1, 3, 8)
    roll = synthetic.make_rollout(cfg, seed=22, obs_shape=(1, 3, 5), hidden_state_size=4)
    roll["returns"] = torch.randn(6, 6, 1, generator=torch.Generator().manual_seed(7))
    adv = torch.randn(5, 6, 1, generator=torch.Generator().manual_seed(8))
    st = fill_storage(A, cfg, roll, (1, 3, 5), 4)
    out = {k: roll[k] for k in roll}
    out["advantages"] = adv
    torch.manual_seed(78)
    for k, mb in enumerate(st.feed_forward_generator(adv, 3)):
        for slot in range(9):
            out[f"mb{k}_{GEN_NAMES[slot]}"] = mb[slot]
    out["num_minibatches"] = k + 1
    out["seed"] = 78
    save("ff_gen_variant_a", **out)

    # recurrent, variant A
    cfg = small_cfg("rec", 6, 8, 2, 3, 4, True, 1, 4, 8)
    roll = synthetic.make_rollout(cfg, seed=23, obs_shape=(2, 3, 3), hidden_state_size=8)
    roll["returns"] = torch.randn(7, 8, 1, generator=torch.Generator().manual_seed(9))
    adv = torch.randn(6, 8, 1, generator=torch.Generator().manual_seed(10))
    st = fill_storage(A, cfg, roll, (2, 3, 3), 8)
    out = {k: roll[k] for k in roll}
    out["advantages"] = adv
    torch.manual_seed(79)
    for k, mb in enumerate(st.recurrent_generator(adv, 4)):
        for slot in range(9):
            out[f"mb{k}_{GEN_NAMES[slot]}"] = mb[slot]
    out["num_minibatches"] = k + 1
    out["seed"] = 79
    save("rec_gen", **out)


# ----------------------------------------------------------------------------- policy
def make_policy(ref, C, A_n, V, recurrent, hidden, seed):
    torch.manual_seed(seed)
    return ref.Policy((C, 84, 84), ref_loader.Discrete(A_n), base=ref.CNNBase,
                      base_kwargs={"recurrent": recurrent, "hidden_size": hidden}, vector_obs_len=V)


def loss_and_grads(pol, obs, vobs, h0, masks, actions, old_v, ret, old_logp, adv, clip, vcoef, ecoef):
    values, logp, ent, hxs = pol.evaluate_actions(obs, vobs, h0, masks, actions)
    ratio = torch.exp(logp - old_logp)
    s1 = ratio * adv
    s2 = torch.clamp(ratio, 1.0 - clip, 1.0 + clip) * adv
    a_loss = -torch.min(s1, s2).mean()
    vclip = old_v + (values - old_v).clamp(-clip, clip)
    v_loss = 0.5 * torch.max((values - ret).pow(2), (vclip - ret).pow(2)).mean()
    pol.zero_grad()
    (v_loss * vcoef + a_loss - ent * ecoef).backward()
    grads = {"grad." + n: p.grad.clone() for n, p in pol.named_parameters()}
    return dict(values=values, logp=logp, entropy=ent, hxs=hxs, value_loss=v_loss, action_loss=a_loss), grads


def golden_policy(A, B):
    g = torch.Generator().manual_seed(31)
    # recurrent + vector obs (variant A), T=5 steps x E=3 envs, resets inside the window
    C, V, An, H, T, E = 2, 3, 5, 32, 5, 3
    pol = make_policy(A, C, An, V, True, H, seed=3)
    Bn = T * E
    obs = torch.randn(Bn, C, 84, 84, generator=g)
    vobs = torch.rand(Bn, V, generator=g)
    h0 = 0.3 * torch.randn(E, H, generator=g)
    masks = torch.ones(T, E, 1)
    masks[0, 1] = 0.0
    masks[2, 0] = 0.0
    masks[3, 2] = 0.0
    masks = masks.reshape(Bn, 1)
    actions = torch.randint(0, An, (Bn, 1), generator=g)
    old_v = torch.randn(Bn, 1, generator=g) * 0.1
    ret = torch.randn(Bn, 1, generator=g) * 0.2
    old_logp = -np.log(An) + 0.02 * torch.randn(Bn, 1, generator=g)
    adv = torch.randn(Bn, 1, generator=g)
    outs, grads = loss_and_grads(pol, obs, vobs, h0, masks, actions, old_v, ret, old_logp, adv, 0.1, 0.5, 0.01)
    with torch.no_grad():
        m1 = torch.tensor([[1.0], [0.0], [1.0]])
        v_act, a_act, lp_act, h_act = pol.act(obs[:E], vobs[:E], h0, m1, deterministic=True)
        v_get = pol.get_value(obs[:E], vobs[:E], h0, m1)
    save("policy_recurrent",
         **{"param." + k: v for k, v in pol.state_dict().items()},
         obs=obs, vobs=vobs, h0=h0, masks=masks, actions=actions, old_v=old_v, ret=ret,
         old_logp=old_logp, adv=adv, clip=0.1, vcoef=0.5, ecoef=0.01, T=T, E=E,
         act_masks=m1, act_value=v_act, act_action=a_act, act_logp=lp_act, act_hxs=h_act, get_value=v_get,
         **{"out." + k: v for k, v in outs.items()}, **grads)

    # feed-forward, variant B (V=0)
    C, An, H, Bn = 1, 6, 32, 10
    pol = make_policy(B, C, An, 0, False, H, seed=4)
    obs = torch.randn(Bn, C, 84, 84, generator=g)
    vobs = torch.zeros(Bn, 0)
    h0 = torch.zeros(Bn, 1)
    masks = torch.ones(Bn, 1)
    actions = torch.randint(0, An, (Bn, 1), generator=g)
    old_v = torch.randn(Bn, 1, generator=g) * 0.1
    ret = torch.randn(Bn, 1, generator=g) * 0.2
    old_logp = -np.log(An) + 0.02 * torch.randn(Bn, 1, generator=g)
    adv = torch.randn(Bn, 1, generator=g)
    outs, grads = loss_and_grads(pol, obs, vobs, h0, masks, actions, old_v, ret, old_logp, adv, 0.1, 0.5, 0.01)
    save("policy_feedforward",
         **{"param." + k: v for k, v in pol.state_dict().items()},
         obs=obs, vobs=vobs, h0=h0, masks=masks, actions=actions, old_v=old_v, ret=ret,
         old_logp=old_logp, adv=adv, clip=0.1, vcoef=0.5, ecoef=0.01,
         **{"out." + k: v for k, v in outs.items()}, **grads)


# ----------------------------------------------------------------------------- PPO.update
def golden_update(A, B):
    for tag, ref, recurrent, C, V, An, T, N, nmb, epochs in (
            ("recurrent", A, True, 2, 3, 5, 8, 4, 2, 2),
            ("feedforward", B, False, 1, 0, 6, 6, 4, 3, 2)):
        H = 32
        cfg = small_cfg("upd_" + tag, T, N, C, V, An, recurrent, epochs, nmb, H)
        pol = make_policy(ref, C, An, V, recurrent, H, seed=5)
        init = {k: v.clone() for k, v in pol.state_dict().items()}
        roll = synthetic.make_rollout(cfg, seed=41, reset_prob=0.1, hidden_state_size=(H if recurrent else 1))
        st = fill_storage(ref, cfg, roll, (C, 84, 84), H if recurrent else 1)
        # self-consistent old log-probs / values: run the policy over the rollout
        with torch.no_grad():
            for t in range(T):
                v, a, lp, h = pol.act(st.obs[t], st.vector_obs[t], st.recurrent_hidden_states[t], st.masks[t],
                                      deterministic=True)
                st.value_preds[t].copy_(v)
                st.action_log_probs[t].copy_(lp)
                st.recurrent_hidden_states[t + 1].copy_(h)
                # keep the seeded random actions but use their log-prob under the policy
                feats = pol.base(st.obs[t], st.vector_obs[t], st.recurrent_hidden_states[t], st.masks[t])[1]
                st.action_log_probs[t].copy_(pol.dist(feats).log_probs(st.actions[t]))
            nv = pol.get_value(st.obs[-1], st.vector_obs[-1], st.recurrent_hidden_states[-1], st.masks[-1])
        st.compute_returns(nv, True, 0.99, 0.95, False)
        pre = {k: getattr(st, k).clone() for k in ("obs", "vector_obs", "recurrent_hidden_states", "rewards",
                                                   "value_preds", "returns", "action_log_probs", "actions",
                                                   "masks", "bad_masks")}
        agent = ref.PPO(pol, 0.1, epochs, nmb, 0.5, 0.01, lr=2.5e-4, eps=1e-5, max_grad_norm=0.5)
        torch.manual_seed(99)
        vl, al, ent = agent.update(st)
        save("update_" + tag,
             **{"init." + k: v for k, v in init.items()},
             **{"final." + k: v for k, v in pol.state_dict().items()},
             **{"roll." + k: v for k, v in pre.items()},
             losses=np.array([vl, al, ent], dtype=np.float64), seed=99,
             T=T, N=N, C=C, V=V, A=An, H=H, nmb=nmb, epochs=epochs,
             clip=0.1, vcoef=0.5, ecoef=0.01, lr=2.5e-4, eps=1e-5, max_grad_norm=0.5)


# ----------------------------------------------------------------------------- full-size init + forward
def golden_full(A, B):
    out = {}
    for tag, ref, recurrent, C, V, An in (("c2", A, True, 3, 15, 8), ("c1", B, False, 1, 0, 8)):
        pol = make_policy(ref, C, An, V, recurrent, 512, seed=0)
        for k, v in pol.state_dict().items():
            out[f"{tag}.sum.{k}"] = v.double().sum()
            out[f"{tag}.abssum.{k}"] = v.double().abs().sum()
        g = torch.Generator().manual_seed(51)
        E, T = 2, 3
        Bn = E * T
        obs = torch.randn(Bn, C, 84, 84, generator=g)
        vobs = torch.rand(Bn, V, generator=g)
        h0 = 0.1 * torch.randn(E if recurrent else Bn, 512 if recurrent else 1, generator=g)
        masks = torch.ones(Bn, 1)
        masks[3] = 0.0
        actions = torch.randint(0, An, (Bn, 1), generator=g)
        with torch.no_grad():
            v, lp, ent, hx = pol.evaluate_actions(obs, vobs, h0, masks, actions)
        out.update({f"{tag}.value": v, f"{tag}.logp": lp, f"{tag}.entropy": ent, f"{tag}.hxs": hx})
    save("init_full", **out)


def main():
    assert ref_loader.reference_present(), "needs /root/reference"
    A = ref_loader.load_variant_a()
    B = ref_loader.load_variant_b()
    golden_returns(A)
    golden_generators(A, B)
    golden_policy(A, B)
    golden_update(A, B)
    golden_full(A, B)


if __name__ == "__main__":
    main()

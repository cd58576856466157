"""GPU: RolloutLoop (the rollout-side inference loop, run.py:168-216) against the step-by-step reference sequence
actor_critic.act -> masks from done flags -> rollouts.insert, eager and CUDA-graph replay."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


def _make(N=8, T=6, C=3, V=15, A=8, seed=0):
    import ppodash_b200 as ppd
    torch.manual_seed(seed)
    pol = ppd.Policy((C, 84, 84), Discrete(A), base_kwargs={"recurrent": True}, vector_obs_len=V)
    pol.cpu_params = {k: v.clone() for k, v in pol.state_dict().items()}
    pol = pol.to(DEV)
    st = ppd.RolloutStorage(T, N, (C, 84, 84), [V], Discrete(A), 512)
    st.to(DEV)
    g = torch.Generator().manual_seed(seed + 1)
    obs0 = torch.randn(N, C, 84, 84, generator=g)
    vobs0 = torch.rand(N, V, generator=g)
    st.obs[0].copy_(obs0)
    st.vector_obs[0].copy_(vobs0)
    env = [dict(obs=torch.randn(N, C, 84, 84, generator=g), vobs=torch.rand(N, V, generator=g),
                rew=torch.randn(N, generator=g), done=(torch.rand(N, generator=g) < 0.3).numpy(),
                bad=(torch.rand(N, generator=g) < 0.1).numpy()) for _ in range(T)]
    return ppd, pol, st, env


def _reference_loop(pol, st, env):
    """run.py:168-216 written out with the reference's own calls (this repo's Policy.act: checks graph == eager)."""
    for e in env:
        s = st.step
        with torch.no_grad():
            value, action, logp, h = pol.act(st.obs[s], st.vector_obs[s], st.recurrent_hidden_states[s], st.masks[s],
                                             deterministic=True)
        masks = torch.FloatTensor([[0.0] if d else [1.0] for d in e["done"]])
        bad_masks = torch.FloatTensor([[0.0] if b else [1.0] for b in e["bad"]])
        st.insert(e["obs"], e["vobs"], h, action, logp, value, e["rew"].unsqueeze(1), masks, bad_masks)


def _oracle_loop(params, obs0, vobs0, env, N, H=512):
    """The same sequence computed by the ORACLE on the CPU (oracle/policy.py `act` = PKG/model.py:54-66, deterministic) with a
    plain-python stand-in for the storage bookkeeping of PKG/storage.py:60-73.  Returns the tensors `insert` would have filled."""
    from oracle import policy as o_pol
    T = len(env)
    out = dict(value_preds=torch.zeros(T + 1, N, 1), action_log_probs=torch.zeros(T, N, 1), actions=torch.zeros(T, N, 1, dtype=torch.int64),
               recurrent_hidden_states=torch.zeros(T + 1, N, H), masks=torch.ones(T + 1, N, 1), bad_masks=torch.ones(T + 1, N, 1),
               rewards=torch.zeros(T, N, 1))
    obs, vobs = obs0, vobs0
    with torch.no_grad():
        for s, e in enumerate(env):
            value, action, logp, h = o_pol.act(params, obs, vobs, out["recurrent_hidden_states"][s], out["masks"][s], True,
                                               deterministic=True)
            out["value_preds"][s], out["actions"][s], out["action_log_probs"][s] = value, action, logp
            out["recurrent_hidden_states"][s + 1] = h
            out["masks"][s + 1] = torch.FloatTensor([[0.0] if d else [1.0] for d in e["done"]])
            out["bad_masks"][s + 1] = torch.FloatTensor([[0.0] if b else [1.0] for b in e["bad"]])
            out["rewards"][s] = e["rew"].unsqueeze(1)
            obs, vobs = e["obs"], e["vobs"]
    return out


@pytest.mark.parametrize("graph", [False, True])
def test_rollout_loop_matches_oracle_sequence(graph):
    """RolloutLoop against the oracle's act -> masks -> insert sequence (not against this repo's own Policy.act)."""
    ppd, pol, st, env = _make()
    obs0, vobs0 = st.obs[0].cpu(), st.vector_obs[0].cpu()
    want = _oracle_loop(pol.cpu_params, obs0, vobs0, env, 8)
    loop = ppd.RolloutLoop(pol, st, deterministic=True, use_cuda_graph=graph)
    for e in env:
        loop.act()
        loop.observe(e["obs"].numpy(), e["vobs"].numpy(), e["rew"].numpy(), e["done"], e["bad"])
    torch.cuda.synchronize()
    for name in ("actions", "masks", "bad_masks", "rewards"):
        assert torch.equal(getattr(st, name).cpu(), want[name]), name
    T = len(env)
    # stated fp32 tolerance: 1e-5 relative (+2e-6 absolute near zero), as for Policy.act in test_gpu_policy_ppo.py
    np.testing.assert_allclose(st.value_preds[:T].cpu().numpy(), want["value_preds"][:T].numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(st.action_log_probs.cpu().numpy(), want["action_log_probs"].numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(st.recurrent_hidden_states.cpu().numpy(), want["recurrent_hidden_states"].numpy(), rtol=1e-5, atol=2e-6)


def test_rollout_graph_survives_update_and_rebind():
    """act (graph) -> PPO.update (forward at T*E rows: grows the engine's scratch pool) -> act (graph) must replay on live
    buffers: the captured graph owns its scratch.  Then the module is moved (.cpu().to()): the flat buffers are re-created and
    the loop must re-capture.  Checked against the eager path on the same inputs after every stage."""
    ppd, pol, st, env = _make(N=8, T=6)
    loop = ppd.RolloutLoop(pol, st, deterministic=True, use_cuda_graph=True)

    def eager_act():
        s = st.step
        with torch.no_grad():
            return [t.clone() for t in pol.act(st.obs[s], st.vector_obs[s], st.recurrent_hidden_states[s], st.masks[s], deterministic=True)]

    def check():
        want = eager_act()
        loop.act()
        got = loop._last
        assert torch.equal(got[1], want[1])
        for a, b in zip(got, want):
            np.testing.assert_allclose(a.float().cpu().numpy(), b.float().cpu().numpy(), rtol=1e-6, atol=1e-7)
        loop._last = None

    for e in env:
        loop.act()
        loop.observe(e["obs"], e["vobs"], e["rew"], e["done"], e["bad"])
    with torch.no_grad():
        nv = pol.get_value(st.obs[-1], st.vector_obs[-1], st.recurrent_hidden_states[-1], st.masks[-1])
    st.compute_returns(nv, True, 0.99, 0.95, False)
    agent = ppd.algo.PPO(pol, 0.1, 2, 2, 0.5, 0.001, lr=1e-3, eps=1e-5, max_grad_norm=0.5)
    g0 = loop._graph
    agent.update(st)                                # parameters change in place; scratch pool of the engine grows
    junk = [torch.randn(1 << 20, device=DEV) for _ in range(8)]          # anything freed by the update is reused here
    st.after_update()
    check()
    assert loop._graph is g0                        # same flat buffers: no re-capture needed, new weights seen through them
    del junk
    pol.cpu()
    pol.to(DEV)                                     # engine.bind() re-creates the flat buffers
    check()
    assert loop._graph is not g0


@pytest.mark.parametrize("graph", [False, True])
def test_rollout_loop_matches_reference_sequence(graph):
    ppd, pol, st_ref, env = _make()
    _reference_loop(pol, st_ref, env)
    _, _, st, _ = _make()
    loop = ppd.RolloutLoop(pol, st, deterministic=True, use_cuda_graph=graph)
    for e in env:
        a = loop.act()
        assert a.is_pinned() and a.shape == (8, 1) and a.dtype == torch.int64
        loop.observe(e["obs"].numpy(), e["vobs"].numpy(), e["rew"].numpy(), e["done"], e["bad"])
    torch.cuda.synchronize()
    assert st.step == st_ref.step
    for name in ("obs", "vector_obs", "rewards", "actions", "masks", "bad_masks"):
        assert torch.equal(getattr(st, name), getattr(st_ref, name)), name
    for name in ("value_preds", "action_log_probs", "recurrent_hidden_states"):     # same kernels, same inputs
        np.testing.assert_allclose(getattr(st, name).cpu().numpy(), getattr(st_ref, name).cpu().numpy(), rtol=1e-6, atol=1e-7,
                                   err_msg=name)


def test_rollout_loop_sampling_is_consistent():
    """Sampled actions come with their own log-probabilities (checked through evaluate_actions) and the loop feeds PPO."""
    ppd, pol, st, env = _make(seed=3)
    loop = ppd.RolloutLoop(pol, st, deterministic=False, use_cuda_graph=True)
    seen = set()
    for e in env:
        a = loop.act()
        seen.update(a.flatten().tolist())
        loop.observe(e["obs"], e["vobs"], e["rew"], e["done"])
    torch.cuda.synchronize()
    assert len(seen) > 1 and all(0 <= x < 8 for x in seen)
    T, N = st.rewards.shape[:2]
    _, logp, _, _ = pol.evaluate_actions(st.obs[0], st.vector_obs[0], st.recurrent_hidden_states[0], st.masks[0], st.actions[0])
    np.testing.assert_allclose(logp.cpu().numpy(), st.action_log_probs[0].cpu().numpy(), rtol=1e-5, atol=1e-6)
    with torch.no_grad():
        nv = pol.get_value(st.obs[-1], st.vector_obs[-1], st.recurrent_hidden_states[-1], st.masks[-1])
    st.compute_returns(nv, True, 0.99, 0.95, False)
    agent = ppd.algo.PPO(pol, 0.1, 1, 2, 0.5, 0.001, lr=1e-4, eps=1e-5, max_grad_norm=0.5)
    out = agent.update(st)
    assert all(np.isfinite(x) for x in out)


@pytest.mark.gpu
def test_staged_upload_matches_direct_copy():
    """RolloutStorage.upload_from with pinned observations (per-env copies on a copy stream, ordered by the first epoch's
    permutation, minibatches waiting only for their envs) gives the same update as a plain copy of every field."""
    from ppodash_b200 import algo, synthetic
    from ppodash_b200.model import Policy
    from ppodash_b200.storage import RolloutStorage

    class Discrete:
        def __init__(self, n):
            self.n = n
            self.shape = ()
    cfg = synthetic.RolloutConfig("staged", 16, 8, 3, 15, 8, True, 2, 4, 1e-4, 0.001)
    roll = synthetic.make_rollout(cfg, seed=3, reset_prob=0.05)
    results = []
    for staged in (False, True):
        torch.manual_seed(0)
        pol = Policy((3, 84, 84), Discrete(8), base_kwargs={"recurrent": True}, vector_obs_len=15).to("cuda:0")
        st = RolloutStorage(cfg.num_steps, cfg.num_envs, (3, 84, 84), [15], Discrete(8), 512)
        st.to("cuda:0")
        host = {k: roll[k].clone().pin_memory() for k in RolloutStorage._FIELDS}
        if staged:
            st.upload_from(host)
            assert getattr(st, "_pending", None) is not None
        else:
            for k in RolloutStorage._FIELDS:
                getattr(st, k).copy_(host[k])
        st.compute_returns(roll["next_value"].to("cuda:0"), True, cfg.gamma, cfg.gae_lambda, False)
        agent = algo.PPO(pol, cfg.clip_param, cfg.ppo_epoch, cfg.num_mini_batch, cfg.value_loss_coef, cfg.entropy_coef,
                         lr=cfg.lr, eps=cfg.eps, max_grad_norm=cfg.max_grad_norm)
        torch.manual_seed(11)
        out = agent.update(st)
        assert getattr(st, "_pending", None) is None
        torch.cuda.synchronize()
        results.append((out, st.obs.clone(), {k: v.clone() for k, v in pol.state_dict().items()}))
    assert results[0][0] == results[1][0]
    assert torch.equal(results[0][1], results[1][1])
    assert torch.equal(results[1][1].cpu(), roll["obs"])
    for k in results[0][2]:
        assert torch.equal(results[0][2][k], results[1][2][k]), k

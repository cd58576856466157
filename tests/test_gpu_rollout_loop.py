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
    pol = ppd.Policy((C, 84, 84), Discrete(A), base_kwargs={"recurrent": True}, vector_obs_len=V).to(DEV)
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
    """run.py:168-216 written out with the reference's own calls."""
    for e in env:
        s = st.step
        with torch.no_grad():
            value, action, logp, h = pol.act(st.obs[s], st.vector_obs[s], st.recurrent_hidden_states[s], st.masks[s],
                                             deterministic=True)
        masks = torch.FloatTensor([[0.0] if d else [1.0] for d in e["done"]])
        bad_masks = torch.FloatTensor([[0.0] if b else [1.0] for b in e["bad"]])
        st.insert(e["obs"], e["vobs"], h, action, logp, value, e["rew"].unsqueeze(1), masks, bad_masks)


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

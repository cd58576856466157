"""GPU: A2C_ACKTR (acktr=False) on the PPO path's kernels against the golden fixture recorded from the reference's
own A2C_ACKTR class and against the oracle at PPO-Dash shapes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import a2c_update as o_a2c  # noqa: E402
from oracle import returns as o_ret  # noqa: E402
import ppodash_b200 as ppd  # noqa: E402
from ppodash_b200 import synthetic  # noqa: E402

DEV = "cuda:0"


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


def T_(x):
    return torch.as_tensor(np.asarray(x))


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
def test_a2c_two_updates_match_reference(golden, precision):
    g = golden("update_a2c")
    C, V, A, H, T, N = (int(g[k]) for k in ("C", "V", "A", "H", "T", "N"))
    pol = ppd.Policy((C, 84, 84), Discrete(A), base_kwargs={"recurrent": True, "hidden_size": H}, vector_obs_len=V)
    pol.load_state_dict({k[5:]: T_(g[k]) for k in g.files if k.startswith("init.")}, strict=True)
    pol = pol.to(DEV)
    pol.engine(precision)
    st = ppd.RolloutStorage(T, N, (C, 84, 84), [V], Discrete(A), H)
    for k in ppd.RolloutStorage._FIELDS:
        getattr(st, k).copy_(T_(g["roll." + k]))
    st.to(DEV)
    lr = float(g["lr"])
    agent = ppd.algo.A2C_ACKTR(pol, float(g["vcoef"]), float(g["ecoef"]), lr=lr, eps=float(g["eps"]), alpha=float(g["alpha"]),
                               max_grad_norm=float(g["max_grad_norm"]))
    out1 = agent.update(st)
    assert all(isinstance(x, float) for x in out1)
    np.testing.assert_allclose(np.array(out1), g["losses1"], rtol=1e-5, atol=1e-7)
    # RMSprop's first step is lr * g / (|g| sqrt(1 - alpha) + eps): entries with |g| ~ eps amplify rounding; stated gate as for Adam
    # in test_gpu_policy_ppo.py: within 5 % of one step
    for name, p in pol.state_dict().items():
        np.testing.assert_allclose(p.cpu().numpy(), g["mid." + name], rtol=0, atol=0.05 * lr, err_msg=name)
    out2 = agent.update(st)
    np.testing.assert_allclose(np.array(out2), g["losses2"], rtol=1e-4, atol=1e-6)
    for name, p in pol.state_dict().items():
        np.testing.assert_allclose(p.cpu().numpy(), g["final." + name], rtol=0, atol=0.1 * lr, err_msg=name)
    with pytest.raises(NotImplementedError):
        ppd.algo.A2C_ACKTR(pol, 0.5, 0.01, lr=lr, eps=1e-5, alpha=0.99, acktr=True)


@pytest.mark.parametrize("u8", [False, True])
def test_a2c_update_ppo_dash_shape_vs_oracle(golden, u8):
    """C = 3, V = 15, A = 8, H = 512, recurrent, 16 steps x 8 envs = 128 rows in one batch; float32 and uint8 storage."""
    T, N, V, A = 16, 8, 15, 8
    cfg = synthetic.RolloutConfig("a2c", T, N, 3, V, A, True, 1, 1, 7e-4, 0.01)
    roll = synthetic.make_rollout(cfg, seed=13, reset_prob=0.05)
    kw = {}
    if u8:
        go = golden("obs_pipeline")
        rng = np.random.RandomState(2)
        frames = torch.from_numpy(rng.randint(0, 256, size=(T + 1, N, 3, 84, 84), dtype=np.uint8))
        mean = np.ascontiguousarray(go["mean"].transpose(2, 0, 1))
        roll["obs"] = ((frames.double() - torch.from_numpy(mean)) / float(go["std"])).float()
        kw = dict(obs_dtype=torch.uint8, obs_mean=mean, obs_std=float(go["std"]))
    torch.manual_seed(0)
    pol = ppd.Policy((3, 84, 84), Discrete(A), base_kwargs={"recurrent": True}, vector_obs_len=V)
    p_cpu = {k: v.clone() for k, v in pol.state_dict().items()}
    pol = pol.to(DEV)
    st = ppd.RolloutStorage(T, N, (3, 84, 84), [V], Discrete(A), 512, **kw)
    st.to(DEV)
    for k in ppd.RolloutStorage._FIELDS:
        getattr(st, k).copy_(frames if (u8 and k == "obs") else roll[k])
    st.compute_returns(roll["next_value"].to(DEV), True, 0.99, 0.95, False)
    ret, v = o_ret.returns_recurrence(roll["rewards"].numpy(), roll["value_preds"].numpy(), roll["masks"].numpy(), roll["bad_masks"].numpy(),
                                      roll["next_value"].numpy(), True, 0.99, 0.95, False)
    cpu_roll = dict(roll)
    cpu_roll["returns"], cpu_roll["value_preds"] = torch.from_numpy(ret), torch.from_numpy(v)
    agent = ppd.algo.A2C_ACKTR(pol, 0.5, 0.01, lr=7e-4, eps=1e-5, alpha=0.99, max_grad_norm=0.5)
    eng = pol.engine()
    grads = {}
    step0 = agent.optimizer.step

    def spy(*a, **k):
        grads.update({n: p.grad.clone() for n, p in pol.named_parameters()})
        return step0(*a, **k)
    agent.optimizer.step = spy
    got = agent.update(st)
    state = o_a2c.A2CState(p_cpu, 7e-4, 1e-5, 0.99)
    ref = {}
    want = o_a2c.a2c_update(state, cpu_roll, recurrent=True, value_loss_coef=0.5, entropy_coef=0.01, max_grad_norm=0.5,
                            on_update=lambda info: ref.update(info))
    np.testing.assert_allclose(np.array(got), np.array(want), rtol=2e-5, atol=1e-7)
    coef = min(1.0, 0.5 / (ref["grad_norm"] + 1e-6))
    for n, gt in grads.items():
        r = (ref["grads"][n] / coef).numpy()
        scale = max(1e-12, float(np.abs(r).max()))
        gn = gt.cpu().numpy()
        err = np.abs(gn - r)
        # natural ReLUs on both sides: one unit flipping at rounding distance from zero shifts every upstream gradient entry by
        # O(1 / rows) (128 rows here) -- the gate of tests/test_gpu_production_size.py: direction within 1e-5 of 1, no entry beyond 1e-2
        # of the scale (the elementwise 1e-5 check with forced ReLU decisions lives there)
        cos = float((gn.astype(np.float64) * r).sum() / (np.linalg.norm(gn.astype(np.float64)) * np.linalg.norm(r.astype(np.float64)) + 1e-300))
        assert cos >= 1.0 - 1e-5 and float(err.max()) <= 1e-2 * scale, (n, cos, float(err.max()) / scale)
    assert eng.rms_state is not None

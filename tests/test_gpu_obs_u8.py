"""GPU: uint8 observation storage with normalise-on-read and device frame stack -- bit-exact against the oracle's
restatement of the reference's env-side float pipeline (oracle/obs_pipeline.py, pinned to the reference wrappers by
tests/golden/obs_pipeline.npz), and a PPO.update that is bit-identical to the float32 storage fed with the reference's values."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import obs_pipeline as o_obs  # noqa: E402
import ppodash_b200 as ppd  # noqa: E402
from ppodash_b200 import synthetic  # noqa: E402

DEV = "cuda:0"


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


def _inputs(g):
    T, N, H, W, C, nstack = (int(x) for x in g["shape"])
    rng = np.random.RandomState(int(g["seed"]))
    frames = rng.randint(0, 256, size=(T + 1, N, H, W, C), dtype=np.uint8)
    dones = rng.rand(T, N) < 0.3
    dones[1, 0] = True
    return frames, dones, nstack


def _filled_storage(frames, dones, nstack, mean_hwc, std, V=0, H=1, A=4):
    """A uint8 RolloutStorage filled the way run.py:140-141,215-216 fills it: slot 0 from reset, then one insert per step."""
    T, N = dones.shape
    C = frames.shape[-1]
    chw = torch.from_numpy(np.ascontiguousarray(frames.transpose(0, 1, 4, 2, 3)))          # uint8 [T+1, N, C, H, W]
    mean_chw = None if mean_hwc is None else np.ascontiguousarray(mean_hwc.transpose(2, 0, 1))
    st = ppd.RolloutStorage(T, N, (nstack * C, 84, 84), [V], Discrete(A), H, obs_dtype=torch.uint8, frame_stack=nstack,
                            obs_mean=mean_chw, obs_std=std)
    st.to(DEV)
    st.obs[0].copy_(chw[0])
    z = lambda *s: torch.zeros(*s)
    for t in range(T):
        masks = torch.FloatTensor([[0.0] if d else [1.0] for d in dones[t]])
        st.insert(chw[t + 1], z(N, V), z(N, H), torch.zeros(N, 1, dtype=torch.int64), z(N, 1), z(N, 1), z(N, 1), masks, torch.ones(N, 1))
    return st, chw


@pytest.mark.parametrize("case", ["norm", "div255", "stack4"])
def test_expand_and_gathers_bit_exact_vs_oracle(golden, case):
    g = golden("obs_pipeline")
    frames, dones, nstack = _inputs(g)
    mean, std, ns = (g["mean"], float(g["std"]), 1) if case == "norm" else ((None, None, 1) if case == "div255" else (g["mean"], float(g["std"]), nstack))
    want = o_obs.rollout_observations(frames, dones, ns, mean, std)                         # float32 [T+1, N, ns*C, H, W]
    st, _ = _filled_storage(frames, dones, ns, mean, std)
    T, N = dones.shape
    for t in range(T + 1):
        assert torch.equal(st.obs_at(t).cpu(), want[t]), (case, t)
    # recurrent generator: rows time-major over the drawn env block
    torch.manual_seed(5)
    rows = [s[0] for s in st.recurrent_generator(None, N)]                                  # N minibatches of one env each
    torch.manual_seed(5)
    perm = torch.randperm(N)
    for k, got in enumerate(rows):
        assert got.dtype == torch.float32 and tuple(got.shape) == (T, want.shape[2], 84, 84)
        assert torch.equal(got.cpu(), want[:T, perm[k]]), (case, "recurrent", k)
    # feed-forward generator: sample-level permutation
    torch.manual_seed(6)
    got = [s[0] for s in st.feed_forward_generator(None, 3)]
    torch.manual_seed(6)
    perm = torch.randperm(T * N)
    mbs = T * N // 3
    flat = want[:T].reshape(T * N, *want.shape[2:])
    for k, gk in enumerate(got):
        assert torch.equal(gk.cpu(), flat[perm[k * mbs:(k + 1) * mbs]]), (case, "ff", k)
    # after_update carries slot T's whole stack into slot 0
    last = st.obs_at(T).clone()
    st.after_update()
    assert torch.equal(st.obs_at(0), last)


def test_update_from_uint8_storage_is_bit_identical_to_float_storage(golden):
    """PPO.update on a uint8 storage == PPO.update on the reference-layout float32 storage holding the reference's normalised
    observations of the same frames (losses and every parameter bit for bit): the gathered minibatches are the same bits."""
    g = golden("obs_pipeline")
    T, N, V, A = 8, 4, 15, 8
    rng = np.random.RandomState(3)
    frames = rng.randint(0, 256, size=(T + 1, N, 84, 84, 3), dtype=np.uint8)
    dones = rng.rand(T, N) < 0.2
    obs_f32 = o_obs.rollout_observations(frames, dones, 1, g["mean"], float(g["std"]))
    cfg = synthetic.RolloutConfig("u8", T, N, 3, V, A, True, 2, 2, 1e-3, 0.001)
    roll = synthetic.make_rollout(cfg, seed=9, with_obs=False)
    roll["masks"][1:] = torch.from_numpy(1.0 - dones.astype(np.float32)).unsqueeze(-1)
    results = []
    for u8 in (False, True):
        torch.manual_seed(0)
        pol = ppd.Policy((3, 84, 84), Discrete(A), base_kwargs={"recurrent": True}, vector_obs_len=V).to(DEV)
        if u8:
            st = ppd.RolloutStorage(T, N, (3, 84, 84), [V], Discrete(A), 512, obs_dtype=torch.uint8,
                                    obs_mean=np.ascontiguousarray(g["mean"].transpose(2, 0, 1)), obs_std=float(g["std"]))
            st.to(DEV)
            st.obs.copy_(torch.from_numpy(np.ascontiguousarray(frames.transpose(0, 1, 4, 2, 3))))
        else:
            st = ppd.RolloutStorage(T, N, (3, 84, 84), [V], Discrete(A), 512)
            st.to(DEV)
            st.obs.copy_(obs_f32)
        for k in ppd.RolloutStorage._FIELDS:
            if k != "obs":
                getattr(st, k).copy_(roll[k])
        st.compute_returns(roll["next_value"].to(DEV), True, 0.99, 0.95, False)
        agent = ppd.algo.PPO(pol, 0.1, cfg.ppo_epoch, cfg.num_mini_batch, 0.5, 0.001, lr=cfg.lr, eps=1e-5, max_grad_norm=0.5)
        torch.manual_seed(4)
        out = agent.update(st)
        torch.cuda.synchronize()
        results.append((out, {k: v.clone() for k, v in pol.state_dict().items()}))
    assert results[0][0] == results[1][0]
    for k in results[0][1]:
        assert torch.equal(results[0][1][k], results[1][1][k]), k


def test_rollout_loop_on_uint8_storage(golden):
    """RolloutLoop over a uint8 storage (frames uploaded as bytes, policy input expanded on the device inside act) fills the same
    values / actions / hidden states as over a float32 storage that is handed the reference's normalised observations."""
    g = golden("obs_pipeline")
    T, N, V, A = 5, 4, 15, 8
    rng = np.random.RandomState(11)
    frames = rng.randint(0, 256, size=(T + 1, N, 84, 84, 3), dtype=np.uint8)
    dones = rng.rand(T, N) < 0.3
    vobs = torch.rand(T + 1, N, V, generator=torch.Generator().manual_seed(1))
    obs_f32 = o_obs.rollout_observations(frames, dones, 1, g["mean"], float(g["std"]))
    chw = np.ascontiguousarray(frames.transpose(0, 1, 4, 2, 3))
    out = []
    for u8 in (False, True):
        torch.manual_seed(0)
        pol = ppd.Policy((3, 84, 84), Discrete(A), base_kwargs={"recurrent": True}, vector_obs_len=V).to(DEV)
        kw = dict(obs_dtype=torch.uint8, obs_mean=np.ascontiguousarray(g["mean"].transpose(2, 0, 1)), obs_std=float(g["std"])) if u8 else {}
        st = ppd.RolloutStorage(T, N, (3, 84, 84), [V], Discrete(A), 512, **kw)
        st.to(DEV)
        st.obs[0].copy_(torch.from_numpy(chw[0]) if u8 else obs_f32[0])
        st.vector_obs[0].copy_(vobs[0])
        loop = ppd.RolloutLoop(pol, st, deterministic=True, use_cuda_graph=True)
        for t in range(T):
            loop.act()
            loop.observe(chw[t + 1] if u8 else obs_f32[t + 1].numpy(), vobs[t + 1].numpy(), np.zeros(N, np.float32), dones[t])
        torch.cuda.synchronize()
        out.append({k: getattr(st, k).clone() for k in ("value_preds", "actions", "action_log_probs", "recurrent_hidden_states", "masks")})
        if u8:
            assert st.obs.dtype == torch.uint8 and torch.equal(st.obs.cpu(), torch.from_numpy(chw))
    for k in out[0]:
        assert torch.equal(out[0][k], out[1][k]), k

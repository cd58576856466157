"""GPU parity: returns/GAE, advantage stats, gathers, fused loss, clip+Adam, obs running-norm.
Every check calls the sm_100a kernels through the C ABI (ctypes) and compares with the oracle
(oracle/, pinned by tests/test_oracle_golden.py) and with the committed golden fixtures."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import minibatch as o_mb  # noqa: E402
from oracle import ppo_update as o_upd  # noqa: E402
from oracle import returns as o_ret  # noqa: E402
from oracle import running_mean_std as o_rms  # noqa: E402
from ppodash_b200 import _lib, synthetic  # noqa: E402
from ppodash_b200.storage import FusedAdvantages, RolloutStorage  # noqa: E402

DEV = "cuda:0"


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


def T_(x):
    return torch.as_tensor(np.asarray(x))


def storage_from(roll, obs_shape, V, A, H, dev=DEV):
    T, N = roll["rewards"].shape[:2]
    st = RolloutStorage(T, N, obs_shape, [V], Discrete(A), H)
    for k in RolloutStorage._FIELDS:
        getattr(st, k).copy_(T_(roll[k]))
    st.to(dev)
    return st


# --------------------------------------------------------------------------- returns
def _ws(T, N):
    w = _lib.workspace(_lib.lib().ppd_compute_returns_workspace(T, N), DEV, "returns", zero=True)
    return w.data_ptr(), w.numel()


@pytest.mark.parametrize("use_gae", [True, False])
@pytest.mark.parametrize("proper", [True, False])
def test_returns_vs_golden(golden, use_gae, proper):
    g = golden("returns")
    roll = {k: g[k] for k in ("rewards", "value_preds", "masks", "bad_masks")}
    T, N = roll["rewards"].shape[:2]
    roll.update(obs=np.zeros((T + 1, N, 1, 2, 2), np.float32), vector_obs=np.zeros((T + 1, N, 0), np.float32),
                recurrent_hidden_states=np.zeros((T + 1, N, 1), np.float32), returns=np.full((T + 1, N, 1), -7.0, np.float32),
                action_log_probs=np.zeros((T, N, 1), np.float32), actions=np.zeros((T, N, 1), np.int64))
    st = storage_from(roll, (1, 2, 2), 0, 4, 1)
    st.compute_returns(T_(g["next_value"]).to(DEV), use_gae, float(g["gamma"]), float(g["gae_lambda"]), proper)
    tag = f"gae{int(use_gae)}_proper{int(proper)}"
    got = st.returns.cpu().numpy()
    # tolerance stated: chunk carries are combined affinely (re-association), so fp32 mixed tolerance
    np.testing.assert_allclose(got, g["returns_" + tag], rtol=1e-5, atol=1e-5)
    assert np.array_equal(st.value_preds.cpu().numpy(), g["value_preds_" + tag])
    # slots the reference leaves untouched keep the sentinel
    if use_gae:
        assert np.all(got[-1] == -7.0)


@pytest.mark.parametrize("T,N", [(1, 1), (16, 32), (17, 33), (300, 70), (512, 32), (1000, 257)])
@pytest.mark.parametrize("use_gae,proper", [(True, False), (True, True), (False, True), (False, False)])
def test_returns_vs_oracle_shapes(T, N, use_gae, proper):
    gen = torch.Generator().manual_seed(T * 1000 + N)
    f = synthetic.scalar_fields(gen, T, N, 4, reset_prob=0.02, bad_prob=0.01 if proper else 0.0)
    want, want_v = o_ret.returns_recurrence(f["rewards"].numpy(), f["value_preds"].numpy(), f["masks"].numpy(),
                                            f["bad_masks"].numpy(), f["next_value"].numpy(), use_gae, 0.99, 0.95, proper)
    L = _lib.lib()
    d = {k: f[k].to(DEV).contiguous() for k in ("rewards", "value_preds", "masks", "bad_masks", "next_value")}
    ret = torch.zeros(T + 1, N, 1, device=DEV)
    _lib.check(L.ppd_compute_returns(d["rewards"].data_ptr(), d["value_preds"].data_ptr(), d["masks"].data_ptr(),
                                     d["bad_masks"].data_ptr(), ret.data_ptr(), d["next_value"].data_ptr(), T, N,
                                     0.99, 0.95, int(use_gae), int(proper), *_ws(T, N), _lib.stream_ptr()))
    np.testing.assert_allclose(ret.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    assert np.array_equal(d["value_preds"].cpu().numpy(), want_v)
    if T <= 16:   # a single chunk: no re-association at all -> bit exact
        assert np.array_equal(ret.cpu().numpy(), want)


@pytest.mark.parametrize("T,N", [(16, 32), (130, 36), (300, 72), (512, 32), (1000, 256), (2048, 128)])
@pytest.mark.parametrize("use_gae,proper", [(True, False), (True, True), (False, True), (False, False)])
def test_returns_persistent_tma_kernel(T, N, use_gae, proper):
    """The persistent TMA-staged variant (selectable, not the default), forced on small shapes ragged in T and N."""
    L = _lib.lib()
    L.ppd_compute_returns_set_tuning(102, 3)
    try:
        gen = torch.Generator().manual_seed(T * 7 + N)
        f = synthetic.scalar_fields(gen, T, N, 4, reset_prob=0.02, bad_prob=0.01 if proper else 0.0)
        want, want_v = o_ret.returns_recurrence(f["rewards"].numpy(), f["value_preds"].numpy(), f["masks"].numpy(),
                                                f["bad_masks"].numpy(), f["next_value"].numpy(), use_gae, 0.99, 0.95, proper)
        d = {k: f[k].to(DEV).contiguous() for k in ("rewards", "value_preds", "masks", "bad_masks", "next_value")}
        ret = torch.zeros(T + 1, N, 1, device=DEV)
        for _ in range(2):      # twice: the workspace re-arms itself between launches
            _lib.check(L.ppd_compute_returns(d["rewards"].data_ptr(), d["value_preds"].data_ptr(), d["masks"].data_ptr(),
                                             d["bad_masks"].data_ptr(), ret.data_ptr(), d["next_value"].data_ptr(), T, N,
                                             0.99, 0.95, int(use_gae), int(proper), *_ws(T, N), _lib.stream_ptr()))
        np.testing.assert_allclose(ret.cpu().numpy(), want, rtol=1e-5, atol=1e-5)
        assert np.array_equal(d["value_preds"].cpu().numpy(), want_v)
    finally:
        L.ppd_compute_returns_set_tuning(100, 3)


def test_returns_edge_masks():
    """all-ones, zero at t=1, zero at t=T, every env zero on the same step, a full zero column."""
    T, N = 64, 40
    gen = torch.Generator().manual_seed(3)
    f = synthetic.scalar_fields(gen, T, N, 4, reset_prob=0.0)
    cases = []
    m = torch.ones(T + 1, N, 1); cases.append(m)
    m = torch.ones(T + 1, N, 1); m[1] = 0; cases.append(m)
    m = torch.ones(T + 1, N, 1); m[T] = 0; cases.append(m)
    m = torch.ones(T + 1, N, 1); m[33] = 0; cases.append(m)
    m = torch.ones(T + 1, N, 1); m[:, 7] = 0; cases.append(m)
    L = _lib.lib()
    for masks in cases:
        want, _ = o_ret.returns_recurrence(f["rewards"].numpy(), f["value_preds"].numpy(), masks.numpy(),
                                           f["bad_masks"].numpy(), f["next_value"].numpy(), True, 0.99, 0.95, False)
        d = {k: f[k].to(DEV).contiguous() for k in ("rewards", "value_preds", "bad_masks", "next_value")}
        md = masks.to(DEV)
        ret = torch.zeros(T + 1, N, 1, device=DEV)
        _lib.check(L.ppd_compute_returns(d["rewards"].data_ptr(), d["value_preds"].data_ptr(), md.data_ptr(),
                                         d["bad_masks"].data_ptr(), ret.data_ptr(), d["next_value"].data_ptr(), T, N,
                                         0.99, 0.95, 1, 0, *_ws(T, N), _lib.stream_ptr()))
        np.testing.assert_allclose(ret.cpu().numpy(), want, rtol=1e-5, atol=1e-5)


def test_returns_full_size_recurrence_property():
    """BASELINE config 4 (4096 envs x 2048 steps): check the defining recurrence
    A_t = delta_t + gamma*lambda*m_{t+1}*A_{t+1} on the kernel's own output (size-independent)."""
    T, N = 2048, 4096
    gen = torch.Generator(device=DEV).manual_seed(0)
    r = (torch.rand(T, N, 1, device=DEV, generator=gen) > 0.97).float() * 0.1
    v = torch.randn(T + 1, N, 1, device=DEV, generator=gen)
    m = (torch.rand(T + 1, N, 1, device=DEV, generator=gen) > 0.002).float()
    nv = torch.randn(N, 1, device=DEV, generator=gen)
    ret = torch.zeros(T + 1, N, 1, device=DEV)
    _lib.check(_lib.lib().ppd_compute_returns(r.data_ptr(), v.data_ptr(), m.data_ptr(), None, ret.data_ptr(),
                                              nv.data_ptr(), T, N, 0.99, 0.95, 1, 0, *_ws(T, N), _lib.stream_ptr()))
    assert torch.equal(v[T], nv)
    A = (ret[:T] - v[:T]).double()
    delta = r.double() + 0.99 * v[1:].double() * m[1:].double() - v[:T].double()
    A_next = torch.cat([A[1:], torch.zeros(1, N, 1, device=DEV, dtype=torch.float64)], 0)
    resid = (A - (delta + 0.99 * 0.95 * m[1:].double() * A_next)).abs().max().item()
    assert resid < 5e-6, resid


# --------------------------------------------------------------------------- advantages
def test_advantage_stats_and_normalize(golden):
    g = golden("returns")
    ret = T_(g["returns_gae1_proper0"]).to(DEV)
    val = T_(g["value_preds_gae1_proper0"]).to(DEV)
    n = ret[:-1].numel()
    L = _lib.lib()
    ws = torch.empty(L.ppd_advantage_moments_workspace(n), dtype=torch.uint8, device=DEV)
    mom = torch.zeros(3, dtype=torch.float64, device=DEV)
    stats = torch.zeros(2, device=DEV)
    out = torch.empty(n, device=DEV)
    _lib.check(L.ppd_advantage_moments(ret.data_ptr(), val.data_ptr(), n, mom.data_ptr(), ws.data_ptr(), ws.numel(),
                                       _lib.stream_ptr()))
    _lib.check(L.ppd_advantage_finalize(mom.data_ptr(), stats.data_ptr(), _lib.stream_ptr()))
    _lib.check(L.ppd_advantage_normalize(ret.data_ptr(), val.data_ptr(), n, stats.data_ptr(), out.data_ptr(),
                                         _lib.stream_ptr()))
    adv = (T_(g["returns_gae1_proper0"]) - T_(g["value_preds_gae1_proper0"]))[:-1]
    assert abs(stats[0].item() - adv.mean().item()) < 1e-6
    assert abs(stats[1].item() - (adv.std().item() + 1e-5)) < 1e-6
    np.testing.assert_allclose(out.cpu().numpy().reshape(adv.shape), g["adv_norm"], rtol=1e-5, atol=1e-6)
    # large n
    x = torch.randn(1 << 22, device=DEV) * 3 + 0.5
    z = torch.zeros_like(x)
    n = x.numel()
    ws = torch.empty(L.ppd_advantage_moments_workspace(n), dtype=torch.uint8, device=DEV)
    _lib.check(L.ppd_advantage_moments(x.data_ptr(), z.data_ptr(), n, mom.data_ptr(), ws.data_ptr(), ws.numel(),
                                       _lib.stream_ptr()))
    _lib.check(L.ppd_advantage_finalize(mom.data_ptr(), stats.data_ptr(), _lib.stream_ptr()))
    assert abs(stats[0].item() - x.double().mean().item()) < 1e-6
    assert abs(stats[1].item() - (x.double().std().item() + 1e-5)) < 1e-5


# --------------------------------------------------------------------------- gathers
NAMES = ("obs", "vector_obs", "recurrent_hidden_states", "actions", "value_preds", "returns",
         "masks", "old_action_log_probs", "adv_targ")


def _golden_roll(g):
    keys = ("obs", "vector_obs", "recurrent_hidden_states", "actions", "value_preds", "returns", "masks",
            "action_log_probs", "rewards", "bad_masks")
    return {k: g[k] for k in keys}


@pytest.mark.parametrize("fixture,nmb,slots", [("ff_gen_variant_b", 4, (0, 3, 4, 5, 6, 7, 8)),
                                               ("ff_gen_variant_a", 3, tuple(range(9)))])
def test_feed_forward_generator_bit_exact_vs_golden(golden, fixture, nmb, slots):
    g = golden(fixture)
    roll = _golden_roll(g)
    st = storage_from(roll, roll["obs"].shape[2:], roll["vector_obs"].shape[2], 4, roll["recurrent_hidden_states"].shape[2])
    torch.manual_seed(int(g["seed"]))
    mbs = list(st.feed_forward_generator(T_(g["advantages"]).to(DEV), nmb))
    assert len(mbs) == int(g["num_minibatches"])
    for k, mb in enumerate(mbs):
        for s in slots:
            assert torch.equal(mb[s].cpu(), T_(g[f"mb{k}_{NAMES[s]}"])), (k, NAMES[s])


def test_recurrent_generator_bit_exact_vs_golden(golden):
    g = golden("rec_gen")
    roll = _golden_roll(g)
    st = storage_from(roll, roll["obs"].shape[2:], 3, 4, 8)
    torch.manual_seed(int(g["seed"]))
    mbs = list(st.recurrent_generator(T_(g["advantages"]).to(DEV), 4))
    assert len(mbs) == int(g["num_minibatches"])
    for k, mb in enumerate(mbs):
        for s in range(9):
            assert torch.equal(mb[s].cpu(), T_(g[f"mb{k}_{NAMES[s]}"])), (k, NAMES[s])


@pytest.mark.parametrize("recurrent", [False, True])
def test_generators_vs_oracle_obstacle_tower_shape(recurrent):
    """84x84 obs (vectorised 16-byte path), V=15, H=512; fused advantage normalisation."""
    cfg = synthetic.RolloutConfig("t", 12, 8, 3, 15, 8, recurrent, 1, 4, 1e-4, 0.001)
    roll = synthetic.make_rollout(cfg, seed=7)
    roll["returns"] = torch.randn(13, 8, 1, generator=torch.Generator().manual_seed(1))
    st = storage_from(roll, (3, 84, 84), 15, 8, 512 if recurrent else 1)
    adv_cpu = T_(o_ret.normalized_advantages(roll["returns"].numpy(), roll["value_preds"].numpy()))
    # explicit advantages tensor: every slot bit-exact
    torch.manual_seed(5)
    want = list(o_mb.recurrent_minibatches(roll, adv_cpu, 4) if recurrent else o_mb.feed_forward_minibatches(roll, adv_cpu, 4))
    torch.manual_seed(5)
    got = list(st.recurrent_generator(adv_cpu.to(DEV), 4) if recurrent else st.feed_forward_generator(adv_cpu.to(DEV), 4))
    assert len(got) == len(want) == 4
    for a, b in zip(got, want):
        for s in range(9):
            assert torch.equal(a[s].cpu(), b[s]), NAMES[s]
    # fused advantages: computed from returns/value_preds + device stats
    L = _lib.lib()
    n = 12 * 8
    ws = torch.empty(L.ppd_advantage_moments_workspace(n), dtype=torch.uint8, device=DEV)
    mom = torch.zeros(3, dtype=torch.float64, device=DEV)
    stats = torch.zeros(2, device=DEV)
    _lib.check(L.ppd_advantage_moments(st.returns.data_ptr(), st.value_preds.data_ptr(), n, mom.data_ptr(), ws.data_ptr(),
                                       ws.numel(), _lib.stream_ptr()))
    _lib.check(L.ppd_advantage_finalize(mom.data_ptr(), stats.data_ptr(), _lib.stream_ptr()))
    torch.manual_seed(5)
    got = list(st.recurrent_generator(FusedAdvantages(stats), 4) if recurrent else st.feed_forward_generator(FusedAdvantages(stats), 4))
    for a, b in zip(got, want):
        for s in range(8):
            assert torch.equal(a[s].cpu(), b[s]), NAMES[s]
        assert torch.allclose(a[8].cpu(), b[8], rtol=1e-5, atol=1e-6)


def test_generator_edge_cases():
    cfg = synthetic.RolloutConfig("t", 5, 6, 1, 0, 4, False, 1, 4, 1e-4, 0.001)
    roll = synthetic.make_rollout(cfg, seed=9, obs_shape=(1, 3, 3))   # unaligned rows -> scalar copy path
    st = storage_from(roll, (1, 3, 3), 0, 4, 1)
    with pytest.raises(AssertionError):
        next(st.feed_forward_generator(None, 31))
    with pytest.raises(AssertionError):
        next(st.recurrent_generator(torch.zeros(5, 6, 1, device=DEV), 7))
    # ragged: 30 samples into 4 minibatches of 7, remainder dropped; advantages=None -> adv_targ None
    torch.manual_seed(1)
    got = list(st.feed_forward_generator(None, 4))
    torch.manual_seed(1)
    want = list(o_mb.feed_forward_minibatches(roll, None, 4))
    assert len(got) == len(want) == 4 and got[0][8] is None
    for a, b in zip(got, want):
        assert a[0].shape[0] == 7 and a[1].shape == (7, 0)
        for s in (0, 2, 3, 4, 5, 6, 7):
            assert torch.equal(a[s].cpu(), b[s])
    # mini_batch_size path (used by GAIL in the reference)
    torch.manual_seed(2)
    got = list(st.feed_forward_generator(None, mini_batch_size=8))
    assert len(got) == 3 and got[0][0].shape[0] == 8
    # N not divisible by num_mini_batch: the reference raises IndexError on the ragged tail block
    gen = st.recurrent_generator(torch.zeros(5, 6, 1, device=DEV), 4)   # E = 1 -> fine, 6 blocks
    assert len(list(gen)) == 6
    cfg5 = synthetic.RolloutConfig("t", 5, 5, 1, 0, 4, True, 1, 2, 1e-4, 0.001)
    roll5 = synthetic.make_rollout(cfg5, seed=9, obs_shape=(1, 3, 3), hidden_state_size=4)
    st5 = storage_from(roll5, (1, 3, 3), 0, 4, 4)
    with pytest.raises(IndexError):
        list(st5.recurrent_generator(torch.zeros(5, 5, 1, device=DEV), 2))


# --------------------------------------------------------------------------- fused loss
def _loss_case(B, A, seed, ties=True, clipped=True):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(B, A + 1, generator=g)
    z[:, :A] *= 0.5
    actions = torch.randint(0, A, (B, 1), generator=g)
    logp_now = torch.log_softmax(z[:, :A], -1).gather(1, actions)
    old_logp = logp_now + 0.15 * torch.randn(B, 1, generator=g)
    adv = torch.randn(B, 1, generator=g)
    old_v = z[:, A:A + 1] + 0.15 * torch.randn(B, 1, generator=g)
    ret = torch.randn(B, 1, generator=g)
    if ties:
        old_logp[::5] = logp_now[::5]          # ratio == 1 exactly: surr1 == surr2
        old_v[::7] = z[::7, A:A + 1]           # v - V_old == 0: e1 vs e2 decided by rounding
        adv[::11] = 0.0
    return z, actions, old_logp.detach(), adv, old_v.detach(), ret


@pytest.mark.parametrize("B,A", [(7, 8), (256, 8), (2048, 8), (1024, 54), (4099, 5)])
@pytest.mark.parametrize("clipped", [True, False])
def test_ppo_loss_fwd_bwd_vs_autograd(B, A, clipped):
    z, actions, old_logp, adv, old_v, ret = _loss_case(B, A, B + A)
    zc = z.clone().requires_grad_(True)
    d = torch.distributions.Categorical(logits=zc[:, :A])
    logp = d.log_prob(actions.squeeze(-1)).unsqueeze(-1)
    ent = d.entropy().mean()
    vl, al = o_upd.ppo_losses(zc[:, A:A + 1], logp, ent, old_v, ret, old_logp, adv, 0.1, clipped)
    (vl * 0.5 + al - ent * 0.01).backward()
    L = _lib.lib()
    zd = z.to(DEV).contiguous()
    dz = torch.empty_like(zd)
    lp_out = torch.empty(B, device=DEV)
    ent_out = torch.empty(B, device=DEV)
    loss = torch.zeros(3, device=DEV)
    ws = torch.empty(L.ppd_ppo_loss_workspace(B), dtype=torch.uint8, device=DEV)
    t = [x.to(DEV).contiguous() for x in (actions, old_logp, adv, old_v, ret)]
    _lib.check(L.ppd_ppo_loss_fwd_bwd(zd.data_ptr(), A + 1, A, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(),
                                      t[3].data_ptr(), t[4].data_ptr(), B, B, 0.1, 0.5, 0.01, int(clipped),
                                      dz.data_ptr(), lp_out.data_ptr(), ent_out.data_ptr(), loss.data_ptr(),
                                      ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    # stated tolerance: fp32, rtol 1e-5 (+ atol 1e-7 on gradients that are O(1/B))
    np.testing.assert_allclose(loss.cpu().numpy(), [vl.item(), al.item(), ent.item()], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(lp_out.cpu().numpy(), logp.detach().squeeze(-1).numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dz.cpu().numpy(), zc.grad.numpy(), rtol=1e-4, atol=2e-8)


def test_ppo_loss_global_rows_scaling():
    """Sharded minibatch: two halves with global_rows = B must sum to the single-GPU result."""
    B, A = 512, 8
    z, actions, old_logp, adv, old_v, ret = _loss_case(B, A, 5)
    L = _lib.lib()

    def run(sl, rows):
        n = sl.stop - sl.start
        t = [x[sl].to(DEV).contiguous() for x in (z, actions, old_logp, adv, old_v, ret)]
        dz = torch.empty_like(t[0])
        loss = torch.zeros(3, device=DEV)
        ws = torch.empty(L.ppd_ppo_loss_workspace(n), dtype=torch.uint8, device=DEV)
        _lib.check(L.ppd_ppo_loss_fwd_bwd(t[0].data_ptr(), A + 1, A, t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(),
                                          t[4].data_ptr(), t[5].data_ptr(), n, rows, 0.1, 0.5, 0.01, 1, dz.data_ptr(),
                                          None, None, loss.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        return dz.cpu(), loss.cpu()
    dz_all, loss_all = run(slice(0, B), B)
    dz_a, loss_a = run(slice(0, B // 2), B)
    dz_b, loss_b = run(slice(B // 2, B), B)
    assert torch.allclose(torch.cat([dz_a, dz_b]), dz_all, rtol=0, atol=0)
    assert torch.allclose(loss_a + loss_b, loss_all, rtol=1e-6, atol=1e-8)


def test_categorical_eval():
    B, A = 300, 54
    g = torch.Generator().manual_seed(0)
    z = torch.randn(B, A, generator=g)
    actions = torch.randint(0, A, (B, 1), generator=g)
    d = torch.distributions.Categorical(logits=z)
    L = _lib.lib()
    zd, ad = z.to(DEV), actions.to(DEV)
    lp = torch.empty(B, device=DEV); ent = torch.empty(B, device=DEV)
    mode = torch.empty(B, dtype=torch.int64, device=DEV); probs = torch.empty(B, A, device=DEV)
    _lib.check(L.ppd_categorical_eval(zd.data_ptr(), A, A, ad.data_ptr(), B, lp.data_ptr(), ent.data_ptr(),
                                      mode.data_ptr(), probs.data_ptr(), _lib.stream_ptr()))
    np.testing.assert_allclose(lp.cpu().numpy(), d.log_prob(actions.squeeze(-1)).numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ent.cpu().numpy(), d.entropy().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(probs.cpu().numpy(), d.probs.numpy(), rtol=1e-5, atol=1e-7)
    assert torch.equal(mode.cpu(), d.probs.argmax(-1))


# --------------------------------------------------------------------------- clip + Adam
@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("n,max_norm", [(1000, 0.5), (2464393, 0.5), (4097, 1e9), (33, 0.0)])
def test_clip_adam_vs_torch(n, max_norm, fused):
    _lib.lib().ppd_clip_adam_set_fused(fused)
    try:
        _clip_adam_case(n, max_norm)
    finally:
        _lib.lib().ppd_clip_adam_set_fused(1)


def _clip_adam_case(n, max_norm):
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g) * 0.05
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=1e-4, eps=1e-5)
    L = _lib.lib()
    p = p0.to(DEV); m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
    ws = torch.empty(L.ppd_clip_adam_workspace(n), dtype=torch.uint8, device=DEV)
    gn = torch.zeros(1, device=DEV)
    loss_in = torch.tensor([1.0, 2.0, 3.0], device=DEV); loss_acc = torch.zeros(3, device=DEV)
    for step in range(1, 6):
        grad = torch.randn(n, generator=g) * (10.0 ** -(step % 3)) / np.sqrt(n)
        ref_p.grad = grad.clone()
        if max_norm > 0:
            want_norm = torch.nn.utils.clip_grad_norm_([ref_p], max_norm)
        else:
            want_norm = grad.norm()
        opt.step()
        gd = grad.to(DEV)
        _lib.check(L.ppd_clip_adam_step(p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, step, 1e-4, 0.9,
                                        0.999, 1e-5, max_norm, gn.data_ptr(), loss_in.data_ptr(), loss_acc.data_ptr(),
                                        ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        # torch's fp32 CPU norm is itself ~4e-5 off the float64 value at 2.4M elements; the kernel
        # (fp32 per thread, fp64 across threads) is checked against float64, and loosely against torch
        exact = grad.double().norm().item()
        assert abs(gn.item() - exact) <= 1e-6 * max(1.0, exact)
        assert abs(gn.item() - want_norm.item()) <= 2e-4 * max(1.0, want_norm.item())
        # stated tolerance: |dp| per step <= lr = 1e-4; agree to 1e-8 absolute (1e-4 of a step) + 1e-6 relative
        # (torch's clip coefficient inherits the ~4e-5 error of its fp32 norm)
        np.testing.assert_allclose(p.cpu().numpy(), ref_p.detach().numpy(), rtol=1e-6, atol=1e-8)
    assert torch.allclose(loss_acc.cpu(), torch.tensor([5.0, 10.0, 15.0]))
    st = opt.state[ref_p]
    # the moments see torch's clip coefficient, which carries the ~4e-5 error of torch's fp32 norm
    np.testing.assert_allclose(m.cpu().numpy(), st["exp_avg"].numpy(), rtol=2e-4, atol=5e-9)
    np.testing.assert_allclose(v.cpu().numpy(), st["exp_avg_sq"].numpy(), rtol=4e-4, atol=1e-14)


# --------------------------------------------------------------------------- obs running norm
def test_obs_rms_update_normalize_vs_oracle():
    rng = np.random.default_rng(0)
    N, shape = 32, (3, 84, 84)
    F = int(np.prod(shape))
    rms = o_rms.RunningMoments(shape=shape)
    L = _lib.lib()
    mean = torch.zeros(F, dtype=torch.float64, device=DEV)
    var = torch.ones(F, dtype=torch.float64, device=DEV)
    count = 1e-4
    for it in range(3):
        obs = (rng.normal(1.5, 2.0, size=(N,) + shape) * (1 + it)).astype(np.float32)
        want = o_rms.obs_filter(rms, obs, clipob=1.0 if it == 1 else 10.0, update=True)
        od = torch.from_numpy(obs).to(DEV).reshape(N, F).contiguous()
        out = torch.empty_like(od)
        _lib.check(L.ppd_obs_rms_update_normalize(od.data_ptr(), N, F, mean.data_ptr(), var.data_ptr(), count, 1, 1e-8,
                                                  1.0 if it == 1 else 10.0, out.data_ptr(), _lib.stream_ptr()))
        count += N
        np.testing.assert_allclose(mean.cpu().numpy().reshape(shape), rms.mean, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(var.cpu().numpy().reshape(shape), rms.var, rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(out.cpu().numpy().reshape((N,) + shape), want.astype(np.float32), rtol=1e-6, atol=1e-6)
    # eval mode: no update
    obs = rng.normal(size=(N,) + shape).astype(np.float32)
    want = o_rms.obs_filter(rms, obs, update=False)
    m0 = mean.clone()
    od = torch.from_numpy(obs).to(DEV).reshape(N, F).contiguous()
    out = torch.empty_like(od)
    _lib.check(L.ppd_obs_rms_update_normalize(od.data_ptr(), N, F, mean.data_ptr(), var.data_ptr(), count, 0, 1e-8, 10.0,
                                              out.data_ptr(), _lib.stream_ptr()))
    assert torch.equal(mean, m0)
    np.testing.assert_allclose(out.cpu().numpy().reshape((N,) + shape), want.astype(np.float32), rtol=1e-6, atol=1e-6)

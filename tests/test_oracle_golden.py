"""CPU: the oracle restatement against fixtures produced by the real reference
(tests/golden/make_golden.py).  Pins oracle/ -- see oracle/__init__.py."""
import numpy as np
import pytest
import torch

from oracle import minibatch, policy, ppo_update, returns, running_mean_std

torch.set_num_threads(2)


def T_(x):
    return torch.as_tensor(np.asarray(x))


@pytest.mark.parametrize("use_gae", [True, False])
@pytest.mark.parametrize("proper", [True, False])
def test_returns_bit_exact(golden, use_gae, proper):
    g = golden("returns")
    sentinel = np.full_like(g["value_preds"], -7.0)
    ret, v = returns.returns_recurrence(g["rewards"], g["value_preds"], g["masks"], g["bad_masks"],
                                        g["next_value"], use_gae, float(g["gamma"]), float(g["gae_lambda"]),
                                        proper, returns_in=sentinel)
    tag = f"gae{int(use_gae)}_proper{int(proper)}"
    assert np.array_equal(ret, g["returns_" + tag])
    assert np.array_equal(v, g["value_preds_" + tag])


def test_advantage_normalisation(golden):
    g = golden("returns")
    adv = returns.normalized_advantages(g["returns_gae1_proper0"], g["value_preds_gae1_proper0"])
    assert np.array_equal(adv, g["adv_norm"])


NAMES = ("obs", "vector_obs", "recurrent_hidden_states", "actions", "value_preds", "returns",
         "masks", "old_action_log_probs", "adv_targ")


def _roll(g):
    return {k: T_(g[k]) for k in ("obs", "vector_obs", "recurrent_hidden_states", "actions", "value_preds",
                                  "returns", "masks", "action_log_probs", "rewards")}


@pytest.mark.parametrize("fixture,nmb,slots", [("ff_gen_variant_b", 4, (0, 3, 4, 5, 6, 7, 8)),
                                               ("ff_gen_variant_a", 3, tuple(range(9)))])
def test_feed_forward_minibatches_bit_exact(golden, fixture, nmb, slots):
    g = golden(fixture)
    torch.manual_seed(int(g["seed"]))
    mbs = list(minibatch.feed_forward_minibatches(_roll(g), T_(g["advantages"]), nmb))
    assert len(mbs) == int(g["num_minibatches"])
    for k, mb in enumerate(mbs):
        for s in slots:
            assert torch.equal(mb[s], T_(g[f"mb{k}_{NAMES[s]}"])), (k, NAMES[s])


def test_recurrent_minibatches_bit_exact(golden):
    g = golden("rec_gen")
    torch.manual_seed(int(g["seed"]))
    mbs = list(minibatch.recurrent_minibatches(_roll(g), T_(g["advantages"]), 4))
    assert len(mbs) == int(g["num_minibatches"])
    for k, mb in enumerate(mbs):
        for s in range(9):
            assert torch.equal(mb[s], T_(g[f"mb{k}_{NAMES[s]}"])), (k, NAMES[s])


def _params(g, prefix="param."):
    return {k[len(prefix):]: T_(g[k]) for k in g.files if k.startswith(prefix)}


def _grads_of(p, g, recurrent, concat):
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    v, lp, ent, hx = policy.evaluate_actions(p, T_(g["obs"]), T_(g["vobs"]), T_(g["h0"]), T_(g["masks"]),
                                             T_(g["actions"]), recurrent, concat)
    vl, al = ppo_update.ppo_losses(v, lp, ent, T_(g["old_v"]), T_(g["ret"]), T_(g["old_logp"]), T_(g["adv"]),
                                   float(g["clip"]))
    (vl * float(g["vcoef"]) + al - ent * float(g["ecoef"])).backward()
    return (v, lp, ent, hx, vl, al), {k: t.grad for k, t in p.items()}


@pytest.mark.parametrize("fixture,recurrent,concat", [("policy_recurrent", True, True),
                                                      ("policy_feedforward", False, False)])
def test_policy_forward_and_grads(golden, fixture, recurrent, concat):
    g = golden(fixture)
    (v, lp, ent, hx, vl, al), grads = _grads_of(_params(g), g, recurrent, concat)
    tol = dict(rtol=1e-5, atol=1e-6)
    assert torch.allclose(v, T_(g["out.values"]), **tol)
    assert torch.allclose(lp, T_(g["out.logp"]), **tol)
    assert torch.allclose(ent, T_(g["out.entropy"]), **tol)
    assert torch.allclose(hx, T_(g["out.hxs"]), **tol)
    assert torch.allclose(vl, T_(g["out.value_loss"]), **tol)
    assert torch.allclose(al, T_(g["out.action_loss"]), **tol)
    for k, gr in grads.items():
        ref = T_(g["grad." + k])
        assert torch.allclose(gr, ref, rtol=1e-4, atol=1e-7), (k, (gr - ref).abs().max())


def test_policy_act_and_value(golden):
    g = golden("policy_recurrent")
    p = _params(g)
    E = int(g["E"])
    with torch.no_grad():
        v, a, lp, h = policy.act(p, T_(g["obs"])[:E], T_(g["vobs"])[:E], T_(g["h0"]), T_(g["act_masks"]), True,
                                 deterministic=True)
        gv = policy.get_value(p, T_(g["obs"])[:E], T_(g["vobs"])[:E], T_(g["h0"]), T_(g["act_masks"]), True)
    assert torch.equal(a, T_(g["act_action"]))
    assert torch.allclose(v, T_(g["act_value"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(lp, T_(g["act_logp"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(h, T_(g["act_hxs"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(gv, T_(g["get_value"]), rtol=1e-5, atol=1e-6)


def test_gru_stepwise_equals_segmented(golden):
    g = golden("policy_recurrent")
    p = _params(g)
    x = torch.cat((policy.trunk(p, T_(g["obs"])), T_(g["vobs"])), 1)
    a, ha = policy.gru_with_resets(p, x, T_(g["h0"]), T_(g["masks"]))
    b, hb = policy.gru_cell_stepwise(p, x, T_(g["h0"]), T_(g["masks"]))
    assert torch.allclose(a, b, atol=1e-6) and torch.allclose(ha, hb, atol=1e-6)


@pytest.mark.parametrize("tag,recurrent,concat", [("recurrent", True, True), ("feedforward", False, False)])
def test_ppo_update(golden, tag, recurrent, concat):
    g = golden("update_" + tag)
    state = ppo_update.UpdateState(_params(g, "init."), lr=float(g["lr"]), eps=float(g["eps"]))
    roll = {k[5:]: T_(g[k]) for k in g.files if k.startswith("roll.")}
    torch.manual_seed(int(g["seed"]))
    out = ppo_update.ppo_update(state, roll, recurrent=recurrent, clip_param=float(g["clip"]),
                                ppo_epoch=int(g["epochs"]), num_mini_batch=int(g["nmb"]),
                                value_loss_coef=float(g["vcoef"]), entropy_coef=float(g["ecoef"]),
                                max_grad_norm=float(g["max_grad_norm"]), concat_vector=concat)
    assert np.allclose(np.array(out), g["losses"], rtol=1e-5, atol=1e-7), (out, g["losses"])
    for k, v in state.params.items():
        ref = T_(g["final." + k])
        assert torch.allclose(v.detach(), ref, rtol=0, atol=2e-6), (k, (v.detach() - ref).abs().max())


@pytest.mark.parametrize("tag,C,V,A,recurrent,concat", [("c2", 3, 15, 8, True, True), ("c1", 1, 0, 8, False, False)])
def test_full_size_init_and_forward(golden, tag, C, V, A, recurrent, concat):
    g = golden("init_full")
    torch.manual_seed(0)
    p = policy.init_params(C, A, V, recurrent, 512, concat)
    for k, v in p.items():
        # same RNG stream as Policy(...); the QR inside orthogonal_ rounds differently with the
        # thread count, hence 1e-3 absolute on sums of ~1e6 elements (a different stream is off by O(10))
        assert np.isclose(float(v.double().sum()), float(g[f"{tag}.sum.{k}"]), rtol=0, atol=1e-3), k
        assert np.isclose(float(v.double().abs().sum()), float(g[f"{tag}.abssum.{k}"]), rtol=1e-6), k
    gen = torch.Generator().manual_seed(51)
    E, Tn = 2, 3
    Bn = E * Tn
    obs = torch.randn(Bn, C, 84, 84, generator=gen)
    vobs = torch.rand(Bn, V, generator=gen)
    h0 = 0.1 * torch.randn(E if recurrent else Bn, 512 if recurrent else 1, generator=gen)
    masks = torch.ones(Bn, 1)
    masks[3] = 0.0
    actions = torch.randint(0, A, (Bn, 1), generator=gen)
    with torch.no_grad():
        v, lp, ent, hx = policy.evaluate_actions(p, obs, vobs, h0, masks, actions, recurrent, concat)
    assert torch.allclose(v, T_(g[f"{tag}.value"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(lp, T_(g[f"{tag}.logp"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(ent, T_(g[f"{tag}.entropy"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(hx, T_(g[f"{tag}.hxs"]), rtol=1e-5, atol=1e-6)


def test_running_moments_pooled_property():
    rng = np.random.default_rng(0)
    rms = running_mean_std.RunningMoments(shape=(3, 4))
    chunks = [rng.normal(2.0, 3.0, size=(n, 3, 4)).astype(np.float32) for n in (5, 1, 17, 8)]
    for c in chunks:
        rms.update(c)
    allx = np.concatenate(chunks, 0).astype(np.float64)
    n = allx.shape[0]
    w = 1e-4
    mean = allx.sum(0) / (n + w)                                   # pseudo-batch: weight 1e-4, mean 0, var 1
    ex2 = ((allx ** 2).sum(0) + w * 1.0) / (n + w)
    assert np.allclose(rms.mean, mean, rtol=1e-12, atol=1e-12)
    assert np.allclose(rms.var, ex2 - mean ** 2, rtol=1e-10, atol=1e-12)
    assert np.isclose(rms.count, n + w)
    out = running_mean_std.obs_filter(rms, chunks[0], clipob=1.0, update=False)
    assert out.min() >= -1.0 and out.max() <= 1.0


# --------------------------------------------------------------------------- uint8 observation pipeline
def _obs_pipeline_inputs(g):
    T, N, H, W, C, nstack = (int(x) for x in g["shape"])
    rng = np.random.RandomState(int(g["seed"]))
    frames = rng.randint(0, 256, size=(T + 1, N, H, W, C), dtype=np.uint8)
    dones = rng.rand(T, N) < 0.3
    dones[1, 0] = True
    assert np.array_equal(dones, g["dones"])
    return frames, dones, nstack


def test_obs_pipeline_oracle_is_bit_exact_with_reference_wrappers(golden):
    """oracle/obs_pipeline.py against the outputs of the reference's NormalizeWrapper / TransposeImage / VecPyTorch cast /
    VecPyTorchFrameStack (tests/golden/make_golden_obs.py): sha256 of the float32 bytes, and the stored slots element for element."""
    import hashlib
    from oracle import obs_pipeline as o_obs
    g = golden("obs_pipeline")
    frames, dones, nstack = _obs_pipeline_inputs(g)
    sha = lambda t: hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()
    norm = o_obs.rollout_observations(frames, dones, 1, g["mean"], g["std"])
    assert np.array_equal(norm[:1].numpy(), g["head_norm"])
    assert sha(norm) == str(g["sha_norm"])
    div = o_obs.rollout_observations(frames, dones, 1)
    assert np.array_equal(div[:1].numpy(), g["head_div255"])
    assert sha(div) == str(g["sha_div255"])
    st = o_obs.rollout_observations(frames, dones, nstack, g["mean"], g["std"])
    assert np.array_equal(st[2, :1].numpy(), g["head_stack"])
    assert sha(st) == str(g["sha_stack"])


# --------------------------------------------------------------------------- A2C
def test_a2c_update_oracle_vs_reference(golden):
    """oracle/a2c_update.py against two consecutive updates of the reference's A2C_ACKTR (tests/golden/make_golden_a2c.py)."""
    from oracle import a2c_update as o_a2c
    g = golden("update_a2c")
    init = {k[5:]: torch.as_tensor(g[k]) for k in g.files if k.startswith("init.")}
    roll = {k[5:]: torch.as_tensor(g[k]) for k in g.files if k.startswith("roll.")}
    st = o_a2c.A2CState(init, float(g["lr"]), float(g["eps"]), float(g["alpha"]))
    kw = dict(recurrent=True, value_loss_coef=float(g["vcoef"]), entropy_coef=float(g["ecoef"]), max_grad_norm=float(g["max_grad_norm"]))
    out1 = o_a2c.a2c_update(st, roll, **kw)
    np.testing.assert_allclose(np.array(out1), g["losses1"], rtol=1e-6, atol=1e-8)
    for k, v in st.params.items():
        # (the fixture was made with one thread, run.py:55; this process may use several: the convolutions' summation order differs,
        #  and RMSprop's g / (sqrt(v) + eps) amplifies that on entries with a gradient near eps -- 2e-3 of lr = 1.4e-6 absolute)
        np.testing.assert_allclose(v.detach().numpy(), g["mid." + k], rtol=1e-6, atol=1.4e-6, err_msg=k)
    out2 = o_a2c.a2c_update(st, roll, **kw)
    np.testing.assert_allclose(np.array(out2), g["losses2"], rtol=1e-5, atol=1e-7)
    for k, v in st.params.items():
        np.testing.assert_allclose(v.detach().numpy(), g["final." + k], rtol=1e-5, atol=3e-6, err_msg=k)

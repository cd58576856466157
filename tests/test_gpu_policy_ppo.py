"""GPU parity of the drop-in classes (Policy / PPO / RolloutStorage) against the golden fixtures
recorded from the reference and against the oracle on seeded synthetic rollouts."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import policy as o_pol  # noqa: E402
from oracle import ppo_update as o_upd  # noqa: E402
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


def params_of(g, prefix):
    return {k[len(prefix):]: T_(g[k]) for k in g.files if k.startswith(prefix)}


def make_policy(C, A, V, recurrent, H, state=None, precision=None):
    pol = ppd.Policy((C, 84, 84), Discrete(A), base_kwargs={"recurrent": recurrent, "hidden_size": H}, vector_obs_len=V)
    if state is not None:
        missing, unexpected = pol.load_state_dict(state, strict=True)
        assert not missing and not unexpected
    pol = pol.to(DEV)
    if precision:
        pol.engine(precision)
    return pol


PARITY_MODES = ["fp32", "tf32x3"]      # both must meet the fp32 parity gate


@pytest.mark.parametrize("precision", PARITY_MODES)
@pytest.mark.parametrize("fixture,recurrent,C,A,V", [("policy_recurrent", True, 2, 5, 3), ("policy_feedforward", False, 1, 6, 0)])
def test_policy_forward_matches_reference(golden, fixture, recurrent, C, A, V, precision):
    g = golden(fixture)
    pol = make_policy(C, A, V, recurrent, 32, params_of(g, "param."), precision)
    d = lambda k: T_(g[k]).to(DEV)
    v, lp, ent, hx = pol.evaluate_actions(d("obs"), d("vobs"), d("h0"), d("masks"), d("actions"))
    tol = dict(rtol=1e-5, atol=2e-6)     # stated fp32 tolerance (1e-5 relative, + 2e-6 absolute near zero)
    np.testing.assert_allclose(v.cpu().numpy(), g["out.values"], **tol)
    np.testing.assert_allclose(lp.cpu().numpy(), g["out.logp"], **tol)
    np.testing.assert_allclose(ent.item(), float(g["out.entropy"]), **tol)
    np.testing.assert_allclose(hx.cpu().numpy(), g["out.hxs"], **tol)
    if recurrent:
        E = int(g["E"])
        v, a, lp, h = pol.act(d("obs")[:E], d("vobs")[:E], d("h0"), d("act_masks"), deterministic=True)
        assert torch.equal(a.cpu(), T_(g["act_action"]))
        np.testing.assert_allclose(v.cpu().numpy(), g["act_value"], **tol)
        np.testing.assert_allclose(lp.cpu().numpy(), g["act_logp"], **tol)
        np.testing.assert_allclose(h.cpu().numpy(), g["act_hxs"], **tol)
        gv = pol.get_value(d("obs")[:E], d("vobs")[:E], d("h0"), d("act_masks"))
        np.testing.assert_allclose(gv.cpu().numpy(), g["get_value"], **tol)
        # stochastic act: shapes / dtypes / log-prob consistency (sampling uses the device RNG)
        v, a, lp, h = pol.act(d("obs")[:E], d("vobs")[:E], d("h0"), d("act_masks"))
        assert a.shape == (E, 1) and a.dtype == torch.int64 and lp.shape == (E, 1) and h.shape == (E, 32)


@pytest.mark.parametrize("precision", PARITY_MODES)
@pytest.mark.parametrize("fixture,recurrent,C,A,V", [("policy_recurrent", True, 2, 5, 3), ("policy_feedforward", False, 1, 6, 0)])
def test_minibatch_gradients_match_reference_autograd(golden, fixture, recurrent, C, A, V, precision):
    g = golden(fixture)
    pol = make_policy(C, A, V, recurrent, 32, params_of(g, "param."), precision)
    eng = pol.engine()
    d = lambda k: T_(g[k]).to(DEV)
    sample = (d("obs"), d("vobs"), d("h0"), d("actions"), d("old_v"), d("ret"), d("masks"), d("old_logp"), d("adv"))
    eng.train_minibatch(sample, float(g["clip"]), float(g["vcoef"]), float(g["ecoef"]))
    loss = eng.flat_grad[eng.loss_off:eng.loss_off + 3].cpu().numpy()
    np.testing.assert_allclose(loss[0], float(g["out.value_loss"]), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(loss[1], float(g["out.action_loss"]), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(loss[2], float(g["out.entropy"]), rtol=1e-5, atol=1e-7)
    for name, p in pol.named_parameters():
        ref = g["grad." + name]
        got = p.grad.cpu().numpy()
        scale = max(1e-6, float(np.abs(ref).max()))
        # stated tolerance: 1e-5 relative to the largest gradient entry of the tensor + 1e-4 elementwise relative
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5 * scale, err_msg=name)


@pytest.mark.parametrize("precision", PARITY_MODES)
@pytest.mark.parametrize("tag,recurrent", [("recurrent", True), ("feedforward", False)])
def test_ppo_update_matches_reference(golden, tag, recurrent, precision):
    g = golden("update_" + tag)
    C, V, A, H, T, N = (int(g[k]) for k in ("C", "V", "A", "H", "T", "N"))
    pol = make_policy(C, A, V, recurrent, H, params_of(g, "init."), precision)
    st = ppd.RolloutStorage(T, N, (C, 84, 84), [V], Discrete(A), H if recurrent else 1)
    for k in ppd.RolloutStorage._FIELDS:
        getattr(st, k).copy_(T_(g["roll." + k]))
    st.to(DEV)
    agent = ppd.algo.PPO(pol, float(g["clip"]), int(g["epochs"]), int(g["nmb"]), float(g["vcoef"]), float(g["ecoef"]),
                         lr=float(g["lr"]), eps=float(g["eps"]), max_grad_norm=float(g["max_grad_norm"]))
    torch.manual_seed(int(g["seed"]))
    out = agent.update(st)
    assert all(isinstance(x, float) for x in out)
    np.testing.assert_allclose(np.array(out), g["losses"], rtol=1e-4, atol=1e-6)
    lr = float(g["lr"])
    for name, p in pol.state_dict().items():
        ref = g["final." + name]
        # stated tolerance: parameters move by at most lr per Adam step and Adam's m/sqrt(v) is ill-conditioned for
            # entries whose gradient is ~eps; after epochs*nmb steps they agree to 5% of ONE step (observed worst
            # case: 1 entry in 32768 at 2% of a step)
        np.testing.assert_allclose(p.cpu().numpy(), ref, rtol=0, atol=0.05 * lr, err_msg=name)


@pytest.mark.parametrize("precision", PARITY_MODES + ["tf32"])
def test_c2_shaped_minibatch_vs_oracle(precision):
    """PPO-Dash full shapes (C=3, V=15, A=8, H=512, recurrent), T=24 x E=4 rows, resets inside."""
    torch.manual_seed(0)
    pol = ppd.Policy((3, 84, 84), Discrete(8), base_kwargs={"recurrent": True}, vector_obs_len=15)
    p_cpu = {k: v.clone() for k, v in pol.state_dict().items()}
    pol = pol.to(DEV)
    pol.engine(precision)
    # stated tolerances.  fp32: rtol 1e-5 (+2e-6 abs near zero).  tf32x3: 1e-5 relative to the tensor's scale
    # (max |ref|): the hi/lo split drops the lo*lo term and 3 bits of lo, i.e. ~2^-21 per product instead of
    # fp32's 2^-24.  tf32 (single pass, truncated 10-bit mantissas, errors do not cancel in long sums):
    # direction of every gradient tensor within cos >= 0.995 and entries within 10% of the tensor's scale.
    loose = precision == "tf32"
    atol_out = {"fp32": 2e-6, "tf32x3": 1e-5, "tf32": 5e-3}[precision]
    T, E = 24, 4
    cfg = synthetic.RolloutConfig("t", T, E, 3, 15, 8, True, 1, 1, 1e-4, 0.001)
    roll = synthetic.make_rollout(cfg, seed=3, reset_prob=0.05)
    B = T * E
    obs = roll["obs"][:T].reshape(B, 3, 84, 84)
    vobs = roll["vector_obs"][:T].reshape(B, 15)
    h0 = roll["recurrent_hidden_states"][0]
    masks = roll["masks"][:T].reshape(B, 1)
    actions = roll["actions"].reshape(B, 1)
    gen = torch.Generator().manual_seed(1)
    old_v = 0.1 * torch.randn(B, 1, generator=gen); ret = 0.3 * torch.randn(B, 1, generator=gen)
    adv = torch.randn(B, 1, generator=gen)
    pr = {k: v.clone().requires_grad_(True) for k, v in p_cpu.items()}
    v, lp, ent, hx = o_pol.evaluate_actions(pr, obs, vobs, h0, masks, actions, True, True)
    old_logp = (lp + 0.05 * torch.randn(B, 1, generator=gen)).detach()
    vl, al = o_upd.ppo_losses(v, lp, ent, old_v, ret, old_logp, adv, 0.1)
    (vl * 0.5 + al - ent * 0.001).backward()
    eng = pol.engine()
    dd = lambda t: t.to(DEV)
    out = eng.train_minibatch((dd(obs), dd(vobs), dd(h0), dd(actions), dd(old_v), dd(ret), dd(masks), dd(old_logp), dd(adv)),
                              0.1, 0.5, 0.001)
    np.testing.assert_allclose(out["value"].cpu().numpy(), v.detach().numpy(), rtol=1e-2 if loose else 1e-5, atol=atol_out)
    np.testing.assert_allclose(out["rnn_hxs"].cpu().numpy(), hx.detach().numpy(), rtol=1e-2 if loose else 1e-5, atol=atol_out)
    loss = eng.flat_grad[eng.loss_off:eng.loss_off + 3].cpu().numpy()
    np.testing.assert_allclose(loss, [vl.item(), al.item(), ent.item()], rtol=1e-2 if loose else 1e-5, atol=1e-7)
    for name, p in pol.named_parameters():
        ref = pr[name].grad.numpy()
        scale = max(1e-6, float(np.abs(ref).max()))
        got = p.grad.cpu().numpy()
        if loose:
            cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
            assert cos >= 0.995 and float(np.abs(got - ref).max()) <= 0.1 * scale, (name, cos)
        else:
            if precision == "fp32":
                np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5 * scale, err_msg=name)
            else:
                # A ~1e-6 difference in a pre-activation that sits at zero flips its ReLU mask, which changes the
                # gradient of every weight feeding that unit by O(1/rows) -- a discrete jump, not rounding.  So:
                # at least 95% of the entries (one flipped unit touches a whole row/channel: 1/32 of a bias) meet the
                # 1e-5 gate and none is off by more than 5e-4 of the scale.
                err = np.abs(got - ref)
                ok = err <= (1e-4 * np.abs(ref) + 2e-5 * scale)
                assert ok.mean() >= 0.95 and float(err.max()) <= 5e-4 * scale, (name, ok.mean(), err.max() / scale)
    # the same minibatch cut into time chunks (GRU chunks on their own stream, pipelined against the trunk),
    # with and without keeping the im2col matrices for backward, gives the same gradients
    g1 = eng.flat_grad.clone()
    sample = (dd(obs), dd(vobs), dd(h0), dd(actions), dd(old_v), dd(ret), dd(masks), dd(old_logp), dd(adv))
    for chunk_rows, time_chunks, budget in ((40, 4, 6 << 30), (2048, 3, 6 << 30), (32, 4, 0)):
        eng.chunk_rows, eng.time_chunks, eng.cols_budget, eng.overlap_gru = chunk_rows, time_chunks, budget, True
        # (the implicit-GEMM convolutions of the tf32x3 mode have no im2col scratch to bound: only the time chunks cut there)
        assert len(eng._chunks(B, E)) >= (1 if precision == "tf32x3" else 3)
        eng.train_minibatch(sample, 0.1, 0.5, 0.001)
        torch.cuda.synchronize()
        np.testing.assert_allclose(eng.flat_grad.cpu().numpy(), g1.cpu().numpy(), rtol=1e-3 if loose else 1e-4,
                                   atol=1e-5 if loose else 2e-7)


def test_state_dict_roundtrip_and_rebind():
    torch.manual_seed(1)
    pol = ppd.Policy((1, 84, 84), Discrete(8), base_kwargs={"recurrent": True, "hidden_size": 32}, vector_obs_len=2)
    ref = {k: v.clone() for k, v in pol.state_dict().items()}
    pol = pol.to(DEV)
    obs = torch.randn(4, 1, 84, 84, device=DEV); vo = torch.rand(4, 2, device=DEV)
    h = torch.zeros(4, 32, device=DEV); m = torch.ones(4, 1, device=DEV)
    v0 = pol.get_value(obs, vo, h, m).clone()
    sd = pol.state_dict()                       # parameters are now views into the flat buffer
    assert set(sd) == set(ref)
    for k in ref:
        assert sd[k].shape == ref[k].shape and torch.equal(sd[k].cpu(), ref[k]), k
    import io
    buf = io.BytesIO(); torch.save(pol, buf); buf.seek(0)          # run.py:259 pickles the whole module
    pol2 = torch.load(buf, weights_only=False)
    assert torch.equal(pol2.get_value(obs, vo, h, m), v0)
    pol.cpu(); pol.to(DEV)                       # moving the module drops the views; the engine re-binds
    assert torch.equal(pol.get_value(obs, vo, h, m), v0)


def test_early_gradient_bucket_is_final_when_handed_over():
    """Data-parallel overlap (engine.train_minibatch(grad_ready=...)): the bucket [fc.w, end) is handed over before the convolution
    backward runs.  A snapshot taken on the hand-over stream must already equal the final gradient, the two buckets must tile the
    buffer, and the gradients must be bit-identical to the run without the callback."""
    torch.manual_seed(0)
    pol = ppd.Policy((3, 84, 84), Discrete(8), base_kwargs={"recurrent": True}, vector_obs_len=15).to(DEV)
    eng = pol.engine("tf32x3")
    T, E = 16, 4
    cfg = synthetic.RolloutConfig("t", T, E, 3, 15, 8, True, 1, 1, 1e-4, 0.001)
    roll = synthetic.make_rollout(cfg, seed=3, reset_prob=0.05)
    B = T * E
    gen = torch.Generator().manual_seed(1)
    dd = lambda t: t.to(DEV)
    sample = (dd(roll["obs"][:T].reshape(B, 3, 84, 84)), dd(roll["vector_obs"][:T].reshape(B, 15)), dd(roll["recurrent_hidden_states"][0]),
              dd(roll["actions"].reshape(B, 1)), dd(0.1 * torch.randn(B, 1, generator=gen)), dd(0.3 * torch.randn(B, 1, generator=gen)),
              dd(roll["masks"][:T].reshape(B, 1)), dd(roll["action_log_probs"].reshape(B, 1)), dd(torch.randn(B, 1, generator=gen)))
    eng.train_minibatch(sample, 0.1, 0.5, 0.001)
    torch.cuda.synchronize()
    want = eng.flat_grad.clone()
    for comm_ctas in (0, 8):
        # comm_ctas > 0: after the hand-over the convolution backward runs on 148 - comm_ctas CTAs (SMs left to the collective) and its
        # split-K plans follow the cap, so that bucket is re-associated (rounding-level differences); 0: the very same launches
        eng.comm_ctas = comm_ctas
        snaps = []

        def grad_ready(lo, hi):
            snaps.append((lo, hi, eng.flat_grad[lo:hi].clone()))          # on the stream the range became final on
        eng.train_minibatch(sample, 0.1, 0.5, 0.001, grad_ready=grad_ready)
        torch.cuda.synchronize()
        final = eng.flat_grad.clone()
        assert len(snaps) == 2
        (lo1, hi1, s1), (lo0, hi0, s0) = snaps
        assert lo1 == eng.segs["fc.w"].off and hi1 == eng.flat_grad.numel() and lo0 == 0 and hi0 == lo1
        assert torch.equal(s1, final[lo1:hi1]) and torch.equal(s0, final[lo0:hi0])          # final when handed over
        assert torch.equal(final[lo1:hi1], want[lo1:hi1])                                   # the early bucket: same launches either way
        if comm_ctas == 0:
            assert torch.equal(final, want)
        else:
            for name in ("conv1.w", "conv2.w", "conv3.w"):
                sg = eng.segs[name]
                a, b = final[sg.off:sg.off + sg.numel], want[sg.off:sg.off + sg.numel]
                assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()), name
        assert (hi1 - lo1) / eng.flat_grad.numel() > 0.95


def test_update_with_prefetched_gathers_is_bit_identical():
    """PPO.update gathers minibatch i+1 on a side stream (into two buffers it owns) while minibatch i trains; that must not change a
    single bit of the update: same losses and same parameters as with the gathers on the main stream."""
    T, N = 32, 8
    cfg = synthetic.RolloutConfig("t", T, N, 3, 15, 8, True, 2, 4, 1e-4, 0.001)
    roll = synthetic.make_rollout(cfg, seed=5, reset_prob=0.05)
    results = []
    for prefetch in (True, False, True):
        torch.manual_seed(0)
        pol = ppd.Policy((3, 84, 84), Discrete(8), base_kwargs={"recurrent": True}, vector_obs_len=15).to(DEV)
        pol.engine("tf32x3")
        st = ppd.RolloutStorage(T, N, (3, 84, 84), [15], Discrete(8), 512)
        for k in ppd.RolloutStorage._FIELDS:
            getattr(st, k).copy_(roll[k])
        st.to(DEV)
        st.compute_returns(roll["next_value"].to(DEV), True, cfg.gamma, cfg.gae_lambda, False)
        agent = ppd.algo.PPO(pol, cfg.clip_param, cfg.ppo_epoch, cfg.num_mini_batch, cfg.value_loss_coef, cfg.entropy_coef,
                             lr=cfg.lr, eps=cfg.eps, max_grad_norm=cfg.max_grad_norm)
        agent.prefetch_gather = prefetch
        torch.manual_seed(11)
        out = [agent.update(st) for _ in range(2)]          # the second update re-uses the buffers of the first
        results.append((out, {k: v.clone() for k, v in pol.state_dict().items()}))
        assert getattr(st, "_gather_bufs", None) is None      # handed back after every update
    (o1, p1), (o0, p0), (o2, p2) = results
    assert o1 == o0 == o2
    for k in p0:
        assert torch.equal(p1[k], p0[k]) and torch.equal(p2[k], p0[k]), k


@pytest.mark.parametrize("recurrent,prefetch", [(True, True), (True, False), (False, False)])
def test_update_through_cuda_graphs_is_bit_identical(recurrent, prefetch):
    """PPO.update with every minibatch replayed from captured CUDA graphs (minibatch_graph.py: graph A up to the GRU recurrence, the
    next minibatch's gather queued behind it, graph B for the rest; one graph without a recurrence) launches the kernels of the eager
    path with the same arguments: losses and parameters must not differ by a bit.  The third update changes clip_param (a by-value
    kernel argument baked into the graphs): the minibatch is captured again, not replayed with the old value."""
    T, N = 32, 8
    C, V = (3, 15) if recurrent else (1, 0)
    cfg = synthetic.RolloutConfig("t", T, N, C, V, 8, recurrent, 2, 4, 1e-4, 0.001)
    roll = synthetic.make_rollout(cfg, seed=5, reset_prob=0.05)
    results = []
    # 2 = one graph per minibatch, the gather's trigger as an external event-record node (MinibatchGraphs.single)
    for graph in ((False, True, 2) if recurrent else (False, True)):
        torch.manual_seed(0)
        pol = ppd.Policy((C, 84, 84), Discrete(8), base_kwargs={"recurrent": recurrent}, vector_obs_len=V).to(DEV)
        pol.engine("tf32x3")
        st = ppd.RolloutStorage(T, N, (C, 84, 84), [V], Discrete(8), 512)
        for k in ppd.RolloutStorage._FIELDS:
            getattr(st, k).copy_(roll[k])
        st.to(DEV)
        st.compute_returns(roll["next_value"].to(DEV), True, cfg.gamma, cfg.gae_lambda, False)
        agent = ppd.algo.PPO(pol, cfg.clip_param, cfg.ppo_epoch, cfg.num_mini_batch, cfg.value_loss_coef, cfg.entropy_coef,
                             lr=cfg.lr, eps=cfg.eps, max_grad_norm=cfg.max_grad_norm)
        agent.prefetch_gather = prefetch
        agent.use_cuda_graph = graph
        torch.manual_seed(11)
        out = [agent.update(st) for _ in range(2)]
        agent.clip_param = 0.05
        out.append(agent.update(st))
        if graph:
            g = agent._graphs
            slots = 2 if (recurrent and prefetch) else 1
            assert g.disabled is None, g.disabled
            assert g.captures == 2 * slots, g.captures              # each slot once per clip_param value
            # 8 minibatches per update; the first two of a new minibatch shape run eagerly, all others are replays
            assert g.replays == 3 * 8 - 2, g.replays
            assert all(len(e.graphs) == (2 if recurrent and graph is True else 1) for e in g.entries.values())
            assert all((e.ext is not None) == (graph == 2) for e in g.entries.values())
        else:
            assert agent._graphs is None
        # graphs off and on again: the minibatch buffers are rebuilt, the old graphs dropped and new ones captured
        agent.use_cuda_graph, agent.static_minibatch = False, False
        out.append(agent.update(st))
        agent.use_cuda_graph, agent.static_minibatch = graph, True
        out.append(agent.update(st))
        if graph:
            assert g.disabled is None, g.disabled
            # (only the prefetching path keeps observation-only buffers while graphs are off; the others allocate per minibatch
            # then and find their buffers -- and graphs -- unchanged afterwards)
            assert (g.captures, len(g.entries)) == ((3 * slots, slots) if prefetch else (2 * slots, 2 * slots))
        results.append((out, {k: v.clone() for k, v in pol.state_dict().items()}))
    o0, p0 = results[0]
    for o1, p1 in results[1:]:
        assert o0 == o1
        for k in p0:
            assert torch.equal(p1[k], p0[k]), k

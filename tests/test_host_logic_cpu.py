"""Host-side logic that needs no GPU: the minibatch buffers algo.PPO owns (the tensors the generators gather into and captured CUDA
graphs point at) have exactly the shapes / dtypes the reference's generators yield (PKG/storage.py:159-160,222-223, restated in
oracle/minibatch.py), and the launch accounting that keeps `gpu_launches` honest when kernels are replayed from CUDA graphs."""
import torch

import ppodash_b200 as ppd
from oracle import minibatch as o_mb
from ppodash_b200 import _lib, synthetic

SLOT_ORDER = ("obs", "vector_obs", "hxs", "actions", "value_preds", "returns", "masks", "logp", "adv")      # the 9-tuple's order


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


def _agent_and_storage(cfg):
    pol = ppd.Policy((cfg.channels, 84, 84), Discrete(cfg.num_actions), base_kwargs={"recurrent": cfg.recurrent},
                     vector_obs_len=cfg.vector_obs_len)
    st = ppd.RolloutStorage(cfg.num_steps, cfg.num_envs, (cfg.channels, 84, 84), [cfg.vector_obs_len], Discrete(cfg.num_actions),
                            512 if cfg.recurrent else 1)
    agent = ppd.algo.PPO(pol, cfg.clip_param, cfg.ppo_epoch, cfg.num_mini_batch, cfg.value_loss_coef, cfg.entropy_coef,
                         lr=cfg.lr, eps=cfg.eps, max_grad_norm=cfg.max_grad_norm)
    return agent, st


def test_minibatch_buffers_have_the_generators_shapes():
    for recurrent, C, V in ((True, 3, 15), (False, 1, 0)):
        cfg = synthetic.RolloutConfig("t", 6, 4, C, V, 8, recurrent, 1, 2, 1e-4, 0.001)
        roll = synthetic.make_rollout(cfg, seed=3, reset_prob=0.1)
        agent, st = _agent_and_storage(cfg)
        T, N, nmb = cfg.num_steps, cfg.num_envs, cfg.num_mini_batch
        adv = torch.zeros(T, N, 1)
        if recurrent:
            want = next(o_mb.recurrent_minibatches(roll, adv, nmb))
            rows, hrows = T * (N // nmb), N // nmb
        else:
            want = next(o_mb.feed_forward_minibatches(roll, adv, nmb))
            rows = hrows = T * N // nmb
        agent._make_slots(st, rows, hrows, 2, True, "cpu")
        assert len(agent._gbufs) == 2
        for b in agent._gbufs:
            assert tuple(b) == SLOT_ORDER
            for name, w in zip(SLOT_ORDER, want):
                assert tuple(b[name].shape) == tuple(w.shape) and b[name].dtype == w.dtype, (recurrent, name)
        # unchanged request: the same buffers (captured graphs point at them); changed request: new ones
        first = [b["obs"].data_ptr() for b in agent._gbufs]
        agent._make_slots(st, rows, hrows, 2, True, "cpu")
        assert [b["obs"].data_ptr() for b in agent._gbufs] == first
        agent._make_slots(st, rows, hrows, 1, False, "cpu")
        assert len(agent._gbufs) == 1 and tuple(agent._gbufs[0]) == ("obs",)
        # a sample lives in a slot only when all nine tensors do (what MinibatchGraphs may bake in)
        assert not agent._in_slot((agent._gbufs[0]["obs"],))
        agent._make_slots(st, rows, hrows, 1, True, "cpu")
        assert agent._in_slot((agent._gbufs[0]["obs"],)) and not agent._in_slot((torch.zeros(3),))


def test_graph_switches_and_defaults(monkeypatch):
    cfg = synthetic.RolloutConfig("t", 4, 2, 3, 15, 8, True, 1, 2, 1e-4, 0.001)
    for env, want in ((None, False), ("0", False), ("1", True), ("2", 2)):
        if env is None:
            monkeypatch.delenv("PPD_GRAPH", raising=False)
        else:
            monkeypatch.setenv("PPD_GRAPH", env)
        agent, _ = _agent_and_storage(cfg)
        assert agent.use_cuda_graph == want and type(agent.use_cuda_graph) is type(want)
        assert agent.static_minibatch is True and agent.prefetch_gather is True


def test_launch_accounting_counts_replayed_kernels():
    _lib.reset_launch_count()
    assert _lib.launch_count() == 0
    _lib.note_replayed_launches(36)             # one replayed minibatch
    _lib.note_replayed_launches(36)
    assert _lib.launch_count() == 72 and int(_lib.lib().ppd_launch_count()) == 0
    _lib.note_replayed_launches(-36)            # a capture: the library counted 36 launches although nothing ran
    assert _lib.launch_count() == 36
    _lib.reset_launch_count()
    assert _lib.launch_count() == 0
    assert not _lib.profiling()
    with _lib.profiled():
        assert _lib.profiling()                 # algo.PPO launches from Python while bench.py attributes time per C-ABI call
    assert not _lib.profiling()

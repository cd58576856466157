"""PPO: drop-in for PKG/algo/ppo.py (same constructor, ``update(rollouts)`` returning
``(value_loss, action_loss, dist_entropy)`` as Python floats, ``.optimizer.param_groups[i]['lr']``
writable for the linear LR schedule, PKG/utils.py:46-50).

What changes underneath (all on the current CUDA stream, no host sync until the final read):
  * advantage mean / unbiased std over all T*N in one reduction kernel; the normalised advantage is
    recomputed on the fly inside the minibatch gather (ppo.py:35-37);
  * per minibatch: gather kernel -> network forward -> fused PPO loss forward+backward -> network
    backward -> (NCCL all-reduce of the flat gradient buffer when data-parallel) -> fused global-norm
    clip + Adam (ppo.py:57-84).  The reference's three ``.item()`` syncs per minibatch
    (ppo.py:86-88) become one 3-float device-to-host read per ``update``.
Data-parallel: every rank holds the envs of its shard in its own RolloutStorage;
advantage moments (3 doubles) are all-reduced once per update and the flat gradient buffer (with
the loss partial sums in its tail) once per minibatch; all ranks then apply the identical step.
"""
import os

import torch

from .. import _lib
from .. import dist as ppd_dist
from .._lib import check, lib
from ..minibatch_graph import MinibatchGraphs
from ..storage import FusedAdvantages

GRAPH_DEFAULT = "0"
GRAPH_MAX_ROWS = 8192


class FusedClipAdam(torch.optim.Optimizer):
    """torch.optim.Adam-compatible facade (param_groups, state_dict) whose step is the fused
    clip + Adam kernel over the policy's flat parameter buffer."""

    def __init__(self, policy, lr=None, eps=None):
        defaults = dict(lr=1e-3 if lr is None else lr, betas=(0.9, 0.999), eps=1e-8 if eps is None else eps)
        super().__init__(list(policy.parameters()), defaults)
        self._policy = policy
        self.max_grad_norm = None

    def zero_grad(self, set_to_none=False):
        eng = self._policy.engine()
        eng.bind()
        eng.flat_grad.zero_()

    @torch.no_grad()
    def step(self, closure=None, loss_acc=None, grad_norm_out=None):
        g = self.param_groups[0]
        self._policy.engine().adam_step(g["lr"], g["betas"], g["eps"], self.max_grad_norm, loss_acc, grad_norm_out)

    def _layout_signature(self):
        eng = self._policy.engine()
        return [(n, sg.off, sg.numel) for n, sg in eng.segs.items()]

    def state_dict(self):
        """{format, layout, step, exp_avg, exp_avg_sq, param_groups}: the moments are flat buffers in the ENGINE's segment layout
        (permuted conv weights, padded w_ih, 256-byte aligned segments), so the layout travels with them and is checked on load."""
        eng = self._policy.engine()
        eng.bind()
        st = eng.adam_state
        return dict(format="ppodash_b200.FusedClipAdam/1", layout=self._layout_signature(), step=st["step"],
                    exp_avg=st["exp_avg"].clone(), exp_avg_sq=st["exp_avg_sq"].clone(),
                    param_groups=[{k: v for k, v in g.items() if k != "params"} for g in self.param_groups])

    def load_state_dict(self, sd):
        if "exp_avg" not in sd or "layout" not in sd:
            raise ValueError("FusedClipAdam.load_state_dict expects the dict FusedClipAdam.state_dict() writes "
                             "(flat moments + layout signature); torch.optim.Adam's per-parameter format is not accepted")
        eng = self._policy.engine()
        eng.bind()
        if [tuple(x) for x in sd["layout"]] != [tuple(x) for x in self._layout_signature()]:
            raise ValueError("FusedClipAdam.load_state_dict: the saved moments were laid out for another network shape")
        eng.adam_state["step"] = int(sd["step"])
        eng.adam_state["exp_avg"].copy_(sd["exp_avg"])
        eng.adam_state["exp_avg_sq"].copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)


class PPO():
    def __init__(self,
                 actor_critic,
                 clip_param,
                 ppo_epoch,
                 num_mini_batch,
                 value_loss_coef,
                 entropy_coef,
                 lr=None,
                 eps=None,
                 max_grad_norm=None,
                 use_clipped_value_loss=True,
                 process_group=None):
        self.actor_critic = actor_critic
        self.clip_param = clip_param
        self.ppo_epoch = ppo_epoch
        self.num_mini_batch = num_mini_batch
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.optimizer = FusedClipAdam(actor_critic, lr=lr, eps=eps)
        # data parallel: None -> use the default process group if one is initialised with >1 ranks
        self.process_group = process_group
        self.last_grad_norm = None
        # gather minibatch i+1 on a side stream while minibatch i trains (update()); PPD_PREFETCH_GATHER=0 turns it off (A/B timing)
        self.prefetch_gather = os.environ.get("PPD_PREFETCH_GATHER", "1") != "0"
        # one minibatch = replayed CUDA graphs instead of 36 launches from Python (minibatch_graph.py); single process only; opt-in
        # (PPD_GRAPH=1): measured 0.2-0.7 ms per update slower on the device than the Python-issued launches
        # (True / 1: two graphs, split where the recurrence starts; 2: one graph with an external event there, see MinibatchGraphs.single)
        self.use_cuda_graph = {"0": False, "1": True}.get(os.environ.get("PPD_GRAPH", GRAPH_DEFAULT), 2)
        # all nine tensors of a minibatch live in buffers this object owns (graphs need that; without graphs it still saves eight
        # allocations per minibatch on the gather stream and the allocator's cross-stream bookkeeping for them: -0.5 ms per update);
        # PPD_STATIC_MINIBATCH=0: only the observations do, the small fields are allocated per minibatch (A/B timing)
        self.static_minibatch = os.environ.get("PPD_STATIC_MINIBATCH", "1") != "0"
        self._side = None
        self._gbufs = None
        self._graphs = None

    def _gather_stream(self, dev):
        if self._side is None or self._side.device != torch.device(dev):
            self._side = torch.cuda.Stream(device=dev)
        return self._side

    def _world(self):
        return ppd_dist.world(self.process_group)[0]

    def advantage_stats(self, rollouts):
        """Device tensor [mean, std + 1e-5] of returns[:-1] - value_preds[:-1] over all ranks."""
        L = lib()
        T, N = rollouts.rewards.size(0), rollouts.rewards.size(1)
        dev = rollouts.returns.device
        n = T * N
        ws = _lib.workspace(L.ppd_advantage_moments_workspace(n), dev, "advmom")
        mom = torch.empty(3, dtype=torch.float64, device=dev)
        stats = torch.empty(2, dtype=torch.float32, device=dev)
        st = _lib.stream_ptr(dev)
        check(L.ppd_advantage_moments(_lib.ptr(rollouts.returns, torch.float32), _lib.ptr(rollouts.value_preds, torch.float32),
                                      n, mom.data_ptr(), ws.data_ptr(), ws.numel(), st), "advantage_moments")
        ppd_dist.all_reduce_sum(mom, self.process_group)        # {sum, sum of squares, count} over all ranks
        check(L.ppd_advantage_finalize(mom.data_ptr(), stats.data_ptr(), st), "advantage_finalize")
        return stats

    def update(self, rollouts):
        pol = self.actor_critic
        eng = pol.engine()
        eng.bind()
        dev = eng.device
        world = self._world()
        stats = self.advantage_stats(rollouts)
        advantages = FusedAdvantages(stats)
        loss_acc = torch.zeros(3, dtype=torch.float32, device=dev)
        gnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.optimizer.max_grad_norm = self.max_grad_norm

        def all_samples():
            for e in range(self.ppo_epoch):
                if pol.is_recurrent:
                    data_generator = rollouts.recurrent_generator(advantages, self.num_mini_batch)
                else:
                    data_generator = rollouts.feed_forward_generator(advantages, self.num_mini_batch)
                for sample in data_generator:
                    yield sample

        # The minibatch gathers depend only on the rollout and the advantage statistics, not on the parameters: minibatch i+1 is
        # gathered on a side stream while minibatch i trains -- queued behind the point where its GRU recurrence starts, which
        # keeps only 16 SMs per env busy (64 of 148), so the HBM-bound gather runs in the idle ones.  The permutations are drawn
        # in the same order as before (one randperm per epoch, PKG/storage.py:138,169).
        main = torch.cuda.current_stream(dev)
        # ... which only pays while the recurrence leaves a good part of the GPU idle (one 16-SM cluster per env of the minibatch):
        # with many envs per minibatch (or no recurrence) there is nothing to fill, and a second in-flight copy of a multi-GB
        # minibatch only costs allocator traffic (measured at 128 envs x 512 steps per minibatch: 2.9 -> 4.0 s per update).
        envs_per_mb = rollouts.rewards.size(1) // self.num_mini_batch
        use_side = self.prefetch_gather and pol.is_recurrent and 16 * envs_per_mb <= 148 - 32
        side = self._gather_stream(dev) if use_side else None
        # CUDA graphs: single process, default precision mode, and not while bench.py attributes time to C-ABI calls
        # (and for minibatches small enough that launch overhead matters: at 65 536 rows the kernels run for milliseconds)
        graphs = None
        T, N = rollouts.rewards.size(0), rollouts.rewards.size(1)
        if pol.is_recurrent:
            rows, hrows = T * envs_per_mb, envs_per_mb            # rows of a minibatch, rows of its hidden-state tensor
        else:
            rows = hrows = (T * N) // self.num_mini_batch
        if self.use_cuda_graph and world == 1 and eng.precision == "tf32x3" and rows <= GRAPH_MAX_ROWS and not _lib.profiling():
            if self._graphs is None or self._graphs.eng is not eng:
                self._graphs = MinibatchGraphs(eng)
            graphs = self._graphs
            graphs.single = self.use_cuda_graph == 2
        own_slots = side is not None or graphs is not None
        if own_slots:
            # caller-owned minibatch buffers ("slots"), used in turn.  Two when the gathers are prefetched: the gather of minibatch
            # i+2 is queued behind an event of minibatch i+1's forward pass, i.e. behind every reader of minibatch i.  All nine
            # tensors live in them: a captured graph bakes their addresses in, and no block crosses streams through the allocator.
            self._make_slots(rollouts, rows, hrows, 2 if side is not None else 1, graphs is not None or self.static_minibatch, dev)
            rollouts.set_gather_buffers(self._gbufs)
        samples = all_samples()

        def fetch(after=None, first=False):
            if side is None:
                s = next(samples, None)
                return None if s is None else (s, None)
            with torch.cuda.stream(side):
                if first:
                    side.wait_stream(main)            # returns, advantage statistics, last update's writes
                if after is not None:
                    for e in (after if isinstance(after, tuple) else (after,)):
                        side.wait_event(e)
                s = next(samples, None)
                if s is None:
                    return None
                ev = torch.cuda.Event()
                ev.record(side)
            return s, ev

        try:
            eng._fast_bind = True                 # the module tree was checked above; it does not change while update() runs
            self._run_minibatches(fetch, main, side, eng, world, loss_acc, gnorm, graphs)
        finally:
            eng._fast_bind = False
            if own_slots:
                rollouts.set_gather_buffers(None)
        num_updates = self.ppo_epoch * self.num_mini_batch
        vals = (loss_acc / num_updates).tolist()          # the only device->host sync of update()
        self.last_grad_norm = gnorm
        return vals[0], vals[1], vals[2]

    def _make_slots(self, rollouts, rows, hrows, n, full, dev):
        dev = torch.device(dev)
        shape = (rows,) + tuple(getattr(rollouts, "policy_obs_shape", None) or rollouts.obs.shape[2:])
        cur = self._gbufs
        if cur is not None and len(cur) == n and cur[0]["obs"].shape == shape and cur[0]["obs"].device == dev \
                and ("actions" in cur[0]) == full and (not full or cur[0]["hxs"].shape[0] == hrows):
            return
        if self._graphs is not None:
            self._graphs.clear()                  # their graphs point into the buffers about to be dropped
        self._gbufs = []
        for _ in range(n):
            b = {"obs": torch.empty(shape, dtype=torch.float32, device=dev)}
            if full:
                like = lambda t, r=rows: torch.empty((r,) + tuple(t.shape[2:]), dtype=t.dtype, device=dev)
                b.update(vector_obs=like(rollouts.vector_obs), hxs=like(rollouts.recurrent_hidden_states, hrows),
                         actions=like(rollouts.actions), value_preds=like(rollouts.value_preds), returns=like(rollouts.returns),
                         masks=like(rollouts.masks), logp=like(rollouts.action_log_probs),
                         adv=torch.empty(rows, 1, dtype=torch.float32, device=dev))
            self._gbufs.append(b)

    def _in_slot(self, sample):
        p = sample[0].data_ptr()
        return any(b["obs"].data_ptr() == p and "actions" in b for b in (self._gbufs or []))

    def _run_minibatches(self, fetch, main, side, eng, world, loss_acc, gnorm, graphs=None):
        cur = fetch(first=True)
        while cur is not None:
            sample, ev = cur
            if ev is not None:
                main.wait_event(ev)
                for t in sample:
                    if t is not None:
                        t.record_stream(main)         # allocated on the side stream, consumed on this one
            nxt = []
            at_gru = None
            if side is not None:
                def at_gru(events=None, _n=nxt):
                    # events: what the gather has to wait for, when the caller has them already (MinibatchGraphs, one-graph mode)
                    if events is None:
                        events = torch.cuda.Event()
                        events.record(main)
                    _n.append(fetch(after=events))
            rows = sample[0].shape[0]
            if graphs is not None and self._in_slot(sample):
                graphs.run(sample, (float(self.clip_param), float(self.value_loss_coef), float(self.entropy_coef),
                                    bool(self.use_clipped_value_loss), rows), at_gru)
                self._after_minibatch(nxt, fetch, main, side, loss_acc, gnorm)
                cur = nxt[0]
                continue
            eng.on_gru_forward = at_gru
            try:
                if world > 1:
                    # two buckets, all-reduced asynchronously from the stream on which each becomes final: [fc.w, end) (FC, GRU,
                    # heads, loss partials: 97 % of the bytes) overlaps the convolution backward, [0, fc.w) follows at the end
                    pending = []

                    def grad_ready(lo, hi, _p=pending):
                        _p.append(ppd_dist.all_reduce_sum_async(eng.flat_grad[lo:hi], self.process_group))
                    eng.train_minibatch(sample, self.clip_param, self.value_loss_coef, self.entropy_coef,
                                        self.use_clipped_value_loss, global_rows=rows * world, grad_ready=grad_ready)
                    for w in pending:
                        w.wait()              # the current stream waits for the collective (no host sync)
                else:
                    eng.train_minibatch(sample, self.clip_param, self.value_loss_coef, self.entropy_coef,
                                        self.use_clipped_value_loss, global_rows=rows)
            finally:
                eng.on_gru_forward = None
            self._after_minibatch(nxt, fetch, main, side, loss_acc, gnorm)
            cur = nxt[0]

    def _after_minibatch(self, nxt, fetch, main, side, loss_acc, gnorm):
        if not nxt:
            # no recurrence in this network, or a chunked minibatch: gather behind the whole minibatch (the event keeps the side
            # stream behind the readers of the buffer it is about to overwrite)
            e1 = None
            if side is not None:
                e1 = torch.cuda.Event()
                e1.record(main)
            nxt.append(fetch(after=e1))
        self.optimizer.step(loss_acc=loss_acc, grad_norm_out=gnorm)

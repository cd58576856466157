"""MinibatchGraphs: one PPO minibatch (PKG/algo/ppo.py:57-81 -- forward, fused loss forward+backward, backward; ~190 kernel
launches and as many ctypes calls on three streams) as TWO replayed CUDA graphs.

Why two: algo.PPO gathers minibatch i+1 on a side stream behind the point where minibatch i's GRU recurrence starts (the recurrence
keeps 16 SMs per env busy and leaves the rest to the HBM-bound gather).  Graph A ends at that point, the caller queues the gather
behind it (`between()`), graph B starts with the recurrence.  A network without a recurrence is one graph.  `single = True`
(PPD_GRAPH=2) keeps the whole minibatch in one graph and marks that point with an external event-record node instead.

What a graph bakes in, and who keeps it valid:
  * the 9 tensors of the minibatch -- algo.PPO owns two sets of them ("slots") and the generators gather into them in turn
    (RolloutStorage.set_gather_buffers); a sample that does not live in a slot runs eagerly;
  * the flat parameter / gradient buffers and the precision mode -- part of the key (PolicyEngine.signature()); bind() after
    .to() / load_state_dict() on a new module gives a new key and the minibatch is captured again;
  * the engine's activation scratch -- one dict per minibatch shape, owned HERE and routed in with PolicyEngine.scratch() for the
    eager warm-up minibatches and the capture alike, so no later call at another batch size can free or move it;
  * the kernel workspaces of _lib.workspace (grow-only, replaced when a larger request comes): every buffer alive at capture time is
    referenced from the entry, so a replaced one is not freed under a graph that still points at it;
  * clip_param, the loss coefficients and global_rows (kernel arguments by value) -- part of the key.
The clip + Adam step stays outside (its step count and learning rate are by-value arguments that change every minibatch); so do the
gathers (their permutation offsets change) and, in data-parallel runs, everything: with a process group the minibatch runs eagerly.

Replays launch exactly the kernels the eager path launches, with the same arguments: results are bit-identical
(tests/test_gpu_policy_ppo.py::test_update_through_cuda_graphs_is_bit_identical).

Measured on B200 at the PPO-Dash shape (tools/graph_ab.py, profiles/r2_minibatch_graphs.md): 196 instead of 2 500 launches issued
from Python per update, and 0.2-0.7 ms per update MORE device time than the Python-issued launches (which keep ahead of the device
anyway), with two graphs per minibatch or one.  Hence opt-in: PPD_GRAPH=1 (two graphs) / 2 (one), or agent.use_cuda_graph.
"""
import warnings

import torch

from . import _lib

WARM_MINIBATCHES = 2      # eager minibatches of a new shape before the first capture (module loading, lazy attribute calls, workspaces)


class _Entry:
    __slots__ = ("graphs", "held", "launches", "out", "ext")


class MinibatchGraphs:
    def __init__(self, engine):
        self.eng = engine
        self.entries = {}        # key (shape key, by-value arguments, the slot's tensor addresses) -> _Entry
        self.scratch = {}        # shape key -> dict for PolicyEngine.scratch()
        self.warm = {}           # shape key -> eager minibatches run so far
        self.cap = None          # the stream every capture runs on (per-stream kernel workspaces are keyed by it: one set for all graphs)
        self.replays = 0
        self.captures = 0
        # single=True: ONE graph per minibatch; the point where the recurrence starts is an external event-record node
        # (cudaEventRecordExternal) that the gather stream waits on from outside the graph.  Halves the graph launches.
        self.single = False
        self.disabled = None     # reason, once a capture has failed: everything runs eagerly from then on

    def clear(self):
        self.entries.clear()
        self.scratch.clear()
        self.warm.clear()

    def _keys(self, sample, hyper):
        eng = self.eng
        shape_key = (eng.signature(), eng.overlap_wgrad, bool(self.single),
                     tuple(None if t is None else (tuple(t.shape), t.dtype) for t in sample))
        return shape_key, (shape_key, hyper, tuple(None if t is None else t.data_ptr() for t in sample))

    def run(self, sample, hyper, between=None):
        """Train on `sample` (forward + loss + backward; gradients left in the flat gradient buffer).  hyper = (clip_param,
        value_coef, entropy_coef, use_clipped_value_loss, global_rows).  between(): called once, with the current stream at the
        point where the GRU recurrence is about to start (right away for a network without one)."""
        eng = self.eng
        shape_key, key = self._keys(sample, hyper)
        bufs = self.scratch.setdefault(shape_key, {})
        ent = self.entries.get(key)
        if ent is None:
            if self.disabled is not None or self.warm.get(shape_key, 0) < WARM_MINIBATCHES:
                self.warm[shape_key] = self.warm.get(shape_key, 0) + 1
                self._eager(sample, hyper, between, bufs)
                return
            if len(self.entries) >= 8:          # shapes / buffers keep changing: do not pile up graphs
                self.entries.clear()
            try:
                ent = self._capture(sample, hyper, bufs)
            except Exception as e:              # nothing ran yet: say so once and train this (and every later) minibatch eagerly
                self.disabled = f"{type(e).__name__}: {e}"
                warnings.warn("ppodash_b200: CUDA-graph capture of the PPO minibatch failed, continuing with eager launches: "
                              + self.disabled)
                self._eager(sample, hyper, between, bufs)
                return
            self.entries[key] = ent
        self.replays += 1
        if ent.ext is not None and between is not None:
            # One graph.  Correctness of what `between` queues must not rest on how a wait issued after the launch sees the record
            # node inside it, so it also gets an ordinary event recorded BEFORE the launch (= behind the previous minibatch, the last
            # reader of the buffer the next gather overwrites); the external event only holds the gather back to the recurrence.
            before = torch.cuda.Event()
            before.record(torch.cuda.current_stream(eng.device))
            ent.graphs[0].replay()
            between((before, ent.ext))
        else:
            ent.graphs[0].replay()
            if between is not None:
                between()
            for g in ent.graphs[1:]:
                g.replay()
        _lib.note_replayed_launches(ent.launches)

    def _eager(self, sample, hyper, between, bufs):
        eng = self.eng
        fired = []

        def at_gru():
            fired.append(1)
            if between is not None:
                between()
        eng.on_gru_forward = at_gru
        try:
            with eng.scratch(bufs):
                eng.train_minibatch(sample, hyper[0], hyper[1], hyper[2], hyper[3], global_rows=hyper[4])
        finally:
            eng.on_gru_forward = None
        if not fired and between is not None:
            between()

    def _capture(self, sample, hyper, bufs):
        eng = self.eng
        dev = eng.device
        main = torch.cuda.current_stream(dev)
        # one memory pool per entry (its graphs share it; they replay in capture order): a pool dies with the last graph that uses
        # it, so a handle kept across clear() would name a pool the allocator has already dropped
        pool = torch.cuda.graph_pool_handle()
        if self.cap is None or self.cap.device != torch.device(dev):
            self.cap = torch.cuda.Stream(device=dev)
        cap = self.cap
        cap.wait_stream(main)
        graphs = [torch.cuda.CUDAGraph()]
        n0 = _lib.launch_count()
        open_graph = [None]

        def begin(g):
            g.capture_begin(pool=pool, capture_error_mode="thread_local")
            open_graph[0] = g

        def end():
            g, open_graph[0] = open_graph[0], None
            g.capture_end()

        ext = [None]

        def split():
            if self.single:
                if ext[0] is None:
                    ext[0] = torch.cuda.Event(external=True)
                    ext[0].record(cap)                      # an event-record node of the graph
                return
            # graph A ends where the recurrence starts (no side-stream work is open here: the forward pass runs on one stream)
            end()
            graphs.append(torch.cuda.CUDAGraph())
            begin(graphs[-1])
        with torch.cuda.stream(cap):
            eng.on_gru_forward = split
            try:
                begin(graphs[0])
                with eng.scratch(bufs):
                    out = eng.train_minibatch(sample, hyper[0], hyper[1], hyper[2], hyper[3], global_rows=hyper[4])
                end()
            except BaseException:
                if open_graph[0] is not None:
                    try:
                        end()
                    except Exception:
                        pass
                raise
            finally:
                eng.on_gru_forward = None
        main.wait_stream(cap)
        ent = _Entry()
        ent.graphs = graphs
        ent.ext = ext[0]
        ent.out = out                                   # tensors allocated inside the capture (graph pool) stay referenced
        ent.held = _lib.live_workspaces()
        ent.launches = _lib.launch_count() - n0
        _lib.note_replayed_launches(-ent.launches)      # the capture counted them once although nothing ran
        self.captures += 1
        return ent

"""RolloutLoop: the rollout-side inference loop of the reference's training script on the device
(ppo-dash-training/pytorch-a2c-ppo-acktr-gail/run.py:168-216).

The reference does, per environment step,

    value, action, logp, h = actor_critic.act(rollouts.obs[step], rollouts.vector_obs[step],
                                              rollouts.recurrent_hidden_states[step], rollouts.masks[step])
    obs, reward, done, infos = envs.step(action)
    masks     = torch.FloatTensor([[0.0] if done_ else [1.0] for done_ in done])            # run.py:205-207
    bad_masks = torch.FloatTensor([[0.0] if 'bad_transition' in info.keys() else [1.0] ...  # run.py:208-210
    rollouts.insert(obs, vector_obs, h, action, logp, value, reward, masks, bad_masks)

i.e. ~20 kernel launches driven from Python for `act`, Python-list tensor builds for the masks and
pageable host->device copies.  Here:

  * `act()` replays ONE CUDA graph (captured on first use) that runs the whole policy forward
    (implicit-GEMM convolutions, FC, GRU step, heads, categorical sample / log-prob) on static
    input buffers, then copies the actions to a pinned host buffer;
  * `observe()` stages observations / rewards / done flags in pinned host memory, uploads them
    with asynchronous copies and builds masks = 1 - done and bad_masks = 1 - bad on the device;
  * both write straight into the RolloutStorage slots `insert` would fill (same bookkeeping,
    `rollouts.step` advances the same way), so `compute_returns` / `PPO.update` follow unchanged.

There is no CPU fallback; `use_cuda_graph=False` runs the same kernels eagerly (used by the tests
to check that the graph replays what the eager path computes).
"""
import numpy as np
import torch

from . import _lib


class RolloutLoop:
    def __init__(self, actor_critic, rollouts, deterministic=False, use_cuda_graph=True):
        self.policy = actor_critic
        self.rollouts = rollouts
        self.deterministic = bool(deterministic)
        self.use_cuda_graph = bool(use_cuda_graph)
        dev = rollouts.obs.device
        if dev.type != "cuda":
            raise _lib.PpdError("RolloutLoop needs the rollout storage on a CUDA device (rollouts.to(device))")
        self.device = dev
        N = rollouts.obs.shape[1]
        self.N = N
        # static inputs / outputs of the captured graph
        u8 = getattr(rollouts, "obs_u8", False)          # uint8 frame storage: the policy input is expanded from it on the device
        obs_in = torch.zeros((N,) + rollouts.policy_obs_shape, device=dev) if u8 else torch.zeros_like(rollouts.obs[0])
        self._in = dict(obs=obs_in, vobs=torch.zeros_like(rollouts.vector_obs[0]),
                        h=torch.zeros_like(rollouts.recurrent_hidden_states[0]), m=torch.ones_like(rollouts.masks[0]))
        self._out = None
        self._graph = None
        self._sig = None
        self._bufs = {}               # scratch of the captured forward pass: owned here, never shared with eager / training calls
        # pinned staging
        pin = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype).pin_memory()
        self._h_obs = pin(*rollouts.obs.shape[1:], dtype=rollouts.obs.dtype)
        self._h_vobs = pin(*rollouts.vector_obs.shape[1:])
        self._h_rew = pin(N, 1)
        self._h_flags = pin(2, N, 1)                          # done, bad_transition as 0/1 floats
        self._h_act = pin(N, rollouts.actions.shape[2], dtype=rollouts.actions.dtype)
        self._d_flags = torch.zeros(2, N, 1, device=dev)
        self._last = None

    # ------------------------------------------------------------------ act
    def _forward(self):
        i = self._in
        with torch.no_grad(), self.policy.engine().scratch(self._bufs):
            return self.policy.act(i["obs"], i["vobs"], i["h"], i["m"], deterministic=self.deterministic)

    def _capture(self):
        # The graph bakes in device addresses: its scratch buffers live in self._bufs (PolicyEngine.scratch), so PPO.update or an
        # eager act() at another batch size cannot free or reuse them, and `_sig` records what else it depends on (flat parameter
        # buffers, precision); act() re-captures when that changes (engine.bind() after .to() / load_state_dict on a new module).
        self._graph, self._out = None, None
        self._bufs.clear()
        self._sig = self.policy.engine().signature()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                          # warm-up: workspaces, lazy binds, tensor maps
            for _ in range(2):
                self._forward()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self._forward()
        self._graph, self._out = g, out

    def act(self):
        """`actor_critic.act` on the storage's current step; returns the actions as a pinned CPU tensor [N, A]
        (valid after this call returns: it waits for the copy)."""
        r, s = self.rollouts, self.rollouts.step
        i = self._in
        r.obs_at(s, out=i["obs"]) if getattr(r, "obs_u8", False) else i["obs"].copy_(r.obs[s])
        i["vobs"].copy_(r.vector_obs[s])
        i["h"].copy_(r.recurrent_hidden_states[s]); i["m"].copy_(r.masks[s])
        if self.use_cuda_graph:
            if self._graph is None or self._sig != self.policy.engine().signature():
                self._capture()
            self._graph.replay()
            out = self._out
        else:
            out = self._forward()
        self._last = out
        self._h_act.copy_(out[1], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._h_act

    # ------------------------------------------------------------------ observe
    @staticmethod
    def _stage(dst, src):
        if isinstance(src, torch.Tensor):
            dst.copy_(src.reshape(dst.shape))
        else:
            np_dtype = np.uint8 if dst.dtype == torch.uint8 else np.float32
            dst.copy_(torch.from_numpy(np.ascontiguousarray(src, dtype=np_dtype)).reshape(dst.shape))

    def observe(self, obs, vector_obs, reward, done, bad_transition=None):
        """The environment's answer to the last `act()`: new observations, rewards, `done` flags and (optionally) the
        'bad_transition' flags of run.py:208-210.  Host arrays / CPU tensors; device tensors are used as they are."""
        if self._last is None:
            raise RuntimeError("observe() follows act()")
        value, action, logp, h = self._last
        dev = self.device

        def up(pinned, x):
            if isinstance(x, torch.Tensor) and x.is_cuda:
                return x.to(pinned.dtype).reshape(pinned.shape)
            self._stage(pinned, x)
            return pinned.to(dev, non_blocking=True)

        self._h_flags[0].copy_(torch.as_tensor(np.asarray(done, dtype=np.float32)).reshape(self.N, 1))
        if bad_transition is None:
            self._h_flags[1].zero_()
        else:
            self._h_flags[1].copy_(torch.as_tensor(np.asarray(bad_transition, dtype=np.float32)).reshape(self.N, 1))
        self._d_flags.copy_(self._h_flags, non_blocking=True)
        masks = 1.0 - self._d_flags[0]
        bad_masks = 1.0 - self._d_flags[1]
        self.rollouts.insert(up(self._h_obs, obs), up(self._h_vobs, vector_obs), h, action, logp, value,
                             up(self._h_rew, reward), masks, bad_masks)
        self._last = None

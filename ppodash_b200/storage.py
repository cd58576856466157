"""RolloutStorage: drop-in for the reference class (PKG/storage.py:9-223).

Same constructor, public tensor attributes, ``insert`` / ``after_update`` /
``compute_returns`` / ``feed_forward_generator`` / ``recurrent_generator``.
The buffers are ordinary time-major torch tensors ([T(+1), N, ...], fp32, int64
actions) so callers can keep indexing and mutating them (run.py:140-141,172-175,237);
the three heavy methods enqueue sm_100a kernels from libppodash_b200.so on the
current CUDA stream and never synchronise:

  compute_returns          -> ppd_compute_returns           (storage.py:82-121)
  feed_forward_generator   -> ppd_gather_feed_forward       (storage.py:123-160)
  recurrent_generator      -> ppd_gather_recurrent          (storage.py:162-223)

The minibatch permutations are drawn on the host from torch's global CPU generator with the
same call the reference makes (one ``torch.randperm`` per epoch), so with equal seeds the index
sets are bit-identical to the reference's.  There is no CPU fallback for the kernels.
"""
import ctypes

import torch

from . import _lib
from ._lib import GatherDesc, ObsU8Desc, check, lib, ptr, stream_ptr


class FusedAdvantages:
    """Handle that lets the generators compute the normalised advantage
    (returns - value_preds - mean) / (std + 1e-5) on the fly (ppo.py:35-37) instead of
    gathering it from a materialised [T, N, 1] tensor.  ``stats`` is a 2-float device tensor."""

    def __init__(self, stats):
        self.stats = stats


def _flatten_helper(T, N, _tensor):
    return _tensor.view(T * N, *_tensor.size()[2:])


class RolloutStorage(object):
    """Reference constructor (PKG/storage.py:10-11) plus one opt-in extension:

    ``obs_dtype=torch.uint8`` keeps the observations as the uint8 frames the environment emits -- each frame ONCE, CHW -- and
    moves the env-side float pipeline of the reference onto the device: ``(frame - obs_mean) / obs_std`` (or ``frame / 255``) in
    float64 rounded once to float32 (NormalizeWrapper + VecPyTorch's ``.float()``), and the ``frame_stack``-deep stack of
    VecPyTorchFrameStack with its zeroing at episode starts.  The float32 values the policy sees are bit-identical to the
    reference's; they are produced when a frame is read (``obs_at(step)`` for act / get_value, the generators for training).
    ``obs_shape`` stays the shape the POLICY sees, ``(frame_stack * C, H, W)``; ``rollouts.obs[t]`` is the newest uint8 frame
    ``[N, C, H, W]`` of slot t and ``insert`` takes the new frame.  ``obs_mean``: float64 ``[C, H, W]`` (the reference's (H, W, C)
    mean file transposed) or None; ``obs_std``: the divisor (None: 255 without a mean -- the reference's ``obs / 255`` -- else 1).
    4x (no stack) to 16x (4-stack) fewer observation bytes in HBM and over PCIe."""

    def __init__(self, num_steps, num_processes, obs_shape, vector_obs_shape, action_space,
                 recurrent_hidden_state_size, obs_dtype=None, frame_stack=1, obs_mean=None, obs_std=None):
        self.obs_u8 = obs_dtype == torch.uint8
        if obs_dtype not in (None, torch.float32, torch.uint8):
            raise ValueError("obs_dtype must be None / torch.float32 (reference layout) or torch.uint8")
        if self.obs_u8:
            if len(obs_shape) != 3 or obs_shape[0] % int(frame_stack) != 0 or frame_stack < 1:
                raise ValueError("uint8 storage needs obs_shape = (frame_stack * C, H, W)")
            self.frame_stack = int(frame_stack)
            c = obs_shape[0] // self.frame_stack
            # frames [T + nstack, N, C, H, W]; slot t's newest frame is frame t + nstack - 1 = self.obs[t]
            self._frames = torch.zeros(num_steps + self.frame_stack, num_processes, c, *obs_shape[1:], dtype=torch.uint8)
            self.obs = self._frames[self.frame_stack - 1:]
            self.obs_age = torch.zeros(num_steps + 1, num_processes, dtype=torch.uint8)     # earlier frames of the same episode
            self.obs_mean = None
            if obs_mean is not None:
                self.obs_mean = torch.as_tensor(obs_mean, dtype=torch.float64).reshape(c, *obs_shape[1:]).contiguous().clone()
            self.obs_div = float(obs_std) if obs_std is not None else (255.0 if obs_mean is None else 1.0)
            self.policy_obs_shape = tuple(obs_shape)
            self._mul_exact = None
        else:
            if frame_stack != 1 or obs_mean is not None or obs_std is not None:
                raise ValueError("frame_stack / obs_mean / obs_std need obs_dtype=torch.uint8")
            self.obs = torch.zeros(num_steps + 1, num_processes, *obs_shape)
        self.vector_obs = torch.zeros(num_steps + 1, num_processes, *vector_obs_shape)
        self.recurrent_hidden_states = torch.zeros(num_steps + 1, num_processes, recurrent_hidden_state_size)
        self.rewards = torch.zeros(num_steps, num_processes, 1)
        self.value_preds = torch.zeros(num_steps + 1, num_processes, 1)
        self.returns = torch.zeros(num_steps + 1, num_processes, 1)
        self.action_log_probs = torch.zeros(num_steps, num_processes, 1)
        if action_space.__class__.__name__ == 'Discrete':
            action_shape = 1
        else:
            action_shape = action_space.shape[0]
        self.actions = torch.zeros(num_steps, num_processes, action_shape)
        if action_space.__class__.__name__ == 'Discrete':
            self.actions = self.actions.long()
        self.masks = torch.ones(num_steps + 1, num_processes, 1)
        # 0 where the episode ended because of a time limit rather than a true terminal state
        self.bad_masks = torch.ones(num_steps + 1, num_processes, 1)
        self.num_steps = num_steps
        self.step = 0

    _FIELDS = ("obs", "vector_obs", "recurrent_hidden_states", "rewards", "value_preds", "returns",
               "action_log_probs", "actions", "masks", "bad_masks")

    def to(self, device):
        self.finish_upload()
        for name in self._FIELDS:
            if name == "obs" and self.obs_u8:
                continue
            setattr(self, name, getattr(self, name).to(device, non_blocking=True))
        if self.obs_u8:
            self._frames = self._frames.to(device, non_blocking=True)
            self.obs = self._frames[self.frame_stack - 1:]             # keep `obs` a view of the frame ring
            self.obs_age = self.obs_age.to(device, non_blocking=True)
            if self.obs_mean is not None:
                self.obs_mean = self.obs_mean.to(device, non_blocking=True)

    # ------------------------------------------------------------------ uint8 frames
    def _u8_desc(self):
        d = ObsU8Desc()
        d.frames, d.age = ptr(self._frames, torch.uint8), ptr(self.obs_age, torch.uint8)
        d.mean = ptr(self.obs_mean, torch.float64) if self.obs_mean is not None else None
        d.divisor = self.obs_div
        d.C, d.HW, d.nstack = self._frames.size(2), self._frames[0, 0, 0].numel(), self.frame_stack
        if self._mul_exact is None:
            # once per (mean, divisor): may the kernels multiply by 1/divisor instead of dividing?  (all 256 x C*H*W cases checked)
            bad = torch.ones(1, dtype=torch.int32, device=self._frames.device)
            check(lib().ppd_obs_u8_certify(d.mean, d.C * d.HW, d.divisor, ptr(bad, torch.int32), stream_ptr(self._frames.device)),
                  "obs_u8_certify")
            self._mul_exact = int(bad.item()) == 0
        d.multiply_exact = 1 if self._mul_exact else 0
        return d

    def obs_at(self, step, out=None):
        """float32 ``[N, frame_stack * C, H, W]`` observation of storage slot `step` -- what the reference keeps in
        ``rollouts.obs[step]`` and feeds to ``actor_critic.act`` (run.py:172-175).  For the reference layout this is ``obs[step]``."""
        if not self.obs_u8:
            return self.obs[step]
        self.finish_upload()
        N = self.rewards.size(1)
        step = int(step) % (self.num_steps + 1)
        if out is None:
            out = torch.empty((N,) + self.policy_obs_shape, dtype=torch.float32, device=self._frames.device)
        d = self._u8_desc()
        check(lib().ppd_obs_u8_expand(ctypes.byref(d), step, N, ptr(out, torch.float32), stream_ptr(self._frames.device)), "obs_u8_expand")
        return out

    # ------------------------------------------------------------------ staged upload of a host rollout
    def upload_from(self, host, staged=True):
        """Fill this (device) storage from `host`: another RolloutStorage or a dict of CPU tensors with the same fields
        (the job of the reference's `rollouts.to(device)`, storage.py:34-46, for a rollout collected on the host).

        Every field but `obs` is a few MB and is copied at once on the current stream.  The observations -- 99 % of the
        bytes -- are uploaded PER ENV on a copy stream when `host["obs"]` is pinned and `staged` is true: the copies are
        queued when the first minibatch generator draws its env permutation, in that order, and a minibatch waits only for
        the envs it gathers, so the upload overlaps compute_returns and the first epoch of `PPO.update` instead of
        preceding them.  `finish_upload()` (called by everything else that touches `obs`) waits for all of it."""
        self.finish_upload()
        get = (lambda k: getattr(host, k)) if not isinstance(host, dict) else (lambda k: host[k])
        dev = self.obs.device
        if dev.type != "cuda":
            raise _lib.PpdError("upload_from needs the storage on a CUDA device (rollouts.to(device))")
        for k in self._FIELDS:
            src = get(k)
            if tuple(src.shape) != tuple(getattr(self, k).shape):
                raise ValueError("upload_from: field {} has shape {}, expected {}".format(k, tuple(src.shape), tuple(getattr(self, k).shape)))
        for k in self._FIELDS:
            if k != "obs":
                getattr(self, k).copy_(get(k), non_blocking=True)
        obs = get("obs")
        if self.obs_u8:
            if obs.dtype != torch.uint8:
                raise TypeError("upload_from: uint8 storage takes uint8 frames")
            # the stack depth of every slot (and, for frame_stack > 1, the nstack - 1 frames before slot 0)
            has = (lambda k: k in host) if isinstance(host, dict) else (lambda k: hasattr(host, k))
            if has("obs_age"):
                self.obs_age.copy_(get("obs_age"), non_blocking=True)
            if self.frame_stack > 1 and has("_frames"):
                self._frames[:self.frame_stack - 1].copy_(get("_frames")[:self.frame_stack - 1], non_blocking=True)
        if staged and obs.device.type == "cpu" and obs.is_pinned() and obs.is_contiguous() and obs.dtype == self.obs.dtype \
                and self.obs.is_contiguous():
            self._pending = {"host": obs, "events": None}
        else:
            self.obs.copy_(obs, non_blocking=True)

    def _issue_upload(self, order=None):
        """Queue the per-env observation copies (env order `order`, default natural) on the copy stream; one event per env."""
        p = getattr(self, "_pending", None)
        if p is None or p["events"] is not None:
            return
        dev = self.obs.device
        N = self.obs.size(1)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs = self._copy_stream
        cs.wait_stream(torch.cuda.current_stream(dev))          # earlier readers / writers of obs on the main stream
        row = self.obs[0, 0].numel() * self.obs.element_size()
        rows = self.obs.size(0)
        order = list(range(N)) if order is None else [int(e) for e in order]
        seen = set(order)
        order += [e for e in range(N) if e not in seen]
        events = [None] * N
        for e in order:
            check(lib().ppd_upload_rows(self.obs.data_ptr() + e * row, N * row, p["host"].data_ptr() + e * row, N * row,
                                        row, rows, cs.cuda_stream), "upload_rows")
            ev = torch.cuda.Event()
            ev.record(cs)
            events[e] = ev
        p["events"] = events

    def _wait_envs(self, envs=None):
        """Make the current stream wait for the uploads of `envs` (all when None)."""
        p = getattr(self, "_pending", None)
        if p is None:
            return
        if p["events"] is None:
            self._issue_upload()
        cur = torch.cuda.current_stream(self.obs.device)
        left = False
        for e, ev in enumerate(p["events"]):
            if ev is None:
                continue
            if envs is None or e in envs:
                cur.wait_event(ev)
                p["events"][e] = None
            else:
                left = True
        if not left:
            self._pending = None

    def finish_upload(self):
        self._wait_envs(None)

    def half(self):
        # The reference's experimental --half_precision path (storage.py:48-58) also casts the int64
        # actions to fp16, which breaks Categorical.log_prob; no published run used it.
        raise NotImplementedError("half-precision rollout storage is out of scope (SURVEY.md section 5)")

    def insert(self, obs, vector_obs, recurrent_hidden_states, actions, action_log_probs,
               value_preds, rewards, masks, bad_masks):
        s = self.step
        self.finish_upload()
        if self.obs_u8:
            if obs.dtype != torch.uint8:
                raise TypeError("uint8 storage: insert() takes the new uint8 frame [N, C, H, W] (normalisation and stacking happen on the device)")
            if tuple(obs.shape[1:]) != tuple(self.obs.shape[2:]):
                raise ValueError("uint8 storage: insert() takes ONE new frame per env, shape {}".format(tuple(self.obs.shape[2:])))
        self.obs[s + 1].copy_(obs, non_blocking=True)
        self.vector_obs[s + 1].copy_(vector_obs, non_blocking=True)
        self.recurrent_hidden_states[s + 1].copy_(recurrent_hidden_states, non_blocking=True)
        self.actions[s].copy_(actions, non_blocking=True)
        self.action_log_probs[s].copy_(action_log_probs, non_blocking=True)
        self.value_preds[s].copy_(value_preds, non_blocking=True)
        self.rewards[s].copy_(rewards, non_blocking=True)
        self.masks[s + 1].copy_(masks, non_blocking=True)
        self.bad_masks[s + 1].copy_(bad_masks, non_blocking=True)
        if self.obs_u8:
            # VecPyTorchFrameStack.step_wait (make_env.py:39-46): an env whose episode just ended starts a fresh stack
            grown = torch.clamp(self.obs_age[s].to(torch.int16) + 1, max=self.frame_stack - 1)
            self.obs_age[s + 1] = torch.where(self.masks[s + 1].reshape(-1) == 0, torch.zeros_like(grown), grown).to(torch.uint8)
        self.step = (self.step + 1) % self.num_steps

    def after_update(self):
        self.finish_upload()
        if self.obs_u8:
            ns = self.frame_stack
            self._frames[:ns].copy_(self._frames[-ns:].clone() if ns > 1 else self._frames[-ns:])    # slot T's whole stack -> slot 0
            self.obs_age[0].copy_(self.obs_age[-1])
        else:
            self.obs[0].copy_(self.obs[-1])
        self.vector_obs[0].copy_(self.vector_obs[-1])
        self.recurrent_hidden_states[0].copy_(self.recurrent_hidden_states[-1])
        self.masks[0].copy_(self.masks[-1])
        self.bad_masks[0].copy_(self.bad_masks[-1])

    # ------------------------------------------------------------------ returns / GAE
    def compute_returns(self, next_value, use_gae, gamma, gae_lambda, use_proper_time_limits=True):
        T, N = self.rewards.size(0), self.rewards.size(1)
        dev = self.rewards.device
        nv = next_value.detach().to(device=dev, dtype=torch.float32).reshape(N).contiguous()
        if dev.type != "cuda":
            raise _lib.PpdError("compute_returns needs the storage on a CUDA device (no CPU fallback)")
        # the look-back workspace re-arms itself between calls (epoch counters), which only works for calls that are ordered: one
        # workspace per stream, so that returns computed concurrently on two streams of a device never share it
        ws = _lib.workspace(lib().ppd_compute_returns_workspace(T, N), dev, "returns@%x" % stream_ptr(dev), zero=True)
        check(lib().ppd_compute_returns(
            ptr(self.rewards, torch.float32), ptr(self.value_preds, torch.float32),
            ptr(self.masks, torch.float32), ptr(self.bad_masks, torch.float32),
            ptr(self.returns, torch.float32), ptr(nv), T, N, float(gamma), float(gae_lambda),
            int(bool(use_gae)), int(bool(use_proper_time_limits)), ws.data_ptr(), ws.numel(),
            stream_ptr(dev)), "compute_returns")

    # ------------------------------------------------------------------ generators
    def _out(self, rows, like, dtype=None):
        return torch.empty((rows,) + tuple(like.shape[2:]), dtype=dtype or like.dtype, device=like.device)

    def _gather(self, mode, perm_dev, start, rows, E, advantages, out=None):
        """Fill one minibatch; returns the 9-tuple in the reference's order (storage.py:159-160)."""
        T, N = self.rewards.size(0), self.rewards.size(1)
        dev = self.obs.device
        o = out or self._next_gather_buffers(rows) or {}
        if self.obs_u8:
            obs_b = o.get("obs") if "obs" in o else torch.empty((rows,) + self.policy_obs_shape, dtype=torch.float32, device=dev)
        else:
            obs_b = o.get("obs") if "obs" in o else self._out(rows, self.obs)
        vobs_b = o.get("vector_obs") if "vector_obs" in o else self._out(rows, self.vector_obs)
        hrows = rows if mode == "ff" else E
        hxs_b = o.get("hxs") if "hxs" in o else self._out(hrows, self.recurrent_hidden_states)
        act_b = o.get("actions") if "actions" in o else self._out(rows, self.actions)
        val_b = o.get("value_preds") if "value_preds" in o else self._out(rows, self.value_preds)
        ret_b = o.get("returns") if "returns" in o else self._out(rows, self.returns)
        msk_b = o.get("masks") if "masks" in o else self._out(rows, self.masks)
        lp_b = o.get("logp") if "logp" in o else self._out(rows, self.action_log_probs)
        d = GatherDesc()
        if self.obs_u8:
            # observations: uint8 frames -> normalised, stacked float32 rows (ppd_gather_obs_u8_*); the small fields follow below
            u = self._u8_desc()
            if mode == "ff":
                check(lib().ppd_gather_obs_u8_feed_forward(ctypes.byref(u), ptr(perm_dev, torch.int64), start, rows, T, N,
                                                           ptr(obs_b, torch.float32), stream_ptr(dev)), "gather_obs_u8")
            else:
                check(lib().ppd_gather_obs_u8_recurrent(ctypes.byref(u), ptr(perm_dev, torch.int64), start, E, T, N,
                                                        ptr(obs_b, torch.float32), stream_ptr(dev)), "gather_obs_u8")
        else:
            obs_row = self.obs[0, 0].numel()
            d.obs, d.obs_out, d.obs_row = ptr(self.obs, torch.float32), ptr(obs_b), obs_row
        vrow = self.vector_obs[0, 0].numel()
        if vrow > 0:
            d.vobs, d.vobs_out, d.vobs_row = ptr(self.vector_obs, torch.float32), vobs_b.data_ptr(), vrow
            d.vobs_out_ld = int(o.get("vector_obs_ld", 0))
        d.hxs, d.hxs_out = ptr(self.recurrent_hidden_states, torch.float32), ptr(hxs_b)
        d.hxs_row = self.recurrent_hidden_states.size(-1)
        d.actions, d.actions_out, d.actions_row = ptr(self.actions, torch.int64), ptr(act_b), self.actions.size(-1)
        d.value_preds, d.value_preds_out = ptr(self.value_preds, torch.float32), ptr(val_b)
        d.returns, d.returns_out = ptr(self.returns, torch.float32), ptr(ret_b)
        d.masks, d.masks_out = ptr(self.masks, torch.float32), ptr(msk_b)
        d.logp, d.logp_out = ptr(self.action_log_probs, torch.float32), ptr(lp_b)
        adv_b = None
        keep = None
        if isinstance(advantages, FusedAdvantages):
            adv_b = o.get("adv") if "adv" in o else torch.empty(rows, 1, dtype=torch.float32, device=dev)
            d.adv_stats, d.adv_out = ptr(advantages.stats, torch.float32), ptr(adv_b)
        elif advantages is not None:
            keep = advantages.detach()
            if keep.dtype != torch.float32 or not keep.is_contiguous():
                keep = keep.float().contiguous()
            adv_b = o.get("adv") if "adv" in o else torch.empty(rows, 1, dtype=torch.float32, device=dev)
            d.adv, d.adv_out = ptr(keep), ptr(adv_b)
        if mode == "ff":
            rc = lib().ppd_gather_feed_forward(ctypes.byref(d), ptr(perm_dev, torch.int64), start, rows, T, N,
                                               stream_ptr(dev))
        else:
            rc = lib().ppd_gather_recurrent(ctypes.byref(d), ptr(perm_dev, torch.int64), start, E, T, N,
                                            stream_ptr(dev))
        check(rc, "minibatch gather")
        return obs_b, vobs_b, hxs_b, act_b, val_b, ret_b, msk_b, lp_b, adv_b

    def set_gather_buffers(self, buffers):
        """Caller-owned output buffers for the generators, used in rotation (None: allocate per minibatch, the default).  `buffers`:
        list of dicts {"obs": float32 [rows, *policy obs shape]} and, optionally, any of "vector_obs", "hxs", "actions", "value_preds",
        "returns", "masks", "logp", "adv" (shapes as the generators yield them; fields left out are allocated per minibatch).  algo.PPO gathers minibatch i+1 on a side stream while minibatch i
        trains and hands two sets over, so that no 100-MB block crosses streams through the caching allocator."""
        self._gather_bufs = list(buffers) if buffers else None
        self._gather_buf_i = 0

    def _next_gather_buffers(self, rows):
        bufs = getattr(self, "_gather_bufs", None)
        if not bufs:
            return None
        b = bufs[self._gather_buf_i % len(bufs)]
        if b["obs"].shape[0] != rows:
            return None
        self._gather_buf_i += 1
        return b

    @staticmethod
    def _perm_to_device(perm, device):
        if not torch.device(device).type == "cuda":
            raise _lib.PpdError("RolloutStorage generators need the storage on a CUDA device "
                                "(call rollouts.to(device) first; there is no CPU fallback)")
        return perm.pin_memory().to(device, non_blocking=True)

    def feed_forward_generator(self, advantages, num_mini_batch=None, mini_batch_size=None):
        num_steps, num_processes = self.rewards.size()[0:2]
        batch_size = num_processes * num_steps
        if mini_batch_size is None:
            assert batch_size >= num_mini_batch, (
                "PPO requires the number of processes ({}) "
                "* number of steps ({}) = {} "
                "to be greater than or equal to the number of PPO mini batches ({})."
                "".format(num_processes, num_steps, num_processes * num_steps, num_mini_batch))
            mini_batch_size = batch_size // num_mini_batch
        # SubsetRandomSampler(range(batch_size)) draws exactly this (CPU global generator); BatchSampler
        # with drop_last=True then cuts it into consecutive blocks (storage.py:138-142).
        perm = torch.randperm(batch_size)
        perm_dev = self._perm_to_device(perm, self.obs.device)
        self.finish_upload()                                       # a feed-forward minibatch samples every env
        for k in range(batch_size // mini_batch_size):
            yield self._gather("ff", perm_dev, k * mini_batch_size, mini_batch_size, 0, advantages)

    def recurrent_generator(self, advantages, num_mini_batch):
        num_processes = self.rewards.size(1)
        assert num_processes >= num_mini_batch, (
            "PPO requires the number of processes ({}) "
            "to be greater than or equal to the number of "
            "PPO mini batches ({}).".format(num_processes, num_mini_batch))
        num_envs_per_batch = num_processes // num_mini_batch
        perm = torch.randperm(num_processes)                       # storage.py:169
        perm_dev = self._perm_to_device(perm, self.obs.device)
        T = self.num_steps
        pending = getattr(self, "_pending", None) is not None
        if pending:
            self._issue_upload(perm.tolist())                      # a staged upload follows this epoch's env order
        for start_ind in range(0, num_processes, num_envs_per_batch):
            if start_ind + num_envs_per_batch > num_processes:
                # the reference runs off the end of `perm` here (storage.py:182)
                raise IndexError("index {} is out of bounds for dimension 0 with size {}".format(
                    num_processes, num_processes))
            if pending:
                self._wait_envs(set(perm[start_ind:start_ind + num_envs_per_batch].tolist()))
            yield self._gather("rec", perm_dev, start_ind, T * num_envs_per_batch, num_envs_per_batch, advantages)

"""Oracle (test infrastructure): minibatch index generation and gathers.

Restates PKG/storage.py:123-160 (feed_forward_generator) and
PKG/storage.py:162-223 (recurrent_generator).  The permutations come from the
CPU global torch RNG exactly as in the reference (one ``torch.randperm`` per
epoch: ``SubsetRandomSampler.__iter__`` for the feed-forward path,
storage.py:169 for the recurrent path), so seeding torch identically yields
bit-identical index sets.

A rollout is passed as a dict of arrays/tensors with the reference attribute
names (obs, vector_obs, recurrent_hidden_states, actions, value_preds,
returns, masks, action_log_probs); everything is handled as torch CPU tensors.
"""
import torch

FIELDS = ("obs", "vector_obs", "recurrent_hidden_states", "actions", "value_preds",
          "returns", "masks", "action_log_probs")


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x)


def feed_forward_indices(T, N, num_mini_batch=None, mini_batch_size=None):
    """List of int64 index tensors, one per minibatch (storage.py:127-142)."""
    batch = T * N
    if mini_batch_size is None:
        assert batch >= num_mini_batch
        mini_batch_size = batch // num_mini_batch
    perm = torch.randperm(batch)                      # SubsetRandomSampler(range(batch))
    nfull = batch // mini_batch_size                  # BatchSampler(drop_last=True)
    return [perm[k * mini_batch_size:(k + 1) * mini_batch_size] for k in range(nfull)]


def feed_forward_minibatches(roll, advantages, num_mini_batch=None, mini_batch_size=None):
    """Yield the 9-tuple of storage.py:159-160 (variant A: every field gathered)."""
    T, N = _t(roll["rewards"]).shape[:2]
    flat = {}
    for k in FIELDS:
        x = _t(roll[k])
        x = x[:T]                                      # [:-1] for the T+1 fields, all of the T fields
        flat[k] = x.reshape(T * N, *x.shape[2:])
    adv = None if advantages is None else _t(advantages).reshape(T * N, 1)
    for idx in feed_forward_indices(T, N, num_mini_batch, mini_batch_size):
        yield (flat["obs"][idx], flat["vector_obs"][idx], flat["recurrent_hidden_states"][idx],
               flat["actions"][idx], flat["value_preds"][idx], flat["returns"][idx],
               flat["masks"][idx], flat["action_log_probs"][idx],
               None if adv is None else adv[idx])


def recurrent_env_blocks(N, num_mini_batch):
    """List of int64 env-index tensors, one per minibatch (storage.py:163-170,182)."""
    assert N >= num_mini_batch
    E = N // num_mini_batch
    perm = torch.randperm(N)
    blocks = []
    for s in range(0, N, E):
        if s + E > N:
            # the reference indexes perm[s + offset] past its end here (storage.py:182)
            raise IndexError("num_processes must be divisible by num_mini_batch")
        blocks.append(perm[s:s + E])
    return blocks


def recurrent_minibatches(roll, advantages, num_mini_batch):
    """Yield the 9-tuple of storage.py:222-223: [T*E,...] time-major rows, h0 [E,H]."""
    T, N = _t(roll["rewards"]).shape[:2]
    adv = _t(advantages)
    for envs in recurrent_env_blocks(N, num_mini_batch):
        E = envs.numel()

        def cols(name):
            x = _t(roll[name])[:T]                     # [T, N, ...]
            y = x[:, envs]                             # [T, E, ...]  == stack of per-env columns, dim 1
            return y.reshape(T * E, *y.shape[2:])

        h0 = _t(roll["recurrent_hidden_states"])[0, envs].reshape(E, -1)
        a = adv[:, envs].reshape(T * E, *adv.shape[2:])
        yield (cols("obs"), cols("vector_obs"), h0, cols("actions"), cols("value_preds"),
               cols("returns"), cols("masks"), cols("action_log_probs"), a)

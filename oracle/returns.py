"""Oracle (test infrastructure): returns / GAE recurrence and advantage normalisation.

Restates PKG/storage.py:82-121 (the four ``compute_returns`` branches) and
PKG/algo/ppo.py:35-37 (advantage normalisation) with numpy float32, keeping
the reference's operation order so results are bit-comparable with torch-CPU.
"""
import numpy as np

F32 = np.float32


def returns_recurrence(rewards, value_preds, masks, bad_masks, next_value,
                       use_gae, gamma, gae_lambda, use_proper_time_limits=True,
                       returns_in=None):
    """Backward recurrence over time for one rollout.

    rewards [T,N,1]; value_preds, masks, bad_masks [T+1,N,1]; next_value [N,1].
    Returns (returns [T+1,N,1], value_preds_out [T+1,N,1]).  Side effects of the
    reference are reproduced: the GAE branches overwrite value_preds[T]
    (storage.py:90,108) and leave returns[T] untouched; the discounted
    branches write returns[T] (storage.py:101,118).
    """
    r = np.asarray(rewards, dtype=F32)
    v = np.array(value_preds, dtype=F32, copy=True)
    m = np.asarray(masks, dtype=F32)
    b = np.asarray(bad_masks, dtype=F32)
    T = r.shape[0]
    ret = np.zeros_like(v) if returns_in is None else np.array(returns_in, dtype=F32, copy=True)
    g = F32(gamma)
    gl = F32(gamma * gae_lambda)           # python-double product, then one rounding (torch scalar semantics)
    nv = np.asarray(next_value, dtype=F32).reshape(v[-1].shape)
    if use_gae:
        v[T] = nv
        acc = np.zeros_like(v[0])
        for t in range(T - 1, -1, -1):
            delta = (r[t] + (g * v[t + 1]) * m[t + 1]) - v[t]          # storage.py:93-95 / 111-113
            acc = delta + (gl * m[t + 1]) * acc                        # storage.py:96-97 / 114-115
            if use_proper_time_limits:
                acc = acc * b[t + 1]                                   # storage.py:98
            ret[t] = acc + v[t]                                        # storage.py:99 / 116
    else:
        ret[T] = nv
        for t in range(T - 1, -1, -1):
            base = (ret[t + 1] * g) * m[t + 1] + r[t]                  # storage.py:103-104 / 120-121
            if use_proper_time_limits:
                base = base * b[t + 1] + (F32(1) - b[t + 1]) * v[t]    # storage.py:104-105
            ret[t] = base
    return ret, v


def normalized_advantages(returns, value_preds):
    """(adv - mean) / (std + 1e-5), std unbiased, over all T*N (ppo.py:35-37).

    Computed through torch-CPU so the mean/std reductions are the reference's.
    """
    import torch
    ret = torch.as_tensor(np.asarray(returns, dtype=F32))
    v = torch.as_tensor(np.asarray(value_preds, dtype=F32))
    adv = ret[:-1] - v[:-1]
    adv = (adv - adv.mean()) / (adv.std() + 1e-5)
    return adv.numpy()

"""CPU: the C-ABI library loads and exports every symbol include/ppodash_b200.h declares
(no compute calls -- there is no GPU here), and the product path refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ppodash_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ppd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ppodash_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    # every declared symbol has a ctypes prototype, and vice versa
    assert set(_lib.exported_symbols()) == set(names)
    assert _lib.lib().ppd_abi_version() == 1


def test_no_cpu_fallback():
    from ppodash_b200 import _lib
    with pytest.raises(_lib.PpdError):
        _lib.ptr(torch.zeros(4))
    from ppodash_b200.storage import RolloutStorage

    class Discrete:
        n = 4
    st = RolloutStorage(4, 2, (1, 2, 2), [0], Discrete(), 1)
    with pytest.raises(_lib.PpdError):
        st.compute_returns(torch.zeros(2, 1), True, 0.99, 0.95, False)
    with pytest.raises(_lib.PpdError):
        next(st.feed_forward_generator(None, 2))
    with pytest.raises(_lib.PpdError):                      # a staged upload targets device storage
        st.upload_from({k: getattr(st, k) for k in RolloutStorage._FIELDS})
    rc = _lib.lib().ppd_upload_rows(None, 0, None, 0, 0, 0, None)
    assert rc == -1 and b"upload_rows" in _lib.lib().ppd_last_error()


def test_bad_arguments_return_error_codes():
    from ppodash_b200 import _lib
    L = _lib.lib()
    rc = L.ppd_compute_returns(None, None, None, None, None, None, 4, 4, 0.99, 0.95, 1, 0, None, 0, None)
    assert rc == -1 and b"null" in L.ppd_last_error()
    assert L.ppd_advantage_moments_workspace(1000) >= 16
    assert L.ppd_clip_adam_workspace(1 << 20) > 256


def test_storage_host_semantics():
    """insert / after_update bookkeeping (PKG/storage.py:60-80) on CPU tensors."""
    from ppodash_b200.storage import RolloutStorage

    class Discrete:
        n = 4
    T, N = 3, 2
    st = RolloutStorage(T, N, (1, 2, 2), [3], Discrete(), 5)
    assert st.actions.dtype == torch.int64 and st.masks.eq(1).all() and st.bad_masks.eq(1).all()
    for t in range(T):
        st.insert(torch.full((N, 1, 2, 2), t + 1.0), torch.full((N, 3), t + 1.0), torch.full((N, 5), t + 1.0),
                  torch.full((N, 1), t, dtype=torch.long), torch.full((N, 1), -t - 1.0), torch.full((N, 1), t + 0.5),
                  torch.full((N, 1), 0.25 * t), torch.zeros(N, 1), torch.ones(N, 1))
    assert st.step == 0
    assert st.obs[3].eq(3).all() and st.actions[2].eq(2).all() and st.rewards[1].eq(0.25).all()
    assert st.masks[1:].eq(0).all() and st.masks[0].eq(1).all()
    st.after_update()
    assert st.obs[0].eq(3).all() and st.masks[0].eq(0).all() and st.recurrent_hidden_states[0].eq(3).all()


def test_uint8_storage_bookkeeping_on_cpu():
    """Host logic of the uint8 storage (no kernel): frame ring / `obs` view aliasing, stack-depth (`age`) bookkeeping of insert
    against VecPyTorchFrameStack's zeroing rule (make_env.py:39-46), after_update carry, and the loud failure without a GPU."""
    import numpy as np
    import torch
    from ppodash_b200 import _lib
    from ppodash_b200.storage import RolloutStorage

    class Discrete:
        def __init__(self, n):
            self.n = n
            self.shape = ()
    T, N, ns = 6, 3, 4
    st = RolloutStorage(T, N, (ns * 3, 84, 84), [0], Discrete(4), 1, obs_dtype=torch.uint8, frame_stack=ns)
    assert st.obs.dtype == torch.uint8 and tuple(st.obs.shape) == (T + 1, N, 3, 84, 84) and tuple(st._frames.shape) == (T + ns, N, 3, 84, 84)
    assert st.obs.data_ptr() == st._frames[ns - 1].data_ptr() and st.obs_div == 255.0
    rng = np.random.RandomState(0)
    dones = rng.rand(T, N) < 0.4
    z = lambda *s: torch.zeros(*s)
    for t in range(T):
        frame = torch.full((N, 3, 84, 84), t + 1, dtype=torch.uint8)
        masks = torch.FloatTensor([[0.0] if d else [1.0] for d in dones[t]])
        st.insert(frame, z(N, 0), z(N, 1), torch.zeros(N, 1, dtype=torch.int64), z(N, 1), z(N, 1), z(N, 1), masks, torch.ones(N, 1))
        assert int(st.obs[t + 1].max()) == t + 1
    # age[t, n] = steps since the last episode start, capped at ns - 1
    want = np.zeros((T + 1, N), np.int64)
    for t in range(T):
        want[t + 1] = np.where(dones[t], 0, np.minimum(want[t] + 1, ns - 1))
    assert np.array_equal(st.obs_age.numpy(), want)
    import pytest
    with pytest.raises(TypeError):
        st.insert(torch.zeros(N, 3, 84, 84), z(N, 0), z(N, 1), torch.zeros(N, 1, dtype=torch.int64), z(N, 1), z(N, 1), z(N, 1), torch.ones(N, 1), torch.ones(N, 1))
    last_frames, last_age = st._frames[-ns:].clone(), st.obs_age[-1].clone()
    st.after_update()
    assert torch.equal(st._frames[:ns], last_frames) and torch.equal(st.obs_age[0], last_age)
    with pytest.raises(_lib.PpdError):
        st.obs_at(0)                      # expanding needs the CUDA kernel: no CPU fallback
    with pytest.raises(ValueError):
        RolloutStorage(T, N, (3, 84, 84), [0], Discrete(4), 1, frame_stack=4)

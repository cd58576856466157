#!/usr/bin/env python
"""Generates tests/golden/obs_pipeline.npz from the UNMODIFIED reference wrapper classes (authoring container only: needs
/root/reference).  gym / baselines / cv2 are not installed, so minimal stand-ins for the base classes the wrappers derive from are
registered first; the classes under test -- NormalizeWrapper (013 study), TransposeImage and VecPyTorchFrameStack (training
make_env.py) -- are the reference's own code, run on seeded uint8 frames.

Stored: the mean / std of ObtRetro-v6 (float64, as np.loadtxt reads them), the seeded inputs' seed and shape, sha256 digests of
the reference outputs (float32 bytes) for three pipelines, and the first storage slots of each in full for debugging.
"""
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
S013 = os.path.join(REF, "ppo-dash-study/013_ra+no_stack+lshp+recurrent+vec_obs+norm_obs+rew_hacking")
TRAIN = os.path.join(REF, "ppo-dash-training/pytorch-a2c-ppo-acktr-gail")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "obs_pipeline.npz")


def install_stubs():
    gym = types.ModuleType("gym")

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                self.low, self.high = np.asarray(low), np.asarray(high)
                shape = self.low.shape
            else:
                self.low, self.high = np.full(shape, low), np.full(shape, high)
            self.shape, self.dtype = tuple(shape), dtype

    class Wrapper:
        def __init__(self, env):
            self.env = env
            self.observation_space = env.observation_space
            self.action_space = getattr(env, "action_space", None)

    class ObservationWrapper(Wrapper):
        def observation(self, obs):
            raise NotImplementedError

    class RewardWrapper(Wrapper):
        pass

    class ActionWrapper(Wrapper):
        pass

    class Env:
        pass

    spaces = types.ModuleType("gym.spaces")
    spaces.Box = Box
    spaces.Discrete = type("Discrete", (), {"__init__": lambda self, n: setattr(self, "n", n)})
    spaces.Dict = dict
    spaces.MultiDiscrete = type("MultiDiscrete", (), {})
    box = types.ModuleType("gym.spaces.box")
    box.Box = Box
    gym.Wrapper, gym.ObservationWrapper, gym.RewardWrapper, gym.ActionWrapper, gym.Env = Wrapper, ObservationWrapper, RewardWrapper, ActionWrapper, Env
    gym.spaces = spaces
    gym.error = types.SimpleNamespace(Error=Exception)
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.spaces.box": box})
    cv2 = types.ModuleType("cv2")
    cv2.ocl = types.SimpleNamespace(setUseOpenCL=lambda flag: None)
    sys.modules["cv2"] = cv2
    for name in ("baselines", "baselines.common", "baselines.common.atari_wrappers", "baselines.common.vec_env", "sohojoe_dummy_vec_env",
                 "sohojoe_shmem_vec_env", "a2c_ppo_acktr", "a2c_ppo_acktr.envs", "inverse_rl"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["baselines.common.atari_wrappers"].LazyFrames = object

    class VecEnvWrapper:
        def __init__(self, venv, observation_space=None, action_space=None):
            self.venv = venv
            self.num_envs = venv.num_envs
            self.observation_space = observation_space or venv.observation_space

    sys.modules["baselines.common.vec_env"].VecEnvWrapper = VecEnvWrapper
    sys.modules["sohojoe_dummy_vec_env"].DummyVecEnv = object
    sys.modules["sohojoe_shmem_vec_env"].ShmemVecEnv = object
    sys.modules["a2c_ppo_acktr.envs"].VecNormalize = object
    sys.modules["inverse_rl"].InverseRL = object
    return gym


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def digest(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def main():
    # The reference runs with torch.set_num_threads(1) (run.py:55).  It matters here: VecPyTorchFrameStack shifts its buffer with an
    # OVERLAPPING in-place copy (make_env.py:41-42), which is only a shift when executed front to back; multi-threaded torch splits
    # the copy into chunks and produces mixtures of frames.
    torch.set_num_threads(1)
    gym = install_stubs()
    wr = load("sohojoe_wrappers", os.path.join(S013, "sohojoe_wrappers.py"))          # NormalizeWrapper of the PPO-Dash study
    me = load("make_env_ref", os.path.join(TRAIN, "make_env.py"))                     # TransposeImage, VecPyTorchFrameStack
    Box = sys.modules["gym.spaces"].Box

    T, N, H, W, C, NSTACK, SEED = 5, 3, 84, 84, 3, 4, 20260
    rng = np.random.RandomState(SEED)
    frames = rng.randint(0, 256, size=(T + 1, N, H, W, C), dtype=np.uint8)
    dones = rng.rand(T, N) < 0.3
    dones[1, 0] = True

    class RawEnv:
        observation_space = Box(0, 255, (H, W, C), dtype=np.uint8)
        action_space = None

    def pipeline(normal_filename):
        """Per-env wrappers as make_env.py stacks them (NormalizeWrapper -> TransposeImage), then the vec env's np.stack and
        VecPyTorch's .float(): float32 [T+1, N, C, H, W]."""
        cwd = os.getcwd()
        os.chdir(S013)                                     # the mean / std files are opened relative to the run directory
        try:
            norm = wr.NormalizeWrapper(RawEnv(), normal_filename)
        finally:
            os.chdir(cwd)
        tr = me.TransposeImage(norm, op=[2, 0, 1])
        out = []
        for t in range(T + 1):
            per_env = [tr.observation(norm.observation(frames[t, n])) for n in range(N)]
            out.append(torch.from_numpy(np.stack(per_env)).float())
        return torch.stack(out), norm

    norm_obs, norm = pipeline("ObtRetro-v6")
    div255_obs, _ = pipeline(None)

    class Venv:
        num_envs = N
        observation_space = Box(-10, 10, (C, H, W), dtype=np.float32)

        def __init__(self, seq):
            self.seq, self.t = seq, 0

        def reset(self):
            self.t = 0
            return self.seq[0]

        def step_wait(self):
            self.t += 1
            return self.seq[self.t], np.zeros(N), dones[self.t - 1], [{} for _ in range(N)]

    fs = me.VecPyTorchFrameStack(Venv(norm_obs), NSTACK)
    stacked = [fs.reset().clone()]
    for t in range(T):
        stacked.append(fs.step_wait()[0].clone())
    stacked = torch.stack(stacked)                         # [T+1, N, NSTACK*C, H, W]

    np.savez_compressed(
        OUT, mean=np.asarray(norm.mean, dtype=np.float64), std=np.float64(norm.std), seed=SEED, shape=np.array([T, N, H, W, C, NSTACK]),
        dones=dones, sha_norm=digest(norm_obs), sha_div255=digest(div255_obs), sha_stack=digest(stacked),
        head_norm=norm_obs[:1].numpy(), head_div255=div255_obs[:1].numpy(), head_stack=stacked[2, :1].numpy())
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", "norm", tuple(norm_obs.shape), "stack", tuple(stacked.shape))


if __name__ == "__main__":
    main()

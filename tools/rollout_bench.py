#!/usr/bin/env python
"""Env-steps/s of the rollout-side inference loop (act + insert, synthetic environment answers): reference-style Python
loop over Policy.act vs RolloutLoop eager vs RolloutLoop with the captured CUDA graph."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppodash_b200 as ppd  # noqa: E402


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


def main():
    dev = "cuda:0"
    N, T, C, V, A = 32, 128, 3, 15, 8
    torch.manual_seed(0)
    pol = ppd.Policy((C, 84, 84), Discrete(A), base_kwargs={"recurrent": True}, vector_obs_len=V).to(dev)
    obs = np.random.randn(N, C, 84, 84).astype(np.float32)
    vobs = np.random.rand(N, V).astype(np.float32)
    rew = np.zeros(N, np.float32)
    done = np.zeros(N, bool)

    def fresh():
        st = ppd.RolloutStorage(T, N, (C, 84, 84), [V], Discrete(A), 512)
        st.to(dev)
        return st

    def ref_loop(st):
        for _ in range(T):
            s = st.step
            with torch.no_grad():
                value, action, logp, h = pol.act(st.obs[s], st.vector_obs[s], st.recurrent_hidden_states[s], st.masks[s])
            action.cpu()                                               # envs.step(action) needs it on the host
            masks = torch.FloatTensor([[0.0] if d else [1.0] for d in done])
            bad = torch.FloatTensor([[1.0] for _ in done])
            st.insert(torch.from_numpy(obs), torch.from_numpy(vobs), h, action, logp, value, torch.from_numpy(rew).unsqueeze(1), masks, bad)

    def loop(st, graph):
        lp = ppd.RolloutLoop(pol, st, use_cuda_graph=graph)
        def run():
            for _ in range(T):
                lp.act()
                lp.observe(obs, vobs, rew, done)
        return run

    res = {}
    for name, mk in (("python_loop_policy_act", lambda st: (lambda: ref_loop(st))), ("rollout_loop_eager", lambda st: loop(st, False)),
                     ("rollout_loop_cuda_graph", lambda st: loop(st, True))):
        st = fresh()
        fn = mk(st)
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        res[name] = dict(ms_per_env_step_batch=round(best / T * 1e3, 4), env_steps_per_s=round(N * T / best))
    print(json.dumps(dict(workload=f"{N} envs x {T} steps, 3x84x84 + 15 vector obs, GRU-512, synthetic env answers from host arrays", **res)))


if __name__ == "__main__":
    main()

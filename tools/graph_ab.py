"""A/B of PPO.update with and without CUDA graphs of the minibatch (ppodash_b200/minibatch_graph.py) in ONE process on one GPU:
the bench workload (C2, uint8 storage), the same agent, `agent.use_cuda_graph` / `agent.static_minibatch` switched between
blocks of steps (three modes: eager with the small minibatch fields allocated per minibatch, eager with all nine minibatch tensors in
buffers PPO owns, graphs).
    python tools/graph_ab.py [--workload c2] [--steps 4] [--rounds 3]
Prints one JSON line per block (device time per update from CUDA events, kernel launches per update and how many of them were
issued from Python) and a summary."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--rounds", type=int, default=3)
    a = ap.parse_args()
    import torch
    import ppodash_b200 as ppd
    from ppodash_b200 import _lib, synthetic
    cfg = synthetic.CONFIGS[a.workload]
    args = argparse.Namespace(obs="auto", precision="tf32x3")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    obs_mode = bench.obs_mode_of(args, cfg)
    work = bench.make_workload(cfg, 1234, obs_mode, need_f32=False)
    st, host, nv_host = bench.build_storage(ppd, torch, cfg, work, dev, obs_mode)
    st.upload_from(host)
    st.finish_upload()
    nv = nv_host.to(dev)
    torch.manual_seed(0)
    pol = ppd.Policy((cfg.channels, cfg.obs_hw, cfg.obs_hw), bench.Discrete(cfg.num_actions),
                     base_kwargs={"recurrent": cfg.recurrent, "hidden_size": cfg.hidden_size}, vector_obs_len=cfg.vector_obs_len).to(dev)
    pol.engine("tf32x3")
    agent = ppd.algo.PPO(pol, cfg.clip_param, cfg.ppo_epoch, cfg.num_mini_batch, cfg.value_loss_coef, cfg.entropy_coef,
                         lr=cfg.lr, eps=cfg.eps, max_grad_norm=cfg.max_grad_norm)

    def step():
        st.compute_returns(nv, True, cfg.gamma, cfg.gae_lambda, False)
        torch.manual_seed(99)
        return agent.update(st)

    MODES = {"eager, small fields allocated per minibatch": (False, False), "eager": (False, True), "graphs": (True, True),
             "one graph": (2, True)}

    def block(mode, n):
        graph, agent.static_minibatch = MODES[mode]
        agent.use_cuda_graph = graph
        step()                                          # settle (captures on the first graph block)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        _lib.reset_launch_count()
        ev[0].record()
        t0 = time.perf_counter()
        for i in range(n):
            out = step()
            ev[i + 1].record()
        wall = time.perf_counter() - t0
        torch.cuda.synchronize()
        each = [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(n)]
        g = agent._graphs
        return dict(mode=mode, ms=round(sum(each) / n, 3), each=each, wall_ms=round(1e3 * wall / n, 3), losses=list(out),
                    launches=_lib.launch_count() // n, from_python=int(_lib.lib().ppd_launch_count()) // n,
                    capture_failed=g.disabled if g else None)

    for _ in range(2):
        step()
    res = []
    for r in range(a.rounds):
        for mode in MODES:
            b = block(mode, a.steps)
            res.append(b)
            print(json.dumps(b), flush=True)
    g = agent._graphs
    summary = {m: [b["ms"] for b in res if b["mode"] == m] for m in MODES}
    summary.update(captures=g.captures if g else 0, capture_failed=g.disabled if g else None)
    print(json.dumps(summary), flush=True)


if __name__ == "__main__":
    main()

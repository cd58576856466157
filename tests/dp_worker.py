"""Worker of tests/test_gpu_dp_nccl.py: launched by torch.distributed.run, one rank per GPU (NCCL).

SURVEY.md 8e parity: G ranks, each holding the envs [r*N/G, (r+1)*N/G) of one global rollout in its own RolloutStorage, run
compute_returns + PPO.update on the REAL CUDA path with NCCL gradient all-reduces; rank 0 then runs the same update in a single
process on the global rollout with the equivalent global permutation (ppodash_b200.dist.equivalent_global_*_blocks, fed through a
patched torch.randperm).  Advantage statistics, per-minibatch gradients (after the all-reduce), losses and final parameters must
agree to fp32 tolerance (only the summation order differs).  Exits non-zero on any mismatch; prints one summary line on rank 0.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ppodash_b200 as ppd  # noqa: E402
from ppodash_b200 import dist as ppd_dist  # noqa: E402
from ppodash_b200 import synthetic  # noqa: E402


class Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


def run_update(cfg, roll, dev, group, perm_source, recurrent):
    """One compute_returns + PPO.update; returns (advantage stats, per-minibatch flat gradients, losses, final params, perms drawn)."""
    torch.manual_seed(0)
    V = cfg.vector_obs_len
    pol = ppd.Policy((cfg.channels, 84, 84), Discrete(cfg.num_actions), base_kwargs={"recurrent": recurrent}, vector_obs_len=V).to(dev)
    N = roll["rewards"].shape[1]
    st = ppd.RolloutStorage(cfg.num_steps, N, (cfg.channels, 84, 84), [V], Discrete(cfg.num_actions), cfg.hidden_size if recurrent else 1)
    for k in ppd.RolloutStorage._FIELDS:
        getattr(st, k).copy_(roll[k])
    st.to(dev)
    st.compute_returns(roll["next_value"].to(dev), True, cfg.gamma, cfg.gae_lambda, False)
    agent = ppd.algo.PPO(pol, cfg.clip_param, cfg.ppo_epoch, cfg.num_mini_batch, cfg.value_loss_coef, cfg.entropy_coef,
                         lr=cfg.lr, eps=cfg.eps, max_grad_norm=cfg.max_grad_norm, process_group=group)
    stats = agent.advantage_stats(st).cpu()
    grads = []
    eng = pol.engine()
    step0 = agent.optimizer.step

    def spy(*a, **kw):
        grads.append(eng.flat_grad.clone())
        return step0(*a, **kw)
    agent.optimizer.step = spy
    drawn = []
    real = torch.randperm

    def randperm(n, *a, **kw):
        p = perm_source(n) if callable(perm_source) else real(n, *a, **kw)
        drawn.append(p.clone())
        return p
    torch.randperm = randperm
    try:
        losses = agent.update(st)
    finally:
        torch.randperm = real
    torch.cuda.synchronize()
    return stats, grads, losses, {k: v.detach().cpu().clone() for k, v in pol.state_dict().items()}, drawn


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "recurrent"
    recurrent = mode == "recurrent"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    solo = dist.new_group([0])
    n_local, T, nmb, epochs = 4, 16, 2, 2
    if recurrent:
        cfg = synthetic.RolloutConfig("dp", T, n_local * world, 3, 15, 8, True, epochs, nmb, 1e-4, 0.001)
    else:
        cfg = synthetic.RolloutConfig("dp_ff", T, n_local * world, 4, 0, 54, False, epochs, nmb, 2.5e-4, 0.01)
    roll = synthetic.make_rollout(cfg, seed=77, reset_prob=0.05)
    N = cfg.num_envs
    sl = ppd_dist.shard_envs(N, rank, world)
    shard = {k: (v[:, sl].contiguous() if v.dim() >= 2 and v.shape[0] in (T, T + 1) else v[sl].contiguous()) for k, v in roll.items()}
    torch.manual_seed(100 + rank)                       # every rank draws its OWN permutations
    stats, grads, losses, params, drawn = run_update(cfg, shard, dev, None, None, recurrent)
    all_drawn = [None] * world
    dist.all_gather_object(all_drawn, [p.tolist() for p in drawn])
    ok = True
    if rank == 0:
        # equivalent global permutation of every epoch
        it = iter(range(epochs))
        def source(n):
            e = next(it)
            perms = [torch.tensor(all_drawn[r][e]) for r in range(world)]
            if recurrent:
                blocks = ppd_dist.equivalent_global_env_blocks(perms, n_local, nmb)
            else:
                blocks = ppd_dist.equivalent_global_sample_blocks(perms, T, n_local, nmb)
            return torch.cat(blocks)
        stats1, grads1, losses1, params1, _ = run_update(cfg, roll, dev, solo, source, recurrent)
        rep = {}
        np.testing.assert_allclose(stats.numpy(), stats1.numpy(), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(np.array(losses), np.array(losses1), rtol=1e-5, atol=1e-7)
        assert len(grads) == len(grads1) == epochs * nmb
        worst = 0.0
        for k, (g, g1) in enumerate(zip(grads, grads1)):
            g, g1 = g.cpu().numpy(), g1.cpu().numpy()
            scale = float(np.abs(g1).max())
            err = np.abs(g - g1)
            worst = max(worst, float(err.max()) / scale)
            # stated fp32 tolerance: 1e-4 relative + 1e-5 of the buffer's largest entry (partial sums are added in another order);
            # later minibatches start from parameters that already differ by rounding, hence the small outlier allowance
            okk = err <= 1e-4 * np.abs(g1) + 1e-5 * scale
            assert okk.mean() >= (1.0 if k == 0 else 0.999) and err.max() <= 5e-4 * scale, (k, float(okk.mean()), float(err.max()) / scale)
        for k in params:
            err = (params[k] - params1[k]).abs().numpy()
            assert (err <= 0.05 * cfg.lr).mean() >= 0.98 and err.max() <= cfg.lr, (k, float(err.max()))
        print(f"dp_nccl_parity ok: mode={mode} world={world} minibatches={len(grads)} worst_grad_err/scale={worst:.2e} "
              f"losses_dp={losses} losses_single={losses1}", flush=True)
    # ranks hold bit-identical parameters after the update (identical reduced gradients, identical Adam step; no broadcast)
    flat = torch.cat([params[k].reshape(-1) for k in sorted(params)]).to(dev)
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = torch.tensor([int(torch.equal(flat, ref))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    assert same.item() == 1, "parameters diverged between ranks"
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

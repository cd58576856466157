"""CPU, world_size 2 over gloo: the env-sharded data-parallel scheme of SURVEY.md 8e.
The collectives and index bookkeeping are the ones PPO.update uses on the GPUs (ppodash_b200.dist);
the per-rank compute is done by the oracle so the test needs no GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import minibatch as o_mb
from oracle import policy as o_pol
from oracle import ppo_update as o_upd
from ppodash_b200 import dist as ppd_dist
from ppodash_b200 import synthetic

T, N, NMB, WORLD = 6, 8, 2, 2
CFG = synthetic.RolloutConfig("dp", T, N, 1, 3, 5, True, 1, NMB, 1e-4, 0.001, hidden_size=16)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    roll = synthetic.make_rollout(CFG, seed=5, reset_prob=0.1)
    roll["returns"] = torch.randn(T + 1, N, 1, generator=torch.Generator().manual_seed(1))
    torch.manual_seed(0)
    params = o_pol.init_params(1, 5, 3, True, 16)
    return roll, params


def _loss_grads(params, mb, adv_rows, global_rows):
    """Oracle loss on the rows of `mb`, with every mean taken over `global_rows` rows."""
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    obs, vobs, h0, actions, old_v, ret, masks, old_logp, _ = mb
    v, lp, ent_mean, _ = o_pol.evaluate_actions(p, obs, vobs, h0, masks, actions, True, True)
    rows = obs.shape[0]
    vl, al = o_upd.ppo_losses(v, lp, None, old_v, ret, old_logp, adv_rows, 0.1)
    scale = rows / global_rows
    loss = (vl * 0.5 + al - ent_mean * 0.001) * scale
    loss.backward()
    flat = torch.cat([p[k].grad.reshape(-1) for k in sorted(p)])
    return flat, torch.tensor([vl.item() * scale, al.item() * scale, ent_mean.item() * scale], dtype=torch.float64)


def _worker(rank, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        torch.set_num_threads(1)
        roll, params = _problem()
        sl = ppd_dist.shard_envs(N, rank, WORLD)
        local = {k: (v[:, sl].contiguous() if v.dim() >= 2 and v.shape[1] == N else v) for k, v in roll.items()}
        # ---- advantage moments: 3 float64 all-reduced once per update
        adv = (local["returns"][:-1] - local["value_preds"][:-1]).double()
        mom = torch.tensor([adv.sum(), (adv * adv).sum(), float(adv.numel())], dtype=torch.float64)
        ppd_dist.all_reduce_sum(mom)
        mean = mom[0] / mom[2]
        std = torch.sqrt((mom[1] - mom[0] * mean) / (mom[2] - 1))
        adv_n = ((local["returns"][:-1] - local["value_preds"][:-1]) - mean.float()) / (std.float() + 1e-5)
        # ---- one epoch: every rank permutes its own envs; gradients summed per minibatch
        torch.manual_seed(100 + rank)
        n_local = N // WORLD
        perm = torch.randperm(n_local)
        perms = [torch.empty_like(perm) for _ in range(WORLD)]
        dist.all_gather(perms, perm)
        torch.manual_seed(100 + rank)
        grads, losses = [], []
        for mb in o_mb.recurrent_minibatches(local, adv_n, NMB):
            rows = mb[0].shape[0]
            g, l = _loss_grads(params, mb, mb[8], rows * WORLD)
            ppd_dist.all_reduce_sum(g)
            ppd_dist.all_reduce_sum(l)
            grads.append(g)
            losses.append(l)
        if rank == 0:
            torch.save(dict(mean=mean, std=std, perms=perms, grads=grads, losses=losses), out)
    finally:
        dist.destroy_process_group()


def test_env_sharded_update_equals_single_process(tmp_path):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(_free_port(), out), nprocs=WORLD, join=True)
    got = torch.load(out)
    roll, params = _problem()
    # single process, same minibatch composition through the equivalent global env blocks
    adv = roll["returns"][:-1] - roll["value_preds"][:-1]
    assert abs(got["mean"].item() - adv.double().mean().item()) < 1e-9
    assert abs(got["std"].item() - adv.double().std().item()) < 1e-9
    adv_n = (adv - adv.mean()) / (adv.std() + 1e-5)
    blocks = ppd_dist.equivalent_global_env_blocks(got["perms"], N // WORLD, NMB)
    assert sorted(torch.cat(blocks).tolist()) == list(range(N))
    for k, envs in enumerate(blocks):
        E = envs.numel()
        cols = lambda name: roll[name][:T][:, envs].reshape(T * E, *roll[name].shape[2:])
        mb = (cols("obs"), cols("vector_obs"), roll["recurrent_hidden_states"][0, envs], cols("actions"),
              cols("value_preds"), cols("returns"), cols("masks"), cols("action_log_probs"), None)
        g, l = _loss_grads(params, mb, adv_n[:, envs].reshape(T * E, 1), T * E)
        scale = float(g.abs().max())
        assert float((got["grads"][k] - g).abs().max()) <= 2e-5 * scale          # fp32 summation order only
        np.testing.assert_allclose(got["losses"][k].numpy(), l.numpy(), rtol=1e-5, atol=1e-8)


def test_index_helpers():
    assert ppd_dist.shard_envs(32, 3, 8) == slice(12, 16)
    with pytest.raises(ValueError):
        ppd_dist.shard_envs(10, 0, 4)
    assert ppd_dist.world() == (1, 0)
    t = torch.ones(3)
    assert ppd_dist.all_reduce_sum(t) is t and torch.equal(t, torch.ones(3))     # single rank: no-op
    # feed-forward: local flat index t*n_local+n  ->  global t*N + r*n_local + n
    perms = [torch.tensor([0, 5, 2, 3, 4, 1]), torch.tensor([5, 4, 3, 2, 1, 0])]
    blocks = ppd_dist.equivalent_global_sample_blocks(perms, T=3, n_local=2, num_mini_batch=2)
    assert sorted(torch.cat(blocks).tolist()) == list(range(12))
    assert blocks[0].tolist() == [0, 9, 4, 11, 10, 7]

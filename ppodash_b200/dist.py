"""Data-parallel plumbing for the PPO update: one process per GPU, envs sharded.

The hot path shards over environments: every per-env quantity (GAE recurrence, GRU unroll, the
env columns of a recurrent minibatch) is independent across envs, and the only cross-env couplings
are three means -- the advantage mean/std over all T*N, the minibatch-mean losses and the resulting
gradients.  So the data path needs exactly two collectives (NCCL over NVLink/NVSwitch on the GPU
box; the same code runs over gloo on CPU tensors in the tests):

  * once per update : all-reduce of {sum(adv), sum(adv^2), count} (3 float64)
  * once per minibatch: all-reduce of the flat gradient buffer, whose tail carries the three loss
    partial sums; kernels already scale by 1/global_rows, so the SUM is the global-mean gradient and
    every rank applies the identical clip + Adam step (parameters stay bit-identical, no broadcast).
"""
import torch
import torch.distributed as dist


def world(process_group=None):
    """(world_size, rank) of the default / given process group; (1, 0) when not initialised."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(process_group), dist.get_rank(process_group)


def shard_envs(num_envs, rank, world_size):
    """Rank r owns envs [r*N/G, (r+1)*N/G); N must divide evenly so that all shards run the same kernels."""
    if num_envs % world_size != 0:
        raise ValueError(f"num_envs ({num_envs}) must be divisible by the number of ranks ({world_size})")
    n = num_envs // world_size
    return slice(rank * n, (rank + 1) * n)


def all_reduce_sum(t, process_group=None):
    """In-place SUM all-reduce (no-op on a single rank)."""
    ws, _ = world(process_group)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=process_group)
    return t


class _Done:
    def wait(self):
        return True


def all_reduce_sum_async(t, process_group=None):
    """In-place SUM all-reduce of a contiguous view, enqueued behind the work of the CURRENT stream; returns a handle whose
    ``wait()`` makes the then-current stream wait for the result (torch.distributed ``async_op=True``).  No-op on a single rank."""
    ws, _ = world(process_group)
    if ws > 1:
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=process_group, async_op=True)
    return _Done()


def equivalent_global_env_blocks(local_perms, n_local, num_mini_batch):
    """The env blocks a SINGLE process must use to see the same minibatches as G ranks that each drew
    `local_perms[r]` (a permutation of its n_local envs) in recurrent_generator: global minibatch k is the
    union over ranks of  r*n_local + perm_r[k*E_local : (k+1)*E_local].  Returns a list of int64 tensors."""
    e_local = n_local // num_mini_batch
    blocks = []
    for k in range(num_mini_batch):
        parts = [r * n_local + p[k * e_local:(k + 1) * e_local] for r, p in enumerate(local_perms)]
        blocks.append(torch.cat(parts))
    return blocks


def equivalent_global_sample_blocks(local_perms, T, n_local, num_mini_batch):
    """Feed-forward counterpart: local flat index i = t*n_local + n maps to global t*N + r*n_local + n."""
    G = len(local_perms)
    N = G * n_local
    mbs = (T * n_local) // num_mini_batch
    blocks = []
    for k in range(num_mini_batch):
        parts = []
        for r, p in enumerate(local_perms):
            idx = p[k * mbs:(k + 1) * mbs]
            t, n = idx // n_local, idx % n_local
            parts.append(t * N + r * n_local + n)
        blocks.append(torch.cat(parts))
    return blocks

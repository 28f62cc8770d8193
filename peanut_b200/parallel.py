"""Multi-GPU plumbing: the path shards over independent environments, one process per GPU.

There is no data-path collective - every environment's frame, maps and pose live on exactly one rank.  The only
exchange is the result gather to the rank that hosts the (CPU) planner, done with ``torch.distributed`` (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  The reference itself is single-process, single-GPU, batch 1
(nav/collect.py:32-33); this module is what replaces "run N copies by hand".
"""
import os

import torch
import torch.distributed as dist


def env_range(num_envs, world_size, rank):
    """Contiguous block partition: environment e lives on rank e // ceil(num_envs / world_size)
    (BASELINE.json configs[3]: 64 envs, 8 per GPU).  Returns (first, last_exclusive) for `rank`; ranks past the
    end get an empty range."""
    if num_envs < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad partition arguments")
    per = -(-num_envs // world_size) if num_envs else 0
    lo = min(rank * per, num_envs)
    return lo, min(lo + per, num_envs)


def owner_of(env, num_envs, world_size):
    per = -(-num_envs // world_size)
    if not (0 <= env < num_envs):
        raise IndexError(env)
    return env // per


def init_from_env(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, local_rank, world_size)."""
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, local, world


def max_over_ranks(values, device=None):
    """Element-wise maximum of a list of floats over all ranks (timing is reported as the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def gather_env_results(local, num_envs, dst=0):
    """local: this rank's per-environment results [E_local, ...] (block partition of `num_envs`).  Returns the
    [num_envs, ...] tensor on rank `dst` (None elsewhere).  Ranks may hold different numbers of environments."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-num_envs // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, out, dst=dst)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        lo, hi = env_range(num_envs, world, r)
        parts.append(out[r][:hi - lo])
    return torch.cat(parts, 0)

"""Env-batch sharding across the GPUs of one box.

Envs are independent, so the step path has NO collective: every rank owns a contiguous range of env indices, one
``VecWindFarmEnv`` (one handle, one stream, one state tensor) on its GPU.  The only exchange is the gather of
per-env evaluation statistics when FarmEval / AgentEval needs joint numbers (``AgentEval.py:363-475`` assembles
one dataset from all conditions): ``gather_env_stats`` below, one ``all_gather_into_tensor`` of a few KB.
Works with any ``torch.distributed`` backend (``nccl`` over NVLink on the GPU box, ``gloo`` in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_envs, rank, world):
    """Contiguous, balanced env range [lo, hi) of ``rank``; the first ``n_envs % world`` ranks get one extra env."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(int(n_envs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_envs, world):
    return [shard_range(n_envs, r, world)[1] - shard_range(n_envs, r, world)[0] for r in range(world)]


def env_seed(seed0, env_index):
    """Seed of one env of the global batch: independent of how the batch is sharded."""
    return int(seed0) + int(env_index)


def gather_env_stats(local, n_envs, group=None):
    """All-gather per-env rows ``local`` [n_local, ...] of every rank into the global [n_envs, ...] tensor
    (same order as the unsharded batch).  Uneven shards are padded to the largest shard for the collective."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_envs, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} envs, expected {sizes[rank]}")
    nmax = max(sizes)
    pad = local
    if local.shape[0] < nmax:
        pad = torch.cat([local, local.new_zeros((nmax - local.shape[0],) + tuple(local.shape[1:]))])
    out = local.new_empty((world * nmax,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    parts = [out[r * nmax:r * nmax + sizes[r]] for r in range(world)]
    return torch.cat(parts)


def max_over_ranks(value, device, group=None):
    """Device-timed durations are reported as the max over ranks (bench contract)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sample_conditions(cfg_wind, env_ids, n_turb, seed0=0, yaw_start=15.0, random_yaw=True):
    """Per-env (ws, ti, wd, yaw0) from ``default_rng(seed0 + env)`` in the reference draw order ws -> ti -> wd -> yaw
    (Wind_Farm_Env.py:564-568, :715).  Depends only on the global env index, never on the sharding."""
    n = len(env_ids)
    ws, ti, wd, yaw0 = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros((n, n_turb))
    for k, e in enumerate(env_ids):
        rng = np.random.default_rng(env_seed(seed0, e))
        ws[k] = rng.uniform(cfg_wind["ws_min"], cfg_wind["ws_max"])
        ti[k] = rng.uniform(cfg_wind["TI_min"], cfg_wind["TI_max"])
        wd[k] = rng.uniform(cfg_wind["wd_min"], cfg_wind["wd_max"])
        if random_yaw:
            yaw0[k] = rng.uniform(-yaw_start, yaw_start, n_turb)
    return ws, ti, wd, yaw0

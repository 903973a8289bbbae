"""CPU, world_size 2 over gloo: env sharding and the one gather of the N>1 path.  No GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from windgym_b200.sharding import gather_env_stats, max_over_ranks, sample_conditions, shard_range, shard_sizes


def test_shard_ranges_partition_the_batch():
    for n, w in ((4096, 8), (4096, 1), (10, 4), (3, 8), (1024, 3)):
        seen = []
        for r in range(w):
            lo, hi = shard_range(n, r, w)
            assert 0 <= lo <= hi <= n
            seen += list(range(lo, hi))
        assert seen == list(range(n))
        assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1
    assert shard_range(4096, 3, 8) == (1536, 2048)        # BASELINE.json cfg 3: 512 envs / GPU
    with pytest.raises(ValueError):
        shard_range(8, 8, 8)


def test_conditions_do_not_depend_on_the_sharding():
    wind = {"ws_min": 7, "ws_max": 15, "TI_min": 0.02, "TI_max": 0.15, "wd_min": 255, "wd_max": 285}
    full = sample_conditions(wind, range(10), 4)
    parts = [sample_conditions(wind, range(*shard_range(10, r, 3)), 4) for r in range(3)]
    for k in range(4):
        assert np.array_equal(full[k], np.concatenate([p[k] for p in parts]))
    # env 1 of the global batch with seed0 = 0 draws exactly what the reference draws for reset(seed=1)
    ws, ti, wd, yaw0 = sample_conditions(wind, [1], 4)
    assert ws[0] == 11.094572997602054 and wd[0] == 259.32478838158903


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_envs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_envs, rank, world)
        # per-env rows: [env index, env index squared, rank]
        local = torch.stack([torch.arange(lo, hi, dtype=torch.float32), torch.arange(lo, hi, dtype=torch.float32) ** 2,
                             torch.full((hi - lo,), float(rank))], dim=1)
        full = gather_env_stats(local, n_envs)
        t = max_over_ranks(1.0 + rank, "cpu")
        q.put((rank, full.numpy(), t))
        with pytest.raises(ValueError):
            gather_env_stats(local[:-1], n_envs)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_envs", [8, 7])
def test_gather_env_stats_world2_gloo(n_envs):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_envs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, full, t in res:
        assert full.shape == (n_envs, 3)
        assert np.array_equal(full[:, 0], np.arange(n_envs))
        assert np.array_equal(full[:, 1], np.arange(n_envs) ** 2)
        owners = np.concatenate([np.full(s, r) for r, s in enumerate(shard_sizes(n_envs, world))])
        assert np.array_equal(full[:, 2], owners)
        assert t == 2.0                                   # max over ranks of (1 + rank)


def test_gather_is_identity_without_process_group():
    x = torch.arange(6.0).reshape(3, 2)
    assert gather_env_stats(x, 3) is x
    assert max_over_ranks(3.5, "cpu") == 3.5


def _eval_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.fake_vec_env import FakeVecEnv
        from windgym_b200.agents import ConstantAgent
        from windgym_b200.evaluate import eval_batched
        wss, wds = [8.0, 10.0, 12.0], [265, 270, 275]          # 9 conditions over 2 ranks: 5 + 4
        lo, hi = shard_range(9, rank, world)
        env = FakeVecEnv(hi - lo, n_turb=3, eval_mode=True)
        ds = eval_batched(env, ConstantAgent([-10, 20, 0]), wss, wds, [0.05], t_sim=5)
        q.put((rank, ds["powerT_a"], ds.coords["time0"]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_eval_batched_sharded_over_two_ranks_equals_single_rank():
    """f-2 on N>1: conditions sharded across ranks, ONE gather assembles the reference-shaped dataset on every rank."""
    from tests.fake_vec_env import FakeVecEnv
    from windgym_b200.agents import ConstantAgent
    from windgym_b200.evaluate import eval_batched
    single = eval_batched(FakeVecEnv(9, n_turb=3, eval_mode=True), ConstantAgent([-10, 20, 0]), [8.0, 10.0, 12.0],
                          [265, 270, 275], [0.05], t_sim=5)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_eval_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, pT, t0 in res:
        assert np.array_equal(pT, single["powerT_a"]) and np.array_equal(t0, single.coords["time0"])

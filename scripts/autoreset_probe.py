"""Auto-reset throughput of the two pools on the bench workload: scripts/autoreset_probe.py [envs] [device|host|both]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from windgym_b200 import DevicePooledVecEnv, PooledVecEnv, V80  # noqa: E402
from windgym_b200.vector import GymVectorEnv  # noqa: E402


def leg(B, device_side, steps):
    cfg = bench.workload_config(4, 4, "Power_avg")
    dev = torch.device("cuda:0")
    cls = DevicePooledVecEnv if device_side else PooledVecEnv
    R = max(384 if device_side else 64, B // 8)
    extra = {}
    if device_side:
        if os.environ.get("POOL_STREAMS"):
            extra["n_streams"] = int(os.environ["POOL_STREAMS"])
        if os.environ.get("POOL_EVERY"):
            extra["refill_every"] = int(os.environ["POOL_EVERY"])
    pool = cls(V80(), B, reserve=R, config=cfg, device="cuda:0", n_passthrough=5, seed=0, **extra)
    genv = GymVectorEnv(venv=pool, as_torch=True)
    genv.reset(seed=0)
    acts = (torch.rand((64, B, 16), generator=torch.Generator().manual_seed(1)) * 2 - 1).cuda()
    tmax = pool.time_max.cpu().numpy() if torch.is_tensor(pool.time_max) else pool.time_max
    pool.state["timestep"][:] = torch.as_tensor((np.random.default_rng(0).uniform(0, 1, B) * tmax).astype(np.int32), device=dev)
    for i in range(120 if device_side else 10):
        genv.step(acts[i % 64])
    torch.cuda.synchronize()
    n = torch.zeros((), dtype=torch.int64, device=dev)
    t0 = time.perf_counter()
    for i in range(steps):
        _, _, _, tr, _ = genv.step(acts[i % 64])
        if device_side:
            n += tr.sum()
        else:
            n += int(tr.sum())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if os.environ.get("POOL_PROFILE"):      # per-kernel device time of the stepping kernels while the pool is working
        inner = pool.inner
        inner.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(256):
            genv.step(acts[i % 64])
        e1.record()
        torch.cuda.synchronize()
        fl, fi, n = inner.profile_read()
        inner.profile_enable(False)
        print(f"   while pooled: flow {1e3 * fl / n:.1f} us, finish {1e3 * fi / n:.1f} us per step; stepping stream busy "
              f"{1e3 * e0.elapsed_time(e1) / 256:.1f} us per step")
        # the swap + copy pair alone (nothing finished: flags cleared), back to back on the stepping stream
        inner.truncated.zero_()
        torch.cuda.synchronize()
        e0.record()
        for i in range(256):
            pool.lib.wg_pool_swap(inner._h, inner._step_ptrs[0], inner._step_ptrs[3], inner._step_ptrs[1],
                                  pool.swapped.data_ptr(), pool._final.data_ptr(), inner._stream())
        e1.record()
        torch.cuda.synchronize()
        print(f"   wg_pool_swap alone (no episode finished): {1e3 * e0.elapsed_time(e1) / 256:.1f} us per call")
    print(f"{'device' if device_side else 'host'} pool, {B} envs: {B * steps / dt / 1e6:.2f} M env-steps/s, {1e3 * dt / steps:.3f} ms/step, "
          f"episodes {int(n)}, stats {dict(pool.stats)}")
    pool.close()
    del genv, pool
    torch.cuda.empty_cache()


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    which = sys.argv[2] if len(sys.argv) > 2 else "both"
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
    if which in ("host", "both"):
        leg(B, False, min(steps, 300))
    if which in ("device", "both"):
        leg(B, True, steps)
    if which == "both":
        leg(B, False, min(steps, 300))

"""Eager steps vs the same steps captured in a CUDA graph (torch.cuda.CUDAGraph) and replayed:
scripts/graph_probe.py [envs] [nx] [ny]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from windgym_b200 import V80, VecWindFarmEnv  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    nx = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    ny = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    T, n, reps = nx * ny, 64, 12
    cfg = bench.workload_config(nx, ny, "Power_avg")
    env = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", n_passthrough=80, seed=0)
    env.reset(seed=0)
    acts = torch.rand((n, B, T), device="cuda:0") * 2 - 1
    for k in range(n):
        env.step(acts[k], info=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(reps):
        for k in range(n):
            env.step(acts[k], info=False)
    torch.cuda.synchronize()
    t_eager = (time.perf_counter() - t0) / (reps * n)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k in range(n):
            env.step(acts[k], info=False)
    g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(reps):
        g.replay()
    torch.cuda.synchronize()
    t_graph = (time.perf_counter() - t0) / (reps * n)
    env.check_flags()
    print(f"{B} envs, {nx}x{ny}: eager {1e6 * t_eager:.1f} us/step ({B / t_eager / 1e6:.2f} M env-steps/s) | graph of {n} steps "
          f"{1e6 * t_graph:.1f} us/step ({B / t_graph / 1e6:.2f} M env-steps/s)")
    env.close()


if __name__ == "__main__":
    main()

"""How much would cross-step overlap pay below one wave of CTAs?  N independent handles of 512/N envs, each on its own
stream, stepped round-robin without host synchronisation, against one handle of 512 envs:
scripts/overlap_probe.py [envs_total] [steps]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from windgym_b200 import V80, VecWindFarmEnv  # noqa: E402


def run(total, n_handles, steps):
    cfg = bench.workload_config(4, 4, "Power_avg")
    B = total // n_handles
    streams = [torch.cuda.Stream() for _ in range(n_handles)]
    envs, acts = [], []
    for i, s in enumerate(streams):
        with torch.cuda.stream(s):
            e = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", n_passthrough=50, seed=i)
            e.reset(seed=i)
            envs.append(e)
            acts.append((torch.rand((32, B, 16), device="cuda:0") * 2 - 1))
    torch.cuda.synchronize()
    for w in range(2):
        if w == 1:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        for k in range(steps if w else 30):
            for e, s, a in zip(envs, streams, acts):
                with torch.cuda.stream(s):
                    e.step(a[k % 32], info=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{n_handles} handle(s) x {B} envs: {total * steps / dt / 1e6:.2f} M env-steps/s, {1e6 * dt / steps:.1f} us per step of all handles")
    for e in envs:
        e.close()


if __name__ == "__main__":
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    for n in (1, 2, 4):
        run(total, n, steps)

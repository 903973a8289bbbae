"""Where does the end-to-end (host in / host out) step time go?  scripts/e2e_probe.py [envs]
Times, per step: the device-only step; VecWindFarmEnv.step_host (Python + C-ABI); the bare wg_step_host C call through
ctypes with prebuilt arguments; the same with WG_NO_ZEROCOPY semantics is a separate process run."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from windgym_b200 import V80, VecWindFarmEnv  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    T, n = 16, 400
    cfg = bench.workload_config(4, 4, "Power_avg")
    ws, ti, wd, yaw0 = bench.sample_conditions(cfg, np.arange(B), T)
    env = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", n_passthrough=bench.n_passthrough_for(4 * n, cfg), seed=0)
    env.reset(wind=(ws, ti, wd), yaw0=yaw0)
    acts = (torch.rand((n, B, T), generator=torch.Generator().manual_seed(1)) * 2 - 1).pin_memory()
    acts_d = acts.cuda()
    for i in range(20):
        env.step(acts_d[i])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        env.step(acts_d[i])
    torch.cuda.synchronize()
    t_dev = (time.perf_counter() - t0) / n
    for i in range(20):
        env.step_host(acts[i])
    t0 = time.perf_counter()
    for i in range(n):
        env.step_host(acts[i])
    t_py = (time.perf_counter() - t0) / n
    h = env._host
    stream = torch._C._cuda_getCurrentRawStream(0)
    fn = env.lib.wg_step_host
    ptrs = [acts[i].data_ptr() for i in range(n)]
    t0 = time.perf_counter()
    for i in range(n):
        fn(env._h, env._step_ptrs[0], ptrs[i], h["act_ptr"], h["out_ptr"], h["res_ptr"], h["n_res"], stream)
    t_c = (time.perf_counter() - t0) / n
    # launch-only cost of the device step (host time per call when the GPU is not waited for)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(50):
        env.step(acts_d[i])
    t_launch = (time.perf_counter() - t0) / 50
    torch.cuda.synchronize()
    print(f"envs {B}: device-only step {1e6 * t_dev:.1f} us | step_host (python) {1e6 * t_py:.1f} us | bare wg_step_host via "
          f"ctypes {1e6 * t_c:.1f} us | host time of an async env.step call {1e6 * t_launch:.1f} us")


if __name__ == "__main__":
    main()

"""Throughput with auto-reset (SURVEY.md 8d metric (ii)): naive masked reset vs the spare-env pool."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from windgym_b200 import V80, VecWindFarmEnv, PooledVecEnv
from windgym_b200.vector import GymVectorEnv
B, T = 4096, 16
cfg = bench.workload_config(4, 4, "Power_avg")
mode = sys.argv[1] if len(sys.argv) > 1 else "pool"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 400
if mode == "pool":
    venv = PooledVecEnv(V80(), B, reserve=int(sys.argv[3]) if len(sys.argv) > 3 else 512, config=cfg, device="cuda:0", n_passthrough=5, seed=0)
else:
    venv = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", n_passthrough=5, seed=0)
env = GymVectorEnv(venv=venv, as_torch=True)
t0 = time.perf_counter(); env.reset(seed=0); torch.cuda.synchronize(); print("full reset s", time.perf_counter() - t0)
acts = torch.rand((B, T), device="cuda:0") * 2 - 1
rng = np.random.default_rng(0)
venv.state["timestep"][:] = torch.as_tensor((rng.uniform(0, 1, B) * venv.time_max).astype(np.int32), device="cuda:0")
for i in range(20):
    env.step(acts)
ndone = 0
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(n):
    obs, r, term, trunc, infos = env.step(acts)
    ndone += int(trunc.sum())
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"autoreset {mode}: {B*n/dt:.0f} env-steps/s, {dt/n*1e3:.3f} ms/step, {ndone} resets in {n} steps", getattr(venv, "stats", ""))

"""Host-side profile of the auto-reset loop (PooledVecEnv behind GymVectorEnv): where does the wall time of a step go?"""
import cProfile, pstats, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from windgym_b200 import V80, PooledVecEnv
from windgym_b200.vector import GymVectorEnv
B, T = 4096, 16
cfg = bench.workload_config(4, 4, "Power_avg")
venv = PooledVecEnv(V80(), B, reserve=512, config=cfg, device="cuda:0", n_passthrough=5, seed=0)
env = GymVectorEnv(venv=venv, as_torch=True)
env.reset(seed=0)
acts = torch.rand((B, T), device="cuda:0") * 2 - 1
rng = np.random.default_rng(0)
venv.state["timestep"][:] = torch.as_tensor((rng.uniform(0, 1, B) * venv.time_max).astype(np.int32), device="cuda:0")
for i in range(30):
    env.step(acts)
torch.cuda.synchronize()
n = 300
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for i in range(n):
    obs, r, term, trunc, infos = env.step(acts)
pr.disable()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"{B*n/dt:.0f} env-steps/s, {dt/n*1e3:.3f} ms/step", venv.stats)
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(28)

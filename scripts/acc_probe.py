"""Accuracy probe: deep-array rotors of a 64-turbine farm and a near-cut-in 4x4 case against the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import oracle_rollout, small_config
from windgym_b200 import V80, VecWindFarmEnv
def run(nx, ny, ws, wd, ti, kw, steps=2):
    T = nx * ny
    cfg = small_config(nx, ny, reward="Power_avg", action="yaw")
    B = len(ws)
    yaw0 = np.random.default_rng(1).uniform(-15, 15, (B, T))
    acts = np.random.default_rng(2).uniform(-1, 1, (steps, B, T * (2 if kw.get("induction_control") else 1))).astype(np.float32)
    env = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", **kw)
    env.reset(wind=(ws, ti, wd), yaw0=yaw0)
    P, U = [], []
    for a in acts:
        env.step(torch.as_tensor(a)); P.append(env.state["power"][:, 0].cpu().numpy().copy()); U.append(env.state["u"][:, 0].cpu().numpy().copy())
    ref = oracle_rollout(cfg, ws, ti, wd, yaw0, acts, **kw)
    P, U = np.array(P).transpose(1, 0, 2), np.array(U).transpose(1, 0, 2)
    dp = np.abs(P - ref["power"]); rel = dp / np.maximum(ref["power"], 2e5); du = np.abs(U - ref["u"])
    print(f"{nx}x{ny}: max |dP|/max(P,10% rated) {rel.max():.3e}   max |du|/ws {(du / ws[:, None, None]).max():.3e}   mean |du|/ws {(du / ws[:, None, None]).mean():.3e}")
run(8, 8, np.array([9.0, 12.0]), np.array([268.0, 272.0]), np.array([0.06, 0.1]), dict(fill_window=2, induction_control=True))
run(4, 4, np.array([7.2312, 8.5]), np.array([271.5082, 270.0]), np.array([0.0348, 0.05]), dict())

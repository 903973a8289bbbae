"""Diagnose one env of the full-size parity test against the oracle: scripts/diag_env.py <env index>"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import small_config  # noqa: E402
from windgym_b200 import V80, VecWindFarmEnv  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 1561
B, T, steps = 4096, 16, 3
rng = np.random.default_rng(2024)
ws, ti, wd, yaw0 = rng.uniform(7, 15, B), rng.uniform(0.02, 0.15, B), rng.uniform(255, 285, B), rng.uniform(-15, 15, (B, T))
acts = np.random.default_rng(7).uniform(-1, 1, (steps, B, T)).astype(np.float32)
cfg = small_config(4, 4, reward="Power_avg", action="wind")
sel = [b]
print("env", b, "ws %.4f ti %.4f wd %.4f" % (ws[b], ti[b], wd[b]))
env = VecWindFarmEnv(V80(), 1, config=cfg, device="cuda:0")
env.reset(wind=(ws[sel], ti[sel], wd[sel]), yaw0=yaw0[sel])
from oracle.env_numpy import WindFarmEnvOracle
from oracle.v80 import V80 as OV80
o = WindFarmEnvOracle(OV80(), cfg, reset_init=False)
o.reset(wind=(ws[b], ti[b], wd[b]), yaw0=yaw0[b])
np.set_printoptions(precision=6, linewidth=200)
print("after reset: u gpu - u oracle:", env.state["u"][0, 0].cpu().numpy() - o.fs.rotor_avg_windspeed[:, 0])
print("count gpu", env.state["count"][0, 0].cpu().numpy(), "retire", env.state["retire"][0, 0].cpu().numpy())
print("count ora", o.fs.count)
for k in range(steps):
    env.step(torch.as_tensor(acts[k, sel]))
    o.step(acts[k, b])
    ug, uo = env.state["u"][0, 0].cpu().numpy().astype(np.float64), o.fs.rotor_avg_windspeed[:, 0]
    pg, po = env.state["power"][0, 0].cpu().numpy().astype(np.float64), o.fs.windTurbines.power()
    rel = np.abs(pg - po) / np.maximum(po, 1.0)
    t = int(np.argmax(rel))
    print(f"step {k}: worst turbine {t}: rel P err {rel[t]:.3e}; u gpu {ug[t]:.6f} oracle {uo[t]:.6f} (du {ug[t] - uo[t]:.2e}); "
          f"P gpu {pg[t]:.1f} oracle {po[t]:.1f}; yaw gpu {env.state['yaw'][0, 0, t].item():.5f} oracle {o.fs.windTurbines.yaw[t]:.5f}")
    print("   du all:", ug - uo)
    print("   count gpu", env.state["count"][0, 0].cpu().numpy() - env.state["retire"][0, 0].cpu().numpy(), " oracle", o.fs.count)
print("xr", env.state["xr"][0].cpu().numpy())

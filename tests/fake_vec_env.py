"""A CPU stand-in with the ``VecWindFarmEnv`` protocol for host-logic tests of the adapters (no CUDA, no physics):
power = ws^3 * cos(yaw)^2 per turbine, observations = [ws/25, yaw/45] per turbine, episodes of ``horizon`` steps."""
import types

import numpy as np
import torch


class FakeVecEnv:
    def __init__(self, n_envs, n_turb=3, horizon=5, baseline=False, eval_mode=False):
        self.n_envs, self.n_turb, self.horizon = n_envs, n_turb, horizon
        self.obs_var = 2 * n_turb
        self.obs_shape = (n_envs, self.obs_var)
        self.device = torch.device("cpu")
        self.Baseline_comp = baseline
        self.n_farms = 2 if baseline else 1
        self.yaw_min, self.yaw_max, self.yaw_step = -45, 45, 1
        self.seed = None
        self.eval_mode = eval_mode
        self.ec = types.SimpleNamespace(dt_sim=1, dt_env=1, turbtype="None", S=1)
        B, T, F = n_envs, n_turb, self.n_farms
        z = lambda *s: torch.zeros(s, dtype=torch.float32)
        self.state = {"yaw": z(B, F, T), "u": z(B, F, T), "v": z(B, F, T), "w": z(B, F, T), "power": z(B, F, T),
                      "n_step": torch.zeros((B, F), dtype=torch.int32), "timestep": torch.zeros(B, dtype=torch.int32),
                      "meas": z(B, 4, T), "xr": z(B, T), "yr": z(B, T)}
        self.ws, self.ti, self.wd = np.full(B, 10.0), np.full(B, 0.05), np.full(B, 270.0)
        self._over = {}
        self.obs = z(B, self.obs_var)
        self.reward = z(B)
        self.truncated = torch.zeros(B, dtype=torch.uint8)
        self.terminated = torch.zeros(B, dtype=torch.bool)
        self.n_resets = np.zeros(B, dtype=np.int64)
        self.closed = False

    def set_wind_vals(self, ws=None, ti=None, wd=None):
        for k, v in (("ws", ws), ("ti", ti), ("wd", wd)):
            if v is not None:
                self._over[k] = np.broadcast_to(np.asarray(v, dtype=np.float64), (self.n_envs,)).copy()

    def _refresh(self):
        s = self.state
        ws = torch.as_tensor(self.ws, dtype=torch.float32)[:, None, None]
        s["u"][:] = ws
        s["power"][:] = ws ** 3 * torch.cos(torch.deg2rad(s["yaw"])) ** 2
        self.obs[:] = torch.cat([s["u"][:, 0] / 25.0, s["yaw"][:, 0] / 45.0], dim=1)

    def reset(self, seed=None, mask=None, wind=None, yaw0=None):
        sel = np.arange(self.n_envs) if mask is None else np.flatnonzero(np.asarray(mask))
        for k in ("ws", "ti", "wd"):
            if k in self._over:
                getattr(self, k)[sel] = self._over[k][sel]
        y0 = np.zeros((self.n_envs, self.n_turb)) if yaw0 is None else np.broadcast_to(yaw0, (self.n_envs, self.n_turb))
        for f in range(self.n_farms):
            self.state["yaw"][sel, f] = torch.as_tensor(np.ascontiguousarray(y0[sel]), dtype=torch.float32)
        self.state["timestep"][sel] = 0
        self.state["n_step"][sel] = torch.as_tensor((100 + self.ws[sel]).astype(np.int32))[:, None]
        self.n_resets[sel] += 1
        self._refresh()
        return self.obs, self._info()

    def step(self, actions):
        a = torch.as_tensor(np.asarray(actions, dtype=np.float32)) if not torch.is_tensor(actions) else actions
        a = a.reshape(self.n_envs, self.n_turb).to(torch.float32)
        s = self.state
        tgt = (a + 1) / 2 * 90 - 45
        s["yaw"][:, 0] = torch.minimum(torch.maximum(tgt, s["yaw"][:, 0] - 1), s["yaw"][:, 0] + 1)
        s["n_step"] += 1
        self._refresh()
        self.reward[:] = s["power"][:, 0].sum(1) / 1000.0
        ts = s["timestep"]
        self.truncated[:] = (ts >= self.horizon - 1).to(torch.uint8) if not self.eval_mode else 0
        ts += 1
        return self.obs, self.reward, self.terminated, self.truncated, self._info()

    def check_flags(self):
        pass

    def _info(self):
        s = self.state
        d = {"yaw angles agent": s["yaw"][:, 0], "Wind speed Global": self.ws, "Wind direction Global": self.wd,
             "Turbulence intensity": self.ti, "Power pr turbine agent": s["power"][:, 0],
             "Wind speed at turbines": s["meas"][:, 0], "Wind direction at turbines": s["meas"][:, 1],
             "Turbine x positions": s["xr"], "Turbine y positions": s["yr"]}
        if self.Baseline_comp:
            d["yaw angles base"] = s["yaw"][:, 1]
            d["Power pr turbine baseline"] = s["power"][:, 1]
            d["Wind speed at turbines baseline"] = s["u"][:, 1]
        return d

    def close(self):
        self.closed = True

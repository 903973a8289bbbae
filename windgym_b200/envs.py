"""Single-env facades with the reference's class surface, backed by a batch-of-one ``VecWindFarmEnv``.

``WindFarmEnv`` (``WindGym/Wind_Farm_Env.py:47``), ``FarmEval`` (``WindGym/FarmEval.py:10``) and
``WindFarmEnvMulti`` (``WindGym/WindEnvMulti.py:17``): same constructor arguments, ``reset()/step()`` return
shapes and dtypes, info keys, and the attributes callers read (``AgentEval.py:83-209``, ``GreedyAgent.py:36``):
``env.fs.windTurbines.{yaw, power(), rotor_avg_windspeed, positions_xyz}``, ``env.fs.time``, ``env.fs_baseline``,
``env.farm_measurements.get_*``.  All compute runs in the CUDA library; these classes only move one env's
numbers to the host.  For throughput use ``VecWindFarmEnv`` directly.

Conscious differences from the reference (SURVEY.md quirk ledger): ``turbtype`` defaults to ``"None"``; the Mann site
types share ONE box per env object (generated at construction, not at every reset); ``"Random"`` raises
``NotImplementedError``; after truncation the env stays
usable (Q11); ``WindFarmEnvMulti`` constructs (Q9-i), declares the observation length it actually returns (Q9-ii)
and keeps the reference's double ``timestep`` increment only with ``compat_double_timestep=True`` (Q9-iii).
"""
import numpy as np
import torch

from .vec_env import VecWindFarmEnv

try:  # gymnasium / pettingzoo are optional: agents only need Box-like spaces
    from gymnasium import Env as _GymEnv
    from gymnasium.spaces import Box
except Exception:  # pragma: no cover - exercised in this image (gymnasium absent)
    class _GymEnv:
        metadata = {}

    class Box:
        """Minimal stand-in for ``gymnasium.spaces.Box`` (low/high/shape/dtype/sample/contains)."""

        def __init__(self, low, high, shape, dtype=np.float32, seed=None):
            self.low = np.full(shape, low, dtype=dtype)
            self.high = np.full(shape, high, dtype=dtype)
            self.shape, self.dtype = tuple(shape), np.dtype(dtype)
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)


def _window_bounds(L, N, W, i):
    """Rolling window i of a history of length L (MesClass.py:85-116)."""
    if i == 0:
        return max(0, L - W), L
    if i == N - 1 and L >= W:
        return 0, W
    if L < W:
        return 0, L
    spacing = max(1, (L - W) // (N - 1))
    pos = min(i * spacing, L - W)
    return pos, pos + W


class _MeasurementView:
    """Host view of the device ring buffers with ``farm_mes``'s getters (MesClass.py:620-677), unscaled values."""

    def __init__(self, venv, b=0):
        self.v, self.b = venv, b
        ec = venv.ec
        self.ch = {"ws": ec.ws_mes, "wd": ec.wd_mes, "yaw": ec.yaw_mes, "power": ec.power_mes}
        self.gate = {"ws": ec.mes_level["turb_ws"], "wd": ec.mes_level["turb_wd"], "yaw": True,
                     "power": ec.mes_level["turb_power"]}
        self.fgate = {"ws": ec.mes_level["farm_ws"], "wd": ec.mes_level["farm_wd"], "power": ec.mes_level["farm_power"]}
        T = venv.n_turb
        self.off, o = {}, 0
        for c in ("ws", "wd", "yaw", "power"):
            H = self.ch[c][f"{c}_history_length"]
            self.off[c] = [o + t * H for t in range(T)]
            o += T * H
        for c in ("ws", "wd", "power"):
            self.off["farm_" + c] = o
            o += self.ch[c][f"{c}_history_length"]

    def _history(self, c, off):
        H = self.ch[c][f"{c}_history_length"]
        n = int(self.v.state["n_push"][self.b])
        ring = self.v.state["rings"][self.b, off:off + H].cpu().numpy().astype(np.float64)
        L = min(n, H)
        return ring[(n - L + np.arange(L)) % H]

    def _mes(self, c, off, gate):
        m = self.ch[c]
        cur, roll = m[f"{c}_current"] and gate, m[f"{c}_rolling_mean"] and gate
        hist = self._history(c, off)
        out = []
        if hist.size == 0:
            return np.array(out, dtype=np.float32)
        if cur:
            out.append(hist[-1])
        if roll:
            N, W = m[f"{c}_history_N"], m[f"{c}_window_length"]
            for i in range(N):
                lo, hi = _window_bounds(hist.size, N, W, i)
                out.append(np.mean(hist[lo:hi]))
        return np.array(out, dtype=np.float32)

    def _turb(self, c):
        return np.array([self._mes(c, o, self.gate[c]) for o in self.off[c]]).flatten()

    def get_ws_turb(self, scaled=False):
        return self._turb("ws")

    def get_wd_turb(self, scaled=False):
        return self._turb("wd")

    def get_yaw_turb(self, scaled=False):
        return self._turb("yaw")

    def get_power_turb(self, scaled=False):
        return self._turb("power")

    def get_ws_farm(self, scaled=False):
        return self._mes("ws", self.off["farm_ws"], self.fgate["ws"])

    def get_wd_farm(self, scaled=False):
        return self._mes("wd", self.off["farm_wd"], self.fgate["wd"])

    def get_power_farm(self, scaled=False):
        return self._mes("power", self.off["farm_power"], self.fgate["power"])

    def get_TI_turb(self, scaled=False):
        out = []
        for o in self.off["ws"]:
            u = self._history("ws", o)
            out.append(np.float32(np.std(u - u.mean()) / u.mean()) if u.size else np.float32(0))
        return np.array(out, dtype=np.float32)

    def get_TI(self, scaled=False):
        return np.array([self.get_TI_turb().mean()], dtype=np.float32)

    def observed_variables(self):
        return self.v.obs_var

    def max_hist(self):
        return self.v.ec.hist_max


class _TurbineView:
    """``fs.windTurbines`` (dynamiks ``PyWakeWindTurbines`` protocol)."""

    def __init__(self, venv, farm, b=0):
        self.v, self.f, self.b = venv, farm, b
        self.types = np.zeros(venv.n_turb, dtype=int)

    @property
    def yaw(self):
        return self.v.state["yaw"][self.b, self.f].cpu().numpy().astype(np.float64)

    @yaw.setter
    def yaw(self, value):
        val = np.broadcast_to(np.asarray(value, dtype=np.float32).reshape(-1), (self.v.n_turb,)) \
            if np.size(value) == 1 else np.asarray(value, dtype=np.float32)
        self.v.state["yaw"][self.b, self.f] = torch.as_tensor(np.ascontiguousarray(val), device=self.v.device)

    def power(self):
        return self.v.state["power"][self.b, self.f].cpu().numpy().astype(np.float64)

    @property
    def rotor_avg_windspeed(self):
        s = self.v.state
        return torch.stack([s["u"][self.b, self.f], s["v"][self.b, self.f], s["w"][self.b, self.f]], dim=1) \
            .cpu().numpy().astype(np.float64)

    @property
    def positions_xyz(self):
        s = self.v.state
        T = self.v.n_turb
        return np.stack([s["xr"][self.b].cpu().numpy().astype(np.float64),
                         s["yr"][self.b].cpu().numpy().astype(np.float64), np.full(T, self.v.ec.hub_height)])

    rotor_positions_xyz = positions_xyz

    def yaw_tilt(self):
        return self.yaw, np.zeros(self.v.n_turb)

    def hub_height(self):
        return self.v.ec.hub_height

    def diameter(self):
        return self.v.ec.D


class _FlowView:
    """``env.fs`` / ``env.fs_baseline``: the dynamiks ``DWMFlowSimulation`` members the env layer and AgentEval read."""

    def __init__(self, venv, farm, b=0):
        self.v, self.f, self.b = venv, farm, b
        self.windTurbines = _TurbineView(venv, farm, b)

    @property
    def time(self):
        return float(int(self.v.state["n_step"][self.b, self.f]) * self.v.ec.dt_sim)

    @property
    def wind_direction(self):
        return float(self.v.wd[self.b])

    def step(self):
        self.v.flow_steps(1)

    def run(self, t):
        self.v.flow_steps(int(round(t / self.v.ec.dt_sim)))

    def get_windspeed(self, view, include_wakes=True, xarray=False):
        """dynamiks ``get_windspeed`` for an ``XYView``-like object (attributes ``x``, ``y``, ``z``): array
        [3, len(x), len(y)] of (u, v, w).  ``xarray=True`` returns a ``FieldArray`` with the ``.x/.y.values`` and
        ``[component]`` access ``_render_frame`` uses (Wind_Farm_Env.py:1056-1063)."""
        x, y = np.asarray(view.x, dtype=np.float64), np.asarray(view.y, dtype=np.float64)
        z = float(np.asarray(getattr(view, "z", self.v.ec.hub_height)).reshape(-1)[0])
        if include_wakes:
            uvw = self.v.flow_field(x, y, z, env=self.b, farm=self.f).cpu().numpy().astype(np.float64)
        else:
            uvw = np.zeros((3, x.size, y.size))
            uvw[0] = float(self.v.ws[self.b])
        return FieldArray(uvw, x, y) if xarray else uvw


class XYView:
    """``dynamiks.views.XYView`` stand-in: a horizontal plane ``x`` x ``y`` at height ``z`` (wind-aligned frame)."""

    def __init__(self, x, y, z, ax=None, adaptive=False):
        self.x, self.y, self.z, self.ax, self.adaptive = np.asarray(x), np.asarray(y), z, ax, adaptive


class _Coord:
    def __init__(self, values):
        self.values = values


class FieldArray:
    """What ``get_windspeed(..., xarray=True)`` returns without xarray: ``uvw[0]`` is the u field [x, y]."""

    def __init__(self, values, x, y):
        self.values, self.x, self.y = values, _Coord(x), _Coord(y)

    def __getitem__(self, i):
        return self.values[i]


def _viridis(t):
    """Tiny viridis-like colour map (5 knots, linear): t in [0, 1] -> uint8 RGB."""
    knots = np.array([[68, 1, 84], [59, 82, 139], [33, 145, 140], [94, 201, 98], [253, 231, 37]], dtype=np.float64)
    t = np.clip(t, 0.0, 1.0) * (len(knots) - 1)
    i = np.minimum(t.astype(int), len(knots) - 2)
    f = (t - i)[..., None]
    return (knots[i] * (1 - f) + knots[i + 1] * f).astype(np.uint8)


class WindFarmEnv(_GymEnv):
    """Drop-in for ``WindGym.WindFarmEnv`` (Wind_Farm_Env.py:47-70): one env, numpy in / numpy out."""

    metadata = {"render_modes": ["human", "rgb_array"]}
    _eval_mode = False
    _multi_agent = False

    def __init__(self, turbine, n_passthrough=5, TI_min_mes=0.0, TI_max_mes=0.50, TurbBox="Default", turbtype="None",
                 yaml_path=None, Baseline_comp=False, yaw_init=None, render_mode=None, seed=None, dt_sim=1, dt_env=1,
                 yaw_step=1, fill_window=True, sample_site=None, HTC_path=None, reset_init=True, config=None,
                 device="cuda:0", turb_box=None, added_turbulence=None, induction_control=False, derate_min=0.5):
        if HTC_path is not None:
            raise NotImplementedError("HAWC2 turbines (HTC_path) are out of scope: external aero-elastic co-simulation")
        if render_mode is not None and render_mode not in self.metadata["render_modes"]:
            raise ValueError(f"render_mode must be one of {self.metadata['render_modes']}")
        self.render_mode = render_mode
        self.vec = VecWindFarmEnv(turbine, 1, yaml_path=yaml_path, config=config, n_passthrough=n_passthrough,
                                  TI_min_mes=TI_min_mes, TI_max_mes=TI_max_mes, TurbBox=TurbBox, turbtype=turbtype,
                                  Baseline_comp=Baseline_comp, yaw_init=yaw_init, seed=seed, dt_sim=dt_sim,
                                  dt_env=dt_env, yaw_step=yaw_step, fill_window=fill_window, device=device,
                                  multi_agent=self._multi_agent, eval_mode=self._eval_mode, sample_site=sample_site,
                                  turb_box=turb_box, added_turbulence=added_turbulence,
                                  induction_control=induction_control, derate_min=derate_min)
        self.sample_site = sample_site
        v, ec = self.vec, self.vec.ec
        self.turbine, self.seed = turbine, seed
        self.n_turb, self.x_pos, self.y_pos = ec.n_turb, ec.x_pos, ec.y_pos
        self.yaw_min, self.yaw_max, self.yaw_step = ec.yaw_min, ec.yaw_max, yaw_step
        self.ws_min, self.ws_max, self.wd_min, self.wd_max = ec.ws_min, ec.ws_max, ec.wd_min, ec.wd_max
        self.TI_min, self.TI_max = ec.TI_min, ec.TI_max
        self.Baseline_comp, self.ActionMethod, self.BaseController = ec.Baseline_comp, ec.ActionMethod, ec.BaseController
        self.dt_sim, self.dt_env, self.sim_steps_per_env_step = dt_sim, dt_env, ec.S
        self.act_var = ec.act_var   # 1 in the reference (Wind_Farm_Env.py:97-99); 2 with induction_control
        self.obs_var = v.obs_var
        self.hist_max, self.steps_on_reset = ec.hist_max, ec.steps_on_reset
        self.maxturbpower = ec.maxturbpower
        self.fs = _FlowView(v, 0)
        if self.Baseline_comp:
            self.fs_baseline = _FlowView(v, 1)
        self.farm_measurements = _MeasurementView(v)
        self.timestep, self.time_max = 0, 0
        self.ws = self.ti = self.wd = None
        self._init_spaces()
        if reset_init:
            self.reset(seed=seed)

    def _init_spaces(self):
        self.observation_space = Box(low=-1.0, high=1.0, shape=(self.obs_var,), dtype=np.float32)
        self.action_space = Box(low=-1, high=1, shape=(self.n_turb * self.act_var,), dtype=np.float32)

    # -------------------------------------------------------------------------------------------- helpers
    def _sync_scalars(self):
        v = self.vec
        self.ws, self.ti, self.wd = float(v.ws[0]), float(v.ti[0]), float(v.wd[0])
        self.time_max = int(v.time_max[0])
        self.rated_power = float(np.asarray(self.turbine.power(self.ws)))

    def _get_info(self):
        s = self.vec.state
        meas = s["meas"][0].cpu().numpy().astype(np.float64)
        self.current_ws, self.current_wd, self.current_yaw = meas[0], meas[1], self.fs.windTurbines.yaw
        fm = self.farm_measurements
        pw = self.fs.windTurbines.power()
        pos = self.fs.windTurbines.positions_xyz
        info = {
            "yaw angles agent": self.current_yaw,
            "yaw angles measured": fm.get_yaw_turb(),
            "Wind speed Global": self.ws,
            "Wind speed at turbines": self.current_ws,
            "Wind speed at turbines measured": fm.get_ws_turb(),
            "Wind speed at farm measured": fm.get_ws_farm(),
            "Wind direction Global": self.wd,
            "Wind direction at turbines": self.current_wd,
            "Wind direction at turbines measured": fm.get_wd_turb(),
            "Wind direction at farm measured": fm.get_wd_farm(),
            "Turbulence intensity": self.ti,
            "Power agent": pw.sum(),
            "Power pr turbine agent": pw,
            "Turbine x positions": pos[0],
            "Turbine y positions": pos[1],
        }
        if self.Baseline_comp:
            wt = self.fs_baseline.windTurbines
            pb = wt.power()
            info["yaw angles base"] = wt.yaw
            info["Power baseline"] = pb.sum()
            info["Power pr turbine baseline"] = pb
            info["Wind speed at turbines baseline"] = wt.rotor_avg_windspeed[:, 0]
        return info

    # -------------------------------------------------------------------------------------------- gym API
    def reset(self, seed=None, options=None):
        """Wind_Farm_Env.py:680-802.  Returns (obs float32[obs_var], info)."""
        if seed is not None:
            # gymnasium re-seeds np_random: same seed -> same episode (check_env determinism); later unseeded resets
            # continue that stream (VecWindFarmEnv.reset stores the seed and rewinds its episode counter)
            self.seed = seed
        obs, _ = self.vec.reset(seed=seed if seed is not None else self._next_seed())
        self.vec.check_flags()
        self.timestep = 0
        self._sync_scalars()
        return self._obs_numpy(obs), self._get_info()

    def _next_seed(self):
        # unseeded reset: continue the stream of the last seed (episode counter advances inside VecWindFarmEnv)
        return self.vec.seed

    def _obs_numpy(self, obs):
        return obs[0].cpu().numpy().astype(np.float32)

    def step(self, action):
        """Wind_Farm_Env.py:920-1034.  Returns (obs, reward, terminated=False, truncated, info)."""
        a = np.asarray(action, dtype=np.float32).reshape(1, -1)
        if a.shape[1] != self.n_turb * self.act_var:
            raise ValueError(f"action must have {self.n_turb * self.act_var} entries")
        obs, rew, _, trunc, _ = self.vec.step(torch.as_tensor(a))
        fl = int(self.vec.state["flags"][0])
        if fl & 1:
            raise Exception("NaN Power")  # Wind_Farm_Env.py:980-981
        self.vec.check_flags()
        self.timestep += 1
        return self._obs_numpy(obs), float(rew[0]), False, bool(trunc[0]), self._get_info()

    def init_render(self):
        """Wind_Farm_Env.py:464-478: a 250 x 250 hub-height view around the (rotated) farm."""
        x_turb, y_turb = self.fs.windTurbines.positions_xyz[:2]
        self.a = np.linspace(-200 + min(x_turb), 1000 + max(x_turb), 250)
        self.b = np.linspace(-200 + min(y_turb), 200 + max(y_turb), 250)
        self.view = XYView(z=self.turbine.hub_height(), x=self.a, y=self.b, adaptive=False)

    def render(self):
        if self.render_mode == "rgb_array":
            return self._render_frame()

    def _render_frame(self, baseline=False):
        """Wind_Farm_Env.py:1040-1083 without matplotlib: the u field of the hub-height view as an RGB image
        [len(y), len(x), 3] uint8 (row 0 = smallest y), rotors drawn as black segments across their yawed planes."""
        fs_use = self.fs_baseline if baseline else self.fs
        self.init_render()
        uvw = fs_use.get_windspeed(self.view, include_wakes=True, xarray=True)
        u = uvw[0].T                                              # [y, x] like pcolormesh(x, y, u.T)
        img = _viridis((u - 0.3 * self.ws) / (0.8 * self.ws))
        wt = fs_use.windTurbines
        xt, yt = wt.positions_xyz[:2]
        R = 0.5 * self.turbine.diameter()
        for x0, y0, g in zip(xt, yt, np.deg2rad(wt.yaw)):
            for s in np.linspace(-R, R, 41):                      # rotor plane: normal turned by the yaw offset
                xi = np.searchsorted(self.a, x0 + s * np.sin(g))
                yi = np.searchsorted(self.b, y0 + s * np.cos(g))
                if 0 <= xi < self.a.size and 0 <= yi < self.b.size:
                    img[yi, xi] = 0
        return img

    def plot_frame(self, baseline=False):
        return self._render_frame(baseline=baseline)

    def close(self):
        self.vec.close()


class FarmEval(WindFarmEnv):
    """Drop-in for ``WindGym.FarmEval`` (FarmEval.py:10-90): pinned wind / yaw, never truncates."""

    _eval_mode = True

    def __init__(self, turbine, TI_min_mes=0.0, TI_max_mes=0.50, yaw_init="Zeros", TurbBox="Default", yaml_path=None,
                 Baseline_comp=False, render_mode=None, turbtype="None", seed=None, dt_sim=1, dt_env=1, yaw_step=1,
                 n_passthrough=5, HTC_path=None, reset_init=True, config=None, device="cuda:0", turb_box=None):
        super().__init__(turbine, n_passthrough=n_passthrough, TI_min_mes=TI_min_mes, TI_max_mes=TI_max_mes,
                         TurbBox=TurbBox, turbtype=turbtype, yaml_path=yaml_path, Baseline_comp=Baseline_comp,
                         yaw_init=yaw_init, render_mode=render_mode, seed=seed, dt_sim=dt_sim, dt_env=dt_env,
                         yaw_step=yaw_step, HTC_path=HTC_path, reset_init=reset_init, config=config, device=device,
                         turb_box=turb_box)

    def set_wind_vals(self, ws=None, ti=None, wd=None):
        self.vec.set_wind_vals(ws=ws, ti=ti, wd=wd)
        for k, val in (("ws", ws), ("ti", ti), ("wd", wd)):
            if val is not None:
                setattr(self, k, val)

    def set_yaw_vals(self, yaw_vals):
        self.vec.set_yaw_vals(yaw_vals)

    def update_tf(self, path):
        """FarmEval.update_tf (FarmEval.py:86-90): pin the turbulence box file used from the next reset on."""
        if self.vec.ec.turbtype == "None":
            raise ValueError("update_tf needs a Mann turbtype (MannLoad)")
        self.vec._attach_turbulence(None, path)


class WindFarmEnvMulti(WindFarmEnv):
    """Drop-in for ``WindGym.WindFarmEnvMulti`` (WindEnvMulti.py:17-249): PettingZoo parallel API, one agent per
    turbine, dict in / dict out, one shared scalar reward."""

    metadata = {"name": "MultiFarm_environment_v0"}
    _multi_agent = True

    def __init__(self, turbine, n_passthrough=20, TI_min_mes=0.0, TI_max_mes=0.50, TurbBox="Default", turbtype="None",
                 yaml_path=None, Baseline_comp=False, yaw_init=None, render_mode=None, seed=None, dt_sim=1, dt_env=1,
                 yaw_step=1, fill_window=True, sample_site=None, config=None, device="cuda:0",
                 compat_double_timestep=False, turb_box=None):
        self.compat_double_timestep = compat_double_timestep
        self.possible_agents, self.agents = [], []
        super().__init__(turbine, n_passthrough=n_passthrough, TI_min_mes=TI_min_mes, TI_max_mes=TI_max_mes,
                         TurbBox=TurbBox, turbtype=turbtype, yaml_path=yaml_path, Baseline_comp=Baseline_comp,
                         yaw_init=yaw_init, render_mode=None, seed=seed, dt_sim=dt_sim, dt_env=dt_env,
                         yaw_step=yaw_step, fill_window=fill_window, sample_site=sample_site, reset_init=False,
                         config=config, device=device, turb_box=turb_box)
        self.possible_agents = ["turbine_" + str(r) for r in range(self.n_turb)]
        self.agent_name_mapping = dict(zip(self.possible_agents, range(self.n_turb)))
        self.reset(seed=seed)

    def _init_spaces(self):
        self._obs_space = Box(low=-1.0, high=1.0, shape=(self.obs_var,), dtype=np.float32)
        self._act_space = Box(low=-1.0, high=1.0, shape=(self.act_var,), dtype=np.float32)

    def observation_space(self, agent):
        return self._obs_space

    def action_space(self, agent):
        return self._act_space

    def _obs_numpy(self, obs):
        o = obs[0].cpu().numpy().astype(np.float32)  # [T, obs_var]
        return {a: o[i] for a, i in self.agent_name_mapping.items() if a in self.agents}

    def _get_infos(self):
        base = WindFarmEnv._get_info(self)
        fm = self.farm_measurements
        per = {c: [fm._mes(c, o, fm.gate[c]) for o in fm.off[c]] for c in ("ws", "wd", "yaw")}
        infos = {}
        for a in self.agents:
            i = self.agent_name_mapping[a]
            infos[a] = {
                "yaw angles agent": base["yaw angles agent"][i],
                "yaw angles measured": per["yaw"][i],
                "Wind speed Global": self.ws,
                "Wind speed at turbine": base["Wind speed at turbines"][i],
                "Wind speed at turbine measured": per["ws"][i],
                "Wind direction Global": self.wd,
                "Wind direction at turbine": base["Wind direction at turbines"][i],
                "Wind direction at turbine measured": per["wd"][i],
                "Wind direction at farm measured": base["Wind direction at farm measured"],
                "Turbulence intensity": self.ti,
                "Power agent": base["Power agent"],
                "Power turbine agent": base["Power pr turbine agent"][i],
                "Turbine x positions": base["Turbine x positions"][i],
                "Turbine y positions": base["Turbine y positions"][i],
            }
        return infos

    def _get_info(self):
        return self._get_infos()

    def reset(self, seed=None, options=None):
        self.agents = list(self.possible_agents)
        obs, infos = super().reset(seed=seed, options=options)
        self.timestep = 0
        return obs, infos

    def step(self, actions):
        all_action = np.array([np.asarray(a, dtype=np.float32).reshape(-1)[0] for a in actions.values()],
                              dtype=np.float32)
        obs, reward, _, truncated, infos = super().step(all_action)
        rewards = {a: reward for a in self.agents}
        truncations = {a: bool(truncated) for a in self.agents}
        terminations = {a: False for a in self.agents}
        if self.compat_double_timestep:  # reference increments twice per step (WindEnvMulti.py:219)
            self.timestep += 1
            st = self.vec.state["timestep"]
            st += 1
        if truncated:
            self.agents = []
        return obs, rewards, terminations, truncations, infos

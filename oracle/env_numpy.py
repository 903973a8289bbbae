"""TEST INFRASTRUCTURE -- numpy restatement of the reference's in-repo env layer (one env, CPU).

This is the oracle that travels to the GPU box (``/root/reference`` does not).  It restates, with citations,

* ``Mes`` / ``turb_mes`` / ``farm_mes``        -- ``WindGym/MesClass.py:23-703``
* ``WindFarmEnv.reset / step`` bookkeeping     -- ``WindGym/Wind_Farm_Env.py:680-802, :920-1034``
* yaw action semantics ``_adjust_yaws``        -- ``Wind_Farm_Env.py:822-864``
* rewards / penalty                            -- ``Wind_Farm_Env.py:804-820, :866-918``
* baseline yaw controllers                     -- ``WindGym/BasicControllers/BasicControllers.py:10-73``
* multi-agent observation split                -- ``WindGym/WindEnvMulti.py:79-103``

over the flow seam in ``oracle.dwm_numpy``.  It is PINNED: ``tests/golden/make_golden.py`` (build container) runs the
unmodified reference files through ``oracle/ref_loader.py`` and writes their outputs to ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` compares this module with them bit for bit, anywhere (the fixtures travel to the GPU box).  Deliberate differences from the reference: wind conditions / initial yaws can be
injected (the batched tests feed identical per-env values to both sides); the measurement-noise RNG is seeded
(the reference's is not); after truncation the object stays usable (Q11).
"""
from collections import deque

import numpy as np

from . import dwm_numpy as dwm


# ------------------------------------------------------------------------------------------------
# MesClass restatement
# ------------------------------------------------------------------------------------------------
def window_bounds(L, N, W, i):
    """Window [lo, hi) of rolling value i for a history holding L samples (MesClass.py:85-116)."""
    if i == 0:
        return max(0, L - W), L
    if i == N - 1 and L >= W:
        return 0, W
    if L < W:
        return 0, L
    spacing = max(1, (L - W) // (N - 1))
    pos = min(i * spacing, L - W)
    return pos, pos + W


class MesO:
    """One scalar history (``Mes``, MesClass.py:23-125)."""

    def __init__(self, current, rolling_mean, history_N, history_length, window_length):
        self.current, self.rolling_mean = bool(current), bool(rolling_mean)
        self.N, self.H, self.W = int(history_N), int(history_length), int(window_length)
        self.q = deque(maxlen=self.H)

    def add(self, v):
        self.q.append(float(v))

    def n_out(self):
        return int(self.current) + int(self.rolling_mean) * self.N

    def get(self):
        out = []
        L = len(self.q)
        if L == 0:
            return np.array(out, dtype=np.float32)
        vals = np.array(self.q, dtype=np.float64)
        if self.current:
            out.append(vals[-1])
        if self.rolling_mean:
            for i in range(self.N):
                lo, hi = window_bounds(L, self.N, self.W, i)
                out.append(np.mean(vals[lo:hi]))
        return np.array(out, dtype=np.float32)


def scale_val(val, lo, hi):
    """``_scale_val`` (MesClass.py:324-326 / WindEnv.py:14-16); float32 in, float32 out."""
    return 2 * (val - lo) / (hi - lo) - 1


class TurbMesO:
    """``turb_mes`` (MesClass.py:128-351)."""

    def __init__(self, ws, wd, yaw, power, ranges, include_TI, power_max):
        self.ws, self.wd, self.yaw, self.power = MesO(*ws), MesO(*wd), MesO(*yaw), MesO(*power)
        (self.ws_min, self.ws_max, self.wd_min, self.wd_max, self.yaw_min, self.yaw_max,
         self.TI_min, self.TI_max) = ranges
        self.include_TI = bool(include_TI)
        self.power_max = power_max

    def n_out(self):
        return self.ws.n_out() + self.wd.n_out() + self.yaw.n_out() + int(self.include_TI) + self.power.n_out()

    def max_hist(self):
        return max(self.ws.H, self.wd.H, self.yaw.H)

    def calc_TI(self, scaled=False):
        u = np.array(self.ws.q, dtype=np.float64)
        U = u.mean()
        TI = np.array([np.std(u - U) / U], dtype=np.float32)
        return scale_val(TI, self.TI_min, self.TI_max) if scaled else TI

    def get(self, scaled):
        ti = self.calc_TI() if self.include_TI else np.array([])
        parts = [self.ws.get(), self.wd.get(), self.yaw.get(), ti, self.power.get()]
        if scaled:
            rng = [(self.ws_min, self.ws_max), (self.wd_min, self.wd_max), (self.yaw_min, self.yaw_max),
                   (self.TI_min, self.TI_max), (0, self.power_max)]
            parts = [scale_val(p, lo, hi) for p, (lo, hi) in zip(parts, rng)]
        return np.concatenate(parts)


class FarmMesO:
    """``farm_mes`` (MesClass.py:354-703)."""

    def __init__(self, n_turb, noise, lv, ws_mes, wd_mes, yaw_mes, power_mes, ranges, power_max, noise_seed=0):
        self.T = n_turb
        self.lv = lv
        self.ranges = ranges
        self.power_max = power_max
        self.noise = noise
        self.rng = np.random.default_rng(noise_seed)
        self.noise_std = {"ws": 0.0, "wd": 2.0, "yaw": 0.0, "power": 0.0}  # MesClass.py:436-439

        def d(m, key, gate):
            return (m[f"{key}_current"] and gate, m[f"{key}_rolling_mean"] and gate, m[f"{key}_history_N"],
                    m[f"{key}_history_length"], m[f"{key}_window_length"])

        self.turb = [
            TurbMesO(d(ws_mes, "ws", lv["turb_ws"]), d(wd_mes, "wd", lv["turb_wd"]), d(yaw_mes, "yaw", True),
                     d(power_mes, "power", lv["turb_power"]), ranges, lv["turb_TI"], power_max)
            for _ in range(n_turb)
        ]
        self.farm = TurbMesO(d(ws_mes, "ws", lv["farm_ws"]), d(wd_mes, "wd", lv["farm_wd"]), d(yaw_mes, "yaw", True),
                             d(power_mes, "power", lv["farm_power"]), ranges, lv["farm_TI"], power_max * n_turb)
        self.farm_n_out = (
            lv["farm_ws"] * (ws_mes["ws_current"] + ws_mes["ws_rolling_mean"] * ws_mes["ws_history_N"])
            + lv["farm_wd"] * (wd_mes["wd_current"] + wd_mes["wd_rolling_mean"] * wd_mes["wd_history_N"])
            + lv["farm_TI"]
            + lv["farm_power"] * (power_mes["power_current"] + power_mes["power_rolling_mean"] * power_mes["power_history_N"])
        )

    def _noise(self, key, n):
        if self.noise == "Normal":
            return self.rng.normal(0.0, self.noise_std[key], size=n)
        return np.zeros(n)

    def max_hist(self):
        return self.turb[0].max_hist()

    def n_out(self):
        return self.turb[0].n_out() * self.T + int(self.farm_n_out)

    def add(self, ws, wd, yaw, power):
        ws = ws + self._noise("ws", self.T)
        wd = wd + self._noise("wd", self.T)
        yaw = yaw + self._noise("yaw", self.T)
        power = power + self._noise("power", self.T)
        for t, tm in enumerate(self.turb):
            tm.ws.add(ws[t]); tm.wd.add(wd[t]); tm.yaw.add(yaw[t]); tm.power.add(power[t])
        self.farm.ws.add(np.mean(ws)); self.farm.wd.add(np.mean(wd)); self.farm.power.add(np.sum(power))

    def get(self, scaled=True):
        """Single-agent observation vector (MesClass.py:679-703); float64 carrier of float32-rounded values."""
        r = self.ranges
        farm = np.array([])
        if self.lv["farm_ws"]:
            v = self.farm.ws.get()
            farm = np.append(farm, scale_val(v, r[0], r[1]) if scaled else v)
        if self.lv["farm_wd"]:
            v = self.farm.wd.get()
            farm = np.append(farm, scale_val(v, r[2], r[3]) if scaled else v)
        if self.lv["farm_TI"]:
            farm = np.append(farm, np.array([tm.calc_TI(scaled) for tm in self.turb]).flatten().mean())
        if self.lv["farm_power"]:
            v = self.farm.power.get()
            farm = np.append(farm, scale_val(v, 0, self.power_max * self.T) if scaled else v)
        turb = np.array([tm.get(scaled) for tm in self.turb]).flatten()
        return np.concatenate([turb, farm])

    def get_multi(self):
        """Per-agent observations of ``WindFarmEnvMulti._get_obs_multi`` (WindEnvMulti.py:79-103)."""
        fb = self.farm.get(True)
        return [np.clip(np.concatenate([tm.get(True), fb]), -1.0, 1.0).astype(np.float32) for tm in self.turb]


# ------------------------------------------------------------------------------------------------
# baseline controllers (BasicControllers.py:10-73)
# ------------------------------------------------------------------------------------------------
def local_yaw_controller(fs, yaw_step=1):
    yaw = fs.windTurbines.yaw
    uvw = fs.windTurbines.rotor_avg_windspeed
    off = np.rad2deg(np.arctan(uvw[:, 1] / uvw[:, 0])) - yaw
    return yaw + np.sign(off) * np.minimum(np.abs(off), yaw_step)


def global_yaw_controller(fs, yaw_step=1):
    yaw = fs.windTurbines.yaw
    return yaw - np.sign(yaw) * np.minimum(np.abs(yaw), yaw_step)


def grid_layout(D, xDist, yDist, nx, ny):
    """Reference layout rule incl. its spacing quirk (Wind_Farm_Env.py:246-252)."""
    x = np.linspace(0, D * xDist * nx, nx)
    y = np.linspace(0, D * yDist * ny, ny)
    xv, yv = np.meshgrid(x, y, indexing="xy")
    return xv.flatten(), yv.flatten()


# ------------------------------------------------------------------------------------------------
# env restatement
# ------------------------------------------------------------------------------------------------
class WindFarmEnvOracle:
    """One-env restatement of ``WindFarmEnv`` (and of ``FarmEval`` with ``eval_mode=True``).

    ``cfg`` is the dict ``yaml.safe_load`` returns for a reference YAML (Wind_Farm_Env.py:349-399).
    """

    def __init__(self, turbine, cfg, n_passthrough=5, TI_min_mes=0.0, TI_max_mes=0.5, turbtype="None",
                 Baseline_comp=False, yaw_init=None, seed=None, dt_sim=1, dt_env=1, yaw_step=1, fill_window=True,
                 eval_mode=False, reset_init=True, noise_seed=0, turb_field=None, added_field=None,
                 induction_control=False, derate_min=0.5):
        if turbtype not in ("None", "MannFixed", "MannGenerate", "MannLoad", "Random"):
            raise ValueError("Invalid turbulence type specified")   # Wind_Farm_Env.py:666-668
        # "Random" (RandomTurbulence(ti, ws, seed), :640-644) runs through the same box interface: turb_field is then a
        # box of independent N(0, 1) cells (the white-noise field frozen on a grid)
        self.turbtype = turbtype
        self.act_var = 2 if induction_control else 1   # extension: [yaw actions | induction actions]
        self.derate_min = derate_min
        self.turb_field = turb_field      # oracle.mann_numpy.MannTurbulenceField for the Mann site types
        self.added_field = added_field    # unit-variance isotropic box of the wake-added turbulence (or None)
        self.turb_offset = (0.0, 0.0, 0.0)
        if turbtype != "None" and turb_field is None:
            raise ValueError("a Mann turbtype needs turb_field (the box the device path was given)")
        self.turbine, self.cfg = turbine, cfg
        self.n_passthrough, self.dt, self.dt_env, self.yaw_step = n_passthrough, dt_sim, dt_env, yaw_step
        if dt_env % dt_sim != 0:
            raise ValueError("dt_env must be a multiple of dt_sim")
        self.S = int(dt_env / dt_sim)
        self.eval_mode, self.noise_seed = eval_mode, noise_seed
        self.yaw_start = 15.0
        self.d_particle = 0.2
        self.maxturbpower = max(turbine.power(np.arange(10, 25, 1)))
        farm, wind = cfg["farm"], cfg["wind"]
        self.yaw_min, self.yaw_max = farm["yaw_min"], farm["yaw_max"]
        self.nx, self.ny = farm["nx"], farm["ny"]
        self.n_turb = self.nx * self.ny
        self.ws_min, self.ws_max = wind["ws_min"], wind["ws_max"]
        self.TI_min, self.TI_max = wind["TI_min"], wind["TI_max"]
        self.wd_min, self.wd_max = wind["wd_min"], wind["wd_max"]
        self.wd_min_mes, self.wd_max_mes = wind["wd_min"], wind["wd_max"]
        self.TI_min_mes, self.TI_max_mes = TI_min_mes, TI_max_mes
        self.action_penalty = cfg["act_pen"]["action_penalty"]
        self.action_penalty_type = cfg["act_pen"]["action_penalty_type"]
        self.Power_scaling = cfg["power_def"]["Power_scaling"]
        self.power_avg = cfg["power_def"]["Power_avg"]
        self.power_reward = cfg["power_def"]["Power_reward"]
        if self.power_reward not in ("Baseline", "Power_avg", "None", "Power_diff"):
            raise ValueError("The Power_reward must be either Baseline, Power_avg, None or Power_diff")
        if self.power_reward == "Power_diff":
            self._power_wSize = self.power_avg // 10
            if self.power_avg < 40:
                raise ValueError("The Power_avg must be larger then 40 for the Power_diff reward.")
        if cfg.get("Track_power"):
            raise NotImplementedError("The Track_power is not implemented yet")
        self.ActionMethod = cfg["ActionMethod"]
        self.BaseController = cfg["BaseController"]
        self.noise = cfg["noise"]
        yi = yaw_init if yaw_init is not None else cfg["yaw_init"]
        self.yaw_init_mode = yi if yi in ("Random", "Defined") else "Zeros"
        self.yaw_initial = [0]
        self.Baseline_comp = bool(self.power_reward == "Baseline" or Baseline_comp)
        if self.Baseline_comp and self.BaseController not in ("Local", "Global"):
            raise ValueError("The BaseController must be either Local or Global... For now")
        self.farm_pow_deq = deque(maxlen=self.power_avg)  # never cleared on reset
        self.base_pow_deq = deque(maxlen=self.power_avg)
        self._new_mes()
        self.hist_max = self.mes.max_hist()
        if fill_window is True:
            self.steps_on_reset = self.hist_max
        elif isinstance(fill_window, int) and not isinstance(fill_window, bool) and fill_window >= 1:
            self.steps_on_reset = min(fill_window, self.hist_max)
        elif fill_window is False:
            self.steps_on_reset = 1
        else:
            raise ValueError("fill_window must be True or a non-negative integer")
        self.D = turbine.diameter()
        self.x_pos, self.y_pos = grid_layout(self.D, farm["xDist"], farm["yDist"], self.nx, self.ny)
        self.obs_var = self.mes.n_out()
        self.np_random = np.random.default_rng(seed)
        self.timestep, self.time_max = 0, 0
        if reset_init:
            self.reset(seed=seed)

    def _new_mes(self):
        c = self.cfg
        ranges = (2.0, 25.0, self.wd_min_mes - 5, self.wd_max_mes + 5, self.yaw_min, self.yaw_max,
                  self.TI_min_mes, self.TI_max_mes)
        self.mes = FarmMesO(self.n_turb, self.noise, c["mes_level"], c["ws_mes"], c["wd_mes"], c["yaw_mes"],
                            c["power_mes"], ranges, self.maxturbpower, noise_seed=self.noise_seed)

    # FarmEval.set_wind_vals / set_yaw_vals (FarmEval.py:63-84)
    def set_wind_vals(self, ws=None, ti=None, wd=None):
        if ws is not None:
            self.ws_min = self.ws_max = ws
        if ti is not None:
            self.TI_min = self.TI_max = ti
        if wd is not None:
            self.wd_min = self.wd_max = wd

    def set_yaw_vals(self, yaw_vals):
        self.yaw_initial = yaw_vals

    def _new_fs(self):
        wts = dwm.PyWakeWindTurbines(self.x_pos, self.y_pos, self.turbine)
        if self.turbtype == "None":  # Wind_Farm_Env.py:661-665
            tf = dwm.RandomTurbulence(ti=0, ws=self.ws)
        else:                        # :612-659: the box is scaled to the episode's TI and wind speed
            tf = self.turb_field
            tf.offset = np.asarray(self.turb_offset, dtype=np.float64)
            tf.scale_TI(TI=self.ti, U=self.ws)
        site = dwm.TurbulenceFieldSite(ws=self.ws, turbulenceField=tf)
        import types
        added = types.SimpleNamespace(field=self.added_field) if self.added_field is not None else None
        return dwm.DWMFlowSimulation(site, wts, wind_direction=self.wd, dt=self.dt, d_particle=self.d_particle,
                                     addedTurbulenceModel=added)

    def _measure(self):
        uvw = self.fs.windTurbines.rotor_avg_windspeed
        self.current_ws = np.linalg.norm(uvw, axis=1)
        self.current_wd = np.rad2deg(np.arctan(uvw[:, 1] / uvw[:, 0])) + self.wd
        self.current_yaw = self.fs.windTurbines.yaw.copy()
        self.current_powers = self.fs.windTurbines.power()

    def _obs(self):
        return np.clip(self.mes.get(True), -1.0, 1.0).astype(np.float32)

    def reset(self, seed=None, wind=None, yaw0=None, turb_offset=None):
        """Wind_Farm_Env.py:680-802.  ``wind=(ws, ti, wd)`` / ``yaw0`` bypass the RNG draws; ``turb_offset`` is the
        env's position inside the shared turbulence box (Mann site types)."""
        if turb_offset is not None:
            self.turb_offset = turb_offset
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        self.timestep = 0
        if wind is None:  # draw order ws -> ti -> wd (:564-568)
            self.ws = self.np_random.uniform(low=self.ws_min, high=self.ws_max)
            self.ti = self.np_random.uniform(low=self.TI_min, high=self.TI_max)
            self.wd = self.np_random.uniform(low=self.wd_min, high=self.wd_max)
        else:
            self.ws, self.ti, self.wd = wind
        self._new_mes()
        self.rated_power = self.turbine.power(self.ws)
        self.fs = self._new_fs()
        if yaw0 is not None:
            y0 = np.asarray(yaw0, dtype=np.float64)
        elif self.yaw_init_mode == "Random":
            y0 = self.np_random.uniform(low=-self.yaw_start, high=self.yaw_start, size=self.n_turb)
        elif self.yaw_init_mode == "Defined":
            yv = np.asarray(self.yaw_initial, dtype=np.float64)
            y0 = yv if yv.size == self.n_turb else np.ones(self.n_turb) * yv[0]
        else:
            y0 = np.zeros(self.n_turb)
        self.fs.windTurbines.yaw = y0
        xr = self.fs.windTurbines.rotor_positions_xyz[0]
        t_inflow = (xr.max() - xr.min()) / self.ws
        t_dev = int(t_inflow * 2)
        self.time_max = 9999999 if self.eval_mode else int(t_inflow * self.n_passthrough)
        self.t_developed = t_dev
        self.fs.run(t_dev)
        for _ in range(self.steps_on_reset):
            acc = [[], [], [], []]
            for _ in range(self.S):
                self.fs.step()
                self._measure()
                for a, v in zip(acc, (self.current_ws, self.current_wd, self.current_yaw, self.current_powers)):
                    a.append(v)
            m = [np.mean(a, axis=0) for a in acc]
            self.mes.add(*m)
            self.farm_pow_deq.append(m[3].sum())
        if self.Baseline_comp:
            self.fs_baseline = self._new_fs()
            self.fs_baseline.windTurbines.yaw = self.fs.windTurbines.yaw
            self.fs_baseline.run(t_dev)
            for _ in range(self.hist_max):  # no controller during the fill
                bp = []
                for _ in range(self.S):
                    self.fs_baseline.step()
                    bp.append(self.fs_baseline.windTurbines.power().sum())
                self.base_pow_deq.append(np.mean(bp, axis=0))
        return self._obs(), self._info()

    def _info(self):
        d = {
            "yaw angles agent": self.current_yaw.copy(),
            "Wind speed Global": self.ws,
            "Wind speed at turbines": self.current_ws,
            "Wind direction Global": self.wd,
            "Wind direction at turbines": self.current_wd,
            "Turbulence intensity": self.ti,
            "Power agent": self.fs.windTurbines.power().sum(),
            "Power pr turbine agent": self.fs.windTurbines.power(),
            "Turbine x positions": self.fs.windTurbines.positions_xyz[0],
            "Turbine y positions": self.fs.windTurbines.positions_xyz[1],
        }
        if self.Baseline_comp:
            d["yaw angles base"] = self.fs_baseline.windTurbines.yaw.copy()
            d["Power baseline"] = self.fs_baseline.windTurbines.power().sum()
            d["Power pr turbine baseline"] = self.fs_baseline.windTurbines.power()
            d["Wind speed at turbines baseline"] = self.fs_baseline.windTurbines.rotor_avg_windspeed[:, 0]
        return d

    def _adjust_yaws(self, action):
        wt = self.fs.windTurbines
        if self.act_var == 2:  # extension: the second half of the action vector sets the induction scale
            ua = np.asarray(action[self.n_turb:], dtype=np.float64)
            wt.derate = np.clip(self.derate_min + 0.5 * (ua + 1.0) * (1.0 - self.derate_min), self.derate_min, 1.0)
            action = action[:self.n_turb]
        if self.ActionMethod == "yaw":
            wt.yaw = np.clip(wt.yaw + action * self.yaw_step, self.yaw_min, self.yaw_max)
        elif self.ActionMethod == "wind":
            new = (action + 1.0) / 2.0 * (self.yaw_max - self.yaw_min) + self.yaw_min
            wt.yaw = np.clip(np.clip(new, wt.yaw - self.yaw_step, wt.yaw + self.yaw_step), self.yaw_min, self.yaw_max)
        elif self.ActionMethod == "absolute":
            raise NotImplementedError("The absolute method is not implemented yet")
        else:
            raise ValueError("The ActionMethod must be yaw, wind or absolute")

    def _power_rew(self):
        if self.power_reward == "Baseline":
            return np.mean(self.farm_pow_deq) / np.mean(self.base_pow_deq) - 1
        if self.power_reward == "Power_avg":
            return np.mean(self.farm_pow_deq) / self.n_turb / self.rated_power
        if self.power_reward == "Power_diff":
            q = list(self.farm_pow_deq)
            return (np.mean(q[self.power_avg - self._power_wSize:self.power_avg]) - np.mean(q[:self._power_wSize])) / self.n_turb
        return 0.0

    def _action_pen(self):
        if self.action_penalty < 0.001:
            return 0
        if self.action_penalty_type == "Change":
            pen = np.mean(np.abs(self.old_yaws - self.fs.windTurbines.yaw))
        elif self.action_penalty_type == "Total":
            pen = np.mean(np.abs(self.fs.windTurbines.yaw)) / self.yaw_max
        return self.action_penalty * pen

    def step(self, action):
        """Wind_Farm_Env.py:920-1034."""
        action = np.asarray(action)
        self.old_yaws = self.fs.windTurbines.yaw.copy()
        self._adjust_yaws(action)
        acc = [[], [], [], []]
        bp = []
        ctrl = local_yaw_controller if self.BaseController == "Local" else global_yaw_controller
        for _ in range(self.S):
            self.fs.step()
            if self.Baseline_comp:
                self.fs_baseline.windTurbines.yaw = ctrl(self.fs_baseline, self.yaw_step)
                self.fs_baseline.step()
                bp.append(self.fs_baseline.windTurbines.power().sum())
            self._measure()
            for a, v in zip(acc, (self.current_ws, self.current_wd, self.current_yaw, self.current_powers)):
                a.append(v)
        m = [np.mean(a, axis=0) for a in acc]
        self.mes.add(*m)
        self.farm_pow_deq.append(m[3].sum())
        if self.Baseline_comp:
            self.base_pow_deq.append(np.mean(bp, axis=0))
        if np.any(np.isnan(self.farm_pow_deq)):
            raise Exception("NaN Power")
        obs, info = self._obs(), self._info()
        reward = self._power_rew() * self.Power_scaling - self._action_pen()
        truncated = bool(self.timestep >= self.time_max)
        self.timestep += 1
        return obs, reward, False, truncated, info

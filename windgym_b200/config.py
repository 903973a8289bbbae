"""Host-side configuration of the batched env: the reference's YAML schema and constructor arguments.

Mirrors ``WindFarmEnv.__init__`` / ``load_config`` (reference ``WindGym/Wind_Farm_Env.py:50-261, :349-399``):
same keys, same defaults, same error behaviour (``ValueError`` / ``NotImplementedError``).
"""
import math
from dataclasses import dataclass, field

import numpy as np
import yaml

K_HILL = 0.4
CT_MAX = 0.96
MARGIN_D = 2.0

_ACTION = {"yaw": 0, "wind": 1, "absolute": 2}
_REWARD = {"None": 0, "Baseline": 1, "Power_avg": 2, "Power_diff": 3}
_PENALTY = {"Change": 0, "Total": 1}
_CONTROLLER = {"Local": 0, "Global": 1}


def load_yaml(path):
    with open(path, "r") as fh:
        return yaml.safe_load(fh)


def grid_layout(D, xDist, yDist, nx, ny):
    """Reference layout rule including its spacing quirk (Wind_Farm_Env.py:246-252)."""
    x = np.linspace(0, D * xDist * nx, nx)
    y = np.linspace(0, D * yDist * ny, ny)
    xv, yv = np.meshgrid(x, y, indexing="xy")
    return xv.flatten().astype(np.float64), yv.flatten().astype(np.float64)


def rotate_layout(x, y, wd):
    """Layout frame -> wind-aligned frame about the centroid (theta = 270 deg - wd)."""
    th = np.deg2rad(270.0 - np.asarray(wd, dtype=np.float64))[..., None]
    dx, dy = x - x.mean(), y - y.mean()
    return dx * np.cos(th) + dy * np.sin(th), -dx * np.sin(th) + dy * np.cos(th)


@dataclass
class EnvConfig:
    """Everything that is fixed for the lifetime of an env object."""
    cfg: dict
    turbine: object
    n_passthrough: float = 5
    TI_min_mes: float = 0.0
    TI_max_mes: float = 0.5
    turbtype: str = "None"
    Baseline_comp: bool = False
    yaw_init: str = None
    dt_sim: float = 1
    dt_env: float = 1
    yaw_step: float = 1
    fill_window: object = True
    eval_mode: bool = False
    multi_agent: bool = False
    noise_seed: int = 0
    induction_control: bool = False   # extension: one derating action per turbine after the yaw actions
    derate_min: float = 0.5
    derived: dict = field(default_factory=dict)

    def __post_init__(self):
        c, t = self.cfg, self.turbine
        if self.dt_env % self.dt_sim != 0:
            raise ValueError("dt_env must be a multiple of dt_sim")
        self.S = int(self.dt_env / self.dt_sim)
        self.act_var = 2 if self.induction_control else 1
        if not (0.0 < self.derate_min <= 1.0):
            raise ValueError("derate_min must be in (0, 1]")
        if self.turbtype not in ("None", "MannLoad", "MannGenerate", "MannFixed", "Random"):
            raise ValueError("Invalid turbulence type specified")  # Wind_Farm_Env.py:666-668
        self.yaw_start = 15.0
        self.d_particle = 0.2
        self.maxturbpower = float(max(t.power(np.arange(10, 25, 1))))
        farm, wind = c["farm"], c["wind"]
        self.yaw_min, self.yaw_max = farm["yaw_min"], farm["yaw_max"]
        self.nx, self.ny = farm["nx"], farm["ny"]
        self.n_turb = self.nx * self.ny
        self.ws_min, self.ws_max = wind["ws_min"], wind["ws_max"]
        self.TI_min, self.TI_max = wind["TI_min"], wind["TI_max"]
        self.wd_min, self.wd_max = wind["wd_min"], wind["wd_max"]
        self.wd_min_mes, self.wd_max_mes = wind["wd_min"], wind["wd_max"]
        self.action_penalty = c["act_pen"]["action_penalty"]
        self.action_penalty_type = c["act_pen"]["action_penalty_type"]
        self.Power_scaling = c["power_def"]["Power_scaling"]
        self.power_avg = c["power_def"]["Power_avg"]
        self.power_reward = c["power_def"]["Power_reward"]
        if c.get("Track_power"):
            raise NotImplementedError("The Track_power is not implemented yet")
        if self.power_reward not in _REWARD:
            raise ValueError("The Power_reward must be either Baseline, Power_avg, None or Power_diff")
        if self.power_reward == "Power_diff" and self.power_avg < 40:
            raise ValueError("The Power_avg must be larger then 40 for the Power_diff reward. "
                             "Also it should probably be way larger my guy")
        self.ActionMethod = c["ActionMethod"]
        if self.ActionMethod == "absolute":
            raise NotImplementedError("The absolute method is not implemented yet")
        if self.ActionMethod not in _ACTION:
            raise ValueError("The ActionMethod must be yaw, wind or absolute")
        self.BaseController = c["BaseController"]
        self.noise = c["noise"]
        yi = self.yaw_init if self.yaw_init is not None else c["yaw_init"]
        self.yaw_init_mode = yi if yi in ("Random", "Defined") else "Zeros"
        self.Baseline_comp = bool(self.power_reward == "Baseline" or self.Baseline_comp)
        if self.Baseline_comp and self.BaseController not in _CONTROLLER:
            raise ValueError("The BaseController must be either Local or Global... For now")
        self.mes_level, self.ws_mes, self.wd_mes = c["mes_level"], c["ws_mes"], c["wd_mes"]
        self.yaw_mes, self.power_mes = c["yaw_mes"], c["power_mes"]
        self.hist_max = max(self.ws_mes["ws_history_length"], self.wd_mes["wd_history_length"],
                            self.yaw_mes["yaw_history_length"])
        fw = self.fill_window
        if fw is True:
            self.steps_on_reset = self.hist_max
        elif isinstance(fw, int) and not isinstance(fw, bool) and fw >= 1:
            self.steps_on_reset = min(fw, self.hist_max)
        elif fw is False:
            self.steps_on_reset = 1
        else:
            raise ValueError("fill_window must be True or a non-negative integer")
        self.D = float(t.diameter())
        self.hub_height = float(t.hub_height())
        self.x_pos, self.y_pos = grid_layout(self.D, farm["xDist"], farm["yDist"], self.nx, self.ny)
        # wake-chain capacity that cannot overflow for ANY wind direction / speed (FarmEval may pin either):
        # farm diagonal + margin, at the tightest particle spacing d_particle*D*(1 - k_hill*2a_max)
        ct_tab = np.asarray(getattr(t, "ct_table", [0.9]), dtype=np.float64)
        a_max = 0.5 * (1.0 - math.sqrt(1.0 - min(float(ct_tab.max()), CT_MAX)))
        self.f_min = 1.0 - K_HILL * 2.0 * a_max
        diag = math.hypot(self.x_pos.max() - self.x_pos.min(), self.y_pos.max() - self.y_pos.min())
        n = int(math.ceil((diag + MARGIN_D * self.D) / (self.d_particle * self.D * self.f_min))) + 3
        self.p_cap = (n + 7) // 8 * 8

    # per-env quantities the reference computes at reset (Wind_Farm_Env.py:723-732), in fp64
    def reset_integers(self, ws, wd):
        ws = np.asarray(ws, dtype=np.float64)
        ws = np.where(ws > 0, ws, 1.0)   # slots that were never reset (spare pool) carry no conditions yet
        xr, _ = rotate_layout(self.x_pos, self.y_pos, wd)
        dist = xr.max(axis=-1) - xr.min(axis=-1)
        t_inflow = dist / ws
        t_dev = (t_inflow * 2).astype(np.int64)
        n_spin = np.rint(t_dev / self.dt_sim).astype(np.int32)
        time_max = np.full(ws.shape, 9999999, dtype=np.int32) if self.eval_mode else \
            (t_inflow * self.n_passthrough).astype(np.int32)
        k_emit = np.maximum(1, np.ceil(self.d_particle * self.D / (ws * self.dt_sim) - 1e-9)).astype(np.int32)
        return n_spin, time_max, k_emit

    def codes(self):
        return dict(action=_ACTION[self.ActionMethod], reward=_REWARD[self.power_reward],
                    penalty=_PENALTY.get(self.action_penalty_type, 0),
                    controller=_CONTROLLER.get(self.BaseController, 0))

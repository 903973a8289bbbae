"""Shared test helpers: reference-schema configs and oracle rollouts (oracle = checker only)."""
import copy

import numpy as np

from oracle.env_numpy import WindFarmEnvOracle
from oracle.v80 import V80 as OracleV80

# Env1.yaml semantics (reference examples/EnvConfigs/Env1.yaml), written out so that the GPU box needs no
# reference checkout.
ENV1 = {
    "yaw_init": "Random", "noise": "None", "BaseController": "Local", "ActionMethod": "wind", "Track_power": False,
    "farm": {"yaw_min": -45, "yaw_max": 45, "xDist": 4, "yDist": 4, "nx": 2, "ny": 2},
    "wind": {"ws_min": 7, "ws_max": 15, "TI_min": 0.02, "TI_max": 0.15, "wd_min": 255, "wd_max": 285},
    "act_pen": {"action_penalty": 0.0, "action_penalty_type": "Change"},
    "power_def": {"Power_reward": "Baseline", "Power_avg": 10, "Power_scaling": 1.0},
    "mes_level": {"turb_ws": True, "turb_wd": False, "turb_TI": False, "turb_power": False,
                  "farm_ws": False, "farm_wd": False, "farm_TI": False, "farm_power": False},
    "ws_mes": {"ws_current": False, "ws_rolling_mean": True, "ws_history_N": 1, "ws_history_length": 25,
               "ws_window_length": 25},
    "wd_mes": {"wd_current": False, "wd_rolling_mean": False, "wd_history_N": 1, "wd_history_length": 20,
               "wd_window_length": 20},
    "yaw_mes": {"yaw_current": False, "yaw_rolling_mean": True, "yaw_history_N": 1, "yaw_history_length": 10,
                "yaw_window_length": 10},
    "power_mes": {"power_current": False, "power_rolling_mean": False, "power_history_N": 1,
                  "power_history_length": 10, "power_window_length": 10},
}


def small_config(nx=2, ny=2, reward="Power_avg", action="wind", **over):
    """Env1-style config with a different grid / reward; ``over`` patches nested keys as 'section.key'=value."""
    c = copy.deepcopy(ENV1)
    c["farm"]["nx"], c["farm"]["ny"] = nx, ny
    c["power_def"]["Power_reward"] = reward
    c["ActionMethod"] = action
    for k, v in over.items():
        if "." in k:
            a, b = k.split(".")
            c[a][b] = v
        else:
            c[k] = v
    return c


def rich_config(nx=2, ny=2, **over):
    """Every measurement channel switched on, several windows -- exercises the whole MesClass surface."""
    c = small_config(nx, ny, **over)
    c["mes_level"] = {k: True for k in c["mes_level"]}
    c["ws_mes"] = {"ws_current": True, "ws_rolling_mean": True, "ws_history_N": 3, "ws_history_length": 12,
                   "ws_window_length": 4}
    c["wd_mes"] = {"wd_current": True, "wd_rolling_mean": True, "wd_history_N": 2, "wd_history_length": 9,
                   "wd_window_length": 3}
    c["yaw_mes"] = {"yaw_current": False, "yaw_rolling_mean": True, "yaw_history_N": 4, "yaw_history_length": 10,
                    "yaw_window_length": 1}
    c["power_mes"] = {"power_current": True, "power_rolling_mean": True, "power_history_N": 2,
                      "power_history_length": 7, "power_window_length": 7}
    return c


def oracle_rollout(cfg, ws, ti, wd, yaw0, acts, multi=False, reset_kw=None, **kw):
    """Run the CPU oracle env for every batch entry.  acts: [steps, B, T].  Returns per-env stacked arrays."""
    B = len(ws)
    out = {k: [] for k in ("obs0", "obs", "reward", "power", "yaw", "ws_turb", "trunc", "power_base", "yaw_base",
                           "time_max", "t_developed", "u")}
    for b in range(B):
        env = WindFarmEnvOracle(OracleV80(), cfg, reset_init=False, **kw)
        o0, _ = env.reset(wind=(ws[b], ti[b], wd[b]), yaw0=yaw0[b], **(reset_kw or {}))
        rec = {k: [] for k in ("obs", "reward", "power", "yaw", "ws_turb", "trunc", "power_base", "yaw_base", "u")}
        if multi:
            o0 = np.stack(env.mes.get_multi())
        for a in acts[:, b]:
            o, r, _, tr, info = env.step(a)
            if multi:
                o = np.stack(env.mes.get_multi())
            rec["obs"].append(o); rec["reward"].append(r); rec["trunc"].append(tr)
            rec["power"].append(info["Power pr turbine agent"]); rec["yaw"].append(info["yaw angles agent"])
            rec["ws_turb"].append(info["Wind speed at turbines"])
            rec["u"].append(np.array(env.fs.windTurbines.rotor_avg_windspeed[:, 0]))
            if env.Baseline_comp:
                rec["power_base"].append(info["Power pr turbine baseline"]); rec["yaw_base"].append(info["yaw angles base"])
        out["obs0"].append(o0)
        for k, v in rec.items():
            out[k].append(np.array(v))
        out["time_max"].append(env.time_max); out["t_developed"].append(env.t_developed)
    return {k: np.array(v) for k, v in out.items()}

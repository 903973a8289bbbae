"""Generate tests/golden/*.npz by running the UNMODIFIED reference env layer (build container only).

    python tests/golden/make_golden.py

What runs: ``/root/reference/WindGym/{Wind_Farm_Env,MesClass,WindEnv,FarmEval,BasicControllers}.py`` exactly as
shipped, imported through ``oracle/ref_loader.py`` (import-only shims for gymnasium / matplotlib / ...), over the
fp64 restatement of the un-vendored ``dynamiks`` seam (``oracle/dwm_numpy.py``, re-exported by
``oracle/shims/dynamiks``).  So these vectors PIN everything the reference repo itself contains on the hot path
(MesClass windows/scaling, yaw action semantics, substep means, rewards, truncation, reset bookkeeping, RNG draw
order, baseline controllers); the flow numbers inside them come from the restated solver (parity unpinned,
see oracle/dwm_numpy.py).

Files (all small):
  mes_golden.npz   farm_mes.add_measurements / get_measurements(scaled=True) on seeded random histories
  env_golden.npz   WindFarmEnv / FarmEval reset + step trajectories on the shipped YAMLs and variants
  multi_golden.npz WindFarmEnvMulti (WindEnvMulti.py) reset + step: per-agent observation split, shared reward,
                   half-length truncation
  cfg1_golden.npz  BASELINE.json configs[0]: shipped 2turb.yaml, seed 1, 200 steps (zero action, then +0.5)
  ppo_golden.npz   the shipped agent examples/PPO_2975000.zip: actor weights + deterministic actions on fixed observations
"""
import json
import os
import sys
import tempfile

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402
from oracle.v80 import V80  # noqa: E402
from tests.helpers import ENV1, rich_config, small_config  # noqa: E402


def mes_cases():
    """(name, T, farm_mes kwargs, n_push)"""
    env1 = dict(noise="None", turb_ws=True, turb_wd=False, turb_TI=False, turb_power=False, farm_ws=False,
                farm_wd=False, farm_TI=False, farm_power=False,
                ws_current=False, ws_rolling_mean=True, ws_history_N=1, ws_history_length=25, ws_window_length=25,
                wd_current=False, wd_rolling_mean=False, wd_history_N=1, wd_history_length=20, wd_window_length=20,
                yaw_current=False, yaw_rolling_mean=True, yaw_history_N=1, yaw_history_length=10, yaw_window_length=10,
                power_current=False, power_rolling_mean=False, power_history_N=1, power_history_length=10,
                power_window_length=10,
                ws_min=2.0, ws_max=25.0, wd_min=250.0, wd_max=290.0, yaw_min=-45, yaw_max=45, TI_min=0.0, TI_max=0.5,
                power_max=2000000.0)
    rich = dict(env1, turb_wd=True, turb_TI=True, turb_power=True, farm_ws=True, farm_wd=True, farm_TI=True,
                farm_power=True, ws_current=True, ws_history_N=3, ws_history_length=12, ws_window_length=4,
                wd_current=True, wd_rolling_mean=True, wd_history_N=2, wd_history_length=9, wd_window_length=3,
                yaw_history_N=4, yaw_history_length=10, yaw_window_length=1,
                power_current=True, power_rolling_mean=True, power_history_N=2, power_history_length=7,
                power_window_length=7)
    hist100 = dict(env1, ws_history_N=100, ws_history_length=100, ws_window_length=1,
                   yaw_rolling_mean=True, yaw_history_N=100, yaw_history_length=100, yaw_window_length=1)
    spaced = dict(env1, ws_history_N=4, ws_history_length=10, ws_window_length=1, ws_current=True,
                  yaw_history_N=3, yaw_history_length=25, yaw_window_length=5, farm_ws=True, farm_power=True,
                  power_rolling_mean=True, power_history_N=5, power_history_length=30, power_window_length=3)
    return [("env1_T4", 4, env1, 40), ("rich_T3", 3, rich, 30), ("hist100_T2", 2, hist100, 130),
            ("spaced_T16", 16, spaced, 45)]


def make_mes(ns, out):
    meta = {}
    for name, T, kw, n_push in mes_cases():
        fm = ns.MesClass.farm_mes(T, **kw)
        rng = np.random.default_rng(abs(hash(name)) % 2**31 if False else sum(map(ord, name)))
        ws = rng.uniform(3.0, 24.0, (n_push, T))
        wd = rng.uniform(252.0, 288.0, (n_push, T))
        yaw = rng.uniform(-44.0, 44.0, (n_push, T))
        pw = rng.uniform(0.0, 2.0e6, (n_push, T))
        obs = []
        for k in range(n_push):
            fm.add_measurements(ws[k].copy(), wd[k].copy(), yaw[k].copy(), pw[k].copy())
            o = np.clip(fm.get_measurements(scaled=True), -1.0, 1.0).astype(np.float32)  # Wind_Farm_Env.py:513-520
            obs.append(o)
        out[f"{name}/ws"], out[f"{name}/wd"], out[f"{name}/yaw"], out[f"{name}/power"] = ws, wd, yaw, pw
        out[f"{name}/obs"] = np.array(obs)
        meta[name] = dict(T=T, kwargs=kw, n_push=n_push, observed_variables=int(fm.observed_variables()))
    out["meta"] = np.array(json.dumps(meta))


def write_yaml(cfg):
    f = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    yaml.safe_dump(cfg, f)
    f.close()
    return f.name


def env_cases(ns):
    ex = ns.examples
    with open(os.path.join(ex, "2turb.yaml")) as fh:
        two = yaml.safe_load(fh)
    with open(os.path.join(ex, "Env1.yaml")) as fh:
        env1 = yaml.safe_load(fh)
    assert env1 == ENV1, "tests/helpers.ENV1 must equal the shipped Env1.yaml"
    cases = [
        dict(name="env1_seed1", cfg=env1, kw=dict(seed=1), steps=14),
        dict(name="2turb_seed1", cfg=two, kw=dict(seed=1), steps=8),
        dict(name="power_avg_yaw_dt2_total", kw=dict(seed=7, dt_env=2, dt_sim=1, yaw_step=2), steps=8,
             cfg=small_config(2, 2, reward="Power_avg", action="yaw",
                              **{"act_pen.action_penalty": 0.1, "act_pen.action_penalty_type": "Total"})),
        dict(name="rich_3x1_global_base", kw=dict(seed=3, Baseline_comp=True), steps=8,
             cfg=rich_config(3, 1, reward="Power_avg", BaseController="Global")),
        dict(name="power_diff_change", kw=dict(seed=11), steps=6,
             cfg=small_config(2, 1, reward="Power_diff", action="yaw",
                              **{"power_def.Power_avg": 40, "act_pen.action_penalty": 0.02,
                                 "ws_mes.ws_history_length": 45})),
        dict(name="truncation_short", kw=dict(seed=5, n_passthrough=0.05), steps=8,
             cfg=small_config(2, 1, reward="Power_avg", action="wind")),
    ]
    return cases


def make_env(ns, out):
    meta = {}
    for c in env_cases(ns):
        path = write_yaml(c["cfg"])
        env = ns.WindFarmEnv(V80(), yaml_path=path, turbtype="None", **c["kw"])  # ctor resets with seed (:259-261)
        obs0, info0 = env.reset(seed=c["kw"]["seed"])
        T = env.n_turb
        yaw0 = np.array(env.fs.windTurbines.yaw, dtype=np.float64).copy()
        rng = np.random.default_rng(100 + c["kw"]["seed"])
        acts = rng.uniform(-1, 1, (c["steps"], T)).astype(np.float32)
        rec = {k: [] for k in ("obs", "reward", "trunc", "power", "yaw", "ws_turb", "wd_turb", "power_base",
                               "yaw_base", "fs_time")}
        for a in acts:
            o, r, term, tr, info = env.step(a)
            assert term is False
            rec["obs"].append(o); rec["reward"].append(r); rec["trunc"].append(tr)
            rec["power"].append(np.array(info["Power pr turbine agent"]))
            rec["yaw"].append(np.array(info["yaw angles agent"]).copy())
            rec["ws_turb"].append(np.array(info["Wind speed at turbines"]))
            rec["wd_turb"].append(np.array(info["Wind direction at turbines"]))
            rec["fs_time"].append(info.get("time_array", [np.nan])[-1] if False else np.nan)
            if env.Baseline_comp:
                rec["power_base"].append(np.array(info["Power pr turbine baseline"]))
                rec["yaw_base"].append(np.array(info["yaw angles base"]).copy())
            if tr:
                break
        n = len(rec["obs"])
        pre = c["name"]
        out[f"{pre}/acts"] = acts[:n]
        out[f"{pre}/obs0"] = obs0
        out[f"{pre}/yaw0"] = yaw0
        for k, v in rec.items():
            if k != "fs_time":
                out[f"{pre}/{k}"] = np.array(v)
        meta[pre] = dict(cfg=c["cfg"], kw=c["kw"], ws=float(env.ws) if n and not rec["trunc"][-1] else None,
                         steps=n, n_turb=T)
        # wind conditions / integers are read before a possible truncation teardown: re-create to read them safely
        env2 = ns.WindFarmEnv(V80(), yaml_path=path, turbtype="None", **c["kw"])
        env2.reset(seed=c["kw"]["seed"])
        meta[pre].update(ws=float(env2.ws), ti=float(env2.ti), wd=float(env2.wd), time_max=int(env2.time_max),
                         fs_time_after_reset=float(env2.fs.time), obs_var=int(env2.obs_var),
                         rated_power=float(env2.rated_power))
        os.unlink(path)

    # FarmEval + ConstantAgent known answer (reference tests/test_basics.py:415-459 scenario)
    path = write_yaml(ENV1)
    fe = ns.FarmEvalCls(V80(), yaml_path=path, turbtype="None", yaw_init="Zeros", seed=2, reset_init=True)
    fe.set_wind_vals(ws=10, ti=0.07, wd=270)
    obs0, _ = fe.reset()
    agent = ns.ConstantAgent(yaw_angles=[-10, 20, 0, 0])
    yaws, powers = [np.array(fe.fs.windTurbines.yaw).copy()], []
    obs = [obs0]
    for _ in range(25):
        a, _ = agent.predict(obs[-1])
        o, r, te, tr, info = fe.step(a)
        obs.append(o)
        yaws.append(np.array(fe.fs.windTurbines.yaw).copy())
        powers.append(np.array(fe.fs.windTurbines.power()))
    out["farmeval/action"] = np.asarray(agent.predict()[0], dtype=np.float64)
    out["farmeval/yaw"] = np.array(yaws)
    out["farmeval/power"] = np.array(powers)
    out["farmeval/obs"] = np.array(obs)
    meta["farmeval"] = dict(fs_time_after_25=float(fe.fs.time), time_max=int(fe.time_max), ws=10, ti=0.07, wd=270)
    os.unlink(path)
    out["meta"] = np.array(json.dumps(meta))


def multi_cases():
    """WindFarmEnvMulti cases: BASELINE.json cfg 5's shape (4x2 farm, Env1-style channels: 2 observations per agent)
    and one with every turbine- and farm-level channel on (exercises the farm block: ghost farm-yaw slots that never
    receive data, farm TI = calc_TI of the farm-mean ws ring -- SURVEY.md Q9-ii, Q10), one that runs into the
    truncation (the reference counts ``timestep`` twice per step, WindEnvMulti.py:219 + Wind_Farm_Env.py:1027:
    half-length episodes -- and raises in the truncating step, see make_multi)."""
    return [
        dict(name="multi_4x2_env1", cfg=small_config(4, 2, reward="Power_avg", action="yaw"), kw=dict(seed=5), steps=8),
        dict(name="multi_rich_2x2", cfg=rich_config(2, 2, reward="Power_avg", action="wind"), kw=dict(seed=9), steps=12),
        dict(name="multi_truncation", cfg=small_config(2, 1, reward="Power_avg", action="wind"),
             kw=dict(seed=5, n_passthrough=0.2), steps=12),
    ]


def make_multi(ns, out):
    """The UNMODIFIED ``WindFarmEnvMulti`` (WindEnvMulti.py:17-249).  Its constructor resets before
    ``possible_agents`` exists; the pettingzoo shim's class attribute ``possible_agents = []`` lets that
    first reset pass with no agents, the explicit ``reset(seed)`` below is the real one."""
    Multi = ns.WindEnvMulti.WindFarmEnvMulti
    meta = {}
    for c in multi_cases():
        path = write_yaml(c["cfg"])
        env = Multi(V80(), yaml_path=path, turbtype="None", **c["kw"])
        obs0, infos0 = env.reset(seed=c["kw"]["seed"])
        agents = list(env.possible_agents)
        T = env.n_turb
        yaw0 = np.array(env.fs.windTurbines.yaw, dtype=np.float64).copy()
        mt = dict(ws=float(env.ws), ti=float(env.ti), wd=float(env.wd), time_max=int(env.time_max),
                  fs_time_after_reset=float(env.fs.time), declared_obs_var=int(env.obs_var), n_turb=T,
                  obs_len=int(len(obs0[agents[0]])), cfg=c["cfg"], kw=c["kw"])
        rng = np.random.default_rng(200 + c["kw"]["seed"])
        acts = rng.uniform(-1, 1, (c["steps"], T)).astype(np.float32)
        rec = {k: [] for k in ("obs", "reward", "trunc", "power", "yaw")}
        for a in acts:
            try:
                o, r, te, tr, info = env.step({ag: a[i:i + 1] for i, ag in enumerate(agents)})
            except AttributeError:
                # The reference cannot return from the truncating step: WindFarmEnv.step tears the episode down
                # (deletes farm_measurements, Wind_Farm_Env.py:1003-1023) and WindFarmEnvMulti.step then
                # reads them (WindEnvMulti.py:201).  What it pins is WHEN: the index of the truncating step, half the
                # single-agent episode length because ``timestep`` is counted twice per step (Q9-iii).
                mt["truncating_step"] = len(rec["obs"])
                break
            assert not any(te.values()) and len(set(r.values())) == 1 and len(set(tr.values())) == 1
            rec["obs"].append(np.stack([o[ag] for ag in agents]))
            rec["reward"].append(r[agents[0]])
            rec["trunc"].append(tr[agents[0]])
            if tr[agents[0]]:
                assert env.agents == []          # the episode is over: infos is empty (no agents)
                break
            rec["power"].append(np.array([info[ag]["Power turbine agent"] for ag in agents]))
            rec["yaw"].append(np.array([info[ag]["yaw angles agent"] for ag in agents]))
        pre = c["name"]
        n = len(rec["obs"])
        out[f"{pre}/acts"], out[f"{pre}/yaw0"] = acts[:n], yaw0
        out[f"{pre}/obs0"] = np.stack([obs0[ag] for ag in agents])
        for k, v in rec.items():
            out[f"{pre}/{k}"] = np.array(v)
        mt["steps"] = n
        meta[pre] = mt
        os.unlink(path)
    out["meta"] = np.array(json.dumps(meta))


def make_cfg1(ns, out):
    """BASELINE.json configs[0] exactly as stated: the shipped 2turb.yaml, seed 1, 200 steps on the CPU through the
    reference's WindFarmEnv -- zero action for 100 steps, then +0.5."""
    with open(os.path.join(ns.examples, "2turb.yaml")) as fh:
        two = yaml.safe_load(fh)
    path = write_yaml(two)
    env = ns.WindFarmEnv(V80(), yaml_path=path, turbtype="None", seed=1)
    obs0, _ = env.reset(seed=1)
    T = env.n_turb
    acts = np.zeros((200, T), dtype=np.float32)
    acts[100:] = 0.5
    rec = {k: [] for k in ("obs", "reward", "trunc", "power", "yaw", "power_base", "yaw_base", "ws_turb")}
    out["cfg1/yaw0"] = np.array(env.fs.windTurbines.yaw, dtype=np.float64).copy()
    for a in acts:
        o, r, te, tr, info = env.step(a)
        assert te is False and not tr
        rec["obs"].append(o); rec["reward"].append(r); rec["trunc"].append(tr)
        rec["power"].append(np.array(info["Power pr turbine agent"]))
        rec["yaw"].append(np.array(info["yaw angles agent"]).copy())
        rec["ws_turb"].append(np.array(info["Wind speed at turbines"]))
        if env.Baseline_comp:
            rec["power_base"].append(np.array(info["Power pr turbine baseline"]))
            rec["yaw_base"].append(np.array(info["yaw angles base"]).copy())
    out["cfg1/acts"], out["cfg1/obs0"] = acts, obs0
    for k, v in rec.items():
        out[f"cfg1/{k}"] = np.array(v)
    out["meta"] = np.array(json.dumps({"cfg1": dict(
        cfg=two, ws=float(env.ws), ti=float(env.ti), wd=float(env.wd), time_max=int(env.time_max),
        obs_var=int(env.obs_var), n_turb=T, steps=200, fs_time_end=float(env.fs.time),
        Baseline_comp=bool(env.Baseline_comp))}))
    os.unlink(path)


def make_ppo(ns, out):
    """The shipped agent ``/root/reference/examples/PPO_2975000.zip`` (SB3 2.3.2 MlpPolicy, 8 -> 64 -> 64 -> 4, tanh):
    its actor weights and its deterministic actions on fixed observations, evaluated HERE with plain numpy from the raw
    state dict (mean = action_net(tanh(W2 tanh(W1 o + b1) + b2)); predict(deterministic=True) = clip(mean, -1, 1)) --
    independent of windgym_b200.agents.SB3MlpPolicy, which the tests compare against it."""
    import io
    import zipfile

    import torch
    path = os.path.join(os.path.dirname(ns.examples), "PPO_2975000.zip")
    with zipfile.ZipFile(path) as z:
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
    w = {k: v.double().numpy() for k, v in sd.items()}
    rng = np.random.default_rng(2975000)
    obs = np.concatenate([rng.uniform(-1, 1, (64, 8)), np.zeros((1, 8)), np.ones((1, 8)), -np.ones((1, 8))]).astype(np.float32)
    h = obs.astype(np.float64)
    for i in (0, 2):
        h = np.tanh(h @ w[f"mlp_extractor.policy_net.{i}.weight"].T + w[f"mlp_extractor.policy_net.{i}.bias"])
    mean = h @ w["action_net.weight"].T + w["action_net.bias"]
    out["obs"], out["actions"] = obs, np.clip(mean, -1.0, 1.0)
    for k in sd:
        if k.startswith("mlp_extractor.policy_net.") or k.startswith("action_net.") or k == "log_std":
            out["sd/" + k] = sd[k].numpy()
    out["meta"] = np.array(json.dumps({"source": "examples/PPO_2975000.zip", "obs_dim": 8, "n_actions": 4,
                                       "hidden": [64, 64], "keys": sorted(sd.keys())}))


def main():
    ns = load_reference()
    only = sys.argv[1:]   # e.g. `make_golden.py multi cfg1` regenerates just those files
    jobs = {"mes": (make_mes, "mes_golden.npz"), "env": (make_env, "env_golden.npz"),
            "multi": (make_multi, "multi_golden.npz"), "cfg1": (make_cfg1, "cfg1_golden.npz"),
            "ppo": (make_ppo, "ppo_golden.npz")}
    for key, (fn, fname) in jobs.items():
        if only and key not in only:
            continue
        data = {}
        fn(ns, data)
        np.savez_compressed(os.path.join(HERE, fname), **data)
        print(fname, os.path.getsize(os.path.join(HERE, fname)), "bytes")


if __name__ == "__main__":
    main()

"""CPU: the oracle (oracle/env_numpy.py over oracle/dwm_numpy.py) against the golden vectors produced by the
UNMODIFIED reference env layer (tests/golden/make_golden.py) and against the reference's own known-answer tests
(/root/reference/tests/test_MesClass.py, restated for the oracle classes).  Runs anywhere (no reference needed)."""
import json
import os

import numpy as np
import pytest

from oracle.env_numpy import FarmMesO, MesO, TurbMesO, WindFarmEnvOracle, scale_val, window_bounds
from oracle.v80 import V80

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    z = np.load(os.path.join(GOLD, name), allow_pickle=False)
    return z, json.loads(str(z["meta"]))


def farm_mes_from_kwargs(T, kw):
    """Build the oracle's FarmMesO from the reference's farm_mes keyword arguments (MesClass.py:360-401)."""
    lv = {k: kw[k] for k in ("turb_ws", "turb_wd", "turb_TI", "turb_power", "farm_ws", "farm_wd", "farm_TI",
                             "farm_power")}
    ch = {c: {k: v for k, v in kw.items() if k.startswith(c + "_") and k not in (c + "_min", c + "_max")}
          for c in ("ws", "wd", "yaw", "power")}
    ranges = (kw["ws_min"], kw["ws_max"], kw["wd_min"], kw["wd_max"], kw["yaw_min"], kw["yaw_max"], kw["TI_min"],
              kw["TI_max"])
    return FarmMesO(T, kw["noise"], lv, ch["ws"], ch["wd"], ch["yaw"], ch["power"], ranges, kw["power_max"])


# ------------------------------------------------------------------------------------------------ MesClass golden
@pytest.mark.parametrize("case", ["env1_T4", "rich_T3", "hist100_T2", "spaced_T16"])
def test_mes_golden(case):
    z, meta = _load("mes_golden.npz")
    m = meta[case]
    fm = farm_mes_from_kwargs(m["T"], m["kwargs"])
    assert fm.n_out() == m["observed_variables"]
    for k in range(m["n_push"]):
        fm.add(z[f"{case}/ws"][k], z[f"{case}/wd"][k], z[f"{case}/yaw"][k], z[f"{case}/power"][k])
        obs = np.clip(fm.get(True), -1.0, 1.0).astype(np.float32)
        ref = z[f"{case}/obs"][k]
        assert obs.shape == ref.shape
        assert np.array_equal(obs, ref), f"{case} push {k}: max diff {np.abs(obs - ref).max():.3e}"


# ------------------------------------------------------------------------------------------------ env golden
ENV_CASES = ["env1_seed1", "2turb_seed1", "power_avg_yaw_dt2_total", "rich_3x1_global_base", "power_diff_change",
             "truncation_short"]


def _oracle_env(meta):
    kw = dict(meta["kw"])
    seed = kw.pop("seed")
    # the reference constructor resets once with the seed, the golden script resets again with the same seed
    env = WindFarmEnvOracle(V80(), meta["cfg"], seed=seed, reset_init=True, **kw)
    return env, seed


@pytest.mark.parametrize("case", ENV_CASES)
def test_env_golden(case):
    z, meta = _load("env_golden.npz")
    m = meta[case]
    env, seed = _oracle_env(m)
    obs0, _ = env.reset(seed=seed)
    assert env.ws == m["ws"] and env.ti == m["ti"] and env.wd == m["wd"]           # RNG draw order (:564-568)
    assert env.time_max == m["time_max"]
    assert env.fs.time == m["fs_time_after_reset"]
    assert env.obs_var == m["obs_var"]
    assert np.array_equal(np.asarray(env.fs.windTurbines.yaw), z[f"{case}/yaw0"])  # yaw init draw (:715)
    assert np.array_equal(obs0, z[f"{case}/obs0"])
    acts = z[f"{case}/acts"]
    for k, a in enumerate(acts):
        o, r, term, tr, info = env.step(a)
        assert term is False
        assert np.array_equal(o, z[f"{case}/obs"][k]), f"{case} step {k} obs"
        rr = z[f"{case}/reward"][k]
        assert (np.isnan(r) and np.isnan(rr)) or r == pytest.approx(rr, rel=1e-12, abs=1e-12), f"{case} step {k} reward"
        assert tr == bool(z[f"{case}/trunc"][k])
        assert np.allclose(info["Power pr turbine agent"], z[f"{case}/power"][k], rtol=1e-12)
        assert np.allclose(info["yaw angles agent"], z[f"{case}/yaw"][k], rtol=0, atol=1e-12)
        assert np.allclose(info["Wind speed at turbines"], z[f"{case}/ws_turb"][k], rtol=1e-12)
        assert np.allclose(info["Wind direction at turbines"], z[f"{case}/wd_turb"][k], rtol=1e-12)
        if env.Baseline_comp:
            assert np.allclose(info["Power pr turbine baseline"], z[f"{case}/power_base"][k], rtol=1e-12)
            assert np.allclose(info["yaw angles base"], z[f"{case}/yaw_base"][k], rtol=0, atol=1e-12)
    assert len(acts) == m["steps"]


def test_truncation_case_actually_truncates():
    z, meta = _load("env_golden.npz")
    assert z["truncation_short/trunc"][-1], "golden case must end in a truncated step"
    assert not z["truncation_short/trunc"][0]


def test_farmeval_constant_agent_known_answer():
    """FarmEval + ConstantAgent([-10, 20, 0, 0]) (reference tests/test_basics.py:415-459): 'wind' action, 1 deg/step."""
    z, meta = _load("env_golden.npz")
    m = meta["farmeval"]
    from tests.helpers import ENV1
    env = WindFarmEnvOracle(V80(), ENV1, yaw_init="Zeros", eval_mode=True, reset_init=False)
    env.set_wind_vals(ws=m["ws"], ti=m["ti"], wd=m["wd"])
    obs, _ = env.reset()
    assert env.time_max == m["time_max"] == 9999999
    action = z["farmeval/action"]
    assert np.allclose(action, [-10 / 45, 20 / 45, 0, 0])          # BaseAgent.scale_yaw (BaseAgent.py:17-23)
    yaws = [np.asarray(env.fs.windTurbines.yaw).copy()]
    assert np.array_equal(obs, z["farmeval/obs"][0])
    for k in range(25):
        obs, r, _, tr, info = env.step(action)
        assert not tr
        yaws.append(np.asarray(env.fs.windTurbines.yaw).copy())
        assert np.array_equal(obs, z["farmeval/obs"][k + 1])
        assert np.allclose(env.fs.windTurbines.power(), z["farmeval/power"][k], rtol=1e-12)
    yaws = np.array(yaws)
    assert np.allclose(yaws, z["farmeval/yaw"], atol=1e-12)
    assert np.allclose(yaws[5], [-5, 5, 0, 0]) and np.allclose(yaws[10], [-10, 10, 0, 0])
    assert np.allclose(yaws[20], [-10, 20, 0, 0]) and np.allclose(yaws[25], [-10, 20, 0, 0])
    assert env.fs.time == m["fs_time_after_25"]


# ------------------------------------------------------------------------------------------------ known answers
# Restated from /root/reference/tests/test_MesClass.py (line numbers in each test) and SURVEY.md 8(c).
def _mes(current, rolling, N=1, H=10, W=1):
    return MesO(current, rolling, N, H, W)


def test_mes_empty():                       # test_MesClass.py:23-27
    assert _mes(True, False).get().size == 0


def test_mes_current_only():                # :29-34
    m = _mes(True, False)
    m.add(5)
    assert np.array_equal(m.get(), np.array([5], dtype=np.float32))


def test_mes_rolling_only():                # :36-52
    m = _mes(False, True, N=3, W=2, H=10)
    m.add(2); m.add(4)
    assert np.array_equal(m.get(), np.array([3, 3, 3], dtype=np.float32))


def test_mes_current_and_rolling():         # :54-68
    m = _mes(True, True, N=3, W=1, H=10)
    for i in range(1, 4):
        m.add(i)
    assert np.array_equal(m.get(), np.array([3, 3, 2, 1], dtype=np.float32))


def test_mes_rolling_short_history():       # :70-84
    m = _mes(False, True, N=3, W=1, H=10)
    m.add(1); m.add(2)
    assert np.array_equal(m.get(), np.array([2, 2, 1], dtype=np.float32))


def test_mes_window_order_quirk():          # latest, ascending-from-oldest, oldest
    m = _mes(False, True, N=4, W=1, H=10)
    for i in range(10):
        m.add(i)
    assert np.array_equal(m.get(), np.array([9, 3, 6, 0], dtype=np.float32))
    m = _mes(True, True, N=3, W=5, H=25)
    for i in range(25):
        m.add(i)
    assert np.array_equal(m.get(), np.array([24, 22, 12, 2], dtype=np.float32))
    m = _mes(False, True, N=100, W=1, H=100)
    for i in range(3):
        m.add(i)
    out = m.get()
    assert out.shape == (100,) and out[0] == 2 and out[1] == 1 and out[-1] == 0 and np.all(out[2:-1] == 2)


def test_window_bounds_cover_history():
    for L in range(1, 40):
        for N in (1, 2, 3, 7):
            for W in (1, 4, 10, 50):
                for i in range(N):
                    lo, hi = window_bounds(L, N, W, i)
                    assert 0 <= lo < hi <= L


def _turb(ws=(False, True, 1, 10, 5), wd=(False, True, 1, 10, 5), yaw=(False, True, 2, 30, 1), power=(False, True, 1, 10, 5),
          include_TI=True):
    return TurbMesO(ws, wd, yaw, power, (0.0, 30.0, 0.0, 360.0, -45, 45, 0.0, 0.5), include_TI, 2000000)


def test_turb_mes_observed_variables_and_max_hist():   # test_MesClass.py:144-151
    t = _turb()
    assert t.n_out() == 6
    assert t.max_hist() == 30


def test_turb_mes_constant_history_gives_zero_TI():    # :153-172
    t = _turb()
    for _ in range(10):
        t.ws.add(10.0); t.wd.add(180.0); t.yaw.add(0.0); t.power.add(1.0e6)
    out = t.get(False)
    assert np.allclose(out, [10.0, 180.0, 0.0, 0.0, 0.0, 1.0e6])
    assert t.calc_TI()[0] == 0.0


def test_scale_val():                                   # :174-187
    assert scale_val(np.float32(15.0), 0.0, 30.0) == 0.0
    assert scale_val(np.float32(0.0), 0.0, 30.0) == -1.0
    assert scale_val(np.float32(30.0), 0.0, 30.0) == 1.0


def test_farm_mes_env1_known_answer():                  # SURVEY.md 8(c) golden vector
    z, meta = _load("mes_golden.npz")
    fm = farm_mes_from_kwargs(4, meta["env1_T4"]["kwargs"])
    fm.add(np.array([5.0, 6, 7, 8]), np.full(4, 270.0), np.array([-10.0, 0, 10, 20]), np.zeros(4))
    exp = np.array([-0.73913044, -0.22222221, -0.65217388, 0, -0.56521738, 0.22222221, -0.47826087, 0.44444442],
                   dtype=np.float32)
    assert np.array_equal(fm.get(True).astype(np.float32), exp)
    assert fm.n_out() == 8


# ------------------------------------------------------------------------------------------------ multi-agent golden
MULTI_CASES = ["multi_4x2_env1", "multi_rich_2x2", "multi_truncation"]


@pytest.mark.parametrize("case", MULTI_CASES)
def test_multi_agent_golden(case):
    """Oracle's per-agent split (FarmMesO.get_multi) == the unmodified WindFarmEnvMulti._get_obs_multi
    (WindEnvMulti.py:79-103), bit for bit; shared reward; the reference's half-length episodes."""
    z, meta = _load("multi_golden.npz")
    m = meta[case]
    kw = dict(m["kw"])
    seed = kw.pop("seed")
    kw.setdefault("n_passthrough", 20)                       # WindFarmEnvMulti's default (WindEnvMulti.py:25)
    env = WindFarmEnvOracle(V80(), m["cfg"], seed=seed, reset_init=True, **kw)
    env.reset(seed=seed)
    assert (env.ws, env.ti, env.wd, env.time_max) == (m["ws"], m["ti"], m["wd"], m["time_max"])
    assert env.fs.time == m["fs_time_after_reset"]
    assert np.array_equal(np.asarray(env.fs.windTurbines.yaw), z[f"{case}/yaw0"])
    o0 = np.stack(env.mes.get_multi())
    # declared obs dim over-counts by the farm-yaw ghost channels that never receive data
    assert o0.shape == (m["n_turb"], m["obs_len"]) and m["obs_len"] <= m["declared_obs_var"]
    assert np.array_equal(o0, z[f"{case}/obs0"])
    for k, a in enumerate(z[f"{case}/acts"]):
        o, r, _, tr, info = env.step(a)
        assert not tr
        assert np.array_equal(np.stack(env.mes.get_multi()), z[f"{case}/obs"][k]), f"{case} step {k}"
        assert r == pytest.approx(float(z[f"{case}/reward"][k]), rel=1e-12, abs=1e-12)
        assert np.allclose(info["Power pr turbine agent"], z[f"{case}/power"][k], rtol=1e-12)
        assert np.allclose(info["yaw angles agent"], z[f"{case}/yaw"][k], rtol=0, atol=1e-12)
        env.timestep += 1                                    # the second increment of WindEnvMulti.py:219
    if "truncating_step" in m:
        # the next step is the one the reference truncates in (and cannot return from, see make_golden.make_multi)
        assert len(z[f"{case}/acts"]) == m["truncating_step"]
        assert env.timestep >= env.time_max and (env.timestep - 2) < env.time_max
        assert m["truncating_step"] == (m["time_max"] + 1) // 2


def test_cfg1_2turb_200_steps_golden():
    """BASELINE.json configs[0]: shipped 2turb.yaml, seed 1, 200 CPU steps (zero action, then +0.5)."""
    z, meta = _load("cfg1_golden.npz")
    m = meta["cfg1"]
    env = WindFarmEnvOracle(V80(), m["cfg"], seed=1, reset_init=True)
    obs0, _ = env.reset(seed=1)
    assert (env.ws, env.ti, env.wd, env.time_max, env.obs_var) == (m["ws"], m["ti"], m["wd"], m["time_max"], m["obs_var"])
    assert np.array_equal(obs0, z["cfg1/obs0"])
    acts = z["cfg1/acts"]
    assert acts.shape == (200, 2) and np.all(acts[:100] == 0) and np.all(acts[100:] == 0.5)
    for k, a in enumerate(acts):
        o, r, term, tr, info = env.step(a)
        assert term is False and not tr
        assert np.array_equal(o, z["cfg1/obs"][k]), f"step {k}"
        assert r == pytest.approx(float(z["cfg1/reward"][k]), rel=1e-12, abs=1e-12)
        assert np.allclose(info["Power pr turbine agent"], z["cfg1/power"][k], rtol=1e-12)
        assert np.allclose(info["Power pr turbine baseline"], z["cfg1/power_base"][k], rtol=1e-12)
        assert np.allclose(info["yaw angles agent"], z["cfg1/yaw"][k], rtol=0, atol=1e-12)
    assert env.fs.time == m["fs_time_end"]
    # the yaw offsets moved by +0.5 * yaw_step per step in the second half ("yaw" action) or towards the set point
    assert not np.allclose(z["cfg1/yaw"][99], z["cfg1/yaw"][199])

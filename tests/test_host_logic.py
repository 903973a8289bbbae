"""CPU: host-side mirror of the reference interface (windgym_b200/config.py, turbines.py): YAML schema, constructor
errors, layout rule, reset integers and RNG draw order -- against the reference's golden values (SURVEY.md 8c,
tests/golden/env_golden.npz).  No GPU, no oracle needed except as the holder of the golden numbers."""
import json
import os

import numpy as np
import pytest

from tests.helpers import ENV1, rich_config, small_config
from windgym_b200 import V80, EnvConfig, grid_layout
from windgym_b200.config import rotate_layout

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_layout_spacing_quirk():
    """x = linspace(0, D*xDist*nx, nx): 2 turbines at xDist=4 are 8 D apart (Wind_Farm_Env.py:246-252)."""
    x, y = grid_layout(80.0, 4, 4, 2, 1)
    assert np.allclose(x, [0, 640]) and np.allclose(y, [0, 0])
    x, y = grid_layout(80.0, 4, 4, 4, 4)
    assert np.allclose(np.unique(x), [0, 426.6666667, 853.3333333, 1280])
    assert x.shape == (16,) and np.allclose(x[:4], np.unique(x)) and np.allclose(y[:4], 0)
    x, y = grid_layout(80.0, 4, 4, 1, 3)
    assert np.allclose(x, 0)


def test_rotation_matches_reference_notebook():
    """wd = 266.83 deg: delta (640, 0) -> (639.02, -35.36) (Example 1 notebook :288-289)."""
    xr, yr = rotate_layout(np.array([0.0, 640.0]), np.array([0.0, 0.0]), np.array(266.83))
    assert xr[1] - xr[0] == pytest.approx(639.02, abs=0.01)
    assert yr[1] - yr[0] == pytest.approx(-35.36, abs=0.05)


def test_v80_tables():
    t = V80()
    assert t.diameter() == 80.0 and t.hub_height() == 70.0
    assert float(max(t.power(np.arange(10, 25, 1)))) == 2.0e6        # maxturbpower (Wind_Farm_Env.py:112)
    assert t.power(10.0) == pytest.approx(1341e3)
    assert t.power(10.0, yaw=30.0) == pytest.approx(np.interp(10 * np.cos(np.pi / 6), t.ws_table, t.power_table_w))
    assert t.ct(8.0, yaw=20.0) == pytest.approx(np.interp(8 * np.cos(np.deg2rad(20)), t.ws_table, t.ct_table)
                                                 * np.cos(np.deg2rad(20)) ** 2)
    assert t.power(2.0) == 0.0 and t.power(30.0) == 2.0e6            # np.interp clamps outside 3..25 m/s


def test_config_defaults_and_derived():
    ec = EnvConfig(ENV1, V80())
    assert ec.n_turb == 4 and ec.S == 1 and ec.Baseline_comp            # Power_reward "Baseline" forces the 2nd farm
    assert ec.hist_max == 25 and ec.steps_on_reset == 25               # max(ws, wd, yaw history) (:224-240)
    assert ec.yaw_init_mode == "Random" and ec.ActionMethod == "wind"
    assert ec.p_cap % 8 == 0 and ec.p_cap >= 64
    ec = EnvConfig(ENV1, V80(), fill_window=5)
    assert ec.steps_on_reset == 5
    ec = EnvConfig(ENV1, V80(), fill_window=False)
    assert ec.steps_on_reset == 1
    ec = EnvConfig(ENV1, V80(), fill_window=1000)
    assert ec.steps_on_reset == 25
    ec = EnvConfig(small_config(2, 2, reward="Power_avg"), V80(), Baseline_comp=True)
    assert ec.Baseline_comp
    ec = EnvConfig(small_config(2, 2, reward="Power_avg"), V80(), yaw_init="Zeros")
    assert ec.yaw_init_mode == "Zeros"


@pytest.mark.parametrize("patch,kw,exc,msg", [
    ({}, dict(dt_env=3, dt_sim=2), ValueError, "dt_env must be a multiple of dt_sim"),                 # :107
    ({"ActionMethod": "absolute"}, {}, NotImplementedError, "absolute method is not implemented"),    # :861
    ({"ActionMethod": "bogus"}, {}, ValueError, "ActionMethod must be yaw, wind or absolute"),        # :864
    ({"power_def.Power_reward": "Bogus"}, {}, ValueError, "Power_reward must be either"),             # :192
    ({"power_def.Power_reward": "Power_diff", "power_def.Power_avg": 10}, {}, ValueError, "larger then 40"),  # :186
    ({"Track_power": True}, {}, NotImplementedError, "Track_power"),                                   # :168
    ({"BaseController": "PyWake"}, {}, ValueError, "BaseController must be either Local or Global"),  # :314
    ({}, dict(fill_window=-3), ValueError, "fill_window must be True or a non-negative integer"),     # :240
    ({}, dict(turbtype="Mann"), ValueError, "Invalid turbulence type"),                                # :666-668
])
def test_config_errors_match_reference(patch, kw, exc, msg):
    cfg = small_config(2, 2, reward="Baseline", **patch)
    with pytest.raises(exc, match=msg):
        EnvConfig(cfg, V80(), **kw)


def test_turbtype_random_is_accepted():
    assert EnvConfig(small_config(2, 2, reward="Baseline"), V80(), turbtype="Random").turbtype == "Random"   # :640-644


def test_reset_integers_match_reference_golden():
    """t_developed, time_max from the unmodified reference (Wind_Farm_Env.py:723-732) on the shipped YAMLs."""
    z = np.load(os.path.join(GOLD, "env_golden.npz"))
    meta = json.loads(str(z["meta"]))
    for case in ("env1_seed1", "2turb_seed1", "power_avg_yaw_dt2_total", "rich_3x1_global_base", "truncation_short"):
        m = meta[case]
        kw = {k: v for k, v in m["kw"].items() if k != "seed"}
        ec = EnvConfig(m["cfg"], V80(), **kw)
        n_spin, time_max, k_emit = ec.reset_integers(np.array([m["ws"]]), np.array([m["wd"]]))
        assert int(time_max[0]) == m["time_max"], case
        # fs.time after reset = t_developed + steps_on_reset * dt_env (:734-766)
        assert n_spin[0] * ec.dt_sim + ec.steps_on_reset * ec.dt_env == m["fs_time_after_reset"], case
        assert k_emit[0] == max(1, int(np.ceil(0.2 * 80.0 / (m["ws"] * ec.dt_sim) - 1e-9)))
    ec = EnvConfig(ENV1, V80(), eval_mode=True)
    assert ec.reset_integers(np.array([10.0]), np.array([270.0]))[1][0] == 9999999   # FarmEval.py:59


def test_survey_golden_wind_draws():
    """reset(seed=1) draw order ws -> ti -> wd -> yaw (SURVEY.md 8c golden values from the unmodified reference)."""
    rng = np.random.default_rng(1)
    w = ENV1["wind"]
    ws = rng.uniform(w["ws_min"], w["ws_max"]); ti = rng.uniform(w["TI_min"], w["TI_max"])
    wd = rng.uniform(w["wd_min"], w["wd_max"]); yaw = rng.uniform(-15, 15, 4)
    assert ws == 11.094572997602054 and ti == 0.1435602805223716 and wd == 259.32478838158903
    assert np.allclose(yaw, [13.45948341, -5.64505644, -2.30020653, 9.83107781])
    ec = EnvConfig(ENV1, V80())
    n_spin, time_max, _ = ec.reset_integers(np.array([ws]), np.array([wd]))
    assert time_max[0] == 336 and n_spin[0] + 25 == 159


def test_p_cap_never_overflows_for_any_direction():
    """Chain capacity covers the farm diagonal + margin at the tightest possible particle spacing."""
    for nx, ny in ((2, 1), (4, 4), (8, 8), (1, 5)):
        ec = EnvConfig(small_config(nx, ny), V80())
        diag = np.hypot(np.ptp(ec.x_pos), np.ptp(ec.y_pos))
        min_spacing = ec.d_particle * ec.D * ec.f_min
        assert ec.p_cap * min_spacing >= diag + 2 * ec.D
    assert EnvConfig(small_config(4, 4), V80()).p_cap == 168
    assert EnvConfig(rich_config(8, 8), V80()).p_cap % 8 == 0


def test_fast_rng_replicates_numpy_streams():
    """windgym_b200.fast_rng evaluates default_rng([seed, env, episode]) for many envs with array arithmetic: bit for
    bit the same doubles as numpy's SeedSequence + PCG64, and sample_conditions gives identical draws on either path."""
    from types import SimpleNamespace
    from windgym_b200.fast_rng import uniform_streams
    from windgym_b200.vec_env import VecWindFarmEnv
    rng = np.random.default_rng(0)
    for seed, ep in [(0, 0), (7, 3), (2 ** 32 - 1, 2 ** 32 - 1), (123456789, 41)]:
        envs = np.concatenate([np.arange(40), rng.integers(0, 2 ** 32, 40)])
        ref = np.array([np.random.default_rng([seed, int(i), ep]).random(11) for i in envs])
        assert np.array_equal(uniform_streams(seed, envs, ep, 11), ref)
    with pytest.raises(ValueError):
        uniform_streams(2 ** 32, [0], 0, 1)
    B, T = 64, 5
    ec = SimpleNamespace(ws_min=7, ws_max=15.5, TI_min=0.02, TI_max=0.15, wd_min=255, wd_max=285.0, turbtype="None",
                         yaw_init_mode="Random", yaw_start=15)
    out = []
    for no_fast in (False, True):
        stub = SimpleNamespace(ec=ec, n_envs=B, n_turb=T, ws=np.zeros(B), ti=np.zeros(B), wd=np.zeros(B), _episode=2,
                               sample_site=None, _wind_override={}, yaw_initial=None, _no_fast_rng=no_fast)
        out.append(VecWindFarmEnv.sample_conditions(stub, 11, range(3, 60)))
    for a, b in zip(*out):
        assert np.array_equal(a, b)
    assert out[0][0][3] != 0 and out[0][0][0] == 0 and np.abs(out[0][3][10]).max() <= 15


# ------------------------------------------------------------------------------------------------ NetCDF-4 / HDF5 boxes
def test_hdf5_min_reads_a_mann_box_file(tmp_path):
    """windgym_b200/hdf5_min.py on a synthetic HDF5 file in the 'earliest' layout (superblock 0, version-1 headers,
    symbol-table group) with a contiguous and a chunked + shuffle + deflate copy of the field."""
    from tests import hdf5_writer
    from windgym_b200 import hdf5_min
    rng = np.random.default_rng(0)
    uvw = rng.normal(size=(3, 12, 6, 5)).astype(np.float32)
    x, y, z = np.arange(12) * 4.0, np.arange(6) * 8.0, np.arange(5) * 8.0 + 10.0
    for chunked in ((), ("uvw",)):
        p = str(tmp_path / f"box_{len(chunked)}.nc")
        hdf5_writer.write(p, {"uvw": uvw, "x": x, "y": y, "z": z, "seed": np.array([7], dtype=np.int64)}, chunked=chunked)
        assert hdf5_min.list_datasets(p) == ["seed", "uvw", "x", "y", "z"]
        got, dxyz = hdf5_min.read_mann_box(p)
        assert got.dtype == np.float32 and np.array_equal(got, uvw) and dxyz == (4.0, 8.0, 8.0)
        assert hdf5_min.read_datasets(p, names=("seed",))["seed"].tolist() == [7]
    with pytest.raises(ValueError):
        (tmp_path / "junk.nc").write_bytes(b"not hdf5" * 100)
        hdf5_min.read_mann_box(str(tmp_path / "junk.nc"))


@pytest.mark.reference
def test_hdf5_min_reads_the_reference_netcdf4_file():
    """The reference's own NetCDF-4 file (xarray -> netCDF4: superblock 2, version-2 object headers, dense links in a
    fractal heap, contiguous float64 datasets): names, shapes and internal consistency of what the reader returns."""
    from windgym_b200 import hdf5_min
    path = "/root/reference/examples/PPO_eval.nc"
    if not os.path.isfile(path):
        pytest.skip("reference checkout not present")
    ds = hdf5_min.read_datasets(path)
    assert set(ds) >= {"time", "ws", "wd", "TI", "turb", "powerF_a", "powerT_a", "yaw_a", "ws_a", "reward", "pct_inc"}
    assert ds["ws"].tolist() == [10, 11, 12, 14] and ds["wd"].tolist() == [260, 265, 270, 275, 280]
    assert ds["time"].shape == (2039,) and np.all(np.diff(ds["time"]) == 1.0)
    assert ds["powerT_a"].shape == (2039, 4, 4, 5, 1, 1) and ds["powerF_a"].shape == (2039, 4, 5, 1, 1)
    ok = np.isfinite(ds["powerF_a"])
    assert ok.mean() > 0.9 and np.array_equal(ds["powerF_a"][ok], ds["powerT_a"].sum(axis=1)[ok])   # farm = sum of turbines
    assert np.nanmax(ds["powerT_a"]) <= 2.0e6 and np.nanmin(ds["yaw_a"]) >= -45 and np.nanmax(ds["yaw_a"]) <= 45

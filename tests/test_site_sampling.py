"""CPU: site-based wind sampling (reference Wind_Farm_Env.py:569-596, tests/test_basics.py:369-407):
sector frequency -> Weibull(A, k) speed -> clip to the YAML limits; TI stays uniform."""
import types

import numpy as np

from windgym_b200.sites import WeibullSite
from windgym_b200.vec_env import VecWindFarmEnv


def _sampler(site, n_envs, seed, **wind):
    """A VecWindFarmEnv shell without a device: only the host-side sampling logic is exercised."""
    env = VecWindFarmEnv.__new__(VecWindFarmEnv)
    w = dict(ws_min=4, ws_max=20, TI_min=0.02, TI_max=0.15, wd_min=200, wd_max=330)
    w.update(wind)
    env.ec = types.SimpleNamespace(yaw_init_mode="Zeros", yaw_start=15.0, turbtype="None", **w)
    env.n_envs, env.n_turb, env.sample_site, env._site_tables = n_envs, 2, site, None
    env.ws, env.ti, env.wd = np.zeros(n_envs), np.zeros(n_envs), np.zeros(n_envs)
    env._wind_override, env._episode, env.yaw_initial = {}, 1, [0]
    return env.sample_conditions(seed)


def test_site_sampling_follows_the_sector_frequencies_and_weibull():
    freq = np.zeros(12); freq[8], freq[9] = 0.25, 0.75            # sectors centred on 240 and 270 deg
    site = WeibullSite(freq, A=np.full(12, 10.0), k=np.full(12, 2.0))
    lw = site.local_wind(wd=np.arange(360), ws=np.arange(3, 25))
    f = lw.Sector_frequency_ilk[0, :, 0]
    assert f.shape == (360,) and abs(f.sum() - 1) < 1e-12 and f[270] > 0 and f[100] == 0
    ws, ti, wd, yaw0 = _sampler(site, 4000, seed=3)
    assert set(np.unique(np.round(wd))) <= set(range(225, 286))      # only the two populated sectors
    assert 0.70 < np.mean(wd >= 255) < 0.80                          # 75 % of the draws from the 270 sector
    # Weibull(A=10, k=2): mean = A*Gamma(1.5) = 8.86 (clipping at [4, 20] moves it slightly up)
    assert 8.6 < ws.mean() < 9.4 and ws.min() >= 4 and ws.max() <= 20
    assert ti.min() >= 0.02 and ti.max() <= 0.15 and yaw0.shape == (4000, 2)


def test_site_sampling_is_clipped_and_seeded():
    site = WeibullSite(np.full(12, 1 / 12), A=np.full(12, 9.0), k=np.full(12, 2.2))
    a = _sampler(site, 300, seed=11, wd_min=260, wd_max=280, ws_min=8, ws_max=9)
    b = _sampler(site, 300, seed=11, wd_min=260, wd_max=280, ws_min=8, ws_max=9)
    c = _sampler(site, 300, seed=12, wd_min=260, wd_max=280, ws_min=8, ws_max=9)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and not np.array_equal(a[0], c[0])
    assert a[2].min() >= 260 and a[2].max() <= 280 and a[0].min() >= 8 and a[0].max() <= 9
    assert (a[2] == 260).any() and (a[2] == 280).any()              # directions outside the range are clipped, not rejected
    assert len(np.unique(a[0])) > 10                                 # wind speeds vary (test_basics.py:369-407)


def _shell(site, n_envs, yaw_mode="Zeros"):
    env = VecWindFarmEnv.__new__(VecWindFarmEnv)
    env.ec = types.SimpleNamespace(yaw_init_mode=yaw_mode, yaw_start=15.0, turbtype="None", ws_min=4, ws_max=20,
                                   TI_min=0.02, TI_max=0.15, wd_min=200, wd_max=330)
    env.n_envs, env.n_turb, env.sample_site, env._site_tables = n_envs, 2, site, None
    env.ws, env.ti, env.wd = np.zeros(n_envs), np.zeros(n_envs), np.zeros(n_envs)
    env._wind_override, env._episode, env.yaw_initial = {}, 1, [0]
    return env


def test_site_sampling_with_pinned_wind_and_defined_yaw_touches_every_env():
    """Regression (round-1 advisor finding): the sector draw used to overwrite the env index list, so that
    set_wind_vals applied to one env and yaw_init='Defined' indexed out of bounds."""
    site = WeibullSite(np.full(12, 1 / 12), A=np.full(12, 9.0), k=np.full(12, 2.2))
    env = _shell(site, 400, yaw_mode="Defined")
    env.set_wind_vals(ws=11.5, wd=271.0)
    env.set_yaw_vals([5.0, -7.0])
    ws, ti, wd, yaw0 = env.sample_conditions(seed=5)
    assert np.all(ws == 11.5) and np.all(wd == 271.0)               # the override reaches all 400 envs
    assert len(np.unique(ti)) > 100                                  # TI still sampled per env
    assert np.array_equal(yaw0, np.tile([5.0, -7.0], (400, 1)))
    env4 = _shell(site, 4, yaw_mode="Defined")                       # B=4: used to raise IndexError (index 37)
    env4.set_yaw_vals([3.0])
    assert np.array_equal(env4.sample_conditions(seed=1)[3], np.full((4, 2), 3.0))
    # masked subset: only the selected envs change
    env.ws[:] = -1.0
    ws2, *_ = env.sample_conditions(seed=5, envs=[3, 7])
    assert np.all(ws2[[3, 7]] == 11.5) and np.all(np.delete(ws2, [3, 7]) == -1.0)

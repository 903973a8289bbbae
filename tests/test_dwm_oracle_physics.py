"""CPU: the restated flow solver (oracle/dwm_numpy.py; PARITY UNPINNED -- dynamiks is not vendored) against analytic
limits and the weak plausibility anchors of SURVEY.md A.5.  These keep the frozen specification honest."""
import numpy as np
import pytest

from oracle import dwm_numpy as dwm
from oracle.v80 import V80


def _sim(x, y, ws=10.0, wd=270.0, yaw=None, steps=0):
    wt = dwm.PyWakeWindTurbines(np.asarray(x, float), np.asarray(y, float), V80())
    fs = dwm.DWMFlowSimulation(dwm.TurbulenceFieldSite(ws, dwm.RandomTurbulence(0, ws)), wt, wind_direction=wd, dt=1,
                               d_particle=0.2)
    if yaw is not None:
        wt.yaw = yaw
    for _ in range(steps):
        fs.step()
    return fs, wt


def test_single_turbine_sees_free_stream():
    fs, wt = _sim([0.0], [0.0], ws=9.0, steps=60)
    assert np.allclose(fs.rotor_avg_windspeed, [[9.0, 0.0, 0.0]])
    assert wt.power()[0] == pytest.approx(996e3)                      # V80 table at 9 m/s
    fs.run(5)
    assert fs.time == 65.0 and fs.n_step == 65                        # fs.time bookkeeping


def test_yaw_reduces_power_by_cos_law():
    fs, wt = _sim([0.0], [0.0], ws=9.0, yaw=[25.0], steps=5)
    assert wt.power()[0] == pytest.approx(V80().power(9.0 * np.cos(np.deg2rad(25.0))))


def test_aligned_row_zero_yaw_is_symmetric_and_waked():
    fs, wt = _sim([0.0, 640.0], [0.0, 0.0], ws=10.0, steps=160)
    u = fs.rotor_avg_windspeed
    assert np.all(u[:, 1] == 0.0) and np.all(u[:, 2] == 0.0)          # no lateral component without yaw
    assert u[0, 0] == pytest.approx(10.0)
    assert 4.0 < u[1, 0] < 9.5                                        # 8 D downstream, zero ambient TI: deep wake
    assert wt.power()[1] < wt.power()[0]


def test_upstream_yaw_deflects_wake_and_helps_downstream_turbine():
    """The Hill-vortex term is what makes yaw matter (reference comment, Wind_Farm_Env.py:711)."""
    _, wt0 = _sim([0.0, 640.0], [0.0, 0.0], ws=9.0, yaw=[0.0, 0.0], steps=200)
    fs1, wt1 = _sim([0.0, 640.0], [0.0, 0.0], ws=9.0, yaw=[25.0, 0.0], steps=200)
    assert wt1.power()[1] > wt0.power()[1]
    assert fs1.rotor_avg_windspeed[1, 1] != 0.0                       # deflected wake induces a lateral component
    # mirror symmetry: opposite yaw gives the mirrored lateral velocity and the same power
    fs2, wt2 = _sim([0.0, 640.0], [0.0, 0.0], ws=9.0, yaw=[-25.0, 0.0], steps=200)
    assert wt2.power()[1] == pytest.approx(wt1.power()[1], rel=1e-9)
    assert fs2.rotor_avg_windspeed[1, 1] == pytest.approx(-fs1.rotor_avg_windspeed[1, 1], rel=1e-9)


def test_ainslie_march_conserves_momentum_deficit():
    """Thin-shear-layer march: the momentum-deficit integral int U (1 - U) r dr is conserved."""
    a = np.array([0.15, 0.25, 0.33])
    U = dwm.inlet_profile(a)
    r = np.arange(dwm.N_R) * dwm.DR
    mom0 = np.sum(U * (1 - U) * r, axis=1)
    xt = np.zeros(3)
    for k in range(200):
        U, nu = dwm.ainslie_march(U, np.full(3, 0.2), xt + 0.1, np.full(3, 0.023 * 0.08 ** 0.3))
        xt += 0.2
        assert np.all(nu > 0)
    mom1 = np.sum(U * (1 - U) * r, axis=1)
    assert np.allclose(mom1, mom0, rtol=2e-2)
    assert np.all(U[:, 0] > 1 - 2 * a) and np.all(U <= 1.0 + 1e-9)    # the wake recovers and never overshoots
    assert np.all(np.diff(U, axis=1) > -1e-9)                         # monotone in r


def test_inlet_profile_is_continuous_in_induction():
    a = np.linspace(0.05, 0.38, 200)
    U = dwm.inlet_profile(a)
    assert np.all(U[:, -1] == 1.0) and np.allclose(U[:, 0], 1 - 2 * a)
    eps = 1e-6                                                       # cell-averaged top hat: no jumps in a
    assert np.abs(dwm.inlet_profile(a + eps) - U).max() < 1e-3


def test_plausibility_anchor_2x2_notebook():
    """Order-of-magnitude anchor from the reference notebook: 2x2 V80 farm, 8 D pitch, ws 9.3,
    wd 266.8: front row ~ free stream, second row between ~50 % and ~95 % of it.  NOT a parity target."""
    fs, wt = _sim([0.0, 640.0, 0.0, 640.0], [0.0, 0.0, 640.0, 640.0], ws=9.30, wd=266.8, steps=250)
    u = fs.rotor_avg_windspeed[:, 0]
    order = np.argsort(fs.positions_xyz[0])
    assert np.allclose(u[order[:2]], 9.30, atol=1e-6)
    assert np.all(u[order[2:]] > 0.45 * 9.30) and np.all(u[order[2:]] < 9.30)
    assert 2.0e6 < wt.power().sum() < 4.0e6                           # notebook: 3.17 MW


def test_chain_capacity_never_overflows():
    fs, wt = _sim(np.linspace(0, 1280, 4), np.zeros(4), ws=7.0, yaw=[30, -30, 30, -30], steps=400)
    assert fs.overflow == 0
    assert fs.count.max() <= fs.P


def test_emission_cadence_sensitivity():
    """The frozen specification releases one particle per chain every k_emit = ceil(16 m / (ws dt)) steps (20-30 m
    apart at dt = 1 s) instead of "every d_particle D = 16 m of travel" (reference Wind_Farm_Env.py:116, :708 hands
    d_particle to dynamiks, whose release rule cannot be read here).  How much does the spacing matter?  The same
    farm with (a) the cadence rule, (b) a particle EVERY step (10-12 m apart: denser than 16 m), (c) a release at the
    first step boundary after the newest particle travelled 16 m: every rotor's speed moves by < 1e-4 ws and
    its power by < 3e-4 (measured: 2e-5 and 1e-4) between a 2x finer and a 1.5x coarser wake discretisation than the
    nominal one -- the wake is resolved by linear interpolation between stations whose profiles change slowly in x.  (DESIGN.md section 2 states the bound.)"""
    x, y = [0.0, 560.0, 1120.0, 0.0, 560.0, 1120.0], [0.0, 0.0, 0.0, 400.0, 400.0, 400.0]
    yaw = [20.0, 0.0, 0.0, -15.0, 10.0, 0.0]
    out = {}
    for ws in (8.5, 12.0):
        for rule in ("cadence", 1, "distance"):
            wt = dwm.PyWakeWindTurbines(np.asarray(x, float), np.asarray(y, float), V80())
            fs = dwm.DWMFlowSimulation(dwm.TurbulenceFieldSite(ws, dwm.RandomTurbulence(0, ws)), wt, wind_direction=268.0,
                                       dt=1, d_particle=0.2, emit_rule=rule, p_cap=400)
            fs.ti = 0.08
            wt.yaw = yaw
            for _ in range(420):
                fs.step()
            assert fs.overflow == 0
            out[(ws, rule)] = (fs.rotor_avg_windspeed[:, 0].copy(), wt.power().copy(), int(fs.count.sum()))
        u_c, p_c, n_c = out[(ws, "cadence")]
        for rule in (1, "distance"):
            u, p, n = out[(ws, rule)]
            du = np.abs(u - u_c) / ws
            dp = np.abs(p - p_c) / np.maximum(p_c, 1.0)
            print(f"ws {ws}: rule {rule!r}: stations {n} vs {n_c}; max |du|/ws {du.max():.2e}; max rel dP {dp.max():.2e}; "
                  f"farm power {abs(p.sum() - p_c.sum()) / p_c.sum():.2e}")
            assert du.max() < 1e-4 and dp.max() < 3e-4 and abs(p.sum() - p_c.sum()) / p_c.sum() < 1e-4
        assert out[(ws, 1)][2] > 1.5 * n_c                       # the dense rule really has (many) more stations

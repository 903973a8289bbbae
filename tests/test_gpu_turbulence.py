"""GPU: ambient Mann turbulence -- device generator vs oracle, and the flow kernel with a box
(meandering wake centres + rotor-plane fluctuations) against the fp64 oracle on the same box, offsets and TI."""
import numpy as np
import pytest

from oracle import mann_numpy as mn
from tests.helpers import oracle_rollout, small_config

pytestmark = pytest.mark.gpu

N, D3 = (256, 64, 32), (4.0, 6.0, 6.0)     # 1024 m x 384 m x 192 m periodic box


def _boxes(lowpass=160.0):
    import torch
    from windgym_b200.mann import MannBox
    noise = mn.box_noise(N, 11)
    ref = mn.mann_box(0.1, 33.6, 3.9, N, D3, noise=noise)
    box = MannBox.generate(0.1, 33.6, 3.9, N, D3, device="cuda:0", noise=noise, lowpass_width=lowpass)
    return ref, box


def test_device_generator_matches_oracle(built_lib):
    ref, box = _boxes()
    got = box.raw[..., :3].permute(3, 0, 1, 2).cpu().numpy()
    assert np.abs(got - ref).max() < 2e-5 * ref.std()
    f = mn.MannTurbulenceField(ref, D3, lowpass_width=160.0)
    assert np.abs(box.lp.permute(3, 0, 1, 2).cpu().numpy() - f.uvw_lp[1:]).max() < 2e-5 * ref.std()
    # production path: seeded device noise, complex64 transform for big boxes -- statistics only
    from windgym_b200.mann import MannBox
    big = MannBox.generate(Nxyz=(512, 128, 64), dxyz=(3.0, 3.0, 3.0), seed=1234, device="cuda:0")
    u = big.raw[..., 0]
    assert abs(float(u.mean())) < 1e-3 and 0.5 < big.std_u < 3.0
    assert float(big.raw[..., 0].std()) > float(big.raw[..., 1].std()) > float(big.raw[..., 2].std())


N_ISO = (64, 32, 32)


def _iso_boxes():
    """Unit-variance isotropic box of the wake-added turbulence, oracle (fp64) and device, from the same noise."""
    from windgym_b200.mann import MannBox
    noise = mn.box_noise(N_ISO, 23)
    ref = mn.mann_box(1.0, 10.0, 0.0, N_ISO, (5.0, 5.0, 5.0), noise=noise)
    ref = ref / ref[0].std()
    box = MannBox.isotropic_unit(80.0, device="cuda:0", Nxyz=N_ISO, noise=noise)
    assert abs(float(box.raw[..., 0].double().std(unbiased=False)) - 1.0) < 1e-5
    return ref, box


@pytest.mark.parametrize("reward,added", [("Power_avg", False), ("Baseline", False), ("Power_avg", True), ("Baseline", True)])
def test_flow_with_turbulence_box_vs_oracle(built_lib, reward, added):
    import torch
    from windgym_b200 import V80, VecWindFarmEnv
    ref_box, box = _boxes()
    ref_iso, iso = _iso_boxes() if added else (None, False)
    cfg = small_config(2, 2, reward=reward, action="wind")
    B, T, steps = 3, 4, 6
    rng = np.random.default_rng(3)
    ws, ti, wd = rng.uniform(8, 13, B), rng.uniform(0.05, 0.12, B), rng.uniform(262, 278, B)
    yaw0 = rng.uniform(-15, 15, (B, T))
    off = rng.uniform(0, 1, (B, 3)) * (np.array(N) * np.array(D3))
    acts = rng.uniform(-1, 1, (steps, B, T)).astype(np.float32)
    env = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", turbtype="MannFixed", turb_box=box, added_turbulence=iso)
    obs0 = env.reset(wind=(ws, ti, wd), yaw0=yaw0, turb_offset=off)[0].cpu().numpy().copy()
    pw, yw, ob, rw, uvw = [], [], [], [], []
    for a in acts:
        o, r, _, _, info = env.step(torch.as_tensor(a))
        pw.append(info["Power pr turbine agent"].cpu().numpy().copy()); ob.append(o.cpu().numpy().copy())
        rw.append(r.cpu().numpy().copy())
        uvw.append(torch.stack([env.state[k][:, 0] for k in ("u", "v", "w")], -1).cpu().numpy().copy())
    env.check_flags()
    pw, ob, rw, uvw = np.array(pw), np.array(ob), np.array(rw), np.array(uvw)
    z = env.state["pmut"][:, :, 0].cpu().numpy()[..., 2]
    assert np.abs(uvw[..., 2]).max() > 1e-3 and np.abs(z[z != 0] - 70.0).max() > 0.05   # w' at rotors, wakes meander in z
    for b in range(B):
        field = mn.MannTurbulenceField(ref_box, D3, lowpass_width=160.0)
        afield = mn.MannTurbulenceField(ref_iso, (5.0, 5.0, 5.0), lowpass_width=5.0) if added else None
        ref = oracle_rollout(cfg, ws[b:b + 1], ti[b:b + 1], wd[b:b + 1], yaw0[b:b + 1], acts[:, b:b + 1],
                             turbtype="MannFixed", turb_field=field, added_field=afield,
                             reset_kw=dict(turb_offset=off[b]))
        rel = np.abs(pw[:, b] - ref["power"][0]) / np.maximum(ref["power"][0], 1.0)
        assert rel.max() < 1e-4, f"env {b}: power rel err {rel.max():.3e}"
        assert np.allclose(obs0[b], ref["obs0"][0], atol=2e-5)
        assert np.allclose(ob[:, b], ref["obs"][0], atol=2e-5)
        assert np.allclose(rw[:, b], ref["reward"][0], rtol=2e-4, atol=2e-5)
    if added:   # the added turbulence acts inside wakes only: the most upstream rotor is untouched by it
        env1 = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", turbtype="MannFixed", turb_box=box,
                              added_turbulence=False)
        env1.reset(wind=(ws, ti, wd), yaw0=yaw0, turb_offset=off)
        for a in acts:
            env1.step(torch.as_tensor(a))
        u1 = env1.state["u"][:, 0].cpu().numpy()
        up_ = env.state["xr"].cpu().numpy().argmin(axis=1)
        assert np.array_equal(u1[np.arange(B), up_], uvw[-1][np.arange(B), up_, 0])
        assert np.abs(u1 - uvw[-1][..., 0]).max() > 1e-3
    # turbulence makes the rotor inflow fluctuate: the same farm without a box sees a steady free-stream front row
    env0 = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0")
    env0.reset(wind=(ws, ti, wd), yaw0=yaw0)
    up = env.state["xr"].cpu().numpy().argmin(axis=1)
    u_front = uvw[:, np.arange(B), up, 0]
    assert np.abs(u_front - ws[None]).max() > 0.05 and np.allclose(env0.state["u"][np.arange(B), 0, up].cpu().numpy(), ws, rtol=1e-6)


def test_facade_with_mann_box_runs_and_meanders(built_lib):
    from windgym_b200 import V80, WindFarmEnv
    _, box = _boxes()
    cfg = small_config(2, 1, reward="Power_avg", action="yaw")
    env = WindFarmEnv(V80(), config=cfg, turbtype="MannGenerate", turb_box=box, seed=2, device="cuda:0")
    assert env.vec.added_box is not None and env.vec.added_box.Nxyz == (128, 64, 64)   # default addedTurbulenceModel
    obs, info = env.reset(seed=2)
    p = []
    for _ in range(20):
        obs, r, term, trunc, info = env.step(np.zeros(2, dtype=np.float32))
        p.append(info["Power agent"])
    assert np.isfinite(p).all() and np.std(p) > 1e-3 * np.mean(p)      # power fluctuates with the inflow
    assert info["Turbulence intensity"] == env.ti and 0.02 <= env.ti <= 0.15
    img_env = env.fs.get_windspeed(type("V", (), dict(x=np.linspace(-400, 800, 40), y=np.linspace(-200, 200, 30), z=70.0))())
    assert np.std(img_env[2]) > 0                                       # w' present in the rendered field
    env.close()


def test_brick_layout_is_bit_identical_to_the_compact_box(built_lib, monkeypatch):
    """Large turbulence boxes are re-laid as 64-byte bricks (8 trilinear corners per cell, wg_set_turbulence) so that a
    wake-centre sample is one aligned read: same values, same interpolation order -> the same bits as gathering the 8
    corners from the caller's compact layout.  Forced on the small test box here."""
    import torch
    from windgym_b200 import V80, VecWindFarmEnv
    _, box = _boxes()
    cfg = small_config(2, 2, reward="Power_avg", action="wind")
    B, T, steps = 8, 4, 12
    rng = np.random.default_rng(5)
    ws, ti, wd = rng.uniform(8, 13, B), rng.uniform(0.05, 0.12, B), rng.uniform(262, 278, B)
    yaw0 = rng.uniform(-15, 15, (B, T))
    off = rng.uniform(0, 1, (B, 3)) * (np.array(N) * np.array(D3))
    off[0] = (np.array(N) * np.array(D3)) - 1e-3          # an env at the periodic seam of the box
    acts = rng.uniform(-1, 1, (steps, B, T)).astype(np.float32)

    def run():
        env = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", turbtype="MannFixed", turb_box=box)
        env.reset(wind=(ws, ti, wd), yaw0=yaw0, turb_offset=off)
        out = []
        for a in acts:
            o, r, _, _, info = env.step(torch.as_tensor(a))
            out.append((o.cpu().numpy().copy(), info["Power pr turbine agent"].cpu().numpy().copy(),
                        env.state["pmut"].cpu().numpy().copy()))
        env.check_flags(); env.close()
        return out
    monkeypatch.setenv("WG_NO_BRICKS", "1")
    compact = run()
    monkeypatch.delenv("WG_NO_BRICKS")
    monkeypatch.setenv("WG_FORCE_BRICKS", "1")
    bricks = run()
    for (o1, p1, s1), (o2, p2, s2) in zip(compact, bricks):
        assert np.array_equal(o1, o2) and np.array_equal(p1, p2) and np.array_equal(s1, s2)


def test_turbtype_random_white_noise_box_vs_oracle(built_lib):
    """turbtype "Random" (Wind_Farm_Env.py:640-644: RandomTurbulence(ti, ws, seed) + the non-synchronised isotropic
    added-turbulence model): a box of independent N(0, 1) cells through the same sampling path; against the oracle on
    the same box.  The fluctuations reach the rotors, the wakes barely meander (white noise has no large scales)."""
    import torch
    from windgym_b200 import V80, VecWindFarmEnv
    from windgym_b200.mann import MannBox
    Nw, Dw = (128, 32, 16), (8.0, 8.0, 8.0)
    box = MannBox.white_noise(Nxyz=Nw, dxyz=Dw, seed=3, device="cuda:0")
    ref_box = box.raw[..., :3].permute(3, 0, 1, 2).cpu().numpy().astype(np.float64)
    assert abs(box.std_u - 1.0) < 0.01
    lp_std = float(box.lp.std())
    assert lp_std < 0.2                                      # the low-pass (meandering) part of white noise is small
    cfg = small_config(2, 2, reward="Power_avg", action="wind")
    B, T, steps = 2, 4, 6
    rng = np.random.default_rng(8)
    ws, ti, wd = rng.uniform(8, 12, B), rng.uniform(0.05, 0.10, B), rng.uniform(264, 276, B)
    yaw0 = rng.uniform(-10, 10, (B, T))
    off = rng.uniform(0, 1, (B, 3)) * (np.array(Nw) * np.array(Dw))
    acts = rng.uniform(-1, 1, (steps, B, T)).astype(np.float32)
    env = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", turbtype="Random", turb_box=box, added_turbulence=False)
    env.reset(wind=(ws, ti, wd), yaw0=yaw0, turb_offset=off)
    pw, uvw = [], []
    for a in acts:
        o, r, _, _, info = env.step(torch.as_tensor(a))
        pw.append(info["Power pr turbine agent"].cpu().numpy().copy())
        uvw.append(torch.stack([env.state[k][:, 0] for k in ("u", "v", "w")], -1).cpu().numpy().copy())
    env.check_flags(); env.close()
    pw, uvw = np.array(pw), np.array(uvw)
    assert np.abs(uvw[..., 2]).max() > 1e-3                   # w' reaches the rotors
    for b in range(B):
        field = mn.MannTurbulenceField(ref_box, Dw, lowpass_width=160.0)
        ref = oracle_rollout(cfg, ws[b:b + 1], ti[b:b + 1], wd[b:b + 1], yaw0[b:b + 1], acts[:, b:b + 1],
                             turbtype="Random", turb_field=field, reset_kw=dict(turb_offset=off[b]))
        rel = np.abs(pw[:, b] - ref["power"][0]) / np.maximum(ref["power"][0], 1.0)
        assert rel.max() < 1e-4, f"env {b}: power rel err {rel.max():.3e}"
    # the facade builds its own white-noise box for turbtype="Random" and consumes the TF_seed draw (:642)
    from windgym_b200 import WindFarmEnv
    fac = WindFarmEnv(V80(), config=small_config(2, 1, reward="Power_avg", action="yaw"), turbtype="Random", seed=4)
    fac.reset(seed=4)
    p = [fac.step(np.zeros(2, dtype=np.float32))[4]["Power agent"] for _ in range(10)]
    assert np.isfinite(p).all() and np.std(p) > 0
    fac.close()


def test_mannload_reads_netcdf4_boxes(built_lib, tmp_path):
    """turbtype "MannLoad" (Wind_Farm_Env.py:611-618): MannTurbulenceField.from_netcdf on the files under TurbBox --
    NetCDF-4 / HDF5 files read by the built-in reader (no netCDF4 / h5py in this image); FarmEval.update_tf
    (FarmEval.py:86-90) switches the file."""
    import torch
    from tests import hdf5_writer
    from windgym_b200 import V80, FarmEval, WindFarmEnv
    from windgym_b200.mann import MannBox
    rng = np.random.default_rng(4)
    shape, d3 = (64, 16, 8), (6.0, 12.0, 12.0)
    axes = [np.arange(n) * d for n, d in zip(shape, d3)]
    boxes = []
    for k in range(2):
        uvw = rng.normal(size=(3,) + shape).astype(np.float32)
        p = str(tmp_path / f"TF_{k}.nc")
        hdf5_writer.write(p, {"uvw": uvw, "x": axes[0], "y": axes[1], "z": axes[2]}, chunked=("uvw",) if k else ())
        boxes.append((p, uvw))
    box = MannBox.from_file(boxes[1][0], device="cuda:0")
    assert box.Nxyz == shape and box.dxyz == d3
    assert np.array_equal(box.raw[..., :3].permute(3, 0, 1, 2).cpu().numpy(), boxes[1][1])
    cfg = small_config(2, 1, reward="Power_avg", action="yaw")
    env = WindFarmEnv(V80(), config=cfg, turbtype="MannLoad", TurbBox=str(tmp_path), seed=3, device="cuda:0")
    assert env.vec.turb_box.Nxyz == shape
    env.reset(seed=3)
    p = [env.step(np.zeros(2, dtype=np.float32))[4]["Power agent"] for _ in range(8)]
    assert np.isfinite(p).all() and np.std(p) > 0
    env.close()
    fe = FarmEval(V80(), config=cfg, turbtype="MannLoad", TurbBox=boxes[0][0], yaw_init="Zeros", seed=1, device="cuda:0")
    fe.update_tf(boxes[1][0])
    assert np.array_equal(fe.vec.turb_box.raw[..., :3].permute(3, 0, 1, 2).cpu().numpy(), boxes[1][1])
    fe.set_wind_vals(ws=10, ti=0.08, wd=270)
    fe.reset()
    assert np.isfinite(fe.step(np.zeros(2, dtype=np.float32))[1])
    fe.close()

"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): per-step power within 1e-4 relative of the fp64 oracle; observations are
float32 in the reference, compared at 2e-5 absolute on the [-1, 1] scale; integer bookkeeping bit-exact.
"""
import numpy as np
import pytest

from tests.helpers import oracle_rollout, rich_config, small_config

pytestmark = pytest.mark.gpu

POWER_RTOL = 1e-4
OBS_ATOL = 2e-5


def _conditions(B, T, seed, wd_lo=258.0, wd_hi=282.0):
    rng = np.random.default_rng(seed)
    return (rng.uniform(7, 14, B), rng.uniform(0.02, 0.15, B), rng.uniform(wd_lo, wd_hi, B),
            rng.uniform(-15, 15, (B, T)))


def _run_gpu(cfg, ws, ti, wd, yaw0, acts, **kw):
    import torch
    from windgym_b200 import V80, VecWindFarmEnv
    B, T = yaw0.shape
    env = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", **kw)
    obs0, _ = env.reset(wind=(ws, ti, wd), yaw0=yaw0)
    out = dict(obs0=obs0.cpu().numpy().copy(), obs=[], reward=[], power=[], yaw=[], trunc=[], power_base=[], yaw_base=[])
    for a in acts:
        o, r, _, tr, info = env.step(torch.as_tensor(a))
        out["obs"].append(o.cpu().numpy().copy()); out["reward"].append(r.cpu().numpy().copy())
        out["trunc"].append(tr.cpu().numpy().copy())
        out["power"].append(info["Power pr turbine agent"].cpu().numpy().copy())
        out["yaw"].append(info["yaw angles agent"].cpu().numpy().copy())
        if env.Baseline_comp:
            out["power_base"].append(info["Power pr turbine baseline"].cpu().numpy().copy())
            out["yaw_base"].append(info["yaw angles base"].cpu().numpy().copy())
    env.check_flags()
    out = {k: np.array(v) for k, v in out.items()}
    return env, out


def _compare(gpu, ref, B, baseline=False):
    # gpu arrays are [steps, B, ...]; oracle arrays are [B, steps, ...]
    for b in range(B):
        p_g, p_r = gpu["power"][:, b], ref["power"][b]
        rel = np.abs(p_g - p_r) / np.maximum(np.abs(p_r), 1.0)
        assert rel.max() < POWER_RTOL, f"env {b}: power rel err {rel.max():.3e}"
        assert np.allclose(gpu["yaw"][:, b], ref["yaw"][b], atol=1e-4), f"env {b}: yaw"
        assert np.allclose(gpu["obs0"][b], ref["obs0"][b], atol=OBS_ATOL), f"env {b}: reset obs"
        assert np.allclose(gpu["obs"][:, b], ref["obs"][b], atol=OBS_ATOL), \
            f"env {b}: obs max err {np.abs(gpu['obs'][:, b] - ref['obs'][b]).max():.3e}"
        assert np.allclose(gpu["reward"][:, b], ref["reward"][b], rtol=2e-4, atol=2e-5), f"env {b}: reward"
        assert np.array_equal(gpu["trunc"][:, b].astype(bool), ref["trunc"][b].astype(bool)), f"env {b}: truncated"
        if baseline:
            relb = np.abs(gpu["power_base"][:, b] - ref["power_base"][b]) / np.maximum(ref["power_base"][b], 1.0)
            assert relb.max() < POWER_RTOL, f"env {b}: baseline power rel err {relb.max():.3e}"
            assert np.allclose(gpu["yaw_base"][:, b], ref["yaw_base"][b], atol=1e-3), f"env {b}: baseline yaw"


@pytest.mark.parametrize("nx,ny,action", [(2, 1, "yaw"), (2, 2, "wind"), (3, 2, "wind")])
def test_step_parity_power_avg(built_lib, nx, ny, action):
    T, B, steps = nx * ny, 3, 8
    cfg = small_config(nx, ny, reward="Power_avg", action=action)
    ws, ti, wd, yaw0 = _conditions(B, T, seed=10 + T)
    acts = np.random.default_rng(1).uniform(-1, 1, (steps, B, T)).astype(np.float32)
    _, gpu = _run_gpu(cfg, ws, ti, wd, yaw0, acts)
    ref = oracle_rollout(cfg, ws, ti, wd, yaw0, acts)
    _compare(gpu, ref, B)


def test_step_parity_baseline_farm(built_lib):
    """Power_reward 'Baseline': second farm with the greedy local controller (Wind_Farm_Env.py:948-954)."""
    nx, ny, B, steps = 2, 2, 3, 8
    cfg = small_config(nx, ny, reward="Baseline", action="wind")
    ws, ti, wd, yaw0 = _conditions(B, nx * ny, seed=3)
    acts = np.random.default_rng(2).uniform(-1, 1, (steps, B, nx * ny)).astype(np.float32)
    env, gpu = _run_gpu(cfg, ws, ti, wd, yaw0, acts)
    assert env.n_farms == 2
    ref = oracle_rollout(cfg, ws, ti, wd, yaw0, acts)
    _compare(gpu, ref, B, baseline=True)


def test_step_parity_substeps_and_penalty(built_lib):
    """dt_env = 3*dt_sim (substep means, Wind_Farm_Env.py:943-969) + 'Change' action penalty (:804-820)."""
    cfg = small_config(2, 2, reward="Power_avg", action="yaw", **{"act_pen.action_penalty": 0.05})
    B, T, steps = 2, 4, 5
    ws, ti, wd, yaw0 = _conditions(B, T, seed=5)
    acts = np.random.default_rng(3).uniform(-1, 1, (steps, B, T)).astype(np.float32)
    _, gpu = _run_gpu(cfg, ws, ti, wd, yaw0, acts, dt_env=3, dt_sim=1)
    ref = oracle_rollout(cfg, ws, ti, wd, yaw0, acts, dt_env=3, dt_sim=1)
    _compare(gpu, ref, B)


def test_induction_action_extension_parity(built_lib):
    """act_var = 2 (extension; BASELINE.json cfg 4 'yaw + induction actions'): [yaw | induction] actions, derated
    rotors a = delta a_tab -- against the oracle's turbine model; delta = 1 reproduces the yaw-only run exactly."""
    import torch
    from windgym_b200 import V80, VecWindFarmEnv
    cfg = small_config(2, 2, reward="Baseline", action="wind")
    B, T, steps = 3, 4, 8
    ws, ti, wd, yaw0 = _conditions(B, T, seed=31)
    rng = np.random.default_rng(6)
    acts = rng.uniform(-1, 1, (steps, B, 2 * T)).astype(np.float32)
    env, gpu = _run_gpu(cfg, ws, ti, wd, yaw0, acts, induction_control=True, derate_min=0.4)
    assert env.ec.act_var == 2
    der = env.state["derate"][:, 0].cpu().numpy()
    assert np.allclose(der, 0.4 + 0.5 * (acts[-1][:, T:] + 1.0) * 0.6, atol=1e-6) and (env.state["derate"][:, 1] == 1).all()
    ref = oracle_rollout(cfg, ws, ti, wd, yaw0, acts, induction_control=True, derate_min=0.4)
    _compare(gpu, ref, B, baseline=True)
    # derating costs power at the derated rotor: compare with the same yaw actions and delta = 1
    full = acts.copy()
    full[:, :, T:] = 1.0
    _, gpu1 = _run_gpu(cfg, ws, ti, wd, yaw0, full, induction_control=True, derate_min=0.4)
    _, gpu0 = _run_gpu(cfg, ws, ti, wd, yaw0, acts[:, :, :T])
    assert np.array_equal(gpu1["power"], gpu0["power"]) and np.array_equal(gpu1["obs"], gpu0["obs"])
    up = env.state["xr"].cpu().numpy().argmin(axis=1)
    assert (gpu["power"][-1][np.arange(B), up] <= gpu0["power"][-1][np.arange(B), up] + 1e-3).all()
    with pytest.raises(ValueError):
        env.step(torch.zeros((B, T)))


def test_rich_observation_parity(built_lib):
    """All measurement channels, current + several windows, TI, farm level (MesClass.py:328-351, :679-703)."""
    cfg = rich_config(2, 2, reward="Power_avg")
    B, T, steps = 2, 4, 6
    ws, ti, wd, yaw0 = _conditions(B, T, seed=8)
    acts = np.random.default_rng(4).uniform(-1, 1, (steps, B, T)).astype(np.float32)
    env, gpu = _run_gpu(cfg, ws, ti, wd, yaw0, acts)
    ref = oracle_rollout(cfg, ws, ti, wd, yaw0, acts)
    assert env.obs_var == ref["obs"].shape[-1]
    _compare(gpu, ref, B)


def test_flow_state_matches_oracle_particles(built_lib):
    """Raw wake state after fs.run(n): particle positions and Ainslie profiles, chain by chain, age by age."""
    from oracle import dwm_numpy as dwm
    from oracle.v80 import V80 as OV80
    from windgym_b200 import V80, VecWindFarmEnv
    from windgym_b200.config import grid_layout
    cfg = small_config(2, 2)
    B, T = 2, 4
    ws, ti, wd, yaw0 = _conditions(B, T, seed=21)
    env = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", fill_window=False)
    env.reset(wind=(ws, ti, wd), yaw0=yaw0)  # spin-up + 1 fill step
    n_spin, _, _ = env.ec.reset_integers(ws, wd)
    x, y = grid_layout(80.0, 4, 4, 2, 2)
    for b in range(B):
        wt = dwm.PyWakeWindTurbines(x, y, OV80())
        fs = dwm.DWMFlowSimulation(dwm.TurbulenceFieldSite(ws[b], dwm.RandomTurbulence(0, ws[b])), wt,
                                   wind_direction=wd[b], dt=1, d_particle=0.2)
        wt.yaw = yaw0[b]
        for _ in range(int(n_spin[b]) + 1):
            fs.step()
        assert int(env.state["n_step"][b, 0]) == fs.n_step
        for t in range(T):
            prof, pmut, pcon = env.profiles_by_age(b, 0, t)
            sl = fs.slots_by_age(t)
            assert prof.shape[0] == len(sl), (b, t, prof.shape[0], len(sl))
            assert np.allclose(pmut[:, :2], fs.pmut[t, sl, :2], atol=2e-2), "particle positions"
            assert np.abs(prof - fs.prof[t, sl]).max() < 2e-5, f"profiles max err {np.abs(prof - fs.prof[t, sl]).max():.2e}"
            assert np.allclose(pcon, fs.pcon[t, sl], atol=1e-4), "emission scalars"
            U = fs.prof[t, sl]
            M = np.sum((1.0 - U) * (np.arange(64) / 16.0)[None], axis=1) / 16.0
            bw = np.sqrt(np.maximum(2.0 * M * (1.0 - U.min(axis=1)), 0.0))
            assert np.allclose(env.last_bw, bw, atol=2e-5), "shear integral carried in the Dirichlet slot"
        assert np.allclose(env.state["u"][b, 0].cpu().numpy(), fs.rotor_avg_windspeed[:, 0], rtol=2e-5)


def _check_work_table(env, load_farm):
    """The work table of a single-substep step: every farm of the active envs appears with parts 0..n-1 exactly once,
    parts are ordered by descending tiles per part, unused entries are -1."""
    work = env.state["work"].cpu().numpy()
    used = work[(work[:, 0] >= 0) & ((work[:, 1] >> 8) > 0)]     # -1: unused entry; nparts 0: beyond the table's end
    farm, part, nparts = used[:, 0], used[:, 1] & 0xff, used[:, 1] >> 8
    U = load_farm.size
    assert set(farm.tolist()) == set(range(U)), "every farm of the active envs must be in the table"
    for u in range(U):
        m = farm == u
        n = int(nparts[m][0])
        assert (nparts[m] == n).all() and sorted(part[m].tolist()) == list(range(n)), f"farm {u}: parts {part[m]}"
    tiles = np.maximum(-(-load_farm // 32), 1)
    per_part = -(-tiles[farm] // nparts)
    assert np.all(np.diff(per_part) <= 0), "longest part first"
    return farm, part, nparts


def test_work_table_launch_order_and_retire_bookkeeping(built_lib, monkeypatch):
    """wg_step launches single-substep steps through a work table (state field `work`, wg_plan_kernel): with fewer
    farms than resident CTA slots the farms are cut into parts that fill the machine; any cut -- and the plain
    index order -- gives bit-identical results (fixed-point rotor sums).  Steps with substeps use the per-env launch
    order (state field `order`).  `retire` (stations the next step drops) never exceeds `count`."""
    import torch
    from windgym_b200 import V80, VecWindFarmEnv
    cfg = small_config(3, 2, reward="Power_avg", action="wind")
    B, T = 96, 6
    ws, ti, wd, yaw0 = _conditions(B, T, seed=11)
    acts = np.random.default_rng(5).uniform(-1, 1, (70, B, T)).astype(np.float32)   # > 64 steps: one periodic rebuild
    # the table is built from the loads the previous flow launch left behind: check it on the first step after a reset
    env_0 = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0")
    env_0.reset(wind=(ws, ti, wd), yaw0=yaw0)
    load0 = env_0.state["load"].cpu().numpy().reshape(-1)
    env_0.step(torch.as_tensor(acts[0]))
    farm, part, nparts = _check_work_table(env_0, load0)
    assert len(farm) > B and nparts.max() > 1, "96 farms leave CTA slots free: the heavy ones must be split"
    assert not env_0.state["part_acc"].any() and not env_0.state["part_keep"].any() and \
        not env_0.state["part_arrive"].any(), "the parts' scratch must be back at rest (all zeros) after the launch"
    env_0.close()
    # dt_env = 2: substeps inside the launch -> per-env order, no split
    env_s = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", dt_env=2)
    env_s.reset(wind=(ws, ti, wd), yaw0=yaw0)
    load_s = env_s.state["load"].cpu().numpy().sum(axis=1)
    env_s.step(torch.as_tensor(acts[0]))
    order = env_s.state["order"].cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(B)), "order must be a permutation of the active envs"
    quantum = max(1, -(-(env_s.n_farms * T * env_s.ec.p_cap) // 1024))
    key = np.minimum(load_s[order] // quantum, 1023)
    assert len(set(load_s.tolist())) > 8 and np.all(np.diff(key) <= 0), "envs must be launched by descending load class"
    env_s.close()
    env_a, out_a = _run_gpu(cfg, ws, ti, wd, yaw0, acts)
    _check_work_table(env_a, env_a.state["load"].cpu().numpy().reshape(-1) * 0 + 1)   # structure only (loads moved on)
    count, retire = env_a.state["count"].cpu().numpy(), env_a.state["retire"].cpu().numpy()
    assert (retire >= 0).all() and (retire <= count).all() and (retire <= 8).all()
    assert retire.sum() > 0 or count.sum() > 0
    env_a.close()
    monkeypatch.setenv("WG_NO_SPLIT", "1")          # read at wg_create: one CTA per farm, per-env order
    env_b, out_b = _run_gpu(cfg, ws, ti, wd, yaw0, acts)
    env_b.close()
    monkeypatch.setenv("WG_NO_ORDER", "1")          # ... and plain index order
    env_c, out_c = _run_gpu(cfg, ws, ti, wd, yaw0, acts)
    env_c.close()
    for k in ("obs", "reward", "power", "yaw", "trunc"):
        assert np.array_equal(out_a[k], out_b[k]), f"{k} depends on how the farms are cut into CTAs"
        assert np.array_equal(out_a[k], out_c[k]), f"{k} depends on the launch order"


@pytest.mark.parametrize("nx,ny,B", [(2, 2, 2), (4, 4, 1)])
def test_long_horizon_drift_1000_steps(built_lib, nx, ny, B):
    """The fp32 wake state is carried from step to step: 1000 steps of random actions (the bench horizon) against the
    fp64 oracle, Baseline reward (both farms).  The per-step power error must stay below 1e-4 for the whole rollout --
    stations live for about one farm transit (150-250 steps), so rounding differences cannot accumulate beyond that.
    The error profile over the rollout is written next to the bench artefacts (gpurun_out/drift_*.json)."""
    import json
    import os
    T, steps = nx * ny, 1000
    cfg = small_config(nx, ny, reward="Baseline", action="wind")
    ws, ti, wd, yaw0 = _conditions(B, T, seed=77 + T)
    acts = np.random.default_rng(9).uniform(-1, 1, (steps, B, T)).astype(np.float32)
    env, gpu = _run_gpu(cfg, ws, ti, wd, yaw0, acts, n_passthrough=60)
    env.close()
    ref = oracle_rollout(cfg, ws, ti, wd, yaw0, acts, n_passthrough=60)
    assert not ref["trunc"].any() and not gpu["trunc"].any()
    rel = np.abs(gpu["power"] - ref["power"].transpose(1, 0, 2)) / np.maximum(ref["power"].transpose(1, 0, 2), 1.0)
    relb = np.abs(gpu["power_base"] - ref["power_base"].transpose(1, 0, 2)) / np.maximum(ref["power_base"].transpose(1, 0, 2), 1.0)
    obs_err = np.abs(gpu["obs"] - ref["obs"].transpose(1, 0, 2))
    rew_err = np.abs(gpu["reward"] - ref["reward"].T)
    prof = {"farm": f"{nx}x{ny}", "envs": B, "steps": steps, "ws": ws.tolist(), "wd": wd.tolist(),
            "max_rel_power_err_per_100_steps": [float(rel[i:i + 100].max()) for i in range(0, steps, 100)],
            "max_rel_base_power_err_per_100_steps": [float(relb[i:i + 100].max()) for i in range(0, steps, 100)],
            "max_obs_err": float(obs_err.max()), "max_reward_err": float(rew_err.max())}
    print(json.dumps(prof))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"drift_{nx}x{ny}.json"), "w") as fh:
            json.dump(prof, fh, indent=1)
    assert rel.max() < POWER_RTOL and relb.max() < POWER_RTOL, prof
    assert obs_err.max() < OBS_ATOL and rew_err.max() < 2e-4, prof
    # no growth: the last 300 steps are no worse than 3x the first 300
    assert rel[-300:].max() < max(3.0 * rel[:300].max(), 2e-5), prof


def test_measurement_noise_normal_distribution_and_seeding(built_lib):
    """noise="Normal" (two of the three shipped YAMLs): farm_mes.add_measurements adds N(0, 2 deg) to the wind
    direction samples and nothing (sigma = 0) to ws / yaw / power (MesClass.py:436-444, :574-577).  The reference's
    noise RNG is unseeded, so what can be pinned is the distribution, which channels it touches, that
    it never reaches the flow, reproducibility under ``noise_seed`` and independence across envs."""
    import torch
    from windgym_b200 import V80, VecWindFarmEnv
    cfg = small_config(2, 2, reward="Power_avg", action="wind",
                       **{"wind.wd_min": 200, "wind.wd_max": 340})      # wide scaling range: the noisy wd never clips
    cfg["mes_level"].update(turb_ws=True, turb_wd=True, turb_power=True)
    for c in ("ws", "wd", "yaw", "power"):
        cfg[f"{c}_mes"].update({f"{c}_current": True, f"{c}_rolling_mean": False})
    B, T, steps = 256, 4, 30
    ws, ti, wd = np.full(B, 9.0), np.full(B, 0.06), np.full(B, 270.0)
    yaw0 = np.zeros((B, T))
    acts = np.random.default_rng(3).uniform(-1, 1, (steps, 1, T)).astype(np.float32).repeat(B, axis=1)

    def run(noise, seed):
        c = dict(cfg, noise=noise)
        env = VecWindFarmEnv(V80(), B, config=c, device="cuda:0", noise_seed=seed)
        env.reset(wind=(ws, ti, wd), yaw0=yaw0)
        obs, pw = [], []
        for a in acts:
            o, r, _, _, info = env.step(torch.as_tensor(a))
            obs.append(o.cpu().numpy().copy()); pw.append(info["Power pr turbine agent"].cpu().numpy().copy())
        env.check_flags(); env.close()
        return np.array(obs).reshape(steps, B, T, 4), np.array(pw)      # per turbine: ws | wd | yaw | power

    clean, p_clean = run("None", 0)
    noisy, p_noisy = run("Normal", 7)
    again, _ = run("Normal", 7)
    other, _ = run("Normal", 8)
    assert np.array_equal(p_clean, p_noisy), "measurement noise must not reach the flow"
    for ch, name in ((0, "ws"), (2, "yaw"), (3, "power")):
        assert np.array_equal(clean[..., ch], noisy[..., ch]), f"{name} observations must be noise-free (sigma = 0)"
    span = (340 + 5) - (200 - 5)                                        # wd scaling range (Wind_Farm_Env.py:443-444)
    noise_deg = (noisy[..., 1].astype(np.float64) - clean[..., 1]) * span / 2.0
    n = noise_deg.size
    sd, mean = noise_deg.std(), noise_deg.mean()
    z = (noise_deg - mean) / sd
    assert abs(sd - 2.0) < 0.04 and abs(mean) < 0.05, (sd, mean)        # 30720 samples: sigma known to ~0.4 %
    assert abs((z ** 3).mean()) < 0.06 and abs((z ** 4).mean() - 3.0) < 0.15, "not a normal distribution"
    assert np.array_equal(noisy, again), "same noise_seed must reproduce the observations"
    assert not np.array_equal(noisy[..., 1], other[..., 1]), "another noise_seed must give other noise"
    # identical envs draw independent noise: across-env correlation of the noise ~ 0, and no two envs share a sample
    a, b = noise_deg[:, 0].ravel(), noise_deg[:, 1].ravel()
    assert abs(np.corrcoef(a, b)[0, 1]) < 0.25 and len(np.unique(np.round(noise_deg[0], 6))) > 0.98 * B * T


@pytest.mark.parametrize("B,reward", [(200, "Baseline"), (1100, "Power_avg")])
def test_free_running_steps_overlap_the_previous_finish_kernel(built_lib, monkeypatch, B, reward):
    """Steps issued back to back without a host synchronisation: the flow kernel of step k+1 is launched as
    programmatic dependent of step k's finish kernel (its prologue and tile loop overlap it, griddepcontrol.wait in
    front of the turbine epilogue); the finish kernel in turn is a programmatic dependent of the step's flow kernel
    (released at CTA start below one wave of CTAs, behind the tile loops above).  Results -- every observation / reward
    of the rollout, the wake state at the end -- must be bit-identical to plain stream order (WG_NO_PDL=1), for a batch
    below one wave of CTAs and above it, with one and two farms per env."""
    import torch
    from windgym_b200 import V80, VecWindFarmEnv
    cfg = small_config(3, 2, reward=reward, action="yaw")
    T, steps = 6, 150
    ws, ti, wd, yaw0 = _conditions(B, T, seed=23)
    acts = torch.as_tensor(np.random.default_rng(8).uniform(-1, 1, (steps, B, T)).astype(np.float32)).cuda()

    def rollout():
        env = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", n_passthrough=50)
        env.reset(wind=(ws, ti, wd), yaw0=yaw0)
        obs = torch.empty((steps, B, env.obs_var), device="cuda:0")
        rew = torch.empty((steps, B), device="cuda:0")
        for k in range(steps):                      # no read-back inside the loop: the launches run ahead of the device
            o, r, _, _, _ = env.step(acts[k], info=False)
            obs[k].copy_(o); rew[k].copy_(r)
        torch.cuda.synchronize()
        env.check_flags()
        st = {k: env.state[k].clone() for k in ("prof", "pmut", "pcon", "yaw", "power", "count", "head", "rings")}
        env.close()
        return obs.cpu().numpy(), rew.cpu().numpy(), st

    obs_a, rew_a, st_a = rollout()
    monkeypatch.setenv("WG_NO_PDL_NEXT", "1")       # read at wg_create: no edge across steps
    obs_c, rew_c, st_c = rollout()
    monkeypatch.setenv("WG_NO_PDL", "1")            # no programmatic launch at all: plain stream order
    obs_b, rew_b, st_b = rollout()
    assert np.isfinite(obs_a).all() and np.abs(obs_a).max() > 0
    for o, r, st, what in ((obs_a, rew_a, st_a, "all edges"), (obs_c, rew_c, st_c, "flow -> finish edge only")):
        assert np.array_equal(o, obs_b) and np.array_equal(r, rew_b), what
        for k in st:
            assert torch.equal(st[k], st_b[k]), f"{what}: state field {k} differs"


def test_steps_captured_in_a_cuda_graph_replay_bit_identically(built_lib):
    """A sequence of steps (work-table rebuild, flow and finish kernels with their programmatic launch edges) can be
    captured in a CUDA graph on the caller's stream and replayed: no call synchronises, allocates or depends on host
    state that a replay would miss.  Two replays with fresh actions equal the same 2 x 24 steps issued eagerly."""
    import torch
    from windgym_b200 import V80, VecWindFarmEnv
    cfg = small_config(3, 2, reward="Power_avg", action="yaw")
    B, T, n = 300, 6, 24
    ws, ti, wd, yaw0 = _conditions(B, T, seed=31)
    acts = torch.as_tensor(np.random.default_rng(3).uniform(-1, 1, (2, n, B, T)).astype(np.float32)).cuda()

    env_e = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", n_passthrough=50)
    env_e.reset(wind=(ws, ti, wd), yaw0=yaw0)
    obs_e = torch.empty((2, n, B, env_e.obs_var), device="cuda:0")
    rew_e = torch.empty((2, n, B), device="cuda:0")
    for r in range(2):
        for k in range(n):
            o, rw, _, _, _ = env_e.step(acts[r, k], info=False)
            obs_e[r, k].copy_(o); rew_e[r, k].copy_(rw)
    torch.cuda.synchronize()

    env_g = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0", n_passthrough=50)
    env_g.reset(wind=(ws, ti, wd), yaw0=yaw0)
    a_static = torch.zeros((n, B, T), device="cuda:0")
    obs_g = torch.empty((n, B, env_g.obs_var), device="cuda:0")
    rew_g = torch.empty((n, B), device="cuda:0")
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for k in range(n):
            o, rw, _, _, _ = env_g.step(a_static[k], info=False)
            obs_g[k].copy_(o); rew_g[k].copy_(rw)
    for r in range(2):
        a_static.copy_(acts[r])
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(obs_g, obs_e[r]) and torch.equal(rew_g, rew_e[r]), f"replay {r}"
    env_g.check_flags()
    for k in ("prof", "pmut", "yaw", "power", "count", "rings"):
        assert torch.equal(env_g.state[k], env_e.state[k]), f"state field {k}"
    env_e.close(); env_g.close()

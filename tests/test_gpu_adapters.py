"""GPU: the f-2 / f-3 rows on the CUDA backend -- vector adapters with masked auto-reset, the batched evaluation
against the single-env FarmEval facade loop (the reference's eval_single_fast pattern) and the oracle, an SB3-style
MLP policy rolled out on the device."""
import numpy as np
import pytest

from tests.helpers import oracle_rollout, small_config

pytestmark = pytest.mark.gpu


def test_vector_env_autoreset_matches_fresh_reset(built_lib):
    import torch
    from windgym_b200 import GymVectorEnv, RecordEpisodeVals, V80, VecWindFarmEnv
    cfg = small_config(2, 1, reward="Power_avg", action="yaw")
    B, T = 4, 2
    # n_passthrough tiny -> time_max = int(dist/ws * n_pass) is a handful of steps, different per env
    # pooled=False: the masked in-step reset (exact reference RNG stream); the default is the device-side pool
    env = RecordEpisodeVals(GymVectorEnv(V80(), B, config=cfg, n_passthrough=0.2, seed=5, device="cuda:0", pooled=False))
    obs, infos = env.reset(seed=5)
    v = env.env.venv
    tmax = v.time_max.copy()
    assert obs.shape == (B, 4) and obs.dtype == np.float32 and (tmax > 0).all() and len(set(tmax.tolist())) > 1
    first_done = None
    for k in range(int(tmax.max()) + 3):
        ws_before = v.ws.copy()
        obs, r, term, trunc, infos = env.step(np.zeros((B, T), dtype=np.float32))
        assert np.isfinite(obs).all() and np.abs(obs).max() <= 1.0 and not term.any()
        assert np.array_equal(trunc, k >= tmax) or first_done is not None
        if trunc.any() and first_done is None:
            first_done = (k, trunc.copy())
            assert np.array_equal(infos["_final_observation"], trunc)
            ts = v.state["timestep"].cpu().numpy()
            assert (ts[trunc] == 0).all() and (ts[~trunc] == k + 1).all()
            # finished envs drew new conditions, the others kept theirs
            assert (v.ws[trunc] != ws_before[trunc]).all() and (v.ws[~trunc] == ws_before[~trunc]).all()
            # the auto-reset observation equals a fresh env reset onto the same conditions
            fresh = VecWindFarmEnv(V80(), B, config=cfg, n_passthrough=0.2, device="cuda:0")
            yaw_now = v.state["yaw"][:, 0].cpu().numpy()
            f_obs, _ = fresh.reset(wind=(v.ws, v.ti, v.wd), yaw0=yaw_now)
            assert np.array_equal(f_obs.cpu().numpy()[trunc], obs[trunc])
            fresh.close()
    assert first_done is not None and len(env.mean_power_queue) >= B
    assert all(p > 0 for p in env.mean_power_queue) and all(l >= 1 for l in env.length_queue)
    v.check_flags()


def test_eval_batched_vs_farmeval_loop_and_oracle(built_lib):
    import torch
    from windgym_b200 import ConstantAgent, FarmEval, V80, VecWindFarmEnv, eval_batched
    cfg = small_config(2, 2, reward="Baseline", action="wind")
    wss, wds, tis, t_sim = [9.0, 12.0], [265.0, 275.0], [0.07], 12
    agent = ConstantAgent([-10, 20, 0, 0])
    env = VecWindFarmEnv(V80(), 4, config=cfg, eval_mode=True, yaw_init="Defined", device="cuda:0")
    ds = eval_batched(env, agent, wss, wds, tis, t_sim=t_sim)
    env.close()
    assert ds["powerT_a"].shape == (t_sim, 4, 2, 2, 1, 1, 1) and "pct_inc" in ds
    # yaw moves one degree per step towards the target (SURVEY.md 8a known answer), baseline farm stays greedy
    assert np.allclose(ds["yaw_a"][5, :, 0, 0, 0, 0, 0], [-5, 5, 0, 0]) and np.allclose(ds["yaw_a"][11, :, 1, 1, 0, 0, 0], [-10, 11, 0, 0])
    for i, ws in enumerate(wss):
        for j, wd in enumerate(wds):
            # --- the reference's serial pattern on the single-env facade: bit-identical (same CUDA path, batch of one)
            fe = FarmEval(V80(), config=cfg, Baseline_comp=True, yaw_init="Defined", reset_init=False, device="cuda:0")
            fe.set_wind_vals(ws=ws, ti=0.07, wd=wd)
            fe.set_yaw_vals(0.0)
            obs, _ = fe.reset()
            pw = [fe.fs.windTurbines.power()]
            pb = [fe.fs_baseline.windTurbines.power()]
            agent.env = fe
            for _ in range(1, t_sim):
                obs, r, _, _, _ = fe.step(agent.predict(obs)[0])
                pw.append(fe.fs.windTurbines.power()); pb.append(fe.fs_baseline.windTurbines.power())
            assert fe.fs.time == ds.coords["time0"][i, j, 0, 0] + (t_sim - 1)
            fe.close()
            assert np.array_equal(ds["powerT_a"][:, :, i, j, 0, 0, 0], np.array(pw))
            assert np.array_equal(ds["powerT_b"][:, :, i, j, 0, 0, 0], np.array(pb))
    # --- one condition against the CPU oracle
    acts = np.tile(agent.predict()[0].astype(np.float32), (t_sim - 1, 1, 1))
    ref = oracle_rollout(cfg, np.array([12.0]), np.array([0.07]), np.array([265.0]), np.zeros((1, 4)), acts,
                         eval_mode=True)
    rel = np.abs(ds["powerT_a"][1:, :, 1, 0, 0, 0, 0] - ref["power"][0]) / np.maximum(ref["power"][0], 1.0)
    assert rel.max() < 1e-4
    relb = np.abs(ds["powerT_b"][1:, :, 1, 0, 0, 0, 0] - ref["power_base"][0]) / np.maximum(ref["power_base"][0], 1.0)
    assert relb.max() < 1e-4
    assert np.allclose(ds["reward"][1:, 1, 0, 0, 0, 0], ref["reward"][0], rtol=2e-4, atol=2e-5)


def test_sb3_mlp_policy_rollout_on_device(built_lib):
    import torch
    from windgym_b200 import GymVectorEnv, SB3MlpPolicy, V80
    cfg = small_config(2, 2, reward="Power_avg", action="yaw")
    env = GymVectorEnv(V80(), 8, config=cfg, device="cuda:0", as_torch=True, seed=0)
    torch.manual_seed(0)
    pol = SB3MlpPolicy(env.single_observation_space.shape[0], 4).to("cuda:0")   # PPO_2975000.zip shape: 8 -> 64 -> 64 -> 4
    obs, _ = env.reset(seed=0)
    assert obs.is_cuda and obs.shape == (8, 8)
    tot = 0.0
    for _ in range(5):
        act = pol.predict_batch(obs)
        assert act.is_cuda and act.shape == (8, 4) and float(act.abs().max()) <= 1.0
        obs, r, term, trunc, infos = env.step(act)
        assert r.is_cuda and infos["Power agent"].is_cuda
        tot += float(infos["Power agent"].sum())
    assert np.isfinite(tot) and tot > 0
    env.close()


def test_flow_field_and_render_vs_oracle(built_lib):
    """f-4: fs.get_windspeed on an XY view (the render path) against the oracle's superposition at points."""
    from oracle import dwm_numpy as dwm
    from oracle.v80 import V80 as OV80
    from windgym_b200 import FarmEval, V80
    from windgym_b200.config import grid_layout
    from windgym_b200.envs import XYView
    cfg = small_config(2, 2, reward="Power_avg", action="wind")
    ws, wd = 9.0, 268.0
    env = FarmEval(V80(), config=cfg, yaw_init="Defined", reset_init=False, render_mode="rgb_array", device="cuda:0")
    env.set_wind_vals(ws=ws, ti=0.07, wd=wd)
    env.set_yaw_vals([20.0, -15.0, 0.0, 10.0])
    env.reset()
    n_spin, _, _ = env.vec.ec.reset_integers(np.array([ws]), np.array([wd]))
    n_flow = int(n_spin[0]) + env.steps_on_reset
    assert env.fs.time == n_flow
    # oracle flow in the same state
    x, y = grid_layout(80.0, 4, 4, 2, 2)
    wt = dwm.PyWakeWindTurbines(x, y, OV80())
    fs = dwm.DWMFlowSimulation(dwm.TurbulenceFieldSite(ws, dwm.RandomTurbulence(0, ws)), wt, wind_direction=wd, dt=1,
                               d_particle=0.2)
    wt.yaw = [20.0, -15.0, 0.0, 10.0]
    for _ in range(n_flow):
        fs.step()
    xt, yt = env.fs.windTurbines.positions_xyz[:2]
    view = XYView(x=np.linspace(xt.min() - 150, xt.max() + 700, 61), y=np.linspace(yt.min() - 150, yt.max() + 150, 47), z=70.0)
    got = env.fs.get_windspeed(view, include_wakes=True, xarray=False)
    ref = fs.get_windspeed(view)
    assert got.shape == ref.shape == (3, 61, 47)
    assert ref[0].min() < 0.75 * ws and ref[0].max() <= ws + 1e-9      # there are wakes in the view
    assert np.abs(got[0] - ref[0]).max() < 2e-4 * ws, np.abs(got[0] - ref[0]).max()
    assert np.abs(got[1] - ref[1]).max() < 2e-4 * ws
    assert np.abs(got[1]).max() > 0.05                                 # yawed rotors deflect: lateral component
    free = env.fs.get_windspeed(view, include_wakes=False)
    assert np.all(free[0] == ws) and np.all(free[1:] == 0)
    fa = env.fs.get_windspeed(view, xarray=True)
    assert np.array_equal(fa[0], got[0]) and np.array_equal(fa.x.values, view.x)
    # render: RGB frame of the reference's 250 x 250 view
    img = env.render()
    assert img.shape == (250, 250, 3) and img.dtype == np.uint8 and len(np.unique(img.reshape(-1, 3), axis=0)) > 20
    assert (img.reshape(-1, 3).sum(1) == 0).sum() >= 4 * 20               # rotors drawn
    env.close()


def test_pooled_autoreset_swaps_in_predeveloped_envs(built_lib):
    """Auto-reset through the spare pool: a swapped-in env is exactly a freshly reset env on its conditions, the
    pool recycles, and a dry pool falls back to the synchronous masked reset."""
    import torch
    from windgym_b200 import GymVectorEnv, PooledVecEnv, V80, VecWindFarmEnv
    cfg = small_config(2, 1, reward="Power_avg", action="yaw")
    B, T = 6, 2
    pool = PooledVecEnv(V80(), B, reserve=4, refill_chunk=2, config=cfg, n_passthrough=0.2, seed=9, device="cuda:0")
    env = GymVectorEnv(venv=pool)
    obs, infos = env.reset(seed=9)
    assert obs.shape == (B, 4) and pool.state["yaw"].shape[0] == B and pool.inner.state["yaw"].shape[0] == B + 4
    checked = 0
    for k in range(60):
        ws_before = pool.ws.copy()
        obs, r, term, trunc, infos = env.step(np.zeros((B, T), dtype=np.float32))
        assert obs.shape == (B, 4) and np.isfinite(obs).all() and infos["Power agent"].shape == (B,)
        if trunc.any():
            ts = pool.state["timestep"].cpu().numpy()
            assert (ts[trunc] == 0).all()
            assert (pool.ws[trunc] != ws_before[trunc]).all() and (pool.ws[~trunc] == ws_before[~trunc]).all()
            if checked < 3:   # the swapped-in state is a genuine reset onto the env's (new) conditions
                fresh = VecWindFarmEnv(V80(), B, config=cfg, n_passthrough=0.2, device="cuda:0")
                f_obs, _ = fresh.reset(wind=(pool.ws, pool.ti, pool.wd), yaw0=pool.state["yaw"][:, 0].cpu().numpy())
                assert np.array_equal(f_obs.cpu().numpy()[trunc], obs[trunc])
                for key in ("count", "head", "n_step", "power"):
                    a, b = fresh.state[key].cpu().numpy(), pool.state[key].cpu().numpy()
                    assert np.array_equal(a[trunc], b[trunc]), key
                assert np.array_equal(fresh.time_max[trunc], pool.time_max[trunc])
                fresh.close()
                checked += 1
    pool.check_flags()
    assert checked == 3 and pool.stats["swapped"] >= 10 and pool.stats["refills"] >= 3
    pool.close()
    # a pool of one spare runs dry when several envs finish together: the rest is reset synchronously
    tiny = PooledVecEnv(V80(), 4, reserve=1, config=cfg, n_passthrough=0.2, seed=1, device="cuda:0")
    tiny.reset(seed=1)
    tiny.reset(mask=np.array([True, True, True, False]))
    assert tiny.stats["swapped"] == 1 and tiny.stats["sync_resets"] == 2
    assert (tiny.state["timestep"].cpu().numpy() == 0).all() and np.isfinite(tiny.obs.cpu().numpy()).all()
    tiny.close()


def test_step_host_equals_step(built_lib):
    """Host-buffer step (numpy in, numpy out, one packed D2H copy) gives the same results as the device step."""
    import torch
    from windgym_b200 import V80, VecWindFarmEnv
    cfg = small_config(2, 2, reward="Power_avg", action="wind")
    B, T = 6, 4
    rng = np.random.default_rng(3)
    ws = rng.uniform(7, 12, B); wd = rng.uniform(262, 278, B); ti = np.full(B, 0.07); yaw0 = rng.uniform(-15, 15, (B, T))
    acts = rng.uniform(-1, 1, (5, B, T)).astype(np.float32)
    a = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0")
    b = VecWindFarmEnv(V80(), B, config=cfg, device="cuda:0")
    a.reset(wind=(ws, ti, wd), yaw0=yaw0)
    b.reset(wind=(ws, ti, wd), yaw0=yaw0)
    pinned = torch.from_numpy(acts).pin_memory()
    for k in range(5):
        o, r, _, tr, _ = a.step(torch.as_tensor(acts[k]))
        # numpy array (staged into the env's pinned buffer), caller's pinned tensor (zero-copy as it is), pageable
        # tensor (copy-engine fallback inside wg_step_host): all three give the device step's bits
        src = (acts[k], pinned[k], torch.from_numpy(acts[k].copy()))[k % 3]
        oh, rh, th = b.step_host(src)
        assert np.array_equal(b.obs.cpu().numpy(), oh)   # the device-side result buffer is written as well
        assert oh.dtype == np.float32 and rh.dtype == np.float32 and th.dtype == np.bool_
        assert np.array_equal(oh, o.cpu().numpy()) and np.array_equal(rh, r.cpu().numpy())
        assert np.array_equal(th, tr.cpu().numpy().astype(bool))
    with pytest.raises(ValueError):
        b.step_host(np.zeros((B, T + 1), dtype=np.float32))
    a.close(); b.close()


@pytest.mark.parametrize("multi", [False, True])
def test_device_pool_autoreset_without_the_host(built_lib, multi):
    """DevicePooledVecEnv (wg_pool_*): finished episodes are paired with ready spares and replaced on the device; a
    swapped-in env is exactly a freshly reset env on the conditions the device drew for it; the finished episode's
    last observation is kept; nothing is read back by the adapter.  Single- and multi-agent observation layouts."""
    import torch
    from windgym_b200 import DevicePooledVecEnv, GymVectorEnv, V80, VecWindFarmEnv
    cfg = small_config(2, 1, reward="Power_avg", action="yaw")
    B, T, R = 6, 2, 6
    env = GymVectorEnv(V80(), B, config=cfg, n_passthrough=0.2, seed=9, device="cuda:0", as_torch=True, reserve=R,
                       refill_every=2, multi_agent=multi)
    pool = env.venv
    assert isinstance(pool, DevicePooledVecEnv) and pool.inner.n_envs == B + R
    obs, infos = env.reset(seed=9)
    oshape = (B, T, 2) if multi else (B, 4)
    assert tuple(obs.shape) == oshape and tuple(env.single_observation_space.shape) == oshape[1:]
    torch.cuda.synchronize()
    assert (pool.inner.state["pool_status"][:B] == 0).all()
    zeros = torch.zeros((B, T), dtype=torch.float32, device="cuda:0")
    checked, n_swapped, prev_obs = 0, 0, obs.clone()
    for k in range(70):
        ws_before = pool.ws.clone()
        ts_before = pool.state["timestep"].clone()
        obs, r, term, trunc, infos = env.step(zeros)
        torch.cuda.synchronize()
        tr = trunc.cpu().numpy().astype(bool)
        assert tuple(obs.shape) == oshape and bool(torch.isfinite(obs).all()) and not bool(term.any())
        ts = pool.state["timestep"].cpu().numpy()
        assert (ts[tr] == 0).all() and (ts[~tr] == ts_before.cpu().numpy()[~tr] + 1).all()
        if tr.any():
            n_swapped += int(tr.sum())
            wsn, wsb = pool.ws.cpu().numpy(), ws_before.cpu().numpy()
            assert (wsn[tr] != wsb[tr]).all() and (wsn[~tr] == wsb[~tr]).all()
            assert (wsn >= 7).all() and (wsn <= 15).all()
            fo = infos["final_observation"]
            assert np.array_equal(infos["_final_observation"].cpu().numpy(), tr)
            assert bool(torch.isfinite(fo[trunc.bool()]).all()) and not torch.equal(fo[trunc.bool()], obs[trunc.bool()])
            if checked < 3:   # the swapped-in state is a genuine reset onto the conditions the device drew
                fresh = VecWindFarmEnv(V80(), B, config=cfg, n_passthrough=0.2, device="cuda:0", multi_agent=multi)
                f_obs, _ = fresh.reset(wind=(pool.ws.cpu().numpy().astype(np.float64), pool.ti.cpu().numpy().astype(np.float64),
                                             pool.wd.cpu().numpy().astype(np.float64)),
                                       yaw0=pool.state["yaw"][:, 0].cpu().numpy())
                assert np.array_equal(f_obs.cpu().numpy()[tr], obs.cpu().numpy()[tr])
                for key in ("count", "head", "n_step", "power", "time_max"):
                    a, b = fresh.state[key].cpu().numpy(), pool.state[key].cpu().numpy()
                    assert np.array_equal(a[tr], b[tr]), key
                fresh.close()
                checked += 1
        prev_obs = obs.clone()
    pool.check_flags()
    st = pool.stats
    assert checked == 3 and st["swapped"] == n_swapped >= 10 and st["refilled"] >= st["swapped"] and st["refill_calls"] >= 5
    status = pool.inner.state["pool_status"].cpu().numpy()
    assert (status[:B] == 0).all() and set(status[B:].tolist()) <= {1, 2, 3, 4}
    env.close()


def test_two_gpus_driven_by_one_process(built_lib):
    """One process, one handle per GPU.  The opt-in to > 48 KB of dynamic shared memory is a
    per-DEVICE function attribute: the 64-turbine-capacity flow kernel variant (T > 16) and a finish kernel with long
    histories need it on every device they run on.  Same inputs -> the same bits on both GPUs."""
    import torch
    from windgym_b200 import V80, VecWindFarmEnv
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run through `gpurun --gpus 2`)")
    cfg = small_config(5, 4, reward="Power_avg", action="wind",
                       **{"ws_mes.ws_history_length": 400, "ws_mes.ws_window_length": 400})   # 20 turbines, 33 KB of rings
    B, T = 8, 20
    rng = np.random.default_rng(2)
    ws, ti, wd = rng.uniform(9, 13, B), rng.uniform(0.04, 0.1, B), rng.uniform(262, 278, B)
    yaw0 = rng.uniform(-10, 10, (B, T))
    acts = rng.uniform(-1, 1, (4, B, T)).astype(np.float32)
    outs = []
    envs = [VecWindFarmEnv(V80(), B, config=cfg, device=f"cuda:{i}", fill_window=3) for i in (0, 1)]
    for env in envs:
        with torch.cuda.device(env.device):
            env.reset(wind=(ws, ti, wd), yaw0=yaw0)
    for a in acts:                                   # interleaved stepping of the two devices
        for env in envs:
            with torch.cuda.device(env.device):
                env.step(torch.as_tensor(a))
    for env in envs:
        with torch.cuda.device(env.device):
            env.check_flags()
            outs.append((env.obs.cpu().numpy().copy(), env.state["power"].cpu().numpy().copy()))
            env.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.isfinite(outs[0][0]).all() and outs[0][1].max() > 1e5


def test_sb3_vec_env_and_ppo_rollout_on_the_device_pool(built_lib):
    """The adapters as a training loop uses them, on the default device-side pool: SB3VecEnv (numpy out, list of info
    dicts with terminal_observation / TimeLimit.truncated) and collect_rollout with per-agent observations (BASELINE.json
    cfg 5's PPO-rollout shape) through real episode ends -- no masked reset, no stall."""
    import torch
    from windgym_b200 import DevicePooledVecEnv, SB3VecEnv, V80
    from windgym_b200.vector import collect_rollout
    cfg = small_config(2, 1, reward="Power_avg", action="yaw")
    B, T = 5, 2
    env = SB3VecEnv(V80(), B, config=cfg, n_passthrough=0.2, seed=2, device="cuda:0", reserve=8, refill_every=1)
    assert isinstance(env.venv, DevicePooledVecEnv)
    obs = env.reset()
    assert obs.shape == (B, 4) and obs.dtype == np.float32
    n_done, saw_terminal = 0, False
    for k in range(40):
        obs, rew, dones, infos = env.step(np.zeros((B, T), dtype=np.float32))
        assert obs.shape == (B, 4) and rew.shape == (B,) and dones.dtype == bool and len(infos) == B
        assert np.isfinite(obs).all() and np.isfinite(rew).all()
        for i in range(B):
            assert infos[i]["TimeLimit.truncated"] == bool(dones[i])
            if dones[i]:
                n_done += 1
                saw_terminal = True
                assert infos[i]["terminal_observation"].shape == (4,) and np.isfinite(infos[i]["terminal_observation"]).all()
            assert infos[i]["Power agent"] > 0
    assert n_done >= 5 and saw_terminal
    env.close()
    # multi-agent rollout buffers on the pool: obs [n, B, T, obs], one action per agent, dones where episodes ended
    pool = DevicePooledVecEnv(V80(), 6, config=small_config(2, 2, reward="Power_avg", action="yaw"), multi_agent=True,
                              n_passthrough=0.3, seed=4, device="cuda:0", reserve=12, refill_every=1)
    pool.reset(seed=4)
    ro = collect_rollout(pool, lambda o: torch.tanh(o.sum(dim=-1)), 48)
    assert tuple(ro["obs"].shape) == (48, 6, 4, 2) and tuple(ro["actions"].shape) == (48, 6, 4)
    assert bool(torch.isfinite(ro["obs"]).all()) and bool(torch.isfinite(ro["rewards"]).all())
    assert int(ro["dones"].sum()) >= 6                      # every env finished at least about once
    st = pool.stats
    assert st["swapped"] == int(ro["dones"].sum()) and st["refilled"] >= st["swapped"]
    pool.check_flags()
    pool.close()

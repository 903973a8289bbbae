#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched WindFarmEnv.step hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # CPU arm: oracle port on the host cores

One "step" = one batched ``VecWindFarmEnv.step`` over all envs of the rank: yaw update, S flow substeps of every
farm (wake advection + Ainslie march + superposition + rotor average + turbine update), MesClass push/extract,
reward, truncation.  Workload = BASELINE.json configs[1]: 16-turbine 4x4 grid, 4096 envs, yaw-only actions,
Env1.yaml observation/yaw semantics, uniform inflow (turbtype "None"), synthetic U(-1,1) actions.  Under torchrun
the 4096 envs are SHARDED over the ranks (strong scaling: BASELINE.json configs[2] = 512 per GPU at N = 8).
Prints ONE JSON line on rank 0 (contract: see the task statement / DESIGN.md section "Measurement").
"""
import argparse
import json
import math
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/s (16-turbine farm, 4096 envs)"
UNIT = "env-steps/s"
STATION_BYTES = (2 * 64 + 8 + 4) * 4   # profile r+w, 8 scalars read, 4 written  = 560 B
TURB_BYTES = 8 * 4                     # per turbine and flow step
HBM_FALLBACK_GBS = 6650.0              # /opt/skills/guides/B200_PROFILING.md fallback


def workload_config(nx, ny, reward):
    """Env1.yaml semantics on an nx x ny grid (reference examples/EnvConfigs/Env1.yaml)."""
    return {
        "yaw_init": "Random", "noise": "None", "BaseController": "Local", "ActionMethod": "wind", "Track_power": False,
        "farm": {"yaw_min": -45, "yaw_max": 45, "xDist": 4, "yDist": 4, "nx": nx, "ny": ny},
        "wind": {"ws_min": 7, "ws_max": 15, "TI_min": 0.02, "TI_max": 0.15, "wd_min": 255, "wd_max": 285},
        "act_pen": {"action_penalty": 0.0, "action_penalty_type": "Change"},
        "power_def": {"Power_reward": reward, "Power_avg": 10, "Power_scaling": 1.0},
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


def sample_conditions(cfg, env_ids, T, seed0=0):
    """Per-env (ws, ti, wd, yaw0) from default_rng(seed0 + env) in the reference draw order."""
    w = cfg["wind"]
    n = len(env_ids)
    ws, ti, wd, yaw0 = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros((n, T))
    for k, e in enumerate(env_ids):
        rng = np.random.default_rng(seed0 + int(e))
        ws[k] = rng.uniform(w["ws_min"], w["ws_max"])
        ti[k] = rng.uniform(w["TI_min"], w["TI_max"])
        wd[k] = rng.uniform(w["wd_min"], w["wd_max"])
        yaw0[k] = rng.uniform(-15.0, 15.0, T)
    return ws, ti, wd, yaw0


def n_passthrough_for(total_steps, cfg, D=80.0):
    """Episode length knob (reference ctor arg): long enough that no env truncates inside the run, so that the
    timed window is steady state (SURVEY.md 8(d) metric (i)).  time_max = int(dist/ws*n_passthrough) >= total."""
    f = cfg["farm"]
    # shortest wind-aligned farm extent over the sampled wd range is bounded below by the aligned extent * cos(15deg)
    ext = D * f["xDist"] * f["nx"] * math.cos(math.radians(16.0)) if f["nx"] > 1 else D * f["yDist"] * f["ny"] * 0.25
    t_inflow_min = ext / cfg["wind"]["ws_max"]
    return max(5, int(math.ceil((total_steps + 8) / t_inflow_min)))


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(index)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- CPU arms
_BARRIER = None  # inherited by the forked workers of run_reference


def _oracle_worker(args):
    """One process of the SubprocVecEnv pattern (reference examples/longer_steps_example.py:194-209): its own env."""
    env_id, nx, ny, reward, steps, warmup, use_barrier, budget_s = args
    barrier = _BARRIER if use_barrier else None
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle.env_numpy import WindFarmEnvOracle
    from oracle.v80 import V80 as OracleV80
    cfg = workload_config(nx, ny, reward)
    T = nx * ny
    ws, ti, wd, yaw0 = sample_conditions(cfg, [env_id], T)
    env = WindFarmEnvOracle(OracleV80(), cfg, reset_init=False, n_passthrough=n_passthrough_for(steps + warmup, cfg))
    env.reset(wind=(ws[0], ti[0], wd[0]), yaw0=yaw0[0])
    rng = np.random.default_rng(1234 + env_id)
    acts = rng.uniform(-1, 1, (steps + warmup, T)).astype(np.float32)
    for a in acts[:warmup]:
        env.step(a)
    if barrier is not None:
        barrier.wait()
    t0 = time.perf_counter()
    done = 0
    for a in acts[warmup:]:
        env.step(a)
        done += 1
        if budget_s and time.perf_counter() - t0 > budget_s:
            break
    return done, time.perf_counter() - t0


def cpu_port_single(nx, ny, reward, budget_s=12.0, max_steps=4000):
    """Oracle port, one env on one core, bounded sample (reported beside the GPU number; not the target)."""
    done, dt = _oracle_worker((0, nx, ny, reward, max_steps, 3, False, budget_s))
    return {"value": done / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"1 env (env 0 of the workload, {nx}x{ny} farm) x {done} steps after reset+3 warm-up steps, "
                      f"{dt:.1f} s; numpy fp64 oracle/env_numpy.py over oracle/dwm_numpy.py"}


def run_reference(a):
    """--impl reference: the CPU implementation of the path on all host cores.  The reference is pure Python over
    un-vendored dynamiks/py_wake (absent here and on the GPU box), so this arm times the oracle PORT
    (oracle/env_numpy.py + oracle/dwm_numpy.py), one env per core in separate processes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    global _BARRIER
    _BARRIER = ctx.Barrier(cores)
    jobs = [(e, a.nx, a.ny, a.reward, a.steps, a.warmup, True, 0.0) for e in range(cores)]
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_oracle_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    value = sum(d / t for d, t in res)
    t_step = max(t for _, t in res) / a.steps
    sample = (f"{cores} envs (one per host core, envs 0..{cores - 1} of the workload) x {a.steps} steps after reset + "
              f"{a.warmup} warm-up steps; whole run {wall:.1f} s; oracle PORT (oracle/env_numpy.py over "
              "oracle/dwm_numpy.py); the unmodified reference env layer over the same restated solver is no faster "
              "(14.7 vs 10.0 ms/step on this workload, profiles/r05_reference_layer_vs_port.txt): a conservative CPU arm")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(a, cores, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)
    return 0


def bench_config(a, envs_per_rank, where, world=1):
    total = envs_per_rank * world
    return {"workload": f"BASELINE.json configs[{1 if world == 1 else 2}]: {a.nx * a.ny}-turbine {a.nx}x{a.ny} grid (V80, "
                        f"reference linspace layout), " +
                        (f"{total} envs total, {envs_per_rank} per GPU x {world} GPU(s)" if where == "gpu" else
                         f"{envs_per_rank} envs per step (one per host core)") +
                        ", yaw-only actions U(-1,1), Env1.yaml obs/yaw semantics, " +
                        ("turbtype None (uniform inflow)" if getattr(a, "turbtype", "None") == "None" else
                         "Mann turbulence box shared by the envs"),
            "n_turb": a.nx * a.ny, "envs_total": total if where == "gpu" else None,
            "envs_per_gpu": envs_per_rank if where == "gpu" else None,
            "farms_per_env": 2 if a.reward == "Baseline" else 1, "power_reward": a.reward,
            "dt_env": 1, "dt_sim": 1, "obs_dim": 2 * a.nx * a.ny, "parallelism": f"env-sharded x{a.gpus}",
            "turbtype": getattr(a, "turbtype", "None"),
            "l2_policy": "working set (wake state, >= 200 MB per step and GPU) exceeds the 126 MB L2; no explicit flush",
            "step_overlap": "free-running timed loop: each step's flow kernel is a programmatic dependent launch of the "
                            "previous step's finish kernel (its wake-advection part overlaps it; bit-identical results); "
                            "e2e waits for every step's results on the host, no overlap there"}


# ---------------------------------------------------------------------------------------------- GPU arm
def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def measure(torch, dist, env, acts_host, K, W, K_e2e, K_prof, world, extra_bytes_per_station=0):
    """One workload on this rank's env: W warm-up + K device-timed steps (actions resident in HBM), then K_e2e steps
    through step_host (host buffers in and out, synchronised every step), then K_prof steps with the library's own
    CUDA events around each kernel.  Times are the max over ranks; station counts are summed over ranks."""
    dev, B = env.device, env.n_envs
    acts_dev = acts_host[:W + K].to(dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        env.step(acts_dev[i])
    barrier()
    l0 = env.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(W, W + K):
        env.step(acts_dev[i])
    e1.record()
    torch.cuda.synchronize()
    launches = env.launch_count - l0
    t_ms = e0.elapsed_time(e1)
    barrier()
    env.check_flags()

    # end to end through the public API with HOST buffers (VecWindFarmEnv.step_host -> C-ABI wg_step_host): the
    # actions come from pinned host memory and obs | reward | truncated are back in host memory, the call returns
    # when they are -- every step
    base = W + K
    for i in range(3):
        obs_h, rew_h, tr_h = env.step_host(acts_host[base + i])
    base += 3
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    any_trunc = False
    for i in range(K_e2e):
        obs_h, rew_h, tr_h = env.step_host(acts_host[base + i])   # host in, host out, synchronised
        any_trunc |= bool(tr_h.any())                             # the caller reads the result every step
    e3.record()
    torch.cuda.synchronize()
    t_e2e_ms = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3)
    base += K_e2e
    h2d = int(acts_host[0].numel()) * 4
    d2h = obs_h.size * 4 + B * 4 + B
    assert not any_trunc and not bool(tr_h.any()), "an env truncated inside the timed window"

    # dominant kernel alone: CUDA events recorded by the library on the launching stream around each kernel
    env.profile_enable(True)
    acts_p = acts_host[base:base + K_prof].to(dev)
    torch.cuda.synchronize()
    for i in range(K_prof):
        env.step(acts_p[i])
    flow_ms, fin_ms, n_prof = env.profile_read()
    env.profile_enable(False)
    # live wake stations (all envs, farms, chains): the ring counts minus the stations the next step drops
    live = int(env.state["count"].sum().item()) - int(env.state["retire"].sum().item())
    env.check_flags()
    F, S, T = env.n_farms, env.ec.S, env.n_turb
    tt = torch.tensor([t_ms, t_e2e_ms, flow_ms / max(n_prof, 1), fin_ms / max(n_prof, 1)], dtype=torch.float64, device=dev)
    cnt = torch.tensor([live, B], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    t_ms, t_e2e_ms, t_flow_ms, t_fin_ms = [float(x) for x in tt.tolist()]
    live_all, B_all = [float(x) for x in cnt.tolist()]
    # algorithmic bytes of one flow launch, summed over the ranks; per GPU = / world
    bytes_flow = S * (live_all * (STATION_BYTES + extra_bytes_per_station) + B_all * F * T * TURB_BYTES)
    peak, peak_src = hbm_peak()
    achieved = bytes_flow / world / (t_flow_ms * 1e-3) / 1e9
    return {
        "value": B_all * K / (t_ms * 1e-3), "ms_per_step": t_ms / K, "launches": int(launches),
        "e2e": {"value": B_all * K_e2e / (t_e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e_ms / K_e2e},
        "roofline": {"kernel": "wg_flow_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "ms_per_launch": t_flow_ms,
                     "finish_kernel_ms": t_fin_ms, "live_stations_per_env_farm": live_all / (B_all * F),
                     "bytes_per_station": STATION_BYTES + extra_bytes_per_station,
                     "algorithmic_bytes_per_launch_per_gpu": bytes_flow / world, "launches_timed": int(n_prof),
                     # SURVEY 8(d) "for the fused step": the same bytes over the whole device step of the timed loop
                     "step_achieved": bytes_flow / world / (t_ms / K * 1e-3) / 1e9,
                     "step_frac": bytes_flow / world / (t_ms / K * 1e-3) / 1e9 / peak,
                     "per": "GPU (bytes of all ranks / n_gpus / max-over-ranks launch time)"},
    }


def make_env(torch, dev, rank, B, nx, ny, reward, total_steps, seed_base=0, **kw):
    """A VecWindFarmEnv of this rank's B envs (global env ids rank*B ...), reset with the workload's conditions."""
    from windgym_b200 import V80, VecWindFarmEnv
    cfg = workload_config(nx, ny, reward)
    n_pass = n_passthrough_for(total_steps, cfg)
    env_ids = np.arange(rank * B, (rank + 1) * B)
    ws, ti, wd, yaw0 = sample_conditions(cfg, env_ids, nx * ny, seed0=seed_base)
    env = VecWindFarmEnv(V80(), B, config=cfg, device=str(dev), n_passthrough=max(n_pass, kw.pop("n_passthrough", 0)),
                         seed=rank, **kw)
    env.reset(wind=(ws, ti, wd), yaw0=yaw0)
    torch.cuda.synchronize()
    assert int(np.min(env.time_max)) > total_steps, "episode would truncate inside the run"
    return env, cfg, n_pass


def side_leg(torch, dist, dev, rank, world, name, B, nx, ny, K, W, what, act_mult=1, extra_bytes_per_station=0, **kw):
    """A secondary configuration measured with the same procedure as the headline one (fewer steps)."""
    K_prof = min(K, 32)
    n_act = W + K + 3 + K + K_prof
    env, cfg, _ = make_env(torch, dev, rank, B, nx, ny, "Power_avg", n_act + 8, **kw)
    gen = torch.Generator(device="cpu").manual_seed(4321 + rank)
    acts = (torch.rand((n_act, B, nx * ny * act_mult), generator=gen, dtype=torch.float32) * 2 - 1).pin_memory()
    m = measure(torch, dist, env, acts, K, W, K, K_prof, world, extra_bytes_per_station)
    env.close()
    del env
    torch.cuda.empty_cache()
    r = m["roofline"]
    return {"what": what, "envs_per_gpu": B, "envs_total": B * world, "n_turb": nx * ny, "steps": K,
            "value": m["value"], "unit": UNIT, "ms_per_step": m["ms_per_step"], "e2e": m["e2e"]["value"],
            "e2e_ms_per_step": m["e2e"]["ms_per_step"], "roofline_frac": r["frac"], "achieved_gbs": r["achieved"],
            "ms_per_launch": r["ms_per_launch"], "finish_kernel_ms": r["finish_kernel_ms"],
            "live_stations_per_env_farm": r["live_stations_per_env_farm"], "bytes_per_station": r["bytes_per_station"]}


def run_gpu(a):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched through torch.distributed.run (one rank per GPU)")
        a.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: windgym_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from windgym_b200 import V80

    # STRONG scaling (BASELINE.json metric: "4096 envs at 1/2/4/8 B200"; configs[2]: 512 per GPU at 8): the batch of
    # --envs-total envs is sharded over the ranks.  --envs pins the per-GPU count instead (weak scaling).
    scaling = "strong"
    if a.envs is not None:
        B, scaling = a.envs, ("weak" if world > 1 else "strong")
    else:
        if a.envs_total % world:
            raise SystemExit(f"--envs-total {a.envs_total} is not divisible by {world} ranks")
        B = a.envs_total // world
    T, K, W = a.nx * a.ny, a.steps, a.warmup
    K_e2e, K_prof = K, min(K, 64)
    n_act = W + K + 3 + K_e2e + K_prof
    kw = {}
    extra_b = 0
    if a.turbtype == "Mann":   # ambient turbulence + wake-added turbulence: MannFixed (Wind_Farm_Env.py:646-656)
        kw, extra_b = mann_kwargs(dev, a.mann_box), MANN_GATHER_BYTES
    env, cfg, n_pass = make_env(torch, dev, rank, B, a.nx, a.ny, a.reward, n_act + 8, **kw)
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    acts_host = (torch.rand((n_act, B, T), generator=gen, dtype=torch.float32) * 2 - 1).pin_memory()
    acts_dev = acts_host[:W + K].to(dev)

    clocks = ClockSampler(local) if rank == 0 else None
    m = measure(torch, dist, env, acts_host, K, W, K_e2e, K_prof, world, extra_b)
    clk = clocks.stop() if clocks else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- SURVEY 8(d) metric (ii): throughput WITH auto-reset (episodes of the reference's length, n_passthrough = 5,
    # envs at random phases of their episodes, finished envs replaced from the spare pool, spin-up in the background)
    auto = None
    env.close(); del env
    torch.cuda.empty_cache()

    def autoreset_leg(device_side):
        from windgym_b200 import DevicePooledVecEnv, PooledVecEnv
        from windgym_b200.vector import GymVectorEnv
        # long enough (>= 0.3 s of stepping) that the ~20 ms of background spin-ups still in flight at the end, which
        # the closing synchronize waits for, do not distort the figure
        K_ar = max(K, 200, int(300.0 / max(m["ms_per_step"], 1e-3)) if device_side else 0)
        R = max(384 if device_side else 64, B // 8)
        cls = DevicePooledVecEnv if device_side else PooledVecEnv
        pool = cls(V80(), B, reserve=R, config=cfg, device=str(dev), n_passthrough=5, seed=rank, **kw)
        genv = GymVectorEnv(venv=pool, as_torch=True)
        genv.reset(seed=rank)
        prng = np.random.default_rng(rank)
        tmax = pool.time_max.cpu().numpy() if torch.is_tensor(pool.time_max) else pool.time_max
        pool.state["timestep"][:] = torch.as_tensor((prng.uniform(0, 1, B) * tmax).astype(np.int32), device=dev)
        for i in range(120):                                 # the spares' first spin-up completes in the background
            genv.step(acts_dev[i % (W + K)])
        barrier()
        n_res = torch.zeros((), dtype=torch.int64, device=dev)
        t0 = time.perf_counter()
        for i in range(K_ar):
            _, _, _, tr_ar, _ = genv.step(acts_dev[i % (W + K)])
            if device_side:
                n_res += tr_ar.sum()                            # stays on the device: no read-back in the loop
            else:
                n_res += int(tr_ar.sum())
        torch.cuda.synchronize()
        t_ar = (time.perf_counter() - t0) * 1e3
        pool.check_flags()
        res = {"ms": t_ar, "steps": K_ar, "resets": int(n_res.item()), "stats": dict(pool.stats), "reserve": R}
        pool.close()
        del genv, pool
        torch.cuda.empty_cache()
        ta = torch.tensor([res["ms"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        res["ms"] = float(ta.item())
        return res

    auto_host = None
    if not a.no_autoreset:
        auto = autoreset_leg(True)
        if world == 1:
            auto_host = autoreset_leg(False)

    # ---- the other BASELINE.json configurations, same procedure, fewer steps (driver-visible: "configs")
    extra = {}
    if not a.no_extras:
        Ks, Ws = max(20, min(K, 40)), 5
        if world == 1:
            extra["cfg3_share_512_envs_1gpu"] = side_leg(
                torch, dist, dev, rank, world, "cfg3", 512, 4, 4, Ks, Ws,
                "one GPU's share of configs[2] (4096 envs sharded 512/GPU over 8 GPUs): 512 envs x 4x4 farm on 1 GPU")
            extra["cfg4_8x8_1024_envs_yaw_induction"] = side_leg(
                torch, dist, dev, rank, world, "cfg4", 1024, 8, 8, Ks, Ws,
                "configs[3]: 64-turbine 8x8 farm, 1024 envs, yaw + induction actions (act_var = 2 extension)",
                act_mult=2, induction_control=True)
            extra["cfg2_mann"] = side_leg(
                torch, dist, dev, rank, world, "mann", 4096, 4, 4, Ks, Ws,
                "configs[1] with turbtype MannFixed: ambient Mann box + wake-added turbulence (meandering wakes); bytes "
                "per station include the 8-corner gather of the low-pass box", extra_bytes_per_station=MANN_GATHER_BYTES,
                **mann_kwargs(dev, a.mann_box))
        if 2048 % world == 0:
            extra["cfg5_multi_agent_4x2_2048_envs"] = side_leg(
                torch, dist, dev, rank, world, "cfg5", 2048 // world, 4, 2, Ks, Ws,
                f"configs[4]: WindFarmEnvMulti 8-turbine 4x2 farm, 2048 envs total ({2048 // world} per GPU x {world}), "
                "per-agent observation rows obs f32[B, 8, 2]", multi_agent=True, n_passthrough=20)
        if world > 1 and not a.no_weak:
            extra["weak_scaling_4096_envs_per_gpu"] = side_leg(
                torch, dist, dev, rank, world, "weak", 4096, 4, 4, Ks, Ws,
                f"weak scaling: 4096 envs PER GPU ({4096 * world} envs total); not the BASELINE metric")

    if rank == 0:
        r = m["roofline"]
        # DRAM bytes of one launch from the committed `ncu --set full` capture of this exact workload (null otherwise)
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "flow_ncu_traffic.json")) as fh:
                tr_ = json.load(fh)
            if (tr_["nx"], tr_["ny"], tr_["envs"], tr_["reward"], tr_["turbtype"]) == (a.nx, a.ny, B, a.reward, a.turbtype):
                traffic, traffic_src = tr_["dram_bytes_read"] + tr_["dram_bytes_write"], tr_["source"]
        except Exception:
            pass
        r["traffic"], r["traffic_source"] = traffic, traffic_src
        # SURVEY.md 8(d) sizes its state model at N_p = 804 stations per 4x4 farm (16 m spacing, wd = 270); the frozen
        # specification releases a particle every ceil(16 m / (ws dt)) steps and holds fewer (the
        # rotor powers move by < 1e-4 between the two spacings).  Time is proportional to stations: the same kernel on
        # the SURVEY's state model would deliver value x live / 804.
        if (a.nx, a.ny) == (4, 4):
            r["survey_np"] = 804
            r["value_at_survey_np_equiv"] = m["value"] * min(1.0, r["live_stations_per_env_farm"] / 804.0)
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": bench_config(a, B, "gpu", world),
            "e2e": m["e2e"], "gpu_launches": m["launches"], "roofline": r, "clocks": clk,
        }
        def ar_entry(r, what):
            return {"value": world * B * r["steps"] / (r["ms"] * 1e-3), "unit": UNIT, "ms_per_step": r["ms"] / r["steps"],
                    "steps": r["steps"], "episodes_finished_rank0": r["resets"], "spare_envs": r["reserve"],
                    "pool": r["stats"], "what": what}
        if auto is not None:
            line["with_autoreset"] = ar_entry(
                auto, "steady-state training loop, device-side pool (DevicePooledVecEnv / wg_pool_*): n_passthrough=5 "
                      "episodes at random phases; finished episodes are paired with pre-developed spares and replaced on "
                      "the device (no flag read-back, no host decision), consumed spares are re-drawn and spun up on 8 background "
                      "streams; wall clock incl. all reset work and the closing synchronize, max over ranks")
        if auto_host is not None:
            line["with_autoreset_host_pool"] = ar_entry(
                auto_host, "same loop with the host-driven pool of round 1 (PooledVecEnv: truncation flags read back and "
                           "swaps decided on the host every step, numpy PCG64 condition draws)")
        if extra:
            line["configs"] = extra
        line["config"]["n_passthrough"] = n_pass
        if world == 1 and not a.no_cpu:
            line["cpu_baseline"] = cpu_port_single(a.nx, a.ny, a.reward)
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


MANN_GATHER_BYTES = 8 * 8   # low-pass box: 8 corners x float2 per station and step (random 32-byte sector reads)


def mann_kwargs(dev, size):
    """turbtype MannFixed.  size 'ref': the reference's box 2048 x 512 x 64 @ 3 m (Wind_Farm_Env.py:646-656);
    'test': the reduced box of the reference's tests, 1024 x 128 x 32 (tests/test_basics.py:38-45)."""
    from windgym_b200.mann import MannBox
    nxyz = (2048, 512, 64) if size == "ref" else (1024, 128, 32)
    return dict(turbtype="MannFixed", turb_box=MannBox.generate(0.1, 33.6, 3.9, Nxyz=nxyz, dxyz=(3.0, 3.0, 3.0), seed=1234,
                                                                device=str(dev)))


_JSON_FD = None


def reserve_stdout():
    """Keep the real stdout for the ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library
    chatter) is sent to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-total", type=int, default=4096, help="envs of the whole job, sharded over the GPUs (strong scaling)")
    ap.add_argument("--envs", type=int, default=None, help="envs PER GPU (overrides --envs-total; weak scaling under torchrun)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configurations (configs key)")
    ap.add_argument("--no-weak", action="store_true", help="N>1: skip the weak-scaling leg (4096 envs per GPU)")
    ap.add_argument("--mann-box", default="ref", choices=["ref", "test"], help="Mann box size of --turbtype Mann / cfg2_mann")
    ap.add_argument("--nx", type=int, default=4)
    ap.add_argument("--ny", type=int, default=4)
    ap.add_argument("--reward", default="Power_avg", choices=["Power_avg", "Baseline"],
                    help="Baseline adds the second (greedy-controller) farm per env: 2x the flow work")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-autoreset", action="store_true", help="skip the with_autoreset leg")
    ap.add_argument("--turbtype", default="None", choices=["None", "Mann"],
                    help="Mann: ambient Mann turbulence box (meandering + rotor fluctuations); not the headline config")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.impl == "reference":
        return run_reference(a)
    return run_gpu(a)


if __name__ == "__main__":
    sys.exit(main())

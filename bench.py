#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched WindFarmEnv.step hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # CPU arm: oracle port on the host cores

One "step" = one batched ``VecWindFarmEnv.step`` over all envs of the rank: yaw update, S flow substeps of every
farm (wake advection + Ainslie march + superposition + rotor average + turbine update), MesClass push/extract,
reward, truncation.  Workload = BASELINE.json configs[1]: 16-turbine 4x4 grid, 4096 envs per GPU, yaw-only actions,
Env1.yaml observation/yaw semantics, uniform inflow (turbtype "None"), synthetic U(-1,1) actions.
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
              f"{a.warmup} warm-up steps; whole run {wall:.1f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(a, cores, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)
    return 0


def bench_config(a, envs_per_rank, where):
    return {"workload": f"BASELINE.json configs[1]: {a.nx * a.ny}-turbine {a.nx}x{a.ny} grid (V80, reference linspace "
                        f"layout), {envs_per_rank} envs per {'GPU' if where == 'gpu' else 'step (one per host core)'}, "
                        "yaw-only actions U(-1,1), Env1.yaml obs/yaw semantics, " +
                        ("turbtype None (uniform inflow)" if getattr(a, "turbtype", "None") == "None" else
                         "Mann turbulence box 1024x128x32 @ 3 m shared by the envs"),
            "n_turb": a.nx * a.ny, "envs_per_gpu": envs_per_rank if where == "gpu" else None,
            "farms_per_env": 2 if a.reward == "Baseline" else 1, "power_reward": a.reward,
            "dt_env": 1, "dt_sim": 1, "obs_dim": 2 * a.nx * a.ny, "parallelism": f"env-sharded x{a.gpus}",
            "turbtype": getattr(a, "turbtype", "None"),
            "l2_policy": "working set (wake state, GBs per step) exceeds the 126 MB L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------- GPU arm
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
    from windgym_b200 import V80, VecWindFarmEnv

    B, T, K, W = a.envs, a.nx * a.ny, a.steps, a.warmup
    K_e2e = K
    K_prof = min(K, 64)
    total = W + K + 3 + K_e2e + K_prof + 8
    cfg = workload_config(a.nx, a.ny, a.reward)
    n_pass = n_passthrough_for(total, cfg)
    env_ids = np.arange(rank * B, (rank + 1) * B)
    ws, ti, wd, yaw0 = sample_conditions(cfg, env_ids, T)
    kw = {}
    if a.turbtype == "Mann":   # ambient turbulence: the reference tests' reduced box (tests/test_basics.py:38-45)
        from windgym_b200.mann import MannBox
        kw = dict(turbtype="MannFixed", turb_box=MannBox.generate(0.1, 33.6, 3.9, Nxyz=(1024, 128, 32), dxyz=(3.0, 3.0, 3.0),
                                                                  seed=1234, device=str(dev)))
    env = VecWindFarmEnv(V80(), B, config=cfg, device=str(dev), n_passthrough=n_pass, seed=rank, **kw)
    env.reset(wind=(ws, ti, wd), yaw0=yaw0)
    torch.cuda.synchronize()
    assert int(np.min(env.time_max)) > total, "episode would truncate inside the run"

    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
    n_act = W + K + 3 + K_e2e + K_prof
    acts_host = (torch.rand((n_act, B, T), generator=gen, dtype=torch.float32) * 2 - 1).pin_memory()
    acts_dev = acts_host[:W + K].to(dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then K timed steps with the inputs resident in HBM
    for i in range(W):
        env.step(acts_dev[i])
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
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

    # ---- end to end through the public API with HOST buffers (VecWindFarmEnv.step_host): pinned actions H2D, the
    # step, obs | reward | truncated D2H (one copy of the packed result buffer), stream synchronised -- every step
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
    t_e2e_wall = (time.perf_counter() - t0) * 1e3
    t_e2e_ms = max(e2.elapsed_time(e3), t_e2e_wall)
    clk = clocks.stop() if clocks else None
    base += K_e2e
    h2d = B * T * 4
    d2h = obs_h.size * 4 + B * 4 + B
    assert not any_trunc and not bool(tr_h.any()), "an env truncated inside the timed window"

    # ---- dominant kernel alone: CUDA events recorded by the library on the launching stream around each kernel
    env.profile_enable(True)
    acts_p = acts_host[base:base + K_prof].to(dev)
    torch.cuda.synchronize()
    for i in range(K_prof):
        env.step(acts_p[i])
    flow_ms, fin_ms, n_prof = env.profile_read()
    env.profile_enable(False)
    # live wake stations of this rank (all envs, farms, chains): the ring counts minus the stations the next step drops
    live = int(env.state["count"].sum().item()) - int(env.state["retire"].sum().item())
    F, S = env.n_farms, env.ec.S
    # with a turbulence box every station also gathers 8 corners x 8 B of the low-pass box (not counted as
    # algorithmic state traffic: the box is shared and L2-resident at this size)
    bytes_flow = S * (live * STATION_BYTES + B * F * T * TURB_BYTES)
    t_flow = flow_ms / max(n_prof, 1) * 1e-3
    env.check_flags()

    # ---- SURVEY 8(d) metric (ii): throughput WITH auto-reset (episodes of the reference's length, n_passthrough = 5,
    # envs at random phases of their episodes, finished envs replaced from the spare pool, spin-up in the background)
    auto = None
    if not a.no_autoreset:
        from windgym_b200 import PooledVecEnv
        from windgym_b200.vector import GymVectorEnv
        env.close(); del env
        torch.cuda.empty_cache()
        K_ar = max(K, 200)
        pool = PooledVecEnv(V80(), B, reserve=max(64, B // 8), config=cfg, device=str(dev), n_passthrough=5, seed=rank, **kw)
        genv = GymVectorEnv(venv=pool, as_torch=True)
        genv.reset(seed=rank)
        prng = np.random.default_rng(rank)
        pool.state["timestep"][:] = torch.as_tensor((prng.uniform(0, 1, B) * pool.time_max).astype(np.int32), device=dev)
        for i in range(10):
            genv.step(acts_dev[i % (W + K)])
        barrier()
        n_res = 0
        t0 = time.perf_counter()
        for i in range(K_ar):
            _, _, _, tr_ar, _ = genv.step(acts_dev[i % (W + K)])
            n_res += int(tr_ar.sum())
        torch.cuda.synchronize()
        t_ar = (time.perf_counter() - t0) * 1e3
        pool.check_flags()
        auto = {"ms": t_ar, "steps": K_ar, "resets": n_res, "stats": dict(pool.stats), "reserve": pool.reserve}
        pool.close()

    # ---- max over ranks
    tt = torch.tensor([t_ms, t_e2e_ms, t_flow * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, t_e2e_ms, t_flow_ms = [float(x) for x in tt.tolist()]

    if rank == 0:
        peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                mp_ = json.load(fh)
            peak, peak_src = float(mp_["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
        achieved = bytes_flow / (t_flow_ms * 1e-3) / 1e9
        # DRAM bytes of one launch from the committed `ncu --set full` capture of this exact workload (null otherwise)
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "flow_ncu_traffic.json")) as fh:
                tr_ = json.load(fh)
            if (tr_["nx"], tr_["ny"], tr_["envs"], tr_["reward"], tr_["turbtype"]) == (a.nx, a.ny, B, a.reward, a.turbtype):
                traffic, traffic_src = tr_["dram_bytes_read"] + tr_["dram_bytes_write"], tr_["source"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": world * B * K / (t_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": bench_config(a, B, "gpu"),
            "e2e": {"value": world * B * K_e2e / (t_e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e_ms / K_e2e},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "wg_flow_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "ms_per_launch": t_flow_ms, "finish_kernel_ms": fin_ms / max(n_prof, 1),
                         "live_stations_per_env_farm": live / (B * F), "bytes_per_station": STATION_BYTES,
                         "algorithmic_bytes_per_launch": bytes_flow, "launches_timed": int(n_prof)},
            "clocks": clk,
        }
        if auto is not None:   # rank 0's share, whole-job value extrapolated over the ranks (no cross-rank coupling)
            line["with_autoreset"] = {
                "value": world * B * auto["steps"] / (auto["ms"] * 1e-3), "unit": UNIT, "ms_per_step": auto["ms"] / auto["steps"],
                "steps": auto["steps"], "episodes_finished": auto["resets"], "spare_envs": auto["reserve"],
                "pool": auto["stats"],
                "what": "steady-state training loop: n_passthrough=5 episodes at random phases, finished envs swapped for "
                        "pre-developed spares (spin-up batched on a background stream), truncation flags read on the host "
                        "every step (wall clock)"}
        line["config"]["n_passthrough"] = n_pass
        if world == 1 and not a.no_cpu:
            line["cpu_baseline"] = cpu_port_single(a.nx, a.ny, a.reward)
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


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
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
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

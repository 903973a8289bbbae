"""VecWindFarmEnv -- thousands of independent WindFarmEnv instances stepped by one CUDA launch.

Batched counterpart of the reference ``WindFarmEnv`` (``WindGym/Wind_Farm_Env.py:47``): same constructor
arguments, same YAML schema, same ``reset()/step()`` contract, with a leading env axis on every array.
torch owns all device memory; the compute is ``libwindgym_b200.so`` behind the C-ABI in
``include/windgym_b200.h`` (no CPU fallback).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .fast_rng import uniform_streams
from .config import EnvConfig, load_yaml

_TORCH_DT = {0: torch.float32, 1: torch.int32}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class VecWindFarmEnv:
    def __init__(self, turbine, n_envs, yaml_path=None, config=None, n_passthrough=5, TI_min_mes=0.0,
                 TI_max_mes=0.50, TurbBox="Default", turbtype="None", Baseline_comp=False, yaw_init=None,
                 seed=None, dt_sim=1, dt_env=1, yaw_step=1, fill_window=True, device="cuda:0",
                 multi_agent=False, eval_mode=False, noise_seed=0, reset_init=False, sample_site=None, turb_box=None,
                 added_turbulence=None, induction_control=False, derate_min=0.5):
        cfg = config if config is not None else load_yaml(yaml_path)
        self.ec = ec = EnvConfig(cfg, turbine, n_passthrough=n_passthrough, TI_min_mes=TI_min_mes,
                                 TI_max_mes=TI_max_mes, turbtype=turbtype, Baseline_comp=Baseline_comp,
                                 yaw_init=yaw_init, dt_sim=dt_sim, dt_env=dt_env, yaw_step=yaw_step,
                                 fill_window=fill_window, eval_mode=eval_mode, multi_agent=multi_agent,
                                 noise_seed=noise_seed, induction_control=induction_control, derate_min=derate_min)
        self.turbine = turbine
        self.n_envs, self.n_turb = int(n_envs), ec.n_turb
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.WgError("VecWindFarmEnv runs on a CUDA device only (no CPU fallback)")
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.seed = seed
        self.sample_site = sample_site
        self._site_tables = None
        self.x_pos, self.y_pos = ec.x_pos, ec.y_pos
        self.yaw_min, self.yaw_max, self.yaw_step = ec.yaw_min, ec.yaw_max, yaw_step
        self.Baseline_comp = ec.Baseline_comp
        self.n_farms = 2 if ec.Baseline_comp else 1
        self.yaw_initial = [0]
        self._wind_override = {}
        self._episode = 0
        self.lib = _lib.load()
        torch.cuda.set_device(self.device)
        self._create()
        self.turb_box = None
        self.added_box = None
        if ec.turbtype != "None":
            self._attach_turbulence(turb_box, TurbBox)
            # addedTurbulenceModel of the Mann site types (Wind_Farm_Env.py:618,:639,:658); a MannBox injects the
            # isotropic box, False switches the model off
            if added_turbulence is not False:
                from .mann import MannBox
                self.added_box = added_turbulence if isinstance(added_turbulence, MannBox) else \
                    MannBox.isotropic_unit(ec.D, seed=4321, device=self.device)
                nx, ny, nz = self.added_box.Nxyz
                _lib.check(self.lib.wg_set_added_turbulence(self._h, _ptr(self.added_box.raw), nx, ny, nz,
                                                            *self.added_box.dxyz, 0.6, 0.35))
        self.obs_shape = (self.n_envs, self.n_turb, self.obs_var) if multi_agent else (self.n_envs, self.obs_var)
        # obs | reward | truncated live in ONE device buffer, so that a host caller gets a step's results with a
        # single device-to-host copy (step_host)
        n_obs = int(np.prod(self.obs_shape))
        self._out = torch.zeros(n_obs * 4 + self.n_envs * 4 + self.n_envs, dtype=torch.uint8, device=self.device)
        self.obs = self._out[:n_obs * 4].view(torch.float32).view(self.obs_shape)
        self.reward = self._out[n_obs * 4:(n_obs + self.n_envs) * 4].view(torch.float32)
        self.truncated = self._out[(n_obs + self.n_envs) * 4:]
        self._host = None   # pinned host staging of step_host (allocated on first use)
        self.terminated = torch.zeros(self.n_envs, dtype=torch.bool, device=self.device)
        self._step_ptrs = (_ptr(self._state), _ptr(self.obs), _ptr(self.reward), _ptr(self.truncated))
        self._n_act_elems = self.n_envs * self.n_turb * self.ec.act_var   # elements step() expects (active envs)
        # host copies of the per-env wind conditions of the current episode
        self.ws = np.zeros(self.n_envs); self.ti = np.zeros(self.n_envs); self.wd = np.zeros(self.n_envs)
        if reset_init:
            self.reset(seed=seed)

    # ------------------------------------------------------------------------------------------ C-ABI plumbing
    def _create(self):
        ec, c = self.ec, _lib
        t = self.turbine
        self._tabs = [np.ascontiguousarray(a, dtype=np.float32) for a in (t.ws_table, t.power_table_w, t.ct_table)]
        self._xy = [np.ascontiguousarray(a, dtype=np.float64) for a in (ec.x_pos, ec.y_pos)]
        codes = ec.codes()

        def chan(m, k):
            return c.MesChannel(int(m[f"{k}_current"]), int(m[f"{k}_rolling_mean"]), int(m[f"{k}_history_N"]),
                                int(m[f"{k}_history_length"]), int(m[f"{k}_window_length"]))

        lv = ec.mes_level
        mes = c.MesConfig(
            ws=chan(ec.ws_mes, "ws"), wd=chan(ec.wd_mes, "wd"), yaw=chan(ec.yaw_mes, "yaw"),
            power=chan(ec.power_mes, "power"),
            turb_ws=int(lv["turb_ws"]), turb_wd=int(lv["turb_wd"]), turb_TI=int(lv["turb_TI"]),
            turb_power=int(lv["turb_power"]), farm_ws=int(lv["farm_ws"]), farm_wd=int(lv["farm_wd"]),
            farm_TI=int(lv["farm_TI"]), farm_power=int(lv["farm_power"]),
            ws_min=2.0, ws_max=25.0, wd_min=ec.wd_min_mes - 5, wd_max=ec.wd_max_mes + 5,  # Wind_Farm_Env.py:440-444
            yaw_min=ec.yaw_min, yaw_max=ec.yaw_max, ti_min=ec.TI_min_mes, ti_max=ec.TI_max_mes,
            power_max=ec.maxturbpower, noise=1 if ec.noise == "Normal" else 0,
            noise_std=(C.c_float * 4)(0.0, 2.0, 0.0, 0.0), noise_seed=int(ec.noise_seed),
            multi_agent=int(ec.multi_agent))
        fp = C.POINTER(C.c_float)
        dp = C.POINTER(C.c_double)
        cfg = c.Config(
            n_envs=self.n_envs, n_turb=ec.n_turb, n_farms=self.n_farms, p_cap=ec.p_cap, substeps=ec.S,
            dt=float(ec.dt_sim), diameter=ec.D, hub_height=ec.hub_height, d_particle=ec.d_particle,
            n_tab=len(self._tabs[0]), tab_ws=self._tabs[0].ctypes.data_as(fp),
            tab_power=self._tabs[1].ctypes.data_as(fp), tab_ct=self._tabs[2].ctypes.data_as(fp),
            x_pos=self._xy[0].ctypes.data_as(dp), y_pos=self._xy[1].ctypes.data_as(dp),
            action_method=codes["action"], yaw_min=float(ec.yaw_min), yaw_max=float(ec.yaw_max),
            yaw_step=float(ec.yaw_step), base_controller=codes["controller"], power_reward=codes["reward"],
            power_avg=int(ec.power_avg), power_scaling=float(ec.Power_scaling),
            action_penalty=float(ec.action_penalty), action_penalty_type=codes["penalty"],
            steps_on_reset=int(ec.steps_on_reset), mes=mes, act_var=int(ec.act_var), derate_min=float(ec.derate_min))
        h = C.c_void_p()
        _lib.check(self.lib.wg_create(C.byref(cfg), C.byref(h)))
        self._h = h
        nbytes = C.c_size_t()
        _lib.check(self.lib.wg_state_bytes(h, C.byref(nbytes)))
        self.state_bytes = nbytes.value
        self._state = torch.zeros(self.state_bytes, dtype=torch.uint8, device=self.device)
        od = C.c_int32()
        _lib.check(self.lib.wg_obs_dim(h, C.byref(od)))
        self.obs_var = od.value
        self.state = {}
        i = 0
        while True:
            nm = self.lib.wg_state_field_name(h, i)
            if nm is None:
                break
            off, dt, nd, shp = C.c_size_t(), C.c_int32(), C.c_int32(), (C.c_int64 * 8)()
            _lib.check(self.lib.wg_state_field(h, nm, C.byref(off), C.byref(dt), C.byref(nd), shp))
            shape = [shp[k] for k in range(nd.value)]
            n = int(np.prod(shape)) * 4
            self.state[nm.decode()] = self._state[off.value:off.value + n].view(_TORCH_DT[dt.value]).view(shape)
            i += 1

    def _attach_turbulence(self, box, TurbBox):
        """``_def_site`` (Wind_Farm_Env.py:598-678) for the Mann site types.  ONE box per handle, shared read-only by
        all envs of the GPU; each env sits at its own offset inside the periodic box and carries its own
        ``scale_TI`` factor (the batched counterpart of one box per env).  ``turb_box`` injects a ready ``MannBox``
        (what the reference's tests do by patching ``MannTurbulenceField.generate``, tests/test_basics.py:56-63)."""
        import glob
        import os
        from .mann import MannBox
        ec = self.ec
        if box is None:
            if ec.turbtype == "MannFixed":      # :647-656
                box = MannBox.generate(0.1, 33.6, 3.9, Nxyz=(2048, 512, 64), dxyz=(3.0, 3.0, 3.0), seed=1234,
                                       device=self.device, lowpass_width=2 * ec.D)
            elif ec.turbtype == "MannGenerate":  # :621-637
                box = MannBox.generate(0.1, 33.6, 3.9, Nxyz=(4096, 512, 64), dxyz=(ec.D / 20, ec.D / 10, ec.D / 10),
                                       seed=0 if self.seed is None else int(self.seed), device=self.device,
                                       lowpass_width=2 * ec.D)
            elif ec.turbtype == "Random":        # :640-644: RandomTurbulence(ti, ws, seed) -- a white-noise box
                box = MannBox.white_noise(seed=0 if self.seed is None else int(self.seed), device=self.device,
                                          lowpass_width=2 * ec.D)
            else:                                 # MannLoad, :612-617: one of the files under TurbBox
                files = [TurbBox] if os.path.isfile(str(TurbBox)) else sorted(
                    f for ext in ("*.npz", "*.npy", "*.nc") for f in glob.glob(os.path.join(str(TurbBox), ext)))
                if not files:
                    raise FileNotFoundError(f"TurbBox={TurbBox!r}: no turbulence box file (.npz/.npy/.nc) found")
                self._tf_files = files
                pick = np.random.default_rng(self.seed).choice(len(files))
                box = MannBox.from_file(files[pick], device=self.device, lowpass_width=2 * ec.D)
        if box.device != self.device:
            raise ValueError(f"turbulence box lives on {box.device}, env on {self.device}")
        self.turb_box = box
        nx, ny, nz = box.Nxyz
        _lib.check(self.lib.wg_set_turbulence(self._h, _ptr(box.raw), _ptr(box.lp), nx, ny, nz, *box.dxyz))
        self.turb_offset = np.zeros((self.n_envs, 3))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.wg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        # raw handle of torch's current stream on this device (a plain C call: this sits on the per-step path)
        return torch._C._cuda_getCurrentRawStream(self._dev_index)

    @property
    def launch_count(self):
        n = C.c_uint64()
        _lib.check(self.lib.wg_launch_count(self._h, C.byref(n)))
        return n.value

    def profile_enable(self, on=True):
        """Record CUDA events around the two kernels of every step() (measurement aid, see wg_profile_enable)."""
        _lib.check(self.lib.wg_profile_enable(self._h, int(bool(on))))

    def profile_read(self):
        """(flow kernel ms, finish kernel ms, steps) summed over the recorded steps; synchronises."""
        a, b, n = C.c_double(), C.c_double(), C.c_uint64()
        _lib.check(self.lib.wg_profile_read(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    # ------------------------------------------------------------------------------------------ FarmEval surface
    def set_wind_vals(self, ws=None, ti=None, wd=None):
        """FarmEval.set_wind_vals (FarmEval.py:63-78); scalars or per-env arrays."""
        for k, v in (("ws", ws), ("ti", ti), ("wd", wd)):
            if v is not None:
                self._wind_override[k] = np.broadcast_to(np.asarray(v, dtype=np.float64), (self.n_envs,)).copy()

    def set_yaw_vals(self, yaw_vals):
        self.yaw_initial = yaw_vals

    # ------------------------------------------------------------------------------------------ reset / step
    def sample_conditions(self, seed=None, envs=None):
        """Per-env (ws, ti, wd, yaw0) with the reference's draw order ws -> ti -> wd -> yaw (Wind_Farm_Env.py:564-568,
        :715); env i of episode e uses np.random.default_rng([seed, i, e]) -- and exactly default_rng(seed) for a
        single env's first episode, which is what gymnasium seeds the reference with."""
        ec, B, T = self.ec, self.n_envs, self.n_turb
        envs = range(B) if envs is None else envs
        if not hasattr(self, "_tf_seed"):
            self._tf_seed = np.zeros(B, dtype=np.int64)
        ws, ti, wd = self.ws.copy(), self.ti.copy(), self.wd.copy()
        yaw0 = np.zeros((B, T))
        # Many envs at once: the same bit streams evaluated with array arithmetic (windgym_b200/fast_rng.py, pinned
        # against numpy in tests/test_host_logic.py) instead of one Generator object per env (17 us each).
        idx = np.fromiter(envs, dtype=np.int64)
        if (seed is not None and not (B == 1 and self._episode == 0) and self.sample_site is None
                and ec.turbtype not in ("MannGenerate", "MannLoad", "Random") and 0 <= int(seed) < 2 ** 32 and idx.size > 8
                and not getattr(self, "_no_fast_rng", False)):
            n_yaw = T if ec.yaw_init_mode == "Random" else 0
            u = uniform_streams(int(seed), idx, self._episode, 3 + n_yaw)
            ws[idx] = ec.ws_min + (ec.ws_max - ec.ws_min) * u[:, 0]      # Generator.uniform: low + (high - low) * u
            ti[idx] = ec.TI_min + (ec.TI_max - ec.TI_min) * u[:, 1]
            wd[idx] = ec.wd_min + (ec.wd_max - ec.wd_min) * u[:, 2]
            if n_yaw:
                yaw0[idx] = -ec.yaw_start + (ec.yaw_start - (-ec.yaw_start)) * u[:, 3:]
            envs = ()
        for i in envs:
            if seed is None:
                rng = np.random.default_rng()
            elif B == 1 and self._episode == 0:
                rng = np.random.default_rng(seed)
            else:
                rng = np.random.default_rng([seed, i, self._episode])
            if self.sample_site is None:
                ws[i] = rng.uniform(low=ec.ws_min, high=ec.ws_max)
                ti[i] = rng.uniform(low=ec.TI_min, high=ec.TI_max)
                wd[i] = rng.uniform(low=ec.wd_min, high=ec.wd_max)
            else:  # site-based sampling, Wind_Farm_Env.py:569-596 (sector frequency -> Weibull A, k -> clip)
                dirs, As, ks, freqs = self._site()
                sec = rng.choice(np.arange(dirs.size), 1, p=freqs)   # sector index (NOT the env index list `idx`)
                wd_s, ws_s = dirs[sec].item(), (As[sec] * rng.weibull(ks[sec])).item()
                wd[i] = np.clip(wd_s, ec.wd_min, ec.wd_max)
                ws[i] = np.clip(ws_s, ec.ws_min, ec.ws_max)
                ti[i] = rng.uniform(low=ec.TI_min, high=ec.TI_max)
            if ec.turbtype in ("MannGenerate", "Random"):    # TF_seed draw sits between wd and yaw (:623, :642)
                self._tf_seed[i] = int(rng.integers(0, 100000))
            elif ec.turbtype == "MannLoad":      # np_random.choice(TF_files) sits there too (:614): keep the stream aligned
                rng.choice(max(1, len(getattr(self, "_tf_files", ()))))
            if ec.yaw_init_mode == "Random":
                yaw0[i] = rng.uniform(low=-ec.yaw_start, high=ec.yaw_start, size=T)
        for k, arr in (("ws", ws), ("ti", ti), ("wd", wd)):
            if k in self._wind_override:
                arr[idx] = self._wind_override[k][idx]
        if ec.yaw_init_mode == "Defined":
            yv = np.asarray(self.yaw_initial, dtype=np.float64)
            if yv.size not in (1, T):
                raise ValueError("So I am pretty sure something has gone wrong here. The specified yaw values "
                                 "are not the right length.")
            yaw0[idx] = yv if yv.size == T else np.ones(T) * yv.reshape(-1)[0]
        return ws, ti, wd, yaw0

    def _site(self):
        """Wind resource tables of ``sample_site`` (py_wake site protocol: ``local_wind(x, y, wd, ws)`` with
        ``Sector_frequency_ilk``, ``Weibull_A_ilk``, ``Weibull_k_ilk``), read once (Wind_Farm_Env.py:571-577).
        The reference draws from the GLOBAL numpy RNG here; this class uses the env's seeded stream."""
        if self._site_tables is None:
            dirs = np.arange(0, 360, 1)
            lw = self.sample_site.local_wind(x=0, y=0, wd=dirs, ws=np.arange(3, 25, 1))
            freqs = np.asarray(lw.Sector_frequency_ilk)[0, :, 0].astype(np.float64)
            self._site_tables = (dirs, np.asarray(lw.Weibull_A_ilk)[0, :, 0].astype(np.float64),
                                 np.asarray(lw.Weibull_k_ilk)[0, :, 0].astype(np.float64), freqs / freqs.sum())
        return self._site_tables

    def reset(self, seed=None, mask=None, wind=None, yaw0=None, turb_offset=None):
        """WindFarmEnv.reset for the masked envs (all when ``mask`` is None).
        ``wind=(ws, ti, wd)`` and ``yaw0`` ([B,T]) inject conditions instead of sampling them; ``turb_offset``
        ([B,3] metres) pins the envs' positions inside the turbulence box."""
        ec, B, T, dev = self.ec, self.n_envs, self.n_turb, self.device
        sel = np.arange(B) if mask is None else np.flatnonzero(np.asarray(mask))
        if seed is None:
            seed = self.seed
        elif mask is None:
            # an explicit seed restarts the stream: later unseeded (and masked auto-) resets continue it, and
            # reset(seed=s) twice gives the same conditions (gymnasium check_env determinism)
            self.seed = seed
            self._episode = 0
        ws, ti, wd, y0 = self.sample_conditions(seed, sel)
        if wind is not None:
            for arr, v in zip((ws, ti, wd), wind):
                arr[sel] = np.broadcast_to(np.asarray(v, dtype=np.float64), (B,))[sel]
        if yaw0 is not None:
            y0[sel] = np.broadcast_to(np.asarray(yaw0, dtype=np.float64), (B, T))[sel]
        self.ws, self.ti, self.wd = ws, ti, wd
        # integer decisions of the reset, for the selected envs only (the kernel ignores the others)
        n_spin, k_emit = np.zeros(B, dtype=np.int32), np.ones(B, dtype=np.int32)
        prev_tm = getattr(self, "time_max", None)
        time_max = np.zeros(B, dtype=np.int32) if prev_tm is None or len(prev_tm) != B else np.array(prev_tm, dtype=np.int32)
        if sel.size:
            n_spin[sel], time_max[sel], k_emit[sel] = ec.reset_integers(ws[sel], wd[sel])
        rated = np.asarray(self.turbine.power(ws), dtype=np.float64)
        ti_flow = np.zeros(B)  # turbtype "None": RandomTurbulence(ti=0) (Wind_Farm_Env.py:661-665)
        tb = {}
        if self.turb_box is not None:
            ti_flow = ti.copy()    # TurbulenceFieldSite over a box scaled to the episode's TI (:617,:638,:657)
            if turb_offset is not None:
                self.turb_offset[sel] = np.broadcast_to(np.asarray(turb_offset, dtype=np.float64), (B, 3))[sel]
            elif B > 1:            # a lone env sits at the box origin like the reference's single simulation
                # one stream per env, [seed ^ salt, env, episode] -- evaluated for all envs of the reset at once
                # (fast_rng: numpy's SeedSequence + PCG64 with array arithmetic), not one Generator object per env
                L = np.array(self.turb_box.Nxyz) * np.array(self.turb_box.dxyz)
                s32 = ((0 if seed is None else int(seed)) ^ 0x5EEDB0C5 ^ (int(self._tf_seed[sel].sum()) * 2654435761)) & 0xFFFFFFFF
                self.turb_offset[sel] = uniform_streams(s32, sel, self._episode & 0xFFFFFFFF, 3) * L
            tb = dict(off=self.turb_offset, scale=self.turb_box.scale_for(ti, ws))
        f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(dev, non_blocking=True)
        i32 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32)).to(dev, non_blocking=True)
        keep = dict(ws=f32(ws), ti=f32(ti_flow), wd=f32(wd), yaw0=f32(y0), rated=f32(rated), k=i32(k_emit),
                    spin=i32(n_spin), tmax=i32(time_max), tb_off=f32(tb["off"]) if tb else None,
                    tb_scale=f32(tb["scale"]) if tb else None)
        m = None
        if mask is not None:
            m = torch.as_tensor(np.ascontiguousarray(np.asarray(mask), dtype=np.uint8)).to(dev)
            keep["mask"] = m
        args = _lib.ResetArgs(mask=_ptr(m), ws=_ptr(keep["ws"]), ti_flow=_ptr(keep["ti"]), wd=_ptr(keep["wd"]),
                              yaw0=_ptr(keep["yaw0"]), rated_power=_ptr(keep["rated"]), k_emit=_ptr(keep["k"]),
                              t_developed=_ptr(keep["spin"]), time_max=_ptr(keep["tmax"]),
                              tb_offset=_ptr(keep["tb_off"]), tb_scale=_ptr(keep["tb_scale"]))
        _lib.check(self.lib.wg_reset(self._h, _ptr(self._state), C.byref(args), _ptr(self.obs), self._stream()))
        self._keep = keep  # inputs stay alive until the stream has consumed them
        self._episode += 1
        self.time_max = time_max
        return self.obs, self._info()

    def step(self, actions, info=True):
        """WindFarmEnv.step for every env: actions float32 [B,T] (device) -> obs, reward, terminated, truncated, info
        (``info=False``: None instead of the dict of live device views, for wrappers that build their own)."""
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions, dtype=np.float32))
        if actions.device != self.device or actions.dtype != torch.float32 or not actions.is_contiguous():
            actions = actions.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
        if actions.numel() != self._n_act_elems:
            raise ValueError(f"actions must have {self._n_act_elems // (self.n_turb * self.ec.act_var)}x"
                             f"{self.n_turb * self.ec.act_var} elements")
        p = self._step_ptrs  # fixed buffers: the ctypes pointers are built once
        rc = self.lib.wg_step(self._h, p[0], actions.data_ptr(), p[1], p[2], p[3],
                              torch._C._cuda_getCurrentRawStream(self._dev_index))
        if rc != 0:
            _lib.check(rc)
        self._last_actions = actions   # keeps the tensor alive until the launch has consumed it
        return self.obs, self.reward, self.terminated, self.truncated, (self._info() if info else None)

    def step_host(self, actions):
        """``step()`` for callers whose buffers live on the HOST (the reference's contract: numpy in, numpy out):
        actions float32 [B, T*act_var] (numpy array or CPU tensor) in, numpy views out (obs [B,obs], reward [B],
        truncated bool [B]; valid until the next ``step_host`` call), returned when the results are in host memory.
        With pinned host buffers (``pin_memory()`` tensors; the result buffer always is) nothing is copied: the flow
        kernel reads the actions from mapped host memory, the finish kernel writes the results there and the host
        polls the step's completion word (``wg_step_host``).  ``info`` stays on the device (``self._info()``)."""
        if self._host is None:
            n_obs = int(np.prod(self.obs_shape))
            res = torch.empty(self._out.numel(), dtype=torch.uint8).pin_memory()
            act = torch.empty((self.n_envs, self.n_turb * self.ec.act_var), dtype=torch.float32).pin_memory()
            self._host = {
                "res": res, "act": act,
                "act_dev": torch.empty(act.shape, dtype=torch.float32, device=self.device),
                "obs": res[:n_obs * 4].view(torch.float32).view(self.obs_shape).numpy(),
                "reward": res[n_obs * 4:(n_obs + self.n_envs) * 4].view(torch.float32).numpy(),
                "truncated": res[(n_obs + self.n_envs) * 4:].numpy().view(np.bool_)}
            nb = C.c_size_t()
            _lib.check(self.lib.wg_result_bytes(self._h, C.byref(nb)))
            assert nb.value == self._out.numel(), "packed result buffer does not match the library's layout"
            self._host.update(act_ptr=_ptr(self._host["act_dev"]), out_ptr=_ptr(self._out),
                              res_ptr=C.c_void_p(res.data_ptr()), n_res=C.c_size_t(nb.value),
                              host_out=(self._host["obs"], self._host["reward"], self._host["truncated"]))
        h = self._host
        n_need = self._n_act_elems
        if (torch.is_tensor(actions) and actions.dtype == torch.float32 and actions.device.type == "cpu"
                and actions.is_contiguous() and actions.numel() == n_need):
            a_ptr = actions.data_ptr()       # fast path: the caller's buffer goes to the C-ABI as it is
        else:
            a = actions if torch.is_tensor(actions) else torch.from_numpy(np.ascontiguousarray(actions, dtype=np.float32))
            if a.numel() != n_need:
                raise ValueError(f"actions must have {n_need // (self.n_turb * self.ec.act_var)}x"
                                 f"{self.n_turb * self.ec.act_var} elements")
            stage = h["act"].view(-1)[:n_need]
            stage.copy_(a.reshape(-1))       # dtype / layout conversion into the pinned staging buffer
            a_ptr = stage.data_ptr()
        # one C-ABI call.  Pinned host buffers (the staging buffers here are; a caller's own pinned tensor too): the
        # kernels read the actions from and write the results to mapped host memory directly and the call returns when
        # the step's completion word arrives.  Pageable buffers: H2D copy, step, D2H copy, stream synchronise.
        rc = self.lib.wg_step_host(self._h, self._step_ptrs[0], a_ptr, h["act_ptr"], h["out_ptr"],
                                   h["res_ptr"], h["n_res"], torch._C._cuda_getCurrentRawStream(self._dev_index))
        if rc != 0:
            _lib.check(rc)
        return h["host_out"]

    def set_active(self, n_active):
        """``wg_set_active``: ``step()`` advances only envs [0, n_active); the other slots are a spare pool
        (``windgym_b200.pool.PooledVecEnv``)."""
        _lib.check(self.lib.wg_set_active(self._h, int(n_active)))
        self.n_active = int(n_active)
        self._n_act_elems = self.n_active * self.n_turb * self.ec.act_var

    def copy_envs(self, src, dst):
        """``wg_copy_envs``: complete per-env state of slot src[k] -> slot dst[k] (device copy on the current stream).
        The index lists go through one pinned staging buffer and one H2D copy; the observation rows follow."""
        src = np.asarray(src, dtype=np.int64).reshape(-1)
        dst = np.asarray(dst, dtype=np.int64).reshape(-1)
        n = int(src.size)
        if n == 0:
            return
        st = getattr(self, "_swap_stage", None)
        if st is None or st["cap"] < n:
            cap = max(64, 2 * n)
            st = {"cap": cap, "host": torch.empty(2 * cap, dtype=torch.int32).pin_memory(),
                  "dev": torch.empty(2 * cap, dtype=torch.int32, device=self.device), "ev": None}
            st["np"] = st["host"].numpy()
            self._swap_stage = st
        if st["ev"] is not None:
            st["ev"].synchronize()           # the previous lists have been read (normally long ago)
        st["np"][:n] = src
        st["np"][n:2 * n] = dst
        dev = st["dev"][:2 * n]
        dev.copy_(st["host"][:2 * n], non_blocking=True)
        st["ev"] = torch.cuda.Event()
        st["ev"].record(torch.cuda.current_stream(self.device))
        _lib.check(self.lib.wg_copy_envs(self._h, _ptr(self._state), C.c_void_p(dev.data_ptr()),
                                         C.c_void_p(dev.data_ptr() + 4 * n), n, self._stream()))
        idx = dev.long()
        self.obs.index_copy_(0, idx[n:], self.obs.index_select(0, idx[:n]))
        for arr in (self.ws, self.ti, self.wd, self.time_max):
            arr[dst] = arr[src]
        if getattr(self, "turb_box", None) is not None:
            self.turb_offset[dst] = self.turb_offset[src]

    def flow_steps(self, n):
        """DWMFlowSimulation.run(n*dt) for every env and farm, without measurement bookkeeping."""
        _lib.check(self.lib.wg_flow_steps(self._h, _ptr(self._state), int(n), self._stream()))

    def flow_field(self, x, y, z=None, env=0, farm=0):
        """``fs.get_windspeed(XYView(x, y, z), include_wakes=True)`` of one env's farm (render path,
        Wind_Farm_Env.py:470-476, :1056): wake-superposed (u, v, w) on the grid x[i] x y[j] at height z
        (default: hub height), wind-aligned frame.  Returns a device tensor [3, len(x), len(y)]."""
        xs = torch.as_tensor(np.asarray(x, dtype=np.float32), device=self.device)
        ys = torch.as_tensor(np.asarray(y, dtype=np.float32), device=self.device)
        gx, gy = torch.meshgrid(xs, ys, indexing="ij")
        px, py = gx.reshape(-1).contiguous(), gy.reshape(-1).contiguous()
        out = torch.empty((3, px.numel()), dtype=torch.float32, device=self.device)
        zz = float(self.ec.hub_height if z is None else z)
        _lib.check(self.lib.wg_flow_field(self._h, _ptr(self._state), int(env), int(farm), _ptr(px), _ptr(py),
                                          int(px.numel()), zz, _ptr(out), self._stream()))
        self._keep_field = (px, py)
        return out.reshape(3, xs.numel(), ys.numel())

    def mes_push_extract(self, ws, wd, yaw, power):
        """farm_mes.add_measurements + get_measurements(scaled=True) + clip on [B,T] device tensors."""
        ts = [x.to(self.device, dtype=torch.float32).contiguous() for x in (ws, wd, yaw, power)]
        _lib.check(self.lib.wg_mes_push_extract(self._h, _ptr(self._state), *[_ptr(x) for x in ts], _ptr(self.obs),
                                                self._stream()))
        self._keep_mes = ts
        return self.obs

    def check_flags(self):
        """Surface device-side error flags at a sync point (reference: Exception('NaN Power'), Wind_Farm_Env.py:981)."""
        fl = self.state["flags"]
        if bool((fl & 1).any()):
            raise Exception("NaN Power")
        if bool((fl & 2).any()):
            raise _lib.WgError("wake particle chain overflow (p_cap too small)")
        if bool((fl & 4).any()):
            raise _lib.WgError("work table overflow in wg_plan_kernel: farms were left unstepped (library bug)")

    def _info(self):
        """Device views with the reference's info keys (Wind_Farm_Env.py:527-555); zero-copy, no sync.  The views
        are live (they always show the current state), so the dict is built once and only the host-side wind
        conditions are refreshed."""
        d = getattr(self, "_info_views", None)
        if d is not None:
            d["Wind speed Global"], d["Wind direction Global"], d["Turbulence intensity"] = self.ws, self.wd, self.ti
            return dict(d)
        s = self.state
        d = {
            "yaw angles agent": s["yaw"][:, 0], "Wind speed Global": self.ws, "Wind direction Global": self.wd,
            "Turbulence intensity": self.ti, "Power pr turbine agent": s["power"][:, 0],
            "Wind speed at turbines": s["meas"][:, 0], "Wind direction at turbines": s["meas"][:, 1],
            "Turbine x positions": s["xr"], "Turbine y positions": s["yr"],
        }
        if self.ec.act_var == 2:
            d["derating agent"] = s["derate"][:, 0]
        if self.Baseline_comp:
            d["yaw angles base"] = s["yaw"][:, 1]
            d["Power pr turbine baseline"] = s["power"][:, 1]
            d["Wind speed at turbines baseline"] = s["u"][:, 1]
        self._info_views = d
        return dict(d)

    # helpers for tests / inspection ------------------------------------------------------------------------
    def profiles_by_age(self, b, f, t):
        """Un-swizzled wake profiles [count,64] and scalars of one chain, youngest first (host numpy)."""
        s = self.state
        P = self.ec.p_cap
        head, cnt = int(s["head"][b, f, t]), int(s["count"][b, f, t])
        par = int(s["n_step"][b, f]) & 1
        slots = (head - 1 - np.arange(cnt)) % P
        raw = s["prof"][b, f, t].cpu().numpy()[slots].reshape(cnt, 16, 4)
        key = (slots & 7)[:, None]
        idx = np.arange(16)[None, :] ^ key
        prof = np.take_along_axis(raw, idx[:, :, None], axis=1).reshape(cnt, 64)
        # node 63 is the Dirichlet node (U = 1); its slot carries the row's shear integral sqrt(2 M (1 - Umin))
        self.last_bw = prof[:, 63].copy()
        prof[:, 63] = 1.0
        pmut = s["pmut"][par, b, f, t].cpu().numpy()[slots]
        pcon = s["pcon"][b, f, t].cpu().numpy()[slots]
        return prof, pmut, pcon

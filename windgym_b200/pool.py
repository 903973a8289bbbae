"""Auto-reset without a stall: a pool of pre-developed spare envs behind a ``VecWindFarmEnv``.

Resetting a wind-farm env is expensive by nature -- ``fs.run(t_developed)`` plus the measurement fill is 30-40 % of
all flow steps of an episode (reference ``Wind_Farm_Env.py:722-766``) -- and in a batch of
thousands of envs some episode ends at almost every step.  A masked reset inside ``step()`` would serialise a ~300
step spin-up of a handful of envs behind every batched step (measured: 13.7 ms per step instead of 0.45 ms).
``PooledVecEnv`` keeps ``reserve`` extra env slots in the same state tensor: they are reset in the background on a
second CUDA stream (batched, overlapping the stepping of the active envs); when an episode ends, a ready spare is
copied over the finished env (``wg_copy_envs``, ~1 MB per env) and the freed spare is recycled.  Conditions of a
spare are drawn when it is prepared instead of at the moment of the swap: the same distribution and, for a given
seed, a deterministic sequence.  If the pool runs dry the remaining envs take the synchronous masked reset.

The object has the ``VecWindFarmEnv`` protocol (the vector adapters of ``windgym_b200.vector`` accept it as
``venv=``); arrays and state views cover the ``n_envs`` active envs only.
"""
import numpy as np
import torch

from .vec_env import VecWindFarmEnv


class PooledVecEnv:
    def __init__(self, turbine, n_envs, reserve=None, refill_chunk=None, **env_kwargs):
        self.n_envs = int(n_envs)
        self.reserve = int(reserve) if reserve is not None else max(32, self.n_envs // 8)
        self.inner = VecWindFarmEnv(turbine, self.n_envs + self.reserve, **env_kwargs)
        v = self.inner
        if len(v.obs_shape) != 2:
            raise ValueError("PooledVecEnv wraps the single-agent observation layout")
        v.set_active(self.n_envs)
        # Spares are recycled in batches of reserve/4 right after a step has been launched (the host side of a masked
        # reset -- condition draws, argument upload, ~50 launches -- then overlaps with the GPU stepping the active
        # envs), on a ring of background streams: one spin-up is ~300 serial flow steps, i.e. tens of ms whatever the
        # batch size, so the supply of ready spares is (envs in preparation) / latency and refills must overlap.
        self.refill_chunk = int(refill_chunk) if refill_chunk is not None else max(1, self.reserve // 4)
        self.device, self.ec, self.n_turb, self.obs_var = v.device, v.ec, v.n_turb, v.obs_var
        self.obs_shape = (self.n_envs, v.obs_var)
        self.Baseline_comp, self.n_farms = v.Baseline_comp, v.n_farms
        self.yaw_min, self.yaw_max, self.yaw_step = v.yaw_min, v.yaw_max, v.yaw_step
        self.x_pos, self.y_pos = v.x_pos, v.y_pos
        self._bgs = [torch.cuda.Stream(device=self.device) for _ in range(4)]
        self._bg_next = 0
        self._ready, self._refilling, self._free = [], [], []      # [(slot, event)], [(slots, event, keep)], [slot]
        self.stats = {"swapped": 0, "sync_resets": 0, "refills": 0}
        B = self.n_envs
        self.state = {k: (t[:, :B] if k == "pmut" else t[:B]) for k, t in v.state.items()}
        self.terminated = v.terminated[:B]

    # ------------------------------------------------------------------------------------------ protocol
    @property
    def seed(self):
        return self.inner.seed

    @seed.setter
    def seed(self, value):
        self.inner.seed = value

    @property
    def ws(self):
        return self.inner.ws[:self.n_envs]

    @property
    def ti(self):
        return self.inner.ti[:self.n_envs]

    @property
    def wd(self):
        return self.inner.wd[:self.n_envs]

    @property
    def time_max(self):
        return self.inner.time_max[:self.n_envs]

    @property
    def obs(self):
        return self.inner.obs[:self.n_envs]

    @property
    def launch_count(self):
        return self.inner.launch_count

    def set_wind_vals(self, ws=None, ti=None, wd=None):
        for v in (ws, ti, wd):
            if v is not None and np.ndim(v) != 0:
                raise ValueError("PooledVecEnv pins scalar wind values only (spares and active envs share them)")
        self.inner.set_wind_vals(ws=ws, ti=ti, wd=wd)

    def check_flags(self):
        fl = self.state["flags"]
        if bool((fl & 1).any()):
            raise Exception("NaN Power")
        self.inner.check_flags()

    def _info(self):
        B = self.n_envs
        d = getattr(self, "_info_views", None)
        if d is None:   # device views are live: slice them once, refresh only the host-side wind conditions
            d = {k: (v[:B] if hasattr(v, "shape") and len(v.shape) and v.shape[0] == B + self.reserve else v)
                 for k, v in self.inner._info().items()}
            self._info_views = d
        d["Wind speed Global"], d["Wind direction Global"], d["Turbulence intensity"] = self.ws, self.wd, self.ti
        return dict(d)

    def close(self):
        torch.cuda.synchronize(self.device)
        self.inner.close()

    # ------------------------------------------------------------------------------------------ reset / step
    def reset(self, seed=None, mask=None, wind=None, yaw0=None):
        """``mask=None``: reset every active env (synchronously) and start preparing the spares.  ``mask`` ([n_envs]
        bool): the envs whose episode ended -- ready spares are swapped in."""
        B, R = self.n_envs, self.reserve
        if mask is None:
            torch.cuda.synchronize(self.device)                     # no background work may touch the slots we reuse
            self._ready, self._refilling, self._free = [], [], []
            m = np.zeros(B + R, dtype=bool)
            m[:B] = True
            if wind is not None:
                wind = tuple(np.concatenate([np.broadcast_to(np.asarray(w, dtype=np.float64), (B,)), np.zeros(R)]) for w in wind)
            if yaw0 is not None:
                yaw0 = np.concatenate([np.broadcast_to(np.asarray(yaw0, dtype=np.float64), (B, self.n_turb)),
                                       np.zeros((R, self.n_turb))])
            if seed is not None:   # spares prepared later continue the seeded stream
                self.inner.seed = seed
                self.inner._episode = 0
            self.inner.reset(seed=seed, mask=m, wind=wind, yaw0=yaw0)
            self._refill(list(range(B, B + R)))
            return self.obs, self._info()
        idx = np.flatnonzero(np.asarray(mask)[:B])
        if idx.size:
            self._swap_in(idx)
        return self.obs, self._info()

    def step(self, actions):
        obs, rew, term, trunc, _ = self.inner.step(actions)
        if len(self._free) >= self.refill_chunk:                   # behind the launch: see refill_chunk
            self._refill(self._free)
            self._free = []
        B = self.n_envs
        return obs[:B], rew[:B], term[:B], trunc[:B], self._info()

    # ------------------------------------------------------------------------------------------ the pool
    def _refill(self, slots):
        """Batched masked reset of spare ``slots`` on the background stream."""
        if not slots:
            return
        main = torch.cuda.current_stream(self.device)
        done_reading = torch.cuda.Event()
        done_reading.record(main)                                   # copies out of these slots are ordered before
        m = np.zeros(self.n_envs + self.reserve, dtype=bool)
        m[slots] = True
        bg = self._bgs[self._bg_next]
        self._bg_next = (self._bg_next + 1) % len(self._bgs)
        with torch.cuda.stream(bg):
            bg.wait_event(done_reading)
            self.inner.reset(mask=m)
            keep = self.inner._keep
            ev = torch.cuda.Event()
            ev.record(bg)
        self._refilling.append((list(slots), ev, keep))
        self.stats["refills"] += 1

    def _collect(self, block=False):
        """Move finished background batches to the ready list (``block``: if none has finished take the oldest one
        anyway -- the main stream then waits for it on the device, the host does not)."""
        done = [r for r in self._refilling if r[1].query()]
        if not done and block and self._refilling:
            done = [self._refilling[0]]
        for r in done:
            self._refilling.remove(r)
            self._ready += [(s, r[1]) for s in r[0]]

    def _swap_in(self, idx):
        self._collect()
        while len(self._ready) < len(idx) and self._refilling:
            self._collect(block=True)
        take = min(len(idx), len(self._ready))
        main = torch.cuda.current_stream(self.device)
        if take:
            pairs = [self._ready.pop() for _ in range(take)]
            for ev in {id(e): e for _, e in pairs}.values():
                main.wait_event(ev)
            src = [s for s, _ in pairs]
            self.inner.copy_envs(src, idx[:take])
            self._free += src
            self.stats["swapped"] += take
        if take < len(idx):                                          # pool ran dry: synchronous masked reset
            m = np.zeros(self.n_envs + self.reserve, dtype=bool)
            m[idx[take:]] = True
            self.inner.reset(mask=m)
            self.stats["sync_resets"] += len(idx) - take


class DevicePooledVecEnv:
    """Auto-reset with the host out of the loop (``wg_pool_*`` in ``include/windgym_b200.h``).

    Same idea as ``PooledVecEnv`` -- ``reserve`` spare env slots behind the ``n_envs`` active ones, spun up in the
    background and copied over envs whose episode ended -- but every per-step decision is made on the device:
    ``wg_pool_swap`` (two launches on the stepping stream) pairs finished episodes with ready spares and copies them,
    ``wg_pool_refill`` (whenever enough spares have been consumed, round robin over background streams) draws new wind conditions
    for the consumed spares with a counter-based generator, runs the masked reset on them and marks them ready.  No
    truncation flags are read back, no host-side bookkeeping per episode.  ``step()`` returns the new episode's first
    observation for the swapped envs and ``truncated`` = the envs that were swapped in this step (an episode that
    finds no ready spare runs on until one is; ``stats["deferred"]`` counts those env-steps); the finished episode's
    last observation is in ``final_obs``.

    Conditions come from the device generator (``seed``, slot, refill count) -- the same distributions as the
    reference's draws (Wind_Farm_Env.py:557-568, :715), not numpy's PCG64 stream; the first episode of every active env
    is a host-seeded ``VecWindFarmEnv.reset`` as before.  Works for the single-agent and the multi-agent observation
    layout (obs [n_envs, obs] or [n_envs, T, obs])."""

    device_autoreset = True

    def __init__(self, turbine, n_envs, reserve=None, refill_every=4, n_streams=8, **env_kwargs):
        import ctypes as C
        from . import _lib
        self._C, self._lib_mod = C, _lib
        self.n_envs = int(n_envs)
        # A spin-up takes ~20 ms whatever the batch size while episodes end at a rate of n_envs / episode length per
        # step: the spares in flight are (finished episodes per second) x (spin-up latency) ~ 250-400 for the 4x4 farm
        # at any batch size from 512 envs up; small batches are bounded by 4 spares per env
        self.reserve = int(reserve) if reserve is not None else max(32, min(384, 4 * self.n_envs), self.n_envs // 8)
        self.refill_every = max(1, int(refill_every))      # steps between attempts to issue a refill (see _refill)
        self.inner = VecWindFarmEnv(turbine, self.n_envs + self.reserve, **env_kwargs)
        v = self.inner
        if v.sample_site is not None:
            raise NotImplementedError("DevicePooledVecEnv draws uniform wind conditions on the device; use "
                                      "PooledVecEnv (host-side draws) with sample_site")
        self.device, self.ec, self.n_turb, self.obs_var = v.device, v.ec, v.n_turb, v.obs_var
        self.obs_shape = (self.n_envs,) + tuple(v.obs_shape[1:])
        self.Baseline_comp, self.n_farms = v.Baseline_comp, v.n_farms
        self.yaw_min, self.yaw_max, self.yaw_step = v.yaw_min, v.yaw_max, v.yaw_step
        self.x_pos, self.y_pos = v.x_pos, v.y_pos
        self.lib = v.lib
        n_streams = min(max(1, int(n_streams)), 8)          # one refill mask row per stream (WG_POOL_MASKS)
        # HIGH priority: a refill is a few dozen long-lived CTAs; behind a stepping grid that always has thousands of CTAs
        # queued they would wait for slots and the spin-up latency (hence the number of spares in flight) would grow
        # without bound -- with priority they take their slots at once and the stepping grid fills the rest
        self._bgs = [torch.cuda.Stream(device=self.device, priority=-1) for _ in range(n_streams)]
        self._bg_events = [None] * n_streams
        self._n_refill = 0
        self._n_tried = 0
        self._steps = 0
        B = self.n_envs
        self.state = {k: (t[:, :B] if k == "pmut" else t[:B]) for k, t in v.state.items()
                      if t.shape[0] == B + self.reserve or (k == "pmut" and t.shape[1] == B + self.reserve)}
        self.terminated = v.terminated[:B]
        self.swapped = torch.zeros(B + self.reserve, dtype=torch.uint8, device=self.device)
        self._final = torch.zeros_like(v.obs)
        self._ready = False
        self._swap_ptrs = (v._step_ptrs[0], v._step_ptrs[3], v._step_ptrs[1], C.c_void_p(self.swapped.data_ptr()),
                           C.c_void_p(self._final.data_ptr()))
        self._views = (v.obs[:B], v.reward[:B], v.terminated[:B], self.swapped[:B])

    # ------------------------------------------------------------------------------------------ protocol
    @property
    def seed(self):
        return self.inner.seed

    @seed.setter
    def seed(self, value):
        self.inner.seed = value

    @property
    def ws(self):
        return self.state["ws"]

    @property
    def ti(self):
        return self.state["ti"]

    @property
    def wd(self):
        return self.state["wd"]

    @property
    def time_max(self):
        return self.state["time_max"]

    @property
    def obs(self):
        return self.inner.obs[:self.n_envs]

    @property
    def final_obs(self):
        """Last observation of the episodes that ended in the latest step (rows where ``truncated`` is set)."""
        return self._final[:self.n_envs]

    @property
    def launch_count(self):
        return self.inner.launch_count

    @property
    def stats(self):
        """{"swapped", "deferred", "refilled"} so far (synchronises the device)."""
        out = (self._C.c_uint64 * 8)()
        self._lib_mod.check(self.lib.wg_pool_stats(self.inner._h, self.inner._step_ptrs[0], out, self.inner._stream()))
        return {"swapped": int(out[0]), "deferred": int(out[1]), "refilled": int(out[2]), "refill_calls": self._n_refill}

    def check_flags(self):
        fl = self.state["flags"]
        if bool((fl & 1).any()):
            raise Exception("NaN Power")
        if bool((fl & 2).any()):
            raise self._lib_mod.WgError("wake particle chain overflow (p_cap too small)")

    def _info(self):
        d = getattr(self, "_info_views", None)
        if d is None:
            B, n_all = self.n_envs, self.n_envs + self.reserve
            d = {k: (val[:B] if hasattr(val, "shape") and len(val.shape) and val.shape[0] == n_all else val)
                 for k, val in self.inner._info().items()}
            # the wind conditions live on the device: they change whenever a spare is swapped in
            d["Wind speed Global"], d["Wind direction Global"], d["Turbulence intensity"] = self.ws, self.wd, self.ti
            self._info_views = d
        return dict(d)

    def close(self):
        torch.cuda.synchronize(self.device)
        self.inner.close()

    # ------------------------------------------------------------------------------------------ reset / step
    def _draw(self, seed):
        ec, v = self.ec, self.inner
        yc = 0.0
        if ec.yaw_init_mode == "Defined":
            yv = np.asarray(v.yaw_initial, dtype=np.float64).reshape(-1)
            if yv.size != 1:
                raise NotImplementedError("per-turbine 'Defined' yaw values need PooledVecEnv (host-side resets)")
            yc = float(yv[0])
        return self._lib_mod.PoolDraw(
            ws_min=ec.ws_min, ws_max=ec.ws_max, ti_min=ec.TI_min, ti_max=ec.TI_max, wd_min=ec.wd_min, wd_max=ec.wd_max,
            yaw_start=ec.yaw_start, n_passthrough=float(ec.n_passthrough),
            tb_std_u=float(v.turb_box.std_u) if v.turb_box is not None else 0.0, yaw_const=yc,
            yaw_random=int(ec.yaw_init_mode == "Random"), eval_mode=int(bool(ec.eval_mode)),
            seed=int(0 if seed is None else seed) & 0xFFFFFFFFFFFFFFFF)

    def reset(self, seed=None, mask=None, wind=None, yaw0=None):
        """``mask=None``: reset every active env (host-seeded, like ``VecWindFarmEnv.reset``) and start preparing the
        spares.  ``mask`` ([n_envs] bool): synchronous masked reset of those envs (the pool itself never needs it)."""
        v, B, R = self.inner, self.n_envs, self.reserve
        torch.cuda.synchronize(self.device)                     # no background work may touch the slots we reuse
        m = np.zeros(B + R, dtype=bool)
        if mask is None:
            m[:B] = True
        else:
            m[:B] = np.asarray(mask, dtype=bool)[:B]
        if wind is not None:
            wind = tuple(np.concatenate([np.broadcast_to(np.asarray(w, dtype=np.float64), (B,)), np.zeros(R)]) for w in wind)
        if yaw0 is not None:
            yaw0 = np.concatenate([np.broadcast_to(np.asarray(yaw0, dtype=np.float64), (B, self.n_turb)),
                                   np.zeros((R, self.n_turb))])
        if seed is not None:
            v.seed, v._episode = seed, 0
        v.reset(seed=seed, mask=m, wind=wind, yaw0=yaw0)
        if mask is None:
            self._lib_mod.check(self.lib.wg_pool_init(v._h, v._step_ptrs[0], B, v._stream()))
            v.n_active = B
            v._n_act_elems = B * self.n_turb * self.ec.act_var
            self._draw_args = self._draw(v.seed)
            self._steps, self._ready = 0, True
            self._refill()                                       # all spares: one batched spin-up
        return self.obs, self._info()

    def _refill(self, wait=False):
        """Issue one refill sequence (claim + draw, masked reset, publish: ~60 launches) on a background stream that is
        idle.  A spin-up takes ~20 ms whatever it holds, so the pace is set by the streams, not by the step count: at
        most one sequence is queued per stream (its completion event is polled, never waited for), what a sequence
        claims is decided on the device when it starts.  No idle stream: nothing is issued (the next attempt is a few
        steps away)."""
        v = self.inner
        for _ in range(len(self._bgs)):
            k = self._n_tried % len(self._bgs)
            self._n_tried += 1
            ev = self._bg_events[k]
            if ev is None or ev.query():
                bg = self._bgs[k]
                self._lib_mod.check(self.lib.wg_pool_refill(v._h, v._step_ptrs[0], self._C.byref(self._draw_args),
                                                            v._step_ptrs[1], k, self._C.c_void_p(bg.cuda_stream)))
                ev = torch.cuda.Event()
                ev.record(bg)
                self._bg_events[k] = ev
                self._n_refill += 1
                return True
        return False

    def step(self, actions):
        if not self._ready:
            raise RuntimeError("reset() must be called before step()")
        v = self.inner
        v.step(actions, info=False)
        sp = self._swap_ptrs
        rc = self.lib.wg_pool_swap(v._h, sp[0], sp[1], sp[2], sp[3], sp[4], v._stream())
        if rc != 0:
            self._lib_mod.check(rc)
        self._steps += 1
        if self._steps % self.refill_every == 0:
            self._refill()
        o = self._views          # fixed buffers: the sliced views are built once
        return o[0], o[1], o[2], o[3], self._info()

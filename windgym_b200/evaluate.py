"""Batched agent evaluation.

The reference evaluates ONE condition per run (``eval_single_fast``, ``WindGym/AgentEval.py:39-477``) and loops
serially over (wd, ws, TI, turbulence box) in ``AgentEval.eval_multiple`` (:579-617) before ``xr.merge``.  Here the
whole condition grid is the env batch of one ``VecWindFarmEnv`` in evaluation mode (``FarmEval`` semantics:
pinned wind, pinned initial yaw, no truncation; ``FarmEval.py:10-90``): one reset, ``t_sim - 1`` batched steps, the
per-step records accumulated on the device and read back once.  With ``torch.distributed`` initialised the
conditions are sharded across ranks and assembled with the ONE gather of the design (``sharding.gather_env_stats``).

The result has the reference's variables and dimension order
    powerF_a, reward, pct_inc, powerF_b : (time, ws, wd, TI, turbbox, model_step)
    powerT_a, yaw_a, ws_a (+ *_b)       : (time, turb, ws, wd, TI, turbbox, model_step)
as plain numpy arrays in an ``EvalDataset`` (``to_xarray()`` when xarray is installed, ``save()`` to ``.npz``).
"""
import itertools

import numpy as np
import torch
import torch.distributed as dist

from .agents import batch_actions
from .sharding import gather_env_stats, shard_range

_DIMS6 = ("time", "ws", "wd", "TI", "turbbox", "model_step")
_DIMS7 = ("time", "turb", "ws", "wd", "TI", "turbbox", "model_step")


class EvalDataset:
    """Minimal labelled-array container mirroring the reference's ``xr.Dataset`` layout."""

    def __init__(self, data_vars, coords):
        self.data_vars, self.coords = data_vars, coords

    def __getitem__(self, name):
        return self.data_vars[name][1]

    def __contains__(self, name):
        return name in self.data_vars

    def dims(self, name):
        return self.data_vars[name][0]

    def to_xarray(self):
        import xarray as xr  # optional
        return xr.Dataset(data_vars=self.data_vars, coords=self.coords)

    def save(self, path):
        """``AgentEval.save_performance`` (:699-707) without xarray: one ``.npz`` with ``var/<name>`` and ``coord/<name>``."""
        out = {f"var/{k}": v[1] for k, v in self.data_vars.items()}
        out.update({f"coord/{k}": np.asarray(v) for k, v in self.coords.items()})
        np.savez_compressed(path, **out)

    @classmethod
    def load(cls, path):
        z = np.load(path, allow_pickle=False)
        coords = {k[6:]: z[k] for k in z.files if k.startswith("coord/")}
        dv = {}
        for k in z.files:
            if k.startswith("var/"):
                a = z[k]
                dv[k[4:]] = (_DIMS7 if a.ndim == 7 else _DIMS6, a)
        return cls(dv, coords)


def condition_grid(windspeeds, winddirs, turbintensities, turbboxes):
    """Flat condition list in (ws, wd, TI, turbbox) C order -- the order the result arrays are reshaped with."""
    return list(itertools.product(windspeeds, winddirs, turbintensities, turbboxes))


def eval_batched(env, model, windspeeds=(10.0,), winddirs=(270,), turbintensities=(0.05,), turbboxes=("Default",),
                 yaw_init=0.0, t_sim=1000, model_step=1, deterministic=False, distributed=None):
    """Evaluate ``model`` on every (ws, wd, TI, turbbox) combination at once.

    ``env`` is a ``VecWindFarmEnv`` built with ``eval_mode=True`` (never truncates) whose batch holds this rank's
    share of the conditions: ``n_envs == len(conditions)`` on one GPU, or ``shard_range(len(conditions), rank, world)``
    envs per rank when ``torch.distributed`` is initialised (``distributed=None`` auto-detects).
    Mirrors ``eval_single_fast`` (:119-209): record t = 0 after reset, then ``t_sim - 1`` steps of
    ``action = model.predict(obs)`` / ``env.step(action)``; values recorded are post-step ``power()``, ``yaw``,
    ``|rotor_avg_windspeed|`` of the agent farm (and the baseline farm when ``Baseline_comp``), the reward and
    ``pct_inc = (P_a - P_b)/P_b*100``.
    """
    conds = condition_grid(windspeeds, winddirs, turbintensities, turbboxes)
    if any(b != "Default" for _, _, _, b in conds) and getattr(env.ec, "turbtype", "None") == "None":
        raise NotImplementedError("named turbulence boxes need a Mann-box site (turbtype != 'None')")
    if len(set(turbboxes)) > 1:
        # every env of a handle samples the handle's ONE shared box: a turbbox axis with several names would hold
        # identical results (the reference loads a box per entry, AgentEval.py:110-118) -- refuse instead
        raise NotImplementedError("eval_batched evaluates one turbulence box per call: run it once per box "
                                  "(env.update_tf / turb_box=) and concatenate along 'turbbox'")
    use_dist = dist.is_initialized() and dist.get_world_size() > 1 if distributed is None else bool(distributed)
    rank, world = (dist.get_rank(), dist.get_world_size()) if use_dist else (0, 1)
    lo, hi = shard_range(len(conds), rank, world)
    if env.n_envs != hi - lo:
        raise ValueError(f"env batch ({env.n_envs}) must equal this rank's share of the conditions ({hi - lo})")
    mine = conds[lo:hi]
    B, T, dev = env.n_envs, env.n_turb, env.device
    base = bool(env.Baseline_comp)
    if hasattr(model, "UseEnv"):  # eval_single_fast :124-128
        model.yaw_max, model.yaw_min, model.env = env.yaw_max, env.yaw_min, env
    env.set_wind_vals(ws=np.array([c[0] for c in mine], dtype=np.float64), ti=np.array([c[2] for c in mine], dtype=np.float64),
                      wd=np.array([c[1] for c in mine], dtype=np.float64))
    y0 = np.broadcast_to(np.asarray(yaw_init, dtype=np.float64), (T,)) if np.ndim(yaw_init) <= 1 else np.asarray(yaw_init)
    obs, _ = env.reset(yaw0=np.broadcast_to(y0, (B, T)))

    n_rec = 3 * T + 1 + (3 * T if base else 0)   # per env and time: power[T], yaw[T], |uvw|[T], reward (+ baseline)
    rec = torch.zeros((t_sim, B, n_rec), dtype=torch.float32, device=dev)

    def record(i, reward):
        s = env.state
        wsn = torch.sqrt(s["u"] ** 2 + s["v"] ** 2 + s["w"] ** 2)
        parts = [s["power"][:, 0], s["yaw"][:, 0], wsn[:, 0], reward.reshape(B, 1)]
        if base:
            parts += [s["power"][:, 1], s["yaw"][:, 1], wsn[:, 1]]
        rec[i] = torch.cat(parts, dim=1)

    record(0, torch.zeros(B, device=dev))
    time0 = env.state["n_step"][:, 0].to(torch.float32) * float(env.ec.dt_sim)
    for i in range(1, t_sim):
        act = batch_actions(model, obs, env, deterministic=deterministic)
        obs, reward, _, _, _ = env.step(act)
        record(i, reward)
    env.check_flags()

    # [B, t_sim, n_rec] rows per env -> the one gather of the multi-GPU design -> global condition order
    rows = rec.permute(1, 0, 2).contiguous()
    if use_dist:
        rows = gather_env_stats(rows, len(conds))
        time0 = gather_env_stats(time0.reshape(-1, 1), len(conds)).reshape(-1)
    rows, time0 = rows.cpu().numpy().astype(np.float64), time0.cpu().numpy().astype(np.float64)

    shape = (len(windspeeds), len(winddirs), len(turbintensities), len(turbboxes))

    def farm(a):   # [C, time] -> (time, ws, wd, TI, box, 1)
        return np.moveaxis(a.reshape(shape + (t_sim,)), -1, 0)[..., None]

    def turb(a):   # [C, time, T] -> (time, turb, ws, wd, TI, box, 1)
        return np.moveaxis(np.moveaxis(a.reshape(shape + (t_sim, T)), -2, 0), -1, 1)[..., None]

    pT, yw, wsn, rw = rows[:, :, :T], rows[:, :, T:2 * T], rows[:, :, 2 * T:3 * T], rows[:, :, 3 * T]
    dv = {"powerF_a": (_DIMS6, farm(pT.sum(-1))), "powerT_a": (_DIMS7, turb(pT)), "yaw_a": (_DIMS7, turb(yw)),
          "ws_a": (_DIMS7, turb(wsn)), "reward": (_DIMS6, farm(rw))}
    if base:
        o = 3 * T + 1
        pTb, ywb, wsb = rows[:, :, o:o + T], rows[:, :, o + T:o + 2 * T], rows[:, :, o + 2 * T:o + 3 * T]
        pa, pb = pT.sum(-1), pTb.sum(-1)
        dv.update({"powerF_b": (_DIMS6, farm(pb)), "powerT_b": (_DIMS7, turb(pTb)), "yaw_b": (_DIMS7, turb(ywb)),
                   "ws_b": (_DIMS7, turb(wsb)), "pct_inc": (_DIMS6, farm((pa - pb) / pb * 100.0))})
    coords = {"ws": np.asarray(windspeeds, dtype=np.float64), "wd": np.asarray(winddirs, dtype=np.float64),
              "turb": np.arange(T), "time": np.arange(t_sim) * float(env.ec.dt_env),
              "TI": np.asarray(turbintensities, dtype=np.float64), "turbbox": np.asarray(list(turbboxes)),
              "model_step": np.array([model_step]),
              # fs.time right after reset per condition (the reference's time coordinate is time0 + time)
              "time0": time0.reshape(shape)}
    return EvalDataset(dv, coords)


class AgentEval:
    """``WindGym.AgentEval`` (:480-715) on the batched backend: same setters and ``eval_single`` / ``eval_multiple``
    / ``save_performance`` / ``load_performance``; ``env_factory(n_envs)`` builds the evaluation batch
    (``VecWindFarmEnv(..., eval_mode=True)``) for the requested number of conditions."""

    def __init__(self, env_factory=None, model=None, name="NoName", t_sim=1000):
        self.ws, self.ti, self.wd, self.yaw, self.turbbox = 10.0, 0.05, 270, 0.0, "Default"
        self.t_sim = t_sim
        self.winddirs, self.windspeeds, self.turbintensities, self.turbboxes = [270], [10], [0.05], ["Default"]
        self.multiple_eval = False
        self.env_factory, self.model, self.name = env_factory, model, name

    def set_conditions(self, winddirs=(), windspeeds=(), turbintensities=(), turbboxes=("Default",)):
        if len(winddirs):
            self.winddirs = list(winddirs)
        if len(windspeeds):
            self.windspeeds = list(windspeeds)
        if len(turbintensities):
            self.turbintensities = list(turbintensities)
        if len(turbboxes):
            self.turbboxes = list(turbboxes)

    def set_condition(self, ws=None, ti=None, wd=None, yaw=None, turbbox=None):
        for k, v in (("ws", ws), ("ti", ti), ("wd", wd), ("yaw", yaw), ("turbbox", turbbox)):
            if v is not None:
                setattr(self, k, v)

    def update_model(self, model):
        self.model = model

    def _run(self, wss, wds, tis, boxes, deterministic):
        n = len(wss) * len(wds) * len(tis) * len(boxes)
        if dist.is_initialized() and dist.get_world_size() > 1:
            lo, hi = shard_range(n, dist.get_rank(), dist.get_world_size())
            n = hi - lo
        env = self.env_factory(n)
        try:
            return eval_batched(env, self.model, wss, wds, tis, boxes, yaw_init=self.yaw, t_sim=self.t_sim,
                                deterministic=deterministic)
        finally:
            env.close()

    def eval_single(self, deterministic=False, **_):
        return self._run([self.ws], [self.wd], [self.ti], [self.turbbox], deterministic)

    def eval_multiple(self, deterministic=False, **_):
        self.multiple_eval = True
        self.multiple_eval_ds = self._run(self.windspeeds, self.winddirs, self.turbintensities, self.turbboxes,
                                          deterministic)
        return self.multiple_eval_ds

    def save_performance(self, path=None):
        if self.multiple_eval:
            self.multiple_eval_ds.save(path or (self.name + "_eval.npz"))
        else:
            print("It doenst look like you have any data to save my guy")

    def load_performance(self, path):
        self.multiple_eval_ds = EvalDataset.load(path)
        self.multiple_eval = True

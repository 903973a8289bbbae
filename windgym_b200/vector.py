"""Vector-env adapters over ``VecWindFarmEnv``.

The reference vectorises by process replication (SB3 ``SubprocVecEnv`` / ``make_vec_env``,
``examples/longer_steps_example.py:194-219``; gymnasium vector envs wrapped by ``RecordEpisodeVals``,
``WindGym/wrappers/recordEpisodeVals.py:8-64``).  Here the batch is one CUDA launch; these classes only give it the
call surface those callers expect:

* ``GymVectorEnv``  -- ``gymnasium.vector.VectorEnv`` protocol: ``reset(seed, options) -> (obs, infos)``,
  ``step(actions) -> (obs, rewards, terminations, truncations, infos)`` with dict-of-arrays infos
  (``infos["Power agent"]`` is the array ``RecordEpisodeVals`` reads, wrappers :44), same-step auto-reset with
  ``infos["final_observation"]`` / ``infos["_final_observation"]``.
* ``SB3VecEnv``     -- stable-baselines3 ``VecEnv`` protocol: ``reset() -> obs``, ``step_async/step_wait`` ->
  ``(obs, rewards, dones, infos)`` with a list of per-env dicts carrying ``terminal_observation`` and
  ``TimeLimit.truncated``; ``get_attr/set_attr/env_method/env_is_wrapped/seed/close``.
* ``RecordEpisodeVals`` -- episode return / length / mean-power queues with the reference wrapper's names.

``as_torch=True`` keeps observations, rewards and flags on the device (no host copies) for on-device policies.
gymnasium / stable-baselines3 are not dependencies: the classes are duck-typed.
"""
import time
from collections import deque

import numpy as np
import torch

from .envs import Box
from .vec_env import VecWindFarmEnv


class _LazyInfos(dict):
    """Info dict whose farm power sums are evaluated on first access (one torch reduction each)."""

    _LAZY = {"Power agent": "Power pr turbine agent", "Power baseline": "Power pr turbine baseline"}

    def __init__(self, raw, baseline):
        super().__init__(raw)
        self._pending = {"Power agent"} | ({"Power baseline"} if baseline else set())

    def __missing__(self, key):
        if key in self._pending:
            self._pending.discard(key)
            val = dict.__getitem__(self, self._LAZY[key]).sum(dim=1)
            self[key] = val
            return val
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._pending

    def keys(self):
        for k in list(self._pending):
            self[k]
        return dict.keys(self)

    def items(self):
        self.keys()
        return dict.items(self)


class _VectorBase:
    metadata = {"render_modes": [], "autoreset_mode": "same_step"}
    render_mode = None

    def __init__(self, turbine=None, n_envs=1, venv=None, as_torch=False, auto_reset=True, pooled=None, **env_kwargs):
        """``venv``: a ready ``VecWindFarmEnv`` / ``PooledVecEnv`` / ``DevicePooledVecEnv``.  Otherwise one is built:
        with ``auto_reset`` (the default) a ``DevicePooledVecEnv`` -- finished episodes are replaced by pre-developed
        spare envs on the device, no step ever waits for a spin-up; ``pooled=False`` (or ``auto_reset=False``) gives the
        plain ``VecWindFarmEnv`` whose auto-reset is the masked in-step reset (exact reference RNG stream, but every
        finished episode stalls the batch for its ~300-step spin-up)."""
        if venv is None:
            use_pool = auto_reset if pooled is None else bool(pooled)
            if use_pool and env_kwargs.get("sample_site") is None:
                from .pool import DevicePooledVecEnv
                venv = DevicePooledVecEnv(turbine, n_envs, **env_kwargs)
            else:
                venv = VecWindFarmEnv(turbine, n_envs, **env_kwargs)
        self.venv = venv
        v = self.venv
        self.num_envs = v.n_envs
        self.as_torch, self.auto_reset = as_torch, auto_reset
        obs_one = tuple(v.obs_shape[1:])          # (obs_var,) single agent, (T, obs_var) multi agent
        n_act = v.n_turb * getattr(v.ec, "act_var", 1)
        self.single_observation_space = Box(low=-1.0, high=1.0, shape=obs_one, dtype=np.float32)
        self.single_action_space = Box(low=-1.0, high=1.0, shape=(n_act,), dtype=np.float32)
        self.observation_space = Box(low=-1.0, high=1.0, shape=(v.n_envs,) + obs_one, dtype=np.float32)
        self.action_space = Box(low=-1.0, high=1.0, shape=(v.n_envs, n_act), dtype=np.float32)
        self._needs_reset = True

    # ------------------------------------------------------------------------------------------ helpers
    def _out(self, t, dtype=None):
        if self.as_torch:
            return t
        a = t.cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def _infos(self):
        """Dict-of-arrays info with the reference's keys (Wind_Farm_Env.py:527-555), batched on axis 0.  With
        ``as_torch`` the entries are the env's live device views and the farm power sums (``"Power agent"``, the array
        ``RecordEpisodeVals`` reads, ``"Power baseline"``) are computed when first read: a rollout loop that does not
        look at the infos launches no extra kernel per step."""
        v = self.venv
        raw = v._info()
        if self.as_torch:
            return _LazyInfos(raw, v.Baseline_comp)
        out = {}
        for k, val in raw.items():
            out[k] = self._out(val) if torch.is_tensor(val) else np.asarray(val)
        pw = raw["Power pr turbine agent"].sum(dim=1)
        out["Power agent"] = self._out(pw)
        if v.Baseline_comp:
            out["Power baseline"] = self._out(raw["Power pr turbine baseline"].sum(dim=1))
        return out

    def _step_core(self, actions):
        v = self.venv
        if self._needs_reset:
            raise RuntimeError("reset() must be called before step()")
        obs, rew, term, trunc, _ = v.step(actions)
        infos = self._infos()
        final_obs, done_mask = None, None
        if getattr(v, "device_autoreset", False):
            # the pool already swapped fresh episodes in on the device: trunc marks them, obs holds their first
            # observation, v.final_obs the finished episodes' last one.  Nothing is read back here.
            if self.auto_reset:
                final_obs, done_mask = v.final_obs, trunc
        elif self.auto_reset:
            done_mask = trunc.cpu().numpy().astype(bool)          # one small D2H per step: who finished?
            if done_mask.any():
                final_obs = obs.clone()
                trunc_keep, rew_keep = trunc.clone(), rew.clone()
                v.reset(mask=done_mask)                            # masked batched spin-up; writes the new obs rows
                obs, trunc, rew = v.obs, trunc_keep, rew_keep
            else:
                done_mask = None
        return obs, rew, term, trunc, infos, final_obs, done_mask

    def close(self):
        self.venv.close()

    @property
    def unwrapped(self):
        return self


class GymVectorEnv(_VectorBase):
    """``gymnasium.vector.VectorEnv``-style surface over the CUDA batch."""

    def reset(self, seed=None, options=None):
        obs, _ = self.venv.reset(seed=seed)
        self._needs_reset = False
        return self._out(obs), self._infos()

    def step(self, actions):
        obs, rew, term, trunc, infos, final_obs, done = self._step_core(actions)
        if final_obs is not None:
            infos["final_observation"] = self._out(final_obs)
            infos["_final_observation"] = ((done.view(torch.bool) if done.dtype == torch.uint8 else done.bool())
                                           if self.as_torch else done.cpu().numpy().astype(bool)) \
                if torch.is_tensor(done) else done
        if self.as_torch:
            return obs, rew, term, (trunc.view(torch.bool) if trunc.dtype == torch.uint8 else trunc.bool()), infos
        return (obs.cpu().numpy(), rew.cpu().numpy().astype(np.float64), term.cpu().numpy(),
                trunc.cpu().numpy().astype(bool), infos)


class SB3VecEnv(_VectorBase):
    """stable-baselines3 ``VecEnv``-style surface (what ``PPO(..., env=...)`` drives)."""

    def reset(self):
        obs, _ = self.venv.reset()
        self._needs_reset = False
        return self._out(obs)

    def step_async(self, actions):
        self._pending = actions

    def step_wait(self):
        obs, rew, term, trunc, infos, final_obs, done = self._step_core(self._pending)
        dones = trunc.cpu().numpy().astype(bool)
        fo = final_obs.cpu().numpy() if (final_obs is not None and dones.any()) else None
        keys = [k for k, val in infos.items() if not np.isscalar(val)]
        host = {k: (infos[k].cpu().numpy() if torch.is_tensor(infos[k]) else np.asarray(infos[k])) for k in keys}
        info_list = []
        for i in range(self.num_envs):
            d = {k: host[k][i] for k in keys if host[k].shape[:1] == (self.num_envs,)}
            d["TimeLimit.truncated"] = bool(dones[i])
            if fo is not None and dones[i]:
                d["terminal_observation"] = fo[i]
            info_list.append(d)
        return self._out(obs), self._out(rew), dones, info_list

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def seed(self, seed=None):
        self.venv.seed = seed
        return [seed] * self.num_envs

    def get_attr(self, name, indices=None):
        val = getattr(self.venv, name)
        n = self.num_envs if indices is None else len(list(np.atleast_1d(indices)))
        return [val] * n

    def set_attr(self, name, value, indices=None):
        setattr(self.venv, name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        res = getattr(self.venv, method_name)(*args, **kwargs)
        n = self.num_envs if indices is None else len(list(np.atleast_1d(indices)))
        return [res] * n

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(list(np.atleast_1d(indices)))
        return [False] * n


class RecordEpisodeVals:
    """``WindGym/wrappers/recordEpisodeVals.py:8-64`` over a ``GymVectorEnv``: ``return_queue``, ``length_queue``,
    ``time_queue`` (gymnasium ``RecordEpisodeStatistics``) and ``mean_power_queue`` = mean of
    ``infos["Power agent"]`` over each finished episode."""

    def __init__(self, env, buffer_length=100):
        self.env = env
        self.num_envs = env.num_envs
        self.return_queue = deque(maxlen=buffer_length)
        self.length_queue = deque(maxlen=buffer_length)
        self.time_queue = deque(maxlen=buffer_length)
        self.mean_power_queue = deque(maxlen=buffer_length)
        self.episode_count = 0
        self._zero()

    def _zero(self):
        n = self.num_envs
        self.episode_returns = np.zeros(n)
        self.episode_lengths = np.zeros(n, dtype=np.int64)
        self.episode_powers = np.zeros(n)
        self.episode_start_times = np.full(n, time.perf_counter())

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        self._zero()
        return obs, info

    def step(self, actions):
        obs, rewards, terminations, truncations, infos = self.env.step(actions)
        to_np = lambda x: x.cpu().numpy() if torch.is_tensor(x) else np.asarray(x)
        r, done = to_np(rewards).astype(np.float64), to_np(terminations) | to_np(truncations).astype(bool)
        self.episode_returns += r
        self.episode_lengths += 1
        self.episode_powers += to_np(infos["Power agent"]).astype(np.float64)
        if done.any():
            idx = np.flatnonzero(done)
            now = time.perf_counter()
            infos["episode"] = {"r": np.where(done, self.episode_returns, 0.0), "l": np.where(done, self.episode_lengths, 0),
                                "t": np.where(done, now - self.episode_start_times, 0.0)}
            infos["_episode"] = done
            for i in idx:
                self.return_queue.append(self.episode_returns[i])
                self.length_queue.append(int(self.episode_lengths[i]))
                self.time_queue.append(now - self.episode_start_times[i])
                self.mean_power_queue.append(self.episode_powers[i] / self.episode_lengths[i])
            self.episode_count += idx.size
            self.episode_returns[idx] = 0.0
            self.episode_lengths[idx] = 0
            self.episode_powers[idx] = 0.0
            self.episode_start_times[idx] = now
        return obs, rewards, terminations, truncations, infos


def collect_rollout(venv, policy, n_steps, auto_reset=True):
    """PPO-style rollout buffer on the device: ``policy(obs) -> actions`` is called on the env's own tensors, nothing
    leaves the GPU.  Returns ``obs [n, B, ...]``, ``actions [n, B, T]``, ``rewards [n, B]``, ``dones [n, B]`` and the
    observation after the last step.  With a ``DevicePooledVecEnv`` (the rollout loop of a training run) finished
    episodes are replaced on the device and the loop never synchronises; with a plain ``VecWindFarmEnv`` and
    ``auto_reset`` they take the masked in-step reset.  Works for the single-agent layout (obs [B, obs_var]) and the multi-agent one
    (``multi_agent=True``: obs [B, T, obs_var], one action per agent -- BASELINE.json cfg 5: obs f32[128, 2048, 8, 2],
    actions f32[128, 2048, 8, 1] after ``unsqueeze(-1)``)."""
    B, T, dev = venv.n_envs, venv.n_turb, venv.device
    obs = venv.obs
    buf_obs = torch.empty((n_steps,) + tuple(obs.shape), dtype=torch.float32, device=dev)
    buf_act = torch.empty((n_steps, B, T), dtype=torch.float32, device=dev)
    buf_rew = torch.empty((n_steps, B), dtype=torch.float32, device=dev)
    buf_done = torch.empty((n_steps, B), dtype=torch.bool, device=dev)
    pooled = getattr(venv, "device_autoreset", False)     # DevicePooledVecEnv: episodes are replaced on the device
    for i in range(n_steps):
        buf_obs[i] = obs
        act = policy(obs).reshape(B, T).to(torch.float32)
        buf_act[i] = act
        obs, rew, term, trunc, _ = venv.step(act)
        buf_rew[i], buf_done[i] = rew, trunc.bool()
        if auto_reset and not pooled:
            done = trunc.cpu().numpy().astype(bool)
            if done.any():
                venv.reset(mask=done)
                obs = venv.obs
    return {"obs": buf_obs, "actions": buf_act, "rewards": buf_rew, "dones": buf_done, "last_obs": obs.clone()}

"""Drop-in test / evaluation agents of the reference (``WindGym/Agents/{BaseAgent,ConstantAgent,RandomAgent,
GreedyAgent}.py``) plus their batched forms, and an SB3-``MlpPolicy``-compatible torch policy that runs on the
device against the env's tensors (the shipped ``examples/PPO_2975000.zip`` holds such a state dict:
obs -> 64 -> 64 tanh -> n_turb actions).

Every agent has the reference's ``predict(obs, deterministic=...) -> (action, state)`` for one env and
``predict_batch(obs[B, obs_var]) -> torch float32 [B, T]`` for a ``VecWindFarmEnv``.
"""
import numpy as np
import torch


class BaseAgent:
    """``Agents/BaseAgent.py:8-23``: holds the yaw range, scales yaw angles to actions in [-1, 1]."""

    def __init__(self, yaw_max=45, yaw_min=-45):
        self.yaw_max = yaw_max
        self.yaw_min = yaw_min

    def predict(self, *args, **kwargs):
        pass

    def scale_yaw(self, yaws):
        return (yaws - self.yaw_min) / (self.yaw_max - self.yaw_min) * 2 - 1


class ConstantAgent(BaseAgent):
    """``Agents/ConstantAgent.py:9-37``: constant target yaw angles (used with ``ActionMethod: "wind"``)."""

    def __init__(self, yaw_angles, yaw_max=45, yaw_min=-45):
        self.UseEnv = True
        self.yaw_max, self.yaw_min = yaw_max, yaw_min
        self.yaw_angles = np.array(yaw_angles) if isinstance(yaw_angles, list) else yaw_angles

    def predict(self, *args, **kwargs):
        return self.scale_yaw(self.yaw_angles), None

    def predict_batch(self, obs, env=None):
        a = torch.as_tensor(np.asarray(self.scale_yaw(np.asarray(self.yaw_angles, dtype=np.float64)), dtype=np.float32))
        return a.to(obs.device).reshape(1, -1).expand(obs.shape[0], -1).contiguous()


class RandomAgent(BaseAgent):
    """``Agents/RandomAgent.py:8-24``: samples the action space."""

    def __init__(self, env=None, seed=None):
        self.UseEnv = True
        self.env = env
        self._gen = torch.Generator().manual_seed(0 if seed is None else int(seed))

    def predict(self, *args, **kwargs):
        return self.env.action_space.sample(), None

    def predict_batch(self, obs, env=None):
        env = env if env is not None else self.env
        a = torch.rand((obs.shape[0], env.n_turb), generator=self._gen, dtype=torch.float32) * 2 - 1
        return a.to(obs.device)


def local_yaw_controller(fs, yaw_step=1):
    """``BasicControllers.py:10-46``: step every turbine's yaw offset towards its local wind direction."""
    wt = fs.windTurbines
    uvw = np.asarray(wt.rotor_avg_windspeed)
    yaw = np.asarray(wt.yaw, dtype=np.float64)
    off = np.rad2deg(np.arctan(uvw[:, 1] / uvw[:, 0])) - yaw
    return yaw + np.sign(off) * np.minimum(np.abs(off), yaw_step)


def global_yaw_controller(fs, yaw_step=1):
    """``BasicControllers.py:49-73``: step every turbine's yaw offset towards zero."""
    yaw = np.asarray(fs.windTurbines.yaw, dtype=np.float64)
    return yaw - np.sign(yaw) * np.minimum(np.abs(yaw), yaw_step)


class GreedyAgent(BaseAgent):
    """``Agents/GreedyAgent.py:13-45``: the baseline controller as an agent."""

    def __init__(self, type="local", yaw_max=45, yaw_min=-45, yaw_step=1, env=None):
        self.UseEnv = True
        self.env = env
        self.yaw_max, self.yaw_min, self.yaw_step = yaw_max, yaw_min, yaw_step
        self.kind = type
        self.controller = local_yaw_controller if type == "local" else global_yaw_controller

    def predict(self, *args, **kwargs):
        return self.scale_yaw(self.controller(fs=self.env.fs, yaw_step=self.yaw_step)), None

    def predict_batch(self, obs, env=None):
        env = env if env is not None else self.env
        s = env.state
        yaw = s["yaw"][:, 0]
        if self.kind == "local":
            off = torch.rad2deg(torch.atan(s["v"][:, 0] / s["u"][:, 0])) - yaw
            goal = yaw + torch.sign(off) * torch.clamp(off.abs(), max=self.yaw_step)
        else:
            goal = yaw - torch.sign(yaw) * torch.clamp(yaw.abs(), max=self.yaw_step)
        return ((goal - self.yaw_min) / (self.yaw_max - self.yaw_min) * 2 - 1).to(torch.float32)


class SB3MlpPolicy(torch.nn.Module):
    """Actor of a stable-baselines3 ``MlpPolicy`` (``policy.pth`` inside the SB3 zip): ``mlp_extractor.policy_net``
    (Linear-tanh stack) -> ``action_net``; deterministic action = clip(mean, -1, 1), stochastic adds
    ``exp(log_std)`` noise.  Runs wherever its weights live -- on the GPU next to the env tensors there is no
    host round trip in the rollout."""

    def __init__(self, obs_dim, n_actions, hidden=(64, 64)):
        super().__init__()
        layers, d = [], obs_dim
        for h in hidden:
            layers += [torch.nn.Linear(d, h), torch.nn.Tanh()]
            d = h
        self.policy_net = torch.nn.Sequential(*layers)
        self.action_net = torch.nn.Linear(d, n_actions)
        self.log_std = torch.nn.Parameter(torch.zeros(n_actions))

    @classmethod
    def from_state_dict(cls, sd):
        """Build from SB3's key names (``mlp_extractor.policy_net.{0,2,..}.weight``, ``action_net.weight``, ``log_std``)."""
        ws = sorted((int(k.split(".")[2]), v) for k, v in sd.items()
                    if k.startswith("mlp_extractor.policy_net.") and k.endswith(".weight"))
        if not ws:
            raise ValueError("state dict holds no mlp_extractor.policy_net.*.weight: not an SB3 MlpPolicy")
        hidden = tuple(int(w.shape[0]) for _, w in ws)
        obs_dim, n_act = int(ws[0][1].shape[1]), int(sd["action_net.weight"].shape[0])
        pol = cls(obs_dim, n_act, hidden)
        own = {}
        for k, v in sd.items():
            if k.startswith("mlp_extractor.policy_net."):
                own["policy_net." + k[len("mlp_extractor.policy_net."):]] = v
            elif k.startswith("action_net.") or k == "log_std":
                own[k] = v
        pol.load_state_dict(own)
        return pol

    @classmethod
    def load(cls, path, device="cpu"):
        """``policy.pth`` extracted from an SB3 zip (plain ``torch.load(weights_only=True)``)."""
        return cls.from_state_dict(torch.load(path, map_location=device, weights_only=True)).to(device)

    @classmethod
    def from_zip(cls, path, device="cpu"):
        """The model file stable-baselines3 writes (``PPO.save`` -> ``*.zip`` holding ``policy.pth``; the reference ships
        ``examples/PPO_2975000.zip`` and loads it with ``PPO.load``, examples/Example 3 .. / AgentEval usage): read
        ``policy.pth`` straight out of the archive, no stable-baselines3 needed."""
        import io
        import zipfile
        with zipfile.ZipFile(path) as z:
            if "policy.pth" not in z.namelist():
                raise ValueError(f"{path}: no policy.pth inside (not a stable-baselines3 model zip)")
            buf = io.BytesIO(z.read("policy.pth"))
        return cls.from_state_dict(torch.load(buf, map_location=device, weights_only=True)).to(device)

    @torch.no_grad()
    def predict_batch(self, obs, env=None, deterministic=True):
        mean = self.action_net(self.policy_net(obs.to(self.log_std.device, torch.float32)))
        if not deterministic:
            mean = mean + torch.randn_like(mean) * self.log_std.exp()
        return mean.clamp(-1.0, 1.0)

    def predict(self, obs, deterministic=True, **kw):
        a = self.predict_batch(torch.as_tensor(np.asarray(obs, dtype=np.float32)).reshape(1, -1), deterministic=deterministic)
        return a[0].cpu().numpy(), None


def batch_actions(model, obs, env, deterministic=True):
    """Actions [B, T] (torch, on the env's device) from any supported agent: batched agents are called once, a plain
    ``predict(obs)`` agent (SB3-style) is called per env on host copies."""
    if hasattr(model, "predict_batch"):
        try:
            a = model.predict_batch(obs, env=env, deterministic=deterministic)
        except TypeError:
            a = model.predict_batch(obs, env=env)
        return a.to(env.device, torch.float32).reshape(env.n_envs, -1)
    o = obs.cpu().numpy()
    acts = np.stack([np.asarray(model.predict(o[i], deterministic=deterministic)[0], dtype=np.float32).reshape(-1)
                     for i in range(o.shape[0])])
    return torch.as_tensor(acts).to(env.device)

"""``check_env`` -- the checks gymnasium's ``utils.env_checker.check_env`` makes on a single env, restated.

The reference runs ``check_env(env.unwrapped)`` on its ``WindFarmEnv`` (``tests/test_basics.py:410-412``); gymnasium is
not a dependency here, so the contract it enforces is written out: Box spaces with finite
bounds and a float32 dtype; ``reset(seed=s)`` returns ``(obs, info)`` with ``obs`` inside the observation space,
exact dtype and shape, and is deterministic for a given seed; ``reset()`` without a seed runs; ``step(action)``
returns ``(obs, float-like, bool, bool, dict)`` with ``obs`` inside the space; a ``render_mode`` attribute exists.
Raises ``AssertionError`` with gymnasium-style messages.
"""
import numbers

import numpy as np


def _check_box(space, name):
    for attr in ("low", "high", "shape", "dtype"):
        assert hasattr(space, attr), f"The {name} space must be a Box-like space with `{attr}`"
    low, high = np.broadcast_to(space.low, space.shape), np.broadcast_to(space.high, space.shape)
    assert np.all(np.isfinite(low)) and np.all(np.isfinite(high)), f"The {name} space must have finite bounds"
    assert np.all(low <= high), f"The {name} space has low > high"
    assert len(space.shape) >= 1, f"The {name} space must not be a scalar space"


def _check_obs(obs, space, where):
    assert isinstance(obs, np.ndarray), f"The observation returned by `{where}` must be a numpy array, got {type(obs)}"
    assert obs.dtype == space.dtype, f"The observation returned by `{where}` has dtype {obs.dtype}, expected {space.dtype}"
    assert obs.shape == tuple(space.shape), f"The observation returned by `{where}` has shape {obs.shape}, expected {space.shape}"
    assert space.contains(obs), f"The observation returned by `{where}` is not within the observation space"
    assert np.all(np.isfinite(obs)), f"The observation returned by `{where}` holds NaN or inf"


def check_env(env, seed=123, n_steps=3):
    """Run the checks on a single-agent env (``WindFarmEnv`` / ``FarmEval``)."""
    assert hasattr(env, "observation_space") and hasattr(env, "action_space"), "The env must define its spaces"
    assert hasattr(env, "render_mode"), "The env must have a `render_mode` attribute"
    assert hasattr(env, "metadata") and "render_modes" in env.metadata, "env.metadata must list `render_modes`"
    _check_box(env.observation_space, "observation")
    _check_box(env.action_space, "action")
    assert env.action_space.dtype == np.float32 and env.observation_space.dtype == np.float32

    out = env.reset(seed=seed)
    assert isinstance(out, tuple) and len(out) == 2, "`reset()` must return a tuple (obs, info)"
    obs_1, info = out
    assert isinstance(info, dict), "The second value returned by `reset()` must be a dict"
    _check_obs(obs_1, env.observation_space, "reset()")
    obs_2, _ = env.reset(seed=seed)
    assert np.array_equal(obs_1, obs_2), "Using `env.reset(seed=s)` twice gave different observations: reset is not deterministic"
    obs_3, _ = env.reset()                      # unseeded reset after a seeded one must run
    _check_obs(obs_3, env.observation_space, "reset()")
    obs_4, _ = env.reset(seed=seed + 1)
    _check_obs(obs_4, env.observation_space, "reset()")

    env.reset(seed=seed)
    rng = np.random.default_rng(seed)
    for _ in range(n_steps):
        action = rng.uniform(env.action_space.low, env.action_space.high).astype(env.action_space.dtype)
        assert env.action_space.contains(action)
        res = env.step(action)
        assert isinstance(res, tuple) and len(res) == 5, "`step()` must return (obs, reward, terminated, truncated, info)"
        obs, reward, terminated, truncated, info = res
        _check_obs(obs, env.observation_space, "step()")
        assert isinstance(reward, numbers.Real) and not isinstance(reward, bool), f"The reward must be a float, got {type(reward)}"
        assert np.isfinite(reward), "The reward is NaN or inf"
        assert isinstance(terminated, (bool, np.bool_)) and isinstance(truncated, (bool, np.bool_)), \
            "`terminated` and `truncated` must be booleans"
        assert isinstance(info, dict), "`info` must be a dict"
        if terminated or truncated:
            env.reset(seed=seed)
    return True

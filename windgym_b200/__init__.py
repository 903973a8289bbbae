"""windgym_b200 -- B200-native batched wind-farm RL environment (WindGym-compatible hot path).

Host side mirrors the reference's interface for the per-step hot path (``WindFarmEnv.reset/step``,
``FarmEval``, ``WindFarmEnvMulti``; reference ``WindGym/Wind_Farm_Env.py``); the compute is hand-written
sm_100a CUDA in ``csrc/`` behind the C-ABI of ``include/windgym_b200.h``.  No CPU fallback.
"""
from ._lib import WgError, LIB_PATH  # noqa: F401
from .config import EnvConfig, grid_layout, load_yaml  # noqa: F401
from .turbines import V80  # noqa: F401


def __getattr__(name):  # torch-dependent classes are imported lazily so `import windgym_b200` stays cheap
    if name == "VecWindFarmEnv":
        from .vec_env import VecWindFarmEnv
        return VecWindFarmEnv
    if name in ("WindFarmEnv", "FarmEval", "WindFarmEnvMulti"):
        from . import envs
        return getattr(envs, name)
    if name in ("PooledVecEnv", "DevicePooledVecEnv"):
        from . import pool
        return getattr(pool, name)
    if name in ("GymVectorEnv", "SB3VecEnv", "RecordEpisodeVals"):
        from . import vector
        return getattr(vector, name)
    if name in ("AgentEval", "eval_batched", "EvalDataset"):
        from . import evaluate
        return getattr(evaluate, name)
    if name in ("BaseAgent", "ConstantAgent", "RandomAgent", "GreedyAgent", "SB3MlpPolicy"):
        from . import agents
        return getattr(agents, name)
    raise AttributeError(name)

"""Import-only stub of the gymnasium names the reference touches (WindEnv.py:2, Wind_Farm_Env.py:3,:458)."""
import numpy as np
from . import spaces  # noqa: F401


class Env:
    metadata = {}
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random = np.random.default_rng()
        return self._np_random

    @np_random.setter
    def np_random(self, v):
        self._np_random = v

    def reset(self, seed=None, options=None):
        # gymnasium.utils.seeding.np_random(seed) == Generator(PCG64(SeedSequence(seed))) == default_rng(seed)
        if seed is not None:
            self._np_random = np.random.default_rng(seed)

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass

"""Vectorised replica of ``np.random.default_rng([seed, env, episode]).random(n)`` for many envs at once.

The per-env condition draws of a reset follow the reference's order (ws -> ti -> wd -> yaw,
``Wind_Farm_Env.py:564-568, :715``) on the stream ``default_rng([seed, env, episode])``.  Building thousands of
``Generator`` objects costs ~17 us each on the host (SeedSequence hashing, PCG64 seeding) -- more than the GPU needs
to step the whole batch.  This module evaluates the same bit streams with numpy array arithmetic: numpy's
``SeedSequence`` (entropy mixing of 32-bit words), the PCG64 seeding and its XSL-RR 128/64 output function, and the
53-bit double conversion.  ``tests/test_host_logic.py`` pins it bit for bit against numpy.
"""
import numpy as np

_M32 = np.uint64(0xFFFFFFFF)
_INIT_A, _MULT_A = 0x43B0D7E5, 0x931E8875
_INIT_B, _MULT_B = 0x8B51F9DD, 0x58F38DED
_MIX_L, _MIX_R = 0xCA01F9DD, 0x4973F715
_PCG_MULT_HI, _PCG_MULT_LO = np.uint64(0x2360ED051FC65DA4), np.uint64(0x4385DF649FCCF645)


def _u32(x):
    return x & _M32


class _Hash:
    """hashmix of SeedSequence: the multiplier evolves with every call (shared by all envs: it is data independent)."""

    def __init__(self, init, mult):
        self.c, self.mult = init, mult

    def __call__(self, value):
        value = _u32(value ^ np.uint64(self.c))
        self.c = (self.c * self.mult) & 0xFFFFFFFF
        value = _u32(value * np.uint64(self.c))
        return _u32(value ^ (value >> np.uint64(16)))


def _mix(x, y):
    r = _u32(np.uint64(_MIX_L) * x - np.uint64(_MIX_R) * y)
    return _u32(r ^ (r >> np.uint64(16)))


def _seed_words(entropy_words):
    """SeedSequence(entropy).generate_state(4, uint64) for entropy given as a list of uint32 word arrays [n]."""
    n = entropy_words[0].shape[0]
    h = _Hash(_INIT_A, _MULT_A)
    pool = [h(entropy_words[i]) if i < len(entropy_words) else h(np.zeros(n, dtype=np.uint64)) for i in range(4)]
    for i_src in range(4):
        for i_dst in range(4):
            if i_src != i_dst:
                pool[i_dst] = _mix(pool[i_dst], h(pool[i_src]))
    for i_src in range(4, len(entropy_words)):
        for i_dst in range(4):
            pool[i_dst] = _mix(pool[i_dst], h(entropy_words[i_src]))
    g = _Hash(_INIT_B, _MULT_B)
    w = [g(pool[i % 4]) for i in range(8)]
    return [w[2 * k] | (w[2 * k + 1] << np.uint64(32)) for k in range(4)]   # little-endian pairs -> uint64


def _mul64(a, b):
    """Full 64 x 64 -> 128 bit product of uint64 arrays: (high, low)."""
    a0, a1 = a & _M32, a >> np.uint64(32)
    b0, b1 = b & _M32, b >> np.uint64(32)
    p00, p01, p10, p11 = a0 * b0, a0 * b1, a1 * b0, a1 * b1
    mid = (p00 >> np.uint64(32)) + (p01 & _M32) + (p10 & _M32)
    lo = (p00 & _M32) | (mid << np.uint64(32))
    hi = p11 + (p01 >> np.uint64(32)) + (p10 >> np.uint64(32)) + (mid >> np.uint64(32))
    return hi, lo


def _step(sh, sl, ih, il):
    """state <- state * MULT + inc (mod 2^128), limbs as uint64 arrays."""
    hi, lo = _mul64(sl, _PCG_MULT_LO)
    hi = hi + sl * _PCG_MULT_HI + sh * _PCG_MULT_LO
    lo2 = lo + il
    carry = (lo2 < lo).astype(np.uint64)
    return hi + ih + carry, lo2


def uniform_streams(seed, env_ids, episode, n_draws):
    """[len(env_ids), n_draws] float64 in [0, 1): row k equals ``default_rng([seed, env_ids[k], episode]).random(n_draws)``.
    All three integers must lie in [0, 2**32)."""
    env = np.asarray(env_ids, dtype=np.uint64).reshape(-1)
    n = env.shape[0]
    if not (0 <= int(seed) < 2 ** 32 and 0 <= int(episode) < 2 ** 32 and (n == 0 or int(env.max()) < 2 ** 32)):
        raise ValueError("uniform_streams: seed, env ids and episode must fit 32 bits")
    with np.errstate(over="ignore"):
        words = [np.full(n, int(seed), dtype=np.uint64), env, np.full(n, int(episode), dtype=np.uint64)]
        s0, s1, s2, s3 = _seed_words(words)          # initstate = (s0, s1), initseq = (s2, s3): (high, low)
        ih = (s2 << np.uint64(1)) | (s3 >> np.uint64(63))
        il = (s3 << np.uint64(1)) | np.uint64(1)
        sh, sl = _step(np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64), ih, il)
        sl2 = sl + s1
        sh = sh + s0 + (sl2 < sl).astype(np.uint64)
        sh, sl = _step(sh, sl2, ih, il)
        out = np.empty((n, n_draws), dtype=np.float64)
        for k in range(n_draws):
            sh, sl = _step(sh, sl, ih, il)
            x = sh ^ sl
            rot = sh >> np.uint64(58)
            r = (x >> rot) | (x << ((np.uint64(64) - rot) & np.uint64(63)))
            out[:, k] = (r >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return out

"""Minimal wind-resource site with the py_wake ``Site.local_wind`` protocol the env reads when ``sample_site`` is
given (reference ``WindGym/Wind_Farm_Env.py:569-577``: ``Sector_frequency_ilk``, ``Weibull_A_ilk``,
``Weibull_k_ilk`` on 1-degree sectors).  py_wake is not a dependency; a real py_wake site works unchanged."""
import types

import numpy as np


class WeibullSite:
    """Sector-wise Weibull site: ``freq``, ``A``, ``k`` given on ``n`` equal sectors (interpolated to the requested
    directions, nearest sector), e.g. the 12-sector tables of py_wake's Hornsrev1 site."""

    def __init__(self, freq, A, k):
        self.freq, self.A, self.k = (np.asarray(v, dtype=np.float64) for v in (freq, A, k))
        if not (self.freq.size == self.A.size == self.k.size):
            raise ValueError("freq, A and k must have one entry per sector")

    def local_wind(self, x=0, y=0, wd=np.arange(0, 360, 1), ws=np.arange(3, 25, 1), **_):
        wd = np.asarray(wd, dtype=np.float64).reshape(-1)
        n = self.freq.size
        sec = np.floor(((wd + 180.0 / n) % 360.0) / (360.0 / n)).astype(int) % n
        per_deg = self.freq[sec] / np.maximum(np.bincount(sec, minlength=n)[sec], 1)
        shape = (1, wd.size, 1)
        return types.SimpleNamespace(Sector_frequency_ilk=(per_deg / per_deg.sum()).reshape(shape),
                                     Weibull_A_ilk=self.A[sec].reshape(shape), Weibull_k_ilk=self.k[sec].reshape(shape))

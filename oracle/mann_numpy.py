"""TEST INFRASTRUCTURE -- fp64 numpy restatement of the Mann turbulence box behind ``dynamiks``'
``MannTurbulenceField`` (reference call sites ``WindGym/Wind_Farm_Env.py:616-659``: ``generate(alphaepsilon, L,
Gamma, Nxyz, dxyz, seed)``, ``from_netcdf``, ``scale_TI(TI, U)``; ``tests/test_basics.py:32-46`` fixture).

PARITY UNPINNED: ``hipersim>=0.1.7`` / ``dynamiks`` are not vendored and not installed.
This file restates the published algorithm:

* J. Mann (1998), "Wind field simulation", Probabilistic Engineering Mechanics 13(4):269-282: the sheared spectral
  tensor (eddy lifetime beta = Gamma (kL)^(-2/3) / sqrt(2F1(1/3, 17/6; 4/3; -(kL)^-2)), k30 = k3 + beta k1,
  von Karman E(k0) = alphaepsilon L^(5/3) (k0 L)^4 / (1 + (k0 L)^2)^(17/6)) and its factorisation C(k) (eq. 46),
  Fourier simulation u(x) = sum_k exp(i k.x) C(k) n(k) sqrt(dk1 dk2 dk3) with complex unit Gaussians n(k).
* Frozen-turbulence sampling (Taylor): the field at (x, y, z, t) is the box at (x - U t, y, z), trilinear, periodic.
* DWM meandering (Larsen et al. 2008): wake centres move with the LARGE scales only -- a box filter of 2 D x 2 D
  in (y, z) applied to (v, w) (``lowpass_yz``).

The CUDA path's generator (``windgym_b200/mann.py``, torch + cuFFT) is checked against ``mann_box`` on the same
noise, the sampling kernels against ``MannTurbulenceField.sample*``.
"""
import numpy as np


def _hyp2f1_beta(kL):
    """2F1(1/3, 17/6; 4/3; -(kL)^-2) (scipy when present, otherwise Pfaff transformation + series)."""
    z = -1.0 / np.maximum(kL, 1e-30) ** 2
    try:
        from scipy.special import hyp2f1
        return hyp2f1(1.0 / 3.0, 17.0 / 6.0, 4.0 / 3.0, z)
    except Exception:  # pragma: no cover
        a, b, c = 1.0 / 3.0, 17.0 / 6.0, 4.0 / 3.0
        w = z / (z - 1.0)                       # in (0, 1): 2F1(a,b;c;z) = (1-z)^-a 2F1(a, c-b; c; w)
        term, tot = np.ones_like(w), np.ones_like(w)
        for n in range(4000):
            term = term * (a + n) * (c - b + n) / ((c + n) * (n + 1.0)) * w
            tot = tot + term
        return (1.0 - z) ** (-a) * tot


def wave_numbers(Nxyz, dxyz):
    """FFT-ordered angular wave numbers of the periodic box (rad/m)."""
    return [2.0 * np.pi * np.fft.fftfreq(n, d) for n, d in zip(Nxyz, dxyz)]


def tensor_factor(k1, k2, k3, alphaepsilon, L, Gamma):
    """C(k) of Mann (1998) eq. 46 for broadcastable wave-number arrays -> [3, 3, ...] (zero at k = 0)."""
    k1, k2, k3 = np.broadcast_arrays(k1, k2, k3)
    kk = k1 * k1 + k2 * k2 + k3 * k3
    zero = kk == 0.0
    kk = np.where(zero, 1.0, kk)
    kL = np.sqrt(kk) * L
    beta = Gamma / (kL ** (2.0 / 3.0) * np.sqrt(_hyp2f1_beta(kL)))
    k30 = k3 + beta * k1
    k0k0 = k1 * k1 + k2 * k2 + k30 * k30
    E = alphaepsilon * L ** (5.0 / 3.0) * (k0k0 * L * L) ** 2 / (1.0 + k0k0 * L * L) ** (17.0 / 6.0)
    kh2 = k1 * k1 + k2 * k2                                    # horizontal wave number squared
    kh2s = np.where(kh2 == 0.0, 1.0, kh2)
    C1 = beta * k1 * k1 * (k0k0 - 2.0 * k30 * k30 + beta * k1 * k30) / (kk * kh2s)
    C2 = k2 * k0k0 / kh2s ** 1.5 * np.arctan2(beta * k1 * np.sqrt(kh2s), k0k0 - k30 * k1 * beta)
    k1s = np.where(k1 == 0.0, 1.0, k1)
    zeta1 = np.where(k1 == 0.0, -beta, C1 - k2 / k1s * C2)     # k1 -> 0 limits (Mann 1998, below eq. 16)
    zeta2 = np.where(k1 == 0.0, 0.0, k2 / k1s * C1 + C2)
    with np.errstate(divide="ignore", invalid="ignore"):
        amp = np.sqrt(E / (4.0 * np.pi)) / k0k0
    C = np.zeros((3, 3) + k1.shape)
    C[0, 0], C[0, 1], C[0, 2] = k2 * zeta1, k30 - k1 * zeta1, -k2
    C[1, 0], C[1, 1], C[1, 2] = k2 * zeta2 - k30, -k1 * zeta2, k1
    C[2, 0], C[2, 1], C[2, 2] = k0k0 * k2 / kk, -k0k0 * k1 / kk, 0.0
    C = C * amp
    C[..., zero] = 0.0
    C[:, :, kh2 == 0.0] = 0.0                                  # purely vertical wave vectors carry no energy here
    return C


def box_noise(Nxyz, seed):
    """Complex unit Gaussians n(k) [3, Nx, Ny, Nz] (fp64 pairs) from ``default_rng(seed)`` -- shared by both paths."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((2, 3) + tuple(Nxyz))


def mann_box(alphaepsilon, L, Gamma, Nxyz, dxyz, seed=1, noise=None):
    """(u, v, w) fluctuations [3, Nx, Ny, Nz] of one periodic Mann box."""
    if noise is None:
        noise = box_noise(Nxyz, seed)
    n = noise[0] + 1j * noise[1]
    k1, k2, k3 = wave_numbers(Nxyz, dxyz)
    C = tensor_factor(k1[:, None, None], k2[None, :, None], k3[None, None, :], alphaepsilon, L, Gamma)
    dk = np.prod([2.0 * np.pi / (nn * d) for nn, d in zip(Nxyz, dxyz)])
    dZ = np.einsum("ij...,j...->i...", C, n) * np.sqrt(dk)
    # u(x) = sum_k exp(i k.x) dZ(k): unnormalised inverse transform; the real part is the field
    return np.real(np.fft.ifftn(dZ, axes=(1, 2, 3))) * np.prod(Nxyz)


def lowpass_yz(uvw, dxyz, width):
    """Periodic box filter of ``width`` metres in y and z (odd number of cells, at least 1) -- DWM meandering scales."""
    out = uvw
    for ax, d in ((2, dxyz[1]), (3, dxyz[2])):
        m = max(1, int(round(width / d)))
        m += 1 - (m % 2)
        m = min(m, out.shape[ax] - 1 + (out.shape[ax] % 2))
        acc = np.zeros_like(out)
        for s in range(-(m // 2), m // 2 + 1):
            acc += np.roll(out, s, axis=ax)
        out = acc / m
    return out


class MannTurbulenceField:
    """``dynamiks.sites.turbulence_fields.MannTurbulenceField`` stand-in: a periodic box + frozen-turbulence sampling."""

    def __init__(self, uvw, dxyz, lowpass_width=160.0, offset=(0.0, 0.0, 0.0)):
        self.uvw = np.asarray(uvw, dtype=np.float64)
        self.dxyz = tuple(float(d) for d in dxyz)
        self.scale = 1.0
        self.offset = np.asarray(offset, dtype=np.float64)
        self.ti = 0.0
        self.lowpass_width = float(lowpass_width)
        self.uvw_lp = lowpass_yz(self.uvw, self.dxyz, lowpass_width)

    @classmethod
    def generate(cls, alphaepsilon=0.1, L=33.6, Gamma=3.9, Nxyz=(64, 32, 16), dxyz=(3.0, 3.0, 3.0), seed=1, **kw):
        return cls(mann_box(alphaepsilon, L, Gamma, Nxyz, dxyz, seed), dxyz, **kw)

    def scale_TI(self, TI, U):
        """Rescale so that std(u) over the box equals TI * U (dynamiks ``scale_TI``)."""
        self.scale = float(TI * U / np.std(self.uvw[0]))
        self.ti = float(TI)

    def _trilinear(self, box, x, y, z):
        n = box.shape[1:]
        pos = [(np.asarray(c, dtype=np.float64) + o) / d for c, o, d in zip((x, y, z), self.offset, self.dxyz)]
        i0 = [np.floor(p).astype(np.int64) for p in pos]
        fr = [p - i for p, i in zip(pos, i0)]
        out = np.zeros((box.shape[0],) + np.shape(pos[0]))
        for a in (0, 1):
            for b in (0, 1):
                for c in (0, 1):
                    w = (fr[0] if a else 1 - fr[0]) * (fr[1] if b else 1 - fr[1]) * (fr[2] if c else 1 - fr[2])
                    out += w * box[:, (i0[0] + a) % n[0], (i0[1] + b) % n[1], (i0[2] + c) % n[2]]
        return out * self.scale

    def sample(self, x, y, z, t, U):
        """(u', v', w') at wind-aligned points and time t (Taylor: the box is advected with U)."""
        return self._trilinear(self.uvw, np.asarray(x) - U * t, y, z)

    def sample_lp(self, x, y, z, t, U):
        """Low-pass filtered (v', w') that move the wake centres."""
        return self._trilinear(self.uvw_lp, np.asarray(x) - U * t, y, z)[1:]

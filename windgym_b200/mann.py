"""Mann turbulence boxes on the device (reference ``WindGym/Wind_Farm_Env.py:598-678``
``_def_site``: ``MannTurbulenceField.generate / from_netcdf / scale_TI`` of dynamiks over hipersim).

Generation is reset-time work: the sheared spectral tensor factor C(k) of Mann (1998, Prob. Eng. Mech. 13:269,
eq. 46) is evaluated on the device slab by slab, multiplied with complex Gaussian noise and inverse-transformed
with cuFFT (``torch.fft.ifftn``).  The per-step consumers are the CUDA kernels: ``wg_flow_kernel`` samples the box
trilinearly with the Taylor shift (meandering of the wake centres from the low-pass filtered (v, w); rotor-plane
fluctuations from the raw box).  Layout handed to the library: ``raw`` [Nx, Ny, Nz, 4] float32 (u, v, w, 0) -- one
16-byte load per corner -- and ``lp`` [Nx, Ny, Nz, 2] float32 (v, w low-passed over 2 D x 2 D in y, z).
Restated algorithm and tolerances: ``oracle/mann_numpy.py`` (parity unpinned, hipersim is not vendored).
"""
import numpy as np
import torch


def _beta_table(n=8192, lo=-4.0, hi=5.0):
    """sqrt(2F1(1/3, 17/6; 4/3; -(kL)^-2)) on a log10(kL) grid (host, once); interpolated on the device."""
    lg = np.linspace(lo, hi, n)
    kL = 10.0 ** lg
    z = -1.0 / kL ** 2
    try:
        from scipy.special import hyp2f1
        h = hyp2f1(1.0 / 3.0, 17.0 / 6.0, 4.0 / 3.0, z)
    except Exception:  # Pfaff transformation + series: 2F1(a,b;c;z) = (1-z)^-a 2F1(a, c-b; c; z/(z-1))
        a, b, c = 1.0 / 3.0, 17.0 / 6.0, 4.0 / 3.0
        w = z / (z - 1.0)
        term, tot = np.ones_like(w), np.ones_like(w)
        for k in range(6000):
            term = term * (a + k) * (c - b + k) / ((c + k) * (k + 1.0)) * w
            tot = tot + term
        h = (1.0 - z) ** (-a) * tot
    return lg, np.sqrt(h)


def _interp_table(lgk, lg, tab):
    t = ((lgk - lg[0]) / (lg[1] - lg[0])).clamp(0, lg.numel() - 1.000001)
    i = t.floor().long()
    f = t - i
    return tab[i] * (1 - f) + tab[i + 1] * f


def tensor_factor(k1, k2, k3, alphaepsilon, L, Gamma, table):
    """C(k) [3, 3, ...] (fp64 torch) for broadcastable wave numbers; mirrors ``oracle.mann_numpy.tensor_factor``."""
    k1, k2, k3 = torch.broadcast_tensors(k1, k2, k3)
    kk = k1 * k1 + k2 * k2 + k3 * k3
    zero = kk == 0
    kk = torch.where(zero, torch.ones_like(kk), kk)
    kL = kk.sqrt() * L
    beta = Gamma / (kL ** (2.0 / 3.0) * _interp_table(torch.log10(kL), *table))
    k30 = k3 + beta * k1
    k0k0 = k1 * k1 + k2 * k2 + k30 * k30
    E = alphaepsilon * L ** (5.0 / 3.0) * (k0k0 * L * L) ** 2 / (1.0 + k0k0 * L * L) ** (17.0 / 6.0)
    kh2 = k1 * k1 + k2 * k2
    kh2s = torch.where(kh2 == 0, torch.ones_like(kh2), kh2)
    C1 = beta * k1 * k1 * (k0k0 - 2.0 * k30 * k30 + beta * k1 * k30) / (kk * kh2s)
    C2 = k2 * k0k0 / kh2s ** 1.5 * torch.atan2(beta * k1 * kh2s.sqrt(), k0k0 - k30 * k1 * beta)
    k1s = torch.where(k1 == 0, torch.ones_like(k1), k1)
    zeta1 = torch.where(k1 == 0, -beta, C1 - k2 / k1s * C2)
    zeta2 = torch.where(k1 == 0, torch.zeros_like(k1), k2 / k1s * C1 + C2)
    amp = (E / (4.0 * np.pi)).sqrt() / torch.where(k0k0 == 0, torch.ones_like(k0k0), k0k0)
    amp = torch.where(zero | (kh2 == 0), torch.zeros_like(amp), amp)
    z0 = torch.zeros_like(k1)
    rows = [[k2 * zeta1, k30 - k1 * zeta1, -k2], [k2 * zeta2 - k30, -k1 * zeta2, k1],
            [k0k0 * k2 / kk, -k0k0 * k1 / kk, z0]]
    return torch.stack([torch.stack([c * amp for c in r]) for r in rows])


class MannBox:
    """One periodic turbulence box resident on a GPU, in the layouts the flow kernel samples."""

    def __init__(self, uvw, dxyz, lowpass_width=160.0):
        """``uvw``: tensor [3, Nx, Ny, Nz] (any float dtype) on the target device."""
        self.device = uvw.device
        self.Nxyz = tuple(int(n) for n in uvw.shape[1:])
        self.dxyz = tuple(float(d) for d in dxyz)
        self.lowpass_width = float(lowpass_width)
        self.std_u = float(uvw[0].double().std(unbiased=False))
        raw = torch.zeros(self.Nxyz + (4,), dtype=torch.float32, device=self.device)
        raw[..., :3] = uvw.permute(1, 2, 3, 0).to(torch.float32)
        self.raw = raw.contiguous()
        vw = uvw[1:].to(torch.float32)
        for ax, d in ((2, self.dxyz[1]), (3, self.dxyz[2])):   # periodic box filter in y, z (oracle: lowpass_yz)
            n = vw.shape[ax]
            m = max(1, int(round(lowpass_width / d)))
            m += 1 - (m % 2)
            m = min(m, n - 1 + (n % 2))
            acc = torch.zeros_like(vw)
            for s in range(-(m // 2), m // 2 + 1):
                acc += torch.roll(vw, s, dims=ax)
            vw = acc / m
        self.lp = vw.permute(1, 2, 3, 0).contiguous()

    def scale_for(self, TI, U):
        """``scale_TI(TI, U)``: factor that makes std(u) = TI * U (per env: TI and U may be arrays)."""
        return np.asarray(TI, dtype=np.float64) * np.asarray(U, dtype=np.float64) / self.std_u

    @property
    def nbytes(self):
        return self.raw.numel() * 4 + self.lp.numel() * 4

    @classmethod
    def generate(cls, alphaepsilon=0.1, L=33.6, Gamma=3.9, Nxyz=(4096, 512, 64), dxyz=(4.0, 8.0, 8.0), seed=1,
                 device="cuda:0", noise=None, lowpass_width=160.0, slab=64):
        """``MannTurbulenceField.generate`` (Wind_Farm_Env.py:625-637 / :647-656).  ``noise`` [2, 3, Nx, Ny, Nz] injects the
        Gaussian draws (parity tests); otherwise they come from a seeded device generator."""
        dev = torch.device(device)
        Nx, Ny, Nz = (int(n) for n in Nxyz)
        lg, tab = _beta_table()
        table = (torch.as_tensor(lg, device=dev), torch.as_tensor(tab, device=dev))
        ks = [torch.as_tensor(2.0 * np.pi * np.fft.fftfreq(n, d), device=dev) for n, d in zip(Nxyz, dxyz)]
        dk = float(np.prod([2.0 * np.pi / (n * d) for n, d in zip(Nxyz, dxyz)]))
        big = noise is None and Nx * Ny * Nz > (1 << 24)
        cdt = torch.complex64 if big else torch.complex128
        dZ = torch.empty((3, Nx, Ny, Nz), dtype=cdt, device=dev)
        gen = torch.Generator(device=dev).manual_seed(int(seed))
        for x0 in range(0, Nx, slab):
            x1 = min(Nx, x0 + slab)
            C = tensor_factor(ks[0][x0:x1, None, None], ks[1][None, :, None], ks[2][None, None, :], alphaepsilon, L,
                              Gamma, table)
            if noise is not None:
                nz = torch.as_tensor(noise[:, :, x0:x1], device=dev, dtype=torch.float64)
            else:
                nz = torch.randn((2, 3, x1 - x0, Ny, Nz), generator=gen, device=dev, dtype=torch.float64)
            n = torch.complex(nz[0], nz[1])
            dZ[:, x0:x1] = (torch.einsum("ijxyz,jxyz->ixyz", C.to(torch.complex128), n) * np.sqrt(dk)).to(cdt)
            del C, nz, n
        uvw = torch.fft.ifftn(dZ, dim=(1, 2, 3)).real * float(Nx * Ny * Nz)
        del dZ
        return cls(uvw, dxyz, lowpass_width=lowpass_width)

    @classmethod
    def white_noise(cls, Nxyz=(1024, 128, 32), dxyz=(8.0, 8.0, 8.0), seed=1, device="cuda:0", lowpass_width=160.0):
        """``RandomTurbulence(ti, ws, seed)`` of turbtype "Random" (Wind_Farm_Env.py:640-644): uncorrelated Gaussian
        fluctuations, here as a periodic box of independent N(0, 1) cells per component (``scale_for`` makes
        std(u) = TI * U).  White noise has no large scales: its low-pass (meandering) part is close to zero by
        construction.  Between cell centres the trilinear sampling of the flow kernel averages neighbouring cells
        (variance at a point between 1/8 and 1 of the cell variance) -- a rotor average over 16 points spread over 80 m
        of 8 m cells is insensitive to that."""
        dev = torch.device(device)
        gen = torch.Generator(device=dev).manual_seed(int(seed))
        uvw = torch.randn((3,) + tuple(int(n) for n in Nxyz), generator=gen, device=dev, dtype=torch.float32)
        return cls(uvw, dxyz, lowpass_width=lowpass_width)

    @classmethod
    def isotropic_unit(cls, D, seed=1, device="cuda:0", Nxyz=(128, 64, 64), noise=None):
        """Box of the wake-added turbulence (``SynchronizedAutoScalingIsotropicMannTurbulence``): isotropic (Gamma = 0),
        small scales (L = D/8, cells of D/16), normalised to unit standard deviation of u."""
        box = cls.generate(1.0, D / 8.0, 0.0, Nxyz=Nxyz, dxyz=(D / 16.0,) * 3, seed=seed, device=device, noise=noise,
                           lowpass_width=D / 16.0)
        box.raw /= box.std_u
        box.lp /= box.std_u
        box.std_u = 1.0
        return box

    @classmethod
    def from_file(cls, path, device="cuda:0", dxyz=None, lowpass_width=160.0):
        """``MannTurbulenceField.from_netcdf`` (Wind_Farm_Env.py:614-617).  ``.npy`` ([3,Nx,Ny,Nz], needs ``dxyz``),
        ``.npz`` (arrays ``uvw`` and ``dxyz``), NetCDF-3 through scipy, or NetCDF-4 / HDF5 as hipersim writes it through the
        built-in reader ``windgym_b200.hdf5_min`` (contiguous / chunked / deflate-compressed numeric datasets)."""
        path = str(path)
        if path.endswith(".npy"):
            uvw, d = np.load(path), dxyz
        elif path.endswith(".npz"):
            z = np.load(path)
            uvw, d = z["uvw"], tuple(z["dxyz"]) if "dxyz" in z.files else dxyz
        else:
            with open(path, "rb") as fh:
                magic = fh.read(8)
            if magic[:3] == b"CDF":          # classic NetCDF-3
                from scipy.io import netcdf_file
                with netcdf_file(path, "r", mmap=False) as nc:
                    uvw = np.array(nc.variables["uvw"][:])
                    d = tuple(float(nc.variables[a][1] - nc.variables[a][0]) for a in ("x", "y", "z"))
            else:                            # NetCDF-4 = HDF5: the built-in minimal reader (windgym_b200/hdf5_min.py)
                from .hdf5_min import read_mann_box
                uvw, d = read_mann_box(path)
        if d is None:
            raise ValueError("dxyz is required for a bare .npy box")
        return cls(torch.as_tensor(np.ascontiguousarray(uvw), device=torch.device(device)), d, lowpass_width=lowpass_width)

"""TEST INFRASTRUCTURE -- fp64 numpy restatement of the ``dynamiks`` seam of DTUWindEnergy/WindGym.

PARITY UNPINNED.  ``dynamiks@77f4f875c6401fde150e351a60143a5b522cf950`` (``/root/reference/pyproject.toml:25``)
is not vendored in the reference and not installed here; the reference's tests pin no numeric flow value
(SURVEY.md section 8c).  This file restates the *published* Dynamic Wake Meandering algorithm behind the
exact object protocol the reference env uses (``Wind_Farm_Env.py:11-21`` imports; ctor ``:702-711``;
``run`` ``:734``; ``step`` ``:745,:945``; ``yaw`` ``:715,:832-858``; ``rotor_avg_windspeed`` ``:486-490``;
``power()`` ``:495,:539``; ``positions_xyz`` ``:541``; ``rotor_positions_xyz`` ``:723``; ``time`` ``:800``).
It is the frozen specification the CUDA path (``windgym_b200/csrc``) is checked against.

Model (all lengths non-dimensionalised by the rotor radius R, wake velocities by the particle's emission
inflow U0e):

* Frame: wind aligned, theta = 270deg - wd; x' = dx cos(theta) + dy sin(theta), y' = -dx sin(theta) + dy cos(theta)
  about the layout centroid (matches the positions printed in the reference notebook).
* Ambient: uniform (ws, 0, 0) (``TurbulenceFieldSite`` over ``RandomTurbulence(ti=0)`` -- the deterministic
  ``turbtype="None"`` path, ``Wind_Farm_Env.py:661-665``), or (ws, 0, 0) + a frozen Mann box advected with ws
  (``oracle/mann_numpy.py``; ``MannLoad`` / ``MannGenerate`` / ``MannFixed``, ``Wind_Farm_Env.py:612-659``): the
  low-pass filtered (v', w') at a particle's centre move it (meandering, Larsen et al. 2008), the raw box averaged
  over the rotor's quadrature points is added to the rotor inflow.  Wake-added turbulence (Madsen et al. 2010):
  inside wakes the rotor points gain w U0e k_mt(r) times a unit-variance isotropic box, k_mt = 0.6 |1-U| + 0.35 |dU/dr|.
* Turbine: tabular P/CT, py_wake ``SimpleYawModel``: P = P_tab(u cos g), CT = CT_tab(u cos g) cos^2 g,
  induction a = (1 - sqrt(1 - CT)) / 2.
* Wake particles (Larsen et al. 2008, Wind Energy 11:377): one chain per turbine; a particle is released at
  the rotor centre every k_emit = ceil(d_particle D / (ws dt)) steps (the uniform-speed limit of "every
  d_particle*D of travel"; keeps emission deterministic).  Particle velocity = ambient + Hill-vortex
  self-induced velocity 0.4 dUc U0e (-cos g0, +sin g0, 0) (Larsen et al. 2020, J. Phys. Conf. Ser.
  1618:062047), dUc = centre-line deficit of the particle's current profile.
* Deficit: each particle carries an axisymmetric profile U(r) on 64 radial nodes (dr = R/16) obeying the
  thin-shear-layer equations U dU/dx + V dU/dr = nu/r d/dr(r dU/dr), dU/dx + 1/r d(rV)/dr = 0
  (Ainslie 1988, J. Wind Eng. Ind. Aerodyn. 27:213), marched implicitly (tridiagonal in r, coefficients
  frozen at the old level; jDWM layout) by the particle's own axial displacement every step; V from
  continuity, solved consistently with the momentum equation at the old level.
  Inlet (IEC 61400-1 ed.4 Annex E): U_w = 1 - 2a inside R_w = f_w sqrt((1-a)/(1-2a)), f_w = 1 - 0.45 a^2
  (cell-averaged top hat so it is continuous in a).
  Eddy viscosity (IEC 61400-1 ed.4 Annex E.2 / Madsen et al. 2010, J. Sol. Energy Eng. 132:041014):
  nu = 0.023 F1(x) TI^0.3 + 0.016 F2(x) (b/R)(1 - Umin), with b(1-Umin) = sqrt(2 M (1-Umin)),
  M = int (1-U) r dr (top-hat-equivalent width).
* Superposition: linear sum over upstream chains; at the rotor plane x_j the two bracketing particles
  (consecutive ages) are interpolated linearly in x (centre and profile); the deficit vector points along
  the emitting rotor's axis (-cos g0, +sin g0).  Rotor average: 16-point equal-area polar quadrature.
"""
import numpy as np

N_R = 64
DR = 1.0 / 16.0
K_HILL = 0.4
K1 = 0.023
K2 = 0.016
CT_MAX = 0.96
MARGIN_D = 2.0
N_Q = 16
DXT_MIN = 1e-6
K_M1 = 0.6    # wake-added turbulence, Madsen et al. (2010) / IEC 61400-1 ed.4 Annex E: k_mt = k_m1 |1-U| + k_m2 |dU/dr|
K_M2 = 0.35


def rotor_points():
    """16 (dy, dz) offsets in rotor radii: 4 equal-area rings x 4 azimuths, equal weights 1/16."""
    pts = np.empty((N_Q, 2))
    for k in range(4):
        rho = np.sqrt((k + 0.5) / 4.0)
        for m in range(4):
            phi = 2.0 * np.pi * (m + 0.5 * (k & 1)) / 4.0 + np.pi / 8.0
            pts[4 * k + m] = (rho * np.cos(phi), rho * np.sin(phi))
    return pts


def f1_filter(xt):
    """IEC 61400-1 ed.4 (E.5): ambient-turbulence filter, xt in rotor radii."""
    s = np.clip(xt / 8.0, 0.0, 1.0) ** 1.5
    return np.where(xt >= 8.0, 1.0, s - np.sin(2.0 * np.pi * s) / (2.0 * np.pi))


def f2_filter(xt):
    """IEC 61400-1 ed.4 (E.6): shear-layer filter, xt in rotor radii."""
    lin = 0.025 * xt - 0.0375
    return np.where(
        xt < 4.0, 0.0625, np.where(xt < 12.0, lin, np.where(xt < 20.0, 0.00105 * (xt - 12.0) ** 3 + lin, 1.0))
    )


def inlet_profile(a):
    """Cell-averaged top-hat inlet U(r) for induction ``a`` (array [n]) -> [n, 64]."""
    a = np.asarray(a, dtype=np.float64)
    fw = 1.0 - 0.45 * a * a
    rw2 = fw * fw * (1.0 - a) / (1.0 - 2.0 * a)
    j = np.arange(N_R)
    rlo = np.maximum(j - 0.5, 0.0) * DR
    rhi = (j + 0.5) * DR
    frac = np.clip((rw2[:, None] - rlo[None] ** 2) / (rhi**2 - rlo**2)[None], 0.0, 1.0)
    U = 1.0 - 2.0 * a[:, None] * frac
    U[:, -1] = 1.0
    return U


def ainslie_march(U, dxt, xt, knu1):
    """Advance profiles U [n,64] by dxt [n] rotor radii at downstream distance xt [n]; returns (U_new, nu)."""
    n = U.shape[0]
    r = np.arange(N_R) * DR
    idr2 = 1.0 / (DR * DR)
    # ---- pass A: Laplacian, continuity-consistent radial velocity (per unit nu), integrals for nu
    Up = np.zeros_like(U)
    L = np.zeros_like(U)
    Up[:, 1:-1] = (U[:, 2:] - U[:, :-2]) * (0.5 / DR)
    L[:, 1:-1] = (U[:, 2:] + U[:, :-2] - 2.0 * U[:, 1:-1]) * idr2 + Up[:, 1:-1] / r[1:-1]
    L[:, 0] = 4.0 * (U[:, 1] - U[:, 0]) * idr2
    Vh = np.zeros_like(U)
    I = np.zeros(n)
    rg_prev = np.zeros(n)
    for j in range(1, N_R - 1):
        Ip = I + 0.5 * DR * rg_prev
        den = U[:, j] - 0.5 * DR * Up[:, j]
        g = (L[:, j] + Up[:, j] * Ip / r[j]) / den
        rg = r[j] * g
        I = Ip + 0.5 * DR * rg
        Vh[:, j] = -I / r[j]
        rg_prev = rg
    M = np.sum((1.0 - U) * r[None], axis=1) * DR
    dmin = 1.0 - U.min(axis=1)
    nu = knu1 * f1_filter(xt) + K2 * f2_filter(xt) * np.sqrt(np.maximum(2.0 * M * dmin, 0.0))
    # ---- pass B: tridiagonal rows (unknowns j = 0..62, U[63] = 1 Dirichlet), Thomas forward sweep
    idx = 1.0 / np.maximum(dxt, DXT_MIN)
    cp = np.zeros_like(U)
    dp = np.zeros_like(U)
    b0 = U[:, 0] * idx + 4.0 * nu * idr2
    cp[:, 0] = (-4.0 * nu * idr2) / b0
    dp[:, 0] = (U[:, 0] * U[:, 0] * idx) / b0
    for j in range(1, N_R - 1):
        am = (1.0 - 0.5 * DR / r[j]) * idr2
        ap = (1.0 + 0.5 * DR / r[j]) * idr2
        Vd = nu * Vh[:, j] * (0.5 / DR)
        a = -nu * am - Vd
        c = -nu * ap + Vd
        b = U[:, j] * idx + 2.0 * nu * idr2
        d = U[:, j] * U[:, j] * idx
        if j == N_R - 2:
            d = d - c  # Dirichlet U[63] = 1
        m = 1.0 / (b - a * cp[:, j - 1])
        cp[:, j] = c * m
        dp[:, j] = (d - a * dp[:, j - 1]) * m
    # ---- pass C: back substitution
    Un = np.ones_like(U)
    Un[:, N_R - 2] = dp[:, N_R - 2]
    for j in range(N_R - 3, -1, -1):
        Un[:, j] = dp[:, j] - cp[:, j] * Un[:, j + 1]
    return Un, nu


def profile_gradient_at(U, r):
    """|dU/dr| (per rotor radius) of the cell that holds r: the slope of the linear interpolant; zero beyond the grid."""
    s = r / DR
    j0 = np.minimum(np.floor(s).astype(np.int64), N_R - 2)
    rows = np.arange(U.shape[0])[:, None]
    return np.where(s >= N_R - 1, 0.0, np.abs(U[rows, j0 + 1] - U[rows, j0]) / DR)


def profile_deficit_at(U, r):
    """Deficit 1-U at radius r (rotor radii) by linear interpolation; zero beyond the grid. U [n,64], r [n,q]."""
    s = r / DR
    j0 = np.minimum(np.floor(s).astype(np.int64), N_R - 2)
    f = s - j0
    rows = np.arange(U.shape[0])[:, None]
    val = (1.0 - U[rows, j0]) * (1.0 - f) + (1.0 - U[rows, j0 + 1]) * f
    return np.where(s >= N_R - 1, 0.0, val)


# ------------------------------------------------------------------------------------------------
# dynamiks object protocol
# ------------------------------------------------------------------------------------------------
class RandomTurbulence:
    """``dynamiks.sites.turbulence_fields.RandomTurbulence`` stand-in.  Only ti == 0 (uniform inflow, the
    reference's deterministic ``turbtype="None"`` site, ``Wind_Farm_Env.py:664``) is restated."""

    def __init__(self, ti=0.0, ws=10.0, seed=None):
        self.ti, self.ws, self.seed = float(ti), float(ws), seed


class MannTurbulenceField:
    """Placeholder: Mann-box inflow is SURVEY.md section 8 row f-1 (next), not part of this oracle yet."""

    @staticmethod
    def generate(*a, **k):
        raise NotImplementedError("Mann turbulence boxes are not restated (SURVEY.md 8 f-1)")

    from_netcdf = generate


class TurbulenceFieldSite:
    def __init__(self, ws, turbulenceField):
        self.ws = float(ws)
        self.turbulenceField = turbulenceField
        self.ti = float(getattr(turbulenceField, "ti", 0.0))


class jDWMAinslieGenerator:  # marker objects: the model they select is the one restated above
    pass


class HillVortexParticleMotion:
    pass


class PyWakeWindTurbines:
    """``dynamiks.wind_turbines.PyWakeWindTurbines`` stand-in (layout frame x east / y north)."""

    def __init__(self, x, y, windTurbine):
        self.x = np.asarray(x, dtype=np.float64).copy()
        self.y = np.asarray(y, dtype=np.float64).copy()
        self.windTurbine = windTurbine
        self.N = self.x.size
        self._yaw = np.zeros(self.N)
        self._uvw = None
        self._fs = None
        self.types = np.zeros(self.N, dtype=int)

    # yaw offset [deg] relative to the global wind direction; assignment copies (no aliasing between farms)
    @property
    def yaw(self):
        return self._yaw

    @yaw.setter
    def yaw(self, value):
        v = np.array(value, dtype=np.float64).reshape(-1)
        self._yaw = np.full(self.N, v[0]) if v.size == 1 else v.copy()

    def yaw_tilt(self):
        return self._yaw, np.zeros(self.N)

    def hub_height(self):
        return self.windTurbine.hub_height()

    def diameter(self):
        return self.windTurbine.diameter()

    @property
    def positions_xyz(self):
        return self._fs.positions_xyz

    @property
    def rotor_positions_xyz(self):
        return self._fs.positions_xyz

    @property
    def rotor_avg_windspeed(self):
        return self._fs.rotor_avg_windspeed

    # Extension (no reference counterpart; BASELINE.json cfg 4 "yaw + induction actions"): ``derate`` scales every
    # rotor's axial induction, a = delta a_tab: CT = 4a(1-a) cos^2(yaw), P = P_tab a(1-a)^2 / (a_tab (1-a_tab)^2).
    # delta = 1 is the reference's turbine (and takes the reference's code path, bit for bit).
    def _derated(self):
        u = self._fs.rotor_avg_windspeed[:, 0]
        co = np.cos(np.deg2rad(self._yaw))
        p_tab = self.windTurbine.power(u * co, yaw=0.0)
        ct_tab = np.clip(self.windTurbine.ct(u * co, yaw=0.0), 0.0, CT_MAX)
        a0 = 0.5 * (1.0 - np.sqrt(1.0 - ct_tab))
        a1 = self.derate * a0
        with np.errstate(divide="ignore", invalid="ignore"):
            scale = np.where(a0 > 0, (a1 * (1 - a1) ** 2) / (a0 * (1 - a0) ** 2), 1.0)
        return p_tab * scale, 4.0 * a1 * (1.0 - a1) * co**2

    def power(self):
        der = getattr(self, "derate", None)
        if der is not None and np.any(der != 1.0):
            p, _ = self._derated()
            full = self.windTurbine.power(self._fs.rotor_avg_windspeed[:, 0], yaw=self._yaw)
            return np.where(der != 1.0, p, full)
        u = self._fs.rotor_avg_windspeed[:, 0]
        return self.windTurbine.power(u, yaw=self._yaw)

    def ct(self):
        u = self._fs.rotor_avg_windspeed[:, 0]
        full = np.minimum(self.windTurbine.ct(u, yaw=self._yaw), CT_MAX)
        der = getattr(self, "derate", None)
        if der is not None and np.any(der != 1.0):
            _, c = self._derated()
            return np.where(der != 1.0, np.minimum(c, CT_MAX), full)
        return full


def rotate_layout(x, y, wd):
    """Layout (east, north) -> wind-aligned frame about the centroid.  Returns (x', y')."""
    th = np.deg2rad(270.0 - wd)
    dx = x - x.mean()
    dy = y - y.mean()
    return dx * np.cos(th) + dy * np.sin(th), -dx * np.sin(th) + dy * np.cos(th)


def emission_cadence(d_particle, D, ws, dt):
    return max(1, int(np.ceil(d_particle * D / (ws * dt) - 1e-9)))


def chain_capacity(extent, D, d_particle, ws_min_spacing, f_min):
    """Slots per chain that can never overflow: farm extent + margin at the slowest possible spacing."""
    n = int(np.ceil((extent + MARGIN_D * D) / (ws_min_spacing * f_min))) + 3
    return (n + 7) // 8 * 8


class DWMFlowSimulation:
    """Restated ``dynamiks.dwm.DWMFlowSimulation`` (reference ctor call ``Wind_Farm_Env.py:702-711``)."""

    def __init__(self, site, windTurbines, wind_direction=270.0, particleDeficitGenerator=None, dt=1,
                 d_particle=0.2, particleMotionModel=None, addedTurbulenceModel=None, p_cap=None, emit_rule="cadence",
                 **_):
        # emit_rule (sensitivity studies only, tests/test_dwm_oracle_physics.py; the frozen specification -- and the
        # CUDA kernel -- is "cadence"):
        #   "cadence"  one particle per chain every k_emit = ceil(d_particle D / (ws dt)) steps
        #   "distance" per chain, at the first step boundary where its newest particle has travelled >= d_particle D
        #   int k      one particle per chain every k steps
        self.emit_rule = emit_rule
        self.d_emit = float(d_particle) * float(windTurbines.diameter())
        wt = windTurbines
        self.site, self.windTurbines = site, wt
        wt._fs = self
        self.wind_direction = float(wind_direction)
        self.dt = float(dt)
        self.time = 0.0
        self.n_step = 0
        self.ws = float(site.ws)
        self.ti = float(site.ti)
        self.T = T = wt.N
        self.D = float(wt.diameter())
        self.R = 0.5 * self.D
        xr, yr = rotate_layout(wt.x, wt.y, self.wind_direction)
        self.xr, self.yr = xr, yr
        self.zh = float(wt.hub_height())
        self.positions_xyz = np.stack([xr, yr, np.full(T, self.zh)])
        self.xmax = xr.max()
        self.k_emit = emission_cadence(d_particle, self.D, self.ws, self.dt)
        if isinstance(emit_rule, (int, np.integer)) and not isinstance(emit_rule, bool):
            self.k_emit = max(1, int(emit_rule))
        tab_ct = getattr(wt.windTurbine, "ct_table", np.array([0.9]))
        a_max = 0.5 * (1.0 - np.sqrt(1.0 - min(float(np.max(tab_ct)), CT_MAX)))
        self.f_min = 1.0 - K_HILL * 2.0 * a_max
        if p_cap is None:
            p_cap = chain_capacity(xr.max() - xr.min(), self.D, d_particle, self.k_emit * self.ws * self.dt, self.f_min)
        self.P = P = int(p_cap)
        # particle state, slot-addressed ring per chain (this is the layout the device mirrors)
        self.prof = np.ones((T, P, N_R))
        self.pmut = np.zeros((T, P, 4))  # x, y, z, uc
        self.pcon = np.zeros((T, P, 4))  # U0e, knu1, cos g0, sin g0
        self.head = np.zeros(T, dtype=np.int64)  # next slot to write
        self.count = np.zeros(T, dtype=np.int64)
        self.overflow = 0
        self.rotor_avg_windspeed = np.tile(np.array([self.ws, 0.0, 0.0]), (T, 1))
        self.qpts = rotor_points()
        self.last_nu = None
        tf = getattr(site, "turbulenceField", None)
        self.tf = tf if hasattr(tf, "sample") else None  # Mann box (oracle/mann_numpy.py); None = uniform inflow
        # wake-added turbulence (SynchronizedAutoScalingIsotropicMannTurbulence, Wind_Farm_Env.py:618): an object with
        # a unit-variance isotropic ``field`` (mann_numpy.MannTurbulenceField) advected with the ambient box
        self.added = getattr(addedTurbulenceModel, "field", None) if self.tf is not None else None

    # -- helpers --------------------------------------------------------------------------------
    def slots_by_age(self, t):
        """Slot indices of chain t from youngest (age 0) to oldest."""
        return (self.head[t] - 1 - np.arange(self.count[t])) % self.P

    def run(self, t):
        for _ in range(int(round(t / self.dt))):
            self.step()

    # -- one DWM step ---------------------------------------------------------------------------
    def step(self):
        T, P, R, dt = self.T, self.P, self.R, self.dt
        wt = self.windTurbines
        # 1. retire particles that will be past the farm (+margin) after this step's move
        for t in range(T):
            while self.count[t] > 0:
                s = (self.head[t] - self.count[t]) % P
                x, _, _, uc = self.pmut[t, s]
                U0e, _, cg, _ = self.pcon[t, s]
                if x + (self.ws - K_HILL * (1.0 - uc) * U0e * cg) * dt > self.xmax + MARGIN_D * self.D:
                    self.count[t] -= 1
                else:
                    break
        # 2. move + march every live particle
        live = np.zeros((T, P), dtype=bool)
        for t in range(T):
            live[t, self.slots_by_age(t)] = True
        ti, si = np.nonzero(live)
        if ti.size:
            U = self.prof[ti, si]
            x, y, z, uc = self.pmut[ti, si].T
            U0e, knu1, cg, sg = self.pcon[ti, si].T
            duc = 1.0 - uc
            vx = self.ws - K_HILL * duc * U0e * cg
            vy = K_HILL * duc * U0e * sg
            vz = np.zeros_like(vy)
            if self.tf is not None:  # meandering: the large-scale lateral / vertical fluctuations carry the wake centre
                vl, wl = self.tf.sample_lp(x, y, z, self.time, self.ws)
                vy = vy + vl
                vz = wl
            dx = vx * dt
            xt_mid = (x + 0.5 * dx - self.xr[ti]) / R
            Un, nu = ainslie_march(U, dx / R, xt_mid, knu1)
            self.last_nu = nu
            self.prof[ti, si] = Un
            self.pmut[ti, si] = np.stack([x + dx, y + vy * dt, z + vz * dt, Un[:, 0]], axis=1)
        # 3. rotor inflow: ambient minus superposed upstream deficits
        du = np.zeros(T)
        dv = np.zeros(T)
        kmt = np.zeros((T, N_Q))  # sum over wakes of w U0e k_mt at every rotor point (wake-added turbulence)
        for i in range(T):
            n = self.count[i]
            if n < 2:
                continue
            sl = self.slots_by_age(i)
            xa, ya, za = self.pmut[i, sl, 0], self.pmut[i, sl, 1], self.pmut[i, sl, 2]
            xp, xn = xa[:-1, None], xa[1:, None]  # younger, older
            xj = self.xr[None, :]
            up = (xp <= xj) & (xj < xn)
            dn = (xn <= xj) & (xj < xp)
            sign = up.astype(np.float64) - dn.astype(np.float64)
            sign[:, i] = 0.0
            pi_, ji = np.nonzero(sign)
            if pi_.size == 0:
                continue
            sg_ = sign[pi_, ji]
            w = (self.xr[ji] - xa[pi_]) / (xa[pi_ + 1] - xa[pi_])
            yc = ya[pi_] * (1 - w) + ya[pi_ + 1] * w
            zc = za[pi_] * (1 - w) + za[pi_ + 1] * w
            ry = (self.yr[ji][:, None] - yc[:, None]) / R + self.qpts[None, :, 0]
            rz = (self.zh - zc[:, None]) / R + self.qpts[None, :, 1]
            rq = np.sqrt(ry * ry + rz * rz)
            for side, wgt in ((0, 1.0 - w), (1, w)):
                s_ = sl[pi_ + side]
                Dq = profile_deficit_at(self.prof[i, s_], rq).mean(axis=1)
                U0e, _, cg, sg0 = self.pcon[i, s_].T
                np.add.at(du, ji, sg_ * wgt * U0e * cg * Dq)
                np.add.at(dv, ji, sg_ * wgt * U0e * sg0 * Dq)
                if self.added is not None:
                    kq = K_M1 * profile_deficit_at(self.prof[i, s_], rq) + K_M2 * profile_gradient_at(self.prof[i, s_], rq)
                    np.add.at(kmt, ji, (sg_ * wgt * U0e)[:, None] * kq)
        amb = np.zeros((3, T))
        if self.tf is not None:  # rotor-averaged ambient fluctuation at the new time level
            py = self.yr[:, None] + self.qpts[None, :, 0] * R
            pz = self.zh + self.qpts[None, :, 1] * R
            px = np.broadcast_to(self.xr[:, None], py.shape)
            amb = self.tf.sample(px, py, pz, self.time + dt, self.ws).mean(axis=2)
            if self.added is not None:
                self.added.offset = self.tf.offset
                amb = amb + (kmt[None] * self.added.sample(px, py, pz, self.time + dt, self.ws)).mean(axis=2)
        self.rotor_avg_windspeed = np.stack([self.ws + amb[0] - du, amb[1] + dv, amb[2]], axis=1)
        # 4./5. turbine update and particle release
        if self.emit_rule == "distance":
            newest = (self.head - 1) % P
            due = (self.count == 0) | (self.pmut[np.arange(T), newest, 0] - self.xr >= self.d_emit)
        else:
            due = np.full(T, self.n_step % self.k_emit == 0)
        if due.any():
            u = self.rotor_avg_windspeed[:, 0]
            ct = np.clip(wt.ct(), 0.0, CT_MAX)
            a = 0.5 * (1.0 - np.sqrt(1.0 - ct))
            Uin = inlet_profile(a)
            g = np.deg2rad(wt.yaw)
            for t in range(T):
                if not due[t]:
                    continue
                s = self.head[t]
                if self.count[t] == P:
                    self.overflow += 1
                else:
                    self.count[t] += 1
                self.prof[t, s] = Uin[t]
                self.pmut[t, s] = (self.xr[t], self.yr[t], self.zh, Uin[t, 0])
                self.pcon[t, s] = (u[t], K1 * self.ti**0.3 if self.ti > 0 else 0.0, np.cos(g[t]), np.sin(g[t]))
                self.head[t] = (s + 1) % P
        self.n_step += 1
        self.time += dt

    # -- render API (``Wind_Farm_Env.py:1056``) --------------------------------------
    def wind_at_points(self, x, y, z):
        """Wake-superposed (u, v, w) at points (x, y, z) of the wind-aligned frame: the superposition rule of
        ``step`` (bracketing pair of consecutive ages, linear interpolation in x of centre and profile, deficit
        along the emitting rotor's axis) evaluated at a point instead of over a rotor's quadrature points."""
        x, y = np.asarray(x, dtype=np.float64).reshape(-1), np.asarray(y, dtype=np.float64).reshape(-1)
        z = np.broadcast_to(np.asarray(z, dtype=np.float64), x.shape)
        du, dv = np.zeros(x.size), np.zeros(x.size)
        for i in range(self.T):
            if self.count[i] < 2:
                continue
            sl = self.slots_by_age(i)
            xa, ya, za = self.pmut[i, sl, 0], self.pmut[i, sl, 1], self.pmut[i, sl, 2]
            xp, xn = xa[:-1, None], xa[1:, None]  # younger, older
            up = (xp <= x[None]) & (x[None] < xn)
            dn = (xn <= x[None]) & (x[None] < xp)
            sign = up.astype(np.float64) - dn.astype(np.float64)
            pi_, ji = np.nonzero(sign)
            if pi_.size == 0:
                continue
            sg_ = sign[pi_, ji]
            w = (x[ji] - xa[pi_]) / (xa[pi_ + 1] - xa[pi_])
            yc = ya[pi_] * (1 - w) + ya[pi_ + 1] * w
            zc = za[pi_] * (1 - w) + za[pi_ + 1] * w
            rq = np.sqrt(((y[ji] - yc) / self.R) ** 2 + ((z[ji] - zc) / self.R) ** 2)[:, None]
            for side, wgt in ((0, 1.0 - w), (1, w)):
                s_ = sl[pi_ + side]
                Dq = profile_deficit_at(self.prof[i, s_], rq)[:, 0]
                U0e, _, cg, sg0 = self.pcon[i, s_].T
                np.add.at(du, ji, sg_ * wgt * U0e * cg * Dq)
                np.add.at(dv, ji, sg_ * wgt * U0e * sg0 * Dq)
        amb = self.tf.sample(x, y, z, self.time, self.ws) if self.tf is not None else np.zeros((3, x.size))
        return np.stack([self.ws + amb[0] - du, amb[1] + dv, amb[2]])

    def get_windspeed(self, view, include_wakes=True, xarray=False):
        """``[3, len(view.x), len(view.y)]`` on an XYView-like object (attributes x, y, z)."""
        gx, gy = np.meshgrid(np.asarray(view.x, dtype=np.float64), np.asarray(view.y, dtype=np.float64), indexing="ij")
        if not include_wakes:
            out = np.zeros((3,) + gx.shape)
            out[0] = self.ws
            return out
        return self.wind_at_points(gx, gy, float(np.asarray(view.z).reshape(-1)[0])).reshape((3,) + gx.shape)

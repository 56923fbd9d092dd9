"""Seeded synthetic inputs for the hot path, as host numpy arrays.

A `Case` is plain data -- sizes, physical settings and parent arrays (C-order (sy, sx), i fastest)
-- so the same arrays can be fed to the GPU library and to the CPU oracle.  Configurations:

  anticyclone_case(N): BASELINE config 1/2 -- examples/ice_advected_by_anticyclone.jl:35-126 scaled
      to N x N (Bounded x Bounded, dx = 4 km, FPlane, wind-stress arrays, SemiImplicitStress ocean drag).
  periodic_case(N):    BASELINE config 3 -- doubly periodic variant with every mask branch live.
  coastline_case(Ny):  BASELINE config 4 -- examples/ice_advected_on_coastline.jl with its immersed coastline.
  latlon_case(N):      a basin on a LatitudeLongitudeGrid (j-dependent metrics).
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field

import numpy as np

from .metrics import latitude_longitude_metrics, spherical_coriolis_f_ff

SEED = 20260417


@dataclass
class Case:
    name: str
    Nx: int
    Ny: int
    Hx: int
    Hy: int
    topology: tuple          # ("Periodic"|"Bounded", "Periodic"|"Bounded")
    Lx: float
    Ly: float
    dt: float = 120.0
    substeps: int = 150
    coriolis_f: float | None = 1e-4
    advection_order: int = 7
    timestepper: str = "SplitRungeKutta3"
    u_bc_value: float | None = None   # ValueBoundaryCondition on u north/south (Bounded y)
    v_bc_value: float | None = None   # ValueBoundaryCondition on v west/east (Bounded x)
    rho_e: float = 1026.0
    Cd: float = 5.5e-3
    fields: dict = field(default_factory=dict)   # h, a, u, v, top_x, top_y, ue, ve parents
    top_const: tuple | None = None       # constant (tau_x, tau_y) when there are no top_x/top_y arrays
    ocean_const: tuple | None = None     # constant (ue, ve) when there are no ue/ve arrays
    mask: object = None                  # immersed mask at centres, uint8 (Ny+2Hy, Nx+2Hx), 1 = land
    immersed_drag: tuple = (0.0, 0.0)    # linear-drag immersed flux BC coefficients for u and v
    top_kind: str = "auto"               # "auto": arrays -> (u=Field, v=Field), top_const -> numbers, else nothing;
                                         # "semi_implicit": top_x/top_y (or top_const) are the atmosphere's u_e, v_e
    top_rho_Cd: tuple = (1.3, 1.2e-3)    # rho_e, Cd of a top SemiImplicitStress
    bottom_kind: str = "semi_implicit"   # "semi_implicit" (ue/ve or ocean_const = ocean velocity), "stress" (they hold tau), "none"
    free_drift: str | None = None        # None, "fields" (fields fd_u, fd_v) or "stress_balance"
    latlon: tuple | None = None          # ((lon0, lon1), (lat0, lat1)) in degrees: LatitudeLongitudeGrid; Lx, Ly = extents in degrees
    rotation_rate: float | None = None   # lat-lon only: HydrostaticSphericalCoriolis(rotation_rate) instead of FPlane(coriolis_f)
    metric_arrays: dict | None = None    # explicit j-indexed metric arrays (a slab's rows of the global grid's metrics)
    fold: dict | None = None             # Folded y axis: dict(maps={(lx, ly): (target, source)}, sign_velocity, sign_external)
    thermo: dict | None = None           # slab thermodynamics on top of the dynamics: scalars bottom_heat_flux, ice_salinity;
                                         # arrays Tu (initial top temperature) and Qtop (external top heat flux) live in `fields`

    def metrics(self):
        """j-indexed metric arrays of a lat-lon case (None on a RectilinearGrid)."""
        if self.metric_arrays is not None:
            return {k: v for k, v in self.metric_arrays.items() if k != "f_ff"}
        if self.latlon is None:
            return None
        return latitude_longitude_metrics(self.Nx, self.Ny, self.Hy, self.latlon[0], self.latlon[1])

    def f_ff(self):
        """Per-row Coriolis parameter at (Face, Face) of a spherical-Coriolis case (None otherwise)."""
        if self.rotation_rate is None or self.latlon is None:
            return None
        if self.metric_arrays is not None and "f_ff" in self.metric_arrays:
            return self.metric_arrays["f_ff"]
        return spherical_coriolis_f_ff(self.Ny, self.Hy, self.latlon[1], self.rotation_rate)

    @property
    def dx(self):
        return self.Lx / self.Nx

    @property
    def dy(self):
        return self.Ly / self.Ny

    def bounded_like(self, axis):
        return self.topology[axis] in ("Bounded", "Folded")

    def parent_shape(self, loc):
        sx = self.Nx + 2 * self.Hx + (1 if (loc[0] and self.topology[0] == "Bounded") else 0)
        sy = self.Ny + 2 * self.Hy + (1 if (loc[1] and self.topology[1] in ("Bounded", "Folded")) else 0)
        return sy, sx

    def nodes(self, loc):
        sy, sx = self.parent_shape(loc)
        i = np.arange(sx) - self.Hx + 1
        j = np.arange(sy) - self.Hy + 1
        x = ((i - 1) if loc[0] else (i - 0.5)) * self.dx
        y = ((j - 1) if loc[1] else (j - 0.5)) * self.dy
        return np.meshgrid(x, y)


LOC = dict(u=(1, 0), v=(0, 1), h=(0, 0), a=(0, 0), top_x=(1, 0), top_y=(0, 1), ue=(1, 0), ve=(0, 1), hs=(0, 0),
           fd_u=(1, 0), fd_v=(0, 1), Tu=(0, 0), Qtop=(0, 0))


def _wrap_periodic(case: Case, arr, loc):
    """Fill halos of a host array with periodic images along Periodic axes (initial state only)."""
    Hx, Hy, Nx, Ny = case.Hx, case.Hy, case.Nx, case.Ny
    if case.topology[0] == "Periodic":
        arr[:, :Hx] = arr[:, Nx:Nx + Hx]
        arr[:, Nx + Hx:Nx + 2 * Hx] = arr[:, Hx:2 * Hx]
    if case.topology[1] == "Periodic":
        arr[:Hy, :] = arr[Ny:Ny + Hy, :]
        arr[Ny + Hy:Ny + 2 * Hy, :] = arr[Hy:2 * Hy, :]
    return arr


def periodic_case(N, Ny=None, H=7, seed=SEED, aice="mixed", substeps=150, dt=120.0, advection_order=7,
                  timestepper="SplitRungeKutta3", moving=True) -> Case:
    """Doubly periodic domain, L = N * 4 km, smooth fields + seeded noise."""
    Ny = N if Ny is None else Ny
    c = Case("periodic", N, Ny, H, H, ("Periodic", "Periodic"), N * 4000.0, Ny * 4000.0, dt=dt, substeps=substeps,
             advection_order=advection_order, timestepper=timestepper)
    rng = np.random.default_rng(seed)
    tp = 2 * np.pi
    X, Y = c.nodes(LOC["h"])
    h = 0.3 + 0.005 * (np.sin(3 * tp * X / c.Lx) + np.sin(2 * tp * Y / c.Ly)) + 1e-3 * rng.uniform(-1, 1, X.shape)
    if aice == "ones":
        a = np.ones_like(h)
    else:
        a = 0.9 + 0.1 * rng.uniform(0, 1, X.shape)
        r = rng.uniform(0, 1, X.shape)
        # a 5 % patch of marginal ice (below minimum_concentration) and 2 % of open water
        patch = (np.abs(X / c.Lx - 0.3) < 0.11) & (np.abs(Y / c.Ly - 0.6) < 0.11)
        a = np.where(patch & (r < 0.9), 5e-4 * r, a)
        zero = (np.abs(X / c.Lx - 0.7) < 0.07) & (np.abs(Y / c.Ly - 0.25) < 0.07)
        a = np.where(zero, 0.0, a)
        h = np.where(zero, 0.0, h)
    Xu, Yu = c.nodes(LOC["u"])
    Xv, Yv = c.nodes(LOC["v"])
    amp = 0.05 if moving else 0.0
    u = amp * np.sin(tp * Yu / c.Ly) * np.cos(tp * Xu / c.Lx)
    v = -amp * np.sin(tp * Xv / c.Lx) * np.cos(tp * Yv / c.Ly)
    ue = 0.01 * np.sin(tp * Yu / c.Ly)
    ve = 0.01 * np.sin(tp * Xv / c.Lx)
    tx = 0.1 * np.sin(tp * Yu / c.Ly)
    ty = 0.1 * np.cos(tp * Xv / c.Lx)
    raw = dict(h=h, a=a, u=u, v=v, ue=ue, ve=ve, top_x=tx, top_y=ty)
    c.fields = {k: _wrap_periodic(c, np.ascontiguousarray(vv, dtype=np.float64), LOC[k]) for k, vv in raw.items()}
    return c


def anticyclone_case(N, H=7, seed=SEED, substeps=150, dt=120.0, advection_order=7, noise=1e-3,
                     timestepper="SplitRungeKutta3") -> Case:
    """examples/ice_advected_by_anticyclone.jl scaled to N x N at dx = 4 km (N = 128 is the shipped case)."""
    L = N * 4000.0
    c = Case("anticyclone", N, N, H, H, ("Bounded", "Bounded"), L, L, dt=dt, substeps=substeps,
             advection_order=advection_order, timestepper=timestepper, u_bc_value=0.0, v_bc_value=0.0)
    rng = np.random.default_rng(seed)
    vo, va = 0.01, 30.0
    s = L / 512e3  # stretch the eddy with the domain so every cell stays active at large N
    Xc, Yc = c.nodes(LOC["h"])
    Xu, Yu = c.nodes(LOC["u"])
    Xv, Yv = c.nodes(LOC["v"])

    def wind(x, y):
        cen = 256e3 * s
        r = np.sqrt((x - cen) ** 2 + (y - cen) ** 2)
        sp = 1 / 100 * np.exp(-r / (100e3 * s))
        ua = -va * sp * (np.cos(np.deg2rad(72)) * (x - cen) + np.sin(np.deg2rad(72)) * (y - cen)) / (1000 * s)
        vaa = -va * sp * (-np.sin(np.deg2rad(72)) * (x - cen) + np.cos(np.deg2rad(72)) * (y - cen)) / (1000 * s)
        return ua, vaa

    uau, vau = wind(Xu, Yu)
    uav, vav = wind(Xv, Yv)
    tx = -uau * np.sqrt(uau ** 2 + vau ** 2) * 1.3 * 1.2e-3
    ty = -vav * np.sqrt(uav ** 2 + vav ** 2) * 1.3 * 1.2e-3
    ue = vo * (2 * Yu - L) / L
    ve = vo * (L - 2 * Xv) / L
    h = 0.3 + 0.005 * (np.sin(60 * Xc / 1000e3) + np.sin(30 * Yc / 1000e3)) + noise * rng.uniform(-1, 1, Xc.shape)
    a = np.ones_like(h)
    raw = dict(h=h, a=a, u=np.zeros_like(Xu), v=np.zeros_like(Xv), ue=ue, ve=ve, top_x=tx, top_y=ty)
    c.fields = {k: np.ascontiguousarray(vv, dtype=np.float64) for k, vv in raw.items()}
    return c


def latlon_case(N=96, H=4, seed=SEED, substeps=150, dt=600.0, advection_order=7, timestepper="SplitRungeKutta3",
                topology=("Bounded", "Bounded"), lon=(0.0, 60.0), lat=(20.0, 70.0), Ny=None) -> Case:
    """A basin on a LatitudeLongitudeGrid (the grid of test/test_rheology_energy_budget.jl:18-24 and of the coupled
    ClimaOcean set-ups): lambda in (0, 60), phi in (20, 70), closed or zonally periodic, rotating wind stress,
    sheared ocean current, variable ice cover.  The metrics vary by a factor ~2.7 across the rows."""
    Ny = N if Ny is None else Ny
    c = Case("latlon", N, Ny, H, H, tuple(topology), lon[1] - lon[0], lat[1] - lat[0], dt=dt, substeps=substeps,
             advection_order=advection_order, timestepper=timestepper, latlon=(lon, lat),
             u_bc_value=0.0 if topology[1] == "Bounded" else None, v_bc_value=0.0 if topology[0] == "Bounded" else None)
    rng = np.random.default_rng(seed)
    tp = 2 * np.pi
    X, Y = c.nodes(LOC["h"])
    Xu, Yu = c.nodes(LOC["u"])
    Xv, Yv = c.nodes(LOC["v"])
    fx = lambda x: x / c.Lx
    fy = lambda y: y / c.Ly
    h = 0.5 + 0.2 * np.sin(tp * fx(X)) * np.cos(tp * fy(Y)) + 1e-3 * rng.uniform(-1, 1, X.shape)
    a = np.clip(0.95 + 0.05 * np.cos(2 * tp * fx(X)) - 0.3 * (fy(Y) < 0.15), 0.0, 1.0)
    tx = 0.1 * np.cos(tp * fy(Yu))
    ty = 0.1 * np.sin(tp * fx(Xv))
    ue = 0.05 * np.sin(tp * fy(Yu))
    ve = 0.05 * np.sin(tp * fx(Xv))
    raw = dict(h=h, a=a, u=np.zeros_like(Xu), v=np.zeros_like(Xv), ue=ue, ve=ve, top_x=tx, top_y=ty)
    c.fields = {k: _wrap_periodic(c, np.ascontiguousarray(vv, dtype=np.float64), LOC[k]) for k, vv in raw.items()}
    return c


def curvilinear_case(N=72, Ny=56, H=4, seed=SEED, substeps=20, dt=600.0, advection_order=7, timestepper="SplitRungeKutta3",
                     topology=("Bounded", "Bounded")) -> Case:
    """An orthogonal curvilinear mesh given by two-dimensional metrics (the layout of Oceananigans' OrthogonalSphericalShellGrid,
    e.g. one panel of a rotated or stretched mesh): spacings vary smoothly by +-25 % in both directions; the areas are the
    products of the local spacings at each of the four locations.  Fields as in latlon_case."""
    c = Case("curvilinear", N, Ny, H, H, tuple(topology), float(N), float(Ny), dt=dt, substeps=substeps,
             advection_order=advection_order, timestepper=timestepper,
             u_bc_value=0.0 if topology[1] == "Bounded" else None, v_bc_value=0.0 if topology[0] == "Bounded" else None)
    tp = 2 * np.pi
    i = np.arange(1 - H, N + H + 2)[None, :].astype(np.float64)
    j = np.arange(1 - H, Ny + H + 2)[:, None].astype(np.float64)

    def spacing(x, y, base, phase):   # x, y: index coordinates of the location (periodic in the periodic directions)
        return base * (1.0 + 0.15 * np.sin(tp * x / N + phase) + 0.10 * np.cos(tp * y / Ny - phase))

    xc, xf, yc, yf = i - 0.5, i - 1.0, j - 0.5, j - 1.0
    M = {}
    for name, (x, y) in dict(cc=(xc, yc), fc=(xf, yc), cf=(xc, yf), ff=(xf, yf)).items():
        M["dx" + name] = spacing(x, y, 4000.0, 0.3) + 0 * (x + y)
        M["dy" + name] = spacing(x, y, 3000.0, 1.1) + 0 * (x + y)
        M["az" + name] = M["dx" + name] * M["dy" + name]
    # along a Periodic axis the halo metrics are exact images of the interior ones (as Oceananigans fills them): index j maps
    # to ((j - 1) mod N) + 1, so face N + 1 is face 1.  A partition relies on it: a rank recomputes its neighbour's rows.
    ii = np.arange(1 - H, N + H + 2)
    jj = np.arange(1 - H, Ny + H + 2)
    for k in list(M):
        if topology[0] == "Periodic":
            M[k] = M[k][:, ((ii - 1) % N) + H]
        if topology[1] == "Periodic":
            M[k] = M[k][((jj - 1) % Ny) + H, :]
    c.metric_arrays = {k: np.ascontiguousarray(v) for k, v in M.items()}
    rng = np.random.default_rng(seed)
    X, Y = c.nodes(LOC["h"])
    Xu, Yu = c.nodes(LOC["u"])
    Xv, Yv = c.nodes(LOC["v"])
    fx, fy = (lambda x: x / c.Lx), (lambda y: y / c.Ly)
    h = 0.5 + 0.2 * np.sin(tp * fx(X)) * np.cos(tp * fy(Y)) + 1e-3 * rng.uniform(-1, 1, X.shape)
    a = np.clip(0.95 + 0.05 * np.cos(2 * tp * fx(X)) - 0.3 * (fy(Y) < 0.15), 0.0, 1.0)
    raw = dict(h=h, a=a, u=np.zeros_like(Xu), v=np.zeros_like(Xv), ue=0.05 * np.sin(tp * fy(Yu)), ve=0.05 * np.sin(tp * fx(Xv)),
               top_x=0.1 * np.cos(tp * fy(Yu)), top_y=0.1 * np.sin(tp * fx(Xv)))
    c.fields = {k: _wrap_periodic(c, np.ascontiguousarray(vv, dtype=np.float64), LOC[k]) for k, vv in raw.items()}
    return c


def example_fold_maps(Nx, Ny, Hx, Hy, topo_x="Periodic", part_y=False):
    """Copy lists of a north fold through the centre row j = Ny (a U-point pivot, the older of the two fold variants of Oceananigans'
    TripolarGrid, written down from memory): an EXAMPLE for the tests of the mechanism -- a real host derives its lists from its own
    fill_halo_regions! (julia/ClimaSeaIceB200.jl: fold_maps), and the library takes whatever lists it is given.
        centres in x:  i' = Nx - i + 1        faces in x:  i' = Nx - i + 2   (wrapped periodically into 1..Nx)
        centres in y:  (i, Ny + k) <- (i', Ny - k), k = 1..Hy, and the pivot row itself, (i, Ny) <- (i', Ny) for i > Nx / 2 (and its periodic images)
        faces in y:    (i, Ny + k) <- (i', Ny - k + 1), k = 1..Hy + 1
    Every parent column is filled, x halos included."""
    maps = {}
    for lx in (0, 1):
        for ly in (0, 1):
            sx = Nx + 2 * Hx + (1 if (lx and topo_x == "Bounded") else 0)
            tg, sr = [], []
            idx = lambda i, j: (j - 1 + Hy) * sx + (i - 1 + Hx)
            for pi in range(sx):
                i = pi + 1 - Hx
                ip = (Nx - i + (2 if lx else 1) - 1) % Nx + 1
                for k in range(1, Hy + (2 if (ly and not part_y) else 1)):   # (a y-slab's Face-y parent has no extra row)
                    tg.append(idx(i, Ny + k)); sr.append(idx(ip, Ny - k + (1 if ly else 0)))
                iw = (i - 1) % Nx + 1   # the periodic image of a halo column
                if not ly and iw > Nx // 2 and ip != iw:
                    tg.append(idx(i, Ny)); sr.append(idx(ip, Ny))
            maps[(lx, ly)] = (np.asarray(tg, dtype=np.int32), np.asarray(sr, dtype=np.int32))
    return maps


def folded_case(Nx=48, Ny=40, H=5, seed=SEED, substeps=12, dt=600.0, timestepper="SplitRungeKutta3", mask=True) -> Case:
    """A tripolar-like case: two-dimensional metrics, zonally periodic, a wall in the south and a fold in the north (example copy
    lists, velocities with sign -1), optionally an immersed island touching the fold.  Exercises the mechanism the library offers
    for Oceananigans' TripolarGrid; not a statement about that grid's index convention."""
    c = curvilinear_case(Nx, Ny, H=H, seed=seed, substeps=substeps, dt=dt, timestepper=timestepper, topology=("Periodic", "Bounded"))
    c.name = "folded"
    c.topology = ("Periodic", "Folded")
    c.u_bc_value = 0.0
    c.fold = dict(maps=example_fold_maps(Nx, Ny, H, H), sign_velocity=-1.0, sign_external=1.0)
    rng = np.random.default_rng(seed + 11)
    c.fields["u"] = _wrap_periodic(c, 0.05 * rng.uniform(-1, 1, c.fields["u"].shape), LOC["u"])
    c.fields["v"] = _wrap_periodic(c, 0.05 * rng.uniform(-1, 1, c.fields["v"].shape), LOC["v"])
    if mask:
        X, Y = c.nodes(LOC["h"])
        land = ((X / c.Lx - 0.3) ** 2 + (Y / c.Ly - 0.95) ** 2) < 0.01
        m = land.astype(np.uint8)
        # the mask obeys the fold like any centre field (sign +1), and periodicity in x
        tg, sr = c.fold["maps"][(0, 0)]
        m = _wrap_periodic(c, m.astype(np.float64), LOC["h"])
        flat = m.reshape(-1)
        flat[tg] = flat[sr]
        c.mask = np.ascontiguousarray(m.astype(np.uint8))
    return c


def arctic_cap_case(Nx=192, Ny=48, H=7, seed=SEED, substeps=20, dt=600.0, timestepper="SplitRungeKutta3") -> Case:
    """BASELINE config 5 in miniature: a zonally periodic lat-lon cap (lambda in (0, 360), phi in (60, 88); the metrics
    shrink 14x towards the pole), HydrostaticSphericalCoriolis, EVP dynamics + WENO advection coupled to bare-ice slab thermodynamics with a
    latitude-dependent surface heat flux (freezing near the pole, melting at the ice edge)."""
    c = latlon_case(Nx, H=H, seed=seed, substeps=substeps, dt=dt, timestepper=timestepper, topology=("Periodic", "Bounded"),
                    lon=(0.0, 360.0), lat=(60.0, 88.0), Ny=Ny)
    c.name = "arctic-cap"
    c.rotation_rate = 7.292115e-5       # HydrostaticSphericalCoriolis
    rng = np.random.default_rng(seed + 5)
    X, Y = c.nodes(LOC["h"])
    fy = Y / c.Ly
    c.fields["Tu"] = -20.0 * fy - 2.0 + 0.1 * rng.uniform(-1, 1, X.shape)
    c.fields["Qtop"] = 150.0 * (fy - 0.35) + 20.0 * np.sin(2 * np.pi * X / c.Lx)      # W m^-2, > 0: heat leaves the ice
    c.thermo = dict(bottom_heat_flux=-4.0, ice_salinity=4.0)
    for k in ("Tu", "Qtop"):
        c.fields[k] = _wrap_periodic(c, np.ascontiguousarray(c.fields[k], dtype=np.float64), LOC[k])
    return c


def marginal_ice_case(N=64, H=5, seed=SEED, substeps=20, dt=120.0, variant="bottom_drag", snow=True,
                      timestepper="SplitRungeKutta3", topology=("Periodic", "Periodic")) -> Case:
    """A marginal ice zone: bands of compact ice, marginal ice (mass or concentration under the dynamical thresholds
    but above eps -- the cells that take the free-drift velocity) and open water, plus a snow layer.  Variants:
      "bottom_drag":  top = wind-stress arrays, bottom = SemiImplicitStress (ocean), StressBalanceFreeDrift  (TISB)
      "top_drag":     top = SemiImplicitStress with wind arrays, bottom = prescribed stress arrays, StressBalanceFreeDrift (BISB)
      "fields":       config-3 stresses, free_drift = (u = Field, v = Field)
      "both_drag":    SemiImplicitStress on both sides, free_drift = nothing
      "const_top_drag": top = SemiImplicitStress with constant wind, bottom = nothing
    """
    c = periodic_case(N, H=H, seed=seed, substeps=substeps, dt=dt, timestepper=timestepper) if topology == ("Periodic", "Periodic") else None
    if c is None:
        c = Case("marginal", N, N, H, H, tuple(topology), N * 4000.0, N * 4000.0, dt=dt, substeps=substeps, timestepper=timestepper,
                 u_bc_value=0.0 if topology[1] == "Bounded" else None, v_bc_value=0.0 if topology[0] == "Bounded" else None)
        base = periodic_case(N, H=H, seed=seed)
        for k, arr in base.fields.items():
            out = np.zeros(c.parent_shape(LOC[k]))
            out[:arr.shape[0], :arr.shape[1]] = arr
            c.fields[k] = out
    c.name = "marginal-" + variant
    rng = np.random.default_rng(seed + 1)
    X, Y = c.nodes(LOC["h"])
    a = c.fields["a"]
    h = c.fields["h"]
    band = (Y / c.Ly > 0.35) & (Y / c.Ly < 0.65)
    a[band] = 2e-4 + 6e-4 * rng.uniform(0, 1, a.shape)[band]            # under minimum_concentration = 1e-3
    thin = (X / c.Lx > 0.4) & (X / c.Lx < 0.6) & ~band
    h[thin] = 5e-4 + 5e-4 * rng.uniform(0, 1, h.shape)[thin]              # mass under minimum_mass = 1 kg m^-2
    tp = 2 * np.pi
    Xu, Yu = c.nodes(LOC["u"])
    Xv, Yv = c.nodes(LOC["v"])
    if variant == "bottom_drag":
        c.free_drift = "stress_balance"
        # a patch without wind exercises the tau == 0 branch of the closed form
        calm_u = (Xu / c.Lx < 0.2)
        calm_v = (Xv / c.Lx < 0.2)
        c.fields["top_x"] = np.where(calm_u, 0.0, c.fields["top_x"])
        c.fields["top_y"] = np.where(calm_v, 0.0, c.fields["top_y"])
    elif variant == "top_drag":
        c.free_drift = "stress_balance"
        c.top_kind = "semi_implicit"
        c.fields["top_x"] = 8.0 * np.sin(tp * Yu / c.Ly) + 2.0          # wind, m/s
        c.fields["top_y"] = 6.0 * np.cos(tp * Xv / c.Lx)
        c.bottom_kind = "stress"
        c.fields["ue"] = 0.02 * np.sin(tp * Yu / c.Ly)                    # prescribed ocean stress, N m^-2
        c.fields["ve"] = 0.02 * np.cos(tp * Xv / c.Lx)
    elif variant == "fields":
        c.free_drift = "fields"
        c.fields["fd_u"] = 0.03 * np.cos(tp * Yu / c.Ly)
        c.fields["fd_v"] = 0.02 * np.sin(tp * Xv / c.Lx)
    elif variant == "both_drag":
        c.top_kind = "semi_implicit"
        c.fields["top_x"] = 8.0 * np.sin(tp * Yu / c.Ly) + 2.0
        c.fields["top_y"] = 6.0 * np.cos(tp * Xv / c.Lx)
    elif variant == "const_top_drag":
        c.top_kind = "semi_implicit"
        c.fields.pop("top_x"); c.fields.pop("top_y")
        c.top_const = (5.0, -3.0)
        c.bottom_kind = "none"
        c.fields.pop("ue"); c.fields.pop("ve")
    else:
        raise ValueError(variant)
    if snow:
        c.fields["hs"] = np.where(a > 0, 0.1 + 0.05 * np.sin(tp * X / c.Lx) * np.cos(tp * Y / c.Ly), 0.0)
    for k in list(c.fields):
        c.fields[k] = _wrap_periodic(c, np.ascontiguousarray(c.fields[k], dtype=np.float64), LOC[k])
    return c


def slab_of(case: Case, rank: int, nranks: int, Hy: int) -> Case:
    """Rank-local y-slab (halo Hy) of a case whose x axis is anything and whose y axis is Periodic (halo rows =
    periodic images) or Bounded (halo rows = the global parent's rows where it has them, zeros beyond)."""
    assert case.Ny % nranks == 0
    ny = case.Ny // nranks
    c = dataclasses.replace(case, name=case.name + f"-slab{rank}", Ny=ny, Hy=Hy, Ly=case.Ly / nranks, fields={}, metric_arrays=None, mask=None)
    if case.fold is not None:
        # only the last slab holds the fold; its lists are the example's, over the slab's own rows and (deeper) halo
        c.fold = dict(case.fold, maps=example_fold_maps(case.Nx, ny, case.Hx, Hy, topo_x=case.topology[0], part_y=True)) if rank == nranks - 1 else None
    if case.latlon is not None:
        # the slab's rows of the GLOBAL grid's metrics (same expressions per global row index => same bits as on one rank);
        # local row jl = 1-Hy .. ny+Hy+1 is global row rank*ny + jl
        G = latitude_longitude_metrics(case.Nx, case.Ny, Hy, case.latlon[0], case.latlon[1])
        if case.rotation_rate is not None:
            G["f_ff"] = spherical_coriolis_f_ff(case.Ny, Hy, case.latlon[1], case.rotation_rate)
        c.metric_arrays = {k: np.ascontiguousarray(v[rank * ny:rank * ny + ny + 2 * Hy + 1]) for k, v in G.items()}
    j = np.arange(rank * ny - Hy, (rank + 1) * ny + Hy)          # 0-based global interior row of every slab row
    for k, arr in case.fields.items():
        if case.topology[1] == "Periodic":
            interior = arr[case.Hy:case.Hy + case.Ny, :]
            c.fields[k] = np.ascontiguousarray(interior[j % case.Ny, :])
        else:
            out = np.zeros((ny + 2 * Hy, arr.shape[1]))
            pj = j + case.Hy                                      # row in the global parent
            ok = (pj >= 0) & (pj < arr.shape[0])
            out[ok, :] = arr[pj[ok], :]
            c.fields[k] = out
    if case.latlon is None and case.metric_arrays is not None:
        # two-dimensional metrics: the slab's rows of the global arrays (local row j is global row rank * ny + j; wrapped
        # along a Periodic axis, clamped into the global parent beyond a wall, where nothing reads them)
        jl = np.arange(1 - Hy, ny + Hy + 2) + rank * ny
        rows = ((jl - 1) % case.Ny) + case.Hy if case.topology[1] == "Periodic" else np.clip(jl - 1 + case.Hy, 0, case.Ny + 2 * case.Hy)
        c.metric_arrays = {k: np.ascontiguousarray(v[rows, :]) for k, v in case.metric_arrays.items()}
    if case.mask is not None:   # the immersed mask (centres) travels with the rows; beyond a Bounded parent nothing is immersed
        if case.topology[1] == "Periodic":
            c.mask = np.ascontiguousarray(case.mask[case.Hy:case.Hy + case.Ny, :][j % case.Ny, :])
        else:
            c.mask = np.zeros((ny + 2 * Hy, case.mask.shape[1]), dtype=np.uint8)
            pj = j + case.Hy
            ok = (pj >= 0) & (pj < case.mask.shape[0])
            c.mask[ok, :] = case.mask[pj[ok], :]
            if c.fold is not None:   # the slab's halo is deeper than the global one: the mask beyond the fold is its image
                tg, sr = c.fold["maps"][(0, 0)]
                c.mask.reshape(-1)[tg] = c.mask.reshape(-1)[sr]
    return c


def block_of(case: Case, rank: int, Rx: int, Ry: int, Hx: int, Hy: int) -> Case:
    """Rank-local block (halos Hx, Hy) of a Rx x Ry partition, rank = ry * Rx + rx, of a regular-grid case: the y-slab of
    row ry, cut along x.  Halo columns are periodic images (Periodic x) or the global parent's columns (Bounded x)."""
    assert case.Nx % Rx == 0 and case.latlon is None
    rx, ry = rank % Rx, rank // Rx
    sl = slab_of(case, ry, Ry, Hy)
    if Rx == 1:
        return sl
    nx = case.Nx // Rx
    c = dataclasses.replace(sl, name=case.name + f"-block{rx}.{ry}", Nx=nx, Hx=Hx, Lx=case.Lx / Rx, fields={}, mask=None)
    i = np.arange(rx * nx - Hx, (rx + 1) * nx + Hx)               # 0-based global interior column of every block column
    if sl.mask is not None:
        if case.topology[0] == "Periodic":
            c.mask = np.ascontiguousarray(sl.mask[:, case.Hx:case.Hx + case.Nx][:, i % case.Nx])
        else:
            c.mask = np.zeros((sl.mask.shape[0], nx + 2 * Hx), dtype=np.uint8)
            pi = i + case.Hx
            ok = (pi >= 0) & (pi < sl.mask.shape[1])
            c.mask[:, ok] = sl.mask[:, pi[ok]]
    for k, arr in sl.fields.items():
        if case.topology[0] == "Periodic":
            interior = arr[:, case.Hx:case.Hx + case.Nx]
            c.fields[k] = np.ascontiguousarray(interior[:, i % case.Nx])
        else:
            out = np.zeros((arr.shape[0], nx + 2 * Hx))
            pi = i + case.Hx
            ok = (pi >= 0) & (pi < arr.shape[1])
            out[:, ok] = arr[:, pi[ok]]
            c.fields[k] = out
    return c


def slab_rows(case: Case, rank: int, nranks: int):
    """Rows of the global parent array that rank's interior covers."""
    ny = case.Ny // nranks
    return slice(case.Hy + rank * ny, case.Hy + (rank + 1) * ny)


def coastline_case(Ny=128, H=4, substeps=150, dt=300.0, seed=SEED, noise=1e-3) -> Case:
    """examples/ice_advected_on_coastline.jl:32-125 scaled to 2Ny x Ny (BASELINE config 4 at Ny = 4096):
    Periodic x Bounded, dx = dy = 2 km, triangular immersed coastline, uniform wind stress, ocean at rest
    (SemiImplicitStress with zero velocity), u = 0 on the walls, linear immersed drag C = 3e-3, h = aice = 1."""
    Nx = 2 * Ny
    c = Case("coastline", Nx, Ny, H, H, ("Periodic", "Bounded"), Nx * 2000.0, Ny * 2000.0, dt=dt, substeps=substeps,
             coriolis_f=None, u_bc_value=0.0, top_const=(-1.3 * 1.2e-3 * 10.0 ** 2, 0.0), ocean_const=(0.0, 0.0),
             immersed_drag=(3e-3, 3e-3))
    rng = np.random.default_rng(seed)
    Xc, Yc = c.nodes(LOC["h"])
    x = Xc - c.Lx / 2                                     # the example's x runs over (-Lx/2, Lx/2)
    land = (Yc <= c.Ly / 2) & (np.abs(x / c.Lx) * Nx + Yc / c.Ly * Ny <= 24 * (Ny / 128))
    c.mask = np.ascontiguousarray(land.astype(np.uint8))
    h = np.ones_like(Xc) + noise * rng.uniform(-1, 1, Xc.shape)
    raw = dict(h=h, a=np.ones_like(h), u=np.zeros(c.parent_shape(LOC["u"])), v=np.zeros(c.parent_shape(LOC["v"])))
    c.fields = {k: _wrap_periodic(c, np.ascontiguousarray(vv, dtype=np.float64), LOC[k]) for k, vv in raw.items()}
    return c

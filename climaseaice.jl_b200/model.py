"""Host-side mirror of the reference's user-facing interface for the hot path.

Julia is not available in this image, so the Julia structs the reference's users touch are
mirrored here with the same names, keyword arguments and defaults; every compute call goes
through the C ABI of libclimaseaice_b200.so (the same entry points a Julia `ccall` shim binds,
see INTEGRATION.md).  torch is used only to own device memory and streams.

    RectilinearGrid(size, x, y, halo, topology)              Oceananigans.Grids
    ElastoViscoPlasticRheology(...)     src/Rheologies/elasto_visco_plastic_rheology.jl:119-137
    SplitExplicitSolver(grid; substeps) src/SeaIceDynamics/split_explicit_momentum_equations.jl:18-46
    SemiImplicitStress(; ue, ve, rho_e, Cd)  src/SeaIceDynamics/sea_ice_external_stress.jl:84-130
    SeaIceMomentumEquation(grid; ...)   src/SeaIceDynamics/sea_ice_momentum_equations.jl:67-94
    SeaIceModel(grid; ...)              src/sea_ice_model.jl:140-297
    time_step!(model, dt)               src/sea_ice_rk_substep.jl:81-94 / src/sea_ice_fe_step.jl:13-34
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field as dc_field

import numpy as np
import torch

from . import _lib as L
from .metrics import METRIC_NAMES, latitude_longitude_metrics, spherical_coriolis_f_ff

Periodic, Bounded, Flat = "Periodic", "Bounded", "Flat"
Folded = "Folded"   # y axis of a tripolar grid: a wall in the south, the fold (Zipper boundary condition) in the north
Center, Face = 0, 1


class RectilinearGrid:
    """RectilinearGrid(size=(Nx, Ny), x=(x0, x1), y=(y0, y1), halo=(Hx, Hy), topology=(TX, TY, Flat))."""

    def __init__(self, size, x, y, halo=(3, 3), topology=(Periodic, Periodic, Flat), device=None, partitioned_y=False, partitioned_x=False):
        # partitioned_y / partitioned_x: this grid is a rank-local block; Face fields then carry no extra row / column on a
        # Bounded axis (the wall point lives in the block's halo), as with Oceananigans' Distributed grids
        self.partitioned_y = bool(partitioned_y)
        self.partitioned_x = bool(partitioned_x)
        self.Nx, self.Ny = int(size[0]), int(size[1])
        self.Hx, self.Hy = int(halo[0]), int(halo[1])
        self.x, self.y = (float(x[0]), float(x[1])), (float(y[0]), float(y[1]))
        self.topology = tuple(topology[:2])
        for k, t in enumerate(self.topology):
            if t not in (Periodic, Bounded) and not (k == 1 and t == Folded):
                raise ValueError(f"unsupported topology {t!r}")
        self.dx = (self.x[1] - self.x[0]) / self.Nx
        self.dy = (self.y[1] - self.y[0]) / self.Ny
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    @property
    def topo_codes(self):
        return tuple(L.PERIODIC if t == Periodic else (L.FOLDED if t == Folded else L.BOUNDED) for t in self.topology)

    def parent_shape(self, loc):
        """(sy, sx) of a field's parent: Face fields carry N+1 points along Bounded axes."""
        sx = self.Nx + 2 * self.Hx + (1 if (loc[0] == Face and self.topology[0] == Bounded and not self.partitioned_x) else 0)
        sy = self.Ny + 2 * self.Hy + (1 if (loc[1] == Face and self.topology[1] in (Bounded, Folded) and not self.partitioned_y) else 0)
        return sy, sx

    def nodes(self, loc, with_halos=True):
        """x(i), y(j) coordinate vectors of the parent array of a field at `loc`."""
        sy, sx = self.parent_shape(loc)
        i = np.arange(sx) - self.Hx + 1
        j = np.arange(sy) - self.Hy + 1
        xs = self.x[0] + ((i - 1) if loc[0] == Face else (i - 0.5)) * self.dx
        ys = self.y[0] + ((j - 1) if loc[1] == Face else (j - 0.5)) * self.dy
        return xs, ys


class LatitudeLongitudeGrid(RectilinearGrid):
    """LatitudeLongitudeGrid(size=(Nx, Ny), longitude=(l0, l1), latitude=(p0, p1), halo, topology, radius).

    Regularly spaced in degrees; the horizontal metrics depend on j only (Oceananigans' precomputed-metrics layout:
    dx = R cos(phi) dlambda at centre / face latitudes, dy = R dphi, Az = R^2 dlambda (sin phi_north - sin phi_south)).
    `metrics[name][j - 1 + Hy]` is the value at index j, for j = 1-Hy .. Ny+Hy+1 -- the arrays csi_config.metrics takes.
    `nodes()` returns degrees.  Only the general (per-kernel) solver formulation supports this grid.
    """

    def __init__(self, size, longitude, latitude, halo=(3, 3), topology=(Bounded, Bounded, Flat), radius=6371e3, device=None,
                 metrics=None, partitioned_y=False):
        super().__init__(size, longitude, latitude, halo=halo, topology=topology, device=device, partitioned_y=partitioned_y)
        self.radius = float(radius)
        if metrics is None:
            metrics = latitude_longitude_metrics(self.Nx, self.Ny, self.Hy, self.x, self.y, self.radius)
        self.metrics = {}
        for n in METRIC_NAMES:
            a = np.ascontiguousarray(metrics[n], dtype=np.float64).copy()
            if a.shape != (self.Ny + 2 * self.Hy + 1,):
                raise ValueError(f"metric {n}: expected {self.Ny + 2 * self.Hy + 1} entries, got {a.shape}")
            self.metrics[n] = a


class OrthogonalSphericalShellGrid(RectilinearGrid):
    """Orthogonal curvilinear grid given by its metrics (Oceananigans' OrthogonalSphericalShellGrid family: rotated,
    stretched or conformally mapped meshes; with topology (Periodic, Folded) and `SeaIceModel(..., fold=...)` a tripolar mesh): `metrics[name]` is a (Ny + 2Hy + 1) x (Nx + 2Hx + 1) array with the
    value at index (i, j) -- halos included -- stored at [j - 1 + Hy, i - 1 + Hx], for the twelve names of METRIC_NAMES
    (dx, dy, Az at the four horizontal locations), i.e. what Oceananigans' `Δxᶜᶜᵃ` ... `Azᶠᶠᵃ` return.  `nodes()` are index
    coordinates.  The fused tile kernel reads the metrics per node from planes in its own layout; a partition along x runs on
    the general (per-kernel) solver formulation."""

    def __init__(self, size, metrics, halo=(3, 3), topology=(Bounded, Bounded, Flat), device=None, partitioned_y=False):
        super().__init__(size, (0.0, float(size[0])), (0.0, float(size[1])), halo=halo, topology=topology, device=device,
                         partitioned_y=partitioned_y)
        self.metrics = {}
        shp = (self.Ny + 2 * self.Hy + 1, self.Nx + 2 * self.Hx + 1)
        for n in METRIC_NAMES:
            a = np.ascontiguousarray(metrics[n], dtype=np.float64).copy()
            if a.shape != shp:
                raise ValueError(f"metric {n}: expected shape {shp}, got {a.shape}")
            self.metrics[n] = a


class Field:
    """Oceananigans-layout field: a dense (sy, sx) float64 parent (i fastest) with halos, on the GPU."""

    def __init__(self, loc, grid, data=None):
        self.loc, self.grid = tuple(loc), grid
        shp = grid.parent_shape(self.loc)
        if data is None:
            self.parent = torch.zeros(shp, dtype=torch.float64, device=grid.device)
        else:
            t = torch.as_tensor(np.ascontiguousarray(data, dtype=np.float64))
            if tuple(t.shape) != shp:
                raise ValueError(f"parent shape {tuple(t.shape)} != {shp}")
            self.parent = t.to(grid.device).contiguous()

    @property
    def interior(self):
        g = self.grid
        sy, sx = self.parent.shape
        return self.parent[g.Hy:sy - g.Hy, g.Hx:sx - g.Hx]

    def set(self, value):
        """set!(field, value): a number, an (x, y) function, or an interior-shaped / parent-shaped array."""
        g = self.grid
        if callable(value):
            xs, ys = g.nodes(self.loc)
            X, Y = np.meshgrid(xs, ys)
            arr = np.asarray(value(X, Y), dtype=np.float64)
            self.interior.copy_(torch.as_tensor(arr[g.Hy:arr.shape[0] - g.Hy, g.Hx:arr.shape[1] - g.Hx]).to(self.parent.device))
        elif np.isscalar(value):
            self.interior.fill_(float(value))
        else:
            arr = torch.as_tensor(np.asarray(value, dtype=np.float64)).to(self.parent.device)
            if tuple(arr.shape) == tuple(self.parent.shape):
                self.parent.copy_(arr)
            else:
                self.interior.copy_(arr)
        return self

    def as_csi(self):
        a = L.csi_array()
        a.ptr = self.parent.data_ptr()
        a.ny_tot, a.nx_tot = self.parent.shape
        a.off_x, a.off_y = self.grid.Hx, self.grid.Hy
        return a

    def numpy(self):
        return self.parent.detach().cpu().numpy()


@dataclass
class ElastoViscoPlasticRheology:
    ice_compressive_strength: float = 27500.0
    ice_compaction_hardening: float = 20.0
    yield_curve_eccentricity: float = 2.0
    minimum_plastic_stress: float = 2e-9
    min_relaxation_parameter: float = 50.0
    max_relaxation_parameter: float = 300.0
    relaxation_strength: float = math.pi ** 2
    pressure_formulation: str = "ReplacementPressure"  # or "IceStrength"


@dataclass
class SplitExplicitSolver:
    substeps: int = 120  # SplitExplicitSolver(grid; substeps=120)


@dataclass
class SemiImplicitStress:
    ue: object = 0.0  # Field (f,c), or a number (ConstantField / ZeroField)
    ve: object = 0.0  # Field (c,f), or a number
    rho_e: float = 1026.0
    Cd: float = 5.5e-3


@dataclass
class StressBalanceFreeDrift:
    """StressBalanceFreeDrift(): tau_a = tau_o closed form (stress_balance_free_drift.jl:61-109).  As in the reference
    (sime.jl:80) it is repointed at the momentum equation's own top/bottom stresses, exactly one of which must be
    a SemiImplicitStress."""
    top_momentum_stress: object = None
    bottom_momentum_stress: object = None


@dataclass
class FPlane:
    f: float = 1e-4


@dataclass
class HydrostaticSphericalCoriolis:
    """HydrostaticSphericalCoriolis(; rotation_rate = Omega_Earth), EnstrophyConserving scheme, on a LatitudeLongitudeGrid:
    f at (Face, Face) = 2 Omega sin(phi_f)."""
    rotation_rate: float = 7.292115e-5
    f_ff_override: object = None   # explicit per-row values (a slab's rows of the global grid's f)

    def f_ff(self, grid):
        if self.f_ff_override is not None:
            return np.ascontiguousarray(self.f_ff_override, dtype=np.float64)
        return spherical_coriolis_f_ff(grid.Ny, grid.Hy, grid.y, self.rotation_rate)


@dataclass
class WENO:
    order: int = 5


@dataclass
class UpwindBiased:
    order: int = 1


@dataclass
class ValueBoundaryCondition:
    value: float = 0.0


# ---- thermodynamics (src/SeaIceThermodynamics) -------------------------------------------------------------------
@dataclass
class PhaseTransitions:
    """PhaseTransitions(; density=917, heat_capacity=2000, liquid_density=999.8, liquid_heat_capacity=4186,
    reference_latent_heat=334e3, reference_temperature=0, liquidus=LinearLiquidus(slope=0.054, T0=0))"""
    density: float = 917.0
    heat_capacity: float = 2000.0
    liquid_density: float = 999.8
    liquid_heat_capacity: float = 4186.0
    reference_latent_heat: float = 334e3
    reference_temperature: float = 0.0
    liquidus_slope: float = 0.054
    liquidus_freshwater_melting_temperature: float = 0.0


@dataclass
class MeltingConstrainedFluxBalance:
    """Top temperature from the flux balance Q_x(T) = Q_i(T), capped by the melting temperature (secant solve)."""
    tolerance: float = 1e-3      # RootSolvers SolutionTolerance default
    maxiters: int = 10000        # RootSolvers find_zero default


@dataclass
class PrescribedTemperature:
    temperature: object = 0.0    # number, or an array of parent shape (bottom only)


@dataclass
class IceWaterThermalEquilibrium:
    salinity: object = 0.0       # number or Field


@dataclass
class ConductiveFlux:
    conductivity: float = 2.0


@dataclass
class RadiativeEmission:
    emissivity: float = 1.0
    stefan_boltzmann_constant: float = 5.67e-8
    reference_temperature: float = 273.15


@dataclass
class LinearHeatFlux:
    """Q = coefficient * (T_top - temperature) [* aice]: the bulk sensible-heat FluxFunction of the reference's tests
    (test/test_energy_conservation.jl:8-13), offered as a built-in because closures cannot cross the C ABI."""
    coefficient: float = 0.0
    temperature: float = 0.0
    times_concentration: bool = True


class SlabThermodynamics:
    """SlabThermodynamics(grid; top_heat_boundary_condition=MeltingConstrainedFluxBalance(),
    bottom_heat_boundary_condition=IceWaterThermalEquilibrium(), internal_heat_flux=ConductiveFlux(conductivity=2))
    (slab_sea_ice_thermodynamics.jl:84-108)."""

    def __init__(self, grid, top_surface_temperature=None, top_heat_boundary_condition=None,
                 bottom_heat_boundary_condition=None, internal_heat_flux=None):
        self.top = top_heat_boundary_condition or MeltingConstrainedFluxBalance()
        self.bottom = bottom_heat_boundary_condition or IceWaterThermalEquilibrium()
        self.internal_heat_flux = internal_heat_flux or ConductiveFlux(2.0)
        self.top_surface_temperature = Field((Center, Center), grid)
        if top_surface_temperature is not None:
            self.top_surface_temperature.set(top_surface_temperature)
        elif isinstance(self.top, PrescribedTemperature):
            self.top_surface_temperature.set(self.top.temperature)


def sea_ice_slab_thermodynamics(grid, **kw):
    return SlabThermodynamics(grid, **kw)


def snow_slab_thermodynamics(grid, conductivity=0.31, **kw):
    return SlabThermodynamics(grid, internal_heat_flux=ConductiveFlux(conductivity), **kw)


class SeaIceMomentumEquation:
    def __init__(self, grid, coriolis=None, rheology=None, top_momentum_stress=None, bottom_momentum_stress=None,
                 free_drift=None, solver=None, minimum_concentration=1e-3, minimum_mass=1.0):
        # free_drift: nothing, (u=Field, v=Field), or StressBalanceFreeDrift()
        if free_drift is not None and not isinstance(free_drift, (dict, StressBalanceFreeDrift)):
            raise TypeError("free_drift must be nothing, dict(u=Field, v=Field) or StressBalanceFreeDrift()")
        if isinstance(free_drift, StressBalanceFreeDrift):
            ts, bs = isinstance(top_momentum_stress, SemiImplicitStress), isinstance(bottom_momentum_stress, SemiImplicitStress)
            if ts and bs:
                raise ValueError("`StressBalanceFreeDrift` supports a `SemiImplicitStress` only for the `top_momentum_stress` "
                                 "or the `bottom_momentum_stress`, not both")
            if not (ts or bs):
                raise ValueError("`StressBalanceFreeDrift` requires using a `SemiImplicitStress` for either the "
                                 "`top_momentum_stress` or the `bottom_momentum_stress`")
        self.free_drift = free_drift
        self.grid = grid
        self.coriolis = coriolis
        self.rheology = rheology or ElastoViscoPlasticRheology()
        self.solver = solver or SplitExplicitSolver(substeps=150)
        self.top = top_momentum_stress
        self.bottom = bottom_momentum_stress
        self.minimum_concentration = float(minimum_concentration)
        self.minimum_mass = float(minimum_mass)
        # Auxiliaries(r::ElastoViscoPlasticRheology, grid): evp.jl:140-173
        c, f = Center, Face
        self.auxiliaries = dict(
            s11=Field((c, c), grid), s22=Field((c, c), grid), s12=Field((f, f), grid),
            zeta_f=Field((f, f), grid), zeta_c=Field((c, c), grid), delta=Field((c, c), grid),
            alpha=Field((c, c), grid), un=Field((f, c), grid), vn=Field((c, f), grid), P=Field((c, c), grid))
        self.auxiliaries["alpha"].parent.fill_(self.rheology.max_relaxation_parameter)  # evp.jl:161


class SeaIceModel:
    """SeaIceModel(grid; dynamics, advection, timestepper=:SplitRungeKutta3, boundary_conditions, ice_density=900)."""

    def __init__(self, grid, dynamics=None, advection=None, timestepper="SplitRungeKutta3", boundary_conditions=None,
                 ice_density=900.0, ice_thermodynamics=None, solver_impl="auto", immersed_mask=None,
                 partition=None, immersed_drag=(0.0, 0.0), snow_thickness=False, snow_thermodynamics=None,
                 top_heat_flux=None, bottom_heat_flux=0.0, snowfall=0.0, snow_density=330.0,
                 ice_consolidation_thickness=0.05, ice_salinity=0.0, phase_transitions=None, fold=None):
        if dynamics is None and ice_thermodynamics is None:
            raise ValueError("pass a SeaIceMomentumEquation as `dynamics` and/or SlabThermodynamics as `ice_thermodynamics`")
        if snow_thermodynamics is not None and ice_thermodynamics is None:
            raise ValueError("snow_thermodynamics needs ice_thermodynamics")
        # a thermodynamics-only model keeps an (unused) momentum equation so that the field set stays uniform
        self._no_dynamics = dynamics is None
        if dynamics is None:
            dynamics = SeaIceMomentumEquation(grid)
        snow_thickness = bool(snow_thickness) or snow_thermodynamics is not None  # sea_ice_model.jl:203
        self.ice_thermodynamics, self.snow_thermodynamics = ice_thermodynamics, snow_thermodynamics
        self.phase_transitions = phase_transitions or PhaseTransitions()
        self.grid, self.dynamics = grid, dynamics
        self.advection = advection
        self.timestepper = timestepper
        if timestepper not in ("SplitRungeKutta3", "ForwardEuler"):
            raise ValueError(timestepper)
        c, f = Center, Face
        self.velocities = dict(u=Field((f, c), grid), v=Field((c, f), grid))
        self.ice_thickness = Field((c, c), grid)
        self.ice_concentration = Field((c, c), grid)
        self.sea_ice_density = float(ice_density)
        self.Gn = dict(h=Field((c, c), grid), a=Field((c, c), grid))
        rk = timestepper == "SplitRungeKutta3"
        self.Psi_m = dict(h=Field((c, c), grid), a=Field((c, c), grid), u=Field((f, c), grid), v=Field((c, f), grid)) if rk else None
        # prognostic snow thickness hs (allocated by the reference when snow_thermodynamics is given, sea_ice_model.jl:203):
        # here only its advection with the ice (tracer_tendency:47-52, fe.jl:84-94)
        self.snow_thickness = Field((c, c), grid) if snow_thickness else None
        if snow_thickness:
            self.Gn["hs"] = Field((c, c), grid)
            if rk:
                self.Psi_m["hs"] = Field((c, c), grid)
        self.iteration = 0
        self.time = 0.0
        bcs = boundary_conditions or {}
        self._u_bc = bcs.get("u", {})
        self._v_bc = bcs.get("v", {})
        self.partition = partition  # (rank, nranks, exchange_every[, Rx]): y-slabs, or Rx x (nranks / Rx) blocks with rank = ry * Rx + rx
        # ImmersedBoundaryCondition with the discrete-form flux -C*u (u: south/north) and -C*v (v: west/east), as in
        # examples/ice_advected_on_coastline.jl:91-98
        self.immersed_drag = (float(immersed_drag[0]), float(immersed_drag[1]))
        self._mask = None if immersed_mask is None else np.ascontiguousarray(immersed_mask, dtype=np.uint8)
        # Folded y axis (tripolar grid): fold = dict(maps={(lx, ly): (target, source)}, sign_velocity=-1.0, sign_external=1.0), the
        # copy lists of the north fold in linear parent indices -- what the host's fill_halo_regions! does on its grid
        self._fold = fold
        holds_fold = partition is None or partition[0] == partition[1] - 1 or (len(partition) > 3 and partition[3] > 1)   # y-slabs: the last one
        if fold is not None and grid.topology[1] != Folded:
            raise ValueError("only a Folded y axis takes `fold`")
        if grid.topology[1] == Folded and holds_fold and fold is None:
            raise ValueError("a Folded y axis needs `fold` (on a partition: on the last y-slab)")
        self._handle = C.c_void_p()
        self._solver_impl = dict(auto=L.SOLVER_AUTO, unfused=L.SOLVER_UNFUSED, fused=L.SOLVER_FUSED)[solver_impl]
        cfg = self._config()
        L.check(L.lib().csi_create(C.byref(cfg), C.byref(self._handle)))
        self._cfg = cfg
        self._thermo = None
        if ice_thermodynamics is not None:
            self._setup_thermodynamics(top_heat_flux, bottom_heat_flux, snowfall, snow_density, ice_consolidation_thickness, ice_salinity)

    # -- thermodynamics -> csi_thermo_config / csi_thermo_fields ----------------------------------
    def _setup_thermodynamics(self, top_heat_flux, bottom_heat_flux, snowfall, snow_density, hc, salinity):
        g, it, st, pt = self.grid, self.ice_thermodynamics, self.snow_thermodynamics, self.phase_transitions
        cc = (Center, Center)
        tc = L.csi_thermo_config()
        for n in ("density", "heat_capacity", "liquid_density", "liquid_heat_capacity", "reference_latent_heat",
                  "reference_temperature", "liquidus_slope", "liquidus_freshwater_melting_temperature"):
            setattr(tc, n, getattr(pt, n))
        kind = lambda bc: L.TOP_PRESCRIBED if isinstance(bc, PrescribedTemperature) else L.TOP_FLUX_BALANCE
        tc.top_heat_bc = kind(it.top)
        tc.snow_top_heat_bc = kind(st.top) if st is not None else L.TOP_FLUX_BALANCE
        solver = next((bc for bc in ((st.top if st is not None else None), it.top) if isinstance(bc, MeltingConstrainedFluxBalance)),
                      MeltingConstrainedFluxBalance())
        tc.secant_tolerance, tc.secant_maxiters = solver.tolerance, solver.maxiters
        tc.layered = 1 if st is not None else 0
        tc.ice_conductivity = it.internal_heat_flux.conductivity
        tc.snow_conductivity = st.internal_heat_flux.conductivity if st is not None else 0.31
        arrays = {}

        def scalar_or_field(value, name, attr):
            if isinstance(value, Field):
                arrays[name] = value
            elif np.ndim(value) > 0:
                arrays[name] = Field(cc, g, value)
            else:
                setattr(tc, attr, float(value))
        if isinstance(it.bottom, PrescribedTemperature):
            tc.bottom_heat_bc = L.BOTTOM_PRESCRIBED
            scalar_or_field(it.bottom.temperature, "Tb", "bottom_temperature")
        else:
            tc.bottom_heat_bc = L.BOTTOM_EQUILIBRIUM
            scalar_or_field(it.bottom.salinity, "Sb", "bottom_salinity")
        # external top flux (sea_ice_model.jl:242-256): default 0, or the conductive flux itself for a bare-ice
        # PrescribedTemperature top (no surface imbalance)
        if top_heat_flux is None:
            top_heat_flux = ("conductive",) if (st is None and isinstance(it.top, PrescribedTemperature)) else 0.0
        terms = top_heat_flux if isinstance(top_heat_flux, (tuple, list)) else (top_heat_flux,)
        if not 1 <= len(terms) <= 2:
            raise NotImplementedError("top_heat_flux: one flux or a tuple of two")
        tc.n_top_terms = len(terms)
        for k, t in enumerate(terms):
            if isinstance(t, RadiativeEmission):
                tc.top_term_kind[k] = L.FLUX_RADIATIVE_EMISSION
                tc.emissivity, tc.stefan_boltzmann_constant, tc.emission_reference_temperature = \
                    t.emissivity, t.stefan_boltzmann_constant, t.reference_temperature
            elif isinstance(t, LinearHeatFlux):
                tc.top_term_kind[k] = L.FLUX_LINEAR
                tc.linear_coefficient, tc.linear_temperature = t.coefficient, t.temperature
                tc.linear_times_concentration = 1 if t.times_concentration else 0
            elif isinstance(t, str) and t == "conductive":
                tc.top_term_kind[k] = L.FLUX_CONDUCTIVE
            elif isinstance(t, Field) or np.ndim(t) > 0:
                tc.top_term_kind[k] = L.FLUX_ARRAY
                arrays["Qtop"] = t if isinstance(t, Field) else Field(cc, g, t)
            elif callable(t):
                raise NotImplementedError("FluxFunction closures cannot cross the C ABI: use a number, an array, "
                                          "RadiativeEmission or a tuple of two of them")
            else:
                tc.top_term_kind[k] = L.FLUX_CONST
                tc.top_flux_const = float(t)
        scalar_or_field(bottom_heat_flux, "Qbot", "bottom_flux_const")
        scalar_or_field(snowfall, "snowfall", "snowfall")
        scalar_or_field(snow_density, "rho_s", "snow_density")
        scalar_or_field(hc, "hc", "ice_consolidation_thickness")
        scalar_or_field(salinity, "S", "ice_salinity")
        self.mass_fluxes = dict(ice=Field(cc, g), snow=Field(cc, g), intercepted_snowfall=Field(cc, g))
        arrays.update(h=self.ice_thickness, a=self.ice_concentration, Tu=it.top_surface_temperature,
                      mf_ice=self.mass_fluxes["ice"], mf_snow=self.mass_fluxes["snow"], mf_snowfall=self.mass_fluxes["intercepted_snowfall"])
        if st is not None:
            arrays.update(hs=self.snow_thickness, Tus=st.top_surface_temperature)
        tf = L.csi_thermo_fields()
        for n, fld in arrays.items():
            setattr(tf, n, fld.as_csi())
        self._thermo = (tc, tf, arrays)
        L.check(L.lib().csi_attach_thermodynamics(self._handle, C.byref(tc), C.byref(tf)), self._handle)

    def thermodynamic_time_step(self, dt):
        """thermodynamic_time_step!(model, model.ice_thermodynamics, model.snow_thermodynamics, dt)"""
        tc, tf, _ = self._thermo
        L.check(L.lib().csi_thermodynamic_time_step(self._handle, C.byref(tc), C.byref(tf), float(dt), self._stream()), self._handle)

    # -- configuration -> csi_config ---------------------------------------------------------
    def _config(self):
        g, d = self.grid, self.dynamics
        r = d.rheology
        cfg = L.csi_config()
        cfg.abi_version = L.ABI_VERSION
        cfg.device = g.device.index or 0
        cfg.Nx, cfg.Ny, cfg.Hx, cfg.Hy = g.Nx, g.Ny, g.Hx, g.Hy
        cfg.topo_x, cfg.topo_y = g.topo_codes
        cfg.dx, cfg.dy = g.dx, g.dy
        if isinstance(g, (LatitudeLongitudeGrid, OrthogonalSphericalShellGrid)):
            cfg.metric_kind = L.METRIC_J if isinstance(g, LatitudeLongitudeGrid) else L.METRIC_IJ
            for k, n in enumerate(METRIC_NAMES):
                cfg.metrics[k] = g.metrics[n].ctypes.data_as(C.POINTER(C.c_double))
        cfg.immersed_mask = self._mask.ctypes.data if self._mask is not None else None
        cfg.ice_compressive_strength = r.ice_compressive_strength
        cfg.ice_compaction_hardening = r.ice_compaction_hardening
        cfg.yield_curve_eccentricity = r.yield_curve_eccentricity
        cfg.minimum_plastic_stress = r.minimum_plastic_stress
        cfg.min_relaxation_parameter = r.min_relaxation_parameter
        cfg.max_relaxation_parameter = r.max_relaxation_parameter
        cfg.relaxation_strength = r.relaxation_strength
        cfg.pressure_formulation = L.ICE_STRENGTH if r.pressure_formulation == "IceStrength" else L.REPLACEMENT_PRESSURE
        cfg.substeps = d.solver.substeps
        cfg.minimum_mass, cfg.minimum_concentration = d.minimum_mass, d.minimum_concentration
        cfg.ice_density = self.sea_ice_density
        if isinstance(d.coriolis, HydrostaticSphericalCoriolis):
            if not isinstance(g, LatitudeLongitudeGrid):
                raise ValueError("HydrostaticSphericalCoriolis needs a LatitudeLongitudeGrid")
            cfg.coriolis_kind = L.CORIOLIS_SPHERICAL
            self._f_ff = d.coriolis.f_ff(g)
            if self._f_ff.shape != (g.Ny + 2 * g.Hy + 1,):
                raise ValueError("coriolis f_ff: expected Ny + 2 Hy + 1 entries")
            cfg.coriolis_f_ff = self._f_ff.ctypes.data_as(C.POINTER(C.c_double))
        else:
            cfg.coriolis_kind = L.CORIOLIS_FPLANE if d.coriolis is not None else L.CORIOLIS_NONE
            cfg.coriolis_f = d.coriolis.f if d.coriolis is not None else 0.0
        # either stress: nothing | (u=Number, v=Number) | (u=Field, v=Field) | SemiImplicitStress   (ext.jl:8-40,84-146)
        def kind_of(st):
            if st is None:
                return L.STRESS_NONE, (0.0, 0.0)
            if isinstance(st, SemiImplicitStress):
                return L.STRESS_SEMI_IMPLICIT, ((0.0, 0.0) if isinstance(st.ue, Field) else (float(st.ue), float(st.ve)))
            if isinstance(st, dict) and isinstance(st["u"], Field) and isinstance(st["v"], Field):
                return L.STRESS_FIELD, (0.0, 0.0)
            if isinstance(st, dict) and not isinstance(st["u"], Field) and not isinstance(st["v"], Field):
                return L.STRESS_CONST, (float(st["u"]), float(st["v"]))
            raise NotImplementedError("a momentum stress must be nothing, (u=, v=) numbers, (u=, v=) Fields or a SemiImplicitStress")
        cfg.top_stress_kind, (cfg.top_tau_x, cfg.top_tau_y) = kind_of(d.top)
        cfg.bottom_stress_kind, (cfg.ue_const, cfg.ve_const) = kind_of(d.bottom)
        if isinstance(d.top, SemiImplicitStress):
            cfg.top_rho_e, cfg.top_Cd = d.top.rho_e, d.top.Cd
        if isinstance(d.bottom, SemiImplicitStress):
            cfg.rho_e, cfg.Cd = d.bottom.rho_e, d.bottom.Cd
        fd = d.free_drift
        cfg.free_drift_kind = L.FD_NONE if fd is None else (L.FD_FIELDS if isinstance(fd, dict) else L.FD_STRESS_BALANCE)
        for side in ("south", "north"):
            if side in self._u_bc:
                cfg.u_south_north_bc, cfg.u_south_north_value = L.BC_VALUE, float(self._u_bc[side].value)
        for side in ("west", "east"):
            if side in self._v_bc:
                cfg.v_west_east_bc, cfg.v_west_east_value = L.BC_VALUE, float(self._v_bc[side].value)
        cfg.advection_order = 0 if self.advection is None else int(self.advection.order)
        cfg.timestepper = L.RK3 if self.timestepper == "SplitRungeKutta3" else L.FE
        cfg.solver_impl = self._solver_impl
        cfg.immersed_drag_u, cfg.immersed_drag_v = self.immersed_drag
        if self._fold is not None:
            self._fold_keep = []
            for (lx, ly), (tg, sr) in self._fold["maps"].items():
                k = lx + 2 * ly
                tg = np.ascontiguousarray(tg, dtype=np.int32); sr = np.ascontiguousarray(sr, dtype=np.int32)
                self._fold_keep += [tg, sr]
                cfg.fold_target[k] = tg.ctypes.data_as(C.POINTER(C.c_int32))
                cfg.fold_source[k] = sr.ctypes.data_as(C.POINTER(C.c_int32))
                cfg.fold_count[k] = tg.size
            cfg.fold_sign_velocity = float(self._fold.get("sign_velocity", -1.0))
            cfg.fold_sign_external = float(self._fold.get("sign_external", 1.0))
        if self.partition:
            cfg.rank, cfg.nranks, cfg.exchange_every = self.partition[:3]
            cfg.partition_x = int(self.partition[3]) if len(self.partition) > 3 else 0
        else:
            cfg.rank, cfg.nranks, cfg.exchange_every = 0, 1, 0
        return cfg

    # -- fields -> csi_fields ----------------------------------------------------------------
    def all_fields(self):
        d = self.dynamics
        out = dict(u=self.velocities["u"], v=self.velocities["v"], h=self.ice_thickness, a=self.ice_concentration,
                   Gh=self.Gn["h"], Ga=self.Gn["a"])
        out.update(d.auxiliaries)
        if self.Psi_m:
            out.update(hm=self.Psi_m["h"], am=self.Psi_m["a"], um=self.Psi_m["u"], vm=self.Psi_m["v"])
        for st, (nx, ny) in ((d.top, ("top_x", "top_y")), (d.bottom, ("ue", "ve"))):
            if isinstance(st, dict) and isinstance(st["u"], Field):
                out.update({nx: st["u"], ny: st["v"]})
            if isinstance(st, SemiImplicitStress) and isinstance(st.ue, Field):
                out.update({nx: st.ue, ny: st.ve})
        if isinstance(d.free_drift, dict):
            out.update(fd_u=d.free_drift["u"], fd_v=d.free_drift["v"])
        if self.snow_thickness is not None:
            out.update(hs=self.snow_thickness, Ghs=self.Gn["hs"])
            if self.Psi_m:
                out.update(hsm=self.Psi_m["hs"])
        return out

    def csi_fields(self):
        f = L.csi_fields()
        for n, fld in self.all_fields().items():
            setattr(f, n, fld.as_csi())
        return f

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.grid.device).cuda_stream)

    # -- the reference's methods -----------------------------------------------------------------
    def set(self, **kw):
        """set!(model, h=..., ℵ=..., u=..., v=...) (use `a` for ℵ)."""
        targets = dict(h=self.ice_thickness, a=self.ice_concentration, u=self.velocities["u"], v=self.velocities["v"])
        if "hs" in kw:
            if self.snow_thickness is None:  # sea_ice_model.jl:307-311
                raise ValueError("Cannot set snow thickness `hs` on a SeaIceModel without snow (model.snow_thickness is nothing).")
            targets["hs"] = self.snow_thickness
        for k, v in kw.items():
            targets["a" if k in ("ℵ", "aice") else k].set(v)
        return self

    def time_step_momentum(self, dt, substeps=None):
        """time_step_momentum!(model, dynamics, dt)"""
        f = self.csi_fields()
        n = self.dynamics.solver.substeps if substeps is None else int(substeps)
        L.check(L.lib().csi_evp_substeps(self._handle, C.byref(f), float(dt), n, self._stream()), self._handle)

    def compute_tracer_tendencies(self):
        f = self.csi_fields()
        L.check(L.lib().csi_compute_tracer_tendencies(self._handle, C.byref(f), self._stream()), self._handle)

    def dynamic_time_step(self, dt):
        f = self.csi_fields()
        L.check(L.lib().csi_dynamic_time_step(self._handle, C.byref(f), float(dt), self._stream()), self._handle)

    def cache_current_fields(self):
        f = self.csi_fields()
        L.check(L.lib().csi_cache_current_fields(self._handle, C.byref(f), self._stream()), self._handle)

    def update_state(self):
        f = self.csi_fields()
        L.check(L.lib().csi_update_state(self._handle, C.byref(f), self._stream()), self._handle)

    def time_step(self, dt):
        """time_step!(model, dt)"""
        if self._no_dynamics:
            return self._time_step_without_dynamics(dt)
        f = self.csi_fields()
        L.check(L.lib().csi_time_step(self._handle, C.byref(f), float(dt), 1 if self.iteration == 0 else 0, self._stream()),
                self._handle)
        self.iteration += 1
        self.time += dt

    def _time_step_without_dynamics(self, dt):
        """time_step! with dynamics = nothing (fe.jl:13-34, rk.jl:81-94): advection with the model's (zero) velocities,
        no momentum step, then the thermodynamic step; driven from the host through the per-stage entry points."""
        if self.iteration == 0:
            self.update_state()
        if self.timestepper == "ForwardEuler":
            self.compute_tracer_tendencies()
            self.dynamic_time_step(dt)
            self.thermodynamic_time_step(dt)
            self.update_state()
        else:
            self.cache_current_fields()
            for beta in (3, 2, 1):
                dtau = dt / beta
                self.compute_tracer_tendencies()
                self.dynamic_time_step(dtau)
                self.thermodynamic_time_step(dtau)
                self.update_state()
        self.iteration += 1
        self.time += dt

    def cell_advection_timescale(self):
        f = self.csi_fields()
        out = C.c_double()
        L.check(L.lib().csi_cell_advection_timescale(self._handle, C.byref(f), C.byref(out), self._stream()), self._handle)
        return out.value

    def diagnostics(self):
        f = self.csi_fields()
        out = (C.c_double * 5)()
        L.check(L.lib().csi_diagnostics(self._handle, C.byref(f), out, self._stream()), self._handle)
        return dict(zip(("sum_h_Az", "sum_a_Az", "sum_ha_Az", "max_abs_u", "max_abs_v"), out))

    def comm_init(self, unique_id: bytes):
        rank, nranks = self.partition[:2]
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        L.check(L.lib().csi_comm_init(self._handle, buf, rank, nranks), self._handle)

    def fused_stats(self):
        """(inputs_failed_validation, tile passes redone with the IEEE operators, tiles per substep) of the last momentum solve."""
        out = (C.c_int64 * 3)()
        L.check(L.lib().csi_fused_stats(self._handle, out), self._handle)
        return tuple(out)

    @property
    def launch_count(self):
        return L.lib().csi_launch_count(self._handle)

    @property
    def last_elapsed_ms(self):
        return L.lib().csi_last_elapsed_ms(self._handle)

    def close(self):
        if self._handle:
            L.lib().csi_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def time_step_b(model, dt):
    """time_step!(model, dt) -- free-function spelling of the reference's API."""
    model.time_step(dt)


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    L.check(L.lib().csi_nccl_unique_id(buf))
    return bytes(buf)

"""ctypes binding of the CPU ORACLE (oracle/csi_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
PARITY UNPINNED -- see oracle/csi_oracle.h.

Arrays are numpy float64, C-order, shape (sy, sx) -- i.e. Oceananigans' column-major parent
(i fastest) seen from numpy as arr[pj, pi].  Element (i, j) (1-based Julia indices) of a field
with halo offsets (ox, oy) is arr[j-1+oy, i-1+ox].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libcsi_oracle.so"

PERIODIC, BOUNDED, FOLDED = 0, 1, 2
STRESS_NONE, STRESS_CONST, STRESS_FIELD, STRESS_SEMI_IMPLICIT = 0, 1, 2, 3
RK3, FE = 0, 1


def build(force: bool = False) -> Path:
    """Compile the oracle with gcc (oracle/Makefile)."""
    srcs = [_HERE / "csi_oracle.c", _HERE / "csi_oracle_weno.c", _HERE / "csi_oracle_thermo.c", _HERE / "csi_oracle.h"]
    if force or not _LIB_PATH.exists() or any(s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


class Field(C.Structure):
    _fields_ = [("p", C.POINTER(C.c_double)), ("sx", C.c_int32), ("sy", C.c_int32), ("ox", C.c_int32), ("oy", C.c_int32)]


class Grid(C.Structure):
    _fields_ = [
        ("Nx", C.c_int32), ("Ny", C.c_int32), ("Hx", C.c_int32), ("Hy", C.c_int32),
        ("topo_x", C.c_int32), ("topo_y", C.c_int32), ("metric_kind", C.c_int32), ("metW", C.c_int32),
        ("dx", C.c_double), ("dy", C.c_double),
    ] + [(n, C.POINTER(C.c_double)) for n in
         ("dxcc", "dxfc", "dxcf", "dxff", "dycc", "dyfc", "dycf", "dyff", "azcc", "azfc", "azcf", "azff")] + [
        ("mask", C.POINTER(C.c_uint8)),
        ("fold_target", C.POINTER(C.c_int32) * 4), ("fold_source", C.POINTER(C.c_int32) * 4), ("fold_count", C.c_int32 * 4),
        ("fold_sign_velocity", C.c_double), ("fold_sign_external", C.c_double)]


class Params(C.Structure):
    _fields_ = [
        ("Pstar", C.c_double), ("C", C.c_double), ("e", C.c_double), ("Dmin", C.c_double),
        ("alpha_min", C.c_double), ("alpha_max", C.c_double), ("c_alpha", C.c_double),
        ("pressure_formulation", C.c_int32), ("substeps", C.c_int32),
        ("min_mass", C.c_double), ("min_conc", C.c_double), ("rho_ice", C.c_double),
        ("coriolis_kind", C.c_int32), ("pad0_", C.c_int32), ("f", C.c_double),
        ("top_kind", C.c_int32), ("pad1_", C.c_int32), ("top_tx", C.c_double), ("top_ty", C.c_double),
        ("top_x", Field), ("top_y", Field),
        ("bot_kind", C.c_int32), ("pad2_", C.c_int32),
        ("rho_e", C.c_double), ("Cd", C.c_double), ("ue_c", C.c_double), ("ve_c", C.c_double),
        ("ue", Field), ("ve", Field),
        ("u_sn_bc", C.c_int32), ("v_we_bc", C.c_int32), ("u_sn_val", C.c_double), ("v_we_val", C.c_double),
        ("advection_order", C.c_int32), ("timestepper", C.c_int32),
        ("imm_drag_u", C.c_double), ("imm_drag_v", C.c_double),
        ("free_drift_kind", C.c_int32), ("pad3_", C.c_int32), ("fd_u", Field), ("fd_v", Field),
        ("top_rho", C.c_double), ("top_Cd", C.c_double), ("f_ff", C.POINTER(C.c_double)),
    ]


_STATE_NAMES = ("u", "v", "h", "a", "s11", "s22", "s12", "zf", "zc", "delta", "alpha", "un", "vn", "P",
                "Gh", "Ga", "hm", "am", "um", "vm", "hs", "Ghs", "hsm")
_SNOW_NAMES = ("hs", "Ghs", "hsm")
FD_NONE, FD_FIELDS, FD_STRESS_BALANCE = 0, 1, 2


class State(C.Structure):
    _fields_ = [(n, Field) for n in _STATE_NAMES]


# location of every state field: (face_x, face_y)
LOC = dict(u=(1, 0), v=(0, 1), h=(0, 0), a=(0, 0), s11=(0, 0), s22=(0, 0), s12=(1, 1), zf=(1, 1), zc=(0, 0),
           delta=(0, 0), alpha=(0, 0), un=(1, 0), vn=(0, 1), P=(0, 0), Gh=(0, 0), Ga=(0, 0),
           hm=(0, 0), am=(0, 0), um=(1, 0), vm=(0, 1), top_x=(1, 0), top_y=(0, 1), ue=(1, 0), ve=(0, 1),
           hs=(0, 0), Ghs=(0, 0), hsm=(0, 0), fd_u=(1, 0), fd_v=(0, 1))

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.csio_exp.restype = C.c_double
        L.csio_exp.argtypes = [C.c_double]
        L.csio_pow4.restype = C.c_double
        L.csio_pow4.argtypes = [C.c_double]
        L.csio_cell_advection_timescale.restype = C.c_double
        L.csio_reconstruct_x.restype = C.c_double
        L.csio_reconstruct_y.restype = C.c_double
        L.csio_reconstruct_x.argtypes = [C.POINTER(Grid), C.c_int, C.c_int, C.POINTER(Field), C.c_int, C.c_int]
        L.csio_reconstruct_y.argtypes = L.csio_reconstruct_x.argtypes
        _lib = L
    return _lib


def _as_field(arr, ox, oy):
    f = Field()
    if arr is None:
        return f
    assert arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"], "float64 C-contiguous (sy, sx) array expected"
    f.p = arr.ctypes.data_as(C.POINTER(C.c_double))
    f.sy, f.sx = arr.shape
    f.ox, f.oy = ox, oy
    return f


def parent_shape(Nx, Ny, Hx, Hy, topo, loc):
    """Oceananigans parent extents: Face fields carry N+1 points along Bounded axes."""
    sx = Nx + 2 * Hx + (1 if (loc[0] and topo[0] == BOUNDED) else 0)
    sy = Ny + 2 * Hy + (1 if (loc[1] and topo[1] in (BOUNDED, FOLDED)) else 0)
    return sy, sx


DEFAULT_PARAMS = dict(
    Pstar=27500.0, C=20.0, e=2.0, Dmin=2e-9, alpha_min=50.0, alpha_max=300.0, c_alpha=float(np.pi) ** 2,
    pressure_formulation=0, substeps=150, min_mass=1.0, min_conc=1e-3, rho_ice=900.0,
    coriolis_kind=0, f=0.0, top_kind=STRESS_NONE, top_tx=0.0, top_ty=0.0,
    bot_kind=STRESS_NONE, rho_e=1026.0, Cd=5.5e-3, ue_c=0.0, ve_c=0.0,
    u_sn_bc=0, v_we_bc=0, u_sn_val=0.0, v_we_val=0.0, advection_order=7, timestepper=RK3, imm_drag_u=0.0, imm_drag_v=0.0,
    free_drift_kind=FD_NONE, top_rho=1.3, top_Cd=1.2e-3,
)


class OracleModel:
    """Owns numpy copies of every field and drives the C oracle on them."""

    def __init__(self, Nx, Ny, Hx, Hy, topo=(PERIODIC, PERIODIC), dx=1.0, dy=1.0, params=None, fields=None,
                 metrics=None, mask=None, fold=None):
        self.Nx, self.Ny, self.Hx, self.Hy, self.topo = Nx, Ny, Hx, Hy, tuple(topo)
        prm = dict(DEFAULT_PARAMS)
        prm.update(params or {})
        self.prm = prm
        fields = fields or {}
        self.arr = {}
        snow = fields.get("hs") is not None
        for n in _STATE_NAMES + ("top_x", "top_y", "ue", "ve", "fd_u", "fd_v"):
            shp = parent_shape(Nx, Ny, Hx, Hy, self.topo, LOC[n])
            if n in fields and fields[n] is not None:
                a = np.ascontiguousarray(fields[n], dtype=np.float64).copy()
                assert a.shape == shp, (n, a.shape, shp)
            elif n in ("top_x", "top_y", "ue", "ve", "fd_u", "fd_v") or (n in _SNOW_NAMES and not snow):
                a = None
            else:
                a = np.zeros(shp)
                if n == "alpha":
                    a[:] = prm["alpha_max"]  # evp.jl:161
            self.arr[n] = a
        self.g = Grid(Nx=Nx, Ny=Ny, Hx=Hx, Hy=Hy, topo_x=self.topo[0], topo_y=self.topo[1], dx=dx, dy=dy)
        self._keep = []
        if metrics is not None:
            two_d = next(iter(metrics.values())).ndim == 2   # orthogonal curvilinear grid: (Ny+2Hy+1) x (Nx+2Hx+1) arrays
            self.g.metric_kind = 2 if two_d else 1
            if two_d:
                self.g.metW = Nx + 2 * Hx + 1
            for k, v in metrics.items():
                v = np.ascontiguousarray(v, dtype=np.float64)
                assert v.shape == ((Ny + 2 * Hy + 1, Nx + 2 * Hx + 1) if two_d else (Ny + 2 * Hy + 1,)), (k, v.shape)
                self._keep.append(v)
                setattr(self.g, k, v.ctypes.data_as(C.POINTER(C.c_double)))
        if mask is not None:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            assert m.shape == (Ny + 2 * Hy, Nx + 2 * Hx)
            self._keep.append(m)
            self.g.mask = m.ctypes.data_as(C.POINTER(C.c_uint8))
        if self.topo[1] == FOLDED:
            # fold = dict(maps={loc: (target, source)}, sign_velocity=-1.0, sign_external=1.0), loc in ((0,0),(1,0),(0,1),(1,1))
            assert fold is not None, "a FOLDED y axis needs its copy lists"
            for loc, (tg, sr) in fold["maps"].items():
                k = loc[0] + 2 * loc[1]
                tg = np.ascontiguousarray(tg, dtype=np.int32); sr = np.ascontiguousarray(sr, dtype=np.int32)
                assert tg.shape == sr.shape
                self._keep += [tg, sr]
                self.g.fold_target[k] = tg.ctypes.data_as(C.POINTER(C.c_int32))
                self.g.fold_source[k] = sr.ctypes.data_as(C.POINTER(C.c_int32))
                self.g.fold_count[k] = tg.size
            self.g.fold_sign_velocity = fold.get("sign_velocity", -1.0)
            self.g.fold_sign_external = fold.get("sign_external", 1.0)
        self.p = Params()
        for k, v in prm.items():
            if k == "f_ff":   # HydrostaticSphericalCoriolis: j-indexed f at (Face, Face)
                v = np.ascontiguousarray(v, dtype=np.float64)
                assert v.shape == (Ny + 2 * Hy + 1,)
                self._keep.append(v)
                self.p.f_ff = v.ctypes.data_as(C.POINTER(C.c_double))
            else:
                setattr(self.p, k, v)
        for n in ("top_x", "top_y", "ue", "ve", "fd_u", "fd_v"):
            setattr(self.p, n, _as_field(self.arr[n], Hx, Hy))
        self.s = State()
        for n in _STATE_NAMES:
            setattr(self.s, n, _as_field(self.arr[n], Hx, Hy))
        self.iteration = 0

    # -- views -------------------------------------------------------------------------------
    def interior(self, name):
        lx, ly = LOC[name]
        nx = self.Nx + (1 if (lx and self.topo[0] == BOUNDED) else 0)
        ny = self.Ny + (1 if (ly and self.topo[1] in (BOUNDED, FOLDED)) else 0)
        return self.arr[name][self.Hy:self.Hy + ny, self.Hx:self.Hx + nx]

    def _refs(self):
        return C.byref(self.g), C.byref(self.p), C.byref(self.s)

    # -- reference entry points --------------------------------------------------------------
    def fill_halo(self, name, which=0):
        lx, ly = LOC[name]
        f = getattr(self.s, name) if name in _STATE_NAMES else getattr(self.p, name)
        lib().csio_fill_halo(C.byref(self.g), C.byref(self.p), C.byref(f), lx, ly, which)

    def initialize_rheology(self):
        lib().csio_initialize_rheology(*self._refs())

    def compute_stresses(self, dt):
        lib().csio_compute_stresses(*self._refs(), C.c_double(dt))

    def u_velocity_step(self, dt):
        lib().csio_u_velocity_step(*self._refs(), C.c_double(dt))

    def v_velocity_step(self, dt):
        lib().csio_v_velocity_step(*self._refs(), C.c_double(dt))

    def time_step_momentum(self, dt, nsub=None):
        lib().csio_time_step_momentum(*self._refs(), C.c_double(dt), int(self.prm["substeps"] if nsub is None else nsub))

    def compute_tracer_tendencies(self):
        lib().csio_compute_tracer_tendencies(*self._refs())

    def dynamic_time_step(self, dt):
        lib().csio_dynamic_time_step(*self._refs(), C.c_double(dt))

    def update_state(self):
        lib().csio_update_state(*self._refs())

    def time_step(self, dt):
        lib().csio_time_step(*self._refs(), C.c_double(dt), 1 if self.iteration == 0 else 0)
        self.iteration += 1

    def cell_advection_timescale(self):
        return lib().csio_cell_advection_timescale(C.byref(self.g), C.byref(self.s))

    def stress_power_budget(self):
        out = (C.c_double * 3)()
        lib().csio_stress_power_budget(C.byref(self.g), C.byref(self.s), out)
        return tuple(out)

    def reconstruct(self, axis, name, order, bias, i, j):
        f = getattr(self.s, name)
        fn = lib().csio_reconstruct_x if axis == 0 else lib().csio_reconstruct_y
        return fn(C.byref(self.g), order, bias, C.byref(f), i, j)


# ---- slab thermodynamics (csi_oracle_thermo.c) ---------------------------------------------------
TOP_FLUX_BALANCE, TOP_PRESCRIBED = 0, 1
BOTTOM_EQUILIBRIUM, BOTTOM_PRESCRIBED = 0, 1
FLUX_CONST, FLUX_ARRAY, FLUX_RADIATIVE_EMISSION, FLUX_CONDUCTIVE, FLUX_LINEAR = 0, 1, 2, 3, 4


class ThermoParams(C.Structure):
    _fields_ = [
        ("density", C.c_double), ("heat_capacity", C.c_double), ("liquid_density", C.c_double),
        ("liquid_heat_capacity", C.c_double), ("reference_latent_heat", C.c_double), ("reference_temperature", C.c_double),
        ("liquidus_T0", C.c_double), ("liquidus_slope", C.c_double),
        ("top_bc", C.c_int32), ("snow_top_bc", C.c_int32), ("bottom_bc", C.c_int32), ("layered", C.c_int32),
        ("ice_conductivity", C.c_double), ("snow_conductivity", C.c_double),
        ("bottom_salinity", C.c_double), ("bottom_temperature", C.c_double),
        ("n_top_terms", C.c_int32), ("top_term_kind", C.c_int32 * 2), ("pad_", C.c_int32),
        ("top_flux_const", C.c_double), ("emissivity", C.c_double), ("stefan_boltzmann", C.c_double),
        ("emission_reference_temperature", C.c_double), ("bottom_flux_const", C.c_double),
        ("snowfall", C.c_double), ("snow_density", C.c_double), ("consolidation_thickness", C.c_double), ("ice_salinity", C.c_double),
        ("secant_tol", C.c_double), ("secant_maxiters", C.c_int32), ("pad2_", C.c_int32),
        ("linear_coefficient", C.c_double), ("linear_temperature", C.c_double),
        ("linear_times_concentration", C.c_int32), ("pad3_", C.c_int32),
    ]


THERMO_FIELDS = ("h", "a", "hs", "Tu", "Tus", "S", "hc", "Qtop", "Qbot", "Sb", "Tb", "snowfall", "rho_s", "mf_ice", "mf_snow", "mf_snowfall")


class ThermoState(C.Structure):
    _fields_ = [(n, Field) for n in THERMO_FIELDS]


# SeaIceModel / PhaseTransitions / SlabThermodynamics defaults (sea_ice_model.jl:66-83, SeaIceThermodynamics.jl:107-114,
# slab_sea_ice_thermodynamics.jl:36-48,84-90); secant defaults are RootSolvers' [RS-recall]
DEFAULT_THERMO = dict(
    density=917.0, heat_capacity=2000.0, liquid_density=999.8, liquid_heat_capacity=4186.0, reference_latent_heat=334e3,
    reference_temperature=0.0, liquidus_T0=0.0, liquidus_slope=0.054,
    top_bc=TOP_FLUX_BALANCE, snow_top_bc=TOP_FLUX_BALANCE, bottom_bc=BOTTOM_EQUILIBRIUM, layered=0,
    ice_conductivity=2.0, snow_conductivity=0.31, bottom_salinity=0.0, bottom_temperature=0.0,
    n_top_terms=1, top_term_kind=(FLUX_CONST, FLUX_CONST), top_flux_const=0.0,
    emissivity=1.0, stefan_boltzmann=5.67e-8, emission_reference_temperature=273.15, bottom_flux_const=0.0,
    snowfall=0.0, snow_density=330.0, consolidation_thickness=0.05, ice_salinity=0.0,
    secant_tol=1e-3, secant_maxiters=10000, linear_coefficient=0.0, linear_temperature=0.0, linear_times_concentration=0,
)


class ThermoOracle:
    """Numpy copies of the thermodynamic fields of one (c,c)-located column set + the C oracle's step."""

    def __init__(self, Nx, Ny, Hx, Hy, params=None, fields=None, rho_ice=900.0, shared=None):
        """`shared`: name -> existing array used in place (e.g. h, a, hs of an OracleModel for a coupled step)."""
        self.Nx, self.Ny, self.Hx, self.Hy = Nx, Ny, Hx, Hy
        shared = shared or {}
        prm = dict(DEFAULT_THERMO)
        prm.update(params or {})
        self.prm, self.rho_ice = prm, float(rho_ice)
        self.g = Grid(Nx=Nx, Ny=Ny, Hx=Hx, Hy=Hy, topo_x=PERIODIC, topo_y=PERIODIC, dx=1.0, dy=1.0)
        self.p = ThermoParams()
        for k, v in prm.items():
            if k == "top_term_kind":
                self.p.top_term_kind[0], self.p.top_term_kind[1] = v
            else:
                setattr(self.p, k, v)
        fields = fields or {}
        shp = (Ny + 2 * Hy, Nx + 2 * Hx)
        self.arr = {}
        self.s = ThermoState()
        always = ("h", "a", "Tu", "mf_ice", "mf_snow", "mf_snowfall") + (("hs", "Tus") if prm["layered"] else ())
        for n in THERMO_FIELDS:
            if n in shared:
                a = shared[n]
                assert a.dtype == np.float64 and a.shape == shp and a.flags["C_CONTIGUOUS"], n
            elif fields.get(n) is not None:
                a = np.ascontiguousarray(fields[n], dtype=np.float64).copy()
                assert a.shape == shp, (n, a.shape, shp)
            elif n in always:
                a = np.zeros(shp)
            else:
                a = None
            self.arr[n] = a
            setattr(self.s, n, _as_field(a, Hx, Hy))

    def interior(self, name):
        return self.arr[name][self.Hy:self.Hy + self.Ny, self.Hx:self.Hx + self.Nx]

    def step(self, dt):
        lib().csio_thermodynamic_time_step(C.byref(self.g), C.byref(self.p), C.byref(self.s), C.c_double(self.rho_ice), C.c_double(dt))


def pow4(x: float) -> float:
    return lib().csio_pow4(float(x))


def exp_cr(x: float) -> float:
    return lib().csio_exp(float(x))


def set_threads(n: int) -> int:
    return lib().csio_set_threads(int(n))

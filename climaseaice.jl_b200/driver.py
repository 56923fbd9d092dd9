"""Glue between synthetic `Case`s and the library: a device-resident SeaIceModel, and a
host-buffer stepper that goes through the `*_host` entry points of the C ABI (the end-to-end path
a Julia host with CPU arrays would take)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .model import (Bounded, FPlane, Field, HydrostaticSphericalCoriolis, LatitudeLongitudeGrid, Periodic, RectilinearGrid, SeaIceModel, SeaIceMomentumEquation, SemiImplicitStress,
                    SlabThermodynamics, SplitExplicitSolver, StressBalanceFreeDrift, UpwindBiased, ValueBoundaryCondition, WENO)
from .synthetic import LOC, Case


def grid_from_case(case: Case, device=None, partitioned_y=False, partitioned_x=False) -> RectilinearGrid:
    if case.metric_arrays is not None and next(iter(case.metric_arrays.values())).ndim == 2:   # orthogonal curvilinear grid
        from .model import OrthogonalSphericalShellGrid
        return OrthogonalSphericalShellGrid(size=(case.Nx, case.Ny), metrics=case.metric_arrays, halo=(case.Hx, case.Hy),
                                            topology=(case.topology[0], case.topology[1], "Flat"), device=device, partitioned_y=partitioned_y)
    if case.latlon is not None:
        return LatitudeLongitudeGrid(size=(case.Nx, case.Ny), longitude=case.latlon[0], latitude=case.latlon[1],
                                     halo=(case.Hx, case.Hy), topology=(case.topology[0], case.topology[1], "Flat"),
                                     device=device, metrics=case.metrics(), partitioned_y=partitioned_y)
    return RectilinearGrid(size=(case.Nx, case.Ny), x=(0, case.Lx), y=(0, case.Ly), halo=(case.Hx, case.Hy),
                           topology=(case.topology[0], case.topology[1], "Flat"), device=device, partitioned_y=partitioned_y, partitioned_x=partitioned_x)


def model_from_case(case: Case, solver_impl="auto", partition=None, device=None) -> SeaIceModel:
    grid = grid_from_case(case, device, partitioned_y=partition is not None, partitioned_x=partition is not None and len(partition) > 3 and partition[3] > 1)
    F = case.fields
    oc = case.ocean_const or (0.0, 0.0)
    fld = lambda n: Field(LOC[n], grid, F[n])
    # top: arrays or constants, read as the stress itself or as the atmosphere's velocity (SemiImplicitStress)
    tu, tv = (fld("top_x"), fld("top_y")) if "top_x" in F else (case.top_const if case.top_const else (None, None))
    if case.top_kind == "semi_implicit":
        top = SemiImplicitStress(ue=tu, ve=tv, rho_e=case.top_rho_Cd[0], Cd=case.top_rho_Cd[1])
    else:
        top = None if tu is None else dict(u=tu, v=tv)
    bu, bv = (fld("ue"), fld("ve")) if "ue" in F else oc
    if case.bottom_kind == "semi_implicit":
        bottom = SemiImplicitStress(ue=bu, ve=bv, rho_e=case.rho_e, Cd=case.Cd)
    elif case.bottom_kind == "stress":
        bottom = dict(u=bu, v=bv)
    else:
        bottom = None
    free_drift = None
    if case.free_drift == "fields":
        free_drift = dict(u=fld("fd_u"), v=fld("fd_v"))
    elif case.free_drift == "stress_balance":
        free_drift = StressBalanceFreeDrift()
    if case.f_ff() is not None:
        coriolis = HydrostaticSphericalCoriolis(case.rotation_rate, f_ff_override=case.f_ff())
    else:
        coriolis = FPlane(case.coriolis_f) if case.coriolis_f is not None else None
    dyn = SeaIceMomentumEquation(grid,
                                 coriolis=coriolis,
                                 top_momentum_stress=top, bottom_momentum_stress=bottom, free_drift=free_drift,
                                 solver=SplitExplicitSolver(substeps=case.substeps))
    bcs = {}
    if case.u_bc_value is not None:
        bcs["u"] = dict(north=ValueBoundaryCondition(case.u_bc_value), south=ValueBoundaryCondition(case.u_bc_value))
    if case.v_bc_value is not None:
        bcs["v"] = dict(west=ValueBoundaryCondition(case.v_bc_value), east=ValueBoundaryCondition(case.v_bc_value))
    adv = None if case.advection_order == 0 else (UpwindBiased(1) if case.advection_order == 1 else WENO(case.advection_order))
    thermo = {}
    if case.thermo is not None:   # coupled slab thermodynamics (bare ice): Tu / Qtop arrays + scalars
        thermo = dict(ice_thermodynamics=SlabThermodynamics(grid, top_surface_temperature=F["Tu"]),
                      top_heat_flux=Field((0, 0), grid, F["Qtop"]) if "Qtop" in F else None,
                      bottom_heat_flux=case.thermo.get("bottom_heat_flux", 0.0), ice_salinity=case.thermo.get("ice_salinity", 0.0))
    m = SeaIceModel(grid, dynamics=dyn, advection=adv, timestepper=case.timestepper, boundary_conditions=bcs,
                    solver_impl=solver_impl, partition=partition, immersed_mask=case.mask, immersed_drag=case.immersed_drag,
                    snow_thickness="hs" in F, fold=case.fold, **thermo)
    m.set(h=F["h"], a=F["a"], u=F["u"], v=F["v"])
    if "hs" in F:
        m.set(hs=F["hs"])
    return m


class HostStepper:
    """Drives csi_time_step_host / csi_evp_substeps_host on pinned host arrays: every call uploads
    the inputs, runs on the GPU and downloads the results (the end-to-end figure of bench.py)."""

    def __init__(self, case: Case, solver_impl="auto", device_index=0, partition=None, unique_id=None):
        self.case = case
        self.model = model_from_case(case, solver_impl=solver_impl, device=f"cuda:{device_index}", partition=partition)
        if partition is not None:   # one rank of a partition: host blocks with halos, NCCL exchange inside the call
            self.model.comm_init(unique_id)
        self.host = {}
        # arrays the host entry points copy in either direction are pinned; the others (P, u^n, zeta, Delta, G^n ...: written and
        # read on the device only) just need a host address and a shape
        copied = {"u", "v", "h", "a", "s11", "s22", "s12", "alpha", "top_x", "top_y", "ue", "ve", "hs", "fd_u", "fd_v", "um", "vm"}
        for n, fld in self.model.all_fields().items():
            t = torch.empty(fld.parent.shape, dtype=torch.float64)
            if n in copied:
                t = t.pin_memory()
                t.copy_(fld.parent)
            self.host[n] = t
        self.fields = L.csi_fields()
        for n, t in self.host.items():
            a = L.csi_array()
            a.ptr = t.data_ptr()
            a.ny_tot, a.nx_tot = t.shape
            a.off_x, a.off_y = case.Hx, case.Hy
            setattr(self.fields, n, a)
        self.iteration = 0

    def last_transfer_bytes(self):
        """(host->device, device->host) bytes of the last call, as counted by the library."""
        a, b = C.c_uint64(), C.c_uint64()
        L.check(L.lib().csi_last_transfer_bytes(self.model._handle, C.byref(a), C.byref(b)), self.model._handle)
        return a.value, b.value

    def time_step(self, dt, nsteps=1):
        h = self.model._handle
        L.check(L.lib().csi_time_step_host(h, C.byref(self.fields), float(dt), int(nsteps), 1 if self.iteration == 0 else 0), h)
        self.iteration += nsteps

    def evp_substeps(self, dt, nsub):
        h = self.model._handle
        L.check(L.lib().csi_evp_substeps_host(h, C.byref(self.fields), float(dt), int(nsub)), h)

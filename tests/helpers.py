"""Test helpers: build the CPU oracle from a synthetic Case, and compare fields.

The oracle is the checker: it is imported only here, in smoke() and in bench.py's CPU legs."""
from __future__ import annotations

import numpy as np

from oracle import oracle as O


def oracle_params(case) -> dict:
    F = case.fields
    p = dict(substeps=case.substeps, advection_order=case.advection_order,
             timestepper=O.RK3 if case.timestepper == "SplitRungeKutta3" else O.FE,
             coriolis_kind=0 if case.coriolis_f is None else 1, f=case.coriolis_f or 0.0,
             rho_e=case.rho_e, Cd=case.Cd, top_rho=case.top_rho_Cd[0], top_Cd=case.top_rho_Cd[1])
    if case.f_ff() is not None:
        p.update(coriolis_kind=2, f_ff=case.f_ff())
    has_top = "top_x" in F or bool(case.top_const)
    if case.top_kind == "semi_implicit":
        p.update(top_kind=O.STRESS_SEMI_IMPLICIT)
    else:
        p.update(top_kind=O.STRESS_FIELD if "top_x" in F else (O.STRESS_CONST if has_top else O.STRESS_NONE))
    if case.top_const and "top_x" not in F:
        p.update(top_tx=case.top_const[0], top_ty=case.top_const[1])
    if case.bottom_kind == "semi_implicit":
        p.update(bot_kind=O.STRESS_SEMI_IMPLICIT)
    elif case.bottom_kind == "stress":
        p.update(bot_kind=O.STRESS_FIELD if "ue" in F else O.STRESS_CONST)
    else:
        p.update(bot_kind=O.STRESS_NONE)
    if case.ocean_const and "ue" not in F:
        p.update(ue_c=case.ocean_const[0], ve_c=case.ocean_const[1])
    p.update(free_drift_kind={None: O.FD_NONE, "fields": O.FD_FIELDS, "stress_balance": O.FD_STRESS_BALANCE}[case.free_drift])
    p.update(imm_drag_u=case.immersed_drag[0], imm_drag_v=case.immersed_drag[1])
    if case.u_bc_value is not None:
        p.update(u_sn_bc=1, u_sn_val=case.u_bc_value)
    if case.v_bc_value is not None:
        p.update(v_we_bc=1, v_we_val=case.v_bc_value)
    return p


def oracle_from_case(case, **overrides) -> O.OracleModel:
    topo = tuple(dict(Periodic=O.PERIODIC, Bounded=O.BOUNDED, Folded=O.FOLDED)[t] for t in case.topology)
    p = oracle_params(case)
    p.update(overrides)
    return O.OracleModel(case.Nx, case.Ny, case.Hx, case.Hy, topo=topo, dx=case.dx, dy=case.dy, params=p,
                         fields={k: v.copy() for k, v in case.fields.items()}, mask=case.mask, metrics=case.metrics(), fold=case.fold)


# GPU field name -> oracle field name
NAME_MAP = dict(u="u", v="v", h="h", a="a", s11="s11", s22="s22", s12="s12", alpha="alpha", zeta_c="zc", zeta_f="zf",
                delta="delta", P="P", un="un", vn="vn", Gh="Gh", Ga="Ga", hs="hs", Ghs="Ghs")


def rel_err(a, b):
    """max|a-b| / max|b| (fields cross zero, SURVEY section 7 hard part 3)."""
    d = float(np.max(np.abs(a - b)))
    s = float(np.max(np.abs(b)))
    return d / s if s > 0 else d


def interior_of(arr, case):
    return arr[case.Hy:arr.shape[0] - case.Hy, case.Hx:arr.shape[1] - case.Hx]


def compare_model(model, oracle, case, names=("u", "v", "h", "a", "s11", "s22", "s12"), interior_only=True):
    out = {}
    F = model.all_fields()
    for n in names:
        g = F[n].numpy()
        r = oracle.arr[NAME_MAP[n]]
        if interior_only:
            g, r = interior_of(g, case), interior_of(r, case)
        out[n] = (rel_err(g, r), bool(np.array_equal(g, r)))
    return out


def thermo_oracle_from_case(case, o):
    """The thermodynamics oracle of a coupled case (Case.thermo), working in place on the dynamics oracle's h, aice."""
    F = case.fields
    prm = dict(bottom_flux_const=case.thermo.get("bottom_heat_flux", 0.0), ice_salinity=case.thermo.get("ice_salinity", 0.0),
               n_top_terms=1, top_term_kind=((O.FLUX_ARRAY if "Qtop" in F else O.FLUX_CONST), O.FLUX_CONST))
    return O.ThermoOracle(case.Nx, case.Ny, case.Hx, case.Hy, params=prm, fields={k: F[k] for k in ("Tu", "Qtop") if k in F},
                          shared=dict(h=o.arr["h"], a=o.arr["a"]), rho_ice=900.0)


def coupled_oracle_step(o, t, dt):
    """time_step!(model, dt) with dynamics and thermodynamics, stage by stage (fe.jl:13-34; rk.jl:29-94, beta = 3, 2, 1)."""
    if o.iteration == 0:
        o.update_state()
    if o.prm["timestepper"] == O.FE:
        stages = [dt]
    else:
        stages = [dt / 3, dt / 2, dt / 1]
        for n in ("h", "a", "u", "v"):
            o.arr[n + "m"][:] = o.arr[n]
    for dtau in stages:
        o.compute_tracer_tendencies()
        o.time_step_momentum(dtau)
        o.dynamic_time_step(dtau)
        t.step(dtau)
        o.update_state()
    o.iteration += 1

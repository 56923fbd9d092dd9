"""GPU parity of the slab thermodynamics kernel against the CPU oracle: bitwise on every
output, on seeded fields that drive every branch (consolidated / thin ice, melt / freeze / extinction, open water,
snow melt, flooding, partial cover), standalone, through a thermodynamics-only model, and coupled to the dynamics."""
import numpy as np
import pytest

import climaseaice_b200 as csi
from climaseaice_b200.driver import grid_from_case, model_from_case
from climaseaice_b200.synthetic import periodic_case
from oracle import oracle as O
from tests.helpers import compare_model, interior_of, oracle_from_case

pytestmark = pytest.mark.gpu


def thermo_fields(N, H, seed, snow):
    rng = np.random.default_rng(seed)
    shp = (N + 2 * H, N + 2 * H)
    u = lambda lo, hi: rng.uniform(lo, hi, shp)
    h = u(0.0, 2.5)
    h[u(0, 1) < 0.15] = 0.01 * rng.uniform(0, 4)            # thinner than the consolidation thickness
    a = np.clip(u(-0.1, 1.2), 0.0, 1.0)
    a[u(0, 1) < 0.1] = 0.0
    h[a == 0] = 0.0
    F = dict(h=h, a=a, Tu=u(-30.0, 0.0), S=u(0.0, 8.0), Qtop=u(-400.0, 300.0), Qbot=u(-50.0, 30.0), Sb=u(28.0, 36.0))
    if snow:
        hs = np.where(a > 0, u(0.0, 0.6), 0.0)
        hs[u(0, 1) < 0.2] = 0.0
        hs[u(0, 1) < 0.1] *= 8.0                             # heavy snow: flooding
        F.update(hs=hs, Tus=u(-35.0, 0.0), snowfall=u(0.0, 5e-5), rho_s=u(250.0, 400.0))
    return F


def build_pair(N, H, F, snow, top_terms, timestepper="ForwardEuler", **kw):
    """The same thermodynamics-only model on the GPU (host mirror) and on the oracle."""
    case = periodic_case(N, H=H)
    grid = grid_from_case(case)
    fld = lambda a: csi.Field((csi.Center, csi.Center), grid, a)
    ice = csi.SlabThermodynamics(grid, top_surface_temperature=F["Tu"], bottom_heat_boundary_condition=csi.IceWaterThermalEquilibrium(fld(F["Sb"])))
    snw = csi.snow_slab_thermodynamics(grid, top_surface_temperature=F["Tus"]) if snow else None
    top = tuple(fld(F["Qtop"]) if t == "array" else t for t in top_terms)
    m = csi.SeaIceModel(grid, ice_thermodynamics=ice, snow_thermodynamics=snw, timestepper=timestepper,
                        top_heat_flux=top if len(top) > 1 else top[0], bottom_heat_flux=fld(F["Qbot"]), ice_salinity=fld(F["S"]),
                        snowfall=fld(F["snowfall"]) if snow else 0.0, snow_density=fld(F["rho_s"]) if snow else 330.0, **kw)
    m.set(h=F["h"], a=F["a"])
    if snow:
        m.set(hs=F["hs"])
    kinds, prm = [], dict(layered=1 if snow else 0, n_top_terms=len(top_terms))
    for t in top_terms:
        if isinstance(t, str):
            kinds.append(O.FLUX_ARRAY)
        elif isinstance(t, csi.RadiativeEmission):
            kinds.append(O.FLUX_RADIATIVE_EMISSION)
        elif isinstance(t, csi.LinearHeatFlux):
            kinds.append(O.FLUX_LINEAR)
            prm.update(linear_coefficient=t.coefficient, linear_temperature=t.temperature, linear_times_concentration=int(t.times_concentration))
        else:
            kinds.append(O.FLUX_CONST)
            prm.update(top_flux_const=float(t))
    prm["top_term_kind"] = tuple(kinds + [O.FLUX_CONST])[:2]
    o = O.ThermoOracle(N, N, H, H, params=prm, fields=F)
    return m, o, case


def assert_same(m, o, case, names):
    _, _, arrays = m._thermo
    for n in names:
        g = interior_of(arrays[n].numpy(), case)
        r = interior_of(o.arr[n], case)
        assert np.array_equal(g, r, equal_nan=True), (n, float(np.nanmax(np.abs(g - r))))


OUT_ICE = ("h", "a", "Tu", "mf_ice", "mf_snow", "mf_snowfall")
OUT_SNOW = OUT_ICE + ("hs", "Tus")


@pytest.mark.parametrize("snow", [False, True])
@pytest.mark.parametrize("top", ["array", "array+emission", "linear+const"])
def test_thermodynamic_kernel_bitwise(snow, top):
    N, H = 96, 4
    F = thermo_fields(N, H, 11 + snow, snow)
    terms = dict(array=("array",), **{"array+emission": ("array", csi.RadiativeEmission()),
                                      "linear+const": (csi.LinearHeatFlux(6.15, -12.0, True), -40.0)})[top]
    m, o, case = build_pair(N, H, F, snow, terms)
    for dt in (3600.0, 600.0, 7200.0):
        m.thermodynamic_time_step(dt)
        o.step(dt)
        assert_same(m, o, case, OUT_SNOW if snow else OUT_ICE)
    h = interior_of(m.ice_thickness.numpy(), case)
    a = interior_of(m.ice_concentration.numpy(), case)
    assert np.isfinite(h).all() and (a >= 0).all() and (a <= 1).all()
    assert (h == 0).any() and (h > 0).any()                      # some columns melted away, most did not
    m.close()


@pytest.mark.parametrize("timestepper", ["ForwardEuler", "SplitRungeKutta3"])
def test_mass_flux_closure_through_the_model(timestepper):
    """test_thermodynamic_mass_fluxes.jl: d/dt (rho_i h aice + rho_s hs aice) = the three recorded fluxes, per column,
    for a thermodynamics-only model under both time steppers (the last RK stage advances the full dt from Psi^-)."""
    N, H, dt = 48, 4, 3600.0
    F = thermo_fields(N, H, 3, True)
    F["rho_s"][:] = 330.0
    m, o, case = build_pair(N, H, F, True, ("array",), timestepper=timestepper)
    I = lambda f: interior_of(f.numpy(), case)
    m.update_state()
    M0 = 900.0 * I(m.ice_thickness) * I(m.ice_concentration) + 330.0 * I(m.snow_thickness) * I(m.ice_concentration)
    m.time_step(dt)
    M1 = 900.0 * I(m.ice_thickness) * I(m.ice_concentration) + 330.0 * I(m.snow_thickness) * I(m.ice_concentration)
    total = I(m.mass_fluxes["ice"]) + I(m.mass_fluxes["snow"]) + I(m.mass_fluxes["intercepted_snowfall"])
    expected = (M1 - M0) / dt
    assert np.all(np.abs(total - expected) <= 1e-12 * np.maximum(1.0, np.abs(expected)))
    m.close()


def test_coupled_dynamics_and_thermodynamics_step():
    """One model with dynamics AND thermodynamics: csi_time_step runs the thermodynamic kernel after dynamic_time_step!
    in every stage (fe.jl:27-30, rk.jl:89-91).  Oracle: the same stage sequence driven from Python."""
    N, H = 48, 7
    case = periodic_case(N, H=H, substeps=10)
    F = thermo_fields(N, H, 21, False)
    m = model_from_case(case, solver_impl="auto")
    grid = m.grid
    # attach thermodynamics to the existing dynamics model (same h, aice arrays)
    m.ice_thermodynamics = csi.SlabThermodynamics(grid, top_surface_temperature=F["Tu"])
    m.snow_thermodynamics, m.phase_transitions = None, csi.PhaseTransitions()
    qtop = csi.Field((csi.Center, csi.Center), grid, F["Qtop"])
    m._setup_thermodynamics(qtop, -5.0, 0.0, 330.0, 0.05, 3.0)
    o = oracle_from_case(case)
    t = O.ThermoOracle(N, N, H, H, params=dict(n_top_terms=1, top_term_kind=(O.FLUX_ARRAY, O.FLUX_CONST), bottom_flux_const=-5.0, ice_salinity=3.0),
                       fields=dict(Tu=F["Tu"], Qtop=F["Qtop"]), shared=dict(h=o.arr["h"], a=o.arr["a"]))
    for step in range(2):
        m.time_step(case.dt)
        if step == 0:
            o.update_state()
        # rk_substep! x 3 with beta = 3, 2, 1  (rk.jl:29-94)
        for n in ("h", "a", "u", "v"):
            o.arr[n + "m"][:] = o.arr[n]
        for beta in (3, 2, 1):
            dtau = case.dt / beta
            o.compute_tracer_tendencies()
            o.time_step_momentum(dtau)
            o.dynamic_time_step(dtau)
            t.step(dtau)
            o.update_state()
    res = compare_model(m, o, case)
    for n, (err, same) in res.items():
        assert same, (n, err)
    assert_same(m, t, case, ("Tu", "mf_ice"))
    m.close()


def test_thermodynamics_argument_errors():
    N, H = 16, 4
    F = thermo_fields(N, H, 1, False)
    m, o, case = build_pair(N, H, F, False, ("array",))
    tc, tf, _ = m._thermo
    import ctypes as C
    from climaseaice_b200 import _lib as L
    bad = L.csi_thermo_config.from_buffer_copy(tc)
    bad.n_top_terms = 3
    assert L.lib().csi_thermodynamic_time_step(m._handle, C.byref(bad), C.byref(tf), 1.0, None) == -1
    bad = L.csi_thermo_config.from_buffer_copy(tc)
    bad.secant_maxiters = 0
    assert L.lib().csi_thermodynamic_time_step(m._handle, C.byref(bad), C.byref(tf), 1.0, None) == -1
    nof = L.csi_thermo_fields.from_buffer_copy(tf)
    nof.Qtop.ptr = None
    assert L.lib().csi_thermodynamic_time_step(m._handle, C.byref(tc), C.byref(nof), 1.0, None) == -1
    assert b"Qtop" in L.lib().csi_last_error(m._handle)
    with pytest.raises(NotImplementedError, match="closures"):
        csi.SeaIceModel(m.grid, ice_thermodynamics=csi.SlabThermodynamics(m.grid), top_heat_flux=lambda *a: 0.0)
    m.close()


@pytest.mark.parametrize("impl", ("unfused", "fused"))
@pytest.mark.parametrize("timestepper", ["SplitRungeKutta3", "ForwardEuler"])
def test_arctic_cap_config5_reduced(impl, timestepper):
    """BASELINE config 5 in miniature: coupled slab thermodynamics + EVP dynamics + WENO advection on a zonally periodic
    lat-lon cap (general kernels: j-dependent metrics), against the staged oracle."""
    from climaseaice_b200.synthetic import arctic_cap_case
    from tests.helpers import coupled_oracle_step, thermo_oracle_from_case
    case = arctic_cap_case(96, 32, H=5, substeps=12, timestepper=timestepper)
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    t = thermo_oracle_from_case(case, o)
    for _ in range(2):
        m.time_step(case.dt)
        coupled_oracle_step(o, t, case.dt)
    for n, (err, same) in compare_model(m, o, case).items():
        assert same, (n, err)
    assert_same(m, t, case, ("Tu", "mf_ice"))
    mf = interior_of(m.mass_fluxes["ice"].numpy(), case)
    assert (mf > 0).any() and (mf < 0).any()           # freezing near the pole, melting at the edge
    m.close()

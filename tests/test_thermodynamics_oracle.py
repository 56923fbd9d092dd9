"""The thermodynamics ORACLE (oracle/csi_oracle_thermo.c) against the properties the reference's own tests hold for
this path -- it ships no known-answer values, so these pin structure, signs and conservation, not bits:
  test/test_thermodynamic_mass_fluxes.jl   mass closure, signs, melt to extinction, lateral growth
  test/test_energy_conservation.jl         per-step energy residual at round-off (bare / snow / precipitation / aice < 1)
  test/test_snow_thermodynamics.jl         flooding, snowfall accumulation, snow melts before ice, insulation
"""
import numpy as np
import pytest

from oracle import oracle as O


def column(params=None, rho_ice=900.0, **state):
    """A single column (RectilinearGrid(size=()) in the reference's tests)."""
    fields = {k: np.full((3, 3), float(v)) for k, v in state.items()}
    return O.ThermoOracle(1, 1, 1, 1, params=params, fields=fields, rho_ice=rho_ice)


def val(o, n):
    return float(o.interior(n)[0, 0])


@pytest.mark.parametrize("top,h0,a0,check", [
    (100.0, 1.0, 1.0, "freezing"), (-200.0, 1.0, 1.0, "melting"), (-1e5, 0.2, 1.0, "extinction"), (300.0, 1.0, 0.95, "partial")])
def test_bare_ice_mass_flux_closure(top, h0, a0, check):
    """test_thermodynamic_mass_fluxes.jl:53-112 (one kernel call = the ForwardEuler step of a column)."""
    dt = 3600.0
    o = column(dict(top_flux_const=top, bottom_flux_const=10.0), h=h0, a=a0)
    m0 = 900.0 * h0 * a0
    o.step(dt)
    h, a = val(o, "h"), val(o, "a")
    expected = (900.0 * h * a - m0) / dt
    total = val(o, "mf_ice") + val(o, "mf_snow") + val(o, "mf_snowfall")
    assert abs(total - expected) <= 1e-12 * max(1.0, abs(expected))
    assert val(o, "mf_snow") == 0 and val(o, "mf_snowfall") == 0
    if check == "freezing":
        assert val(o, "mf_ice") > 0
    elif check == "melting":
        assert val(o, "mf_ice") < 0
    elif check == "extinction":
        assert h == 0 and a == 0
    else:
        assert a > 0.95 and val(o, "mf_ice") > 0


@pytest.mark.parametrize("top,h0,hs0,check", [(100.0, 1.0, 0.1, "freezing"), (-200.0, 1.0, 0.1, "melting"), (-1e5, 0.2, 0.05, "extinction")])
def test_snow_ice_mass_flux_closure(top, h0, hs0, check):
    """test_thermodynamic_mass_fluxes.jl:114-177"""
    dt, Ps = 3600.0, 1e-5
    o = column(dict(layered=1, top_flux_const=top, bottom_flux_const=10.0, snowfall=Ps), h=h0, a=1.0, hs=hs0)
    m0 = 900.0 * h0 + 330.0 * hs0
    o.step(dt)
    h, a, hs = val(o, "h"), val(o, "a"), val(o, "hs")
    expected = ((900.0 * h * a + 330.0 * hs * a) - m0) / dt
    total = val(o, "mf_ice") + val(o, "mf_snow") + val(o, "mf_snowfall")
    assert abs(total - expected) <= 1e-12 * max(1.0, abs(expected))
    if check == "extinction":
        assert h == 0 and a == 0 and hs == 0
    else:
        assert abs(val(o, "mf_snowfall") - Ps * a) <= 1e-12 * Ps
        if check == "freezing":
            assert val(o, "mf_ice") > 0
        else:
            assert val(o, "mf_snow") < 0


def _energy_residual(snow, precipitation, melting, a0=1.0, hs0=0.2, per_cell=False):
    """test_energy_conservation.jl:21-89,117-176: E = -aice (rho_i L h + rho_s L hs); dE = (-Qa + Ql + Qp) dt."""
    Ta = 5.0 if melting else -15.0
    coef = 1e-3 * 1.225 * 1004 * 5
    Qb = -20.0 if melting else -5.0
    Ps = 6e-5 if precipitation else 0.0
    prm = dict(layered=1 if snow else 0, n_top_terms=1, top_term_kind=(O.FLUX_LINEAR, O.FLUX_CONST), linear_coefficient=coef,
               linear_temperature=Ta, linear_times_concentration=1, bottom_flux_const=Qb, snowfall=Ps, consolidation_thickness=0.05)
    state = dict(h=1.0, a=a0)
    if snow:
        state["hs"] = hs0
    o = column(prm, **state)
    L, rho_i, rho_s = 334e3, 900.0, (330.0 if snow else 0.0)       # latent_heat(pt, 0) = reference_latent_heat
    dt, worst = 600.0, 0.0
    for _ in range(200):
        h0_, a0_, hs0_ = val(o, "h"), val(o, "a"), (val(o, "hs") if snow else 0.0)
        E0 = -a0_ * (rho_i * L * h0_ + rho_s * L * hs0_)
        o.step(dt)
        h1, a1, hs1 = val(o, "h"), val(o, "a"), (val(o, "hs") if snow else 0.0)
        E1 = -a1 * (rho_i * L * h1 + rho_s * L * hs1)
        # the flux the kernel last evaluated: at the converged top temperature, with the concentration of the step's start
        Tu = val(o, "Tus") if snow else val(o, "Tu")
        Qa = coef * (Tu - Ta) * a0_
        Qp = -L * Ps if (precipitation and a1 > 0) else 0.0
        expected = (-Qa + Qb + Qp) * dt
        scale = max(abs(E0), abs(E1), abs(expected), 1.0)
        worst = max(worst, abs((E1 - E0) - expected) / scale)
        if h1 <= 0 and a1 <= 0:
            break
    return worst


@pytest.mark.parametrize("snow,precipitation,melting", [(False, False, False), (False, False, True), (True, False, False),
                                                         (True, False, True), (True, True, False), (True, True, True)])
def test_energy_conservation(snow, precipitation, melting):
    # the reference's own tolerance is 1e-15 (test_energy_conservation.jl:92); its bare-ice latent heat at Tu != 0 and
    # Tb != 0 is evaluated at 0 degC there (S = 0 => Tb = 0), as here
    assert _energy_residual(snow, precipitation, melting) < 1e-15


@pytest.mark.parametrize("a0,hs0,melting", [(0.5, 0.15, True), (0.7, 0.3, True), (0.5, 0.15, False)])
def test_energy_conservation_partial_cover(a0, hs0, melting):
    """test_energy_conservation.jl:178-199: the closed-form self-consistent solve keeps the residual at round-off for aice < 1."""
    assert _energy_residual(True, False, melting, a0=a0, hs0=hs0) < 1e-13


def test_flooding_snowfall_and_snow_melts_first():
    """test_snow_thermodynamics.jl:104-186"""
    o = column(dict(layered=1, top_bc=O.TOP_PRESCRIBED), h=0.5, a=1.0, hs=1.0, Tu=-5.0)
    o.step(1.0)
    assert val(o, "h") > 0.5 and val(o, "hs") < 1.0                 # negative freeboard: snow turns into ice
    o = column(dict(layered=1, snowfall=1e-5), h=1.0, a=1.0, hs=0.0)
    o.step(3600.0)
    assert val(o, "hs") > 0                                          # snowfall accumulates
    o = column(dict(layered=1, top_flux_const=-100.0), h=2.0, a=1.0, hs=0.1)
    o.step(3600.0)
    assert val(o, "hs") < 0.1                                        # incoming heat melts snow first
    assert val(o, "h") >= 2.0 - 1e-3


def test_interface_temperature_lies_between_surface_and_base():
    """test_snow_thermodynamics.jl:76-102, through the kernel: prescribed snow-surface temperature, no external flux.
    Tsi = Tb + (Tu - Tb) Ri / (Rs + Ri), stored as the ice slab's top temperature; = Tu without snow."""
    prm = dict(layered=1, snow_top_bc=O.TOP_PRESCRIBED, bottom_bc=O.BOTTOM_PRESCRIBED, bottom_temperature=-1.8, consolidation_thickness=0.05)
    for hs in (0.0, 0.3, 1.0):
        o = column(prm, h=1.0, a=1.0, hs=hs, Tus=-10.0)
        o.step(600.0)
        Tsi = val(o, "Tu")
        Ri, Rs = 1.0 / 2.0, hs / 0.31
        assert Tsi == pytest.approx(-1.8 + (-10.0 + 1.8) * Ri / (Rs + Ri), abs=1e-12)
        assert (Tsi == pytest.approx(-10.0, abs=1e-12)) if hs == 0.0 else (-10.0 < Tsi < -1.8)
        assert np.isfinite(val(o, "h")) and 0.99 < val(o, "a") <= 1.0


def test_radiative_emission_balance_and_pow4():
    """RadiativeEmission (boundary_fluxes.jl:117-144) balanced against conduction: the secant solve lands on
    eps sigma (T + Tr)^4 = -k (T - Tb) / h; (T + Tr)^4 is correctly rounded (Julia's compensated Float64^Int)."""
    from fractions import Fraction
    rng = np.random.default_rng(5)
    for x in rng.uniform(150, 320, 3000):
        assert O.pow4(float(x)) == float(Fraction(float(x)) ** 4)
    prm = dict(top_term_kind=(O.FLUX_RADIATIVE_EMISSION, O.FLUX_CONST), n_top_terms=2, top_flux_const=-250.0,
               secant_tol=1e-10, bottom_bc=O.BOTTOM_PRESCRIBED, bottom_temperature=-1.8)
    o = column(prm, h=1.5, a=1.0, Tu=-20.0)
    o.step(60.0)
    T = val(o, "Tu")
    lhs = 5.67e-8 * (T + 273.15) ** 4 - 250.0
    rhs = -2.0 * (T - (-1.8)) / 1.5
    assert lhs == pytest.approx(rhs, rel=1e-9) and -60 < T < -1.8


def test_unconsolidated_ice_and_prescribed_temperature_paths():
    # thinner than the consolidation thickness: Tu = bottom temperature, no conduction (slab_thermodynamics_tendencies.jl:112-114)
    o = column(dict(top_flux_const=50.0, bottom_salinity=30.0), h=0.02, a=0.5, Tu=-7.0)
    o.step(600.0)
    assert val(o, "Tu") == 0.0 - 0.054 * 30.0
    # PrescribedTemperature top with the model's default external flux (= the conductive flux): no top imbalance,
    # the slab grows from below by conduction only
    prm = dict(top_bc=O.TOP_PRESCRIBED, top_term_kind=(O.FLUX_CONDUCTIVE, O.FLUX_CONST))
    o = column(prm, h=1.0, a=1.0, Tu=-10.0)
    o.step(3600.0)
    growth = val(o, "h") - 1.0
    assert val(o, "Tu") == -10.0
    assert growth == pytest.approx(2.0 * 10.0 / 1.0 / (900.0 * 334e3) * 3600.0, rel=1e-9)

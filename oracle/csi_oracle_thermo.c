/*
 * csi_oracle_thermo.c -- CPU ORACLE for the slab thermodynamics step.
 *
 * TEST INFRASTRUCTURE ONLY (see csi_oracle.h).  PARITY UNPINNED, twice over:
 *   - the in-tree physics (src/SeaIceThermodynamics/ *.jl) is restated literally, left to right, but the
 *     reference ships no known-answer values for it, only closure/sign properties
 *     (test/test_thermodynamic_mass_fluxes.jl, test/test_snow_thermodynamics.jl), which tests/ restates;
 *   - the surface-temperature solve calls RootSolvers.jl (`SecantMethod`, `find_zero`, `CompactSolution`;
 *     Project.toml:11,23 allows "0.3, 0.4, 1.0", no Manifest => version unpinned, package not in the tree).
 *     Its published algorithm is restated in `secant_find_zero` below [RS-recall]: iterate
 *     x1 <- x1 - y1 * (x1 - x0) / (y1 - y0) from (x0, x1), stop when |x0 - x1| < tol (SolutionTolerance,
 *     default 1e-3) or after maxiters (default 10 000) iterations, return the last x1.
 *   - `(T + Tr)^4` of RadiativeEmission is Julia's Float64^Int `pow_body` (compensated squaring with fma)
 *     [Julia-recall], restated in `jl_pow4`.
 *
 * One function per reference function; file:line cites are relative to /root/reference.
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#include "csi_oracle.h"

#define F(f, i, j) ((f).p[(size_t)((i)-1 + (f).ox) + (size_t)((j)-1 + (f).oy) * (size_t)(f).sx])
typedef const csio_grid *G;
typedef const csio_thermo_params *TP;

/* ---- Julia scalar semantics ---- */
static inline double jl_max(double a, double b)
{
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return signbit(a) ? b : a;
    return a > b ? a : b;
}
static inline double jl_min(double a, double b)
{ /* Base.min: NaN-propagating, min(-0.0, +0.0) = -0.0 */
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return signbit(a) ? a : b;
    return a < b ? a : b;
}
static inline double jl_mul_bool(double x, int b) { return b ? x : copysign(0.0, x); }

/* x^4 as Base.pow_body(x::Float64, 4) [Julia-recall]: two compensated squarings, then x + err */
static inline double jl_pow4(double x)
{
    double xnlo = 0.0, err, hi;
    /* n = 4: bit 0 clear */
    err = x * 2 * xnlo;
    hi = x * x;
    xnlo = fma(x, x, -hi);
    x = hi;
    xnlo += err;
    /* n = 2: bit 0 clear */
    err = x * 2 * xnlo;
    hi = x * x;
    double lo = fma(x, x, -hi);
    x = hi;
    xnlo = lo + err;
    /* n = 1: leave the loop; y = 1, ynlo = 0: err = muladd(y, xnlo, x * ynlo) = xnlo */
    err = 1.0 * xnlo + x * 0.0;
    return (isfinite(x) && isfinite(err)) ? x * 1.0 + err : x * 1.0;
}

/* ---- per-cell inputs: array when present, else the constant ---- */
static inline double opt(csio_field f, double c, int i, int j) { return f.p ? F(f, i, j) : c; }

typedef struct {
    G g;
    TP p;
    const csio_thermo_state *s;
    int i, j;
    double rho_i;
} cellctx;

/* melting_temperature(::LinearLiquidus, S): SeaIceThermodynamics.jl:58-60 */
static inline double melting_temperature(TP p, double S) { return p->liquidus_T0 - p->liquidus_slope * S; }

/* latent_heat: SeaIceThermodynamics.jl:158-167 */
static inline double latent_heat(TP p, double T)
{
    return p->reference_latent_heat + (p->liquid_density * p->liquid_heat_capacity / p->density - p->heat_capacity) * (T - p->reference_temperature);
}

/* bottom_temperature: HeatBoundaryConditions/bottom_heat_boundary_conditions.jl:32-38 */
static inline double bottom_temperature(const cellctx *c)
{
    if (c->p->bottom_bc == CSIO_BOTTOM_PRESCRIBED_TEMPERATURE) return opt(c->s->Tb, c->p->bottom_temperature, c->i, c->j);
    return melting_temperature(c->p, opt(c->s->Sb, c->p->bottom_salinity, c->i, c->j));
}

/* slab_internal_heat_flux(ConductiveFlux): slab_heat_and_tracer_fluxes.jl:8-31; reads fields.h */
static inline double ice_conductive_flux(const cellctx *c, double Tu)
{
    double Tb = bottom_temperature(c);
    double hi = F(c->s->h, c->i, c->j);
    return hi <= 0 ? 0.0 : (-c->p->ice_conductivity) * (Tu - Tb) / hi;
}
/* ice_snow_conductive_flux: slab_heat_and_tracer_fluxes.jl:49-67; reads fields.h, fields.hs */
static inline double ice_snow_conductive_flux(const cellctx *c, double Tu)
{
    double Tb = bottom_temperature(c);
    double hi = F(c->s->h, c->i, c->j), hs = F(c->s->hs, c->i, c->j);
    double R = hs / c->p->snow_conductivity + hi / c->p->ice_conductivity;
    return R <= 0 ? 0.0 : (Tb - Tu) / R;
}
/* interface_temperature: slab_heat_and_tracer_fluxes.jl:71-85 */
static inline double interface_temperature(const cellctx *c, double Tu)
{
    double Tb = bottom_temperature(c);
    double hi = F(c->s->h, c->i, c->j), hs = F(c->s->hs, c->i, c->j);
    double Ri = hi / c->p->ice_conductivity, Rs = hs / c->p->snow_conductivity, R = Rs + Ri;
    return R <= 0 ? Tb : Tb + (Tu - Tb) * Ri / R;
}

/* getflux of one term of the external top flux: HeatBoundaryConditions/boundary_fluxes.jl:8-12,139-144;
 * CONDUCTIVE = the default `internal_flux_function` wrapper of sea_ice_model.jl:244-252 */
static inline double top_term(const cellctx *c, int kind, double T)
{
    switch (kind) {
    case CSIO_FLUX_CONST: return c->p->top_flux_const;
    case CSIO_FLUX_ARRAY: return F(c->s->Qtop, c->i, c->j);
    case CSIO_FLUX_RADIATIVE_EMISSION: return c->p->emissivity * c->p->stefan_boltzmann * jl_pow4(T + c->p->emission_reference_temperature);
    case CSIO_FLUX_CONDUCTIVE: return ice_conductive_flux(c, T);
    case CSIO_FLUX_LINEAR: {
        double q = c->p->linear_coefficient * (T - c->p->linear_temperature);
        return c->p->linear_times_concentration ? q * F(c->s->a, c->i, c->j) : q;
    }
    default: return 0.0;
    }
}
/* getflux(::Tuple{A}) / (::Tuple{A,B}) = getflux(A) + getflux(B): boundary_fluxes.jl:15-17 */
static inline double top_external_flux(const cellctx *c, double T)
{
    double q = top_term(c, c->p->top_term_kind[0], T);
    if (c->p->n_top_terms > 1) q = q + top_term(c, c->p->top_term_kind[1], T);
    return q;
}
static inline double bottom_external_flux(const cellctx *c) { return opt(c->s->Qbot, c->p->bottom_flux_const, c->i, c->j); }

/* top_surface_temperature: HeatBoundaryConditions/top_heat_boundary_conditions.jl:82-100 with RootSolvers'
 * SecantMethod [RS-recall].  `combined` selects the internal flux: ice-only or snow+ice. */
static inline double flux_balance(const cellctx *c, int combined, double T)
{
    return top_external_flux(c, T) - (combined ? ice_snow_conductive_flux(c, T) : ice_conductive_flux(c, T));
}
static double secant_find_zero(const cellctx *c, int combined, double x0, double x1)
{
    double y0 = flux_balance(c, combined, x0);
    double y1 = flux_balance(c, combined, x1);
    for (int it = 1; it <= c->p->secant_maxiters; it++) {
        double dx = x1 - x0, dy = y1 - y0;
        x0 = x1;
        y0 = y1;
        x1 -= y1 * dx / dy;
        y1 = flux_balance(c, combined, x1);
        if (fabs(x0 - x1) < c->p->secant_tol) return x1;
    }
    return x1;
}
static inline double top_surface_temperature(const cellctx *c, int combined, double Tu)
{
    double T1 = Tu + 1, T2 = Tu - 0;
    return secant_find_zero(c, combined, T1, T2);
}

/* concentration_thermodynamic_step(::ProportionalEvolution): thermodynamic_time_step.jl:355-369 */
static inline double concentration_step(double dV, double an, double hn, double hc, double dt)
{
    int freezing = dV >= 0, melting = dV < 0;
    double daf = jl_mul_bool((1 - an) / hc * dV, freezing);
    double dam = jl_mul_bool(an / (2 * hn) * dV, melting);
    double ap = an + dt * (daf + dam);
    return jl_max(0.0, ap);
}
/* ice_volume_update: thermodynamic_time_step.jl:297-318 */
static inline void ice_volume_update(double dV, double hn, double an, double hc, double dt, double *h1, double *a1)
{
    double V = hn * an + dt * dV;
    V = jl_max(0.0, V);
    dV = (V - hn * an) / dt;
    double ap = concentration_step(dV, an, hn, hc, dt);
    double hp = V / ap;
    hp = ap <= 0 ? 0.0 : hp;
    ap = dV == 0 ? an : ap;
    hp = dV == 0 ? hn : hp;
    ap = hp == 0 ? 0.0 : ap;
    hp = ap == 0 ? 0.0 : hp;
    *a1 = ap > 1 ? 1.0 : ap;
    *h1 = ap > 1 ? hp * ap : hp;
}

/* ice_melt_freeze_tendency: slab_thermodynamics_tendencies.jl:30-69 (Qui, Qbi already scalars) */
static inline double ice_melt_freeze_tendency(const cellctx *c, double Tui, double Qui, double Qbi)
{
    double hi = F(c->s->h, c->i, c->j), hc = opt(c->s->hc, c->p->consolidation_thickness, c->i, c->j);
    int consolidated = hi >= hc;
    double Tbi = bottom_temperature(c);
    double Eb = c->rho_i * latent_heat(c->p, Tbi);
    double Eu = c->rho_i * latent_heat(c->p, Tui);
    double Qii = consolidated ? ice_conductive_flux(c, Tui) : 0.0; /* ice_interior_heat_flux: :12-21 */
    double wu = (Qui - Qii) / Eu;
    double wb = (Qii - Qbi) / Eb;
    return wu + wb;
}

/* thermodynamic_tendency: slab_thermodynamics_tendencies.jl:75-135 */
static inline double thermodynamic_tendency(const cellctx *c)
{
    int i = c->i, j = c->j;
    double hi = F(c->s->h, i, j), hc = opt(c->s->hc, c->p->consolidation_thickness, i, j);
    double Si = opt(c->s->S, c->p->ice_salinity, i, j);
    int consolidated = hi >= hc;
    if (c->p->top_bc != CSIO_TOP_PRESCRIBED_TEMPERATURE) {
        double Tun;
        if (consolidated) {
            Tun = top_surface_temperature(c, 0, F(c->s->Tu, i, j));
            Tun = jl_min(Tun, melting_temperature(c->p, Si));
        } else {
            Tun = bottom_temperature(c);
        }
        F(c->s->Tu, i, j) = Tun;
    }
    double Tui = F(c->s->Tu, i, j);
    double Qui = top_external_flux(c, Tui);
    double Qbi = bottom_external_flux(c);
    return ice_melt_freeze_tendency(c, Tui, Qui, Qbi);
}

/* snow_ice_formation: thermodynamic_time_step.jl:331-351 */
static inline void snow_ice_formation(double hi, double hs, double rho_i, double rho_s, double rho_w, double *hi1, double *hs1)
{
    double hf = hi * (1 - rho_i / rho_w) - hs * rho_s / rho_w;
    int flooding = hf < 0;
    double dhs = flooding ? -hf * rho_i / rho_s : 0.0;
    double hsp = jl_max(0.0, hs - dhs);
    dhs = hs - hsp;
    double dhi = dhs * rho_s / rho_i;
    *hi1 = hi + dhi;
    *hs1 = hsp;
}

/* _ice_thermodynamic_time_step!: thermodynamic_time_step.jl:76-118 */
static void ice_cell(const cellctx *c, double dt)
{
    int i = c->i, j = c->j;
    double hn = F(c->s->h, i, j), an = F(c->s->a, i, j), hc = opt(c->s->hc, c->p->consolidation_thickness, i, j);
    double dV = thermodynamic_tendency(c);
    double h1, a1;
    ice_volume_update(dV, hn, an, hc, dt, &h1, &a1);
    F(c->s->a, i, j) = a1;
    F(c->s->h, i, j) = h1;
    if (c->s->mf_ice.p) F(c->s->mf_ice, i, j) = c->rho_i * (h1 * a1 - hn * an) / dt;
    if (c->s->mf_snow.p) F(c->s->mf_snow, i, j) = 0.0;
    if (c->s->mf_snowfall.p) F(c->s->mf_snowfall, i, j) = 0.0;
}

/* _layered_thermodynamic_time_step!: thermodynamic_time_step.jl:132-291 */
static void layered_cell(const cellctx *c, double dt)
{
    int i = c->i, j = c->j;
    TP p = c->p;
    double hin = F(c->s->h, i, j), an = F(c->s->a, i, j), hc = opt(c->s->hc, p->consolidation_thickness, i, j);
    double hsn = F(c->s->hs, i, j);
    double Vin = hin * an, Vsn = hsn * an;
    int consolidated = hin >= hc;
    double Si = opt(c->s->S, p->ice_salinity, i, j);
    double Tb = bottom_temperature(c);
    double Tm = melting_temperature(p, Si);
    Tm = hsn > 0 ? 0.0 : Tm;
    if (p->snow_top_bc != CSIO_TOP_PRESCRIBED_TEMPERATURE) {
        double Tun;
        if (consolidated) {
            Tun = top_surface_temperature(c, 1, F(c->s->Tus, i, j));
            Tun = jl_min(Tun, Tm);
        } else {
            Tun = Tb;
        }
        F(c->s->Tus, i, j) = Tun;
    }
    double Tus = F(c->s->Tus, i, j);
    double Tsi = interface_temperature(c, Tus);
    F(c->s->Tu, i, j) = Tsi;
    double Qis = consolidated ? ice_snow_conductive_flux(c, Tus) : 0.0;
    double Qui = top_external_flux(c, Tus);
    double Qui_per_ice = an > 0 ? Qui / an : 0.0;
    double dQ = Qui_per_ice - Qis;
    double melt_energy = jl_max(0.0, -dQ);
    double rho_s = opt(c->s->rho_s, p->snow_density, i, j);
    double Ls = p->reference_latent_heat;
    double cap = rho_s * Ls * hsn / dt;
    double Qs = jl_min(melt_energy, cap);
    double Gsm = Qs / (rho_s * Ls);
    double rho_i = c->rho_i, rhoL = rho_i * Ls;
    double Qbi = bottom_external_flux(c);
    double alpha = (Qui - Qbi) / rhoL, beta = Qs / rhoL;
    double Cm = hin > 0 ? an / (2 * hin) : 0.0;
    double Cf = hc > 0 ? (1 - an) / hc : 0.0;
    double Km = dt * Cm, Kf = dt * Cf;
    double eps = 2.220446049250313e-16;
    double Dm = 1 - Km * beta, Df = 1 - Kf * beta;
    double am = fabs(Dm) > eps ? (an + Km * alpha) / Dm : an + Km * alpha;
    double af = fabs(Df) > eps ? (an + Kf * alpha) / Df : an + Kf * alpha;
    double dVm = alpha + beta * am;
    int melting = dVm < 0;
    double atmp = melting ? am : af;
    double Qeff = Qui + Qs * atmp;
    double dV = ice_melt_freeze_tendency(c, Tsi, Qeff, Qbi);
    double hi1, a1;
    ice_volume_update(dV, hin, an, hc, dt, &hi1, &a1);
    hsn = a1 > 0 ? hsn * an / a1 : 0.0;
    double Ps = opt(c->s->snowfall, p->snowfall, i, j);
    double Gsp = a1 > 0 ? Ps / rho_s : 0.0; /* snow_accumulation: :325-328 */
    double hsp = hsn + dt * (Gsp - Gsm);
    hsp = jl_max(0.0, hsp);
    snow_ice_formation(hi1, hsp, rho_i, rho_s, p->liquid_density, &hi1, &hsp);
    hsp = a1 <= 0 ? 0.0 : hsp;
    F(c->s->a, i, j) = a1;
    F(c->s->h, i, j) = hi1;
    F(c->s->hs, i, j) = hsp;
    double Pabs = rho_s * Gsp * a1;
    if (c->s->mf_ice.p) F(c->s->mf_ice, i, j) = rho_i * (hi1 * a1 - Vin) / dt;
    if (c->s->mf_snow.p) F(c->s->mf_snow, i, j) = rho_s * (hsp * a1 - Vsn) / dt - Pabs;
    if (c->s->mf_snowfall.p) F(c->s->mf_snowfall, i, j) = Pabs;
}

/* thermodynamic_time_step!: thermodynamic_time_step.jl:6-59 (launch over :xy) */
int csio_thermodynamic_time_step(const csio_grid *g, const csio_thermo_params *p, csio_thermo_state *s, double rho_ice, double dt)
{
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= g->Ny; j++)
        for (int i = 1; i <= g->Nx; i++) {
            cellctx c = {g, p, s, i, j, rho_ice};
            if (p->layered) layered_cell(&c, dt);
            else ice_cell(&c, dt);
        }
    return 0;
}

/* exposed for unit tests */
double csio_pow4(double x) { return jl_pow4(x); }

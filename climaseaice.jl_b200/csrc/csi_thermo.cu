// csi_thermo.cu -- slab thermodynamics step: one pointwise kernel per call.
//
//   thermodynamic_time_step!              src/SeaIceThermodynamics/thermodynamic_time_step.jl:6-59
//   _ice_thermodynamic_time_step!         :76-118      (bare ice)
//   _layered_thermodynamic_time_step!     :132-291     (snow + ice, resistors in series)
//   ice_volume_update, snow_ice_formation, concentration_thermodynamic_step   :297-369
//   thermodynamic_tendency, ice_melt_freeze_tendency   src/SeaIceThermodynamics/slab_thermodynamics_tendencies.jl:30-135
//   ConductiveFlux, IceSnowConductiveFlux, interface_temperature   src/SeaIceThermodynamics/slab_heat_and_tracer_fluxes.jl
//   latent_heat, melting_temperature      src/SeaIceThermodynamics/SeaIceThermodynamics.jl:58-60,158-167
//   top_surface_temperature (secant), bottom_temperature, getflux, RadiativeEmission
//                                         src/SeaIceThermodynamics/HeatBoundaryConditions/*.jl
//
// Every expression keeps the reference's left-to-right association (-fmad=false; `/` is the IEEE operator); the only
// fused operations are the two written in Julia's Float64^Int power (pow4).  One thread per column; HBM-bound:
// ~10 doubles per column for bare ice, ~14 with snow.
#include "csi_internal.h"
#include "csi_math.cuh"

namespace csi {

namespace {

__device__ __forceinline__ double jl_min(double a, double b)
{  // Base.min: NaN-propagating, min(-0.0, +0.0) = -0.0
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return signbit(a) ? a : b;
    return a < b ? a : b;
}

// x^4 as Base.pow_body(x::Float64, 4): two compensated squarings, then x + err
__device__ __forceinline__ double jl_pow4(double x)
{
    double xnlo = 0.0, err, hi;
    err = x * 2 * xnlo;
    hi = x * x;
    xnlo = __fma_rn(x, x, -hi);
    x = hi;
    xnlo += err;
    err = x * 2 * xnlo;
    hi = x * x;
    const double lo = __fma_rn(x, x, -hi);
    x = hi;
    xnlo = lo + err;
    err = 1.0 * xnlo + x * 0.0;
    return (isfinite(x) && isfinite(err)) ? x * 1.0 + err : x * 1.0;
}

struct Cell {
    const csi_thermo_config &p;
    const DThermoFields &f;
    int i, j;
    double rho_i;

    __device__ __forceinline__ double opt(const DArr &a, double c) const { return a.p ? at(a, i, j) : c; }
    __device__ __forceinline__ double melting_temperature(double S) const { return p.liquidus_freshwater_melting_temperature - p.liquidus_slope * S; }
    __device__ __forceinline__ double latent_heat(double T) const
    {
        return p.reference_latent_heat + (p.liquid_density * p.liquid_heat_capacity / p.density - p.heat_capacity) * (T - p.reference_temperature);
    }
    __device__ __forceinline__ double bottom_temperature() const
    {
        if (p.bottom_heat_bc == CSI_BOTTOM_PRESCRIBED_TEMPERATURE) return opt(f.Tb, p.bottom_temperature);
        return melting_temperature(opt(f.Sb, p.bottom_salinity));
    }
    __device__ __forceinline__ double ice_conductive_flux(double Tu) const
    {
        const double Tb = bottom_temperature();
        const double hi = at(f.h, i, j);
        return hi <= 0 ? 0.0 : (-p.ice_conductivity) * (Tu - Tb) / hi;
    }
    __device__ __forceinline__ double ice_snow_conductive_flux(double Tu) const
    {
        const double Tb = bottom_temperature();
        const double hi = at(f.h, i, j), hs = at(f.hs, i, j);
        const double R = hs / p.snow_conductivity + hi / p.ice_conductivity;
        return R <= 0 ? 0.0 : (Tb - Tu) / R;
    }
    __device__ __forceinline__ double interface_temperature(double Tu) const
    {
        const double Tb = bottom_temperature();
        const double hi = at(f.h, i, j), hs = at(f.hs, i, j);
        const double Ri = hi / p.ice_conductivity, Rs = hs / p.snow_conductivity, R = Rs + Ri;
        return R <= 0 ? Tb : Tb + (Tu - Tb) * Ri / R;
    }
    __device__ __forceinline__ double top_term(int kind, double T) const
    {
        switch (kind) {
        case CSI_FLUX_CONST: return p.top_flux_const;
        case CSI_FLUX_ARRAY: return at(f.Qtop, i, j);
        case CSI_FLUX_RADIATIVE_EMISSION: return p.emissivity * p.stefan_boltzmann_constant * jl_pow4(T + p.emission_reference_temperature);
        case CSI_FLUX_CONDUCTIVE: return ice_conductive_flux(T);
        case CSI_FLUX_LINEAR: {
            const double q = p.linear_coefficient * (T - p.linear_temperature);
            return p.linear_times_concentration ? q * at(f.a, i, j) : q;
        }
        default: return 0.0;
        }
    }
    __device__ __forceinline__ double top_external_flux(double T) const
    {
        double q = top_term(p.top_term_kind[0], T);
        if (p.n_top_terms > 1) q = q + top_term(p.top_term_kind[1], T);
        return q;
    }
    __device__ __forceinline__ double bottom_external_flux() const { return opt(f.Qbot, p.bottom_flux_const); }
    __device__ __forceinline__ double flux_balance(bool combined, double T) const
    {
        return top_external_flux(T) - (combined ? ice_snow_conductive_flux(T) : ice_conductive_flux(T));
    }
    // RootSolvers.find_zero(f, SecantMethod(T+1, T-0), CompactSolution()): SolutionTolerance, maxiters
    __device__ double top_surface_temperature(bool combined, double Tu) const
    {
        double x0 = Tu + 1, x1 = Tu - 0;
        double y0 = flux_balance(combined, x0);
        double y1 = flux_balance(combined, x1);
        for (int it = 1; it <= p.secant_maxiters; it++) {
            const double dx = x1 - x0, dy = y1 - y0;
            x0 = x1;
            y0 = y1;
            x1 -= y1 * dx / dy;
            y1 = flux_balance(combined, x1);
            if (fabs(x0 - x1) < p.secant_tolerance) return x1;
        }
        return x1;
    }
    __device__ __forceinline__ double ice_melt_freeze_tendency(double Tui, double Qui, double Qbi) const
    {
        const double hi = at(f.h, i, j), hc = opt(f.hc, p.ice_consolidation_thickness);
        const bool consolidated = hi >= hc;
        const double Tbi = bottom_temperature();
        const double Eb = rho_i * latent_heat(Tbi);
        const double Eu = rho_i * latent_heat(Tui);
        const double Qii = consolidated ? ice_conductive_flux(Tui) : 0.0;
        const double wu = (Qui - Qii) / Eu;
        const double wb = (Qii - Qbi) / Eb;
        return wu + wb;
    }
};

__device__ __forceinline__ double concentration_step(double dV, double an, double hn, double hc, double dt)
{
    const bool freezing = dV >= 0, melting = dV < 0;
    const double daf = jl_mul_bool((1 - an) / hc * dV, freezing);
    const double dam = jl_mul_bool(an / (2 * hn) * dV, melting);
    const double ap = an + dt * (daf + dam);
    return jl_max(0.0, ap);
}
__device__ __forceinline__ void ice_volume_update(double dV, double hn, double an, double hc, double dt, double &h1, double &a1)
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
    a1 = ap > 1 ? 1.0 : ap;
    h1 = ap > 1 ? hp * ap : hp;
}
__device__ __forceinline__ void snow_ice_formation(double hi, double hs, double rho_i, double rho_s, double rho_w, double &hi1, double &hs1)
{
    const double hf = hi * (1 - rho_i / rho_w) - hs * rho_s / rho_w;
    const bool flooding = hf < 0;
    double dhs = flooding ? -hf * rho_i / rho_s : 0.0;
    const double hsp = jl_max(0.0, hs - dhs);
    dhs = hs - hsp;
    const double dhi = dhs * rho_s / rho_i;
    hi1 = hi + dhi;
    hs1 = hsp;
}

__global__ void __launch_bounds__(256) k_thermodynamics(const __grid_constant__ DGrid g, const __grid_constant__ csi_thermo_config p,
                                                        const __grid_constant__ DThermoFields f, double rho_i, double dt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.Nx || j > g.Ny) return;
    const Cell c{p, f, i, j, rho_i};
    const double hn = at(f.h, i, j), an = at(f.a, i, j), hc = c.opt(f.hc, p.ice_consolidation_thickness);
    const double Si = c.opt(f.S, p.ice_salinity);
    const bool consolidated = hn >= hc;
    if (!p.layered) {
        // thermodynamic_tendency + _ice_thermodynamic_time_step!
        if (p.top_heat_bc != CSI_TOP_PRESCRIBED_TEMPERATURE) {
            double Tun;
            if (consolidated) {
                Tun = c.top_surface_temperature(false, at(f.Tu, i, j));
                Tun = jl_min(Tun, c.melting_temperature(Si));
            } else {
                Tun = c.bottom_temperature();
            }
            at(f.Tu, i, j) = Tun;
        }
        const double Tui = at(f.Tu, i, j);
        const double Qui = c.top_external_flux(Tui);
        const double Qbi = c.bottom_external_flux();
        const double dV = c.ice_melt_freeze_tendency(Tui, Qui, Qbi);
        double h1, a1;
        ice_volume_update(dV, hn, an, hc, dt, h1, a1);
        at(f.a, i, j) = a1;
        at(f.h, i, j) = h1;
        if (f.mf_ice.p) at(f.mf_ice, i, j) = rho_i * (h1 * a1 - hn * an) / dt;
        if (f.mf_snow.p) at(f.mf_snow, i, j) = 0.0;
        if (f.mf_snowfall.p) at(f.mf_snowfall, i, j) = 0.0;
        return;
    }
    // _layered_thermodynamic_time_step!
    double hsn = at(f.hs, i, j);
    const double Vin = hn * an, Vsn = hsn * an;
    const double Tb = c.bottom_temperature();
    double Tm = c.melting_temperature(Si);
    Tm = hsn > 0 ? 0.0 : Tm;
    if (p.snow_top_heat_bc != CSI_TOP_PRESCRIBED_TEMPERATURE) {
        double Tun;
        if (consolidated) {
            Tun = c.top_surface_temperature(true, at(f.Tus, i, j));
            Tun = jl_min(Tun, Tm);
        } else {
            Tun = Tb;
        }
        at(f.Tus, i, j) = Tun;
    }
    const double Tus = at(f.Tus, i, j);
    const double Tsi = c.interface_temperature(Tus);
    at(f.Tu, i, j) = Tsi;
    const double Qis = consolidated ? c.ice_snow_conductive_flux(Tus) : 0.0;
    const double Qui = c.top_external_flux(Tus);
    const double Qui_per_ice = an > 0 ? Qui / an : 0.0;
    const double dQ = Qui_per_ice - Qis;
    const double melt_energy = jl_max(0.0, -dQ);
    const double rho_s = c.opt(f.rho_s, p.snow_density);
    const double Ls = p.reference_latent_heat;
    const double cap = rho_s * Ls * hsn / dt;
    const double Qs = jl_min(melt_energy, cap);
    const double Gsm = Qs / (rho_s * Ls);
    const double rhoL = rho_i * Ls;
    const double Qbi = c.bottom_external_flux();
    const double alpha = (Qui - Qbi) / rhoL, beta = Qs / rhoL;
    const double Cm = hn > 0 ? an / (2 * hn) : 0.0;
    const double Cf = hc > 0 ? (1 - an) / hc : 0.0;
    const double Km = dt * Cm, Kf = dt * Cf;
    const double eps = 2.220446049250313e-16;
    const double Dm = 1 - Km * beta, Df = 1 - Kf * beta;
    const double am = fabs(Dm) > eps ? (an + Km * alpha) / Dm : an + Km * alpha;
    const double af = fabs(Df) > eps ? (an + Kf * alpha) / Df : an + Kf * alpha;
    const double dVm = alpha + beta * am;
    const bool melting = dVm < 0;
    const double atmp = melting ? am : af;
    const double Qeff = Qui + Qs * atmp;
    const double dV = c.ice_melt_freeze_tendency(Tsi, Qeff, Qbi);
    double hi1, a1;
    ice_volume_update(dV, hn, an, hc, dt, hi1, a1);
    hsn = a1 > 0 ? hsn * an / a1 : 0.0;
    const double Ps = c.opt(f.snowfall, p.snowfall);
    const double Gsp = a1 > 0 ? Ps / rho_s : 0.0;
    double hsp = hsn + dt * (Gsp - Gsm);
    hsp = jl_max(0.0, hsp);
    snow_ice_formation(hi1, hsp, rho_i, rho_s, p.liquid_density, hi1, hsp);
    hsp = a1 <= 0 ? 0.0 : hsp;
    at(f.a, i, j) = a1;
    at(f.h, i, j) = hi1;
    at(f.hs, i, j) = hsp;
    const double Pabs = rho_s * Gsp * a1;
    if (f.mf_ice.p) at(f.mf_ice, i, j) = rho_i * (hi1 * a1 - Vin) / dt;
    if (f.mf_snow.p) at(f.mf_snow, i, j) = rho_s * (hsp * a1 - Vsn) / dt - Pabs;
    if (f.mf_snowfall.p) at(f.mf_snowfall, i, j) = Pabs;
}

}  // namespace

void launch_thermodynamics(const LaunchCtx &c, const DGrid &g, const csi_thermo_config &p, const DThermoFields &f, double rho_i, double dt)
{
    k_thermodynamics<<<dim3((g.Nx + 255) / 256, g.Ny), 256, 0, c.stream>>>(g, p, f, rho_i, dt);
    ++*c.launches;
}

}  // namespace csi

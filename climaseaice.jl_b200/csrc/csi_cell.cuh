// csi_cell.cuh -- per-node formulas of the EVP substep, evaluated straight from global memory.
//
// Used by the UNFUSED kernels (csi_unfused.cu), which keep the reference's launch structure
// (stress kernel, u kernel, v kernel, halo fills) and serve as the general path for every
// topology.  Each function names the reference function it stands for
// (evp = src/Rheologies/elasto_visco_plastic_rheology.jl, isd = src/Rheologies/ice_stress_divergence.jl,
//  mt = src/SeaIceDynamics/momentum_tendencies_kernel_functions.jl,
//  ext = src/SeaIceDynamics/sea_ice_external_stress.jl, se = src/SeaIceDynamics/split_explicit_momentum_equations.jl).
// Expression trees keep Julia's left-to-right association; compiled with -fmad=false.
#pragma once
#include "../../include/climaseaice_b200.h"
#include "csi_math.cuh"
#include "csi_types.cuh"

namespace csi {

// ---- node activity (Oceananigans inactive_cell / peripheral_node) ----------------------------
__device__ __forceinline__ bool outside_domain(const DGrid &g, int i, int j)
{
    // a partition's connected side is a rank boundary, not a wall
    return (g.topo_x == CSI_BOUNDED && ((i < 1 && !g.conn_w) || (i > g.Nx && !g.conn_e))) ||
           (g.topo_y == CSI_BOUNDED && ((j < 1 && !g.conn_s) || (j > g.Ny && !g.conn_n)));
}
__device__ __forceinline__ bool immersed_cell(const DGrid &g, int i, int j)
{
    if (!g.mask) return false;
    const int sx = g.Nx + 2 * g.Hx, sy = g.Ny + 2 * g.Hy;
    int pi = min(max(i - 1 + g.Hx, 0), sx - 1), pj = min(max(j - 1 + g.Hy, 0), sy - 1);
    return g.mask[(size_t)pi + (size_t)pj * sx] != 0;
}
__device__ __forceinline__ bool inactive_cell(const DGrid &g, int i, int j) { return outside_domain(g, i, j) || immersed_cell(g, i, j); }
__device__ __forceinline__ bool peripheral_fc(const DGrid &g, int i, int j) { return inactive_cell(g, i - 1, j) || inactive_cell(g, i, j); }
__device__ __forceinline__ bool peripheral_cf(const DGrid &g, int i, int j) { return inactive_cell(g, i, j - 1) || inactive_cell(g, i, j); }
__device__ __forceinline__ bool imm_peripheral_cc(const DGrid &g, int i, int j) { return inactive_cell(g, i, j) && !outside_domain(g, i, j); }
__device__ __forceinline__ bool imm_peripheral_ff(const DGrid &g, int i, int j)
{
    const bool per = inactive_cell(g, i - 1, j - 1) || inactive_cell(g, i, j - 1) || inactive_cell(g, i - 1, j) || inactive_cell(g, i, j);
    const bool und = outside_domain(g, i - 1, j - 1) || outside_domain(g, i, j - 1) || outside_domain(g, i - 1, j) || outside_domain(g, i, j);
    return per && !und;
}
__device__ __forceinline__ bool imm_peripheral_fc(const DGrid &g, int i, int j)
{
    return peripheral_fc(g, i, j) && !(outside_domain(g, i - 1, j) || outside_domain(g, i, j));
}
__device__ __forceinline__ bool imm_peripheral_cf(const DGrid &g, int i, int j)
{
    return peripheral_cf(g, i, j) && !(outside_domain(g, i, j - 1) || outside_domain(g, i, j));
}

// ---- averages (Oceananigans: y-average of x-averages) -----------------------------------------
template <class Fn> __device__ __forceinline__ double avg_ff(Fn q, int i, int j) { return ((q(i - 1, j - 1) + q(i, j - 1)) / 2 + (q(i - 1, j) + q(i, j)) / 2) / 2; }
template <class Fn> __device__ __forceinline__ double avg_cc(Fn q, int i, int j) { return ((q(i, j) + q(i + 1, j)) / 2 + (q(i, j + 1) + q(i + 1, j + 1)) / 2) / 2; }
template <class Fn> __device__ __forceinline__ double avg_fc(Fn q, int i, int j) { return ((q(i - 1, j) + q(i, j)) / 2 + (q(i - 1, j + 1) + q(i, j + 1)) / 2) / 2; }
template <class Fn> __device__ __forceinline__ double avg_cf(Fn q, int i, int j) { return ((q(i, j - 1) + q(i + 1, j - 1)) / 2 + (q(i, j) + q(i + 1, j)) / 2) / 2; }

// ---- strain rates: evp:360-375 (metrics: scalars on a RectilinearGrid, functions of j on a lat-lon grid) ----
__device__ __forceinline__ double eps_D(const DGrid &g, const DArr &u, const DArr &v, int i, int j)
{
    return ((dyfc(g, i + 1, j) * at(u, i + 1, j) - dyfc(g, i, j) * at(u, i, j)) + (dxcf(g, i, j + 1) * at(v, i, j + 1) - dxcf(g, i, j) * at(v, i, j))) / azcc(g, i, j);
}
__device__ __forceinline__ double eps_T(const DGrid &g, const DArr &u, const DArr &v, int i, int j)
{
    const double dy = dycc(g, i, j), dx = dxcc(g, i, j);
    return (dy * dy * (at(u, i + 1, j) / dyfc(g, i + 1, j) - at(u, i, j) / dyfc(g, i, j)) - dx * dx * (at(v, i, j + 1) / dxcf(g, i, j + 1) - at(v, i, j) / dxcf(g, i, j))) /
           azcc(g, i, j);
}
__device__ __forceinline__ double eps_S(const DGrid &g, const DArr &u, const DArr &v, int i, int j)
{
    const double dx = dxff(g, i, j), dy = dyff(g, i, j);
    return (dx * dx * (at(u, i, j) / dxfc(g, i, j) - at(u, i, j - 1) / dxfc(g, i, j - 1)) + dy * dy * (at(v, i, j) / dycf(g, i, j) - at(v, i - 1, j) / dycf(g, i - 1, j))) /
           azff(g, i, j);
}
__device__ __forceinline__ double strain_xx(const DGrid &g, const DArr &u, const DArr &v, int i, int j) { return (eps_D(g, u, v, i, j) + eps_T(g, u, v, i, j)) / 2; }
__device__ __forceinline__ double strain_yy(const DGrid &g, const DArr &u, const DArr &v, int i, int j) { return (eps_D(g, u, v, i, j) - eps_T(g, u, v, i, j)) / 2; }
__device__ __forceinline__ double strain_xy(const DGrid &g, const DArr &u, const DArr &v, int i, int j) { return eps_S(g, u, v, i, j) / 2; }

// ice_mass: src/ClimaSeaIce.jl:42
__device__ __forceinline__ double ice_mass(const DParams &p, const DFields &f, int i, int j) { return at(f.h, i, j) * p.rho_i * at(f.a, i, j); }

// ---- _compute_evp_viscosities! + _compute_evp_stresses! at one node: evp:236-354 --------------
__device__ __forceinline__ void evp_stress_node(const DGrid &g, const DParams &p, const DFields &f, double dt, int i, int j)
{
    const DArr &u = f.u, &v = f.v;
    auto exx = [&](int a, int b) { return strain_xx(g, u, v, a, b); };
    auto eyy = [&](int a, int b) { return strain_yy(g, u, v, a, b); };
    auto exy = [&](int a, int b) { return strain_xy(g, u, v, a, b); };
    auto PP = [&](int a, int b) { return at(f.P, a, b); };
    auto mm = [&](int a, int b) { return ice_mass(p, f, a, b); };

    const double e11c = exx(i, j), e22c = eyy(i, j), e12f = exy(i, j);
    const double e11f = avg_ff(exx, i, j), e22f = avg_ff(eyy, i, j), e12c = avg_cc(exy, i, j);
    const double dc = e11c + e22c, df = e11f + e22f;
    const double sc = sqrt((e11c - e22c) * (e11c - e22c) + 4 * (e12c * e12c));
    const double sf = sqrt((e11f - e22f) * (e11f - e22f) + 4 * (e12f * e12f));
    const double Dc = jl_max(sqrt(dc * dc + sc * sc * p.em2), p.Dmin);
    const double Df = jl_max(sqrt(df * df + sf * sf * p.em2), p.Dmin);
    const double Pc = PP(i, j), Pf = avg_ff(PP, i, j);
    const double zf = Pf / (2 * Df), zc = Pc / (2 * Dc);
    at(f.zf, i, j) = zf;
    at(f.zc, i, j) = zc;
    at(f.delta, i, j) = Dc;

    // stresses (the strain rates of evp:310-312 are the same expressions as e11c, e22c, e12f)
    const double Pr = p.pform == CSI_ICE_STRENGTH ? Pc : Pc * Dc / (Dc + p.Dmin);
    const double ec = zc * p.em2, ef = zf * p.em2;
    const double s11n = 2 * ec * e11c + ((zc - ec) * (e11c + e22c) - Pr / 2);
    const double s22n = 2 * ec * e22c + ((zc - ec) * (e11c + e22c) - Pr / 2);
    const double s12n = 2 * ef * e12f;
    const double mc = mm(i, j), mf = avg_ff(mm, i, j);
    double g2c = zc * p.ca * dt / mc / azcc(g, i, j);
    g2c = (g2c != g2c) ? p.amax * p.amax : g2c;
    const double gc = jl_clamp(sqrt(g2c), p.amin, p.amax);
    double g2f = zf * p.ca * dt / mf / azff(g, i, j);
    g2f = (g2f != g2f) ? p.amax * p.amax : g2f;
    const double gf = jl_clamp(sqrt(g2f), p.amin, p.amax);
    const double d11 = (s11n - at(f.s11, i, j)) / gc;
    const double d22 = (s22n - at(f.s22, i, j)) / gc;
    const double d12 = (s12n - at(f.s12, i, j)) / gf;
    at(f.s11, i, j) += (mc > 0 ? d11 : 0.0);
    at(f.s22, i, j) += (mc > 0 ? d22 : 0.0);
    at(f.s12, i, j) += (mf > 0 ? d12 : 0.0);
    at(f.alpha, i, j) = gc;
}

// ---- stress divergence: isd:16-51 ------------------------------------------------------------
__device__ __forceinline__ double stress_cc(const DGrid &g, const DArr &s, int i, int j) { return (g.mask && imm_peripheral_cc(g, i, j)) ? 0.0 : at(s, i, j); }
__device__ __forceinline__ double stress_ff(const DGrid &g, const DArr &s, int i, int j) { return (g.mask && imm_peripheral_ff(g, i, j)) ? 0.0 : at(s, i, j); }
__device__ __forceinline__ double sigD(const DGrid &g, const DFields &f, int i, int j) { return stress_cc(g, f.s11, i, j) + stress_cc(g, f.s22, i, j); }
__device__ __forceinline__ double sigT(const DGrid &g, const DFields &f, int i, int j) { return stress_cc(g, f.s11, i, j) - stress_cc(g, f.s22, i, j); }
__device__ __forceinline__ double div_sigma_1j(const DGrid &g, const DFields &f, int i, int j)
{
    const double d = dyfc(g, i, j) * (sigD(g, f, i, j) - sigD(g, f, i - 1, j)) / 2;
    const double t = (dycc(g, i, j) * dycc(g, i, j) * sigT(g, f, i, j) - dycc(g, i - 1, j) * dycc(g, i - 1, j) * sigT(g, f, i - 1, j)) / dyfc(g, i, j) / 2;
    const double S = (dxff(g, i, j + 1) * dxff(g, i, j + 1) * stress_ff(g, f.s12, i, j + 1) - dxff(g, i, j) * dxff(g, i, j) * stress_ff(g, f.s12, i, j)) / dxfc(g, i, j);
    return (d + t + S) / azfc(g, i, j);
}
__device__ __forceinline__ double div_sigma_2j(const DGrid &g, const DFields &f, int i, int j)
{
    const double d = dxcf(g, i, j) * (sigD(g, f, i, j) - sigD(g, f, i, j - 1)) / 2;
    const double t = -(dxcc(g, i, j) * dxcc(g, i, j) * sigT(g, f, i, j) - dxcc(g, i, j - 1) * dxcc(g, i, j - 1) * sigT(g, f, i, j - 1)) / dxcf(g, i, j) / 2;
    const double S = (dyff(g, i + 1, j) * dyff(g, i + 1, j) * stress_ff(g, f.s12, i + 1, j) - dyff(g, i, j) * dyff(g, i, j) * stress_ff(g, f.s12, i, j)) / dycf(g, i, j);
    return (d + t + S) / azcf(g, i, j);
}

// ---- immersed stress divergence: isd:57-123 with the linear-drag FluxBoundaryCondition -C*u (coastline example) ----
__device__ __forceinline__ double immersed_div_sigma_1j(const DGrid &g, const DParams &p, const DFields &f, int i, int j)
{
    if (!g.mask || p.imm_u == 0.0) return 0.0;
    const double bc = (-p.imm_u) * at(f.u, i, j);
    const double qW = 0.0 * (dycc(g, i - 1, j) * 1.0), qE = 0.0 * (dycc(g, i, j) * 1.0);
    const double qS = (imm_peripheral_ff(g, i, j) ? -bc : 0.0) * (dxff(g, i, j) * 1.0);
    const double qN = (imm_peripheral_ff(g, i, j + 1) ? bc : 0.0) * (dxff(g, i, j + 1) * 1.0);
    return (qE - qW + qN - qS) / (azfc(g, i, j) * 1.0);
}
__device__ __forceinline__ double immersed_div_sigma_2j(const DGrid &g, const DParams &p, const DFields &f, int i, int j)
{
    if (!g.mask || p.imm_v == 0.0) return 0.0;
    const double bc = (-p.imm_v) * at(f.v, i, j);
    const double qW = (imm_peripheral_ff(g, i, j) ? -bc : 0.0) * (dyff(g, i, j) * 1.0);
    const double qE = (imm_peripheral_ff(g, i + 1, j) ? bc : 0.0) * (dyff(g, i + 1, j) * 1.0);
    const double qS = 0.0 * (dxcc(g, i, j - 1) * 1.0), qN = 0.0 * (dxcc(g, i, j) * 1.0);
    return (qE - qW + qN - qS) / (azcf(g, i, j) * 1.0);
}

// ---- external stresses: ext:8-40,176-210 -----------------------------------------------------------
// Either side (TOP: atmosphere, BOT: ocean) is nothing / a Number pair / a pair of arrays / a SemiImplicitStress.
// ext_x/ext_y is the side's array-or-constant: the stress itself for CONST and FIELD, u_e / v_e for SEMI_IMPLICIT.
enum { TOP = 0, BOT = 1 };
__device__ __forceinline__ int stress_kind(const DParams &p, int w) { return w == TOP ? p.top_kind : p.bot_kind; }
__device__ __forceinline__ double ext_x(const DParams &p, const DFields &f, int w, int i, int j)
{
    if (w == TOP) return f.top_x.p ? at(f.top_x, i, j) : p.ttx;
    return f.ue.p ? at(f.ue, i, j) : p.ue_c;
}
__device__ __forceinline__ double ext_y(const DParams &p, const DFields &f, int w, int i, int j)
{
    if (w == TOP) return f.top_y.p ? at(f.top_y, i, j) : p.tty;
    return f.ve.p ? at(f.ve, i, j) : p.ve_c;
}
__device__ __forceinline__ double ext_rho(const DParams &p, int w) { return w == TOP ? p.top_rho : p.rho_e; }
__device__ __forceinline__ double ext_Cd(const DParams &p, int w) { return w == TOP ? p.top_Cd : p.Cd; }
__device__ __forceinline__ double sis_speed_x(const DParams &p, const DFields &f, int w, int i, int j)
{
    auto ey = [&](int a, int b) { return ext_y(p, f, w, a, b); };
    auto vv = [&](int a, int b) { return at(f.v, a, b); };
    const double du = ext_x(p, f, w, i, j) - at(f.u, i, j);
    const double dv = avg_fc(ey, i, j) - avg_fc(vv, i, j);
    return sqrt(du * du + dv * dv);
}
__device__ __forceinline__ double sis_speed_y(const DParams &p, const DFields &f, int w, int i, int j)
{
    auto ex = [&](int a, int b) { return ext_x(p, f, w, a, b); };
    auto uu = [&](int a, int b) { return at(f.u, a, b); };
    const double dv = ext_y(p, f, w, i, j) - at(f.v, i, j);
    const double du = avg_cf(ex, i, j) - avg_cf(uu, i, j);
    return sqrt(du * du + dv * dv);
}
// implicit_tau_x/y_coefficient (zero unless SemiImplicitStress) and explicit_tau_x/y
__device__ __forceinline__ double implicit_tx(const DParams &p, const DFields &f, int w, int i, int j)
{
    return stress_kind(p, w) == CSI_STRESS_SEMI_IMPLICIT ? ext_rho(p, w) * ext_Cd(p, w) * sis_speed_x(p, f, w, i, j) : 0.0;
}
__device__ __forceinline__ double implicit_ty(const DParams &p, const DFields &f, int w, int i, int j)
{
    return stress_kind(p, w) == CSI_STRESS_SEMI_IMPLICIT ? ext_rho(p, w) * ext_Cd(p, w) * sis_speed_y(p, f, w, i, j) : 0.0;
}
__device__ __forceinline__ double explicit_tx(const DParams &p, const DFields &f, int w, int i, int j, double coef)
{
    const int k = stress_kind(p, w);
    if (k == CSI_STRESS_NONE) return 0.0;
    if (k == CSI_STRESS_SEMI_IMPLICIT) return coef * ext_x(p, f, w, i, j);  // (rho * Cd * sqrt(..)) * u_e, coef = implicit_tx
    return ext_x(p, f, w, i, j);
}
__device__ __forceinline__ double explicit_ty(const DParams &p, const DFields &f, int w, int i, int j, double coef)
{
    const int k = stress_kind(p, w);
    if (k == CSI_STRESS_NONE) return 0.0;
    if (k == CSI_STRESS_SEMI_IMPLICIT) return coef * ext_y(p, f, w, i, j);
    return ext_y(p, f, w, i, j);
}
// x/y_momentum_stress of a side that is not a SemiImplicitStress (ext:34-38): explicit - zero(grid) * u
__device__ __forceinline__ double x_momentum_stress(const DParams &p, const DFields &f, int w, int i, int j)
{
    return explicit_tx(p, f, w, i, j, 0.0) - 0.0 * at(f.u, i, j);
}
__device__ __forceinline__ double y_momentum_stress(const DParams &p, const DFields &f, int w, int i, int j)
{
    return explicit_ty(p, f, w, i, j, 0.0) - 0.0 * at(f.v, i, j);
}

// ---- free drift: stress_balance_free_drift.jl:61-129 ---------------------------------------------
// nothing -> 0; (u=, v=) arrays -> the array value; StressBalanceFreeDrift on the model's own stresses:
// with d the SemiImplicitStress side and o the other, U_d - tau_o / sqrt(C_d * |tau_o|)
__device__ __forceinline__ double free_drift_u(const DParams &p, const DFields &f, int i, int j)
{
    if (p.fd_kind == CSI_FD_NONE) return 0.0;
    if (p.fd_kind == CSI_FD_FIELDS) return at(f.fd_u, i, j);
    const int d = p.bot_kind == CSI_STRESS_SEMI_IMPLICIT ? BOT : TOP, o = 1 - d;
    auto yms = [&](int a, int b) { return y_momentum_stress(p, f, o, a, b); };
    const double tx = x_momentum_stress(p, f, o, i, j);
    const double ty = avg_fc(yms, i, j);
    const double t = sqrt(tx * tx + ty * ty);
    const double Ud = ext_x(p, f, d, i, j);
    const double Cdrag = ext_rho(p, d) * ext_Cd(p, d);
    return Ud - (t == 0 ? t : tx / sqrt(Cdrag * t));
}
__device__ __forceinline__ double free_drift_v(const DParams &p, const DFields &f, int i, int j)
{
    if (p.fd_kind == CSI_FD_NONE) return 0.0;
    if (p.fd_kind == CSI_FD_FIELDS) return at(f.fd_v, i, j);
    const int d = p.bot_kind == CSI_STRESS_SEMI_IMPLICIT ? BOT : TOP, o = 1 - d;
    auto xms = [&](int a, int b) { return x_momentum_stress(p, f, o, a, b); };
    const double tx = avg_cf(xms, i, j);
    const double ty = y_momentum_stress(p, f, o, i, j);
    const double t = sqrt(tx * tx + ty * ty);
    const double Ud = ext_y(p, f, d, i, j);
    const double Cdrag = ext_rho(p, d) * ext_Cd(p, d);
    return Ud - (t == 0 ? t : ty / sqrt(Cdrag * t));
}

// ---- Coriolis [OCN-recall]: FPlane, or HydrostaticSphericalCoriolis with the EnstrophyConserving scheme ------
//   x_f_cross_U = -Iy^c(f^ff) * Ix^f(Iy^c(dx^cf v)) / dx^fc ;  y_f_cross_U = +Ix^c(f^ff) * Iy^f(Ix^c(dy^fc u)) / dy^cf
__device__ __forceinline__ double x_f_cross_U(const DGrid &g, const DParams &p, const DFields &f, int i, int j)
{
    if (p.cor == CSI_CORIOLIS_NONE) return 0.0;
    if (p.cor == CSI_CORIOLIS_SPHERICAL) {
        auto dxv = [&](int a, int b) { return (dxcf(g, a, b) * at(f.v, a, b) + dxcf(g, a, b + 1) * at(f.v, a, b + 1)) / 2; };
        const double fbar = (__ldg(p.fff + (j - 1 + g.Hy)) + __ldg(p.fff + (j + g.Hy))) / 2;
        return -fbar * ((dxv(i - 1, j) + dxv(i, j)) / 2) / dxfc(g, i, j);
    }
    auto vv = [&](int a, int b) { return at(f.v, a, b); };
    return -p.f * avg_fc(vv, i, j);
}
__device__ __forceinline__ double y_f_cross_U(const DGrid &g, const DParams &p, const DFields &f, int i, int j)
{
    if (p.cor == CSI_CORIOLIS_NONE) return 0.0;
    if (p.cor == CSI_CORIOLIS_SPHERICAL) {
        auto dyu = [&](int a, int b) { return (dyfc(g, a, b) * at(f.u, a, b) + dyfc(g, a + 1, b) * at(f.u, a + 1, b)) / 2; };
        const double fj = __ldg(p.fff + (j - 1 + g.Hy));
        const double fbar = (fj + fj) / 2;
        return fbar * ((dyu(i, j - 1) + dyu(i, j)) / 2) / dycf(g, i, j);
    }
    auto uu = [&](int a, int b) { return at(f.u, a, b); };
    return p.f * avg_cf(uu, i, j);
}

// ---- _u_velocity_step! at one face: se:197-229 with mt:11-41, evp:384,391-395 -------------------
__device__ __forceinline__ void u_step_node(const DGrid &g, const DParams &p, const DFields &f, double dt, int i, int j)
{
    auto mm = [&](int a, int b) { return ice_mass(p, f, a, b); };
    const double mi = (mm(i, j) + mm(i - 1, j)) / 2;
    const double ai = (at(f.a, i, j) + at(f.a, i - 1, j)) / 2;
    const double abar = (at(f.alpha, i, j) + at(f.alpha, i - 1, j)) / 2;
    const double dtau = dt / abar;
    const double cbot = implicit_tx(p, f, BOT, i, j), ctop = implicit_tx(p, f, TOP, i, j);  // implicit_tx_coefficient
    const double tbot = explicit_tx(p, f, BOT, i, j, cbot), ttop = explicit_tx(p, f, TOP, i, j, ctop);
    const double xcross = x_f_cross_U(g, p, f, i, j);
    const double rheo = (at(f.un, i, j) - at(f.u, i, j)) / dtau / abar;
    double Gu = -xcross - ttop / mi * ai + tbot / mi * ai + div_sigma_1j(g, f, i, j) / mi + immersed_div_sigma_1j(g, p, f, i, j) / mi + (0.0 + rheo);
    Gu = mi <= 0 ? 0.0 : Gu;
    double tau = (cbot - ctop) / mi * ai;
    tau = mi <= 0 ? 0.0 : tau;
    const double uD = (at(f.u, i, j) + dtau * Gu) / (1 + dtau * tau);
    const double uF = free_drift_u(p, f, i, j);
    const bool marginal = (mi > 2.220446049250313e-16) & (ai > 2.220446049250313e-16);
    const bool active_ice = (mi >= p.min_mass) & (ai >= p.min_conc);
    const bool active = !peripheral_fc(g, i, j);
    at(f.u, i, j) = jl_mul_bool(active_ice ? uD : (marginal ? uF : 0.0), active);
}

// ---- _v_velocity_step! at one face: se:231-264 with mt:44-74, evp:385,397-401 -------------------
__device__ __forceinline__ void v_step_node(const DGrid &g, const DParams &p, const DFields &f, double dt, int i, int j)
{
    auto mm = [&](int a, int b) { return ice_mass(p, f, a, b); };
    const double mi = (mm(i, j) + mm(i, j - 1)) / 2;
    const double ai = (at(f.a, i, j) + at(f.a, i, j - 1)) / 2;
    const double abar = (at(f.alpha, i, j) + at(f.alpha, i, j - 1)) / 2;
    const double dtau = dt / abar;
    const double cbot = implicit_ty(p, f, BOT, i, j), ctop = implicit_ty(p, f, TOP, i, j);
    const double tbot = explicit_ty(p, f, BOT, i, j, cbot), ttop = explicit_ty(p, f, TOP, i, j, ctop);
    const double ycross = y_f_cross_U(g, p, f, i, j);
    const double rheo = (at(f.vn, i, j) - at(f.v, i, j)) / dtau / abar;
    double Gv = -ycross - ttop / mi * ai + tbot / mi * ai + div_sigma_2j(g, f, i, j) / mi + immersed_div_sigma_2j(g, p, f, i, j) / mi + (0.0 + rheo);
    Gv = mi <= 0 ? 0.0 : Gv;
    double tau = (cbot - ctop) / mi * ai;
    tau = mi <= 0 ? 0.0 : tau;
    const double vD = (at(f.v, i, j) + dtau * Gv) / (1 + dtau * tau);
    const double vF = free_drift_v(p, f, i, j);
    const bool marginal = (mi > 2.220446049250313e-16) & (ai > 2.220446049250313e-16);
    const bool active_ice = (mi >= p.min_mass) & (ai >= p.min_conc);
    const bool active = !peripheral_cf(g, i, j);
    at(f.v, i, j) = jl_mul_bool(active_ice ? vD : (marginal ? vF : 0.0), active);
}

}  // namespace csi

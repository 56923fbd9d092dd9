// csi_types.cuh -- device-side views of the grid, parameters and fields (passed by value).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace csi {

// Oceananigans parent array: element (i, j) (1-based) at p[(i-1+ox) + (j-1+oy)*sx]
struct DArr {
    double *p;
    int sx, sy, ox, oy;
};
__device__ __forceinline__ double &at(const DArr &a, int i, int j)
{
    return a.p[(size_t)(i - 1 + a.ox) + (size_t)(j - 1 + a.oy) * (size_t)a.sx];
}
__device__ __forceinline__ double ld(const DArr &a, int i, int j)
{
    return __ldg(a.p + ((size_t)(i - 1 + a.ox) + (size_t)(j - 1 + a.oy) * (size_t)a.sx));
}

struct DGrid {
    int Nx, Ny, Hx, Hy;
    int topo_x, topo_y;     // CSI_PERIODIC / CSI_BOUNDED
    int conn_s, conn_n;     // partition: south / north side is a rank boundary (halo exchanged)
    int conn_w, conn_e;     // 2-D partition: west / east side is a rank boundary
    double dx, dy, az;      // regular rectilinear metrics; az = dx*dy
    const uint8_t *mask;    // optional immersed mask at centres (parent-shaped), device pointer
    const uint8_t *mask_host;  // the same mask in host memory (plan construction only)
    // j-dependent metrics (LatitudeLongitudeGrid): 12 arrays of metL doubles, entry for index j at [j-1+Hy];
    // order dxcc dxfc dxcf dxff dycc dyfc dycf dyff azcc azfc azcf azff.  NULL on a regular RectilinearGrid.
    const double *met;
    int metL, pad_;
    const double *met_host;  // host copies (plan construction only): the 12 x metL metrics, and f at (Face, Face) or NULL
    const double *fff_host;
};

// grid metrics at row j (they do not depend on i on the supported grids)
enum { M_DXCC = 0, M_DXFC, M_DXCF, M_DXFF, M_DYCC, M_DYFC, M_DYCF, M_DYFF, M_AZCC, M_AZFC, M_AZCF, M_AZFF };
__device__ __forceinline__ double metric(const DGrid &g, int which, int j, double regular)
{
    return g.met ? __ldg(g.met + (size_t)which * g.metL + (j - 1 + g.Hy)) : regular;
}
__device__ __forceinline__ double dxcc(const DGrid &g, int j) { return metric(g, M_DXCC, j, g.dx); }
__device__ __forceinline__ double dxfc(const DGrid &g, int j) { return metric(g, M_DXFC, j, g.dx); }
__device__ __forceinline__ double dxcf(const DGrid &g, int j) { return metric(g, M_DXCF, j, g.dx); }
__device__ __forceinline__ double dxff(const DGrid &g, int j) { return metric(g, M_DXFF, j, g.dx); }
__device__ __forceinline__ double dycc(const DGrid &g, int j) { return metric(g, M_DYCC, j, g.dy); }
__device__ __forceinline__ double dyfc(const DGrid &g, int j) { return metric(g, M_DYFC, j, g.dy); }
__device__ __forceinline__ double dycf(const DGrid &g, int j) { return metric(g, M_DYCF, j, g.dy); }
__device__ __forceinline__ double dyff(const DGrid &g, int j) { return metric(g, M_DYFF, j, g.dy); }
__device__ __forceinline__ double azcc(const DGrid &g, int j) { return metric(g, M_AZCC, j, g.az); }
__device__ __forceinline__ double azfc(const DGrid &g, int j) { return metric(g, M_AZFC, j, g.az); }
__device__ __forceinline__ double azcf(const DGrid &g, int j) { return metric(g, M_AZCF, j, g.az); }
__device__ __forceinline__ double azff(const DGrid &g, int j) { return metric(g, M_AZFF, j, g.az); }

struct DParams {
    double Pstar, C, em2, Dmin, amin, amax, ca;
    int pform, cor;
    double min_mass, min_conc, rho_i, f;
    int top_kind, bot_kind;
    double ttx, tty;
    double rho_e, Cd, ue_c, ve_c;
    int u_sn_bc, v_we_bc;
    double u_sn_val, v_we_val;
    int adv_order, pad_;
    double imm_u, imm_v;  // immersed linear-drag flux BC coefficients (0 = none)
    int fd_kind, pad2_;   // free drift: CSI_FD_NONE / FIELDS / STRESS_BALANCE
    double top_rho, top_Cd;  // top SemiImplicitStress (u_e, v_e = top_x/top_y arrays or ttx/tty constants)
    const double *fff;       // HydrostaticSphericalCoriolis: f at (Face, Face) per row (device), row j at [j-1+Hy]
};

struct DFields {
    DArr u, v, h, a, s11, s22, s12, zf, zc, delta, alpha, un, vn, P;
    DArr top_x, top_y, ue, ve, Gh, Ga, hm, am, um, vm;
    DArr hs, Ghs, hsm, fd_u, fd_v;
};

// csi_thermo_fields on the device, same member order
struct DThermoFields {
    DArr h, a, hs, Tu, Tus, S, hc, Qtop, Qbot, Sb, Tb, snowfall, rho_s, mf_ice, mf_snow, mf_snowfall;
};

// index window a kernel runs over (inclusive, 1-based reference indices)
struct Range2 {
    int i0, i1, j0, j1;
};

}  // namespace csi

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
    int metL, metW;         // rows of the metric arrays; columns per row (0: the metrics depend on j only)
    const char *met_fused_why;  // NULL, or why the metric arrays rule the fused kernel out (checked once at csi_create)
    const double *met_host;  // host copies (plan construction only): the 12 x metL metrics, and f at (Face, Face) or NULL
    const double *fff_host;
    // north fold of a tripolar grid (CSI_FOLDED; topo_y is then CSI_BOUNDED with conn_n = 1: the north side is no wall): copy lists
    // per location lx + 2 ly, applied by the halo fill after the fills of the other sides
    int fold;
    const int32_t *fold_t[4], *fold_s[4];
    const int32_t *fold_t_host[4], *fold_s_host[4];  // host copies of the lists (plan construction of the fused solver)
    int fold_n[4];
    double fold_sv, fold_se;
};

// grid metrics at (i, j): constants on a RectilinearGrid, functions of j on a LatitudeLongitudeGrid (metW == 0), full 2-D
// arrays on an orthogonal curvilinear grid (metW = Nx + 2 Hx + 1 columns per row, entry (i, j) at [(j-1+Hy) metW + (i-1+Hx)])
enum { M_DXCC = 0, M_DXFC, M_DXCF, M_DXFF, M_DYCC, M_DYFC, M_DYCF, M_DYFF, M_AZCC, M_AZFC, M_AZCF, M_AZFF };
__device__ __forceinline__ double metric(const DGrid &g, int which, int i, int j, double regular)
{
    if (!g.met) return regular;
    if (g.metW) {
        const int pi = min(max(i - 1 + g.Hx, 0), g.metW - 1);
        return __ldg(g.met + ((size_t)which * g.metL + (j - 1 + g.Hy)) * g.metW + pi);
    }
    return __ldg(g.met + (size_t)which * g.metL + (j - 1 + g.Hy));
}
#define CSI_METRIC_FN(name, which, regular) \
    __device__ __forceinline__ double name(const DGrid &g, int i, int j) { return metric(g, which, i, j, regular); }
CSI_METRIC_FN(dxcc, M_DXCC, g.dx) CSI_METRIC_FN(dxfc, M_DXFC, g.dx) CSI_METRIC_FN(dxcf, M_DXCF, g.dx) CSI_METRIC_FN(dxff, M_DXFF, g.dx)
CSI_METRIC_FN(dycc, M_DYCC, g.dy) CSI_METRIC_FN(dyfc, M_DYFC, g.dy) CSI_METRIC_FN(dycf, M_DYCF, g.dy) CSI_METRIC_FN(dyff, M_DYFF, g.dy)
CSI_METRIC_FN(azcc, M_AZCC, g.az) CSI_METRIC_FN(azfc, M_AZFC, g.az) CSI_METRIC_FN(azcf, M_AZCF, g.az) CSI_METRIC_FN(azff, M_AZFF, g.az)
#undef CSI_METRIC_FN

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

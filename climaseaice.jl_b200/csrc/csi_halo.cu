// csi_halo.cu -- fill_halo_regions! for one field (Oceananigans BoundaryConditions, restated in
// SURVEY.md Appendix A): call sites src/SeaIceDynamics/split_explicit_momentum_equations.jl:170-187,
// src/Rheologies/elasto_visco_plastic_rheology.jl:275-280, src/sea_ice_model.jl:379-394.
//
// Order: non-periodic sides first over 1:N of the other axis, periodic sides last over the full
// parent extent of the other axis, so corners hold periodic images.  Sides that are rank
// boundaries of a slab partition are left to csi_exchange_halos (`only_local_halos = true`).
//   no-flux (Center, Bounded):  c[0] = c[1], c[N+1] = c[N]
//   value   (tangential velocity): c[0] = c[1] + ((c[1]-val)/(D/2))*(-D), c[N+1] = c[N] + ((val-c[N])/(D/2))*D
//   impenetrable (normal velocity): c[1] = 0, c[N+1] = 0
#include "csi_internal.h"

namespace csi {

enum { FILL_NONE = 0, FILL_NOFLUX = 1, FILL_VALUE = 2, FILL_IMPENETRABLE = 3 };

__global__ void k_fill_x_periodic(DArr a, int Nx, int Hx)
{
    const int pj = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y + 1;
    if (pj >= a.sy) return;
    const int j = pj + 1 - a.oy;
    at(a, 1 - k, j) = at(a, Nx + 1 - k, j);
    at(a, Nx + k, j) = at(a, k, j);
}

__global__ void k_fill_y_periodic(DArr a, int Ny, int Hy)
{
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y + 1;
    if (pi >= a.sx) return;
    const int i = pi + 1 - a.ox;
    at(a, i, 1 - k) = at(a, i, Ny + 1 - k);
    at(a, i, Ny + k) = at(a, i, k);
}

// j0..j1 / i0..i1: 1..N of the other axis, widened over the halo of its connected sides on a partition -- the substep loop
// computes there too (se:40-46), and a wall's boundary condition belongs to every column / row that is computed
__global__ void k_fill_x_bounded(DGrid g, DArr a, int Nx, int j0, int j1, int mode, double val, int do_west, int do_east)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x + j0;
    if (j > j1) return;
    const double Dw = dxff(g, 1, j), De = dxff(g, Nx + 1, j);  // Delta x at (Face, Face) on the walls, as Oceananigans' left/right_gradient uses
    if (mode == FILL_NOFLUX) {
        if (do_west) at(a, 0, j) = at(a, 1, j);
        if (do_east) at(a, Nx + 1, j) = at(a, Nx, j);
    } else if (mode == FILL_VALUE) {
        const double c1 = at(a, 1, j), cN = at(a, Nx, j);
        if (do_west) at(a, 0, j) = c1 + ((c1 - val) / (Dw / 2)) * (-Dw);
        if (do_east) at(a, Nx + 1, j) = cN + ((val - cN) / (De / 2)) * De;
    } else if (mode == FILL_IMPENETRABLE) {
        if (do_west) at(a, 1, j) = 0.0;
        if (do_east) at(a, Nx + 1, j) = 0.0;
    }
}

__global__ void k_fill_y_bounded(DGrid g, DArr a, int i0, int i1, int Ny, int mode, double val, int do_south, int do_north)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + i0;
    if (i > i1) return;
    const double Ds = dyff(g, i, 1), Dn = dyff(g, i, Ny + 1);
    if (mode == FILL_NOFLUX) {
        if (do_south) at(a, i, 0) = at(a, i, 1);
        if (do_north) at(a, i, Ny + 1) = at(a, i, Ny);
    } else if (mode == FILL_VALUE) {
        const double c1 = at(a, i, 1), cN = at(a, i, Ny);
        if (do_south) at(a, i, 0) = c1 + ((c1 - val) / (Ds / 2)) * (-Ds);
        if (do_north) at(a, i, Ny + 1) = cN + ((val - cN) / (Dn / 2)) * Dn;
    } else if (mode == FILL_IMPENETRABLE) {
        if (do_south) at(a, i, 1) = 0.0;
        if (do_north) at(a, i, Ny + 1) = 0.0;
    }
}

// the north fold (CSI_FOLDED): parent[target] = sign * parent[source] for every entry of the location's copy list
__global__ void k_fill_fold(DArr a, const int32_t *target, const int32_t *source, int n, double sign)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) a.p[target[k]] = sign * a.p[source[k]];
}

void launch_fold_list(const LaunchCtx &c, const DArr &a, const int32_t *target, const int32_t *source, int n, double sign)
{
    if (n <= 0) return;
    k_fill_fold<<<(n + 127) / 128, 128, 0, c.stream>>>(a, target, source, n, sign);
    ++*c.launches;
}

// which: 0 default boundary conditions, 1 = u, 2 = v (their own wall conditions, fold sign of velocities), 3 = an external stress /
// velocity array (default conditions, fold sign of external fields)
void launch_fill_halo(const LaunchCtx &c, const DGrid &g, const DParams &p, const DArr &a, int lx, int ly, int which)
{
    if (!a.p) return;
    const int T = 128;
    if (g.topo_x == CSI_BOUNDED && !(g.conn_w && g.conn_e)) {
        int mode = FILL_NONE;
        double val = 0.0;
        if (lx == 0) {
            mode = FILL_NOFLUX;
            if (which == 2 && p.v_we_bc == CSI_BC_VALUE) { mode = FILL_VALUE; val = p.v_we_val; }
        } else if (which == 1) {
            mode = FILL_IMPENETRABLE;
        }
        if (mode != FILL_NONE) {
            const int j0 = g.conn_s ? 1 - g.Hy : 1, j1 = g.conn_n ? g.Ny + g.Hy : g.Ny;
            k_fill_x_bounded<<<(j1 - j0 + T) / T, T, 0, c.stream>>>(g, a, g.Nx, j0, j1, mode, val, !g.conn_w, !g.conn_e);
            ++*c.launches;
        }
    }
    if (g.topo_y == CSI_BOUNDED && !(g.conn_s && g.conn_n)) {
        int mode = FILL_NONE;
        double val = 0.0;
        if (ly == 0) {
            mode = FILL_NOFLUX;
            if (which == 1 && p.u_sn_bc == CSI_BC_VALUE) { mode = FILL_VALUE; val = p.u_sn_val; }
        } else if (which == 2) {
            mode = FILL_IMPENETRABLE;
        }
        if (mode != FILL_NONE) {
            const int i0 = g.conn_w ? 1 - g.Hx : 1, i1 = g.conn_e ? g.Nx + g.Hx : g.Nx;
            k_fill_y_bounded<<<(i1 - i0 + T) / T, T, 0, c.stream>>>(g, a, i0, i1, g.Ny, mode, val, !g.conn_s, !g.conn_n);
            ++*c.launches;
        }
    }
    if (g.topo_x == CSI_PERIODIC && !g.conn_w && !g.conn_e) {
        k_fill_x_periodic<<<dim3((a.sy + T - 1) / T, g.Hx), T, 0, c.stream>>>(a, g.Nx, g.Hx);
        ++*c.launches;
    }
    if (g.topo_y == CSI_PERIODIC && !g.conn_s && !g.conn_n) {
        k_fill_y_periodic<<<dim3((a.sx + T - 1) / T, g.Hy), T, 0, c.stream>>>(a, g.Ny, g.Hy);
        ++*c.launches;
    }
    if (g.fold) {  // last: its sources are interior cells, its targets include the corners the periodic fill has just written
        const int loc = (lx ? 1 : 0) + (ly ? 2 : 0);
        if (g.fold_n[loc] > 0) {
            const double sign = (which == 1 || which == 2) ? g.fold_sv : (which == 3 ? g.fold_se : 1.0);
            k_fill_fold<<<(g.fold_n[loc] + T - 1) / T, T, 0, c.stream>>>(a, g.fold_t[loc], g.fold_s[loc], g.fold_n[loc], sign);
            ++*c.launches;
        }
    }
}

// mask_immersed_field_xy!: zero a field on peripheral nodes of its own location (immersed grids)
__global__ void k_mask_immersed(DGrid g, DArr a, int lx, int ly)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.Nx || j > g.Ny) return;
    auto imm = [&](int ii, int jj) {
        const int sx = g.Nx + 2 * g.Hx, sy = g.Ny + 2 * g.Hy;
        const int pi = min(max(ii - 1 + g.Hx, 0), sx - 1), pj = min(max(jj - 1 + g.Hy, 0), sy - 1);
        const bool out = (g.topo_x == CSI_BOUNDED && ((ii < 1 && !g.conn_w) || (ii > g.Nx && !g.conn_e))) ||
                         (g.topo_y == CSI_BOUNDED && ((jj < 1 && !g.conn_s) || (jj > g.Ny && !g.conn_n)));
        return out || g.mask[(size_t)pi + (size_t)pj * sx] != 0;
    };
    bool per;
    if (lx && !ly) per = imm(i - 1, j) || imm(i, j);
    else if (!lx && ly) per = imm(i, j - 1) || imm(i, j);
    else per = imm(i, j);
    if (per) at(a, i, j) = 0.0;
}

void launch_mask_immersed(const LaunchCtx &c, const DGrid &g, const DArr &a, int lx, int ly)
{
    if (!g.mask || !a.p) return;
    k_mask_immersed<<<dim3((g.Nx + 127) / 128, g.Ny), 128, 0, c.stream>>>(g, a, lx, ly);
    ++*c.launches;
}

}  // namespace csi

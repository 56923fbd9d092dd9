// csi_unfused.cu -- the EVP substep as separate kernels, one thread per node.
//
// Keeps the launch structure of the reference (initialize, stress, u, v; halo fills in
// csi_halo.cu) and works for every topology the library accepts.  The viscosity and stress
// kernels of the reference (evp:236-273, 294-354) are merged into one launch: the stress update
// reads zeta, Delta and P only at its own node (evp:314-318, 286), so the merge is order-neutral.
#include "csi_cell.cuh"
#include "csi_internal.h"

namespace csi {

static constexpr int BX = 32, BY = 8;

static dim3 grid_for(const Range2 &r) { return dim3((r.i1 - r.i0 + BX) / BX, (r.j1 - r.j0 + BY) / BY); }

// _initialize_evp_rhology!: evp:211-219, over the whole parent of P
__global__ void __launch_bounds__(BX *BY) k_initialize_rheology(const __grid_constant__ DGrid g, const __grid_constant__ DParams p,
                                                                 const __grid_constant__ DFields f, Range2 r)
{
    const int i = r.i0 + blockIdx.x * BX + threadIdx.x, j = r.j0 + blockIdx.y * BY + threadIdx.y;
    if (i > r.i1 || j > r.j1) return;
    at(f.P, i, j) = p.Pstar * at(f.h, i, j) * exp_cr(-p.C * (1 - at(f.a, i, j)));
    at(f.un, i, j) = at(f.u, i, j);
    at(f.vn, i, j) = at(f.v, i, j);
}

__global__ void __launch_bounds__(BX *BY) k_evp_stress(const __grid_constant__ DGrid g, const __grid_constant__ DParams p,
                                                        const __grid_constant__ DFields f, double dt, Range2 r)
{
    const int i = r.i0 + blockIdx.x * BX + threadIdx.x, j = r.j0 + blockIdx.y * BY + threadIdx.y;
    if (i > r.i1 || j > r.j1) return;
    evp_stress_node(g, p, f, dt, i, j);
}

__global__ void __launch_bounds__(BX *BY) k_u_step(const __grid_constant__ DGrid g, const __grid_constant__ DParams p,
                                                    const __grid_constant__ DFields f, double dt, Range2 r)
{
    const int i = r.i0 + blockIdx.x * BX + threadIdx.x, j = r.j0 + blockIdx.y * BY + threadIdx.y;
    if (i > r.i1 || j > r.j1) return;
    u_step_node(g, p, f, dt, i, j);
}

__global__ void __launch_bounds__(BX *BY) k_v_step(const __grid_constant__ DGrid g, const __grid_constant__ DParams p,
                                                    const __grid_constant__ DFields f, double dt, Range2 r)
{
    const int i = r.i0 + blockIdx.x * BX + threadIdx.x, j = r.j0 + blockIdx.y * BY + threadIdx.y;
    if (i > r.i1 || j > r.j1) return;
    v_step_node(g, p, f, dt, i, j);
}

void launch_initialize_rheology(const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f)
{
    Range2 r{1 - g.Hx, g.Nx + g.Hx, 1 - g.Hy, g.Ny + g.Hy};
    k_initialize_rheology<<<grid_for(r), dim3(BX, BY), 0, c.stream>>>(g, p, f, r);
    ++*c.launches;
}

void launch_evp_stress(const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f, double dt)
{
    Range2 r{-g.Hx + 2, g.Nx + g.Hx - 1, -g.Hy + 2, g.Ny + g.Hy - 1};  // evp:145
    k_evp_stress<<<grid_for(r), dim3(BX, BY), 0, c.stream>>>(g, p, f, dt, r);
    ++*c.launches;
}

void launch_u_step(const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f, double dt, Range2 r)
{
    k_u_step<<<grid_for(r), dim3(BX, BY), 0, c.stream>>>(g, p, f, dt, r);
    ++*c.launches;
}

void launch_v_step(const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f, double dt, Range2 r)
{
    k_v_step<<<grid_for(r), dim3(BX, BY), 0, c.stream>>>(g, p, f, dt, r);
    ++*c.launches;
}

}  // namespace csi

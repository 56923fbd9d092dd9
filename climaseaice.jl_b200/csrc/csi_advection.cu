// csi_advection.cu -- h / aice flux-form advection.
//
//   compute_tracer_tendencies!  src/tracer_tendency_kernel_functions.jl:9-45
//   horizontal_div_Uc           src/sea_ice_advection.jl:51-58
//   _dynamic_step_tracers!      src/sea_ice_fe_step.jl:56-82 (drivers fe:36-50, rk:134-152)
//
// The face reconstruction is Oceananigans' upwind-biased WENO-Z (order 3/5/7) or first-order
// upwind, restated from SURVEY.md Appendix A.  One CTA computes a TX x TY block of cells from a
// shared-memory tile of h, aice (halo B) and u, v; the x- and y-face fluxes are evaluated once per
// face into shared memory and differenced, so each flux is computed once per tile instead of
// twice per cell as in the reference's per-cell formulation (identical bits: same expression).
#include "csi_cell.cuh"
#include "csi_internal.h"

namespace csi {

namespace weno {
__device__ constexpr double EPS = 1e-8;
__device__ constexpr double R2[2][2] = {{1.0 / 2.0, 1.0 / 2.0}, {-1.0 / 2.0, 3.0 / 2.0}};
__device__ constexpr double R3[3][3] = {{1.0 / 3.0, 5.0 / 6.0, -1.0 / 6.0}, {-1.0 / 6.0, 5.0 / 6.0, 1.0 / 3.0}, {1.0 / 3.0, -7.0 / 6.0, 11.0 / 6.0}};
__device__ constexpr double R4[4][4] = {{1.0 / 4.0, 13.0 / 12.0, -5.0 / 12.0, 1.0 / 12.0},
                                        {-1.0 / 12.0, 7.0 / 12.0, 7.0 / 12.0, -1.0 / 12.0},
                                        {1.0 / 12.0, -5.0 / 12.0, 13.0 / 12.0, 1.0 / 4.0},
                                        {-1.0 / 4.0, 13.0 / 12.0, -23.0 / 12.0, 25.0 / 12.0}};
__device__ constexpr double C2[2] = {2.0 / 3.0, 1.0 / 3.0};
__device__ constexpr double C3[3] = {3.0 / 10.0, 3.0 / 5.0, 1.0 / 10.0};
__device__ constexpr double C4[4] = {4.0 / 35.0, 18.0 / 35.0, 12.0 / 35.0, 1.0 / 35.0};
__device__ constexpr double S3[3][6] = {{10, -31, 11, 25, -19, 4}, {4, -13, 5, 13, -13, 4}, {4, -19, 11, 25, -31, 10}};
__device__ constexpr double S4[4][10] = {{2.107, -9.402, 7.042, -1.854, 11.003, -17.246, 4.642, 7.043, -3.882, 0.547},
                                         {0.547, -2.522, 1.922, -0.494, 3.443, -5.966, 1.602, 2.843, -1.642, 0.267},
                                         {0.267, -1.642, 1.602, -0.494, 2.843, -5.966, 1.922, 3.443, -2.522, 0.547},
                                         {0.547, -3.882, 4.642, -1.854, 7.043, -17.246, 7.042, 11.003, -9.402, 2.107}};

__device__ __forceinline__ double zalpha(double beta, double tau, double C)
{
    const double r = tau / (beta + EPS);
    return C * (1 + r * r);
}

// q[n], n = 0..2B-2: the biased stencil ordered so that candidate k, slot m is q[(B-1-k)+m]
template <int B> __device__ __forceinline__ double reconstruct(const double *q);
template <> __device__ __forceinline__ double reconstruct<1>(const double *q) { return q[0]; }
template <> __device__ __forceinline__ double reconstruct<2>(const double *q)
{
    double beta[2], pk[2], al[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const double *s = q + (1 - k);
        beta[k] = s[0] * (1.0 * s[0] + -2.0 * s[1]) + 1.0 * (s[1] * s[1]);
        pk[k] = R2[k][0] * s[0] + R2[k][1] * s[1];
    }
    const double tau = fabs(beta[0] - beta[1]);
#pragma unroll
    for (int k = 0; k < 2; k++) al[k] = zalpha(beta[k], tau, C2[k]);
    const double sum = al[0] + al[1];
    return (al[0] / sum) * pk[0] + (al[1] / sum) * pk[1];
}
template <> __device__ __forceinline__ double reconstruct<3>(const double *q)
{
    double beta[3], pk[3], al[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double *s = q + (2 - k);
        const double *c = S3[k];
        beta[k] = s[0] * (c[0] * s[0] + c[1] * s[1] + c[2] * s[2]) + s[1] * (c[3] * s[1] + c[4] * s[2]) + c[5] * (s[2] * s[2]);
        pk[k] = R3[k][0] * s[0] + R3[k][1] * s[1] + R3[k][2] * s[2];
    }
    const double tau = fabs(beta[0] - beta[2]);
#pragma unroll
    for (int k = 0; k < 3; k++) al[k] = zalpha(beta[k], tau, C3[k]);
    const double sum = al[0] + al[1] + al[2];
    return (al[0] / sum) * pk[0] + (al[1] / sum) * pk[1] + (al[2] / sum) * pk[2];
}
template <> __device__ __forceinline__ double reconstruct<4>(const double *q)
{
    double beta[4], pk[4], al[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double *s = q + (3 - k);
        const double *c = S4[k];
        beta[k] = s[0] * (c[0] * s[0] + c[1] * s[1] + c[2] * s[2] + c[3] * s[3]) + s[1] * (c[4] * s[1] + c[5] * s[2] + c[6] * s[3]) +
                  s[2] * (c[7] * s[2] + c[8] * s[3]) + c[9] * (s[3] * s[3]);
        pk[k] = R4[k][0] * s[0] + R4[k][1] * s[1] + R4[k][2] * s[2] + R4[k][3] * s[3];
    }
    const double tau = fabs(beta[0] + 3 * beta[1] - 3 * beta[2] - beta[3]);
#pragma unroll
    for (int k = 0; k < 4; k++) al[k] = zalpha(beta[k], tau, C4[k]);
    const double sum = al[0] + al[1] + al[2] + al[3];
    return (al[0] / sum) * pk[0] + (al[1] / sum) * pk[1] + (al[2] / sum) * pk[2] + (al[3] / sum) * pk[3];
}

// face value from a strided line of cells: cm points at cell (face-1), stride between cells
template <int B> __device__ __forceinline__ double face_value(const double *cm, int stride, bool left)
{
    double q[2 * B - 1];
#pragma unroll
    for (int n = 0; n < 2 * B - 1; n++) q[n] = left ? cm[(n - (B - 1)) * stride] : cm[(B - n) * stride];
    return reconstruct<B>(q);
}
__device__ __forceinline__ double face_value_dyn(int b, const double *cm, int stride, bool left)
{
    switch (b) {
    case 1: return face_value<1>(cm, stride, left);
    case 2: return face_value<2>(cm, stride, left);
    case 3: return face_value<3>(cm, stride, left);
    default: return face_value<4>(cm, stride, left);
    }
}
}  // namespace weno

// order reduction next to Bounded walls: buffer = min(B, face-1, N+1-face), at least 1; a slab's connected
// side is not a wall
__device__ __forceinline__ int buffer_at(int B, bool wall_lo, bool wall_hi, int N, int face)
{
    int b = B;
    if (wall_lo) b = min(b, face - 1);
    if (wall_hi) b = min(b, N + 1 - face);
    return b < 1 ? 1 : b;
}

// ImmersedBoundaryGrid: a scheme of buffer b is used at a face only if none of the 2b cells of its left- and right-biased
// stencils (face-b .. face+b-1 along the stencil direction) is inactive, else the next lower order is tried (Oceananigans'
// near_*_immersed_boundary_biased + buffer_scheme chain, call site src/sea_ice_advection.jl:51-58); first order is never tested
template <bool XDIR> __device__ __forceinline__ int buffer_immersed(const DGrid &g, int b0, int i, int j)
{
    for (int b = b0; b >= 2; b--) {
        bool bad = false;
        for (int c = -b; c <= b - 1; c++) bad = bad || (XDIR ? inactive_cell(g, i + c, j) : inactive_cell(g, i, j + c));
        if (!bad) return b;
    }
    return 1;
}

// A CTA of 32 x 8 threads owns 31 x 7 cells: its 32 x 7 x-faces and 31 x 8 y-faces are then ONE sweep of the threads each.  (With a
// 32 x 8 cell tile the 33 x 8 and 32 x 9 faces took a second, almost empty sweep, and the block barrier behind it was the kernel's
// largest stall: profiles/r02_tracer_tendencies_kernel_ncu.md.)
static constexpr int BTX = 32, BTY = 8;          // threads
static constexpr int ATX = 31, ATY = 7, AH = 4;  // cells per tile and the widest stencil halo (WENO7)
static_assert((ATX + 1) * ATY <= BTX * BTY && ATX * (ATY + 1) <= BTX * BTY, "one sweep per face direction");

// G^n.h = -div(U h), G^n.aice = -div(U aice) and, with snow (NQ = 3), G^n.hs = -div(U hs)  (tracer_tendency:27-52)
template <int B, int NQ>
__global__ void __launch_bounds__(BTX *BTY) k_tracer_tendencies(const __grid_constant__ DGrid g, const __grid_constant__ DParams p,
                                                                 const __grid_constant__ DFields f)
{
    constexpr int SX = ATX + 2 * AH, SY = ATY + 2 * AH;
    __shared__ double sh[NQ][SY][SX];          // h, aice (, hs) with halo
    __shared__ double fx[NQ][ATY][ATX + 1];    // x-face fluxes
    __shared__ double fy[NQ][ATY + 1][ATX];    // y-face fluxes
    const int i0 = blockIdx.x * ATX + 1, j0 = blockIdx.y * ATY + 1;
    const int tid = threadIdx.y * BTX + threadIdx.x;
    for (int t = tid; t < SX * SY; t += BTX * BTY) {
        const int li = t % SX, lj = t / SX;
        int gi = i0 - AH + li, gj = j0 - AH + lj;
        // stay inside the parent array (cells beyond the halo are never used by a valid stencil)
        gi = min(max(gi, 1 - g.Hx), g.Nx + g.Hx);
        gj = min(max(gj, 1 - g.Hy), g.Ny + g.Hy);
        sh[0][lj][li] = at(f.h, gi, gj);
        sh[1][lj][li] = at(f.a, gi, gj);
        if (NQ > 2) sh[NQ - 1][lj][li] = at(f.hs, gi, gj);
    }
    __syncthreads();
    const bool bx_lo = g.topo_x == CSI_BOUNDED && !g.conn_w, bx_hi = g.topo_x == CSI_BOUNDED && !g.conn_e, by_lo = g.topo_y == CSI_BOUNDED && !g.conn_s, by_hi = g.topo_y == CSI_BOUNDED && !g.conn_n;
    // x faces: (ATX+1) x ATY
    for (int t = tid; t < (ATX + 1) * ATY; t += BTX * BTY) {
        const int li = t % (ATX + 1), lj = t / (ATX + 1);
        const int i = i0 + li, j = j0 + lj;
        if (i <= g.Nx + 1 && j <= g.Ny) {
            const double U = at(f.u, min(i, f.u.sx - f.u.ox), j);
            int b = buffer_at(B, bx_lo, bx_hi, g.Nx, i);
            if (g.mask) b = buffer_immersed<true>(g, b, i, j);
            const bool imm = g.mask && imm_peripheral_fc(g, i, j);
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const double ct = weno::face_value_dyn(b, &sh[q][lj + AH][li + AH - 1], 1, U > 0);
                const double fl = (dyfc(g, i, j) * 1.0) * U * ct;
                fx[q][lj][li] = imm ? 0.0 : fl;
            }
        }
    }
    // y faces: ATX x (ATY+1)
    for (int t = tid; t < ATX * (ATY + 1); t += BTX * BTY) {
        const int li = t % ATX, lj = t / ATX;
        const int i = i0 + li, j = j0 + lj;
        if (i <= g.Nx && j <= g.Ny + 1) {
            const double V = at(f.v, i, min(j, f.v.sy - f.v.oy));
            int b = buffer_at(B, by_lo, by_hi, g.Ny, j);
            if (g.mask) b = buffer_immersed<false>(g, b, i, j);
            const bool imm = g.mask && imm_peripheral_cf(g, i, j);
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const double ct = weno::face_value_dyn(b, &sh[q][lj + AH - 1][li + AH], SX, V > 0);
                const double fl = (dxcf(g, i, j) * 1.0) * V * ct;
                fy[q][lj][li] = imm ? 0.0 : fl;
            }
        }
    }
    __syncthreads();
    const int li = threadIdx.x, lj = threadIdx.y, i = i0 + li, j = j0 + lj;
    if (li < ATX && lj < ATY && i <= g.Nx && j <= g.Ny) {
        const double V = azcc(g, i, j) * 1.0;
        at(f.Gh, i, j) = -(1 / V * ((fx[0][lj][li + 1] - fx[0][lj][li]) + (fy[0][lj + 1][li] - fy[0][lj][li])));
        at(f.Ga, i, j) = -(1 / V * ((fx[1][lj][li + 1] - fx[1][lj][li]) + (fy[1][lj + 1][li] - fy[1][lj][li])));
        if (NQ > 2) at(f.Ghs, i, j) = -(1 / V * ((fx[NQ - 1][lj][li + 1] - fx[NQ - 1][lj][li]) + (fy[NQ - 1][lj + 1][li] - fy[NQ - 1][lj][li])));
    }
}

__global__ void __launch_bounds__(256) k_zero_tendencies(const __grid_constant__ DGrid g, const __grid_constant__ DFields f)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.Nx || j > g.Ny) return;
    at(f.Gh, i, j) = -0.0;  // -zero(grid)
    at(f.Ga, i, j) = -0.0;
    if (f.hs.p) at(f.Ghs, i, j) = -0.0;
}

void launch_tracer_tendencies(const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f)
{
    dim3 grid((g.Nx + ATX - 1) / ATX, (g.Ny + ATY - 1) / ATY), block(BTX, BTY);
    switch (p.adv_order) {
    case 0: k_zero_tendencies<<<dim3((g.Nx + 255) / 256, g.Ny), 256, 0, c.stream>>>(g, f); break;
#define CSI_TT(B_)                                                                     \
    if (f.hs.p) k_tracer_tendencies<B_, 3><<<grid, block, 0, c.stream>>>(g, p, f);     \
    else k_tracer_tendencies<B_, 2><<<grid, block, 0, c.stream>>>(g, p, f);            \
    break
    case 1: CSI_TT(1);
    case 3: CSI_TT(2);
    case 5: CSI_TT(3);
    default: CSI_TT(4);
#undef CSI_TT
    }
    ++*c.launches;
}

// _dynamic_step_tracers!: fe:56-82, dynamic_step_snow!: fe:84-94
__global__ void __launch_bounds__(256) k_dynamic_step(const __grid_constant__ DGrid g, const __grid_constant__ DFields f, DArr hn, DArr an, DArr hsn, double dt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > g.Nx || j > g.Ny) return;
    double hp = at(hn, i, j) + dt * at(f.Gh, i, j);
    double ap = at(an, i, j) + dt * at(f.Ga, i, j);
    ap = jl_max(0.0, ap);
    hp = jl_max(0.0, hp);
    ap = (hp == 0) ? 0.0 : ap;
    hp = (ap == 0) ? 0.0 : hp;
    const double Vp = hp * ap;
    const double anew = ap > 1 ? 1.0 : ap;
    at(f.a, i, j) = anew;
    at(f.h, i, j) = ap > 1 ? Vp : hp;
    if (f.hs.p) {
        double sp = at(hsn, i, j) + dt * at(f.Ghs, i, j);
        sp = jl_max(0.0, sp);
        sp = (anew <= 0) ? 0.0 : sp;
        at(f.hs, i, j) = sp;
    }
}

void launch_dynamic_step(const LaunchCtx &c, const DGrid &g, const DFields &f, const DArr &hn, const DArr &an, const DArr &hsn, double dt)
{
    k_dynamic_step<<<dim3((g.Nx + 255) / 256, g.Ny), 256, 0, c.stream>>>(g, f, hn, an, hsn, dt);
    ++*c.launches;
}

}  // namespace csi

// csi_reduce.cu -- grid-wide diagnostics with warp-shuffle reductions.
//
//   cell_advection_timescale(model)  src/ClimaSeaIce.jl:66-69  -> min over cells of 1/(|u|/dx + |v|/dy)
//   conservation diagnostics: sum h*Az, sum aice*Az, sum h*aice*Az, max|u|, max|v|
//
// Deterministic by construction: pass 1 gives every CTA a fixed, contiguous set of rows and
// reduces them in a fixed order (thread-serial, then shuffle tree, then a fixed-order combine of
// the warp partials); pass 2 is one CTA folding the per-CTA partials in index order.  No atomics.
// Nothing reduced here ever feeds back into the model state.
#include "csi_internal.h"

namespace csi {

static constexpr int RT = 256;

struct Acc {
    double s0, s1, s2, m0, m1, tmin;
};
__device__ __forceinline__ Acc acc_identity()
{
    Acc a;
    a.s0 = a.s1 = a.s2 = 0.0;
    a.m0 = a.m1 = 0.0;
    a.tmin = INFINITY;
    return a;
}
__device__ __forceinline__ Acc acc_merge(const Acc &a, const Acc &b)
{
    Acc r;
    r.s0 = a.s0 + b.s0;
    r.s1 = a.s1 + b.s1;
    r.s2 = a.s2 + b.s2;
    r.m0 = fmax(a.m0, b.m0);
    r.m1 = fmax(a.m1, b.m1);
    r.tmin = fmin(a.tmin, b.tmin);
    return r;
}
__device__ __forceinline__ Acc acc_shfl_down(const Acc &a, int d)
{
    Acc r;
    r.s0 = __shfl_down_sync(0xffffffffu, a.s0, d);
    r.s1 = __shfl_down_sync(0xffffffffu, a.s1, d);
    r.s2 = __shfl_down_sync(0xffffffffu, a.s2, d);
    r.m0 = __shfl_down_sync(0xffffffffu, a.m0, d);
    r.m1 = __shfl_down_sync(0xffffffffu, a.m1, d);
    r.tmin = __shfl_down_sync(0xffffffffu, a.tmin, d);
    return r;
}
__device__ __forceinline__ Acc block_reduce(Acc a)
{
    __shared__ Acc warp_part[RT / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a = acc_merge(a, acc_shfl_down(a, d));
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = a;
    __syncthreads();
    Acc r = acc_identity();
    if (threadIdx.x == 0)
        for (int w = 0; w < RT / 32; w++) r = acc_merge(r, warp_part[w]);
    return r;  // valid in thread 0
}

__global__ void __launch_bounds__(RT) k_reduce_pass1(const __grid_constant__ DGrid g, const __grid_constant__ DFields f, Acc *part)
{
    Acc a = acc_identity();
    const int rows_per_cta = (g.Ny + gridDim.x - 1) / gridDim.x;
    const int j0 = 1 + blockIdx.x * rows_per_cta, j1 = min(g.Ny, j0 + rows_per_cta - 1);
    for (int j = j0; j <= j1; j++)
        for (int i = 1 + threadIdx.x; i <= g.Nx; i += RT) {
            const double h = f.h.p ? at(f.h, i, j) : 0.0, c = f.a.p ? at(f.a, i, j) : 0.0;
            const double u = at(f.u, i, j), v = at(f.v, i, j);
            Acc b;
            b.s0 = h * azcc(g, i, j);
            b.s1 = c * azcc(g, i, j);
            b.s2 = h * c * azcc(g, i, j);
            b.m0 = fabs(u);
            b.m1 = fabs(v);
            b.tmin = 1 / (fabs(u) / dxfc(g, i, j) + fabs(v) / dycf(g, i, j));
            a = acc_merge(a, b);
        }
    a = block_reduce(a);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
}

__global__ void __launch_bounds__(RT) k_reduce_pass2(const Acc *part, int n, double *out6)
{
    Acc a = acc_identity();
    for (int k = threadIdx.x; k < n; k += RT) a = acc_merge(a, part[k]);
    a = block_reduce(a);
    if (threadIdx.x == 0) {
        out6[0] = a.s0;
        out6[1] = a.s1;
        out6[2] = a.s2;
        out6[3] = a.m0;
        out6[4] = a.m1;
        out6[5] = a.tmin;
    }
}

static void reduce_all(const LaunchCtx &c, const DGrid &g, const DFields &f, double *scratch, int nscratch, double *out6)
{
    int nblk = (int)(nscratch * sizeof(double) / sizeof(Acc));
    if (nblk > 592) nblk = 592;  // 4 CTAs per SM x 148 SMs
    if (nblk > g.Ny) nblk = g.Ny;
    k_reduce_pass1<<<nblk, RT, 0, c.stream>>>(g, f, reinterpret_cast<Acc *>(scratch));
    k_reduce_pass2<<<1, RT, 0, c.stream>>>(reinterpret_cast<const Acc *>(scratch), nblk, out6);
    *c.launches += 2;
}

void launch_cfl(const LaunchCtx &c, const DGrid &g, const DFields &f, double *scratch, int nscratch, double *out_dev)
{
    reduce_all(c, g, f, scratch, nscratch, out_dev);
}
void launch_diagnostics(const LaunchCtx &c, const DGrid &g, const DFields &f, double *scratch, int nscratch, double *out_dev)
{
    reduce_all(c, g, f, scratch, nscratch, out_dev);
}

}  // namespace csi

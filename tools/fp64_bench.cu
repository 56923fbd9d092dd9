// Microbenchmark: issue rate of FP64 DFMA / DADD / DMUL / division / sqrt on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/bin/fp64_bench tools/fp64_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int ILP> __global__ void k(double *out, double a, double b, int iters)
{
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = a + threadIdx.x * 1e-9 + k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            if (OP == 0) x[k] = __fma_rn(x[k], b, a);
            else if (OP == 1) x[k] = __dadd_rn(x[k], b);
            else if (OP == 2) x[k] = __dmul_rn(x[k], b);
            else if (OP == 3) x[k] = a / x[k] + b;
            else if (OP == 4) x[k] = sqrt(x[k]) + b;
            else if (OP == 5) x[k] = __fma_rn(x[k], 1.0, b);      // add written as fma
            else if (OP == 6) x[k] = __fma_rn(x[k], b, -0.0);     // mul written as fma
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP, int ILP> void run(const char *name, int blocks_per_sm, int threads)
{
    int nsm = 148;
    double *out;
    cudaMalloc(&out, sizeof(double) * nsm * blocks_per_sm * threads);
    const int iters = 4096;
    k<OP, ILP><<<nsm * blocks_per_sm, threads>>>(out, 1.000001, 0.9999999, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP, ILP><<<nsm * blocks_per_sm, threads>>>(out, 1.000001, 0.9999999, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)nsm * blocks_per_sm * threads * iters * ILP;
    printf("%-28s ILP=%d warps/SM=%2d : %8.2f G op/s  (%.1f thread-ops/clk/SM @1.965GHz)\n", name, ILP, blocks_per_sm * threads / 32, ops / ms / 1e6,
           ops / ms / 1e6 * 1e9 / 148 / 1.965e9);
    cudaFree(out);
}
int main()
{
    run<0, 8>("DFMA", 4, 256);
    run<1, 8>("DADD", 4, 256);
    run<2, 8>("DMUL", 4, 256);
    run<5, 8>("DADD as fma(x,1,b)", 4, 256);
    run<6, 8>("DMUL as fma(x,b,-0)", 4, 256);
    run<0, 1>("DFMA", 1, 128);
    run<0, 1>("DFMA", 2, 256);
    run<0, 2>("DFMA", 2, 256);
    run<0, 4>("DFMA", 2, 256);
    run<0, 1>("DFMA", 8, 256);
    run<1, 1>("DADD", 1, 128);
    run<3, 4>("div (a/x+b)", 4, 256);
    run<3, 1>("div (a/x+b)", 1, 128);
    run<4, 4>("sqrt(x)+b", 4, 256);
    return 0;
}

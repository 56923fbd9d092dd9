// csi_bench.cu -- device microbenchmarks behind the roofline denominators of bench.py (instrumentation, not on the hot path).
//
// csi_measure_fp64_rate: thread-level FP64 FMA instructions per second of this device, with eight independent chains per
// thread and 16 warps per SM sub-partition (enough to cover the pipe's dependent-issue latency), as a burst (best short
// launch, boost clock) and sustained (mean over the second half of `seconds` of back-to-back launches, i.e. under the
// power cap the EVP kernel also runs into).  The fused substep kernel's FP64 instruction rate is reported against it.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/climaseaice_b200.h"

namespace {

__global__ void __launch_bounds__(512) k_fp64_chains(double *out, int iters, double b, double c)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3, a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
#pragma unroll 4
    for (int k = 0; k < iters; k++) {
        a0 = __fma_rn(a0, b, c); a1 = __fma_rn(a1, b, c); a2 = __fma_rn(a2, b, c); a3 = __fma_rn(a3, b, c);
        a4 = __fma_rn(a4, b, c); a5 = __fma_rn(a5, b, c); a6 = __fma_rn(a6, b, c); a7 = __fma_rn(a7, b, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) out[0] = s;  // keeps the chains alive
}

}  // namespace

extern "C" int csi_measure_fp64_rate(int32_t device, double seconds, double *burst_fma_per_s, double *sustained_fma_per_s)
{
    if (!burst_fma_per_s || !sustained_fma_per_s || !(seconds > 0) || seconds > 30) return CSI_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return CSI_ERR_NO_DEVICE;
    }
    cudaSetDevice(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return CSI_ERR_NO_DEVICE;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double)) != cudaSuccess) return 1;
    const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 4096;
    const double fma_per_launch = (double)blocks * threads * iters * 8.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    std::vector<float> ms;
    double total_ms = 0;
    k_fp64_chains<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);  // module load
    cudaDeviceSynchronize();
    while (total_ms < seconds * 1e3 && ms.size() < 100000) {
        cudaEventRecord(e0);
        for (int r = 0; r < 8; r++) k_fp64_chains<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        ms.push_back(t / 8);
        total_ms += t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (ms.empty() || cudaGetLastError() != cudaSuccess) return 1;
    const float best = *std::min_element(ms.begin(), ms.end());
    double tail = 0;
    const size_t half = ms.size() / 2;
    for (size_t k = half; k < ms.size(); k++) tail += ms[k];
    tail /= (double)(ms.size() - half);
    *burst_fma_per_s = fma_per_launch / (best * 1e-3);
    *sustained_fma_per_s = fma_per_launch / (tail * 1e-3);
    return CSI_OK;
}

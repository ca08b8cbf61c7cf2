// peaks.cuh -- measured FP64 arithmetic peaks of the device this library runs on, for the roofline denominators of the
// latency / issue-bound whole-solve kernels (bench.py `step_roofline`): MEASURED_PEAKS.json only carries HBM GB/s and dense
// bf16 TF/s, and the 40 TFLOP/s fp64 figure of the data sheet is nominal.
//   dfma_kernel : 16 independent DFMA chains per thread, 8 warps per CTA, 4 CTAs per SM (the FP64 vector pipe saturated)
//   dmma_kernel : 4 independent mma.sync.m8n8k4.f64 accumulator chains per warp (SASS DMMA.8x8x4), same residency
#pragma once
#include <cuda_runtime.h>

namespace riptrm {
namespace peaks {

constexpr int kChains = 16;
constexpr int kUnroll = 64;

__global__ void __launch_bounds__(256, 4) dfma_kernel(double* sink, int reps, double a, double b) {
    double acc[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) acc[i] = (double)(threadIdx.x + i);
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
#pragma unroll
            for (int i = 0; i < kChains; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) s += acc[i];
    if (s == 123.456) sink[0] = s;   // never true: keeps the chains alive
}

__global__ void __launch_bounds__(256, 4) dmma_kernel(double* sink, int reps, double a, double b) {
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = (double)i;
    const double fa = a + (double)(threadIdx.x & 3) * 1e-9, fb = b;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(fa), "d"(fb));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

}  // namespace peaks
}  // namespace riptrm

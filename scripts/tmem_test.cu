// tmem_test.cu -- can Tensor Memory hold the 50 x 50 fp64 matrix of a warp-per-pair solve, and is S.v from TMEM faster
// than from shared memory when every SM is full?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cuda_runtime.h>
#include "../riemannian-interior-point-trust-region-method_b200/csrc/tmem.cuh"
using namespace riptrm;
constexpr int N = 50, REPS = 200;

// MODE 0: S in shared memory (20 KB per warp), MODE 1: S in TMEM (200 columns per warp)
template <int MODE>
__global__ void __launch_bounds__(128, 2) kern(const double* gS, double* out, long long* cyc, int* err) {
    extern __shared__ __align__(16) double smem[];
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* S = smem + warp * (N * N + 64);           // staging / resident copy
    double* vb = smem + 4 * (N * N + 64) + warp * 64;
    for (int i = lane; i < N * N + 64; i += 32) S[i] = (i < N * N) ? gS[i] : 0.0;
    vb[lane] = 0; vb[lane + 32] = 0;
    if (MODE == 1) {
        if (warp == 0) tmem::alloc(&tbase, 256);
        tmem::fence_before_sync();
        __syncthreads();
        tmem::fence_after_sync();
    }
    __syncwarp();
    const uint32_t taddr = (MODE == 1) ? (tbase + ((uint32_t)(32 * warp) << 16)) : 0u;
    if (MODE == 1) {
        // lane l holds S[j][2l], S[j][2l+1] for all j: columns 4j .. 4j+3 of its TMEM lane
        const double2* row = reinterpret_cast<const double2*>(S) + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            uint32_t r[64];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const double2 s = row[(16 * c + jj) * (N / 2)];
                r[4 * jj + 0] = (uint32_t)__double2loint(s.x); r[4 * jj + 1] = (uint32_t)__double2hiint(s.x);
                r[4 * jj + 2] = (uint32_t)__double2loint(s.y); r[4 * jj + 3] = (uint32_t)__double2hiint(s.y);
            }
            tmem::tmem_st_x64(taddr + 64 * c, r);
        }
        {
            uint32_t r[8];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                const double2 s = row[(48 + jj) * (N / 2)];
                r[4 * jj + 0] = (uint32_t)__double2loint(s.x); r[4 * jj + 1] = (uint32_t)__double2hiint(s.x);
                r[4 * jj + 2] = (uint32_t)__double2loint(s.y); r[4 * jj + 3] = (uint32_t)__double2hiint(s.y);
            }
            tmem::tmem_st_x8(taddr + 192, r);
        }
        tmem::wait_st();
    }
    double v0 = 1.0 + lane * 1e-3, v1 = 0.5 - lane * 1e-3;
    if (lane >= 25) v0 = v1 = 0.0;
    const long long t0 = clock64();
    for (int rep = 0; rep < REPS; ++rep) {
        reinterpret_cast<double2*>(vb)[lane] = make_double2(v0, v1);
        __syncwarp();
        double a0x = 0, a0y = 0, a1x = 0, a1y = 0;
        if (MODE == 0) {
            const double2* row = reinterpret_cast<const double2*>(S) + lane;
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                const double2 vj = *reinterpret_cast<const double2*>(vb + j);
                const double2 s0 = row[j * (N / 2)], s1 = row[(j + 1) * (N / 2)];
                a0x = fma(s0.x, vj.x, a0x); a0y = fma(s0.y, vj.x, a0y);
                a1x = fma(s1.x, vj.y, a1x); a1y = fma(s1.y, vj.y, a1y);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                uint32_t r[64];
                tmem::tmem_ld_x64(taddr + 64 * c, r);
                tmem::wait_ld();
#pragma unroll
                for (int jj = 0; jj < 16; jj += 2) {
                    const double2 vj = *reinterpret_cast<const double2*>(vb + 16 * c + jj);
                    a0x = fma(__hiloint2double(r[4 * jj + 1], r[4 * jj + 0]), vj.x, a0x);
                    a0y = fma(__hiloint2double(r[4 * jj + 3], r[4 * jj + 2]), vj.x, a0y);
                    a1x = fma(__hiloint2double(r[4 * jj + 5], r[4 * jj + 4]), vj.y, a1x);
                    a1y = fma(__hiloint2double(r[4 * jj + 7], r[4 * jj + 6]), vj.y, a1y);
                }
            }
            {
                uint32_t r[8];
                tmem::tmem_ld_x8(taddr + 192, r);
                tmem::wait_ld();
                const double2 vj = *reinterpret_cast<const double2*>(vb + 48);
                a0x = fma(__hiloint2double(r[1], r[0]), vj.x, a0x);
                a0y = fma(__hiloint2double(r[3], r[2]), vj.x, a0y);
                a1x = fma(__hiloint2double(r[5], r[4]), vj.y, a1x);
                a1y = fma(__hiloint2double(r[7], r[6]), vj.y, a1y);
            }
        }
        __syncwarp();
        const double o0 = a0x + a1x, o1 = a0y + a1y;
        v0 = (lane < 25) ? o0 * 0.03 : 0.0;
        v1 = (lane < 25) ? o1 * 0.03 : 0.0;
    }
    const long long t1 = clock64();
    if (lane < 25) {
        out[(blockIdx.x * 4 + warp) * 64 + 2 * lane] = v0;
        out[(blockIdx.x * 4 + warp) * 64 + 2 * lane + 1] = v1;
    }
    if (lane == 0) cyc[blockIdx.x * 4 + warp] = t1 - t0;
    if (MODE == 1) {
        tmem::fence_before_sync();
        __syncthreads();
        if (warp == 0) tmem::dealloc(tbase, 256);
    }
    (void)err;
}

int main() {
    const int grid = 148 * 2;
    double *gS, *o0, *o1;
    long long *c0, *c1;
    cudaMalloc(&gS, N * N * 8); cudaMalloc(&o0, grid * 4 * 64 * 8); cudaMalloc(&o1, grid * 4 * 64 * 8);
    cudaMalloc(&c0, grid * 4 * 8); cudaMalloc(&c1, grid * 4 * 8);
    static double hS[N * N];
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) hS[i * N + j] = ((i * 31 + j * 17) % 97) * 1e-2 - 0.4 + ((i == j) ? 1.0 : 0.0);
    for (int i = 0; i < N; ++i) for (int j = 0; j < i; ++j) hS[i * N + j] = hS[j * N + i];
    cudaMemcpy(gS, hS, sizeof hS, cudaMemcpyHostToDevice);
    const size_t smem = (4 * (N * N + 64) + 4 * 64) * 8;
    cudaFuncSetAttribute(kern<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kern<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int full = 0; full < 2; ++full) {
        const int g = full ? grid : 1;
        cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
        kern<0><<<g, 128, smem>>>(gS, o0, c0, nullptr); kern<1><<<g, 128, smem>>>(gS, o1, c1, nullptr);
        cudaEventRecord(e0); kern<0><<<g, 128, smem>>>(gS, o0, c0, nullptr);
        cudaEventRecord(e1); kern<1><<<g, 128, smem>>>(gS, o1, c1, nullptr);
        cudaEventRecord(e2); cudaDeviceSynchronize();
        float ms0, ms1; cudaEventElapsedTime(&ms0, e0, e1); cudaEventElapsedTime(&ms1, e1, e2);
        static double h0[148 * 2 * 4 * 64], h1[148 * 2 * 4 * 64];
        static long long hc0[148 * 2 * 4], hc1[148 * 2 * 4];
        cudaMemcpy(h0, o0, g * 4 * 64 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(h1, o1, g * 4 * 64 * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(hc0, c0, g * 4 * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hc1, c1, g * 4 * 8, cudaMemcpyDeviceToHost);
        int bad = 0; double mx = 0;
        for (int i = 0; i < g * 4 * 64; ++i) { if ((i % 64) < 50) { if (h0[i] != h1[i]) ++bad; if (fabs(h0[i]) > mx) mx = fabs(h0[i]); } }
        double a0 = 0, a1 = 0; for (int i = 0; i < g * 4; ++i) { a0 += hc0[i]; a1 += hc1[i]; }
        printf("%s: smem %.1f cycles/matvec (%.3f ms)  tmem %.1f cycles/matvec (%.3f ms)  mismatches %d  max|v| %.3e  err %s\n",
               full ? "8 warps/SM x 148 SMs" : "4 warps, 1 SM", a0 / (g * 4) / REPS, ms0, a1 / (g * 4) / REPS, ms1, bad, mx,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

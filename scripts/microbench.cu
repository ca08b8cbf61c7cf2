// microbench.cu -- single-warp latency probes on sm_100a (fp64 dependent chains, 64-bit shuffles, LDS, div, sqrt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o microbench scripts/microbench.cu && ./microbench
#include <cstdio>
#include <cuda_runtime.h>

#define N_ITER 256
__device__ __forceinline__ double shx(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

__global__ void probe(double* out, long long* cyc, double seed, const double* gS) {
    __shared__ __align__(16) double S[2600];
    __shared__ __align__(16) double vb[64];
    const int lane = threadIdx.x;
    for (int i = lane; i < 2600; i += 32) S[i] = gS[i];
    vb[lane] = seed + lane;
    vb[lane + 32] = seed - lane;
    __syncwarp();
    double a = seed + lane, b = seed * 0.5, c = 1.0000001;
    long long t0, t1;
    int k = 0;
    // 0: dependent DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_ITER; ++i) a = fma(a, c, b);
    t1 = clock64();
    cyc[k++] = t1 - t0;
    // 1: dependent DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_ITER; ++i) a = a + b;
    t1 = clock64();
    cyc[k++] = t1 - t0;
    // 2: dependent 64-bit shuffle (xor 1) chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_ITER; ++i) a = shx(a, 1);
    t1 = clock64();
    cyc[k++] = t1 - t0;
    // 3: butterfly all-reduce (5 levels: shuffle + add), dependent chain of N_ITER/8 reductions
    t0 = clock64();
    for (int i = 0; i < N_ITER / 8; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a = a + shx(a, o);
        a = a * 0.03125;
    }
    t1 = clock64();
    cyc[k++] = t1 - t0;
    // 4: three simultaneous butterflies
    {
        double x = a, y = a + 1.0, z = a - 1.0;
        t0 = clock64();
        for (int i = 0; i < N_ITER / 8; ++i) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double tx = shx(x, o), ty = shx(y, o), tz = shx(z, o);
                x = x + tx;
                y = y + ty;
                z = z + tz;
            }
            x = x * 0.03125;
            y = y * 0.03125;
            z = z * 0.03125;
        }
        t1 = clock64();
        cyc[k++] = t1 - t0;
        a = x + y + z;
    }
    // 5: dependent division chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N_ITER / 8; ++i) a = b / (a + 3.0);
    t1 = clock64();
    cyc[k++] = t1 - t0;
    // 6: dependent sqrt chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N_ITER / 8; ++i) a = sqrt(a + 3.0);
    t1 = clock64();
    cyc[k++] = t1 - t0;
    // 7: dependent LDS.64 chain (pointer chasing through shared memory)
    {
        int idx = lane;
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < N_ITER; ++i) idx = ((int)S[idx]) & 1023;
        t1 = clock64();
        cyc[k++] = t1 - t0;
        a += idx;
    }
    // 8: matvec n=50, K=2, current scheme (2 chains per element), repeated 16 times dependent
    {
        double v0 = a * 1e-3, v1 = b;
        const int n = 50;
        t0 = clock64();
        for (int rep = 0; rep < 16; ++rep) {
            vb[lane] = v0;
            if (lane < 18) vb[32 + lane] = v1;
            __syncwarp();
            double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
            const double* row = S + lane;
            for (int j = 0; j + 1 < n; j += 2) {
                const double2 vj = *reinterpret_cast<const double2*>(vb + j);
                a0 = fma(row[0], vj.x, a0);
                b0 = fma(row[32], vj.x, b0);
                a1 = fma(row[n], vj.y, a1);
                b1 = fma(row[n + 32], vj.y, b1);
                row += 2 * n;
            }
            __syncwarp();
            v0 = a0 + a1;
            v1 = b0 + b1;
        }
        t1 = clock64();
        cyc[k++] = t1 - t0;
        a += v0 + v1;
    }
    // 9: matvec n=50 fully unrolled with 4 chains per element (rows j mod 4)
    {
        double v0 = a * 1e-3, v1 = b;
        t0 = clock64();
        for (int rep = 0; rep < 16; ++rep) {
            vb[lane] = v0;
            if (lane < 18) vb[32 + lane] = v1;
            __syncwarp();
            double acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
            const double* row = S + lane;
#pragma unroll
            for (int j = 0; j < 48; j += 4) {
                const double2 va = *reinterpret_cast<const double2*>(vb + j);
                const double2 vc = *reinterpret_cast<const double2*>(vb + j + 2);
                acc0[0] = fma(row[(j + 0) * 50], va.x, acc0[0]);
                acc1[0] = fma(row[(j + 0) * 50 + 32], va.x, acc1[0]);
                acc0[1] = fma(row[(j + 1) * 50], va.y, acc0[1]);
                acc1[1] = fma(row[(j + 1) * 50 + 32], va.y, acc1[1]);
                acc0[2] = fma(row[(j + 2) * 50], vc.x, acc0[2]);
                acc1[2] = fma(row[(j + 2) * 50 + 32], vc.x, acc1[2]);
                acc0[3] = fma(row[(j + 3) * 50], vc.y, acc0[3]);
                acc1[3] = fma(row[(j + 3) * 50 + 32], vc.y, acc1[3]);
            }
            {
                const double2 va = *reinterpret_cast<const double2*>(vb + 48);
                acc0[0] = fma(row[48 * 50], va.x, acc0[0]);
                acc1[0] = fma(row[48 * 50 + 32], va.x, acc1[0]);
                acc0[1] = fma(row[49 * 50], va.y, acc0[1]);
                acc1[1] = fma(row[49 * 50 + 32], va.y, acc1[1]);
            }
            __syncwarp();
            v0 = (acc0[0] + acc0[1]) + (acc0[2] + acc0[3]);
            v1 = (acc1[0] + acc1[1]) + (acc1[2] + acc1[3]);
        }
        t1 = clock64();
        cyc[k++] = t1 - t0;
        a += v0 + v1;
    }
    // 10: empty clock pair
    t0 = clock64();
    t1 = clock64();
    cyc[k++] = t1 - t0;
    out[lane] = a;
}

int main() {
    double *out, *gS;
    long long* cyc;
    cudaMalloc(&out, 32 * 8);
    cudaMalloc(&gS, 2600 * 8);
    cudaMalloc(&cyc, 16 * 8);
    double hS[2600];
    for (int i = 0; i < 2600; ++i) hS[i] = (double)((i * 37) % 1000) * 1e-3;
    cudaMemcpy(gS, hS, sizeof hS, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) probe<<<1, 32>>>(out, cyc, 1.25, gS);
    cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const char* names[] = {"DFMA dependent", "DADD dependent", "SHFL64 xor dependent", "butterfly allreduce (per reduction)",
                           "3 butterflies together (per triple)", "DDIV dependent (+1 add)", "DSQRT dependent (+1 add)",
                           "LDS.64 + cvt dependent", "matvec n=50 2-chain (per matvec)", "matvec n=50 4-chain unrolled (per matvec)",
                           "clock overhead"};
    const double div[] = {N_ITER, N_ITER, N_ITER, N_ITER / 8, N_ITER / 8, N_ITER / 8, N_ITER / 8, N_ITER, 16, 16, 1};
    for (int i = 0; i < 11; ++i) printf("%-45s %8.1f cycles\n", names[i], (double)h[i] / div[i]);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

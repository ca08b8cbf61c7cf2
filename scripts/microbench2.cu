// microbench2.cu -- single-warp S.v (n = 50) variants and shared-memory / fp64 issue rates on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define REPS 16
__global__ void probe(double* out, long long* cyc, double seed, const double* gS) {
    __shared__ __align__(16) double S[50 * 50 + 128];
    __shared__ __align__(16) double vb[64];
    const int lane = threadIdx.x;
    for (int i = lane; i < 2628; i += 32) S[i] = gS[i];
    vb[lane] = 0;
    vb[lane + 32] = 0;
    __syncwarp();
    long long t0, t1;
    int k = 0;
    double acc = 0;
    const int n = 50, ns2 = 25;
    // A: pair layout, unroll 5 (the kernel's loop)
    {
        double v0 = seed, v1 = seed * 0.5;
        t0 = clock64();
        for (int rep = 0; rep < REPS; ++rep) {
            reinterpret_cast<double2*>(vb)[lane] = make_double2(v0, v1);
            __syncwarp();
            double a0x = 0, a0y = 0, a1x = 0, a1y = 0;
            const double2* row = reinterpret_cast<const double2*>(S) + lane;
#pragma unroll 5
            for (int j = 0; j + 1 < n; j += 2) {
                const double2 vj = *reinterpret_cast<const double2*>(vb + j);
                const double2 s0 = row[0], s1 = row[ns2];
                a0x = fma(s0.x, vj.x, a0x);
                a0y = fma(s0.y, vj.x, a0y);
                a1x = fma(s1.x, vj.y, a1x);
                a1y = fma(s1.y, vj.y, a1y);
                row += 2 * ns2;
            }
            __syncwarp();
            v0 = (a0x + a1x) * 1e-2;
            v1 = (a0y + a1y) * 1e-2;
        }
        t1 = clock64();
        cyc[k++] = t1 - t0;
        acc += v0 + v1;
    }
    // B: pair layout, fully unrolled
    {
        double v0 = seed, v1 = seed * 0.5;
        t0 = clock64();
        for (int rep = 0; rep < REPS; ++rep) {
            reinterpret_cast<double2*>(vb)[lane] = make_double2(v0, v1);
            __syncwarp();
            double a0x = 0, a0y = 0, a1x = 0, a1y = 0;
            const double2* row = reinterpret_cast<const double2*>(S) + lane;
#pragma unroll
            for (int j = 0; j < 50; j += 2) {
                const double2 vj = *reinterpret_cast<const double2*>(vb + j);
                const double2 s0 = row[j * ns2], s1 = row[(j + 1) * ns2];
                a0x = fma(s0.x, vj.x, a0x);
                a0y = fma(s0.y, vj.x, a0y);
                a1x = fma(s1.x, vj.y, a1x);
                a1y = fma(s1.y, vj.y, a1y);
            }
            __syncwarp();
            v0 = (a0x + a1x) * 1e-2;
            v1 = (a0y + a1y) * 1e-2;
        }
        t1 = clock64();
        cyc[k++] = t1 - t0;
        acc += v0 + v1;
    }
    // C: explicit two-stage software pipeline (loads of block b+1 issued before the fmas of block b), blocks of 5 row pairs
    {
        double v0 = seed, v1 = seed * 0.5;
        t0 = clock64();
        for (int rep = 0; rep < REPS; ++rep) {
            reinterpret_cast<double2*>(vb)[lane] = make_double2(v0, v1);
            __syncwarp();
            double a0x = 0, a0y = 0, a1x = 0, a1y = 0;
            const double2* row = reinterpret_cast<const double2*>(S) + lane;
            double2 cs0[5], cs1[5], cv[5], ns0[5], ns1[5], nv[5];
#pragma unroll
            for (int u = 0; u < 5; ++u) {
                cv[u] = *reinterpret_cast<const double2*>(vb + 2 * u);
                cs0[u] = row[(2 * u) * ns2];
                cs1[u] = row[(2 * u + 1) * ns2];
            }
#pragma unroll
            for (int b = 0; b < 5; ++b) {
                if (b < 4) {
#pragma unroll
                    for (int u = 0; u < 5; ++u) {
                        nv[u] = *reinterpret_cast<const double2*>(vb + 10 * (b + 1) + 2 * u);
                        ns0[u] = row[(10 * (b + 1) + 2 * u) * ns2];
                        ns1[u] = row[(10 * (b + 1) + 2 * u + 1) * ns2];
                    }
                }
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    a0x = fma(cs0[u].x, cv[u].x, a0x);
                    a0y = fma(cs0[u].y, cv[u].x, a0y);
                    a1x = fma(cs1[u].x, cv[u].y, a1x);
                    a1y = fma(cs1[u].y, cv[u].y, a1y);
                }
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    cv[u] = nv[u];
                    cs0[u] = ns0[u];
                    cs1[u] = ns1[u];
                }
            }
            __syncwarp();
            v0 = (a0x + a1x) * 1e-2;
            v1 = (a0y + a1y) * 1e-2;
        }
        t1 = clock64();
        cyc[k++] = t1 - t0;
        acc += v0 + v1;
    }
    // D: v kept in registers and broadcast by shuffle (no vbuf), S by LDS.128, fully unrolled
    {
        double v0 = seed, v1 = seed * 0.5;
        t0 = clock64();
        for (int rep = 0; rep < REPS; ++rep) {
            double a0x = 0, a0y = 0, a1x = 0, a1y = 0;
            const double2* row = reinterpret_cast<const double2*>(S) + lane;
#pragma unroll
            for (int j = 0; j < 50; j += 2) {
                const double vjx = __shfl_sync(0xffffffffu, v0, j >> 1), vjy = __shfl_sync(0xffffffffu, v1, j >> 1);
                const double2 s0 = row[j * ns2], s1 = row[(j + 1) * ns2];
                a0x = fma(s0.x, vjx, a0x);
                a0y = fma(s0.y, vjx, a0y);
                a1x = fma(s1.x, vjy, a1x);
                a1y = fma(s1.y, vjy, a1y);
            }
            v0 = (a0x + a1x) * 1e-2;
            v1 = (a0y + a1y) * 1e-2;
        }
        t1 = clock64();
        cyc[k++] = t1 - t0;
        acc += v0 + v1;
    }
    // E: 64 independent LDS.128 (throughput), F: 64 independent LDS.64, G: 64 independent DFMA (8 chains)
    {
        const double2* row = reinterpret_cast<const double2*>(S) + lane;
        double2 r[32];
        t0 = clock64();
#pragma unroll
        for (int u = 0; u < 32; ++u) r[u] = row[u * ns2];
        double s = 0;
#pragma unroll
        for (int u = 0; u < 32; ++u) s += r[u].x + r[u].y;
        t1 = clock64();
        cyc[k++] = t1 - t0;
        acc += s;
        const double* rowd = S + lane;
        double rd[32];
        t0 = clock64();
#pragma unroll
        for (int u = 0; u < 32; ++u) rd[u] = rowd[u * 50];
        s = 0;
#pragma unroll
        for (int u = 0; u < 32; ++u) s += rd[u];
        t1 = clock64();
        cyc[k++] = t1 - t0;
        acc += s;
        double c8[8] = {1, 2, 3, 4, 5, 6, 7, 8};
        t0 = clock64();
#pragma unroll
        for (int u = 0; u < 64; ++u) c8[u & 7] = fma(c8[u & 7], 1.0000001, seed);
        t1 = clock64();
        cyc[k++] = t1 - t0;
        for (int u = 0; u < 8; ++u) acc += c8[u];
    }
    // H: one full reduction-heavy "tCG iteration skeleton": matvec(B) + 5 dependent butterflies + 2 divisions + sqrt
    {
        double v0 = seed, v1 = seed * 0.5, z = 1.0;
        t0 = clock64();
        for (int rep = 0; rep < REPS; ++rep) {
            reinterpret_cast<double2*>(vb)[lane] = make_double2(v0, v1);
            __syncwarp();
            double a0x = 0, a0y = 0, a1x = 0, a1y = 0;
            const double2* row = reinterpret_cast<const double2*>(S) + lane;
#pragma unroll
            for (int j = 0; j < 50; j += 2) {
                const double2 vj = *reinterpret_cast<const double2*>(vb + j);
                const double2 s0 = row[j * ns2], s1 = row[(j + 1) * ns2];
                a0x = fma(s0.x, vj.x, a0x);
                a0y = fma(s0.y, vj.x, a0y);
                a1x = fma(s1.x, vj.y, a1x);
                a1y = fma(s1.y, vj.y, a1y);
            }
            __syncwarp();
            double w0 = a0x + a1x, w1 = a0y + a1y;
            for (int r5 = 0; r5 < 5; ++r5) {
                double p = fma(w0, v0, w1 * v1);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) p = p + __shfl_xor_sync(0xffffffffu, p, o);
                w0 = w0 + p * 1e-3;
                w1 = w1 - p * 1e-3;
                if (r5 == 2) z = 1.0 / (p + 3.0);
                if (r5 == 3) z = sqrt(p * p + 1.0) / (z + 2.0);
            }
            v0 = w0 * 1e-2 * z;
            v1 = w1 * 1e-2;
        }
        t1 = clock64();
        cyc[k++] = t1 - t0;
        acc += v0 + v1;
    }
    out[lane] = acc;
}
int main() {
    double *out, *gS;
    long long* cyc;
    cudaMalloc(&out, 32 * 8);
    cudaMalloc(&gS, 2700 * 8);
    cudaMalloc(&cyc, 16 * 8);
    double hS[2700];
    for (int i = 0; i < 2700; ++i) hS[i] = (double)((i * 37) % 1000) * 1e-3;
    cudaMemcpy(gS, hS, sizeof hS, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) probe<<<1, 32>>>(out, cyc, 1.25, gS);
    cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const char* names[] = {"A matvec pair-layout unroll 5", "B matvec pair-layout full unroll", "C matvec explicit sw pipeline",
                           "D matvec v by shuffle", "E 32 independent LDS.128 + sum", "F 32 independent LDS.64 + sum",
                           "G 64 DFMA in 8 chains", "H tCG-iteration skeleton"};
    const double div[] = {REPS, REPS, REPS, REPS, 1, 1, 1, REPS};
    for (int i = 0; i < 8; ++i) printf("%-40s %8.1f cycles\n", names[i], (double)h[i] / div[i]);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

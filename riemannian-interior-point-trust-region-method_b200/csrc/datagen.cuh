// datagen.cuh -- synthetic NonnegPCA sweeps drawn on the device.
//
// The law is the reference generator's (src/NonnegPCA/generator.py:9-65, config_dataset.yaml:6-8):
//   support S: floor(delta n) indices without replacement, v_S = 1/sqrt(|S|)                      (:12-16)
//   Z = sqrt(snr) v v' + N / sqrt(n),  N_ij ~ N(0,1),  diagonal replaced by N(0,1) 2/sqrt(n)      (:19-28; NOT symmetrised)
//   x0 = |u / ||u|||,  u ~ U(0,1)^n ;  y0 = 1                                                      (:46-51, :63)
// The reference draws from NumPy's global, unseeded generator.  Here every number is a pure function of
// (instance id, stream, index): Philox4x32-10 keyed by the instance id, so a sweep is reproducible, any rank can draw
// exactly its own share, and tests/helpers.py restates the generator in NumPy bit for bit (normals by Marsaglia's
// polar method on counter sub-indices; log is det_log, everything else IEEE; no contraction: -fmad=false).
#pragma once
#include "common.cuh"

namespace riptrm {
namespace gen {

enum { STREAM_SUPPORT = 0, STREAM_NOISE = 1, STREAM_DIAG = 2, STREAM_X0 = 3 };  // + point index for x0

struct U4 {
    uint32_t a, b, c, d;
};

__device__ __forceinline__ U4 philox(uint64_t key, uint64_t index, uint32_t stream, uint32_t sub) {
    uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = stream, c3 = sub;
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return U4{c0, c1, c2, c3};
}
// 53-bit uniform in [0, 1)
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double uniform(uint64_t key, uint64_t index, uint32_t stream) {
    const U4 r = philox(key, index, stream, 0);
    return u53(r.a, r.b);
}
__device__ __forceinline__ double normal(uint64_t key, uint64_t index, uint32_t stream) {
    for (uint32_t t = 0; t < 64; ++t) {
        const U4 r = philox(key, index, stream, t);
        const double v1 = 2.0 * u53(r.a, r.b) - 1.0, v2 = 2.0 * u53(r.c, r.d) - 1.0;
        const double s = v1 * v1 + v2 * v2;
        if (s < 1.0 && s > 0.0) return v1 * sqrt((-2.0 * det_log(s)) / s);
    }
    return 0.0;  // probability (1 - pi/4)^64
}

// one CTA per instance
__global__ void nonnegpca_kernel(int n, long long first_instance, int points, double snr, double delta, double* __restrict__ Z,
                                 double* __restrict__ x0, double* __restrict__ y0) {
    extern __shared__ double sh[];  // v [n], u [n], perm (int) [n]
    double* v = sh;
    double* u = sh + n;
    int* perm = reinterpret_cast<int*>(sh + 2 * n);
    const long long inst = first_instance + blockIdx.x;
    const uint64_t key = (uint64_t)inst;
    const int k = (int)floor(delta * (double)n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        v[i] = 0.0;
        perm[i] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // partial Fisher-Yates: the first k entries are a uniformly random k-subset
        for (int i = 0; i < k; ++i) {
            int j = i + (int)(uniform(key, (uint64_t)i, STREAM_SUPPORT) * (double)(n - i));
            if (j > n - 1) j = n - 1;
            const int t = perm[i];
            perm[i] = perm[j];
            perm[j] = t;
        }
        const double val = 1.0 / sqrt((double)k);
        for (int i = 0; i < k; ++i) v[perm[i]] = val;
    }
    __syncthreads();
    const double rs = sqrt(snr), rn = sqrt((double)n);
    double* Zi = Z + (size_t)blockIdx.x * n * n;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e - i * n;
        const double noise = (i == j) ? (normal(key, (uint64_t)i, STREAM_DIAG) * 2.0) / rn : normal(key, (uint64_t)e, STREAM_NOISE) / rn;
        Zi[e] = rs * (v[i] * v[j]) + noise;
    }
    for (int pt = 0; pt < points; ++pt) {
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) u[i] = uniform(key, (uint64_t)i, STREAM_X0 + pt);
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) s = s + u[i] * u[i];
            v[0] = sqrt(s);  // v is no longer needed
        }
        __syncthreads();
        const double nrm = v[0];
        const size_t row = ((size_t)blockIdx.x * points + pt) * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            x0[row + i] = fabs(u[i] / nrm);
            y0[row + i] = 1.0;
        }
    }
}

}  // namespace gen
}  // namespace riptrm

// common.cuh -- warp-level primitives shared by the warp-per-instance RIPTRM kernels.
//
// Arithmetic contract (DESIGN.md "Determinism"): this translation unit is compiled with
// -fmad=false, so `a*b+c` is a rounded multiply followed by a rounded add, exactly like the
// NumPy elementwise expressions of the reference.  Fused multiply-adds appear ONLY where they
// are written explicitly (`fma(...)`): inside dot products and matrix-vector products, whose
// summation order is the fixed tree defined here (lane-strided partial sums, then an xor
// butterfly).  The deterministic CPU oracle (oracle/c/riptrm_det.c) follows the same tree,
// which is what makes iteration traces bit-identical between the two.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace riptrm {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// xor-butterfly sum: every lane ends with the same bits (IEEE add is commutative).
__device__ __forceinline__ double wsum(double p) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) p = p + __shfl_xor_sync(kFull, p, off);
    return p;
}
__device__ __forceinline__ void wsum2(double& a, double& b) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ta = __shfl_xor_sync(kFull, a, off);
        double tb = __shfl_xor_sync(kFull, b, off);
        a = a + ta;
        b = b + tb;
    }
}
__device__ __forceinline__ void wsum3(double& a, double& b, double& c) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ta = __shfl_xor_sync(kFull, a, off);
        double tb = __shfl_xor_sync(kFull, b, off);
        double tc = __shfl_xor_sync(kFull, c, off);
        a = a + ta;
        b = b + tb;
        c = c + tc;
    }
}
template <int N>
__device__ __forceinline__ void wsumN(double (&v)[N]) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double t[N];
#pragma unroll
        for (int i = 0; i < N; ++i) t[i] = __shfl_xor_sync(kFull, v[i], off);
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = v[i] + t[i];
    }
}
__device__ __forceinline__ double wmin(double p) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) p = fmin(p, __shfl_xor_sync(kFull, p, off));
    return p;
}
__device__ __forceinline__ double wmax(double p) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) p = fmax(p, __shfl_xor_sync(kFull, p, off));
    return p;
}
__device__ __forceinline__ bool wall(bool pred) { return __all_sync(kFull, pred); }
__device__ __forceinline__ bool wany(bool pred) { return __any_sync(kFull, pred); }
__device__ __forceinline__ double wbcast(double v, int src) { return __shfl_sync(kFull, v, src); }

// A vector of up to 32*K doubles owned by one warp: element e = k*32 + lane lives in v[k].
// Elements beyond the logical length hold +0.0 at all times (they add exactly 0 to sums).
template <int K>
struct WVec {
    double v[K];
};

template <int K>
__device__ __forceinline__ WVec<K> wzero() {
    WVec<K> r;
#pragma unroll
    for (int k = 0; k < K; ++k) r.v[k] = 0.0;
    return r;
}

// <a,b>: per-lane partial p = a0*b0, then p = fma(ak, bk, p) for k = 1..K-1, then butterfly.
template <int K>
__device__ __forceinline__ double wdot_partial(const WVec<K>& a, const WVec<K>& b) {
    double p = a.v[0] * b.v[0];
#pragma unroll
    for (int k = 1; k < K; ++k) p = fma(a.v[k], b.v[k], p);
    return p;
}
template <int K>
__device__ __forceinline__ double wdot(const WVec<K>& a, const WVec<K>& b) {
    return wsum(wdot_partial(a, b));
}
template <int K>
__device__ __forceinline__ double wsum_vec(const WVec<K>& a) {
    double p = a.v[0];
#pragma unroll
    for (int k = 1; k < K; ++k) p = p + a.v[k];
    return wsum(p);
}

// Deterministic natural logarithm (same operation sequence on the GPU and in the C oracle,
// no fused multiply-adds): the classic argument reduction x = 2^k (1+f), s = f/(2+f),
// log(1+f) = f - hfsq + s (hfsq + R(s^2)) with a degree-14 minimax R (Sun fdlibm's
// published coefficients).  < 1 ulp.  x <= 0 -> -inf / NaN like log().
__host__ __device__ __forceinline__ double det_log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                 Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                 Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    union {
        double d;
        uint64_t u;
    } w;
    w.d = x;
    int64_t k = 0;
    if (x != x) return x;
    if (x < 0.0) {                                 // NaN
        w.u = 0x7ff8000000000000ull;
        return w.d;
    }
    if (x == 0.0) {                                // -inf
        w.u = 0xfff0000000000000ull;
        return w.d;
    }
    if ((w.u >> 52) == 0x7ff) return x;           // +inf
    if ((w.u >> 52) == 0) {                        // subnormal: scale up by 2^54
        w.d = x * 18014398509481984.0;
        k -= 54;
    }
    uint64_t hx = w.u >> 32;
    k += (int64_t)(hx >> 20) - 1023;
    hx &= 0x000fffff;
    uint64_t i = (hx + 0x95f64) & 0x100000;        // normalise to [sqrt(2)/2, sqrt(2))
    w.u = ((hx | (i ^ 0x3ff00000)) << 32) | (w.u & 0xffffffffull);
    k += (int64_t)(i >> 20);
    double f = w.d - 1.0;
    double dk = (double)k;
    double s = f / (2.0 + f);
    double z = s * s;
    double ww = z * z;
    double t1 = ww * (Lg2 + ww * (Lg4 + ww * Lg6));
    double t2 = z * (Lg1 + ww * (Lg3 + ww * (Lg5 + ww * Lg7)));
    double R = t2 + t1;
    double hfsq = 0.5 * f * f;
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

}  // namespace riptrm

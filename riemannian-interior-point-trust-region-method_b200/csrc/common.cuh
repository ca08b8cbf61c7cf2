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
#include <type_traits>

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
// Sums of 4 (8) values over the warp with half the shuffles of independent butterflies: at the xor-16 (and xor-8,
// xor-4) level each lane keeps half of its values and trades the other half, so value i ends up being reduced by
// the 8 (4) lanes whose bits 4,3(,2) spell i, through exactly the butterfly's pairing (l, l^16), (l, l^8), ...;
// a final round of broadcasts hands every total to every lane.  Each total therefore has the same summation tree --
// hence the same bits -- as wsum() of that value; the cost is 20 (30) SHFL.32 instead of 40 (80), which matters
// because a single warp issues one shuffle per ~4 cycles (scripts/microbench.cu).
__device__ __forceinline__ void wsum4x(double (&v)[4]) {
    const int lane = lane_id();
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
    double ka = h16 ? v[2] : v[0], kb = h16 ? v[3] : v[1];
    const double sa = h16 ? v[0] : v[2], sb = h16 ? v[1] : v[3];
    const double ra = __shfl_xor_sync(kFull, sa, 16), rb = __shfl_xor_sync(kFull, sb, 16);
    ka = ka + ra;
    kb = kb + rb;
    double k = h8 ? kb : ka;
    const double s = h8 ? ka : kb;
    k = k + __shfl_xor_sync(kFull, s, 8);
    k = k + __shfl_xor_sync(kFull, k, 4);
    k = k + __shfl_xor_sync(kFull, k, 2);
    k = k + __shfl_xor_sync(kFull, k, 1);
    v[0] = __shfl_sync(kFull, k, 0);
    v[1] = __shfl_sync(kFull, k, 8);
    v[2] = __shfl_sync(kFull, k, 16);
    v[3] = __shfl_sync(kFull, k, 24);
}
// 8 slots; only the first NV totals are broadcast back (the rest are padding the caller set to 0)
template <int NV>
__device__ __forceinline__ void wsum8x(double (&v)[8]) {
    const int lane = lane_id();
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
    double k4[4], s4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        k4[i] = h16 ? v[4 + i] : v[i];
        s4[i] = h16 ? v[i] : v[4 + i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) s4[i] = __shfl_xor_sync(kFull, s4[i], 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) k4[i] = k4[i] + s4[i];
    double k2[2], s2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        k2[i] = h8 ? k4[2 + i] : k4[i];
        s2[i] = h8 ? k4[i] : k4[2 + i];
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) s2[i] = __shfl_xor_sync(kFull, s2[i], 8);
#pragma unroll
    for (int i = 0; i < 2; ++i) k2[i] = k2[i] + s2[i];
    double k = h4 ? k2[1] : k2[0];
    const double s = h4 ? k2[0] : k2[1];
    k = k + __shfl_xor_sync(kFull, s, 4);
    k = k + __shfl_xor_sync(kFull, k, 2);
    k = k + __shfl_xor_sync(kFull, k, 1);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __shfl_sync(kFull, k, ((i & 4) ? 16 : 0) + ((i & 2) ? 8 : 0) + ((i & 1) ? 4 : 0));
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

// Explicit shared-memory accesses through 32-bit shared-space addresses: a pointer that has been through a spilled context
// struct loses its address space and the compiler falls back to generic LD / ST (profiles/r02c_stableid_*: 15 % of the
// instructions of the small-matrix kernel).  "memory" keeps them ordered with respect to __syncwarp().
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
// the same with a compile-time byte offset folded into the instruction's immediate
template <int OFF>
__device__ __forceinline__ double lds_f64_off(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts_f64_off(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "d"(v) : "memory");
}

// 128-bit forms
template <int OFF>
__device__ __forceinline__ double2 lds_v2f64_off(uint32_t a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts_v2f64_off(uint32_t a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "d"(x), "d"(y) : "memory");
}

// compile-time loop: f(std::integral_constant<int, I>) for I = 0 .. N-1 (indices usable as template arguments / immediates)
template <int I, int N, class Fn>
__device__ __forceinline__ void static_for(Fn&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}
// a value the compiler must keep (or spill) instead of recomputing it from %tid at every use
__device__ __forceinline__ uint32_t opaque_u32(uint32_t x) {
    asm volatile("" : "+r"(x));
    return x;
}

// Warp sums of NV (4 or 6) values THROUGH SHARED MEMORY, same summation tree -- hence the same bits -- as wsum() of each value
// (pairing (l, l^16), (l, l^8), (l, l^4), (l, l^2), (l, l^1)), in ~30 instructions instead of the ~65 (SHFL + FSEL + moves) of
// wsum8x / ~50 of wsum4x: every lane stores its NV partials (NV/2 STS.128), lane 4v + g then adds the eight partials of value v
// held by the lanes {g + 4k} in the tree's order (the xor-16, xor-8, xor-4 levels: 8 LDS.64, 7 DADD), two butterfly steps over
// g finish the tree, and the totals go back through shared memory (one STS.64, NV/2 broadcast LDS.128).  The whole-solve
// Sphere kernels are bound by issue slots, not by shared-memory bandwidth.
// Scratch: kRedDoubles doubles per warp, 16-byte aligned: three rows (one per value pair) of 36 double2 -- 32 lanes' partials,
// then the pair's totals -- which puts the rows 16 banks apart (two wavefronts per LDS.64, the minimum for 24 lanes x 8 bytes).
// The three addresses are per-thread constants the caller computes once (`WarpScratch`).
constexpr int kRedRow = 36 * 16;                   // bytes per row
constexpr int kRedTotOff = 32 * 16;                // byte offset of a row's totals
constexpr int kRedDoubles = 3 * kRedRow / 8;       // 216
struct WarpScratch {
    uint32_t wb;   // shared-space address of the warp's scratch: [0, 512) broadcast operand of S.v, then the reduction rows
    uint32_t wl;   // wb + 16 * lane
    uint32_t rd;   // wb + 512 + row(v) + 8 * (v & 1) + 16 * g with v = min(lane / 4, 5), g = lane % 4: this lane's tree reads
    __device__ __forceinline__ void init(const void* base) {
        const int lane = lane_id();
        const int vv = min(lane >> 2, 5), g = lane & 3;
        const uint32_t b = smem_u32(base);
        wb = opaque_u32(b);
        wl = opaque_u32(b + 16u * lane);
        rd = opaque_u32(b + 512u + (uint32_t)((vv >> 1) * kRedRow + (vv & 1) * 8 + g * 16));
    }
};
constexpr int kWarpScratchDoubles = 64 + kRedDoubles;   // 280
template <int NV>
__device__ __forceinline__ void wsum_smem(double (&v)[NV], const WarpScratch& ws) {
    static_assert(NV == 4 || NV == 6, "wsum_smem: 4 or 6 values");
    sts_v2f64_off<512>(ws.wl, v[0], v[1]);
    sts_v2f64_off<512 + kRedRow>(ws.wl, v[2], v[3]);
    if (NV == 6) sts_v2f64_off<512 + 2 * kRedRow>(ws.wl, v[4], v[5]);
    __syncwarp();
    const uint32_t rd = ws.rd;   // with NV = 4 the lanes of values 4, 5 add up stale data nobody reads
    const double p0 = lds_f64_off<0>(rd), p1 = lds_f64_off<64>(rd), p2 = lds_f64_off<128>(rd), p3 = lds_f64_off<192>(rd);
    const double p4 = lds_f64_off<256>(rd), p5 = lds_f64_off<320>(rd), p6 = lds_f64_off<384>(rd), p7 = lds_f64_off<448>(rd);
    const double a0 = p0 + p4, a1 = p1 + p5, a2 = p2 + p6, a3 = p3 + p7;   // lanes i, i ^ 16
    const double b0 = a0 + a2, b1 = a1 + a3;                               // i, i ^ 8
    double c = b0 + b1;                                                    // i, i ^ 4
    c = c + __shfl_xor_sync(kFull, c, 2);
    c = c + __shfl_xor_sync(kFull, c, 1);
    if ((lane_id() & 3) == 0) sts_f64_off<kRedTotOff>(rd, c);
    __syncwarp();
    const double2 t0 = lds_v2f64_off<512 + kRedTotOff>(ws.wb), t1 = lds_v2f64_off<512 + kRedRow + kRedTotOff>(ws.wb);
    v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
    if (NV == 6) {
        const double2 t2 = lds_v2f64_off<512 + 2 * kRedRow + kRedTotOff>(ws.wb);
        v[4] = t2.x; v[5] = t2.y;
    }
}

__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

}  // namespace riptrm

// fam_columns.cuh -- NonnegPCA with ONE large data matrix shared by p unit-norm columns
// (RIPTRM_FAMILY_NONNEGPCA_COLUMNS; BASELINE config 4, n = 20000, p = 10).
//
// Reading of config 4 (SURVEY.md fact 11 / App. A.4 "Oblique / multi-start"): X in R^{n x p}, every column a
// point of Sphere(n) with the constraints x_ic + eps >= 0, f = -tr(X'ZX).  The problem decouples into p
// reference-exact Sphere problems (src/NonnegPCA/coordinator.py:37-95) that share Z, so per column the
// formulas are those of fam_sphere.cuh, while the expensive operator is ONE dense contraction S.V,
// S = Z + Z' (n x n), V = [delta_1 .. delta_p]: the HBM-bound Hessian-vector product.
//
// Kernel shape: a persistent cooperative grid, one CTA of 9 warps per SM.
//   * S.V pass ("stream"): S is cut into tiles of TJ=32 rows x TW=128 columns (32 KB).  The tiles, linearised
//     (column block major, row block minor), are dealt to the CTAs in equal contiguous ranges (stream-K), so
//     148 SMs stay equally loaded for any n.  Warp 8 is the producer: its 32 lanes issue one 1 KB bulk
//     async copy (TMA, cp.async.bulk -> SASS UBLKCP) per tile row plus one for the 32 x p slice of V into a
//     5-stage shared-memory ring guarded by full/empty mbarriers.  Warps 0..7 consume: S is symmetric, so
//     (S V)[i,:] = sum_j S[j,i] V[j,:] and a warp reads row j of the tile with lanes along i (conflict-free
//     LDS.128) while V[j,0..p) is a shared-memory broadcast; each thread owns 4 columns i x p accumulators
//     (40 DFMA per 7 LDS.128).  At a column-block boundary the 8 warps' accumulators are summed in warp
//     order and written as a partial; the owner of each row later adds the partials in CTA order.
//   * every other step of the tCG iteration is a "vector phase" over the n x p arrays (L2 resident, 1.6 MB
//     each), rows dealt to CTAs in contiguous chunks, with per-column dot products reduced per CTA in a
//     fixed order and then over CTAs in CTA order.  Phases are separated by grid-wide barriers; all CTAs
//     compute the per-column scalars redundantly and bit-identically, so control flow stays uniform.
// All sums have a fixed order for a given grid size: results are deterministic run to run.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "solver_warp.cuh"

namespace riptrm {
namespace col {

namespace cg = cooperative_groups;

constexpr int TW = 128;     // tile width  (columns i of S per tile; 4 per consumer lane)
constexpr int TJ = 32;      // tile height (rows j of S per tile; one bulk copy per row)
constexpr int STAGES = 5;   // shared-memory ring depth (5 x 32 KB in flight per SM)
constexpr int NCW = 8;      // consumer warps
constexpr int NT = (NCW + 1) * 32;
constexpr int MAXP = 16;
constexpr int MAXQ = 4;     // dot products per phase and column

struct Params {
    int n, n_pad, ld, p;           // ld: leading dimension of S (multiple of TW); n_pad: multiple of TJ
    int NIB, NJT;                  // column blocks, row blocks
    long long total_tiles;
    int R;                         // rows of the n x p arrays owned by one CTA
    int slots;                     // column blocks one CTA's tile range can touch
    const double* S;               // [NIB * NJT tiles][TJ][TW]: S = Z + Z' in streaming order, zero padded
    double eps;
    int embedded;
    // n_pad x p arrays (row-major, rows >= n are zero and never written)
    double *X, *Y, *Sx, *ys, *c, *V, *Sv, *t, *Hd, *eta, *Heta, *r, *eta2, *Heta2, *r2;
    const long long* tbeg;         // [grid + 1] first tile of each CTA's range
    const int* tib0;               // [grid] tbeg[g] / NJT
    double* mv_part;               // [grid][slots][TW][p]
    double* dot_part;              // [2][grid][MAXQ * MAXP]
    // hook arguments
    double mu, Delta;
    const double* vin;             // hessvec operand [n][p]
    double* out;                   // [n][p]
    double* info;                  // [p][4]
    unsigned long long* passes;    // number of S.V passes executed (for the roofline arithmetic)
    // tCG options
    int tcg_mininner, tcg_maxinner;
    double tcg_theta, tcg_kappa;
};

// ------------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + bulk async copy (TMA)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory"); }

// ------------------------------------------------------------------------------------------------------
// shared memory
// ------------------------------------------------------------------------------------------------------
template <int P>
struct Smem {
    alignas(128) double S[STAGES][TJ][TW];
    alignas(16) double V[STAGES][TJ][P];
    alignas(16) double flush[TW][P];
    double redv[MAXQ][NT];
    double scal[MAXQ * P];
    alignas(8) uint64_t full[STAGES];
    alignas(8) uint64_t empty[STAGES];
};

struct Pipe {
    int stage;
    uint32_t phase;
    __device__ __forceinline__ void advance() {
        if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
        }
    }
};

// ------------------------------------------------------------------------------------------------------
// One S.V pass: mv_part <- this CTA's partial sums of S V over its tile range
// ------------------------------------------------------------------------------------------------------
template <int P>
__device__ __forceinline__ void stream_pass(const Params& prm, Smem<P>& sm, Pipe& pipe) {
    const int g = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = prm.tbeg[g], t1 = prm.tbeg[g + 1];
    const uint32_t tx_bytes = (uint32_t)(TJ * TW * sizeof(double) + TJ * P * sizeof(double));
    if (warp == NCW) {
        // ---- producer: lane r copies tile row r; lane 0 also arms the barrier and copies the V slice
        int jt = (int)(t0 % prm.NJT);
        for (long long t = t0; t < t1; ++t, jt = (jt + 1 == prm.NJT) ? 0 : jt + 1) {
            mbar_wait(&sm.empty[pipe.stage], pipe.phase ^ 1u);
            if (lane == 0) {
                // S is stored tile by tile in streaming order: one 32 KB bulk copy per tile
                mbar_expect_tx(&sm.full[pipe.stage], tx_bytes);
                bulk_g2s(&sm.S[pipe.stage][0][0], prm.S + (size_t)t * (TJ * TW), (uint32_t)(TJ * TW * sizeof(double)),
                         &sm.full[pipe.stage]);
                bulk_g2s(&sm.V[pipe.stage][0][0], prm.V + (size_t)jt * TJ * P, (uint32_t)(TJ * P * sizeof(double)),
                         &sm.full[pipe.stage]);
            }
            __syncwarp();
            pipe.advance();
        }
    } else {
        double acc[4][P];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int c = 0; c < P; ++c) acc[q][c] = 0.0;
        const int ib0 = prm.tib0[g];
        int cur_ib = ib0;
        int jt = (int)(t0 - (long long)ib0 * prm.NJT), ibn = ib0;  // (ibn, jt): block coordinates of tile t
        for (long long t = t0; t <= t1; ++t) {
            const int ib = (t < t1) ? ibn : -1;
            if (++jt == prm.NJT) {
                jt = 0;
                ++ibn;
            }
            if (t > t0 && ib != cur_ib) {
                // ---- column-block boundary: sum the 8 warps' accumulators in warp order, write the partial
                for (int w = 0; w < NCW; ++w) {
                    if (warp == w) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int col = (q >> 1) * 64 + 2 * lane + (q & 1);
#pragma unroll
                            for (int c = 0; c < P; ++c)
                                sm.flush[col][c] = (w == 0) ? acc[q][c] : (sm.flush[col][c] + acc[q][c]);
                        }
                    }
                    consumer_bar();
                }
                double* dst = prm.mv_part + ((size_t)g * prm.slots + (cur_ib - ib0)) * (TW * P);
                const double* src = &sm.flush[0][0];
                for (int e = threadIdx.x; e < TW * P; e += NCW * 32) dst[e] = src[e];
                consumer_bar();
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int c = 0; c < P; ++c) acc[q][c] = 0.0;
                cur_ib = ib;
            }
            if (t == t1) break;
            mbar_wait(&sm.full[pipe.stage], pipe.phase);
#pragma unroll
            for (int rr = 0; rr < TJ / NCW; ++rr) {
                const int r = warp + rr * NCW;
                const double2 s01 = reinterpret_cast<const double2*>(&sm.S[pipe.stage][r][0])[lane];
                const double2 s23 = reinterpret_cast<const double2*>(&sm.S[pipe.stage][r][64])[lane];
                const double* v = &sm.V[pipe.stage][r][0];
#pragma unroll
                for (int c = 0; c < P; ++c) {
                    const double vc = v[c];
                    acc[0][c] = fma(s01.x, vc, acc[0][c]);
                    acc[1][c] = fma(s01.y, vc, acc[1][c]);
                    acc[2][c] = fma(s23.x, vc, acc[2][c]);
                    acc[3][c] = fma(s23.y, vc, acc[3][c]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[pipe.stage]);
            pipe.advance();
        }
    }
    if (threadIdx.x == 0 && g == 0) atomicAdd(prm.passes, 1ull);
}

// (S V)[row, c]: partials of the CTAs whose tile ranges meet the row's column block, added in CTA order.
// tbeg[g] = first tile of CTA g (host-computed, G + 1 entries), tib0[g] = tbeg[g] / NJT.
__device__ __forceinline__ double gather_elem(const Params& prm, int row, int c, int P) {
    const int G = gridDim.x;
    const int ib = row / TW, ii = row - ib * TW;
    const long long tb = (long long)ib * prm.NJT, te = tb + prm.NJT;  // tiles of this column block
    int g = (int)(((double)tb * G) / (double)prm.total_tiles);
    g = min(max(g, 0), G - 1);
    while (g > 0 && prm.tbeg[g] > tb) --g;
    while (g < G - 1 && prm.tbeg[g + 1] <= tb) ++g;
    double out = 0.0;
    bool first = true;
    for (; g < G; ++g) {
        const long long b = prm.tbeg[g], e = prm.tbeg[g + 1];
        if (b >= te) break;
        if (e <= b) continue;
        const double v = prm.mv_part[(((size_t)g * prm.slots + (ib - prm.tib0[g])) * TW + ii) * P + c];
        out = first ? v : (out + v);
        first = false;
    }
    return out;
}

// ------------------------------------------------------------------------------------------------------
// Vector phases: thread `tid` < NTV = (NT / P) * P owns the elements e = ebeg + tid + k * NTV of the CTA's
// row chunk, all of column c = tid % P.  Per-thread partial dot products are summed per column in thread
// order, written per CTA, and after the grid barrier summed over CTAs (lane-strided, then the xor butterfly).
// ------------------------------------------------------------------------------------------------------
template <int P, int Q>
__device__ __forceinline__ void block_reduce_store(const Params& prm, Smem<P>& sm, const double (&part)[Q], int buf) {
    constexpr int NTV = (NT / P) * P;
#pragma unroll
    for (int q = 0; q < Q; ++q) sm.redv[q][threadIdx.x] = part[q];
    __syncthreads();
    if (threadIdx.x < Q * P) {
        const int q = threadIdx.x / P, c = threadIdx.x - q * P;
        double s = sm.redv[q][c];
        for (int i = c + P; i < NTV; i += P) s = s + sm.redv[q][i];
        prm.dot_part[((size_t)buf * gridDim.x + blockIdx.x) * (MAXQ * MAXP) + threadIdx.x] = s;
    }
}

template <int P, int Q>
__device__ __forceinline__ void gather_scalars(const Params& prm, Smem<P>& sm, int buf) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x;
    for (int k = warp; k < Q * P; k += NCW + 1) {
        double s = 0.0;
        for (int g = lane; g < G; g += 32) s = s + prm.dot_part[((size_t)buf * G + g) * (MAXQ * MAXP) + k];
        s = wsum(s);
        if (lane == 0) sm.scal[k] = s;
    }
    __syncthreads();
}

// per-column tCG state, identical in every CTA
struct ColState {
    double e_Pe, e_Pd, d_Pd, z_r, r_r, norm_r0, nr_theta, target, model_value, alpha, beta, kappa, xSx, a, b, d;
    int done, iters, stop;
};

// ------------------------------------------------------------------------------------------------------
// the cooperative kernel.  MODE 1: out = Hw[vin] at (X, Y, mu).  MODE 2: one tCG solve per column at
// (X, Y, mu, Delta) -> out = eta, info.  MODE 3: bare S.V passes (diagnostic).
// ------------------------------------------------------------------------------------------------------
template <int P, int MODE>
__global__ void __launch_bounds__(NT, 1) columns_kernel(Params prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<P>& sm = *reinterpret_cast<Smem<P>*>(smem_raw);
    __shared__ ColState cs[P];
    cg::grid_group grid = cg::this_grid();
    constexpr int NTV = (NT / P) * P;
    const int g = blockIdx.x;
    const int tid = threadIdx.x;
    const int n = prm.n;
    const int row_lo = min(n, g * prm.R), row_hi = min(n, row_lo + prm.R);
    const size_t ebeg = (size_t)row_lo * P + tid, eend = (size_t)row_hi * P;
    const bool vthread = tid < NTV;
    const int myc = tid % P;
#define FOR_ELEMS(e) for (size_t e = ebeg; vthread && e < eend; e += NTV)

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    Pipe pipe{0, 0u};
    int buf = 0;
    if (MODE == 3) {
        // diagnostic: tcg_maxinner bare S.V passes (no vector phases) -- the streaming ceiling of this kernel
        for (int j = 0; j < prm.tcg_maxinner; ++j) {
            stream_pass<P>(prm, sm, pipe);
            grid.sync();
        }
        return;
    }

    // ---- point cache (eval_point + begin_step of fam_sphere.cuh), operand V := X -----------------------
    FOR_ELEMS(e) prm.V[e] = prm.X[e];
    fence_proxy_async();
    grid.sync();
    stream_pass<P>(prm, sm, pipe);
    grid.sync();
    {
        double part[3] = {0.0, 0.0, 0.0};
        FOR_ELEMS(e) {
            const double sx = gather_elem(prm, (int)(e / P), myc, P);
            const double x = prm.X[e], y = prm.Y[e];
            const double s = x + prm.eps;
            const double w = prm.mu * (1.0 / s);
            prm.Sx[e] = sx;
            prm.ys[e] = y / s;
            part[0] = fma(x, sx, part[0]);  // x'Sx
            part[1] = fma(x, w, part[1]);   // <x, mu/s>
            part[2] = fma(y, x, part[2]);   // y'x
        }
        block_reduce_store<P, 3>(prm, sm, part, buf);
    }
    grid.sync();
    gather_scalars<P, 3>(prm, sm, buf);
    buf ^= 1;
    if (tid < P) {
        cs[tid].xSx = sm.scal[0 * P + tid];
        cs[tid].kappa = sm.scal[0 * P + tid] + sm.scal[2 * P + tid];
        cs[tid].a = sm.scal[1 * P + tid];  // xw, used just below
        cs[tid].done = (tid >= prm.p) ? 1 : 0;  // padding columns (P > p) never run
        cs[tid].iters = 0;
        cs[tid].stop = RIPTRM_TCG_MAX_INNER_ITER;
    }
    __syncthreads();
    {
        // c = grad f - G_x(mu/s);  r = c, delta = -c, eta = Heta = 0;  r_r = <c, c>
        double part[1] = {0.0};
        const double xSx = cs[myc].xSx, xw = cs[myc].a;
        FOR_ELEMS(e) {
            const double x = prm.X[e];
            const double s = x + prm.eps;
            const double w = prm.mu * (1.0 / s);
            const double gradf = -prm.Sx[e] + xSx * x;
            const double Gw = w - xw * x;
            const double cc = gradf - Gw;
            prm.c[e] = cc;
            if (MODE == 2) {
                prm.r[e] = cc;
                prm.V[e] = -cc;
                prm.eta[e] = 0.0;
                prm.Heta[e] = 0.0;
                part[0] = fma(cc, cc, part[0]);
            } else {
                prm.V[e] = prm.vin[e];
            }
        }
        if (MODE == 2) block_reduce_store<P, 1>(prm, sm, part, buf);
    }
    fence_proxy_async();
    grid.sync();
    const double Delta2 = prm.Delta * prm.Delta;
    int maxinner = prm.tcg_maxinner < 0 ? (n - 1) : prm.tcg_maxinner;
    if (MODE == 2) {
        gather_scalars<P, 1>(prm, sm, buf);
        buf ^= 1;
        if (tid < P) {
            ColState& s = cs[tid];
            s.r_r = sm.scal[tid];
            s.norm_r0 = sqrt(s.r_r);
            s.z_r = s.r_r;
            s.d_Pd = s.r_r;
            s.e_Pe = 0.0;
            s.e_Pd = 0.0;
            s.model_value = 0.0;
            s.nr_theta = (prm.tcg_theta == 1.0) ? s.norm_r0 : pow(s.norm_r0, prm.tcg_theta);
            s.target = s.norm_r0 * fmin(s.nr_theta, prm.tcg_kappa);
        }
        __syncthreads();
    } else {
        maxinner = 1;
    }

    for (int j = 0; j < maxinner; ++j) {
        // ---- T1: Sv = S V ------------------------------------------------------------------------------
        stream_pass<P>(prm, sm, pipe);
        grid.sync();
        // ---- T2: a = <x, Sv>, b = <x, v> ---------------------------------------------------------------
        {
            double part[2] = {0.0, 0.0};
            FOR_ELEMS(e) {
                const double sv = gather_elem(prm, (int)(e / P), myc, P);
                const double x = prm.X[e];
                prm.Sv[e] = sv;
                part[0] = fma(x, sv, part[0]);
                part[1] = fma(x, prm.V[e], part[1]);
            }
            block_reduce_store<P, 2>(prm, sm, part, buf);
        }
        grid.sync();
        gather_scalars<P, 2>(prm, sm, buf);
        buf ^= 1;
        if (tid < P) {
            cs[tid].a = sm.scal[tid];
            cs[tid].b = sm.scal[P + tid];
        }
        __syncthreads();
        // ---- T3: t = (y/s) * G*[v], d = <x, t> ---------------------------------------------------------
        {
            double part[1] = {0.0};
            const double b = cs[myc].b;
            FOR_ELEMS(e) {
                const double x = prm.X[e], v = prm.V[e];
                const double ga = prm.embedded ? v : (v - x * b);
                const double tt = prm.ys[e] * ga;
                prm.t[e] = tt;
                part[0] = fma(x, tt, part[0]);
            }
            block_reduce_store<P, 1>(prm, sm, part, buf);
        }
        grid.sync();
        gather_scalars<P, 1>(prm, sm, buf);
        buf ^= 1;
        if (tid < P) cs[tid].d = sm.scal[tid];
        __syncthreads();
        // ---- T4: Hd = Hw[v];  d_Hd = <v, Hd> -----------------------------------------------------------
        {
            double part[1] = {0.0};
            const double a = cs[myc].a, d = cs[myc].d, kappa = cs[myc].kappa;
            FOR_ELEMS(e) {
                const double x = prm.X[e], v = prm.V[e];
                const double hl = (-prm.Sv[e] + a * x) + kappa * v;
                const double gg = prm.t[e] - d * x;
                const double hd = hl + gg;
                if (MODE == 1) {
                    prm.out[e] = hd;
                } else {
                    prm.Hd[e] = hd;
                    part[0] = fma(v, hd, part[0]);
                }
            }
            if (MODE == 2) block_reduce_store<P, 1>(prm, sm, part, buf);
        }
        if (MODE == 1) break;
        grid.sync();
        gather_scalars<P, 1>(prm, sm, buf);
        buf ^= 1;
        // ---- T5: step length / boundary exit; tentative eta, Heta, r; three dot products -----------------
        if (tid < P && !cs[tid].done) {
            ColState& s = cs[tid];
            const double d_Hd = sm.scal[tid];
            s.alpha = 0.0;
            double e_Pe_new = s.e_Pe;
            if (d_Hd != 0.0) {
                s.alpha = s.z_r / d_Hd;
                e_Pe_new = (s.e_Pe + (2.0 * s.alpha) * s.e_Pd) + (s.alpha * s.alpha) * s.d_Pd;
            }
            s.iters = j + 1;
            if (d_Hd <= 0.0 || e_Pe_new >= Delta2) {
                // tau solve (RIPTRM.py:123-125); alpha carries tau into the update below
                s.alpha = (-s.e_Pd + sqrt(s.e_Pd * s.e_Pd + s.d_Pd * (Delta2 - s.e_Pe))) / s.d_Pd;
                s.stop = (d_Hd <= 0.0) ? RIPTRM_TCG_NEGATIVE_CURVATURE : RIPTRM_TCG_EXCEEDED_TR;
                s.done = 2;  // boundary exit: commit eta + tau delta in this phase
            } else {
                s.e_Pe = e_Pe_new;
            }
        }
        __syncthreads();
        {
            double part[3] = {0.0, 0.0, 0.0};
            const int done = cs[myc].done;
            const double al = cs[myc].alpha;
            if (done != 1) {
                FOR_ELEMS(e) {
                    const double hd = prm.Hd[e];
                    const double ne = prm.eta[e] + al * prm.V[e];
                    const double nh = prm.Heta[e] + al * hd;
                    if (done == 2) {
                        prm.eta[e] = ne;
                        prm.Heta[e] = nh;
                    } else {
                        const double nr = prm.r[e] + al * hd;
                        prm.eta2[e] = ne;
                        prm.Heta2[e] = nh;
                        prm.r2[e] = nr;
                        part[0] = fma(ne, prm.c[e], part[0]);
                        part[1] = fma(ne, nh, part[1]);
                        part[2] = fma(nr, nr, part[2]);
                    }
                }
            }
            block_reduce_store<P, 3>(prm, sm, part, buf);
        }
        grid.sync();
        gather_scalars<P, 3>(prm, sm, buf);
        buf ^= 1;
        // ---- T6: model test, residual test, beta; delta_new = -r + beta delta; <x, delta_new> -------------
        if (tid < P) {
            ColState& s = cs[tid];
            if (s.done == 2) {
                s.done = 1;
            } else if (!s.done) {
                const double new_model = sm.scal[tid] + 0.5 * sm.scal[P + tid];
                if (new_model >= s.model_value) {
                    s.stop = RIPTRM_TCG_MODEL_INCREASED;
                    s.done = 1;
                } else {
                    s.model_value = new_model;
                    s.r_r = sm.scal[2 * P + tid];
                    const double norm_r = sqrt(s.r_r);
                    s.done = 3;  // commit eta2/Heta2/r2 in this phase
                    if (j >= prm.tcg_mininner && norm_r <= s.target) {
                        s.stop = (prm.tcg_kappa < s.nr_theta) ? RIPTRM_TCG_REACHED_TARGET_LINEAR
                                                               : RIPTRM_TCG_REACHED_TARGET_SUPERLINEAR;
                        s.done = 4;  // commit, then finished
                    } else {
                        const double zold = s.z_r;
                        s.z_r = s.r_r;
                        s.beta = s.z_r / zold;
                    }
                }
            }
        }
        __syncthreads();
        {
            double part[1] = {0.0};
            const int done = cs[myc].done;
            const double beta = cs[myc].beta;
            if (done >= 3) {
                FOR_ELEMS(e) {
                    const double rr = prm.r2[e];
                    prm.eta[e] = prm.eta2[e];
                    prm.Heta[e] = prm.Heta2[e];
                    prm.r[e] = rr;
                    if (done == 3) {
                        const double dn = -rr + beta * prm.V[e];
                        prm.t[e] = dn;  // pre-projection delta
                        part[0] = fma(prm.X[e], dn, part[0]);
                    }
                }
            }
            block_reduce_store<P, 1>(prm, sm, part, buf);
        }
        grid.sync();
        gather_scalars<P, 1>(prm, sm, buf);
        buf ^= 1;
        // ---- T7: delta = P_x(delta_new); conjugacy scalars ------------------------------------------------
        if (cs[myc].done == 3) {
            const double xd = sm.scal[myc];
            FOR_ELEMS(e) prm.V[e] = prm.t[e] - xd * prm.X[e];
        }
        __syncthreads();
        if (tid < P) {
            ColState& s = cs[tid];
            if (s.done == 3) {
                s.e_Pd = s.beta * (s.e_Pd + s.alpha * s.d_Pd);
                s.d_Pd = s.z_r + (s.beta * s.beta) * s.d_Pd;
                s.done = 0;
            } else if (s.done == 4) {
                s.done = 1;
            }
        }
        __syncthreads();
        bool all_done = true;
#pragma unroll
        for (int c = 0; c < P; ++c) all_done = all_done && (cs[c].done == 1);
        fence_proxy_async();
        grid.sync();
        if (all_done) break;
    }

    if (MODE == 2) {
        // ---- outputs: eta, ||eta||, info ---------------------------------------------------------------------
        double part[1] = {0.0};
        FOR_ELEMS(e) {
            const double v = prm.eta[e];
            prm.out[e] = v;
            part[0] = fma(v, v, part[0]);
        }
        block_reduce_store<P, 1>(prm, sm, part, buf);
        grid.sync();
        gather_scalars<P, 1>(prm, sm, buf);
        if (g == 0 && tid < P && tid < prm.p && prm.info != nullptr) {
            prm.info[tid * 4 + 0] = (double)cs[tid].iters;
            prm.info[tid * 4 + 1] = (double)cs[tid].stop;
            prm.info[tid * 4 + 2] = sqrt(sm.scal[tid]);
            prm.info[tid * 4 + 3] = cs[tid].model_value;
        }
    }
#undef FOR_ELEMS
}

// S = Z + Z' into the streaming layout: tile (ib, jt) = rows [jt*TJ, +TJ) x columns [ib*TW, +TW) stored
// contiguously ([TJ][TW] row-major) at tile index ib * NJT + jt; padding stays zero.
__global__ void build_S_kernel(const double* __restrict__ Z, double* __restrict__ S, int n, int NJT) {
    __shared__ double tile[32][33];
    const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int k = ty; k < 32; k += 8) {
        const int i = bj + k, j = bi + tx;  // tile[k][tx] = Z[bj + k][bi + tx]
        tile[k][tx] = (i < n && j < n) ? Z[(size_t)i * n + j] : 0.0;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int i = bi + k, j = bj + tx;
        if (i < n && j < n) {
            const size_t tile_id = (size_t)(j / TW) * NJT + (i / TJ);
            S[(tile_id * TJ + (i % TJ)) * TW + (j % TW)] = Z[(size_t)i * n + j] + tile[tx][k];
        }
    }
}

}  // namespace col
}  // namespace riptrm

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
//   * S.V pass ("stream"): S is cut into tiles of TJ=32 rows x TW=128 columns (32 KB), stored tile by tile in
//     streaming order.  The tiles, linearised (column block major, row block minor), are dealt to the CTAs in
//     equal contiguous ranges (stream-K), so 148 SMs stay equally loaded for any n.  Warp 8 is the producer: one
//     32 KB bulk async copy (TMA, cp.async.bulk -> SASS UBLKCP) per tile plus one for the 32 x p slice of V into a
//     5-stage shared-memory ring guarded by full/empty mbarriers.  Warps 0..7 consume.  S is symmetric, so
//     (S V)[i,:] = sum_j S[j,i] V[j,:].
//       P >= 8: FP64 tensor cores, mma.sync.m8n8k4.f64 (SASS DMMA) with M <-> i, K <-> j, N <-> column of V; inside a
//       tile S is stored in A-fragment order, so a fragment is one coalesced LDS.64; a warp owns 16 columns i of the
//       tile and at most 8 accumulator doubles per thread.
//       P = 1, 2, 4: DFMA; a warp reads row j of the tile with lanes along i (conflict-free LDS.128) while V[j,0..p)
//       is a shared-memory broadcast; each thread owns 4 columns i x p accumulators, the 8 warps' accumulators are
//       summed in warp order at a column-block boundary.
//     The partial of a column block is written per CTA; the owner of each row later adds the partials in CTA order.
//   * every other step of the tCG iteration is a "vector phase" over the n x p arrays (L2 resident, 1.6 MB
//     each), rows dealt to CTAs in contiguous chunks, with per-column dot products reduced per CTA in a
//     fixed order and then over CTAs in CTA order.  Phases are separated by grid-wide barriers (four per tCG
//     iteration, thanks to the merged reductions of fam_sphere.cuh); all CTAs compute the per-column scalars
//     redundantly and bit-identically, so control flow stays uniform.
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
// FP64 tensor-core (DMMA) consumers for P >= 8 (N = 8 or 16, padded); plain DFMA consumers for P = 1, 2, 4
template <int P>
constexpr bool kTensor = (P >= 8);
constexpr int MAXQ = 10;    // reductions per phase and column (sums first, then mins)

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
    // whole-solve mode (riptrm_solve on this family): per-column state and the solver options
    int solve;                     // 1: mu / Delta / activity per column come from colstate, the point cache is valid
    double* colstate;              // [MAXP][CS_FIELDS]
    double *XN, *YN, *SxN, *Xinit, *Yinit, *Sxinit, *Xprev;
    const double *mu_sched, *tolL_sched, *tolC_sched;
    int maxiter, inner_maxiter, trace_mode, trace_capacity;
    double tolresid, initial_tr_radius, minimal_initial_tr_radius, maximal_tr_radius, rho, reduction_regularization,
        gamma, const_left, const_right;
    double* trace;                 // [p][trace_capacity][RIPTRM_TRACE_FIELDS] or nullptr
    double* summary;               // [p][RIPTRM_SUMMARY_FIELDS]
    int* all_done;                 // device flag read by the host loop
    // time limits (RIPTRM.py:822-834, base_solver.py:85-106): the host loop measures the time since the start of the solve
    // when it polls `all_done` and passes it in, so every CTA takes the same branch
    double now_s, maxtime, inner_maxtime;   // inner_maxtime < 0: None (the outer limit applies to the inner loop)
    // device-side loop (riptrm_api.cu columns_solve_graph): the launches of a whole solve are the body of a conditional WHILE
    // graph node; seconds since the start come from a device clock stamp written ahead of every iteration, and the post
    // kernel ends the loop
    int reuse_heta;                // COLUMNS whole solve: the Hw[dx] of RIPTRM.py:659 is the Hw[eta] the tCG accumulated (one S.V
                                   // pass and three reduction rounds fewer per trust-region iteration; solver_warp.cuh
                                   // inner_step has the argument); 0 = a fresh product (RIPTRM_RECOMPUTE_HDX=1)
    const double* now_ptr;         // nullptr: `now_s` above (host clock at enqueue time)
    unsigned long long cond;       // cudaGraphConditionalHandle
    int cond_on;
};

// seconds since the start of the solve, device clock: written by one thread ahead of every trust-region iteration so that every
// CTA of the following launches reads the same value
__global__ void stamp_kernel(double* now_s, unsigned long long* t0, int init) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (init) {
        *t0 = t;
        *now_s = 0.0;
    } else {
        *now_s = (double)(t - *t0) * 1e-9;
    }
}

// per-column solver state kept in global memory between launches
enum {
    CS_IT = 0, CS_K, CS_DELTA, CS_XSX, CS_COST, CS_CNT_INNER, CS_CNT_TCG, CS_CNT_AUX, CS_ROWS, CS_FINISHED, CS_STOP,
    CS_DELTA_INIT, CS_XSX_INIT, CS_COST_INIT, CS_TCG_ITERS, CS_TCG_STOP, CS_KAPPA, CS_T_INNER, CS_FIELDS = 24
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
    // who contributes partials to the column blocks this CTA's rows fall into (built once per launch)
    unsigned long long gt_base[4][8];
    int gt_n[4];
    int gt_ib_first, gt_ok;
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
    } else if (kTensor<P>) {
        // ---- FP64 tensor-core consumers (mma.sync.m8n8k4.f64, SASS DMMA): out[i, c] = sum_j S[j][i] V[j][c] with
        //      M <-> i (8), K <-> j (4), N <-> c (8).  Warp w owns the column groups ig = 2w, 2w+1 of the tile; per row
        //      group jg it loads two A fragments (one coalesced LDS.64 each: S is stored in fragment order) and two B
        //      fragments (V slice, row-major; columns >= P masked) and issues four DMMAs.  At most 8 accumulator
        //      doubles per thread instead of 4 P.
        constexpr int NCG = (P + 7) / 8;  // column groups of V (N tiles)
        double acc[2][NCG][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < NCG; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        const int ib0 = prm.tib0[g];
        int cur_ib = ib0;
        int jt = (int)(t0 - (long long)ib0 * prm.NJT), ibn = ib0;
        for (long long t = t0; t <= t1; ++t) {
            const int ib = (t < t1) ? ibn : -1;
            if (++jt == prm.NJT) {
                jt = 0;
                ++ibn;
            }
            if (t > t0 && ib != cur_ib) {
                // accumulator (a, b, h): i = 8 (2 warp + a) + (lane >> 2), c = 8 b + 2 (lane & 3) + h; warps own
                // disjoint columns i, so the flush needs no cross-warp sum
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < NCG; ++b) {
                        const int col = 8 * (2 * warp + a) + (lane >> 2), c0 = 8 * b + 2 * (lane & 3);
                        if (c0 < P) sm.flush[col][c0] = acc[a][b][0];
                        if (c0 + 1 < P) sm.flush[col][c0 + 1] = acc[a][b][1];
                        acc[a][b][0] = acc[a][b][1] = 0.0;
                    }
                consumer_bar();
                double* dst = prm.mv_part + ((size_t)g * prm.slots + (cur_ib - ib0)) * (TW * P);
                const double* src = &sm.flush[0][0];
                for (int e = threadIdx.x; e < TW * P; e += NCW * 32) dst[e] = src[e];
                consumer_bar();
                cur_ib = ib;
            }
            if (t == t1) break;
            mbar_wait(&sm.full[pipe.stage], pipe.phase);
            const double* St = &sm.S[pipe.stage][0][0];
            const double* Vt = &sm.V[pipe.stage][0][0];
#pragma unroll
            for (int jg = 0; jg < TJ / 4; ++jg) {
                const double a0 = St[(jg * 16 + 2 * warp) * 32 + lane];
                const double a1 = St[(jg * 16 + 2 * warp + 1) * 32 + lane];
#pragma unroll
                for (int b = 0; b < NCG; ++b) {
                    const int cc = 8 * b + (lane >> 2);
                    const double bf = (cc < P) ? Vt[(4 * jg + (lane & 3)) * P + cc] : 0.0;
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc[0][b][0]), "+d"(acc[0][b][1]) : "d"(a0), "d"(bf));
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(acc[1][b][0]), "+d"(acc[1][b][1]) : "d"(a1), "d"(bf));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[pipe.stage]);
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

// The same sum through a per-CTA table of contributors (no index arithmetic or schedule loads per element)
template <int P>
__device__ __forceinline__ void build_gather_tab(const Params& prm, Smem<P>& sm, int row_lo, int row_hi) {
    if (threadIdx.x == 0) {
        const int G = gridDim.x;
        sm.gt_ok = 1;
        sm.gt_ib_first = row_lo / TW;
        const int ib_last = (row_hi > row_lo) ? (row_hi - 1) / TW : sm.gt_ib_first;
        if (ib_last - sm.gt_ib_first >= 4) sm.gt_ok = 0;
        for (int ib = sm.gt_ib_first; sm.gt_ok && ib <= ib_last; ++ib) {
            const long long tb = (long long)ib * prm.NJT, te = tb + prm.NJT;
            int g = (int)(((double)tb * G) / (double)prm.total_tiles);
            g = min(max(g, 0), G - 1);
            while (g > 0 && prm.tbeg[g] > tb) --g;
            while (g < G - 1 && prm.tbeg[g + 1] <= tb) ++g;
            int cnt = 0;
            for (; g < G; ++g) {
                const long long b = prm.tbeg[g], e = prm.tbeg[g + 1];
                if (b >= te) break;
                if (e <= b) continue;
                if (cnt == 8) {
                    sm.gt_ok = 0;
                    break;
                }
                sm.gt_base[ib - sm.gt_ib_first][cnt++] =
                    ((unsigned long long)g * prm.slots + (unsigned long long)(ib - prm.tib0[g])) * TW * P;
            }
            sm.gt_n[ib - sm.gt_ib_first] = cnt;
        }
    }
    __syncthreads();
}
template <int P>
__device__ __forceinline__ double gather_fast(const Params& prm, const Smem<P>& sm, int row, int c) {
    if (!sm.gt_ok) return gather_elem(prm, row, c, P);
    const int ib = row / TW, ii = row - ib * TW, k = ib - sm.gt_ib_first;
    const int cnt = sm.gt_n[k];
    double out = prm.mv_part[sm.gt_base[k][0] + (size_t)ii * P + c];
    for (int q = 1; q < cnt; ++q) out = out + prm.mv_part[sm.gt_base[k][q] + (size_t)ii * P + c];
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

// after a grid barrier: dst[k] = sum (k < first_min) or minimum over the CTAs of value k of their partial records
// (`stride` doubles per CTA and buffer).  A warp takes 32 consecutive values (coalesced loads of one CTA's record) over a
// contiguous range of CTAs, sixteen loads in flight, added in CTA order; the ranges' results go through shared memory
// (`scratch`, at least NT doubles) and are added in range order.  Same order in every CTA: the totals are bit-identical
// across the grid.
__device__ __forceinline__ void gather_values(const double* __restrict__ part, int stride, double* scratch, double* dst,
                                              int nval, int first_min) {
    constexpr int NW = NT / 32;
    const int G = gridDim.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (nval + 31) >> 5;
    const int gsplit = (nkb >= NW) ? 1 : NW / nkb;
    for (int item = warp; item < nkb * gsplit; item += NW) {
        const int kb = item % nkb, gs = item / nkb;
        const int k = kb * 32 + lane;
        const int g_lo = (int)((long long)G * gs / gsplit), g_hi = (int)((long long)G * (gs + 1) / gsplit);
        const bool is_min = k >= first_min;
        const double neutral = is_min ? CUDART_INF : 0.0;
        double s = neutral;
        if (k < nval) {
            const double* src = part + k;
            for (int g0 = g_lo; g0 < g_hi; g0 += 16) {
                double v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = (g0 + u < g_hi) ? src[(size_t)(g0 + u) * stride] : neutral;
#pragma unroll
                for (int u = 0; u < 16; ++u) s = is_min ? fmin(s, v[u]) : (s + v[u]);
            }
        }
        scratch[gs * (nkb * 32) + k] = s;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nval; k += NT) {
        const bool is_min = k >= first_min;
        double s = scratch[k];
        for (int gs = 1; gs < gsplit; ++gs) {
            const double t = scratch[gs * (nkb * 32) + k];
            s = is_min ? fmin(s, t) : (s + t);
        }
        dst[k] = s;
    }
    __syncthreads();
}

// scal[q*P+c] = sum over CTAs of the Q per-column partial sums
template <int P, int Q>
__device__ __forceinline__ void gather_scalars(const Params& prm, Smem<P>& sm, int buf) {
    gather_values(prm.dot_part + (size_t)buf * gridDim.x * (MAXQ * MAXP), MAXQ * MAXP, &sm.redv[0][0], sm.scal, Q * P, Q * P);
}

// per-column tCG state, identical in every CTA
struct ColState {
    double e_Pe, e_Pd, d_Pd, z_r, r_r, norm_r0, nr_theta, target, model_value, alpha, beta, kappa, xSx, a, b, d, mu, Delta2, q, xd;
    int done, iters, stop;
};

// ------------------------------------------------------------------------------------------------------
// the cooperative kernel.  MODE 1: out = Hw[vin] at (X, Y, mu).  MODE 2: one tCG solve per column at
// (X, Y, mu, Delta) -> out = eta, info.  MODE 3: bare S.V passes (diagnostic).
// ------------------------------------------------------------------------------------------------------
template <int P, int MODE>
__global__ void __launch_bounds__(NT, 1) columns_kernel(Params prm) {
    // whole-solve mode: the host enqueues a few trust-region iterations ahead of the "all done" flag it polls; the launches
    // behind the last iteration find the flag set and return (uniform over the grid: written by the previous launch)
    if (prm.solve && *reinterpret_cast<const volatile int*>(prm.all_done) != 0) return;
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
    build_gather_tab<P>(prm, sm, row_lo, row_hi);
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

    // per-column barrier parameter / radius / activity: hook arguments, or the solver state in whole-solve mode
    if (tid < P) {
        cs[tid].mu = prm.mu;
        cs[tid].Delta2 = prm.Delta * prm.Delta;
        cs[tid].done = (tid >= prm.p) ? 1 : 0;  // padding columns (P > p) never run
        if (prm.solve) {
            const double* st = prm.colstate + tid * CS_FIELDS;
            const int it = (int)st[CS_IT];
            cs[tid].mu = prm.mu_sched[it > 0 ? it - 1 : 0];
            cs[tid].Delta2 = st[CS_DELTA] * st[CS_DELTA];
            if (st[CS_FINISHED] != 0.0) cs[tid].done = 1;
        }
    }
    __syncthreads();
    // ---- point cache (eval_point + begin_step of fam_sphere.cuh), operand V := X -----------------------
    if (!prm.solve) {
        FOR_ELEMS(e) prm.V[e] = prm.X[e];
        fence_proxy_async();
        grid.sync();
        stream_pass<P>(prm, sm, pipe);
        grid.sync();
    }
    {
        double part[3] = {0.0, 0.0, 0.0};
        const double mu_c = cs[myc].mu;
        FOR_ELEMS(e) {
            const double sx = prm.solve ? prm.Sx[e] : gather_fast<P>(prm, sm, (int)(e / P), myc);
            const double x = prm.X[e], y = prm.Y[e];
            const double s = x + prm.eps;
            const double w = mu_c * (1.0 / s);
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
        cs[tid].iters = 0;
        cs[tid].stop = RIPTRM_TCG_MAX_INNER_ITER;
    }
    __syncthreads();
    {
        // c = grad f - G_x(mu/s);  r = c, delta = -c, eta = Heta = 0;  r_r = <c, c>
        double part[2] = {0.0, 0.0};
        const double xSx = cs[myc].xSx, xw = cs[myc].a, mu_c = cs[myc].mu;
        FOR_ELEMS(e) {
            const double x = prm.X[e];
            const double s = x + prm.eps;
            const double w = mu_c * (1.0 / s);
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
                part[1] = fma(x * prm.ys[e], x, part[1]);  // q = <x*(y/s), x>
            } else {
                prm.V[e] = prm.vin[e];
            }
        }
        if (MODE == 2) block_reduce_store<P, 2>(prm, sm, part, buf);
    }
    fence_proxy_async();
    grid.sync();
    int maxinner = prm.tcg_maxinner < 0 ? (n - 1) : prm.tcg_maxinner;
    if (MODE == 2) {
        gather_scalars<P, 2>(prm, sm, buf);
        buf ^= 1;
        if (tid < P) {
            ColState& s = cs[tid];
            s.r_r = sm.scal[tid];
            s.q = sm.scal[P + tid];
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

    // One Hessian-vector product in the reference's operation order (hook riptrm_hessvec)
    for (int j = 0; MODE == 1 && j < maxinner; ++j) {
        // ---- T1: Sv = S V ------------------------------------------------------------------------------
        stream_pass<P>(prm, sm, pipe);
        grid.sync();
        // ---- T2: a = <x, Sv>, b = <x, v> ---------------------------------------------------------------
        {
            double part[2] = {0.0, 0.0};
            FOR_ELEMS(e) {
                const double sv = gather_fast<P>(prm, sm, (int)(e / P), myc);
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
            if (d_Hd <= 0.0 || e_Pe_new >= s.Delta2) {
                // tau solve (RIPTRM.py:123-125); alpha carries tau into the update below
                s.alpha = (-s.e_Pd + sqrt(s.e_Pd * s.e_Pd + s.d_Pd * (s.Delta2 - s.e_Pe))) / s.d_Pd;
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

    // Lock-step tCG with merged reductions (the arithmetic of SphereFam::tcg, fam_sphere.cuh): per iteration one S.V
    // pass, one 6-value and one 4-value reduction round, three vector phases, four grid barriers.
    for (int j = 0; MODE == 2 && j < maxinner; ++j) {
#ifdef RIPTRM_COLUMNS_TIMING  // build with -DRIPTRM_COLUMNS_TIMING: CTA 0 prints the phase times of iteration 5
        const bool dbg = g == 0 && tid == 0 && j == 5;
        uint64_t tdbg[8];
#define TDBG(i) if (dbg) tdbg[i] = global_timer_ns()
#else
#define TDBG(i)
#endif
        TDBG(0);
        stream_pass<P>(prm, sm, pipe);
        TDBG(1);
        grid.sync();
        TDBG(2);
        // ---- M1: Sv; a=<x,Sv> b=<x,v> g1=<w,v> h1=<v,Sv> h2=<v,v> h3=<v,(y/s)v> ------------------------------------
        {
            double part[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            if (cs[myc].done != 1) {
                FOR_ELEMS(e) {
                    const double sv = gather_fast<P>(prm, sm, (int)(e / P), myc);
                    const double x = prm.X[e], v = prm.V[e], ys = prm.ys[e];
                    prm.Sv[e] = sv;
                    part[0] = fma(x, sv, part[0]);
                    part[1] = fma(x, v, part[1]);
                    part[2] = fma(x * ys, v, part[2]);
                    part[3] = fma(v, sv, part[3]);
                    part[4] = fma(v, v, part[4]);
                    part[5] = fma(v, ys * v, part[5]);
                }
            }
            block_reduce_store<P, 6>(prm, sm, part, buf);
        }
        TDBG(3);
        grid.sync();
        gather_scalars<P, 6>(prm, sm, buf);
        TDBG(4);
        buf ^= 1;
        if (tid < P && !cs[tid].done) {
            ColState& s = cs[tid];
            const double a = sm.scal[tid], b = sm.scal[P + tid], g1 = sm.scal[2 * P + tid], h1 = sm.scal[3 * P + tid],
                         h2 = sm.scal[4 * P + tid], h3 = sm.scal[5 * P + tid];
            s.a = a;
            s.b = b;
            s.d = prm.embedded ? g1 : (g1 - b * s.q);
            const double dt = prm.embedded ? h3 : (h3 - b * g1);
            const double d_Hd = (((-h1 + a * b) + s.kappa * h2) + dt) - s.d * b;
            s.alpha = 0.0;
            double e_Pe_new = s.e_Pe;
            if (d_Hd != 0.0) {
                s.alpha = s.z_r / d_Hd;
                e_Pe_new = (s.e_Pe + (2.0 * s.alpha) * s.e_Pd) + (s.alpha * s.alpha) * s.d_Pd;
            }
            s.iters = j + 1;
            if (d_Hd <= 0.0 || e_Pe_new >= s.Delta2) {
                s.alpha = (-s.e_Pd + sqrt(s.e_Pd * s.e_Pd + s.d_Pd * (s.Delta2 - s.e_Pe))) / s.d_Pd;  // tau
                s.stop = (d_Hd <= 0.0) ? RIPTRM_TCG_NEGATIVE_CURVATURE : RIPTRM_TCG_EXCEEDED_TR;
                s.done = 2;
            } else {
                s.e_Pe = e_Pe_new;
            }
        }
        __syncthreads();
        // ---- M2: Hd = Hw[v]; tentative eta, Heta, r; <eta',c> <eta',Heta'> <r',r'> <x,r'> --------------------------
        {
            double part[4] = {0.0, 0.0, 0.0, 0.0};
            const int done = cs[myc].done;
            if (done != 1) {
                const double a = cs[myc].a, b = cs[myc].b, d = cs[myc].d, kappa = cs[myc].kappa, al = cs[myc].alpha;
                FOR_ELEMS(e) {
                    const double x = prm.X[e], v = prm.V[e];
                    const double ga = prm.embedded ? v : (v - x * b);
                    const double tt = prm.ys[e] * ga;
                    const double hl = (-prm.Sv[e] + a * x) + kappa * v;
                    const double gg = tt - d * x;
                    const double hd = hl + gg;
                    const double ne = prm.eta[e] + al * v;
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
                        part[3] = fma(x, nr, part[3]);
                    }
                }
            }
            block_reduce_store<P, 4>(prm, sm, part, buf);
        }
        TDBG(5);
        grid.sync();
        gather_scalars<P, 4>(prm, sm, buf);
        TDBG(6);
        buf ^= 1;
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
                    s.done = 3;
                    if (j >= prm.tcg_mininner && norm_r <= s.target) {
                        s.stop = (prm.tcg_kappa < s.nr_theta) ? RIPTRM_TCG_REACHED_TARGET_LINEAR
                                                               : RIPTRM_TCG_REACHED_TARGET_SUPERLINEAR;
                        s.done = 4;
                    } else {
                        const double zold = s.z_r;
                        s.z_r = s.r_r;
                        s.beta = s.z_r / zold;
                        s.xd = -sm.scal[3 * P + tid] + s.beta * s.b;  // <x, -r' + beta v>
                    }
                }
            }
        }
        __syncthreads();
        // ---- M3: commit; v = P_x(-r + beta v) ----------------------------------------------------------------------
        {
            const int done = cs[myc].done;
            if (done >= 3) {
                const double beta = cs[myc].beta, xd = cs[myc].xd;
                FOR_ELEMS(e) {
                    const double rr = prm.r2[e];
                    prm.eta[e] = prm.eta2[e];
                    prm.Heta[e] = prm.Heta2[e];
                    prm.r[e] = rr;
                    if (done == 3) {
                        const double dn = -rr + beta * prm.V[e];
                        prm.V[e] = dn - xd * prm.X[e];
                    }
                }
            }
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
#ifdef RIPTRM_COLUMNS_TIMING
        if (dbg) {
            tdbg[7] = global_timer_ns();
            printf("columns tCG it 5 (CTA 0, ns): stream %llu | sync %llu | M1 %llu | sync+gather %llu | M2 %llu | sync+gather %llu | M3+sync %llu | total %llu\n",
                   (unsigned long long)(tdbg[1] - tdbg[0]), (unsigned long long)(tdbg[2] - tdbg[1]),
                   (unsigned long long)(tdbg[3] - tdbg[2]), (unsigned long long)(tdbg[4] - tdbg[3]),
                   (unsigned long long)(tdbg[5] - tdbg[4]), (unsigned long long)(tdbg[6] - tdbg[5]),
                   (unsigned long long)(tdbg[7] - tdbg[6]), (unsigned long long)(tdbg[7] - tdbg[0]));
        }
#endif
#undef TDBG
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
        if (prm.solve && g == 0 && tid < prm.p && prm.colstate[tid * CS_FIELDS + CS_FINISHED] == 0.0) {
            double* st = prm.colstate + tid * CS_FIELDS;
            st[CS_TCG_ITERS] = (double)cs[tid].iters;
            st[CS_TCG_STOP] = (double)cs[tid].stop;
            st[CS_KAPPA] = cs[tid].kappa;
            st[CS_CNT_TCG] += (double)cs[tid].iters;
        }
        if (g == 0 && tid < P && tid < prm.p && prm.info != nullptr) {
            prm.info[tid * 4 + 0] = (double)cs[tid].iters;
            prm.info[tid * 4 + 1] = (double)cs[tid].stop;
            prm.info[tid * 4 + 2] = sqrt(sm.scal[tid]);
            prm.info[tid * 4 + 3] = cs[tid].model_value;
        }
    }
#undef FOR_ELEMS
}

// ------------------------------------------------------------------------------------------------------
// Whole solve on this family: the host (riptrm_api.cu) alternates the lock-step tCG launch above with this
// kernel, which performs everything else of one trust-region iteration for every column (RIPTRM.py:735-783,
// :574-705) and the outer-iteration bookkeeping of columns whose inner loop ended (:785-896, utils.py:237-368),
// with the same vector-phase machinery (fixed-order reductions, grid barriers, redundant bit-identical scalars).
// Two S.V passes per call: S x_new for the new point's cache and S dx for the reference's extra Hw(dx) (:659).
// INIT = true: the start of the solve instead (point cache of x0, log row 0, outer iteration 1 begins).
// ------------------------------------------------------------------------------------------------------
template <int P, int QS, int QM>
__device__ __forceinline__ void reduce_store_mixed(const Params& prm, Smem<P>& sm, const double (&part)[QS + QM], int buf) {
    constexpr int NTV = (NT / P) * P;
#pragma unroll
    for (int q = 0; q < QS + QM; ++q) sm.redv[q][threadIdx.x] = part[q];
    __syncthreads();
    if (threadIdx.x < (QS + QM) * P) {
        const int q = threadIdx.x / P, c = threadIdx.x - q * P;
        double s = sm.redv[q][c];
        if (q < QS)
            for (int i = c + P; i < NTV; i += P) s = s + sm.redv[q][i];
        else
            for (int i = c + P; i < NTV; i += P) s = fmin(s, sm.redv[q][i]);
        prm.dot_part[((size_t)buf * gridDim.x + blockIdx.x) * (MAXQ * MAXP) + threadIdx.x] = s;
    }
}
template <int P, int QS, int QM>
__device__ __forceinline__ void gather_mixed(const Params& prm, Smem<P>& sm, int buf) {
    gather_values(prm.dot_part + (size_t)buf * gridDim.x * (MAXQ * MAXP), MAXQ * MAXP, &sm.redv[0][0], sm.scal,
                  (QS + QM) * P, QS * P);
}

struct PostState {  // per column, identical in every CTA
    double it, k, Delta, xSx, cost, cnt_inner, cnt_tcg, cnt_aux, rows, finished, stop, Delta_init, xSx_init, cost_init;
    double mu, tolL, tolC, kappa, normdx, nrm, bdot, xSxN, costN, minx, miny, compl_v, xy, ngl, pl_cur, pl_new, a, b, d;
    double ared_pred, radius_update, inner_status, dual_clipping, tcg_iters, tcg_stop, DeltaNext;
    double radius0;  // trust-region radius before the step (the log's TR_radius)
    double t_inner;  // start of the current inner loop, seconds since the start of the solve
    int path;      // 0 idle (finished), 1 converged, 2 primal infeasible, 3 normal (rho test)
    int accept, boundary, rollback;
    int evalc;     // evaluate (utils.py:342-368) this column in this call: boundary, or every step with trace_mode 1
};

template <int P, bool INIT>
__global__ void __launch_bounds__(NT, 1) columns_post_kernel(Params prm) {
    if (!INIT && *reinterpret_cast<const volatile int*>(prm.all_done) != 0) return;   // enqueued ahead of the flag (see columns_kernel)
    const double now_s = (prm.now_ptr != nullptr) ? *prm.now_ptr : prm.now_s;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<P>& sm = *reinterpret_cast<Smem<P>*>(smem_raw);
    __shared__ PostState ps[P];
    cg::grid_group grid = cg::this_grid();
    constexpr int NTV = (NT / P) * P;
    const int g = blockIdx.x, tid = threadIdx.x, n = prm.n;
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
    if (tid < P) {
        PostState& s = ps[tid];
        const double* st = prm.colstate + tid * CS_FIELDS;
        s.it = st[CS_IT]; s.k = st[CS_K]; s.Delta = st[CS_DELTA]; s.xSx = st[CS_XSX]; s.cost = st[CS_COST];
        s.cnt_inner = st[CS_CNT_INNER]; s.cnt_tcg = st[CS_CNT_TCG]; s.cnt_aux = st[CS_CNT_AUX]; s.rows = st[CS_ROWS];
        s.finished = (tid >= prm.p) ? 1.0 : st[CS_FINISHED]; s.stop = st[CS_STOP];
        s.Delta_init = st[CS_DELTA_INIT]; s.xSx_init = st[CS_XSX_INIT]; s.cost_init = st[CS_COST_INIT];
        s.tcg_iters = st[CS_TCG_ITERS]; s.tcg_stop = st[CS_TCG_STOP]; s.kappa = st[CS_KAPPA];
        s.t_inner = st[CS_T_INNER];
        const int it = (int)s.it;
        s.mu = prm.mu_sched[it > 0 ? it - 1 : 0];
        s.tolL = prm.tolL_sched[it > 0 ? it - 1 : 0];
        s.tolC = prm.tolC_sched[it > 0 ? it - 1 : 0];
        s.path = (s.finished != 0.0) ? 0 : 3;
        s.radius0 = s.Delta;
        s.accept = s.boundary = s.rollback = 0;
        s.evalc = 0;
        s.ared_pred = s.radius_update = s.dual_clipping = CUDART_NAN;
        s.inner_status = CUDART_NAN;
        s.normdx = s.minx = s.miny = s.compl_v = CUDART_NAN;
    }
    __syncthreads();
    build_gather_tab<P>(prm, sm, row_lo, row_hi);
    Pipe pipe{0, 0u};
    int buf = 0;

    if (INIT) {
        // x0 -> X, V; S x0; cost, x'Sx; the first outer iteration starts after the evaluation below
        FOR_ELEMS(e) prm.V[e] = prm.X[e];
        fence_proxy_async();
        grid.sync();
        stream_pass<P>(prm, sm, pipe);
        grid.sync();
        {
            double part[1] = {0.0};
            FOR_ELEMS(e) {
                const double sx = gather_fast<P>(prm, sm, (int)(e / P), myc);
                prm.Sx[e] = sx;
                prm.Xprev[e] = prm.X[e];
                part[0] = fma(prm.X[e], sx, part[0]);
            }
            block_reduce_store<P, 1>(prm, sm, part, buf);
        }
        grid.sync();
        gather_scalars<P, 1>(prm, sm, buf);
        buf ^= 1;
        if (tid < P) {
            PostState& s = ps[tid];
            s.xSx = sm.scal[tid];
            s.cost = -0.5 * s.xSx;
            s.it = 0.0;
            s.k = 0.0;
            s.Delta = prm.initial_tr_radius > 0.0 ? prm.initial_tr_radius : 3.141592653589793 / 8.0;
            s.cnt_inner = s.cnt_tcg = s.cnt_aux = s.rows = 0.0;
            s.finished = (tid >= prm.p) ? 1.0 : 0.0;
            s.stop = (double)RIPTRM_STOP_RUNNING;
            s.mu = prm.mu_sched[0];
            s.boundary = (s.finished == 0.0);
            s.path = 0;
        }
        __syncthreads();
    } else {
        // ---- P1: ||dx||^2, <x,dx>, ||x+dx||^2  (RIPTRM.py:735, :743, :744) ---------------------------------
        {
            double part[3] = {0.0, 0.0, 0.0};
            if (ps[myc].path) {
                FOR_ELEMS(e) {
                    const double x = prm.X[e], dx = prm.eta[e], a = x + dx;
                    part[0] = fma(dx, dx, part[0]);
                    part[1] = fma(x, dx, part[1]);
                    part[2] = fma(a, a, part[2]);
                }
            }
            block_reduce_store<P, 3>(prm, sm, part, buf);
        }
        grid.sync();
        gather_scalars<P, 3>(prm, sm, buf);
        buf ^= 1;
        if (tid < P && ps[tid].path) {
            ps[tid].normdx = sqrt(sm.scal[tid]);
            ps[tid].bdot = sm.scal[P + tid];
            ps[tid].nrm = sqrt(sm.scal[2 * P + tid]);
        }
        __syncthreads();
        // ---- P2: yNew, xNew; operand V := xNew ------------------------------------------------------------
        if (ps[myc].path) {
            const double b = ps[myc].bdot, nrm = ps[myc].nrm, mu = ps[myc].mu;
            FOR_ELEMS(e) {
                const double x = prm.X[e], y = prm.Y[e], dx = prm.eta[e];
                const double s = x + prm.eps;
                const double ga = prm.embedded ? dx : (dx - x * b);
                const double dy = (-y + mu * (1.0 / s)) - (y * ga) / s;
                const double xn = (x + dx) / nrm;
                prm.YN[e] = y + dy;
                prm.XN[e] = xn;
                prm.V[e] = xn;
            }
        }
        fence_proxy_async();
        grid.sync();
        stream_pass<P>(prm, sm, pipe);  // S x_new
        grid.sync();
        // ---- P3: Sx_new, x_new' S x_new ---------------------------------------------------------------------
        {
            double part[1] = {0.0};
            if (ps[myc].path) {
                FOR_ELEMS(e) {
                    const double sx = gather_fast<P>(prm, sm, (int)(e / P), myc);
                    prm.SxN[e] = sx;
                    part[0] = fma(prm.XN[e], sx, part[0]);
                }
            }
            block_reduce_store<P, 1>(prm, sm, part, buf);
        }
        grid.sync();
        gather_scalars<P, 1>(prm, sm, buf);
        buf ^= 1;
        if (tid < P && ps[tid].path) {
            ps[tid].xSxN = sm.scal[tid];
            ps[tid].costN = -0.5 * sm.scal[tid];
        }
        __syncthreads();
        // ---- P4: feasibility, complementarity, <x_new, y_new>  (:591-596) ----------------------------------------
        {
            double part[4] = {0.0, 0.0, CUDART_INF, CUDART_INF};
            if (ps[myc].path) {
                const double mu = ps[myc].mu;
                FOR_ELEMS(e) {
                    const double xn = prm.XN[e], yn = prm.YN[e];
                    const double sn = xn + prm.eps;
                    const double cv = yn * sn - mu;
                    part[0] = part[0] + cv * cv;
                    part[1] = fma(xn, yn, part[1]);
                    part[2] = fmin(part[2], sn);
                    part[3] = fmin(part[3], yn);
                }
            }
            reduce_store_mixed<P, 2, 2>(prm, sm, part, buf);
        }
        grid.sync();
        gather_mixed<P, 2, 2>(prm, sm, buf);
        buf ^= 1;
        if (tid < P && ps[tid].path) {
            ps[tid].compl_v = sqrt(sm.scal[tid]);
            ps[tid].xy = sm.scal[P + tid];
            ps[tid].minx = sm.scal[2 * P + tid];
            ps[tid].miny = sm.scal[3 * P + tid];
        }
        __syncthreads();
        // ---- P5: || grad L(x_new, y_new) ||  (:593) ----------------------------------------------------------------
        {
            double part[1] = {0.0};
            if (ps[myc].path) {
                const double xSxN = ps[myc].xSxN, xy = ps[myc].xy;
                FOR_ELEMS(e) {
                    const double xn = prm.XN[e];
                    const double gl = (-prm.SxN[e] + xSxN * xn) - (prm.YN[e] - xy * xn);
                    part[0] = fma(gl, gl, part[0]);
                }
            }
            block_reduce_store<P, 1>(prm, sm, part, buf);
        }
        grid.sync();
        gather_scalars<P, 1>(prm, sm, buf);
        buf ^= 1;
        if (tid < P && ps[tid].path) {
            PostState& s = ps[tid];
            s.ngl = sqrt(sm.scal[tid]);
            const bool xfe = s.minx > 0.0, yfe = s.miny > 0.0;
            if (xfe && yfe && s.ngl <= s.tolL && s.compl_v <= s.tolC) s.path = 1;        // :762-766
            else if (!xfe) s.path = 2;                                                     // :769-775
            else s.path = 3;
        }
        __syncthreads();
        // ---- P6: log-barrier sums; operand V := dx for the extra Hessian-vector product (:644-659) -----------------------
        bool any_normal = false;
#pragma unroll
        for (int c = 0; c < P; ++c) any_normal = any_normal || (ps[c].path == 3);
        if (any_normal) {
            {
                double part[2] = {0.0, 0.0};
                if (ps[myc].path == 3) {
                    FOR_ELEMS(e) {
                        part[0] = part[0] + det_log(prm.X[e] + prm.eps);
                        part[1] = part[1] + det_log(prm.XN[e] + prm.eps);
                        prm.V[e] = prm.eta[e];
                    }
                }
                block_reduce_store<P, 2>(prm, sm, part, buf);
            }
            fence_proxy_async();
            grid.sync();
            gather_scalars<P, 2>(prm, sm, buf);
            buf ^= 1;
            if (tid < P && ps[tid].path == 3) {
                ps[tid].pl_cur = sm.scal[tid];
                ps[tid].pl_new = sm.scal[P + tid];
            }
            __syncthreads();
            if (prm.reuse_heta) {
                // <Hw dx, dx>, <c, dx> with the Hw[eta] the tCG kernel accumulated beside eta
                double part[2] = {0.0, 0.0};
                if (ps[myc].path == 3) {
                    FOR_ELEMS(e) {
                        const double v = prm.eta[e];
                        part[0] = fma(prm.Heta[e], v, part[0]);
                        part[1] = fma(prm.c[e], v, part[1]);
                    }
                }
                block_reduce_store<P, 2>(prm, sm, part, buf);
                grid.sync();
                gather_scalars<P, 2>(prm, sm, buf);
                buf ^= 1;
            } else {
            stream_pass<P>(prm, sm, pipe);  // S dx
            grid.sync();
            // ---- P7-P9: Hw[dx] as in the tCG kernel, then <Hw dx, dx>, <c, dx> ------------------------------------------
            {
                double part[2] = {0.0, 0.0};
                if (ps[myc].path == 3) {
                    FOR_ELEMS(e) {
                        const double sv = gather_fast<P>(prm, sm, (int)(e / P), myc);
                        const double x = prm.X[e];
                        prm.Sv[e] = sv;
                        part[0] = fma(x, sv, part[0]);
                        part[1] = fma(x, prm.eta[e], part[1]);
                    }
                }
                block_reduce_store<P, 2>(prm, sm, part, buf);
            }
            grid.sync();
            gather_scalars<P, 2>(prm, sm, buf);
            buf ^= 1;
            if (tid < P && ps[tid].path == 3) {
                ps[tid].a = sm.scal[tid];
                ps[tid].b = sm.scal[P + tid];
            }
            __syncthreads();
            {
                double part[1] = {0.0};
                if (ps[myc].path == 3) {
                    const double b = ps[myc].b;
                    FOR_ELEMS(e) {
                        const double x = prm.X[e], v = prm.eta[e];
                        const double ga = prm.embedded ? v : (v - x * b);
                        const double tt = prm.ys[e] * ga;
                        prm.t[e] = tt;
                        part[0] = fma(x, tt, part[0]);
                    }
                }
                block_reduce_store<P, 1>(prm, sm, part, buf);
            }
            grid.sync();
            gather_scalars<P, 1>(prm, sm, buf);
            buf ^= 1;
            if (tid < P && ps[tid].path == 3) ps[tid].d = sm.scal[tid];
            __syncthreads();
            {
                double part[2] = {0.0, 0.0};
                if (ps[myc].path == 3) {
                    const double a = ps[myc].a, d = ps[myc].d, kappa = ps[myc].kappa;
                    FOR_ELEMS(e) {
                        const double x = prm.X[e], v = prm.eta[e];
                        const double hl = (-prm.Sv[e] + a * x) + kappa * v;
                        const double gg = prm.t[e] - d * x;
                        const double hd = hl + gg;
                        part[0] = fma(hd, v, part[0]);
                        part[1] = fma(prm.c[e], v, part[1]);
                    }
                }
                block_reduce_store<P, 2>(prm, sm, part, buf);
            }
            grid.sync();
            gather_scalars<P, 2>(prm, sm, buf);
            buf ^= 1;
            }
        }
        // ---- P10: rho test, radius update (:660-677), acceptance, inner-loop bookkeeping (:808-842) --------------------
        if (tid < P && ps[tid].path) {
            PostState& s = ps[tid];
            s.k += 1.0;
            s.cnt_inner += 1.0;
            s.DeltaNext = s.Delta;
            if (s.path == 1) {
                s.inner_status = (double)RIPTRM_INNER_CONVERGED;
                s.accept = 1;
                s.boundary = 1;
            } else if (s.path == 2) {
                s.inner_status = (double)RIPTRM_INNER_PRIMAL_INFEASIBLE;
                s.DeltaNext = prm.gamma * s.normdx;
            } else {
                if (!prm.reuse_heta) s.cnt_aux += 1.0;
                const double phi_cur = s.cost - s.mu * s.pl_cur, phi_new = s.costN - s.mu * s.pl_new;
                double ared = phi_cur - phi_new;
                double pred = (0.0 - 0.5 * sm.scal[tid]) - sm.scal[P + tid];
                const double reg = (fmax(1.0, fabs(phi_cur)) * 2.220446049250313e-16) * prm.reduction_regularization;
                ared = ared + reg;
                pred = pred + reg;
                s.ared_pred = ared / pred;
                if (ared < 0.25 * pred) {
                    s.radius_update = (double)RIPTRM_RADIUS_REDUCED;
                    s.DeltaNext = 0.25 * s.Delta;
                } else if (ared >= 0.75 * pred && fabs(s.normdx - s.Delta) <= 1e-15) {
                    s.radius_update = (double)RIPTRM_RADIUS_EXPANDED;
                    s.DeltaNext = fmin(2.0 * s.Delta, prm.maximal_tr_radius);
                } else {
                    s.radius_update = (double)RIPTRM_RADIUS_UNCHANGED;
                }
                if (ared > prm.rho * pred) {
                    s.inner_status = (double)RIPTRM_INNER_SUCCESSFUL;
                    s.accept = 2;  // accept with dual clipping
                } else {
                    s.inner_status = (double)RIPTRM_INNER_UNSUCCESSFUL;
                }
            }
            {   // :822-834 (after the step, as the reference)
                const double rt = (prm.inner_maxtime < 0.0) ? now_s : (now_s - s.t_inner);
                const double lim = (prm.inner_maxtime < 0.0) ? prm.maxtime : prm.inner_maxtime;
                if (rt >= lim) {
                    s.inner_status = (double)RIPTRM_INNER_MAX_TIME;
                    s.rollback = 1;
                    s.boundary = 1;
                }
            }
            if (prm.inner_maxiter >= 0 && (int)s.k >= prm.inner_maxiter) {  // :835-842
                s.inner_status = (double)RIPTRM_INNER_MAX_ITER;
                s.rollback = 1;
                s.boundary = 1;
            }
        }
        __syncthreads();
        // ---- P10b: commit (x, y) / clip the duals (:681-696) / roll back -------------------------------------------------
        {
            double part[1] = {0.0};
            const PostState& s = ps[myc];
            if (s.path) {
                const double I_right = fmax(prm.const_right, prm.const_right / s.mu);
                FOR_ELEMS(e) {
                    if (s.rollback) {
                        prm.X[e] = prm.Xinit[e];
                        prm.Y[e] = prm.Yinit[e];
                        prm.Sx[e] = prm.Sxinit[e];
                    } else if (s.accept == 1) {
                        prm.X[e] = prm.XN[e];
                        prm.Y[e] = prm.YN[e];
                        prm.Sx[e] = prm.SxN[e];
                    } else if (s.accept == 2) {
                        const double yn = prm.YN[e], xn = prm.XN[e];
                        const double I_left = prm.const_left * fmin(fmin(prm.Y[e], s.mu / (xn + prm.eps)), 1.0);
                        const double cl = fmin(fmax(yn, I_left), I_right);
                        if (!(cl == yn)) part[0] = part[0] + 1.0;
                        prm.X[e] = xn;
                        prm.Y[e] = cl;
                        prm.Sx[e] = prm.SxN[e];
                    }
                }
            }
            block_reduce_store<P, 1>(prm, sm, part, buf);
        }
        grid.sync();
        gather_scalars<P, 1>(prm, sm, buf);
        buf ^= 1;
        if (tid < P && ps[tid].path) {
            PostState& s = ps[tid];
            if (s.rollback) {
                s.xSx = s.xSx_init;
                s.cost = s.cost_init;
                s.Delta = s.Delta_init;
            } else {
                if (s.accept) {
                    s.xSx = s.xSxN;
                    s.cost = s.costN;
                }
                if (s.accept == 2) s.dual_clipping = (sm.scal[tid] > 0.0) ? 1.0 : 0.0;
                if (s.path != 1) s.Delta = s.DeltaNext;   // converged: the radius is returned unchanged (:762-766)
            }
        }
        __syncthreads();
    }

    // ---- evaluation (utils.py:342-368) of the columns at an outer-iteration boundary (inner loop just ended, or INIT) and,
    //      with trace_mode 1, of every running column (one log row per trust-region iteration, RIPTRM.py:812-818);
    //      then the log row, the stop tests (base_solver.py:85-106) and the barrier update (:890-894) -----------------------
    if (tid < P) ps[tid].evalc = ps[tid].boundary || (prm.trace_mode == 1 && !INIT && ps[tid].path != 0);
    __syncthreads();
    {
        double part[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, CUDART_INF, CUDART_INF};
        if (ps[myc].evalc) {
            FOR_ELEMS(e) {
                const double x = prm.X[e], y = prm.Y[e];
                const double gi = -(x + prm.eps);
                const double cv = y * gi, nv = fmax(-y, 0.0), iv = fmax(gi, 0.0);
                part[0] = fma(x, y, part[0]);
                part[1] = fma(x, x, part[1]);
                part[2] = fma(prm.Xprev[e], x, part[2]);
                part[3] = part[3] + cv * cv;
                part[4] = part[4] + nv * nv;
                part[5] = part[5] + iv * iv;
                part[6] = part[6] + iv;
                part[7] = fmin(part[7], -iv);
                part[8] = fmin(part[8], -fabs(y));
            }
        }
        reduce_store_mixed<P, 7, 2>(prm, sm, part, buf);
    }
    grid.sync();
    gather_mixed<P, 7, 2>(prm, sm, buf);
    buf ^= 1;
    __shared__ double ev[P][9];
    if (tid < P && ps[tid].evalc) {
        for (int q = 0; q < 9; ++q) ev[tid][q] = sm.scal[q * P + tid];
    }
    __syncthreads();
    {
        double part[1] = {0.0};
        if (ps[myc].evalc) {
            const double xSx = ps[myc].xSx, xy = ev[myc][0];
            FOR_ELEMS(e) {
                const double x = prm.X[e];
                const double gl = (-prm.Sx[e] + xSx * x) - (prm.Y[e] - xy * x);
                part[0] = fma(gl, gl, part[0]);
            }
        }
        block_reduce_store<P, 1>(prm, sm, part, buf);
    }
    grid.sync();
    gather_scalars<P, 1>(prm, sm, buf);
    buf ^= 1;
    if (tid < P && ps[tid].evalc) {
        PostState& s = ps[tid];
        const double gradnorm = sqrt(sm.scal[tid]);
        const double p_compl = ev[tid][3], p_nonneg = ev[tid][4], p_ineq = ev[tid][5];
        const double man_v = sqrt(ev[tid][1]) - 1.0;
        const double residual = sqrt(((((gradnorm * gradnorm + p_compl) + p_nonneg) + p_ineq) + 0.0) + man_v * man_v);
        const double max_v = -ev[tid][7], mean_v = ev[tid][6] / (double)n;
        const double distance = acos(fmax(fmin(ev[tid][2], 1.0), -1.0));
        const int it = (int)s.it;
        const double mu_next = prm.mu_sched[it];  // the barrier parameter after the update (:890-893), mu_sched[0] at row 0
        // trace_mode 1: row 0 and one row per trust-region iteration; trace_mode 2: one row per outer iteration
        const bool inner_row = prm.trace_mode == 1 && !INIT;
        const bool write_row = (prm.trace_mode == 1) || (prm.trace_mode == 2 && s.boundary);
        if (write_row) {
            if (prm.trace != nullptr && g == 0 && (int)s.rows < prm.trace_capacity) {
                double* row = prm.trace + ((size_t)tid * prm.trace_capacity + (size_t)s.rows) * RIPTRM_TRACE_FIELDS;
                for (int f = 0; f < RIPTRM_TRACE_FIELDS; ++f) row[f] = CUDART_NAN;
                row[RIPTRM_TR_ITERATION] = (double)it;
                row[RIPTRM_TR_MU] = inner_row ? s.mu : mu_next;
                if (!INIT) {
                    row[RIPTRM_TR_NUM_INNER] = s.k;
                    row[RIPTRM_TR_INNER_STATUS] = s.inner_status;
                    row[RIPTRM_TR_RADIUS] = inner_row ? s.radius0 : s.Delta;
                }
                if (inner_row) {
                    row[RIPTRM_TR_NUM_INNER] = s.rollback ? (double)prm.inner_maxiter : s.k;
                    row[RIPTRM_TR_DXTYPE] = s.tcg_stop;
                    row[RIPTRM_TR_TCG_ITERS] = s.tcg_iters;
                    row[RIPTRM_TR_NORMDX] = s.normdx;
                    row[RIPTRM_TR_MINXFEASI] = s.minx;
                    row[RIPTRM_TR_MINYFEASI] = s.miny;
                    row[RIPTRM_TR_COMPL] = s.compl_v;
                    row[RIPTRM_TR_ARED_PRED] = s.ared_pred;
                    row[RIPTRM_TR_RADIUS_UPDATE] = s.radius_update;
                    row[RIPTRM_TR_DUAL_CLIPPING] = s.dual_clipping;
                }
                row[RIPTRM_TR_MAXABSLAGMULT] = -ev[tid][8];
                row[RIPTRM_TR_COST] = s.cost;
                row[RIPTRM_TR_DISTANCE] = distance;
                row[RIPTRM_TR_RESIDUAL] = residual;
                row[RIPTRM_TR_GRADNORM] = gradnorm;
                row[RIPTRM_TR_COMPLVIOLATION] = sqrt(p_compl);
                row[RIPTRM_TR_DUALVIOLATION] = sqrt(p_nonneg);
                row[RIPTRM_TR_MANVIOLATION] = man_v;
                row[RIPTRM_TR_MAXVIOLATION] = max_v;
                row[RIPTRM_TR_MEANVIOLATION] = mean_v;
                row[RIPTRM_TR_TIME] = now_s;
            }
            s.rows += 1.0;
        }
        if (s.boundary) {
            int stop = RIPTRM_STOP_RUNNING;
            if (now_s >= prm.maxtime) stop = RIPTRM_STOP_MAXTIME;
            else if (it >= prm.maxiter) stop = RIPTRM_STOP_MAXITER;
            if (residual <= prm.tolresid) stop = RIPTRM_STOP_TOLRESID;
            if (g == 0 && prm.summary != nullptr) {
                double* sm_ = prm.summary + (size_t)tid * RIPTRM_SUMMARY_FIELDS;
                sm_[RIPTRM_SM_COST] = s.cost;
                sm_[RIPTRM_SM_RESIDUAL] = residual;
                sm_[RIPTRM_SM_GRADNORM] = gradnorm;
                sm_[RIPTRM_SM_COMPLVIOLATION] = sqrt(p_compl);
                sm_[RIPTRM_SM_DUALVIOLATION] = sqrt(p_nonneg);
                sm_[RIPTRM_SM_MANVIOLATION] = man_v;
                sm_[RIPTRM_SM_MAXVIOLATION] = max_v;
                sm_[RIPTRM_SM_MEANVIOLATION] = mean_v;
                sm_[RIPTRM_SM_MU] = mu_next;
                sm_[RIPTRM_SM_RADIUS] = fmax(s.Delta, it > 0 ? prm.minimal_initial_tr_radius : s.Delta);
                sm_[RIPTRM_SM_OUTER_ITERS] = (double)it;
                sm_[RIPTRM_SM_INNER_ITERS] = s.cnt_inner;
                sm_[RIPTRM_SM_TCG_ITERS] = s.cnt_tcg;
                sm_[RIPTRM_SM_AUX_HESSVECS] = s.cnt_aux;
                sm_[RIPTRM_SM_STOP_REASON] = (double)stop;
                sm_[RIPTRM_SM_TRACE_ROWS] = s.rows;
            }
            if (it > 0) s.Delta = fmax(s.Delta, prm.minimal_initial_tr_radius);   // :894
            if (stop != RIPTRM_STOP_RUNNING) {
                s.finished = 1.0;
                s.stop = (double)stop;
            } else {
                s.it = (double)(it + 1);
                s.k = 0.0;
                s.Delta_init = s.Delta;
                s.xSx_init = s.xSx;
                s.cost_init = s.cost;
                s.t_inner = now_s;
            }
        }
    }
    __syncthreads();
    // start-of-inner-run copies for the rollback (:794-796) and the previous iterate for `distance`
    if (ps[myc].boundary && ps[myc].finished == 0.0) {
        FOR_ELEMS(e) {
            const double x = prm.X[e];
            prm.Xinit[e] = x;
            prm.Yinit[e] = prm.Y[e];
            prm.Sxinit[e] = prm.Sx[e];
            prm.Xprev[e] = x;
        }
    } else if (prm.trace_mode == 1 && ps[myc].evalc) {
        FOR_ELEMS(e) prm.Xprev[e] = prm.X[e];  // previous inner iterate for the next row's `distance` (RIPTRM.py:819)
    }
    if (g == 0 && tid < P) {
        const PostState& s = ps[tid];
        double* st = prm.colstate + tid * CS_FIELDS;
        st[CS_IT] = s.it; st[CS_K] = s.k; st[CS_DELTA] = s.Delta; st[CS_XSX] = s.xSx; st[CS_COST] = s.cost;
        st[CS_CNT_INNER] = s.cnt_inner; st[CS_CNT_TCG] = s.cnt_tcg; st[CS_CNT_AUX] = s.cnt_aux; st[CS_ROWS] = s.rows;
        st[CS_FINISHED] = s.finished; st[CS_STOP] = s.stop;
        st[CS_DELTA_INIT] = s.Delta_init; st[CS_XSX_INIT] = s.xSx_init; st[CS_COST_INIT] = s.cost_init;
        st[CS_T_INNER] = s.t_inner;
    }
    if (g == 0 && tid == 0) {
        bool all = true;
        for (int c = 0; c < prm.p; ++c) all = all && (ps[c].finished != 0.0);
        *prm.all_done = all ? 1 : 0;
        if (prm.cond_on) cudaGraphSetConditional((cudaGraphConditionalHandle)prm.cond, all ? 0u : 1u);
        if (prm.now_ptr != nullptr) {   // the clock stamp the next iteration's launches read (device-side loop)
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            double* nw = const_cast<double*>(prm.now_ptr);
            nw[0] = (double)(t - *reinterpret_cast<const unsigned long long*>(nw + 1)) * 1e-9;
        }
    }
#undef FOR_ELEMS
}

// S = Z + Z' into the streaming layout: tile (ib, jt) = rows [jt*TJ, +TJ) x columns [ib*TW, +TW) stored
// contiguously ([TJ][TW] row-major) at tile index ib * NJT + jt; padding stays zero.
// frag = 1 (P = 16, DMMA consumers): inside a tile the 4 x 8 blocks (rows 4jg.., columns 8ig..) are stored as the 32
// consecutive doubles of an mma.m8n8k4 A-fragment: entry (jj, ii) of block (jg, ig) at ((jg*16 + ig)*32 + ii*4 + jj).
__global__ void build_S_kernel(const double* __restrict__ Z, double* __restrict__ S, int n, int NJT, int frag) {
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
            const int r = i % TJ, cidx = j % TW;
            const size_t within = frag ? (size_t)(((r >> 2) * 16 + (cidx >> 3)) * 32 + (cidx & 7) * 4 + (r & 3))
                                       : (size_t)r * TW + cidx;
            S[tile_id * (TJ * TW) + within] = Z[(size_t)i * n + j] + tile[tx][k];
        }
    }
}

}  // namespace col
}  // namespace riptrm

// fam_stiefel.cuh -- NonnegPCA on the Stiefel manifold St(n, p) with one large data matrix
// (RIPTRM_FAMILY_NONNEGPCA_STIEFEL; BASELINE config 4 "n = 20000, p = 10 on Stiefel").
//
//   minimise f(X) = -tr(X'ZX)  over X'X = I_p  subject to  g_ij(X) = -X_ij - eps <= 0   (m = n p constraints).
//
// ONE RIPTRM run (RIPTRM.py:909-976) whose tangent vectors are n x p matrices.  With S = Z + Z', s = X + eps,
// P_X U = U - X sym(X'U) (pymanopt Stiefel: projection = to_tangent_space = egrad2rgrad; ehess2rhess =
// P_X(ehess - V sym(X' egrad)); retraction = qf(X + V) with positive diagonal R; metric = Frobenius):
//   c      = grad f - G_X(mu / s)                     = P_X(-S X - mu / s)                      (RIPTRM.py:730)
//   Hw[V]  = Hess L[V] + G_X(y o G*_X[V] / s)         = P_X(-S V + V C1 + (Y / s) o G*_X[V])    (RIPTRM.py:729)
//            C1 = sym(X'SX) + sym(X'Y),  G*_X[V] = P_X V  (V itself under is_euclidean_embedded)
//   dy     = -y + mu / s - y o G*_X[dx] / s                                                     (RIPTRM.py:743)
// The expensive operator is the same dense contraction S.V as in fam_columns.cuh, and this file reuses that file's
// persistent cooperative grid: stream-K tile schedule, TMA producer warp + 5-stage mbarrier ring, DMMA (P >= 8) or DFMA
// consumers, per-CTA partials added in CTA order.  What differs is everything between two S.V passes: the tCG state
// is one set of scalars (inner products are Frobenius sums over all columns), and every projection needs a p x p
// matrix X'U, reduced over the grid exactly like the dot products (thread (row, c) accumulates column c of X'U).
// Per tCG iteration: one S.V pass, three reduction rounds, four grid barriers.
// All sums have a fixed order for a given grid size: results are deterministic run to run.
#pragma once
#include "fam_columns.cuh"

namespace riptrm {
namespace stf {

using namespace col;

constexpr int KMAXQ = 3 * MAXP + 6;         // reduced quantities per round and column, at most
constexpr int DOT_STRIDE = KMAXQ * MAXP;    // doubles per CTA and buffer in dot_part
constexpr int SS_C1 = CS_FIELDS;            // colstate[SS_C1 ...]: C1 (P x P) of the current point, for the post kernel

// per-thread sums -> per-CTA sums per column c, written to dot_part[buf][cta][(qoff + q) * P + c].
// The last KMIN quantities are minima.  Chunks of MAXQ quantities go through sm.redv.
template <int P, int K, int KMIN>
__device__ __forceinline__ void breduce(const Params& prm, Smem<P>& sm, const double (&part)[K], int buf, int qoff = 0) {
    constexpr int NTV = (NT / P) * P;
#pragma unroll
    for (int q0 = 0; q0 < K; q0 += MAXQ) {
        const int nq = (K - q0 < MAXQ) ? (K - q0) : MAXQ;
#pragma unroll
        for (int q = 0; q < MAXQ; ++q)
            if (q0 + q < K) sm.redv[q][threadIdx.x] = part[q0 + q];
        __syncthreads();
        if (threadIdx.x < nq * P) {
            const int q = threadIdx.x / P, c = threadIdx.x - q * P;
            const bool is_min = (q0 + q) >= K - KMIN;
            double s = sm.redv[q][c];
            if (is_min)
                for (int i = c + P; i < NTV; i += P) s = fmin(s, sm.redv[q][i]);
            else
                for (int i = c + P; i < NTV; i += P) s = s + sm.redv[q][i];
            prm.dot_part[((size_t)buf * gridDim.x + blockIdx.x) * DOT_STRIDE + (size_t)(qoff + q0 + q) * P + c] = s;
        }
        __syncthreads();
    }
}

// after a grid barrier: tot[q * P + c] = sum (q < first_min) or minimum over the CTAs (col::gather_values)
template <int P>
__device__ __forceinline__ void gather_tot(const Params& prm, Smem<P>& sm, double* tot, int buf, int K, int first_min) {
    gather_values(prm.dot_part + (size_t)buf * gridDim.x * DOT_STRIDE, DOT_STRIDE, &sm.redv[0][0], tot, K * P, first_min * P);
}

// sum over the columns of one reduced quantity (column order)
template <int P>
__device__ __forceinline__ double colsum(const double* tot, int q) {
    double s = tot[q * P];
#pragma unroll
    for (int c = 1; c < P; ++c) s = s + tot[q * P + c];
    return s;
}

// acc[a] += xrow[a] * val: column c of X'U accumulated by the thread that owns U[row, c]
template <int P>
__device__ __forceinline__ void xt_acc(double* acc, const double* __restrict__ xrow, double val) {
#pragma unroll
    for (int a = 0; a < P; ++a) acc[a] = fma(xrow[a], val, acc[a]);
}
// sum_a row[a] * M[a][c] with column c of M in registers
template <int P>
__device__ __forceinline__ double row_dot(const double* __restrict__ row, const double (&mc)[P]) {
    double s = row[0] * mc[0];
#pragma unroll
    for (int a = 1; a < P; ++a) s = fma(row[a], mc[a], s);
    return s;
}
template <int P>
__device__ __forceinline__ void load_col(double (&mc)[P], const double* M, int c) {
#pragma unroll
    for (int a = 0; a < P; ++a) mc[a] = M[a * P + c];
}
// sum_{a, c} A[a][c] * B[a][c]
template <int P>
__device__ __forceinline__ double mat_dot(const double* A, const double* B) {
    double s = 0.0;
    for (int k = 0; k < P * P; ++k) s = fma(A[k], B[k], s);
    return s;
}

struct TcgState {
    double e_Pe, e_Pd, d_Pd, z_r, r_r, norm_r0, nr_theta, target, model_value, alpha, beta, mu, Delta2;
    int done, iters, stop;
};

// ------------------------------------------------------------------------------------------------------
// MODE 1: out = Hw[vin] at (X, Y, mu).  MODE 2: one tCG solve at (X, Y, mu, Delta) -> out = eta, info.
// ------------------------------------------------------------------------------------------------------
template <int P, int MODE>
__global__ void __launch_bounds__(NT, 1) stiefel_kernel(Params prm) {
    if (prm.solve && *reinterpret_cast<const volatile int*>(prm.all_done) != 0) return;   // enqueued ahead of the flag (columns_kernel)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<P>& sm = *reinterpret_cast<Smem<P>*>(smem_raw);
    __shared__ TcgState ts;
    __shared__ double C1[P * P];
    cg::grid_group grid = cg::this_grid();
    constexpr int NTV = (NT / P) * P;
    constexpr int RSTEP = NTV / P;
    const int g = blockIdx.x, tid = threadIdx.x, n = prm.n;
    const int row_lo = min(n, g * prm.R), row_hi = min(n, row_lo + prm.R);
    const int myc = tid % P;
    const bool vthread = tid < NTV && myc < prm.p;
    const int row0 = row_lo + tid / P;
#define FOR_ROWS(row, e) \
    for (int row = row0; vthread && row < row_hi; row += RSTEP) \
        for (size_t e = (size_t)row * P + myc, once_ = 1; once_; once_ = 0)
    // p x p scratch in the flush buffer of the S.V pass: valid between two passes only
    double* tot = &sm.flush[0][0];
    double* symA = tot + (3 * P + 6) * P;
    double* symB = symA + P * P;
    double* symC = symB + P * P;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ts.mu = prm.mu;
        ts.Delta2 = prm.Delta * prm.Delta;
        ts.done = 0;
        if (prm.solve) {
            const int it = (int)prm.colstate[CS_IT];
            ts.mu = prm.mu_sched[it > 0 ? it - 1 : 0];
            ts.Delta2 = prm.colstate[CS_DELTA] * prm.colstate[CS_DELTA];
            if (prm.colstate[CS_FINISHED] != 0.0) ts.done = 1;
        }
    }
    __syncthreads();
    build_gather_tab<P>(prm, sm, row_lo, row_hi);
    Pipe pipe{0, 0u};
    int buf = 0;
    if (ts.done) return;  // uniform over the grid

    // ---- point cache: S X (kept by the solver between calls), Y / s, the three p x p matrices of the point -------
    if (!prm.solve) {
        FOR_ROWS(row, e) prm.V[e] = prm.X[e];
        fence_proxy_async();
        grid.sync();
        stream_pass<P>(prm, sm, pipe);
        grid.sync();
    }
    const double mu = ts.mu;
    {
        double part[3 * P];
#pragma unroll
        for (int q = 0; q < 3 * P; ++q) part[q] = 0.0;
        FOR_ROWS(row, e) {
            const double sx = prm.solve ? prm.Sx[e] : gather_fast<P>(prm, sm, row, myc);
            const double x = prm.X[e], y = prm.Y[e];
            const double s = x + prm.eps;
            const double w = mu * (1.0 / s);
            prm.Sx[e] = sx;
            prm.ys[e] = y / s;
            const double* xr = prm.X + (size_t)row * P;
#pragma unroll
            for (int a = 0; a < P; ++a) {
                const double xa = xr[a];
                part[a] = fma(xa, sx, part[a]);          // X'SX
                part[P + a] = fma(xa, y, part[P + a]);   // X'Y
                part[2 * P + a] = fma(xa, w, part[2 * P + a]);  // X'(mu/s)
            }
        }
        breduce<P, 3 * P, 0>(prm, sm, part, buf);
    }
    grid.sync();
    gather_tot<P>(prm, sm, tot, buf, 3 * P, 3 * P);
    buf ^= 1;
    if (tid < P * P) {
        const int a = tid / P, c = tid - a * P;
        const double sxx = tot[a * P + c], sxxT = tot[c * P + a];
        const double xy = tot[(P + a) * P + c], xyT = tot[(P + c) * P + a];
        const double xw = tot[(2 * P + a) * P + c], xwT = tot[(2 * P + c) * P + a];
        const double c1 = 0.5 * (sxx + sxxT) + 0.5 * (xy + xyT);
        C1[tid] = c1;
        symA[tid] = 0.5 * ((-sxx - xw) + (-sxxT - xwT));  // sym(X'(-SX - mu/s))
        if (prm.solve && g == 0) prm.colstate[SS_C1 + tid] = c1;
    }
    __syncthreads();
    {
        // c = P_X(-SX - mu/s);  MODE 2: r = c, delta = -c, eta = Heta = 0;  <c, c>;  X'delta
        double part[P + 1];
#pragma unroll
        for (int q = 0; q < P + 1; ++q) part[q] = 0.0;
        double mc[P];
        load_col<P>(mc, symA, myc);
        FOR_ROWS(row, e) {
            const double x = prm.X[e];
            const double w = mu * (1.0 / (x + prm.eps));
            const double* xr = prm.X + (size_t)row * P;
            const double cc = (-prm.Sx[e] - w) - row_dot<P>(xr, mc);
            prm.c[e] = cc;
            double v;
            if (MODE == 2) {
                prm.r[e] = cc;
                prm.eta[e] = 0.0;
                prm.Heta[e] = 0.0;
                v = -cc;
                part[P] = fma(cc, cc, part[P]);
            } else {
                v = prm.vin[e];
            }
            prm.V[e] = v;
            xt_acc<P>(part, xr, v);
        }
        breduce<P, P + 1, 0>(prm, sm, part, buf);
    }
    fence_proxy_async();
    grid.sync();
    const int dim = n * prm.p - (prm.p * (prm.p + 1)) / 2;
    int maxinner = prm.tcg_maxinner < 0 ? dim : prm.tcg_maxinner;
    if (MODE == 1) maxinner = 1;
    if (MODE == 2 && tid == 0) {
        ts.iters = 0;
        ts.stop = RIPTRM_TCG_MAX_INNER_ITER;
    }

    for (int j = 0; j < maxinner; ++j) {
#ifdef RIPTRM_COLUMNS_TIMING  // build with -DRIPTRM_COLUMNS_TIMING: CTA 0 prints the phase times of iteration 5
        const bool dbg = g == 0 && tid == 0 && j == 5;
        uint64_t tdbg[12];
#define TDBG(i) if (dbg) tdbg[i] = global_timer_ns()
#else
#define TDBG(i)
#endif
        TDBG(0);
        stream_pass<P>(prm, sm, pipe);
        TDBG(1);
        grid.sync();
        TDBG(2);
        // X'delta of the direction just streamed (partials written before the barrier that preceded the pass)
        gather_tot<P>(prm, sm, tot, buf, P + 1, P + 1);
        TDBG(3);
        buf ^= 1;
        if (tid < P * P) {
            const int a = tid / P, c = tid - a * P;
            symB[tid] = 0.5 * (tot[a * P + c] + tot[c * P + a]);
        }
        if (MODE == 2 && j == 0 && tid == 0) {
            TcgState& s = ts;
            s.r_r = colsum<P>(tot, P);
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
        // ---- M1: W = -S V + V C1 + (Y/s) o G*[V];  X'W, <V, W> ------------------------------------------------
        {
            double part[P + 1];
#pragma unroll
            for (int q = 0; q < P + 1; ++q) part[q] = 0.0;
            double c1c[P], bc[P];
            load_col<P>(c1c, C1, myc);
            load_col<P>(bc, symB, myc);
            FOR_ROWS(row, e) {
                const double sv = gather_fast<P>(prm, sm, row, myc);
                const double* xr = prm.X + (size_t)row * P;
                const double* vr = prm.V + (size_t)row * P;
                const double v = prm.V[e];
                const double ga = prm.embedded ? v : (v - row_dot<P>(xr, bc));
                const double w = (-sv + row_dot<P>(vr, c1c)) + prm.ys[e] * ga;
                prm.t[e] = w;
                xt_acc<P>(part, xr, w);
                part[P] = fma(v, w, part[P]);
            }
            TDBG(4);
            breduce<P, P + 1, 0>(prm, sm, part, buf);
        }
        TDBG(5);
        grid.sync();
        TDBG(6);
        gather_tot<P>(prm, sm, tot, buf, P + 1, P + 1);
        TDBG(7);
        buf ^= 1;
        if (tid < P * P) {
            const int a = tid / P, c = tid - a * P;
            symA[tid] = 0.5 * (tot[a * P + c] + tot[c * P + a]);  // M = sym(X'W):  Hw[V] = W - X M
        }
        __syncthreads();
        if (MODE == 1) {
            double mc[P];
            load_col<P>(mc, symA, myc);
            FOR_ROWS(row, e) prm.out[e] = prm.t[e] - row_dot<P>(prm.X + (size_t)row * P, mc);
            break;
        }
        if (tid == 0) {
            TcgState& s = ts;
            // <V, Hw V> = <V, W> - <X'V, M>
            const double d_Hd = colsum<P>(tot, P) - mat_dot<P>(symB, symA);
            s.alpha = 0.0;
            double e_Pe_new = s.e_Pe;
            if (d_Hd != 0.0) {
                s.alpha = s.z_r / d_Hd;
                e_Pe_new = (s.e_Pe + (2.0 * s.alpha) * s.e_Pd) + (s.alpha * s.alpha) * s.d_Pd;
            }
            s.iters = j + 1;
            if (d_Hd <= 0.0 || e_Pe_new >= s.Delta2) {
                s.alpha = (-s.e_Pd + sqrt(s.e_Pd * s.e_Pd + s.d_Pd * (s.Delta2 - s.e_Pe))) / s.d_Pd;  // tau (:123-125)
                s.stop = (d_Hd <= 0.0) ? RIPTRM_TCG_NEGATIVE_CURVATURE : RIPTRM_TCG_EXCEEDED_TR;
                s.done = 2;
            } else {
                s.e_Pe = e_Pe_new;
            }
        }
        __syncthreads();
        // ---- M2: Hd = W - X M; tentative eta, Heta, r;  X'r', <eta', c>, <eta', Heta'>, <r', r'> ------------------
        {
            double part[P + 3];
#pragma unroll
            for (int q = 0; q < P + 3; ++q) part[q] = 0.0;
            const int done = ts.done;
            const double al = ts.alpha;
            double mc[P];
            load_col<P>(mc, symA, myc);
            FOR_ROWS(row, e) {
                const double* xr = prm.X + (size_t)row * P;
                const double hd = prm.t[e] - row_dot<P>(xr, mc);
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
                    xt_acc<P>(part, xr, nr);
                    part[P] = fma(ne, prm.c[e], part[P]);
                    part[P + 1] = fma(ne, nh, part[P + 1]);
                    part[P + 2] = fma(nr, nr, part[P + 2]);
                }
            }
            breduce<P, P + 3, 0>(prm, sm, part, buf);
        }
        TDBG(8);
        grid.sync();
        gather_tot<P>(prm, sm, tot, buf, P + 3, P + 3);
        TDBG(9);
        buf ^= 1;
        if (tid == 0) {
            TcgState& s = ts;
            if (s.done == 2) {
                s.done = 1;
            } else {
                const double new_model = colsum<P>(tot, P) + 0.5 * colsum<P>(tot, P + 1);
                if (new_model >= s.model_value) {
                    s.stop = RIPTRM_TCG_MODEL_INCREASED;   // :162-165, the previous eta is returned
                    s.done = 1;
                } else {
                    s.model_value = new_model;
                    s.r_r = colsum<P>(tot, P + 2);
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
                    }
                }
            }
        }
        __syncthreads();
        if (tid < P * P && ts.done == 3) {
            const int a = tid / P, c = tid - a * P;
            // sym(X'(-r' + beta V)) for the re-projection of the new direction (:210)
            symC[tid] = -0.5 * (tot[a * P + c] + tot[c * P + a]) + ts.beta * symB[tid];
        }
        __syncthreads();
        // ---- M3: commit;  V = P_X(-r' + beta V);  X'V ------------------------------------------------------------
        {
            double part[P + 1];
#pragma unroll
            for (int q = 0; q < P + 1; ++q) part[q] = 0.0;
            const int done = ts.done;
            if (done >= 3) {
                const double beta = ts.beta;
                double mc[P];
                load_col<P>(mc, symC, myc);
                FOR_ROWS(row, e) {
                    const double rr = prm.r2[e];
                    prm.eta[e] = prm.eta2[e];
                    prm.Heta[e] = prm.Heta2[e];
                    prm.r[e] = rr;
                    if (done == 3) {
                        const double* xr = prm.X + (size_t)row * P;
                        const double dn = (-rr + beta * prm.V[e]) - row_dot<P>(xr, mc);
                        prm.V[e] = dn;
                        xt_acc<P>(part, xr, dn);
                    }
                }
            }
            breduce<P, P + 1, 0>(prm, sm, part, buf);
        }
        if (tid == 0) {
            TcgState& s = ts;
            if (s.done == 3) {
                s.e_Pd = s.beta * (s.e_Pd + s.alpha * s.d_Pd);
                s.d_Pd = s.z_r + (s.beta * s.beta) * s.d_Pd;
                s.done = 0;
            } else if (s.done == 4) {
                s.done = 1;
            }
        }
        __syncthreads();
        const bool finished = ts.done == 1;
        fence_proxy_async();
        TDBG(10);
        grid.sync();
#ifdef RIPTRM_COLUMNS_TIMING
        if (dbg) {
            tdbg[11] = global_timer_ns();
            printf("stiefel tCG it 5 (CTA 0, ns): stream %llu | sync %llu | gather B %llu | M1 body %llu | M1 reduce %llu | sync %llu | "
                   "gather %llu | M2 %llu | sync+gather %llu | M3 %llu | sync %llu | total %llu\n",
                   (unsigned long long)(tdbg[1] - tdbg[0]), (unsigned long long)(tdbg[2] - tdbg[1]),
                   (unsigned long long)(tdbg[3] - tdbg[2]), (unsigned long long)(tdbg[4] - tdbg[3]),
                   (unsigned long long)(tdbg[5] - tdbg[4]), (unsigned long long)(tdbg[6] - tdbg[5]),
                   (unsigned long long)(tdbg[7] - tdbg[6]), (unsigned long long)(tdbg[8] - tdbg[7]),
                   (unsigned long long)(tdbg[9] - tdbg[8]), (unsigned long long)(tdbg[10] - tdbg[9]),
                   (unsigned long long)(tdbg[11] - tdbg[10]), (unsigned long long)(tdbg[11] - tdbg[0]));
        }
#endif
#undef TDBG
        if (finished) break;
    }

    if (MODE == 2) {
        double part[1] = {0.0};
        FOR_ROWS(row, e) {
            const double v = prm.eta[e];
            prm.out[e] = v;
            part[0] = fma(v, v, part[0]);
        }
        breduce<P, 1, 0>(prm, sm, part, buf);
        grid.sync();
        gather_tot<P>(prm, sm, tot, buf, 1, 1);
        if (g == 0 && tid == 0) {
            if (prm.solve) {
                prm.colstate[CS_TCG_ITERS] = (double)ts.iters;
                prm.colstate[CS_TCG_STOP] = (double)ts.stop;
                prm.colstate[CS_CNT_TCG] += (double)ts.iters;
            }
            if (prm.info != nullptr) {
                prm.info[0] = (double)ts.iters;
                prm.info[1] = (double)ts.stop;
                prm.info[2] = sqrt(colsum<P>(tot, 0));
                prm.info[3] = ts.model_value;
            }
        }
    }
#undef FOR_ROWS
}

// ------------------------------------------------------------------------------------------------------
// Whole solve: everything of one trust-region iteration except the tCG (RIPTRM.py:735-783, :574-705) and the
// outer-iteration bookkeeping when the inner loop ends (:785-896, utils.py:237-368).  Two S.V passes per call:
// S x_new (the new point's cache) and S dx (the reference's extra Hw(dx), :659).  INIT = true: start of the solve.
// ------------------------------------------------------------------------------------------------------
struct PostState {
    double it, k, Delta, cost, cnt_inner, cnt_tcg, cnt_aux, rows, finished, stop, Delta_init, cost_init;
    double mu, tolL, tolC, normdx, costN, minx, miny, compl_v, ngl, pl_cur, pl_new, hdx, cdx;
    double ared_pred, radius_update, inner_status, dual_clipping, tcg_iters, tcg_stop, DeltaNext, radius0, t_inner;
    int path;      // 0 idle (finished), 1 converged, 2 primal infeasible, 3 normal (rho test)
    int accept, boundary, rollback, evalc;
};

template <int P, bool INIT>
__global__ void __launch_bounds__(NT, 1) stiefel_post_kernel(Params prm) {
    if (!INIT && *reinterpret_cast<const volatile int*>(prm.all_done) != 0) return;   // enqueued ahead of the flag (columns_kernel)
    const double now_s = (prm.now_ptr != nullptr) ? *prm.now_ptr : prm.now_s;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<P>& sm = *reinterpret_cast<Smem<P>*>(smem_raw);
    __shared__ PostState ps;
    __shared__ double C1[P * P];
    __shared__ double keepB[P * P];  // sym(X'dx), needed on both sides of an S.V pass
    __shared__ double ev[16];
    cg::grid_group grid = cg::this_grid();
    constexpr int NTV = (NT / P) * P;
    constexpr int RSTEP = NTV / P;
    const int g = blockIdx.x, tid = threadIdx.x, n = prm.n, p = prm.p;
    const int row_lo = min(n, g * prm.R), row_hi = min(n, row_lo + prm.R);
    const int myc = tid % P;
    const bool vthread = tid < NTV && myc < p;
    const int row0 = row_lo + tid / P;
#define FOR_ROWS(row, e) \
    for (int row = row0; vthread && row < row_hi; row += RSTEP) \
        for (size_t e = (size_t)row * P + myc, once_ = 1; once_; once_ = 0)
    double* tot = &sm.flush[0][0];
    double* symA = tot + (3 * P + 6) * P;
    double* symB = symA + P * P;
    double* symC = symB + P * P;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        PostState& s = ps;
        const double* st = prm.colstate;
        s.it = st[CS_IT]; s.k = st[CS_K]; s.Delta = st[CS_DELTA]; s.cost = st[CS_COST];
        s.cnt_inner = st[CS_CNT_INNER]; s.cnt_tcg = st[CS_CNT_TCG]; s.cnt_aux = st[CS_CNT_AUX]; s.rows = st[CS_ROWS];
        s.finished = st[CS_FINISHED]; s.stop = st[CS_STOP];
        s.Delta_init = st[CS_DELTA_INIT]; s.cost_init = st[CS_COST_INIT];
        s.tcg_iters = st[CS_TCG_ITERS]; s.tcg_stop = st[CS_TCG_STOP];
        s.t_inner = st[CS_T_INNER];
        const int it = (int)s.it;
        s.mu = prm.mu_sched[it > 0 ? it - 1 : 0];
        s.tolL = prm.tolL_sched[it > 0 ? it - 1 : 0];
        s.tolC = prm.tolC_sched[it > 0 ? it - 1 : 0];
        s.path = (s.finished != 0.0) ? 0 : 3;
        s.radius0 = s.Delta;
        s.accept = s.boundary = s.rollback = s.evalc = 0;
        s.ared_pred = s.radius_update = s.dual_clipping = s.inner_status = CUDART_NAN;
        s.normdx = s.minx = s.miny = s.compl_v = CUDART_NAN;
    }
    if (!INIT && tid < P * P) C1[tid] = prm.colstate[SS_C1 + tid];
    __syncthreads();
    build_gather_tab<P>(prm, sm, row_lo, row_hi);
    Pipe pipe{0, 0u};
    int buf = 0;

    if (INIT) {
        FOR_ROWS(row, e) prm.V[e] = prm.X[e];
        fence_proxy_async();
        grid.sync();
        stream_pass<P>(prm, sm, pipe);
        grid.sync();
        {
            double part[1] = {0.0};
            FOR_ROWS(row, e) {
                const double sx = gather_fast<P>(prm, sm, row, myc);
                prm.Sx[e] = sx;
                prm.Xprev[e] = prm.X[e];
                part[0] = fma(prm.X[e], sx, part[0]);
            }
            breduce<P, 1, 0>(prm, sm, part, buf);
        }
        grid.sync();
        gather_tot<P>(prm, sm, tot, buf, 1, 1);
        buf ^= 1;
        if (tid == 0) {
            PostState& s = ps;
            s.cost = -0.5 * colsum<P>(tot, 0);
            s.it = 0.0;
            s.k = 0.0;
            s.Delta = prm.initial_tr_radius > 0.0 ? prm.initial_tr_radius : sqrt((double)p) / 8.0;  // typical_dist / 8
            s.cnt_inner = s.cnt_tcg = s.cnt_aux = s.rows = 0.0;
            s.finished = 0.0;
            s.stop = (double)RIPTRM_STOP_RUNNING;
            s.mu = prm.mu_sched[0];
            s.boundary = 1;
            s.path = 0;
        }
        __syncthreads();
    } else if (ps.path) {  // uniform over the grid
        const double mu = ps.mu;
        // ---- P1: <dx, dx>, X'dx, (X + dx)'(X + dx)  (RIPTRM.py:735, :743, :744) --------------------------------
        {
            double part[2 * P + 1];
#pragma unroll
            for (int q = 0; q < 2 * P + 1; ++q) part[q] = 0.0;
            FOR_ROWS(row, e) {
                const double* xr = prm.X + (size_t)row * P;
                const double* dr = prm.eta + (size_t)row * P;
                const double dx = prm.eta[e], ac = prm.X[e] + dx;
#pragma unroll
                for (int a = 0; a < P; ++a) {
                    const double xa = xr[a];
                    part[a] = fma(xa, dx, part[a]);
                    part[P + a] = fma(xa + dr[a], ac, part[P + a]);
                }
                part[2 * P] = fma(dx, dx, part[2 * P]);
            }
            breduce<P, 2 * P + 1, 0>(prm, sm, part, buf);
        }
        grid.sync();
        gather_tot<P>(prm, sm, tot, buf, 2 * P + 1, 2 * P + 1);
        buf ^= 1;
        if (tid < P * P) {
            const int a = tid / P, c = tid - a * P;
            keepB[tid] = 0.5 * (tot[a * P + c] + tot[c * P + a]);
            symA[tid] = 0.5 * (tot[(P + a) * P + c] + tot[(P + c) * P + a]);  // Gram matrix of X + dx
            symC[tid] = 0.0;
        }
        __syncthreads();
        if (tid == 0) {
            // qf(X + dx) = (X + dx) R^-1 with R'R = Gram, diag(R) > 0 -- numpy's QR with the sign fix of
            // pymanopt's Stiefel retraction.  The Gram matrix is I + dx'dx up to rounding: well conditioned.
            ps.normdx = sqrt(colsum<P>(tot, 2 * P));
            double* R = symA;
            for (int j = 0; j < p; ++j) {
                double d = R[j * P + j];
                for (int k = 0; k < j; ++k) d = d - R[k * P + j] * R[k * P + j];
                d = sqrt(d);
                R[j * P + j] = d;
                for (int i = j + 1; i < p; ++i) {
                    double v = R[j * P + i];
                    for (int k = 0; k < j; ++k) v = v - R[k * P + j] * R[k * P + i];
                    R[j * P + i] = v / d;
                }
            }
            double* Ri = symC;  // upper triangular inverse, column by column
            for (int j = 0; j < p; ++j) {
                Ri[j * P + j] = 1.0 / R[j * P + j];
                for (int i = j - 1; i >= 0; --i) {
                    double v = 0.0;
                    for (int k = i + 1; k <= j; ++k) v = fma(R[i * P + k], Ri[k * P + j], v);
                    Ri[i * P + j] = -v / R[i * P + i];
                }
            }
        }
        __syncthreads();
        // ---- P2: yNew, xNew; operand V := xNew -------------------------------------------------------------------
        {
            double bc[P], ric[P];
            load_col<P>(bc, keepB, myc);
            load_col<P>(ric, symC, myc);
            FOR_ROWS(row, e) {
                const double* xr = prm.X + (size_t)row * P;
                const double* dr = prm.eta + (size_t)row * P;
                const double x = prm.X[e], y = prm.Y[e], dx = prm.eta[e];
                const double s = x + prm.eps;
                const double ga = prm.embedded ? dx : (dx - row_dot<P>(xr, bc));
                const double dy = (-y + mu * (1.0 / s)) - (y * ga) / s;
                double xn = 0.0;
#pragma unroll
                for (int a = 0; a < P; ++a)
                    if (a <= myc) xn = fma(xr[a] + dr[a], ric[a], xn);
                prm.YN[e] = y + dy;
                prm.XN[e] = xn;
                prm.V[e] = xn;
            }
        }
        fence_proxy_async();
        grid.sync();
        stream_pass<P>(prm, sm, pipe);  // S x_new
        grid.sync();
        // ---- P3: S x_new;  xNew'S xNew, xNew'yNew, complementarity, feasibility (:591-596) -------------------------
        {
            double part[2 * P + 3];
#pragma unroll
            for (int q = 0; q < 2 * P + 1; ++q) part[q] = 0.0;
            part[2 * P + 1] = part[2 * P + 2] = CUDART_INF;
            FOR_ROWS(row, e) {
                const double sx = gather_fast<P>(prm, sm, row, myc);
                prm.SxN[e] = sx;
                const double* xr = prm.XN + (size_t)row * P;
                const double xn = prm.XN[e], yn = prm.YN[e];
                const double sn = xn + prm.eps;
                const double cv = yn * sn - mu;
#pragma unroll
                for (int a = 0; a < P; ++a) {
                    const double xa = xr[a];
                    part[a] = fma(xa, sx, part[a]);
                    part[P + a] = fma(xa, yn, part[P + a]);
                }
                part[2 * P] = part[2 * P] + cv * cv;
                part[2 * P + 1] = fmin(part[2 * P + 1], sn);
                part[2 * P + 2] = fmin(part[2 * P + 2], yn);
            }
            breduce<P, 2 * P + 3, 2>(prm, sm, part, buf);
        }
        grid.sync();
        gather_tot<P>(prm, sm, tot, buf, 2 * P + 3, 2 * P + 1);
        buf ^= 1;
        if (tid < P * P) {
            const int a = tid / P, c = tid - a * P;
            // sym(xNew'(-S xNew - yNew)) for grad L(xNew, yNew) = P(-S xNew - yNew)
            symA[tid] = 0.5 * ((-tot[a * P + c] - tot[(P + a) * P + c]) + (-tot[c * P + a] - tot[(P + c) * P + a]));
        }
        if (tid == 0) {
            double tr = 0.0;
            for (int a = 0; a < p; ++a) tr = tr + tot[a * P + a];
            ps.costN = -0.5 * tr;
            ps.compl_v = sqrt(colsum<P>(tot, 2 * P));
            double mx = CUDART_INF, my = CUDART_INF;
            for (int c = 0; c < p; ++c) {
                mx = fmin(mx, tot[(2 * P + 1) * P + c]);
                my = fmin(my, tot[(2 * P + 2) * P + c]);
            }
            ps.minx = mx;
            ps.miny = my;
        }
        __syncthreads();
        // ---- P5: || grad L(xNew, yNew) ||  (:593) ----------------------------------------------------------------------
        {
            double part[1] = {0.0};
            double mc[P];
            load_col<P>(mc, symA, myc);
            FOR_ROWS(row, e) {
                const double gl = (-prm.SxN[e] - prm.YN[e]) - row_dot<P>(prm.XN + (size_t)row * P, mc);
                part[0] = fma(gl, gl, part[0]);
            }
            breduce<P, 1, 0>(prm, sm, part, buf);
        }
        grid.sync();
        gather_tot<P>(prm, sm, tot, buf, 1, 1);
        buf ^= 1;
        if (tid == 0) {
            PostState& s = ps;
            s.ngl = sqrt(colsum<P>(tot, 0));
            const bool xfe = s.minx > 0.0, yfe = s.miny > 0.0;
            if (xfe && yfe && s.ngl <= s.tolL && s.compl_v <= s.tolC) s.path = 1;        // :762-766
            else if (!xfe) s.path = 2;                                                     // :769-775
            else s.path = 3;
        }
        __syncthreads();
        if (ps.path == 3) {
            // ---- P6: log-barrier sums; operand V := dx for the extra Hessian-vector product (:644-659) -----------------
            {
                double part[2] = {0.0, 0.0};
                FOR_ROWS(row, e) {
                    part[0] = part[0] + det_log(prm.X[e] + prm.eps);
                    part[1] = part[1] + det_log(prm.XN[e] + prm.eps);
                    prm.V[e] = prm.eta[e];
                }
                breduce<P, 2, 0>(prm, sm, part, buf);
            }
            fence_proxy_async();
            grid.sync();
            gather_tot<P>(prm, sm, tot, buf, 2, 2);
            buf ^= 1;
            if (tid == 0) {
                ps.pl_cur = colsum<P>(tot, 0);
                ps.pl_new = colsum<P>(tot, 1);
            }
            __syncthreads();
            if (prm.reuse_heta) {
                // <Hw dx, dx>, <c, dx> with the Hw[eta] the tCG kernel accumulated beside eta (fam_columns.cuh Params::reuse_heta)
                double part[2] = {0.0, 0.0};
                FOR_ROWS(row, e) {
                    const double v = prm.eta[e];
                    part[0] = fma(prm.Heta[e], v, part[0]);
                    part[1] = fma(prm.c[e], v, part[1]);
                }
                breduce<P, 2, 0>(prm, sm, part, buf);
                grid.sync();
                gather_tot<P>(prm, sm, tot, buf, 2, 2);
                buf ^= 1;
                if (tid == 0) {
                    ps.hdx = colsum<P>(tot, 0);
                    ps.cdx = colsum<P>(tot, 1);
                }
                __syncthreads();
            } else {
            stream_pass<P>(prm, sm, pipe);  // S dx
            grid.sync();
            // ---- P7: W of Hw[dx];  X'W, <dx, W>, <c, dx> ---------------------------------------------------------------
            {
                double part[P + 2];
#pragma unroll
                for (int q = 0; q < P + 2; ++q) part[q] = 0.0;
                double c1c[P], bc[P];
                load_col<P>(c1c, C1, myc);
                load_col<P>(bc, keepB, myc);
                FOR_ROWS(row, e) {
                    const double sv = gather_fast<P>(prm, sm, row, myc);
                    const double* xr = prm.X + (size_t)row * P;
                    const double* dr = prm.eta + (size_t)row * P;
                    const double v = prm.eta[e];
                    const double ga = prm.embedded ? v : (v - row_dot<P>(xr, bc));
                    const double w = (-sv + row_dot<P>(dr, c1c)) + prm.ys[e] * ga;
                    xt_acc<P>(part, xr, w);
                    part[P] = fma(v, w, part[P]);
                    part[P + 1] = fma(prm.c[e], v, part[P + 1]);
                }
                breduce<P, P + 2, 0>(prm, sm, part, buf);
            }
            grid.sync();
            gather_tot<P>(prm, sm, tot, buf, P + 2, P + 2);
            buf ^= 1;
            if (tid < P * P) {
                const int a = tid / P, c = tid - a * P;
                symA[tid] = 0.5 * (tot[a * P + c] + tot[c * P + a]);
            }
            __syncthreads();
            if (tid == 0) {
                ps.hdx = colsum<P>(tot, P) - mat_dot<P>(keepB, symA);  // <Hw dx, dx>
                ps.cdx = colsum<P>(tot, P + 1);
            }
            __syncthreads();
            }
        }
        // ---- P10: rho test, radius update (:660-677), acceptance, inner-loop bookkeeping (:808-842) --------------------
        if (tid == 0) {
            PostState& s = ps;
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
                double pred = (0.0 - 0.5 * s.hdx) - s.cdx;
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
                    s.accept = 2;
                } else {
                    s.inner_status = (double)RIPTRM_INNER_UNSUCCESSFUL;
                }
            }
            {   // :822-834
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
            const int accept = ps.accept, rollback = ps.rollback;
            const double I_right = fmax(prm.const_right, prm.const_right / mu);
            FOR_ROWS(row, e) {
                if (rollback) {
                    prm.X[e] = prm.Xinit[e];
                    prm.Y[e] = prm.Yinit[e];
                    prm.Sx[e] = prm.Sxinit[e];
                } else if (accept == 1) {
                    prm.X[e] = prm.XN[e];
                    prm.Y[e] = prm.YN[e];
                    prm.Sx[e] = prm.SxN[e];
                } else if (accept == 2) {
                    const double yn = prm.YN[e], xn = prm.XN[e];
                    const double I_left = prm.const_left * fmin(fmin(prm.Y[e], mu / (xn + prm.eps)), 1.0);
                    const double cl = fmin(fmax(yn, I_left), I_right);
                    if (!(cl == yn)) part[0] = part[0] + 1.0;
                    prm.X[e] = xn;
                    prm.Y[e] = cl;
                    prm.Sx[e] = prm.SxN[e];
                }
            }
            breduce<P, 1, 0>(prm, sm, part, buf);
        }
        grid.sync();
        gather_tot<P>(prm, sm, tot, buf, 1, 1);
        buf ^= 1;
        if (tid == 0) {
            PostState& s = ps;
            if (s.rollback) {
                s.cost = s.cost_init;
                s.Delta = s.Delta_init;
            } else {
                if (s.accept) s.cost = s.costN;
                if (s.accept == 2) s.dual_clipping = (colsum<P>(tot, 0) > 0.0) ? 1.0 : 0.0;
                if (s.path != 1) s.Delta = s.DeltaNext;   // converged: the radius is returned unchanged (:762-766)
            }
        }
        __syncthreads();
    }

    // ---- evaluation (utils.py:342-368) at an outer-iteration boundary (or INIT) and, with trace_mode 1, after every
    //      trust-region iteration; then the log row, the stop tests (base_solver.py:85-106), the barrier update (:890-894)
    if (tid == 0) ps.evalc = ps.boundary || (prm.trace_mode == 1 && !INIT && ps.path != 0);
    __syncthreads();
    if (ps.evalc) {  // uniform over the grid
        {
            double part[2 * P];
#pragma unroll
            for (int q = 0; q < 2 * P; ++q) part[q] = 0.0;
            FOR_ROWS(row, e) {
                const double* xr = prm.X + (size_t)row * P;
                const double sx = prm.Sx[e], y = prm.Y[e];
#pragma unroll
                for (int a = 0; a < P; ++a) {
                    const double xa = xr[a];
                    part[a] = fma(xa, sx, part[a]);
                    part[P + a] = fma(xa, y, part[P + a]);
                }
            }
            breduce<P, 2 * P, 0>(prm, sm, part, buf, 0);
        }
        {
            double part[P + 6];
#pragma unroll
            for (int q = 0; q < P + 4; ++q) part[q] = 0.0;
            part[P + 4] = part[P + 5] = CUDART_INF;
            FOR_ROWS(row, e) {
                const double* xr = prm.X + (size_t)row * P;
                const double x = prm.X[e], y = prm.Y[e];
                const double gi = -(x + prm.eps);
                const double cv = y * gi, nv = fmax(-y, 0.0), iv = fmax(gi, 0.0);
                xt_acc<P>(part, xr, x);                     // X'X
                part[P] = part[P] + cv * cv;
                part[P + 1] = part[P + 1] + nv * nv;
                part[P + 2] = part[P + 2] + iv * iv;
                part[P + 3] = part[P + 3] + iv;
                part[P + 4] = fmin(part[P + 4], -iv);
                part[P + 5] = fmin(part[P + 5], -fabs(y));
            }
            breduce<P, P + 6, 2>(prm, sm, part, buf, 2 * P);
        }
        grid.sync();
        gather_tot<P>(prm, sm, tot, buf, 3 * P + 6, 3 * P + 4);
        buf ^= 1;
        if (tid < P * P) {
            const int a = tid / P, c = tid - a * P;
            symA[tid] = 0.5 * ((-tot[a * P + c] - tot[(P + a) * P + c]) + (-tot[c * P + a] - tot[(P + c) * P + a]));
        }
        if (tid == 0) {
            double mv = 0.0;  // || X'X - I ||_F
            for (int a = 0; a < p; ++a)
                for (int c = 0; c < p; ++c) {
                    const double d = tot[(2 * P + a) * P + c] - (a == c ? 1.0 : 0.0);
                    mv = fma(d, d, mv);
                }
            ev[0] = sqrt(mv);
            ev[1] = colsum<P>(tot, 3 * P);       // sum (y g)^2
            ev[2] = colsum<P>(tot, 3 * P + 1);   // sum max(-y, 0)^2
            ev[3] = colsum<P>(tot, 3 * P + 2);   // sum max(g, 0)^2
            ev[4] = colsum<P>(tot, 3 * P + 3);   // sum max(g, 0)
            double mx = CUDART_INF, my = CUDART_INF;
            for (int c = 0; c < p; ++c) {
                mx = fmin(mx, tot[(3 * P + 4) * P + c]);
                my = fmin(my, tot[(3 * P + 5) * P + c]);
            }
            ev[5] = -mx;                         // max violation
            ev[6] = -my;                         // max |y|
        }
        __syncthreads();
        {
            double part[1] = {0.0};
            double mc[P];
            load_col<P>(mc, symA, myc);
            FOR_ROWS(row, e) {
                const double gl = (-prm.Sx[e] - prm.Y[e]) - row_dot<P>(prm.X + (size_t)row * P, mc);
                part[0] = fma(gl, gl, part[0]);
            }
            breduce<P, 1, 0>(prm, sm, part, buf);
        }
        grid.sync();
        gather_tot<P>(prm, sm, tot, buf, 1, 1);
        buf ^= 1;
        if (tid == 0) {
            PostState& s = ps;
            const double gradnorm = sqrt(colsum<P>(tot, 0));
            const double p_compl = ev[1], p_nonneg = ev[2], p_ineq = ev[3], man_v = ev[0];
            const double residual = sqrt(((((gradnorm * gradnorm + p_compl) + p_nonneg) + p_ineq) + 0.0) + man_v * man_v);
            const double max_v = ev[5], mean_v = ev[4] / ((double)n * (double)p);
            const int it = (int)s.it;
            const double mu_next = prm.mu_sched[it];
            const bool inner_row = prm.trace_mode == 1 && !INIT;
            const bool write_row = (prm.trace_mode == 1) || (prm.trace_mode == 2 && s.boundary);
            if (write_row) {
                if (prm.trace != nullptr && g == 0 && (int)s.rows < prm.trace_capacity) {
                    double* row = prm.trace + (size_t)s.rows * RIPTRM_TRACE_FIELDS;
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
                    row[RIPTRM_TR_MAXABSLAGMULT] = ev[6];
                    row[RIPTRM_TR_COST] = s.cost;
                    row[RIPTRM_TR_DISTANCE] = CUDART_NAN;  // pymanopt's Stiefel has no dist()
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
                    double* sm_ = prm.summary;
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
                    s.cost_init = s.cost;
                    s.t_inner = now_s;
                }
            }
        }
        __syncthreads();
    }
    // start-of-inner-run copies for the rollback (:794-796)
    if (ps.boundary && ps.finished == 0.0) {
        FOR_ROWS(row, e) {
            prm.Xinit[e] = prm.X[e];
            prm.Yinit[e] = prm.Y[e];
            prm.Sxinit[e] = prm.Sx[e];
        }
    }
    if (g == 0 && tid == 0) {
        const PostState& s = ps;
        double* st = prm.colstate;
        st[CS_IT] = s.it; st[CS_K] = s.k; st[CS_DELTA] = s.Delta; st[CS_COST] = s.cost;
        st[CS_CNT_INNER] = s.cnt_inner; st[CS_CNT_TCG] = s.cnt_tcg; st[CS_CNT_AUX] = s.cnt_aux; st[CS_ROWS] = s.rows;
        st[CS_FINISHED] = s.finished; st[CS_STOP] = s.stop;
        st[CS_DELTA_INIT] = s.Delta_init; st[CS_COST_INIT] = s.cost_init;
        st[CS_T_INNER] = s.t_inner;
        *prm.all_done = (s.finished != 0.0) ? 1 : 0;
        if (prm.cond_on) cudaGraphSetConditional((cudaGraphConditionalHandle)prm.cond, (s.finished != 0.0) ? 0u : 1u);
        if (prm.now_ptr != nullptr) {   // the clock stamp the next iteration's launches read (device-side loop)
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            double* nw = const_cast<double*>(prm.now_ptr);
            nw[0] = (double)(t - *reinterpret_cast<const unsigned long long*>(nw + 1)) * 1e-9;
        }
    }
#undef FOR_ROWS
}

}  // namespace stf
}  // namespace riptrm

// fam_stableid.cuh -- the reference's StableIdentification workload: fit A = (J - R) Q to trajectory data on
// Product[SkewSymmetric(d), SPD(d), SPD(d)] with box / "two-box" constraints on entries of A
// (src/StableIdentification/coordinator.py:34-179; closed forms SURVEY.md App. A.3, manifold formulas App. B).
//
//   f(J,R,Q) = tr(E E') / N,  E = X' - (I + h A) X                                          (:92-98)
//   G_A = df/dA = -2h E X'/N ;  along (dJ,dR,dQ): dA = (dJ-dR) Q + (J-R) dQ,  dG_A = 2h^2 dA (X X')/N
//   chain rule for any phi(A) with Phi = dphi/dA:  d/dJ = Phi Q', d/dR = -Phi Q', d/dQ = (J-R)' Phi
//   constraints (:108-152), Phi_i = coef_i E_{rc}:  g = -A_rc + a (coef -1) | A_rc - a (coef +1) |
//                                                   -(A_rc - a)^2 + b (coef -2 (A_rc - a), d coef = -2 dA_rc)
//   Riemannian conversions: Skew: skew(.) ; SPD (affine-invariant metric): rgrad = P sym(eg) P,
//   rhess = P sym(eh) P + sym(V sym(eg) P), <A,B>_P = tr(P^-1 A P^-1 B), retraction sym(P + V + V P^-1 V / 2).
//
// The conversions are linear in (egrad, ehess), so Hess L[v] is formed from the Euclidean gradient / Hessian of
// the Lagrangian (Phi_L = G_A + sum_i y_i Phi_i), which is the reference's `do_euclidean_lincomb` form
// (RIPTRM.py:497-517) and equals its default per-constraint sum up to rounding.  With tangent v,
//   G*_x[v]_i = <grad s_i, v>_x = -coef_i dA_rc      (the SPD metric cancels against P sym(eg) P)
//   G_x(w)    = -rgrad( pull( sum_i w_i Phi_i ) ).
//
// One warp per instance.  A point / tangent vector is WVec<3>: slot k = component (J, R, Q), entry (i,j) of the
// d x d matrix on lane i*d + j (d*d <= 32); constraint i lives on lane i (m <= 32).  Products go through the warp's
// shared-memory scratch (smallmat.cuh); X, X' and X X' are staged once per CTA.
#pragma once
#include "smallmat.cuh"
#include "solver_warp.cuh"

namespace riptrm {

struct StableIdFam {
    static constexpr int K = 3;
    static constexpr int MK = 1;
    static constexpr int DMAX = 5;
    static constexpr int kComponents = 3;
    static constexpr int kSlots = 8;
    using Vec = WVec<3>;
    using CVec = WVec<1>;
    using LM = double;  // "lane matrix": a d x d matrix with one entry per lane

    struct Ctx {
        int d, dd, m, N;
        double h;
        bool embedded;
        const double* X;    // shared memory [d][N]
        const double* XP;   // shared memory [d][N]
        const double* XXt;  // shared memory [d][d]
        double* sc;         // scratch: kSlots slots of 32 doubles + E [d][N]
        double* E;
        bool generic_tcg;   // measurement switch: tCG in the reference's operation order on unwhitened vectors (tcg_generic)
        // this lane's constraint (lane < m)
        int kind, rc;
        double ca, cb;
        // lane geometry of a d x d lane matrix: row / column of this lane's entry, the lane holding the transposed entry
        int li, lj, tlane;
        // scatter: the (at most 4) constraints whose entry (r, c) is this lane's, in constraint order (-1: none); when an
        // entry is hit by more than 4 constraints nsrc = -1 and scatter falls back to the serial shared-memory form
        int src[4];
        int nsrc;
    };
    struct Pt {
        Vec x;
        CVec s;
        double cost;
        LM A, GA, JmR;       // (J-R) Q, df/dA, J - R
        LM Rinv, Qinv;
        LM LR, LRi, LQ, LQi; // Cholesky factors of R, Q and their inverses (R^-1 = LRi' LRi): one factorisation per point
                             // serves the metric, the retraction, the positive-definiteness test and the whitened tCG
        CVec coef;           // Phi_i = coef_i E_rc
        bool spd_ok;
    };
    struct Step {
        Vec c;
        CVec ys;
        LM PhiL;             // G_A + sum_i y_i Phi_i
    };

    static __host__ __device__ constexpr int smem_doubles(int d, int N) { return kSlots * 32 + 3 * d * N + d * d; }

    template <class Params>
    static __device__ __forceinline__ Ctx make_ctx(const Params& P, const DevOpts& o, double* smem) {
        Ctx c;
        c.d = P.n;
        c.dd = P.n * P.n;
        c.m = P.m;
        c.N = P.N;
        c.h = P.hstep;
        c.embedded = o.is_euclidean_embedded != 0;
        c.generic_tcg = P.generic_tcg != 0;
        c.sc = smem;
        double* X = smem + kSlots * 32;
        double* XP = X + c.d * c.N;
        c.E = XP + c.d * c.N;
        double* XXt = c.E + c.d * c.N;
        for (int e = lane_id(); e < c.d * c.N; e += 32) {
            X[e] = P.Xd[e];
            XP[e] = P.XPd[e];
        }
        __syncwarp();
        sm::mm(XXt, X, X, c.d, c.N, c.d, false, true);
        c.X = X;
        c.XP = XP;
        c.XXt = XXt;
        c.kind = 0;
        c.rc = 0;
        c.ca = c.cb = 0.0;
        if (lane_id() < c.m) {
            const double* row = P.conspec + lane_id() * 5;
            c.kind = (int)row[0];
            c.rc = (int)row[1] * c.d + (int)row[2];
            c.ca = row[3];
            c.cb = row[4];
        }
        const int l = lane_id();
        c.li = l / c.d;
        c.lj = l - c.li * c.d;
        c.tlane = (l < c.dd) ? c.lj * c.d + c.li : l;
        int cnt = 0, over = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) c.src[q] = -1;
        for (int i = 0; i < c.m; ++i) {
            const int rci = __shfl_sync(kFull, c.rc, i);
            if (rci == l && l < c.dd) {
                if (cnt < 4) c.src[cnt] = i;
                else over = 1;
                ++cnt;
            }
        }
        c.nsrc = __any_sync(kFull, over) ? -1 : 0;
        return c;
    }

    static __device__ __forceinline__ double* slot(const Ctx& c, int i) { return c.sc + 32 * i; }
    static __device__ __forceinline__ bool on(const Ctx& c) { return lane_id() < c.dd; }
    static __device__ __forceinline__ bool active(const Ctx& c, int) { return on(c); }
    static __device__ __forceinline__ bool cactive(const Ctx& c, int) { return lane_id() < c.m; }
    static __device__ __forceinline__ int dim(const Ctx& c) { return c.d * (c.d - 1) / 2 + c.d * (c.d + 1); }
    static __device__ __forceinline__ int num_constraints(const Ctx& c) { return c.m; }
    static __device__ __forceinline__ double typical_dist(const Ctx& c) { return sqrt((double)dim(c)); }
    static __device__ __forceinline__ bool domain_ok(const Ctx&, const Pt&) { return true; }

    static __device__ __forceinline__ Vec load_x(const Ctx& c, const double* g) {
        Vec r;
#pragma unroll
        for (int k = 0; k < 3; ++k) r.v[k] = on(c) ? g[k * c.dd + lane_id()] : 0.0;
        return r;
    }
    static __device__ __forceinline__ void store_x(const Ctx& c, double* g, const Vec& v) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (on(c)) g[k * c.dd + lane_id()] = v.v[k];
    }
    static __device__ __forceinline__ CVec load_y(const Ctx& c, const double* g) {
        CVec r;
        r.v[0] = (lane_id() < c.m) ? g[lane_id()] : 0.0;
        return r;
    }
    static __device__ __forceinline__ void store_y(const Ctx& c, double* g, const CVec& v) {
        if (lane_id() < c.m) g[lane_id()] = v.v[0];
    }

    // ---- lane-matrix helpers ---------------------------------------------------------------------------
    static __device__ __forceinline__ void put(const Ctx& c, int s, LM a) {
        slot(c, s)[lane_id()] = a;
        __syncwarp();
    }
    // op(A) op(B) for d x d lane matrices.  Out of line (see smallmat.cuh): called from ~150 sites of the solve.
    template <bool TA, bool TB>
    static __device__ __forceinline__ double dot5(const double* A, const double* B, int i, int j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 5; ++k) s = fma(TA ? A[k * 5 + i] : A[i * 5 + k], TB ? B[j * 5 + k] : B[k * 5 + j], s);
        return s;
    }
    static __device__ __noinline__ LM mul_impl(double* sc, int d, LM a, LM b, bool tA, bool tB) {
        const int l = lane_id();
        sc[l] = a;
        sc[32 + l] = b;
        __syncwarp();
        double s = 0.0;
        const double* A = sc;
        const double* B = sc + 32;
        if (d == 5) {
            // the reference's dimension: unrolled, same order of additions as the loop below
            if (l < 25) {
                const int i = l / 5, j = l - i * 5;
                s = tA ? (tB ? dot5<true, true>(A, B, i, j) : dot5<true, false>(A, B, i, j))
                       : (tB ? dot5<false, true>(A, B, i, j) : dot5<false, false>(A, B, i, j));
            }
        } else if (l < d * d) {
            const int i = l / d, j = l - i * d;
            for (int k = 0; k < d; ++k) s = fma(tA ? A[k * d + i] : A[i * d + k], tB ? B[j * d + k] : B[k * d + j], s);
        }
        __syncwarp();
        return s;
    }
    // d == 5 (the reference's dimension): inlined, the transposition flags are compile-time constants at every call site and
    // the row / column offsets come from the lane geometry in Ctx -- 2 STS + 10 LDS + 5 DFMA + 2 barriers, no index arithmetic
    // (the out-of-line general version spent 55 instructions per product, most of them integer division by a runtime d;
    // profiles/r02b_stableid_*).  Same order of additions as mul_impl.
    static __device__ __forceinline__ LM mul(const Ctx& c, LM a, LM b, bool tA = false, bool tB = false) {
        if (c.d != 5) return mul_impl(c.sc, c.d, a, b, tA, tB);
        const int l = lane_id();
        double* sc = c.sc;
        sc[l] = a;
        sc[32 + l] = b;
        __syncwarp();
        double s = 0.0;
        if (l < 25) {
            const double* A = tA ? sc + c.li : sc + 5 * c.li;
            const double* B = tB ? sc + 32 + 5 * c.lj : sc + 32 + c.lj;
#pragma unroll
            for (int k = 0; k < 5; ++k) s = fma(tA ? A[5 * k] : A[k], tB ? B[k] : B[5 * k], s);
        }
        __syncwarp();
        return s;
    }
    static __device__ __noinline__ LM transpose_impl(double* sc, int d, LM a) {
        const int l = lane_id();
        sc[l] = a;
        __syncwarp();
        double r = 0.0;
        if (l < d * d) {
            const int i = l / d, j = l - i * d;
            r = sc[j * d + i];
        }
        __syncwarp();
        return r;
    }
    // the transposed entry lives on lane `tlane`: one 64-bit shuffle (was a shared-memory round trip, 57 instructions)
    static __device__ __forceinline__ LM transpose(const Ctx& c, LM a) { return __shfl_sync(kFull, a, c.tlane); }
    static __device__ __forceinline__ LM symm(const Ctx& c, LM a) { return 0.5 * (a + transpose(c, a)); }
    static __device__ __forceinline__ LM skew(const Ctx& c, LM a) { return 0.5 * (a - transpose(c, a)); }
    static __device__ __forceinline__ LM inverse(const Ctx& c, LM a, bool& ok) {
        put(c, 2, a);
        ok = sm::inverse<DMAX>(slot(c, 3), slot(c, 2), c.d);
        const double r = (on(c) && ok) ? slot(c, 3)[lane_id()] : 0.0;
        __syncwarp();
        return r;
    }
    // sum_i wc_i E_{r_i c_i} (+ base) as a lane matrix; contributions added in constraint order
    static __device__ __noinline__ LM scatter_impl(double* sc, int m, int dd, int my_rc, double my_wc, LM base) {
        double* P = sc + 32 * 4;
        double* wc = sc + 32 * 5;
        int* rc = reinterpret_cast<int*>(sc + 32 * 6);
        const int l = lane_id();
        P[l] = base;
        wc[l] = (l < m) ? my_wc : 0.0;
        rc[l] = my_rc;
        __syncwarp();
        if (l == 0)
            for (int i = 0; i < m; ++i) P[rc[i]] = P[rc[i]] + wc[i];
        __syncwarp();
        const double r = (l < dd) ? P[l] : 0.0;
        __syncwarp();
        return r;
    }
    static __device__ __forceinline__ LM scatter(const Ctx& c, const CVec& w, const CVec& coef, LM base) {
        const double wc = w.v[0] * coef.v[0];
        if (c.nsrc < 0) return scatter_impl(c.sc, c.m, c.dd, c.rc, wc, base);
        // every entry gathers its constraints' weights in constraint order: the same sums as the serial form
        double r = base;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double v = __shfl_sync(kFull, wc, c.src[q] < 0 ? 0 : c.src[q]);
            if (c.src[q] >= 0) r = r + v;
        }
        return (lane_id() < c.dd) ? r : 0.0;
    }
    // value of a lane matrix at this lane's constraint entry (r_i, c_i)
    static __device__ __noinline__ double at_constraint_impl(double* sc, int m, int my_rc, LM a) {
        sc[lane_id()] = a;
        __syncwarp();
        const double r = (lane_id() < m) ? sc[my_rc] : 0.0;
        __syncwarp();
        return r;
    }
    static __device__ __forceinline__ double at_constraint(const Ctx& c, LM a) {
        const double v = __shfl_sync(kFull, a, c.rc);      // entry (r, c) of a lane matrix lives on lane r d + c
        return (lane_id() < c.m) ? v : 0.0;
    }

    // Euclidean -> Riemannian gradient, componentwise
    static __device__ __forceinline__ Vec egrad2rgrad(const Ctx& c, const Pt& pt, LM egJ, LM egR, LM egQ) {
        Vec r;
        r.v[0] = skew(c, egJ);
        r.v[1] = mul(c, mul(c, pt.x.v[1], symm(c, egR)), pt.x.v[1]);
        r.v[2] = mul(c, mul(c, pt.x.v[2], symm(c, egQ)), pt.x.v[2]);
        return r;
    }
    // pull(Phi) = [Phi Q', -Phi Q', (J-R)' Phi], then egrad2rgrad
    static __device__ __forceinline__ Vec rgrad_of_phi(const Ctx& c, const Pt& pt, LM Phi) {
        const LM gJ = mul(c, Phi, pt.x.v[2], false, true);
        const LM gQ = mul(c, pt.JmR, Phi, true, false);
        return egrad2rgrad(c, pt, gJ, -gJ, gQ);
    }

    static __device__ __forceinline__ void eval_point(const Ctx& c, const Vec& x, Pt& pt) {
        pt.x = x;
        pt.JmR = x.v[0] - x.v[1];
        pt.A = mul(c, pt.JmR, x.v[2]);
        // E = X' - (I + h A) X  (d x N), cost = tr(E E')/N, G_A = -2h E X'/N
        put(c, 2, pt.A);
        const double* A = slot(c, 2);
        const int d = c.d, N = c.N;
        double part = 0.0;
        for (int e = lane_id(); e < d * N; e += 32) {
            const int i = e / N, t = e - i * N;
            double ax = 0.0;
            for (int k = 0; k < d; ++k) ax = fma(((i == k) ? 1.0 : 0.0) + c.h * A[i * d + k], c.X[k * N + t], ax);
            const double ev = c.XP[e] - ax;
            c.E[e] = ev;
            part = fma(ev, ev, part);
        }
        __syncwarp();
        pt.cost = wsum(part) / (double)N;
        double ga = 0.0;
        if (on(c)) {
            const int i = lane_id() / d, j = lane_id() - i * d;
            double s = 0.0;
            for (int t = 0; t < N; ++t) s = fma(c.E[i * N + t], c.X[j * N + t], s);
            ga = (-2.0 * c.h) * s / (double)N;
        }
        __syncwarp();
        pt.GA = ga;
        // constraints
        const double Arc = at_constraint(c, pt.A);
        double g = 0.0, coef = 0.0;
        if (lane_id() < c.m) {
            if (c.kind == 0) {
                g = -Arc + c.ca;
                coef = -1.0;
            } else if (c.kind == 1) {
                g = Arc - c.ca;
                coef = 1.0;
            } else {
                g = -((Arc - c.ca) * (Arc - c.ca)) + c.cb;
                coef = -2.0 * (Arc - c.ca);
            }
        }
        pt.s.v[0] = -g;
        pt.coef.v[0] = coef;
        // one Cholesky factorisation per SPD block: positive definite iff it succeeds, P^-1 = L^-T L^-1
        const bool okR = chol_and_inverse(c.sc, c.d, x.v[1], pt.LR, pt.LRi);
        const bool okQ = chol_and_inverse(c.sc, c.d, x.v[2], pt.LQ, pt.LQi);
        pt.Rinv = mul(c, pt.LRi, pt.LRi, true, false);
        pt.Qinv = mul(c, pt.LQi, pt.LQi, true, false);
        pt.spd_ok = okR && okQ;
    }

    // <a, b>_x = <aJ, bJ> + tr(R^-1 aR R^-1 bR) + tr(Q^-1 aQ Q^-1 bQ): per-lane partial
    // (taking this and the gradient conversions out of line as well costs registers -- 12 instead of 16 warps per SM --
    // and was slower: 442 vs 367 ms for the 2048-pair sweep)
    static __device__ __forceinline__ double inner_partial(const Ctx& c, const Pt& pt, const Vec& a, const Vec& b) {
        const LM tR = mul(c, mul(c, pt.Rinv, a.v[1]), pt.Rinv);
        const LM tQ = mul(c, mul(c, pt.Qinv, a.v[2]), pt.Qinv);
        const LM bRt = transpose(c, b.v[1]), bQt = transpose(c, b.v[2]);
        return (a.v[0] * b.v[0] + tR * bRt) + tQ * bQt;
    }
    static __device__ __forceinline__ double inner(const Ctx& c, const Pt& pt, const Vec& a, const Vec& b) {
        return wsum(inner_partial(c, pt, a, b));
    }

    static __device__ __forceinline__ Vec project(const Ctx& c, const Pt&, const Vec& v) {
        Vec r;
        r.v[0] = skew(c, v.v[0]);
        r.v[1] = symm(c, v.v[1]);
        r.v[2] = symm(c, v.v[2]);
        return r;
    }

    static __device__ __forceinline__ void begin_step(const Ctx& c, const Pt& pt, const CVec& y, double mu, Step& st) {
        CVec w;
        const bool onc = cactive(c, 0);
        w.v[0] = onc ? mu * (1.0 / pt.s.v[0]) : 0.0;
        st.ys.v[0] = onc ? y.v[0] / pt.s.v[0] : 0.0;
        st.PhiL = scatter(c, y, pt.coef, pt.GA);
        const Vec gradf = rgrad_of_phi(c, pt, pt.GA);
        const Vec gw = rgrad_of_phi(c, pt, scatter(c, w, pt.coef, 0.0));  // = -G_x(mu/s)
#pragma unroll
        for (int k = 0; k < 3; ++k) st.c.v[k] = gradf.v[k] + gw.v[k];      // grad f - G_x(mu/s)
    }

    // dA along v
    static __device__ __forceinline__ LM dA_of(const Ctx& c, const Pt& pt, const Vec& v) {
        return mul(c, v.v[0] - v.v[1], pt.x.v[2]) + mul(c, pt.JmR, v.v[2]);
    }

    // 'is_euclidean_embedded' (RIPTRM.py:553-571): G*[v]_i = <-egrad g_i, v>_x with the EUCLIDEAN gradient inside the
    // manifold's inner product.  On the SPD components that metric is tr(P^-1 A P^-1 B), so with -egrad g_i = -coef_i pull(E_rc)
    // the entry of dA is taken with v_R -> R^-1 v_R R^-1 and v_Q -> Q^-1 v_Q Q^-1 (not the same operator as the Riemannian
    // form on this manifold -- the option is meant for submanifolds with the embedded metric -- but it is what the reference
    // evaluates when the key is set).
    static __device__ __forceinline__ LM dA_embedded(const Ctx& c, const Pt& pt, const Vec& v) {
        const LM wR = mul(c, mul(c, pt.Rinv, v.v[1]), pt.Rinv);
        const LM wQ = mul(c, mul(c, pt.Qinv, v.v[2]), pt.Qinv);
        return mul(c, v.v[0] - wR, pt.x.v[2]) + mul(c, pt.JmR, wQ);
    }

    static __device__ __forceinline__ CVec gadj(const Ctx& c, const Pt& pt, const Vec& v) {
        CVec g;
        g.v[0] = -pt.coef.v[0] * at_constraint(c, c.embedded ? dA_embedded(c, pt, v) : dA_of(c, pt, v));
        return g;
    }

    static __device__ __forceinline__ Vec Hw(const Ctx& c, const Pt& pt, const CVec& y, const Step& st, const Vec& v) {
        const LM dA = dA_of(c, pt, v);
        // Euclidean Hessian of the Lagrangian along v: dPhi_L = dG_A + sum_i y_i dcoef_i E_rc
        put(c, 7, dA);
        double dga = 0.0;
        if (on(c)) {
            const int d = c.d, i = lane_id() / d, j = lane_id() - i * d;
            double s = 0.0;
            for (int k = 0; k < d; ++k) s = fma(slot(c, 7)[i * d + k], c.XXt[k * d + j], s);
            dga = (2.0 * c.h * c.h) * s / (double)c.N;
        }
        __syncwarp();
        const double dArc = at_constraint(c, dA);
        CVec dcoef, ones;
        dcoef.v[0] = (lane_id() < c.m && c.kind == 2) ? -2.0 * dArc : 0.0;
        ones.v[0] = 1.0;
        CVec ydc;
        ydc.v[0] = y.v[0] * dcoef.v[0];
        const LM dPhiL = scatter(c, ydc, ones, dga);
        // pull_d: hJ = dPhi Q' + Phi dQ' ; hR = -hJ ; hQ = (dJ - dR)' Phi + (J - R)' dPhi
        const LM hJ = mul(c, dPhiL, pt.x.v[2], false, true) + mul(c, st.PhiL, v.v[2], false, true);
        const LM hQ = mul(c, v.v[0] - v.v[1], st.PhiL, true, false) + mul(c, pt.JmR, dPhiL, true, false);
        // Euclidean gradient of the Lagrangian
        const LM gJ = mul(c, st.PhiL, pt.x.v[2], false, true);
        const LM gQ = mul(c, pt.JmR, st.PhiL, true, false);
        Vec hl;
        hl.v[0] = skew(c, hJ);
        {
            const LM R = pt.x.v[1];
            hl.v[1] = mul(c, mul(c, R, symm(c, -hJ)), R) + symm(c, mul(c, mul(c, v.v[1], symm(c, -gJ)), R));
            const LM Q = pt.x.v[2];
            hl.v[2] = mul(c, mul(c, Q, symm(c, hQ)), Q) + symm(c, mul(c, mul(c, v.v[2], symm(c, gQ)), Q));
        }
        // condensed barrier term G_x((y/s) * G*[v]) = -rgrad(pull(sum_i w_i Phi_i))
        CVec w;
        const double gArc = c.embedded ? at_constraint(c, dA_embedded(c, pt, v)) : dArc;
        w.v[0] = st.ys.v[0] * (-pt.coef.v[0] * gArc);
        const Vec gw = rgrad_of_phi(c, pt, scatter(c, w, pt.coef, 0.0));
        Vec out;
#pragma unroll
        for (int k = 0; k < 3; ++k) out.v[k] = hl.v[k] - gw.v[k];
        return out;
    }

    // ---- tCG in whitened coordinates ---------------------------------------------------------------------------------------
    // With R = LR LR', Q = LQ LQ' (Cholesky) the map v -> vh = (vJ, LR^-1 vR LR^-T, LQ^-1 vQ LQ^-T) is an isometry from the
    // tangent space with the product metric (Frobenius + two affine-invariant traces) onto matrices with the FROBENIUS inner
    // product.  Steihaug-Toint tCG (RIPTRM.py:41-216) on vh is the same iteration with
    //   * every inner product a lane-wise product and one butterfly (tcg_generic: two whitening congruences per operand);
    //   * the Hessian-vector product conjugated: LR'[R-block of the Euclidean part]LR has the manifold's P . P cancelled, the
    //     curvature term sym(V sym(eg) P) becomes sym(vh Gh) with Gh = LR' sym(eg) LR cached per call, and the barrier term
    //     folds into the Euclidean Hessian of the Lagrangian before the pull-back (both are linear in Phi):
    //         dPhi' = dG_A + sum_i (y_i dcoef_i - w_i coef_i) E_rc,   w_i = (y_i / s_i) G*[v]_i
    //     17 5 x 5 products per iteration instead of 22 + 16 (4 per inner product).
    // Same algorithm, same exits; rounding differs from tcg_generic by the usual few ulp.
    struct White {
        LM LR, LRi, LQ, LQi;   // Cholesky factors of R, Q and their inverses (the point's)
        LM GR, GQ;             // LR' sym(-gJ) LR, LQ' sym(gQ) LQ: whitened Riemannian gradient blocks of the Lagrangian
    };
    static __device__ __forceinline__ void white_setup(const Ctx& c, const Pt& pt, const Step& st, White& w) {
        w.LR = pt.LR;
        w.LRi = pt.LRi;
        w.LQ = pt.LQ;
        w.LQi = pt.LQi;
        const LM gJ = mul(c, st.PhiL, pt.x.v[2], false, true);
        const LM gQ = mul(c, pt.JmR, st.PhiL, true, false);
        w.GR = mul(c, mul(c, w.LR, symm(c, -gJ), true, false), w.LR);
        w.GQ = mul(c, mul(c, w.LQ, symm(c, gQ), true, false), w.LQ);
    }
    static __device__ __forceinline__ Vec whiten(const Ctx& c, const White& w, const Vec& v) {
        Vec r;
        r.v[0] = v.v[0];
        r.v[1] = mul(c, mul(c, w.LRi, v.v[1]), w.LRi, false, true);
        r.v[2] = mul(c, mul(c, w.LQi, v.v[2]), w.LQi, false, true);
        return r;
    }
    static __device__ __forceinline__ Vec unwhiten(const Ctx& c, const White& w, const Vec& v) {
        Vec r;
        r.v[0] = v.v[0];
        r.v[1] = mul(c, mul(c, w.LR, v.v[1]), w.LR, false, true);
        r.v[2] = mul(c, mul(c, w.LQ, v.v[2]), w.LQ, false, true);
        return r;
    }
    // whitened Hw: vh -> whiten(Hw[unwhiten(vh)])
    static __device__ __forceinline__ Vec Hw_white(const Ctx& c, const Pt& pt, const CVec& y, const Step& st, const White& w,
                                                   const Vec& vh) {
        const LM vR = mul(c, mul(c, w.LR, vh.v[1]), w.LR, false, true);
        const LM vQ = mul(c, mul(c, w.LQ, vh.v[2]), w.LQ, false, true);
        const LM D1 = vh.v[0] - vR;
        const LM dA = mul(c, D1, pt.x.v[2]) + mul(c, pt.JmR, vQ);
        put(c, 7, dA);
        double dga = 0.0;
        if (on(c)) {
            const int d = c.d, i = lane_id() / d, j = lane_id() - i * d;
            double s = 0.0;
            for (int k = 0; k < d; ++k) s = fma(slot(c, 7)[i * d + k], c.XXt[k * d + j], s);
            dga = (2.0 * c.h * c.h) * s / (double)c.N;
        }
        __syncwarp();
        const double dArc = at_constraint(c, dA);
        double gArc = dArc;
        if (c.embedded) {
            Vec v;
            v.v[0] = vh.v[0];
            v.v[1] = vR;
            v.v[2] = vQ;
            gArc = at_constraint(c, dA_embedded(c, pt, v));
        }
        CVec wc, ones;
        const double dcoef = (lane_id() < c.m && c.kind == 2) ? -2.0 * dArc : 0.0;
        const double wi = st.ys.v[0] * (-pt.coef.v[0] * gArc);
        wc.v[0] = y.v[0] * dcoef - wi * pt.coef.v[0];
        ones.v[0] = 1.0;
        const LM dPhi = scatter(c, wc, ones, dga);
        const LM hJ = mul(c, dPhi, pt.x.v[2], false, true) + mul(c, st.PhiL, vQ, false, true);
        const LM hQ = mul(c, D1, st.PhiL, true, false) + mul(c, pt.JmR, dPhi, true, false);
        const LM hJt = transpose(c, hJ);
        Vec out;
        out.v[0] = 0.5 * (hJ - hJt);
        out.v[1] = mul(c, mul(c, w.LR, -0.5 * (hJ + hJt), true, false), w.LR) + symm(c, mul(c, vh.v[1], w.GR));
        out.v[2] = mul(c, mul(c, w.LQ, symm(c, hQ), true, false), w.LQ) + symm(c, mul(c, vh.v[2], w.GQ));
        return out;
    }
    static __device__ __forceinline__ double dotw(const Vec& a, const Vec& b) {
        return fma(a.v[2], b.v[2], fma(a.v[1], b.v[1], a.v[0] * b.v[0]));
    }

    static __device__ __forceinline__ TcgResult tcg(const Ctx& ctx, const DevOpts& o, const Pt& pt, const CVec& y,
                                                    const Step& st, double Delta, Vec& eta, Vec& Heta) {
        if (ctx.generic_tcg) return tcg_generic<StableIdFam>(ctx, o, pt, y, st, Delta, eta, Heta);
        White w;
        white_setup(ctx, pt, st, w);
        const Vec ch = whiten(ctx, w, st.c);
        Vec e = wzero<3>(), He = wzero<3>();                  // :47
        Vec r = ch, delta;                                    // :48
#pragma unroll
        for (int k = 0; k < 3; ++k) delta.v[k] = -r.v[k];     // :71
        double r_r = wsum(dotw(r, r));                        // :56
        const double norm_r0 = sqrt(r_r);
        double z_r = r_r, d_Pd = r_r, e_Pe = 0.0, e_Pd = 0.0, model_value = 0.0;
        TcgResult res;
        res.stop = RIPTRM_TCG_MAX_INNER_ITER;                 // :95
        const int maxinner = o.tcg_maxinner < 0 ? dim(ctx) : o.tcg_maxinner;
        const double Delta2 = Delta * Delta;
        const double nr_theta = (o.tcg_theta == 1.0) ? norm_r0 : pow(norm_r0, o.tcg_theta);
        const double target = norm_r0 * fmin(nr_theta, o.tcg_kappa);
        int j = 0;
        for (; j < maxinner; ++j) {                           // :98
            const Vec Hd = Hw_white(ctx, pt, y, st, w, delta);   // :100
            const double d_Hd = wsum(dotw(delta, Hd));        // :103
            double alpha = 0.0, e_Pe_new = e_Pe;
            if (d_Hd != 0.0) {                                // :106-114
                alpha = z_r / d_Hd;
                e_Pe_new = (e_Pe + (2.0 * alpha) * e_Pd) + (alpha * alpha) * d_Pd;
            }
            if (d_Hd <= 0.0 || e_Pe_new >= Delta2) {          // :118
                const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta2 - e_Pe))) / d_Pd;   // :123-125
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    e.v[k] = e.v[k] + tau * delta.v[k];       // :127
                    He.v[k] = He.v[k] + tau * Hd.v[k];        // :132
                }
                res.stop = (d_Hd <= 0.0) ? RIPTRM_TCG_NEGATIVE_CURVATURE : RIPTRM_TCG_EXCEEDED_TR;
                ++j;
                break;
            }
            e_Pe = e_Pe_new;                                  // :149
            Vec ne, nHe, rn;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                ne.v[k] = e.v[k] + alpha * delta.v[k];        // :150
                nHe.v[k] = He.v[k] + alpha * Hd.v[k];         // :154
                rn.v[k] = r.v[k] + alpha * Hd.v[k];           // :172
            }
            double ip_ec = dotw(ne, ch), ip_eh = dotw(ne, nHe), ip_rr = dotw(rn, rn);
            wsum3(ip_ec, ip_eh, ip_rr);
            const double new_model = ip_ec + 0.5 * ip_eh;     // :86-87, :162
            if (new_model >= model_value) {                   // :163
                res.stop = RIPTRM_TCG_MODEL_INCREASED;
                ++j;
                break;
            }
            e = ne;                                           // :167-169
            He = nHe;
            model_value = new_model;
            r = rn;
            r_r = ip_rr;                                      // :175
            if (j >= o.tcg_mininner && sqrt(r_r) <= target) { // :183-191
                res.stop = (o.tcg_kappa < nr_theta) ? RIPTRM_TCG_REACHED_TARGET_LINEAR
                                                    : RIPTRM_TCG_REACHED_TARGET_SUPERLINEAR;
                ++j;
                break;
            }
            const double beta = r_r / z_r;                    // :200-205
            z_r = r_r;
            // :206, :210 -- the re-projection onto the tangent space is skew / sym of the blocks in these coordinates
            const LM dJ = -r.v[0] + beta * delta.v[0], dR = -r.v[1] + beta * delta.v[1], dQ = -r.v[2] + beta * delta.v[2];
            delta.v[0] = skew(ctx, dJ);
            delta.v[1] = symm(ctx, dR);
            delta.v[2] = symm(ctx, dQ);
            e_Pd = beta * (e_Pd + alpha * d_Pd);              // :213
            d_Pd = z_r + (beta * beta) * d_Pd;                // :214
        }
        eta = unwhiten(ctx, w, e);
        Heta = unwhiten(ctx, w, He);
        res.iters = j;
        res.model_value = model_value;
        return res;
    }

    // ---- Exact_RepMat: orthonormal tangent basis of the product (riptrm_b200/basis.py) ----------------------------------------
    //   Skew(d):           (E_ab - E_ba) / sqrt 2, a < b                                  d (d-1) / 2 coordinates
    //   SPD(d) at P = LL': L E L' with E = E_aa, (E_ab + E_ba) / sqrt 2                   d (d+1) / 2 coordinates each
    // (<L E L', L F L'>_P = tr(P^-1 L E L' P^-1 L F L') = tr(E F): orthonormal in the affine-invariant metric.)
    // Coordinate order: skew pairs (a < b, row-major), then for R and for Q: the d diagonal entries, the pairs.
    struct Coord {
        LM LR, LRi, LQ, LQi;   // Cholesky factors and their inverses, one entry per lane
    };
    // Cholesky factor L (P = L L') and L^-1 of a d x d lane matrix, every lane redundantly on a private copy (25 flops of
    // latency-bound serial work either way); returns false when P is not positive definite.  D > 0: compile-time size, the
    // private matrices stay in registers (the runtime-d version indexes local memory).
    template <int D>
    static __device__ __forceinline__ bool chol_core(const double* sc, int d_rt, int li, int lj, LM& L_out, LM& Li_out) {
        constexpr int DM = (D > 0) ? D : DMAX;
        const int d = (D > 0) ? D : d_rt;
        double L[DM][DM], Li[DM][DM];
        bool ok = true;
#pragma unroll
        for (int i = 0; i < DM; ++i)
#pragma unroll
            for (int j = 0; j < DM; ++j) L[i][j] = Li[i][j] = 0.0;
#pragma unroll
        for (int i = 0; i < DM; ++i) {
            if (i >= d) break;
#pragma unroll
            for (int j = 0; j < DM; ++j) {
                if (j > i) break;
                double s = sc[i * d + j];
#pragma unroll
                for (int k = 0; k < DM; ++k) {
                    if (k >= j) break;
                    s -= L[i][k] * L[j][k];
                }
                if (i == j) {
                    ok = ok && (s > 0.0);
                    L[i][i] = sqrt(s);
                } else {
                    L[i][j] = s / L[j][j];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < DM; ++j) {          // forward substitution L Li = I, column by column
            if (j >= d) break;
#pragma unroll
            for (int i = 0; i < DM; ++i) {
                if (i < j) continue;
                if (i >= d) break;
                double s = (i == j) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < DM; ++k) {
                    if (k < j) continue;
                    if (k >= i) break;
                    s -= L[i][k] * Li[k][j];
                }
                Li[i][j] = s / L[i][i];
            }
        }
        double lo = 0.0, lio = 0.0;
#pragma unroll
        for (int i = 0; i < DM; ++i)
#pragma unroll
            for (int j = 0; j < DM; ++j)
                if (i == li && j == lj) {
                    lo = L[i][j];
                    lio = Li[i][j];
                }
        L_out = lo;
        Li_out = lio;
        return ok;
    }
    static __device__ __noinline__ bool chol_and_inverse(double* sc, int d, LM P, LM& L_out, LM& Li_out) {
        const int l = lane_id();
        sc[l] = P;
        __syncwarp();
        const int li = l / d, lj = l - li * d;
        const bool ok = (d == 5) ? chol_core<5>(sc, d, li, lj, L_out, Li_out) : chol_core<0>(sc, d, li, lj, L_out, Li_out);
        __syncwarp();
        if (l >= d * d) L_out = Li_out = 0.0;
        return ok;
    }
    static __device__ __forceinline__ void coord_setup(const Ctx&, const Pt& pt, Coord& cc) {
        cc.LR = pt.LR;
        cc.LRi = pt.LRi;
        cc.LQ = pt.LQ;
        cc.LQi = pt.LQi;
    }
    static __device__ __forceinline__ int pair_index(int d, int a, int b) { return a * d - (a * (a + 1)) / 2 + (b - a - 1); }
    // symmetric / skew lane matrix from packed coordinates (diag first, then pairs scaled by 1 / sqrt 2)
    static __device__ __forceinline__ LM unpack(const Ctx& c, const double* coef, bool skewpart) {
        const int l = lane_id(), d = c.d;
        if (l >= c.dd) return 0.0;
        const int i = l / d, j = l - i * d;
        const double r = 0.70710678118654752440;
        if (skewpart) {
            if (i == j) return 0.0;
            return (i < j) ? coef[pair_index(d, i, j)] * r : -(coef[pair_index(d, j, i)] * r);
        }
        if (i == j) return coef[i];
        return coef[d + pair_index(d, i < j ? i : j, i < j ? j : i)] * r;
    }
    static __device__ __forceinline__ Vec from_coords(const Ctx& c, const Pt&, const Coord& cc, const double* coef) {
        const int ns = c.d * (c.d - 1) / 2, np = c.d * (c.d + 1) / 2;
        Vec v;
        v.v[0] = unpack(c, coef, true);
        const LM KR = unpack(c, coef + ns, false), KQ = unpack(c, coef + ns + np, false);
        v.v[1] = mul(c, mul(c, cc.LR, KR), cc.LR, false, true);
        v.v[2] = mul(c, mul(c, cc.LQ, KQ), cc.LQ, false, true);
        return v;
    }
    static __device__ __forceinline__ void pack(const Ctx& c, LM a, double* out, bool skewpart) {
        put(c, 7, a);
        const int l = lane_id(), d = c.d;
        const double r = 0.70710678118654752440;
        const double* A = slot(c, 7);
        if (l < c.dd) {
            const int i = l / d, j = l - i * d;
            if (skewpart) {
                if (i < j) out[pair_index(d, i, j)] = (A[i * d + j] - A[j * d + i]) * r;
            } else {
                if (i == j) out[i] = A[i * d + i];
                else if (i < j) out[d + pair_index(d, i, j)] = (A[i * d + j] + A[j * d + i]) * r;
            }
        }
        __syncwarp();
    }
    static __device__ __forceinline__ void to_coords(const Ctx& c, const Pt&, const Coord& cc, const Vec& v, double* out) {
        const int ns = c.d * (c.d - 1) / 2, np = c.d * (c.d + 1) / 2;
        pack(c, v.v[0], out, true);
        pack(c, mul(c, mul(c, cc.LRi, v.v[1]), cc.LRi, false, true), out + ns, false);
        pack(c, mul(c, mul(c, cc.LQi, v.v[2]), cc.LQi, false, true), out + ns + np, false);
    }

    static __device__ __forceinline__ Vec retract(const Ctx& c, const Pt& pt, const Vec& dx) {
        Vec r;
        r.v[0] = pt.x.v[0] + dx.v[0];
        r.v[1] = symm(c, (pt.x.v[1] + dx.v[1]) + mul(c, dx.v[1], mul(c, pt.Rinv, dx.v[1])) / 2.0);
        r.v[2] = symm(c, (pt.x.v[2] + dx.v[2]) + mul(c, dx.v[2], mul(c, pt.Qinv, dx.v[2])) / 2.0);
        return r;
    }

    static __device__ __forceinline__ double gradL_xy_partial(const Ctx&, const Pt&, const CVec&) { return 0.0; }
    static __device__ __forceinline__ double gradL_norm_given(const Ctx& c, const Pt& pt, const CVec& y, double) {
        const Vec g = rgrad_of_phi(c, pt, scatter(c, y, pt.coef, pt.GA));
        return sqrt(inner(c, pt, g, g));
    }
    static __device__ __forceinline__ double gradL_norm(const Ctx& c, const Pt& pt, const CVec& y) {
        return gradL_norm_given(c, pt, y, 0.0);
    }

    // src/StableIdentification/simulator.py:11-33
    static __device__ __forceinline__ double manvio(const Ctx& c, const Pt& pt) {
        const LM a = pt.x.v[0] + transpose(c, pt.x.v[0]);
        const LM b = pt.x.v[1] - transpose(c, pt.x.v[1]);
        const LM q = pt.x.v[2] - transpose(c, pt.x.v[2]);
        double pa = a * a, pb = b * b, pq = q * q;
        wsum3(pa, pb, pq);
        const double v = (sqrt(pa) + sqrt(pb)) + sqrt(pq);
        return pt.spd_ok ? v : CUDART_INF;
    }

    // eigenvalues of P^{-1/2} B P^{-1/2} (= those of chol(P)^-1 B chol(P)^-T): || log w ||
    static __device__ __forceinline__ double spd_dist(const Ctx& c, LM P, LM B) {
        put(c, 2, P);
        double w[DMAX], V[DMAX][DMAX];
        sm::jacobi_eig<DMAX>(slot(c, 2), c.d, w, V);
        __syncwarp();
        double ih = 0.0;
        if (on(c)) {
            const int d = c.d, i = lane_id() / d, j = lane_id() - i * d;
            for (int k = 0; k < d; ++k) ih = fma(V[i][k] * (1.0 / sqrt(w[k])), V[j][k], ih);
        }
        const LM M = symm(c, mul(c, mul(c, ih, B), ih));
        put(c, 2, M);
        sm::jacobi_eig<DMAX>(slot(c, 2), c.d, w, V);
        __syncwarp();
        double acc = 0.0;
        for (int k = 0; k < c.d; ++k) {
            const double lg = log(w[k]);
            acc += lg * lg;
        }
        return sqrt(acc);
    }
    static __device__ __forceinline__ double dist(const Ctx& c, const Vec& xPrev, const Pt& pt) {
        const double dj = xPrev.v[0] - pt.x.v[0];
        const double nJ = sqrt(wsum(dj * dj));
        const double nR = spd_dist(c, xPrev.v[1], pt.x.v[1]);
        const double nQ = spd_dist(c, xPrev.v[2], pt.x.v[2]);
        return sqrt((nJ * nJ + nR * nR) + nQ * nQ);
    }
};

}  // namespace riptrm

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
    static constexpr bool kTcgReturnsHw = true;   // tcg() returns Hw[eta] (unwhitened) beside eta: solver_warp.cuh inner_step
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
        // d == 5: shared-space addresses of this lane's operands in the product scratch (two independent scratch pairs, so
        // that two products share one pair of barriers): write slots, row / column starts for op(A), op(B)
        // (five registers; the second scratch pair and the B operand are compile-time offsets of these)
        uint32_t sW, sAn, sAt, sBn, sBt;
        uint32_t sChol;      // 192 doubles after the scratch slots: R, Q, L and L^-1 of both
        LM XXt_lm;           // X X' as a lane matrix
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

    static constexpr int kCholDoubles = 192;
    static __host__ __device__ constexpr int smem_doubles(int d, int N) { return kSlots * 32 + kCholDoubles + 3 * d * N + d * d; }

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
        double* X = smem + kSlots * 32 + kCholDoubles;
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
        const uint32_t sb = smem_u32(smem);
        c.sW = sb + 8 * l;                    // slot 0 / 1 (+ 256) hold A / B of the first product, slots 2 / 3 (+ 512) of the second
        c.sAn = sb + 40 * c.li;
        c.sAt = sb + 8 * c.li;
        c.sBn = sb + 256 + 8 * c.lj;
        c.sBt = sb + 256 + 40 * c.lj;
        c.sChol = sb + 8 * (kSlots * 32);
        c.XXt_lm = (l < c.dd) ? XXt[l] : 0.0;
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
    // d == 5 (the reference's dimension): inlined, the transposition flags are compile-time constants at every call site, the
    // operand addresses are per-lane constants of Ctx and the accesses are explicit ld / st.shared -- 2 STS + 10 LDS + 5 DFMA
    // + 2 barriers, no index arithmetic, no divergence (lanes 25..31 compute on in-bounds scratch and return 0).  The
    // out-of-line general version spent 55 instructions per product, most of them integer division by a runtime d
    // (profiles/r02b_stableid_*).  Same order of additions as mul_impl.
    template <int Q, bool TA, bool TB>
    static __device__ __forceinline__ double dot5s(uint32_t a0, uint32_t b0) {
        constexpr int SA = TA ? 40 : 8, SB = TB ? 8 : 40, O = 512 * Q;
        double s = 0.0;
        s = fma(lds_f64_off<O>(a0), lds_f64_off<O>(b0), s);
        s = fma(lds_f64_off<O + SA>(a0), lds_f64_off<O + SB>(b0), s);
        s = fma(lds_f64_off<O + 2 * SA>(a0), lds_f64_off<O + 2 * SB>(b0), s);
        s = fma(lds_f64_off<O + 3 * SA>(a0), lds_f64_off<O + 3 * SB>(b0), s);
        s = fma(lds_f64_off<O + 4 * SA>(a0), lds_f64_off<O + 4 * SB>(b0), s);
        return s;
    }
    static __device__ __forceinline__ LM mul(const Ctx& c, LM a, LM b, bool tA = false, bool tB = false) {
        if (c.d != 5) return mul_impl(c.sc, c.d, a, b, tA, tB);
        sts_f64_off<0>(c.sW, a);
        sts_f64_off<256>(c.sW, b);
        __syncwarp();
        const uint32_t a0 = tA ? c.sAt : c.sAn, b0 = tB ? c.sBt : c.sBn;
        const double s = tA ? (tB ? dot5s<0, true, true>(a0, b0) : dot5s<0, true, false>(a0, b0))
                            : (tB ? dot5s<0, false, true>(a0, b0) : dot5s<0, false, false>(a0, b0));
        __syncwarp();
        return (lane_id() < 25) ? s : 0.0;
    }
    // Two independent products behind one pair of barriers (the R and Q blocks of every formula come in such pairs): twice the
    // instruction-level parallelism of two mul() calls at half the barrier latency.  OUT OF LINE, one copy per combination
    // of transposition flags, operands and addresses by value: with every product inlined the kernel was 35 k instructions
    // (570 KB) and a third of all stall cycles were instruction-cache misses (profiles/r02d_stableid_*).
    template <bool TA1, bool TB1, bool TA2, bool TB2>
    static __device__ __noinline__ double2 mul2_d5(uint32_t sW, uint32_t a1s, uint32_t b1s, uint32_t a2s, uint32_t b2s, double a1,
                                                   double b1, double a2, double b2) {
        sts_f64_off<0>(sW, a1);
        sts_f64_off<256>(sW, b1);
        sts_f64_off<512>(sW, a2);
        sts_f64_off<768>(sW, b2);
        __syncwarp();
        const double s1 = dot5s<0, TA1, TB1>(a1s, b1s);
        const double s2 = dot5s<1, TA2, TB2>(a2s, b2s);
        __syncwarp();
        const bool in = lane_id() < 25;
        return make_double2(in ? s1 : 0.0, in ? s2 : 0.0);
    }
    template <bool TA1, bool TB1, bool TA2, bool TB2>
    static __device__ __forceinline__ void mul2(const Ctx& c, LM a1, LM b1, LM a2, LM b2, LM& o1, LM& o2) {
        if (c.d != 5) {
            o1 = mul_impl(c.sc, c.d, a1, b1, TA1, TB1);
            o2 = mul_impl(c.sc, c.d, a2, b2, TA2, TB2);
            return;
        }
        const double2 r = mul2_d5<TA1, TB1, TA2, TB2>(c.sW, TA1 ? c.sAt : c.sAn, TB1 ? c.sBt : c.sBn, TA2 ? c.sAt : c.sAn,
                                                      TB2 ? c.sBt : c.sBn, a1, b1, a2, b2);
        o1 = r.x;
        o2 = r.y;
    }
    // op(A1) M1 op(C1) and op(A2) M2 op(C2): two chained product pairs (congruences L . L', L' . L, P . P, ...)
    template <bool TA1, bool TC1, bool TA2, bool TC2>
    static __device__ __forceinline__ void chain2(const Ctx& c, LM a1, LM m1, LM c1, LM a2, LM m2, LM c2, LM& o1, LM& o2) {
        LM u, v;
        mul2<TA1, false, TA2, false>(c, a1, m1, a2, m2, u, v);
        mul2<false, TC1, false, TC2>(c, u, c1, v, c2, o1, o2);
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
    // every entry gathers its constraints' weights in constraint order: the same sums as the serial form (out of line: ~30
    // call sites)
    static __device__ __noinline__ double scatter_gather(double wc, double base, int s0, int s1, int s2, int s3, int dd) {
        double r = base;
        const double v0 = __shfl_sync(kFull, wc, s0 < 0 ? 0 : s0);
        const double v1 = __shfl_sync(kFull, wc, s1 < 0 ? 0 : s1);
        const double v2 = __shfl_sync(kFull, wc, s2 < 0 ? 0 : s2);
        const double v3 = __shfl_sync(kFull, wc, s3 < 0 ? 0 : s3);
        if (s0 >= 0) r = r + v0;
        if (s1 >= 0) r = r + v1;
        if (s2 >= 0) r = r + v2;
        if (s3 >= 0) r = r + v3;
        return (lane_id() < dd) ? r : 0.0;
    }
    static __device__ __forceinline__ LM scatter(const Ctx& c, const CVec& w, const CVec& coef, LM base) {
        const double wc = w.v[0] * coef.v[0];
        if (c.nsrc < 0) return scatter_impl(c.sc, c.m, c.dd, c.rc, wc, base);
        return scatter_gather(wc, base, c.src[0], c.src[1], c.src[2], c.src[3], c.dd);
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
        chain2<false, false, false, false>(c, pt.x.v[1], symm(c, egR), pt.x.v[1], pt.x.v[2], symm(c, egQ), pt.x.v[2], r.v[1], r.v[2]);
        return r;
    }
    // pull(Phi) = [Phi Q', -Phi Q', (J-R)' Phi], then egrad2rgrad
    static __device__ __forceinline__ Vec rgrad_of_phi(const Ctx& c, const Pt& pt, LM Phi) {
        LM gJ, gQ;
        mul2<false, true, true, false>(c, Phi, pt.x.v[2], pt.JmR, Phi, gJ, gQ);
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
        pt.spd_ok = chol_pair(c, x.v[1], x.v[2], pt.LR, pt.LRi, pt.LQ, pt.LQi);
        mul2<true, false, true, false>(c, pt.LRi, pt.LRi, pt.LQi, pt.LQi, pt.Rinv, pt.Qinv);
    }

    // <a, b>_x = <aJ, bJ> + tr(R^-1 aR R^-1 bR) + tr(Q^-1 aQ Q^-1 bQ): per-lane partial
    // (taking this and the gradient conversions out of line as well costs registers -- 12 instead of 16 warps per SM --
    // and was slower: 442 vs 367 ms for the 2048-pair sweep)
    static __device__ __forceinline__ double inner_partial(const Ctx& c, const Pt& pt, const Vec& a, const Vec& b) {
        LM tR, tQ;
        chain2<false, false, false, false>(c, pt.Rinv, a.v[1], pt.Rinv, pt.Qinv, a.v[2], pt.Qinv, tR, tQ);
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
        LM p1, p2;
        mul2<false, false, false, false>(c, v.v[0] - v.v[1], pt.x.v[2], pt.JmR, v.v[2], p1, p2);
        return p1 + p2;
    }

    // 'is_euclidean_embedded' (RIPTRM.py:553-571): G*[v]_i = <-egrad g_i, v>_x with the EUCLIDEAN gradient inside the
    // manifold's inner product.  On the SPD components that metric is tr(P^-1 A P^-1 B), so with -egrad g_i = -coef_i pull(E_rc)
    // the entry of dA is taken with v_R -> R^-1 v_R R^-1 and v_Q -> Q^-1 v_Q Q^-1 (not the same operator as the Riemannian
    // form on this manifold -- the option is meant for submanifolds with the embedded metric -- but it is what the reference
    // evaluates when the key is set).
    static __device__ __forceinline__ LM dA_embedded(const Ctx& c, const Pt& pt, const Vec& v) {
        LM wR, wQ, p1, p2;
        chain2<false, false, false, false>(c, pt.Rinv, v.v[1], pt.Rinv, pt.Qinv, v.v[2], pt.Qinv, wR, wQ);
        mul2<false, false, false, false>(c, v.v[0] - wR, pt.x.v[2], pt.JmR, wQ, p1, p2);
        return p1 + p2;
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
        LM h1, h2, h3, h4, gJ, gQ;
        mul2<false, true, false, true>(c, dPhiL, pt.x.v[2], st.PhiL, v.v[2], h1, h2);
        mul2<true, false, true, false>(c, v.v[0] - v.v[1], st.PhiL, pt.JmR, dPhiL, h3, h4);
        const LM hJ = h1 + h2, hQ = h3 + h4;
        // Euclidean gradient of the Lagrangian
        mul2<false, true, true, false>(c, st.PhiL, pt.x.v[2], pt.JmR, st.PhiL, gJ, gQ);
        Vec hl;
        hl.v[0] = skew(c, hJ);
        {
            const LM R = pt.x.v[1], Q = pt.x.v[2];
            LM a1, a2, b1, b2;
            chain2<false, false, false, false>(c, R, symm(c, -hJ), R, Q, symm(c, hQ), Q, a1, a2);
            chain2<false, false, false, false>(c, v.v[1], symm(c, -gJ), R, v.v[2], symm(c, gQ), Q, b1, b2);
            hl.v[1] = a1 + symm(c, b1);
            hl.v[2] = a2 + symm(c, b2);
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
        LM gJ, gQ;
        mul2<false, true, true, false>(c, st.PhiL, pt.x.v[2], pt.JmR, st.PhiL, gJ, gQ);
        chain2<true, false, true, false>(c, w.LR, symm(c, -gJ), w.LR, w.LQ, symm(c, gQ), w.LQ, w.GR, w.GQ);
    }
    static __device__ __forceinline__ Vec whiten(const Ctx& c, const White& w, const Vec& v) {
        Vec r;
        r.v[0] = v.v[0];
        chain2<false, true, false, true>(c, w.LRi, v.v[1], w.LRi, w.LQi, v.v[2], w.LQi, r.v[1], r.v[2]);
        return r;
    }
    static __device__ __forceinline__ Vec unwhiten(const Ctx& c, const White& w, const Vec& v) {
        Vec r;
        r.v[0] = v.v[0];
        chain2<false, true, false, true>(c, w.LR, v.v[1], w.LR, w.LQ, v.v[2], w.LQ, r.v[1], r.v[2]);
        return r;
    }
    // whitened Hw: vh -> whiten(Hw[unwhiten(vh)])
    static __device__ __forceinline__ Vec Hw_white(const Ctx& c, const Pt& pt, const CVec& y, const Step& st, const White& w,
                                                   const Vec& vh) {
        LM vR, vQ, p1, p2;
        chain2<false, true, false, true>(c, w.LR, vh.v[1], w.LR, w.LQ, vh.v[2], w.LQ, vR, vQ);
        const LM D1 = vh.v[0] - vR;
        // dA = D1 Q + (J-R) vQ, and dA (X X') for dG_A = 2 h^2 dA (X X') / N: XXt sits in the product scratch as a lane matrix
        mul2<false, false, false, false>(c, D1, pt.x.v[2], pt.JmR, vQ, p1, p2);
        const LM dA = p1 + p2;
        const double dga = (2.0 * c.h * c.h) * mul(c, dA, c.XXt_lm) / (double)c.N;
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
        LM h1, h2, h3, h4, g1, g2, o1, o2;
        mul2<false, true, false, true>(c, dPhi, pt.x.v[2], st.PhiL, vQ, h1, h2);
        mul2<true, false, true, false>(c, D1, st.PhiL, pt.JmR, dPhi, h3, h4);
        const LM hJ = h1 + h2, hQ = h3 + h4;
        const LM hJt = transpose(c, hJ);
        mul2<false, false, false, false>(c, vh.v[1], w.GR, vh.v[2], w.GQ, g1, g2);
        chain2<true, false, true, false>(c, w.LR, -0.5 * (hJ + hJt), w.LR, w.LQ, symm(c, hQ), w.LQ, o1, o2);
        Vec out;
        out.v[0] = 0.5 * (hJ - hJt);
        out.v[1] = o1 + symm(c, g1);
        out.v[2] = o2 + symm(c, g2);
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
    // Cholesky factors and their inverses of BOTH SPD blocks at once: lanes 0..15 factor R, lanes 16..31 factor Q (every lane
    // of a half redundantly, in registers: d == 5 is unrolled), one reciprocal square root per pivot and no division; lanes 0
    // and 16 publish their half's result and every lane picks up its entries.  (Two calls of the general routine were 10 %
    // of the kernel's time: 30 dependent divisions / square roots each.)
    static __device__ __noinline__ bool chol_pair(const Ctx& c, LM R, LM Q, LM& LR, LM& LRi, LM& LQ, LM& LQi) {
        if (c.d != 5) {
            const bool a = chol_and_inverse(c.sc, c.d, R, LR, LRi);
            const bool b = chol_and_inverse(c.sc, c.d, Q, LQ, LQi);
            return a && b;
        }
        const int l = lane_id();
        const uint32_t base = c.sChol;
        sts_f64(base + 8 * l, R);              // [0, 32): R, [32, 64): Q
        sts_f64(base + 256 + 8 * l, Q);
        __syncwarp();
        const uint32_t mine = base + ((l >= 16) ? 256u : 0u);
        double P[5][5], L[5][5], Li[5][5], ri[5];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) P[i][j] = lds_f64(mine + 8 * (5 * i + j));
        bool ok = true;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            double s = P[j][j];
#pragma unroll
            for (int k = 0; k < j; ++k) s = fma(-L[j][k], L[j][k], s);
            ok = ok && (s > 0.0);
            ri[j] = rsqrt(s);
            L[j][j] = s * ri[j];
#pragma unroll
            for (int i = j + 1; i < 5; ++i) {
                double t = P[i][j];
#pragma unroll
                for (int k = 0; k < j; ++k) t = fma(-L[i][k], L[j][k], t);
                L[i][j] = t * ri[j];
            }
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) {          // L Li = I, column by column
            Li[j][j] = ri[j];
#pragma unroll
            for (int i = j + 1; i < 5; ++i) {
                double t = 0.0;
#pragma unroll
                for (int k = j; k < i; ++k) t = fma(L[i][k], Li[k][j], t);
                Li[i][j] = -t * ri[i];
            }
        }
        __syncwarp();
        if ((l & 15) == 0) {                    // [64, 96): L, [96, 128): Li of R ; Q 32 doubles further on -- 2 x 64 = 128
            const uint32_t o = base + 512 + ((l >= 16) ? 512u : 0u);
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    sts_f64(o + 8 * (5 * i + j), (j <= i) ? L[i][j] : 0.0);
                    sts_f64(o + 256 + 8 * (5 * i + j), (j <= i) ? Li[i][j] : 0.0);
                }
        }
        __syncwarp();
        const bool in = l < 25;
        const uint32_t e = base + 512 + 8 * (in ? l : 0);
        LR = in ? lds_f64(e) : 0.0;
        LRi = in ? lds_f64(e + 256) : 0.0;
        LQ = in ? lds_f64(e + 512) : 0.0;
        LQi = in ? lds_f64(e + 768) : 0.0;
        const bool okR = __shfl_sync(kFull, ok ? 1 : 0, 0) != 0, okQ = __shfl_sync(kFull, ok ? 1 : 0, 16) != 0;
        __syncwarp();
        return okR && okQ;
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
        chain2<false, true, false, true>(c, cc.LR, KR, cc.LR, cc.LQ, KQ, cc.LQ, v.v[1], v.v[2]);
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
        LM kR, kQ;
        chain2<false, true, false, true>(c, cc.LRi, v.v[1], cc.LRi, cc.LQi, v.v[2], cc.LQi, kR, kQ);
        pack(c, v.v[0], out, true);
        pack(c, kR, out + ns, false);
        pack(c, kQ, out + ns + np, false);
    }

    static __device__ __forceinline__ Vec retract(const Ctx& c, const Pt& pt, const Vec& dx) {
        Vec r;
        r.v[0] = pt.x.v[0] + dx.v[0];
        LM qR, qQ;
        chain2<false, false, false, false>(c, dx.v[1], pt.Rinv, dx.v[1], dx.v[2], pt.Qinv, dx.v[2], qR, qQ);
        r.v[1] = symm(c, (pt.x.v[1] + dx.v[1]) + qR / 2.0);
        r.v[2] = symm(c, (pt.x.v[2] + dx.v[2]) + qQ / 2.0);
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
    // (out of line, scalars by value: only evaluated when a trace row is written)
    static __device__ __noinline__ double spd_dist_impl(double* sc, int d, int tlane, LM P, LM B) {
        const int l = lane_id();
        double* s2 = sc + 64;
        s2[l] = P;
        __syncwarp();
        double w[DMAX], V[DMAX][DMAX];
        sm::jacobi_eig<DMAX>(s2, d, w, V);
        __syncwarp();
        double ih = 0.0;
        if (l < d * d) {
            const int i = l / d, j = l - i * d;
            for (int k = 0; k < d; ++k) ih = fma(V[i][k] * (1.0 / sqrt(w[k])), V[j][k], ih);
        }
        const LM t = mul_impl(sc, d, mul_impl(sc, d, ih, B, false, false), ih, false, false);
        const LM M = 0.5 * (t + __shfl_sync(kFull, t, tlane));
        s2[l] = M;
        __syncwarp();
        sm::jacobi_eig<DMAX>(s2, d, w, V);
        __syncwarp();
        double acc = 0.0;
        for (int k = 0; k < d; ++k) {
            const double lg = log(w[k]);
            acc += lg * lg;
        }
        return sqrt(acc);
    }
    static __device__ __forceinline__ double spd_dist(const Ctx& c, LM P, LM B) { return spd_dist_impl(c.sc, c.d, c.tlane, P, B); }
    static __device__ __forceinline__ double dist(const Ctx& c, const Vec& xPrev, const Pt& pt) {
        const double dj = xPrev.v[0] - pt.x.v[0];
        const double nJ = sqrt(wsum(dj * dj));
        const double nR = spd_dist(c, xPrev.v[1], pt.x.v[1]);
        const double nQ = spd_dist(c, xPrev.v[2], pt.x.v[2]);
        return sqrt((nJ * nJ + nR * nR) + nQ * nQ);
    }
};

}  // namespace riptrm

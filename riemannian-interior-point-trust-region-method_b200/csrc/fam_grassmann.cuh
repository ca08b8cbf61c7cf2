// fam_grassmann.cuh -- the reference's "Rosenbrock" workload: a quadratic chain cost on Grassmann(n, p),
// vec(X)_i >= -offset  (src/Rosenbrock/coordinator.py:33-91; config_simulation.yaml:10-12: n=5, k=3, alpha=1e7).
//
//   f(X)  = sum_{i<np-1} alpha (v_{i+1}-v_i)^2 + (1-v_i)^2,  v = X.flatten() (row-major)          (:44-51)
//   egrad = T v + b, ehess[V] = T vec(V)   (T tridiagonal, closed form SURVEY.md App. A.2)
//   g_i   = -v_i - offset  => s = v + offset, egrad g_i = -E_i, ehess g_i = 0                     (:58-63)
//   P_X U = U - X (X'U);  rhess = P_X(ehess) - V (X' egrad)          (pymanopt Grassmann; oracle/manifolds.py)
//   Hess L[V] = P_X(T[V]) + V (X'Y - X' egrad f),  Y = y reshaped n x p          (RIPTRM.py:491-523)
//   G_X(w) = P_X(W), G*_X[V]_i = <P_X E_i, V> = (P_X V)_i                           (:525-571)
//   retraction: polar factor of X + V = (X+V) ((X+V)'(X+V))^{-1/2}   (pymanopt: u @ vt of the thin SVD)
//
// One warp per instance, entry e = i*p + j on lane e (n*p <= 32), matrix products through the warp's
// shared-memory scratch (smallmat.cuh).
#pragma once
#include "smallmat.cuh"
#include "solver_warp.cuh"

namespace riptrm {

struct GrassmannFam {
    static constexpr bool kTcgReturnsHw = true;   // tcg_generic accumulates Hw[eta] beside eta: solver_warp.cuh inner_step
    static constexpr int K = 1;
    static constexpr int MK = 1;
    static constexpr int PMAX = 5;
    using Vec = WVec<1>;
    using CVec = WVec<1>;

    struct Ctx {
        int n, p, np;
        double alpha, offset;
        bool embedded;
        double* sc;  // shared-memory scratch: 6 slots of 32 doubles
        double* perp;  // Exact_RepMat: n x (n-p) orthonormal complement of the current point (after the scratch slots)
    };
    struct Pt {
        Vec x;
        CVec s;
        double cost;
        Vec eg;    // Euclidean gradient of f at x
        Vec XtG;   // X' egrad f (p x p, entry a*p+b on lane a*p+b)
    };
    struct Step {
        Vec c;
        CVec ys;
        Vec M;  // X'Y - X'egrad f  (p x p)
    };

    static constexpr int kScratchDoubles = 6 * 32;
    static constexpr int kPerpDoubles = 128;   // n (n - p) <= 128 for the exact path (the reference's 5 x 2 = 10)
    static __host__ __device__ constexpr int smem_doubles(int, int) { return kScratchDoubles + kPerpDoubles; }
    static constexpr int kComponents = 1;
    template <class Params>
    static __device__ __forceinline__ Ctx make_ctx(const Params& P, const DevOpts& o, double* smem) {
        Ctx c;
        c.n = P.n;
        c.p = P.p;
        c.np = P.n * P.p;
        c.alpha = P.alpha;
        c.offset = P.offset;
        c.embedded = o.is_euclidean_embedded != 0;
        c.sc = smem;
        c.perp = smem + kScratchDoubles;
        return c;
    }
    static __device__ __forceinline__ Vec load_x(const Ctx& c, const double* g) {
        Vec r;
        r.v[0] = (lane_id() < c.np) ? g[lane_id()] : 0.0;
        return r;
    }
    static __device__ __forceinline__ CVec load_y(const Ctx& c, const double* g) { return load_x(c, g); }
    static __device__ __forceinline__ void store_x(const Ctx& c, double* g, const Vec& v) {
        if (lane_id() < c.np) g[lane_id()] = v.v[0];
    }
    static __device__ __forceinline__ void store_y(const Ctx& c, double* g, const CVec& v) { store_x(c, g, v); }
    static __device__ __forceinline__ double* slot(const Ctx& c, int i) { return c.sc + 32 * i; }
    static __device__ __forceinline__ bool active(const Ctx& c, int) { return lane_id() < c.np; }
    static __device__ __forceinline__ bool cactive(const Ctx& c, int k) { return active(c, k); }
    static __device__ __forceinline__ int dim(const Ctx& c) { return c.np - c.p * c.p; }
    static __device__ __forceinline__ int num_constraints(const Ctx& c) { return c.np; }
    static __device__ __forceinline__ double typical_dist(const Ctx& c) { return sqrt((double)c.p); }
    static __device__ __forceinline__ bool domain_ok(const Ctx&, const Pt&) { return true; }

    static __device__ __forceinline__ void put(const Ctx& c, int s, const Vec& v) {
        slot(c, s)[lane_id()] = v.v[0];
        __syncwarp();
    }
    static __device__ __forceinline__ Vec get(const Ctx& c, int s, int len) {
        Vec r;
        r.v[0] = (lane_id() < len) ? slot(c, s)[lane_id()] : 0.0;
        __syncwarp();
        return r;
    }

    // T applied to a flat vector held one entry per lane (+ b for the gradient): the order of the NumPy
    // restatement (oracle/problems.py RosenbrockProblem._apply_T)
    static __device__ __forceinline__ Vec apply_T(const Ctx& c, const Vec& v, bool with_b) {
        const int l = lane_id(), L = c.np;
        put(c, 5, v);
        const double* a = slot(c, 5);
        double out = 0.0;
        if (l < L) {
            if (l < L - 1) out = out + (-2.0 * c.alpha * (a[l + 1] - a[l]) + 2.0 * a[l]);
            if (l > 0) out = out + 2.0 * c.alpha * (a[l] - a[l - 1]);
            if (with_b && l < L - 1) out = out + (-2.0);
        }
        __syncwarp();
        Vec r;
        r.v[0] = out;
        return r;
    }

    // X'U (p x p) for n x p matrices held one entry per lane
    static __device__ __forceinline__ Vec XtU(const Ctx& c, const Vec& X, const Vec& U) {
        put(c, 0, X);
        put(c, 1, U);
        sm::mm(slot(c, 2), slot(c, 0), slot(c, 1), c.p, c.n, c.p, true, false);
        return get(c, 2, c.p * c.p);
    }
    // V M for V n x p, M p x p
    static __device__ __forceinline__ Vec VM(const Ctx& c, const Vec& V, const Vec& M) {
        put(c, 0, V);
        put(c, 1, M);
        sm::mm(slot(c, 2), slot(c, 0), slot(c, 1), c.n, c.p, c.p);
        return get(c, 2, c.np);
    }

    static __device__ __forceinline__ Vec project(const Ctx& c, const Pt& pt, const Vec& v) {
        const Vec XtV = XtU(c, pt.x, v);
        const Vec XXtV = VM(c, pt.x, XtV);
        Vec r;
        r.v[0] = v.v[0] - XXtV.v[0];
        return r;
    }

    static __device__ __forceinline__ void eval_point(const Ctx& c, const Vec& x, Pt& pt) {
        pt.x = x;
        // cost: the reference's sequential accumulation over i (coordinator.py:48-50)
        put(c, 5, x);
        const double* v = slot(c, 5);
        double val = 0.0;
        for (int i = 0; i < c.np - 1; ++i) {
            const double d = v[i + 1] - v[i], e = 1.0 - v[i];
            val = (val + c.alpha * (d * d)) + e * e;
        }
        __syncwarp();
        pt.cost = val;
        pt.eg = apply_T(c, x, true);
        pt.XtG = XtU(c, x, pt.eg);
        pt.s.v[0] = active(c, 0) ? (x.v[0] + c.offset) : 0.0;
    }

    static __device__ __forceinline__ double inner_partial(const Ctx&, const Pt&, const Vec& a, const Vec& b) {
        return a.v[0] * b.v[0];
    }
    static __device__ __forceinline__ double inner(const Ctx& c, const Pt& pt, const Vec& a, const Vec& b) {
        return wsum(inner_partial(c, pt, a, b));
    }

    static __device__ __forceinline__ void begin_step(const Ctx& c, const Pt& pt, const CVec& y, double mu, Step& st) {
        Vec w;
        const bool on = active(c, 0);
        w.v[0] = on ? mu * (1.0 / pt.s.v[0]) : 0.0;
        st.ys.v[0] = on ? y.v[0] / pt.s.v[0] : 0.0;
        const Vec gradf = project(c, pt, pt.eg);
        const Vec Gw = project(c, pt, w);
        st.c.v[0] = gradf.v[0] - Gw.v[0];
        const Vec XtY = XtU(c, pt.x, y);
        st.M.v[0] = XtY.v[0] - pt.XtG.v[0];
    }

    static __device__ __forceinline__ CVec gadj(const Ctx& c, const Pt& pt, const Vec& v) {
        if (c.embedded) return v;
        return project(c, pt, v);
    }

    static __device__ __forceinline__ Vec Hw(const Ctx& c, const Pt& pt, const CVec&, const Step& st, const Vec& v) {
        const Vec Tv = apply_T(c, v, false);
        const Vec PTv = project(c, pt, Tv);
        const Vec vM = VM(c, v, st.M);
        const CVec ga = gadj(c, pt, v);
        Vec w;
        w.v[0] = st.ys.v[0] * ga.v[0];
        const Vec Gw = project(c, pt, w);
        Vec out;
        out.v[0] = (PTv.v[0] + vM.v[0]) + Gw.v[0];
        return out;
    }

    static __device__ __forceinline__ TcgResult tcg(const Ctx& ctx, const DevOpts& o, const Pt& pt, const CVec& y,
                                                    const Step& st, double Delta, Vec& eta, Vec& Heta) {
        return tcg_generic<GrassmannFam>(ctx, o, pt, y, st, Delta, eta, Heta);
    }

    // ---- Exact_RepMat: orthonormal tangent basis X_perp E_ab (riptrm_b200/basis.py grassmann_basis) -------------------------
    // X_perp = the last n - p columns of the complete Householder QR factor of X: T_X = { X_perp K }, and K -> X_perp K is an
    // isometry from R^{(n-p) x p} (Frobenius) since X_perp has orthonormal columns.  n x p is tiny: lane 0 factorises.
    struct Coord {};
    static __device__ __noinline__ void coord_setup(const Ctx& c, const Pt& pt, Coord&) {
        put(c, 0, pt.x);
        const int n = c.n, p = c.p, q = c.n - c.p;
        if (lane_id() == 0) {
            double* R = slot(c, 1);      // working copy of X (n x p)
            double* Vh = slot(c, 2);     // Householder vectors, column j in Vh[:, j]
            double* E = c.perp;          // n x q
            for (int e = 0; e < n * p; ++e) {
                R[e] = slot(c, 0)[e];
                Vh[e] = 0.0;
            }
            for (int j = 0; j < p; ++j) {
                double nrm2 = 0.0;
                for (int i = j; i < n; ++i) nrm2 = fma(R[i * p + j], R[i * p + j], nrm2);
                const double x0 = R[j * p + j];
                const double alpha = (x0 >= 0.0) ? -sqrt(nrm2) : sqrt(nrm2);
                for (int i = j; i < n; ++i) Vh[i * p + j] = R[i * p + j];
                Vh[j * p + j] = x0 - alpha;
                double vv = 0.0;
                for (int i = j; i < n; ++i) vv = fma(Vh[i * p + j], Vh[i * p + j], vv);
                if (vv > 0.0) {
                    for (int cidx = j; cidx < p; ++cidx) {
                        double dot = 0.0;
                        for (int i = j; i < n; ++i) dot = fma(Vh[i * p + j], R[i * p + cidx], dot);
                        const double f = 2.0 * dot / vv;
                        for (int i = j; i < n; ++i) R[i * p + cidx] = R[i * p + cidx] - f * Vh[i * p + j];
                    }
                }
            }
            for (int i = 0; i < n; ++i)
                for (int b = 0; b < q; ++b) E[i * q + b] = (i == p + b) ? 1.0 : 0.0;
            for (int j = p - 1; j >= 0; --j) {   // Q[:, p:] = H_0 ... H_{p-1} [e_p ... e_{n-1}]
                double vv = 0.0;
                for (int i = j; i < n; ++i) vv = fma(Vh[i * p + j], Vh[i * p + j], vv);
                if (!(vv > 0.0)) continue;
                for (int b = 0; b < q; ++b) {
                    double dot = 0.0;
                    for (int i = j; i < n; ++i) dot = fma(Vh[i * p + j], E[i * q + b], dot);
                    const double f = 2.0 * dot / vv;
                    for (int i = j; i < n; ++i) E[i * q + b] = E[i * q + b] - f * Vh[i * p + j];
                }
            }
        }
        __syncwarp();
    }
    static __device__ __forceinline__ Vec from_coords(const Ctx& c, const Pt&, const Coord&, const double* coef) {
        sm::mm(slot(c, 2), c.perp, coef, c.n, c.n - c.p, c.p);          // X_perp K
        return get(c, 2, c.np);
    }
    static __device__ __forceinline__ void to_coords(const Ctx& c, const Pt&, const Coord&, const Vec& v, double* out) {
        put(c, 0, v);
        sm::mm(out, c.perp, slot(c, 0), c.n - c.p, c.n, c.p, true, false);   // X_perp' V
    }

    // (A'A)^{-1/2} applied to A = X + V: the polar factor (pymanopt: u @ vt of the thin SVD)
    static __device__ __forceinline__ Vec retract(const Ctx& c, const Pt& pt, const Vec& dx) {
        Vec A;
        A.v[0] = pt.x.v[0] + dx.v[0];
        const Vec AtA = XtU(c, A, A);
        put(c, 3, AtA);
        double w[PMAX], V[PMAX][PMAX];
        sm::jacobi_eig<PMAX>(slot(c, 3), c.p, w, V);
        __syncwarp();
        const int l = lane_id();
        if (l < c.p * c.p) {
            const int a = l / c.p, b = l - a * c.p;
            double s = 0.0;
            for (int k = 0; k < c.p; ++k) s = fma(V[a][k] * (1.0 / sqrt(w[k])), V[b][k], s);
            slot(c, 4)[l] = s;
        }
        __syncwarp();
        const Vec Minv = get(c, 4, c.p * c.p);
        return VM(c, A, Minv);
    }

    static __device__ __forceinline__ double gradL_xy_partial(const Ctx&, const Pt&, const CVec&) { return 0.0; }
    static __device__ __forceinline__ double gradL_norm_given(const Ctx& c, const Pt& pt, const CVec& y, double) {
        // grad f + sum_i y_i grad g_i = P_X(egrad f) - P_X(Y)
        const Vec a = project(c, pt, pt.eg), b = project(c, pt, y);
        Vec g;
        g.v[0] = a.v[0] - b.v[0];
        return sqrt(wsum(g.v[0] * g.v[0]));
    }
    static __device__ __forceinline__ double gradL_norm(const Ctx& c, const Pt& pt, const CVec& y) {
        return gradL_norm_given(c, pt, y, 0.0);
    }

    // src/Rosenbrock/simulator.py:107-114: 0 when rank(X) == p, else inf
    static __device__ __forceinline__ double manvio(const Ctx& c, const Pt& pt) {
        const Vec XtX = XtU(c, pt.x, pt.x);
        put(c, 3, XtX);
        double w[PMAX], V[PMAX][PMAX];
        sm::jacobi_eig<PMAX>(slot(c, 3), c.p, w, V);
        __syncwarp();
        double wmin_ = w[0], wmax_ = w[0];
        for (int k = 1; k < c.p; ++k) {
            wmin_ = fmin(wmin_, w[k]);
            wmax_ = fmax(wmax_, w[k]);
        }
        // numpy.linalg.matrix_rank tolerance: sigma_max * max(n, p) * eps on singular values (= sqrt of these)
        const double tol = sqrt(wmax_) * (double)c.n * 2.220446049250313e-16;
        return (wmin_ > 0.0 && sqrt(wmin_) > tol) ? 0.0 : CUDART_INF;
    }

    // pymanopt Grassmann.dist: || arccos(min(svd(X'Y), 1)) ||
    static __device__ __forceinline__ double dist(const Ctx& c, const Vec& xPrev, const Pt& pt) {
        const Vec XtY = XtU(c, xPrev, pt.x);
        const Vec G = XtU_pp(c, XtY);
        put(c, 3, G);
        double w[PMAX], V[PMAX][PMAX];
        sm::jacobi_eig<PMAX>(slot(c, 3), c.p, w, V);
        __syncwarp();
        double acc = 0.0;
        for (int k = 0; k < c.p; ++k) {
            const double sv = fmin(sqrt(fmax(w[k], 0.0)), 1.0);
            const double a = acos(sv);
            acc += a * a;
        }
        return sqrt(acc);
    }
    // M'M for a p x p matrix held on lanes
    static __device__ __forceinline__ Vec XtU_pp(const Ctx& c, const Vec& M) {
        put(c, 0, M);
        sm::mm(slot(c, 2), slot(c, 0), slot(c, 0), c.p, c.p, c.p, true, false);
        return get(c, 2, c.p * c.p);
    }
};

}  // namespace riptrm

/*
 * riptrm_det.c -- TEST / BENCH INFRASTRUCTURE (CPU oracle, not product code).
 *
 * Deterministic plain-C restatement of the reference's RIPTRM tCG path for the NonnegPCA / Sphere workload
 * with closed-form derivatives:
 *     run / outer_step / inner_run   /root/reference/src/solver/RIPTRM.py:909-976, :866-896, :785-847
 *     inner_step                     :707-783
 *     truncated_conjugate_gradient   :41-216
 *     compute_inner_stoppingcriteria :574-629
 *     update_xy_TR_radius            :631-705
 *     evaluation / compute_residual  /root/reference/src/solver/utils.py:342-368, :269-340, :237-267
 *     problem                        /root/reference/src/NonnegPCA/coordinator.py:37-95
 *
 * What "deterministic" means here: every floating-point operation, its order and its fused-multiply-add
 * policy are specified (DESIGN.md "Arithmetic specification"): a vector of n <= 128 doubles is laid out on
 * 32 lanes in pairs, lane l owning the elements e(l,k) = 64*(k>>1) + 2*l + (k&1), k < K (K = 2 for n <= 64,
 * else 4); a dot product is a per-lane fma chain over k followed by a halving tree
 * over lanes (what an xor-butterfly of warp shuffles computes); S.v accumulates even and odd rows in two
 * chains per element; elementwise expressions are evaluated without contraction (compile with
 * -ffp-contract=off).  The CUDA kernels implement the same specification, which is what allows bit-for-bit
 * comparison of whole iteration traces (parity tier T1).  Tier T2 compares this file with the NumPy oracle
 * (oracle/riptrm_oracle.py), itself bit-identical to the unmodified reference on the reference's dataset.
 *
 * PARITY PIN: tests/test_oracle_c.py checks this file against tests/golden/nonnegpca_1_a_K40.json (output of
 * the unmodified reference) and against the NumPy oracle on generated instances.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LANES 32
#define MAXK 4
#define MAXN (LANES * MAXK)
#define ELEM(l, k) (64 * ((k) >> 1) + 2 * (l) + ((k) & 1))

#define TRACE_FIELDS 25
#define SUMMARY_FIELDS 16

enum { TCG_MAX_INNER_ITER = 0, TCG_NEGATIVE_CURVATURE, TCG_EXCEEDED_TR, TCG_MODEL_INCREASED, TCG_REACHED_TARGET_LINEAR,
       TCG_REACHED_TARGET_SUPERLINEAR };
enum { INNER_NONE = 0, INNER_CONVERGED, INNER_PRIMAL_INFEASIBLE, INNER_SUCCESSFUL, INNER_UNSUCCESSFUL, INNER_MAX_TIME,
       INNER_MAX_ITER };
enum { RADIUS_NONE = 0, RADIUS_REDUCED, RADIUS_EXPANDED, RADIUS_UNCHANGED };
enum { STOP_RUNNING = 0, STOP_MAXTIME, STOP_MAXITER, STOP_TOLRESID, STOP_NUMERICAL };

typedef struct {
    int32_t maxiter, inner_maxiter, tcg_mininner, tcg_maxinner, is_euclidean_embedded, trace_mode, trace_capacity,
        reserved0;
    double tolresid, maxtime, inner_maxtime, initial_tr_radius, minimal_initial_tr_radius, maximal_tr_radius, rho,
        reduction_regularization, gamma, const_left, const_right, tcg_theta, tcg_kappa;
    const double* mu_sched;
    const double* tol_lagrangian_sched;
    const double* tol_complementarity_sched;
} det_options; /* same layout as riptrm_options of include/riptrm_b200.h */

typedef struct {
    int n, K;
    double eps;
    int embedded;
    double S[MAXN * MAXN]; /* row-major n x n, S = Z + Z' */
} Ctx;

typedef struct {
    double x[MAXN], s[MAXN], Sx[MAXN];
    double cost, xSx;
} Pt;

typedef struct {
    double c[MAXN], ys[MAXN];
    double kappa;
} Step;

/* ---- lane-structured reductions ------------------------------------------------------------------- */
static double lane_tree(double* p) {
    for (int off = 16; off > 0; off >>= 1)
        for (int l = 0; l < off; ++l) p[l] = p[l] + p[l + off];
    return p[0];
}
static double vdot(const Ctx* c, const double* a, const double* b) {
    double p[LANES];
    for (int l = 0; l < LANES; ++l) {
        double q = a[ELEM(l, 0)] * b[ELEM(l, 0)];
        for (int k = 1; k < c->K; ++k) q = fma(a[ELEM(l, k)], b[ELEM(l, k)], q);
        p[l] = q;
    }
    return lane_tree(p);
}

/* out = S v: two fma chains per element (even rows, odd rows), summed at the end */
static void matvec(const Ctx* c, const double* v, double* out) {
    const int n = c->n;
    for (int e = 0; e < c->K * LANES; ++e) {
        if (e >= n) {
            out[e] = 0.0;
            continue;
        }
        double a0 = 0.0, a1 = 0.0;
        int j = 0;
        for (; j + 1 < n; j += 2) {
            a0 = fma(c->S[j * n + e], v[j], a0);
            a1 = fma(c->S[(j + 1) * n + e], v[j + 1], a1);
        }
        if (j < n) a0 = fma(c->S[j * n + e], v[j], a0);
        out[e] = a0 + a1;
    }
}

/* same operation sequence as the CUDA det_log (common.cuh): fdlibm-style log without contraction */
static double det_log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    union {
        double d;
        uint64_t u;
    } w;
    w.d = x;
    int64_t k = 0;
    if (x != x) return x;
    if (x < 0.0) {
        w.u = 0x7ff8000000000000ull;
        return w.d;
    }
    if (x == 0.0) {
        w.u = 0xfff0000000000000ull;
        return w.d;
    }
    if ((w.u >> 52) == 0x7ff) return x;
    if ((w.u >> 52) == 0) {
        w.d = x * 18014398509481984.0;
        k -= 54;
    }
    uint64_t hx = w.u >> 32;
    k += (int64_t)(hx >> 20) - 1023;
    hx &= 0x000fffff;
    uint64_t i = (hx + 0x95f64) & 0x100000;
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

/* ---- problem pieces (fam_sphere.cuh) --------------------------------------------------------------- */
static void eval_point(const Ctx* c, const double* x, Pt* pt) {
    const int N = c->K * LANES;
    memcpy(pt->x, x, sizeof(double) * N);
    matvec(c, x, pt->Sx);
    pt->xSx = vdot(c, x, pt->Sx);
    pt->cost = -0.5 * pt->xSx;
    for (int e = 0; e < N; ++e) pt->s[e] = (e < c->n) ? (x[e] + c->eps) : 0.0;
}

static void begin_step(const Ctx* c, const Pt* pt, const double* y, double mu, Step* st) {
    const int N = c->K * LANES;
    double w[MAXN];
    for (int e = 0; e < N; ++e) {
        const int on = e < c->n;
        w[e] = on ? mu * (1.0 / pt->s[e]) : 0.0;
        st->ys[e] = on ? y[e] / pt->s[e] : 0.0;
    }
    const double xw = vdot(c, pt->x, w), yx = vdot(c, y, pt->x);
    st->kappa = pt->xSx + yx;
    for (int e = 0; e < N; ++e) {
        const double gradf = -pt->Sx[e] + pt->xSx * pt->x[e];
        const double Gw = w[e] - xw * pt->x[e];
        st->c[e] = gradf - Gw;
    }
}

static void gadj(const Ctx* c, const Pt* pt, const double* v, double* g) {
    const int N = c->K * LANES;
    if (c->embedded) {
        memcpy(g, v, sizeof(double) * N);
        return;
    }
    const double b = vdot(c, pt->x, v);
    for (int e = 0; e < N; ++e) g[e] = v[e] - pt->x[e] * b;
}

static void Hw(const Ctx* c, const Pt* pt, const Step* st, const double* v, double* out) {
    const int N = c->K * LANES;
    double Sv[MAXN], t[MAXN];
    matvec(c, v, Sv);
    const double a = vdot(c, pt->x, Sv), b = vdot(c, pt->x, v);
    for (int e = 0; e < N; ++e) {
        const double ga = c->embedded ? v[e] : (v[e] - pt->x[e] * b);
        t[e] = st->ys[e] * ga;
    }
    const double d = vdot(c, pt->x, t);
    for (int e = 0; e < N; ++e) {
        const double hl = (-Sv[e] + a * pt->x[e]) + st->kappa * v[e];
        const double g = t[e] - d * pt->x[e];
        out[e] = hl + g;
    }
}

static void project(const Ctx* c, const Pt* pt, double* v) {
    const int N = c->K * LANES;
    const double a = vdot(c, pt->x, v);
    for (int e = 0; e < N; ++e) v[e] = v[e] - a * pt->x[e];
}

static void retract(const Ctx* c, const Pt* pt, const double* dx, double* out) {
    const int N = c->K * LANES;
    for (int e = 0; e < N; ++e) out[e] = pt->x[e] + dx[e];
    const double nrm = sqrt(vdot(c, out, out));
    for (int e = 0; e < N; ++e) out[e] = out[e] / nrm;
}

static double gradL_norm(const Ctx* c, const Pt* pt, const double* y) {
    const int N = c->K * LANES;
    double g[MAXN];
    const double xy = vdot(c, pt->x, y);
    for (int e = 0; e < N; ++e) g[e] = (-pt->Sx[e] + pt->xSx * pt->x[e]) - (y[e] - xy * pt->x[e]);
    return sqrt(vdot(c, g, g));
}

/* ---- Steihaug-Toint tCG (RIPTRM.py:41-216; solver_warp.cuh tcg<F>) ---------------------------------- */
typedef struct {
    int iters, stop;
    double model_value;
} TcgResult;

static TcgResult tcg(const Ctx* c, const det_options* o, const Pt* pt, const Step* st, double Delta, double* eta,
                     double* Heta) {
    const int N = c->K * LANES;
    double r[MAXN], delta[MAXN], Hd[MAXN], new_eta[MAXN], new_Heta[MAXN];
    memset(eta, 0, sizeof(double) * N);
    memset(Heta, 0, sizeof(double) * N);
    memcpy(r, st->c, sizeof(double) * N);
    double e_Pe = 0.0;
    double r_r = vdot(c, r, r);
    const double norm_r0 = sqrt(r_r);
    double z_r = r_r, d_Pd = z_r;
    for (int e = 0; e < N; ++e) delta[e] = -r[e];
    double e_Pd = 0.0, model_value = 0.0;
    TcgResult res;
    res.stop = TCG_MAX_INNER_ITER;
    const int maxinner = o->tcg_maxinner < 0 ? (c->n - 1) : o->tcg_maxinner;
    const double Delta2 = Delta * Delta;
    const double nr_theta = (o->tcg_theta == 1.0) ? norm_r0 : pow(norm_r0, o->tcg_theta);
    const double target = norm_r0 * fmin(nr_theta, o->tcg_kappa);
    int j = 0;
#ifndef FAITHFUL_TCG
    /* Merged-reduction form (what the CUDA Sphere kernel runs): algebraically the loop of RIPTRM.py:98-214, with
     * <x,t>, <delta,H delta> and the projection coefficient assembled from 6 + 4 inner products taken in two
     * reduction rounds per iteration instead of five dependent ones.  -DFAITHFUL_TCG compiles the loop in the
     * reference's operation order instead (same results to rounding; tests/test_oracle_c.py compares both). */
    double wv[MAXN], Sv[MAXN], tmp[MAXN], r_new[MAXN];
    const double target_sq = target * target;
    double inv_zr = 1.0 / z_r;
    for (int e = 0; e < N; ++e) wv[e] = pt->x[e] * st->ys[e];
    const double q = vdot(c, wv, pt->x);
#ifdef RIPTRM_TCG_FMA
    /* EXPERIMENT (round 2, not adopted): the same loop with the axpy / update lines written as fused multiply-adds, which
     * would remove ~29 of the ~225 FP64 instructions of a tCG iteration on the GPU.  Measured against the reference's golden
     * run (scripts/parity_report.py --engine c with this switch): the discrete trace leaves at outer iteration 19 instead of
     * 20, the KKT residual agrees to 3.5e-8 / 2.6e-3 instead of 9.5e-9 / 9.5e-7 over outer iterations 1-9 / 10-19 -- the
     * parity windows SHRINK, so the kernels keep the unfused form. */
    for (; j < maxinner; ++j) {
        matvec(c, delta, Sv);
        const double a = vdot(c, pt->x, Sv), b = vdot(c, pt->x, delta), g1 = vdot(c, wv, delta);
        const double h1 = vdot(c, delta, Sv), h2 = vdot(c, delta, delta);
        for (int e = 0; e < N; ++e) tmp[e] = st->ys[e] * delta[e];
        const double h3 = vdot(c, delta, tmp);
        const double d = c->embedded ? g1 : fma(-b, q, g1);
        for (int e = 0; e < N; ++e) {
            const double ga = c->embedded ? delta[e] : fma(-pt->x[e], b, delta[e]);
            const double t = st->ys[e] * ga;
            const double hl = fma(st->kappa, delta[e], fma(a, pt->x[e], -Sv[e]));
            const double g = fma(-d, pt->x[e], t);
            Hd[e] = hl + g;
        }
        const double dt = c->embedded ? h3 : fma(-b, g1, h3);
        const double d_Hd = fma(-d, b, fma(st->kappa, h2, fma(a, b, -h1)) + dt);
        double alpha = 0.0, e_Pe_new = e_Pe;
        if (d_Hd != 0.0) {
            alpha = z_r / d_Hd;
            e_Pe_new = fma(alpha * alpha, d_Pd, fma(2.0 * alpha, e_Pd, e_Pe));
        }
        if (d_Hd <= 0.0 || e_Pe_new >= Delta2) {
            const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta2 - e_Pe))) / d_Pd;
            for (int e = 0; e < N; ++e) {
                eta[e] = fma(tau, delta[e], eta[e]);
                Heta[e] = fma(tau, Hd[e], Heta[e]);
            }
            res.stop = (d_Hd <= 0.0) ? TCG_NEGATIVE_CURVATURE : TCG_EXCEEDED_TR;
            ++j;
            break;
        }
        e_Pe = e_Pe_new;
        for (int e = 0; e < N; ++e) {
            new_eta[e] = fma(alpha, delta[e], eta[e]);
            new_Heta[e] = fma(alpha, Hd[e], Heta[e]);
            r_new[e] = fma(alpha, Hd[e], r[e]);
        }
        const double new_model = fma(0.5, vdot(c, new_eta, new_Heta), vdot(c, new_eta, st->c));
        const double rr_new = vdot(c, r_new, r_new), xr = vdot(c, pt->x, r_new);
        if (new_model >= model_value) {
            res.stop = TCG_MODEL_INCREASED;
            ++j;
            break;
        }
        memcpy(eta, new_eta, sizeof(double) * N);
        memcpy(Heta, new_Heta, sizeof(double) * N);
        memcpy(r, r_new, sizeof(double) * N);
        model_value = new_model;
        r_r = rr_new;
        /* residual test on squares (||r|| <= target <=> r_r <= target^2): no square root on the critical path */
        if (j >= o->tcg_mininner && r_r <= target_sq) {
            res.stop = (o->tcg_kappa < nr_theta) ? TCG_REACHED_TARGET_LINEAR : TCG_REACHED_TARGET_SUPERLINEAR;
            ++j;
            break;
        }
        /* beta = z_r / z_r_old as a product with the reciprocal formed one iteration earlier (off the critical path) */
        const double beta = r_r * inv_zr;
        z_r = r_r;
        inv_zr = 1.0 / z_r;
        const double xd = fma(beta, b, -xr); /* <x, -r + beta delta> */
        for (int e = 0; e < N; ++e) {
            const double dn = fma(beta, delta[e], -r[e]);
            delta[e] = fma(-xd, pt->x[e], dn);
        }
        e_Pd = beta * fma(alpha, d_Pd, e_Pd);
        d_Pd = fma(beta * beta, d_Pd, z_r);
    }
#else
    for (; j < maxinner; ++j) {
        matvec(c, delta, Sv);
        const double a = vdot(c, pt->x, Sv), b = vdot(c, pt->x, delta), g1 = vdot(c, wv, delta);
        const double h1 = vdot(c, delta, Sv), h2 = vdot(c, delta, delta);
        for (int e = 0; e < N; ++e) tmp[e] = st->ys[e] * delta[e];
        const double h3 = vdot(c, delta, tmp);
        const double d = c->embedded ? g1 : (g1 - b * q);
        for (int e = 0; e < N; ++e) {
            const double ga = c->embedded ? delta[e] : (delta[e] - pt->x[e] * b);
            const double t = st->ys[e] * ga;
            const double hl = (-Sv[e] + a * pt->x[e]) + st->kappa * delta[e];
            const double g = t - d * pt->x[e];
            Hd[e] = hl + g;
        }
        const double dt = c->embedded ? h3 : (h3 - b * g1);
        const double d_Hd = (((-h1 + a * b) + st->kappa * h2) + dt) - d * b;
        double alpha = 0.0, e_Pe_new = e_Pe;
        if (d_Hd != 0.0) {
            alpha = z_r / d_Hd;
            e_Pe_new = (e_Pe + (2.0 * alpha) * e_Pd) + (alpha * alpha) * d_Pd;
        }
        if (d_Hd <= 0.0 || e_Pe_new >= Delta2) {
            const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta2 - e_Pe))) / d_Pd;
            for (int e = 0; e < N; ++e) {
                eta[e] = eta[e] + tau * delta[e];
                Heta[e] = Heta[e] + tau * Hd[e];
            }
            res.stop = (d_Hd <= 0.0) ? TCG_NEGATIVE_CURVATURE : TCG_EXCEEDED_TR;
            ++j;
            break;
        }
        e_Pe = e_Pe_new;
        for (int e = 0; e < N; ++e) {
            new_eta[e] = eta[e] + alpha * delta[e];
            new_Heta[e] = Heta[e] + alpha * Hd[e];
            r_new[e] = r[e] + alpha * Hd[e];
        }
        const double new_model = vdot(c, new_eta, st->c) + 0.5 * vdot(c, new_eta, new_Heta);
        const double rr_new = vdot(c, r_new, r_new), xr = vdot(c, pt->x, r_new);
        if (new_model >= model_value) {
            res.stop = TCG_MODEL_INCREASED;
            ++j;
            break;
        }
        memcpy(eta, new_eta, sizeof(double) * N);
        memcpy(Heta, new_Heta, sizeof(double) * N);
        memcpy(r, r_new, sizeof(double) * N);
        model_value = new_model;
        r_r = rr_new;
        /* residual test on squares (||r|| <= target <=> r_r <= target^2): no square root on the critical path */
        if (j >= o->tcg_mininner && r_r <= target_sq) {
            res.stop = (o->tcg_kappa < nr_theta) ? TCG_REACHED_TARGET_LINEAR : TCG_REACHED_TARGET_SUPERLINEAR;
            ++j;
            break;
        }
        /* beta = z_r / z_r_old as a product with the reciprocal formed one iteration earlier (off the critical path) */
        const double beta = r_r * inv_zr;
        z_r = r_r;
        inv_zr = 1.0 / z_r;
        const double xd = -xr + beta * b; /* <x, -r + beta delta> */
        for (int e = 0; e < N; ++e) {
            const double dn = -r[e] + beta * delta[e];
            delta[e] = dn - xd * pt->x[e];
        }
        e_Pd = beta * (e_Pd + alpha * d_Pd);
        d_Pd = z_r + (beta * beta) * d_Pd;
    }
#endif
#else
    for (; j < maxinner; ++j) {
        Hw(c, pt, st, delta, Hd);
        const double d_Hd = vdot(c, delta, Hd);
        double alpha = 0.0, e_Pe_new = e_Pe;
        if (d_Hd != 0.0) {
            alpha = z_r / d_Hd;
            e_Pe_new = (e_Pe + (2.0 * alpha) * e_Pd) + (alpha * alpha) * d_Pd;
        }
        if (d_Hd <= 0.0 || e_Pe_new >= Delta2) {
            const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta2 - e_Pe))) / d_Pd;
            for (int e = 0; e < N; ++e) {
                eta[e] = eta[e] + tau * delta[e];
                Heta[e] = Heta[e] + tau * Hd[e];
            }
            res.stop = (d_Hd <= 0.0) ? TCG_NEGATIVE_CURVATURE : TCG_EXCEEDED_TR;
            ++j;
            break;
        }
        e_Pe = e_Pe_new;
        for (int e = 0; e < N; ++e) {
            new_eta[e] = eta[e] + alpha * delta[e];
            new_Heta[e] = Heta[e] + alpha * Hd[e];
        }
        const double new_model = vdot(c, new_eta, st->c) + 0.5 * vdot(c, new_eta, new_Heta);
        if (new_model >= model_value) {
            res.stop = TCG_MODEL_INCREASED;
            ++j;
            break;
        }
        memcpy(eta, new_eta, sizeof(double) * N);
        memcpy(Heta, new_Heta, sizeof(double) * N);
        model_value = new_model;
        for (int e = 0; e < N; ++e) r[e] = r[e] + alpha * Hd[e];
        r_r = vdot(c, r, r);
        const double norm_r = sqrt(r_r);
        if (j >= o->tcg_mininner && norm_r <= target) {
            res.stop = (o->tcg_kappa < nr_theta) ? TCG_REACHED_TARGET_LINEAR : TCG_REACHED_TARGET_SUPERLINEAR;
            ++j;
            break;
        }
        const double zold_rold = z_r;
        z_r = r_r;
        const double beta = z_r / zold_rold;
        for (int e = 0; e < N; ++e) delta[e] = -r[e] + beta * delta[e];
        project(c, pt, delta);
        e_Pd = beta * (e_Pd + alpha * d_Pd);
        d_Pd = z_r + (beta * beta) * d_Pd;
    }
#endif
    res.iters = j;
    res.model_value = model_value;
    return res;
}

/* ---- observers (utils.py:237-368; solver_warp.cuh evaluate<F>) -------------------------------------- */
typedef struct {
    double cost, distance, residual, gradnorm, compl_v, dual_v, man_v, max_v, mean_v;
} EvalRow;

static double lane_fold_sum(double* p) { return lane_tree(p); }

static EvalRow evaluate(const Ctx* c, const Pt* pt, const double* y, const double* xPrev) {
    EvalRow ev;
    double pc[LANES], pn[LANES], pi[LANES], ps[LANES];
    double pmax = 0.0;
    ev.cost = pt->cost;
    {
        double ip = vdot(c, xPrev, pt->x);
        ip = fmax(fmin(ip, 1.0), -1.0);
        ev.distance = acos(ip);
    }
    ev.gradnorm = gradL_norm(c, pt, y);
    for (int l = 0; l < LANES; ++l) {
        double a = 0.0, b = 0.0, d = 0.0, s = 0.0;
        for (int k = 0; k < c->K; ++k) {
            const int e = ELEM(l, k);
            if (e < c->n) {
                const double g = -pt->s[e];
                const double cv = y[e] * g;
                a = a + cv * cv;
                const double nv = fmax(-y[e], 0.0);
                b = b + nv * nv;
                const double iv = fmax(g, 0.0);
                d = d + iv * iv;
                s = s + iv;
                pmax = fmax(pmax, iv);
            }
        }
        pc[l] = a;
        pn[l] = b;
        pi[l] = d;
        ps[l] = s;
    }
    const double p_compl = lane_fold_sum(pc), p_nonneg = lane_fold_sum(pn), p_ineq = lane_fold_sum(pi),
                 p_sum = lane_fold_sum(ps);
    ev.compl_v = sqrt(p_compl);
    ev.dual_v = sqrt(p_nonneg);
    ev.man_v = sqrt(vdot(c, pt->x, pt->x)) - 1.0;
    ev.residual = sqrt(((((ev.gradnorm * ev.gradnorm + p_compl) + p_nonneg) + p_ineq) + 0.0) + ev.man_v * ev.man_v);
    ev.max_v = pmax;
    ev.mean_v = p_sum / (double)c->n;
    return ev;
}

typedef struct {
    double num_inner, radius, dxtype, tcg_iters, normdx, minxfeasi, minyfeasi, compl_v, ared_pred, radius_update,
        inner_status, dual_clipping;
} InnerInfo;

static InnerInfo empty_info(void) {
    InnerInfo i;
    i.num_inner = i.radius = i.dxtype = i.tcg_iters = i.normdx = i.minxfeasi = i.minyfeasi = i.compl_v = NAN;
    i.ared_pred = i.radius_update = i.inner_status = i.dual_clipping = NAN;
    return i;
}

static double max_abs_mult(const Ctx* c, const double* y) {
    double m = -INFINITY;
    for (int e = 0; e < c->n; ++e) m = fmax(m, fabs(y[e]));
    return m;
}

static void write_trace_row(double* row, int iteration, double mu, const InnerInfo* in, double maxabs, const EvalRow* ev) {
    row[0] = (double)iteration;
    row[1] = in->num_inner;
    row[2] = mu;
    row[3] = in->radius;
    row[4] = in->dxtype;
    row[5] = in->tcg_iters;
    row[6] = in->normdx;
    row[7] = in->minxfeasi;
    row[8] = in->minyfeasi;
    row[9] = in->compl_v;
    row[10] = in->ared_pred;
    row[11] = in->radius_update;
    row[12] = in->inner_status;
    row[13] = in->dual_clipping;
    row[14] = maxabs;
    row[15] = ev->cost;
    row[16] = ev->distance;
    row[17] = ev->residual;
    row[18] = ev->gradnorm;
    row[19] = ev->compl_v;
    row[20] = ev->dual_v;
    row[21] = ev->man_v;
    row[22] = ev->max_v;
    row[23] = ev->mean_v;
    row[24] = 0.0; /* time: the oracle does not model wall clock */
}

typedef struct {
    double inner, tcg, aux;
} Counters;

/* per-lane sequential sum over k of f(e), then the lane tree */
#define LANE_SUM(RESULT, EXPR_ACTIVE)                             \
    do {                                                          \
        double _p[LANES];                                         \
        for (int l = 0; l < LANES; ++l) {                         \
            double _a = 0.0;                                      \
            for (int k = 0; k < c->K; ++k) {                      \
                const int e = ELEM(l, k);                      \
                if (e < c->n) _a = _a + (EXPR_ACTIVE);            \
            }                                                     \
            _p[l] = _a;                                           \
        }                                                         \
        RESULT = lane_tree(_p);                                   \
    } while (0)

/* ---- one trust-region iteration (RIPTRM.py:707-783; solver_warp.cuh inner_step<F>) -------------------- */
static int inner_step(const Ctx* c, const det_options* o, Pt* pt, double* y, double mu, double* Delta, double tolL,
                      double tolC, int k_inner, InnerInfo* info, Counters* cnt) {
    const int N = c->K * LANES;
    *info = empty_info();
    info->num_inner = (double)k_inner;
    info->radius = *Delta;
    Step st;
    begin_step(c, pt, y, mu, &st);
    double dx[MAXN], Hdx_tcg[MAXN];
    const TcgResult tr = tcg(c, o, pt, &st, *Delta, dx, Hdx_tcg);
    cnt->tcg += (double)tr.iters;
    info->dxtype = (double)tr.stop;
    info->tcg_iters = (double)tr.iters;
    const double normdx = sqrt(vdot(c, dx, dx));
    info->normdx = normdx;
    double ga[MAXN], yNew[MAXN], xN[MAXN];
    gadj(c, pt, dx, ga);
    for (int e = 0; e < N; ++e) {
        if (e < c->n) {
            const double dy = (-y[e] + mu * (1.0 / pt->s[e])) - (y[e] * ga[e]) / pt->s[e];
            yNew[e] = y[e] + dy;
        } else {
            yNew[e] = 0.0;
        }
    }
    Pt ptN;
    retract(c, pt, dx, xN);
    eval_point(c, xN, &ptN);
    int xfe = 1, yfe = 1;
    double mins = INFINITY, miny = INFINITY, p_c;
    for (int e = 0; e < c->n; ++e) {
        xfe = xfe && (ptN.s[e] > 0.0);
        yfe = yfe && (yNew[e] > 0.0);
        mins = fmin(mins, ptN.s[e]);
        miny = fmin(miny, yNew[e]);
    }
    LANE_SUM(p_c, ((yNew[e] * ptN.s[e] - mu) * (yNew[e] * ptN.s[e] - mu)));
    const double compl_v = sqrt(p_c);
    const double ngl = gradL_norm(c, &ptN, yNew);
    info->minxfeasi = mins;
    info->minyfeasi = miny;
    info->compl_v = compl_v;
    if (xfe && yfe && ngl <= tolL && compl_v <= tolC) {
        info->inner_status = (double)INNER_CONVERGED;
        *pt = ptN;
        memcpy(y, yNew, sizeof(double) * N);
        return 1;
    }
    if (!xfe) {
        info->inner_status = (double)INNER_PRIMAL_INFEASIBLE;
        *Delta = o->gamma * normdx;
        return 0;
    }
    double pl_cur, pl_new;
    LANE_SUM(pl_cur, det_log(pt->s[e]));
    LANE_SUM(pl_new, det_log(ptN.s[e]));
    const double phi_cur = pt->cost - mu * pl_cur;
    const double phi_new = ptN.cost - mu * pl_new;
    double ared = phi_cur - phi_new;
    /* the Hw[dx] of RIPTRM.py:659: the product the tCG accumulated beside eta (solver_warp.cuh inner_step), or a fresh one
     * with RIPTRM_RECOMPUTE_HDX=1 in the environment (the reference's form; the switch the CUDA library reads too) */
    double Hdx[MAXN];
    if (getenv("RIPTRM_RECOMPUTE_HDX") == NULL) {
        memcpy(Hdx, Hdx_tcg, sizeof(Hdx));
    } else {
        Hw(c, pt, &st, dx, Hdx);
        cnt->aux += 1.0;
    }
    double pred = (0.0 - 0.5 * vdot(c, Hdx, dx)) - vdot(c, st.c, dx);
    const double reg = (fmax(1.0, fabs(phi_cur)) * 2.220446049250313e-16) * o->reduction_regularization;
    ared = ared + reg;
    pred = pred + reg;
    info->ared_pred = ared / pred;
    double DeltaNext;
    if (ared < 0.25 * pred) {
        info->radius_update = (double)RADIUS_REDUCED;
        DeltaNext = 0.25 * *Delta;
    } else if (ared >= 0.75 * pred && fabs(normdx - *Delta) <= 1e-15) {
        info->radius_update = (double)RADIUS_EXPANDED;
        DeltaNext = fmin(2.0 * *Delta, o->maximal_tr_radius);
    } else {
        info->radius_update = (double)RADIUS_UNCHANGED;
        DeltaNext = *Delta;
    }
    if (ared > o->rho * pred) {
        info->inner_status = (double)INNER_SUCCESSFUL;
        const double I_right = fmax(o->const_right, o->const_right / mu);
        int clipped_any = 0;
        for (int e = 0; e < c->n; ++e) {
            const double I_left = o->const_left * fmin(fmin(y[e], mu / ptN.s[e]), 1.0);
            const double cl = fmin(fmax(yNew[e], I_left), I_right);
            clipped_any = clipped_any || !(cl == yNew[e]);
            yNew[e] = cl;
        }
        info->dual_clipping = clipped_any ? 1.0 : 0.0;
        *pt = ptN;
        memcpy(y, yNew, sizeof(double) * N);
    } else {
        info->inner_status = (double)INNER_UNSUCCESSFUL;
    }
    *Delta = DeltaNext;
    return 0;
}

/* ---- whole solve (RIPTRM.py:909-976; solver_warp.cuh solve_instance<F>) ------------------------------- */
int riptrm_det_solve_nonnegpca(int n, const double* Z, const double* x0, const double* y0, double eps,
                               const det_options* o, double* x_out, double* y_out, double* summary, double* trace) {
    if (n <= 0 || n > MAXN) return -1;
    Ctx ctx;
    Ctx* c = &ctx;
    c->n = n;
    c->K = (n <= 64) ? 2 : 4;
    c->eps = eps;
    c->embedded = o->is_euclidean_embedded != 0;
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) {
            const double s = Z[i * n + j] + Z[j * n + i];
            c->S[i * n + j] = s;
            c->S[j * n + i] = s;
        }
    const int N = c->K * LANES;
    double xin[MAXN], y[MAXN], xPrev[MAXN];
    memset(xin, 0, sizeof xin);
    memset(y, 0, sizeof y);
    memcpy(xin, x0, sizeof(double) * n);
    memcpy(y, y0, sizeof(double) * n);
    Pt pt;
    eval_point(c, xin, &pt);
    double Delta = o->initial_tr_radius > 0.0 ? o->initial_tr_radius : 3.141592653589793 / 8.0;
    memcpy(xPrev, pt.x, sizeof(double) * N);
    InnerInfo info = empty_info();
    Counters cnt = {0.0, 0.0, 0.0};
    int it = 0, rows = 0, stop_reason = STOP_RUNNING;
    double mu = o->mu_sched[0];
    EvalRow ev;
    for (;;) {
        ev = evaluate(c, &pt, y, xPrev);
        if (o->trace_mode != 0 && (it == 0 || o->trace_mode == 2)) {
            if (trace && rows < o->trace_capacity)
                write_trace_row(trace + (size_t)rows * TRACE_FIELDS, it, mu, &info, max_abs_mult(c, y), &ev);
            ++rows;
        }
        memcpy(xPrev, pt.x, sizeof(double) * N);
        if (it >= o->maxiter) stop_reason = STOP_MAXITER;
        if (ev.residual <= o->tolresid) stop_reason = STOP_TOLRESID;
        if (stop_reason != STOP_RUNNING) break;
        it += 1;
        mu = o->mu_sched[it - 1];
        const double tolL = o->tol_lagrangian_sched[it - 1], tolC = o->tol_complementarity_sched[it - 1];
        const Pt pt_init = pt;
        double y_init[MAXN], xPrevInner[MAXN];
        memcpy(y_init, y, sizeof(double) * N);
        const double Delta_init = Delta;
        memcpy(xPrevInner, pt.x, sizeof(double) * N);
        int k = 0;
        for (;;) {
            k += 1;
            int done = inner_step(c, o, &pt, y, mu, &Delta, tolL, tolC, k, &info, &cnt);
            cnt.inner += 1.0;
            if (o->trace_mode == 1) {
                if (trace && rows < o->trace_capacity) {
                    const EvalRow evi = evaluate(c, &pt, y, xPrevInner);
                    write_trace_row(trace + (size_t)rows * TRACE_FIELDS, it, mu, &info, max_abs_mult(c, y), &evi);
                }
                ++rows;
            }
            memcpy(xPrevInner, pt.x, sizeof(double) * N);
            if (o->inner_maxiter >= 0 && k >= o->inner_maxiter) {
                info.inner_status = (double)INNER_MAX_ITER;
                done = 1;
                pt = pt_init;
                memcpy(y, y_init, sizeof(double) * N);
                Delta = Delta_init;
            }
            if (done) break;
        }
        mu = o->mu_sched[it];
        Delta = fmax(Delta, o->minimal_initial_tr_radius);
    }
    if (x_out) memcpy(x_out, pt.x, sizeof(double) * n);
    if (y_out) memcpy(y_out, y, sizeof(double) * n);
    if (summary) {
        summary[0] = ev.cost;
        summary[1] = ev.residual;
        summary[2] = ev.gradnorm;
        summary[3] = ev.compl_v;
        summary[4] = ev.dual_v;
        summary[5] = ev.man_v;
        summary[6] = ev.max_v;
        summary[7] = ev.mean_v;
        summary[8] = mu;
        summary[9] = Delta;
        summary[10] = (double)it;
        summary[11] = cnt.inner;
        summary[12] = cnt.tcg;
        summary[13] = cnt.aux;
        summary[14] = (double)stop_reason;
        summary[15] = (double)rows;
    }
    return 0;
}

/* Hooks for unit parity: one Hessian-vector product / one tCG solve at (x, y, mu[, Delta]) */
int riptrm_det_hessvec(int n, const double* Z, const double* x, const double* y, double eps, double mu, const double* v,
                       double* out) {
    det_options o;
    memset(&o, 0, sizeof o);
    double one = 0.0;
    o.mu_sched = o.tol_lagrangian_sched = o.tol_complementarity_sched = &one;
    if (n <= 0 || n > MAXN) return -1;
    static Ctx ctx; /* not reentrant: hooks are single-threaded test helpers */
    Ctx* c = &ctx;
    c->n = n;
    c->K = (n <= 64) ? 2 : 4;
    c->eps = eps;
    c->embedded = 0;
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) {
            const double s = Z[i * n + j] + Z[j * n + i];
            c->S[i * n + j] = s;
            c->S[j * n + i] = s;
        }
    double xin[MAXN], yin[MAXN], vin[MAXN], hv[MAXN];
    memset(xin, 0, sizeof xin);
    memset(yin, 0, sizeof yin);
    memset(vin, 0, sizeof vin);
    memcpy(xin, x, sizeof(double) * n);
    memcpy(yin, y, sizeof(double) * n);
    memcpy(vin, v, sizeof(double) * n);
    Pt pt;
    Step st;
    eval_point(c, xin, &pt);
    begin_step(c, &pt, yin, mu, &st);
    Hw(c, &pt, &st, vin, hv);
    memcpy(out, hv, sizeof(double) * n);
    return 0;
}

/* Solves `count` pairs one after another (the CPU baseline leg calls this from one thread per chunk).
 * Z [count][n][n], x0/y0 [count][n], summary [count][16]. */
int riptrm_det_solve_many(int count, int n, const double* Z, const double* x0, const double* y0, double eps,
                          const det_options* o, double* x_out, double* y_out, double* summary) {
    for (int i = 0; i < count; ++i) {
        int rc = riptrm_det_solve_nonnegpca(n, Z + (size_t)i * n * n, x0 + (size_t)i * n, y0 + (size_t)i * n, eps, o,
                                            x_out ? x_out + (size_t)i * n : 0, y_out ? y_out + (size_t)i * n : 0,
                                            summary ? summary + (size_t)i * SUMMARY_FIELDS : 0, 0);
        if (rc) return rc;
    }
    return 0;
}

// solver_warp.cuh -- the whole RIPTRM solve for ONE (instance, initialpoint) pair, executed by
// ONE warp, generic over a problem family F (fam_*.cuh).  Everything the reference does per
// solve lives here: the outer barrier loop (RIPTRM.py:909-976, :866-896), the inner
// trust-region loop (:785-847), one trust-region iteration (:707-783), the Steihaug-Toint
// truncated CG with its boundary-crossing tau solve (:41-216), the stopping tests (:574-629),
// the rho test / radius update / dual clipping (:631-705), and the observers that produce the
// reference's log row (utils.py:237-368; RIPTRM.py:980-1024).
//
// A family F provides (all warp-collective, every lane calls):
//   Vec / CVec           tangent-/constraint-space vectors (WVec<K> / WVec<MK>)
//   Ctx, Pt, Step        per-instance data, per-point cache (x, s, cost, ...), per-step cache (c, y/s, ...)
//   eval_point, begin_step, Hw, gadj, inner, project, retract, gradL_norm, manvio, dist,
//   cactive(ctx,k), dim(ctx), typical_dist(ctx)
#pragma once
#include <type_traits>
#include "common.cuh"
#include "dense_trs.cuh"
#include "../../include/riptrm_b200.h"

namespace riptrm {

struct DevOpts {
    int maxiter, inner_maxiter, tcg_mininner, tcg_maxinner, is_euclidean_embedded, trace_mode, trace_capacity;
    double tolresid, maxtime, inner_maxtime, initial_tr_radius, minimal_initial_tr_radius, maximal_tr_radius;
    double rho, reduction_regularization, gamma, const_left, const_right, tcg_theta, tcg_kappa;
    const double* mu;    // device, length maxiter + 1
    const double* tolL;  // device
    const double* tolC;  // device
    // TRS_solver='Exact_RepMat' (RIPTRM.py:431-444) and the second-order test (:599-617)
    int second_order;
    double trs_tolhardcase;
    const double* tolS;  // device: forcing_function_second_order(mu) per outer iteration
    // the Hw[dx] of the model decrease (RIPTRM.py:659): 0 = the product the tCG has accumulated (families whose tCG returns
    // it, `F::kTcgReturnsHw`), 1 = a fresh Hessian-vector product as the reference writes it (RIPTRM_RECOMPUTE_HDX=1)
    int recompute_hdx;
};

// families that take the inner products a trust-region iteration needs from dx alone -- <dx,dx> (:735), the projection
// coefficient of G*[dx] (:743), the retraction's norm (:744), <Hw dx, dx> and <c, dx> (:659-660) -- in ONE reduction round
// instead of four declare `static constexpr bool kMergedStepDots = true` and provide step_dots / gadj_given / retract_given
// (same summation tree per value, hence the same bits)
template <class F, class = void>
struct MergedStepDots { static constexpr bool value = false; };
template <class F>
struct MergedStepDots<F, std::enable_if_t<F::kMergedStepDots>> { static constexpr bool value = true; };

// families whose tcg() hands back Hw[eta] next to eta declare `static constexpr bool kTcgReturnsHw = true`
template <class F, class = void>
struct TcgReturnsHw { static constexpr bool value = false; };
template <class F>
struct TcgReturnsHw<F, std::enable_if_t<F::kTcgReturnsHw>> { static constexpr bool value = true; };

// ------------------------------------------------------------------------------------------
// Exact_RepMat: the representation matrix of Hw in an orthonormal tangent basis and its eigen-decomposition.
// A family provides the isometry between T_x M and R^dim:
//   Coord                                   per-point data of the basis (registers / the family's scratch)
//   coord_setup(ctx, pt, Coord&)            basisfun(manifold, x)                              (RIPTRM.py:433, :600)
//   from_coords(ctx, pt, cc, coef) -> Vec   sum_i coef_i b_i                                   (:441-443)
//   to_coords(ctx, pt, cc, v, out)          out_i = <v, b_i>_x                                  (:436-438, utils.py:570)
// `RepWork` is one warp's shared-memory workspace; slot `cur` holds the decomposition that belongs to the current (x, y)
// (the reference's is_RepMat_available / basisxCur / HwCurmatrix / cxCurvector, :415-421), slot 1 - cur the one built at
// the trial point for the eigenvalue test (basisxNew / HwNewmatrix / cxNewvector, :613-615).
// ------------------------------------------------------------------------------------------
struct RepWork {
    double* W;       // d x ld scratch: the matrix handed to jacobi_sym; hard-case system afterwards
    double* VT[2];   // d x ld each: rows = eigenvectors
    double* D[2];    // eigenvalues
    double* al[2];   // coordinates of c in the eigenbasis
    double* coef;    // d
    double* col;     // d
    double* ws;      // 6 d
    int d, ld;
    int cur;
    bool mat_valid;  // (D, VT)[cur] belong to the current (x, y)
    bool al_valid;   // al[cur] belongs to the current (x, mu)
};
__host__ __device__ __forceinline__ int rep_ld(int d) { return (d + 1) | 1; }
__host__ __device__ __forceinline__ size_t rep_doubles(int d) { return (size_t)3 * d * rep_ld(d) + (size_t)12 * d; }
__device__ __forceinline__ void rep_init(RepWork& rw, double* mem, int d) {
    const int ld = rep_ld(d);
    rw.d = d;
    rw.ld = ld;
    rw.W = mem;
    rw.VT[0] = mem + (size_t)d * ld;
    rw.VT[1] = mem + (size_t)2 * d * ld;
    double* v = mem + (size_t)3 * d * ld;
    rw.D[0] = v;
    rw.D[1] = v + d;
    rw.al[0] = v + 2 * d;
    rw.al[1] = v + 3 * d;
    rw.coef = v + 4 * d;
    rw.col = v + 5 * d;
    rw.ws = v + 6 * d;
    rw.cur = 0;
    rw.mat_valid = rw.al_valid = false;
}
struct NoRep {};

// selfadj_operator2matrix (utils.py:565-573) + eigen-decomposition into slot `slot`: column j = coordinates of Hw[b_j], the
// upper triangle is kept and mirrored.  Returns the number of Hessian-vector products (dim).
template <class F>
__device__ __forceinline__ int rep_build(const typename F::Ctx& ctx, const typename F::Pt& pt, const typename F::CVec& y,
                                         const typename F::Step& st, const typename F::Coord& cc, RepWork& rw, int slot) {
    const int d = rw.d, ld = rw.ld, lane = lane_id();
    for (int j = 0; j < d; ++j) {
        for (int k = lane; k < d; k += 32) rw.coef[k] = (k == j) ? 1.0 : 0.0;
        __syncwarp();
        const typename F::Vec b = F::from_coords(ctx, pt, cc, rw.coef);
        const typename F::Vec hb = F::Hw(ctx, pt, y, st, b);
        F::to_coords(ctx, pt, cc, hb, rw.col);
        for (int i = lane; i <= j; i += 32) {
            rw.W[i * ld + j] = rw.col[i];
            rw.W[j * ld + i] = rw.col[i];
        }
        __syncwarp();
    }
    dense::jacobi_sym(rw.W, rw.VT[slot], d, ld);
    for (int k = lane; k < d; k += 32) rw.D[slot][k] = rw.W[k * ld + k];
    __syncwarp();
    return d;
}
template <class F>
__device__ __forceinline__ void rep_rhs(const typename F::Ctx& ctx, const typename F::Pt& pt, const typename F::Step& st,
                                        const typename F::Coord& cc, RepWork& rw, int slot) {
    F::to_coords(ctx, pt, cc, st.c, rw.col);                         // cxCurvector (:436-438)
    dense::rows_dot(rw.VT[slot], rw.d, rw.ld, rw.col, rw.al[slot]);
}
__device__ __forceinline__ double rep_mineig(const RepWork& rw, int slot) {
    double m = CUDART_INF;
    for (int k = lane_id(); k < rw.d; k += 32) m = fmin(m, rw.D[slot][k]);
    return wmin(m);
}

struct TcgResult {
    int iters;  // j + 1
    int stop;   // riptrm_tcg_stop
    double model_value;
};

// ------------------------------------------------------------------------------------------
// Steihaug-Toint truncated CG, eta0 = 0, identity preconditioner  (RIPTRM.py:41-216), in the reference's
// operation order, for any family.  A family's F::tcg either forwards here or provides a form with merged
// reductions (fam_sphere.cuh).
// ------------------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ TcgResult tcg_generic(const typename F::Ctx& ctx, const DevOpts& o, const typename F::Pt& pt,
                                         const typename F::CVec& y, const typename F::Step& st, double Delta,
                                         typename F::Vec& eta, typename F::Vec& Heta) {
    using Vec = typename F::Vec;
    constexpr int K = F::K;
    eta = wzero<K>();
    Heta = wzero<K>();                                     // :47
    Vec r = st.c;                                          // :48
    double e_Pe = 0.0;                                     // :49
    double r_r = F::inner(ctx, pt, r, r);                  // :56
    const double norm_r0 = sqrt(r_r);                      // :57-58
    double z_r = r_r;                                      // :62,:67 (z = r: identity preconditioner)
    double d_Pd = z_r;                                     // :68
    Vec delta;
#pragma unroll
    for (int k = 0; k < K; ++k) delta.v[k] = -r.v[k];      // :71
    double e_Pd = 0.0;                                     // :73
    double model_value = 0.0;                              // :90
    TcgResult res;
    res.stop = RIPTRM_TCG_MAX_INNER_ITER;                  // :95
    const int maxinner = o.tcg_maxinner < 0 ? F::dim(ctx) : o.tcg_maxinner;
    const double Delta2 = Delta * Delta;
    // min(norm_r0**theta, kappa)  (:183-185); pow only when theta != 1
    const double nr_theta = (o.tcg_theta == 1.0) ? norm_r0 : pow(norm_r0, o.tcg_theta);
    const double target = norm_r0 * fmin(nr_theta, o.tcg_kappa);
    int j = 0;
    for (; j < maxinner; ++j) {                            // :98
        Vec Hd = F::Hw(ctx, pt, y, st, delta);             // :100
        const double d_Hd = F::inner(ctx, pt, delta, Hd);  // :103
        double alpha = 0.0, e_Pe_new = e_Pe;
        if (d_Hd != 0.0) {                                 // :106-114
            alpha = z_r / d_Hd;
            e_Pe_new = (e_Pe + (2.0 * alpha) * e_Pd) + (alpha * alpha) * d_Pd;
        }
        if (d_Hd <= 0.0 || e_Pe_new >= Delta2) {           // :118
            const double tau = (-e_Pd + sqrt(e_Pd * e_Pd + d_Pd * (Delta2 - e_Pe))) / d_Pd;  // :123-125
#pragma unroll
            for (int k = 0; k < K; ++k) {
                eta.v[k] = eta.v[k] + tau * delta.v[k];    // :127
                Heta.v[k] = Heta.v[k] + tau * Hd.v[k];     // :132
            }
            res.stop = (d_Hd <= 0.0) ? RIPTRM_TCG_NEGATIVE_CURVATURE : RIPTRM_TCG_EXCEEDED_TR;  // :142-145
            ++j;
            break;
        }
        e_Pe = e_Pe_new;                                   // :149
        Vec new_eta, new_Heta;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            new_eta.v[k] = eta.v[k] + alpha * delta.v[k];  // :150
            new_Heta.v[k] = Heta.v[k] + alpha * Hd.v[k];   // :154
        }
        // r_new is formed before the model test so that its norm rides in the same butterfly as the two
        // model-value dot products (three sums, one reduction latency); the values are those of :162 / :175
        Vec r_new;
#pragma unroll
        for (int k = 0; k < K; ++k) r_new.v[k] = r.v[k] + alpha * Hd.v[k];  // :172
        double ip_ec = F::inner_partial(ctx, pt, new_eta, st.c);
        double ip_eh = F::inner_partial(ctx, pt, new_eta, new_Heta);
        double ip_rr = F::inner_partial(ctx, pt, r_new, r_new);
        wsum3(ip_ec, ip_eh, ip_rr);
        const double new_model = ip_ec + 0.5 * ip_eh;  // :86-87,:162
        if (new_model >= model_value) {                    // :163
            res.stop = RIPTRM_TCG_MODEL_INCREASED;
            ++j;
            break;
        }
        eta = new_eta;                                     // :167-169
        Heta = new_Heta;
        model_value = new_model;
        r = r_new;
        r_r = ip_rr;                                       // :175
        const double norm_r = sqrt(r_r);
        if (j >= o.tcg_mininner && norm_r <= target) {     // :183-191
            res.stop = (o.tcg_kappa < nr_theta) ? RIPTRM_TCG_REACHED_TARGET_LINEAR
                                                : RIPTRM_TCG_REACHED_TARGET_SUPERLINEAR;
            ++j;
            break;
        }
        const double zold_rold = z_r;                      // :200
        z_r = r_r;                                         // :202
        const double beta = z_r / zold_rold;               // :205
#pragma unroll
        for (int k = 0; k < K; ++k) delta.v[k] = -r.v[k] + beta * delta.v[k];  // :206
        delta = F::project(ctx, pt, delta);                // :210
        e_Pd = beta * (e_Pd + alpha * d_Pd);               // :213
        d_Pd = z_r + (beta * beta) * d_Pd;                 // :214
    }
    res.iters = j;  // j + 1 of the reference (loop index of the last executed iteration, plus one)
    res.model_value = model_value;
    return res;
}

// ------------------------------------------------------------------------------------------
// The condensed Newton system of the Riemannian interior-point method (RIPM.py:484-511):
//     Aw[dx] = Hess_x L(x, z)[dx] + G_x( G*_x[dx] * z / s ) = c
// is the operator of this file's trust-region model with the slack s an INDEPENDENT variable (RIPM keeps (x, y, z, s)); `st` is
// a Step whose ys holds z / s.  Two solvers, as in the reference:
//   newton_repmat  RepresentMatMethod (RIPM.py:238-300, no equality constraints: T_mat = Aw_mat): the representation matrix
//                  in the tangent basis, solved through its symmetric eigen-decomposition (scipy.linalg.solve(assume_a='sym'))
//   newton_cr      TangentSpaceConjResMethod (utils.py:582-618; Saad, Iterative Methods, Alg. 6.20): conjugate residuals on the
//                  tangent space, v0 = 0, stop at |r| / |c| < tol or after maxiter iterations
// ------------------------------------------------------------------------------------------
struct NewtonOut {
    double iters, rel_res, mineig;
};

template <class F>
__device__ __forceinline__ NewtonOut newton_repmat(const typename F::Ctx& ctx, const typename F::Pt& pt, const typename F::CVec& z,
                                                   const typename F::Step& st, const typename F::Vec& c, RepWork& rw,
                                                   typename F::Vec& dx) {
    typename F::Coord cc;
    F::coord_setup(ctx, pt, cc);
    rep_build<F>(ctx, pt, z, st, cc, rw, 0);
    F::to_coords(ctx, pt, cc, c, rw.col);                           // tangent2vec (utils.py:575-580)
    dense::rows_dot(rw.VT[0], rw.d, rw.ld, rw.col, rw.al[0]);
    for (int k = lane_id(); k < rw.d; k += 32) rw.col[k] = rw.al[0][k] / rw.D[0][k];
    __syncwarp();
    dense::cols_dot(rw.VT[0], rw.d, rw.ld, rw.col, rw.coef);
    dx = F::from_coords(ctx, pt, cc, rw.coef);
    // A solve through the eigen-decomposition carries a forward error of eps * cond in EVERY direction (the reference's LAPACK
    // sysv is backward stable: residual eps |A| |dx|); two steps of iterative refinement on the operator restore that
    typename F::Vec r;
    for (int pass = 0; pass < 3; ++pass) {
        const typename F::Vec Adx = F::Hw(ctx, pt, z, st, dx);
#pragma unroll
        for (int k = 0; k < F::K; ++k) r.v[k] = c.v[k] - Adx.v[k];
        if (pass == 2) break;
        F::to_coords(ctx, pt, cc, r, rw.col);
        dense::rows_dot(rw.VT[0], rw.d, rw.ld, rw.col, rw.al[0]);
        for (int k = lane_id(); k < rw.d; k += 32) rw.col[k] = rw.al[0][k] / rw.D[0][k];
        __syncwarp();
        dense::cols_dot(rw.VT[0], rw.d, rw.ld, rw.col, rw.coef);
        const typename F::Vec corr = F::from_coords(ctx, pt, cc, rw.coef);
#pragma unroll
        for (int k = 0; k < F::K; ++k) dx.v[k] = dx.v[k] + corr.v[k];
    }
    NewtonOut out;
    out.iters = 0.0;
    out.rel_res = sqrt(F::inner(ctx, pt, r, r)) / sqrt(F::inner(ctx, pt, c, c));
    out.mineig = rep_mineig(rw, 0);
    return out;
}

template <class F>
__device__ __forceinline__ NewtonOut newton_cr(const typename F::Ctx& ctx, const typename F::Pt& pt, const typename F::CVec& z,
                                               const typename F::Step& st, const typename F::Vec& b, double tol, int maxiter,
                                               typename F::Vec& v) {
    using Vec = typename F::Vec;
    constexpr int K = F::K;
    v = wzero<K>();                                                 // v0 = 0 (RIPM.py:431 self.v0)
    Vec r = b, p = b;                                               // r = b - A(v0)
    const double b_norm = sqrt(F::inner(ctx, pt, b, b));
    Vec Ar = F::Hw(ctx, pt, z, st, r), Ap = Ar;
    double rAr = F::inner(ctx, pt, r, Ar);
    double rel_res = 1.0;
    int t = 0;
    while (true) {
        t += 1;
        const double a = rAr / F::inner(ctx, pt, Ap, Ap);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            v.v[k] = v.v[k] + a * p.v[k];
            r.v[k] = r.v[k] - a * Ap.v[k];
        }
        rel_res = sqrt(F::inner(ctx, pt, r, r)) / b_norm;
        if (rel_res < tol || t == maxiter) break;
        Ar = F::Hw(ctx, pt, z, st, r);
        const double old = rAr;
        rAr = F::inner(ctx, pt, r, Ar);
        const double beta = rAr / old;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            p.v[k] = r.v[k] + beta * p.v[k];
            Ap.v[k] = Ar.v[k] + beta * Ap.v[k];
        }
    }
    NewtonOut out;
    out.iters = (double)t;
    out.rel_res = rel_res;
    out.mineig = CUDART_NAN;
    return out;
}

// ------------------------------------------------------------------------------------------
// Observers: utils.evaluation / compute_residual / compute_maxmeanviolations (utils.py:237-368)
// ------------------------------------------------------------------------------------------
struct EvalRow {
    double cost, distance, residual, gradnorm, compl_v, dual_v, man_v, max_v, mean_v;
};

template <class F>
__device__ __forceinline__ EvalRow evaluate(const typename F::Ctx& ctx, const typename F::Pt& pt,
                                            const typename F::CVec& y, const typename F::Vec& xPrev, bool want_distance = true) {
    constexpr int MK = F::MK;
    EvalRow ev;
    ev.cost = pt.cost;
    // `distance` only goes to the log; without a trace the SPD / Grassmann geodesic distance (two Jacobi eigen-solves per
    // component: a fifth of the StableIdentification kernel's instructions) is not computed
    ev.distance = want_distance ? F::dist(ctx, xPrev, pt) : 0.0;
    ev.gradnorm = F::gradL_norm(ctx, pt, y);
    double p_compl = 0.0, p_nonneg = 0.0, p_ineq = 0.0, p_sum = 0.0, p_max = 0.0;
#pragma unroll
    for (int k = 0; k < MK; ++k) {
        if (F::cactive(ctx, k)) {
            const double g = -pt.s.v[k];
            const double cv = y.v[k] * g;
            p_compl = p_compl + cv * cv;                    // utils.py:297-301
            const double nv = fmax(-y.v[k], 0.0);
            p_nonneg = p_nonneg + nv * nv;                  // :305-309
            const double iv = fmax(g, 0.0);
            p_ineq = p_ineq + iv * iv;                      // :313-318
            p_sum = p_sum + iv;                             // :250-255
            p_max = fmax(p_max, iv);
        }
    }
    wsum3(p_compl, p_nonneg, p_ineq);
    p_sum = wsum(p_sum);
    p_max = wmax(p_max);
    ev.compl_v = sqrt(p_compl);
    ev.dual_v = sqrt(p_nonneg);
    ev.man_v = F::manvio(ctx, pt);
    ev.residual = sqrt(((((ev.gradnorm * ev.gradnorm + p_compl) + p_nonneg) + p_ineq) + 0.0) +
                       ev.man_v * ev.man_v);                // :332-338
    ev.max_v = p_max;
    ev.mean_v = p_sum / (double)F::num_constraints(ctx);   // :264-265
    return ev;
}

struct InnerInfo {
    double num_inner, radius, dxtype, tcg_iters, normdx, minxfeasi, minyfeasi, compl_v, ared_pred, radius_update,
        inner_status, dual_clipping, mineig;
};

__device__ __forceinline__ InnerInfo empty_info() {
    const double nan = CUDART_NAN;
    InnerInfo i;
    i.num_inner = i.radius = i.dxtype = i.tcg_iters = i.normdx = i.minxfeasi = i.minyfeasi = i.compl_v = nan;
    i.ared_pred = i.radius_update = i.inner_status = i.dual_clipping = i.mineig = nan;
    return i;
}

// One coalesced 200-byte store: lane f writes field f.
__device__ __forceinline__ void write_trace_row(double* row, int iteration, double mu, const InnerInfo& in,
                                                double maxabs, const EvalRow& ev, double time_s) {
    const int l = lane_id();
    double v = 0.0;
    switch (l) {
        case RIPTRM_TR_ITERATION: v = (double)iteration; break;
        case RIPTRM_TR_NUM_INNER: v = in.num_inner; break;
        case RIPTRM_TR_MU: v = mu; break;
        case RIPTRM_TR_RADIUS: v = in.radius; break;
        case RIPTRM_TR_DXTYPE: v = in.dxtype; break;
        case RIPTRM_TR_TCG_ITERS: v = in.tcg_iters; break;
        case RIPTRM_TR_NORMDX: v = in.normdx; break;
        case RIPTRM_TR_MINXFEASI: v = in.minxfeasi; break;
        case RIPTRM_TR_MINYFEASI: v = in.minyfeasi; break;
        case RIPTRM_TR_COMPL: v = in.compl_v; break;
        case RIPTRM_TR_ARED_PRED: v = in.ared_pred; break;
        case RIPTRM_TR_RADIUS_UPDATE: v = in.radius_update; break;
        case RIPTRM_TR_INNER_STATUS: v = in.inner_status; break;
        case RIPTRM_TR_DUAL_CLIPPING: v = in.dual_clipping; break;
        case RIPTRM_TR_MAXABSLAGMULT: v = maxabs; break;
        case RIPTRM_TR_COST: v = ev.cost; break;
        case RIPTRM_TR_DISTANCE: v = ev.distance; break;
        case RIPTRM_TR_RESIDUAL: v = ev.residual; break;
        case RIPTRM_TR_GRADNORM: v = ev.gradnorm; break;
        case RIPTRM_TR_COMPLVIOLATION: v = ev.compl_v; break;
        case RIPTRM_TR_DUALVIOLATION: v = ev.dual_v; break;
        case RIPTRM_TR_MANVIOLATION: v = ev.man_v; break;
        case RIPTRM_TR_MAXVIOLATION: v = ev.max_v; break;
        case RIPTRM_TR_MEANVIOLATION: v = ev.mean_v; break;
        case RIPTRM_TR_TIME: v = time_s; break;
        case RIPTRM_TR_MINEIGVALHW: v = in.mineig; break;
        default: break;
    }
    if (l < RIPTRM_TRACE_FIELDS) row[l] = v;
}

template <class F>
__device__ __forceinline__ double max_abs_mult(const typename F::Ctx& ctx, const typename F::CVec& y) {
    double m = -CUDART_INF;  // RIPTRM.py:1020-1022
#pragma unroll
    for (int k = 0; k < F::MK; ++k)
        if (F::cactive(ctx, k)) m = fmax(m, fabs(y.v[k]));
    return wmax(m);
}

// ------------------------------------------------------------------------------------------
// One trust-region iteration (RIPTRM.py:707-783).  Updates (pt, y, Delta) in place.
// Returns true when the inner loop converged.
// ------------------------------------------------------------------------------------------
struct Counters {
    double inner, tcg, aux;
};

template <class F, bool EXACT = false, class Rep = NoRep>
__device__ __forceinline__ bool inner_step(const typename F::Ctx& ctx, const DevOpts& o, typename F::Pt& pt,
                                           typename F::CVec& y, double mu, double& Delta, double tolL, double tolC,
                                           int k_inner, InnerInfo& info, Counters& cnt, Rep& rw, double tolS = 0.0) {
    using Vec = typename F::Vec;
    using CVec = typename F::CVec;
    constexpr int MK = F::MK;
    info = empty_info();
    info.num_inner = (double)k_inner;
    info.radius = Delta;                                            // :708, :394

    typename F::Step st;
    F::begin_step(ctx, pt, y, mu, st);                              // s, grad f, c  (:724-730)

    Vec dx, Hdx_tcg;
    if constexpr (EXACT) {
        // compute_direction, Exact_RepMat branch (:431-444)
        typename F::Coord cc;
        F::coord_setup(ctx, pt, cc);
        if (!rw.mat_valid) {
            cnt.aux += (double)rep_build<F>(ctx, pt, y, st, cc, rw, rw.cur);
            rw.mat_valid = true;
            rw.al_valid = false;
        }
        if (!rw.al_valid) {
            rep_rhs<F>(ctx, pt, st, cc, rw, rw.cur);
            rw.al_valid = true;
        }
        const dense::TrsOut to = dense::trs_eig(rw.D[rw.cur], rw.al[rw.cur], rw.d, Delta, o.trs_tolhardcase, rw.col, rw.ws,
                                                rw.W, rw.ld);
        dense::cols_dot(rw.VT[rw.cur], rw.d, rw.ld, rw.col, rw.coef);
        dx = F::from_coords(ctx, pt, cc, rw.coef);                  // :441-443
        info.dxtype = (double)to.kind;
    } else {
        const TcgResult tr = F::tcg(ctx, o, pt, y, st, Delta, dx, Hdx_tcg);  // :733 -> :445-452
        cnt.tcg += (double)tr.iters;
        info.dxtype = (double)tr.stop;
        info.tcg_iters = (double)tr.iters;
    }
    constexpr bool MERGED = !EXACT && MergedStepDots<F>::value && TcgReturnsHw<F>::value;
    double dots[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // MERGED: <dx,dx>, <x,dx>, <x+dx,x+dx>, <Hw dx,dx>, <c,dx>
    double normdx;
    CVec ga;
    if constexpr (MERGED) {
        F::step_dots(ctx, pt, st, dx, Hdx_tcg, dots);
        normdx = sqrt(dots[0]);                                     // :735
        ga = F::gadj_given(ctx, pt, dx, dots[1]);
    } else {
        normdx = sqrt(F::inner(ctx, pt, dx, dx));                   // :735
        ga = F::gadj(ctx, pt, dx);
    }
    info.normdx = normdx;

    // dy = -y + mu * (1/s) - y * G*[dx] / s ; yNew = y + dy        (:743, :745)
    CVec yNew;
#pragma unroll
    for (int k = 0; k < MK; ++k) {
        if (F::cactive(ctx, k)) {
            const double dy = (-y.v[k] + mu * (1.0 / pt.s.v[k])) - (y.v[k] * ga.v[k]) / pt.s.v[k];
            yNew.v[k] = y.v[k] + dy;
        } else {
            yNew.v[k] = 0.0;
        }
    }
    typename F::Pt ptN;
    if constexpr (MERGED) F::eval_point(ctx, F::retract_given(ctx, pt, dx, dots[2]), ptN);
    else F::eval_point(ctx, F::retract(ctx, pt, dx), ptN);          // :744 (+ cost, s at xNew)

    // compute_inner_stoppingcriteria (:574-629)
    bool xfe = true, yfe = true;
    double mins = CUDART_INF, miny = CUDART_INF, p_c = 0.0;
#pragma unroll
    for (int k = 0; k < MK; ++k) {
        if (F::cactive(ctx, k)) {
            xfe = xfe && (ptN.s.v[k] > 0.0);                        // :591
            yfe = yfe && (yNew.v[k] > 0.0);                         // :592
            mins = fmin(mins, ptN.s.v[k]);
            miny = fmin(miny, yNew.v[k]);
            const double cv = yNew.v[k] * ptN.s.v[k] - mu;          // :595
            p_c = p_c + cv * cv;
        }
    }
    xfe = wall(xfe) && F::domain_ok(ctx, ptN);
    yfe = wall(yfe);
    double p_xy = F::gradL_xy_partial(ctx, ptN, yNew);
    wsum2(p_c, p_xy);
    const double compl_v = sqrt(p_c);
    const double ngl = F::gradL_norm_given(ctx, ptN, yNew, p_xy);   // :593
    info.minxfeasi = wmin(mins);
    info.minyfeasi = wmin(miny);
    info.compl_v = compl_v;
    bool eig_ok = true;
    if constexpr (EXACT) {
        // :599-617: the matrix of Hw at (xNew, yNew) and its smallest eigenvalue.  Not formed at an infeasible trial point
        // (the reference forms it there too, from negative y/s weights; the value only reaches the log: NaN here)
        if (o.second_order && xfe) {
            typename F::Step stN;
            F::begin_step(ctx, ptN, yNew, mu, stN);
            typename F::Coord cn;
            F::coord_setup(ctx, ptN, cn);
            cnt.aux += (double)rep_build<F>(ctx, ptN, yNew, stN, cn, rw, 1 - rw.cur);
            rep_rhs<F>(ctx, ptN, stN, cn, rw, 1 - rw.cur);
            const double mineig = rep_mineig(rw, 1 - rw.cur);
            info.mineig = mineig;
            eig_ok = mineig >= -tolS;                               // :611
        }
    }

    if (xfe && yfe && ngl <= tolL && compl_v <= tolC && eig_ok) {   // :762-766
        info.inner_status = (double)RIPTRM_INNER_CONVERGED;
        pt = ptN;
        y = yNew;
        if constexpr (EXACT) {
            // Hw does not depend on mu: the decomposition built at (xNew, yNew) serves the next outer iteration, whose
            // inner_preprocess (:415-421) would rebuild the same matrix; only c changes with mu
            if (o.second_order) {
                rw.cur = 1 - rw.cur;
                rw.mat_valid = true;
            } else {
                rw.mat_valid = false;
            }
            rw.al_valid = false;
        }
        return true;
    }
    if (!xfe) {                                                     // :769-775
        info.inner_status = (double)RIPTRM_INNER_PRIMAL_INFEASIBLE;
        Delta = o.gamma * normdx;
        return false;
    }

    // update_xy_TR_radius (:631-705)
    double pl_cur = 0.0, pl_new = 0.0;
#pragma unroll
    for (int k = 0; k < MK; ++k) {
        if (F::cactive(ctx, k)) {
            pl_cur = pl_cur + det_log(pt.s.v[k]);
            pl_new = pl_new + det_log(ptN.s.v[k]);
        }
    }
    wsum2(pl_cur, pl_new);
    const double phi_cur = pt.cost - mu * pl_cur;                   // :649
    const double phi_new = ptN.cost - mu * pl_new;
    double ared = phi_cur - phi_new;                                // :658
    // The Hessian-vector product of :659.  Hw is linear and the tCG accumulates Hw[eta] = sum_j alpha_j Hw[delta_j] beside eta
    // (:132, :154), so on the tCG path that vector is at hand (pymanopt's own trust-region solver uses it the same way): one
    // S.v fewer per trust-region iteration.  It differs from a fresh product by rounding only -- `pred` moves in its last
    // digits, no decision of the reference dataset's 342 trust-region iterations does (profiles/parity_r02.md is unchanged
    // to every digit); DevOpts::recompute_hdx restores the reference's fresh product.
    Vec Hdx;
    bool fresh = true;
    if constexpr (!EXACT && TcgReturnsHw<F>::value) {
        if (!o.recompute_hdx) {
            Hdx = Hdx_tcg;
            fresh = false;
        }
    }
    double ip_hd, ip_cd;
    if (MERGED && !fresh) {
        ip_hd = dots[3];                                            // taken in the merged round right after the tCG
        ip_cd = dots[4];
    } else {
        if (fresh) {
            Hdx = F::Hw(ctx, pt, y, st, dx);
            cnt.aux += 1.0;
        }
        ip_hd = F::inner_partial(ctx, pt, Hdx, dx);
        ip_cd = F::inner_partial(ctx, pt, st.c, dx);
        wsum2(ip_hd, ip_cd);
    }
    double pred = (0.0 - 0.5 * ip_hd) - ip_cd;
    const double reg = (fmax(1.0, fabs(phi_cur)) * 2.220446049250313e-16) * o.reduction_regularization;  // :660
    ared = ared + reg;
    pred = pred + reg;
    info.ared_pred = ared / pred;                                   // :666
    double DeltaNext;
    if (ared < 0.25 * pred) {                                       // :667-675
        info.radius_update = (double)RIPTRM_RADIUS_REDUCED;
        DeltaNext = 0.25 * Delta;
    } else if (ared >= 0.75 * pred && fabs(normdx - Delta) <= 1e-15) {
        info.radius_update = (double)RIPTRM_RADIUS_EXPANDED;
        DeltaNext = fmin(2.0 * Delta, o.maximal_tr_radius);
    } else {
        info.radius_update = (double)RIPTRM_RADIUS_UNCHANGED;
        DeltaNext = Delta;
    }
    if (ared > o.rho * pred) {                                      // :677
        info.inner_status = (double)RIPTRM_INNER_SUCCESSFUL;
        // :681-684.  np.maximum(a, b, out) quirk: I_right = max(const_right, const_right / mu)
        const double I_right = fmax(o.const_right, o.const_right / mu);
        bool clipped_any = false;
#pragma unroll
        for (int k = 0; k < MK; ++k) {
            if (F::cactive(ctx, k)) {
                const double I_left = o.const_left * fmin(fmin(y.v[k], mu / ptN.s.v[k]), 1.0);
                const double cl = fmin(fmax(yNew.v[k], I_left), I_right);
                clipped_any = clipped_any || !(cl == yNew.v[k]);
                yNew.v[k] = cl;
            }
        }
        const bool clipped = wany(clipped_any);
        info.dual_clipping = clipped ? 1.0 : 0.0;                   // :685-695
        pt = ptN;
        y = yNew;
        if constexpr (EXACT) {
            if (!clipped && o.second_order) {                       // :687-692: HwNewmatrix becomes HwCurmatrix
                rw.cur = 1 - rw.cur;
                rw.mat_valid = rw.al_valid = true;
            } else {
                rw.mat_valid = rw.al_valid = false;
            }
        }
    } else {
        info.inner_status = (double)RIPTRM_INNER_UNSUCCESSFUL;      // :697-702
    }
    Delta = DeltaNext;
    return false;
}

// ------------------------------------------------------------------------------------------
// Whole solve: RIPTRM.run (:909-976) with outer_step (:866-896) and inner_run (:785-847)
//
// Pause / resume: a batch is solved in two launches (riptrm_api.cu: the first few outer iterations of every
// pair, then the remainder in order of decreasing work so far, the longest pairs first).  A pair pauses right
// after the loop-top bookkeeping of outer iteration `pause_at` (evaluation done, stop tests passed); its
// state is (x, y) plus the 8 doubles of `PauseState`.  Everything else is a deterministic function of those,
// so a resumed solve is bit-identical to an uninterrupted one.
// ------------------------------------------------------------------------------------------
constexpr int kPauseFields = 8;
struct PauseState {
    double Delta, it, inner, tcg, aux, rows, elapsed_s, paused;
};

template <class F, bool EXACT = false, class Rep = NoRep>
__device__ __forceinline__ void solve_instance(const typename F::Ctx& ctx, const DevOpts& o,
                                               const typename F::Vec& x0, const typename F::CVec& y0,
                                               typename F::Pt& pt, typename F::CVec& y, double* summary,
                                               double* trace /* this instance's rows or nullptr */,
                                               double* pause /* kPauseFields doubles or nullptr */, bool resume,
                                               int pause_at /* outer iteration to pause at; < 0: never */, Rep* rwp = nullptr) {
    using Vec = typename F::Vec;
    using CVec = typename F::CVec;
    F::eval_point(ctx, x0, pt);                                     // outer_preprocess (:849-864)
    y = y0;
    double Delta = o.initial_tr_radius > 0.0 ? o.initial_tr_radius : F::typical_dist(ctx) / 8.0;
    Vec xPrev = pt.x;
    InnerInfo info = empty_info();
    Counters cnt = {0.0, 0.0, 0.0};
    int it = 0, rows = 0, stop_reason = RIPTRM_STOP_RUNNING;
    double mu = o.mu[0];
    double elapsed0 = 0.0;
    if (resume) {
        Delta = pause[0];
        it = (int)pause[1];
        cnt.inner = pause[2];
        cnt.tcg = pause[3];
        cnt.aux = pause[4];
        rows = (int)pause[5];
        elapsed0 = pause[6];
        mu = o.mu[it];
    }
    const uint64_t t_start = global_timer_ns();
    EvalRow ev;
    bool skip_top = resume;
    while (true) {                                                  // :931
        if (!skip_top) {
            ev = evaluate<F>(ctx, pt, y, xPrev, trace != nullptr && o.trace_mode != 0);   // :933
            if (o.trace_mode != 0 && (it == 0 || o.trace_mode == 2)) {  // :936-941
                if (trace != nullptr && rows < o.trace_capacity)
                    write_trace_row(trace + (size_t)rows * RIPTRM_TRACE_FIELDS, it, mu, info, max_abs_mult<F>(ctx, y),
                                    ev, elapsed0 + (double)(global_timer_ns() - t_start) * 1e-9);
                ++rows;
            }
            xPrev = pt.x;                                           // :946
            // base_solver.check_stoppingcriterion (:85-106) + residual criterion (:942-945)
            const double run_time = elapsed0 + (double)(global_timer_ns() - t_start) * 1e-9;
            if (run_time >= o.maxtime) stop_reason = RIPTRM_STOP_MAXTIME;
            else if (it >= o.maxiter) stop_reason = RIPTRM_STOP_MAXITER;
            if (ev.residual <= o.tolresid) stop_reason = RIPTRM_STOP_TOLRESID;
            if (stop_reason != RIPTRM_STOP_RUNNING) break;
            if (it == pause_at) {                                   // hand over to the second launch
                const int l = lane_id();
                const double v = (l == 0) ? Delta : (l == 1) ? (double)it : (l == 2) ? cnt.inner : (l == 3) ? cnt.tcg
                               : (l == 4) ? cnt.aux : (l == 5) ? (double)rows : (l == 6) ? run_time : 1.0;
                if (l < kPauseFields) pause[l] = v;
                return;
            }
        }
        skip_top = false;
        it += 1;                                                    // :959
        // ---- outer_step (:866-896)
        mu = o.mu[it - 1];
        const double tolL = o.tolL[it - 1], tolC = o.tolC[it - 1];  // :881-885
        double tolS = 0.0;
        if constexpr (EXACT) {
            tolS = o.second_order ? o.tolS[it - 1] : 0.0;
            rwp->al_valid = false;                                  // c depends on mu
        }
        // ---- inner_run (:785-847)
        const typename F::Pt pt_init = pt;
        const CVec y_init = y;
        const double Delta_init = Delta;
        Vec xPrevInner = pt.x;
        const uint64_t t_inner = global_timer_ns();
        int k = 0;
        while (true) {
            k += 1;                                                 // :808
            bool done;
            if constexpr (EXACT) {
                done = inner_step<F, true, Rep>(ctx, o, pt, y, mu, Delta, tolL, tolC, k, info, cnt, *rwp, tolS);
            } else {
                NoRep none;
                done = inner_step<F>(ctx, o, pt, y, mu, Delta, tolL, tolC, k, info, cnt, none);  // :810
            }
            cnt.inner += 1.0;
            if (o.trace_mode == 1) {                                // :812-818
                if (trace != nullptr && rows < o.trace_capacity) {
                    const EvalRow evi = evaluate<F>(ctx, pt, y, xPrevInner);
                    write_trace_row(trace + (size_t)rows * RIPTRM_TRACE_FIELDS, it, mu, info,
                                    max_abs_mult<F>(ctx, y), evi,
                                    elapsed0 + (double)(global_timer_ns() - t_start) * 1e-9);
                }
                ++rows;
            }
            xPrevInner = pt.x;                                      // :819
            bool rollback = false;
            const uint64_t now = global_timer_ns();                 // :822-834
            const double rt = (o.inner_maxtime < 0.0) ? elapsed0 + (double)(now - t_start) * 1e-9
                                                      : (double)(now - t_inner) * 1e-9;
            const double lim = (o.inner_maxtime < 0.0) ? o.maxtime : o.inner_maxtime;
            if (rt >= lim) {
                info.inner_status = (double)RIPTRM_INNER_MAX_TIME;
                rollback = true;
            }
            if (o.inner_maxiter >= 0 && k >= o.inner_maxiter) {     // :835-842
                info.inner_status = (double)RIPTRM_INNER_MAX_ITER;
                rollback = true;
            }
            if (rollback) {
                done = true;
                pt = pt_init;
                y = y_init;
                Delta = Delta_init;
                if constexpr (EXACT) rwp->mat_valid = rwp->al_valid = false;
            }
            if (done) break;                                        // :844-845
        }
        mu = o.mu[it];                                              // :890-893 (host-evaluated schedule)
        Delta = fmax(Delta, o.minimal_initial_tr_radius);           // :894
    }
    if (pause != nullptr && lane_id() == 7) pause[7] = 0.0;        // finished, nothing left for a second launch
    if (summary != nullptr) {
        const int l = lane_id();
        double v = 0.0;
        switch (l) {
            case RIPTRM_SM_COST: v = ev.cost; break;
            case RIPTRM_SM_RESIDUAL: v = ev.residual; break;
            case RIPTRM_SM_GRADNORM: v = ev.gradnorm; break;
            case RIPTRM_SM_COMPLVIOLATION: v = ev.compl_v; break;
            case RIPTRM_SM_DUALVIOLATION: v = ev.dual_v; break;
            case RIPTRM_SM_MANVIOLATION: v = ev.man_v; break;
            case RIPTRM_SM_MAXVIOLATION: v = ev.max_v; break;
            case RIPTRM_SM_MEANVIOLATION: v = ev.mean_v; break;
            case RIPTRM_SM_MU: v = mu; break;
            case RIPTRM_SM_RADIUS: v = Delta; break;
            case RIPTRM_SM_OUTER_ITERS: v = (double)it; break;
            case RIPTRM_SM_INNER_ITERS: v = cnt.inner; break;
            case RIPTRM_SM_TCG_ITERS: v = cnt.tcg; break;
            case RIPTRM_SM_AUX_HESSVECS: v = cnt.aux; break;
            case RIPTRM_SM_STOP_REASON: v = (double)stop_reason; break;
            case RIPTRM_SM_TRACE_ROWS: v = (double)rows; break;
            default: break;
        }
        if (l < RIPTRM_SUMMARY_FIELDS) summary[l] = v;
    }
}

}  // namespace riptrm

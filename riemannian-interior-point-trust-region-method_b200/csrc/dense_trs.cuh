// dense_trs.cuh -- the dense linear algebra of the reference's `Exact_RepMat` trust-region solver, one warp per pair:
//
//   * `jacobi_sym`  symmetric eigen-decomposition of the representation matrix (cyclic Jacobi in shared memory) -- replaces
//                   scipy.linalg.eigh in the second-order test (RIPTRM.py:609-611) and feeds the trust-region solve;
//   * `trs_eig`     the trust-region subproblem  min x'Ax/2 + a'x, |x| <= Delta  in the eigenbasis of A -- replaces `TRSgep`
//                   (RIPTRM.py:218-299), which takes the rightmost eigenpair of a 2 dim x 2 dim generalized eigenproblem
//                   (Adachi, Iwata, Nakatsukasa, Takeda 2017) with scipy.linalg.eig.
//
// Why not the pencil: with B = I its rightmost eigenvalue lam1 is the multiplier of the boundary solution, i.e. the root of the
// secular equation  sum_i alpha_i^2 / (d_i + lam)^2 = Delta^2  on (-d_min, inf)  (alpha = V'a, A = V diag(d) V'), and its
// eigenvector is (y1; y2) ~ ((A + lam1)^-1 a; (A + lam1)^-2 a).  One symmetric eigen-decomposition therefore gives everything
// TRSgep takes from the nonsymmetric QZ iteration -- lam1, the boundary point x = -Delta y1/|y1|, the hard-case test
// |y1| / |(y1; y2)| < tolhardcase, the hard-case vector x1 = y2 -- and the same decomposition IS the eigenvalue test of the
// second-order stationarity check and is reused when a step is rejected (only Delta changes).  The interior candidate of
// :244 (scipy.sparse.linalg.cg, rtol 1e-5, at most 10 dim iterations: an INEXACT Newton point, which is what the reference's
// iterates follow) is the same conjugate-gradient recurrence run on the diagonalised system.
//
// All vectors of length d live in shared memory (lanes stride over the entries); every sum is a per-lane serial sum followed
// by the xor butterfly of common.cuh, so results are deterministic.
#pragma once
#include "common.cuh"

namespace riptrm {
namespace dense {

enum { TRS_BOUNDARY = 6, TRS_INTERIOR = 7, TRS_HARD_1 = 8, TRS_HARD_3 = 9, TRS_HARD_6 = 10, TRS_HARD_9 = 11 };

__device__ __forceinline__ double vsum(double p) { return wsum(p); }

// A = V diag(w) V' for the symmetric d x d matrix W (row-major, leading dimension ld, both triangles valid).  On exit the
// diagonal of W holds the eigenvalues (unsorted) and row k of VT the eigenvector of W[k][k].  Rotations are applied to whole
// rows (contiguous, conflict-free) and mirrored into the columns.
static __device__ __noinline__ void jacobi_sym(double* W, double* VT, int d, int ld) {
    const int lane = lane_id();
    for (int e = lane; e < d * ld; e += 32) {
        const int i = e / ld, j = e - i * ld;
        VT[e] = (i == j) ? 1.0 : 0.0;
    }
    __syncwarp();
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int e = lane; e < d * d; e += 32) {
            const int i = e / d, j = e - i * d;
            const double v = W[i * ld + j];
            if (i != j) off = fma(v, v, off);
            else dg = fma(v, v, dg);
        }
        wsum2(off, dg);
        if (!(off == off) || off <= 1e-300 || off <= 1e-34 * dg) break;
        for (int p = 0; p < d - 1; ++p) {
            for (int q = p + 1; q < d; ++q) {
                const double apq = W[p * ld + q], app = W[p * ld + p], aqq = W[q * ld + q];
                __syncwarp();   // every lane has read the pivot block before anyone rewrites it
                // a rotation that cannot change either diagonal entry any more is skipped (uniform: all lanes read the same)
                if (apq == 0.0 || (sweep > 3 && fabs(apq) <= 1e-19 * sqrt(fabs(app) * fabs(aqq)))) continue;
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = lane; k < d; k += 32) {
                    if (k != p && k != q) {
                        const double akp = W[p * ld + k], akq = W[q * ld + k];
                        const double np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
                        W[p * ld + k] = np_;
                        W[q * ld + k] = nq_;
                        W[k * ld + p] = np_;
                        W[k * ld + q] = nq_;
                    }
                    const double vp = VT[p * ld + k], vq = VT[q * ld + k];
                    VT[p * ld + k] = c * vp - s * vq;
                    VT[q * ld + k] = s * vp + c * vq;
                }
                if (lane == 0) {
                    W[p * ld + p] = app - t * apq;
                    W[q * ld + q] = aqq + t * apq;
                    W[p * ld + q] = 0.0;
                    W[q * ld + p] = 0.0;
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();
}

// out[k] = sum_j M[k][j] v[j]  (rows of M = VT: coordinates of v in the eigenbasis); lanes stride over k, ld odd
__device__ __forceinline__ void rows_dot(const double* M, int d, int ld, const double* v, double* out) {
    for (int k = lane_id(); k < d; k += 32) {
        double s = 0.0;
        for (int j = 0; j < d; ++j) s = fma(M[k * ld + j], v[j], s);
        out[k] = s;
    }
    __syncwarp();
}
// out[j] = sum_k M[k][j] z[k]  (back from the eigenbasis)
__device__ __forceinline__ void cols_dot(const double* M, int d, int ld, const double* z, double* out) {
    for (int j = lane_id(); j < d; j += 32) {
        double s = 0.0;
        for (int k = 0; k < d; ++k) s = fma(M[k * ld + j], z[k], s);
        out[j] = s;
    }
    __syncwarp();
}

__device__ __forceinline__ double vdot(const double* a, const double* b, int d) {
    double s = 0.0;
    for (int k = lane_id(); k < d; k += 32) s = fma(a[k], b[k], s);
    return wsum(s);
}

struct TrsOut {
    int kind;        // TRS_*
    double lam1;     // multiplier of the boundary solution (0 for the interior one), as TRSgep returns it
};

// Solves (diag(m) + lam u u') x = b for x by Gaussian elimination with partial pivoting in the scratch matrix H
// (d x (d + 1), leading dimension ldh >= d + 1); the hard-case system of RIPTRM.py:268-270, rare.
static __device__ __noinline__ void solve_diag_plus_rank1(double* H, int ldh, int d, const double* m, double lam, const double* u,
                                                        const double* b, double* x) {
    const int lane = lane_id();
    for (int e = lane; e < d * (d + 1); e += 32) {
        const int i = e / (d + 1), j = e - i * (d + 1);
        H[i * ldh + j] = (j == d) ? b[i] : (lam * u[i] * u[j] + ((i == j) ? m[i] : 0.0));
    }
    __syncwarp();
    for (int c = 0; c < d; ++c) {
        int piv = c;
        double best = fabs(H[c * ldh + c]);
        for (int r = c + 1; r < d; ++r) {      // redundantly on every lane: uniform pivot
            const double v = fabs(H[r * ldh + c]);
            if (v > best) {
                best = v;
                piv = r;
            }
        }
        __syncwarp();
        if (piv != c) {
            for (int j = lane; j <= d; j += 32) {
                const double t = H[c * ldh + j];
                H[c * ldh + j] = H[piv * ldh + j];
                H[piv * ldh + j] = t;
            }
            __syncwarp();
        }
        const double pv = H[c * ldh + c];
        for (int r = c + 1 + lane; r < d; r += 32) {
            const double f = H[r * ldh + c] / pv;
            for (int j = c; j <= d; ++j) H[r * ldh + j] = H[r * ldh + j] - f * H[c * ldh + j];
        }
        __syncwarp();
    }
    if (lane == 0) {
        for (int i = d - 1; i >= 0; --i) {
            double s = H[i * ldh + d];
            for (int j = i + 1; j < d; ++j) s = s - H[i * ldh + j] * x[j];
            x[i] = s / H[i * ldh + i];
        }
    }
    __syncwarp();
}

// The trust-region subproblem in the eigenbasis.  D[k]: eigenvalues, al[k]: coordinates of a; z (out): coordinates of the
// solution; ws: 6 d doubles of scratch; H: scratch matrix for the hard case (may alias nothing live).
static __device__ __noinline__ TrsOut trs_eig(const double* D, const double* al, int d, double Delta, double tolhard, double* z,
                                            double* ws, double* H, int ldh) {
    const int lane = lane_id();
    double* xs = ws;            // CG iterate / later x2
    double* rs = ws + d;        // CG residual
    double* ps = ws + 2 * d;    // CG direction
    double* g = ws + 3 * d;     // gaps d_k - d_min
    double* w = ws + 4 * d;     // y1 ~ al / (g + t)
    double* x1 = ws + 5 * d;    // y2
    TrsOut out;
    const double Delta2 = Delta * Delta;
    // ---- interior candidate: scipy.sparse.linalg.cg(A, -a), x0 = 0, atol = 1e-5 |a|, maxiter = 10 d  (RIPTRM.py:244) -------
    const double bn = sqrt(vdot(al, al, d));
    bool newton_ok = false;
    double newton_obj = 0.0;
    if (bn != 0.0) {
        for (int k = lane; k < d; k += 32) {
            xs[k] = 0.0;
            rs[k] = -al[k];
        }
        __syncwarp();
        const double atol = 1e-5 * bn;
        double rho_prev = 0.0;
        for (int it = 0; it < 10 * d; ++it) {
            double rr = 0.0;
            for (int k = lane; k < d; k += 32) rr = fma(rs[k], rs[k], rr);
            rr = wsum(rr);
            if (sqrt(rr) < atol) break;
            const double beta = (it > 0) ? rr / rho_prev : 0.0;
            double pq = 0.0;
            for (int k = lane; k < d; k += 32) {
                const double p = (it > 0) ? rs[k] + beta * ps[k] : rs[k];
                ps[k] = p;
                pq = fma(p, D[k] * p, pq);
            }
            pq = wsum(pq);
            const double a_cg = rr / pq;
            for (int k = lane; k < d; k += 32) {
                xs[k] = xs[k] + a_cg * ps[k];
                rs[k] = rs[k] - a_cg * (D[k] * ps[k]);
            }
            rho_prev = rr;
            __syncwarp();
        }
        // :245-249: accepted when |A p + a| / |a| < 1e-5 and p'p < Delta^2
        double res = 0.0, pp = 0.0, obj = 0.0;
        for (int k = lane; k < d; k += 32) {
            const double r = D[k] * xs[k] + al[k];
            res = fma(r, r, res);
            pp = fma(xs[k], xs[k], pp);
            obj = obj + (0.5 * (D[k] * xs[k]) * xs[k] + al[k] * xs[k]);
        }
        wsum3(res, pp, obj);
        newton_ok = (sqrt(res) / bn < 1e-5) && !(pp >= Delta2);
        newton_obj = obj;
        if (newton_ok)
            for (int k = lane; k < d; k += 32) z[k] = xs[k];
        __syncwarp();
    }
    // ---- boundary solution: secular equation in t = lam + d_min > 0 --------------------------------------------------------
    double dmin = CUDART_INF;
    for (int k = lane; k < d; k += 32) dmin = fmin(dmin, D[k]);
    dmin = wmin(dmin);
    for (int k = lane; k < d; k += 32) g[k] = D[k] - dmin;
    __syncwarp();
    // psi(t) = sum al^2 / (g + t)^2 decreases from psi(0+) (inf unless a is orthogonal to the bottom eigenspace) to 0
    auto psi = [&](double t, double& s3) {
        double s2 = 0.0, c3 = 0.0;
        for (int k = lane; k < d; k += 32) {
            const double q = al[k] / (g[k] + t);
            s2 = fma(q, q, s2);
            c3 = fma(q * q, 1.0 / (g[k] + t), c3);
        }
        wsum2(s2, c3);
        s3 = c3;
        return s2;
    };
    // 1/sqrt(psi) is concave and increasing in t, so Newton on it (More-Sorensen) started LEFT of the root increases
    // monotonically to it.  A left point is found geometrically below hi = |a| / Delta (psi(hi) <= Delta^2); when psi stays
    // below Delta^2 down to hi 2^-1000 there is no root: a is (numerically) orthogonal to the bottom eigenspace and Delta is
    // beyond the largest step the other eigenvectors can supply -- the hard case proper.
    double t = 0.0;
    bool no_root = false;
    if (!(bn > 0.0)) {
        no_root = true;                // a = 0: the eigenvector of the pencil is (0; v_min)
    } else {
        double hi = bn / Delta, s3 = 0.0;
        double lo = hi;
        int steps = 0;
        while (true) {
            const double s2 = psi(lo, s3);
            if (s2 > Delta2) break;
            hi = lo;
            lo = lo * 0.0625;
            if (++steps > 250) {
                no_root = true;
                break;
            }
        }
        if (!no_root) {
            t = lo;
            for (int it = 0; it < 100; ++it) {
                const double s2 = psi(t, s3);
                if (!(s2 > Delta2)) {           // rounding carried the iterate past the root: it is the root to working precision
                    break;
                }
                double tn = t + (s2 / s3) * ((sqrt(s2) - Delta) / Delta);
                if (!(tn < hi)) tn = 0.5 * (t + hi);
                if (!(tn > t)) break;           // no representable progress
                t = tn;
            }
        }
    }
    // y1 = w = al / (g + t), y2 = w / (g + t), (y1; y2) scaled to unit length as scipy.linalg.eig returns it
    double n1 = 0.0, n2 = 0.0;
    if (no_root) {
        // bottom eigenvector: the lowest index attaining d_min
        int kmin = 1 << 30;
        for (int k = lane; k < d; k += 32)
            if (g[k] == 0.0) kmin = min(kmin, k);
        for (int off = 16; off > 0; off >>= 1) kmin = min(kmin, __shfl_xor_sync(kFull, kmin, off));
        for (int k = lane; k < d; k += 32) {
            w[k] = 0.0;
            x1[k] = (k == kmin) ? 1.0 : 0.0;
        }
        n1 = 0.0;
        n2 = 1.0;
    } else {
        for (int k = lane; k < d; k += 32) {
            const double q = al[k] / (g[k] + t);
            w[k] = q;
            x1[k] = q / (g[k] + t);
            n1 = fma(q, q, n1);
            n2 = fma(x1[k], x1[k], n2);
        }
        wsum2(n1, n2);
    }
    __syncwarp();
    const double lam1 = t - dmin;
    const double scale = 1.0 / sqrt(n1 + n2);
    const double normx = sqrt(n1) * scale;                                 // :258 with (y1; y2) of unit length
    int kind = TRS_BOUNDARY;
    double bobj = 0.0;
    if (!(normx < tolhard)) {
        // x = -Delta y1 / |y1|  (:259-261: scaled to the boundary, sign such that x'a <= 0)
        const double f = -Delta / sqrt(n1);
        for (int k = lane; k < d; k += 32) {
            const double xv = f * w[k];
            xs[k] = xv;
            bobj = bobj + (0.5 * (D[k] * xv) * xv + al[k] * xv);
        }
        bobj = wsum(bobj);
    } else {
        // ---- hard case (:263-290) ------------------------------------------------------------------------------------------
        for (int k = lane; k < d; k += 32) {
            x1[k] = x1[k] * scale;
            rs[k] = g[k] + t;           // diagonal of A + lam1 I in the eigenbasis
            ps[k] = -al[k];
        }
        __syncwarp();
        solve_diag_plus_rank1(H, ldh, d, rs, lam1, x1, ps, xs);            // H = A + lam1 I + lam1 x1 x1' (:268-270)
        kind = TRS_HARD_1;
        auto resid = [&]() {
            double r2 = 0.0;
            for (int k = lane; k < d; k += 32) {
                const double r = rs[k] * xs[k] + al[k];
                r2 = fma(r, r, r2);
            }
            return sqrt(wsum(r2)) / bn;
        };
        if (resid() > tolhard) {                                           // :274-283
            for (int ii = 3; ii <= 9; ii += 3) {
                // the ii smallest eigenvalues get lam1 added on the diagonal (P = their eigenvectors: H is diagonal here)
                for (int k = lane; k < d; k += 32) {
                    int rank = 0;
                    for (int j = 0; j < d; ++j) rank += (D[j] < D[k] || (D[j] == D[k] && j < k)) ? 1 : 0;
                    xs[k] = -al[k] / (rs[k] + ((rank < ii) ? lam1 : 0.0));
                }
                __syncwarp();
                kind = (ii == 3) ? TRS_HARD_3 : (ii == 6) ? TRS_HARD_6 : TRS_HARD_9;
                if (resid() < tolhard) break;
            }
        }
        double aa = 0.0, bb = 0.0, cc = 0.0;
        for (int k = lane; k < d; k += 32) {
            aa = fma(x1[k], x1[k], aa);
            bb = fma(xs[k], x1[k], bb);
            cc = fma(xs[k], xs[k], cc);
        }
        wsum3(aa, bb, cc);
        bb = 2.0 * bb;
        cc = cc - Delta2;
        const double alp = (-bb + sqrt(bb * bb - 4.0 * aa * cc)) / (2.0 * aa);   // :289
        for (int k = lane; k < d; k += 32) {
            const double xv = xs[k] + alp * x1[k];
            xs[k] = xv;
            bobj = bobj + (0.5 * (D[k] * xv) * xv + al[k] * xv);
        }
        bobj = wsum(bobj);
    }
    __syncwarp();
    // ---- :293-299: the interior candidate wins when its model value is not larger ------------------------------------------
    if (newton_ok && newton_obj <= bobj) {
        out.kind = TRS_INTERIOR;
        out.lam1 = 0.0;
    } else {
        for (int k = lane; k < d; k += 32) z[k] = xs[k];
        out.kind = kind;
        out.lam1 = lam1;
    }
    __syncwarp();
    return out;
}

}  // namespace dense
}  // namespace riptrm

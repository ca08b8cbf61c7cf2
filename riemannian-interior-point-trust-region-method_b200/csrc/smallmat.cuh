// smallmat.cuh -- tiny dense linear algebra for the small-manifold families (Grassmann(5,3), Skew/SPD(5)):
// matrices live row-major in a warp's shared-memory scratch, the 32 lanes split the output entries, every
// entry is a serial (fixed-order) sum, and every routine ends with __syncwarp().  Sizes are <= 5 x 5 here:
// latency, not throughput, is what matters and determinism is free.
// The routines are NOT inlined: the whole-solve kernels call them from hundreds of sites, and with everything inlined
// the StableIdentification kernel was 1 MB of SASS whose warps, each at a different place, waited on instruction
// fetch 83 % of the time (profiles/r01k_stableid_*).
#pragma once
#include "common.cuh"

namespace riptrm {
namespace sm {

// C[m x n] = op(A) * op(B); op(A) is m x k (A stored k x m when tA), op(B) is k x n (B stored n x k when tB)
static __device__ __noinline__ void mm(double* C, const double* A, const double* B, int m, int k, int n, bool tA = false,
                                   bool tB = false) {
    for (int e = lane_id(); e < m * n; e += 32) {
        const int i = e / n, j = e - i * n;
        double s = 0.0;
        for (int l = 0; l < k; ++l) {
            const double a = tA ? A[l * m + i] : A[i * k + l];
            const double b = tB ? B[j * k + l] : B[l * n + j];
            s = fma(a, b, s);
        }
        C[e] = s;
    }
    __syncwarp();
}

// C = alpha * A + beta * B (elementwise, len entries); C may alias A or B
__device__ __forceinline__ void axpby(double* C, double alpha, const double* A, double beta, const double* B, int len) {
    for (int e = lane_id(); e < len; e += 32) C[e] = alpha * A[e] + beta * B[e];
    __syncwarp();
}

// C = (A + A') / 2 or (A - A') / 2 for an n x n matrix; C must not alias A
__device__ __forceinline__ void sym(double* C, const double* A, int n, double sign = 1.0) {
    for (int e = lane_id(); e < n * n; e += 32) {
        const int i = e / n, j = e - i * n;
        C[e] = 0.5 * (A[e] + sign * A[j * n + i]);
    }
    __syncwarp();
}

// Cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 5), executed redundantly by every lane on
// private copies: A = V diag(w) V'.  Returns false if a non-finite entry appears.
template <int NMAX>
__device__ __noinline__ bool jacobi_eig(const double* Ain, int n, double (&w)[NMAX], double (&V)[NMAX][NMAX]) {
    double A[NMAX][NMAX];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            A[i][j] = Ain[i * n + j];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                if (i != j) off += A[i][j] * A[i][j];
                else diag += A[i][j] * A[i][j];
            }
        if (!(off == off)) return false;
        if (off <= 1e-60 || off <= 1e-34 * diag) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = A[i][i];
    return true;
}

// In-place inverse of an n x n matrix (n <= 5) by Gauss-Jordan with partial pivoting, redundantly per lane on a
// private copy, result written back by lane 0.  Returns false when singular / non-finite.
template <int NMAX>
__device__ __noinline__ bool inverse(double* Ainv, const double* Ain, int n) {
    double M[NMAX][2 * NMAX];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            M[i][j] = Ain[i * n + j];
            M[i][n + j] = (i == j) ? 1.0 : 0.0;
        }
    bool ok = true;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        double best = fabs(M[c][c]);
        for (int r = c + 1; r < n; ++r)
            if (fabs(M[r][c]) > best) {
                best = fabs(M[r][c]);
                piv = r;
            }
        if (!(best > 0.0)) {
            ok = false;
            break;
        }
        if (piv != c)
            for (int j = 0; j < 2 * n; ++j) {
                const double t = M[c][j];
                M[c][j] = M[piv][j];
                M[piv][j] = t;
            }
        const double d = M[c][c];
        for (int j = 0; j < 2 * n; ++j) M[c][j] = M[c][j] / d;
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = M[r][c];
            if (f != 0.0)
                for (int j = 0; j < 2 * n; ++j) M[r][j] = M[r][j] - f * M[c][j];
        }
    }
    __syncwarp();
    if (lane_id() == 0 && ok)
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) Ainv[i * n + j] = M[i][n + j];
    __syncwarp();
    return ok;
}

// true iff the symmetric n x n matrix is positive definite (Cholesky succeeds), redundantly per lane
template <int NMAX>
__device__ __noinline__ bool is_spd(const double* Ain, int n) {
    double L[NMAX][NMAX];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = Ain[i * n + j];
            for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
            if (i == j) {
                if (!(s > 0.0)) return false;
                L[i][i] = sqrt(s);
            } else {
                L[i][j] = s / L[j][j];
            }
        }
    return true;
}

}  // namespace sm
}  // namespace riptrm

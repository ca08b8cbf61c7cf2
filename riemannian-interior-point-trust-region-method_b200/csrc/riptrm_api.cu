// riptrm_api.cu -- C-ABI entry points (include/riptrm_b200.h) and the warp-per-instance kernels.
//
// Kernel shape for the batched small families: ONE warp per CTA, one (instance, initialpoint)
// pair per warp at a time, a persistent grid (SM count x resident CTAs per SM) that pulls
// instances from an atomic queue because per-instance work varies 3-5x.
// There is no inter-warp synchronisation anywhere on the solve path.
#include <cuda.h>           // types of the stream memory operations only: the entry point comes from cudaGetDriverEntryPoint
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/riptrm_b200.h"
#include "datagen.cuh"
#include "fam_columns.cuh"
#include "fam_stiefel.cuh"
#include "fam_grassmann.cuh"
#include "fam_sphere.cuh"
#include "fam_stableid.cuh"
#include "peaks.cuh"

using namespace riptrm;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                               \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return fail(RIPTRM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    } while (0)

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
struct riptrm_handle {
    int family = 0, n = 0, p = 0, m = 0, batch = 0, device = 0;
    int vec_len = 0;  // n*p doubles per point
    int num_sms = 0;
    // problem data
    double* dZ = nullptr;  // [batch_z][n][n]
    bool ownZ = false;
    size_t z_bytes = 0;
    int batch_z = 0;
    double eps = 0.0;
    double ros_alpha = 0.0, ros_offset = 0.0;
    double* d_sid = nullptr;  // StableIdentification: X [d][N], XP [d][N], conspec [m][5]
    int sid_N = 0;
    double sid_h = 0.0;
    bool have_problem = false;
    // options
    riptrm_options opts{};
    bool have_opts = false;
    double* d_sched = nullptr;  // 3 * (maxiter + 1)
    int sched_len = 0;
    // staging for RIPTRM_HOST calls
    double *d_x0 = nullptr, *d_y0 = nullptr, *d_x = nullptr, *d_y = nullptr, *d_summary = nullptr, *d_trace = nullptr;
    double* d_v = nullptr;      // hook operand
    double* d_info = nullptr;   // hook info
    size_t trace_bytes = 0;
    int* d_counter = nullptr;   // [2]: work queue of the main kernel, work queue of the fast lane
    int* d_fast_order = nullptr;  // fast lane: the pairs of the longest units
    int* d_lane_state = nullptr;  // fast lane: election state (2 + 2 * 256 ints) + per-CTA placement records
    unsigned long long* d_lane_times = nullptr;   // entry / exit globaltimer of the recorded CTAs
    unsigned int* d_lane_arrive = nullptr;        // arrival count of the lane kernel's CTAs (see SphereParams::lane_arrive)
    int lane_debug_len = 0;
    cudaStream_t lane_stream = nullptr;
    cudaEvent_t lane_ev0 = nullptr, lane_ev1 = nullptr;
    // two-launch schedule of the batched families (longest pairs first in the second launch)
    double* d_pause = nullptr;   // [batch][kPauseFields]
    float *d_keys = nullptr, *d_keys_sorted = nullptr;
    int *d_idx = nullptr, *d_order = nullptr;
    void* d_sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    // COLUMNS family workspace
    double* dS = nullptr;        // [n_pad][ld] = Z + Z'
    double* d_colbuf = nullptr;  // 15 arrays of n_pad x P + partials
    unsigned long long* d_passes = nullptr;
    int colP = 0, n_pad = 0, ld = 0, col_slots = 0, col_R = 0, col_grid = 0;
    unsigned long long last_passes = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // COLUMNS / STIEFEL whole solves: the host runs kColLookahead trust-region iterations ahead of the flag it polls
    int* h_done = nullptr;               // pinned, kColLookahead ints
    cudaEvent_t done_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool no_launch_events = false;       // inside a whole solve: the launches do not record ev0 / ev1 (the solve does)
    // device-side loop of the whole solve (columns_solve_graph): the instantiated graph, the parameter block it was built
    // for, the device clock stamp
    cudaGraphExec_t col_graph = nullptr;
    col::Params col_graph_prm{};
    double* d_now = nullptr;             // [0] seconds since the start of the solve, [1] (as bits) globaltimer at the start
    int64_t launches = 0;
    double last_ms = 0.0;
};

struct SphereParams {
    const double* Z;  // device [batch_z][n][n]
    int batch_z;
    int n;
    int batch;
    double eps;
    const double* x0;
    const double* y0;
    double* x;
    double* y;
    double* summary;
    double* trace;
    // two-launch schedule
    const int* order;  // slot -> pair (second launch) or nullptr
    double* pause;     // [batch][kPauseFields] or nullptr
    int resume;        // second launch: continue the paused pairs from (x, y, pause)
    int pause_at;      // first launch: outer iteration to pause at (< 0: run to the end)
    int sibling_units; // 1: the queue (and `order`) counts units of two consecutive pairs that share a Z (sphere_tmem2_kernel)
    int queue_len;     // entries of the work queue (0: batch, or batch / 2 units)
    // fast lane inside sphere_tmem2_kernel: `lane_ctas` CTAs, each alone on its SM, solve the pairs of `fast_order` one warp
    // per scheduler; lane_state = {lanes taken, CTAs arrived, arrivals per SM [256], role per SM [256]}; lane_debug [grid]
    int lane_ctas, fast_len;
    const int* fast_order;
    int* lane_state;
    int* lane_debug;
    unsigned long long* lane_times;   // [CTA][2]: globaltimer at entry / exit of the CTAs that write a lane_debug record
    unsigned int* lane_arrive;        // two-kernel lane: every lane CTA counts itself in when it starts; the main kernel's stream
                                      // waits on the count (cuStreamWaitValue32), so it cannot start before the lane is resident
    // hooks
    const double* v;
    double mu;
    double Delta;
    double* out;
    double* info;
    // riptrm_newton: explicit slacks [batch][m], solver (0 RepMat, 1 conjugate residuals), CR tolerance / iteration cap
    const double* slack;
    int newton_method, kr_maxiter;
    double kr_tol;
};

enum { kLaneMain = 1, kLaneLane = 2, kLaneExit = 3 };   // roles in the fast-lane placement record (riptrm_lane_placement)
constexpr int kLaneMaxSms = 256;

// Loads Z of one instance into shared memory and symmetrises it in place: S = Z + Z'
// (Z is not symmetric: src/NonnegPCA/generator.py:25-28; Hessian of -x'Zx is -(Z+Z')).
__device__ __forceinline__ void load_S(const double* __restrict__ Zg, double* S, int n, int ns, int pad) {
    const int lane = lane_id();
    const int nn = n * n;
    for (int idx = lane; idx < nn; idx += 32) {
        const int i = idx / n, j = idx - i * n;
        S[i * ns + j] = __ldg(Zg + idx);
    }
    if (ns != n)
        for (int i = lane; i < n; i += 32) S[i * ns + n] = 0.0;
    for (int i = lane; i < pad; i += 32) S[n * ns + i] = 0.0;
    __syncwarp();
    for (int idx = lane; idx < nn; idx += 32) {
        const int i = idx / n, j = idx - i * n;
        if (i <= j) {
            const double s = S[i * ns + j] + S[j * ns + i];
            S[i * ns + j] = s;
            S[j * ns + i] = s;
        }
    }
    __syncwarp();
}

// global [len] <-> pair layout (element 64*(k>>1) + 2*lane + (k&1))
template <int K>
__device__ __forceinline__ WVec<K> load_vec(const double* __restrict__ g, int len) {
    WVec<K> r;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int e = SphereFam<K, 0>::elem(k);
        r.v[k] = (e < len) ? g[e] : 0.0;
    }
    return r;
}
template <int K>
__device__ __forceinline__ void store_vec(double* g, const WVec<K>& r, int len) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int e = SphereFam<K, 0>::elem(k);
        if (e < len) g[e] = r.v[k];
    }
}

// One exact trust-region solve at (pt, y0, mu, Delta) for the riptrm_trs hook: dx and {type, lam1, ||dx||, smallest eigenvalue}
template <class F>
__device__ __forceinline__ void trs_hook(const typename F::Ctx& ctx, const DevOpts& o, const typename F::Pt& pt,
                                         const typename F::CVec& y0, const typename F::Step& st, double Delta, RepWork& rw,
                                         typename F::Vec& dx, double* info4) {
    typename F::Coord cc;
    F::coord_setup(ctx, pt, cc);
    rep_build<F>(ctx, pt, y0, st, cc, rw, 0);
    rep_rhs<F>(ctx, pt, st, cc, rw, 0);
    const dense::TrsOut to = dense::trs_eig(rw.D[0], rw.al[0], rw.d, Delta, o.trs_tolhardcase, rw.col, rw.ws, rw.W, rw.ld);
    dense::cols_dot(rw.VT[0], rw.d, rw.ld, rw.col, rw.coef);
    dx = F::from_coords(ctx, pt, cc, rw.coef);
    const double nrm = sqrt(F::inner(ctx, pt, dx, dx));
    const double mineig = rep_mineig(rw, 0);
    const int lane = lane_id();
    if (info4 != nullptr && lane < 4)
        info4[lane] = (lane == 0) ? (double)to.kind : (lane == 1) ? to.lam1 : (lane == 2) ? nrm : mineig;
}

// riptrm_newton: st.ys <- z / s with the caller's slacks, the right-hand side made tangent, then one of the two solvers
template <class F>
__device__ __forceinline__ void newton_hook(const typename F::Ctx& ctx, const typename F::Pt& pt, const typename F::CVec& z,
                                            typename F::Step& st, const typename F::CVec& sl, const typename F::Vec& rhs, int method,
                                            double tol, int maxiter, RepWork& rw, typename F::Vec& dx, double* info4) {
#pragma unroll
    for (int k = 0; k < F::MK; ++k) st.ys.v[k] = F::cactive(ctx, k) ? z.v[k] / sl.v[k] : 0.0;
    const typename F::Vec c = F::project(ctx, pt, rhs);
    const NewtonOut no = (method == 0) ? newton_repmat<F>(ctx, pt, z, st, c, rw, dx) : newton_cr<F>(ctx, pt, z, st, c, tol, maxiter, dx);
    const double nrm = sqrt(F::inner(ctx, pt, dx, dx));
    const int lane = lane_id();
    if (info4 != nullptr && lane < 4)
        info4[lane] = (lane == 0) ? no.iters : (lane == 1) ? no.rel_res : (lane == 2) ? nrm : no.mineig;
}

// mode 0: whole solve; 1: one Hessian-vector product; 2: one tCG solve; 3: one exact trust-region solve (EXACT only);
// 4: one condensed Newton-system solve of the interior-point method (EXACT only: it shares the RepWork workspace)
// EXACT: TRS_solver='Exact_RepMat' -- the representation-matrix workspace (RepWork) follows S in shared memory
template <int K, int MODE, int NFIX, bool EXACT = false>
__global__ void __launch_bounds__(32, EXACT ? 1 : ((K == 2) ? 10 : 4)) sphere_kernel(SphereParams P, DevOpts o, int* counter) {
    using F = SphereFam<K, NFIX>;
    extern __shared__ __align__(16) double smem[];
    const int n = P.n;
    const int ns = (n + 1) & ~1;
    const int pad = 32 * K;
    typename F::Ctx ctx;
    ctx.S = smem;
    ctx.vbuf = smem + n * ns + pad;
    ctx.n = n;
    ctx.ns = ns;
    ctx.eps = P.eps;
    ctx.embedded = o.is_euclidean_embedded != 0;
    const int lane = lane_id();
    int loaded_z = -1;
    RepWork rw;
    if (EXACT) rep_init(rw, smem + n * ns + pad + 32 * K, n - 1);
    while (true) {
        int inst = 0;
        if (lane == 0) inst = atomicAdd(counter, 1);
        inst = __shfl_sync(kFull, inst, 0);
        if (inst >= P.batch) break;
        if (P.order != nullptr) inst = P.order[inst];
        double* pause = (P.pause != nullptr) ? P.pause + (size_t)inst * kPauseFields : nullptr;
        if (MODE == 0 && P.resume && pause[7] == 0.0) continue;  // finished in the first launch
        const int zi = inst / (P.batch / P.batch_z);  // consecutive pairs of one instance share its Z
        if (zi != loaded_z) {
            load_S(P.Z + (size_t)zi * n * n, smem, n, ns, pad);
            loaded_z = zi;
        }
        const bool resume = (MODE == 0) && P.resume;
        const typename F::Vec x0 = load_vec<K>((resume ? P.x : P.x0) + (size_t)inst * n, n);
        const typename F::CVec y0 = load_vec<K>((resume ? P.y : P.y0) + (size_t)inst * n, n);
        if (MODE == 0) {
            typename F::Pt pt;
            typename F::CVec y;
            double* tr = (P.trace != nullptr && o.trace_mode != 0)
                             ? P.trace + (size_t)inst * o.trace_capacity * RIPTRM_TRACE_FIELDS
                             : nullptr;
            if constexpr (EXACT) {
                rw.cur = 0;
                rw.mat_valid = rw.al_valid = false;
                solve_instance<F, true, RepWork>(ctx, o, x0, y0, pt, y,
                                                 P.summary ? P.summary + (size_t)inst * RIPTRM_SUMMARY_FIELDS : nullptr, tr, nullptr,
                                                 false, -1, &rw);
            } else {
                solve_instance<F>(ctx, o, x0, y0, pt, y, P.summary ? P.summary + (size_t)inst * RIPTRM_SUMMARY_FIELDS : nullptr,
                                  tr, pause, resume, P.pause_at);
            }
            if (P.x) store_vec<K>(P.x + (size_t)inst * n, pt.x, n);
            if (P.y) store_vec<K>(P.y + (size_t)inst * n, y, n);
        } else {
            typename F::Pt pt;
            F::eval_point(ctx, x0, pt);
            typename F::Step st;
            F::begin_step(ctx, pt, y0, P.mu, st);
            if (MODE == 1) {
                const typename F::Vec v = load_vec<K>(P.v + (size_t)inst * n, n);
                const typename F::Vec hv = F::Hw(ctx, pt, y0, st, v);
                store_vec<K>(P.out + (size_t)inst * n, hv, n);
            } else if (MODE == 3) {
                if constexpr (EXACT) {
                    typename F::Vec dx;
                    trs_hook<F>(ctx, o, pt, y0, st, P.Delta, rw, dx, P.info ? P.info + (size_t)inst * 4 : nullptr);
                    store_vec<K>(P.out + (size_t)inst * n, dx, n);
                }
            } else if (MODE == 4) {
                if constexpr (EXACT) {
                    typename F::Vec dx;
                    const typename F::CVec sl = load_vec<K>(P.slack + (size_t)inst * n, n);
                    const typename F::Vec rhs = load_vec<K>(P.v + (size_t)inst * n, n);
                    newton_hook<F>(ctx, pt, y0, st, sl, rhs, P.newton_method, P.kr_tol, P.kr_maxiter, rw, dx,
                                   P.info ? P.info + (size_t)inst * 4 : nullptr);
                    store_vec<K>(P.out + (size_t)inst * n, dx, n);
                }
            } else {
                typename F::Vec eta, Heta;
                const TcgResult r = F::tcg(ctx, o, pt, y0, st, P.Delta, eta, Heta);
                store_vec<K>(P.out + (size_t)inst * n, eta, n);
                const double nrm = sqrt(F::inner(ctx, pt, eta, eta));  // warp-collective: all lanes
                if (P.info != nullptr && lane < 4) {
                    const double val = (lane == 0) ? (double)r.iters : (lane == 1) ? (double)r.stop : (lane == 2) ? nrm : r.model_value;
                    P.info[(size_t)inst * 4 + lane] = val;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
extern "C" int riptrm_abi_version(void) { return RIPTRM_ABI_VERSION; }
extern "C" const char* riptrm_last_error(void) { return g_last_error.c_str(); }

extern "C" int riptrm_create(int family, int n, int p, int m, int batch, int device, riptrm_handle** out) {
    if (out == nullptr) return fail(RIPTRM_E_INVALID, "out is NULL");
    *out = nullptr;
    if (n <= 0 || p <= 0 || m < 0 || batch <= 0) return fail(RIPTRM_E_INVALID, "n, p, batch must be positive");
    if (family == RIPTRM_FAMILY_NONNEGPCA_SPHERE) {
        if (p != 1 || m != n) return fail(RIPTRM_E_INVALID, "NonnegPCA/Sphere needs p == 1 and m == n");
        if (n > 128) return fail(RIPTRM_E_UNSUPPORTED, "Sphere family: n <= 128 (use RIPTRM_FAMILY_NONNEGPCA_COLUMNS for large n)");
    } else if (family == RIPTRM_FAMILY_ROSENBROCK_GRASSMANN) {
        if (n * p > 32 || p > GrassmannFam::PMAX || p > n) return fail(RIPTRM_E_UNSUPPORTED, "Grassmann family: n*p <= 32, p <= 5");
        if (m != n * p) return fail(RIPTRM_E_INVALID, "Rosenbrock/Grassmann needs m == n * p");
    } else if (family == RIPTRM_FAMILY_STABLEID_PRODUCT) {
        if (p != 3 || n > StableIdFam::DMAX || m > 32) return fail(RIPTRM_E_UNSUPPORTED, "StableIdentification family: p == 3 blocks, d <= 5, m <= 32");
    } else if (family == RIPTRM_FAMILY_NONNEGPCA_COLUMNS) {
        if (batch != 1) return fail(RIPTRM_E_INVALID, "COLUMNS family: batch must be 1 (the p columns are the batch)");
        if (m != n * p) return fail(RIPTRM_E_INVALID, "COLUMNS family needs m == n * p");
        if (p > col::MAXP) return fail(RIPTRM_E_UNSUPPORTED, "COLUMNS family: p <= 16");
    } else if (family == RIPTRM_FAMILY_NONNEGPCA_STIEFEL) {
        if (batch != 1) return fail(RIPTRM_E_INVALID, "STIEFEL family: batch must be 1 (one n x p matrix iterate)");
        if (m != n * p) return fail(RIPTRM_E_INVALID, "STIEFEL family needs m == n * p");
        if (p > col::MAXP || p > n) return fail(RIPTRM_E_UNSUPPORTED, "STIEFEL family: p <= 16, p <= n");
    } else {
        return fail(RIPTRM_E_UNSUPPORTED, "family not built into this library");
    }
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(RIPTRM_E_INVALID, "no such CUDA device");
    CUDA_TRY(cudaSetDevice(device));
    riptrm_handle* h = new riptrm_handle();
    h->family = family;
    h->n = n;
    h->p = p;
    h->m = m;
    h->batch = batch;
    h->device = device;
    h->vec_len = (family == RIPTRM_FAMILY_STABLEID_PRODUCT) ? 3 * n * n : n * p;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaMalloc(&h->d_counter, 2 * sizeof(int)));
    CUDA_TRY(cudaEventCreate(&h->ev0));
    CUDA_TRY(cudaEventCreate(&h->ev1));
    *out = h;
    return RIPTRM_OK;
}

static void free_dev(double*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}
template <class T>
static void free_any(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

extern "C" int riptrm_destroy(riptrm_handle* h) {
    if (h != nullptr) {
        if (h->col_graph != nullptr) cudaGraphExecDestroy(h->col_graph);
        free_dev(h->d_now);
        if (h->h_done != nullptr) cudaFreeHost(h->h_done);
        for (auto& e : h->done_ev) if (e != nullptr) cudaEventDestroy(e);
    }
    if (h == nullptr) return RIPTRM_OK;
    cudaSetDevice(h->device);
    if (h->ownZ) free_dev(h->dZ);
    free_dev(h->d_sched);
    free_dev(h->d_x0);
    free_dev(h->d_y0);
    free_dev(h->d_x);
    free_dev(h->d_y);
    free_dev(h->d_summary);
    free_dev(h->d_trace);
    free_dev(h->d_v);
    free_dev(h->d_info);
    free_dev(h->dS);
    free_dev(h->d_colbuf);
    free_dev(h->d_pause);
    free_dev(h->d_sid);
    free_any(h->d_keys);
    free_any(h->d_keys_sorted);
    free_any(h->d_idx);
    free_any(h->d_order);
    free_any(h->d_sort_tmp);
    free_any(h->d_passes);
    free_any(h->d_counter);
    free_any(h->d_fast_order);
    free_any(h->d_lane_state);
    free_any(h->d_lane_times);
    free_any(h->d_lane_arrive);
    if (h->lane_stream) cudaStreamDestroy(h->lane_stream);
    if (h->lane_ev0) cudaEventDestroy(h->lane_ev0);
    if (h->lane_ev1) cudaEventDestroy(h->lane_ev1);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
    return RIPTRM_OK;
}


// ------------------------------------------------------------------------------------------
// COLUMNS family (fam_columns.cuh): host side
// ------------------------------------------------------------------------------------------
static int finish_timing(riptrm_handle* h, bool sync);
static int ensure(double*& p, size_t bytes);

constexpr int kColArrays = 24;  // n_pad x P work arrays of the COLUMNS family

static int columns_template_p(int p) {
    const int sizes[] = {1, 2, 4, 8, 10, 16};
    for (int s : sizes)
        if (p <= s) return s;
    return -1;
}

static bool is_stiefel(const riptrm_handle* h) { return h->family == RIPTRM_FAMILY_NONNEGPCA_STIEFEL; }
// the Stiefel kernels are built for P = 4, 10, 16
static int stiefel_template_p(int p) { return p <= 4 ? 4 : (p <= 10 ? 10 : 16); }

static int columns_setup(riptrm_handle* h, const double* Z, double eps, int where) {
    const int n = h->n;
    h->colP = is_stiefel(h) ? stiefel_template_p(h->p) : columns_template_p(h->p);
    h->n_pad = (n + col::TJ - 1) / col::TJ * col::TJ;
    h->ld = (n + col::TW - 1) / col::TW * col::TW;
    h->col_grid = h->num_sms;
    const long long NIB = h->ld / col::TW, NJT = h->n_pad / col::TJ, total = NIB * NJT;
    const long long per_cta = (total + h->col_grid - 1) / h->col_grid;
    h->col_slots = (int)((per_cta + NJT - 2) / NJT + 1);
    h->col_R = (n + h->col_grid - 1) / h->col_grid;
    const size_t sbytes = (size_t)h->n_pad * h->ld * sizeof(double);
    free_dev(h->dS);
    free_dev(h->d_colbuf);
    free_dev(h->d_pause);
    free_any(h->d_keys);
    free_any(h->d_keys_sorted);
    free_any(h->d_idx);
    free_any(h->d_order);
    free_any(h->d_sort_tmp);
    CUDA_TRY(cudaMalloc(&h->dS, sbytes));
    CUDA_TRY(cudaMemset(h->dS, 0, sbytes));
    const double* dZ = Z;
    double* tmpZ = nullptr;
    if (where != RIPTRM_DEVICE) {
        CUDA_TRY(cudaMalloc(&tmpZ, (size_t)n * n * sizeof(double)));
        const cudaError_t ce = cudaMemcpy(tmpZ, Z, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) {
            cudaFree(tmpZ);
            return fail(RIPTRM_E_CUDA, std::string("cudaMemcpy(Z): ") + cudaGetErrorString(ce));
        }
        dZ = tmpZ;
    }
    dim3 grid((n + 31) / 32, (n + 31) / 32), block(32, 8);
    col::build_S_kernel<<<grid, block>>>(dZ, h->dS, n, h->n_pad / col::TJ, h->colP >= 8 ? 1 : 0);
    cudaError_t be = cudaGetLastError();
    if (be == cudaSuccess) be = cudaDeviceSynchronize();
    if (tmpZ) cudaFree(tmpZ);
    if (be != cudaSuccess) return fail(RIPTRM_E_CUDA, std::string("build_S_kernel: ") + cudaGetErrorString(be));
    h->launches += 1;
    const size_t arr = (size_t)h->n_pad * h->colP;
    const size_t mv = (size_t)h->col_grid * h->col_slots * col::TW * h->colP;
    const size_t dots = (size_t)2 * h->col_grid * stf::DOT_STRIDE;
    const size_t total_d = kColArrays * arr + mv + dots + 4 * col::MAXP + (size_t)2 * (h->col_grid + 2) +
                           (size_t)col::MAXP * col::CS_FIELDS + (size_t)col::MAXP * RIPTRM_SUMMARY_FIELDS + 2;
    CUDA_TRY(cudaMalloc(&h->d_colbuf, total_d * sizeof(double)));
    CUDA_TRY(cudaMemset(h->d_colbuf, 0, total_d * sizeof(double)));
    {
        // stream-K schedule: CTA g streams tiles [tbeg[g], tbeg[g+1]) of the linearised tile order
        std::vector<long long> tb(h->col_grid + 1);
        std::vector<int> ib0(h->col_grid + 2, 0);
        for (int g = 0; g <= h->col_grid; ++g) tb[g] = total * g / h->col_grid;
        for (int g = 0; g < h->col_grid; ++g) ib0[g] = (int)(tb[g] / NJT);
        double* tail = h->d_colbuf + kColArrays * arr + mv + dots + 4 * col::MAXP;
        CUDA_TRY(cudaMemcpy(tail, tb.data(), tb.size() * sizeof(long long), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(tail + (h->col_grid + 2), ib0.data(), ib0.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (h->d_passes == nullptr) CUDA_TRY(cudaMalloc(&h->d_passes, sizeof(unsigned long long)));
    CUDA_TRY(cudaMemset(h->d_passes, 0, sizeof(unsigned long long)));
    h->eps = eps;
    h->have_problem = true;
    return RIPTRM_OK;
}

template <int P, int MODE>
static int columns_launch(riptrm_handle* h, col::Params& prm, cudaStream_t st) {
    auto kern = col::columns_kernel<P, MODE>;
    const size_t smem = sizeof(col::Smem<P>);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, col::NT, smem));
    if (per_sm < 1) return fail(RIPTRM_E_UNSUPPORTED, "columns kernel does not fit on an SM");
    void* args[] = {&prm};
    if (!h->no_launch_events) CUDA_TRY(cudaEventRecord(h->ev0, st));
    CUDA_TRY(cudaLaunchCooperativeKernel((void*)kern, dim3(h->col_grid), dim3(col::NT), args, smem, st));
    if (!h->no_launch_events) CUDA_TRY(cudaEventRecord(h->ev1, st));
    h->launches += 1;
    return RIPTRM_OK;
}

template <int P, int MODE>
static int stiefel_launch(riptrm_handle* h, col::Params& prm, cudaStream_t st) {
    auto kern = stf::stiefel_kernel<P, MODE>;
    const size_t smem = sizeof(col::Smem<P>);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, col::NT, smem));
    if (per_sm < 1) return fail(RIPTRM_E_UNSUPPORTED, "stiefel kernel does not fit on an SM");
    void* args[] = {&prm};
    if (!h->no_launch_events) CUDA_TRY(cudaEventRecord(h->ev0, st));
    CUDA_TRY(cudaLaunchCooperativeKernel((void*)kern, dim3(h->col_grid), dim3(col::NT), args, smem, st));
    if (!h->no_launch_events) CUDA_TRY(cudaEventRecord(h->ev1, st));
    h->launches += 1;
    return RIPTRM_OK;
}

template <int MODE>
static int columns_dispatch(riptrm_handle* h, col::Params& prm, cudaStream_t st) {
#ifdef RIPTRM_DEV_SPHERE_ONLY   // diagnostic builds (seconds instead of minutes): only the n = 50 TMEM kernels
    return fail(RIPTRM_E_UNSUPPORTED, "RIPTRM_DEV_SPHERE_ONLY build");
#else
    if (is_stiefel(h)) {
        if (MODE == 3) return fail(RIPTRM_E_UNSUPPORTED, "stream-only diagnostic: COLUMNS family only");
        constexpr int M = (MODE == 3) ? 2 : MODE;
        switch (h->colP) {
            case 4: return stiefel_launch<4, M>(h, prm, st);
            case 10: return stiefel_launch<10, M>(h, prm, st);
            case 16: return stiefel_launch<16, M>(h, prm, st);
        }
        return fail(RIPTRM_E_UNSUPPORTED, "unsupported column count");
    }
    switch (h->colP) {
        case 1: return columns_launch<1, MODE>(h, prm, st);
        case 2: return columns_launch<2, MODE>(h, prm, st);
        case 4: return columns_launch<4, MODE>(h, prm, st);
        case 8: return columns_launch<8, MODE>(h, prm, st);
        case 10: return columns_launch<10, MODE>(h, prm, st);
        case 16: return columns_launch<16, MODE>(h, prm, st);
    }
    return fail(RIPTRM_E_UNSUPPORTED, "unsupported column count");
#endif
}

// copies a caller array [n][p] (host or device) into the padded device layout [n_pad][P]
static int columns_import(riptrm_handle* h, double* dst, const double* src, int where, cudaStream_t st) {
    const cudaMemcpyKind kind = (where == RIPTRM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)h->colP * sizeof(double), src, (size_t)h->p * sizeof(double),
                               (size_t)h->p * sizeof(double), h->n, kind, st));
    return RIPTRM_OK;
}
static int columns_export(riptrm_handle* h, double* dst, const double* src, int where, cudaStream_t st) {
    const cudaMemcpyKind kind = (where == RIPTRM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)h->p * sizeof(double), src, (size_t)h->colP * sizeof(double),
                               (size_t)h->p * sizeof(double), h->n, kind, st));
    return RIPTRM_OK;
}

// pointers into the handle's COLUMNS workspace
struct ColPtrs {
    double *vin, *out, *info, *colstate, *summary;
    int* all_done;
};

static ColPtrs columns_fill_params(riptrm_handle* h, col::Params& prm) {
    const size_t arr = (size_t)h->n_pad * h->colP;
    double* b = h->d_colbuf;
    prm.n = h->n;
    prm.n_pad = h->n_pad;
    prm.ld = h->ld;
    prm.p = h->p;
    prm.NIB = h->ld / col::TW;
    prm.NJT = h->n_pad / col::TJ;
    prm.total_tiles = (long long)prm.NIB * prm.NJT;
    prm.R = h->col_R;
    prm.slots = h->col_slots;
    prm.S = h->dS;
    prm.eps = h->eps;
    prm.embedded = h->have_opts ? h->opts.is_euclidean_embedded : 0;
    double** fields[] = {&prm.X, &prm.Y, &prm.Sx, &prm.ys, &prm.c, &prm.V, &prm.Sv, &prm.t, &prm.Hd, &prm.eta,
                         &prm.Heta, &prm.r, &prm.eta2, &prm.Heta2, &prm.r2, &prm.XN, &prm.YN, &prm.SxN, &prm.Xinit,
                         &prm.Yinit, &prm.Sxinit, &prm.Xprev};
    for (int i = 0; i < 22; ++i) *fields[i] = b + i * arr;
    ColPtrs q{};
    q.vin = b + 22 * arr;
    q.out = b + 23 * arr;
    prm.mv_part = b + kColArrays * arr;
    prm.dot_part = prm.mv_part + (size_t)h->col_grid * h->col_slots * col::TW * h->colP;
    q.info = prm.dot_part + (size_t)2 * h->col_grid * stf::DOT_STRIDE;
    prm.tbeg = reinterpret_cast<const long long*>(q.info + 4 * col::MAXP);
    prm.tib0 = reinterpret_cast<const int*>(q.info + 4 * col::MAXP + (h->col_grid + 2));
    q.colstate = q.info + 4 * col::MAXP + (size_t)2 * (h->col_grid + 2);
    q.summary = q.colstate + (size_t)col::MAXP * col::CS_FIELDS;
    q.all_done = reinterpret_cast<int*>(q.summary + (size_t)col::MAXP * RIPTRM_SUMMARY_FIELDS);
    prm.vin = q.vin;
    prm.out = q.out;
    prm.info = q.info;
    prm.colstate = q.colstate;
    prm.summary = q.summary;
    prm.all_done = q.all_done;
    prm.passes = h->d_passes;
    prm.tcg_mininner = h->have_opts ? h->opts.tcg_mininner : 1;
    prm.tcg_maxinner = h->have_opts ? h->opts.tcg_maxinner : -1;
    prm.tcg_theta = h->have_opts ? h->opts.tcg_theta : 1.0;
    prm.tcg_kappa = h->have_opts ? h->opts.tcg_kappa : 0.1;
    return q;
}

static int columns_run(riptrm_handle* h, int mode, const double* x, const double* y, double mu, double Delta,
                       const double* v, double* out, double* info, int where, cudaStream_t st) {
    col::Params prm{};
    const ColPtrs q = columns_fill_params(h, prm);
    double* d_vin = q.vin;
    double* d_out = q.out;
    double* d_info = q.info;
    prm.mu = mu;
    prm.Delta = Delta;
    prm.solve = 0;
    int rc;
    if ((rc = columns_import(h, prm.X, x, where, st)) || (rc = columns_import(h, prm.Y, y, where, st))) return rc;
    if (mode == 1 && (rc = columns_import(h, d_vin, v, where, st))) return rc;
    static const bool stream_only = getenv("RIPTRM_COLUMNS_STREAM_ONLY") != nullptr;  // diagnostic, see fam_columns.cuh
    rc = (mode == 1) ? columns_dispatch<1>(h, prm, st)
                     : (stream_only ? columns_dispatch<3>(h, prm, st) : columns_dispatch<2>(h, prm, st));
    if (rc) return rc;
    if ((rc = columns_export(h, out, d_out, where, st))) return rc;
    if (info != nullptr && mode == 2) {
        const cudaMemcpyKind kind = (where == RIPTRM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        CUDA_TRY(cudaMemcpyAsync(info, d_info, (size_t)(is_stiefel(h) ? 1 : h->p) * 4 * sizeof(double), kind, st));
    }
    if (where == RIPTRM_DEVICE) return RIPTRM_OK;
    CUDA_TRY(cudaStreamSynchronize(st));
    return finish_timing(h, true);
}


template <int P, bool INIT>
static int columns_launch_post(riptrm_handle* h, col::Params& prm, cudaStream_t st) {
    auto kern = col::columns_post_kernel<P, INIT>;
    const size_t smem = sizeof(col::Smem<P>);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {&prm};
    CUDA_TRY(cudaLaunchCooperativeKernel((void*)kern, dim3(h->col_grid), dim3(col::NT), args, smem, st));
    h->launches += 1;
    return RIPTRM_OK;
}
template <int P, bool INIT>
static int stiefel_launch_post(riptrm_handle* h, col::Params& prm, cudaStream_t st) {
    auto kern = stf::stiefel_post_kernel<P, INIT>;
    const size_t smem = sizeof(col::Smem<P>);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {&prm};
    CUDA_TRY(cudaLaunchCooperativeKernel((void*)kern, dim3(h->col_grid), dim3(col::NT), args, smem, st));
    h->launches += 1;
    return RIPTRM_OK;
}
template <bool INIT>
static int columns_dispatch_post(riptrm_handle* h, col::Params& prm, cudaStream_t st) {
#ifdef RIPTRM_DEV_SPHERE_ONLY   // diagnostic builds (seconds instead of minutes): only the n = 50 TMEM kernels
    return fail(RIPTRM_E_UNSUPPORTED, "RIPTRM_DEV_SPHERE_ONLY build");
#else
    if (is_stiefel(h)) {
        switch (h->colP) {
            case 4: return stiefel_launch_post<4, INIT>(h, prm, st);
            case 10: return stiefel_launch_post<10, INIT>(h, prm, st);
            case 16: return stiefel_launch_post<16, INIT>(h, prm, st);
        }
        return fail(RIPTRM_E_UNSUPPORTED, "unsupported column count");
    }
    switch (h->colP) {
        case 1: return columns_launch_post<1, INIT>(h, prm, st);
        case 2: return columns_launch_post<2, INIT>(h, prm, st);
        case 4: return columns_launch_post<4, INIT>(h, prm, st);
        case 8: return columns_launch_post<8, INIT>(h, prm, st);
        case 10: return columns_launch_post<10, INIT>(h, prm, st);
        case 16: return columns_launch_post<16, INIT>(h, prm, st);
    }
    return fail(RIPTRM_E_UNSUPPORTED, "unsupported column count");
#endif
}

// ---- device-side loop of the whole solve ------------------------------------------------------------------------------
// kernel entry points by family and column template (what columns_dispatch / columns_dispatch_post launch)
static void* columns_tcg_entry(const riptrm_handle* h, size_t* smem) {
#ifdef RIPTRM_DEV_SPHERE_ONLY
    return nullptr;
#else
#define RIPTRM_ENTRY(NS, KERN, PP) case PP: *smem = sizeof(col::Smem<PP>); return (void*)NS::KERN<PP, 2>;
    if (is_stiefel(h)) {
        switch (h->colP) { RIPTRM_ENTRY(stf, stiefel_kernel, 4) RIPTRM_ENTRY(stf, stiefel_kernel, 10) RIPTRM_ENTRY(stf, stiefel_kernel, 16) }
        return nullptr;
    }
    switch (h->colP) {
        RIPTRM_ENTRY(col, columns_kernel, 1) RIPTRM_ENTRY(col, columns_kernel, 2) RIPTRM_ENTRY(col, columns_kernel, 4)
        RIPTRM_ENTRY(col, columns_kernel, 8) RIPTRM_ENTRY(col, columns_kernel, 10) RIPTRM_ENTRY(col, columns_kernel, 16)
    }
    return nullptr;
#undef RIPTRM_ENTRY
#endif
}
template <bool INIT>
static void* columns_post_entry(const riptrm_handle* h) {
#ifdef RIPTRM_DEV_SPHERE_ONLY
    return nullptr;
#else
#define RIPTRM_ENTRY(NS, KERN, PP) case PP: return (void*)NS::KERN<PP, INIT>;
    if (is_stiefel(h)) {
        switch (h->colP) { RIPTRM_ENTRY(stf, stiefel_post_kernel, 4) RIPTRM_ENTRY(stf, stiefel_post_kernel, 10) RIPTRM_ENTRY(stf, stiefel_post_kernel, 16) }
        return nullptr;
    }
    switch (h->colP) {
        RIPTRM_ENTRY(col, columns_post_kernel, 1) RIPTRM_ENTRY(col, columns_post_kernel, 2) RIPTRM_ENTRY(col, columns_post_kernel, 4)
        RIPTRM_ENTRY(col, columns_post_kernel, 8) RIPTRM_ENTRY(col, columns_post_kernel, 10) RIPTRM_ENTRY(col, columns_post_kernel, 16)
    }
    return nullptr;
#undef RIPTRM_ENTRY
#endif
}

// The whole solve as ONE graph launch: [clock stamps, post<INIT>] then a conditional WHILE node whose body is [tCG launch,
// post launch] -- all cooperative launches; the post kernel ends the loop with cudaGraphSetConditional when every
// run has finished.  No host involvement between the first launch and the last: `RIPTRM_DEVICE` solves of the COLUMNS /
// STIEFEL families only enqueue work, and the time limits are tested against a device clock.  The graph is kept with the
// parameter block it was built for and rebuilt when a pointer or option in it changes.  Returns > 0 when conditional graph
// nodes are not available (the caller falls back to the host-sequenced loop).
static int columns_solve_graph(riptrm_handle* h, col::Params prm, cudaStream_t st) {
    if (getenv("RIPTRM_COLUMNS_HOST_LOOP") != nullptr) return 1;   // A/B switch
    size_t smem = 0;
    void* tcg_fn = columns_tcg_entry(h, &smem);
    void* post0_fn = columns_post_entry<true>(h);
    void* post_fn = columns_post_entry<false>(h);
    if (tcg_fn == nullptr || post0_fn == nullptr || post_fn == nullptr) return fail(RIPTRM_E_UNSUPPORTED, "unsupported column count");
    if (h->d_now == nullptr) CUDA_TRY(cudaMalloc(&h->d_now, 2 * sizeof(double)));
    prm.now_ptr = h->d_now;
    prm.now_s = 0.0;
    prm.cond_on = 1;
    prm.cond = 0;
    if (h->col_graph != nullptr) {   // same launches as last time?  (the handle sits inside the block: compare without it)
        col::Params a = h->col_graph_prm;
        a.cond = 0;
        if (memcmp(&a, &prm, sizeof(prm)) != 0) {
            cudaGraphExecDestroy(h->col_graph);
            h->col_graph = nullptr;
        }
    }
    if (h->col_graph == nullptr) {
        for (void* fn : {tcg_fn, post0_fn, post_fn})
            CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaGraph_t g = nullptr;
        CUDA_TRY(cudaGraphCreate(&g, 0));
        cudaGraphConditionalHandle ch;
        if (cudaGraphConditionalHandleCreate(&ch, g, 1, cudaGraphCondAssignDefault) != cudaSuccess) {
            cudaGetLastError();
            cudaGraphDestroy(g);
            return 1;
        }
        col::Params gp = prm;
        gp.cond = (unsigned long long)ch;
        double* now = h->d_now;
        unsigned long long* t0 = reinterpret_cast<unsigned long long*>(h->d_now + 1);
        int one = 1, zero = 0;
        auto add_kernel = [&](cudaGraph_t graph, cudaGraphNode_t* dep, void* fn, dim3 grid, dim3 block, size_t sm, void** args,
                              bool coop, cudaGraphNode_t* out) -> cudaError_t {
            cudaKernelNodeParams kp = {};
            kp.func = fn;
            kp.gridDim = grid;
            kp.blockDim = block;
            kp.sharedMemBytes = (unsigned)sm;
            kp.kernelParams = args;
            cudaError_t e = cudaGraphAddKernelNode(out, graph, dep, dep ? 1 : 0, &kp);
            if (e == cudaSuccess && coop) {
                cudaLaunchAttributeValue v = {};
                v.cooperative = 1;
                e = cudaGraphKernelNodeSetAttribute(*out, cudaLaunchAttributeCooperative, &v);
            }
            return e;
        };
        void* stamp_fn = (void*)col::stamp_kernel;
        void* a_stamp0[] = {&now, &t0, &one};
        void* a_stamp[] = {&now, &t0, &zero};
        void* a_prm[] = {&gp};
        // start of the clock, then a first reading (> 0: a run whose time budget is already spent must stop at once); every
        // post launch leaves the reading for the next iteration
        cudaGraphNode_t n_stamp0, n_stamp1, n_post0, n_while, b_tcg, b_post;
        cudaError_t e = add_kernel(g, nullptr, stamp_fn, dim3(1), dim3(1), 0, a_stamp0, false, &n_stamp0);
        if (e == cudaSuccess) e = add_kernel(g, &n_stamp0, stamp_fn, dim3(1), dim3(1), 0, a_stamp, false, &n_stamp1);
        if (e == cudaSuccess) e = add_kernel(g, &n_stamp1, post0_fn, dim3(h->col_grid), dim3(col::NT), smem, a_prm, true, &n_post0);
        cudaGraphNodeParams wp = {cudaGraphNodeTypeConditional};
        wp.conditional.handle = ch;
        wp.conditional.type = cudaGraphCondTypeWhile;
        wp.conditional.size = 1;
        if (e == cudaSuccess) e = cudaGraphAddNode(&n_while, g, &n_post0, 1, &wp);
        if (e == cudaSuccess) {
            cudaGraph_t body = wp.conditional.phGraph_out[0];
            e = add_kernel(body, nullptr, tcg_fn, dim3(h->col_grid), dim3(col::NT), smem, a_prm, true, &b_tcg);
            if (e == cudaSuccess) e = add_kernel(body, &b_tcg, post_fn, dim3(h->col_grid), dim3(col::NT), smem, a_prm, true, &b_post);
        }
        if (e == cudaSuccess) e = cudaGraphInstantiate(&h->col_graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) {
            cudaGetLastError();
            h->col_graph = nullptr;
            if (getenv("RIPTRM_DEBUG_GRAPH") != nullptr) fprintf(stderr, "[riptrm] columns graph: %s\n", cudaGetErrorString(e));
            return 1;   // e.g. a driver without conditional nodes: host-sequenced loop
        }
        h->col_graph_prm = gp;
    }
    CUDA_TRY(cudaGraphLaunch(h->col_graph, st));
    h->launches += 3;   // the graph launch; the launches inside it are counted on the device (matvec passes, inner iterations)
    return RIPTRM_OK;
}

// riptrm_solve on the COLUMNS family: p independent RIPTRM runs sharing S, advanced in lock-step.  One graph launch with a
// device-side loop (columns_solve_graph); where conditional graph nodes are missing the host sequences the launches (tCG for
// all columns, then the rest of the trust-region iteration) a few iterations ahead of the flag it polls.
static int columns_solve(riptrm_handle* h, const double* x0, const double* y0, double* x, double* y, double* summary,
                         double* trace, int where, cudaStream_t st) {
    const riptrm_options& o = h->opts;
    col::Params prm{};
    const ColPtrs q = columns_fill_params(h, prm);
    prm.solve = 1;
    prm.reuse_heta = (getenv("RIPTRM_RECOMPUTE_HDX") == nullptr) ? 1 : 0;
    prm.mu_sched = h->d_sched;
    prm.tolL_sched = h->d_sched + h->sched_len;
    prm.tolC_sched = h->d_sched + 2 * h->sched_len;
    prm.maxiter = o.maxiter;
    prm.inner_maxiter = o.inner_maxiter;
    prm.trace_mode = o.trace_mode;
    prm.trace_capacity = o.trace_capacity;
    prm.tolresid = o.tolresid;
    prm.initial_tr_radius = o.initial_tr_radius;
    prm.minimal_initial_tr_radius = o.minimal_initial_tr_radius;
    prm.maximal_tr_radius = o.maximal_tr_radius;
    prm.rho = o.rho;
    prm.reduction_regularization = o.reduction_regularization;
    prm.gamma = o.gamma;
    prm.const_left = o.const_left;
    prm.const_right = o.const_right;
    prm.maxtime = o.maxtime;
    prm.inner_maxtime = o.inner_maxtime;
    const auto wall0 = std::chrono::steady_clock::now();
    auto seconds_since_start = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count(); };
    const int nruns = is_stiefel(h) ? 1 : h->p;  // STIEFEL: one run; COLUMNS: one per column
    const size_t tb = (o.trace_mode != 0 && trace != nullptr)
                          ? (size_t)nruns * o.trace_capacity * RIPTRM_TRACE_FIELDS * sizeof(double) : 0;
    if (tb != 0) {
        if (where == RIPTRM_DEVICE) {
            prm.trace = trace;
        } else {
            if (h->d_trace != nullptr && h->trace_bytes != tb) free_dev(h->d_trace);
            int rc0;
            if ((rc0 = ensure(h->d_trace, tb))) return rc0;
            h->trace_bytes = tb;
            prm.trace = h->d_trace;
        }
    }
    int rc;
    if ((rc = columns_import(h, prm.X, x0, where, st)) || (rc = columns_import(h, prm.Y, y0, where, st))) return rc;
    CUDA_TRY(cudaMemsetAsync(q.colstate, 0, (size_t)col::MAXP * col::CS_FIELDS * sizeof(double), st));
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    const int grc = columns_solve_graph(h, prm, st);
    if (grc < 0) return grc;
    const bool host_loop = grc > 0;
    prm.now_s = seconds_since_start();
    if (host_loop && (rc = columns_dispatch_post<true>(h, prm, st))) return rc;
    const long long max_rounds = (long long)(o.maxiter + 1) * (o.inner_maxiter > 0 ? o.inner_maxiter : 100000);
    // The host sequences the launches (one tCG launch and one post launch per trust-region iteration) kColLookahead
    // iterations AHEAD of the "all done" flag: after every iteration the flag is copied to a pinned slot and an event is
    // recorded; before enqueuing iteration r the host waits for the flag of iteration r - kColLookahead only, so the stream
    // never runs dry (round 1 synchronised on every iteration: 4-14 % of a config-4 solve).  The few launches enqueued
    // behind the last iteration see the flag on the device and return at once.
    constexpr int kColLookahead = 4;
    if (h->h_done == nullptr) {
        CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&h->h_done), kColLookahead * sizeof(int), cudaHostAllocDefault));
        for (int i = 0; i < kColLookahead; ++i) CUDA_TRY(cudaEventCreateWithFlags(&h->done_ev[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < kColLookahead; ++i) h->h_done[i] = 0;
    h->no_launch_events = true;
    for (long long round = 0; host_loop && round < max_rounds; ++round) {
        const int slot = (int)(round % kColLookahead);
        if (round >= kColLookahead) {
            if (cudaEventSynchronize(h->done_ev[slot]) != cudaSuccess) { rc = fail(RIPTRM_E_CUDA, "columns solve: flag poll failed"); break; }
            if (h->h_done[slot]) break;
        }
        prm.now_s = seconds_since_start();   // as of enqueue time: the time limits are tested a few iterations late at most
        if ((rc = columns_dispatch<2>(h, prm, st))) break;
        if ((rc = columns_dispatch_post<false>(h, prm, st))) break;
        if (cudaMemcpyAsync(&h->h_done[slot], q.all_done, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaEventRecord(h->done_ev[slot], st) != cudaSuccess) { rc = fail(RIPTRM_E_CUDA, "columns solve: flag copy failed"); break; }
    }
    h->no_launch_events = false;
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    if (x != nullptr && (rc = columns_export(h, x, prm.X, where, st))) return rc;
    if (y != nullptr && (rc = columns_export(h, y, prm.Y, where, st))) return rc;
    const cudaMemcpyKind kind = (where == RIPTRM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (summary != nullptr)
        CUDA_TRY(cudaMemcpyAsync(summary, q.summary, (size_t)nruns * RIPTRM_SUMMARY_FIELDS * sizeof(double), kind, st));
    if (tb != 0 && where != RIPTRM_DEVICE) CUDA_TRY(cudaMemcpyAsync(trace, h->d_trace, tb, cudaMemcpyDeviceToHost, st));
    if (where == RIPTRM_DEVICE && !host_loop) return RIPTRM_OK;   // enqueue-only, like the batched families
    CUDA_TRY(cudaStreamSynchronize(st));
    return finish_timing(h, true);
}

extern "C" int riptrm_set_nonnegpca(riptrm_handle* h, const double* Z, int batch_z, double eps, int where) {
    if (h == nullptr || Z == nullptr) return fail(RIPTRM_E_INVALID, "NULL argument");
    if (h->family != RIPTRM_FAMILY_NONNEGPCA_SPHERE && h->family != RIPTRM_FAMILY_NONNEGPCA_COLUMNS && !is_stiefel(h))
        return fail(RIPTRM_E_INVALID, "handle is not a NonnegPCA family");
    if (batch_z < 1 || h->batch % batch_z != 0) return fail(RIPTRM_E_INVALID, "batch_z must divide batch (batch_z instances x batch/batch_z initial points)");
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->family == RIPTRM_FAMILY_NONNEGPCA_COLUMNS || is_stiefel(h)) return columns_setup(h, Z, eps, where);
    const size_t bytes = (size_t)batch_z * h->n * h->n * sizeof(double);
    if (where == RIPTRM_DEVICE) {
        if (h->ownZ) free_dev(h->dZ);
        h->dZ = const_cast<double*>(Z);
        h->ownZ = false;
        h->z_bytes = 0;
    } else {
        // staging buffer is kept across calls (re-binding new data every step must not pay cudaMalloc / cudaFree)
        if (!h->ownZ || h->z_bytes != bytes) {
            if (h->ownZ) free_dev(h->dZ);
            h->dZ = nullptr;
            CUDA_TRY(cudaMalloc(&h->dZ, bytes));
            h->ownZ = true;
            h->z_bytes = bytes;
        }
        CUDA_TRY(cudaMemcpyAsync(h->dZ, Z, bytes, cudaMemcpyHostToDevice, 0));
        CUDA_TRY(cudaStreamSynchronize(0));  // the caller may reuse its buffer on return
    }
    h->batch_z = batch_z;
    h->eps = eps;
    h->have_problem = true;
    return RIPTRM_OK;
}

extern "C" int riptrm_set_rosenbrock(riptrm_handle* h, double alpha, double offset) {
    if (h == nullptr) return fail(RIPTRM_E_INVALID, "NULL argument");
    if (h->family != RIPTRM_FAMILY_ROSENBROCK_GRASSMANN) return fail(RIPTRM_E_INVALID, "handle is not the Rosenbrock family");
    h->ros_alpha = alpha;
    h->ros_offset = offset;
    h->have_problem = true;
    return RIPTRM_OK;
}
extern "C" int riptrm_set_stableid(riptrm_handle* h, const double* X, const double* XP, int N, double hstep,
                                   const double* conspec, int m, int where) {
    if (h == nullptr || X == nullptr || XP == nullptr || conspec == nullptr) return fail(RIPTRM_E_INVALID, "NULL argument");
    if (h->family != RIPTRM_FAMILY_STABLEID_PRODUCT) return fail(RIPTRM_E_INVALID, "handle is not the StableIdentification family");
    if (m != h->m || N <= 0) return fail(RIPTRM_E_INVALID, "m must match riptrm_create, N > 0");
    for (int i = 0; where != RIPTRM_DEVICE && i < m; ++i) {
        const double* r = conspec + 5 * i;
        if (!(r[0] == 0 || r[0] == 1 || r[0] == 2) || r[1] < 0 || r[1] >= h->n || r[2] < 0 || r[2] >= h->n)
            return fail(RIPTRM_E_INVALID, "conspec row: kind in {0,1,2}, 0 <= row, col < d");
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t dn = (size_t)h->n * N;
    free_dev(h->d_sid);
    CUDA_TRY(cudaMalloc(&h->d_sid, (2 * dn + 5 * (size_t)m) * sizeof(double)));
    const cudaMemcpyKind kind = (where == RIPTRM_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CUDA_TRY(cudaMemcpy(h->d_sid, X, dn * sizeof(double), kind));
    CUDA_TRY(cudaMemcpy(h->d_sid + dn, XP, dn * sizeof(double), kind));
    CUDA_TRY(cudaMemcpy(h->d_sid + 2 * dn, conspec, 5 * (size_t)m * sizeof(double), kind));
    h->sid_N = N;
    h->sid_h = hstep;
    h->have_problem = true;
    return RIPTRM_OK;
}

extern "C" int riptrm_set_options(riptrm_handle* h, const riptrm_options* o) {
    if (h == nullptr || o == nullptr) return fail(RIPTRM_E_INVALID, "NULL argument");
    if (o->maxiter < 0) return fail(RIPTRM_E_INVALID, "maxiter < 0");
    if (o->mu_sched == nullptr || o->tol_lagrangian_sched == nullptr || o->tol_complementarity_sched == nullptr)
        return fail(RIPTRM_E_INVALID, "schedules are required (length maxiter + 1)");
    if (o->trace_mode < 0 || o->trace_mode > 2) return fail(RIPTRM_E_INVALID, "trace_mode must be 0, 1 or 2");
    if (o->trace_mode != 0 && o->trace_capacity <= 0) return fail(RIPTRM_E_INVALID, "trace_capacity must be positive");
    if (o->trs_solver != RIPTRM_TRS_SOLVER_TCG && o->trs_solver != RIPTRM_TRS_SOLVER_EXACT_REPMAT)
        return fail(RIPTRM_E_INVALID, "trs_solver must be RIPTRM_TRS_SOLVER_TCG or RIPTRM_TRS_SOLVER_EXACT_REPMAT");
    if (o->second_order_stationarity && o->trs_solver != RIPTRM_TRS_SOLVER_EXACT_REPMAT)
        return fail(RIPTRM_E_INVALID, "second_order_stationarity needs trs_solver = EXACT_REPMAT (RIPTRM.py:599)");
    if (o->second_order_stationarity && o->tol_second_order_sched == nullptr)
        return fail(RIPTRM_E_INVALID, "second_order_stationarity needs tol_second_order_sched (length maxiter + 1)");
    if (o->trs_solver == RIPTRM_TRS_SOLVER_EXACT_REPMAT) {
        if (h->family == RIPTRM_FAMILY_NONNEGPCA_COLUMNS || is_stiefel(h))
            return fail(RIPTRM_E_UNSUPPORTED, "Exact_RepMat needs a dim x dim representation matrix: not built for the large-n families");
        if (h->family == RIPTRM_FAMILY_NONNEGPCA_SPHERE && h->n > 64)
            return fail(RIPTRM_E_UNSUPPORTED, "Exact_RepMat on Sphere(n): n <= 64 (three dim x dim matrices per pair in shared memory)");
        if (h->family == RIPTRM_FAMILY_ROSENBROCK_GRASSMANN && h->n * (h->n - h->p) > GrassmannFam::kPerpDoubles)
            return fail(RIPTRM_E_UNSUPPORTED, "Exact_RepMat on Grassmann(n, p): n (n - p) <= 128");
    }
    CUDA_TRY(cudaSetDevice(h->device));
    const int len = o->maxiter + 1;
    free_dev(h->d_sched);
    CUDA_TRY(cudaMalloc(&h->d_sched, (size_t)4 * len * sizeof(double)));
    CUDA_TRY(cudaMemset(h->d_sched, 0, (size_t)4 * len * sizeof(double)));
    if (o->tol_second_order_sched != nullptr)
        CUDA_TRY(cudaMemcpy(h->d_sched + 3 * len, o->tol_second_order_sched, len * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d_sched, o->mu_sched, len * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d_sched + len, o->tol_lagrangian_sched, len * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(h->d_sched + 2 * len, o->tol_complementarity_sched, len * sizeof(double), cudaMemcpyHostToDevice));
    h->sched_len = len;
    h->opts = *o;
    h->opts.mu_sched = h->opts.tol_lagrangian_sched = h->opts.tol_complementarity_sched = nullptr;
    h->opts.tol_second_order_sched = nullptr;
    h->have_opts = true;
    return RIPTRM_OK;
}

static DevOpts make_devopts(const riptrm_handle* h) {
    const riptrm_options& s = h->opts;
    DevOpts o;
    o.maxiter = s.maxiter;
    o.inner_maxiter = s.inner_maxiter;
    o.tcg_mininner = s.tcg_mininner;
    o.tcg_maxinner = s.tcg_maxinner;
    o.is_euclidean_embedded = s.is_euclidean_embedded;
    o.recompute_hdx = getenv("RIPTRM_RECOMPUTE_HDX") != nullptr ? 1 : 0;   // A/B switch: the reference's fresh Hw[dx] at :659
    o.trace_mode = s.trace_mode;
    o.trace_capacity = s.trace_capacity;
    o.tolresid = s.tolresid;
    o.maxtime = s.maxtime;
    o.inner_maxtime = s.inner_maxtime;
    o.initial_tr_radius = s.initial_tr_radius;
    o.minimal_initial_tr_radius = s.minimal_initial_tr_radius;
    o.maximal_tr_radius = s.maximal_tr_radius;
    o.rho = s.rho;
    o.reduction_regularization = s.reduction_regularization;
    o.gamma = s.gamma;
    o.const_left = s.const_left;
    o.const_right = s.const_right;
    o.tcg_theta = s.tcg_theta;
    o.tcg_kappa = s.tcg_kappa;
    o.mu = h->d_sched;
    o.tolL = h->d_sched + h->sched_len;
    o.tolC = h->d_sched + 2 * h->sched_len;
    o.tolS = h->d_sched + 3 * h->sched_len;
    o.second_order = s.second_order_stationarity;
    o.trs_tolhardcase = s.trs_tolhardcase;
    return o;
}
static bool exact_repmat(const riptrm_handle* h) { return h->have_opts && h->opts.trs_solver == RIPTRM_TRS_SOLVER_EXACT_REPMAT; }

static int ensure(double*& p, size_t bytes) {
    if (p != nullptr) return RIPTRM_OK;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(RIPTRM_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    return RIPTRM_OK;
}

// n = 50 with S in Tensor Memory: 4 independent warps per CTA (one per TMEM sub-partition), 2 CTAs per SM (each
// allocates 256 of the 512 TMEM columns), every warp pulling pairs from the same atomic queue as sphere_kernel.
#ifndef RIPTRM_RS1
#define RIPTRM_RS1 1     // sphere_tmem_kernel (one warp per copy; the fast lane, small batches): tCG reductions through shared memory
#endif
constexpr int kTmem1Scratch = 64 + kRedDoubles;   // per warp after the staging copy: broadcast operand + reduction rows
template <int MODE>
__global__ void __launch_bounds__(128, 2) sphere_tmem_kernel(SphereParams P, DevOpts o, int* counter) {
    using F = SphereFam<2, 50, true, 64, RIPTRM_RS1 != 0>;
    constexpr int K = 2;
    extern __shared__ __align__(16) double smem[];
    __shared__ uint32_t tmem_base;
    const int n = 50, ns = 50, pad = 64;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    double* my = smem + (size_t)warp * (n * ns + pad + kTmem1Scratch);
    if (P.lane_debug != nullptr && threadIdx.x == 0) {   // placement record of the two-kernel lane: (smid << 4) | role
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        P.lane_debug[blockIdx.x] = (int)(smid << 4) | kLaneLane;
        if (P.lane_times != nullptr) P.lane_times[2 * blockIdx.x] = global_timer_ns();
    }
    if (P.lane_arrive != nullptr && threadIdx.x == 0) {   // this CTA is resident: count it in for the main kernel's stream wait
        atomicAdd(P.lane_arrive, 1u);
        __threadfence_system();
    }
    if (warp == 0) tmem::alloc(&tmem_base, 256);
    tmem::fence_before_sync();
    __syncthreads();
    tmem::fence_after_sync();
    typename F::Ctx ctx;
    ctx.S = my;                       // staging copy; S.v reads the TMEM copy
    ctx.vbuf = my + n * ns + pad;
    ctx.ws.init(ctx.vbuf);
    ctx.n = n;
    ctx.ns = ns;
    ctx.eps = P.eps;
    ctx.embedded = o.is_euclidean_embedded != 0;
    ctx.taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
    int loaded_z = -1;
    while (true) {
        int inst = 0;
        if (lane == 0) inst = atomicAdd(counter, 1);
        inst = __shfl_sync(kFull, inst, 0);
        if (inst >= (P.queue_len > 0 ? P.queue_len : P.batch)) break;
        if (P.order != nullptr) inst = P.order[inst];
        double* pause = (P.pause != nullptr) ? P.pause + (size_t)inst * kPauseFields : nullptr;
        if (MODE == 0 && P.resume && pause[7] == 0.0) continue;
        const int zi = inst / (P.batch / P.batch_z);
        if (zi != loaded_z) {
            load_S(P.Z + (size_t)zi * n * n, my, n, ns, pad);
            F::stage_to_tmem(ctx);
            loaded_z = zi;
        }
        const bool resume = (MODE == 0) && P.resume;
        const typename F::Vec x0 = load_vec<K>((resume ? P.x : P.x0) + (size_t)inst * n, n);
        const typename F::CVec y0 = load_vec<K>((resume ? P.y : P.y0) + (size_t)inst * n, n);
        if (MODE == 0) {
            typename F::Pt pt;
            typename F::CVec y;
            double* tr = (P.trace != nullptr && o.trace_mode != 0)
                             ? P.trace + (size_t)inst * o.trace_capacity * RIPTRM_TRACE_FIELDS
                             : nullptr;
            solve_instance<F>(ctx, o, x0, y0, pt, y, P.summary ? P.summary + (size_t)inst * RIPTRM_SUMMARY_FIELDS : nullptr,
                              tr, pause, resume, P.pause_at);
            if (P.x) store_vec<K>(P.x + (size_t)inst * n, pt.x, n);
            if (P.y) store_vec<K>(P.y + (size_t)inst * n, y, n);
        } else {
            typename F::Pt pt;
            F::eval_point(ctx, x0, pt);
            typename F::Step st;
            F::begin_step(ctx, pt, y0, P.mu, st);
            if (MODE == 1) {
                const typename F::Vec v = load_vec<K>(P.v + (size_t)inst * n, n);
                const typename F::Vec hv = F::Hw(ctx, pt, y0, st, v);
                store_vec<K>(P.out + (size_t)inst * n, hv, n);
            } else {
                typename F::Vec eta, Heta;
                const TcgResult r = F::tcg(ctx, o, pt, y0, st, P.Delta, eta, Heta);
                store_vec<K>(P.out + (size_t)inst * n, eta, n);
                const double nrm = sqrt(F::inner(ctx, pt, eta, eta));
                if (P.info != nullptr && lane < 4) {
                    const double val = (lane == 0) ? (double)r.iters : (lane == 1) ? (double)r.stop : (lane == 2) ? nrm : r.model_value;
                    P.info[(size_t)inst * 4 + lane] = val;
                }
            }
        }
    }
    tmem::fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem::dealloc(tmem_base, 256);
    if (P.lane_times != nullptr && threadIdx.x == 0) P.lane_times[2 * blockIdx.x + 1] = global_timer_ns();
}

// The same with two warps per copy of S.  Consecutive pairs of one instance (its initial points) share Z, and a warp of the
// same TMEM sub-partition in the same CTA can read what its neighbour staged: 8 warps per CTA, warps w and w + 4 work on
// the two pairs of a unit from one copy of S in sub-partition w % 4.  16 resident warps per SM instead of 8 (the
// register file then allows 128 registers per thread), TMEM and staging traffic per pair halved.  The pairs of an
// instance do nearly the same work (correlation 0.99), so the warps of a unit finish together.
// Fast lane BY CONSTRUCTION (round 2; replaces a second kernel on a priority stream whose CTAs were given their SMs by a
// 30 us sleep kernel and a 150 KB shared-memory request).  The grid is two CTAs per SM, which is what fits (128 registers x 256
// threads, 86 KB).  The first CTA to arrive on an SM (an atomic per %smid) claims one of `lane_ctas` lane slots or declares
// the SM a main SM; the second CTA reads that decision -- the first one is running, so the wait is bounded -- and, on a lane
// SM, steps aside: it waits until all CTAs of the grid have arrived (none is pending that could be placed next to the lane
// CTA; the wait is capped at 200 us so that a smaller residency than expected costs speed, never progress) and exits.
// A lane CTA therefore owns its SM: its warps
// 0..3 -- one per scheduler, the residency at which a tCG iteration takes 1.0 us instead of 2.4 -- solve the pairs of the
// longest units (`fast_order`), its warps 4..7 have nothing to do.  `lane_debug` records (smid, role) per CTA for the test.

#ifndef RIPTRM_RS2
#define RIPTRM_RS2 1     // the two reductions of a tCG iteration through shared memory (wsum_smem) instead of shuffles
#endif
#ifndef RIPTRM_TCH2
#define RIPTRM_TCH2 16   // TMEM columns per chunk of S in S.v (16 / 32; same order of additions).  16: the chunk's registers no
                         // longer force ptxas to shuffle S data around (83 -> 39 register moves per tCG iteration)
#endif
template <int MODE>
__global__ void __launch_bounds__(256, 2) sphere_tmem2_kernel(SphereParams P, DevOpts o, int* counter) {
    using F = SphereFam<2, 50, true, RIPTRM_TCH2, RIPTRM_RS2 != 0>;
    constexpr int K = 2;
    extern __shared__ __align__(16) double smem[];
    __shared__ uint32_t tmem_base;
    __shared__ int unit_slot[4];
    __shared__ int cta_role;
    const int n = 50, ns = 50, pad = 64;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int q = warp & 3, role = warp >> 2;
    double* stage = smem + (size_t)q * (n * ns + pad);
    if (P.lane_ctas > 0) {
        if (threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            int* st = P.lane_state;
            int r = kLaneMain;
            if (smid < (unsigned)kLaneMaxSms) {
                const int slot = atomicAdd(st + 2 + smid, 1);
                if (slot == 0) {
                    r = (atomicAdd(st, 1) < P.lane_ctas) ? kLaneLane : kLaneMain;
                    __threadfence();
                    atomicExch(st + 2 + kLaneMaxSms + smid, r);
                } else {
                    int first;
                    do {
                        first = atomicAdd(st + 2 + kLaneMaxSms + smid, 0);   // the first CTA of this SM is running: bounded wait
                    } while (first == 0);
                    r = (first == kLaneLane) ? kLaneExit : kLaneMain;
                }
            }
            atomicAdd(st + 1, 1);
            if (r == kLaneExit) {   // all CTAs placed: nobody can follow us onto this SM
                const uint64_t t0 = global_timer_ns();
                while (atomicAdd(st + 1, 0) < (int)gridDim.x && global_timer_ns() - t0 < 200000ull) __nanosleep(200);
            }
            if (P.lane_debug != nullptr) P.lane_debug[blockIdx.x] = (int)(smid << 4) | r;
            if (P.lane_times != nullptr) P.lane_times[2 * blockIdx.x] = global_timer_ns();
            cta_role = r;
        }
        __syncthreads();
        if (cta_role == kLaneExit) return;
    } else if (threadIdx.x == 0) {
        cta_role = kLaneMain;
        if (P.lane_debug != nullptr) {   // main kernel next to the two-kernel lane: record where this CTA runs
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            P.lane_debug[blockIdx.x] = (int)(smid << 4) | kLaneMain;
            if (P.lane_times != nullptr) P.lane_times[2 * blockIdx.x] = global_timer_ns();
        }
    }
    if (warp == 0) tmem::alloc(&tmem_base, 256);
    tmem::fence_before_sync();
    __syncthreads();
    tmem::fence_after_sync();
    typename F::Ctx ctx;
    ctx.S = stage;                    // staging copy of the sub-partition; S.v reads the TMEM copy
    ctx.vbuf = smem + (size_t)4 * (n * ns + pad) + (size_t)warp * kWarpScratchDoubles;
    ctx.ws.init(ctx.vbuf);
    ctx.n = n;
    ctx.ns = ns;
    ctx.eps = P.eps;
    ctx.embedded = o.is_euclidean_embedded != 0;
    ctx.taddr = tmem_base + ((uint32_t)(32 * q) << 16);
    const int units = P.batch / 2, ipp = P.batch / P.batch_z;
    auto pair_bar = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); };
    int loaded_z = -1;
    const bool lane_cta = cta_role == kLaneLane;
    while (lane_cta && role == 0) {
        // lane worker: one pair at a time from the fast queue, this warp's own copy of S in its TMEM sub-partition
        int i = 0;
        if (lane == 0) i = atomicAdd(counter + 1, 1);
        i = __shfl_sync(kFull, i, 0);
        if (i >= P.fast_len) break;
        const int inst = P.fast_order[i];
        double* pause = P.pause + (size_t)inst * kPauseFields;
        if (pause[7] != 2.0) continue;           // finished in an earlier launch
        const int zi = inst / ipp;
        if (zi != loaded_z) {
            load_S(P.Z + (size_t)zi * n * n, stage, n, ns, pad);
            F::stage_to_tmem(ctx);
            loaded_z = zi;
        }
        const typename F::Vec x0 = load_vec<K>(P.x + (size_t)inst * n, n);
        const typename F::CVec y0 = load_vec<K>(P.y + (size_t)inst * n, n);
        typename F::Pt pt;
        typename F::CVec y;
        double* tr = (P.trace != nullptr && o.trace_mode != 0) ? P.trace + (size_t)inst * o.trace_capacity * RIPTRM_TRACE_FIELDS
                                                               : nullptr;
        solve_instance<F>(ctx, o, x0, y0, pt, y, P.summary ? P.summary + (size_t)inst * RIPTRM_SUMMARY_FIELDS : nullptr, tr,
                          pause, true, P.pause_at);
        if (P.x) store_vec<K>(P.x + (size_t)inst * n, pt.x, n);
        if (P.y) store_vec<K>(P.y + (size_t)inst * n, y, n);
    }
    while (!lane_cta) {
        if (role == 0 && lane == 0) unit_slot[q] = atomicAdd(counter, 1);
        pair_bar();
        int unit = unit_slot[q];
        pair_bar();                   // both warps have read the slot (and are done with the previous unit's S)
        if (unit >= units) break;
        if (P.order != nullptr) unit = P.order[unit];
        const int inst = 2 * unit + role;
        const bool resume = P.resume != 0;
        // paused[7]: 1 paused, 0 finished in an earlier launch, 2 handed to the fast lane of this launch
        const bool skip0 = resume && P.pause[(size_t)(2 * unit) * kPauseFields + 7] != 1.0;
        const bool skip1 = resume && P.pause[(size_t)(2 * unit + 1) * kPauseFields + 7] != 1.0;
        if (skip0 && skip1) continue;  // both finished in an earlier launch (uniform over the two warps)
        const int zi = (2 * unit) / ipp;
        if (zi != loaded_z) {
            if (role == 0) {
                load_S(P.Z + (size_t)zi * n * n, stage, n, ns, pad);
                F::stage_to_tmem(ctx);
            }
            tmem::fence_before_sync();
            pair_bar();
            tmem::fence_after_sync();
            loaded_z = zi;
        }
        if (role == 0 ? skip0 : skip1) continue;
        double* pause = (P.pause != nullptr) ? P.pause + (size_t)inst * kPauseFields : nullptr;
        const typename F::Vec x0 = load_vec<K>((resume ? P.x : P.x0) + (size_t)inst * n, n);
        const typename F::CVec y0 = load_vec<K>((resume ? P.y : P.y0) + (size_t)inst * n, n);
        typename F::Pt pt;
        typename F::CVec y;
        double* tr = (P.trace != nullptr && o.trace_mode != 0) ? P.trace + (size_t)inst * o.trace_capacity * RIPTRM_TRACE_FIELDS
                                                               : nullptr;
        solve_instance<F>(ctx, o, x0, y0, pt, y, P.summary ? P.summary + (size_t)inst * RIPTRM_SUMMARY_FIELDS : nullptr, tr,
                          pause, resume, P.pause_at);
        if (P.x) store_vec<K>(P.x + (size_t)inst * n, pt.x, n);
        if (P.y) store_vec<K>(P.y + (size_t)inst * n, y, n);
    }
    tmem::fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem::dealloc(tmem_base, 256);
    if (P.lane_times != nullptr && threadIdx.x == 0) P.lane_times[2 * blockIdx.x + 1] = global_timer_ns();
}

static int launch_sphere_tmem2(riptrm_handle* h, const SphereParams& P, const DevOpts& o, cudaStream_t st) {
    const size_t smem = ((size_t)4 * (50 * 50 + 64) + 8 * kWarpScratchDoubles) * sizeof(double);
    auto kern = sphere_tmem2_kernel<0>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int grid = h->num_sms * 2;
    const int need = (h->batch / 2 + 3) / 4;
    if (grid > need) grid = need;
    CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, sizeof(int), st));
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    if (P.lane_ctas > 0) grid = h->num_sms * 2;   // the lane election counts on two CTAs per SM
    kern<<<grid, 256, smem, st>>>(P, o, h->d_counter);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    h->launches += 1;
    return RIPTRM_OK;
}

// can the lane be carved out of the main kernel's own grid?  (SM ids within the election table, enough units to fill the grid)
static bool lane_in_kernel_possible(riptrm_handle* h) {
    return h->num_sms <= kLaneMaxSms && h->batch / 2 >= h->num_sms * 8;
}

template <int MODE>
static int launch_sphere_tmem(riptrm_handle* h, const SphereParams& P, const DevOpts& o, cudaStream_t st) {
    const size_t smem = (size_t)4 * (50 * 50 + 64 + kTmem1Scratch) * sizeof(double);
    auto kern = sphere_tmem_kernel<MODE>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int grid = h->num_sms * 2;
    const int need = (h->batch + 3) / 4;
    if (grid > need) grid = need;
    CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, sizeof(int), st));
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    kern<<<grid, 128, smem, st>>>(P, o, h->d_counter);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    h->launches += 1;
    return RIPTRM_OK;
}

template <int K, int MODE, int NFIX, bool EXACT = false>
static int launch_sphere(riptrm_handle* h, const SphereParams& P, const DevOpts& o, cudaStream_t st) {
    const int n = h->n;
    const int ns = (n + 1) & ~1;
    const size_t smem = ((size_t)(n * ns + 32 * K + 32 * K) + (EXACT ? rep_doubles(n - 1) : 0)) * sizeof(double);
    auto kern = sphere_kernel<K, MODE, NFIX, EXACT>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
    if (per_sm < 1) return fail(RIPTRM_E_UNSUPPORTED, "instance does not fit in shared memory");
    int grid = h->num_sms * per_sm;
    if (grid > h->batch) grid = h->batch;
    CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, sizeof(int), st));
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    kern<<<grid, 32, smem, st>>>(P, o, h->d_counter);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    h->launches += 1;
    return RIPTRM_OK;
}

// sort key of the second launch: work spent in the first one (tCG iterations + 2 per trust-region iteration);
// pairs that already finished sort last
__global__ void schedule_keys_kernel(const double* __restrict__ pause, float* keys, int* idx, int batch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const double* p = pause + (size_t)i * kPauseFields;
    keys[i] = (p[7] != 0.0) ? (float)(p[3] + 2.0 * p[2]) : -1.0f;
    idx[i] = i;
}

// the same per unit of two consecutive pairs (sphere_tmem2_kernel): the sum of the two pairs' work
__global__ void schedule_unit_keys_kernel(const double* __restrict__ pause, float* keys, int* idx, int units) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= units) return;
    const double* p0 = pause + (size_t)(2 * u) * kPauseFields;
    const double* p1 = p0 + kPauseFields;
    const double w0 = (p0[7] != 0.0) ? p0[3] + 2.0 * p0[2] : 0.0, w1 = (p1[7] != 0.0) ? p1[3] + 2.0 * p1[2] : 0.0;
    keys[u] = (p0[7] != 0.0 || p1[7] != 0.0) ? (float)(w0 + w1) : -1.0f;
    idx[u] = u;
}

// Fast lane of the last launch.  With 16 warps per SM a tCG iteration of one pair takes 2.4 us instead of 1.4 (1.0 with
// one warp per scheduler), and the longest pairs of a batch (30-75 k iterations where the mean is 6 k) become its
// critical path.  By outer iteration 20 they are at the head of the sorted list (the truly longest pair ranked 2-14 in six
// batches; at 14 it can still rank 900th), with 80-90 % of their work ahead.  Their units are taken out of the main queue
// and run by the one-warp-per-copy kernel on a few SMs of their own, one warp per scheduler, next to the main kernel.
__global__ void hold_kernel(unsigned ns) {
    const uint64_t t0 = global_timer_ns();
    while (global_timer_ns() - t0 < ns) __nanosleep(1000);
}

__global__ void mark_fast_lane_kernel(const int* __restrict__ unit_order, int units, double* pause, int* fast_order) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= units) return;
    for (int r = 0; r < 2; ++r) {
        const int i = 2 * unit_order[j] + r;
        if (pause[(size_t)i * kPauseFields + 7] == 1.0) pause[(size_t)i * kPauseFields + 7] = 2.0;
        fast_order[2 * j + r] = i;   // finished pairs stay in the list and are skipped by the kernel
    }
}

// two warps per copy of S when the initial points of an instance come in pairs (see sphere_tmem2_kernel)
static bool sibling_units(const riptrm_handle* h) {
    if (h->n != 50 || h->batch_z < 1 || getenv("RIPTRM_SPHERE_NO_TMEM") != nullptr || getenv("RIPTRM_SPHERE_NO_SIBLINGS") != nullptr)
        return false;
    const int ipp = h->batch / h->batch_z;
    return ipp >= 2 && ipp % 2 == 0 && h->batch >= 2 * h->num_sms * 8;
}

// TRS_solver='Exact_RepMat': one warp per CTA with the representation-matrix workspace in shared memory, one launch
template <int MODE>
static int dispatch_sphere_exact(riptrm_handle* h, const SphereParams& P, const DevOpts& o, cudaStream_t st) {
#ifdef RIPTRM_DEV_SPHERE_ONLY   // diagnostic builds (seconds instead of minutes): only the n = 50 TMEM kernels
    return fail(RIPTRM_E_UNSUPPORTED, "RIPTRM_DEV_SPHERE_ONLY build");
#else
    if (h->n == 50) return launch_sphere<2, MODE, 50, true>(h, P, o, st);
    if (h->n <= 64) return launch_sphere<2, MODE, 0, true>(h, P, o, st);
    return fail(RIPTRM_E_UNSUPPORTED, "Exact_RepMat on Sphere(n): n <= 64");
#endif
}

template <int MODE>
static int dispatch_sphere(riptrm_handle* h, const SphereParams& P, const DevOpts& o, cudaStream_t st) {
    const int n = h->n;
    const bool no_tmem = getenv("RIPTRM_SPHERE_NO_TMEM") != nullptr;  // A/B switch (measurements, tests)
    if (MODE == 0 && P.sibling_units) return launch_sphere_tmem2(h, P, o, st);
    if (n == 50 && !no_tmem) return launch_sphere_tmem<MODE>(h, P, o, st);      // the reference's dim, S in TMEM
#ifdef RIPTRM_DEV_SPHERE_ONLY
    return fail(RIPTRM_E_UNSUPPORTED, "RIPTRM_DEV_SPHERE_ONLY build");
#else
    if (n == 50) return launch_sphere<2, MODE, 50>(h, P, o, st);  // the reference's dim (config_dataset.yaml:6)
    if (n <= 64) return launch_sphere<2, MODE, 0>(h, P, o, st);
    return launch_sphere<4, MODE, 0>(h, P, o, st);
#endif
}

static int finish_timing(riptrm_handle* h, bool sync) {
    if (sync) {
        CUDA_TRY(cudaEventSynchronize(h->ev1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        h->last_ms = ms;
    }
    return RIPTRM_OK;
}


// ------------------------------------------------------------------------------------------
// small-manifold families (Grassmann, Product[Skew, SPD, SPD]): one warp per pair, vectors on lanes, matrix
// products through a per-warp shared-memory scratch; same generic solve (solver_warp.cuh)
// ------------------------------------------------------------------------------------------
struct SmallParams {
    int n, p, m, batch;
    double alpha, offset;  // Rosenbrock
    const double *Xd, *XPd, *conspec;  // StableIdentification
    int N;
    double hstep;
    const double* x0;
    const double* y0;
    double* x;
    double* y;
    double* summary;
    double* trace;
    const double* v;
    double mu, Delta;
    double* out;
    double* info;
    const double* slack;           // riptrm_newton (see SphereParams)
    int newton_method, kr_maxiter;
    double kr_tol;
    int generic_tcg;               // StableIdentification: tCG on unwhitened vectors (measurement switch RIPTRM_STABLEID_GENERIC_TCG)
};

#ifndef RIPTRM_SMALL_MINBLOCKS
#define RIPTRM_SMALL_MINBLOCKS 16
#endif
template <class F, int MODE, bool EXACT = false>
__global__ void __launch_bounds__(32, EXACT ? 4 : RIPTRM_SMALL_MINBLOCKS) small_kernel(SmallParams P, DevOpts o, int* counter) {  // <= 128 registers: 16 warps per SM
    extern __shared__ __align__(16) double smem[];
    typename F::Ctx ctx = F::make_ctx(P, o, smem);
    const int lane = lane_id();
    const int xl = (F::kComponents == 1) ? P.n * P.p : P.n * P.n * F::kComponents;
    RepWork rw;
    if (EXACT) rep_init(rw, smem + F::smem_doubles(P.n, P.N), F::dim(ctx));
    while (true) {
        int inst = 0;
        if (lane == 0) inst = atomicAdd(counter, 1);
        inst = __shfl_sync(kFull, inst, 0);
        if (inst >= P.batch) break;
        const typename F::Vec x0 = F::load_x(ctx, P.x0 + (size_t)inst * xl);
        const typename F::CVec y0 = F::load_y(ctx, P.y0 + (size_t)inst * P.m);
        if (MODE == 0) {
            typename F::Pt pt;
            typename F::CVec y;
            double* tr = (P.trace != nullptr && o.trace_mode != 0)
                             ? P.trace + (size_t)inst * o.trace_capacity * RIPTRM_TRACE_FIELDS
                             : nullptr;
            if constexpr (EXACT) {
                rw.cur = 0;
                rw.mat_valid = rw.al_valid = false;
                solve_instance<F, true, RepWork>(ctx, o, x0, y0, pt, y,
                                                 P.summary ? P.summary + (size_t)inst * RIPTRM_SUMMARY_FIELDS : nullptr, tr, nullptr,
                                                 false, -1, &rw);
            } else {
                solve_instance<F>(ctx, o, x0, y0, pt, y, P.summary ? P.summary + (size_t)inst * RIPTRM_SUMMARY_FIELDS : nullptr,
                                  tr, nullptr, false, -1);
            }
            if (P.x) F::store_x(ctx, P.x + (size_t)inst * xl, pt.x);
            if (P.y) F::store_y(ctx, P.y + (size_t)inst * P.m, y);
        } else {
            typename F::Pt pt;
            F::eval_point(ctx, x0, pt);
            typename F::Step st;
            F::begin_step(ctx, pt, y0, P.mu, st);
            if (MODE == 1) {
                const typename F::Vec v = F::load_x(ctx, P.v + (size_t)inst * xl);
                const typename F::Vec hv = F::Hw(ctx, pt, y0, st, v);
                F::store_x(ctx, P.out + (size_t)inst * xl, hv);
            } else if (MODE == 3) {
                if constexpr (EXACT) {
                    typename F::Vec dx;
                    trs_hook<F>(ctx, o, pt, y0, st, P.Delta, rw, dx, P.info ? P.info + (size_t)inst * 4 : nullptr);
                    F::store_x(ctx, P.out + (size_t)inst * xl, dx);
                }
            } else if (MODE == 4) {
                if constexpr (EXACT) {
                    typename F::Vec dx;
                    const typename F::CVec sl = F::load_y(ctx, P.slack + (size_t)inst * P.m);
                    const typename F::Vec rhs = F::load_x(ctx, P.v + (size_t)inst * xl);
                    newton_hook<F>(ctx, pt, y0, st, sl, rhs, P.newton_method, P.kr_tol, P.kr_maxiter, rw, dx,
                                   P.info ? P.info + (size_t)inst * 4 : nullptr);
                    F::store_x(ctx, P.out + (size_t)inst * xl, dx);
                }
            } else {
                typename F::Vec eta, Heta;
                const TcgResult r = F::tcg(ctx, o, pt, y0, st, P.Delta, eta, Heta);
                F::store_x(ctx, P.out + (size_t)inst * xl, eta);
                const double nrm = sqrt(F::inner(ctx, pt, eta, eta));
                if (P.info != nullptr && lane < 4) {
                    const double val = (lane == 0) ? (double)r.iters : (lane == 1) ? (double)r.stop : (lane == 2) ? nrm : r.model_value;
                    P.info[(size_t)inst * 4 + lane] = val;
                }
            }
        }
    }
}

template <class F>
static int small_dim(const SmallParams& P);
template <>
int small_dim<GrassmannFam>(const SmallParams& P) { return P.n * P.p - P.p * P.p; }
template <>
int small_dim<StableIdFam>(const SmallParams& P) { return P.n * (P.n - 1) / 2 + P.n * (P.n + 1); }

template <class F, int MODE, bool EXACT = false>
static int launch_small(riptrm_handle* h, const SmallParams& P, const DevOpts& o, cudaStream_t st) {
    auto kern = small_kernel<F, MODE, EXACT>;
    const size_t smem = ((size_t)F::smem_doubles(P.n, P.N) + (EXACT ? rep_doubles(small_dim<F>(P)) : 0)) * sizeof(double);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // one warp per CTA, as many CTAs per SM as registers / shared memory allow (the warps are latency-bound)
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
    if (per_sm < 1) return fail(RIPTRM_E_UNSUPPORTED, "small-family kernel does not fit on an SM");
    if (const char* e = getenv("RIPTRM_SMALL_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e)));  // diagnostic
    int grid = h->num_sms * per_sm;
    if (grid > h->batch) grid = h->batch;
    CUDA_TRY(cudaMemsetAsync(h->d_counter, 0, sizeof(int), st));
    CUDA_TRY(cudaEventRecord(h->ev0, st));
    kern<<<grid, 32, smem, st>>>(P, o, h->d_counter);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ev1, st));
    h->launches += 1;
    return RIPTRM_OK;
}

static int dispatch_small(riptrm_handle* h, int mode, const SmallParams& P, const DevOpts& o, cudaStream_t st) {
#ifdef RIPTRM_DEV_SPHERE_ONLY   // diagnostic builds (seconds instead of minutes): only the n = 50 TMEM kernels
    return fail(RIPTRM_E_UNSUPPORTED, "RIPTRM_DEV_SPHERE_ONLY build");
#else
    if (mode == 3 || mode == 4 || (mode == 0 && exact_repmat(h))) {   // exact trust-region solver (hook / whole solve), Newton hook
        if (h->family == RIPTRM_FAMILY_ROSENBROCK_GRASSMANN)
            return mode == 3 ? launch_small<GrassmannFam, 3, true>(h, P, o, st)
                 : mode == 4 ? launch_small<GrassmannFam, 4, true>(h, P, o, st) : launch_small<GrassmannFam, 0, true>(h, P, o, st);
        if (h->family == RIPTRM_FAMILY_STABLEID_PRODUCT)
            return mode == 3 ? launch_small<StableIdFam, 3, true>(h, P, o, st)
                 : mode == 4 ? launch_small<StableIdFam, 4, true>(h, P, o, st) : launch_small<StableIdFam, 0, true>(h, P, o, st);
        return fail(RIPTRM_E_UNSUPPORTED, "family not built into this library");
    }
    if (h->family == RIPTRM_FAMILY_ROSENBROCK_GRASSMANN) {
        if (mode == 0) return launch_small<GrassmannFam, 0>(h, P, o, st);
        if (mode == 1) return launch_small<GrassmannFam, 1>(h, P, o, st);
        return launch_small<GrassmannFam, 2>(h, P, o, st);
    }
    if (h->family == RIPTRM_FAMILY_STABLEID_PRODUCT) {
        if (mode == 0) return launch_small<StableIdFam, 0>(h, P, o, st);
        if (mode == 1) return launch_small<StableIdFam, 1>(h, P, o, st);
        return launch_small<StableIdFam, 2>(h, P, o, st);
    }
    return fail(RIPTRM_E_UNSUPPORTED, "family not built into this library");
#endif
}

static SmallParams small_params(const riptrm_handle* h) {
    SmallParams P{};
    P.n = h->n;
    P.p = h->p;
    P.m = h->m;
    P.batch = h->batch;
    P.alpha = h->ros_alpha;
    P.offset = h->ros_offset;
    P.generic_tcg = getenv("RIPTRM_STABLEID_GENERIC_TCG") != nullptr ? 1 : 0;
    if (h->d_sid != nullptr) {
        const size_t dn = (size_t)h->n * h->sid_N;
        P.Xd = h->d_sid;
        P.XPd = h->d_sid + dn;
        P.conspec = h->d_sid + 2 * dn;
        P.N = h->sid_N;
        P.hstep = h->sid_h;
    }
    return P;
}

// election state + placement records (main-kernel CTAs in [0, 2 sms), lane-kernel CTAs of the two-kernel form after them)
static int ensure_lane_buffers(riptrm_handle* h, cudaStream_t st) {
    const int records = h->num_sms * 2 + 64, total = 2 + 2 * kLaneMaxSms + records;
    if (h->d_lane_state == nullptr || h->lane_debug_len < records) {
        free_any(h->d_lane_state);
        free_any(h->d_lane_times);
        CUDA_TRY(cudaMalloc(&h->d_lane_state, (size_t)total * sizeof(int)));
        CUDA_TRY(cudaMalloc(&h->d_lane_times, (size_t)2 * records * sizeof(unsigned long long)));
        h->lane_debug_len = records;
    }
    CUDA_TRY(cudaMemsetAsync(h->d_lane_state, 0, (size_t)total * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(h->d_lane_times, 0, (size_t)2 * records * sizeof(unsigned long long), st));
    return RIPTRM_OK;
}

// Fast lane of the last launch, carved out of the main kernel's grid (see sphere_tmem2_kernel): marks the pairs of the
// `fast_units` longest units for the lane, resets the election state and fills the lane fields of P.
static int prepare_lane_in_kernel(riptrm_handle* h, SphereParams& P, cudaStream_t st) {
    int fast_units = 16;   // 32 pairs: 8 lane CTAs x 4 warps
    if (const char* e = getenv("RIPTRM_FAST_UNITS")) fast_units = std::max(1, std::min(64, atoi(e)));   // tuning knob
    fast_units = std::min(fast_units, h->batch / 2);
    if (h->d_fast_order == nullptr) CUDA_TRY(cudaMalloc(&h->d_fast_order, 2 * 64 * sizeof(int)));
    int rc0;
    if ((rc0 = ensure_lane_buffers(h, st))) return rc0;
    mark_fast_lane_kernel<<<(fast_units + 31) / 32, 32, 0, st>>>(h->d_order, fast_units, h->d_pause, h->d_fast_order);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemsetAsync(h->d_counter + 1, 0, sizeof(int), st));
    P.lane_ctas = (2 * fast_units + 3) / 4;
    P.fast_len = 2 * fast_units;
    P.fast_order = h->d_fast_order;
    P.lane_state = h->d_lane_state;
    P.lane_debug = h->d_lane_state + 2 + 2 * kLaneMaxSms;
    P.lane_times = h->d_lane_times;
    h->launches += 1;
    return RIPTRM_OK;
}

typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static WaitValue32Fn lane_wait_value_entry() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return reinterpret_cast<WaitValue32Fn>(fn);
}

// the two-kernel lane: a second kernel on a priority stream whose CTAs take an SM each (150 KB of shared memory: no main CTA
// fits beside them); the main kernel's stream waits until they have all started
static int launch_fast_lane(riptrm_handle* h, const SphereParams& P, const DevOpts& o, cudaStream_t st) {
    int fast_units = 16;   // 32 pairs: 8 CTAs of the 4-warp kernel on 8 SMs of their own
    if (const char* e = getenv("RIPTRM_FAST_UNITS")) fast_units = std::max(1, std::min(64, atoi(e)));   // tuning knob
    fast_units = std::min(fast_units, h->batch / 2);
    if (h->lane_stream == nullptr) {
        int lo = 0, hi = 0;   // highest priority: its CTAs are placed before the main kernel's when both are pending
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&h->lane_stream, cudaStreamNonBlocking, hi));
        CUDA_TRY(cudaEventCreateWithFlags(&h->lane_ev0, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->lane_ev1, cudaEventDisableTiming));
        CUDA_TRY(cudaMalloc(&h->d_fast_order, 2 * 64 * sizeof(int)));
    }
    mark_fast_lane_kernel<<<(fast_units + 31) / 32, 32, 0, st>>>(h->d_order, fast_units, h->d_pause, h->d_fast_order);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemsetAsync(h->d_counter + 1, 0, sizeof(int), st));
    if (h->d_lane_arrive == nullptr) CUDA_TRY(cudaMalloc(&h->d_lane_arrive, sizeof(unsigned int)));
    CUDA_TRY(cudaMemsetAsync(h->d_lane_arrive, 0, sizeof(unsigned int), st));
    CUDA_TRY(cudaEventRecord(h->lane_ev0, st));
    CUDA_TRY(cudaStreamWaitEvent(h->lane_stream, h->lane_ev0, 0));
    SphereParams F = P;
    F.lane_arrive = h->d_lane_arrive;
    F.sibling_units = 0;
    F.order = h->d_fast_order;
    F.queue_len = 2 * fast_units;
    F.lane_debug = (h->d_lane_state != nullptr) ? h->d_lane_state + 2 + 2 * kLaneMaxSms + h->num_sms * 2 : nullptr;
    F.lane_times = (h->d_lane_times != nullptr) ? h->d_lane_times + 2 * (h->num_sms * 2) : nullptr;
    // 150 KB of shared memory per CTA: no CTA of the main kernel (86 KB) fits next to it
    const size_t smem = 150 * 1024;
    auto kern = sphere_tmem_kernel<0>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const int nctas = (2 * fast_units + 3) / 4;
    kern<<<nctas, 128, smem, h->lane_stream>>>(F, o, h->d_counter + 1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->lane_ev1, h->lane_stream));
    // The main kernel follows on `st` once the lane's CTAs have their SMs -- BY CONSTRUCTION: `st` waits until the arrival
    // count the lane CTAs increment at entry has reached their number (a stream memory operation; round 1 slept 30 us in a
    // one-warp kernel instead).  Driver entry point through the runtime, so the library does not link libcuda; without it
    // (or with RIPTRM_LANE_HOLD_BY_SLEEP=1, an A/B switch) the sleep kernel is the fall-back.
    static WaitValue32Fn wait_value = lane_wait_value_entry();
    if (wait_value != nullptr && getenv("RIPTRM_LANE_HOLD_BY_SLEEP") == nullptr) {
        const CUresult wr = wait_value((CUstream)st, (CUdeviceptr)(uintptr_t)h->d_lane_arrive, (cuuint32_t)nctas, CU_STREAM_WAIT_VALUE_GEQ);
        if (wr != CUDA_SUCCESS) return fail(RIPTRM_E_CUDA, "cuStreamWaitValue32 failed");
        h->launches += 2;
    } else {
        hold_kernel<<<1, 32, 0, st>>>(30000);
        CUDA_TRY(cudaGetLastError());
        h->launches += 3;
    }
    return RIPTRM_OK;
}

// One launch, or several with the pairs re-ordered in between (see solve_instance): every pair is advanced to outer
// iteration s1, then -- longest first, by the work spent so far -- to s2, then to the end.  Bit-identical results either
// way.  Work per pair varies 3-5x and is only partly predictable from its first iterations: on six 16384-pair batches of
// the bench workload the makespan over the ideal (total work / resident warps) was 1.34 / 1.22 / ... for one launch in
// index order, 1.06-1.23 for one split at 6, 1.04-1.08 at 8, 1.03-1.04 for splits at 8 and 14, 1.005 with perfect
// knowledge (scripts/schedule_probe.py + offline list scheduling of the measured per-pair work).
static int solve_scheduled(riptrm_handle* h, SphereParams P, const DevOpts& o, cudaStream_t st) {
    if (exact_repmat(h)) {
        P.order = nullptr;
        P.pause = nullptr;
        P.resume = 0;
        P.pause_at = -1;
        return dispatch_sphere_exact<0>(h, P, o, st);
    }
    const int user = h->opts.schedule_split;  //: 0 auto, < 0 off, > 0 one split at that outer iteration
    const int maxiter = h->opts.maxiter;
    const int resident = h->num_sms * 10;
    P.sibling_units = sibling_units(h) ? 1 : 0;
    int splits[3] = {-1, -1, -1};
    int nsplit = 0;
    if (user > 0 && user < maxiter) {
        splits[nsplit++] = user;
    } else if (user == 0 && h->batch > resident && maxiter > 12) {
        splits[nsplit++] = (4 * maxiter + 7) / 15;   // 8 of 30
        splits[nsplit++] = (7 * maxiter + 7) / 15;   // 14 of 30
        if (P.sibling_units) splits[nsplit++] = (2 * maxiter) / 3;   // 20 of 30: the fast lane needs the late ranking
        if (const char* e = getenv("RIPTRM_SPLITS")) {   // tuning knob: up to three comma-separated outer iterations
            nsplit = 0;
            for (const char* c = e; *c != 0 && nsplit < 3;) {
                const int v = atoi(c);
                if (v > 0 && v < maxiter) splits[nsplit++] = v;
                while (*c != 0 && *c != ',') ++c;
                if (*c == ',') ++c;
            }
        }
    }
    if (nsplit == 0) {
        P.order = nullptr;
        P.pause = nullptr;
        P.resume = 0;
        P.pause_at = -1;
        return dispatch_sphere<0>(h, P, o, st);
    }
    const size_t B = h->batch;
    int rc;
    if ((rc = ensure(h->d_pause, B * kPauseFields * sizeof(double)))) return rc;
    if (h->d_keys == nullptr) {
        CUDA_TRY(cudaMalloc(&h->d_keys, B * sizeof(float)));
        CUDA_TRY(cudaMalloc(&h->d_keys_sorted, B * sizeof(float)));
        CUDA_TRY(cudaMalloc(&h->d_idx, B * sizeof(int)));
        CUDA_TRY(cudaMalloc(&h->d_order, B * sizeof(int)));
        CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(nullptr, h->sort_tmp_bytes, h->d_keys, h->d_keys_sorted, h->d_idx,
                                                           h->d_order, (int)B, 0, 32, st));
        CUDA_TRY(cudaMalloc(&h->d_sort_tmp, h->sort_tmp_bytes));
    }
    // the paused iterates travel in the x / y output buffers
    if (P.x == nullptr) {
        if ((rc = ensure(h->d_x, B * h->vec_len * sizeof(double)))) return rc;
        P.x = h->d_x;
    }
    if (P.y == nullptr) {
        if ((rc = ensure(h->d_y, B * h->m * sizeof(double)))) return rc;
        P.y = h->d_y;
    }
    cudaEvent_t first_start = nullptr;
    bool swapped = false;
    CUDA_TRY(cudaEventCreateWithFlags(&first_start, cudaEventDefault));
    P.pause = h->d_pause;
    for (int phase = 0; phase <= nsplit; ++phase) {
        P.order = (phase == 0) ? nullptr : h->d_order;
        P.resume = (phase == 0) ? 0 : 1;
        P.pause_at = (phase < nsplit) ? splits[phase] : -1;
        const bool lane = P.sibling_units && user == 0 && phase == nsplit && nsplit >= 1 && splits[nsplit - 1] * 3 >= maxiter * 2 - 2 &&
                          getenv("RIPTRM_SPHERE_NO_FAST_LANE") == nullptr;   // needs the ranking of outer iteration ~ 2/3 maxiter
        // two forms of the lane: a second kernel on a priority stream (default: its one-warp-per-copy
        // code at 254 registers is 2-4 % faster end to end) or CTAs elected inside the main kernel (RIPTRM_FAST_LANE_IN_KERNEL=1)
        const bool legacy_lane = getenv("RIPTRM_FAST_LANE_IN_KERNEL") == nullptr;
        if (getenv("RIPTRM_DEBUG_LANE") != nullptr)
            fprintf(stderr, "[riptrm] phase %d/%d lane %d sibling %d user %d last split %d maxiter %d\n", phase, nsplit, (int)lane,
                    P.sibling_units, user, nsplit ? splits[nsplit - 1] : -1, maxiter);
        const bool in_kernel = lane && !legacy_lane && lane_in_kernel_possible(h);
        P.lane_ctas = 0;
        if (in_kernel && (rc = prepare_lane_in_kernel(h, P, st))) break;
        if (lane && !in_kernel) {
            if ((rc = ensure_lane_buffers(h, st))) break;
            if ((rc = launch_fast_lane(h, P, o, st))) break;
            P.lane_debug = h->d_lane_state + 2 + 2 * kLaneMaxSms;
            P.lane_times = h->d_lane_times;
        }
        if ((rc = dispatch_sphere<0>(h, P, o, st))) break;
        P.lane_ctas = 0;
        P.lane_debug = nullptr;
        P.lane_times = nullptr;
        if (lane && !in_kernel && cudaStreamWaitEvent(st, h->lane_ev1, 0) != cudaSuccess) { rc = fail(RIPTRM_E_CUDA, "fast lane join failed"); break; }
        if (lane && !in_kernel) cudaEventRecord(h->ev1, st);   // the reported time ends when both kernels have
        if (phase == 0) {   // keep the start of the first launch: the reported time spans all
            std::swap(first_start, h->ev0);
            swapped = true;
        }
        if (phase < nsplit) {
            const int items = P.sibling_units ? (int)(B / 2) : (int)B;
            if (P.sibling_units)
                schedule_unit_keys_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(h->d_pause, h->d_keys, h->d_idx, items);
            else
                schedule_keys_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(h->d_pause, h->d_keys, h->d_idx, items);
            if (cudaGetLastError() != cudaSuccess) { rc = fail(RIPTRM_E_CUDA, "schedule_keys_kernel launch failed"); break; }
            if (cub::DeviceRadixSort::SortPairsDescending(h->d_sort_tmp, h->sort_tmp_bytes, h->d_keys, h->d_keys_sorted,
                                                          h->d_idx, h->d_order, items, 0, 32, st) != cudaSuccess) {
                rc = fail(RIPTRM_E_CUDA, "radix sort of the schedule keys failed");
                break;
            }
            h->launches += 2;
        }
    }
    if (swapped) std::swap(first_start, h->ev0);   // ev0 = start of the first launch again; ev1 was recorded by the last one
    cudaEventDestroy(first_start);
    return rc;
}

extern "C" int riptrm_generate_nonnegpca(int device, int n, long long first_instance, int instances, int points_per_instance,
                                         double snr, double delta, double* Z, double* x0, double* y0, void* stream) {
    if (Z == nullptr || x0 == nullptr || y0 == nullptr) return fail(RIPTRM_E_INVALID, "NULL argument");
    if (n < 2 || n > 2048 || instances < 1 || points_per_instance < 1 || first_instance < 0)
        return fail(RIPTRM_E_INVALID, "generate_nonnegpca: 2 <= n <= 2048, instances >= 1, points_per_instance >= 1");
    if (!(snr >= 0.0) || !(delta > 0.0 && delta <= 1.0) || (int)floor(delta * n) < 1)
        return fail(RIPTRM_E_INVALID, "generate_nonnegpca: snr >= 0, 0 < delta <= 1, floor(delta n) >= 1");
    CUDA_TRY(cudaSetDevice(device));
    const size_t smem = (size_t)n * (2 * sizeof(double) + sizeof(int));
    gen::nonnegpca_kernel<<<instances, 128, smem, (cudaStream_t)stream>>>(n, first_instance, points_per_instance, snr, delta, Z, x0, y0);
    CUDA_TRY(cudaGetLastError());
    return RIPTRM_OK;
}

extern "C" int riptrm_solve(riptrm_handle* h, const double* x0, const double* y0, double* x, double* y,
                            double* summary, double* trace, int where, void* stream) {
    if (h == nullptr || x0 == nullptr || y0 == nullptr) return fail(RIPTRM_E_INVALID, "NULL argument");
    if (!h->have_problem) return fail(RIPTRM_E_STATE, "riptrm_set_<family> has not been called");
    if (!h->have_opts) return fail(RIPTRM_E_STATE, "riptrm_set_options has not been called");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (h->family == RIPTRM_FAMILY_NONNEGPCA_COLUMNS || is_stiefel(h)) return columns_solve(h, x0, y0, x, y, summary, trace, where, st);
    const size_t B = h->batch;
    const size_t xb = B * h->vec_len * sizeof(double), yb = B * h->m * sizeof(double);
    const size_t sb = B * RIPTRM_SUMMARY_FIELDS * sizeof(double);
    const size_t tb = (h->opts.trace_mode != 0) ? B * (size_t)h->opts.trace_capacity * RIPTRM_TRACE_FIELDS * sizeof(double) : 0;
    const bool small = (h->family == RIPTRM_FAMILY_ROSENBROCK_GRASSMANN || h->family == RIPTRM_FAMILY_STABLEID_PRODUCT);
    SphereParams P{};
    P.Z = h->dZ;
    P.batch_z = h->batch_z;
    P.n = h->n;
    P.batch = h->batch;
    P.eps = h->eps;
    const DevOpts o = make_devopts(h);
    if (small && where == RIPTRM_DEVICE) {
        SmallParams Q = small_params(h);
        Q.x0 = x0; Q.y0 = y0; Q.x = x; Q.y = y; Q.summary = summary; Q.trace = trace;
        return dispatch_small(h, 0, Q, o, st);
    }
    if (where == RIPTRM_DEVICE) {
        P.x0 = x0;
        P.y0 = y0;
        P.x = x;
        P.y = y;
        P.summary = summary;
        P.trace = trace;
        return solve_scheduled(h, P, o, st);
    }
    int rc;
    if ((rc = ensure(h->d_x0, xb)) || (rc = ensure(h->d_y0, yb)) || (rc = ensure(h->d_x, xb)) ||
        (rc = ensure(h->d_y, yb)) || (rc = ensure(h->d_summary, sb)))
        return rc;
    if (tb != 0 && trace != nullptr) {
        if (h->d_trace != nullptr && h->trace_bytes != tb) free_dev(h->d_trace);
        if ((rc = ensure(h->d_trace, tb))) return rc;
        h->trace_bytes = tb;
    }
    CUDA_TRY(cudaMemcpyAsync(h->d_x0, x0, xb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->d_y0, y0, yb, cudaMemcpyHostToDevice, st));
    P.x0 = h->d_x0;
    P.y0 = h->d_y0;
    P.x = h->d_x;
    P.y = h->d_y;
    P.summary = h->d_summary;
    P.trace = (tb != 0 && trace != nullptr) ? h->d_trace : nullptr;
    if (small) {
        SmallParams Q = small_params(h);
        Q.x0 = P.x0; Q.y0 = P.y0; Q.x = P.x; Q.y = P.y; Q.summary = P.summary; Q.trace = P.trace;
        if ((rc = dispatch_small(h, 0, Q, o, st))) return rc;
    } else if ((rc = solve_scheduled(h, P, o, st))) {
        return rc;
    }
    if (x) CUDA_TRY(cudaMemcpyAsync(x, h->d_x, xb, cudaMemcpyDeviceToHost, st));
    if (y) CUDA_TRY(cudaMemcpyAsync(y, h->d_y, yb, cudaMemcpyDeviceToHost, st));
    if (summary) CUDA_TRY(cudaMemcpyAsync(summary, h->d_summary, sb, cudaMemcpyDeviceToHost, st));
    if (P.trace) CUDA_TRY(cudaMemcpyAsync(trace, h->d_trace, tb, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return finish_timing(h, true);
}

static int run_hook(riptrm_handle* h, int mode, const double* x, const double* y, double mu, double Delta,
                    const double* v, double* out, double* info, int where, cudaStream_t st) {
    if (h == nullptr || x == nullptr || y == nullptr || out == nullptr) return fail(RIPTRM_E_INVALID, "NULL argument");
    if (!h->have_problem) return fail(RIPTRM_E_STATE, "riptrm_set_<family> has not been called");
    if (mode == 2 && !h->have_opts) return fail(RIPTRM_E_STATE, "riptrm_set_options has not been called");
    CUDA_TRY(cudaSetDevice(h->device));
    if (mode == 3 && (h->family == RIPTRM_FAMILY_NONNEGPCA_COLUMNS || is_stiefel(h)))
        return fail(RIPTRM_E_UNSUPPORTED, "riptrm_trs: not built for the large-n families");
    if (h->family == RIPTRM_FAMILY_NONNEGPCA_COLUMNS || is_stiefel(h)) return columns_run(h, mode, x, y, mu, Delta, v, out, info, where, st);
    const size_t B = h->batch;
    const size_t xb = B * h->vec_len * sizeof(double), yb = B * h->m * sizeof(double), ib = B * 4 * sizeof(double);
    SphereParams P{};
    P.Z = h->dZ;
    P.batch_z = h->batch_z;
    P.n = h->n;
    P.batch = h->batch;
    P.eps = h->eps;
    P.mu = mu;
    P.Delta = Delta;
    DevOpts o{};
    if (h->have_opts) o = make_devopts(h);
    else { o.tcg_maxinner = -1; o.tcg_theta = 1.0; o.tcg_kappa = 0.1; o.tcg_mininner = 1; o.trs_tolhardcase = 1e-8; }
    int rc;
    const bool small = (h->family == RIPTRM_FAMILY_ROSENBROCK_GRASSMANN || h->family == RIPTRM_FAMILY_STABLEID_PRODUCT);
    if (where == RIPTRM_DEVICE) {
        if (small) {
            SmallParams Q = small_params(h);
            Q.x0 = x; Q.y0 = y; Q.v = v; Q.out = out; Q.info = info; Q.mu = mu; Q.Delta = Delta;
            return dispatch_small(h, mode, Q, o, st);
        }
        P.x0 = x; P.y0 = y; P.v = v; P.out = out; P.info = info;
        if (mode == 3) return dispatch_sphere_exact<3>(h, P, o, st);
        return mode == 1 ? dispatch_sphere<1>(h, P, o, st) : dispatch_sphere<2>(h, P, o, st);
    }
    if ((rc = ensure(h->d_x0, xb)) || (rc = ensure(h->d_y0, yb)) || (rc = ensure(h->d_x, xb)) ||
        (rc = ensure(h->d_v, xb)) || (rc = ensure(h->d_info, ib)))
        return rc;
    CUDA_TRY(cudaMemcpyAsync(h->d_x0, x, xb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->d_y0, y, yb, cudaMemcpyHostToDevice, st));
    if (v) CUDA_TRY(cudaMemcpyAsync(h->d_v, v, xb, cudaMemcpyHostToDevice, st));
    P.x0 = h->d_x0; P.y0 = h->d_y0; P.v = h->d_v; P.out = h->d_x; P.info = info ? h->d_info : nullptr;
    if (small) {
        SmallParams Q = small_params(h);
        Q.x0 = P.x0; Q.y0 = P.y0; Q.v = P.v; Q.out = P.out; Q.info = P.info; Q.mu = mu; Q.Delta = Delta;
        rc = dispatch_small(h, mode, Q, o, st);
    } else if (mode == 3)
        rc = dispatch_sphere_exact<3>(h, P, o, st);
    else
        rc = mode == 1 ? dispatch_sphere<1>(h, P, o, st) : dispatch_sphere<2>(h, P, o, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, h->d_x, xb, cudaMemcpyDeviceToHost, st));
    if (info) CUDA_TRY(cudaMemcpyAsync(info, h->d_info, ib, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return finish_timing(h, true);
}

extern "C" int riptrm_hessvec(riptrm_handle* h, const double* x, const double* y, double mu, const double* v,
                              double* out, int where, void* stream) {
    if (v == nullptr) return fail(RIPTRM_E_INVALID, "NULL argument");
    return run_hook(h, 1, x, y, mu, 0.0, v, out, nullptr, where, (cudaStream_t)stream);
}

extern "C" int riptrm_tcg(riptrm_handle* h, const double* x, const double* y, double mu, double Delta, double* eta,
                          double* info, int where, void* stream) {
    return run_hook(h, 2, x, y, mu, Delta, nullptr, eta, info, where, (cudaStream_t)stream);
}

// Batched dense trust-region subproblems, one warp each: A [count][d][d] symmetric, a [count][d] -> x [count][d],
// info [count][4] = {type, lam1, |x|, smallest eigenvalue of A}
__global__ void __launch_bounds__(32) trs_dense_kernel(const double* __restrict__ A, const double* __restrict__ a, int d, int count,
                                                      double Delta, double tolhard, double* x, double* info) {
    extern __shared__ __align__(16) double smem[];
    RepWork rw;
    rep_init(rw, smem, d);
    const int lane = lane_id();
    for (int b = blockIdx.x; b < count; b += gridDim.x) {
        const double* Ab = A + (size_t)b * d * d;
        for (int e = lane; e < d * d; e += 32) rw.W[(e / d) * rw.ld + (e % d)] = Ab[e];
        for (int k = lane; k < d; k += 32) rw.col[k] = a[(size_t)b * d + k];
        __syncwarp();
        dense::jacobi_sym(rw.W, rw.VT[0], d, rw.ld);
        for (int k = lane; k < d; k += 32) rw.D[0][k] = rw.W[k * rw.ld + k];
        __syncwarp();
        dense::rows_dot(rw.VT[0], d, rw.ld, rw.col, rw.al[0]);
        const dense::TrsOut to = dense::trs_eig(rw.D[0], rw.al[0], d, Delta, tolhard, rw.col, rw.ws, rw.W, rw.ld);
        dense::cols_dot(rw.VT[0], d, rw.ld, rw.col, rw.coef);
        const double nrm = sqrt(dense::vdot(rw.coef, rw.coef, d));
        const double mineig = rep_mineig(rw, 0);
        for (int k = lane; k < d; k += 32) x[(size_t)b * d + k] = rw.coef[k];
        if (lane < 4) info[(size_t)b * 4 + lane] = (lane == 0) ? (double)to.kind : (lane == 1) ? to.lam1 : (lane == 2) ? nrm : mineig;
        __syncwarp();
    }
}

extern "C" int riptrm_trs_dense(int device, int d, int count, const double* A, const double* a, double Delta, double tolhardcase,
                                double* x, double* info, int where, void* stream) {
    if (A == nullptr || a == nullptr || x == nullptr || info == nullptr) return fail(RIPTRM_E_INVALID, "NULL argument");
    if (d < 1 || d > 64 || count < 1 || !(Delta > 0.0)) return fail(RIPTRM_E_INVALID, "trs_dense: 1 <= d <= 64, count >= 1, Delta > 0");
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ab = (size_t)count * d * d * sizeof(double), vb = (size_t)count * d * sizeof(double), ib = (size_t)count * 4 * sizeof(double);
    double *dA = const_cast<double*>(A), *da = const_cast<double*>(a), *dx = x, *di = info;
    if (where != RIPTRM_DEVICE) {
        CUDA_TRY(cudaMalloc(&dA, ab + 2 * vb + ib));
        da = dA + (size_t)count * d * d;
        dx = da + (size_t)count * d;
        di = dx + (size_t)count * d;
        CUDA_TRY(cudaMemcpyAsync(dA, A, ab, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(da, a, vb, cudaMemcpyHostToDevice, st));
    }
    const size_t smem = rep_doubles(d) * sizeof(double);
    CUDA_TRY(cudaFuncSetAttribute(trs_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    trs_dense_kernel<<<count < 1184 ? count : 1184, 32, smem, st>>>(dA, da, d, count, Delta, tolhardcase, dx, di);
    CUDA_TRY(cudaGetLastError());
    if (where != RIPTRM_DEVICE) {
        CUDA_TRY(cudaMemcpyAsync(x, dx, vb, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(info, di, ib, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        cudaFree(dA);
    }
    return RIPTRM_OK;
}

// RIPM's condensed Newton system on the same Hessian-vector operator (RIPM.py:484-511)
extern "C" int riptrm_newton(riptrm_handle* h, const double* x, const double* z, const double* s, const double* c, int method,
                             double tol, int maxiter, double* dx, double* info, int where, void* stream) {
    if (h == nullptr || x == nullptr || z == nullptr || s == nullptr || c == nullptr || dx == nullptr)
        return fail(RIPTRM_E_INVALID, "NULL argument");
    if (method != 0 && method != 1) return fail(RIPTRM_E_INVALID, "riptrm_newton: method 0 (RepresentMatMethod) or 1 (TangentSpaceConjResMethod)");
    if (method == 1 && (!(tol >= 0.0) || maxiter < 1)) return fail(RIPTRM_E_INVALID, "riptrm_newton: tol >= 0, maxiter >= 1");
    if (!h->have_problem) return fail(RIPTRM_E_STATE, "riptrm_set_<family> has not been called");
    if (h->family == RIPTRM_FAMILY_NONNEGPCA_COLUMNS || is_stiefel(h))
        return fail(RIPTRM_E_UNSUPPORTED, "riptrm_newton: not built for the large-n families");
    if (h->family == RIPTRM_FAMILY_NONNEGPCA_SPHERE && h->n > 64) return fail(RIPTRM_E_UNSUPPORTED, "riptrm_newton on Sphere(n): n <= 64");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t B = h->batch;
    const size_t xb = B * h->vec_len * sizeof(double), yb = B * h->m * sizeof(double), ib = B * 4 * sizeof(double);
    DevOpts o{};
    if (h->have_opts) o = make_devopts(h);
    const bool small = (h->family == RIPTRM_FAMILY_ROSENBROCK_GRASSMANN || h->family == RIPTRM_FAMILY_STABLEID_PRODUCT);
    const double *px = x, *pz = z, *ps = s, *pc = c;
    double *pdx = dx, *pinfo = info;
    double* stage = nullptr;
    if (where != RIPTRM_DEVICE) {
        CUDA_TRY(cudaMalloc(&stage, 3 * xb + 2 * yb + ib));
        double* q = stage;
        CUDA_TRY(cudaMemcpyAsync(q, x, xb, cudaMemcpyHostToDevice, st)); px = q; q += B * h->vec_len;
        CUDA_TRY(cudaMemcpyAsync(q, c, xb, cudaMemcpyHostToDevice, st)); pc = q; q += B * h->vec_len;
        pdx = q; q += B * h->vec_len;
        CUDA_TRY(cudaMemcpyAsync(q, z, yb, cudaMemcpyHostToDevice, st)); pz = q; q += B * h->m;
        CUDA_TRY(cudaMemcpyAsync(q, s, yb, cudaMemcpyHostToDevice, st)); ps = q; q += B * h->m;
        pinfo = q;
    }
    int rc;
    if (small) {
        SmallParams Q = small_params(h);
        Q.x0 = px; Q.y0 = pz; Q.v = pc; Q.out = pdx; Q.info = pinfo; Q.slack = ps;
        Q.newton_method = method; Q.kr_tol = tol; Q.kr_maxiter = maxiter;
        rc = dispatch_small(h, 4, Q, o, st);
    } else {
        SphereParams P{};
        P.Z = h->dZ; P.batch_z = h->batch_z; P.n = h->n; P.batch = h->batch; P.eps = h->eps;
        P.x0 = px; P.y0 = pz; P.v = pc; P.out = pdx; P.info = pinfo; P.slack = ps;
        P.newton_method = method; P.kr_tol = tol; P.kr_maxiter = maxiter;
        rc = dispatch_sphere_exact<4>(h, P, o, st);
    }
    if (rc == RIPTRM_OK && where != RIPTRM_DEVICE) {
        cudaError_t e = cudaMemcpyAsync(dx, pdx, xb, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && info != nullptr) e = cudaMemcpyAsync(info, pinfo, ib, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(RIPTRM_E_CUDA, std::string("riptrm_newton: ") + cudaGetErrorString(e));
        else rc = finish_timing(h, true);
    }
    if (stage) cudaFree(stage);
    return rc;
}

extern "C" int riptrm_trs(riptrm_handle* h, const double* x, const double* y, double mu, double Delta, double* dx,
                          double* info, int where, void* stream) {
    return run_hook(h, 3, x, y, mu, Delta, nullptr, dx, info, where, (cudaStream_t)stream);
}

// Placement record of the last solve's fast lane (sphere_tmem2_kernel): out[i] = (smid << 4) | role for CTA i, role 1 main,
// 2 lane, 3 stepped aside; returns the number of records (0: the last solve had no in-kernel lane)
extern "C" int riptrm_lane_placement(riptrm_handle* h, int* out, unsigned long long* times, int capacity) {
    if (h == nullptr || out == nullptr || h->d_lane_state == nullptr) return 0;
    const int n = std::min(capacity, h->lane_debug_len);
    if (cudaMemcpy(out, h->d_lane_state + 2 + 2 * kLaneMaxSms, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (times != nullptr && h->d_lane_times != nullptr &&
        cudaMemcpy(times, h->d_lane_times, (size_t)2 * n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1;
    return n;
}

extern "C" int64_t riptrm_launch_count(const riptrm_handle* h) { return h ? h->launches : 0; }
extern "C" int64_t riptrm_matvec_passes(riptrm_handle* h) {
    if (h == nullptr || h->d_passes == nullptr) return 0;
    unsigned long long v = 0;
    if (cudaMemcpy(&v, h->d_passes, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)v;
}
extern "C" double riptrm_last_kernel_ms(riptrm_handle* h) {
    if (h == nullptr || h->launches == 0) return 0.0;
    if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0;
    h->last_ms = ms;
    return h->last_ms;
}


// Measured FP64 peaks (peaks.cuh): out[0] = DFMA TFLOP/s (vector pipe), out[1] = DMMA TFLOP/s (mma.sync m8n8k4 f64), each the
// best of `repeats` launches of ~`ms_target` ms timed with CUDA events on `stream`.
extern "C" int riptrm_measure_fp64_peaks(int device, double ms_target, int repeats, double* out, void* stream) {
    if (out == nullptr || repeats < 1 || !(ms_target > 0.0)) return fail(RIPTRM_E_INVALID, "measure_fp64_peaks: bad argument");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    cudaStream_t st = (cudaStream_t)stream;
    double* sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int grid = prop.multiProcessorCount * 4, block = 256;
    int rc = RIPTRM_OK;
    for (int which = 0; which < 2 && rc == RIPTRM_OK; ++which) {
        // flops per repetition of the kernel's outer loop, whole grid
        const double flops_rep = (which == 0)
            ? 2.0 * peaks::kChains * peaks::kUnroll * (double)grid * block
            : 2.0 * 8 * 8 * 4 * 4 * peaks::kUnroll * (double)grid * (block / 32);
        int reps = 64;
        double best = 0.0;
        for (int it = 0; it < repeats + 2; ++it) {   // two calibration rounds size `reps` for ms_target
            cudaEventRecord(e0, st);
            if (which == 0) peaks::dfma_kernel<<<grid, block, 0, st>>>(sink, reps, 0.999999, 1e-9);
            else peaks::dmma_kernel<<<grid, block, 0, st>>>(sink, reps, 0.999999, 1e-9);
            cudaEventRecord(e1, st);
            if (cudaEventSynchronize(e1) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
                rc = fail(RIPTRM_E_CUDA, "measure_fp64_peaks: kernel failed");
                break;
            }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it < 2) {
                reps = (int)fmin(1e6, fmax(16.0, reps * ms_target / fmax(ms, 1e-3)));
            } else {
                best = fmax(best, flops_rep * reps / (ms * 1e-3) / 1e12);
            }
        }
        out[which] = best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return rc;
}

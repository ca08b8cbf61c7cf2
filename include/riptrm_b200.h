/*
 * riptrm_b200.h -- C ABI of the B200-native RIPTRM trust-region path.
 *
 * The reference (shirokumakur0/Riemannian-interior-point-trust-region-method) is
 * pure Python and has NO FFI; its drop-in boundary is
 *     solver = RIPTRM(option); output = solver.run(problem)
 * (src/base/base_simulator.py:64-66, src/NonnegPCA/simulator.py:38,
 *  src/solver/RIPTRM.py:303,909).  The entry points below are what a ctypes
 * binding on that boundary needs (INTEGRATION.md shows the stub); each one cites
 * the reference code it replaces.
 *
 * Conventions
 *   - plain C, no torch types; all arrays are double (fp64) unless stated.
 *   - every pointer argument carries a `where` flag: RIPTRM_HOST (pageable or
 *     pinned host memory; the library stages it) or RIPTRM_DEVICE (a device
 *     pointer valid on the handle's GPU, e.g. a torch tensor's data_ptr()).
 *   - all functions return 0 on success or a negative RIPTRM_E_* code;
 *     riptrm_last_error() returns a human-readable message for the calling
 *     thread's last failure.  No exceptions cross the boundary
 *     (reference error behaviour: RIPTRM.py:961-966).
 *   - a handle is bound to one device and is not thread-safe; batching over
 *     (problem_instance, initialpoint) pairs is the concurrency mechanism.
 */
#ifndef RIPTRM_B200_H
#define RIPTRM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIPTRM_ABI_VERSION 2

/* memory space of a pointer argument */
#define RIPTRM_HOST 0
#define RIPTRM_DEVICE 1

/* error codes */
#define RIPTRM_OK 0
#define RIPTRM_E_INVALID (-1)     /* bad argument / unsupported option */
#define RIPTRM_E_CUDA (-2)        /* CUDA runtime failure (message has the cudaError) */
#define RIPTRM_E_UNSUPPORTED (-3) /* problem family / size not covered by a kernel */
#define RIPTRM_E_STATE (-4)       /* call order (e.g. solve before set_problem) */

/* problem families (the workloads the reference solves with RIPTRM) */
typedef enum {
    /* min -x'Zx on Sphere(n), x_i + eps >= 0   -- src/NonnegPCA/coordinator.py:37-95 */
    RIPTRM_FAMILY_NONNEGPCA_SPHERE = 1,
    /* quadratic chain on Grassmann(n,k), vec(X)_i >= -0.01 -- src/Rosenbrock/coordinator.py:33-91 */
    RIPTRM_FAMILY_ROSENBROCK_GRASSMANN = 2,
    /* A=(J-R)Q on Product[Skew(d),SPD(d),SPD(d)] -- src/StableIdentification/coordinator.py:34-179 */
    RIPTRM_FAMILY_STABLEID_PRODUCT = 3,
    /* min -tr(X'ZX), X n x p with unit columns sharing one large Z (Oblique / multi-start
     * reading of BASELINE config 4): the HBM-bound path */
    RIPTRM_FAMILY_NONNEGPCA_COLUMNS = 4,
    /* min -tr(X'ZX) on Stiefel(n,p), X_ij + eps >= 0, one large Z (the Stiefel reading of BASELINE config 4;
     * SURVEY.md App. A.4): ONE run with n x p matrix iterates; batch = 1, m = n*p; x, y, v, out are [n][p]
     * row-major, info is [4], summary [1][..], trace [1][capacity][..].  Same streaming kernel as COLUMNS. */
    RIPTRM_FAMILY_NONNEGPCA_STIEFEL = 5
} riptrm_family;

/* tCG stop reasons -- RIPTRM.py:95,143,145,164,188,190 (dxtype = "tCG_" + name) */
typedef enum {
    RIPTRM_TCG_MAX_INNER_ITER = 0,
    RIPTRM_TCG_NEGATIVE_CURVATURE = 1,
    RIPTRM_TCG_EXCEEDED_TR = 2,
    RIPTRM_TCG_MODEL_INCREASED = 3,
    RIPTRM_TCG_REACHED_TARGET_LINEAR = 4,
    RIPTRM_TCG_REACHED_TARGET_SUPERLINEAR = 5
} riptrm_tcg_stop;

/* dxtype of the exact trust-region solver -- the `type` string TRSgep returns (RIPTRM.py:262, :271, :280, :298), stored in
 * the same trace field as the tCG stop reasons */
typedef enum {
    RIPTRM_TRS_BOUNDARY = 6,
    RIPTRM_TRS_INTERIOR = 7,
    RIPTRM_TRS_HARDCASE_1 = 8,
    RIPTRM_TRS_HARDCASE_3 = 9,
    RIPTRM_TRS_HARDCASE_6 = 10,
    RIPTRM_TRS_HARDCASE_9 = 11
} riptrm_trs_type;

/* 'TRS_solver' (RIPTRM.py:323, :431, :445) */
#define RIPTRM_TRS_SOLVER_TCG 0
#define RIPTRM_TRS_SOLVER_EXACT_REPMAT 1

/* inner_status -- RIPTRM.py:763,770,678,698,829,837 */
typedef enum {
    RIPTRM_INNER_NONE = 0,
    RIPTRM_INNER_CONVERGED = 1,
    RIPTRM_INNER_PRIMAL_INFEASIBLE = 2,
    RIPTRM_INNER_SUCCESSFUL = 3,
    RIPTRM_INNER_UNSUCCESSFUL = 4,
    RIPTRM_INNER_MAX_TIME = 5,
    RIPTRM_INNER_MAX_ITER = 6
} riptrm_inner_status;

/* radius_update -- RIPTRM.py:667-675 */
typedef enum {
    RIPTRM_RADIUS_NONE = 0,
    RIPTRM_RADIUS_REDUCED = 1,
    RIPTRM_RADIUS_EXPANDED = 2,
    RIPTRM_RADIUS_UNCHANGED = 3
} riptrm_radius_update;

/* why the outer loop ended -- base_solver.py:85-106, RIPTRM.py:944-957 */
typedef enum {
    RIPTRM_STOP_RUNNING = 0,
    RIPTRM_STOP_MAXTIME = 1,
    RIPTRM_STOP_MAXITER = 2,
    RIPTRM_STOP_TOLRESID = 3,
    RIPTRM_STOP_NUMERICAL = 4 /* non-finite state; the reference would raise inside outer_step (:961-966) */
} riptrm_stop_reason;

/* Solver options: the numeric keys of RIPTRM.py:305-358.  The callable keys
 * (forcing_function_*, barrier update rule :890-893) are evaluated by the host into the
 * per-outer-iteration schedules below, so any Python callable remains usable. */
typedef struct {
    int32_t maxiter;                    /* 'maxiter' */
    int32_t inner_maxiter;              /* 'inner_maxiter'; < 0 == None */
    int32_t tcg_mininner;               /* 'tCG_mininner' */
    int32_t tcg_maxinner;               /* < 0 == manifold.dim (RIPTRM.py:447) */
    int32_t is_euclidean_embedded;      /* 'is_euclidean_embedded' */
    int32_t trace_mode;                 /* 0 none; 1 one row per inner iteration ('save_inner_iteration'=True);
                                           2 one row per outer iteration (False) */
    int32_t trace_capacity;             /* rows per instance, incl. row 0 */
    int32_t schedule_split;             /* not a reference key.  Batches larger than the GPU's resident capacity are solved
                                           in several launches: every pair is advanced to an outer iteration, the pairs are
                                           sorted by work so far, and the remainder runs longest first.  Results are
                                           bit-identical either way.  0 = automatic (splits at 4/15 and 7/15 of maxiter, and
                                           at 2/3 where two warps share a copy of S), k > 0 = one split at outer
                                           iteration k, < 0 = single launch */
    double tolresid;                    /* 'tolresid' */
    double maxtime;                     /* 'maxtime' seconds, measured on the device clock */
    double inner_maxtime;               /* 'inner_maxtime'; < 0 == None */
    double initial_tr_radius;           /* 'initial_TR_radius'; <= 0 == manifold.typical_dist / 8 (:855-862) */
    double minimal_initial_tr_radius;   /* 'minimal_initial_TR_radius' */
    double maximal_tr_radius;           /* 'maximal_TR_radius' */
    double rho;                         /* 'rho' */
    double reduction_regularization;    /* 'reduction_regularization' */
    double gamma;                       /* 'gamma' */
    double const_left;                  /* 'const_left' */
    double const_right;                 /* 'const_right' */
    double tcg_theta;                   /* 'tCG_theta' */
    double tcg_kappa;                   /* 'tCG_kappa' */
    /* host arrays of length maxiter+1: mu_sched[k] is the barrier parameter used by outer
     * iteration k+1 (mu_sched[0] = 'initial_barrier_parameter'); tol_*[k] the forcing-function
     * values for it (RIPTRM.py:881-885). */
    const double* mu_sched;
    const double* tol_lagrangian_sched;
    const double* tol_complementarity_sched;
    /* ---- ABI 2: the exact trust-region solver on the representation matrix (Sphere n <= 64, Grassmann, Product) ---- */
    int32_t trs_solver;                 /* 'TRS_solver': RIPTRM_TRS_SOLVER_TCG | RIPTRM_TRS_SOLVER_EXACT_REPMAT (:431-444).  The
                                           tangent basis ('basisfun', random in the reference) is fixed per manifold
                                           (riptrm_b200/basis.py); step and eigenvalues do not depend on the basis */
    int32_t second_order_stationarity;  /* 'second_order_stationarity' (:599-617); needs EXACT_REPMAT as in the reference */
    double trs_tolhardcase;             /* 'TRS_tolhardcase' (:432, :263) */
    const double* tol_second_order_sched; /* forcing_function_second_order(mu) per outer iteration (:886-887); may be NULL
                                           when second_order_stationarity == 0 */
} riptrm_options;

/* One trace row = RIPTRM_TRACE_FIELDS doubles; field indices below.  Mirrors the reference
 * log row (base_solver.py:58-76; utils.py:356-364; RIPTRM.py:980-1024) + tcg_iters. */
#define RIPTRM_TRACE_FIELDS 26
enum {
    RIPTRM_TR_ITERATION = 0, RIPTRM_TR_NUM_INNER = 1, RIPTRM_TR_MU = 2, RIPTRM_TR_RADIUS = 3,
    RIPTRM_TR_DXTYPE = 4, RIPTRM_TR_TCG_ITERS = 5, RIPTRM_TR_NORMDX = 6, RIPTRM_TR_MINXFEASI = 7,
    RIPTRM_TR_MINYFEASI = 8, RIPTRM_TR_COMPL = 9, RIPTRM_TR_ARED_PRED = 10, RIPTRM_TR_RADIUS_UPDATE = 11,
    RIPTRM_TR_INNER_STATUS = 12, RIPTRM_TR_DUAL_CLIPPING = 13, RIPTRM_TR_MAXABSLAGMULT = 14,
    RIPTRM_TR_COST = 15, RIPTRM_TR_DISTANCE = 16, RIPTRM_TR_RESIDUAL = 17, RIPTRM_TR_GRADNORM = 18,
    RIPTRM_TR_COMPLVIOLATION = 19, RIPTRM_TR_DUALVIOLATION = 20, RIPTRM_TR_MANVIOLATION = 21,
    RIPTRM_TR_MAXVIOLATION = 22, RIPTRM_TR_MEANVIOLATION = 23,
    RIPTRM_TR_TIME = 24, /* seconds since the solve started, device clock (base_solver.py:71) */
    RIPTRM_TR_MINEIGVALHW = 25 /* 'mineigvalHw' (RIPTRM.py:609-610, :1003); NaN on the tCG path */
};
/* absent values (Python None) are stored as NaN; dual_clipping: 0 False, 1 True, NaN None */

/* Per-instance summary = RIPTRM_SUMMARY_FIELDS doubles */
#define RIPTRM_SUMMARY_FIELDS 16
enum {
    RIPTRM_SM_COST = 0, RIPTRM_SM_RESIDUAL = 1, RIPTRM_SM_GRADNORM = 2, RIPTRM_SM_COMPLVIOLATION = 3,
    RIPTRM_SM_DUALVIOLATION = 4, RIPTRM_SM_MANVIOLATION = 5, RIPTRM_SM_MAXVIOLATION = 6,
    RIPTRM_SM_MEANVIOLATION = 7, RIPTRM_SM_MU = 8, RIPTRM_SM_RADIUS = 9, RIPTRM_SM_OUTER_ITERS = 10,
    RIPTRM_SM_INNER_ITERS = 11, RIPTRM_SM_TCG_ITERS = 12, /* sum of (j+1) over tCG calls */
    RIPTRM_SM_AUX_HESSVECS = 13,                         /* fresh Hw(dx) products formed for RIPTRM.py:659 (0 where the
                                                            tCG's accumulated Hw[eta] is used: Sphere, TRS_solver='tCG') */
    RIPTRM_SM_STOP_REASON = 14, RIPTRM_SM_TRACE_ROWS = 15 /* rows produced (may exceed capacity) */
};

typedef struct riptrm_handle riptrm_handle;

/* ---- lifetime ------------------------------------------------------------------------- */
int riptrm_abi_version(void);
const char* riptrm_last_error(void);

/* Creates a solver for `batch` independent (instance, initialpoint) pairs of one family on
 * CUDA device `device`.  Replaces RIPTRM.__init__ + outer_preprocess (RIPTRM.py:303-365,849-864).
 *   n, p : point shape (Sphere: n, p=1; Grassmann: n x p; StableId: n=d, p=3 blocks; columns: n x p)
 *   m    : number of inequality constraints per instance */
int riptrm_create(int family, int n, int p, int m, int batch, int device, riptrm_handle** out);
int riptrm_destroy(riptrm_handle* h);

/* ---- problem data (replaces the coordinator's closures) -------------------------------- */
/* NonnegPCA (families 1 and 4): Z is [batch_z][n][n] row-major, NOT symmetric
 * (src/NonnegPCA/coordinator.py:50-54); batch_z divides `batch`: the batch is batch_z problem instances times
 * batch / batch_z initial points each, pair i using Z[i / (batch / batch_z)] (batch_z == batch: one Z per pair;
 * batch_z == 1: one Z shared by all pairs).
 * eps is the constraint offset: g_i = -x_i - eps (0 in the reference). */
int riptrm_set_nonnegpca(riptrm_handle* h, const double* Z, int batch_z, double eps, int where);

/* Rosenbrock (family 2): alpha of src/Rosenbrock/config_simulation.yaml:12; offset of :62 */
int riptrm_set_rosenbrock(riptrm_handle* h, double alpha, double offset);

/* StableIdentification (family 3): X, XP are [d][N] row-major (coordinator.py:73-88), h the
 * step (:53).  conspec is [m][5] = {kind, row, col, a, b}, one row per inequality constraint in
 * the order coordinator.py:132-152 appends them: kind 0: g = -A[row,col] + a ; kind 1:
 * g = A[row,col] - a ; kind 2: g = -(A[row,col] - a)^2 + b.  Shared by the whole batch. */
int riptrm_set_stableid(riptrm_handle* h, const double* X, const double* XP, int N, double hstep,
                        const double* conspec, int m, int where);

int riptrm_set_options(riptrm_handle* h, const riptrm_options* opts);

/* ---- the solve (replaces RIPTRM.run, RIPTRM.py:909-976) --------------------------------- */
/* x0 [batch][n*p] (Product: J,R,Q concatenated, each row-major), y0 [batch][m].
 * Outputs (any may be NULL): x [batch][n*p], y [batch][m], summary [batch][RIPTRM_SUMMARY_FIELDS],
 * trace [batch][trace_capacity][RIPTRM_TRACE_FIELDS].  `stream` is a cudaStream_t (or NULL).
 * With `where` == RIPTRM_HOST the call returns after the results are in the host buffers;
 * with RIPTRM_DEVICE every family only enqueues work on `stream`.  The COLUMNS / STIEFEL families (4, 5) run their whole
 * solve as ONE graph launch: a conditional WHILE node whose body is the tCG launch and the post launch of a trust-region
 * iteration (cooperative kernels), ended from the device when every run has finished; 'maxtime' / 'inner_maxtime' are tested
 * against a device clock.  Where conditional graph nodes are missing (or with RIPTRM_COLUMNS_HOST_LOOP=1) the host sequences
 * the launches, four iterations ahead of the one-int "all done" flag it polls, and the call returns when the solve has
 * finished. */
int riptrm_solve(riptrm_handle* h, const double* x0, const double* y0, double* x, double* y,
                 double* summary, double* trace, int where, void* stream);

/* ---- unit hooks on the hot path (parity + roofline) ------------------------------------- */
/* out = Hw[v] at (x, y, mu): Hess_x L(x,y)[v] + G_x(y * G*_x[v] / s)   (RIPTRM.py:729) */
int riptrm_hessvec(riptrm_handle* h, const double* x, const double* y, double mu, const double* v,
                   double* out, int where, void* stream);
/* One tCG solve at (x, y, mu, Delta) (RIPTRM.py:41-216 via :445-452):
 * eta [batch][n*p], info [batch][4] = {j+1, stop reason, ||eta||, <eta,Hw eta> model value}
 * (COLUMNS: info [p][4], one tCG per column; STIEFEL: info [1][4]) */
int riptrm_tcg(riptrm_handle* h, const double* x, const double* y, double mu, double Delta, double* eta,
               double* info, int where, void* stream);

/* One exact trust-region solve at (x, y, mu, Delta) (RIPTRM.py:431-444: representation matrix of Hw in the fixed tangent
 * basis, `TRSgep` :218-299): dx [batch][n*p], info [batch][4] = {type (riptrm_trs_type), lam1, ||dx||_x, smallest eigenvalue of
 * the representation matrix of Hw at (x, y)}.  Sphere (n <= 64), Grassmann and Product families. */
int riptrm_trs(riptrm_handle* h, const double* x, const double* y, double mu, double Delta, double* dx,
               double* info, int where, void* stream);

/* The condensed Newton system of the reference's Riemannian interior-point method on the same operator (SURVEY.md section 8f
 * rank 4; src/solver/RIPM.py:484-511):   Aw[dx] = Hess_x L(x, z)[dx] + G_x( G*_x[dx] * z / s ) = c,   with the slacks s an
 * independent variable [batch][m] and c [batch][n*p] the right-hand side (made tangent on entry).
 *   method 0: RepresentMatMethod (RIPM.py:238-300; no equality constraints) -- representation matrix in the tangent basis, dense
 *             symmetric solve;   method 1: TangentSpaceConjResMethod (utils.py:582-618) -- conjugate residuals from v0 = 0 until
 *             |r| / |c| < tol ('KrylovTolrelresid') or `maxiter` ('KrylovMaxIteration') iterations.
 * dx [batch][n*p]; info [batch][4] = {iterations (0 for method 0), |c - Aw dx| / |c|, ||dx||_x, smallest eigenvalue of the
 * matrix of Aw (method 0; NaN for method 1)}.  Sphere (n <= 64), Grassmann and Product families. */
int riptrm_newton(riptrm_handle* h, const double* x, const double* z, const double* s, const double* c, int method,
                  double tol, int maxiter, double* dx, double* info, int where, void* stream);

/* The dense core of the above on caller-supplied data -- `TRSgep(A, a, I, Delta, tolhardcase)` (RIPTRM.py:218-299) for `count`
 * independent problems, one warp each: A [count][d][d] symmetric, a [count][d] -> x [count][d], info [count][4] = {type, lam1,
 * ||x||, smallest eigenvalue of A}.  d <= 64. */
int riptrm_trs_dense(int device, int d, int count, const double* A, const double* a, double Delta, double tolhardcase,
                     double* x, double* info, int where, void* stream);

/* ---- synthetic sweeps drawn on the device (src/NonnegPCA/generator.py:9-65) ----------------
 * `instances` NonnegPCA instances first_instance .. by the reference generator's law (snr, delta as in
 * config_dataset.yaml:7-8), each with `points_per_instance` feasible starting points: Z [instances][n][n],
 * x0 / y0 [instances * points_per_instance][n] (instance-major), all DEVICE pointers on `device`.  Every number is a
 * function of (instance id, stream, index) through Philox4x32-10, so ranks draw their own shares independently. */
int riptrm_generate_nonnegpca(int device, int n, long long first_instance, int instances, int points_per_instance,
                              double snr, double delta, double* Z, double* x0, double* y0, void* stream);

/* Diagnostic: where and when the CTAs of the last launch of a Sphere solve with a fast lane ran.  out[i] = (SM id << 4) | role
 * (role 1: main CTA, 2: lane CTA, 3: stepped aside for a lane CTA; 0: no record), times[2 i], times[2 i + 1] = %globaltimer at
 * the CTA's entry and exit (ns; may be NULL).  Records [0, 2 x SMs) are the main kernel's CTAs, the lane kernel's CTAs of the
 * two-kernel form follow.  Returns the number of records written (0: no lane in the last solve, < 0: CUDA error). */
int riptrm_lane_placement(riptrm_handle* h, int* out, unsigned long long* times, int capacity);

/* number of kernel launches the handle has issued (for bench.py's gpu_launches) */
int64_t riptrm_launch_count(const riptrm_handle* h);
/* COLUMNS and STIEFEL families: number of full S.V streaming passes (one per Hessian-vector product, plus the S.X of the
 * point cache) the handle has executed since riptrm_set_nonnegpca -- the unit of the HBM roofline */
int64_t riptrm_matvec_passes(riptrm_handle* h);
/* milliseconds the last riptrm_solve / hessvec / tcg kernel took on its stream (CUDA events
 * recorded around the launch); blocks until that kernel has finished */
double riptrm_last_kernel_ms(riptrm_handle* h);

/* ---- measurement utility (no reference counterpart) -------------------------------------------------------------------
 * FP64 arithmetic peaks of `device`, measured by saturating micro-kernels (csrc/peaks.cuh): out[0] = DFMA TFLOP/s (the FP64
 * vector pipe the whole-solve kernels run on), out[1] = DMMA TFLOP/s (mma.sync.m8n8k4.f64, the tensor path of the COLUMNS /
 * STIEFEL consumers).  Best of `repeats` launches of about `ms_target` ms each, CUDA events on `stream`.  bench.py uses them
 * as the denominators of `step_roofline` (MEASURED_PEAKS.json has no fp64 figure; the data sheet's 40 TFLOP/s is nominal). */
int riptrm_measure_fp64_peaks(int device, double ms_target, int repeats, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RIPTRM_B200_H */

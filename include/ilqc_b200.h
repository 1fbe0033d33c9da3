/* ilqc_b200 -- C ABI of the B200-native batched iLQR/DDP engine.
 *
 * The reference (vroulet/ilqc) is pure Python: it has no FFI, its boundary is the Python API
 *   run_min_algo(env, max_iter, algo, ...)            algorithms/run_min_algo.py:15-100
 *   DiffEnv.forward(cmd, approx)                      envs/forward.py:100-139
 *   lin_quad_backward / quad_backward                 envs/backward.py:6-150
 *   roll_out_lin / roll_out_exact                     envs/rollout.py:8-88
 *   mpc_step                                          model_predictive_control/mpc_pipeline.py:18-55
 * Each entry point below cites the reference function it replaces for a *batch* of independent
 * problems.  The Python package ilqc_b200 binds these symbols with ctypes (ilqc_b200/_lib.py) and keeps
 * the reference's names/signatures on top (ilqc_b200/algorithms, ilqc_b200/envs, ...).
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host;
 *  - float64 everywhere; structure-of-arrays with the batch index innermost ("[t][c][B]" means element
 *    (t,c,b) lives at ((t*C)+c)*B+b) so that thread-per-problem kernels read coalesced;
 *  - stream is a cudaStream_t passed as void*; kernels are enqueued on it, nothing synchronises unless
 *    stated; no hidden device allocations: scratch memory is supplied by the caller (workspace);
 *  - return value: 0 = ok, <0 = invalid argument (ILQC_E_*), >0 = cudaError_t of a failed launch.
 *    Nothing throws across the ABI.
 */
#ifndef ILQC_B200_H
#define ILQC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ILQC_ABI_VERSION 3

enum { ILQC_E_ARG = -1, ILQC_E_UNSUPPORTED = -2, ILQC_E_WORKSPACE = -3 };

/* models (envs/pendulum.py:13, :115 ; envs/car.py:134, :148 ; envs/lin_quad.py:7) */
enum { ILQC_MODEL_PENDULUM = 0, ILQC_MODEL_CARTPOLE = 1, ILQC_MODEL_CAR_SIMPLE = 2, ILQC_MODEL_CAR_BICYCLE = 3 };
/* integrators (envs/discretization.py:8, :42, :22) */
enum { ILQC_INT_EULER = 0, ILQC_INT_RK4_CST = 1, ILQC_INT_RK4 = 2 };
/* Car cost (envs/car.py:190-253) */
enum { ILQC_COST_CONTOURING = 0, ILQC_COST_EXACT = 1 };
/* Car time penalty (envs/car.py:209-224) */
enum { ILQC_TP_VREF_SQUARED = 0, ILQC_TP_TREF = 1, ILQC_TP_DREF = 2, ILQC_TP_DREF_SQUARED = 3 };
/* spline termination (envs/tracks/torch_spline.py:341-347) */
enum { ILQC_TERM_LOOP = 0, ILQC_TERM_REPEAT_END = 1 };

/* Problem family description: env scalars + device pointers to the shared spline tables and to the
 * optional per-problem parameter arrays.  Mirrors the constructor arguments of Pendulum / CartPendulum
 * / Car (envs/pendulum.py:22-54, :123-176 ; envs/car.py:27-114). */
typedef struct ilqc_model_desc {
  int32_t model, integrator, cost, time_penalty;
  int32_t horizon;          /* number of control steps H */
  int32_t t0;               /* env.init_time_iter (forward.py:149) unless pp_t0 is given */
  int32_t stay_put_start;   /* horizon - int(stay_put_time/dt), computed on the host (pendulum.py:44-48) */
  int32_t has_x_limits;     /* CartPendulum x_limits is not None */
  int32_t n_seg;            /* spline segments = n_knots-1 */
  int32_t termination;
  int32_t reg_ctrl_is_tuple; /* Car reg_ctrl given as (reg, reg_tref) (car.py:235-249) */
  int32_t reserved0;
  double dt;
  /* Pendulum / CartPendulum */
  double reg_speed, reg_ctrl, reg_ctrl_tref, reg_barrier, x_min, x_max;
  double phys[8];           /* pendulum: g,m,l,mu ; cart-pendulum: g,M,m,b,I,l */
  /* Car */
  double reg_cont, reg_lag, reg_bar, reg_obs, vref, constrain_angle, acc_lo, acc_hi;
  double Cm1, Cm2, Cr0, Cr2, Br, Cr, Dr, Bf, Cf, Df, m, Iz, lf, lr, car_l, car_w;
  double obs_x, obs_y, track_len;
  const double* knots;      /* [n_seg+1] shared knot vector (tracks.py:50-53) */
  const double* coef;       /* [n_seg][3 splines: centre,inner,outer][2 channels][4: a,b,c,d] */
  /* optional per-problem overrides, NULL = use the scalars above */
  const double* pp_weights; /* [3][B]: reg_cont, reg_lag, reg_speed (car.py:52-56) */
  const int32_t* pp_t0;     /* [B] */
} ilqc_model_desc;

/* algorithm grammar of run_min_algo.py:267-292 */
enum { ILQC_ALGO_GD = 0, ILQC_ALGO_CLASSIC = 1, ILQC_ALGO_DDP = 2 };
enum { ILQC_APPROX_LINQUAD = 0, ILQC_APPROX_QUAD = 1 };
enum { ILQC_STEP_DIR = 0, ILQC_STEP_REG = 1, ILQC_STEP_DIRVAR = 2, ILQC_STEP_REGVAR = 3 };
/* per-problem status words (check_cvg_status, run_min_algo.py:195-264) */
enum { ILQC_RUNNING = 0, ILQC_CONVERGED = 1, ILQC_DIVERGED = 2, ILQC_STUCK = 3, ILQC_NO_DIRECTION = 4 };
/* backward-pass flavours (backward.py:6, :135, :144) */
enum { ILQC_BWD_LINQUAD = 0, ILQC_BWD_QUAD_NEWTON = 1, ILQC_BWD_QUAD_DDP = 2 };
/* handle_bad_dir of bell_step (backward.py:219-240): what a positive rest (non-convex step) does.
 * check_total_value (run_min_algo's default): accept the step, quad flavours are feasible iff jcst < 0 ;
 * flag: the pass is infeasible as soon as one step has rest > 0 ; let_it_be: always feasible.
 * ('modify_hess' calls the removed torch.eig in the reference and is dead there.)
 * For ilqc_backward the value is or-ed into `mode` as (handle_bad_dir << 4). */
enum { ILQC_BAD_DIR_CHECK_TOTAL = 0, ILQC_BAD_DIR_FLAG = 1, ILQC_BAD_DIR_LET_IT_BE = 2 };
/* roll-out flavours (rollout.py:8, :50 ; algos_steps.py:22-43) */
enum { ILQC_ROLL_EXACT = 0, ILQC_ROLL_LIN = 1, ILQC_ROLL_GD = 2 };

typedef struct ilqc_algo_desc {
  int32_t algo_type, approx, step_mode;
  int32_t n_candidates;     /* L: line-search candidates s*rho^c, c<L, evaluated per launch (>=1) */
  int32_t line_search;      /* run_min_algo(line_search=...) ; 0 = fixed stepsize */
  int32_t max_ls_trips;     /* safety cap on line-search trips (reference: unbounded, stops at 1e-32) */
  int32_t compute_grad_norm;/* collect_info's ||grad_u f|| each iteration (run_min_algo.py:156-157) */
  int32_t scheduler;        /* 0 auto, 1 lock-step iterations, 2 asynchronous per-problem state machines
                               (regularised step modes and gd only); identical results either way */
  int32_t handle_bad_dir;   /* ILQC_BAD_DIR_* (run_min_algo(handle_bad_dir=...), run_min_algo.py:21) */
  int32_t expansion;        /* 0: one thread per problem walks the horizon (roll-out + expansion fused; best for
                               large batches).  1: time-parallel, for small batches (MPC windows, single problems):
                               a roll-out pass stores the trajectory, then one thread per (problem, time step)
                               expands its step.  Same mathematics; the two modes differ by rounding (1e-14). */
  int32_t reserved1, reserved2;
  const double* pp_cost0;   /* [B] device, nullable.  Restart (run_min_algo(past_metrics=...), run_min_algo.py:72-75):
                               the ORIGINAL metrics['cost'][0] of every problem, against which check_cost_status tests
                               divergence (cost > cost[0] + 1e3, run_min_algo.py:240-243).  NULL: row 0 of this call. */
} ilqc_algo_desc;

int ilqc_abi_version(void);

/* dimensions implied by the model/integrator (Car: nx 6/8, nu 3 or 9 for rk4 ; car.py:88-101) */
int ilqc_dims(const ilqc_model_desc* m, int32_t* nx, int32_t* nu);
/* doubles per time step per problem of the packed (sparsity-aware) expansion written by ilqc_linearize */
int ilqc_lin_stride(const ilqc_model_desc* m);
int ilqc_gain_stride(const ilqc_model_desc* m);

/* DiffEnv.forward(cmd) without approx (forward.py:100-139, :94-96): roll-out + costs.
 * x0 [nx][B], cmd [H][nu][B] ; out traj [H+1][nx][B] (nullable), cost_t [H+1][B] (nullable; cost_t[0]=0),
 * total [B] = Python sum(costs), left to right (run_min_algo.py:154). */
int ilqc_rollout_cost(const ilqc_model_desc* m, int B, const double* x0, const double* cmd, double* traj,
                      double* cost_t, double* total, void* stream);

/* DiffEnv.forward(cmd, approx='linquad') (forward.py:84-92, :188-216, :274-321): analytic first-order
 * expansion of the dynamics and second-order expansion of the costs in the packed layout
 * lin [H][ilqc_lin_stride][B]; also traj [H+1][nx][B] and total [B]. */
int ilqc_linearize(const ilqc_model_desc* m, int B, const double* x0, const double* cmd, double* lin,
                   double* traj, double* total, void* stream);

/* Same expansion, dense, as the reference stores it on traj[t+1] / costs[t] (forward.py:209-216, :299-319),
 * plus (order==2) the dynamics-Hessian closures hess_dyn_state/ctrl/state_ctrl(lam) (forward.py:218-269)
 * evaluated at lam [H][nx][B].  For tests and external callers.
 * A [H][nx][nx][B] (=grad_dyn_state.T), Bm [H][nx][nu][B], p [H+1][nx][B], P [H+1][nx][nx][B],
 * q [H][nu][B], Q [H][nu][nu][B], Hxx [H][nx][nx][B], Hxu [H][nx][nu][B], Huu [H][nu][nu][B]. */
int ilqc_expand_dense(const ilqc_model_desc* m, int B, const double* x0, const double* cmd, int order,
                      double* A, double* Bm, double* p, double* P, double* q, double* Q, const double* lam,
                      double* Hxx, double* Hxu, double* Huu, void* stream);

/* lin_quad_backward (backward.py:6-57) / bell_step+bell_core (:153-257) on DENSE expansions of any small
 * size (instantiated (nx, nu): (2,1) (4,1) (4,3) (4,4) (6,3) (8,3); other sizes return ILQC_E_UNSUPPORTED): one thread per (problem, candidate).  n_slots = B*L, slot = b*L+c uses
 * problem b and reg_ctrl[slot].  R is nullable ([H][nx][nu][B], quad_backward's cross term).
 * out K [H][nu][nx][n_slots], k [H][nu][n_slots], jcst [n_slots]. */
int ilqc_lq_backward_dense(int nx, int nu, int H, int B, int L, const double* A, const double* Bm,
                           const double* p, const double* P, const double* q, const double* Q,
                           const double* R, const double* reg_ctrl, double reg_state, double* K, double* k,
                           double* jcst, void* stream);
/* roll_out_lin (rollout.py:8-47) on dense expansions; step [n_slots]; out diff [H][nu][n_slots]. */
int ilqc_rollout_lin_dense(int nx, int nu, int H, int B, int L, const double* A, const double* Bm,
                           const double* K, const double* k, const double* step, double* diff, void* stream);

/* Backward pass on the packed expansion of a model (backward.py:6-150).  mode = ILQC_BWD_*.
 * n_slots threads; slot s serves problem prob[s] (NULL: s / L... identity) with reg_ctrl[s].
 * The quad modes recompute the dynamics Hessians contracted with the running lambda from (traj, cmd).
 * out gains [H][ilqc_gain_stride][n_slots], jcst [n_slots], feasible [n_slots] (int32). */
int ilqc_backward(const ilqc_model_desc* m, int mode, int B, int n_slots, const int32_t* prob,
                  const double* reg_ctrl, double reg_state, const double* lin, const double* traj,
                  const double* cmd, double* gains, double* jcst, int32_t* feasible, void* stream);

/* roll_out_exact / roll_out_lin / gd direction (rollout.py:50-88, :8-47 ; algos_steps.py:22-43) fused with
 * the objective of the candidate, obj(cmd+diff) (algos_steps.py:329-331, :416).
 * gslot [n_slots] (nullable = identity): which gains slot each candidate reads.
 * out diff [H][nu][n_slots], newval [n_slots], diffsq [n_slots] (= ||diff||^2). */
int ilqc_rollout(const ilqc_model_desc* m, int kind, int B, int n_slots, const int32_t* prob,
                 const int32_t* gslot, const double* step, const double* x0, const double* cmd,
                 const double* traj, const double* lin, const double* gains, const double* grad,
                 double* diff, double* newval, double* diffsq, void* stream);

/* grad_u sum(costs) through the adjoint recursion on the packed expansion (replaces
 * torch.autograd.grad(sum(costs), cmd), run_min_algo.py:157, algos_steps.py:37) and the norm used to
 * scale the regularised line search (algos_steps.py:342-351).
 * out grad [H][nu][B] (nullable), gnorm [B] = ||grad||, cnorm [B] (nullable) = sqrt(sum_t |grad_ctrl_t|^2+|grad_state_t|^2). */
int ilqc_grad_obj(const ilqc_model_desc* m, int B, const double* lin, double* grad, double* gnorm,
                  double* cnorm, void* stream);

/* run_min_algo for B problems at once (run_min_algo.py:15-100 with min_step, algos_steps.py:277-437).
 * cmd [H][nu][B] in/out, stepsize [B] in/out (the optimiser's auxiliary variable),
 * metrics out: cost/gnorm/step_rep [n_iter+1][B] (NaN-padded after a problem stops), n_done [B] iterations
 * executed, status [B]; ls_trips [n_iter][B] (nullable) line-search trips per iteration.
 * workspace: device scratch of at least ilqc_solve_workspace_bytes(). Synchronises the stream (it reads
 * back the number of problems still searching after every line-search round). */
size_t ilqc_solve_workspace_bytes(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B);
int ilqc_solve(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, const double* x0, double* cmd,
               double* stepsize, int n_iter, double* cost, double* gnorm, double* step_rep, int32_t* n_done,
               int32_t* status, int32_t* ls_trips, void* workspace, size_t workspace_bytes, void* stream);

/* Same call for HOST buffers in the reference's own layout (x0_host [B][nx], cmd [B][H][nu],
 * stepsize_host [B], metrics [B][n_iter+1]): copies in, transposes on the device, solves, copies back.
 * cmd0_host: initial controls, or NULL to start from zeros like run_min_algo(prev_cmd=None)
 * (run_min_algo.py:53-57; nothing is copied in then); cmd_host: output (may alias cmd0_host).
 * The desc pointers (spline tables, per-problem arrays) are still device pointers. */
int ilqc_solve_host(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, const double* x0_host,
                    const double* cmd0_host, double* cmd_host, double* stepsize_host, int n_iter, double* cost_host,
                    double* gnorm_host, double* step_rep_host, int32_t* n_done_host, int32_t* status_host,
                    void* workspace, size_t workspace_bytes, void* stream);
size_t ilqc_solve_host_workspace_bytes(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, int n_iter);

/* run_mpc (model_predictive_control/mpc_pipeline.py:87-115) for B concurrent closed-loop episodes, the plant being
 * the model: windows of growing full_horizon h <- min(max(h + window_size - overlap, window_size), full_horizon),
 * every window = mpc_step (mpc_pipeline.py:18-55): warm start from the last `overlap` controls of the previous
 * window + (full_horizon - previous length) copies of its last control (or zeros, keep_applying_last_ctrl = 0),
 * initial state re-anchored on the trajectory of the previous controls (Python's negative-index behaviour
 * included when the previous window is shorter than the overlap), init_time_iter = previous length - overlap,
 * result spliced behind the frozen prefix.  optim_on_full_window = 1 re-optimises the whole horizon from x0.
 * The first window is solved on m->horizon steps (run_min_algo sizes cmd from env.horizon, run_min_algo.py:53).
 * The window loop, the warm-start shift / padding kernel and the re-anchoring roll-out run inside this call:
 * no per-window host work besides the launches.
 * x0 [nx][B]; out cmd [full_horizon][nu][B] (rows >= the last window's horizon untouched), cmd_len (host int,
 * number of valid rows); per-window metrics (nullable): cost [W][max_iter+1][B], status [W][B], n_done [W][B]
 * with W = ilqc_mpc_num_windows(...), or max_windows if 0 < max_windows < W.  m->pp_t0 must be NULL. */
int ilqc_mpc_num_windows(int full_horizon, int window_size, int overlap);
size_t ilqc_mpc_workspace_bytes(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, int full_horizon,
                                int window_size, int overlap, int max_iter, int optim_on_full_window);
int ilqc_mpc(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, const double* x0, int full_horizon,
             int window_size, int overlap, int max_iter, int optim_on_full_window, int keep_applying_last_ctrl,
             int max_windows, double* cmd, int32_t* cmd_len_host, double* cost, int32_t* status, int32_t* n_done,
             void* workspace, size_t workspace_bytes, void* stream);

/* newton_oracle (algorithms/algos_steps.py:46-102; the dense cross-check of tests/test_newton.py): Newton step on the
 * flattened problem.  The (H nu x H nu) Hessian of the total cost is assembled from the analytic second-order expansion
 * (costate + one second-order adjoint sweep per column) instead of autodiff, then (Hessian + reg I) d = -grad is solved
 * by LU with partial pivoting in one thread block per problem (torch.linalg.solve).  A diagnostic: nx <= 8, nu <= 9,
 * memory grows like B (H nu)^2.
 * x0 [nx][B], cmd [H][nu][B], reg [B] ; out dir [H][nu][B], opt [B] = dir . grad / 2 ; optional grad [H][nu][B],
 * hess [B][H nu][H nu] (row-major per problem). */
size_t ilqc_newton_dense_workspace_bytes(const ilqc_model_desc* m, int B);
int ilqc_newton_dense(const ilqc_model_desc* m, int B, const double* x0, const double* cmd, const double* reg,
                      double* dir, double* opt, double* grad, double* hess, void* workspace, size_t workspace_bytes,
                      void* stream);
/* Jacobian of the trajectory traj[1:] with respect to the controls, the matrix compute_min_sing_eigval_jac
 * (algorithms/run_min_algo.py:103-124) builds with auto_multi_grad: jac [B][H nx][H nu], row (t-1) nx + i = state i of
 * x_t, column s nu + k = control k of step s (zero for s >= t). */
size_t ilqc_traj_jacobian_workspace_bytes(const ilqc_model_desc* m, int B);
int ilqc_traj_jacobian(const ilqc_model_desc* m, int B, const double* x0, const double* cmd, double* jac,
                       void* workspace, size_t workspace_bytes, void* stream);

/* [B][R][C] (reference layout, batch outermost) <-> [R][C][B] (engine layout) on the device. */
int ilqc_to_soa(const double* src, double* dst, int B, int n, void* stream);
int ilqc_from_soa(const double* src, double* dst, int B, int n, void* stream);

/* Optional CUDA-event timing of the kernel classes launched by ilqc_solve (0 roll-out+cost, 1 expansion,
 * 2 backward, 3 candidate roll-out, 4 adjoint).  ilqc_profile_read fills ms/launches/slots (arrays of >=5),
 * resets the counters and returns the number of classes. */
int ilqc_profile_enable(int enable);
int ilqc_profile_read(double* ms, int64_t* launches, int64_t* units, int n);

/* Engine tuning knobs (no effect on results): "coop_max" = launches of the linear-quadratic backward pass with at most
 * this many candidates use the warp-cooperative kernel (8 lanes per candidate, csrc/coop.cuh; identical gains),
 * "feed_small" = ... the cp.async operand feed.  A negative value restores the default.  Returns ILQC_E_ARG for an unknown name.  (Engine-side only: the
 * reference has no counterpart; its backward pass is envs/backward.py:6-57.) */
int ilqc_set_tuning(const char* name, long long value);

/* Number of kernels launched by this library since the last reset (all entry points, bookkeeping included). */
long long ilqc_launch_count(int reset);

/* Self-test of the device math layer (csrc/fastmath.cuh: branch-free FP64 sin/cos, atan2, atan, division,
 * reciprocal, sqrt + 1/sqrt used by the car models in place of the CUDA math library on their stated
 * argument ranges).  fn: 0 sincos(a) -> o0 = sin, o1 = cos ; 1 atan2(a, b) ; 2 atan(a) ; 3 a / b ; 4 1 / a ;
 * 5 o0 = sqrt(a), o1 = 1/sqrt(a).  n elements, device pointers (b, o1 may be NULL where unused). */
int ilqc_fastmath_eval(int fn, int n, const double* a, const double* b, double* o0, double* o1, void* stream);

/* FP64 DFMA micro-benchmark used for the roofline denominator: returns achieved FLOP/s (2 per DFMA). */
int ilqc_fp64_peak(int iters, double* flops_per_s, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ILQC_B200_H */

// extern "C" entry points declared in include/ilqc_b200.h.  Shared by the product library (api.cu, nvcc)
// and the test-only host simulator (tests/hostsim/api_hostsim.cpp).
#include "launch.cuh"
#include "fastmath.cuh"
#include "solver.inl"
#include "solver_async.inl"
#include "mpc.inl"
#include "newton_dense.inl"

namespace ilqc {
int dense_lq_backward(int nx, int nu, int H, int B, int L, const double* A, const double* Bm, const double* p,
                      const double* P, const double* q, const double* Q, const double* R, const double* reg_ctrl,
                      double reg_state, double* K, double* k, double* jcst, stream_t st);
int dense_rollout_lin(int nx, int nu, int H, int B, int L, const double* A, const double* Bm, const double* K,
                      const double* k, const double* step, double* diff, stream_t st);
}


namespace ilqc {
// self-test of the device math layer (fastmath.cuh): evaluates one of its functions elementwise
ILQC_GLOBAL fastmath_eval_kernel(int fn, int n, const double* a, const double* b, double* o0, double* o1) {
  const int i = ILQC_TID;
  if (i >= n) return;
  const double x[1] = {a[i]}, y[1] = {b ? b[i] : 0.0};
  double r0[1] = {0.0}, r1[1] = {0.0};
  switch (fn) {
    case 0: fm::sincos<1>(x, r0, r1); break;
    case 1: fm::atan2<1>(x, y, r0); break;   // atan2(a, b)
    case 2: fm::atan<1>(x, r0); break;
    case 3: r0[0] = fm::div(x[0], y[0]); break;
    case 4: r0[0] = fm::rcp(x[0]); break;
    case 5: fm::sqrt_rsqrt(x[0], &r0[0], &r1[0]); break;
    default: break;
  }
  o0[i] = r0[0];
  if (o1) o1[i] = r1[0];
}
}  // namespace ilqc

using namespace ilqc;

// scheduler choice (ilqc_algo_desc.scheduler): 0 = automatic, 1 = lock-step iterations, 2 = asynchronous
// per-problem state machines on two streams (regularised step modes and gd only).  Both give bit-identical
// results; on a B200 they are within noise of each other once the speculation capacity is tuned
// (profiles/r01_tuning_scheduler.txt), so automatic = the simpler lock-step driver.
static int solve_dispatch(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, const double* x0, double* cmd,
                          double* stepsize, int n_iter, double* cost, double* gnorm, double* step_rep, int32_t* n_done,
                          int32_t* status, int32_t* ls_trips, void* workspace, size_t workspace_bytes, stream_t st) {
  if (!a) return ILQC_E_ARG;
  const bool regmode = a->algo_type == ILQC_ALGO_GD || a->step_mode == ILQC_STEP_REG || a->step_mode == ILQC_STEP_REGVAR;
  int sched = a->scheduler;
  if (sched < 0 || sched > 2) return ILQC_E_ARG;
  if (sched == 0) sched = 1;
  if (sched == 2 && !regmode) return ILQC_E_UNSUPPORTED;  // direction oracles: lock-step driver only
  if (sched == 2)
    return solve_async_impl(m, a, B, x0, cmd, stepsize, n_iter, cost, gnorm, step_rep, n_done, status, ls_trips,
                            workspace, workspace_bytes, st);
  return solve_impl(m, a, B, x0, cmd, stepsize, n_iter, cost, gnorm, step_rep, n_done, status, ls_trips, workspace,
                    workspace_bytes, st);
}

extern "C" {

int ilqc_abi_version(void) { return ILQC_ABI_VERSION; }

int ilqc_dims(const ilqc_model_desc* m, int32_t* nx, int32_t* nu) {
  Dims d;
  int rc = get_dims(m, &d);
  if (rc) return rc;
  if (nx) *nx = d.nx;
  if (nu) *nu = d.nu;
  return 0;
}
int ilqc_lin_stride(const ilqc_model_desc* m) {
  Dims d;
  int rc = get_dims(m, &d);
  return rc ? rc : d.lin_stride;
}
int ilqc_gain_stride(const ilqc_model_desc* m) {
  Dims d;
  int rc = get_dims(m, &d);
  return rc ? rc : d.gain_stride;
}

int ilqc_rollout_cost(const ilqc_model_desc* m, int B, const double* x0, const double* cmd, double* traj,
                      double* cost_t, double* total, void* stream) {
  Dims d;
  int rc = get_dims(m, &d);
  if (rc) return rc;
  if (B <= 0 || !x0 || !cmd) return ILQC_E_ARG;
  ILQC_FOR_MODEL(effective_model(*m), return LM::forward(0, *m, B, B, nullptr, x0, cmd, nullptr, traj, cost_t, total, nullptr, (stream_t)stream));
  return 0;
}

int ilqc_linearize(const ilqc_model_desc* m, int B, const double* x0, const double* cmd, double* lin, double* traj,
                   double* total, void* stream) {
  Dims d;
  int rc = get_dims(m, &d);
  if (rc) return rc;
  if (B <= 0 || !x0 || !cmd || !lin) return ILQC_E_ARG;
  ILQC_FOR_MODEL(effective_model(*m), return LM::forward(1, *m, B, B, nullptr, x0, cmd, lin, traj, nullptr, total, nullptr, (stream_t)stream));
  return 0;
}

int ilqc_expand_dense(const ilqc_model_desc* m, int B, const double* x0, const double* cmd, int order, double* A,
                      double* Bm, double* p, double* P, double* q, double* Q, const double* lam, double* Hxx,
                      double* Hxu, double* Huu, void* stream) {
  Dims d;
  int rc = get_dims(m, &d);
  if (rc) return rc;
  if (B <= 0 || !x0 || !cmd || !A || !Bm || !p || !P || !q || !Q) return ILQC_E_ARG;
  if (order >= 2 && (!lam || !Hxx || !Hxu || !Huu)) return ILQC_E_ARG;
  ILQC_FOR_MODEL(effective_model(*m), return LM::expand_dense(*m, B, x0, cmd, order, A, Bm, p, P, q, Q, lam, Hxx, Hxu, Huu,
                                                   (stream_t)stream));
  return 0;
}

int ilqc_lq_backward_dense(int nx, int nu, int H, int B, int L, const double* A, const double* Bm, const double* p,
                           const double* P, const double* q, const double* Q, const double* R,
                           const double* reg_ctrl, double reg_state, double* K, double* k, double* jcst,
                           void* stream) {
  if (H <= 0 || B <= 0 || L <= 0 || !A || !Bm || !p || !P || !q || !Q || !reg_ctrl || !K || !k || !jcst)
    return ILQC_E_ARG;
  return dense_lq_backward(nx, nu, H, B, L, A, Bm, p, P, q, Q, R, reg_ctrl, reg_state, K, k, jcst, (stream_t)stream);
}

int ilqc_rollout_lin_dense(int nx, int nu, int H, int B, int L, const double* A, const double* Bm, const double* K,
                           const double* k, const double* step, double* diff, void* stream) {
  if (H <= 0 || B <= 0 || L <= 0 || !A || !Bm || !K || !k || !step || !diff) return ILQC_E_ARG;
  return dense_rollout_lin(nx, nu, H, B, L, A, Bm, K, k, step, diff, (stream_t)stream);
}

int ilqc_backward(const ilqc_model_desc* m, int mode, int B, int n_slots, const int32_t* prob,
                  const double* reg_ctrl, double reg_state, const double* lin, const double* traj,
                  const double* cmd, double* gains, double* jcst, int32_t* feasible, void* stream) {
  Dims d;
  int rc = get_dims(m, &d);
  if (rc) return rc;
  if (B <= 0 || n_slots <= 0 || !reg_ctrl || !lin || !gains || !jcst) return ILQC_E_ARG;
  const int flavour = mode & 0xf, bad_dir = mode >> ILQC_BWD_BAD_DIR_SHIFT;
  if (flavour < ILQC_BWD_LINQUAD || flavour > ILQC_BWD_QUAD_DDP || bad_dir < 0 || bad_dir > ILQC_BAD_DIR_LET_IT_BE) return ILQC_E_ARG;
  if (flavour != ILQC_BWD_LINQUAD && (!traj || !cmd)) return ILQC_E_ARG;
  ILQC_FOR_MODEL(effective_model(*m), return LM::backward(mode | (prob ? 0 : ILQC_BWD_DENSE_HINT), *m, B, n_slots, prob, nullptr, n_slots, reg_ctrl, reg_state, lin,
                                               traj, cmd, nullptr, gains, jcst, feasible, (stream_t)stream));
  return 0;
}

int ilqc_rollout(const ilqc_model_desc* m, int kind, int B, int n_slots, const int32_t* prob, const int32_t* gslot,
                 const double* step, const double* x0, const double* cmd, const double* traj, const double* lin,
                 const double* gains, const double* grad, double* diff, double* newval, double* diffsq,
                 void* stream) {
  Dims d;
  int rc = get_dims(m, &d);
  if (rc) return rc;
  if (B <= 0 || n_slots <= 0 || !step || !x0 || !cmd || !diff || !newval || !diffsq) return ILQC_E_ARG;
  if (kind == ILQC_ROLL_EXACT && (!traj || !gains)) return ILQC_E_ARG;
  if (kind == ILQC_ROLL_LIN && (!lin || !gains)) return ILQC_E_ARG;
  if (kind == ILQC_ROLL_GD && !grad) return ILQC_E_ARG;
  ILQC_FOR_MODEL(effective_model(*m), return LM::rollout(kind, *m, B, n_slots, prob, gslot, n_slots, nullptr, n_slots, step, x0, cmd,
                                              traj, lin, gains, grad, diff, newval, diffsq, (stream_t)stream));
  return 0;
}

int ilqc_grad_obj(const ilqc_model_desc* m, int B, const double* lin, double* grad, double* gnorm, double* cnorm,
                  void* stream) {
  Dims d;
  int rc = get_dims(m, &d);
  if (rc) return rc;
  if (B <= 0 || !lin || !gnorm) return ILQC_E_ARG;
  if (cnorm) ILQC_FOR_MODEL(effective_model(*m), { int rc2 = LM::cnorm_lin(*m, B, lin, cnorm, (stream_t)stream); if (rc2) return rc2; });
  ILQC_FOR_MODEL(effective_model(*m), return LM::adjoint(*m, B, B, nullptr, lin, grad, gnorm, (stream_t)stream));
  return 0;
}

size_t ilqc_solve_workspace_bytes(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B) {
  Dims d;
  if (get_dims(m, &d) || !a || B <= 0 || a->n_candidates < 1) return 0;
  Workspace w;
  carve(&w, nullptr, d, m->horizon, B, a->n_candidates, needs_hess(a), a->expansion == 1);
  return w.bytes;
}

int ilqc_solve(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, const double* x0, double* cmd,
               double* stepsize, int n_iter, double* cost, double* gnorm, double* step_rep, int32_t* n_done,
               int32_t* status, int32_t* ls_trips, void* workspace, size_t workspace_bytes, void* stream) {
  if (!m || !a || !x0 || !cmd || !stepsize || !cost || !gnorm || !step_rep || !n_done || !status) return ILQC_E_ARG;
  return solve_dispatch(m, a, B, x0, cmd, stepsize, n_iter, cost, gnorm, step_rep, n_done, status, ls_trips, workspace,
                        workspace_bytes, (stream_t)stream);
}

// per-kernel-class CUDA-event timing of ilqc_solve (classes: 0 roll-out+cost, 1 expansion, 2 backward,
// 3 candidate roll-out, 4 adjoint).  enable: 1/0 ; read returns cumulative ms / launches / slots and resets.
int ilqc_profile_enable(int enable) { prof_state().enabled = enable; return 0; }
int ilqc_set_tuning(const char* name, long long value) {
  if (!name) return ILQC_E_ARG;
  const Tuning def = tuning_defaults();   // (a negative value restores the default)
  if (!strcmp(name, "coop_max")) tuning().coop_max = value < 0 ? def.coop_max : value;
  else if (!strcmp(name, "feed_small")) tuning().feed_small = value < 0 ? def.feed_small : value;
  else return ILQC_E_ARG;
  return 0;
}
int ilqc_profile_read(double* ms, int64_t* launches, int64_t* units, int n) {
  if (!ms || !launches || !units || n < PROF_N) return ILQC_E_ARG;
  ProfState& p = prof_state();
  for (int i = 0; i < PROF_N; ++i) {
    ms[i] = p.ms[i]; launches[i] = p.launches[i]; units[i] = p.units[i];
    p.ms[i] = 0; p.launches[i] = 0; p.units[i] = 0;
  }
  return PROF_N;
}

int ilqc_fastmath_eval(int fn, int n, const double* a, const double* b, double* o0, double* o1, void* stream) {
  if (fn < 0 || fn > 5 || n <= 0 || !a || !o0 || ((fn == 1 || fn == 3) && !b)) return ILQC_E_ARG;
  ILQC_LAUNCH((fastmath_eval_kernel), n, 256, (stream_t)stream, fn, n, a, b, o0, o1);
  return ILQC_LAUNCH_OK();
}

long long ilqc_launch_count(int reset) {
  const long long n = launch_counter();
  if (reset) launch_counter() = 0;
  return n;
}

int ilqc_to_soa(const double* src, double* dst, int B, int n, void* stream) {
  if (!src || !dst || B <= 0 || n <= 0) return ILQC_E_ARG;
  ILQC_LAUNCH((to_soa_kernel), B * n, 256, (stream_t)stream, src, dst, B, n);
  return ILQC_LAUNCH_OK();
}
int ilqc_from_soa(const double* src, double* dst, int B, int n, void* stream) {
  if (!src || !dst || B <= 0 || n <= 0) return ILQC_E_ARG;
  ILQC_LAUNCH((from_soa_kernel), B * n, 256, (stream_t)stream, src, dst, B, n);
  return ILQC_LAUNCH_OK();
}

// ---- host-buffer entry point ---------------------------------------------------------------------------
// device staging: x0 [B][nx] + cmd [B][H][nu] + stepsize + metrics (3 x [n_iter+1][B]) in both layouts
static size_t host_stage_doubles(const Dims& d, int H, int B, int n_iter) {
  return (size_t)B * d.nx * 2 + (size_t)B * H * d.nu * 2 + (size_t)B + (size_t)3 * (n_iter + 1) * B * 2;
}
size_t ilqc_solve_host_workspace_bytes(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, int n_iter) {
  Dims d;
  if (get_dims(m, &d) || !a || B <= 0) return 0;
  const size_t ws = ilqc_solve_workspace_bytes(m, a, B);
  return ws + align_up(host_stage_doubles(d, m->horizon, B, n_iter) * sizeof(double)) + align_up((size_t)2 * B * 4) + 512;
}
int ilqc_solve_host(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, const double* x0_host,
                    const double* cmd0_host, double* cmd_host, double* stepsize_host, int n_iter, double* cost_host, double* gnorm_host, double* step_rep_host,
                    int32_t* n_done_host, int32_t* status_host, void* workspace, size_t workspace_bytes,
                    void* stream) {
  Dims d;
  int rc = get_dims(m, &d);
  if (rc) return rc;
  if (!a || B <= 0 || !x0_host || !cmd_host || !stepsize_host || !cost_host || !status_host) return ILQC_E_ARG;
  if (workspace_bytes < ilqc_solve_host_workspace_bytes(m, a, B, n_iter)) return ILQC_E_WORKSPACE;
  const int H = m->horizon;
  stream_t st = (stream_t)stream;
  const size_t ws = align_up(ilqc_solve_workspace_bytes(m, a, B));
  char* base = (char*)workspace + ws;
  double* x0_aos = (double*)base;
  double* x0_soa = x0_aos + (size_t)B * d.nx;
  double* cmd_aos = x0_soa + (size_t)B * d.nx;
  double* cmd_soa = cmd_aos + (size_t)B * H * d.nu;
  double* step_d = cmd_soa + (size_t)B * H * d.nu;
  const size_t mrows = (size_t)(n_iter + 1) * B;
  double* met_soa = step_d + B;      // cost, gnorm, step_rep  [n_iter+1][B]
  double* met_aos = met_soa + 3 * mrows;  // [B][n_iter+1]
  int32_t* ints = (int32_t*)(base + align_up(host_stage_doubles(d, H, B, n_iter) * sizeof(double)));
  int32_t* n_done_d = ints;
  int32_t* status_d = ints + B;
#define ILQC_CHECK(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
  ILQC_CHECK(dev_h2d(x0_aos, x0_host, sizeof(double) * B * d.nx, st));
  ILQC_CHECK(dev_h2d(step_d, stepsize_host, sizeof(double) * B, st));
  ILQC_CHECK(ilqc_to_soa(x0_aos, x0_soa, B, d.nx, stream));
  if (cmd0_host) {
    ILQC_CHECK(dev_h2d(cmd_aos, cmd0_host, sizeof(double) * B * H * d.nu, st));
    ILQC_CHECK(ilqc_to_soa(cmd_aos, cmd_soa, B, H * d.nu, stream));
  } else {
    ILQC_CHECK(dev_memset(cmd_soa, 0, sizeof(double) * B * H * d.nu, st));  // prev_cmd=None: zeros (run_min_algo.py:57)
  }
  ILQC_CHECK(solve_dispatch(m, a, B, x0_soa, cmd_soa, step_d, n_iter, met_soa, met_soa + mrows, met_soa + 2 * mrows,
                            n_done_d, status_d, nullptr, workspace, ws, st));
  ILQC_CHECK(ilqc_from_soa(cmd_soa, cmd_aos, B, H * d.nu, stream));
  for (int k = 0; k < 3; ++k) ILQC_CHECK(ilqc_from_soa(met_soa + k * mrows, met_aos + k * mrows, B, n_iter + 1, stream));
  ILQC_CHECK(dev_d2h(cmd_host, cmd_aos, sizeof(double) * B * H * d.nu, st));
  ILQC_CHECK(dev_d2h(stepsize_host, step_d, sizeof(double) * B, st));
  ILQC_CHECK(dev_d2h(cost_host, met_aos, sizeof(double) * mrows, st));
  if (gnorm_host) ILQC_CHECK(dev_d2h(gnorm_host, met_aos + mrows, sizeof(double) * mrows, st));
  if (step_rep_host) ILQC_CHECK(dev_d2h(step_rep_host, met_aos + 2 * mrows, sizeof(double) * mrows, st));
  if (n_done_host) ILQC_CHECK(dev_d2h(n_done_host, n_done_d, sizeof(int32_t) * B, st));
  ILQC_CHECK(dev_d2h(status_host, status_d, sizeof(int32_t) * B, st));
  ILQC_CHECK(dev_sync(st));
#undef ILQC_CHECK
  return 0;
}

// ---- closed-loop MPC episodes ----------------------------------------------------------------------------
int ilqc_mpc_num_windows(int full_horizon, int window_size, int overlap) {
  if (full_horizon <= 0 || window_size <= 0 || overlap < 0 || overlap >= window_size) return ILQC_E_ARG;
  int n = 0;
  for (int h = 0; h < full_horizon; ++n) h = mpc_next_horizon(h, window_size, overlap, full_horizon);
  return n;
}
// walks the window sequence on the host: longest solve horizon (and validity of every window)
static int mpc_max_window(int first_h, int full_horizon, int window_size, int overlap, bool full) {
  int curr = 0, maxH = first_h;
  for (int h = 0; h < full_horizon;) {
    h = mpc_next_horizon(h, window_size, overlap, full_horizon);
    if (curr == 0) { curr = first_h; continue; }
    MpcWindow w;
    if (!mpc_window(curr, h, overlap, full, &w) || w.Hw <= 0) return -1;
    if (w.Hw > maxH) maxH = w.Hw;
    curr = w.keep + w.Hw;
  }
  return maxH;
}
enum { kMpcAdvanceChunk = 16 };
struct MpcCarve { size_t solve, wcmd, xfrozen, traj, step, met, ints, total; };
static bool mpc_carve(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, int full_horizon, int window_size,
                      int overlap, int max_iter, int full, MpcCarve* c, int* maxH_out) {
  Dims d;
  if (get_dims(m, &d) || !a || B <= 0 || max_iter < 0) return false;
  if (ilqc_mpc_num_windows(full_horizon, window_size, overlap) <= 0) return false;
  const int maxH = mpc_max_window(m->horizon, full_horizon, window_size, overlap, full != 0);
  if (maxH <= 0) return false;
  ilqc_model_desc md = *m;
  md.horizon = maxH;
  size_t off = 0;
  c->solve = off; off += align_up(ilqc_solve_workspace_bytes(&md, a, B));
  c->wcmd = off; off += align_up(sizeof(double) * (size_t)maxH * d.nu * B);
  c->xfrozen = off; off += align_up(sizeof(double) * (size_t)d.nx * B);
  c->traj = off; off += align_up(sizeof(double) * (size_t)(kMpcAdvanceChunk + 1) * d.nx * B);
  c->step = off; off += align_up(sizeof(double) * (size_t)B);
  c->met = off; off += align_up(sizeof(double) * (size_t)3 * (max_iter + 1) * B);
  c->ints = off; off += align_up(sizeof(int32_t) * (size_t)2 * B);
  c->total = off;
  if (maxH_out) *maxH_out = maxH;
  return true;
}
size_t ilqc_mpc_workspace_bytes(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, int full_horizon, int window_size,
                                int overlap, int max_iter, int optim_on_full_window) {
  MpcCarve c;
  return mpc_carve(m, a, B, full_horizon, window_size, overlap, max_iter, optim_on_full_window, &c, nullptr) ? c.total : 0;
}
int ilqc_mpc(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, const double* x0, int full_horizon, int window_size,
             int overlap, int max_iter, int optim_on_full_window, int keep_applying_last_ctrl, int max_windows, double* cmd,
             int32_t* cmd_len_host, double* cost, int32_t* status, int32_t* n_done, void* workspace, size_t workspace_bytes,
             void* stream) {
  Dims d;
  int rc = get_dims(m, &d);
  if (rc) return rc;
  if (!a || !x0 || !cmd || B <= 0 || m->pp_t0) return ILQC_E_ARG;
  MpcCarve c;
  int maxH = 0;
  if (!mpc_carve(m, a, B, full_horizon, window_size, overlap, max_iter, optim_on_full_window, &c, &maxH)) return ILQC_E_ARG;
  if (!workspace || workspace_bytes < c.total) return ILQC_E_WORKSPACE;
  stream_t st = (stream_t)stream;
  char* base = (char*)workspace;
  double* wcmd = (double*)(base + c.wcmd);
  double* xfrozen = (double*)(base + c.xfrozen);
  double* trajs = (double*)(base + c.traj);
  double* step = (double*)(base + c.step);
  double* met = (double*)(base + c.met);
  int32_t* ints = (int32_t*)(base + c.ints);
  const size_t row = (size_t)d.nu * B, mrows = (size_t)(max_iter + 1) * B;
  const bool full = optim_on_full_window != 0;
  const bool regstep = a->algo_type != ILQC_ALGO_GD && (a->step_mode == ILQC_STEP_REG || a->step_mode == ILQC_STEP_REGVAR);
  const double step0 = regstep ? 100.0 : 1.0;  // run_min_algo's default first stepsize (run_min_algo.py:59)
#define ILQC_CHECK(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
  int curr = 0, k_frozen = 0, n_win = 0;
  const double* xf = x0;  // state after the first k_frozen controls of cmd
  for (int h = 0; h < full_horizon;) {
    h = mpc_next_horizon(h, window_size, overlap, full_horizon);
    ilqc_model_desc md = *m;
    const double* xw = x0;
    MpcWindow w;
    if (curr == 0) {  // first window: run_min_algo(env, max_iter, algo) on env.horizon zeros
      w.Hw = m->horizon; w.keep = 0; w.a = 0;
      ILQC_CHECK(dev_memset(wcmd, 0, sizeof(double) * (size_t)w.Hw * row, st));
    } else {
      if (!mpc_window(curr, h, overlap, full, &w)) return ILQC_E_ARG;
      // warm start: tail of the previous controls, then copies of the last control (or zeros)
      if (w.n_tail > 0) ILQC_CHECK(dev_d2d(wcmd, cmd + (size_t)w.tail_start * row, sizeof(double) * (size_t)w.n_tail * row, st));
      if (w.extra > 0) {
        const int n = w.extra * (int)row;
        ILQC_LAUNCH((mpc_pad_kernel), n, 256, st, n, (int)row, cmd + (size_t)(curr - 1) * row, wcmd + (size_t)w.n_tail * row, keep_applying_last_ctrl);
      }
      if (!full) {
        // re-anchor: state after a_pos controls of the previous solution (the plant is the model)
        if (w.a_pos < k_frozen) { k_frozen = 0; xf = x0; }
        while (k_frozen < w.a_pos) {
          const int n = w.a_pos - k_frozen < kMpcAdvanceChunk ? w.a_pos - k_frozen : (int)kMpcAdvanceChunk;
          ilqc_model_desc ma = *m;
          ma.horizon = n; ma.t0 = k_frozen;
          ILQC_FOR_MODEL(effective_model(ma), ILQC_CHECK(LM::forward(0, ma, B, B, nullptr, xf, cmd + (size_t)k_frozen * row, nullptr, trajs, nullptr, nullptr, nullptr, st)));
          ILQC_CHECK(dev_d2d(xfrozen, trajs + (size_t)n * d.nx * B, sizeof(double) * (size_t)d.nx * B, st));
          xf = xfrozen;
          k_frozen += n;
        }
        xw = xf;
        md.t0 = w.a;
      } else {
        md.t0 = m->t0;
      }
    }
    md.horizon = w.Hw;
    ILQC_LAUNCH((fill_kernel), B, 256, st, B, step, step0);
    double* cost_w = cost ? cost + (size_t)n_win * mrows : met;
    int32_t* status_w = status ? status + (size_t)n_win * B : ints;
    int32_t* n_done_w = n_done ? n_done + (size_t)n_win * B : ints + B;
    ILQC_CHECK(solve_dispatch(&md, a, B, xw, wcmd, step, max_iter, cost_w, met + mrows, met + 2 * mrows, n_done_w, status_w,
                              nullptr, base + c.solve, c.wcmd - c.solve, st));
    // cmd_opt = prev_cmd[:-overlap] ++ cmd
    ILQC_CHECK(dev_d2d(cmd + (size_t)w.keep * row, wcmd, sizeof(double) * (size_t)w.Hw * row, st));
    curr = w.keep + w.Hw;
    ++n_win;
    if (max_windows > 0 && n_win >= max_windows) break;
  }
  ILQC_CHECK(dev_sync(st));
#undef ILQC_CHECK
  if (cmd_len_host) *cmd_len_host = curr;
  return ILQC_LAUNCH_OK();
}

// ---- dense Newton step / trajectory Jacobian (diagnostics of the reference) --------------------------------
size_t ilqc_newton_dense_workspace_bytes(const ilqc_model_desc* m, int B) {
  Dims d;
  if (get_dims(m, &d) || B <= 0 || d.nx > kNdMaxNx || d.nu > kNdMaxNu) return 0;
  NdCarve c;
  nd_carve(&c, d.nx, d.nu, m->horizon, B, true);
  return c.total;
}
int ilqc_newton_dense(const ilqc_model_desc* m, int B, const double* x0, const double* cmd, const double* reg, double* dir,
                      double* opt, double* grad, double* hess, void* workspace, size_t workspace_bytes, void* stream) {
  Dims d;
  int rc = get_dims(m, &d);
  if (rc) return rc;
  if (B <= 0 || !x0 || !cmd || !reg || !dir || !opt) return ILQC_E_ARG;
  if (d.nx > kNdMaxNx || d.nu > kNdMaxNu) return ILQC_E_UNSUPPORTED;
  const int H = m->horizon, n = H * d.nu;
  if ((long long)B * n > 0x7fffffffLL) return ILQC_E_ARG;
  NdCarve c;
  nd_carve(&c, d.nx, d.nu, H, B, true);
  if (!workspace || workspace_bytes < c.total) return ILQC_E_WORKSPACE;
  stream_t st = (stream_t)stream;
  char* base = (char*)workspace;
  auto at = [&](size_t off) { return (double*)(base + off); };
  double* g = at(c.grad);
  double* hs = at(c.hess);
#define ILQC_CHECK(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
  // first-order pass: A, B, p, P, q, Q ; costate and gradient ; second pass: dynamics Hessians contracted with it
  ILQC_FOR_MODEL(effective_model(*m), ILQC_CHECK(LM::expand_dense(*m, B, x0, cmd, 1, at(c.A), at(c.Bm), at(c.p), at(c.P), at(c.q), at(c.Q),
                                                           nullptr, nullptr, nullptr, nullptr, st)));
  NdDense E{d.nx, d.nu, H, B, at(c.A), at(c.Bm), at(c.p), at(c.P), at(c.q), at(c.Q), at(c.Hxx), at(c.Hxu), at(c.Huu)};
  ILQC_LAUNCH((nd_costate_kernel), B, 64, st, E, at(c.lam), g);
  ILQC_FOR_MODEL(effective_model(*m), ILQC_CHECK(LM::expand_dense(*m, B, x0, cmd, 2, at(c.A), at(c.Bm), at(c.p), at(c.P), at(c.q), at(c.Q),
                                                           at(c.lam), at(c.Hxx), at(c.Hxu), at(c.Huu), st)));
  ILQC_LAUNCH((nd_tangent_kernel), B * n, 128, st, E, at(c.D), (double*)nullptr);
  ILQC_LAUNCH((nd_hess_kernel), B * n, 128, st, E, at(c.D), hs);
#if defined(ILQC_HOSTSIM)
  ILQC_LAUNCH((nd_solve_kernel), B, 1, st, B, n, d.nu, hs, g, reg, at(c.M), at(c.rhs), dir, opt);
#else
  nd_solve_kernel<<<B, kNdSolveThreads, 0, st>>>(B, n, d.nu, hs, g, reg, at(c.M), at(c.rhs), dir, opt);
  ++launch_counter();
#endif
  if (grad) ILQC_CHECK(dev_d2d(grad, g, sizeof(double) * (size_t)n * B, st));
  if (hess) ILQC_CHECK(dev_d2d(hess, hs, sizeof(double) * (size_t)n * n * B, st));
#undef ILQC_CHECK
  return ILQC_LAUNCH_OK();
}
size_t ilqc_traj_jacobian_workspace_bytes(const ilqc_model_desc* m, int B) {
  Dims d;
  if (get_dims(m, &d) || B <= 0 || d.nx > kNdMaxNx || d.nu > kNdMaxNu) return 0;
  NdCarve c;
  nd_carve(&c, d.nx, d.nu, m->horizon, B, false);
  return c.total;
}
int ilqc_traj_jacobian(const ilqc_model_desc* m, int B, const double* x0, const double* cmd, double* jac, void* workspace,
                       size_t workspace_bytes, void* stream) {
  Dims d;
  int rc = get_dims(m, &d);
  if (rc) return rc;
  if (B <= 0 || !x0 || !cmd || !jac) return ILQC_E_ARG;
  if (d.nx > kNdMaxNx || d.nu > kNdMaxNu) return ILQC_E_UNSUPPORTED;
  const int H = m->horizon, n = H * d.nu;
  if ((long long)B * n > 0x7fffffffLL) return ILQC_E_ARG;
  NdCarve c;
  nd_carve(&c, d.nx, d.nu, H, B, false);
  if (!workspace || workspace_bytes < c.total) return ILQC_E_WORKSPACE;
  stream_t st = (stream_t)stream;
  char* base = (char*)workspace;
  auto at = [&](size_t off) { return (double*)(base + off); };
  ILQC_FOR_MODEL(effective_model(*m), { int rc2 = LM::expand_dense(*m, B, x0, cmd, 1, at(c.A), at(c.Bm), at(c.p), at(c.P), at(c.q), at(c.Q),
                                                                nullptr, nullptr, nullptr, nullptr, st); if (rc2) return rc2; });
  NdDense E{d.nx, d.nu, H, B, at(c.A), at(c.Bm), at(c.p), at(c.P), at(c.q), at(c.Q), nullptr, nullptr, nullptr};
  ILQC_LAUNCH((nd_tangent_kernel), B * n, 128, st, E, (double*)nullptr, jac);
  return ILQC_LAUNCH_OK();
}

}  // extern "C"

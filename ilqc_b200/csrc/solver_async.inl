// Asynchronous batched run_min_algo for the regularised step modes ('reg', 'regvar') and gd.
//
// Same per-problem arithmetic and decisions as the lock-step driver of solver.inl (bit-identical results,
// tests/test_host_logic.py::test_schedulers_agree), different scheduling: the problems do NOT advance
// iteration by iteration together.  Every problem is a small state machine
//     LIN  (needs its re-expansion + metrics row)  ->  LS (searching its step)  ->  LIN -> ... -> DONE
// and every "tick" of the driver serves all problems at once, whatever their iteration:
//     stream A: candidates of the problems in LS      (make_slots, backward, roll-out, select, apply)
//     stream B: re-expansion of the problems in LIN   (forward<1>, adjoint, status + next line-search set-up)
// The two groups touch disjoint problems, so the streams run concurrently.  In the lock-step driver an
// iteration lasts as long as its slowest line search (7 rounds for a median of 4 trips on the bicycle
// workload) and the expansion kernels run alone on a mostly idle GPU; here a problem that accepted its step
// starts its next iteration on the next tick.  Reference semantics per problem are unchanged
// (algorithms/run_min_algo.py:78-96, algorithms/algos_steps.py:277-437).
#pragma once
#include "solver.inl"

namespace ilqc {

#if !defined(ILQC_HOSTSIM)
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
};
inline SideStream& side_stream() {
  static SideStream ss[16];
  int dev = 0;
  cudaGetDevice(&dev);
  SideStream& x = ss[dev & 15];
  if (!x.s) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    const char* e = getenv("ILQC_SIDE_PRIO");  // tuning knob: 1 = side stream at the highest priority
    cudaStreamCreateWithPriority(&x.s, cudaStreamNonBlocking, (e && atoi(e) == 1) ? hi : lo);
    cudaEventCreateWithFlags(&x.ev_a, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&x.ev_b, cudaEventDisableTiming);
  }
  return x;
}
#endif

inline int solve_async_impl(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, const double* x0, double* cmd,
                            double* stepsize, int n_iter, double* cost, double* gnorm_out, double* step_rep,
                            int32_t* n_done, int32_t* status, int32_t* ls_trips, void* workspace,
                            size_t workspace_bytes, stream_t st) {
  Dims d;
  int rc = get_dims(m, &d);
  if (rc) return rc;
  if (!a || B <= 0 || n_iter < 0 || a->n_candidates < 1) return ILQC_E_ARG;
  const int H = m->horizon, L = a->n_candidates;
  Workspace w;
  if (a->expansion != 0) return ILQC_E_UNSUPPORTED;  // the time-parallel expansion belongs to the lock-step driver
  carve(&w, (char*)workspace, d, H, B, L, needs_hess(a));
  if (!workspace || workspace_bytes < w.bytes) return ILQC_E_WORKSPACE;

  ProfScope prof(st);
  LsParams P;
  P.algo_type = a->algo_type; P.approx = a->approx; P.step_mode = a->step_mode; P.L = L;
  P.line_search = a->line_search; P.H = H; P.nu = d.nu;
  if (a->algo_type == ILQC_ALGO_GD) P.step_mode = ILQC_STEP_REGVAR;
  const bool regmode = (P.step_mode == ILQC_STEP_REG || P.step_mode == ILQC_STEP_REGVAR);
  if (!regmode) return ILQC_E_UNSUPPORTED;  // direction oracles use the lock-step driver
  P.rho = 0.5;
  const int max_trips = a->max_ls_trips > 0 ? a->max_ls_trips : 2000;
  const bool quad = a->approx == ILQC_APPROX_QUAD && a->algo_type != ILQC_ALGO_GD;
  if (a->handle_bad_dir < ILQC_BAD_DIR_CHECK_TOTAL || a->handle_bad_dir > ILQC_BAD_DIR_LET_IT_BE) return ILQC_E_ARG;
  const int bwd_mode = (!quad ? ILQC_BWD_LINQUAD
                              : (a->algo_type == ILQC_ALGO_DDP ? ILQC_BWD_QUAD_DDP : ILQC_BWD_QUAD_NEWTON)) |
                       (a->handle_bad_dir << ILQC_BWD_BAD_DIR_SHIFT);
  const bool need_grad = a->compute_grad_norm || a->algo_type == ILQC_ALGO_GD;
  const ilqc_model_desc& md = *m;
  AsyncState A{w.phase, w.iter, ls_trips, n_done, cost, gnorm_out, step_rep};

#if defined(ILQC_HOSTSIM)
  stream_t sb = st;
#else
  SideStream& ss = side_stream();
  stream_t sb = ss.s;
#endif
#define ILQC_CHECK(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

  // stragglers allowed to lag behind the bulk: ~2 % of the batch, at least one block's worth
  int lag = B / 50 < 64 ? 64 : B / 50;
  if (const char* e = getenv("ILQC_LAG")) lag = atoi(e);  // tuning knob
  if (lag < 0) lag = 0;
  const int n_rows = (n_iter + 1) * B;
  ILQC_LAUNCH((fill_i32_kernel), B, 256, st, B, status, (int32_t)ILQC_RUNNING);
  ILQC_LAUNCH((fill_i32_kernel), B, 256, st, B, n_done, (int32_t)0);
  ILQC_LAUNCH((fill_i32_kernel), B, 256, st, B, w.phase, (int32_t)PHASE_LIN);
  ILQC_LAUNCH((fill_i32_kernel), B, 256, st, B, w.iter, (int32_t)0);
  ILQC_LAUNCH((fill_i32_kernel), B, 256, st, B, w.feas_prob, (int32_t)1);
  ILQC_LAUNCH((fill_kernel), B, 256, st, B, w.gnorm, (double)NAN);
  ILQC_LAUNCH((fill_kernel), n_rows, 256, st, n_rows, cost, (double)NAN);       // rows after a stop: NaN padding
  ILQC_LAUNCH((fill_kernel), n_rows, 256, st, n_rows, gnorm_out, (double)NAN);
  ILQC_LAUNCH((fill_kernel), n_rows, 256, st, n_rows, step_rep, (double)NAN);
  if (ls_trips && n_iter > 0) ILQC_LAUNCH((fill_i32_kernel), n_iter * B, 256, st, n_iter * B, ls_trips, (int32_t)0);
  // objective of the initial controls with the arithmetic of every candidate evaluation
  ILQC_PROF(PROF_FORWARD0, B, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::forward(0, md, B, B, nullptr, x0, cmd, nullptr, nullptr, nullptr, w.curr_val, nullptr, st))));

  for (;;) {
    ILQC_LAUNCH((plan_async_kernel), 1, 1024, st, B, w.phase, w.trips, L, (int)slot_capacity(B, L), w.list, w.slot_cnt, w.slot_off,
                w.ident, w.count);
    int counts[3] = {0, 0, 0};
    ILQC_CHECK(dev_d2h_sync(counts, w.count, 3 * sizeof(int), st));
    const int n_ls = counts[0], n_lin = counts[1], n_slots = counts[2];
    if (n_ls == 0 && n_lin == 0) break;
    // Bulk-synchronous with deferred stragglers: the accepted problems wait for their (full-size, efficient)
    // re-expansion until at most `lag` problems are still searching; those stragglers then keep searching on
    // stream A while stream B re-expands the bulk, and join the bulk's next rounds with their own (wider)
    // candidate counts, one iteration behind.  The lock-step driver instead spends 2 latency-bound rounds per
    // iteration on them with the rest of the GPU idle.
    const bool do_lin = n_lin > 0 && (n_ls <= lag || n_ls == 0);
    if (trace_rounds()) fprintf(stderr, "ilqc tick: t %.3f ms ls %d lin %d slots %d do_lin %d\n", trace_ms(), n_ls, n_lin, n_slots, (int)do_lin);
#if !defined(ILQC_HOSTSIM)
    cudaEventRecord(ss.ev_a, st);
    cudaStreamWaitEvent(sb, ss.ev_a, 0);
#endif
    // ---- stream B: re-expansion, gradient, metrics row and next line-search set-up of the LIN problems ----
    if (do_lin) {
      ILQC_PROF_ON(sb, PROF_LINEARIZE, n_lin, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::forward(1, md, B, n_lin, w.ident, x0, cmd, w.lin, w.traj, nullptr, nullptr, w.cnorm, sb))));
      if (quad) ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::hess(md, B, n_lin, w.ident, w.traj, cmd, w.hess, sb)));
      if (need_grad) ILQC_PROF_ON(sb, PROF_ADJOINT, n_lin, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::adjoint(md, B, n_lin, w.ident, w.lin, w.dir, w.gnorm, sb))));
      ILQC_LAUNCH((status_async_kernel), n_lin, 256, sb, P, n_lin, w.ident, B, n_iter, w.curr_val, w.gnorm, w.cnorm, stepsize,
                  w.s_try, w.dec_unit, w.trips, status, A, a->pp_cost0);
    }
    // ---- stream A: line-search candidates of the LS problems ----
    // (ILQC_DEFER=1: the stragglers skip the re-expansion tick and continue inside the bulk's next big launch)
    static const bool defer = getenv("ILQC_DEFER") != nullptr;
    if (n_ls > 0 && !(defer && do_lin)) {
      P.L = 0;
      P.slot_cnt = w.slot_cnt;
      P.slot_off = w.slot_off;
      ILQC_LAUNCH((make_slots_kernel), n_ls, 256, st, P, n_ls, w.list, w.s_try, w.prob, w.gslot, w.step_true, w.step_roll, w.reg);
      if (a->algo_type == ILQC_ALGO_GD) {
        ILQC_PROF(PROF_ROLLOUT, n_slots, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::rollout(ILQC_ROLL_GD, md, B, n_slots, w.prob, nullptr, 0, nullptr, n_slots, w.step_roll, x0,
                                                        cmd, w.traj, w.lin, nullptr, w.dir, w.diff, w.newval, w.diffsq, st))));
      } else {
        ILQC_PROF(PROF_BACKWARD, n_slots, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::backward(bwd_mode | ((size_t)n_ls * 10 >= (size_t)B * 9 ? ILQC_BWD_DENSE_HINT : 0), md, B, n_slots, w.prob, nullptr, n_slots, w.reg, 0.0, w.lin, w.traj, cmd,
                                                         w.hess, w.gains, w.jcst, w.feas_slot, st))));
        const int kind = a->algo_type == ILQC_ALGO_DDP ? ILQC_ROLL_EXACT : ILQC_ROLL_LIN;
        ILQC_PROF(PROF_ROLLOUT, n_slots, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::rollout(kind, md, B, n_slots, w.prob, nullptr, n_slots, nullptr, n_slots, w.step_roll, x0, cmd,
                                                        w.traj, w.lin, w.gains, nullptr, w.diff, w.newval, w.diffsq, st))));
      }
      ILQC_LAUNCH((ls_select_kernel), n_ls, 128, st, P, n_ls, B, n_slots, w.list, w.step_true, w.newval, w.diffsq, w.jcst,
                  w.dec_unit, w.feas_slot, w.feas_prob, w.cnorm, w.curr_val, stepsize, w.s_try, status, w.flags, w.trips,
                  w.accept, max_trips, A);
      ILQC_LAUNCH((apply_kernel), n_ls * (H * d.nu), 256, st, n_ls, H * d.nu, B, n_slots, w.list, w.accept, w.diff, cmd);
    }
#if !defined(ILQC_HOSTSIM)
    cudaEventRecord(ss.ev_b, sb);
    cudaStreamWaitEvent(st, ss.ev_b, 0);
#endif
  }
  ILQC_CHECK(dev_sync(st));
  prof.flush();
  return ILQC_LAUNCH_OK();
#undef ILQC_CHECK
}

}  // namespace ilqc

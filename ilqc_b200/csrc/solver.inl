// Batched run_min_algo: the outer loop, the per-problem line-search state machine and the metrics /
// convergence bookkeeping of the reference, for B independent problems at once.
//
// Reference semantics restated (file:line in /root/reference):
//   run_min_algo, collect_info, check_cvg_status     algorithms/run_min_algo.py:15-100, :127-264
//   min_step (line-search set-up per step mode)      algorithms/algos_steps.py:277-381
//   backtracking_line_search                         algorithms/algos_steps.py:384-437
//   gd_oracle / classic_oracle / ddp_oracle          algorithms/algos_steps.py:22-43, :105-274
//
// B200-first structure: every problem is a row of a structure-of-arrays workspace resident in HBM; one
// solver iteration is a handful of launches (expansion, L line-search candidates per still-searching
// problem = backward + roll-out, selection, adjoint, status) with thread-per-(problem,candidate)
// kernels.  Problems whose line search needs more trips are compacted into a dense index list so later
// rounds launch only the threads that still have work.  The only host<->device traffic inside the loop
// is one 4-byte counter per line-search round.
#pragma once
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "launch.cuh"

namespace ilqc {

// ---- memory plumbing (stream-ordered) -----------------------------------------------------------------
#if defined(ILQC_HOSTSIM)
inline int dev_memset(void* p, int v, size_t n, stream_t) { memset(p, v, n); return 0; }
inline int dev_d2h_sync(void* dst, const void* src, size_t n, stream_t) { memcpy(dst, src, n); return 0; }
inline int dev_h2d(void* dst, const void* src, size_t n, stream_t) { memcpy(dst, src, n); return 0; }
inline int dev_d2h(void* dst, const void* src, size_t n, stream_t) { memcpy(dst, src, n); return 0; }
inline int dev_d2d(void* dst, const void* src, size_t n, stream_t) { memcpy(dst, src, n); return 0; }
inline int dev_sync(stream_t) { return 0; }
#else
inline int dev_memset(void* p, int v, size_t n, stream_t st) { return (int)cudaMemsetAsync(p, v, n, st); }
inline int dev_d2h_sync(void* dst, const void* src, size_t n, stream_t st) {
  cudaError_t e = cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaStreamSynchronize(st);
}
inline int dev_h2d(void* dst, const void* src, size_t n, stream_t st) {
  return (int)cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, st);
}
inline int dev_d2h(void* dst, const void* src, size_t n, stream_t st) {
  return (int)cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, st);
}
inline int dev_d2d(void* dst, const void* src, size_t n, stream_t st) {
  return (int)cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, st);
}
inline int dev_sync(stream_t st) { return (int)cudaStreamSynchronize(st); }
#endif

// ---- model dispatch -----------------------------------------------------------------------------------
// effective model id: public model id, + ILQC_MODEL_VAR_OFFSET for the three-control-set `rk4` integrator
#ifndef ILQC_MODEL_VAR_OFFSET
#define ILQC_MODEL_VAR_OFFSET 4
#endif
inline int effective_model(const ilqc_model_desc& m) {
  return m.model + (m.integrator == ILQC_INT_RK4 ? ILQC_MODEL_VAR_OFFSET : 0);
}
#define ILQC_FOR_MODEL(model_id, CALL)                                   \
  switch (model_id) {                                                    \
    case ILQC_MODEL_PENDULUM: { using LM = Launch<ILQC_MODEL_PENDULUM>; CALL; } break;       \
    case ILQC_MODEL_CARTPOLE: { using LM = Launch<ILQC_MODEL_CARTPOLE>; CALL; } break;       \
    case ILQC_MODEL_CAR_SIMPLE: { using LM = Launch<ILQC_MODEL_CAR_SIMPLE>; CALL; } break;   \
    case ILQC_MODEL_CAR_BICYCLE: { using LM = Launch<ILQC_MODEL_CAR_BICYCLE>; CALL; } break; \
    case ILQC_MODEL_CARTPOLE + ILQC_MODEL_VAR_OFFSET: { using LM = Launch<ILQC_MODEL_CARTPOLE + ILQC_MODEL_VAR_OFFSET>; CALL; } break;       \
    case ILQC_MODEL_CAR_SIMPLE + ILQC_MODEL_VAR_OFFSET: { using LM = Launch<ILQC_MODEL_CAR_SIMPLE + ILQC_MODEL_VAR_OFFSET>; CALL; } break;   \
    case ILQC_MODEL_CAR_BICYCLE + ILQC_MODEL_VAR_OFFSET: { using LM = Launch<ILQC_MODEL_CAR_BICYCLE + ILQC_MODEL_VAR_OFFSET>; CALL; } break; \
    default: return ILQC_E_ARG;                                          \
  }

// ---- optional per-kernel-class timing with CUDA events (ilqc_profile_*; used by bench.py for the roofline) ----
enum { PROF_FORWARD0 = 0, PROF_LINEARIZE = 1, PROF_BACKWARD = 2, PROF_ROLLOUT = 3, PROF_ADJOINT = 4, PROF_N = 5 };
struct ProfState {
  int enabled = 0;
  double ms[PROF_N] = {0, 0, 0, 0, 0};
  long long launches[PROF_N] = {0, 0, 0, 0, 0};
  long long units[PROF_N] = {0, 0, 0, 0, 0};  // slots (threads with work) summed over launches
};
inline ProfState& prof_state() { static ProfState s; return s; }
#if defined(ILQC_HOSTSIM)
struct ProfScope { ProfScope(stream_t) {} void begin(int, long long) {} void begin_on(int, long long, stream_t) {} void end() {} void flush() {} };
#else
struct ProfScope {
  struct Rec { int cls; long long units; cudaEvent_t e0, e1; };
  std::vector<Rec> recs;
  stream_t st;
  bool on;
  explicit ProfScope(stream_t s) : st(s), on(prof_state().enabled != 0) {}
  stream_t cur;
  void begin(int cls, long long units) { begin_on(cls, units, st); }
  void begin_on(int cls, long long units, stream_t s) {
    if (!on) return;
    Rec r; r.cls = cls; r.units = units;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cur = s;
    cudaEventRecord(r.e0, s);
    recs.push_back(r);
  }
  void end() { if (on) cudaEventRecord(recs.back().e1, cur); }
  void flush() {  // call after the stream has been synchronised
    for (auto& r : recs) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, r.e0, r.e1);
      prof_state().ms[r.cls] += ms; prof_state().launches[r.cls] += 1; prof_state().units[r.cls] += r.units;
      cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    recs.clear();
  }
};
#endif
#define ILQC_PROF(cls, units, stmt) do { prof.begin(cls, units); stmt; prof.end(); } while (0)
#define ILQC_PROF_ON(stream, cls, units, stmt) do { prof.begin_on(cls, units, stream); stmt; prof.end(); } while (0)

struct Dims { int nx, nu, lin_stride, gain_stride, hess_stride; };
inline int get_dims(const ilqc_model_desc* m, Dims* d) {
  if (!m) return ILQC_E_ARG;
  if (m->integrator < ILQC_INT_EULER || m->integrator > ILQC_INT_RK4) return ILQC_E_ARG;
  if (m->integrator == ILQC_INT_RK4 && m->model == ILQC_MODEL_PENDULUM) return ILQC_E_UNSUPPORTED;  // Pendulum is Euler only (pendulum.py:68-71)
  ILQC_FOR_MODEL(effective_model(*m), LM::dims(&d->nx, &d->nu, &d->lin_stride, &d->gain_stride, &d->hess_stride));
  return 0;
}

// ---- bookkeeping kernels ------------------------------------------------------------------------------
namespace {

struct LsParams {
  int algo_type, approx, step_mode, L, line_search, H, nu;
  double rho;
  // variable candidate counts (asynchronous scheduler): candidates of active problem a occupy the slots
  // [slot_off[a], slot_off[a] + slot_cnt[a]); NULL: uniform L, slots a * L + c
  const int32_t* slot_cnt = nullptr;
  const int32_t* slot_off = nullptr;
};
// per-problem scheduler state of the asynchronous driver (all NULL in the lock-step driver)
enum { PHASE_LIN = 0, PHASE_LS = 1, PHASE_DONE = 2 };
struct AsyncState {
  int32_t *phase, *iter, *ls_trips, *n_done;
  double *cost, *gn, *step_rep;
};

// min_step's line-search set-up (algos_steps.py:333-367): first trial step of every running problem
ILQC_GLOBAL ls_init_kernel(LsParams P, int B, const int32_t* status, const double* stepsize, const double* cnorm,
                           double* s_try, int32_t* flags, int32_t* trips) {
  const int b = ILQC_TID;
  if (b >= B) return;
  int act = status[b] == ILQC_RUNNING;
  if (act) {
    double s = stepsize[b];
    if (P.line_search) {
      if (P.step_mode == ILQC_STEP_DIR) s = 1.0;
      else if (P.step_mode == ILQC_STEP_REG) s = (8.0 * s) / cnorm[b];
      else if (P.step_mode == ILQC_STEP_REGVAR) s = s * 8.0;
      else s = ::fmin(4.0 * s, 1.0);
    }
    s_try[b] = s;
  }
  trips[b] = 0;  // (stopped problems report 0 trips for the iterations they no longer run)
  flags[b] = act;
}

// ordered compaction of the flagged problems into list[0..count)
#if defined(ILQC_HOSTSIM)
ILQC_GLOBAL compact_kernel(int B, const int32_t* flags, int32_t* list, int32_t* count) {
  int n = 0;
  for (int b = 0; b < B; ++b)
    if (flags[b]) list[n++] = b;
  *count = n;
}
#else
// one block of 32 warps; warp w owns the contiguous range [w*per, (w+1)*per) and walks it 32 flags at a time
// (coalesced loads, ballot + popc), so the list stays sorted.  pred(b) decides membership.
template <class Pred>
__device__ __forceinline__ void compact_block(int B, Pred pred, int32_t* list, int32_t* count) {
  __shared__ int wsum[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int per = (((B + 31) / 32) + 31) & ~31;
  const int lo = wid * per, hi = min(B, lo + per);
  int n = 0;
  for (int base = lo; base < hi; base += 32) {
    const int b = base + lane;
    n += __popc(__ballot_sync(0xffffffffu, b < hi && pred(b)));
  }
  if (lane == 0) wsum[wid] = n;
  __syncthreads();
  int pos = 0, total = 0;
  for (int w = 0; w < 32; ++w) {
    const int v = wsum[w];
    if (w < wid) pos += v;
    total += v;
  }
  for (int base = lo; base < hi; base += 32) {
    const int b = base + lane;
    const bool f = b < hi && pred(b);
    const unsigned mask = __ballot_sync(0xffffffffu, f);
    if (f) list[pos + __popc(mask & ((1u << lane) - 1u))] = b;
    pos += __popc(mask);
  }
  if (threadIdx.x == 0) *count = total;
}
__global__ void __launch_bounds__(1024) compact_kernel(int B, const int32_t* flags, int32_t* list, int32_t* count) {
  compact_block(B, [=](int b) { return flags[b] != 0; }, list, count);
}
#endif

// candidates c < L of every still-searching problem: s_c = s_try * rho^c by repeated multiplication
// (algos_steps.py:420 `stepsize *= decrease_fac`), reg_ctrl = 1/s_c for the regularised oracles (:164, :244)
ILQC_GLOBAL make_slots_kernel(LsParams P, int n_active, const int32_t* list, const double* s_try, int32_t* prob,
                              int32_t* gslot, double* step_true, double* step_roll, double* reg) {
  const int a = ILQC_TID;
  if (a >= n_active) return;
  const int b = list[a];
  double s = s_try[b];
  const bool regmode = (P.step_mode == ILQC_STEP_REG || P.step_mode == ILQC_STEP_REGVAR);
  const int L = P.slot_cnt ? P.slot_cnt[a] : P.L, off = P.slot_off ? P.slot_off[a] : a * P.L;
  for (int c = 0; c < L; ++c) {
    const int slot = off + c;
    prob[slot] = b;
    step_true[slot] = s;
    if (P.algo_type == ILQC_ALGO_GD) {
      step_roll[slot] = -s;  // diff = -stepsize * grad (algos_steps.py:40)
      gslot[slot] = b;
    } else if (regmode) {
      step_roll[slot] = 1.0;  // roll_out_*(…, stepsize=1.0) with gains of reg = 1/s
      reg[slot] = 1.0 / s;
      gslot[slot] = slot;
    } else {
      step_roll[slot] = s;  // gains of the single backward pass, scaled offsets
      gslot[slot] = b;
    }
    s *= P.rho;
  }
}

// backtracking_line_search's accept / stop tests (algos_steps.py:411-434) for the L candidates of every
// still-searching problem, in the order the sequential reference would try them.
ILQC_GLOBAL ls_select_kernel(LsParams P, int n_active, int B, int n_slots, const int32_t* list,
                             const double* step_true, const double* newval, const double* diffsq,
                             const double* jcst_slot, const double* dec_unit, const int32_t* feas_slot,
                             const int32_t* feas_prob, const double* cnorm, double* curr_val, double* stepsize,
                             double* s_try, int32_t* status, int32_t* flags, int32_t* trips, int32_t* accept,
                             int max_trips, AsyncState A) {
  const int a = ILQC_TID;
  if (a >= n_active) return;
  const int b = list[a];
  const bool regmode = (P.step_mode == ILQC_STEP_REG || P.step_mode == ILQC_STEP_REGVAR);
  const double curr = curr_val[b];
  int chosen = -1;
  bool chosen_feasible = false;
  int ntrip = trips[b];
  const int L = P.slot_cnt ? P.slot_cnt[a] : P.L, off = P.slot_off ? P.slot_off[a] : a * P.L;
  for (int c = 0; c < L && chosen < 0; ++c) {
    const int slot = off + c;
    const double s = step_true[slot];
    bool feasible = true;
    if (P.algo_type != ILQC_ALGO_GD) feasible = regmode ? (feas_slot[slot] != 0) : (feas_prob[b] != 0);
    ++ntrip;
    bool decrease = false, stop = false;
    if (feasible) {
      // dec_obj: regularised oracles return the model decrease jcst of their own backward pass; direction
      // oracles scale the unit-step value (stepsize*opt_step, :157 / :237 ; gd: stepsize*fixed_obj, :41)
      const double dec = regmode && P.algo_type != ILQC_ALGO_GD ? jcst_slot[slot] : s * dec_unit[b];
      const double dirn = s > 0.0 ? 1.0 : (s < 0.0 ? -1.0 : 0.0);
      decrease = dirn * (newval[slot] - (curr + dec)) < 0.0;
      stop = ::sqrt(diffsq[slot]) < 1e-32;
    } else {
      stop = ::fabs(s) < 1e-32;
    }
    if (!P.line_search) { decrease = true; }
    if (decrease || stop || ntrip >= max_trips) {
      chosen = c;
      chosen_feasible = feasible;
    }
  }
  trips[b] = ntrip;
  accept[a] = -1;
  if (chosen >= 0) {
    const int slot = off + chosen;
    if (chosen_feasible) {
      accept[a] = slot;  // cmd += diff[slot] is done by apply_kernel, coalesced over the active problems
      curr_val[b] = newval[slot];
      if (P.line_search) {
        const double s = step_true[slot];
        stepsize[b] = (P.step_mode == ILQC_STEP_REG && P.algo_type != ILQC_ALGO_GD) ? s * cnorm[b] : s;
      }
      if (A.phase) {  // asynchronous scheduler: this problem moves on to its re-expansion
        const int it = A.iter[b];
        if (A.ls_trips) A.ls_trips[(size_t)it * B + b] = ntrip;
        A.iter[b] = it + 1;
        A.phase[b] = PHASE_LIN;
      }
    } else {
      status[b] = ILQC_NO_DIRECTION;  // min_step returns None (algos_steps.py:379-380)
      curr_val[b] = NAN;
      if (P.line_search) {  // ... with the last trial step (rescaled in 'reg' mode, algos_steps.py:372-373)
        const double s = step_true[slot];
        stepsize[b] = (P.step_mode == ILQC_STEP_REG && P.algo_type != ILQC_ALGO_GD) ? s * cnorm[b] : s;
      }
      if (A.phase) {
        const int it = A.iter[b];
        if (A.ls_trips) A.ls_trips[(size_t)it * B + b] = ntrip;
        const size_t row = (size_t)(it + 1) * B + b;  // collect_info with costs=None: NaN cost once
        A.cost[row] = NAN; A.gn[row] = NAN; A.step_rep[row] = stepsize[b];
        A.n_done[b] = it + 1;
        A.phase[b] = PHASE_DONE;
      }
    }
    flags[b] = 0;
  } else {
    s_try[b] = step_true[off + L - 1] * P.rho;
    flags[b] = 1;
  }
}

// ordered compaction of the problems with phase[b] == value (asynchronous scheduler)
#if defined(ILQC_HOSTSIM)
ILQC_GLOBAL compact_eq_kernel(int B, const int32_t* phase, int value, int32_t* list, int32_t* count) {
  int n = 0;
  for (int b = 0; b < B; ++b)
    if (phase[b] == value) list[n++] = b;
  *count = n;
}
#else
__global__ void __launch_bounds__(1024) compact_eq_kernel(int B, const int32_t* phase, int value, int32_t* list, int32_t* count) {
  compact_block(B, [=](int b) { return phase[b] == value; }, list, count);
}
#endif

// asynchronous scheduler, one tick: ordered lists of the problems in their line search (list) and of those
// waiting for their re-expansion (ident), and for the former the number of candidates of this tick from the
// trips already spent in the current search (round policy of round_candidates: 4, then 3, then 16, then 128 at
// a time), their slot offsets and the total.  When the total would exceed the slot capacity every count is capped.
ILQC_HD int tick_candidates(int trips_so_far) { return trips_so_far < 4 ? 4 : (trips_so_far < 7 ? 3 : (trips_so_far < 23 ? 16 : 128)); }
#if defined(ILQC_HOSTSIM)
ILQC_GLOBAL plan_async_kernel(int B, const int32_t* phase, const int32_t* trips, int Lmax, int cap, int32_t* list,
                              int32_t* cnt, int32_t* off, int32_t* ident, int32_t* count) {
  int cls[4] = {0, 0, 0, 0};
  for (int b = 0; b < B; ++b)
    if (phase[b] == PHASE_LS) { const int d = tick_candidates(trips[b]); ++cls[d == 4 ? 0 : (d == 3 ? 1 : (d == 16 ? 2 : 3))]; }
  static const int caps[9] = {128, 64, 32, 16, 8, 4, 3, 2, 1};
  int Lc = 1;
  for (int i = 0; i < 9; ++i) {
    const int c = caps[i];
    const long long D = (long long)cls[0] * (4 < c ? 4 : c) + (long long)cls[1] * (3 < c ? 3 : c) +
                        (long long)cls[2] * (16 < c ? 16 : c) + (long long)cls[3] * (128 < c ? 128 : c);
    if (D <= cap) { Lc = c; break; }
  }
  int n_ls = 0, n_lin = 0, slots = 0;
  for (int b = 0; b < B; ++b) {
    if (phase[b] == PHASE_LS) {
      int L = tick_candidates(trips[b]);
      L = L < Lc ? L : Lc;
      L = L < Lmax ? L : Lmax;
      list[n_ls] = b; cnt[n_ls] = L; off[n_ls] = slots;
      slots += L; ++n_ls;
    } else if (phase[b] == PHASE_LIN) {
      ident[n_lin++] = b;
    }
  }
  count[0] = n_ls; count[1] = n_lin; count[2] = slots;
}
#else
__global__ void __launch_bounds__(1024) plan_async_kernel(int B, const int32_t* phase, const int32_t* trips, int Lmax, int cap,
                                                          int32_t* list, int32_t* cnt, int32_t* off, int32_t* ident,
                                                          int32_t* count) {
  __shared__ int w_ls[32], w_lin[32], w_slots[32], w_cls[32][4];
  __shared__ int s_Lc;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int per = (((B + 31) / 32) + 31) & ~31;
  const int lo = wid * per, hi = min(B, lo + per);
  // pass 1: members and candidate classes of this warp's range
  int n_ls = 0, n_lin = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  for (int base = lo; base < hi; base += 32) {
    const int b = base + lane;
    const int ph = b < hi ? phase[b] : -1;
    const bool ls = ph == PHASE_LS;
    const int d = ls ? tick_candidates(trips[b]) : 0;
    n_ls += __popc(__ballot_sync(0xffffffffu, ls));
    n_lin += __popc(__ballot_sync(0xffffffffu, ph == PHASE_LIN));
    c0 += __popc(__ballot_sync(0xffffffffu, d == 4));
    c1 += __popc(__ballot_sync(0xffffffffu, d == 3));
    c2 += __popc(__ballot_sync(0xffffffffu, d == 16));
    c3 += __popc(__ballot_sync(0xffffffffu, d == 128));
  }
  if (lane == 0) { w_ls[wid] = n_ls; w_lin[wid] = n_lin; w_cls[wid][0] = c0; w_cls[wid][1] = c1; w_cls[wid][2] = c2; w_cls[wid][3] = c3; }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long cls[4] = {0, 0, 0, 0};
    for (int w = 0; w < 32; ++w)
      for (int k = 0; k < 4; ++k) cls[k] += w_cls[w][k];
    const int caps[9] = {128, 64, 32, 16, 8, 4, 3, 2, 1};
    int Lc = 1;
    for (int i = 0; i < 9; ++i) {
      const int c = caps[i];
      const long long D = cls[0] * min(4, c) + cls[1] * min(3, c) + cls[2] * min(16, c) + cls[3] * min(128, c);
      if (D <= cap) { Lc = c; break; }
    }
    s_Lc = Lc;
  }
  __syncthreads();
  const int Lc = s_Lc;
  // pass 2: slots of this warp's range
  int slots = 0;
  for (int base = lo; base < hi; base += 32) {
    const int b = base + lane;
    int L = (b < hi && phase[b] == PHASE_LS) ? min(min(tick_candidates(trips[b]), Lc), Lmax) : 0;
    for (int o = 16; o > 0; o >>= 1) L += __shfl_xor_sync(0xffffffffu, L, o);
    slots += L;
  }
  if (lane == 0) w_slots[wid] = slots;
  __syncthreads();
  int pos_ls = 0, pos_lin = 0, pos_slot = 0, tot_ls = 0, tot_lin = 0, tot_slot = 0;
  for (int w = 0; w < 32; ++w) {
    if (w < wid) { pos_ls += w_ls[w]; pos_lin += w_lin[w]; pos_slot += w_slots[w]; }
    tot_ls += w_ls[w]; tot_lin += w_lin[w]; tot_slot += w_slots[w];
  }
  // pass 3: ordered lists, counts and offsets
  const unsigned lt = (1u << lane) - 1u;
  for (int base = lo; base < hi; base += 32) {
    const int b = base + lane;
    const int ph = b < hi ? phase[b] : -1;
    const bool ls = ph == PHASE_LS, lin = ph == PHASE_LIN;
    const int L = ls ? min(min(tick_candidates(trips[b]), Lc), Lmax) : 0;
    int incl = L;  // inclusive warp scan of L
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const unsigned m_ls = __ballot_sync(0xffffffffu, ls), m_lin = __ballot_sync(0xffffffffu, lin);
    if (ls) {
      const int a = pos_ls + __popc(m_ls & lt);
      list[a] = b; cnt[a] = L; off[a] = pos_slot + incl - L;
    }
    if (lin) ident[pos_lin + __popc(m_lin & lt)] = b;
    pos_ls += __popc(m_ls);
    pos_lin += __popc(m_lin);
    pos_slot += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (threadIdx.x == 0) { count[0] = tot_ls; count[1] = tot_lin; count[2] = tot_slot; }
}
#endif

// asynchronous scheduler: collect_info + check_cvg_status for the problems that were just re-expanded, each
// at its own iteration index, followed by the line-search set-up of its next iteration (ls_init_kernel)
ILQC_GLOBAL status_async_kernel(LsParams P, int n_list, const int32_t* list, int B, int n_iter, const double* curr_val,
                                const double* gnorm, const double* cnorm, const double* stepsize, double* s_try,
                                double* dec_unit, int32_t* trips, int32_t* status, AsyncState A, const double* cost0) {
  const int a = ILQC_TID;
  if (a >= n_list) return;
  const int b = list[a];
  const int it = A.iter[b];
  const size_t row = (size_t)it * B + b;
  const double c = curr_val[b];
  double srep = stepsize[b];
  if (P.algo_type != ILQC_ALGO_GD && P.step_mode == ILQC_STEP_REG && it > 0) srep = srep * A.gn[(size_t)(it - 1) * B + b];
  A.cost[row] = c;
  A.gn[row] = gnorm[b];
  A.step_rep[row] = srep;
  A.n_done[b] = it;
  int st = ILQC_RUNNING;
  const double c0 = cost0 ? cost0[b] : A.cost[b];
  if (c != c || c > c0 + 1e3) st = ILQC_DIVERGED;
  if (it > 0) {
    const double cp = A.cost[(size_t)(it - 1) * B + b];
    if (::fabs(c - cp) / ::fabs(cp) < 1e-24) st = ILQC_CONVERGED;
  }
  if (st == ILQC_RUNNING && srep < 1e-24) st = ILQC_STUCK;
  status[b] = st;
  if (st == ILQC_RUNNING && it < n_iter) {
    double s = stepsize[b];
    if (P.line_search) {
      if (P.step_mode == ILQC_STEP_DIR) s = 1.0;
      else if (P.step_mode == ILQC_STEP_REG) s = (8.0 * s) / cnorm[b];
      else if (P.step_mode == ILQC_STEP_REGVAR) s = s * 8.0;
      else s = ::fmin(4.0 * s, 1.0);
    }
    s_try[b] = s;
    trips[b] = 0;
    if (P.algo_type == ILQC_ALGO_GD) dec_unit[b] = -0.5 * (gnorm[b] * gnorm[b]);
    A.phase[b] = PHASE_LS;
  } else {
    A.phase[b] = PHASE_DONE;
  }
}

// next_cmd = cmd + diff_cmd (algos_steps.py:377) for the problems whose line search just ended: thread
// (a, e) with the active index a fastest, so both the candidate read and the cmd write are coalesced
ILQC_GLOBAL apply_kernel(int n_active, int n_elem, int B, int n_slots, const int32_t* list, const int32_t* accept,
                         const double* diff, double* cmd) {
  const int i = ILQC_TID;
  if (i >= n_active * n_elem) return;
  const int a = i % n_active, e = i / n_active;
  const int slot = accept[a];
  if (slot >= 0) cmd[(size_t)e * B + list[a]] += diff[(size_t)e * n_slots + slot];
}

// collect_info + check_cvg_status (run_min_algo.py:127-264) for metrics row `it`
ILQC_GLOBAL status_kernel(LsParams P, int B, int it, const double* curr_val, const double* gnorm,
                          const double* stepsize, double* cost, double* gn, double* step_rep, int32_t* status,
                          int32_t* n_done, const double* cost0) {
  const int b = ILQC_TID;
  if (b >= B) return;
  const size_t row = (size_t)it * B + b;
  const int st_prev = status[b];
  if (it > 0 && st_prev != ILQC_RUNNING) {
    // stopped in an earlier iteration: pad.  A problem whose line search found no direction in THIS
    // iteration records a NaN cost once (collect_info with costs=None, run_min_algo.py:154).
    const bool failed_now = st_prev == ILQC_NO_DIRECTION && n_done[b] == it - 1;
    cost[row] = NAN; gn[row] = NAN; step_rep[row] = failed_now ? stepsize[b] : NAN;
    if (failed_now) n_done[b] = it;
    return;
  }
  const double c = curr_val[b];
  double srep = stepsize[b];
  // the reported step of the 'reg' mode is rescaled by the previous gradient norm (run_min_algo.py:158-161)
  if (P.algo_type != ILQC_ALGO_GD && P.step_mode == ILQC_STEP_REG && it > 0) srep = srep * gn[(size_t)(it - 1) * B + b];
  cost[row] = c;
  gn[row] = gnorm[b];
  step_rep[row] = srep;
  n_done[b] = it;
  int st = ILQC_RUNNING;
  const double c0 = cost0 ? cost0[b] : cost[b];  // (restart: the original metrics['cost'][0], ilqc_algo_desc.pp_cost0)
  if (c != c || c > c0 + 1e3) st = ILQC_DIVERGED;
  if (it > 0) {
    const double cp = cost[(size_t)(it - 1) * B + b];
    if (::fabs(c - cp) / ::fabs(cp) < 1e-24) st = ILQC_CONVERGED;
  }
  if (st == ILQC_RUNNING && srep < 1e-24) st = ILQC_STUCK;
  status[b] = st;
}

ILQC_GLOBAL fill_kernel(int n, double* p, double v) {
  const int i = ILQC_TID;
  if (i < n) p[i] = v;
}
ILQC_GLOBAL fill_i32_kernel(int n, int32_t* p, int32_t v) {
  const int i = ILQC_TID;
  if (i < n) p[i] = v;
}
// gd: dec_unit = -||grad||^2/2 (algos_steps.py:38) ; direction oracles: dec_unit = jcst of the backward pass
ILQC_GLOBAL gd_dec_kernel(int B, const double* gnorm, double* dec_unit) {
  const int b = ILQC_TID;
  if (b < B) dec_unit[b] = -0.5 * (gnorm[b] * gnorm[b]);
}
// direction oracles escalate reg_ctrl 1e-6, 1e-5, ... while the backward pass is infeasible and reg<1e6
// (algos_steps.py:140-144, :220-224): flag the problems that still need another pass
ILQC_GLOBAL escalate_kernel(int B, const int32_t* status, const int32_t* feas, double* reg, int32_t* flags) {
  const int b = ILQC_TID;
  if (b >= B) return;
  int f = 0;
  if (status[b] == ILQC_RUNNING && !feas[b] && reg[b] < 1e6) {
    reg[b] = reg[b] > 0.0 ? 10.0 * reg[b] : 1e-6;
    f = 1;
  }
  flags[b] = f;
}
// [B][n] -> [n][B] and back (reference layout <-> engine layout)
ILQC_GLOBAL to_soa_kernel(const double* src, double* dst, int B, int n) {
  const int i = ILQC_TID;
  if (i >= B * n) return;
  const int b = i % B, e = i / B;
  dst[i] = src[(size_t)b * n + e];
}
ILQC_GLOBAL from_soa_kernel(const double* src, double* dst, int B, int n) {
  const int i = ILQC_TID;
  if (i >= B * n) return;
  const int e = i % n, b = i / n;
  dst[i] = src[(size_t)e * B + b];
}

}  // namespace

// ---- workspace ----------------------------------------------------------------------------------------
struct Workspace {
  double *lin, *traj, *hess, *gsq, *gains, *diff, *dir, *newval, *diffsq, *jcst, *step_true, *step_roll, *reg, *curr_val,
      *cnorm, *gnorm, *s_try, *dec_unit, *reg_prob;
  int32_t *prob, *gslot, *feas_slot, *feas_prob, *list, *flags, *trips, *count, *ident, *accept, *phase, *iter, *slot_cnt,
      *slot_off;
  size_t bytes;
};
inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
// Threads one wave of the register-heavy kernels keeps resident (256 per SM at 255 registers/thread).
// Line-search rounds with fewer searching problems than this evaluate several candidates per problem at
// once: the speculative candidates ride on SMs that would otherwise idle.
// 512 threads per SM = exactly one resident wave of the roll-out kernel (128 registers) and two of the
// backward kernel; measured best on the bicycle workload (profiles/r01_tuning_scheduler.txt).
#ifndef ILQC_WAVE_THREADS_PER_SM
#define ILQC_WAVE_THREADS_PER_SM 512
#endif
inline int wave_slots() {
#if defined(ILQC_HOSTSIM)
  return 4096;
#else
  static int cached_by_dev[64] = {};  // (SM count of the CURRENT device: engines on several GPUs share this code)
  int dev = 0;
  cudaGetDevice(&dev);
  int& cached = cached_by_dev[dev & 63];
  if (!cached) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cached = sms * ILQC_WAVE_THREADS_PER_SM;
    if (const char* e = getenv("ILQC_WAVE_SLOTS")) {  // tuning knob (tools/gpu_tune.sh)
      const int v = atoi(e);
      if (v > 0) cached = v;
    }
  }
  return cached;
#endif
}
// ILQC_TRACE_ROUNDS=1: one stderr line per line-search round (tuning aid)
inline double trace_ms() {  // host clock, ms since the first call
  static const auto t0 = std::chrono::steady_clock::now();
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
inline bool trace_rounds() {
  static int v = -1;
  if (v < 0) v = getenv("ILQC_TRACE_ROUNDS") ? 1 : 0;
  return v == 1;
}
// Candidates per problem in line-search round `round` (0-based) of an iteration with n_active searching
// problems, at most Lmax (the caller's n_candidates).  Every candidate up to the accepted one has to be
// evaluated anyway (the sequential search of the reference tries them in order), so speculation only costs
// what lies beyond the accepted trip.  Policy from the trip statistics of the bicycle workload (trips 3-5 for
// 92 % of the problems once the iterates settle, a 0.2 % tail of 30-110 trips) and the measured launch costs
// (tools/sim_round_policy.py, profiles/r01_round_policy.txt):
//   round 0: clamp(49152 / n_active, 4, 16) candidates within 1.75 waves: 4 for a full batch (most problems
//            need the first 3-4 steps 8s, 4s, 2s, s); small batches leave the GPU idle, their launches are
//            latency-bound (the cost of a launch is flat below ~20 k slots), so a wider first round is free and
//            saves whole rounds (simulated -28 % at 4096 problems, -50 % at <= 1024)
//   round 1: up to 3 within one wave                (trip 5-7)
//   round 2: up to 16 within one wave
//   later, or as soon as <= 148 problems are left: up to 128 within one wave (stragglers: finish the 30-110
//            halvings in one launch)
// round < 0 (asynchronous scheduler, no round structure): one wave, Lmax.
inline size_t round_capacity(int round) {
  const size_t wave = (size_t)wave_slots();
  return round == 0 ? wave + (3 * wave) / 4 : wave;
}
// `done` = candidates every still-searching problem has already tried in this search.
// Small-batch regime (the first round affords more than 4 candidates per problem, i.e. fewer than ~12 k searching
// problems: MPC windows, single problems): launches are latency-bound, a round costs the same whether it carries 1 k or
// 20 k slots, so rounds are what to save - the first round takes up to 24 candidates within 65536 slots and every later
// round four times what has been tried so far, within one wave (tools/sim_round_policy.py on the trip counts of
// warm-started MPC windows, profiles/r02_scheduling_analysis.txt: -12 to -23 % line-search time against the 4/3/16/128 rule).
inline int round_candidates(int round, int n_active, int Lmax, int done = 0, bool small_batch = false) {
  const int n = n_active > 0 ? n_active : 1;
  if (round >= 0 && small_batch) {
    int c;
    if (round == 0) {
      c = 65536 / n;
      c = c < 4 ? 4 : (c > 24 ? 24 : c);
    } else {
      c = 4 * done;
      c = c > 128 ? 128 : c;
    }
    const int cap = (int)(round_capacity(round) / (size_t)n);  // (the workspace holds slot_capacity() slots)
    c = c > cap ? cap : c;
    c = c < 1 ? 1 : c;
    return c < Lmax ? c : Lmax;
  }
  if (round >= 0) {
    int c;
    if (round == 0) {
      c = 49152 / n;
      c = c < 4 ? 4 : (c > 16 ? 16 : c);
    } else if (n <= 148 || round >= 3) {
      c = 128;
    } else {
      c = round == 1 ? 3 : 16;
    }
    if (c < Lmax) Lmax = c;
  }
  int L = (int)(round_capacity(round) / (size_t)n);
  return L < 1 ? 1 : (L > Lmax ? Lmax : L);
}
// small-batch regime of an iteration that starts its search with n_active problems (see round_candidates)
inline bool small_batch_regime(int n_active) { return n_active > 0 && 49152 / n_active > 4; }
// slots the workspace must hold: n_active * round_candidates(...) <= max(B, largest round capacity)
inline size_t slot_capacity(int B, int Lmax) {
  // (two waves: round 0 of the bulk = 1.75 waves + the deferred stragglers of the asynchronous scheduler)
  size_t cap = (size_t)B * Lmax, wave = 2 * (size_t)wave_slots();
  size_t s = cap < wave ? cap : wave;
  return s < (size_t)B ? (size_t)B : s;
}
// quad: second-order algorithms also keep the second-order tensor of the dynamics [H][hess_stride][B]
inline bool needs_hess(const ilqc_algo_desc* a) { return a->approx == ILQC_APPROX_QUAD && a->algo_type != ILQC_ALGO_GD; }
inline void carve(Workspace* w, char* base, const Dims& d, int H, int B, int L, bool quad, bool tpar = false) {
  size_t off = 0;
  auto takeD = [&](size_t n) { double* p = (double*)(base + off); off = align_up(off + n * sizeof(double)); return p; };
  auto takeI = [&](size_t n) { int32_t* p = (int32_t*)(base + off); off = align_up(off + n * sizeof(int32_t)); return p; };
  const size_t S = slot_capacity(B, L);
  w->lin = takeD((size_t)H * d.lin_stride * B);
  w->traj = takeD((size_t)(H + 1) * d.nx * B);
  w->hess = quad ? takeD((size_t)H * d.hess_stride * B) : nullptr;
  w->gsq = tpar ? takeD((size_t)H * B) : nullptr;  // scratch of the time-parallel expansion
  w->gains = takeD((size_t)H * d.gain_stride * S);
  w->diff = takeD((size_t)H * d.nu * S);
  w->dir = takeD((size_t)H * d.nu * B);
  w->newval = takeD(S); w->diffsq = takeD(S); w->jcst = takeD(S); w->step_true = takeD(S); w->step_roll = takeD(S);
  w->reg = takeD(S);
  w->curr_val = takeD(B); w->cnorm = takeD(B); w->gnorm = takeD(B); w->s_try = takeD(B); w->dec_unit = takeD(B);
  w->reg_prob = takeD(B);
  w->prob = takeI(S); w->gslot = takeI(S); w->feas_slot = takeI(S);
  w->feas_prob = takeI(B); w->list = takeI(B); w->flags = takeI(B); w->trips = takeI(B); w->ident = takeI(B); w->accept = takeI(B);
  w->phase = takeI(B); w->iter = takeI(B); w->slot_cnt = takeI(B); w->slot_off = takeI(B);
  w->count = takeI(64);
  w->bytes = off;
}

inline int solve_impl(const ilqc_model_desc* m, const ilqc_algo_desc* a, int B, const double* x0, double* cmd,
                      double* stepsize, int n_iter, double* cost, double* gnorm_out, double* step_rep,
                      int32_t* n_done, int32_t* status, int32_t* ls_trips, void* workspace, size_t workspace_bytes,
                      stream_t st) {
  Dims d;
  int rc = get_dims(m, &d);
  if (rc) return rc;
  if (!a || B <= 0 || n_iter < 0 || a->n_candidates < 1 || a->expansion < 0 || a->expansion > 1) return ILQC_E_ARG;
  const int H = m->horizon, L = a->n_candidates;
  Workspace w;
  carve(&w, (char*)workspace, d, H, B, L, needs_hess(a), a->expansion == 1);
  if (!workspace || workspace_bytes < w.bytes) return ILQC_E_WORKSPACE;

  ProfScope prof(st);
  LsParams P;
  P.algo_type = a->algo_type; P.approx = a->approx; P.step_mode = a->step_mode; P.L = L;
  P.line_search = a->line_search; P.H = H; P.nu = d.nu;
  if (a->algo_type == ILQC_ALGO_GD) P.step_mode = ILQC_STEP_REGVAR;  // min_step forces it (algos_steps.py:311)
  const bool regmode = (P.step_mode == ILQC_STEP_REG || P.step_mode == ILQC_STEP_REGVAR);
  P.rho = regmode ? 0.5 : 0.9;
  const int max_trips = a->max_ls_trips > 0 ? a->max_ls_trips : 2000;
  const bool quad = a->approx == ILQC_APPROX_QUAD && a->algo_type != ILQC_ALGO_GD;
  if (a->handle_bad_dir < ILQC_BAD_DIR_CHECK_TOTAL || a->handle_bad_dir > ILQC_BAD_DIR_LET_IT_BE) return ILQC_E_ARG;
  const int bwd_mode = (!quad ? ILQC_BWD_LINQUAD
                              : (a->algo_type == ILQC_ALGO_DDP ? ILQC_BWD_QUAD_DDP : ILQC_BWD_QUAD_NEWTON)) |
                       (a->handle_bad_dir << ILQC_BWD_BAD_DIR_SHIFT);
  const bool need_grad = a->compute_grad_norm || a->algo_type == ILQC_ALGO_GD;
  const ilqc_model_desc& md = *m;

#define ILQC_CHECK(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

  ILQC_LAUNCH((fill_i32_kernel), B, 256, st, B, status, (int32_t)ILQC_RUNNING);
  ILQC_LAUNCH((fill_i32_kernel), B, 256, st, B, n_done, (int32_t)0);
  ILQC_LAUNCH((fill_i32_kernel), B, 256, st, B, w.feas_prob, (int32_t)1);
  ILQC_LAUNCH((fill_kernel), B, 256, st, B, w.gnorm, (double)NAN);
  // (rows of iterations that never run - every problem stopped early - report 0 trips, like the asynchronous driver)
  if (ls_trips && n_iter > 0) ILQC_LAUNCH((fill_i32_kernel), n_iter * B, 256, st, n_iter * B, ls_trips, (int32_t)0);

  // iteration 0: objective of the initial controls (same arithmetic as every candidate evaluation), its
  // expansion and the metrics row 0 (run_min_algo.py:65-76)
  ILQC_PROF(PROF_FORWARD0, B, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::forward(0, md, B, B, nullptr, x0, cmd, nullptr, nullptr, nullptr, w.curr_val, nullptr, st))));
  if (a->expansion == 1) { ILQC_PROF(PROF_LINEARIZE, B, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::expand_parallel(md, B, B, nullptr, x0, cmd, w.lin, w.traj, w.gsq, w.cnorm, st)))); }
  else { ILQC_PROF(PROF_LINEARIZE, B, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::forward(1, md, B, B, nullptr, x0, cmd, w.lin, w.traj, nullptr, nullptr, w.cnorm, st)))); }
  if (quad) ILQC_PROF(PROF_LINEARIZE, 0, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::hess(md, B, B, nullptr, w.traj, cmd, w.hess, st))));
  if (need_grad) ILQC_PROF(PROF_ADJOINT, B, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::adjoint(md, B, B, nullptr, w.lin, w.dir, w.gnorm, st))));
  ILQC_LAUNCH((status_kernel), B, 256, st, P, B, 0, w.curr_val, w.gnorm, stepsize, cost, gnorm_out, step_rep, status,
              n_done, a->pp_cost0);

  for (int it = 1; it <= n_iter; ++it) {
    ILQC_LAUNCH((ls_init_kernel), B, 256, st, P, B, status, stepsize, w.cnorm, w.s_try, w.flags, w.trips);
    ILQC_LAUNCH((compact_kernel), 1, 1024, st, B, w.flags, w.list, w.count);
    int n_active = 0;
    ILQC_CHECK(dev_d2h_sync(&n_active, w.count, sizeof(int), st));
    if (n_active == 0) {
      for (int r = it; r <= n_iter; ++r)
        ILQC_LAUNCH((status_kernel), B, 256, st, P, B, r, w.curr_val, w.gnorm, stepsize, cost, gnorm_out, step_rep,
                    status, n_done, a->pp_cost0);
      break;
    }
    if (a->algo_type == ILQC_ALGO_GD) {
      ILQC_LAUNCH((gd_dec_kernel), B, 256, st, B, w.gnorm, w.dec_unit);
    } else if (!regmode) {
      // direction oracles: one backward pass per problem (reg_ctrl = 0, escalated while infeasible)
      ILQC_LAUNCH((fill_kernel), B, 256, st, B, w.reg_prob, 0.0);
      const int32_t* plist = w.list;
      int n_pass = n_active;
      for (int esc = 0; esc < 16 && n_pass > 0; ++esc) {
        // slot == problem for these gains: thread a serves problem plist[a] and writes slot plist[a]
        ILQC_PROF(PROF_BACKWARD, n_pass, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::backward(bwd_mode, md, B, n_pass, plist, plist, B, w.reg_prob, 0.0, w.lin,
                                                         w.traj, cmd, w.hess, w.gains, w.dec_unit, w.feas_prob, st))));
        // always feasible: linear-quadratic passes unless 'flag', every pass under 'let_it_be'
        if ((!quad && a->handle_bad_dir != ILQC_BAD_DIR_FLAG) || a->handle_bad_dir == ILQC_BAD_DIR_LET_IT_BE) break;
        ILQC_LAUNCH((escalate_kernel), B, 256, st, B, status, w.feas_prob, w.reg_prob, w.flags);
        ILQC_LAUNCH((compact_kernel), 1, 1024, st, B, w.flags, w.ident, w.count);
        ILQC_CHECK(dev_d2h_sync(&n_pass, w.count, sizeof(int), st));
        plist = w.ident;
      }
      if (a->algo_type == ILQC_ALGO_CLASSIC) {
        // diff_cmd = roll_out_lin(traj, gains) once (algos_steps.py:153), candidates scale it
        ILQC_LAUNCH((fill_kernel), B, 256, st, B, w.step_roll, 1.0);
        ILQC_PROF(PROF_ROLLOUT, n_active, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::rollout(ILQC_ROLL_LIN, md, B, n_active, w.list, w.list, B, w.list, B,
                                                        w.step_roll, x0, cmd, w.traj, w.lin, w.gains, nullptr, w.dir,
                                                        w.newval, w.diffsq, st))));
      }
      // restore the active flags consumed by the escalation loop
      ILQC_LAUNCH((ls_init_kernel), B, 256, st, P, B, status, stepsize, w.cnorm, w.s_try, w.flags, w.trips);
    }

    // ---- line search rounds ----
    const bool small_batch = small_batch_regime(n_active);
    for (int round = 0, tried = 0; n_active > 0; ++round, tried += P.L) {
      P.L = round_candidates(round, n_active, L, tried, small_batch);
      const int n_slots = n_active * P.L;
      if (trace_rounds()) fprintf(stderr, "ilqc round: t %.3f ms iter %d active %d candidates %d\n", trace_ms(), it, n_active, P.L);
      ILQC_LAUNCH((make_slots_kernel), n_active, 256, st, P, n_active, w.list, w.s_try, w.prob, w.gslot, w.step_true,
                  w.step_roll, w.reg);
      if (a->algo_type == ILQC_ALGO_GD) {
        ILQC_PROF(PROF_ROLLOUT, n_slots, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::rollout(ILQC_ROLL_GD, md, B, n_slots, w.prob, nullptr, 0, nullptr, n_slots, w.step_roll, x0,
                                                        cmd, w.traj, w.lin, nullptr, w.dir, w.diff, w.newval, w.diffsq,
                                                        st))));
      } else if (regmode) {
        ILQC_PROF(PROF_BACKWARD, n_slots, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::backward(bwd_mode | ((size_t)n_active * 10 >= (size_t)B * 9 ? ILQC_BWD_DENSE_HINT : 0), md, B, n_slots, w.prob, nullptr, n_slots, w.reg, 0.0, w.lin, w.traj, cmd,
                                                         w.hess, w.gains, w.jcst, w.feas_slot, st))));
        const int kind = a->algo_type == ILQC_ALGO_DDP ? ILQC_ROLL_EXACT : ILQC_ROLL_LIN;
        ILQC_PROF(PROF_ROLLOUT, n_slots, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::rollout(kind, md, B, n_slots, w.prob, nullptr, n_slots, nullptr, n_slots, w.step_roll, x0, cmd,
                                                        w.traj, w.lin, w.gains, nullptr, w.diff, w.newval, w.diffsq,
                                                        st))));
      } else {
        const int kind = a->algo_type == ILQC_ALGO_DDP ? ILQC_ROLL_EXACT : ILQC_ROLL_GD;
        ILQC_PROF(PROF_ROLLOUT, n_slots, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::rollout(kind, md, B, n_slots, w.prob, w.gslot, B, nullptr, n_slots, w.step_roll, x0, cmd,
                                                        w.traj, w.lin, w.gains, w.dir, w.diff, w.newval, w.diffsq, st))));
      }
      ILQC_LAUNCH((ls_select_kernel), n_active, 128, st, P, n_active, B, n_slots, w.list, w.step_true, w.newval, w.diffsq,
                  w.jcst, w.dec_unit, w.feas_slot, w.feas_prob, w.cnorm, w.curr_val, stepsize, w.s_try, status, w.flags,
                  w.trips, w.accept, max_trips, AsyncState{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr});
      ILQC_LAUNCH((apply_kernel), n_active * (H * d.nu), 256, st, n_active, H * d.nu, B, n_slots, w.list, w.accept, w.diff, cmd);
      ILQC_LAUNCH((compact_kernel), 1, 1024, st, B, w.flags, w.list, w.count);
      ILQC_CHECK(dev_d2h_sync(&n_active, w.count, sizeof(int), st));
    }
    if (ls_trips) ILQC_CHECK(dev_d2d(ls_trips + (size_t)(it - 1) * B, w.trips, sizeof(int32_t) * B, st));

    // re-expand at the accepted controls (algos_steps.py:376-378) and collect the metrics of this iteration
    if (a->expansion == 1) { ILQC_PROF(PROF_LINEARIZE, B, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::expand_parallel(md, B, B, nullptr, x0, cmd, w.lin, w.traj, w.gsq, w.cnorm, st)))); }
    else { ILQC_PROF(PROF_LINEARIZE, B, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::forward(1, md, B, B, nullptr, x0, cmd, w.lin, w.traj, nullptr, nullptr, w.cnorm, st)))); }
  if (quad) ILQC_PROF(PROF_LINEARIZE, 0, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::hess(md, B, B, nullptr, w.traj, cmd, w.hess, st))));
    if (need_grad) ILQC_PROF(PROF_ADJOINT, B, ILQC_FOR_MODEL(effective_model(md), ILQC_CHECK(LM::adjoint(md, B, B, nullptr, w.lin, w.dir, w.gnorm, st))));
    ILQC_LAUNCH((status_kernel), B, 256, st, P, B, it, w.curr_val, w.gnorm, stepsize, cost, gnorm_out, step_rep, status,
                n_done, a->pp_cost0);
  }
  ILQC_CHECK(dev_sync(st));
  prof.flush();
  return ILQC_LAUNCH_OK();
#undef ILQC_CHECK
}

}  // namespace ilqc

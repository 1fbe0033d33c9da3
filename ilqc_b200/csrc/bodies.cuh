// Per-slot bodies of the hot-path kernels (one CUDA thread = one problem or one (problem, candidate)).
// HBM layout everywhere: [time][component][slot] with the slot (batch) index innermost, so a warp reads
// 32 consecutive doubles per component (fully coalesced 256 B requests).
//
// Reference functions restated (file:line in /root/reference):
//   forward_body   DiffEnv.forward/step, diff_discrete_dyn, diff_cost      envs/forward.py:62-139, :188-321
//   backward_body  lin_quad_backward, quad_backward(_ddp/_newton)          envs/backward.py:6-150
//   rollout_body   roll_out_exact, roll_out_lin, gd direction + obj(...)   envs/rollout.py:8-88,
//                                                                          algorithms/algos_steps.py:22-43, :329-331
//   adjoint_body   torch.autograd.grad(sum(costs), cmd)                    algorithms/run_min_algo.py:156-157
#pragma once
#include "common.cuh"
#include "jet.cuh"
#include "models.cuh"
#include "riccati.cuh"

namespace ilqc {

// ----------------------------------------------------------------------------------------------------
// expansion of one step: fills the packed record `rec` (Layout<M>) for step t from (x_t, u_t) and
// returns x_{t+1}, the step cost and |grad_ctrl|^2+|grad_state|^2 (algos_steps.py:342-351)
// ----------------------------------------------------------------------------------------------------
template <class M, class Store>
ILQC_HD void expand_step(const ilqc_model_desc& m, const Weights& w, const double (&x)[M::NX],
                         const double (&u)[M::NU], int time_next, double (&xn)[M::NX], double* cost,
                         double* gsq, Store&& store) {
  using L = Layout<M>;
  constexpr int NX = M::NX, NU = M::NU;
  {
    // first-order jets over the model's active variables; the heading (M::kRotVar) is not one of them: its column of
    // the Jacobian is a rotation of the position increments (models.cuh)
    constexpr int R = M::kRotVar;
    constexpr int NVJ = R >= 0 ? M::NV - 1 : M::NV;
    using S = Jet<NVJ, 1>;
    auto red = [](int v) constexpr { return (v < 0 || v == R) ? -1 : ((R >= 0 && v > R) ? v - 1 : v); };
    S xs[NX], us[NU], xns[NX];
    static_for<NX>([&](auto I) { xs[I] = make_var<S>(x[I], red(M::sv(I))); });
    static_for<NU>([&](auto I) { us[I] = make_var<S>(u[I], red(M::uv(I))); });
    M::template step<S>(m, xs, us, xns);
#pragma unroll
    for (int i = 0; i < NX; ++i) xn[i] = xns[i].v;
    // N = A - I, B
    static_for<NX>([&](auto I) {
      constexpr int i = I;
      static_for<NX>([&](auto Jc) {
        constexpr int j = Jc;
        if constexpr (M::Nm(i, j)) {
          double v;
          if constexpr (R >= 0 && M::sv(j) == R) v = (i == 0) ? -(xn[1] - x[1]) : ((i == 1) ? (xn[0] - x[0]) : 0.0);
          else if constexpr (M::sv(j) >= 0) v = xns[i].g[red(M::sv(j))] - (i == j ? 1.0 : 0.0);
          else v = passive_N_of<M>(m, i, j);
          store(L::oN + M::Nm.slot(i, j), v);
        }
      });
      static_for<NU>([&](auto Kc) {
        constexpr int k = Kc;
        if constexpr (M::Bm(i, k)) {
          double v;
          if constexpr (M::uv(k) >= 0) v = xns[i].g[red(M::uv(k))];
          else v = passive_B_of<M>(m, i, k);
          store(L::oB + M::Bm.slot(i, k), v);
        }
      });
    });
  }
  double gs = 0.0;
  {
    double q[NU], Qd[NU];
    const double cc = M::ctrl_cost(m, u, q, Qd);
    double gc = 0.0;
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      store(L::oq + i, q[i]);
      store(L::oQ + i, Qd[i]);
      gc += q[i] * q[i];
    }
    double cv, cg[M::NVC], ch[sym_size(M::NVC)];
    M::cost_expansion(m, w, xn, time_next, &cv, cg, ch);
    static_for<NX>([&](auto I) {
      constexpr int i = I;
      if constexpr (M::pm(0, i)) {
        const double v = cg[M::cv(i)];
        store(L::op + M::pm.slot(0, i), v);
        gs += v * v;
      }
      static_for<NX>([&](auto Jc) {
        constexpr int j = Jc;
        if constexpr (j >= i && M::Pm(i, j)) store(L::oP + M::Pm.slot(i, j), ch[sym_idx(M::cv(i), M::cv(j), M::NVC)]);
      });
    });
    *cost = cv + cc;
    *gsq = gc + gs;
  }
}

// ----------------------------------------------------------------------------------------------------
// forward: ORD==0 roll-out + costs ; ORD==1 additionally the packed expansion
// ----------------------------------------------------------------------------------------------------
template <class M, int ORD>
ILQC_HD void forward_body(const ilqc_model_desc& m, int b, int B, const double* __restrict__ x0,
                          const double* __restrict__ cmd, double* __restrict__ lin, double* __restrict__ traj,
                          double* __restrict__ cost_t, double* __restrict__ total, double* __restrict__ cnorm) {
  constexpr int NX = M::NX, NU = M::NU;
  const int H = m.horizon;
  const Weights w = load_weights(m, b, B);
  const int t0 = load_t0(m, b);
  double x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    x[i] = x0[(size_t)i * B + b];
    if (traj) traj[(size_t)i * B + b] = x[i];
  }
  if (cost_t) cost_t[b] = 0.0;
  double tot = 0.0, gsum = 0.0;
  for (int t = 0; t < H; ++t) {
    double u[NU], xn[NX], c;
#pragma unroll
    for (int k = 0; k < NU; ++k) u[k] = cmd[((size_t)t * NU + k) * B + b];
    if constexpr (ORD == 0) {
      M::template step<double>(m, x, u, xn);
      double q[NU], Qd[NU];
      const double cs = M::template state_cost<double>(m, w, xn, t0 + t + 1);
      c = cs + M::ctrl_cost(m, u, q, Qd);
    } else {
      double gsq;
      double* rec = lin + (size_t)t * Layout<M>::stride * B + b;
      expand_step<M>(m, w, x, u, t0 + t + 1, xn, &c, &gsq, [&](int slot, double v) { rec[(size_t)slot * B] = v; });
      gsum += gsq;
    }
    tot = tot + c;
    if (cost_t) cost_t[(size_t)(t + 1) * B + b] = c;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      x[i] = xn[i];
      if (traj) traj[((size_t)(t + 1) * NX + i) * B + b] = xn[i];
    }
  }
  if (total) total[b] = tot;
  if (cnorm) cnorm[b] = ::sqrt(gsum);
}

// ----------------------------------------------------------------------------------------------------
// time-parallel expansion (ilqc_algo_desc.expansion = 1): one thread per (problem, time step) expands step t
// from the stored trajectory; gsq[t][b] = |grad_ctrl|^2 + |grad_state|^2 of the step (summed in time order by
// cnorm_body, like the sequential kernel does)
// ----------------------------------------------------------------------------------------------------
template <class M>
ILQC_HD void expand_body(const ilqc_model_desc& m, int b, int B, int t, const double* __restrict__ traj,
                         const double* __restrict__ cmd, double* __restrict__ lin, double* __restrict__ gsq) {
  constexpr int NX = M::NX, NU = M::NU;
  const Weights w = load_weights(m, b, B);
  const int t0 = load_t0(m, b);
  double x[NX], u[NU], xn[NX], c, g;
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = traj[((size_t)t * NX + i) * B + b];
#pragma unroll
  for (int k = 0; k < NU; ++k) u[k] = cmd[((size_t)t * NU + k) * B + b];
  double* rec = lin + (size_t)t * Layout<M>::stride * B + b;
  expand_step<M>(m, w, x, u, t0 + t + 1, xn, &c, &g, [&](int slot, double v) { rec[(size_t)slot * B] = v; });
  gsq[(size_t)t * B + b] = g;
}
ILQC_HD void cnorm_body(int b, int B, int H, const double* __restrict__ gsq, double* __restrict__ cnorm) {
  double s = 0.0;
  for (int t = 0; t < H; ++t) s += gsq[(size_t)t * B + b];
  cnorm[b] = ::sqrt(s);
}

// norm that scales the regularised line search, from a stored packed expansion (algos_steps.py:342-351):
// sqrt(sum_{t>=1} |grad_ctrl_t|^2 + |grad_state_t|^2), summed in time order like forward_body does
template <class M>
ILQC_HD void cnorm_lin_body(const ilqc_model_desc& m, int b, int B, const double* __restrict__ lin, double* __restrict__ cnorm) {
  using L = Layout<M>;
  double gsum = 0.0;
  for (int t = 0; t < m.horizon; ++t) {
    const double* rec = lin + (size_t)t * L::stride * B + b;
    double gc = 0.0, gs = 0.0;
#pragma unroll
    for (int i = 0; i < M::NU; ++i) { const double v = rec[(size_t)(L::oq + i) * B]; gc += v * v; }
#pragma unroll
    for (int i = 0; i < L::np; ++i) { const double v = rec[(size_t)(L::op + i) * B]; gs += v * v; }
    gsum += gc + gs;
  }
  cnorm[b] = ::sqrt(gsum);
}

// ----------------------------------------------------------------------------------------------------
// loads of the packed record into the Riccati operands.  `rd(slot)` returns entry `slot` of the record.
// ----------------------------------------------------------------------------------------------------
struct GlobalRec {  // record in HBM: [component][problem], this thread's problem
  const double* p;
  size_t B;
  ILQC_HD double operator()(int slot) const { return p[(size_t)slot * B]; }
};
template <class M, class BellT, class Rd>
ILQC_HD void load_dyn(BellT& bl, const Rd& rd) {
  using L = Layout<M>;
  static_for<M::NX>([&](auto I) {
    constexpr int i = I;
    static_for<M::NX>([&](auto Jc) {
      constexpr int j = Jc;
      if constexpr (M::Nm(i, j)) bl.N[i][j] = rd(L::oN + M::Nm.slot(i, j));
    });
    static_for<M::NU>([&](auto Kc) {
      constexpr int k = Kc;
      if constexpr (M::Bm(i, k)) bl.B[i][k] = rd(L::oB + M::Bm.slot(i, k));
    });
  });
}
template <class M, class Rd>
ILQC_HD void load_state_cost(double (&p)[M::NX], double (&P)[M::NX][M::NX], const Rd& rd) {
  using L = Layout<M>;
  static_for<M::NX>([&](auto I) {
    constexpr int i = I;
    if constexpr (M::pm(0, i)) p[i] = rd(L::op + M::pm.slot(0, i)); else p[i] = 0.0;
    static_for<M::NX>([&](auto Jc) {
      constexpr int j = Jc;
      if constexpr (j >= i) {
        if constexpr (M::Pm(i, j)) P[i][j] = rd(L::oP + M::Pm.slot(i, j)); else P[i][j] = 0.0;
      }
    });
  });
}

// Operand feed of the backward recursion.  Step t consumes chunk(t) = {N, B, q, Qd of record t} + {p, P of
// record t-1} (the state cost charged on x_t), Layout<M>::stride doubles.
//   DirectFeed: plain global loads at the point of use (test-only host build; models whose chunk does not
//               fit the shared-memory budget).
//   AsyncFeed : (CUDA) every thread copies ITS OWN next chunk into shared memory with cp.async while it
//               computes the current step, double-buffered: [2][stride][block] doubles, no barrier needed
//               because a thread only ever reads what it copied itself.  The recursion runs at the
//               255-register limit with 8 warps per SM, so nothing else can cover the HBM latency of the
//               ~50 dependent-free loads per step (ncu before: 3.6 of 6.5 stall cycles per issue were
//               long_scoreboard at 47 % of the HBM peak).
#ifndef ILQC_BWD_PREFETCH
#define ILQC_BWD_PREFETCH 2   // 0 off, 1 prefetch.global.L1, 2 prefetch.global.L2
#endif
template <class M>
struct DirectFeed {
  const double* lin_b;  // lin + b
  size_t B;
  int t_cur = 0;
  ILQC_HD void prefetch(int t) {
#if ILQC_BWD_PREFETCH && defined(__CUDA_ARCH__)
    // pull chunk(t) into L2 one step ahead (no registers, no shared memory): the gathered rounds spent 57 % of their stall
    // samples in long_scoreboard at the first uses of N, B and of P, p (ncu source page, round 2); backward class
    // 52.4 -> 50.1 ms per solve (profiles/r02_tuning_small_rounds.txt; .L1 and .L2 measured equal)
    using L = Layout<M>;
    const double* rt = lin_b + (size_t)t * L::stride * B;
#pragma unroll
    for (int c = 0; c < L::op; ++c) {
#if ILQC_BWD_PREFETCH == 1
      asm volatile("prefetch.global.L1 [%0];" ::"l"(rt + (size_t)c * B));
#else
      asm volatile("prefetch.global.L2 [%0];" ::"l"(rt + (size_t)c * B));
#endif
    }
    if (t > 0) {
      const double* rp = rt - (size_t)L::stride * B;
#pragma unroll
      for (int c = L::op; c < L::stride; ++c) {
#if ILQC_BWD_PREFETCH == 1
        asm volatile("prefetch.global.L1 [%0];" ::"l"(rp + (size_t)c * B));
#else
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (size_t)c * B));
#endif
      }
    }
#else
    (void)t;
#endif
  }
  ILQC_HD void acquire(int t, bool) { t_cur = t; }
  ILQC_HD GlobalRec dyn() const { return GlobalRec{lin_b + (size_t)t_cur * Layout<M>::stride * B, B}; }
  ILQC_HD GlobalRec cost() const { return GlobalRec{lin_b + (size_t)(t_cur - 1) * Layout<M>::stride * B, B}; }
};
#if defined(__CUDA_ARCH__)
struct SmemRec {
  const double* p;  // this thread's column of the staged chunk: [component][block]
  ILQC_D double operator()(int slot) const { return p[slot * ILQC_BLOCK]; }
};
template <class M>
struct AsyncFeed {
  using L = Layout<M>;
  const double* lin_b;
  size_t B;
  double* sm;  // + threadIdx.x
  int stage = 1;  // buffer holding the chunk being consumed
  int fill = 1;   // buffer the latest prefetch went to
  static ILQC_D void cp8(double* dst, const double* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
  }
  // issue the copies of chunk(t) into the other buffer
  ILQC_D void prefetch(int t) {
    fill ^= 1;
    double* dst = sm + (size_t)fill * L::stride * ILQC_BLOCK;
    const double* rt = lin_b + (size_t)t * L::stride * B;
#pragma unroll
    for (int c = 0; c < L::op; ++c) cp8(dst + c * ILQC_BLOCK, rt + (size_t)c * B);
    if (t > 0) {
      const double* rp = rt - (size_t)L::stride * B;
#pragma unroll
      for (int c = L::op; c < L::stride; ++c) cp8(dst + c * ILQC_BLOCK, rp + (size_t)c * B);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // make chunk(t) readable: it is the older of the (at most) two groups in flight
  ILQC_D void acquire(int, bool one_in_flight_behind) {
    if (one_in_flight_behind) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    stage ^= 1;
  }
  ILQC_D SmemRec dyn() const { return SmemRec{sm + (size_t)stage * L::stride * ILQC_BLOCK}; }
  ILQC_D SmemRec cost() const { return dyn(); }
};
#endif

// ----------------------------------------------------------------------------------------------------
// second-order tensor of the dynamics at (x_t, u_t) of the current iterate: packed Hessians over the NV jet
// variables of the first NXH rows of phi, layout [t][row * sym_size(NV) + entry][problem].  One thread per
// (problem, time step, group of Hessian entries): the steps are independent once the trajectory is known.  Restates the closures
// hess_dyn_state / hess_dyn_ctrl / hess_dyn_state_ctrl of forward.py:218-269 before their contraction with lambda.
// ----------------------------------------------------------------------------------------------------
// models may provide a structured second-order step (M::kStructuredHess, M::step_hess)
// (-DILQC_HESS_GENERIC: generic jets through the whole step for every model - the A/B and cross-check build)
template <class M, class = void> struct structured_hess : std::false_type {};
#if !defined(ILQC_HESS_GENERIC)
template <class M> struct structured_hess<M, std::enable_if_t<M::kStructuredHess>> : std::true_type {};
#endif
// per-thread scratch of the structured step (the RK4 accumulator): a column of a shared-memory tile in the CUDA kernel
// (element i of the thread at i * ILQC_BLOCK: conflict-free), a plain array in the host build
template <int N> struct HessAccLocal {
  double a[N > 0 ? N : 1];
  ILQC_HD double& operator[](int i) { return a[i]; }
};
struct HessAccShared {
  double* col;
  ILQC_HD double& operator[](int i) { return col[(size_t)i * ILQC_BLOCK]; }
};
template <class M, int NG, int G>
constexpr int hess_acc_doubles() {
  if constexpr (structured_hess<M>::value) {
    constexpr int NVJ = M::kRotVar >= 0 ? M::NV - 1 : M::NV;
    return M::template kHessAccDoubles<Jet<NVJ, 2, NG, G>>;
  } else {
    return 0;
  }
}
template <class M, int NG, int G, class ACC>
ILQC_HD void hess_body(const ilqc_model_desc& m, int b, int B, int t, const double* __restrict__ traj,
                       const double* __restrict__ cmd, double* __restrict__ hess, ACC& acc) {
  constexpr int NX = M::NX, NU = M::NU, NH = sym_size(M::NV);
  constexpr int R = M::kRotVar;                  // heading variable, handled in closed form (models.cuh), or -1
  constexpr int NVJ = R >= 0 ? M::NV - 1 : M::NV;  // jet variables actually carried
  using S = Jet<NVJ, 2, NG, G>;  // group G of NG groups of Hessian entries (jet.cuh)
  double x[NX], u[NU];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = traj[((size_t)t * NX + i) * B + b];
#pragma unroll
  for (int k = 0; k < NU; ++k) u[k] = cmd[((size_t)t * NU + k) * B + b];
  // jet index without the heading variable
  auto red = [](int v) constexpr { return (v < 0 || v == R) ? -1 : ((R >= 0 && v > R) ? v - 1 : v); };
  S xs[NX], us[NU], xns[NX];
  static_for<NX>([&](auto I) { xs[I] = make_var<S>(x[I], red(M::sv(I))); });
  static_for<NU>([&](auto I) { us[I] = make_var<S>(u[I], red(M::uv(I))); });
  if constexpr (structured_hess<M>::value) M::template step_hess<S>(m, xs, us, xns, acc);  // (models.cuh: bicycle model)
  else M::template step<S>(m, xs, us, xns);
  // entries among the carried variables only: the heading entries are closed forms of the first-order expansion and are
  // rebuilt by the consumer (backward_body)
  double* T = hess + (size_t)t * Layout<M>::hess_stride * B + b;
  static_for<M::NXH>([&](auto I) {
    constexpr int i = I;
    static_for<S::NH>([&](auto Kk) {
      constexpr int k = Kk;
      if constexpr (S::valid(k)) T[(size_t)(i * Layout<M>::hess_row + S::entry(k)) * B] = xns[i].h[k];
    });
  });
}

// ----------------------------------------------------------------------------------------------------
// backward pass (MODE = ILQC_BWD_*).  slot s serves problem b with regularisation reg_ctrl.
// gains record per step: [K row-major NU x NX | k NU], layout [t][gain_stride][n_slots]
// ----------------------------------------------------------------------------------------------------
// Feed of the stored dynamics-Hessian tensor into the second-order backward passes: step t contracts the
// Layout<M>::hess_stride doubles of (problem, t) with its costate.
//   DirectTensor: plain global loads at the point of use.  (A cp.async feed through shared memory - the thread's rows of
//   step t-1 copied while step t computes, [hess_stride][block] doubles - was measured SLOWER on a B200: 3.01 ms against
//   2.51 ms for the full first round; the 4 candidates of a problem copy from equal addresses and 8-byte cp.async
//   requests of equal addresses serialise in the LSU, profiles/r02_tuning_small_rounds.txt.)
struct DirectTensor {
  const double* hess_b;  // hess + b
  size_t B;
  int stride;
  const double* cur = nullptr;
  ILQC_HD void prefetch(int) {}   // (prefetch.global.L2 of the next step's 90 rows: measured slower, 47.5 -> 53.4 ms per solve)
  ILQC_HD void acquire(int t) { cur = hess_b + (size_t)t * stride * B; }
  ILQC_HD double operator()(int idx) const { return cur[(size_t)idx * B]; }
};
struct NoTensor {
  ILQC_HD void prefetch(int) {}
  ILQC_HD void acquire(int) {}
  ILQC_HD double operator()(int) const { return 0.0; }
};

template <class M, int MODE, bool TENSOR, class Feed, class TFeed>
ILQC_HD void backward_body(const ilqc_model_desc& m, int s, int n_slots, int b, int B, double reg_ctrl,
                           double reg_state, const double* __restrict__ lin, const double* __restrict__ traj,
                           const double* __restrict__ cmd, const double* __restrict__ hess, double* __restrict__ gains,
                           double* __restrict__ jcst_out, int32_t* __restrict__ feas_out, int bad_dir, Feed& feed, TFeed& tfeed) {
  constexpr int NX = M::NX, NU = M::NU;
  constexpr bool QUAD = (MODE != ILQC_BWD_LINQUAD);
  using L = Layout<M>;
  using Pat = PatOf<M, QUAD>;
  const int H = m.horizon;
  Bell<Pat> bl;
  double J[sym_size(NX)], j[NX];
  {
    double pH[NX], PH[NX][NX];
    load_state_cost<M>(pH, PH, GlobalRec{lin + (size_t)(H - 1) * L::stride * B + b, (size_t)B});
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      j[i] = pH[i];
#pragma unroll
      for (int c = i; c < NX; ++c) J[sym_idx(i, c, NX)] = PH[i][c] + (i == c ? reg_state : 0.0);
    }
  }
  // constant cost-Hessian entry outside the stored pattern (Car time_penalty='dref_squared', car.py:213-216)
  double p_cross = 0.0;
  if constexpr (M::kPassivePRow >= 0) {
    p_cross = M::passive_P(m, load_weights(m, b, B));
    J[sym_idx(M::kPassivePRow, M::kPassivePCol, NX)] += p_cross;
  }
  bool any_pos = false;  // handle_bad_dir='flag': some step had rest > 0 (backward.py:220-223)
  double lam[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) lam[i] = j[i];
  double jcst = 0.0;
  feed.prefetch(H - 1);
  if constexpr (TENSOR) tfeed.prefetch(H - 1);
  for (int t = H - 1; t >= 0; --t) {
    if (t > 0) feed.prefetch(t - 1);
    feed.acquire(t, t > 0);
    {
      const auto rec = feed.dyn();
      load_dyn<M>(bl, rec);
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        bl.q[u] = rec(L::oq + u);
#pragma unroll
        for (int v = u; v < NU; ++v) bl.Q[u][v] = 0.0;
        bl.Q[u][u] = rec(L::oQ + u) + reg_ctrl;
      }
    }
    double hxx[QUAD ? sym_size(M::NV) : 1];  // state block of the dynamics Hessian, added to P in phase 2
    if constexpr (QUAD) {
      // Hessian of lam^T phi at (x_t, u_t) (forward.py:218-269)
      if constexpr (TENSOR) {
        // contraction of the stored second-order tensor of the dynamics (hess_body): it depends on the
        // iterate only, not on lambda or the regularisation, so every candidate of every round shares it
        tfeed.acquire(t);
        constexpr int R = M::kRotVar;
        static_for<sym_size(M::NV)>([&](auto E) {
          constexpr int e = E;
          constexpr int p = SymTable<M::NV>().i[e], q = SymTable<M::NV>().j[e];
          if constexpr (p != R && q != R) {
            constexpr int er = sym_idx(L::hess_red(p), L::hess_red(q), L::hess_nvj);
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < M::NXH; ++i) acc += lam[i] * tfeed(i * L::hess_row + er);
            hxx[e] = acc;
          }
        });
        if (t > 0) tfeed.prefetch(t - 1);   // (the buffer is free: every entry of step t has been read)
        if constexpr (R >= 0) {
          // entries with the heading phi (models.cuh kRotVar): (x' - x, y' - y) = R(phi) w with w independent of phi, so
          // d/dphi (x' - x) = -(y' - y), d/dphi (y' - y) = x' - x and every mixed derivative is the same rotation of the
          // first-order expansion; the other rows do not depend on phi
          const double dx = traj[((size_t)(t + 1) * NX + 0) * B + b] - traj[((size_t)t * NX + 0) * B + b];
          const double dy = traj[((size_t)(t + 1) * NX + 1) * B + b] - traj[((size_t)t * NX + 1) * B + b];
          static_for<M::NV>([&](auto Wv) {
            constexpr int w = Wv;
            constexpr int e = sym_idx(w < R ? w : R, w < R ? R : w, M::NV);
            if constexpr (w == R) {
              hxx[e] = lam[0] * (-dx) + lam[1] * (-dy);
            } else {
              double gx = 0.0, gy = 0.0;   // d x' / d w, d y' / d w from the packed first-order expansion
              static_for<NX>([&](auto C) {
                constexpr int c = C;
                if constexpr (M::sv(c) == w) {
                  if constexpr (M::Nm(0, c)) gx = bl.N[0][c];
                  if constexpr (M::Nm(1, c)) gy = bl.N[1][c];
                }
              });
              static_for<NU>([&](auto Kc) {
                constexpr int k = Kc;
                if constexpr (M::uv(k) == w) {
                  if constexpr (M::Bm(0, k)) gx = bl.B[0][k];
                  if constexpr (M::Bm(1, k)) gy = bl.B[1][k];
                }
              });
              hxx[e] = lam[0] * (-gy) + lam[1] * gx;
            }
          });
        }
      } else {
        // ... recomputed here by second-order jets (stand-alone ilqc_backward)
        using S = Jet<M::NV, 2>;
        double x[NX], u[NU];
#pragma unroll
        for (int i = 0; i < NX; ++i) x[i] = traj[((size_t)t * NX + i) * B + b];
#pragma unroll
        for (int k = 0; k < NU; ++k) u[k] = cmd[((size_t)t * NU + k) * B + b];
        S xs[NX], us[NU], xns[NX];
        static_for<NX>([&](auto I) { xs[I] = make_var<S>(x[I], M::sv(I)); });
        static_for<NU>([&](auto I) { us[I] = make_var<S>(u[I], M::uv(I)); });
        M::template step<S>(m, xs, us, xns);
#pragma unroll
        for (int e = 0; e < sym_size(M::NV); ++e) {
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < M::NXH; ++i) acc += lam[i] * xns[i].h[e];
          hxx[e] = acc;
        }
      }
      static_for<NX>([&](auto I) {
        constexpr int i = I;
        if constexpr (M::sv(i) >= 0) {
          static_for<NU>([&](auto Kc) {
            constexpr int k = Kc;
            if constexpr (M::uv(k) >= 0) bl.R[i][k] = hxx[sym_idx(M::sv(i), M::uv(k), M::NV)];
          });
        }
      });
      static_for<NU>([&](auto U) {
        constexpr int u1 = U;
        if constexpr (M::uv(u1) >= 0) {
          static_for<NU>([&](auto V) {
            constexpr int v1 = V;
            if constexpr (v1 >= u1 && M::uv(v1) >= 0) bl.Q[u1][v1] += hxx[sym_idx(M::uv(u1), M::uv(v1), M::NV)];
          });
        }
      });
    }
    // newton-mode costate: lam <- p + A^T lam = p + lam + N^T lam (backward.py:122-123); the N part now,
    // p once the state cost has been loaded
    double lam_next[NX];
    if constexpr (MODE == ILQC_BWD_QUAD_NEWTON) {
      static_for<NX>([&](auto I) {
        constexpr int i = I;
        double acc = lam[i];
        static_for<NX>([&](auto Rr) {
          constexpr int r = Rr;
          if constexpr (M::Nm(r, i)) acc += bl.N[r][i] * lam[r];
        });
        lam_next[i] = acc;
      });
    }
    bl.phase1(J, j);
    // state cost on x_t (costs[t] of the reference; zero at t = 0), loaded only now to keep N, B, W and the
    // cost operands from being live at the same time
    if (t > 0) {
      load_state_cost<M>(bl.p, bl.P, feed.cost());
    } else {
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        bl.p[i] = 0.0;
#pragma unroll
        for (int c = i; c < NX; ++c) bl.P[i][c] = 0.0;
      }
    }
    if constexpr (MODE == ILQC_BWD_QUAD_NEWTON) {
#pragma unroll
      for (int i = 0; i < NX; ++i) lam_next[i] += bl.p[i];
    }
    // (P + state block of the dynamics Hessian) + reg_state on the diagonal, applied where phase 2 uses P (riccati.cuh)
    if constexpr (!QUAD) {
      bl.reg_diag = reg_state;
    } else {
      // P gets the state block of the dynamics Hessian and reg_state only for t>0 (backward.py:102-108)
      bl.reg_diag = t > 0 ? reg_state : 0.0;
      static_for<NX>([&](auto I) {
        constexpr int i = I;
        if constexpr (M::sv(i) >= 0) {
          static_for<NX>([&](auto Jc) {
            constexpr int c = Jc;
            if constexpr (c >= i && M::sv(c) >= 0) bl.Padd[i][c] = t > 0 ? hxx[sym_idx(M::sv(i), M::sv(c), M::NV)] : 0.0;
          });
        }
      });
    }
    const double rest = bl.phase2(J, j);
    if constexpr (M::kPassivePRow >= 0) {
      if (t > 0) J[sym_idx(M::kPassivePRow, M::kPassivePCol, NX)] += p_cross;
    }
    any_pos |= rest > 0.0;
    jcst = jcst + rest;
    double* g = gains + (size_t)t * L::gain_stride * n_slots + s;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
#pragma unroll
      for (int c = 0; c < NX; ++c) g[(size_t)(u * NX + c) * n_slots] = bl.K[u][c];
      g[(size_t)(NU * NX + u) * n_slots] = bl.k[u];
    }
    if constexpr (MODE == ILQC_BWD_QUAD_NEWTON) {
#pragma unroll
      for (int i = 0; i < NX; ++i) lam[i] = lam_next[i];
    } else if constexpr (MODE == ILQC_BWD_QUAD_DDP) {
#pragma unroll
      for (int i = 0; i < NX; ++i) lam[i] = j[i];
    }
  }
  jcst_out[s] = jcst;
  // handle_bad_dir == 'check_total_value' (run_min_algo.py:21): linquad is always feasible,
  // the quad variants require a negative model decrease (backward.py:130-131)
  // 'flag': infeasible as soon as one step is non-convex; 'let_it_be': always feasible
  if (feas_out) {
    int feas = QUAD ? (jcst < 0.0 ? 1 : 0) : 1;
    if (bad_dir == ILQC_BAD_DIR_FLAG) feas = any_pos ? 0 : 1;
    else if (bad_dir == ILQC_BAD_DIR_LET_IT_BE) feas = 1;
    feas_out[s] = feas;
  }
}

// ----------------------------------------------------------------------------------------------------
// roll-outs fused with the objective of the candidate.
//   EXACT: v_t = K_t (x_t - xbar_t) + step*k_t along the true dynamics              (rollout.py:50-88)
//   LIN  : v_t = K_t dx_t + step*k_t ; dx_{t+1} = A_t dx_t + B_t v_t                (rollout.py:8-47)
//   DIR  : v_t = step * dir_t   (classic 'dir' oracles and gd, algos_steps.py:38-41, :153-160)
// In every case newval = sum of costs along the true dynamics under cmd+v (algos_steps.py:329-331).
// ----------------------------------------------------------------------------------------------------
template <class M, int KIND>
ILQC_HD void rollout_body(const ilqc_model_desc& m, int s, int n_slots, int b, int B, int gs, int n_gslots,
                          double step, const double* __restrict__ x0, const double* __restrict__ cmd,
                          const double* __restrict__ traj, const double* __restrict__ lin,
                          const double* __restrict__ gains, const double* __restrict__ dir,
                          double* __restrict__ diff, double* __restrict__ newval, double* __restrict__ diffsq) {
  constexpr int NX = M::NX, NU = M::NU;
  using L = Layout<M>;
  const int H = m.horizon;
  const Weights w = load_weights(m, b, B);
  const int t0 = load_t0(m, b);
  double x[NX], dx[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    x[i] = x0[(size_t)i * B + b];
    dx[i] = 0.0;
  }
  double tot = 0.0, nsq = 0.0;
  for (int t = 0; t < H; ++t) {
    double v[NU], u[NU];
    if constexpr (KIND == ILQC_ROLL_GD) {
#pragma unroll
      for (int k = 0; k < NU; ++k) v[k] = step * dir[((size_t)t * NU + k) * B + b];
    } else {
      const double* g = gains + (size_t)t * L::gain_stride * n_gslots + gs;
#pragma unroll
      for (int k = 0; k < NU; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < NX; ++c) acc += g[(size_t)(k * NX + c) * n_gslots] * dx[c];
        v[k] = acc + step * g[(size_t)(NU * NX + k) * n_gslots];
      }
    }
#pragma unroll
    for (int k = 0; k < NU; ++k) {
      diff[((size_t)t * NU + k) * n_slots + s] = v[k];
      nsq += v[k] * v[k];
      u[k] = cmd[((size_t)t * NU + k) * B + b] + v[k];
    }
    if constexpr (KIND == ILQC_ROLL_LIN) {
      // dx <- A dx + B v = dx + N dx + B v on the packed expansion
      const double* rec = lin + (size_t)t * L::stride * B + b;
      double dn[NX];
      static_for<NX>([&](auto I) {
        constexpr int i = I;
        double acc = dx[i];
        static_for<NX>([&](auto Jc) {
          constexpr int c = Jc;
          if constexpr (M::Nm(i, c)) acc += rec[(size_t)(L::oN + M::Nm.slot(i, c)) * B] * dx[c];
        });
        static_for<NU>([&](auto Kc) {
          constexpr int k = Kc;
          if constexpr (M::Bm(i, k)) acc += rec[(size_t)(L::oB + M::Bm.slot(i, k)) * B] * v[k];
        });
        dn[i] = acc;
      });
#pragma unroll
      for (int i = 0; i < NX; ++i) dx[i] = dn[i];
    }
    double xn[NX], q[NU], Qd[NU];
    M::template step<double>(m, x, u, xn);
    // reference state that closes the step (dx = x_{t+1} - xbar_{t+1}), requested before the cost is evaluated so that its
    // HBM latency hides behind the ~300 instructions of the cost (ncu source page, round 2: the subtraction below was the
    // instruction with the most stall samples of the kernel, all long_scoreboard)
    double xr[KIND == ILQC_ROLL_EXACT ? NX : 1];
    if constexpr (KIND == ILQC_ROLL_EXACT) {
#pragma unroll
      for (int i = 0; i < NX; ++i) xr[i] = traj[((size_t)(t + 1) * NX + i) * B + b];
    }
    const double cs = M::template state_cost<double>(m, w, xn, t0 + t + 1);
    const double c = cs + M::ctrl_cost(m, u, q, Qd);
    tot = tot + c;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      x[i] = xn[i];
      if constexpr (KIND == ILQC_ROLL_EXACT) dx[i] = xn[i] - xr[i];
    }
  }
  newval[s] = tot;
  diffsq[s] = nsq;
}

// ----------------------------------------------------------------------------------------------------
// gradient of the total cost wrt the controls by the adjoint recursion on the packed expansion
// ----------------------------------------------------------------------------------------------------
template <class M>
ILQC_HD void adjoint_body(const ilqc_model_desc& m, int b, int B, const double* __restrict__ lin,
                          double* __restrict__ grad, double* __restrict__ gnorm) {
  constexpr int NX = M::NX, NU = M::NU;
  using L = Layout<M>;
  const int H = m.horizon;
  double carry[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) carry[i] = 0.0;
  double nsq = 0.0;
  for (int t = H - 1; t >= 0; --t) {
    const double* rec = lin + (size_t)t * L::stride * B + b;
    double lam[NX];
    static_for<NX>([&](auto I) {
      constexpr int i = I;
      lam[i] = carry[i];
      if constexpr (M::pm(0, i)) lam[i] += rec[(size_t)(L::op + M::pm.slot(0, i)) * B];
    });
    static_for<NU>([&](auto Kc) {
      constexpr int k = Kc;
      double acc = rec[(size_t)(L::oq + k) * B];
      static_for<NX>([&](auto I) {
        constexpr int i = I;
        if constexpr (M::Bm(i, k)) acc += rec[(size_t)(L::oB + M::Bm.slot(i, k)) * B] * lam[i];
      });
      if (grad) grad[((size_t)t * NU + k) * B + b] = acc;
      nsq += acc * acc;
    });
    static_for<NX>([&](auto I) {
      constexpr int i = I;
      double acc = lam[i];
      static_for<NX>([&](auto Rr) {
        constexpr int r = Rr;
        if constexpr (M::Nm(r, i)) acc += rec[(size_t)(L::oN + M::Nm.slot(r, i)) * B] * lam[r];
      });
      carry[i] = acc;
    });
  }
  gnorm[b] = ::sqrt(nsq);
}

}  // namespace ilqc

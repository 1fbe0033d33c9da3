// Kernels + launchers of ONE model (ILQC_MODEL_INST), included by model_<k>.cu (product, nvcc) and by
// the test-only host simulator.  Each kernel is a thread-per-slot wrapper around a body of bodies.cuh.
#include "launch.cuh"
#include "bodies.cuh"
#include "coop.cuh"

#ifndef ILQC_MODEL_INST
#error "define ILQC_MODEL_INST"
#endif

namespace ilqc {

namespace {
using MI = Model<ILQC_MODEL_INST>;

template <class M, int ORD>
ILQC_GLOBAL_LB(ILQC_LB_FWD1) forward_kernel(ilqc_model_desc m, int B, int n_threads, const int32_t* prob, const double* x0,
                                            const double* cmd, double* lin, double* traj, double* cost_t, double* total,
                                            double* cnorm) {
  const int a = ILQC_TID;
  if (a >= n_threads) return;
  const int b = prob ? prob[a] : a;
  forward_body<M, ORD>(m, b, B, x0, cmd, lin, traj, cost_t, total, cnorm);
}

// shared-memory budget of the double-buffered operand feed (bodies.cuh AsyncFeed): ILQC_LB_BWD resident blocks
// of [2][stride][ILQC_BLOCK] doubles must fit in the SM's 227 KB (1 KB per block is reserved by the runtime)
template <class M>
constexpr bool kBwdAsyncFeed =
#if defined(ILQC_HOSTSIM)
    false;
#else
    (size_t)ILQC_LB_BWD * (2 * Layout<M>::stride * ILQC_BLOCK * sizeof(double) + 1024) <= 227 * 1024;
#endif
template <class M>
constexpr size_t kBwdSmemBytes = kBwdAsyncFeed<M> ? 2 * Layout<M>::stride * ILQC_BLOCK * sizeof(double) : 0;

// ASYNC: operands staged through shared memory by cp.async (bodies.cuh AsyncFeed).  Measured on a B200
// (profiles/r01_tuning_backward_feed.txt): 16-26 % faster than direct loads when the slots map to (nearly)
// consecutive problems with few candidates each, i.e. the first line-search round of an iteration, and
// slower for gathered / many-candidates-per-problem rounds (8-byte cp.async of equal or scattered addresses
// serialises in the LSU), so the launcher picks it only for dense rounds of the linear-quadratic mode.
template <class M, int MODE, bool ASYNC, bool TENSOR>
ILQC_GLOBAL_LB(ILQC_LB_BWD) backward_kernel(ilqc_model_desc m, int B, int n_threads, const int32_t* prob, const int32_t* oslot,
                            int n_oslots, const double* reg_ctrl, double reg_state, const double* lin,
                            const double* traj, const double* cmd, const double* hess, double* gains, double* jcst,
                            int32_t* feasible, int bad_dir) {
  const int s = ILQC_TID;
  if (s >= n_threads) return;
  const int b = prob ? prob[s] : s;
  const int os = oslot ? oslot[s] : s;
#if defined(__CUDA_ARCH__)
  if constexpr (ASYNC) {
    extern __shared__ double ilqc_bwd_smem[];
    AsyncFeed<M> feed{lin + b, (size_t)B, ilqc_bwd_smem + threadIdx.x};
    NoTensor nt;
    backward_body<M, MODE, TENSOR>(m, os, n_oslots, b, B, reg_ctrl[os], reg_state, lin, traj, cmd, hess, gains, jcst, feasible, bad_dir, feed, nt);
    return;
  }
#endif
  DirectFeed<M> feed{lin + b, (size_t)B};
  DirectTensor tf{TENSOR ? hess + b : nullptr, (size_t)B, Layout<M>::hess_stride};
  backward_body<M, MODE, TENSOR>(m, os, n_oslots, b, B, reg_ctrl[os], reg_state, lin, traj, cmd, hess, gains, jcst, feasible, bad_dir, feed, tf);
}

template <class M, int KIND>
ILQC_GLOBAL_LB(ILQC_LB_ROLL) rollout_kernel(ilqc_model_desc m, int B, int n_threads, const int32_t* prob, const int32_t* gslot,
                           int n_gslots, const int32_t* oslot, int n_oslots, const double* step, const double* x0,
                           const double* cmd, const double* traj, const double* lin, const double* gains,
                           const double* dir, double* diff, double* newval, double* diffsq) {
  const int s = ILQC_TID;
  if (s >= n_threads) return;
  const int b = prob ? prob[s] : s;
  const int gs = gslot ? gslot[s] : s;
  const int os = oslot ? oslot[s] : s;
  rollout_body<M, KIND>(m, os, n_oslots, b, B, gs, n_gslots, step[s], x0, cmd, traj, lin, gains, dir, diff, newval,
                        diffsq);
}

template <class M>
ILQC_GLOBAL_LB(ILQC_LB_FWD1) expand_kernel(ilqc_model_desc m, int B, int n_probs, const int32_t* prob, const double* traj,
                                           const double* cmd, double* lin, double* gsq) {
  const long long i = (long long)ILQC_BLOCK_BASE + ILQC_LOCAL_TID;
  if (i >= (long long)n_probs * m.horizon) return;
  const int a = (int)(i % n_probs), t = (int)(i / n_probs);
  expand_body<M>(m, prob ? prob[a] : a, B, t, traj, cmd, lin, gsq);
}
template <class M>
ILQC_GLOBAL cnorm_kernel(ilqc_model_desc m, int B, int n_probs, const int32_t* prob, const double* gsq, double* cnorm) {
  const int a = ILQC_TID;
  if (a >= n_probs) return;
  cnorm_body(prob ? prob[a] : a, B, m.horizon, gsq, cnorm);
}

// one thread per (problem, time step) and Hessian-entry group G; consecutive threads = consecutive problems
// of the same step.  Groups: jets with more than 16 second-order components are split in 3 (28-double jets of
// the bicycle model: 9.6 KB of spills per thread in one piece).
#ifndef ILQC_HESS_GROUPS
#define ILQC_HESS_GROUPS 3
#endif
template <class M>
constexpr int kHessGroups = sym_size(M::kRotVar >= 0 ? M::NV - 1 : M::NV) > 12 ? ILQC_HESS_GROUPS : 1;
template <class M, int NG, int G>
ILQC_GLOBAL_LB(ILQC_LB_BWD) hess_kernel(ilqc_model_desc m, int B, int n_probs, const int32_t* prob, const double* traj,
                                        const double* cmd, double* hess) {
  const long long i = (long long)ILQC_BLOCK_BASE + ILQC_LOCAL_TID;
  constexpr int NACC = hess_acc_doubles<M, NG, G>();
#if defined(ILQC_HOSTSIM)
  HessAccLocal<NACC> acc;
#else
  __shared__ double acc_tile[(NACC > 0 ? NACC : 1) * ILQC_BLOCK];
  HessAccShared acc{acc_tile + threadIdx.x};
#endif
  if (i >= (long long)n_probs * m.horizon) return;
  const int a = (int)(i % n_probs), t = (int)(i / n_probs);
  const int b = prob ? prob[a] : a;
  hess_body<M, NG, G>(m, b, B, t, traj, cmd, hess, acc);
}

template <class M>
ILQC_GLOBAL cnorm_lin_kernel(ilqc_model_desc m, int B, const double* lin, double* cnorm) {
  const int b = ILQC_TID;
  if (b < B) cnorm_lin_body<M>(m, b, B, lin, cnorm);
}

template <class M>
ILQC_GLOBAL adjoint_kernel(ilqc_model_desc m, int B, int n_threads, const int32_t* prob, const double* lin,
                           double* grad, double* gnorm) {
  const int a = ILQC_TID;
  if (a >= n_threads) return;
  const int b = prob ? prob[a] : a;
  adjoint_body<M>(m, b, B, lin, grad, gnorm);
}

// Dense dump of the expansion in the reference's own shapes (forward.py:209-216, :299-319) and of the
// dynamics-Hessian closures at given lambdas (forward.py:218-269).  Test / interop path, not tuned.
template <class M>
ILQC_GLOBAL expand_dense_kernel(ilqc_model_desc m, int B, const double* x0, const double* cmd, int order,
                                double* A, double* Bm, double* p, double* P, double* q, double* Q,
                                const double* lam, double* Hxx, double* Hxu, double* Huu) {
  const int b = ILQC_TID;
  if (b >= B) return;
  constexpr int NX = M::NX, NU = M::NU;
  using L = Layout<M>;
  const int H = m.horizon;
  const Weights w = load_weights(m, b, B);
  const int t0 = load_t0(m, b);
  double x[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = x0[(size_t)i * B + b];
  // costs[0]: zero gradient / Hessian (forward.py:130-134)
  for (int i = 0; i < NX; ++i) {
    p[(size_t)i * B + b] = 0.0;
    for (int j = 0; j < NX; ++j) P[((size_t)i * NX + j) * B + b] = 0.0;
  }
  for (int t = 0; t < H; ++t) {
    double u[NU], xn[NX], c, gsq, rec[L::stride];
#pragma unroll
    for (int k = 0; k < NU; ++k) u[k] = cmd[((size_t)t * NU + k) * B + b];
    expand_step<M>(m, w, x, u, t0 + t + 1, xn, &c, &gsq, [&](int slot, double v) { rec[slot] = v; });
    static_for<NX>([&](auto I) {
      constexpr int i = I;
      static_for<NX>([&](auto Jc) {
        constexpr int j = Jc;
        double a = (i == j) ? 1.0 : 0.0;
        if constexpr (M::Nm(i, j)) a += rec[L::oN + M::Nm.slot(i, j)];
        A[(((size_t)t * NX + i) * NX + j) * B + b] = a;
        double pv = 0.0;
        if constexpr (j >= i) { if constexpr (M::Pm(i, j)) pv = rec[L::oP + M::Pm.slot(i, j)]; }
        else { if constexpr (M::Pm(j, i)) pv = rec[L::oP + M::Pm.slot(j, i)]; }
        if constexpr (M::kPassivePRow >= 0) {
          if ((i == M::kPassivePRow && j == M::kPassivePCol) || (j == M::kPassivePRow && i == M::kPassivePCol))
            pv += M::passive_P(m, w);
        }
        P[(((size_t)(t + 1) * NX + i) * NX + j) * B + b] = pv;
      });
      static_for<NU>([&](auto Kc) {
        constexpr int k = Kc;
        double bv = 0.0;
        if constexpr (M::Bm(i, k)) bv = rec[L::oB + M::Bm.slot(i, k)];
        Bm[(((size_t)t * NX + i) * NU + k) * B + b] = bv;
      });
      double pv = 0.0;
      if constexpr (M::pm(0, i)) pv = rec[L::op + M::pm.slot(0, i)];
      p[((size_t)(t + 1) * NX + i) * B + b] = pv;
    });
    for (int k = 0; k < NU; ++k) {
      q[((size_t)t * NU + k) * B + b] = rec[L::oq + k];
      for (int k2 = 0; k2 < NU; ++k2) Q[(((size_t)t * NU + k) * NU + k2) * B + b] = (k == k2) ? rec[L::oQ + k] : 0.0;
    }
    if (order >= 2) {
      using S = Jet<M::NV, 2>;
      S xs[NX], us[NU], xns[NX];
      static_for<NX>([&](auto I) { xs[I] = make_var<S>(x[I], M::sv(I)); });
      static_for<NU>([&](auto I) { us[I] = make_var<S>(u[I], M::uv(I)); });
      M::template step<S>(m, xs, us, xns);
      double hl[sym_size(M::NV)];
      for (int e = 0; e < sym_size(M::NV); ++e) {
        double acc = 0.0;
        for (int i = 0; i < NX; ++i) acc += lam[((size_t)t * NX + i) * B + b] * xns[i].h[e];
        hl[e] = acc;
      }
      static_for<NX>([&](auto I) {
        constexpr int i = I;
        static_for<NX>([&](auto Jc) {
          constexpr int j = Jc;
          double v = 0.0;
          if constexpr (M::sv(i) >= 0 && M::sv(j) >= 0) v = hl[sym_idx(M::sv(i), M::sv(j), M::NV)];
          Hxx[(((size_t)t * NX + i) * NX + j) * B + b] = v;
        });
        static_for<NU>([&](auto Kc) {
          constexpr int k = Kc;
          double v = 0.0;
          if constexpr (M::sv(i) >= 0 && M::uv(k) >= 0) v = hl[sym_idx(M::sv(i), M::uv(k), M::NV)];
          Hxu[(((size_t)t * NX + i) * NU + k) * B + b] = v;
        });
      });
      static_for<NU>([&](auto Kc) {
        constexpr int k = Kc;
        static_for<NU>([&](auto K2) {
          constexpr int k2 = K2;
          double v = 0.0;
          if constexpr (M::uv(k) >= 0 && M::uv(k2) >= 0) v = hl[sym_idx(M::uv(k), M::uv(k2), M::NV)];
          Huu[(((size_t)t * NU + k) * NU + k2) * B + b] = v;
        });
      });
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = xn[i];
  }
}
}  // namespace

template <>
int Launch<ILQC_MODEL_INST>::forward(int order, const ilqc_model_desc& m, int B, int n_threads, const int32_t* prob,
                                     const double* x0, const double* cmd, double* lin, double* traj, double* cost_t,
                                     double* total, double* cnorm, stream_t st) {
  if (order == 0)
    ILQC_LAUNCH((forward_kernel<MI, 0>), n_threads, kBlock, st, m, B, n_threads, prob, x0, cmd, lin, traj, cost_t, total, cnorm);
  else
    ILQC_LAUNCH((forward_kernel<MI, 1>), n_threads, kBlock, st, m, B, n_threads, prob, x0, cmd, lin, traj, cost_t, total, cnorm);
  return ILQC_LAUNCH_OK();
}

template <>
int Launch<ILQC_MODEL_INST>::expand_parallel(const ilqc_model_desc& m, int B, int n_probs, const int32_t* prob,
                                             const double* x0, const double* cmd, double* lin, double* traj, double* gsq,
                                             double* cnorm, stream_t st) {
  // trajectory (roll-out kernel, one thread per problem), then every step of every problem at once
  ILQC_LAUNCH((forward_kernel<MI, 0>), n_probs, kBlock, st, m, B, n_probs, prob, x0, cmd, nullptr, traj, nullptr, nullptr, nullptr);
  const long long n = (long long)n_probs * m.horizon;
  ILQC_LAUNCH((expand_kernel<MI>), n, kBlock, st, m, B, n_probs, prob, traj, cmd, lin, gsq);
  if (cnorm) ILQC_LAUNCH((cnorm_kernel<MI>), n_probs, 256, st, m, B, n_probs, prob, gsq, cnorm);
  return ILQC_LAUNCH_OK();
}

template <>
int Launch<ILQC_MODEL_INST>::expand_dense(const ilqc_model_desc& m, int B, const double* x0, const double* cmd,
                                          int order, double* A, double* Bm, double* p, double* P, double* q,
                                          double* Q, const double* lam, double* Hxx, double* Hxu, double* Huu,
                                          stream_t st) {
  ILQC_LAUNCH((expand_dense_kernel<MI>), B, kBlock, st, m, B, x0, cmd, order, A, Bm, p, P, q, Q, lam, Hxx, Hxu, Huu);
  return ILQC_LAUNCH_OK();
}

template <>
int Launch<ILQC_MODEL_INST>::backward(int mode, const ilqc_model_desc& m, int B, int n_slots, const int32_t* prob,
                                      const int32_t* oslot, int n_oslots, const double* reg_ctrl, double reg_state,
                                      const double* lin, const double* traj, const double* cmd, const double* hess,
                                      double* gains, double* jcst, int32_t* feasible, stream_t st) {
  const bool dense = (mode & ILQC_BWD_DENSE_HINT) != 0;
  const int bad_dir = (mode >> ILQC_BWD_BAD_DIR_SHIFT) & 3;
  mode &= 0xf;
#define ILQC_BWD_ARGS m, B, n_slots, prob, oslot, n_oslots, reg_ctrl, reg_state, lin, traj, cmd, hess, gains, jcst, feasible, bad_dir
  if (mode == ILQC_BWD_LINQUAD) {
#if !defined(ILQC_HOSTSIM)
    if constexpr (MI::NX == 8) {
      // small (latency-bound) launches: 8 lanes per candidate (coop.cuh); identical results.  ILQC_COOP_MAX=<slots> tunes
      // the switch-over (0 = never); measured in profiles/r02_tuning_small_rounds.txt
      if ((long long)n_slots <= tuning().coop_max) {
        constexpr int CPB = ILQC_COOP_BLOCK / 8;
        backward_coop_kernel<MI, ILQC_BWD_LINQUAD><<<(n_slots + CPB - 1) / CPB, ILQC_COOP_BLOCK, 0, st>>>(
            m, B, n_slots, prob, oslot, n_oslots, reg_ctrl, reg_state, lin, gains, jcst, feasible, bad_dir);
        ++::ilqc::launch_counter();
        return ILQC_LAUNCH_OK();
      }
    }
#endif
    if constexpr (kBwdAsyncFeed<MI>) {
      // (ILQC_FEED_SMALL=<slots>: tuning knob - also stage the operands of launches of at most that many slots, which are
      //  latency-bound whatever their access pattern; measured in profiles/r02_tuning_small_rounds.txt)
      if ((dense && (long long)n_slots <= 5LL * B) || (long long)n_slots <= tuning().feed_small) {
        ILQC_LAUNCH_SMEM((backward_kernel<MI, ILQC_BWD_LINQUAD, true, false>), n_slots, kBlock, kBwdSmemBytes<MI>, st, ILQC_BWD_ARGS);
        return ILQC_LAUNCH_OK();
      }
    }
    ILQC_LAUNCH((backward_kernel<MI, ILQC_BWD_LINQUAD, false, false>), n_slots, kBlock, st, ILQC_BWD_ARGS);
  } else if (mode == ILQC_BWD_QUAD_NEWTON) {
    // hess != NULL: contraction of the stored second-order tensor (Launch::hess); NULL: recomputed in the kernel
    if (hess) ILQC_LAUNCH((backward_kernel<MI, ILQC_BWD_QUAD_NEWTON, false, true>), n_slots, kBlock, st, ILQC_BWD_ARGS);
    else ILQC_LAUNCH((backward_kernel<MI, ILQC_BWD_QUAD_NEWTON, false, false>), n_slots, kBlock, st, ILQC_BWD_ARGS);
  } else if (mode == ILQC_BWD_QUAD_DDP) {
    if (hess) ILQC_LAUNCH((backward_kernel<MI, ILQC_BWD_QUAD_DDP, false, true>), n_slots, kBlock, st, ILQC_BWD_ARGS);
    else ILQC_LAUNCH((backward_kernel<MI, ILQC_BWD_QUAD_DDP, false, false>), n_slots, kBlock, st, ILQC_BWD_ARGS);
  } else {
    return ILQC_E_ARG;
  }
#undef ILQC_BWD_ARGS
  return ILQC_LAUNCH_OK();
}

template <>
int Launch<ILQC_MODEL_INST>::hess(const ilqc_model_desc& m, int B, int n_probs, const int32_t* prob, const double* traj,
                                  const double* cmd, double* hess, stream_t st) {
  const long long n = (long long)n_probs * m.horizon;
  constexpr int NG = kHessGroups<MI>;
  ILQC_LAUNCH((hess_kernel<MI, NG, 0>), n, kBlock, st, m, B, n_probs, prob, traj, cmd, hess);
  if constexpr (NG > 1) ILQC_LAUNCH((hess_kernel<MI, NG, 1>), n, kBlock, st, m, B, n_probs, prob, traj, cmd, hess);
  if constexpr (NG > 2) ILQC_LAUNCH((hess_kernel<MI, NG, 2>), n, kBlock, st, m, B, n_probs, prob, traj, cmd, hess);
  if constexpr (NG > 3) ILQC_LAUNCH((hess_kernel<MI, NG, 3>), n, kBlock, st, m, B, n_probs, prob, traj, cmd, hess);
  static_assert(NG <= 4, "at most 4 Hessian-entry groups");
  return ILQC_LAUNCH_OK();
}

template <>
int Launch<ILQC_MODEL_INST>::rollout(int kind, const ilqc_model_desc& m, int B, int n_slots, const int32_t* prob,
                                     const int32_t* gslot, int n_gslots, const int32_t* oslot, int n_oslots,
                                     const double* step, const double* x0, const double* cmd, const double* traj,
                                     const double* lin, const double* gains, const double* dir, double* diff,
                                     double* newval, double* diffsq, stream_t st) {
  if (kind == ILQC_ROLL_EXACT)
    ILQC_LAUNCH((rollout_kernel<MI, ILQC_ROLL_EXACT>), n_slots, kBlock, st, m, B, n_slots, prob, gslot, n_gslots, oslot, n_oslots,
                step, x0, cmd, traj, lin, gains, dir, diff, newval, diffsq);
  else if (kind == ILQC_ROLL_LIN)
    ILQC_LAUNCH((rollout_kernel<MI, ILQC_ROLL_LIN>), n_slots, kBlock, st, m, B, n_slots, prob, gslot, n_gslots, oslot, n_oslots,
                step, x0, cmd, traj, lin, gains, dir, diff, newval, diffsq);
  else if (kind == ILQC_ROLL_GD)
    ILQC_LAUNCH((rollout_kernel<MI, ILQC_ROLL_GD>), n_slots, kBlock, st, m, B, n_slots, prob, gslot, n_gslots, oslot, n_oslots, step,
                x0, cmd, traj, lin, gains, dir, diff, newval, diffsq);
  else
    return ILQC_E_ARG;
  return ILQC_LAUNCH_OK();
}

template <>
int Launch<ILQC_MODEL_INST>::adjoint(const ilqc_model_desc& m, int B, int n_threads, const int32_t* prob,
                                     const double* lin, double* grad, double* gnorm, stream_t st) {
  ILQC_LAUNCH((adjoint_kernel<MI>), n_threads, kBlock, st, m, B, n_threads, prob, lin, grad, gnorm);
  return ILQC_LAUNCH_OK();
}

template <>
int Launch<ILQC_MODEL_INST>::cnorm_lin(const ilqc_model_desc& m, int B, const double* lin, double* cnorm, stream_t st) {
  ILQC_LAUNCH((cnorm_lin_kernel<MI>), B, 256, st, m, B, lin, cnorm);
  return ILQC_LAUNCH_OK();
}

template <>
void Launch<ILQC_MODEL_INST>::dims(int* nx, int* nu, int* lin_stride, int* gain_stride, int* hess_stride) {
  *nx = MI::NX;
  *nu = MI::NU;
  *lin_stride = Layout<MI>::stride;
  *gain_stride = Layout<MI>::gain_stride;
  *hess_stride = Layout<MI>::hess_stride;
}

}  // namespace ilqc

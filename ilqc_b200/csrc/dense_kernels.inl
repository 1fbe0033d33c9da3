// Generic dense Riccati backward pass / linear roll-out for small systems (lin_quad_backward,
// quad_backward's bell_step with a cross term R, roll_out_lin: envs/backward.py:6-57, :153-257,
// envs/rollout.py:8-47).  Used by external callers and by the LinQuadEnv tests (envs/lin_quad.py,
// tests/test_auto_lin_quad.py); the solver itself uses the sparsity-aware model kernels.
#include "launch.cuh"
#include "riccati.cuh"

namespace ilqc {
namespace {

template <int NX, int NU, bool HAS_R>
ILQC_GLOBAL lq_backward_dense_kernel(int H, int B, int L, const double* A, const double* Bm, const double* p,
                                     const double* P, const double* q, const double* Q, const double* R,
                                     const double* reg_ctrl, double reg_state, double* K, double* k,
                                     double* jcst) {
  const int s = ILQC_TID;
  const int n_slots = B * L;
  if (s >= n_slots) return;
  const int b = s / L;
  using Pat = DensePat<NX, NU, HAS_R>;
  Bell<Pat> bl;
  double J[sym_size(NX)], j[NX];
  // J_H = hess_state(costs[H]) + reg_state I ; j_H = grad_state(costs[H])
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    j[i] = p[((size_t)H * NX + i) * B + b];
#pragma unroll
    for (int c = i; c < NX; ++c)
      J[sym_idx(i, c, NX)] = P[(((size_t)H * NX + i) * NX + c) * B + b] + (i == c ? reg_state : 0.0);
  }
  const double reg = reg_ctrl[s];
  double acc_cst = 0.0;
  for (int t = H - 1; t >= 0; --t) {
#pragma unroll
    for (int i = 0; i < NX; ++i) {
#pragma unroll
      for (int c = 0; c < NX; ++c)
        bl.N[i][c] = A[(((size_t)t * NX + i) * NX + c) * B + b] - (i == c ? 1.0 : 0.0);
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        bl.B[i][u] = Bm[(((size_t)t * NX + i) * NU + u) * B + b];
        if constexpr (HAS_R) bl.R[i][u] = R[(((size_t)t * NX + i) * NU + u) * B + b];
      }
      bl.p[i] = p[((size_t)t * NX + i) * B + b];
#pragma unroll
      for (int c = i; c < NX; ++c)
        bl.P[i][c] = P[(((size_t)t * NX + i) * NX + c) * B + b] + (i == c ? reg_state : 0.0);
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      bl.q[u] = q[((size_t)t * NU + u) * B + b];
#pragma unroll
      for (int v = u; v < NU; ++v) bl.Q[u][v] = Q[(((size_t)t * NU + u) * NU + v) * B + b] + (u == v ? reg : 0.0);
    }
    acc_cst = acc_cst + bl.step(J, j);
#pragma unroll
    for (int u = 0; u < NU; ++u) {
#pragma unroll
      for (int c = 0; c < NX; ++c) K[(((size_t)t * NU + u) * NX + c) * n_slots + s] = bl.K[u][c];
      k[((size_t)t * NU + u) * n_slots + s] = bl.k[u];
    }
  }
  jcst[s] = acc_cst;
}

template <int NX, int NU>
ILQC_GLOBAL rollout_lin_dense_kernel(int H, int B, int L, const double* A, const double* Bm, const double* K,
                                     const double* k, const double* step, double* diff) {
  const int s = ILQC_TID;
  const int n_slots = B * L;
  if (s >= n_slots) return;
  const int b = s / L;
  double dx[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) dx[i] = 0.0;
  const double st = step[s];
  for (int t = 0; t < H; ++t) {
    double v[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < NX; ++c) acc += K[(((size_t)t * NU + u) * NX + c) * n_slots + s] * dx[c];
      v[u] = acc + st * k[((size_t)t * NU + u) * n_slots + s];
      diff[((size_t)t * NU + u) * n_slots + s] = v[u];
    }
    double dn[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < NX; ++c) acc += A[(((size_t)t * NX + i) * NX + c) * B + b] * dx[c];
#pragma unroll
      for (int u = 0; u < NU; ++u) acc += Bm[(((size_t)t * NX + i) * NU + u) * B + b] * v[u];
      dn[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) dx[i] = dn[i];
  }
}

template <int NX, int NU>
int launch_lq_dense(int H, int B, int L, const double* A, const double* Bm, const double* p, const double* P,
                    const double* q, const double* Q, const double* R, const double* reg_ctrl, double reg_state,
                    double* K, double* k, double* jcst, stream_t st) {
  const int n = B * L;
  if (R) ILQC_LAUNCH((lq_backward_dense_kernel<NX, NU, true>), n, kBlock, st, H, B, L, A, Bm, p, P, q, Q, R, reg_ctrl,
                     reg_state, K, k, jcst);
  else ILQC_LAUNCH((lq_backward_dense_kernel<NX, NU, false>), n, kBlock, st, H, B, L, A, Bm, p, P, q, Q, R, reg_ctrl,
                   reg_state, K, k, jcst);
  return ILQC_LAUNCH_OK();
}
template <int NX, int NU>
int launch_rol_dense(int H, int B, int L, const double* A, const double* Bm, const double* K, const double* k,
                     const double* step, double* diff, stream_t st) {
  ILQC_LAUNCH((rollout_lin_dense_kernel<NX, NU>), B * L, kBlock, st, H, B, L, A, Bm, K, k, step, diff);
  return ILQC_LAUNCH_OK();
}
}  // namespace

#define ILQC_DENSE_SIZES(X) X(2, 1) X(4, 1) X(4, 3) X(4, 4) X(6, 3) X(8, 3)

int dense_lq_backward(int nx, int nu, int H, int B, int L, const double* A, const double* Bm, const double* p,
                      const double* P, const double* q, const double* Q, const double* R, const double* reg_ctrl,
                      double reg_state, double* K, double* k, double* jcst, stream_t st) {
#define X(a, b) \
  if (nx == a && nu == b) return launch_lq_dense<a, b>(H, B, L, A, Bm, p, P, q, Q, R, reg_ctrl, reg_state, K, k, jcst, st);
  ILQC_DENSE_SIZES(X)
#undef X
  return ILQC_E_UNSUPPORTED;
}
int dense_rollout_lin(int nx, int nu, int H, int B, int L, const double* A, const double* Bm, const double* K,
                      const double* k, const double* step, double* diff, stream_t st) {
#define X(a, b) \
  if (nx == a && nu == b) return launch_rol_dense<a, b>(H, B, L, A, Bm, K, k, step, diff, st);
  ILQC_DENSE_SIZES(X)
#undef X
  return ILQC_E_UNSUPPORTED;
}

}  // namespace ilqc

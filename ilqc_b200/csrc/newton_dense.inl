// Dense Newton step on the flattened control problem and the Jacobian of the trajectory with respect to the
// controls, for B problems at once.  Diagnostics of the reference, not the solver's hot path:
//   newton_oracle                   algorithms/algos_steps.py:46-102   (tests/test_newton.py cross-checks the structured
//                                                                        Newton direction of quad_backward against it)
//   compute_min_sing_eigval_jac     algorithms/run_min_algo.py:103-124
//
// The reference obtains the (H nu x H nu) Hessian of the total cost by differentiating its autograd graph twice.  Here
// it comes from the analytic expansion (ilqc_expand_dense): with the costate lam_{t+1} = d f / d x_{t+1}
//   lam_H = p_H ;  lam_t = p_t + A_t^T lam_{t+1} ;  grad_t = q_t + B_t^T lam_{t+1}
// column (s, k) of the Hessian is the second-order adjoint of the tangent dx of the unit control e_(s,k):
//   dx_{t+1} = A_t dx_t + B_t v_t ,  v_t = e_k [t == s]
//   dlam_H = P_H dx_H ;  dlam_t = (P_t + <lam_{t+1}, phi_xx>) dx_t + <lam_{t+1}, phi_xu> v_t + A_t^T dlam_{t+1}
//   (H v)_t = (Q_t + <lam_{t+1}, phi_uu>) v_t + <lam_{t+1}, phi_xu>^T dx_t + B_t^T dlam_{t+1}
// one thread per (problem, column); the linear system (Hessian + reg I) d = -grad is then solved by one thread block
// per problem with LU + partial pivoting (what torch.linalg.solve does).
#pragma once
#include "launch.cuh"

namespace ilqc {
namespace {

constexpr int kNdMaxNx = 8, kNdMaxNu = 9;

struct NdDense {  // dense expansion in the engine layout ([..][B], ilqc_expand_dense) + scratch
  int nx, nu, H, B;
  const double *A, *Bm, *p, *P, *q, *Q, *Hxx, *Hxu, *Huu;
};

// costate lam[t] = d f / d x_{t+1} (t < H) and the gradient of the total cost
ILQC_GLOBAL nd_costate_kernel(NdDense E, double* lam, double* grad) {
  const int b = ILQC_TID;
  if (b >= E.B) return;
  const int nx = E.nx, nu = E.nu, H = E.H;
  const size_t B = (size_t)E.B;
  double l[kNdMaxNx], ln[kNdMaxNx];
  for (int i = 0; i < nx; ++i) l[i] = E.p[((size_t)H * nx + i) * B + b];
  for (int t = H - 1; t >= 0; --t) {
    for (int i = 0; i < nx; ++i) lam[((size_t)t * nx + i) * B + b] = l[i];
    for (int k = 0; k < nu; ++k) {
      double acc = E.q[((size_t)t * nu + k) * B + b];
      for (int i = 0; i < nx; ++i) acc += E.Bm[(((size_t)t * nx + i) * nu + k) * B + b] * l[i];
      grad[((size_t)t * nu + k) * B + b] = acc;
    }
    if (t > 0) {
      for (int i = 0; i < nx; ++i) {
        double acc = E.p[((size_t)t * nx + i) * B + b];
        for (int r = 0; r < nx; ++r) acc += E.A[(((size_t)t * nx + r) * nx + i) * B + b] * l[r];
        ln[i] = acc;
      }
      for (int i = 0; i < nx; ++i) l[i] = ln[i];
    }
  }
}

// tangent of the unit control (s, k): dx_t for t = s+1 .. H.  jac (nullable): [B][H nx][H nu] (row = state
// (t-1, i) of traj[1:], column = control (s, k)) ; D (nullable): scratch [H+1][nx][n B] for the Hessian kernel
ILQC_GLOBAL nd_tangent_kernel(NdDense E, double* D, double* jac) {
  const int id = ILQC_TID;
  const int nx = E.nx, nu = E.nu, H = E.H, n = H * nu;
  if (id >= E.B * n) return;
  const int b = id / n, col = id % n, s = col / nu, k = col % nu;
  const size_t B = (size_t)E.B, nB = (size_t)n * B;
  double dx[kNdMaxNx], dn[kNdMaxNx];
  for (int i = 0; i < nx; ++i) dx[i] = 0.0;
  for (int t = 0; t <= H; ++t) {
    if (t > 0) {
      const int tm = t - 1;  // dx_t = A_{t-1} dx_{t-1} + B_{t-1} v_{t-1}
      if (tm < s) {
        for (int i = 0; i < nx; ++i) dn[i] = 0.0;
      } else if (tm == s) {
        for (int i = 0; i < nx; ++i) dn[i] = E.Bm[(((size_t)tm * nx + i) * nu + k) * B + b];
      } else {
        for (int i = 0; i < nx; ++i) {
          double acc = 0.0;
          for (int r = 0; r < nx; ++r) acc += E.A[(((size_t)tm * nx + i) * nx + r) * B + b] * dx[r];
          dn[i] = acc;
        }
      }
      for (int i = 0; i < nx; ++i) dx[i] = dn[i];
      if (jac)
        for (int i = 0; i < nx; ++i) jac[((size_t)b * H * nx + (size_t)tm * nx + i) * n + col] = dx[i];
    }
    if (D)
      for (int i = 0; i < nx; ++i) D[((size_t)t * nx + i) * nB + id] = dx[i];
  }
}

// second-order adjoint of column (s, k): hess [B][n][n] (row-major per problem; the Hessian is symmetric)
ILQC_GLOBAL nd_hess_kernel(NdDense E, const double* D, double* hess) {
  const int id = ILQC_TID;
  const int nx = E.nx, nu = E.nu, H = E.H, n = H * nu;
  if (id >= E.B * n) return;
  const int b = id / n, col = id % n, s = col / nu, k = col % nu;
  const size_t B = (size_t)E.B, nB = (size_t)n * B;
  double dl[kNdMaxNx], dn[kNdMaxNx], dx[kNdMaxNx];
  // dlam_H = P_H dx_H
  for (int i = 0; i < nx; ++i) dx[i] = D[((size_t)H * nx + i) * nB + id];
  for (int i = 0; i < nx; ++i) {
    double acc = 0.0;
    for (int r = 0; r < nx; ++r) acc += E.P[(((size_t)H * nx + i) * nx + r) * B + b] * dx[r];
    dl[i] = acc;
  }
  double* hb = hess + (size_t)b * n * n;
  for (int t = H - 1; t >= 0; --t) {
    for (int i = 0; i < nx; ++i) dx[i] = D[((size_t)t * nx + i) * nB + id];
    // (H v)_t = (Q + Huu) v_t + Hxu^T dx_t + B_t^T dlam_{t+1}
    for (int u = 0; u < nu; ++u) {
      double acc = 0.0;
      if (t == s) acc = E.Q[(((size_t)t * nu + u) * nu + k) * B + b] + E.Huu[(((size_t)t * nu + u) * nu + k) * B + b];
      for (int i = 0; i < nx; ++i) {
        acc += E.Hxu[(((size_t)t * nx + i) * nu + u) * B + b] * dx[i];
        acc += E.Bm[(((size_t)t * nx + i) * nu + u) * B + b] * dl[i];
      }
      hb[(size_t)(t * nu + u) * n + col] = acc;
    }
    if (t > 0) {
      // dlam_t = (P_t + Hxx_t) dx_t + Hxu_t v_t + A_t^T dlam_{t+1}
      for (int i = 0; i < nx; ++i) {
        double acc = (t == s) ? E.Hxu[(((size_t)t * nx + i) * nu + k) * B + b] : 0.0;
        for (int r = 0; r < nx; ++r) {
          acc += (E.P[(((size_t)t * nx + i) * nx + r) * B + b] + E.Hxx[(((size_t)t * nx + i) * nx + r) * B + b]) * dx[r];
          acc += E.A[(((size_t)t * nx + r) * nx + i) * B + b] * dl[r];
        }
        dn[i] = acc;
      }
      for (int i = 0; i < nx; ++i) dl[i] = dn[i];
    }
  }
}

// (hess + reg I) x = grad by LU with partial pivoting ; dir = -x, opt = dir . grad / 2 (algos_steps.py:64-66).
// M: scratch [B][n][n], rhs: scratch [B][n].
#if defined(ILQC_HOSTSIM)
ILQC_GLOBAL nd_solve_kernel(int B, int n, int nu, const double* hess, const double* grad, const double* reg, double* M,
                            double* rhs, double* dir, double* opt) {
  const int b = ILQC_TID;
  if (b >= B) return;
  double* Mb = M + (size_t)b * n * n;
  double* rb = rhs + (size_t)b * n;
  for (int r = 0; r < n; ++r) {
    for (int c = 0; c < n; ++c) Mb[(size_t)r * n + c] = hess[((size_t)b * n + r) * n + c] + (r == c ? reg[b] : 0.0);
    rb[r] = grad[(size_t)r * B + b];
  }
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (::fabs(Mb[(size_t)r * n + c]) > ::fabs(Mb[(size_t)piv * n + c])) piv = r;
    if (piv != c) {
      for (int cc = 0; cc < n; ++cc) { const double t = Mb[(size_t)c * n + cc]; Mb[(size_t)c * n + cc] = Mb[(size_t)piv * n + cc]; Mb[(size_t)piv * n + cc] = t; }
      const double t = rb[c]; rb[c] = rb[piv]; rb[piv] = t;
    }
    for (int r = c + 1; r < n; ++r) {
      const double f = Mb[(size_t)r * n + c] / Mb[(size_t)c * n + c];
      for (int cc = c + 1; cc < n; ++cc) Mb[(size_t)r * n + cc] -= f * Mb[(size_t)c * n + cc];
      rb[r] -= f * rb[c];
    }
  }
  double o = 0.0;
  for (int r = n - 1; r >= 0; --r) {
    double acc = rb[r];
    for (int cc = r + 1; cc < n; ++cc) acc -= Mb[(size_t)r * n + cc] * rb[cc];
    rb[r] = acc / Mb[(size_t)r * n + r];
  }
  for (int r = 0; r < n; ++r) {
    const double d = -rb[r];
    dir[(size_t)r * B + b] = d;
    o += d * grad[(size_t)r * B + b];
  }
  opt[b] = 0.5 * o;
}
constexpr int kNdSolveThreads = 1;
#else
constexpr int kNdSolveThreads = 256;
__global__ void __launch_bounds__(kNdSolveThreads) nd_solve_kernel(int B, int n, int nu, const double* hess, const double* grad,
                                                                   const double* reg, double* M, double* rhs, double* dir,
                                                                   double* opt) {
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  if (b >= B) return;
  double* Mb = M + (size_t)b * n * n;
  double* rb = rhs + (size_t)b * n;
  __shared__ double s_val[kNdSolveThreads];
  __shared__ int s_idx[kNdSolveThreads];
  __shared__ double s_piv;
  for (int e = tid; e < n * n; e += nt) {
    const int r = e / n, c = e % n;
    Mb[e] = hess[(size_t)b * n * n + e] + (r == c ? reg[b] : 0.0);
  }
  for (int r = tid; r < n; r += nt) rb[r] = grad[(size_t)r * B + b];
  __syncthreads();
  for (int c = 0; c < n; ++c) {
    // pivot: first row of maximal |M[r][c]|, r >= c
    double best = -1.0;
    int bi = c;
    for (int r = c + tid; r < n; r += nt) {
      const double v = ::fabs(Mb[(size_t)r * n + c]);
      if (v > best) { best = v; bi = r; }
    }
    s_val[tid] = best; s_idx[tid] = bi;
    __syncthreads();
    for (int o = nt / 2; o > 0; o >>= 1) {
      if (tid < o) {
        const double v2 = s_val[tid + o];
        const int i2 = s_idx[tid + o];
        if (v2 > s_val[tid] || (v2 == s_val[tid] && i2 < s_idx[tid])) { s_val[tid] = v2; s_idx[tid] = i2; }
      }
      __syncthreads();
    }
    const int piv = s_idx[0];
    if (piv != c) {
      for (int cc = tid; cc < n; cc += nt) {
        const double t = Mb[(size_t)c * n + cc];
        Mb[(size_t)c * n + cc] = Mb[(size_t)piv * n + cc];
        Mb[(size_t)piv * n + cc] = t;
      }
      if (tid == 0) { const double t = rb[c]; rb[c] = rb[piv]; rb[piv] = t; }
    }
    __syncthreads();
    if (tid == 0) s_piv = Mb[(size_t)c * n + c];
    __syncthreads();
    const double pv = s_piv, rc = rb[c];
    // one warp-strided pass over the rows below the pivot: thread (r, cc-chunk)
    const int rows = n - c - 1, cols = n - c - 1;
    for (int r = c + 1 + tid / 32; r < n; r += nt / 32) {
      const double f = Mb[(size_t)r * n + c] / pv;
      for (int cc = c + 1 + (tid & 31); cc < n; cc += 32) Mb[(size_t)r * n + cc] -= f * Mb[(size_t)c * n + cc];
      if ((tid & 31) == 0) rb[r] -= f * rc;
    }
    (void)rows; (void)cols;
    __syncthreads();
  }
  // back substitution (one row at a time; the dot product is shared by the block)
  for (int r = n - 1; r >= 0; --r) {
    double acc = 0.0;
    for (int cc = r + 1 + tid; cc < n; cc += nt) acc += Mb[(size_t)r * n + cc] * rb[cc];
    s_val[tid] = acc;
    __syncthreads();
    for (int o = nt / 2; o > 0; o >>= 1) {
      if (tid < o) s_val[tid] += s_val[tid + o];
      __syncthreads();
    }
    if (tid == 0) rb[r] = (rb[r] - s_val[0]) / Mb[(size_t)r * n + r];
    __syncthreads();
  }
  double o = 0.0;
  for (int r = tid; r < n; r += nt) {
    const double d = -rb[r];
    dir[(size_t)r * B + b] = d;
    o += d * grad[(size_t)r * B + b];
  }
  s_val[tid] = o;
  __syncthreads();
  for (int w = nt / 2; w > 0; w >>= 1) {
    if (tid < w) s_val[tid] += s_val[tid + w];
    __syncthreads();
  }
  if (tid == 0) opt[b] = 0.5 * s_val[0];
}
#endif

struct NdCarve { size_t A, Bm, p, P, q, Q, lam, Hxx, Hxu, Huu, grad, D, hess, M, rhs, total; };
inline void nd_carve(NdCarve* c, int nx, int nu, int H, int B, bool need_second_order) {
  size_t off = 0;
  const size_t n = (size_t)H * nu;
  auto take = [&](size_t doubles) { const size_t o = off; off = (off + doubles * sizeof(double) + 255) & ~(size_t)255; return o; };
  c->A = take((size_t)H * nx * nx * B); c->Bm = take((size_t)H * nx * nu * B);
  c->p = take((size_t)(H + 1) * nx * B); c->P = take((size_t)(H + 1) * nx * nx * B);
  c->q = take((size_t)H * nu * B); c->Q = take((size_t)H * nu * nu * B);
  if (need_second_order) {
    c->lam = take((size_t)H * nx * B);
    c->Hxx = take((size_t)H * nx * nx * B); c->Hxu = take((size_t)H * nx * nu * B); c->Huu = take((size_t)H * nu * nu * B);
    c->grad = take(n * B);
    c->D = take((size_t)(H + 1) * nx * n * B);
    c->hess = take(n * n * B); c->M = take(n * n * B); c->rhs = take(n * B);
  }
  c->total = off;
}

}  // namespace
}  // namespace ilqc

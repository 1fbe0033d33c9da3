// One step of the regularised Riccati / Bellman recursion, thread-per-problem, everything in registers.
//
// Restates bell_step + bell_core (envs/backward.py:153-257):
//     H = Q + B^T J B ;  K = -H^{-1}(R^T + B^T J A) ;  k = -H^{-1}(q + B^T j)
//     J' = P + A^T J A + (R + A^T J B) K ;  j' = p + A^T j + A^T J B k + R k ;  rest = (q + B^T j).k / 2
// with torch.linalg.solve = LU with partial pivoting (two right-hand sides solved together here).
//
// B200-first formulation: the discretised dynamics are always x' = x + (...), so the Jacobian is kept as
// A = I + N and only the structurally non-zero entries of N, B, P, p (compile-time masks of the model,
// models.cuh) are ever touched:
//     W = J N ;  A^T J A = J + W + W^T + N^T W ;  G = R^T + B^T (J + W)  (= (R + A^T J B)^T, J symmetric)
//     J' = P + J + W + W^T + N^T W + G^T K ;  j' = p + j + N^T j + G^T k
// For the bicycle car this is ~0.75 kFMA per step instead of 2.2 kFMA for the dense formulas.
// J is carried as a packed symmetric matrix (the reference's J is symmetric up to rounding).
#pragma once
#include "common.cuh"
#include "models.cuh"

namespace ilqc {

// Pattern adapter: masks seen by the recursion for model M, with (QUAD) or without the dynamics-Hessian
// blocks of the second-order variants (backward.py:102-114).
template <class M, bool QUAD_>
struct PatOf {
  static constexpr int NX = M::NX, NU = M::NU;
  static constexpr bool QUAD = QUAD_;
  static constexpr auto Nm = M::Nm;
  static constexpr auto Bm = M::Bm;
  static constexpr auto pm = M::pm;
  static constexpr auto Pm = QUAD_ ? mask_or(M::Pm, Layout<M>::hxx()) : M::Pm;
  static constexpr auto Rm = Layout<M>::hxu();
  // Q: diagonal cost Hessian (+ full symmetric block on the jet controls when QUAD)
  static constexpr bool Qm(int u, int v) { return u == v || (QUAD_ && Layout<M>::huu()(u < v ? u : v, u < v ? v : u)); }
};
template <int NX_, int NU_, bool HAS_R>
struct DensePat {
  static constexpr int NX = NX_, NU = NU_;
  static constexpr bool QUAD = HAS_R;
  static constexpr auto Nm = full_mask<NX_, NX_>();
  static constexpr auto Bm = full_mask<NX_, NU_>();
  static constexpr auto pm = full_mask<1, NX_>();
  static constexpr auto Pm = upper_of(full_mask<NX_, NX_>());
  static constexpr auto Rm = full_mask<NX_, NU_>();
  static constexpr bool Qm(int, int) { return true; }
};

template <class Pat>
struct Bell {
  static constexpr int NX = Pat::NX, NU = Pat::NU;
  static constexpr int NJ = sym_size(NX);

  // operands (register arrays; entries outside the masks are never read)
  double N[NX][NX];
  double B[NX][NU];
  double P[NX][NX];  // upper triangle
  double p[NX];
  double Q[NU][NU];  // upper triangle
  double q[NU];
  double R[NX][NU];
  // results
  double K[NU][NX];
  double k[NU];

  // J (packed upper, sym_idx) and j are updated in place; returns `rest`
  ILQC_HD double step(double (&J)[NJ], double (&j)[NX]) {
    auto Jat = [&](int a, int b) -> double { return J[sym_idx(a, b, NX)]; };
    // W = J N (only the active columns of N)
    double W[NX][NX];
    static_for<NX>([&](auto C) {
      constexpr int c = C;
      if constexpr (Pat::Nm.col_any(c)) {
        static_for<NX>([&](auto I) {
          constexpr int i = I;
          double acc = 0.0;
          static_for<NX>([&](auto Rr) {
            constexpr int r = Rr;
            if constexpr (Pat::Nm(r, c)) acc += Jat(i, r) * N[r][c];
          });
          W[i][c] = acc;
        });
      }
    });
    // JB = J B
    double JB[NX][NU];
    static_for<NU>([&](auto U) {
      constexpr int u = U;
      static_for<NX>([&](auto I) {
        constexpr int i = I;
        double acc = 0.0;
        static_for<NX>([&](auto Rr) {
          constexpr int r = Rr;
          if constexpr (Pat::Bm(r, u)) acc += Jat(i, r) * B[r][u];
        });
        JB[i][u] = acc;
      });
    });
    // G = R^T + B^T (J + W)   (NU x NX)
    double G[NU][NX];
    static_for<NU>([&](auto U) {
      constexpr int u = U;
      static_for<NX>([&](auto C) {
        constexpr int c = C;
        double acc = JB[c][u];
        if constexpr (Pat::QUAD && Pat::Rm(c, u)) acc += R[c][u];
        if constexpr (Pat::Nm.col_any(c)) {
          static_for<NX>([&](auto Rr) {
            constexpr int r = Rr;
            if constexpr (Pat::Bm(r, u)) acc += B[r][u] * W[r][c];
          });
        }
        G[u][c] = acc;
      });
    });
    // H = Q + B^T J B (symmetric: upper computed, mirrored) ; gq = q + B^T j
    double Hm[NU][NU];
    double gq[NU];
    static_for<NU>([&](auto U) {
      constexpr int u = U;
      static_for<NU>([&](auto V) {
        constexpr int v = V;
        if constexpr (v >= u) {
          double acc = 0.0;
          if constexpr (Pat::Qm(u, v)) acc = Q[u][v];
          static_for<NX>([&](auto Rr) {
            constexpr int r = Rr;
            if constexpr (Pat::Bm(r, u)) acc += B[r][u] * JB[r][v];
          });
          Hm[u][v] = acc;
        }
      });
      double acc = q[u];
      static_for<NX>([&](auto Rr) {
        constexpr int r = Rr;
        if constexpr (Pat::Bm(r, u)) acc += B[r][u] * j[r];
      });
      gq[u] = acc;
    });
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int v = 0; v < u; ++v) Hm[u][v] = Hm[v][u];

    // solve H [X | x] = [G | gq] by LU with partial pivoting ; K = -X, k = -x
    double rhs[NU][NX + 1];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
#pragma unroll
      for (int c = 0; c < NX; ++c) rhs[u][c] = G[u][c];
      rhs[u][NX] = gq[u];
    }
    double ipiv[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) {
#pragma unroll
      for (int r = c + 1; r < NU; ++r) {
        const bool sw = ::fabs(Hm[r][c]) > ::fabs(Hm[c][c]);
#pragma unroll
        for (int cc = 0; cc < NU; ++cc) {
          const double a = Hm[c][cc], b = Hm[r][cc];
          Hm[c][cc] = sw ? b : a;
          Hm[r][cc] = sw ? a : b;
        }
#pragma unroll
        for (int cc = 0; cc < NX + 1; ++cc) {
          const double a = rhs[c][cc], b = rhs[r][cc];
          rhs[c][cc] = sw ? b : a;
          rhs[r][cc] = sw ? a : b;
        }
      }
      // one reciprocal per pivot (an FP64 division is ~25 instructions; the NX+1 right-hand sides would
      // otherwise pay it NU*(NX+1) times per step)
      ipiv[c] = 1.0 / Hm[c][c];
#pragma unroll
      for (int r = c + 1; r < NU; ++r) {
        const double f = Hm[r][c] * ipiv[c];
#pragma unroll
        for (int cc = c + 1; cc < NU; ++cc) Hm[r][cc] -= f * Hm[c][cc];
#pragma unroll
        for (int cc = 0; cc < NX + 1; ++cc) rhs[r][cc] -= f * rhs[c][cc];
      }
    }
#pragma unroll
    for (int r = NU - 1; r >= 0; --r) {
#pragma unroll
      for (int cc = 0; cc < NX + 1; ++cc) {
        double acc = rhs[r][cc];
#pragma unroll
        for (int c2 = r + 1; c2 < NU; ++c2) acc -= Hm[r][c2] * rhs[c2][cc];
        rhs[r][cc] = acc * ipiv[r];
      }
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
#pragma unroll
      for (int c = 0; c < NX; ++c) K[u][c] = -rhs[u][c];
      k[u] = -rhs[u][NX];
    }

    // rest = (q + B^T j) . k / 2
    double rest = 0.0;
#pragma unroll
    for (int u = 0; u < NU; ++u) rest += gq[u] * k[u];
    rest *= 0.5;

    // j' = p + j + N^T j + G^T k
    double jn[NX];
    static_for<NX>([&](auto I) {
      constexpr int i = I;
      double acc = j[i];
      if constexpr (Pat::pm(0, i)) acc += p[i];
      static_for<NX>([&](auto Rr) {
        constexpr int r = Rr;
        if constexpr (Pat::Nm(r, i)) acc += N[r][i] * j[r];
      });
#pragma unroll
      for (int u = 0; u < NU; ++u) acc += G[u][i] * k[u];
      jn[i] = acc;
    });
    // J' = P + J + W + W^T + N^T W + G^T K  (upper triangle)
    double Jn[NJ];
    static_for<NX>([&](auto I) {
      constexpr int i = I;
      static_for<NX>([&](auto C) {
        constexpr int c = C;
        if constexpr (c >= i) {
          double acc = Jat(i, c);
          if constexpr (Pat::Pm(i, c)) acc += P[i][c];
          if constexpr (Pat::Nm.col_any(c)) acc += W[i][c];
          if constexpr (Pat::Nm.col_any(i)) acc += W[c][i];
          if constexpr (Pat::Nm.col_any(c) && Pat::Nm.col_any(i)) {
            static_for<NX>([&](auto Rr) {
              constexpr int r = Rr;
              if constexpr (Pat::Nm(r, i)) acc += N[r][i] * W[r][c];
            });
          }
#pragma unroll
          for (int u = 0; u < NU; ++u) acc += G[u][i] * K[u][c];
          Jn[sym_idx(i, c, NX)] = acc;
        }
      });
    });
#pragma unroll
    for (int i = 0; i < NJ; ++i) J[i] = Jn[i];
#pragma unroll
    for (int i = 0; i < NX; ++i) j[i] = jn[i];
    return rest;
  }
};

}  // namespace ilqc

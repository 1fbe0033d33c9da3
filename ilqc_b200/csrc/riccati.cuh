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
  // entries of P that receive the state block of the dynamics Hessian (Bell::Padd)
  static constexpr bool padd(int i, int c) { return QUAD_ && Layout<M>::hxx()(i, c); }
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
  static constexpr bool padd(int, int) { return false; }
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
  // added to P at its point of use in phase 2, in this order: (P + Padd) + reg_diag on the diagonal.  Kept apart from P
  // so that nothing touches the freshly loaded P, p before the LU factorisation has been issued: an early
  // `P[i][i] += reg_state` made the warp wait for the load right there (18 % of the stall samples of the gathered
  // rounds, ncu source page round 2) instead of behind the three divisions of the LU.
  double Padd[NX][NX];  // upper triangle, entries Pat::padd only (state block of the dynamics Hessian)
  double reg_diag = 0.0;
  // intermediates carried from phase1 to phase2
  double G[NU][NX];
  double Hm[NU][NU];
  double gq[NU];
  double jt[NX];
  // results
  double K[NU][NX];
  double k[NU];

  // The step is split in two phases so that the state-cost operands (P, p) can be loaded AFTER everything
  // that needs N, B and W has been consumed: the recursion sits at the 255-register limit and every value
  // kept live across the whole step turns into local-memory traffic (ncu: half of the backward kernel's
  // stall cycles were long_scoreboard on spills).  J is updated in place: J'(i,c) only reads J(i,c).
  //
  // phase1 needs N, B, Q, q (and R when Pat::QUAD):
  //   W = J N ; JB = J B ; H = Q + B^T JB ; gq = q + B^T j ; G = R^T + JB^T + B^T W ; jt = j + N^T j ;
  //   J += W + W^T + N^T W
  ILQC_HD void phase1(double (&J)[NJ], const double (&j)[NX]) {
    auto Jat = [&](int a, int b) -> double { return J[sym_idx(a, b, NX)]; };
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
    // H = Q + B^T J B (upper computed, mirrored) ; gq = q + B^T j
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
    // G = R^T + B^T (J + W)   (NU x NX);  B^T J = (J B)^T because J is symmetric
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
    // jt = j + N^T j
    static_for<NX>([&](auto I) {
      constexpr int i = I;
      double acc = j[i];
      static_for<NX>([&](auto Rr) {
        constexpr int r = Rr;
        if constexpr (Pat::Nm(r, i)) acc += N[r][i] * j[r];
      });
      jt[i] = acc;
    });
    // J += W + W^T + N^T W  (upper triangle, in place)
    static_for<NX>([&](auto I) {
      constexpr int i = I;
      static_for<NX>([&](auto C) {
        constexpr int c = C;
        if constexpr (c >= i && (Pat::Nm.col_any(c) || Pat::Nm.col_any(i))) {
          double acc = J[sym_idx(i, c, NX)];
          if constexpr (Pat::Nm.col_any(c)) acc += W[i][c];
          if constexpr (Pat::Nm.col_any(i)) acc += W[c][i];
          if constexpr (Pat::Nm.col_any(c) && Pat::Nm.col_any(i)) {
            static_for<NX>([&](auto Rr) {
              constexpr int r = Rr;
              if constexpr (Pat::Nm(r, i)) acc += N[r][i] * W[r][c];
            });
          }
          J[sym_idx(i, c, NX)] = acc;
        }
      });
    });
  }

  // phase2 needs P, p:  K = -H^{-1} G ; k = -H^{-1} gq (LU with partial pivoting, both right-hand sides at
  // once) ; J += P + G^T K ; j = jt + p + G^T k ; returns rest = gq.k / 2
  ILQC_HD double phase2(double (&J)[NJ], double (&j)[NX]) {
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
    double rest = 0.0;
#pragma unroll
    for (int u = 0; u < NU; ++u) rest += gq[u] * k[u];
    rest *= 0.5;
    // j = jt + p + G^T k
    static_for<NX>([&](auto I) {
      constexpr int i = I;
      double acc = jt[i];
      if constexpr (Pat::pm(0, i)) acc += p[i];
#pragma unroll
      for (int u = 0; u < NU; ++u) acc += G[u][i] * k[u];
      j[i] = acc;
    });
    // J += P + G^T K  (upper triangle, in place)
    static_for<NX>([&](auto I) {
      constexpr int i = I;
      static_for<NX>([&](auto C) {
        constexpr int c = C;
        if constexpr (c >= i) {
          double acc = J[sym_idx(i, c, NX)];
          // (the diagonal also outside the stored pattern: it carries reg_state, backward.py:35 / :105)
          if constexpr (Pat::Pm(i, c) || i == c) {
            double pv = P[i][c];
            if constexpr (Pat::padd(i, c)) pv += Padd[i][c];
            if constexpr (i == c) pv += reg_diag;
            acc += pv;
          }
#pragma unroll
          for (int u = 0; u < NU; ++u) acc += G[u][i] * K[u][c];
          J[sym_idx(i, c, NX)] = acc;
        }
      });
    });
    return rest;
  }

  // whole step with all operands already loaded
  ILQC_HD double step(double (&J)[NJ], double (&j)[NX]) {
    phase1(J, j);
    return phase2(J, j);
  }
};

}  // namespace ilqc

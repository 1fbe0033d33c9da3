// Warp-cooperative backward pass: NX lanes per line-search candidate (lane i owns row i of the cost-to-go matrix J).
//
// Why: the thread-per-candidate recursion (bodies.cuh backward_body + riccati.cuh) is a chain of ~1400 dependent
// instructions per time step at 255 registers.  That is the right shape when every SM holds its 8 warps (first round of
// an iteration: 131072 candidates), but the later line-search rounds and MPC windows launch a few hundred to a few thousand
// candidates: one warp per scheduler or less, and the launch then lasts H x (latency of that chain) whatever its size
// (0.4 ms for 1 candidate or 8000, profiles/r01_round_policy.txt).  Here 8 lanes share one candidate, each lane runs a
// ~4x shorter chain at ~90 registers, and the rows meet through a 1 KB shared-memory tile per candidate (warp-level
// synchronisation only: the 8 lanes of a candidate sit in one warp).
//
// Same arithmetic as riccati.cuh, entry by entry: every dot product is accumulated by ONE lane in the order of the
// thread-per-candidate code (no tree reductions), J' is computed on its upper triangle and mirrored, the 3x3 LU runs
// redundantly on every lane.  The two kernels therefore return identical gains (tests/test_gpu_parity.py::test_coop_backward),
// which keeps results independent of which kernel a round happens to use.
//
// Restates bell_step + bell_core (envs/backward.py:153-257) of the reference, like riccati.cuh.
#pragma once
#include "bodies.cuh"

namespace ilqc {

#if defined(__CUDACC__) && !defined(ILQC_HOSTSIM)

#ifndef ILQC_COOP_BLOCK
#define ILQC_COOP_BLOCK 128  // threads per block = 16 candidates
#endif

// compile-time tables of the packed record, looked up with the (runtime) lane index once per thread
template <class M>
struct CoopTables {
  using L = Layout<M>;
  static constexpr int NX = M::NX;
  // largest number of non-zeros in a column of N
  static constexpr int kn() {
    int best = 0;
    for (int c = 0; c < NX; ++c) {
      int n = 0;
      for (int r = 0; r < NX; ++r) n += M::Nm(r, c) ? 1 : 0;
      best = n > best ? n : best;
    }
    return best;
  }
  static constexpr int KN = kn();
  // k-th non-zero of column c of N: row (or -1) and slot in the record
  static constexpr int ncol_row(int c, int k) {
    int n = 0;
    for (int r = 0; r < NX; ++r)
      if (M::Nm(r, c)) {
        if (n == k) return r;
        ++n;
      }
    return -1;
  }
  static constexpr int ncol_slot(int c, int k) {
    const int r = ncol_row(c, k);
    return r < 0 ? -1 : L::oN + M::Nm.slot(r, c);
  }
  static constexpr int p_slot(int i) { return M::pm(0, i) ? L::op + M::pm.slot(0, i) : -1; }
  static constexpr int P_slot(int i, int c) { return (c >= i && M::Pm(i, c)) ? L::oP + M::Pm.slot(i, c) : -1; }
};

template <class M, int MODE>
__global__ void __launch_bounds__(ILQC_COOP_BLOCK)
backward_coop_kernel(ilqc_model_desc m, int B, int n_slots_active, const int32_t* __restrict__ prob,
                     const int32_t* __restrict__ oslot, int n_oslots, const double* __restrict__ reg_ctrl, double reg_state,
                     const double* __restrict__ lin, double* __restrict__ gains, double* __restrict__ jcst_out,
                     int32_t* __restrict__ feas_out, int bad_dir) {
  static_assert(MODE == ILQC_BWD_LINQUAD, "cooperative backward pass: linear-quadratic mode");
  constexpr int NX = M::NX, NU = M::NU;
  static_assert(NX == 8, "one lane per state row, 8 lanes per candidate");
  using L = Layout<M>;
  using Pat = PatOf<M, false>;
  using Tb = CoopTables<M>;
  constexpr int KN = Tb::KN;
  constexpr int CPB = ILQC_COOP_BLOCK / NX;  // candidates per block
  constexpr int TS = NX + 1;                  // padded row stride of the 8x8 tile (bank conflicts)
  __shared__ double s_tile[CPB][NX * TS];     // W (phase 1), then J rows for the mirror
  __shared__ double s_JB[CPB][NX * NU];
  __shared__ double s_j[CPB][NX];
  __shared__ double s_K[CPB][NU * NX];

  const int cand = threadIdx.x / NX, i = threadIdx.x % NX;
  const int a = blockIdx.x * CPB + cand;
  const bool live = a < n_slots_active;
  // (all lanes of the warp take part in the warp-level synchronisations: dead candidates run on candidate 0's data)
  const int aa = live ? a : 0;
  const int b = prob ? prob[aa] : aa;
  const int os = oslot ? oslot[aa] : aa;
  const double reg = reg_ctrl[os];
  const int H = m.horizon;
  double* tile = s_tile[cand];
  double* JBs = s_JB[cand];
  double* js = s_j[cand];
  double* Ks = s_K[cand];
  const double* lin_b = lin + b;
  const size_t Bs = (size_t)B;

  // per-lane slots of the record (column i of N, row i of P, p_i), resolved once
  int nrow[KN], nslot[KN], pslot = -1, Pslot[NX];
  static_for<NX>([&](auto I) {
    constexpr int ii = I;
    if (i == ii) {
      static_for<KN>([&](auto Kk) { nrow[Kk] = Tb::ncol_row(ii, Kk); nslot[Kk] = Tb::ncol_slot(ii, Kk); });
      pslot = Tb::p_slot(ii);
      static_for<NX>([&](auto C) { Pslot[C] = Tb::P_slot(ii, C); });
    }
  });
  auto rd = [&](const double* rec, int slot) -> double { return slot >= 0 ? rec[(size_t)slot * Bs] : 0.0; };

  // terminal condition: J = P_H + reg_state I (+ the constant cross term of the passive states), j = p_H
  double Jr[NX], jv;
  {
    const double* rec = lin_b + (size_t)(H - 1) * L::stride * Bs;
    jv = rd(rec, pslot);
#pragma unroll
    for (int c = 0; c < NX; ++c) Jr[c] = rd(rec, Pslot[c]) + (c == i ? reg_state : 0.0);
  }
  double p_cross = 0.0;
  if constexpr (M::kPassivePRow >= 0) {
    p_cross = M::passive_P(m, load_weights(m, b, B));
    if (i == M::kPassivePRow) Jr[M::kPassivePCol] += p_cross;
  }
  // mirror the upper triangle into the full rows
  auto mirror = [&]() {
#pragma unroll
    for (int c = 0; c < NX; ++c) tile[i * TS + c] = Jr[c];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      const double v = tile[c * TS + i];
      Jr[c] = c < i ? v : Jr[c];
    }
    __syncwarp();
  };
  mirror();

  bool any_pos = false;
  double jcst = 0.0;
  for (int t = H - 1; t >= 0; --t) {
    const double* rec = lin_b + (size_t)t * L::stride * Bs;
    // ---- operands of the step (every lane reads the whole N, B: the 8 lanes of a candidate hit the same addresses)
    double N[NX][NX], Bm[NX][NU], q[NU], Qd[NU];
    static_for<NX>([&](auto R_) {
      constexpr int r = R_;
      static_for<NX>([&](auto C) {
        constexpr int c = C;
        if constexpr (M::Nm(r, c)) N[r][c] = rec[(size_t)(L::oN + M::Nm.slot(r, c)) * Bs];
      });
      static_for<NU>([&](auto U) {
        constexpr int u = U;
        if constexpr (M::Bm(r, u)) Bm[r][u] = rec[(size_t)(L::oB + M::Bm.slot(r, u)) * Bs];
      });
    });
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      q[u] = rec[(size_t)(L::oq + u) * Bs];
      Qd[u] = rec[(size_t)(L::oQ + u) * Bs] + reg;
    }
    double ncol[KN];  // column i of N
#pragma unroll
    for (int k = 0; k < KN; ++k) ncol[k] = rd(rec, nslot[k]);

    // ---- phase 1: row i of W = J N and of J B
    double W[NX], JB[NU];
    static_for<NX>([&](auto C) {
      constexpr int c = C;
      double acc = 0.0;
      if constexpr (Pat::Nm.col_any(c)) {
        static_for<NX>([&](auto R_) {
          constexpr int r = R_;
          if constexpr (Pat::Nm(r, c)) acc += Jr[r] * N[r][c];
        });
      }
      W[c] = acc;
      tile[i * TS + c] = acc;
    });
    static_for<NU>([&](auto U) {
      constexpr int u = U;
      double acc = 0.0;
      static_for<NX>([&](auto R_) {
        constexpr int r = R_;
        if constexpr (Pat::Bm(r, u)) acc += Jr[r] * Bm[r][u];
      });
      JB[u] = acc;
      JBs[i * NU + u] = acc;
    });
    js[i] = jv;
    __syncwarp();
    // everything the lane needs from the other rows
    double jall[NX], JBall[NX][NU], Wcol[NX];
#pragma unroll
    for (int r = 0; r < NX; ++r) {
      jall[r] = js[r];
      Wcol[r] = tile[r * TS + i];
#pragma unroll
      for (int u = 0; u < NU; ++u) JBall[r][u] = JBs[r * NU + u];
    }
    // H = Q + B^T J B (upper computed, mirrored) ; gq = q + B^T j   (redundantly on every lane)
    double Hm[NU][NU], gq[NU];
    static_for<NU>([&](auto U) {
      constexpr int u = U;
      static_for<NU>([&](auto V) {
        constexpr int v = V;
        if constexpr (v >= u) {
          double acc = 0.0;
          if constexpr (u == v) acc = Qd[u];
          static_for<NX>([&](auto R_) {
            constexpr int r = R_;
            if constexpr (Pat::Bm(r, u)) acc += Bm[r][u] * JBall[r][v];
          });
          Hm[u][v] = acc;
        }
      });
      double acc = q[u];
      static_for<NX>([&](auto R_) {
        constexpr int r = R_;
        if constexpr (Pat::Bm(r, u)) acc += Bm[r][u] * jall[r];
      });
      gq[u] = acc;
    });
#pragma unroll
    for (int u = 0; u < NU; ++u)
#pragma unroll
      for (int v = 0; v < u; ++v) Hm[u][v] = Hm[v][u];
    // column i of G = B^T (J + W)   (B^T J = (J B)^T)
    double G[NU];
    static_for<NU>([&](auto U) {
      constexpr int u = U;
      double acc = JB[u];
      static_for<NX>([&](auto R_) {
        constexpr int r = R_;
        if constexpr (Pat::Bm(r, u)) acc += Bm[r][u] * Wcol[r];
      });
      G[u] = acc;
    });
    // jt = j + N^T j ; row i of J += W + W^T + N^T W (column i of N from the per-lane table, ascending rows)
    double jt = jv;
#pragma unroll
    for (int k = 0; k < KN; ++k) {
      const int r = nrow[k] < 0 ? 0 : nrow[k];
      jt += ncol[k] * js[r];
    }
    {
      double accJ[NX];
#pragma unroll
      for (int c = 0; c < NX; ++c) accJ[c] = (Jr[c] + W[c]) + Wcol[c];
#pragma unroll
      for (int k = 0; k < KN; ++k) {
        const int r = nrow[k] < 0 ? 0 : nrow[k];
#pragma unroll
        for (int c = 0; c < NX; ++c) accJ[c] += ncol[k] * tile[r * TS + c];
      }
#pragma unroll
      for (int c = 0; c < NX; ++c) Jr[c] = accJ[c];
    }

    // ---- phase 2: state cost on x_t (zero at t = 0), K = -H^{-1} G (own column), k = -H^{-1} gq
    double pv = 0.0, Prow[NX];
    if (t > 0) {
      const double* recc = lin_b + (size_t)(t - 1) * L::stride * Bs;
      pv = rd(recc, pslot);
#pragma unroll
      for (int c = 0; c < NX; ++c) Prow[c] = rd(recc, Pslot[c]);
    } else {
#pragma unroll
      for (int c = 0; c < NX; ++c) Prow[c] = 0.0;
    }
#pragma unroll
    for (int c = 0; c < NX; ++c) Prow[c] += (c == i ? reg_state : 0.0);
    double rhs[NU][2];
#pragma unroll
    for (int u = 0; u < NU; ++u) { rhs[u][0] = G[u]; rhs[u][1] = gq[u]; }
    double ipiv[NU];
#pragma unroll
    for (int c = 0; c < NU; ++c) {
#pragma unroll
      for (int r = c + 1; r < NU; ++r) {
        const bool sw = ::fabs(Hm[r][c]) > ::fabs(Hm[c][c]);
#pragma unroll
        for (int cc = 0; cc < NU; ++cc) {
          const double x = Hm[c][cc], y = Hm[r][cc];
          Hm[c][cc] = sw ? y : x;
          Hm[r][cc] = sw ? x : y;
        }
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const double x = rhs[c][cc], y = rhs[r][cc];
          rhs[c][cc] = sw ? y : x;
          rhs[r][cc] = sw ? x : y;
        }
      }
      ipiv[c] = 1.0 / Hm[c][c];
#pragma unroll
      for (int r = c + 1; r < NU; ++r) {
        const double f = Hm[r][c] * ipiv[c];
#pragma unroll
        for (int cc = c + 1; cc < NU; ++cc) Hm[r][cc] -= f * Hm[c][cc];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) rhs[r][cc] -= f * rhs[c][cc];
      }
    }
#pragma unroll
    for (int r = NU - 1; r >= 0; --r) {
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        double acc = rhs[r][cc];
#pragma unroll
        for (int c2 = r + 1; c2 < NU; ++c2) acc -= Hm[r][c2] * rhs[c2][cc];
        rhs[r][cc] = acc * ipiv[r];
      }
    }
    double Kc[NU], kk[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) { Kc[u] = -rhs[u][0]; kk[u] = -rhs[u][1]; Ks[u * NX + i] = Kc[u]; }
    double rest = 0.0;
#pragma unroll
    for (int u = 0; u < NU; ++u) rest += gq[u] * kk[u];
    rest *= 0.5;
    // j = jt + p + G^T k
    {
      double acc = jt;
      if (pslot >= 0) acc += pv;
#pragma unroll
      for (int u = 0; u < NU; ++u) acc += G[u] * kk[u];
      jv = acc;
    }
    __syncwarp();
    // row i of J += P + G^T K
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      double acc = Jr[c];
      if (Pslot[c] >= 0 || c == i) acc += Prow[c];   // (stored pattern + the diagonal, which carries reg_state: riccati.cuh phase2)
#pragma unroll
      for (int u = 0; u < NU; ++u) acc += G[u] * Ks[u * NX + c];
      Jr[c] = acc;
    }
    if constexpr (M::kPassivePRow >= 0) {
      if (t > 0 && i == M::kPassivePRow) Jr[M::kPassivePCol] += p_cross;
    }
    any_pos |= rest > 0.0;
    jcst = jcst + rest;
    if (live) {
      double* g = gains + (size_t)t * L::gain_stride * n_oslots + os;
#pragma unroll
      for (int u = 0; u < NU; ++u) g[(size_t)(u * NX + i) * n_oslots] = Kc[u];
#pragma unroll
      for (int u = 0; u < NU; ++u)
        if (i == u) g[(size_t)(NU * NX + u) * n_oslots] = kk[u];
    }
    __syncwarp();  // (the tile and the K buffer are rewritten by the mirror / the next step)
    mirror();
  }
  if (live && i == 0) {
    jcst_out[os] = jcst;
    if (feas_out) {
      int feas = 1;
      if (bad_dir == ILQC_BAD_DIR_FLAG) feas = any_pos ? 0 : 1;
      feas_out[os] = feas;
    }
  }
}

#endif  // CUDA

}  // namespace ilqc

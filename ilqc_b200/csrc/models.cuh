// Problem definitions: continuous dynamics, integrators, costs and the compile-time sparsity patterns of
// their expansions.  Written once, templated on the scalar type S (double or Jet<N,ORD>).
//
// Reference semantics restated here (file:line in /root/reference):
//   Pendulum.dyn/costs            envs/pendulum.py:56-87
//   CartPendulum.dyn/costs        envs/pendulum.py:180-238   (state ordering quirk :181 vs :197 kept)
//   Car.simple_model              envs/car.py:134-146
//   Car.bicycle_model             envs/car.py:148-188 + envs/bicycle_model.json
//   Car.costs / Car.barrier       envs/car.py:190-302        (detached pieces :262, :267 kept)
//   euler / runge_kutta4_cst_ctrl envs/discretization.py:8-19, :42-57
//   NaturalCubicSpline            envs/tracks/torch_spline.py:341-375
#pragma once
#include "common.cuh"
#include "jet.cuh"
#include "../../include/ilqc_b200.h"

namespace ilqc {

// ---------------------------------------------------------------------------------------------------
// compile-time masks written as strings, one row per line ('x' = structurally non-zero)
// ---------------------------------------------------------------------------------------------------
template <int R, int C>
struct Mask {
  bool m[R][C];
  constexpr Mask() : m{} {}
  constexpr bool operator()(int i, int j) const { return m[i][j]; }
  constexpr int count() const {
    int n = 0;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < C; ++j) n += m[i][j] ? 1 : 0;
    return n;
  }
  // rank of entry (i,j) among the set entries, row-major
  constexpr int slot(int i, int j) const {
    int n = 0;
    for (int a = 0; a < R; ++a)
      for (int b = 0; b < C; ++b) {
        if (a == i && b == j) return n;
        n += m[a][b] ? 1 : 0;
      }
    return n;
  }
  constexpr bool col_any(int j) const {
    for (int i = 0; i < R; ++i)
      if (m[i][j]) return true;
    return false;
  }
  constexpr bool row_any(int i) const {
    for (int j = 0; j < C; ++j)
      if (m[i][j]) return true;
    return false;
  }
};
template <int R, int C>
constexpr Mask<R, C> parse_mask(const char* s) {
  Mask<R, C> k;
  int i = 0, j = 0;
  for (const char* p = s; *p; ++p) {
    if (*p == '/' || *p == ' ') {
      if (*p == '/') { ++i; j = 0; }
      continue;
    }
    k.m[i][j] = (*p == 'x');
    ++j;
  }
  return k;
}
template <int R, int C>
constexpr Mask<R, C> full_mask() {
  Mask<R, C> k;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) k.m[i][j] = true;
  return k;
}
// symmetric masks keep only i<=j set (upper triangle); query with ordered indices
template <int R>
constexpr Mask<R, R> upper_of(Mask<R, R> a) {
  Mask<R, R> k;
  for (int i = 0; i < R; ++i)
    for (int j = i; j < R; ++j) k.m[i][j] = a.m[i][j] || a.m[j][i];
  return k;
}
template <int R>
constexpr Mask<R, R> outer_sym(const int (&v)[R]) {  // v[i]!=0 marks members of the block
  Mask<R, R> k;
  for (int i = 0; i < R; ++i)
    for (int j = i; j < R; ++j) k.m[i][j] = v[i] && v[j];
  return k;
}
template <int R, int C>
constexpr Mask<R, C> mask_or(Mask<R, C> a, Mask<R, C> b) {
  Mask<R, C> k;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) k.m[i][j] = a.m[i][j] || b.m[i][j];
  return k;
}

// per-problem cost weights (Car seed!=0 perturbs them, car.py:49-56)
struct Weights {
  double reg_cont, reg_lag, reg_speed;
};
ILQC_HD Weights load_weights(const ilqc_model_desc& m, int b, int B) {
  Weights w;
  if (m.pp_weights) {
    w.reg_cont = ILQC_LDG(m.pp_weights + b);
    w.reg_lag = ILQC_LDG(m.pp_weights + B + b);
    w.reg_speed = ILQC_LDG(m.pp_weights + 2 * B + b);
  } else {
    w.reg_cont = m.reg_cont; w.reg_lag = m.reg_lag; w.reg_speed = m.reg_speed;
  }
  return w;
}
ILQC_HD int load_t0(const ilqc_model_desc& m, int b) { return m.pp_t0 ? ILQC_LDG(m.pp_t0 + b) : m.t0; }

// ---------------------------------------------------------------------------------------------------
// integrators (discretization.py).  f(const S* x, S* dx) evaluates the continuous dynamics with the
// controls captured; N = number of integrated components.
// ---------------------------------------------------------------------------------------------------
// VAR = runge_kutta4 with three control sets per step (discretization.py:22-39): set 0 for k1, set 1 for k2
// and k3, set 2 for k4.  f(set, s, k) evaluates the continuous dynamics at stage state s with control set `set`.
template <class S, int N, bool VAR, class F>
ILQC_HD void integrate(int integ, double dt, F&& f, const S (&x)[N], S (&xn)[N]) {
  if (!VAR && integ == ILQC_INT_EULER) {
    S k[N];
    f(0, x, k);
#pragma unroll
    for (int i = 0; i < N; ++i) xn[i] = x[i] + dt * k[i];
  } else {
    // state + dt*(k1 + 2*k2 + 2*k3 + k4)/6 with stage states state + dt*k1/2, state + dt*k2/2, state + dt*k3.
    // Written as a loop over the four stages so that the (large, transcendental-heavy) dynamics body can be
    // instantiated once: 4x less code than the unrolled form.  Bit-identical to the unrolled formulas:
    // x/2 == x*0.5 and 1*k, 2*k are exact.  Measured on B200 (profiles/r01_tuning_kernels.txt): rolled is
    // 9-12 % faster for the plain roll-out and cuts the spills of the second-order jets by a third, but is 5 %
    // slower for the first-order expansion, which keeps the unrolled form (as do the three-control-set
    // integrators, whose per-stage control set must be a compile-time index).
#ifndef ILQC_STAGE_UNROLL_JET1
#define ILQC_STAGE_UNROLL_JET1 4
#endif
#ifndef ILQC_STAGE_UNROLL_MAX_DOUBLES
#define ILQC_STAGE_UNROLL_MAX_DOUBLES 16
#endif
    constexpr int kStageUnroll =
        VAR ? 4 : ((is_jet<S>::value && sizeof(S) <= sizeof(double) * ILQC_STAGE_UNROLL_MAX_DOUBLES) ? ILQC_STAGE_UNROLL_JET1 : 1);
    S k[N], s[N], acc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = x[i];
#pragma unroll(kStageUnroll)
    for (int stage = 0; stage < 4; ++stage) {
      const int set = VAR ? (stage == 0 ? 0 : (stage == 3 ? 2 : 1)) : 0;
      f(set, s, k);
      const double wgt = (stage == 1 || stage == 2) ? 2.0 : 1.0;  // k1 + 2 k2 + 2 k3 + k4
      const double half = (stage < 2) ? 0.5 : 1.0;                 // next stage state: dt*k/2, dt*k/2, dt*k
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (stage == 0) acc[i] = k[i]; else acc[i] = acc[i] + wgt * k[i];
        s[i] = x[i] + (dt * k[i]) * half;
      }
    }
    MathHot mh;  // (dt * acc) / 6: quotient by a constant, correctly rounded in 3 FMAs (fastmath.cuh)
#pragma unroll
    for (int i = 0; i < N; ++i) xn[i] = x[i] + mh.divc(dt * acc[i], 6.0, 1.0 / 6.0);
  }
}

// ---------------------------------------------------------------------------------------------------
// natural cubic spline lookup (torch_spline.py:341-375).  One segment search serves the three splines
// of a track (shared knots, tracks.py:50-53).
// ---------------------------------------------------------------------------------------------------
struct Seg {
  const double* c;  // 24 coefficients of the segment: [spline][channel][a,b,c,d]
};
// smooth_relu with eps = 1e-6 (torch_utils.py:44-69): x for x > eps, 0 for x < -eps, a cubic blend in between
// (the reference multiplies the branches by boolean masks; the inactive branch contributes an exact 0)
template <class S>
ILQC_HD S smooth_relu(const S& x) {
  const double eps = 1e-6;
  const double xv = val(x);
  if (xv > eps) return x;
  if (xv >= -eps) {
    const double x1 = -eps, x2 = eps, y1 = 0.0, y2 = eps;
    const double a = 0.0 * (x2 - x1) - (y2 - y1), b = -1.0 * (x2 - x1) + (y2 - y1);
    const S tx = (x - x1) / (x2 - x1);
    const S omt = 1.0 - tx;
    return omt * y1 + tx * y2 + (tx * omt) * (omt * a + tx * b);
  }
  return make_cst<S>(0.0);
}
template <class S>
ILQC_HD Seg spline_locate(const ilqc_model_desc& m, const S& t, S* frac) {
  double tv = val(t);
  S tw = t;
  if (m.termination == ILQC_TERM_LOOP) {
    // torch.remainder: result has the sign of the divisor; derivative 1
    const double L = m.track_len;
    double r = ::fmod(tv, L);
    if (r != 0.0 && ((r < 0.0) != (L < 0.0))) r += L;
    set_val(tw, r);  // shifts the value, keeps the derivatives
    tv = r;
  } else {
    // 'repeat_end_point': t <- smooth_min(t, L) = -smooth_relu(-t + L) + L (torch_utils.py:72, torch_spline.py:345)
    const double L = m.track_len;
    tw = -smooth_relu<S>(-t + L) + L;
    tv = val(tw);
  }
  // bucketize(t, knots) (right=False) - 1, clamped: the last knot strictly below t, i.e. the segment with
  // knots[idx] < t <= knots[idx+1] (t <= knots[0] -> 0, t > knots[n_seg] -> n_seg-1).  The knots are arc
  // lengths of nearly equidistant way-points, so a proportional guess is off by a segment or two at most;
  // two short walks fix it with 2-3 table reads instead of the 9 dependent reads of a binary search.
  const int last = m.n_seg - 1;
  int idx = (int)(tv * ((double)m.n_seg / m.track_len));
  idx = idx < 0 ? 0 : (idx > last ? last : idx);
  while (idx > 0 && !(ILQC_LDG(m.knots + idx) < tv)) --idx;
  while (idx < last && ILQC_LDG(m.knots + idx + 1) < tv) ++idx;
  *frac = tw - ILQC_LDG(m.knots + idx);
  Seg s;
  s.c = m.coef + (size_t)idx * 24;
  return s;
}
template <class S>
ILQC_HD S spline_eval(const Seg& s, int spline, int ch, const S& f) {
  const double* c = s.c + (spline * 2 + ch) * 4;
  S inner = ILQC_LDG(c + 2) + ILQC_LDG(c + 3) * f;
  inner = ILQC_LDG(c + 1) + inner * f;
  return ILQC_LDG(c + 0) + inner * f;
}
template <class S>
ILQC_HD S spline_deriv(const Seg& s, int spline, int ch, const S& f) {
  const double* c = s.c + (spline * 2 + ch) * 4;
  S inner = 2.0 * ILQC_LDG(c + 2) + (3.0 * ILQC_LDG(c + 3)) * f;
  return ILQC_LDG(c + 1) + inner * f;
}

// ===================================================================================================
// Model traits.  Each model declares
//   NX, NU            state / control dimension
//   NV                number of jet variables of the dynamics; SV[i] / UV[k] = jet index of state i /
//                     control k (-1: the dynamics depend on it only through constant-coefficient terms)
//   NVC, CV[i]        same for the state cost
//   masks             N = A - I, B, p, P (upper), and for the second-order path R, Hxx, Huu
//   step<S>()         one discrete step ; state_cost<S>() ; ctrl_cost()
// ===================================================================================================
// value, gradient and packed Hessian of M::state_cost wrt its NVC cost variables through second-order jets
template <class M>
ILQC_HD void jet_cost_expansion(const ilqc_model_desc& m, const Weights& w, const double (&x)[M::NX], int time_iter,
                                double* cval, double (&g)[M::NVC], double (&h)[sym_size(M::NVC)]) {
  using SC = Jet<M::NVC, 2>;
  SC xc[M::NX];
  static_for<M::NX>([&](auto I) { xc[I] = make_var<SC>(x[I], M::cv(I)); });
  const SC c = M::template state_cost<SC>(m, w, xc, time_iter);
  *cval = c.v;
#pragma unroll
  for (int i = 0; i < M::NVC; ++i) g[i] = c.g[i];
#pragma unroll
  for (int i = 0; i < sym_size(M::NVC); ++i) h[i] = c.h[i];
}

template <int MODEL>
struct Model;
// internal ids of the three-control-set (`rk4`) variants: public model id + ILQC_MODEL_VAR_OFFSET
#define ILQC_MODEL_VAR_OFFSET 4

// ---------------------------------------------------------------------------------------------------
template <>
struct Model<ILQC_MODEL_PENDULUM> {
  static constexpr int NX = 2, NU = 1, NV = 3, NVC = 2;
  static constexpr int NXH = NX;  // state rows whose dynamics can have non-zero second derivatives
  static constexpr int kRotVar = -1;  // no heading variable (see CarSimpleT)
  static constexpr int sv(int i) { constexpr int t[NX] = {0, 1}; return t[i]; }
  static constexpr int uv(int i) { constexpr int t[NU] = {2}; return t[i]; }
  static constexpr int cv(int i) { constexpr int t[NX] = {0, 1}; return t[i]; }
  static constexpr auto Nm = full_mask<NX, NX>();
  static constexpr auto Bm = full_mask<NX, NU>();
  static constexpr auto pm = full_mask<1, NX>();
  static constexpr auto Pm = upper_of(full_mask<NX, NX>());
  static ILQC_HD double passive_N(const ilqc_model_desc&, int, int) { return 0.0; }
  static ILQC_HD double passive_B(const ilqc_model_desc&, int, int) { return 0.0; }
  static constexpr int kPassivePRow = -1, kPassivePCol = -1;  // no constant cost-Hessian entry outside Pm
  static ILQC_HD double passive_P(const ilqc_model_desc&, const Weights&) { return 0.0; }

  // dyn (pendulum.py:56-66) + euler
  template <class S>
  static ILQC_HD void step(const ilqc_model_desc& m, const S (&x)[NX], const S (&u)[NU], S (&xn)[NX]) {
    const double g = m.phys[0], mm = m.phys[1], l = m.phys[2], mu = m.phys[3];
    const double c_sin = -g / l, c_fr = mu / (mm * (l * l)), c_u = 1.0 / (mm * (l * l));
    auto f = [&](int, const S* s, S* d) {
      const S a[1] = {s[0] + M_PI};
      S sn[1], cs[1];
      MathHot mx;
      mx.sincos(a, sn, cs);
      if constexpr (MathHot::kFast) {
        if (ILQC_UNLIKELY(mx.bad)) sn[0] = jsin(a[0]);  // |theta| >= 1e5: CUDA math library
      }
      d[0] = s[1];
      d[1] = c_sin * sn[0] - c_fr * s[1] + c_u * u[0];
    };
    integrate<S, NX, false>(m.integrator, m.dt, f, x, xn);
  }
  // cost_state (pendulum.py:82-87): time_iter is the index of the state the cost is charged on
  template <class S>
  static ILQC_HD S state_cost(const ilqc_model_desc& m, const Weights&, const S (&x)[NX], int time_iter) {
    if (time_iter >= m.stay_put_start) return jsq(x[0]) + m.reg_speed * jsq(x[1]);
    return make_cst<S>(0.0);
  }
  static ILQC_HD void cost_expansion(const ilqc_model_desc& m, const Weights& w, const double (&x)[NX], int time_iter,
                                     double* cval, double (&g)[NVC], double (&h)[sym_size(NVC)]) {
    jet_cost_expansion<Model<ILQC_MODEL_PENDULUM>>(m, w, x, time_iter, cval, g, h);
  }
  static ILQC_HD double ctrl_cost(const ilqc_model_desc& m, const double (&u)[NU], double (&q)[NU],
                                  double (&Qd)[NU]) {
    q[0] = 2.0 * m.reg_ctrl * u[0];
    Qd[0] = 2.0 * m.reg_ctrl;
    return m.reg_ctrl * (u[0] * u[0]);
  }
};

// ---------------------------------------------------------------------------------------------------
// CartPendulum (pendulum.py:115-238).  VAR: discretization='rk4', one control per control set (nu = 3).
template <bool VAR>
struct CartPoleT {
  static constexpr int NSET = VAR ? 3 : 1;
  static constexpr int NX = 4, NU = NSET, NV = 3 + NSET, NVC = 3;
  static constexpr int NXH = NX;
  static constexpr int kRotVar = -1;
  static constexpr int sv(int i) { constexpr int t[NX] = {-1, 0, 1, 2}; return t[i]; }
  static constexpr int uv(int i) { return 3 + i; }
  static constexpr int cv(int i) { constexpr int t[NX] = {0, 1, -1, 2}; return t[i]; }
  static constexpr auto Nm = parse_mask<NX, NX>(".xxx/.xxx/.xxx/.xxx");
  static constexpr auto Bm = full_mask<NX, NU>();
  static constexpr auto pm = parse_mask<1, NX>("xx.x");
  static constexpr auto Pm = parse_mask<NX, NX>("x.../.x../..../...x");
  static ILQC_HD double passive_N(const ilqc_model_desc&, int, int) { return 0.0; }
  static ILQC_HD double passive_B(const ilqc_model_desc&, int, int) { return 0.0; }
  static constexpr int kPassivePRow = -1, kPassivePCol = -1;
  static ILQC_HD double passive_P(const ilqc_model_desc&, const Weights&) { return 0.0; }

  // dyn (pendulum.py:180-198), literal ordering: unpack (x, xdot, th, thdot) = s0..s3, return
  // (xdot, thdot, dxdot, dthdot)
  template <class S>
  static ILQC_HD void step(const ilqc_model_desc& m, const S (&x)[NX], const S (&u)[NU], S (&xn)[NX]) {
    const double g = m.phys[0], M = m.phys[1], mm = m.phys[2], b = m.phys[3], I = m.phys[4], l = m.phys[5];
    const double a11 = M + mm, a22 = I + mm * (l * l), ml = mm * l;
    const double d0 = I * (M + mm) + mm * (l * l) * M, d1 = (mm * mm) * (l * l);
    const double mgl = -mm * g * l;
    auto dyn = [&](auto& mx, int set, const S* s, S* d) {
      const S xdot = s[1], thdot = s[3];
      const S th[1] = {s[2]};
      S sn[1], cs[1];
      mx.sincos(th, sn, cs);
      const S sth = sn[0], cth = cs[0];
      const S a12 = ml * cth;
      const S detA = d0 + d1 * jsq(sth);
      const S b1 = (ml * jsq(thdot)) * sth - b * xdot + u[set];
      const S b2 = mgl * sth;
      d[0] = xdot;
      d[1] = thdot;
      d[2] = mx.div(a22 * b1 - a12 * b2, detA);
      d[3] = mx.div(a11 * b2 - a12 * b1, detA);
    };
    auto f = [&](int set, const S* s, S* d) {
      MathHot mx;
      dyn(mx, set, s, d);
      if constexpr (MathHot::kFast) {
        if (ILQC_UNLIKELY(mx.bad)) {  // an argument outside the fast functions' range: CUDA math library
          MathExact me;
          dyn(me, set, s, d);
        }
      }
    };
    integrate<S, NX, VAR>(m.integrator, m.dt, f, x, xn);
  }
  // cost_state (pendulum.py:225-238)
  template <class S>
  static ILQC_HD S state_cost(const ilqc_model_desc& m, const Weights&, const S (&x)[NX], int time_iter) {
    S c = make_cst<S>(0.0);
    if (time_iter >= m.stay_put_start) c = jsq(x[1] + M_PI) + m.reg_speed * jsq(x[3]);
    if (m.has_x_limits) {
      const S bar = m.reg_barrier * (jcube(jmax0(-(x[0] - m.x_min))) + jcube(jmax0(-(m.x_max - x[0]))));
      c = c + bar;
    }
    return c;
  }
  static ILQC_HD void cost_expansion(const ilqc_model_desc& m, const Weights& w, const double (&x)[NX], int time_iter,
                                     double* cval, double (&g)[NVC], double (&h)[sym_size(NVC)]) {
    jet_cost_expansion<CartPoleT<VAR>>(m, w, x, time_iter, cval, g, h);
  }
  // reg_ctrl * ctrl.dot(ctrl) (pendulum.py:217-222)
  static ILQC_HD double ctrl_cost(const ilqc_model_desc& m, const double (&u)[NU], double (&q)[NU],
                                  double (&Qd)[NU]) {
    double dot = 0.0;
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      dot += u[i] * u[i];
      q[i] = 2.0 * m.reg_ctrl * u[i];
      Qd[i] = 2.0 * m.reg_ctrl;
    }
    return m.reg_ctrl * dot;
  }
};
template <> struct Model<ILQC_MODEL_CARTPOLE> : CartPoleT<false> {};
template <> struct Model<ILQC_MODEL_CARTPOLE + ILQC_MODEL_VAR_OFFSET> : CartPoleT<true> {};

// ---------------------------------------------------------------------------------------------------
// Car: shared cost code (car.py:190-302).
// ---------------------------------------------------------------------------------------------------
// Contouring cost (car.py:196-227) + barrier (car.py:255-302) as a function of the reference time `tref`, which
// is the only argument that enters through the splines: S = double (roll-outs) or Jet<1,2> seeded on tref (cost
// expansion).  x, y enter through ex = x - xr(tref), ey = y - yr(tref) only (the barrier's `point` is detached,
// car.py:262) and vtref through separable terms, so their derivatives are closed forms of the quantities
// returned in `aux` (car_cost_expansion below) and a 4-variable second-order jet is never needed.
template <class S>
struct CarCostAux {
  S ec, el, sp, cp;  // contouring / lag error and sin / cos of the track heading, as functions of tref
};
template <class S, class MX>
ILQC_HD S car_contouring_cost(MX& mx, const ilqc_model_desc& m, const Weights& w, double px, double py, const S& tref,
                              double vtref, CarCostAux<S>* aux) {
  S fr;
  const Seg sg = spline_locate<S>(m, tref, &fr);
  const S xr = spline_eval<S>(sg, 0, 0, fr), yr = spline_eval<S>(sg, 0, 1, fr);
  const S dxr = spline_deriv<S>(sg, 0, 0, fr), dyr = spline_deriv<S>(sg, 0, 1, fr);
  S sp, cp;
  if constexpr (MX::kFast) {
    // sin / cos of phi_ref = atan2(dy, dx) are dy / |d| and dx / |d|: one reciprocal square root instead of
    // atan2 + sin + cos (equal to the reference's composition up to rounding)
    S nrm, inrm;
    mx.sqrt_inv(jsq(dxr) + jsq(dyr), &nrm, &inrm);
    sp = dyr * inrm;
    cp = dxr * inrm;
  } else {
    const S phi_ref = jatan2(dyr, dxr);
    sp = jsin(phi_ref);
    cp = jcos(phi_ref);
  }
  const S ex = px - xr, ey = py - yr;
  const S cont_err = sp * ex - cp * ey;
  const S lag_err = -cp * ex - sp * ey;
  if (aux) { aux->ec = cont_err; aux->el = lag_err; aux->sp = sp; aux->cp = cp; }
  S c = w.reg_cont * jsq(cont_err) + w.reg_lag * jsq(lag_err);
  if (m.time_penalty == ILQC_TP_VREF_SQUARED) c = c + (w.reg_speed * jsq(vtref - m.vref)) * (m.dt * m.dt);
  else if (m.time_penalty == ILQC_TP_TREF) c = c - w.reg_speed * tref;
  else if (m.time_penalty == ILQC_TP_DREF) c = c - (w.reg_speed * vtref) * m.dt;
  else c = c - ((w.reg_speed * vtref) * m.dt) * tref;
  // barrier (car.py:255-302)
  S bar = make_cst<S>(-1e-6 * ::log(vtref));
  const double carwidth = 1.5 * m.car_w;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const S bx = spline_eval<S>(sg, 1 + i, 0, fr), by = spline_eval<S>(sg, 1 + i, 1, fr);
    const S dbx = spline_deriv<S>(sg, 1 + i, 0, fr), dby = spline_deriv<S>(sg, 1 + i, 1, fr);
    // normal = torch.tensor([dpos[1], -dpos[0]]) / v : numerator detached (car.py:267)
    S n0, n1;
    if constexpr (MX::kFast) {
      S v, iv;
      mx.sqrt_inv(jsq(dbx) + jsq(dby), &v, &iv);
      n0 = val(dby) * iv;
      n1 = (-val(dbx)) * iv;
    } else {
      const S v = jsqrt(jsq(dbx) + jsq(dby));
      n0 = val(dby) / v;
      n1 = (-val(dbx)) / v;
    }
    const S dot = (px - bx) * n0 + (py - by) * n1;
    if (i == 0) bar = bar + m.reg_bar * jcube(jmax0(carwidth / 2.0 - dot));
    else bar = bar + m.reg_bar * jcube(jmax0(dot + carwidth / 2.0));
  }
  if (m.reg_obs > 0.0) {
    const double d2 = (px - m.obs_x) * (px - m.obs_x) + (py - m.obs_y) * (py - m.obs_y);
    const double o = jmax0(carwidth * carwidth - d2);
    bar = bar + m.reg_obs * ((o * o) * o);
  }
  return c + bar;
}
template <class S, class MX>
ILQC_HD S car_contouring_cost_checked(const ilqc_model_desc& m, const Weights& w, double px, double py, const S& tref,
                                      double vtref, CarCostAux<S>* aux) {
  MathHot mx;
  S c = car_contouring_cost<S>(mx, m, w, px, py, tref, vtref, aux);
  if constexpr (MathHot::kFast) {
    if (ILQC_UNLIKELY(mx.bad)) {  // an argument outside the fast functions' range: CUDA math library
      MathExact me;
      c = car_contouring_cost<S>(me, m, w, px, py, tref, vtref, aux);
    }
  }
  return c;
}
// reference point of the `exact` cost: time = time_iter * vref * dt is a constant of the graph (car.py:229-231)
ILQC_HD void car_exact_ref(const ilqc_model_desc& m, int time_iter, double* xr, double* yr) {
  const double time = ((double)time_iter * m.vref) * m.dt;
  double fr;
  const Seg sg = spline_locate<double>(m, time, &fr);
  *xr = spline_eval<double>(sg, 0, 0, fr);
  *yr = spline_eval<double>(sg, 0, 1, fr);
}
// value of the state cost (roll-outs)
template <class S, int NX>
ILQC_HD S car_state_cost(const ilqc_model_desc& m, const Weights& w, const S (&x)[NX], int time_iter) {
  static_assert(!is_jet<S>::value, "derivatives of the car cost come from car_cost_expansion");
  if (m.cost == ILQC_COST_EXACT) {
    double xr, yr;
    car_exact_ref(m, time_iter, &xr, &yr);
    return jsq(x[0] - xr) + jsq(x[1] - yr);
  }
  return car_contouring_cost_checked<double, MathHot>(m, w, x[0], x[1], x[NX - 2], x[NX - 1], nullptr);
}
// value, gradient and Hessian of the state cost wrt the cost variables (x, y, tref, vtref) = what autograd
// returns in the reference (forward.py:274-321): g[4], h = packed upper triangle of the 4 x 4 Hessian
template <int NX>
ILQC_HD void car_cost_expansion(const ilqc_model_desc& m, const Weights& w, const double (&x)[NX], int time_iter,
                                double* cval, double (&g)[4], double (&h)[10]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 10; ++i) h[i] = 0.0;
  const double px = x[0], py = x[1], tref = x[NX - 2], vtref = x[NX - 1];
  if (m.cost == ILQC_COST_EXACT) {
    double xr, yr;
    car_exact_ref(m, time_iter, &xr, &yr);
    const double ex = px - xr, ey = py - yr;
    *cval = ex * ex + ey * ey;
    g[0] = 2.0 * ex;
    g[1] = 2.0 * ey;
    h[sym_idx(0, 0, 4)] = 2.0;
    h[sym_idx(1, 1, 4)] = 2.0;
    return;
  }
  using S1 = Jet<1, 2>;
  CarCostAux<S1> a;
  const S1 c = car_contouring_cost_checked<S1, MathHot>(m, w, px, py, make_var<S1>(tref, 0), vtref, &a);
  *cval = c.v;
  const double qC = w.reg_cont, qL = w.reg_lag, qV = w.reg_speed;
  const double sp = a.sp.v, cp = a.cp.v, dsp = a.sp.g[0], dcp = a.cp.g[0];
  const double ec = a.ec.v, el = a.el.v, dec = a.ec.g[0], del = a.el.g[0];
  // ec = sp ex - cp ey, el = -cp ex - sp ey  =>  d ec/dx = sp, d ec/dy = -cp, d el/dx = -cp, d el/dy = -sp
  g[0] = 2.0 * (qC * (ec * sp) - qL * (el * cp));
  g[1] = -2.0 * (qC * (ec * cp) + qL * (el * sp));
  g[2] = c.g[0];
  h[sym_idx(0, 0, 4)] = 2.0 * (qC * (sp * sp) + qL * (cp * cp));
  h[sym_idx(1, 1, 4)] = 2.0 * (qC * (cp * cp) + qL * (sp * sp));
  h[sym_idx(0, 1, 4)] = 2.0 * ((qL - qC) * (sp * cp));
  h[sym_idx(0, 2, 4)] = 2.0 * (qC * (dec * sp + ec * dsp) - qL * (del * cp + el * dcp));
  h[sym_idx(1, 2, 4)] = -2.0 * (qC * (dec * cp + ec * dcp) + qL * (del * sp + el * dsp));
  h[sym_idx(2, 2, 4)] = c.h[0];
  // vtref: separable terms (time penalty + log barrier)
  const double iv = 1.0 / vtref;
  double gv = -1e-6 * iv, hv = 1e-6 * (iv * iv);
  if (m.time_penalty == ILQC_TP_VREF_SQUARED) {
    gv += (2.0 * (w.reg_speed * (vtref - m.vref))) * (m.dt * m.dt);
    hv += (2.0 * qV) * (m.dt * m.dt);
  } else if (m.time_penalty == ILQC_TP_DREF) {
    gv -= qV * m.dt;
  } else if (m.time_penalty == ILQC_TP_DREF_SQUARED) {
    gv -= (qV * m.dt) * tref;  // (the constant mixed entry -(qV dt) is car_passive_P)
  }
  g[3] = gv;
  h[sym_idx(3, 3, 4)] = hv;
}

// time_penalty='dref_squared' (car.py:213-216): -((reg_speed * vtref) * dt) * tref has the constant mixed second
// derivative -(reg_speed * dt) wrt (tref, vtref), an entry outside the stored P pattern
ILQC_HD double car_passive_P(const ilqc_model_desc& m, const Weights& w) {
  return (m.cost == ILQC_COST_CONTOURING && m.time_penalty == ILQC_TP_DREF_SQUARED) ? -(w.reg_speed * m.dt) : 0.0;
}

// reg_ctrl * ctrl.dot(ctrl), or the tuple form reg[0]*(acc^2 + delta^2) + reg[1]*atref^2 summed over the
// control sets (car.py:235-251: ctrl[::3], ctrl[1::3], ctrl[2::3])
template <int NU>
ILQC_HD double car_ctrl_cost(const ilqc_model_desc& m, const double (&u)[NU], double (&q)[NU], double (&Qd)[NU]) {
  double c0 = 0.0, c1 = 0.0;
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    const bool is_tref = (i % 3 == 2);
    const double r = (m.reg_ctrl_is_tuple && is_tref) ? m.reg_ctrl_tref : m.reg_ctrl;
    if (m.reg_ctrl_is_tuple && is_tref) c1 += u[i] * u[i]; else c0 += u[i] * u[i];
    q[i] = 2.0 * r * u[i];
    Qd[i] = 2.0 * r;
  }
  return m.reg_ctrl_is_tuple ? (m.reg_ctrl * c0 + m.reg_ctrl_tref * c1) : m.reg_ctrl * c0;
}

// reference-time double integrator (tref, vtref ; atref per control set), the same for both car models
// (car.py:143-146, :186-188)
template <bool VAR>
ILQC_HD void car_tref_step(const ilqc_model_desc& m, double tref, double vtref, const double (&atref)[VAR ? 3 : 1],
                           double* trefn, double* vtrefn) {
  const double x[2] = {tref, vtref};
  double xn[2];
  auto f = [&](int set, const double* s, double* d) { d[0] = s[1]; d[1] = atref[set]; };
  integrate<double, 2, VAR>(m.integrator, m.dt, f, x, xn);
  *trefn = xn[0];
  *vtrefn = xn[1];
}
// its constant Jacobian wrt the control set `set`:
//   constant control: d tref'/d atref = dt^2/2 (rk4) or 0 (euler), d vtref'/d atref = dt
//   three sets:       d tref'/d atref_i = (dt^2/6, dt^2/3, 0),      d vtref'/d atref_i = dt/6 * (1, 4, 1)
template <bool VAR>
ILQC_HD double car_tref_B(const ilqc_model_desc& m, int row_is_vtref, int set) {
  const double dt = m.dt;
  if (!VAR) {
    if (row_is_vtref) return dt;
    return m.integrator == ILQC_INT_EULER ? 0.0 : (dt * dt) / 2.0;
  }
  if (row_is_vtref) return (dt / 6.0) * (set == 1 ? 4.0 : 1.0);
  return set == 0 ? (dt * dt) / 6.0 : (set == 1 ? (dt * dt) / 3.0 : 0.0);
}
// B mask of a car model: vehicle rows x control (a, delta) of every set, (tref, vtref) rows x atref of every
// set; ACC_ROW = a row that only the acceleration enters (simple car's v), -1 if none
template <int NX, int NU, int NVEH, int ACC_ROW>
constexpr Mask<NX, NU> car_B_mask() {
  Mask<NX, NU> k;
  for (int i = 0; i < NX; ++i)
    for (int j = 0; j < NU; ++j) {
      if (i < NVEH) k.m[i][j] = (j % 3 != 2) && (i != ACC_ROW || j % 3 == 0);
      else k.m[i][j] = (j % 3 == 2);
    }
  return k;
}

// ---------------------------------------------------------------------------------------------------
template <bool VAR>
struct CarSimpleT {
  // state (x, y, phi, v, tref, vtref), ctrl (a, delta, atref) per control set
  static constexpr int NSET = VAR ? 3 : 1;
  static constexpr int NX = 6, NU = 3 * NSET, NV = 2 + 2 * NSET, NVC = 4;
  static constexpr int NXH = 4;  // (tref, vtref) evolve linearly
  // jet index of the heading phi.  (x' - x, y' - y) = R(phi) w with w independent of phi (phi enters the dynamics only
  // through the rotation of the velocity, and phi' - phi does not depend on phi), so every derivative wrt phi is a
  // rotation of quantities the other variables already give: d(x'-x)/dphi = -(y'-y), d(y'-y)/dphi = x'-x.  The
  // Hessian-tensor kernel therefore runs its second-order jets without phi (hess_body).
  static constexpr int kRotVar = 0;
  static constexpr int sv(int i) { constexpr int t[NX] = {-1, -1, 0, 1, -1, -1}; return t[i]; }
  static constexpr int uv(int k) { return (k % 3 == 2) ? -1 : 2 + 2 * (k / 3) + (k % 3); }
  static constexpr int cv(int i) { constexpr int t[NX] = {0, 1, -1, -1, 2, 3}; return t[i]; }
  static constexpr auto Nm = parse_mask<NX, NX>("..xx../..xx../...x../....../.....x/......");
  static constexpr auto Bm = car_B_mask<NX, NU, 4, 3>();
  static constexpr auto pm = parse_mask<1, NX>("xx..xx");
  static constexpr auto Pm = parse_mask<NX, NX>("xx..x./.x..x./....../....../....x./.....x");

  // steering: delta = 2/pi atan(u) * constrain_angle ; tan(delta) = sin / cos
  template <class S, class MX>
  static ILQC_HD void controls(MX& mx, const ilqc_model_desc& m, const S (&u)[NU], S (&tand)[NSET]) {
    S ud[NSET], at[NSET];
#pragma unroll
    for (int g = 0; g < NSET; ++g) ud[g] = u[3 * g + 1];
    if constexpr (MX::kFast) {
      mx.atan(ud, at);
      S delta[NSET], sd[NSET], cd[NSET];
#pragma unroll
      for (int g = 0; g < NSET; ++g) delta[g] = ((2.0 / M_PI) * at[g]) * m.constrain_angle;
      mx.sincos(delta, sd, cd);
#pragma unroll
      for (int g = 0; g < NSET; ++g) tand[g] = mx.div(sd[g], cd[g]);
    } else {
#pragma unroll
      for (int g = 0; g < NSET; ++g) tand[g] = jtan(((2.0 / M_PI) * jatan(ud[g])) * m.constrain_angle);
    }
  }
  template <class S, class MX>
  static ILQC_HD void dyn(MX& mx, double carlength, double rcarlength, const S& tand, const S& acc, const S* s, S* d) {
    const S ph[1] = {s[2]};
    S sphi[1], cphi[1];
    mx.sincos(ph, sphi, cphi);
    d[0] = s[3] * cphi[0];
    d[1] = s[3] * sphi[0];
    d[2] = mx.divc(s[3] * tand, carlength, rcarlength);
    d[3] = acc;
  }
  template <class S>
  static ILQC_HD void step(const ilqc_model_desc& m, const S (&x)[NX], const S (&u)[NU], S (&xn)[NX]) {
    S tand[NSET];
    double atref[NSET];
#pragma unroll
    for (int g = 0; g < NSET; ++g) atref[g] = val(u[3 * g + 2]);
    {
      MathHot mx;
      controls<S>(mx, m, u, tand);
      if constexpr (MathHot::kFast) {
        if (ILQC_UNLIKELY(mx.bad)) {
          MathExact me;
          controls<S>(me, m, u, tand);
        }
      }
    }
    const double carlength = 1.5 * m.car_l, rcarlength = 1.0 / carlength;
    // vehicle part: (x, y, phi, v)
    const S xv[4] = {x[0], x[1], x[2], x[3]};
    S xvn[4];
    auto f = [&](int set, const S* s, S* d) {
      MathHot mx;
      dyn<S>(mx, carlength, rcarlength, tand[set], u[3 * set], s, d);
      if constexpr (MathHot::kFast) {
        if (ILQC_UNLIKELY(mx.bad)) {
          MathExact me;
          dyn<S>(me, carlength, rcarlength, tand[set], u[3 * set], s, d);
        }
      }
    };
    integrate<S, 4, VAR>(m.integrator, m.dt, f, xv, xvn);
#pragma unroll
    for (int i = 0; i < 4; ++i) xn[i] = xvn[i];
    double tn, vn;
    car_tref_step<VAR>(m, val(x[4]), val(x[5]), atref, &tn, &vn);
    xn[4] = make_cst<S>(tn);
    xn[5] = make_cst<S>(vn);
  }
  // constant-coefficient part of (N, B) that the jets do not carry
  static constexpr int kPassivePRow = 4, kPassivePCol = 5;
  static ILQC_HD double passive_P(const ilqc_model_desc& m, const Weights& w) { return car_passive_P(m, w); }
  static ILQC_HD double passive_N(const ilqc_model_desc& m, int i, int j) { return (i == 4 && j == 5) ? m.dt : 0.0; }
  static ILQC_HD double passive_B(const ilqc_model_desc& m, int i, int k) {
    if (k % 3 != 2 || i < 4) return 0.0;
    return car_tref_B<VAR>(m, i == 5, k / 3);
  }
  template <class S>
  static ILQC_HD S state_cost(const ilqc_model_desc& m, const Weights& w, const S (&x)[NX], int time_iter) {
    return car_state_cost<S, NX>(m, w, x, time_iter);
  }
  static ILQC_HD void cost_expansion(const ilqc_model_desc& m, const Weights& w, const double (&x)[NX], int time_iter,
                                     double* cval, double (&g)[NVC], double (&h)[sym_size(NVC)]) {
    car_cost_expansion<NX>(m, w, x, time_iter, cval, g, h);
  }
  static ILQC_HD double ctrl_cost(const ilqc_model_desc& m, const double (&u)[NU], double (&q)[NU],
                                  double (&Qd)[NU]) {
    return car_ctrl_cost<NU>(m, u, q, Qd);
  }
};
template <> struct Model<ILQC_MODEL_CAR_SIMPLE> : CarSimpleT<false> {};
template <> struct Model<ILQC_MODEL_CAR_SIMPLE + ILQC_MODEL_VAR_OFFSET> : CarSimpleT<true> {};

// ---------------------------------------------------------------------------------------------------
template <bool VAR>
struct CarBicycleT {
  // state (x, y, phi, vx, vy, vphi, tref, vtref), ctrl (a, delta, atref) per control set
  static constexpr int NSET = VAR ? 3 : 1;
  static constexpr int NX = 8, NU = 3 * NSET, NV = 4 + 2 * NSET, NVC = 4;
  static constexpr int NXH = 6;  // (tref, vtref) evolve linearly
  static constexpr int kRotVar = 0;  // heading phi, see CarSimpleT
  static constexpr int sv(int i) { constexpr int t[NX] = {-1, -1, 0, 1, 2, 3, -1, -1}; return t[i]; }
  static constexpr int uv(int k) { return (k % 3 == 2) ? -1 : 4 + 2 * (k / 3) + (k % 3); }
  static constexpr int cv(int i) { constexpr int t[NX] = {0, 1, -1, -1, -1, -1, 2, 3}; return t[i]; }
  static constexpr auto Nm =
      parse_mask<NX, NX>("..xxxx../..xxxx../...xxx../...xxx../...xxx../...xxx../.......x/........");
  static constexpr auto Bm = car_B_mask<NX, NU, 6, -1>();
  static constexpr auto pm = parse_mask<1, NX>("xx....xx");
  static constexpr auto Pm =
      parse_mask<NX, NX>("xx....x./.x....x./......../......../......../......../......x./.......x");

  // control squashing (car.py:150-158): does not depend on the stage state, evaluated once per control set
  template <class S, class MX>
  static ILQC_HD void controls(MX& mx, const ilqc_model_desc& m, const S (&u)[NU], S (&delta)[NSET], S (&acc)[NSET],
                               S (&sdel)[NSET], S (&cdel)[NSET]) {
    const double gap = m.acc_hi - m.acc_lo, rgap = 1.0 / gap;
    S ud[NSET], at[NSET];
#pragma unroll
    for (int g = 0; g < NSET; ++g) ud[g] = u[3 * g + 1];
    mx.atan(ud, at);
#pragma unroll
    for (int g = 0; g < NSET; ++g) {
      delta[g] = ((2.0 / M_PI) * at[g]) * m.constrain_angle;
      acc[g] = gap * mx.sigmoid(mx.divc(4.0 * u[3 * g], gap, rgap)) + m.acc_lo;
    }
    mx.sincos(delta, sdel, cdel);
  }
  struct Par {
    double Cm1, Cm2, Cr0, Cr2, Br, Cr, Dr, Bf, Cf, Df, mass, Iz, lf, lr, rmass, rIz;
  };
  // bicycle_model (car.py:160-188) at one stage state; the front / rear tyre chains (atan2 -> atan -> sin) run
  // as two lanes of the L-lane functions
  template <class S, class MX>
  static ILQC_HD void dyn(MX& mx, const Par& p, const S& delta, const S& acc, const S& sdel, const S& cdel,
                          const S* s, S* d) {
    const S phi = s[2], vx = s[3], vy = s[4], vphi = s[5];
    const S ys[2] = {vphi * p.lf + vy, vphi * p.lr - vy}, xs[2] = {vx, vx};
    S al[2];
    mx.atan2(ys, xs, al);
    const S alphaf = delta - al[0];
    const S alphar = al[1];
    const S ba[2] = {p.Br * alphar, p.Bf * alphaf};
    S at[2];
    mx.atan(ba, at);
    const S sa[3] = {p.Cr * at[0], p.Cf * at[1], phi};
    S sn[3], cs[3];
    mx.sincos(sa, sn, cs);
    const S Fry = p.Dr * sn[0];
    const S Ffy = p.Df * sn[1];
    const S Frx = (p.Cm1 - p.Cm2 * vx) * acc - p.Cr0 - p.Cr2 * jsq(vx);
    const S sphi = sn[2], cphi = cs[2];
    d[0] = cphi * vx - sphi * vy;
    d[1] = sphi * vx + cphi * vy;
    d[2] = vphi;
    d[3] = mx.divc(Frx - Ffy * sdel + (p.mass * vy) * vphi, p.mass, p.rmass);
    d[4] = mx.divc(Fry + Ffy * cdel - (p.mass * vx) * vphi, p.mass, p.rmass);
    d[5] = mx.divc((Ffy * p.lf) * cdel - Fry * p.lr, p.Iz, p.rIz);
  }
  // ---- structured second-order expansion of one stage (hess_kernel) -------------------------------------------------
  // Second-order jets through the whole bicycle model carry 11-21 doubles per scalar over ~150 operations.  The model
  // has far more structure than that: each tyre force is a UNIVARIATE chain (atan -> sin) of one slip angle, which is
  // an atan2 of two linear forms of the stage state.  So the forces are expanded in their own 2 / 3 variables
  // (rear: (vphi lr - vy, vx); front: (vphi lf + vy, vx, delta)), lifted through the linear maps to the 5 local
  // variables (vx, vy, vphi, a, delta), combined by a handful of 5-variable jet products into the three body
  // accelerations, and only then composed with the stage-state jets S (chain rule with the local gradient / Hessian).
  // ~2 kFMA per stage instead of ~11 kFMA, and the stage state is the only wide quantity that stays live.
  static constexpr bool kStructuredHess = !VAR;
  // f(z(zeta)) for a local jet f over NL variables; Z(integral_constant<c>) -> const S& selects inner jet c
  template <int NL, class S, class ZSel>
  static ILQC_HD S composeN(const Jet<NL, 2>& f, ZSel Z) {
    constexpr int N = S::NV;
    S r;
    r.v = f.v;
    static_for<N>([&](auto K) __attribute__((always_inline)) {
      double a = 0.0;
      static_for<NL>([&](auto C) __attribute__((always_inline)) { a += f.g[C] * Z(C).g[K]; });
      r.g[K] = a;
    });
    double Mh[NL][N];   // Mh[c][j] = sum_d f.H[c][d] dz_d/dzeta_j
    static_for<NL>([&](auto C) __attribute__((always_inline)) {
      static_for<N>([&](auto J) __attribute__((always_inline)) {
        double a = 0.0;
        static_for<NL>([&](auto D) __attribute__((always_inline)) { a += f.h[sym_idx(C, D, NL)] * Z(D).g[J]; });
        Mh[C][J] = a;
      });
    });
    static_for<S::NH>([&](auto K) __attribute__((always_inline)) {
      constexpr int k = K;
      if constexpr (S::valid(k)) {
        constexpr int i = SymTable<N>().i[S::entry(k)], j = SymTable<N>().j[S::entry(k)];
        double a = 0.0;
        static_for<NL>([&](auto C) __attribute__((always_inline)) { a += f.g[C] * Z(C).h[k]; });
        static_for<NL>([&](auto C) __attribute__((always_inline)) { a += Z(C).g[i] * Mh[C][j]; });
        r.h[k] = a;
      } else {
        r.h[k] = 0.0;
      }
    });
    return r;
  }
  // Stage dynamics on the Hessian-group jets S with the tyre forces expanded in their own 2 / 3 variables (full small
  // Hessians: 6 / 10 doubles) and composed straight into S; the remaining algebra (a dozen products) runs on S.
  // control-derived factors (acceleration, steering angle, its sine / cosine) depend on ONE jet variable each: carried as
  // univariate jets (value, first, second derivative) they cost 3 doubles instead of a group jet's 11 - four of them stay
  // live across all RK4 stages.  a * b for a group jet a and a univariate b in jet variable VAR:
  using U1 = Jet<1, 2>;
  template <int JV, class S>
  static ILQC_HD S mulU(const S& a, const U1& b) {
    constexpr int N = S::NV;
    S r;
    r.v = a.v * b.v;
    static_for<N>([&](auto I) __attribute__((always_inline)) {
      if constexpr (I == JV) r.g[I] = a.g[I] * b.v + a.v * b.g[0];
      else r.g[I] = a.g[I] * b.v;
    });
    static_for<S::NH>([&](auto K) __attribute__((always_inline)) {
      constexpr int k = K;
      if constexpr (S::valid(k)) {
        constexpr int i = SymTable<N>().i[S::entry(k)], j = SymTable<N>().j[S::entry(k)];
        double h = a.h[k] * b.v;
        if constexpr (j == JV) h += a.g[i] * b.g[0];
        if constexpr (i == JV) h += a.g[j] * b.g[0];
        if constexpr (i == JV && j == JV) h += a.v * b.h[0];
        r.h[k] = h;
      } else {
        r.h[k] = 0.0;
      }
    });
    return r;
  }
  template <int JV, class S>
  static ILQC_HD S toS(const U1& b) {
    S r = make_cst<S>(b.v);
    r.g[JV] = b.g[0];
    static_for<S::NH>([&](auto K) __attribute__((always_inline)) {
      constexpr int k = K;
      if constexpr (S::valid(k)) {
        if constexpr (SymTable<S::NV>().i[S::entry(k)] == JV && SymTable<S::NV>().j[S::entry(k)] == JV) r.h[k] = b.h[0];
      }
    });
    return r;
  }
  template <class S, class MX>
  static ILQC_HD void dyn_hess_vel(MX& mx, const Par& p, const U1& delta, const U1& acc, const U1& sdel, const U1& cdel,
                                   const S& vx, const S& vy, const S& vphi, S& d3, S& d4, S& d5) {
    static_assert(kRotVar == 0, "jet variables without the heading: (vx, vy, vphi, u0, u1)");
    constexpr int VA = uv(0) - 1, VD = uv(1) - 1;   // jet variables of the acceleration / steering control
    S Fry, Ffy;
    {
      using R2 = Jet<2, 2>;  // rear tyre force over (vphi lr - vy, vx)
      const S yr = vphi * p.lr - vy;
      const R2 ys[1] = {make_var<R2>(yr.v, 0)}, xs[1] = {make_var<R2>(vx.v, 1)};
      R2 al[1], at[1], sn[1], cs[1];
      mx.atan2(ys, xs, al);
      const R2 ba[1] = {p.Br * al[0]};
      mx.atan(ba, at);
      const R2 sa[1] = {p.Cr * at[0]};
      mx.sincos(sa, sn, cs);
      Fry = composeN<2, S>(p.Dr * sn[0], [&](auto C) __attribute__((always_inline)) -> const S& {
        if constexpr (C == 0) return yr; else return vx;
      });
    }
    {
      using F3 = Jet<3, 2>;  // front tyre force over (vphi lf + vy, vx, delta)
      const S yf = vphi * p.lf + vy;
      const S dl = toS<VD, S>(delta);
      const F3 ys[1] = {make_var<F3>(yf.v, 0)}, xs[1] = {make_var<F3>(vx.v, 1)};
      F3 al[1], at[1], sn[1], cs[1];
      mx.atan2(ys, xs, al);
      const F3 ba[1] = {p.Bf * (make_var<F3>(delta.v, 2) - al[0])};
      mx.atan(ba, at);
      const F3 sa[1] = {p.Cf * at[0]};
      mx.sincos(sa, sn, cs);
      Ffy = composeN<3, S>(p.Df * sn[0], [&](auto C) __attribute__((always_inline)) -> const S& {
        if constexpr (C == 0) return yf; else if constexpr (C == 1) return vx; else return dl;
      });
    }
    const S Frx = mulU<VA>(p.Cm1 - p.Cm2 * vx, acc) - p.Cr0 - p.Cr2 * jsq(vx);
    const S FfyC = mulU<VD>(Ffy, cdel);
    d3 = mx.divc(Frx - mulU<VD>(Ffy, sdel) + (p.mass * vy) * vphi, p.mass, p.rmass);
    d4 = mx.divc(Fry + FfyC - (p.mass * vx) * vphi, p.mass, p.rmass);
    d5 = mx.divc(FfyC * p.lf - Fry * p.lr, p.Iz, p.rIz);
  }
  template <class S, class MX>
  static ILQC_HD void dyn_hess_pos(MX& mx, const S& phi, const S& vx, const S& vy, S& d0, S& d1) {
    const S ph[1] = {phi};
    S sp[1], cp[1];
    mx.sincos(ph, sp, cp);
    d0 = cp[0] * vx - sp[0] * vy;
    d1 = sp[0] * vx + cp[0] * vy;
  }
  // one discrete step on second-order jets S for the dynamics-Hessian tensor (constant control, RK4 or Euler).
  // The RK4 accumulator k1 + 2 k2 + 2 k3 + k4 (6 jets) is kept in `acc` - shared memory in the CUDA kernel, 33 KB per
  // 64-thread block (bodies.cuh HessAcc) - instead of registers: with it the stage state, the stage slope and the
  // accumulator are 3 x 66 doubles live across the model evaluation, more than the 127 a thread can hold.  Same
  // formulas and operation order as integrate<> (bit-identical results).
  template <class S>
  static constexpr int kHessAccDoubles = 6 * (1 + S::NV + S::NH);
  template <class S, class ACC>
  static ILQC_HD void step_hess(const ilqc_model_desc& m, const S (&x)[NX], const S (&u)[NU], S (&xn)[NX], ACC& acc) {
    U1 u1[NU], delta[NSET], accel[NSET], sdel[NSET], cdel[NSET];   // (every control is its own variable here)
#pragma unroll
    for (int k = 0; k < NU; ++k) u1[k] = make_var<U1>(u[k].v, 0);
    {
      MathHot mx;
      controls<U1>(mx, m, u1, delta, accel, sdel, cdel);
      if constexpr (MathHot::kFast) {
        if (ILQC_UNLIKELY(mx.bad)) {
          MathExact me;
          controls<U1>(me, m, u1, delta, accel, sdel, cdel);
        }
      }
    }
    const Par p = {m.Cm1, m.Cm2, m.Cr0, m.Cr2, m.Br, m.Cr, m.Dr, m.Bf, m.Cf, m.Df, m.m, m.Iz, m.lf, m.lr, 1.0 / m.m, 1.0 / m.Iz};
    constexpr int NE = 1 + S::NV + S::NH;   // doubles per jet
    auto elem = [](auto& jet, auto E) __attribute__((always_inline)) -> auto& {   // e-th double of a jet: v, g[0..NV), h[0..NH)
      constexpr int e = E;
      if constexpr (e == 0) return jet.v;
      else if constexpr (e <= S::NV) return jet.g[e - 1];
      else return jet.h[e - 1 - S::NV];
    };
    const double dt = m.dt;
    // Stage slopes are consumed as soon as they exist, in an order that keeps the live set small: the position rows
    // (they feed only the accumulator: x, y do not enter the dynamics), then the heading row (a copy of the yaw rate),
    // then the three body accelerations, which replace the stage velocities.  (always_inline lambdas: left to the
    // inliner's size heuristic the stage bodies become real calls and the jets they pass by address live in local memory.)
    auto pos_rows = [&](const S& phi, const S& vx, const S& vy, S& d0, S& d1) __attribute__((always_inline)) {
      MathHot mx;
      dyn_hess_pos<S>(mx, phi, vx, vy, d0, d1);
      if constexpr (MathHot::kFast) {
        if (ILQC_UNLIKELY(mx.bad)) {
          MathExact me;
          dyn_hess_pos<S>(me, phi, vx, vy, d0, d1);
        }
      }
    };
    auto vel_rows = [&](const S& vx, const S& vy, const S& w, S& d3, S& d4, S& d5) __attribute__((always_inline)) {
      MathHot mx;
      dyn_hess_vel<S>(mx, p, delta[0], accel[0], sdel[0], cdel[0], vx, vy, w, d3, d4, d5);
      if constexpr (MathHot::kFast) {
        if (ILQC_UNLIKELY(mx.bad)) {
          MathExact me;
          dyn_hess_vel<S>(me, p, delta[0], accel[0], sdel[0], cdel[0], vx, vy, w, d3, d4, d5);
        }
      }
    };
    if (m.integrator == ILQC_INT_EULER) {
      S k[6];
      pos_rows(x[2], x[3], x[4], k[0], k[1]);
      k[2] = x[5];
      vel_rows(x[3], x[4], x[5], k[3], k[4], k[5]);
#pragma unroll
      for (int i = 0; i < 6; ++i) xn[i] = x[i] + dt * k[i];
    } else {
      S s2 = x[2], s3 = x[3], s4 = x[4], s5 = x[5];
      static_for<4>([&](auto ST) __attribute__((always_inline)) {
        constexpr int stage = ST;
        constexpr double wgt = (stage == 1 || stage == 2) ? 2.0 : 1.0;   // k1 + 2 k2 + 2 k3 + k4
        constexpr double half = (stage < 2) ? 0.5 : 1.0;                 // next stage state: dt*k/2, dt*k/2, dt*k
        auto accumulate = [&](auto I, const S& k) __attribute__((always_inline)) {
          static_for<NE>([&](auto E) __attribute__((always_inline)) {
            const double ke = elem(k, E);
            if constexpr (stage == 0) acc[I * NE + E] = ke;
            else acc[I * NE + E] = acc[I * NE + E] + wgt * ke;
          });
        };
        auto advance = [&](auto I, const S& k, S& sn) __attribute__((always_inline)) {
          if constexpr (stage < 3)
            static_for<NE>([&](auto E) __attribute__((always_inline)) { elem(sn, E) = elem(x[I], E) + (dt * elem(k, E)) * half; });
        };
        {
          S d0, d1;
          pos_rows(s2, s3, s4, d0, d1);
          accumulate(std::integral_constant<int, 0>{}, d0);
          accumulate(std::integral_constant<int, 1>{}, d1);
        }
        accumulate(std::integral_constant<int, 2>{}, s5);
        advance(std::integral_constant<int, 2>{}, s5, s2);
        S d3, d4, d5;
        vel_rows(s3, s4, s5, d3, d4, d5);
        accumulate(std::integral_constant<int, 3>{}, d3);
        accumulate(std::integral_constant<int, 4>{}, d4);
        accumulate(std::integral_constant<int, 5>{}, d5);
        advance(std::integral_constant<int, 3>{}, d3, s3);
        advance(std::integral_constant<int, 4>{}, d4, s4);
        advance(std::integral_constant<int, 5>{}, d5, s5);
      });
      MathHot mh;  // (dt * acc) / 6 as in integrate<>
      static_for<6>([&](auto I) __attribute__((always_inline)) {
        static_for<NE>([&](auto E) __attribute__((always_inline)) {
          const double a = dt * acc[I * NE + E];
          if constexpr (E == 0) elem(xn[I], E) = elem(x[I], E) + mh.divc_v(a, 6.0, 1.0 / 6.0);
          else elem(xn[I], E) = elem(x[I], E) + a * (1.0 / 6.0);
        });
      });
    }
    xn[6] = make_cst<S>(0.0);   // (tref, vtref) evolve linearly: no second-order terms, rows not stored)
    xn[7] = make_cst<S>(0.0);
  }
  template <class S>
  static ILQC_HD void step(const ilqc_model_desc& m, const S (&x)[NX], const S (&u)[NU], S (&xn)[NX]) {
    S delta[NSET], acc[NSET], sdel[NSET], cdel[NSET];
    double atref[NSET];
#pragma unroll
    for (int g = 0; g < NSET; ++g) atref[g] = val(u[3 * g + 2]);
    {
      MathHot mx;
      controls<S>(mx, m, u, delta, acc, sdel, cdel);
      if constexpr (MathHot::kFast) {
        if (ILQC_UNLIKELY(mx.bad)) {
          MathExact me;
          controls<S>(me, m, u, delta, acc, sdel, cdel);
        }
      }
    }
    const Par p = {m.Cm1, m.Cm2, m.Cr0, m.Cr2, m.Br, m.Cr, m.Dr, m.Bf, m.Cf, m.Df, m.m, m.Iz, m.lf, m.lr, 1.0 / m.m, 1.0 / m.Iz};
    const S xv[6] = {x[0], x[1], x[2], x[3], x[4], x[5]};
    S xvn[6];
    auto f = [&](int set, const S* s, S* d) {
      MathHot mx;
      dyn<S>(mx, p, delta[set], acc[set], sdel[set], cdel[set], s, d);
      if constexpr (MathHot::kFast) {
        if (ILQC_UNLIKELY(mx.bad)) {  // an argument outside the fast functions' range: CUDA math library
          MathExact me;
          dyn<S>(me, p, delta[set], acc[set], sdel[set], cdel[set], s, d);
        }
      }
    };
    integrate<S, 6, VAR>(m.integrator, m.dt, f, xv, xvn);
#pragma unroll
    for (int i = 0; i < 6; ++i) xn[i] = xvn[i];
    double tn, vn;
    car_tref_step<VAR>(m, val(x[6]), val(x[7]), atref, &tn, &vn);
    xn[6] = make_cst<S>(tn);
    xn[7] = make_cst<S>(vn);
  }
  static constexpr int kPassivePRow = 6, kPassivePCol = 7;
  static ILQC_HD double passive_P(const ilqc_model_desc& m, const Weights& w) { return car_passive_P(m, w); }
  static ILQC_HD double passive_N(const ilqc_model_desc& m, int i, int j) { return (i == 6 && j == 7) ? m.dt : 0.0; }
  static ILQC_HD double passive_B(const ilqc_model_desc& m, int i, int k) {
    if (k % 3 != 2 || i < 6) return 0.0;
    return car_tref_B<VAR>(m, i == 7, k / 3);
  }
  template <class S>
  static ILQC_HD S state_cost(const ilqc_model_desc& m, const Weights& w, const S (&x)[NX], int time_iter) {
    return car_state_cost<S, NX>(m, w, x, time_iter);
  }
  static ILQC_HD void cost_expansion(const ilqc_model_desc& m, const Weights& w, const double (&x)[NX], int time_iter,
                                     double* cval, double (&g)[NVC], double (&h)[sym_size(NVC)]) {
    car_cost_expansion<NX>(m, w, x, time_iter, cval, g, h);
  }
  static ILQC_HD double ctrl_cost(const ilqc_model_desc& m, const double (&u)[NU], double (&q)[NU],
                                  double (&Qd)[NU]) {
    return car_ctrl_cost<NU>(m, u, q, Qd);
  }
};
template <> struct Model<ILQC_MODEL_CAR_BICYCLE> : CarBicycleT<false> {};
template <> struct Model<ILQC_MODEL_CAR_BICYCLE + ILQC_MODEL_VAR_OFFSET> : CarBicycleT<true> {};

template <class M>
ILQC_HD double passive_N_of(const ilqc_model_desc& m, int i, int j) { return M::passive_N(m, i, j); }
template <class M>
ILQC_HD double passive_B_of(const ilqc_model_desc& m, int i, int k) { return M::passive_B(m, i, k); }

// ---------------------------------------------------------------------------------------------------
// derived compile-time layout of the packed expansion of model M (doubles per time step per problem):
//   [ N entries | B entries | q (NU) | Qd (NU) | p_next entries | P_next entries ]
// "next" = expansion of the state cost charged on x_{t+1} (costs[t+1] in the reference).
// ---------------------------------------------------------------------------------------------------
template <class M>
struct Layout {
  static constexpr int nN = M::Nm.count(), nB = M::Bm.count(), np = M::pm.count(), nP = M::Pm.count();
  static constexpr int oN = 0, oB = oN + nN, oq = oB + nB, oQ = oq + M::NU, op = oQ + M::NU, oP = op + np;
  static constexpr int stride = oP + nP;
  static constexpr int gain_stride = M::NU * M::NX + M::NU;
  // second-order tensor of the dynamics: packed Hessians of the first NXH state rows over the jet variables WITHOUT the
  // heading (M::kRotVar): rows >= 2 do not depend on it and the heading entries of rows 0, 1 are rotations of the
  // first-order expansion (bodies.cuh backward_body), so 6 x 15 instead of 6 x 21 doubles per step for the bicycle car
  static constexpr int hess_nvj = M::kRotVar >= 0 ? M::NV - 1 : M::NV;
  static constexpr int hess_row = sym_size(hess_nvj);
  static constexpr int hess_stride = M::NXH * hess_row;
  static constexpr int hess_red(int v) { return (v < 0 || v == M::kRotVar) ? -1 : ((M::kRotVar >= 0 && v > M::kRotVar) ? v - 1 : v); }
  // masks of the second-order path: Hessian of lambda^T phi lives on the jet variables
  static constexpr Mask<M::NX, M::NX> hxx() {
    Mask<M::NX, M::NX> k;
    for (int i = 0; i < M::NX; ++i)
      for (int j = i; j < M::NX; ++j) k.m[i][j] = M::sv(i) >= 0 && M::sv(j) >= 0;
    return k;
  }
  static constexpr Mask<M::NX, M::NU> hxu() {
    Mask<M::NX, M::NU> k;
    for (int i = 0; i < M::NX; ++i)
      for (int j = 0; j < M::NU; ++j) k.m[i][j] = M::sv(i) >= 0 && M::uv(j) >= 0;
    return k;
  }
  static constexpr Mask<M::NU, M::NU> huu() {
    Mask<M::NU, M::NU> k;
    for (int i = 0; i < M::NU; ++i)
      for (int j = i; j < M::NU; ++j) k.m[i][j] = M::uv(i) >= 0 && M::uv(j) >= 0;
    return k;
  }
};

}  // namespace ilqc

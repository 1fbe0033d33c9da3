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
template <class S, int N, class F>
ILQC_HD void integrate_cst_ctrl(int integ, double dt, F&& f, const S (&x)[N], S (&xn)[N]) {
  if (integ == ILQC_INT_EULER) {
    S k[N];
    f(x, k);
#pragma unroll
    for (int i = 0; i < N; ++i) xn[i] = x[i] + dt * k[i];
  } else {
    // runge_kutta4_cst_ctrl: state + dt*(k1 + 2*k2 + 2*k3 + k4)/6 with stage states state + dt*k1/2,
    // state + dt*k2/2, state + dt*k3.  Written as a rolled loop over the four stages so that the (large,
    // transcendental-heavy) dynamics body is instantiated once: 4x less code than the unrolled form, which
    // matters for the instruction cache (ncu: 22 % of the expansion kernel's stall cycles were
    // no_instruction).  Bit-identical to the unrolled formulas: x/2 == x*0.5 and 1*k, 2*k are exact.
    // (measured on B200, profiles/r01_tuning_kernels.txt: rolled is 9-12 % faster for the plain roll-out and
    // cuts the spills of the second-order jets by a third, but 5 % slower for the first-order expansion,
    // which keeps the unrolled form)
    constexpr int kStageUnroll = (is_jet<S>::value && sizeof(S) <= sizeof(double) * 16) ? 4 : 1;
    S k[N], s[N], acc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = x[i];
#pragma unroll(kStageUnroll)
    for (int stage = 0; stage < 4; ++stage) {
      f(s, k);
      const double wgt = (stage == 1 || stage == 2) ? 2.0 : 1.0;  // k1 + 2 k2 + 2 k3 + k4
      const double half = (stage < 2) ? 0.5 : 1.0;                 // next stage state: dt*k/2, dt*k/2, dt*k
#pragma unroll
      for (int i = 0; i < N; ++i) {
        if (stage == 0) acc[i] = k[i]; else acc[i] = acc[i] + wgt * k[i];
        s[i] = x[i] + (dt * k[i]) * half;
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) xn[i] = x[i] + (dt * acc[i]) / 6.0;
  }
}

// ---------------------------------------------------------------------------------------------------
// natural cubic spline lookup (torch_spline.py:341-375).  One segment search serves the three splines
// of a track (shared knots, tracks.py:50-53).
// ---------------------------------------------------------------------------------------------------
struct Seg {
  const double* c;  // 24 coefficients of the segment: [spline][channel][a,b,c,d]
};
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
template <int MODEL>
struct Model;

// ---------------------------------------------------------------------------------------------------
template <>
struct Model<ILQC_MODEL_PENDULUM> {
  static constexpr int NX = 2, NU = 1, NV = 3, NVC = 2;
  static constexpr int sv(int i) { constexpr int t[NX] = {0, 1}; return t[i]; }
  static constexpr int uv(int i) { constexpr int t[NU] = {2}; return t[i]; }
  static constexpr int cv(int i) { constexpr int t[NX] = {0, 1}; return t[i]; }
  static constexpr bool Q_DIAG = true;
  static constexpr auto Nm = full_mask<NX, NX>();
  static constexpr auto Bm = full_mask<NX, NU>();
  static constexpr auto pm = full_mask<1, NX>();
  static constexpr auto Pm = upper_of(full_mask<NX, NX>());

  // dyn (pendulum.py:56-66) + euler
  template <class S>
  static ILQC_HD void step(const ilqc_model_desc& m, const S (&x)[NX], const S (&u)[NU], S (&xn)[NX]) {
    const double g = m.phys[0], mm = m.phys[1], l = m.phys[2], mu = m.phys[3];
    const double c_sin = -g / l, c_fr = mu / (mm * (l * l)), c_u = 1.0 / (mm * (l * l));
    auto f = [&](const S* s, S* d) {
      d[0] = s[1];
      d[1] = c_sin * jsin(s[0] + M_PI) - c_fr * s[1] + c_u * u[0];
    };
    integrate_cst_ctrl<S, NX>(m.integrator, m.dt, f, x, xn);
  }
  // cost_state (pendulum.py:82-87): time_iter is the index of the state the cost is charged on
  template <class S>
  static ILQC_HD S state_cost(const ilqc_model_desc& m, const Weights&, const S (&x)[NX], int time_iter) {
    if (time_iter >= m.stay_put_start) return jsq(x[0]) + m.reg_speed * jsq(x[1]);
    return make_cst<S>(0.0);
  }
  static ILQC_HD double ctrl_cost(const ilqc_model_desc& m, const double (&u)[NU], double (&q)[NU],
                                  double (&Qd)[NU]) {
    q[0] = 2.0 * m.reg_ctrl * u[0];
    Qd[0] = 2.0 * m.reg_ctrl;
    return m.reg_ctrl * (u[0] * u[0]);
  }
};

// ---------------------------------------------------------------------------------------------------
template <>
struct Model<ILQC_MODEL_CARTPOLE> {
  static constexpr int NX = 4, NU = 1, NV = 4, NVC = 3;
  static constexpr int sv(int i) { constexpr int t[NX] = {-1, 0, 1, 2}; return t[i]; }
  static constexpr int uv(int i) { constexpr int t[NU] = {3}; return t[i]; }
  static constexpr int cv(int i) { constexpr int t[NX] = {0, 1, -1, 2}; return t[i]; }
  static constexpr bool Q_DIAG = true;
  static constexpr auto Nm = parse_mask<NX, NX>(".xxx/.xxx/.xxx/.xxx");
  static constexpr auto Bm = full_mask<NX, NU>();
  static constexpr auto pm = parse_mask<1, NX>("xx.x");
  static constexpr auto Pm = parse_mask<NX, NX>("x.../.x../..../...x");

  // dyn (pendulum.py:180-198), literal ordering: unpack (x, xdot, th, thdot) = s0..s3, return
  // (xdot, thdot, dxdot, dthdot)
  template <class S>
  static ILQC_HD void step(const ilqc_model_desc& m, const S (&x)[NX], const S (&u)[NU], S (&xn)[NX]) {
    const double g = m.phys[0], M = m.phys[1], mm = m.phys[2], b = m.phys[3], I = m.phys[4], l = m.phys[5];
    const double a11 = M + mm, a22 = I + mm * (l * l), ml = mm * l;
    const double d0 = I * (M + mm) + mm * (l * l) * M, d1 = (mm * mm) * (l * l);
    const double mgl = -mm * g * l;
    auto f = [&](const S* s, S* d) {
      const S xdot = s[1], th = s[2], thdot = s[3];
      const S sth = jsin(th), cth = jcos(th);
      const S a12 = ml * cth;
      const S detA = d0 + d1 * jsq(sth);
      const S b1 = (ml * jsq(thdot)) * sth - b * xdot + u[0];
      const S b2 = mgl * sth;
      d[0] = xdot;
      d[1] = thdot;
      d[2] = (a22 * b1 - a12 * b2) / detA;
      d[3] = (a11 * b2 - a12 * b1) / detA;
    };
    integrate_cst_ctrl<S, NX>(m.integrator, m.dt, f, x, xn);
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
  static ILQC_HD double ctrl_cost(const ilqc_model_desc& m, const double (&u)[NU], double (&q)[NU],
                                  double (&Qd)[NU]) {
    q[0] = 2.0 * m.reg_ctrl * u[0];
    Qd[0] = 2.0 * m.reg_ctrl;
    return m.reg_ctrl * (u[0] * u[0]);
  }
};

// ---------------------------------------------------------------------------------------------------
// Car: shared cost code (car.py:190-302).  XI/YI/TI/VI = indices of x, y, tref, vtref in the state.
// ---------------------------------------------------------------------------------------------------
template <class S, int NX>
ILQC_HD S car_state_cost(const ilqc_model_desc& m, const Weights& w, const S (&x)[NX], int time_iter) {
  const S& px = x[0];
  const S& py = x[1];
  if (m.cost == ILQC_COST_EXACT) {
    // time = time_iter * vref * dt is a constant of the graph (car.py:229-231)
    const double time = ((double)time_iter * m.vref) * m.dt;
    double fr;
    const Seg sg = spline_locate<double>(m, time, &fr);
    const double xr = spline_eval<double>(sg, 0, 0, fr), yr = spline_eval<double>(sg, 0, 1, fr);
    return jsq(px - xr) + jsq(py - yr);
  }
  const S& tref = x[NX - 2];
  const S& vtref = x[NX - 1];
  S fr;
  const Seg sg = spline_locate<S>(m, tref, &fr);
  const S xr = spline_eval<S>(sg, 0, 0, fr), yr = spline_eval<S>(sg, 0, 1, fr);
  const S dxr = spline_deriv<S>(sg, 0, 0, fr), dyr = spline_deriv<S>(sg, 0, 1, fr);
  const S phi_ref = jatan2(dyr, dxr);
  const S sp = jsin(phi_ref), cp = jcos(phi_ref);
  const S ex = px - xr, ey = py - yr;
  const S cont_err = sp * ex - cp * ey;
  const S lag_err = -cp * ex - sp * ey;
  S c = w.reg_cont * jsq(cont_err) + w.reg_lag * jsq(lag_err);
  if (m.time_penalty == ILQC_TP_VREF_SQUARED) c = c + (w.reg_speed * jsq(vtref - m.vref)) * (m.dt * m.dt);
  else if (m.time_penalty == ILQC_TP_TREF) c = c - w.reg_speed * tref;
  else if (m.time_penalty == ILQC_TP_DREF) c = c - (w.reg_speed * vtref) * m.dt;
  else c = c - ((w.reg_speed * vtref) * m.dt) * tref;
  // barrier (car.py:255-302)
  S bar = -1e-6 * jlog(vtref);
  const double carwidth = 1.5 * m.car_w;
  const double pxv = val(px), pyv = val(py);  // point = torch.tensor([x, y]) is detached (car.py:262)
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const S bx = spline_eval<S>(sg, 1 + i, 0, fr), by = spline_eval<S>(sg, 1 + i, 1, fr);
    const S dbx = spline_deriv<S>(sg, 1 + i, 0, fr), dby = spline_deriv<S>(sg, 1 + i, 1, fr);
    const S v = jsqrt(jsq(dbx) + jsq(dby));
    // normal = torch.tensor([dpos[1], -dpos[0]]) / v : numerator detached (car.py:267)
    const S n0 = val(dby) / v, n1 = (-val(dbx)) / v;
    const S dot = (pxv - bx) * n0 + (pyv - by) * n1;
    if (i == 0) bar = bar + m.reg_bar * jcube(jmax0(carwidth / 2.0 - dot));
    else bar = bar + m.reg_bar * jcube(jmax0(dot + carwidth / 2.0));
  }
  if (m.reg_obs > 0.0) {
    const double d2 = (pxv - m.obs_x) * (pxv - m.obs_x) + (pyv - m.obs_y) * (pyv - m.obs_y);
    const double o = jmax0(carwidth * carwidth - d2);
    bar = bar + m.reg_obs * ((o * o) * o);
  }
  return c + bar;
}

template <int NU>
ILQC_HD double car_ctrl_cost(const ilqc_model_desc& m, const double (&u)[NU], double (&q)[NU], double (&Qd)[NU]) {
  if (m.reg_ctrl_is_tuple) {
    // reg_ctrl[0]*(a^2+delta^2) + reg_ctrl[1]*atref^2 (car.py:235-241)
    const double r0 = m.reg_ctrl, r1 = m.reg_ctrl_tref;
    q[0] = 2.0 * r0 * u[0]; q[1] = 2.0 * r0 * u[1]; q[2] = 2.0 * r1 * u[2];
    Qd[0] = 2.0 * r0; Qd[1] = 2.0 * r0; Qd[2] = 2.0 * r1;
    return r0 * (u[0] * u[0] + u[1] * u[1]) + r1 * (u[2] * u[2]);
  }
  double dot = 0.0;
#pragma unroll
  for (int i = 0; i < NU; ++i) {
    dot += u[i] * u[i];
    q[i] = 2.0 * m.reg_ctrl * u[i];
    Qd[i] = 2.0 * m.reg_ctrl;
  }
  return m.reg_ctrl * dot;
}

// reference-time double integrator (tref, vtref ; atref), the same for both car models (car.py:143-146)
ILQC_HD void car_tref_step(const ilqc_model_desc& m, double tref, double vtref, double atref, double* trefn,
                           double* vtrefn) {
  const double x[2] = {tref, vtref};
  double xn[2];
  auto f = [&](const double* s, double* d) { d[0] = s[1]; d[1] = atref; };
  integrate_cst_ctrl<double, 2>(m.integrator, m.dt, f, x, xn);
  *trefn = xn[0];
  *vtrefn = xn[1];
}
// its constant Jacobian: d tref'/d vtref = dt ; d tref'/d atref = dt^2/2 (rk4) or 0 (euler) ; d vtref'/d atref = dt
ILQC_HD double car_tref_B(const ilqc_model_desc& m) { return m.integrator == ILQC_INT_EULER ? 0.0 : (m.dt * m.dt) / 2.0; }

// ---------------------------------------------------------------------------------------------------
template <>
struct Model<ILQC_MODEL_CAR_SIMPLE> {
  // state (x, y, phi, v, tref, vtref), ctrl (a, delta, atref)
  static constexpr int NX = 6, NU = 3, NV = 4, NVC = 4;
  static constexpr int sv(int i) { constexpr int t[NX] = {-1, -1, 0, 1, -1, -1}; return t[i]; }
  static constexpr int uv(int i) { constexpr int t[NU] = {2, 3, -1}; return t[i]; }
  static constexpr int cv(int i) { constexpr int t[NX] = {0, 1, -1, -1, 2, 3}; return t[i]; }
  static constexpr bool Q_DIAG = true;
  static constexpr auto Nm = parse_mask<NX, NX>("..xx../..xx../...x../....../.....x/......");
  static constexpr auto Bm = parse_mask<NX, NU>("xx./xx./xx./x../..x/..x");
  static constexpr auto pm = parse_mask<1, NX>("xx..xx");
  static constexpr auto Pm = parse_mask<NX, NX>("xx..x./.x..x./....../....../....x./.....x");

  template <class S>
  static ILQC_HD void step(const ilqc_model_desc& m, const S (&x)[NX], const S (&u)[NU], S (&xn)[NX]) {
    const S delta = ((2.0 / M_PI) * jatan(u[1])) * m.constrain_angle;
    const S tand = jtan(delta);
    const double carlength = 1.5 * m.car_l;
    // vehicle part: (x, y, phi, v)
    const S xv[4] = {x[0], x[1], x[2], x[3]};
    S xvn[4];
    auto f = [&](const S* s, S* d) {
      const S sphi = jsin(s[2]), cphi = jcos(s[2]);
      d[0] = s[3] * cphi;
      d[1] = s[3] * sphi;
      d[2] = (s[3] * tand) / carlength;
      d[3] = u[0];
    };
    integrate_cst_ctrl<S, 4>(m.integrator, m.dt, f, xv, xvn);
#pragma unroll
    for (int i = 0; i < 4; ++i) xn[i] = xvn[i];
    double tn, vn;
    car_tref_step(m, val(x[4]), val(x[5]), val(u[2]), &tn, &vn);
    xn[4] = make_cst<S>(tn);
    xn[5] = make_cst<S>(vn);
  }
  // constant-coefficient part of (N, B) that the jets do not carry
  static ILQC_HD double passive_N(const ilqc_model_desc& m, int i, int j) { return (i == 4 && j == 5) ? m.dt : 0.0; }
  static ILQC_HD double passive_B(const ilqc_model_desc& m, int i, int k) {
    if (k != 2) return 0.0;
    return i == 4 ? car_tref_B(m) : (i == 5 ? m.dt : 0.0);
  }
  template <class S>
  static ILQC_HD S state_cost(const ilqc_model_desc& m, const Weights& w, const S (&x)[NX], int time_iter) {
    return car_state_cost<S, NX>(m, w, x, time_iter);
  }
  static ILQC_HD double ctrl_cost(const ilqc_model_desc& m, const double (&u)[NU], double (&q)[NU],
                                  double (&Qd)[NU]) {
    return car_ctrl_cost<NU>(m, u, q, Qd);
  }
};

// ---------------------------------------------------------------------------------------------------
template <>
struct Model<ILQC_MODEL_CAR_BICYCLE> {
  // state (x, y, phi, vx, vy, vphi, tref, vtref), ctrl (a, delta, atref)
  static constexpr int NX = 8, NU = 3, NV = 6, NVC = 4;
  static constexpr int sv(int i) { constexpr int t[NX] = {-1, -1, 0, 1, 2, 3, -1, -1}; return t[i]; }
  static constexpr int uv(int i) { constexpr int t[NU] = {4, 5, -1}; return t[i]; }
  static constexpr int cv(int i) { constexpr int t[NX] = {0, 1, -1, -1, -1, -1, 2, 3}; return t[i]; }
  static constexpr bool Q_DIAG = true;
  static constexpr auto Nm =
      parse_mask<NX, NX>("..xxxx../..xxxx../...xxx../...xxx../...xxx../...xxx../.......x/........");
  static constexpr auto Bm = parse_mask<NX, NU>("xx./xx./xx./xx./xx./xx./..x/..x");
  static constexpr auto pm = parse_mask<1, NX>("xx....xx");
  static constexpr auto Pm =
      parse_mask<NX, NX>("xx....x./.x....x./......../......../......../......../......x./.......x");

  template <class S>
  static ILQC_HD void step(const ilqc_model_desc& m, const S (&x)[NX], const S (&u)[NU], S (&xn)[NX]) {
    // control squashing is stage-independent under constant control: evaluate once
    const S delta = ((2.0 / M_PI) * jatan(u[1])) * m.constrain_angle;
    const double gap = m.acc_hi - m.acc_lo;
    const S a = gap * jsigmoid((4.0 * u[0]) / gap) + m.acc_lo;
    const S sdel = jsin(delta), cdel = jcos(delta);
    const double Cm1 = m.Cm1, Cm2 = m.Cm2, Cr0 = m.Cr0, Cr2 = m.Cr2, Br = m.Br, Cr = m.Cr, Dr = m.Dr, Bf = m.Bf,
                 Cf = m.Cf, Df = m.Df, mass = m.m, Iz = m.Iz, lf = m.lf, lr = m.lr;
    const S xv[6] = {x[0], x[1], x[2], x[3], x[4], x[5]};
    S xvn[6];
    auto f = [&](const S* s, S* d) {
      const S phi = s[2], vx = s[3], vy = s[4], vphi = s[5];
      const S alphaf = delta - jatan2(vphi * lf + vy, vx);
      const S alphar = jatan2(vphi * lr - vy, vx);
      const S Fry = Dr * jsin(Cr * jatan(Br * alphar));
      const S Ffy = Df * jsin(Cf * jatan(Bf * alphaf));
      const S Frx = (Cm1 - Cm2 * vx) * a - Cr0 - Cr2 * jsq(vx);
      const S sphi = jsin(phi), cphi = jcos(phi);
      d[0] = cphi * vx - sphi * vy;
      d[1] = sphi * vx + cphi * vy;
      d[2] = vphi;
      d[3] = (Frx - Ffy * sdel + (mass * vy) * vphi) / mass;
      d[4] = (Fry + Ffy * cdel - (mass * vx) * vphi) / mass;
      d[5] = ((Ffy * lf) * cdel - Fry * lr) / Iz;
    };
    integrate_cst_ctrl<S, 6>(m.integrator, m.dt, f, xv, xvn);
#pragma unroll
    for (int i = 0; i < 6; ++i) xn[i] = xvn[i];
    double tn, vn;
    car_tref_step(m, val(x[6]), val(x[7]), val(u[2]), &tn, &vn);
    xn[6] = make_cst<S>(tn);
    xn[7] = make_cst<S>(vn);
  }
  static ILQC_HD double passive_N(const ilqc_model_desc& m, int i, int j) { return (i == 6 && j == 7) ? m.dt : 0.0; }
  static ILQC_HD double passive_B(const ilqc_model_desc& m, int i, int k) {
    if (k != 2) return 0.0;
    return i == 6 ? car_tref_B(m) : (i == 7 ? m.dt : 0.0);
  }
  template <class S>
  static ILQC_HD S state_cost(const ilqc_model_desc& m, const Weights& w, const S (&x)[NX], int time_iter) {
    return car_state_cost<S, NX>(m, w, x, time_iter);
  }
  static ILQC_HD double ctrl_cost(const ilqc_model_desc& m, const double (&u)[NU], double (&q)[NU],
                                  double (&Qd)[NU]) {
    return car_ctrl_cost<NU>(m, u, q, Qd);
  }
};

// Pendulum / CartPendulum carry every dependence in their jets
template <class M>
ILQC_HD double passive_N_of(const ilqc_model_desc& m, int i, int j) {
  if constexpr (M::NX >= 6) return M::passive_N(m, i, j); else return 0.0;
}
template <class M>
ILQC_HD double passive_B_of(const ilqc_model_desc& m, int i, int k) {
  if constexpr (M::NX >= 6) return M::passive_B(m, i, k); else return 0.0;
}

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

// Forward-mode Taylor arithmetic ("jets") used to build the analytic first/second-order expansions
// that the reference obtains from torch.autograd (envs/forward.py:188-321, envs/torch_utils.py:8-41).
//
// Jet<N,1> carries value + gradient wrt N seeded variables; Jet<N,2> additionally the packed symmetric
// Hessian.  The model code (models.cuh) is written once, templated on the scalar type S, and is
// instantiated with S=double for plain roll-outs and S=Jet<N,ORD> for the expansions, so values and
// derivatives come from the same arithmetic.  All loops have compile-time bounds and fully unroll, the
// jets live in registers (thread-per-problem mapping).
#pragma once
#include "common.cuh"

namespace ilqc {

template <int N, int ORD>
struct Jet {
  static constexpr int NV = N;
  static constexpr int NH = (ORD >= 2) ? sym_size(N) : 1;
  double v;
  double g[N];
  double h[NH];
};

template <class S> struct is_jet : std::false_type {};
template <int N, int O> struct is_jet<Jet<N, O>> : std::true_type {};

// ---- construction ------------------------------------------------------------------------------------
template <class S>
ILQC_HD S make_cst(double v) {
  if constexpr (is_jet<S>::value) {
    S r;
    r.v = v;
#pragma unroll
    for (int i = 0; i < S::NV; ++i) r.g[i] = 0.0;
#pragma unroll
    for (int i = 0; i < S::NH; ++i) r.h[i] = 0.0;
    return r;
  } else {
    return v;
  }
}
template <class S>
ILQC_HD S make_var(double v, int idx) {  // idx < 0: unseeded constant
  if constexpr (is_jet<S>::value) {
    S r = make_cst<S>(v);
#pragma unroll
    for (int i = 0; i < S::NV; ++i) r.g[i] = (i == idx) ? 1.0 : 0.0;
    return r;
  } else {
    return v;
  }
}
ILQC_HD double val(double a) { return a; }
template <int N, int O> ILQC_HD double val(const Jet<N, O>& a) { return a.v; }
ILQC_HD void set_val(double& a, double v) { a = v; }
template <int N, int O> ILQC_HD void set_val(Jet<N, O>& a, double v) { a.v = v; }
// autograd's .detach()/torch.tensor(x): keep the value, drop the derivatives (envs/car.py:262, :267)
template <class S> ILQC_HD S detach(const S& a) { return make_cst<S>(val(a)); }

// ---- generic chain rules -----------------------------------------------------------------------------
template <int N, int O>
ILQC_HD Jet<N, O> chain1(const Jet<N, O>& a, double f0, double f1, double f2) {
  Jet<N, O> r;
  r.v = f0;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = f1 * a.g[i];
  if constexpr (O >= 2) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) {
        const int k = sym_idx(i, j, N);
        r.h[k] = f1 * a.h[k] + f2 * (a.g[i] * a.g[j]);
      }
  }
  return r;
}
template <int N, int O>
ILQC_HD Jet<N, O> chain2(const Jet<N, O>& a, const Jet<N, O>& b, double f0, double fa, double fb, double faa,
                         double fab, double fbb) {
  Jet<N, O> r;
  r.v = f0;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = fa * a.g[i] + fb * b.g[i];
  if constexpr (O >= 2) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) {
        const int k = sym_idx(i, j, N);
        r.h[k] = fa * a.h[k] + fb * b.h[k] + faa * (a.g[i] * a.g[j]) + fbb * (b.g[i] * b.g[j]) +
                 fab * (a.g[i] * b.g[j] + a.g[j] * b.g[i]);
      }
  }
  return r;
}

// ---- arithmetic --------------------------------------------------------------------------------------
template <int N, int O>
ILQC_HD Jet<N, O> operator+(const Jet<N, O>& a, const Jet<N, O>& b) {
  Jet<N, O> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = a.g[i] + b.g[i];
  if constexpr (O >= 2) {
#pragma unroll
    for (int i = 0; i < Jet<N, O>::NH; ++i) r.h[i] = a.h[i] + b.h[i];
  }
  return r;
}
template <int N, int O>
ILQC_HD Jet<N, O> operator-(const Jet<N, O>& a, const Jet<N, O>& b) {
  Jet<N, O> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = a.g[i] - b.g[i];
  if constexpr (O >= 2) {
#pragma unroll
    for (int i = 0; i < Jet<N, O>::NH; ++i) r.h[i] = a.h[i] - b.h[i];
  }
  return r;
}
template <int N, int O>
ILQC_HD Jet<N, O> operator-(const Jet<N, O>& a) {
  Jet<N, O> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = -a.g[i];
  if constexpr (O >= 2) {
#pragma unroll
    for (int i = 0; i < Jet<N, O>::NH; ++i) r.h[i] = -a.h[i];
  }
  return r;
}
template <int N, int O> ILQC_HD Jet<N, O> operator+(const Jet<N, O>& a, double b) { Jet<N, O> r = a; r.v = a.v + b; return r; }
template <int N, int O> ILQC_HD Jet<N, O> operator+(double b, const Jet<N, O>& a) { Jet<N, O> r = a; r.v = b + a.v; return r; }
template <int N, int O> ILQC_HD Jet<N, O> operator-(const Jet<N, O>& a, double b) { Jet<N, O> r = a; r.v = a.v - b; return r; }
template <int N, int O> ILQC_HD Jet<N, O> operator-(double b, const Jet<N, O>& a) { Jet<N, O> r = -a; r.v = b - a.v; return r; }
template <int N, int O>
ILQC_HD Jet<N, O> operator*(const Jet<N, O>& a, double b) {
  Jet<N, O> r;
  r.v = a.v * b;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = a.g[i] * b;
  if constexpr (O >= 2) {
#pragma unroll
    for (int i = 0; i < Jet<N, O>::NH; ++i) r.h[i] = a.h[i] * b;
  }
  return r;
}
template <int N, int O> ILQC_HD Jet<N, O> operator*(double b, const Jet<N, O>& a) { return a * b; }
template <int N, int O>
ILQC_HD Jet<N, O> operator/(const Jet<N, O>& a, double b) {
  Jet<N, O> r;
  const double ib = 1.0 / b;
  r.v = a.v / b;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = a.g[i] * ib;
  if constexpr (O >= 2) {
#pragma unroll
    for (int i = 0; i < Jet<N, O>::NH; ++i) r.h[i] = a.h[i] * ib;
  }
  return r;
}
template <int N, int O>
ILQC_HD Jet<N, O> operator*(const Jet<N, O>& a, const Jet<N, O>& b) {
  return chain2(a, b, a.v * b.v, b.v, a.v, 0.0, 1.0, 0.0);
}
template <int N, int O>
ILQC_HD Jet<N, O> operator/(const Jet<N, O>& a, const Jet<N, O>& b) {
  const double ib = 1.0 / b.v;
  const double q = a.v / b.v;
  // q = a/b : q_a = 1/b, q_b = -q/b, q_aa = 0, q_ab = -1/b^2, q_bb = 2q/b^2
  return chain2(a, b, q, ib, -q * ib, 0.0, -ib * ib, 2.0 * q * ib * ib);
}
template <int N, int O>
ILQC_HD Jet<N, O> operator/(double a, const Jet<N, O>& b) {
  const double ib = 1.0 / b.v;
  const double q = a / b.v;
  return chain1(b, q, -q * ib, 2.0 * q * ib * ib);
}

// ---- elementary functions (double overloads first, so model code can be generic in S) ----------------
ILQC_HD double jsin(double a) { return ::sin(a); }
ILQC_HD double jcos(double a) { return ::cos(a); }
ILQC_HD double jtan(double a) { return ::tan(a); }
ILQC_HD double jatan(double a) { return ::atan(a); }
ILQC_HD double jatan2(double y, double x) { return ::atan2(y, x); }
ILQC_HD double jexp(double a) { return ::exp(a); }
ILQC_HD double jlog(double a) { return ::log(a); }
ILQC_HD double jsqrt(double a) { return ::sqrt(a); }
ILQC_HD double jsigmoid(double a) { return 1.0 / (1.0 + ::exp(-a)); }
ILQC_HD double jsq(double a) { return a * a; }             // torch x**2 == x*x
ILQC_HD double jcube(double a) { return (a * a) * a; }      // torch x**3 == x*x*x
ILQC_HD double jmax0(double a) { return a > 0.0 ? a : 0.0; }  // relu / maximum(x, 0)
ILQC_HD void jsincos(double a, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  ::sincos(a, s, c);
#else
  *s = ::sin(a);
  *c = ::cos(a);
#endif
}

template <int N, int O> ILQC_HD Jet<N, O> jsin(const Jet<N, O>& a) {
  double s, c; jsincos(a.v, &s, &c);
  return chain1(a, s, c, -s);
}
template <int N, int O> ILQC_HD Jet<N, O> jcos(const Jet<N, O>& a) {
  double s, c; jsincos(a.v, &s, &c);
  return chain1(a, c, -s, -c);
}
template <int N, int O> ILQC_HD Jet<N, O> jtan(const Jet<N, O>& a) {
  const double t = ::tan(a.v);
  const double d = 1.0 + t * t;
  return chain1(a, t, d, 2.0 * t * d);
}
template <int N, int O> ILQC_HD Jet<N, O> jatan(const Jet<N, O>& a) {
  const double d = 1.0 / (1.0 + a.v * a.v);
  return chain1(a, ::atan(a.v), d, -2.0 * a.v * d * d);
}
template <int N, int O> ILQC_HD Jet<N, O> jatan2(const Jet<N, O>& y, const Jet<N, O>& x) {
  const double ir2 = 1.0 / (x.v * x.v + y.v * y.v);
  const double fy = x.v * ir2, fx = -y.v * ir2;
  const double fyy = -2.0 * x.v * y.v * ir2 * ir2;
  const double fxy = (y.v * y.v - x.v * x.v) * ir2 * ir2;
  return chain2(y, x, ::atan2(y.v, x.v), fy, fx, fyy, fxy, -fyy);
}
template <int N, int O> ILQC_HD Jet<N, O> jexp(const Jet<N, O>& a) {
  const double e = ::exp(a.v);
  return chain1(a, e, e, e);
}
template <int N, int O> ILQC_HD Jet<N, O> jlog(const Jet<N, O>& a) {
  const double i = 1.0 / a.v;
  return chain1(a, ::log(a.v), i, -i * i);
}
template <int N, int O> ILQC_HD Jet<N, O> jsqrt(const Jet<N, O>& a) {
  const double s = ::sqrt(a.v);
  const double d = 0.5 / s;
  return chain1(a, s, d, -0.5 * d / a.v);
}
template <int N, int O> ILQC_HD Jet<N, O> jsigmoid(const Jet<N, O>& a) {
  const double s = 1.0 / (1.0 + ::exp(-a.v));
  const double d = s * (1.0 - s);
  return chain1(a, s, d, d * (1.0 - 2.0 * s));
}
template <int N, int O> ILQC_HD Jet<N, O> jsq(const Jet<N, O>& a) { return chain1(a, a.v * a.v, 2.0 * a.v, 2.0); }
template <int N, int O> ILQC_HD Jet<N, O> jcube(const Jet<N, O>& a) {
  return chain1(a, (a.v * a.v) * a.v, 3.0 * a.v * a.v, 6.0 * a.v);
}
template <int N, int O> ILQC_HD Jet<N, O> jmax0(const Jet<N, O>& a) {
  return (a.v > 0.0) ? a : make_cst<Jet<N, O>>(0.0);
}

}  // namespace ilqc

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
#include "fastmath.cuh"

namespace ilqc {

// (i, j) of the packed upper-triangle index e of a symmetric matrix of order N
template <int N>
struct SymTable {
  int i[sym_size(N) > 0 ? sym_size(N) : 1], j[sym_size(N) > 0 ? sym_size(N) : 1];
  constexpr SymTable() : i{}, j{} {
    int e = 0;
    for (int a = 0; a < N; ++a)
      for (int b = a; b < N; ++b) { i[e] = a; j[e] = b; ++e; }
  }
};
// NG > 1: the Hessian entries are split into NG groups of NHG consecutive packed entries and this jet carries
// group G only (value and gradient are complete).  Every Hessian entry propagates independently given the
// gradients, so NG threads with jets of NG-times fewer second-order components compute one full Hessian
// (hess_kernel: the 28-double jets of the bicycle model spill 9.6 KB per thread, the 14-double ones do not).
template <int N, int ORD, int NG = 1, int G = 0>
struct Jet {
  static constexpr int NV = N;
  static constexpr int NHG = (sym_size(N) + NG - 1) / NG;  // entries per group
  static constexpr int NH = (ORD >= 2) ? NHG : 1;
  static constexpr int E0 = G * NHG;                       // first packed entry of this group
  double v;
  double g[N];
  double h[NH];
  // packed index of local entry k, and whether it exists (the last group may be padded)
  static constexpr int entry(int k) { return E0 + k; }
  static constexpr bool valid(int k) { return E0 + k < sym_size(N); }
};

template <class S> struct is_jet : std::false_type {};
template <int N, int O, int NG, int G> struct is_jet<Jet<N, O, NG, G>> : std::true_type {};

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
template <int N, int O, int NG, int G> ILQC_HD double val(const Jet<N, O, NG, G>& a) { return a.v; }
ILQC_HD void set_val(double& a, double v) { a = v; }
template <int N, int O, int NG, int G> ILQC_HD void set_val(Jet<N, O, NG, G>& a, double v) { a.v = v; }
// autograd's .detach()/torch.tensor(x): keep the value, drop the derivatives (envs/car.py:262, :267)
template <class S> ILQC_HD S detach(const S& a) { return make_cst<S>(val(a)); }

// ---- generic chain rules -----------------------------------------------------------------------------
template <int N, int O, int NG, int G>
ILQC_HD Jet<N, O, NG, G> chain1(const Jet<N, O, NG, G>& a, double f0, double f1, double f2) {
  Jet<N, O, NG, G> r;
  r.v = f0;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = f1 * a.g[i];
  if constexpr (O >= 2) {
    using J = Jet<N, O, NG, G>;
    static_for<J::NH>([&](auto K) {
      constexpr int k = K;
      if constexpr (J::valid(k)) {
        constexpr int i = SymTable<N>().i[J::entry(k)], j = SymTable<N>().j[J::entry(k)];
        r.h[k] = f1 * a.h[k] + f2 * (a.g[i] * a.g[j]);
      } else {
        r.h[k] = 0.0;
      }
    });
  }
  return r;
}
template <int N, int O, int NG, int G>
ILQC_HD Jet<N, O, NG, G> chain2(const Jet<N, O, NG, G>& a, const Jet<N, O, NG, G>& b, double f0, double fa, double fb, double faa,
                         double fab, double fbb) {
  Jet<N, O, NG, G> r;
  r.v = f0;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = fa * a.g[i] + fb * b.g[i];
  if constexpr (O >= 2) {
    using J = Jet<N, O, NG, G>;
    static_for<J::NH>([&](auto K) {
      constexpr int k = K;
      if constexpr (J::valid(k)) {
        constexpr int i = SymTable<N>().i[J::entry(k)], j = SymTable<N>().j[J::entry(k)];
        r.h[k] = fa * a.h[k] + fb * b.h[k] + faa * (a.g[i] * a.g[j]) + fbb * (b.g[i] * b.g[j]) +
                 fab * (a.g[i] * b.g[j] + a.g[j] * b.g[i]);
      } else {
        r.h[k] = 0.0;
      }
    });
  }
  return r;
}

// ---- arithmetic --------------------------------------------------------------------------------------
template <int N, int O, int NG, int G>
ILQC_HD Jet<N, O, NG, G> operator+(const Jet<N, O, NG, G>& a, const Jet<N, O, NG, G>& b) {
  Jet<N, O, NG, G> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = a.g[i] + b.g[i];
  if constexpr (O >= 2) {
#pragma unroll
    for (int i = 0; i < Jet<N, O, NG, G>::NH; ++i) r.h[i] = a.h[i] + b.h[i];
  }
  return r;
}
template <int N, int O, int NG, int G>
ILQC_HD Jet<N, O, NG, G> operator-(const Jet<N, O, NG, G>& a, const Jet<N, O, NG, G>& b) {
  Jet<N, O, NG, G> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = a.g[i] - b.g[i];
  if constexpr (O >= 2) {
#pragma unroll
    for (int i = 0; i < Jet<N, O, NG, G>::NH; ++i) r.h[i] = a.h[i] - b.h[i];
  }
  return r;
}
template <int N, int O, int NG, int G>
ILQC_HD Jet<N, O, NG, G> operator-(const Jet<N, O, NG, G>& a) {
  Jet<N, O, NG, G> r;
  r.v = -a.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = -a.g[i];
  if constexpr (O >= 2) {
#pragma unroll
    for (int i = 0; i < Jet<N, O, NG, G>::NH; ++i) r.h[i] = -a.h[i];
  }
  return r;
}
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> operator+(const Jet<N, O, NG, G>& a, double b) { Jet<N, O, NG, G> r = a; r.v = a.v + b; return r; }
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> operator+(double b, const Jet<N, O, NG, G>& a) { Jet<N, O, NG, G> r = a; r.v = b + a.v; return r; }
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> operator-(const Jet<N, O, NG, G>& a, double b) { Jet<N, O, NG, G> r = a; r.v = a.v - b; return r; }
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> operator-(double b, const Jet<N, O, NG, G>& a) { Jet<N, O, NG, G> r = -a; r.v = b - a.v; return r; }
template <int N, int O, int NG, int G>
ILQC_HD Jet<N, O, NG, G> operator*(const Jet<N, O, NG, G>& a, double b) {
  Jet<N, O, NG, G> r;
  r.v = a.v * b;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = a.g[i] * b;
  if constexpr (O >= 2) {
#pragma unroll
    for (int i = 0; i < Jet<N, O, NG, G>::NH; ++i) r.h[i] = a.h[i] * b;
  }
  return r;
}
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> operator*(double b, const Jet<N, O, NG, G>& a) { return a * b; }
template <int N, int O, int NG, int G>
ILQC_HD Jet<N, O, NG, G> operator/(const Jet<N, O, NG, G>& a, double b) {
  Jet<N, O, NG, G> r;
  const double ib = 1.0 / b;
  r.v = a.v / b;
#pragma unroll
  for (int i = 0; i < N; ++i) r.g[i] = a.g[i] * ib;
  if constexpr (O >= 2) {
#pragma unroll
    for (int i = 0; i < Jet<N, O, NG, G>::NH; ++i) r.h[i] = a.h[i] * ib;
  }
  return r;
}
template <int N, int O, int NG, int G>
ILQC_HD Jet<N, O, NG, G> operator*(const Jet<N, O, NG, G>& a, const Jet<N, O, NG, G>& b) {
  return chain2(a, b, a.v * b.v, b.v, a.v, 0.0, 1.0, 0.0);
}
template <int N, int O, int NG, int G>
ILQC_HD Jet<N, O, NG, G> operator/(const Jet<N, O, NG, G>& a, const Jet<N, O, NG, G>& b) {
  const double ib = 1.0 / b.v;
  const double q = a.v / b.v;
  // q = a/b : q_a = 1/b, q_b = -q/b, q_aa = 0, q_ab = -1/b^2, q_bb = 2q/b^2
  return chain2(a, b, q, ib, -q * ib, 0.0, -ib * ib, 2.0 * q * ib * ib);
}
template <int N, int O, int NG, int G>
ILQC_HD Jet<N, O, NG, G> operator/(double a, const Jet<N, O, NG, G>& b) {
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

template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> jsin(const Jet<N, O, NG, G>& a) {
  double s, c; jsincos(a.v, &s, &c);
  return chain1(a, s, c, -s);
}
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> jcos(const Jet<N, O, NG, G>& a) {
  double s, c; jsincos(a.v, &s, &c);
  return chain1(a, c, -s, -c);
}
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> jtan(const Jet<N, O, NG, G>& a) {
  const double t = ::tan(a.v);
  const double d = 1.0 + t * t;
  return chain1(a, t, d, 2.0 * t * d);
}
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> jatan(const Jet<N, O, NG, G>& a) {
  const double d = 1.0 / (1.0 + a.v * a.v);
  return chain1(a, ::atan(a.v), d, -2.0 * a.v * d * d);
}
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> jatan2(const Jet<N, O, NG, G>& y, const Jet<N, O, NG, G>& x) {
  const double ir2 = 1.0 / (x.v * x.v + y.v * y.v);
  const double fy = x.v * ir2, fx = -y.v * ir2;
  const double fyy = -2.0 * x.v * y.v * ir2 * ir2;
  const double fxy = (y.v * y.v - x.v * x.v) * ir2 * ir2;
  return chain2(y, x, ::atan2(y.v, x.v), fy, fx, fyy, fxy, -fyy);
}
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> jexp(const Jet<N, O, NG, G>& a) {
  const double e = ::exp(a.v);
  return chain1(a, e, e, e);
}
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> jlog(const Jet<N, O, NG, G>& a) {
  const double i = 1.0 / a.v;
  return chain1(a, ::log(a.v), i, -i * i);
}
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> jsqrt(const Jet<N, O, NG, G>& a) {
  const double s = ::sqrt(a.v);
  const double d = 0.5 / s;
  return chain1(a, s, d, -0.5 * d / a.v);
}
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> jsigmoid(const Jet<N, O, NG, G>& a) {
  const double s = 1.0 / (1.0 + ::exp(-a.v));
  const double d = s * (1.0 - s);
  return chain1(a, s, d, d * (1.0 - 2.0 * s));
}
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> jsq(const Jet<N, O, NG, G>& a) { return chain1(a, a.v * a.v, 2.0 * a.v, 2.0); }
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> jcube(const Jet<N, O, NG, G>& a) {
  return chain1(a, (a.v * a.v) * a.v, 3.0 * a.v * a.v, 6.0 * a.v);
}
template <int N, int O, int NG, int G> ILQC_HD Jet<N, O, NG, G> jmax0(const Jet<N, O, NG, G>& a) {
  return (a.v > 0.0) ? a : make_cst<Jet<N, O, NG, G>>(0.0);
}

// ======================================================================================================
// Math context: the elementary functions the car models call, in two flavours.
//   MathCtx<true>   branch-free L-lane implementations of fastmath.cuh, valid on a stated argument range;
//                   `bad` accumulates "an argument was outside that range" and the caller then redoes the
//                   block with
//   MathCtx<false>  the CUDA math library / plain IEEE division (the arithmetic of every other model).
// Both act on doubles and on jets (chain rules on top of the value functions).
// ======================================================================================================
#ifndef ILQC_FASTMATH
#define ILQC_FASTMATH 1
#endif
#if defined(__CUDA_ARCH__) || !defined(__CUDACC__)
#define ILQC_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define ILQC_UNLIKELY(x) (x)
#endif

template <bool FAST>
struct MathCtx {
  static constexpr bool kFast = FAST;
  bool bad = false;

  // ---- value level -----------------------------------------------------------------------------------
  template <int L>
  ILQC_HD void sincos_v(const double (&a)[L], double (&s)[L], double (&c)[L]) {
    if constexpr (FAST) {
#pragma unroll
      for (int l = 0; l < L; ++l) bad |= !(::fabs(a[l]) < fm::kSinCosMaxArg);
      fm::sincos<L>(a, s, c);
    } else {
#pragma unroll
      for (int l = 0; l < L; ++l) { s[l] = ::sin(a[l]); c[l] = ::cos(a[l]); }
    }
  }
  template <int L>
  ILQC_HD void atan2_v(const double (&y)[L], const double (&x)[L], double (&o)[L]) {
    if constexpr (FAST) {
#pragma unroll
      for (int l = 0; l < L; ++l) bad |= fm::exp_out_of_range(::fabs(x[l]) + ::fabs(y[l]));
      fm::atan2<L>(y, x, o);
    } else {
#pragma unroll
      for (int l = 0; l < L; ++l) o[l] = ::atan2(y[l], x[l]);
    }
  }
  template <int L>
  ILQC_HD void atan_v(const double (&a)[L], double (&o)[L]) {
    if constexpr (FAST) {
#pragma unroll
      for (int l = 0; l < L; ++l) bad |= fm::too_big(a[l]);
      fm::atan<L>(a, o);
    } else {
#pragma unroll
      for (int l = 0; l < L; ++l) o[l] = ::atan(a[l]);
    }
  }
  ILQC_HD double rcp_v(double b) {
    if constexpr (FAST) {
      bad |= fm::exp_out_of_range(b);
      return fm::rcp(b);
    } else {
      return 1.0 / b;
    }
  }
  // q = a/b and ib = 1/b with one reciprocal
  ILQC_HD void divrcp_v(double a, double b, double* q, double* ib) {
    if constexpr (FAST) {
      bad |= fm::exp_out_of_range(b) | fm::too_big(a);
      const double r = fm::rcp_raw(b);
      *q = fm::div_with_rcp(a, b, r);
      *ib = fm::div_with_rcp(1.0, b, r);
    } else {
      *q = a / b;
      *ib = 1.0 / b;
    }
  }
  ILQC_HD double div_v(double a, double b) {
    if constexpr (FAST) {
      bad |= fm::exp_out_of_range(b) | fm::too_big(a);
      return fm::div(a, b);
    } else {
      return a / b;
    }
  }
  // a / b for a loop-invariant b in the safe exponent window with rb = 1/b precomputed (3 FMA-class
  // instructions, correctly rounded for finite a)
  ILQC_HD double divc_v(double a, double b, double rb) {
    if constexpr (FAST) return fm::div_with_rcp(a, b, rb);
    else return a / b;
  }

  // ---- generic in S (double or Jet) ------------------------------------------------------------------------
  template <class S, int L>
  ILQC_HD void sincos(const S (&a)[L], S (&s)[L], S (&c)[L]) {
    double av[L], sv[L], cv[L];
#pragma unroll
    for (int l = 0; l < L; ++l) av[l] = val(a[l]);
    sincos_v<L>(av, sv, cv);
#pragma unroll
    for (int l = 0; l < L; ++l) {
      if constexpr (is_jet<S>::value) {
        s[l] = chain1(a[l], sv[l], cv[l], -sv[l]);
        c[l] = chain1(a[l], cv[l], -sv[l], -cv[l]);
      } else {
        s[l] = sv[l];
        c[l] = cv[l];
      }
    }
  }
  template <class S, int L>
  ILQC_HD void atan2(const S (&y)[L], const S (&x)[L], S (&o)[L]) {
    double yv[L], xv[L], ov[L];
#pragma unroll
    for (int l = 0; l < L; ++l) { yv[l] = val(y[l]); xv[l] = val(x[l]); }
    atan2_v<L>(yv, xv, ov);
#pragma unroll
    for (int l = 0; l < L; ++l) {
      if constexpr (is_jet<S>::value) {
        const double ir2 = rcp_v(xv[l] * xv[l] + yv[l] * yv[l]);
        const double fy = xv[l] * ir2, fx = -yv[l] * ir2;
        const double fyy = -2.0 * xv[l] * yv[l] * ir2 * ir2;
        const double fxy = (yv[l] * yv[l] - xv[l] * xv[l]) * ir2 * ir2;
        o[l] = chain2(y[l], x[l], ov[l], fy, fx, fyy, fxy, -fyy);
      } else {
        o[l] = ov[l];
      }
    }
  }
  template <class S, int L>
  ILQC_HD void atan(const S (&a)[L], S (&o)[L]) {
    double av[L], ov[L];
#pragma unroll
    for (int l = 0; l < L; ++l) av[l] = val(a[l]);
    atan_v<L>(av, ov);
#pragma unroll
    for (int l = 0; l < L; ++l) {
      if constexpr (is_jet<S>::value) {
        const double d = rcp_v(1.0 + av[l] * av[l]);
        o[l] = chain1(a[l], ov[l], d, -2.0 * av[l] * d * d);
      } else {
        o[l] = ov[l];
      }
    }
  }
  template <class S>
  ILQC_HD S div(const S& a, const S& b) {
    if constexpr (is_jet<S>::value) {
      double q, ib;
      divrcp_v(a.v, b.v, &q, &ib);
      return chain2(a, b, q, ib, -q * ib, 0.0, -ib * ib, 2.0 * q * ib * ib);
    } else {
      return div_v(a, b);
    }
  }
  template <class S>
  ILQC_HD S div_cst_by(double a, const S& b) {
    if constexpr (is_jet<S>::value) {
      double q, ib;
      divrcp_v(a, b.v, &q, &ib);
      return chain1(b, q, -q * ib, 2.0 * q * ib * ib);
    } else {
      return div_v(a, b);
    }
  }
  template <class S>
  ILQC_HD S divc(const S& a, double b, double rb) {
    if constexpr (is_jet<S>::value) {
      S r;
      r.v = divc_v(a.v, b, rb);
#pragma unroll
      for (int i = 0; i < S::NV; ++i) r.g[i] = a.g[i] * rb;
#pragma unroll
      for (int i = 0; i < S::NH; ++i) r.h[i] = a.h[i] * rb;
      return r;
    } else {
      return divc_v(a, b, rb);
    }
  }
  // torch.sigmoid: 1 / (1 + exp(-a))
  template <class S>
  ILQC_HD S sigmoid(const S& a) {
    const double s = rcp_v(1.0 + ::exp(-val(a)));
    if constexpr (is_jet<S>::value) {
      const double d = s * (1.0 - s);
      return chain1(a, s, d, d * (1.0 - 2.0 * s));
    } else {
      return s;
    }
  }
  // v = sqrt(w) and iv = 1/sqrt(w)
  template <class S>
  ILQC_HD void sqrt_inv(const S& w, S* v, S* iv) {
    double s, is;
    if constexpr (FAST) {
      bad |= fm::exp_out_of_range(val(w)) | (val(w) < 0.0);
      fm::sqrt_rsqrt(val(w), &s, &is);
    } else {
      s = ::sqrt(val(w));
      is = 1.0 / s;
    }
    if constexpr (is_jet<S>::value) {
      const double is2 = is * is;
      *v = chain1(w, s, 0.5 * is, (-0.25 * is) * is2);
      *iv = chain1(w, is, (-0.5 * is) * is2, ((0.75 * is) * is2) * is2);
    } else {
      *v = s;
      *iv = is;
    }
  }
  template <class S>
  ILQC_HD S log(const S& a) {
    if constexpr (is_jet<S>::value) {
      const double i = rcp_v(a.v);
      return chain1(a, ::log(a.v), i, -i * i);
    } else {
      return ::log(a);
    }
  }
};
#if ILQC_FASTMATH
using MathHot = MathCtx<true>;
#else
using MathHot = MathCtx<false>;
#endif
using MathExact = MathCtx<false>;

}  // namespace ilqc

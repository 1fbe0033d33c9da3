// Shared macros / compile-time helpers for the ilqc_b200 kernels.
//
// Every numerical routine is written as an ILQC_HD function of the *problem slot* it serves, so that the
// __global__ kernels are one-line wrappers (thread-per-problem mapping) and the very same templates can
// be instantiated by the test-only host simulator (tests/hostsim, -DILQC_HOSTSIM) to debug the math in
// a container without a GPU.  The product library never contains the host simulator.
#pragma once
#include <cmath>
#include <cstdint>
#include <utility>
#include <type_traits>

#if defined(__CUDACC__)
#define ILQC_HD __host__ __device__ __forceinline__
#define ILQC_D __device__ __forceinline__
#else
#define ILQC_HD inline __attribute__((always_inline))
#define ILQC_D inline __attribute__((always_inline))
#endif

#if defined(__CUDA_ARCH__)
#define ILQC_LDG(p) __ldg(p)
#else
#define ILQC_LDG(p) (*(p))
#endif

namespace ilqc {

// ---- compile-time loop: static_for<0,N>([&](auto I){ constexpr int i = I; ... }) --------------------
template <int... Is, class F>
ILQC_HD void static_for_impl(std::integer_sequence<int, Is...>, F&& f) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
ILQC_HD void static_for(F&& f) {
  if constexpr (N > 0) static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

// packed index of the symmetric pair (i,j) in a row-major upper triangle of order n
ILQC_HD constexpr int sym_idx(int i, int j, int n) {
  return (i <= j) ? (i * n - (i * (i - 1)) / 2 + (j - i)) : (j * n - (j * (j - 1)) / 2 + (i - j));
}
ILQC_HD constexpr int sym_size(int n) { return n * (n + 1) / 2; }

}  // namespace ilqc

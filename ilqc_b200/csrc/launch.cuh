// Launch plumbing shared by all translation units.
//
// Product build (nvcc, sm_100a): ILQC_LAUNCH enqueues a __global__ kernel, one thread per slot, on the
// caller's stream.  Test-only build (-DILQC_HOSTSIM, g++): the same "kernel" functions are called in a
// loop on the CPU so that the container without a GPU can check the arithmetic against the golden
// vectors (tests/hostsim).  The host simulator is never part of libilqc_b200.so.
#pragma once
#include <cstdlib>
#include <cstring>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include "common.cuh"
#include "../../include/ilqc_b200.h"

#if defined(ILQC_HOSTSIM)
namespace ilqc {
extern thread_local int g_sim_tid;
typedef void* stream_t;
inline long long& launch_counter() { static long long n = 0; return n; }
}
#define ILQC_GLOBAL static void
#define ILQC_GLOBAL_LB(minblocks) static void
#define ILQC_TID (::ilqc::g_sim_tid)
#define ILQC_LOCAL_TID (::ilqc::g_sim_tid)
#define ILQC_BLOCK_BASE 0
#define ILQC_UNPAREN(...) __VA_ARGS__
#define ILQC_LAUNCH(kern, n, block, stream, ...)                 \
  do {                                                           \
    for (int _i = 0; _i < (int)(n); ++_i) {                      \
      ::ilqc::g_sim_tid = _i;                                    \
      ILQC_UNPAREN kern(__VA_ARGS__);                            \
    }                                                            \
  } while (0)
#define ILQC_LAUNCH_SMEM(kern, n, block, smem, stream, ...) ILQC_LAUNCH(kern, n, block, stream, __VA_ARGS__)
#define ILQC_LAUNCH_OK() 0
#else
#include <cuda_runtime.h>
namespace ilqc {
typedef cudaStream_t stream_t;
// number of kernels this library has launched (ilqc_launch_count; bench.py reports it as gpu_launches)
inline long long& launch_counter() { static long long n = 0; return n; }
}
#define ILQC_GLOBAL __global__ void
#define ILQC_GLOBAL_LB(minblocks) __global__ void __launch_bounds__(ILQC_BLOCK, minblocks)
#define ILQC_TID ((int)(blockIdx.x * blockDim.x + threadIdx.x))
#define ILQC_LOCAL_TID ((long long)threadIdx.x)
#define ILQC_BLOCK_BASE ((long long)blockIdx.x * blockDim.x)
#define ILQC_UNPAREN(...) __VA_ARGS__
#define ILQC_LAUNCH(kern, n, block, stream, ...)                                                   \
  do {                                                                                             \
    if ((n) > 0) {                                                                                 \
      ILQC_UNPAREN kern<<<((n) + (block)-1) / (block), (block), 0, (stream)>>>(__VA_ARGS__);        \
      ++::ilqc::launch_counter();                                                                  \
    }                                                                                              \
  } while (0)
// same with dynamic shared memory (opt-in above 48 KB is requested once per kernel)
#define ILQC_LAUNCH_SMEM(kern, n, block, smem, stream, ...)                                                   \
  do {                                                                                                         \
    if ((n) > 0) {                                                                                             \
      static bool _attr_set[64] = {};  /* the attribute is per device */                                      \
      int _dev = 0;                                                                                            \
      cudaGetDevice(&_dev);                                                                                    \
      if ((smem) > 0 && !_attr_set[_dev & 63]) {                                                               \
        cudaFuncSetAttribute(ILQC_UNPAREN kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem));     \
        _attr_set[_dev & 63] = true;                                                                           \
      }                                                                                                        \
      ILQC_UNPAREN kern<<<((n) + (block)-1) / (block), (block), (smem), (stream)>>>(__VA_ARGS__);               \
      ++::ilqc::launch_counter();                                                                              \
    }                                                                                                          \
  } while (0)
#define ILQC_LAUNCH_OK() ((int)cudaGetLastError())
#endif

namespace ilqc {

// thread-block size of the thread-per-problem kernels.  The big bodies use ~255 registers, so an SM holds
// 256 threads; small blocks keep all 148 SMs busy for batches that are not a multiple of 148*256.
#ifndef ILQC_BLOCK
#define ILQC_BLOCK 64
#endif
constexpr int kBlock = ILQC_BLOCK;
// Occupancy targets (resident blocks per SM) of the three heavy kernel families; they set the register
// budget ptxas works with: 65536 / (ILQC_BLOCK * blocks) registers per thread.  Tuned on a B200 with
// tools/tune_kernels.py (profiles/r01_tuning.txt).
#ifndef ILQC_LB_FWD1
#define ILQC_LB_FWD1 4
#endif
#ifndef ILQC_LB_BWD
#define ILQC_LB_BWD 4
#endif
#ifndef ILQC_LB_ROLL
#define ILQC_LB_ROLL 8
#endif

// launches of at most this many candidates use the warp-cooperative backward pass (coop.cuh)
#ifndef ILQC_COOP_MAX_DEFAULT
#define ILQC_COOP_MAX_DEFAULT 0
#endif
// engine tuning knobs (ilqc_set_tuning; initial values from the environment: ILQC_COOP_MAX, ILQC_FEED_SMALL)
struct Tuning {
  long long coop_max, feed_small;
};
inline Tuning tuning_defaults() {
  Tuning x;
  const char* e = getenv("ILQC_COOP_MAX");
  x.coop_max = e ? atoll(e) : (long long)ILQC_COOP_MAX_DEFAULT;
  e = getenv("ILQC_FEED_SMALL");
  x.feed_small = e ? atoll(e) : 0LL;
  return x;
}
inline Tuning& tuning() {
  static Tuning t = tuning_defaults();
  return t;
}
// internal flag or-ed into the backward mode: the slots map to (nearly) consecutive problems (all problems
// active, candidates of a problem adjacent) -> the shared-memory operand feed pays off
#define ILQC_BWD_DENSE_HINT 0x100
// bits 4-5 of the backward mode: handle_bad_dir (ILQC_BAD_DIR_*)
#define ILQC_BWD_BAD_DIR_SHIFT 4

// Per-model launchers, defined once per model in model_kernels.inl (one translation unit per model so
// that the heavy templates compile in parallel).
template <int MODEL>
struct Launch {
  // n_threads threads; thread a serves problem prob[a] (prob == NULL: a, n_threads == B)
  static int forward(int order, const ilqc_model_desc& m, int B, int n_threads, const int32_t* prob,
                     const double* x0, const double* cmd, double* lin, double* traj, double* cost_t, double* total,
                     double* cnorm, stream_t st);
  // time-parallel expansion (ilqc_algo_desc.expansion = 1): gsq = scratch [H][B]
  static int expand_parallel(const ilqc_model_desc& m, int B, int n_probs, const int32_t* prob, const double* x0,
                             const double* cmd, double* lin, double* traj, double* gsq, double* cnorm, stream_t st);
  static int expand_dense(const ilqc_model_desc& m, int B, const double* x0, const double* cmd, int order,
                          double* A, double* Bm, double* p, double* P, double* q, double* Q, const double* lam,
                          double* Hxx, double* Hxu, double* Huu, stream_t st);
  // n_threads threads; thread s serves problem prob[s] (NULL: s) and reads/writes the per-slot arrays
  // (reg_ctrl, gains, jcst, feasible / diff, newval, diffsq) at index oslot[s] (NULL: s) with slot
  // stride n_oslots; roll-outs read their gains at gslot[s] (NULL: s) with stride n_gslots.
  static int backward(int mode, const ilqc_model_desc& m, int B, int n_threads, const int32_t* prob,
                      const int32_t* oslot, int n_oslots, const double* reg_ctrl, double reg_state,
                      const double* lin, const double* traj, const double* cmd, const double* hess, double* gains,
                      double* jcst, int32_t* feasible, stream_t st);
  // second-order tensor of the dynamics of the problems prob[0..n_probs) (NULL: all), [H][hess_stride][B]
  static int hess(const ilqc_model_desc& m, int B, int n_probs, const int32_t* prob, const double* traj,
                  const double* cmd, double* hess, stream_t st);
  static int rollout(int kind, const ilqc_model_desc& m, int B, int n_threads, const int32_t* prob,
                     const int32_t* gslot, int n_gslots, const int32_t* oslot, int n_oslots, const double* step,
                     const double* x0, const double* cmd, const double* traj, const double* lin,
                     const double* gains, const double* dir, double* diff, double* newval, double* diffsq,
                     stream_t st);
  static int adjoint(const ilqc_model_desc& m, int B, int n_threads, const int32_t* prob, const double* lin,
                     double* grad, double* gnorm, stream_t st);
  static int cnorm_lin(const ilqc_model_desc& m, int B, const double* lin, double* cnorm, stream_t st);
  static void dims(int* nx, int* nu, int* lin_stride, int* gain_stride, int* hess_stride);
};

}  // namespace ilqc

// TEST INFRASTRUCTURE ONLY -- CPU emulation of the kernels of ilqc_b200/csrc.
// The very same templates (bodies.cuh, riccati.cuh, solver.inl, api.inl) are compiled with g++ and
// -DILQC_HOSTSIM: every "kernel" becomes a loop over its slots.  It lets the `-m "not gpu"` tests check the
// arithmetic of the CUDA code against the golden vectors in a container without a GPU.  The product
// package (ilqc_b200/_lib.py) never loads this library and fails loudly without the CUDA one.
#include "../../ilqc_b200/csrc/api.inl"
namespace ilqc { thread_local int g_sim_tid = 0; }
extern "C" int ilqc_fp64_peak(int, double*, void*) { return ILQC_E_UNSUPPORTED; }
extern "C" int ilqc_is_hostsim(void) { return 1; }

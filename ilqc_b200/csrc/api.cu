// libilqc_b200.so: C-ABI entry points + solver driver (product build, nvcc, sm_100a).
#include "api.inl"

namespace ilqc {
// FP64 FMA micro-benchmark: 8 independent dependency chains per thread, 2 flop per DFMA.  Gives the
// measured FP64 roofline denominator on the box the bench runs on (BASELINE.md section 3 asks for it).
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace ilqc

extern "C" int ilqc_fp64_peak(int iters, double* flops_per_s, void* stream) {
  if (iters <= 0 || !flops_per_s) return ILQC_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int threads = 256, blocks = sms * 8;
  double* buf = nullptr;
  cudaError_t e = cudaMalloc(&buf, sizeof(double) * threads * blocks);
  if (e != cudaSuccess) return (int)e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  fp64_peak_kernel<<<blocks, threads, 0, st>>>(buf, iters / 8 + 1, 0.999999, 1e-9);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, st);
    fp64_peak_kernel<<<blocks, threads, 0, st>>>(buf, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16.0 * 8.0 * (double)iters * threads * blocks;
    const double rate = flops / (ms * 1e-3);
    if (rate > best) best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *flops_per_s = best;
  return (int)cudaGetLastError();
}

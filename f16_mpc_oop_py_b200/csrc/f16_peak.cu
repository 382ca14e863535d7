// f16_peak.cu -- FP64 FMA micro-benchmark: the measured denominator of the FP64-pipe roofline.
// MEASURED_PEAKS.json carries HBM and bf16 numbers only, and this path is bound by the FP64 pipe
// (SURVEY.md 8d), so the library measures the sustained DFMA rate itself: 8 independent register-only
// FMA chains per thread, 1024 threads per CTA, two CTAs per SM.
#include "f16_kernels.cuh"

namespace f16 {

__global__ void __launch_bounds__(1024, 2) dfma_peak_kernel(long long iters, double seed, double* sink) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (long long i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.678) sink[0] = s;  // keeps the chains alive, practically never taken
}

cudaError_t launch_dfma_peak(cudaStream_t stream, int sm_count, long long iters, double* sink, double* flops) {
  const int blocks = sm_count * 2, threads = 1024;
  dfma_peak_kernel<<<blocks, threads, 0, stream>>>(iters, 1.0, sink);
  *flops = (double)blocks * threads * (double)iters * 64.0 * 2.0;
  return cudaGetLastError();
}

}  // namespace f16

// mb_issue.cu -- does a DFMA (half-rate pipe: 16 lanes per sub-partition) leave the second issue cycle free for another
// pipe on B200?  Per-iteration cost of D DFMAs + I integer ops (+ F FP32 ops), 8 warps per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

template <int D, int I, int F, int SETP>
__global__ void mix_kernel(int iters, long long* cycles, double* sink) {
  double a[8];
  int k[8];
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { a[j] = threadIdx.x * 1e-3 + j; k[j] = threadIdx.x + j; f[j] = threadIdx.x * 0.5f + j; }
  const double b = 1.0000001, c = 1e-9;
  int cnt = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int d = 0; d < D; d++) a[(u + d) & 7] = fma(a[(u + d) & 7], b, c);
#pragma unroll
      for (int d = 0; d < SETP; d++) cnt += (a[(u + d) & 7] > 1.5 + d) ? 1 : 0;
#pragma unroll
      for (int d = 0; d < I; d++) k[(u + d) & 7] = (k[(u + d) & 7] ^ (i + d)) + 0x9e3779b9;
#pragma unroll
      for (int d = 0; d < F; d++) f[(u + d) & 7] = fmaf(f[(u + d) & 7], 1.0001f, 0.5f);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  double s = 0; int ks = cnt; float fs = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) { s += a[j]; ks += k[j]; fs += f[j]; }
  if (s == 1.2345 || ks == 12345 || fs == 1.25f) sink[0] = s;
}

template <int D, int I, int F, int SETP>
void run(const char* name, long long* d_c, double* d_s) {
  const int iters = 2000, warps = 32;
  mix_kernel<D, I, F, SETP><<<1, warps * 32>>>(iters, d_c, d_s);
  long long c = 0;
  cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
  // per sub-partition: warps/4 warps each executing iters*8 groups
  printf("%-34s %.2f issue cycles per group per warp (per sub-partition)\n", name, (double)c / (iters * 8.0 * (warps / 4)));
}

int main() {
  long long* d_c; double* d_s;
  cudaMalloc(&d_c, 1024 * 8); cudaMalloc(&d_s, 64);
  run<1, 0, 0, 0>("1 DFMA", d_c, d_s);
  run<2, 0, 0, 0>("2 DFMA", d_c, d_s);
  run<0, 1, 0, 0>("1 INT(2 ops: LOP3+IADD)", d_c, d_s);
  run<0, 2, 0, 0>("2 INT", d_c, d_s);
  run<1, 1, 0, 0>("1 DFMA + 1 INT", d_c, d_s);
  run<2, 1, 0, 0>("2 DFMA + 1 INT", d_c, d_s);
  run<2, 2, 0, 0>("2 DFMA + 2 INT", d_c, d_s);
  run<0, 0, 1, 0>("1 FFMA", d_c, d_s);
  run<1, 0, 1, 0>("1 DFMA + 1 FFMA", d_c, d_s);
  run<1, 0, 2, 0>("1 DFMA + 2 FFMA", d_c, d_s);
  run<2, 0, 2, 0>("2 DFMA + 2 FFMA", d_c, d_s);
  run<2, 1, 1, 0>("2 DFMA + 1 INT + 1 FFMA", d_c, d_s);
  run<0, 0, 0, 1>("1 DSETP(+sel/add)", d_c, d_s);
  run<1, 0, 0, 1>("1 DFMA + 1 DSETP(+sel/add)", d_c, d_s);
  run<2, 0, 0, 2>("2 DFMA + 2 DSETP(+sel/add)", d_c, d_s);
  return 0;
}

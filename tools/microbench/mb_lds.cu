// mb_lds.cu -- how many shared-memory wavefronts does a warp-wide LDS.128 / LDS.64 cost on B200 when the lanes read
// 1, 2, 4, 8 ... distinct addresses (the table-gather pattern of the F-16 step kernel)?  And the dependent-issue
// latency of DFMA.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_lds mb_lds.cu ; run on one GPU.
#include <cstdio>
#include <cuda_runtime.h>

extern __shared__ __align__(16) double sm[];

// pattern: lane l reads 16 B at (group(l) * stride_bytes); groups = number of distinct addresses
template <int VEC>
__global__ void lds_kernel(int groups, int stride_dbl, int interleave, int iters, long long* cycles, double* sink) {
  for (int i = threadIdx.x; i < 20000; i += blockDim.x) sm[i] = i * 0.5;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int g = interleave ? (lane % groups) : (lane * groups / 32);
  const double* p = sm + g * stride_dbl;
  double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const double* q = p + u * 2 + (it & 1) * 64;
      if (VEC == 2) {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(q)));
        if (u & 1) { acc0 += v.x; acc1 += v.y; } else { acc2 += v.x; acc3 += v.y; }
      } else {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(q)));
        if (u & 1) acc0 += v; else acc2 += v;
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc0 + acc1 + acc2 + acc3 == 1.2345) sink[0] = acc0;
}

__global__ void dfma_latency(int iters, long long* cycles, double* sink) {
  double a = threadIdx.x * 1e-3, b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 32; u++) a = fma(a, b, c);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (a == 1.2345) sink[0] = a;
}

int main() {
  long long* d_c; double* d_s;
  cudaMalloc(&d_c, 1024 * 8); cudaMalloc(&d_s, 64);
  cudaFuncSetAttribute(lds_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
  cudaFuncSetAttribute(lds_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
  const int iters = 2000;
  printf("# LDS throughput: cycles per warp-wide load instruction per SM (all warps of one CTA issuing), 1 CTA\n");
  for (int vec = 2; vec >= 1; vec--)
    for (int warps : {1, 4, 8, 16})
      for (int inter = 0; inter < 2; inter++)
        for (int groups : {1, 2, 4, 8, 32})
          for (int stride : {2, 42, 546, 16}) {  // doubles: adjacent 16 B, next alpha cell, next beta node, 128 B
            if (groups == 1 && (stride != 2 || inter)) continue;
            if (vec == 2) lds_kernel<2><<<1, warps * 32, 200000>>>(groups, stride, inter, iters, d_c, d_s);
            else lds_kernel<1><<<1, warps * 32, 200000>>>(groups, stride, inter, iters, d_c, d_s);
            long long c = 0;
            cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            printf("LDS.%d warps=%2d %s groups=%2d stride=%4d B : %.2f cyc per load instr (per SM, all warps)\n", vec * 64, warps,
                   inter ? "interleaved" : "blocked    ", groups, stride * 8, (double)c / (iters * 16.0 * warps));
          }
  for (int warps : {1, 2, 4}) {
    dfma_latency<<<1, warps * 32>>>(2000, d_c, d_s);
    long long c = 0;
    cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent chain, %d warp(s) on the SM (one per sub-partition up to 4): %.2f cyc per DFMA\n", warps, (double)c / (2000 * 32.0));
  }
  for (int warps : {8, 16}) {
    dfma_latency<<<1, warps * 32>>>(2000, d_c, d_s);
    long long c = 0;
    cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent chain, %d warps on the SM: %.2f cyc per DFMA per warp\n", warps, (double)c / (2000 * 32.0));
  }
  return 0;
}

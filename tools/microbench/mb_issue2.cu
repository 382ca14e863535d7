// mb_issue2.cu -- issue-slot cost model of the FP64-bound step kernel on B200: cycles per warp per sub-partition of
// instruction groups built from inline PTX (8 warps per sub-partition, independent chains).
#include <cstdio>
#include <cuda_runtime.h>

extern __shared__ __align__(16) double sm[];

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int KIND>
__global__ void k(int iters, long long* cycles, double* sink) {
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 0.5 + 1.0;
  __syncthreads();
  double a[8], m[8];
  int q[8];
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { a[j] = threadIdx.x * 1e-3 + j; m[j] = 1.0 + j * 1e-9; q[j] = threadIdx.x + j; f[j] = j; }
  const double b = 1.0000001, c = 1e-9;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 3) * 4368;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      // every group starts with one DFMA
      if (KIND < 20) a[u] = fma(a[u], b, c);
      if (KIND == 20) a[u] = fma(a[u], m[u], m[(u + 1) & 7]);                                         // DFMA, 3 register sources
      if (KIND == 21) a[u] = fma(a[u], m[u], c);                                                      // DFMA, 2 register sources
      if (KIND == 22) { a[u] = fma(a[u], m[u], m[(u + 1) & 7]); q[u] = (q[u] ^ i); }                  // 3-reg DFMA + LOP3
      if (KIND == 23) { a[u] = fma(a[u], m[u], m[(u + 1) & 7]); q[u] = (q[u] ^ i) + 0x9e3779b9; }     // 3-reg DFMA + LOP3 + IADD
      if (KIND == 24) { a[u] = fma(a[u], m[u], m[(u + 1) & 7]); m[u] = (q[u] > i) ? m[u] : a[(u + 3) & 7]; }  // 3-reg DFMA + ISETP + 2 SEL
      if (KIND == 25) { a[u] = fma(a[u], m[u], m[(u + 1) & 7]); a[(u+4)&7] = a[(u+4)&7] * m[(u + 2) & 7]; }  // 3-reg DFMA + 2-reg DMUL
      if (KIND == 1) asm volatile("mul.f64 %0, %0, %1;" : "+d"(m[u]) : "d"(b));                      // DFMA + DMUL
      if (KIND == 2) asm volatile("add.f64 %0, %0, %1;" : "+d"(m[u]) : "d"(c));                      // DFMA + DADD
      if (KIND == 3) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sbase + u * 8 + (i & 7) * 64)); m[u] += v; }   // + LDS.64 (+DADD)
      if (KIND == 4) { double v, w; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v), "=d"(w) : "r"(sbase + u * 16 + (i & 7) * 128)); m[u] += v; m[(u + 1) & 7] += w; }  // + LDS.128 (+2 DADD)
      if (KIND == 6) q[u] = (q[u] ^ i) + 0x9e3779b9;                                                 // + LOP3 + IADD
      if (KIND == 7) q[u] = q[u] * 3 + i;                                                            // + IMAD
      if (KIND == 8) f[u] = fmaf(f[u], 1.0001f, 0.5f);                                               // + FFMA
      if (KIND == 9) { q[u] = q[u] * 3 + i; f[u] = fmaf(f[u], 1.0001f, 0.5f); }                      // + IMAD + FFMA
      if (KIND == 10) { q[u] = (q[u] ^ i); }                                                         // + LOP3
      if (KIND == 11) { q[u] = max(q[u], i) ; }                                                      // + IMNMX
      if (KIND == 12) { m[u] = (q[u] > i) ? m[u] : a[(u + 3) & 7]; }                                 // + ISETP + 2 SEL
      if (KIND == 90) asm volatile("mul.f64 %0, %0, %1;" : "+d"(m[u]) : "d"(b));                      // DMUL only
      if (KIND == 91) asm volatile("add.f64 %0, %0, %1;" : "+d"(m[u]) : "d"(c));                      // DADD only
      if (KIND == 92) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sbase + u * 8 + (i & 7) * 64)); q[u] ^= __double2loint(v); }  // LDS.64 + LOP3
      if (KIND == 93) { double v, w; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v), "=d"(w) : "r"(sbase + u * 16 + (i & 7) * 128)); q[u] ^= __double2loint(v) ^ __double2hiint(w); }  // LDS.128 + LOP3
      if (KIND == 94) { int p; asm volatile("{ .reg .pred p; setp.gt.f64 p, %1, %2; selp.s32 %0, 1, 0, p; }" : "=r"(p) : "d"(a[u]), "d"(m[u])); q[u] += p; }  // DSETP + SEL + IADD
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  double s = 0; int ks = 0; float fs = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) { s += a[j] + m[j]; ks += q[j]; fs += f[j]; }
  if (s == 1.2345 || ks == 12345 || fs == 1.25f) sink[0] = s;
}

template <int KIND>
void run(const char* name, long long* d_c, double* d_s) {
  const int iters = 2000, warps = 32;
  k<KIND><<<1, warps * 32, 40000>>>(iters, d_c, d_s);
  long long c = 0;
  cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("%-40s %.2f cycles per group per warp (per sub-partition)%s\n", name, (double)c / (iters * 8.0 * (warps / 4)), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_c; double* d_s;
  cudaMalloc(&d_c, 1024 * 8); cudaMalloc(&d_s, 64);
  run<0>("DFMA", d_c, d_s);
  run<21>("DFMA 2 reg sources", d_c, d_s);
  run<20>("DFMA 3 reg sources", d_c, d_s);
  run<22>("DFMA(3 reg) + LOP3", d_c, d_s);
  run<23>("DFMA(3 reg) + LOP3 + IADD", d_c, d_s);
  run<24>("DFMA(3 reg) + ISETP + 2 SEL", d_c, d_s);
  run<25>("DFMA(3 reg) + DMUL(2 reg)", d_c, d_s);
  run<90>("DMUL", d_c, d_s);
  run<91>("DADD", d_c, d_s);
  run<94>("DSETP + SEL + IADD", d_c, d_s);
  run<92>("LDS.64(4 addr) + LOP3", d_c, d_s);
  run<93>("LDS.128(4 addr) + 2 LOP3", d_c, d_s);
  run<1>("DFMA + DMUL", d_c, d_s);
  run<2>("DFMA + DADD", d_c, d_s);
  run<3>("DFMA + LDS.64(4 addr) + DADD", d_c, d_s);
  run<4>("DFMA + LDS.128(4 addr) + 2 DADD", d_c, d_s);
  run<6>("DFMA + LOP3 + IADD", d_c, d_s);
  run<10>("DFMA + LOP3", d_c, d_s);
  run<7>("DFMA + IMAD", d_c, d_s);
  run<8>("DFMA + FFMA", d_c, d_s);
  run<9>("DFMA + IMAD + FFMA", d_c, d_s);
  run<11>("DFMA + IMNMX", d_c, d_s);
  run<12>("DFMA + ISETP + 2 SEL", d_c, d_s);
  return 0;
}

// f16_partition.cu -- a mixed batch partitioned by fidelity before the fused step ("batch partitioned by
// fidelity first so warps stay uniform").  The step kernels take one aircraft per thread; with per-aircraft fidelity flags
// in arbitrary order a warp would hold both models and each of the two launches would idle the other model's lanes for
// all K steps.  Instead: a stable three-way partition of the aircraft indices (hifi | lofi | invalid flag), one gather of
// the 22 state / input planes into that order, the two launches on contiguous sub-ranges with every lane busy, and one
// scatter back.  The reorder moves 2 x 320 B per aircraft once per call -- HBM-bound, a few steps' worth of time.
#include <stdint.h>

#include "f16_kernels.cuh"

namespace f16 {
namespace partition {

constexpr int THREADS = 1024;

__device__ __forceinline__ int class_of(unsigned char f) { return f == 1 ? 0 : (f == 0 ? 1 : 2); }  // hifi, lofi, invalid

// per-CTA class counts: cta_counts[b][3]
__global__ void __launch_bounds__(THREADS)
count_kernel(const unsigned char* __restrict__ fi, long long N, unsigned* __restrict__ cta_counts) {
  __shared__ unsigned wc[THREADS / 32][3];
  const long long n = (long long)blockIdx.x * THREADS + threadIdx.x;
  const int c = n < N ? class_of(fi[n]) : 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const unsigned m = __ballot_sync(0xffffffffu, c == k);
    if (lane == 0) wc[warp][k] = __popc(m);
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    unsigned s = 0;
    for (int w = 0; w < THREADS / 32; w++) s += wc[w][threadIdx.x];
    cta_counts[(long long)blockIdx.x * 3 + threadIdx.x] = s;
  }
}

// one CTA: exclusive scan of the per-CTA counts, class by class; cta_off[b][k] = position of CTA b's first element of
// class k in the permutation; totals[3] = class sizes
__global__ void __launch_bounds__(THREADS)
scan_kernel(const unsigned* __restrict__ cta_counts, int n_cta, unsigned* __restrict__ cta_off, long long* __restrict__ totals) {
  __shared__ unsigned long long part[THREADS];
  __shared__ unsigned long long class_base[4];
  // class totals first (so that class k starts after classes < k)
  for (int k = 0; k < 3; k++) {
    unsigned long long s = 0;
    for (int b = threadIdx.x; b < n_cta; b += THREADS) s += cta_counts[(long long)b * 3 + k];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = THREADS / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      totals[k] = (long long)part[0];
      class_base[k + 1] = (k == 0 ? 0 : class_base[k]) + part[0];
      if (k == 0) class_base[0] = 0;
    }
    __syncthreads();
  }
  // exclusive scan per class: every thread owns a contiguous chunk of CTAs
  const int chunk = (n_cta + THREADS - 1) / THREADS;
  const int b0 = threadIdx.x * chunk, b1 = min(n_cta, b0 + chunk);
  for (int k = 0; k < 3; k++) {
    unsigned long long s = 0;
    for (int b = b0; b < b1; b++) s += cta_counts[(long long)b * 3 + k];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {  // 1024 partial sums: a serial exclusive scan is a few microseconds
      unsigned long long run = class_base[k];
      for (int t = 0; t < THREADS; t++) {
        const unsigned long long v = part[t];
        part[t] = run;
        run += v;
      }
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (int b = b0; b < b1; b++) {
      cta_off[(long long)b * 3 + k] = (unsigned)run;
      run += cta_counts[(long long)b * 3 + k];
    }
    __syncthreads();
  }
}

// perm[position] = aircraft index; stable inside every class
__global__ void __launch_bounds__(THREADS)
place_kernel(const unsigned char* __restrict__ fi, long long N, const unsigned* __restrict__ cta_off, unsigned* __restrict__ perm) {
  __shared__ unsigned wbase[THREADS / 32][3];
  const long long n = (long long)blockIdx.x * THREADS + threadIdx.x;
  const int c = n < N ? class_of(fi[n]) : 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned rank = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const unsigned m = __ballot_sync(0xffffffffu, c == k);
    if (c == k) rank = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) wbase[warp][k] = __popc(m);
  }
  __syncthreads();
  if (threadIdx.x < 3) {  // exclusive scan over the 32 warps
    unsigned run = cta_off[(long long)blockIdx.x * 3 + threadIdx.x];
    for (int w = 0; w < THREADS / 32; w++) {
      const unsigned v = wbase[w][threadIdx.x];
      wbase[w][threadIdx.x] = run;
      run += v;
    }
  }
  __syncthreads();
  if (c < 3) perm[wbase[warp][c] + rank] = (unsigned)n;
}

// dst[p][t] = src[p][perm[t]] (gather) or dst[p][perm[t]] = src[p][t] (scatter), `planes` planes of 8- or 4-byte elements
template <typename T, bool GATHER>
__global__ void __launch_bounds__(256)
move_kernel(const T* __restrict__ src, long long ld_src, T* __restrict__ dst, long long ld_dst, int planes,
            const unsigned* __restrict__ perm, long long N) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) {
    const long long n = perm[t];
    for (int p = 0; p < planes; p++) {
      if (GATHER) dst[p * ld_dst + t] = src[p * ld_src + n];
      else dst[p * ld_dst + n] = src[p * ld_src + t];
    }
  }
}

cudaError_t launch_build(const LaunchCfg& cfg, const unsigned char* fi, long long N, unsigned* perm, long long* totals_dev,
                         unsigned* scratch /* 6 * n_cta unsigned */) {
  const int n_cta = (int)((N + THREADS - 1) / THREADS);
  unsigned* counts = scratch;
  unsigned* off = scratch + (size_t)3 * n_cta;
  count_kernel<<<n_cta, THREADS, 0, cfg.stream>>>(fi, N, counts);
  scan_kernel<<<1, THREADS, 0, cfg.stream>>>(counts, n_cta, off, totals_dev);
  place_kernel<<<n_cta, THREADS, 0, cfg.stream>>>(fi, N, off, perm);
  if (cfg.launch_counter) *cfg.launch_counter += 3;
  return cudaGetLastError();
}

static int move_grid(const LaunchCfg& cfg, long long N) {
  long long b = (N + 255) / 256;
  const long long cap = (long long)cfg.sm_count * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

cudaError_t launch_gather_f64(const LaunchCfg& cfg, const double* src, long long ld_src, double* dst, long long ld_dst, int planes,
                              const unsigned* perm, long long N) {
  if (N <= 0) return cudaSuccess;
  move_kernel<double, true><<<move_grid(cfg, N), 256, 0, cfg.stream>>>(src, ld_src, dst, ld_dst, planes, perm, N);
  if (cfg.launch_counter) ++*cfg.launch_counter;
  return cudaGetLastError();
}
cudaError_t launch_scatter_f64(const LaunchCfg& cfg, const double* src, long long ld_src, double* dst, long long ld_dst, int planes,
                               const unsigned* perm, long long N) {
  if (N <= 0) return cudaSuccess;
  move_kernel<double, false><<<move_grid(cfg, N), 256, 0, cfg.stream>>>(src, ld_src, dst, ld_dst, planes, perm, N);
  if (cfg.launch_counter) ++*cfg.launch_counter;
  return cudaGetLastError();
}
cudaError_t launch_scatter_i32(const LaunchCfg& cfg, const int* src, int* dst, const unsigned* perm, long long N) {
  if (N <= 0) return cudaSuccess;
  move_kernel<int, false><<<move_grid(cfg, N), 256, 0, cfg.stream>>>(src, N, dst, N, 1, perm, N);
  if (cfg.launch_counter) ++*cfg.launch_counter;
  return cudaGetLastError();
}

__global__ void fill_i32_kernel(int* __restrict__ dst, int value, long long N) {
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) dst[n] = value;
}
cudaError_t launch_fill_i32(const LaunchCfg& cfg, int* dst, int value, long long N) {
  if (N <= 0) return cudaSuccess;
  fill_i32_kernel<<<move_grid(cfg, N), 256, 0, cfg.stream>>>(dst, value, N);
  if (cfg.launch_counter) ++*cfg.launch_counter;
  return cudaGetLastError();
}

cudaError_t launch_gather_i32(const LaunchCfg& cfg, const int* src, int* dst, const unsigned* perm, long long N) {
  if (N <= 0) return cudaSuccess;
  move_kernel<int, true><<<move_grid(cfg, N), 256, 0, cfg.stream>>>(src, N, dst, N, 1, perm, N);
  if (cfg.launch_counter) ++*cfg.launch_counter;
  return cudaGetLastError();
}

// ---- survivors of a Monte-Carlo run (f16_api.cu: step_compacting) -------------------------------------------------------
// after a chunk of k_launch[n] steps that started at step `base`: flag[n] = 1 for an aircraft still flying (status 0), 0 for one
// that has stopped; k_total[n] (< 0 = still flying) records the step at which an aircraft stopped, once
__global__ void mark_kernel(const int* __restrict__ status, const int* __restrict__ k_launch, int* __restrict__ k_total, int base,
                            long long N, unsigned char* __restrict__ flag) {
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
    const bool flying = status[n] == 0;
    flag[n] = flying ? 1 : 0;
    if (!flying && k_total[n] < 0) k_total[n] = base + k_launch[n];
  }
}
cudaError_t launch_mark_survivors(const LaunchCfg& cfg, const int* status, const int* k_launch, int* k_total, int base, long long N,
                                  unsigned char* flag) {
  if (N <= 0) return cudaSuccess;
  mark_kernel<<<move_grid(cfg, N), 256, 0, cfg.stream>>>(status, k_launch, k_total, base, N, flag);
  if (cfg.launch_counter) ++*cfg.launch_counter;
  return cudaGetLastError();
}
// end of the run: an aircraft that never stopped has taken all K steps
__global__ void close_steps_kernel(int* __restrict__ k_total, int K, long long N) {
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x)
    if (k_total[n] < 0) k_total[n] = K;
}
cudaError_t launch_close_steps(const LaunchCfg& cfg, int* k_total, int K, long long N) {
  if (N <= 0) return cudaSuccess;
  close_steps_kernel<<<move_grid(cfg, N), 256, 0, cfg.stream>>>(k_total, K, N);
  if (cfg.launch_counter) ++*cfg.launch_counter;
  return cudaGetLastError();
}

int n_cta(long long N) { return (int)((N + THREADS - 1) / THREADS); }

}  // namespace partition
}  // namespace f16

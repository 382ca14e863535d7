// f16_stats.cu -- the summary a Monte-Carlo run ends with, reduced on the device: over the
// aircraft whose status word is 0 ("alive": the reference would not have exit()ed, env.py:117-124), per state
// min, max, mean and M2 = sum (x - mean)^2, plus the two counts.  One row of 74 doubles
//     [n, alive, min[18], max[18], mean[18], M2[18]]
// is what a rank contributes to the only collective of a run (an all-gather of these rows, shard.py::merge_summaries).
//
// HBM-bound: every plane of the SoA state and the status words are read once per pass, 148 B per aircraft, fully
// coalesced (thread n reads x[i][n]; the status words are read by both state halves); two passes (sums -> mean, then squared deviations about that mean: no cancellation)
// = 296 B per aircraft, the second one mostly out of L2 for batches below ~0.8 Mi aircraft.  Reductions are a fixed
// tree (per thread -> warp shuffles -> shared memory -> one partial per CTA -> one finishing CTA, lane-strided), so the
// result is bit-reproducible for a given device and batch size; no atomics.
#include <math.h>
#include <stdint.h>

#include "f16_kernels.cuh"

namespace f16 {
namespace stats {

constexpr int THREADS = 256;
constexpr int NS = 18;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// pass 1: partial[b] = {sum[18], min[18], max[18], alive count} of CTA column b   (55 doubles, stride 56)
// pass 2: partial[b] = {sum (x - mean)^2 [18]}                                    (18 doubles, stride 56)
// blockIdx.y picks nine of the eighteen states: 27 accumulators per thread instead of 54 keep three CTAs resident per SM
// (the loads in flight, not the arithmetic, set the pace).
constexpr int NH = NS / 2;
template <int PASS>
__global__ void __launch_bounds__(THREADS, 3)
partial_kernel(const double* __restrict__ x, long long ld, long long N, const int* __restrict__ status,
               const double* __restrict__ row /* pass 2: the row with the means filled in */, double* __restrict__ partial) {
  __shared__ double red[THREADS / 32][3 * NH + 1];
  const int s0 = blockIdx.y * NH;  // first state of this half
  x += (long long)s0 * ld;
  double s[NH], mn[NH], mx[NH], mean[NH];
  double cnt = 0.0;
#pragma unroll
  for (int i = 0; i < NH; i++) {
    s[i] = 0.0;
    mn[i] = INFINITY;
    mx[i] = -INFINITY;
    mean[i] = PASS == 2 ? row[38 + s0 + i] : 0.0;
  }
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
    if (status && status[n] != 0) continue;
    cnt += 1.0;
#pragma unroll
    for (int i = 0; i < NH; i++) {
      const double v = x[i * ld + n];
      if (PASS == 1) {
        s[i] += v;
        mn[i] = v < mn[i] ? v : mn[i];  // a NaN never wins a comparison: skipped, as fmin / fmax would
        mx[i] = v > mx[i] ? v : mx[i];
      } else {
        const double d = v - mean[i];
        s[i] = fma(d, d, s[i]);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  cnt = warp_sum(cnt);
  if (lane == 0) red[warp][3 * NH] = cnt;
#pragma unroll
  for (int i = 0; i < NH; i++) {
    const double a = warp_sum(s[i]);
    if (lane == 0) red[warp][i] = a;
    if (PASS == 1) {
      const double b = warp_min(mn[i]), c = warp_max(mx[i]);
      if (lane == 0) {
        red[warp][NH + i] = b;
        red[warp][2 * NH + i] = c;
      }
    }
  }
  __syncthreads();
  const int nfield = PASS == 1 ? 3 * NH + 1 : NH;
  if (threadIdx.x < nfield) {
    const int f = threadIdx.x, kind = f / NH;  // 0 sum, 1 min, 2 max, 3 count
    double a = red[0][f];
    for (int w = 1; w < THREADS / 32; w++) {
      const double b = red[w][f];
      a = (PASS == 1 && kind == 1) ? fmin(a, b) : (PASS == 1 && kind == 2) ? fmax(a, b) : a + b;
    }
    // field layout of a partial: sum[18] | min[18] | max[18] | count
    const int dst = kind == 3 ? 3 * NS : kind * NS + s0 + (f - kind * NH);
    if (kind != 3 || blockIdx.y == 0) partial[(long long)blockIdx.x * 56 + dst] = a;
  }
}

// one CTA of 32 warps: warp w folds fields w and w + 32 of the per-CTA partials (lanes stride over the partials, then the
// same shuffle tree) into the 74-double row
template <int PASS>
__global__ void __launch_bounds__(1024)
finish_kernel(const double* __restrict__ partial, int n_part, long long N, double* __restrict__ row) {
  __shared__ double tot[3 * NS + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nfield = PASS == 1 ? 3 * NS + 1 : NS;
  for (int f = warp; f < nfield; f += 32) {
    const bool is_min = PASS == 1 && f >= NS && f < 2 * NS, is_max = PASS == 1 && f >= 2 * NS && f < 3 * NS;
    double a = is_min ? INFINITY : is_max ? -INFINITY : 0.0;
    for (int b = lane; b < n_part; b += 32) {
      const double v = partial[(long long)b * 56 + f];
      a = is_min ? fmin(a, v) : is_max ? fmax(a, v) : a + v;
    }
    a = is_min ? warp_min(a) : is_max ? warp_max(a) : warp_sum(a);
    if (lane == 0) tot[f] = a;
  }
  __syncthreads();
  const int f = threadIdx.x;
  if (PASS == 1) {
    if (f == 0) {
      row[0] = (double)N;
      row[1] = tot[3 * NS];
    }
    if (f < NS) row[38 + f] = tot[3 * NS] > 0.0 ? tot[f] / tot[3 * NS] : 0.0;  // mean = sum / alive
    if (f >= NS && f < 3 * NS) row[2 + (f - NS)] = tot[f];                       // min -> row[2..19], max -> row[20..37]
  } else if (f < NS) {
    row[56 + f] = tot[f];
  }
}

// row: 74 doubles on the device; scratch: at least 56 * grid doubles on the device
cudaError_t launch_summary(const LaunchCfg& cfg, const double* x, long long ld, long long N, const int* status, double* row,
                           double* scratch, int grid) {
  partial_kernel<1><<<dim3(grid, 2), THREADS, 0, cfg.stream>>>(x, ld, N, status, nullptr, scratch);
  finish_kernel<1><<<1, 1024, 0, cfg.stream>>>(scratch, grid, N, row);
  partial_kernel<2><<<dim3(grid, 2), THREADS, 0, cfg.stream>>>(x, ld, N, status, row, scratch);
  finish_kernel<2><<<1, 1024, 0, cfg.stream>>>(scratch, grid, N, row);
  if (cfg.launch_counter) *cfg.launch_counter += 4;
  return cudaGetLastError();
}

int summary_grid(const LaunchCfg& cfg, long long N) {
  // grid.x x 2 (state halves) CTAs of 256 threads: six resident per SM at most, nine independent loads per thread and
  // iteration in flight; never more CTAs than there is work for
  long long want = (N + THREADS - 1) / THREADS;
  const long long cap = (long long)cfg.sm_count * 3;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace stats
}  // namespace f16

// f16_step_fast.cu -- the hifi fused Euler step of F16_MATH_FAST (f16_fast.cuh).
// Compiled with -fmad=false: every fused multiply-add in this kernel is an explicit fma() in the source, so the
// result does not depend on which mul/add pairs ptxas would have chosen to contract in a given instantiation
// (CTA size, table staging) -- all variants of the kernel, and the host emulation of the tests, agree bit for bit.
#include "f16_kernels_common.cuh"
#include "f16_fast.cuh"

namespace f16 {
namespace fast {

__constant__ fastmath::LqrDense c_lqr_fast;

// ------------------------------------------------------------------------------------------------------
// step_batch, hifi, F16_MATH_FAST: the same K fused Euler steps on the re-associated arithmetic of f16_fast.cuh
// (fast table image: 171 KB in shared memory, one CTA per SM).  The per-step checks are the cheap "all inside"
// form; the exact status word is rebuilt from the frozen state when an aircraft stops.
// ------------------------------------------------------------------------------------------------------
constexpr int FAST_SMEM_BYTES = F16_FI_BYTES + 16;

template <bool SMEM, bool LQR, int THREADS, int COLMASK = 0>
__global__ void __launch_bounds__(THREADS, 1)
step_hifi_fast_kernel(DevTables tabs, BatchSel sel, double* __restrict__ x_g, long long ld_x,
                      const double* __restrict__ u_g, long long ld_u, long long N, int K, double dt,
                      int* __restrict__ status, int* __restrict__ steps_done) {
  const double* img = tabs.hifi_fast;
  if (SMEM) {
    stage_tables_tma<F16_FI_BYTES>(f16_smem, img, reinterpret_cast<unsigned long long*>(f16_smem + F16_FI_BYTES));
    img = reinterpret_cast<const double*>(f16_smem);
  }
#if defined(F16_FAST_LDS64)
  img += tabs.zero;  // always 0; keeps the table gathers 8-byte loads (see fastmath::fd)
#endif
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
    const int own = owns<1>(sel, n);
    if (own == 0) continue;
    if (own < 0) {
      if (status) status[n] = (int)ST_FIDELITY;
      if (steps_done) steps_done[n] = 0;
      continue;
    }
    double x[18], u_in[4];
#pragma unroll
    for (int i = 0; i < 18; i++) x[i] = x_g[i * ld_x + n];
#pragma unroll
    for (int i = 0; i < 4; i++) u_in[i] = u_g[i * ld_u + n];
    const double xcg = sel.xcg ? sel.xcg[n] : sel.xcg_default;
    int k;
    const unsigned st = fastmath::step_aircraft<LQR, 1, COLMASK>(img, x, u_in, LQR ? &c_lqr_fast : nullptr, xcg, dt, K, k);
#pragma unroll
    for (int i = 0; i < 18; i++) x_g[i * ld_x + n] = x[i];
    if (status) status[n] = (int)st;
    if (steps_done) steps_done[n] = k;
  }
}

// ------------------------------------------------------------------------------------------------------
// step_batch, lofi, F16_MATH_FAST: the Stevens-Lewis model on the same arithmetic (fastmath::calc_xdot_lofi).  Its step
// image is 7 KB (lofi tables + the centre table of half_rho), copied into shared memory by the CTA itself.
// ------------------------------------------------------------------------------------------------------
template <bool LQR, int THREADS, int COLMASK = 0>
__global__ void __launch_bounds__(THREADS, 1)
step_lofi_fast_kernel(DevTables tabs, BatchSel sel, double* __restrict__ x_g, long long ld_x,
                      const double* __restrict__ u_g, long long ld_u, long long N, int K, double dt,
                      int* __restrict__ status, int* __restrict__ steps_done) {
  double* img = reinterpret_cast<double*>(f16_smem);
  for (int i = threadIdx.x; i < F16_LOFI_STEP_IMG_DOUBLES; i += THREADS)
    img[i] = i < F16_IMG_LOFI_DOUBLES ? tabs.lofi[i] : tabs.hifi_fast[F16_FI_POW + (i - F16_IMG_LOFI_DOUBLES)];
  __syncthreads();
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
    if (owns<0>(sel, n) == 0) continue;  // fidelity flags other than 0 / 1 are reported by the hifi kernel
    double x[18], u_in[4];
#pragma unroll
    for (int i = 0; i < 18; i++) x[i] = x_g[i * ld_x + n];
#pragma unroll
    for (int i = 0; i < 4; i++) u_in[i] = u_g[i * ld_u + n];
    const double xcg = sel.xcg ? sel.xcg[n] : sel.xcg_default;
    int k;
    const unsigned st = fastmath::step_aircraft<LQR, 0, COLMASK>(img, x, u_in, LQR ? &c_lqr_fast : nullptr, xcg, dt, K, k);
#pragma unroll
    for (int i = 0; i < 18; i++) x_g[i * ld_x + n] = x[i];
    if (status) status[n] = (int)st;
    if (steps_done) steps_done[n] = k;
  }
}

// ------------------------------------------------------------------------------------------------------
// trim_batch, F16_MATH_FAST: the same Nelder-Mead search (f16_model.cuh::nelder_mead_trim_with) with the objective on the
// step kernel's arithmetic -- a third of the instructions per evaluation, and the search is ~2000 dependent evaluations
// per thread.  obj_func clips thrust, the three surfaces and alpha into their ranges (env.py:240-250) and leaves the
// flap state unclipped, so the precondition here is what calc_xdot_hifi / _lofi actually need (altitude inside the
// density table, no NaN), not the step bounds; a trial point outside it takes the reference-order objective on the
// standard image in global memory.
// ------------------------------------------------------------------------------------------------------
template <int FI>
static __device__ __noinline__ double trim_cost_reference_order(const double* img_std, double h, double V, double lef_q, double u0,
                                                                double u1, double u2, double u3, double u4, double xcg,
                                                                unsigned* st) {
  TrimPoint t;
  t.h = h;
  t.V = V;
  t.lef_q = lef_q;
  const double ux[5] = {u0, u1, u2, u3, u4};
  unsigned s = 0;
  const double c = trim_cost<FI>(img_std, t, ux, xcg, s);
  *st = s;
  return c;
}

template <int FI>
struct TrimCostFast {
  const double* img;      // the step image of this fidelity, in shared memory
  const double* img_std;  // the standard image, in global memory (rare path)
  __device__ __forceinline__ double operator()(const TrimPoint& t, const double (&ux)[5], double xcg, unsigned& st) const {
    double x[18], u[4], xd[18];
    trim_cost_inputs(t, ux, x, u);
    // trim_state: phi = psi = beta = p = q = r = 0, theta = alpha; what can be out of range is the caller's h and V
    const bool ok = (x[2] >= 0.0) & (x[2] <= 100000.0) & !either_nan(x[6], x[7]) & !either_nan(x[12], x[13]) &
                    !either_nan(x[14], x[15]) & !either_nan(x[16], x[17]);
    if (!ok) {
      unsigned s = 0;
      const double c = trim_cost_reference_order<FI>(img_std, t.h, t.V, t.lef_q, ux[0], ux[1], ux[2], ux[3], ux[4], xcg, &s);
      st = s;
      return c;
    }
    double uc[4];
    fastmath::clip_commands(u, uc);
    const bool inside = FI ? fastmath::calc_xdot_hifi<false>(img, x, uc, xcg, xd) : fastmath::calc_xdot_lofi<false>(img, x, uc, xcg, xd);
    if (!inside) {
      const double a = x[7] * (180.0 / 3.141592653589793), b = x[8] * (180.0 / 3.141592653589793);
      st = FI ? hifi_envelope(a, b, x[13]) : lofi_envelope(a, b, x[13]);
      return __builtin_huge_val();
    }
    st = 0;
    return trim_cost_sum(xd);
  }
};

struct TrimGuessFast {
  double ux[5];
};

template <int FI>
__global__ void __launch_bounds__(256, 1)
trim_fast_kernel(DevTables tabs, BatchSel sel, const double* __restrict__ h_g, const double* __restrict__ v_g, long long N,
                 double tol, int maxiter, TrimGuessFast guess, double* __restrict__ x_g, long long ld_x,
                 double* __restrict__ info_g, long long ld_info, int* __restrict__ status) {
  const double* img;
  if (FI) {
    stage_tables_tma<F16_FI_BYTES>(f16_smem, tabs.hifi_fast, reinterpret_cast<unsigned long long*>(f16_smem + F16_FI_BYTES));
    img = reinterpret_cast<const double*>(f16_smem);
#if defined(F16_FAST_LDS64)
    img += tabs.zero;
#endif
  } else {
    double* li = reinterpret_cast<double*>(f16_smem);
    for (int i = threadIdx.x; i < F16_LOFI_STEP_IMG_DOUBLES; i += blockDim.x)
      li[i] = i < F16_IMG_LOFI_DOUBLES ? tabs.lofi[i] : tabs.hifi_fast[F16_FI_POW + (i - F16_IMG_LOFI_DOUBLES)];
    __syncthreads();
    img = li;
  }
  const TrimCostFast<FI> cost{img, FI ? tabs.hifi : tabs.lofi};
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
    const int own = owns<FI>(sel, n);
    if (own == 0) continue;
    double x[18];
    TrimResult r;
    if (own == 1) {
      double ux[5];
#pragma unroll
      for (int k = 0; k < 5; k++) ux[k] = guess.ux[k];
      const TrimPoint t = trim_point(h_g[n], v_g[n]);
      r = nelder_mead_trim_with(cost, t, sel.xcg ? sel.xcg[n] : sel.xcg_default, tol, maxiter, ux);
      trim_state(t, ux, x);  // env.py:275-288: the optimiser's (unclipped) point
    } else {
      r.cost = qnan(); r.iterations = 0; r.fcalls = 0; r.converged = 0; r.status = ST_FIDELITY;
#pragma unroll
      for (int i = 0; i < 18; i++) x[i] = qnan();
    }
#pragma unroll
    for (int i = 0; i < 18; i++) x_g[i * ld_x + n] = x[i];
    if (info_g) {
      info_g[n] = r.cost;
      info_g[ld_info + n] = (double)r.iterations;
      info_g[2 * ld_info + n] = (double)r.fcalls;
      info_g[3 * ld_info + n] = (double)r.converged;
    }
    if (status) status[n] = (int)r.status;
  }
}

cudaError_t launch_trim_fast(const LaunchCfg& cfg, const DevTables& tabs, const BatchSel& sel, int FI, const double* h,
                             const double* v, long long N, double tol, int maxiter, const double* ux0, double* x_trim,
                             long long ld_x, double* info, long long ld_info, int* status) {
  if (N <= 0) return cudaSuccess;
  TrimGuessFast g;
  for (int k = 0; k < 5; k++) g.ux[k] = ux0[k];
  if (FI)
    return launch_persistent(cfg, trim_fast_kernel<1>, 256, FAST_SMEM_BYTES, N, 256, tabs, sel, h, v, N, tol, maxiter, g, x_trim,
                             ld_x, info, ld_info, status);
  return launch_persistent(cfg, trim_fast_kernel<0>, 256, F16_LOFI_STEP_IMG_DOUBLES * 8, N, 256, tabs, sel, h, v, N, tol, maxiter,
                           g, x_trim, ld_x, info, ld_info, status);
}

// ------------------------------------------------------------------------------------------------------
// f16_fast_probe: the fast image and its cell search as the step kernel uses them (fastmath::probe_hifi), one query per
// thread, image read from global memory.  coef [44][N], cells [4][N], lam [4][N]; a query outside the hifi tables gets
// NaN / -1 and the status word of hifi_envelope().
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fast_probe_kernel(DevTables tabs, const double* __restrict__ alpha, const double* __restrict__ beta,
                  const double* __restrict__ el, long long N, double* __restrict__ coef, int* __restrict__ cells,
                  double* __restrict__ lam, int* __restrict__ status) {
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
    const double a = alpha[n], b = beta[n], e = el[n];
    const unsigned st = hifi_envelope(a, b, e);
    double o[44], l[4];
    int c[4];
    if (st) {
      for (int i = 0; i < 44; i++) o[i] = qnan();
      for (int i = 0; i < 4; i++) { c[i] = -1; l[i] = qnan(); }
    } else {
      fastmath::probe_hifi(tabs.hifi_fast, a, b, e, o, c, l);
    }
    for (int i = 0; i < 44; i++) coef[i * N + n] = o[i];
    for (int i = 0; i < 4; i++) { cells[i * N + n] = c[i]; lam[i * N + n] = l[i]; }
    if (status) status[n] = (int)st;
  }
}

cudaError_t launch_fast_probe(const LaunchCfg& cfg, const DevTables& tabs, const double* alpha, const double* beta,
                              const double* el, long long N, double* coef, int* cells, double* lam, int* status) {
  if (N <= 0) return cudaSuccess;
  long long g = (N + 255) / 256;
  if (g > cfg.sm_count * 8) g = cfg.sm_count * 8;
  fast_probe_kernel<<<(unsigned)g, 256, 0, cfg.stream>>>(tabs, alpha, beta, el, N, coef, cells, lam, status);
  if (cfg.launch_counter) ++*cfg.launch_counter;
  return cudaGetLastError();
}

using StepKern = void (*)(DevTables, BatchSel, double*, long long, const double*, long long, long long, int, double, int*,
                          int*);

template <bool LQR>
static StepKern pick_step_hifi_fast(bool smem_tables, int& threads) {
  if (!smem_tables) { threads = 256; return step_hifi_fast_kernel<false, LQR, 256>; }
  if (threads <= 256) { threads = 256; return step_hifi_fast_kernel<true, LQR, 256>; }
  if (threads <= 384) { threads = 384; return step_hifi_fast_kernel<true, LQR, 384>; }
  if (threads <= 512) { threads = 512; return step_hifi_fast_kernel<true, LQR, 512>; }
  if (threads <= 640) { threads = 640; return step_hifi_fast_kernel<true, LQR, 640>; }
  threads = 768;
  return step_hifi_fast_kernel<true, LQR, 768>;
}

cudaError_t launch_step_hifi_fast(const LaunchCfg& cfg, const DevTables& tabs, const BatchSel& sel, double* x, long long ld_x,
                                  const double* u, long long ld_u, long long N, int K, double dt, const LqrLaw* lqr_host,
                                  int* status, int* steps_done) {
  if (N <= 0) return cudaSuccess;
  if (lqr_host) {
    static fastmath::LqrDense dense;  // caller holds the library mutex; the copy is enqueued before the launch below
    cudaError_t e = cudaStreamSynchronize(cfg.stream);  // a previous launch may still read the symbol's source
    if (e != cudaSuccess) return e;
    fastmath::make_dense_law(*lqr_host, dense);
    e = cudaMemcpyToSymbolAsync(c_lqr_fast, &dense, sizeof(dense), 0, cudaMemcpyHostToDevice, cfg.stream);
    if (e != cudaSuccess) return e;
  }
  int threads = cfg.step_threads;
  StepKern k = lqr_host ? pick_step_hifi_fast<true>(cfg.smem_tables, threads) : pick_step_hifi_fast<false>(cfg.smem_tables, threads);
  if (lqr_host && cfg.smem_tables && threads == 384) {  // the reference's own column set at the default CTA size: compile-time columns
    fastmath::LqrDense d;
    fastmath::make_dense_law(*lqr_host, d);
    if (d.colmask == F16_LQR_MPC_COLMASK) k = step_hifi_fast_kernel<true, true, 384, F16_LQR_MPC_COLMASK>;
  }
  const int smem = cfg.smem_tables ? FAST_SMEM_BYTES : 0;
  return launch_persistent(cfg, k, threads, smem, N, threads, tabs, sel, x, ld_x, u, ld_u, N, K, dt, status, steps_done);
}

// the caller (launch_step) has uploaded nothing yet for the law: same symbol, same protocol as above
cudaError_t launch_step_lofi_fast(const LaunchCfg& cfg, const DevTables& tabs, const BatchSel& sel, double* x, long long ld_x,
                                  const double* u, long long ld_u, long long N, int K, double dt, const LqrLaw* lqr_host,
                                  int* status, int* steps_done) {
  if (N <= 0) return cudaSuccess;
  if (lqr_host) {
    static fastmath::LqrDense dense;
    cudaError_t e = cudaStreamSynchronize(cfg.stream);
    if (e != cudaSuccess) return e;
    fastmath::make_dense_law(*lqr_host, dense);
    e = cudaMemcpyToSymbolAsync(c_lqr_fast, &dense, sizeof(dense), 0, cudaMemcpyHostToDevice, cfg.stream);
    if (e != cudaSuccess) return e;
  }
  const int smem = F16_LOFI_STEP_IMG_DOUBLES * 8;
  StepKern k = lqr_host ? step_lofi_fast_kernel<true, 384> : step_lofi_fast_kernel<false, 384>;
  if (lqr_host) {  // the reference's own column set: compile-time columns (as in the hifi kernel)
    fastmath::LqrDense d;
    fastmath::make_dense_law(*lqr_host, d);
    if (d.colmask == F16_LQR_MPC_COLMASK) k = step_lofi_fast_kernel<true, 384, F16_LQR_MPC_COLMASK>;
  }
  return launch_persistent(cfg, k, 384, smem, N, 384, tabs, sel, x, ld_x, u, ld_u, N, K, dt, status, steps_done);
}

}  // namespace fast
}  // namespace f16

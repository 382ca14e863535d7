// f16_api.cu -- the C ABI of libf16_b200.so (include/f16_b200.h): library state, table upload, host<->device
// plumbing and dispatch into the strict / fast kernel builds.  No arithmetic of the plant happens on the host.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/f16_b200.h"
#include "f16_kernels.cuh"
#include "f16_tables_host.h"

static_assert(sizeof(f16_lqr_t) == sizeof(f16::LqrLaw), "f16_lqr_t and the device-side law must have one layout");

namespace {

struct DevBuf {  // grow-only device scratch for the host-pointer entry points
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct State {
  bool ready = false;
  int init_rc = F16_ERR_NOINIT;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<double> payload;
  std::string table_source;
  double* d_hifi = nullptr;
  double* d_lofi = nullptr;
  double* d_hifi_fast = nullptr;
  int math_mode = F16_MATH_STRICT;
  int clr_mode = F16_CLR_AS_BUILT;
  bool smem_tables = true;
  int step_threads = 384;
  int lin_variant = 0;
  bool step_chunking = true;
  double default_xcg = 0.25;
  int last_status = 0;
  unsigned long long launches = 0;
  // legacy single-aircraft path: mapped pinned host memory, the kernel reads and writes it directly
  double* pin = nullptr;      // [17 in | 18 out | 3 atmos in/out ...]
  double* pin_dev = nullptr;
  DevBuf b_in, b_in2, b_out, b_fi, b_xcg, b_st, b_st2, b_a, b_b, b_flush, b_l1, b_l2, b_l3, b_l4, b_l5, b_sum, b_perm, b_pscr, b_px, b_pu, b_pxcg, b_pst, b_pk, b_redo, b_prog;
};

State G;
std::mutex G_mu;
thread_local std::string t_err;

void set_err(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_err = buf;
}

int cuda_fail(cudaError_t e, const char* what) {
  set_err("%s: %s", what, cudaGetErrorString(e));
  return F16_ERR_CUDA;
}
#define CK(call)                                          \
  do {                                                    \
    cudaError_t e__ = (call);                             \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

std::string lib_dir() {
  Dl_info info;
  if (dladdr((void*)&f16_init, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    size_t s = p.rfind('/');
    return s == std::string::npos ? std::string(".") : p.substr(0, s);
  }
  return std::string();
}

int upload_tables() {
  std::vector<double> img;
  f16::build_hifi_image(G.payload, G.clr_mode == F16_CLR_FROM_FILE, img);
  if (!G.d_hifi) CK(cudaMalloc(&G.d_hifi, F16_IMG_HIFI_BYTES));
  CK(cudaMemcpy(G.d_hifi, img.data(), F16_IMG_HIFI_BYTES, cudaMemcpyHostToDevice));
  f16::build_hifi_fast_image(G.payload, G.clr_mode == F16_CLR_FROM_FILE, img);
  if (!G.d_hifi_fast) CK(cudaMalloc(&G.d_hifi_fast, F16_FI_BYTES));
  CK(cudaMemcpy(G.d_hifi_fast, img.data(), F16_FI_BYTES, cudaMemcpyHostToDevice));
  f16::build_lofi_image(img);
  if (!G.d_lofi) CK(cudaMalloc(&G.d_lofi, F16_IMG_LOFI_BYTES));
  CK(cudaMemcpy(G.d_lofi, img.data(), F16_IMG_LOFI_BYTES, cudaMemcpyHostToDevice));
  return F16_OK;
}

int init_locked(const char* table_path, int device) {
  if (G.ready) return F16_OK;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_err("no CUDA device: %s (libf16_b200 has no CPU path)", e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    return G.init_rc = F16_ERR_CUDA;
  }
  if (device < 0) {
    const char* v = getenv("F16_DEVICE");
    if (!v || !*v) v = getenv("LOCAL_RANK");
    device = (v && *v) ? atoi(v) % ndev : 0;
  }
  if (device >= ndev) { set_err("device %d out of range (%d present)", device, ndev); return G.init_rc = F16_ERR_ARG; }
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_err("device %d is sm_%d%d; libf16_b200 is built for sm_100a only", device, prop.major, prop.minor);
    return G.init_rc = F16_ERR_CUDA;
  }
  G.device = device;
  G.sm_count = prop.multiProcessorCount;

  std::string err;
  if (!f16::load_canonical(table_path, lib_dir(), G.payload, G.table_source, err)) {
    set_err("%s", err.c_str());
    return G.init_rc = F16_ERR_TABLES;
  }
  if (!f16::check_grids(G.payload, err)) { set_err("%s", err.c_str()); return G.init_rc = F16_ERR_TABLES; }

  CK(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&G.ev0));
  CK(cudaEventCreate(&G.ev1));
  CK(cudaHostAlloc((void**)&G.pin, 64 * sizeof(double), cudaHostAllocMapped));
  CK(cudaHostGetDevicePointer((void**)&G.pin_dev, G.pin, 0));

  if (const char* v = getenv("F16_XCG")) G.default_xcg = atof(v);
  if (const char* v = getenv("F16_MATH")) G.math_mode = (!strcmp(v, "fast") || !strcmp(v, "1")) ? F16_MATH_FAST : F16_MATH_STRICT;
  if (const char* v = getenv("F16_CLR")) G.clr_mode = (!strcmp(v, "file") || !strcmp(v, "1")) ? F16_CLR_FROM_FILE : F16_CLR_AS_BUILT;
  if (const char* v = getenv("F16_STEP_THREADS")) G.step_threads = atoi(v);
  if (const char* v = getenv("F16_TABLE_STAGING")) G.smem_tables = atoi(v) != 0;
  if (const char* v = getenv("F16_STEP_CHUNKING")) G.step_chunking = atoi(v) != 0;
  if (const char* v = getenv("F16_LIN_VARIANT")) G.lin_variant = (atoi(v) == 1 || atoi(v) == 2) ? atoi(v) : 0;

  int rc = upload_tables();
  if (rc != F16_OK) return G.init_rc = rc;
  G.ready = true;
  return G.init_rc = F16_OK;
}

int ensure() {
  if (G.ready) return cudaSetDevice(G.device) == cudaSuccess ? F16_OK : F16_ERR_CUDA;
  return init_locked(nullptr, -1);
}

f16::LaunchCfg cfg(bool smem_tables) {
  f16::LaunchCfg c;
  c.stream = G.stream;
  c.sm_count = G.sm_count;
  c.step_threads = G.step_threads;
  c.smem_tables = smem_tables;
  c.lin_variant = G.lin_variant;
  c.launch_counter = &G.launches;
  c.step_chunking = G.step_chunking;
  c.step_progress = (int*)G.b_prog.p;
  c.step_progress_cap = (long long)(G.b_prog.cap / 4);
  return c;
}
f16::DevTables tabs() { return f16::DevTables{G.d_hifi, G.d_lofi, G.d_hifi_fast, 0}; }
f16::BatchSel sel_of(const unsigned char* fi, int fi_default, const double* xcg, double xcg_default) {
  return f16::BatchSel{fi, fi_default, xcg, xcg_default};
}

// host -> device staging of the optional per-aircraft selectors
int stage_sel(const unsigned char* fi, const double* xcg, long long N, const unsigned char** d_fi, const double** d_xcg) {
  *d_fi = nullptr;
  *d_xcg = nullptr;
  if (fi) {
    CK(G.b_fi.reserve((size_t)N));
    CK(cudaMemcpyAsync(G.b_fi.p, fi, (size_t)N, cudaMemcpyHostToDevice, G.stream));
    *d_fi = (const unsigned char*)G.b_fi.p;
  }
  if (xcg) {
    CK(G.b_xcg.reserve((size_t)N * 8));
    CK(cudaMemcpyAsync(G.b_xcg.p, xcg, (size_t)N * 8, cudaMemcpyHostToDevice, G.stream));
    *d_xcg = (const double*)G.b_xcg.p;
  }
  return F16_OK;
}

#define DISPATCH(fn, ...) (G.math_mode == F16_MATH_FAST ? f16::fast::fn(__VA_ARGS__) : f16::strict::fn(__VA_ARGS__))

// the one-shot entry points: in F16_MATH_FAST a batch worth staging tables for runs the TMA-tiled kernels on the arithmetic of
// f16_fast.cuh (f16_step_fast.cu); a handful of aircraft (the legacy Nlplant symbol among them) read the tables through L2
bool oneshot_fast(long long N) { return G.math_mode == F16_MATH_FAST && G.smem_tables && N >= 4096; }
cudaError_t run_nlplant(const f16::BatchSel& sel, const double* xu, long long ld_in, double* xdot, long long ld_out, long long N,
                        int* status) {
  if (oneshot_fast(N)) {
    cudaError_t e = G.b_redo.reserve((size_t)((N + 31) / 32) * 4);
    if (e != cudaSuccess) return e;
    return f16::fast::launch_xdot_fast(cfg(true), tabs(), sel, xu, ld_in, nullptr, 0, xdot, ld_out, N, status, (unsigned*)G.b_redo.p);
  }
  return DISPATCH(launch_nlplant, cfg(G.smem_tables && N >= 4096), tabs(), sel, xu, ld_in, xdot, ld_out, N, status);
}
// linearise_batch.  F16_MATH_STRICT: the staged reference-order kernels (variant 0 / 2: CTA per 32 aircraft, 1: warp per
// aircraft).  F16_MATH_FAST, variant 0: the two-aircraft-per-warp kernel on the arithmetic of f16_fast.cuh
// (f16_linearise_fast.cu); variants 1 and 2 keep the strict kernels in fast mode as well.
cudaError_t run_linearise(const f16::BatchSel& sel, const double* x, long long ld_x, const double* u, long long ld_u, long long N,
                          double eps, int scheme, double* A, double* B, int* status) {
  if (G.math_mode == F16_MATH_FAST && G.lin_variant == 0 && N < (1LL << 31)) {
    cudaError_t e = G.b_redo.reserve((size_t)((N + 1) / 2) * 4);
    if (e != cudaSuccess) return e;
    return f16::fast::launch_linearise_fast(cfg(true), tabs(), sel, x, ld_x, u, ld_u, N, eps, scheme, A, B, status,
                                            (unsigned*)G.b_redo.p);
  }
  return f16::strict::launch_linearise(cfg(true), tabs(), sel, x, ld_x, u, ld_u, N, eps, scheme, A, B, status);
}
cudaError_t run_calc_xdot(const f16::BatchSel& sel, const double* x, long long ld_x, const double* u, long long ld_u, double* xdot,
                          long long ld_out, long long N, int* status) {
  if (oneshot_fast(N)) {
    cudaError_t e = G.b_redo.reserve((size_t)((N + 31) / 32) * 4);
    if (e != cudaSuccess) return e;
    return f16::fast::launch_xdot_fast(cfg(true), tabs(), sel, x, ld_x, u, ld_u, xdot, ld_out, N, status, (unsigned*)G.b_redo.p);
  }
  return DISPATCH(launch_calc_xdot, cfg(G.smem_tables && N >= 4096), tabs(), sel, x, ld_x, u, ld_u, xdot, ld_out, N, status);
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int f16_init(const char* table_path, int device) {
  std::lock_guard<std::mutex> lk(G_mu);
  return init_locked(table_path, device);
}

void f16_shutdown(void) {
  std::lock_guard<std::mutex> lk(G_mu);
  if (!G.ready) return;
  cudaSetDevice(G.device);
  cudaStreamSynchronize(G.stream);
  for (DevBuf* b : {&G.b_in, &G.b_in2, &G.b_out, &G.b_fi, &G.b_xcg, &G.b_st, &G.b_st2, &G.b_a, &G.b_b, &G.b_flush, &G.b_l1, &G.b_l2,
                    &G.b_l3, &G.b_l4, &G.b_l5, &G.b_sum, &G.b_perm, &G.b_pscr, &G.b_px, &G.b_pu, &G.b_pxcg, &G.b_pst, &G.b_pk, &G.b_redo, &G.b_prog})
    b->release();
  if (G.d_hifi) cudaFree(G.d_hifi);
  if (G.d_lofi) cudaFree(G.d_lofi);
  if (G.d_hifi_fast) cudaFree(G.d_hifi_fast);
  if (G.pin) cudaFreeHost(G.pin);
  cudaEventDestroy(G.ev0);
  cudaEventDestroy(G.ev1);
  cudaStreamDestroy(G.stream);
  G = State();
}

const char* f16_last_error(void) { return t_err.c_str(); }
int f16_last_status(void) { return G.last_status; }
int f16_device(void) { return G.ready ? G.device : -1; }
int f16_sm_count(void) { return G.sm_count; }
unsigned long long f16_launch_count(void) { return G.launches; }
void* f16_stream(void) { return (void*)G.stream; }

int f16_set_math_mode(int mode) {
  std::lock_guard<std::mutex> lk(G_mu);
  int prev = G.math_mode;
  G.math_mode = mode == F16_MATH_FAST ? F16_MATH_FAST : F16_MATH_STRICT;
  return prev;
}

int f16_set_clr_mode(int mode) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  int prev = G.clr_mode;
  G.clr_mode = mode == F16_CLR_FROM_FILE ? F16_CLR_FROM_FILE : F16_CLR_AS_BUILT;
  if (G.clr_mode != prev) {
    cudaStreamSynchronize(G.stream);
    rc = upload_tables();
    if (rc != F16_OK) return rc;
  }
  return prev;
}

void f16_set_default_xcg(double xcg) { G.default_xcg = xcg; }

int f16_set_table_staging(int mode) {
  std::lock_guard<std::mutex> lk(G_mu);
  int prev = G.smem_tables ? 1 : 0;
  G.smem_tables = mode != 0;
  return prev;
}

int f16_set_step_threads(int threads) {
  std::lock_guard<std::mutex> lk(G_mu);
  int prev = G.step_threads;
  G.step_threads = threads;
  return prev;
}

int f16_set_step_chunking(int on) {
  std::lock_guard<std::mutex> lk(G_mu);
  int prev = G.step_chunking ? 1 : 0;
  G.step_chunking = on != 0;
  return prev;
}

int f16_set_linearise_variant(int variant) {
  std::lock_guard<std::mutex> lk(G_mu);
  int prev = G.lin_variant;
  G.lin_variant = (variant == 1 || variant == 2) ? variant : 0;
  return prev;
}

int f16_tables_sha256(char* out65) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  f16::payload_sha256_hex(G.payload, out65);
  return F16_OK;
}

// ---- legacy reference ABI -------------------------------------------------------------------------------
void f16_nlplant_xcg(const double* xu, double* xdot, int fidelity, double xcg) {
  std::lock_guard<std::mutex> lk(G_mu);
  const double nan = __builtin_nan("");
  if (ensure() != F16_OK) {
    for (int i = 0; i < 18; i++) xdot[i] = nan;
    G.last_status = F16_ST_NAN;
    fprintf(stderr, "libf16_b200: Nlplant: %s\n", t_err.c_str());
    return;
  }
  memcpy(G.pin, xu, 17 * sizeof(double));
  int* st_host = reinterpret_cast<int*>(G.pin + 40);
  int* st_dev = reinterpret_cast<int*>(G.pin_dev + 40);
  // one aircraft: tables read through L2 (no 105 KB staging), inputs and outputs in mapped pinned memory
  cudaError_t e = DISPATCH(launch_nlplant, cfg(false), tabs(), sel_of(nullptr, fidelity, nullptr, xcg), G.pin_dev, 1,
                           G.pin_dev + 17, 1, 1, st_dev);
  if (e == cudaSuccess) e = cudaStreamSynchronize(G.stream);
  if (e != cudaSuccess) {
    cuda_fail(e, "Nlplant");
    fprintf(stderr, "libf16_b200: Nlplant: %s\n", t_err.c_str());
    for (int i = 0; i < 18; i++) xdot[i] = nan;
    G.last_status = F16_ST_NAN;
    return;
  }
  memcpy(xdot, G.pin + 17, 18 * sizeof(double));
  G.last_status = *st_host;
}

void Nlplant(double* xu, double* xdot, int fidelity) { f16_nlplant_xcg(xu, xdot, fidelity, G.default_xcg); }

void atmos(double alt, double vt, double* coeff) { f16_atmos(alt, vt, coeff); }

void f16_atmos(double alt, double vt, double* coeff) {
  std::lock_guard<std::mutex> lk(G_mu);
  const double nan = __builtin_nan("");
  coeff[0] = coeff[1] = coeff[2] = nan;
  if (ensure() != F16_OK) {
    fprintf(stderr, "libf16_b200: atmos: %s\n", t_err.c_str());
    return;
  }
  G.pin[48] = alt;
  G.pin[49] = vt;
  cudaError_t e = DISPATCH(launch_atmos, cfg(false), G.pin_dev + 48, G.pin_dev + 49, 1, G.pin_dev + 50);
  if (e == cudaSuccess) e = cudaStreamSynchronize(G.stream);
  if (e != cudaSuccess) {
    cuda_fail(e, "atmos");
    fprintf(stderr, "libf16_b200: atmos: %s\n", t_err.c_str());
    return;
  }
  coeff[0] = G.pin[50];
  coeff[1] = G.pin[51];
  coeff[2] = G.pin[52];
}

// ---- device-pointer entry points ---------------------------------------------------------------------------
int Nlplant_batch_dev(const double* xu_soa, long long ld_in, double* xdot_soa, long long ld_out, const unsigned char* fi,
                      int fi_default, const double* xcg, double xcg_default, long long N, int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || (N > 0 && (!xu_soa || !xdot_soa)) || ld_in < N || ld_out < N) { set_err("Nlplant_batch_dev: bad argument"); return F16_ERR_ARG; }
  CK(run_nlplant(sel_of(fi, fi_default, xcg, xcg_default), xu_soa, ld_in, xdot_soa, ld_out, N, status));
  return F16_OK;
}

int calc_xdot_batch_dev(const double* x_soa, long long ld_x, const double* u_soa, long long ld_u, double* xdot_soa,
                        long long ld_out, const unsigned char* fi, int fi_default, const double* xcg, double xcg_default,
                        long long N, int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || (N > 0 && (!x_soa || !u_soa || !xdot_soa)) || ld_x < N || ld_u < N || ld_out < N) { set_err("calc_xdot_batch_dev: bad argument"); return F16_ERR_ARG; }
  CK(run_calc_xdot(sel_of(fi, fi_default, xcg, xcg_default), x_soa, ld_x, u_soa, ld_u, xdot_soa, ld_out, N, status));
  return F16_OK;
}

// scratch of the time-chunked step schedule (one int per 32 aircraft); without it the launch falls back to the plain kernel
static void reserve_step_progress(long long N) {
  if (G.step_chunking && N > 0 && G.b_prog.reserve((size_t)((N + 31) / 32) * 4) != cudaSuccess) cudaGetLastError();
}

// A mixed batch (per-aircraft fidelity flags) for the fused step: ordered by fidelity first, so that each of the two launches
// runs on a contiguous range with every lane busy (f16_partition.cu).  Applies when the reorder is amortised
// (N >= 4096 aircraft, K >= 8 steps); *handled = false leaves the call to the per-lane masking of the kernels.
static int step_partitioned(double* d_x, long long ld_x, const double* d_u, long long ld_u, long long N, int K, double dt,
                            const f16_lqr_t* lqr, const unsigned char* d_fi, const double* d_xcg, double xcg_default,
                            int* d_status, int* d_steps, bool* handled) {
  *handled = false;
  if (!d_fi || N < 4096 || K < 8 || N >= (1LL << 31)) return F16_OK;
  const size_t n = (size_t)N;
  const int n_cta = f16::partition::n_cta(N);
  // scratch is optional: when the device cannot hold a second copy of the batch the masked path does the job
  if (G.b_perm.reserve(n * 4) != cudaSuccess || G.b_pscr.reserve((size_t)n_cta * 6 * 4 + 64) != cudaSuccess ||
      G.b_px.reserve(18 * n * 8) != cudaSuccess || G.b_pu.reserve(4 * n * 8) != cudaSuccess ||
      G.b_pst.reserve(n * 4) != cudaSuccess || G.b_pk.reserve(n * 4) != cudaSuccess ||
      (d_xcg && G.b_pxcg.reserve(n * 8) != cudaSuccess)) {
    cudaGetLastError();
    return F16_OK;
  }
  const f16::LaunchCfg c = cfg(G.smem_tables);
  unsigned* perm = (unsigned*)G.b_perm.p;
  unsigned* scr = (unsigned*)G.b_pscr.p;
  long long* totals_dev = (long long*)(scr + (size_t)n_cta * 6 + 2);  // 8-byte aligned: n_cta * 24 + 8 bytes in
  CK(f16::partition::launch_build(c, d_fi, N, perm, totals_dev, scr));
  long long tot[3];
  CK(cudaMemcpyAsync(tot, totals_dev, sizeof(tot), cudaMemcpyDeviceToHost, G.stream));
  CK(cudaStreamSynchronize(G.stream));
  const long long n1 = tot[0], n0 = tot[1], nbad = tot[2];
  double* px = (double*)G.b_px.p;
  double* pu = (double*)G.b_pu.p;
  double* pxcg = d_xcg ? (double*)G.b_pxcg.p : nullptr;
  int* pst = (int*)G.b_pst.p;
  int* pk = (int*)G.b_pk.p;
  CK(f16::partition::launch_gather_f64(c, d_x, ld_x, px, N, 18, perm, N));
  CK(f16::partition::launch_gather_f64(c, d_u, ld_u, pu, N, 4, perm, N));
  if (d_xcg) CK(f16::partition::launch_gather_f64(c, d_xcg, N, pxcg, N, 1, perm, N));
  const f16::LqrLaw* law = reinterpret_cast<const f16::LqrLaw*>(lqr);
  if (n1 > 0)
    CK(DISPATCH(launch_step, c, tabs(), sel_of(nullptr, 1, pxcg, xcg_default), px, N, pu, N, n1, K, dt, law, pst, pk));
  if (n0 > 0)
    CK(DISPATCH(launch_step, c, tabs(), sel_of(nullptr, 0, pxcg ? pxcg + n1 : nullptr, xcg_default), px + n1, N, pu + n1, N, n0, K,
                dt, law, pst + n1, pk + n1));
  if (nbad > 0) {  // neither model: the state stays as it is, status = F16_ST_FIDELITY, no step taken
    std::vector<int> bad((size_t)nbad, (int)F16_ST_FIDELITY);
    CK(cudaMemcpyAsync(pst + n1 + n0, bad.data(), (size_t)nbad * 4, cudaMemcpyHostToDevice, G.stream));
    CK(cudaMemsetAsync(pk + n1 + n0, 0, (size_t)nbad * 4, G.stream));
    CK(cudaStreamSynchronize(G.stream));  // `bad` is pageable and goes out of scope
  }
  CK(f16::partition::launch_scatter_f64(c, px, N, d_x, ld_x, 18, perm, N));
  if (d_status) CK(f16::partition::launch_scatter_i32(c, pst, d_status, perm, N));
  if (d_steps) CK(f16::partition::launch_scatter_i32(c, pk, d_steps, perm, N));
  *handled = true;
  return F16_OK;
}

int step_batch_dev(double* x_soa, long long ld_x, const double* u_soa, long long ld_u, long long N, int K, double dt,
                   const f16_lqr_t* lqr, const unsigned char* fi, int fi_default, const double* xcg, double xcg_default,
                   int* status, int* steps_done) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || K < 0 || (N > 0 && (!x_soa || !u_soa)) || ld_x < N || ld_u < N) { set_err("step_batch_dev: bad argument"); return F16_ERR_ARG; }
  if (lqr && (lqr->n_sel < 0 || lqr->n_sel > 18)) { set_err("step_batch_dev: lqr.n_sel out of range"); return F16_ERR_ARG; }
  if (lqr) for (int j = 0; j < lqr->n_sel; j++) if (lqr->sel[j] < 0 || lqr->sel[j] > 17) { set_err("step_batch_dev: lqr.sel out of range"); return F16_ERR_ARG; }
  // tables go to shared memory whenever the launch does real work; a handful of aircraft-steps read them via L2
  reserve_step_progress(N);
  bool handled = false;
  if ((rc = step_partitioned(x_soa, ld_x, u_soa, ld_u, N, K, dt, lqr, fi, xcg, xcg_default, status, steps_done, &handled)) != F16_OK)
    return rc;
  if (handled) return F16_OK;
  const bool smem = G.smem_tables && (N * (long long)(K > 0 ? K : 1) >= 4096);
  CK(DISPATCH(launch_step, cfg(smem), tabs(), sel_of(fi, fi_default, xcg, xcg_default), x_soa, ld_x, u_soa, ld_u, N, K, dt,
              reinterpret_cast<const f16::LqrLaw*>(lqr), status, steps_done));
  return F16_OK;
}

int linearise_batch_dev(const double* x_soa, long long ld_x, const double* u_soa, long long ld_u, long long N, double eps,
                        int scheme, double* A, double* B, const unsigned char* fi, int fi_default, const double* xcg,
                        double xcg_default, int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || (N > 0 && (!x_soa || !u_soa || !A || !B)) || ld_x < N || ld_u < N || (scheme != 0 && scheme != 1) || !(eps >= 1e-12 && eps <= 1e3)) {
    set_err("linearise_batch_dev: bad argument");
    return F16_ERR_ARG;
  }
  CK(run_linearise(sel_of(fi, fi_default, xcg, xcg_default), x_soa, ld_x, u_soa, ld_u, N, eps, scheme, A, B, status));
  return F16_OK;
}

// ---- host-pointer entry points: H2D, kernels, D2H on the library stream ---------------------------------------
#define H2D(dst, src, bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, G.stream))
#define D2H(dst, src, bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, G.stream))

int Nlplant_batch(const double* xu_soa, double* xdot_soa, const unsigned char* fi, int fi_default, const double* xcg,
                  double xcg_default, long long N, int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || (N > 0 && (!xu_soa || !xdot_soa))) { set_err("Nlplant_batch: bad argument"); return F16_ERR_ARG; }
  if (N == 0) return F16_OK;
  const size_t n = (size_t)N;
  CK(G.b_in.reserve(17 * n * 8));
  CK(G.b_out.reserve(18 * n * 8));
  CK(G.b_st.reserve(n * 4));
  const unsigned char* d_fi;
  const double* d_xcg;
  if ((rc = stage_sel(fi, xcg, N, &d_fi, &d_xcg)) != F16_OK) return rc;
  H2D(G.b_in.p, xu_soa, 17 * n * 8);
  CK(run_nlplant(sel_of(d_fi, fi_default, d_xcg, xcg_default), (const double*)G.b_in.p, N, (double*)G.b_out.p, N, N,
                 (int*)G.b_st.p));
  D2H(xdot_soa, G.b_out.p, 18 * n * 8);
  if (status) D2H(status, G.b_st.p, n * 4);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

int calc_xdot_batch(const double* x_soa, const double* u_soa, double* xdot_soa, const unsigned char* fi, int fi_default,
                    const double* xcg, double xcg_default, long long N, int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || (N > 0 && (!x_soa || !u_soa || !xdot_soa))) { set_err("calc_xdot_batch: bad argument"); return F16_ERR_ARG; }
  if (N == 0) return F16_OK;
  const size_t n = (size_t)N;
  CK(G.b_in.reserve(18 * n * 8));
  CK(G.b_in2.reserve(4 * n * 8));
  CK(G.b_out.reserve(18 * n * 8));
  CK(G.b_st.reserve(n * 4));
  const unsigned char* d_fi;
  const double* d_xcg;
  if ((rc = stage_sel(fi, xcg, N, &d_fi, &d_xcg)) != F16_OK) return rc;
  H2D(G.b_in.p, x_soa, 18 * n * 8);
  H2D(G.b_in2.p, u_soa, 4 * n * 8);
  CK(run_calc_xdot(sel_of(d_fi, fi_default, d_xcg, xcg_default), (const double*)G.b_in.p, N, (const double*)G.b_in2.p, N,
                   (double*)G.b_out.p, N, N, (int*)G.b_st.p));
  D2H(xdot_soa, G.b_out.p, 18 * n * 8);
  if (status) D2H(status, G.b_st.p, n * 4);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

int step_batch(double* x_soa, const double* u_soa, long long N, int K, double dt, const f16_lqr_t* lqr,
               const unsigned char* fi, int fi_default, const double* xcg, double xcg_default, int* status,
               int* steps_done) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || K < 0 || (N > 0 && (!x_soa || !u_soa))) { set_err("step_batch: bad argument"); return F16_ERR_ARG; }
  if (lqr && (lqr->n_sel < 0 || lqr->n_sel > 18)) { set_err("step_batch: lqr.n_sel out of range"); return F16_ERR_ARG; }
  if (lqr) for (int j = 0; j < lqr->n_sel; j++) if (lqr->sel[j] < 0 || lqr->sel[j] > 17) { set_err("step_batch: lqr.sel out of range"); return F16_ERR_ARG; }
  if (N == 0) return F16_OK;
  const size_t n = (size_t)N;
  CK(G.b_in.reserve(18 * n * 8));
  CK(G.b_in2.reserve(4 * n * 8));
  CK(G.b_st.reserve(n * 4));
  CK(G.b_st2.reserve(n * 4));
  const unsigned char* d_fi;
  const double* d_xcg;
  if ((rc = stage_sel(fi, xcg, N, &d_fi, &d_xcg)) != F16_OK) return rc;
  H2D(G.b_in.p, x_soa, 18 * n * 8);
  H2D(G.b_in2.p, u_soa, 4 * n * 8);
  reserve_step_progress(N);
  bool handled = false;
  if ((rc = step_partitioned((double*)G.b_in.p, N, (const double*)G.b_in2.p, N, N, K, dt, lqr, d_fi, d_xcg, xcg_default,
                             (int*)G.b_st.p, (int*)G.b_st2.p, &handled)) != F16_OK)
    return rc;
  if (!handled) {
    const bool smem = G.smem_tables && (N * (long long)(K > 0 ? K : 1) >= 4096);
    CK(DISPATCH(launch_step, cfg(smem), tabs(), sel_of(d_fi, fi_default, d_xcg, xcg_default), (double*)G.b_in.p, N,
                (const double*)G.b_in2.p, N, N, K, dt, reinterpret_cast<const f16::LqrLaw*>(lqr), (int*)G.b_st.p,
                (int*)G.b_st2.p));
  }
  D2H(x_soa, G.b_in.p, 18 * n * 8);
  if (status) D2H(status, G.b_st.p, n * 4);
  if (steps_done) D2H(steps_done, G.b_st2.p, n * 4);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

// K Euler steps with a snapshot of the whole state every `snap_every` steps: traj [K / snap_every][18][N].
// Launches of snap_every steps back to back on one stream (step_batch is restartable bit for bit at any K boundary),
// each snapshot copied out while the next chunk runs.
int step_batch_traj(double* x_soa, const double* u_soa, long long N, int K, int snap_every, double dt, const f16_lqr_t* lqr,
                    const unsigned char* fi, int fi_default, const double* xcg, double xcg_default, double* traj, int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || K < 0 || snap_every < 1 || (N > 0 && (!x_soa || !u_soa || (!traj && K >= snap_every)))) {
    set_err("step_batch_traj: bad argument");
    return F16_ERR_ARG;
  }
  if (lqr && (lqr->n_sel < 0 || lqr->n_sel > 18)) { set_err("step_batch_traj: lqr.n_sel out of range"); return F16_ERR_ARG; }
  if (lqr) for (int j = 0; j < lqr->n_sel; j++) if (lqr->sel[j] < 0 || lqr->sel[j] > 17) { set_err("step_batch_traj: lqr.sel out of range"); return F16_ERR_ARG; }
  if (N == 0) return F16_OK;
  const size_t n = (size_t)N;
  CK(G.b_in.reserve(18 * n * 8));
  CK(G.b_in2.reserve(4 * n * 8));
  CK(G.b_out.reserve(2 * 18 * n * 8));  // two snapshot slots: copy-out of one overlaps the next chunk
  CK(G.b_st.reserve(n * 4));
  const unsigned char* d_fi;
  const double* d_xcg;
  if ((rc = stage_sel(fi, xcg, N, &d_fi, &d_xcg)) != F16_OK) return rc;
  H2D(G.b_in.p, x_soa, 18 * n * 8);
  H2D(G.b_in2.p, u_soa, 4 * n * 8);
  const bool smem = G.smem_tables && (N * (long long)snap_every >= 4096);
  const f16::BatchSel sel = sel_of(d_fi, fi_default, d_xcg, xcg_default);
  int done = 0, snap = 0;
  while (done < K) {
    const int k = (K - done) < snap_every ? (K - done) : snap_every;
    CK(DISPATCH(launch_step, cfg(smem), tabs(), sel, (double*)G.b_in.p, N, (const double*)G.b_in2.p, N, N, k, dt,
                reinterpret_cast<const f16::LqrLaw*>(lqr), (int*)G.b_st.p, nullptr));
    done += k;
    if (k == snap_every) {
      double* slot = (double*)G.b_out.p + (size_t)(snap & 1) * 18 * n;
      // same stream: the slot's previous copy-out has completed before this copy starts
      CK(cudaMemcpyAsync(slot, G.b_in.p, 18 * n * 8, cudaMemcpyDeviceToDevice, G.stream));
      CK(cudaMemcpyAsync(traj + (size_t)snap * 18 * n, slot, 18 * n * 8, cudaMemcpyDeviceToHost, G.stream));
      snap++;
    }
  }
  D2H(x_soa, G.b_in.p, 18 * n * 8);
  if (status) D2H(status, G.b_st.p, n * 4);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

int linearise_batch(const double* x_soa, const double* u_soa, long long N, double eps, int scheme, double* A, double* B,
                    const unsigned char* fi, int fi_default, const double* xcg, double xcg_default, int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || (N > 0 && (!x_soa || !u_soa || !A || !B)) || (scheme != 0 && scheme != 1) || !(eps >= 1e-12 && eps <= 1e3)) {
    set_err("linearise_batch: bad argument");
    return F16_ERR_ARG;
  }
  if (N == 0) return F16_OK;
  const size_t n = (size_t)N;
  CK(G.b_in.reserve(18 * n * 8));
  CK(G.b_in2.reserve(4 * n * 8));
  CK(G.b_a.reserve(324 * n * 8));
  CK(G.b_b.reserve(72 * n * 8));
  CK(G.b_st.reserve(n * 4));
  const unsigned char* d_fi;
  const double* d_xcg;
  if ((rc = stage_sel(fi, xcg, N, &d_fi, &d_xcg)) != F16_OK) return rc;
  H2D(G.b_in.p, x_soa, 18 * n * 8);
  H2D(G.b_in2.p, u_soa, 4 * n * 8);
  CK(run_linearise(sel_of(d_fi, fi_default, d_xcg, xcg_default), (const double*)G.b_in.p, N, (const double*)G.b_in2.p, N, N, eps,
                   scheme, (double*)G.b_a.p, (double*)G.b_b.p, (int*)G.b_st.p));
  D2H(A, G.b_a.p, 324 * n * 8);
  D2H(B, G.b_b.p, 72 * n * 8);
  if (status) D2H(status, G.b_st.p, n * 4);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

// ---- trim_batch: env.py::trim for N flight conditions ------------------------------------------------------------
static const double kTrimGuess[5] = {5000, -0.09, 8.49, -0.01, 0.01};  // env.py:264-271 (unpacked as P3, dh, da, dr, alpha)

int trim_batch_dev(const double* h, const double* V, long long N, double tol, int maxiter, const double* ux0, double* x_trim_soa,
                   long long ld_x, double* info_soa, long long ld_info, const unsigned char* fi, int fi_default, const double* xcg,
                   double xcg_default, int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || (N > 0 && (!h || !V || !x_trim_soa)) || ld_x < N || (info_soa && ld_info < N) || !(tol >= 0) || maxiter < 1) {
    set_err("trim_batch_dev: bad argument");
    return F16_ERR_ARG;
  }
  CK(DISPATCH(launch_trim, cfg(G.smem_tables && N >= 1024), tabs(), sel_of(fi, fi_default, xcg, xcg_default), h, V, N, tol, maxiter,
              ux0 ? ux0 : kTrimGuess, x_trim_soa, ld_x, info_soa, ld_info, status));
  return F16_OK;
}

int trim_batch(const double* h, const double* V, long long N, double tol, int maxiter, const double* ux0, double* x_trim_soa,
               double* info_soa, const unsigned char* fi, int fi_default, const double* xcg, double xcg_default, int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || (N > 0 && (!h || !V || !x_trim_soa)) || !(tol >= 0) || maxiter < 1) { set_err("trim_batch: bad argument"); return F16_ERR_ARG; }
  if (N == 0) return F16_OK;
  const size_t n = (size_t)N;
  CK(G.b_in.reserve(2 * n * 8));
  CK(G.b_out.reserve(18 * n * 8));
  CK(G.b_a.reserve(4 * n * 8));
  CK(G.b_st.reserve(n * 4));
  const unsigned char* d_fi;
  const double* d_xcg;
  if ((rc = stage_sel(fi, xcg, N, &d_fi, &d_xcg)) != F16_OK) return rc;
  double* d = (double*)G.b_in.p;
  H2D(d, h, n * 8);
  H2D(d + n, V, n * 8);
  CK(DISPATCH(launch_trim, cfg(G.smem_tables && N >= 1024), tabs(), sel_of(d_fi, fi_default, d_xcg, xcg_default), d, d + n, N, tol,
              maxiter, ux0 ? ux0 : kTrimGuess, (double*)G.b_out.p, N, (double*)G.b_a.p, N, (int*)G.b_st.p));
  D2H(x_trim_soa, G.b_out.p, 18 * n * 8);
  if (info_soa) D2H(info_soa, G.b_a.p, 4 * n * 8);
  if (status) D2H(status, G.b_st.p, n * 4);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

// ---- end-of-run statistics, reduced on the device (f16_stats.cu) ------------------------------------------------------
// enqueue the four reduction kernels; row_dev (74 doubles) must not alias the scratch (56 doubles per CTA column)
static int enqueue_summary(const double* d_x, long long ld, long long N, const int* d_status, double* row_dev, double* scratch,
                           int grid) {
  CK(f16::stats::launch_summary(cfg(false), d_x, ld, N, d_status, row_dev, scratch, grid));
  return F16_OK;
}

static int summary_common(const double* d_x, long long ld, long long N, const int* d_status, double* row_host) {
  const int grid = f16::stats::summary_grid(cfg(false), N);
  CK(G.b_sum.reserve(((size_t)grid * 56 + 80) * 8));
  double* scratch = (double*)G.b_sum.p;
  double* row = scratch + (size_t)grid * 56;
  int rc = enqueue_summary(d_x, ld, N, d_status, row, scratch, grid);
  if (rc != F16_OK) return rc;
  D2H(row_host, row, 74 * 8);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

// K steps in chunks of snap_every, one summary row per chunk; x / u / status on the device.  rows_host [K / snap_every][74].
static int step_stats_core(double* d_x, long long ld_x, const double* d_u, long long ld_u, long long N, int K, int snap_every,
                           double dt, const f16_lqr_t* lqr, const f16::BatchSel& sel, int* d_status, double* rows_host) {
  const int n_rows = K / snap_every;
  const int grid = f16::stats::summary_grid(cfg(false), N);
  CK(G.b_sum.reserve(((size_t)grid * 56 + (size_t)(n_rows > 0 ? n_rows : 1) * 74 + 8) * 8));
  double* scratch = (double*)G.b_sum.p;
  double* rows_dev = scratch + (size_t)grid * 56;
  const bool smem = G.smem_tables && (N * (long long)snap_every >= 4096);
  int done = 0, snap = 0;
  while (done < K) {
    const int k = (K - done) < snap_every ? (K - done) : snap_every;
    CK(DISPATCH(launch_step, cfg(smem), tabs(), sel, d_x, ld_x, d_u, ld_u, N, k, dt, reinterpret_cast<const f16::LqrLaw*>(lqr),
                d_status, nullptr));
    done += k;
    if (k == snap_every) {
      int rc = enqueue_summary(d_x, ld_x, N, d_status, rows_dev + (size_t)snap * 74, scratch, grid);
      if (rc != F16_OK) return rc;
      snap++;
    }
  }
  if (n_rows > 0) D2H(rows_host, rows_dev, (size_t)n_rows * 74 * 8);
  return F16_OK;
}

static bool lqr_ok(const f16_lqr_t* lqr) {
  if (!lqr) return true;
  if (lqr->n_sel < 0 || lqr->n_sel > 18) return false;
  for (int j = 0; j < lqr->n_sel; j++)
    if (lqr->sel[j] < 0 || lqr->sel[j] > 17) return false;
  return true;
}

int step_batch_stats_dev(double* x_soa, long long ld_x, const double* u_soa, long long ld_u, long long N, int K, int snap_every,
                         double dt, const f16_lqr_t* lqr, const unsigned char* fi, int fi_default, const double* xcg,
                         double xcg_default, double* rows, int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || K < 0 || snap_every < 1 || (N > 0 && (!x_soa || !u_soa || !status)) || ld_x < N || ld_u < N || !lqr_ok(lqr) ||
      (!rows && K >= snap_every)) {
    set_err("step_batch_stats_dev: bad argument");
    return F16_ERR_ARG;
  }
  rc = step_stats_core(x_soa, ld_x, u_soa, ld_u, N, K, snap_every, dt, lqr, sel_of(fi, fi_default, xcg, xcg_default), status, rows);
  if (rc != F16_OK) return rc;
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

int step_batch_stats(double* x_soa, const double* u_soa, long long N, int K, int snap_every, double dt, const f16_lqr_t* lqr,
                     const unsigned char* fi, int fi_default, const double* xcg, double xcg_default, double* rows, int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || K < 0 || snap_every < 1 || (N > 0 && (!x_soa || !u_soa)) || !lqr_ok(lqr) || (!rows && K >= snap_every)) {
    set_err("step_batch_stats: bad argument");
    return F16_ERR_ARG;
  }
  const size_t n = (size_t)N;
  CK(G.b_in.reserve((n ? 18 * n : 1) * 8));
  CK(G.b_in2.reserve((n ? 4 * n : 1) * 8));
  CK(G.b_st.reserve((n ? n : 1) * 4));
  const unsigned char* d_fi = nullptr;
  const double* d_xcg = nullptr;
  if (n) {
    if ((rc = stage_sel(fi, xcg, N, &d_fi, &d_xcg)) != F16_OK) return rc;
    H2D(G.b_in.p, x_soa, 18 * n * 8);
    H2D(G.b_in2.p, u_soa, 4 * n * 8);
    CK(cudaMemsetAsync(G.b_st.p, 0, n * 4, G.stream));
  }
  rc = step_stats_core((double*)G.b_in.p, N, (const double*)G.b_in2.p, N, N, K, snap_every, dt, lqr,
                       sel_of(d_fi, fi_default, d_xcg, xcg_default), (int*)G.b_st.p, rows);
  if (rc != F16_OK) return rc;
  if (n) {
    D2H(x_soa, G.b_in.p, 18 * n * 8);
    if (status) D2H(status, G.b_st.p, n * 4);
  }
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

int state_summary_batch_dev(const double* x_soa, long long ld_x, long long N, const int* status, double* row) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || !row || (N > 0 && !x_soa) || ld_x < N) { set_err("state_summary_batch_dev: bad argument"); return F16_ERR_ARG; }
  return summary_common(x_soa, ld_x, N, status, row);
}

int state_summary_batch(const double* x_soa, long long N, const int* status, double* row) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || !row || (N > 0 && !x_soa)) { set_err("state_summary_batch: bad argument"); return F16_ERR_ARG; }
  const size_t n = (size_t)N;
  CK(G.b_in.reserve((n ? 18 * n : 1) * 8));
  if (n) H2D(G.b_in.p, x_soa, 18 * n * 8);
  const int* d_st = nullptr;
  if (status && n) {
    CK(G.b_st.reserve(n * 4));
    H2D(G.b_st.p, status, n * 4);
    d_st = (const int*)G.b_st.p;
  }
  return summary_common((const double*)G.b_in.p, N, N, d_st, row);
}

// ---- between linearise and the LQR law: reduced model, zero-order hold, discrete LQR gain --------------------------------
static bool dims_ok(int n, int m) { return n >= 1 && m >= 1 && n <= 18 && n + m <= 22; }

int reduce_jacobian_batch_dev(const double* A, long long N, double* A_na, double* B_na) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || (N > 0 && (!A || !A_na || !B_na))) { set_err("reduce_jacobian_batch_dev: bad argument"); return F16_ERR_ARG; }
  CK(f16::linalg::launch_reduce_jacobian(cfg(false), A, N, A_na, B_na));
  return F16_OK;
}

int discretise_batch_dev(const double* A, const double* B, int n, int m, long long N, double dt, double* Ad, double* Bd) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || !dims_ok(n, m) || (N > 0 && (!A || !B || !Ad || !Bd))) { set_err("discretise_batch_dev: bad argument"); return F16_ERR_ARG; }
  CK(f16::linalg::launch_zoh(cfg(false), A, B, n, m, N, dt, Ad, Bd));
  return F16_OK;
}

int dlqr_batch_dev(const double* Ad, const double* Bd, const double* Q /* device, n x n */, const double* R /* device, m x m */,
                   int n, int m, long long N, double* K, double* P, int* info) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || !dims_ok(n, m) || (N > 0 && (!Ad || !Bd || !Q || !R || !K))) { set_err("dlqr_batch_dev: bad argument"); return F16_ERR_ARG; }
  CK(f16::linalg::launch_dlqr(cfg(false), Ad, Bd, Q, R, n, m, N, 64, 1e-15, K, P, info));
  return F16_OK;
}

int reduce_jacobian_batch(const double* A, long long N, double* A_na, double* B_na) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || (N > 0 && (!A || !A_na || !B_na))) { set_err("reduce_jacobian_batch: bad argument"); return F16_ERR_ARG; }
  if (N == 0) return F16_OK;
  const size_t n = (size_t)N;
  CK(G.b_l1.reserve(n * 324 * 8));
  CK(G.b_l2.reserve(n * 81 * 8));
  CK(G.b_l3.reserve(n * 27 * 8));
  H2D(G.b_l1.p, A, n * 324 * 8);
  CK(f16::linalg::launch_reduce_jacobian(cfg(false), (const double*)G.b_l1.p, N, (double*)G.b_l2.p, (double*)G.b_l3.p));
  D2H(A_na, G.b_l2.p, n * 81 * 8);
  D2H(B_na, G.b_l3.p, n * 27 * 8);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

int discretise_batch(const double* A, const double* B, int n, int m, long long N, double dt, double* Ad, double* Bd) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || !dims_ok(n, m) || (N > 0 && (!A || !B || !Ad || !Bd))) { set_err("discretise_batch: bad argument"); return F16_ERR_ARG; }
  if (N == 0) return F16_OK;
  const size_t na = (size_t)N * n * n * 8, nb = (size_t)N * n * m * 8;
  CK(G.b_l1.reserve(na));
  CK(G.b_l2.reserve(nb));
  CK(G.b_l3.reserve(na));
  CK(G.b_l4.reserve(nb));
  H2D(G.b_l1.p, A, na);
  H2D(G.b_l2.p, B, nb);
  CK(f16::linalg::launch_zoh(cfg(false), (const double*)G.b_l1.p, (const double*)G.b_l2.p, n, m, N, dt, (double*)G.b_l3.p,
                             (double*)G.b_l4.p));
  D2H(Ad, G.b_l3.p, na);
  D2H(Bd, G.b_l4.p, nb);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

int dlqr_batch(const double* Ad, const double* Bd, const double* Q, const double* R, int n, int m, long long N, double* K, double* P,
               int* info) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || !dims_ok(n, m) || (N > 0 && (!Ad || !Bd || !Q || !R || !K))) { set_err("dlqr_batch: bad argument"); return F16_ERR_ARG; }
  if (N == 0) return F16_OK;
  const size_t na = (size_t)N * n * n * 8, nb = (size_t)N * n * m * 8, nk = (size_t)N * m * n * 8;
  CK(G.b_l1.reserve(na));
  CK(G.b_l2.reserve(nb));
  CK(G.b_l3.reserve(nk));
  CK(G.b_l4.reserve(na));
  CK(G.b_l5.reserve((size_t)(n * n + m * m) * 8));
  CK(G.b_st.reserve((size_t)N * 8));
  H2D(G.b_l1.p, Ad, na);
  H2D(G.b_l2.p, Bd, nb);
  double* dQ = (double*)G.b_l5.p;
  H2D(dQ, Q, (size_t)n * n * 8);
  H2D(dQ + n * n, R, (size_t)m * m * 8);
  CK(f16::linalg::launch_dlqr(cfg(false), (const double*)G.b_l1.p, (const double*)G.b_l2.p, dQ, dQ + n * n, n, m, N, 64, 1e-15,
                              (double*)G.b_l3.p, P ? (double*)G.b_l4.p : nullptr, info ? (int*)G.b_st.p : nullptr));
  D2H(K, G.b_l3.p, nk);
  if (P) D2H(P, G.b_l4.p, na);
  if (info) D2H(info, G.b_st.p, (size_t)N * 8);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

// F16._calc_LQR_gain (env.py:344-358) for N operating points, every stage on the device:
// forward linearise -> reduced 9-state / 3-input model -> cont2discrete(dt) -> K = -dlqr(Ad, Bd, C'C = I, R = I)
int lqr_gain_batch(const double* x_soa, const double* u_soa, long long N, double dt, double* K /* [N][3][9] */, const unsigned char* fi,
                   int fi_default, const double* xcg, double xcg_default, int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N < 0 || (N > 0 && (!x_soa || !u_soa || !K)) || !(dt > 0)) { set_err("lqr_gain_batch: bad argument"); return F16_ERR_ARG; }
  if (N == 0) return F16_OK;
  const size_t n = (size_t)N;
  CK(G.b_in.reserve(18 * n * 8));
  CK(G.b_in2.reserve(4 * n * 8));
  CK(G.b_a.reserve(324 * n * 8));
  CK(G.b_b.reserve(72 * n * 8));
  CK(G.b_st.reserve(n * 8));
  CK(G.b_st2.reserve(n * 4));
  CK(G.b_l1.reserve(n * 81 * 8));
  CK(G.b_l2.reserve(n * 27 * 8));
  CK(G.b_l3.reserve(n * 81 * 8));
  CK(G.b_l4.reserve(n * 27 * 8));
  CK(G.b_l5.reserve((81 + 9) * 8));
  CK(G.b_out.reserve(n * 27 * 8));
  const unsigned char* d_fi;
  const double* d_xcg;
  if ((rc = stage_sel(fi, xcg, N, &d_fi, &d_xcg)) != F16_OK) return rc;
  // the reduced model evaluates at X with the actuator states replaced by the inputs (env.py:175-177)
  std::vector<double> xs(x_soa, x_soa + 18 * n);
  for (int k = 0; k < 3; k++) memcpy(&xs[(13 + k) * n], &u_soa[(1 + k) * n], n * 8);
  double QR[90];
  for (int i = 0; i < 81; i++) QR[i] = (i % 10 == 0) ? 1.0 : 0.0;
  for (int i = 0; i < 9; i++) QR[81 + i] = (i % 4 == 0) ? 1.0 : 0.0;
  CK(cudaMemcpyAsync(G.b_in.p, xs.data(), 18 * n * 8, cudaMemcpyHostToDevice, G.stream));
  CK(cudaStreamSynchronize(G.stream));  // xs is a temporary
  H2D(G.b_in2.p, u_soa, 4 * n * 8);
  CK(cudaMemcpyAsync(G.b_l5.p, QR, sizeof QR, cudaMemcpyHostToDevice, G.stream));
  CK(cudaStreamSynchronize(G.stream));  // QR is a temporary
  CK(f16::strict::launch_linearise(cfg(true), tabs(), sel_of(d_fi, fi_default, d_xcg, xcg_default), (const double*)G.b_in.p, N,
                                   (const double*)G.b_in2.p, N, N, 1e-5, F16_FD_FORWARD, (double*)G.b_a.p, (double*)G.b_b.p,
                                   (int*)G.b_st2.p));
  CK(f16::linalg::launch_reduce_jacobian(cfg(false), (const double*)G.b_a.p, N, (double*)G.b_l1.p, (double*)G.b_l2.p));
  CK(f16::linalg::launch_zoh(cfg(false), (const double*)G.b_l1.p, (const double*)G.b_l2.p, 9, 3, N, dt, (double*)G.b_l3.p,
                             (double*)G.b_l4.p));
  double* dQ = (double*)G.b_l5.p;
  CK(f16::linalg::launch_dlqr(cfg(false), (const double*)G.b_l3.p, (const double*)G.b_l4.p, dQ, dQ + 81, 9, 3, N, 64, 1e-15,
                              (double*)G.b_out.p, nullptr, (int*)G.b_st.p));
  std::vector<double> k(27 * n);
  std::vector<int> info(2 * n), st(n);
  CK(cudaMemcpyAsync(k.data(), G.b_out.p, 27 * n * 8, cudaMemcpyDeviceToHost, G.stream));
  CK(cudaMemcpyAsync(info.data(), G.b_st.p, n * 8, cudaMemcpyDeviceToHost, G.stream));
  CK(cudaMemcpyAsync(st.data(), G.b_st2.p, n * 4, cudaMemcpyDeviceToHost, G.stream));
  CK(cudaStreamSynchronize(G.stream));
  for (size_t i = 0; i < 27 * n; i++) K[i] = -k[i];  // env.py:356: K = - dlqr(A, B, Q, R)
  if (status)
    for (size_t i = 0; i < n; i++) status[i] = st[i] ? st[i] : (info[2 * i] < 0 ? (int)F16_ST_NAN : 0);
  return F16_OK;
}

// ---- parity probes -----------------------------------------------------------------------------------------------
int f16_hifi_probe(const double* alpha_deg, const double* beta_deg, const double* el, long long N, double* coef, int* cells,
                   int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N <= 0 || !alpha_deg || !beta_deg || !el || !coef || !cells) { set_err("f16_hifi_probe: bad argument"); return F16_ERR_ARG; }
  const size_t n = (size_t)N;
  CK(G.b_in.reserve(3 * n * 8));
  CK(G.b_out.reserve(44 * n * 8));
  CK(G.b_st.reserve(n * 4));
  CK(G.b_st2.reserve(8 * n * 4));
  double* d = (double*)G.b_in.p;
  H2D(d, alpha_deg, n * 8);
  H2D(d + n, beta_deg, n * 8);
  H2D(d + 2 * n, el, n * 8);
  CK(DISPATCH(launch_hifi_probe, cfg(false), tabs(), d, d + n, d + 2 * n, N, (double*)G.b_out.p, (int*)G.b_st2.p,
              (int*)G.b_st.p));
  D2H(coef, G.b_out.p, 44 * n * 8);
  D2H(cells, G.b_st2.p, 8 * n * 4);
  if (status) D2H(status, G.b_st.p, n * 4);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

int f16_fast_probe(const double* alpha_deg, const double* beta_deg, const double* el, long long N, double* coef, int* cells,
                   double* lam, int* status) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N <= 0 || !alpha_deg || !beta_deg || !el || !coef || !cells || !lam) { set_err("f16_fast_probe: bad argument"); return F16_ERR_ARG; }
  const size_t n = (size_t)N;
  CK(G.b_in.reserve(3 * n * 8));
  CK(G.b_out.reserve(48 * n * 8));
  CK(G.b_st.reserve(n * 4));
  CK(G.b_st2.reserve(4 * n * 4));
  double* d = (double*)G.b_in.p;
  H2D(d, alpha_deg, n * 8);
  H2D(d + n, beta_deg, n * 8);
  H2D(d + 2 * n, el, n * 8);
  double* o = (double*)G.b_out.p;
  CK(f16::fast::launch_fast_probe(cfg(false), tabs(), d, d + n, d + 2 * n, N, o, (int*)G.b_st2.p, o + 44 * n, (int*)G.b_st.p));
  D2H(coef, o, 44 * n * 8);
  D2H(lam, o + 44 * n, 4 * n * 8);
  D2H(cells, G.b_st2.p, 4 * n * 4);
  if (status) D2H(status, G.b_st.p, n * 4);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

int f16_lofi_probe(const double* alpha_deg, const double* beta_deg, const double* el, const double* dail,
                   const double* drud, long long N, double* out) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N <= 0 || !alpha_deg || !beta_deg || !el || !dail || !drud || !out) { set_err("f16_lofi_probe: bad argument"); return F16_ERR_ARG; }
  const size_t n = (size_t)N;
  CK(G.b_in.reserve(5 * n * 8));
  CK(G.b_out.reserve(19 * n * 8));
  double* d = (double*)G.b_in.p;
  H2D(d, alpha_deg, n * 8);
  H2D(d + n, beta_deg, n * 8);
  H2D(d + 2 * n, el, n * 8);
  H2D(d + 3 * n, dail, n * 8);
  H2D(d + 4 * n, drud, n * 8);
  CK(DISPATCH(launch_lofi_probe, cfg(false), tabs(), d, d + n, d + 2 * n, d + 3 * n, d + 4 * n, N, (double*)G.b_out.p));
  D2H(out, G.b_out.p, 19 * n * 8);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

int f16_div_probe(const double* a, const double* b, long long N, double* out) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N <= 0 || !a || !b || !out) { set_err("f16_div_probe: bad argument"); return F16_ERR_ARG; }
  const size_t n = (size_t)N;
  CK(G.b_in.reserve(2 * n * 8));
  CK(G.b_out.reserve(3 * n * 8));
  double* d = (double*)G.b_in.p;
  H2D(d, a, n * 8);
  H2D(d + n, b, n * 8);
  CK(f16::strict::launch_div_probe(cfg(false), d, d + n, N, (double*)G.b_out.p));  // always the strict build's helpers
  D2H(out, G.b_out.p, 3 * n * 8);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

int atmos_batch(const double* alt, const double* vt, long long N, double* coeff_soa) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  if (N <= 0 || !alt || !vt || !coeff_soa) { set_err("atmos_batch: bad argument"); return F16_ERR_ARG; }
  const size_t n = (size_t)N;
  CK(G.b_in.reserve(2 * n * 8));
  CK(G.b_out.reserve(3 * n * 8));
  double* d = (double*)G.b_in.p;
  H2D(d, alt, n * 8);
  H2D(d + n, vt, n * 8);
  CK(DISPATCH(launch_atmos, cfg(false), d, d + n, N, (double*)G.b_out.p));
  D2H(coeff_soa, G.b_out.p, 3 * n * 8);
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}

// ---- memory / timing helpers ------------------------------------------------------------------------------------------
void* f16_dev_alloc(unsigned long long bytes) {
  std::lock_guard<std::mutex> lk(G_mu);
  if (ensure() != F16_OK) return nullptr;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) { cuda_fail(e, "cudaMalloc"); return nullptr; }
  return p;
}
void f16_dev_free(void* p) {
  if (p) cudaFree(p);
}
void* f16_host_alloc_pinned(unsigned long long bytes) {
  std::lock_guard<std::mutex> lk(G_mu);
  if (ensure() != F16_OK) return nullptr;
  void* p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) { cuda_fail(e, "cudaHostAlloc"); return nullptr; }
  return p;
}
void f16_host_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}
int f16_memcpy_h2d(void* dst_dev, const void* src_host, unsigned long long bytes) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  CK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, G.stream));
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}
int f16_memcpy_d2h(void* dst_host, const void* src_dev, unsigned long long bytes) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  CK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, G.stream));
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}
int f16_memcpy_d2d(void* dst_dev, const void* src_dev, unsigned long long bytes) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  CK(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, G.stream));
  return F16_OK;
}
int f16_memset_dev(void* dst_dev, int value, unsigned long long bytes) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  CK(cudaMemsetAsync(dst_dev, value, bytes, G.stream));
  return F16_OK;
}
int f16_sync(void) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  CK(cudaStreamSynchronize(G.stream));
  return F16_OK;
}
int f16_timer_start(void) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  CK(cudaEventRecord(G.ev0, G.stream));
  return F16_OK;
}
int f16_timer_stop(float* ms) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  CK(cudaEventRecord(G.ev1, G.stream));
  CK(cudaEventSynchronize(G.ev1));
  CK(cudaEventElapsedTime(ms, G.ev0, G.ev1));
  return F16_OK;
}

int f16_measure_fp64_peak(double ms, double* tflops) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  CK(G.b_st.reserve(64));
  double flops = 0;
  float t = 0;
  long long iters = 2000;
  // calibrate, then run for about `ms`
  for (int pass = 0; pass < 2; pass++) {
    CK(cudaEventRecord(G.ev0, G.stream));
    CK(f16::launch_dfma_peak(G.stream, G.sm_count, iters, (double*)G.b_st.p, &flops));
    ++G.launches;
    CK(cudaEventRecord(G.ev1, G.stream));
    CK(cudaEventSynchronize(G.ev1));
    CK(cudaEventElapsedTime(&t, G.ev0, G.ev1));
    if (pass == 0) {
      double want = ms > 1 ? ms : 1;
      iters = (long long)(iters * want / (t > 1e-3f ? t : 1e-3f));
      if (iters < 100) iters = 100;
    }
  }
  *tflops = flops / (t * 1e-3) / 1e12;
  return F16_OK;
}

int f16_flush_l2(void) {
  std::lock_guard<std::mutex> lk(G_mu);
  int rc = ensure();
  if (rc != F16_OK) return rc;
  const size_t bytes = 256u << 20;  // > 126 MB L2
  CK(G.b_flush.reserve(bytes));
  CK(cudaMemsetAsync(G.b_flush.p, 0, bytes, G.stream));
  return F16_OK;
}

#pragma GCC visibility pop
}  // extern "C"

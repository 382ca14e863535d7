/* cfg5_closed_loop_multi_gpu.c -- BASELINE cfg 5 from plain C on every GPU of the box: a closed-loop LQR Monte Carlo
 * (u = u0 - K (x - x_trim) fused into the Euler step) over N aircraft, split by the LIBRARY over the device contexts that
 * f16_init_devices creates -- the caller sees one array and one call.  The gain is the reference's own
 * F16._calc_LQR_gain() (env.py:344-358), computed on the device by lqr_gain_batch at the trim point.
 *
 *   gcc -O2 -Iinclude examples/cfg5_closed_loop_multi_gpu.c -o /tmp/cfg5 -Lf16_mpc_oop_py_b200 -lf16_b200 \
 *       -Wl,-rpath,$PWD/f16_mpc_oop_py_b200 && /tmp/cfg5 [aircraft (default 1048576)] [steps (default 10000)]
 *
 * The same batch is flown twice, open loop and closed loop.  xcg 0.35 is the statically unstable airframe: open loop a part of
 * a +-5 % batch leaves the flight envelope within 10 s, the regulator keeps every aircraft inside.  Exits 0 when that is what
 * happened; 2 without a B200. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "f16_b200.h"

static double uniform(unsigned long long *s) { /* xorshift64*, [-1, 1) */
  *s ^= *s >> 12; *s ^= *s << 25; *s ^= *s >> 27;
  return (double)((*s * 2685821657736338717ULL) >> 11) / 4503599627370496.0 - 1.0;
}

int main(int argc, char **argv) {
  const long long N = argc > 1 ? atoll(argv[1]) : 1048576;
  const int K = argc > 2 ? atoi(argv[2]) : 10000;
  if (f16_init_devices(NULL, NULL, 0) != F16_OK) { /* every visible GPU */
    fprintf(stderr, "f16_init_devices: %s\n", f16_last_error());
    return 2;
  }
  f16_set_math_mode(F16_MATH_FAST);
  printf("%d device context(s)\n", f16_device_count());

  const double h = 10000.0, V = 700.0, xcg = 0.35;
  double xt[18], ut[4], Kg[27];
  int st = 0;
  if (trim_batch(&h, &V, 1, 1e-10, 50000, NULL, xt, NULL, NULL, 1, NULL, xcg, &st) != F16_OK || st != 0) return 1;
  for (int i = 0; i < 4; i++) ut[i] = xt[12 + i];
  if (lqr_gain_batch(xt, ut, 1, 0.001, Kg, NULL, 1, NULL, xcg, &st) != F16_OK || st != 0) return 1;

  /* test_env.py:294: u = u0 - K (x - x_ref) on the nine MPC states (parameters.py:135) and the three surfaces; the regulator
   * gain is -K_lqr because _calc_LQR_gain returns -dlqr(...) (env.py:356) */
  static const int mpc[9] = {3, 4, 7, 8, 9, 10, 11, 17, 16};
  f16_lqr_t law = {0};
  law.n_sel = 9;
  law.row_mask = 0xE; /* rows dh, da, dr */
  for (int j = 0; j < 9; j++) {
    law.sel[j] = mpc[j];
    law.x_ref[j] = xt[mpc[j]];
    for (int r = 0; r < 3; r++) law.K[1 + r][j] = -Kg[r * 9 + j];
  }
  for (int r = 0; r < 4; r++) law.u0[r] = ut[r];

  double *x = (double *)f16_host_alloc_pinned((unsigned long long)N * 18 * 8);
  double *u = (double *)f16_host_alloc_pinned((unsigned long long)N * 4 * 8);
  int *status = (int *)malloc((size_t)N * 4);
  if (!x || !u || !status) return 1;
  unsigned long long seed = 0xF16C5ULL;
  for (int i = 0; i < 18; i++)
    for (long long n = 0; n < N; n++) {
      const double r = 0.05 * uniform(&seed);
      x[i * N + n] = i < 2 ? 0.0 : (xt[i] != 0.0 ? xt[i] * (1.0 + r) : r);
    }
  for (int i = 0; i < 4; i++)
    for (long long n = 0; n < N; n++) u[i * N + n] = ut[i];

  double *x0 = (double *)malloc((size_t)N * 18 * 8);
  if (!x0) return 1;
  memcpy(x0, x, (size_t)N * 18 * 8);
  double open_row[74], closed_row[74];
  struct timespec t0, t1;
  for (int closed = 0; closed < 2; closed++) {
    memcpy(x, x0, (size_t)N * 18 * 8);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (step_batch(x, u, N, K, 0.001, closed ? &law : NULL, NULL, 1, NULL, xcg, status, NULL) != F16_OK) {
      fprintf(stderr, "step_batch: %s\n", f16_last_error());
      return 1;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double s = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    double *row = closed ? closed_row : open_row;
    state_summary_batch(x, N, status, row);
    printf("%s loop: %lld aircraft x %d steps in %.3f s (host wall, copies included) = %.3e aircraft-steps/s; %.0f of %.0f inside "
           "the envelope, alpha in [%.4f, %.4f] rad\n", closed ? "closed" : "open  ", N, K, s, (double)N * K / s, row[1], row[0],
           row[2 + 7], row[20 + 7]);
  }
  const int ok = closed_row[1] == (double)N && open_row[1] < (double)N;
  printf("%s\n", ok ? "closed loop holds the whole batch" : "MISMATCH");
  free(x0);
  f16_host_free_pinned(x);
  f16_host_free_pinned(u);
  free(status);
  f16_shutdown();
  return ok ? 0 : 1;
}

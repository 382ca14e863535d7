/* cfg1_open_loop.c -- BASELINE cfg 1 from plain C through the C ABI of include/f16_b200.h: one hifi F-16 (xcg 0.35),
 * trimmed at 10000 ft / 700 ft/s by trim_batch, flown open loop for 10 s of dt = 0.001 Euler steps by ONE step_batch
 * call, then linearised.  It is what the reference does in Python with test_env.py::test_control (env.py:105-130 called
 * 10000 times) and env.py::linearise -- host code in C, no Python, no PyTorch.
 *
 *   gcc -O2 -Iinclude examples/cfg1_open_loop.c -o /tmp/cfg1 -Lf16_mpc_oop_py_b200 -lf16_b200 \
 *       -Wl,-rpath,$PWD/f16_mpc_oop_py_b200 && /tmp/cfg1
 *
 * Prints the final state (SURVEY.md 8c known answer 3: npos 7000.002274 ft, h 9999.939040 ft) and exits 0 when it
 * matches; exits 2 when no B200 is usable (the library has no CPU path). */
#include <math.h>
#include <stdio.h>

#include "f16_b200.h"

int main(void) {
  if (f16_init(NULL, -1) != F16_OK) {
    fprintf(stderr, "f16_init: %s\n", f16_last_error());
    return 2;
  }
  /* env.py:198-292 -- Nelder-Mead trim, the reference's tolerance / iteration cap (env.py:273) */
  const double h = 10000.0, V = 700.0, xcg = 0.35;
  double x[18], info[4];
  int st = 0;
  if (trim_batch(&h, &V, 1, 1e-10, 50000, NULL, x, info, NULL, 1, NULL, xcg, &st) != F16_OK || st != 0) {
    fprintf(stderr, "trim_batch: %s (status %d)\n", f16_last_error(), st);
    return 1;
  }
  printf("trim: alpha %.10f rad  T %.6f lb  dh %.9f deg  cost %.3g  (%d iterations)\n", x[7], x[12], x[13], info[0], (int)info[1]);

  /* env.py:294-342 at the trim point, forward differences as the reference */
  double u[4] = {x[12], x[13], x[14], x[15]}, A[18 * 18], B[18 * 4];
  if (linearise_batch(x, u, 1, 1e-5, F16_FD_FORWARD, A, B, NULL, 1, NULL, xcg, &st) != F16_OK || st != 0) {
    fprintf(stderr, "linearise_batch: %s (status %d)\n", f16_last_error(), st);
    return 1;
  }
  printf("A[q][alpha] %.9f  A[alpha][q] %.9f  B[dh][dh_cmd] %.9f\n", A[10 * 18 + 7], A[7 * 18 + 10], B[13 * 4 + 1]);

  /* test_env.py:444-465 -- 10 s open loop, inputs held at trim */
  int steps = 0;
  if (step_batch(x, u, 1, 10000, 0.001, NULL, NULL, 1, NULL, xcg, &st, &steps) != F16_OK) {
    fprintf(stderr, "step_batch: %s\n", f16_last_error());
    return 1;
  }
  printf("after %d steps (status %d): npos %.6f  epos %.6f  h %.6f  theta %.9f  V %.7f  alpha %.9f\n", steps, st, x[0], x[1],
         x[2], x[4], x[6], x[7]);
  const int ok = st == 0 && steps == 10000 && fabs(x[0] - 7000.002274) < 2e-5 && fabs(x[2] - 9999.939040) < 2e-5;
  printf("%s\n", ok ? "matches the reference's 10 s trajectory" : "MISMATCH");
  f16_shutdown();
  return ok ? 0 : 1;
}

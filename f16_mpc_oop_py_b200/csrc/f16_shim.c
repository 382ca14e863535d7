/* f16_shim.c -- the two file names the reference loads (parameters.py:108-114):
 *     C/nlplant_xcg25.so  (-DF16_SHIM_XCG=0.25)      C/nlplant_xcg35.so  (-DF16_SHIM_XCG=0.35)
 * Each exports exactly the reference ABI (C/nlplant.c:8,14) and forwards to libf16_b200.so, found through
 * an $ORIGIN-relative rpath.  xcg is a compile-time constant in the reference too (C/nlplant.c:34). */
#ifndef F16_SHIM_XCG
#error "compile with -DF16_SHIM_XCG=0.25 or 0.35"
#endif
void f16_nlplant_xcg(const double *xu, double *xdot, int fidelity, double xcg);
void f16_atmos(double alt, double vt, double *coeff);

void Nlplant(double *xu, double *xdot, int fidelity) { f16_nlplant_xcg(xu, xdot, fidelity, F16_SHIM_XCG); }
void atmos(double alt, double vt, double *coeff) { f16_atmos(alt, vt, coeff); }

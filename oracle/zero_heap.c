/* zero_heap.c -- linked into the oracle/_ref builds of the reference with -Wl,--wrap=malloc.
 *
 * The reference interpolates one table it never fills: at C/hifi_F16_AeroData.c:965-972 the error printf after
 * `if(fp==NULL)` is commented out, so the fscanf loop that should load CL1320_ALPHA1_606.dat became the body of
 * that `if`; _CLr's DATA stays as malloc(160) returned it.  In a fresh process that memory is zero (its binaries
 * then compute Clr = 0 bit for bit); inside a long-lived process (pytest) it is recycled garbage and the reference's
 * roll/yaw derivatives become noise.  The checker must be deterministic, so the reference's malloc returns zeroed
 * memory here -- its sources are compiled unmodified, only this allocation policy is pinned. */
#include <stdlib.h>
void *__wrap_malloc(size_t n) { return calloc(1, n); }

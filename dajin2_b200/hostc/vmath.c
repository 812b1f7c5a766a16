/* Four-wide exp / log that return libm's bits, for the output pass of the MLP evaluation (fastmlp.c: fm_output).
 *
 * scipy's expit and xlogy call libm's exp and log once per training point and evaluation; after the other passes were
 * fused these two calls were the largest term of an evaluation.  glibc's double-precision exp and log (2.28 and
 * later) are the table-driven routines of ARM's optimized-routines: branch-free apart from rare special cases, a
 * 128-entry table each, short polynomials.  They are restated here on vectors of four doubles with the SAME
 * operations in the SAME order -- every lane computes what the scalar routine computes -- using the tables of the
 * libm that is loaded in this process (dajin2_b200/fastmlp.py locates them in the library file; nothing is copied
 * into this repository).  This file is compiled with -ffp-contract=fast, as glibc is: the fused multiply-adds of
 * glibc's build are the ones gcc forms from the same expressions here.
 *
 * Nothing is taken on trust: fm_vmath_init() compares both functions with libm on 2 x 10^6 arguments (the ranges the
 * MLP produces, the special-case boundaries, random bit patterns) and the vector pass is only switched on when not
 * one bit differs; every fit then still checks its first and last evaluation against scikit-learn's own
 * (fastmlp.py).  Lanes outside the fast range (|x| < 2^-54, |x| >= 512, non-finite, subnormal or non-positive log
 * arguments) are handed to libm itself. */
#include <stdint.h>

#if defined(__AVX2__) && defined(__FMA__)
#define VL 4
#define VM_NAME(name) fm_##name
#include "vmath_lanes.h" /* exp_v / log_v on four doubles, fm_vmath_init, fm_vmath_ready, fm_output_rows_vec */

#else /* no AVX2 + FMA: the scalar pass of fastmlp.c is used */

int64_t fm_vmath_init(const double *exp_hdr, const uint64_t *exp_tab, const double *log_hdr, const double *log_tab,
                      int64_t test_groups) {
    (void)exp_hdr; (void)exp_tab; (void)log_hdr; (void)log_tab; (void)test_groups;
    return -1;
}
int fm_vmath_ready(void) { return 0; }
void fm_output_rows_vec(double *z, double b, const double *y, double *delta, double *terms, int64_t lo, int64_t m) {
    (void)z; (void)b; (void)y; (void)delta; (void)terms; (void)lo; (void)m;
}
#endif

/* The eight-wide (AVX-512) form of vmath.c: the same body (vmath_lanes.h), table look-ups as gathers.  Compiled for
 * AVX-512 on any x86-64 build machine with -ffp-contract=fast and only entered after fastmlp512.c's run-time CPU check
 * and its own self test against libm (fm512_vmath_init: 0 differences required). */
#include <stdint.h>

#if defined(__x86_64__) && defined(__AVX512F__) && defined(__AVX512DQ__) && defined(__AVX512VL__) && defined(__FMA__)
#define VL 8
#define VM_NAME(name) fm512_##name
#include "vmath_lanes.h"

#else

int64_t fm512_vmath_init(const double *exp_hdr, const uint64_t *exp_tab, const double *log_hdr, const double *log_tab,
                         int64_t test_groups) {
    (void)exp_hdr; (void)exp_tab; (void)log_hdr; (void)log_tab; (void)test_groups;
    return -1;
}
int fm512_vmath_ready(void) { return 0; }
void fm512_output_rows_vec(double *z, double b, const double *y, double *delta, double *terms, int64_t lo, int64_t m) {
    (void)z; (void)b; (void)y; (void)delta; (void)terms; (void)lo; (void)m;
}
#endif

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
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>

typedef double v4d __attribute__((vector_size(32)));
typedef long long v4i __attribute__((vector_size(32)));
typedef unsigned long long v4u __attribute__((vector_size(32)));

static double EXP_HDR[8];      /* invln2N, shift, negln2hiN, negln2loN, C2..C5 */
static uint64_t EXP_TAB[256];  /* tail bits, scale bits of 2^(k/128) */
static double LOG_HDR[18];     /* ln2hi, ln2lo, A[5], B[11] */
static double LOG_TAB[256];    /* invc, logc */
static int g_ready = 0;

static inline uint64_t asu(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline v4d bcast(double x) { return (v4d){x, x, x, x}; }

static inline v4d exp4(v4d x) {
    const v4u ax = ((v4u)x >> 52) & 0x7ffULL;
    const v4u bad = (v4u)((ax - 0x3c9ULL) >= (v4u){0x3f, 0x3f, 0x3f, 0x3f});
    if (bad[0] | bad[1] | bad[2] | bad[3]) return (v4d){exp(x[0]), exp(x[1]), exp(x[2]), exp(x[3])};
    const v4d z = bcast(EXP_HDR[0]) * x;
    v4d kd = z + bcast(EXP_HDR[1]);
    const v4u ki = (v4u)kd;
    kd -= bcast(EXP_HDR[1]);
    const v4d r = x + kd * bcast(EXP_HDR[2]) + kd * bcast(EXP_HDR[3]);
    const v4u idx = 2 * (ki % 128);
    const v4u top = ki << 45;
    const v4d tail = (v4d)(v4u){EXP_TAB[idx[0]], EXP_TAB[idx[1]], EXP_TAB[idx[2]], EXP_TAB[idx[3]]};
    const v4u sbits = (v4u){EXP_TAB[idx[0] + 1], EXP_TAB[idx[1] + 1], EXP_TAB[idx[2] + 1], EXP_TAB[idx[3] + 1]} + top;
    const v4d r2 = r * r;
    const v4d tmp = tail + r + r2 * (bcast(EXP_HDR[4]) + r * bcast(EXP_HDR[5])) +
                    r2 * r2 * (bcast(EXP_HDR[6]) + r * bcast(EXP_HDR[7]));
    const v4d scale = (v4d)sbits;
    return scale + scale * tmp;
}

static inline v4d log4(v4d x) {
    const double *A = LOG_HDR + 2, *B = LOG_HDR + 7;
    const v4u ix = (v4u)x;
    const v4u top = ix >> 48;
    const v4u bad = (v4u)((top - 0x10ULL) >= (v4u){0x7fe0, 0x7fe0, 0x7fe0, 0x7fe0});
    if (bad[0] | bad[1] | bad[2] | bad[3]) return (v4d){log(x[0]), log(x[1]), log(x[2]), log(x[3])};
    const uint64_t LO = asu(1.0 - 0x1p-4), HI = asu(1.0 + 0x1.09p-4);
    const v4i near = (v4i)((ix - LO) < (v4u){HI - LO, HI - LO, HI - LO, HI - LO});
    v4d ynear = bcast(0.0), yfar = bcast(0.0);
    if (near[0] | near[1] | near[2] | near[3]) { /* arguments close to 1 */
        const v4d r = x - bcast(1.0), r2 = r * r, r3 = r * r2;
        v4d y = r3 * (bcast(B[1]) + r * bcast(B[2]) + r2 * bcast(B[3]) +
                      r3 * (bcast(B[4]) + r * bcast(B[5]) + r2 * bcast(B[6]) +
                            r3 * (bcast(B[7]) + r * bcast(B[8]) + r2 * bcast(B[9]) + r3 * bcast(B[10]))));
        v4d w = r * bcast(0x1p27);
        const v4d rhi = r + w - w, rlo = r - rhi;
        w = rhi * rhi * bcast(B[0]);
        const v4d hi = r + w;
        v4d lo = r - hi + w;
        lo += bcast(B[0]) * rlo * (rhi + r);
        y += lo;
        y += hi;
        const v4i one = (v4i)(ix == (v4u){asu(1.0), asu(1.0), asu(1.0), asu(1.0)}); /* log(1) = +0 */
        ynear = (v4d)(((v4i)y) & ~one);
    }
    if (~(near[0] & near[1] & near[2] & near[3])) {
        const v4u tmp = ix - 0x3fe6000000000000ULL;
        const v4u i = (tmp >> 45) % 128;
        const v4i k = (v4i)tmp >> 52;
        const v4u iz = ix - (tmp & (0xfffULL << 52));
        const v4d invc = (v4d){LOG_TAB[2 * i[0]], LOG_TAB[2 * i[1]], LOG_TAB[2 * i[2]], LOG_TAB[2 * i[3]]};
        const v4d logc = (v4d){LOG_TAB[2 * i[0] + 1], LOG_TAB[2 * i[1] + 1], LOG_TAB[2 * i[2] + 1], LOG_TAB[2 * i[3] + 1]};
        const v4d z = (v4d)iz;
        const v4d r = (v4d)_mm256_fmadd_pd((__m256d)z, (__m256d)invc, (__m256d)bcast(-1.0));
        const v4d kd = (v4d){(double)k[0], (double)k[1], (double)k[2], (double)k[3]};
        const v4d w = kd * bcast(LOG_HDR[0]) + logc;
        const v4d hi = w + r;
        const v4d lo = w - hi + r + kd * bcast(LOG_HDR[1]);
        const v4d r2 = r * r;
        yfar = lo + r2 * bcast(A[0]) + r * r2 * (bcast(A[1]) + r * bcast(A[2]) + r2 * (bcast(A[3]) + r * bcast(A[4]))) + hi;
    }
    return (v4d)(((v4i)ynear & near) | ((v4i)yfar & ~near));
}

/* ---- self test against the libm of this process --------------------------------------------------------- */
static uint64_t g_seed;
static inline double rnd(void) {
    g_seed ^= g_seed << 13;
    g_seed ^= g_seed >> 7;
    g_seed ^= g_seed << 17;
    return (double)(g_seed >> 11) * (1.0 / 9007199254740992.0);
}

static int64_t self_test(int64_t groups) {
    int64_t bad = 0;
    g_seed = 88172645463325252ULL;
    for (int64_t g = 0; g < groups; g++) {
        v4d x, y;
        for (int l = 0; l < 4; l++) {
            const int64_t c = g + l;
            double t = (rnd() - 0.5) * 100.0;       /* the range of -(z + b) */
            if (c % 7 == 0) t = (rnd() - 0.5) * 2e-3;
            if (c % 11 == 0) t = (rnd() - 0.5) * 1400.0; /* across the overflow / underflow boundaries */
            if (c % 101 == 0) { uint64_t u = g_seed; memcpy(&t, &u, 8); } /* any bit pattern */
            x[l] = t;
            double p = rnd();                        /* probabilities */
            if (c % 3 == 0) p = 1.0 - rnd() * 0.07;  /* the close-to-one branch */
            if (c % 5 == 0) p = exp((rnd() - 0.5) * 80.0);
            if (c % 13 == 0) p = 2.220446049250313e-16 * (1.0 + rnd());
            if (c % 1009 == 0) p = 1.0;
            if (c % 103 == 0) { uint64_t u = g_seed >> 1; memcpy(&p, &u, 8); }
            y[l] = p;
        }
        const v4d e = exp4(x), q = log4(y);
        for (int l = 0; l < 4; l++) {
            const double er = exp(x[l]), qr = log(y[l]);
            /* NaN payloads and signs are libm's own business: any NaN equals any NaN here */
            if (asu(e[l]) != asu(er) && !(e[l] != e[l] && er != er)) bad++;
            if (asu(q[l]) != asu(qr) && !(q[l] != q[l] && qr != qr)) bad++;
        }
    }
    return bad;
}

int64_t fm_vmath_init(const double *exp_hdr, const uint64_t *exp_tab, const double *log_hdr, const double *log_tab,
                      int64_t test_groups) {
    memcpy(EXP_HDR, exp_hdr, sizeof(EXP_HDR));
    memcpy(EXP_TAB, exp_tab, sizeof(EXP_TAB));
    memcpy(LOG_HDR, log_hdr, sizeof(LOG_HDR));
    memcpy(LOG_TAB, log_tab, sizeof(LOG_TAB));
    const int64_t bad = self_test(test_groups);
    g_ready = bad == 0;
    return bad;
}

int fm_vmath_ready(void) { return g_ready; }

/* The element-wise part of fm_output on rows [lo, lo + m) for labels that are exactly 0 or 1:
 *   a = 1 / (1 + exp(-(z + b)));  p = clip(a, eps, 1 - eps);  terms = log(y == 1 ? p : 1 - p);  delta = a - y
 * (for such labels xlogy(y, p) + xlogy(1 - y, 1 - p) is exactly that one logarithm: the other term is the 0 that
 * scipy's xlogy returns for a zero factor, and x + 0 = x for the non-zero logarithm of a clipped probability). */
void fm_output_rows_vec(double *z, double b, const double *y, double *delta, double *terms, int64_t lo, int64_t m) {
    const double eps = 2.220446049250313e-16, hi = 1.0 - eps;
    int64_t i = lo;
    for (; i + 4 <= lo + m; i += 4) {
        v4d v, yi;
        memcpy(&v, z + i, 32);
        memcpy(&yi, y + i, 32);
        v = v + bcast(b);
        const v4d a = bcast(1.0) / (bcast(1.0) + exp4(-v));
        v4d p = a;
        p = (v4d)(((v4i)bcast(eps) & (v4i)(a < bcast(eps))) | ((v4i)p & ~(v4i)(a < bcast(eps))));
        const v4i over = (v4i)(p > bcast(hi));
        p = (v4d)(((v4i)bcast(hi) & over) | ((v4i)p & ~over));
        const v4i is_one = (v4i)(yi == bcast(1.0));
        const v4d q = bcast(1.0) - p;
        const v4d arg = (v4d)(((v4i)p & is_one) | ((v4i)q & ~is_one));
        const v4d t = log4(arg), d = a - yi;
        memcpy(z + i, &a, 32);
        memcpy(terms + i, &t, 32);
        memcpy(delta + i, &d, 32);
    }
    for (; i < lo + m; i++) { /* the last rows of the range: the scalar statement of fm_output */
        const double v = z[i] + b;
        const double a = 1.0 / (1.0 + exp(-v));
        z[i] = a;
        const double p = a < eps ? eps : (a > hi ? hi : a);
        terms[i] = log(y[i] == 1.0 ? p : 1.0 - p);
        delta[i] = a - y[i];
    }
}

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

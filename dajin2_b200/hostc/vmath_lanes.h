/* Width-generic body of vmath.c (VL = 4: AVX2 + FMA) and vmath512.c (VL = 8: AVX-512): exp / log on VL doubles that
 * return libm's bits, their self test, and the element-wise part of the MLP's output pass.  See vmath.c for what
 * this is and why it may be trusted.  The including file defines VL and VM_NAME(name) and is compiled with
 * -ffp-contract=fast, as glibc is. */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <immintrin.h>

typedef double vd __attribute__((vector_size(8 * VL)));
typedef long long vi __attribute__((vector_size(8 * VL)));
typedef unsigned long long vu __attribute__((vector_size(8 * VL)));

static double EXP_HDR[8];      /* invln2N, shift, negln2hiN, negln2loN, C2..C5 */
static uint64_t EXP_TAB[256];  /* tail bits, scale bits of 2^(k/128) */
static double LOG_HDR[18];     /* ln2hi, ln2lo, A[5], B[11] */
static double LOG_TAB[256];    /* invc, logc */
static int g_ready = 0;

static inline uint64_t asu(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline vd bcast(double x) { return (vd){0} + x; }
static inline vu bcastu(uint64_t x) { return (vu){0} + x; }
static inline int any_lane(vi m) {
    long long acc = 0;
    for (int l = 0; l < VL; l++) acc |= m[l];
    return acc != 0;
}
static inline int all_lanes(vi m) {
    long long acc = -1;
    for (int l = 0; l < VL; l++) acc &= m[l];
    return acc != 0;
}
#if VL == 8
static inline vd fma_v(vd a, vd b, vd c) { return (vd)_mm512_fmadd_pd((__m512d)a, (__m512d)b, (__m512d)c); }
static inline vd gather_d(const double *base, vu idx) { return (vd)_mm512_i64gather_pd((__m512i)idx, base, 8); }
static inline vu gather_u(const uint64_t *base, vu idx) { return (vu)_mm512_i64gather_epi64((__m512i)idx, base, 8); }
#else
static inline vd fma_v(vd a, vd b, vd c) { return (vd)_mm256_fmadd_pd((__m256d)a, (__m256d)b, (__m256d)c); }
static inline vd gather_d(const double *base, vu idx) {
    vd r;
    for (int l = 0; l < VL; l++) r[l] = base[idx[l]];
    return r;
}
static inline vu gather_u(const uint64_t *base, vu idx) {
    vu r;
    for (int l = 0; l < VL; l++) r[l] = base[idx[l]];
    return r;
}
#endif

static inline vd exp_v(vd x) {
    const vu ax = ((vu)x >> 52) & 0x7ffULL;
    const vi bad = (vi)((ax - 0x3c9ULL) >= bcastu(0x3f));
    if (any_lane(bad)) {
        vd r;
        for (int l = 0; l < VL; l++) r[l] = exp(x[l]);
        return r;
    }
    const vd z = bcast(EXP_HDR[0]) * x;
    vd kd = z + bcast(EXP_HDR[1]);
    const vu ki = (vu)kd;
    kd -= bcast(EXP_HDR[1]);
    const vd r = x + kd * bcast(EXP_HDR[2]) + kd * bcast(EXP_HDR[3]);
    const vu idx = 2 * (ki % 128);
    const vu top = ki << 45;
    const vd tail = (vd)gather_u(EXP_TAB, idx);
    const vu sbits = gather_u(EXP_TAB + 1, idx) + top;
    const vd r2 = r * r;
    const vd tmp = tail + r + r2 * (bcast(EXP_HDR[4]) + r * bcast(EXP_HDR[5])) +
                   r2 * r2 * (bcast(EXP_HDR[6]) + r * bcast(EXP_HDR[7]));
    const vd scale = (vd)sbits;
    return scale + scale * tmp;
}

static inline vd log_v(vd x) {
    const double *A = LOG_HDR + 2, *B = LOG_HDR + 7;
    const vu ix = (vu)x;
    const vu top = ix >> 48;
    const vi bad = (vi)((top - 0x10ULL) >= bcastu(0x7fe0));
    if (any_lane(bad)) {
        vd r;
        for (int l = 0; l < VL; l++) r[l] = log(x[l]);
        return r;
    }
    const uint64_t LO = asu(1.0 - 0x1p-4), HI = asu(1.0 + 0x1.09p-4);
    const vi near = (vi)((ix - LO) < bcastu(HI - LO));
    vd ynear = bcast(0.0), yfar = bcast(0.0);
    if (any_lane(near)) { /* arguments close to 1 */
        const vd r = x - bcast(1.0), r2 = r * r, r3 = r * r2;
        vd y = r3 * (bcast(B[1]) + r * bcast(B[2]) + r2 * bcast(B[3]) +
                     r3 * (bcast(B[4]) + r * bcast(B[5]) + r2 * bcast(B[6]) +
                           r3 * (bcast(B[7]) + r * bcast(B[8]) + r2 * bcast(B[9]) + r3 * bcast(B[10]))));
        vd w = r * bcast(0x1p27);
        const vd rhi = r + w - w, rlo = r - rhi;
        w = rhi * rhi * bcast(B[0]);
        const vd hi = r + w;
        vd lo = r - hi + w;
        lo += bcast(B[0]) * rlo * (rhi + r);
        y += lo;
        y += hi;
        const vi one = (vi)(ix == bcastu(asu(1.0))); /* log(1) = +0 */
        ynear = (vd)(((vi)y) & ~one);
    }
    if (!all_lanes(near)) {
        const vu tmp = ix - 0x3fe6000000000000ULL;
        const vu i = (tmp >> 45) % 128;
        const vi k = (vi)tmp >> 52;
        const vu iz = ix - (tmp & (0xfffULL << 52));
        const vd invc = gather_d(LOG_TAB, 2 * i);
        const vd logc = gather_d(LOG_TAB + 1, 2 * i);
        const vd z = (vd)iz;
        const vd r = fma_v(z, invc, bcast(-1.0));
        const vd kd = __builtin_convertvector(k, vd);
        const vd w = kd * bcast(LOG_HDR[0]) + logc;
        const vd hi = w + r;
        const vd lo = w - hi + r + kd * bcast(LOG_HDR[1]);
        const vd r2 = r * r;
        yfar = lo + r2 * bcast(A[0]) + r * r2 * (bcast(A[1]) + r * bcast(A[2]) + r2 * (bcast(A[3]) + r * bcast(A[4]))) + hi;
    }
    return (vd)(((vi)ynear & near) | ((vi)yfar & ~near));
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
        vd x, y;
        for (int l = 0; l < VL; l++) {
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
        const vd e = exp_v(x), q = log_v(y);
        for (int l = 0; l < VL; l++) {
            const double er = exp(x[l]), qr = log(y[l]);
            /* NaN payloads and signs are libm's own business: any NaN equals any NaN here */
            if (asu(e[l]) != asu(er) && !(e[l] != e[l] && er != er)) bad++;
            if (asu(q[l]) != asu(qr) && !(q[l] != q[l] && qr != qr)) bad++;
        }
    }
    return bad;
}

int64_t VM_NAME(vmath_init)(const double *exp_hdr, const uint64_t *exp_tab, const double *log_hdr, const double *log_tab,
                            int64_t test_groups) {
    memcpy(EXP_HDR, exp_hdr, sizeof(EXP_HDR));
    memcpy(EXP_TAB, exp_tab, sizeof(EXP_TAB));
    memcpy(LOG_HDR, log_hdr, sizeof(LOG_HDR));
    memcpy(LOG_TAB, log_tab, sizeof(LOG_TAB));
    const int64_t bad = self_test(test_groups);
    g_ready = bad == 0;
    return bad;
}

int VM_NAME(vmath_ready)(void) { return g_ready; }

/* The element-wise part of fm_output on rows [lo, lo + m) for labels that are exactly 0 or 1:
 *   a = 1 / (1 + exp(-(z + b)));  p = clip(a, eps, 1 - eps);  terms = log(y == 1 ? p : 1 - p);  delta = a - y
 * (for such labels xlogy(y, p) + xlogy(1 - y, 1 - p) is exactly that one logarithm: the other term is the 0 that
 * scipy's xlogy returns for a zero factor, and x + 0 = x for the non-zero logarithm of a clipped probability). */
void VM_NAME(output_rows_vec)(double *z, double b, const double *y, double *delta, double *terms, int64_t lo, int64_t m) {
    const double eps = 2.220446049250313e-16, hi = 1.0 - eps;
    int64_t i = lo;
    for (; i + VL <= lo + m; i += VL) {
        vd v, yi;
        memcpy(&v, z + i, sizeof(vd));
        memcpy(&yi, y + i, sizeof(vd));
        v = v + bcast(b);
        const vd a = bcast(1.0) / (bcast(1.0) + exp_v(-v));
        vd p = a;
        const vi under = (vi)(a < bcast(eps));
        p = (vd)(((vi)bcast(eps) & under) | ((vi)p & ~under));
        const vi over = (vi)(p > bcast(hi));
        p = (vd)(((vi)bcast(hi) & over) | ((vi)p & ~over));
        const vi is_one = (vi)(yi == bcast(1.0));
        const vd q = bcast(1.0) - p;
        const vd arg = (vd)(((vi)p & is_one) | ((vi)q & ~is_one));
        const vd t = log_v(arg), d = a - yi;
        memcpy(z + i, &a, sizeof(vd));
        memcpy(terms + i, &t, sizeof(vd));
        memcpy(delta + i, &d, sizeof(vd));
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

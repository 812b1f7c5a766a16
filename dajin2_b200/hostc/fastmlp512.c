/* AVX-512 forms of the two one-thread passes of fastmlp.c for the reference's 2-5-2 model (forward_252, backward_252).
 *
 * Same rule as there: a vector lane is ONE output element (a row of the forward pass, a unit's chain of the backward
 * pass) and sees exactly the scalar code's operations in the scalar code's order -- fused multiply-adds where
 * OpenBLAS' dgemm makes them, plain adds and multiplies elsewhere (compiled with -ffp-contract=off) -- so the bits are
 * those of fastmlp.c's AVX2 and scalar forms (tests/test_fastmlp.py compares the three).  What changes is the
 * instruction count: the five units of the first hidden layer fit one register (no scalar side channel for unit 4),
 * the forward pass takes eight rows per trip, and the backward pass forms the output-unit deltas of eight rows at a
 * time before it walks the rows' chains.  This file is compiled for AVX-512 whatever the build machine is and is only
 * entered when the running CPU has AVX-512 F / DQ / VL (fm512_supported, checked by fastmlp.c at run time). */
#include <stdint.h>

#if defined(__x86_64__) && defined(__AVX512F__) && defined(__AVX512DQ__) && defined(__AVX512VL__)
#include <immintrin.h>

int fm512_supported(void) {
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl") &&
           __builtin_cpu_supports("fma");
}

static inline __m512d relu8(__m512d v) { /* v unless v < 0 (NaN and -0.0 pass through) */
    return _mm512_maskz_mov_pd(_mm512_cmp_pd_mask(v, _mm512_setzero_pd(), _CMP_NLT_UQ), v);
}

/* rows [0, n8) with n8 a multiple of 8; X row-major with two columns (xs_row == 2, xs_col == 1), column-major
 * (xs_row == 1) or any other element strides */
void fm512_forward_252(const double *X, int64_t xs_row, int64_t xs_col, int64_t n8, const double *W1, const double *b1,
                       const double *W2, const double *b2, double *a1, double *a2) {
    const __m512d zero = _mm512_setzero_pd();
    const __m512i even = _mm512_set_epi64(14, 12, 10, 8, 6, 4, 2, 0), odd = _mm512_set_epi64(15, 13, 11, 9, 7, 5, 3, 1);
    const __m512i il_lo = _mm512_set_epi64(11, 3, 10, 2, 9, 1, 8, 0), il_hi = _mm512_set_epi64(15, 7, 14, 6, 13, 5, 12, 4);
    /* a1 is row-major n x 5: element e = 5 r + j of a block of eight rows is unit j of row r.  Output register k holds
     * e = 8 k .. 8 k + 7, gathered from the five unit registers with two two-source permutes, a blend and a merge. */
    __m512i idx01[5], idx23[5], idx4[5];
    __mmask8 m23[5], m4[5];
    for (int k = 0; k < 5; k++) {
        long long i01[8], i23[8], i4[8];
        unsigned b23 = 0, b4 = 0;
        for (int pos = 0; pos < 8; pos++) {
            const int e = 8 * k + pos, r = e / 5, j = e % 5;
            i01[pos] = j == 1 ? 8 + r : r;
            i23[pos] = j == 3 ? 8 + r : r;
            i4[pos] = r;
            if (j == 2 || j == 3) b23 |= 1u << pos;
            if (j == 4) b4 |= 1u << pos;
        }
        idx01[k] = _mm512_loadu_si512(i01);
        idx23[k] = _mm512_loadu_si512(i23);
        idx4[k] = _mm512_loadu_si512(i4);
        m23[k] = (__mmask8)b23;
        m4[k] = (__mmask8)b4;
    }
    for (int64_t i = 0; i < n8; i += 8) {
        __m512d x0v, x1v;
        if (xs_row == 2 && xs_col == 1) {
            const __m512d p = _mm512_loadu_pd(X + 2 * i), q = _mm512_loadu_pd(X + 2 * i + 8);
            x0v = _mm512_permutex2var_pd(p, even, q);
            x1v = _mm512_permutex2var_pd(p, odd, q);
        } else if (xs_row == 1) { /* column-major X: what numpy builds from the reference's concatenation of .T views */
            x0v = _mm512_loadu_pd(X + i);
            x1v = _mm512_loadu_pd(X + i + xs_col);
        } else {
            double t0[8], t1[8];
            for (int r = 0; r < 8; r++) {
                t0[r] = X[(i + r) * xs_row];
                t1[r] = X[(i + r) * xs_row + xs_col];
            }
            x0v = _mm512_loadu_pd(t0);
            x1v = _mm512_loadu_pd(t1);
        }
        __m512d h[5];
        for (int j = 0; j < 5; j++)
            h[j] = relu8(_mm512_add_pd(_mm512_fmadd_pd(x1v, _mm512_set1_pd(W1[5 + j]),
                                                       _mm512_fmadd_pd(x0v, _mm512_set1_pd(W1[j]), zero)),
                                       _mm512_set1_pd(b1[j])));
        __m512d o0 = zero, o1 = zero;
        for (int k = 0; k < 5; k++) {
            o0 = _mm512_fmadd_pd(h[k], _mm512_set1_pd(W2[2 * k]), o0);
            o1 = _mm512_fmadd_pd(h[k], _mm512_set1_pd(W2[2 * k + 1]), o1);
        }
        o0 = relu8(_mm512_add_pd(o0, _mm512_set1_pd(b2[0])));
        o1 = relu8(_mm512_add_pd(o1, _mm512_set1_pd(b2[1])));
        for (int k = 0; k < 5; k++) {
            __m512d v = _mm512_permutex2var_pd(h[0], idx01[k], h[1]);
            v = _mm512_mask_blend_pd(m23[k], v, _mm512_permutex2var_pd(h[2], idx23[k], h[3]));
            v = _mm512_mask_permutexvar_pd(v, m4[k], idx4[k], h[4]);
            _mm512_storeu_pd(a1 + 5 * i + 8 * k, v);
        }
        _mm512_storeu_pd(a2 + 2 * i, _mm512_permutex2var_pd(o0, il_lo, o1));
        _mm512_storeu_pd(a2 + 2 * i + 8, _mm512_permutex2var_pd(o0, il_hi, o1));
    }
}

/* one row of the backward sweep: the deltas of the five units in the lanes, then the 27 chains */
#define FM512_ROW(row, E0, E1)                                                                              \
    do {                                                                                                    \
        const __m512d e0v = _mm512_set1_pd(E0), e1v = _mm512_set1_pd(E1);                                   \
        const __m512d a = _mm512_maskz_loadu_pd(m5, a1 + (row) * 5);                                        \
        __m512d d = _mm512_fmadd_pd(e1v, wb, _mm512_fmadd_pd(e0v, wa, zero));                               \
        d = _mm512_maskz_mov_pd(_mm512_cmp_pd_mask(a, zero, _CMP_NEQ_UQ), d);                               \
        c0 = _mm512_fmadd_pd(a, e0v, c0);                                                                   \
        c1 = _mm512_fmadd_pd(a, e1v, c1);                                                                   \
        p0 = _mm512_fmadd_pd(_mm512_set1_pd(X[(row) * xs_row]), d, p0);                                     \
        p1 = _mm512_fmadd_pd(_mm512_set1_pd(X[(row) * xs_row + xs_col]), d, p1);                            \
        s0 += (E0);                                                                                         \
        s1 += (E1);                                                                                         \
        t = _mm512_add_pd(t, d);                                                                            \
    } while (0)

void fm512_backward_252(const double *X, int64_t xs_row, int64_t xs_col, int64_t n, const double *a1, const double *a2,
                        const double *W2, const double *W3, const double *d3, double *cs2, double *g2, double *cs1,
                        double *g1) {
    const __mmask8 m5 = 0x1F;
    const __m512d zero = _mm512_setzero_pd();
    const __m512d wa = _mm512_set_pd(0, 0, 0, W2[8], W2[6], W2[4], W2[2], W2[0]);
    const __m512d wb = _mm512_set_pd(0, 0, 0, W2[9], W2[7], W2[5], W2[3], W2[1]);
    const double w30 = W3[0], w31 = W3[1];
    const __m512d w30v = _mm512_set1_pd(w30), w31v = _mm512_set1_pd(w31);
    const __m512i even = _mm512_set_epi64(14, 12, 10, 8, 6, 4, 2, 0), odd = _mm512_set_epi64(15, 13, 11, 9, 7, 5, 3, 1);
    __m512d c0 = zero, c1 = zero, p0 = zero, p1 = zero, t = zero; /* lanes 0..4: the chains of units 0..4 */
    double s0 = 0, s1 = 0;                                        /* the two sums of d2 */
    int64_t i = 0;
    for (; i + 8 <= n; i += 8) {
        /* d2 of eight rows: e = (a2 == 0) ? 0 : 0 + d3 * W3 (an inner dimension of 1: numpy's own loop, out = 0 + a * b) */
        const __m512d g = _mm512_loadu_pd(d3 + i);
        const __m512d p = _mm512_loadu_pd(a2 + 2 * i), q = _mm512_loadu_pd(a2 + 2 * i + 8);
        const __m512d a20 = _mm512_permutex2var_pd(p, even, q), a21 = _mm512_permutex2var_pd(p, odd, q);
        const __m512d e0 = _mm512_maskz_mov_pd(_mm512_cmp_pd_mask(a20, zero, _CMP_NEQ_UQ),
                                               _mm512_add_pd(zero, _mm512_mul_pd(g, w30v)));
        const __m512d e1 = _mm512_maskz_mov_pd(_mm512_cmp_pd_mask(a21, zero, _CMP_NEQ_UQ),
                                               _mm512_add_pd(zero, _mm512_mul_pd(g, w31v)));
        double e0b[8] __attribute__((aligned(64))), e1b[8] __attribute__((aligned(64)));
        _mm512_store_pd(e0b, e0);
        _mm512_store_pd(e1b, e1);
        for (int r = 0; r < 8; r++) FM512_ROW(i + r, e0b[r], e1b[r]);
    }
    for (; i < n; i++) {
        const double gi = d3[i];
        const double e0 = a2[2 * i] == 0.0 ? 0.0 : 0.0 + gi * w30;
        const double e1 = a2[2 * i + 1] == 0.0 ? 0.0 : 0.0 + gi * w31;
        FM512_ROW(i, e0, e1);
    }
    double v0[8], v1[8], q0[8], q1[8], tt[8];
    _mm512_storeu_pd(v0, c0);
    _mm512_storeu_pd(v1, c1);
    _mm512_storeu_pd(q0, p0);
    _mm512_storeu_pd(q1, p1);
    _mm512_storeu_pd(tt, t);
    for (int k = 0; k < 5; k++) {
        g2[2 * k] = v0[k];
        g2[2 * k + 1] = v1[k];
        g1[k] = q0[k];
        g1[5 + k] = q1[k];
        cs1[k] = tt[k];
    }
    cs2[0] = s0;
    cs2[1] = s1;
}

#else /* not an x86-64 AVX-512 build: fastmlp.c never enters this file */

int fm512_supported(void) { return 0; }
void fm512_forward_252(const double *X, int64_t xs_row, int64_t xs_col, int64_t n8, const double *W1, const double *b1,
                       const double *W2, const double *b2, double *a1, double *a2) {
    (void)X; (void)xs_row; (void)xs_col; (void)n8; (void)W1; (void)b1; (void)W2; (void)b2; (void)a1; (void)a2;
}
void fm512_backward_252(const double *X, int64_t xs_row, int64_t xs_col, int64_t n, const double *a1, const double *a2,
                        const double *W2, const double *W3, const double *d3, double *cs2, double *g2, double *cs1,
                        double *g1) {
    (void)X; (void)xs_row; (void)xs_col; (void)n; (void)a1; (void)a2; (void)W2; (void)W3; (void)d3; (void)cs2; (void)g2;
    (void)cs1; (void)g1;
}
#endif

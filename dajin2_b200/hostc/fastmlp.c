/* scikit-learn's lbfgs MLP loss/gradient (float64, two relu hidden layers, one logistic output), bit for bit.
 *
 * The reference selects candidate loci with sklearn's MLPClassifier(solver="lbfgs", hidden_layer_sizes=(5, 2))
 * (src/DAJIN2/core/preprocess/mutation_processing/anomaly_detector.py:89-92).  One loss/gradient evaluation of
 * that model on the 2 x (10 000 + L) training points is ~60 numpy calls on (n, 5) arrays and costs ~4.5 ms, almost
 * all of it numpy's per-call and per-element overhead.  This file restates the arithmetic of
 * sklearn/neural_network/_multilayer_perceptron.py (_forward_pass, _backprop, _compute_loss_grad) and _base.py
 * (binary_log_loss, inplace_relu, inplace_relu_derivative) with the SAME floating-point operations in the SAME
 * order, so that the optimiser follows the same trajectory to the last bit:
 *   - matrix products (numpy -> OpenBLAS dgemm): every output element is ONE chain of fused multiply-adds over
 *     the inner dimension, started from +0.0 (probed with scripts/micro/blas_order_probe.py: no k-blocking and no
 *     split accumulators are visible up to K = 30 000).  The two matrix x vector products of the output unit go
 *     through dgemv, whose order differs, and stay with numpy (called from Python between the stages here);
 *   - an inner dimension of 1 takes numpy's non-BLAS loop: out = 0 + a * b;
 *   - np.sum / np.mean over a contiguous axis is numpy's pairwise summation (blocks of 128, 8 accumulators),
 *     restated in pairwise_sum(); a sum over axis 0 of an (n, h > 1) array is a plain row-by-row accumulation;
 *   - expit and xlogy are scipy.special's: 1 / (1 + exp(-x)) and x * log(y) with libm's exp / log.
 * None of this is taken on trust: dajin2_b200/fastmlp.py compares loss and gradient with scikit-learn's own function
 * at the first and the last iterate of every fit and falls back to scikit-learn on any difference.
 * Compile WITHOUT -ffast-math, with -ffp-contract=off (and -mfma -mavx2 where the CPU has them: the one-thread
 * vector passes of the reference's 2-5-2 model below need both; without them the generic threaded passes run).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;
void fm_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
static int g_forward_threads = 1; /* the forward pass shares its rows among the fit's threads (0: one thread) */
void fm_set_forward_threads(int on) { g_forward_threads = on ? 1 : 0; }

/* fastmlp512.c: the one-thread passes of the 2-5-2 model on AVX-512, entered only when the running CPU has it */
int fm512_supported(void);
void fm512_forward_252(const double *X, int64_t xs_row, int64_t xs_col, int64_t n8, const double *W1, const double *b1,
                       const double *W2, const double *b2, double *a1, double *a2);
void fm512_backward_252(const double *X, int64_t xs_row, int64_t xs_col, int64_t n, const double *a1, const double *a2,
                        const double *W2, const double *W3, const double *d3, double *cs2, double *g2, double *cs1,
                        double *g1);
static int g_avx512 = -1; /* -1: not asked yet */
void fm_set_avx512(int on) { g_avx512 = (on && fm512_supported()) ? 1 : 0; }
int fm_get_avx512(void) {
    if (g_avx512 < 0) g_avx512 = fm512_supported() ? 1 : 0;
    return g_avx512;
}

/* numpy/_core/src/umath/loops_utils.h.src: @TYPE@_pairwise_sum */
static double pairwise_sum(const double *a, int64_t n) {
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        int64_t i;
        r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = a[3];
        r[4] = a[4]; r[5] = a[5]; r[6] = a[6]; r[7] = a[7];
        for (i = 8; i < n - (n % 8); i += 8) {
            r[0] += a[i + 0]; r[1] += a[i + 1]; r[2] += a[i + 2]; r[3] += a[i + 3];
            r[4] += a[i + 4]; r[5] += a[i + 5]; r[6] += a[i + 6]; r[7] += a[i + 7];
        }
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
    }
}

/* (onehot * mask).sum(axis=1) of the consensus control filter (reference consensus/similarity_searcher.py:45-62):
 * row r of the implicit matrix holds mask[i] where the code byte of read r at locus i is of class `cls` (class_of[256]:
 * 0 '+', 1 '-', 2 '*', 3 none) and 0.0 elsewhere; each row is summed the way numpy sums a contiguous float64 row. */
void dj_masked_row_sums(const uint8_t *codes, int64_t pitch, int64_t n_rows, int32_t n_loci, const uint8_t *class_of,
                        int32_t cls, const double *mask, double *out, int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
    {
        double *tmp = (double *)malloc(sizeof(double) * (size_t)(n_loci > 0 ? n_loci : 1));
#pragma omp for schedule(static)
        for (int64_t r = 0; r < n_rows; r++) {
            const uint8_t *row = codes + r * pitch;
            for (int32_t i = 0; i < n_loci; i++) tmp[i] = class_of[row[i]] == cls ? mask[i] : 0.0;
            out[r] = pairwise_sum(tmp, n_loci);
        }
        free(tmp);
    }
}

/* np.maximum(v, 0): v unless v < 0 (NaN and -0.0 pass through, as numpy's loop does) */
static inline double relu(double v) { return v < 0.0 ? 0.0 : v; }

#define FM_MAX_H 16

/* Forward pass through the two hidden layers:
 *   a1 = relu(X @ W1 + b1)   (n x h1),   a2 = relu(a1 @ W2 + b2)   (n x h2)
 * X has n_in columns and arbitrary element strides (xs_row, xs_col, in doubles); W row-major. */
/* rows [i0, i1) of the forward pass; called with literal layer sizes for the reference's model so that the inner
 * loops unroll (the OpenMP region is outside: an outlined region would turn the sizes back into variables) */
static inline __attribute__((always_inline)) void forward_rows(
    int64_t i0, int64_t i1, const double *X, int64_t xs_row, int64_t xs_col, const int n_in, const double *W1,
    const double *b1, const int h1, const double *W2, const double *b2, const int h2, double *a1, double *a2) {
    for (int64_t i = i0; i < i1; i++) {
        double *r1 = a1 + i * h1, *r2 = a2 + i * h2;
        for (int j = 0; j < h1; j++) {
            double acc = 0.0;
            for (int k = 0; k < n_in; k++) acc = fma(X[i * xs_row + k * xs_col], W1[k * h1 + j], acc);
            r1[j] = relu(acc + b1[j]);
        }
        for (int j = 0; j < h2; j++) {
            double acc = 0.0;
            for (int k = 0; k < h1; k++) acc = fma(r1[k], W2[k * h2 + j], acc);
            r2[j] = relu(acc + b2[j]);
        }
    }
}

/* ---- the reference's model (2 inputs, hidden layers of 5 and 2) on ONE thread with AVX2 + FMA: the lanes of a vector
 * are different output elements (the units of a layer, the chains of a reduction), never a split of one sum, so every
 * element sees the scalar code's operations in the scalar code's order.  Measured on the 16-core host of a B200 box:
 * spreading these two passes over threads makes the arrays they share (a1, a2 and the deltas) travel from core to
 * core every evaluation, and that traffic, not the arithmetic, bounded them (scripts/mlp_breakdown.py). ---- */
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define FM_HAVE_252_AVX2 1

static inline __m256d relu4(__m256d v) { /* v unless v < 0 (NaN and -0.0 pass through) */
    return _mm256_blendv_pd(v, _mm256_setzero_pd(), _mm256_cmp_pd(v, _mm256_setzero_pd(), _CMP_LT_OQ));
}

static void forward_252(const double *X, int64_t xs_row, int64_t xs_col, int64_t n, const double *W1, const double *b1,
                        const double *W2, const double *b2, double *a1, double *a2) {
    const __m256d w0 = _mm256_loadu_pd(W1), w1 = _mm256_loadu_pd(W1 + 5), bb = _mm256_loadu_pd(b1);
    const double w04 = W1[4], w14 = W1[9], b4 = b1[4];
    const __m128d b2v = _mm_loadu_pd(b2);
    __m128d w2[5];
    for (int k = 0; k < 5; k++) w2[k] = _mm_loadu_pd(W2 + 2 * k);
    const __m256d zero = _mm256_setzero_pd();
    int64_t i = 0;
    /* four rows at a time, one row per lane (each lane = the scalar code on its row), results transposed back into the
     * row-major a1 / a2 the backward pass and numpy read */
    for (; i + 4 <= n; i += 4) {
        __m256d x0v, x1v;
        if (xs_row == 2 && xs_col == 1) { /* row-major X, as scikit-learn hands it over */
            const __m256d p = _mm256_loadu_pd(X + 2 * i), q = _mm256_loadu_pd(X + 2 * i + 4);
            x0v = _mm256_permute4x64_pd(_mm256_unpacklo_pd(p, q), 0xD8);
            x1v = _mm256_permute4x64_pd(_mm256_unpackhi_pd(p, q), 0xD8);
        } else if (xs_row == 1) { /* column-major X: what numpy builds from the reference's concatenation of .T views */
            x0v = _mm256_loadu_pd(X + i);
            x1v = _mm256_loadu_pd(X + i + xs_col);
        } else {
            x0v = _mm256_set_pd(X[(i + 3) * xs_row], X[(i + 2) * xs_row], X[(i + 1) * xs_row], X[i * xs_row]);
            x1v = _mm256_set_pd(X[(i + 3) * xs_row + xs_col], X[(i + 2) * xs_row + xs_col], X[(i + 1) * xs_row + xs_col],
                                X[i * xs_row + xs_col]);
        }
        __m256d h[5];
        for (int j = 0; j < 5; j++)
            h[j] = relu4(_mm256_add_pd(_mm256_fmadd_pd(x1v, _mm256_set1_pd(W1[5 + j]),
                                                       _mm256_fmadd_pd(x0v, _mm256_set1_pd(W1[j]), zero)),
                                       _mm256_set1_pd(b1[j])));
        __m256d o0 = zero, o1 = zero;
        for (int k = 0; k < 5; k++) {
            o0 = _mm256_fmadd_pd(h[k], _mm256_set1_pd(W2[2 * k]), o0);
            o1 = _mm256_fmadd_pd(h[k], _mm256_set1_pd(W2[2 * k + 1]), o1);
        }
        o0 = relu4(_mm256_add_pd(o0, _mm256_set1_pd(b2[0])));
        o1 = relu4(_mm256_add_pd(o1, _mm256_set1_pd(b2[1])));
        const __m256d t0 = _mm256_unpacklo_pd(h[0], h[1]), t1 = _mm256_unpackhi_pd(h[0], h[1]);
        const __m256d t2 = _mm256_unpacklo_pd(h[2], h[3]), t3 = _mm256_unpackhi_pd(h[2], h[3]);
        _mm256_storeu_pd(a1 + 5 * i, _mm256_permute2f128_pd(t0, t2, 0x20));
        _mm256_storeu_pd(a1 + 5 * i + 5, _mm256_permute2f128_pd(t1, t3, 0x20));
        _mm256_storeu_pd(a1 + 5 * i + 10, _mm256_permute2f128_pd(t0, t2, 0x31));
        _mm256_storeu_pd(a1 + 5 * i + 15, _mm256_permute2f128_pd(t1, t3, 0x31));
        double h4[4];
        _mm256_storeu_pd(h4, h[4]);
        a1[5 * i + 4] = h4[0];
        a1[5 * i + 9] = h4[1];
        a1[5 * i + 14] = h4[2];
        a1[5 * i + 19] = h4[3];
        const __m256d lo = _mm256_unpacklo_pd(o0, o1), hi = _mm256_unpackhi_pd(o0, o1);
        _mm256_storeu_pd(a2 + 2 * i, _mm256_permute2f128_pd(lo, hi, 0x20));
        _mm256_storeu_pd(a2 + 2 * i + 4, _mm256_permute2f128_pd(lo, hi, 0x31));
    }
    for (; i < n; i++) { /* the last rows, one at a time with the units of a layer in the lanes */
        const double x0 = X[i * xs_row], x1 = X[i * xs_row + xs_col];
        const __m256d acc = _mm256_fmadd_pd(_mm256_set1_pd(x1), w1, _mm256_fmadd_pd(_mm256_set1_pd(x0), w0, zero));
        const __m256d h = relu4(_mm256_add_pd(acc, bb));
        const double h4 = relu(fma(x1, w14, fma(x0, w04, 0.0)) + b4);
        _mm256_storeu_pd(a1 + i * 5, h);
        a1[i * 5 + 4] = h4;
        double hh[4];
        _mm256_storeu_pd(hh, h);
        __m128d o = _mm_fmadd_pd(_mm_set1_pd(hh[0]), w2[0], _mm_setzero_pd());
        o = _mm_fmadd_pd(_mm_set1_pd(hh[1]), w2[1], o);
        o = _mm_fmadd_pd(_mm_set1_pd(hh[2]), w2[2], o);
        o = _mm_fmadd_pd(_mm_set1_pd(hh[3]), w2[3], o);
        o = _mm_fmadd_pd(_mm_set1_pd(h4), w2[4], o);
        o = _mm_add_pd(o, b2v);
        o = _mm_blendv_pd(o, _mm_setzero_pd(), _mm_cmplt_pd(o, _mm_setzero_pd()));
        _mm_storeu_pd(a2 + i * 2, o);
    }
}

/* Backward pass and all 27 reduction chains in one sweep: d2 = d3 W3^T masked by a2 and d1 = d2 W2^T masked by a1 are
 * formed per row in registers and feed g2 = a1^T d2, g1 = X^T d1 and the column sums of d2 and d1 directly. */
static void backward_252(const double *X, int64_t xs_row, int64_t xs_col, int64_t n, const double *a1, const double *a2,
                         const double *W2, const double *W3, const double *d3, double *cs2, double *g2, double *cs1,
                         double *g1) {
    const __m256d zero = _mm256_setzero_pd();
    const __m256d wa = _mm256_set_pd(W2[6], W2[4], W2[2], W2[0]), wb = _mm256_set_pd(W2[7], W2[5], W2[3], W2[1]);
    const double wa4 = W2[8], wb4 = W2[9], w30 = W3[0], w31 = W3[1];
    __m256d c0 = zero, c1 = zero, p0 = zero, p1 = zero, t = zero; /* chains of units 0..3 */
    double c40 = 0, c41 = 0, p04 = 0, p14 = 0, t4 = 0, s0 = 0, s1 = 0; /* unit 4 and the two sums of d2 */
    for (int64_t i = 0; i < n; i++) {
        const double g = d3[i];
        const double e0 = a2[2 * i] == 0.0 ? 0.0 : 0.0 + g * w30;
        const double e1 = a2[2 * i + 1] == 0.0 ? 0.0 : 0.0 + g * w31;
        const __m256d e0v = _mm256_set1_pd(e0), e1v = _mm256_set1_pd(e1);
        const __m256d a = _mm256_loadu_pd(a1 + i * 5);
        const double a4 = a1[i * 5 + 4];
        __m256d d = _mm256_fmadd_pd(e1v, wb, _mm256_fmadd_pd(e0v, wa, zero));
        d = _mm256_blendv_pd(d, zero, _mm256_cmp_pd(a, zero, _CMP_EQ_OQ));
        const double d4 = a4 == 0.0 ? 0.0 : fma(e1, wb4, fma(e0, wa4, 0.0));
        const double x0 = X[i * xs_row], x1 = X[i * xs_row + xs_col];
        c0 = _mm256_fmadd_pd(a, e0v, c0);
        c1 = _mm256_fmadd_pd(a, e1v, c1);
        c40 = fma(a4, e0, c40);
        c41 = fma(a4, e1, c41);
        p0 = _mm256_fmadd_pd(_mm256_set1_pd(x0), d, p0);
        p1 = _mm256_fmadd_pd(_mm256_set1_pd(x1), d, p1);
        p04 = fma(x0, d4, p04);
        p14 = fma(x1, d4, p14);
        s0 += e0;
        s1 += e1;
        t = _mm256_add_pd(t, d);
        t4 += d4;
    }
    double v0[4], v1[4];
    _mm256_storeu_pd(v0, c0);
    _mm256_storeu_pd(v1, c1);
    for (int k = 0; k < 4; k++) {
        g2[2 * k] = v0[k];
        g2[2 * k + 1] = v1[k];
    }
    g2[8] = c40;
    g2[9] = c41;
    _mm256_storeu_pd(g1, p0);
    g1[4] = p04;
    _mm256_storeu_pd(g1 + 5, p1);
    g1[9] = p14;
    cs2[0] = s0;
    cs2[1] = s1;
    _mm256_storeu_pd(cs1, t);
    cs1[4] = t4;
}
#endif

#define FM_CHUNKS 64

int fm_forward(const double *X, int64_t xs_row, int64_t xs_col, int64_t n, int n_in, const double *W1, const double *b1,
               int h1, const double *W2, const double *b2, int h2, double *a1, double *a2) {
    if (h1 > FM_MAX_H || h2 > FM_MAX_H || n_in > FM_MAX_H) return 1;
    const int64_t step = (n + FM_CHUNKS - 1) / FM_CHUNKS;
#ifdef FM_HAVE_252_AVX2
    if (n_in == 2 && h1 == 5 && h2 == 2) {
        int64_t done = 0;
        if (fm_get_avx512()) { /* eight rows per trip; the last n % 8 rows go through the four-row form */
            done = n - n % 8;
            const int T = g_forward_threads ? g_threads : 1;
            if (T > 1 && done >= 4096) { /* rows are independent: the same trips, shared among the fit's threads */
                const int64_t per = ((done / 8 + T - 1) / T) * 8;
#pragma omp parallel for num_threads(T) schedule(static)
                for (int t = 0; t < T; t++) {
                    const int64_t lo = t * per, hi = lo + per < done ? lo + per : done;
                    if (lo < hi)
                        fm512_forward_252(X + lo * xs_row, xs_row, xs_col, hi - lo, W1, b1, W2, b2, a1 + 5 * lo, a2 + 2 * lo);
                }
            } else {
                fm512_forward_252(X, xs_row, xs_col, done, W1, b1, W2, b2, a1, a2);
            }
        }
        if (done < n) forward_252(X + done * xs_row, xs_row, xs_col, n - done, W1, b1, W2, b2, a1 + 5 * done, a2 + 2 * done);
        return 0;
    }
#endif
    if (n_in == 2 && h1 == 5 && h2 == 2) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int c = 0; c < FM_CHUNKS; c++)
            forward_rows(c * step, (c + 1) * step < n ? (c + 1) * step : n, X, xs_row, xs_col, 2, W1, b1, 5, W2, b2, 2, a1, a2);
    } else {
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int c = 0; c < FM_CHUNKS; c++)
            forward_rows(c * step, (c + 1) * step < n ? (c + 1) * step : n, X, xs_row, xs_col, n_in, W1, b1, h1, W2, b2, h2,
                         a1, a2);
    }
    return 0;
}

/* Output unit: a = expit(z + b); binary_log_loss(y, a); delta = a - y.
 * Returns -mean(xlogy(y, p) + xlogy(1 - y, 1 - p)) with p = clip(a, eps, 1 - eps); *sum_delta = np.sum(delta, axis=0).
 * z is overwritten with a; `terms` is scratch of n doubles. */
/* numpy's pairwise sum splits a long array in halves (the first rounded down to a multiple of 8) until at most 128
 * elements are left: the leaves are independent sums, only their combination is a fixed tree.  So the leaves are
 * summed by the threads that produced their elements, and one thread adds the leaf sums in numpy's order. */
static int64_t leaves_of(int64_t lo, int64_t n, int64_t *off, int64_t *len, int64_t k) {
    if (n <= 128) {
        off[k] = lo;
        len[k] = n;
        return k + 1;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    k = leaves_of(lo, n2, off, len, k);
    return leaves_of(lo + n2, n - n2, off, len, k);
}

static double combine_leaves(const double *leaf, int64_t n, int64_t *cursor) {
    if (n <= 128) return leaf[(*cursor)++];
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    const double left = combine_leaves(leaf, n2, cursor);
    return left + combine_leaves(leaf, n - n2, cursor);
}

static int64_t g_leaf_n = -1, g_n_leaves = 0, *g_leaf_off = 0, *g_leaf_len = 0;
static double *g_leaf_a = 0, *g_leaf_b = 0;

/* vmath.c: the same element-wise arithmetic four rows at a time, with exp / log that return libm's bits */
int fm_vmath_ready(void);
void fm_output_rows_vec(double *z, double b, const double *y, double *delta, double *terms, int64_t lo, int64_t m);
int fm512_vmath_ready(void); /* vmath512.c: eight rows at a time */
void fm512_output_rows_vec(double *z, double b, const double *y, double *delta, double *terms, int64_t lo, int64_t m);

double fm_output(double *z, double b, const double *y, int64_t n, double *delta, double *terms, double *sum_delta,
                 int use_vec) {
    const double eps = 2.220446049250313e-16;
    const double hi = 1.0 - eps;
    /* the caller asks for it only when every label is exactly 0 or 1; 2 = the eight-wide form */
    const int vec = !use_vec ? 0 : (use_vec == 2 && fm_get_avx512() && fm512_vmath_ready()) ? 2 : fm_vmath_ready();
    if (n != g_leaf_n) { /* the leaf table of this length (one training-set size per fit) */
        const int64_t cap = n / 32 + 4;
        free(g_leaf_off); free(g_leaf_len); free(g_leaf_a); free(g_leaf_b);
        g_leaf_off = (int64_t *)malloc(sizeof(int64_t) * cap);
        g_leaf_len = (int64_t *)malloc(sizeof(int64_t) * cap);
        g_leaf_a = (double *)malloc(sizeof(double) * cap);
        g_leaf_b = (double *)malloc(sizeof(double) * cap);
        g_n_leaves = n > 0 ? leaves_of(0, n, g_leaf_off, g_leaf_len, 0) : 0;
        g_leaf_n = n;
    }
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (int64_t l = 0; l < g_n_leaves; l++) {
        const int64_t lo = g_leaf_off[l], m = g_leaf_len[l];
        if (vec == 2) fm512_output_rows_vec(z, b, y, delta, terms, lo, m);
        else if (vec) fm_output_rows_vec(z, b, y, delta, terms, lo, m);
        else for (int64_t i = lo; i < lo + m; i++) {
            double v = z[i] + b;
            double a = 1.0 / (1.0 + exp(-v));
            z[i] = a;
            double p = a < eps ? eps : (a > hi ? hi : a);
            double yi = y[i], t1, t2;
            /* scipy.special.xlogy: 0 when x == 0 and y is not NaN, else x * log(y) */
            t1 = (yi == 0.0 && p == p) ? 0.0 : yi * log(p);
            double q = 1.0 - p, w = 1.0 - yi;
            t2 = (w == 0.0 && q == q) ? 0.0 : w * log(q);
            terms[i] = t1 + t2;
            delta[i] = a - yi;
        }
        g_leaf_a[l] = pairwise_sum(delta + lo, m);
        g_leaf_b[l] = pairwise_sum(terms + lo, m);
    }
    if (n <= 0) {
        *sum_delta = 0.0;
        return -(0.0 / (double)n);
    }
    int64_t cursor = 0;
    *sum_delta = combine_leaves(g_leaf_a, n, &cursor);
    cursor = 0;
    return -(combine_leaves(g_leaf_b, n, &cursor) / (double)n);
}

/* Backward pass below the output unit (d3 = delta of the output unit, n x 1; W3 = h2 weights of the output unit):
 *   d2 = (d3 @ W3.T) with d2[a2 == 0] = 0          cs2 = sum(d2, axis=0)        g2 = a1.T @ d2   (h1 x h2)
 *   d1 = (d2 @ W2.T) with d1[a1 == 0] = 0          cs1 = sum(d1, axis=0)        g1 = X.T @ d1    (n_in x h1)
 * d2 / d1 are scratch (n x h2, n x h1).  The column sums and the products over n are sequential chains. */
static inline __attribute__((always_inline)) void backward_rows(
    int64_t i0, int64_t i1, const double *a1, const int h1, const double *a2, const int h2, const double *W2,
    const double *W3, const double *d3, double *d2, double *d1) {
    for (int64_t i = i0; i < i1; i++) {
        double *r2 = d2 + i * h2, *r1 = d1 + i * h1;
        for (int j = 0; j < h2; j++) {
            double v = 0.0 + d3[i] * W3[j];
            r2[j] = a2[i * h2 + j] == 0.0 ? 0.0 : v;
        }
        for (int j = 0; j < h1; j++) {
            double acc = 0.0;
            for (int k = 0; k < h2; k++) acc = fma(r2[k], W2[j * h2 + k], acc);
            r1[j] = a1[i * h1 + j] == 0.0 ? 0.0 : acc;
        }
    }
}

/* The reductions over the n rows: one sequential chain per output element.  Chains are independent of each other, so
 * they are dealt to threads in four tasks (a chain is never split: its order is the result). */
static inline __attribute__((always_inline)) void reduce_colsums(int64_t n, const int h1, const int h2, const double *d2,
                                                                 const double *d1, double *cs2, double *cs1) {
    double s2[FM_MAX_H], s1[FM_MAX_H];
    for (int j = 0; j < h2; j++) s2[j] = 0.0;
    for (int j = 0; j < h1; j++) s1[j] = 0.0;
    for (int64_t i = 0; i < n; i++) {
        for (int j = 0; j < h2; j++) s2[j] += d2[i * h2 + j];
        for (int j = 0; j < h1; j++) s1[j] += d1[i * h1 + j];
    }
    for (int j = 0; j < h2; j++) cs2[j] = s2[j];
    for (int j = 0; j < h1; j++) cs1[j] = s1[j];
}

/* g[k, j] = sum_i left[i, k] * right[i, j] for k in [k0, k1): fma chains over i (left has element strides) */
static inline __attribute__((always_inline)) void reduce_product(int64_t n, const double *left, int64_t ls_row,
                                                                 int64_t ls_col, const int k0, const int k1,
                                                                 const double *right, const int h, double *g) {
    double acc[FM_MAX_H * FM_MAX_H];
    for (int c = 0; c < (k1 - k0) * h; c++) acc[c] = 0.0;
    for (int64_t i = 0; i < n; i++) {
        for (int k = k0; k < k1; k++) {
            const double x = left[i * ls_row + k * ls_col];
            for (int j = 0; j < h; j++) acc[(k - k0) * h + j] = fma(x, right[i * h + j], acc[(k - k0) * h + j]);
        }
    }
    for (int k = k0; k < k1; k++)
        for (int j = 0; j < h; j++) g[k * h + j] = acc[(k - k0) * h + j];
}

/* one of the four reduction tasks (generic layer sizes) */
static inline __attribute__((always_inline)) void reduce_task(
    const int task, const double *X, int64_t xs_row, int64_t xs_col, int64_t n, const int n_in, const int half,
    const double *a1, const int h1, const int h2, const double *d2, const double *d1, double *cs2, double *g2,
    double *cs1, double *g1) {
    if (task == 0) reduce_product(n, a1, h1, 1, 0, h1, d2, h2, g2);
    else if (task == 1) reduce_product(n, X, xs_row, xs_col, 0, half, d1, h1, g1);
    else if (task == 2) reduce_product(n, X, xs_row, xs_col, half, n_in, d1, h1, g1);
    else reduce_colsums(n, h1, h2, d2, d1, cs2, cs1);
}

/* The same four tasks for the reference's model (2 inputs, hidden layers of 5 and 2) with every chain in a register:
 * the generic loops above keep their accumulators in memory, which puts a store-to-load round trip into each step of
 * each chain (4x slower).  The operations and their order per chain are the same. */
static void reduce_task_252(const int task, const double *X, int64_t xs_row, int64_t xs_col, int64_t n, const double *a1,
                            const double *d2, const double *d1, double *cs2, double *g2, double *cs1, double *g1) {
    if (task == 0) { /* g2[k][j] = sum_i a1[i][k] * d2[i][j] */
        double c00 = 0, c01 = 0, c10 = 0, c11 = 0, c20 = 0, c21 = 0, c30 = 0, c31 = 0, c40 = 0, c41 = 0;
        for (int64_t i = 0; i < n; i++) {
            const double *x = a1 + i * 5, y0 = d2[i * 2], y1 = d2[i * 2 + 1];
            c00 = fma(x[0], y0, c00); c01 = fma(x[0], y1, c01);
            c10 = fma(x[1], y0, c10); c11 = fma(x[1], y1, c11);
            c20 = fma(x[2], y0, c20); c21 = fma(x[2], y1, c21);
            c30 = fma(x[3], y0, c30); c31 = fma(x[3], y1, c31);
            c40 = fma(x[4], y0, c40); c41 = fma(x[4], y1, c41);
        }
        g2[0] = c00; g2[1] = c01; g2[2] = c10; g2[3] = c11; g2[4] = c20; g2[5] = c21; g2[6] = c30; g2[7] = c31;
        g2[8] = c40; g2[9] = c41;
    } else if (task == 1 || task == 2) { /* g1[k][j] = sum_i X[i][k] * d1[i][j] for k = task - 1 */
        const int k = task - 1;
        double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
        for (int64_t i = 0; i < n; i++) {
            const double x = X[i * xs_row + k * xs_col], *y = d1 + i * 5;
            c0 = fma(x, y[0], c0); c1 = fma(x, y[1], c1); c2 = fma(x, y[2], c2); c3 = fma(x, y[3], c3);
            c4 = fma(x, y[4], c4);
        }
        double *o = g1 + k * 5;
        o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3; o[4] = c4;
    } else { /* column sums of d2 and d1 */
        double s0 = 0, s1 = 0, t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
        for (int64_t i = 0; i < n; i++) {
            const double *p = d2 + i * 2, *q = d1 + i * 5;
            s0 += p[0]; s1 += p[1];
            t0 += q[0]; t1 += q[1]; t2 += q[2]; t3 += q[3]; t4 += q[4];
        }
        cs2[0] = s0; cs2[1] = s1;
        cs1[0] = t0; cs1[1] = t1; cs1[2] = t2; cs1[3] = t3; cs1[4] = t4;
    }
}

int fm_backward(const double *X, int64_t xs_row, int64_t xs_col, int64_t n, int n_in, const double *a1, int h1,
                const double *a2, int h2, const double *W2, const double *W3, const double *d3, double *d2, double *d1,
                double *cs2, double *g2, double *cs1, double *g1) {
    if (h1 > FM_MAX_H || h2 > FM_MAX_H || n_in > FM_MAX_H) return 1;
    const int64_t step = (n + FM_CHUNKS - 1) / FM_CHUNKS;
#ifdef FM_HAVE_252_AVX2
    if (n_in == 2 && h1 == 5 && h2 == 2) {
        /* (The 27 chains are independent and could be walked in groups by the fit's threads without changing a bit.
         * Measured, that loses badly -- fit 58 -> 117 ms on the 16-core Xeon of the B200 boxes, 76 -> 110 ms in the
         * build container: every thread then pulls all of a1 / a2 / d3 out of the caches of the cores that wrote
         * them, and the next forward pass takes the lines back.) */
        if (fm_get_avx512()) fm512_backward_252(X, xs_row, xs_col, n, a1, a2, W2, W3, d3, cs2, g2, cs1, g1);
        else backward_252(X, xs_row, xs_col, n, a1, a2, W2, W3, d3, cs2, g2, cs1, g1);
        return 0;
    }
#endif
    if (n_in == 2 && h1 == 5 && h2 == 2) {
#pragma omp parallel num_threads(g_threads)
        {
#pragma omp for schedule(static)
            for (int c = 0; c < FM_CHUNKS; c++)
                backward_rows(c * step, (c + 1) * step < n ? (c + 1) * step : n, a1, 5, a2, 2, W2, W3, d3, d2, d1);
#pragma omp for schedule(dynamic, 1)
            for (int task = 0; task < 4; task++)
                reduce_task_252(task, X, xs_row, xs_col, n, a1, d2, d1, cs2, g2, cs1, g1);
        }
    } else {
        const int half = (n_in + 1) / 2;
#pragma omp parallel num_threads(g_threads)
        {
#pragma omp for schedule(static)
            for (int c = 0; c < FM_CHUNKS; c++)
                backward_rows(c * step, (c + 1) * step < n ? (c + 1) * step : n, a1, h1, a2, h2, W2, W3, d3, d2, d1);
#pragma omp for schedule(dynamic, 1)
            for (int task = 0; task < 4; task++)
                reduce_task(task, X, xs_row, xs_col, n, n_in, half, a1, h1, h2, d2, d1, cs2, g2, cs1, g1);
        }
    }
    return 0;
}

/* ---- entry points on scikit-learn's packed parameter vector: [W1 | W2 | W3 | b1 | b2 | b3], row-major ---- */

int fm_forward_packed(const double *X, int64_t xs_row, int64_t xs_col, int64_t n, int n_in, int h1, int h2,
                      const double *packed, double *a1, double *a2) {
    const double *W1 = packed, *W2 = W1 + n_in * h1, *W3 = W2 + h1 * h2, *b1 = W3 + h2, *b2 = b1 + h1;
    return fm_forward(X, xs_row, xs_col, n, n_in, W1, b1, h1, W2, b2, h2, a1, a2);
}

/* g3_raw = a2.T @ d3 (numpy's dgemv), sum3 = np.sum(d3, axis=0).  Writes the packed gradient:
 * coef_grads[l] = (raw + alpha * W_l) / n, intercept_grads[l] = colsum / n (_compute_loss_grad). */
int fm_backward_packed(const double *X, int64_t xs_row, int64_t xs_col, int64_t n, int n_in, int h1, int h2,
                       const double *packed, const double *a1, const double *a2, const double *d3, double *d2, double *d1,
                       const double *g3_raw, double sum3, double alpha, double *grad) {
    const double *W1 = packed, *W2 = W1 + n_in * h1, *W3 = W2 + h1 * h2;
    double cs1[FM_MAX_H], cs2[FM_MAX_H], g1[FM_MAX_H * FM_MAX_H], g2[FM_MAX_H * FM_MAX_H];
    int rc = fm_backward(X, xs_row, xs_col, n, n_in, a1, h1, a2, h2, W2, W3, d3, d2, d1, cs2, g2, cs1, g1);
    if (rc) return rc;
    const double dn = (double)n;
    double *o = grad;
    for (int c = 0; c < n_in * h1; c++) *o++ = (g1[c] + alpha * W1[c]) / dn;
    for (int c = 0; c < h1 * h2; c++) *o++ = (g2[c] + alpha * W2[c]) / dn;
    for (int c = 0; c < h2; c++) *o++ = (g3_raw[c] + alpha * W3[c]) / dn;
    for (int c = 0; c < h1; c++) *o++ = cs1[c] / dn;
    for (int c = 0; c < h2; c++) *o++ = cs2[c] / dn;
    *o++ = sum3 / dn;
    return 0;
}

double fm_pairwise_sum(const double *a, int64_t n) { return pairwise_sum(a, n); }

/* ---- one call per evaluation ------------------------------------------------------------------------------------
 * _fast_loss_grad (dajin2_b200/fastmlp.py) made three calls into this file per evaluation and, between them, two
 * matrix-vector products and three dot products through numpy (numpy -> OpenBLAS): ~70 us of interpreter time around
 * 0.3 ms of arithmetic.  fm_eval does the same sequence in one call.  The two products and the dots are NOT restated:
 * they are made by the very OpenBLAS instance numpy is linked against, through its own entry points (cblas_dgemv /
 * cblas_ddot, handed over as function pointers), with the arguments numpy's matmul / dot pass for these shapes
 * (numpy/_core/src/umath/matmul.c.src, DOUBLE_gemv: a C-contiguous (n, h) matrix times a column goes out as
 * ColMajor / Trans with lda = h; its transposed view times a column as RowMajor / Trans with lda = h) -- same code,
 * same arguments, same bits; fastmlp.py compares both against numpy at load time and keeps the three-call path
 * otherwise, and every fit still compares its evaluation with scikit-learn's. */
typedef void (*fm_dgemv_fn)(int order, int trans, int64_t m, int64_t n, double alpha, const double *a, int64_t lda,
                            const double *x, int64_t incx, double beta, double *y, int64_t incy);
typedef double (*fm_ddot_fn)(int64_t n, const double *x, int64_t incx, const double *y, int64_t incy);

typedef struct {
    const double *X;
    int64_t xs_row, xs_col, n;
    int32_t n_in, h1, h2, vec;
    const double *y;
    double *a1, *a2, *a3, *d3, *d2, *d1, *terms;
    double alpha;
    fm_dgemv_fn dgemv;
    fm_ddot_fn ddot;
} FmEvalCtx;

/* loss (return value) and packed gradient of MLPClassifier._loss_grad_lbfgs for the 2-hidden-layer logistic model */
double fm_eval(const FmEvalCtx *c, const double *packed, double *grad, int *status) {
    const int n_in = c->n_in, h1 = c->h1, h2 = c->h2;
    const int64_t n = c->n;
    const int o_w3 = n_in * h1 + h1 * h2, o_b3 = o_w3 + h2 + h1 + h2;
    *status = fm_forward_packed(c->X, c->xs_row, c->xs_col, n, n_in, h1, h2, packed, c->a1, c->a2);
    if (*status) return 0.0;
    /* a3 = a2 @ W3 */
    c->dgemv(102 /* ColMajor */, 112 /* Trans */, h2, n, 1.0, c->a2, h2, packed + o_w3, 1, 0.0, c->a3, 1);
    double sum3 = 0.0;
    double loss = fm_output(c->a3, packed[o_b3], c->y, n, c->d3, c->terms, &sum3, c->vec);
    /* values = 0; for s in coefs_: values += np.dot(s.ravel(), s.ravel());  loss += (0.5 * alpha) * values / n */
    double values = 0.0;
    const int sizes[3] = {n_in * h1, h1 * h2, h2};
    int lo = 0;
    for (int k = 0; k < 3; k++) {
        values += 0.0 + c->ddot(sizes[k], packed + lo, 1, packed + lo, 1); /* numpy: sum = 0.0; sum += cblas_ddot(...) */
        lo += sizes[k];
    }
    loss += (0.5 * c->alpha) * values / (double)n;
    /* g3_raw = a2.T @ d3 */
    double g3_raw[FM_MAX_H];
    c->dgemv(101 /* RowMajor */, 112 /* Trans */, n, h2, 1.0, c->a2, h2, c->d3, 1, 0.0, g3_raw, 1);
    *status = fm_backward_packed(c->X, c->xs_row, c->xs_col, n, n_in, h1, h2, packed, c->a1, c->a2, c->d3, c->d2, c->d1,
                                 g3_raw, sum3, c->alpha, grad);
    return loss;
}

/* Host helper of the clustering control flow (see dajin2_b200/clusterer.py). */
#include <math.h>
#include <stdint.h>

/* ---- numpy's legacy RandomState.choice(n, size, replace=False, p=uniform) in O(log n) ------------------------------
 * scikit-learn seeds every bisection with random_state.choice(n_samples, 2, replace=False, p=w / w.sum())
 * (sklearn/cluster/_kmeans.py:1030-1037) with unit weights.  numpy (random/mtrand.pyx, choice) then builds
 * cdf = cumsum(p); cdf /= cdf[-1] and returns cdf.searchsorted(x, side="right") for uniform draws x: O(n) time and
 * memory per bisection, which at 10^6 .. 10^7 reads is the largest term of the clustering stage.
 *
 * The cumulative sum is a plain sequential float64 accumulation c_k = fl(c_{k-1} + p) of the constant p = 1.0 / n.
 * While c stays inside one binade its ulp u is fixed, c is a multiple of u, and fl(c + p) adds the same multiple of u
 * every time: p rounded to a multiple of u -- to nearest, and in the tie case to the even neighbour, which after one
 * step is the same neighbour every time.  So inside a binade the sequence is an exact arithmetic progression; only
 * the few steps around a binade boundary are made with real floating-point additions.  The table of progressions
 * (about three entries per binade, ~100 in all) gives c_k for any k, the search for a draw is a binary search over k. */
#define FM_MAX_SEGMENTS 512

typedef struct {
    int64_t k;  /* index of the first term of the segment: c_k = start (k counts terms, c_1 = p) */
    int64_t len; /* c_{k+j} = start + j * step for 0 <= j <= len */
    double start, step;
} fm_segment;

static int build_segments(int64_t n, double p, fm_segment *seg) {
    int n_seg = 0;
    int64_t k = 1;
    double c = p; /* numpy's cumsum starts with the first element itself */
    while (k <= n) {
        /* two real steps: c -> c1 -> c2; if all three lie in one binade the increment c2 - c1 repeats inside it */
        if (n_seg >= FM_MAX_SEGMENTS - 2) return -1;
        seg[n_seg++] = (fm_segment){k, 0, c, 0.0};
        if (k == n) break;
        const double c1 = c + p;
        int e0, e1, e2;
        frexp(c, &e0);
        frexp(c1, &e1);
        if (k + 1 == n || e1 != e0) { /* boundary step (or the last term): keep it as a real step */
            k += 1;
            c = c1;
            continue;
        }
        const double c2 = c1 + p;
        frexp(c2, &e2);
        if (e2 != e1) {
            seg[n_seg++] = (fm_segment){k + 1, 0, c1, 0.0};
            k += 2;
            c = c2;
            continue;
        }
        const double d = c2 - c1;            /* exact: both are multiples of the binade's ulp */
        const double top = ldexp(1.0, e2);   /* the binade is [top / 2, top) */
        const double ulp = ldexp(1.0, e2 - 53);
        /* largest j with c2 + j * d < top, in units of the ulp (all quantities are integers below 2^53) */
        const int64_t room = (int64_t)((top - c2) / ulp) - 1, du = (int64_t)(d / ulp);
        int64_t j = du > 0 ? room / du : 0;
        if (j > n - (k + 2)) j = n - (k + 2);
        if (j < 0) j = 0;
        seg[n_seg++] = (fm_segment){k + 1, 0, c1, 0.0};
        seg[n_seg++] = (fm_segment){k + 2, j, c2, d};
        k = k + 2 + j;
        c = c2 + (double)j * d; /* exact */
        if (k >= n) break;
        /* the next term is made with a real addition (it may leave the binade) */
        k += 1;
        c = c + p;
    }
    return n_seg;
}

static double term(const fm_segment *seg, int n_seg, int64_t k) { /* c_k, 1 <= k <= n */
    int lo = 0, hi = n_seg - 1; /* last segment with seg.k <= k */
    while (lo < hi) {
        const int mid = (lo + hi + 1) / 2;
        if (seg[mid].k <= k) lo = mid; else hi = mid - 1;
    }
    return seg[lo].start + (double)(k - seg[lo].k) * seg[lo].step;
}

int fm_uniform_choice(int64_t n, const double *x, int n_draws, int64_t *out) {
    if (n <= 0) return 1;
    const double p = 1.0 / (double)n;
    fm_segment seg[FM_MAX_SEGMENTS];
    const int n_seg = build_segments(n, p, seg);
    if (n_seg <= 0) return 2;
    const double total = term(seg, n_seg, n);
    for (int d = 0; d < n_draws; d++) {
        /* searchsorted(x, side="right") on cdf[i] = c_{i+1} / total: the first i whose value exceeds x, else n */
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = lo + (hi - lo) / 2;
            if (term(seg, n_seg, mid + 1) / total > x[d]) hi = mid; else lo = mid + 1;
        }
        out[d] = lo;
    }
    return 0;
}

/* c_k for tests: the k-th value of np.cumsum(np.full(n, 1.0 / n)) */
double fm_uniform_cumsum_term(int64_t n, int64_t k) {
    fm_segment seg[FM_MAX_SEGMENTS];
    const int n_seg = build_segments(n, 1.0 / (double)n, seg);
    if (n_seg <= 0 || k < 1 || k > n) return -1.0;
    return term(seg, n_seg, k);
}

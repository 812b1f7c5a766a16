/* Host helper of the clustering control flow (see dajin2_b200/clusterer.py). */
#include <stdint.h>

/* ---- numpy's legacy RandomState.choice(n, size, replace=False, p=uniform) without its O(n) temporaries -----------
 * scikit-learn seeds every bisection with random_state.choice(n_samples, 2, replace=False, p=w / w.sum())
 * (sklearn/cluster/_kmeans.py:1030-1037) with unit weights.  numpy (random/mtrand.pyx, choice) then builds
 * cdf = cumsum(p); cdf /= cdf[-1] and returns cdf.searchsorted(x, side="right") for uniform draws x.  The cumulative sum
 * is a plain sequential float64 accumulation of p = 1.0 / n, so the same indices are found here with one pass that
 * keeps a checkpoint every FM_CDF_STRIDE terms and a short re-accumulation inside the block that holds each draw. */
#define FM_CDF_STRIDE 4096

int fm_uniform_choice(int64_t n, const double *x, int n_draws, int64_t *out) {
    if (n <= 0) return 1;
    const double p = 1.0 / (double)n;
    const int64_t n_ck = n / FM_CDF_STRIDE + 1;
    double stack_ck[1024];
    double *ck = n_ck <= 1024 ? stack_ck : 0;
    if (!ck) return 2; /* n > 4M: the caller uses numpy */
    double c = 0.0;
    for (int64_t i = 0; i < n; i++) {
        if (i % FM_CDF_STRIDE == 0) ck[i / FM_CDF_STRIDE] = c; /* sum of the first i terms */
        c += p;
    }
    const double total = c;
    for (int d = 0; d < n_draws; d++) {
        /* first block whose END value (normalised) exceeds x: normalised values are non-decreasing */
        int64_t lo = 0, hi = n_ck - 1; /* block b covers terms [b*S, min(n, (b+1)*S)) */
        while (lo < hi) {
            int64_t mid = (lo + hi) / 2;
            /* value after the last term of block mid = checkpoint of block mid+1 (or total) */
            double end = (mid + 1 < n_ck && (mid + 1) * FM_CDF_STRIDE <= n) ? ck[mid + 1] : total;
            if (end / total > x[d]) hi = mid; else lo = mid + 1;
        }
        double v = ck[lo];
        int64_t i = lo * FM_CDF_STRIDE, stop = i + FM_CDF_STRIDE < n ? i + FM_CDF_STRIDE : n;
        int64_t found = n; /* searchsorted returns n when no element exceeds x */
        for (; i < stop; i++) {
            v += p;
            if (v / total > x[d]) { found = i; break; }
        }
        if (found == n && stop < n) return 3; /* cannot happen: the block was chosen by its end value */
        out[d] = found;
    }
    return 0;
}

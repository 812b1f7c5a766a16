// Clustering of the read x locus score matrix: row de-duplication, bisecting 2-means, silhouette, member
// selection and label gathering.
//
// Score rows are discrete (each entry is one of a handful of per-locus score values) and heavily duplicated, so
// the production path first collapses the N rows to U distinct weighted points (dj_dedup_rows, one pass over
// the uint16 id matrix) and runs scikit-learn's arithmetic on those points in float64:
//   BisectingKMeans  sklearn/cluster/_bisect_k_means.py:298-456, Lloyd sklearn/cluster/_kmeans.py:630-758,
//                    _k_means_lloyd.pyx:221-420, _k_means_common.pyx:55-312
//   silhouette_score sklearn/metrics/cluster/_unsupervised.py:59-300
// which is what the reference calls at src/DAJIN2/core/clustering/clustering.py:35,42,68.
#include <float.h>
#include <string.h>

#include "dj_common.cuh"

namespace {

constexpr int kThreads = 256;

int grid_for(int64_t n_items, int threads) {
    int64_t blocks = (n_items + threads - 1) / threads;
    const int64_t cap = (int64_t)dj_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ---------------------------------------------------------------------------------------------------------
// row de-duplication
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
dedup_kernel(const uint16_t *__restrict__ ids, int64_t n_rows, int n_cols, const uint8_t *__restrict__ strand,
             const int32_t *__restrict__ rows, DjRowTableView tb, int32_t *__restrict__ slot_of_row,
             uint32_t *__restrict__ overflow) {
    const int lane = threadIdx.x & 31;
    const int64_t n_padded = (n_rows + 31) & ~(int64_t)31;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_padded;
         t += (int64_t)gridDim.x * blockDim.x) {
        const bool valid = t < n_rows;
        const unsigned active = __ballot_sync(0xffffffffu, valid);
        if (!valid) continue;
        const uint16_t *row = ids + t * (int64_t)n_cols;
        uint64_t h = 0x9E3779B97F4A7C15ULL ^ (uint64_t)n_cols;
        for (int c = 0; c < n_cols; ++c) h = dj_hash_combine(h, row[c]);
        const unsigned long long fp = h | 0x8000000000000000ULL;
        bool claimed;
        const uint32_t slot = dj_table_insert(tb.fp, tb.capacity, fp, &claimed);
        if (claimed) atomicAdd(tb.n_entries, 1u);
        const bool plus = strand[rows ? rows[t] : t] == '+';
        // one atomic per distinct slot in the warp instead of one per row
        const unsigned same = __match_any_sync(active, slot);
        const unsigned plus_mask = __ballot_sync(active, plus);
        if (slot == tb.capacity) {
            atomicAdd(overflow, 1u);
            slot_of_row[t] = -1;
            continue;
        }
        slot_of_row[t] = (int32_t)slot;
        if (lane == __ffs(same) - 1) {
            atomicAdd(&tb.count[slot], (uint32_t)__popc(same));
            const int n_plus = __popc(same & plus_mask);
            if (n_plus) atomicAdd(&tb.plus[slot], (uint32_t)n_plus);
            atomicMin(&tb.first[slot], (uint32_t)t);  // the leader holds the smallest t of its group
        }
        if (lane == 31 - __clz(same)) atomicMax(&tb.last[slot], (uint32_t)t);  // ... this lane the largest
    }
}

__global__ void dedup_export_kernel(DjRowTableView tb, DjRowRecord *out, uint32_t max_records, uint32_t *n_out) {
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < tb.capacity; slot += gridDim.x * blockDim.x) {
        if (tb.fp[slot] == 0ULL) continue;
        const uint32_t k = atomicAdd(n_out, 1u);
        if (k >= max_records) continue;
        DjRowRecord rec;
        rec.slot = slot;
        rec.count = tb.count[slot];
        rec.count_plus = tb.plus[slot];
        rec.first_row = tb.first[slot];
        rec.last_row = tb.last[slot];
        out[k] = rec;
    }
}

// ---------------------------------------------------------------------------------------------------------
// block-level deterministic reductions (fixed tree order)
// ---------------------------------------------------------------------------------------------------------
__device__ double block_sum(double v, double *scratch) {
    scratch[threadIdx.x] = v;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) scratch[threadIdx.x] += scratch[threadIdx.x + s];
        __syncthreads();
    }
    const double r = scratch[0];
    __syncthreads();
    return r;
}

// sklearn/cluster/_k_means_common.pyx:16-43 (_euclidean_dense_dense): groups of four, remainder one by one
__device__ double euclid_dense(const double *a, const double *b, int n, bool squared) {
    double result = 0.0;
    const int n4 = n / 4;
    for (int i = 0; i < n4; ++i) {
        const double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2], d3 = a[3] - b[3];
        result += (d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
        a += 4;
        b += 4;
    }
    for (int i = 0; i < n % 4; ++i) result += (a[i] - b[i]) * (a[i] - b[i]);
    return squared ? result : sqrt(result);
}

// sklearn/cluster/_k_means_common.pyx:55-79 (_euclidean_sparse_dense) on a dense row: zeros contribute nothing
__device__ double euclid_sparse(const double *x, const double *c, double c_sq, int n) {
    double result = 0.0;
    for (int k = 0; k < n; ++k) {
        const double a = x[k];
        if (a != 0.0) {
            const double bi = c[k];
            const double tmp = a - bi;
            result += tmp * tmp - bi * bi;
        }
    }
    result += c_sq;
    return result < 0.0 ? 0.0 : result;
}

constexpr int kBisectThreads = 512;

// One CTA runs the whole 2-means of one bisection.  The E-step is parallel over points, the M-step over columns
// (each column sums its points in point order, so the result does not depend on the thread count).
__global__ void __launch_bounds__(kBisectThreads)
bisect_kernel(const double *__restrict__ X, const double *__restrict__ weight, int n_points, int n_cols,
              int32_t *leaf, const int32_t *__restrict__ last_row, int target, int seed0, int seed1,
              double tol, int max_iter,
              int8_t *__restrict__ lab, int8_t *__restrict__ lab_old, double *__restrict__ dist,
              double *__restrict__ centers_out, double *__restrict__ stats, double *__restrict__ big_centers,
              const int32_t *__restrict__ seed_dev, int child_a, int child_b) {
    // dj_cluster_step: the seeds were located on the device (seed_dev[0..1]) and the members of the bisected leaf move
    // to the two child leaves at the end (child_a / child_b >= 0); dj_bisect passes seeds as arguments and keeps `leaf`
    if (seed_dev) {
        seed0 = seed_dev[0];
        seed1 = seed_dev[1];
    }
    extern __shared__ double smem[];
    // the four centre rows live in shared memory; with more score columns than fit there (one column per selected
    // locus: multi-kb deletions or knock-ins) they live in global memory -- one CTA, so __syncthreads orders them
    double *c_old = big_centers ? big_centers : smem;                             // [2][n_cols]
    double *c_new = (big_centers ? big_centers : smem) + 2 * n_cols;              // [2][n_cols]
    double *scratch = big_centers ? smem : smem + 4 * n_cols;                     // [kBisectThreads]
    __shared__ double cn[2], wsum[2], shift[2];
    __shared__ int s_flag, s_far;
    __shared__ int s_idx[kBisectThreads];
    __shared__ double s_max;
    const int tid = threadIdx.x;

    for (int k = tid; k < n_cols; k += blockDim.x) {
        c_old[k] = X[(int64_t)seed0 * n_cols + k];
        c_old[n_cols + k] = X[(int64_t)seed1 * n_cols + k];
    }
    for (int p = tid; p < n_points; p += blockDim.x)
        if (leaf[p] == target) lab_old[p] = -1;
    __syncthreads();

    bool strict = false;
    int iter = 0;
    for (; iter < max_iter; ++iter) {
        if (tid < 2) {
            double s = 0.0;
            for (int k = 0; k < n_cols; ++k) s += c_old[tid * n_cols + k] * c_old[tid * n_cols + k];
            cn[tid] = s;
        }
        __syncthreads();
        // E-step (_k_means_lloyd.pyx:389-409): argmin of |c|^2 - 2 x.c, ties to the first centre
        for (int p = tid; p < n_points; p += blockDim.x) {
            if (leaf[p] != target) continue;
            const double *x = X + (int64_t)p * n_cols;
            double d0 = 0.0, d1 = 0.0;
            for (int k = 0; k < n_cols; ++k) {
                const double a = x[k];
                if (a != 0.0) {
                    d0 += c_old[k] * a;
                    d1 += c_old[n_cols + k] * a;
                }
            }
            d0 = cn[0] - 2 * d0;
            d1 = cn[1] - 2 * d1;
            lab[p] = d1 < d0 ? 1 : 0;
        }
        __syncthreads();
        // M-step (_k_means_lloyd.pyx:413-418)
        for (int k = tid; k < n_cols; k += blockDim.x) {
            double s0 = 0.0, s1 = 0.0;
            for (int p = 0; p < n_points; ++p) {
                if (leaf[p] != target) continue;
                const double v = X[(int64_t)p * n_cols + k] * weight[p];
                if (lab[p]) s1 += v; else s0 += v;
            }
            c_new[k] = s0;
            c_new[n_cols + k] = s1;
        }
        if (tid == blockDim.x - 1) {
            double w0 = 0.0, w1 = 0.0;
            for (int p = 0; p < n_points; ++p) {
                if (leaf[p] != target) continue;
                if (lab[p]) w1 += weight[p]; else w0 += weight[p];
            }
            wsum[0] = w0;
            wsum[1] = w1;
        }
        __syncthreads();
        // empty-cluster relocation (_k_means_common.pyx:214-272): the farthest row (weight 1) seeds the empty cluster.
        // scikit-learn finds it with np.argpartition(distances, -1), whose kth == n - 1 case is a linear scan that
        // keeps the LAST of equal maxima (numpy/_core/src/npysort/selection.cpp, introselect_): among points at the
        // same distance the one whose last row comes last in row order (last_row[]) wins
        if (wsum[0] == 0.0 || wsum[1] == 0.0) {
            const int empty = wsum[0] == 0.0 ? 0 : 1;
            double best = -1.0;
            int best_p = -1;
            for (int p = tid; p < n_points; p += blockDim.x) {
                if (leaf[p] != target) continue;
                const int j = lab[p];
                const double d = euclid_sparse(X + (int64_t)p * n_cols, c_old + j * n_cols, cn[j], n_cols);
                dist[p] = d;
                if (d > best || (d == best && last_row[p] > last_row[best_p])) { best = d; best_p = p; }
            }
            // fixed-order arg-max: largest distance, largest last row among equals
            scratch[tid] = best;
            s_idx[tid] = best_p;
            __syncthreads();
            if (tid == 0) {
                double m = -1.0;
                int far = -1;
                for (int q = 0; q < (int)blockDim.x; ++q) {
                    if (s_idx[q] < 0) continue;
                    if (scratch[q] > m || (scratch[q] == m && last_row[s_idx[q]] > last_row[far])) {
                        m = scratch[q];
                        far = s_idx[q];
                    }
                }
                s_max = m;
                s_far = far;
            }
            __syncthreads();
            if (s_max > 0.0) {
                const int far = s_far;
                const int old = lab[far];
                for (int k = tid; k < n_cols; k += blockDim.x) {
                    const double v = X[(int64_t)far * n_cols + k];  // sample_weight of one row = 1
                    if (v != 0.0) {
                        c_new[old * n_cols + k] -= v;
                        c_new[empty * n_cols + k] = v;
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    wsum[empty] = 1.0;
                    wsum[old] -= 1.0;
                }
            }
            __syncthreads();
        }
        // _average_centers (_k_means_common.pyx:274-295): sequential over the two clusters
        const int big = wsum[1] > wsum[0] ? 1 : 0;
        for (int j = 0; j < 2; ++j) {
            if (wsum[j] > 0.0) {
                const double alpha = 1.0 / wsum[j];
                for (int k = tid; k < n_cols; k += blockDim.x) c_new[j * n_cols + k] *= alpha;
            } else {
                for (int k = tid; k < n_cols; k += blockDim.x) c_new[j * n_cols + k] = c_new[big * n_cols + k];
            }
            __syncthreads();
        }
        if (tid < 2) shift[tid] = euclid_dense(c_new + tid * n_cols, c_old + tid * n_cols, n_cols, false);
        // labels unchanged since the previous iteration?
        int differs = 0;
        for (int p = tid; p < n_points; p += blockDim.x)
            if (leaf[p] == target && lab[p] != lab_old[p]) differs = 1;
        if (tid == 0) s_flag = 0;
        __syncthreads();
        if (differs) s_flag = 1;
        __syncthreads();
        {   // centers, centers_new = centers_new, centers
            double *tmp = c_old; c_old = c_new; c_new = tmp;
        }
        if (s_flag == 0) { strict = true; ++iter; break; }
        const double shift_tot = shift[0] * shift[0] + shift[1] * shift[1];
        if (shift_tot <= tol) { ++iter; break; }
        for (int p = tid; p < n_points; p += blockDim.x)
            if (leaf[p] == target) lab_old[p] = lab[p];
        __syncthreads();
    }
    if (iter > max_iter) iter = max_iter;

    if (tid < 2) {
        double s = 0.0;
        for (int k = 0; k < n_cols; ++k) s += c_old[tid * n_cols + k] * c_old[tid * n_cols + k];
        cn[tid] = s;
    }
    __syncthreads();
    if (!strict) {  // rerun the E-step so that labels match the final centres (_kmeans.py:739-751)
        for (int p = tid; p < n_points; p += blockDim.x) {
            if (leaf[p] != target) continue;
            const double *x = X + (int64_t)p * n_cols;
            double d0 = 0.0, d1 = 0.0;
            for (int k = 0; k < n_cols; ++k) {
                const double a = x[k];
                if (a != 0.0) {
                    d0 += c_old[k] * a;
                    d1 += c_old[n_cols + k] * a;
                }
            }
            lab[p] = (cn[1] - 2 * d1) < (cn[0] - 2 * d0) ? 1 : 0;
        }
        __syncthreads();
    }
    // inertia per half (_bisect_k_means.py:266-296 -> _inertia_sparse with single_label)
    double in0 = 0.0, in1 = 0.0, w0 = 0.0, w1 = 0.0;
    for (int p = tid; p < n_points; p += blockDim.x) {
        if (leaf[p] != target) continue;
        const int j = lab[p];
        const double d = euclid_sparse(X + (int64_t)p * n_cols, c_old + j * n_cols, cn[j], n_cols) * weight[p];
        if (j) { in1 += d; w1 += weight[p]; } else { in0 += d; w0 += weight[p]; }
    }
    in0 = block_sum(in0, scratch);
    in1 = block_sum(in1, scratch);
    w0 = block_sum(w0, scratch);
    w1 = block_sum(w1, scratch);
    if (centers_out)
        for (int k = tid; k < 2 * n_cols; k += blockDim.x) centers_out[k] = c_old[k];
    if (tid == 0) {
        stats[0] = in0;
        stats[1] = in1;
        stats[2] = (double)iter;
        stats[3] = w0;
        stats[4] = w1;
    }
    if (child_a >= 0) {
        __syncthreads();
        for (int p = tid; p < n_points; p += blockDim.x)
            if (leaf[p] == target) leaf[p] = lab[p] ? child_b : child_a;
    }
}

// ---------------------------------------------------------------------------------------------------------
// silhouette on weighted points
// ---------------------------------------------------------------------------------------------------------
constexpr int kSilThreads = 128;

__global__ void __launch_bounds__(kSilThreads)
silhouette_kernel(const double *__restrict__ X, const double *__restrict__ weight, const int32_t *__restrict__ label,
                  int n_points, int n_cols, int n_clusters, int tile_rows, double *__restrict__ s_out,
                  double *__restrict__ big_acc) {
    extern __shared__ double smem[];
    // per-thread sums over the clusters: shared memory, or (many clusters) this CTA's slice of a global buffer
    double *acc = big_acc ? big_acc + (size_t)blockIdx.x * kSilThreads * n_clusters : smem;  // [kSilThreads][n_clusters]
    double *cw = big_acc ? smem : smem + (size_t)kSilThreads * n_clusters;                   // [n_clusters] cluster weights
    double *tile = cw + n_clusters;                                                          // [tile_rows][n_cols]
    const int tid = threadIdx.x;
    const int i = blockIdx.x * blockDim.x + tid;
    for (int c = tid; c < n_clusters; c += blockDim.x) cw[c] = 0.0;
    for (int c = 0; c < n_clusters; ++c) acc[(size_t)tid * n_clusters + c] = 0.0;
    __syncthreads();
    if (tid == 0)
        for (int p = 0; p < n_points; ++p) cw[label[p]] += weight[p];
    const double *xi = X + (int64_t)(i < n_points ? i : 0) * n_cols;
    for (int j0 = 0; j0 < n_points; j0 += tile_rows) {
        const int rows = min(tile_rows, n_points - j0);
        __syncthreads();
        for (int e = tid; e < rows * n_cols; e += blockDim.x) tile[e] = X[(int64_t)j0 * n_cols + e];
        __syncthreads();
        if (i < n_points) {
            for (int jj = 0; jj < rows; ++jj) {
                const double *xj = tile + jj * n_cols;
                double d = 0.0;
                for (int k = 0; k < n_cols; ++k) {
                    const double t = xi[k] - xj[k];
                    d += t * t;
                }
                acc[(size_t)tid * n_clusters + label[j0 + jj]] += weight[j0 + jj] * sqrt(d);
            }
        }
    }
    __syncthreads();
    if (i >= n_points) return;
    const int own = label[i];
    const double n_own = cw[own];
    double inter = DBL_MAX;
    for (int c = 0; c < n_clusters; ++c) {
        if (c == own || cw[c] <= 0.0) continue;
        const double m = acc[(size_t)tid * n_clusters + c] / cw[c];
        if (m < inter) inter = m;
    }
    double s = 0.0;
    if (n_own - 1.0 > 0.0 && inter != DBL_MAX) {
        const double intra = acc[(size_t)tid * n_clusters + own] / (n_own - 1.0);
        const double den = intra > inter ? intra : inter;
        if (den > 0.0) s = (inter - intra) / den;
    }
    s_out[i] = s;
}

__global__ void __launch_bounds__(kBisectThreads)
weighted_mean_kernel(const double *__restrict__ s, const double *__restrict__ weight, int n, double *out) {
    __shared__ double scratch[kBisectThreads];
    double num = 0.0, den = 0.0;
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        num += s[p] * weight[p];
        den += weight[p];
    }
    num = block_sum(num, scratch);
    den = block_sum(den, scratch);
    if (threadIdx.x == 0) out[0] = den > 0.0 ? num / den : 0.0;
}

// ---------------------------------------------------------------------------------------------------------
// member selection / label gathering / brute-force assignment
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
member_count_kernel(const int32_t *__restrict__ slot_of_row, int64_t n_rows, const int32_t *__restrict__ leaf_of_slot,
                    int target, int32_t *__restrict__ block_counts) {
    const int64_t t = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int member = (t < n_rows && leaf_of_slot[slot_of_row[t]] == target) ? 1 : 0;
    const int n = __syncthreads_count(member);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = n;
}

__global__ void __launch_bounds__(1024)
member_locate_kernel(const int32_t *__restrict__ slot_of_row, int64_t n_rows, const int32_t *__restrict__ leaf_of_slot,
                     int target, int block, int rank, int32_t *__restrict__ out_row) {
    __shared__ int warp_counts[32];
    const int64_t t = (int64_t)block * 1024 + threadIdx.x;
    const int member = (t < n_rows && leaf_of_slot[slot_of_row[t]] == target) ? 1 : 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, member);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_counts[warp] = __popc(ballot);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_counts[w];
    before += __popc(ballot & ((1u << lane) - 1u));
    if (member && before == rank) out_row[0] = (int32_t)t;
}

__global__ void gather_labels_kernel(const int32_t *__restrict__ slot_of_row, int64_t n_rows,
                                     const int32_t *__restrict__ label_of_slot, int32_t *__restrict__ labels) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_rows; t += (int64_t)gridDim.x * blockDim.x)
        labels[t] = label_of_slot[slot_of_row[t]];
}

// read x centroid block on the raw (not de-duplicated) id matrix: label + per-centroid sums
__global__ void __launch_bounds__(kThreads)
assign_rows_kernel(const uint16_t *__restrict__ ids, int64_t n_rows, int n_cols, const double *__restrict__ value_of_id,
                   const double *__restrict__ centers, int n_centers, int32_t *__restrict__ label,
                   double *__restrict__ sums, double *__restrict__ counts) {
    extern __shared__ double s_c[];  // [n_centers][n_cols] + [n_centers] squared norms
    double *s_n = s_c + (size_t)n_centers * n_cols;
    for (int e = threadIdx.x; e < n_centers * n_cols; e += blockDim.x) s_c[e] = centers[e];
    __syncthreads();
    for (int j = threadIdx.x; j < n_centers; j += blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < n_cols; ++k) s += s_c[j * n_cols + k] * s_c[j * n_cols + k];
        s_n[j] = s;
    }
    __syncthreads();
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_rows; t += (int64_t)gridDim.x * blockDim.x) {
        const uint16_t *row = ids + t * (int64_t)n_cols;
        int best = 0;
        double best_d = DBL_MAX;
        for (int j = 0; j < n_centers; ++j) {
            double dot = 0.0;
            for (int k = 0; k < n_cols; ++k) dot += s_c[j * n_cols + k] * value_of_id[row[k]];
            const double d = s_n[j] - 2 * dot;
            if (d < best_d) { best_d = d; best = j; }
        }
        label[t] = best;
        if (sums) {
            atomicAdd(&counts[best], 1.0);
            for (int k = 0; k < n_cols; ++k) {
                const double v = value_of_id[row[k]];
                if (v != 0.0) atomicAdd(&sums[(int64_t)best * n_cols + k], v);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// dj_cluster_step: one iteration of optimize_labels without a host round trip in between
// ---------------------------------------------------------------------------------------------------------
// members of leaf `target` among the sample rows, counted per 1024 rows; a row belongs to the leaf of its point
__global__ void __launch_bounds__(1024)
step_member_count_kernel(const int32_t *__restrict__ slot_of_row, int64_t n_rows, const int32_t *__restrict__ point_of_slot,
                         const int32_t *__restrict__ leaf, int target, int32_t *__restrict__ block_counts) {
    const int64_t t = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int member = (t < n_rows && leaf[point_of_slot[slot_of_row[t]]] == target) ? 1 : 0;
    const int n = __syncthreads_count(member);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = n;
}

// CTA i resolves seed i: a point given by the host (seed_given[i] >= 0), or the point of the seed_rank[i]-th member
// row of the leaf (scikit-learn draws ROW indices; rows are in file order)
__global__ void __launch_bounds__(1024)
step_seed_kernel(const int32_t *__restrict__ slot_of_row, int64_t n_rows, const int32_t *__restrict__ point_of_slot,
                 const int32_t *__restrict__ leaf, int target, const int32_t *__restrict__ block_counts, int n_blocks,
                 int given0, int given1, long long rank0, long long rank1, int32_t *__restrict__ seed_out) {
    __shared__ long long s_before[1024];
    __shared__ int s_block, s_rank;
    __shared__ int warp_counts[32];
    const int given = blockIdx.x == 0 ? given0 : given1;
    const long long rank = blockIdx.x == 0 ? rank0 : rank1;
    if (given >= 0) {
        if (threadIdx.x == 0) seed_out[blockIdx.x] = given;
        return;
    }
    // strips of the block counts: thread q owns blocks [q * per, (q + 1) * per)
    const int per = (n_blocks + 1023) / 1024;
    long long mine = 0;
    for (int b = threadIdx.x * per; b < min(n_blocks, (threadIdx.x + 1) * per); ++b) mine += block_counts[b];
    s_before[threadIdx.x] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long acc = 0;
        int q = 0;
        for (; q < 1024; ++q) {
            if (acc + s_before[q] > rank) break;
            acc += s_before[q];
        }
        int b = q * per;
        for (; b < n_blocks; ++b) {
            const int c = block_counts[b];
            if (acc + c > rank) break;
            acc += c;
        }
        s_block = b;
        s_rank = (int)(rank - acc);
    }
    __syncthreads();
    if (s_block >= n_blocks) return;  // rank beyond the members: seed_out stays -1 and the host raises
    const int64_t t = (int64_t)s_block * 1024 + threadIdx.x;
    const int pt = t < n_rows ? point_of_slot[slot_of_row[t]] : -1;
    const int member = (pt >= 0 && leaf[pt] == target) ? 1 : 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, member);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_counts[warp] = __popc(ballot);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_counts[w];
    before += __popc(ballot & ((1u << lane) - 1u));
    if (member && before == s_rank) seed_out[blockIdx.x] = pt;
}

// labels = position of each point's leaf in the left-to-right order of the leaves; clusters that hold more than 1 %
// of the control rows (clustering.py:43-45); whether the sample points carry more than one label
__global__ void __launch_bounds__(kBisectThreads)
step_relabel_kernel(const int32_t *__restrict__ leaf, const int32_t *__restrict__ order_of_leaf, int n_points,
                    int n_sample_points, const double *__restrict__ weight, int n_clusters, double control_coverage,
                    int32_t *__restrict__ label, double *__restrict__ ctrl_weight, int32_t *__restrict__ out_flags) {
    __shared__ int s_distinct, s_ctrl;
    const int tid = threadIdx.x;
    if (tid == 0) { s_distinct = 0; s_ctrl = 0; }
    for (int c = tid; c < n_clusters; c += blockDim.x) ctrl_weight[c] = 0.0;
    __syncthreads();
    const int first = n_sample_points > 0 ? order_of_leaf[leaf[0]] : -1;
    for (int p = tid; p < n_points; p += blockDim.x) {
        const int lab = order_of_leaf[leaf[p]];
        label[p] = lab;
        if (p < n_sample_points) {
            if (lab != first) s_distinct = 1;
        } else {
            atomicAdd(&ctrl_weight[lab], weight[p]);  // integer-valued weights: the sum is exact in any order
        }
    }
    __syncthreads();
    int n = 0;
    if (control_coverage > 0.0)
        for (int c = tid; c < n_clusters; c += blockDim.x)
            if (ctrl_weight[c] / control_coverage > 0.01) ++n;
    if (n) atomicAdd(&s_ctrl, n);
    __syncthreads();
    if (tid == 0) {
        out_flags[0] = s_ctrl;
        out_flags[1] = s_distinct;
    }
}

}  // namespace

extern "C" int64_t dj_row_table_bytes(uint32_t capacity) { return 64 + (int64_t)capacity * 24; }

extern "C" int dj_row_table_clear(void *table, uint32_t capacity, void *stream) {
    if (capacity == 0 || (capacity & (capacity - 1)) != 0) {
        dj_set_error("dj_row_table_clear: capacity must be a power of two");
        return 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    DjRowTableView tb = dj_row_view(table, capacity);
    DJ_CUDA_CHECK(cudaMemsetAsync(table, 0, 64 + (size_t)capacity * 16, st));
    DJ_CUDA_CHECK(cudaMemsetAsync(tb.first, 0xFF, (size_t)capacity * 4, st));
    DJ_CUDA_CHECK(cudaMemsetAsync(tb.last, 0, (size_t)capacity * 4, st));
    return 0;
}

extern "C" int dj_dedup_rows(const uint16_t *ids, int64_t n_rows, int32_t n_cols, const uint8_t *strand,
                             const int32_t *rows, void *table, uint32_t capacity, int32_t *slot_of_row,
                             void *stream) {
    if (n_rows <= 0) return 0;
    DjRowTableView tb = dj_row_view(table, capacity);
    dedup_kernel<<<grid_for(n_rows, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        ids, n_rows, n_cols, strand, rows, tb, slot_of_row, tb.n_entries + 1);
    DJ_LAUNCH_CHECK();
    return 0;
}

extern "C" int dj_dedup_export(const void *table, uint32_t capacity, DjRowRecord *h_out, uint32_t max_records,
                               uint32_t *h_n_records, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    DjRowTableView tb = dj_row_view(const_cast<void *>(table), capacity);
    uint32_t header[2] = {0, 0};
    DJ_CUDA_CHECK(cudaMemcpyAsync(header, tb.n_entries, sizeof(header), cudaMemcpyDeviceToHost, st));
    DJ_CUDA_CHECK(cudaStreamSynchronize(st));
    if (header[1] != 0) {
        dj_set_error("dj_dedup_export: row table overflow (%u rows lost); use a larger capacity", header[1]);
        return 2;
    }
    *h_n_records = header[0];
    if (header[0] == 0) return 0;
    if (header[0] > max_records) {
        dj_set_error("dj_dedup_export: %u records do not fit in %u", header[0], max_records);
        return 3;
    }
    DjRowRecord *d_out = nullptr;
    uint32_t *d_n = nullptr;
    if (dj_scratch_pool_init() != 0) return 1;
    // stream-ordered scratch: cudaMalloc / cudaFree synchronise the device and cost tens of milliseconds in a
    // process that holds gigabytes of allocations
    DJ_CUDA_CHECK(cudaMallocAsync(&d_out, sizeof(DjRowRecord) * (size_t)header[0] + 16, st));
    d_n = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(d_out) + sizeof(DjRowRecord) * (size_t)header[0]);
    DJ_CUDA_CHECK(cudaMemsetAsync(d_n, 0, sizeof(uint32_t), st));
    dedup_export_kernel<<<grid_for(capacity, kThreads), kThreads, 0, st>>>(tb, d_out, header[0], d_n);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess)
        err = cudaMemcpyAsync(h_out, d_out, sizeof(DjRowRecord) * (size_t)header[0], cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    cudaFreeAsync(d_out, st);
    if (err != cudaSuccess) {
        dj_set_error("dj_dedup_export: %s", cudaGetErrorString(err));
        return 1;
    }
    return 0;
}

extern "C" int dj_bisect(const double *X, const double *weight, int32_t n_points, int32_t n_cols,
                         const int32_t *leaf, const int32_t *last_row, int32_t target, int32_t seed0, int32_t seed1,
                         double tol, int32_t max_iter, int8_t *sub_label, double *centers, double *stats, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_points <= 0 || n_cols <= 0) {
        dj_set_error("dj_bisect: empty problem");
        return 1;
    }
    size_t smem = sizeof(double) * ((size_t)4 * n_cols + kBisectThreads);
    const bool big = smem > 200 * 1024;  // centres in global memory instead
    if (big) smem = sizeof(double) * kBisectThreads;
    DJ_CUDA_CHECK(cudaFuncSetAttribute(bisect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int8_t *lab_old = nullptr;
    double *dist = nullptr, *big_centers = nullptr;
    DJ_CUDA_CHECK(cudaMallocAsync(&lab_old, (size_t)n_points, st));
    DJ_CUDA_CHECK(cudaMallocAsync(&dist, sizeof(double) * (size_t)n_points, st));
    if (big) DJ_CUDA_CHECK(cudaMallocAsync(&big_centers, sizeof(double) * 4 * (size_t)n_cols, st));
    bisect_kernel<<<1, kBisectThreads, smem, st>>>(X, weight, n_points, n_cols, const_cast<int32_t *>(leaf), last_row,
                                                   target, seed0, seed1, tol, max_iter, sub_label, lab_old, dist, centers,
                                                   stats, big_centers, nullptr, -1, -1);
    cudaError_t err = cudaGetLastError();
    cudaFreeAsync(lab_old, st);
    cudaFreeAsync(dist, st);
    if (big_centers) cudaFreeAsync(big_centers, st);
    if (err != cudaSuccess) {
        dj_set_error("dj_bisect: %s", cudaGetErrorString(err));
        return 1;
    }
    return 0;
}

extern "C" int dj_silhouette(const double *X, const double *weight, const int32_t *label, int32_t n_points,
                             int32_t n_cols, int32_t n_clusters, double *score, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_points <= 0 || n_clusters <= 0) {
        dj_set_error("dj_silhouette: empty problem");
        return 1;
    }
    int tile_rows = 16384 / (n_cols > 0 ? n_cols : 1);
    if (tile_rows < 1) tile_rows = 1;
    if (tile_rows > 64) tile_rows = 64;
    size_t smem = sizeof(double) * ((size_t)kSilThreads * n_clusters + n_clusters + (size_t)tile_rows * n_cols);
    const bool big = smem > 200 * 1024;  // the per-thread cluster sums move to global memory
    if (big) {
        while (tile_rows > 1 && sizeof(double) * ((size_t)n_clusters + (size_t)tile_rows * n_cols) > 200 * 1024) tile_rows /= 2;
        smem = sizeof(double) * ((size_t)n_clusters + (size_t)tile_rows * n_cols);
        if (smem > 200 * 1024) {
            dj_set_error("dj_silhouette: %d clusters x %d columns exceed the shared-memory budget", n_clusters, n_cols);
            return 1;
        }
    }
    DJ_CUDA_CHECK(cudaFuncSetAttribute(silhouette_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (unsigned)((n_points + kSilThreads - 1) / kSilThreads);
    double *s = nullptr, *big_acc = nullptr;
    DJ_CUDA_CHECK(cudaMallocAsync(&s, sizeof(double) * (size_t)n_points, st));
    if (big) DJ_CUDA_CHECK(cudaMallocAsync(&big_acc, sizeof(double) * (size_t)blocks * kSilThreads * n_clusters, st));
    silhouette_kernel<<<blocks, kSilThreads, smem, st>>>(X, weight, label, n_points, n_cols, n_clusters, tile_rows, s,
                                                         big_acc);
    cudaError_t err = cudaGetLastError();
    if (big_acc) cudaFreeAsync(big_acc, st);
    if (err == cudaSuccess) {
        weighted_mean_kernel<<<1, kBisectThreads, 0, st>>>(s, weight, n_points, score);
        err = cudaGetLastError();
    }
    cudaFreeAsync(s, st);
    if (err != cudaSuccess) {
        dj_set_error("dj_silhouette: %s", cudaGetErrorString(err));
        return 1;
    }
    return 0;
}

extern "C" int dj_member_counts(const int32_t *slot_of_row, int64_t n_rows, const int32_t *leaf_of_slot,
                                int32_t target, int32_t *block_counts, void *stream) {
    if (n_rows <= 0) return 0;
    const int64_t blocks = (n_rows + 1023) / 1024;
    member_count_kernel<<<(unsigned)blocks, 1024, 0, (cudaStream_t)stream>>>(slot_of_row, n_rows, leaf_of_slot, target,
                                                                            block_counts);
    DJ_LAUNCH_CHECK();
    return 0;
}

extern "C" int dj_member_locate(const int32_t *slot_of_row, int64_t n_rows, const int32_t *leaf_of_slot,
                                int32_t target, int32_t block, int32_t rank_in_block, int32_t *out_row,
                                void *stream) {
    member_locate_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(slot_of_row, n_rows, leaf_of_slot, target, block,
                                                              rank_in_block, out_row);
    DJ_LAUNCH_CHECK();
    return 0;
}

extern "C" int dj_gather_labels(const int32_t *slot_of_row, int64_t n_rows, const int32_t *label_of_slot,
                                int32_t *labels, void *stream) {
    if (n_rows <= 0) return 0;
    gather_labels_kernel<<<grid_for(n_rows, kThreads), kThreads, 0, (cudaStream_t)stream>>>(slot_of_row, n_rows,
                                                                                          label_of_slot, labels);
    DJ_LAUNCH_CHECK();
    return 0;
}

extern "C" int dj_assign_rows(const uint16_t *ids, int64_t n_rows, int32_t n_cols, const double *value_of_id,
                              const double *centers, int32_t n_centers, int32_t *label, double *sums,
                              double *counts, void *stream) {
    if (n_rows <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(double) * ((size_t)n_centers * n_cols + n_centers);
    if (smem > 200 * 1024) {
        dj_set_error("dj_assign_rows: centroid block does not fit in shared memory");
        return 1;
    }
    DJ_CUDA_CHECK(cudaFuncSetAttribute(assign_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (sums) {
        DJ_CUDA_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)n_centers * n_cols, st));
        DJ_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(double) * (size_t)n_centers, st));
    }
    assign_rows_kernel<<<grid_for(n_rows, kThreads), kThreads, smem, st>>>(ids, n_rows, n_cols, value_of_id, centers,
                                                                         n_centers, label, sums, counts);
    DJ_LAUNCH_CHECK();
    return 0;
}

// One iteration of the reference's optimize_labels loop (clustering/clustering.py:29-55) as one stream of launches
// and ONE device -> host copy: the two random seeds of the bisection located among the sample rows (or given as
// points), the 2-means of the leaf (dj_bisect's kernel), the left-to-right cluster labels, the control-share rule and
// the silhouette score.  `leaf` (device, int32 per point) is updated in place: members of `target` move to child_a /
// child_b.  Host outputs: sub_host[p] (0 / 1 for the members of the leaf as they were split, untouched elsewhere) and
// out (DjClusterStepOut).
extern "C" int dj_cluster_step(const double *X, const double *weight, int32_t n_points, int32_t n_sample_points,
                               int32_t n_cols, int32_t *leaf, const int32_t *last_row, const int32_t *slot_of_row,
                               int64_t n_rows, const int32_t *point_of_slot, int32_t target, int32_t child_a,
                               int32_t child_b, const int32_t *seed_given, const int64_t *seed_rank,
                               const int32_t *order_of_leaf_host, int32_t n_leaf_ids, int32_t n_clusters,
                               double control_coverage, double tol, int32_t max_iter, int32_t want_silhouette,
                               int8_t *sub_host, DjClusterStepOut *out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_points <= 0 || n_cols <= 0 || n_clusters <= 0 || n_leaf_ids <= 0) {
        dj_set_error("dj_cluster_step: empty problem");
        return 1;
    }
    const bool locate = seed_given[0] < 0 || seed_given[1] < 0;
    if (locate && (slot_of_row == nullptr || point_of_slot == nullptr || n_rows <= 0)) {
        dj_set_error("dj_cluster_step: a seed is given as a row rank but there are no rows");
        return 1;
    }
    // pinned staging for the step's inputs and results (grown on demand, kept for the life of the process)
    static uint8_t *h_pin = nullptr;
    static size_t h_pin_bytes = 0;
    const size_t order_bytes = sizeof(int32_t) * (size_t)n_leaf_ids;
    const size_t res_off = (order_bytes + 15) & ~(size_t)15;
    const size_t res_bytes = 64 + (((size_t)n_points + 15) & ~(size_t)15);  // stats[5], score, flags[2], seeds[2]; sub labels
    const size_t need = res_off + res_bytes;
    if (need > h_pin_bytes) {
        if (h_pin) cudaFreeHost(h_pin);
        h_pin = nullptr;
        h_pin_bytes = 0;
        DJ_CUDA_CHECK(cudaMallocHost(&h_pin, need * 2));
        h_pin_bytes = need * 2;
    }
    memcpy(h_pin, order_of_leaf_host, order_bytes);

    const int64_t n_blocks = locate ? (n_rows + 1023) / 1024 : 0;
    // one device allocation for everything the step needs
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_order = take(order_bytes), o_res = take(res_bytes), o_labold = take((size_t)n_points),
                 o_dist = take(sizeof(double) * (size_t)n_points), o_label = take(sizeof(int32_t) * (size_t)n_points),
                 o_ctrl = take(sizeof(double) * (size_t)n_clusters), o_counts = take(sizeof(int32_t) * (size_t)(n_blocks + 1)),
                 o_sil = take(sizeof(double) * (size_t)n_points);
    size_t bis_smem = sizeof(double) * ((size_t)4 * n_cols + kBisectThreads);
    const bool bis_big = bis_smem > 200 * 1024;
    if (bis_big) bis_smem = sizeof(double) * kBisectThreads;
    const size_t o_bigc = take(bis_big ? sizeof(double) * 4 * (size_t)n_cols : 0);
    int tile_rows = 16384 / n_cols;
    if (tile_rows < 1) tile_rows = 1;
    if (tile_rows > 64) tile_rows = 64;
    size_t sil_smem = sizeof(double) * ((size_t)kSilThreads * n_clusters + n_clusters + (size_t)tile_rows * n_cols);
    const bool sil_big = sil_smem > 200 * 1024;
    const unsigned sil_blocks = (unsigned)((n_points + kSilThreads - 1) / kSilThreads);
    if (sil_big) {
        while (tile_rows > 1 && sizeof(double) * ((size_t)n_clusters + (size_t)tile_rows * n_cols) > 200 * 1024) tile_rows /= 2;
        sil_smem = sizeof(double) * ((size_t)n_clusters + (size_t)tile_rows * n_cols);
        if (sil_smem > 200 * 1024) {
            dj_set_error("dj_cluster_step: %d clusters x %d columns exceed the shared-memory budget", n_clusters, n_cols);
            return 1;
        }
    }
    const size_t o_bigacc = take(sil_big ? sizeof(double) * (size_t)sil_blocks * kSilThreads * n_clusters : 0);
    uint8_t *ws = nullptr;
    DJ_CUDA_CHECK(cudaMallocAsync(&ws, off + 256, st));
    int32_t *d_order = reinterpret_cast<int32_t *>(ws + o_order);
    uint8_t *d_res = ws + o_res;
    double *d_stats = reinterpret_cast<double *>(d_res);           // [0..4] bisect, [5] silhouette
    int32_t *d_flags = reinterpret_cast<int32_t *>(d_res + 48);    // ctrl clusters, distinct sample labels
    int32_t *d_seeds = reinterpret_cast<int32_t *>(d_res + 56);    // the two seed points
    int8_t *d_sub = reinterpret_cast<int8_t *>(d_res + 64);
    int32_t *d_label = reinterpret_cast<int32_t *>(ws + o_label);

    cudaError_t err = cudaMemcpyAsync(d_order, h_pin, order_bytes, cudaMemcpyHostToDevice, st);
    if (err == cudaSuccess) err = cudaMemsetAsync(d_res, 0xFF, res_bytes, st);  // seeds = -1, sub labels = 0xFF
    if (err == cudaSuccess && locate) {
        step_member_count_kernel<<<(unsigned)n_blocks, 1024, 0, st>>>(slot_of_row, n_rows, point_of_slot, leaf, target,
                                                                      reinterpret_cast<int32_t *>(ws + o_counts));
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) {
        step_seed_kernel<<<2, 1024, 0, st>>>(slot_of_row, n_rows, point_of_slot, leaf, target,
                                             reinterpret_cast<int32_t *>(ws + o_counts), (int)n_blocks, seed_given[0],
                                             seed_given[1], (long long)seed_rank[0], (long long)seed_rank[1], d_seeds);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaFuncSetAttribute(bisect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bis_smem);
    if (err == cudaSuccess) {
        bisect_kernel<<<1, kBisectThreads, bis_smem, st>>>(
            X, weight, n_points, n_cols, leaf, last_row, target, 0, 0, tol, max_iter, d_sub,
            reinterpret_cast<int8_t *>(ws + o_labold), reinterpret_cast<double *>(ws + o_dist), nullptr, d_stats,
            bis_big ? reinterpret_cast<double *>(ws + o_bigc) : nullptr, d_seeds, child_a, child_b);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) {
        step_relabel_kernel<<<1, kBisectThreads, 0, st>>>(leaf, d_order, n_points, n_sample_points, weight, n_clusters,
                                                          control_coverage, d_label,
                                                          reinterpret_cast<double *>(ws + o_ctrl), d_flags);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess && want_silhouette) {
        err = cudaFuncSetAttribute(silhouette_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sil_smem);
        if (err == cudaSuccess) {
            silhouette_kernel<<<sil_blocks, kSilThreads, sil_smem, st>>>(
                X, weight, d_label, n_points, n_cols, n_clusters, tile_rows, reinterpret_cast<double *>(ws + o_sil),
                sil_big ? reinterpret_cast<double *>(ws + o_bigacc) : nullptr);
            err = cudaGetLastError();
        }
        if (err == cudaSuccess) {
            weighted_mean_kernel<<<1, kBisectThreads, 0, st>>>(reinterpret_cast<double *>(ws + o_sil), weight, n_points,
                                                               d_stats + 5);
            err = cudaGetLastError();
        }
    }
    if (err == cudaSuccess) err = cudaMemcpyAsync(h_pin + res_off, d_res, res_bytes, cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(ws, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    if (err != cudaSuccess) {
        dj_set_error("dj_cluster_step: %s", cudaGetErrorString(err));
        return 1;
    }
    const double *h_stats = reinterpret_cast<const double *>(h_pin + res_off);
    const int32_t *h_flags = reinterpret_cast<const int32_t *>(h_pin + res_off + 48);
    out->inertia[0] = h_stats[0];
    out->inertia[1] = h_stats[1];
    out->weight[0] = h_stats[3];
    out->weight[1] = h_stats[4];
    out->silhouette = h_stats[5];
    out->iterations = (int32_t)h_stats[2];
    out->control_clusters = h_flags[0];
    out->sample_labels_distinct = h_flags[1];
    out->seed_point[0] = h_flags[2];
    out->seed_point[1] = h_flags[3];
    memcpy(sub_host, h_pin + res_off + 64, (size_t)n_points);
    if (out->seed_point[0] < 0 || out->seed_point[1] < 0) {
        dj_set_error("dj_cluster_step: a seed rank lies beyond the members of leaf %d", target);
        return 1;
    }
    return 0;
}

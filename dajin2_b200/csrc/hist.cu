// Per-locus histograms of the code matrix (dj_hist) and the 10-wide window cosine distances (dj_window_cosine).
//
// dj_hist: each thread owns four consecutive loci (one 32-bit load per read), a CTA owns 1024 consecutive loci
// and walks down the reads with a fixed stride, so every request is a fully coalesced 1 KB segment of one row.
// The ten counters per locus (4 classes with the reference's skip rule + 3 startswith classes x 2 strands) are
// kept as packed 8-bit fields in registers -- one table look-up and two adds per token -- and spilled to the
// 32-bit global counters every 255 reads.  Algorithmic bytes per read: code_pitch + 1.
//
// Reference: count_indels (src/DAJIN2/core/preprocess/mutation_processing/indel_counter.py:15-31) and
// count_strandless_of_indels (src/DAJIN2/core/preprocess/error_correction/strand_bias_handler.py:18-33).
#include "dj_common.cuh"

namespace {

constexpr int kHistThreads = 256;
constexpr int kLociPerCta = kHistThreads * 4;

__global__ void __launch_bounds__(kHistThreads)
hist_kernel(const uint8_t *__restrict__ codes, int code_pitch, const uint8_t *__restrict__ strand,
            const int32_t *__restrict__ rows, int64_t n_rows, int n_loci, int32_t *__restrict__ counts) {
    // per code byte: low word = '=','+','-','*' increments (skip rule applied), high word = startswith '+','-','*'
    __shared__ uint2 lut[256];
    {
        const int c = threadIdx.x;
        const int hc = dj_code_hist_class(c);
        const int sc = dj_code_startswith_class(c);
        lut[c] = make_uint2(hc < 4 ? 1u << (8 * hc) : 0u, sc < 3 ? 1u << (8 * sc) : 0u);
    }
    __syncthreads();

    const int locus0 = blockIdx.x * kLociPerCta + threadIdx.x * 4;
    if (locus0 >= code_pitch) return;
    uint32_t acc_h[4] = {0, 0, 0, 0}, acc_p[4] = {0, 0, 0, 0}, acc_m[4] = {0, 0, 0, 0};
    int pending = 0;

    auto flush = [&]() {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int locus = locus0 + j;
            if (locus < n_loci) {
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    const uint32_t v = (acc_h[j] >> (8 * f)) & 0xFFu;
                    if (v) atomicAdd(&counts[(DJ_HIST_MATCH + f) * (int64_t)n_loci + locus], (int32_t)v);
                }
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    const uint32_t vp = (acc_p[j] >> (8 * f)) & 0xFFu;
                    const uint32_t vm = (acc_m[j] >> (8 * f)) & 0xFFu;
                    if (vp) atomicAdd(&counts[(DJ_HIST_SW_INS_P + f) * (int64_t)n_loci + locus], (int32_t)vp);
                    if (vm) atomicAdd(&counts[(DJ_HIST_SW_INS_M + f) * (int64_t)n_loci + locus], (int32_t)vm);
                }
            }
            acc_h[j] = acc_p[j] = acc_m[j] = 0;
        }
        pending = 0;
    };

    for (int64_t t = blockIdx.y; t < n_rows; t += gridDim.y) {
        const int64_t r = rows ? rows[t] : t;
        const uint32_t word = *reinterpret_cast<const uint32_t *>(codes + r * (int64_t)code_pitch + locus0);
        const bool plus = strand[r] == '+';
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint2 inc = lut[(word >> (8 * j)) & 0xFFu];
            acc_h[j] += inc.x;
            if (plus) acc_p[j] += inc.y; else acc_m[j] += inc.y;
        }
        if (++pending == 255) flush();
    }
    flush();
}

// cosine_distance of mutation_processing/anomaly_detector.py:11-25 on a window that starts at `i`
__device__ double window_distance(const double *x, const double *y, int begin, int end) {
    if (end <= begin) return 0.0;
    double dot = 0.0, nx = 0.0, ny = 0.0;
    for (int k = begin; k < end; ++k) {
        dot += x[k] * y[k];
        nx += x[k] * x[k];
        ny += y[k] * y[k];
    }
    const double den = sqrt(nx) * sqrt(ny);
    if (den == 0.0 || isnan(den)) return 1.0;
    double sim = dot / (den + 2.220446049250313e-16);
    sim = fmin(1.0, fmax(-1.0, sim));
    return 1.0 - sim;
}

__global__ void window_cosine_kernel(const double *__restrict__ sample, const double *__restrict__ control,
                                     int n_loci, double *__restrict__ dist, double *__restrict__ dist_tail) {
    // stage the tile (+9 halo) through shared memory: every value is used by ten windows
    extern __shared__ double tile[];
    double *xs = tile, *ys = tile + blockDim.x + 9;
    const int i0 = blockIdx.x * blockDim.x;
    for (int k = threadIdx.x; k < (int)blockDim.x + 9; k += blockDim.x) {
        const int g = i0 + k;
        xs[k] = g < n_loci ? sample[g] : 0.0;
        ys[k] = g < n_loci ? control[g] : 0.0;
    }
    __syncthreads();
    const int i = i0 + threadIdx.x;
    if (i >= n_loci) return;
    const int width = min(10, n_loci - i);  // python slicing truncates at the end of the array
    dist[i] = window_distance(xs, ys, threadIdx.x, threadIdx.x + width);
    dist_tail[i] = window_distance(xs, ys, threadIdx.x + 1, threadIdx.x + width);
}

}  // namespace

extern "C" int dj_hist(const uint8_t *codes, int32_t code_pitch, const uint8_t *strand, const int32_t *rows,
                       int64_t n_rows, int32_t n_loci, int32_t *counts, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    DJ_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * DJ_HIST_ROWS * (size_t)n_loci, st));
    if (n_rows <= 0 || n_loci <= 0) return 0;
    if ((code_pitch & 3) != 0 || ((uintptr_t)codes & 3) != 0) {
        dj_set_error("dj_hist: code matrix must be 4-byte aligned with a pitch that is a multiple of 4");
        return 1;
    }
    const int sms = dj_sm_count();
    if (sms <= 0) return 1;
    const int gx = (code_pitch + kLociPerCta - 1) / kLociPerCta;
    int64_t gy = ((int64_t)sms * 8 + gx - 1) / gx;  // ~8 resident CTAs per SM
    if (gy > n_rows) gy = n_rows;
    if (gy > 65535) gy = 65535;
    hist_kernel<<<dim3(gx, (unsigned)gy), kHistThreads, 0, st>>>(codes, code_pitch, strand, rows, n_rows, n_loci, counts);
    DJ_LAUNCH_CHECK();
    return 0;
}

extern "C" int dj_window_cosine(const double *sample, const double *control, int32_t n_loci, double *dist,
                                double *dist_tail, void *stream) {
    if (n_loci <= 0) return 0;
    const int threads = 128;
    const int blocks = (n_loci + threads - 1) / threads;
    window_cosine_kernel<<<blocks, threads, sizeof(double) * 2 * (threads + 9), (cudaStream_t)stream>>>(
        sample, control, n_loci, dist, dist_tail);
    DJ_LAUNCH_CHECK();
    return 0;
}

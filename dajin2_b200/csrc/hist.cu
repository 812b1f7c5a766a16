// Per-locus histograms of the code matrix (dj_hist) and the 10-wide window cosine distances (dj_window_cosine).
//
// dj_hist.  98 tokens in 100 are plain matches "=A|C|G|T" (code bytes 0..3) and need no counter of their own: their
// number at a locus is the number of reads minus everything else seen there.  So the pass over the matrix only
// LOOKS for the other tokens -- three byte-parallel operations per four loci -- and counts those:
//   * a CTA of 8 warps owns 512 consecutive loci (16 loci = one 16-byte load per thread); its warps take blocks of
//     32 reads from a counter shared by the CTAs of the same tile column; every request of a warp is a fully coalesced 512-byte segment of one row;
//   * a thread whose 16 loci hold an other-than-plain token appends its 16 bytes to its warp's queue in shared
//     memory (one ballot per row); whenever 32 groups are queued the warp drains them with all lanes busy: one
//     shared-memory atomic (two for lowercase / N-anchored tokens) per queued token on the CTA's 10 x 512 tile;
//   * the tile goes to a scratch buffer as one partial per slab (no global atomics), and a second small kernel sums
//     the partials and derives the ten output rows.
// Algorithmic bytes per read: code_pitch + 1; the partials add 20 KB per CTA.
//
// Reference: count_indels (src/DAJIN2/core/preprocess/mutation_processing/indel_counter.py:15-31) and
// count_strandless_of_indels (src/DAJIN2/core/preprocess/error_correction/strand_bias_handler.py:18-33).
#include "dj_common.cuh"

namespace {

constexpr int kHistThreads = 256;
constexpr int kHistWarps = kHistThreads / 32;  // one warp = 512 loci of one row; the 8 warps take rows t, t+1, ... t+7
constexpr int kTileLoci = 512;
constexpr int kRowUnroll = 4;
constexpr int kRowBlock = 32;  // reads a warp takes at a time
constexpr int kQueue = 64;  // 16-byte groups per warp queue (a drain takes 32)
// tile rows: what is counted in shared memory
constexpr int kTileSwPlus = 0;   // 0..2  startswith '+','-','*' on '+' strand reads (strand_bias_handler.py:25-28)
constexpr int kTileSwMinus = 3;  // 3..5  the same on '-' strand reads
constexpr int kTileSkipped = 6;  // 6..8  of those, tokens count_indels skips (lowercase or "=N" anchor, :20-23)
constexpr int kTileOther = 9;    // tokens that are neither plain matches nor start with '+','-','*' ("=N", "=a", invalid)
constexpr int kTileRows = 10;

struct HistSmem {
    int32_t tile[kTileRows][kTileLoci];
    uint4 qdata[kHistWarps][kQueue];
    uint2 qmeta[kHistWarps][kQueue];  // x = first locus of the group inside the tile | strand '-' << 31, y = byte flags
    // per code byte and strand ('+' first): low byte = 1 + tile row of the token's counter, high byte = 1 + tile row
    // of its second counter (0 = none)
    uint16_t class_lut[2][256];
};

// 0x80 in every byte of w that is not a plain match (code byte >= 4)
__device__ __forceinline__ uint32_t other_flags(uint32_t w) {
    return (((w | 0x80808080u) - 0x04040404u) | w) & 0x80808080u;
}

template <bool kGather>
__global__ void __launch_bounds__(kHistThreads, 4)
hist_partial_kernel(const uint8_t *__restrict__ codes, int code_pitch, const uint8_t *__restrict__ strand,
                    const int32_t *__restrict__ rows, int64_t n_rows, int32_t *__restrict__ partial, int partial_pitch,
                    int32_t *__restrict__ next_row) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    HistSmem &sm = *reinterpret_cast<HistSmem *>(smem_raw);
    for (int e = threadIdx.x; e < kTileRows * kTileLoci; e += kHistThreads) (&sm.tile[0][0])[e] = 0;
    for (int c = threadIdx.x; c < 512; c += kHistThreads) {
        const int code = c & 255, minus = c >> 8;
        const int hc = dj_code_hist_class(code), sc = dj_code_startswith_class(code);
        uint32_t ent = 0;
        if (code >= 4)
            ent = sc < 3 ? (uint32_t)((minus ? kTileSwMinus : kTileSwPlus) + sc + 1) |
                               (hc == 4 ? (uint32_t)(kTileSkipped + sc + 1) << 8 : 0u)
                         : (uint32_t)(kTileOther + 1);
        sm.class_lut[minus][code] = (uint16_t)ent;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int local0 = lane * 16;  // first of this thread's 16 loci inside the tile
    const int locus0 = blockIdx.x * kTileLoci + local0;
    const bool active = locus0 < code_pitch;  // code_pitch is a multiple of 16: the 16-byte load stays inside the row
    uint4 *qdata = sm.qdata[warp];
    uint2 *qmeta = sm.qmeta[warp];
    const uint8_t *qbytes = reinterpret_cast<const uint8_t *>(qdata);
    const uint32_t lanes_below = (1u << lane) - 1u;
    uint32_t q_in = 0, q_out = 0;  // groups appended / drained so far (warp-uniform)

    // Count the flagged tokens of (up to) 32 queued groups, one group per lane.
    auto drain = [&](uint32_t n) {
        __syncwarp();
        if ((uint32_t)lane < n) {
            const uint32_t slot = (q_out + lane) & (kQueue - 1);
            const uint2 meta = qmeta[slot];
            const uint16_t *lut = sm.class_lut[meta.x >> 31];
            int32_t *cell = &sm.tile[0][meta.x & 0x7FFFFFFFu];
            const uint8_t *bytes = qbytes + slot * 16;
            uint32_t m = meta.y;  // bit 8*j + 4 + k: byte j of word k, i.e. locus first + 4*k + j
            while (m) {
                const int bit = 31 - __clz(m);
                m ^= 1u << bit;
                const int pos = ((bit & 7) - 4) * 4 + (bit >> 3);
                const uint32_t ent = lut[bytes[pos]];
                atomicAdd(cell + ((int)(ent & 0xFFu) - 1) * kTileLoci + pos, 1);
                if (ent >> 8) atomicAdd(cell + ((int)(ent >> 8) - 1) * kTileLoci + pos, 1);
            }
        }
        q_out += n;
        __syncwarp();
    };

    // Work distribution: every warp of the CTAs that share this tile column takes the next kRowBlock reads of the
    // matrix from the column's counter, so all SMs finish together whatever their speed.
    for (;;) {
        int64_t t = 0;
        if (lane == 0) t = (int64_t)atomicAdd(next_row + blockIdx.x, kRowBlock);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_rows) break;
        const int64_t t_end = min(n_rows, t + kRowBlock);
        const uint8_t *row_ptr = codes + t * (int64_t)code_pitch + locus0;  // !kGather: walks down the matrix
        for (; t < t_end; t += kRowUnroll) {
            uint4 v[kRowUnroll];
            uint32_t minus[kRowUnroll];
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u) {
                v[u] = make_uint4(0, 0, 0, 0);  // reads past the block and loci past the matrix read as plain matches
                minus[u] = 0;
                if (t + u < t_end && active) {
                    if (kGather) {
                        const int64_t r = rows[t + u];
                        v[u] = __ldg(reinterpret_cast<const uint4 *>(codes + r * (int64_t)code_pitch + locus0));
                        minus[u] = __ldg(strand + r) == '+' ? 0u : 0x80000000u;
                    } else {
                        v[u] = __ldg(reinterpret_cast<const uint4 *>(row_ptr + u * (int64_t)code_pitch));
                        minus[u] = __ldg(strand + t + u) == '+' ? 0u : 0x80000000u;
                    }
                }
            }
            row_ptr += kRowUnroll * (int64_t)code_pitch;
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u) {
                const uint32_t m = (other_flags(v[u].x) >> 3) | (other_flags(v[u].y) >> 2) |
                                   (other_flags(v[u].z) >> 1) | other_flags(v[u].w);
                const uint32_t vote = __ballot_sync(0xffffffffu, m != 0);
                if (m != 0) {
                    const uint32_t slot = (q_in + __popc(vote & lanes_below)) & (kQueue - 1);
                    qdata[slot] = v[u];
                    qmeta[slot] = make_uint2((uint32_t)local0 | minus[u], m);
                }
                q_in += __popc(vote);
                if (q_in - q_out >= 32) drain(32);
            }
        }
    }
    drain(q_in - q_out);
    __syncthreads();
    int32_t *out = partial + (int64_t)blockIdx.y * kTileRows * partial_pitch + blockIdx.x * kTileLoci;
    for (int e = threadIdx.x; e < kTileRows * kTileLoci; e += kHistThreads) {
        const int row = e / kTileLoci, col = e % kTileLoci;
        if (blockIdx.x * kTileLoci + col < partial_pitch) out[(int64_t)row * partial_pitch + col] = sm.tile[row][col];
    }
}

// Sum the slab partials and derive the ten output rows (see include/dajin2_b200.h for their meaning).
__global__ void hist_reduce_kernel(const int32_t *__restrict__ partial, int partial_pitch, int n_slabs, int64_t n_rows,
                                   int n_loci, int32_t *__restrict__ counts) {
    const int locus = blockIdx.x * blockDim.x + threadIdx.x;
    if (locus >= n_loci) return;
    int32_t sum[kTileRows];
#pragma unroll
    for (int k = 0; k < kTileRows; ++k) sum[k] = 0;
    for (int s = 0; s < n_slabs; ++s) {
        const int32_t *p = partial + (int64_t)s * kTileRows * partial_pitch + locus;
#pragma unroll
        for (int k = 0; k < kTileRows; ++k) sum[k] += __ldg(p + (int64_t)k * partial_pitch);
    }
    int32_t starts = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int32_t both = sum[kTileSwPlus + c] + sum[kTileSwMinus + c];
        starts += both;
        counts[(DJ_HIST_INS + c) * (int64_t)n_loci + locus] = both - sum[kTileSkipped + c];
        counts[(DJ_HIST_SW_INS_P + c) * (int64_t)n_loci + locus] = sum[kTileSwPlus + c];
        counts[(DJ_HIST_SW_INS_M + c) * (int64_t)n_loci + locus] = sum[kTileSwMinus + c];
    }
    counts[DJ_HIST_MATCH * (int64_t)n_loci + locus] = (int32_t)n_rows - starts - sum[kTileOther];
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
    if (n_loci <= 0) return 0;
    if (n_rows <= 0) {
        DJ_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * DJ_HIST_ROWS * (size_t)n_loci, st));
        return 0;
    }
    if ((code_pitch & 15) != 0 || ((uintptr_t)codes & 15) != 0 || n_rows > 0x7FFFFFFF) {
        dj_set_error("dj_hist: code matrix must be 16-byte aligned with a pitch that is a multiple of 16 (< 2^31 reads)");
        return 1;
    }
    const int sms = dj_sm_count();
    if (sms <= 0) return 1;
    static bool configured = false;
    if (!configured) {
        DJ_CUDA_CHECK(cudaFuncSetAttribute(hist_partial_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(HistSmem)));
        DJ_CUDA_CHECK(cudaFuncSetAttribute(hist_partial_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(HistSmem)));
        // keep the scratch for the slab partials in the stream-ordered pool across synchronisations
        int dev = 0;
        cudaMemPool_t pool;
        DJ_CUDA_CHECK(cudaGetDevice(&dev));
        DJ_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = UINT64_MAX;
        DJ_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        configured = true;
    }
    const int gx = (code_pitch + kTileLoci - 1) / kTileLoci;
    // four CTAs per SM, all resident at once; they share the reads of their tile column through next_row[]
    int64_t gy = ((int64_t)sms * 4) / gx;
    if (gy > (n_rows + 255) / 256) gy = (n_rows + 255) / 256;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    const int partial_pitch = (code_pitch + 31) & ~31;
    const size_t partial_bytes = sizeof(int32_t) * (size_t)gy * kTileRows * partial_pitch;
    int32_t *partial = nullptr;
    DJ_CUDA_CHECK(cudaMallocAsync(&partial, partial_bytes + sizeof(int32_t) * gx, st));
    int32_t *next_row = partial + partial_bytes / sizeof(int32_t);
    cudaError_t err = cudaMemsetAsync(next_row, 0, sizeof(int32_t) * gx, st);
    if (err == cudaSuccess) {
        if (rows)
            hist_partial_kernel<true><<<dim3(gx, (unsigned)gy), kHistThreads, sizeof(HistSmem), st>>>(
                codes, code_pitch, strand, rows, n_rows, partial, partial_pitch, next_row);
        else
            hist_partial_kernel<false><<<dim3(gx, (unsigned)gy), kHistThreads, sizeof(HistSmem), st>>>(
                codes, code_pitch, strand, rows, n_rows, partial, partial_pitch, next_row);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) {
        hist_reduce_kernel<<<(n_loci + 127) / 128, 128, 0, st>>>(partial, partial_pitch, (int)gy, n_rows, n_loci, counts);
        err = cudaGetLastError();
    }
    cudaFreeAsync(partial, st);
    if (err != cudaSuccess) {
        dj_set_error("dj_hist: %s", cudaGetErrorString(err));
        return 1;
    }
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

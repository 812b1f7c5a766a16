// Per-locus histograms of the code matrix (dj_hist) and the 10-wide window cosine distances (dj_window_cosine).
//
// dj_hist.  98 tokens in 100 are plain matches "=A|C|G|T" (code bytes 0..3) and need no counter of their own: their
// number at a locus is the number of reads minus everything else seen there.  So the pass over the matrix only
// LOOKS for the other tokens -- two 3-input ORs and one masked test per 16 loci -- and counts those:
//   * the matrix is cut into tile columns of 512 loci (16 loci = one 16-byte load per thread).  A CTA of 8 warps
//     works on one column at a time; its warps take blocks of 32 reads from the column's counter and keep eight
//     512-byte row segments in flight per warp; the strands of a block are one coalesced load and one ballot;
//   * a thread whose 16 loci hold an other-than-plain token appends its 16 bytes to its warp's queue in shared
//     memory (one ballot per row); whenever 32 groups are queued the warp drains them with all lanes busy: one
//     shared-memory atomic (two for lowercase / N-anchored tokens) per queued token on the CTA's 10 x 512 tile;
//   * columns differ in work (reads truncated with "=N" tails, indel-rich regions): when a column runs out of
//     reads its CTAs add their tile to the global 10 x L tile (one red.add per non-zero cell) and move to the
//     column that is furthest behind, so that all SMs finish together (scripts/micro/hist_trace.cu: with fixed
//     columns the kernel waited 2.7x longer for its slowest column);
//   * a second small kernel derives the ten output rows from the global tile.
// Algorithmic bytes per read: code_pitch + 1.
//
// Reference: count_indels (src/DAJIN2/core/preprocess/mutation_processing/indel_counter.py:15-31) and
// count_strandless_of_indels (src/DAJIN2/core/preprocess/error_correction/strand_bias_handler.py:18-33).
#include "dj_common.cuh"

namespace {

constexpr int kHistThreads = 256;
constexpr int kHistWarps = kHistThreads / 32;
constexpr int kTileLoci = 512;  // one warp request = 512 loci of one row
constexpr int kTilePad = 16;    // columns 512..527 absorb the lanes of the last tile that lie past the matrix
constexpr int kTilePitch = kTileLoci + kTilePad;
constexpr int kRowUnroll = 8;   // row segments in flight per warp
constexpr int kRowBlock = 32;   // reads a warp takes at a time
constexpr int kQueue = 64;      // 16-byte groups per warp queue (a drain takes 32)
// tile rows: what is counted in shared memory
constexpr int kTileSwPlus = 0;   // 0..2  startswith '+','-','*' on '+' strand reads (strand_bias_handler.py:25-28)
constexpr int kTileSwMinus = 3;  // 3..5  the same on '-' strand reads
constexpr int kTileSkipped = 6;  // 6..8  of those, tokens count_indels skips (lowercase or "=N" anchor, :20-23)
constexpr int kTileOther = 9;    // tokens that are neither plain matches nor start with '+','-','*' ("=N", "=a", invalid)
constexpr int kTileRows = 10;

struct HistSmem {
    int32_t tile[kTileRows][kTilePitch];
    uint4 qdata[kHistWarps][kQueue];
    uint32_t qmeta[kHistWarps][kQueue];  // first locus of the group inside the tile | strand '-' << 31
    // per code byte and strand ('+' first): low byte = 1 + tile row of the token's counter, high byte = 1 + tile row
    // of its second counter (0 = none)
    uint16_t class_lut[2][256];
    int next_col;
};

#ifdef DJ_HIST_TRACE  // scripts/micro/hist_trace.cu: per-CTA time stamps (start, loop start, loop end, end, SM id)
__device__ unsigned long long dj_hist_trace[8192 * 5];
__device__ __forceinline__ void trace_mark(int k) {
    if (threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const int cta = blockIdx.x;
        if (cta < 8192) dj_hist_trace[cta * 5 + k] = t;
        if (k == 0 && cta < 8192) {
            uint32_t sm;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            dj_hist_trace[cta * 5 + 4] = sm;
        }
    }
}
#define DJ_TRACE(k) trace_mark(k)
#else
#define DJ_TRACE(k)
#endif

// 0x80 in every byte of w that is not a plain match (code byte >= 4)
__device__ __forceinline__ uint32_t other_flags(uint32_t w) {
    return (((w | 0x80808080u) - 0x04040404u) | w) & 0x80808080u;
}

template <bool kGather>
__global__ void __launch_bounds__(kHistThreads, 4)
hist_tile_kernel(const uint8_t *__restrict__ codes, int code_pitch, const uint8_t *__restrict__ strand,
                 const int32_t *__restrict__ rows, int64_t n_rows, int32_t *__restrict__ gtile, int gtile_pitch,
                 int32_t *__restrict__ next_row, int n_cols) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    HistSmem &sm = *reinterpret_cast<HistSmem *>(smem_raw);
    DJ_TRACE(0);
    for (int e = threadIdx.x; e < kTileRows * kTilePitch; e += kHistThreads) (&sm.tile[0][0])[e] = 0;
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
    DJ_TRACE(1);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *qdata = sm.qdata[warp];
    uint32_t *qmeta = sm.qmeta[warp];
    const uint8_t *qbytes = reinterpret_cast<const uint8_t *>(qdata);
    const uint32_t lanes_below = (1u << lane) - 1u;
    uint32_t q_in = 0, q_out = 0;  // groups appended / drained so far (warp-uniform)

    // Count the other-than-plain tokens of (up to) 32 queued groups, one group per lane.
    auto drain = [&](uint32_t n) {
        __syncwarp();
        if ((uint32_t)lane < n) {
            const uint32_t slot = (q_out + lane) & (kQueue - 1);
            const uint32_t meta = qmeta[slot];
            const uint4 d = qdata[slot];
            const uint16_t *lut = sm.class_lut[meta >> 31];
            int32_t *cell = &sm.tile[0][meta & 0x7FFFFFFFu];
            const uint8_t *bytes = qbytes + slot * 16;
            // bit 8*j + 4 + k: byte j of word k, i.e. locus first + 4*k + j
            uint32_t m = (other_flags(d.x) >> 3) | (other_flags(d.y) >> 2) | (other_flags(d.z) >> 1) | other_flags(d.w);
            while (m) {
                const int bit = 31 - __clz(m);
                m ^= 1u << bit;
                const int pos = ((bit & 7) - 4) * 4 + (bit >> 3);
                const uint32_t ent = lut[bytes[pos]];
                atomicAdd(cell + ((int)(ent & 0xFFu) - 1) * kTilePitch + pos, 1);
                if (ent >> 8) atomicAdd(cell + ((int)(ent >> 8) - 1) * kTilePitch + pos, 1);
            }
        }
        q_out += n;
        __syncwarp();
    };

    int col = blockIdx.x % n_cols;
    for (;;) {
        // code_pitch is a multiple of 16, so a 16-byte load stays inside the row.  Lanes of the last column that
        // lie past the matrix re-read the row's first bytes and count into the tile's padding columns.
        const bool active = col * kTileLoci + lane * 16 < code_pitch;
        const int col0 = active ? col * kTileLoci + lane * 16 : 0;
        const uint32_t local0 = active ? (uint32_t)lane * 16u : (uint32_t)kTileLoci;

        // One row segment: queue it if any of its 16 code bytes is >= 4.
        auto scan = [&](const uint4 &v, uint32_t minus_bit31) {
            const bool any = ((v.x | v.y | v.z | v.w) & 0xFCFCFCFCu) != 0;
            const uint32_t vote = __ballot_sync(0xffffffffu, any);
            if (any) {
                const uint32_t slot = (q_in + __popc(vote & lanes_below)) & (kQueue - 1);
                qdata[slot] = v;
                qmeta[slot] = local0 | minus_bit31;
            }
            q_in += __popc(vote);
            if (q_in - q_out >= 32) drain(32);
        };

        // Every warp of the CTAs that work on this column takes the next kRowBlock reads from the column's counter.
        for (;;) {
            int64_t t = 0;
            if (lane == 0) t = (int64_t)atomicAdd(next_row + col, kRowBlock);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= n_rows) break;
            const int n_block = (int)min((int64_t)kRowBlock, n_rows - t);
            // lane u holds read t + u of the block: its matrix row and its strand
            int64_t my_row = t + min(lane, n_block - 1);
            if (kGather) my_row = __ldg(rows + my_row);
            const uint32_t minus_bits = __ballot_sync(0xffffffffu, __ldg(strand + my_row) != '+');
            const uint8_t *seg = codes + col0 + (kGather ? (int64_t)0 : t * (int64_t)code_pitch);
            if (n_block == kRowBlock) {
#pragma unroll 1
                for (int u0 = 0; u0 < kRowBlock; u0 += kRowUnroll) {
                    uint4 v[kRowUnroll];
#pragma unroll
                    for (int u = 0; u < kRowUnroll; ++u) {
                        if (kGather) {
                            const int64_t r = __shfl_sync(0xffffffffu, my_row, u0 + u);
                            v[u] = __ldg(reinterpret_cast<const uint4 *>(seg + r * (int64_t)code_pitch));
                        } else {
                            v[u] = __ldg(reinterpret_cast<const uint4 *>(seg + u * (int64_t)code_pitch));
                        }
                    }
                    if (!kGather) seg += kRowUnroll * (int64_t)code_pitch;
                    const uint32_t mb = minus_bits >> u0;
#pragma unroll
                    for (int u = 0; u < kRowUnroll; ++u) scan(v[u], (mb << (31 - u)) & 0x80000000u);
                }
            } else {  // the last, ragged block of the matrix
                for (int u = 0; u < n_block; ++u) {
                    const int64_t r = kGather ? __shfl_sync(0xffffffffu, my_row, u) : (int64_t)u;
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(seg + r * (int64_t)code_pitch));
                    scan(v, (minus_bits << (31 - u)) & 0x80000000u);
                }
            }
        }
        drain(q_in - q_out);
        __syncthreads();
        DJ_TRACE(2);
        // the column has no reads left: add this CTA's tile to the global one, clear it, and pick the column
        // that is furthest behind
        int32_t *out = gtile + col * kTileLoci;
        for (int e = threadIdx.x; e < kTileRows * kTileLoci; e += kHistThreads) {
            const int row = e / kTileLoci, c = e % kTileLoci;
            const int32_t v = sm.tile[row][c];
            if (v != 0) {
                if (col * kTileLoci + c < gtile_pitch) atomicAdd(out + (int64_t)row * gtile_pitch + c, v);
                sm.tile[row][c] = 0;
            }
        }
        if (threadIdx.x < kTileRows * kTilePad) sm.tile[threadIdx.x / kTilePad][kTileLoci + threadIdx.x % kTilePad] = 0;
        if (threadIdx.x == 0) {
            int best = -1;
            int64_t best_next = n_rows;
            for (int k = 1; k < n_cols; ++k) {
                const int c = (col + k) % n_cols;
                const int64_t nx = *((volatile int32_t *)(next_row + c));
                if (nx < best_next) {
                    best_next = nx;
                    best = c;
                }
            }
            sm.next_col = best;
        }
        __syncthreads();
        col = sm.next_col;
        if (col < 0) break;
    }
    DJ_TRACE(3);
}

// Derive the ten output rows (see include/dajin2_b200.h for their meaning) from the global tile.
__global__ void hist_reduce_kernel(const int32_t *__restrict__ gtile, int gtile_pitch, int64_t n_rows, int n_loci,
                                   int32_t *__restrict__ counts) {
    const int locus = blockIdx.x * blockDim.x + threadIdx.x;
    if (locus >= n_loci) return;
    int32_t sum[kTileRows];
#pragma unroll
    for (int k = 0; k < kTileRows; ++k) sum[k] = __ldg(gtile + (int64_t)k * gtile_pitch + locus);
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
    if ((code_pitch & 15) != 0 || ((uintptr_t)codes & 15) != 0 || n_rows > 0x7FF00000) {
        dj_set_error("dj_hist: code matrix must be 16-byte aligned with a pitch that is a multiple of 16 (< 2^31 reads)");
        return 1;
    }
    const int sms = dj_sm_count();
    if (sms <= 0) return 1;
    static bool configured = false;
    if (!configured) {
        DJ_CUDA_CHECK(cudaFuncSetAttribute(hist_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(HistSmem)));
        DJ_CUDA_CHECK(cudaFuncSetAttribute(hist_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(HistSmem)));
        if (dj_scratch_pool_init() != 0) return 1;
        configured = true;
    }
    const int n_cols = (code_pitch + kTileLoci - 1) / kTileLoci;
    // four CTAs per SM, all resident at once; fewer when there are not enough blocks of reads to go round
    int64_t ctas = (int64_t)sms * 4;
    const int64_t blocks = ((n_rows + kRowBlock - 1) / kRowBlock) * n_cols;
    if (ctas > (blocks + kHistWarps - 1) / kHistWarps) ctas = (blocks + kHistWarps - 1) / kHistWarps;
    if (ctas < n_cols) ctas = n_cols;
    const int gtile_pitch = (code_pitch + 31) & ~31;
    const size_t scratch_ints = (size_t)kTileRows * gtile_pitch + (size_t)n_cols;
    int32_t *gtile = nullptr;
    DJ_CUDA_CHECK(cudaMallocAsync(&gtile, sizeof(int32_t) * scratch_ints, st));
    int32_t *next_row = gtile + (size_t)kTileRows * gtile_pitch;
    cudaError_t err = cudaMemsetAsync(gtile, 0, sizeof(int32_t) * scratch_ints, st);
    if (err == cudaSuccess) {
        if (rows)
            hist_tile_kernel<true><<<(unsigned)ctas, kHistThreads, sizeof(HistSmem), st>>>(
                codes, code_pitch, strand, rows, n_rows, gtile, gtile_pitch, next_row, n_cols);
        else
            hist_tile_kernel<false><<<(unsigned)ctas, kHistThreads, sizeof(HistSmem), st>>>(
                codes, code_pitch, strand, rows, n_rows, gtile, gtile_pitch, next_row, n_cols);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) {
        hist_reduce_kernel<<<(n_loci + 127) / 128, 128, 0, st>>>(gtile, gtile_pitch, n_rows, n_loci, counts);
        err = cudaGetLastError();
    }
    cudaFreeAsync(gtile, st);
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

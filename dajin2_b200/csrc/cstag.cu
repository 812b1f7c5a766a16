// dj_cs_midsv_measure / dj_cs_midsv_write: SAM cs tags (long form) -> MIDSV text in HBM, laid out as dj_encode
// reads it (text + row_off + row_len).  One warp per read, persistent grid; the per-read routine is cstag_core.cuh
// (also executed on the host by the CPU suite, see there).  HBM-bound streaming: reads ~1 byte and writes ~3 bytes
// per aligned base (16-byte loads, 16-byte stores out of a per-warp shared-memory stage).
#include "dj_common.cuh"

#include "cstag_core.cuh"

namespace {

constexpr int kWarps = 8;  // per CTA: 8 x DJ_CS_STAGE = 12.5 KB of shared memory; two CTAs per SM (128 registers)

template <bool kWrite>
__global__ void __launch_bounds__(kWarps * 32, 2) cs_midsv_kernel(const uint8_t *__restrict__ cs,
                                                                const int64_t *__restrict__ aln_off,
                                                                const int32_t *__restrict__ aln_pos,
                                                                const uint8_t *__restrict__ aln_lower,
                                                                const int64_t *__restrict__ read_first, int64_t n_reads,
                                                                int32_t ref_len, const uint8_t *__restrict__ ref_seq,
                                                                const int64_t *__restrict__ row_off,
                                                                int32_t *__restrict__ row_len,
                                                                int32_t *__restrict__ row_status,
                                                                int32_t *__restrict__ row_n,
                                                                uint8_t *__restrict__ text) {
    __shared__ __align__(16) uint8_t stage_all[kWrite ? kWarps * DJ_CS_STAGE : 16];
    __shared__ DjCsLut lut;  // the grammar's class and transition tables (cstag_core.cuh)
    djcs_lut_fill(&lut, threadIdx.x, blockDim.x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t *stage = stage_all + (kWrite ? warp * DJ_CS_STAGE : 0);
    const int64_t n_warps = (int64_t)gridDim.x * kWarps;
    for (int64_t r = (int64_t)blockIdx.x * kWarps + warp; r < n_reads; r += n_warps) {
        if (kWrite) {
            int32_t len = row_len[r];
            if (len <= 0) continue;  // dropped by the measuring pass
            djcs_read<true>(lut, cs, aln_off, aln_pos, aln_lower, read_first[r], read_first[r + 1], ref_len, ref_seq,
                            stage, text + row_off[r], &len, nullptr, nullptr, lane);
        } else {
            djcs_read<false>(lut, cs, aln_off, aln_pos, aln_lower, read_first[r], read_first[r + 1], ref_len, ref_seq,
                             stage, nullptr, row_len + r, row_status + r, row_n ? row_n + r : nullptr, lane);
        }
    }
}

int grid_for(int64_t n_reads) {
    const int sms = dj_sm_count();
    if (sms <= 0) return 0;
    const int64_t want = (n_reads + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)sms * 8;  // 8 CTAs of 256 threads per SM
    return (int)(want < cap ? want : cap);
}

}  // namespace

extern "C" int dj_cs_midsv_measure(const uint8_t *cs, const int64_t *aln_off, const int32_t *aln_pos,
                                   const uint8_t *aln_lower, const int64_t *read_first, int64_t n_reads,
                                   int32_t ref_len, const uint8_t *ref_seq, int32_t *row_len, int32_t *row_status,
                                   int32_t *row_n, void *stream) {
    if (n_reads <= 0) return 0;
    if (((uintptr_t)cs & 15) != 0) {
        dj_set_error("dj_cs_midsv_measure: the cs buffer must be 16-byte aligned");
        return 1;
    }
    const int grid = grid_for(n_reads);
    if (grid <= 0) return 1;
    cs_midsv_kernel<false><<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>(
        cs, aln_off, aln_pos, aln_lower, read_first, n_reads, ref_len, ref_seq, nullptr, row_len, row_status, row_n,
        nullptr);
    DJ_LAUNCH_CHECK();
    return 0;
}

extern "C" int dj_cs_midsv_write(const uint8_t *cs, const int64_t *aln_off, const int32_t *aln_pos,
                                 const uint8_t *aln_lower, const int64_t *read_first, int64_t n_reads, int32_t ref_len,
                                 const uint8_t *ref_seq, const int64_t *row_off, const int32_t *row_len, uint8_t *text,
                                 void *stream) {
    if (n_reads <= 0) return 0;
    if (((uintptr_t)cs & 15) != 0) {
        dj_set_error("dj_cs_midsv_write: the cs buffer must be 16-byte aligned");
        return 1;
    }
    const int grid = grid_for(n_reads);
    if (grid <= 0) return 1;
    cs_midsv_kernel<true><<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>(
        cs, aln_off, aln_pos, aln_lower, read_first, n_reads, ref_len, ref_seq, row_off, const_cast<int32_t *>(row_len),
        nullptr, nullptr, text);
    DJ_LAUNCH_CHECK();
    return 0;
}

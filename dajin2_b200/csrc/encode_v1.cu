// Round-1 generic encoder, kept as an on-device cross-check of the unit kernel (dj_encode_v1): one warp streams one
// read, 1 KB of text per trip, EVERY 16-byte chunk goes through the eight comma slots (encode_common.cuh).  It does
// not write checkpoints (its trips are 1 KB, the checkpoint grid of dj_encode is DJ_ENC_CHUNK = 1536 bytes).
// Measured on B200: 0.53 of the HBM copy peak, instruction-bound (profiles/r01d_ncu_encode_kernel.summary.txt).
#include "encode_common.cuh"

#ifndef DJ_ENC_PREFETCH_DIST
#define DJ_ENC_PREFETCH_DIST 1  // trips (of 1 KB per warp) the text is prefetched ahead of its use; 0 = off
#endif
#ifndef DJ_ENC_PREFETCH_LEVEL
#define DJ_ENC_PREFETCH_LEVEL "L1"
#endif

namespace {

constexpr int kWarpsPerCta = 20;  // two CTAs per SM = 40 warps at 48 registers
constexpr int kStageBytes = 512;  // per warp: < 32 carried bytes + two chunks of at most kMaxTokensPerChunk

__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
encode_v1_kernel(const uint8_t *__restrict__ text, const int64_t *__restrict__ row_off,
              const int32_t *__restrict__ row_len, int64_t n_reads, int n_loci, int code_pitch,
              uint8_t *__restrict__ codes, int32_t *__restrict__ match_score, int32_t *__restrict__ n_tokens,
              uint32_t *__restrict__ flags, uint32_t c_four, uint32_t c_shift) {
    // c_four = 4 and c_shift = 2^kLutBits arrive as arguments so that the table address is computed with integer
    // multiply-adds (FMA pipe) instead of being strength-reduced to shifts and masks on the saturated ALU pipe
    __shared__ uint32_t lut[kLutSize];
    __shared__ __align__(16) uint8_t stage_all[kWarpsPerCta][kStageBytes];
    build_token_lut(lut);

    const int lane = threadIdx.x & 31;
    uint8_t *stage = stage_all[threadIdx.x >> 5];
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(lut);
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * kWarpsPerCta;

    for (int64_t r = warp; r < n_reads; r += n_warps) {
        const int64_t off = row_off[r];
        const int len = row_len[r];
        const int head = (int)(off & 15);
        const uint4 *base = reinterpret_cast<const uint4 *>(text + (off - head));
        const int nbytes = head + len;
        const int n_iter = (nbytes >> 9) + 1;   // the chunk that holds byte `nbytes` closes the last token
        const int n_pairs = (n_iter + 1) >> 1;  // two 512-byte chunks per trip; a chunk past the row is all filler
        uint8_t *out = codes + r * (int64_t)code_pitch;

        int tokbase = 0;                  // commas seen so far = tokens closed so far
        uint32_t counts = 0;              // anchors' calc_match contributions (low half) | substitutions (high half)
        uint32_t sticky = 0;              // OR of the kSticky* bits, of rank errors (bits 0..3) and of specials (8..12)
        uint32_t high = 0;                // OR of every text word: bit 7 of a byte set <=> the row is not 7-bit ASCII
        uint32_t carry_w = kFill;         // the word in front of this trip's first byte
        uint8_t *outp = out + lane * 16;  // where this lane's 16 bytes of the next flush go
        int pair = 0;

        // One trip = two chunks of 512 bytes.  stage[0] always holds locus (tokbase & ~31) of the row.
        auto trip = [&](const uint4 &ca, const uint4 &cb) {
            const uint32_t zza = chunk_commas(ca, high), zzb = chunk_commas(cb, high);
            // the word in front of each lane's chunk
            uint32_t wpa = __shfl_up_sync(0xffffffffu, ca.w, 1);
            uint32_t wpb = __shfl_up_sync(0xffffffffu, cb.w, 1);
            const uint32_t last_a = __shfl_sync(0xffffffffu, ca.w, 31);
            if (lane == 0) {
                wpa = carry_w;
                wpb = last_a;
            }
            carry_w = __shfl_sync(0xffffffffu, cb.w, 31);

            // ranks of this lane's commas: one scan for both chunks (16-bit fields)
            const uint32_t cnt_a = __popc(zza), cnt_b = __popc(zzb);
            const uint32_t cnt = cnt_a | (cnt_b << 16);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            const uint32_t totals = __shfl_sync(0xffffffffu, incl, 31);
            const int total_a = (int)(totals & 0xFFFFu), total_b = (int)(totals >> 16);
            const uint32_t excl = incl - cnt;

            // not MIDSV: more commas than a chunk of tokens can hold, or more tokens than the row has room for
            // (tested BEFORE anything is staged or stored: a row with more tokens than the pitch never writes past
            // its own row of the code matrix)
            if (total_a > kMaxTokensPerChunk || total_b > kMaxTokensPerChunk ||
                tokbase + total_a + total_b > code_pitch) {
                sticky |= kStickyBroken;
                pair = 0x3FFFFFFF;  // leave the row: it is flagged, its codes are not defined
                return;
            }
            uint32_t acc = 0;
            const uint32_t sp = stage_s + (uint32_t)(tokbase & 31);
            const uint32_t wrote_a = emit_chunk(acc, sp + (excl & 0xFFFFu), lut_s, wpa, ca, zza, c_four, c_shift);
            const uint32_t wrote_b =
                emit_chunk(acc, sp + (uint32_t)total_a + (excl >> 16), lut_s, wpb, cb, zzb, c_four, c_shift);
            // two commas among bytes 1..3 of a word (an empty or one-character token) are not MIDSV: bits 0..3;
            // the table's "special" counter: bits 8..12
            sticky |= (wrote_a ^ cnt_a) | (wrote_b ^ cnt_b) | ((acc >> 15) & 0x1F00u);
            counts += ((acc >> kStatShift) & 0x3Fu) | ((acc >> 2) & 0x1F0000u);
            // complete 32-byte sectors leave; the incomplete one moves to the front of the staging buffer
            __syncwarp();
            const int staged = (tokbase & 31) + total_a + total_b;
            const int flush_bytes = staged & ~31;
            if (lane * 16 < flush_bytes) *reinterpret_cast<uint4 *>(outp) = *reinterpret_cast<const uint4 *>(stage + lane * 16);
            const uint8_t b = stage[flush_bytes + lane];  // lanes past the tail read stale bytes, never stored
            __syncwarp();
            if (lane < staged - flush_bytes) stage[lane] = b;
            tokbase += total_a + total_b;
            outp += flush_bytes;
        };

        // first trip (the row may start inside its first chunk), trips that lie entirely inside the row, last trips
        trip(load_chunk(base, lane, head, nbytes), load_chunk(base, 32 + lane, head, nbytes));
        const int n_inside = nbytes >> 10;
#pragma unroll 1
        for (++pair; pair < n_inside; ++pair) {
            const uint4 *p = base + pair * 64 + lane;
#if DJ_ENC_PREFETCH_DIST > 0
            // the loads of the trip after next (never past the row: its last trips are not "inside")
            if (pair + DJ_ENC_PREFETCH_DIST < n_inside) {
                asm volatile("prefetch.global." DJ_ENC_PREFETCH_LEVEL " [%0];" ::"l"(p + 64 * DJ_ENC_PREFETCH_DIST));
                asm volatile("prefetch.global." DJ_ENC_PREFETCH_LEVEL " [%0];" ::"l"(p + 64 * DJ_ENC_PREFETCH_DIST + 32));
            }
#endif
            trip(__ldg(p), __ldg(p + 32));
        }
#pragma unroll 1
        for (; pair < n_pairs; ++pair)
            trip(load_chunk(base, pair * 64 + lane, head, nbytes), load_chunk(base, pair * 64 + 32 + lane, head, nbytes));

        const bool broken = (sticky & kStickyBroken) != 0;
        if (!broken) {  // what is left is less than one sector: pad it (deterministically) and store it
            __syncwarp();
            const int staged = tokbase & 31;
            if (lane < 32 - staged) stage[staged + lane] = 0;  // reads as a plain match: dj_hist looks at every byte of the pitch
            __syncwarp();
            if (staged > 0 && lane < 2 && (tokbase & ~31) < code_pitch)
                *reinterpret_cast<uint4 *>(outp) = *reinterpret_cast<const uint4 *>(stage + lane * 16);
        }

        counts = __reduce_add_sync(0xffffffffu, counts);  // 16-bit fields: dj_encode limits n_loci accordingly
        sticky = __reduce_or_sync(0xffffffffu, sticky | (high & 0x80808080u));
        const uint32_t mism = counts & 0xFFFFu, n_sub = counts >> 16;
        // every inserted base is the three bytes "+B|": what the anchors and commas do not explain is 3 x inserted
        const int ins3 = len - 3 * tokbase + 1 - (int)n_sub;
        bool bad = (sticky & (kStickyBroken | 0x8080808Fu)) != 0 || ins3 < 0 || ins3 % 3 != 0;
        uint32_t mismatches = mism + (uint32_t)(ins3 / 3);
        if ((sticky & 0x1F00u) != 0 && !broken) {
            __syncwarp();
            bool has_invalid = false;
            mismatches = recheck_row(text + off, len, out, min(tokbase, code_pitch), lane, &has_invalid);
            bad = bad || has_invalid;
        }
        if (lane == 0) {
            match_score[r] = -(int32_t)mismatches;
            n_tokens[r] = tokbase;
            flags[r] = (tokbase != n_loci ? DJ_FLAG_TOKEN_COUNT : 0u) | (bad ? DJ_FLAG_BAD_TOKEN : 0u);
        }
        __syncwarp();
    }
}


}  // namespace

extern "C" int dj_encode_v1(const uint8_t *text, const int64_t *row_off, const int32_t *row_len, int64_t n_reads,
                            int32_t n_loci, int32_t code_pitch, uint8_t *codes, int32_t *match_score,
                            int32_t *n_tokens, uint32_t *flags, void *stream) {
    if (n_reads <= 0) return 0;
    if (((uintptr_t)text & 15) != 0 || ((uintptr_t)codes & 15) != 0) {
        dj_set_error("dj_encode_v1: text and codes must be 16-byte aligned");
        return 1;
    }
    if (n_loci > 400000 || code_pitch < n_loci || (code_pitch & 31) != 0) {
        dj_set_error("dj_encode_v1: unsupported n_loci %d / code_pitch %d", n_loci, code_pitch);
        return 1;
    }
    const int sms = dj_sm_count();
    if (sms <= 0) return 1;
    int64_t ctas = (n_reads + kWarpsPerCta - 1) / kWarpsPerCta;
    const int64_t resident = (int64_t)sms * 2;
    if (ctas > resident) ctas = resident;
    encode_v1_kernel<<<(unsigned)ctas, kWarpsPerCta * 32, 0, (cudaStream_t)stream>>>(
        text, row_off, row_len, n_reads, n_loci, code_pitch, codes, match_score, n_tokens, flags, 4u,
        (uint32_t)kLutSize);
    DJ_LAUNCH_CHECK();
    return 0;
}

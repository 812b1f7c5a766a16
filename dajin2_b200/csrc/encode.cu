// MIDSV text -> read x locus code matrix (dj_encode).
//
// One warp streams one read.  Every iteration the 32 lanes load 32 consecutive, 16-byte aligned chunks of the
// row (one coalesced 512-byte request, software-prefetched one iteration ahead) and work on the COMMAS of
// their chunk: a comma closes exactly one token, the rank of the comma inside the row is the token's locus,
// and the four bytes in front of it identify the token's anchor ("=X", "-X" or "*XY") and whether the token is
// an insertion ("+B|+B|...|<anchor>": the byte in front of the anchor is '|').  So a token of any length is
// translated with ONE perfect-hash look-up keyed on the bytes that precede its closing comma -- there is no
// per-token loop, no slow path for insertions and no divergence: a 32-bit word holds at most two commas
// (tokens are >= 2 characters), one at byte 0 and one among bytes 1..3, which gives eight predicated slots per
// lane.  Code bytes are staged in shared memory and leave the SM as aligned 16-byte stores of complete 32-byte
// sectors.  HBM traffic is the text once in and the code bytes once out (algorithmic bytes per read:
// row_len + n_loci + 16).
//
// calc_match (src/DAJIN2/core/classification/classifier.py:11-24) = #'+' + #'-' + #'*' + #lowercase characters
// falls out of the same pass without touching the bytes again: '-' and '*' only occur as anchor operators and
// the anchors' lowercase letters are known from the look-up (summed as 8-bit fields of one register), and the
// number of inserted bases follows from the row length, because every inserted base is the three bytes "+B|":
//     row_len = 3 * n_inserted + 2 * n_tokens + n_substitutions + (n_tokens - 1).
// Only reads that carry a lowercase insertion (an insertion inside an inversion) are recounted byte by byte.
//
// Reference behaviour covered here: the tokenisation every pass of the reference starts with
// (`json.loads` + `MIDSV.split(",")`, src/DAJIN2/utils/fileio.py:49-52) and calc_match.
#include "dj_common.cuh"

namespace {

constexpr int kWarpsPerCta = 16;
constexpr int kLutBits = 11;
constexpr int kLutSize = 1 << kLutBits;
// injective over the 300 keys enumerated in build_token_lut (tests/test_host_logic.py re-derives this)
constexpr uint32_t kLutMagic = 0x0ccd5009u;
constexpr uint32_t kCommas = 0x2C2C2C2Cu;
constexpr int kStageBytes = 512;     // per warp: < 32 carried bytes + two chunks of at most kMaxTokensPerIter + 16
constexpr int kMaxTokensPerIter = 176;  // a valid chunk of 512 bytes closes at most 171 tokens

// per-token statistics packed in the second word of a table entry (8-bit fields, drained every 8 iterations)
constexpr uint32_t kStatSub = 1u << 8;        // anchor is a substitution "*XY"
constexpr uint32_t kStatLowerIns = 1u << 16;  // lowercase insertion: its inserted bases need the byte recount

__device__ __forceinline__ uint32_t lut_slot(uint32_t key) { return (key * kLutMagic) >> (32 - kLutBits); }

// Key of the token closed by a comma: high nibble of the byte four positions before the comma, then the three
// bytes in front of the comma, then the comma's own low nibble (0xC).  For "=X"/"-X" anchors the three bytes are
// <separator> <op> <X>; for "*XY" they are '*' X Y and the separator (',' or '|') is the nibble.
__device__ __forceinline__ uint32_t make_key(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
    return (b0 >> 4) | (b1 << 4) | (b2 << 12) | (b3 << 20) | (0xCu << 28);
}

// Table entry: x = (key & ~0xFF) | code byte, y = statistics.  Empty slots carry a key no token can produce.
__device__ void build_token_lut(uint2 *lut) {
    for (int e = threadIdx.x; e < kLutSize; e += blockDim.x) lut[e] = make_uint2(DJ_CODE_INVALID, 0u);
    __syncthreads();
    for (int k = threadIdx.x; k < 300; k += blockDim.x) {
        uint32_t key, code;
        bool lower;
        if (k < 200) {  // <any byte> <',' or '|'> <'=' or '-'> <base>
            // the byte before the separator is the last letter of the previous token, an inserted base, or the
            // ',' padding in front of the row: high nibbles 2 (','), 4 (ACGN), 5 (T), 6 (acgn), 7 (t)
            const uint32_t nib = k / 40;  // 0 -> 0x2_, 1 -> 0x4_, 2 -> 0x5_, 3 -> 0x6_, 4 -> 0x7_
            const uint32_t nib_byte = nib == 0 ? 0x20u : (nib + 3u) << 4;
            const int rem = k % 40, x = rem % 10;
            const bool ins = rem >= 20, del = (rem % 20) >= 10;
            lower = x >= 5;
            key = make_key(nib_byte, ins ? '|' : ',', del ? '-' : '=', (uint8_t)dj_base_char(x % 5, lower));
            code = (ins ? DJ_CODE_INS : 0) | (uint32_t)((del ? 10 : 0) + 5 * (lower ? 1 : 0) + x % 5);
        } else {  // <',' or '|'> '*' <ref base> <read base>
            const int j = k - 200, rem = j % 50;
            const bool ins = j >= 50;
            lower = rem >= 25;
            const int r = (rem % 25) / 5, b = rem % 5;
            key = make_key(ins ? '|' : ',', '*', (uint8_t)dj_base_char(r, lower), (uint8_t)dj_base_char(b, lower));
            code = (ins ? DJ_CODE_INS : 0) | (uint32_t)(20 + 25 * (lower ? 1 : 0) + 5 * r + b);
        }
        const int anchor = code & 0x7F;
        uint32_t stat = (uint32_t)dj_anchor_mismatch(anchor);
        if (anchor >= 20) stat |= kStatSub;
        if ((code & DJ_CODE_INS) && lower) stat |= kStatLowerIns;
        lut[lut_slot(key)] = make_uint2((key & 0xFFFFFF00u) | code, stat);
    }
    __syncthreads();
}

// 0x80 in every byte of w that equals ','.  Exact when every byte of w is 7-bit ASCII; rows with other bytes
// are flagged DJ_FLAG_BAD_TOKEN by the caller.
__device__ __forceinline__ uint32_t comma_flags(uint32_t w) {
    return ~((w ^ kCommas) + 0x7F7F7F7Fu) & 0x80808080u;
}

// 0xFF in every byte of the word at byte offset g whose position lies inside [lo, hi)
__device__ __forceinline__ uint32_t valid_bytes(int g, int lo, int hi) {
    uint32_t m = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (g + b >= lo && g + b < hi) m |= 0xFFu << (8 * b);
    return m;
}

// Chunk `c` (16 bytes) of the row's aligned span, with every byte outside the row replaced by ','.
__device__ __forceinline__ uint4 load_chunk(const uint4 *base, int c, int head, int nbytes) {
    const int g = c * 16;
    if (g >= nbytes) return make_uint4(kCommas, kCommas, kCommas, kCommas);
    uint4 v = __ldg(base + c);
    if (g < head || g + 16 > nbytes) {
        uint32_t m;
        m = valid_bytes(g, head, nbytes);      v.x = (v.x & m) | (kCommas & ~m);
        m = valid_bytes(g + 4, head, nbytes);  v.y = (v.y & m) | (kCommas & ~m);
        m = valid_bytes(g + 8, head, nbytes);  v.z = (v.z & m) | (kCommas & ~m);
        m = valid_bytes(g + 12, head, nbytes); v.w = (v.w & m) | (kCommas & ~m);
    }
    return v;
}

// calc_match counted byte by byte (only for reads with a lowercase insertion)
__device__ __noinline__ uint32_t recount_mismatches(const uint8_t *row, int len, int lane) {
    uint32_t n = 0;
    for (int p = lane; p < len; p += 32) {
        const uint32_t c = row[p];
        n += (c == '+' || c == '-' || c == '*' || (c >= 'a' && c <= 'z')) ? 1u : 0u;
    }
    return __reduce_add_sync(0xffffffffu, n);
}

// Translate the token closed by one comma.  `key` = the 32 bits in front of the comma (see make_key), `flag` != 0
// when the slot holds a comma.  Written in PTX so that the slot stays branch-free: one table look-up, and the
// statistics, the key check, the one-byte store of the code and the staging pointer are predicated.
__device__ __forceinline__ void emit_token(uint32_t &acc, uint32_t &diff, uint32_t &sp, uint32_t lut_s, uint32_t key,
                                           uint32_t flag) {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " .reg .u32 h, ex, ey;\n"
        " setp.ne.u32 p, %5, 0;\n"
        " mul.lo.u32 h, %3, %6;\n"
        " shr.u32 h, h, %7;\n"
        " and.b32 h, h, %8;\n"
        " add.u32 h, h, %4;\n"
        " ld.shared.v2.u32 {ex, ey}, [h];\n"
        " @p add.u32 %0, %0, ey;\n"
        " @p lop3.b32 %1, %1, ex, %3, 0xF6;\n"
        " @p st.shared.u8 [%2], ex;\n"
        " @p add.u32 %2, %2, 1;\n"
        "}\n"
        : "+r"(acc), "+r"(diff), "+r"(sp)
        : "r"(key), "r"(lut_s), "r"(flag), "n"(kLutMagic), "n"(32 - kLutBits - 3), "n"((kLutSize - 1) << 3)
        : "memory");
}

__device__ __forceinline__ uint32_t find_msb(uint32_t m) {
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(m));
    return r;
}

// The (at most two) commas of one word: byte 0, then the single comma bytes 1..3 can hold.
__device__ __forceinline__ void emit_word(uint32_t &acc, uint32_t &diff, uint32_t &sp, uint32_t lut_s, uint32_t prev,
                                          uint32_t word, uint32_t z) {
    emit_token(acc, diff, sp, lut_s, __funnelshift_r(prev, word, 4), z & 0x80u);
    const uint32_t m = z & 0x80808000u;
    emit_token(acc, diff, sp, lut_s, __funnelshift_r(prev, word, find_msb(m) - 3u), m);
}

__device__ __forceinline__ int emit_chunk(uint32_t &acc, uint32_t &diff, uint32_t sp, uint32_t lut_s, uint32_t wp,
                                          const uint4 &c, const uint32_t (&z)[4]) {
    const uint32_t sp0 = sp;
    emit_word(acc, diff, sp, lut_s, wp, c.x, z[0]);
    emit_word(acc, diff, sp, lut_s, c.x, c.y, z[1]);
    emit_word(acc, diff, sp, lut_s, c.y, c.z, z[2]);
    emit_word(acc, diff, sp, lut_s, c.z, c.w, z[3]);
    return (int)(sp - sp0);
}

__device__ __forceinline__ void chunk_commas(const uint4 &c, uint32_t (&z)[4], uint32_t &high) {
    high |= c.x | c.y;
    high |= c.z | c.w;
    z[0] = comma_flags(c.x);
    z[1] = comma_flags(c.y);
    z[2] = comma_flags(c.z);
    z[3] = comma_flags(c.w);
}

// commas count inside [head, nbytes] only: the row's own plus one terminator right after it
__device__ __forceinline__ void mask_commas(uint32_t (&z)[4], int g0, int head, int nbytes) {
#pragma unroll
    for (int k = 0; k < 4; ++k) z[k] &= valid_bytes(g0 + 4 * k, head, nbytes + 1);
}

__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
encode_kernel(const uint8_t *__restrict__ text, const int64_t *__restrict__ row_off,
              const int32_t *__restrict__ row_len, int64_t n_reads, int n_loci, int code_pitch, int ckpt_pitch,
              uint8_t *__restrict__ codes, int32_t *__restrict__ match_score, int32_t *__restrict__ n_tokens,
              uint32_t *__restrict__ flags, uint32_t *__restrict__ ckpt) {
    __shared__ uint2 lut[kLutSize];
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
        const int n_iter = (nbytes >> 9) + 1;  // the chunk that holds byte `nbytes` closes the last token
        const int n_pairs = (n_iter + 1) >> 1;  // two 512-byte chunks per trip; a chunk past the row is all padding
        uint8_t *out = codes + r * (int64_t)code_pitch;
        uint32_t *ck = ckpt + r * (int64_t)ckpt_pitch;

        int tokbase = 0;      // commas seen so far = tokens closed so far
        int staged_from = 0;  // locus of stage[0] (a multiple of 32)
        uint32_t acc = 0, diff = 0, mism = 0, n_sub = 0, n_lower_ins = 0, high = 0;
        bool broken = false;
        uint32_t carry_w = kCommas;  // the word in front of this trip's first byte
        uint4 ca = load_chunk(base, lane, head, nbytes);
        uint4 cb = load_chunk(base, 32 + lane, head, nbytes);

        for (int pair = 0; pair < n_pairs; ++pair) {
            const uint4 na = load_chunk(base, (2 * pair + 2) * 32 + lane, head, nbytes);
            const uint4 nb = load_chunk(base, (2 * pair + 3) * 32 + lane, head, nbytes);
            const bool last = pair == n_pairs - 1;
            uint32_t za[4], zb[4];
            chunk_commas(ca, za, high);
            chunk_commas(cb, zb, high);
            if (pair == 0 || last) {
                mask_commas(za, (2 * pair * 32 + lane) * 16, head, nbytes);
                mask_commas(zb, ((2 * pair + 1) * 32 + lane) * 16, head, nbytes);
            }
            // the word in front of each lane's chunk
            uint32_t wpa = __shfl_up_sync(0xffffffffu, ca.w, 1);
            uint32_t wpb = __shfl_up_sync(0xffffffffu, cb.w, 1);
            const uint32_t last_a = __shfl_sync(0xffffffffu, ca.w, 31);
            const uint32_t front_a = carry_w;
            if (lane == 0) {
                wpa = carry_w;
                wpb = last_a;
            }
            carry_w = __shfl_sync(0xffffffffu, cb.w, 31);

            // ranks of this lane's commas: one scan for both chunks (16-bit fields)
            const uint32_t cnt_a = __popc(za[0] | (za[1] >> 1) | (za[2] >> 2) | (za[3] >> 3));
            const uint32_t cnt_b = __popc(zb[0] | (zb[1] >> 1) | (zb[2] >> 2) | (zb[3] >> 3));
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

            // checkpoints: tokens that START before each chunk's first byte (dj_find_token resumes from them)
            if (lane < 2) {
                const uint32_t front = lane == 0 ? front_a : last_a;
                uint32_t v = (uint32_t)tokbase + (lane == 0 ? 0u : (uint32_t)total_a) + ((front >> 24) == ',' ? 0u : 1u);
                if (pair == 0 && lane == 0) v = 0;
                ck[2 * pair + lane] = v;
            }

            if (total_a > kMaxTokensPerIter || total_b > kMaxTokensPerIter) broken = true;  // not MIDSV
            if (!broken) {
                const uint32_t sp = stage_s + (uint32_t)(tokbase - staged_from);
                const int wrote_a = emit_chunk(acc, diff, sp + (excl & 0xFFFFu), lut_s, wpa, ca, za);
                const int wrote_b = emit_chunk(acc, diff, sp + (uint32_t)total_a + (excl >> 16), lut_s, wpb, cb, zb);
                // two commas among bytes 1..3 of a word (an empty or one-character token) are not MIDSV
                if ((uint32_t)wrote_a != cnt_a || (uint32_t)wrote_b != cnt_b) diff |= 0x100u;
            }
            tokbase += total_a + total_b;

            if ((pair & 3) == 3 || last) {  // drain the 8-bit statistics fields (at most 48 tokens per lane since)
                mism += acc & 0xFFu;
                n_sub += (acc >> 8) & 0xFFu;
                n_lower_ins += acc >> 16;
                acc = 0;
            }
            if (!broken) {
                __syncwarp();
                const int staged = tokbase - staged_from;
                if (last) {
                    if (lane < 16) stage[staged + lane] = DJ_CODE_INVALID;  // deterministic padding of the row
                    __syncwarp();
                }
                const int flush_bytes = last ? ((staged + 15) & ~15) : (staged & ~31);
                if (lane * 16 < flush_bytes) {
                    const int o = staged_from + lane * 16;
                    if (o < code_pitch)
                        *reinterpret_cast<uint4 *>(out + o) = *reinterpret_cast<const uint4 *>(stage + lane * 16);
                }
                if (!last && flush_bytes > 0) {  // carry the incomplete sector to the front of the staging buffer
                    const int tail = staged - flush_bytes;
                    uint8_t b = 0;
                    if (lane < tail) b = stage[flush_bytes + lane];
                    __syncwarp();
                    if (lane < tail) stage[lane] = b;
                    staged_from += flush_bytes;
                }
                __syncwarp();
            }
            ca = na;
            cb = nb;
        }

        mism = __reduce_add_sync(0xffffffffu, mism);
        n_sub = __reduce_add_sync(0xffffffffu, n_sub);
        n_lower_ins = __reduce_add_sync(0xffffffffu, n_lower_ins);
        diff = __reduce_or_sync(0xffffffffu, diff);
        high = __reduce_or_sync(0xffffffffu, high);
        // every inserted base is the three bytes "+B|": what the anchors and commas do not explain is 3 x inserted
        const int ins3 = len - 3 * tokbase + 1 - (int)n_sub;
        const bool bad = broken || (diff >> 8) != 0 || (high & 0x80808080u) != 0 || ins3 < 0 || ins3 % 3 != 0;
        uint32_t mismatches = mism + (uint32_t)(ins3 / 3);
        if (n_lower_ins != 0 && !bad) mismatches = recount_mismatches(text + off, len, lane);
        if (lane == 0) {
            match_score[r] = -(int32_t)mismatches;
            n_tokens[r] = tokbase;
            flags[r] = (tokbase != n_loci ? DJ_FLAG_TOKEN_COUNT : 0u) | (bad ? DJ_FLAG_BAD_TOKEN : 0u);
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" int32_t dj_code_pitch(int32_t n_loci) { return (n_loci + 31) & ~31; }

extern "C" int32_t dj_ckpt_pitch(int32_t max_row_len) { return (max_row_len + 15) / DJ_ENC_CHUNK + 3; }

extern "C" int dj_encode(const uint8_t *text, const int64_t *row_off, const int32_t *row_len, int64_t n_reads,
                         int32_t n_loci, int32_t code_pitch, int32_t ckpt_pitch, uint8_t *codes,
                         int32_t *match_score, int32_t *n_tokens, uint32_t *flags, uint32_t *ckpt, void *stream) {
    if (n_reads <= 0) return 0;
    if (((uintptr_t)text & 15) != 0 || ((uintptr_t)codes & 15) != 0) {
        dj_set_error("dj_encode: text and codes must be 16-byte aligned");
        return 1;
    }
    if (code_pitch < n_loci || (code_pitch & 31) != 0) {
        dj_set_error("dj_encode: code_pitch %d must be a multiple of 32 and >= n_loci %d", code_pitch, n_loci);
        return 1;
    }
    const int sms = dj_sm_count();
    if (sms <= 0) return 1;
    int64_t ctas = (n_reads + kWarpsPerCta - 1) / kWarpsPerCta;
    const int64_t resident = (int64_t)sms * 2;  // 2 CTAs of 512 threads = 32 warps per SM (persistent)
    if (ctas > resident) ctas = resident;
    encode_kernel<<<(unsigned)ctas, kWarpsPerCta * 32, 0, (cudaStream_t)stream>>>(
        text, row_off, row_len, n_reads, n_loci, code_pitch, ckpt_pitch, codes, match_score, n_tokens, flags, ckpt);
    DJ_LAUNCH_CHECK();
    return 0;
}

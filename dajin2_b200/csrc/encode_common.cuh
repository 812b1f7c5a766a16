// Pieces shared by the MIDSV encoders (encode.cu: the unit kernel behind dj_encode; encode_v1.cu: the round-1
// generic kernel kept as an on-device cross-check, dj_encode_v1).
//
// Tokens are identified FROM THEIR CLOSING COMMA: a comma closes exactly one token, the rank of the comma inside the
// row is the token's locus, and the four bytes in front of it identify the token's anchor ("=X", "-X" or "*XY") and
// whether the token is an insertion ("+B|+B|...|<anchor>": the byte in front of the anchor is '|').  So a token of
// any length is translated with ONE perfect-hash look-up keyed on the bytes that precede its closing comma.  A
// 32-bit word holds at most two commas (tokens are >= 2 characters), one at byte 0 and one among bytes 1..3, which
// gives eight predicated, branch-free slots per 16-byte chunk (emit_chunk).
//
// calc_match (src/DAJIN2/core/classification/classifier.py:11-24) = #'+' + #'-' + #'*' + #lowercase characters
// falls out of the same pass: '-' and '*' only occur as anchor operators and the anchors' lowercase letters are
// known from the look-up (the table entry carries them in bit fields that are summed with one add per token), and
// the number of inserted bases follows from the row length, because every inserted base is the three bytes "+B|":
//     row_len = 3 * n_inserted + 2 * n_tokens + n_substitutions + (n_tokens - 1).
// Only reads that carry a lowercase insertion (an insertion inside an inversion) are recounted byte by byte.
//
// Validation: a token whose anchor is not in the table reads an empty slot (code 0x7F) with probability
// 1 - 320/8192 = 0.96 and the row is flagged; rows are also flagged when the token count, the length identity
// above or 7-bit ASCII fail.  (The reference does not validate its input at all.)
#pragma once
#include "dj_common.cuh"

#ifndef DJ_ENC_PRED_LDS
#define DJ_ENC_PRED_LDS 0  // 1: slots without a comma do not read the table (fewer bank conflicts, one more move)
#endif
#if DJ_ENC_PRED_LDS == 2
#define DJ_ENC_LDS " @p ld.shared.u32 e, [h];\n"
#elif DJ_ENC_PRED_LDS
#define DJ_ENC_LDS " mov.u32 e, 0;\n @p ld.shared.u32 e, [h];\n"
#else
#define DJ_ENC_LDS " ld.shared.u32 e, [h];\n"
#endif
#ifndef DJ_ENC_HASH_MODE
#define DJ_ENC_HASH_MODE 1  // 0: mul, shr, and   1: mul, shr, mad   2: mul, mul.hi, mad (no ALU-pipe op)
#endif

namespace {

constexpr int kLutBits = 13;
constexpr int kLutSize = 1 << kLutBits;
// (key * kLutMagic) >> 19 is injective over the 320 keys enumerated in build_token_lut and spreads the frequent
// ones over distinct shared-memory banks (tests/test_host_logic.py re-derives both properties)
constexpr uint32_t kLutMagic = 0xda7e7ab5u;
constexpr uint32_t kCommas = 0x2C2C2C2Cu;
constexpr uint32_t kFill = 0x20202020u;  // bytes outside the row read as ' ' (never a comma, high nibble 2)
constexpr int kMaxTokensPerChunk = 176;  // a valid chunk of 512 bytes closes at most 171 tokens

// Table entry (one 32-bit word): bits 0..7 the code byte; above it three counters that are summed with one add
// per token and drained after every trip (at most 16 slots per lane and trip, so the code bytes sum to < 2^12):
constexpr int kStatShift = 12;
constexpr uint32_t kStatMismatch = 1u << 12;  // bits 12..17: calc_match contribution of the anchor (0..3 per token)
constexpr uint32_t kStatSub = 1u << 18;       // bits 18..22: anchor is a substitution "*XY"
constexpr uint32_t kStatSpecial = 1u << 23;   // bits 23..27: empty slot (unknown anchor) or lowercase insertion
constexpr uint32_t kStickyBroken = 1u << 16;  // per-row flag word: see the main loop

__device__ __forceinline__ uint32_t lut_slot(uint32_t key) { return (key * kLutMagic) >> (32 - kLutBits); }

// Key of the token closed by a comma: high nibble of the byte four positions before the comma, then the three
// bytes in front of the comma, then the comma's own low nibble (0xC).  For "=X"/"-X" anchors the three bytes are
// <separator> <op> <X>; for "*XY" they are '*' X Y and the separator (',' or '|') is the nibble.
__device__ __forceinline__ uint32_t make_key(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
    return (b0 >> 4) | (b1 << 4) | (b2 << 12) | (b3 << 20) | (0xCu << 28);
}

__device__ void build_token_lut(uint32_t *lut) {
    for (int e = threadIdx.x; e < kLutSize; e += blockDim.x) lut[e] = DJ_CODE_INVALID | kStatSpecial;
    __syncthreads();
    for (int k = threadIdx.x; k < 320; k += blockDim.x) {
        uint32_t key, code;
        bool lower;
        if (k < 200 || k >= 300) {  // <any byte> <separator> <'=' or '-'> <base>
            // the byte before the separator is the last letter of the previous token or an inserted base: high
            // nibbles 4 (ACGN), 5 (T), 6 (acgn), 7 (t); nibble 2 is kept for a ',' there.  k >= 300: the first
            // token of the row, which follows the ' ' filler instead of a separator.
            const bool first = k >= 300;
            const int kk = first ? k - 300 : k;
            const uint32_t nib = first ? 0 : kk / 40;  // 0 -> 0x2_, 1 -> 0x4_, 2 -> 0x5_, 3 -> 0x6_, 4 -> 0x7_
            const uint32_t nib_byte = nib == 0 ? 0x20u : (nib + 3u) << 4;
            const int rem = kk % 40, x = rem % 10;
            const bool ins = !first && rem >= 20, del = (rem % 20) >= 10;
            lower = x >= 5;
            key = make_key(nib_byte, first ? ' ' : (ins ? '|' : ','), del ? '-' : '=',
                           (uint8_t)dj_base_char(x % 5, lower));
            code = (ins ? DJ_CODE_INS : 0) | (uint32_t)((del ? 10 : 0) + 5 * (lower ? 1 : 0) + x % 5);
        } else {  // <',' or '|' or the filler> '*' <ref base> <read base>
            const int j = k - 200, rem = j % 50;
            const bool ins = j >= 50;
            lower = rem >= 25;
            const int r = (rem % 25) / 5, b = rem % 5;
            key = make_key(ins ? '|' : ',', '*', (uint8_t)dj_base_char(r, lower), (uint8_t)dj_base_char(b, lower));
            code = (ins ? DJ_CODE_INS : 0) | (uint32_t)(20 + 25 * (lower ? 1 : 0) + 5 * r + b);
        }
        const int anchor = code & 0x7F;
        uint32_t entry = code | (uint32_t)dj_anchor_mismatch(anchor) * kStatMismatch;
        if (anchor >= 20) entry |= kStatSub;
        if ((code & DJ_CODE_INS) && lower) entry |= kStatSpecial;
        lut[lut_slot(key)] = entry;
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

// word of the padded row at byte offset g: the row's bytes inside [head, nbytes), ONE ',' at `nbytes` (it closes
// the last token) and the ' ' filler everywhere else
__device__ __forceinline__ uint32_t pad_word(uint32_t w, int g, int head, int nbytes) {
    const uint32_t m = valid_bytes(g, head, nbytes);
    const uint32_t t = valid_bytes(g, nbytes, nbytes + 1);
    return (w & m) | (((kCommas & t) | (kFill & ~t)) & ~m);
}

// Chunk `c` (16 bytes) of the row's aligned span, padded as described above.
__device__ __forceinline__ uint4 load_chunk(const uint4 *base, int c, int head, int nbytes) {
    const int g = c * 16;
    if (g >= nbytes) return make_uint4(g == nbytes ? (kFill ^ 0x20u ^ 0x2Cu) : kFill, kFill, kFill, kFill);
    uint4 v = __ldg(base + c);
    if (g < head || g + 16 > nbytes) {
        v.x = pad_word(v.x, g, head, nbytes);
        v.y = pad_word(v.y, g + 4, head, nbytes);
        v.z = pad_word(v.z, g + 8, head, nbytes);
        v.w = pad_word(v.w, g + 12, head, nbytes);
    }
    return v;
}

// Rare path (a token read an empty table slot, or a lowercase insertion): look for the invalid code in the row's
// codes and count calc_match byte by byte.
__device__ __noinline__ uint32_t recheck_row(const uint8_t *row, int len, const uint8_t *out, int n_codes, int lane,
                                             bool *has_invalid) {
    uint32_t n = 0, invalid = 0;
    for (int p = lane; p < len; p += 32) {
        const uint32_t c = row[p];
        n += (c == '+' || c == '-' || c == '*' || (c >= 'a' && c <= 'z')) ? 1u : 0u;
    }
    for (int p = lane; p < n_codes; p += 32) invalid |= (__ldcg(out + p) & 0x7F) == DJ_CODE_INVALID ? 1u : 0u;
    *has_invalid = __reduce_or_sync(0xffffffffu, invalid) != 0;
    return __reduce_add_sync(0xffffffffu, n);
}

// Translate the token closed by one comma.  `key` = the 32 bits in front of the comma (see make_key), `flag` != 0
// when the slot holds a comma.  Written in PTX so that the slot stays branch-free: one table read; the add of the
// counters, the one-byte store of the code and the staging pointer are predicated.
__device__ __forceinline__ void emit_token(uint32_t &acc, uint32_t &sp, uint32_t lut_s, uint32_t key, uint32_t flag,
                                           uint32_t c_four, uint32_t c_shift) {
#if DJ_ENC_HASH_MODE == 0
    (void)c_four; (void)c_shift;
    asm volatile(
        "{\n .reg .pred p;\n .reg .u32 h, e;\n"
        " setp.ne.u32 p, %4, 0;\n"
        " mul.lo.u32 h, %2, %5;\n"
        " shr.u32 h, h, %6;\n"
        " and.b32 h, h, %7;\n"
        " add.u32 h, h, %3;\n"
        DJ_ENC_LDS
        " @p add.u32 %0, %0, e;\n"
        " @p st.shared.u8 [%1], e;\n"
        " @p add.u32 %1, %1, 1;\n}\n"
        : "+r"(acc), "+r"(sp)
        : "r"(key), "r"(lut_s), "r"(flag), "n"(kLutMagic), "n"(32 - kLutBits - 2), "n"((kLutSize - 1) << 2)
        : "memory");
#elif DJ_ENC_HASH_MODE == 1
    (void)c_shift;
    asm volatile(
        "{\n .reg .pred p;\n .reg .u32 h, e;\n"
        " setp.ne.u32 p, %4, 0;\n"
        " mul.lo.u32 h, %2, %6;\n"
        " shr.u32 h, h, %7;\n"
        " mad.lo.u32 h, h, %5, %3;\n"
        DJ_ENC_LDS
        " @p add.u32 %0, %0, e;\n"
        " @p st.shared.u8 [%1], e;\n"
        " @p add.u32 %1, %1, 1;\n}\n"
        : "+r"(acc), "+r"(sp)
        : "r"(key), "r"(lut_s), "r"(flag), "r"(c_four), "n"(kLutMagic), "n"(32 - kLutBits)
        : "memory");
#else
    asm volatile(
        "{\n .reg .pred p;\n .reg .u32 h, e;\n"
        " setp.ne.u32 p, %4, 0;\n"
        " mul.lo.u32 h, %2, %7;\n"
        " mul.hi.u32 h, h, %6;\n"
        " mad.lo.u32 h, h, %5, %3;\n"
        DJ_ENC_LDS
        " @p add.u32 %0, %0, e;\n"
        " @p st.shared.u8 [%1], e;\n"
        " @p add.u32 %1, %1, 1;\n}\n"
        : "+r"(acc), "+r"(sp)
        : "r"(key), "r"(lut_s), "r"(flag), "r"(c_four), "r"(c_shift), "n"(kLutMagic)
        : "memory");
#endif
}

__device__ __forceinline__ uint32_t find_msb(uint32_t m) {
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(m));
    return r;
}

// Comma flags of one 16-byte chunk in one word: bit 7-k of byte j is set when byte j of word k is a comma.
__device__ __forceinline__ uint32_t chunk_commas(const uint4 &c, uint32_t &high) {
    high |= c.x | c.y;
    high |= c.z | c.w;
    return comma_flags(c.x) | (comma_flags(c.y) >> 1) | (comma_flags(c.z) >> 2) | (comma_flags(c.w) >> 3);
}

// The (at most two) commas of word k: byte 0, then the single comma bytes 1..3 can hold.
template <int k>
__device__ __forceinline__ void emit_word(uint32_t &acc, uint32_t &sp, uint32_t lut_s, uint32_t prev, uint32_t word,
                                          uint32_t zz, uint32_t c_four, uint32_t c_shift) {
    emit_token(acc, sp, lut_s, __funnelshift_r(prev, word, 4), zz & (0x80u >> k), c_four, c_shift);
    const uint32_t m = zz & (0x80808000u >> k);
    emit_token(acc, sp, lut_s, __funnelshift_r(prev, word, find_msb(m) + (uint32_t)(k - 3)), m, c_four, c_shift);
}

__device__ __forceinline__ uint32_t emit_chunk(uint32_t &acc, uint32_t sp, uint32_t lut_s, uint32_t wp, const uint4 &c,
                                               uint32_t zz, uint32_t c_four, uint32_t c_shift) {
    const uint32_t sp0 = sp;
    emit_word<0>(acc, sp, lut_s, wp, c.x, zz, c_four, c_shift);
    emit_word<1>(acc, sp, lut_s, c.x, c.y, zz, c_four, c_shift);
    emit_word<2>(acc, sp, lut_s, c.y, c.z, zz, c_four, c_shift);
    emit_word<3>(acc, sp, lut_s, c.z, c.w, zz, c_four, c_shift);
    return sp - sp0;
}

// The same for a piece of kW <= 8 words (w[0..kW-1], wp the word in front): zz holds the comma flags of word k in
// bit 7-k of every byte (piece_commas_exact).
template <int kW>
__device__ __forceinline__ uint32_t emit_piece(uint32_t &acc, uint32_t sp, uint32_t lut_s, uint32_t wp,
                                               const uint32_t (&w)[kW], uint32_t zz, uint32_t c_four, uint32_t c_shift) {
    static_assert(kW == 4 || kW == 6, "piece sizes of the unit encoder");
    const uint32_t sp0 = sp;
    emit_word<0>(acc, sp, lut_s, wp, w[0], zz, c_four, c_shift);
    emit_word<1>(acc, sp, lut_s, w[0], w[1], zz, c_four, c_shift);
    emit_word<2>(acc, sp, lut_s, w[1], w[2], zz, c_four, c_shift);
    emit_word<3>(acc, sp, lut_s, w[2], w[3], zz, c_four, c_shift);
    if constexpr (kW == 6) {
        emit_word<4>(acc, sp, lut_s, w[3], w[4], zz, c_four, c_shift);
        emit_word<5>(acc, sp, lut_s, w[4], w[5], zz, c_four, c_shift);
    }
    return sp - sp0;
}

}  // namespace

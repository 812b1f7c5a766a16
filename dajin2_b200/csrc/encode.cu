// MIDSV text -> read x locus code matrix (dj_encode).
//
// One warp streams one read.  Every trip of the main loop the 32 lanes load two runs of 32 consecutive,
// 16-byte aligned chunks of the row (two coalesced 512-byte requests) and work on the COMMAS of their chunks: a
// comma closes exactly one token, the rank of the comma inside the row is the token's locus, and the four bytes
// in front of it identify the token's anchor ("=X", "-X" or "*XY") and whether the token is an insertion
// ("+B|+B|...|<anchor>": the byte in front of the anchor is '|').  So a token of any length is translated with
// ONE perfect-hash look-up keyed on the bytes that precede its closing comma -- there is no per-token loop, no
// slow path for insertions and no divergence: a 32-bit word holds at most two commas (tokens are >= 2
// characters), one at byte 0 and one among bytes 1..3, which gives eight predicated slots per lane and chunk.
// Code bytes are staged in shared memory and leave the SM as aligned 16-byte stores of complete 32-byte sectors.
// HBM traffic is the text once in and the code bytes once out (algorithmic bytes per read:
// row_len + n_loci + 16).
//
// The kernel is bound by the integer (ALU) pipe and the shared-memory pipe, not by HBM (profiles/): the slot is
// therefore written in PTX and kept to one 4-byte table read, one 1-byte staging store and integer multiply-adds
// (FMA pipe) for the hash and the address.
//
// calc_match (src/DAJIN2/core/classification/classifier.py:11-24) = #'+' + #'-' + #'*' + #lowercase characters
// falls out of the same pass without touching the bytes again: '-' and '*' only occur as anchor operators and
// the anchors' lowercase letters are known from the look-up (the table entry carries them in bit fields that are
// summed with one add per token), and the number of inserted bases follows from the row length, because every
// inserted base is the three bytes "+B|":
//     row_len = 3 * n_inserted + 2 * n_tokens + n_substitutions + (n_tokens - 1).
// Only reads that carry a lowercase insertion (an insertion inside an inversion) are recounted byte by byte.
//
// Validation: a token whose anchor is not in the table reads an empty slot (code 0x7F) with probability
// 1 - 320/8192 = 0.96 and the row is flagged; rows are also flagged when the token count, the length identity
// above or 7-bit ASCII fail.  (The reference does not validate its input at all.)
//
// Reference behaviour covered here: the tokenisation every pass of the reference starts with
// (`json.loads` + `MIDSV.split(",")`, src/DAJIN2/utils/fileio.py:49-52) and calc_match.
#include "dj_common.cuh"

#ifndef DJ_ENC_PREFETCH_DIST
#define DJ_ENC_PREFETCH_DIST 1  // trips (of 1 KB per warp) the text is prefetched ahead of its use; 0 = off
#endif
#ifndef DJ_ENC_PREFETCH_LEVEL
#define DJ_ENC_PREFETCH_LEVEL "L1"
#endif
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

#ifndef DJ_ENC_WARPS
#define DJ_ENC_WARPS 20  // warps per CTA; two CTAs per SM = 40 warps at 48 registers (16: 63 registers, 1 % slower)
#endif
constexpr int kWarpsPerCta = DJ_ENC_WARPS;
constexpr int kLutBits = 13;
constexpr int kLutSize = 1 << kLutBits;
// (key * kLutMagic) >> 19 is injective over the 320 keys enumerated in build_token_lut and spreads the frequent
// ones over distinct shared-memory banks (tests/test_host_logic.py re-derives both properties)
constexpr uint32_t kLutMagic = 0xda7e7ab5u;
constexpr uint32_t kCommas = 0x2C2C2C2Cu;
constexpr uint32_t kFill = 0x20202020u;  // bytes outside the row read as ' ' (never a comma, high nibble 2)
constexpr int kStageBytes = 512;         // per warp: < 32 carried bytes + two chunks of at most kMaxTokensPerChunk
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

__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
encode_kernel(const uint8_t *__restrict__ text, const int64_t *__restrict__ row_off,
              const int32_t *__restrict__ row_len, int64_t n_reads, int n_loci, int code_pitch, int ckpt_pitch,
              uint8_t *__restrict__ codes, int32_t *__restrict__ match_score, int32_t *__restrict__ n_tokens,
              uint32_t *__restrict__ flags, uint32_t *__restrict__ ckpt, uint32_t c_four, uint32_t c_shift) {
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
        uint32_t *ck = ckpt + r * (int64_t)ckpt_pitch;

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
            // checkpoint: tokens that START before this trip's first byte (dj_find_token resumes from it)
            if (lane == 0) {
                ck[pair] = pair == 0 ? 0u : (uint32_t)tokbase + ((carry_w >> 24) == ',' ? 0u : 1u);
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
            if (total_a > kMaxTokensPerChunk || total_b > kMaxTokensPerChunk || tokbase >= code_pitch) {
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


// ---------------------------------------------------------------------------------------------------------
// Reference-assisted encoder (dj_encode_ref)
//
// 98 tokens in 100 are plain matches "=X" whose X is, by the definition of MIDSV, the reference base of their locus:
// a stretch of plain tokens is byte for byte the reference rendered as "=X,=X,...".  So when the caller supplies
// the allele sequence, a lane does not translate its tokens one by one; it VERIFIES them: the 16 bytes of its chunk
// plus the 4 bytes in front of them are compared with the rendered reference at the offset the token ranks dictate
// (20 bytes, six conflict-free shared-memory loads, five funnel shifts, five XOR/ORs).  If they are equal, every
// token that closes in the chunk is a plain match of the reference and its code is already in the row: each row
// starts as a copy of the reference's code row (coalesced 16-byte stores).  If they differ -- a mismatch, an indel,
// an inversion, an N, a row edge -- the chunk goes to the warp's queue (32 bytes in shared memory), and whenever
// 32 chunks are queued the warp translates them with the eight-slot code of the generic kernel, one chunk per
// lane, all lanes busy, and patches the row with byte stores.  Nothing is taken on trust: a read whose plain
// tokens disagree with the sequence just takes the generic path chunk by chunk.  Output identical to dj_encode.
// ---------------------------------------------------------------------------------------------------------

constexpr int kRefFrontPad = 32;   // ' ' bytes in front of the rendered reference (a row starts after the filler)
constexpr int kRefBlockWords = 41; // a block = 32 words of text + 9 words repeated from the next block: lanes read
                                   // 4 words apart, the 9-word skew puts the 32 lanes of a request in 32 banks, and
                                   // a 6-word window never leaves its block
constexpr int kQueueEntries = 64;  // per warp; an entry = 2 x uint4: (prev word, chunk), (first locus, comma flags)

__host__ __device__ inline int ref_text_blocks(int n_loci) { return (kRefFrontPad + 3 * n_loci + 64 + 127) / 128; }

// byte k of the padded rendered reference: ' ' x 32, then "=X," per locus, then ' '.  Loci whose base is not one of
// ACGTN (upper case) get a byte no text holds, so their tokens always take the generic path.
__device__ __forceinline__ uint32_t ref_text_byte(const uint8_t *__restrict__ ref, int n_loci, int k) {
    const int pos = k - kRefFrontPad;
    if (pos < 0 || pos >= 3 * n_loci) return ' ';
    const int locus = pos / 3, m = pos - 3 * locus;
    if (m == 0) return '=';
    if (m == 2) return ',';
    const uint32_t b = ref[locus];
    return (b == 'A' || b == 'C' || b == 'G' || b == 'T' || b == 'N') ? b : 0x01u;
}

__device__ __forceinline__ uint32_t ref_code(uint32_t b) {
    return b == 'A' ? 0u : b == 'C' ? 1u : b == 'G' ? 2u : b == 'T' ? 3u : b == 'N' ? 4u : 0u;
}

// the generic slot with the code byte going to global memory (queued chunks patch the row)
__device__ __forceinline__ void emit_token_g(uint32_t &acc, uint8_t *&gp, uint32_t lut_s, uint32_t key, uint32_t flag,
                                             uint32_t c_four) {
    asm volatile(
        "{\n .reg .pred p;\n .reg .u32 h, e;\n"
        " setp.ne.u32 p, %4, 0;\n"
        " mul.lo.u32 h, %2, %6;\n"
        " shr.u32 h, h, %7;\n"
        " mad.lo.u32 h, h, %5, %3;\n"
        " ld.shared.u32 e, [h];\n"
        " @p add.u32 %0, %0, e;\n"
        " @p st.global.u8 [%1], e;\n"
        " @p add.u64 %1, %1, 1;\n}\n"
        : "+r"(acc), "+l"(gp)
        : "r"(key), "r"(lut_s), "r"(flag), "r"(c_four), "n"(kLutMagic), "n"(32 - kLutBits)
        : "memory");
}

template <int k>
__device__ __forceinline__ void emit_word_g(uint32_t &acc, uint8_t *&gp, uint32_t lut_s, uint32_t prev, uint32_t word,
                                            uint32_t zz, uint32_t c_four) {
    emit_token_g(acc, gp, lut_s, __funnelshift_r(prev, word, 4), zz & (0x80u >> k), c_four);
    const uint32_t m = zz & (0x80808000u >> k);
    emit_token_g(acc, gp, lut_s, __funnelshift_r(prev, word, find_msb(m) + (uint32_t)(k - 3)), m, c_four);
}

__global__ void __launch_bounds__(kWarpsPerCta * 32, 2)
encode_ref_kernel(const uint8_t *__restrict__ text, const int64_t *__restrict__ row_off,
                  const int32_t *__restrict__ row_len, int64_t n_reads, int n_loci, int code_pitch, int ckpt_pitch,
                  const uint8_t *__restrict__ ref, uint8_t *__restrict__ codes, int32_t *__restrict__ match_score,
                  int32_t *__restrict__ n_tokens, uint32_t *__restrict__ flags, uint32_t *__restrict__ ckpt,
                  uint32_t c_four) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    // layout: token table | per-warp queues | code row of the reference | rendered reference (swizzled)
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem_raw);
    uint4 *queue_all = reinterpret_cast<uint4 *>(smem_raw + kLutSize * 4);
    uint8_t *ref_codes = smem_raw + kLutSize * 4 + kWarpsPerCta * kQueueEntries * 32;
    uint32_t *ref_text = reinterpret_cast<uint32_t *>(ref_codes + code_pitch);
    build_token_lut(lut);
    for (int i = threadIdx.x; i < code_pitch; i += blockDim.x) ref_codes[i] = i < n_loci ? (uint8_t)ref_code(ref[i]) : 0;
    const int n_blocks = ref_text_blocks(n_loci);
    for (int w = threadIdx.x; w < n_blocks * kRefBlockWords; w += blockDim.x) {
        const int blk = w / kRefBlockWords, in = w - blk * kRefBlockWords;
        const int k = (blk * 32 + in) * 4;  // words 32..40 of a block repeat the head of the next one
        ref_text[w] = ref_text_byte(ref, n_loci, k) | (ref_text_byte(ref, n_loci, k + 1) << 8) |
                      (ref_text_byte(ref, n_loci, k + 2) << 16) | (ref_text_byte(ref, n_loci, k + 3) << 24);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    uint4 *queue = queue_all + (threadIdx.x >> 5) * (kQueueEntries * 2);
    const uint32_t lut_s = (uint32_t)__cvta_generic_to_shared(lut);
    const uint32_t ref_text_s = (uint32_t)__cvta_generic_to_shared(ref_text);
    const uint32_t lanes_below = (1u << lane) - 1u;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * kWarpsPerCta;

    for (int64_t r = warp; r < n_reads; r += n_warps) {
        const int64_t off = row_off[r];
        const int len = row_len[r];
        const int head = (int)(off & 15);
        const uint4 *base = reinterpret_cast<const uint4 *>(text + (off - head));
        const int nbytes = head + len;
        const int n_iter = (nbytes >> 9) + 1;
        const int n_pairs = (n_iter + 1) >> 1;
        uint8_t *out = codes + r * (int64_t)code_pitch;
        uint32_t *ck = ckpt + r * (int64_t)ckpt_pitch;

        // the row starts as the reference's code row
        for (int b = lane * 16; b < code_pitch; b += 512)
            *reinterpret_cast<uint4 *>(out + b) = *reinterpret_cast<const uint4 *>(ref_codes + b);
        __syncwarp();

        int tokbase = 0;
        uint32_t counts = 0, sticky = 0, high = 0;
        uint32_t carry_w = kFill;
        uint32_t q_in = 0, q_out = 0;  // chunks queued / translated so far (warp-uniform)
        int pair = 0;

        // translate (up to) 32 queued chunks, one per lane, with the generic slots; codes go to the row
        auto drain = [&](uint32_t n) {
            __syncwarp();
            if ((uint32_t)lane < n) {
                const uint32_t slot = (q_out + lane) & (kQueueEntries - 1);
                const uint4 e0 = queue[slot * 2], e1 = queue[slot * 2 + 1];
                const uint32_t first = e1.y, zz = e1.z;
                const uint32_t cnt = __popc(zz);
                if (first + cnt <= (uint32_t)code_pitch) {
                    uint8_t *gp = out + first;
                    uint8_t *const gp0 = gp;
                    uint32_t acc = 0;
                    emit_word_g<0>(acc, gp, lut_s, e0.x, e0.y, zz, c_four);
                    emit_word_g<1>(acc, gp, lut_s, e0.y, e0.z, zz, c_four);
                    emit_word_g<2>(acc, gp, lut_s, e0.z, e0.w, zz, c_four);
                    emit_word_g<3>(acc, gp, lut_s, e0.w, e1.x, zz, c_four);
                    sticky |= ((uint32_t)(gp - gp0) ^ cnt) | ((acc >> 15) & 0x1F00u);
                    counts += ((acc >> kStatShift) & 0x3Fu) | ((acc >> 2) & 0x1F0000u);
                } else {
                    sticky |= kStickyBroken;
                }
            }
            q_out += n;
            __syncwarp();
        };

        // One chunk: verify it against the rendered reference or queue it.  `first` = rank of the first token that
        // closes in the chunk, `wp` = the word in front of the chunk.
        auto check = [&](const uint4 &c, uint32_t wp, uint32_t zz, uint32_t cnt, uint32_t first) {
            const uint32_t t = zz & 0x00808080u;     // a plain chunk closes a token within its first three bytes
            const uint32_t p8 = t ? find_msb(t) - 7u : 0u;  // 8 x (position of that comma)
            // byte offset, in the padded rendered reference, of the byte 4 in front of the chunk (ranks past the
            // reference are clamped: such a chunk is never plain, the loads just have to stay inside the table)
            const uint32_t a = 3u * min(first, (uint32_t)n_loci) + (uint32_t)(kRefFrontPad + 2 - 4) - (p8 >> 3);
            const uint32_t q = a >> 2, sh = (a & 3u) << 3;
            const uint32_t addr = ref_text_s + ((q + 9u * (q >> 5)) << 2);
            uint32_t r0, r1, r2, r3, r4, r5;
            asm volatile("ld.shared.u32 %0, [%6];\n ld.shared.u32 %1, [%6+4];\n ld.shared.u32 %2, [%6+8];\n"
                         " ld.shared.u32 %3, [%6+12];\n ld.shared.u32 %4, [%6+16];\n ld.shared.u32 %5, [%6+20];\n"
                         : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5) : "r"(addr));
            uint32_t d = (wp ^ __funnelshift_r(r0, r1, sh)) & (0xFFFFFFFFu << ((p8 + 8u) & 31u));
            d |= c.x ^ __funnelshift_r(r1, r2, sh);
            d |= c.y ^ __funnelshift_r(r2, r3, sh);
            d |= c.z ^ __funnelshift_r(r3, r4, sh);
            d |= c.w ^ __funnelshift_r(r4, r5, sh);
            const bool plain = d == 0 && t != 0 && first < (uint32_t)n_loci;
            const bool push = !plain && cnt != 0;
            const uint32_t vote = __ballot_sync(0xffffffffu, push);
            if (push) {
                const uint32_t slot = (q_in + __popc(vote & lanes_below)) & (kQueueEntries - 1);
                queue[slot * 2] = make_uint4(wp, c.x, c.y, c.z);
                queue[slot * 2 + 1] = make_uint4(c.w, first, zz, 0u);
            }
            q_in += __popc(vote);
            if (q_in - q_out >= 32) drain(32);
        };

        auto trip = [&](const uint4 &ca, const uint4 &cb) {
            const uint32_t zza = chunk_commas(ca, high), zzb = chunk_commas(cb, high);
            uint32_t wpa = __shfl_up_sync(0xffffffffu, ca.w, 1);
            uint32_t wpb = __shfl_up_sync(0xffffffffu, cb.w, 1);
            const uint32_t last_a = __shfl_sync(0xffffffffu, ca.w, 31);
            if (lane == 0) {
                ck[pair] = pair == 0 ? 0u : (uint32_t)tokbase + ((carry_w >> 24) == ',' ? 0u : 1u);
                wpa = carry_w;
                wpb = last_a;
            }
            carry_w = __shfl_sync(0xffffffffu, cb.w, 31);

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
            if (total_a > kMaxTokensPerChunk || total_b > kMaxTokensPerChunk || tokbase >= code_pitch) {
                sticky |= kStickyBroken;
                pair = 0x3FFFFFFF;
                return;
            }
            check(ca, wpa, zza, cnt_a, (uint32_t)tokbase + (excl & 0xFFFFu));
            check(cb, wpb, zzb, cnt_b, (uint32_t)tokbase + (uint32_t)total_a + (excl >> 16));
            tokbase += total_a + total_b;
        };

        trip(load_chunk(base, lane, head, nbytes), load_chunk(base, 32 + lane, head, nbytes));
        const int n_inside = nbytes >> 10;
#pragma unroll 1
        for (++pair; pair < n_inside; ++pair) {
            const uint4 *p = base + pair * 64 + lane;
#if DJ_ENC_PREFETCH_DIST > 0
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
        drain(q_in - q_out);

        const bool broken = (__reduce_or_sync(0xffffffffu, sticky) & kStickyBroken) != 0;
        counts = __reduce_add_sync(0xffffffffu, counts);
        sticky = __reduce_or_sync(0xffffffffu, sticky | (high & 0x80808080u));
        const uint32_t mism = counts & 0xFFFFu, n_sub = counts >> 16;
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

extern "C" int32_t dj_code_pitch(int32_t n_loci) { return (n_loci + 31) & ~31; }

extern "C" int32_t dj_ckpt_pitch(int32_t max_row_len) { return (max_row_len + 15) / DJ_ENC_CHUNK + 2; }

extern "C" int dj_encode(const uint8_t *text, const int64_t *row_off, const int32_t *row_len, int64_t n_reads,
                         int32_t n_loci, int32_t code_pitch, int32_t ckpt_pitch, uint8_t *codes,
                         int32_t *match_score, int32_t *n_tokens, uint32_t *flags, uint32_t *ckpt, void *stream) {
    if (n_reads <= 0) return 0;
    if (((uintptr_t)text & 15) != 0 || ((uintptr_t)codes & 15) != 0) {
        dj_set_error("dj_encode: text and codes must be 16-byte aligned");
        return 1;
    }
    if (n_loci > 400000) {  // the per-row counters are reduced as 16-bit fields: 3 * n_loci / 32 must fit
        dj_set_error("dj_encode: n_loci %d is beyond the supported 400000", n_loci);
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
        text, row_off, row_len, n_reads, n_loci, code_pitch, ckpt_pitch, codes, match_score, n_tokens, flags, ckpt, 4u,
        (uint32_t)kLutSize);
    DJ_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t dj_encode_ref_smem_bytes(int32_t n_loci, int32_t code_pitch) {
    return (int64_t)kLutSize * 4 + (int64_t)kWarpsPerCta * kQueueEntries * 32 + code_pitch +
           (int64_t)ref_text_blocks(n_loci) * kRefBlockWords * 4;
}

extern "C" int dj_encode_ref(const uint8_t *text, const int64_t *row_off, const int32_t *row_len, int64_t n_reads,
                             int32_t n_loci, int32_t code_pitch, int32_t ckpt_pitch, const uint8_t *ref_seq,
                             uint8_t *codes, int32_t *match_score, int32_t *n_tokens, uint32_t *flags, uint32_t *ckpt,
                             void *stream) {
    const int64_t smem = dj_encode_ref_smem_bytes(n_loci, code_pitch);
    // without a reference, or when the rendered reference does not fit next to a second CTA: the generic kernel
    if (ref_seq == nullptr || smem > 110 * 1024)
        return dj_encode(text, row_off, row_len, n_reads, n_loci, code_pitch, ckpt_pitch, codes, match_score, n_tokens,
                         flags, ckpt, stream);
    if (n_reads <= 0) return 0;
    if (((uintptr_t)text & 15) != 0 || ((uintptr_t)codes & 15) != 0) {
        dj_set_error("dj_encode_ref: text and codes must be 16-byte aligned");
        return 1;
    }
    if (code_pitch < n_loci || (code_pitch & 31) != 0) {
        dj_set_error("dj_encode_ref: code_pitch %d must be a multiple of 32 and >= n_loci %d", code_pitch, n_loci);
        return 1;
    }
    const int sms = dj_sm_count();
    if (sms <= 0) return 1;
    static int64_t configured = 0;
    if (configured < smem) {
        DJ_CUDA_CHECK(cudaFuncSetAttribute(encode_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int64_t ctas = (n_reads + kWarpsPerCta - 1) / kWarpsPerCta;
    const int64_t resident = (int64_t)sms * 2;
    if (ctas > resident) ctas = resident;
    encode_ref_kernel<<<(unsigned)ctas, kWarpsPerCta * 32, (size_t)smem, (cudaStream_t)stream>>>(
        text, row_off, row_len, n_reads, n_loci, code_pitch, ckpt_pitch, ref_seq, codes, match_score, n_tokens, flags,
        ckpt, 4u);
    DJ_LAUNCH_CHECK();
    return 0;
}

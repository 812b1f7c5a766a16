// MIDSV text -> read x locus code matrix (dj_encode).
//
// One warp streams one read: every iteration the 32 lanes load 32 consecutive, 16-byte aligned chunks of the
// row (one coalesced 512-byte request), find the token starts of their chunk with byte-parallel arithmetic,
// rank them with a warp scan and translate each token through a perfect-hash table held in shared memory.
// Nothing is staged and nothing is synchronised across warps; the only HBM traffic is the text once in and the
// code bytes once out (algorithmic bytes per read: row_len + n_loci).
//
// Reference behaviour covered here: the tokenisation every pass of the reference starts with
// (`json.loads` + `MIDSV.split(",")`, src/DAJIN2/utils/fileio.py:49-52) and calc_match
// (src/DAJIN2/core/classification/classifier.py:11-24).
#include "dj_common.cuh"

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kLutBits = 11;
constexpr int kLutSize = 1 << kLutBits;
constexpr uint32_t kLutMagic = 0xc324c985u;  // collision-free over the 80 token heads below (tests/test_codes.py)
constexpr uint32_t kEntIns = 0x4000u;
constexpr uint32_t kEntBad = 0x8000u;
constexpr uint32_t kCommas = 0x2C2C2C2Cu;

__device__ __forceinline__ uint32_t lut_slot(uint32_t head3) { return (head3 * kLutMagic) >> (32 - kLutBits); }

// Table entry: bits 0..7 code, bits 8..9 calc_match contribution, bit 14 insertion head, bit 15 unsupported.
__device__ void build_token_lut(uint16_t *lut) {
    for (int e = threadIdx.x; e < kLutSize; e += blockDim.x) lut[e] = (uint16_t)(kEntBad | DJ_CODE_INVALID);
    __syncthreads();
    const int k = threadIdx.x;
    if (k < 80) {
        uint32_t c0, c1, c2, entry;
        if (k < 20) {  // "=B," and "-B,"
            const int j = k % 10;
            c0 = k < 10 ? '=' : '-';
            c1 = (uint8_t)dj_base_char(j % 5, j >= 5);
            c2 = ',';
            entry = (uint32_t)k | ((uint32_t)dj_anchor_mismatch(k) << 8);
        } else if (k < 70) {  // "*RB"
            const int j = k - 20;
            const bool lower = j >= 25;
            c0 = '*';
            c1 = (uint8_t)dj_base_char((j % 25) / 5, lower);
            c2 = (uint8_t)dj_base_char(j % 5, lower);
            entry = (uint32_t)k | ((uint32_t)dj_anchor_mismatch(k) << 8);
        } else {  // "+B|" : head of an insertion token
            const int j = k - 70;
            c0 = '+';
            c1 = (uint8_t)dj_base_char(j % 5, j >= 5);
            c2 = '|';
            entry = kEntIns | DJ_CODE_INVALID;
        }
        lut[lut_slot(c0 | (c1 << 8) | (c2 << 16))] = (uint16_t)entry;
    }
    __syncthreads();
}

// 0x80 in every byte of w that equals ',' (exact for any byte value)
__device__ __forceinline__ uint32_t comma_flags(uint32_t w) {
    const uint32_t x = w ^ kCommas;
    const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | x) & 0x80808080u;
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

__device__ __forceinline__ bool is_base_letter(uint32_t c) {
    const uint32_t u = c & 0xDFu;
    return u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'N';
}

// Slow path: "+B|+B|...|<anchor>" starting at byte p of the row.  Returns the table-style entry.
__device__ __noinline__ uint32_t parse_insertion(const uint8_t *row, int len, int p, const uint16_t *lut) {
    int q = p, n_ins = 0, n_lower = 0, n_upper = 0;
    bool bad = false;
    while (q < len && row[q] == '+') {
        if (q + 2 >= len || row[q + 2] != '|' || !is_base_letter(row[q + 1])) { bad = true; break; }
        if (row[q + 1] & 0x20) ++n_lower; else ++n_upper;
        ++n_ins;
        q += 3;
    }
    int alen = 0;
    while (q + alen < len && row[q + alen] != ',') ++alen;
    if (bad || n_ins == 0 || alen < 2 || alen > 3) return kEntBad | DJ_CODE_INVALID;
    const uint32_t head3 = row[q] | ((uint32_t)row[q + 1] << 8) | ((uint32_t)(alen == 3 ? row[q + 2] : ',') << 16);
    const uint32_t e = lut[lut_slot(head3)];
    if (e & (kEntBad | kEntIns)) return kEntBad | DJ_CODE_INVALID;
    const int anchor = e & 0x7F;
    // the inserted bases must share the anchor's case (an all-lowercase token is an inversion; mixed is unsupported)
    if (dj_anchor_is_lower(anchor) ? n_upper > 0 : n_lower > 0) return kEntBad | DJ_CODE_INVALID;
    // calc_match: one '+' per inserted base, every lowercase letter, plus the anchor's own contribution
    const uint32_t mismatch = (uint32_t)(n_ins + n_lower) + ((e >> 8) & 3u);
    return (uint32_t)(DJ_CODE_INS | anchor) | (mismatch << 16);
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
encode_kernel(const uint8_t *__restrict__ text, const int64_t *__restrict__ row_off,
              const int32_t *__restrict__ row_len, int64_t n_reads, int n_loci, int code_pitch, int ckpt_pitch,
              uint8_t *__restrict__ codes, int32_t *__restrict__ match_score, int32_t *__restrict__ n_tokens,
              uint32_t *__restrict__ flags, uint32_t *__restrict__ ckpt) {
    __shared__ uint16_t lut[kLutSize];
    build_token_lut(lut);

    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * kWarpsPerCta;

    for (int64_t r = warp; r < n_reads; r += n_warps) {
        const int64_t off = row_off[r];
        const int len = row_len[r];
        const uint8_t *row = text + off;
        const int head = (int)(off & 15);
        const uint4 *base = reinterpret_cast<const uint4 *>(text + (off - head));
        const int nbytes = head + len;
        const int n_iter = (nbytes + DJ_ENC_CHUNK - 1) / DJ_ENC_CHUNK;
        uint8_t *out = codes + r * (int64_t)code_pitch;
        uint32_t *ck = ckpt + r * (int64_t)ckpt_pitch;

        int tokbase = 0;
        uint32_t mismatch = 0, bad = 0;
        // comma flag of the byte just before this iteration's first byte: the row start counts as one
        uint32_t carry = head == 0 ? 0x80000000u : 0u;
        uint4 cur = load_chunk(base, lane, head, nbytes);

        for (int it = 0; it < n_iter; ++it) {
            const uint4 nxt = load_chunk(base, (it + 1) * 32 + lane, head, nbytes);
            uint32_t w[5];
            w[0] = cur.x; w[1] = cur.y; w[2] = cur.z; w[3] = cur.w;
            // look-ahead word: first word of the next lane's chunk (lane 31 takes it from the next iteration)
            w[4] = __shfl_down_sync(0xffffffffu, w[0], 1);
            const uint32_t nxt0 = __shfl_sync(0xffffffffu, nxt.x, 0);
            if (lane == 31) w[4] = nxt0;

            uint32_t z[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) z[k] = comma_flags(w[k]);
            uint32_t zprev = __shfl_up_sync(0xffffffffu, z[3], 1);
            if (lane == 0) zprev = carry;
            carry = __shfl_sync(0xffffffffu, z[3], 31);

            // start flags: bit 7 of byte j is set when byte j-1 is a comma
            uint32_t s[4];
            s[0] = __funnelshift_l(zprev, z[0], 8);
            s[1] = __funnelshift_l(z[0], z[1], 8);
            s[2] = __funnelshift_l(z[1], z[2], 8);
            s[3] = __funnelshift_l(z[2], z[3], 8);
            const int g0 = (it * 32 + lane) * 16;
            if (g0 < head || g0 + 16 > nbytes) {  // only the first and last chunks of a row
#pragma unroll
                for (int k = 0; k < 4; ++k) s[k] &= valid_bytes(g0 + 4 * k, head, nbytes);
            }

            const int count = __popc(s[0] | (s[1] >> 1) | (s[2] >> 2) | (s[3] >> 3));
            int incl = count;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0) ck[it] = (uint32_t)tokbase;
            int tok = tokbase + incl - count;

#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t m = s[k];
                while (m) {
                    const int sh = __ffs(m) - 8;  // 0, 8, 16, 24
                    m &= m - 1;
                    const uint32_t head3 = __funnelshift_r(w[k], w[k + 1], sh) & 0x00FFFFFFu;
                    uint32_t e = lut[lut_slot(head3)];
                    if (e & kEntIns) {
                        e = parse_insertion(row, len, g0 + 4 * k + (sh >> 3) - head, lut);
                        mismatch += e >> 16;
                    } else {
                        mismatch += (e >> 8) & 3u;
                    }
                    bad |= e;
                    if (tok < n_loci) out[tok] = (uint8_t)e;
                    ++tok;
                }
            }
            tokbase += total;
            cur = nxt;
        }

#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mismatch += __shfl_xor_sync(0xffffffffu, mismatch, d);
            bad |= __shfl_xor_sync(0xffffffffu, bad, d);
        }
        if (lane == 0) {
            match_score[r] = -(int32_t)mismatch;
            n_tokens[r] = tokbase;
            flags[r] = (tokbase != n_loci ? DJ_FLAG_TOKEN_COUNT : 0u) | ((bad & kEntBad) ? DJ_FLAG_BAD_TOKEN : 0u);
        }
    }
}

}  // namespace

extern "C" int32_t dj_code_pitch(int32_t n_loci) { return (n_loci + 31) & ~31; }

extern "C" int32_t dj_ckpt_pitch(int32_t max_row_len) { return (max_row_len + 15 + DJ_ENC_CHUNK - 1) / DJ_ENC_CHUNK + 1; }

extern "C" int dj_encode(const uint8_t *text, const int64_t *row_off, const int32_t *row_len, int64_t n_reads,
                         int32_t n_loci, int32_t code_pitch, int32_t ckpt_pitch, uint8_t *codes,
                         int32_t *match_score, int32_t *n_tokens, uint32_t *flags, uint32_t *ckpt, void *stream) {
    if (n_reads <= 0) return 0;
    if (((uintptr_t)text & 15) != 0) {
        dj_set_error("dj_encode: text must be 16-byte aligned");
        return 1;
    }
    if (code_pitch < n_loci) {
        dj_set_error("dj_encode: code_pitch %d < n_loci %d", code_pitch, n_loci);
        return 1;
    }
    const int sms = dj_sm_count();
    if (sms <= 0) return 1;
    int64_t ctas = (n_reads + kWarpsPerCta - 1) / kWarpsPerCta;
    const int64_t resident = (int64_t)sms * 8;  // 8 CTAs of 256 threads = 64 warps per SM
    if (ctas > resident) ctas = resident;
    encode_kernel<<<(unsigned)ctas, kWarpsPerCta * 32, 0, (cudaStream_t)stream>>>(
        text, row_off, row_len, n_reads, n_loci, code_pitch, ckpt_pitch, codes, match_score, n_tokens, flags, ckpt);
    DJ_LAUNCH_CHECK();
    return 0;
}

// Byte-parallel arithmetic of the unit encoder (encode.cu), written so that the same functions compile for the
// host: scripts/encode_unit_host_test.cpp runs them against a plain tokenizer without a GPU.
//
// A UNIT is 48 consecutive bytes of a row (twelve 32-bit words, one lane of a warp).  48 is a multiple of 3 and of
// 16, so in a stretch of plain tokens ("=X," or "-X,", three bytes each) every unit of the stretch sees the same
// phase, holds exactly sixteen closing commas and therefore sixteen tokens.
//
//   comma bitmap   every byte of the unit is compared with ',' (exact for any byte value); the 48 flags are packed
//                  into two words: bit 8*b + (k & 7) of F[k >> 3] <=> byte b of word k is a comma.  popc = tokens the
//                  unit closes, for ANY unit; equality with the periodic pattern of phase c0 <=> the commas sit at
//                  bytes c0, c0+3, ..., c0+45 and nowhere else.
//   alignment      the byte stream is shifted so that a window starts two bytes in front of the first comma: the
//                  twelve aligned words are four GROUPS "=X,=" "X,=X" ",=X,".
//   group          two byte-permutes gather the four bases, two the four operators; a base is translated through a
//                  byte-permute look-up keyed on bits 1..3 of the letter (A C G T N -> 0 1 3 2 7) and verified by
//                  translating the code back to the letter; the operator must be '=' or '-' ('-' adds 10 to the code).
//   unit is PLAIN  <=> bitmap == pattern, every operator / base verified, and the byte in front of the window is ','
//                  (the first token is not the anchor of an insertion that began in the previous unit).  Its sixteen code bytes are exactly what
//                  the generic comma-slot path (encode_common.cuh) produces for the same text.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define DJU_HD __host__ __device__ __forceinline__
#else
#define DJU_HD static inline
#endif

namespace dju {

constexpr int kUnitBytes = 48;
constexpr int kUnitWords = 12;
constexpr int kUnitsPerTrip = 32;
constexpr int kTripBytes = kUnitBytes * kUnitsPerTrip;  // 1536 = DJ_ENC_CHUNK

DJU_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, s);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t n = (s >> (4 * i)) & 0xFu;
        uint32_t byte = (uint32_t)(v >> (8 * (n & 7u))) & 0xFFu;
        if (n & 8u) byte = (byte & 0x80u) ? 0xFFu : 0u;
        r |= byte << (8 * i);
    }
    return r;
#endif
}

// low word of (hi:lo) >> s, s clamped to 32
DJU_HD uint32_t shr_pair(uint32_t lo, uint32_t hi, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_rc(lo, hi, s);
#else
    if (s >= 32) return hi;
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> s);
#endif
}

// high word of (hi:lo) << (s & 31)
DJU_HD uint32_t shl_pair(uint32_t lo, uint32_t hi, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, s);
#else
    s &= 31u;
    return (uint32_t)(((((uint64_t)hi << 32) | lo) << s) >> 32);
#endif
}

DJU_HD uint32_t popc(uint32_t x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}

// 0x80 in every byte of w that equals ',' -- exact for every byte value
DJU_HD uint32_t comma_flags_exact(uint32_t w) {
    const uint32_t y = ((w ^ 0x2C2C2C2Cu) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(y | w) & 0x80808080u;
}

// acc + (f >> n): the flags of one word land in their slot of the bitmap
DJU_HD uint32_t add_shifted(uint32_t acc, uint32_t f, int n) { return acc + (f >> n); }

struct UnitMap {
    uint32_t f0, f1;  // comma bitmap (see the header comment)
};

DJU_HD UnitMap comma_bitmap(const uint32_t *w) {
    UnitMap m;
    m.f0 = comma_flags_exact(w[7]);
    m.f1 = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 7; ++k) m.f0 = add_shifted(m.f0, comma_flags_exact(w[k]), 7 - k);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 8; k < 12; ++k) m.f1 = add_shifted(m.f1, comma_flags_exact(w[k]), 15 - k);
    return m;
}

// bit of the bitmap that stands for byte p (0..47) of the unit: word index in .x, mask in .y
DJU_HD uint32_t bitmap_word(int p) { return (uint32_t)(p >> 5); }
DJU_HD uint32_t bitmap_mask(int p) { return 1u << (8 * (p & 3) + ((p >> 2) & 7)); }

// the bitmap of a plain unit whose first comma is byte c0 (0..2)
DJU_HD UnitMap plain_pattern(int c0) {
    UnitMap m;
    m.f0 = 0;
    m.f1 = 0;
    for (int p = c0; p < kUnitBytes; p += 3) {
        if (bitmap_word(p)) m.f1 |= bitmap_mask(p); else m.f0 |= bitmap_mask(p);
    }
    return m;
}

// index (0..3) into the pattern table from the comma flags of bytes 0..2 of the unit: one comma at byte c0 -> c0;
// anything else -> an index whose pattern cannot match the bitmap (3 holds an impossible pattern)
DJU_HD uint32_t phase_index(uint32_t f0) {
    const uint32_t m = f0 & 0x00010101u;
    return ((m >> 8) | (m >> 15)) & 3u;
}

// bits the byte stream is shifted right by so that the window starts two bytes in front of the first comma
DJU_HD uint32_t phase_shift_bits(int c0) { return 8u * (uint32_t)(c0 + 2); }

struct GroupOut {
    uint32_t codes;  // four code bytes (valid when the unit turns out plain)
    uint32_t bad;    // != 0: an operator is not '=' / '-', or a base is not one of A C G T N
    uint32_t dels;   // 0x10 in every byte whose operator is '-'
};

// one aligned group: v0 = "=X,=", v1 = "X,=X", v2 = ",=X,"  (commas are vouched for by the bitmap)
DJU_HD GroupOut decode_group(uint32_t v0, uint32_t v1, uint32_t v2) {
    GroupOut g;
    const uint32_t x = prmt(prmt(v0, v1, 0x0741u), v2, 0x6210u);   // bases: v0.b1 v1.b0 v1.b3 v2.b2
    const uint32_t op = prmt(prmt(v0, v1, 0x0630u), v2, 0x5210u);  // operators: v0.b0 v0.b3 v1.b2 v2.b1
    const uint32_t idx = (x >> 1) & 0x07070707u;                   // A 0, C 1, T 2, G 3, N 7
    const uint32_t nib = idx + (idx >> 4);                         // byte 0 = i0 | i1 << 4, byte 2 = i2 | i3 << 4
    const uint32_t sel = prmt(nib, 0u, 0x4420u);
    const uint32_t code = prmt(0x02030100u, 0x04FFFFFFu, sel);
    const uint32_t back = prmt(0x47544341u, 0x4E000000u, sel);     // "ACTG", 'N' at 7
    g.dels = ~op & 0x10101010u;                                    // '-' = 0x2D, '=' = 0x3D
    g.bad = ((op | 0x10101010u) ^ 0x3D3D3D3Du) | (back ^ x);
    g.codes = code + ((g.dels * 10u) >> 4);                        // "-X" = 10 + base
    return g;
}

struct UnitOut {
    uint32_t c[4];   // sixteen code bytes
    uint32_t bad;    // != 0: not plain
    uint32_t n_del;  // deletions among the sixteen tokens (calc_match counts one per deletion)
};

// w[-1] is the word in front of the unit (prev); shift = phase_shift_bits(c0)
DJU_HD UnitOut decode_unit(uint32_t prev, const uint32_t *w, uint32_t shift) {
    uint32_t v[12];
    v[0] = shr_pair(prev, w[0], shift);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = 1; j < 12; ++j) v[j] = shr_pair(w[j - 1], w[j], shift);
    UnitOut u;
    // the byte in front of the window separates the first token from its predecessor: ',' or the token is an
    // insertion whose anchor alone lies in the window ('|'), or the first token of the row (filler)
    u.bad = ((prev >> (shift - 8u)) & 0xFFu) ^ 0x2Cu;
    uint32_t dsum = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int g = 0; g < 4; ++g) {
        const GroupOut o = decode_group(v[3 * g], v[3 * g + 1], v[3 * g + 2]);
        u.c[g] = o.codes;
        u.bad |= o.bad;
        dsum += o.dels << g;  // 0x10, 0x20, 0x40, 0x80: one bit per (group, byte)
    }
    u.n_del = popc(dsum);
    return u;
}

// sixteen code bytes shifted to byte offset a (0..3) of a word-aligned destination: five words, the first holds
// zeros in its low a bytes, the last only its low a bytes
DJU_HD void rotate_codes(const uint32_t *c, uint32_t a, uint32_t *out5) {
    const uint32_t s = 8u * a;
    out5[0] = shl_pair(0u, c[0], s);
    out5[1] = shl_pair(c[0], c[1], s);
    out5[2] = shl_pair(c[1], c[2], s);
    out5[3] = shl_pair(c[2], c[3], s);
    out5[4] = shl_pair(c[3], 0u, s);
}

}  // namespace dju

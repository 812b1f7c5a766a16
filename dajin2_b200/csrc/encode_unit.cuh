// Byte-parallel arithmetic of the unit encoder (encode.cu), written so that the same functions compile for the
// host: scripts/encode_unit_host_test.cpp runs them against a plain tokenizer without a GPU.
//
// A UNIT is 48 consecutive bytes of a row (twelve 32-bit words, one lane of a warp).  48 is a multiple of 3 and of
// 16, so in a stretch of plain tokens ("=X," or "-X,", three bytes each) every unit of the stretch sees the same
// phase, holds exactly sixteen closing commas and therefore sixteen tokens.
//
//   phase          the first comma of a plain unit is byte 0, 1 or 2 (c0); a four-entry table keyed on the comma
//                  flags of bytes 0 and 1 gives the shift that aligns the byte stream and the mask of the unit's
//                  trailing bytes (behind the last token the window closes), which must not hold a comma.
//   alignment      the byte stream is shifted so that a window starts two bytes in front of the first comma: the
//                  twelve aligned words are four GROUPS "=X,=" "X,=X" ",=X,".
//   group          two byte-permutes each gather the four bases, the four operators and the four commas; a base is
//                  translated through a byte-permute look-up keyed on its low three bits (A C G T N -> 1 3 7 4 6) and
//                  verified by translating the look-up index back to the letter; the operator must be '=' or '-' ('-'
//                  adds 10 to the code), the commas must be commas.
//   unit is PLAIN  <=> every group verified, no comma in the trailing bytes, and the byte in front of the window is
//                  ',' (the first token is not the anchor of an insertion that began in the previous unit).  Then the
//                  unit closes exactly sixteen tokens and its sixteen code bytes are exactly what the generic
//                  comma-slot path (encode_common.cuh) produces for the same text.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define DJU_HD __host__ __device__ __forceinline__
#else
#define DJU_HD static inline
#endif

#ifndef DJ_ENC_FMA
#define DJ_ENC_FMA 0  // 1: the shift-and-add steps of a group as multiply-adds with run-time multipliers (see Muls)
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

// Phase table entry for the comma flags of bytes 0 and 1 of the unit (index = [byte 0 is ','] + 2 * [byte 1 is ',']):
//   .shift   bits the byte stream is shifted right by so that the window starts two bytes in front of the first comma
//   .trail   comma flags (0x80 per byte) of the unit's last word that must be clear: the bytes behind the window
//   .bad     != 0: no plain unit starts like this
//   .n0      commas a plain stretch of this phase has among the first 16 bytes of the unit (tokens its first chunk closes)
struct Phase {
    uint32_t shift, trail, bad, n0;
};

DJU_HD Phase phase_entry(uint32_t index) {
    Phase p;
    p.n0 = index == 1 ? 6u : 5u;  // commas at bytes 0,3,..,15 / 1,..,13 / 2,..,14
    p.bad = 0;
    if (index == 0) {  // neither: the first comma can only be byte 2 (the window starts at byte 0)
        p.shift = 32; p.trail = 0;
    } else if (index == 1) {  // byte 0: the window starts two bytes in front of the unit; bytes 46, 47 trail
        p.shift = 16; p.trail = 0x80800000u;
    } else if (index == 2) {  // byte 1: byte 47 trails
        p.shift = 24; p.trail = 0x80000000u;
    } else {  // ",," is not MIDSV
        p.shift = 16; p.trail = 0; p.bad = 1;
    }
    return p;
}

DJU_HD uint32_t phase_index(uint32_t w0) {
    const uint32_t f = comma_flags_exact(w0);
    return ((f >> 7) | (f >> 14)) & 3u;
}

// Multipliers that arrive as kernel arguments (DJ_ENC_FMA = 1): with a multiplier it cannot see, ptxas keeps the
// two shift-and-add steps of a group on the multiply-add pipe (IMAD.HI) instead of turning them into LEA.HI on the
// ALU pipe, which is the pipe that bounds the kernel (profiles/: ALU 78 %, FMA 20 % busy).
struct Muls {
    uint32_t fold;  // 1 << 28: mad.hi(x, fold, x) = x + (x >> 4)
    uint32_t del;   // 10 << 28: mad.hi(d, del, c) = c + ((d * 10) >> 4)
};

// idx + (idx >> 4): byte 0 = i0 | i1 << 4, byte 2 = i2 | i3 << 4 (a multiply-add, off the ALU pipe on the device)
DJU_HD uint32_t fold_nibbles(uint32_t idx, const Muls &m) {
#if defined(__CUDA_ARCH__) && DJ_ENC_FMA
    uint32_t r;
    asm("mad.hi.u32 %0, %1, %2, %1;" : "=r"(r) : "r"(idx), "r"(m.fold));
    return r;
#else
    (void)m;
    return idx + (idx >> 4);
#endif
}

// code + ((dels * 10) >> 4): dels holds 0x10 in every byte whose operator is '-' ("-X" = 10 + base)
DJU_HD uint32_t add_deletions(uint32_t code, uint32_t dels, const Muls &m) {
#if defined(__CUDA_ARCH__) && DJ_ENC_FMA
    uint32_t r;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(dels), "r"(m.del), "r"(code));
    return r;
#else
    (void)m;
    return code + ((dels * 10u) >> 4);
#endif
}

// byte permute without the selector masking of __byte_perm (the selectors built here never set bit 3 of a nibble)
DJU_HD uint32_t prmt_raw(uint32_t a, uint32_t b, uint32_t s) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
#else
    return prmt(a, b, s);
#endif
}

struct GroupOut {
    uint32_t codes;  // four code bytes (valid when bad == 0)
    uint32_t bad;    // != 0: a comma is missing, an operator is not '=' / '-', or a base is not one of A C G T N
    uint32_t dels;   // 0x10 in every byte whose operator is '-'
};

// one aligned group: v0 = "=X,=", v1 = "X,=X", v2 = ",=X,"
DJU_HD GroupOut decode_group(uint32_t v0, uint32_t v1, uint32_t v2, const Muls &m) {
    GroupOut g;
    const uint32_t x = prmt_raw(prmt_raw(v0, v1, 0x0741u), v2, 0x6210u);   // bases: v0.b1 v1.b0 v1.b3 v2.b2
    const uint32_t op = prmt_raw(prmt_raw(v0, v1, 0x0630u), v2, 0x5210u);  // operators: v0.b0 v0.b3 v1.b2 v2.b1
    const uint32_t cm = prmt_raw(prmt_raw(v0, v1, 0x0052u), v2, 0x7410u);  // commas: v0.b2 v1.b1 v2.b0 v2.b3
    const uint32_t idx = x & 0x07070707u;                                   // A 1, C 3, T 4, N 6, G 7
    const uint32_t sel = prmt_raw(fold_nibbles(idx, m), 0u, 0x4420u);
    const uint32_t code = prmt_raw(0x01FF00FFu, 0x0204FF03u, sel);          // index -> code, 0xFF for 0 / 2 / 5
    const uint32_t back = prmt_raw(0x43004100u, 0x474E0054u, sel);          // index -> letter
    g.dels = ~op & 0x10101010u;                                             // '-' = 0x2D, '=' = 0x3D
    g.bad = (cm ^ 0x2C2C2C2Cu) | ((op | 0x10101010u) ^ 0x3D3D3D3Du) | (back ^ x);
    g.codes = add_deletions(code, g.dels, m);                               // "-X" = 10 + base
    return g;
}

struct UnitOut {
    uint32_t c[4];        // sixteen code bytes
    uint32_t bad;         // != 0: not plain
    uint32_t bad_head;    // != 0: the first two groups (eight tokens, bytes 0..21+ of the unit) are not plain either
    uint32_t n_del;       // deletions among the sixteen tokens (calc_match counts one per deletion)
    uint32_t n_del_head;  // deletions among the first ph.n0 tokens (those the unit's first 16-byte chunk closes)
};

// prev is the word in front of the unit; ph = phase_entry(phase_index(w[0]))
DJU_HD UnitOut decode_unit(uint32_t prev, const uint32_t *w, const Phase &ph, const Muls &m = Muls{1u << 28, 10u << 28}) {
    const uint32_t shift = ph.shift;
    uint32_t v[12];
    v[0] = shr_pair(prev, w[0], shift);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = 1; j < 12; ++j) v[j] = shr_pair(w[j - 1], w[j], shift);
    UnitOut u;
    // the byte in front of the window separates the first token from its predecessor: ',' -- or the token is an
    // insertion whose anchor alone lies in the window ('|'), or the first token of the row (filler)
    u.bad_head = ph.bad | (((prev >> (shift - 8u)) & 0xFFu) ^ 0x2Cu);
    u.bad = comma_flags_exact(w[11]) & ph.trail;
    uint32_t dsum = 0, dhead = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int g = 0; g < 4; ++g) {
        const GroupOut o = decode_group(v[3 * g], v[3 * g + 1], v[3 * g + 2], m);
        u.c[g] = o.codes;
        if (g < 2) u.bad_head |= o.bad; else u.bad |= o.bad;
        dsum += o.dels << g;  // 0x10, 0x20, 0x40, 0x80: one bit per (group, byte)
        if (g == 0) dhead = o.dels;
        if (g == 1) dhead += (o.dels & ((1u << (8u * (ph.n0 - 4u))) - 1u)) << 1;
    }
    u.bad |= u.bad_head;
    u.n_del = popc(dsum);
    u.n_del_head = popc(dhead);
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

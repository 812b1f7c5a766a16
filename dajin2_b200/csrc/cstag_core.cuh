// SAM cs tag (long form) -> MIDSV text, one warp per read (reference call site
// preprocess/alignment/midsv_caller.py:84-85, third-party midsv.transform: PARITY UNPINNED, the rules are the ones
// DESIGN.md section 4.5 states and the test suite holds it to).
//
// The body below is written once and compiled twice: by nvcc for sm_100a (cstag.cu: warp intrinsics, shared-memory
// stage) and by g++ for scripts/cstag_host_test.cpp, where the 32 lanes of a warp are 32 host threads and every
// collective is a barrier + exchange (DJW_* below).  The CPU suite therefore executes this very code -- lane
// arithmetic, carries across trips, the staged row writer -- against the CPU restatement without a GPU.
//
// One trip of a warp covers 512 cs bytes (one aligned 16-byte load per lane).  Every byte belongs to the operator
// character before it ('=', '*', '-', '+', '~'): each lane finds the last operator among its 16 bytes, one ballot +
// two shuffles give it the operator (and the distance to it) it starts under, the carry of lane 31 goes to the next
// trip.  A byte then knows what it emits -- "=X," / "-X," / "+X|" (3 bytes), "*R" or "Q," (2 bytes), nothing under
// '~' -- and a warp prefix sum of the lane totals places the bytes in the warp's stage.  The stage is kept congruent
// to the global address modulo 16, so all but the first and the last sector of a row leave as 16-byte stores; the
// remainder (< 16 bytes) stays in the stage for the next trip.  Introns ("~gt123ag" -> 123 x "=N") cut a trip in two
// and are written by the same run filler that pads gaps between alignments and the flanks of the row.
#pragma once

#include <stdint.h>

#ifndef INT64_MAX
#define INT64_MAX 0x7fffffffffffffffLL
#endif

#ifdef __CUDACC__
#define DJW_FN __device__ __forceinline__
#define DJW_FULL 0xffffffffu
#define DJW_SHFL(v, src) __shfl_sync(DJW_FULL, (v), (src))
#define DJW_SHFL_UP(v, d) __shfl_up_sync(DJW_FULL, (v), (d))
#define DJW_SHFL_XOR(v, d) __shfl_xor_sync(DJW_FULL, (v), (d))
#define DJW_BALLOT(p) __ballot_sync(DJW_FULL, (p))
#define DJW_SYNC() __syncwarp()
#define DJW_CLZ(x) __clz((int)(x))
DJW_FN uint4 djw_load16(const uint8_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
#else
// host emulation: provided by the including file (scripts/cstag_host_test.cpp) before this header
#endif

#define DJ_CS_TRIP 512                       // cs bytes per trip of a warp
#define DJ_CS_STAGE (16 + 3 * DJ_CS_TRIP + 16) // stage bytes per warp (multiple of 16)

#define DJ_CS_OVERLAP 1   // two alignments of the read share a reference base: dropped
#define DJ_CS_GRAMMAR 2   // not a long-form cs tag, or an insertion with nothing after it at the end of the row
#define DJ_CS_TOO_LONG 4  // the alignments reach past the end of the reference, or the row needs >= 2^31 bytes
#define DJ_CS_HAS_N 8     // informational (the row is kept): an alignment holds "=N" tokens of its own

template <bool kWrite>
struct DjCsWriter {
    uint8_t *stage;    // this warp's stage; stage[k] goes to gbase[k]
    uint8_t *gbase;    // 16-byte aligned
    int fill;          // bytes in the stage, the first head_skip of them not ours
    int head_skip;
    int64_t row_pos;   // bytes of the row produced so far
    int64_t last;      // row position of the separator that ends the row ('\n' instead of ','); -1 when measuring
    const uint8_t *seq;  // reference sequence: tokens between the pieces of a read become "-<base>" (else "=N")
    int64_t n_tokens;  // "=N" tokens written by the run filler
};

template <bool kWrite>
DJW_FN uint8_t djcs_sep(const DjCsWriter<kWrite> &w, int64_t pos) {
    return pos == w.last ? (uint8_t)'\n' : (uint8_t)',';
}

// full 16-byte sectors of the stage leave; the remainder moves to the front (or leaves byte by byte at the end)
template <bool kWrite>
DJW_FN void djcs_flush(DjCsWriter<kWrite> &w, int lane, bool final) {
    if (!kWrite) return;
    DJW_SYNC();
    const int n_vec = w.fill >> 4;
    for (int v = lane; v < n_vec; v += 32) {
        if (v == 0 && w.head_skip) {
            for (int b = w.head_skip; b < 16; ++b) w.gbase[b] = w.stage[b];
        } else {
            *reinterpret_cast<uint4 *>(w.gbase + 16 * v) = *reinterpret_cast<const uint4 *>(w.stage + 16 * v);
        }
    }
    const int rem = w.fill & 15;
    if (final) {
        const int lo = n_vec * 16 + (n_vec == 0 ? w.head_skip : 0);
        for (int b = lo + lane; b < w.fill; b += 32) w.gbase[b] = w.stage[b];
    } else if (n_vec > 0) {
        uint8_t keep = 0;
        if (lane < rem) keep = w.stage[n_vec * 16 + lane];
        DJW_SYNC();
        if (lane < rem) w.stage[lane] = keep;
        w.gbase += n_vec * 16;
        w.fill = rem;
        w.head_skip = 0;
    }
    DJW_SYNC();
}

// n tokens "=N," (lower: "=n,"); with a reference sequence and inside == true (a gap between two alignments, an
// intron) the tokens are "-<base>," of reference positions ref_pos .. ref_pos + n - 1 instead: what
// replace_internal_n_to_d (midsv_caller.py:108-126) makes of them
template <bool kWrite>
DJW_FN void djcs_fill(DjCsWriter<kWrite> &w, int lane, int64_t n_tokens, bool lower, bool inside, int64_t ref_pos) {
    const bool as_deletion = inside && w.seq != nullptr;
    if (!as_deletion) w.n_tokens += n_tokens;
    if (!kWrite) {
        w.row_pos += 3 * n_tokens;
        return;
    }
    const uint8_t base = lower ? (uint8_t)'n' : (uint8_t)'N';
    for (int64_t t0 = 0; t0 < n_tokens; t0 += DJ_CS_TRIP) {
        const int m = (int)(n_tokens - t0 < DJ_CS_TRIP ? n_tokens - t0 : DJ_CS_TRIP);
        for (int t = lane; t < m; t += 32) {
            uint8_t *p = w.stage + w.fill + 3 * t;
            p[0] = as_deletion ? (uint8_t)'-' : (uint8_t)'=';
            p[1] = as_deletion ? w.seq[ref_pos + t0 + t] : base;
            p[2] = djcs_sep(w, w.row_pos + 3 * t + 2);
        }
        w.fill += 3 * m;
        w.row_pos += 3 * m;
        djcs_flush(w, lane, false);
    }
}

// non-zero iff a byte of x is zero
DJW_FN uint32_t djcs_has_zero_byte(uint32_t x) { return (x - 0x01010101u) & ~x & 0x80808080u; }

// byte mask of the word that holds bytes j0 .. j0 + 3 of a lane: 0xFF where j_lo <= j < j_hi
DJW_FN uint32_t djcs_valid_mask(int j0, int j_lo, int j_hi) {
    uint32_t m = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        if (j0 + b >= j_lo && j0 + b < j_hi) m |= 0xFFu << (8 * b);
    }
    return m;
}

DJW_FN int djcs_byte(const uint4 &v, int j) {
    const uint32_t word = j < 4 ? v.x : j < 8 ? v.y : j < 12 ? v.z : v.w;
    return (int)((word >> ((j & 3) * 8)) & 0xFFu);
}

// ---------------------------------------------------------------------------------------------------------
// The grammar as a table-driven state machine.  A byte has a class (cls[256]); a lane walks its sixteen bytes through
// tr[state][class], which yields the next state and what the byte emits:
//   code 1 "=X,"  2 "-X,"  3 "+X|"  (three bytes)   4 "*X"  5 "X,"  (the two letters of a substitution)   0 nothing
// The codes of a lane are kept as sixteen nibbles; sizes, reference tokens and grammar errors are summed on the way,
// so the bytes are written by a second, state-free walk over the nibbles once the warp's prefix sum has placed them.
// ---------------------------------------------------------------------------------------------------------
enum { DJCS_S_NONE = 0, DJCS_S_EQ, DJCS_S_DEL, DJCS_S_INS, DJCS_S_SUB0, DJCS_S_SUB1, DJCS_S_SUB2, DJCS_S_INTRON };
enum { DJCS_C_OUT = 0, DJCS_C_EQ, DJCS_C_DEL, DJCS_C_INS, DJCS_C_SUB, DJCS_C_INTRON, DJCS_C_LETTER, DJCS_C_DIGIT,
       DJCS_C_OTHER, DJCS_C_LETTER_N };

struct DjCsLut {
    // one look-up per byte: entry[state * 256 + byte], addressed in bytes as (state << 10) | (byte << 2).
    //   [3:0] code   [12:10] next state (already in address position)   [15] grammar error
    //   [17:16] bytes emitted   [22] reference token   [27] an "=N" token (is_n_tag, utils/midsv_handler.py:8-14)
    // so that sizes, tokens and "=N" tokens add up in one register (fields 16..21, 22..26, 27..31 for sixteen bytes)
    uint32_t entry[8 * 256];
};

DJW_FN uint8_t djcs_class_of(int c) {
    if (c == 0) return DJCS_C_OUT;
    if (c == '=') return DJCS_C_EQ;
    if (c == '-') return DJCS_C_DEL;
    if (c == '+') return DJCS_C_INS;
    if (c == '*') return DJCS_C_SUB;
    if (c == '~') return DJCS_C_INTRON;
    if ((c | 0x20) == 'n') return DJCS_C_LETTER_N;
    if ((unsigned)((c | 0x20) - 'a') < 26u) return DJCS_C_LETTER;
    if ((unsigned)(c - '0') < 10u) return DJCS_C_DIGIT;
    return DJCS_C_OTHER;
}

DJW_FN uint32_t djcs_transition(int state, int cls) {
    const uint32_t kBad = 1u << 15, kRef = 1u << 22, kIsN = 1u << 27, kThree = 3u << 16, kTwo = 2u << 16;
    int next = state;
    uint32_t out = 0;
    if (cls == DJCS_C_OUT) {
    } else if (cls >= DJCS_C_EQ && cls <= DJCS_C_INTRON) {  // an operator: a substitution must be complete
        if (state == DJCS_S_SUB0 || state == DJCS_S_SUB1) out |= kBad;
        next = cls == DJCS_C_EQ ? DJCS_S_EQ : cls == DJCS_C_DEL ? DJCS_S_DEL : cls == DJCS_C_INS ? DJCS_S_INS
             : cls == DJCS_C_SUB ? DJCS_S_SUB0 : DJCS_S_INTRON;
    } else if (cls == DJCS_C_LETTER || cls == DJCS_C_LETTER_N) {
        if (state == DJCS_S_EQ) out = 1u | kThree | kRef | (cls == DJCS_C_LETTER_N ? kIsN : 0u);
        else if (state == DJCS_S_DEL) out = 2u | kThree | kRef;
        else if (state == DJCS_S_INS) out = 3u | kThree;
        else if (state == DJCS_S_SUB0) { out = 4u | kTwo; next = DJCS_S_SUB1; }
        else if (state == DJCS_S_SUB1) { out = 5u | kTwo | kRef; next = DJCS_S_SUB2; }
        else if (state == DJCS_S_INTRON) out = 0;
        else out = kBad;  // no operator yet, or a third letter of a substitution
    } else if (cls == DJCS_C_DIGIT) {
        if (state != DJCS_S_INTRON) out = kBad;
    } else {
        out = kBad;
    }
    return out | ((uint32_t)next << 10);
}

// entries [first, 2048) in steps of `step` (a CTA fills its shared-memory copy with all its threads)
DJW_FN void djcs_lut_fill(DjCsLut *lut, int first, int step) {
    for (int k = first; k < 8 * 256; k += step) lut->entry[k] = djcs_transition(k >> 8, djcs_class_of(k & 255));
}

// state a lane starts in, from the last operator before it and the bytes since
DJW_FN int djcs_state_of(int op, int since) {
    if (op == '=') return DJCS_S_EQ;
    if (op == '-') return DJCS_S_DEL;
    if (op == '+') return DJCS_S_INS;
    if (op == '~') return DJCS_S_INTRON;
    if (op == '*') return DJCS_S_SUB0 + (since < 2 ? since : 2);
    return DJCS_S_NONE;
}

// (byte j of the lane) << 2
DJW_FN uint32_t djcs_byte4(const uint4 &v, int j) {
    const uint32_t word = j < 4 ? v.x : j < 8 ? v.y : j < 12 ? v.z : v.w;
    const int b = j & 3;
    return b == 0 ? (word << 2) & 0x3FCu : (word >> (8 * b - 2)) & 0x3FCu;
}

// Walk the lane's bytes through the table.  `active` (bit j = byte j counts) limits what is emitted, not the
// states.  Returns bytes emitted << 16 | reference tokens << 22 | "=N" tokens << 27 | grammar error << 15; the codes
// go to *lo (bytes 0..7) and *hi (bytes 8..15).
template <bool kMasked>
DJW_FN uint32_t djcs_classify(const DjCsLut &lut, const uint4 &v, int state, uint32_t active, uint32_t *lo,
                              uint32_t *hi) {
    const uint8_t *table = reinterpret_cast<const uint8_t *>(lut.entry);
    uint32_t c_lo = 0, c_hi = 0, acc = 0, bad = 0;
    uint32_t t = (uint32_t)state << 10;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        t = *reinterpret_cast<const uint32_t *>(table + ((t & 0x1C00u) | djcs_byte4(v, j)));
        if (kMasked && !((active >> j) & 1u)) t &= 0x9C00u;  // keeps the state and the error, emits nothing
        if (j < 8) c_lo |= (t & 15u) << (4 * j);
        else c_hi |= (t & 15u) << (4 * (j - 8));
        acc += t & 0xFFFF0000u;
        bad |= t;
    }
    *lo = c_lo;
    *hi = c_hi;
    return acc | (bad & 0x8000u);
}

// the bytes of the lane's codes at dst; `term` = offset (from dst) of the separator that ends the row, -1 if none
DJW_FN void djcs_emit(const uint4 &v, uint32_t lo, uint32_t hi, bool lower, uint8_t *dst, int term) {
    int o = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint32_t code = ((j < 8 ? lo : hi) >> (4 * (j & 7))) & 15u;
        if (code == 0) continue;
        const int c = djcs_byte(v, j);
        const uint8_t L = (uint8_t)(lower ? (c | 0x20) : (c & 0xDF));
        if (code <= 3) {
            dst[o] = code == 1 ? (uint8_t)'=' : code == 2 ? (uint8_t)'-' : (uint8_t)'+';
            dst[o + 1] = L;
            dst[o + 2] = code == 3 ? (uint8_t)'|' : (o + 2 == term ? (uint8_t)'\n' : (uint8_t)',');
            o += 3;
        } else if (code == 4) {
            dst[o] = '*';
            dst[o + 1] = L;
            o += 2;
        } else {
            dst[o] = L;
            dst[o + 1] = o + 1 == term ? (uint8_t)'\n' : (uint8_t)',';
            o += 2;
        }
    }
}

template <bool kWrite, bool kMasked>
DJW_FN void djcs_emit_range(const DjCsLut &lut, const uint4 &v, uint32_t active, int in_state, bool lower,
                            DjCsWriter<kWrite> &w, int lane, int *refs, int *bad) {
    uint32_t lo, hi;
    const uint32_t acc = djcs_classify<kMasked>(lut, v, in_state, active, &lo, &hi);
    const int mine = (int)((acc >> 16) & 0x3Fu);
    *refs += (int)((acc >> 22) & 0x1Fu);
    if (acc >> 27) *bad |= DJ_CS_HAS_N;
    if (acc & 0x8000u) *bad |= DJ_CS_GRAMMAR;
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = DJW_SHFL_UP(incl, d);
        if (lane >= d) incl += up;
    }
    const int total = DJW_SHFL(incl, 31);
    if (kWrite) {
        const int excl = incl - mine;
        const int64_t to_term = w.last - (w.row_pos + excl);
        djcs_emit(v, lo, hi, lower, w.stage + w.fill + excl, to_term >= 0 && to_term < 48 ? (int)to_term : -1);
        w.fill += total;
    }
    w.row_pos += total;
    djcs_flush(w, lane, false);
}

// One alignment: cs[a_off, a_end).  Adds its reference tokens to *refs (per lane: sum over the warp afterwards);
// returns the operator the alignment ends under (warp-uniform).
template <bool kWrite>
DJW_FN int djcs_expand(const DjCsLut &lut, const uint8_t *cs, int64_t a_off, int64_t a_end, bool lower, int64_t ref_start,
                       DjCsWriter<kWrite> &w, int lane, int *refs, int *bad) {
    int c_op = 0, c_since = 0;
    for (int64_t chunk = a_off & ~(int64_t)15; chunk < a_end; chunk += DJ_CS_TRIP) {
        const int64_t p0 = chunk + lane * 16;
        const bool interior = chunk >= a_off && chunk + DJ_CS_TRIP <= a_end;  // every byte of the trip is a tag byte
        uint4 v;
        v.x = v.y = v.z = v.w = 0;
        if (interior) {
            v = djw_load16(cs + p0);
        } else if (p0 < a_end && p0 + 16 > a_off) {
            v = djw_load16(cs + p0);
            const int j_lo = a_off > p0 ? (int)(a_off - p0) : 0;  // bytes outside [a_off, a_end) -> 0
            const int j_hi = a_end - p0 < 16 ? (int)(a_end - p0) : 16;
            v.x &= djcs_valid_mask(0, j_lo, j_hi);
            v.y &= djcs_valid_mask(4, j_lo, j_hi);
            v.z &= djcs_valid_mask(8, j_lo, j_hi);
            v.w &= djcs_valid_mask(12, j_lo, j_hi);
        }
        // '~' anywhere in the lane (introns are rare; they and the trips at the ends of a tag take the byte-wise scan)
        const uint32_t tilde_bytes = djcs_has_zero_byte(v.x ^ 0x7E7E7E7Eu) | djcs_has_zero_byte(v.y ^ 0x7E7E7E7Eu) |
                                     djcs_has_zero_byte(v.z ^ 0x7E7E7E7Eu) | djcs_has_zero_byte(v.w ^ 0x7E7E7E7Eu);
        const bool introns = DJW_BALLOT(tilde_bytes != 0) != 0;
        // NUL bytes inside a tag (the slack alignment.scan_sam leaves behind a shortened tag) are no operators either
        const uint32_t nul_bytes = djcs_has_zero_byte(v.x) | djcs_has_zero_byte(v.y) | djcs_has_zero_byte(v.z) |
                                   djcs_has_zero_byte(v.w);
        const bool plain_bytes = interior && !introns && c_op != '~' && DJW_BALLOT(nul_bytes != 0) == 0;
        // the last operator among the lane's bytes
        int last_idx = -1, last_op = 0;
        if (plain_bytes) {
            // Outside introns a tag byte without bit 6 is an operator ('=' '*' '-' '+'; letters and '~' have the bit;
            // digits and anything else without it are grammar errors the table flags, so a wrong carry cannot matter).
            const uint32_t z0 = ~v.x & 0x40404040u, z1 = ~v.y & 0x40404040u, z2 = ~v.z & 0x40404040u,
                           z3 = ~v.w & 0x40404040u;
            const uint32_t z = z3 ? z3 : z2 ? z2 : z1 ? z1 : z0;
            const uint32_t word = z3 ? v.w : z2 ? v.z : z1 ? v.y : v.x;
            if (z) {
                const int b = (31 - DJW_CLZ(z)) >> 3;
                last_idx = (z3 ? 12 : z2 ? 8 : z1 ? 4 : 0) + b;
                last_op = (int)((word >> (8 * b)) & 0xFFu);
            }
        } else {
            for (int j = 0; j < 16; ++j) {
                const int c = djcs_byte(v, j);
                if (c == '=' || c == '*' || c == '-' || c == '+' || c == '~') {
                    last_idx = j;
                    last_op = c;
                }
            }
        }
        const uint32_t has = DJW_BALLOT(last_idx >= 0);
        const uint32_t below = has & ((1u << lane) - 1u);
        const int src = below ? 31 - DJW_CLZ(below) : 0;
        const int s_op = DJW_SHFL(last_op, src);
        const int s_since = DJW_SHFL(15 - last_idx, src);
        const int in_op = below ? s_op : c_op;
        const int in_since = below ? s_since + 16 * (lane - src - 1) : c_since + 16 * lane;
        const int out_op = last_idx >= 0 ? last_op : in_op;
        const int out_since = last_idx >= 0 ? 15 - last_idx : in_since + 16;
        c_op = DJW_SHFL(out_op, 31);
        c_since = DJW_SHFL(out_since, 31);

        const int64_t trip_end = chunk + DJ_CS_TRIP < a_end ? chunk + DJ_CS_TRIP : a_end;
        const int in_state = djcs_state_of(in_op, in_since);
        if (!introns) {
            djcs_emit_range<kWrite, false>(lut, v, 0xFFFFu, in_state, lower, w, lane, refs, bad);
            continue;
        }
        // introns in this trip: bytes up to the '~', the run of "=N", the rest
        uint32_t tilde = 0;
        for (int j = 0; j < 16; ++j) {
            if (djcs_byte(v, j) == '~') tilde |= 1u << j;
        }
        int64_t lo = a_off;
        for (;;) {
            int64_t mine = INT64_MAX;
            for (int j = 15; j >= 0; --j) {
                if (((tilde >> j) & 1u) && p0 + j >= lo) mine = p0 + j;
            }
            int64_t first = mine;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                const int64_t other = DJW_SHFL_XOR(first, d);
                if (other < first) first = other;
            }
            {
                const int64_t hi = first < trip_end ? first : trip_end;
                const int j_lo = lo <= p0 ? 0 : lo - p0 >= 16 ? 16 : (int)(lo - p0);
                const int j_hi = hi <= p0 ? 0 : hi - p0 >= 16 ? 16 : (int)(hi - p0);
                const uint32_t active = j_hi > j_lo ? ((1u << j_hi) - 1u) & ~((1u << j_lo) - 1u) : 0u;
                djcs_emit_range<kWrite, true>(lut, v, active, in_state, lower, w, lane, refs, bad);
            }
            if (first == INT64_MAX) break;
            int64_t n = 0;
            bool digits = false;
            for (int64_t q = first + 3; q < a_end && q < first + 13; ++q) {  // "~" + two letters, then the length
                const int c = cs[q];
                if (c < '0' || c > '9') break;
                n = n * 10 + (c - '0');
                digits = true;
            }
            if (!digits) *bad |= DJ_CS_GRAMMAR;
            int before = *refs;  // reference tokens of this alignment in front of the intron
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) before += DJW_SHFL_XOR(before, d);
            djcs_fill(w, lane, n, lower, true, ref_start + before);
            if (lane == 0) *refs += (int)(n > 0x3fffffff ? 0x3fffffff : n);
            lo = first + 1;
        }
    }
    return c_op;
}

// One read: flank, alignments in POS order with the gaps between them, flank.  kWrite == false measures
// (row_len_io receives the row length, 0 for a dropped read; status its DJ_CS_* bits; row_n, if not null, the number
// of "=N" tokens the filler wrote); kWrite == true writes the row of a read whose measured length is *row_len_io
// (> 0) at `row`.  ref_seq (may be null): see djcs_fill.
template <bool kWrite>
DJW_FN void djcs_read(const DjCsLut &lut, const uint8_t *cs, const int64_t *aln_off, const int32_t *aln_pos, const uint8_t *aln_lower,
                      int64_t first_aln, int64_t end_aln, int32_t ref_len, const uint8_t *ref_seq, uint8_t *stage,
                      uint8_t *row, int32_t *row_len_io, int32_t *status, int32_t *row_n, int lane) {
    DjCsWriter<kWrite> w;
    w.stage = stage;
    w.row_pos = 0;
    w.last = -1;
    w.fill = w.head_skip = 0;
    w.gbase = nullptr;
    w.seq = ref_seq;
    w.n_tokens = 0;
    if (kWrite) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(row);
        w.gbase = reinterpret_cast<uint8_t *>(addr & ~(uintptr_t)15);
        w.fill = w.head_skip = (int)(addr & 15);
        w.last = *row_len_io;
    }
    int64_t cur = 0;  // reference tokens of the row so far
    int bad = 0;
    int ends = 0;
    for (int64_t a = first_aln; a < end_aln; ++a) {
        const int64_t start = aln_pos[a];
        if (start < cur) {
            bad |= DJ_CS_OVERLAP;
            break;
        }
        if (start > cur) {
            djcs_fill(w, lane, start - cur, false, a > first_aln, cur);
            cur = start;
        }
        int refs = 0;
        ends = djcs_expand<kWrite>(lut, cs, aln_off[a], aln_off[a + 1], aln_lower[a] != 0, cur, w, lane, &refs, &bad);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            refs += DJW_SHFL_XOR(refs, d);
            bad |= DJW_SHFL_XOR(bad, d);
        }
        cur += refs;
        if (cur > ref_len) {
            bad |= DJ_CS_TOO_LONG;
            break;
        }
    }
    if (!(bad & (DJ_CS_OVERLAP | DJ_CS_TOO_LONG))) {
        if (cur < ref_len) {
            djcs_fill(w, lane, (int64_t)ref_len - cur, false, false, cur);
        } else if (ends == '+' || first_aln == end_aln) {
            bad |= DJ_CS_GRAMMAR;  // an insertion with nothing after it (or ref_len == 0)
        }
    }
    if (kWrite) {
        djcs_flush(w, lane, true);
    } else if (lane == 0) {
        if (w.row_pos - 1 > 0x7fffffff - 64) bad |= DJ_CS_TOO_LONG;
        *status = bad;
        *row_len_io = (bad & ~DJ_CS_HAS_N) ? 0 : (int32_t)(w.row_pos - 1);
        if (row_n) *row_n = (int32_t)(w.n_tokens > 0x7fffffff ? 0x7fffffff : w.n_tokens);
    }
}

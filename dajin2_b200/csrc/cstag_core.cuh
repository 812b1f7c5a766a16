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

template <bool kWrite>
struct DjCsWriter {
    uint8_t *stage;    // this warp's stage; stage[k] goes to gbase[k]
    uint8_t *gbase;    // 16-byte aligned
    int fill;          // bytes in the stage, the first head_skip of them not ours
    int head_skip;
    int64_t row_pos;   // bytes of the row produced so far
    int64_t last;      // row position of the separator that ends the row ('\n' instead of ','); -1 when measuring
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

// n tokens "=N," (lower: "=n,")
template <bool kWrite>
DJW_FN void djcs_fill(DjCsWriter<kWrite> &w, int lane, int64_t n_tokens, bool lower) {
    if (!kWrite) {
        w.row_pos += 3 * n_tokens;
        return;
    }
    const uint8_t base = lower ? (uint8_t)'n' : (uint8_t)'N';
    for (int64_t t0 = 0; t0 < n_tokens; t0 += DJ_CS_TRIP) {
        const int m = (int)(n_tokens - t0 < DJ_CS_TRIP ? n_tokens - t0 : DJ_CS_TRIP);
        for (int t = lane; t < m; t += 32) {
            uint8_t *p = w.stage + w.fill + 3 * t;
            p[0] = '=';
            p[1] = base;
            p[2] = djcs_sep(w, w.row_pos + 3 * t + 2);
        }
        w.fill += 3 * m;
        w.row_pos += 3 * m;
        djcs_flush(w, lane, false);
    }
}

DJW_FN bool djcs_is_op(int c) { return c == '=' || c == '*' || c == '-' || c == '+' || c == '~'; }

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

// The bytes of this lane with absolute positions in [lo, hi): sizes only (emit == false) or sizes + bytes.
// Returns the bytes produced; *refs / *bad collect reference tokens and grammar errors.
template <bool kWrite>
DJW_FN int djcs_walk(const uint4 &v, int64_t p0, int64_t lo, int64_t hi, int op, int since, bool lower, bool emit,
                     const DjCsWriter<kWrite> &w, uint8_t *dst, int64_t dst_pos, int *refs, int *bad) {
    int total = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int c = djcs_byte(v, j);
        if (c == 0) {  // outside the alignment
            ++since;
            continue;
        }
        if (djcs_is_op(c)) {
            if (op == '*' && since != 2 && !emit) *bad |= DJ_CS_GRAMMAR;  // a substitution is "*" + two letters
            op = c;
            since = 0;
            continue;
        }
        ++since;
        const int64_t p = p0 + j;
        if (p < lo || p >= hi || op == '~') continue;
        const bool letter = (unsigned)((c | 0x20) - 'a') < 26u;
        const uint8_t L = (uint8_t)(lower ? (c | 0x20) : (c & 0xDF));
        int n = 0;
        if (!letter || op == 0) {
            if (!emit) *bad |= DJ_CS_GRAMMAR;
        } else if (op == '*') {
            if (since == 1) {
                n = 2;
                if (kWrite && emit) {
                    dst[total] = '*';
                    dst[total + 1] = L;
                }
            } else if (since == 2) {
                n = 2;
                if (!emit) ++*refs;
                if (kWrite && emit) {
                    dst[total] = L;
                    dst[total + 1] = djcs_sep(w, dst_pos + total + 1);
                }
            } else if (!emit) {
                *bad |= DJ_CS_GRAMMAR;
            }
        } else {
            n = 3;
            if (op != '+' && !emit) ++*refs;
            if (kWrite && emit) {
                dst[total] = (uint8_t)op;
                dst[total + 1] = L;
                dst[total + 2] = op == '+' ? (uint8_t)'|' : djcs_sep(w, dst_pos + total + 2);
            }
        }
        total += n;
    }
    return total;
}

template <bool kWrite>
DJW_FN void djcs_emit_range(const uint4 &v, int64_t p0, int64_t lo, int64_t hi, int in_op, int in_since, bool lower,
                            DjCsWriter<kWrite> &w, int lane, int *refs, int *bad) {
    const int mine = djcs_walk<kWrite>(v, p0, lo, hi, in_op, in_since, lower, false, w, nullptr, 0, refs, bad);
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = DJW_SHFL_UP(incl, d);
        if (lane >= d) incl += up;
    }
    const int total = DJW_SHFL(incl, 31);
    if (kWrite) {
        const int excl = incl - mine;
        int dummy_refs = 0, dummy_bad = 0;
        djcs_walk<kWrite>(v, p0, lo, hi, in_op, in_since, lower, true, w, w.stage + w.fill + excl, w.row_pos + excl,
                          &dummy_refs, &dummy_bad);
        w.fill += total;
    }
    w.row_pos += total;
    djcs_flush(w, lane, false);
}

// One alignment: cs[a_off, a_end).  Adds its reference tokens to *refs (per lane: sum over the warp afterwards);
// returns the operator the alignment ends under (warp-uniform).
template <bool kWrite>
DJW_FN int djcs_expand(const uint8_t *cs, int64_t a_off, int64_t a_end, bool lower, DjCsWriter<kWrite> &w, int lane,
                       int *refs, int *bad) {
    int c_op = 0, c_since = 0;
    for (int64_t chunk = a_off & ~(int64_t)15; chunk < a_end; chunk += DJ_CS_TRIP) {
        const int64_t p0 = chunk + lane * 16;
        uint4 v;
        v.x = v.y = v.z = v.w = 0;
        if (p0 < a_end && p0 + 16 > a_off) {
            v = djw_load16(cs + p0);
            const int j_lo = a_off > p0 ? (int)(a_off - p0) : 0;  // bytes outside [a_off, a_end) -> 0
            const int j_hi = a_end - p0 < 16 ? (int)(a_end - p0) : 16;
            v.x &= djcs_valid_mask(0, j_lo, j_hi);
            v.y &= djcs_valid_mask(4, j_lo, j_hi);
            v.z &= djcs_valid_mask(8, j_lo, j_hi);
            v.w &= djcs_valid_mask(12, j_lo, j_hi);
        }
        int last_idx = -1, last_op = 0;
        uint32_t tilde = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int c = djcs_byte(v, j);
            if (djcs_is_op(c)) {
                last_idx = j;
                last_op = c;
                if (c == '~') tilde |= 1u << j;
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
        if (DJW_BALLOT(tilde != 0) == 0) {
            djcs_emit_range<kWrite>(v, p0, a_off, trip_end, in_op, in_since, lower, w, lane, refs, bad);
            continue;
        }
        // introns in this trip: bytes up to the '~', the run of "=N", the rest
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
            djcs_emit_range<kWrite>(v, p0, lo, first < trip_end ? first : trip_end, in_op, in_since, lower, w, lane, refs,
                                    bad);
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
            djcs_fill(w, lane, n, lower);
            if (lane == 0) *refs += (int)(n > 0x3fffffff ? 0x3fffffff : n);
            lo = first + 1;
        }
    }
    return c_op;
}

// One read: flank, alignments in POS order with the gaps between them, flank.  kWrite == false measures
// (row_len_io receives the row length, 0 for a dropped read; status its DJ_CS_* bits); kWrite == true writes the row
// of a read whose measured length is *row_len_io (> 0) at `row`.
template <bool kWrite>
DJW_FN void djcs_read(const uint8_t *cs, const int64_t *aln_off, const int32_t *aln_pos, const uint8_t *aln_lower,
                      int64_t first_aln, int64_t end_aln, int32_t ref_len, uint8_t *stage, uint8_t *row,
                      int32_t *row_len_io, int32_t *status, int lane) {
    DjCsWriter<kWrite> w;
    w.stage = stage;
    w.row_pos = 0;
    w.last = -1;
    w.fill = w.head_skip = 0;
    w.gbase = nullptr;
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
            djcs_fill(w, lane, start - cur, false);
            cur = start;
        }
        int refs = 0;
        ends = djcs_expand<kWrite>(cs, aln_off[a], aln_off[a + 1], aln_lower[a] != 0, w, lane, &refs, &bad);
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
            djcs_fill(w, lane, (int64_t)ref_len - cur, false);
        } else if (ends == '+' || first_aln == end_aln) {
            bad |= DJ_CS_GRAMMAR;  // an insertion with nothing after it (or ref_len == 0)
        }
    }
    if (kWrite) {
        djcs_flush(w, lane, true);
    } else if (lane == 0) {
        if (w.row_pos - 1 > 0x7fffffff - 64) bad |= DJ_CS_TOO_LONG;
        *status = bad;
        *row_len_io = bad ? 0 : (int32_t)(w.row_pos - 1);
    }
}

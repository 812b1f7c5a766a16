// MIDSV text -> read x locus code matrix (dj_encode): the UNIT encoder.
//
// One warp streams one read.  The text of the read arrives in shared memory through a per-warp ring of bulk
// asynchronous copies (cp.async.bulk + mbarrier, one elected lane issues them, the ring runs kStages - 1 trips -- and
// across row boundaries -- ahead of the arithmetic): no thread issues a global load for text, nothing waits on one.
// A trip is 1536 bytes = 32 UNITS of 48 bytes, one per lane (three conflict-free 16-byte shared loads).
//
// Fast path (encode_unit.cuh): 98 tokens in 100 are "=X," (or "-X,").  Every lane builds the comma bitmap of its
// unit (count of closed tokens = popc, for any unit), tests it against the periodic pattern of its phase and, with
// byte permutes, translates and verifies the sixteen bases / operators of a plain unit.  Plain lanes write their
// sixteen code bytes as whole words into the warp's staging buffer (rotated to the byte offset their rank dictates;
// the word shared with a plain neighbour is merged through one shuffle).
//
// Slow path: the units that hold anything else (an insertion, a substitution, lowercase, N-free garbage, a row edge)
// are cut into 16-byte chunks, the chunks are COMPACTED over the lanes of the warp (a trip has ~4 such units of 32 on
// the benchmark profile, i.e. ~13 chunks: one pass) and go through the generic comma slots of encode_common.cuh: a
// token is identified from its closing comma with one perfect-hash look-up.  Both paths produce the same bytes
// (tests/test_gpu_parity.py holds dj_encode to the generic kernel dj_encode_v1 and to the host tokenizer).
//
// Staged code bytes leave the SM as aligned 16-byte stores of complete 32-byte sectors.  HBM traffic is the text once
// in and the code bytes once out (algorithmic bytes per read: row_len + n_loci + 16).
//
// calc_match (src/DAJIN2/core/classification/classifier.py:11-24), validation flags and checkpoints: as described in
// encode_common.cuh; the checkpoint grid is the trip (DJ_ENC_CHUNK = 1536 bytes).
//
// Reference behaviour covered here: the tokenisation every pass of the reference starts with
// (`json.loads` + `MIDSV.split(",")`, src/DAJIN2/utils/fileio.py:49-52) and calc_match.
#include <type_traits>

#include "encode_common.cuh"
#include "encode_unit.cuh"

#ifndef DJ_ENC_WARPS
#define DJ_ENC_WARPS 24  // warps per CTA, one CTA per SM (80 registers; 28 warps at 72 registers measured the same)
#endif
#ifndef DJ_ENC_STAGES
#define DJ_ENC_STAGES 2  // slots of the per-warp text ring (one trip in flight while one is translated)
#endif
#ifndef DJ_ENC_UNITS
#define DJ_ENC_UNITS 2   // units per lane and trip
#endif
#ifndef DJ_ENC_BULK
#define DJ_ENC_BULK 1    // 1: the ring is filled by bulk copies (cp.async.bulk + mbarrier); 0: by per-lane 16-byte cp.async
#endif

static_assert(DJ_ENC_CHUNK == dju::kTripBytes, "the checkpoint grid is one unit per lane of the unit encoder");

namespace {

constexpr int kWarps = DJ_ENC_WARPS;
constexpr int kStages = DJ_ENC_STAGES;
constexpr int kU = DJ_ENC_UNITS;
constexpr int kHalf = dju::kTripBytes;             // 1536 bytes: one unit per lane (the checkpoint grid)
constexpr int kTrip = kU * kHalf;                  // bytes of text per trip of the main loop
constexpr int kSlotBytes = 16 + kTrip;             // 16 bytes in front of the text: the word before the trip lives there
constexpr int kMaxTokensPerTrip = kU * 544;        // a valid trip closes at most kU * 512 + 1 tokens
constexpr int kStageBytes = 64 + kMaxTokensPerTrip;  // 31 carried bytes + the trip's codes + one rotated spill word
constexpr int kListBytes = kU * 32;                // unit ids of the slow path, compacted
constexpr int kCntBytes = kU * 32 * 4;             // per unit: commas of its three chunks (one byte each)
constexpr int kExclBytes = kU * 32 * 2;            // per unit: rank of its first token inside the trip
constexpr int kWarpBytes = kStages * kSlotBytes + kStageBytes + kListBytes + kCntBytes + kExclBytes;
constexpr int kPatOffset = kLutSize * 4;           // four phase entries (uint4)
constexpr int kBarOffset = kPatOffset + 64;
constexpr int kWarpOffset = (kBarOffset + kWarps * kStages * 8 + 127) & ~127;
constexpr int kSmemBytes = kWarpOffset + kWarps * kWarpBytes;
static_assert(kWarpBytes % 16 == 0 && kSlotBytes % 16 == 0 && kStageBytes % 16 == 0, "16-byte alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared memory of the unit encoder");
static_assert(kU >= 1 && kU <= 4, "units per lane");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n"
        "DJ_WAIT:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DJ_DONE;\n"
        " bra DJ_WAIT;\n"
        "DJ_DONE:\n}\n" ::"r"(bar), "r"(parity)
        : "memory");
}

// the whole warp calls (warp-uniform arguments); one elected lane arms the barrier with the byte count and starts the
// bulk copy global -> shared
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "{\n .reg .pred p;\n"
        " elect.sync _|p, 0xffffffff;\n"
        " @p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n"
        " @p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n}\n" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// DJ_ENC_BULK = 0: every lane copies 16 bytes global -> shared (LDGSTS, L1 bypassed); completion is tracked per thread
// in commit groups
__device__ __forceinline__ void lane_copy16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// comma flags of a piece of kW words in the layout emit_piece expects (bit 7-k of byte j <=> byte j of word k), exact
template <int kW>
__device__ __forceinline__ uint32_t piece_commas_exact(const uint32_t (&w)[kW], uint32_t &high) {
    uint32_t zz = 0;
#pragma unroll
    for (int k = 0; k < kW; ++k) {
        high |= w[k];
        zz |= dju::comma_flags_exact(w[k]) >> k;
    }
    return zz;
}

// The text of the slow path is cut into PIECES, one per lane and pass: 16-byte thirds of a unit (kW = 4 words, eight
// comma slots) or 24-byte halves (kW = 6, twelve slots).  A pass costs about 50 + 30 instructions per slot whatever
// the number of busy lanes, so the size is chosen per trip from the number of slow units: whichever of thirds /
// halves needs fewer slot rounds (halves for 11..16 units: one pass at 1.4 times the cost instead of two).  On the
// benchmark profile a trip has 10.3 slow units on average.  Measured and dropped: 12-byte quarters for trips with at
// most eight slow units (0.603 against 0.616 of peak).
template <int kW>
__device__ __forceinline__ void load_piece(const uint8_t *p, uint32_t (&w)[kW]) {
    if constexpr (kW == 4) {
        const uint4 q = *reinterpret_cast<const uint4 *>(p);
        w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k < kW; k += 2) {
            const uint2 q = *reinterpret_cast<const uint2 *>(p + 4 * k);
            w[k] = q.x; w[k + 1] = q.y;
        }
    }
}

// inclusive warp scan, one shuffle and one predicated add per step
__device__ __forceinline__ uint32_t warp_scan_incl(uint32_t x) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
        asm volatile("{\n .reg .pred p;\n .reg .u32 t;\n shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n @p add.u32 %0, %0, t;\n}\n"
                     : "+r"(x) : "r"(d));
    return x;
}

struct RowGeom {
    const uint8_t *base;  // 16-byte aligned start of the row's span
    int head;             // bytes of the span in front of the row
    int nbytes;           // head + row length: position of the virtual closing comma
    int n_trips;
};

__device__ __forceinline__ RowGeom row_geom(const uint8_t *text, int64_t off, int len) {
    RowGeom g;
    g.head = (int)(off & 15);
    g.base = text + (off - g.head);
    g.nbytes = g.head + len;
    g.n_trips = g.nbytes / kTrip + 1;  // the unit that holds byte `nbytes` closes the last token
    return g;
}

__global__ void __launch_bounds__(kWarps * 32, 1)
encode_unit_kernel(const uint8_t *__restrict__ text, const int64_t *__restrict__ row_off,
                   const int32_t *__restrict__ row_len, int64_t n_reads, int n_loci, int code_pitch, int ckpt_pitch,
                   uint8_t *__restrict__ codes, int32_t *__restrict__ match_score, int32_t *__restrict__ n_tokens,
                   uint32_t *__restrict__ flags, uint32_t *__restrict__ ckpt, uint32_t c_four, uint32_t c_shift,
                   dju::Muls muls) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem);
    uint4 *pat = reinterpret_cast<uint4 *>(smem + kPatOffset);
    build_token_lut(lut);
    if (threadIdx.x < 4) {
        const dju::Phase ph = dju::phase_entry(threadIdx.x);
        pat[threadIdx.x] = make_uint4(ph.shift, ph.trail, ph.bad, ph.n0);
    }

    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    uint8_t *wbase = smem + kWarpOffset + wid * kWarpBytes;
    uint8_t *stage = wbase + kStages * kSlotBytes;
    uint8_t *list = stage + kStageBytes;
    uint8_t *cntbuf = list + kListBytes;
    uint16_t *exclbuf = reinterpret_cast<uint16_t *>(cntbuf + kCntBytes);
    const uint32_t ring_s = smem_u32(wbase);
    const uint32_t stage_s = smem_u32(stage);
    const uint32_t lut_s = smem_u32(lut);
    const uint32_t bar_s = smem_u32(smem + kBarOffset) + (uint32_t)wid * (kStages * 8);
#pragma unroll
    for (int h = 0; h < kU; ++h) reinterpret_cast<uint32_t *>(cntbuf)[h * 32 + lane] = 0u;
    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(bar_s + 8u * s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int64_t warp = (int64_t)blockIdx.x * kWarps + wid;
    const int64_t n_warps = (int64_t)gridDim.x * kWarps;
    const uint32_t lanes_below = (1u << lane) - 1u;

    // ---- producer: the sequence of (row, trip) this warp will consume, issued kStages - 1 trips ahead -------------
    int64_t p_row = warp;
    const uint8_t *p_src = text;  // next 16-byte aligned piece of the producer's row
    int p_left = 0;               // bytes of the row's span (rounded up to 16) from p_src on; may run below 16
    int p_trips = 0;              // trips of the producer's row still to issue; 0: no row left
    int p_slot = 0;
    int64_t p_next_off = 0;
    int p_next_len = 0;
    auto p_open = [&](int64_t off, int len) {
        const RowGeom g = row_geom(text, off, len);
        p_src = g.base;
        p_left = (g.nbytes + 15) & ~15;
        p_trips = g.n_trips;
        if (p_row + n_warps < n_reads) {
            p_next_off = row_off[p_row + n_warps];
            p_next_len = row_len[p_row + n_warps];
        }
    };
    if (p_row < n_reads) p_open(row_off[p_row], row_len[p_row]);
    auto produce = [&]() {
        if (p_trips == 0) {
#if !DJ_ENC_BULK
            asm volatile("cp.async.commit_group;" ::: "memory");  // the consumer counts groups: one per call
#endif
            return;
        }
        const int bytes = p_left > kTrip ? kTrip : (p_left < 16 ? 16 : p_left);
#if DJ_ENC_BULK
        bulk_load(ring_s + (uint32_t)(p_slot * kSlotBytes + 16), p_src, (uint32_t)bytes, bar_s + 8u * p_slot);
#else
        {
            const uint32_t dst = ring_s + (uint32_t)(p_slot * kSlotBytes + 16 + 16 * lane);
            const uint8_t *src = p_src + 16 * lane;
            if (bytes == kTrip) {
#pragma unroll
                for (int k = 0; k < kTrip / 512; ++k) lane_copy16(dst + 512 * k, src + 512 * k);
            } else {
#pragma unroll 1
                for (int b = 16 * lane; b < bytes; b += 512) lane_copy16(dst + (uint32_t)(b - 16 * lane), src + (b - 16 * lane));
            }
        }
#endif
#if !DJ_ENC_BULK
        asm volatile("cp.async.commit_group;" ::: "memory");
#endif
        p_slot = p_slot + 1 == kStages ? 0 : p_slot + 1;
        p_src += kTrip;
        p_left -= kTrip;
        if (--p_trips == 0) {
            p_row += n_warps;
            if (p_row < n_reads) p_open(p_next_off, p_next_len);
        }
    };
#pragma unroll 1
    for (int s = 0; s < kStages - 1; ++s) produce();

    int c_slot = 0;
#if DJ_ENC_BULK
    uint32_t c_phase = 0;  // bit s: parity the consumer waits for on slot s
#endif

    for (int64_t r = warp; r < n_reads; r += n_warps) {
        const int64_t off = row_off[r];
        const int len = row_len[r];
        const RowGeom geom = row_geom(text, off, len);
        const int head = geom.head, nbytes = geom.nbytes;
        uint8_t *out = codes + r * (int64_t)code_pitch;
        uint32_t *ck = ckpt + r * (int64_t)ckpt_pitch;

        int tokbase = 0;                  // commas seen so far = tokens closed so far
        uint32_t mism = 0, n_sub = 0;     // calc_match contributions of the anchors; substitutions
        uint32_t sticky = 0;              // kStickyBroken, rank errors (bits 0..3), specials (bits 8..12)
        uint32_t high = 0;                // OR of the text words of the slow path: bit 7 of a byte <=> not 7-bit ASCII
        uint32_t carry_w = kFill;         // the word in front of the trip's first byte
        uint8_t *outp = out + lane * 16;  // where this lane's 16 bytes of the next flush go
        bool broken = false;

#pragma unroll 1
        for (int t = 0; t < geom.n_trips; ++t) {
            produce();  // refills the slot the previous trip read (all of its reads are behind a __syncwarp)
#if DJ_ENC_BULK
            mbar_wait(bar_s + 8u * c_slot, (c_phase >> c_slot) & 1u);
            c_phase ^= 1u << c_slot;
#else
            asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");  // this trip's group has landed
#endif
            uint8_t *sb = wbase + c_slot * kSlotBytes;
            c_slot = c_slot + 1 == kStages ? 0 : c_slot + 1;
            if (broken) {
                __syncwarp();
                continue;
            }
            const int g0 = t * kTrip;
            const bool edge_trip = (t == 0 && head > 0) || g0 + kTrip > nbytes;
            if (lane == 0) *reinterpret_cast<uint32_t *>(sb + 12) = carry_w;
            if (edge_trip) {
                // bytes of the trip that lie outside the row: ' ' in front of the row, ONE ',' at `nbytes` (it closes
                // the last token) and ' ' behind it up to the end of its unit; units further back are dead
                if (t == 0 && lane < head) sb[16 + lane] = ' ';
                const int rel = nbytes - g0;  // >= 0; inside this trip <=> this is the row's last trip
                if (rel < kTrip) {
                    const int unit_end = (rel / dju::kUnitBytes + 1) * dju::kUnitBytes;
                    if (rel + lane < unit_end) sb[16 + rel + lane] = lane == 0 ? ',' : ' ';
                    if (rel + 32 + lane < unit_end) sb[16 + rel + 32 + lane] = ' ';
                }
            }
            __syncwarp();

            // ---- every lane: its units; sixteen codes each if the unit is plain ---------------------------------
            uint32_t cw[kU][4];
            uint32_t n_del[kU];
            bool plain[kU], dead[kU];
            bool tail[kU];  // the unit holds the row's end: nothing is stored behind it
#pragma unroll
            for (int h = 0; h < kU; ++h) {
                const uint8_t *up = sb + 16 + h * kHalf + dju::kUnitBytes * lane;
                uint32_t w[12];
                {
                    const uint4 q0 = reinterpret_cast<const uint4 *>(up)[0], q1 = reinterpret_cast<const uint4 *>(up)[1],
                                q2 = reinterpret_cast<const uint4 *>(up)[2];
                    w[0] = q0.x; w[1] = q0.y; w[2] = q0.z; w[3] = q0.w;
                    w[4] = q1.x; w[5] = q1.y; w[6] = q1.z; w[7] = q1.w;
                    w[8] = q2.x; w[9] = q2.y; w[10] = q2.z; w[11] = q2.w;
                }
                const uint32_t wprev = *reinterpret_cast<const uint32_t *>(up - 4);
                const uint4 pe = pat[dju::phase_index(w[0])];
                dju::Phase ph;
                ph.shift = pe.x; ph.trail = pe.y; ph.bad = pe.z; ph.n0 = pe.w;
                const dju::UnitOut uo = dju::decode_unit(wprev, w, ph, muls);
                const int g = g0 + h * kHalf + dju::kUnitBytes * lane;
                dead[h] = g > nbytes;  // past the virtual closing comma
                tail[h] = g + dju::kUnitBytes > nbytes;
                plain[h] = !dead[h] && uo.bad == 0;
                n_del[h] = uo.n_del;
#pragma unroll
                for (int q = 0; q < 4; ++q) cw[h][q] = uo.c[q];
            }

            // ---- ranks: a plain unit closes sixteen tokens; the chunks of the other units are compacted over the
            //      lanes and counted (the same compaction translates them further down) ------------------------------
            uint32_t slow_mask[kU], cnt[kU];
            uint32_t zz_first = 0;
            int n_pieces = 0;
            int per_unit = 3;  // pieces per slow unit in this trip: 3 thirds or 2 halves
            bool force = false;
#pragma unroll 1
            for (;;) {
                int n_slow = 0;
#pragma unroll
                for (int h = 0; h < kU; ++h) {
                    const bool p = plain[h] && !force;
                    slow_mask[h] = __ballot_sync(0xffffffffu, !p && !dead[h]);
                    if (!p && !dead[h]) list[n_slow + __popc(slow_mask[h] & lanes_below)] = (uint8_t)(h * 32 + lane);
                    n_slow += __popc(slow_mask[h]);
                }
                // halves where ceil(2 n / 32) passes of halves cost less than ceil(3 n / 32) passes of thirds (cost ratio
                // 10 : 7): n = 11..16, 22..32, 43..48, 54..64 -- bit n of the masks (n - 32 in the second word)
                per_unit = (((n_slow < 32 ? 0xFFC1F800u : 0xFFC1F801u) >> (n_slow & 31)) & 1u) != 0u ? 2 : 3;
                n_pieces = per_unit * n_slow;
                if (n_slow > 0) {
                    __syncwarp();
                    auto count_pass = [&](auto kw_tag) {
                        constexpr int kW = decltype(kw_tag)::value;
                        constexpr int kPer = 12 / kW;  // pieces per unit
#pragma unroll 1
                        for (int j0 = 0; j0 < n_pieces; j0 += 32) {
                            const int j = j0 + lane;
                            if (j < n_pieces) {
                                const int ui = kPer == 2 ? j >> 1 : (int)(((uint32_t)j * 43691u) >> 17);  // j / kPer
                                const int c = j - kPer * ui;
                                const int u = list[ui];
                                uint32_t w[kW];
                                load_piece<kW>(sb + 16 + dju::kUnitBytes * u + 4 * kW * c, w);
                                const uint32_t zz = piece_commas_exact<kW>(w, high);
                                cntbuf[4 * u + c] = (uint8_t)__popc(zz);
                                if (kPer == 2 && c == 1) cntbuf[4 * u + 2] = 0;
                                if (j0 == 0) zz_first = zz;
                            }
                        }
                    };
                    if (per_unit == 3) count_pass(std::integral_constant<int, 4>{});
                    else count_pass(std::integral_constant<int, 6>{});
                    __syncwarp();
                }
                bool small = false;
#pragma unroll
                for (int h = 0; h < kU; ++h) {
                    const bool p = plain[h] && !force;
                    const uint32_t cb = reinterpret_cast<const uint32_t *>(cntbuf)[h * 32 + lane];
                    const uint32_t slow_cnt = __dp4a(cb, 0x01010101u, 0u);  // the fourth byte is never written (zero)
                    cnt[h] = dead[h] ? 0u : (p ? 16u : slow_cnt);
                    small = small || (!tail[h] && !p && slow_cnt < 4u);
                }
                // a unit that closes fewer than four tokens (the inside of a long insertion) does not cover the words
                // it shares with its neighbours: such a trip goes through the slow path as a whole (the unit that holds
                // the row's end has no neighbour behind it)
                if (force || !__any_sync(0xffffffffu, small)) break;
                bool mine_plain = false;
#pragma unroll
                for (int h = 0; h < kU; ++h) mine_plain = mine_plain || plain[h];
                if (!__any_sync(0xffffffffu, mine_plain)) break;
                force = true;
            }

            // inclusive scans of the units' token counts over the lanes, two 16-bit fields per word; unit (h, lane)
            // follows every unit of the rows h' < h and the lanes below it in its own row
            constexpr int kScanWords = (kU + 1) / 2;
            uint32_t incl[kScanWords];
#pragma unroll
            for (int q = 0; q < kScanWords; ++q)
                incl[q] = warp_scan_incl(cnt[2 * q] | (2 * q + 1 < kU ? cnt[(2 * q + 1) % kU] << 16 : 0u));
            uint32_t row_totals[kScanWords];
#pragma unroll
            for (int q = 0; q < kScanWords; ++q) row_totals[q] = __shfl_sync(0xffffffffu, incl[q], 31);
            int row_base[kU];  // tokens the rows of units in front of row h close
            int total = 0;
            uint32_t excl[kU];
#pragma unroll
            for (int h = 0; h < kU; ++h) {
                row_base[h] = total;
                excl[h] = (uint32_t)total + ((h & 1) ? incl[h / 2] >> 16 : incl[h / 2] & 0xFFFFu) - cnt[h];
                total += (int)((h & 1) ? row_totals[h / 2] >> 16 : row_totals[h / 2] & 0xFFFFu);
            }
            // not MIDSV: more commas than a trip of tokens can hold, or more tokens than the row has room for (tested
            // BEFORE anything is staged or stored: a row never writes past its own row of the code matrix)
            if (total > kMaxTokensPerTrip || tokbase + total > code_pitch) {
                sticky |= kStickyBroken;
                broken = true;
#if DJ_ENC_BULK
                if (edge_trip) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
                __syncwarp();
                continue;
            }
            if (lane < kU) {
                // checkpoints: tokens that START before byte j * DJ_ENC_CHUNK of the span (dj_find_token resumes there);
                // lane h writes the one of row h of units: the tokens closed in front of it, and one more unless the byte
                // in front of the row is a comma (for h = 0 that byte is the last of the previous trip, at sb + 15)
                int before = tokbase;
#pragma unroll
                for (int h = 1; h < kU; ++h)
                    if (lane == h) before = tokbase + row_base[h];
                const uint32_t v = (uint32_t)before + (sb[15 + lane * kHalf] == ',' ? 0u : 1u);
                if (lane == 0 ? true : g0 + lane * kHalf <= nbytes) ck[lane] = lane == 0 && t == 0 ? 0u : v;
            }
            ck += kU;
            const uint32_t t0 = (uint32_t)tokbase & 31u;  // stage[0] holds locus (tokbase & ~31) of the row

            // ---- plain units: sixteen code bytes as whole words (the word shared with a plain neighbour is merged
            //      through a shuffle; a word shared with another unit is overwritten by the slow path afterwards) -------
            uint32_t last_spill = 0;  // spill word of lane 31 of the previous row of units (zero unless it is plain)
#pragma unroll
            for (int h = 0; h < kU; ++h) {
                const bool p = plain[h] && !force;
                const uint32_t o = t0 + excl[h];
                uint32_t rot[5];
                dju::rotate_codes(cw[h], o & 3u, rot);
                const uint32_t spill_out = p ? rot[4] : 0u;
                // the bytes in front of this unit inside its first word: the spill of the plain unit in front of it
                // (lane - 1, or lane 31 of the previous row), the codes carried over from the previous trip for the
                // trip's first unit, nothing behind a unit of the slow path (it writes its own bytes later)
                uint32_t before = __shfl_up_sync(0xffffffffu, spill_out, 1);
                if (h == 0) {
                    const uint32_t carried =
                        *reinterpret_cast<const uint32_t *>(stage + (t0 & ~3u)) & ((1u << (8u * (t0 & 3u))) - 1u);
                    if (lane == 0) before = carried;
                } else if (lane == 0) {
                    before = last_spill;
                }
                uint32_t *dst = reinterpret_cast<uint32_t *>(stage + (o & ~3u));
                // the spill word first: if the next unit is plain too, its first word (the same word, merged with this
                // spill) is stored behind the __syncwarp and stands
                if (p) dst[4] = rot[4];
                __syncwarp();
                if (p) {
                    dst[0] = rot[0] | before;
                    dst[1] = rot[1];
                    dst[2] = rot[2];
                    dst[3] = rot[3];
                    mism += n_del[h];
                } else if (!dead[h]) {
                    exclbuf[h * 32 + lane] = (uint16_t)excl[h];
                }
                if (h + 1 < kU) last_spill = __shfl_sync(0xffffffffu, spill_out, 31);
            }
            __syncwarp();

            // ---- slow path: the pieces of the other units go through the generic comma slots --------------------
            auto translate_pass = [&](auto kw_tag) {
                constexpr int kW = decltype(kw_tag)::value;
                constexpr int kPer = 12 / kW;
#pragma unroll 1
                for (int j0 = 0; j0 < n_pieces; j0 += 32) {
                    const int j = j0 + lane;
                    if (j < n_pieces) {
                        const int ui = kPer == 2 ? j >> 1 : (int)(((uint32_t)j * 43691u) >> 17);
                        const int c = j - kPer * ui;
                        const int u = list[ui];
                        const uint8_t *cp = sb + 16 + dju::kUnitBytes * u + 4 * kW * c;
                        uint32_t w[kW];
                        load_piece<kW>(cp, w);
                        const uint32_t wp = *reinterpret_cast<const uint32_t *>(cp - 4);
                        const uint32_t zz = j0 == 0 ? zz_first : piece_commas_exact<kW>(w, high);
                        const uint32_t cb = reinterpret_cast<const uint32_t *>(cntbuf)[u];
                        const uint32_t rel =
                            exclbuf[u] + (c == 0 ? 0u : (c == 1 ? cb & 0xFFu : (cb & 0xFFu) + ((cb >> 8) & 0xFFu)));
                        uint32_t acc = 0;
                        const uint32_t wrote = emit_piece<kW>(acc, stage_s + t0 + rel, lut_s, wp, w, zz, c_four, c_shift);
                        // two commas among bytes 1..3 of a word (an empty or one-character token) are not MIDSV: bits
                        // 0..3; the table's "special" counter: bits 8..12
                        sticky |= (wrote ^ (uint32_t)__popc(zz)) | ((acc >> 15) & 0x1F00u);
                        mism += (acc >> kStatShift) & 0x3Fu;
                        n_sub += (acc >> 18) & 0x1Fu;
                    }
                }
            };
            if (per_unit == 3) translate_pass(std::integral_constant<int, 4>{});
            else translate_pass(std::integral_constant<int, 6>{});
            __syncwarp();

            // ---- complete 32-byte sectors leave; the incomplete one moves to the front of the staging buffer ---
            const int staged = (int)t0 + total;
            const int flush_bytes = staged & ~31;
#pragma unroll 1
            for (int fb = lane * 16; fb < flush_bytes; fb += 512)
                *reinterpret_cast<uint4 *>(outp + (fb - lane * 16)) = *reinterpret_cast<const uint4 *>(stage + fb);
            const uint8_t b = stage[flush_bytes + lane];  // lanes past the tail read stale bytes, never stored
            carry_w = *reinterpret_cast<const uint32_t *>(sb + 16 + kTrip - 4);
            __syncwarp();
            if (lane < staged - flush_bytes) stage[lane] = b;
            tokbase += total;
            outp += flush_bytes;
#if DJ_ENC_BULK
            if (edge_trip) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // padded words were written
#endif
            __syncwarp();
        }

        if (!broken) {  // what is left is less than one sector: pad it (deterministically) and store it
            const int staged = tokbase & 31;
            if (lane < 32 - staged) stage[staged + lane] = 0;  // reads as a plain match: dj_hist looks at every byte of the pitch
            __syncwarp();
            if (staged > 0 && lane < 2 && (tokbase & ~31) < code_pitch)
                *reinterpret_cast<uint4 *>(outp) = *reinterpret_cast<const uint4 *>(stage + lane * 16);
        }

        mism = __reduce_add_sync(0xffffffffu, mism);
        n_sub = __reduce_add_sync(0xffffffffu, n_sub);
        sticky = __reduce_or_sync(0xffffffffu, sticky | (high & 0x80808080u));
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
    if (code_pitch < n_loci || (code_pitch & 31) != 0) {
        dj_set_error("dj_encode: code_pitch %d must be a multiple of 32 and >= n_loci %d", code_pitch, n_loci);
        return 1;
    }
    const int sms = dj_sm_count();
    if (sms <= 0) return 1;
    static bool configured = false;
    if (!configured) {
        DJ_CUDA_CHECK(cudaFuncSetAttribute(encode_unit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        configured = true;
    }
    int64_t ctas = (n_reads + kWarps - 1) / kWarps;
    if (ctas > sms) ctas = sms;  // persistent: one CTA of kWarps warps per SM
    encode_unit_kernel<<<(unsigned)ctas, kWarps * 32, kSmemBytes, (cudaStream_t)stream>>>(
        text, row_off, row_len, n_reads, n_loci, code_pitch, ckpt_pitch, codes, match_score, n_tokens, flags, ckpt, 4u,
        (uint32_t)kLutSize, dju::Muls{1u << 28, 10u << 28});
    DJ_LAUNCH_CHECK();
    return 0;
}

// Round 1 had a second, reference-assisted encoder behind this entry point (plain stretches verified against the
// allele sequence rendered as text).  The unit encoder verifies plain stretches against their own periodic structure
// at a third of the instructions and needs no sequence, so the entry point is kept for ABI compatibility and calls
// dj_encode: same output for any input, with or without ref_seq.
extern "C" int64_t dj_encode_ref_smem_bytes(int32_t n_loci, int32_t code_pitch) {
    (void)n_loci;
    (void)code_pitch;
    return (int64_t)kSmemBytes;
}

extern "C" int dj_encode_ref(const uint8_t *text, const int64_t *row_off, const int32_t *row_len, int64_t n_reads,
                             int32_t n_loci, int32_t code_pitch, int32_t ckpt_pitch, const uint8_t *ref_seq,
                             uint8_t *codes, int32_t *match_score, int32_t *n_tokens, uint32_t *flags, uint32_t *ckpt,
                             void *stream) {
    (void)ref_seq;
    return dj_encode(text, row_off, row_len, n_reads, n_loci, code_pitch, ckpt_pitch, codes, match_score, n_tokens,
                     flags, ckpt, stream);
}

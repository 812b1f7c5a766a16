// Per-read scans of the code matrix that the reference makes as further passes over the MIDSV text:
//
//   dj_read_n_features   N count and longest N run of every read -- the two features of the sequence-error read
//                        filter (convert_nm_tag / extract_n_features,
//                        src/DAJIN2/core/preprocess/error_correction/sequence_error_handler.py:26-41) and the N
//                        count of the consensus stage's control filter
//                        (consensus/consensus_mutation_analyzer.py:85-88);
//   dj_sv_scan           start index and size of every structural-variant feature of every read
//                        (get_index_and_sv_size, preprocess/structural_variants/sv_detector.py:53-97): runs of >= 10
//                        deletions on loci that carry '-' in mutation_loci, runs of >= 10 lowercase (inverted)
//                        tokens, insertions of >= 10 bases on loci that carry '+'.  The reference makes twelve passes
//                        over the two files for this (three SV types x {collect indices, extract features} x
//                        {sample, control}); the records of ONE scan serve all of them.
//
// Both kernels are one warp per read, four 512-token trips in flight: token classes are found sixteen code bytes at a
// time with byte-parallel range tests, rows become bit strings in shared memory (one bit per locus and class) only
// where a class occurs, and runs are found on the words of those bit strings -- no per-token loop anywhere.  Reads are independent; HBM traffic is the code matrix once
// (algorithmic bytes per read: code_pitch).
#include "dj_common.cuh"

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr uint32_t kOnes = 0x01010101u;
constexpr uint32_t kHigh = 0x80808080u;

// 0x80 in every byte of y (7-bit values) that lies in [lo, hi)
template <int lo, int hi>
__device__ __forceinline__ uint32_t bytes_in_range(uint32_t y) {
    const uint32_t ge = y + (uint32_t)(0x80 - lo) * kOnes;
    const uint32_t lt = ~(y + (uint32_t)(0x80 - hi) * kOnes);
    return ge & lt & kHigh;
}

// bit j = bit 7 of byte j
__device__ __forceinline__ uint32_t movemask4(uint32_t z) { return ((z >> 7) * 0x01020408u) >> 24; }

enum { kClassN = 0, kClassDel = 1, kClassLower = 2, kClassIns = 3 };

// bit 7 of each byte: the token class tests on four code bytes
__device__ __forceinline__ uint32_t class_flags(uint32_t w, int cls) {
    const uint32_t y = w & 0x7F7F7F7Fu;
    switch (cls) {
        case kClassN:  // is_n_tag (utils/midsv_handler.py:8-14): anchor "=N" / "=n", inserted or not
            return bytes_in_range<4, 5>(y) | bytes_in_range<9, 10>(y);
        case kClassDel:  // startswith("-"): a deletion anchor that is not an insertion's anchor
            return bytes_in_range<10, 20>(y) & ~w;
        case kClassLower:  // str.islower(): the lowercase anchors (an insertion takes the case of its anchor)
            return bytes_in_range<5, 10>(y) | bytes_in_range<15, 20>(y) | bytes_in_range<45, 70>(y);
        default:  // startswith("+")
            return w & kHigh;
    }
}

__device__ __forceinline__ uint4 class_flags4(const uint4 &v, int cls) {
    return make_uint4(class_flags(v.x, cls), class_flags(v.y, cls), class_flags(v.z, cls), class_flags(v.w, cls));
}

__device__ __forceinline__ uint32_t movemask16(const uint4 &f) {
    return movemask4(f.x) | (movemask4(f.y) << 4) | (movemask4(f.z) << 8) | (movemask4(f.w) << 12);
}

// 0x80 in every byte of w that is not a plain match (code byte >= 4)
__device__ __forceinline__ uint32_t nonplain_flags(uint32_t w) {
    return (((w | kHigh) - 0x04040404u) | w) & kHigh;
}

// Bit strings of one row: bits[c - kFirst][k] bit b = token 32 * k + b belongs to class c.  `words` words per class plus two
// zero words behind them.  Returns, per class, whether any bit is set (warp-uniform).
template <int kFirst, int kLast>
__device__ __forceinline__ uint32_t build_bits(const uint8_t *__restrict__ row, int code_pitch, int n_loci, int words,
                                               uint32_t *bits, int stride, int lane) {
    constexpr int kInFlight = 4;  // 512-token trips loaded before the first is looked at (2 KB per warp in flight)
    uint32_t any = 0;
    const int n_tokens = words * 32;
    for (int base0 = 0; base0 < n_tokens; base0 += 512 * kInFlight) {
        uint4 v[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            const int pos = base0 + u * 512 + lane * 16;
            v[u] = make_uint4(0, 0, 0, 0);
            if (pos < code_pitch) v[u] = __ldg(reinterpret_cast<const uint4 *>(row + pos));
        }
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            const int base = base0 + u * 512;
            if (base >= n_tokens) break;
            const int pos = base + lane * 16;
            // tokens past n_loci (row padding) never count
            uint32_t valid = 0xFFFFu;
            if (pos + 16 > n_loci) valid = pos >= n_loci ? 0u : (0xFFFFu >> (pos + 16 - n_loci));
            const bool plain = ((v[u].x | v[u].y | v[u].z | v[u].w) & 0xFCFCFCFCu) == 0;  // plain matches: no class
            const int k = (base >> 5) + (lane >> 1);
            if (__ballot_sync(0xffffffffu, !plain) == 0) {  // nothing but plain matches in these 512 tokens
#pragma unroll
                for (int c = kFirst; c <= kLast; ++c)
                    if ((lane & 1) == 0) bits[(c - kFirst) * stride + k] = 0u;
                continue;
            }
#pragma unroll
            for (int c = kFirst; c <= kLast; ++c) {
                uint4 f = make_uint4(0, 0, 0, 0);
                if (!plain) f = class_flags4(v[u], c);
                if (__ballot_sync(0xffffffffu, (f.x | f.y | f.z | f.w) != 0) == 0) {  // no token of this class here
                    if ((lane & 1) == 0) bits[(c - kFirst) * stride + k] = 0u;
                    continue;
                }
                const uint32_t m = movemask16(f) & valid;
                const uint32_t hi = __shfl_down_sync(0xffffffffu, m, 1);
                if ((lane & 1) == 0) bits[(c - kFirst) * stride + k] = m | (hi << 16);
                any |= (__ballot_sync(0xffffffffu, m != 0) != 0 ? 1u : 0u) << c;
            }
        }
    }
    return any;
}

struct RunSummary {  // of a bit string: leading run (from bit 0), trailing run, longest run, length
    int pre, suf, best, len;
};

__device__ __forceinline__ RunSummary combine(const RunSummary &a, const RunSummary &b) {  // a then b
    RunSummary r;
    r.pre = a.pre == a.len ? a.len + b.pre : a.pre;
    r.suf = b.suf == b.len ? b.len + a.suf : b.suf;
    r.best = max(max(a.best, b.best), a.suf + b.pre);
    r.len = a.len + b.len;
    return r;
}

__device__ __forceinline__ RunSummary summarize_word(uint32_t w) {
    RunSummary s;
    s.len = 32;
    if (w == 0u) {
        s.pre = s.suf = s.best = 0;
    } else if (w == 0xFFFFFFFFu) {
        s.pre = s.suf = s.best = 32;
    } else {
        s.pre = __ffs(~w) - 1;
        s.suf = __clz(~w);
        int k = 0;
        for (uint32_t t = w; t; t &= t << 1) ++k;
        s.best = k;
    }
    return s;
}

__global__ void __launch_bounds__(kScanThreads)
n_features_kernel(const uint8_t *__restrict__ codes, int code_pitch, const int32_t *__restrict__ rows, int64_t n_rows,
                  int n_loci, int words, int32_t *__restrict__ n_count, int32_t *__restrict__ n_maxrun) {
    extern __shared__ uint32_t smem_bits[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = words + 2;
    uint32_t *bits = smem_bits + warp * stride;
    const int64_t n_warps = (int64_t)gridDim.x * kScanWarps;
    for (int64_t t = (int64_t)blockIdx.x * kScanWarps + warp; t < n_rows; t += n_warps) {
        const int64_t r = rows ? (int64_t)__ldg(rows + t) : t;
        const uint32_t any = build_bits<kClassN, kClassN>(codes + r * (int64_t)code_pitch, code_pitch, n_loci, words,
                                                          bits, stride, lane);
        int count = 0, best = 0;
        if (any) {
            __syncwarp();
            RunSummary acc = {0, 0, 0, 0};
            for (int k0 = 0; k0 < words; k0 += 32) {
                const uint32_t w = k0 + lane < words ? bits[k0 + lane] : 0u;
                count += __popc(w);
                RunSummary s = summarize_word(w);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {  // ordered reduction: lane i holds lanes i .. i + 2d - 1
                    RunSummary o;
                    o.pre = __shfl_down_sync(0xffffffffu, s.pre, d);
                    o.suf = __shfl_down_sync(0xffffffffu, s.suf, d);
                    o.best = __shfl_down_sync(0xffffffffu, s.best, d);
                    o.len = __shfl_down_sync(0xffffffffu, s.len, d);
                    if (lane + d < 32) s = combine(s, o);
                }
                acc = combine(acc, s);  // lane 0's value is the block's; the others are not used
            }
            count = __reduce_add_sync(0xffffffffu, count);
            best = __shfl_sync(0xffffffffu, acc.best, 0);
            __syncwarp();
        }
        if (lane == 0) {
            n_count[t] = count;
            n_maxrun[t] = best;
        }
    }
}

// Length in bytes of token `idx` of read r, found by the whole warp (every lane passes the same arguments): the
// checkpoint table gives the 1 KB segment the token starts in, its commas are located 32 bytes per lane.  The serial
// dj_find_token costs one DRAM round trip per 32-byte sector; this costs two per token.  Returns -1 if absent.
__device__ int warp_token_length(const DjTextRef &tr, int64_t r, int idx, int lane) {
    const int64_t off = __ldg(tr.row_off + r);
    const int len = __ldg(tr.row_len + r);
    const int head = (int)(off & 15);
    const int n_iter = (head + len + DJ_ENC_CHUNK - 1) / DJ_ENC_CHUNK;
    const uint32_t *ck = tr.ckpt + r * (int64_t)tr.ckpt_pitch;
    int lo = 0;  // the last segment whose checkpoint is <= idx (checkpoints do not decrease)
    for (int j0 = 0; j0 < n_iter; j0 += 32) {
        const int j = j0 + lane;
        const uint32_t le = __ballot_sync(0xffffffffu, j < n_iter && (int)__ldg(ck + j) <= idx);
        if (le == 0u) break;
        lo = j0 + 31 - __clz(le);
        if (le != 0xffffffffu) break;
    }
    const int first = max(lo * DJ_ENC_CHUNK - head, 0);  // row-relative byte the search starts at
    const uint8_t *row = tr.text + off;
    // a token starts at `first` when the byte in front of it is a comma (or `first` is the row start)
    const int starts_here = (first == 0 || __ldg(row + first - 1) == ',') ? 1 : 0;
    // commas at positions >= first are numbered 0, 1, ...; the token ends at comma number `end_no` and starts behind
    // comma number end_no - 1 (at `first` itself when end_no == 0, which requires starts_here)
    const int end_no = idx - (int)__ldg(ck + lo) - starts_here + 1;
    if (end_no < 0 || (end_no == 0 && !starts_here)) return -1;
    int start = end_no == 0 ? first : -1, seen = 0;
    const uint8_t *aligned = row - head;  // 16-byte aligned
    for (int w0 = (first + head) & ~15; w0 < head + len; w0 += 1024) {  // offsets relative to `aligned`
        uint32_t commas = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int p = w0 + lane * 32 + h * 16;  // this lane's sixteen bytes
            if (p >= head + len) continue;
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(aligned + p));
            const uint32_t word[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t x = word[k] ^ 0x2C2C2C2Cu;
                const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & kHigh;  // 0x80 where the byte is ','
                commas |= movemask4(z) << (h * 16 + k * 4);
            }
        }
        // only bytes of the row from `first` on count
        const int lane0 = w0 + lane * 32 - head;  // row-relative position of this lane's first byte
        if (lane0 < first) commas &= first - lane0 >= 32 ? 0u : 0xFFFFFFFFu << (first - lane0);
        if (lane0 + 32 > len) commas &= lane0 >= len ? 0u : 0xFFFFFFFFu >> (lane0 + 32 - len);
        const int mine = __popc(commas);
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        const int before = seen + incl - mine;  // number of the first comma of this lane
        int found_start = -1, found_end = -1;
        if (end_no - 1 >= before && end_no - 1 < before + mine) found_start = lane0 + (int)__fns(commas, 0, end_no - before) + 1;
        if (end_no >= before && end_no < before + mine) found_end = lane0 + (int)__fns(commas, 0, end_no - before + 1);
        const uint32_t vs = __ballot_sync(0xffffffffu, found_start >= 0), ve = __ballot_sync(0xffffffffu, found_end >= 0);
        if (vs) start = __shfl_sync(0xffffffffu, found_start, __ffs(vs) - 1);
        if (ve) return __shfl_sync(0xffffffffu, found_end, __ffs(ve) - 1) - start;
        seen += __shfl_sync(0xffffffffu, incl, 31);
    }
    return start >= 0 ? len - start : -1;  // the last token of the row has no closing comma
}

__device__ __forceinline__ void emit_record(int32_t *records, uint32_t *n_records, uint32_t capacity, int type,
                                            int read, int start, int size) {
    const uint32_t slot = atomicAdd(n_records, 1u);
    if (slot < capacity) {
        int32_t *p = records + (size_t)slot * 4;
        p[0] = type;
        p[1] = read;
        p[2] = start;
        p[3] = size;
    }
}

__global__ void __launch_bounds__(kScanThreads)
sv_scan_kernel(const uint8_t *__restrict__ codes, int code_pitch, const int32_t *__restrict__ rows, int64_t n_rows,
               int n_loci, int words, const uint32_t *__restrict__ del_loci, const uint32_t *__restrict__ ins_loci,
               int min_size, DjTextRef text, int32_t *__restrict__ records, uint32_t *__restrict__ n_records,
               uint32_t capacity) {
    extern __shared__ uint32_t smem_bits[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int stride = words + 2;
    uint32_t *bits = smem_bits + warp * 2 * stride;  // deletion and lowercase classes
    if (lane < 4) bits[(lane >> 1) * stride + words + (lane & 1)] = 0u;  // the zero words behind both bit strings
    const int64_t n_warps = (int64_t)gridDim.x * kScanWarps;
    for (int64_t t = (int64_t)blockIdx.x * kScanWarps + warp; t < n_rows; t += n_warps) {
        const int64_t r = rows ? (int64_t)__ldg(rows + t) : t;
        const uint8_t *row = codes + r * (int64_t)code_pitch;
        // Pass A, every read: (i) insertion tokens on loci that hold '+' (:76-78) are looked up at once; (ii) a run of
        // seven or more deletions / lowercase tokens covers an aligned group of four code bytes none of which is a
        // plain match, so reads without such a group (all but the SV-carrying and N-tailed ones) are finished here.
        bool has_full = min_size < 7;
        constexpr int kInFlight = 4;
        for (int base0 = 0; base0 < words * 32; base0 += 512 * kInFlight) {
            uint4 v[kInFlight];
#pragma unroll
            for (int u = 0; u < kInFlight; ++u) {
                const int pos = base0 + u * 512 + lane * 16;
                v[u] = make_uint4(0, 0, 0, 0);
                if (pos < code_pitch) v[u] = __ldg(reinterpret_cast<const uint4 *>(row + pos));
            }
#pragma unroll
            for (int u = 0; u < kInFlight; ++u) {
                const int pos = base0 + u * 512 + lane * 16;
                if (base0 + u * 512 >= words * 32) break;
                const bool full = nonplain_flags(v[u].x) == kHigh || nonplain_flags(v[u].y) == kHigh ||
                                  nonplain_flags(v[u].z) == kHigh || nonplain_flags(v[u].w) == kHigh;
                uint32_t cand = 0;
                if (((v[u].x | v[u].y | v[u].z | v[u].w) & kHigh) != 0)
                    cand = (__ldg(ins_loci + (pos >> 5)) >> (pos & 31)) & 0xFFFFu;  // loci past n_loci are never set
                const uint32_t votes = __ballot_sync(0xffffffffu, full);
                has_full = has_full || votes != 0;
                if (cand) cand &= movemask16(make_uint4(v[u].x & kHigh, v[u].y & kHigh, v[u].z & kHigh, v[u].w & kHigh));
                // candidates are rare: the warp looks each one up together
                for (uint32_t busy = __ballot_sync(0xffffffffu, cand != 0); busy; busy = __ballot_sync(0xffffffffu, cand != 0)) {
                    const int src = __ffs(busy) - 1;
                    const int locus = __shfl_sync(0xffffffffu, pos + __ffs(cand) - 1, src);
                    if (lane == src) cand &= cand - 1;
                    const int len = warp_token_length(text, r, locus, lane);
                    if (len < 0 || lane != src) continue;
                    const int anchor = row[locus] & 0x7F;
                    const int n_ins = (len - (anchor < 20 ? 2 : 3)) / 3;  // every inserted base is "+B|"
                    if (n_ins >= min_size) emit_record(records, n_records, capacity, 2, (int)t, locus, n_ins);
                }
            }
        }
        if (!has_full) continue;
        // Pass B, the few reads left: bit strings of the deletion and lowercase classes, then their runs.
        const uint32_t any = build_bits<kClassDel, kClassLower>(row, code_pitch, n_loci, words, bits, stride, lane);
        if (any == 0) continue;
        __syncwarp();
        // deletions only count on loci whose mutation_loci holds '-' (sv_detector.py:67)
        if (any & (1u << kClassDel))
            for (int k = lane; k < words; k += 32) bits[k] &= __ldg(del_loci + k);
        __syncwarp();
#pragma unroll 1
        for (int c = kClassDel; c <= kClassLower; ++c) {  // runs of at least min_size tokens (:65-74, :80-82)
            if (!(any & (1u << c))) continue;
            const uint32_t *b = bits + (c - kClassDel) * stride;
            for (int k = lane; k < words; k += 32) {
                const uint32_t lo = b[k];
                if (lo == 0u) continue;
                // starts of runs inside this word: a set bit whose predecessor is clear
                const uint32_t prev = k > 0 ? b[k - 1] >> 31 : 0u;
                uint32_t starts = lo & ~((lo << 1) | prev);
                while (starts) {
                    const int bit = __ffs(starts) - 1;
                    starts &= starts - 1;
                    int len = 0, kk = k, bb = bit;  // length of the run that starts here
                    for (;;) {
                        const uint32_t w = ~(b[kk] >> bb);  // zero word behind the row ends every run
                        const int ones = w ? __ffs(w) - 1 : 32;
                        const int room = 32 - bb;
                        if (ones < room) {
                            len += ones;
                            break;
                        }
                        len += room;
                        ++kk;
                        bb = 0;
                    }
                    if (len >= min_size) emit_record(records, n_records, capacity, c == kClassDel ? 0 : 1, (int)t,
                                                     k * 32 + bit, len);
                }
            }
        }
        __syncwarp();
    }
}

int scan_geometry(const char *who, int32_t code_pitch, int32_t n_loci, int classes, int *words, size_t *smem) {
    if (n_loci <= 0 || code_pitch < n_loci || (code_pitch & 15) != 0) {
        dj_set_error("%s: code_pitch %d must be a multiple of 16 and >= n_loci %d > 0", who, code_pitch, n_loci);
        return 1;
    }
    *words = ((n_loci + 511) / 512) * 16;  // whole 512-token trips
    *smem = (size_t)kScanWarps * classes * (*words + 2) * sizeof(uint32_t);
    if (*smem > 200 * 1024) {
        dj_set_error("%s: n_loci %d is beyond the %d loci the shared-memory bit strings hold", who, n_loci,
                     (int)(200 * 1024 / (kScanWarps * classes * 4) - 2) * 32);
        return 1;
    }
    return 0;
}

}  // namespace

extern "C" int dj_read_n_features(const uint8_t *codes, int32_t code_pitch, const int32_t *rows, int64_t n_rows,
                                  int32_t n_loci, int32_t *n_count, int32_t *n_maxrun, void *stream) {
    if (n_rows <= 0) return 0;
    int words = 0;
    size_t smem = 0;
    if (scan_geometry("dj_read_n_features", code_pitch, n_loci, 1, &words, &smem)) return 1;
    if (((uintptr_t)codes & 15) != 0) {
        dj_set_error("dj_read_n_features: the code matrix must be 16-byte aligned");
        return 1;
    }
    const int sms = dj_sm_count();
    if (sms <= 0) return 1;
    static size_t configured = 0;
    if (configured < smem) {
        DJ_CUDA_CHECK(cudaFuncSetAttribute(n_features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int64_t ctas = (n_rows + kScanWarps - 1) / kScanWarps;
    if (ctas > (int64_t)sms * 8) ctas = (int64_t)sms * 8;
    n_features_kernel<<<(unsigned)ctas, kScanThreads, smem, (cudaStream_t)stream>>>(codes, code_pitch, rows, n_rows,
                                                                                  n_loci, words, n_count, n_maxrun);
    DJ_LAUNCH_CHECK();
    return 0;
}

extern "C" int dj_sv_scan(const uint8_t *codes, int32_t code_pitch, const int32_t *rows, int64_t n_rows,
                          int32_t n_loci, const uint8_t *text, const int64_t *row_off, const int32_t *row_len,
                          const uint32_t *ckpt, int32_t ckpt_pitch, const uint32_t *del_loci, const uint32_t *ins_loci,
                          int32_t min_size, int32_t *records, uint32_t capacity, uint32_t *n_records, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    DJ_CUDA_CHECK(cudaMemsetAsync(n_records, 0, sizeof(uint32_t), st));
    if (n_rows <= 0) return 0;
    int words = 0;
    size_t smem = 0;
    if (scan_geometry("dj_sv_scan", code_pitch, n_loci, 2, &words, &smem)) return 1;
    if (((uintptr_t)codes & 15) != 0) {
        dj_set_error("dj_sv_scan: the code matrix must be 16-byte aligned");
        return 1;
    }
    const int sms = dj_sm_count();
    if (sms <= 0) return 1;
    static size_t configured = 0;
    if (configured < smem) {
        DJ_CUDA_CHECK(cudaFuncSetAttribute(sv_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int64_t ctas = (n_rows + kScanWarps - 1) / kScanWarps;
    if (ctas > (int64_t)sms * 8) ctas = (int64_t)sms * 8;
    DjTextRef tr{text, row_off, row_len, ckpt, ckpt_pitch};
    sv_scan_kernel<<<(unsigned)ctas, kScanThreads, smem, st>>>(codes, code_pitch, rows, n_rows, n_loci, words, del_loci,
                                                              ins_loci, min_size, tr, records, n_records, capacity);
    DJ_LAUNCH_CHECK();
    return 0;
}

extern "C" int32_t dj_loci_words(int32_t n_loci) { return ((n_loci + 511) / 512) * 16; }

// Shared device/host helpers for the dajin2_b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dajin2_b200.h"

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
void dj_set_error(const char *fmt, ...);

#define DJ_CUDA_CHECK(expr)                                                                    \
    do {                                                                                       \
        cudaError_t err_ = (expr);                                                             \
        if (err_ != cudaSuccess) {                                                             \
            dj_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err_), __FILE__, __LINE__); \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

#define DJ_LAUNCH_CHECK() DJ_CUDA_CHECK(cudaGetLastError())

int dj_sm_count();
int dj_scratch_pool_init();  // 0 on success; call before cudaMallocAsync

// ---------------------------------------------------------------------------------------------------------
// code byte arithmetic (see include/dajin2_b200.h for the layout)
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int dj_code_opmask(int code) {
    if (code & DJ_CODE_INS) return DJ_MASK_INS;
    int a = code & 0x7F;
    if (a < 10) return 0;
    if (a < 20) return DJ_MASK_DEL;
    if (a < 70) return DJ_MASK_SUB;
    return 0;
}

__host__ __device__ __forceinline__ bool dj_anchor_is_lower(int a) {
    if (a < 10) return a >= 5;
    if (a < 20) return a >= 15;
    if (a < 70) return a >= 45;
    return false;
}

__host__ __device__ __forceinline__ bool dj_anchor_is_n(int a) { return a == 4 || a == 9; }

// histogram class with the reference's skip rule (indel_counter.py:20-30): 0 '=', 1 '+', 2 '-', 3 '*', 4 none
__host__ __device__ __forceinline__ int dj_code_hist_class(int code) {
    int a = code & 0x7F;
    if (a >= 70) return 4;
    if (dj_anchor_is_n(a) || dj_anchor_is_lower(a)) return 4;
    if (code & DJ_CODE_INS) return 1;
    if (a < 10) return 0;
    if (a < 20) return 2;
    return 3;
}

// startswith class without skip rule (error_correction/strand_bias_handler.py:25-28): 0 '+', 1 '-', 2 '*', 3 none
__host__ __device__ __forceinline__ int dj_code_startswith_class(int code) {
    if ((code & 0x7F) >= 70) return 3;
    if (code & DJ_CODE_INS) return 0;
    int a = code & 0x7F;
    if (a < 10) return 3;
    if (a < 20) return 1;
    return 2;
}

// calc_match contribution of a non-insertion token (classifier.py:19-22: '+','-','*' chars + lowercase chars)
__host__ __device__ __forceinline__ int dj_anchor_mismatch(int a) {
    if (a < 10) return a >= 5 ? 1 : 0;   // "=x" has one lowercase char
    if (a < 20) return a >= 15 ? 2 : 1;  // '-' (+ lowercase base)
    if (a < 70) return a >= 45 ? 3 : 1;  // '*' (+ two lowercase bases)
    return 0;
}

__host__ __device__ __forceinline__ char dj_base_char(int b, bool lower) {
    const char up[5] = {'A', 'C', 'G', 'T', 'N'};
    char c = up[b];
    return lower ? (char)(c | 0x20) : c;
}

// ---------------------------------------------------------------------------------------------------------
// hashing
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t dj_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__host__ __device__ __forceinline__ uint64_t dj_hash_combine(uint64_t h, uint64_t v) {
    return dj_mix64(h ^ (v + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2)));
}

// ---------------------------------------------------------------------------------------------------------
// raw-token access through the encoder's checkpoints
// ---------------------------------------------------------------------------------------------------------
struct DjTextRef {
    const uint8_t *text;
    const int64_t *row_off;
    const int32_t *row_len;
    const uint32_t *ckpt;
    int32_t ckpt_pitch;
};

// Locate token `idx` of read `r`: returns byte offset (relative to the row start) and length; false if absent.
__device__ inline bool dj_find_token(const DjTextRef &tr, int64_t r, int idx, int *tok_off, int *tok_len) {
    const int64_t off = tr.row_off[r];
    const int len = tr.row_len[r];
    const int head = (int)(off & 15);
    const int n_iter = (head + len + DJ_ENC_CHUNK - 1) / DJ_ENC_CHUNK;
    const uint32_t *ck = tr.ckpt + r * (int64_t)tr.ckpt_pitch;
    int lo = 0, hi = n_iter - 1;  // largest it with ck[it] <= idx
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if ((int)ck[mid] <= idx) lo = mid; else hi = mid - 1;
    }
    int p = lo * DJ_ENC_CHUNK - head;
    int count = (int)ck[lo];
    if (p < 0) p = 0;
    const uint8_t *row = tr.text + off;
    for (; p < len; ++p) {
        bool start = (p == 0) || (row[p - 1] == ',');
        if (!start) continue;
        if (count == idx) {
            int q = p;
            while (q < len && row[q] != ',') ++q;
            *tok_off = p;
            *tok_len = q - p;
            return true;
        }
        ++count;
    }
    return false;
}

__device__ inline uint64_t dj_hash_token(const DjTextRef &tr, int64_t r, int idx) {
    int o, n;
    if (!dj_find_token(tr, r, idx, &o, &n)) return 0x1234567ULL;
    const uint8_t *p = tr.text + tr.row_off[r] + o;
    uint64_t h = 0xcbf29ce484222325ULL ^ (uint64_t)n;
    for (int k = 0; k < n; ++k) h = (h ^ p[k]) * 0x100000001b3ULL;
    return dj_mix64(h);
}

// ---------------------------------------------------------------------------------------------------------
// 3-mer key (clustering/kmer_generator.py:35-56): "prev,cur,next" with '@' for tokens that are not selected
// ---------------------------------------------------------------------------------------------------------
#define DJ_KMER_AT 256
#define DJ_KMER_RAW 257

struct DjKmerKey {
    bool at;           // the k-mer is "@,@,@"
    uint64_t payload;  // sel_index << 32 | prev << 20 | cur << 10 | next
    uint64_t fp;       // fingerprint used for probing (top bit set, never 0)
};

__device__ inline DjKmerKey dj_kmer_key(const uint8_t *row_codes, int i, const uint8_t *locimask, int sel_index,
                                        const DjTextRef &tr, int64_t read) {
    DjKmerKey key;
    key.at = true;
    key.payload = 0;
    key.fp = 0;
    const int cur = row_codes[i];
    if (!(dj_code_opmask(cur) & locimask[i])) return key;
    const int prev = row_codes[i - 1], next = row_codes[i + 1];
    const bool cur_ins = (cur & DJ_CODE_INS) != 0;
    int prev_repr = DJ_KMER_AT, next_repr = DJ_KMER_AT;
    uint64_t raw = 0;
    if (dj_code_opmask(prev) & locimask[i - 1]) {
        prev_repr = prev;
        // a visible previous insertion was squashed to "+I<anchor>" when it was the centre (kmer_generator.py:
        // 45-47) -- except token 0, which is never a centre: it stays raw unless this centre is an insertion.
        if ((prev & DJ_CODE_INS) && i - 1 == 0 && !cur_ins) {
            prev_repr = DJ_KMER_RAW;
            raw = dj_hash_combine(raw ^ 0x51ULL, dj_hash_token(tr, read, i - 1));
        }
    }
    if (dj_code_opmask(next) & locimask[i + 1]) {
        next_repr = next;
        // the next token is still the raw insertion string unless the centre itself is an insertion
        if ((next & DJ_CODE_INS) && !cur_ins) {
            next_repr = DJ_KMER_RAW;
            raw = dj_hash_combine(raw ^ 0xA7ULL, dj_hash_token(tr, read, i + 1));
        }
    }
    key.at = false;
    key.payload = ((uint64_t)(uint32_t)sel_index << 32) | ((uint64_t)prev_repr << 20) | ((uint64_t)cur << 10) |
                  (uint64_t)next_repr;
    key.fp = dj_hash_combine(dj_mix64(key.payload + 0x2545F4914F6CDD1DULL), raw) | 0x8000000000000000ULL;
    return key;
}

// ---------------------------------------------------------------------------------------------------------
// open-addressing tables: header (64 B) followed by structure-of-arrays columns
// ---------------------------------------------------------------------------------------------------------
struct DjKmerTableView {
    uint32_t *n_entries;
    unsigned long long *fp;
    unsigned long long *payload;
    uint32_t *cnt0;
    uint32_t *cnt1;
    uint32_t *example;
    uint32_t capacity;
};

__host__ __device__ inline DjKmerTableView dj_kmer_view(void *base, uint32_t capacity) {
    DjKmerTableView v;
    uint8_t *p = (uint8_t *)base;
    v.n_entries = (uint32_t *)p;
    p += 64;
    v.fp = (unsigned long long *)p;
    p += (size_t)capacity * 8;
    v.payload = (unsigned long long *)p;
    p += (size_t)capacity * 8;
    v.cnt0 = (uint32_t *)p;
    p += (size_t)capacity * 4;
    v.cnt1 = (uint32_t *)p;
    p += (size_t)capacity * 4;
    v.example = (uint32_t *)p;
    v.capacity = capacity;
    return v;
}

struct DjRowTableView {
    uint32_t *n_entries;
    unsigned long long *fp;
    uint32_t *count;
    uint32_t *plus;
    uint32_t *first;
    uint32_t *last;
    uint32_t capacity;
};

__host__ __device__ inline DjRowTableView dj_row_view(void *base, uint32_t capacity) {
    DjRowTableView v;
    uint8_t *p = (uint8_t *)base;
    v.n_entries = (uint32_t *)p;
    p += 64;
    v.fp = (unsigned long long *)p;
    p += (size_t)capacity * 8;
    v.count = (uint32_t *)p;
    p += (size_t)capacity * 4;
    v.plus = (uint32_t *)p;
    p += (size_t)capacity * 4;
    v.first = (uint32_t *)p;
    p += (size_t)capacity * 4;
    v.last = (uint32_t *)p;
    v.capacity = capacity;
    return v;
}

// Find or claim the slot of fingerprint fp.  *claimed is set when this thread created the slot.
// Returns capacity when the table is full.
__device__ inline uint32_t dj_table_insert(unsigned long long *fps, uint32_t capacity, unsigned long long fp,
                                           bool *claimed) {
    uint32_t slot = (uint32_t)(dj_mix64(fp) & (capacity - 1));
    *claimed = false;
    for (uint32_t probe = 0; probe < capacity; ++probe) {
        unsigned long long cur = *((volatile unsigned long long *)&fps[slot]);
        if (cur == fp) return slot;
        if (cur == 0ULL) {
            unsigned long long old = atomicCAS(&fps[slot], 0ULL, fp);
            if (old == 0ULL) {
                *claimed = true;
                return slot;
            }
            if (old == fp) return slot;
        }
        slot = (slot + 1) & (capacity - 1);
    }
    return capacity;
}

__device__ inline uint32_t dj_table_find(const unsigned long long *fps, uint32_t capacity, unsigned long long fp) {
    uint32_t slot = (uint32_t)(dj_mix64(fp) & (capacity - 1));
    for (uint32_t probe = 0; probe < capacity; ++probe) {
        unsigned long long cur = fps[slot];
        if (cur == fp) return slot;
        if (cur == 0ULL) return capacity;
        slot = (slot + 1) & (capacity - 1);
    }
    return capacity;
}

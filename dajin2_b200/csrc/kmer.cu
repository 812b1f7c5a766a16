// 3-mer counting at the selected loci (dj_kmer_count), score-row annotation (dj_annotate) and raw token
// retrieval (dj_fetch_tokens).
//
// Counting is staged through shared memory: a selected locus sees a few dozen distinct 3-mers however many reads
// there are, so every CTA of the persistent grid counts into its own 1024-slot table (fingerprint, payload, count,
// smallest read) with shared-memory atomics and adds its entries to the global table ONCE, when it has run out of
// reads: a 25 %-allele locus costs (CTAs x distinct 3-mers) global atomics instead of one per read.  A 3-mer that
// finds no room in the CTA's table (32 probes) goes to the global table directly.
//
// Only tokens whose operation is listed in mutation_loci[i] form a k-mer other than "@,@,@"
// (src/DAJIN2/core/clustering/kmer_generator.py:40-54), so both kernels read three code bytes per
// (read, selected locus) -- lanes run over consecutive selected loci of one read so that neighbouring loci share
// sectors -- and touch the hash table only for mutated tokens.  Raw insertion strings (the reference keeps the
// *uncompressed* next token unless the centre is itself an insertion) are re-read from the text through the
// encoder's checkpoints and enter the key as a 64-bit hash.
#include "dj_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kLocalSlots = 1024;  // per-CTA staging table (24 KB of shared memory)
constexpr int kLocalProbes = 32;

// find or claim the slot of fp in a CTA-local table; kLocalSlots when 32 probes found neither
__device__ __forceinline__ uint32_t local_insert(unsigned long long *fps, unsigned long long fp, bool *claimed) {
    uint32_t slot = (uint32_t)(dj_mix64(fp) & (kLocalSlots - 1));
    *claimed = false;
    for (int probe = 0; probe < kLocalProbes; ++probe) {
        const unsigned long long cur = *((volatile unsigned long long *)&fps[slot]);
        if (cur == fp) return slot;
        if (cur == 0ULL) {
            const unsigned long long old = atomicCAS(&fps[slot], 0ULL, fp);
            if (old == 0ULL) {
                *claimed = true;
                return slot;
            }
            if (old == fp) return slot;
        }
        slot = (slot + 1) & (kLocalSlots - 1);
    }
    return kLocalSlots;
}

__device__ __forceinline__ void global_count(DjKmerTableView &tb, unsigned long long fp, unsigned long long payload,
                                             uint32_t count, uint32_t example, int which, uint32_t *overflow) {
    bool claimed;
    const uint32_t slot = dj_table_insert(tb.fp, tb.capacity, fp, &claimed);
    if (slot == tb.capacity) {
        atomicAdd(overflow, count);
        return;
    }
    if (claimed) {
        tb.payload[slot] = payload;
        atomicAdd(tb.n_entries, 1u);
    }
    atomicAdd(which == 0 ? &tb.cnt0[slot] : &tb.cnt1[slot], count);
    atomicMin(&tb.example[slot], example);
}

__global__ void __launch_bounds__(kThreads)
kmer_count_kernel(const uint8_t *__restrict__ codes, int code_pitch, const int32_t *__restrict__ rows,
                  int64_t n_rows, DjTextRef tr, const uint8_t *__restrict__ locimask,
                  const int32_t *__restrict__ sel, int n_sel, int which, DjKmerTableView tb,
                  uint32_t *__restrict__ overflow) {
    __shared__ unsigned long long l_fp[kLocalSlots], l_payload[kLocalSlots];
    __shared__ uint32_t l_cnt[kLocalSlots], l_example[kLocalSlots];
    for (int e = threadIdx.x; e < kLocalSlots; e += blockDim.x) {
        l_fp[e] = 0ULL;
        l_cnt[e] = 0u;
        l_example[e] = 0xFFFFFFFFu;
    }
    __syncthreads();
    const int64_t n_items = n_rows * (int64_t)n_sel;
    for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < n_items;
         item += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = item / n_sel;
        const int s = (int)(item - t * n_sel);
        const int64_t r = rows ? rows[t] : t;
        const DjKmerKey key = dj_kmer_key(codes + r * (int64_t)code_pitch, sel[s], locimask, s, tr, r);
        if (key.at) continue;
        const uint32_t example = ((uint32_t)which << 31) | (uint32_t)r;
        bool claimed;
        const uint32_t slot = local_insert(l_fp, key.fp, &claimed);
        if (slot == kLocalSlots) {  // no room in this CTA's table
            global_count(tb, key.fp, key.payload, 1u, example, which, overflow);
            continue;
        }
        if (claimed) l_payload[slot] = key.payload;
        atomicAdd(&l_cnt[slot], 1u);
        atomicMin(&l_example[slot], example);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kLocalSlots; e += blockDim.x)
        if (l_fp[e] != 0ULL) global_count(tb, l_fp[e], l_payload[e], l_cnt[e], l_example[e], which, overflow);
}

__global__ void kmer_export_kernel(DjKmerTableView tb, DjKmerRecord *out, uint32_t max_records, uint32_t *n_out) {
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < tb.capacity; slot += gridDim.x * blockDim.x) {
        if (tb.fp[slot] == 0ULL) continue;
        const uint32_t k = atomicAdd(n_out, 1u);
        if (k >= max_records) continue;
        const unsigned long long p = tb.payload[slot];
        DjKmerRecord rec;
        rec.sel_index = (uint32_t)(p >> 32);
        rec.prev = (uint16_t)((p >> 20) & 0x3FF);
        rec.cur = (uint16_t)((p >> 10) & 0x3FF);
        rec.next = (uint16_t)(p & 0x3FF);
        rec.reserved = 0;
        rec.count_sample = tb.cnt0[slot];
        rec.count_control = tb.cnt1[slot];
        rec.example = tb.example[slot];
        rec.slot = slot;
        out[k] = rec;
    }
}

__global__ void __launch_bounds__(kThreads)
annotate_kernel(const uint8_t *__restrict__ codes, int code_pitch, const int32_t *__restrict__ rows,
                int64_t n_rows, DjTextRef tr, const uint8_t *__restrict__ locimask, int n_loci,
                const int32_t *__restrict__ col_locus, const int32_t *__restrict__ col_sel,
                const uint16_t *__restrict__ col_at_id, int n_cols, int is_control, DjKmerTableView tb,
                const uint16_t *__restrict__ slot_value_id, uint16_t *__restrict__ ids) {
    const int64_t n_items = n_rows * (int64_t)n_cols;
    for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < n_items;
         item += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = item / n_cols;
        const int c = (int)(item - t * n_cols);
        const int64_t r = rows ? rows[t] : t;
        const int locus = col_locus[c];
        uint16_t id = col_at_id[c];
        if (col_sel[c] >= 0 && locus >= 1 && locus <= n_loci - 2) {
            const DjKmerKey key = dj_kmer_key(codes + r * (int64_t)code_pitch, locus, locimask, col_sel[c], tr, r);
            if (!key.at) {
                id = 0;
                if (!is_control) {  // control rows never score a selected mutation (score_handler.py:143-144)
                    const uint32_t slot = dj_table_find(tb.fp, tb.capacity, key.fp);
                    if (slot != tb.capacity) id = slot_value_id[slot];
                }
            }
        }
        ids[item] = id;
    }
}

__global__ void fetch_tokens_kernel(DjTextRef tr, const int32_t *__restrict__ reads, const int32_t *__restrict__ loci,
                                    int n, uint8_t *__restrict__ out, int32_t *__restrict__ out_len, int max_len) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int o = 0, len = 0;
    if (!dj_find_token(tr, reads[k], loci[k], &o, &len)) {
        out_len[k] = -1;
        return;
    }
    out_len[k] = len;
    const uint8_t *p = tr.text + tr.row_off[reads[k]] + o;
    for (int j = 0; j < len && j < max_len; ++j) out[(int64_t)k * max_len + j] = p[j];
}

int grid_for(int64_t n_items) {
    int64_t blocks = (n_items + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)dj_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

extern "C" int64_t dj_kmer_table_bytes(uint32_t capacity) { return 64 + (int64_t)capacity * 28; }

extern "C" int dj_kmer_table_clear(void *table, uint32_t capacity, void *stream) {
    if (capacity == 0 || (capacity & (capacity - 1)) != 0) {
        dj_set_error("dj_kmer_table_clear: capacity must be a power of two");
        return 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    DjKmerTableView tb = dj_kmer_view(table, capacity);
    // header, fingerprints, payload and both counters -> 0 ; example -> 0xFFFFFFFF (atomicMin target)
    DJ_CUDA_CHECK(cudaMemsetAsync(table, 0, 64 + (size_t)capacity * 24, st));
    DJ_CUDA_CHECK(cudaMemsetAsync(tb.example, 0xFF, (size_t)capacity * 4, st));
    return 0;
}

extern "C" int dj_kmer_count(const uint8_t *codes, int32_t code_pitch, const int32_t *rows, int64_t n_rows,
                             const uint8_t *text, const int64_t *row_off, const int32_t *row_len,
                             const uint32_t *ckpt, int32_t ckpt_pitch, const uint8_t *locimask, int32_t n_loci,
                             const int32_t *sel, int32_t n_sel, int32_t which, void *table, uint32_t capacity,
                             void *stream) {
    (void)n_loci;
    if (n_rows <= 0 || n_sel <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    DjKmerTableView tb = dj_kmer_view(table, capacity);
    DjTextRef tr{text, row_off, row_len, ckpt, ckpt_pitch};
    // the overflow counter lives in the table header (word 1)
    kmer_count_kernel<<<grid_for(n_rows * (int64_t)n_sel), kThreads, 0, st>>>(
        codes, code_pitch, rows, n_rows, tr, locimask, sel, n_sel, which, tb, tb.n_entries + 1);
    DJ_LAUNCH_CHECK();
    return 0;
}

extern "C" int dj_kmer_export(const void *table, uint32_t capacity, DjKmerRecord *h_out, uint32_t max_records,
                              uint32_t *h_n_records, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    DjKmerTableView tb = dj_kmer_view(const_cast<void *>(table), capacity);
    uint32_t header[2] = {0, 0};
    DJ_CUDA_CHECK(cudaMemcpyAsync(header, tb.n_entries, sizeof(header), cudaMemcpyDeviceToHost, st));
    DJ_CUDA_CHECK(cudaStreamSynchronize(st));
    if (header[1] != 0) {
        dj_set_error("dj_kmer_export: 3-mer table overflow (%u inserts lost); use a larger capacity", header[1]);
        return 2;
    }
    *h_n_records = header[0];
    if (header[0] == 0) return 0;
    if (header[0] > max_records) {
        dj_set_error("dj_kmer_export: %u records do not fit in %u", header[0], max_records);
        return 3;
    }
    DjKmerRecord *d_out = nullptr;
    uint32_t *d_n = nullptr;
    if (dj_scratch_pool_init() != 0) return 1;
    // stream-ordered scratch: cudaMalloc / cudaFree synchronise the device and cost tens of milliseconds in a
    // process that holds gigabytes of allocations
    DJ_CUDA_CHECK(cudaMallocAsync(&d_out, sizeof(DjKmerRecord) * (size_t)header[0] + 16, st));
    d_n = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(d_out) + sizeof(DjKmerRecord) * (size_t)header[0]);
    DJ_CUDA_CHECK(cudaMemsetAsync(d_n, 0, sizeof(uint32_t), st));
    kmer_export_kernel<<<grid_for(capacity), kThreads, 0, st>>>(tb, d_out, header[0], d_n);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess)
        err = cudaMemcpyAsync(h_out, d_out, sizeof(DjKmerRecord) * (size_t)header[0], cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    cudaFreeAsync(d_out, st);
    if (err != cudaSuccess) {
        dj_set_error("dj_kmer_export: %s", cudaGetErrorString(err));
        return 1;
    }
    return 0;
}

extern "C" int dj_annotate(const uint8_t *codes, int32_t code_pitch, const int32_t *rows, int64_t n_rows,
                           const uint8_t *text, const int64_t *row_off, const int32_t *row_len,
                           const uint32_t *ckpt, int32_t ckpt_pitch, const uint8_t *locimask, int32_t n_loci,
                           const int32_t *col_locus, const int32_t *col_sel, const uint16_t *col_at_id,
                           int32_t n_cols, int32_t is_control, const void *table, uint32_t capacity,
                           const uint16_t *slot_value_id, uint16_t *ids, void *stream) {
    if (n_rows <= 0 || n_cols <= 0) return 0;
    DjKmerTableView tb = dj_kmer_view(const_cast<void *>(table), capacity);
    DjTextRef tr{text, row_off, row_len, ckpt, ckpt_pitch};
    annotate_kernel<<<grid_for(n_rows * (int64_t)n_cols), kThreads, 0, (cudaStream_t)stream>>>(
        codes, code_pitch, rows, n_rows, tr, locimask, n_loci, col_locus, col_sel, col_at_id, n_cols, is_control, tb,
        slot_value_id, ids);
    DJ_LAUNCH_CHECK();
    return 0;
}

extern "C" int dj_fetch_tokens(const uint8_t *text, const int64_t *row_off, const int32_t *row_len,
                               const uint32_t *ckpt, int32_t ckpt_pitch, const int32_t *reads, const int32_t *loci,
                               int32_t n, uint8_t *out, int32_t *out_len, int32_t max_len, void *stream) {
    if (n <= 0) return 0;
    DjTextRef tr{text, row_off, row_len, ckpt, ckpt_pitch};
    fetch_tokens_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(tr, reads, loci, n, out, out_len, max_len);
    DJ_LAUNCH_CHECK();
    return 0;
}

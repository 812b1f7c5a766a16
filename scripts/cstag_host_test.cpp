// Host execution of dajin2_b200/csrc/cstag_core.cuh (the per-read routine of dj_cs_midsv_measure / _write): the 32
// lanes of a warp are 32 threads, every warp collective is a barrier + exchange through a shared slot array.  Built
// by tests/test_host_logic.py as a shared library (g++ -O1 -pthread -shared -fPIC) and driven through ctypes with
// the same arguments as the C ABI, so the CPU suite runs the device code's own lane arithmetic, carries and staged
// row writer against the oracle.  TEST INFRASTRUCTURE (not linked into the product).
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

struct uint4 {
    uint32_t x, y, z, w;
};

static pthread_barrier_t g_bar;
static thread_local int t_lane = 0;
static int64_t g_slot[32];

static inline void emu_wait() { pthread_barrier_wait(&g_bar); }

template <class T>
static inline T emu_shfl(T v, int src) {
    g_slot[t_lane] = (int64_t)v;
    emu_wait();
    const T r = (T)g_slot[src & 31];
    emu_wait();
    return r;
}

template <class T>
static inline T emu_shfl_up(T v, int d) {
    return emu_shfl(v, t_lane - d >= 0 ? t_lane - d : t_lane);
}

static inline uint32_t emu_ballot(bool p) {
    g_slot[t_lane] = p ? 1 : 0;
    emu_wait();
    uint32_t m = 0;
    for (int k = 0; k < 32; ++k) m |= (uint32_t)(g_slot[k] != 0) << k;
    emu_wait();
    return m;
}

#define DJW_FN static inline
#define DJW_SHFL(v, src) emu_shfl((v), (src))
#define DJW_SHFL_UP(v, d) emu_shfl_up((v), (d))
#define DJW_SHFL_XOR(v, d) emu_shfl((v), t_lane ^ (d))
#define DJW_BALLOT(p) emu_ballot((p))
#define DJW_SYNC() emu_wait()
#define DJW_CLZ(x) __builtin_clz((unsigned)(x))
static inline uint4 djw_load16(const uint8_t *p) {
    uint4 v;
    memcpy(&v, p, 16);
    return v;
}

#include "../dajin2_b200/csrc/cstag_core.cuh"

alignas(16) static uint8_t g_stage[DJ_CS_STAGE];
static DjCsLut g_lut;

template <bool kWrite>
static void run(const uint8_t *cs, const int64_t *aln_off, const int32_t *aln_pos, const uint8_t *aln_lower,
                const int64_t *read_first, int64_t n_reads, int32_t ref_len, const uint8_t *ref_seq, const int64_t *row_off,
                int32_t *row_len, int32_t *row_status, int32_t *row_n, uint8_t *text) {
    pthread_barrier_init(&g_bar, nullptr, 32);
    djcs_lut_fill(&g_lut, 0, 1);
    std::vector<std::thread> lanes;
    for (int lane = 0; lane < 32; ++lane) {
        lanes.emplace_back([=]() {
            t_lane = lane;
            for (int64_t r = 0; r < n_reads; ++r) {
                if (kWrite) {
                    int32_t len = row_len[r];
                    if (len <= 0) continue;
                    djcs_read<true>(g_lut, cs, aln_off, aln_pos, aln_lower, read_first[r], read_first[r + 1], ref_len, ref_seq,
                                    g_stage, text + row_off[r], &len, nullptr, nullptr, lane);
                } else {
                    djcs_read<false>(g_lut, cs, aln_off, aln_pos, aln_lower, read_first[r], read_first[r + 1], ref_len, ref_seq,
                                     g_stage, nullptr, row_len + r, row_status + r, row_n ? row_n + r : nullptr, lane);
                }
            }
        });
    }
    for (auto &t : lanes) t.join();
    pthread_barrier_destroy(&g_bar);
}

extern "C" int cs_host_measure(const uint8_t *cs, const int64_t *aln_off, const int32_t *aln_pos, const uint8_t *aln_lower,
                               const int64_t *read_first, int64_t n_reads, int32_t ref_len, const uint8_t *ref_seq,
                               int32_t *row_len, int32_t *row_status, int32_t *row_n) {
    run<false>(cs, aln_off, aln_pos, aln_lower, read_first, n_reads, ref_len, ref_seq, nullptr, row_len, row_status, row_n,
               nullptr);
    return 0;
}

extern "C" int cs_host_write(const uint8_t *cs, const int64_t *aln_off, const int32_t *aln_pos, const uint8_t *aln_lower,
                             const int64_t *read_first, int64_t n_reads, int32_t ref_len, const uint8_t *ref_seq,
                             const int64_t *row_off, int32_t *row_len, uint8_t *text) {
    run<true>(cs, aln_off, aln_pos, aln_lower, read_first, n_reads, ref_len, ref_seq, row_off, row_len, nullptr, nullptr,
              text);
    return 0;
}

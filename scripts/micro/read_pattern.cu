// Micro-benchmark: how fast can a uint8[N][pitch] matrix be streamed with different warp access patterns?
// (decides the tiling of the histogram kernel).  nvcc -arch=sm_100a -O3 read_pattern.cu -o read_pattern
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t fold(uint4 v) { return (v.x | v.y | v.z | v.w) & 0xFCFCFCFCu; }

// column tiles of `TILE` bytes; a CTA (8 warps) owns one tile column, warps take blocks of 32 rows dynamically
template <int TILE, int UNROLL>
__global__ void __launch_bounds__(256) col_tiled(const uint8_t *m, int pitch, int64_t n_rows, int *next_row, unsigned long long *sink) {
    constexpr int WARPS_X = TILE / 512;          // warps side by side on one row
    constexpr int WARPS_Y = 8 / WARPS_X;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wx = warp % WARPS_X, wy = warp / WARPS_X;
    int col = blockIdx.x * TILE + wx * 512 + lane * 16;
    if (col >= pitch) col = 0;
    unsigned long long acc = 0;
    __shared__ int64_t claim[8];
    for (;;) {
        int64_t t;
        if (WARPS_X == 1) {
            t = 0;
            if (lane == 0) t = atomicAdd(next_row + blockIdx.x, 32);
            t = __shfl_sync(0xffffffffu, t, 0);
        } else {  // the WARPS_X warps of a row group move together
            __syncthreads();
            if (threadIdx.x < WARPS_Y) claim[threadIdx.x] = atomicAdd(next_row + blockIdx.x, 32);
            __syncthreads();
            t = claim[wy];
        }
        if (t >= n_rows) break;
        const uint8_t *p = m + t * (int64_t)pitch + col;
        const int nb = (int)min((int64_t)32, n_rows - t);
        for (int u0 = 0; u0 < nb; u0 += UNROLL) {
            uint4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) v[u] = u0 + u < nb ? __ldg((const uint4 *)(p + (int64_t)(u0 + u) * pitch)) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) acc += __popc(__ballot_sync(0xffffffffu, fold(v[u]) != 0));
        }
    }
    if (acc == 0x123456789ULL) *sink = acc;
}

// a warp streams whole rows: pitch bytes, 512 at a time
template <int UNROLL>
__global__ void __launch_bounds__(256) row_major(const uint8_t *m, int pitch, int64_t n_rows, int *next_row, unsigned long long *sink) {
    const int lane = threadIdx.x & 31;
    unsigned long long acc = 0;
    const int n_seg = (pitch + 511) / 512;
    for (;;) {
        int64_t t = 0;
        if (lane == 0) t = atomicAdd(next_row, 4);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_rows) break;
        const int nb = (int)min((int64_t)4, n_rows - t);
        for (int r = 0; r < nb; ++r) {
            const uint8_t *p = m + (t + r) * (int64_t)pitch + lane * 16;
            for (int s0 = 0; s0 < n_seg; s0 += UNROLL) {
                uint4 v[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) v[u] = (s0 + u) * 512 + lane * 16 < pitch ? __ldg((const uint4 *)(p + (s0 + u) * 512)) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) acc += __popc(__ballot_sync(0xffffffffu, fold(v[u]) != 0));
            }
        }
    }
    if (acc == 0x123456789ULL) *sink = acc;
}

template <typename F>
float time_it(F launch, int *next_row, int n_counters) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int it = 0; it < 6; ++it) {
        cudaMemset(next_row, 0, sizeof(int) * n_counters);
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (it > 0 && ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv) {
    const int64_t n_rows = argc > 1 ? atoll(argv[1]) : 1000000;
    const int pitch = argc > 2 ? atoi(argv[2]) : 5024;
    uint8_t *m; int *next_row; unsigned long long *sink;
    cudaMalloc(&m, n_rows * pitch + 4096);
    cudaMemset(m, 1, n_rows * pitch + 4096);
    cudaMalloc(&next_row, 4096); cudaMalloc(&sink, 8);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double gb = (double)n_rows * pitch / 1e9;
    auto report = [&](const char *name, float ms) { printf("%-28s %8.3f ms %8.1f GB/s\n", name, ms, gb / ms * 1e3); };
    {
        const int gx = (pitch + 511) / 512, gy = sms * 4 / gx;
        report("col512 unroll4", time_it([&] { col_tiled<512, 4><<<dim3(gx, gy), 256>>>(m, pitch, n_rows, next_row, sink); }, next_row, gx));
        report("col512 unroll8", time_it([&] { col_tiled<512, 8><<<dim3(gx, gy), 256>>>(m, pitch, n_rows, next_row, sink); }, next_row, gx));
        const int gy8 = sms * 8 / gx;
        report("col512 unroll4 8cta/sm", time_it([&] { col_tiled<512, 4><<<dim3(gx, gy8), 256>>>(m, pitch, n_rows, next_row, sink); }, next_row, gx));
    }
    {
        const int gx = (pitch + 1023) / 1024, gy = sms * 4 / gx;
        report("col1024 unroll8", time_it([&] { col_tiled<1024, 8><<<dim3(gx, gy), 256>>>(m, pitch, n_rows, next_row, sink); }, next_row, gx));
    }
    {
        const int gx = (pitch + 2047) / 2048, gy = sms * 4 / gx;
        report("col2048 unroll8", time_it([&] { col_tiled<2048, 8><<<dim3(gx, gy), 256>>>(m, pitch, n_rows, next_row, sink); }, next_row, gx));
    }
    report("row-major unroll5 4cta/sm", time_it([&] { row_major<5><<<sms * 4, 256>>>(m, pitch, n_rows, next_row, sink); }, next_row, 1));
    report("row-major unroll10 4cta/sm", time_it([&] { row_major<10><<<sms * 4, 256>>>(m, pitch, n_rows, next_row, sink); }, next_row, 1));
    report("row-major unroll10 8cta/sm", time_it([&] { row_major<10><<<sms * 8, 256>>>(m, pitch, n_rows, next_row, sink); }, next_row, 1));
    // plain copy-like read for reference
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}

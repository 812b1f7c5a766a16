// Per-CTA timeline of hist_partial_kernel on a synthetic code matrix (2 % other-than-plain codes, 1 % N tails).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DDJ_HIST_TRACE hist_trace.cu ../../dajin2_b200/csrc/cabi.cu -o hist_trace
#include "../../dajin2_b200/csrc/hist.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

int main(int argc, char **argv) {
    const int64_t n_rows = argc > 1 ? atoll(argv[1]) : 100000;
    const int n_loci = 5000, pitch = 5024;
    const double rare = argc > 2 ? atof(argv[2]) : 0.02;
    std::vector<uint8_t> h((size_t)n_rows * pitch), hs(n_rows);
    uint64_t x = 88172645463325252ULL;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    for (int64_t r = 0; r < n_rows; ++r) {
        hs[r] = (rnd() & 1) ? '+' : '-';
        uint8_t *row = &h[(size_t)r * pitch];
        for (int i = 0; i < pitch; ++i) {
            row[i] = rnd() & 3;
            if ((rnd() % 100000) < rare * 100000) row[i] = 10 + (rnd() % 50);
        }
        if (rnd() % 100 == 0) { int tail = n_loci * (5 + rnd() % 25) / 100; for (int i = n_loci - tail; i < n_loci; ++i) row[i] = 4; }
        for (int i = n_loci; i < pitch; ++i) row[i] = 0x7F;
    }
    uint8_t *d, *ds; int32_t *counts;
    cudaMalloc(&d, h.size()); cudaMalloc(&ds, n_rows); cudaMalloc(&counts, sizeof(int32_t) * 10 * n_loci);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice); cudaMemcpy(ds, hs.data(), n_rows, cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(a);
        int rc = dj_hist(d, pitch, ds, nullptr, n_rows, n_loci, counts, nullptr);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("iter %d rc %d: %.3f ms  %.1f GB/s\n", it, rc, ms, (double)n_rows * pitch / ms / 1e6);
    }
    std::vector<unsigned long long> tr(8192 * 5);
    cudaMemcpyFromSymbol(tr.data(), dj_hist_trace, tr.size() * 8);
    unsigned long long t0 = ~0ULL, t3 = 0; int n = 0;
    for (int c = 0; c < 8192; ++c) if (tr[c * 5]) { t0 = std::min(t0, tr[c * 5]); t3 = std::max(t3, tr[c * 5 + 3]); n = c + 1; }
    printf("ctas %d, kernel span %.1f us\n", n, (t3 - t0) / 1e3);
    // histogram of start / loop-start / loop-end / end offsets (us)
    for (int k = 0; k < 4; ++k) {
        std::vector<double> v;
        for (int c = 0; c < n; ++c) v.push_back((tr[c * 5 + k] - t0) / 1e3);
        std::sort(v.begin(), v.end());
        printf("mark %d: min %.1f p10 %.1f p50 %.1f p90 %.1f max %.1f us\n", k, v[0], v[n / 10], v[n / 2], v[n * 9 / 10], v[n - 1]);
    }
    return 0;
}

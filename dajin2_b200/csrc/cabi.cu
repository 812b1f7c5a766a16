// Error plumbing and device queries shared by every entry point of libdajin2_b200.so.
#include <stdarg.h>
#include <stdio.h>

#include "dj_common.cuh"

namespace {
thread_local char g_error[512] = "";
}

void dj_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int dj_sm_count() {
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, sms = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess || sms <= 0) {
        dj_set_error("no usable CUDA device: %s", cudaGetErrorString(err));
        return 0;
    }
    cached = sms;
    return sms;
}

// Scratch buffers come from the device's stream-ordered pool (cudaMallocAsync): keep freed blocks in the pool across
// synchronisations instead of returning them to the driver.
int dj_scratch_pool_init() {
    static bool done = false;
    if (done) return 0;
    int dev = 0;
    cudaMemPool_t pool;
    DJ_CUDA_CHECK(cudaGetDevice(&dev));
    DJ_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t keep = UINT64_MAX;
    DJ_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    done = true;
    return 0;
}

extern "C" int dj_abi_version(void) { return DJ_ABI_VERSION; }

extern "C" const char *dj_last_error(void) { return g_error; }

extern "C" int dj_device_info(int32_t *h_out) {
    int dev = 0;
    DJ_CUDA_CHECK(cudaGetDevice(&dev));
    int sms = 0, major = 0, minor = 0;
    DJ_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DJ_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    DJ_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    h_out[0] = sms;
    h_out[1] = major;
    h_out[2] = minor;
    return 0;
}

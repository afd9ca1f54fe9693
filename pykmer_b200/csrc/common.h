// common.h -- error plumbing shared by the translation units of libpykmer_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pykmer_b200.h"

#define PK_API extern "C" __attribute__((visibility("default")))

int pk_set_error(int code, const char *fmt, ...);

// gram_i8.cu: gram (int64 N x N, zeroed or holding a partial sum) += B * B^T on the
// tensor cores (tcgen05 kind::i8); nsamples <= 256.
int pk_gram_i8_launch(const uint32_t *bits_dev, int nsamples, size_t words, size_t stride_words,
                      int64_t *gram_dev, int device, cudaStream_t st);
// gram_f4.cu: the same contraction on the block-scaled FP4 path (tcgen05 kind::mxf4) over TILED masks
// bits[word / 32][nsamples][32]; words a multiple of 32; any nsamples (more than 256: block pairs).
int pk_gram_f4_launch(const uint32_t *bits_dev, int nsamples, size_t words, int64_t *gram_dev, int device,
                      cudaStream_t st);
// 1 / 0: FP32 accumulation of 0/1 products is / is not exact up to 2^24 on this device (checked once)
int pk_gram_f4_exact(int device);

#define PK_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess)                                                         \
            return pk_set_error(e__ == cudaErrorMemoryAllocation ? PK_ERR_NOMEM         \
                                                                 : PK_ERR_CUDA,         \
                                "%s:%d %s -> %s", __FILE__, __LINE__, #call,            \
                                cudaGetErrorString(e__));                               \
    } while (0)

#define PK_REQUIRE(cond, ...)                                                           \
    do {                                                                                \
        if (!(cond)) return pk_set_error(PK_ERR_ARG, __VA_ARGS__);                      \
    } while (0)

static inline int pk_sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0)
        n = 148;
    return n;
}

// restores the caller's current device on scope exit
struct pk_device_guard {
    int prev = -1;
    bool ok = true;
    explicit pk_device_guard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
        want = device;
    }
    ~pk_device_guard() {
        if (prev >= 0 && prev != want) cudaSetDevice(prev);
    }
    int want = -1;
};

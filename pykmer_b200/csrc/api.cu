// api.cu -- general entry points of libpykmer_b200.so (errors, devices, pinned memory)
#include <string.h>

#include "common.h"

static thread_local char g_error[512] = "";

int pk_set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
    return code;
}

PK_API int pk_abi_version(void) { return PK_ABI_VERSION; }

PK_API const char *pk_last_error(void) { return g_error; }

PK_API int pk_device_count(int *count) {
    PK_REQUIRE(count != nullptr, "pk_device_count: NULL output");
    *count = 0;
    PK_CUDA(cudaGetDeviceCount(count));
    return PK_OK;
}

PK_API int pk_device_info(int device, char *name, size_t name_len, int *sm_count,
                          size_t *total_mem_bytes, int *cc_major, int *cc_minor) {
    cudaDeviceProp prop;
    PK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (name && name_len) {
        strncpy(name, prop.name, name_len - 1);
        name[name_len - 1] = '\0';
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (total_mem_bytes) *total_mem_bytes = prop.totalGlobalMem;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return PK_OK;
}

PK_API int pk_host_alloc(void **ptr, size_t bytes) {
    PK_REQUIRE(ptr != nullptr, "pk_host_alloc: NULL output");
    *ptr = nullptr;
    PK_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return PK_OK;
}

PK_API int pk_host_free(void *ptr) {
    if (ptr) PK_CUDA(cudaFreeHost(ptr));
    return PK_OK;
}

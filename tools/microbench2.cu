// microbench2.cu -- second round of B200 measurements for the indexer's window flush:
// can the saturating 8-bit table itself (not a 32-bit counter array) be the L2-resident
// counting target?  Measures, on L2-sized windows:
//   atom_u8_packed   atomicAdd WITH return on 4 packed byte lanes (carry detection needs the old word)
//   red_u8_packed    the same add without return (upper bound)
//   red_f16x2        red.global.add.noftz.f16x2 (carry-free 16-bit lanes, exact up to 2048)
//   window cycle     zero-fill W bytes -> random byte adds -> histogram read, window after window over a
//                    large table, with and without a persisting-L2 access policy window
// Stand-alone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench2 microbench2.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

__global__ void k_atom_u8_packed(uint8_t *t, uint64_t mask, uint64_t nops, uint64_t seed, uint32_t *ovf) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nops; i += stride) {
        const uint64_t idx = mix(i + seed) & mask;
        const uint32_t sh = 8 * (idx & 3);
        const uint32_t old = atomicAdd(reinterpret_cast<uint32_t *>(t + (idx & ~3ull)), 1u << sh);
        if (((old >> sh) & 0xFFu) == 0xFFu) atomicAdd(ovf, 1u);
    }
}
__global__ void k_red_u8_packed(uint8_t *t, uint64_t mask, uint64_t nops, uint64_t seed) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nops; i += stride) {
        const uint64_t idx = mix(i + seed) & mask;
        atomicAdd(reinterpret_cast<uint32_t *>(t + (idx & ~3ull)), 1u << (8 * (idx & 3)));
    }
}
__global__ void k_red_f16x2(__half2 *t, uint64_t mask, uint64_t nops, uint64_t seed) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nops; i += stride) {
        const uint64_t idx = mix(i + seed) & mask;          // index of a 16-bit lane
        const __half2 v = (idx & 1) ? __halves2half2(__float2half(0.f), __float2half(1.f))
                                    : __halves2half2(__float2half(1.f), __float2half(0.f));
        atomicAdd(t + (idx >> 1), v);
    }
}
// the same adds fed from a coalesced entry list (what the flush really does)
__global__ void k_atom_list(uint8_t *t, const uint32_t *ent, uint32_t n, uint32_t *ovf) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t e = __ldcs(ent + i);
        const uint32_t sh = 8 * (e & 3);
        const uint32_t old = atomicAdd(reinterpret_cast<uint32_t *>(t + (e & ~3u)), 1u << sh);
        if (((old >> sh) & 0xFFu) == 0xFFu) atomicAdd(ovf, 1u);
    }
}
__global__ void k_red_list(uint8_t *t, const uint32_t *ent, uint32_t n) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t e = __ldcs(ent + i);
        atomicAdd(reinterpret_cast<uint32_t *>(t + (e & ~3u)), 1u << (8 * (e & 3)));
    }
}
__global__ void k_fill_list(uint32_t *ent, uint32_t n, uint32_t mask, uint64_t seed) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        ent[i] = (uint32_t)mix(i + seed) & mask;
}
__global__ void k_zero(uint4 *dst, size_t nvec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride)
        dst[i] = make_uint4(0, 0, 0, 0);
}
__global__ void __launch_bounds__(256) k_hist(const uint4 *src, size_t nvec, unsigned long long *bins) {
    __shared__ uint32_t sh[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    uint32_t c1 = 0, c2 = 0, c3 = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 q = __ldcg(src + i);
        const uint32_t ws[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t x = ws[k];
            if (!x) continue;
            c1 += __popc(__vcmpeq4(x, 0x01010101u)) >> 3;
            c2 += __popc(__vcmpeq4(x, 0x02020202u)) >> 3;
            c3 += __popc(__vcmpeq4(x, 0x03030303u)) >> 3;
            if (x & 0xFCFCFCFCu) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const uint32_t val = (x >> (8 * b)) & 0xFFu;
                    if (val > 3u) atomicAdd(&sh[warp][val], 1u);
                }
            }
        }
    }
    if (c1) atomicAdd(&sh[warp][1], c1);
    if (c2) atomicAdd(&sh[warp][2], c2);
    if (c3) atomicAdd(&sh[warp][3], c3);
    __syncthreads();
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += sh[k][threadIdx.x];
    if (s && threadIdx.x) atomicAdd(&bins[threadIdx.x], s);
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
    void start(cudaStream_t s = 0) { CK(cudaEventRecord(a, s)); }
    float stop(cudaStream_t s = 0) { CK(cudaEventRecord(b, s)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }
};

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, 0);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, 0);
    printf("device %s sms %d l2 %d B max_persist %d B max_policy_window %d B\n", prop.name, sms, prop.l2CacheSize, max_persist, max_window);
    Timer tm;
    uint32_t *ovf;
    CK(cudaMalloc(&ovf, 4));
    CK(cudaMemset(ovf, 0, 4));
    unsigned long long *bins;
    CK(cudaMalloc(&bins, 256 * 8));
    CK(cudaMemset(bins, 0, 256 * 8));
    const size_t big = 8ull << 30;
    uint8_t *buf;
    CK(cudaMalloc(&buf, big));
    const int grid = sms * 8, block = 256;
    const uint64_t nops = 1ull << 27;       // sparse enough that no byte overflows in a >= 16 MiB window

    for (size_t sz : {64ull << 20, 32ull << 20, 16ull << 20}) {
        CK(cudaMemsetAsync(buf, 0, sz));
        k_red_u8_packed<<<grid, block>>>(buf, sz - 1, nops / 8, 1);
        CK(cudaMemsetAsync(buf, 0, sz));
        tm.start(); k_red_u8_packed<<<grid, block>>>(buf, sz - 1, nops, 2); float ms = tm.stop();
        printf("red_u8_packed     window %3zu MiB ops %llu ms %.3f Gop/s %.2f\n", sz >> 20, (unsigned long long)nops, ms, nops / ms / 1e6);
        CK(cudaMemsetAsync(buf, 0, sz));
        tm.start(); k_atom_u8_packed<<<grid, block>>>(buf, sz - 1, nops, 3, ovf); ms = tm.stop();
        printf("atom_u8_packed    window %3zu MiB ops %llu ms %.3f Gop/s %.2f\n", sz >> 20, (unsigned long long)nops, ms, nops / ms / 1e6);
        for (int mult : {16, 32}) {
            CK(cudaMemsetAsync(buf, 0, sz));
            tm.start(); k_atom_u8_packed<<<sms * mult / 4, 1024>>>(buf, sz - 1, nops, 3, ovf); ms = tm.stop();
            printf("atom_u8_packed    window %3zu MiB 1024thr x%d/4 per SM ms %.3f Gop/s %.2f\n", sz >> 20, mult, ms, nops / ms / 1e6);
        }
        CK(cudaMemsetAsync(buf, 0, sz));
        tm.start(); k_red_f16x2<<<grid, block>>>((__half2 *)buf, sz / 2 - 1, nops, 4); ms = tm.stop();
        printf("red_f16x2         window %3zu MiB ops %llu ms %.3f Gop/s %.2f\n", sz >> 20, (unsigned long long)nops, ms, nops / ms / 1e6);
    }

    // window cycle over a large table: per window zero -> adds (from a list) -> histogram
    uint32_t *list;
    const uint32_t list_cap = 64u << 20;
    CK(cudaMalloc(&list, (size_t)list_cap * 4));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (int persist = 0; persist < 2; persist++) {
        for (size_t wsz : {64ull << 20, 32ull << 20}) {
            for (uint32_t per_win : {750000u, 3000000u, 48000000u}) {
                for (int ret = 0; ret < 2; ret++) {
                    if (per_win * (wsz >> 20) / 64 > list_cap) continue;
                    const uint32_t n_ent = (uint32_t)((uint64_t)per_win * (wsz >> 20) / 64);
                    k_fill_list<<<grid, block, 0, st>>>(list, n_ent, (uint32_t)wsz - 1, 77);
                    const int nwin = (int)std::min<size_t>(big / wsz, 64);
                    if (persist) {
                        const size_t grant = std::min<size_t>(wsz, (size_t)max_persist);
                        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, grant));
                    }
                    cudaLaunchAttribute attr[1];
                    cudaLaunchConfig_t cfg;
                    memset(&cfg, 0, sizeof cfg);
                    cfg.blockDim = dim3(256);
                    cfg.stream = st;
                    cfg.attrs = attr;
                    float best = 1e9f;
                    for (int rep = 0; rep < 3; rep++) {
                        CK(cudaStreamSynchronize(st));
                        tm.start(st);
                        for (int w = 0; w < nwin; w++) {
                            uint8_t *win = buf + (size_t)w * wsz;
                            cfg.numAttrs = 0;
                            if (persist) {
                                attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
                                attr[0].val.accessPolicyWindow.base_ptr = win;
                                attr[0].val.accessPolicyWindow.num_bytes = wsz;
                                attr[0].val.accessPolicyWindow.hitRatio = (size_t)max_persist >= wsz ? 1.0f : (float)max_persist / (float)wsz;
                                attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                                attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                                cfg.numAttrs = 1;
                            }
                            cfg.gridDim = dim3(sms * 4);
                            CK(cudaLaunchKernelEx(&cfg, k_zero, (uint4 *)win, wsz / 16));
                            cfg.gridDim = dim3(sms * 8);
                            if (ret) CK(cudaLaunchKernelEx(&cfg, k_atom_list, win, (const uint32_t *)list, n_ent, ovf));
                            else     CK(cudaLaunchKernelEx(&cfg, k_red_list, win, (const uint32_t *)list, n_ent));
                            cfg.gridDim = dim3(sms * 4);
                            CK(cudaLaunchKernelEx(&cfg, k_hist, (const uint4 *)win, wsz / 16, bins));
                        }
                        const float ms = tm.stop(st);
                        best = ms < best ? ms : best;
                    }
                    printf("window_cycle persist %d window %2zu MiB entries/window %8u %s: %d windows %.3f ms = %.2f us/window, %.1f GB/s of table, %.1f G adds/s\n",
                           persist, wsz >> 20, n_ent, ret ? "atom" : "red ", nwin, best, best * 1e3 / nwin,
                           (double)nwin * wsz / best / 1e6, (double)nwin * n_ent / best / 1e6);
                }
            }
        }
    }
    if (max_persist) cudaCtxResetPersistingL2Cache();
    uint32_t h_ovf = 0;
    CK(cudaMemcpy(&h_ovf, ovf, 4, cudaMemcpyDeviceToHost));
    printf("overflow events seen %u\n", h_ovf);
    return 0;
}

// microbench.cu -- measures the B200 primitives that decide the indexer's counting
// scheme and the merger's Gram kernel ("the first gpurun must measure ...").
// Stand-alone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
// Prints one line per experiment: name, size, ops, ms, Gop/s (and GB/s where meaningful).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// ---- global atomics -------------------------------------------------------------------
__global__ void k_red_u32(uint32_t *t, uint64_t mask, uint64_t nops) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nops; i += stride)
        atomicAdd(t + (mix(i) & mask), 1u);
}

__device__ __forceinline__ void sat_add_u8(uint8_t *table, uint64_t idx, uint32_t cnt) {
    uint32_t *wp = reinterpret_cast<uint32_t *>(table + (idx & ~3ull));
    const uint32_t sh = (uint32_t)(idx & 3) * 8;
    uint32_t old = __ldcg(wp);
    for (;;) {
        const uint32_t b = (old >> sh) & 0xFFu;
        if (b == 255u) return;
        const uint32_t nb = min(255u, b + cnt);
        const uint32_t assumed = old;
        old = atomicCAS(wp, assumed, (assumed & ~(0xFFu << sh)) | (nb << sh));
        if (old == assumed) return;
    }
}
__global__ void k_cas_u8(uint8_t *t, uint64_t mask, uint64_t nops) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nops; i += stride)
        sat_add_u8(t, mix(i) & mask, 1u);
}
// blind byte add without saturation handling (upper bound for any byte-atomic scheme)
__global__ void k_red_u8_blind(uint8_t *t, uint64_t mask, uint64_t nops) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nops; i += stride) {
        const uint64_t idx = mix(i) & mask;
        atomicAdd(reinterpret_cast<uint32_t *>(t + (idx & ~3ull)), 1u << (8 * (idx & 3)));
    }
}
// plain random 4-byte loads (what the table fetch alone costs)
__global__ void k_load_u32(const uint32_t *t, uint64_t mask, uint64_t nops, uint32_t *sink) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nops; i += stride)
        acc += __ldcg(t + (mix(i) & mask));
    if (acc == 0x12345678u) *sink = acc;
}

// ---- shared-memory atomics --------------------------------------------------------------
template <int WORDS>
__global__ void __launch_bounds__(1024) k_smem_atomic(uint64_t nops_per_block, uint32_t *sink) {
    extern __shared__ uint32_t sm[];
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * nops_per_block;
    for (uint64_t i = threadIdx.x; i < nops_per_block; i += blockDim.x)
        atomicAdd(&sm[mix(base + i) & (WORDS - 1)], 1u);
    __syncthreads();
    uint32_t acc = 0;
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) acc += sm[i];
    if (acc == 0xFFFFFFFFu) *sink = acc;
}
// 16-bit lanes in 32-bit words with the "read first, skip when >= 255" rule
template <int WORDS>
__global__ void __launch_bounds__(1024) k_smem_u16_checked(uint64_t nops_per_block, uint32_t *sink) {
    extern __shared__ uint32_t sm[];
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * nops_per_block;
    for (uint64_t i = threadIdx.x; i < nops_per_block; i += blockDim.x) {
        const uint32_t idx = (uint32_t)(mix(base + i) & (2 * WORDS - 1));
        const uint32_t sh = (idx & 1) * 16;
        if (((sm[idx >> 1] >> sh) & 0xFFFFu) < 255u) atomicAdd(&sm[idx >> 1], 1u << sh);
    }
    __syncthreads();
    uint32_t acc = 0;
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) acc += sm[i];
    if (acc == 0xFFFFFFFFu) *sink = acc;
}

// ---- ALU: AND + POPC + ADD vs carry-save ----------------------------------------------------
__global__ void k_popc(const uint32_t *in, uint32_t *out, int iters) {
    uint32_t a = in[threadIdx.x], b = in[threadIdx.x + 32], acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    for (int i = 0; i < iters; i++) {
        acc0 += __popc(a & b); a += 0x9E3779B9u;
        acc1 += __popc(a & b); b ^= a;
        acc2 += __popc(a & b); a += 0x7F4A7C15u;
        acc3 += __popc(a & b); b += a;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
}

// ---- streaming ------------------------------------------------------------------------------
__global__ void k_stream_write(uint4 *dst, size_t nvec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride)
        dst[i] = make_uint4((uint32_t)i, 1, 2, 3);
}
__global__ void k_stream_read(const uint4 *src, size_t nvec, uint32_t *sink) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 q = __ldcs(src + i);
        acc += q.x ^ q.y ^ q.z ^ q.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

struct Timer {
    cudaEvent_t a, b;
    Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
    void start() { CK(cudaEventRecord(a)); }
    float stop() { CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }
};

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s sms %d\n", prop.name, sms);
    Timer tm;
    uint32_t *sink;
    CK(cudaMalloc(&sink, 4));
    const size_t big = 4ull << 30;
    uint8_t *buf;
    CK(cudaMalloc(&buf, big));
    const int grid = sms * 8, block = 256;

    // streaming reference points
    for (int rep = 0; rep < 2; rep++) {
        tm.start(); k_stream_write<<<sms * 16, 256>>>((uint4 *)buf, big / 16); float ms = tm.stop();
        if (rep) printf("stream_write      bytes %zu ms %.3f GB/s %.1f\n", big, ms, big / ms / 1e6);
        tm.start(); k_stream_read<<<sms * 16, 256>>>((const uint4 *)buf, big / 16, sink); ms = tm.stop();
        if (rep) printf("stream_read       bytes %zu ms %.3f GB/s %.1f\n", big, ms, big / ms / 1e6);
        tm.start(); CK(cudaMemsetAsync(buf, 0, big)); ms = tm.stop();
        if (rep) printf("memset            bytes %zu ms %.3f GB/s %.1f\n", big, ms, big / ms / 1e6);
    }

    const uint64_t nops = 1ull << 28;
    const size_t sizes[] = {4ull << 30, 1ull << 30, 256ull << 20, 96ull << 20, 64ull << 20, 32ull << 20, 16ull << 20, 4ull << 20};
    for (size_t sz : sizes) {
        // power-of-two masks; 96 MiB uses a 64 MiB mask + offset pattern -> skip non powers of two
        if (sz & (sz - 1)) continue;
        CK(cudaMemsetAsync(buf, 0, sz));
        k_red_u32<<<grid, block>>>((uint32_t *)buf, sz / 4 - 1, nops / 8);     // warm
        tm.start(); k_red_u32<<<grid, block>>>((uint32_t *)buf, sz / 4 - 1, nops); float ms = tm.stop();
        printf("red_u32           table %5zu MiB ops %llu ms %.3f Gop/s %.2f\n", sz >> 20, (unsigned long long)nops, ms, nops / ms / 1e6);
        CK(cudaMemsetAsync(buf, 0, sz));
        k_cas_u8<<<grid, block>>>(buf, sz - 1, nops / 8);
        CK(cudaMemsetAsync(buf, 0, sz));
        tm.start(); k_cas_u8<<<grid, block>>>(buf, sz - 1, nops); ms = tm.stop();
        printf("cas_u8_sat        table %5zu MiB ops %llu ms %.3f Gop/s %.2f\n", sz >> 20, (unsigned long long)nops, ms, nops / ms / 1e6);
        CK(cudaMemsetAsync(buf, 0, sz));
        tm.start(); k_red_u8_blind<<<grid, block>>>(buf, sz - 1, nops); ms = tm.stop();
        printf("red_u8_blind      table %5zu MiB ops %llu ms %.3f Gop/s %.2f\n", sz >> 20, (unsigned long long)nops, ms, nops / ms / 1e6);
        tm.start(); k_load_u32<<<grid, block>>>((const uint32_t *)buf, sz / 4 - 1, nops, sink); ms = tm.stop();
        printf("load_u32          table %5zu MiB ops %llu ms %.3f Gop/s %.2f\n", sz >> 20, (unsigned long long)nops, ms, nops / ms / 1e6);
    }
    // occupancy sweep for the DRAM-resident CAS (latency bound?)
    for (int mult : {2, 4, 8, 16, 32}) {
        CK(cudaMemsetAsync(buf, 0, 1ull << 30));
        tm.start(); k_cas_u8<<<sms * mult, 256>>>(buf, (1ull << 30) - 1, nops); float ms = tm.stop();
        printf("cas_u8_sat 1GiB   blocks/SM %2d ms %.3f Gop/s %.2f\n", mult, ms, nops / ms / 1e6);
    }

    // shared-memory atomics: one block of 1024 threads per SM (x2 waves), random words
    {
        const uint64_t per_block = 1ull << 22;
        const int blocks = sms * 2;
        CK(cudaFuncSetAttribute(k_smem_atomic<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
        CK(cudaFuncSetAttribute(k_smem_u16_checked<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
        for (int rep = 0; rep < 2; rep++) {
            tm.start(); k_smem_atomic<32768><<<blocks, 1024, 32768 * 4>>>(per_block, sink); float ms = tm.stop();
            if (rep) printf("smem_atomic_u32   words 32768 (128 KiB) ops %llu ms %.3f Gop/s %.2f\n", (unsigned long long)(per_block * blocks), ms, per_block * blocks / ms / 1e6);
            tm.start(); k_smem_atomic<8192><<<blocks, 1024, 8192 * 4>>>(per_block, sink); ms = tm.stop();
            if (rep) printf("smem_atomic_u32   words  8192 ( 32 KiB) ops %llu ms %.3f Gop/s %.2f\n", (unsigned long long)(per_block * blocks), ms, per_block * blocks / ms / 1e6);
            tm.start(); k_smem_u16_checked<32768><<<blocks, 1024, 32768 * 4>>>(per_block, sink); ms = tm.stop();
            if (rep) printf("smem_u16_checked  lanes 65536 (128 KiB) ops %llu ms %.3f Gop/s %.2f\n", (unsigned long long)(per_block * blocks), ms, per_block * blocks / ms / 1e6);
            tm.start(); k_smem_atomic<8192><<<sms * 4, 512, 8192 * 4>>>(per_block / 2, sink); ms = tm.stop();
            if (rep) printf("smem_atomic_u32   words  8192 4x512thr     ops %llu ms %.3f Gop/s %.2f\n", (unsigned long long)(per_block / 2 * sms * 4), ms, per_block / 2 * sms * 4 / ms / 1e6);
        }
    }
    // AND+POPC+ADD rate
    {
        uint32_t *in, *out;
        CK(cudaMalloc(&in, 64 * 4)); CK(cudaMalloc(&out, sms * 16 * 256 * 4));
        CK(cudaMemset(in, 0x5A, 64 * 4));
        const int iters = 1 << 16;
        for (int rep = 0; rep < 2; rep++) {
            tm.start(); k_popc<<<sms * 16, 256>>>(in, out, iters); float ms = tm.stop();
            const double ops = (double)sms * 16 * 256 * iters * 4;
            if (rep) printf("and_popc_add      ops %.3e ms %.3f Gpopc/s %.1f\n", ops, ms, ops / ms / 1e6);
        }
    }
    return 0;
}

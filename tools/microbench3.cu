// microbench3.cu -- measurement tool, not product code: the rate of the legacy binary tensor-core
// instruction mma.sync.aligned.m16n8k256.b1.and.popc on sm_100a ("if it is not
// emulated it needs no expansion at all; record the number either way").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench3 tools/microbench3.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_b1(int iters, int *out) {
    unsigned a[4] = {threadIdx.x * 2654435761u, blockIdx.x * 40503u + 1u, 0x55555555u, threadIdx.x + 7u};
    unsigned b[2] = {threadIdx.x * 0x9E3779B9u + 3u, 0x33333333u};
    int c[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) c[j][i] = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            asm volatile(
                "mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc "
                "{%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
                : "r"(a[0] + j), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    int s = 0;
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) s += c[j][i];
    if (s == 0x7fffffff) out[0] = s;
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int *out;
    cudaMalloc(&out, 4);
    const int iters = 20000, blocks = sms * 4, threads = 256;
    k_b1<<<blocks, threads>>>(100, out);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_b1<<<blocks, threads>>>(iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const cudaError_t err = cudaGetLastError();
    const double mmas = (double)iters * 4 * blocks * (threads / 32);
    const double macs = mmas * 16 * 8 * 256;
    printf("mma.sync.m16n8k256.b1.and.popc: %s, %.3f ms, %.3e AND+POPC bit-MACs/s (%.1f T/s), %.2f MMA/clk/SM at 1.965 GHz\n",
           cudaGetErrorString(err), ms, macs / (ms * 1e-3), macs / (ms * 1e-3) / 1e12,
           mmas / (ms * 1e-3) / sms / 1.965e9);
    printf("for scale: the N=255, K=15 Gram needs 2^30 * 255^2 = %.3e bit-MACs; FP4 tcgen05 issues 3 * 128 * 128 * 2^30 = %.3e\n",
           1073741824.0 * 255 * 255, 3.0 * 128 * 128 * 1073741824.0);
    return 0;
}

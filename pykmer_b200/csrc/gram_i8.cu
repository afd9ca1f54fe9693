// gram_i8.cu -- the merger's Gram matrix on the 5th-generation tensor cores.
//
// Replaces the pair loop of merge (merger.py:136-176) over Header.calculate_distance
// (tools.py:439-493): G = B * B^T with B the N x 4^K matrix of 0/1 presence bits.  This
// really is a dense contraction, so it runs as tcgen05.mma kind::i8 (u8 x u8 -> s32 in
// TMEM): exact, because every product is 0 or 1 and one CTA accumulates fewer than 2^31
// of them.
//
// One persistent CTA per SM walks a contiguous slab of the k-mer axis:
//   warps 0-3  producers: read the packed bitmask words (32 k-mers each) of every sample,
//              blow every bit up to one byte (0/1) straight into the canonical K-major,
//              no-swizzle UMMA core-matrix layout in shared memory -- a 32-bit word is
//              exactly one 32-byte row of one K=32 MMA step -- then fence.proxy.async
//              and arrive on the stage's "full" mbarrier.
//   warp 4     one thread issues the MMAs: A and B descriptors point at the SAME tile
//              (it is a Gram matrix), D[128 x NP] s32 accumulates in TMEM;
//              tcgen05.commit frees the stage.
//   warps 0-3  epilogue: tcgen05.ld the accumulator rows, add them into the global
//              int64 Gram matrix.
// NP (samples padded) is 64, 128 or 256; 256 uses two M=128 accumulators (all 512 TMEM
// columns); 64 uses M=64 MMAs (half the A-operand traffic from shared memory, which is what
// bounds that small tile).
#include <algorithm>
#include <stdlib.h>

#include "common.h"
#include "tcgen05_util.h"

namespace {

using namespace pk_umma;

// per-config constants live in GramCfg<NP>: K=32 steps (= bitmask words) per pipeline stage,
// producer warps, stage ring depth

// instruction descriptor, kind::i8: D = S32, A = B = unsigned 8 bit, both K-major, M x N tile
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 4 bits -> 4 bytes of 0/1 (bit j -> byte j)
__device__ __forceinline__ uint32_t spread4(uint32_t nib) {
    return (nib * 0x00204081u) & 0x01010101u;
}
// one bitmask word -> one 32-byte K-major row: k-mers 0..15 in chunk 0, 16..31 in chunk 1
__device__ __forceinline__ void expand_word(uint32_t w, uint8_t *row_chunk0, uint32_t lbo) {
    uint4 lo, hi;
    lo.x = spread4(w & 15u);         lo.y = spread4((w >> 4) & 15u);
    lo.z = spread4((w >> 8) & 15u);  lo.w = spread4((w >> 12) & 15u);
    hi.x = spread4((w >> 16) & 15u); hi.y = spread4((w >> 20) & 15u);
    hi.z = spread4((w >> 24) & 15u); hi.w = spread4(w >> 28);
    *reinterpret_cast<uint4 *>(row_chunk0) = lo;
    *reinterpret_cast<uint4 *>(row_chunk0 + lbo) = hi;
}

template <int NP>
struct GramCfg {
    // The bit -> byte expansion is ALU work with short dependent chains: it needs several warps per
    // scheduler to run near issue rate, so small NP (cheap MMAs, 8 x 32 cycles per stage at NP=64)
    // gets 8 producer warps, each filling whole stages on its own.
    static constexpr int kProducerWarps = NP == 256 ? 4 : 8;
    static constexpr int kProducerThreads = kProducerWarps * 32;
    static constexpr int kThreads = kProducerThreads + 32;            // + the MMA-issuing warp
    static constexpr int kGroups = NP == 256 ? 1 : 8;     // producer groups filling stages in parallel
    static constexpr int kKB = NP == 128 ? 4 : 8;         // K=32 steps (bitmask words) per stage
    static constexpr int kStages = NP == 256 ? 3 : 8;
    static constexpr int kTileBytes = NP * 32;            // one K=32 step of all NP rows
    static constexpr int kStageBytes = kKB * kTileBytes;
    static constexpr int kMTiles = NP == 256 ? 2 : 1;
    // NP = 64 runs M = 64 MMAs: same 32 cycles as M = 128, but only half the A operand is read
    // from shared memory, which is what bounds this small tile (4 KB instead of 6 KB per MMA).
    // D then sits in "half sub-partition" layout: row r in TMEM lane (r % 16) + 32 * (r / 16).
    static constexpr int kM = NP == 64 ? 64 : 128;
    static constexpr int kTmemCols = NP == 256 ? 512 : NP;
    static constexpr int kPad = 4096;                     // the M=128 descriptor of a 64-row tile overruns
    static constexpr size_t kSmem = (size_t)kStages * kStageBytes + kPad + 256 + 1024;
};

template <int NP, int AHEAD>
__global__ void __launch_bounds__(GramCfg<NP>::kThreads, 1) k_gram_i8(const uint32_t *__restrict__ bits, int nsamples,
                                                          size_t words, size_t stride_words,
                                                          unsigned long long *__restrict__ gram) {
    using C = GramCfg<NP>;
    constexpr int kKB = C::kKB;
    constexpr int kProducerThreads = C::kProducerThreads;
    constexpr int kMmaWarp = C::kProducerWarps;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *ctrl = smem + (size_t)C::kStages * C::kStageBytes + C::kPad;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(ctrl);               // [kStages]
    uint64_t *empty_bar = full_bar + C::kStages;                            // [kStages]
    uint64_t *done_bar = empty_bar + C::kStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // this CTA's slab of words, in whole stages of kKB words
    const size_t total_stages = (words + kKB - 1) / kKB;
    const size_t per_cta = (total_stages + gridDim.x - 1) / gridDim.x;
    const size_t st0 = min(total_stages, (size_t)blockIdx.x * per_cta);
    const size_t st1 = min(total_stages, st0 + per_cta);
    const size_t nst = st1 - st0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; s++) {
            mbar_init(smem_u32(&full_bar[s]), kProducerThreads / C::kGroups);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(done_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero the pad once: it is read (and ignored) by the over-running A descriptor
    for (int i = threadIdx.x; i < C::kPad / 16; i += blockDim.x)
        reinterpret_cast<uint4 *>(smem + (size_t)C::kStages * C::kStageBytes)[i] = make_uint4(0, 0, 0, 0);
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(tmem_slot)), "n"(C::kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kMmaWarp) {
        // ------------------------------------------------------------ producers
        // Producer GROUPS work on different stages at the same time (a stage's latency chain
        // -- wait, expand, proxy fence, arrive -- is much longer than its 8 x 32 MMA cycles at
        // NP=64): for NP <= 128 every warp is a group and fills whole stages on its own, for
        // NP = 256 the four warps fill one stage together.  Group g takes stages g, g+G, ...
        constexpr int kGroups = C::kGroups;
        constexpr int kGroupThreads = kProducerThreads / kGroups;
        constexpr int kRowsPerThread = NP / kGroupThreads;
        constexpr int kAhead = AHEAD;              // stages of global loads in flight per group
        const int group = threadIdx.x / kGroupThreads, tg = threadIdx.x % kGroupThreads;
        const bool vec_ok = ((stride_words & 3) == 0) && (((uintptr_t)bits & 15u) == 0);
        uint32_t wv[kAhead][kRowsPerThread][kKB];

        auto fetch = [&](uint32_t (&dst)[kRowsPerThread][kKB], size_t it) {
            const size_t w0 = (st0 + it) * kKB;
#pragma unroll
            for (int rr = 0; rr < kRowsPerThread; rr++) {
                const int row = tg + rr * kGroupThreads;
#pragma unroll
                for (int k = 0; k < kKB; k++) dst[rr][k] = 0;
                if (row < nsamples && it < nst) {
                    const uint32_t *src = bits + (size_t)row * stride_words + w0;
                    if (vec_ok && w0 + kKB <= words) {
#pragma unroll
                        for (int k = 0; k < kKB; k += 4) {
                            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(src + k));
                            dst[rr][k] = q.x; dst[rr][k + 1] = q.y; dst[rr][k + 2] = q.z; dst[rr][k + 3] = q.w;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < kKB; k++)
                            if (w0 + k < words) dst[rr][k] = __ldg(src + k);
                    }
                }
            }
        };
        auto produce = [&](const uint32_t (&src)[kRowsPerThread][kKB], size_t it) {
            const int s = (int)(it % C::kStages);
            const uint32_t phase = (uint32_t)((it / C::kStages) & 1);
            mbar_wait(smem_u32(&empty_bar[s]), phase ^ 1u);
            uint8_t *stage = smem + (size_t)s * C::kStageBytes;
#pragma unroll
            for (int rr = 0; rr < kRowsPerThread; rr++) {
                const int row = tg + rr * kGroupThreads;
                uint8_t *dst = stage + (size_t)(row >> 3) * 256 + (size_t)(row & 7) * 16;
#pragma unroll
                for (int k = 0; k < kKB; k++)
                    expand_word(src[rr][k], dst + (size_t)k * C::kTileBytes, 128);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(smem_u32(&full_bar[s]));
        };

#pragma unroll
        for (int p = 0; p < kAhead; p++) fetch(wv[p], (size_t)group + (size_t)p * kGroups);
        for (size_t it = group; it < nst; it += (size_t)kAhead * kGroups) {
#pragma unroll
            for (int p = 0; p < kAhead; p++) {
                const size_t cur = it + (size_t)p * kGroups;
                if (cur < nst) {
                    produce(wv[p], cur);
                    fetch(wv[p], cur + (size_t)kAhead * kGroups);
                }
            }
        }
    } else {
        // ------------------------------------------------------------ MMA issuer
        // The whole warp walks the pipeline (warp-uniform control flow keeps descriptors and
        // barrier addresses in uniform registers); one elected lane issues the instructions.
        const uint32_t idesc = make_idesc(C::kM, NP);
        const uint64_t desc0 = make_desc(smem_u32(smem), 128, 256);
        for (size_t it = 0; it < nst; it++) {
            const int s = (int)(it % C::kStages);
            const uint32_t phase = (uint32_t)((it / C::kStages) & 1);
            mbar_wait(smem_u32(&full_bar[s]), phase);
            __syncwarp();          // lanes leave the wait loop one by one; elect.sync wants the whole warp
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                // descriptor start address is in 16-byte units: stage and K-step are plain adds
                const uint64_t dstage = desc0 + (uint64_t)((s * C::kStageBytes) >> 4);
#pragma unroll
                for (int kb = 0; kb < kKB; kb++) {
                    const uint64_t bdesc = dstage + (uint64_t)((kb * C::kTileBytes) >> 4);
#pragma unroll
                    for (int mt = 0; mt < C::kMTiles; mt++)
                        mma_i8(tmem_base + mt * NP, bdesc + (uint64_t)((mt * 128 * 32) >> 4), bdesc, idesc,
                               (it | kb) ? 1u : 0u);
                }
                mma_commit(smem_u32(&empty_bar[s]));      // frees the stage when the MMAs retire
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(smem_u32(done_bar));
        __syncwarp();
    }

    if (warp < 4 && nst) {
        // ------------------------------------------------------------ epilogue
        mbar_wait(smem_u32(done_bar), 0);
        __syncwarp();              // tcgen05.ld below is .aligned
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int mt = 0; mt < C::kMTiles; mt++) {
            // M = 128: lane l of warp w holds row 32 w + l; M = 64: lanes 0..15 hold rows 16 w + l
            const int row_base = C::kM == 64 ? warp * 16 : mt * 128 + warp * 32;
            const bool lane_has_row = C::kM == 64 ? lane < 16 : true;
            const int row = row_base + lane;
            if (row_base >= nsamples) continue;                   // warp-uniform
#pragma unroll 1
            for (int c0 = 0; c0 < NP; c0 += 32) {
                if (c0 >= nsamples) break;
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * NP + c0);
                tmem_ld32(taddr, v);
                if (lane_has_row && row < nsamples) {
#pragma unroll
                    for (int c = 0; c < 32; c++)
                        if (c0 + c < nsamples && v[c])
                            atomicAdd(&gram[(size_t)row * nsamples + c0 + c], (unsigned long long)v[c]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "n"(C::kTmemCols));
    }
}

template <int NP, int AHEAD>
int launch_gram_i8(const uint32_t *bits, int nsamples, size_t words, size_t stride_words,
                   unsigned long long *gram, int device, cudaStream_t st) {
    using C = GramCfg<NP>;
    constexpr int kKB = C::kKB;
    PK_CUDA(cudaFuncSetAttribute(k_gram_i8<NP, AHEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));
    const size_t total_stages = (words + kKB - 1) / kKB;
    // a CTA's s32 accumulators must stay below 2^31: at most 2^31 / 32 words each
    const size_t min_ctas = (words + ((1ull << 25) - 1)) >> 25;
    size_t grid = (size_t)pk_sm_count(device);
    grid = std::max(grid, min_ctas);
    grid = std::max<size_t>(1, std::min(grid, total_stages));
    k_gram_i8<NP, AHEAD><<<(unsigned)grid, C::kThreads, C::kSmem, st>>>(bits, nsamples, words, stride_words, gram);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

}  // namespace

// gram (int64, nsamples x nsamples, already zeroed or holding a partial sum) += B * B^T
int pk_gram_i8_launch(const uint32_t *bits_dev, int nsamples, size_t words, size_t stride_words,
                      int64_t *gram_dev, int device, cudaStream_t st) {
    unsigned long long *g = reinterpret_cast<unsigned long long *>(gram_dev);
    // stages of bitmask loads a producer keeps in flight.  Measured (N=50: 6.60 / 6.54 / 6.49 ms at
    // 2 / 4 / 6, N=100: 12.0 / 13.3 / 13.6 ms, N=255: 34.5 / 36.6 ms at 2 / 3): the producers are not
    // load-latency bound (at 64 rows shared-memory bandwidth is: 2 KB stored + 4 KB read per MMA).
    int ahead = 0;
    if (const char *env = getenv("PYKMER_B200_GRAM_AHEAD")) ahead = atoi(env);
    if (nsamples <= 64) {
        if (ahead == 2) return launch_gram_i8<64, 2>(bits_dev, nsamples, words, stride_words, g, device, st);
        if (ahead == 4) return launch_gram_i8<64, 4>(bits_dev, nsamples, words, stride_words, g, device, st);
        return launch_gram_i8<64, 6>(bits_dev, nsamples, words, stride_words, g, device, st);
    }
    if (nsamples <= 128) {
        if (ahead == 4) return launch_gram_i8<128, 4>(bits_dev, nsamples, words, stride_words, g, device, st);
        return launch_gram_i8<128, 2>(bits_dev, nsamples, words, stride_words, g, device, st);
    }
    if (nsamples <= 256) {
        if (ahead == 3) return launch_gram_i8<256, 3>(bits_dev, nsamples, words, stride_words, g, device, st);
        return launch_gram_i8<256, 2>(bits_dev, nsamples, words, stride_words, g, device, st);
    }
    return pk_set_error(PK_ERR_ARG, "pk_gram_i8_launch: %d samples exceed one tensor-core tile (256)", nsamples);
}

// gram_f4.cu -- the merger's Gram matrix G = B * B^T (merger.py:136-176 over tools.py:473-482) on the
// block-scaled FP4 tensor path: the presence bits are expanded to 4-bit E2M1 numbers (tcgen05.mma
// kind::mxf4, K = 64 per instruction, all scale factors 1.0).  The default for <= 256 samples, fed
// with TILED masks (pk_gram_tiled_device); PYKMER_B200_GRAM=f4 runs it on row-major masks.
//
// Why not a byte per bit (gram_i8.cu): that kernel is bound by shared-memory traffic -- a K=32 step of
// R rows stores 32 R bytes and the MMAs read them back -- with the tensor pipe level with it at
// R = 256.  A nibble per bit halves the bytes per k-mer on both sides and the FP4 pipe runs at twice
// the int8 rate.  Measured (K=15): N=255 34.1 -> 17.0 ms, N=50 6.50 -> 3.69 ms.
//
// Why tiled masks: with bits[row][word] every row is its own stream and 148 CTAs x 256 rows keep
// ~38,000 of them open -- a DRAM page is opened for one 128-byte line, and the gather stops at
// 1.1 TB/s whatever the prefetch depth, load width or row stride (profiles/r01f_gram_sweep_*.txt).  With
// bits[word / 32][row][32 words] the lines of all rows for the same 1024 k-mers lie side by side and
// a CTA reads one sequential stream (tile_rows below).
//
// Exactness: a product is 0 or 1 and the FP32 accumulator of one CTA sees at most 2^24 k-mers
// (the launch sizes the grid for that), so every partial sum is an integer FP32 holds exactly --
// PROVIDED the tensor core adds into the full 24-bit significand.  That is a property of the
// hardware, not of the instruction set, and it is tested, not assumed: tests/ run this kernel
// against gram_i8 / the oracle / the reference's golden matrices, on all-ones masks (every partial
// sum from 64 up to 2^24 occurs: test_gram_f4_every_partial_sum_is_exact) and on random ones.
//
// Layout, roles and pipeline are those of gram_i8.cu.  Differences:
//   * one 32-bit mask word -> 16 bytes (32 nibbles, 0x2 = 1.0 in E2M1) by two masks and four
//     PRMT table look-ups (2 bits -> 1 byte); WHICH nibble a k-mer lands in is irrelevant as
//     long as A and B agree, and they are the same tile;
//   * the scale factors (UE8M0, 0x7F = 2^0) live in TMEM; every one of them is 1.0, so 32
//     columns are filled with 0x7F bytes once and every MMA points both operands at them --
//     whatever the layout, it reads ones;
//   * 129..256 samples use the symmetry: three 128 x 128 accumulators (0,0), (0,1), (1,1) instead
//     of two 128 x 256 ones -- 384 TMEM columns, which leaves room for the scale factors -- and the
//     epilogue mirrors (0,1) into (1,0).
#include <algorithm>
#include <stdlib.h>

#include "common.h"
#include "tcgen05_util.h"

namespace {

using namespace pk_umma;

// block-scaled instruction descriptor (kind::mxf4): A = B = E2M1 (format 1), both K-major,
// scale format UE8M0, N >> 3 at bit 17, M >> 4 at bit 24, K = 64 (bit 31 clear), SF ids 0
__host__ __device__ constexpr uint32_t make_idesc_f4(int m, int n) {
    return (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_f4(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate, uint32_t sfa_tmem, uint32_t sfb_tmem) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
        : "memory");
}

// 2 mask bits -> one byte of two E2M1 nibbles: 00 -> 0x00, 01 -> 0x02, 10 -> 0x20, 11 -> 0x22
constexpr uint32_t kPairLut = 0x22200200u;

// one bitmask word -> 16 bytes = 32 nibbles (one 16-byte K chunk of a row)
__device__ __forceinline__ uint4 expand_word_f4(uint32_t w) {
    const uint32_t s0 = w & 0x33333333u, s1 = (w >> 2) & 0x33333333u;       // PRMT selectors 0..3
    uint4 q;
    q.x = __byte_perm(kPairLut, 0u, s0);
    q.y = __byte_perm(kPairLut, 0u, s0 >> 16);
    q.z = __byte_perm(kPairLut, 0u, s1);
    q.w = __byte_perm(kPairLut, 0u, s1 >> 16);
    return q;
}

template <int NP>
struct F4Cfg {
    static constexpr int kProducerWarps = 8;
    static constexpr int kProducerThreads = kProducerWarps * 32;
    static constexpr int kThreads = kProducerThreads + 32;            // + the MMA-issuing warp
    // What bounds the producers is the gather of the mask words -- every row is its own stream,
    // so a warp's load touches 32 different lines -- and that gather runs at about one line MISS
    // per 8 cycles per SM however many loads are in flight (measured: 4 bytes per cycle per SM
    // with 32 bytes used per miss, tools/gram_sweep.py).  So a thread always takes a whole 128-byte
    // line of its row: for <= 128 rows a stage covers 16 K=64 steps (32 words = one line per row), for
    // 256 rows -- where such a stage would be 128 KB -- a thread reads a line into registers and feeds
    // four 4-step stages from it.
    static constexpr int kGroups = NP == 64 ? 4 : (NP == 128 ? 2 : 1);   // groups fill different stages; <= kStages
    static constexpr int kKB = NP == 256 ? 4 : 16;        // K=64 steps per stage (2 mask words per row each)
    static constexpr int kWords = 2 * kKB;                // mask words per row per stage
    static constexpr int kLineStages = 32 / kWords;       // stages one 128-byte line of a row feeds
    static constexpr int kStages = NP == 64 ? 4 : (NP == 128 ? 3 : 5);
    static constexpr int kTileBytes = NP * 32;            // one K=64 step of all NP rows (32 bytes per row)
    static constexpr int kStageBytes = kKB * kTileBytes;
    static constexpr int kAccN = NP == 256 ? 128 : NP;    // columns of one accumulator
    static constexpr int kAccs = NP == 256 ? 3 : 1;
    static constexpr int kSfCol = kAccs * kAccN;          // scale factors behind the accumulators
    static constexpr int kTmemCols = NP == 256 ? 512 : (NP == 128 ? 256 : 128);
    static constexpr int kPad = 4096;                     // the M=128 descriptor of a 64-row tile overruns
    static constexpr size_t kSmem = (size_t)kStages * kStageBytes + kPad + 256 + 1024;
    static_assert(kGroups <= kStages && kSmem <= 227 * 1024, "a group may not lap the ring; shared memory budget");
};

template <int NP, int AHEAD>
__global__ void __launch_bounds__(F4Cfg<NP>::kThreads, 1) k_gram_f4(const uint32_t *__restrict__ bits, int nsamples,
                                                        size_t words, size_t stride_words,
                                                        unsigned long long *__restrict__ gram, int diag, int tile_rows) {
    // tile_rows: 0 = bits[row][word] with a row stride; R > 0 = the tiled layout
    // bits[word / 32][R rows][32 words] (include/pykmer_b200.h) -- the lines of all rows for the same
    // 1024 k-mers lie side by side, so a CTA's gather is one sequential stream instead of one per row.
    // diag (PYKMER_B200_GRAM_DIAG, timing experiments only -- the result is then meaningless):
    // bit 0 = producers skip the global loads, bit 1 = the issuer skips the MMAs, bit 2 = producers
    // skip the shared-memory stores
    using C = F4Cfg<NP>;
    constexpr int kKB = C::kKB, kWords = C::kWords;
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

    // this CTA's slab of words, in whole stages of kWords words
    const size_t total_stages = (words + kWords - 1) / kWords;
    size_t per_cta = (total_stages + gridDim.x - 1) / gridDim.x;
    per_cta = (per_cta + C::kLineStages - 1) / C::kLineStages * C::kLineStages;     // slabs start on a line
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

    // scale factors: 32 columns of 0x7F bytes (UE8M0 2^0) on all 128 lanes; warp w owns lanes 32 w ..
    if (warp < 4) {
        const uint32_t one = 0x7F7F7F7Fu;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)C::kSfCol;
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
            "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(one)
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp < kMmaWarp) {
        // ------------------------------------------------------------ producers (see gram_i8.cu)
        constexpr int kGroups = C::kGroups;
        constexpr int kGroupThreads = kProducerThreads / kGroups;
        constexpr int kRowsPerThread = NP / kGroupThreads;
        constexpr int kAhead = AHEAD;              // stages of global loads in flight per group
        const int group = threadIdx.x / kGroupThreads, tg = threadIdx.x % kGroupThreads;
        const bool vec_ok = ((stride_words & 3) == 0) && (((uintptr_t)bits & 15u) == 0);
        uint32_t wv[kAhead][kRowsPerThread][kWords];
        auto row_ptr = [&](int row, size_t w0) -> const uint32_t * {       // w0: a multiple of kWords
            if (tile_rows) return bits + (w0 >> 5) * ((size_t)tile_rows * 32) + (size_t)row * 32 + (w0 & 31);
            return bits + (size_t)row * stride_words + w0;
        };

        auto fetch = [&](uint32_t (&dst)[kRowsPerThread][kWords], size_t it) {
            const size_t w0 = (st0 + it) * kWords;
#pragma unroll
            for (int rr = 0; rr < kRowsPerThread; rr++) {
                const int row = tg + rr * kGroupThreads;
#pragma unroll
                for (int k = 0; k < kWords; k++) dst[rr][k] = 0;
                if (diag & 1) continue;
                if (row < nsamples && it < nst) {
                    const uint32_t *src = row_ptr(row, w0);
                    if (vec_ok && w0 + kWords <= words) {
#pragma unroll
                        for (int k = 0; k < kWords; k += 4) {
                            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(src + k));
                            dst[rr][k] = q.x; dst[rr][k + 1] = q.y; dst[rr][k + 2] = q.z; dst[rr][k + 3] = q.w;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < kWords; k++)
                            if (w0 + k < words) dst[rr][k] = __ldg(src + k);
                    }
                }
            }
        };
        auto produce = [&](const uint32_t (&src)[kRowsPerThread][kWords], size_t it) {
            const int s = (int)(it % C::kStages);
            const uint32_t phase = (uint32_t)((it / C::kStages) & 1);
            mbar_wait(smem_u32(&empty_bar[s]), phase ^ 1u);
            uint8_t *stage = smem + (size_t)s * C::kStageBytes;
#pragma unroll
            for (int rr = 0; rr < kRowsPerThread; rr++) {
                const int row = tg + rr * kGroupThreads;
                uint8_t *dst = stage + (size_t)(row >> 3) * 256 + (size_t)(row & 7) * 16;
                if (diag & 4) continue;
#pragma unroll
                for (int k = 0; k < kKB; k++) {            // K=64 step k: words 2k, 2k+1 -> chunks 0, 1
                    *reinterpret_cast<uint4 *>(dst + (size_t)k * C::kTileBytes) = expand_word_f4(src[rr][2 * k]);
                    *reinterpret_cast<uint4 *>(dst + (size_t)k * C::kTileBytes + 128) =
                        expand_word_f4(src[rr][2 * k + 1]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(smem_u32(&full_bar[s]));
        };

        if constexpr (C::kLineStages == 1) {
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
            // one row per thread, one group: line L of the row = stages kLineStages * L ..
            static_assert(C::kLineStages == 1 || (kGroups == 1 && kRowsPerThread == 1), "line mode: one row per thread");
            constexpr int kLS = C::kLineStages;
            uint32_t line[2][kLS][1][kWords];
            auto fetch_line = [&](uint32_t (&dst)[kLS][1][kWords], size_t ln) {
                const size_t w0 = (st0 + ln * kLS) * kWords;
                const int row = tg;
#pragma unroll
                for (int j = 0; j < kLS; j++)
#pragma unroll
                    for (int k = 0; k < kWords; k++) dst[j][0][k] = 0;
                if ((diag & 1) || row >= nsamples || ln * kLS >= nst) return;
                const uint32_t *src = row_ptr(row, w0);
                if (vec_ok && w0 + kLS * kWords <= words) {
#pragma unroll
                    for (int j = 0; j < kLS; j++)
#pragma unroll
                        for (int k = 0; k < kWords; k += 4) {
                            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(src + j * kWords + k));
                            dst[j][0][k] = q.x; dst[j][0][k + 1] = q.y; dst[j][0][k + 2] = q.z; dst[j][0][k + 3] = q.w;
                        }
                } else {
#pragma unroll
                    for (int j = 0; j < kLS; j++)
#pragma unroll
                        for (int k = 0; k < kWords; k++)
                            if (w0 + j * kWords + k < words) dst[j][0][k] = __ldg(src + j * kWords + k);
                }
            };
            fetch_line(line[0], 0);
            fetch_line(line[1], 1);
            for (size_t ln = 0; ln * kLS < nst; ln += 2) {
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    const size_t s0 = (ln + b) * kLS;
                    if (s0 < nst) {
#pragma unroll
                        for (int j = 0; j < kLS; j++)
                            if (s0 + j < nst) produce(line[b][j], s0 + j);
                        fetch_line(line[b], ln + b + 2);
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------ MMA issuer
        const uint32_t idesc = make_idesc_f4(128, C::kAccN);
        const uint64_t desc0 = make_desc(smem_u32(smem), 128, 256);
        const uint32_t sf = tmem_base + (uint32_t)C::kSfCol;
        for (size_t it = 0; it < nst; it++) {
            const int s = (int)(it % C::kStages);
            const uint32_t phase = (uint32_t)((it / C::kStages) & 1);
            mbar_wait(smem_u32(&full_bar[s]), phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint64_t dstage = desc0 + (uint64_t)((s * C::kStageBytes) >> 4);
#pragma unroll
                for (int kb = 0; kb < kKB; kb++) {
                    const uint64_t lo = dstage + (uint64_t)((kb * C::kTileBytes) >> 4);   // rows 0..127
                    const uint32_t acc = (it | kb) ? 1u : 0u;
                    if (diag & 2) continue;
                    if (C::kAccs == 1) {
                        mma_f4(tmem_base, lo, lo, idesc, acc, sf, sf);
                    } else {
                        const uint64_t hi = lo + (uint64_t)((128 * 32) >> 4);            // rows 128..255
                        mma_f4(tmem_base, lo, lo, idesc, acc, sf, sf);                   // (0,0)
                        mma_f4(tmem_base + 128, lo, hi, idesc, acc, sf, sf);             // (0,1)
                        mma_f4(tmem_base + 256, hi, hi, idesc, acc, sf, sf);             // (1,1)
                    }
                }
                mma_commit(smem_u32(&empty_bar[s]));      // frees the stage when the MMAs retire
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(smem_u32(done_bar));
        __syncwarp();
    }

    if (warp < kMmaWarp && nst) {
        // ------------------------------------------------------------ epilogue
        // warp w reads TMEM lanes 32 (w % 4) ..: row (w % 4) * 32 + lane of an accumulator; the two
        // warps that share a lane quarter take alternate 32-column chunks
        mbar_wait(smem_u32(done_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3, half = warp >> 2;
        const int r = quarter * 32 + lane;
#pragma unroll 1
        for (int a = 0; a < C::kAccs; a++) {
            const int row0 = a == 2 ? 128 : 0, col0 = a == 0 ? 0 : (C::kAccs == 1 ? 0 : 128);
            if (row0 + quarter * 32 >= nsamples) continue;            // warp-uniform
#pragma unroll 1
            for (int c0 = half * 32; c0 < C::kAccN; c0 += 64) {
                if (col0 + c0 >= nsamples) break;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * C::kAccN + c0), v);
                const int row = row0 + r;
                if (row < nsamples) {
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        const int col = col0 + c0 + c;
                        const unsigned long long x = __float2ull_rn(__uint_as_float(v[c]));
                        if (col < nsamples && x) {
                            atomicAdd(&gram[(size_t)row * nsamples + col], x);
                            if (a == 1) atomicAdd(&gram[(size_t)col * nsamples + row], x);   // mirror of (0,1)
                        }
                    }
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
int launch_gram_f4(const uint32_t *bits, int nsamples, size_t words, size_t stride_words,
                   unsigned long long *gram, int device, cudaStream_t st, int tile_rows) {
    using C = F4Cfg<NP>;
    PK_CUDA(cudaFuncSetAttribute(k_gram_f4<NP, AHEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));
    const size_t total_stages = (words + C::kWords - 1) / C::kWords;
    // a CTA's FP32 accumulators must stay exact integers: at most 2^24 k-mers = 2^19 words each
    // (slabs are rounded up to whole lines and stages inside the kernel: keep 64 words of margin)
    const size_t slab_max = (1ull << 19) - 64;
    const size_t min_ctas = (words + slab_max - 1) / slab_max;
    size_t grid = (size_t)pk_sm_count(device);
    grid = std::max(grid, min_ctas);
    grid = std::max<size_t>(1, std::min(grid, total_stages));
    int diag = 0;
    if (const char *env = getenv("PYKMER_B200_GRAM_DIAG")) diag = atoi(env);
    if (tile_rows && (words & 31))
        return pk_set_error(PK_ERR_ARG, "gram_f4: the tiled layout needs a multiple of 32 words, got %zu", words);
    k_gram_f4<NP, AHEAD><<<(unsigned)grid, C::kThreads, C::kSmem, st>>>(bits, nsamples, words, stride_words, gram,
                                                                        diag, tile_rows);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

}  // namespace

// gram (int64, nsamples x nsamples, already zeroed or holding a partial sum) += B * B^T
int pk_gram_f4_launch(const uint32_t *bits_dev, int nsamples, size_t words, size_t stride_words,
                      int64_t *gram_dev, int device, cudaStream_t st, int tile_rows) {
    unsigned long long *g = reinterpret_cast<unsigned long long *>(gram_dev);
    if (nsamples <= 64) return launch_gram_f4<64, 2>(bits_dev, nsamples, words, stride_words, g, device, st, tile_rows);
    if (nsamples <= 128) return launch_gram_f4<128, 2>(bits_dev, nsamples, words, stride_words, g, device, st, tile_rows);
    if (nsamples <= 256) return launch_gram_f4<256, 2>(bits_dev, nsamples, words, stride_words, g, device, st, tile_rows);
    return pk_set_error(PK_ERR_ARG, "pk_gram_f4_launch: %d samples exceed one tensor-core tile (256)", nsamples);
}

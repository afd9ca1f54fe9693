// gram_f4.cu -- the merger's Gram matrix G = B * B^T (merger.py:136-176 over tools.py:473-482) on the
// block-scaled FP4 tensor path: the presence bits are expanded to 4-bit E2M1 numbers (tcgen05.mma
// kind::mxf4, K = 64 per instruction, all scale factors 1.0).  The merger's default, fed with TILED
// masks (pk_gram_tiled_device): bits[word / 32][row][32 words], so that the 128-byte lines of all
// rows for the same 1024 k-mers lie side by side and a CTA reads one sequential stream (with
// row-major masks 148 CTAs x 256 rows keep ~38,000 DRAM streams open and the gather stops at
// 1.1 TB/s whatever the prefetch depth, load width or row stride: profiles/r01f_gram_sweep_*.txt).
//
// Round 2: what bounded round 1's kernel was SHARED-MEMORY bandwidth, not the tensor pipe.  With both
// operands in shared memory a K=64 step of 256 rows stores 8 KB and its three MMAs read 24 KB back
// (A and B of each), 256 cycles at 128 B/clk against 192 cycles of tensor pipe (measured: 300); at
// <= 64 samples the M=128 instruction read 4 KB of A, half of it padding, + 2 KB of B per 2 KB stored
// (64 cycles, measured 65, against 32 of tensor pipe).  So now
//   * the A operand comes from TENSOR MEMORY: the producer thread that expands row r of a tile is
//     the thread that owns TMEM lane r (warp w reaches lanes 32 (w % 4) ..), so it writes the same
//     16 + 16 bytes it stores to shared memory (the B operand) into 8 TMEM columns of its lane with
//     tcgen05.st; the MMAs then read only B from shared memory (12 KB instead of 24 KB per step at
//     256 rows);
//   * <= 64 samples run as a DUAL slab: lanes 0..63 carry the samples over the first half of the
//     CTA's k-mers, lanes 64..127 the same samples over the second half, one M=128 x N=128
//     instruction covers two K=64 steps, and the epilogue adds the two diagonal 64 x 64 blocks -- no
//     operand row is padding;
//   * the row blocks are arguments (any two 128-row blocks of one tiled buffer), so more than 256
//     samples run on the tensor cores too, block pair by block pair (pk_gram_f4_launch).
//
// Exactness: a product is 0 or 1 and an FP32 accumulator sees at most 2^24 k-mers (the launch sizes
// the grid for that), so every partial sum is an integer FP32 holds exactly -- PROVIDED the tensor
// core adds into the full 24-bit significand.  That is a property of the hardware, not of the
// instruction set: pk_gram_f4_exact() checks it once per device before the first launch (one CTA
// driven through every integer from 2^23 to 2^24 - 2^15, odd ones included) and the merger falls back
// to the integer kernels if it ever fails; tests/ run this kernel against gram_i8, AND + popcount, the
// oracle and the reference's golden matrices, at BASELINE size too (tests/test_gpu_at_scale.py).
//
// Which nibble a k-mer lands in is irrelevant as long as A and B agree: both are written from the
// same registers (one 32-bit mask word -> 16 bytes = 32 nibbles, 0x2 = 1.0 in E2M1, by two masks and
// four PRMT look-ups), 16 bytes = one K chunk of the shared-memory core matrix = 4 TMEM columns.
// The scale factors (UE8M0, 0x7F = 2^0) live in TMEM; every one of them is 1.0, so 32 columns are
// filled with 0x7F bytes once and every MMA points both operands at them.
#include <algorithm>
#include <mutex>
#include <stdlib.h>
#include <vector>

#include <cuda.h>

#include "common.h"
#include "tcgen05_util.h"

namespace {

using namespace pk_umma;

// block-scaled instruction descriptor (kind::mxf4): A = B = E2M1 (format 1), both K-major,
// scale format UE8M0, N >> 3 at bit 17, M >> 4 at bit 24, K = 64 (bit 31 clear), SF ids 0
__host__ __device__ constexpr uint32_t make_idesc_f4(int m, int n) {
    return (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((uint32_t)(m >> 4) << 24);
}

// D[tmem] (+)= A[smem descriptor] * B[smem descriptor]^T
__device__ __forceinline__ void mma_f4_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate, uint32_t sfa_tmem, uint32_t sfb_tmem) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
        : "memory");
}

// D[tmem] (+)= A[tmem: lane = row, 8 columns = 64 E2M1] * B[smem descriptor]^T
__device__ __forceinline__ void mma_f4_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate, uint32_t sfa_tmem, uint32_t sfb_tmem) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%5], [%6], p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
        : "memory");
}

// 8 consecutive TMEM columns of this thread's lane <- two expanded mask words
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4 a, const uint4 b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}

// 2 mask bits -> one byte of two E2M1 nibbles: 00 -> 0x00, 01 -> 0x02, 10 -> 0x20, 11 -> 0x22
constexpr uint32_t kPairLut = 0x22200200u;

// one bitmask word -> 16 bytes = 32 nibbles (one 16-byte K chunk of a row): 2 LOP3 + 3 SHF + 4 PRMT, all on
// the ALU pipe (one warp instruction per two cycles per scheduler) -- what the producers spend most of
// their issue slots on.  Moving the three shifts to the FMA pipe as high multiplies (IMAD.HI) was measured
// and is slower (N=255: 268 against 243 cycles per step, profiles/r02h_gram_sweep_imad_hi_shifts_lost.txt).
__device__ __forceinline__ uint4 expand_word_f4(uint32_t w) {
    const uint32_t s0 = w & 0x33333333u, s1 = (w >> 2) & 0x33333333u;       // PRMT selectors 0..3
    uint4 q;
    q.x = __byte_perm(kPairLut, 0u, s0);
    q.y = __byte_perm(kPairLut, 0u, s0 >> 16);
    q.z = __byte_perm(kPairLut, 0u, s1);
    q.w = __byte_perm(kPairLut, 0u, s1 >> 16);
    return q;
}

// One TILE = 32 mask words per row = 1024 k-mers = 16 K=64 steps.  The raw tile (all staged rows x 128
// bytes) is fetched by TMA (cp.async.bulk.tensor, 64-row boxes, SWIZZLE_128B) into a ring in shared
// memory; a producer thread owns one row of the staged block, reads ITS 128-byte line back (the swizzle
// makes a warp's reads of 32 different rows conflict-free) and expands it.
//   NP = 256: rows = two 128-row blocks (lo, hi); accumulators (lo,lo) | (lo,hi) -- ONE N=256 instruction with
//             A = lo from tensor memory -- and (hi,hi), whose A operand stays in shared memory: the 96 TMEM
//             columns left beside 384 of accumulators and 32 of scale factors hold three 4-step stages of
//             the lo block, and what bounds the kernel is how many stages are in flight, not their bytes.
//   NP = 128: one 128-row block, one accumulator, A from tensor memory.
//             DUAL: the block is [<= 64 samples over the first half of the CTA's tiles; the same samples
//             over the second half].
// A stage is big (4 / 8 K=64 steps): the producers' chain per stage -- wait for the slot, store, drain the
// tensor-memory and shared-memory stores, proxy fence, arrive -- costs several hundred cycles whatever
// the stage holds (measured, profiles/r02b_gram_sweep_v3_tma_3stage_ring.txt: 2-step stages ran at 220
// cycles per step with NO MMAs).  Two producer groups of NP threads take alternate stages, so one
// expands while the other sits in that chain.
// Warps: producers, then one MMA-issuing warp, then one TMA-issuing warp.
constexpr int kTileWords = 32;
constexpr int kTileSteps = 16;                          // K=64 steps per tile
constexpr int kBoxRows = 64;                            // rows per TMA box
constexpr int kBoxBytes = kBoxRows * 128;

template <int NP>
struct F4Cfg {
    // Two producer groups of NP threads (each reaches all the TMEM lanes of its rows); group g takes
    // stages g, g + 2, ...  NP = 128 with FOUR groups and 4-step stages was measured and is slower
    // (N=128: 110 against 100 cycles per step, profiles/r02g_gram_sweep_np128_four_groups.txt): the stage
    // handshake, not the number of warps, is what a smaller stage makes worse.
    static constexpr int kGroups = 2;
    static constexpr int kGroupWarps = NP / 32;                       // one thread per staged row
    static constexpr int kGroupThreads = kGroupWarps * 32;
    static constexpr int kProducerWarps = kGroups * kGroupWarps;      // 16 / 8
    static constexpr int kMmaWarp = kProducerWarps;                   // the TMA warp follows it
    static constexpr int kThreads = (kProducerWarps + 2) * 32;        // 576 / 320
    static constexpr int kKB = NP == 256 ? 4 : 8;                     // K=64 steps per stage
    static constexpr int kStageChunks = kKB / 2;                      // 16-byte chunks of a line per stage
    static constexpr int kTileStages = kTileSteps / kKB;              // stages one tile feeds: 4 / 2
    static constexpr int kMyStages = kTileStages / kGroups;           // ... of which a group takes 2 / 1
    static constexpr int kStages = NP == 256 ? 3 : 4;
    static constexpr int kStepBytes = NP * 32;                        // one K=64 step of all NP rows
    static constexpr int kStageBytes = kKB * kStepBytes;              // 32 KB either way
    static constexpr int kRawBytes = NP * 128;                        // one raw tile: 32 KB / 16 KB
    static constexpr int kRawSlots = NP == 256 ? 3 : 4;
    static constexpr int kRawReaders = kGroups * kGroupThreads;       // every producer reads part of every raw tile
    static constexpr int kAccs = NP == 256 ? 3 : 1;
    static constexpr int kSfCol = kAccs * 128;                        // scale factors behind the accumulators
    static constexpr int kACol = kSfCol + 32;                         // ring of A operands (the lo block) behind them
    static constexpr int kAStageCols = kKB * 8;                       // 8 columns per step
    static constexpr int kTmemCols = 512;
    static constexpr size_t kSmem = (size_t)kStages * kStageBytes + (size_t)kRawSlots * kRawBytes + 512 + 1024;
    static_assert(kACol + kStages * kAStageCols <= kTmemCols, "TMEM budget");
    static_assert(kTileStages % kGroups == 0 && kStages >= kGroups - 1 && kSmem <= 227 * 1024, "ring depth / shared memory budget");
};

// 32 arrivals of one warp on one mbarrier word serialise like shared-memory atomics on one address; one
// arrival per warp (every lane fences, __syncwarp, lane 0 arrives) removes that.  Measured per kernel form
// (profiles/r02l_gram_sweep_arrivals.txt, K=13): 129..256 rows 3.85 -> 3.50 ms with both barriers per warp
// (one of the two alone: 3.67 / 3.84); <= 128 rows 1.42 per thread against 1.54 per warp -- there the
// __syncwarp costs more than the arrivals, the stages are half as long.
__host__ __device__ constexpr bool arrive_per_warp(int np) { return np > 128; }

struct GramArgs {
    CUtensorMap tmap;          // tiled masks as a 3-D tensor {32 words, tile_rows, tiles}, box {32, 64, 1}, SWIZZLE_128B
    size_t tiles;              // tiles per row (words / 32)
    int tile_rows;             // rows of the tiled buffer
    int row_lo, n_lo;          // first 128-row block: rows [row_lo, row_lo + n_lo)
    int row_hi, n_hi;          // second block (NP = 256)
    int acc_mask;              // NP = 256: bit 0 (lo,lo), bit 1 (lo,hi), bit 2 (hi,hi) -- which blocks to compute
    unsigned long long *gram;  // [ld][ld]
    int ld;
    size_t tiles_per_cta;
    int diag;                  // PYKMER_B200_GRAM_DIAG (timing experiments only -- the result is then meaningless):
                               // bit 0 = no TMA loads, bit 1 = no MMAs, bit 2 = no operand stores,
                               // bit 3 = (lo,lo) and (lo,hi) as two N=128 instructions instead of one N=256
};

__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
    return v;
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// one 64-row x 128-byte box of tile `tile`, rows [row, row + 64) (rows past the buffer read as zeros)
__device__ __forceinline__ void tma_load_box(uint32_t dst, const CUtensorMap *tmap, int row, int tile, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tmap), "r"(0), "r"(row), "r"(tile), "r"(bar)
        : "memory");
}

template <int NP, bool DUAL>
__global__ void __launch_bounds__(F4Cfg<NP>::kThreads, 1) k_gram_f4(const __grid_constant__ GramArgs g) {
    using C = F4Cfg<NP>;
    static_assert(!DUAL || NP == 128, "the dual slab is a form of the one-block kernel");
    constexpr int kKB = C::kKB;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *raw = smem + (size_t)C::kStages * C::kStageBytes;             // [kRawSlots][kRawBytes], 1024-aligned
    uint8_t *ctrl = raw + (size_t)C::kRawSlots * C::kRawBytes;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(ctrl);               // [kStages]
    uint64_t *empty_bar = full_bar + C::kStages;                            // [kStages]
    uint64_t *raw_full = empty_bar + C::kStages;                            // [kRawSlots]
    uint64_t *raw_empty = raw_full + C::kRawSlots;                          // [kRawSlots]
    uint64_t *done_bar = raw_empty + C::kRawSlots;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // one mbarrier arrival per producer WARP (after __syncwarp) or one per thread, see arrive_per_warp()
    constexpr bool warp_full = arrive_per_warp(NP), warp_raw = arrive_per_warp(NP);

    // this CTA's tiles [t0, t1); DUAL: first half [t0, tm) on lanes 0..63, second half [tm, t1) on 64..127
    const size_t t0 = min(g.tiles, (size_t)blockIdx.x * g.tiles_per_cta);
    const size_t t1 = min(g.tiles, t0 + g.tiles_per_cta);
    const size_t tm = DUAL ? t0 + (t1 - t0 + 1) / 2 : t1;
    const size_t ntiles = tm - t0;                                          // tiles this CTA steps through
    const size_t nst = ntiles * C::kTileStages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; s++) {
            mbar_init(smem_u32(&full_bar[s]), warp_full ? C::kGroupWarps : C::kGroupThreads);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int s = 0; s < C::kRawSlots; s++) {
            mbar_init(smem_u32(&raw_full[s]), 1);
            mbar_init(smem_u32(&raw_empty[s]), warp_raw ? C::kRawReaders / 32 : C::kRawReaders);
        }
        mbar_init(smem_u32(done_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == C::kMmaWarp) {
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

    if (warp < C::kProducerWarps) {
        // ------------------------------------------------------------ producers
        // thread -> staged row srow; TMEM lane r = srow % 128 lies in the quarter (warp % 4) this warp reaches
        const int group = warp / C::kGroupWarps, wg = warp % C::kGroupWarps;
        const int quarter = wg & 3, half = wg >> 2;                         // half: lo / hi block (NP = 256)
        const int r = quarter * 32 + lane;
        bool valid;
        size_t my_tiles = ntiles;
        if (NP == 256) valid = r < (half ? g.n_hi : g.n_lo);
        else if (DUAL) { valid = (r & 63) < g.n_lo; if (r >= 64) my_tiles = t1 - tm; }
        else valid = r < g.n_lo;
        const int srow = NP == 256 ? half * 128 + r : r;                    // row inside the staged step / raw tile
        const uint32_t soff = (uint32_t)(srow >> 3) * 256u + (uint32_t)(srow & 7) * 16u;
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)C::kACol;
        const bool to_tmem = NP == 128 || half == 0;                        // the lo block is the A operand in TMEM
        // this thread's line inside a raw slot: box srow / 64, row srow % 64; chunk c sits at c ^ (row & 7)
        const uint32_t roff = (uint32_t)(srow / kBoxRows) * kBoxBytes + (uint32_t)(srow % kBoxRows) * 128u;
        const uint32_t rxor = (uint32_t)(srow & 7);

        auto produce_stage = [&](size_t it, const uint4 *ch) {              // ch: kStageChunks chunks = kKB steps
            const int s = (int)(it % C::kStages);
            const uint32_t phase = (uint32_t)((it / C::kStages) & 1);
            mbar_wait(smem_u32(&empty_bar[s]), phase ^ 1u);
            __syncwarp();          // lanes leave the wait loop one by one; tcgen05.st / wait::st below are .aligned
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint8_t *dst = smem + (size_t)s * C::kStageBytes + soff;
            const uint32_t tdst = tlane + (uint32_t)(s * C::kAStageCols);
            if (!(g.diag & 4)) {
#pragma unroll
                for (int k = 0; k < kKB; k++) {            // K=64 step k of the stage: two words -> chunks 0, 1
                    const uint4 q = ch[k >> 1];
                    const uint4 c0 = expand_word_f4((k & 1) ? q.z : q.x);
                    const uint4 c1 = expand_word_f4((k & 1) ? q.w : q.y);
                    *reinterpret_cast<uint4 *>(dst + (size_t)k * C::kStepBytes) = c0;
                    *reinterpret_cast<uint4 *>(dst + (size_t)k * C::kStepBytes + 128) = c1;
                    if (to_tmem) tmem_st8(tdst + (uint32_t)(k * 8), c0, c1);
                }
            }
            if (to_tmem) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            // every lane has drained and fenced its own stores; ONE arrival per warp -- 32 arrivals on one
            // mbarrier word serialise like any shared-memory atomic on one address
            if (warp_full) {
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
            } else {
                mbar_arrive(smem_u32(&full_bar[s]));
            }
        };

        constexpr int kMyChunks = C::kMyStages * C::kStageChunks;           // chunks of a line this thread expands: 4
        for (size_t t = 0; t < ntiles; t++) {
            const int rs = (int)(t % C::kRawSlots);
            mbar_wait(smem_u32(&raw_full[rs]), (uint32_t)((t / C::kRawSlots) & 1));
            __syncwarp();
            uint4 ch[kMyChunks];
            const uint32_t line = smem_u32(raw) + (uint32_t)rs * C::kRawBytes + roff;
            const bool have = valid && t < my_tiles && !(g.diag & 1);
#pragma unroll
            for (int m = 0; m < kMyChunks; m++) {
                // my m-th chunk: stage (m / kStageChunks) * 2 + group of the tile, chunk m % kStageChunks of it
                const uint32_t c = (uint32_t)(((m / C::kStageChunks) * C::kGroups + group) * C::kStageChunks +
                                              m % C::kStageChunks);
                ch[m] = have ? lds128(line + ((c ^ rxor) << 4)) : make_uint4(0, 0, 0, 0);
            }
            // The slot may be refilled -- once the loads above have RETURNED.  mbarrier.arrive does not wait
            // for outstanding loads by itself (SASS: LD ... SYNCS.ARRIVE back to back), and a TMA write that
            // overtakes them showed up as one stale row in a few launches out of ten (round 2, pass c).
            __threadfence_block();
            if (warp_raw) {
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&raw_empty[rs]));
            } else {
                mbar_arrive(smem_u32(&raw_empty[rs]));
            }
#pragma unroll
            for (int m = 0; m < C::kMyStages; m++)
                produce_stage(t * C::kTileStages + (size_t)(m * C::kGroups + group), &ch[m * C::kStageChunks]);
        }
    } else if (warp == C::kMmaWarp) {
        // ------------------------------------------------------------ MMA issuer
        const uint32_t idesc = make_idesc_f4(128, 128), idesc256 = make_idesc_f4(128, 256);
        const uint64_t desc0 = make_desc(smem_u32(smem), 128, 256);
        const uint32_t sf = tmem_base + (uint32_t)C::kSfCol;
        const uint32_t a0 = tmem_base + (uint32_t)C::kACol;
        const bool fused = NP == 256 && (g.acc_mask & 3) == 3 && !(g.diag & 8);
        for (size_t it = 0; it < nst; it++) {
            const int s = (int)(it % C::kStages);
            const uint32_t phase = (uint32_t)((it / C::kStages) & 1);
            mbar_wait(smem_u32(&full_bar[s]), phase);
            __syncwarp();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint64_t dstage = desc0 + (uint64_t)((s * C::kStageBytes) >> 4);
                const uint32_t astage = a0 + (uint32_t)(s * C::kAStageCols);
#pragma unroll
                for (int kb = 0; kb < kKB; kb++) {
                    const uint64_t lo = dstage + (uint64_t)((kb * C::kStepBytes) >> 4);   // rows 0..127
                    const uint32_t alo = astage + kb * 8;
                    const uint32_t acc = (it | kb) ? 1u : 0u;
                    if (g.diag & 2) continue;
                    if (NP == 128) {
                        mma_f4_ts(tmem_base, alo, lo, idesc, acc, sf, sf);
                    } else {
                        const uint64_t hi = lo + (uint64_t)((128 * 32) >> 4);            // rows 128..255
                        if (fused) {
                            mma_f4_ts(tmem_base, alo, lo, idesc256, acc, sf, sf);        // (lo,lo) | (lo,hi): B = all 256 rows
                        } else {
                            if (g.acc_mask & 1) mma_f4_ts(tmem_base, alo, lo, idesc, acc, sf, sf);         // (lo,lo)
                            if (g.acc_mask & 2) mma_f4_ts(tmem_base + 128, alo, hi, idesc, acc, sf, sf);   // (lo,hi)
                        }
                        if (g.acc_mask & 4) mma_f4_ss(tmem_base + 256, hi, hi, idesc, acc, sf, sf);        // (hi,hi)
                    }
                }
                mma_commit(smem_u32(&empty_bar[s]));      // frees the stage when the MMAs retire
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(smem_u32(done_bar));
        __syncwarp();
    } else {
        // ------------------------------------------------------------ TMA issuer: raw tiles into the ring
        if (lane == 0 && !(g.diag & 1)) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmap) : "memory");
            const int nb_lo = (min(g.n_lo, 128) + kBoxRows - 1) / kBoxRows;
            const int nb_hi = NP == 256 ? (min(g.n_hi, 128) + kBoxRows - 1) / kBoxRows : 0;
            for (size_t t = 0; t < ntiles; t++) {
                const int rs = (int)(t % C::kRawSlots);
                mbar_wait(smem_u32(&raw_empty[rs]), (uint32_t)(((t / C::kRawSlots) & 1) ^ 1));
                const uint32_t bar = smem_u32(&raw_full[rs]);
                const uint32_t dst = smem_u32(raw + (size_t)rs * C::kRawBytes);
                const int tile = (int)(t0 + t);
                if (DUAL) {
                    const bool second = tm + t < t1;
                    mbar_expect_tx(bar, (second ? 2u : 1u) * kBoxBytes);
                    tma_load_box(dst, &g.tmap, g.row_lo, tile, bar);
                    if (second) tma_load_box(dst + kBoxBytes, &g.tmap, g.row_lo, (int)(tm + t), bar);
                } else {
                    mbar_expect_tx(bar, (uint32_t)(nb_lo + nb_hi) * kBoxBytes);
                    for (int b = 0; b < nb_lo; b++)
                        tma_load_box(dst + b * kBoxBytes, &g.tmap, g.row_lo + b * kBoxRows, tile, bar);
                    for (int b = 0; b < nb_hi; b++)
                        tma_load_box(dst + (2 + b) * kBoxBytes, &g.tmap, g.row_hi + b * kBoxRows, tile, bar);
                }
            }
        } else if (lane == 0) {
            // diag & 1: no loads -- just hand the (uninitialised) slots round
            for (size_t t = 0; t < ntiles; t++) {
                const int rs = (int)(t % C::kRawSlots);
                mbar_wait(smem_u32(&raw_empty[rs]), (uint32_t)(((t / C::kRawSlots) & 1) ^ 1));
                mbar_arrive(smem_u32(&raw_full[rs]));
            }
        }
        __syncwarp();
    }

    if (warp < C::kProducerWarps && nst && !(g.diag & 2)) {
        // ------------------------------------------------------------ epilogue
        // a producer warp reads the TMEM lanes of its quarter: row quarter * 32 + lane of an accumulator; the
        // warps that share a quarter take different 32-column chunks
        mbar_wait(smem_u32(done_bar), 0);
        __syncwarp();              // tcgen05.ld below is .aligned
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int group = warp / C::kGroupWarps, wg = warp % C::kGroupWarps;
        const int quarter = wg & 3;
        // NP = 256: four warps share a quarter (lo/hi half x 2 groups), one 32-column chunk each; NP = 128: two
        const int cstart = NP == 256 ? 32 * ((wg >> 2) + 2 * group) : 32 * group;
        constexpr int cstep = NP == 256 ? 128 : 64;
        const int r = quarter * 32 + lane;
        const size_t ld = (size_t)g.ld;
        if (NP == 128 && DUAL) {
            // rows 0..63 x columns 0..63 and rows 64..127 x columns 64..127 are the two halves' Gram blocks
            const int i = r & 63;
            const int c0 = (r >= 64 ? 64 : 0) + cstart;
            if ((c0 & 63) < g.n_lo) {                                   // warp-uniform
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
                if (i < g.n_lo) {
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        const int j = (c0 & 63) + c;
                        const unsigned long long x = __float2ull_rn(__uint_as_float(v[c]));
                        if (j < g.n_lo && x) atomicAdd(&g.gram[(size_t)(g.row_lo + i) * ld + g.row_lo + j], x);
                    }
                }
            }
        } else {
#pragma unroll 1
            for (int a = 0; a < C::kAccs; a++) {
                if (NP == 256 && !((g.acc_mask >> a) & 1)) continue;
                const int rbase = a == 2 ? g.row_hi : g.row_lo, nr = a == 2 ? g.n_hi : g.n_lo;
                const int cbase = a == 0 ? g.row_lo : g.row_hi, nc = a == 0 ? g.n_lo : g.n_hi;
                if (quarter * 32 >= nr) continue;                        // warp-uniform
#pragma unroll 1
                for (int c0 = cstart; c0 < 128; c0 += cstep) {
                    if (c0 >= nc) break;
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * 128 + c0), v);
                    if (r < nr) {
#pragma unroll
                        for (int c = 0; c < 32; c++) {
                            const unsigned long long x = __float2ull_rn(__uint_as_float(v[c]));
                            if (c0 + c < nc && x) {
                                atomicAdd(&g.gram[(size_t)(rbase + r) * ld + cbase + c0 + c], x);
                                if (a == 1) atomicAdd(&g.gram[(size_t)(cbase + c0 + c) * ld + rbase + r], x);   // mirror of (lo,hi)
                            }
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == C::kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "n"(C::kTmemCols));
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*pk_encode_tiled_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);

pk_encode_tiled_t encode_tiled() {
    static pk_encode_tiled_t fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<pk_encode_tiled_t>(p);
    }();
    return fn;
}

int make_tmap(CUtensorMap *map, const uint32_t *bits, int tile_rows, size_t tiles) {
    pk_encode_tiled_t enc = encode_tiled();
    if (!enc) return pk_set_error(PK_ERR_CUDA, "gram_f4: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {kTileWords, (cuuint64_t)tile_rows, (cuuint64_t)tiles};
    const cuuint64_t strides[2] = {128, (cuuint64_t)tile_rows * 128};       // bytes, dimensions 1 and 2
    const cuuint32_t box[3] = {kTileWords, kBoxRows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    // L2 promotion no wider than the 128-byte lines: a tile starts at tile_rows * 128 bytes, which is only
    // 128-byte aligned when the number of samples is odd, and with 256-byte promotion the FIRST line of a
    // box then came back stale in about half of the launches (N = 253, 255: row 0 of the Gram matrix off
    // by a few counts; N = 254, 256 never; none / 128-byte promotion never -- 8 launches each, round 2)
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint32_t *>(bits), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return pk_set_error(PK_ERR_CUDA, "gram_f4: cuTensorMapEncodeTiled failed (%d) for %d rows x %zu tiles", (int)r,
                            tile_rows, tiles);
    return PK_OK;
}

template <int NP, bool DUAL>
int launch_one(GramArgs &a, int device, cudaStream_t st, size_t force_grid) {
    using C = F4Cfg<NP>;
    // an FP32 accumulator must stay an exact integer: at most 2^24 k-mers = 2^14 tiles each (the dual
    // slab has one accumulator block per half)
    const size_t cap = (size_t)(DUAL ? 2 : 1) << 14;
    size_t grid = std::max((size_t)pk_sm_count(device), (a.tiles + cap - 1) / cap);
    grid = std::max<size_t>(1, std::min(grid, a.tiles));
    if (force_grid) grid = force_grid;
    a.tiles_per_cta = (a.tiles + grid - 1) / grid;
    if (a.tiles_per_cta > cap)
        return pk_set_error(PK_ERR_ARG, "gram_f4: %zu tiles per CTA exceed the exact range of an FP32 accumulator", a.tiles_per_cta);
    a.diag = 0;
    if (const char *env = getenv("PYKMER_B200_GRAM_DIAG")) a.diag = atoi(env);
    PK_CUDA(cudaFuncSetAttribute(k_gram_f4<NP, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem));
    k_gram_f4<NP, DUAL><<<(unsigned)grid, C::kThreads, C::kSmem, st>>>(a);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

// ---- the once-per-device exactness check ------------------------------------------------------------
// Three rows over 2^14 - 32 tiles (16,744,448 k-mers), ONE CTA, one accumulator:
//   A  2^23 ones, then exactly one set bit per K=64 instruction      B  all ones
//   C  2^23 ones, then one set bit (A's) every third instruction
// (A,A) and (A,B) walk through every integer from 2^23 to 2^24 - 32768 - 1, (A,C), (C,C) through every
// third; the closed forms below are what exact accumulation gives.
constexpr size_t kCheckTiles = (1u << 14) - 32;
constexpr size_t kCheckOnes = 1u << 13;                  // tiles of ones = 2^23 k-mers

__global__ void k_selfcheck_fill(uint32_t *bits) {
    // tile t, row r, word k  ->  bits[(t * 3 + r) * 32 + k]
    const size_t n = kCheckTiles * 3 * 32;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t t = i / 96;
        const int r = (int)((i / 32) % 3), k = (int)(i % 32);
        uint32_t w = 0xFFFFFFFFu;
        if (r != 1 && t >= kCheckOnes) {
            const size_t step = (t - kCheckOnes) * 16 + (size_t)(k >> 1);      // K=64 step after the ones
            w = (k & 1) ? 0u : 1u << (step & 31);
            if (r == 2 && step % 3 != 0) w = 0u;
        }
        bits[i] = w;
    }
}

std::mutex g_check_mutex;
int g_exact[64];                                         // 0 = not checked, 1 = exact, -1 = NOT exact

int selfcheck(int device) {
    uint32_t *bits = nullptr;
    unsigned long long *gram = nullptr, h[9];
    cudaStream_t st = nullptr;
    PK_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaError_t e = cudaMalloc(&bits, kCheckTiles * 3 * 32 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&gram, sizeof h);
    if (e == cudaSuccess) e = cudaMemsetAsync(gram, 0, sizeof h, st);
    int rc = PK_OK;
    if (e == cudaSuccess) {
        k_selfcheck_fill<<<296, 256, 0, st>>>(bits);
        GramArgs a{};
        a.tiles = kCheckTiles; a.tile_rows = 3; a.row_lo = 0; a.n_lo = 3; a.acc_mask = 1;
        a.gram = gram; a.ld = 3;
        rc = make_tmap(&a.tmap, bits, 3, kCheckTiles);
        if (rc == PK_OK) rc = launch_one<128, false>(a, device, st, 1);
        if (rc == PK_OK) e = cudaMemcpyAsync(h, gram, sizeof h, cudaMemcpyDeviceToHost, st);
        if (rc == PK_OK && e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    cudaFree(bits); cudaFree(gram); cudaStreamDestroy(st);
    if (rc != PK_OK) return rc;
    if (e != cudaSuccess) return pk_set_error(PK_ERR_CUDA, "gram_f4 exactness check: %s", cudaGetErrorString(e));
    const unsigned long long ones = kCheckOnes * 1024ull, steps = (kCheckTiles - kCheckOnes) * 16ull,
                             third = (steps + 2) / 3, all = kCheckTiles * 1024ull;
    const unsigned long long want[9] = {ones + steps, ones + steps, ones + third,
                                        ones + steps, all,          ones + third,
                                        ones + third, ones + third, ones + third};
    bool ok = true;
    for (int i = 0; i < 9; i++) ok = ok && h[i] == want[i];
    return ok ? 1 : -1;
}

}  // namespace

// 1 = FP32 accumulation of 0/1 products is exact up to 2^24 on this device (checked once, cached),
// 0 = it is not: the callers must use the integer kernels; negative = error
int pk_gram_f4_exact(int device) {
    if (device < 0 || device >= 64) return pk_set_error(PK_ERR_ARG, "pk_gram_f4_exact: device %d", device);
    std::lock_guard<std::mutex> lock(g_check_mutex);
    if (g_exact[device] == 0) {
        if (const char *e = getenv("PYKMER_B200_F4_EXACT")) {      // test hook: pretend the check failed / passed
            g_exact[device] = atoi(e) ? 1 : -1;
        } else {
            pk_device_guard guard(device);
            const int rc = selfcheck(device);
            if (rc != 1 && rc != -1) return rc;
            g_exact[device] = rc;
        }
    }
    return g_exact[device] == 1 ? 1 : 0;
}

// gram (int64, nsamples x nsamples, already zeroed or holding a partial sum) += B * B^T for tiled masks
// of `nsamples` rows and `words` (a multiple of 32) words per row.  Any nsamples: up to 256 in one
// launch, more block pair by block pair -- pairs (2i, 2i+1) of 128-row blocks with all three
// accumulators, then every remaining cross pair with the (lo,hi) accumulator alone.
int pk_gram_f4_launch(const uint32_t *bits_dev, int nsamples, size_t words, int64_t *gram_dev, int device,
                      cudaStream_t st) {
    if (words & 31)
        return pk_set_error(PK_ERR_ARG, "gram_f4: the tiled layout needs a multiple of 32 words, got %zu", words);
    const int exact = pk_gram_f4_exact(device);
    if (exact < 0) return exact;
    if (!exact)
        return pk_set_error(PK_ERR_STATE, "gram_f4: this device does not accumulate FP4 products exactly up to 2^24 "
                            "(pk_gram_f4_exact); use the integer Gram kernels (row-major masks, pk_gram_device)");
    if ((uintptr_t)bits_dev & 15u) return pk_set_error(PK_ERR_ARG, "gram_f4: the masks must be 16-byte aligned");
    GramArgs a{};
    a.tiles = words / 32; a.tile_rows = nsamples;
    a.gram = reinterpret_cast<unsigned long long *>(gram_dev); a.ld = nsamples;
    if (a.tiles == 0) return PK_OK;
    {
        const int rc = make_tmap(&a.tmap, bits_dev, nsamples, a.tiles);
        if (rc != PK_OK) return rc;
    }
    if (nsamples <= 64) {
        a.row_lo = 0; a.n_lo = nsamples; a.acc_mask = 1;
        return a.tiles >= 2 ? launch_one<128, true>(a, device, st, 0) : launch_one<128, false>(a, device, st, 0);
    }
    if (nsamples <= 128) {
        a.row_lo = 0; a.n_lo = nsamples; a.acc_mask = 1;
        return launch_one<128, false>(a, device, st, 0);
    }
    const int nb = (nsamples + 127) / 128;
    auto rows = [&](int b) { return std::min(128, nsamples - b * 128); };
    for (int b = 0; b + 1 < nb; b += 2) {                    // (2i, 2i+1): the two diagonal blocks and their cross block
        a.row_lo = b * 128; a.n_lo = rows(b); a.row_hi = (b + 1) * 128; a.n_hi = rows(b + 1); a.acc_mask = 7;
        const int rc = launch_one<256, false>(a, device, st, 0);
        if (rc != PK_OK) return rc;
    }
    if (nb & 1) {                                            // the last block's own diagonal
        a.row_lo = (nb - 1) * 128; a.n_lo = rows(nb - 1); a.acc_mask = 1;
        const int rc = launch_one<128, false>(a, device, st, 0);
        if (rc != PK_OK) return rc;
    }
    for (int i = 0; i < nb; i++)
        for (int j = i + 1; j < nb; j++) {
            if ((i & 1) == 0 && j == i + 1) continue;        // done above
            a.row_lo = i * 128; a.n_lo = rows(i); a.row_hi = j * 128; a.n_hi = rows(j); a.acc_mask = 2;
            const int rc = launch_one<256, false>(a, device, st, 0);
            if (rc != PK_OK) return rc;
        }
    return PK_OK;
}

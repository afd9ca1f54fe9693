// merger.cu -- the merger hot path on B200 (sm_100a).
//
// Replaces, from the reference (sauloal/pykmer): Header.calculate_distance
// (tools.py:439-493: s_valid/o_valid/c_valid and the three sums) and the pair
// loop of merge (merger.py:136-176), recast as a thresholded presence-bitmask
// Gram matrix so every sample is read once instead of N-1 times:
//   k_threshold_pack   uint8 table -> 1 bit per k-mer  (min <= c <= max), row-major or tiled masks
//   (gram_f4.cu, gram_i8.cu: the Gram matrix on the tensor cores -- the default up to 256 samples)
//   k_gram_popc        G[k][l] += popcount(bits[k] & bits[l]), 4x4 register tiles (any N)
//   k_pair_counts      the literal three sums for one pair of tables
//   k_synth_table      deterministic synthetic tables for the benchmark
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "common.h"

namespace {

// ------------------------------------------------------------------ threshold + pack
// One thread turns 32 table bytes (two 16-byte loads) into one 32-bit word.
__device__ __forceinline__ uint32_t valid_nibble(uint32_t w, uint32_t lo4, uint32_t hi4) {
    // 0xFF in every byte with lo <= byte <= hi, then one bit per byte
    const uint32_t m = __vcmpgeu4(w, lo4) & __vcmpleu4(w, hi4);
    return ((m & 0x08040201u) * 0x01010101u) >> 24;      // byte b -> bit b
}

// nrows = 0: word wi goes to bits[wi].  nrows = R: the tiled layout -- word g = first_word + wi of
// sample `row` goes to bits[(g / 32) * R * 32 + row * 32 + g % 32].
__global__ void __launch_bounds__(256) k_threshold_pack(const uint8_t *__restrict__ table, size_t n,
                                                        uint32_t lo, uint32_t hi,
                                                        uint32_t *__restrict__ bits, size_t first_word,
                                                        uint32_t row, uint32_t nrows) {
    const uint32_t lo4 = lo * 0x01010101u, hi4 = hi * 0x01010101u;
    const size_t words = (n + 31) / 32;
    const size_t full_words = n / 32;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint4 *v = reinterpret_cast<const uint4 *>(table);
    for (size_t wi = (size_t)blockIdx.x * blockDim.x + threadIdx.x; wi < words; wi += stride) {
        uint32_t out = 0;
        if (wi < full_words) {
            const uint4 a = __ldcs(v + 2 * wi), b = __ldcs(v + 2 * wi + 1);
            out  = valid_nibble(a.x, lo4, hi4)       | valid_nibble(a.y, lo4, hi4) << 4;
            out |= valid_nibble(a.z, lo4, hi4) << 8  | valid_nibble(a.w, lo4, hi4) << 12;
            out |= valid_nibble(b.x, lo4, hi4) << 16 | valid_nibble(b.y, lo4, hi4) << 20;
            out |= valid_nibble(b.z, lo4, hi4) << 24 | valid_nibble(b.w, lo4, hi4) << 28;
        } else {
            for (size_t i = wi * 32; i < n; i++) {
                const uint32_t c = table[i];
                if (c >= lo && c <= hi) out |= 1u << (i & 31);
            }
        }
        if (nrows) {
            const size_t g = first_word + wi;
            bits[((g >> 5) * nrows + row) * 32 + (g & 31)] = out;
        } else {
            bits[wi] = out;
        }
    }
}

// ------------------------------------------------------------------ Gram by AND + popcount
// Grid: x = word slabs (persistent, grid-stride over chunks of kWT words),
//       y = pairs (I, J) of 64-sample panels with I <= J.
// A block stages kWT words of both panels in shared memory as [sample][word]
// (row stride kWT + 4 words, 16-byte loads and stores, conflict free).  A thread
// owns a 4x4 tile of sample pairs with STRIDED rows {ti + 16x} x {tj + 16y}: the
// eight lanes of a quarter warp then read eight different 16-byte bank groups,
// every operand fetch is one LDS.128 covering 4 words, and the 16 partial sums
// stay in registers for the whole slab; nothing is reduced across lanes.
constexpr int kPanel = 64;          // samples per panel
constexpr int kWT = 64;             // words staged per step
constexpr int kRow = kWT + 4;       // padded shared-memory row, in words
constexpr int kGramThreads = 256;

__global__ void __launch_bounds__(kGramThreads) k_gram_popc(const uint32_t *__restrict__ bits,
                                                            int nsamples, size_t words,
                                                            size_t stride_words,
                                                            unsigned long long *__restrict__ gram) {
    __shared__ __align__(16) uint32_t sA[kPanel * kRow];
    __shared__ __align__(16) uint32_t sB[kPanel * kRow];

    // decode the panel pair from blockIdx.y (upper triangle, row-major)
    const int npanels = (nsamples + kPanel - 1) / kPanel;
    int I = 0, rem = blockIdx.y;
    while (rem >= npanels - I) { rem -= npanels - I; I++; }
    const int J = I + rem;
    const bool diag = (I == J);
    const int rowsA = min(kPanel, nsamples - I * kPanel);
    const int rowsB = min(kPanel, nsamples - J * kPanel);

    // thread -> tile (ti, tj) in 16 x 16; on a diagonal panel only ti <= tj
    int ti = -1, tj = -1;
    {
        int t = threadIdx.x;
        if (diag) {
            int r = 0;
            while (r < 16 && t >= 16 - r) { t -= 16 - r; r++; }
            if (r < 16) { ti = r; tj = r + t; }
        } else {
            ti = t >> 4; tj = t & 15;
        }
    }
    const bool active = ti >= 0 && ti < rowsA && tj < rowsB;

    uint32_t acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
        for (int y = 0; y < 4; y++) acc[x][y] = 0;

    const bool vec_ok = ((stride_words & 3) == 0) && ((((uintptr_t)bits) & 15u) == 0);
    const size_t nchunks = (words + kWT - 1) / kWT;
    for (size_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const size_t w0 = ch * kWT;
        __syncthreads();
        // stage: 16 lanes x 16 bytes cover one sample's 64 words
        for (int idx = threadIdx.x; idx < kPanel * (kWT / 4); idx += kGramThreads) {
            const int s = idx / (kWT / 4), w = (idx % (kWT / 4)) * 4;
            uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
            const size_t gw = w0 + w;
            if (s < rowsA && gw < words) {
                const uint32_t *src = bits + (size_t)(I * kPanel + s) * stride_words + gw;
                if (vec_ok && gw + 4 <= words) va = __ldg(reinterpret_cast<const uint4 *>(src));
                else {
                    va.x = __ldg(src);
                    if (gw + 1 < words) va.y = __ldg(src + 1);
                    if (gw + 2 < words) va.z = __ldg(src + 2);
                    if (gw + 3 < words) va.w = __ldg(src + 3);
                }
            }
            *reinterpret_cast<uint4 *>(&sA[s * kRow + w]) = va;
            if (!diag) {
                if (s < rowsB && gw < words) {
                    const uint32_t *src = bits + (size_t)(J * kPanel + s) * stride_words + gw;
                    if (vec_ok && gw + 4 <= words) vb = __ldg(reinterpret_cast<const uint4 *>(src));
                    else {
                        vb.x = __ldg(src);
                        if (gw + 1 < words) vb.y = __ldg(src + 1);
                        if (gw + 2 < words) vb.z = __ldg(src + 2);
                        if (gw + 3 < words) vb.w = __ldg(src + 3);
                    }
                }
                *reinterpret_cast<uint4 *>(&sB[s * kRow + w]) = vb;
            }
        }
        __syncthreads();
        if (active) {
            const uint32_t *pB = diag ? sA : sB;
#pragma unroll 2
            for (int w = 0; w < kWT; w += 4) {
                uint4 a[4], b[4];
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    a[x] = *reinterpret_cast<const uint4 *>(&sA[(ti + 16 * x) * kRow + w]);
                    b[x] = *reinterpret_cast<const uint4 *>(&pB[(tj + 16 * x) * kRow + w]);
                }
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++)
                        acc[x][y] += __popc(a[x].x & b[y].x) + __popc(a[x].y & b[y].y) +
                                     __popc(a[x].z & b[y].z) + __popc(a[x].w & b[y].w);
            }
        }
    }
    if (active) {
#pragma unroll
        for (int x = 0; x < 4; x++)
#pragma unroll
            for (int y = 0; y < 4; y++) {
                if (diag && ti == tj && x > y) continue;       // mirrored inside the tile
                const int k = I * kPanel + ti + 16 * x, l = J * kPanel + tj + 16 * y;
                if (ti + 16 * x >= rowsA || tj + 16 * y >= rowsB || !acc[x][y]) continue;
                atomicAdd(&gram[(size_t)k * nsamples + l], (unsigned long long)acc[x][y]);
                if (k != l) atomicAdd(&gram[(size_t)l * nsamples + k], (unsigned long long)acc[x][y]);
            }
    }
}

// ------------------------------------------------------------------ one pair, literal form
__global__ void __launch_bounds__(256) k_pair_counts(const uint8_t *__restrict__ s,
                                                     const uint8_t *__restrict__ o, size_t n,
                                                     uint32_t lo, uint32_t hi,
                                                     unsigned long long *__restrict__ out) {
    const uint32_t lo4 = lo * 0x01010101u, hi4 = hi * 0x01010101u;
    unsigned long long cs = 0, co = 0, cc = 0;
    const size_t nvec = n / 16;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint4 *vs = reinterpret_cast<const uint4 *>(s), *vo = reinterpret_cast<const uint4 *>(o);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 a = __ldcs(vs + i), b = __ldcs(vo + i);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t ma = __vcmpgeu4(aw[k], lo4) & __vcmpleu4(aw[k], hi4);   // tools.py:473
            const uint32_t mb = __vcmpgeu4(bw[k], lo4) & __vcmpleu4(bw[k], hi4);   // tools.py:474
            cs += __popc(ma) >> 3;
            co += __popc(mb) >> 3;
            cc += __popc(ma & mb) >> 3;                                            // tools.py:475
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (size_t i = nvec * 16; i < n; i++) {
            const bool va = s[i] >= lo && s[i] <= hi, vb = o[i] >= lo && o[i] <= hi;
            cs += va; co += vb; cc += (va && vb);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        cs += __shfl_xor_sync(0xFFFFFFFFu, cs, d);
        co += __shfl_xor_sync(0xFFFFFFFFu, co, d);
        cc += __shfl_xor_sync(0xFFFFFFFFu, cc, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cs) atomicAdd(out + 0, cs);
        if (co) atomicAdd(out + 1, co);
        if (cc) atomicAdd(out + 2, cc);
    }
}

// ------------------------------------------------------------------ synthetic tables
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) k_synth_table(uint8_t *__restrict__ dst, int sample,
                                                     uint64_t lo, uint64_t hi) {
    const uint64_t k1 = 0x9E3779B97F4A7C15ull * (uint64_t)(sample % 5 + 1);
    const uint64_t k2 = 0xD1B54A32D192ED03ull * (uint64_t)(sample + 1);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
        const uint64_t h1 = splitmix64(i ^ k1), h2 = splitmix64(i ^ k2);
        const bool present = ((h1 & 0xFFFFull) < 5243ull) || ((h2 & 0xFFFFull) < 5898ull);
        const uint32_t t = (uint32_t)((h2 >> 16) & 0xFFFFFFull);
        const uint32_t nt = (~t) & 0xFFFFFFu;
        const uint32_t ones = nt ? (uint32_t)(__ffs((int)nt) - 1) : 24u;   // trailing ones of t
        uint32_t val = min(255u, 1u + ones);
        if (((h2 >> 40) & 0xFFull) < 8ull) val = 51u + (uint32_t)((h2 >> 32) & 0x7Full);
        if ((h2 >> 48) < 43ull) val = 255u;
        dst[i - lo] = present ? (uint8_t)val : (uint8_t)0;
    }
}

int grid_for(size_t items, int threads, int device, int per_sm) {
    const size_t want = (items + threads - 1) / threads;
    const size_t cap = (size_t)pk_sm_count(device) * per_sm;
    return (int)std::max<size_t>(1, std::min(want, cap));
}

}  // namespace

PK_API int pk_threshold_pack_device(const uint8_t *table_dev, size_t n, int min_count, int max_count,
                                    uint32_t *bits_dev, pk_stream stream) {
    PK_REQUIRE(table_dev != nullptr && bits_dev != nullptr, "pk_threshold_pack_device: NULL pointer");
    // merger.py:90-91: min_count >= 1, max_count <= 255
    PK_REQUIRE(min_count >= 1 && max_count <= 255, "pk_threshold_pack_device: thresholds [%d, %d] outside 1..255",
               min_count, max_count);
    PK_REQUIRE(((uintptr_t)table_dev & 15u) == 0, "pk_threshold_pack_device: table_dev must be 16-byte aligned");
    if (n == 0) return PK_OK;
    int device = 0;
    PK_CUDA(cudaGetDevice(&device));
    const size_t words = (n + 31) / 32;
    k_threshold_pack<<<grid_for(words, 256, device, 8), 256, 0, (cudaStream_t)stream>>>(
        table_dev, n, (uint32_t)min_count, (uint32_t)max_count, bits_dev, 0, 0u, 0u);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

PK_API int pk_threshold_pack_tiled_device(const uint8_t *table_dev, size_t n, size_t first_word, int min_count,
                                          int max_count, uint32_t *bits_tiled_dev, int row, int nrows,
                                          pk_stream stream) {
    PK_REQUIRE(table_dev != nullptr && bits_tiled_dev != nullptr, "pk_threshold_pack_tiled_device: NULL pointer");
    PK_REQUIRE(min_count >= 1 && max_count <= 255, "pk_threshold_pack_tiled_device: thresholds [%d, %d] outside 1..255",
               min_count, max_count);                                           // merger.py:90-91
    PK_REQUIRE(((uintptr_t)table_dev & 15u) == 0, "pk_threshold_pack_tiled_device: table_dev must be 16-byte aligned");
    PK_REQUIRE(nrows >= 1 && row >= 0 && row < nrows, "pk_threshold_pack_tiled_device: row %d of %d", row, nrows);
    if (n == 0) return PK_OK;
    int device = 0;
    PK_CUDA(cudaGetDevice(&device));
    const size_t words = (n + 31) / 32;
    k_threshold_pack<<<grid_for(words, 256, device, 8), 256, 0, (cudaStream_t)stream>>>(
        table_dev, n, (uint32_t)min_count, (uint32_t)max_count, bits_tiled_dev, first_word, (uint32_t)row,
        (uint32_t)nrows);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

PK_API int pk_gram_tiled_exact(int device, int *exact) {
    PK_REQUIRE(exact != nullptr, "pk_gram_tiled_exact: NULL output");
    int ndev = 0;
    PK_CUDA(cudaGetDeviceCount(&ndev));
    PK_REQUIRE(device >= 0 && device < ndev, "pk_gram_tiled_exact: device %d of %d", device, ndev);
    const int rc = pk_gram_f4_exact(device);
    if (rc < 0) return rc;
    *exact = rc;
    return PK_OK;
}

PK_API int pk_gram_tiled_device(const uint32_t *bits_tiled_dev, int nsamples, size_t words, int64_t *gram_dev,
                                int accumulate, pk_stream stream) {
    PK_REQUIRE(bits_tiled_dev != nullptr && gram_dev != nullptr, "pk_gram_tiled_device: NULL pointer");
    PK_REQUIRE(nsamples >= 1 && nsamples <= PK_TILED_MAX_SAMPLES,
               "pk_gram_tiled_device: %d samples outside 1..%d", nsamples, PK_TILED_MAX_SAMPLES);
    PK_REQUIRE(((uintptr_t)bits_tiled_dev & 15u) == 0, "pk_gram_tiled_device: bits must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0;
    PK_CUDA(cudaGetDevice(&device));
    if (!accumulate)
        PK_CUDA(cudaMemsetAsync(gram_dev, 0, (size_t)nsamples * nsamples * sizeof(int64_t), st));
    if (words == 0) return PK_OK;
    const size_t padded = (words + 31) & ~(size_t)31;             // whole tiles; the padding words are zero
    return pk_gram_f4_launch(bits_tiled_dev, nsamples, padded, gram_dev, device, st);
}

PK_API int pk_gram_device(const uint32_t *bits_dev, int nsamples, size_t words, size_t stride_words,
                          int64_t *gram_dev, int accumulate, pk_stream stream) {
    PK_REQUIRE(bits_dev != nullptr && gram_dev != nullptr, "pk_gram_device: NULL pointer");
    PK_REQUIRE(nsamples >= 1 && nsamples <= 4096, "pk_gram_device: nsamples %d outside 1..4096", nsamples);
    PK_REQUIRE(stride_words >= words, "pk_gram_device: stride %zu < words %zu", stride_words, words);
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0;
    PK_CUDA(cudaGetDevice(&device));
    if (!accumulate)
        PK_CUDA(cudaMemsetAsync(gram_dev, 0, (size_t)nsamples * nsamples * sizeof(int64_t), st));
    if (words == 0) return PK_OK;
    // Row-major masks -- the two integer implementations the FP4 kernel on tiled masks
    // (pk_gram_tiled_device, the merger's default) is checked against, and its stand-ins on a device
    // that fails pk_gram_f4_exact: tcgen05 kind::i8 (N <= 256; integer accumulators) and AND +
    // popcount on the ALUs (any N).  PYKMER_B200_GRAM=popc forces the latter (test hook).
    const char *algo = getenv("PYKMER_B200_GRAM");
    const bool want_popc = algo && strcmp(algo, "popc") == 0;
    if (!want_popc && nsamples <= 256)
        return pk_gram_i8_launch(bits_dev, nsamples, words, stride_words, gram_dev, device, st);
    const int npanels = (nsamples + kPanel - 1) / kPanel;
    const int npairs = npanels * (npanels + 1) / 2;
    const size_t nchunks = (words + kWT - 1) / kWT;
    // 32-bit partial sums: a block sees at most 2^26 words (2^31 bits)
    const size_t min_x = (words + ((1ull << 26) - 1)) >> 26;
    size_t gx = std::max<size_t>(1, (size_t)pk_sm_count(device) * 4 / (size_t)npairs);
    gx = std::max(gx, min_x);
    gx = std::min(gx, nchunks);
    dim3 grid((unsigned)gx, (unsigned)npairs);
    k_gram_popc<<<grid, kGramThreads, 0, st>>>(bits_dev, nsamples, words, stride_words,
                                               reinterpret_cast<unsigned long long *>(gram_dev));
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

PK_API int pk_pair_counts_device(const uint8_t *s_dev, const uint8_t *o_dev, size_t n, int min_count,
                                 int max_count, uint64_t out_host[3], pk_stream stream) {
    PK_REQUIRE(s_dev != nullptr && o_dev != nullptr && out_host != nullptr, "pk_pair_counts_device: NULL pointer");
    PK_REQUIRE(min_count >= 1 && max_count <= 255, "pk_pair_counts_device: thresholds [%d, %d] outside 1..255",
               min_count, max_count);
    PK_REQUIRE((((uintptr_t)s_dev | (uintptr_t)o_dev) & 15u) == 0, "pk_pair_counts_device: tables must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0;
    PK_CUDA(cudaGetDevice(&device));
    unsigned long long *d = nullptr, h[3] = {0, 0, 0};
    PK_CUDA(cudaMalloc(&d, 3 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), st);
    if (e == cudaSuccess && n) {
        k_pair_counts<<<grid_for(n / 16 + 1, 256, device, 8), 256, 0, st>>>(
            s_dev, o_dev, n, (uint32_t)min_count, (uint32_t)max_count, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return pk_set_error(PK_ERR_CUDA, "pk_pair_counts_device: %s", cudaGetErrorString(e));
    out_host[0] = h[0]; out_host[1] = h[1]; out_host[2] = h[2];
    return PK_OK;
}

PK_API int pk_merge_host(const uint8_t *const *tables_host, int nsamples, size_t n, int min_count,
                         int max_count, int device, uint64_t *matrix_host) {
    PK_REQUIRE(tables_host != nullptr && matrix_host != nullptr, "pk_merge_host: NULL pointer");
    PK_REQUIRE(nsamples >= 1, "pk_merge_host: no samples");                      // merger.py:94
    PK_REQUIRE(min_count >= 1 && max_count <= 255, "pk_merge_host: thresholds [%d, %d] outside 1..255",
               min_count, max_count);                                           // merger.py:90-91
    for (int s = 0; s < nsamples; s++)
        PK_REQUIRE(tables_host[s] != nullptr, "pk_merge_host: table %d is NULL", s);
    int ndev = 0;
    PK_CUDA(cudaGetDeviceCount(&ndev));
    PK_REQUIRE(device >= 0 && device < ndev, "pk_merge_host: device %d of %d", device, ndev);
    pk_device_guard guard(device);
    if (!guard.ok) return pk_set_error(PK_ERR_CUDA, "cudaSetDevice(%d) failed", device);

    // tiled masks + the FP4 Gram kernel (gram_f4.cu) unless the device fails its exactness check;
    // PYKMER_B200_GRAM=i8|popc (test hook) keeps row-major masks and the integer kernel named
    const int exact = pk_gram_f4_exact(device);
    if (exact < 0) return exact;
    const bool tiled = exact == 1 && nsamples <= PK_TILED_MAX_SAMPLES && getenv("PYKMER_B200_GRAM") == nullptr;
    // The masks of ALL samples for one stretch of the k-mer axis must be resident for the contraction
    // (N / 8 bytes per k-mer).  When the whole axis does not fit (K=17 x 255 samples: 548 GB) it is cut
    // into chunks of whole 1024-k-mer tiles that do, and the chunks' Gram matrices are added up -- the
    // sums of tools.py:480-482 are sums over that axis, like the reference's own 100 M-entry blocks
    // (tools.py:449-489).  PYKMER_B200_MASK_BUDGET (bytes; test hook) forces small chunks.
    size_t free_b = 0, total_b = 0;
    PK_CUDA(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = free_b / 2;
    if (const char *env = getenv("PYKMER_B200_MASK_BUDGET")) budget = (size_t)strtoull(env, nullptr, 10);
    size_t per_chunk = std::max<size_t>(1024, (budget * 8 / (size_t)nsamples) / 1024 * 1024);
    per_chunk = std::min(per_chunk, std::max<size_t>(n, 1));
    const size_t words = (per_chunk + 31) / 32;                      // mask words per sample and chunk
    const size_t stride_words = (words + 3) & ~(size_t)3;
    const size_t mask_words = tiled ? ((words + 31) / 32) * 32 * (size_t)nsamples : (size_t)nsamples * stride_words;
    const size_t chunk = std::min<size_t>(per_chunk, 64u << 20);     // bytes per staged copy, multiple of 32
    uint8_t *stage[2] = {nullptr, nullptr};
    uint32_t *bits = nullptr;
    int64_t *gram = nullptr;
    cudaStream_t copy_st = nullptr, work_st = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    std::vector<int64_t> G((size_t)nsamples * nsamples);
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    step(cudaMalloc(&bits, mask_words * sizeof(uint32_t)));
    step(cudaMalloc(&gram, G.size() * sizeof(int64_t)));
    step(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
    step(cudaStreamCreateWithFlags(&work_st, cudaStreamNonBlocking));
    if (e == cudaSuccess) step(cudaMemsetAsync(gram, 0, G.size() * sizeof(int64_t), work_st));
    for (int i = 0; i < 2; i++) {
        step(cudaMalloc(&stage[i], chunk ? chunk : 32));
        step(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
        step(cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming));
        if (e == cudaSuccess) step(cudaEventRecord(consumed[i], work_st));
    }
    int rc = PK_OK, buf = 0;
    for (size_t c0 = 0; c0 < n && e == cudaSuccess && rc == PK_OK; c0 += per_chunk) {
        const size_t cn = std::min(per_chunk, n - c0);               // k-mers of this chunk
        if (tiled) step(cudaMemsetAsync(bits, 0, mask_words * sizeof(uint32_t), work_st));   // tile padding
        for (int s = 0; s < nsamples && e == cudaSuccess && rc == PK_OK; s++) {
            for (size_t off = 0; off < cn && e == cudaSuccess && rc == PK_OK; off += chunk) {
                const size_t len = std::min(chunk, cn - off);
                step(cudaStreamWaitEvent(copy_st, consumed[buf], 0));
                step(cudaMemcpyAsync(stage[buf], tables_host[s] + c0 + off, len, cudaMemcpyHostToDevice, copy_st));
                step(cudaEventRecord(copied[buf], copy_st));
                step(cudaStreamWaitEvent(work_st, copied[buf], 0));
                if (e == cudaSuccess)
                    rc = tiled ? pk_threshold_pack_tiled_device(stage[buf], len, off / 32, min_count, max_count, bits, s,
                                                                nsamples, work_st)
                               : pk_threshold_pack_device(stage[buf], len, min_count, max_count,
                                                          bits + (size_t)s * stride_words + off / 32, work_st);
                step(cudaEventRecord(consumed[buf], work_st));
                buf ^= 1;
            }
        }
        const size_t cw = (cn + 31) / 32;
        if (e == cudaSuccess && rc == PK_OK)
            rc = tiled ? pk_gram_tiled_device(bits, nsamples, cw, gram, 1, work_st)
                       : pk_gram_device(bits, nsamples, cw, stride_words, gram, 1, work_st);
    }
    if (e == cudaSuccess && rc == PK_OK)
        step(cudaMemcpyAsync(G.data(), gram, G.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, work_st));
    if (e == cudaSuccess && rc == PK_OK) step(cudaStreamSynchronize(work_st));
    if (copy_st) cudaStreamSynchronize(copy_st);
    if (work_st) cudaStreamSynchronize(work_st);
    for (int i = 0; i < 2; i++) {
        cudaFree(stage[i]);
        if (copied[i]) cudaEventDestroy(copied[i]);
        if (consumed[i]) cudaEventDestroy(consumed[i]);
    }
    cudaFree(bits); cudaFree(gram);
    if (copy_st) cudaStreamDestroy(copy_st);
    if (work_st) cudaStreamDestroy(work_st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return pk_set_error(e == cudaErrorMemoryAllocation ? PK_ERR_NOMEM : PK_ERR_CUDA,
                            "pk_merge_host: %s", cudaGetErrorString(e));
    }
    if (rc != PK_OK) return rc;
    for (int k = 0; k < nsamples; k++)
        for (int l = 0; l < nsamples; l++) {
            uint64_t *c = matrix_host + ((size_t)k * nsamples + l) * 3;
            c[0] = (uint64_t)G[(size_t)k * nsamples + k];                       // merger.py:175-176
            c[1] = (uint64_t)G[(size_t)l * nsamples + l];
            c[2] = (uint64_t)G[(size_t)k * nsamples + l];
        }
    return PK_OK;
}

PK_API int pk_synth_table_device(uint8_t *dst_dev, int sample, uint64_t lo, uint64_t hi,
                                 pk_stream stream) {
    PK_REQUIRE(dst_dev != nullptr, "pk_synth_table_device: NULL pointer");
    PK_REQUIRE(sample >= 0 && lo <= hi, "pk_synth_table_device: bad arguments");
    if (lo == hi) return PK_OK;
    int device = 0;
    PK_CUDA(cudaGetDevice(&device));
    k_synth_table<<<grid_for((size_t)(hi - lo), 256, device, 16), 256, 0, (cudaStream_t)stream>>>(
        dst_dev, sample, lo, hi);
    PK_CUDA(cudaGetLastError());
    return PK_OK;
}

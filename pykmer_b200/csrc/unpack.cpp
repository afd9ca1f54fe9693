// unpack.cpp -- host side of the packed table transfer (pk_indexer_finalize_to_host).
//
// A count table (the .kin bytes, tools.py:196,240-243) is mostly zeros: the tomato-sized genome fills 25 %
// of the 4^15 entries, 3 % of 4^17.  Crossing PCIe it is the largest transfer of an indexing job (1 GiB out
// against 0.78 GB in at K=15), so the device sends a *packed* slice instead -- k_table_pack, indexer.cu:
//     bitmap     one bit per entry (bit i of 64-bit word j <=> entry 64 j + i is non-zero)
//     chunk_off  per chunk of 1024 entries: where its non-zero bytes start in `nz`, in 16-byte units
//     nz         the non-zero bytes of every chunk in entry order, each chunk padded to 16 bytes
// and the host rebuilds the bytes here on all cores, straight into the caller's buffer.  With AVX-512 VBMI2
// a 64-entry group is one expand-load + one streaming store; otherwise BMI2 `pdep` does 8 entries per step;
// otherwise a plain loop.  Host code only: no CUDA call, usable (and tested) without a GPU.
#include <atomic>
#include <immintrin.h>
#include <string.h>
#include <thread>
#include <vector>

#include "common.h"

namespace {

constexpr size_t kChunk = 1024;                            // entries per chunk (PK_PACK_CHUNK)

void unpack_chunks_scalar(const uint64_t *bitmap, const uint32_t *chunk_off, const uint8_t *nz, uint8_t *dst,
                          size_t c0, size_t c1) {
    for (size_t c = c0; c < c1; c++) {
        const uint8_t *src = nz + (size_t)chunk_off[c] * 16;
        uint8_t *out = dst + c * kChunk;
        memset(out, 0, kChunk);
        for (size_t j = 0; j < kChunk / 64; j++) {
            uint64_t m = bitmap[c * (kChunk / 64) + j];
            while (m) {
                out[j * 64 + (size_t)__builtin_ctzll(m)] = *src++;
                m &= m - 1;
            }
        }
    }
}

__attribute__((target("bmi2,popcnt")))
void unpack_chunks_bmi2(const uint64_t *bitmap, const uint32_t *chunk_off, const uint8_t *nz, uint8_t *dst,
                        size_t c0, size_t c1, const uint8_t *nz_end) {
    // spread[b]: 0xFF in byte i for every set bit i of b -- the pdep mask that drops the next popcount(b)
    // source bytes into the right lanes of 8 output bytes
    static uint64_t spread[256];
    static std::atomic<int> ready(0);
    if (!ready.load(std::memory_order_acquire)) {
        for (int b = 0; b < 256; b++) {
            uint64_t v = 0;
            for (int i = 0; i < 8; i++)
                if (b >> i & 1) v |= 0xFFull << (8 * i);
            spread[b] = v;
        }
        ready.store(1, std::memory_order_release);
    }
    for (size_t c = c0; c < c1; c++) {
        const uint8_t *src = nz + (size_t)chunk_off[c] * 16;
        uint64_t *out = reinterpret_cast<uint64_t *>(dst + c * kChunk);
        const uint8_t *bm = reinterpret_cast<const uint8_t *>(bitmap + c * (kChunk / 64));
        for (size_t j = 0; j < kChunk / 8; j++) {
            const unsigned b = bm[j];
            uint64_t w = 0;
            if (b) {
                uint64_t s;
                if (src + 8 <= nz_end) memcpy(&s, src, 8);
                else { s = 0; memcpy(&s, src, (size_t)(nz_end - src)); }
                w = _pdep_u64(s, spread[b]);
                src += __builtin_popcount(b);
            }
            memcpy(out + j, &w, 8);
        }
    }
}

__attribute__((target("avx512f,avx512bw,avx512vbmi2,popcnt")))
void unpack_chunks_vbmi2(const uint64_t *bitmap, const uint32_t *chunk_off, const uint8_t *nz, uint8_t *dst,
                         size_t c0, size_t c1, const uint8_t *nz_end) {
    const bool aligned = ((uintptr_t)dst & 63) == 0;
    for (size_t c = c0; c < c1; c++) {
        const uint8_t *src = nz + (size_t)chunk_off[c] * 16;
        uint8_t *out = dst + c * kChunk;
        const uint64_t *bm = bitmap + c * (kChunk / 64);
        // the register form of the expand (plain 64-byte load first) runs several times faster than the
        // memory form, but reads up to 63 bytes past the chunk's last byte: only where those exist
        if (aligned && src + kChunk + 64 <= nz_end) {
#pragma GCC unroll 4
            for (size_t j = 0; j < kChunk / 64; j++) {
                const __mmask64 m = bm[j];
                const __m512i v = _mm512_maskz_expand_epi8(m, _mm512_loadu_si512(src));
                _mm512_stream_si512(reinterpret_cast<__m512i *>(out + j * 64), v);   // the buffer is written once
                src += __builtin_popcountll(m);
            }
        } else {
            for (size_t j = 0; j < kChunk / 64; j++) {
                const __mmask64 m = bm[j];
                _mm512_storeu_si512(out + j * 64, _mm512_maskz_expandloadu_epi8(m, src));
                src += __builtin_popcountll(m);
            }
        }
    }
    _mm_sfence();
}

int unpack_isa() {
    static const int isa = [] {
        __builtin_cpu_init();
        if (const char *v = getenv("PYKMER_B200_UNPACK")) {          // test hook: force a code path
            if (!strcmp(v, "scalar")) return 0;
            if (!strcmp(v, "bmi2") && __builtin_cpu_supports("bmi2")) return 1;
        }
        if (__builtin_cpu_supports("avx512vbmi2") && __builtin_cpu_supports("avx512bw")) return 2;
        if (__builtin_cpu_supports("bmi2")) return 1;
        return 0;
    }();
    return isa;
}

}  // namespace

// chunks [c0, c1) of a packed slice -> dst[c0 * 1024, c1 * 1024); nz[0, nz_readable) may be read (never
// past it), the caller vouches for the offsets.  Shared with indexer.cu's transfer pipeline.
void pk_unpack_chunks(const uint64_t *bitmap, const uint32_t *chunk_off, const uint8_t *nz, size_t nz_readable,
                      uint8_t *dst, size_t c0, size_t c1) {
    switch (unpack_isa()) {
    case 2: unpack_chunks_vbmi2(bitmap, chunk_off, nz, dst, c0, c1, nz + nz_readable); break;
    case 1: unpack_chunks_bmi2(bitmap, chunk_off, nz, dst, c0, c1, nz + nz_readable); break;
    default: unpack_chunks_scalar(bitmap, chunk_off, nz, dst, c0, c1);
    }
}

PK_API int pk_table_unpack(const uint64_t *bitmap, const uint32_t *chunk_off, const uint8_t *nz, size_t nz_bytes,
                           size_t n, uint8_t *dst, int threads) {
    PK_REQUIRE(n % kChunk == 0, "pk_table_unpack: %zu entries are not a multiple of the %zu-entry chunk", n, kChunk);
    if (n == 0) return PK_OK;
    PK_REQUIRE(bitmap != nullptr && chunk_off != nullptr && dst != nullptr && (nz != nullptr || nz_bytes == 0),
               "pk_table_unpack: NULL argument");
    const size_t nchunks = n / kChunk;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, 64));
    const size_t grain = 256;                              // chunks per grab: 256 KiB of output
    if ((size_t)nt > (nchunks + grain - 1) / grain) nt = (int)((nchunks + grain - 1) / grain);
    std::atomic<size_t> next(0), bad(SIZE_MAX);
    auto work = [&]() {
        for (;;) {
            const size_t c0 = next.fetch_add(grain);
            if (c0 >= nchunks || bad.load() != SIZE_MAX) return;
            const size_t c1 = std::min(nchunks, c0 + grain);
            // every chunk must lie inside nz: offsets and populations are checked before its bytes are written
            for (size_t c = c0; c < c1; c++) {
                size_t pop = 0;
                for (size_t j = 0; j < kChunk / 64; j++) pop += (size_t)__builtin_popcountll(bitmap[c * (kChunk / 64) + j]);
                if ((size_t)chunk_off[c] * 16 + pop > nz_bytes) { bad.store(c); return; }
            }
            pk_unpack_chunks(bitmap, chunk_off, nz, nz_bytes, dst, c0, c1);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (bad.load() != SIZE_MAX)
        return pk_set_error(PK_ERR_ARG, "pk_table_unpack: chunk %zu runs past the %zu packed bytes (offset %zu x 16)",
                            bad.load(), nz_bytes, (size_t)chunk_off[bad.load()]);
    return PK_OK;
}

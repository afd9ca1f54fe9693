// ingest.cpp -- host side of the indexer's ingest: the two passes over
// the FASTA text that dominate the CLI's wall time once the kernels run in milliseconds.
//
// Replaces, from the reference: gzip.open(path, 'rt') of a .gz / .bgz input
// (read_fasta, indexer.py:108-115) -- for BGZF, whose members are independent and carry their
// own sizes, so they inflate in parallel -- and the per-line strip + concatenation of
// parse_fasta (indexer.py:55-95) for the common case of sequence lines without inner white
// space.  The record logic (headers, names, text before the first header, universal newlines
// with stripping) stays in pykmer_b200/fasta.py, which calls these two functions on whole
// blocks of text and falls back to its own line-by-line path whenever pk_fasta_clean reports
// anything unusual.  Host code only: no CUDA call, usable without a GPU.
#include <atomic>
#include <immintrin.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>
#include <zlib.h>

#include "common.h"

namespace {

struct Block {
    size_t in_off, in_len;      // whole gzip member inside the compressed buffer
    size_t out_off, out_len;    // where its ISIZE bytes go
};

// total size of the BGZF member at p (needs the 'BC' extra subfield), 0 = incomplete header,
// -1 = not BGZF
long long bgzf_member_size(const uint8_t *p, size_t avail) {
    if (avail < 18) return 0;
    if (p[0] != 0x1F || p[1] != 0x8B || p[2] != 8 || !(p[3] & 4)) return -1;
    const size_t xlen = (size_t)p[10] | ((size_t)p[11] << 8);
    if (12 + xlen > avail) return 0;
    size_t q = 12;
    const size_t end = 12 + xlen;
    while (q + 4 <= end) {
        const size_t slen = (size_t)p[q + 2] | ((size_t)p[q + 3] << 8);
        if (p[q] == 'B' && p[q + 1] == 'C' && slen == 2) return (long long)((size_t)p[q + 4] | ((size_t)p[q + 5] << 8)) + 1;
        q += 4 + slen;
    }
    return -1;
}

int inflate_member(const uint8_t *src, size_t len, uint8_t *dst, size_t out_len) {
    const size_t xlen = (size_t)src[10] | ((size_t)src[11] << 8);
    const size_t hdr = 12 + xlen;
    if (len < hdr + 8) return -1;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return -2;
    zs.next_in = const_cast<Bytef *>(src + hdr);
    zs.avail_in = (uInt)(len - hdr - 8);
    zs.next_out = dst;
    zs.avail_out = (uInt)out_len;
    const int rc = inflate(&zs, Z_FINISH);
    const size_t got = out_len - zs.avail_out;
    inflateEnd(&zs);
    if (rc != Z_STREAM_END || got != out_len) return -3;
    uint32_t crc;
    memcpy(&crc, src + len - 8, 4);                       // little endian hosts only (x86-64, aarch64)
    if ((uint32_t)crc32(0L, dst, (uInt)out_len) != crc) return -4;
    return 0;
}

int thread_count(int asked, size_t work_items) {
    int t = asked > 0 ? asked : (int)std::thread::hardware_concurrency();
    if (t < 1) t = 1;
    if (t > 64) t = 64;
    if ((size_t)t > work_items) t = (int)(work_items ? work_items : 1);
    return t;
}

}  // namespace

// Inflate as many WHOLE BGZF members of comp[0, comp_len) as fit into out[0, out_cap), in
// parallel.  *consumed = compressed bytes used up (a prefix of whole members), *produced =
// bytes written.  A truncated last member is simply left unconsumed.
PK_API int pk_bgzf_inflate(const uint8_t *comp, size_t comp_len, uint8_t *out, size_t out_cap,
                           size_t *consumed, size_t *produced, int threads) {
    PK_REQUIRE(consumed != nullptr && produced != nullptr, "pk_bgzf_inflate: NULL output");
    PK_REQUIRE(comp_len == 0 || comp != nullptr, "pk_bgzf_inflate: NULL input");
    PK_REQUIRE(out_cap == 0 || out != nullptr, "pk_bgzf_inflate: NULL destination");
    *consumed = 0;
    *produced = 0;
    std::vector<Block> blocks;
    size_t pos = 0, opos = 0;
    while (pos < comp_len) {
        const long long size = bgzf_member_size(comp + pos, comp_len - pos);
        if (size < 0) return pk_set_error(PK_ERR_ARG, "pk_bgzf_inflate: not a BGZF block at offset %zu", pos);
        if (size == 0 || pos + (size_t)size > comp_len) break;          // incomplete: next call
        uint32_t isize;
        memcpy(&isize, comp + pos + (size_t)size - 4, 4);
        if (opos + isize > out_cap) break;                                // destination full
        blocks.push_back({pos, (size_t)size, opos, (size_t)isize});
        pos += (size_t)size;
        opos += isize;
    }
    if (blocks.empty()) return PK_OK;
    const int nt = thread_count(threads, blocks.size());
    std::atomic<size_t> next(0);
    std::atomic<long long> bad(-1);
    std::atomic<int> bad_rc(0);
    auto work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= blocks.size() || bad.load() >= 0) return;
            const Block &b = blocks[i];
            const int rc = inflate_member(comp + b.in_off, b.in_len, out + b.out_off, b.out_len);
            if (rc != 0) { bad_rc.store(rc); bad.store((long long)b.in_off); return; }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (bad.load() >= 0)
        return pk_set_error(PK_ERR_ARG, "pk_bgzf_inflate: BGZF block at offset %lld fails its %s check",
                            bad.load(), bad_rc.load() == -4 ? "CRC" : "length / deflate");
    *consumed = pos;
    *produced = opos;
    return PK_OK;
}

namespace {

// AVX-512 (BW + VBMI2) forms of the two passes of pk_fasta_clean over src[a, b): 64 bytes per step.
// count: bytes kept (everything but \n and \r) and the class flags (2 = other strip()-able white space,
// 4 = non-ASCII); a block holding any byte below 0x21 other than the two terminators is classified byte
// by byte (rare in sequence text).  compact: VPCOMPRESSB in its register form + one store; the store
// is masked where 64 bytes would run into the next thread's part of dst.
__attribute__((target("avx512f,avx512bw,avx512vbmi2,popcnt")))
void clean_count_avx512(const uint8_t *src, size_t a, size_t b, const uint8_t *cls, size_t *kept_out, uint32_t *flags_out) {
    const __m512i nl = _mm512_set1_epi8('\n'), cr = _mm512_set1_epi8('\r'), low = _mm512_set1_epi8(0x21);
    size_t k = 0, i = a;
    uint32_t f = 0;
    for (; i + 64 <= b; i += 64) {
        const __m512i v = _mm512_loadu_si512(src + i);
        const __mmask64 term = _mm512_cmpeq_epi8_mask(v, nl) | _mm512_cmpeq_epi8_mask(v, cr);
        k += 64 - (size_t)__builtin_popcountll(term);
        if (_mm512_movepi8_mask(v)) f |= 4u;
        if (_mm512_cmplt_epu8_mask(v, low) & ~term)
            for (size_t j = i; j < i + 64; j++) f |= cls[src[j]];
    }
    for (; i < b; i++) {
        const uint8_t c = cls[src[i]];
        k += c != 1;
        f |= c;
    }
    *kept_out = k;
    *flags_out = f & 6u;
}

__attribute__((target("avx512f,avx512bw,avx512vbmi2,popcnt")))
void clean_compact_avx512(const uint8_t *src, size_t a, size_t b, uint8_t *o, uint8_t *o_end) {
    const __m512i nl = _mm512_set1_epi8('\n'), cr = _mm512_set1_epi8('\r');
    size_t i = a;
    for (; i + 64 <= b; i += 64) {
        const __m512i v = _mm512_loadu_si512(src + i);
        const __mmask64 keep = ~(_mm512_cmpeq_epi8_mask(v, nl) | _mm512_cmpeq_epi8_mask(v, cr));
        const __m512i packed = _mm512_maskz_compress_epi8(keep, v);
        const size_t cnt = (size_t)__builtin_popcountll(keep);
        if (o + 64 <= o_end) _mm512_storeu_si512(o, packed);
        else _mm512_mask_storeu_epi8(o, cnt >= 64 ? ~0ull : ((1ull << cnt) - 1), packed);
        o += cnt;
    }
    for (; i < b; i++)
        if (src[i] != '\n' && src[i] != '\r') *o++ = src[i];
}

bool clean_has_avx512() {
    static const bool ok = [] {
        __builtin_cpu_init();
        if (const char *v = getenv("PYKMER_B200_CLEAN")) if (!strcmp(v, "scalar")) return false;   // test hook
        return __builtin_cpu_supports("avx512vbmi2") && __builtin_cpu_supports("avx512bw");
    }();
    return ok;
}

}  // namespace

// Sequence lines without inner white space: dst = src minus '\n' and '\r' (every line is then
// already stripped, so the record's bases are just the remaining bytes, indexer.py:56-58,84).
// flags bit 0 = src holds a strip()-able byte other than the two line terminators
// (blank, tab, \v, \f, 0x1c..0x1f), bit 1 = src holds a byte >= 0x80.  With either bit set the
// caller must not use dst (fasta.py then goes line by line, or rejects the input).
PK_API int pk_fasta_clean(const uint8_t *src, size_t n, uint8_t *dst, size_t *n_out, uint32_t *flags,
                          int threads) {
    PK_REQUIRE(n_out != nullptr && flags != nullptr, "pk_fasta_clean: NULL output");
    PK_REQUIRE(n == 0 || (src != nullptr && dst != nullptr), "pk_fasta_clean: NULL buffer");
    *n_out = 0;
    *flags = 0;
    if (n == 0) return PK_OK;
    const size_t grain = (size_t)1 << 20;
    const int nt = thread_count(threads, (n + grain - 1) / grain);
    std::vector<size_t> kept((size_t)nt + 1, 0);
    std::vector<uint32_t> fl((size_t)nt, 0);
    auto range = [&](int t, size_t &a, size_t &b) {
        a = n / (size_t)nt * (size_t)t;
        b = t == nt - 1 ? n : n / (size_t)nt * (size_t)(t + 1);
    };
    // byte classes: 1 = line terminator, 2 = other strip()-able white space, 4 = non-ASCII
    static uint8_t cls[256];
    static std::atomic<int> cls_ready(0);
    if (!cls_ready.load()) {
        uint8_t tmp[256];
        memset(tmp, 0, sizeof tmp);
        tmp['\n'] = tmp['\r'] = 1;
        tmp[' '] = tmp['\t'] = tmp[0x0B] = tmp[0x0C] = tmp[0x1C] = tmp[0x1D] = tmp[0x1E] = tmp[0x1F] = 2;
        for (int c = 128; c < 256; c++) tmp[c] = 4;
        memcpy(cls, tmp, sizeof tmp);
        cls_ready.store(1);
    }
    const bool simd = clean_has_avx512();
    auto count = [&](int t) {
        size_t a, b;
        range(t, a, b);
        if (simd) {
            clean_count_avx512(src, a, b, cls, &kept[(size_t)t + 1], &fl[(size_t)t]);
            return;
        }
        size_t k = 0;
        uint32_t f = 0;
        for (size_t i = a; i < b; i++) {
            const uint8_t c = cls[src[i]];
            k += c != 1;
            f |= c;
        }
        kept[(size_t)t + 1] = k;
        fl[(size_t)t] = f;
    };
    auto run = [&](auto fn) {
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; t++) pool.emplace_back(fn, t);
        fn(0);
        for (auto &th : pool) th.join();
    };
    run(count);
    uint32_t f = 0;
    for (int t = 0; t < nt; t++) { kept[(size_t)t + 1] += kept[(size_t)t]; f |= fl[(size_t)t]; }
    *flags = ((f & 2u) ? 1u : 0u) | ((f & 4u) ? 2u : 0u);
    *n_out = kept[(size_t)nt];
    if (*flags) return PK_OK;
    auto compact = [&](int t) {
        size_t a, b;
        range(t, a, b);
        uint8_t *o = dst + kept[(size_t)t];
        if (simd) {
            clean_compact_avx512(src, a, b, o, dst + kept[(size_t)t + 1]);
            return;
        }
        size_t i = a;
        while (i < b) {                                    // copy line by line
            const uint8_t *nl = (const uint8_t *)memchr(src + i, '\n', b - i);
            size_t e = nl ? (size_t)(nl - src) : b;
            size_t len = e - i;
            if (len && src[e - 1] == '\r') len--;          // \r\n
            if (len && memchr(src + i, '\r', len)) {       // a lone \r ends a line as well: byte by byte
                for (size_t k = i; k < i + len; k++)
                    if (src[k] != '\r') *o++ = src[k];
            } else {
                memcpy(o, src + i, len);
                o += len;
            }
            i = e + 1;
        }
    };
    run(compact);
    return PK_OK;
}

// ---------------------------------------------------------------------------------- writer
// The output side: the documented workflow runs `bgzip -l 9` over every
// .kin (README.md:26,261-269; data/README.md:24) and the merger then reads the .kin.bgz.  A BGZF
// file is a chain of independent gzip members of at most 0xFF00 input bytes, so the members of
// one batch are deflated on all cores into fixed 64 KiB slots and then closed up in order.
//
// src[0, n) -> whole members in out[0, *produced) (no EOF member: the caller appends it once,
// at the end of the file).  out_cap >= ceil(n / 0xFF00) * 65536.  member_sizes, when not NULL,
// receives the compressed size of each of the ceil(n / 0xFF00) members -- what a .gzi index
// (gzireader.py:12-19: pairs of compressed / uncompressed block offsets) is made of.
PK_API int pk_bgzf_deflate(const uint8_t *src, size_t n, uint8_t *out, size_t out_cap, size_t *produced,
                           uint32_t *member_sizes, int level, int threads) {
    constexpr size_t kIn = 0xFF00, kSlot = 65536, kHead = 18, kTail = 8;
    PK_REQUIRE(produced != nullptr, "pk_bgzf_deflate: NULL output");
    PK_REQUIRE(n == 0 || (src != nullptr && out != nullptr), "pk_bgzf_deflate: NULL buffer");
    PK_REQUIRE(level >= 0 && level <= 9, "pk_bgzf_deflate: level %d outside 0..9", level);
    *produced = 0;
    const size_t nblk = (n + kIn - 1) / kIn;
    if (nblk == 0) return PK_OK;
    PK_REQUIRE(out_cap / kSlot >= nblk, "pk_bgzf_deflate: destination holds %zu bytes, %zu members need %zu",
               out_cap, nblk, nblk * kSlot);
    std::vector<uint32_t> sizes(nblk, 0);
    const int nt = thread_count(threads, nblk);
    std::atomic<size_t> next(0);
    std::atomic<int> bad(0);
    auto work = [&]() {
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad.store(1); return; }
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= nblk || bad.load()) break;
            const uint8_t *in = src + i * kIn;
            const size_t len = i + 1 == nblk ? n - i * kIn : kIn;
            uint8_t *slot = out + i * kSlot;
            deflateReset(&zs);
            zs.next_in = const_cast<Bytef *>(in);
            zs.avail_in = (uInt)len;
            zs.next_out = slot + kHead;
            zs.avail_out = (uInt)(kSlot - kHead - kTail);
            if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { bad.store(2); break; }
            const size_t payload = kSlot - kHead - kTail - zs.avail_out;
            const size_t total = kHead + payload + kTail;
            static const uint8_t head[16] = {0x1F, 0x8B, 8, 4, 0, 0, 0, 0, 0, 0xFF, 6, 0, 'B', 'C', 2, 0};
            memcpy(slot, head, 16);
            slot[16] = (uint8_t)((total - 1) & 0xFF);
            slot[17] = (uint8_t)((total - 1) >> 8);
            const uint32_t crc = (uint32_t)crc32(0L, in, (uInt)len), isize = (uint32_t)len;
            memcpy(slot + kHead + payload, &crc, 4);                 // little endian hosts only
            memcpy(slot + kHead + payload + 4, &isize, 4);
            sizes[i] = (uint32_t)total;
        }
        deflateEnd(&zs);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (bad.load())
        return pk_set_error(PK_ERR_ARG, "pk_bgzf_deflate: zlib %s", bad.load() == 1 ? "could not start" : "overran a 64 KiB member");
    size_t opos = 0;                                                 // close the gaps, in order
    for (size_t i = 0; i < nblk; i++) {
        if (opos != i * kSlot) memmove(out + opos, out + i * kSlot, sizes[i]);
        opos += sizes[i];
        if (member_sizes) member_sizes[i] = sizes[i];
    }
    *produced = opos;
    return PK_OK;
}

// Offsets of the headers of a block of whole lines: every '>' that starts a line (offset 0, or right
// after '\n' / '\r') -- the places where parse_fasta opens a record (indexer.py:62-80; a '>' anywhere
// else in a line is just an invalid base).  Ascending; *count is the number found even when it
// exceeds cap (the first cap are stored; the caller retries with more room).
PK_API int pk_fasta_find_headers(const uint8_t *text, size_t n, uint64_t *pos_out, size_t cap, size_t *count,
                                 int threads) {
    PK_REQUIRE(count != nullptr, "pk_fasta_find_headers: NULL count");
    PK_REQUIRE(n == 0 || text != nullptr, "pk_fasta_find_headers: NULL text");
    PK_REQUIRE(cap == 0 || pos_out != nullptr, "pk_fasta_find_headers: NULL output");
    *count = 0;
    if (n == 0) return PK_OK;
    const size_t grain = (size_t)4 << 20;
    const int nt = thread_count(threads, (n + grain - 1) / grain);
    std::vector<std::vector<uint64_t>> found((size_t)nt);
    auto scan = [&](int t) {
        const size_t a = n / (size_t)nt * (size_t)t, b = t == nt - 1 ? n : n / (size_t)nt * (size_t)(t + 1);
        size_t i = a;
        while (i < b) {
            const uint8_t *hit = (const uint8_t *)memchr(text + i, '>', b - i);
            if (!hit) break;
            const size_t p = (size_t)(hit - text);
            if (p == 0 || text[p - 1] == '\n' || text[p - 1] == '\r') found[(size_t)t].push_back(p);
            i = p + 1;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(scan, t);
    scan(0);
    for (auto &th : pool) th.join();
    size_t k = 0;
    for (const auto &v : found)
        for (uint64_t p : v) {
            if (k < cap) pos_out[k] = p;
            k++;
        }
    *count = k;
    return PK_OK;
}

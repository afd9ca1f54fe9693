// kmer_bits.h -- the bit algebra of the indexer scan, shared by the CUDA kernels
// and by a host-side unit test (tests/host/test_kmer_bits.cpp), so that the index
// arithmetic can be checked on a machine without a GPU.
//
// Replaces, from the reference (sauloal/pykmer): the CONV lookup
// (indexer.py:36-41) and the per-window arithmetic of gen_kmers
// (indexer.py:141-150) followed by pos = min(fwd, rev) (indexer.py:341).
//
// Layout.  The base stream is cut into GROUPS of 16 bases.  A group is encoded as
//   codes : 32 bits, base 0 of the group in bits 31:30 ... base 15 in bits 1:0
//   vmask : 16 bits, bit (15 - j) set iff base j is one of ACGTacgt
// so "earlier base = more significant digit", which is exactly the reference's
// first-base-most-significant encoding of a window (indexer.py:131,149): a window
// is a contiguous slice of the concatenation prev2:prev1:cur of three groups.
// The reverse complement is a slice of the digit-reversed, complemented
// concatenation rcw(cur):rcw(prev1):rcw(prev2).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PK_HD __host__ __device__ __forceinline__
#else
#define PK_HD inline
#endif

#define PK_GROUP 16

// CONV (indexer.py:36-41): bits 1:0 = code (A0 C1 G2 T3), bit 2 = valid.
PK_HD uint32_t pk_lut_entry(uint32_t c) {
    switch (c) {
        case 'A': case 'a': return 4u | 0u;
        case 'C': case 'c': return 4u | 1u;
        case 'G': case 'g': return 4u | 2u;
        case 'T': case 't': return 4u | 3u;
        default: return 0u;
    }
}

// Encode 16 bytes (w[0] holds bytes 0..3, little endian) through a 256-entry LUT.
template <typename LutT>
PK_HD void pk_encode16(const uint32_t w[4], const LutT* lut, uint32_t& codes, uint32_t& vmask) {
    uint32_t c = 0, v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
            uint32_t e = lut[(w[i] >> (8 * b)) & 0xFFu];
            c = (c << 2) | (e & 3u);
            v = (v << 1) | (e >> 2);
        }
    }
    codes = c;
    vmask = v;
}

PK_HD uint32_t pk_brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}

// Reverse the 16 two-bit digits of a group and complement them (digit d -> 3 - d).
PK_HD uint32_t pk_rcw(uint32_t x) {
    x = pk_brev32(x);                                            // digits reversed, bits in a digit swapped
    x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);     // swap back inside each digit
    return ~x;
}

// vcat: validity bits of prev2:prev1:cur (bit 15-j = base j of cur, +16 per group
// further back).  Returns W with bit p = AND(vcat[p .. p+K-1]); the window that
// ENDS at base j of cur is valid iff bit (15 - j) of W is set (indexer.py:144).
PK_HD uint64_t pk_valid_windows(uint64_t vcat, int K) {
    uint64_t W = vcat;
    int run = 1;
    while (run * 2 <= K) { W &= W >> run; run *= 2; }
    if (run < K) W &= W >> (K - run);
    return W;
}

PK_HD uint64_t pk_kmer_mask(int K) { return K >= 32 ? ~0ull : ((1ull << (2 * K)) - 1ull); }

// Forward value of the window ending at base j of cur: digits j-K+1 .. j, first
// base most significant (indexer.py:149).  pc2 is only needed when K > 17.
PK_HD uint64_t pk_fwd_at(uint32_t pc2, uint32_t pc1, uint32_t cc, int j, int K) {
    const int s = 2 * (15 - j);                                  // 0..30
    uint64_t lo = (((uint64_t)pc1 << 32) | cc) >> s;
    if (s) lo |= (uint64_t)pc2 << (64 - s);
    return lo & pk_kmer_mask(K);
}

// Reverse-complement value of the same window (indexer.py:150): a slice of
// r0:r1:r2 = rcw(cur):rcw(prev1):rcw(prev2) starting at bit t = 66 + 2j - 2K.
PK_HD uint64_t pk_rc_at(uint32_t r2, uint32_t r1, uint32_t r0, int j, int K) {
    const int t = 66 + 2 * j - 2 * K;                            // >= 4 for K <= 31
    uint64_t v;
    if (t >= 64)      v = (uint64_t)r0 >> (t - 64);
    else if (t >= 32) v = (((uint64_t)r0 << 32) | r1) >> (t - 32);
    else              v = ((((uint64_t)r1 << 32) | r2) >> t) | ((uint64_t)r0 << (64 - t));
    return v & pk_kmer_mask(K);
}

// 32-bit fast forms for K <= 16 (one halo group).
PK_HD uint32_t pk_fwd32_at(uint64_t cat /* pc1:cc */, int j, uint32_t mask) {
    return (uint32_t)(cat >> (2 * (15 - j))) & mask;
}
PK_HD uint32_t pk_rc32_at(uint64_t rcat /* r0:r1 */, int j, int K, uint32_t mask) {
    return (uint32_t)(rcat >> (34 + 2 * j - 2 * K)) & mask;      // 2 <= shift <= 62 for K <= 16
}

// ---------------------------------------------------------------------------------------
// All windows that END inside one 16-base group, given the group's own encoding
// (cc, cv) and its one or two predecessors.  Windows whose canonical value lies in
// [lo, lo + span) are counted; consecutive equal canonical values are merged into runs.
//   emit(slot, off, cnt)   one run: off = canon - lo, cnt = 1..16 windows.  `slot`
//                          (0..16) is a compile-time constant after unrolling, so a
//                          caller may keep per-slot state in registers.  off is uint32_t
//                          when !WIDE (K <= 16), uint64_t otherwise.
// FULL = the whole table (lo = 0, span = 4^K): no range test at all.
// Returns the 16-bit mask of counted windows (bit 15-j = the window ending at base j);
// its popcount is the contribution to num_kmers (indexer.py:342).
template <bool WIDE, bool FULL, typename Emit>
PK_HD uint32_t pk_scan_group(int K, uint64_t lo, uint64_t span, uint32_t cc, uint32_t cv,
                             uint32_t pc1, uint32_t pv1, uint32_t pc2, uint32_t pv2, Emit emit) {
    const uint64_t vcat = ((uint64_t)pv2 << 32) | ((uint64_t)pv1 << 16) | cv;
    const uint32_t Wm = (uint32_t)pk_valid_windows(vcat, K) & 0xFFFFu;
    if (!Wm) return 0;
    const uint32_t r0 = pk_rcw(cc), r1 = pk_rcw(pc1);
    uint32_t counted = 0, pend = 0;
    if (WIDE) {
        const uint32_t r2 = pk_rcw(pc2);
        // pk_rc_at slices r0:r1:r2 at bit 66 + 2j - 2K: shift the 96-bit value once per group by the
        // K-dependent part (4..32 bits for K = 17..31), so that the per-window shifts are the constants 2j
        const int sh0 = 66 - 2 * K;
        const uint64_t rp_lo = ((((uint64_t)r1 << 32) | r2) >> sh0) | ((uint64_t)r0 << (64 - sh0));
        const uint64_t rp_hi = (uint64_t)r0 >> sh0;
        const uint64_t kmask = pk_kmer_mask(K);
        if (FULL && Wm == 0xFFFFu) {                              // see the 32-bit form below
            uint64_t prev;
            {
                const uint64_t f = pk_fwd_at(pc2, pc1, cc, 0, K), r = rp_lo & kmask;
                prev = f < r ? f : r;
            }
#pragma unroll
            for (int j = 1; j < 16; j++) {
                const uint64_t f = pk_fwd_at(pc2, pc1, cc, j, K);
                const uint64_t r = ((rp_lo >> (2 * j)) | (rp_hi << (64 - 2 * j))) & kmask;
                const uint64_t off = f < r ? f : r;
                if (off == prev) {
                    pend++;
                } else {
                    emit(j, prev, pend + 1u);
                    prev = off;
                    pend = 0;
                }
            }
            emit(16, prev, pend + 1u);
            return 0xFFFFu;
        }
        uint64_t prev = ~0ull;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (!((Wm >> (15 - j)) & 1u)) continue;
            const uint64_t f = pk_fwd_at(pc2, pc1, cc, j, K);
            const uint64_t r = (j ? (rp_lo >> (2 * j)) | (rp_hi << (64 - 2 * j)) : rp_lo) & kmask;
            const uint64_t off = (f < r ? f : r) - lo;            // indexer.py:341
            if (!FULL && off >= span) continue;                   // another shard's k-mer
            counted |= 1u << (15 - j);
            if (off == prev) {
                pend++;
            } else {
                if (pend) emit(j, prev, pend);
                prev = off;
                pend = 1;
            }
        }
        if (pend) emit(16, prev, pend);
    } else {
        // K <= 16: everything fits 32 bits; the window slices are funnel shifts by constants
        const uint32_t mask = (uint32_t)pk_kmer_mask(K);
        const uint64_t cat = ((uint64_t)pc1 << 32) | cc;
        const uint64_t rsh = (((uint64_t)r0 << 32) | r1) >> (34 - 2 * K);
        const uint32_t lo32 = (uint32_t)lo;
        const uint32_t span_m1 = (uint32_t)(span - 1);           // span <= 2^32
        if (FULL && Wm == 0xFFFFu) {
            // all 16 windows valid and no range to test -- nearly every group of a genome: no validity
            // test per window, and a run is always open after the first one
            uint32_t prev = (uint32_t)(cat >> 30) & mask;
            {
                const uint32_t r = (uint32_t)rsh & mask;
                prev = prev < r ? prev : r;
            }
#pragma unroll
            for (int j = 1; j < 16; j++) {
                const uint32_t f = (uint32_t)(cat >> (2 * (15 - j))) & mask;
                const uint32_t r = (uint32_t)(rsh >> (2 * j)) & mask;
                const uint32_t off = f < r ? f : r;               // indexer.py:341
                if (off == prev) {
                    pend++;
                } else {
                    emit(j, prev, pend + 1u);
                    prev = off;
                    pend = 0;
                }
            }
            emit(16, prev, pend + 1u);
            return 0xFFFFu;
        }
        uint32_t prev = 0xFFFFFFFFu;                             // no run yet (never a valid offset + count)
        bool have = false;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (!((Wm >> (15 - j)) & 1u)) continue;
            const uint32_t f = (uint32_t)(cat >> (2 * (15 - j))) & mask;
            const uint32_t r = (uint32_t)(rsh >> (2 * j)) & mask;
            const uint32_t off = (f < r ? f : r) - lo32;          // indexer.py:341
            if (!FULL && off > span_m1) continue;                 // another shard's k-mer
            counted |= 1u << (15 - j);
            if (have && off == prev) {
                pend++;
            } else {
                if (have) emit(j, prev, pend);
                prev = off;
                pend = 1;
                have = true;
            }
        }
        if (have) emit(16, prev, pend);
    }
    return counted;
}

// Which runs of counted windows may lie in different records: calls visit(j) once for the
// first counted window after every invalid base of the group (and for the group's first
// counted window).  cv = validity bits of the group, counted = mask from pk_scan_group.
template <typename Visit>
PK_HD void pk_for_each_record_run(uint32_t cv, uint32_t counted, Visit visit) {
    while (counted) {
        int top = 31;
        while (!((counted >> top) & 1u)) top--;                  // earliest remaining window
        const int j = 15 - top;
        visit(j);
        // drop every counted window up to the next invalid base after j
        uint32_t inv = (~cv) & ((1u << top) - 1u) & 0xFFFFu;     // invalid bases later than j
        if (!inv) break;
        int q = 31;
        while (!((inv >> q) & 1u)) q--;                          // first invalid base after j
        counted &= (1u << q) - 1u;                               // keep windows ending after it
    }
}

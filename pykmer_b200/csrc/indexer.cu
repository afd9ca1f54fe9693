// indexer.cu -- the indexer hot path on B200 (sm_100a).
//
// Replaces, from the reference (sauloal/pykmer): CONV (indexer.py:36-41),
// gen_kmers (indexer.py:130-160), pos = min(fwd, rev) / num_kmers
// (indexer.py:341-342), the `chromosomes` rule (indexer.py:349-351),
// process_kmers' saturating accumulate (indexer.py:239,262) and
// Header.update_stats (tools.py:246-263).
//
// Every scan kernel shares one front end: a thread encodes one 16-base group (one
// coalesced 16-byte load, 256-byte LUT in shared memory), fetches the K-1 base halo
// from its neighbour lanes with warp shuffles, and slices every window that ends in
// its group out of the 2-bit concatenation (pk_scan_group, kmer_bits.h).
//
// Two counting schemes (measured on B200, profiles/r01_microbench_b200.txt):
//   DIRECT     k_scan_count_direct: byte-granular compare-and-swap straight into the
//              table.  Random DRAM sectors: ~20 G updates/s.  Used for small tables
//              (they sit in L2) and as the fallback.
//   PARTITION  the table is cut into WINDOWS of 2^24 entries whose 32-bit counters
//              (64 MiB) stay resident in the 126 MB L2, where red.global.add.u32 runs
//              at ~190 G/s instead of ~20 G/s:
//                k_scan_bucket_count  pass 1: how many k-mers fall in each window
//                k_bucket_offsets     exclusive scan -> segment offsets in the pool
//                k_scan_scatter       pass 2: write (offset-in-window, run length)
//                                     entries, coalesced, into per-window segments
//                k_window_count       per window: L2-resident red.add.u32
//                k_window_commit      per window: clamp to 255 (saturating accumulate),
//                                     write the table once, re-zero the counters, and
//                                     histogram the final bytes (fused update_stats)
//   k_table_stats              stand-alone 256-bin histogram pass (DIRECT mode).
//   k_update_carry             keeps the last 32 stream bytes for the next feed.
#include <algorithm>
#include <atomic>
#include <new>
#include <thread>
#include <type_traits>
#include <vector>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "kmer_bits.h"

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr size_t kStageBytes = 32u << 20;   // pinned-copy chunk of feed_host
constexpr size_t kPackChunk = 1024;          // entries per chunk of a packed table slice (unpack.cpp: kChunk)
constexpr int kCarry = 32;                  // bytes of stream tail kept between feeds
constexpr int kMaxBuckets = 16384;          // windows per handle in PARTITION / SCAN mode
constexpr int kMaxSegments = 128;           // feeds buffered between two flushes
constexpr size_t kMaxFeed = 1u << 30;       // bases per partition pass (one segment each)
constexpr int kTileEntries = kScanWarps * 31 * 16;   // most entries one block tile can emit
// a buffered k-mer entry: (run length - 1) << 28 | offset inside the window.  A run never
// exceeds the 16 windows of one group (pk_scan_group), windows never exceed 2^26 entries.
constexpr uint32_t kEntShift = 28;
constexpr uint32_t kEntMask = (1u << kEntShift) - 1u;

struct ScanParams {
    const uint8_t *seq;       // this feed (16-byte aligned)
    size_t n;
    const uint8_t *carry;     // last kCarry bytes of everything fed before
    int K;
    uint64_t lo, span;        // canonical range [lo, lo + span) owned by this handle
    uint8_t *table;           // [span]                                  (DIRECT)
    unsigned long long *num_kmers;
    unsigned long long *bins;    // [256] histogram of the table, kept as transitions   (DIRECT)
    const uint64_t *rec_starts;
    size_t nrec;
    uint8_t *rec_flags;
    uint64_t stream_off;      // stream offset of seq[0]
    long long ngroups, ntiles;   // 16-base groups / warp tiles of this feed (the host divides once: the kernels
                                 // otherwise redo the 64-bit division by 31 in every tile iteration, ncu source view)
    // PARTITION
    uint32_t win_log2, nbuckets;
    uint32_t *seg_cnt;        // [nbuckets] entries of this feed per window (pass 1 out)
    const uint32_t *seg_off;  // [nbuckets] pool index of each segment      (pass 2 in)
    uint32_t *seg_fill;       // [nbuckets] fill cursors                    (pass 2)
    // estimated pass 1 (see indexer_launch_scan): pass 1 looks at one tile in 2^sample_shift and
    // does no bookkeeping; pass 2 checks its reservations against seg_cap, counts num_kmers into
    // ctl and flags the records.  Kernels given run_if return at once when *run_if == 0.
    uint32_t sample_shift;
    uint32_t pass1_tally;     // pass 1 adds to num_kmers and flags records (the exact pass)
    uint32_t pass2_tally;     // pass 2 does (the estimated pass)
    const uint32_t *seg_cap;  // [nbuckets] room reserved per window, NULL = exact sizes
    uint32_t *ctl;            // [0] pool cursor [1] overflow flag [2] cursor before this feed [4..5] num_kmers of pass 2
    const uint32_t *run_if;
    uint32_t *pool;
    // pass 2 with remote destinations (sequence-sharded multi-GPU): window b's entries go to
    // peer[win_owner[b]] + dest_off[b] -- peer-mapped pools of the window owners (NVLink stores)
    const uint32_t *win_owner;
    const uint32_t *dest_off;
    uint32_t *peer[16];
};

__device__ __forceinline__ void load_group(const ScanParams &p, long long g, long long ngroups,
                                           uint32_t w[4]) {
    if (g < 0) {                                         // halo from the previous feed
        const uint4 q = *reinterpret_cast<const uint4 *>(p.carry + kCarry + 16 * g);
        w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
        return;
    }
    const size_t off = (size_t)g * 16;
    if (g < ngroups && off + 16 <= p.n) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p.seq + off));
        w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
        return;
    }
    w[0] = w[1] = w[2] = w[3] = 0;                       // past the end: invalid bases
    if (g < ngroups) {
#pragma unroll
        for (int b = 0; b < 16; b++)
            if (off + b < p.n) w[b >> 2] |= (uint32_t)p.seq[off + b] << (8 * (b & 3));
    }
}

// The histogram (tools.py:250) is kept incrementally: an atomic that returns the old value tells
// exactly how its lane moved, so bins[old]--, bins[new]++ follows the table without ever reading
// it back.  Per-block transitions; 0 -> 1 (most of a sparse table) stays in a register.
struct HistTally {
    uint32_t *sh;                   // [256] wrapping counters of the block
    unsigned long long *g;          // global bins[256]
    uint32_t c1;
    __device__ __forceinline__ void move(uint32_t from, uint32_t to) {
        if (from == 0u && to == 1u) { c1++; return; }
        if (from == to) return;
        if (from) atomicSub(sh + from, 1u);
        atomicAdd(sh + to, 1u);
    }
    // all threads of the block: add the block's transitions to the global bins
    __device__ __forceinline__ void flush() {
        if (c1) atomicAdd(sh + 1, c1);
        c1 = 0;
        __syncthreads();
        for (uint32_t v = threadIdx.x; v < 256; v += blockDim.x) {
            const int d = (int)sh[v];
            if (d) atomicAdd(g + v, (unsigned long long)(long long)d);
            sh[v] = 0;
        }
        __syncthreads();
    }
};

__device__ __forceinline__ long long find_record(const uint64_t *starts, size_t nrec, uint64_t pos) {
    size_t lo = 0, hi = nrec;                            // upper_bound - 1
    while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if (__ldg(starts + mid) <= pos) lo = mid + 1; else hi = mid;
    }
    return (long long)lo - 1;
}

// indexer.py:349-351: a record is listed once a k-mer of it has been counted.  Runs of
// counted windows separated by an invalid base (a separator, maybe) are looked up apart.
__device__ __forceinline__ void flag_records(const ScanParams &p, long long g, uint32_t cv,
                                             uint32_t counted) {
    if (!p.rec_flags || !counted) return;
    pk_for_each_record_run(cv, counted, [&](int j) {
        const long long r = find_record(p.rec_starts, p.nrec, p.stream_off + (uint64_t)g * 16 + j);
        if (r >= 0 && !p.rec_flags[r]) p.rec_flags[r] = 1;
    });
}

// The same for a whole warp tile (all 32 lanes, converged; cm = 0 for lanes that count nothing).
// A tile covers 496 bases and records are long, so nearly always the first and the last counted
// window of the tile lie in one record, and nearly always it is the record the warp met last
// time: RecCache keeps that record's extent (warp-uniform), two compares settle the tile.
struct RecCache {
    uint64_t lo = 1, hi = 0;        // stream extent [lo, hi) of the cached record (empty at first)
    long long r = -1;
};

__device__ __forceinline__ void flag_records_warp(const ScanParams &p, long long g, uint32_t cv, uint32_t cm,
                                                  RecCache &rc) {
    if (!p.rec_flags) return;
    const unsigned have = __ballot_sync(0xFFFFFFFFu, cm != 0u);
    if (!have) return;
    const int lane = threadIdx.x & 31, first = __ffs((int)have) - 1, last = 31 - __clz((int)have);
    const uint32_t cm_first = __shfl_sync(0xFFFFFFFFu, cm, first), cm_last = __shfl_sync(0xFFFFFFFFu, cm, last);
    const uint64_t base = p.stream_off + (uint64_t)(g - lane) * 16;          // group of lane 0
    const uint64_t pos_lo = base + (uint64_t)first * 16 + (uint64_t)(15 - (31 - __clz((int)cm_first)));
    const uint64_t pos_hi = base + (uint64_t)last * 16 + (uint64_t)(15 - (__ffs((int)cm_last) - 1));
    if (pos_lo >= rc.lo && pos_hi < rc.hi) return;        // the record flagged last time
    const long long r_lo = find_record(p.rec_starts, p.nrec, pos_lo);
    const long long r_hi = find_record(p.rec_starts, p.nrec, pos_hi);
    if (r_lo == r_hi) {
        if (r_lo < 0) return;
        if (lane == 0 && !p.rec_flags[r_lo]) p.rec_flags[r_lo] = 1;
        rc.r = r_lo;
        rc.lo = __ldg(p.rec_starts + r_lo);
        rc.hi = (size_t)(r_lo + 1) < p.nrec ? __ldg(p.rec_starts + r_lo + 1) : ~0ull;
    } else {
        flag_records(p, g, cv, cm);
    }
}

// One warp tile: lane l encodes group tile*GPW - H + l and receives its halo by shuffle.
template <bool WIDE>
struct WarpTile {
    static constexpr int H = WIDE ? 2 : 1;      // halo groups: K-1 <= 16*H bases
    static constexpr int GPW = 32 - H;          // groups a warp emits per tile
    long long g;
    uint32_t cc, cv, pc1, pv1, pc2, pv2;
    bool emits;
    __device__ __forceinline__ void load(const ScanParams &p, long long tile, long long ngroups,
                                         const uint8_t *lut) {
        const int lane = threadIdx.x & 31;
        g = tile * GPW - H + lane;
        uint32_t w[4];
        load_group(p, g, ngroups, w);
        pk_encode16(w, lut, cc, cv);
        pc1 = __shfl_up_sync(0xFFFFFFFFu, cc, 1);
        pv1 = __shfl_up_sync(0xFFFFFFFFu, cv, 1);
        pc2 = 0; pv2 = 0;
        if (WIDE) {
            pc2 = __shfl_up_sync(0xFFFFFFFFu, cc, 2);
            pv2 = __shfl_up_sync(0xFFFFFFFFu, cv, 2);
        }
        emits = lane >= H && g >= 0 && g < ngroups;
    }
};

__device__ __forceinline__ void add_num_kmers(unsigned long long *dst, unsigned long long counted) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) counted += __shfl_xor_sync(0xFFFFFFFFu, counted, o);
    if ((threadIdx.x & 31) == 0 && counted) atomicAdd(dst, counted);
}

// ------------------------------------------------------------------------------ DIRECT
// A warp first queues the runs of its tile in shared memory and only then applies them, every
// lane taking queue entries in turn: the table updates are independent memory operations in
// flight together (four per lane) instead of a chain of DRAM round trips inside the window
// loop, and all 32 lanes work even when a shard keeps one k-mer in eight (measured at K=19, one
// shard of eight: 24 ms chained).  SPARSE (K > 16): the table is almost empty, so the first
// attempt is a blind compare-and-swap against zero -- one round trip instead of load +
// compare-and-swap.  Queue entry: (run length - 1) << 60 | offset (offsets stay below the
// shard's span, far below 2^60).
template <bool WIDE, bool FULL>
__global__ void __launch_bounds__(kScanThreads) k_scan_count_direct(const ScanParams p) {
    __shared__ uint8_t lut[256];
    __shared__ uint32_t s_bins[256];
    __shared__ unsigned long long s_q[kScanWarps][32 * 17];
    __shared__ uint32_t s_qn[kScanWarps];
    lut[threadIdx.x] = (uint8_t)pk_lut_entry(threadIdx.x);
    s_bins[threadIdx.x] = 0;
    __syncthreads();
    HistTally ht{s_bins, p.bins, 0u};
    using WT = WarpTile<WIDE>;
    constexpr bool SPARSE = WIDE;
    constexpr int U = 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long *q = s_q[warp];
    const long long ngroups = p.ngroups, ntiles = p.ntiles;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    unsigned long long counted = 0;
    RecCache rcache;
    for (long long tile = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < ntiles;
         tile += nwarps) {
        if (lane == 0) s_qn[warp] = 0;
        __syncwarp();
        WT t;
        t.load(p, tile, ngroups, lut);
        uint32_t cm = 0;
        if (t.emits) {
            cm = pk_scan_group<WIDE, FULL>(
                p.K, p.lo, p.span, t.cc, t.cv, t.pc1, t.pv1, t.pc2, t.pv2,
                [&](int, auto o, uint32_t c) {
                    q[atomicAdd(&s_qn[warp], 1u)] = ((unsigned long long)(c - 1u) << 60) | (unsigned long long)o;
                });
            counted += __popc(cm);
        }
        __syncwarp();
        flag_records_warp(p, t.g, t.cv, cm, rcache);
        const uint32_t nq = s_qn[warp];
        for (uint32_t base = 0; base < nq; base += 32 * U) {
            unsigned long long off[U];
            uint32_t c[U], seen[U];
#pragma unroll
            for (int u = 0; u < U; u++) {                 // first round: independent operations
                const uint32_t i = base + u * 32 + lane;
                const unsigned long long e = i < nq ? q[i] : ~0ull;
                off[u] = e == ~0ull ? ~0ull - (unsigned)lane : e & ((1ull << 60) - 1ull);   // dead lanes: unique
                c[u] = e == ~0ull ? 0u : (uint32_t)(e >> 60) + 1u;
                // microsatellites put the same two or three k-mers into every lane: one update per
                // distinct offset of the batch (cheap screen first, MATCH.ANY only when it fires)
                const uint32_t h = (uint32_t)off[u] ^ (uint32_t)(off[u] >> 32);
                const uint32_t h1 = __shfl_up_sync(0xFFFFFFFFu, h, 1), h2 = __shfl_up_sync(0xFFFFFFFFu, h, 2),
                               h3 = __shfl_up_sync(0xFFFFFFFFu, h, 3);
                const bool dup = (lane >= 1 && h1 == h) || (lane >= 2 && h2 == h) || (lane >= 3 && h3 == h);
                if (__any_sync(0xFFFFFFFFu, dup)) {
                    const unsigned peers = __match_any_sync(0xFFFFFFFFu, off[u]);
                    if (peers != (1u << lane)) {
                        const uint32_t sum = __reduce_add_sync(peers, c[u]);
                        c[u] = lane == (__ffs((int)peers) - 1) ? min(sum, 255u) : 0u;
                    }
                }
                if (!c[u]) continue;
                uint32_t *wp = reinterpret_cast<uint32_t *>(p.table + (off[u] & ~3ull));
                seen[u] = SPARSE ? atomicCAS(wp, 0u, c[u] << (8u * ((uint32_t)off[u] & 3u))) : __ldcg(wp);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {                 // second round: settle against what was seen
                if (!c[u]) continue;
                if (SPARSE && seen[u] == 0u) { ht.move(0u, c[u]); continue; }
                uint32_t *wp = reinterpret_cast<uint32_t *>(p.table + (off[u] & ~3ull));
                const uint32_t sh = 8u * ((uint32_t)off[u] & 3u);
                uint32_t old = seen[u];
                for (;;) {                                // table[idx] = min(255, table[idx] + c), indexer.py:239,262;
                                                          // counters only grow, so a read of 255 is final
                    const uint32_t b = (old >> sh) & 0xFFu;
                    if (b == 255u) break;
                    const uint32_t nb = min(255u, b + c[u]);
                    const uint32_t assumed = old;
                    old = atomicCAS(wp, assumed, (assumed & ~(0xFFu << sh)) | (nb << sh));
                    if (old == assumed) { ht.move(b, nb); break; }
                }
            }
        }
        __syncwarp();
    }
    add_num_kmers(p.num_kmers, counted);
    ht.flush();                                            // histogram of the table, kept as transitions
}

// ------------------------------------------------------------------------------ PARTITION
// pass 1: per-window entry counts of this feed (+ num_kmers and record flags); with
// sample_shift, of a pseudo-random 1 / 2^sample_shift of the tiles only
__device__ __forceinline__ uint32_t tile_hash(long long tile) {
    uint32_t x = (uint32_t)tile * 0x9E3779B1u;
    x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13;
    return x;
}

template <bool WIDE, bool FULL>
__global__ void __launch_bounds__(kScanThreads) k_scan_bucket_count(const ScanParams p) {
    if (p.run_if && *p.run_if == 0u) return;
    extern __shared__ uint32_t sm[];
    uint32_t *s_cnt = sm;                                  // [nbuckets]
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = (uint8_t)pk_lut_entry(threadIdx.x);
    for (uint32_t b = threadIdx.x; b < p.nbuckets; b += blockDim.x) s_cnt[b] = 0;
    __syncthreads();
    using WT = WarpTile<WIDE>;
    const long long ngroups = p.ngroups, ntiles = p.ntiles;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const uint32_t wl = p.win_log2;
    const uint32_t smask = (1u << p.sample_shift) - 1u;
    unsigned long long counted = 0;
    RecCache rcache;
    for (long long tile = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < ntiles;
         tile += nwarps) {
        if (smask && (tile_hash(tile) & smask)) continue;  // warp-uniform
        WT t;
        t.load(p, tile, ngroups, lut);
        uint32_t cm = 0;
        if (t.emits)
            cm = pk_scan_group<WIDE, FULL>(
                p.K, p.lo, p.span, t.cc, t.cv, t.pc1, t.pv1, t.pc2, t.pv2,
                [&](int, auto off, uint32_t) { atomicAdd(&s_cnt[(uint32_t)(off >> wl)], 1u); });
        __syncwarp();
        if (p.pass1_tally) {
            counted += __popc(cm);
            flag_records_warp(p, t.g, t.cv, cm, rcache);
        }
    }
    if (p.pass1_tally) add_num_kmers(p.num_kmers, counted);
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < p.nbuckets; b += blockDim.x)
        if (s_cnt[b]) atomicAdd(&p.seg_cnt[b], s_cnt[b]);
}

// exclusive scan of one feed's window counts -> pool offsets; advances the pool cursor.
// ESTIMATE: cnt holds the sampled counts; every window gets room for
// est * 2^shift * (1 + 1/16) + kCapSlack entries (stored in cap, cnt goes back to zero), unless
// that adds up to more than `budget` entries -- then the overflow flag sends the feed down the
// exact path straight away.
constexpr uint32_t kCapSlack = 16384;

template <bool ESTIMATE>
__global__ void __launch_bounds__(256) k_bucket_offsets(uint32_t *__restrict__ cnt,
                                                        uint32_t *__restrict__ off, uint32_t nb,
                                                        uint32_t *__restrict__ ctl, uint32_t *__restrict__ cap,
                                                        uint32_t shift, uint32_t budget, uint32_t test,
                                                        const uint32_t *__restrict__ run_if) {
    if (run_if && *run_if == 0u) return;
    __shared__ uint32_t part[256];
    __shared__ uint32_t s_over;
    const uint32_t per = (nb + 255) / 256;
    const uint32_t b0 = threadIdx.x * per, b1 = min(nb, b0 + per);
    auto room = [&](uint32_t c) -> uint32_t {
        if (!ESTIMATE) return c;
        const unsigned long long e = (unsigned long long)c << shift;
        const unsigned long long r = (test & 2u) ? e >> 1 : e + (e >> 4) + kCapSlack;   // test & 2: too little room
        return r > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)r;
    };
    unsigned long long s64 = 0;
    for (uint32_t b = b0; b < b1; b++) s64 += room(cnt[b]);
    part[threadIdx.x] = s64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)s64;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = ctl[0];
        unsigned long long total = 0;
        if (ESTIMATE) ctl[2] = run;                       // where this feed began
        for (int i = 0; i < 256; i++) { const uint32_t v = part[i]; part[i] = run; run += v; total += v; }
        s_over = ESTIMATE && (total > budget || (test & 1u)) ? 1u : 0u;   // test & 1: as if over budget
        if (s_over) ctl[1] = 1u; else ctl[0] = run;
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (uint32_t b = b0; b < b1; b++) {
        const uint32_t r = s_over ? 0u : room(cnt[b]);
        off[b] = run;
        run += r;
        if (ESTIMATE) { cap[b] = r; cnt[b] = 0; }
    }
}

// after the estimated pass 2: either adopt its fill counts (and its num_kmers), or -- a window
// overflowed its reservation -- forget the attempt, so that the exact kernels, which are queued
// behind and do nothing otherwise, redo the feed
__global__ void __launch_bounds__(256) k_feed_settle(uint32_t *__restrict__ cnt, uint32_t *__restrict__ fill,
                                                     uint32_t nb, uint32_t *__restrict__ ctl,
                                                     unsigned long long *__restrict__ num_kmers) {
    const bool over = ctl[1] != 0u;
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) {
        cnt[b] = over ? 0u : fill[b];
        if (over) fill[b] = 0u;
    }
    if (threadIdx.x == 0) {
        unsigned long long *tmp = reinterpret_cast<unsigned long long *>(ctl + 4);
        if (over) ctl[0] = ctl[2]; else *num_kmers += *tmp;
        *tmp = 0ull;
    }
}

// ---- routed scan (sequence-sharded multi-GPU without a host round trip) -------------------------
// Every (source rank, window) owns a fixed region of the window owner's k-mer buffer, sized once from
// a planning scan; pass 2 stores straight into it (capacity-checked like the estimated pass 1).  The
// fill counts then travel the same way: k_publish_fill writes them into a small table at the tail of
// every owner's buffer, and after one stream-ordered collective k_adopt_published turns that table
// into the owner's segment tables.  kPubEntries 32-bit slots: [source rank][window of the whole job].
constexpr size_t kPubEntries = (size_t)1 << 18;        // 16 ranks x 16384 windows

struct PublishPeers {
    uint32_t *pool[16];
    unsigned long long pub_base[16];
};

__global__ void __launch_bounds__(256) k_publish_fill(const uint32_t *__restrict__ fill,
                                                      const uint32_t *__restrict__ cap, uint32_t nb,
                                                      const uint32_t *__restrict__ owner, const PublishPeers pp,
                                                      uint32_t self_rank, uint32_t *__restrict__ ctl,
                                                      unsigned long long *__restrict__ num_kmers,
                                                      uint32_t *__restrict__ status) {
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
        const uint32_t o = owner[b];
        pp.pool[o][pp.pub_base[o] + (size_t)self_rank * nb + b] = min(fill[b], cap[b]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long *tmp = reinterpret_cast<unsigned long long *>(ctl + 4);
        *num_kmers += *tmp;
        *tmp = 0ull;
        if (status) {
            status[0] = ctl[1];                           // 1: a region overflowed, the step must be redone exactly
            status[1] = (uint32_t)*num_kmers;             // this scanner's windows counted since its reset,
            status[2] = (uint32_t)(*num_kmers >> 32);     // so that no separate read-back is needed
        }
    }
}

__global__ void __launch_bounds__(256) k_adopt_published(const uint32_t *__restrict__ tail, uint32_t nranks,
                                                         uint32_t nb_total, uint32_t w0, uint32_t nb_local,
                                                         const uint32_t *__restrict__ import_off,
                                                         uint32_t *__restrict__ seg_cnt0,
                                                         uint32_t *__restrict__ seg_off0) {
    const uint32_t n = nranks * nb_local;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t src = i / nb_local, w = i % nb_local;
        seg_cnt0[i] = tail[(size_t)src * nb_total + w0 + w];
        seg_off0[i] = import_off[i];
    }
}

// pass 2: the same scan again; entries are ranked per window in shared memory, staged
// window by window, and written out as contiguous runs into the segments pass 1 sized.
// entry = (run length - 1) << kEntShift | offset inside the window.
template <bool WIDE, bool FULL>
__global__ void __launch_bounds__(kScanThreads, 4) k_scan_scatter(const ScanParams p) {
    if (p.run_if && *p.run_if == 0u) return;
    if (p.seg_cap && p.ctl[1] != 0u) return;               // over budget already: straight to the exact path
    extern __shared__ uint32_t sm[];
    const uint32_t nb = p.nbuckets;
    uint32_t *s_cnt = sm;                                  // [nb] entries of this tile per window
    uint32_t *s_toff = sm + nb;                            // [nb] exclusive offsets inside the tile
    uint32_t *s_gbase = sm + 2 * nb;                       // [nb] pool index reserved for the tile
    uint32_t *s_ent = sm + 3 * nb;                         // [kTileEntries]
    uint16_t *s_bid = reinterpret_cast<uint16_t *>(s_ent + kTileEntries);   // [kTileEntries]
    __shared__ uint8_t lut[256];
    __shared__ uint32_t s_part[kScanThreads];
    __shared__ uint32_t s_total;
    __shared__ uint32_t *s_peer[16];                       // the owners' buffers (remote destinations only)
    lut[threadIdx.x] = (uint8_t)pk_lut_entry(threadIdx.x);
    if (threadIdx.x < 16) s_peer[threadIdx.x] = p.peer[threadIdx.x];

    using WT = WarpTile<WIDE>;
    const long long ngroups = p.ngroups, ntiles = p.ntiles;
    const long long nblock_tiles = (ntiles + kScanWarps - 1) / kScanWarps;
    const uint32_t wl = p.win_log2, wmask = (1u << wl) - 1u;
    const int warp = threadIdx.x >> 5;
    const uint32_t per = (nb + kScanThreads - 1) / kScanThreads;
    unsigned long long counted = 0;
    RecCache rcache;

    for (long long bt = blockIdx.x; bt < nblock_tiles; bt += gridDim.x) {
        for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) s_cnt[b] = 0;
        __syncthreads();
        // A: scan; keep every run in registers with its window and its rank in the tile
        uint32_t ent[17], key[17];
#pragma unroll
        for (int s = 0; s < 17; s++) key[s] = 0xFFFFFFFFu;
        const long long tile = bt * kScanWarps + warp;
        if (tile < ntiles) {
            WT t;
            t.load(p, tile, ngroups, lut);
            {   // the bases of this warp's NEXT tile: into L1 while this tile is sorted (the load at the top of a
                // tile was 12 % of the kernel's stall samples, long scoreboard)
                const long long gn = t.g + (long long)gridDim.x * kScanWarps * WT::GPW;
                if (gn >= 0 && gn < ngroups) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.seq + (size_t)gn * 16));
            }
            uint32_t cm = 0;
            if (t.emits)
                cm = pk_scan_group<WIDE, FULL>(
                    p.K, p.lo, p.span, t.cc, t.cv, t.pc1, t.pv1, t.pc2, t.pv2,
                    [&](int slot, auto off, uint32_t cnt) {
                        const uint32_t b = (uint32_t)(off >> wl);
                        const uint32_t rank = atomicAdd(&s_cnt[b], 1u);
                        ent[slot] = ((uint32_t)off & wmask) | ((cnt - 1u) << kEntShift);
                        key[slot] = (b << 16) | rank;
                    });
            __syncwarp();
            if (p.pass2_tally) {
                counted += __popc(cm);
                flag_records_warp(p, t.g, t.cv, cm, rcache);
            }
        }
        __syncthreads();
        // B: exclusive scan over the windows; reserve the tile's share of every segment
        uint32_t def_c = 0, def_room = 0, def_pos = 0;     // deferred reservation of this thread's window (per == 1)
        {
            const uint32_t b0 = threadIdx.x * per, b1 = min(nb, b0 + per);
            uint32_t s = 0;
            for (uint32_t b = b0; b < b1; b++) s += s_cnt[b];
            s_part[threadIdx.x] = s;
            __syncthreads();
            if (warp == 0) {                               // scan 256 partials with one warp
                uint32_t v[8], sum = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) { v[i] = s_part[threadIdx.x * 8 + i]; sum += v[i]; }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if ((int)(threadIdx.x & 31) >= o) incl += up;
                }
                uint32_t run = incl - sum;
#pragma unroll
                for (int i = 0; i < 8; i++) { s_part[threadIdx.x * 8 + i] = run; run += v[i]; }
                if (threadIdx.x == 31) s_total = incl;
            }
            __syncthreads();
            uint32_t run = s_part[threadIdx.x];
            auto reserve = [&](uint32_t b, uint32_t c, uint32_t room, uint32_t pos) {
                if (pos + c > room) {                               // the estimate fell short: redo exactly
                    if (*reinterpret_cast<volatile uint32_t *>(p.ctl + 1) == 0u) atomicExch(p.ctl + 1, 1u);
                    s_gbase[b] = 0xFFFFFFFFu;
                } else {
                    s_gbase[b] = (p.win_owner ? p.dest_off[b] : p.seg_off[b]) + pos;
                }
            };
            if (per == 1) {
                // one window per thread (up to 256 windows): the reservation's round trip to L2 runs under the
                // staging below -- its result is only needed by the write-out (the barrier stalls around this
                // block were 12 % of the kernel's stall samples)
                if (b0 < b1) {
                    def_c = s_cnt[b0];
                    s_toff[b0] = run;
                    if (def_c) {
                        def_room = p.seg_cap ? __ldg(p.seg_cap + b0) : 0xFFFFFFFFu;
                        def_pos = atomicAdd(&p.seg_fill[b0], def_c);
                    }
                }
            } else {
                for (uint32_t b = b0; b < b1; b++) {
                    const uint32_t c = s_cnt[b];
                    s_toff[b] = run;
                    run += c;
                    if (c) {
                        const uint32_t room = p.seg_cap ? __ldg(p.seg_cap + b) : 0xFFFFFFFFu;   // before the atomic's round trip
                        reserve(b, c, room, atomicAdd(&p.seg_fill[b], c));
                    }
                }
            }
        }
        __syncthreads();
        // C: stage the runs window by window
#pragma unroll
        for (int s = 0; s < 17; s++) {
            if (key[s] != 0xFFFFFFFFu) {
                const uint32_t b = key[s] >> 16, pos = s_toff[b] + (key[s] & 0xFFFFu);
                s_ent[pos] = ent[s];
                s_bid[pos] = (uint16_t)b;
            }
        }
        if (per == 1 && def_c) {                            // the deferred reservation has come back by now
            const uint32_t b = threadIdx.x;
            if (def_pos + def_c > def_room) {
                if (*reinterpret_cast<volatile uint32_t *>(p.ctl + 1) == 0u) atomicExch(p.ctl + 1, 1u);
                s_gbase[b] = 0xFFFFFFFFu;
            } else {
                s_gbase[b] = (p.win_owner ? p.dest_off[b] : p.seg_off[b]) + def_pos;
            }
        }
        __syncthreads();
        // D: contiguous runs go out with coalesced stores.  The destination is chosen OUTSIDE the loop: picking
        // p.peer[owner] out of the kernel's parameter block per entry cost 11 % of the kernel's instructions
        // also where there are no peers at all (ncu source view, profiles/r02p_ncu_scan_scatter_source.txt)
        const uint32_t total = s_total;
        if (!p.win_owner) {
            uint32_t *const pool = p.pool;
            for (uint32_t i = threadIdx.x; i < total; i += kScanThreads) {   // (four entries in flight per thread: no faster)
                const uint32_t b = s_bid[i], base = s_gbase[b];
                if (base != 0xFFFFFFFFu) pool[base + (i - s_toff[b])] = s_ent[i];
            }
        } else {
            for (uint32_t i = threadIdx.x; i < total; i += kScanThreads) {
                const uint32_t b = s_bid[i], base = s_gbase[b];
                uint32_t *const dst = s_peer[__ldg(p.win_owner + b)];
                if (base != 0xFFFFFFFFu) dst[base + (i - s_toff[b])] = s_ent[i];
            }
        }
        // the next iteration only touches s_cnt before its first barrier
    }
    if (p.pass2_tally) add_num_kmers(reinterpret_cast<unsigned long long *>(p.ctl + 4), counted);
}

// per window: every buffered entry of the window bumps its L2-resident 32-bit counter.
// The L2 atomic units are the bottleneck (~190 G/s on distinct addresses, far less on one
// address), the SMs idle: so a warp first merges its equal addresses (microsatellites put
// the same two or three k-mers into every lane) and issues one red.add per distinct one.
__device__ __forceinline__ void window_add(uint32_t *scratch, uint32_t e, bool live, uint32_t lane) {
    const uint32_t addr = live ? (e & kEntMask) : (0x80000000u | lane);    // dead lanes: unique
    uint32_t val = live ? (e >> kEntShift) + 1u : 0u;
    // cheap screen first (MATCH.ANY is slow): short-period repeats show up as an equal
    // address one, two or three lanes away
    const uint32_t a1 = __shfl_up_sync(0xFFFFFFFFu, addr, 1), a2 = __shfl_up_sync(0xFFFFFFFFu, addr, 2),
                   a3 = __shfl_up_sync(0xFFFFFFFFu, addr, 3);
    const bool dup = (lane >= 1 && a1 == addr) || (lane >= 2 && a2 == addr) || (lane >= 3 && a3 == addr);
    if (__any_sync(0xFFFFFFFFu, dup)) {
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, addr);
        if (peers != (1u << lane)) val = __reduce_add_sync(peers, val);
        if (live && lane == (uint32_t)(__ffs((int)peers) - 1)) atomicAdd(&scratch[addr], val);
    } else if (live) {
        atomicAdd(&scratch[addr], val);
    }
}

__global__ void __launch_bounds__(256) k_window_count(const uint32_t *__restrict__ pool,
                                                      const uint32_t *__restrict__ seg_off,
                                                      const uint32_t *__restrict__ seg_cnt,
                                                      int nseg, uint32_t nb, uint32_t b,
                                                      uint32_t *__restrict__ scratch) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t span = gridDim.x * blockDim.x;                 // entries per grid sweep
    constexpr int U = 4;
    for (int f = 0; f < nseg; f++) {
        const uint32_t off = seg_off[(size_t)f * nb + b], cnt = seg_cnt[(size_t)f * nb + b];
        const uint32_t *src = pool + off;
        // warp-uniform trip count: every lane of a warp runs the same iterations
        for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < cnt; base += U * span) {
            uint32_t e[U];
            bool live[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t i = base + u * span + lane;
                live[u] = i < cnt;
                e[u] = live[u] ? __ldcs(src + i) : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (base + u * span < cnt) window_add(scratch, e[u], live[u], lane);
        }
    }
}

// per window: table = min(255, [table +] counters) (indexer.py:239,262), counters back to
// zero, and -- when bins != NULL -- the histogram of the bytes just written
// (tools.py:250), so the final table is never read back.  One thread turns four
// consecutive counters (one 16-byte load, coalesced across the warp) into four table
// bytes; four independent loads are in flight per thread.
__device__ __forceinline__ uint32_t commit_quad(const uint4 c, uint32_t old, bool accum) {
    uint32_t v0 = c.x, v1 = c.y, v2 = c.z, v3 = c.w;
    if (accum) {
        v0 += old & 0xFFu; v1 += (old >> 8) & 0xFFu; v2 += (old >> 16) & 0xFFu; v3 += old >> 24;
    }
    return min(v0, 255u) | (min(v1, 255u) << 8) | (min(v2, 255u) << 16) | (min(v3, 255u) << 24);
}

template <bool ACCUM>
__global__ void __launch_bounds__(256) k_window_commit(uint32_t *__restrict__ scratch,
                                                       uint8_t *__restrict__ table, size_t n,
                                                       unsigned long long *__restrict__ bins) {
    // bins: [gridDim.x][256] partial histograms, one row per block (launches of one flush are
    // serialised on one stream, so a block owns its row); k_reduce_bins sums the rows.
    __shared__ uint32_t sh[8][256];
    if (bins) {
        for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5;
    uint32_t c1 = 0, c2 = 0, c3 = 0;
    const size_t nq = n / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint4 *sv = reinterpret_cast<uint4 *>(scratch);
    uint32_t *tw = reinterpret_cast<uint32_t *>(table);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    constexpr int U = 4;
    auto tally = [&](uint32_t x) {
        if (!bins || !x) return;
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
    };
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < nq; i += U * stride) {
        uint4 c[U];
        uint32_t old[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            c[u] = __ldcg(sv + i + u * stride);
            old[u] = ACCUM ? tw[i + u * stride] : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t x = commit_quad(c[u], old[u], ACCUM);
            if (c[u].x | c[u].y | c[u].z | c[u].w) sv[i + u * stride] = zero;   // sparse tables: mostly clean
            __stcs(tw + i + u * stride, x);
            tally(x);
        }
    }
    for (; i < nq; i += stride) {
        const uint4 c = __ldcg(sv + i);
        const uint32_t x = commit_quad(c, ACCUM ? tw[i] : 0u, ACCUM);
        if (c.x | c.y | c.z | c.w) sv[i] = zero;
        __stcs(tw + i, x);
        tally(x);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {            // < 4 tail entries
        for (size_t k = nq * 4; k < n; k++) {
            uint32_t val = scratch[k];
            scratch[k] = 0;
            if (ACCUM) val += table[k];
            val = min(val, 255u);
            table[k] = (uint8_t)val;
            if (bins && val) atomicAdd(&sh[0][val], 1u);
        }
    }
    if (!bins) return;
    if (c1) atomicAdd(&sh[warp][1], c1);
    if (c2) atomicAdd(&sh[warp][2], c2);
    if (c3) atomicAdd(&sh[warp][3], c3);
    __syncthreads();
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += sh[k][threadIdx.x];
    if (s && threadIdx.x) bins[(size_t)blockIdx.x * 256 + threadIdx.x] += s;
}

__global__ void __launch_bounds__(256) k_reduce_bins(const unsigned long long *__restrict__ part,
                                                     int rows, unsigned long long *__restrict__ bins) {
    unsigned long long s = 0;
    for (int r = 0; r < rows; r++) s += part[(size_t)r * 256 + threadIdx.x];
    bins[threadIdx.x] = s;
}

// ------------------------------------------------------------------------------ packed transfer
// A finished table slice on its way to the host (pk_indexer_finalize_to_host): one bit per entry, the
// non-zero bytes in entry order, and per chunk of 1024 entries where its bytes start -- the format
// csrc/unpack.cpp rebuilds on the host's cores.  One warp per chunk: two coalesced 512-byte loads, the
// lane's non-zero bytes compacted into shared memory at the lane's rank (two warp scans), one atomicAdd
// reserves the chunk's room (16-byte units, so the copy out is whole uint4s; chunks land in any order),
// 128 bytes of bitmap.  Reads 1 B per entry (the slice was just written: mostly L2 hits), writes
// 1/8 + fill B per entry.
__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t w) {     // bit i <=> byte i of w is not zero
    return ((__vcmpne4(w, 0u) & 0x08040201u) * 0x01010101u) >> 24;
}
__device__ __forceinline__ uint32_t nonzero16(const uint4 v) {
    return nonzero_nibble(v.x) | (nonzero_nibble(v.y) << 4) | (nonzero_nibble(v.z) << 8) | (nonzero_nibble(v.w) << 12);
}
__device__ __forceinline__ void compact16(const uint4 v, uint8_t *dst) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const uint32_t x = (w[k] >> (8 * b)) & 0xFFu;
            if (x) *dst++ = (uint8_t)x;
        }
}

__global__ void __launch_bounds__(256) k_table_pack(const uint8_t *__restrict__ table, size_t nchunks,
                                                    uint16_t *__restrict__ bitmap16, uint32_t *__restrict__ chunk_off,
                                                    uint8_t *__restrict__ nz, uint32_t *__restrict__ cursor) {
    __shared__ __align__(16) uint8_t stage[8][kPackChunk + 16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *sb = stage[warp];
    for (size_t c = (size_t)blockIdx.x * 8 + warp; c < nchunks; c += (size_t)gridDim.x * 8) {
        const uint4 *src = reinterpret_cast<const uint4 *>(table + c * kPackChunk);
        const uint4 a = __ldcs(src + lane), b = __ldcs(src + 32 + lane);
        const uint32_t ma = nonzero16(a), mb = nonzero16(b);
        const uint32_t ca = __popc(ma), cb = __popc(mb);
        uint32_t pa = ca, pb = cb;                          // inclusive scans over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t ua = __shfl_up_sync(0xFFFFFFFFu, pa, d), ub = __shfl_up_sync(0xFFFFFFFFu, pb, d);
            if (lane >= d) { pa += ua; pb += ub; }
        }
        const uint32_t tot_a = __shfl_sync(0xFFFFFFFFu, pa, 31), total = tot_a + __shfl_sync(0xFFFFFFFFu, pb, 31);
        if (ca) compact16(a, sb + (pa - ca));
        if (cb) compact16(b, sb + tot_a + (pb - cb));
        const uint32_t units = (total + 15u) >> 4;
        if (lane < 16 && total + lane < units * 16u) sb[total + lane] = 0;      // padding of the last unit
        uint32_t off = 0;
        if (lane == 0) {
            off = atomicAdd(cursor, units);
            chunk_off[c] = off;
        }
        off = __shfl_sync(0xFFFFFFFFu, off, 0);
        __syncwarp();
        uint4 *dst = reinterpret_cast<uint4 *>(nz) + off;
        for (uint32_t i = lane; i < units; i += 32) dst[i] = reinterpret_cast<const uint4 *>(sb)[i];
        bitmap16[c * (kPackChunk / 16) + lane] = (uint16_t)ma;
        bitmap16[c * (kPackChunk / 16) + 32 + lane] = (uint16_t)mb;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------ byte windows
// Sparse tables (K >= 17: a few k-mers per hundred entries) pay for the 32-bit counters twice:
// a window covers only 2^24 entries, and every commit reads 4 bytes to write one.  Here the
// L2-resident window holds the table's own 8-bit lanes, four to a word, so 64 MiB of L2 cover
// 2^26 entries.  atomicAdd(word, cnt << 8*lane) is not saturating, so the add RETURNS the old
// word (ATOMG, measured 128 G/s against 190 G/s for RED): the thread whose add carried out of a
// lane sees it, and books the exact correction
//     true(lane) = physical(lane) + 256 * carries_out(lane) - carries_in(lane)
// as +256 / -1 deltas in a small hash table keyed by lane.  The last block to finish a
// window's count applies the deltas -- lane = min(255, true).  Carries are rare (a k-mer must
// pass 255 inside one window); if the hash table ever fills up, the last block recounts the
// whole window with the exact compare-and-swap rule.
//
// First flush of a table (the normal case): the window IS the table slice -- k_window_zero
// puts it into L2 as zeros, k_window_count8<true> counts in place, and since every returned old
// word tells exactly how each lane moved, the histogram (tools.py:250) is kept as transitions
// old -> new (bins[old]--, bins[new]++): no commit pass, no statistics pass, the table is
// written to DRAM once when L2 evicts it.  Later flushes onto a table that already holds
// counts go through a zeroed scratch window and k_window_commit8<true> (saturating byte add).
struct OvfTable {
    uint32_t *keys;                 // [cap] lane index inside the window, 0xFFFFFFFF = empty
    unsigned long long *vals;       // [cap] signed delta (two's complement)
    uint32_t *list;                 // [cap] occupied slots in insertion order
    uint32_t *meta;                 // [0] occupied slots, [1] finished blocks, [2] table overflowed
    uint32_t mask;                  // cap - 1
};

__device__ __forceinline__ void ovf_add(const OvfTable &t, uint32_t lane_idx, long long delta) {
    if (*reinterpret_cast<volatile uint32_t *>(t.meta + 2)) return;   // already bound for the exact recount
    uint32_t h = (lane_idx * 0x9E3779B1u) >> 7;
    const uint32_t max_probe = min(t.mask, 255u);          // a long probe chain counts as a full table
    for (uint32_t probe = 0; probe <= max_probe; probe++, h++) {
        const uint32_t slot = h & t.mask;
        uint32_t k = __ldcg(t.keys + slot);
        if (k == 0xFFFFFFFFu) {
            k = atomicCAS(t.keys + slot, 0xFFFFFFFFu, lane_idx);
            if (k == 0xFFFFFFFFu) {
                t.list[atomicAdd(t.meta, 1u)] = slot;
                k = lane_idx;
            }
        }
        if (k == lane_idx) {
            atomicAdd(t.vals + slot, (unsigned long long)delta);
            return;
        }
    }
    atomicExch(t.meta + 2, 1u);                           // full: the window is recounted exactly
}

// one lane moved from `from` to `to`: histogram bookkeeping in global memory (rare paths)
__device__ __forceinline__ void bins_move(unsigned long long *g_bins, uint32_t from, uint32_t to) {
    if (from == to) return;
    if (from) atomicAdd(g_bins + from, ~0ull);            // -1
    if (to) atomicAdd(g_bins + to, 1ull);
}

// the add `val << 8*(lane_idx & 3)` onto `old` carried out of its lane: book every carry of the
// ripple (and, with g_bins, every lane the ripple moved)
__device__ __noinline__ void ovf_book(const OvfTable &t, uint32_t old, uint32_t lane_idx, uint32_t val,
                                      unsigned long long *g_bins) {
    const uint32_t first = lane_idx & 3u, base = lane_idx & ~3u;
    uint32_t add = val;
    for (uint32_t j = first; j < 4 && add; j++) {
        const uint32_t was = (old >> (8 * j)) & 0xFFu;
        const uint32_t s = was + add;
        if (g_bins) bins_move(g_bins, was, s & 0xFFu);
        add = s >> 8;                                     // carry into the next lane
        if (add) {
            ovf_add(t, base + j, 256);
            if (j < 3) ovf_add(t, base + j + 1, -1);      // a carry out of lane 3 leaves the word
        }
    }
}

template <bool HIST>
__device__ __forceinline__ void lane_add(uint32_t *win, uint32_t lane_idx, uint32_t val, const OvfTable &t,
                                         HistTally &ht) {
    const uint32_t sh = 8u * (lane_idx & 3u);
    const uint32_t old = atomicAdd(win + (lane_idx >> 2), val << sh);
    const uint32_t was = (old >> sh) & 0xFFu;
    if (was + val > 255u) ovf_book(t, old, lane_idx, val, HIST ? ht.g : nullptr);
    else if (HIST) ht.move(was, was + val);
}

// same duplicate merging as window_add; a merged count beyond 255 saturates the lane anyway
template <bool HIST>
__device__ __forceinline__ void window_add8(uint32_t *win, uint32_t e, bool live, uint32_t lane,
                                            const OvfTable &t, HistTally &ht) {
    const uint32_t addr = live ? (e & kEntMask) : (0x80000000u | lane);
    uint32_t val = live ? (e >> kEntShift) + 1u : 0u;
    const uint32_t a1 = __shfl_up_sync(0xFFFFFFFFu, addr, 1), a2 = __shfl_up_sync(0xFFFFFFFFu, addr, 2),
                   a3 = __shfl_up_sync(0xFFFFFFFFu, addr, 3);
    const bool dup = (lane >= 1 && a1 == addr) || (lane >= 2 && a2 == addr) || (lane >= 3 && a3 == addr);
    if (__any_sync(0xFFFFFFFFu, dup)) {
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, addr);
        if (peers != (1u << lane)) val = __reduce_add_sync(peers, val);
        if (live && lane == (uint32_t)(__ffs((int)peers) - 1)) lane_add<HIST>(win, addr, min(val, 255u), t, ht);
    } else if (live) {
        lane_add<HIST>(win, addr, val, t, ht);
    }
}

// exact saturating add on a packed lane (compare-and-swap; a read of 255 is final) for the recount
template <bool HIST>
__device__ __forceinline__ void lane_add_exact(uint32_t *win, uint32_t lane_idx, uint32_t val, HistTally &ht) {
    uint32_t *wp = win + (lane_idx >> 2);
    const uint32_t sh = 8u * (lane_idx & 3u);
    uint32_t old = __ldcg(wp);
    for (;;) {
        const uint32_t b = (old >> sh) & 0xFFu;
        if (b == 255u) return;
        const uint32_t nb = min(255u, b + val);
        const uint32_t assumed = old;
        old = atomicCAS(wp, assumed, (assumed & ~(0xFFu << sh)) | (nb << sh));
        if (old == assumed) {
            if (HIST) ht.move(b, nb);
            return;
        }
    }
}

__global__ void __launch_bounds__(256) k_window_zero(uint4 *__restrict__ dst, size_t nvec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) dst[i] = zero;
}

// HIST = false: `win` is the zeroed scratch window (k_window_commit8 follows).
// HIST = true:  `win` is the zeroed table slice itself; g_bins follows every lane.
// win_words = 32-bit words of the window (a multiple of 4; the allocation is padded).
template <bool HIST>
__global__ void __launch_bounds__(256) k_window_count8(const uint32_t *__restrict__ pool,
                                                       const uint32_t *__restrict__ seg_off,
                                                       const uint32_t *__restrict__ seg_cnt,
                                                       int nseg, uint32_t nb, uint32_t b,
                                                       uint32_t *__restrict__ win, size_t win_words,
                                                       const OvfTable t, unsigned long long *__restrict__ g_bins) {
    __shared__ uint32_t s_bins[256];
    __shared__ uint32_t s_last;
    HistTally ht{s_bins, g_bins, 0u};
    if (HIST) {
        s_bins[threadIdx.x] = 0;
        __syncthreads();
    }
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t span = gridDim.x * blockDim.x;
    constexpr int U = 4;
    for (int f = 0; f < nseg; f++) {
        const uint32_t off = seg_off[(size_t)f * nb + b], cnt = seg_cnt[(size_t)f * nb + b];
        const uint32_t *src = pool + off;
        for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < cnt; base += U * span) {
            uint32_t e[U];
            bool live[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t i = base + u * span + lane;
                live[u] = i < cnt;
                e[u] = live[u] ? __ldcs(src + i) : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (base + u * span < cnt) window_add8<HIST>(win, e[u], live[u], lane, t, ht);
        }
    }
    if (HIST) ht.flush();
    // the last block to get here settles the carries
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(t.meta + 1, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const uint32_t nkeys = __ldcg(t.meta), failed = __ldcg(t.meta + 2);
    if (failed) {
        // exact recount by this one block (slow; only when > cap lanes overflowed in one window)
        uint4 *z = reinterpret_cast<uint4 *>(win);
        for (size_t i = threadIdx.x; i < win_words / 4; i += blockDim.x) {
            if (HIST) {                                   // take back what the lanes stood for
                const uint4 q = __ldcg(z + i);
                const uint32_t ws[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (ws[k])
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint32_t v = (ws[k] >> (8 * j)) & 0xFFu;
                            if (v) atomicSub(s_bins + v, 1u);
                        }
            }
            z[i] = make_uint4(0, 0, 0, 0);
        }
        for (uint32_t i = threadIdx.x; i <= t.mask; i += blockDim.x) { t.keys[i] = 0xFFFFFFFFu; t.vals[i] = 0ull; }
        __threadfence();
        __syncthreads();
        for (int f = 0; f < nseg; f++) {
            const uint32_t off = seg_off[(size_t)f * nb + b], cnt = seg_cnt[(size_t)f * nb + b];
            for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
                const uint32_t e = __ldcs(pool + off + i);
                lane_add_exact<HIST>(win, e & kEntMask, (e >> kEntShift) + 1u, ht);
            }
        }
        if (HIST) ht.flush();
    } else {
        for (uint32_t i = threadIdx.x; i < nkeys; i += blockDim.x) {
            const uint32_t slot = __ldcg(t.list + i);
            const uint32_t lane_idx = __ldcg(t.keys + slot);
            const long long delta = (long long)__ldcg(t.vals + slot);
            const uint32_t sh = 8u * (lane_idx & 3u);
            const long long phys = (long long)((__ldcg(win + (lane_idx >> 2)) >> sh) & 0xFFu);
            long long want = phys + delta;                 // the lane's true count, >= 0
            want = want > 255 ? 255 : want;
            atomicAdd(win + (lane_idx >> 2), (uint32_t)(int32_t)(want - phys) << sh);
            if (HIST) bins_move(g_bins, (uint32_t)phys, (uint32_t)want);
            t.keys[slot] = 0xFFFFFFFFu;
            t.vals[slot] = 0ull;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { t.meta[0] = 0; t.meta[1] = 0; t.meta[2] = 0; }
}

// per window: table = [table +sat] lanes (indexer.py:239,262), lanes back to zero, histogram of the
// bytes written (tools.py:250).  One thread moves 16 table entries per step.
template <bool ACCUM>
__global__ void __launch_bounds__(256) k_window_commit8(uint32_t *__restrict__ scratch,
                                                        uint8_t *__restrict__ table, size_t n,
                                                        unsigned long long *__restrict__ bins) {
    __shared__ uint32_t sh[8][256];
    if (bins) {
        for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5;
    uint32_t c1 = 0, c2 = 0, c3 = 0;
    auto tally = [&](uint32_t x) {
        if (!bins || !x) return;
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
    };
    const size_t nv = n / 16;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    uint4 *sv = reinterpret_cast<uint4 *>(scratch);
    uint4 *tv = reinterpret_cast<uint4 *>(table);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    constexpr int U = 4;
    auto one = [&](size_t i, uint4 c, uint4 old) {
        if (c.x | c.y | c.z | c.w) sv[i] = zero;            // sparse tables: mostly clean already
        if (ACCUM) {
            c.x = __vaddus4(c.x, old.x); c.y = __vaddus4(c.y, old.y);
            c.z = __vaddus4(c.z, old.z); c.w = __vaddus4(c.w, old.w);
        }
        __stcs(tv + i, c);
        tally(c.x); tally(c.y); tally(c.z); tally(c.w);
    };
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < nv; i += U * stride) {
        uint4 c[U], old[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            c[u] = __ldcg(sv + i + u * stride);
            old[u] = ACCUM ? __ldcs(tv + i + u * stride) : zero;
        }
#pragma unroll
        for (int u = 0; u < U; u++) one(i + u * stride, c[u], old[u]);
    }
    for (; i < nv; i += stride) one(i, __ldcg(sv + i), ACCUM ? __ldcs(tv + i) : zero);
    if (blockIdx.x == 0 && threadIdx.x == 0) {             // < 16 tail entries
        uint8_t *sb = reinterpret_cast<uint8_t *>(scratch);
        for (size_t k = nv * 16; k < n; k++) {
            uint32_t val = sb[k];
            sb[k] = 0;
            if (ACCUM) val = min(255u, val + table[k]);
            table[k] = (uint8_t)val;
            if (bins && val) atomicAdd(&sh[0][val], 1u);
        }
    }
    if (!bins) return;
    if (c1) atomicAdd(&sh[warp][1], c1);
    if (c2) atomicAdd(&sh[warp][2], c2);
    if (c3) atomicAdd(&sh[warp][3], c3);
    __syncthreads();
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += sh[k][threadIdx.x];
    if (s && threadIdx.x) bins[(size_t)blockIdx.x * 256 + threadIdx.x] += s;
}

// ------------------------------------------------------------------------------ smem flush
// Second level of the partition (dense tables): the entries of every 2^24 window are split
// once more into 512 sub-buckets of 2^15 table entries, whose 32-bit counters fit in the
// shared memory of one CTA (128 KB).  Counting then runs on shared-memory atomics (~1 T/s
// measured) instead of L2 atomics (~0.16 T/s), and one CTA turns its counters straight into
// 32 KB of table bytes + histogram: no counter array in L2, no separate commit pass.
constexpr int kSubLog2 = 15;                    // table entries per sub-bucket
constexpr int kSubs = 1 << (24 - kSubLog2);     // sub-buckets per 2^24 window (512)
constexpr int kSubTile = 8192;                  // entries one CTA ranks per step in k_sub_scatter
constexpr int kSubThreads = 512;

// entries of window (w_first + blockIdx.y) -> per-sub-bucket counts
__global__ void __launch_bounds__(256) k_sub_count(const uint32_t *__restrict__ pool,
                                                   const uint32_t *__restrict__ seg_off,
                                                   const uint32_t *__restrict__ seg_cnt, int nseg,
                                                   uint32_t nb, uint32_t w_first, uint32_t win_log2,
                                                   uint32_t *__restrict__ sub_cnt) {
    __shared__ uint32_t s_cnt[kSubs];
    for (int i = threadIdx.x; i < kSubs; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    const uint32_t b = w_first + blockIdx.y;
    const uint32_t wmask = (1u << win_log2) - 1u;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (int f = 0; f < nseg; f++) {
        const uint32_t off = seg_off[(size_t)f * nb + b], cnt = seg_cnt[(size_t)f * nb + b];
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += stride)
            atomicAdd(&s_cnt[(__ldg(pool + off + i) & wmask) >> kSubLog2], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSubs; i += blockDim.x)
        if (s_cnt[i]) atomicAdd(&sub_cnt[(size_t)blockIdx.y * kSubs + i], s_cnt[i]);
}

// exclusive scan over nwin * 512 sub-bucket counts -> offsets in the second pool
__global__ void __launch_bounds__(1024) k_sub_offsets(const uint32_t *__restrict__ cnt,
                                                      uint32_t *__restrict__ off, uint32_t n) {
    __shared__ uint32_t part[1024];
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t b0 = threadIdx.x * per, b1 = min(n, b0 + per);
    uint32_t s = 0;
    for (uint32_t b = b0; b < b1; b++) s += cnt[b];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t v[32], sum = 0;
#pragma unroll
        for (int i = 0; i < 32; i++) { v[i] = part[threadIdx.x * 32 + i]; sum += v[i]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if ((int)threadIdx.x >= o) incl += up;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int i = 0; i < 32; i++) { part[threadIdx.x * 32 + i] = run; run += v[i]; }
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (uint32_t b = b0; b < b1; b++) { off[b] = run; run += cnt[b]; }
}

// entries of window (w_first + blockIdx.y) -> second pool, grouped by sub-bucket
__global__ void __launch_bounds__(kSubThreads) k_sub_scatter(const uint32_t *__restrict__ pool,
                                                             const uint32_t *__restrict__ seg_off,
                                                             const uint32_t *__restrict__ seg_cnt, int nseg,
                                                             uint32_t nb, uint32_t w_first, uint32_t win_log2,
                                                             const uint32_t *__restrict__ sub_off,
                                                             uint32_t *__restrict__ sub_fill,
                                                             uint32_t *__restrict__ pool2) {
    extern __shared__ uint32_t sm2[];
    uint32_t *s_cnt = sm2;                         // [kSubs]
    uint32_t *s_toff = sm2 + kSubs;                // [kSubs]
    uint32_t *s_gbase = sm2 + 2 * kSubs;           // [kSubs]
    uint32_t *s_ent = sm2 + 3 * kSubs;             // [kSubTile]
    uint16_t *s_sid = reinterpret_cast<uint16_t *>(s_ent + kSubTile);   // [kSubTile]
    __shared__ uint32_t s_total;
    const uint32_t b = w_first + blockIdx.y;
    const uint32_t wmask = (1u << win_log2) - 1u;
    const size_t sbase = (size_t)blockIdx.y * kSubs;
    constexpr int kPer = kSubTile / kSubThreads;   // 16 entries per thread
    for (int f = 0; f < nseg; f++) {
        const uint32_t off = seg_off[(size_t)f * nb + b], cnt = seg_cnt[(size_t)f * nb + b];
        const uint32_t ntiles = (cnt + kSubTile - 1) / kSubTile;
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            __syncthreads();                                  // previous tile fully written out
            for (int i = threadIdx.x; i < kSubs; i += blockDim.x) s_cnt[i] = 0;
            __syncthreads();
            const uint32_t t0 = tile * kSubTile;
            uint32_t ent[kPer], key[kPer];
#pragma unroll
            for (int k = 0; k < kPer; k++) {
                const uint32_t i = t0 + k * kSubThreads + threadIdx.x;      // coalesced
                key[k] = 0xFFFFFFFFu;
                if (i < cnt) {
                    ent[k] = __ldcs(pool + off + i);
                    const uint32_t sub = (ent[k] & wmask) >> kSubLog2;
                    key[k] = (sub << 16) | atomicAdd(&s_cnt[sub], 1u);
                }
            }
            __syncthreads();
            if (threadIdx.x < 32) {                           // scan the 512 counts with one warp
                uint32_t v[kSubs / 32], sum = 0;
#pragma unroll
                for (int i = 0; i < kSubs / 32; i++) { v[i] = s_cnt[threadIdx.x * (kSubs / 32) + i]; sum += v[i]; }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if ((int)threadIdx.x >= o) incl += up;
                }
                uint32_t run = incl - sum;
#pragma unroll
                for (int i = 0; i < kSubs / 32; i++) { s_toff[threadIdx.x * (kSubs / 32) + i] = run; run += v[i]; }
                if (threadIdx.x == 31) s_total = incl;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < kSubs; i += blockDim.x) {
                const uint32_t c = s_cnt[i];
                if (c) s_gbase[i] = sub_off[sbase + i] + atomicAdd(&sub_fill[sbase + i], c);
            }
#pragma unroll
            for (int k = 0; k < kPer; k++) {
                if (key[k] != 0xFFFFFFFFu) {
                    const uint32_t sub = key[k] >> 16, pos = s_toff[sub] + (key[k] & 0xFFFFu);
                    s_ent[pos] = ent[k];
                    s_sid[pos] = (uint16_t)sub;
                }
            }
            __syncthreads();
            const uint32_t total = s_total;
            for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
                const uint32_t sub = s_sid[i];
                pool2[s_gbase[sub] + (i - s_toff[sub])] = s_ent[i];
            }
        }
    }
}

// one CTA per sub-bucket of window w: count in shared memory, write 32 KB of the table
template <bool ACCUM>
__global__ void __launch_bounds__(1024, 1) k_sub_tally(const uint32_t *__restrict__ pool2,
                                                       const uint32_t *__restrict__ sub_off,
                                                       const uint32_t *__restrict__ sub_cnt,
                                                       uint32_t w_local, uint8_t *__restrict__ table_win,
                                                       size_t n_win, unsigned long long *__restrict__ bins) {
    extern __shared__ uint32_t sm3[];
    uint32_t *cnt = sm3;                                     // [2^15]
    uint32_t(*sh)[256] = reinterpret_cast<uint32_t(*)[256]>(sm3 + (1 << kSubLog2));   // [8][256]
    const uint32_t sub = blockIdx.x;
    const size_t first = (size_t)sub << kSubLog2;            // first table entry of this sub-bucket
    if (first >= n_win) return;
    const uint32_t n_here = (uint32_t)min((size_t)(1u << kSubLog2), n_win - first);
    {
        uint4 *z = reinterpret_cast<uint4 *>(cnt);
        for (int i = threadIdx.x; i < (1 << kSubLog2) / 4; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        if (bins) for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
    }
    __syncthreads();
    const size_t si = (size_t)w_local * kSubs + sub;
    const uint32_t off = sub_off[si], m = sub_cnt[si];
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        const uint32_t e = __ldcs(pool2 + off + i);
        atomicAdd(&cnt[e & ((1u << kSubLog2) - 1u)], (e >> kEntShift) + 1u);
    }
    __syncthreads();
    const int warp = (threadIdx.x >> 5) & 7;
    uint32_t c1 = 0, c2 = 0, c3 = 0;
    uint32_t *tw = reinterpret_cast<uint32_t *>(table_win + first);
    const uint32_t nq = n_here / 4;
    for (uint32_t q = threadIdx.x; q < nq; q += blockDim.x) {
        const uint4 c = *reinterpret_cast<const uint4 *>(cnt + 4 * q);
        const uint32_t x = commit_quad(c, ACCUM ? tw[q] : 0u, ACCUM);
        __stcs(tw + q, x);
        if (bins && x) {
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
    if (threadIdx.x == 0) {                                  // < 4 tail entries
        for (uint32_t k = nq * 4; k < n_here; k++) {
            uint32_t val = cnt[k];
            if (ACCUM) val += table_win[first + k];
            val = min(val, 255u);
            table_win[first + k] = (uint8_t)val;
            if (bins && val) atomicAdd(&sh[0][val], 1u);
        }
    }
    if (!bins) return;
    if (c1) atomicAdd(&sh[warp][1], c1);
    if (c2) atomicAdd(&sh[warp][2], c2);
    if (c3) atomicAdd(&sh[warp][3], c3);
    __syncthreads();
    if (threadIdx.x < 256) {
        unsigned long long s = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) s += sh[k][threadIdx.x];
        if (s && threadIdx.x) bins[(size_t)blockIdx.x * 256 + threadIdx.x] += s;
    }
}

// new carry = last kCarry bytes of (old carry ++ seq[0..n))
__global__ void k_update_carry(uint8_t *carry, const uint8_t *seq, size_t n, uint32_t *ctl) {
    const int i = threadIdx.x;                            // 32 threads
    if (ctl && i == 0) ctl[1] = 0u;                       // the feed is settled either way
    const long long pos = (long long)n - kCarry + i;
    const uint8_t v = pos >= 0 ? seq[pos] : carry[kCarry + pos];
    __syncwarp();
    carry[i] = v;
}

// 256-bin histogram of a byte table.  Zero bytes are never counted (derived on
// the host from n); 1, 2 and 3 -- the bulk of a k-mer table -- stay in registers.
__global__ void __launch_bounds__(256) k_table_stats(const uint8_t *__restrict__ table, size_t n,
                                                     unsigned long long *__restrict__ hist) {
    __shared__ uint32_t sh[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    uint32_t c1 = 0, c2 = 0, c3 = 0;
    const size_t nvec = n / 16;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint4 *v = reinterpret_cast<const uint4 *>(table);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const uint4 q = __ldcs(v + i);
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
    if (blockIdx.x == 0 && threadIdx.x == 0) {            // < 16 tail bytes
        for (size_t i = nvec * 16; i < n; i++) {
            const uint32_t val = table[i];
            if (val) atomicAdd(&sh[0][val], 1u);
        }
    }
    if (c1) atomicAdd(&sh[warp][1], c1);
    if (c2) atomicAdd(&sh[warp][2], c2);
    if (c3) atomicAdd(&sh[warp][3], c3);
    __syncthreads();
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += sh[k][threadIdx.x];
    if (s && threadIdx.x) atomicAdd(&hist[threadIdx.x], s);
}

int stats_from_bins(const unsigned long long bins_in[256], size_t n, int64_t hist[255],
                    uint64_t stats[4]) {
    unsigned long long nz = 0, sum = 0;
    int mx = 0, mn_nz = 0;
    for (int v = 1; v < 256; v++) {
        hist[v - 1] = (int64_t)bins_in[v];                // tools.py:250 hist[i] = #{== i+1}
        if (bins_in[v]) {
            nz += bins_in[v];
            sum += bins_in[v] * (unsigned long long)v;
            mx = v;
            if (!mn_nz) mn_nz = v;
        }
    }
    stats[0] = sum;                                       // vals_sum   tools.py:260
    stats[1] = nz;                                        // vals_count tools.py:261
    stats[2] = (nz < n) ? 0 : (uint64_t)mn_nz;            // vals_min   tools.py:262
    stats[3] = (uint64_t)mx;                              // vals_max   tools.py:263
    return PK_OK;
}

size_t scatter_smem_bytes(uint32_t nb) {
    return (size_t)3 * nb * sizeof(uint32_t) + (size_t)kTileEntries * (sizeof(uint32_t) + sizeof(uint16_t));
}

}  // namespace

static int g_persist_users[64];               // handles per device holding a persisting-L2 carve-out

struct pk_indexer {
    int K = 0, device = 0, mode = PK_MODE_DIRECT;
    uint64_t lo = 0, hi = 0;
    size_t table_bytes = 0;
    uint8_t *table = nullptr;
    uint8_t *carry = nullptr;                  // kCarry bytes
    unsigned long long *counters = nullptr;    // [0] num_kmers, [1..256] histogram bins
    unsigned long long *h_counters = nullptr;  // pinned mirror
    uint64_t stream_off = 0;
    uint64_t *rec_starts = nullptr;
    uint8_t *rec_flags = nullptr;
    size_t nrec = 0, rec_cap = 0;
    uint64_t last_rec_start = 0;
    uint8_t *stage[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr, work_stream = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    cudaEvent_t joined = nullptr;              // orders the caller's stream against work_stream
    cudaStream_t last_stream = nullptr;        // stream of the most recent feed / reset
    int sm_count = 148;
    uint64_t launches = 0;
    bool fed = false;
    // PARTITION mode
    uint32_t win_log2 = 24, nbuckets = 0;
    uint32_t *pool = nullptr;                  // buffered entries
    const uint32_t *pool_ext = nullptr;        // entries imported from other ranks (pk_indexer_import_segments)
    // fused exchange: peer-mapped pools of the window owners and the per-window routing
    uint32_t *peer_pool[16] = {nullptr};
    void *peer_ipc[16] = {nullptr};            // cudaIpcOpenMemHandle bases to close
    uint32_t *route = nullptr;                 // device: [3][kMaxBuckets] owner, destination offset, room (routed scan)
    bool routed = false;                       // pk_indexer_set_route has been called
    uint32_t *routed_status = nullptr;         // caller's device word receiving the overflow flag of a routed scan
    int self_rank = 0;
    unsigned long long pub_base[16] = {0};     // entry index of the published-count table in each owner's buffer
    uint32_t *import_off = nullptr;            // owner: device [import_nseg][nbuckets] layout of what peers store here
    uint32_t import_nseg = 0, import_w0 = 0, import_nb_total = 0;
    const uint8_t *p1_seq = nullptr;           // sequence of the pending pass 1
    size_t p1_n = 0;
    size_t pool_cap = 0, pool_ub = 0;          // capacity / upper bound of entries in use
    uint32_t *seg = nullptr;                   // 4 x [kMaxSegments][nbuckets]: cnt, off, fill, cap
    uint32_t *cursor = nullptr;                // device control words: pool cursor, overflow flag, ... (ScanParams::ctl)
    uint32_t est_shift = 4;                    // estimated pass 1 looks at one tile in 2^est_shift (0 = always exact)
    size_t est_min = (size_t)1 << 24;          // ... for feeds of at least this many bases
    uint32_t est_test = 0;                     // test hook: 1 = as if over budget, 2 = reserve too little room
    uint32_t *scratch = nullptr;               // one window of 32-bit counters
    unsigned long long *bins_part = nullptr;   // [8 * sm_count][256] partial histograms
    uint32_t *pool2 = nullptr;                 // smem flush: entries regrouped by sub-bucket
    uint32_t *sub = nullptr;                   // smem flush: 3 x [64 * 512] counts, offsets, cursors
    bool flush_smem = false;                   // second-level shared-memory flush (else L2 counters)
    bool count8 = false;                       // byte windows: the L2 window holds 8-bit lanes (k_window_count8)
    int cnt8_blocks_per_sm = 4;                // grid of k_window_count8: K=17 step 21.1 / 18.7 / 20.1 / 20.3 ms at 2 / 4 / 8 / 16
    int commit_blocks_per_sm = 4, zero_blocks_per_sm = 4;   // k_window_commit (K=15: 1.68 / 1.47 / 1.56 ms at 2 / 4 / 8) and k_window_zero grids
    int cnt_blocks_per_sm = 12;                // grid of k_window_count: K=15 step 9.86 / 9.77 / 9.43 / 9.45 ms at 6 / 8 / 12 / 16
    size_t scratch_bytes = 0;
    OvfTable ovf = {nullptr, nullptr, nullptr, nullptr, 0};
    bool sub_smem_set = false;
    cudaEvent_t committed[2] = {nullptr, nullptr};
    struct TableShipper *shipper = nullptr;    // packed device-to-host transfer of finished table slices
    uint64_t xfer[4] = {0, 0, 0, 0};           // last finalize_to_host: bytes device-to-host, packed / raw slices, unpack threads
    int nseg = 0;
    size_t l2_persist_bytes = 0;               // persisting-L2 carve-out granted for `scratch`
    bool table_valid = false;                  // every window has been written since reset
    bool stats_valid = false;                  // bins hold the histogram of the current table
    bool scatter_smem_set = false;
    // optional per-kernel-class timing (pk_indexer_set_profiling)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;      // pairs (begin, end)
    std::vector<int> prof_tags;
};

enum { PROF_SCAN_DIRECT = 0, PROF_BUCKET_COUNT, PROF_OFFSETS, PROF_SCATTER, PROF_WINDOW_COUNT,
       PROF_WINDOW_COMMIT, PROF_TABLE_STATS, PROF_CARRY, PROF_CLASSES };

// brackets one kernel launch with events on its stream when profiling is on
struct prof_scope {
    pk_indexer *ix; cudaStream_t st; cudaEvent_t end = nullptr;
    prof_scope(pk_indexer *ix_, cudaStream_t st_, int tag) : ix(ix_), st(st_) {
        if (!ix->profiling) return;
        cudaEvent_t b = nullptr;
        if (cudaEventCreate(&b) != cudaSuccess || cudaEventCreate(&end) != cudaSuccess) { end = nullptr; return; }
        cudaEventRecord(b, st);
        ix->prof_events.push_back(b);
        ix->prof_events.push_back(end);
        ix->prof_tags.push_back(tag);
    }
    ~prof_scope() { if (end) cudaEventRecord(end, st); }
};

// zero the whole table in pieces of 16 GiB: one cudaMemsetAsync over tens of GiB slows down to
// ~2.4 TB/s (measured: 51 GiB in 22 ms, against 32 GiB in 4.8 ms = 7.2 TB/s)
static cudaError_t table_zero(pk_indexer *ix, cudaStream_t st) {
    const size_t bytes = (ix->table_bytes + 255) & ~(size_t)255, piece = (size_t)16 << 30;
    for (size_t off = 0; off < bytes; off += piece) {
        const cudaError_t e = cudaMemsetAsync(ix->table + off, 0, std::min(piece, bytes - off), st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// make work_stream wait for whatever the caller's stream was last given
static int indexer_join(pk_indexer *ix) {
    if (ix->last_stream != ix->work_stream) {
        PK_CUDA(cudaEventRecord(ix->joined, ix->last_stream));
        PK_CUDA(cudaStreamWaitEvent(ix->work_stream, ix->joined, 0));
        ix->last_stream = ix->work_stream;
    }
    return PK_OK;
}

static uint32_t *seg_cnt(pk_indexer *ix, int f) { return ix->seg + (size_t)f * ix->nbuckets; }
static uint32_t *seg_off(pk_indexer *ix, int f) {
    return ix->seg + ((size_t)kMaxSegments + f) * ix->nbuckets;
}
static uint32_t *seg_fill(pk_indexer *ix, int f) {
    return ix->seg + ((size_t)2 * kMaxSegments + f) * ix->nbuckets;
}
static uint32_t *seg_cap(pk_indexer *ix, int f) {
    return ix->seg + ((size_t)3 * kMaxSegments + f) * ix->nbuckets;
}

// The window's 32-bit counters must stay in L2 while the k-mer entries and the table
// stream past them: mark them persisting (and everything that misses the carve-out
// streaming) for the two window kernels.
static unsigned window_launch_attr(pk_indexer *ix, cudaLaunchAttribute *attr) {
    if (!ix->l2_persist_bytes) return 0;
    const size_t bytes = ix->scratch_bytes;
    attr->id = cudaLaunchAttributeAccessPolicyWindow;
    attr->val.accessPolicyWindow.base_ptr = ix->scratch;
    attr->val.accessPolicyWindow.num_bytes = bytes;
    attr->val.accessPolicyWindow.hitRatio =
        ix->l2_persist_bytes >= bytes ? 1.0f : (float)ix->l2_persist_bytes / (float)bytes;
    attr->val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr->val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    return 1;
}

// ---- packed device-to-host transfer of a finished table (pk_indexer_finalize_to_host) ---------------
// The table is the largest thing an indexing job moves over PCIe and it is mostly zeros, so every
// finished slice (a window) is packed on the device (k_table_pack), only bitmap + chunk offsets +
// non-zero bytes cross the bus into a pinned slot, and a team of host threads rebuilds the bytes in the
// caller's buffer (unpack.cpp) while the next windows are still being counted and copied.  The caller's
// thread runs kLag windows behind the launches: it waits for a window's packed size (4 bytes the pack
// kernel's cursor sends home), queues the copies, and hands the slot to the team.  A slice goes RAW --
// the plain cudaMemcpyAsync of the table bytes, as before -- when it is dense (packed size above
// 5/8 of the bytes; 1/4 when more than two ranks share the host), when it is no whole number of chunks, or when its slot is still being unpacked:
// the host's cores and the bus then share the work in whatever ratio keeps both busy.
// Measured on the 16-core GPU box, K=15 (profiles/r02n_*): 33.8 ms end to end with plain copies, 26.4 ms with
// this pipeline.  The 4-byte cursor copy on the work stream queues on the copy engine behind the bulk copies
// of the earlier windows, so the kernels run at most a window or two ahead of the transfer.  Taking that
// coupling away (cursor stored straight into pinned memory, 8 device / 6 host slots, one raw copy in
// flight at a time) let the flush, the copies and 15 streaming-store threads all hit host memory at once
// and measured 29.4 .. 30.9 ms whatever the lag and slot counts; coding the counts as nibbles + escapes
// (0.29 GB instead of 0.41 GB on the bus) made the team compute-bound and measured 30.3 .. 33.4 ms; counting
// the first 5/8 of the stream under the host-to-device copy (an early flush) changed nothing, the phase
// after the last input byte is bound by the host's memory writes, not by the GPU.  All three were removed.
void pk_unpack_chunks(const uint64_t *bitmap, const uint32_t *chunk_off, const uint8_t *nz, size_t nz_readable,
                      uint8_t *dst, size_t c0, size_t c1);          // unpack.cpp

struct TableShipper {
    static constexpr int kSlots = 4, kLag = 2;
    struct Job {
        int slot = 0;
        size_t n = 0, nz_bytes = 0;
        uint8_t *dst = nullptr;
        std::atomic<int> state{0};              // 0 = nobody waits for the copy yet, 1 = one thread does, 2 = arrived
        std::atomic<size_t> next{0}, done{0};   // chunks handed out / rebuilt
    };
    int device = 0;
    size_t max_n = 0;                           // entries of the largest slice
    size_t off_bytes = 0, nz_at = 0, nz_cap = 0, dev_bytes = 0, host_bytes = 0;
    uint8_t *d_buf[kSlots] = {nullptr}, *h_buf[kSlots] = {nullptr};
    uint32_t *d_cur = nullptr, *h_cur = nullptr;   // device cursors (one per slot); pinned: one word per slice
    size_t h_cur_cap = 0;
    cudaEvent_t packed[kSlots] = {nullptr}, copied[kSlots] = {nullptr};
    bool slot_copied[kSlots] = {false};         // copied[slot] has been recorded in this run
    std::atomic<int> slot_busy[kSlots];
    // one run = one flush
    std::vector<Job> jobs;
    std::atomic<size_t> njobs{0}, jobs_done{0};
    std::atomic<bool> closing{false};
    std::atomic<int> failed{0};
    std::vector<std::thread> team;
    struct Pending { size_t idx; int slot; const uint8_t *src; uint8_t *dst; size_t n; bool packable; };
    std::vector<Pending> pending;
    size_t issued = 0, serviced = 0, packed_slices = 0, raw_slices = 0, d2h_bytes = 0;
    int team_size = 0;
    int dense_eighths = 5;                      // a window goes packed up to this many eighths of its bytes
    bool running = false;

    static int threads() {
        if (const char *v = getenv("PYKMER_B200_UNPACK_THREADS"))
            if (atoi(v) > 0) return std::min(atoi(v), 64);
        int hc = (int)std::thread::hardware_concurrency(), local = 1;
        if (const char *v = getenv("LOCAL_WORLD_SIZE")) local = std::max(1, atoi(v));
        return std::max(1, std::min(hc / local - 1, 32));     // the caller's thread keeps a core
    }

    int ensure(int device_, size_t max_n_, size_t nslices) {
        const int rc = ensure_buffers(device_, max_n_, nslices);
        if (rc != PK_OK) release_buffers();                 // nothing half-allocated survives a failure
        return rc;
    }

    int ensure_buffers(int device_, size_t max_n_, size_t nslices) {
        if (max_n_ > max_n) {
            release_buffers();
            device = device_;
            max_n = max_n_;
            const size_t bm = max_n / 8;
            off_bytes = max_n / kPackChunk * sizeof(uint32_t);
            nz_at = (bm + off_bytes + 255) & ~(size_t)255;
            dev_bytes = nz_at + max_n + max_n / kPackChunk * 16 + 256;   // every chunk may pad up to 15 bytes
            nz_cap = max_n / 2 + 4096;                                    // denser slices go raw
            host_bytes = nz_at + nz_cap + 256;
            for (int i = 0; i < kSlots; i++) {
                PK_CUDA(cudaMalloc(&d_buf[i], dev_bytes));
                PK_CUDA(cudaHostAlloc(&h_buf[i], host_bytes, cudaHostAllocDefault));
                PK_CUDA(cudaEventCreateWithFlags(&packed[i], cudaEventDisableTiming));
                PK_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
            }
            PK_CUDA(cudaMalloc(&d_cur, kSlots * sizeof(uint32_t)));
        }
        if (nslices > h_cur_cap) {
            if (h_cur) cudaFreeHost(h_cur);
            h_cur = nullptr;
            PK_CUDA(cudaHostAlloc(&h_cur, nslices * sizeof(uint32_t), cudaHostAllocDefault));
            h_cur_cap = nslices;
        }
        return PK_OK;
    }

    void release_buffers() {
        for (int i = 0; i < kSlots; i++) {
            if (d_buf[i]) cudaFree(d_buf[i]);
            if (h_buf[i]) cudaFreeHost(h_buf[i]);
            if (packed[i]) cudaEventDestroy(packed[i]);
            if (copied[i]) cudaEventDestroy(copied[i]);
            d_buf[i] = h_buf[i] = nullptr;
            packed[i] = copied[i] = nullptr;
        }
        if (d_cur) cudaFree(d_cur);
        d_cur = nullptr;
        max_n = 0;
    }

    ~TableShipper() {
        stop_team();
        release_buffers();
        if (h_cur) cudaFreeHost(h_cur);
    }

    void team_main() {
        cudaSetDevice(device);
        for (size_t j = 0;; j++) {
            while (j >= njobs.load(std::memory_order_acquire)) {
                if (closing.load(std::memory_order_acquire) && j >= njobs.load(std::memory_order_acquire)) return;
                std::this_thread::yield();
            }
            Job &job = jobs[j];
            int expect = 0;
            if (job.state.compare_exchange_strong(expect, 1)) {          // one thread waits for the copy
                if (cudaEventSynchronize(copied[job.slot]) != cudaSuccess) failed.store(1);
                job.state.store(2, std::memory_order_release);
            } else {
                while (job.state.load(std::memory_order_acquire) != 2) std::this_thread::yield();
            }
            const size_t nchunks = job.n / kPackChunk, grain = 128;      // 128 KiB of table per grab
            const uint8_t *h = h_buf[job.slot];
            for (;;) {
                const size_t c0 = job.next.fetch_add(grain);
                if (c0 >= nchunks) break;
                const size_t c1 = std::min(nchunks, c0 + grain);
                if (!failed.load(std::memory_order_relaxed))
                    pk_unpack_chunks(reinterpret_cast<const uint64_t *>(h), reinterpret_cast<const uint32_t *>(h + max_n / 8),
                                     h + nz_at, job.nz_bytes + 64, job.dst, c0, c1);
                if (job.done.fetch_add(c1 - c0) + (c1 - c0) == nchunks) {  // last piece: the slot is free again
                    slot_busy[job.slot].store(0, std::memory_order_release);
                    jobs_done.fetch_add(1, std::memory_order_release);
                }
            }
        }
    }

    void begin(size_t nslices) {
        if (!team.empty()) stop_team();                     // a run that ended in an error left its team behind
        jobs = std::vector<Job>(nslices);
        njobs.store(0); jobs_done.store(0); closing.store(false); failed.store(0);
        for (int i = 0; i < kSlots; i++) { slot_busy[i].store(0); slot_copied[i] = false; }
        pending.clear();
        issued = serviced = packed_slices = raw_slices = d2h_bytes = 0;
        const int nt = team_size = threads();
        for (int t = 0; t < nt; t++) team.emplace_back([this] { team_main(); });
        running = true;
    }

    void stop_team() {
        closing.store(true, std::memory_order_release);
        for (auto &t : team) t.join();
        team.clear();
    }

    // after the slice's last kernel has been queued on st: pack it behind that kernel
    int pack(const uint8_t *slice_dev, uint8_t *dst_host, size_t n, cudaStream_t st, int sm_count) {
        const size_t idx = issued++;
        const int slot = (int)(idx % kSlots);
        const bool packable = n % kPackChunk == 0 && n <= max_n && n >= kPackChunk;
        if (packable) {
            if (slot_copied[slot]) PK_CUDA(cudaStreamWaitEvent(st, copied[slot], 0));   // its last copy has left the slot
            PK_CUDA(cudaMemsetAsync(d_cur + slot, 0, sizeof(uint32_t), st));
            const size_t nchunks = n / kPackChunk;
            const int grid = (int)std::min<size_t>((size_t)sm_count * 8, (nchunks + 7) / 8);
            k_table_pack<<<grid, 256, 0, st>>>(slice_dev, nchunks, reinterpret_cast<uint16_t *>(d_buf[slot]),
                                                 reinterpret_cast<uint32_t *>(d_buf[slot] + max_n / 8), d_buf[slot] + nz_at,
                                                 d_cur + slot);
            PK_CUDA(cudaGetLastError());
            PK_CUDA(cudaMemcpyAsync(h_cur + idx, d_cur + slot, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        } else {
            h_cur[idx] = 0xFFFFFFFFu;
        }
        PK_CUDA(cudaEventRecord(packed[slot], st));
        pending.push_back({idx, slot, slice_dev, dst_host, n, packable});
        return PK_OK;
    }

    // queue the copies of every slice issued at least `lag` slices ago
    int service(cudaStream_t copy_stream, size_t lag) {
        while (serviced + lag < issued) {
            const Pending p = pending[serviced++];
            const bool packable = p.packable;
            PK_CUDA(cudaEventSynchronize(packed[p.slot]));
            const size_t nz_bytes = packable ? (size_t)h_cur[p.idx] * 16 : 0;
            const size_t packed_bytes = p.n / 8 + p.n / kPackChunk * sizeof(uint32_t) + nz_bytes;
            const bool go_packed = packable && nz_bytes <= nz_cap && packed_bytes * 8 <= p.n * (size_t)dense_eighths &&
                                   slot_busy[p.slot].load(std::memory_order_acquire) == 0;
            PK_CUDA(cudaStreamWaitEvent(copy_stream, packed[p.slot], 0));
            if (!go_packed) {
                PK_CUDA(cudaMemcpyAsync(p.dst, p.src, p.n, cudaMemcpyDeviceToHost, copy_stream));
                raw_slices++;
                d2h_bytes += p.n + (packable ? sizeof(uint32_t) : 0);
                continue;
            }
            slot_busy[p.slot].store(1, std::memory_order_release);
            uint8_t *h = h_buf[p.slot];
            const uint8_t *d = d_buf[p.slot];
            PK_CUDA(cudaMemcpyAsync(h, d, p.n / 8, cudaMemcpyDeviceToHost, copy_stream));
            PK_CUDA(cudaMemcpyAsync(h + max_n / 8, d + max_n / 8, p.n / kPackChunk * sizeof(uint32_t),
                                    cudaMemcpyDeviceToHost, copy_stream));
            if (nz_bytes) PK_CUDA(cudaMemcpyAsync(h + nz_at, d + nz_at, nz_bytes, cudaMemcpyDeviceToHost, copy_stream));
            PK_CUDA(cudaEventRecord(copied[p.slot], copy_stream));
            slot_copied[p.slot] = true;
            Job &job = jobs[njobs.load(std::memory_order_relaxed)];
            job.slot = p.slot; job.n = p.n; job.nz_bytes = nz_bytes; job.dst = p.dst;
            njobs.fetch_add(1, std::memory_order_release);
            packed_slices++;
            d2h_bytes += packed_bytes + sizeof(uint32_t);
        }
        return PK_OK;
    }

    // drain: the remaining copies, then every slot rebuilt (raw copies are joined by the caller's stream sync)
    int finish(cudaStream_t copy_stream) {
        running = false;
        const int rc = service(copy_stream, 0);
        if (const char *v = getenv("PYKMER_B200_VERBOSE"))
            if (atoi(v) >= 1)
                fprintf(stderr, "[pykmer_b200] packed transfer: %zu slices (%zu packed, %zu raw), %.1f MB, %d host threads\n",
                        issued, packed_slices, raw_slices, d2h_bytes / 1e6, team_size);
        if (rc == PK_OK)
            while (jobs_done.load(std::memory_order_acquire) < njobs.load(std::memory_order_acquire)) std::this_thread::yield();
        stop_team();
        pending.clear();
        if (rc != PK_OK) return rc;
        if (failed.load()) return pk_set_error(PK_ERR_CUDA, "packed table transfer: waiting for a copy failed");
        return PK_OK;
    }
};

// start a packed transfer for this flush, or return with ix->shipper idle (plain copies)
static int shipper_begin(pk_indexer *ix, uint8_t *table_host, size_t slice_entries, size_t nslices) {
    if (!table_host || ix->table_bytes < ((size_t)1 << 22)) return PK_OK;
    // What the packed form is worth depends on how many cores rebuild and on how sparse the table is.  One GPU
    // (16 cores): K=15 26.4 .. 30.1 ms against 33.8 .. 34.3, K=17 140 against 316; two GPUs (24 cores): 19.1
    // against 23.4, 132 against 202.  With four or eight ranks on 32 cores (7 / 3 threads per rank) the ranks
    // fight over the host's memory: K=17 (3 % of the entries in use, packed size 1/6) still wins, 145 against 193 ms
    // and 171 against 172, K=15 (25 % in use, packed size 0.4) loses, 19.7 against 17.8 and 17.9 against 15.7 ms.
    // So with more than two ranks per host only windows that pack to a quarter of their bytes go packed.
    // PYKMER_B200_PACKED_D2H=0 turns the packed form off, =1 takes the 5/8 rule whatever the number of ranks.
    const char *v = getenv("PYKMER_B200_PACKED_D2H");
    if (v && atoi(v) == 0) return PK_OK;
    int local_ranks = 1;
    if (const char *lw = getenv("LOCAL_WORLD_SIZE")) local_ranks = std::max(1, atoi(lw));
    if (!ix->shipper) ix->shipper = new (std::nothrow) TableShipper();
    if (!ix->shipper) return pk_set_error(PK_ERR_NOMEM, "packed table transfer: out of host memory");
    ix->shipper->dense_eighths = (v || local_ranks <= 2) ? 5 : 2;
    const int rc = ix->shipper->ensure(ix->device, std::min(slice_entries, ix->table_bytes), nslices);
    if (rc != PK_OK) return rc;
    ix->shipper->begin(nslices);
    return PK_OK;
}

// one finished slice: packed behind its last kernel when a transfer is running, else the plain copy
static int ship_slice(pk_indexer *ix, cudaStream_t st, uint32_t b, const uint8_t *slice_dev, uint8_t *dst_host, size_t n) {
    if (ix->shipper && ix->shipper->running) {
        const int rc = ix->shipper->pack(slice_dev, dst_host, n, st, ix->sm_count);
        if (rc != PK_OK) return rc;
        ix->launches++;
        return ix->shipper->service(ix->copy_stream, TableShipper::kLag);
    }
    PK_CUDA(cudaEventRecord(ix->committed[b & 1u], st));
    PK_CUDA(cudaStreamWaitEvent(ix->copy_stream, ix->committed[b & 1u], 0));
    PK_CUDA(cudaMemcpyAsync(dst_host, slice_dev, n, cudaMemcpyDeviceToHost, ix->copy_stream));
    ix->xfer[0] += n;
    ix->xfer[2] += 1;
    return PK_OK;
}

// PARTITION: drain the buffered entries window by window into the table: count the
// window's entries into the L2-resident counters, then commit them (clamp, write the
// table slice once, re-zero, histogram).  Counting and committing two windows side by
// side on two streams was measured and is slower: both live off the same L2.
// When table_host != NULL every committed window is copied out at once on the copy
// stream, so the device-to-host transfer of the table overlaps the rest of the flush.
static int indexer_flush_finish(pk_indexer *ix, cudaStream_t st, bool with_stats, uint8_t *table_host);

// byte windows, first flush of a table: zero the table slice into L2, count in place with the
// histogram kept as transitions; nothing is committed and nothing is read back
static int indexer_flush_inplace(pk_indexer *ix, cudaStream_t st, uint8_t *table_host) {
    const size_t win = (size_t)1 << ix->win_log2;
    PK_CUDA(cudaMemsetAsync(ix->counters + 1, 0, 256 * sizeof(unsigned long long), st));
    const uint32_t *src = ix->pool_ext ? ix->pool_ext : (const uint32_t *)ix->pool;
    {
        const int rc = shipper_begin(ix, table_host, win, ix->nbuckets);
        if (rc != PK_OK) return rc;
    }
    for (uint32_t b = 0; b < ix->nbuckets; b++) {
        const size_t n = std::min(win, ix->table_bytes - (size_t)b * win);
        const size_t nvec = (n + 15) / 16;                  // the allocation is padded to 256 bytes
        uint8_t *tw = ix->table + (size_t)b * win;
        {
            prof_scope ps(ix, st, PROF_WINDOW_COMMIT);
            // 15 us per 64 MiB whatever the grid (2..16 blocks per SM) and also with cudaMemsetAsync:
            // the previous window's dirty lines leave L2 for DRAM at the same time
            const int zgrid = (int)std::max<size_t>(1, std::min<size_t>((size_t)ix->sm_count * ix->zero_blocks_per_sm, (nvec + 255) / 256));
            k_window_zero<<<zgrid, 256, 0, st>>>(reinterpret_cast<uint4 *>(tw), nvec);
        }
        ix->launches++;
        if (ix->nseg) {
            prof_scope ps(ix, st, PROF_WINDOW_COUNT);
            k_window_count8<true><<<ix->sm_count * ix->cnt8_blocks_per_sm, 256, 0, st>>>(
                src, seg_off(ix, 0), seg_cnt(ix, 0), ix->nseg, ix->nbuckets, b, reinterpret_cast<uint32_t *>(tw),
                nvec * 4, ix->ovf, ix->counters + 1);
            ix->launches++;
        }
        if (table_host) {                                   // ship this window while the next is counted
            const int rc = ship_slice(ix, st, b, tw, table_host + (size_t)b * win, n);
            if (rc != PK_OK) return rc;
        }
    }
    PK_CUDA(cudaGetLastError());
    const int rc = indexer_flush_finish(ix, st, false, table_host);
    ix->stats_valid = true;                                 // counters + 1 already hold the histogram
    return rc;
}

static int indexer_flush_l2(pk_indexer *ix, cudaStream_t st, bool with_stats, uint8_t *table_host) {
    if (ix->count8 && !ix->table_valid) return indexer_flush_inplace(ix, st, table_host);
    if (ix->count8 && !ix->scratch) {
        // a later flush onto a table that already holds counts (k-mer buffer overflow, feeding after
        // finalize): count into a zeroed scratch window, then add it in with byte saturation
        ix->scratch_bytes = std::max<size_t>((size_t)1 << ix->win_log2, 16);
        PK_CUDA(cudaMalloc(&ix->scratch, ix->scratch_bytes));
        PK_CUDA(cudaMemsetAsync(ix->scratch, 0, ix->scratch_bytes, st));
    }
    const size_t win = (size_t)1 << ix->win_log2;
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = window_launch_attr(ix, attr);
    const int rows = ix->sm_count * ix->commit_blocks_per_sm;   // commit grid = rows of partial bins
    if (with_stats)
        PK_CUDA(cudaMemsetAsync(ix->bins_part, 0, (size_t)8 * ix->sm_count * 256 * sizeof(unsigned long long), st));
    unsigned long long *bins = with_stats ? ix->bins_part : nullptr;
    const int grid = ix->sm_count * (ix->count8 ? ix->cnt8_blocks_per_sm : ix->cnt_blocks_per_sm);
    {
        const int rc = shipper_begin(ix, table_host, win, ix->nbuckets);
        if (rc != PK_OK) return rc;
    }
    for (uint32_t b = 0; b < ix->nbuckets; b++) {
        const size_t n = std::min(win, ix->table_bytes - (size_t)b * win);
        if (ix->nseg) {
            prof_scope ps(ix, st, PROF_WINDOW_COUNT);
            cfg.gridDim = dim3(grid);
            const uint32_t *src = ix->pool_ext ? ix->pool_ext : (const uint32_t *)ix->pool;
            if (ix->count8)
                PK_CUDA(cudaLaunchKernelEx(&cfg, k_window_count8<false>, src, (const uint32_t *)seg_off(ix, 0),
                                           (const uint32_t *)seg_cnt(ix, 0), ix->nseg, ix->nbuckets, b, ix->scratch,
                                           ix->scratch_bytes / sizeof(uint32_t), ix->ovf,
                                           (unsigned long long *)nullptr));
            else
                PK_CUDA(cudaLaunchKernelEx(&cfg, k_window_count, src, (const uint32_t *)seg_off(ix, 0),
                                           (const uint32_t *)seg_cnt(ix, 0), ix->nseg, ix->nbuckets, b, ix->scratch));
            ix->launches++;
        }
        uint8_t *tw = ix->table + (size_t)b * win;
        const size_t per_thread = ix->count8 ? 16 : 4;       // table entries one commit thread moves per step
        const int cgrid = (int)std::max<size_t>(1, std::min<size_t>((size_t)rows, (n / per_thread + 255) / 256));
        {
            prof_scope ps(ix, st, PROF_WINDOW_COMMIT);
            cfg.gridDim = dim3(cgrid);
            if (ix->count8) {
                if (ix->table_valid) PK_CUDA(cudaLaunchKernelEx(&cfg, k_window_commit8<true>, ix->scratch, tw, n, bins));
                else                 PK_CUDA(cudaLaunchKernelEx(&cfg, k_window_commit8<false>, ix->scratch, tw, n, bins));
            } else {
                if (ix->table_valid) PK_CUDA(cudaLaunchKernelEx(&cfg, k_window_commit<true>, ix->scratch, tw, n, bins));
                else                 PK_CUDA(cudaLaunchKernelEx(&cfg, k_window_commit<false>, ix->scratch, tw, n, bins));
            }
        }
        ix->launches++;
        if (table_host) {                                   // ship this window while the next is counted
            const int rc = ship_slice(ix, st, b, tw, table_host + (size_t)b * win, n);
            if (rc != PK_OK) return rc;
        }
    }
    return indexer_flush_finish(ix, st, with_stats, table_host);
}

// common end of a flush: join the table copies, reduce the histogram, recycle the buffers
static int indexer_flush_finish(pk_indexer *ix, cudaStream_t st, bool with_stats, uint8_t *table_host) {
    const int rows = ix->sm_count * 8;                      // every row of bins_part (unused ones are zero)
    if (ix->shipper && ix->shipper->running) {              // the last copies, and every slot rebuilt on the host
        const int rc = ix->shipper->finish(ix->copy_stream);
        if (rc != PK_OK) return rc;
        ix->xfer[0] = ix->shipper->d2h_bytes;
        ix->xfer[1] = ix->shipper->packed_slices;
        ix->xfer[2] = ix->shipper->raw_slices;
        ix->xfer[3] = (uint64_t)ix->shipper->team_size;
    }
    if (table_host) {
        PK_CUDA(cudaEventRecord(ix->committed[0], ix->copy_stream));
        PK_CUDA(cudaStreamWaitEvent(st, ix->committed[0], 0));
    }
    if (with_stats) {
        k_reduce_bins<<<1, 256, 0, st>>>(ix->bins_part, rows, ix->counters + 1);
        ix->launches++;
    }
    PK_CUDA(cudaGetLastError());
    if (const char *v = getenv("PYKMER_B200_VERBOSE")) {
        if (atoi(v) >= 2) {                                // debugging aid: synchronises
            uint32_t used = 0;
            std::vector<uint32_t> cnt((size_t)kMaxSegments * ix->nbuckets);
            cudaStreamSynchronize(st);
            cudaMemcpy(&used, ix->cursor, sizeof used, cudaMemcpyDeviceToHost);
            cudaMemcpy(cnt.data(), seg_cnt(ix, 0), cnt.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[pykmer_b200] flush: %u entries in %d segments, %u windows\n", used, ix->nseg, ix->nbuckets);
            for (uint32_t b = 0; b < ix->nbuckets; b++) {
                unsigned long long t = 0;
                for (int f = 0; f < ix->nseg; f++) t += cnt[(size_t)f * ix->nbuckets + b];
                fprintf(stderr, "[pykmer_b200] window %u entries %llu\n", b, t);
            }
        }
    }
    PK_CUDA(cudaMemsetAsync(ix->seg, 0, (size_t)4 * kMaxSegments * ix->nbuckets * sizeof(uint32_t), st));
    PK_CUDA(cudaMemsetAsync(ix->cursor, 0, 32, st));
    ix->nseg = 0;
    ix->pool_ub = 0;
    ix->pool_ext = nullptr;
    ix->table_valid = true;
    ix->stats_valid = with_stats;
    return PK_OK;
}

// PARTITION, dense tables: second-level split + shared-memory counting (k_sub_*).
static int indexer_flush_smem(pk_indexer *ix, cudaStream_t st, bool with_stats, uint8_t *table_host) {
    const size_t win = (size_t)1 << ix->win_log2;
    const int rows = ix->sm_count * 4;                      // >= kSubs rows of partial bins
    if (with_stats)
        PK_CUDA(cudaMemsetAsync(ix->bins_part, 0, (size_t)8 * ix->sm_count * 256 * sizeof(unsigned long long), st));
    unsigned long long *bins = with_stats ? ix->bins_part : nullptr;
    const uint32_t *src = ix->pool_ext ? ix->pool_ext : ix->pool;
    constexpr uint32_t kGroup = 64;                         // windows regrouped per pass over the pool
    uint32_t *sub_cnt = ix->sub, *sub_off = ix->sub + kGroup * kSubs, *sub_fill = ix->sub + 2 * kGroup * kSubs;
    const size_t smem_scatter = (size_t)3 * kSubs * sizeof(uint32_t) + (size_t)kSubTile * 6;
    const size_t smem_tally = ((size_t)(1 << kSubLog2) + 8 * 256) * sizeof(uint32_t);
    if (!ix->sub_smem_set) {
        PK_CUDA(cudaFuncSetAttribute(k_sub_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scatter));
        PK_CUDA(cudaFuncSetAttribute(k_sub_tally<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tally));
        PK_CUDA(cudaFuncSetAttribute(k_sub_tally<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tally));
        ix->sub_smem_set = true;
    }
    for (uint32_t w0 = 0; w0 < ix->nbuckets; w0 += kGroup) {
        const uint32_t W = std::min(kGroup, ix->nbuckets - w0);
        PK_CUDA(cudaMemsetAsync(ix->sub, 0, (size_t)3 * kGroup * kSubs * sizeof(uint32_t), st));
        if (ix->nseg) {
            {
                prof_scope ps(ix, st, PROF_WINDOW_COUNT);
                k_sub_count<<<dim3(32, W), 256, 0, st>>>(src, seg_off(ix, 0), seg_cnt(ix, 0), ix->nseg, ix->nbuckets,
                                                        w0, ix->win_log2, sub_cnt);
                k_sub_offsets<<<1, 1024, 0, st>>>(sub_cnt, sub_off, W * kSubs);
                k_sub_scatter<<<dim3(16, W), kSubThreads, smem_scatter, st>>>(
                    src, seg_off(ix, 0), seg_cnt(ix, 0), ix->nseg, ix->nbuckets, w0, ix->win_log2, sub_off, sub_fill,
                    ix->pool2);
            }
            ix->launches += 3;
        }
        for (uint32_t b = 0; b < W; b++) {
            const size_t n = std::min(win, ix->table_bytes - (size_t)(w0 + b) * win);
            uint8_t *tw = ix->table + (size_t)(w0 + b) * win;
            const unsigned nsub = (unsigned)((n + (1u << kSubLog2) - 1) >> kSubLog2);
            {
                prof_scope ps(ix, st, PROF_WINDOW_COMMIT);
                if (ix->table_valid) k_sub_tally<true><<<nsub, 1024, smem_tally, st>>>(ix->pool2, sub_off, sub_cnt, b, tw, n, bins);
                else                 k_sub_tally<false><<<nsub, 1024, smem_tally, st>>>(ix->pool2, sub_off, sub_cnt, b, tw, n, bins);
            }
            ix->launches++;
            if (table_host) {                               // ship this window while the next is counted
                PK_CUDA(cudaEventRecord(ix->committed[b & 1u], st));
                PK_CUDA(cudaStreamWaitEvent(ix->copy_stream, ix->committed[b & 1u], 0));
                PK_CUDA(cudaMemcpyAsync(table_host + (size_t)(w0 + b) * win, tw, n, cudaMemcpyDeviceToHost, ix->copy_stream));
            }
        }
    }
    return indexer_flush_finish(ix, st, with_stats, table_host);
}

static int indexer_flush(pk_indexer *ix, cudaStream_t st, bool with_stats, uint8_t *table_host) {
    return ix->flush_smem ? indexer_flush_smem(ix, st, with_stats, table_host)
                          : indexer_flush_l2(ix, st, with_stats, table_host);
}

enum { SCAN_BOTH = 0, SCAN_PASS1 = 1, SCAN_PASS2_REMOTE = 2, SCAN_ROUTED = 3 };

static int indexer_launch_scan(pk_indexer *ix, const uint8_t *seq_dev, size_t n, cudaStream_t st,
                               int phase = SCAN_BOTH) {
    if (n == 0) return PK_OK;
    ScanParams p;
    memset(&p, 0, sizeof p);
    p.seq = seq_dev; p.n = n; p.carry = ix->carry; p.K = ix->K; p.lo = ix->lo; p.span = ix->hi - ix->lo;
    p.table = ix->table; p.num_kmers = ix->counters; p.bins = ix->counters + 1;
    p.rec_starts = ix->rec_starts; p.nrec = ix->nrec; p.rec_flags = ix->nrec ? ix->rec_flags : nullptr;
    p.stream_off = ix->stream_off;
    const bool wide = ix->K > 16;
    const bool full = ix->lo == 0 && ix->hi == (1ull << (2 * ix->K));
#define PK_LAUNCH_SCAN(KERNEL, GRID, SMEM)                                                     \
    do {                                                                                       \
        if (wide) { if (full) KERNEL<true, true><<<GRID, kScanThreads, SMEM, st>>>(p);         \
                    else      KERNEL<true, false><<<GRID, kScanThreads, SMEM, st>>>(p); }      \
        else      { if (full) KERNEL<false, true><<<GRID, kScanThreads, SMEM, st>>>(p);        \
                    else      KERNEL<false, false><<<GRID, kScanThreads, SMEM, st>>>(p); }     \
    } while (0)
#define PK_OCCUPANCY(KERNEL, SMEM, OUT)                                                        \
    do {                                                                                       \
        cudaError_t oe__;                                                                      \
        if (wide) oe__ = full ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&OUT, KERNEL<true, true>, kScanThreads, SMEM)    \
                              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&OUT, KERNEL<true, false>, kScanThreads, SMEM);  \
        else      oe__ = full ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&OUT, KERNEL<false, true>, kScanThreads, SMEM)   \
                              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&OUT, KERNEL<false, false>, kScanThreads, SMEM); \
        if (oe__ != cudaSuccess || OUT < 1) { cudaGetLastError(); OUT = 1; }                   \
    } while (0)
#define PK_SMEM_OPT_IN(KERNEL, BYTES)                                                          \
    do {                                                                                       \
        PK_CUDA(cudaFuncSetAttribute(KERNEL<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES)));   \
        PK_CUDA(cudaFuncSetAttribute(KERNEL<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES)));  \
        PK_CUDA(cudaFuncSetAttribute(KERNEL<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES)));  \
        PK_CUDA(cudaFuncSetAttribute(KERNEL<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES))); \
    } while (0)
    const long long gpw = wide ? 30 : 31;
    const long long ngroups = (long long)((n + 15) / 16);
    const long long ntiles = (ngroups + gpw - 1) / gpw;
    const long long want = (ntiles + kScanWarps - 1) / kScanWarps;
    p.ngroups = ngroups;
    p.ntiles = ntiles;
    // grid-stride kernels: exactly as many blocks as are resident at once (no partial second wave)
    if (ix->mode == PK_MODE_DIRECT) {
        int per_sm = 0;
        PK_OCCUPANCY(k_scan_count_direct, 0, per_sm);
        const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)ix->sm_count * per_sm));
        {
            prof_scope ps(ix, st, PROF_SCAN_DIRECT);
            PK_LAUNCH_SCAN(k_scan_count_direct, grid, 0);
        }
        PK_CUDA(cudaGetLastError());
        ix->launches += 1;
    } else {
        // Estimated pass 1 (PARTITION, big feeds): pass 1 only has to size the segments, so it
        // looks at one tile in 16 and every window gets room for its estimate + 1/16 + kCapSlack;
        // pass 2 checks each reservation against that room and does the bookkeeping (num_kmers,
        // record flags).  If a window ever overflows, k_feed_settle forgets the attempt and the
        // exact kernels queued behind -- which otherwise return at once -- redo the feed.
        const size_t est_need = n + n / 8 + (size_t)ix->nbuckets * kCapSlack;
        bool estimate = phase == SCAN_BOTH && ix->mode == PK_MODE_PARTITION && ix->est_shift > 0 &&
                        n >= ix->est_min && est_need <= ix->pool_cap && est_need < (1ull << 32);
        const size_t need = estimate ? est_need : n;
        const bool remote = phase == SCAN_PASS2_REMOTE || phase == SCAN_ROUTED;
        if (!remote && (ix->nseg == kMaxSegments || ix->pool_ub + need > ix->pool_cap)) {
            if (ix->mode == PK_MODE_SCAN)
                return pk_set_error(PK_ERR_STATE, "scan-only handle: k-mer buffer full (%zu entries, %d segments); "
                                    "export and reset before feeding more", ix->pool_cap, ix->nseg);
            const int rc = indexer_flush(ix, st, false, nullptr);
            if (rc != PK_OK) return rc;
        }
        const int f = ix->nseg;
        p.win_log2 = ix->win_log2; p.nbuckets = ix->nbuckets; p.pool = ix->pool;
        p.seg_cnt = seg_cnt(ix, f); p.seg_off = seg_off(ix, f); p.seg_fill = seg_fill(ix, f);
        p.ctl = ix->cursor;
        p.pass1_tally = 1;
        const size_t smem1 = (size_t)ix->nbuckets * sizeof(uint32_t);
        const size_t smem2 = scatter_smem_bytes(ix->nbuckets);
        if (!ix->scatter_smem_set) {
            PK_SMEM_OPT_IN(k_scan_scatter, smem2);
            if (smem1 > 48 * 1024) PK_SMEM_OPT_IN(k_scan_bucket_count, smem1);
            ix->scatter_smem_set = true;
        }
        int per_sm1 = 0, per_sm2 = 0;
        PK_OCCUPANCY(k_scan_bucket_count, smem1, per_sm1);
        PK_OCCUPANCY(k_scan_scatter, smem2, per_sm2);
        // grid-stride / persistent kernels: exactly as many blocks as are resident at once
        const int grid1 = (int)std::max<long long>(1, std::min<long long>(want, (long long)ix->sm_count * per_sm1));
        const int grid2 = (int)std::max<long long>(1, std::min<long long>(want, (long long)ix->sm_count * per_sm2));
        if (estimate) {
            ScanParams q = p;
            q.sample_shift = ix->est_shift; q.pass1_tally = 0; q.pass2_tally = 1; q.seg_cap = seg_cap(ix, f);
            {
                prof_scope ps(ix, st, PROF_BUCKET_COUNT);
                const ScanParams keep = p; p = q;
                PK_LAUNCH_SCAN(k_scan_bucket_count, grid1, smem1);
                p = keep;
            }
            {
                prof_scope ps(ix, st, PROF_OFFSETS);
                k_bucket_offsets<true><<<1, 256, 0, st>>>(q.seg_cnt, seg_off(ix, f), ix->nbuckets, ix->cursor,
                                                          seg_cap(ix, f), ix->est_shift, (uint32_t)est_need,
                                                          ix->est_test, nullptr);
            }
            {
                prof_scope ps(ix, st, PROF_SCATTER);
                const ScanParams keep = p; p = q;
                PK_LAUNCH_SCAN(k_scan_scatter, grid2, smem2);
                p = keep;
            }
            {
                prof_scope ps(ix, st, PROF_OFFSETS);
                k_feed_settle<<<1, 256, 0, st>>>(q.seg_cnt, q.seg_fill, ix->nbuckets, ix->cursor, ix->counters);
            }
            ix->launches += 4;
            p.run_if = ix->cursor + 1;                     // the exact kernels below run only after an overflow
        }
        if (!remote) {
            prof_scope ps(ix, st, PROF_BUCKET_COUNT);
            PK_LAUNCH_SCAN(k_scan_bucket_count, grid1, smem1);
            ix->launches += 1;
        }
        if (phase == SCAN_BOTH) {
            prof_scope ps(ix, st, PROF_OFFSETS);
            k_bucket_offsets<false><<<1, 256, 0, st>>>(p.seg_cnt, seg_off(ix, f), ix->nbuckets, ix->cursor,
                                                       nullptr, 0u, 0u, 0u, p.run_if);
            ix->launches += 1;
        }
        if (phase == SCAN_PASS1) {
            PK_CUDA(cudaGetLastError());
            ix->p1_seq = seq_dev;
            ix->p1_n = n;
            ix->last_stream = st;
            return PK_OK;                              // pass 2 follows once the routing is known
        }
        if (remote) {
            p.win_owner = ix->route;
            p.dest_off = ix->route + kMaxBuckets;
            for (int i = 0; i < 16; i++) p.peer[i] = ix->peer_pool[i];
        }
        if (phase == SCAN_ROUTED) {                        // one pass: fixed, capacity-checked regions; pass 2 tallies
            p.seg_cap = ix->route + 2 * kMaxBuckets;
            p.pass1_tally = 0;
            p.pass2_tally = 1;
        }
        {
            prof_scope ps(ix, st, PROF_SCATTER);
            PK_LAUNCH_SCAN(k_scan_scatter, grid2, smem2);
        }
        PK_CUDA(cudaGetLastError());
        ix->launches += 1;
        if (phase == SCAN_ROUTED) {
            PublishPeers pp;
            for (int i = 0; i < 16; i++) { pp.pool[i] = ix->peer_pool[i]; pp.pub_base[i] = ix->pub_base[i]; }
            prof_scope ps(ix, st, PROF_OFFSETS);
            k_publish_fill<<<(ix->nbuckets + 255) / 256, 256, 0, st>>>(
                seg_fill(ix, f), ix->route + 2 * kMaxBuckets, ix->nbuckets, ix->route, pp, (uint32_t)ix->self_rank,
                ix->cursor, ix->counters, ix->routed_status);
            PK_CUDA(cudaGetLastError());
            ix->launches += 1;
        }
        if (phase == SCAN_BOTH) {
            ix->nseg++;
            ix->pool_ub += need;
        }
        ix->stats_valid = false;
    }
    {
        prof_scope ps(ix, st, PROF_CARRY);
        k_update_carry<<<1, 32, 0, st>>>(ix->carry, seq_dev, n, ix->mode != PK_MODE_DIRECT ? ix->cursor : nullptr);
    }
    PK_CUDA(cudaGetLastError());
    ix->launches += 1;
    ix->stream_off += n;
    ix->fed = true;
    ix->last_stream = st;
    return PK_OK;
#undef PK_LAUNCH_SCAN
#undef PK_SMEM_OPT_IN
#undef PK_OCCUPANCY
}

// feeds are cut so that one partition pass never exceeds kMaxFeed bases
static int indexer_feed_device(pk_indexer *ix, const uint8_t *seq_dev, size_t n, cudaStream_t st) {
    const size_t step = ix->mode != PK_MODE_DIRECT ? std::min(kMaxFeed, ix->pool_cap) : n;
    size_t off = 0;
    while (off < n) {
        const size_t len = std::min(step, n - off);
        const int rc = indexer_launch_scan(ix, seq_dev + off, len, st);
        if (rc != PK_OK) return rc;
        off += len;
    }
    return PK_OK;
}

PK_API int pk_indexer_create(pk_indexer **out, int kmer_len, int device, uint64_t range_lo,
                             uint64_t range_hi, int mode) {
    PK_REQUIRE(out != nullptr, "pk_indexer_create: out is NULL");
    *out = nullptr;
    // tools.py:165-167: kmer_len > 0 and odd; 2K bits must fit the 64-bit k-mer word
    PK_REQUIRE(kmer_len > 0 && (kmer_len % 2) == 1 && kmer_len <= 31,
               "pk_indexer_create: kmer_len must be odd and in 1..31, got %d", kmer_len);
    const uint64_t T = 1ull << (2 * kmer_len);
    if (range_hi == 0) range_hi = T;
    PK_REQUIRE(range_lo < range_hi && range_hi <= T, "pk_indexer_create: bad range [%llu, %llu) for 4^K = %llu",
               (unsigned long long)range_lo, (unsigned long long)range_hi, (unsigned long long)T);
    PK_REQUIRE(mode == PK_MODE_AUTO || mode == PK_MODE_DIRECT || mode == PK_MODE_PARTITION ||
               mode == PK_MODE_SCAN, "pk_indexer_create: unknown mode %d", mode);
    int ndev = 0;
    PK_CUDA(cudaGetDeviceCount(&ndev));
    PK_REQUIRE(device >= 0 && device < ndev, "pk_indexer_create: device %d of %d", device, ndev);
    pk_device_guard guard(device);
    if (!guard.ok) return pk_set_error(PK_ERR_CUDA, "cudaSetDevice(%d) failed", device);

    pk_indexer *ix = new (std::nothrow) pk_indexer();
    if (!ix) return pk_set_error(PK_ERR_NOMEM, "out of host memory");
    ix->K = kmer_len; ix->device = device; ix->lo = range_lo; ix->hi = range_hi;
    ix->table_bytes = (size_t)(range_hi - range_lo);
    ix->sm_count = pk_sm_count(device);

    // window size: 2^24 counters (64 MiB of u32) stay L2-resident on B200; the
    // environment override exists so that tests can force many windows on small tables
    // how a window is counted (PYKMER_B200_FLUSH = l2 | byte | smem):
    //   l2    32-bit counters kept in L2 (default for K <= 15)
    //   byte  the window holds the table's own 8-bit lanes, 2^26 entries per window (default for
    //         K >= 17, where a window sees few k-mers and the per-window passes dominate).  The
    //         rule depends on K alone so that every handle of a multi-GPU job picks the same windows.
    //   smem  second-level split + shared-memory counters.  Measured on config 2
    //         (profiles/r01_ncu_sub_kernels.txt): L2 flush 6.5 ms, smem flush 13.8 ms; kept as the
    //         second, independent implementation the tests cross-check.
    const char *fe = getenv("PYKMER_B200_FLUSH");
    ix->flush_smem = fe && strcmp(fe, "smem") == 0;
    ix->count8 = fe ? strcmp(fe, "byte") == 0 : kmer_len >= 17;
    if (const char *ge = getenv("PYKMER_B200_CNT8_GRID")) {
        const int v = atoi(ge);
        if (v >= 1 && v <= 32) ix->cnt8_blocks_per_sm = v;
    }
    if (const char *ge = getenv("PYKMER_B200_COMMIT_GRID")) {
        const int v = atoi(ge);
        if (v >= 1 && v <= 8) ix->commit_blocks_per_sm = v;
    }
    if (const char *ge = getenv("PYKMER_B200_ZERO_GRID")) {
        const int v = atoi(ge);
        if (v >= 1 && v <= 32) ix->zero_blocks_per_sm = v;
    }
    if (const char *ge = getenv("PYKMER_B200_CNT_GRID")) {
        const int v = atoi(ge);
        if (v >= 1 && v <= 32) ix->cnt_blocks_per_sm = v;
    }
    uint32_t win_log2 = ix->count8 ? 26 : 24;
    if (const char *env = getenv("PYKMER_B200_WINDOW_LOG2")) {
        const int v = atoi(env);
        if (v >= 4 && v <= (ix->count8 ? 26 : 24)) win_log2 = (uint32_t)v;
    }
    const uint64_t nb = (ix->table_bytes + ((1ull << win_log2) - 1)) >> win_log2;
    // AUTO: tiny tables (K <= 9) and very sparse ones (K >= 19: a genome fills well under 1 % of
    // 4^19 entries, so sweeping the table window by window costs more than one random DRAM update
    // per k-mer) count DIRECT; everything between is PARTITIONed -- also K = 11 and 13, whose
    // tables would fit L2: their k-mers repeat so often that compare-and-swap keeps colliding
    // (measured on the 782 Mbp stream: K=11 26.0 ms DIRECT / 10.6 ms PARTITION, K=13 27.9 / 9.5).
    if (mode == PK_MODE_AUTO)
        mode = (ix->table_bytes >= (1ull << 22) && kmer_len < 19 && nb <= (uint64_t)kMaxBuckets)
                   ? PK_MODE_PARTITION : PK_MODE_DIRECT;
    if ((mode == PK_MODE_PARTITION || mode == PK_MODE_SCAN) && nb > (uint64_t)kMaxBuckets) {
        delete ix;
        return pk_set_error(PK_ERR_ARG, "pk_indexer_create: range needs %llu windows, at most %d "
                            "are supported in PARTITION mode", (unsigned long long)nb, kMaxBuckets);
    }
    ix->mode = mode;
    ix->win_log2 = win_log2;
    ix->nbuckets = (uint32_t)nb;

    const size_t alloc = (ix->table_bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    if (mode != PK_MODE_SCAN) step(cudaMalloc(&ix->table, alloc));     // a scanner owns no table
    step(cudaMalloc(&ix->carry, 64));
    step(cudaMalloc(&ix->counters, 257 * sizeof(unsigned long long)));
    step(cudaHostAlloc(&ix->h_counters, 257 * sizeof(unsigned long long), cudaHostAllocDefault));
    step(cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking));
    step(cudaStreamCreateWithFlags(&ix->work_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        step(cudaEventCreateWithFlags(&ix->copied[i], cudaEventDisableTiming));
        step(cudaEventCreateWithFlags(&ix->consumed[i], cudaEventDisableTiming));
    }
    step(cudaEventCreateWithFlags(&ix->joined, cudaEventDisableTiming));
    ix->last_stream = ix->work_stream;
    if ((mode == PK_MODE_PARTITION || mode == PK_MODE_SCAN) && e == cudaSuccess) {
        size_t free_b = 0, total_b = 0;
        step(cudaMemGetInfo(&free_b, &total_b));
        size_t cap = (size_t)1 << 30;                      // 2^30 entries = 4 GiB
        if (const char *env = getenv("PYKMER_B200_POOL_LOG2")) {
            const int v = atoi(env);
            if (v >= 12 && v <= 30) cap = (size_t)1 << v;
        }
        while (cap > ((size_t)1 << 20) && cap * sizeof(uint32_t) > free_b / 4) cap >>= 1;
        ix->pool_cap = cap;
        // test hooks of the estimated pass 1: PYKMER_B200_EST = "shift[:min_log2[:test]]"
        // (shift 0 = always exact; test 1 = as if over budget, 2 = reserve too little room)
        if (const char *env = getenv("PYKMER_B200_EST")) {
            int sh = 4, ml = 24, test = 0;
            sscanf(env, "%d:%d:%d", &sh, &ml, &test);
            ix->est_shift = (uint32_t)std::min(std::max(sh, 0), 8);
            ix->est_min = (size_t)1 << std::min(std::max(ml, 0), 40);
            ix->est_test = (uint32_t)test;
        }
        const size_t seg_bytes = (size_t)4 * kMaxSegments * ix->nbuckets * sizeof(uint32_t);
        step(cudaMalloc(&ix->pool, cap * sizeof(uint32_t)));
        step(cudaMalloc(&ix->seg, seg_bytes));
        step(cudaMalloc(&ix->cursor, 256));
        if (mode == PK_MODE_PARTITION) {
            step(cudaMalloc(&ix->bins_part, (size_t)8 * ix->sm_count * 256 * sizeof(unsigned long long)));
            if (ix->flush_smem) {
                step(cudaMalloc(&ix->pool2, cap * sizeof(uint32_t)));
                step(cudaMalloc(&ix->sub, (size_t)3 * 64 * kSubs * sizeof(uint32_t)));
            } else if (!ix->count8) {                       // byte windows count in the table itself
                ix->scratch_bytes = sizeof(uint32_t) << win_log2;
                step(cudaMalloc(&ix->scratch, ix->scratch_bytes));
            }
            if (ix->count8) {
                uint32_t ovf_log2 = 16;
                if (const char *env = getenv("PYKMER_B200_OVF_LOG2")) {
                    const int v = atoi(env);
                    if (v >= 1 && v <= 22) ovf_log2 = (uint32_t)v;
                }
                const size_t oc = (size_t)1 << ovf_log2;
                ix->ovf.mask = (uint32_t)oc - 1u;
                step(cudaMalloc(&ix->ovf.keys, oc * sizeof(uint32_t)));
                step(cudaMalloc(&ix->ovf.vals, oc * sizeof(unsigned long long)));
                step(cudaMalloc(&ix->ovf.list, oc * sizeof(uint32_t)));
                step(cudaMalloc(&ix->ovf.meta, 256));
                if (e == cudaSuccess) {
                    step(cudaMemsetAsync(ix->ovf.keys, 0xFF, oc * sizeof(uint32_t), ix->work_stream));
                    step(cudaMemsetAsync(ix->ovf.vals, 0, oc * sizeof(unsigned long long), ix->work_stream));
                    step(cudaMemsetAsync(ix->ovf.meta, 0, 256, ix->work_stream));
                }
            }
        }
        for (int i = 0; i < 2; i++)
            step(cudaEventCreateWithFlags(&ix->committed[i], cudaEventDisableTiming));
        // persisting-L2 carve-out for the counters (PYKMER_B200_L2_PERSIST=0 disables it)
        const char *pe = getenv("PYKMER_B200_L2_PERSIST");
        if (e == cudaSuccess && mode == PK_MODE_PARTITION && !ix->flush_smem && !(pe && atoi(pe) == 0)) {
            int max_persist = 0, max_window = 0;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device);
            const size_t want = ix->scratch_bytes;
            size_t grant = std::min<size_t>(want, (size_t)std::max(max_persist, 0));
            if (grant && (size_t)max_window >= want &&
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, grant) == cudaSuccess) {
                ix->l2_persist_bytes = grant;
                g_persist_users[device & 63]++;
            } else {
                cudaGetLastError();
            }
            if (getenv("PYKMER_B200_VERBOSE"))
                fprintf(stderr, "[pykmer_b200] L2 persist: max %d B, max window %d B, window counters "
                        "%zu B, granted %zu B\n", max_persist, max_window, want, ix->l2_persist_bytes);
        }
        if (e == cudaSuccess) {
            step(cudaMemsetAsync(ix->seg, 0, seg_bytes, ix->work_stream));
            step(cudaMemsetAsync(ix->cursor, 0, 256, ix->work_stream));
            if (ix->scratch)
                step(cudaMemsetAsync(ix->scratch, 0, ix->scratch_bytes, ix->work_stream));
        }
    }
    if (e == cudaSuccess) {
        if (mode == PK_MODE_DIRECT) step(table_zero(ix, ix->work_stream));
        step(cudaMemsetAsync(ix->carry, 0, 64, ix->work_stream));
        step(cudaMemsetAsync(ix->counters, 0, 257 * sizeof(unsigned long long), ix->work_stream));
        step(cudaStreamSynchronize(ix->work_stream));
    }
    if (e != cudaSuccess) {
        const int code = pk_set_error(e == cudaErrorMemoryAllocation ? PK_ERR_NOMEM : PK_ERR_CUDA,
                                      "pk_indexer_create(K=%d, %zu table bytes): %s", kmer_len,
                                      ix->table_bytes, cudaGetErrorString(e));
        cudaGetLastError();
        pk_indexer_destroy(ix);
        return code;
    }
    *out = ix;
    return PK_OK;
}

PK_API int pk_indexer_destroy(pk_indexer *ix) {
    if (!ix) return PK_OK;
    pk_device_guard guard(ix->device);
    if (ix->work_stream) cudaStreamSynchronize(ix->work_stream);
    if (ix->copy_stream) cudaStreamSynchronize(ix->copy_stream);
    cudaFree(ix->table); cudaFree(ix->carry); cudaFree(ix->counters);
    cudaFree(ix->rec_starts); cudaFree(ix->rec_flags);
    cudaFree(ix->stage[0]); cudaFree(ix->stage[1]);
    cudaFree(ix->pool); cudaFree(ix->seg); cudaFree(ix->cursor); cudaFree(ix->scratch);
    cudaFree(ix->bins_part); cudaFree(ix->route); cudaFree(ix->pool2); cudaFree(ix->sub);
    cudaFree(ix->import_off);
    delete ix->shipper;
    cudaFree(ix->ovf.keys); cudaFree(ix->ovf.vals); cudaFree(ix->ovf.list); cudaFree(ix->ovf.meta);
    if (ix->l2_persist_bytes) {
        // give the carve-out back with its last user: left in place it shrinks the L2 of everything that
        // runs later in the process (measured: K=17's in-place byte windows after a K=15 handle,
        // 23.2 ms against 17.2 ms, profiles/r02a_bench_default_line_1gpu.json)
        cudaCtxResetPersistingL2Cache();
        if (--g_persist_users[ix->device & 63] <= 0) {
            g_persist_users[ix->device & 63] = 0;
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
        }
    }
    for (int i = 0; i < 16; i++)
        if (ix->peer_ipc[i]) cudaIpcCloseMemHandle(ix->peer_ipc[i]);
    for (int i = 0; i < 2; i++)
        if (ix->committed[i]) cudaEventDestroy(ix->committed[i]);
    if (ix->h_counters) cudaFreeHost(ix->h_counters);
    for (int i = 0; i < 2; i++) {
        if (ix->copied[i]) cudaEventDestroy(ix->copied[i]);
        if (ix->consumed[i]) cudaEventDestroy(ix->consumed[i]);
    }
    if (ix->joined) cudaEventDestroy(ix->joined);
    for (cudaEvent_t e : ix->prof_events) cudaEventDestroy(e);
    if (ix->copy_stream) cudaStreamDestroy(ix->copy_stream);
    if (ix->work_stream) cudaStreamDestroy(ix->work_stream);
    delete ix;
    return PK_OK;
}

PK_API int pk_indexer_reset(pk_indexer *ix, pk_stream stream) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_reset: NULL handle");
    pk_device_guard guard(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    // order after everything already queued on the handle, then hand over to `stream`
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    PK_CUDA(cudaEventRecord(ix->joined, ix->work_stream));
    PK_CUDA(cudaStreamWaitEvent(st, ix->joined, 0));
    if (ix->mode == PK_MODE_DIRECT) {
        PK_CUDA(table_zero(ix, st));
    } else {
        // the first flush rewrites every window, so the table itself needs no memset
        PK_CUDA(cudaMemsetAsync(ix->seg, 0, (size_t)4 * kMaxSegments * ix->nbuckets * sizeof(uint32_t), st));
        PK_CUDA(cudaMemsetAsync(ix->cursor, 0, 32, st));
        ix->nseg = 0;
        ix->pool_ub = 0;
        ix->pool_ext = nullptr;
        ix->p1_seq = nullptr;
        ix->p1_n = 0;
        ix->table_valid = false;
        ix->stats_valid = false;
    }
    PK_CUDA(cudaMemsetAsync(ix->carry, 0, 64, st));
    PK_CUDA(cudaMemsetAsync(ix->counters, 0, 257 * sizeof(unsigned long long), st));
    if (ix->nrec) PK_CUDA(cudaMemsetAsync(ix->rec_flags, 0, ix->nrec, st));
    ix->stream_off = 0;
    ix->fed = false;
    ix->last_stream = st;
    return PK_OK;
}

PK_API int pk_indexer_set_records(pk_indexer *ix, const uint64_t *rec_starts_host, size_t nrec) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_set_records: NULL handle");
    PK_REQUIRE(nrec == 0 || rec_starts_host != nullptr, "pk_indexer_set_records: NULL table");
    PK_REQUIRE(nrec >= ix->nrec, "pk_indexer_set_records: the table may only grow (%zu -> %zu)",
               ix->nrec, nrec);
    for (size_t i = 1; i < nrec; i++)
        PK_REQUIRE(rec_starts_host[i - 1] <= rec_starts_host[i],
                   "pk_indexer_set_records: offsets must ascend (entry %zu)", i);
    if (nrec == 0) return PK_OK;
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    PK_CUDA(cudaStreamSynchronize(ix->work_stream));       // kernels may still read the old table
    if (nrec > ix->rec_cap) {
        const size_t cap = std::max<size_t>(std::max<size_t>(nrec, 2 * ix->rec_cap), 1024);
        uint64_t *starts = nullptr;
        uint8_t *flags = nullptr;
        PK_CUDA(cudaMalloc(&starts, cap * sizeof(uint64_t)));
        cudaError_t e = cudaMalloc(&flags, cap);
        if (e == cudaSuccess) e = cudaMemset(flags, 0, cap);
        if (e == cudaSuccess && ix->nrec)
            e = cudaMemcpy(flags, ix->rec_flags, ix->nrec, cudaMemcpyDeviceToDevice);
        if (e != cudaSuccess) {
            cudaFree(starts); cudaFree(flags);
            return pk_set_error(PK_ERR_CUDA, "pk_indexer_set_records: %s", cudaGetErrorString(e));
        }
        cudaFree(ix->rec_starts); cudaFree(ix->rec_flags);
        ix->rec_starts = starts; ix->rec_flags = flags; ix->rec_cap = cap;
    }
    PK_CUDA(cudaMemcpy(ix->rec_starts, rec_starts_host, nrec * sizeof(uint64_t), cudaMemcpyHostToDevice));
    ix->nrec = nrec;
    ix->last_rec_start = rec_starts_host[nrec - 1];
    return PK_OK;
}

PK_API int pk_indexer_append_records(pk_indexer *ix, const uint64_t *new_starts_host, size_t n) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_append_records: NULL handle");
    PK_REQUIRE(n == 0 || new_starts_host != nullptr, "pk_indexer_append_records: NULL table");
    if (n == 0) return PK_OK;
    for (size_t i = 1; i < n; i++)
        PK_REQUIRE(new_starts_host[i - 1] <= new_starts_host[i],
                   "pk_indexer_append_records: offsets must ascend (entry %zu)", i);
    PK_REQUIRE(ix->nrec == 0 || ix->last_rec_start <= new_starts_host[0],
               "pk_indexer_append_records: offsets must ascend (first new entry)");
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    const size_t nrec = ix->nrec + n;
    if (nrec > ix->rec_cap) {
        PK_CUDA(cudaStreamSynchronize(ix->work_stream));   // kernels may still read the old table
        const size_t cap = std::max<size_t>(std::max<size_t>(nrec, 2 * ix->rec_cap), 1024);
        uint64_t *starts = nullptr;
        uint8_t *flags = nullptr;
        PK_CUDA(cudaMalloc(&starts, cap * sizeof(uint64_t)));
        cudaError_t e = cudaMalloc(&flags, cap);
        if (e == cudaSuccess) e = cudaMemset(flags, 0, cap);
        if (e == cudaSuccess && ix->nrec) {
            e = cudaMemcpy(flags, ix->rec_flags, ix->nrec, cudaMemcpyDeviceToDevice);
            if (e == cudaSuccess)
                e = cudaMemcpy(starts, ix->rec_starts, ix->nrec * sizeof(uint64_t), cudaMemcpyDeviceToDevice);
        }
        if (e != cudaSuccess) {
            cudaFree(starts); cudaFree(flags);
            return pk_set_error(PK_ERR_CUDA, "pk_indexer_append_records: %s", cudaGetErrorString(e));
        }
        cudaFree(ix->rec_starts); cudaFree(ix->rec_flags);
        ix->rec_starts = starts; ix->rec_flags = flags; ix->rec_cap = cap;
    }
    // entries beyond nrec are not read by kernels already queued (they were launched with the old count)
    PK_CUDA(cudaMemcpyAsync(ix->rec_starts + ix->nrec, new_starts_host, n * sizeof(uint64_t), cudaMemcpyHostToDevice,
                            ix->work_stream));
    PK_CUDA(cudaStreamSynchronize(ix->work_stream));       // the host array may go away
    ix->last_rec_start = new_starts_host[n - 1];
    ix->nrec = nrec;
    return PK_OK;
}

PK_API int pk_indexer_flush(pk_indexer *ix) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_flush: NULL handle");
    if (ix->mode != PK_MODE_PARTITION || (ix->nseg == 0 && ix->table_valid)) return PK_OK;   // nothing buffered
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    return indexer_flush(ix, ix->work_stream, false, nullptr);
}

PK_API int pk_indexer_feed_device(pk_indexer *ix, const uint8_t *seq_dev, size_t n, pk_stream stream) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_feed_device: NULL handle");
    PK_REQUIRE(n == 0 || seq_dev != nullptr, "pk_indexer_feed_device: NULL sequence");
    PK_REQUIRE(((uintptr_t)seq_dev & 15u) == 0, "pk_indexer_feed_device: seq_dev must be 16-byte aligned");
    pk_device_guard guard(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (st != ix->last_stream) {                          // keep feeds ordered across streams
        PK_CUDA(cudaEventRecord(ix->joined, ix->last_stream));
        PK_CUDA(cudaStreamWaitEvent(st, ix->joined, 0));
    }
    return indexer_feed_device(ix, seq_dev, n, st);
}

PK_API int pk_indexer_feed_host(pk_indexer *ix, const uint8_t *seq_host, size_t n) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_feed_host: NULL handle");
    PK_REQUIRE(n == 0 || seq_host != nullptr, "pk_indexer_feed_host: NULL sequence");
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    for (int i = 0; i < 2; i++)
        if (!ix->stage[i]) {
            PK_CUDA(cudaMalloc(&ix->stage[i], kStageBytes));
            PK_CUDA(cudaEventRecord(ix->consumed[i], ix->work_stream));
        }
    size_t off = 0;
    int buf = 0;
    while (off < n) {
        const size_t len = std::min(kStageBytes, n - off);
        PK_CUDA(cudaStreamWaitEvent(ix->copy_stream, ix->consumed[buf], 0));
        PK_CUDA(cudaMemcpyAsync(ix->stage[buf], seq_host + off, len, cudaMemcpyHostToDevice,
                                ix->copy_stream));
        PK_CUDA(cudaEventRecord(ix->copied[buf], ix->copy_stream));
        PK_CUDA(cudaStreamWaitEvent(ix->work_stream, ix->copied[buf], 0));
        const int rc = indexer_feed_device(ix, ix->stage[buf], len, ix->work_stream);
        if (rc != PK_OK) return rc;
        PK_CUDA(cudaEventRecord(ix->consumed[buf], ix->work_stream));
        off += len;
        buf ^= 1;
    }
    return PK_OK;
}

PK_API int pk_indexer_sync(pk_indexer *ix) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_sync: NULL handle");
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    PK_CUDA(cudaStreamSynchronize(ix->copy_stream));
    PK_CUDA(cudaStreamSynchronize(ix->work_stream));
    return PK_OK;
}

static int indexer_finalize(pk_indexer *ix, int64_t hist_host[255], uint64_t stats_host[5],
                            uint8_t *table_host) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_finalize: NULL handle");
    PK_REQUIRE(hist_host != nullptr && stats_host != nullptr, "pk_indexer_finalize: NULL output");
    if (ix->mode == PK_MODE_SCAN)
        return pk_set_error(PK_ERR_STATE, "pk_indexer_finalize: a scan-only handle has no table; export its "
                            "segments to the handles that own the windows");
    pk_device_guard guard(ix->device);
    bool copied = false;
    {
        const int rc = indexer_join(ix);                   // feeds may sit on the caller's stream
        if (rc != PK_OK) return rc;
    }
    cudaStream_t st = ix->work_stream;
    if (table_host) memset(ix->xfer, 0, sizeof ix->xfer);
    if (ix->mode == PK_MODE_PARTITION && (ix->nseg || !ix->table_valid || !ix->stats_valid)) {
        const int rc = indexer_flush(ix, st, true, table_host);   // statistics fused into the commit
        if (rc != PK_OK) return rc;
        copied = table_host != nullptr;
    }
    // DIRECT: counters + 1 already hold the histogram (kept as transitions by every update)
    if (table_host && !copied) {
        PK_CUDA(cudaMemcpyAsync(table_host, ix->table, ix->table_bytes, cudaMemcpyDeviceToHost, st));
        ix->xfer[0] += ix->table_bytes;
        ix->xfer[2] += 1;
    }
    PK_CUDA(cudaMemcpyAsync(ix->h_counters, ix->counters, 257 * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, st));
    PK_CUDA(cudaStreamSynchronize(st));
    uint64_t st4[4];
    stats_from_bins(ix->h_counters + 1, ix->table_bytes, hist_host, st4);
    stats_host[0] = ix->h_counters[0];
    for (int i = 0; i < 4; i++) stats_host[1 + i] = st4[i];
    return PK_OK;
}

PK_API int pk_indexer_finalize(pk_indexer *ix, int64_t hist_host[255], uint64_t stats_host[5]) {
    return indexer_finalize(ix, hist_host, stats_host, nullptr);
}

PK_API int pk_indexer_finalize_to_host(pk_indexer *ix, int64_t hist_host[255], uint64_t stats_host[5],
                                       uint8_t *table_host) {
    PK_REQUIRE(table_host != nullptr, "pk_indexer_finalize_to_host: NULL table buffer");
    return indexer_finalize(ix, hist_host, stats_host, table_host);
}

// the packing half of the packed transfer on its own (tests, other consumers): nz_dev needs room for
// n + n / 64 bytes; *nz_units_host = 16-byte units written.  Synchronises `stream`.
PK_API int pk_table_pack_device(const uint8_t *table_dev, size_t n, uint64_t *bitmap_dev, uint32_t *chunk_off_dev,
                                uint8_t *nz_dev, uint32_t *nz_units_host, pk_stream stream) {
    PK_REQUIRE(n % kPackChunk == 0, "pk_table_pack_device: %zu entries are not a multiple of the %zu-entry chunk", n, kPackChunk);
    PK_REQUIRE(nz_units_host != nullptr, "pk_table_pack_device: NULL output");
    *nz_units_host = 0;
    if (n == 0) return PK_OK;
    PK_REQUIRE(table_dev && bitmap_dev && chunk_off_dev && nz_dev, "pk_table_pack_device: NULL buffer");
    PK_REQUIRE(((uintptr_t)table_dev & 15) == 0 && ((uintptr_t)nz_dev & 15) == 0, "pk_table_pack_device: table and nz must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0;
    PK_CUDA(cudaGetDevice(&device));
    uint32_t *cursor = nullptr;
    PK_CUDA(cudaMalloc(&cursor, sizeof(uint32_t)));
    cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(uint32_t), st);
    const size_t nchunks = n / kPackChunk;
    if (e == cudaSuccess) {
        const int grid = (int)std::min<size_t>((size_t)pk_sm_count(device) * 8, (nchunks + 7) / 8);
        k_table_pack<<<grid, 256, 0, st>>>(table_dev, nchunks, reinterpret_cast<uint16_t *>(bitmap_dev), chunk_off_dev, nz_dev, cursor);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(nz_units_host, cursor, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(cursor);
    PK_CUDA(e);
    return PK_OK;
}

PK_API int pk_indexer_transfer_stats(pk_indexer *ix, uint64_t stats_host[4]) {
    PK_REQUIRE(ix != nullptr && stats_host != nullptr, "pk_indexer_transfer_stats: NULL argument");
    memcpy(stats_host, ix->xfer, sizeof ix->xfer);
    return PK_OK;
}

PK_API int pk_indexer_record_flags(pk_indexer *ix, uint8_t *flags_host, size_t nrec) {
    PK_REQUIRE(ix != nullptr && flags_host != nullptr, "pk_indexer_record_flags: NULL argument");
    PK_REQUIRE(nrec == ix->nrec, "pk_indexer_record_flags: %zu records registered, %zu asked",
               ix->nrec, nrec);
    if (nrec == 0) return PK_OK;
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    PK_CUDA(cudaMemcpyAsync(flags_host, ix->rec_flags, nrec, cudaMemcpyDeviceToHost, ix->work_stream));
    PK_CUDA(cudaStreamSynchronize(ix->work_stream));
    return PK_OK;
}

PK_API int pk_indexer_table_device(pk_indexer *ix, const uint8_t **table_dev, size_t *bytes) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_table_device: NULL handle");
    if (table_dev) *table_dev = ix->table;
    if (bytes) *bytes = ix->table_bytes;
    return PK_OK;
}

PK_API int pk_indexer_table_to_host(pk_indexer *ix, uint8_t *dst_host, size_t offset, size_t bytes) {
    PK_REQUIRE(ix != nullptr && dst_host != nullptr, "pk_indexer_table_to_host: NULL argument");
    PK_REQUIRE(offset <= ix->table_bytes && bytes <= ix->table_bytes - offset,
               "pk_indexer_table_to_host: [%zu, +%zu) outside %zu table bytes", offset, bytes,
               ix->table_bytes);
    if (ix->mode == PK_MODE_SCAN)
        return pk_set_error(PK_ERR_STATE, "pk_indexer_table_to_host: a scan-only handle has no table");
    if (ix->mode == PK_MODE_PARTITION && (ix->nseg || !ix->table_valid))
        return pk_set_error(PK_ERR_STATE, "pk_indexer_table_to_host: call pk_indexer_finalize first "
                            "(k-mers are still buffered)");
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    PK_CUDA(cudaMemcpyAsync(dst_host, ix->table + offset, bytes, cudaMemcpyDeviceToHost, ix->work_stream));
    PK_CUDA(cudaStreamSynchronize(ix->work_stream));
    return PK_OK;
}

PK_API int pk_indexer_launch_count(pk_indexer *ix, uint64_t *launches) {
    PK_REQUIRE(ix != nullptr && launches != nullptr, "pk_indexer_launch_count: NULL argument");
    *launches = ix->launches;
    return PK_OK;
}

// ---- sequence-sharded multi-GPU: scan anywhere, count where the window lives ----------------
PK_API int pk_indexer_prime(pk_indexer *ix, const uint8_t *halo_dev, size_t n, uint64_t stream_off,
                            pk_stream stream) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_prime: NULL handle");
    PK_REQUIRE(n <= (size_t)kCarry && (n == 0 || halo_dev != nullptr), "pk_indexer_prime: at most %d halo bytes", kCarry);
    pk_device_guard guard(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (st != ix->last_stream) {
        PK_CUDA(cudaEventRecord(ix->joined, ix->last_stream));
        PK_CUDA(cudaStreamWaitEvent(st, ix->joined, 0));
    }
    PK_CUDA(cudaMemsetAsync(ix->carry, 0, 64, st));
    if (n) {
        k_update_carry<<<1, 32, 0, st>>>(ix->carry, halo_dev, n, nullptr);
        PK_CUDA(cudaGetLastError());
        ix->launches += 1;
    }
    if (ix->mode == PK_MODE_SCAN && ix->nseg == 0) {
        // a scanner begins a new slice: its per-window counts / cursors start from zero again, while
        // num_kmers and the record flags keep accumulating (pk_indexer_reset clears those)
        PK_CUDA(cudaMemsetAsync(ix->seg, 0, (size_t)4 * kMaxSegments * ix->nbuckets * sizeof(uint32_t), st));
        PK_CUDA(cudaMemsetAsync(ix->cursor, 0, 32, st));
        ix->p1_seq = nullptr;
        ix->p1_n = 0;
        ix->fed = false;
    }
    ix->stream_off = stream_off;
    ix->last_stream = st;
    return PK_OK;
}

PK_API int pk_indexer_scan_result(pk_indexer *ix, uint64_t *num_kmers) {
    PK_REQUIRE(ix != nullptr && num_kmers != nullptr, "pk_indexer_scan_result: NULL argument");
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    PK_CUDA(cudaMemcpyAsync(ix->h_counters, ix->counters, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            ix->work_stream));
    PK_CUDA(cudaStreamSynchronize(ix->work_stream));
    *num_kmers = ix->h_counters[0];
    return PK_OK;
}

PK_API int pk_indexer_export_segments(pk_indexer *ix, const uint32_t **entries_dev, uint32_t *nseg,
                                      uint32_t *nwindows, uint32_t *seg_off_host, uint32_t *seg_cnt_host,
                                      size_t capacity) {
    PK_REQUIRE(ix != nullptr && entries_dev && nseg && nwindows, "pk_indexer_export_segments: NULL argument");
    PK_REQUIRE(ix->mode != PK_MODE_DIRECT, "pk_indexer_export_segments: DIRECT mode buffers no k-mers");
    const size_t need = (size_t)ix->nseg * ix->nbuckets;
    *entries_dev = ix->pool;
    *nseg = (uint32_t)ix->nseg;
    *nwindows = ix->nbuckets;
    if (!seg_off_host || !seg_cnt_host) return PK_OK;      // size query
    PK_REQUIRE(capacity >= need, "pk_indexer_export_segments: need room for %zu values, got %zu", need, capacity);
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    if (need) {
        PK_CUDA(cudaMemcpyAsync(seg_cnt_host, seg_cnt(ix, 0), need * sizeof(uint32_t), cudaMemcpyDeviceToHost, ix->work_stream));
        PK_CUDA(cudaMemcpyAsync(seg_off_host, seg_off(ix, 0), need * sizeof(uint32_t), cudaMemcpyDeviceToHost, ix->work_stream));
    }
    PK_CUDA(cudaStreamSynchronize(ix->work_stream));
    return PK_OK;
}

PK_API int pk_indexer_import_segments(pk_indexer *ix, const uint32_t *entries_dev, uint32_t nseg,
                                      const uint32_t *seg_off_host, const uint32_t *seg_cnt_host) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_import_segments: NULL handle");
    PK_REQUIRE(ix->mode == PK_MODE_PARTITION, "pk_indexer_import_segments: the handle must count in PARTITION mode");
    PK_REQUIRE(ix->nseg == 0, "pk_indexer_import_segments: the handle already buffers k-mers of its own");
    PK_REQUIRE(nseg <= (uint32_t)kMaxSegments, "pk_indexer_import_segments: at most %d segments", kMaxSegments);
    PK_REQUIRE(nseg == 0 || (seg_off_host && seg_cnt_host), "pk_indexer_import_segments: NULL argument");
    if (!entries_dev) entries_dev = ix->pool;             // peers stored the entries into this handle's pool
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    const size_t n = (size_t)nseg * ix->nbuckets;
    if (n) {
        PK_CUDA(cudaMemcpyAsync(seg_cnt(ix, 0), seg_cnt_host, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->work_stream));
        PK_CUDA(cudaMemcpyAsync(seg_off(ix, 0), seg_off_host, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->work_stream));
        PK_CUDA(cudaStreamSynchronize(ix->work_stream));   // the host tables may go away
    }
    ix->pool_ext = entries_dev;
    ix->nseg = (int)nseg;
    ix->stats_valid = false;
    return PK_OK;
}

// ---- fused exchange: pass 2 stores straight into the window owners' pools over NVLink --------
PK_API int pk_indexer_pool_ipc_handle(pk_indexer *ix, void *handle64, size_t *capacity_entries) {
    PK_REQUIRE(ix != nullptr && handle64 != nullptr, "pk_indexer_pool_ipc_handle: NULL argument");
    PK_REQUIRE(ix->pool != nullptr, "pk_indexer_pool_ipc_handle: the handle has no k-mer buffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    pk_device_guard guard(ix->device);
    cudaIpcMemHandle_t h;
    PK_CUDA(cudaIpcGetMemHandle(&h, ix->pool));
    memcpy(handle64, &h, 64);
    if (capacity_entries) *capacity_entries = ix->pool_cap;
    return PK_OK;
}

PK_API int pk_indexer_open_peer_pool(pk_indexer *scanner, int peer, const void *handle64,
                                     pk_indexer *local_owner) {
    PK_REQUIRE(scanner != nullptr && peer >= 0 && peer < 16, "pk_indexer_open_peer_pool: bad argument");
    PK_REQUIRE((handle64 != nullptr) != (local_owner != nullptr),
               "pk_indexer_open_peer_pool: give either an IPC handle or the local owner handle");
    pk_device_guard guard(scanner->device);
    if (scanner->peer_ipc[peer]) { cudaIpcCloseMemHandle(scanner->peer_ipc[peer]); scanner->peer_ipc[peer] = nullptr; }
    if (local_owner) {
        scanner->peer_pool[peer] = local_owner->pool;
    } else {
        cudaIpcMemHandle_t h;
        memcpy(&h, handle64, 64);
        void *base = nullptr;
        PK_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        scanner->peer_ipc[peer] = base;
        scanner->peer_pool[peer] = static_cast<uint32_t *>(base);
    }
    return PK_OK;
}

PK_API int pk_indexer_scan_pass1(pk_indexer *ix, const uint8_t *seq_dev, size_t n, pk_stream stream) {
    PK_REQUIRE(ix != nullptr && ix->mode == PK_MODE_SCAN, "pk_indexer_scan_pass1: needs a scan-only handle");
    PK_REQUIRE(seq_dev != nullptr && n > 0 && n <= kMaxFeed && n <= ix->pool_cap,
               "pk_indexer_scan_pass1: between 1 and 2^30 bases per pass");
    PK_REQUIRE(((uintptr_t)seq_dev & 15u) == 0, "pk_indexer_scan_pass1: seq_dev must be 16-byte aligned");
    PK_REQUIRE(ix->nseg == 0 && ix->p1_seq == nullptr, "pk_indexer_scan_pass1: a pass is already pending");
    pk_device_guard guard(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (st != ix->last_stream) {
        PK_CUDA(cudaEventRecord(ix->joined, ix->last_stream));
        PK_CUDA(cudaStreamWaitEvent(st, ix->joined, 0));
    }
    return indexer_launch_scan(ix, seq_dev, n, st, SCAN_PASS1);
}

PK_API int pk_indexer_pass1_counts(pk_indexer *ix, uint32_t *counts_host, size_t nwindows) {
    PK_REQUIRE(ix != nullptr && counts_host != nullptr, "pk_indexer_pass1_counts: NULL argument");
    PK_REQUIRE(ix->p1_seq != nullptr, "pk_indexer_pass1_counts: no pass 1 pending");
    PK_REQUIRE(nwindows == ix->nbuckets, "pk_indexer_pass1_counts: the handle has %u windows", ix->nbuckets);
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    PK_CUDA(cudaMemcpyAsync(counts_host, seg_cnt(ix, 0), nwindows * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                            ix->work_stream));
    PK_CUDA(cudaStreamSynchronize(ix->work_stream));
    return PK_OK;
}

PK_API int pk_indexer_scan_pass2_remote(pk_indexer *ix, int nranks, const uint32_t *owner_host,
                                        const uint32_t *dest_off_host, pk_stream stream) {
    PK_REQUIRE(ix != nullptr && owner_host != nullptr && dest_off_host != nullptr,
               "pk_indexer_scan_pass2_remote: NULL argument");
    PK_REQUIRE(ix->p1_seq != nullptr, "pk_indexer_scan_pass2_remote: no pass 1 pending");
    PK_REQUIRE(nranks >= 1 && nranks <= 16, "pk_indexer_scan_pass2_remote: 1..16 ranks");
    for (uint32_t b = 0; b < ix->nbuckets; b++) {
        PK_REQUIRE(owner_host[b] < (uint32_t)nranks && ix->peer_pool[owner_host[b]] != nullptr,
                   "pk_indexer_scan_pass2_remote: window %u routed to rank %u whose pool is not open", b, owner_host[b]);
    }
    pk_device_guard guard(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (st != ix->last_stream) {
        PK_CUDA(cudaEventRecord(ix->joined, ix->last_stream));
        PK_CUDA(cudaStreamWaitEvent(st, ix->joined, 0));
    }
    if (!ix->route) PK_CUDA(cudaMalloc(&ix->route, (size_t)3 * kMaxBuckets * sizeof(uint32_t)));
    PK_CUDA(cudaMemcpyAsync(ix->route, owner_host, ix->nbuckets * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PK_CUDA(cudaMemcpyAsync(ix->route + kMaxBuckets, dest_off_host, ix->nbuckets * sizeof(uint32_t),
                            cudaMemcpyHostToDevice, st));
    ix->routed = false;                        // the routing of a routed scan, if any, has been replaced
    const uint8_t *seq = ix->p1_seq;
    const size_t n = ix->p1_n;
    ix->p1_seq = nullptr;
    ix->p1_n = 0;
    const int rc = indexer_launch_scan(ix, seq, n, st, SCAN_PASS2_REMOTE);
    if (rc != PK_OK) return rc;
    PK_CUDA(cudaStreamSynchronize(st));        // the pageable routing tables may go away; stores have landed
    return PK_OK;
}

// ---- routed scan: fixed regions, no host round trip per step -------------------------------------
PK_API int pk_indexer_pub_base(pk_indexer *ix, uint64_t *entry_index) {
    PK_REQUIRE(ix != nullptr && entry_index != nullptr, "pk_indexer_pub_base: NULL argument");
    PK_REQUIRE(ix->pool != nullptr && ix->pool_cap > 2 * kPubEntries, "pk_indexer_pub_base: the handle's k-mer buffer is too small");
    *entry_index = ix->pool_cap - kPubEntries;
    return PK_OK;
}

PK_API int pk_indexer_set_route(pk_indexer *ix, int nranks, int self_rank, const uint32_t *owner_host,
                                const uint32_t *dest_off_host, const uint32_t *cap_host,
                                const uint64_t *pub_base_host) {
    PK_REQUIRE(ix != nullptr && ix->mode == PK_MODE_SCAN, "pk_indexer_set_route: needs a scan-only handle");
    PK_REQUIRE(owner_host && dest_off_host && cap_host && pub_base_host, "pk_indexer_set_route: NULL argument");
    PK_REQUIRE(nranks >= 1 && nranks <= 16 && self_rank >= 0 && self_rank < nranks, "pk_indexer_set_route: rank %d of %d", self_rank, nranks);
    PK_REQUIRE((size_t)nranks * ix->nbuckets <= kPubEntries, "pk_indexer_set_route: %d ranks x %u windows exceed the published-count table", nranks, ix->nbuckets);
    for (uint32_t b = 0; b < ix->nbuckets; b++) {
        PK_REQUIRE(owner_host[b] < (uint32_t)nranks && ix->peer_pool[owner_host[b]] != nullptr,
                   "pk_indexer_set_route: window %u routed to rank %u whose buffer is not open", b, owner_host[b]);
        PK_REQUIRE((uint64_t)dest_off_host[b] + cap_host[b] <= pub_base_host[owner_host[b]],
                   "pk_indexer_set_route: the region of window %u runs into the published-count table", b);
    }
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    PK_CUDA(cudaStreamSynchronize(ix->work_stream));
    if (!ix->route) PK_CUDA(cudaMalloc(&ix->route, (size_t)3 * kMaxBuckets * sizeof(uint32_t)));
    PK_CUDA(cudaMemcpy(ix->route, owner_host, ix->nbuckets * sizeof(uint32_t), cudaMemcpyHostToDevice));
    PK_CUDA(cudaMemcpy(ix->route + kMaxBuckets, dest_off_host, ix->nbuckets * sizeof(uint32_t), cudaMemcpyHostToDevice));
    PK_CUDA(cudaMemcpy(ix->route + 2 * kMaxBuckets, cap_host, ix->nbuckets * sizeof(uint32_t), cudaMemcpyHostToDevice));
    for (int i = 0; i < 16; i++) ix->pub_base[i] = i < nranks ? pub_base_host[i] : 0;
    ix->self_rank = self_rank;
    ix->routed = true;
    return PK_OK;
}

PK_API int pk_indexer_scan_routed(pk_indexer *ix, const uint8_t *seq_dev, size_t n, uint32_t *status_dev,
                                  pk_stream stream) {
    PK_REQUIRE(ix != nullptr && ix->mode == PK_MODE_SCAN && ix->routed, "pk_indexer_scan_routed: needs a scan-only handle with a route (pk_indexer_set_route)");
    PK_REQUIRE(seq_dev != nullptr && n > 0 && n <= kMaxFeed, "pk_indexer_scan_routed: between 1 and 2^30 bases per pass");
    PK_REQUIRE(((uintptr_t)seq_dev & 15u) == 0, "pk_indexer_scan_routed: seq_dev must be 16-byte aligned");
    PK_REQUIRE(ix->nseg == 0 && ix->p1_seq == nullptr && !ix->fed, "pk_indexer_scan_routed: one routed scan per reset");
    pk_device_guard guard(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (st != ix->last_stream) {
        PK_CUDA(cudaEventRecord(ix->joined, ix->last_stream));
        PK_CUDA(cudaStreamWaitEvent(st, ix->joined, 0));
    }
    ix->routed_status = status_dev;
    const int rc = indexer_launch_scan(ix, seq_dev, n, st, SCAN_ROUTED);
    ix->routed_status = nullptr;
    return rc;
}

PK_API int pk_indexer_set_import_layout(pk_indexer *ix, uint32_t nseg, const uint32_t *seg_off_host,
                                        uint32_t first_window, uint32_t nwindows_total) {
    PK_REQUIRE(ix != nullptr && ix->mode == PK_MODE_PARTITION, "pk_indexer_set_import_layout: the handle must count in PARTITION mode");
    PK_REQUIRE(nseg >= 1 && nseg <= 16 && seg_off_host != nullptr, "pk_indexer_set_import_layout: 1..16 source ranks");
    PK_REQUIRE((size_t)nseg * nwindows_total <= kPubEntries && first_window + ix->nbuckets <= nwindows_total,
               "pk_indexer_set_import_layout: windows [%u, +%u) of %u", first_window, ix->nbuckets, nwindows_total);
    PK_REQUIRE(ix->pool_cap > 2 * kPubEntries, "pk_indexer_set_import_layout: the handle's k-mer buffer is too small");
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);
        if (rc != PK_OK) return rc;
    }
    PK_CUDA(cudaStreamSynchronize(ix->work_stream));
    cudaFree(ix->import_off);
    ix->import_off = nullptr;
    const size_t n = (size_t)nseg * ix->nbuckets;
    PK_CUDA(cudaMalloc(&ix->import_off, n * sizeof(uint32_t)));
    PK_CUDA(cudaMemcpy(ix->import_off, seg_off_host, n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    ix->import_nseg = nseg; ix->import_w0 = first_window; ix->import_nb_total = nwindows_total;
    return PK_OK;
}

PK_API int pk_indexer_import_published(pk_indexer *ix, pk_stream stream) {
    PK_REQUIRE(ix != nullptr && ix->import_off != nullptr, "pk_indexer_import_published: no import layout (pk_indexer_set_import_layout)");
    PK_REQUIRE(ix->nseg == 0, "pk_indexer_import_published: the handle already buffers k-mers of its own");
    pk_device_guard guard(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (st != ix->last_stream) {
        PK_CUDA(cudaEventRecord(ix->joined, ix->last_stream));
        PK_CUDA(cudaStreamWaitEvent(st, ix->joined, 0));
    }
    const uint32_t n = ix->import_nseg * ix->nbuckets;
    k_adopt_published<<<(n + 255) / 256, 256, 0, st>>>(ix->pool + (ix->pool_cap - kPubEntries), ix->import_nseg,
                                                       ix->import_nb_total, ix->import_w0, ix->nbuckets, ix->import_off,
                                                       seg_cnt(ix, 0), seg_off(ix, 0));
    PK_CUDA(cudaGetLastError());
    ix->launches += 1;
    ix->pool_ext = ix->pool;
    ix->nseg = (int)ix->import_nseg;
    ix->stats_valid = false;
    ix->last_stream = st;
    return PK_OK;
}

static void prof_clear(pk_indexer *ix) {
    for (cudaEvent_t e : ix->prof_events) cudaEventDestroy(e);
    ix->prof_events.clear();
    ix->prof_tags.clear();
}

PK_API int pk_indexer_set_profiling(pk_indexer *ix, int enable) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_set_profiling: NULL handle");
    pk_device_guard guard(ix->device);
    prof_clear(ix);
    ix->profiling = enable != 0;
    return PK_OK;
}

PK_API int pk_indexer_profile(pk_indexer *ix, double ms_host[8], uint32_t launches_host[8]) {
    PK_REQUIRE(ix != nullptr && ms_host != nullptr && launches_host != nullptr,
               "pk_indexer_profile: NULL argument");
    pk_device_guard guard(ix->device);
    PK_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < 8; i++) { ms_host[i] = 0.0; launches_host[i] = 0; }
    for (size_t i = 0; i < ix->prof_tags.size(); i++) {
        float ms = 0.f;
        PK_CUDA(cudaEventElapsedTime(&ms, ix->prof_events[2 * i], ix->prof_events[2 * i + 1]));
        ms_host[ix->prof_tags[i]] += ms;
        launches_host[ix->prof_tags[i]] += 1;
        if (const char *v = getenv("PYKMER_B200_VERBOSE"))
            if (atoi(v) >= 2) fprintf(stderr, "[pykmer_b200] launch %zu class %d %.4f ms\n", i, ix->prof_tags[i], ms);
    }
    prof_clear(ix);
    return PK_OK;
}

PK_API int pk_indexer_mode(pk_indexer *ix, int *mode, int *windows) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_mode: NULL handle");
    if (mode) *mode = ix->mode;
    if (windows) *windows = ix->mode != PK_MODE_DIRECT ? (int)ix->nbuckets : 0;
    return PK_OK;
}

PK_API int pk_indexer_window_log2(pk_indexer *ix, int *window_log2) {
    PK_REQUIRE(ix != nullptr && window_log2 != nullptr, "pk_indexer_window_log2: NULL argument");
    *window_log2 = ix->mode != PK_MODE_DIRECT ? (int)ix->win_log2 : 0;
    return PK_OK;
}

PK_API int pk_table_stats_device(const uint8_t *table_dev, size_t n, int64_t hist_host[255],
                                 uint64_t stats_host[4], pk_stream stream) {
    PK_REQUIRE(hist_host != nullptr && stats_host != nullptr, "pk_table_stats_device: NULL output");
    PK_REQUIRE(n == 0 || table_dev != nullptr, "pk_table_stats_device: NULL table");
    PK_REQUIRE(((uintptr_t)table_dev & 15u) == 0, "pk_table_stats_device: table_dev must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int device = 0;
    PK_CUDA(cudaGetDevice(&device));
    unsigned long long *bins = nullptr;
    unsigned long long h[256];
    PK_CUDA(cudaMalloc(&bins, 256 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(bins, 0, 256 * sizeof(unsigned long long), st);
    if (e == cudaSuccess && n) {
        k_table_stats<<<pk_sm_count(device) * 8, 256, 0, st>>>(table_dev, n, bins);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, bins, sizeof h, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(bins);
    if (e != cudaSuccess) return pk_set_error(PK_ERR_CUDA, "pk_table_stats_device: %s", cudaGetErrorString(e));
    return stats_from_bins(h, n, hist_host, stats_host);
}

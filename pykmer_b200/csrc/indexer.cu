// indexer.cu -- the indexer hot path on B200 (sm_100a).
//
// Replaces, from the reference (sauloal/pykmer): CONV (indexer.py:36-41),
// gen_kmers (indexer.py:130-160), pos = min(fwd, rev) / num_kmers
// (indexer.py:341-342), the `chromosomes` rule (indexer.py:349-351),
// process_kmers' saturating accumulate (indexer.py:239,262) and
// Header.update_stats (tools.py:246-263).
//
// Kernels
//   k_scan_count_direct<WIDE>  fused encode -> canonical k-mer -> saturating count.
//       One thread encodes one 16-base group (one 16-byte coalesced load), gets the
//       K-1 base halo from its neighbour lanes by warp shuffle, slices every window
//       out of the 2-bit concatenation (kmer_bits.h) and bumps table[canon] with a
//       byte-granular compare-and-swap that never touches a saturated counter.
//   k_table_stats              one streaming pass: 256-bin histogram -> hist/vals_*.
//   k_update_carry             keeps the last 32 stream bytes for the next feed.
#include <algorithm>
#include <new>
#include <string.h>

#include "common.h"
#include "kmer_bits.h"

namespace {

constexpr int kScanThreads = 256;
constexpr size_t kStageBytes = 32u << 20;   // pinned-copy chunk of feed_host
constexpr int kCarry = 32;                  // bytes of stream tail kept between feeds

struct ScanParams {
    const uint8_t *seq;       // this feed (16-byte aligned)
    size_t n;
    const uint8_t *carry;     // last kCarry bytes of everything fed before
    int K;
    uint64_t lo, hi;          // canonical range owned by this handle
    uint8_t *table;           // [hi - lo]
    unsigned long long *num_kmers;
    const uint64_t *rec_starts;
    size_t nrec;
    uint8_t *rec_flags;
    uint64_t stream_off;      // stream offset of seq[0]
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

// table[idx] = min(255, table[idx] + cnt)  (indexer.py:239,262).  Counters only
// grow, so a (possibly stale) read of 255 is final and needs no atomic at all.
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

__device__ __forceinline__ long long find_record(const uint64_t *starts, size_t nrec, uint64_t pos) {
    size_t lo = 0, hi = nrec;                            // upper_bound - 1
    while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if (__ldg(starts + mid) <= pos) lo = mid + 1; else hi = mid;
    }
    return (long long)lo - 1;
}

template <bool WIDE>
__global__ void __launch_bounds__(kScanThreads) k_scan_count_direct(const ScanParams p) {
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = (uint8_t)pk_lut_entry(threadIdx.x);
    __syncthreads();

    constexpr int H = WIDE ? 2 : 1;          // halo groups: K-1 <= 16*H bases
    constexpr int GPW = 32 - H;              // groups a warp emits per tile
    const unsigned full = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int K = p.K;
    const long long ngroups = (long long)((p.n + 15) / 16);
    const long long ntiles = (ngroups + GPW - 1) / GPW;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const uint64_t mask64 = pk_kmer_mask(K);
    const uint32_t mask32 = (uint32_t)mask64;
    unsigned long long counted = 0;

    for (long long tile = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < ntiles;
         tile += nwarps) {
        const long long g = tile * GPW - H + lane;
        uint32_t w[4];
        load_group(p, g, ngroups, w);
        uint32_t cc, cv;
        pk_encode16(w, lut, cc, cv);
        const uint32_t pc1 = __shfl_up_sync(full, cc, 1);
        const uint32_t pv1 = __shfl_up_sync(full, cv, 1);
        uint32_t pc2 = 0, pv2 = 0;
        if (WIDE) {
            pc2 = __shfl_up_sync(full, cc, 2);
            pv2 = __shfl_up_sync(full, cv, 2);
        }
        if (lane < H || g >= ngroups) continue;

        const uint64_t vcat = ((uint64_t)pv2 << 32) | ((uint64_t)pv1 << 16) | cv;
        const uint32_t Wm = (uint32_t)pk_valid_windows(vcat, K) & 0xFFFFu;
        if (!Wm) continue;

        const uint32_t r0 = pk_rcw(cc), r1 = pk_rcw(pc1), r2 = WIDE ? pk_rcw(pc2) : 0u;
        const uint64_t cat = ((uint64_t)pc1 << 32) | cc;
        const uint64_t rcat = ((uint64_t)r0 << 32) | r1;
        uint64_t prev = ~0ull;
        uint32_t pend = 0;
        long long cur_rec = -1;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (!((cv >> (15 - j)) & 1u)) cur_rec = -1;      // a separator may have passed
            if (!((Wm >> (15 - j)) & 1u)) continue;
            uint64_t canon;
            if (WIDE) {
                const uint64_t f = pk_fwd_at(pc2, pc1, cc, j, K);
                const uint64_t r = pk_rc_at(r2, r1, r0, j, K);
                canon = f < r ? f : r;                        // indexer.py:341
            } else {
                const uint32_t f = pk_fwd32_at(cat, j, mask32);
                const uint32_t r = pk_rc32_at(rcat, j, K, mask32);
                canon = f < r ? f : r;
            }
            if (canon < p.lo || canon >= p.hi) continue;      // another shard's k-mer
            counted++;                                        // indexer.py:342
            if (canon == prev) {
                pend++;                                       // homopolymer run: one update
            } else {
                if (pend) sat_add_u8(p.table, prev - p.lo, pend);
                prev = canon;
                pend = 1;
            }
            if (p.rec_flags && cur_rec < 0) {                 // indexer.py:349-351
                cur_rec = find_record(p.rec_starts, p.nrec, p.stream_off + (uint64_t)g * 16 + j);
                if (cur_rec >= 0 && !p.rec_flags[cur_rec]) p.rec_flags[cur_rec] = 1;
            }
        }
        if (pend) sat_add_u8(p.table, prev - p.lo, pend);
    }
    (void)mask64;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) counted += __shfl_xor_sync(full, counted, o);
    if (lane == 0 && counted) atomicAdd(p.num_kmers, counted);
}

// new carry = last kCarry bytes of (old carry ++ seq[0..n))
__global__ void k_update_carry(uint8_t *carry, const uint8_t *seq, size_t n) {
    const int i = threadIdx.x;                            // 32 threads
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

}  // namespace

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
    uint8_t *stage[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr, work_stream = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    cudaEvent_t joined = nullptr;              // orders the caller's stream against work_stream
    cudaStream_t last_stream = nullptr;        // stream of the most recent feed / reset
    int sm_count = 148;
    uint64_t launches = 0;
    bool fed = false;
};

static int indexer_launch_scan(pk_indexer *ix, const uint8_t *seq_dev, size_t n, cudaStream_t st) {
    if (n == 0) return PK_OK;
    ScanParams p;
    p.seq = seq_dev; p.n = n; p.carry = ix->carry; p.K = ix->K; p.lo = ix->lo; p.hi = ix->hi;
    p.table = ix->table; p.num_kmers = ix->counters;
    p.rec_starts = ix->rec_starts; p.nrec = ix->nrec; p.rec_flags = ix->nrec ? ix->rec_flags : nullptr;
    p.stream_off = ix->stream_off;
    const bool wide = ix->K > 16;
    const long long gpw = wide ? 30 : 31;
    const long long ngroups = (long long)((n + 15) / 16);
    const long long ntiles = (ngroups + gpw - 1) / gpw;
    const long long want = (ntiles + (kScanThreads / 32) - 1) / (kScanThreads / 32);
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)ix->sm_count * 8));
    if (wide) k_scan_count_direct<true><<<grid, kScanThreads, 0, st>>>(p);
    else      k_scan_count_direct<false><<<grid, kScanThreads, 0, st>>>(p);
    PK_CUDA(cudaGetLastError());
    k_update_carry<<<1, 32, 0, st>>>(ix->carry, seq_dev, n);
    PK_CUDA(cudaGetLastError());
    ix->launches += 2;
    ix->stream_off += n;
    ix->fed = true;
    ix->last_stream = st;
    return PK_OK;
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
    PK_REQUIRE(mode == PK_MODE_AUTO || mode == PK_MODE_DIRECT,
               "pk_indexer_create: mode %d not available in this build", mode);
    int ndev = 0;
    PK_CUDA(cudaGetDeviceCount(&ndev));
    PK_REQUIRE(device >= 0 && device < ndev, "pk_indexer_create: device %d of %d", device, ndev);
    pk_device_guard guard(device);
    if (!guard.ok) return pk_set_error(PK_ERR_CUDA, "cudaSetDevice(%d) failed", device);

    pk_indexer *ix = new (std::nothrow) pk_indexer();
    if (!ix) return pk_set_error(PK_ERR_NOMEM, "out of host memory");
    ix->K = kmer_len; ix->device = device; ix->lo = range_lo; ix->hi = range_hi;
    ix->mode = PK_MODE_DIRECT;
    ix->table_bytes = (size_t)(range_hi - range_lo);
    ix->sm_count = pk_sm_count(device);
    const size_t alloc = (ix->table_bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    step(cudaMalloc(&ix->table, alloc));
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
    if (e == cudaSuccess) {
        step(cudaMemsetAsync(ix->table, 0, alloc, ix->work_stream));
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
    if (ix->h_counters) cudaFreeHost(ix->h_counters);
    for (int i = 0; i < 2; i++) {
        if (ix->copied[i]) cudaEventDestroy(ix->copied[i]);
        if (ix->consumed[i]) cudaEventDestroy(ix->consumed[i]);
    }
    if (ix->joined) cudaEventDestroy(ix->joined);
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
    PK_CUDA(cudaMemsetAsync(ix->table, 0, ix->table_bytes, st));
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
    return PK_OK;
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
    return indexer_launch_scan(ix, seq_dev, n, st);
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
        const int rc = indexer_launch_scan(ix, ix->stage[buf], len, ix->work_stream);
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

PK_API int pk_indexer_finalize(pk_indexer *ix, int64_t hist_host[255], uint64_t stats_host[5]) {
    PK_REQUIRE(ix != nullptr, "pk_indexer_finalize: NULL handle");
    PK_REQUIRE(hist_host != nullptr && stats_host != nullptr, "pk_indexer_finalize: NULL output");
    pk_device_guard guard(ix->device);
    {
        const int rc = indexer_join(ix);                   // feeds may sit on the caller's stream
        if (rc != PK_OK) return rc;
    }
    cudaStream_t st = ix->work_stream;
    PK_CUDA(cudaMemsetAsync(ix->counters + 1, 0, 256 * sizeof(unsigned long long), st));
    k_table_stats<<<ix->sm_count * 8, 256, 0, st>>>(ix->table, ix->table_bytes, ix->counters + 1);
    PK_CUDA(cudaGetLastError());
    ix->launches += 1;
    PK_CUDA(cudaMemcpyAsync(ix->h_counters, ix->counters, 257 * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, st));
    PK_CUDA(cudaStreamSynchronize(st));
    uint64_t st4[4];
    stats_from_bins(ix->h_counters + 1, ix->table_bytes, hist_host, st4);
    stats_host[0] = ix->h_counters[0];
    for (int i = 0; i < 4; i++) stats_host[1 + i] = st4[i];
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

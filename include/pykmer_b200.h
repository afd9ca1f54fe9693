/*
 * pykmer_b200.h -- C ABI of libpykmer_b200.so, the B200 (sm_100a) replacement for
 * the two data-parallel hot paths of sauloal/pykmer.
 *
 * The reference is pure Python and exposes no FFI; the drop-in boundary is its
 * two CLIs and four file formats.  This header is the seam
 * a maintainer binds with ctypes at the reference's two natural function seams;
 * INTEGRATION.md shows the stub.  Each entry point cites the reference code it
 * replaces (file:line into sauloal/pykmer).
 *
 * Conventions: plain C symbols; every function returns PK_OK (0) or a negative
 * PK_ERR_* and leaves a message for pk_last_error() (thread-local); nothing is
 * thrown or allocated across the ABI except through the create/alloc calls
 * below; `*_dev` pointers are device pointers on the handle's / current device,
 * `*_host` pointers are host pointers; pk_stream is a cudaStream_t (NULL = the
 * legacy default stream).  A pk_indexer belongs to one GPU and is not
 * thread-safe; independent handles are independent.
 */
#ifndef PYKMER_B200_H
#define PYKMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PK_OK          0
#define PK_ERR_ARG    -1   /* bad argument (the reference raises AssertionError/ValueError) */
#define PK_ERR_CUDA   -2   /* CUDA runtime failure, text in pk_last_error() */
#define PK_ERR_STATE  -3   /* call out of order (e.g. feed after finalize) */
#define PK_ERR_NOMEM  -4   /* device or pinned-host allocation failed */

#define PK_ABI_VERSION 1

typedef void *pk_stream;
typedef struct pk_indexer pk_indexer;

/* ------------------------------------------------------------------ general */
int         pk_abi_version(void);
const char *pk_last_error(void);
int         pk_device_count(int *count);
/* name_len bytes of name are filled (NUL-terminated); any out pointer may be NULL */
int         pk_device_info(int device, char *name, size_t name_len, int *sm_count,
                           size_t *total_mem_bytes, int *cc_major, int *cc_minor);
/* pinned (page-locked) host memory for the streaming copies */
int         pk_host_alloc(void **ptr, size_t bytes);
int         pk_host_free(void *ptr);

/* ------------------------------------------------------------------ indexer
 * Replaces gen_kmers + the consumer loop of create_fasta_index + process_kmers
 * (indexer.py:130-160, 340-384, 162-297) and Header.update_stats
 * (tools.py:246-263).
 *
 * The sequence is fed as a cleaned byte STREAM: the host has applied the
 * reference's text rules (indexer.py:55-95) and joined the records with one
 * byte outside ACGTacgt.  Any such byte voids the K windows containing it
 * (indexer.py:144), so the separator also keeps windows inside records
 * (indexer.py:133-141).  Feeds concatenate: the last K-1 bases carry over.
 *
 * The handle owns table[range_hi - range_lo] (uint8, saturating at 255,
 * indexer.py:239,262) covering canonical k-mer values [range_lo, range_hi);
 * windows whose canonical value falls outside are ignored.  range 0..4^K is the
 * whole .kin; a sub-range is one shard of the k-mer-axis partition.
 */
#define PK_MODE_AUTO      0   /* PARTITION for 11 <= K <= 17, else DIRECT */
#define PK_MODE_DIRECT    1   /* saturating byte compare-and-swap straight into the table */
#define PK_MODE_PARTITION 2   /* bucket k-mers by table window, count each window in L2 and write
                                 it once: 2^24-entry windows with 32-bit counters for K <= 15,
                                 2^26-entry windows holding the table's own 8-bit lanes (carries
                                 settled exactly) for K >= 17; pk_indexer_window_log2 tells.
                                 (PYKMER_B200_WINDOW_LOG2 / PYKMER_B200_POOL_LOG2 /
                                 PYKMER_B200_OVF_LOG2 shrink the window, the k-mer buffer and the
                                 carry table, PYKMER_B200_FLUSH = l2 | byte | smem forces the
                                 scheme; they exist for the tests.) */

#define PK_MODE_SCAN      3   /* scan + bucket only (no table): the scanning half of the
                                 sequence-sharded multi-GPU path, see pk_indexer_export_segments */

int pk_indexer_create(pk_indexer **out, int kmer_len, int device,
                      uint64_t range_lo, uint64_t range_hi, int mode);
int pk_indexer_destroy(pk_indexer *ix);
/* zero the table, counters, carry and stream offset; keeps allocations */
int pk_indexer_reset(pk_indexer *ix, pk_stream stream);

/* Optional record table (stream offsets of each record's first byte, ascending).
 * When set, pk_indexer_record_flags reports which records produced >= 1 counted
 * window -- the rule by which the reference lists `chromosomes`
 * (indexer.py:349-351).  May be called again between feeds with a LONGER table
 * (records are discovered as the file is parsed); entries already registered
 * must not change. */
int pk_indexer_set_records(pk_indexer *ix, const uint64_t *rec_starts_host, size_t nrec);
/* The same table grown by n more records: only the new offsets are checked and uploaded (a draft
 * assembly or a read set has millions of records and arrives in hundreds of pieces). */
int pk_indexer_append_records(pk_indexer *ix, const uint64_t *new_starts_host, size_t n);

/* seq_dev must be 16-byte aligned.  Asynchronous on `stream`. */
int pk_indexer_feed_device(pk_indexer *ix, const uint8_t *seq_dev, size_t n, pk_stream stream);
/* Streams seq_host through double-buffered cudaMemcpyAsync on a side stream,
 * overlapping copy and counting.  Pinned memory (pk_host_alloc) gives full PCIe
 * speed.  Returns after the last chunk is enqueued; the buffer must stay
 * untouched until pk_indexer_sync / pk_indexer_finalize. */
int pk_indexer_feed_host(pk_indexer *ix, const uint8_t *seq_host, size_t n);
int pk_indexer_sync(pk_indexer *ix);
/* PARTITION mode: count everything buffered so far into the table now (asynchronous; statistics are
 * left to pk_indexer_finalize).  Lets pk_indexer_import_segments be called once per piece of a stream. */
int pk_indexer_flush(pk_indexer *ix);

/* Finish counting and compute the statistics of tools.py:246-263 over the
 * handle's range.  hist_host[i] = #{table == i+1}, i in 0..254.
 * stats_host = {num_kmers (indexer.py:342), vals_sum, vals_count, vals_min,
 * vals_max}.  Synchronises.  Feeding after finalize is allowed (statistics are
 * recomputed by the next finalize).  In PARTITION mode the table is only
 * complete after finalize. */
int pk_indexer_finalize(pk_indexer *ix, int64_t hist_host[255], uint64_t stats_host[5]);
/* Same, and the whole table (range_hi - range_lo bytes) lands in table_host (pinned
 * memory for full speed).  In PARTITION mode each table window leaves as soon as it is
 * committed -- packed, see below -- so the transfer overlaps the counting of the later windows. */
int pk_indexer_finalize_to_host(pk_indexer *ix, int64_t hist_host[255], uint64_t stats_host[5],
                                uint8_t *table_host);
/* what the last pk_indexer_finalize_to_host moved: {bytes device-to-host, windows sent packed, windows sent
 * as they are, host threads that rebuilt the packed ones} */
int pk_indexer_transfer_stats(pk_indexer *ix, uint64_t stats_host[4]);
int pk_indexer_record_flags(pk_indexer *ix, uint8_t *flags_host, size_t nrec);

/* The packed form in which pk_indexer_finalize_to_host moves a finished table over PCIe (the .kin bytes of
 * tools.py:196,240-243 are mostly zeros: 25 % of the entries are in use at K=15, 3 % at K=17), per slice of
 * n entries, n a multiple of 1024:
 *   bitmap[n / 64]       bit i of word j <=> entry 64 j + i is non-zero
 *   chunk_off[n / 1024]  where the non-zero bytes of entries [1024 c, +1024) start in nz, in 16-byte units
 *   nz                   those bytes in entry order, every chunk padded to a multiple of 16 (chunks in any order)
 * finalize_to_host packs every window behind its last kernel, copies only the packed form into pinned slots
 * and rebuilds the bytes in table_host on the host's cores while later windows are counted; dense windows
 * (packed size above 5/8 of the bytes), and windows whose slot is still being rebuilt, are copied as they
 * are.  With more than two ranks per host (LOCAL_WORLD_SIZE) only windows that pack to a quarter of their bytes go
 * packed.  PYKMER_B200_PACKED_D2H=0 turns the packed form off, =1 keeps the 5/8 rule whatever the number of ranks;
 * PYKMER_B200_UNPACK_THREADS sets the team size (default: all cores but one, divided by LOCAL_WORLD_SIZE).
 * The two halves on their own: */
int pk_table_pack_device(const uint8_t *table_dev, size_t n, uint64_t *bitmap_dev, uint32_t *chunk_off_dev,
                         uint8_t *nz_dev /* room for n + n / 64 bytes */, uint32_t *nz_units_host, pk_stream stream);
/* host only (no CUDA call): dst[0, n) from the packed form; offsets are checked against nz_bytes */
int pk_table_unpack(const uint64_t *bitmap, const uint32_t *chunk_off, const uint8_t *nz, size_t nz_bytes,
                    size_t n, uint8_t *dst, int threads);

/* Device view of the table (valid after finalize), and a copy to host memory. */
int pk_indexer_table_device(pk_indexer *ix, const uint8_t **table_dev, size_t *bytes);
int pk_indexer_table_to_host(pk_indexer *ix, uint8_t *dst_host, size_t offset, size_t bytes);
/* launches of this library's kernels issued through the handle so far */
int pk_indexer_launch_count(pk_indexer *ix, uint64_t *launches);
/* ---- sequence-sharded multi-GPU indexing -----------------------------------------------
 * Instead of every GPU scanning the whole sequence, each rank scans 1/N of it with a
 * PK_MODE_SCAN handle over the FULL k-mer range, the bucketed k-mer entries are exchanged
 * (all-to-all over NVLink, done by the caller, e.g. torch.distributed / NCCL) so that every
 * entry reaches the rank that owns its table window, and a PK_MODE_PARTITION handle over the
 * rank's own k-mer range counts them.  Window w of the scanner covers k-mers
 * [range_lo + w * 2^24, +2^24); shard boundaries must be multiples of 2^24.
 *
 * pk_indexer_prime: start a slice in the middle of the stream -- the up-to-32 bytes that
 *   precede it become the window carry (no k-mer is counted for them) and stream_off is the
 *   stream offset of the next byte fed (record flags use it).  On a scan-only handle with nothing
 *   buffered it also starts the per-window counts from zero again (num_kmers and the record flags
 *   keep accumulating), so one scanner can take a slice of every piece of a long stream.
 * pk_indexer_export_segments: device pointer of the entry buffer and, per fed segment and
 *   window, the offset and count of its entries (host arrays of nseg * nwindows uint32;
 *   pass NULL arrays to query nseg / nwindows).  Synchronises.
 * pk_indexer_import_segments: hand a PARTITION handle entries gathered elsewhere: segment s,
 *   local window w has seg_cnt[s * nwindows + w] entries at entries_dev + seg_off[...].
 *   entries_dev must stay valid until the next finalize.
 * pk_indexer_scan_result: number of windows counted by a scanner (its share of num_kmers). */
int pk_indexer_prime(pk_indexer *ix, const uint8_t *halo_dev, size_t n, uint64_t stream_off,
                     pk_stream stream);
int pk_indexer_scan_result(pk_indexer *ix, uint64_t *num_kmers);
int pk_indexer_export_segments(pk_indexer *ix, const uint32_t **entries_dev, uint32_t *nseg,
                               uint32_t *nwindows, uint32_t *seg_off_host, uint32_t *seg_cnt_host,
                               size_t capacity);
int pk_indexer_import_segments(pk_indexer *ix, const uint32_t *entries_dev, uint32_t nseg,
                               const uint32_t *seg_off_host, const uint32_t *seg_cnt_host);

/* Fused exchange (no separate all-to-all): pass 2 of the scanner stores every entry straight
 * into the k-mer buffer of the handle that owns its window, through CUDA-IPC peer mappings
 * over NVLink.  Per step: pk_indexer_scan_pass1 (count per window) -> pk_indexer_pass1_counts
 * -> the ranks all-gather the counts and derive the routing -> pk_indexer_scan_pass2_remote
 * (owner rank and destination offset per window; synchronises) -> barrier ->
 * pk_indexer_import_segments(owner, NULL, ...) on every owner (NULL = its own buffer).
 * pk_indexer_pool_ipc_handle exports a PARTITION handle's buffer (64-byte cudaIpcMemHandle_t),
 * pk_indexer_open_peer_pool maps it into a scanner as destination `peer` (pass the local
 * owner handle instead of an IPC handle for the rank's own windows). */
int pk_indexer_pool_ipc_handle(pk_indexer *ix, void *handle64, size_t *capacity_entries);
int pk_indexer_open_peer_pool(pk_indexer *scanner, int peer, const void *handle64,
                              pk_indexer *local_owner);
int pk_indexer_scan_pass1(pk_indexer *ix, const uint8_t *seq_dev, size_t n, pk_stream stream);
int pk_indexer_pass1_counts(pk_indexer *ix, uint32_t *counts_host, size_t nwindows);
int pk_indexer_scan_pass2_remote(pk_indexer *ix, int nranks, const uint32_t *owner_host,
                                 const uint32_t *dest_off_host, pk_stream stream);

/* Routed scan: the same fused exchange with NO host round trip inside a step -- for jobs that index
 * the same kind of stream again and again (bench.py's steady state; a planning scan sizes it once).
 * Every (source rank, window) owns a fixed region of the window owner's k-mer buffer:
 * pk_indexer_set_route gives a scanner, per window, the owner rank, the region's offset in the owner's
 * buffer and its room in entries (plus, per rank, where the owner's published-count table starts:
 * pk_indexer_pub_base, the last 2^18 entries of its buffer).  pk_indexer_scan_routed is then ONE pass:
 * scan, store into the regions (capacity-checked), count num_kmers, flag records, and write the fill
 * counts into every owner's table -- all asynchronous on `stream`; status_dev (four 32-bit words of the
 * caller's device memory) receives [0] = 1 if a region overflowed, in which case the step must be redone
 * with the exact two-pass protocol above, [1], [2] = low / high word of the scanner's num_kmers so far.  After ONE stream-ordered collective over all ranks (e.g. an NCCL
 * all-reduce of the status words: every rank's stores have landed when it completes)
 * pk_indexer_import_published turns the table into the owner's segment tables (layout given once by
 * pk_indexer_set_import_layout: seg_off_host[source][local window], first_window = global index of the
 * owner's window 0, nwindows_total = windows of the whole job) and pk_indexer_finalize counts. */
int pk_indexer_pub_base(pk_indexer *ix, uint64_t *entry_index);
int pk_indexer_set_route(pk_indexer *scanner, int nranks, int self_rank, const uint32_t *owner_host,
                         const uint32_t *dest_off_host, const uint32_t *cap_host,
                         const uint64_t *pub_base_host);
int pk_indexer_scan_routed(pk_indexer *scanner, const uint8_t *seq_dev, size_t n, uint32_t *status_dev,
                           pk_stream stream);
int pk_indexer_set_import_layout(pk_indexer *owner, uint32_t nseg, const uint32_t *seg_off_host,
                                 uint32_t first_window, uint32_t nwindows_total);
int pk_indexer_import_published(pk_indexer *owner, pk_stream stream);

/* Per-kernel-class device time, measured with CUDA events on the launching stream
 * around every launch made through the handle while enabled.  Classes (index into
 * ms_host / launches_host): 0 scan_count_direct, 1 scan_bucket_count, 2 bucket_offsets,
 * 3 scan_scatter, 4 window_count, 5 window_commit, 6 table_stats, 7 update_carry.
 * pk_indexer_profile synchronises the device, returns the sums and clears them. */
int pk_indexer_set_profiling(pk_indexer *ix, int enable);
int pk_indexer_profile(pk_indexer *ix, double ms_host[8], uint32_t launches_host[8]);
/* the counting scheme PK_MODE_AUTO resolved to, and its number of table windows */
int pk_indexer_mode(pk_indexer *ix, int *mode, int *windows);
/* log2 of the table entries per window (0 in DIRECT mode); every handle of one K agrees on it */
int pk_indexer_window_log2(pk_indexer *ix, int *window_log2);

/* Header.update_stats (tools.py:246-263) over any device table.
 * stats_host = {vals_sum, vals_count, vals_min, vals_max}.  Synchronises `stream`. */
int pk_table_stats_device(const uint8_t *table_dev, size_t n, int64_t hist_host[255],
                          uint64_t stats_host[4], pk_stream stream);

/* ------------------------------------------------------------------ merger
 * Replaces Header.calculate_distance (tools.py:439-493) and the pair loop of
 * merge (merger.py:136-176): valid = (min <= count <= max) (tools.py:473-474);
 * matrix[k][l] = (Total_k, Total_l, Shared_kl) = (G[k][k], G[l][l], G[k][l]) with
 * G = B * B^T over the 0/1 presence rows B.
 */
/* bit (i & 31) of bits_dev[i >> 5] = (min <= table_dev[i] <= max); n need not be
 * a multiple of 32 (tail bits are 0); table_dev 16-byte aligned. */
int pk_threshold_pack_device(const uint8_t *table_dev, size_t n, int min_count, int max_count,
                             uint32_t *bits_dev, pk_stream stream);
/* gram_dev[k * nsamples + l] (+)= popcount(bits[k] & bits[l]) over `words` 32-bit
 * words; sample k starts at bits_dev + k * stride_words.  accumulate = 0
 * overwrites gram_dev, 1 adds to it (k-mer-axis slabs / shards). */
int pk_gram_device(const uint32_t *bits_dev, int nsamples, size_t words, size_t stride_words,
                   int64_t *gram_dev, int accumulate, pk_stream stream);
/* Tiled masks -- the layout the tensor-core Gram kernel streams (any number of samples up to
 * PK_TILED_MAX_SAMPLES; more than 256 run block pair by block pair).  With row-major masks every
 * sample is its own stream and a CTA gathers 128 bytes from each of them per step: DRAM sees
 * ~38,000 interleaved streams and delivers ~1.1 TB/s.  Here the
 * words of ALL samples for the same 1024 k-mers lie side by side,
 *     word g of sample r  ->  bits_tiled_dev[(g / 32) * nrows * 32 + r * 32 + g % 32],
 * so the gather is one sequential stream.  The buffer holds ceil(words / 32) * nrows * 32 words and
 * must be zeroed before packing (the padding of the last tile is read).
 * pk_threshold_pack_tiled_device packs n entries of sample `row` whose first entry is k-mer
 * 32 * first_word of that sample (slabs of a table may be packed by separate calls).
 * pk_gram_tiled_device is pk_gram_device on such a buffer (tcgen05 kind::mxf4 with FP32 accumulators
 * that stay exact integers, see gram_f4.cu); `words` = words per sample.
 * pk_gram_tiled_exact: *exact = 1 if the device accumulates 0/1 FP4 products exactly up to 2^24 (checked
 * once per device by driving one accumulator through every integer from 2^23 up; cached), 0 if not --
 * pk_gram_tiled_device then refuses (PK_ERR_STATE) and the caller packs row-major masks for
 * pk_gram_device's integer kernels instead (pk_merge_host does so by itself). */
#define PK_TILED_MAX_SAMPLES 4096
int pk_gram_tiled_exact(int device, int *exact);
int pk_threshold_pack_tiled_device(const uint8_t *table_dev, size_t n, size_t first_word, int min_count,
                                   int max_count, uint32_t *bits_tiled_dev, int row, int nrows,
                                   pk_stream stream);
int pk_gram_tiled_device(const uint32_t *bits_tiled_dev, int nsamples, size_t words, int64_t *gram_dev,
                         int accumulate, pk_stream stream);
/* One pair, straight from two device tables: out_host = {Total_s, Total_o, Shared}
 * -- the return value of Header.calculate_distance (tools.py:493).  Synchronises. */
int pk_pair_counts_device(const uint8_t *s_dev, const uint8_t *o_dev, size_t n, int min_count,
                          int max_count, uint64_t out_host[3], pk_stream stream);
/* Whole merge from host tables: streams every table once through pinned-copy +
 * threshold_pack, then one Gram pass.  matrix_host is uint64[N][N][3]
 * (merger.py:136; the diagonal, which the reference leaves uninitialised, is
 * (T_k, T_k, T_k)). */
int pk_merge_host(const uint8_t *const *tables_host, int nsamples, size_t n, int min_count,
                  int max_count, int device, uint64_t *matrix_host);

/* ------------------------------------------------------------------ host ingest
 * Host-only helpers of the FASTA reader (pykmer_b200/fasta.py); no CUDA call, usable without
 * a GPU.  They take over the two passes over the text that dominate the CLI's wall time:
 * gunzipping a .bgz input (read_fasta, indexer.py:108-115 -- BGZF members are independent, so
 * they inflate in parallel) and stripping + joining sequence lines (parse_fasta,
 * indexer.py:55-95) when the lines hold no inner white space.
 *
 * pk_bgzf_inflate: inflate as many WHOLE BGZF members of comp[0, comp_len) as fit into
 * out[0, out_cap) on `threads` threads (0 = all cores); CRC and length of every member are
 * checked.  *consumed = compressed bytes used (a prefix of whole members; a truncated last
 * member is left for the next call), *produced = bytes written.
 * pk_fasta_clean: dst = src minus '\n' and '\r'; *n_out = bytes kept.  flags bit 0: src holds
 * other strip()-able white space (blank, tab, \v, \f, 0x1c-0x1f), bit 1: a byte >= 0x80 -- with
 * either set dst is not written and the caller goes line by line / rejects the input. */
int pk_bgzf_inflate(const uint8_t *comp, size_t comp_len, uint8_t *out, size_t out_cap,
                    size_t *consumed, size_t *produced, int threads);
int pk_fasta_clean(const uint8_t *src, size_t n, uint8_t *dst, size_t *n_out, uint32_t *flags,
                   int threads);
/* pk_fasta_find_headers: offsets of every '>' that starts a line of text[0, n) (offset 0 or right after
 * '\n' / '\r') -- where parse_fasta opens a record (indexer.py:62-80); ascending, on `threads` threads.
 * *count = number found, also when it exceeds cap (only the first cap are stored). */
int pk_fasta_find_headers(const uint8_t *text, size_t n, uint64_t *pos_out, size_t cap, size_t *count,
                          int threads);
/* pk_bgzf_deflate: the output side -- what the documented workflow does with the external
 * `bgzip -l 9` after every indexer run (README.md:26,261-269; data/README.md:24) and what the
 * merger reads back (tools.py:296-302).  src[0, n) becomes ceil(n / 0xFF00) independent BGZF
 * members, deflated on `threads` threads (0 = all cores) and written back to back into
 * out[0, *produced); out_cap >= ceil(n / 0xFF00) * 65536.  No EOF member is appended (the
 * caller writes it once per file).  member_sizes (may be NULL) receives each member's
 * compressed size: the material of a .gzi index (gzireader.py:12-19). */
int pk_bgzf_deflate(const uint8_t *src, size_t n, uint8_t *out, size_t out_cap, size_t *produced,
                    uint32_t *member_sizes, int level, int threads);

/* Deterministic synthetic count table (benchmark input):
 * entries [lo, hi) of sample `sample`, bit-identical to pykmer_b200/synth.py. */
int pk_synth_table_device(uint8_t *dst_dev, int sample, uint64_t lo, uint64_t hi,
                          pk_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* PYKMER_B200_H */

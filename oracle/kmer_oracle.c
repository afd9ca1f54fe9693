/*
 * kmer_oracle.c -- CPU restatement of the two pykmer hot paths.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: it may be called
 * from tests/, from __graft_entry__.smoke() and from bench.py's cpu_baseline /
 * `--impl reference` leg, and from nowhere else.  The product path
 * (pykmer_b200/) never links, imports or falls back to anything in oracle/.
 *
 * Parity status: PINNED.  oracle/make_golden.py runs the reference's own
 * indexer.py / merger.py (imported from /root/reference with three
 * non-arithmetic shims) and tests/test_oracle_golden.py checks every function
 * below against the committed outputs under tests/golden/.
 *
 * All citations are file:line into the reference (sauloal/pykmer).
 *
 * Sequence "stream" convention used by every indexer entry point: the host has
 * already applied the reference's text rules (indexer.py:55-95: strip lines,
 * drop headers, join the lines of a record) and concatenated the records with
 * ONE byte that is not in ACGTacgt between them.  Because any non-ACGT byte
 * voids the K windows that contain it (indexer.py:144), that separator also
 * enforces "windows never cross records" (indexer.py:133-141).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define OK_EXPORT __attribute__((visibility("default")))

/* minimal pthread parallel-for (dynamic, one item at a time) */
typedef void (*ok_item_fn)(long item, void *ctx);
typedef struct { ok_item_fn fn; void *ctx; long n; long next; } ok_pf;
static void *ok_pf_worker(void *arg) {
    ok_pf *pf = (ok_pf *)arg;
    for (;;) {
        long i = __atomic_fetch_add(&pf->next, 1, __ATOMIC_RELAXED);
        if (i >= pf->n) break;
        pf->fn(i, pf->ctx);
    }
    return NULL;
}
static void ok_parallel_for(long n, int threads, ok_item_fn fn, void *ctx) {
    ok_pf pf = { fn, ctx, n, 0 };
    if (threads <= 1 || n <= 1) { ok_pf_worker(&pf); return; }
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    int started = 0;
    for (int t = 0; t < threads - 1; t++)
        if (pthread_create(&tid[started], NULL, ok_pf_worker, &pf) == 0) started++;
    ok_pf_worker(&pf);
    for (int t = 0; t < started; t++) pthread_join(tid[t], NULL);
}

/* indexer.py:36-41 -- CONV: A/a->0 C/c->1 G/g->2 T/t->3, everything else None */
static inline int ok_code(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

/* indexer.py:239,262 -- counts are clamped to 255 and added with
 * a + min(255 - a, v); for v == 1 that is a saturating increment. */
static inline void ok_sat_inc(uint8_t *slot) {
    if (*slot != 255) (*slot)++;
}

/* upper_bound(rec_starts, pos) - 1: which record a stream position is in */
static size_t ok_record_of(const uint64_t *rec_starts, size_t nrec, uint64_t pos) {
    size_t lo = 0, hi = nrec;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (rec_starts[mid] <= pos) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

/*
 * Direct restatement of gen_kmers (indexer.py:130-160) + the consumer loop of
 * create_fasta_index (indexer.py:340-342,377): for every window start, skip it
 * if it holds a None, otherwise fwd = sum pos_val[p]*j, rev = sum
 * pos_val[K-1-p]*(3-j), pos = min(fwd, rev), num_kmers += 1.  O(K) per base,
 * kept deliberately literal; ok_index_rolling below is the fast form and is
 * tested equal to this one.
 *
 * table covers canonical indices [range_lo, range_hi); windows whose canonical
 * index falls outside are not counted (neither in table nor in num_kmers) --
 * that is the k-mer-range sharding of SURVEY section 8(e); the full table is
 * range_lo = 0, range_hi = 4^K.
 * rec_starts/rec_flags (optional): rec_flags[r] is set to 1 when record r
 * yields at least one counted window (indexer.py:349-351, "chromosomes").
 */
OK_EXPORT int ok_index_direct(const uint8_t *seq, size_t n, int K,
                              uint64_t range_lo, uint64_t range_hi,
                              uint8_t *table, uint64_t *num_kmers,
                              const uint64_t *rec_starts, size_t nrec,
                              uint8_t *rec_flags) {
    if (K <= 0 || K > 31) return -1;
    uint64_t pos_val[32];
    for (int p = 0; p < K; p++) pos_val[p] = 1ull << (2 * (K - p - 1)); /* indexer.py:131 */
    uint64_t count = 0;
    if (n >= (size_t)K) {
        for (size_t i = 0; i + K <= n; i++) {                          /* indexer.py:141 */
            int bad = 0;
            uint64_t fwd = 0, rev = 0;
            for (int p = 0; p < K; p++) {
                int j = ok_code(seq[i + p]);
                if (j < 0) { bad = 1; break; }                          /* indexer.py:144 */
                fwd += pos_val[p] * (uint64_t)j;                        /* indexer.py:149 */
                rev += pos_val[K - p - 1] * (uint64_t)(3 - j);          /* indexer.py:150 */
            }
            if (bad) continue;
            uint64_t pos = fwd < rev ? fwd : rev;                       /* indexer.py:341 */
            if (pos < range_lo || pos >= range_hi) continue;
            count++;                                                    /* indexer.py:342 */
            ok_sat_inc(&table[pos - range_lo]);
            if (rec_flags) rec_flags[ok_record_of(rec_starts, nrec, i)] = 1;
        }
    }
    *num_kmers += count;
    return 0;
}

/*
 * Rolling form of the same computation: shift the new base into the low end
 * of fwd, the complement into the high end of rev, and keep a run length of
 * consecutive valid bases.  Same outputs as ok_index_direct.
 */
static void ok_roll_span(const uint8_t *seq, size_t from, size_t to, size_t n, int K,
                         uint64_t range_lo, uint64_t range_hi,
                         uint8_t *table, uint64_t *num_kmers, int atomic,
                         const uint64_t *rec_starts, size_t nrec, uint8_t *rec_flags) {
    /* counts windows whose END position e lies in [from, to) */
    const uint64_t mask = (K == 32) ? ~0ull : ((1ull << (2 * K)) - 1);
    const int top = 2 * (K - 1);
    uint64_t fwd = 0, rev = 0, count = 0;
    size_t run = 0;
    size_t begin = from >= (size_t)(K - 1) ? from - (K - 1) : 0;
    (void)n;
    for (size_t e = begin; e < to; e++) {
        int j = ok_code(seq[e]);
        if (j < 0) { run = 0; fwd = 0; rev = 0; continue; }
        fwd = ((fwd << 2) | (uint64_t)j) & mask;
        rev = (rev >> 2) | ((uint64_t)(3 - j) << top);
        run++;
        if (run < (size_t)K || e < from) continue;
        uint64_t pos = fwd < rev ? fwd : rev;
        if (pos < range_lo || pos >= range_hi) continue;
        count++;
        uint8_t *slot = &table[pos - range_lo];
        if (!atomic) {
            ok_sat_inc(slot);
        } else {
            uint8_t cur = __atomic_load_n(slot, __ATOMIC_RELAXED);
            while (cur != 255 &&
                   !__atomic_compare_exchange_n(slot, &cur, (uint8_t)(cur + 1), 1,
                                                __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { }
        }
        if (rec_flags) rec_flags[ok_record_of(rec_starts, nrec, e)] = 1;
    }
    if (atomic) __atomic_fetch_add(num_kmers, count, __ATOMIC_RELAXED);
    else *num_kmers += count;
}

OK_EXPORT int ok_index_rolling(const uint8_t *seq, size_t n, int K,
                               uint64_t range_lo, uint64_t range_hi,
                               uint8_t *table, uint64_t *num_kmers,
                               const uint64_t *rec_starts, size_t nrec,
                               uint8_t *rec_flags) {
    if (K <= 0 || K > 31) return -1;
    ok_roll_span(seq, 0, n, n, K, range_lo, range_hi, table, num_kmers, 0,
                 rec_starts, nrec, rec_flags);
    return 0;
}

/*
 * Threaded rolling form for the CPU baseline leg of bench.py ("all the host
 * threads it can use"): the stream is cut into spans, each thread rolls its
 * span (re-reading a K-1 halo) and increments the shared table with a
 * saturating compare-exchange.  Same outputs as ok_index_rolling.
 */
typedef struct {
    const uint8_t *seq; size_t n; int K; uint64_t lo, hi;
    uint8_t *table; uint64_t *num_kmers; size_t spans, span;
} ok_span_ctx;
static void ok_span_item(long s, void *ctx) {
    ok_span_ctx *c = (ok_span_ctx *)ctx;
    size_t from = (size_t)s * c->span, to = from + c->span;
    if (from >= c->n) return;
    if (to > c->n) to = c->n;
    ok_roll_span(c->seq, from, to, c->n, c->K, c->lo, c->hi, c->table, c->num_kmers, 1,
                 NULL, 0, NULL);
}

OK_EXPORT int ok_index_rolling_mt(const uint8_t *seq, size_t n, int K,
                                  uint64_t range_lo, uint64_t range_hi,
                                  uint8_t *table, uint64_t *num_kmers, int threads) {
    if (K <= 0 || K > 31) return -1;
    if (threads < 1) threads = 1;
    ok_span_ctx c;
    c.spans = (size_t)threads * 16;
    c.span = (n + c.spans - 1) / c.spans;
    if (c.span == 0) c.span = 1;
    c.seq = seq; c.n = n; c.K = K; c.lo = range_lo; c.hi = range_hi;
    c.table = table; c.num_kmers = num_kmers;
    ok_parallel_for((long)c.spans, threads, ok_span_item, &c);
    return 0;
}

/*
 * Header.update_stats (tools.py:246-263): hist = histogram(arr, bins=255,
 * range=(1,255)) i.e. hist[i] = #{arr == i+1}; vals_sum, vals_count (non-zero
 * entries), vals_min, vals_max.  stats = {vals_sum, vals_count, vals_min,
 * vals_max}.
 */
OK_EXPORT int ok_table_stats(const uint8_t *table, size_t n, int64_t hist[255],
                             uint64_t stats[4]) {
    uint64_t bins[256];
    memset(bins, 0, sizeof bins);
    for (size_t i = 0; i < n; i++) bins[table[i]]++;
    uint64_t sum = 0, cnt = 0;
    int mn = 255, mx = 0;
    for (int v = 0; v < 256; v++) {
        if (v > 0) hist[v - 1] = (int64_t)bins[v];
        if (bins[v]) {
            sum += bins[v] * (uint64_t)v;
            if (v > 0) cnt += bins[v];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
    }
    if (n == 0) { mn = 0; mx = 0; }
    stats[0] = sum; stats[1] = cnt; stats[2] = (uint64_t)mn; stats[3] = (uint64_t)mx;
    return 0;
}

/* threaded form of ok_table_stats for the CPU baseline leg (private bins per slab) */
typedef struct { const uint8_t *table; size_t n, slab; uint64_t (*bins)[256]; } ok_stats_ctx;
static void ok_stats_item(long s, void *ctx) {
    ok_stats_ctx *c = (ok_stats_ctx *)ctx;
    size_t from = (size_t)s * c->slab, to = from + c->slab;
    if (to > c->n) to = c->n;
    uint64_t *b = c->bins[s];
    for (size_t i = from; i < to; i++) b[c->table[i]]++;
}
OK_EXPORT int ok_table_stats_mt(const uint8_t *table, size_t n, int64_t hist[255],
                                uint64_t stats[4], int threads) {
    if (threads < 1) threads = 1;
    long slabs = (long)threads * 4;
    ok_stats_ctx c;
    c.table = table; c.n = n; c.slab = (n + (size_t)slabs - 1) / (size_t)slabs;
    if (c.slab == 0) c.slab = 1;
    c.bins = (uint64_t (*)[256])calloc((size_t)slabs, sizeof(uint64_t[256]));
    if (!c.bins) return -1;
    ok_parallel_for(slabs, threads, ok_stats_item, &c);
    uint64_t bins[256];
    memset(bins, 0, sizeof bins);
    for (long s = 0; s < slabs; s++)
        for (int v = 0; v < 256; v++) bins[v] += c.bins[s][v];
    free(c.bins);
    uint64_t sum = 0, cnt = 0;
    int mn = 255, mx = 0;
    for (int v = 0; v < 256; v++) {
        if (v > 0) hist[v - 1] = (int64_t)bins[v];
        if (bins[v]) {
            sum += bins[v] * (uint64_t)v;
            if (v > 0) cnt += bins[v];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
    }
    if (n == 0) { mn = 0; mx = 0; }
    stats[0] = sum; stats[1] = cnt; stats[2] = (uint64_t)mn; stats[3] = (uint64_t)mx;
    return 0;
}

/*
 * Header.calculate_distance (tools.py:439-493), the arithmetic of one pair:
 * s_valid = (s >= min) & (s <= max), o_valid likewise, c_valid = both
 * (tools.py:473-475); returns the three sums (tools.py:480-482).
 */
OK_EXPORT int ok_pair_counts(const uint8_t *s, const uint8_t *o, size_t n,
                             int min_count, int max_count, uint64_t out[3]) {
    uint64_t sc = 0, oc = 0, cc = 0;
    for (size_t i = 0; i < n; i++) {
        int sv = s[i] >= min_count && s[i] <= max_count;
        int ov = o[i] >= min_count && o[i] <= max_count;
        sc += (uint64_t)sv; oc += (uint64_t)ov; cc += (uint64_t)(sv & ov);
    }
    out[0] = sc; out[1] = oc; out[2] = cc;
    return 0;
}

/*
 * merger.py:136-176 -- the (N, N, 3) matrix: for every k < l the worker's
 * triple (Total_k, Total_l, Shared_kl) is stored at [k,l] and mirrored at
 * [l,k] as (Total_l, Total_k, Shared_kl).  The reference leaves the diagonal
 * uninitialised (merger.py:136); here it is defined as (T_k, T_k, T_k).
 * tables: N rows of n bytes, row stride `stride` bytes.  One task per pair,
 * as the reference's Pool does (merger.py:139-153), spread over `threads`.
 */
typedef struct {
    const uint8_t *tables; int N; size_t n, stride; int min_count, max_count;
    uint64_t *matrix; const int *pk, *pl;
} ok_pair_ctx;
static void ok_pair_item(long p, void *ctx) {
    ok_pair_ctx *c = (ok_pair_ctx *)ctx;
    int k = c->pk[p], l = c->pl[p];
    uint64_t out[3];
    ok_pair_counts(c->tables + (size_t)k * c->stride, c->tables + (size_t)l * c->stride,
                   c->n, c->min_count, c->max_count, out);
    uint64_t *a = c->matrix + ((size_t)k * c->N + l) * 3;
    uint64_t *b = c->matrix + ((size_t)l * c->N + k) * 3;
    a[0] = out[0]; a[1] = out[1]; a[2] = out[2];      /* merger.py:175 */
    b[0] = out[1]; b[1] = out[0]; b[2] = out[2];      /* merger.py:176 */
}

OK_EXPORT int ok_merge_matrix(const uint8_t *tables, int N, size_t n, size_t stride,
                              int min_count, int max_count, uint64_t *matrix, int threads) {
    if (threads < 1) threads = 1;
    long npairs = (long)N * (N - 1) / 2;
    int *pk = (int *)malloc(sizeof(int) * (size_t)(npairs > 0 ? npairs : 1));
    int *pl = (int *)malloc(sizeof(int) * (size_t)(npairs > 0 ? npairs : 1));
    long t = 0;
    for (int k = 0; k < N - 1; k++)
        for (int l = k + 1; l < N; l++) { pk[t] = k; pl[t] = l; t++; }
    for (int k = 0; k < N; k++) {
        uint64_t out[3];
        ok_pair_counts(tables + (size_t)k * stride, tables + (size_t)k * stride, n,
                       min_count, max_count, out);
        uint64_t *c = matrix + ((size_t)k * N + k) * 3;
        c[0] = out[0]; c[1] = out[0]; c[2] = out[0];
    }
    ok_pair_ctx c = { tables, N, n, stride, min_count, max_count, matrix, pk, pl };
    ok_parallel_for(npairs, threads, ok_pair_item, &c);
    free(pk); free(pl);
    return 0;
}

/* threshold + pack (the bitmask form of tools.py:473-474): bit (i & 31) of
 * word i >> 5 is 1 iff min <= table[i] <= max.  n need not be a multiple of 32. */
OK_EXPORT int ok_threshold_pack(const uint8_t *table, size_t n, int min_count,
                                int max_count, uint32_t *bits) {
    size_t words = (n + 31) / 32;
    memset(bits, 0, words * sizeof(uint32_t));
    for (size_t i = 0; i < n; i++)
        if (table[i] >= min_count && table[i] <= max_count)
            bits[i >> 5] |= 1u << (i & 31);
    return 0;
}

OK_EXPORT int ok_max_threads(void) {
    long v = sysconf(_SC_NPROCESSORS_ONLN);
    return v > 0 ? (int)v : 1;
}

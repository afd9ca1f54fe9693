// Host-side check of pykmer_b200/csrc/kmer_bits.h against a literal restatement of
// indexer.py:141-150 (no GPU needed).  Exit code 0 = all windows agree.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "kmer_bits.h"

static int code_of(uint8_t c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1;
                 case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return -1; }
}

int main() {
    uint8_t lut[256];
    for (int c = 0; c < 256; c++) lut[c] = (uint8_t)pk_lut_entry((uint32_t)c);
    const char alphabet[] = "ACGTacgtACGTACGTACGTNn>*\n";
    uint64_t rng = 88172645463325252ull;
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    const size_t n = 16 * 4096;
    std::vector<uint8_t> seq(n);
    for (size_t i = 0; i < n; i++) seq[i] = (uint8_t)alphabet[next() % (sizeof(alphabet) - 1)];
    for (size_t i = 5000; i < 9000; i++) seq[i] = "ACGT"[next() & 3];     // long valid stretch
    for (size_t i = 20000; i < 20400; i++) seq[i] = 'A';
    size_t groups = n / 16;
    std::vector<uint32_t> codes(groups), vm(groups);
    for (size_t g = 0; g < groups; g++) {
        uint32_t w[4];
        memcpy(w, &seq[16 * g], 16);
        pk_encode16(w, lut, codes[g], vm[g]);
    }
    long checked = 0, valid = 0;
    for (int K = 1; K <= 31; K += 2) {
        for (size_t g = 2; g < groups; g++) {
            uint32_t pc2 = codes[g - 2], pc1 = codes[g - 1], cc = codes[g];
            uint64_t vcat = ((uint64_t)vm[g - 2] << 32) | ((uint64_t)vm[g - 1] << 16) | vm[g];
            uint64_t W = pk_valid_windows(vcat, K);
            uint32_t r0 = pk_rcw(cc), r1 = pk_rcw(pc1), r2 = pk_rcw(pc2);
            for (int j = 0; j < 16; j++) {
                size_t e = 16 * g + j;            // window end
                bool ok = true;
                uint64_t fwd = 0, rev = 0;
                for (int p = 0; p < K; p++) {
                    int c = code_of(seq[e - K + 1 + p]);
                    if (c < 0) { ok = false; break; }
                    fwd += ((uint64_t)c) << (2 * (K - 1 - p));
                    rev += ((uint64_t)(3 - c)) << (2 * p);
                }
                bool got_ok = (W >> (15 - j)) & 1;
                checked++;
                if (got_ok != ok) { printf("validity mismatch K=%d e=%zu\n", K, e); return 1; }
                if (!ok) continue;
                valid++;
                uint64_t f = pk_fwd_at(pc2, pc1, cc, j, K), r = pk_rc_at(r2, r1, r0, j, K);
                if (f != fwd || r != rev) {
                    printf("value mismatch K=%d e=%zu fwd %llx/%llx rc %llx/%llx\n", K, e,
                           (unsigned long long)f, (unsigned long long)fwd,
                           (unsigned long long)r, (unsigned long long)rev);
                    return 1;
                }
                if (K <= 16) {
                    uint32_t m = (uint32_t)pk_kmer_mask(K);
                    uint64_t cat = ((uint64_t)pc1 << 32) | cc, rcat = ((uint64_t)r0 << 32) | r1;
                    uint64_t W32 = pk_valid_windows(((uint64_t)vm[g - 1] << 16) | vm[g], K);
                    if (!((W32 >> (15 - j)) & 1)) { printf("W32 mismatch K=%d\n", K); return 1; }
                    if (pk_fwd32_at(cat, j, m) != (uint32_t)fwd || pk_rc32_at(rcat, j, K, m) != (uint32_t)rev) {
                        printf("32-bit form mismatch K=%d e=%zu\n", K, e);
                        return 1;
                    }
                }
            }
        }
    }
    // pk_scan_group: runs expanded back must equal the per-window brute force, per group
    long runs = 0, slots_ok = 0;
    for (int K = 1; K <= 31; K += 2) {
        const uint64_t T = 1ull << (2 * K);
        const uint64_t los[2] = {0, T / 4}, his[2] = {T, T / 2 + 7};
        for (int rg = 0; rg < 2; rg++) {
            const uint64_t lo = los[rg], hi = his[rg];
            for (size_t g = 2; g < groups; g++) {
                std::vector<uint64_t> want, got;
                std::vector<int> fresh_want;
                bool fresh = true;
                for (int j = 0; j < 16; j++) {
                    size_t e = 16 * g + j;
                    if (code_of(seq[e]) < 0) fresh = true;
                    bool ok = true;
                    uint64_t fwd = 0, rev = 0;
                    for (int p = 0; p < K; p++) {
                        int c = code_of(seq[e - K + 1 + p]);
                        if (c < 0) { ok = false; break; }
                        fwd += ((uint64_t)c) << (2 * (K - 1 - p));
                        rev += ((uint64_t)(3 - c)) << (2 * p);
                    }
                    if (!ok) continue;
                    uint64_t canon = fwd < rev ? fwd : rev;
                    if (canon < lo || canon >= hi) continue;
                    want.push_back(canon - lo);
                    fresh_want.push_back(fresh ? 1 : 0);
                    fresh = false;
                }
                int last_slot = -1;
                bool slot_order = true;
                auto emit = [&](int slot, uint64_t off, uint32_t cnt) {
                    if (slot <= last_slot || slot > 16 || cnt < 1 || cnt > 16) slot_order = false;
                    last_slot = slot;
                    runs++;
                    for (uint32_t c = 0; c < cnt; c++) got.push_back(off);
                };
                const bool full = (rg == 0);
                uint32_t cm;
                if (K > 16) {
                    cm = full ? pk_scan_group<true, true>(K, lo, hi - lo, codes[g], vm[g], codes[g - 1], vm[g - 1], codes[g - 2], vm[g - 2], emit)
                              : pk_scan_group<true, false>(K, lo, hi - lo, codes[g], vm[g], codes[g - 1], vm[g - 1], codes[g - 2], vm[g - 2], emit);
                } else {
                    cm = full ? pk_scan_group<false, true>(K, lo, hi - lo, codes[g], vm[g], codes[g - 1], vm[g - 1], 0u, 0u, emit)
                              : pk_scan_group<false, false>(K, lo, hi - lo, codes[g], vm[g], codes[g - 1], vm[g - 1], 0u, 0u, emit);
                }
                // counted mask -> fresh positions, as the record flagger will visit them
                std::vector<int> visit_want, visit_got;
                {
                    int wi = 0;
                    for (int j = 0; j < 16; j++)
                        if ((cm >> (15 - j)) & 1u) { if (fresh_want[wi]) visit_want.push_back(j); wi++; }
                }
                pk_for_each_record_run(vm[g], cm, [&](int j) { visit_got.push_back(j); });
                uint32_t n = 0;
                for (int b = 0; b < 16; b++) n += (cm >> b) & 1u;
                if (!slot_order || n != want.size() || got != want || visit_got != visit_want) {
                    printf("scan_group mismatch K=%d g=%zu range %d: n=%u want=%zu\n", K, g, rg, n, want.size());
                    return 1;
                }
                slots_ok++;
            }
        }
    }
    printf("ok: %ld windows checked, %ld valid; %ld groups scanned, %ld runs\n", checked, valid, slots_ok, runs);
    return 0;
}

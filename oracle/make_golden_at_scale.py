#!/usr/bin/env python3
"""Digests of the oracle's outputs at BASELINE.json's FULL sizes -> tests/golden/at_scale.json.

TEST INFRASTRUCTURE ONLY.  The oracle (oracle/kmer_oracle.c, pinned against the reference's own
indexer.py / merger.py by tests/golden/, see make_golden.py) is run here, on the CPU, on the whole
synthetic 782,520,033 bp stream of configs 2 / 4 / 5 (pykmer_b200/synth.py, seeded), and what it
produced is committed as sha256 digests + statistics.  tests/test_gpu_at_scale.py then holds the
CUDA path to them at full size on the GPU box (where it also re-runs the oracle live), so a later
round cannot drift from what this round verified.

    python oracle/make_golden_at_scale.py          # ~2 min, ~20 GB of RAM
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from pykmer_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "at_scale.json")

# the table slices checked at K=17 (config 4; four 4 GiB quarters = the whole 16 GiB table) and at
# K=19 (config 5; the table has 256 GiB -- two disjoint 2^28-entry ranges, one in the crowded low
# end of the canonical axis, one in the middle)
K17_RANGES = [(q << 32, (q + 1) << 32) for q in range(4)]
K19_RANGES = [(0, 1 << 28), ((1 << 37) + (5 << 28), (1 << 37) + (6 << 28))]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(memoryview(np.ascontiguousarray(a))).hexdigest()


def digest_range(stream: np.ndarray, K: int, lo: int, hi: int, threads: int) -> dict:
    table, num, _ = oracle.index_stream(stream, K, range_lo=lo, range_hi=hi, method="mt", threads=threads)
    hist, st = oracle.table_stats(table, threads=threads)
    return {"range": [lo, hi], "sha256": sha(table), "num_kmers": num, "hist": hist,
            "vals_sum": st["vals_sum"], "vals_count": st["vals_count"], "vals_min": st["vals_min"],
            "vals_max": st["vals_max"]}


def main() -> None:
    threads = oracle.max_threads()
    recs = synth.syn782m_records(scale=1.0)
    stream, starts, lengths, names = synth.records_to_stream(recs)
    doc = {"stream": {"bytes": int(stream.size), "bp": int(sum(lengths)), "records": len(lengths),
                      "sha256": sha(stream), "seed": hex(synth.SYN782M_SEED)},
           "generator": "oracle/make_golden_at_scale.py (oracle/kmer_oracle.c ok_index_rolling_mt)"}
    print("K=15 full table", flush=True)
    doc["k15"] = digest_range(stream, 15, 0, 4 ** 15, threads)
    # record flags (indexer.py:349-351) need the single-threaded form of the oracle
    _, num, flags = oracle.index_stream(stream, 15, rec_starts=starts)
    assert num == doc["k15"]["num_kmers"]
    doc["k15"]["record_flags"] = [int(f) for f in flags]
    doc["k17"] = []
    for lo, hi in K17_RANGES:
        print(f"K=17 [{lo}, {hi})", flush=True)
        doc["k17"].append(digest_range(stream, 17, lo, hi, threads))
    doc["k19"] = []
    for lo, hi in K19_RANGES:
        print(f"K=19 [{lo}, {hi})", flush=True)
        doc["k19"].append(digest_range(stream, 19, lo, hi, threads))
    with open(OUT, "w") as fh:
        json.dump(doc, fh, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

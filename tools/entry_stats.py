#!/usr/bin/env python3
"""CPU-side look at the synthetic stream: how skewed are the per-window k-mer counts?
(analysis aid, not product code)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import oracle
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.125
stream, starts, lengths = bench.load_stream(scale, 0, 1)
K = 15
tot = 0
rows = []
for b in range(64):
    t, n, _ = oracle.index_stream(stream, K, range_lo=b << 24, range_hi=(b + 1) << 24)
    rows.append((b, n, int((t == 255).sum())))
    tot += n
print("total", tot)
for b, n, sat in rows:
    print(b, n, "%.2f%%" % (100.0 * n / tot), "saturated", sat)

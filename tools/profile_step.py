#!/usr/bin/env python3
"""One indexer step (reset -> feed_device -> finalize) on a scaled config-2 stream; the
short command line ncu is pointed at.
    python tools/profile_step.py [scale] [K] [mode] [steps] [lo:hi]      (lo:hi = k-mer range in GiB)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from pykmer_b200 import device as dev  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
K = int(sys.argv[2]) if len(sys.argv) > 2 else 15
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
lo, hi = (int(float(x) * 2 ** 30) for x in sys.argv[5].split(":")) if len(sys.argv) > 5 else (0, 4 ** K)
stream, starts, lengths = bench.load_stream(scale, 0, 1)
d = torch.from_numpy(stream).cuda()
with dev.Indexer(K, mode=mode, range_lo=lo, range_hi=hi) as ix:
    ix.set_records(starts)
    for _ in range(steps):
        ix.reset()
        ix.feed_device(d)
        hist, st = ix.finalize()
    torch.cuda.synchronize()
    print("mode", ix.mode(), "num_kmers", st["num_kmers"], "vals_count", st["vals_count"])

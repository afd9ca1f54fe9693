#!/usr/bin/env python3
"""Per-launch device time of one full-size indexer step (library profiling, verbose dump on stderr)."""
import os, sys
os.environ["PYKMER_B200_VERBOSE"] = "2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from pykmer_b200 import device as dev
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
stream, starts, lengths = bench.load_stream(scale, 0, 1)
d = torch.from_numpy(stream).cuda()
with dev.Indexer(15) as ix:
    ix.set_records(starts)
    for rep in range(2):
        ix.reset(); ix.feed_device(d); ix.finalize()
    ix.set_profiling(True)
    ix.reset(); ix.feed_device(d); ix.finalize()
    print(ix.profile())

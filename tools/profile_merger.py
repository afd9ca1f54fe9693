#!/usr/bin/env python3
"""One merger pass (pack + Gram) on synthetic samples; the short command line ncu is pointed at.
    python tools/profile_merger.py [N] [K]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from pykmer_b200 import device as dev  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
K = int(sys.argv[2]) if len(sys.argv) > 2 else 15
T = 4 ** K
words = T // 32
raw = torch.empty(T, dtype=torch.uint8, device="cuda")
if dev.use_tiled_masks(N):                      # the merger's default for <= 256 samples
    bits = dev.tiled_masks(words, N)
    for s in range(N):
        dev.synth_table(s, 0, T, out=raw)
        dev.threshold_pack_tiled(raw, 1, 50, bits, s, N)
    G = dev.gram_tiled(bits, N, words)
    G = dev.gram_tiled(bits, N, words)
else:
    bits = torch.zeros((N, words), dtype=torch.int32, device="cuda")
    for s in range(N):
        dev.synth_table(s, 0, T, out=raw)
        dev.threshold_pack(raw, 1, 50, out=bits[s])
    G = dev.gram(bits)
    G = dev.gram(bits)
torch.cuda.synchronize()
print("N", N, "K", K, "trace", int(G.diagonal().sum()), "G01", int(G[0, 1]))

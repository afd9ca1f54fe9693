#!/usr/bin/env python3
"""Development aid (1 GPU): what does k_scan_scatter's REMOTE form cost when every destination is local?
Separates the cost of the per-window owner / peer-pointer indirection from the cost of NVLink stores
(at 2 GPUs the scatter of half a stream takes 2.39 ms against 1.69 for the same bases stored locally)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from pykmer_b200 import device as dev, dist as pdist, _native as nat  # noqa: E402

K = 15
stream, starts, lengths = bench.load_stream(1.0, 0, 1)
half = (stream.size // 2) // 16 * 16
d = torch.from_numpy(np.ascontiguousarray(stream[:half])).cuda()

with dev.Indexer(K) as ix:                                   # the single-GPU path on the same half stream
    for _ in range(2):
        ix.reset(); ix.feed_device(d); ix.finalize()
    ix.set_profiling(True)
    ix.reset(); ix.feed_device(d); ix.finalize()
    print("local pass 2 (estimated pass 1)   :", {k: round(v[0], 3) for k, v in ix.profile().items()})

sc = dev.Indexer(K, mode=nat.PK_MODE_SCAN)
ct = dev.Indexer(K, mode=nat.PK_MODE_PARTITION)
sc.open_peer_pool(0, local_owner=ct)
nwin = sc.mode()[1]
for rep in range(3):
    sc.reset(); ct.reset(); sc.prime(None, 0)
    if rep == 2:
        sc.set_profiling(True)
    sc.scan_pass1(d)
    cnt = sc.pass1_counts().astype(np.int64)[None, :]
    owner_of, dest_off, imp_off, imp_cnt, landed = pdist.plan_fused(cnt, [(0, nwin)], 0)
    sc.scan_pass2_remote(1, owner_of, dest_off)
print("remote form, every destination local:", {k: round(v[0], 3) for k, v in sc.profile().items()})
# routed form
pub = [ct.pub_base()]
owner_of, dest_off, cap, imp = pdist.plan_routed(cnt, [(0, nwin)], 0, pub)
sc.set_route(1, 0, owner_of, dest_off, cap, pub)
status = torch.zeros(4, dtype=torch.int32, device="cuda")
for rep in range(3):
    sc.reset(); sc.prime(None, 0)
    if rep == 2:
        sc.set_profiling(True)
    sc.scan_routed(d, status)
print("routed form, every destination local:", {k: round(v[0], 3) for k, v in sc.profile().items()}, "status", status.tolist())

#!/usr/bin/env python3
"""Golden vector of the packed table form (include/pykmer_b200.h: pk_table_pack_device / pk_table_unpack):
a seeded 8192-entry table with genome-like counts, a full chunk, an empty chunk and saturated entries, and its
packed form as oracle.pack_table lays it out (chunks in ascending order).  Pins the format across rounds:
tests/test_host_cpu.py rebuilds the table from these bytes with the library's host unpacker and checks
that the oracle still packs to exactly these bytes.
    python oracle/make_golden_packed.py        # writes tests/golden/packed_table.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402


def table() -> np.ndarray:
    rng = np.random.default_rng(20261019)
    t = np.minimum(rng.geometric(0.5, 8192), 255).astype(np.uint8)
    t[rng.random(8192) >= 0.25] = 0
    t[1024:2048] = rng.integers(1, 256, 1024, dtype=np.uint8)      # a chunk without a zero
    t[2048:3072] = 0                                               # a chunk without a count
    t[3072:3080] = 255
    t[8191] = 7
    return t


if __name__ == "__main__":
    t = table()
    bitmap, chunk_off, nz = oracle.pack_table(t)
    out = os.path.join(ROOT, "tests", "golden", "packed_table.npz")
    np.savez_compressed(out, table=t, bitmap=bitmap, chunk_off=chunk_off, nz=nz)
    print(out, t.size, "entries ->", bitmap.nbytes + chunk_off.nbytes + nz.nbytes, "packed bytes")

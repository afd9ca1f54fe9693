#!/usr/bin/env python3
"""Manufacture tests/golden/cnidaria/ by running the UNMODIFIED reference
calculate_distances_cnidaria.py (statholder :40-473, attachMethodName :475-549, init :552-569) on
a fixed list of (totalX, totalY, countX, countY, val) tuples: pairs taken from the committed golden
.kma matrices, degenerate tables (nothing shared, identical samples, empty samples), tables at
the scale of 4^15-entry indexes, and seeded random ones.

TEST INFRASTRUCTURE ONLY, build-container only (reads /root/reference).
Complex results (the reference's `** .5` of a negative number) are stored as [re, im].

usage: python oracle/make_golden_cnidaria.py
"""
import contextlib
import glob
import io
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PYKMER_REFERENCE", "/root/reference")


def cases():
    out = []
    for src in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "merger", "matrix_*.npz"))):
        m = np.load(src)["matrix"]
        n = m.shape[0]
        for k in range(n):
            for l in range(n):
                if k != l:
                    tk, tl, sh = (int(v) for v in m[k, l])
                    out.append((tk, tl, tk, tl, sh))
    out += [(0, 0, 0, 0, 0), (10, 10, 10, 10, 0), (10, 10, 10, 10, 10), (10, 0, 10, 0, 0), (1, 1, 1, 1, 1),
            (5, 9, 5, 9, 5), (9, 5, 9, 5, 5), (100, 100, 60, 70, 50), (2, 3, 2, 3, 1), (1000, 7, 1000, 7, 7)]
    rng = random.Random(0x5EED)
    for _ in range(300):                                   # K=15-sized tables, count = total
        a = rng.randrange(1, 1 << 28)
        b = rng.randrange(1, 1 << 28)
        sh = rng.randrange(0, min(a, b) + 1)
        out.append((a, b, a, b, sh))
    for _ in range(200):                                   # count < total (the cnidaria use)
        tx, ty = rng.randrange(1, 10 ** 6), rng.randrange(1, 10 ** 6)
        cx, cy = rng.randrange(0, tx + 1), rng.randrange(0, ty + 1)
        out.append((tx, ty, cx, cy, rng.randrange(0, min(cx, cy) + 1)))
    return out


def encode(v):
    if isinstance(v, complex):
        return [v.real, v.imag]
    return v


def main() -> None:
    sys.path.insert(0, REF)
    with contextlib.redirect_stdout(io.StringIO()):        # the module prints its method table on import
        import calculate_distances_cnidaria as ref
        names = sorted(ref.methods_available.keys())
        ref.init(names)
    tuples = cases()
    values = {}
    for name in names:
        f = ref.methods_enabled[name]
        col = []
        for (tx, ty, cx, cy, val) in tuples:
            dissi = {name: [[0.0]]}
            f(dissi, 0, 0, 0, tx, ty, cx, cy, val)
            col.append(encode(dissi[name][0][0]))
        values[name] = col
    out = os.path.join(ROOT, "tests", "golden", "cnidaria")
    os.makedirs(out, exist_ok=True)
    import gzip
    with gzip.GzipFile(os.path.join(out, "methods.json.gz"), "wb", mtime=0) as fh:
        fh.write(json.dumps({"methods": names, "cases": tuples, "values": values}).encode())
    print(f"{len(names)} methods x {len(tuples)} cases -> {out}/methods.json.gz")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Manufacture tests/golden/distance/ by running the UNMODIFIED reference calculate_distance.py
(calc_distance, calculate_distance.py:42-109) on the committed golden .kma matrices.

TEST INFRASTRUCTURE ONLY, build-container only (reads /root/reference).  scikit-bio and ete3 are
not installed here and calc_distance does not use them, so they are replaced by empty modules
for the import; the tree half of the reference (cluster_distance, which is all scikit-bio /
ete3 calls) cannot run and is therefore not part of the goldens.

usage: python oracle/make_golden_distance.py
"""
import glob
import os
import shutil
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PYKMER_REFERENCE", "/root/reference")


def main() -> None:
    for name, attrs in (("skbio", ("DistanceMatrix",)), ("skbio.tree", ("nj",)),
                        ("ete3", ("Tree", "TreeStyle", "TextFace"))):
        mod = types.ModuleType(name)
        for a in attrs:
            setattr(mod, a, None)
        sys.modules[name] = mod
    sys.path.insert(0, REF)
    import calculate_distance as ref                     # the reference module itself

    out = os.path.join(ROOT, "tests", "golden", "distance")
    os.makedirs(out, exist_ok=True)
    from pathlib import Path
    for src in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "merger", "matrix_*.npz"))):
        with tempfile.TemporaryDirectory() as tmp:
            kma = os.path.join(tmp, os.path.basename(src)[:-4] + ".kma")
            shutil.copy(src, kma)
            matrix = ref.get_matrix(Path(kma))
            with np.errstate(divide="ignore", invalid="ignore"):
                basefile, dist = ref.calc_distance(Path(kma), matrix, fill_diagonal=True)
            saved = np.load(f"{basefile}.npz")["distance"]
            assert np.array_equal(saved, dist, equal_nan=True)
            np.savez(os.path.join(out, os.path.basename(src)[:-4] + ".dist.jaccard.npz"), distance=dist)
            print(os.path.basename(src), dist.shape, float(np.nanmin(dist)), float(np.nanmax(dist)))


if __name__ == "__main__":
    main()

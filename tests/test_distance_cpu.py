"""Downstream consumer: .kma -> Jaccard distance -> neighbour-joining tree.
The Jaccard half is pinned against the reference's own calc_distance (tests/golden/distance/,
made by oracle/make_golden_distance.py); the tree half is checked through properties of the
neighbour-joining algorithm (additive distances are reproduced exactly), because scikit-bio and
ete3 -- which the reference calls for it -- are not installed."""
import glob
import itertools
import json
import os
import shutil

import numpy as np
import pytest

from pykmer_b200 import distance as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "merger", "matrix_*.npz")))


@pytest.mark.parametrize("case", CASES)
def test_jaccard_distance_matches_reference_golden(case, tmp_path):
    kma = tmp_path / (case + ".kma")
    shutil.copy(os.path.join(GOLD, "merger", case + ".npz"), kma)
    matrix = D.get_matrix(kma)
    basefile, dist = D.calc_distance(kma, matrix, fill_diagonal=True)
    want = np.load(os.path.join(GOLD, "distance", case + ".dist.jaccard.npz"))["distance"]
    assert dist.dtype == np.float64 and np.array_equal(dist, want, equal_nan=True)      # bit-exact
    assert str(basefile) == f"{kma}.dist.jaccard"
    assert np.array_equal(np.load(f"{basefile}.npz")["distance"], want, equal_nan=True)


def test_jaccard_formula_on_known_counts():
    m = np.zeros((2, 2, 3), dtype=np.uint64)
    m[0, 1] = (10, 20, 5)
    m[1, 0] = (20, 10, 5)
    m[0, 0] = m[1, 1] = (123456789, 7, 3)                 # the reference leaves the diagonal as garbage
    d = D.jaccard_distance(m)
    assert d[0, 0] == 0.0 and d[1, 1] == 0.0
    assert d[0, 1] == d[1, 0] == 1.0 - 5.0 / (10 + 20 - 5)


def _tree_distances(newick, ids):
    """Pairwise path lengths between the leaves of a Newick tree."""
    root = D._parse_newick(newick.strip())
    paths = {}

    def walk(n, trail):
        label, length, children = n
        trail = trail + [(id(n), length or 0.0)]
        if not children:
            paths[label] = trail
        for c in children:
            walk(c, trail)

    walk(root, [])
    out = np.zeros((len(ids), len(ids)))
    for (i, a), (j, b) in itertools.combinations(enumerate(ids), 2):
        pa, pb = paths[a], paths[b]
        k = 0
        while k < min(len(pa), len(pb)) and pa[k][0] == pb[k][0]:
            k += 1
        out[i, j] = out[j, i] = sum(l for _, l in pa[k:]) + sum(l for _, l in pb[k:])
    return out


def test_neighbor_joining_reproduces_an_additive_tree():
    # ((a:2, b:3):1, (c:4, (d:2, e:1):2):3) as an unrooted tree: additive distances
    ids = list("abcde")
    d = np.array([[0, 5, 10, 10, 9],
                  [5, 0, 11, 11, 10],
                  [10, 11, 0, 8, 7],
                  [10, 11, 8, 0, 3],
                  [9, 10, 7, 3, 0]], dtype=np.float64)
    newick = D.neighbor_joining(d, ids)
    assert newick.endswith(";") and all(newick.count(x + ":") == 1 for x in ids)
    assert np.allclose(_tree_distances(newick, ids), d, atol=1e-5)
    assert set(D.ascii_tree(newick).split()) >= {"\\-e", "+"} or "e" in D.ascii_tree(newick)


def test_neighbor_joining_three_taxa_and_errors():
    d = np.array([[0, 3, 4], [3, 0, 5], [4, 5, 0]], dtype=np.float64)
    assert D.neighbor_joining(d, ["x", "y", "z"]) == "(x:1.000000, y:2.000000, z:3.000000);"
    with pytest.raises(ValueError):
        D.neighbor_joining(d[:2, :2], ["x", "y"])


def test_cluster_outputs_and_cli(tmp_path):
    case = "matrix_K07_001-255"
    kma = tmp_path / "proj.001-255.kma"
    shutil.copy(os.path.join(GOLD, "merger", case + ".npz"), kma)
    n = D.get_matrix(kma).shape[0]
    names = [f"sample{i}.fa.gz" for i in range(n)]
    (tmp_path / "proj.001-255.kma.json").write_text(json.dumps(
        {"project_name": "proj", "min_count": 1, "max_count": 255,
         "data": [{"pos": i, "header": {"input_file_name": names[i]}} for i in range(n)]}))
    (tmp_path / "proj.001-255.kma.names.tsv").write_text("sample0.fa.gz\tSample zero\nbroken line\n")
    D.main([str(kma)])
    base = f"{kma}.dist.jaccard"
    want = np.load(os.path.join(GOLD, "distance", case + ".dist.jaccard.npz"))["distance"]
    assert np.array_equal(np.load(base + ".mat.redundant.np"), want)
    cond = np.load(base + ".mat.condensed.np")
    assert cond.shape == (n * (n - 1) // 2,) and np.array_equal(cond, want[np.triu_indices(n, 1)])
    assert np.allclose(np.loadtxt(base + ".mat.condensed.txt"), cond)
    rows = open(base + ".mat.redundant.lsmat").read().splitlines()
    assert rows[0].split("\t") == ["", "Sample zero"] + names[1:]              # names file applied
    assert [float(x) for x in rows[1].split("\t")[1:]] == want[0].tolist()
    newick = open(base + ".newick").read()
    assert newick.count(",") == n - 1 and "Sample zero:" in newick
    assert os.path.getsize(base + ".tree") > 0

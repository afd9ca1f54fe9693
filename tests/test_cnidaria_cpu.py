"""pykmer_b200/cnidaria_stats.py against the reference's calculate_distances_cnidaria.py
(goldens made by oracle/make_golden_cnidaria.py, which runs the reference itself): the same method
names, the same value bit for bit on 846 tables -- degenerate ones (ZeroDivisionError / ValueError
-> 1, :535-543) and the complex results of its `** .5` included."""
import gzip
import json
import math
import os

import numpy as np
import pytest

from pykmer_b200 import cnidaria_stats as cs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "cnidaria", "methods.json.gz")).read())


def _same(got, want) -> bool:
    if isinstance(want, list):                                   # complex, stored as [re, im]
        return isinstance(got, complex) and _same(got.real, want[0]) and _same(got.imag, want[1])
    if isinstance(got, complex):
        return False
    if isinstance(want, float) and math.isnan(want):
        return isinstance(got, float) and math.isnan(got)
    return got == want and (isinstance(got, float) == isinstance(want, float) or got == want)


def test_method_names_are_the_reference_s():
    assert sorted(cs.methods_available) == GOLD["methods"] and len(GOLD["methods"]) == 71


@pytest.mark.parametrize("name", GOLD["methods"])
def test_every_method_matches_the_reference_bit_for_bit(name):
    for case, want in zip(GOLD["cases"], GOLD["values"][name]):
        got = cs.evaluate(name, *case)
        assert _same(got, want), (name, case, got, want)


def test_reference_calling_convention_accumulates_into_dissi():
    """init / methods_enabled[name](dissi, num_kmers, x, y, totalX, totalY, countX, countY, val)
    (calculate_distances_cnidaria.py:483-547, 552-569): results are ADDED to dissi[name][x][y]."""
    cs.methods_enabled.clear()
    cs.init(["D_jaccard", "S_dice"])
    dissi = {"D_jaccard": [[0.0, 0.0], [0.0, 0.0]], "S_dice": [[0.0, 0.0], [0.0, 0.0]]}
    for _ in range(2):
        cs.methods_enabled["D_jaccard"](dissi, 0, 0, 1, 100, 100, 60, 70, 50)
    cs.methods_enabled["S_dice"](dissi, 0, 1, 0, 100, 100, 60, 70, 50)
    assert dissi["D_jaccard"][0][1] == 2 * (1 - 50 / 80) and dissi["D_jaccard"][1][0] == 0.0
    assert dissi["S_dice"][1][0] == 1.0 - 100.0 / 130.0
    assert cs.stats_data.a == 50 and cs.stats_data.b == 10 and cs.stats_data.c == 20 and cs.stats_data.d == 80
    with pytest.raises(KeyError):                                  # :557-560
        cs.init(["no_such_method"])
    with pytest.raises(KeyError):                                  # :478-481
        cs.attachMethodName("never_enabled", lambda: 0)


def test_apply_over_a_kma_matrix():
    m = np.load(os.path.join(ROOT, "tests", "golden", "merger", "matrix_K07_001-255.npz"))["matrix"]
    n = m.shape[0]
    out = cs.apply(m, ["D_jaccard", "S_ochiai_I", "D_hamming"])
    for k in range(n):
        for l in range(n):
            if k == l:
                assert out["D_jaccard"][k, l] == 0.0
                continue
            tk, tl, sh = (int(v) for v in m[k, l])
            assert out["D_jaccard"][k, l] == cs.evaluate("D_jaccard", tk, tl, tk, tl, sh)
            assert out["D_hamming"][k, l] == tk + tl - 2 * sh
    # Jaccard distance agrees with the reference's own consumer (calculate_distance.py:82-84)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "distance", "matrix_K07_001-255.dist.jaccard.npz"))
    key = list(gold.keys())[0]
    assert np.allclose(out["D_jaccard"], gold[key], rtol=0, atol=1e-15)
    with pytest.raises(KeyError):
        cs.apply(m, ["nope"])
    assert len(cs.apply(m[:2, :2])) == 71                          # every method, tiny matrix

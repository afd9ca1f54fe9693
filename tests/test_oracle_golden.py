"""Pins the CPU oracle against what the reference itself wrote (tests/golden/,
made by oracle/make_golden.py from /root/reference's indexer.py / merger.py)."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

INDEXER_CASES = sorted(
    os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLD, "indexer", "*.json"))
    if not os.path.basename(p).startswith("syn10M"))


def _split(case):
    fname, kk = case.rsplit(".", 1)
    return fname, int(kk)


@pytest.mark.parametrize("case", INDEXER_CASES)
@pytest.mark.parametrize("method", ["rolling", "direct"])
def test_indexer_matches_reference(case, method):
    fname, K = _split(case)
    if method == "direct" and K > 11:
        pytest.skip("direct form only on the small cases")
    gold = json.load(open(os.path.join(GOLD, "indexer", case + ".json")))
    res = oracle.index_fasta(os.path.join(GOLD, "inputs", fname), K, method=method)
    assert res["num_kmers"] == gold["num_kmers"]
    assert res["chromosomes"] == gold["chromosomes"]
    assert res["hist"] == gold["hist"]
    for k in ("hist_sum", "hist_count", "hist_min", "hist_max",
              "vals_sum", "vals_count", "vals_min", "vals_max"):
        assert res[k] == gold[k], k
    assert res["table"].size == gold["data_size"] == 4 ** K
    assert hashlib.sha256(res["table"].tobytes()).hexdigest() == gold["output_file_cheksum"]
    npz = os.path.join(GOLD, "indexer", case + ".kin.npz")
    if os.path.exists(npz):
        assert np.array_equal(np.load(npz)["table"], res["table"])
    assert oracle.frag_size_rule(K) == gold["frag_size"]


def test_stats_numpy_and_c_agree():
    rng = np.random.default_rng(3)
    t = rng.integers(0, 256, size=100_003, dtype=np.uint8)
    t[rng.random(t.size) < 0.8] = 0
    assert oracle.table_stats(t) == oracle.table_stats_numpy(t) == oracle.table_stats(t, threads=3)


def test_frag_size_rule_known_values():
    # values computed with the reference's own Header class
    for K, v in {3: 1000, 11: 2_098_000, 13: 33_555_000, 15: 357_914_000,
                 17: 954_438_000, 19: 999_557_000}.items():
        assert oracle.frag_size_rule(K) == v


def test_allkmers_closed_form():
    """test.py construction: table[i] == 2 iff i <= rc(i) (K odd => no palindromes)."""
    K = 5
    res = oracle.index_fasta(os.path.join(GOLD, "inputs", "allkmers_05.fasta.gz"), K)
    idx = np.arange(4 ** K)
    rc = np.zeros_like(idx)
    for p in range(K):
        rc |= (3 - ((idx >> (2 * p)) & 3)) << (2 * (K - 1 - p))
    assert np.array_equal(res["table"], np.where(idx <= rc, 2, 0).astype(np.uint8))
    assert res["num_kmers"] == 4 ** K and res["vals_count"] == 4 ** K // 2


def test_range_sharding_concatenates():
    """k-mer-range shards concatenate to the full table."""
    rng = np.random.default_rng(5)
    seq = np.frombuffer(b"ACGTNacgt", dtype=np.uint8)[rng.integers(0, 9, size=20_000)]
    K = 7
    full, n_full, _ = oracle.index_stream(seq, K)
    T = 4 ** K
    cuts = [0, T // 3 + 1, T // 2, T]
    parts, total = [], 0
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        t, n, _ = oracle.index_stream(seq, K, range_lo=lo, range_hi=hi)
        parts.append(t)
        total += n
    assert np.array_equal(np.concatenate(parts), full) and total == n_full


def test_threaded_rolling_equals_serial():
    rng = np.random.default_rng(6)
    seq = np.frombuffer(b"ACGTNa", dtype=np.uint8)[rng.integers(0, 6, size=300_000)]
    seq[1000:2500] = ord("A")
    a, na, _ = oracle.index_stream(seq, 9)
    b, nb, _ = oracle.index_stream(seq, 9, method="mt", threads=4)
    assert na == nb and np.array_equal(a, b)


MERGER_CASES = sorted(glob.glob(os.path.join(GOLD, "merger", "matrix_*.npz")))


@pytest.mark.parametrize("path", MERGER_CASES, ids=[os.path.basename(p) for p in MERGER_CASES])
def test_merger_matches_reference(path):
    gold = np.load(path)["matrix"]
    meta = json.load(open(path[:-4] + ".json"))
    samples = np.load(os.path.join(GOLD, "merger", "samples_K07.npz"))
    tables = samples["tables"]
    assert [str(n) for n in samples["names"]] == meta["order"]
    lo, hi = meta["min_count"], meta["max_count"]
    m = oracle.merge_matrix(tables, lo, hi, threads=2)
    N = m.shape[0]
    off = ~np.eye(N, dtype=bool)
    assert gold.dtype == np.uint64 and gold.shape == (N, N, 3)
    assert np.array_equal(m[off], gold[off])
    # the Gram restatement gives the same off-diagonal triples
    bits = np.stack([oracle.threshold_pack(t, lo, hi) for t in tables])
    m2 = oracle.matrix_from_gram(oracle.gram_from_bits(bits))
    assert np.array_equal(m2[off], gold[off])
    assert np.array_equal(m2, m)
    # one pair through the NumPy form the reference uses
    assert oracle.pair_counts(tables[0], tables[3], lo, hi) == \
        oracle.pair_counts_numpy(tables[0], tables[3], lo, hi) == tuple(int(v) for v in gold[0, 3])


def test_config1_syn10M(tmp_path):
    """BASELINE config 1 (10 Mbp bgzip multi-FASTA, K=11): the FASTA is regenerated
    from its seed, must hash to what the reference was fed, and the oracle must
    reproduce the reference's .kin digest and statistics."""
    from pykmer_b200 import synth
    gold = json.load(open(os.path.join(GOLD, "indexer", "syn10M.fa.bgz.11.json")))
    src = str(tmp_path / "syn10M.fa.bgz")
    synth.write_fasta(src, synth.syn10m_records(), line_width=60, level=1)
    assert hashlib.sha256(open(src, "rb").read()).hexdigest() == gold["fasta_sha256"]
    res = oracle.index_fasta(src, 11)
    assert res["num_kmers"] == gold["num_kmers"]
    assert res["chromosomes"] == gold["chromosomes"]
    assert res["hist"] == gold["hist"]
    assert res["vals_sum"] == gold["vals_sum"] and res["vals_max"] == gold["vals_max"]
    assert hashlib.sha256(res["table"].tobytes()).hexdigest() == gold["output_file_cheksum"]


def test_tiled_mask_layout_formula():
    """The tiled layout the merger packs into (include/pykmer_b200.h): word g of sample r sits at
    (g // 32) * N * 32 + r * 32 + g % 32; Gram matrices do not depend on the layout."""
    rng = np.random.default_rng(9)
    for n, words in ((1, 1), (3, 31), (5, 32), (7, 33), (50, 100)):
        rows = rng.integers(0, 2 ** 32, size=(n, words), dtype=np.uint64).astype(np.uint32)
        flat = oracle.tile_masks(rows)
        assert flat.size == -(-words // 32) * n * 32
        for r in range(n):
            for g in (0, words // 2, words - 1):
                assert flat[(g // 32) * n * 32 + r * 32 + g % 32] == rows[r, g]
        back = flat.reshape(-1, n, 32).transpose(1, 0, 2).reshape(n, -1)
        assert np.array_equal(back[:, :words], rows) and not back[:, words:].any()
        assert np.array_equal(oracle.gram_from_bits(back), oracle.gram_from_bits(rows))

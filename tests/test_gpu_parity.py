"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the
committed golden outputs of the reference.  Everything is integer work: the bar is
bit-exact equality."""
import glob
import hashlib
import json
import os
import shutil

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    from oracle import oracle
    from pykmer_b200 import device, _native
    assert _native.device_count() >= 1
    return {"torch": torch, "oracle": oracle, "dev": device, "nat": _native}


def _index_stream(dev, stream, K, lo=0, hi=None, starts=None, pieces=None, host=False, mode=0,
                  window_log2=None, pool_log2=None, flush=None, ovf_log2=None, est=None):
    import torch
    hi = 4 ** K if hi is None else hi
    # test hooks of the library: small table windows / small k-mer buffer / flush scheme
    # (read at create time)
    for name, val in (("PYKMER_B200_WINDOW_LOG2", window_log2), ("PYKMER_B200_POOL_LOG2", pool_log2),
                      ("PYKMER_B200_FLUSH", flush), ("PYKMER_B200_OVF_LOG2", ovf_log2),
                      ("PYKMER_B200_EST", est)):
        if val is None:
            os.environ.pop(name, None)
        else:
            os.environ[name] = str(val)
    try:
        ix = dev.Indexer(K, device=0, range_lo=lo, range_hi=hi, mode=mode)
    finally:
        os.environ.pop("PYKMER_B200_WINDOW_LOG2", None)
        os.environ.pop("PYKMER_B200_POOL_LOG2", None)
        os.environ.pop("PYKMER_B200_FLUSH", None)
        os.environ.pop("PYKMER_B200_OVF_LOG2", None)
        os.environ.pop("PYKMER_B200_EST", None)
    with ix:
        if starts is not None:
            ix.set_records(starts)
        cuts = [0, len(stream)] if not pieces else [0] + sorted(pieces) + [len(stream)]
        keep = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            part = np.ascontiguousarray(stream[a:b])
            if host:
                ix.feed_host(part)
            elif b > a:
                t = torch.from_numpy(part.copy()).cuda()
                keep.append(t)
                ix.feed_device(t)
        hist, st = ix.finalize()
        flags = ix.record_flags() if starts is not None else None
        table = ix.table_to_host().numpy().copy()
    return table, hist, st, flags


INDEXER_CASES = sorted(os.path.basename(p)[:-5] for p in glob.glob(os.path.join(GOLD, "indexer", "*.json"))
                       if not os.path.basename(p).startswith("syn10M"))


@pytest.mark.parametrize("case", INDEXER_CASES)
def test_indexer_stream_matches_reference_golden(env, case):
    """cleaned stream -> CUDA -> table/stats == what the reference's indexer.py wrote."""
    from pykmer_b200 import fasta
    fname, kk = case.rsplit(".", 1)
    K = int(kk)
    gold = json.load(open(os.path.join(GOLD, "indexer", case + ".json")))
    stream, names, starts, lengths = fasta.read_fasta_stream(os.path.join(GOLD, "inputs", fname))
    table, hist, st, flags = _index_stream(env["dev"], stream, K, starts=starts)
    assert st["num_kmers"] == gold["num_kmers"]
    assert hist == gold["hist"]
    for k in ("vals_sum", "vals_count", "vals_min", "vals_max"):
        assert st[k] == gold[k], k
    assert hashlib.sha256(table.tobytes()).hexdigest() == gold["output_file_cheksum"]
    chrom = [[names[i], lengths[i]] for i in range(len(names)) if flags[i]]
    assert chrom == gold["chromosomes"]


@pytest.mark.parametrize("case", ["tiny_mixed.fa.07", "saturating.fa.gz.11", "rand200k.fa.bgz.13",
                                  "allkmers_05.fasta.gz.05"])
def test_indexer_cli_writes_reference_files(env, case, tmp_path):
    """indexer.py CLI: .kin bytes and every deterministic .kin.json key equal the reference's."""
    from pykmer_b200 import indexer
    fname, kk = case.rsplit(".", 1)
    K = int(kk)
    gold = json.load(open(os.path.join(GOLD, "indexer", case + ".json")))
    src = str(tmp_path / fname)
    shutil.copy(os.path.join(GOLD, "inputs", fname), src)
    indexer.main([src, "sample", str(K)])
    kin = f"{src}.{K:02d}.kin"
    assert os.path.exists(kin) and not os.path.exists(kin + ".tmp")
    meta = json.load(open(kin + ".json"))
    assert sorted(meta.keys()) == gold["all_keys"]
    for k, v in gold.items():
        if k in ("all_keys",):
            continue
        got = meta[k]
        if k == "project_name":
            got = os.path.basename(got)
        assert got == v, k
    assert gen_sha(kin) == gold["output_file_cheksum"] == meta["output_file_cheksum"]
    # the reference's own self-check (read_fasta_index -> Header.check_data) passes
    indexer.read_fasta_index(src, input_file=src, kmer_len=K)


def gen_sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        for blk in iter(lambda: fh.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def test_indexer_config1_syn10M(env, tmp_path):
    """BASELINE config 1 end to end through the CLI: 10 Mbp bgzip multi-FASTA, K=11."""
    from pykmer_b200 import indexer, synth
    gold = json.load(open(os.path.join(GOLD, "indexer", "syn10M.fa.bgz.11.json")))
    src = str(tmp_path / "syn10M.fa.bgz")
    synth.write_fasta(src, synth.syn10m_records(), line_width=60, level=1)
    assert gen_sha(src) == gold["fasta_sha256"]
    indexer.main([src, "syn10M", "11"])
    meta = json.load(open(src + ".11.kin.json"))
    for k in ("num_kmers", "chromosomes", "hist", "hist_sum", "hist_count", "hist_min", "hist_max",
              "vals_sum", "vals_count", "vals_min", "vals_max", "output_file_cheksum",
              "input_file_cheksum", "frag_size", "flush_every", "data_size"):
        assert meta[k] == gold[k], k
    assert gen_sha(src + ".11.kin") == gold["output_file_cheksum"]


def _random_stream(rng, n, alphabet=b"ACGTacgtACGTACGTNn>", runs=True):
    a = np.frombuffer(alphabet, dtype=np.uint8)
    s = a[rng.integers(0, len(a), size=n)].copy()
    if runs and n > 5000:
        s[1000:2600] = ord("A")                       # saturates a counter
        s[3000:4200] = np.frombuffer(b"AT" * 600, dtype=np.uint8)
    return s


@pytest.mark.parametrize("K", [1, 3, 5, 7, 9, 11, 13])
@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 495, 496, 497, 100_003])
def test_indexer_random_streams_vs_oracle(env, K, n):
    rng = np.random.default_rng(1000 * K + n)
    s = _random_stream(rng, n)
    want, num, _ = env["oracle"].index_stream(s, K)
    table, hist, st, _ = _index_stream(env["dev"], s, K)
    assert st["num_kmers"] == num
    assert np.array_equal(table, want)
    oh, ost = env["oracle"].table_stats(want)
    assert hist == oh and all(st[k] == ost[k] for k in ("vals_sum", "vals_count", "vals_min", "vals_max"))


@pytest.mark.parametrize("K", [15, 17, 19, 21, 25, 31])
def test_indexer_wide_k_on_a_range_vs_oracle(env, K):
    """K > 13: the table is too big for a unit test, so count one k-mer range
    (exactly what a shard does) over a sequence built to land in it."""
    rng = np.random.default_rng(K)
    parts = []
    for _ in range(400):
        parts.append(np.full(int(rng.integers(K - 3, 3 * K)), ord("A"), dtype=np.uint8))
        parts.append(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(rng.integers(1, 12)))])
        if rng.random() < 0.1:
            parts.append(np.frombuffer(b"N", dtype=np.uint8))
    s = np.concatenate(parts)
    lo, hi = 0, 1 << 22
    want, num, _ = env["oracle"].index_stream(s, K, range_lo=lo, range_hi=hi)
    assert num > 1000
    table, hist, st, _ = _index_stream(env["dev"], s, K, lo=lo, hi=hi)
    assert st["num_kmers"] == num and np.array_equal(table, want)
    lo2 = (4 ** K) // 2 - (1 << 20)
    lo2 -= lo2 % 4
    want2, num2, _ = env["oracle"].index_stream(s, K, range_lo=lo2, range_hi=lo2 + (1 << 21))
    t2, _, st2, _ = _index_stream(env["dev"], s, K, lo=lo2, hi=lo2 + (1 << 21))
    assert st2["num_kmers"] == num2 and np.array_equal(t2, want2)


@pytest.mark.parametrize("K", [5, 11, 17])
@pytest.mark.parametrize("host", [False, True])
def test_indexer_chunked_feeds_carry_the_window(env, K, host):
    rng = np.random.default_rng(77 + K)
    s = _random_stream(rng, 50_000)
    hi = min(4 ** K, 1 << 22)
    want, num, _ = env["oracle"].index_stream(s, K, range_hi=hi)
    cuts = sorted(set(rng.integers(1, len(s) - 1, size=40).tolist()) | {1, 2, 3, 17, 33, 48, 49, 64})
    table, _, st, _ = _index_stream(env["dev"], s, K, hi=hi, pieces=cuts, host=host)
    assert st["num_kmers"] == num and np.array_equal(table, want)


def test_indexer_record_flags_many_records(env):
    """test.py construction at K=7: 16384 records of exactly K bases, every one listed."""
    from pykmer_b200 import fasta
    stream, names, starts, lengths = fasta.read_fasta_stream(os.path.join(GOLD, "inputs", "allkmers_07.fasta.gz"))
    for K, expect_all in ((7, True), (9, False)):
        _, _, st, flags = _index_stream(env["dev"], stream, K, starts=starts)
        _, num, oflags = env["oracle"].index_stream(stream, K, rec_starts=starts)
        assert st["num_kmers"] == num and np.array_equal(flags, oflags)
        assert bool(flags.all()) == expect_all


def test_indexer_range_shards_concatenate(env):
    rng = np.random.default_rng(5)
    s = _random_stream(rng, 200_000)
    K = 9
    T = 4 ** K
    full, _, st_full, _ = _index_stream(env["dev"], s, K)
    cuts = [0, T // 8, T // 2 + 4, T]
    parts, total = [], 0
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        t, _, st, _ = _index_stream(env["dev"], s, K, lo=lo, hi=hi)
        parts.append(t)
        total += st["num_kmers"]
    assert np.array_equal(np.concatenate(parts), full) and total == st_full["num_kmers"]


def test_indexer_reset_and_reuse(env):
    import torch
    rng = np.random.default_rng(9)
    dev, oracle = env["dev"], env["oracle"]
    with dev.Indexer(9) as ix:
        for rep in range(3):
            s = _random_stream(rng, 30_000 + rep)
            t = torch.from_numpy(s).cuda()
            ix.reset()
            ix.feed_device(t)
            hist, st = ix.finalize()
            want, num, _ = oracle.index_stream(s, 9)
            assert st["num_kmers"] == num
            assert np.array_equal(ix.table_to_host().numpy(), want)
        assert ix.launch_count() >= 6          # scan + carry per feed (the histogram rides along)


def test_indexer_rejects_bad_arguments(env):
    dev, nat = env["dev"], env["nat"]
    for K in (0, 4, 33, -3):
        with pytest.raises(ValueError):
            dev.Indexer(K)
    with pytest.raises(ValueError):
        dev.Indexer(5, range_lo=10, range_hi=5)
    with pytest.raises(ValueError):
        dev.Indexer(5, range_hi=4 ** 5 + 1)


@pytest.mark.parametrize("n", [1, 15, 16, 31, 4096, 1_000_003])
def test_table_stats_vs_oracle(env, n):
    rng = np.random.default_rng(n)
    t = rng.integers(0, 256, size=n, dtype=np.uint8)
    t[rng.random(n) < 0.7] = 0
    t[rng.random(n) < 0.1] = 1
    hist, st = env["dev"].table_stats(t)
    oh, ost = env["oracle"].table_stats(t)
    assert hist == oh
    assert st == (ost["vals_sum"], ost["vals_count"], ost["vals_min"], ost["vals_max"])
    full = np.full(max(n, 16), 7, dtype=np.uint8)            # no zero entry: vals_min = 7
    assert env["dev"].table_stats(full)[1][2] == 7


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 65_536, 1_000_003])
@pytest.mark.parametrize("lohi", [(1, 255), (2, 10), (255, 255), (1, 1), (7, 3)])
def test_threshold_pack_vs_oracle(env, n, lohi):
    import torch
    rng = np.random.default_rng(n)
    t = rng.integers(0, 256, size=n, dtype=np.uint8)
    t[rng.random(n) < 0.5] = 0
    lo, hi = lohi
    bits = env["dev"].threshold_pack(torch.from_numpy(t).cuda(), lo, hi)
    want = env["oracle"].threshold_pack(t, lo, hi)
    assert np.array_equal(bits.cpu().numpy().view(np.uint32), want)


@pytest.fixture(params=["i8", "popc"])
def gram_algo(request):
    """The two integer Gram implementations on row-major masks: tcgen05 kind::i8 tensor cores and
    AND + popcount.  (The merger's default -- tcgen05 kind::mxf4 on tiled masks -- is tested below.)"""
    os.environ["PYKMER_B200_GRAM"] = request.param
    yield request.param
    os.environ.pop("PYKMER_B200_GRAM", None)


@pytest.mark.parametrize("N,words", [(1, 1), (2, 3), (3, 64), (5, 65), (50, 1000), (64, 257),
                                     (65, 130), (130, 70), (255, 33), (50, 150_001), (128, 40_000),
                                     (255, 30_000), (256, 9_999)])
def test_gram_vs_oracle(env, gram_algo, N, words):
    import torch
    rng = np.random.default_rng(N * 1000 + words)
    stride = (words + 3) & ~3
    bits = rng.integers(0, 2 ** 32, size=(N, stride), dtype=np.uint64).astype(np.uint32)
    bits[:, words:] = 0xFFFFFFFF                                  # padding must be ignored
    d = torch.from_numpy(bits.view(np.int32)).cuda()
    G = env["dev"].gram(d, words=words).cpu().numpy()
    if N * words > 2_000_000:                                     # big cases: exact Gram by float64 matmul
        B = np.unpackbits(bits[:, :words].view(np.uint8), axis=1, bitorder="little").astype(np.float32)
        want = (B.astype(np.float64) @ B.T.astype(np.float64)).astype(np.int64)
    else:
        want = env["oracle"].gram_from_bits(bits[:, :words])
    assert np.array_equal(G, want)
    G2 = env["dev"].gram(d, words=words, out=torch.from_numpy(want.copy()).cuda(), accumulate=True)
    assert np.array_equal(G2.cpu().numpy(), 2 * want)


def _tile(rows_t):
    """(N, words) row-major int32 CUDA masks -> the tiled buffer (words a multiple of 32)."""
    N, words = rows_t.shape
    return rows_t.view(N, words // 32, 32).permute(1, 0, 2).contiguous().view(-1)


@pytest.mark.parametrize("N", [3, 100, 130])
def test_gram_f4_every_partial_sum_is_exact(env, N):
    """All-ones masks, 2^19 words per accumulator (<= 64 samples: a CTA runs two half slabs, each
    with its own accumulator block): the FP32 accumulators walk through every multiple of 64 up to
    2^24 -- the largest value an accumulator may reach -- so any lost low bit shows up in G.
    (The odd values are walked in tests/test_gpu_at_scale.py.)"""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    words = sms << (20 if N <= 64 else 19)
    d = torch.full((N, words), -1, dtype=torch.int32, device="cuda")
    d[N - 1, 1::2] = 0x55555555                                   # one sample with half the bits
    G = env["dev"].gram_tiled(_tile(d), N, words).cpu().numpy()
    want = np.full((N, N), 32 * words, dtype=np.int64)
    want[N - 1, :] = want[:, N - 1] = 16 * words + 16 * (words // 2)
    assert np.array_equal(G, want)


def test_gram_f4_exactness_check_and_fallback(env, tmp_path):
    """pk_gram_tiled_exact: the once-per-device self-check passes on this GPU; a process in which it
    fails (test hook) refuses the FP4 kernel loudly and the merger takes the integer kernels."""
    import subprocess
    import sys
    assert env["dev"].gram_tiled_exact(0) is True
    code = (
        "import numpy as np, torch\n"
        "from pykmer_b200 import device as dev, _native as nat, synth\n"
        "assert dev.gram_tiled_exact(0) is False and not dev.use_tiled_masks(6)\n"
        "try:\n"
        "    dev.gram_tiled(dev.tiled_masks(32, 3), 3, 32)\n"
        "    raise SystemExit('the FP4 kernel ran on a device marked inexact')\n"
        "except nat.PkError as e:\n"
        "    assert 'exact' in str(e)\n"
        "from oracle import oracle\n"
        "tables = np.stack([synth.synth_table(s, 9) for s in range(6)])\n"
        "m = dev.merge_host(list(tables), 1, 50, device=0)\n"
        "assert np.array_equal(m, oracle.merge_matrix(tables, 1, 50))\n"
        "print('fallback ok')\n")
    res = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, PYKMER_B200_F4_EXACT="0"))
    assert res.returncode == 0 and "fallback ok" in res.stdout, res.stderr[-2000:]


@pytest.mark.parametrize("N,n", [(1, 5), (2, 100), (3, 1024), (7, 33_000), (50, 70_001), (64, 40_000),
                                 (65, 50_000), (128, 98_304), (200, 77_777), (256, 40_000), (50, 3_000_001)])
def test_tiled_masks_pack_and_gram_vs_oracle(env, N, n):
    """The merger's default path for <= 256 samples: tables -> tiled masks (packed in two slabs) ->
    FP4 Gram.  Layout and matrix against the oracle, table lengths that are not whole words / tiles."""
    import torch
    dev, oracle = env["dev"], env["oracle"]
    rng = np.random.default_rng(N * 7919 + n)
    tables = (rng.integers(0, 12, size=(N, n), dtype=np.uint8) * (rng.random((N, n)) < 0.4)).astype(np.uint8)
    lo, hi = 2, 9
    words = (n + 31) // 32
    bits = dev.tiled_masks(words, N)
    cut = (n // 2) // 1024 * 1024                                 # slabs start on whole words
    for s in range(N):
        t = torch.from_numpy(tables[s]).cuda()
        if cut:
            dev.threshold_pack_tiled(t[:cut].contiguous(), lo, hi, bits, s, N)
            dev.threshold_pack_tiled(t[cut:].contiguous(), lo, hi, bits, s, N, first_word=cut // 32)
        else:
            dev.threshold_pack_tiled(t, lo, hi, bits, s, N)
    rows = np.stack([oracle.threshold_pack(tables[s], lo, hi) for s in range(N)])
    want_tiled = oracle.tile_masks(rows)
    assert np.array_equal(bits.cpu().numpy().view(np.uint32)[:want_tiled.size], want_tiled)
    G = dev.gram_tiled(bits, N, words).cpu().numpy()
    if N * words > 2_000_000:
        B = np.unpackbits(rows.view(np.uint8), axis=1, bitorder="little").astype(np.float64)
        want = (B @ B.T).astype(np.int64)
    else:
        want = oracle.gram_from_bits(rows)
    assert np.array_equal(G, want)
    G2 = dev.gram_tiled(bits, N, words, out=torch.from_numpy(want.copy()).cuda(), accumulate=True)
    assert np.array_equal(G2.cpu().numpy(), 2 * want)


@pytest.mark.parametrize("N,words", [(257, 64), (300, 4096), (384, 2048), (513, 320), (640, 96)])
def test_tiled_gram_more_than_256_samples_vs_oracle(env, N, words):
    """More than 256 samples stay on the tensor cores: 128-row block pairs of one tiled buffer
    (merger.py:139-153 merges any N)."""
    import torch
    dev = env["dev"]
    rng = np.random.default_rng(N + words)
    rows = rng.integers(0, 2 ** 32, size=(N, words), dtype=np.uint64).astype(np.uint32)
    rows[rng.random(N) < 0.1] = 0                                  # a few empty samples
    d = torch.from_numpy(rows.view(np.int32)).cuda()
    G = dev.gram_tiled(_tile(d), N, words).cpu().numpy()
    B = np.unpackbits(rows.view(np.uint8), axis=1, bitorder="little").astype(np.float32)
    want = (B.astype(np.float64) @ B.T.astype(np.float64)).astype(np.int64)
    assert np.array_equal(G, want)
    G2 = dev.gram_tiled(_tile(d), N, words, out=torch.from_numpy(want.copy()).cuda(), accumulate=True)
    assert np.array_equal(G2.cpu().numpy(), 2 * want)
    with pytest.raises(ValueError):
        dev.gram_tiled(dev.tiled_masks(1, 4097), 4097, 1)


MERGER_CASES = sorted(glob.glob(os.path.join(GOLD, "merger", "matrix_*.npz")))


@pytest.mark.parametrize("path", MERGER_CASES, ids=[os.path.basename(p) for p in MERGER_CASES])
def test_merge_matches_reference_golden(env, path):
    gold = np.load(path)["matrix"]
    meta = json.load(open(path[:-4] + ".json"))
    tables = np.load(os.path.join(GOLD, "merger", "samples_K07.npz"))["tables"]
    lo, hi = meta["min_count"], meta["max_count"]
    m = env["dev"].merge_host(list(tables), lo, hi)
    N = m.shape[0]
    off = ~np.eye(N, dtype=bool)
    assert m.dtype == np.uint64 and np.array_equal(m[off], gold[off])
    assert np.array_equal(m, env["oracle"].merge_matrix(tables, lo, hi))
    assert env["dev"].pair_counts(tables[1], tables[4], lo, hi) == tuple(int(v) for v in gold[1, 4])


@pytest.mark.parametrize("budget", [1, 4096, 40_000])
def test_merge_in_k_axis_chunks_matches_reference_golden(env, budget, tmp_path):
    """A working set larger than the device memory set aside for masks (K=17 x 255 samples would be
    548 GB) is contracted chunk by chunk along the k-mer axis and summed, like the reference's own
    100 M-entry blocks (tools.py:449-489).  Forced here with tiny mask budgets -- 1 byte = one
    1024-k-mer tile per chunk, 16 chunks at K=7 -- through pk_merge_host and through
    merger.merge_tables, against the reference's own .kma matrices and the oracle."""
    from pykmer_b200 import merger, tools
    tables = np.load(os.path.join(GOLD, "merger", "samples_K07.npz"))["tables"]
    N = tables.shape[0]
    off = ~np.eye(N, dtype=bool)
    os.environ["PYKMER_B200_MASK_BUDGET"] = str(budget)
    try:
        for path in MERGER_CASES[:3]:
            gold = np.load(path)["matrix"]
            meta = json.load(open(path[:-4] + ".json"))
            m = env["dev"].merge_host(list(tables), meta["min_count"], meta["max_count"])
            assert np.array_equal(m[off], gold[off])
    finally:
        os.environ.pop("PYKMER_B200_MASK_BUDGET", None)
    # the CLI's route: files on disk, merge_tables with the same budget
    headers = []
    for s in range(N):
        src = tmp_path / f"s{s}.fa"
        src.write_text(">x\nACGT\n")
        h = tools.Header(str(src), input_file=str(src), kmer_len=7)
        tables[s].tofile(h.index_file_root)
        headers.append(h)
    want = env["oracle"].merge_matrix(tables, 2, 9)
    for tail in (budget, None):
        m = merger.merge_tables(headers, 2, 9, mask_budget_bytes=tail, slab_bytes=4096)
        assert np.array_equal(m, want)


def test_merger_cli_writes_reference_files(env, tmp_path):
    """merger.py CLI on .kin/.kin.bgz files: .kma matrix and .kma.json layout as the reference."""
    import gzip
    from pykmer_b200 import merger
    from pykmer_b200.tools import Header
    samples = np.load(os.path.join(GOLD, "merger", "samples_K07.npz"))
    template = json.load(open(os.path.join(GOLD, "indexer", "tiny_mixed.fa.07.json")))
    kins = []
    for name, table in zip(samples["names"], samples["tables"]):
        name = str(name)
        packed = name.endswith(".bgz")
        # same directory layout as the golden run, so that sorting the paths gives its order
        sub = tmp_path / "merge" if name.startswith("synth") else tmp_path
        sub.mkdir(exist_ok=True)
        kin = str(sub / (name[:-4] if packed else name))
        base = kin[:-len(".07.kin")]
        open(base, "w").close()
        meta = {k: template.get(k) for k in Header.HEADER_FIXED + Header.HEADER_DATA}
        meta.update(input_file_path=base, input_file_name=os.path.basename(base), kmer_len=7)
        json.dump(meta, open(kin + ".json", "w"))
        if packed:
            with gzip.open(kin + ".bgz", "wb") as fz:
                fz.write(table.tobytes())
            kins.append(kin + ".bgz")
        else:
            table.tofile(kin)
            kins.append(kin)
    for lo, hi in ((2, 10), (1, 255)):
        gold = np.load(os.path.join(GOLD, "merger", f"matrix_K07_{lo:03d}-{hi:03d}.npz"))["matrix"]
        gmeta = json.load(open(os.path.join(GOLD, "merger", f"matrix_K07_{lo:03d}-{hi:03d}.json")))
        proj = str(tmp_path / f"proj{lo}")
        merger.main([proj] + list(reversed(kins)) + [f"--min-count={lo}", f"--max-count={hi}"])
        kma = f"{proj}.{lo:03d}-{hi:03d}.kma"
        m = np.load(kma)["matrix"]
        desc = json.load(open(kma + ".json"))
        N = m.shape[0]
        off = ~np.eye(N, dtype=bool)
        assert m.dtype == np.uint64 and m.shape == gold.shape and np.array_equal(m[off], gold[off])
        assert sorted(desc.keys()) == gmeta["top_keys"]
        assert sorted(desc["data"][0].keys()) == gmeta["data_keys"]
        assert sorted(desc["data"][0]["header"].keys()) == gmeta["header_keys"]
        assert [os.path.basename(d["index_file"]) for d in desc["data"]] == gmeta["order"]
        assert [d["pos"] for d in desc["data"]] == gmeta["pos"]
        with pytest.raises(AssertionError):                     # refuses to overwrite (merger.py:99)
            merger.main([proj] + kins + [f"--min-count={lo}", f"--max-count={hi}"])
    # one pair through the mirror of merger.calculate_distance
    assert merger.calculate_distance(kins[0], kins[1], 2, 10) == \
        env["oracle"].pair_counts(samples["tables"][0], samples["tables"][1], 2, 10)
    # and through Header.calculate_distance2 (tools.py:495-512) 
    from pykmer_b200.tools import Header
    h0, h1 = Header("p", index_file=kins[0]), Header("p", index_file=kins[1])
    assert h0.calculate_distance2(h1, 2, 10) == h0.calculate_distance(h1, 2, 10) == \
        env["oracle"].pair_counts(samples["tables"][0], samples["tables"][1], 2, 10)


def test_synth_table_kernel_matches_numpy(env):
    from pykmer_b200 import synth
    for s in (0, 3, 49, 254):
        d = env["dev"].synth_table(s, 12_345, 12_345 + 200_000)
        assert np.array_equal(d.cpu().numpy(), synth.synth_table_slice(s, 12_345, 12_345 + 200_000))


def test_merge_pipeline_synthetic_k11_vs_oracle(env):
    """50 synthetic samples at K=11, --max-count=50 (BASELINE config 3 at reduced K)."""
    import torch
    from pykmer_b200 import synth
    dev = env["dev"]
    K, N = 11, 50
    T = 4 ** K
    words = T // 32
    bits = torch.zeros((N, words), dtype=torch.int32, device="cuda")
    tables = np.empty((N, T), dtype=np.uint8)
    for s in range(N):
        d = dev.synth_table(s, 0, T)
        dev.threshold_pack(d, 1, 50, out=bits[s])
        tables[s] = d.cpu().numpy()
    assert np.array_equal(tables[7], synth.synth_table(7, K))
    m = dev.matrix_from_gram(dev.gram(bits).cpu().numpy())
    want = env["oracle"].merge_matrix(tables, 1, 50, threads=env["oracle"].max_threads())
    assert np.array_equal(m, want)


def test_indexer_scaled_config2_vs_oracle(env):
    """BASELINE config 2 at 1/16 scale (49 Mbp tomato-like multi-FASTA stream, K=15, full
    1 GiB table): table bytes, num_kmers, record flags and statistics against the oracle."""
    import torch
    from pykmer_b200 import synth
    dev, oracle = env["dev"], env["oracle"]
    recs = synth.syn782m_records(scale=1 / 16)
    stream, starts, lengths, names = synth.records_to_stream(recs)
    want, num, _ = oracle.index_stream(stream, 15, method="mt", threads=oracle.max_threads())
    table, hist, st, flags = _index_stream(dev, stream, 15, starts=starts, host=True)
    assert st["num_kmers"] == num and flags.all()
    assert np.array_equal(table, want)
    oh, ost = oracle.table_stats(want)
    assert hist == oh and st["vals_sum"] == ost["vals_sum"] and st["vals_max"] == 255
    # size-independent properties of any index
    assert st["vals_sum"] <= st["num_kmers"] and sum(hist) == st["vals_count"]
    assert st["vals_min"] == 0


# ------------------------------------------------------------------ PARTITION counting mode

PART = 2   # PK_MODE_PARTITION


@pytest.mark.parametrize("K,wlog", [(5, 4), (5, 24), (9, 6), (9, 10), (11, 10), (11, 16), (13, 14), (13, 24)])
@pytest.mark.parametrize("n", [0, 17, 4000, 100_003, 1_000_003])
@pytest.mark.parametrize("plog", [12, 30])
@pytest.mark.parametrize("flush", ["smem", "l2", "byte"])
def test_partition_mode_random_streams_vs_oracle(env, K, wlog, n, plog, flush):
    """Window-partitioned counting (many windows, and a tiny k-mer buffer that forces
    repeated saturating flushes) gives the oracle's table, statistics and num_kmers."""
    if plog == 12 and n > 200_000:
        pytest.skip("tiny buffer only on the smaller streams")
    rng = np.random.default_rng(31 * K + n + wlog)
    s = _random_stream(rng, n)
    want, num, _ = env["oracle"].index_stream(s, K)
    table, hist, st, _ = _index_stream(env["dev"], s, K, mode=PART, window_log2=wlog, pool_log2=plog, flush=flush)
    assert st["num_kmers"] == num
    assert np.array_equal(table, want)
    oh, ost = env["oracle"].table_stats(want)
    assert hist == oh and all(st[k] == ost[k] for k in ("vals_sum", "vals_count", "vals_min", "vals_max"))


@pytest.mark.parametrize("case", ["tiny_mixed.fa.07", "saturating.fa.gz.11", "allkmers_07.fasta.gz.07",
                                  "rand200k.fa.bgz.11"])
def test_partition_mode_matches_reference_golden(env, case):
    from pykmer_b200 import fasta
    fname, kk = case.rsplit(".", 1)
    K = int(kk)
    gold = json.load(open(os.path.join(GOLD, "indexer", case + ".json")))
    stream, names, starts, lengths = fasta.read_fasta_stream(os.path.join(GOLD, "inputs", fname))
    wlog = 4 if K <= 7 else 12
    for pieces, flush in ((None, "smem"), ([7, 100, 101, 5000], "smem"), ([7, 100, 101, 5000], "l2"),
                          (None, "byte"), ([7, 100, 101, 5000], "byte")):
        pieces = [c for c in (pieces or []) if c < len(stream)] or None
        table, hist, st, flags = _index_stream(env["dev"], stream, K, starts=starts, mode=PART,
                                               window_log2=wlog, pieces=pieces, flush=flush)
        assert st["num_kmers"] == gold["num_kmers"] and hist == gold["hist"]
        assert hashlib.sha256(table.tobytes()).hexdigest() == gold["output_file_cheksum"]
        chrom = [[names[i], lengths[i]] for i in range(len(names)) if flags[i]]
        assert chrom == gold["chromosomes"]


@pytest.mark.parametrize("K", [17, 19, 31])
def test_partition_mode_wide_k_range(env, K):
    rng = np.random.default_rng(K + 500)
    parts = []
    for _ in range(600):
        parts.append(np.full(int(rng.integers(K - 3, 3 * K)), ord("A"), dtype=np.uint8))
        parts.append(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(rng.integers(1, 12)))])
    s = np.concatenate(parts)
    hi = 1 << 22
    want, num, _ = env["oracle"].index_stream(s, K, range_hi=hi)
    for host in (False, True):
        table, hist, st, _ = _index_stream(env["dev"], s, K, hi=hi, mode=PART, window_log2=12, host=host,
                                           pieces=[1000, 1001, 20_000])
        assert st["num_kmers"] == num and np.array_equal(table, want)


def _repeat_heavy_stream(rng, n_motifs, copies, spacer=True):
    """Many k-mers far beyond 255 occurrences: homopolymers, (AT)n, (AAT)n and random motifs
    repeated `copies` times -- what makes the 8-bit lanes of the byte windows carry."""
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    parts = [np.full(3000, ord("A"), dtype=np.uint8), np.frombuffer(b"N", dtype=np.uint8),
             np.tile(np.frombuffer(b"AT", dtype=np.uint8), 1200), np.frombuffer(b"N", dtype=np.uint8),
             np.tile(np.frombuffer(b"AAT", dtype=np.uint8), 900), np.frombuffer(b"N", dtype=np.uint8),
             np.full(700, ord("c"), dtype=np.uint8)]
    motifs = [acgt[rng.integers(0, 4, size=int(rng.integers(20, 60)))] for _ in range(n_motifs)]
    order = rng.integers(0, n_motifs, size=n_motifs * copies)
    for m in order:
        parts.append(motifs[m])
        if spacer and rng.random() < 0.3:
            parts.append(acgt[rng.integers(0, 4, size=int(rng.integers(1, 9)))])
    return np.concatenate(parts)


@pytest.mark.parametrize("K,wlog", [(7, 8), (9, 18), (11, 12), (13, 26), (17, 26)])
@pytest.mark.parametrize("ovf_log2", [None, 10, 3, 1])
@pytest.mark.parametrize("pieces", [None, [5000, 90_001]])
def test_byte_windows_settle_carries_exactly(env, K, wlog, ovf_log2, pieces):
    """Byte windows (PYKMER_B200_FLUSH=byte): 8-bit lanes that carry into their neighbours are
    corrected exactly -- through the carry table, and through the compare-and-swap recount when the
    table is too small (ovf_log2 1 and 3) -- also across several saturating flushes."""
    rng = np.random.default_rng(1000 + K)
    s = _repeat_heavy_stream(rng, 40, 400)
    hi = None if K <= 13 else 1 << 28
    want, num, _ = env["oracle"].index_stream(s, K, range_hi=hi) if hi else env["oracle"].index_stream(s, K)
    assert (want == 255).sum() > 20                    # the case really saturates many lanes
    table, hist, st, _ = _index_stream(env["dev"], s, K, hi=hi, mode=PART, window_log2=wlog, flush="byte",
                                       ovf_log2=ovf_log2, pieces=pieces, pool_log2=18 if pieces else None)
    assert st["num_kmers"] == num
    assert np.array_equal(table, want)
    oh, ost = env["oracle"].table_stats(want)
    assert hist == oh and all(st[k] == ost[k] for k in ("vals_sum", "vals_count", "vals_min", "vals_max"))
    ref, _, st2, _ = _index_stream(env["dev"], s, K, hi=hi, mode=PART, window_log2=min(wlog, 24), flush="l2")
    assert np.array_equal(ref, table) and st2 == st


@pytest.mark.parametrize("K,wlog", [(11, 16), (13, 20), (13, 24), (15, 24), (17, 26)])
@pytest.mark.parametrize("est", ["4:12:0", "2:12:0", "4:12:1", "4:12:2", "0:12:0"])
@pytest.mark.parametrize("pieces", [None, [300_001, 1_200_000]])
def test_estimated_pass1_and_its_exact_fallback(env, K, wlog, est, pieces):
    """PARTITION mode sizes its segments from a sampled pass 1 (one tile in 2^shift); pass 2 then
    does the bookkeeping.  Normal case, the two ways the estimate can fall short (over budget:
    test 1, a window overflowing its room: test 2 -- the exact kernels queued behind redo the feed)
    and the always-exact path all give the oracle's table, statistics, num_kmers and record flags."""
    rng = np.random.default_rng(77 + K)
    recs = [_random_stream(rng, int(n)) for n in (700_000, 3, 400_000, 900_000)]
    sep = np.frombuffer(b">", dtype=np.uint8)
    s = np.concatenate([np.concatenate([r, sep]) for r in recs])
    starts = np.cumsum([0] + [len(r) + 1 for r in recs[:-1]]).astype(np.uint64)
    hi = None if K <= 13 else 1 << 28
    want, num, flags_want = env["oracle"].index_stream(s, K, range_hi=hi, rec_starts=starts) if hi else \
        env["oracle"].index_stream(s, K, rec_starts=starts)
    table, hist, st, flags = _index_stream(env["dev"], s, K, hi=hi, starts=starts, mode=PART, window_log2=wlog,
                                           est=est, pieces=pieces)
    assert st["num_kmers"] == num
    assert np.array_equal(table, want)
    assert hist == env["oracle"].table_stats(want)[0]
    assert list(flags) == list(flags_want)


def test_partition_mode_feed_after_finalize_and_reset(env):
    import torch
    rng = np.random.default_rng(12)
    dev, oracle = env["dev"], env["oracle"]
    a, b = _random_stream(rng, 40_000), _random_stream(rng, 30_000)
    os.environ["PYKMER_B200_WINDOW_LOG2"] = "10"
    try:
        ix = dev.Indexer(9, mode=PART)
    finally:
        os.environ.pop("PYKMER_B200_WINDOW_LOG2", None)
    with ix:
        assert ix.mode() == (PART, 4 ** 9 >> 10)
        ix.feed_device(torch.from_numpy(a).cuda())
        h1, s1 = ix.finalize()
        w1, n1, _ = oracle.index_stream(a, 9)
        assert np.array_equal(ix.table_to_host().numpy(), w1) and s1["num_kmers"] == n1
        assert ix.finalize() == (h1, s1)                       # idempotent
        ix.feed_device(torch.from_numpy(b).cuda())             # keeps accumulating (saturating)
        with pytest.raises(RuntimeError):
            ix.table_to_host()                                 # k-mers still buffered
        h2, s2 = ix.finalize()
        w2, n2, _ = oracle.index_stream(np.concatenate([a, b]), 9)
        assert np.array_equal(ix.table_to_host().numpy(), w2) and s2["num_kmers"] == n2
        assert h2 == oracle.table_stats(w2)[0]
        ix.reset()
        ix.feed_device(torch.from_numpy(b).cuda())
        h3, s3 = ix.finalize()
        w3, n3, _ = oracle.index_stream(b, 9)
        assert np.array_equal(ix.table_to_host().numpy(), w3) and s3["num_kmers"] == n3


def test_auto_mode_picks_partition_for_k15(env):
    """AUTO: DIRECT for tiny (K <= 9) and very sparse (K >= 19) tables, PARTITION between --
    2^24-entry windows of 32-bit counters up to K=15, 2^26-entry byte windows at K=17."""
    with env["dev"].Indexer(15) as ix:
        assert ix.mode() == (PART, 64) and ix.window_log2() == 24
    with env["dev"].Indexer(11) as ix:
        assert ix.mode() == (PART, 1)
    with env["dev"].Indexer(9) as ix:
        assert ix.mode()[0] == 1 and ix.window_log2() == 0
    with env["dev"].Indexer(17, range_hi=1 << 30) as ix:
        assert ix.mode() == (PART, 16) and ix.window_log2() == 26
    with env["dev"].Indexer(19, range_hi=1 << 30) as ix:
        assert ix.mode()[0] == 1


def test_indexer_scaled_config2_both_modes(env):
    """1/16-scale config 2 stream, K=15, full table: DIRECT == PARTITION == oracle."""
    from pykmer_b200 import synth
    dev, oracle = env["dev"], env["oracle"]
    recs = synth.syn782m_records(scale=1 / 16)
    stream, starts, lengths, names = synth.records_to_stream(recs)
    want, num, _ = oracle.index_stream(stream, 15, method="mt", threads=oracle.max_threads())
    for mode in (1, 2):
        table, hist, st, flags = _index_stream(dev, stream, 15, starts=starts, mode=mode)
        assert st["num_kmers"] == num and flags.all()
        assert np.array_equal(table, want), f"mode {mode}"
        assert hist == oracle.table_stats(want, threads=oracle.max_threads())[0]


def test_full_size_config2_modes_agree(env):
    """BASELINE config 2 at full size (782.5 Mbp, K=15, 1 GiB table): the two independent
    counting schemes produce the same table; size-independent invariants hold."""
    import torch
    import bench
    dev = env["dev"]
    stream, starts, lengths = bench.load_stream(1.0, 0, 1)
    d = torch.from_numpy(stream).cuda()
    digests, stats = [], []
    for mode, flush in ((1, None), (2, "smem"), (2, "l2"), (2, "byte")):
        if flush:
            os.environ["PYKMER_B200_FLUSH"] = flush
        try:
            ix = dev.Indexer(15, mode=mode)
        finally:
            os.environ.pop("PYKMER_B200_FLUSH", None)
        with ix:
            ix.set_records(starts)
            ix.feed_device(d)
            hist, st = ix.finalize()
            assert ix.record_flags().all()
            p, n = ix.table_ptr()
            t = torch.empty(0, dtype=torch.uint8, device="cuda")
            host = ix.table_to_host().numpy()
            digests.append(hashlib.sha256(host.tobytes()).hexdigest())
            stats.append((hist, st))
            assert st["vals_sum"] <= st["num_kmers"] and sum(hist) == st["vals_count"]
            assert st["vals_sum"] == sum((i + 1) * h for i, h in enumerate(hist))
            assert st["vals_min"] == 0 and st["vals_max"] == 255
            assert st["num_kmers"] <= sum(max(0, l - 14) for l in lengths)
            # canonical k-mers only: an entry whose reverse complement is smaller must be 0
            idx = np.random.default_rng(0).integers(0, 4 ** 15, size=200_000)
            rc = np.zeros_like(idx)
            for q in range(15):
                rc |= (3 - ((idx >> (2 * q)) & 3)) << (2 * (14 - q))
            assert not host[idx[idx > rc]].any()
    assert len(set(digests)) == 1 and all(s == stats[0] for s in stats)


def test_full_size_k17_shard_schemes_agree(env):
    """The 782.5 Mbp stream at K=17 on the lowest quarter of the k-mer axis (a 4 GiB shard, ~45 % of
    the k-mers, the microsatellite k-mers with millions of occurrences among them): byte windows
    counted in place (the default), 32-bit windows and the DIRECT scan give the same table and the
    same statistics -- the carries of the 8-bit lanes and the incremental histograms hold at scale."""
    import torch
    import bench
    dev = env["dev"]
    stream, starts, lengths = bench.load_stream(1.0, 0, 1)
    d = torch.from_numpy(stream).cuda()
    hi = 1 << 32
    digests, stats = [], []
    for mode, flush in ((2, None), (2, "l2"), (1, None)):
        if flush:
            os.environ["PYKMER_B200_FLUSH"] = flush
        try:
            ix = dev.Indexer(17, range_hi=hi, mode=mode)
        finally:
            os.environ.pop("PYKMER_B200_FLUSH", None)
        with ix:
            ix.feed_device(d)
            hist, st = ix.finalize()
            host = ix.table_to_host().numpy()
            digests.append(hashlib.sha256(host.tobytes()).hexdigest())
            stats.append((hist, st))
            assert sum(hist) == st["vals_count"] and st["vals_max"] == 255 and hist[254] > 0
            assert st["vals_sum"] == sum((i + 1) * h for i, h in enumerate(hist))
            assert st["vals_sum"] == int(host.sum(dtype=np.uint64))
            assert st["vals_count"] == int(np.count_nonzero(host))
    assert len(set(digests)) == 1 and all(s == stats[0] for s in stats)


@pytest.mark.parametrize("mode,wlog,bases,packed,threads", [
    (1, None, 300_000, "1", None), (2, 10, 300_000, "1", "3"), (2, 24, 300_000, "1", None), (2, 14, 300_000, "1", "1"),
    (2, 16, 300_000, "0", None),                     # the plain copy of every window
    (2, 18, 40_000_000, "1", None),                  # every canonical 11-mer present: the dense low windows are sent as they are
    (2, 16, 2_500_000, "1", "2"),                    # about half full: packed and raw windows side by side
])
def test_finalize_to_host_streams_the_table(env, mode, wlog, bases, packed, threads):
    """pk_indexer_finalize_to_host: windows leave packed (bitmap + non-zero bytes, rebuilt by a team of host
    threads) or as they are -- dense windows, busy slots, PYKMER_B200_PACKED_D2H=0 -- and land the same bytes."""
    rng = np.random.default_rng(321)
    s = _random_stream(rng, bases) if bases < 40_000_000 else _random_stream(rng, bases, alphabet=b"ACGT")
    K = 11
    want, num, _ = env["oracle"].index_stream(s, K)
    dev = env["dev"]
    if wlog is not None:
        os.environ["PYKMER_B200_WINDOW_LOG2"] = str(wlog)
    os.environ["PYKMER_B200_PACKED_D2H"] = packed
    if threads:
        os.environ["PYKMER_B200_UNPACK_THREADS"] = threads
    try:
        ix = dev.Indexer(K, mode=mode)
        with ix:
            out = dev.pinned_empty(4 ** K + 64)[37:37 + 4 ** K] if threads == "2" else dev.pinned_empty(4 ** K)
            for rep in range(2):
                out.fill_(9)
                ix.reset()
                ix.feed_host(s)
                hist, st = ix.finalize(table_out=out)
                assert st["num_kmers"] == num and np.array_equal(out.numpy(), want)
                assert hist == env["oracle"].table_stats(want)[0]
                x = ix.transfer_stats()
                if mode == 2:
                    assert x["packed_windows"] + x["raw_windows"] == max(1, (4 ** K) >> wlog)
                    if packed == "0":
                        assert x["packed_windows"] == 0 and x["d2h_bytes"] >= 4 ** K
                    elif bases == 40_000_000:                # the low windows are full, the high ones hold no canonical k-mer
                        assert x["raw_windows"] >= 4
                    elif bases == 300_000:
                        assert x["packed_windows"] > 0 and (threads or x["d2h_bytes"] < 4 ** K // 2)
                    if threads and x["packed_windows"]:
                        assert x["unpack_threads"] == int(threads)
            hist2, st2 = ix.finalize(table_out=out)              # nothing pending: plain copy
            assert (hist2, st2) == (hist, st) and np.array_equal(out.numpy(), want)
    finally:
        for k in ("PYKMER_B200_WINDOW_LOG2", "PYKMER_B200_PACKED_D2H", "PYKMER_B200_UNPACK_THREADS"):
            os.environ.pop(k, None)


def test_packed_transfer_is_stricter_when_ranks_share_the_host(env):
    """With more than two ranks per host (LOCAL_WORLD_SIZE) only windows that pack to a quarter of their bytes
    leave packed; PYKMER_B200_PACKED_D2H=1 keeps the 5/8 rule.  Same bytes either way."""
    rng = np.random.default_rng(99)
    s = _random_stream(rng, 2_500_000)
    K = 11
    want, num, _ = env["oracle"].index_stream(s, K)
    dev = env["dev"]
    counts = {}
    for name, extra in (("loose", {"PYKMER_B200_PACKED_D2H": "1"}), ("strict", {})):
        os.environ.update({"PYKMER_B200_WINDOW_LOG2": "16", "LOCAL_WORLD_SIZE": "4", **extra})
        try:
            with dev.Indexer(K, mode=2) as ix:
                out = dev.pinned_empty(4 ** K)
                out.fill_(3)
                ix.feed_host(s)
                hist, st = ix.finalize(table_out=out)
                assert st["num_kmers"] == num and np.array_equal(out.numpy(), want)
                counts[name] = ix.transfer_stats()
        finally:
            for k in ("PYKMER_B200_WINDOW_LOG2", "LOCAL_WORLD_SIZE", "PYKMER_B200_PACKED_D2H"):
                os.environ.pop(k, None)
    # how many windows go packed also depends on how far the (small) team falls behind, so only the totals are fixed
    for c in counts.values():
        assert c["packed_windows"] + c["raw_windows"] == (4 ** K) >> 16 and c["d2h_bytes"] <= 4 ** K + 4 * 64 + 2056


@pytest.mark.parametrize("n,fill", [(1024, 0.0), (1024, 1.0), (1 << 16, 0.5), (1 << 22, 0.25), (3 << 20, 0.03), (1 << 24, 0.9)])
def test_table_pack_kernel_and_host_unpack(env, n, fill):
    """k_table_pack against the NumPy restatement (chunks may land in any order: compared through the
    literal inverse), and the device-packed slice through the host's pk_table_unpack."""
    dev, oracle = env["dev"], env["oracle"]
    rng = np.random.default_rng(n % 1000 + int(fill * 100))
    t = rng.integers(1, 256, n, dtype=np.uint8) if fill == 0.5 else np.minimum(rng.geometric(0.3, n), 255).astype(np.uint8)
    t[rng.random(n) >= fill] = 0
    bm, off, nz = dev.table_pack(t)
    o_bm, o_off, o_nz = oracle.pack_table(t)
    assert np.array_equal(bm, o_bm) and nz.size == o_nz.size
    assert np.array_equal(oracle.unpack_table(bm, off, nz, n), t)
    assert np.array_equal(dev.table_unpack(bm, off, nz, n), t)


@pytest.mark.parametrize("K,wlog,nranks", [(9, 10, 3), (11, 12, 2), (17, 12, 4)])
def test_sequence_sharded_scan_exchange_count(env, K, wlog, nranks):
    """The multi-GPU indexer path emulated rank after rank on one GPU: every 'rank' scans its
    slice of the stream (primed with the preceding bytes), the bucketed entries are routed to
    the owners of their windows exactly as dist.exchange_entries does (the all-to-all is
    emulated with tensor copies), each owner counts its windows; concatenated shards, num_kmers
    and record flags equal the oracle's single pass."""
    import torch
    from pykmer_b200 import dist as pdist, _native as nat
    dev, oracle = env["dev"], env["oracle"]
    rng = np.random.default_rng(K * 7 + nranks)
    s = _random_stream(rng, 120_000)
    if K > 13:                                                  # make a narrow k-mer range busy
        s[::7] = ord("A"); s[1::7] = ord("A"); s[2::7] = ord("A")
    starts = np.array([0, 30_000, 30_001, 90_000], dtype=np.uint64)
    hi_all = min(4 ** K, 1 << 22)
    want, num, oflags = oracle.index_stream(s, K, range_hi=hi_all, rec_starts=starts)
    os.environ["PYKMER_B200_WINDOW_LOG2"] = str(wlog)
    try:
        scanners, exports = [], []
        for r in range(nranks):
            a, b = pdist.slice_bounds(len(s), r, nranks)
            sc = dev.Indexer(K, range_hi=hi_all, mode=nat.PK_MODE_SCAN)
            sc.set_records(starts)
            halo = torch.from_numpy(s[max(0, a - 32):a].copy()).cuda() if a > 0 else None
            sc.prime(halo, a)
            piece = torch.from_numpy(s[a:b].copy()).cuda()
            sc.feed_device(piece[:len(piece) // 2 // 16 * 16])     # two feeds -> two segments
            sc.feed_device(piece[len(piece) // 2 // 16 * 16:])
            scanners.append(sc)
            exports.append(sc.export_segments())
            with pytest.raises(RuntimeError):
                sc.finalize()                                    # a scanner has no table
        nwin = scanners[0].mode()[1]
        nseg = max(e[2].shape[0] for e in exports)
        all_cnt = np.zeros((nranks, nseg, nwin), dtype=np.int64)
        for r, (_, off, cnt) in enumerate(exports):
            all_cnt[r, :cnt.shape[0]] = cnt
        owners = pdist.balanced_window_owners(all_cnt.sum(axis=(0, 1)), nranks, overhead=10)
        tables, total, flags = [], 0, np.zeros(len(starts), dtype=np.uint8)
        for d in range(nranks):
            send_d, recv_d, imp_off, imp_cnt, tot = pdist.plan_exchange(all_cnt, owners, d)
            buf = torch.empty(max(tot, 1), dtype=torch.int32, device="cuda")
            pos = 0
            for f in range(nseg):
                for src in range(nranks):
                    entries, off, cnt = exports[src]
                    send_s = pdist.plan_exchange(all_cnt, owners, src)[0]
                    if f >= cnt.shape[0]:
                        continue
                    start = int(off[f, 0]) + int(send_s[f, :d].sum())
                    n = int(send_s[f, d])
                    buf[pos:pos + n] = entries[start:start + n]
                    pos += n
            w0, w1 = owners[d]
            lo, hi = w0 << wlog, min(hi_all, w1 << wlog)
            ct = dev.Indexer(K, range_lo=lo, range_hi=hi, mode=nat.PK_MODE_PARTITION)
            ct.import_segments(buf, imp_off, imp_cnt)
            hist, st = ct.finalize()
            tables.append(ct.table_to_host().numpy().copy())
            ct.close()
        for sc in scanners:
            total += sc.scan_result()
            flags |= sc.record_flags()
            sc.close()
    finally:
        os.environ.pop("PYKMER_B200_WINDOW_LOG2", None)
    assert total == num and np.array_equal(flags, oflags)
    assert np.array_equal(np.concatenate(tables), want)


@pytest.mark.parametrize("K,wlog,nranks", [(9, 10, 3), (11, 12, 2), (17, 12, 4)])
def test_fused_exchange_routing_on_one_gpu(env, K, wlog, nranks):
    """The fused multi-GPU path (pass 2 stores straight into the window owners' buffers) with all
    'ranks' living on one GPU: the same kernels, routing tables and import tables as the
    NVLink run, peers mapped locally instead of through CUDA IPC."""
    import torch
    from pykmer_b200 import dist as pdist, _native as nat
    dev, oracle = env["dev"], env["oracle"]
    rng = np.random.default_rng(K * 11 + nranks)
    s = _random_stream(rng, 150_000)
    if K > 13:
        s[::7] = ord("A"); s[1::7] = ord("A"); s[2::7] = ord("A")
    starts = np.array([0, 30_000, 30_001, 90_000], dtype=np.uint64)
    hi_all = min(4 ** K, 1 << 22)
    want, num, oflags = oracle.index_stream(s, K, range_hi=hi_all, rec_starts=starts)
    os.environ["PYKMER_B200_WINDOW_LOG2"] = str(wlog)
    try:
        scanners, pieces = [], []
        for r in range(nranks):
            a, b = pdist.slice_bounds(len(s), r, nranks)
            sc = dev.Indexer(K, range_hi=hi_all, mode=nat.PK_MODE_SCAN)
            sc.set_records(starts)
            sc.prime(torch.from_numpy(s[max(0, a - 32):a].copy()).cuda() if a > 0 else None, a)
            piece = torch.from_numpy(s[a:b].copy()).cuda()
            sc.scan_pass1(piece)
            scanners.append(sc)
            pieces.append(piece)
        all_cnt = np.stack([sc.pass1_counts() for sc in scanners]).astype(np.int64)
        owners = pdist.balanced_window_owners(all_cnt.sum(axis=0), nranks, overhead=10)
        counters = []
        for d in range(nranks):
            w0, w1 = owners[d]
            counters.append(dev.Indexer(K, range_lo=w0 << wlog, range_hi=min(hi_all, w1 << wlog),
                                        mode=nat.PK_MODE_PARTITION))
        for r, sc in enumerate(scanners):
            for d in range(nranks):
                sc.open_peer_pool(d, local_owner=counters[d])
            owner_of, dest_off, _, _, _ = pdist.plan_fused(all_cnt, owners, r)
            sc.scan_pass2_remote(nranks, owner_of, dest_off)
        tables = []
        for d, ct in enumerate(counters):
            _, _, imp_off, imp_cnt, landed = pdist.plan_fused(all_cnt, owners, d)
            assert landed == int(all_cnt[:, owners[d][0]:owners[d][1]].sum())
            ct.import_own_pool(imp_off, imp_cnt)
            ct.finalize()
            tables.append(ct.table_to_host().numpy().copy())
        total = sum(sc.scan_result() for sc in scanners)
        flags = np.zeros(len(starts), dtype=np.uint8)
        for sc in scanners:
            flags |= sc.record_flags()
        for h in scanners + counters:
            h.close()
    finally:
        os.environ.pop("PYKMER_B200_WINDOW_LOG2", None)
    assert total == num and np.array_equal(flags, oflags)
    assert np.array_equal(np.concatenate(tables), want)


@pytest.mark.parametrize("K,wlog,nranks", [(9, 10, 3), (11, 12, 2), (15, 24, 4), (17, 12, 4)])
def test_routed_exchange_on_one_gpu(env, K, wlog, nranks):
    """The routed multi-GPU path -- fixed, capacity-checked regions per (source rank, window) sized
    from a planning scan, fill counts published into the owners' buffers, no host round trip -- with
    all 'ranks' living on one GPU.  Also: regions that are too small raise the overflow status
    instead of writing past their room."""
    import torch
    from pykmer_b200 import dist as pdist, _native as nat
    dev, oracle = env["dev"], env["oracle"]
    rng = np.random.default_rng(K * 13 + nranks)
    s = _random_stream(rng, 150_000)
    if K > 13:
        s[::7] = ord("A"); s[1::7] = ord("A"); s[2::7] = ord("A")
    starts = np.array([0, 30_000, 30_001, 90_000], dtype=np.uint64)
    hi_all = min(4 ** K, 1 << (22 if wlog < 24 else 26))
    want, num, oflags = oracle.index_stream(s, K, range_hi=hi_all, rec_starts=starts)
    os.environ["PYKMER_B200_WINDOW_LOG2"] = str(wlog)
    try:
        scanners, pieces, halos, offs = [], [], [], []
        for r in range(nranks):
            a, b = pdist.slice_bounds(len(s), r, nranks)
            sc = dev.Indexer(K, range_hi=hi_all, mode=nat.PK_MODE_SCAN)
            sc.set_records(starts)
            halos.append(torch.from_numpy(s[max(0, a - 32):a].copy()).cuda() if a > 0 else None)
            offs.append(a)
            pieces.append(torch.from_numpy(s[a:b].copy()).cuda())
            sc.prime(halos[r], a)
            sc.scan_pass1(pieces[r])                           # the planning scan: exact counts per window
            scanners.append(sc)
        all_cnt = np.stack([sc.pass1_counts() for sc in scanners]).astype(np.int64)
        nwin = all_cnt.shape[1]
        owners = pdist.balanced_window_owners(all_cnt.sum(axis=0), nranks, overhead=10)
        counters = []
        for d in range(nranks):
            w0, w1 = owners[d]
            counters.append(dev.Indexer(K, range_lo=w0 << wlog, range_hi=min(hi_all, w1 << wlog),
                                        mode=nat.PK_MODE_PARTITION))
        pub = [ct.pub_base() for ct in counters]
        status = torch.zeros((nranks, 4), dtype=torch.int32, device="cuda")
        for attempt, (counts, slack) in enumerate(((all_cnt // 3, 0), (all_cnt, 64))):
            for r, sc in enumerate(scanners):
                for d in range(nranks):
                    sc.open_peer_pool(d, local_owner=counters[d])
                owner_of, dest_off, cap, _ = pdist.plan_routed(counts, owners, r, pub, slack=slack)
                sc.set_route(nranks, r, owner_of, dest_off, cap, pub)
                sc.reset()
                sc.prime(halos[r], offs[r])
                sc.scan_routed(pieces[r], status[r])
            for d, ct in enumerate(counters):
                _, _, _, imp_off = pdist.plan_routed(counts, owners, d, pub, slack=slack)
                ct.reset()
                ct.set_import_layout(imp_off, owners[d][0], nwin)
                ct.import_published()
            flagged = status[:, 0].cpu().numpy()
            if attempt == 0:
                assert flagged.any(), "regions a third of the size they need must overflow"
                for ct in counters:
                    ct.finalize()                              # harmless: the step is discarded
                continue
            assert not flagged.any()
            # the status words also carry every scanner's num_kmers (no separate read-back in a step)
            assert pdist.read_routed_status(status)[2] == [sc.scan_result() for sc in scanners]
            tables = []
            for ct in counters:
                ct.finalize()
                tables.append(ct.table_to_host().numpy().copy())
        total = sum(sc.scan_result() for sc in scanners)
        flags = np.zeros(len(starts), dtype=np.uint8)
        for sc in scanners:
            flags |= sc.record_flags()
        for h in scanners + counters:
            h.close()
    finally:
        os.environ.pop("PYKMER_B200_WINDOW_LOG2", None)
    assert total == num and np.array_equal(flags, oflags)
    assert np.array_equal(np.concatenate(tables), want)


# ---------------------------------------------------------------------------------------------
# The CLIs as multi-rank jobs with the real library: two ranks (gloo rendezvous, both on cuda:0)
# each own half of the canonical k-mer axis.  The host protocol alone is covered on the CPU in
# tests/test_dist_gloo.py; here the range-restricted handles and the sliced merge do the work.

@pytest.mark.parametrize("case,nranks,hooks,scheme", [
    ("rand200k.fa.bgz.13", 2, {}, "sequence slices"),                      # 4 windows of 2^24 entries
    ("rand200k.fa.bgz.13", 3, {"PYKMER_B200_WINDOW_LOG2": "14"}, "sequence slices"),
    ("saturating.fa.gz.11", 2, {"PYKMER_B200_WINDOW_LOG2": "12"}, "sequence slices"),
    ("tiny_mixed.fa.09", 2, {"PYKMER_B200_WINDOW_LOG2": "8"}, "sequence slices"),
    ("rand200k.fa.bgz.13", 2, {"PYKMER_B200_SHARD": "kmer"}, "k-mer ranges"),
    ("saturating.fa.gz.11", 2, {}, "k-mer ranges"),                        # one window: fewer than ranks
    ("tiny_mixed.fa.03", 2, {}, "k-mer ranges"),
])
def test_indexer_cli_two_ranks_on_one_gpu(env, case, nranks, hooks, scheme, tmp_path):
    """torchrun indexer.py as a multi-rank job with the real library (all ranks on cuda:0, gloo
    rendezvous): K <= 17 with enough table windows runs sequence slices + the fused exchange through
    CUDA-IPC peer mappings, everything else k-mer ranges with a replicated scan; outputs equal the
    reference's own files either way."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import multirank
    fname, kk = case.rsplit(".", 1)
    K = int(kk)
    gold = json.load(open(os.path.join(GOLD, "indexer", case + ".json")))
    src = str(tmp_path / fname)
    shutil.copy(os.path.join(GOLD, "inputs", fname), src)
    os.environ.update(hooks)
    try:
        res = multirank.run_cli(tmp_path, "indexer", [src, "sample", K], nranks=nranks, fake=False)
    finally:
        for k in hooks:
            os.environ.pop(k, None)
    assert all(rc == 0 for rc, _ in res), "\n".join(out for _, out in res)
    assert scheme in res[0][1], res[0][1][-600:]
    kin = f"{src}.{K:02d}.kin"
    meta = json.load(open(kin + ".json"))
    for k, v in gold.items():
        if k in ("all_keys", "project_name"):
            continue
        assert meta[k] == v, k
    assert gen_sha(kin) == gold["output_file_cheksum"] == meta["output_file_cheksum"]


def test_merger_cli_two_ranks_on_one_gpu(env, tmp_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import multirank
    kins, _ = multirank.golden_merger_inputs(tmp_path, bgzf_packed=True)
    lo, hi = 1, 50
    gold = np.load(os.path.join(GOLD, "merger", f"matrix_K07_{lo:03d}-{hi:03d}.npz"))["matrix"]
    proj = str(tmp_path / "proj")
    res = multirank.run_cli(tmp_path, "merger", [proj] + kins + [f"--min-count={lo}", f"--max-count={hi}"],
                            nranks=2, fake=False)
    assert all(rc == 0 for rc, _ in res), "\n".join(out for _, out in res)
    m = np.load(f"{proj}.{lo:03d}-{hi:03d}.kma")["matrix"]
    off = ~np.eye(m.shape[0], dtype=bool)
    assert m.dtype == np.uint64 and np.array_equal(m[off], gold[off])

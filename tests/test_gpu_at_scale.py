"""GPU parity at BASELINE.json's FULL sizes.

The CUDA path against (1) the CPU oracle run live on the GPU box on the same 782,520,033 bp
stream and (2) the digests the oracle produced when this suite was written
(tests/golden/at_scale.json, made by oracle/make_golden_at_scale.py), so the two cannot drift
apart silently:

  config 2   K=15, the whole 1 GiB table                        (indexer.py:341-342,239,262)
  config 4   K=17, the whole 16 GiB table, quarter by quarter   (byte windows counted in place)
  config 5   K=19, two disjoint 2^28-entry ranges of the 256 GiB table (DIRECT scan)
  config 3/4 merger K=15, N=255 and N=50: the default FP4 tensor-core Gram against the integer
             tensor-core kernel, against AND + popcount, and against the oracle's literal pair sums
             (tools.py:473-482) on whole 1 GiB tables
  FP4 exactness: partial sums that walk through ODD integers between 2^23 and 2^24

All integer work: bit-exact or failed.
"""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "at_scale.json")))


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    from oracle import oracle
    from pykmer_b200 import device, _native
    import bench
    stream, starts, lengths = bench.load_stream(1.0, 0, 1)
    assert hashlib.sha256(memoryview(stream)).hexdigest() == GOLD["stream"]["sha256"], \
        "the synthetic stream differs from the one the golden digests were made from"
    return {"torch": torch, "oracle": oracle, "dev": device, "nat": _native, "stream": stream,
            "starts": starts, "lengths": lengths, "d_stream": torch.from_numpy(stream).cuda(),
            "threads": oracle.max_threads()}


def _sha(a) -> str:
    return hashlib.sha256(memoryview(np.ascontiguousarray(a))).hexdigest()


def _stats_equal(hist, st, gold) -> None:
    assert hist == gold["hist"]
    for key in ("num_kmers", "vals_sum", "vals_count", "vals_min", "vals_max"):
        assert st[key] == gold[key], key


def test_config2_full_size_vs_oracle_and_golden(env):
    """BASELINE config 2 as the CLI runs it (PK_MODE_AUTO): every byte of the 1 GiB table, hist,
    vals_*, num_kmers and the record flags equal the oracle's, live and as committed."""
    dev, oracle, gold = env["dev"], env["oracle"], GOLD["k15"]
    with dev.Indexer(15) as ix:
        ix.set_records(env["starts"])
        ix.feed_device(env["d_stream"])
        hist, st = ix.finalize()
        flags = ix.record_flags()
        table = ix.table_to_host().numpy()
    _stats_equal(hist, st, gold)
    assert flags.tolist() == gold["record_flags"]
    assert _sha(table) == gold["sha256"]
    want, num, _ = oracle.index_stream(env["stream"], 15, method="mt", threads=env["threads"])
    assert num == st["num_kmers"] and np.array_equal(table, want)
    o_hist, o_st = oracle.table_stats(want, threads=env["threads"])
    assert hist == o_hist and all(st[k] == o_st[k] for k in ("vals_sum", "vals_count", "vals_min", "vals_max"))


def test_config2_full_size_through_host_buffers(env):
    """The e2e route of bench.py (pinned stream in, table windows streamed out while the rest is
    counted) lands the same bytes."""
    dev, gold = env["dev"], GOLD["k15"]
    h_stream = dev.pinned_empty(env["stream"].size)
    h_stream.numpy()[:] = env["stream"]
    h_table = dev.pinned_empty(4 ** 15)
    with dev.Indexer(15) as ix:
        ix.feed_host(h_stream)
        hist, st = ix.finalize(table_out=h_table)
    _stats_equal(hist, st, gold)
    assert _sha(h_table.numpy()) == gold["sha256"]


def test_config4_k17_full_table_vs_oracle_and_golden(env):
    """BASELINE config 4's indexer half: the whole 16 GiB table of K=17 in one pass (byte windows
    counted in place, histogram kept as transitions).  Every quarter of the table against the
    committed oracle digest; the lowest quarter (45 % of the k-mers, the microsatellite k-mers with
    millions of occurrences among them) also against the oracle run live."""
    dev, oracle = env["dev"], env["oracle"]
    with dev.Indexer(17) as ix:
        ix.set_records(env["starts"])
        ix.feed_device(env["d_stream"])
        hist, st = ix.finalize()
        assert ix.record_flags().all()
        quarter = dev.pinned_empty(1 << 32)
        for q, gold in enumerate(GOLD["k17"]):
            lo, hi = gold["range"]
            ix.table_to_host(dst=quarter, offset=lo, nbytes=hi - lo)
            host = quarter.numpy()
            assert _sha(host) == gold["sha256"], f"quarter {q}"
            if q == 0:
                want, num, _ = oracle.index_stream(env["stream"], 17, range_lo=lo, range_hi=hi, method="mt",
                                                   threads=env["threads"])
                assert num == gold["num_kmers"] and np.array_equal(host, want)
                del want
    # the whole-table statistics are the sums of the quarters'
    assert hist == [sum(g["hist"][i] for g in GOLD["k17"]) for i in range(255)]
    for key in ("num_kmers", "vals_sum", "vals_count"):
        assert st[key] == sum(g[key] for g in GOLD["k17"]), key
    assert st["vals_max"] == max(g["vals_max"] for g in GOLD["k17"]) and st["vals_min"] == 0


@pytest.mark.parametrize("which", [0, 1])
def test_config5_k19_ranges_vs_oracle_and_golden(env, which):
    """BASELINE config 5: K=19 shards of the k-mer axis (DIRECT scan of a very sparse table) on the
    full stream: one range in the crowded low end of the canonical axis, one in the middle."""
    dev, oracle, gold = env["dev"], env["oracle"], GOLD["k19"][which]
    lo, hi = gold["range"]
    with dev.Indexer(19, range_lo=lo, range_hi=hi) as ix:
        assert ix.mode()[0] == env["nat"].PK_MODE_DIRECT
        ix.feed_device(env["d_stream"])
        hist, st = ix.finalize()
        table = ix.table_to_host().numpy()
    _stats_equal(hist, st, gold)
    assert _sha(table) == gold["sha256"]
    want, num, _ = oracle.index_stream(env["stream"], 19, range_lo=lo, range_hi=hi, method="mt",
                                       threads=env["threads"])
    assert num == st["num_kmers"] and np.array_equal(table, want)


# ------------------------------------------------------------------------------------------ merger

def _device_masks(dev, torch, N, K, max_count, tiled_too=True):
    """N synthetic K-mer tables (pykmer_b200/synth.py, generated on the device) -> row-major masks
    and, optionally, tiled masks of the same tables."""
    T = 4 ** K
    words = T // 32
    raw = torch.empty(T, dtype=torch.uint8, device="cuda")
    rows = torch.zeros((N, words), dtype=torch.int32, device="cuda")
    tiled = dev.tiled_masks(words, N) if tiled_too else None
    for s in range(N):
        dev.synth_table(s, 0, T, out=raw)
        dev.threshold_pack(raw, 1, max_count, out=rows[s])
        if tiled_too:
            dev.threshold_pack_tiled(raw, 1, max_count, tiled, s, N)
    return rows, tiled, words, raw


@pytest.mark.parametrize("N,max_count", [(255, 255), (50, 50)])
def test_merger_baseline_size_f4_vs_i8_vs_popc_vs_oracle(env, N, max_count):
    """BASELINE configs 3 and 4 (merger half) at full size, K=15: the default path (tiled masks, FP4
    tensor-core Gram, FP32 accumulators that see 7.25 M k-mers per CTA at ~16 % density) equals the
    integer tensor-core kernel on every cell, AND + popcount on a 50-sample subset, and the oracle's
    literal pair sums on two pairs of whole tables."""
    torch, dev, oracle = env["torch"], env["dev"], env["oracle"]
    K = 15
    rows, tiled, words, raw = _device_masks(dev, torch, N, K, max_count)
    G_f4 = dev.gram_tiled(tiled, N, words).cpu().numpy()
    del tiled
    os.environ["PYKMER_B200_GRAM"] = "i8"
    try:
        G_i8 = dev.gram(rows, words=words).cpu().numpy()
        os.environ["PYKMER_B200_GRAM"] = "popc"
        n_sub = min(N, 50)
        G_popc = dev.gram(rows[:n_sub].contiguous(), words=words).cpu().numpy()
    finally:
        os.environ.pop("PYKMER_B200_GRAM", None)
    assert np.array_equal(G_f4, G_i8), "FP4 Gram differs from the integer tensor-core Gram"
    assert np.array_equal(G_f4[:n_sub, :n_sub], G_popc), "FP4 Gram differs from AND + popcount"
    assert (np.diag(G_f4) > 100_000_000).all() and G_f4.max() < 2 ** 30
    # the oracle's pair loop on whole tables (tools.py:473-482)
    for k, l in ((0, 1), (N - 2, N - 1)):
        t_k = dev.synth_table(k, 0, 4 ** K, out=raw).cpu().numpy()
        t_l = dev.synth_table(l, 0, 4 ** K).cpu().numpy()
        tot_k, tot_l, shared = oracle.pair_counts(t_k, t_l, 1, max_count)
        assert (int(G_f4[k, k]), int(G_f4[l, l]), int(G_f4[k, l])) == (tot_k, tot_l, shared)


def _adversarial_period(W):
    """Three rows of one period of W words (W * 32 k-mers < 2^24):
    A  2^23 ones, then exactly one set bit per 64-bit step;  B  all ones;
    C  2^23 ones, then one set bit (the same one as A's) every third step."""
    H = 1 << 18
    assert W > H and W % 2 == 0
    a = np.zeros(W, dtype=np.uint32)
    a[:H] = 0xFFFFFFFF
    steps = (W - H) // 2
    j = np.arange(steps, dtype=np.uint32)
    a[H::2] = np.uint32(1) << (j % np.uint32(32))
    c = a.copy()
    keep = (j % 3) == 0
    c[H::2] = np.where(keep, a[H::2], np.uint32(0))
    b = np.full(W, 0xFFFFFFFF, dtype=np.uint32)
    return a, b, c


def _popc(x: np.ndarray) -> int:
    return int(np.unpackbits(x.view(np.uint8)).sum(dtype=np.int64))


@pytest.mark.parametrize("N", [3, 100, 130])
def test_gram_f4_odd_partial_sums_up_to_2_24(env, N):
    """FP32 accumulation of 0/1 products is exact only if the tensor core keeps all 24 significand
    bits.  Every accumulator gets a slab of 2^19 - 1024 words (one per CTA; two per CTA, one for each
    half slab, at <= 64 samples): the first 2^23 k-mers are ones on every row,
    then row A contributes exactly ONE product per K=64 instruction -- the accumulators of (A, B),
    (A, A) walk through every integer, odd ones included, from 2^23 to 2^24 - 40960, those of (A, C),
    (C, C) through every third.  A lost low bit anywhere shows in G.  Compared with the closed form
    and with the integer tensor-core kernel."""
    torch, dev = env["torch"], env["dev"]
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    W = (1 << 19) - 1024
    a, b, c = _adversarial_period(W)
    periods = sms * (2 if N <= 64 else 1)
    words = periods * W
    kinds = ["B"] * N
    kinds[0], kinds[N - 1] = "A", "C"
    if N > 3:
        kinds[N // 2] = "A"                              # another A: (A, A) off the diagonal, both halves at N=130
    pat = {"A": torch.from_numpy(a.view(np.int32)).cuda(), "B": torch.from_numpy(b.view(np.int32)).cuda(),
           "C": torch.from_numpy(c.view(np.int32)).cuda()}
    rows = torch.empty((N, words), dtype=torch.int32, device="cuda")
    for r, kind in enumerate(kinds):
        rows[r].view(periods, W).copy_(pat[kind].unsqueeze(0).expand(periods, W))
    pc = {(x, y): periods * _popc({"A": a, "B": b, "C": c}[x] & {"A": a, "B": b, "C": c}[y])
          for x in "ABC" for y in "ABC"}
    want = np.array([[pc[(kinds[i], kinds[j])] for j in range(N)] for i in range(N)], dtype=np.int64)
    os.environ["PYKMER_B200_GRAM"] = "i8"
    try:
        G_i8 = dev.gram(rows, words=words).cpu().numpy()
    finally:
        os.environ.pop("PYKMER_B200_GRAM", None)
    assert np.array_equal(G_i8, want)
    tiled = rows.view(N, words // 32, 32).permute(1, 0, 2).contiguous().view(-1)
    del rows
    G = dev.gram_tiled(tiled, N, words).cpu().numpy()
    assert np.array_equal(G, want), "FP4 accumulators lost a bit below 2^24"

"""world_size-2 gloo test (CPU): the host-side logic of the k-mer-axis sharding -- shard
ranges tile the axis, per-shard statistics / record flags / partial Gram matrices reduce to
those of the whole job.  The per-shard numbers come from the CPU oracle (this is a test of
the reduction logic, not of the kernels)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %(root)r)
    import numpy as np, torch, torch.distributed as dist
    from oracle import oracle
    from pykmer_b200 import dist as pdist
    dist.init_process_group("gloo")
    rank, world = pdist.world()
    K = 9
    T = 4 ** K
    rng = np.random.default_rng(11)
    seq = np.frombuffer(b"ACGTNacgt>", dtype=np.uint8)[rng.integers(0, 10, size=60_000)].copy()
    seq[5000:6000] = ord("A")
    starts = np.array([0, 10_000, 10_001, 40_000], dtype=np.uint64)
    lo, hi = pdist.shard_range(T, rank, world, align=64)
    table, num, flags = oracle.index_stream(seq, K, range_lo=lo, range_hi=hi, rec_starts=starts)
    hist, st = oracle.table_stats(table)
    st = {"num_kmers": num, "vals_sum": st["vals_sum"], "vals_count": st["vals_count"],
          "vals_min": st["vals_min"], "vals_max": st["vals_max"]}
    hist_all, st_all = pdist.reduce_index_stats(hist, st)
    flags_all = pdist.reduce_flags(flags)
    # merger: each rank contracts its slice of the k-mer axis
    tables = np.stack([np.random.default_rng(100 + s).integers(0, 4, size=T, dtype=np.uint8) for s in range(5)])
    bits = np.stack([oracle.threshold_pack(t[lo:hi], 1, 2) for t in tables])
    G = torch.from_numpy(oracle.gram_from_bits(bits))
    pdist.reduce_gram(G)
    # stream broadcast
    chunk = torch.from_numpy(seq.copy()) if rank == 0 else None
    got = pdist.broadcast_stream(chunk, seq.size)
    out = {"rank": rank, "lo": lo, "hi": hi, "hist": hist_all, "st": st_all, "flags": flags_all.tolist(),
           "G": G.tolist(), "bcast_ok": bool(np.array_equal(got.numpy(), seq))}
    json.dump(out, open(os.path.join(%(out)r, f"rank{rank}.json"), "w"))
    dist.destroy_process_group()
""")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_ranges_tile_the_axis():
    from pykmer_b200 import dist as pdist
    for total in (4 ** 5, 4 ** 9, 4 ** 15, 4 ** 19):
        for n in (1, 2, 3, 4, 8):
            cuts = [pdist.shard_range(total, r, n) for r in range(n)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
            assert all(lo % 4096 == 0 for lo, _ in cuts)


def test_two_rank_reductions_equal_the_whole_job(tmp_path):
    import json
    from oracle import oracle
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": str(tmp_path)})
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out
    res = [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]
    # the whole job on one "rank"
    K, T = 9, 4 ** 9
    rng = np.random.default_rng(11)
    seq = np.frombuffer(b"ACGTNacgt>", dtype=np.uint8)[rng.integers(0, 10, size=60_000)].copy()
    seq[5000:6000] = ord("A")
    starts = np.array([0, 10_000, 10_001, 40_000], dtype=np.uint64)
    table, num, flags = oracle.index_stream(seq, K, rec_starts=starts)
    hist, st = oracle.table_stats(table)
    tables = np.stack([np.random.default_rng(100 + s).integers(0, 4, size=T, dtype=np.uint8) for s in range(5)])
    G = oracle.gram_from_bits(np.stack([oracle.threshold_pack(t, 1, 2) for t in tables]))
    assert res[0]["hi"] == res[1]["lo"] and res[0]["lo"] == 0 and res[1]["hi"] == T
    for r in res:
        assert r["bcast_ok"]
        assert r["hist"] == hist
        assert r["st"] == {"num_kmers": num, "vals_sum": st["vals_sum"], "vals_count": st["vals_count"],
                           "vals_min": st["vals_min"], "vals_max": st["vals_max"]}
        assert r["flags"] == flags.tolist()
        assert np.array_equal(np.array(r["G"]), G)


def test_exchange_plan_routes_every_entry_to_its_window_owner():
    """plan_exchange (pure host logic of the sequence-sharded indexer): emulate the all-to-all
    in NumPy and check that every rank ends up with exactly the entries of its windows, and
    that the imported segment tables index them."""
    from pykmer_b200 import dist as pdist
    rng = np.random.default_rng(3)
    nranks, nseg, nwin = 4, 2, 11
    all_cnt = rng.integers(0, 50, size=(nranks, nseg, nwin)).astype(np.int64)
    all_cnt[2, 1] = 0                                           # a rank with an empty segment
    per_window = all_cnt.sum(axis=(0, 1))
    for owners in (pdist.window_owner_ranges(nwin, nranks),
                   pdist.balanced_window_owners(per_window, nranks, overhead=5)):
        assert owners[0][0] == 0 and owners[-1][1] == nwin
        assert all(a[1] == b[0] and a[0] < a[1] for a, b in zip(owners[:-1], owners[1:]))
        # each rank's pool: per segment, windows in order; an entry is tagged (src, seg, window, i)
        pools = {}
        for s in range(nranks):
            for f in range(nseg):
                pools[s, f] = [(s, f, w, i) for w in range(nwin) for i in range(all_cnt[s, f, w])]
        plans = [pdist.plan_exchange(all_cnt, owners, r) for r in range(nranks)]
        for d in range(nranks):
            send_d, recv_d, imp_off, imp_cnt, total = plans[d]
            buf = []
            for f in range(nseg):
                for s in range(nranks):
                    send_s = plans[s][0]
                    start = int(send_s[f, :d].sum())
                    chunk = pools[s, f][start:start + int(send_s[f, d])]
                    assert len(chunk) == recv_d[f, s]
                    buf += chunk
            assert len(buf) == total
            w0, w1 = owners[d]
            for seg in range(nseg * nranks):
                f, s = divmod(seg, nranks)
                for wl in range(w1 - w0):
                    got = buf[int(imp_off[seg, wl]):int(imp_off[seg, wl]) + int(imp_cnt[seg, wl])]
                    assert got == [(s, f, w0 + wl, i) for i in range(all_cnt[s, f, w0 + wl])]
    bal = pdist.balanced_window_owners(np.array([100, 1, 1, 1, 1, 1, 1, 1]), 2, overhead=0)
    assert bal == [(0, 1), (1, 8)]


def test_balanced_kmer_ranges_tile_the_axis_and_even_out_cost():
    """Shards of the DIRECT (K >= 19) multi-GPU path: contiguous, window aligned, they tile
    [0, 4^K), and with k-mers crowding the low windows no shard costs much more than the mean."""
    from pykmer_b200 import dist as pdist
    K, wl = 19, 26
    T = 4 ** K
    nwin = T >> wl
    x = (np.arange(nwin) + 0.5) / nwin
    per_window = (780e6 * 2 * (1 - x) / nwin).astype(np.int64)         # density of min(fwd, rc)
    for n in (2, 4, 8):
        ranges = pdist.balanced_kmer_ranges(per_window, n, wl, T)
        assert ranges[0][0] == 0 and ranges[-1][1] == T
        assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:]))
        assert all(lo % (1 << wl) == 0 and hi > lo for lo, hi in ranges)
        cost = [per_window[lo >> wl:hi >> wl].sum() + 230_000 * ((hi - lo) >> wl) for lo, hi in ranges]
        assert max(cost) < 1.05 * (sum(cost) / n)
        equal = [per_window[nwin * r // n:nwin * (r + 1) // n].sum() for r in range(n)]
        assert max(equal) > 1.3 * (sum(equal) / n)                       # what balancing avoids


# ---------------------------------------------------------------------------------------------
# The drop-in CLIs as world_size-N jobs (torchrun indexer.py / merger.py): host logic only -- the
# device layer is tests/fake_device.py (CPU oracle behind the same names), the collectives are
# gloo.  Outputs are compared with what the reference's own indexer.py / merger.py wrote
# (tests/golden).  The same jobs run with the real library in tests/test_gpu_parity.py.

import hashlib
import json
import shutil

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import multirank  # noqa: E402

GOLD = multirank.GOLD


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("case,nranks", [("rand200k.fa.bgz.11", 2), ("tiny_mixed.fa.07", 2),
                                         ("tiny_mixed.fa.03", 2), ("saturating.fa.gz.07", 3)])
def test_indexer_cli_as_a_multi_rank_job_writes_reference_files(tmp_path, case, nranks):
    fname, kk = case.rsplit(".", 1)
    K = int(kk)
    gold = json.load(open(os.path.join(GOLD, "indexer", case + ".json")))
    src = str(tmp_path / fname)
    shutil.copy(os.path.join(GOLD, "inputs", fname), src)
    res = multirank.run_cli(tmp_path, "indexer", [src, "sample", K], nranks=nranks)
    assert all(rc == 0 for rc, _ in res), "\n".join(out for _, out in res)
    kin = f"{src}.{K:02d}.kin"
    assert os.path.exists(kin) and not os.path.exists(kin + ".tmp")
    assert os.path.getsize(kin) == 4 ** K
    meta = json.load(open(kin + ".json"))
    assert sorted(meta.keys()) == gold["all_keys"]
    for k, v in gold.items():
        if k == "all_keys":
            continue
        got = os.path.basename(meta[k]) if k == "project_name" else meta[k]
        assert got == v, k
    assert _sha(kin) == gold["output_file_cheksum"] == meta["output_file_cheksum"]


def test_indexer_multi_rank_job_stops_on_every_rank_when_the_reader_fails(tmp_path):
    bad = str(tmp_path / "bad.fa")
    open(bad, "wb").write(b">a\nACGTACGTAC\xc3\xa9GTACGT\n")           # non-ASCII inside sequence
    res = multirank.run_cli(tmp_path, "indexer", [bad, "s", 5], nranks=2, timeout=120)
    assert all(rc not in (0, None) for rc, _ in res), res
    assert "timed out" not in "".join(out for _, out in res)
    assert "ValueError" in res[0][1] and "rank 0 failed" in res[1][1]
    empty = str(tmp_path / "empty.fa")                                   # no k-mer at all: tools.py:367-368
    open(empty, "wb").write(b">a\nNNNNNNNN\n")
    res = multirank.run_cli(tmp_path, "indexer", [empty, "s", 5], nranks=2, timeout=120)
    assert all(rc not in (0, None) for rc, _ in res) and "AssertionError" in res[0][1]


@pytest.mark.parametrize("nranks,bgzf_packed", [(2, False), (2, True), (3, True)])
def test_merger_cli_as_a_multi_rank_job_writes_reference_files(tmp_path, nranks, bgzf_packed):
    kins, _ = multirank.golden_merger_inputs(tmp_path, bgzf_packed=bgzf_packed)
    lo, hi = 2, 10
    gold = np.load(os.path.join(GOLD, "merger", f"matrix_K07_{lo:03d}-{hi:03d}.npz"))["matrix"]
    gmeta = json.load(open(os.path.join(GOLD, "merger", f"matrix_K07_{lo:03d}-{hi:03d}.json")))
    proj = str(tmp_path / "proj")
    argv = [proj] + list(reversed(kins)) + [f"--min-count={lo}", f"--max-count={hi}"]
    res = multirank.run_cli(tmp_path, "merger", argv, nranks=nranks)
    assert all(rc == 0 for rc, _ in res), "\n".join(out for _, out in res)
    kma = f"{proj}.{lo:03d}-{hi:03d}.kma"
    m = np.load(kma)["matrix"]
    off = ~np.eye(m.shape[0], dtype=bool)
    assert m.dtype == np.uint64 and m.shape == gold.shape and np.array_equal(m[off], gold[off])
    desc = json.load(open(kma + ".json"))
    assert sorted(desc.keys()) == gmeta["top_keys"]
    assert [os.path.basename(d["index_file"]) for d in desc["data"]] == gmeta["order"]
    assert res[0][1].count("matrix Total") == m.shape[0] * (m.shape[0] - 1) // 2
    assert "matrix Total" not in res[1][1]                      # one rank reports and writes
    res = multirank.run_cli(tmp_path, "merger", argv, nranks=nranks)       # refuses to overwrite (merger.py:99)
    assert all(rc not in (0, None) for rc, _ in res) and "AssertionError" in res[0][1]


def test_analytic_kmer_ranges_tile_the_axis_and_balance_the_model():
    """dist.analytic_kmer_ranges (the multi-GPU CLI at K >= 19): aligned cuts that tile the axis, every
    shard non-empty, costs within a window of each other under the stated model, and the shares it
    assumes (7/16, 5/16, 3/16, 1/16 by leading base) hold for canonical k-mers of random sequence."""
    from pykmer_b200 import dist as pdist
    T, align = 4 ** 19, 1 << 26
    for n in (1, 2, 3, 4, 8):
        for kmers in (0.0, 1e6, 780e6, 3e9):
            rs = pdist.analytic_kmer_ranges(T, n, kmers)
            assert rs[0][0] == 0 and rs[-1][1] == T and all(a[1] == b[0] for a, b in zip(rs[:-1], rs[1:]))
            assert all(lo < hi and lo % align == 0 for lo, hi in rs)
            if n > 1 and kmers == 780e6:
                def cost(lo, hi):
                    k = 0.0
                    for q, share in enumerate((7 / 16, 5 / 16, 3 / 16, 1 / 16)):
                        a, b = max(lo, q * T // 4), min(hi, (q + 1) * T // 4)
                        k += share * max(b - a, 0) / (T // 4)
                    return kmers * k / 20e9 + (hi - lo) / 7e12
                costs = [cost(lo, hi) for lo, hi in rs]
                window = kmers * (7 / 16) * align / (T // 4) / 20e9 + align / 7e12
                assert max(costs) - min(costs) <= 2 * window + 1e-12
                assert rs[0][1] - rs[0][0] < rs[-1][1] - rs[-1][0]          # crowded low end: narrower shards
    small = pdist.analytic_kmer_ranges(4 ** 9, 4, 1e6)                       # tiny table: plain equal ranges
    assert small[0][0] == 0 and small[-1][1] == 4 ** 9
    # the assumed shares, on random 9-mers
    rng = np.random.default_rng(1)
    codes = rng.integers(0, 4, size=(200_000, 9))
    w = 4 ** np.arange(8, -1, -1)
    fwd, rc = codes @ w, (3 - codes[:, ::-1]) @ w
    lead = np.minimum(fwd, rc) // 4 ** 8
    got = np.bincount(lead, minlength=4) / lead.size
    assert np.allclose(got, [7 / 16, 5 / 16, 3 / 16, 1 / 16], atol=0.01)


def test_routed_plan_gives_every_source_its_own_region_in_the_owner():
    """plan_routed (host logic of the routed exchange): the regions of all (source rank, window)
    pairs tile each owner's buffer without overlap, leave room for count * (1 + 1/16) + slack, stay
    below the owner's published-count table, and every rank derives the same layout."""
    from pykmer_b200 import dist as pdist
    rng = np.random.default_rng(5)
    nranks, nwin = 4, 23
    all_cnt = rng.integers(0, 100_000, size=(nranks, nwin)).astype(np.int64)
    all_cnt[1, 5:9] = 0
    owners = pdist.balanced_window_owners(all_cnt.sum(axis=0), nranks, overhead=10)
    pub = [1 << 30] * nranks
    plans = [pdist.plan_routed(all_cnt, owners, r, pub, slack=64) for r in range(nranks)]
    for d, (w0, w1) in enumerate(owners):
        spans = []
        for s in range(nranks):
            owner_of, dest_off, cap, _ = plans[s]
            assert (owner_of[w0:w1] == d).all()
            for w in range(w0, w1):
                assert cap[w] >= all_cnt[s, w] + all_cnt[s, w] // 16 + 64
                spans.append((int(dest_off[w]), int(dest_off[w]) + int(cap[w])))
                # the owner's import layout names the same place
                assert int(plans[d][3][s, w - w0]) == int(dest_off[w])
        spans.sort()
        assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert spans[-1][1] <= pub[d]
    with pytest.raises(ValueError):                                  # an owner whose buffer is too small
        pdist.plan_routed(all_cnt, owners, 0, [1000] * nranks)


def test_analytic_window_owners_follow_the_leading_base_shares():
    """Window ownership without a planning scan: contiguous, every rank owns a window, and the low
    windows -- where canonical k-mers crowd (7/16 start with A) -- are shared out more finely."""
    from pykmer_b200 import dist as pdist
    for nwin, nranks in ((64, 2), (64, 8), (256, 8), (4, 4), (5, 3), (3, 2)):
        owners = pdist.analytic_window_owners(nwin, nranks, kmers=7.8e8)
        assert owners[0][0] == 0 and owners[-1][1] == nwin
        assert all(a[1] == b[0] and a[0] < a[1] for a, b in zip(owners, owners[1:])) and owners[-1][0] < nwin
    owners = pdist.analytic_window_owners(64, 8, kmers=7.8e8)
    widths = [b - a for a, b in owners]
    assert widths[0] < widths[-1] and widths[0] <= min(widths) + 1
    # expected cost per rank (share of k-mers + 3 M per window) within 25 % of the mean
    per = np.zeros(64)
    for q, share in enumerate((7 / 16, 5 / 16, 3 / 16, 1 / 16)):
        per[16 * q:16 * q + 16] = 7.8e8 * share / 16
    cost = [per[a:b].sum() + 3e6 * (b - a) for a, b in owners]
    assert max(cost) < 1.25 * (sum(cost) / len(cost))


AGREE_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %(root)r)
    import torch.distributed as dist
    from pykmer_b200 import dist as pdist
    dist.init_process_group("gloo")
    rank, world = pdist.world()
    pdist.agree(None)                                   # nobody failed: returns on every rank
    try:
        pdist.agree(MemoryError("shard does not fit") if rank == 1 else None)
    except MemoryError as e:
        print("own", e)
    except RuntimeError as e:
        print("other", e)
    else:
        print("no error seen")
    dist.destroy_process_group()
""")


def test_agree_raises_on_every_rank_when_any_rank_fails(tmp_path):
    """dist.agree: a failure on a rank other than 0 (an allocation, a write) stops all ranks instead
    of leaving them in the next collective."""
    script = tmp_path / "agree_worker.py"
    script.write_text(AGREE_WORKER % {"root": ROOT})
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, str(script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                              env=dict(os.environ, RANK=str(r), WORLD_SIZE="3", MASTER_ADDR="127.0.0.1",
                                       MASTER_PORT=str(port)))
             for r in range(3)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "own shard does not fit" in outs[1]
    assert "other rank 1 failed: MemoryError: shard does not fit" in outs[0]
    assert "other rank 1 failed" in outs[2]


def test_read_routed_status_decodes_flags_and_64_bit_counts():
    """The status words every rank all-gathers in a routed step: overflow flag, num_kmers low / high."""
    import torch
    from pykmer_b200 import dist as pdist
    big = 5_000_000_123                                            # needs the high word
    rows = [[0, 781_061_717 & 0xFFFFFFFF, 0, 0], [0, big & 0xFFFFFFFF, big >> 32, 0], [0, 0, 0, 0]]
    every = torch.tensor(rows, dtype=torch.int64).to(torch.int32)  # low words above 2^31 wrap to negative int32
    overflow, total, per_rank = pdist.read_routed_status(every)
    assert not overflow and per_rank == [781_061_717, big, 0] and total == 781_061_717 + big
    rows[2][0] = 1
    assert pdist.read_routed_status(torch.tensor(rows, dtype=torch.int64).to(torch.int32))[0]

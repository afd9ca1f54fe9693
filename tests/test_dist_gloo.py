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

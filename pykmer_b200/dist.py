"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  Both hot paths shard along the canonical k-mer axis
(SURVEY.md 8e); the only exchanges are

  indexer  one all-reduce of hist[255] + num_kmers / vals_sum / vals_count (sum) and of
           vals_min / vals_max (min / max)  -- about 2 KB;
  merger   one all-reduce (sum) of the N x N int64 partial Gram matrices.

The sequence itself is either read by every rank or broadcast once from rank 0
(`broadcast_stream`).  Nothing here touches the data path of a single GPU.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

ALIGN = 4096          # shard boundaries are multiples of this many table entries


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env() -> Tuple[int, int, int]:
    """-> (rank, world, local device).  Under torchrun (WORLD_SIZE > 1) joins the job's process
    group -- NCCL, one rank per GPU; PYKMER_B200_DIST_BACKEND=gloo lets the CPU tests and a
    single-GPU box run several ranks -- and selects this rank's GPU; otherwise (0, 1, current)."""
    import os
    n = int(os.environ.get("WORLD_SIZE", "1"))
    ndev = torch.cuda.device_count() if torch.cuda.is_available() else 0
    local = int(os.environ.get("LOCAL_RANK", "0")) % max(ndev, 1)
    if ndev:
        torch.cuda.set_device(local)
    if n > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(os.environ.get("PYKMER_B200_DIST_BACKEND", "nccl"))
    rank, n = world()
    return rank, n, local


def raise_together(error: Optional[BaseException], src: int = 0, group=None) -> None:
    """Rank `src` reports whether its host-side step (opening files, parsing text) failed; every
    rank then raises -- nobody is left waiting in a collective for a rank that has gone."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if error is not None:
            raise error
        return
    box = [None if error is None else f"{type(error).__name__}: {error}"]
    dist.broadcast_object_list(box, src=src, group=group)
    if error is not None:
        raise error
    if box[0] is not None:
        raise RuntimeError(f"rank {src} failed: {box[0]}")


def agree(error: Optional[BaseException], group=None) -> None:
    """EVERY rank reports whether its last stage failed (allocating a 32 GiB shard, feeding, writing its
    slice of the table ...); if any did, every rank raises -- nobody is left waiting in the next
    collective for a rank that has gone.  The failing rank re-raises its own exception."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if error is not None:
            raise error
        return
    mine = None if error is None else f"{type(error).__name__}: {error}"
    every = [None] * dist.get_world_size(group)
    dist.all_gather_object(every, mine, group=group)
    if error is not None:
        raise error
    for r, msg in enumerate(every):
        if msg is not None:
            raise RuntimeError(f"rank {r} failed: {msg}")


def shard_range(total: int, rank: int, nranks: int, align: int = ALIGN) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the k-mer axis owned by `rank`; slices tile [0, total)."""
    def cut(r: int) -> int:
        if r <= 0:
            return 0
        if r >= nranks:
            return total
        return (total * r // nranks) // align * align
    return cut(rank), cut(rank + 1)


def _device_for_backend() -> torch.device:
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def reduce_index_stats(hist: List[int], st: Dict[str, int], group=None) -> Tuple[List[int], Dict[str, int]]:
    """Combine per-shard statistics into those of the whole table (tools.py:246-263):
    hist / num_kmers / vals_sum / vals_count add up, vals_min / vals_max are min / max."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(hist), dict(st)
    dev = _device_for_backend()
    n = dist.get_world_size(group)
    mine = torch.tensor(list(hist) + [st["num_kmers"], st["vals_sum"], st["vals_count"],
                                      st["vals_min"], st["vals_max"]], dtype=torch.int64, device=dev)
    every = torch.empty((n, mine.numel()), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(every.view(-1), mine, group=group)      # one collective, one copy back
    e = every.cpu().numpy()
    s = e[:, :258].sum(axis=0).tolist()
    return s[:255], {"num_kmers": s[255], "vals_sum": s[256], "vals_count": s[257],
                     "vals_min": int(e[:, 258].min()), "vals_max": int(e[:, 259].max())}


def reduce_flags(flags: np.ndarray, group=None) -> np.ndarray:
    """A record is listed if ANY shard counted one of its k-mers (indexer.py:349-351)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flags
    t = torch.from_numpy(flags.astype(np.int32)).to(_device_for_backend())
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.cpu().numpy().astype(np.uint8)


def reduce_gram(G: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the partial Gram matrices of the k-mer-axis shards, in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if G.is_cuda and dist.get_backend(group) != "nccl":       # gloo: through host memory
            g = G.cpu()
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            G.copy_(g)
        else:
            dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)
    return G


def broadcast_stream(chunk: Optional[torch.Tensor], nbytes: int, src: int = 0, group=None) -> torch.Tensor:
    """Broadcast one chunk of the cleaned sequence stream from `src` (NCCL: over NVLink).
    Ranks other than `src` pass chunk=None and receive a fresh uint8 tensor of nbytes."""
    dev = _device_for_backend()
    if chunk is None:
        chunk = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(chunk, src=src, group=group)
    return chunk


# ---------------------------------------------------------------------------------------------
# Sequence-sharded indexing: every rank scans 1/N of the stream (PK_MODE_SCAN handle over the
# full k-mer range), the bucketed entries travel to the rank that owns their table window with
# ONE all-to-all over NVLink, and each rank counts its own windows (PK_MODE_PARTITION handle).

def window_owner_ranges(nwindows: int, nranks: int) -> List[Tuple[int, int]]:
    """Contiguous window ranges [w0, w1) per rank; they tile [0, nwindows)."""
    return [(nwindows * r // nranks, nwindows * (r + 1) // nranks) for r in range(nranks)]


def balanced_window_owners(per_window: np.ndarray, nranks: int, overhead: int = 3_000_000) -> List[Tuple[int, int]]:
    """Contiguous window ranges with about equal cost, cost(window) = its k-mer entries + a
    fixed per-window overhead (the commit pass).  Canonical k-mers crowd the low windows
    (first-base shares 7/16, 5/16, 3/16, 1/16), so equal window counts would leave rank 0 with
    ~1.9x the mean."""
    cost = per_window.astype(np.int64) + overhead
    nwin = len(cost)
    cum = np.concatenate(([0], np.cumsum(cost)))
    cuts = [0]
    for r in range(1, nranks):
        target = cum[-1] * r / nranks
        w = int(np.searchsorted(cum, target))
        w = min(max(w, cuts[-1] + 1), nwin - (nranks - r))      # every rank owns >= 1 window
        cuts.append(w)
    cuts.append(nwin)
    return [(cuts[r], cuts[r + 1]) for r in range(nranks)]


def analytic_window_owners(nwindows: int, nranks: int, kmers: float, overhead: int = 3_000_000) -> List[Tuple[int, int]]:
    """balanced_window_owners WITHOUT a planning scan (the CLI sees the stream once): the expected
    entries per window follow the leading base of a canonical k-mer -- min(fwd, rc) starts with A, C,
    G, T with probability 7/16, 5/16, 3/16, 1/16 -- spread evenly inside each quarter of the axis."""
    if nwindows < 4 * 1 or nwindows < nranks:
        return window_owner_ranges(nwindows, nranks)
    shares = (7 / 16, 5 / 16, 3 / 16, 1 / 16)
    per_window = np.zeros(nwindows, dtype=np.float64)
    edges = [nwindows * q // 4 for q in range(5)]
    for q in range(4):
        n = edges[q + 1] - edges[q]
        if n:
            per_window[edges[q]:edges[q + 1]] = kmers * shares[q] / n
    return balanced_window_owners(per_window.astype(np.int64), nranks, overhead=overhead)


def balanced_kmer_ranges(per_window: np.ndarray, nranks: int, window_log2: int, total: int,
                         window_cost: int = 230_000) -> List[Tuple[int, int]]:
    """k-mer-axis shards [lo, hi) of about equal cost for the DIRECT counting scheme of very
    sparse tables (K >= 19), cut at window boundaries.  cost = k-mers that fall into the shard +
    `window_cost` k-mer equivalents per 2^window_log2 table entries (zero-filling 64 MiB takes
    about as long as 230 k table updates of the DIRECT scan: 9.3 us against 41 ps each, measured).  Equal
    ranges would leave rank 0 with 2.2x the mean (canonical k-mers crowd the low values)."""
    owners = balanced_window_owners(per_window, nranks, overhead=window_cost)
    return [(w0 << window_log2, min(total, w1 << window_log2)) for w0, w1 in owners]


def analytic_kmer_ranges(total: int, nranks: int, kmers: float, align: int = 1 << 26,
                         update_rate: float = 20e9, fill_rate: float = 7e12) -> List[Tuple[int, int]]:
    """k-mer-axis shards [lo, hi) of about equal cost WITHOUT a planning pass, for the CLI's DIRECT
    counting of very sparse tables (K >= 19).  cost(shard) = its share of the `kmers` table updates at
    `update_rate` per second (measured 20 G/s) + its bytes of zero-fill at `fill_rate`
    (measured 7.2 TB/s).  The share comes from the leading base of a canonical k-mer: min(fwd, rc)
    starts with A, C, G, T with probability 7/16, 5/16, 3/16, 1/16 (either of two independent first
    bases is the smaller one), uniform inside a quarter to first order -- equal ranges would leave
    ranks 0 and 1 with 1.75x the mean.  Every rank computes the same cuts from the same three numbers;
    cuts are multiples of `align` entries (a counting window of any mode)."""
    if nranks <= 1 or total < nranks * align:
        return [shard_range(total, r, nranks, align=min(align, ALIGN)) for r in range(nranks)]
    units = total // align                                   # the axis in whole windows
    quarter = units / 4.0
    shares = (7 / 16, 5 / 16, 3 / 16, 1 / 16)
    per_unit_fill = align / fill_rate

    def cost_upto(u: float) -> float:                        # cost of [0, u) windows
        k = 0.0
        for q, share in enumerate(shares):
            k += share * min(max(u - q * quarter, 0.0), quarter) / quarter
        return kmers * k / update_rate + u * per_unit_fill

    whole = cost_upto(float(units))
    cuts = [0]
    for r in range(1, nranks):
        want, lo, hi = whole * r / nranks, cuts[-1] + 1, units - (nranks - r)
        while lo < hi:                                        # smallest u with cost_upto(u) >= want
            mid = (lo + hi) // 2
            if cost_upto(float(mid)) < want:
                lo = mid + 1
            else:
                hi = mid
        cuts.append(lo)
    cuts.append(units)
    return [(cuts[r] * align, total if r == nranks - 1 else cuts[r + 1] * align) for r in range(nranks)]


def slice_bounds(nbytes: int, rank: int, nranks: int, align: int = 16) -> Tuple[int, int]:
    """Byte slice [a, b) of the stream scanned by `rank` (16-byte aligned starts)."""
    def cut(r: int) -> int:
        if r <= 0:
            return 0
        if r >= nranks:
            return nbytes
        return (nbytes * r // nranks) // align * align
    return cut(rank), cut(rank + 1)


def plan_exchange(all_cnt: np.ndarray, owners: List[Tuple[int, int]], rank: int):
    """all_cnt[s, f, w] = entries of source rank s, segment f, window w (uint32/int64).
    -> (send_counts[f][d], recv_counts[f][s], imp_off, imp_cnt) where imp_* are the segment
    tables (nranks * nseg, local windows) of the receive buffer laid out segment by segment,
    source by source, window by window."""
    nranks, nseg, _ = all_cnt.shape
    w0, w1 = owners[rank]
    send = np.zeros((nseg, nranks), dtype=np.int64)
    recv = np.zeros((nseg, nranks), dtype=np.int64)
    for f in range(nseg):
        for d, (a, b) in enumerate(owners):
            send[f, d] = int(all_cnt[rank, f, a:b].sum())
        for s in range(nranks):
            recv[f, s] = int(all_cnt[s, f, w0:w1].sum())
    imp_cnt = np.zeros((nseg * nranks, w1 - w0), dtype=np.uint32)
    imp_off = np.zeros((nseg * nranks, w1 - w0), dtype=np.uint32)
    base = 0
    for f in range(nseg):
        for s in range(nranks):
            c = all_cnt[s, f, w0:w1].astype(np.int64)
            imp_cnt[f * nranks + s] = c
            imp_off[f * nranks + s] = base + np.concatenate(([0], np.cumsum(c)[:-1]))
            base += int(c.sum())
    return send, recv, imp_off, imp_cnt, base


def gather_window_counts(scanner, group=None) -> np.ndarray:
    """all_cnt[s, f, w]: entries of rank s, segment f, window w (a few KB, all-gathered)."""
    rank, nranks = world()
    entries, off, cnt = scanner.export_segments()
    nwin = cnt.shape[1] if cnt.ndim == 2 and cnt.shape[0] else scanner.mode()[1]
    dev = entries.device
    nseg_t = torch.tensor([cnt.shape[0]], dtype=torch.int64, device=dev)
    dist.all_reduce(nseg_t, op=dist.ReduceOp.MAX, group=group)
    nseg = int(nseg_t.item())
    mine = np.zeros((nseg, nwin), dtype=np.int64)
    mine[:cnt.shape[0]] = cnt
    gathered = torch.empty((nranks, nseg, nwin), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered.view(-1), torch.from_numpy(mine).to(dev).view(-1), group=group)
    return gathered.cpu().numpy()


def exchange_entries(scanner, counter, owners: Optional[List[Tuple[int, int]]] = None, group=None):
    """All-to-all of the scanner's bucketed k-mer entries to the window owners; the counter of
    this rank is left holding (importing) everything that falls into its windows.
    owners: window range per rank (default: equal window counts).
    Returns the receive buffer (keep it alive until counter.finalize())."""
    import os
    import time
    verbose = os.environ.get("PYKMER_B200_VERBOSE")
    t = [time.perf_counter()]
    rank, nranks = world()
    entries, off, cnt = scanner.export_segments()
    t.append(time.perf_counter())
    nwin = cnt.shape[1] if cnt.ndim == 2 and cnt.shape[0] else scanner.mode()[1]
    dev = entries.device
    # every rank needs everybody's per-window counts (a few KB); row 0 of the payload carries
    # the segment count so that one all-gather is enough
    cap = 8
    assert cnt.shape[0] <= cap, "more than 8 feed segments per exchange"
    mine = np.zeros((cap + 1, nwin), dtype=np.int64)
    mine[0, 0] = cnt.shape[0]
    mine[1:1 + cnt.shape[0]] = cnt
    gathered = torch.empty((nranks, cap + 1, nwin), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered.view(-1), torch.from_numpy(mine).to(dev).view(-1), group=group)
    g = gathered.cpu().numpy()
    nseg = int(g[:, 0, 0].max())
    all_cnt = g[:, 1:1 + nseg, :]
    t.append(time.perf_counter())
    if owners is None:
        owners = window_owner_ranges(nwin, nranks)
    send, recv, imp_off, imp_cnt, total = plan_exchange(all_cnt, owners, rank)
    recv_buf = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    t.append(time.perf_counter())
    pos = 0
    for f in range(nseg):
        if f < cnt.shape[0]:
            a = int(off[f, 0])
            src = entries[a:a + int(send[f].sum())]          # windows of one segment are contiguous
        else:
            src = entries[:0]
        n_in = int(recv[f].sum())
        dist.all_to_all_single(recv_buf[pos:pos + n_in], src, output_split_sizes=recv[f].tolist(),
                               input_split_sizes=send[f].tolist(), group=group)
        pos += n_in
    if verbose:
        torch.cuda.synchronize()
    t.append(time.perf_counter())
    counter.import_segments(recv_buf, imp_off, imp_cnt)
    t.append(time.perf_counter())
    if verbose and rank == 0:
        names = ("export", "gather", "plan", "all_to_all", "import")
        print("[pykmer_b200] exchange ms: " + ", ".join(f"{n} {1e3 * (b - a):.3f}" for n, a, b in zip(names, t[:-1], t[1:]))
              + f"; sent {4 * int(send.sum() - send[:, rank].sum())} B", flush=True)
    return recv_buf


# ---------------------------------------------------------------------------------------------
# Fused exchange: no all-to-all at all.  Pass 2 of every scanner stores its entries directly
# into the k-mer buffer of the rank that owns the window (CUDA-IPC peer mapping, NVLink), so the
# transfer rides on the scatter kernel's own coalesced stores.

def plan_fused(all_cnt: np.ndarray, owners: List[Tuple[int, int]], rank: int):
    """all_cnt[s, w] = entries of source rank s in window w.  Destination buffers are laid out
    source by source, window by window.  -> (owner_of[w], dest_off_mine[w], imp_off, imp_cnt,
    landed) with imp_* the (nranks, local windows) tables of what lands in THIS rank's buffer."""
    nranks, nwin = all_cnt.shape
    owner_of = np.zeros(nwin, dtype=np.uint32)
    dest_off = np.zeros(nwin, dtype=np.uint32)
    for d, (a, b) in enumerate(owners):
        owner_of[a:b] = d
        base = 0
        for s in range(nranks):
            c = all_cnt[s, a:b].astype(np.int64)
            offs = base + np.concatenate(([0], np.cumsum(c)[:-1])) if b > a else np.zeros(0, dtype=np.int64)
            if s == rank:
                dest_off[a:b] = offs
            base += int(c.sum())
    w0, w1 = owners[rank]
    imp_cnt = all_cnt[:, w0:w1].astype(np.uint32)
    imp_off = np.zeros_like(imp_cnt)
    base = 0
    for s in range(nranks):
        c = all_cnt[s, w0:w1].astype(np.int64)
        imp_off[s] = base + np.concatenate(([0], np.cumsum(c)[:-1])) if w1 > w0 else 0
        base += int(c.sum())
    return owner_of, dest_off, imp_off, imp_cnt, base


def connect_peer_pools(scanner, counter, group=None) -> None:
    """Once per job: every scanner maps every counter's k-mer buffer (IPC handles all-gathered)."""
    rank, nranks = world()
    handle, cap = counter.pool_ipc_handle()
    dev = _device_for_backend()
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    allh = torch.empty((nranks, 64), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allh.view(-1), mine, group=group)
    allh = allh.cpu().numpy()
    for d in range(nranks):
        if d == rank:
            scanner.open_peer_pool(d, local_owner=counter)
        else:
            scanner.open_peer_pool(d, handle=allh[d].tobytes())


def exchange_fused(scanner, counter, seq: Optional[torch.Tensor], owners: List[Tuple[int, int]], group=None) -> np.ndarray:
    """One sequence-sharded scan step with the exchange fused into pass 2.  `seq` is this rank's
    slice (the scanner must have been reset / primed); None or empty = this rank has nothing to scan
    in this step but still takes part in the collectives.  Leaves the counter importing what
    landed in its buffer; returns all_cnt."""
    rank, nranks = world()
    have = seq is not None and seq.numel() > 0
    if have:
        scanner.scan_pass1(seq)
        cnt = scanner.pass1_counts()
    else:
        cnt = np.zeros(scanner.mode()[1], dtype=np.uint32)
    dev = _device_for_backend()
    gathered = torch.empty((nranks, cnt.size), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered.view(-1), torch.from_numpy(cnt.astype(np.int64)).to(dev), group=group)
    all_cnt = gathered.cpu().numpy()
    owner_of, dest_off, imp_off, imp_cnt, landed = plan_fused(all_cnt, owners, rank)
    if have:
        scanner.scan_pass2_remote(nranks, owner_of, dest_off)      # synchronises: stores have landed
    dist.barrier(group=group)                                       # ... on every rank
    counter.import_own_pool(imp_off, imp_cnt)
    return all_cnt


# ---------------------------------------------------------------------------------------------
# Routed exchange: the fused exchange without a host round trip inside a step.  A planning scan
# (exact per-window counts of every rank, once, untimed) sizes a fixed region per (source rank,
# window) in the owner's buffer; a step is then ONE scan pass per rank that stores into the regions
# and publishes its fill counts into a table at the tail of every owner's buffer, ONE stream-ordered
# all-reduce (it carries the overflow flags and orders every rank's stores before every owner's
# count), and the owners' count + commit.  The host only reads the statistics at the end.

def plan_routed(all_cnt: np.ndarray, owners: List[Tuple[int, int]], rank: int, pub_base: Sequence[int],
                slack: int = 4096):
    """all_cnt[s, w] = entries of source rank s in window w (planning scan).  Room of a region =
    count * (1 + 1/16) + slack.  -> (owner_of[w], dest_off_mine[w], cap_mine[w], imp_off[s, local w])."""
    nranks, nwin = all_cnt.shape
    cnt = all_cnt.astype(np.int64)
    cap = cnt + cnt // 16 + slack
    owner_of = np.zeros(nwin, dtype=np.uint32)
    off = np.zeros((nranks, nwin), dtype=np.int64)
    for d, (a, b) in enumerate(owners):
        owner_of[a:b] = d
        flat = cap[:, a:b].reshape(-1)                       # source by source, window by window
        start = np.concatenate(([0], np.cumsum(flat)[:-1])) if flat.size else np.zeros(0, dtype=np.int64)
        off[:, a:b] = start.reshape(nranks, b - a)
        total = int(flat.sum())
        if total > int(pub_base[d]):
            raise ValueError(f"routed exchange: rank {d} would receive up to {total} entries, its buffer holds {int(pub_base[d])}")
    assert off.max(initial=0) < 2 ** 32 and cap.max(initial=0) < 2 ** 32
    w0, w1 = owners[rank]
    return owner_of, off[rank].astype(np.uint32), cap[rank].astype(np.uint32), off[:, w0:w1].astype(np.uint32)


def setup_routed(scanner, counter, all_cnt: np.ndarray, owners: List[Tuple[int, int]], group=None) -> None:
    """Once per job, after connect_peer_pools: exchange the owners' table positions, give the
    scanner its route and the counter its import layout."""
    rank, nranks = world()
    dev = _device_for_backend()
    mine = torch.tensor([counter.pub_base()], dtype=torch.int64, device=dev)
    every = torch.empty(nranks, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(every, mine, group=group)
    pub = [int(v) for v in every.cpu().tolist()]
    owner_of, dest_off, cap, imp_off = plan_routed(all_cnt, owners, rank, pub)
    scanner.set_route(nranks, rank, owner_of, dest_off, cap, pub)
    counter.set_import_layout(imp_off, owners[rank][0], all_cnt.shape[1])


def exchange_routed(scanner, counter, seq: torch.Tensor, status: torch.Tensor, group=None) -> torch.Tensor:
    """One step of the routed exchange (scanner reset + primed, counter reset; setup_routed done).
    Nothing here waits on the host.  Returns the status words of every rank, (nranks, 4) int32 on the
    device: after counter.finalize() the caller reads them once -- read_routed_status()."""
    scanner.scan_routed(seq, status)
    rank, nranks = world()
    every = torch.empty((nranks, status.numel()), dtype=status.dtype, device=status.device)
    # ONE small collective: it carries the overflow flags and every scanner's num_kmers, and it orders
    # every rank's stores (entries and fill counts) before every owner's count
    dist.all_gather_into_tensor(every.view(-1), status, group=group)
    counter.import_published()
    return every


def read_routed_status(every: torch.Tensor) -> Tuple[bool, int, List[int]]:
    """-> (some region overflowed on some rank: redo the step with exchange_fused,
           num_kmers of the whole job, num_kmers per rank)."""
    e = every.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    per_rank = [int(lo) | (int(hi) << 32) for lo, hi in zip(e[:, 1], e[:, 2])]
    return bool(e[:, 0].any()), sum(per_rank), per_rank

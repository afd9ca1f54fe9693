"""indexer: FASTA -> dense saturating uint8 k-mer count table (.kin + .kin.json).

Host-side mirror of the reference's indexer.py (create_fasta_index
indexer.py:299-414, read_fasta_index :416-444, main :475-495): same arguments,
same output files and JSON keys, same errors.  The per-base work -- CONV,
gen_kmers, min(fwd, rev), the saturating accumulate of process_kmers and
Header.update_stats -- runs in the CUDA library (pykmer_b200/csrc/indexer.cu)
through the C ABI; this module only parses text, moves bytes through pinned
buffers and writes files.
"""
from __future__ import annotations

import hashlib
import os
import sys
import time
from typing import List, Optional

import numpy as np

from .fasta import FastaStream
from .tools import Header


def _write_table(path: str, table: np.ndarray, block: int = 64 << 20) -> str:
    """Write the table over the (sparse) tmp file and return its sha256 (the hash runs on a
    second thread beside the writes; both release the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    mv = memoryview(table)

    def digest() -> str:
        h = hashlib.sha256()
        for off in range(0, len(mv), block):
            h.update(mv[off:off + block])
        return h.hexdigest()

    with ThreadPoolExecutor(max_workers=1) as pool:
        job = pool.submit(digest)
        with open(path, "r+b") as fh:
            for off in range(0, len(mv), block):
                fh.write(mv[off:off + block])
        return job.result()


def create_fasta_index(project_name: str, sample_name: Optional[str], input_file: str,
                       kmer_len: int, overwrite: bool,
                       flush_every: int = Header.DEFAULT_FLUSH_EVERY,
                       min_frag_size: int = Header.DEFAULT_MIN_FRAG_SIZE,
                       max_frag_size: int = Header.DEFAULT_MAX_FRAG_SIZE,
                       buffer_size: int = Header.DEFAULT_BUFFER_SIZE,
                       debug: bool = False, device: int = 0,
                       chunk_bytes: int = 64 << 20) -> Header:
    """Index one FASTA file (indexer.py:299-414).  flush_every / frag sizes only end
    up in the JSON, as in the reference where results never depended on them."""
    from . import device as dev          # needs the CUDA library; no fallback
    from . import dist as pdist

    header = Header(project_name, sample_name=sample_name, input_file=input_file,
                    kmer_len=kmer_len, flush_every=flush_every, min_frag_size=min_frag_size,
                    max_frag_size=max_frag_size, buffer_size=buffer_size)
    print(f"project_name {header.project_name} sample_name {header.sample_name} "
          f"kmer_len {header.kmer_len:15,d} kmer_size {header.kmer_size:15,d}")
    if pdist.world()[1] > 1:             # torchrun indexer.py ...: one rank per GPU, k-mer-axis shards
        return _create_fasta_index_sharded(header, input_file, overwrite, device, chunk_bytes)
    header.init_index_tmp_file(overwrite=overwrite)
    t_start = time.perf_counter()

    # off the critical path: the sha256 of the input file (tools.py:279) and the pinned buffer
    # the table lands in are produced by helper threads while the genome is read and indexed
    from concurrent.futures import ThreadPoolExecutor
    from .tools import gen_checksum
    helpers = ThreadPoolExecutor(max_workers=2)
    input_sum = helpers.submit(gen_checksum, header.input_file_path)
    out_buf = helpers.submit(dev.pinned_empty, header.data_size)

    fs = FastaStream(input_file, chunk_bytes=chunk_bytes)
    with dev.Indexer(kmer_len, device=device) as ix:
        # the reader writes its pieces straight into two pinned buffers in turn; a piece is fed
        # (pinned -> device, asynchronous) and the handle drained before its buffer comes round again
        cap = chunk_bytes + (chunk_bytes >> 3) + (1 << 17)
        ring = [dev.pinned_empty(cap) for _ in range(2)]
        views = [r.numpy() for r in ring]
        nrec = 0
        for piece in fs.pieces(buffers=views):
            ix.append_records(fs.starts[nrec:])        # only the records opened since the last piece
            nrec = len(fs.starts)
            ix.feed_host(piece)
            ix.sync()                     # the buffer may be rewritten from here on
            header.timer.update(fs.bases)
        out = out_buf.result()
        hist, st = ix.finalize(table_out=out)      # table windows stream out as they are committed
        flags = ix.record_flags() if fs.starts else np.zeros(0, dtype=np.uint8)
        table = out.numpy()

    header.num_kmers = st["num_kmers"]
    # indexer.py:349-351: a record is listed once its first k-mer arrives
    header.chromosomes = [[fs.names[i], fs.lengths[i]] for i in range(len(fs.names)) if flags[i]]
    header.set_stats(hist, st["vals_sum"], st["vals_count"], st["vals_min"], st["vals_max"])
    print(f"project_name {header.project_name} kmer_len {header.kmer_len:15,d} "
          f"num_kmers {header.num_kmers:15,d} kmer_size {header.kmer_size:15,d}")

    t_gpu = time.perf_counter()
    checksum = _write_table(header.index_tmp_file, table)
    t_write = time.perf_counter()
    header.write_metadata_index_tmp_file(output_checksum=checksum, input_checksum=input_sum.result())
    helpers.shutdown()
    os.rename(header.index_tmp_file, header.index_file)          # indexer.py:412
    t_end = time.perf_counter()
    header.wall_seconds = {"ingest_and_gpu": t_gpu - t_start, "write_and_sha256_table": t_write - t_gpu,
                           "metadata_and_input_sha256": t_end - t_write, "total": t_end - t_start}
    print("  wall: ingest+GPU {ingest_and_gpu:.2f} s, table write+sha256 {write_and_sha256_table:.2f} s, "
          "metadata {metadata_and_input_sha256:.2f} s, total {total:.2f} s".format(**header.wall_seconds))
    return header


def _create_fasta_index_sharded(header: Header, input_file: str, overwrite: bool, device: int,
                                chunk_bytes: int) -> Header:
    """The same job inside a torch.distributed group.  Rank 0 reads the FASTA and
    broadcasts every cleaned piece of the stream (NCCL: over NVLink); the table is partitioned along
    the canonical k-mer axis, every rank writes its slice of the .kin at its own offset, and the
    statistics meet in one small collective (hist / num_kmers / vals_sum / vals_count add up,
    vals_min / vals_max are min / max, a record is listed if any rank counted one of its k-mers).
    Rank 0 writes the JSON.  Two schemes, the ones bench.py measures:

    * K <= 17 -- sequence slices + fused exchange: rank r scans only slice r of every piece (scan-only
      handle over the full k-mer range) and stores the bucketed k-mer entries straight into the
      buffer of the rank that owns their table window (CUDA-IPC peer mapping, pykmer_b200/dist.py
      exchange_fused); each rank counts its own windows.  Windows are owned in contiguous ranges
      balanced on the leading-base shares of canonical k-mers (dist.analytic_window_owners).
    * K >= 19 (BASELINE config 5: the 256 GiB table does not fit one GPU), tables with fewer windows
      than ranks, and PYKMER_B200_SHARD=kmer -- k-mer ranges, replicated scan: every rank scans the
      whole piece and counts only the canonical values of its own range (balanced analytically for
      K >= 19, dist.analytic_kmer_ranges).

    A failure on any rank at any stage (refusing to overwrite, text the reader rejects, a shard that
    does not fit, a full disk) is agreed on by all ranks (dist.agree) and raised everywhere."""
    import torch
    import torch.distributed as tdist
    from . import device as dev, dist as pdist
    from . import _native as nat
    from .tools import gen_checksum

    rank, world = pdist.world()
    K, T = header.kmer_len, header.data_size
    size = os.path.getsize(header.input_file_path)
    packed = header.input_file_path.endswith((".gz", ".bgz"))
    est_kmers = size * (3.5 if packed else 1.0)              # about 3.5 bases per compressed byte
    t_start = time.perf_counter()
    error = None
    if rank == 0:
        try:
            header.init_index_tmp_file(overwrite=overwrite)
        except (ValueError, OSError) as exc:
            error = exc
    pdist.agree(error)

    from concurrent.futures import ThreadPoolExecutor
    helpers = ThreadPoolExecutor(max_workers=2)
    input_sum = helpers.submit(gen_checksum, header.input_file_path) if rank == 0 else None

    fs, pieces = None, None
    if rank == 0:
        fs = FastaStream(input_file, chunk_bytes=chunk_bytes)
        cap = chunk_bytes + (chunk_bytes >> 3) + (1 << 17)
        ring = [dev.pinned_empty(cap) for _ in range(2)]
        pieces = fs.pieces(buffers=[r.numpy() for r in ring])

    with dev.device_scope(device):
        # ---- handles ---------------------------------------------------------------------------
        scanner, ix, owners, error = None, None, None, None
        lo, hi = 0, 0
        try:
            seq_mode = (K <= 17 and getattr(dev, "HAS_SCAN_MODE", False)
                        and os.environ.get("PYKMER_B200_SHARD", "sequence") != "kmer")
            if seq_mode:
                scanner = dev.Indexer(K, device=device, mode=nat.PK_MODE_SCAN)
                wl, nwin = scanner.window_log2(), scanner.mode()[1]
                if nwin < world:                              # fewer table windows than ranks
                    scanner.close()
                    scanner, seq_mode = None, False
            if seq_mode:
                owners = pdist.analytic_window_owners(nwin, world, est_kmers)
                lo, hi = owners[rank][0] << wl, min(T, owners[rank][1] << wl)
                ix = dev.Indexer(K, device=device, range_lo=lo, range_hi=hi, mode=nat.PK_MODE_PARTITION)
            else:
                if K >= 19:
                    # very sparse table, counted DIRECT: balance updates + zero-fill analytically
                    lo, hi = pdist.analytic_kmer_ranges(T, world, est_kmers)[rank]
                else:
                    lo, hi = pdist.shard_range(T, rank, world)
                ix = dev.Indexer(K, device=device, range_lo=lo, range_hi=hi) if hi > lo else None
        except Exception as exc:                              # e.g. the shard does not fit this GPU
            error = exc
        pdist.agree(error)
        if seq_mode:
            pdist.connect_peer_pools(scanner, ix)
        tagger = scanner if seq_mode else ix                  # the handle that flags records

        # ---- the stream, piece by piece ---------------------------------------------------------
        nrec, stream_off = 0, 0
        tail = None                                           # the last 32 bytes fed so far (device)
        while True:
            note, piece, error = [None], None, None
            if rank == 0:
                try:
                    piece = next(pieces, None)
                except (ValueError, OSError) as exc:       # text the reader rejects, damaged .bgz
                    error = exc
                if piece is not None:
                    new = [int(v) for v in fs.starts[nrec:]]
                    note = [(int(piece.size), new)]        # only the records opened since the last piece
                elif error is not None:
                    note = [("error",)]
            tdist.broadcast_object_list(note, src=0)
            if note[0] is not None and note[0][0] == "error":
                pdist.agree(error)
            if note[0] is None:
                break
            nbytes, new_starts = note[0]
            nrec += len(new_starts)
            error = None
            try:
                if tagger is not None and new_starts:
                    tagger.append_records(new_starts)
                if nbytes:
                    chunk = None
                    if rank == 0:
                        chunk = torch.from_numpy(piece)
                        if tdist.get_backend() == "nccl":
                            chunk = dev.upload(chunk)
                    chunk = pdist.broadcast_stream(chunk, nbytes, src=0)
                    d_chunk = dev.upload(chunk)
                    if seq_mode:
                        # rank r scans slice r of the piece; a short piece goes to rank 0 whole
                        a, b = pdist.slice_bounds(nbytes, rank, world) if nbytes >= 4096 * world else \
                            ((0, nbytes) if rank == 0 else (nbytes, nbytes))
                        ext = d_chunk if tail is None else torch.cat((tail, d_chunk))
                        shift = 0 if tail is None else tail.numel()
                        halo = ext[max(0, a + shift - 32):a + shift].clone() if a + shift > 0 else None
                        scanner.prime(halo, stream_off + a)
                        part = d_chunk[a:b]
                        pdist.exchange_fused(scanner, ix, part if b > a else None, owners)
                        ix.flush()                         # count what landed; the buffer is free for the next piece
                        ix.sync()
                        scanner.sync()
                        tail = ext[-32:].clone()
                        del ext, part
                    elif ix is not None:
                        ix.feed_device(d_chunk)
                        ix.sync()                          # the piece may be dropped from here on
                    stream_off += nbytes
                    del d_chunk, chunk
            except Exception as exc:
                error = exc
            pdist.agree(error)
            if rank == 0:
                header.timer.update(fs.bases)

        # ---- statistics ------------------------------------------------------------------------
        error = None
        try:
            if ix is not None:
                hist, st = ix.finalize()
            else:                                              # tiny table, more ranks than slices
                hist = [0] * 255
                st = {"num_kmers": 0, "vals_sum": 0, "vals_count": 0, "vals_min": 255, "vals_max": 0}
            if seq_mode:
                st["num_kmers"] = scanner.scan_result()
            flags = tagger.record_flags() if (tagger is not None and nrec) else np.zeros(nrec, dtype=np.uint8)
        except Exception as exc:
            error = exc
        pdist.agree(error)
        hist, st = pdist.reduce_index_stats(hist, st)
        if nrec:
            flags = pdist.reduce_flags(flags)
        t_gpu = time.perf_counter()

        # every rank writes its own slice of the table at its own offset of the (sparse) tmp file
        error = None
        try:
            if ix is not None:
                step = 256 << 20
                stage = [dev.pinned_empty(min(hi - lo, step)) for _ in range(2)]
                fd = os.open(header.index_tmp_file, os.O_WRONLY)
                try:
                    jobs = [None, None]
                    for i, off in enumerate(range(0, hi - lo, step)):
                        n = min(step, hi - lo - off)
                        if jobs[i & 1] is not None:
                            jobs[i & 1].result()               # its buffer is free again
                        ix.table_to_host(dst=stage[i & 1], offset=off, nbytes=n)
                        jobs[i & 1] = helpers.submit(os.pwrite, fd, memoryview(stage[i & 1].numpy())[:n], lo + off)
                    for job in jobs:
                        if job is not None:
                            job.result()
                finally:
                    os.close(fd)
        except Exception as exc:                               # a full disk, a vanished directory
            error = exc
        pdist.agree(error)
        for h in (scanner, ix):
            if h is not None:
                h.close()
    t_write = time.perf_counter()

    header.num_kmers = st["num_kmers"]
    header.set_stats(hist, st["vals_sum"], st["vals_count"], st["vals_min"], st["vals_max"])
    error = None
    if rank == 0:
        # indexer.py:349-351: a record is listed once its first k-mer arrives
        header.chromosomes = [[fs.names[i], fs.lengths[i]] for i in range(len(fs.names)) if flags[i]]
        print(f"project_name {header.project_name} kmer_len {header.kmer_len:15,d} "
              f"num_kmers {header.num_kmers:15,d} kmer_size {header.kmer_size:15,d}")
        try:
            header.write_metadata_index_tmp_file(input_checksum=input_sum.result())
            os.rename(header.index_tmp_file, header.index_file)      # indexer.py:412
        except (AssertionError, OSError) as exc:           # tools.py:367-368: no k-mer at all
            error = exc
    helpers.shutdown()
    pdist.agree(error)
    t_end = time.perf_counter()
    header.wall_seconds = {"ingest_and_gpu": t_gpu - t_start, "write_table": t_write - t_gpu,
                           "metadata_and_sha256": t_end - t_write, "total": t_end - t_start}
    if rank == 0:
        scheme = "sequence slices + fused exchange" if seq_mode else "k-mer ranges, replicated scan"
        print("  wall ({} ranks, {}): ingest+GPU {ingest_and_gpu:.2f} s, table write {write_table:.2f} s, "
              "metadata + sha256 {metadata_and_sha256:.2f} s, total {total:.2f} s".format(world, scheme, **header.wall_seconds))
    return header


def read_fasta_index(project_name: str, input_file: Optional[str] = None,
                     kmer_len: Optional[int] = None, index_file: Optional[str] = None,
                     debug: bool = False) -> Header:
    """Self-check of a .kin against its JSON (indexer.py:416-444)."""
    header = Header(project_name, input_file=input_file, kmer_len=kmer_len, index_file=index_file)
    header.read_metadata()
    header.check_data_index()
    print("OK")
    return header


def main(argv: Optional[List[str]] = None) -> None:
    """indexer.py <input_file> <sample_name> <kmer_len>   (indexer.py:475-495)
    The older two-argument form `indexer.py <file> <K>` of the README is accepted too."""
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) == 2:
        input_file, sample_name, kmer_len = argv[0], os.path.basename(argv[0]), int(argv[1])
    elif len(argv) == 3:
        input_file, sample_name, kmer_len = argv[0], argv[1], int(argv[2])
    else:
        print("usage: indexer.py <input_file> <sample_name> <kmer_len>", file=sys.stderr)
        sys.exit(1)
    project_name = input_file                                       # indexer.py:484
    print(f"project_name {project_name:s} input_file {input_file:s} sample_name {sample_name:s} "
          f"kmer_len {kmer_len:15,d}")
    device = 0
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:                  # torchrun indexer.py ...
        from . import dist as pdist
        device = pdist.init_from_env()[2]
    create_fasta_index(project_name, sample_name, input_file, kmer_len, buffer_size=2 ** 16,
                       overwrite=True, debug=False, device=device)
    print()


if __name__ == "__main__":
    main()

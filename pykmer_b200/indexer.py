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
        for piece in fs.pieces(buffers=views):
            ix.set_records(fs.starts)
            ix.feed_host(piece)
            ix.sync()                     # the buffer may be rewritten from here on
            header.timer.update(sum(fs.lengths))
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
    """The same job inside a torch.distributed group (BASELINE config 5: the
    256 GiB table of K=19 does not fit one GPU).  Rank g owns the canonical k-mer values
    [g * 4^K / G, (g+1) * 4^K / G): rank 0 reads the FASTA and broadcasts every cleaned piece of
    the stream (NCCL: over NVLink), every rank scans it and counts the k-mers of its own range,
    writes its slice of the .kin at its own offset, and the statistics meet in one small
    collective (hist / num_kmers / vals_sum / vals_count add up, vals_min / vals_max are min /
    max, a record is listed if any rank counted one of its k-mers).  Rank 0 writes the JSON."""
    import torch
    import torch.distributed as tdist
    from . import device as dev, dist as pdist
    from .tools import gen_checksum

    rank, world = pdist.world()
    if header.kmer_len >= 19:
        # very sparse table, counted DIRECT: balance updates + zero-fill analytically (canonical k-mers
        # crowd the low values).  The k-mer count is not known before the file is read; every rank
        # derives the same estimate from the size of the input (about 3.5 bases per compressed byte).
        size = os.path.getsize(header.input_file_path)
        packed = header.input_file_path.endswith((".gz", ".bgz"))
        lo, hi = pdist.analytic_kmer_ranges(header.data_size, world, size * (3.5 if packed else 1.0))[rank]
    else:
        lo, hi = pdist.shard_range(header.data_size, rank, world)
    t_start = time.perf_counter()
    error = None
    if rank == 0:
        try:
            header.init_index_tmp_file(overwrite=overwrite)
        except (ValueError, OSError) as exc:
            error = exc
    pdist.raise_together(error)

    from concurrent.futures import ThreadPoolExecutor
    helpers = ThreadPoolExecutor(max_workers=2)
    input_sum = helpers.submit(gen_checksum, header.input_file_path) if rank == 0 else None

    fs, pieces = None, None
    if rank == 0:
        fs = FastaStream(input_file, chunk_bytes=chunk_bytes)
        cap = chunk_bytes + (chunk_bytes >> 3) + (1 << 17)
        ring = [dev.pinned_empty(cap) for _ in range(2)]
        pieces = fs.pieces(buffers=[r.numpy() for r in ring])
    nrec = 0
    with dev.device_scope(device):
        ix = dev.Indexer(header.kmer_len, device=device, range_lo=lo, range_hi=hi) if hi > lo else None
        while True:
            note, piece, error = [None], None, None
            if rank == 0:
                try:
                    piece = next(pieces, None)
                except (ValueError, OSError) as exc:       # text the reader rejects, damaged .bgz
                    error = exc
                if piece is not None:
                    note = [(int(piece.size), [int(v) for v in fs.starts])]
            pdist.raise_together(error)
            tdist.broadcast_object_list(note, src=0)
            if note[0] is None:
                break
            nbytes, starts = note[0]
            nrec = len(starts)
            if nbytes == 0:
                continue
            chunk = None
            if rank == 0:
                chunk = torch.from_numpy(piece)
                if tdist.get_backend() == "nccl":
                    chunk = dev.upload(chunk)
            chunk = pdist.broadcast_stream(chunk, nbytes, src=0)
            if ix is not None:
                d_chunk = dev.upload(chunk)
                ix.set_records(starts)
                ix.feed_device(d_chunk)
                ix.sync()                                  # the piece may be dropped from here on
                del d_chunk
            if rank == 0:
                header.timer.update(sum(fs.lengths))

        if ix is not None:
            hist, st = ix.finalize()
            flags = ix.record_flags() if nrec else np.zeros(0, dtype=np.uint8)
        else:                                              # tiny table, more ranks than slices
            hist = [0] * 255
            st = {"num_kmers": 0, "vals_sum": 0, "vals_count": 0, "vals_min": 255, "vals_max": 0}
            flags = np.zeros(nrec, dtype=np.uint8)
        hist, st = pdist.reduce_index_stats(hist, st)
        if nrec:
            flags = pdist.reduce_flags(flags)
        t_gpu = time.perf_counter()

        # every rank writes its own slice of the table at its own offset of the (sparse) tmp file
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
            ix.close()
    tdist.barrier()
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
    pdist.raise_together(error)
    t_end = time.perf_counter()
    header.wall_seconds = {"ingest_and_gpu": t_gpu - t_start, "write_table": t_write - t_gpu,
                           "metadata_and_sha256": t_end - t_write, "total": t_end - t_start}
    if rank == 0:
        print("  wall ({} ranks): ingest+GPU {ingest_and_gpu:.2f} s, table write {write_table:.2f} s, "
              "metadata + sha256 {metadata_and_sha256:.2f} s, total {total:.2f} s".format(world, **header.wall_seconds))
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

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

    header = Header(project_name, sample_name=sample_name, input_file=input_file,
                    kmer_len=kmer_len, flush_every=flush_every, min_frag_size=min_frag_size,
                    max_frag_size=max_frag_size, buffer_size=buffer_size)
    print(f"project_name {header.project_name} sample_name {header.sample_name} "
          f"kmer_len {header.kmer_len:15,d} kmer_size {header.kmer_size:15,d}")
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
    create_fasta_index(project_name, sample_name, input_file, kmer_len, buffer_size=2 ** 16,
                       overwrite=True, debug=False)
    print()


if __name__ == "__main__":
    main()

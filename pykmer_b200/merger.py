"""merger: N .kin tables -> (N, N, 3) matrix of (Total #1, Total #2, Shared) (.kma + .kma.json).

Host-side mirror of the reference's merger.py (argparse merger.py:51-59,
calculate_distance :62-78, merge :80-210, main :213-239): same command line,
validation, output names and file contents.  Where the reference scans every
pair of files (N(N-1)/2 tasks in a multiprocessing.Pool), this module reads each
table once, thresholds and bit-packs it on the GPU, and obtains the whole matrix
from one Gram contraction G = B * B^T (pykmer_b200/csrc/merger.cu, gram_f4.cu, gram_i8.cu):
matrix[k, l] = (G[k,k], G[l,l], G[k,l]).

The reference leaves the diagonal of the matrix uninitialised (merger.py:136);
here it is (T_k, T_k, T_k).
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .tools import Header

KIN, PACKED = "." + Header.IND_EXT, "." + Header.COMP_EXT
EXTS = (KIN, KIN + PACKED, ".kma", ".kma" + PACKED)                 # merger.py:38-43

DEFAULT_MIN_COUNT = Header.DEFAULT_MIN_COUNT
DEFAULT_MAX_COUNT = Header.DEFAULT_MAX_COUNT
DEFAULT_BUFFER_SIZE = Header.DEFAULT_BUFFER_SIZE
DEFAULT_BLOCK_SIZE = Header.DEFAULT_BLOCK_SIZE
DEFAULT_THREADS = 4

# option, default, help text (the reference's CLI, merger.py:51-59)
_OPTIONS = (
    ("--min-count", DEFAULT_MIN_COUNT, "Minimum Kmer Count"),
    ("--max-count", DEFAULT_MAX_COUNT, "Maximum Kmer Count"),
    ("--buffer-size", DEFAULT_BUFFER_SIZE, "Buffer size"),
    ("--block-size", DEFAULT_BLOCK_SIZE, "Block size"),
    ("--threads", DEFAULT_THREADS, "Threads (kept for compatibility: the GPU path reads every table once)"),
)


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="Merge kmer databases.")
    parser.add_argument("Project_Name", metavar="P", type=str, help="Project name")
    for dest, nargs in (("Kmer_1", 1), ("Kmer_N", "+")):
        parser.add_argument(dest, metavar="K", type=Path, nargs=nargs, help="List of kin files")
    for flag, default, text in _OPTIONS:
        parser.add_argument(flag, type=int, default=default, nargs="?", help=f"{text} [{default}]")
    return parser


def calculate_distance(k_index_file, l_index_file, min_count: int = DEFAULT_MIN_COUNT,
                       max_count: int = DEFAULT_MAX_COUNT, buffer_size: int = DEFAULT_BUFFER_SIZE,
                       block_size: int = DEFAULT_BLOCK_SIZE) -> Tuple[int, int, int]:
    """One pair of index files -> (Total k, Total l, Shared), as the Pool worker of the
    reference returns it (merger.py:62-78)."""
    pair = [Header(str(f), index_file=str(f), buffer_size=buffer_size) for f in (k_index_file, l_index_file)]
    return pair[0].calculate_distance(pair[1], min_count=min_count, max_count=max_count,
                                      block_size=block_size, threading=True)


def _description_file(kin: Path) -> Path:
    """<x>.kin[.bgz] -> <x>.kin.json (merger.py:112-113)."""
    name = str(kin)
    if name.endswith(PACKED):
        name = name[:-len(PACKED)]
    return Path(f"{name}.{Header.DESC_EXT}")


def _check_inputs(indexes: Sequence[Path], buffer_size: int) -> List[Dict]:
    """The reference's per-input checks (merger.py:104-129): extension, sibling JSON, equal K."""
    entries: List[Dict] = []
    for pos, kin in enumerate(indexes):
        print(f"verifying {kin}")
        assert str(kin).endswith(EXTS), f"all files must be .{Header.IND_EXT}[.bgz]: {kin}"
        desc = _description_file(kin)
        assert desc.exists(), (f"all .{Header.IND_EXT}[.{Header.COMP_EXT}] files must have a associated "
                               f".{Header.IND_EXT}.{Header.DESC_EXT}: {desc}")
        header = Header(str(kin), index_file=str(kin), buffer_size=buffer_size)
        if entries:
            want = entries[0]["header"].kmer_len
            assert header.kmer_len == want, f"kmer_length differs. expected {want}, got {header.kmer_len}"
        entries.append({"pos": pos, "index_file": str(kin), "description_file": str(desc), "header": header})
    return entries


def _save(outfile: Path, project_name: str, min_count: int, max_count: int, entries: List[Dict],
          matrix: np.ndarray) -> None:
    """.kma.json then .kma, each through a .tmp + rename (merger.py:190-208)."""
    for e in entries:
        e["header"] = e["header"].to_dict(lean=True)
    doc = {"project_name": project_name, "min_count": min_count, "max_count": max_count, "data": entries}
    json_path = Path(f"{outfile}.json")
    print(f"saving {json_path}")
    tmp = Path(f"{json_path}.tmp")
    with tmp.open(mode="wt") as fh:
        json.dump(doc, fh, sort_keys=True, indent=1)
    tmp.rename(json_path)
    print(f"saving {outfile}")
    tmp = Path(f"{outfile}.tmp")
    with tmp.open(mode="wb") as fh:
        np.savez_compressed(fh, matrix=matrix)
    tmp.rename(outfile)


def merge(project_name: str, indexes: List[Path], min_count: int = DEFAULT_MIN_COUNT,
          max_count: int = DEFAULT_MAX_COUNT, buffer_size: int = DEFAULT_BUFFER_SIZE,
          block_size: int = DEFAULT_BLOCK_SIZE, threads: int = DEFAULT_THREADS,
          device: int = 0):
    """merge (merger.py:80-210): validate, compute the matrix on the GPU, write the two files."""
    assert min_count >= 1                                     # merger.py:90-94
    assert max_count <= 255
    assert buffer_size > 0
    assert block_size > 0
    assert len(indexes) > 0

    outfile = Path(f"{project_name}.{min_count:03d}-{max_count:03d}.kma")
    assert not Path(project_name).exists(), \
        f"project name ({project_name}) is a file. maybe forgot to pass project name as first argument?"
    assert not outfile.exists(), f"project output file ({outfile}) already exists. not overwriting."

    indexes = [Path(p) for p in indexes]
    assert all(i.exists() for i in indexes)
    entries = _check_inputs(indexes, buffer_size)

    from . import dist as pdist
    rank, world = pdist.world()
    matrix = merge_tables([e["header"] for e in entries], min_count, max_count, device=device)
    if rank == 0:                                             # every rank holds the matrix; one writes
        n = len(entries)
        for k, l in ((k, l) for k in range(n - 1) for l in range(k + 1, n)):
            t_k, t_l, shared = (int(v) for v in matrix[k, l])
            print(f"   matrix Total #{k:3d} {t_k:15,d} Total #{l:3d} {t_l:15,d} Shared {shared:15,d}")
        _save(outfile, project_name, min_count, max_count, entries, matrix)
    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
    return entries, matrix


class _SliceSource:
    """Entries [lo, hi) of one sample's table (the file its JSON names, .bgz preferred:
    tools.py:185-196), handed out slab by slab into the caller's buffers."""

    def __init__(self, header: Header, lo: int, hi: int):
        path = header.index_file
        self.left = hi - lo
        self.arr, self.fh, self.pos = None, None, 0
        if path.endswith(PACKED):
            whole = lo == 0 and hi == header.data_size
            self.arr = header.read_table(path) if whole else header.read_table_slice(lo, hi, path)
        else:
            import os
            assert os.path.getsize(path) == header.data_size, \
                f"{path}: {os.path.getsize(path)} bytes, expected {header.data_size}"
            self.fh = open(path, "rb", buffering=0)
            self.fh.seek(lo)

    def read_into(self, buf: np.ndarray) -> int:
        n = min(buf.size, self.left)
        if self.arr is not None:
            buf[:n] = self.arr[self.pos:self.pos + n]
        else:
            view, got = memoryview(buf)[:n], 0
            while got < n:
                step = self.fh.readinto(view[got:])
                assert step, "table file ended early"
                got += step
        self.pos += n
        self.left -= n
        return n

    def close(self) -> None:
        if self.fh is not None:
            self.fh.close()
        self.arr = None


def merge_tables(headers: List[Header], min_count: int, max_count: int, device: int = 0,
                 slab_bytes: int = 256 << 20, mask_budget_bytes: Optional[int] = None) -> np.ndarray:
    """Read each sample's table once (the file named by its JSON, .bgz preferred:
    tools.py:185-196), threshold + pack it on the GPU, and contract: G = B * B^T.
    -> (N, N, 3) uint64.

    The presence masks of ALL samples for one stretch of the k-mer axis must be resident for the
    contraction, N / 8 bytes per k-mer.  When the whole axis does not fit (K=17 x 255 samples is
    548 GB) the axis is cut into chunks that do and the Gram matrices of the chunks are added up
    (pk_gram_tiled_device(accumulate=1)) -- the sums of tools.py:480-482 are sums over that axis,
    exactly like the reference's own 100 M-entry blocks (tools.py:449-489).  mask_budget_bytes
    (default: 40 % of the free device memory) sets the chunk; the tests force small ones.

    Inside a torch.distributed job (torchrun merger.py ...) rank r takes slice r of the k-mer
    axis of every sample, for the same reason (one all-reduce of N x N int64), and
    reads only that slice of each file (Header.read_table_slice)."""
    import torch
    from . import device as dev          # needs the CUDA library; no fallback
    from . import dist as pdist

    N = len(headers)
    T = headers[0].data_size
    for h in headers:
        assert h.data_size == T
    rank, world = pdist.world()
    lo, hi = pdist.shard_range(T, rank, world) if world > 1 else (0, T)
    n_own = hi - lo
    with dev.device_scope(device):
        tiled = dev.use_tiled_masks(N, device)     # tiled masks + the FP4 tensor-core Gram kernel
        G = dev.zeros((N, N), torch.int64)
        if mask_budget_bytes is None:
            mask_budget_bytes = int(dev.free_memory_bytes() * 0.4)
        # k-mers per chunk: whole tiles of 1024 (the tiled layout), at least one
        per_chunk = max(1024, (mask_budget_bytes * 8 // max(N, 1)) // 1024 * 1024)
        chunks = [(c, min(hi, c + per_chunk)) for c in range(lo, hi, per_chunk)]
        c_words = (min(per_chunk, n_own) + 31) // 32
        stride = max(4, (c_words + 3) & ~3)
        bits = None
        if chunks:
            bits = dev.tiled_masks(c_words, N) if tiled else dev.zeros((N, stride), torch.int32)
        # Two pinned slabs in turn: a helper thread reads / inflates the next slab straight into one
        # while the other crosses PCIe.  A raw .kin is read at its offset with readinto (no
        # intermediate array); a .kin.bgz has to be inflated first (the slice only, Header.read_table_slice).
        ring = [dev.pinned_empty(max(1, min(n_own, per_chunk, slab_bytes))) for _ in range(2)]
        jobs = [(ci, s, off) for ci, (c0, c1) in enumerate(chunks) for s in range(N)
                for off in range(0, c1 - c0, slab_bytes)]
        source: Dict[int, _SliceSource] = {}

        def read(j: int) -> int:
            ci, s, off = jobs[j]
            c0, c1 = chunks[ci]
            if off == 0:
                source[s] = _SliceSource(headers[s], c0, c1)
            n = source[s].read_into(ring[j & 1].numpy())
            if off + n >= c1 - c0:
                source.pop(s).close()
            return n

        from concurrent.futures import ThreadPoolExecutor
        reader = ThreadPoolExecutor(max_workers=1)
        ahead = reader.submit(read, 0) if jobs else None
        for j, (ci, s, off) in enumerate(jobs):
            c0, c1 = chunks[ci]
            n = ahead.result()
            assert n == min(slab_bytes, c1 - c0 - off)
            if j + 1 < len(jobs):
                dev.stream_sync()                              # the other slab has crossed: refill it
                ahead = reader.submit(read, j + 1)
            if ci and s == 0 and off == 0:
                bits.zero_()                                   # a shorter last chunk must not see stale words
            d = dev.upload(ring[j & 1][:n], non_blocking=True)
            if tiled:
                dev.threshold_pack_tiled(d, min_count, max_count, bits, s, N, first_word=off // 32)
            else:
                dev.threshold_pack(d, min_count, max_count, out=bits[s, off // 32:])
            if s == N - 1 and off + n >= c1 - c0:              # the chunk is complete: contract it
                w = (c1 - c0 + 31) // 32
                if tiled:
                    dev.gram_tiled(bits, N, w, out=G, accumulate=True)
                else:
                    dev.gram(bits, words=w, out=G, accumulate=True)
        dev.stream_sync()
        reader.shutdown()
        pdist.reduce_gram(G)
        Gh = G.cpu().numpy()
    return dev.matrix_from_gram(Gh)


def main(argv: Optional[List[str]] = None) -> None:
    args = build_parser().parse_args(argv)
    indexes: List[Path] = sorted(args.Kmer_1 + args.Kmer_N)   # merger.py:228
    if len(indexes) <= 1:
        print("needs at least 2 files")
        sys.exit(1)
    import os
    device = 0
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:            # torchrun merger.py ...: one rank per GPU
        from . import dist as pdist
        device = pdist.init_from_env()[2]
    merge(args.Project_Name, indexes, min_count=args.min_count, max_count=args.max_count,
          buffer_size=args.buffer_size, block_size=args.block_size, threads=args.threads, device=device)


if __name__ == "__main__":
    main()

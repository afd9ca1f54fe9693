"""merger: N .kin tables -> (N, N, 3) matrix of (Total #1, Total #2, Shared) (.kma + .kma.json).

Host-side mirror of the reference's merger.py (argparse merger.py:51-59,
calculate_distance :62-78, merge :80-210, main :213-239): same arguments,
validation, output names and file contents.  The reference scans every pair of
files (N(N-1)/2 tasks in a multiprocessing.Pool); here every table is read
once, thresholded and bit-packed on the GPU, and the whole matrix is one Gram
contraction G = B * B^T (pykmer_b200/csrc/merger.cu):
matrix[k, l] = (G[k,k], G[l,l], G[k,l]).

The reference leaves the diagonal of the matrix uninitialised (merger.py:136);
here it is (T_k, T_k, T_k).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np

from .tools import Header

EXTS = ("." + Header.IND_EXT, "." + Header.IND_EXT + "." + Header.COMP_EXT,
        ".kma", ".kma." + Header.COMP_EXT)

DEFAULT_MIN_COUNT = Header.DEFAULT_MIN_COUNT
DEFAULT_MAX_COUNT = Header.DEFAULT_MAX_COUNT
DEFAULT_BUFFER_SIZE = Header.DEFAULT_BUFFER_SIZE
DEFAULT_BLOCK_SIZE = Header.DEFAULT_BLOCK_SIZE
DEFAULT_THREADS = 4


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Merge kmer databases.")
    p.add_argument("Project_Name", metavar="P", type=str, help="Project name")
    p.add_argument("Kmer_1", metavar="K", type=Path, nargs=1, help="List of kin files")
    p.add_argument("Kmer_N", metavar="K", type=Path, nargs="+", help="List of kin files")
    p.add_argument("--min-count", type=int, default=DEFAULT_MIN_COUNT, nargs="?",
                   help=f"Minimum Kmer Count [{DEFAULT_MIN_COUNT}]")
    p.add_argument("--max-count", type=int, default=DEFAULT_MAX_COUNT, nargs="?",
                   help=f"Maximum Kmer Count [{DEFAULT_MAX_COUNT}]")
    p.add_argument("--buffer-size", type=int, default=DEFAULT_BUFFER_SIZE, nargs="?",
                   help=f"Buffer size [{DEFAULT_BUFFER_SIZE}]")
    p.add_argument("--block-size", type=int, default=DEFAULT_BLOCK_SIZE, nargs="?",
                   help=f"Block size [{DEFAULT_BLOCK_SIZE}]")
    p.add_argument("--threads", type=int, default=DEFAULT_THREADS, nargs="?",
                   help=f"Threads [{DEFAULT_THREADS}] (accepted for compatibility; the GPU path "
                        "reads each table once)")
    return p


def calculate_distance(k_index_file, l_index_file, min_count: int = DEFAULT_MIN_COUNT,
                       max_count: int = DEFAULT_MAX_COUNT, buffer_size: int = DEFAULT_BUFFER_SIZE,
                       block_size: int = DEFAULT_BLOCK_SIZE) -> Tuple[int, int, int]:
    """One pair (merger.py:62-78) -> (Total k, Total l, Shared)."""
    k_header = Header(str(k_index_file), index_file=str(k_index_file), buffer_size=buffer_size)
    l_header = Header(str(l_index_file), index_file=str(l_index_file), buffer_size=buffer_size)
    return k_header.calculate_distance(l_header, min_count=min_count, max_count=max_count,
                                       block_size=block_size, threading=True)


def merge(project_name: str, indexes: List[Path], min_count: int = DEFAULT_MIN_COUNT,
          max_count: int = DEFAULT_MAX_COUNT, buffer_size: int = DEFAULT_BUFFER_SIZE,
          block_size: int = DEFAULT_BLOCK_SIZE, threads: int = DEFAULT_THREADS,
          device: int = 0):
    """merge (merger.py:80-210): validate, compute the matrix, write .kma.json and .kma."""
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

    data = []
    kmer_len = None
    for k, kin in enumerate(indexes):
        print(f"verifying {kin}")
        kins = str(kin)
        assert kins.endswith(EXTS), f"all files must be .{Header.IND_EXT}[.bgz]: {kin}"
        packed = "." + Header.COMP_EXT
        desc = Path(f"{kins[:-len(packed)] if kins.endswith(packed) else kins}.{Header.DESC_EXT}")
        assert desc.exists(), (f"all .{Header.IND_EXT}[.{Header.COMP_EXT}] files must have a "
                               f"associated .{Header.IND_EXT}.{Header.DESC_EXT}: {desc}")
        header = Header(kins, index_file=kins, buffer_size=buffer_size)
        if kmer_len is None:
            kmer_len = header.kmer_len
        assert header.kmer_len == kmer_len, \
            f"kmer_length differs. expected {kmer_len}, got {header.kmer_len}"
        data.append({"pos": k, "index_file": kins, "description_file": str(desc), "header": header})

    matrix = merge_tables([d["header"] for d in data], min_count, max_count, device=device)

    for k in range(len(data) - 1):
        for l in range(k + 1, len(data)):
            print(f"   matrix Total #{k:3d} {int(matrix[k, l, 0]):15,d} Total #{l:3d} "
                  f"{int(matrix[k, l, 1]):15,d} Shared {int(matrix[k, l, 2]):15,d}")

    for v in data:
        v["header"] = v["header"].to_dict(lean=True)
    output = {"project_name": project_name, "min_count": min_count, "max_count": max_count,
              "data": data}

    outfile_json = Path(f"{outfile}.json")
    outfile_json_tmp = Path(f"{outfile_json}.tmp")
    print(f"saving {outfile_json}")
    with outfile_json_tmp.open(mode="wt") as fh:
        json.dump(output, fh, sort_keys=True, indent=1)
    outfile_json_tmp.rename(outfile_json)

    print(f"saving {outfile}")
    outfile_tmp = Path(f"{outfile}.tmp")
    with outfile_tmp.open(mode="wb") as fh:
        np.savez_compressed(fh, matrix=matrix)
    outfile_tmp.rename(outfile)
    return data, matrix


def merge_tables(headers: List[Header], min_count: int, max_count: int, device: int = 0,
                 slab_bytes: int = 256 << 20) -> np.ndarray:
    """Read each sample's table once (the file named by its JSON, .bgz preferred:
    tools.py:185-196), threshold + pack it on the GPU, then one Gram pass.
    -> (N, N, 3) uint64."""
    import torch
    from . import device as dev          # needs the CUDA library; no fallback

    N = len(headers)
    T = headers[0].data_size
    words = (T + 31) // 32
    stride = (words + 3) & ~3
    with torch.cuda.device(device):
        bits = torch.zeros((N, stride), dtype=torch.int32, device="cuda")
        stage = dev.pinned_empty(min(T, slab_bytes))
        for s, h in enumerate(headers):
            table = h.read_table()
            assert table.size == T
            for off in range(0, T, slab_bytes):
                n = min(slab_bytes, T - off)
                torch.cuda.current_stream().synchronize()      # stage is reused
                stage.numpy()[:n] = table[off:off + n]
                d = stage[:n].to("cuda", non_blocking=True)
                dev.threshold_pack(d, min_count, max_count, out=bits[s, off // 32:])
        G = dev.gram(bits, words=words)
        Gh = G.cpu().numpy()
    return dev.matrix_from_gram(Gh)


def main(argv: Optional[List[str]] = None) -> None:
    args = build_parser().parse_args(argv)
    indexes: List[Path] = args.Kmer_1 + args.Kmer_N
    if len(indexes) <= 1:
        print("needs at least 2 files")
        sys.exit(1)
    indexes.sort()                                             # merger.py:228
    merge(args.Project_Name, indexes, min_count=args.min_count, max_count=args.max_count,
          buffer_size=args.buffer_size, block_size=args.block_size, threads=args.threads)


if __name__ == "__main__":
    main()

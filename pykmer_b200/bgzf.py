"""BGZF (blocked gzip, what `bgzip` writes) for the .kin.bgz side of the workflow.

The reference's documented pipeline compresses every .kin with the external `bgzip -l 9`
(README.md:26, data/README.md:24) and its merger gunzips .kin.bgz inputs with Python's gzip
(tools.py:296-302).  The container is a chain of independent <= 64 KiB gzip members with a 'BC'
extra field, so both directions parallelise over host threads (zlib releases the GIL):

    python -m pykmer_b200.bgzf genome.fa.15.kin          # -> genome.fa.15.kin.bgz
    python -m pykmer_b200.bgzf -d genome.fa.15.kin.bgz   # -> genome.fa.15.kin
"""
from __future__ import annotations

import os
import struct
import sys
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import Optional

from .fasta import bgzf_chunks, is_bgzf

BLOCK = 0xFF00
EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_HEAD = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00"


def _deflate_block(args) -> bytes:
    chunk, level = args
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    payload = co.compress(chunk) + co.flush()
    return b"".join((_HEAD, struct.pack("<H", len(payload) + 25), payload,
                     struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))))


def compress_file(src: str, dst: Optional[str] = None, level: int = 6, threads: Optional[int] = None,
                  batch: int = 1024) -> str:
    dst = dst or src + ".bgz"
    threads = threads or min(32, os.cpu_count() or 1)
    tmp = dst + ".tmp"
    with open(src, "rb") as fi, open(tmp, "wb") as fo, ThreadPoolExecutor(max_workers=threads) as pool:
        while True:
            data = fi.read(BLOCK * batch)
            if not data:
                break
            mv = memoryview(data)
            jobs = [(mv[o:o + BLOCK], level) for o in range(0, len(mv), BLOCK)]
            for blk in pool.map(_deflate_block, jobs):
                fo.write(blk)
        fo.write(EOF_BLOCK)
    os.replace(tmp, dst)
    return dst


def decompress_file(src: str, dst: Optional[str] = None) -> str:
    assert is_bgzf(src), f"{src} is not a BGZF file"
    dst = dst or (src[:-4] if src.endswith(".bgz") else src + ".out")
    tmp = dst + ".tmp"
    with open(tmp, "wb") as fo:
        for chunk in bgzf_chunks(src):
            fo.write(chunk)
    os.replace(tmp, dst)
    return dst


def read_all(path: str) -> bytes:
    return b"".join(bgzf_chunks(path))


def main(argv=None) -> None:
    argv = sys.argv[1:] if argv is None else argv
    if argv and argv[0] == "-d":
        for p in argv[1:]:
            print(decompress_file(p))
    else:
        level = 6
        if argv and argv[0] == "-l":
            level, argv = int(argv[1]), argv[2:]
        for p in argv:
            print(compress_file(p, level=level))


if __name__ == "__main__":
    main()

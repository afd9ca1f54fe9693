"""BGZF (blocked gzip, what `bgzip` writes) for the .kin.bgz side of the workflow.

The reference's documented pipeline compresses every .kin with the external `bgzip -l 9`
(README.md:26,261-269; data/README.md:24), its merger gunzips .kin.bgz inputs with Python's gzip
(tools.py:296-302), and gzireader.py prints the `.gzi` block index `bgzip -i` leaves beside them
(gzireader.py:12-37).  The container is a chain of independent <= 64 KiB gzip members with a 'BC'
extra field, so both directions parallelise over host cores: pk_bgzf_deflate / pk_bgzf_inflate
in libpykmer_b200.so (csrc/ingest.cpp) when the library is built, zlib on a thread pool otherwise
(host utilities -- no device arithmetic is involved either way).

    python -m pykmer_b200.bgzf [-l 9] [-i] genome.fa.15.kin   # -> genome.fa.15.kin.bgz [+ .gzi]
    python -m pykmer_b200.bgzf -d genome.fa.15.kin.bgz        # -> genome.fa.15.kin
    python -m pykmer_b200.bgzf -r genome.fa.15.kin.bgz        # (re)build the .gzi of an existing file

With the block index a slice [lo, hi) of the table is read by inflating only the members that
hold it (`read_range`) -- which is how each rank of a multi-GPU merge takes its share of the
k-mer axis out of every sample without inflating the other seven eighths.
"""
from __future__ import annotations

import ctypes
import os
import struct
import sys
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .fasta import _bgzf_block_size, _inflate_block, _native_lib, bgzf_chunks, is_bgzf

BLOCK = 0xFF00
SLOT = 1 << 16
EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_HEAD = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00"
GZI_EXT = ".gzi"


def _deflate_block(args) -> bytes:
    chunk, level = args
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    payload = co.compress(chunk) + co.flush()
    return b"".join((_HEAD, struct.pack("<H", len(payload) + 25), payload,
                     struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk))))


class _Deflater:
    """One batch of input -> (BGZF members back to back, their compressed sizes)."""

    def __init__(self, level: int, threads: int, batch: int, native: Optional[bool]):
        self.level, self.threads = level, threads
        nat = _native_lib()
        self.nat = nat if (native is None or native) else None
        if native and nat is None:
            raise ImportError("pykmer_b200.bgzf: native deflate asked for, libpykmer_b200.so is missing")
        if self.nat is not None:
            self.out = np.empty(batch * SLOT, dtype=np.uint8)
            self.sizes = np.empty(batch, dtype=np.uint32)
        else:
            self.pool = ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 1))

    def run(self, data: bytes) -> Tuple[memoryview, Sequence[int]]:
        nblk = (len(data) + BLOCK - 1) // BLOCK
        if self.nat is not None:
            src = np.frombuffer(data, dtype=np.uint8)
            made = ctypes.c_size_t(0)
            self.nat.check(self.nat.lib.pk_bgzf_deflate(src.ctypes.data, src.size, self.out.ctypes.data,
                                                        self.out.size, ctypes.byref(made),
                                                        self.sizes.ctypes.data, self.level, self.threads))
            return memoryview(self.out)[:made.value], self.sizes[:nblk].tolist()
        mv = memoryview(data)
        blocks = list(self.pool.map(_deflate_block, [(mv[o:o + BLOCK], self.level) for o in range(0, len(mv), BLOCK)]))
        return memoryview(b"".join(blocks)), [len(b) for b in blocks]

    def close(self) -> None:
        if self.nat is None:
            self.pool.shutdown()


def compress_file(src: str, dst: Optional[str] = None, level: int = 6, threads: Optional[int] = None,
                  batch: int = 1024, index: bool = False, native: Optional[bool] = None) -> str:
    """src -> dst (default src + '.bgz'), written under a .tmp name and renamed.  index=True also
    writes dst + '.gzi' (`bgzip -i`)."""
    dst = dst or src + ".bgz"
    tmp = dst + ".tmp"
    worker = _Deflater(level, threads or 0, batch, native)
    comp_off, raw_off, entries = 0, 0, []
    try:
        with open(src, "rb") as fi, open(tmp, "wb") as fo:
            while True:
                data = fi.read(BLOCK * batch)
                if not data:
                    break
                blob, sizes = worker.run(data)
                fo.write(blob)
                for i, size in enumerate(sizes):
                    entries.append((comp_off, raw_off))
                    comp_off += size
                    raw_off += min(BLOCK, len(data) - i * BLOCK)
            fo.write(EOF_BLOCK)
    finally:
        worker.close()
    os.replace(tmp, dst)
    if index:
        write_index(dst + GZI_EXT, entries)
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


# ------------------------------------------------------------------------------- .gzi index
# Layout (gzireader.py:12-19, htslib): uint64 number_entries, then number_entries pairs of
# uint64 (compressed_offset, uncompressed_offset), little endian -- the start of every member
# but the first, whose (0, 0) is implied; the empty EOF member has no entry.

def write_index(index_file: str, entries: Sequence[Tuple[int, int]]) -> str:
    """entries = (compressed_offset, uncompressed_offset) of EVERY data member, first included."""
    body = [e for e in entries if e != (0, 0)]
    with open(index_file + ".tmp", "wb") as fh:
        fh.write(struct.pack("<Q", len(body)))
        fh.write(np.asarray(body, dtype="<u8").reshape(-1, 2).tobytes())
    os.replace(index_file + ".tmp", index_file)
    return index_file


def read_index(index_file: str) -> List[Tuple[int, int]]:
    """-> (compressed_offset, uncompressed_offset) of every data member, the implied first included."""
    with open(index_file, "rb") as fh:
        blob = fh.read()
    if len(blob) < 8:
        raise OSError(f"{index_file}: too short for a .gzi index")
    (count,) = struct.unpack_from("<Q", blob)
    if len(blob) != 8 + 16 * count:
        raise OSError(f"{index_file}: {count} entries announced, {len(blob) - 8} bytes of entries")
    pairs = np.frombuffer(blob, dtype="<u8", offset=8).reshape(-1, 2)
    return [(0, 0)] + [(int(c), int(u)) for c, u in pairs]


def scan_index(path: str) -> List[Tuple[int, int]]:
    """The same list from the member headers of the file itself (`bgzip -r`).  Walks the members with
    seeks and two small reads each (header: BSIZE; trailer: ISIZE) -- the compressed payload, hundreds
    of MB for a K >= 15 table, is never read, so a rank that wants 1/N of a table does not pay for all
    of it."""
    entries, comp_off, raw_off = [], 0, 0
    total = os.path.getsize(path)
    with open(path, "rb", buffering=0) as fh:
        while comp_off < total:
            fh.seek(comp_off)
            head = fh.read(18)
            size = _bgzf_block_size(head, 0)
            if size == 0 and len(head) == 18:                    # a longer extra field than the usual 6 bytes
                xlen = head[10] | (head[11] << 8)
                fh.seek(comp_off)
                size = _bgzf_block_size(fh.read(12 + xlen), 0)
            if size <= 0 or comp_off + size > total:
                raise OSError(f"{path}: not a whole BGZF block at offset {comp_off}")
            fh.seek(comp_off + size - 4)
            isize = int.from_bytes(fh.read(4), "little")
            if isize:
                entries.append((comp_off, raw_off))
            comp_off += size
            raw_off += isize
    return entries


def build_index(path: str) -> str:
    return write_index(path + GZI_EXT, scan_index(path))


def load_index(path: str) -> List[Tuple[int, int]]:
    """The block index of a BGZF file: its .gzi if there is one, else from the headers."""
    gzi = path + GZI_EXT
    if os.path.exists(gzi) and os.path.getmtime(gzi) >= os.path.getmtime(path):
        return read_index(gzi)
    return scan_index(path)


def print_index(index_file: str) -> None:
    """gzireader.py:21-37: the entries of <file>.gzi and the size of <file>."""
    entries = read_index(index_file)[1:]
    filesize = os.path.getsize(index_file[:-len(GZI_EXT)])
    print(f"number_entries: {len(entries):15,d}")
    print(f"filesize      : {filesize:15,d}")
    for pos, (comp, raw) in enumerate(entries):
        print(f"pos: {pos:15,d} compressed_offset {comp:15,d} uncompressed_offset {raw:15,d}")
    print(f"number_entries: {len(entries):15,d}")
    print(f"filesize      : {filesize:15,d}")


def read_range(path: str, lo: int, hi: int, index: Optional[Sequence[Tuple[int, int]]] = None,
               threads: int = 0) -> np.ndarray:
    """Bytes [lo, hi) of the decompressed contents, inflating only the members that hold them."""
    assert 0 <= lo <= hi
    if hi == lo:
        return np.empty(0, dtype=np.uint8)
    index = load_index(path) if index is None else index
    raws = np.fromiter((u for _, u in index), dtype=np.int64, count=len(index))
    first = int(np.searchsorted(raws, lo, side="right")) - 1
    last = int(np.searchsorted(raws, hi, side="left"))          # first member starting at or after hi
    comp_lo = index[first][0]
    raw_lo = index[first][1]
    with open(path, "rb") as fh:
        fh.seek(comp_lo)
        comp = fh.read(index[last][0] - comp_lo) if last < len(index) else fh.read()
    nat = _native_lib()
    if nat is not None:
        cap = ((index[last][1] if last < len(index) else raw_lo + (last - first) * SLOT) - raw_lo)
        out = np.empty(cap, dtype=np.uint8)
        used, made = ctypes.c_size_t(0), ctypes.c_size_t(0)
        cbuf = np.frombuffer(comp, dtype=np.uint8)
        try:
            nat.check(nat.lib.pk_bgzf_inflate(cbuf.ctypes.data, cbuf.size, out.ctypes.data, out.size,
                                              ctypes.byref(used), ctypes.byref(made), threads))
        except ValueError as exc:
            raise OSError(f"{path}: {exc}") from None
        out = out[:made.value]
    else:
        mv, pos, blocks = memoryview(comp), 0, []
        while pos < len(mv):
            size = _bgzf_block_size(mv, pos)
            if size <= 0 or pos + size > len(mv):
                raise OSError(f"{path}: not a whole BGZF block at offset {comp_lo + pos}")
            blocks.append(mv[pos:pos + size])
            pos += size
        with ThreadPoolExecutor(max_workers=threads or min(16, os.cpu_count() or 1)) as pool:
            out = np.frombuffer(b"".join(pool.map(_inflate_block, blocks)), dtype=np.uint8)
    if raw_lo + out.size < hi:
        raise OSError(f"{path}: holds {raw_lo + out.size} bytes, [{lo}, {hi}) asked for")
    return out[lo - raw_lo:hi - raw_lo]


def main(argv=None) -> None:
    argv = sys.argv[1:] if argv is None else list(argv)
    if argv and argv[0] == "-d":
        for p in argv[1:]:
            print(decompress_file(p))
        return
    if argv and argv[0] == "-r":
        for p in argv[1:]:
            print(build_index(p))
        return
    level, index = 6, False
    while argv and argv[0] in ("-l", "-i"):
        if argv[0] == "-l":
            level, argv = int(argv[1]), argv[2:]
        else:
            index, argv = True, argv[1:]
    for p in argv:
        print(compress_file(p, level=level, index=index))


if __name__ == "__main__":
    main()

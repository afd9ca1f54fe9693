"""Host ingest: FASTA text -> the cleaned byte stream the CUDA indexer eats.

Mirrors the text rules of the reference's reader byte for byte
(read_fasta indexer.py:101-128, parse_fasta indexer.py:45-99):
  * .gz / .bgz inputs are gunzipped (multi-member BGZF included), others read raw;
  * universal newlines: \\n, \\r\\n and a lone \\r all end a line;
  * every line is stripped of leading/trailing white space; blank lines vanish;
  * a stripped line starting with '>' opens a record named by the rest of the
    line (description kept); everything before the first header is dropped;
  * the other lines of a record are concatenated, so windows span line breaks;
  * every opened record is reported, empty ones included.
What leaves this module is the concatenation of the records' bytes with ONE
separator byte ('>', never a base) after each record, plus the record table
(name, stream offset, length).  Mapping characters to 2-bit codes, the
None-resets-the-window rule (indexer.py:144) and everything after it happens on
the GPU.

The reference decodes text and indexes CONV by code point (indexer.py:75-76), so a
non-ASCII byte inside sequence lines either fails to decode, raises IndexError or
shifts lengths; this reader rejects such input with ValueError instead of
guessing.  Header lines may hold any UTF-8.
"""
from __future__ import annotations

import gzip
import io
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import BinaryIO, Iterator, List, Optional, Tuple

import numpy as np

SEPARATOR = ord(">")
_STRIP = b"\t\n\x0b\x0c\r\x1c\x1d\x1e\x1f "          # str.strip() white space below 0x80
_NONNL_WS = tuple(bytes([c]) for c in (32, 9, 11, 12, 28, 29, 30, 31))   # strip set minus \n \r


def _has_inline_whitespace(buf: bytes) -> bool:
    """Any blank / tab / ... in the chunk?  (eight memchr passes; the common answer is no)"""
    return any(buf.find(c) >= 0 for c in _NONNL_WS)


def open_binary(path: str) -> BinaryIO:
    """read_fasta (indexer.py:108-115): gzip for .gz/.bgz, plain otherwise."""
    if path.endswith((".gz", ".bgz")):
        return gzip.open(path, "rb")
    return open(path, "rb")


# ---------------------------------------------------------------------------------- BGZF
# `bgzip` output is a chain of independent gzip members of at most 64 KiB, each carrying its
# compressed size in a 'BC' extra subfield.  The reference reads it with Python's gzip, one
# member after the other; the members are independent, so they are inflated here on a thread
# pool (zlib releases the GIL) -- same bytes out, several times faster.  Anything that is not
# BGZF (plain gzip, plain text) takes the ordinary path.

def _bgzf_block_size(buf, pos: int) -> int:
    """Total size of the BGZF block starting at pos, 0 if the header is not BGZF / incomplete."""
    if pos + 18 > len(buf):
        return 0
    if buf[pos] != 0x1F or buf[pos + 1] != 0x8B or buf[pos + 2] != 8 or not (buf[pos + 3] & 4):
        return -1
    xlen = buf[pos + 10] | (buf[pos + 11] << 8)
    p, end = pos + 12, pos + 12 + xlen
    if end > len(buf):
        return 0
    while p + 4 <= end:
        slen = buf[p + 2] | (buf[p + 3] << 8)
        if buf[p] == 66 and buf[p + 1] == 67 and slen == 2:            # 'B', 'C'
            return (buf[p + 4] | (buf[p + 5] << 8)) + 1
        p += 4 + slen
    return -1


def is_bgzf(path: str) -> bool:
    try:
        with open(path, "rb") as fh:
            head = fh.read(64)
    except OSError:
        return False
    return _bgzf_block_size(head, 0) > 0


def _inflate_block(block) -> bytes:
    xlen = block[10] | (block[11] << 8)
    data = zlib.decompress(block[12 + xlen:len(block) - 8], wbits=-15)
    crc = int.from_bytes(block[len(block) - 8:len(block) - 4], "little")
    isize = int.from_bytes(block[len(block) - 4:], "little")
    if len(data) != isize or (zlib.crc32(data) & 0xFFFFFFFF) != crc:
        raise OSError("BGZF block fails its CRC / length check")      # gzip raises BadGzipFile (an OSError)
    return data


def bgzf_chunks(path: str, chunk_bytes: int = 64 << 20, threads: Optional[int] = None) -> Iterator[bytes]:
    """Decompressed contents of a BGZF file, in order, in pieces of roughly chunk_bytes."""
    import os
    threads = threads or min(16, os.cpu_count() or 1)
    read_size = max(1 << 20, chunk_bytes // 3)
    with open(path, "rb") as fh, ThreadPoolExecutor(max_workers=threads) as pool:
        pending = b""
        eof = False
        while not eof or pending:
            if not eof:
                blk = fh.read(read_size)
                eof = not blk
                pending += blk
            mv = memoryview(pending)
            blocks, pos = [], 0
            while True:
                size = _bgzf_block_size(mv, pos)
                if size < 0:
                    raise OSError(f"{path}: not a BGZF block at offset {pos} of the current chunk")
                if size == 0 or pos + size > len(mv):
                    break
                blocks.append(mv[pos:pos + size])
                pos += size
            if not blocks and eof:
                if pending:
                    raise OSError(f"{path}: truncated BGZF block at end of file")
                break
            out = b"".join(pool.map(_inflate_block, blocks))
            pending = bytes(mv[pos:])
            if out:
                yield out


class FastaStream:
    """Incremental FASTA -> stream converter.

    for piece in fs.pieces(): ...   yields uint8 arrays (cleaned bases + separators)
    fs.names / fs.starts / fs.lengths describe the records seen so far; a
    record's entry exists as soon as its header has been read.
    """

    def __init__(self, path: str, chunk_bytes: int = 64 << 20):
        self.path = path
        self.chunk_bytes = chunk_bytes
        self.names: List[str] = []
        self.starts: List[int] = []
        self.lengths: List[int] = []
        self._pos = 0            # stream offset of the next byte to be emitted
        self._open = False       # a record is open (a header has been seen)

    # -- record bookkeeping ---------------------------------------------------------
    def _close_record(self, out: List[bytes]) -> None:
        if self._open:
            out.append(bytes([SEPARATOR]))
            self._pos += 1

    def _open_record(self, raw_name: bytes, out: List[bytes]) -> None:
        self._close_record(out)
        # name = stripped_line[1:]  (indexer.py:80); the line is already left-stripped
        self.names.append(raw_name.decode("utf-8").rstrip())
        self.starts.append(self._pos)
        self.lengths.append(0)
        self._open = True

    def _emit(self, seq: bytes, out: List[bytes]) -> None:
        if not self._open or not seq:      # text before the first header (indexer.py:82)
            return
        out.append(seq)
        self.lengths[-1] += len(seq)
        self._pos += len(seq)

    # -- chunk processing ------------------------------------------------------------
    def _sequence_block(self, seg: bytes, out: List[bytes]) -> None:
        """Lines between two headers.  Without blanks/tabs/... every line is already stripped and
        the block is just its bytes minus the line terminators; otherwise go line by line."""
        if _has_inline_whitespace(seg):
            self._process_slow(seg, out)
        else:
            self._emit(self._clean(seg), out)

    def _process_fast(self, buf: bytes, out: List[bytes]) -> None:
        """Split the chunk at the headers that start a line ('>' at the start of the chunk or right
        after a line terminator); a header hidden behind leading blanks stays inside a sequence
        block, which then contains white space and is handled line by line."""
        cur = 0
        n = len(buf)
        p = buf.find(b">")
        while p >= 0:
            if p == 0 or buf[p - 1] in (10, 13):
                if p > cur:
                    self._sequence_block(buf[cur:p], out)
                e1, e2 = buf.find(b"\n", p), buf.find(b"\r", p)
                end = min(x for x in (e1, e2, n) if x >= 0)
                self._open_record(buf[p + 1:end], out)
                cur = end
                p = buf.find(b">", end)
            else:
                p = buf.find(b">", p + 1)     # a '>' inside a sequence line: an invalid base
        if cur < n:
            self._sequence_block(buf[cur:], out)

    @staticmethod
    def _clean(seg: bytes) -> bytes:
        seq = seg.translate(None, b"\r\n")
        if not seq.isascii():
            raise ValueError("non-ASCII byte inside a sequence line; the reference cannot "
                             "index such input (indexer.py:37,75-76)")
        return seq

    def _process_slow(self, buf: bytes, out: List[bytes]) -> None:
        """General path: per line strip, exactly as the reference does."""
        for line in buf.replace(b"\r\n", b"\n").replace(b"\r", b"\n").split(b"\n"):
            line = line.strip(_STRIP)
            if not line:
                continue
            if line[:1] == b">":
                self._open_record(line[1:], out)
            else:
                if not line.isascii():
                    raise ValueError("non-ASCII byte inside a sequence line; the reference "
                                     "cannot index such input (indexer.py:37,75-76)")
                self._emit(line, out)

    def _raw_chunks(self) -> Iterator[bytes]:
        """Decompressed file contents in order, then one empty chunk as the end marker."""
        if self.path.endswith((".gz", ".bgz")) and is_bgzf(self.path):
            yield from bgzf_chunks(self.path, self.chunk_bytes)
        else:
            with open_binary(self.path) as fh:
                while True:
                    blk = fh.read(self.chunk_bytes)
                    if not blk:
                        break
                    yield blk
        yield b""

    def _prefetched_chunks(self, depth: int = 2) -> Iterator[bytes]:
        """_raw_chunks produced by a helper thread, so inflating overlaps parsing."""
        import queue
        import threading
        q: "queue.Queue" = queue.Queue(maxsize=depth)

        def work():
            try:
                for blk in self._raw_chunks():
                    q.put(blk)
            except BaseException as exc:          # re-raised in the consumer
                q.put(exc)

        th = threading.Thread(target=work, daemon=True)
        th.start()
        while True:
            item = q.get()
            if isinstance(item, BaseException):
                raise item
            yield item
            if not item:
                break
        th.join()

    def pieces(self) -> Iterator[np.ndarray]:
        tail = b""
        if True:
            for blk in self._prefetched_chunks():
                last = not blk
                buf = tail + blk
                if not last:
                    # keep the trailing partial line for the next round (a \r\n pair
                    # split across blocks only leaves a blank line, which vanishes)
                    cut = max(buf.rfind(b"\n"), buf.rfind(b"\r"))
                    if cut < 0:
                        tail = buf
                        continue
                    buf, tail = buf[:cut + 1], buf[cut + 1:]
                out: List[bytes] = []
                if buf:
                    self._process_fast(buf, out)
                if last:
                    self._close_record(out)
                    self._open = False
                if out:
                    yield np.frombuffer(b"".join(out), dtype=np.uint8)
                if last:
                    break

    def read_all(self) -> np.ndarray:
        parts = list(self.pieces())
        return np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)


def read_fasta_stream(path: str):
    """Whole file at once -> (stream uint8[], names, starts uint64[], lengths)."""
    fs = FastaStream(path)
    stream = fs.read_all()
    return stream, fs.names, np.asarray(fs.starts, dtype=np.uint64), fs.lengths

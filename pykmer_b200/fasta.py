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
_NONNL_WS = np.zeros(256, dtype=bool)
_NONNL_WS[[9, 11, 12, 28, 29, 30, 31, 32]] = True


def open_binary(path: str) -> BinaryIO:
    """read_fasta (indexer.py:108-115): gzip for .gz/.bgz, plain otherwise."""
    if path.endswith((".gz", ".bgz")):
        return gzip.open(path, "rb")
    return open(path, "rb")


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
    def _process_fast(self, buf: bytes, arr: np.ndarray, out: List[bytes]) -> None:
        """No blank/tab/etc. anywhere in the chunk: a line is stripped already."""
        gts = np.flatnonzero(arr == SEPARATOR)
        if gts.size:
            prev = arr[np.maximum(gts - 1, 0)]
            is_hdr = (gts == 0) | (prev == 10) | (prev == 13)
            hdrs = gts[is_hdr].tolist()
        else:
            hdrs = []
        cur = 0
        n = len(buf)
        for p in hdrs:
            if p < cur:          # a '>' inside a header line already consumed
                continue
            if p > cur:
                self._emit(self._clean(buf[cur:p]), out)
            e1, e2 = buf.find(b"\n", p), buf.find(b"\r", p)
            end = min(x for x in (e1, e2, n) if x >= 0)
            self._open_record(buf[p + 1:end], out)
            cur = end
        if cur < n:
            self._emit(self._clean(buf[cur:]), out)

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

    def pieces(self) -> Iterator[np.ndarray]:
        tail = b""
        with open_binary(self.path) as fh:
            while True:
                blk = fh.read(self.chunk_bytes)
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
                    arr = np.frombuffer(buf, dtype=np.uint8)
                    if _NONNL_WS[arr].any():
                        self._process_slow(buf, out)
                    else:
                        self._process_fast(buf, arr, out)
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

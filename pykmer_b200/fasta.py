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


_NATIVE = []


def _native_lib():
    """libpykmer_b200.so's host-side ingest helpers, or None when the library is not built."""
    if not _NATIVE:
        try:
            from . import _native
            _NATIVE.append(_native)
        except (ImportError, OSError):
            _NATIVE.append(None)
    return _NATIVE[0]


def _addr(buf, offset: int = 0) -> int:
    """Address of byte `offset` of a bytearray / numpy array (kept alive by the caller)."""
    arr = buf if isinstance(buf, np.ndarray) else np.frombuffer(buf, dtype=np.uint8)
    return arr.ctypes.data + offset


def bgzf_read_into(path: str, out: np.ndarray, threads: int = 0) -> int:
    """Inflate a whole BGZF file into `out` (uint8, large enough) on all cores through
    pk_bgzf_inflate; returns the number of bytes produced.  OSError on damage / truncation,
    ValueError if the contents do not fit."""
    import ctypes
    nat = _native_lib()
    if nat is None:
        raise ImportError("pykmer_b200.fasta.bgzf_read_into needs libpykmer_b200.so")
    assert out.dtype == np.uint8 and out.flags.c_contiguous and out.flags.writeable
    fill, comp, eof = 0, b"", False
    with open(path, "rb") as fh:
        while not (eof and not comp):
            if not eof and len(comp) < (8 << 20):
                blk = fh.read(32 << 20)
                eof = not blk
                comp += blk
            used, made = ctypes.c_size_t(0), ctypes.c_size_t(0)
            cbuf = np.frombuffer(comp, dtype=np.uint8)
            try:
                nat.check(nat.lib.pk_bgzf_inflate(cbuf.ctypes.data if comp else None, len(comp),
                                                  out.ctypes.data + fill, out.size - fill,
                                                  ctypes.byref(used), ctypes.byref(made), threads))
            except ValueError as exc:                      # PK_ERR_ARG: not BGZF / CRC / length
                raise OSError(f"{path}: {exc}") from None
            comp = comp[used.value:]
            fill += made.value
            if used.value == 0 and (eof or len(comp) >= (8 << 20)):
                if comp and out.size - fill < (1 << 16):
                    raise ValueError(f"{path}: more than the expected {out.size} bytes")
                if comp:
                    raise OSError(f"{path}: truncated BGZF block at end of file")
    return fill


class FastaStream:
    """Incremental FASTA -> stream converter.

    for piece in fs.pieces(): ...   yields uint8 arrays (cleaned bases + separators)
    fs.names / fs.starts / fs.lengths describe the records seen so far; a
    record's entry exists as soon as its header has been read.
    """

    def __init__(self, path: str, chunk_bytes: int = 64 << 20, threads: Optional[int] = None,
                 native: Optional[bool] = None):
        self.path = path
        self.chunk_bytes = chunk_bytes
        self.threads = threads or 0              # 0 = every core
        # native = the two passes over the text (BGZF inflate, newline stripping) run in
        # libpykmer_b200.so on all cores (csrc/ingest.cpp); False = this module's own Python
        # path, which is also what every unusual block of text falls back to.
        import queue
        self._text_pool: "queue.SimpleQueue" = queue.SimpleQueue()   # text buffers handed back by the consumer
        self.native = _native_lib() is not None if native is None else bool(native)
        if self.native and _native_lib() is None:
            raise ImportError("pykmer_b200.fasta: native ingest asked for, libpykmer_b200.so is missing")
        self.names: List[str] = []
        self.starts: List[int] = []
        self.lengths: List[int] = []
        self.bases = 0                          # running sum of self.lengths
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
        self.bases += len(seq)
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
                end = buf.find(b"\n", p)
                end = n if end < 0 else end
                cr = buf.find(b"\r", p, end)
                end = cr if cr >= 0 else end
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

    # -- native path ------------------------------------------------------------------
    def _native_text_chunks(self) -> Iterator[Tuple[bytearray, int, bool]]:
        """(text, n, last): text[:n] is the next run of whole lines (it ends at a line terminator
        unless last).  BGZF members are inflated on all cores straight into the buffer."""
        import ctypes
        nat = _native_lib()
        bgzf = self.path.endswith((".gz", ".bgz")) and is_bgzf(self.path)
        tail = b""
        comp = b""
        eof = False
        with (open(self.path, "rb") if bgzf else open_binary(self.path)) as fh:
            while True:
                # a BGZF member inflates to at most 64 KiB: keep room for one beyond the tail.  The
                # buffers go round (a fresh 64 MiB bytearray is zero-filled and page-faulted in each
                # time); nothing beyond `fill` is ever read, so stale bytes do no harm.
                want = len(tail) + (max(self.chunk_bytes, 1 << 16) if bgzf else self.chunk_bytes)
                text = None
                while text is None or len(text) < want:
                    try:
                        text = self._text_pool.get_nowait()
                    except Exception:
                        text = bytearray(want + (1 << 12))
                text[:len(tail)] = tail
                fill = len(tail)
                if bgzf:
                    while True:
                        if not eof and len(comp) < (8 << 20):
                            blk = fh.read(max(1 << 20, self.chunk_bytes // 3))
                            eof = not blk
                            comp += blk
                        used, made = ctypes.c_size_t(0), ctypes.c_size_t(0)
                        cbuf = np.frombuffer(comp, dtype=np.uint8)
                        nat.check(nat.lib.pk_bgzf_inflate(cbuf.ctypes.data if comp else None, len(comp),
                                                          _addr(text, fill), len(text) - fill,
                                                          ctypes.byref(used), ctypes.byref(made), self.threads))
                        comp = comp[used.value:]
                        fill += made.value
                        if eof and not comp:
                            break
                        if used.value == 0:
                            if len(text) - fill < (1 << 16):
                                break                      # the buffer is full
                            if eof:                        # room, no more input, yet no whole member
                                raise OSError(f"{self.path}: truncated BGZF block at end of file")
                    done = eof and not comp
                else:
                    got = fh.readinto(memoryview(text)[fill:])
                    fill += got or 0
                    done = not got
                if done:
                    yield text, fill, True
                    return
                cut = text.rfind(b"\n", 0, fill)           # last line terminator: a \r counts only if it
                cut = max(cut, text.rfind(b"\r", max(cut, 0), fill))   # comes after the last \n
                if cut < 0:                                # no whole line yet: grow the tail
                    tail = bytes(text[:fill])
                    continue
                tail = bytes(text[cut + 1:fill])
                yield text, cut + 1, False

    def _clean_into(self, text: bytearray, a: int, b: int, dst: np.ndarray, opos: int) -> int:
        """Lines text[a:b] between two line-initial headers -> dst[opos:]; returns the new opos."""
        import ctypes
        if b <= a:
            return opos
        nat = _native_lib()
        kept, flags = ctypes.c_size_t(0), ctypes.c_uint32(0)
        nat.check(nat.lib.pk_fasta_clean(_addr(text, a), b - a, _addr(dst, opos), ctypes.byref(kept),
                                         ctypes.byref(flags), self.threads))
        if flags.value == 0:
            if not self._open:                             # text before the first header (indexer.py:82)
                return opos
            n = kept.value
            self.lengths[-1] += n
            self.bases += n
            self._pos += n
            return opos + n
        # inner white space (maybe a header behind leading blanks) or non-ASCII bytes: this
        # module's line-by-line path, which keeps the record table itself
        out: List[bytes] = []
        self._sequence_block(bytes(text[a:b]), out)
        blob = b"".join(out)
        dst[opos:opos + len(blob)] = np.frombuffer(blob, dtype=np.uint8)
        return opos + len(blob)

    def _pieces_native(self, buffers=None) -> Iterator[np.ndarray]:
        turn = 0
        for text, n, last in self._prefetch(self._native_text_chunks()):
            need = n + 16
            dst = None
            if buffers is not None:
                dst = buffers[turn % len(buffers)]
                turn += 1
                if dst.size < need:                        # a line longer than the chunk: own array
                    dst = None
            if dst is None:
                dst = np.empty(need, dtype=np.uint8)
            opos = 0
            cur = 0
            for p in self._header_offsets(text, n):        # every '>' that starts a line
                if p < cur:
                    continue                               # inside the previous header line (lone \r quirks)
                opos = self._clean_into(text, cur, p, dst, opos)
                end = text.find(b"\n", p, n)               # the header line ends at the first \n or \r;
                end = n if end < 0 else end                # look for \r only before that \n (not to
                cr = text.find(b"\r", p, end)              # the end of a 64 MiB chunk of Unix text)
                end = cr if cr >= 0 else end
                if self._open:                             # close the previous record
                    dst[opos] = SEPARATOR
                    opos += 1
                    self._pos += 1
                self.names.append(bytes(text[p + 1:end]).decode("utf-8").rstrip())
                self.starts.append(self._pos)
                self.lengths.append(0)
                self._open = True
                cur = end
            opos = self._clean_into(text, cur, n, dst, opos)
            if last and self._open:
                dst[opos] = SEPARATOR
                opos += 1
                self._pos += 1
                self._open = False
            self._text_pool.put(text)                      # the producer may fill it again
            if opos:
                yield dst[:opos]

    def _header_offsets(self, text: bytearray, n: int):
        """Offsets in text[:n] of the '>' that start a line, ascending (pk_fasta_find_headers, all cores)."""
        import ctypes
        nat = _native_lib()
        cap = 1 << 12
        while True:
            pos = np.empty(cap, dtype=np.uint64)
            count = ctypes.c_size_t(0)
            nat.check(nat.lib.pk_fasta_find_headers(_addr(text, 0), n, pos.ctypes.data, cap, ctypes.byref(count),
                                                    self.threads))
            if count.value <= cap:
                return pos[:count.value].tolist()
            cap = count.value

    @staticmethod
    def _prefetch(gen, depth: int = 2):
        """Run a generator on a helper thread so that producing overlaps consuming."""
        import queue
        import threading
        q: "queue.Queue" = queue.Queue(maxsize=depth)
        end = object()

        def work():
            try:
                for item in gen:
                    q.put(item)
                q.put(end)
            except BaseException as exc:                   # re-raised in the consumer
                q.put(exc)

        th = threading.Thread(target=work, daemon=True)
        th.start()
        while True:
            item = q.get()
            if item is end:
                break
            if isinstance(item, BaseException):
                raise item
            yield item
        th.join()

    def pieces(self, buffers=None) -> Iterator[np.ndarray]:
        """The cleaned stream piece by piece.  buffers: optional ring of uint8 numpy arrays (e.g.
        views of pinned memory) of at least chunk_bytes + chunk_bytes/8 bytes each that the pieces
        are written into in turn -- a piece is then only valid until its buffer comes round again."""
        if self.native:
            yield from self._pieces_native(buffers)
            return
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

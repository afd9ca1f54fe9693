"""CPU oracle for the pykmer hot paths -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl
reference` leg may import this module.  The product (pykmer_b200/) never does.

Parity status: PINNED against the reference's own indexer.py / merger.py run in
the build container (oracle/make_golden.py -> tests/golden/).

The arithmetic lives in kmer_oracle.c (ctypes-loaded from libkmer_oracle.so);
this file adds the text rules of the reference's FASTA reader, restated in
plain Python, and small NumPy restatements used as cross-checks.
All citations are file:line into the reference (sauloal/pykmer).
"""
from __future__ import annotations

import ctypes
import gzip
import math
import os
import subprocess
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkmer_oracle.so")
_SRC = os.path.join(_HERE, "kmer_oracle.c")
_lib = None

SEPARATOR = ord(">")  # any byte outside ACGTacgt works as the record separator


def build(force: bool = False) -> str:
    """Compile kmer_oracle.c -> libkmer_oracle.so (gcc, a second or two)."""
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(_SRC)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libkmer_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        u8p, u64p, i64p, u32p = (ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_uint64),
                                 ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_uint32))
        idx_args = [u8p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64,
                    u8p, u64p, u64p, ctypes.c_size_t, u8p]
        L.ok_index_direct.argtypes = idx_args
        L.ok_index_rolling.argtypes = idx_args
        L.ok_index_rolling_mt.argtypes = [u8p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64,
                                          ctypes.c_uint64, u8p, u64p, ctypes.c_int]
        L.ok_table_stats.argtypes = [u8p, ctypes.c_size_t, i64p, u64p]
        L.ok_table_stats_mt.argtypes = [u8p, ctypes.c_size_t, i64p, u64p, ctypes.c_int]
        L.ok_pair_counts.argtypes = [u8p, u8p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, u64p]
        L.ok_merge_matrix.argtypes = [u8p, ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t,
                                      ctypes.c_int, ctypes.c_int, u64p, ctypes.c_int]
        L.ok_threshold_pack.argtypes = [u8p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, u32p]
        L.ok_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray], ct):
    if a is None:
        return ctypes.cast(None, ctypes.POINTER(ct))
    return a.ctypes.data_as(ctypes.POINTER(ct))


def max_threads() -> int:
    return int(lib().ok_max_threads())


# --------------------------------------------------------------------------- indexer

def index_stream(seq: np.ndarray, K: int, range_lo: int = 0, range_hi: Optional[int] = None,
                 rec_starts: Optional[np.ndarray] = None, method: str = "rolling",
                 table: Optional[np.ndarray] = None, threads: int = 1):
    """Count canonical K-mers of a cleaned byte stream (see kmer_oracle.c header).

    Returns (table uint8[range_hi-range_lo], num_kmers, rec_flags or None).
    """
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    if range_hi is None:
        range_hi = 4 ** K
    if table is None:
        table = np.zeros(range_hi - range_lo, dtype=np.uint8)
    num = np.zeros(1, dtype=np.uint64)
    flags = None
    nrec = 0
    if rec_starts is not None:
        rec_starts = np.ascontiguousarray(rec_starts, dtype=np.uint64)
        nrec = len(rec_starts)
        flags = np.zeros(max(nrec, 1), dtype=np.uint8)
    L = lib()
    if method == "mt":
        assert rec_starts is None
        rc = L.ok_index_rolling_mt(_p(seq, ctypes.c_uint8), seq.size, K, range_lo, range_hi,
                                   _p(table, ctypes.c_uint8), _p(num, ctypes.c_uint64), threads)
    else:
        fn = L.ok_index_direct if method == "direct" else L.ok_index_rolling
        rc = fn(_p(seq, ctypes.c_uint8), seq.size, K, range_lo, range_hi,
                _p(table, ctypes.c_uint8), _p(num, ctypes.c_uint64),
                _p(rec_starts, ctypes.c_uint64), nrec, _p(flags, ctypes.c_uint8))
    if rc != 0:
        raise ValueError(f"oracle rejected K={K}")
    return table, int(num[0]), (flags[:nrec] if flags is not None else None)


def table_stats(table: np.ndarray, threads: int = 1):
    """Header.update_stats (tools.py:246-263) -> (hist list[255], dict of 8 scalars)."""
    table = np.ascontiguousarray(table, dtype=np.uint8)
    hist = np.zeros(255, dtype=np.int64)
    st = np.zeros(4, dtype=np.uint64)
    if threads > 1:
        lib().ok_table_stats_mt(_p(table, ctypes.c_uint8), table.size, _p(hist, ctypes.c_int64),
                                _p(st, ctypes.c_uint64), threads)
    else:
        lib().ok_table_stats(_p(table, ctypes.c_uint8), table.size, _p(hist, ctypes.c_int64),
                             _p(st, ctypes.c_uint64))
    h = hist.tolist()
    return h, {
        "hist_sum": int(sum(h)), "hist_count": int(sum(1 for v in h if v)),
        "hist_min": int(min(h)), "hist_max": int(max(h)),
        "vals_sum": int(st[0]), "vals_count": int(st[1]),
        "vals_min": int(st[2]), "vals_max": int(st[3]),
    }


def table_stats_numpy(table: np.ndarray):
    """Same statistics with the NumPy calls the reference uses (tools.py:250-263)."""
    hist_v, _ = np.histogram(table, bins=255, range=(1, 255))
    return hist_v.tolist(), {
        "hist_sum": int(np.sum(hist_v)), "hist_count": int(np.count_nonzero(hist_v)),
        "hist_min": int(np.min(hist_v)), "hist_max": int(np.max(hist_v)),
        "vals_sum": int(np.sum(table, dtype=np.uint64)), "vals_count": int(np.count_nonzero(table)),
        "vals_min": int(np.min(table)), "vals_max": int(np.max(table)),
    }


def frag_size_rule(K: int, min_frag: int = 500_000_000, max_frag: int = 1_000_000_000) -> int:
    """Header.__init__ frag_size rule (tools.py:169-182)."""
    T = 4 ** K
    frag = T // 10
    frag = min(frag, max_frag)
    frag = max(frag, min_frag)
    frag = min(frag, T)
    if (T % frag) < (T // 2):
        pieces = T // frag
        frag = T // (pieces + 1)
        frag = frag + (pieces + 1) + 1
        frag = int(math.ceil(frag / 1000) * 1000)
    return frag


# FASTA text rules ---------------------------------------------------------------------

def open_text(path: str):
    """read_fasta (indexer.py:101-128): gzip text for .gz/.bgz, plain text otherwise."""
    if path.endswith((".gz", ".bgz")):
        return gzip.open(path, "rt")
    return open(path, "rt")


def parse_records(path: str) -> Iterator[Tuple[str, str]]:
    """parse_fasta (indexer.py:45-99) restated: strip every line, skip blanks, a
    stripped line starting with '>' opens a record named by the rest of the line,
    other lines are appended to the open record; text before the first header is
    dropped; every opened record is yielded, including empty ones."""
    name = None
    parts: List[str] = []
    with open_text(path) as fh:
        for raw in fh:
            line = raw.strip()
            if not line:
                continue
            if line.startswith(">"):
                if name is not None:
                    yield name, "".join(parts)
                name = line[1:]
                parts = []
            else:
                parts.append(line)
    if name is not None:
        yield name, "".join(parts)


def records_to_stream(records: Sequence[Tuple[str, str]]):
    """Concatenate records into the byte stream + record table the kernels eat.

    Code points >= 255 raise IndexError in the reference (CONV has 255 slots,
    indexer.py:37); the oracle raises the same way.
    """
    chunks, starts, lengths, names = [], [], [], []
    pos = 0
    for name, seq in records:
        for ch in seq:
            if ord(ch) >= 255:
                raise IndexError("list index out of range")
        b = seq.encode("latin-1")
        names.append(name)
        starts.append(pos)
        lengths.append(len(b))
        chunks.append(b)
        chunks.append(bytes([SEPARATOR]))
        pos += len(b) + 1
    stream = np.frombuffer(b"".join(chunks), dtype=np.uint8)
    return stream, np.asarray(starts, dtype=np.uint64), lengths, names


def index_fasta(path: str, K: int, method: str = "rolling"):
    """Whole indexer path on the CPU: what indexer.py writes for <path> <K>.

    Returns dict(table, num_kmers, chromosomes, hist, + the 8 statistics).
    chromosomes lists [name, length] only for records that produced at least
    one k-mer (indexer.py:349-351).
    """
    assert K > 0 and K % 2 == 1                                   # tools.py:165-167
    recs = list(parse_records(path))
    stream, starts, lengths, names = records_to_stream(recs)
    table, num, flags = index_stream(stream, K, rec_starts=starts, method=method)
    chrom = [[names[i], lengths[i]] for i in range(len(names)) if flags[i]]
    hist, st = table_stats(table)
    out = {"table": table, "num_kmers": num, "chromosomes": chrom, "hist": hist}
    out.update(st)
    return out


# --------------------------------------------------------------------------- merger

def pair_counts(s: np.ndarray, o: np.ndarray, min_count: int = 1, max_count: int = 255):
    """Header.calculate_distance arithmetic (tools.py:473-482) for one pair."""
    s = np.ascontiguousarray(s, dtype=np.uint8)
    o = np.ascontiguousarray(o, dtype=np.uint8)
    assert s.shape == o.shape
    out = np.zeros(3, dtype=np.uint64)
    lib().ok_pair_counts(_p(s, ctypes.c_uint8), _p(o, ctypes.c_uint8), s.size, min_count,
                         max_count, _p(out, ctypes.c_uint64))
    return tuple(int(v) for v in out)


def pair_counts_numpy(s, o, min_count=1, max_count=255):
    sv = (s >= min_count) & (s <= max_count)
    ov = (o >= min_count) & (o <= max_count)
    return int(np.sum(sv)), int(np.sum(ov)), int(np.sum(sv & ov))


def merge_matrix(tables: np.ndarray, min_count: int = 1, max_count: int = 255, threads: int = 1):
    """(N, N, 3) uint64 matrix of merger.py:136-176 (diagonal defined as T_k)."""
    tables = np.ascontiguousarray(tables, dtype=np.uint8)
    N, n = tables.shape
    m = np.zeros((N, N, 3), dtype=np.uint64)
    lib().ok_merge_matrix(_p(tables, ctypes.c_uint8), N, n, tables.strides[0], min_count,
                          max_count, _p(m, ctypes.c_uint64), threads)
    return m


def threshold_pack(table: np.ndarray, min_count: int = 1, max_count: int = 255):
    table = np.ascontiguousarray(table, dtype=np.uint8)
    bits = np.zeros((table.size + 31) // 32, dtype=np.uint32)
    lib().ok_threshold_pack(_p(table, ctypes.c_uint8), table.size, min_count, max_count,
                            _p(bits, ctypes.c_uint32))
    return bits


def tile_masks(rows: np.ndarray) -> np.ndarray:
    """Row-major mask words (N, words) -> the tiled layout of include/pykmer_b200.h, flat:
    word g of sample r at [(g // 32) * N * 32 + r * 32 + g % 32]; the last tile is zero-padded."""
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    n, words = rows.shape
    tiles = (words + 31) // 32
    padded = np.zeros((n, tiles * 32), dtype=np.uint32)
    padded[:, :words] = rows
    return np.ascontiguousarray(padded.reshape(n, tiles, 32).transpose(1, 0, 2)).reshape(-1)


def gram_from_bits(bits: np.ndarray) -> np.ndarray:
    """G[k,l] = popcount(bits[k] & bits[l]) -- the Gram form of the pair loop."""
    N = bits.shape[0]
    G = np.zeros((N, N), dtype=np.int64)
    for k in range(N):
        for l in range(k, N):
            G[k, l] = G[l, k] = int(np.bitwise_count(bits[k] & bits[l]).sum(dtype=np.int64))
    return G


def matrix_from_gram(G: np.ndarray) -> np.ndarray:
    """matrix[k,l] = (G[k,k], G[l,l], G[k,l]) (merger.py:175-176)."""
    N = G.shape[0]
    d = np.diag(G).astype(np.uint64)
    m = np.zeros((N, N, 3), dtype=np.uint64)
    m[:, :, 0] = d[:, None]
    m[:, :, 1] = d[None, :]
    m[:, :, 2] = G.astype(np.uint64)
    return m


def pack_table(table: np.ndarray):
    """The packed form of a table slice (include/pykmer_b200.h: pk_table_pack_device) restated with NumPy,
    chunks in ascending order: -> (bitmap uint64[n/64], chunk_off uint32[n/1024], nz uint8[]).  The device
    may lay the chunks out in any order, so compare through unpack_table, not array by array."""
    t = np.ascontiguousarray(table, dtype=np.uint8)
    assert t.size % 1024 == 0
    present = t != 0
    bitmap = np.packbits(present, bitorder="little").view(np.uint64)
    per_chunk = present.reshape(-1, 1024).sum(axis=1).astype(np.int64)
    units = (per_chunk + 15) // 16
    chunk_off = np.concatenate(([0], np.cumsum(units)[:-1])).astype(np.uint32)
    nz = np.zeros(int(units.sum()) * 16, dtype=np.uint8)
    vals = t[present]
    starts = np.concatenate(([0], np.cumsum(per_chunk)[:-1]))
    # byte k of chunk c goes to chunk_off[c] * 16 + k
    idx = np.repeat(chunk_off.astype(np.int64) * 16 - starts, per_chunk) + np.arange(vals.size)
    nz[idx] = vals
    return bitmap, chunk_off, nz


def unpack_table(bitmap: np.ndarray, chunk_off: np.ndarray, nz: np.ndarray, n: int) -> np.ndarray:
    """Literal inverse of pack_table for chunks laid out in any order."""
    present = np.unpackbits(np.ascontiguousarray(bitmap).view(np.uint8), bitorder="little")[:n].astype(bool)
    per_chunk = present.reshape(-1, 1024).sum(axis=1).astype(np.int64)
    starts = np.concatenate(([0], np.cumsum(per_chunk)[:-1]))
    idx = np.repeat(chunk_off.astype(np.int64) * 16 - starts, per_chunk) + np.arange(int(per_chunk.sum()))
    out = np.zeros(n, dtype=np.uint8)
    out[present] = nz[idx]
    return out

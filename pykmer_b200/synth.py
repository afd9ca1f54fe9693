"""Deterministic synthetic inputs for the parity tests and the benchmark.

Nothing here is on the hot path: it only manufactures inputs of the shapes
BASELINE.json names -- a tomato-sized multi-FASTA for
the indexer and sets of k-mer count tables for the merger -- because the real
genomes cannot be downloaded.  Everything is seeded so that the CPU oracle and
the GPU see byte-identical inputs.
"""
from __future__ import annotations

import gzip
import io
import struct
import zlib
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

# ---------------------------------------------------------------------------- sequences

# config 1: 10 Mbp, K=11
SYN10M_SEED = 0x5EED0000
SYN10M_LENGTHS = (3_000_000, 2_000_000, 1_500_000, 1_200_000, 1_000_000, 800_000, 400_000, 100_000)
# config 2/4/5: tomato-sized, ch00 and ch12 pinned by the reference README
SYN782M_SEED = 0x5EED0001
SYN782M_LENGTHS = (9_643_250, 90_863_682, 53_473_368, 65_298_490, 64_459_972, 65_269_487,
                   47_258_699, 67_883_646, 63_995_357, 68_513_564, 64_792_705, 54_379_777,
                   66_688_036)

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_IUPAC = np.frombuffer(b"RYKMSW", dtype=np.uint8)


def synth_sequence(rng: np.random.Generator, length: int, *, gc: float = 0.35,
                   repeat_frac: float = 0.40, library: Optional[List[np.ndarray]] = None,
                   micro_frac: float = 0.01, n_gaps_per_mbp: float = 0.25,
                   lower_frac: float = 0.30, iupac_rate: float = 1e-6) -> np.ndarray:
    """One record of upper/lower-case ACGT with repeats, microsatellites, N gaps and
    a sprinkle of IUPAC codes, as uint8 ASCII."""
    at, cg = (1.0 - gc) / 2.0, gc / 2.0
    seq = _BASES[rng.choice(4, size=length, p=[at, cg, cg, at]).astype(np.uint8)]
    if length == 0:
        return seq
    # dispersed repeats: copies drawn from a shared library, 3 % substitutions per copy
    if library and repeat_frac > 0:
        budget = int(length * repeat_frac)
        while budget > 0:
            el = library[int(rng.integers(len(library)))]
            n = min(len(el), length)
            at_pos = int(rng.integers(0, length - n + 1))
            copy = el[:n].copy()
            nsub = int(n * 0.03)
            if nsub:
                where = rng.integers(0, n, size=nsub)
                copy[where] = _BASES[rng.integers(0, 4, size=nsub)]
            seq[at_pos:at_pos + n] = copy
            budget -= n
    # microsatellites (A)n (AT)n (AAT)n: hot addresses, counters past 255
    if micro_frac > 0:
        budget = int(length * micro_frac)
        units = (b"A", b"AT", b"AAT")
        while budget > 0:
            run = int(rng.integers(50, 2001))
            run = min(run, length)
            unit = units[int(rng.integers(3))]
            rep = np.frombuffer((unit * (run // len(unit) + 1))[:run], dtype=np.uint8)
            at_pos = int(rng.integers(0, length - run + 1))
            seq[at_pos:at_pos + run] = rep
            budget -= run
    # N gaps
    ngaps = int(round(length / 1e6 * n_gaps_per_mbp))
    for _ in range(ngaps):
        g = int(min(length, np.exp(rng.uniform(np.log(100), np.log(50_000)))))
        at_pos = int(rng.integers(0, length - g + 1))
        seq[at_pos:at_pos + g] = ord("N")
    # IUPAC ambiguity codes
    nia = int(rng.poisson(length * iupac_rate)) if iupac_rate > 0 else 0
    if nia:
        seq[rng.integers(0, length, size=nia)] = _IUPAC[rng.integers(0, len(_IUPAC), size=nia)]
    # soft-masked (lower-case) blocks
    if lower_frac > 0:
        budget = int(length * lower_frac)
        while budget > 0:
            n = int(min(length, rng.integers(500, 20_001)))
            at_pos = int(rng.integers(0, length - n + 1))
            seq[at_pos:at_pos + n] |= 0x20
            budget -= n
    return seq


def repeat_library(rng: np.random.Generator, elements: int = 2000) -> List[np.ndarray]:
    lens = np.exp(rng.uniform(np.log(200), np.log(8000), size=elements)).astype(np.int64)
    return [_BASES[rng.integers(0, 4, size=int(n))] for n in lens]


def synth_genome(seed: int, lengths: Sequence[int], name_fmt: str = "SL4.0ch{:02d}",
                 library_elements: int = 2000) -> List[Tuple[str, np.ndarray]]:
    """List of (header text, uint8 sequence).  Headers carry a description after a
    blank, as real assemblies do (the reference keeps it, indexer.py:80)."""
    rng = np.random.default_rng(seed)
    lib = repeat_library(rng, library_elements)
    out = []
    for i, n in enumerate(lengths):
        gaps = 0.25 if n >= 1_000_000 else 1.0
        out.append((name_fmt.format(i) + f" synthetic len={n}",
                    synth_sequence(rng, int(n), library=lib, n_gaps_per_mbp=gaps)))
    return out


def syn10m_records() -> List[Tuple[str, np.ndarray]]:
    """Config 1: eight records (10.0 Mbp) plus three degenerate ones that must not
    show up in `chromosomes`: shorter than K, empty, and all-N."""
    recs = synth_genome(SYN10M_SEED, SYN10M_LENGTHS, name_fmt="syn10M_chr{:02d}",
                        library_elements=200)
    recs.insert(3, ("tiny shorter than K", np.frombuffer(b"ACGTA", dtype=np.uint8)))
    recs.insert(6, ("empty record", np.zeros(0, dtype=np.uint8)))
    recs.append(("allN gap only", np.full(1000, ord("N"), dtype=np.uint8)))
    return recs


def syn782m_records(scale: float = 1.0) -> List[Tuple[str, np.ndarray]]:
    """Config 2: 13 tomato-like records, 782,520,033 bp at scale 1."""
    lengths = [max(1, int(n * scale)) for n in SYN782M_LENGTHS]
    return synth_genome(SYN782M_SEED, lengths)


def records_to_stream(records: Iterable[Tuple[str, np.ndarray]], separator: int = ord(">")):
    """Cleaned byte stream (records joined by one non-ACGT byte) + record table."""
    parts, starts, lengths, names = [], [], [], []
    pos = 0
    sep = np.array([separator], dtype=np.uint8)
    for name, seq in records:
        names.append(name)
        starts.append(pos)
        lengths.append(int(seq.size))
        parts.append(seq)
        parts.append(sep)
        pos += int(seq.size) + 1
    stream = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
    return stream, np.asarray(starts, dtype=np.uint64), lengths, names


# ---------------------------------------------------------------------------- FASTA / BGZF

def fasta_bytes(records: Iterable[Tuple[str, np.ndarray]], line_width: int = 60,
                newline: bytes = b"\n") -> bytes:
    out = io.BytesIO()
    for name, seq in records:
        out.write(b">" + name.encode() + newline)
        n = int(seq.size)
        if n == 0:
            continue
        full = n // line_width
        if full:
            body = np.empty((full, line_width + len(newline)), dtype=np.uint8)
            body[:, :line_width] = seq[:full * line_width].reshape(full, line_width)
            body[:, line_width:] = np.frombuffer(newline, dtype=np.uint8)
            out.write(body.tobytes())
        if n % line_width:
            out.write(seq[full * line_width:].tobytes() + newline)
    return out.getvalue()


_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_compress(data: bytes, level: int = 6, block: int = 0xFF00) -> bytes:
    """BGZF container (gzip members of <= 64 KiB with the 'BC' extra field and the
    canonical EOF block) -- what `bgzip` writes; Python's gzip reads it as
    concatenated members."""
    out = io.BytesIO()
    mv = memoryview(data)
    for off in range(0, len(mv), block):
        chunk = mv[off:off + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        payload = co.compress(chunk) + co.flush()
        bsize = len(payload) + 25  # total block size - 1
        out.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00")
        out.write(struct.pack("<H", bsize))
        out.write(payload)
        out.write(struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    out.write(_BGZF_EOF)
    return out.getvalue()


def write_fasta(path: str, records: Iterable[Tuple[str, np.ndarray]], line_width: int = 60,
                newline: bytes = b"\n", level: int = 6) -> None:
    """Plain text, gzip (.gz) or BGZF (.bgz), chosen by the file name."""
    raw = fasta_bytes(records, line_width, newline)
    if path.endswith(".bgz"):
        raw = bgzf_compress(raw, level)
    elif path.endswith(".gz"):
        raw = gzip.compress(raw, compresslevel=level, mtime=0)
    with open(path, "wb") as fh:
        fh.write(raw)


# ---------------------------------------------------------------------------- merger tables

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_table_slice(sample: int, lo: int, hi: int) -> np.ndarray:
    """Entries [lo, hi) of synthetic count table `sample`.

    Counter-based: entry i of sample s depends only on (s, i), so any slice can be
    produced anywhere (CPU here, CUDA in pk_synth_table) bit-identically.  About
    16 % non-zero; ~47 % ones; 3 % in 51..178; a few saturated at 255.
    """
    with np.errstate(over="ignore"):
        i = np.arange(lo, hi, dtype=np.uint64)
        k1 = np.uint64((0x9E3779B97F4A7C15 * (sample % 5 + 1)) & 0xFFFFFFFFFFFFFFFF)
        k2 = np.uint64((0xD1B54A32D192ED03 * (sample + 1)) & 0xFFFFFFFFFFFFFFFF)
        h1 = _splitmix64(i ^ k1)
        h2 = _splitmix64(i ^ k2)
        present = ((h1 & np.uint64(0xFFFF)) < np.uint64(5243)) | \
                  ((h2 & np.uint64(0xFFFF)) < np.uint64(5898))
        t = (h2 >> np.uint64(16)) & np.uint64(0xFFFFFF)
        # trailing ones of t = trailing zeros of ~t
        nt = (~t) & np.uint64(0xFFFFFF)
        low = nt & (~nt + np.uint64(1))
        tz = np.where(nt == 0, 24, np.log2(np.maximum(low, np.uint64(1)).astype(np.float64))
                      ).astype(np.uint64)
        val = np.minimum(np.uint64(255), np.uint64(1) + tz)
        mid = ((h2 >> np.uint64(40)) & np.uint64(0xFF)) < np.uint64(8)
        val = np.where(mid, np.uint64(51) + ((h2 >> np.uint64(32)) & np.uint64(0x7F)), val)
        sat = (h2 >> np.uint64(48)) < np.uint64(43)
        val = np.where(sat, np.uint64(255), val)
        return np.where(present, val, np.uint64(0)).astype(np.uint8)


def synth_table(sample: int, K: int) -> np.ndarray:
    return synth_table_slice(sample, 0, 4 ** K)

"""File formats and metadata of pykmer -- the host-side mirror of the reference's
tools.py (Header / HeaderVars / Timer / gen_checksum, tools.py:24-556).

Same names, argument meaning, file naming, JSON keys and error behaviour as the
reference, so that .kin / .kin.json written here are drop-ins; the array
arithmetic the reference does with NumPy over a memmap (update_stats
tools.py:246-263, calculate_distance tools.py:439-493) runs in the CUDA library
instead.  There is no CPU fallback for those.
"""
from __future__ import annotations

import datetime
import gzip
import hashlib
import io
import json
import math
import os
import socket
from typing import Any, Dict, Iterator, List, Optional, Tuple

import numpy as np


class Timer:
    """Elapsed / delta speed bookkeeping (tools.py:24-64)."""

    def __init__(self) -> None:
        now = datetime.datetime.now()
        self.time_begin = now
        self.time_last = now
        self.val_last = 0
        self.val_delta = 0
        self.time_ela = datetime.timedelta(0)
        self.time_delta = datetime.timedelta(0)
        self.time_ela_s = "none"
        self.time_delta_s = "none"
        self.speed_ela = 0
        self.speed_delta = 0

    @property
    def time_delta_seconds(self) -> float:
        return (datetime.datetime.now() - self.time_last).total_seconds()

    def update(self, val: int) -> None:
        now = datetime.datetime.now()
        self.time_ela = now - self.time_begin
        self.time_delta = now - self.time_last
        self.time_ela_s = str(self.time_ela).split(".", 2)[0]
        self.time_delta_s = str(self.time_delta).split(".", 2)[0]
        self.val_delta = val - self.val_last
        ela = max(self.time_ela.total_seconds(), 1e-9)
        dlt = max(self.time_delta.total_seconds(), 1e-9)
        self.speed_ela = int(val // ela)
        self.speed_delta = int(self.val_delta // dlt)
        self.time_last = now
        self.val_last = val

    def __str__(self) -> str:
        return (f"ela   time {self.time_ela_s} val {self.val_last:15,d} speed {self.speed_ela:15,d}\n"
                f"delta time {self.time_delta_s} val {self.val_delta:15,d} speed {self.speed_delta:15,d}")


def gen_checksum(filename: str, chunk_size: int = 1 << 20) -> str:
    """sha256 of a file (tools.py:548-556)."""
    h = hashlib.sha256()
    with open(filename, "rb") as fh:
        for blk in iter(lambda: fh.read(chunk_size), b""):
            h.update(blk)
    return h.hexdigest()


class HeaderVars:
    """Constants of the format (tools.py:67-106)."""
    HEADER_VER = "KMER001"
    HEADER_FIXED = ["file_ver", "kmer_size", "data_size", "max_size"]
    HEADER_DATA = [
        "project_name", "kmer_len", "flush_every", "frag_size",
        "input_file_name", "input_file_path", "input_file_size", "input_file_ctime",
        "input_file_cheksum", "output_file_size", "output_file_ctime", "output_file_cheksum",
        "num_kmers", "chromosomes",
        "creation_time_start", "creation_time_end", "creation_duration", "creation_speed",
        "hostname", "checksum_script",
        "hist", "hist_sum", "hist_count", "hist_min", "hist_max",
        "vals_sum", "vals_count", "vals_min", "vals_max",
    ]
    NOT_LEAN = ["chromosomes"]

    IND_EXT = "kin"
    DESC_EXT = "json"
    TMP = "tmp"
    COMP_EXT = "bgz"

    DEFAULT_FLUSH_EVERY = 100_000_000
    DEFAULT_MIN_FRAG_SIZE = 500_000_000
    DEFAULT_MAX_FRAG_SIZE = 1_000_000_000
    DEFAULT_BUFFER_SIZE = io.DEFAULT_BUFFER_SIZE
    DEFAULT_MIN_COUNT = 1
    DEFAULT_MAX_COUNT = 255
    DEFAULT_BLOCK_SIZE = 100_000_000


def frag_size_rule(data_size: int, min_frag_size: Optional[int], max_frag_size: Optional[int]) -> int:
    """The fragment size the reference echoes into .kin.json (tools.py:169-182).
    Results never depended on it; it is reproduced because the key is compared."""
    frag = data_size // 10
    if max_frag_size is not None:
        frag = min(frag, max_frag_size)
    if min_frag_size is not None:
        frag = max(frag, min_frag_size)
    frag = min(frag, data_size)
    if data_size % frag < data_size // 2:
        pieces = data_size // frag + 1
        frag = data_size // pieces + pieces + 1
        frag = int(math.ceil(frag / 1_000) * 1_000)
    return frag


class Header(HeaderVars):
    """Names, sizes and metadata of one k-mer index (tools.py:110-545)."""

    def __init__(self, project_name: str, input_file: Optional[str] = None,
                 kmer_len: Optional[int] = None, index_file: Optional[str] = None,
                 frag_size: Optional[int] = None,
                 flush_every: int = HeaderVars.DEFAULT_FLUSH_EVERY,
                 min_frag_size: Optional[int] = HeaderVars.DEFAULT_MIN_FRAG_SIZE,
                 max_frag_size: Optional[int] = HeaderVars.DEFAULT_MAX_FRAG_SIZE,
                 buffer_size: int = HeaderVars.DEFAULT_BUFFER_SIZE,
                 sample_name: Optional[str] = None):
        self.project_name = project_name
        self.sample_name = sample_name      # accepted (indexer.py:311-322), never persisted
        self.input_file_name = os.path.basename(input_file) if input_file else input_file
        self.input_file_path = os.path.abspath(input_file) if input_file else input_file
        self.kmer_len = kmer_len
        self.flush_every = flush_every
        self._buffer_size = buffer_size
        for key in ("input_file_size", "input_file_ctime", "input_file_cheksum",
                    "output_file_size", "output_file_ctime", "output_file_cheksum",
                    "num_kmers", "chromosomes", "creation_time_start", "creation_time_end",
                    "creation_duration", "creation_speed", "hostname", "checksum_script",
                    "hist", "hist_sum", "hist_count", "hist_min", "hist_max",
                    "vals_sum", "vals_count", "vals_min", "vals_max"):
            setattr(self, key, None)
        self.timer = Timer()

        if index_file is not None:
            self._parse_index_file_name(index_file)
            self.read_metadata()

        assert self.kmer_len                       # tools.py:165-167
        assert self.kmer_len > 0
        assert self.kmer_len % 2 == 1

        if frag_size is not None:
            self.frag_size = frag_size
        else:
            self.frag_size = frag_size_rule(self.data_size, min_frag_size, max_frag_size)

    # -- names (tools.py:185-202) ------------------------------------------------------
    @property
    def index_file_root(self) -> str:
        return f"{self.input_file_path}.{self.kmer_len:02d}.{self.IND_EXT}"

    @property
    def index_file(self) -> str:
        packed = f"{self.index_file_root}.{self.COMP_EXT}"
        return packed if os.path.exists(packed) else self.index_file_root

    @property
    def index_file_basename(self) -> str:
        return os.path.basename(self.index_file)

    @property
    def index_tmp_file(self) -> str:
        return f"{self.index_file_root}.{self.TMP}"

    @property
    def metadata_file(self) -> str:
        return f"{self.index_file_root}.{self.DESC_EXT}"

    # -- sizes (tools.py:204-217) ------------------------------------------------------
    @property
    def kmer_size(self) -> int:
        return 4 ** self.kmer_len

    @property
    def data_size(self) -> int:
        return self.kmer_size

    @property
    def max_size(self) -> int:
        return self.data_size

    @property
    def file_ver(self) -> str:
        return self.HEADER_VER

    @property
    def max_val(self) -> int:
        return 255

    def _parse_index_file_name(self, index_file: str) -> None:
        """<input>.<KK>.kin[.bgz] -> input path and K (tools.py:220-238)."""
        suffix = "." + self.COMP_EXT
        if index_file.endswith(suffix):
            index_file = index_file[:-len(suffix)]
        tail = len(f".00.{self.IND_EXT}")           # ".KK.kin"
        if self.input_file_name is None:
            stem = index_file[:-tail]
            self.input_file_name = os.path.basename(stem)
            self.input_file_path = os.path.abspath(stem)
        if self.kmer_len is None:
            self.kmer_len = int(index_file[-tail + 1:-tail + 3])

    # -- files (tools.py:294-350) ------------------------------------------------------
    def open_file(self, index_file: str, mode: str = "r+b") -> Iterator[Any]:
        """Yields one binary handle; .bgz files are gunzipped on the fly."""
        with open(index_file, mode, buffering=self._buffer_size) as fh:
            if index_file.endswith("." + self.COMP_EXT):
                with gzip.open(fh, "rb") as fz:
                    yield fz
            else:
                yield fh

    def open_index_file(self, mode: str = "r+b"):
        return self.open_file(self.index_file, mode=mode)

    def open_index_tmp_file(self, mode: str = "r+b"):
        return self.open_file(self.index_tmp_file, mode=mode)

    def _init_clean(self, overwrite: bool = False) -> None:
        for path in (self.index_file, self.index_file_root):
            if os.path.exists(path):
                if not overwrite:
                    raise ValueError(f"file {path} already exists and overwritting disabled")
                os.remove(path)
        for path in (self.metadata_file, self.index_tmp_file):
            if os.path.exists(path):
                os.remove(path)

    def init_file(self, index_file: str, mode: str = "r+b") -> None:
        """Sparse file of max_size bytes (tools.py:333-341)."""
        with open(index_file, "ab"):
            pass
        with open(index_file, mode) as fh:
            fh.seek(self.max_size - 1)
            fh.write(b"\0")

    def init_index_file(self, overwrite: bool = False, mode: str = "r+b") -> None:
        """tools.py:344-346 (the reference's own call passes self twice and cannot run)."""
        self._init_clean(overwrite=overwrite)
        self.init_file(self.index_file_root, mode=mode)

    def init_index_tmp_file(self, overwrite: bool = False, mode: str = "r+b") -> None:
        self._init_clean(overwrite=overwrite)
        self.init_file(self.index_tmp_file, mode=mode)

    # -- memmap views (tools.py:240-243, 353-363): host access to the raw table ----------
    def _get_mmap(self, fhd, offset: int = 0, mode: str = "r+") -> Iterator[np.memmap]:
        view = np.memmap(fhd, dtype=np.uint8, mode=mode, offset=offset, shape=(self.data_size,))
        yield view
        del view

    def get_array_from_fhd(self, fhd, mode: str = "r+") -> Iterator[np.memmap]:
        yield from self._get_mmap(fhd, offset=0, mode=mode)

    def get_array_from_index_file(self, fhd_mode: str = "r+b", mm_mode: str = "r+") -> Iterator[np.memmap]:
        for fhd in self.open_index_file(mode=fhd_mode):
            yield from self.get_array_from_fhd(fhd, mode=mm_mode)

    def get_array_from_index_tmp_file(self, fhd_mode: str = "r+b", mm_mode: str = "r+") -> Iterator[np.memmap]:
        for fhd in self.open_index_tmp_file(mode=fhd_mode):
            yield from self.get_array_from_fhd(fhd, mode=mm_mode)

    def read_table(self, index_file: Optional[str] = None) -> np.ndarray:
        """The uint8[4^K] table of a .kin / .kin.bgz file."""
        path = index_file or self.index_file
        if path.endswith("." + self.COMP_EXT):
            from .fasta import bgzf_read_into, is_bgzf
            if is_bgzf(path):                       # independent blocks: inflate on all host cores,
                arr = np.empty(self.data_size, dtype=np.uint8)      # straight into the table
                got = bgzf_read_into(path, arr)
                assert got == self.data_size, f"{path}: {got} bytes, expected {self.data_size}"
            else:                                   # plain gzip stream, as the reference reads it
                with gzip.open(path, "rb") as fz:
                    arr = np.frombuffer(fz.read(), dtype=np.uint8)
        else:
            arr = np.fromfile(path, dtype=np.uint8)
        assert arr.size == self.data_size, f"{path}: {arr.size} bytes, expected {self.data_size}"
        return arr

    def read_table_slice(self, lo: int, hi: int, index_file: Optional[str] = None) -> np.ndarray:
        """Entries [lo, hi) of the table -- the share of the k-mer axis one rank of a multi-GPU
        merge works on.  A BGZF file gives up a slice by inflating only the members that hold it
        (bgzf.read_range, with the file's .gzi when there is one); a raw .kin is read at an offset."""
        path = index_file or self.index_file
        assert 0 <= lo <= hi <= self.data_size
        if path.endswith("." + self.COMP_EXT):
            from .fasta import is_bgzf
            if is_bgzf(path):
                from .bgzf import read_range
                arr = read_range(path, lo, hi)
            else:
                arr = self.read_table(path)[lo:hi]
        else:
            assert os.path.getsize(path) == self.data_size, \
                f"{path}: {os.path.getsize(path)} bytes, expected {self.data_size}"
            arr = np.fromfile(path, dtype=np.uint8, count=hi - lo, offset=lo)
        assert arr.size == hi - lo, f"{path}: {arr.size} bytes in [{lo}, {hi})"
        return arr

    # -- statistics (tools.py:246-263), computed on the GPU ---------------------------------
    def set_stats(self, hist: List[int], vals_sum: int, vals_count: int, vals_min: int,
                  vals_max: int) -> None:
        self.hist = [int(v) for v in hist]
        self.hist_sum = int(sum(self.hist))
        self.hist_count = int(sum(1 for v in self.hist if v))
        self.hist_min = int(min(self.hist))
        self.hist_max = int(max(self.hist))
        self.vals_sum, self.vals_count = int(vals_sum), int(vals_count)
        self.vals_min, self.vals_max = int(vals_min), int(vals_max)

    def update_stats(self, fhd) -> None:
        from . import device
        table = np.frombuffer(fhd.read(), dtype=np.uint8)
        assert table.size == self.data_size
        hist, st = device.table_stats(table)
        self.set_stats(hist, *st)

    def update_stats_index_file(self) -> None:
        for fhd in self.open_index_file():
            self.update_stats(fhd)

    def update_stats_index_tmp_file(self) -> None:
        for fhd in self.open_index_tmp_file():
            self.update_stats(fhd)

    # -- metadata (tools.py:273-291, 366-401) ----------------------------------------------
    def update_metadata(self, index_file: str, output_checksum: Optional[str] = None,
                        input_checksum: Optional[str] = None) -> None:
        self.input_file_size = os.path.getsize(self.input_file_path)
        self.input_file_ctime = os.path.getctime(self.input_file_path)
        self.input_file_cheksum = input_checksum or gen_checksum(self.input_file_path)
        self.output_file_size = os.path.getsize(index_file)
        self.output_file_ctime = os.path.getctime(index_file)
        self.output_file_cheksum = output_checksum or gen_checksum(index_file)
        self.hostname = socket.gethostname()
        self.checksum_script = gen_checksum(os.path.abspath(__file__))
        end = datetime.datetime.now()
        self.creation_time_start = str(self.timer.time_begin)
        self.creation_time_end = str(end)
        self.creation_duration = str(end - self.timer.time_begin)
        self.creation_speed = self.timer.speed_ela

    def write_metadata_file(self, index_file: str, output_checksum: Optional[str] = None,
                            recompute_stats: bool = False, input_checksum: Optional[str] = None) -> None:
        assert self.num_kmers          # tools.py:367-368: an input without k-mers is an error
        assert self.chromosomes
        self.update_metadata(index_file, output_checksum, input_checksum)
        if recompute_stats or self.hist is None:
            for fhd in self.open_file(index_file):
                self.update_stats(fhd)
        with open(self.metadata_file, "wt") as fh:
            json.dump(self.to_dict(), fh, indent=1, sort_keys=True)

    def write_metadata_index_file(self) -> None:
        self.write_metadata_file(self.index_file)

    def write_metadata_index_tmp_file(self, output_checksum: Optional[str] = None,
                                      input_checksum: Optional[str] = None) -> None:
        self.write_metadata_file(self.index_tmp_file, output_checksum, input_checksum=input_checksum)

    def read_metadata(self) -> None:
        with open(self.metadata_file, "rt") as fh:
            data = json.load(fh)
        for key in self.HEADER_DATA:
            setattr(self, key, data[key])            # KeyError on a missing key, as the reference
        for key in self.HEADER_FIXED:
            mine, theirs = getattr(self, key), data[key]
            assert mine == theirs, f"self.{key} != header_data[{key}]: {mine} != {theirs}"

    def check_data(self, fhd) -> None:
        """Recompute the statistics of a .kin and compare with its JSON (tools.py:404-426)."""
        self.read_metadata()
        other = self.__class__(self.project_name, input_file=self.input_file_path,
                               kmer_len=self.kmer_len)
        other.read_metadata()
        for key in ("project_name", "input_file_name", "input_file_path", "kmer_len", "num_kmers"):
            assert getattr(self, key) == getattr(other, key), key
        other.update_stats(fhd)
        for key in ("hist", "hist_sum", "hist_count", "hist_min", "hist_max",
                    "vals_sum", "vals_count", "vals_min", "vals_max"):
            assert getattr(self, key) == getattr(other, key), key

    def check_data_file(self, filename: str) -> None:
        for fhd in self.open_file(filename, mode="rb"):
            self.check_data(fhd)

    def check_data_index(self) -> None:
        self.check_data_file(self.index_file)

    def check_data_index_tmp(self) -> None:
        self.check_data_file(self.index_tmp_file)

    # -- distance (tools.py:439-493), computed on the GPU ------------------------------------
    def calculate_distance(self, other: "Header", min_count: int = HeaderVars.DEFAULT_MIN_COUNT,
                           max_count: int = HeaderVars.DEFAULT_MAX_COUNT,
                           block_size: int = HeaderVars.DEFAULT_BLOCK_SIZE,
                           threading: bool = False) -> Tuple[int, int, int]:
        """(Total self, Total other, Shared) under min_count <= count <= max_count."""
        from . import device
        assert self.data_size == other.data_size
        return device.pair_counts(self.read_table(), other.read_table(), min_count, max_count)

    def calculate_distance2(self, other: "Header", min_count: int = HeaderVars.DEFAULT_MIN_COUNT,
                            max_count: int = HeaderVars.DEFAULT_MAX_COUNT) -> Tuple[int, int, int]:
        """The reference's entry-by-entry restatement of calculate_distance (tools.py:495-512,
        unused there).  Same three sums, so it is the same kernel here."""
        return self.calculate_distance(other, min_count=min_count, max_count=max_count)

    def __iter__(self) -> Iterator[int]:
        """The table's bytes as ints, .bgz inflated on the way (tools.py:527-533)."""
        for fhd in self.open_index_file(mode="rb"):
            for blk in iter(lambda: fhd.read(max(self._buffer_size, 1 << 16)), b""):
                yield from blk

    def to_dict(self, lean: bool = False) -> Dict[str, Any]:
        keys = self.HEADER_FIXED + self.HEADER_DATA
        return {k: getattr(self, k) for k in keys if not (lean and k in self.NOT_LEAN)}

    def to_json(self, indent: int = 1, sort_keys: bool = True) -> str:
        return json.dumps(self.to_dict(), indent=indent, sort_keys=sort_keys)

    def __str__(self) -> str:
        rows = []
        for k, v in self.to_dict().items():
            rows.append(f"{k:20s}: {v:15,d}" if isinstance(v, int) and not isinstance(v, bool)
                        else f"{k:20s}: {str(v)[:50]}")
        return "\n".join(rows) + "\n"

    __repr__ = __str__

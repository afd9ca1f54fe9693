"""Thin Python layer over the C ABI: PyTorch owns the buffers (pinned host and
device tensors, streams), libpykmer_b200.so does the work.  Nothing here computes;
if the CUDA library or a GPU is missing the calls fail loudly.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nat

lib = nat.lib
HAS_SCAN_MODE = True       # sequence-sharded multi-GPU indexing (scan-only handles, peer-mapped buffers)


def _stream_ptr(stream: Optional["torch.cuda.Stream"] = None) -> int:
    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)


def pinned_empty(nbytes: int) -> torch.Tensor:
    return torch.empty(int(nbytes), dtype=torch.uint8, pin_memory=True)


def device_scope(device: int):
    """`with device_scope(d):` -- allocations and launches inside go to GPU d."""
    return torch.cuda.device(device)


def zeros(shape, dtype=torch.int32) -> torch.Tensor:
    return torch.zeros(shape, dtype=dtype, device="cuda")


def upload(t: torch.Tensor, non_blocking: bool = False) -> torch.Tensor:
    """CPU tensor (ideally pinned) -> current GPU; a CUDA tensor passes through."""
    return t if t.is_cuda else t.to("cuda", non_blocking=non_blocking)


def free_memory_bytes() -> int:
    """Free memory of the current GPU."""
    return int(torch.cuda.mem_get_info()[0])


def stream_sync() -> None:
    torch.cuda.current_stream().synchronize()


def to_device_u8(arr, device: Optional[int] = None) -> torch.Tensor:
    """uint8 numpy array / tensor -> contiguous CUDA tensor (torch does the copy)."""
    if isinstance(arr, torch.Tensor):
        t = arr
    else:
        a = np.ascontiguousarray(arr, dtype=np.uint8)
        t = torch.from_numpy(a) if a.flags.writeable else torch.from_numpy(a.copy())
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    return t.to(dev, non_blocking=False).contiguous()


class _DeviceAlias:
    """Minimal __cuda_array_interface__ carrier: lets torch view library-owned device memory."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False),
                                         "version": 2, "strides": None}


def _alias_device_memory(ptr: int, n_int32: int, device: int) -> torch.Tensor:
    if n_int32 == 0 or ptr == 0:
        return torch.empty(0, dtype=torch.int32, device=torch.device("cuda", device))
    with torch.cuda.device(device):
        return torch.as_tensor(_DeviceAlias(ptr, n_int32), device=torch.device("cuda", device))


class Indexer:
    """One pk_indexer handle (one GPU, one k-mer range)."""

    def __init__(self, kmer_len: int, device: int = 0, range_lo: int = 0,
                 range_hi: Optional[int] = None, mode: int = nat.PK_MODE_AUTO):
        self.kmer_len = kmer_len
        self.device = device
        self.range_lo = range_lo
        # range_hi = 0 lets the library pick 4^K (and reject a bad K with its own message)
        self.range_hi = (4 ** kmer_len if 0 < kmer_len <= 31 else 0) if range_hi is None else range_hi
        self._h = ctypes.c_void_p()
        nat.check(lib.pk_indexer_create(ctypes.byref(self._h), kmer_len, device, range_lo,
                                        self.range_hi, mode))
        self._nrec = 0
        self._keep: List[object] = []        # host buffers that async feeds still read

    def close(self) -> None:
        if self._h:
            lib.pk_indexer_destroy(self._h)
            self._h = ctypes.c_void_p()
        self._keep.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self, stream=None) -> None:
        nat.check(lib.pk_indexer_reset(self._h, _stream_ptr(stream)))

    def set_records(self, starts) -> None:
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        nat.check(lib.pk_indexer_set_records(self._h, starts.ctypes.data, starts.size))
        self._nrec = int(starts.size)

    def append_records(self, new_starts) -> None:
        """Grow the record table by the offsets in new_starts (only these are checked and uploaded)."""
        new_starts = np.ascontiguousarray(new_starts, dtype=np.uint64)
        if new_starts.size:
            nat.check(lib.pk_indexer_append_records(self._h, new_starts.ctypes.data, new_starts.size))
            self._nrec += int(new_starts.size)

    def flush(self) -> None:
        """PARTITION mode: count what is buffered into the table now (asynchronous)."""
        nat.check(lib.pk_indexer_flush(self._h))
        self._keep.clear()

    def feed_device(self, seq: torch.Tensor, stream=None) -> None:
        assert seq.is_cuda and seq.dtype == torch.uint8 and seq.is_contiguous()
        nat.check(lib.pk_indexer_feed_device(self._h, seq.data_ptr(), seq.numel(),
                                             _stream_ptr(stream)))

    def feed_host(self, seq) -> None:
        """seq: uint8 numpy array or (ideally pinned) CPU tensor.  Asynchronous."""
        if isinstance(seq, torch.Tensor):
            assert seq.dtype == torch.uint8 and seq.is_contiguous() and not seq.is_cuda
            p, n = seq.data_ptr(), seq.numel()
        else:
            seq = np.ascontiguousarray(seq, dtype=np.uint8)
            p, n = seq.ctypes.data, seq.size
        self._keep.append(seq)
        nat.check(lib.pk_indexer_feed_host(self._h, p, n))

    def sync(self) -> None:
        nat.check(lib.pk_indexer_sync(self._h))
        self._keep.clear()

    def finalize(self, table_out=None) -> Tuple[List[int], dict]:
        """Finish counting -> (hist[255], statistics).  table_out (CPU tensor / numpy array of
        range_hi - range_lo bytes, ideally pinned) also receives the table, window by window
        while the rest is still being counted."""
        hist = np.zeros(255, dtype=np.int64)
        st = np.zeros(5, dtype=np.uint64)
        if table_out is None:
            nat.check(lib.pk_indexer_finalize(self._h, hist.ctypes.data, st.ctypes.data))
        else:
            n = table_out.numel() if isinstance(table_out, torch.Tensor) else table_out.size
            assert n >= self.range_hi - self.range_lo
            p = table_out.data_ptr() if isinstance(table_out, torch.Tensor) else table_out.ctypes.data
            nat.check(lib.pk_indexer_finalize_to_host(self._h, hist.ctypes.data, st.ctypes.data, p))
        self._keep.clear()
        return hist.tolist(), {"num_kmers": int(st[0]), "vals_sum": int(st[1]),
                               "vals_count": int(st[2]), "vals_min": int(st[3]),
                               "vals_max": int(st[4])}

    def transfer_stats(self) -> dict:
        """What the last finalize(table_out=...) moved over PCIe (pk_indexer_transfer_stats)."""
        st = np.zeros(4, dtype=np.uint64)
        nat.check(lib.pk_indexer_transfer_stats(self._h, st.ctypes.data))
        return {"d2h_bytes": int(st[0]), "packed_windows": int(st[1]), "raw_windows": int(st[2]),
                "unpack_threads": int(st[3])}

    def record_flags(self) -> np.ndarray:
        flags = np.zeros(max(self._nrec, 1), dtype=np.uint8)
        nat.check(lib.pk_indexer_record_flags(self._h, flags.ctypes.data, self._nrec))
        return flags[:self._nrec]

    def table_ptr(self) -> Tuple[int, int]:
        p, n = ctypes.c_void_p(), ctypes.c_size_t(0)
        nat.check(lib.pk_indexer_table_device(self._h, ctypes.byref(p), ctypes.byref(n)))
        return int(p.value), int(n.value)

    def table_to_host(self, dst=None, offset: int = 0, nbytes: Optional[int] = None):
        """Copy table[offset:offset+nbytes] into dst (numpy / CPU tensor; pinned is fastest)."""
        total = self.range_hi - self.range_lo
        nbytes = total - offset if nbytes is None else nbytes
        if dst is None:
            dst = pinned_empty(nbytes)
        p = dst.data_ptr() if isinstance(dst, torch.Tensor) else dst.ctypes.data
        nat.check(lib.pk_indexer_table_to_host(self._h, p, offset, nbytes))
        return dst

    def mode(self) -> Tuple[int, int]:
        """(counting scheme in use, number of table windows)."""
        m, w = ctypes.c_int(0), ctypes.c_int(0)
        nat.check(lib.pk_indexer_mode(self._h, ctypes.byref(m), ctypes.byref(w)))
        return m.value, w.value

    def window_log2(self) -> int:
        """log2 of the table entries per counting window (0 in DIRECT mode)."""
        v = ctypes.c_int(0)
        nat.check(lib.pk_indexer_window_log2(self._h, ctypes.byref(v)))
        return v.value

    # ---- sequence-sharded multi-GPU (see include/pykmer_b200.h) ------------------------------
    def prime(self, halo: Optional[torch.Tensor], stream_off: int, stream=None) -> None:
        """Begin a slice in mid-stream: `halo` = the <= 32 bytes preceding it (CUDA uint8)."""
        n = 0 if halo is None else halo.numel()
        p = 0 if halo is None else halo.data_ptr()
        nat.check(lib.pk_indexer_prime(self._h, p, n, stream_off, _stream_ptr(stream)))

    def scan_result(self) -> int:
        n = ctypes.c_uint64(0)
        nat.check(lib.pk_indexer_scan_result(self._h, ctypes.byref(n)))
        return int(n.value)

    def export_segments(self):
        """-> (entries: int32 CUDA tensor aliasing the handle's k-mer buffer,
                seg_off, seg_cnt: uint32 numpy arrays of shape (nseg, nwindows))."""
        p, ns, nw = ctypes.c_void_p(), ctypes.c_uint32(0), ctypes.c_uint32(0)
        nat.check(lib.pk_indexer_export_segments(self._h, ctypes.byref(p), ctypes.byref(ns),
                                                 ctypes.byref(nw), None, None, 0))
        off = np.zeros((ns.value, nw.value), dtype=np.uint32)
        cnt = np.zeros((ns.value, nw.value), dtype=np.uint32)
        if off.size:
            nat.check(lib.pk_indexer_export_segments(self._h, ctypes.byref(p), ctypes.byref(ns),
                                                     ctypes.byref(nw), off.ctypes.data, cnt.ctypes.data,
                                                     off.size))
        used = int((off.astype(np.int64) + cnt).max()) if off.size else 0
        entries = _alias_device_memory(int(p.value or 0), used, self.device)
        return entries, off, cnt

    def import_segments(self, entries: torch.Tensor, seg_off: np.ndarray, seg_cnt: np.ndarray) -> None:
        """Count k-mer entries gathered from other ranks; `entries` must outlive finalize()."""
        seg_off = np.ascontiguousarray(seg_off, dtype=np.uint32)
        seg_cnt = np.ascontiguousarray(seg_cnt, dtype=np.uint32)
        assert seg_off.shape == seg_cnt.shape and seg_off.ndim == 2
        self._keep.append(entries)
        nat.check(lib.pk_indexer_import_segments(self._h, entries.data_ptr() if entries.numel() else 0,
                                                 seg_off.shape[0], seg_off.ctypes.data, seg_cnt.ctypes.data))

    # ---- fused exchange: pass 2 stores into the window owners' buffers over NVLink ------------
    def pool_ipc_handle(self) -> Tuple[bytes, int]:
        buf = ctypes.create_string_buffer(64)
        cap = ctypes.c_size_t(0)
        nat.check(lib.pk_indexer_pool_ipc_handle(self._h, buf, ctypes.byref(cap)))
        return buf.raw, int(cap.value)

    def open_peer_pool(self, peer: int, handle: Optional[bytes] = None, local_owner: "Indexer" = None) -> None:
        if local_owner is not None:
            nat.check(lib.pk_indexer_open_peer_pool(self._h, peer, None, local_owner._h))
        else:
            buf = ctypes.create_string_buffer(handle, 64)
            nat.check(lib.pk_indexer_open_peer_pool(self._h, peer, buf, None))

    def scan_pass1(self, seq: torch.Tensor, stream=None) -> None:
        assert seq.is_cuda and seq.dtype == torch.uint8 and seq.is_contiguous()
        self._keep.append(seq)
        nat.check(lib.pk_indexer_scan_pass1(self._h, seq.data_ptr(), seq.numel(), _stream_ptr(stream)))

    def pass1_counts(self) -> np.ndarray:
        nwin = self.mode()[1]
        cnt = np.zeros(nwin, dtype=np.uint32)
        nat.check(lib.pk_indexer_pass1_counts(self._h, cnt.ctypes.data, nwin))
        return cnt

    def scan_pass2_remote(self, nranks: int, owner: np.ndarray, dest_off: np.ndarray, stream=None) -> None:
        owner = np.ascontiguousarray(owner, dtype=np.uint32)
        dest_off = np.ascontiguousarray(dest_off, dtype=np.uint32)
        nat.check(lib.pk_indexer_scan_pass2_remote(self._h, nranks, owner.ctypes.data, dest_off.ctypes.data,
                                                   _stream_ptr(stream)))
        self._keep.clear()

    def import_own_pool(self, seg_off: np.ndarray, seg_cnt: np.ndarray) -> None:
        """Count entries that peers stored into this handle's own k-mer buffer."""
        seg_off = np.ascontiguousarray(seg_off, dtype=np.uint32)
        seg_cnt = np.ascontiguousarray(seg_cnt, dtype=np.uint32)
        nat.check(lib.pk_indexer_import_segments(self._h, None, seg_off.shape[0], seg_off.ctypes.data,
                                                 seg_cnt.ctypes.data))

    # ---- routed scan: fixed regions in the owners' buffers, no host round trip per step ---------
    def pub_base(self) -> int:
        v = ctypes.c_uint64(0)
        nat.check(lib.pk_indexer_pub_base(self._h, ctypes.byref(v)))
        return int(v.value)

    def set_route(self, nranks: int, self_rank: int, owner: np.ndarray, dest_off: np.ndarray, cap: np.ndarray,
                  pub_base: Sequence[int]) -> None:
        owner = np.ascontiguousarray(owner, dtype=np.uint32)
        dest_off = np.ascontiguousarray(dest_off, dtype=np.uint32)
        cap = np.ascontiguousarray(cap, dtype=np.uint32)
        pb = np.ascontiguousarray(pub_base, dtype=np.uint64)
        assert pb.size == nranks
        nat.check(lib.pk_indexer_set_route(self._h, nranks, self_rank, owner.ctypes.data, dest_off.ctypes.data,
                                           cap.ctypes.data, pb.ctypes.data))

    def scan_routed(self, seq: torch.Tensor, status: torch.Tensor, stream=None) -> None:
        """One asynchronous pass: scan `seq`, store the entries into the owners' regions, publish the
        fill counts; status (int32 CUDA tensor) [0] = 1 if a region overflowed."""
        assert seq.is_cuda and seq.dtype == torch.uint8 and seq.is_contiguous()
        assert status.is_cuda and status.element_size() == 4
        self._keep.append(seq)
        nat.check(lib.pk_indexer_scan_routed(self._h, seq.data_ptr(), seq.numel(), status.data_ptr(),
                                             _stream_ptr(stream)))

    def set_import_layout(self, seg_off: np.ndarray, first_window: int, nwindows_total: int) -> None:
        seg_off = np.ascontiguousarray(seg_off, dtype=np.uint32)
        assert seg_off.ndim == 2
        nat.check(lib.pk_indexer_set_import_layout(self._h, seg_off.shape[0], seg_off.ctypes.data, first_window,
                                                   nwindows_total))

    def import_published(self, stream=None) -> None:
        nat.check(lib.pk_indexer_import_published(self._h, _stream_ptr(stream)))

    PROFILE_CLASSES = ("scan_count_direct", "scan_bucket_count", "bucket_offsets", "scan_scatter",
                       "window_count", "window_commit", "table_stats", "update_carry")

    def set_profiling(self, enable: bool) -> None:
        nat.check(lib.pk_indexer_set_profiling(self._h, 1 if enable else 0))

    def profile(self) -> dict:
        """{kernel class: (total ms, launches)} since profiling was enabled / last read."""
        ms = np.zeros(8, dtype=np.float64)
        cnt = np.zeros(8, dtype=np.uint32)
        nat.check(lib.pk_indexer_profile(self._h, ms.ctypes.data, cnt.ctypes.data))
        return {n: (float(ms[i]), int(cnt[i])) for i, n in enumerate(self.PROFILE_CLASSES) if cnt[i]}

    def launch_count(self) -> int:
        n = ctypes.c_uint64(0)
        nat.check(lib.pk_indexer_launch_count(self._h, ctypes.byref(n)))
        return int(n.value)


def table_stats(table, device: Optional[int] = None):
    """Header.update_stats arithmetic (tools.py:246-263) on the GPU.
    -> (hist list[255], (vals_sum, vals_count, vals_min, vals_max))"""
    t = to_device_u8(table, device)
    hist = np.zeros(255, dtype=np.int64)
    st = np.zeros(4, dtype=np.uint64)
    with torch.cuda.device(t.device):
        nat.check(lib.pk_table_stats_device(t.data_ptr(), t.numel(), hist.ctypes.data,
                                            st.ctypes.data, _stream_ptr()))
    return hist.tolist(), tuple(int(v) for v in st)


def table_pack(table, device: Optional[int] = None):
    """The packed form a finished table crosses PCIe in (include/pykmer_b200.h: pk_table_pack_device):
    -> (bitmap uint64[n/64], chunk_off uint32[n/1024], nz uint8[...]) as NumPy arrays."""
    t = to_device_u8(table, device)
    n = t.numel()
    bitmap = torch.empty(n // 64, dtype=torch.int64, device=t.device)
    chunk_off = torch.empty(n // 1024, dtype=torch.int32, device=t.device)
    nz = torch.empty(n + n // 64 + 16, dtype=torch.uint8, device=t.device)
    units = ctypes.c_uint32(0)
    with torch.cuda.device(t.device):
        nat.check(lib.pk_table_pack_device(t.data_ptr(), n, bitmap.data_ptr(), chunk_off.data_ptr(), nz.data_ptr(),
                                           ctypes.byref(units), _stream_ptr()))
    return (bitmap.cpu().numpy().view(np.uint64), chunk_off.cpu().numpy().view(np.uint32),
            nz[: units.value * 16].cpu().numpy())


def table_unpack(bitmap: np.ndarray, chunk_off: np.ndarray, nz: np.ndarray, n: int, threads: int = 0,
                 out: Optional[np.ndarray] = None) -> np.ndarray:
    """Host side of the packed transfer (pk_table_unpack; no CUDA call): the n table bytes."""
    bitmap = np.ascontiguousarray(bitmap, dtype=np.uint64)
    chunk_off = np.ascontiguousarray(chunk_off, dtype=np.uint32)
    nz = np.ascontiguousarray(nz, dtype=np.uint8)
    if out is None:
        out = np.empty(n, dtype=np.uint8)
    assert out.size == n and out.dtype == np.uint8 and out.flags.c_contiguous
    nat.check(lib.pk_table_unpack(bitmap.ctypes.data, chunk_off.ctypes.data, nz.ctypes.data if nz.size else None,
                                  nz.size, n, out.ctypes.data, threads))
    return out


def pair_counts(s, o, min_count: int = 1, max_count: int = 255, device: Optional[int] = None):
    """Header.calculate_distance arithmetic (tools.py:473-482) on the GPU."""
    a, b = to_device_u8(s, device), to_device_u8(o, device)
    assert a.numel() == b.numel()
    out = np.zeros(3, dtype=np.uint64)
    with torch.cuda.device(a.device):
        nat.check(lib.pk_pair_counts_device(a.data_ptr(), b.data_ptr(), a.numel(), min_count,
                                            max_count, out.ctypes.data, _stream_ptr()))
    return tuple(int(v) for v in out)


def threshold_pack(table: torch.Tensor, min_count: int, max_count: int,
                   out: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
    """uint8 CUDA table -> int32 CUDA bitmask words (bit i&31 of word i>>5)."""
    assert table.is_cuda and table.dtype == torch.uint8 and table.is_contiguous()
    n = table.numel()
    words = (n + 31) // 32
    if out is None:
        out = torch.empty(words, dtype=torch.int32, device=table.device)
    assert out.is_cuda and out.numel() >= words and out.element_size() == 4
    with torch.cuda.device(table.device):
        nat.check(lib.pk_threshold_pack_device(table.data_ptr(), n, min_count, max_count,
                                               out.data_ptr(), _stream_ptr(stream)))
    return out


def gram(bits: torch.Tensor, words: Optional[int] = None, out: Optional[torch.Tensor] = None,
         accumulate: bool = False, stream=None) -> torch.Tensor:
    """bits: (N, stride_words) int32 CUDA -> (N, N) int64 CUDA Gram matrix."""
    assert bits.is_cuda and bits.dim() == 2 and bits.element_size() == 4 and bits.is_contiguous()
    N, stride = bits.shape
    words = stride if words is None else words
    if out is None:
        out = torch.zeros((N, N), dtype=torch.int64, device=bits.device)
        accumulate = False
    with torch.cuda.device(bits.device):
        nat.check(lib.pk_gram_device(bits.data_ptr(), N, words, stride, out.data_ptr(),
                                     1 if accumulate else 0, _stream_ptr(stream)))
    return out


TILED_MAX_SAMPLES = 4096


def tiled_mask_words(words: int, nrows: int) -> int:
    """int32 words of a tiled mask buffer for `nrows` samples of `words` words each."""
    return ((words + 31) // 32) * nrows * 32


def tiled_masks(words: int, nrows: int) -> torch.Tensor:
    """Zeroed tiled mask buffer (include/pykmer_b200.h: word g of sample r at
    [(g // 32) * nrows * 32 + r * 32 + g % 32])."""
    return torch.zeros(max(4, tiled_mask_words(words, nrows)), dtype=torch.int32, device="cuda")


def threshold_pack_tiled(table: torch.Tensor, min_count: int, max_count: int, out: torch.Tensor,
                         row: int, nrows: int, first_word: int = 0, stream=None) -> torch.Tensor:
    """uint8 CUDA table slab of sample `row` -> its words of the tiled mask buffer `out`."""
    assert table.is_cuda and table.dtype == torch.uint8 and table.is_contiguous()
    assert out.is_cuda and out.element_size() == 4 and out.is_contiguous()
    words = (table.numel() + 31) // 32
    assert out.numel() >= tiled_mask_words(first_word + words, nrows)
    with torch.cuda.device(table.device):
        nat.check(lib.pk_threshold_pack_tiled_device(table.data_ptr(), table.numel(), first_word, min_count,
                                                     max_count, out.data_ptr(), row, nrows, _stream_ptr(stream)))
    return out


def gram_tiled(bits: torch.Tensor, nsamples: int, words: int, out: Optional[torch.Tensor] = None,
               accumulate: bool = False, stream=None) -> torch.Tensor:
    """tiled mask buffer -> (N, N) int64 CUDA Gram matrix (any N <= TILED_MAX_SAMPLES; more than 256
    samples run block pair by block pair)."""
    assert bits.is_cuda and bits.element_size() == 4 and bits.is_contiguous()
    assert bits.numel() >= tiled_mask_words(words, nsamples)
    if out is None:
        out = torch.zeros((nsamples, nsamples), dtype=torch.int64, device=bits.device)
        accumulate = False
    with torch.cuda.device(bits.device):
        nat.check(lib.pk_gram_tiled_device(bits.data_ptr(), nsamples, words, out.data_ptr(),
                                           1 if accumulate else 0, _stream_ptr(stream)))
    return out


_F4_EXACT = {}


def gram_tiled_exact(device: Optional[int] = None) -> bool:
    """pk_gram_tiled_exact: does this device accumulate 0/1 FP4 products exactly up to 2^24?  (One
    self-check per device, at first use.)"""
    d = torch.cuda.current_device() if device is None else device
    if d not in _F4_EXACT:
        ok = ctypes.c_int(0)
        nat.check(lib.pk_gram_tiled_exact(d, ctypes.byref(ok)))
        _F4_EXACT[d] = bool(ok.value)
    return _F4_EXACT[d]


def use_tiled_masks(nsamples: int, device: Optional[int] = None) -> bool:
    """The merger's default: tiled masks + the FP4 tensor-core Gram kernel.  Row-major masks and the
    integer kernels (tcgen05 kind::i8 up to 256 samples, AND + popcount beyond) serve a device that
    fails the exactness check, and PYKMER_B200_GRAM=i8|popc (test hook)."""
    import os
    return (nsamples <= TILED_MAX_SAMPLES and os.environ.get("PYKMER_B200_GRAM") is None
            and gram_tiled_exact(device))


def matrix_from_gram(G: np.ndarray) -> np.ndarray:
    """(N, N, 3) uint64: [k, l] = (G[k,k], G[l,l], G[k,l])  (merger.py:175-176)."""
    G = np.asarray(G)
    d = np.diag(G).astype(np.uint64)
    N = G.shape[0]
    m = np.empty((N, N, 3), dtype=np.uint64)
    m[:, :, 0] = d[:, None]
    m[:, :, 1] = d[None, :]
    m[:, :, 2] = G.astype(np.uint64)
    return m


def merge_host(tables: Sequence[np.ndarray], min_count: int = 1, max_count: int = 255,
               device: int = 0) -> np.ndarray:
    """Whole merge from host tables through pk_merge_host -> (N, N, 3) uint64."""
    N = len(tables)
    arrs = [np.ascontiguousarray(t, dtype=np.uint8) for t in tables]
    n = arrs[0].size
    assert all(a.size == n for a in arrs)
    ptrs = (ctypes.c_void_p * N)(*[a.ctypes.data for a in arrs])
    m = np.zeros((N, N, 3), dtype=np.uint64)
    nat.check(lib.pk_merge_host(ptrs, N, n, min_count, max_count, device, m.ctypes.data))
    return m


def synth_table(sample: int, lo: int, hi: int, out: Optional[torch.Tensor] = None,
                device: Optional[int] = None, stream=None) -> torch.Tensor:
    if out is None:
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        out = torch.empty(hi - lo, dtype=torch.uint8, device=dev)
    with torch.cuda.device(out.device):
        nat.check(lib.pk_synth_table_device(out.data_ptr(), sample, lo, hi, _stream_ptr(stream)))
    return out

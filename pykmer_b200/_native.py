"""ctypes binding of libpykmer_b200.so (the C ABI declared in include/pykmer_b200.h).

There is no CPU fallback: if the CUDA library has not been built the import
fails loudly, and every call raises PkError with the library's own message.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpykmer_b200.so")

PK_MODE_AUTO, PK_MODE_DIRECT, PK_MODE_PARTITION, PK_MODE_SCAN = 0, 1, 2, 3

# every symbol include/pykmer_b200.h declares (tests check that all of them resolve)
SYMBOLS = (
    "pk_abi_version", "pk_last_error", "pk_device_count", "pk_device_info",
    "pk_host_alloc", "pk_host_free",
    "pk_indexer_create", "pk_indexer_destroy", "pk_indexer_reset", "pk_indexer_set_records",
    "pk_indexer_append_records", "pk_indexer_flush",
    "pk_indexer_feed_device", "pk_indexer_feed_host", "pk_indexer_sync", "pk_indexer_finalize",
    "pk_indexer_finalize_to_host", "pk_indexer_transfer_stats",
    "pk_indexer_record_flags", "pk_indexer_table_device", "pk_indexer_table_to_host",
    "pk_indexer_launch_count", "pk_indexer_mode", "pk_indexer_window_log2", "pk_indexer_set_profiling", "pk_indexer_profile",
    "pk_indexer_prime", "pk_indexer_scan_result", "pk_indexer_export_segments",
    "pk_indexer_import_segments", "pk_indexer_pool_ipc_handle", "pk_indexer_open_peer_pool",
    "pk_indexer_scan_pass1", "pk_indexer_pass1_counts", "pk_indexer_scan_pass2_remote",
    "pk_indexer_pub_base", "pk_indexer_set_route", "pk_indexer_scan_routed", "pk_indexer_set_import_layout",
    "pk_indexer_import_published",
    "pk_table_stats_device", "pk_table_pack_device", "pk_table_unpack",
    "pk_threshold_pack_device", "pk_gram_device", "pk_threshold_pack_tiled_device", "pk_gram_tiled_device", "pk_gram_tiled_exact", "pk_pair_counts_device", "pk_merge_host",
    "pk_synth_table_device", "pk_bgzf_inflate", "pk_fasta_clean", "pk_bgzf_deflate", "pk_fasta_find_headers",
)


class PkError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libpykmer_b200 error {code}: {message}")
        self.code = code


class PkArgumentError(PkError, ValueError):
    """PK_ERR_ARG: where the reference raises AssertionError / ValueError."""


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m pykmer_b200.build` "
            "(nvcc, sm_100a). pykmer_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, sz, u64, i32 = c.c_void_p, c.c_size_t, c.c_uint64, c.c_int
    L.pk_abi_version.restype = i32
    L.pk_last_error.restype = c.c_char_p
    sig = {
        "pk_device_count": [c.POINTER(i32)],
        "pk_device_info": [i32, c.c_char_p, sz, c.POINTER(i32), c.POINTER(sz), c.POINTER(i32),
                           c.POINTER(i32)],
        "pk_host_alloc": [c.POINTER(vp), sz],
        "pk_host_free": [vp],
        "pk_indexer_create": [c.POINTER(vp), i32, i32, u64, u64, i32],
        "pk_indexer_destroy": [vp],
        "pk_indexer_reset": [vp, vp],
        "pk_indexer_set_records": [vp, vp, sz],
        "pk_indexer_append_records": [vp, vp, sz],
        "pk_indexer_flush": [vp],
        "pk_indexer_feed_device": [vp, vp, sz, vp],
        "pk_indexer_feed_host": [vp, vp, sz],
        "pk_indexer_sync": [vp],
        "pk_indexer_finalize": [vp, vp, vp],
        "pk_indexer_finalize_to_host": [vp, vp, vp, vp],
        "pk_indexer_transfer_stats": [vp, vp],
        "pk_indexer_record_flags": [vp, vp, sz],
        "pk_indexer_table_device": [vp, c.POINTER(vp), c.POINTER(sz)],
        "pk_indexer_table_to_host": [vp, vp, sz, sz],
        "pk_indexer_launch_count": [vp, c.POINTER(u64)],
        "pk_indexer_mode": [vp, c.POINTER(i32), c.POINTER(i32)],
        "pk_indexer_window_log2": [vp, c.POINTER(i32)],
        "pk_indexer_set_profiling": [vp, i32],
        "pk_indexer_profile": [vp, vp, vp],
        "pk_indexer_prime": [vp, vp, sz, u64, vp],
        "pk_indexer_scan_result": [vp, c.POINTER(u64)],
        "pk_indexer_export_segments": [vp, c.POINTER(vp), c.POINTER(c.c_uint32), c.POINTER(c.c_uint32),
                                       vp, vp, sz],
        "pk_indexer_import_segments": [vp, vp, c.c_uint32, vp, vp],
        "pk_indexer_pool_ipc_handle": [vp, vp, c.POINTER(sz)],
        "pk_indexer_open_peer_pool": [vp, i32, vp, vp],
        "pk_indexer_scan_pass1": [vp, vp, sz, vp],
        "pk_indexer_pass1_counts": [vp, vp, sz],
        "pk_indexer_scan_pass2_remote": [vp, i32, vp, vp, vp],
        "pk_indexer_pub_base": [vp, c.POINTER(u64)],
        "pk_indexer_set_route": [vp, i32, i32, vp, vp, vp, vp],
        "pk_indexer_scan_routed": [vp, vp, sz, vp, vp],
        "pk_indexer_set_import_layout": [vp, c.c_uint32, vp, c.c_uint32, c.c_uint32],
        "pk_indexer_import_published": [vp, vp],
        "pk_table_stats_device": [vp, sz, vp, vp, vp],
        "pk_table_pack_device": [vp, sz, vp, vp, vp, c.POINTER(c.c_uint32), vp],
        "pk_table_unpack": [vp, vp, vp, sz, sz, vp, i32],
        "pk_threshold_pack_device": [vp, sz, i32, i32, vp, vp],
        "pk_gram_device": [vp, i32, sz, sz, vp, i32, vp],
        "pk_threshold_pack_tiled_device": [vp, sz, sz, i32, i32, vp, i32, i32, vp],
        "pk_gram_tiled_device": [vp, i32, sz, vp, i32, vp],
        "pk_gram_tiled_exact": [i32, c.POINTER(i32)],
        "pk_pair_counts_device": [vp, vp, sz, i32, i32, vp, vp],
        "pk_merge_host": [c.POINTER(vp), i32, sz, i32, i32, i32, vp],
        "pk_synth_table_device": [vp, i32, u64, u64, vp],
        "pk_bgzf_inflate": [vp, sz, vp, sz, c.POINTER(sz), c.POINTER(sz), i32],
        "pk_fasta_clean": [vp, sz, vp, c.POINTER(sz), c.POINTER(c.c_uint32), i32],
        "pk_bgzf_deflate": [vp, sz, vp, sz, c.POINTER(sz), vp, i32, i32],
        "pk_fasta_find_headers": [vp, sz, vp, sz, c.POINTER(sz), i32],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = i32
    return L


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        msg = (lib.pk_last_error() or b"").decode(errors="replace")
        raise (PkArgumentError if rc == -1 else PkError)(rc, msg)


def device_count() -> int:
    n = ctypes.c_int(0)
    check(lib.pk_device_count(ctypes.byref(n)))
    return n.value


def device_info(device: int = 0) -> dict:
    name = ctypes.create_string_buffer(256)
    sm, maj, mnr = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    mem = ctypes.c_size_t(0)
    check(lib.pk_device_info(device, name, 256, ctypes.byref(sm), ctypes.byref(mem),
                             ctypes.byref(maj), ctypes.byref(mnr)))
    return {"name": name.value.decode(), "sm_count": sm.value, "total_mem": mem.value,
            "cc": (maj.value, mnr.value)}


def ptr(x) -> Optional[int]:
    """Raw address of a torch tensor / numpy array / int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(f"cannot take the address of {type(x)}")

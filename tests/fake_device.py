"""A stand-in for pykmer_b200.device on machines without a GPU -- TEST INFRASTRUCTURE ONLY.

The world_size-2 gloo tests (tests/test_dist_gloo.py) check the HOST logic of the multi-rank
CLIs: who reads, what is broadcast, which slice of the .kin each rank writes, how statistics and
partial matrices meet, that a failure on one rank stops all of them.  That logic calls the device
layer through a handful of names; here those names are backed by the CPU oracle (oracle/), so the
protocol can run under gloo in the build container.  Nothing in the product imports this module.
"""
import numpy as np
import torch
from contextlib import nullcontext

from oracle import oracle


def device_scope(device):
    return nullcontext()


def pinned_empty(nbytes):
    return torch.empty(int(nbytes), dtype=torch.uint8)


def zeros(shape, dtype=torch.int32):
    return torch.zeros(shape, dtype=dtype)


def upload(t, non_blocking=False):
    return t.clone()


def stream_sync():
    pass


class Indexer:
    def __init__(self, kmer_len, device=0, range_lo=0, range_hi=None, mode=0):
        self.K, self.lo = kmer_len, range_lo
        self.hi = 4 ** kmer_len if range_hi is None else range_hi
        self.parts, self.starts, self.table = [], np.zeros(0, dtype=np.uint64), None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        pass

    def set_records(self, starts):
        self.starts = np.asarray(starts, dtype=np.uint64)

    def append_records(self, new_starts):
        self.starts = np.concatenate((self.starts, np.asarray(new_starts, dtype=np.uint64)))

    def feed_device(self, seq, stream=None):
        self.parts.append(seq.numpy().copy())

    def sync(self):
        pass

    def finalize(self, table_out=None):
        seq = np.concatenate(self.parts) if self.parts else np.zeros(0, dtype=np.uint8)
        self.table, num, self.flags = oracle.index_stream(seq, self.K, range_lo=self.lo, range_hi=self.hi,
                                                          rec_starts=self.starts)
        hist, st = oracle.table_stats(self.table)
        return hist, {"num_kmers": num, "vals_sum": st["vals_sum"], "vals_count": st["vals_count"],
                      "vals_min": st["vals_min"], "vals_max": st["vals_max"]}

    def record_flags(self):
        return np.asarray(self.flags, dtype=np.uint8)

    def table_to_host(self, dst=None, offset=0, nbytes=None):
        nbytes = self.table.size - offset if nbytes is None else nbytes
        dst.numpy()[:nbytes] = self.table[offset:offset + nbytes]
        return dst


def use_tiled_masks(nsamples, device=None):
    return False


def free_memory_bytes():
    return 1 << 30


def threshold_pack(table, min_count, max_count, out=None, stream=None):
    words = torch.from_numpy(oracle.threshold_pack(table.numpy(), min_count, max_count).view(np.int32))
    out.view(-1)[:words.numel()] = words
    return out


def gram(bits, words=None, out=None, accumulate=False, stream=None):
    b = bits.numpy().view(np.uint32)[:, :bits.shape[1] if words is None else words]
    G = torch.from_numpy(oracle.gram_from_bits(np.ascontiguousarray(b)).astype(np.int64))
    if out is None:
        return G
    if accumulate:
        out += G
    else:
        out.copy_(G)
    return out


def matrix_from_gram(G):
    return oracle.matrix_from_gram(G)

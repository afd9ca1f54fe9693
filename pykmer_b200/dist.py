"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  Both hot paths shard along the canonical k-mer axis
(SURVEY.md 8e); the only exchanges are

  indexer  one all-reduce of hist[255] + num_kmers / vals_sum / vals_count (sum) and of
           vals_min / vals_max (min / max)  -- about 2 KB;
  merger   one all-reduce (sum) of the N x N int64 partial Gram matrices.

The sequence itself is either read by every rank or broadcast once from rank 0
(`broadcast_stream`).  Nothing here touches the data path of a single GPU.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

ALIGN = 4096          # shard boundaries are multiples of this many table entries


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total: int, rank: int, nranks: int, align: int = ALIGN) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the k-mer axis owned by `rank`; slices tile [0, total)."""
    def cut(r: int) -> int:
        if r <= 0:
            return 0
        if r >= nranks:
            return total
        return (total * r // nranks) // align * align
    return cut(rank), cut(rank + 1)


def _device_for_backend() -> torch.device:
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def reduce_index_stats(hist: List[int], st: Dict[str, int], group=None) -> Tuple[List[int], Dict[str, int]]:
    """Combine per-shard statistics into those of the whole table (tools.py:246-263):
    hist / num_kmers / vals_sum / vals_count add up, vals_min / vals_max are min / max."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(hist), dict(st)
    dev = _device_for_backend()
    sums = torch.tensor(list(hist) + [st["num_kmers"], st["vals_sum"], st["vals_count"]],
                        dtype=torch.int64, device=dev)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    mm = torch.tensor([-st["vals_min"], st["vals_max"]], dtype=torch.int64, device=dev)
    dist.all_reduce(mm, op=dist.ReduceOp.MAX, group=group)
    s = sums.cpu().tolist()
    m = mm.cpu().tolist()
    return s[:255], {"num_kmers": s[255], "vals_sum": s[256], "vals_count": s[257],
                     "vals_min": -m[0], "vals_max": m[1]}


def reduce_flags(flags: np.ndarray, group=None) -> np.ndarray:
    """A record is listed if ANY shard counted one of its k-mers (indexer.py:349-351)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flags
    t = torch.from_numpy(flags.astype(np.int32)).to(_device_for_backend())
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.cpu().numpy().astype(np.uint8)


def reduce_gram(G: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the partial Gram matrices of the k-mer-axis shards, in place."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)
    return G


def broadcast_stream(chunk: Optional[torch.Tensor], nbytes: int, src: int = 0, group=None) -> torch.Tensor:
    """Broadcast one chunk of the cleaned sequence stream from `src` (NCCL: over NVLink).
    Ranks other than `src` pass chunk=None and receive a fresh uint8 tensor of nbytes."""
    dev = _device_for_backend()
    if chunk is None:
        chunk = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(chunk, src=src, group=group)
    return chunk

"""Multi-GPU sharding of C = A*B: the host-side logic around the engine's panel entry points.

The path shards by independent units: every tile row of C depends only on the same tile row of A
and on (read-only, replicated) B.  A is cut into contiguous tile-row panels with balanced work (flop + tile products)
(`pem_partition_panels`, device-side flop count), each rank multiplies its panel with
`pem_spgemm_panel`, and the ONLY exchange is an all-gather of three integers per rank
{nnz(C shard), C' tiles, pairs}, whose exclusive scan gives every shard's offset in the global C
(north_star: "NCCL used only to gather per-shard C nnz and offsets").  The reference has no
multi-GPU code.

Works with any torch.distributed backend: NCCL on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class ShardLayout:
    rank: int
    world: int
    sizes: np.ndarray      # int64[world, 3]: nnz, tiles, pairs of every shard
    offsets: np.ndarray    # int64[world, 3]: exclusive scan over ranks (this shard's start in global C)
    totals: np.ndarray     # int64[3]

    @property
    def nnz_offset(self) -> int:
        return int(self.offsets[self.rank, 0])


def layout_from_sizes(sizes: np.ndarray, rank: int) -> ShardLayout:
    sizes = np.asarray(sizes, np.int64).reshape(-1, 3)
    offsets = np.cumsum(sizes, axis=0) - sizes
    return ShardLayout(rank, sizes.shape[0], sizes, offsets, sizes.sum(axis=0))


_xbuf = {}


def exchange_shard_sizes(nnz: int, tiles: int, pairs: int, device=None) -> ShardLayout:
    """All-gather this rank's {nnz, tiles, pairs}; returns every shard's sizes and global offsets.
    Without an initialised process group this is the single-shard layout.  On CUDA devices the three integers
    travel through persistent pinned / device buffers (no pageable staging copy, one stream synchronisation)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return layout_from_sizes(np.array([[nnz, tiles, pairs]], np.int64), 0)
    world, rank = dist.get_world_size(), dist.get_rank()
    if device is not None and torch.device(device).type == "cuda":
        key = (str(device), world)
        if key not in _xbuf:
            _xbuf[key] = (torch.empty(3, dtype=torch.int64, pin_memory=True), torch.empty(3, dtype=torch.int64, device=device),
                          torch.empty(world * 3, dtype=torch.int64, device=device),
                          torch.empty(world * 3, dtype=torch.int64, pin_memory=True))
        h_mine, d_mine, d_all, h_all = _xbuf[key]
        h_mine[0], h_mine[1], h_mine[2] = nnz, tiles, pairs
        d_mine.copy_(h_mine, non_blocking=True)
        dist.all_gather_into_tensor(d_all, d_mine)
        h_all.copy_(d_all, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return layout_from_sizes(h_all.numpy().copy(), rank)
    mine = torch.tensor([nnz, tiles, pairs], dtype=torch.int64, device=device)
    allsz = torch.empty(world * 3, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allsz, mine)
    return layout_from_sizes(allsz.cpu().numpy(), rank)


def split_by_weight(weights: np.ndarray, nparts: int) -> np.ndarray:
    """Host restatement of pem_partition_panels' rule, for tests and for hosts that already hold the
    per-tile-row flop counts: contiguous panels, panel p closes at the first row where the running
    weight (each row counted as weight+1) reaches p/nparts of the total."""
    w = np.asarray(weights, np.int64) + 1
    total = int(w.sum())
    run = np.cumsum(w)
    bounds = np.zeros(nparts + 1, np.int32)
    bounds[nparts] = w.size
    for p in range(1, nparts):
        # first r with run[r] * nparts >= total * p
        bounds[p] = min(int(np.searchsorted(run * nparts, total * p, side="left")) + 1, w.size)
    return np.maximum.accumulate(bounds)


def upload_coo_sharded(I, J, V, device):
    """Multi-GPU ingest of one host COO: every rank uploads only its 1/world slice over its own PCIe
    link (pinned host -> device) and the slices are all-gathered over NVLink, so each GPU ends up with
    the full COO (B is replicated, A's panel is cut later by tile row) after 1/world of the host->device
    traffic.  I/J/V are pinned torch tensors holding the FULL arrays on every rank.  Returns device
    tensors (I, J, V) of the full length and the number of bytes this rank copied from the host."""
    import torch
    import torch.distributed as dist
    n = I.numel()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = [t.to(device, non_blocking=True) for t in (I, J, V)]
        return out[0], out[1], out[2], sum(t.numel() * t.element_size() for t in (I, J, V))
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-n // world)
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
    full, copied = [], 0
    for t in (I, J, V):
        g = torch.empty(world * per, dtype=t.dtype, device=device)
        mine = g[rank * per: (rank + 1) * per]          # all_gather in place: own slice lives inside the output
        mine[: hi - lo].copy_(t[lo:hi], non_blocking=True)
        copied += (hi - lo) * t.element_size()
        dist.all_gather_into_tensor(g, mine)
        full.append(g[:n])                               # only the last slice is padded, and the padding is the tail
    return full[0], full[1], full[2], copied

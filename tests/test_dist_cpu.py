"""CPU tests of the N>1 host logic: world_size-2 gloo all-gather of shard sizes -> global offsets,
and the panel split rule."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from pem_spgemm_b200 import dist as pdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sizes = [(1000 + rank, 10 + rank, 20 * (rank + 1)) for _ in range(1)][0]
        lay = pdist.exchange_shard_sizes(*sizes)
        q.put((rank, lay.sizes.tolist(), lay.offsets.tolist(), lay.totals.tolist(), lay.nnz_offset))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_exchange_of_shard_sizes(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_sizes = [[1000 + r, 10 + r, 20 * (r + 1)] for r in range(world)]
    want_off = (np.cumsum(want_sizes, axis=0) - np.array(want_sizes)).tolist()
    for rank, sizes, offs, totals, nnz_off in out:
        assert sizes == want_sizes and offs == want_off
        assert totals == np.sum(want_sizes, axis=0).tolist()
        assert nnz_off == want_off[rank][0]


def test_single_process_layout():
    lay = pdist.exchange_shard_sizes(7, 3, 5)
    assert lay.world == 1 and lay.nnz_offset == 0 and lay.totals.tolist() == [7, 3, 5]


def test_split_by_weight_is_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    w = rng.integers(0, 1000, size=5000)
    w[:50] *= 40                       # heavy head, as in the hub-first webbase-like matrix
    for n in (1, 2, 4, 8):
        b = pdist.split_by_weight(w, n)
        assert b[0] == 0 and b[-1] == w.size and np.all(np.diff(b) >= 0)
        part = np.add.reduceat(w + 1, b[:-1])[: n]
        assert part.max() <= (w + 1).sum() / n + (w + 1).max()
    assert pdist.split_by_weight(np.zeros(0, np.int64), 3).tolist() == [0, 0, 0, 0]


def _worker_upload(rank, world, port, q):
    import torch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1001                                    # not a multiple of the world size: last slice is padded
        I = torch.arange(n, dtype=torch.int32); J = I * 3; V = I.to(torch.float64) * 0.5
        dI, dJ, dV, copied = pdist.upload_coo_sharded(I, J, V, torch.device("cpu"))
        ok = bool(torch.equal(dI, I) and torch.equal(dJ, J) and torch.equal(dV, V))
        q.put((rank, ok, int(copied)))
    finally:
        dist.destroy_process_group()


def test_gloo_sharded_upload_reassembles_the_coo():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_upload, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in out)
    assert sum(c for _, _, c in out) == 1001 * 16      # every byte of the COO crossed a host link exactly once

"""CPU: host logic of the multi-GPU path — tile planning (pure arithmetic in the C library) and the neighbour
exchange / marking chain of bcd_b200/tiles.py over torch.distributed with the gloo backend, world_size 2 and 3."""
import os
import socket

import numpy as np
import pytest


def test_plan_covers_frame_with_aligned_bands():
    from bcd_b200 import capi, tiles
    for (W, H, n, s) in [(1920, 1080, 8, 3), (4096, 4096, 8, 3), (640, 481, 3, 2), (320, 200, 1, 3), (512, 1024, 4, 1)]:
        p = capi.make_params(scales=s)
        ts = [tiles.plan(W, H, i, n, p) for i in range(n)]
        align, halo = 1 << (s - 1), (p.search_radius + p.patch_radius) << (s - 1)
        assert ts[0].own_row0 == 0 and ts[-1].own_row1 == H
        for a, b in zip(ts, ts[1:]):
            assert a.own_row1 == b.own_row0 and a.own_row1 % align == 0
        for t in ts:
            assert t.sub_row0 == max(0, t.own_row0 - halo) and t.sub_row1 == min(H, t.own_row1 + halo)
            assert t.sub_row0 % align == 0
            assert ((t.own_row1 - t.own_row0) >> (s - 1)) >= p.search_radius + p.patch_radius + 2 or n == 1
        rows = [t.own_row1 - t.own_row0 for t in ts]
        assert max(rows) - min(rows) <= 2 * align


def test_plan_rejects_bad_requests():
    from bcd_b200 import capi, tiles
    with pytest.raises(capi.BcdError):
        tiles.plan(64, 64, 0, 8, capi.make_params(scales=3))
    with pytest.raises(capi.BcdError):
        tiles.plan(64, 64, 3, 2, capi.make_params())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _exchange_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from bcd_b200 import tiles
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex = tiles.DistExchange(dist, rank, world)
        # --- accumulator halo: every rank holds partial sums on its band +- 3 rows; after the swap+add the owned rows
        #     must equal the sum over ranks of the global partial images
        H, W, halo = 12 * world, 5, 3
        rng = np.random.default_rng(123)
        partial = rng.standard_normal((world, H, W)).astype(np.float32)      # rank r's contributions to every row
        for r in range(world):                                              # rank r only touches its band +- halo
            lo, hi = 12 * r - halo, 12 * (r + 1) + halo
            partial[r, :max(lo, 0)] = 0
            partial[r, min(hi, H):] = 0
        mine = torch.from_numpy(partial[rank].copy())
        own0, own1 = 12 * rank, 12 * (rank + 1)
        send = {tiles.UP: mine[own0 - halo:own0].contiguous() if rank > 0 else None,
                tiles.DOWN: mine[own1:own1 + halo].contiguous() if rank < world - 1 else None}
        recv = ex.swap(send)
        if recv[tiles.UP] is not None:
            mine[own0:own0 + halo] += recv[tiles.UP]
        if recv[tiles.DOWN] is not None:
            mine[own1 - halo:own1] += recv[tiles.DOWN]
        ok_acc = np.allclose(mine[own0:own1].numpy(), partial.sum(0)[own0:own1], atol=1e-6)
        # --- marking chain: rank r must see what rank r-1 produced, in order
        token = torch.zeros(4, dtype=torch.int64)
        if rank > 0:
            token = ex.recv_from_up(token)
        token = token + torch.tensor([1, rank, rank * rank, 7])
        if rank < world - 1:
            ex.send_down(token)
        q.put((rank, bool(ok_acc), token.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_neighbour_exchange_and_marking_chain_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_acc, token in res:
        assert ok_acc
        assert token == [rank + 1, sum(range(rank + 1)), sum(i * i for i in range(rank + 1)), 7 * (rank + 1)]

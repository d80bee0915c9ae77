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


# ---------------------------------------------------------------------------------------------------------------------
# run_tile's sequencing (bcd_b200/tiles.py) with a CPU stand-in for the tile worker: every rank logs the stages, the
# strips are small tensors whose content says who produced them for which level. Checks over gloo (world 2 and 3): no
# deadlock, every per-level ordering rule of the row-tile API, the marking chain arrives in rank order on every level,
# and every received strip is the one the neighbour exported for that (level, kind, exchange number).
class _FakeParams:
    def __init__(self, s, m):
        self.nb_scales, self.skip_prob = s, m


class _FakeWorker:
    def __init__(self, rank, world, s, m):
        import torch
        from bcd_b200 import tiles
        self.torch, self.tiles = torch, tiles
        self.rank, self.world = rank, world
        self.params = _FakeParams(s, m)
        self.has = {tiles.UP: rank > 0, tiles.DOWN: rank < world - 1}
        self.main_stream, self.aux_streams, self.device = None, [None] * s, "cpu"
        self.log, self.chain, self.received, self.nexport = [], {}, [], {}

    def begin(self): self.log.append(("begin", -1))
    def end(self): self.log.append(("end", -1))

    def stage(self, name, level):
        self.log.append((name, level))
        if name == "mark_enqueue":  # this rank's link: what came from above + own contribution
            self.chain[level] = self.chain.get(level, 0) + 10 ** self.rank

    def strip_bytes(self, level, what): return 8

    def export(self, level, what, side):
        if what == self.tiles.AS_ROWS:
            return self.torch.tensor([self.chain[level], level], dtype=self.torch.int64)
        n = self.nexport.get((level, what, side), 0)  # exchange number of this (level, kind, side)
        self.nexport[(level, what, side)] = n + 1
        return self.torch.tensor([self.rank, level, what, n, side], dtype=self.torch.int64)

    def import_(self, level, what, side, buf):
        if what == self.tiles.AS_ROWS:
            assert int(buf[1]) == level
            self.chain[level] = int(buf[0])
        else:
            self.received.append((level, what, side, buf.tolist()))


class _FakeRecvEx:
    """DistExchange whose chain receive buffer has the fake strip's shape."""
    def __init__(self, ex):
        self.ex = ex
    def swap(self, send): return self.ex.swap(send)
    def send_down(self, buf): return self.ex.send_down(buf)
    def recv_from_up(self, like):
        import torch
        return self.ex.recv_from_up(torch.zeros(2, dtype=torch.int64))


def _sequencing_worker(rank, world, port, s, m, q):
    import torch.distributed as dist
    from bcd_b200 import tiles
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = _FakeWorker(rank, world, s, m)
        tiles.run_tile(w, _FakeRecvEx(tiles.DistExchange(dist, rank, world)))
        q.put((rank, w.log, w.chain, w.received))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,s,m", [(2, 3, 1.0), (3, 2, 1.0), (2, 2, 0.0)])
def test_run_tile_sequencing_gloo(world, s, m):
    import torch.multiprocessing as mp
    from bcd_b200 import tiles
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sequencing_worker, args=(r, world, port, s, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, log, chain, received in res:
        pos = {e: i for i, e in enumerate(log)}
        assert log[0] == ("begin", -1) and log[-1] == ("end", -1)
        for k in range(s):
            order = [pos[(n, k)] for n in ("select", "mark_enqueue", "mark_wait", "filter", "finalize")]
            assert order == sorted(order)
            if k < s - 1:  # merge(k) needs both levels finalised and the coarser one merged
                assert pos[("merge", k)] > pos[("finalize", k)] and pos[("merge", k)] > pos[("finalize", k + 1)]
                if k < s - 2:
                    assert pos[("merge", k)] > pos[("merge", k + 1)]
        assert ("merge", s - 1) not in pos
        # marking chain: rank r holds the contributions of ranks 0..r (m > 0), or only its own (m = 0: no chain)
        want = sum(10 ** i for i in range(rank + 1)) if m > 0 else 10 ** rank
        assert chain == {k: want for k in range(s)}
        # every strip came from the right neighbour, for the same level / kind / exchange number, facing us
        seen = {}
        for level, what, side, payload in received:
            n = seen.get((level, what, side), 0)
            seen[(level, what, side)] = n + 1
            src = rank - 1 if side == tiles.UP else rank + 1
            assert payload == [src, level, what, n, tiles.DOWN if side == tiles.UP else tiles.UP]
        for k in range(s):
            sides = [sd for sd, has in ((tiles.UP, rank > 0), (tiles.DOWN, rank < world - 1)) if has]
            for sd in sides:
                assert seen[(k, tiles.ACC_HALO, sd)] == 1
                assert seen[(k, tiles.OUT_ROWS, sd)] == (2 if k < s - 1 else 1)

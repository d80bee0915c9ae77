"""Python side of the row-tiled multi-GPU path.

The driver itself is C++ (bcd_b200/csrc/bcd_b200_multi.cu): one handle and one host thread per tile, strips exchanged
through peer memory over NVLink, nothing between the stages of a frame on the host. What lives here:

  * `plan`, `peer_setup / peer_connect / peer_run`, `peer_connect_dist`  — thin ctypes wrappers of bcd_b200_tile_plan /
    _tile_setup / _tile_connect / _tile_run, and the one-time all_gather of the tiles' 96-byte arena tokens;
  * `bench_tiled`  — bench.py's N > 1 arm: one process per GPU (torchrun), a ranked bcd_b200_group per process;
  * the round-1 exchange, kept as the portable fallback and A/B twin (`BCD_B200_EXCHANGE=nccl`): `TileWorker`,
    `DistExchange`, `run_tile`, `run_tiles_loopback` sequence the per-stage entry points (tile_select, tile_mark, ...)
    and move the three kinds of strips with torch.distributed send/recv (NCCL on GPUs, gloo in the CPU tests):

        per level, finest first:      select;  [m > 0] AS rows: recv from rank-1 -> mark -> send to rank+1
        per level, 0 then s-1 .. 1:   wait for the work counters; filter; accumulator halo swap + add; finalize; output rows
        per level s-2 .. 0:           merge; output rows again
"""
import ctypes as C

import numpy as np

from . import capi

ACC_HALO, AS_ROWS, OUT_ROWS = 0, 1, 2
UP, DOWN = 0, 1


class Tile(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("image_width", "image_height", "own_row0", "own_row1", "sub_row0", "sub_row1",
                                         "tile_index", "tile_count")]


def _lib():
    L = capi.lib()
    if not getattr(L, "_tiles_ready", False):
        L.bcd_b200_tile_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(capi.Params), C.POINTER(Tile)]
        L.bcd_b200_tile_begin.argtypes = [C.c_void_p, C.POINTER(capi.Params), C.POINTER(Tile)]
        L.bcd_b200_set_level_aux_stream.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        for n in ("select", "mark", "mark_enqueue", "mark_wait", "filter", "finalize", "merge"):
            getattr(L, "bcd_b200_tile_" + n).argtypes = [C.c_void_p, C.c_int]
        L.bcd_b200_tile_xfer_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.bcd_b200_tile_export.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.bcd_b200_tile_import.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.bcd_b200_tile_end.argtypes = [C.c_void_p]
        L.bcd_b200_tile_download.argtypes = [C.c_void_p, C.c_void_p]
        L.bcd_b200_tile_setup.argtypes = [C.c_void_p, C.POINTER(capi.Params), C.POINTER(Tile), C.c_int, C.c_void_p]
        L.bcd_b200_tile_connect.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.bcd_b200_tile_run.argtypes = [C.c_void_p, C.POINTER(capi.Params)]
        L._tiles_ready = True
    return L


def make_aux_streams(torch_device, levels):
    """One high-priority torch stream per pyramid level for the marking chains."""
    import torch
    return [torch.cuda.Stream(device=torch_device, priority=-1) for _ in range(levels)]


def plan(W, H, index, count, params):
    """Rows of tile `index` of `count` (pure host arithmetic, no GPU needed) -> Tile."""
    t = Tile()
    capi._check(_lib().bcd_b200_tile_plan(W, H, index, count, C.byref(params), C.byref(t)))
    return t


class TileWorker:
    """One tile = one capi.Handle in row-tile mode. Strips travel as 1-D uint8 torch tensors on `device`."""

    def __init__(self, handle, params, tile, torch_device, main_stream=None, aux_stream=None):
        """main_stream / aux_stream: torch.cuda.Stream objects shared with the exchange (NCCL orders its transfers
        against the current torch stream, so the strips of the marking chain of level k must be moved under
        `aux_streams[k]`). aux_stream: one stream (all levels share it), a list with one stream per level, or None
        (one high-priority stream per level is created)."""
        import torch
        self.torch = torch
        self.h, self.params, self.tile, self.device = handle, params, tile, torch_device
        self.has = {UP: tile.tile_index > 0, DOWN: tile.tile_index < tile.tile_count - 1}
        self.main_stream = main_stream if main_stream is not None else torch.cuda.Stream(device=torch_device)
        if aux_stream is None:
            aux_stream = make_aux_streams(torch_device, params.nb_scales)
        elif not isinstance(aux_stream, (list, tuple)):
            aux_stream = [aux_stream] * params.nb_scales
        self.aux_streams = list(aux_stream)
        self.aux_stream = self.aux_streams[0]
        handle.set_streams(self.main_stream.cuda_stream, self.aux_stream.cuda_stream)
        for k, st in enumerate(self.aux_streams):
            capi._check(_lib().bcd_b200_set_level_aux_stream(handle._h, k, C.c_void_p(st.cuda_stream)))

    def begin(self):
        capi._check(_lib().bcd_b200_tile_begin(self.h._h, C.byref(self.params), C.byref(self.tile)))

    def stage(self, name, level):
        capi._check(getattr(_lib(), "bcd_b200_tile_" + name)(self.h._h, level))

    def strip_bytes(self, level, what):
        n = C.c_size_t()
        capi._check(_lib().bcd_b200_tile_xfer_bytes(self.h._h, level, what, C.byref(n)))
        return n.value

    def export(self, level, what, side):
        buf = self.torch.empty(self.strip_bytes(level, what), dtype=self.torch.uint8, device=self.device)
        capi._check(_lib().bcd_b200_tile_export(self.h._h, level, what, side, C.c_void_p(buf.data_ptr())))
        return buf

    def import_(self, level, what, side, buf):
        capi._check(_lib().bcd_b200_tile_import(self.h._h, level, what, side, C.c_void_p(buf.data_ptr())))

    def end(self):
        capi._check(_lib().bcd_b200_tile_end(self.h._h))

    def download(self, out=None):
        rows = self.tile.own_row1 - self.tile.own_row0
        if out is None:
            out = np.empty((rows, self.tile.image_width, 3), np.float32)
        capi._check(_lib().bcd_b200_tile_download(self.h._h, C.c_void_p(out.ctypes.data if isinstance(out, np.ndarray) else out)))
        return out


# ------------------------------------------------------------------------------------------ peer mode (no host hop)
TOKEN_BYTES = 96


def peer_setup(handle, params, tile, hist_depth=60):
    """bcd_b200_tile_setup: sizes the tile's buffers and returns the 96-byte token of its arena."""
    tok = C.create_string_buffer(TOKEN_BYTES)
    capi._check(_lib().bcd_b200_tile_setup(handle._h, C.byref(params), C.byref(tile), hist_depth, tok))
    return tok.raw


def peer_connect(handle, side, token):
    """bcd_b200_tile_connect: token = the neighbour's arena token (bytes), or None to disconnect."""
    buf = C.create_string_buffer(token, TOKEN_BYTES) if token is not None else None
    capi._check(_lib().bcd_b200_tile_connect(handle._h, side, buf))


def peer_run(handle, params=None):
    """bcd_b200_tile_run: the whole tile schedule; strips travel through peer memory, the host never waits in between."""
    capi._check(_lib().bcd_b200_tile_run(handle._h, C.byref(params) if params is not None else None))


def peer_connect_dist(handle, params, tile, dist, rank, world, hist_depth=60):
    """One tile per rank: setup, all_gather of the arena tokens (which is also the barrier tile_run needs: every rank
    has finished its setup), connect to the two neighbours (CUDA IPC)."""
    tok = peer_setup(handle, params, tile, hist_depth)
    gathered = [None] * world
    dist.all_gather_object(gathered, tok)
    if rank > 0:
        peer_connect(handle, UP, gathered[rank - 1])
    if rank < world - 1:
        peer_connect(handle, DOWN, gathered[rank + 1])
    dist.barrier()


# ------------------------------------------------------------------------------------------------------ exchanges
class DistExchange:
    """Neighbour exchange over torch.distributed (one tile per rank)."""

    def __init__(self, dist, rank, world):
        self.dist, self.rank, self.world = dist, rank, world

    def swap(self, send):
        """send: {UP: tensor|None, DOWN: tensor|None} -> received {UP: tensor|None, DOWN: tensor|None}."""
        import torch
        ops, recv = [], {UP: None, DOWN: None}
        for side, peer in ((UP, self.rank - 1), (DOWN, self.rank + 1)):
            if send.get(side) is None or peer < 0 or peer >= self.world:
                continue
            recv[side] = torch.empty_like(send[side])
            ops.append(self.dist.P2POp(self.dist.isend, send[side], peer))
            ops.append(self.dist.P2POp(self.dist.irecv, recv[side], peer))
        if ops:
            for r in self.dist.batch_isend_irecv(ops):
                r.wait()
        return recv

    # The links of the marking chain go through batch_isend_irecv as well: with an eagerly initialised NCCL group
    # torch treats *unbatched* send/recv as collectives and serialises them with every other op of the group, which
    # would chain the levels' marking streams together. wait() orders the current stream (NCCL) / blocks (gloo).
    def recv_from_up(self, like):
        for r in self.dist.batch_isend_irecv([self.dist.P2POp(self.dist.irecv, like, self.rank - 1)]):
            r.wait()
        return like

    def send_down(self, buf):
        for r in self.dist.batch_isend_irecv([self.dist.P2POp(self.dist.isend, buf, self.rank + 1)]):
            r.wait()


def level_order(s):
    """Order in which the levels are filtered: level 0 first (its marking chain was started first and its filter is the
    longest, so the chains of the coarser levels finish under it), then coarsest .. 1."""
    return [0] + list(range(s - 1, 0, -1))


def run_tile(worker, ex):
    """Drives one tile through all levels with a DistExchange (every rank calls this)."""
    import torch
    p = worker.params
    s = p.nb_scales
    with torch.cuda.stream(worker.main_stream):
        worker.begin()
        for k in range(s):  # selection + this rank's link of the marking chain, finest level first
            worker.stage("select", k)
            with torch.cuda.stream(worker.aux_streams[k]):
                if p.skip_prob > 0 and worker.has[UP]:
                    buf = torch.empty(worker.strip_bytes(k, AS_ROWS), dtype=torch.uint8, device=worker.device)
                    worker.import_(k, AS_ROWS, UP, ex.recv_from_up(buf))
                worker.stage("mark_enqueue", k)
                if p.skip_prob > 0 and worker.has[DOWN]:
                    ex.send_down(worker.export(k, AS_ROWS, DOWN))
        for k in level_order(s):
            worker.stage("mark_wait", k)  # host: this rank's work counters of the level
            worker.stage("filter", k)
            _swap_and_import(worker, ex, k, ACC_HALO)
            worker.stage("finalize", k)
            _swap_and_import(worker, ex, k, OUT_ROWS)
        for k in range(s - 2, -1, -1):
            worker.stage("merge", k)
            _swap_and_import(worker, ex, k, OUT_ROWS)
        worker.end()


def _swap_and_import(worker, ex, level, what):
    send = {side: (worker.export(level, what, side) if worker.has[side] else None) for side in (UP, DOWN)}
    recv = ex.swap(send)
    for side in (UP, DOWN):
        if recv[side] is not None:
            worker.import_(level, what, side, recv[side])


def run_tiles_loopback(workers):
    """All tiles of a frame inside this process (e.g. several handles on one GPU): same stage order as run_tile,
    strips handed over directly. Used by the single-GPU parity test of the tiling."""
    p = workers[0].params
    s = p.nb_scales
    n = len(workers)

    def swap_all(level, what):
        sent = [{side: (w.export(level, what, side) if w.has[side] else None) for side in (UP, DOWN)} for w in workers]
        for i, w in enumerate(workers):
            if w.has[UP]:
                w.import_(level, what, UP, sent[i - 1][DOWN])
            if w.has[DOWN]:
                w.import_(level, what, DOWN, sent[i + 1][UP])

    for w in workers:
        w.begin()
    for k in range(s):
        for w in workers:
            w.stage("select", k)
        for i, w in enumerate(workers):  # the marking chain, top to bottom
            with w.torch.cuda.stream(w.aux_streams[k]):
                if p.skip_prob > 0 and i > 0:
                    w.import_(k, AS_ROWS, UP, workers[i - 1].export(k, AS_ROWS, DOWN))
                w.stage("mark_enqueue", k)
    for k in level_order(s):
        for w in workers:
            w.stage("mark_wait", k)
            w.stage("filter", k)
        swap_all(k, ACC_HALO)
        for w in workers:
            w.stage("finalize", k)
        swap_all(k, OUT_ROWS)
    for k in range(s - 2, -1, -1):
        for w in workers:
            w.stage("merge", k)
        swap_all(k, OUT_ROWS)
    for w in workers:
        w.end()
    assert n == workers[0].tile.tile_count


def make_worker(inputs_full, params, index, count, device_index=0, main_stream=None, aux_stream=None, streamed=False):
    """Uploads the rows of tile `index` of a full-frame dict(color,nsamples,hist,cov) and returns its TileWorker.
    Workers that exchange strips in one process must share their two streams (explicit torch.cuda.Stream objects)."""
    import torch
    H, W, _ = inputs_full["color"].shape
    t = plan(W, H, index, count, params)
    sub = {k: np.ascontiguousarray(v[t.sub_row0:t.sub_row1]) for k, v in inputs_full.items()}
    h = capi.Handle(device_index)
    worker = TileWorker(h, params, t, torch.device("cuda", device_index), main_stream, aux_stream)
    if streamed:
        h.upload_async(**sub)  # bcd_b200_upload_async: the tile's level-0 selection runs behind the arriving rows
    else:
        h.upload(**sub)
    return worker


# ---------------------------------------------------------------------------------------------------------- bench
def bench_tiled(a, rank, world, local_rank, W, H, config, workload, dist, saved_stdout_fd=None):
    """bench.py's N > 1 arm: one process per GPU, each holding its share of the row tiles in a ranked group
    (bcd_b200_group_*: tile j on GPU j mod N, linked through CUDA IPC arenas over NVLink); barrier + synchronize on both
    sides, max over ranks. BCD_B200_EXCHANGE=nccl selects the round-1 exchange (one tile per rank, torch.distributed
    send/recv per strip) for A/B runs."""
    import json
    import os
    import sys
    import time
    import numpy as np
    import torch
    from . import host
    import bench as B  # clocks sampler / constants of the N=1 arm

    dev = torch.device("cuda", local_rank)
    params = capi.make_params(w=a.w, b=a.b, m=a.m, scales=a.scales)
    Bd = 60
    use_nccl = os.environ.get("BCD_B200_EXCHANGE", "peer") == "nccl"
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    if use_nccl:
        infos = [plan(W, H, rank, world, params)]
        infos = [dict(own_row0=t.own_row0, own_row1=t.own_row1, sub_row0=t.sub_row0, sub_row1=t.sub_row1, tile_index=rank) for t in infos]
        g = None
    else:
        g = capi.Group([local_rank], rank=rank, world=world, tiles_per_gpu=a.tiles_per_gpu)
        mine = g.setup(W, H, Bd, params)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)  # also the barrier between every process' setup and the first run
        tokens = {}
        for part in gathered:
            for idx, tok in part:
                tokens[idx] = tok
        g.connect([tokens[j] for j in range(len(tokens))])
        infos = [g.tile_info(i) for i in range(g.local_tiles())]
        dist.barrier()

    # ---- inputs of the local tiles, generated straight into pinned host memory
    t0 = time.time()
    pins, outs = [], []
    for t in infos:
        rows = t["sub_row1"] - t["sub_row0"]
        pin = dict(color=torch.empty((rows, W, 3), dtype=torch.float32).pin_memory(),
                   nsamples=torch.empty((rows, W, 1), dtype=torch.float32).pin_memory(),
                   hist=torch.empty((rows, W, Bd), dtype=torch.float32).pin_memory(),
                   cov=torch.empty((rows, W, 6), dtype=torch.float32).pin_memory())
        host.synth(W, H, spp=a.spp, row0=t["sub_row0"], rows=rows, out={k: v.numpy() for k, v in pin.items()})
        pins.append(pin)
        outs.append(torch.empty((t["own_row1"] - t["own_row0"], W, 3), dtype=torch.float32).pin_memory())
    gen_s = time.time() - t0

    if use_nccl:
        h = capi.Handle(local_rank)
        tl = plan(W, H, rank, world, params)
        worker = TileWorker(h, params, tl, dev, stream, None)  # one marking stream per level
        ex = DistExchange(dist, rank, world)
        pin, rows = pins[0], infos[0]["sub_row1"] - infos[0]["sub_row0"]
        devt = {k: v.to(dev, non_blocking=True) for k, v in pin.items()}
        torch.cuda.synchronize()

        def prepare_resident():
            h.bind_device(devt["color"].data_ptr(), devt["nsamples"].data_ptr(), devt["hist"].data_ptr(), devt["cov"].data_ptr(), W, rows, Bd)

        def step_resident():
            run_tile(worker, ex)

        def step_e2e():
            h.upload_ptrs(pin["color"].data_ptr(), pin["nsamples"].data_ptr(), pin["hist"].data_ptr(), pin["cov"].data_ptr(), W, rows, Bd,
                          streamed=True)
            run_tile(worker, ex)
            worker.download(outs[0].data_ptr())

        def stats_timings():
            return h.stats(), h.timings()
    else:
        for i, (t, pin, out) in enumerate(zip(infos, pins, outs)):
            rows = t["sub_row1"] - t["sub_row0"]
            g.set_tile_buffers(i, pin["color"].data_ptr(), pin["nsamples"].data_ptr(), pin["hist"].data_ptr(), pin["cov"].data_ptr(),
                               W, rows, Bd, out.data_ptr())

        def prepare_resident():
            for i, (t, pin) in enumerate(zip(infos, pins)):
                rows = t["sub_row1"] - t["sub_row0"]
                g.upload_tile(i, pin["color"].data_ptr(), pin["nsamples"].data_ptr(), pin["hist"].data_ptr(), pin["cov"].data_ptr(), W, rows, Bd)

        def step_resident():
            g.run(params, from_host=False)

        def step_e2e():
            g.run(params, from_host=True)

        def stats_timings():
            return g.stats(), g.timings()

    def timed(fn, resident):
        if resident:
            prepare_resident()
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        sampler = B.ClockSampler(local_rank, B.gpu_uuid(local_rank))
        sampler.start()
        stage_acc, launches = {}, 0
        torch.cuda.synchronize()
        dist.barrier()
        # a frame spans the streams of several tile handles, so the K steps are bracketed by two events on an otherwise
        # idle stream, each recorded while the whole device is idle (synchronised just before): their distance is the
        # device-clock time from "before the first kernel" to "after the last one" (every step ends with all local tile
        # streams drained by bcd_b200_group_run)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            fn()  # nothing but the step itself inside the timed region
        torch.cuda.synchronize()
        e1.record(stream)
        torch.cuda.synchronize()
        ms_local = e0.elapsed_time(e1)
        dist.barrier()
        clocks = sampler.stop()
        ms = torch.tensor([ms_local], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # per-stage device times and launch counts from extra, untimed steps (every rank runs them: the tiles are coupled)
        probes = 3
        for _ in range(probes):
            fn()
            st, tm = stats_timings()
            for k, v in tm.items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v * a.steps / probes
            launches = st["kernel_launches"] * a.steps
        ln = torch.tensor([float(launches)], device=dev)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        return float(ms.item()), stage_acc, int(ln.item()), clocks

    ms_res, stages, launches, clocks = timed(step_resident, True)
    stats, _ = stats_timings()
    ms_e2e, _, _, _ = timed(step_e2e, False)
    # ---- parity of the N-GPU frame against the SAME frame denoised by one GPU (rank 0 gathers the bands)
    bands_local = [(t["own_row0"], torch.from_numpy(o.numpy().copy())) for t, o in zip(infos, outs)]
    gathered = [None] * world
    dist.all_gather_object(gathered, [(r0, b.numpy()) for r0, b in bands_local])
    parity = None
    if rank == 0:
        frame_n = np.empty((H, W, 3), np.float32)
        for part in gathered:
            for r0, band in part:
                frame_n[r0:r0 + band.shape[0]] = band
        full = host.synth(W, H, spp=a.spp)
        h1 = capi.Handle(local_rank)
        h1.upload(**full)
        h1.run(params)
        frame_1 = h1.download()
        st1 = h1.stats()
        h1.close()
        del full
        inner = (slice(2, -2), slice(2, -2))
        den = float(np.abs(frame_1[inner]).max())
        parity = {"rel_linf_vs_1gpu": float(np.abs(frame_n[inner] - frame_1[inner]).max() / den),
                  "what": "the N-GPU frame of this run against the same frame denoised by one GPU (bcd_b200_run); "
                          "only the order of the float atomics differs", "counters_1gpu": st1}
    st = torch.tensor([float(stats[k]) for k in ("interior", "strong", "active", "full", "fallback", "sum_n_full")], device=dev, dtype=torch.float64)
    dist.all_reduce(st, op=dist.ReduceOp.SUM)
    pair = torch.tensor([stages["pair"] / a.steps], device=dev)
    dist.all_reduce(pair, op=dist.ReduceOp.MAX)
    h2d = torch.tensor([float(sum(v.numel() * 4 for pin in pins for v in pin.values())), float(sum(o.numel() * 4 for o in outs))], device=dev)
    dist.all_reduce(h2d, op=dist.ReduceOp.SUM)
    per_rank = [None] * world
    dist.all_gather_object(per_rank, {k: round(v / a.steps, 4) for k, v in stages.items()})
    if rank == 0:
        value = W * H * a.steps / (ms_res * 1e-3) / 1e6
        e2e_val = W * H * a.steps / (ms_e2e * 1e-3) / 1e6
        pk, pk_src = B.peaks()
        npix_all = sum((W >> k) * (H >> k) for k in range(a.scales))
        alg_bytes = npix_all * (Bd + 1) * 4 / world          # per GPU
        achieved = alg_bytes / (float(pair.item()) * 1e-3) / 1e9
        counters = dict(zip(("interior", "strong", "active", "full", "fallback", "sum_n_full"), [int(v) for v in st.tolist()]))
        if parity is not None:
            parity["counters_equal"] = all(counters[k] == parity["counters_1gpu"][k] for k in counters)
        ntiles = g.tiles() if g is not None else world
        config = dict(config)
        config["parallelism"] = "%d row tiles dealt out cyclically to %d GPUs (one process per GPU)" % (ntiles, world)
        line = {"metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_res / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
                "timing": "CUDA events recorded on an idle device right before the first and right after the last of the K steps (a frame spans the streams of several tile handles), max over ranks",
                "e2e": {"value": e2e_val, "unit": B.UNIT, "h2d_bytes_per_step": int(h2d[0].item()), "d2h_bytes_per_step": int(h2d[1].item()),
                        "ms_per_step": ms_e2e / a.steps,
                        "api": "per tile: bcd_b200_upload_async(tile rows, pinned) + bcd_b200_tile_run + bcd_b200_tile_download (bcd_b200_group_run(from_host=1))"},
                "gpu_launches": launches, "tiles": ntiles,
                "exchange": "nccl send/recv per strip (A/B mode)" if use_nccl else
                            "peer memory: strips pushed into the neighbour's arena over NVLink (CUDA IPC), epoch flags, marking ring chained GPU to GPU; no NCCL call and no host hop inside a frame",
                "roofline": {"kernel": "k_pair_terms_sparse (per GPU, slowest rank; sum over its tiles)", "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"],
                             "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "peak_source": pk_src, "traffic": None,
                             "note": "FP32-issue-bound kernel; see the N=1 line for the FP32 figure"},
                "stage_ms_per_step_rank0": {k: v / a.steps for k, v in stages.items()},
                "stage_ms_per_step_by_rank": per_rank,
                "work_counters": counters, "parity": parity,
                "input_generation_s": gen_s}
        sys.stdout.flush()
        if saved_stdout_fd is not None:
            os.dup2(saved_stdout_fd, 1)
        print(json.dumps(line), flush=True)
    dist.barrier()
    if g is not None:
        g.close()
    else:
        h.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0

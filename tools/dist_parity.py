"""Multi-process check of the row-tile path (run under torchrun on the GPU box): every rank denoises its band; rank 0
also denoises the full frame on its own GPU and compares. Default: peer mode (CUDA IPC arenas over NVLink,
bcd_b200_tile_setup / _connect / _run — what bench.py's N > 1 arm uses); BCD_B200_EXCHANGE=nccl: the previous exchange
(torch.distributed send/recv per strip)."""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bcd_b200 import capi, host, tiles  # noqa: E402
import oracles as O  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    stream, aux = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
    torch.cuda.set_stream(stream)
    W, H = 640, 64 * world + 96
    ok = True
    for kw in (dict(m=1.0, scales=3), dict(m=0.0, scales=2)):
        params = capi.make_params(**kw)
        t = tiles.plan(W, H, rank, world, params)
        rows = t.sub_row1 - t.sub_row0
        sub = host.synth(W, H, spp=24, row0=t.sub_row0, rows=rows)
        h = capi.Handle(lr)
        if os.environ.get("BCD_B200_EXCHANGE", "peer") == "nccl":
            worker = tiles.TileWorker(h, params, t, torch.device("cuda", lr), stream, None)  # one marking stream per level
            h.upload(**sub)
            tiles.run_tile(worker, tiles.DistExchange(dist, rank, world))
        else:
            worker = tiles.TileWorker.__new__(tiles.TileWorker)
            worker.h, worker.tile = h, t
            tiles.peer_connect_dist(h, params, t, dist, rank, world)
            for _ in range(3):  # several frames: the epochs advance in step on every rank
                h.upload_async(**sub)
                tiles.peer_run(h, params)
        mine = torch.from_numpy(worker.download()).cuda()
        sizes = [tiles.plan(W, H, i, world, params) for i in range(world)]
        parts = [torch.empty((s.own_row1 - s.own_row0, W, 3), device="cuda") for s in sizes] if rank == 0 else None
        if rank == 0:
            parts[0].copy_(mine)
            for i in range(1, world):
                dist.recv(parts[i], src=i)
        else:
            dist.send(mine, dst=0)
        st = torch.tensor([float(h.stats()[k]) for k in ("interior", "strong", "active", "full", "fallback")], device="cuda", dtype=torch.float64)
        dist.all_reduce(st)
        if os.environ.get("BCD_B200_EXCHANGE", "peer") != "nccl":
            tiles.peer_connect(h, tiles.UP, None)
            tiles.peer_connect(h, tiles.DOWN, None)
            dist.barrier()
        h.close()
        if rank == 0:
            full = host.synth(W, H, spp=24)
            hf = capi.Handle(lr)
            hf.set_stream(stream.cuda_stream)
            hf.upload(**full)
            hf.run(params)
            ref = hf.download()
            sref = hf.stats()
            hf.close()
            out = torch.cat(parts, 0).cpu().numpy()
            err = O.rel_linf(out, ref)
            same = [int(v) for v in st.tolist()] == [sref[k] for k in ("interior", "strong", "active", "full", "fallback")]
            print("dist_parity world=%d %s: rel_linf=%.3e counters_equal=%s" % (world, kw, err, same), flush=True)
            ok = ok and err <= 2e-6 and same
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_PARITY", "OK" if ok else "FAILED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

"""SURVEY §8(d) measurement table on one GPU: C2 (1080p, s=1, defaults), C3 (1080p, s=3), the C5 sweep
b in {6,10,14} x w in {1,2,3} at 1080p s=1, each at m=1 and (optionally) m=0 — inputs resident in HBM, CUDA events on the
launching stream, one JSON line per configuration (Mpix/s, ms, stage times, work counters).

    python tools/sweep.py [--size 1920x1080] [--m0] [--steps 2] > profiles/rNN_sweep.jsonl
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bcd_b200 import capi, host  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--m0", action="store_true", help="also run every configuration with the marking strategy off")
    ap.add_argument("--max-ms", type=float, default=20000.0, help="skip the m=0 twin of a configuration slower than this")
    a = ap.parse_args()
    W, H = (int(v) for v in a.size.split("x"))
    B = 60
    pin = dict(color=torch.empty((H, W, 3)).pin_memory(), nsamples=torch.empty((H, W, 1)).pin_memory(),
               hist=torch.empty((H, W, B)).pin_memory(), cov=torch.empty((H, W, 6)).pin_memory())
    host.synth(W, H, spp=a.spp, out={k: v.numpy() for k, v in pin.items()})
    dev = {k: v.cuda() for k, v in pin.items()}
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    h = capi.Handle(0)
    h.set_stream(stream.cuda_stream)
    h.bind_device(dev["color"].data_ptr(), dev["nsamples"].data_ptr(), dev["hist"].data_ptr(), dev["cov"].data_ptr(), W, H, B)

    configs = [("C2 1080p s=1 defaults", dict(scales=1, b=6, w=1)), ("C3 1080p s=3", dict(scales=3, b=6, w=1))]
    for b in (6, 10, 14):
        for w in (1, 2, 3):
            if (b, w) != (6, 1):
                configs.append(("C5 s=1 b=%d w=%d" % (b, w), dict(scales=1, b=b, w=w)))

    def run(name, kw, m):
        params = capi.make_params(m=m, **kw)
        try:
            h.run(params)  # warm-up (also sizes the buffers)
        except capi.BcdError as e:
            print(json.dumps({"config": name, "m": m, "error": str(e)}), flush=True)
            return None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        stages = {}
        for _ in range(a.steps):
            h.run(params)
            for k, v in h.timings().items():
                stages[k] = stages.get(k, 0.0) + v / a.steps
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        print(json.dumps({"config": name, "size": a.size, "m": m, **kw, "ms_per_frame": ms, "mpix_per_s": W * H / ms / 1e3,
                          "stage_ms": {k: round(v, 3) for k, v in stages.items()}, "work_counters": h.stats()}), flush=True)
        return ms

    for name, kw in configs:
        ms = run(name, kw, 1.0)
        if a.m0 and ms is not None and kw["w"] == 1 and ms * 40 < a.max_ms:  # m=0 filters every strong pixel (only the d=27 rows)
            run(name, kw, 0.0)
    h.close()


if __name__ == "__main__":
    main()

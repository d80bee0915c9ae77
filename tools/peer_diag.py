"""Diagnostic of the peer-mode row tiles (bcd_b200_group_*): N tiles on the listed devices against the whole-frame
run, with the per-tile stage timeline (BCD_B200_TRACE=1) on stderr.  python tools/peer_diag.py [devices] [WxH] [scales] [m]"""
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")  # a lazy kernel load waits for an idle device (see bcd_b200_tile_setup)
os.environ.setdefault("BCD_B200_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from bcd_b200 import capi, host

devices = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,0").split(",")]
W, H = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "120x208").split("x")]
scales = int(sys.argv[3]) if len(sys.argv) > 3 else 1
m = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
inp = host.synth(W, H, spp=24)
h = capi.Handle(devices[0])
h.upload(**inp)
h.run(m=m, scales=scales)
ref, st_ref = h.download(), h.stats()
h.close()
print("whole frame:", st_ref, flush=True)
g = capi.Group(devices)
for it in range(3):
    t0 = time.time()
    try:
        out = g.denoise(**inp, m=m, scales=scales)
    except capi.BcdError as e:
        print("frame %d FAILED after %.2f s: %s" % (it, time.time() - t0, e), flush=True)
        continue
    st = g.stats()
    err = float(np.abs(out - ref)[2:-2, 2:-2].max() / np.abs(ref).max())
    print("frame %d: %.2f ms wall, tiles %d, rel Linf vs whole frame %.2e, counters equal %s, slowest tile %.3f ms" % (
        it, (time.time() - t0) * 1e3, g.tiles(), err, all(st[k] == st_ref[k] for k in st if k != "kernel_launches"), g.timings()["total"]), flush=True)
g.close()

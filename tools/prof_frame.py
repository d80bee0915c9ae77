"""Profiling target: the benchmark frame (G(1920,1080,spp), s=3, defaults) resident on the GPU, N frames through
bcd_b200_run — short enough to sit under `ncu` (see profiles/ for the captures made with it).
    python tools/prof_frame.py [frames] [WxH] [scales] [m] [spp]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcd_b200 import capi, host  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
W, H = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1920x1080").split("x")]
scales = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
spp = int(sys.argv[5]) if len(sys.argv) > 5 else 64
inp = host.synth(W, H, spp=spp)
h = capi.Handle(0)
h.upload(**inp)
p = capi.make_params(m=m, scales=scales)
for _ in range(frames):
    h.run(p)
print(h.stats(), h.timings())
h.close()

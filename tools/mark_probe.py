#!/usr/bin/env python
"""Marking-stage timing probe: frames of 1920 columns and 1..N row bands, single scale, m = 1 (K2 alone is what varies)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bcd_b200 import capi, host
W = int(os.environ.get("PROBE_W", "1920"))
for H in [int(v) for v in os.environ.get("PROBE_H", "34,66,130,258,514,1080").split(",")]:
    inp = host.synth(W, H, spp=16)
    h = capi.Handle(0)
    h.upload(**inp)
    best = None
    for _ in range(6):
        h.run(m=1.0, scales=1)
        t = h.timings()
        best = t["marking"] if best is None else min(best, t["marking"])
    print(json.dumps({"W": W, "H": H, "bands": (H - 2 + 31) // 32, "marking_ms": round(best, 4), "kernel": os.environ.get("BCD_B200_MARKING", "sys"),
                      "nw": os.environ.get("BCD_B200_MARKING_NW", "4")}))
    h.close()

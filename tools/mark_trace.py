#!/usr/bin/env python
"""One traced marking launch (BCD_B200_MARK_TRACE=1): per-band timeline of the systolic kernel on a W x H frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BCD_B200_MARK_TRACE", "0")
from bcd_b200 import capi, host
W, H = int(sys.argv[1]), int(sys.argv[2])
inp = host.synth(W, H, spp=16)
os.environ["BCD_B200_MARK_TRACE"] = "0"
h = capi.Handle(0)
h.upload(**inp)
for _ in range(2):
    h.run(m=1.0, scales=1)
h.close()
os.environ["BCD_B200_MARK_TRACE"] = "1"
h = capi.Handle(0)  # re-reads the hook
h.upload(**inp)
h.run(m=1.0, scales=1)
print("marking stage ms:", h.timings()["marking"])
h.close()

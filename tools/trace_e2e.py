"""Stage timeline of one resident and one streamed-upload frame (BCD_B200_TRACE=1): which stages overlap, where the
main stream waits. Usage: BCD_B200_TRACE=1 python tools/trace_e2e.py [W H]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bcd_b200 import capi, host  # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
B = 60
pin = dict(color=torch.empty((H, W, 3)).pin_memory(), nsamples=torch.empty((H, W, 1)).pin_memory(),
           hist=torch.empty((H, W, B)).pin_memory(), cov=torch.empty((H, W, 6)).pin_memory())
host.synth(W, H, spp=32, out={k: v.numpy() for k, v in pin.items()})
out = torch.empty((H, W, 3)).pin_memory()
h = capi.Handle(0)
params = capi.make_params(m=1.0, scales=3)
ptrs = [pin[k].data_ptr() for k in ("color", "nsamples", "hist", "cov")]
os.environ.pop("BCD_B200_TRACE", None)
for _ in range(2):
    h.upload_ptrs(*ptrs, W, H, B, streamed=True)
    h.run(params)
    h.download_ptr(out.data_ptr())
os.environ["BCD_B200_TRACE"] = "1"
print("==== streamed upload + run", file=sys.stderr)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
h.upload_ptrs(*ptrs, W, H, B, streamed=True)
h.run(params)
h.download_ptr(out.data_ptr())
e1.record()
torch.cuda.synchronize()
print("e2e frame: %.3f ms" % e0.elapsed_time(e1), file=sys.stderr)
print("==== resident run", file=sys.stderr)
h.upload_ptrs(*ptrs, W, H, B)
h.run(params)
h.close()

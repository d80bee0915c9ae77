"""First-contact diagnostic for the GPU box: runs every stage on the golden crops and prints detailed comparisons
(so one gpurun call tells as much as possible). Not a test; see tests/test_gpu_parity.py."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracles as O  # noqa: E402
from bcd_b200 import capi  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def inputs_of(d):
    return {k: d[k] for k in ("color", "nsamples", "hist", "cov")}


def run(inp, **kw):
    h = capi.Handle(0)
    h.upload(**inp)
    t = time.time()
    h.run(**kw)
    dt = time.time() - t
    out = h.download()
    return h, out, dt


def main():
    print("devices:", capi.lib().bcd_b200_device_count())
    a = dict(np.load(os.path.join(G, "testraw_crop_a.npz")))
    inp = inputs_of(a)
    # ---- selection
    h, out, dt = run(inp, m=0.0)
    masks, counts = h.read_selection()
    print("m=0 run %.1f ms stats %s" % (dt * 1e3, h.stats()))
    print("timings", {k: round(v, 3) for k, v in h.timings().items()})
    diff = masks ^ a["port_masks"]
    nb = int(np.unpackbits(diff.view(np.uint8)).sum())
    print("selection: mismatching bits =", nb, " count mismatches =", int((counts != a["port_counts"]).sum()))
    if nb:
        bad = np.argwhere(diff.any(axis=2))
        print("  first bad pixels", bad[:10].tolist())
        y, x = bad[0]
        print("  gpu ", [hex(v) for v in masks[y, x]], counts[y, x])
        print("  port", [hex(v) for v in a["port_masks"][y, x]], a["port_counts"][y, x])
    print("e2e s1 m0 rel_linf", O.rel_linf(out, a["ref_s1_m0"]), "nan gpu/ref", int(np.isnan(out).sum()), int(np.isnan(a["ref_s1_m0"]).sum()))
    err = np.abs(out - a["ref_s1_m0"]).max(axis=2)
    yx = np.unravel_index(np.nanargmax(err), err.shape)
    print("  worst pixel", yx, out[yx], a["ref_s1_m0"][yx])
    h.close()
    # ---- marking
    h, out, dt = run(inp, m=1.0)
    act = h.read_active()
    print("m=1 run %.1f ms stats %s" % (dt * 1e3, h.stats()))
    print("timings", {k: round(v, 3) for k, v in h.timings().items()})
    bad = np.argwhere(act != a["port_active_m1"])
    print("marking mismatches:", bad.shape[0], bad[:10].tolist())
    print("e2e s1 m1 rel_linf", O.rel_linf(out, a["ref_s1_m1"]))
    h.close()
    for key, kw in [("ref_s2_m0", dict(m=0.0, scales=2)), ("ref_s3_m1", dict(m=1.0, scales=3))]:
        h, out, dt = run(inp, **kw)
        print(key, "rel_linf", O.rel_linf(out, a[key]), "%.1f ms" % (dt * 1e3), h.stats())
        h.close()
    b = dict(np.load(os.path.join(G, "testraw_crop_b.npz")))
    for key, kw in [("ref_s2_m0", dict(m=0.0, scales=2)), ("ref_s2_m1", dict(m=1.0, scales=2)),
                    ("ref_s1_m0_b4", dict(m=0.0, b=4)), ("ref_s1_m1_tau2", dict(m=1.0, tau=2.0))]:
        h, out, dt = run(inputs_of(b), **kw)
        print("B", key, "rel_linf", O.rel_linf(out, b[key]), "%.1f ms" % (dt * 1e3))
        h.close()
    c = dict(np.load(os.path.join(G, "testraw_crop_c.npz")))
    h, out, dt = run(inputs_of(c), m=0.0, w=2, b=5)
    print("C w2 rel_linf", O.rel_linf(out, c["ref_s1_m0_w2b3"]), h.stats())
    h.close()
    if O.have_ref():
        # live reference on the box (prebuilt oracle/_ref travels with the snapshot)
        ref = O.ref_denoise(**inputs_of(b), cores=0, m=0.0)
        print("live ref vs golden(b4?) sanity:", O.rel_linf(ref, ref))


if __name__ == "__main__":
    main()

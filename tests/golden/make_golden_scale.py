"""Golden fixtures at benchmark scale and for the wide patches — run HERE (build container), where
oracle/_ref/libbcd_ref.so (the unmodified reference compiled in place) exists; the GPU box only reads the .npz files.

Inputs are NOT stored: they come from the seeded generator G(W,H,spp,seed) (bcd_b200.host.synth, counter based, so any
machine regenerates the same floats — the tests assert a checksum of the regenerated inputs before comparing).
Stored: the reference's float outputs in its deterministic settings (raster order; m=1 on ONE thread, or m=0).

    python tests/golden/make_golden_scale.py            # ~5 minutes on 8 cores
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracles as O  # noqa: E402
from bcd_b200 import host  # noqa: E402


def checksum(inp):
    """Order-independent fingerprint of the regenerated inputs (float64 sums per image)."""
    return np.array([float(np.asarray(inp[k], np.float64).sum()) for k in ("color", "nsamples", "hist", "cov")])


def window(inp, y0, x0, h, w):
    return {k: np.ascontiguousarray(v[y0:y0 + h, x0:x0 + w]) for k, v in inp.items()}


def centre(inp, c):
    H, W, _ = inp["color"].shape
    return window(inp, (H - c) // 2, (W - c) // 2, c, c)


def main():
    t0 = time.time()
    # C1 (BASELINE.json configs[0]): G(256,256) single scale, b=6 w=1 — both skipping modes
    c1 = host.synth(256, 256, spp=64)
    np.savez_compressed(os.path.join(HERE, "scale_c1_256.npz"), checksum=checksum(c1),
                        ref_s1_m1=O.ref_denoise(**c1, cores=1, m=1.0, scales=1),
                        ref_s1_m0=O.ref_denoise(**c1, cores=8, m=0.0, scales=1))
    print("C1 done %.0f s" % (time.time() - t0), flush=True)
    # the benchmark frame itself: 512 x 512 centre window of G(1920,1080), s=3 m=1 b=6 w=1, one thread
    frame = host.synth(1920, 1080, spp=64)
    win = centre(frame, 512)
    np.savez_compressed(os.path.join(HERE, "scale_bench_512.npz"), checksum=checksum(win),
                        ref_s3_m1=O.ref_denoise(**win, cores=1, m=1.0, scales=3))
    print("bench window done %.0f s" % (time.time() - t0), flush=True)
    # wide patches on the smooth (left) half of C1, where the similar sets are large enough for d = 75 / 147
    w2 = window(c1, 96, 24, 64, 64)
    r2 = O.ref_denoise(**w2, cores=8, m=0.0, scales=1, w=2, b=6)
    w3 = window(c1, 100, 30, 44, 44)
    r3 = O.ref_denoise(**w3, cores=8, m=0.0, scales=1, w=3, b=7)
    r2m1 = O.ref_denoise(**w2, cores=1, m=1.0, scales=2, w=2, b=6)
    np.savez_compressed(os.path.join(HERE, "scale_wide_patches.npz"), checksum_w2=checksum(w2), checksum_w3=checksum(w3),
                        ref_w2_b6_m0=r2, ref_w3_b7_m0=r3, ref_w2_b6_s2_m1=r2m1)
    print("wide patches done %.0f s" % (time.time() - t0), flush=True)
    for f in sorted(os.listdir(HERE)):
        if f.startswith("scale_"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()

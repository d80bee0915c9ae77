"""Generates the committed golden fixtures under tests/golden/ — run HERE (build container), where /root/reference
exists; the GPU box only reads the .npz files.

Inputs : crops of the reference's only fixture data/raw/test.raw (320x240, 81 spp) pushed through the reference's own
         SamplesAccumulator (oracle/_ref).
Outputs: the UNMODIFIED reference CPU path (oracle/_ref/libbcd_ref.so) in its deterministic settings
         (useRandomPixelOrder=false; m=1 with nbOfCores=1, or m=0), plus the selection masks / counts / marking
         outcome of the C restatement (oracle/liboracle.so), whose work counters equal the reference's own statistics.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracles as O  # noqa: E402

RAW = "/root/reference/data/raw/test.raw"


def main():
    samples = O.load_test_raw(RAW)
    full = O.ref_accumulate(samples)

    def crop(y0, y1, x0, x1):
        return {k: np.ascontiguousarray(v[y0:y1, x0:x1]) for k, v in full.items()}

    # A: 96 x 128 crop mixing strong and weak regions (even sizes: 96x128 -> 48x64 -> 24x32)
    a = crop(48, 144, 64, 192)
    out = dict(a)
    out["ref_s1_m0"] = O.ref_denoise(**a, cores=8, m=0.0, scales=1)
    out["ref_s1_m1"] = O.ref_denoise(**a, cores=1, m=1.0, scales=1)
    out["ref_s3_m1"] = O.ref_denoise(**a, cores=1, m=1.0, scales=3)
    out["ref_s2_m0"] = O.ref_denoise(**a, cores=8, m=0.0, scales=2)
    masks, counts = O.port_select(a["hist"], a["nsamples"])
    out["port_masks"] = masks
    out["port_counts"] = counts
    out["port_active_m1"] = O.port_marking(masks, counts)
    np.savez_compressed(os.path.join(HERE, "testraw_crop_a.npz"), **out)

    # B: odd-sized 51 x 67 crop (integer halving + clamped interpolation in the pyramid), other radii
    b = crop(100, 151, 30, 97)
    out = dict(b)
    out["ref_s2_m0"] = O.ref_denoise(**b, cores=8, m=0.0, scales=2)
    out["ref_s2_m1"] = O.ref_denoise(**b, cores=1, m=1.0, scales=2)
    out["ref_s1_m0_b4"] = O.ref_denoise(**b, cores=8, m=0.0, scales=1, b=4)
    out["ref_s1_m1_tau2"] = O.ref_denoise(**b, cores=1, m=1.0, scales=1, tau=2.0)
    np.savez_compressed(os.path.join(HERE, "testraw_crop_b.npz"), **out)

    # C: tiny 24 x 28 crop with w=2 (d=75) — exercises the wide-patch kernels
    c = crop(60, 84, 150, 178)
    out = dict(c)
    out["ref_s1_m0_w2b3"] = O.ref_denoise(**c, cores=8, m=0.0, scales=1, w=2, b=5)
    np.savez_compressed(os.path.join(HERE, "testraw_crop_c.npz"), **out)
    # D (row f1): SpikeRemovalFilter of the compiled reference on the seeded synthetic frame G(96,64,spp=32) — the input is
    # regenerated at test time by bcd_b200.host.synth (deterministic); only the reference's outputs are stored
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from bcd_b200 import host
    g = host.synth(96, 64, spp=32)
    r = O.ref_spike_removal(**g, factor=2.0)
    changed = (r["color"] != g["color"]).any(-1)
    np.savez_compressed(os.path.join(HERE, "spike_synth_96x64.npz"), color=r["color"], nsamples=r["nsamples"], cov=r["cov"],
                        hist_channel_sums=r["hist"].astype(np.float64).sum(axis=(0, 1)), changed=changed,
                        ref_denoised_after_spike=O.ref_denoise(**r, cores=1, m=1.0, scales=2))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()

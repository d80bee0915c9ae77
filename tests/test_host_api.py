"""CPU: host-side pieces — the re-implemented bcd::SamplesAccumulator (bit-exact against the reference's), the
synthetic generator's workload (SURVEY §8d counters) and the C++ drop-in's validation path."""
import numpy as np
import pytest

import oracles as O
from bcd_b200 import host


@pytest.mark.skipif(not O.have_ref(), reason="compiled reference (oracle/_ref) not built here")
def test_samples_accumulator_bit_exact_vs_reference():
    rng = np.random.default_rng(7)
    s = (rng.lognormal(-1.0, 0.9, size=(9, 11, 17, 3))).astype(np.float32)
    s[0, 0, 0] = (-1.0, 0.0, 50.0)  # negative, zero, saturated
    a, b = O.ref_accumulate(s), host.accumulate(s)
    for k in a:
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    a, b = O.ref_accumulate(s, nb_bins=12, gamma=1.0, max_value=4.0), host.accumulate(s, nb_bins=12, gamma=1.0, max_value=4.0)
    for k in a:
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k


def test_synthetic_generator_is_reproducible_and_tileable():
    a = host.synth(40, 24, spp=16)
    b = host.synth(40, 24, spp=16)
    top = host.synth(40, 24, spp=16, row0=0, rows=10)
    bot = host.synth(40, 24, spp=16, row0=10, rows=14)
    for k in a:
        assert np.array_equal(a[k], b[k])
        assert np.array_equal(a[k], np.concatenate([top[k], bot[k]], axis=0))
    assert np.allclose(a["nsamples"], 16) and np.allclose(a["hist"].reshape(24, 40, 3, 20).sum(-1), 16, atol=1e-3)


def test_synthetic_workload_counters_match_survey():
    # SURVEY §8d measured with the reference semantics at 256^2: 29,951 strong of 64,516 interior;
    # m=1 raster: 850 full + 34,236 fallback. A 128x96 window must show the same ~46 % mix.
    g = host.synth(128, 96)
    masks, counts = O.port_select(g["hist"], g["nsamples"])
    interior = np.zeros_like(counts, bool)
    interior[1:-1, 1:-1] = True
    frac = ((counts >= 28) & interior).sum() / interior.sum()
    assert 0.3 < frac < 0.7


def _build_dropin_check(tmp_path):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "dropin_check")
    subprocess.run(["g++", "-std=c++11", "-O1", "-I" + os.path.join(root, "include", "bcd", "core"), "-I" + os.path.join(root, "include"),
                    os.path.join(root, "tests", "cxx", "dropin_check.cpp"), "-o", exe, "-L" + os.path.join(root, "bcd_b200"),
                    "-lbcdcore_b200", "-lbcd_b200", "-Wl,-rpath," + os.path.join(root, "bcd_b200")], check=True)
    return exe


def test_cxx_dropin_compiles_like_a_host_renderer_and_validates(tmp_path):
    """A C++ program written against the reference's headers (bare names, bcd:: classes) builds against ours and gets
    the reference's validation behaviour; without a GPU denoise() fails loudly instead of falling back."""
    import subprocess
    from bcd_b200 import capi
    exe = _build_dropin_check(tmp_path)
    gpu = capi.lib().bcd_b200_device_count() > 0
    r = subprocess.run([exe] + (["gpu"] if gpu else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Aborting denoising: nullptr for input color image" in r.stderr
    assert "input number of samples image is 41x32but input color image is 40x32" in r.stderr
    assert "Aborting denoising: input histogram image is empty" in r.stderr
    if not gpu:
        assert "no CPU path" in r.stderr


@pytest.mark.gpu
def test_cxx_dropin_denoises_on_gpu(tmp_path):
    import subprocess
    exe = _build_dropin_check(tmp_path)
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cxx_dropin_matches_c_abi(crop_b):
    from conftest import inputs_of
    from bcd_b200 import capi
    inp = inputs_of(crop_b)
    ok, out, calls = host.cxx_denoise(**inp, m=1.0, scales=2)
    assert ok and calls >= 2
    assert O.rel_linf(out, capi.denoise(**inp, m=1.0, scales=2)) <= 1e-5
    assert O.rel_linf(out, crop_b["ref_s2_m1"]) <= 1e-4

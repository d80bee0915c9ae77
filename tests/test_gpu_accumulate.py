"""GPU: SamplesAccumulator on the device (bcd_b200_accumulate_samples) against the host accumulator,
which is bit-exact with the reference's (tests/test_host_api.py::test_samples_accumulator_bit_exact_vs_reference)."""
import numpy as np
import pytest

import oracles as O

pytestmark = pytest.mark.gpu


def _samples(H, W, spp, seed):
    rng = np.random.default_rng(seed)
    base = rng.uniform(0.02, 1.2, (H, W, 1, 3)).astype(np.float32)
    s = base * np.exp(rng.normal(0.0, 0.6, (H, W, spp, 3))).astype(np.float32)
    s[rng.uniform(size=(H, W, spp)) < 2e-3] *= 20.0      # fireflies (saturated bins)
    s[0, 0, 0] = (-0.5, 0.0, 3.0e3)                       # negative / zero / far beyond the last bin
    return s.astype(np.float32)


def test_device_accumulation_matches_the_host_accumulator():
    from bcd_b200 import capi, host
    H, W, spp = 40, 56, 24
    samples = _samples(H, W, spp, 5)
    ref = host.accumulate(samples)
    h = capi.Handle(0)
    try:
        h.accumulate_samples(samples)
        got = h.download_inputs()
        # everything that does not go through pow() is the reference's number bit for bit
        assert np.array_equal(got["color"], ref["color"])
        assert np.array_equal(got["nsamples"], ref["nsamples"])
        assert np.array_equal(got["cov"], ref["cov"])
        # histogram: device powf vs glibc powf moves at most ~1e-6 of a sample between two neighbouring bins
        assert np.abs(got["hist"] - ref["hist"]).max() <= 4e-6
        assert np.allclose(got["hist"].reshape(H, W, 3, 20).sum(-1), spp, atol=1e-4)
        # the statistics stay resident: denoising them equals denoising the host-accumulated frame
        h.run(m=1.0, scales=2)
        out = h.download()
        h.upload(**ref)
        h.run(m=1.0, scales=2)
        assert O.rel_linf(out, h.download()) <= 1e-5
    finally:
        h.close()


def _cudart():
    import ctypes
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            return ctypes.CDLL(name)
        except OSError:
            continue
    pytest.skip("no CUDA runtime library to allocate a device block with")


def test_device_accumulation_from_device_memory_with_weights():
    """The samples (and per-sample weights) already live in device memory, like a GPU renderer's would."""
    import ctypes
    from bcd_b200 import capi
    rt = _cudart()
    H, W, spp = 24, 32, 8
    samples = _samples(H, W, spp, 9)
    weights = np.random.default_rng(3).uniform(0.25, 2.0, (H, W, spp)).astype(np.float32)
    h = capi.Handle(0)  # creates the context
    ds, dw = ctypes.c_void_p(), ctypes.c_void_p()
    assert rt.cudaMalloc(ctypes.byref(ds), ctypes.c_size_t(samples.nbytes)) == 0
    assert rt.cudaMalloc(ctypes.byref(dw), ctypes.c_size_t(weights.nbytes)) == 0
    try:
        assert rt.cudaMemcpy(ds, ctypes.c_void_p(samples.ctypes.data), ctypes.c_size_t(samples.nbytes), 1) == 0
        assert rt.cudaMemcpy(dw, ctypes.c_void_p(weights.ctypes.data), ctypes.c_size_t(weights.nbytes), 1) == 0
        h.accumulate_samples(None, device_ptrs=(ds.value, dw.value, H, W, spp))
        got = h.download_inputs()
    finally:
        h.close()
        rt.cudaFree(ds)
        rt.cudaFree(dw)
    # weighted mean / count against a float64 evaluation (the reference's accumulator with weights, in float32 order)
    wsum = weights.astype(np.float64).sum(-1)
    mean = (weights[..., None].astype(np.float64) * samples).sum(2) / wsum[..., None]
    assert np.allclose(got["nsamples"][..., 0], wsum, rtol=1e-6)
    assert np.allclose(got["color"], mean, rtol=2e-5, atol=1e-6)
    assert np.allclose(got["hist"].reshape(H, W, 3, 20).sum(-1), wsum[..., None], rtol=1e-5)

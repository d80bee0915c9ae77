"""CPU: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bcd_(?:b200|synth|host)_[a-z0-9_]+)\s*\(", src)))


def test_bcd_b200_h_symbols_exported():
    from bcd_b200 import capi
    names = _declared("bcd_b200.h")
    assert len(names) >= 15
    L = capi.lib()
    for n in names:
        assert hasattr(L, n), "symbol %s declared in include/bcd_b200.h is not exported" % n
    assert sorted(capi.SYMBOLS) == names
    assert L.bcd_b200_abi_version() == 2


def test_host_header_symbols_exported():
    from bcd_b200 import host
    names = _declared("bcd_b200_host.h")
    L = host.lib()
    for n in names:
        assert hasattr(L, n), n
    assert sorted(host.SYMBOLS) == names


def test_cxx_api_symbols_exported():
    """The bcd:: drop-in classes are real exported C++ symbols (a host renderer links against them)."""
    import subprocess
    out = subprocess.run(["nm", "-DC", os.path.join(ROOT, "bcd_b200", "libbcdcore_b200.so")], capture_output=True, text=True).stdout
    for sym in ("bcd::Denoiser::denoise()", "bcd::MultiscaleDenoiser::denoise()", "bcd::Denoiser::inputsOutputsAreOk()",
                "bcd::SamplesAccumulator::addSample(int, int, float, float, float, float)",
                "bcd::SamplesAccumulator::extractSamplesStatistics()"):
        assert sym in out, sym


def test_fails_loudly_without_device_or_bad_args():
    from bcd_b200 import capi
    L = capi.lib()
    p = capi.make_params()
    assert p.search_radius == 6 and p.patch_radius == 1 and p.nb_scales == 1
    d = capi.Params()
    L.bcd_b200_default_params(ctypes.byref(d))
    assert (d.tau, d.patch_radius, d.search_radius, d.skip_prob, d.nb_scales, d.random_order) == (1.0, 1, 6, 1.0, 1, 1)
    assert abs(d.min_eig - 1e-8) < 1e-12
    assert L.bcd_b200_mask_words(6) == 6
    if L.bcd_b200_device_count() == 0:
        import pytest
        with pytest.raises(capi.BcdError) as e:
            capi.Handle(0)
        assert "no CPU path" in str(e.value)
    # null inputs are rejected with the reference's message, before any device work
    rc = L.bcd_b200_denoise(None, ctypes.byref(p), None, 0, None, None)
    assert rc == -1
    img = capi.Images(None, None, None, None, 4, 4, 60)
    import numpy as np
    out = np.zeros(48, np.float32)
    rc = L.bcd_b200_denoise(ctypes.byref(img), ctypes.byref(p), out.ctypes.data, 0, None, None)
    assert rc == -1 and b"nullptr for input color image" in L.bcd_b200_last_error()

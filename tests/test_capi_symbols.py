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


def test_group_api_fails_loudly_without_device_and_plans_on_the_host():
    """The multi-GPU entry points: no device -> BCD_B200_ERR_NO_DEVICE (there is no CPU path), bad arguments are refused
    before any device work, and the tile plan (pure host arithmetic) is consistent with what a group would cut."""
    import pytest
    from bcd_b200 import capi, tiles
    L = capi.lib()
    if L.bcd_b200_device_count() == 0:
        with pytest.raises(capi.BcdError) as e:
            capi.Group([0, 1])
        assert e.value.code == -3 and "no CPU path" in str(e.value)
        import numpy as np
        z = np.zeros((8, 8, 3), np.float32)
        with pytest.raises(capi.BcdError) as e:
            capi.denoise_multi(z, np.zeros((8, 8, 1), np.float32), np.zeros((8, 8, 60), np.float32), np.zeros((8, 8, 6), np.float32),
                               devices=[0, 0])
        assert e.value.code == -3
    p = capi.make_params(scales=3)
    # a null group / a ranked group with a bad rank never touch a device
    assert L.bcd_b200_group_denoise(None, None, ctypes.byref(p), None, None, None) == -1
    g = ctypes.c_void_p()
    ids = (ctypes.c_int32 * 1)(0)
    assert L.bcd_b200_group_create_ranked(ctypes.byref(g), ids, 1, 3, 2) in (-1, -3)
    # 8 tiles of a 1080p frame at 3 scales: bands on multiples of 4 rows, 28 halo rows, cyclic owners are the caller's business
    ts = [tiles.plan(1920, 1080, i, 8, p) for i in range(8)]
    assert [t.own_row1 - t.own_row0 for t in ts] == [132, 136, 136, 136, 132, 136, 136, 136]
    assert all(t.sub_row0 == max(0, t.own_row0 - 28) and t.sub_row1 == min(1080, t.own_row1 + 28) for t in ts)

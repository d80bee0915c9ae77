"""GPU parity tests proper: the CUDA path, called through the C ABI (bcd_b200/capi.py -> libbcd_b200.so), against
  (1) the committed golden fixtures produced by the unmodified reference CPU path (tests/golden/make_golden.py),
  (2) the C restatement's selection masks / counts / marking outcome, bit for bit.
Tolerance: relative L-infinity (max|a-ref| / max|ref|) <= 1e-4 for the float output (north_star); 0 mismatching bits
for the discrete stages.
"""
import numpy as np
import pytest

from conftest import inputs_of
import oracles as O

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def capi():
    from bcd_b200 import capi
    if capi.lib().bcd_b200_device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on the B200 box")
    return capi


def _run(capi, inp, **kw):
    h = capi.Handle(0)
    try:
        h.upload(**inp)
        h.run(**kw)
        out = h.download()
        return h, out
    except Exception:
        h.close()
        raise


def test_selection_bit_exact(capi, crop_a):
    h, _ = _run(capi, inputs_of(crop_a), m=0.0)
    masks, counts = h.read_selection()
    st = h.stats()
    h.close()
    ref_masks, ref_counts = crop_a["port_masks"], crop_a["port_counts"]
    diff_bits = int(np.unpackbits((masks ^ ref_masks).view(np.uint8)).sum())
    bad = np.argwhere(counts != ref_counts)
    assert diff_bits == 0, "%d mismatching selection bits, first at %s" % (diff_bits, bad[:5].tolist())
    assert (counts == ref_counts).all()
    assert st["strong"] == int((ref_counts >= 28).sum())


@pytest.mark.parametrize("dense", [False, True])
def test_pair_terms_bit_exact(capi, crop_a, crop_b, dense, monkeypatch):
    """K1a alone: every pair term (sum bits and count) of every pixel and offset equals the restated reference term
    (DenoisingUnit.cpp:360-386) — the selection bits downstream only see these through a threshold, which would hide a
    last-bit difference such as a contracted multiply-add in the packed-arithmetic path."""
    monkeypatch.setenv("BCD_B200_DENSE_PAIR", "1" if dense else "0")
    try:
        for crop, bsr in ((crop_a, 6), (crop_b, 4), (crop_b, 14)):
            inp = inputs_of(crop)
            h, _ = _run(capi, inp, m=0.0, scales=1, b=bsr)
            sums, counts = h.read_pair_terms(b=bsr)
            h.close()
            ref_s, ref_c = O.port_pair_terms(inp["hist"], inp["nsamples"], b=bsr)
            valid = ref_c != 255
            assert valid.sum() > 0.6 * valid.size
            assert np.array_equal(counts[valid], ref_c[valid])
            assert np.array_equal(sums[valid].view(np.uint32), ref_s[valid].view(np.uint32))
    finally:
        monkeypatch.setenv("BCD_B200_DENSE_PAIR", "0")
        capi.Handle(0).close()  # reset the process-wide hook


@pytest.mark.parametrize("kern", ["v2", "v1"])
def test_pair_terms_uniform_and_ragged_sample_counts(capi, kern, monkeypatch):
    """K1a's uniform-2^k-sample-count shortcut ((b1-b2)^2/(b1+b2), k_distance.cu) against the restated reference term,
    bit for bit: frames where it applies everywhere (n = 64; n = 1/8), nowhere (n = 48; n = 2^20 fails the range flag),
    and tile by tile (left half 64, right half 48, scattered odd pixels) — plus the previous kernel as the A/B twin."""
    from bcd_b200 import host
    monkeypatch.setenv("BCD_B200_PAIR", kern)
    try:
        base = host.synth(150, 41, spp=64)
        H, W = base["nsamples"].shape[:2]
        rng = np.random.default_rng(5)
        cases = {}
        for name, n in (("n64", 64.0), ("n_eighth", 0.125), ("n48", 48.0), ("n2^20", 1048576.0)):
            cases[name] = np.full((H, W, 1), n, np.float32)
        mixed = np.full((H, W, 1), 64.0, np.float32)
        mixed[:, W // 2:] = 48.0
        ys, xs = rng.integers(0, H, 12), rng.integers(0, W // 2, 12)
        mixed[ys, xs, 0] = rng.integers(1, 200, 12).astype(np.float32)
        cases["mixed"] = mixed
        for name, ns in cases.items():
            inp = dict(base)
            inp["nsamples"] = np.ascontiguousarray(ns.reshape(base["nsamples"].shape))
            h, _ = _run(capi, inp, m=0.0, scales=1)
            sums, counts = h.read_pair_terms(b=6)
            h.close()
            ref_s, ref_c = O.port_pair_terms(inp["hist"], inp["nsamples"], b=6)
            valid = ref_c != 255
            assert np.array_equal(counts[valid], ref_c[valid]), name
            bad = np.argwhere((sums.view(np.uint32) != ref_s.view(np.uint32)) & valid)
            assert bad.shape[0] == 0, "%s: %d pair sums differ, first (j,y,x) %s" % (name, bad.shape[0], bad[:4].tolist())
    finally:
        monkeypatch.setenv("BCD_B200_PAIR", "v2")
        capi.Handle(0).close()  # reset the process-wide hook


def test_marking_bit_exact(capi, crop_a):
    h, _ = _run(capi, inputs_of(crop_a), m=1.0)
    active = h.read_active()
    st = h.stats()
    h.close()
    ref = crop_a["port_active_m1"]
    bad = np.argwhere(active != ref)
    assert bad.shape[0] == 0, "%d marking mismatches, first (row,col) %s" % (bad.shape[0], bad[:8].tolist())
    strong = crop_a["port_counts"] >= 28
    assert st["full"] == int((ref.astype(bool) & strong).sum())
    assert st["fallback"] == int((ref.astype(bool) & ~strong).sum())


@pytest.mark.parametrize("key,kw", [
    ("ref_s1_m0", dict(m=0.0, scales=1)),
    ("ref_s1_m1", dict(m=1.0, scales=1)),
    ("ref_s2_m0", dict(m=0.0, scales=2)),
    ("ref_s3_m1", dict(m=1.0, scales=3)),
])
def test_end_to_end_crop_a(capi, crop_a, key, kw):
    h, out = _run(capi, inputs_of(crop_a), **kw)
    h.close()
    err = O.rel_linf(out, crop_a[key])
    assert err <= TOL, "%s: rel Linf %.3e" % (key, err)


@pytest.mark.parametrize("key,kw", [
    ("ref_s2_m0", dict(m=0.0, scales=2)),
    ("ref_s2_m1", dict(m=1.0, scales=2)),
    ("ref_s1_m0_b4", dict(m=0.0, scales=1, b=4)),
    ("ref_s1_m1_tau2", dict(m=1.0, scales=1, tau=2.0)),
])
def test_end_to_end_crop_b_odd_size(capi, crop_b, key, kw):
    h, out = _run(capi, inputs_of(crop_b), **kw)
    h.close()
    err = O.rel_linf(out, crop_b[key])
    assert err <= TOL, "%s: rel Linf %.3e" % (key, err)


def test_wide_patch_w2(capi, crop_c):
    h, out = _run(capi, inputs_of(crop_c), m=0.0, w=2, b=5)
    st = h.stats()
    h.close()
    err = O.rel_linf(out, crop_c["ref_s1_m0_w2b3"])
    assert err <= TOL, "w=2: rel Linf %.3e (stats %s)" % (err, st)


def test_one_shot_entry_matches_handle(capi, crop_b):
    inp = inputs_of(crop_b)
    out1 = capi.denoise(**inp, m=0.0)
    h, out2 = _run(capi, inp, m=0.0)
    h.close()
    assert O.rel_linf(out1, out2) <= 1e-5


def test_invalid_inputs_are_rejected(capi, crop_c):
    inp = inputs_of(crop_c)
    with pytest.raises(capi.BcdError):
        capi.denoise(**inp, b=0)
    with pytest.raises(capi.BcdError):
        capi.denoise(**inp, scales=0)


@pytest.mark.parametrize("nb_bins", [7, 24, 22])  # B = 21 (sparse kernel, odd depth), 72 (dense tiled), 66 (generic)
def test_other_histogram_depths(capi, nb_bins):
    from bcd_b200 import host
    inp = host.synth(56, 44, spp=24, nb_bins=nb_bins)
    h, out = _run(capi, inp, m=1.0)
    masks, counts = h.read_selection()
    active = h.read_active()
    h.close()
    ref_masks, ref_counts = O.port_select(inp["hist"], inp["nsamples"])
    assert int(np.unpackbits((masks ^ ref_masks).view(np.uint8)).sum()) == 0
    assert np.array_equal(counts, ref_counts)
    assert np.array_equal(active, O.port_marking(ref_masks, ref_counts))
    assert O.rel_linf(out, O.port_denoise(**inp, m=1.0)) <= TOL


def test_dense_pair_kernel_path(capi, crop_a, monkeypatch):
    monkeypatch.setenv("BCD_B200_DENSE_PAIR", "1")
    h, _ = _run(capi, inputs_of(crop_a), m=0.0)
    masks, counts = h.read_selection()
    h.close()
    monkeypatch.setenv("BCD_B200_DENSE_PAIR", "0")
    capi.Handle(0).close()  # reset the process-wide hook
    assert int(np.unpackbits((masks ^ crop_a["port_masks"]).view(np.uint8)).sum()) == 0


def test_handle_reuse_across_frame_sizes_and_parameters(capi, crop_a, crop_b, crop_c):
    """One handle, frames of different sizes / radii / scales back to back (buffers and streams persist and are
    re-dimensioned): every result still matches its reference golden."""
    h = capi.Handle(0)
    try:
        for crop, key, kw, streamed in ((crop_a, "ref_s3_m1", dict(m=1.0, scales=3), False),
                                        (crop_b, "ref_s2_m0", dict(m=0.0, scales=2), True),
                                        (crop_c, "ref_s1_m0_w2b3", dict(m=0.0, w=2, b=5), False),
                                        (crop_b, "ref_s1_m0_b4", dict(m=0.0, scales=1, b=4), True),
                                        (crop_a, "ref_s1_m1", dict(m=1.0, scales=1), True)):
            inp = inputs_of(crop)
            (h.upload_async if streamed else h.upload)(**inp)
            h.run(**kw)
            err = O.rel_linf(h.download(), crop[key])
            assert err <= TOL, "%s: rel Linf %.3e" % (key, err)
    finally:
        h.close()


def test_host_register_makes_caller_buffers_streamable(capi, crop_a):
    """bcd_b200_host_register / _unregister: page-lock the caller's own arrays, stream them, release them."""
    import ctypes as C
    inp = {k: np.ascontiguousarray(v) for k, v in inputs_of(crop_a).items()}
    L = capi.lib()
    for v in inp.values():
        assert L.bcd_b200_host_register(C.c_void_p(v.ctypes.data), v.nbytes) == 0
        assert L.bcd_b200_host_register(C.c_void_p(v.ctypes.data), v.nbytes) == 0  # idempotent
    try:
        h = capi.Handle(0)
        h.upload_async(**inp)
        h.run(m=1.0, scales=1)
        out = h.download()
        h.close()
    finally:
        for v in inp.values():
            assert L.bcd_b200_host_unregister(C.c_void_p(v.ctypes.data)) == 0
    assert O.rel_linf(out, crop_a["ref_s1_m1"]) <= TOL
    assert L.bcd_b200_host_register(None, 16) != 0


def test_streamed_upload_matches_serial_upload(capi, crop_a):
    """bcd_b200_upload_async: level-0 pair terms launched band by band behind the arriving histogram rows — same
    selection bits and the same output as the serial upload, and within tolerance of the reference golden."""
    inp = inputs_of(crop_a)
    h = capi.Handle(0)
    h.upload_async(**inp)
    h.run(m=1.0, scales=3)
    out_a = h.download()
    masks_a, counts_a = h.read_selection()
    h.upload(**inp)
    h.run(m=1.0, scales=3)
    out_s = h.download()
    masks_s, counts_s = h.read_selection()
    h.upload_async(**inp)  # a pending upload consumed by the spike prefilter / input read-back instead of a run
    back = h.download_inputs()
    h.close()
    assert np.array_equal(masks_a, masks_s) and np.array_equal(counts_a, counts_s)
    assert O.rel_linf(out_a, out_s) <= 1e-5  # the scatter accumulation uses float atomics
    assert O.rel_linf(out_a, crop_a["ref_s3_m1"]) <= TOL
    assert np.array_equal(back["hist"], inp["hist"]) and np.array_equal(back["color"], inp["color"])


@pytest.mark.parametrize("size", [(640, 400), (333, 517)])
def test_streamed_upload_streams_the_coarser_levels_too(capi, size, monkeypatch):
    """Under bcd_b200_upload_async the selection of the coarser pyramid levels also runs band by band behind the arriving
    rows (bcd_b200_capi.cu: coarseBand), on frames tall enough for a dozen upload chunks and odd sizes: identical work
    counters on every level, identical level-0 selection, same output as the serial upload and as the A/B hook."""
    from bcd_b200 import host
    W, H = size
    inp = host.synth(W, H, spp=16)
    got = {}
    for mode in ("serial", "streamed", "streamed_no_coarse"):
        monkeypatch.setenv("BCD_B200_STREAM_COARSE", "0" if mode == "streamed_no_coarse" else "1")
        h = capi.Handle(0)
        try:
            (h.upload if mode == "serial" else h.upload_async)(**inp)
            h.run(m=1.0, scales=3)
            got[mode] = (h.download(), h.stats(), h.read_selection())
        finally:
            h.close()
    monkeypatch.setenv("BCD_B200_STREAM_COARSE", "1")
    capi.Handle(0).close()  # reset the process-wide hook
    for mode in ("streamed", "streamed_no_coarse"):
        for key in ("interior", "strong", "active", "full", "fallback", "sum_n_full"):
            assert got[mode][1][key] == got["serial"][1][key], (mode, key)
        assert np.array_equal(got[mode][2][0], got["serial"][2][0]) and np.array_equal(got[mode][2][1], got["serial"][2][1])
        assert O.rel_linf(got[mode][0], got["serial"][0]) <= 1e-5  # float atomics only


@pytest.mark.parametrize("group", [8, 16, 32])
def test_eig_chain_layouts_agree_with_reference(capi, crop_a, crop_b, group, monkeypatch):
    """K3e: 4 / 2 / 1 matrices per warp (k_eig_sub.cu, k_filter.cu) — each within tolerance of the reference golden,
    also on the odd-sized crop whose full list is not a multiple of the matrices per warp."""
    monkeypatch.setenv("BCD_B200_EIG_GROUP", str(group))
    try:
        h, out = _run(capi, inputs_of(crop_a), m=1.0, scales=1)
        h.close()
        assert O.rel_linf(out, crop_a["ref_s1_m1"]) <= TOL
        h, out = _run(capi, inputs_of(crop_b), m=0.0, scales=2)
        h.close()
        assert O.rel_linf(out, crop_b["ref_s2_m0"]) <= TOL
    finally:
        monkeypatch.setenv("BCD_B200_EIG_GROUP", "16")
        capi.Handle(0).close()  # reset the process-wide hook to the default


@pytest.mark.parametrize("b", [10, 3, 14])
def test_other_search_radii_bit_exact(capi, crop_b, b):
    inp = inputs_of(crop_b)
    h, out = _run(capi, inp, m=1.0, b=b)
    masks, counts = h.read_selection()
    active = h.read_active()
    h.close()
    ref_masks, ref_counts = O.port_select(inp["hist"], inp["nsamples"], b=b)
    assert int(np.unpackbits((masks ^ ref_masks).view(np.uint8)).sum()) == 0
    assert np.array_equal(active, O.port_marking(ref_masks, ref_counts, b=b))
    assert O.rel_linf(out, O.port_denoise(**inp, m=1.0, b=b)) <= TOL


@pytest.mark.parametrize("size", [(1003, 205), (260, 517), (64, 40)])
def test_marking_on_large_frames_both_kernels(capi, size, monkeypatch):
    """K2 alone at sizes with many row bands (the crops of the reference fixture have three): the GPU's own selection
    masks go through the restated sequential process (DenoisingUnit.cpp:164-173,690 in raster order) on the CPU and the
    active map must match bit for bit — for the systolic kernel (default at b = 6) and the shared-memory wavefront
    (BCD_B200_MARKING=v4), widths that are not a multiple of 4 and heights that are not a multiple of 32."""
    from bcd_b200 import host
    W, H = size
    inp = host.synth(W, H, spp=16)
    got = {}
    try:
        for kern in ("sys", "v4"):
            monkeypatch.setenv("BCD_B200_MARKING", kern)
            h, out = _run(capi, inp, m=1.0)
            masks, counts = h.read_selection()
            active = h.read_active()
            st = h.stats()
            h.close()
            ref = O.port_marking(masks, counts)
            bad = np.argwhere(active != ref)
            assert bad.shape[0] == 0, "%s: %d marking mismatches, first (row,col) %s" % (kern, bad.shape[0], bad[:8].tolist())
            assert st["active"] == int(ref.sum()) and st["full"] == int((ref.astype(bool) & (counts >= 28)).sum())
            got[kern] = (active, out)
        assert np.array_equal(got["sys"][0], got["v4"][0])
        # bands per CTA of the systolic kernel: 1 (every hand-off through the global ring), 2, 8 (default 4)
        monkeypatch.setenv("BCD_B200_MARKING", "sys")
        for nw in ("1", "2", "8"):
            monkeypatch.setenv("BCD_B200_MARKING_NW", nw)
            h, _ = _run(capi, inp, m=1.0)
            active = h.read_active()
            h.close()
            assert np.array_equal(active, got["sys"][0]), "bands per CTA = %s" % nw
        monkeypatch.setenv("BCD_B200_MARKING_NW", "4")
        # fractional m: both kernels use the same counter-based hash, so they must agree with each other
        frac = {}
        for kern in ("sys", "v4"):
            monkeypatch.setenv("BCD_B200_MARKING", kern)
            h, _ = _run(capi, inp, m=0.5)
            frac[kern] = h.read_active()
            h.close()
        assert np.array_equal(frac["sys"], frac["v4"])
    finally:
        monkeypatch.setenv("BCD_B200_MARKING", "sys")
        capi.Handle(0).close()  # reset the process-wide hook


def test_fractional_skip_probability_is_deterministic(capi, crop_b):
    inp = inputs_of(crop_b)
    h, out1 = _run(capi, inp, m=0.5)
    s1 = h.stats()
    h.close()
    h, out2 = _run(capi, inp, m=0.5)
    s2 = h.stats()
    h.close()
    assert s1 == s2 and O.rel_linf(out1, out2) <= 1e-5
    h, _ = _run(capi, inp, m=1.0)
    sm1 = h.stats()
    h.close()
    h, _ = _run(capi, inp, m=0.0)
    sm0 = h.stats()
    h.close()
    assert sm1["active"] <= s1["active"] <= sm0["active"]


def test_spike_removal_prefilter_bit_exact(capi):
    """Row f1 on the GPU: bit-exact against the stored reference output (GCC-build semantics) and against the
    restatement in float-abs mode; the filtered inputs stay resident and feed the denoiser."""
    import os
    from conftest import GOLDEN
    from bcd_b200 import host
    gold = dict(np.load(os.path.join(GOLDEN, "spike_synth_96x64.npz")))
    g = host.synth(96, 64, spp=32)
    h = capi.Handle(0)
    h.upload(**g)
    h.spike_removal(2.0, int_abs=True)
    f = h.download_inputs()
    assert np.array_equal(f["color"].view(np.uint32), gold["color"].view(np.uint32))
    assert np.array_equal(f["cov"].view(np.uint32), gold["cov"].view(np.uint32))
    assert np.array_equal(f["nsamples"], gold["nsamples"])
    port, _ = O.port_spike_removal(**g, factor=2.0, int_abs=True)
    assert np.array_equal(f["hist"].view(np.uint32), port["hist"].view(np.uint32))
    h.run(m=1.0, scales=2)
    out = h.download()
    assert O.rel_linf(out, gold["ref_denoised_after_spike"]) <= TOL
    # float-abs mode, starting again from the raw inputs
    h.upload(**g)
    h.spike_removal(2.0, int_abs=False)
    ff = h.download_inputs()
    h.close()
    portf, _ = O.port_spike_removal(**g, factor=2.0, int_abs=False)
    for k in ff:
        assert np.array_equal(ff[k].view(np.uint32), portf[k].view(np.uint32)), k

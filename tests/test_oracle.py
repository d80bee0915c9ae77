"""CPU: pins the oracles. The reference ships no tests or golden vectors, so
  * the C restatement (oracle/liboracle.so) is checked against the committed outputs of the UNMODIFIED reference
    (tests/golden/*.npz, produced by tests/golden/make_golden.py from data/raw/test.raw) and, where the compiled
    reference is present (oracle/_ref), against it live on seeded synthetic inputs;
  * its discrete stages are checked for the invariants the GPU kernels rely on (symmetry, closed-form marking)."""
import numpy as np
import pytest

from conftest import inputs_of
import oracles as O

TOL = 1e-4


def test_port_matches_reference_goldens(crop_a):
    inp = inputs_of(crop_a)
    out, st = O.port_denoise(**inp, m=1.0, want_stats=True)
    assert O.rel_linf(out, crop_a["ref_s1_m1"]) <= TOL
    # closed-form marking == literal sequential process
    act = crop_a["port_active_m1"].astype(bool)
    strong = crop_a["port_counts"] >= 28
    assert st["active"] == int(act.sum()) and st["full"] == int((act & strong).sum())
    assert st["fallback"] == int((act & ~strong).sum())
    assert O.rel_linf(O.port_denoise(**inp, m=0.0), crop_a["ref_s1_m0"]) <= TOL


def test_port_multiscale_matches_reference_goldens(crop_a, crop_b):
    assert O.rel_linf(O.port_denoise(**inputs_of(crop_a), m=1.0, scales=3), crop_a["ref_s3_m1"]) <= TOL
    assert O.rel_linf(O.port_denoise(**inputs_of(crop_b), m=0.0, scales=2), crop_b["ref_s2_m0"]) <= TOL  # odd sizes
    assert O.rel_linf(O.port_denoise(**inputs_of(crop_b), m=1.0, tau=2.0), crop_b["ref_s1_m1_tau2"]) <= TOL


def test_port_wide_patch(crop_c):
    assert O.rel_linf(O.port_denoise(**inputs_of(crop_c), m=0.0, w=2, b=5), crop_c["ref_s1_m0_w2b3"]) <= TOL


def test_selection_is_symmetric_and_matches_golden(crop_a):
    masks, counts = O.port_select(crop_a["hist"], crop_a["nsamples"])
    assert np.array_equal(masks, crop_a["port_masks"]) and np.array_equal(counts, crop_a["port_counts"])
    H, W = counts.shape
    bits = np.unpackbits(masks.view(np.uint8), axis=2, bitorder="little")[:, :, :169].reshape(H, W, 13, 13)
    # q in S(p)  <=>  p in S(q)
    for dy in range(-6, 7):
        for dx in range(-6, 7):
            a = bits[max(0, -dy):H - max(0, dy), max(0, -dx):W - max(0, dx), dy + 6, dx + 6]
            b = bits[max(0, dy):H - max(0, -dy), max(0, dx):W - max(0, -dx), 6 - dy, 6 - dx]
            assert np.array_equal(a, b)
    assert bits[1:-1, 1:-1, 6, 6].all()  # p in S(p) everywhere on this crop


def test_eigh_restatement_against_lapack():
    rng = np.random.default_rng(1)
    for n in (3, 27, 75):
        X = rng.standard_normal((n, 2 * n)).astype(np.float32)
        A = (X @ X.T / (2 * n)).astype(np.float32)
        A -= np.float32(0.7) * np.eye(n, dtype=np.float32)  # indefinite, like C - Sigma_n
        rc, ev, V = O.port_eigh(A)
        assert rc == 0
        evn = np.linalg.eigvalsh(A.astype(np.float64))
        assert np.abs(ev - evn).max() <= 2e-5 * np.abs(evn).max()
        assert np.abs(V @ np.diag(ev) @ V.T - A).max() <= 2e-5 * np.abs(A).max()
        assert np.abs(V.T @ V - np.eye(n)).max() <= 2e-5


@pytest.mark.skipif(not O.have_ref(), reason="compiled reference (oracle/_ref) not built here")
def test_port_matches_live_reference_on_synthetic():
    from bcd_b200 import host
    inp = host.synth(72, 60, spp=32)
    for kw in (dict(m=1.0, scales=1), dict(m=0.0, scales=2)):
        ref = O.ref_denoise(**inp, cores=1, **kw)
        assert O.rel_linf(O.port_denoise(**inp, **kw), ref) <= TOL


def test_degenerate_inputs_produce_nan_like_the_reference():
    # a 2x2 image has no interior at w=1: count == 0 everywhere -> inf * 0 = NaN (Denoiser.cpp:463-466)
    z = dict(color=np.ones((2, 2, 3), np.float32), nsamples=np.full((2, 2, 1), 4, np.float32),
             hist=np.ones((2, 2, 60), np.float32), cov=np.ones((2, 2, 6), np.float32))
    out = O.port_denoise(**z, m=0.0)
    assert np.isnan(out).all()
    if O.have_ref():
        assert np.isnan(O.ref_denoise(**z, cores=1, m=0.0)).all()


def test_spike_removal_restatement_matches_reference_golden():
    """Row f1: the reference's unqualified abs() binds to ::abs(int) under GCC/glibc (the compiled oracle); the
    restatement's int_abs mode must reproduce the stored reference output bit for bit."""
    import os
    from conftest import GOLDEN
    from bcd_b200 import host
    gold = dict(np.load(os.path.join(GOLDEN, "spike_synth_96x64.npz")))
    g = host.synth(96, 64, spp=32)
    p, src = O.port_spike_removal(**g, factor=2.0, int_abs=True)
    assert np.array_equal(p["color"].view(np.uint32), gold["color"].view(np.uint32))
    assert np.array_equal(p["cov"].view(np.uint32), gold["cov"].view(np.uint32))
    assert np.array_equal(p["nsamples"], gold["nsamples"])
    assert np.allclose(p["hist"].astype(np.float64).sum(axis=(0, 1)), gold["hist_channel_sums"], rtol=0, atol=0)
    assert gold["changed"].sum() > 0 and np.array_equal(src != np.arange(96 * 64).reshape(64, 96), gold["changed"])
    # the float-abs reading replaces many more pixels (MSVC builds of the reference behave like this)
    pf, srcf = O.port_spike_removal(**g, factor=2.0, int_abs=False)
    assert (srcf != np.arange(96 * 64).reshape(64, 96)).sum() > gold["changed"].sum()
    if O.have_ref():
        r = O.ref_spike_removal(**g, factor=2.0)
        for k in r:
            assert np.array_equal(r[k].view(np.uint32), p[k].view(np.uint32)), k


def test_pair_term_planes_reproduce_the_golden_selection(crop_a):
    """bcdo_pair_terms_all (the layout of the GPU intermediate) is pinned by recomputing the selection from it: the
    row-major chain sum of the 3x3 pair terms / count <= tau (DenoisingUnit.cpp:336-358) must give the committed masks."""
    from conftest import inputs_of
    inp = inputs_of(crop_a)
    b, w, tau = 6, 1, np.float32(1.0)
    sums, counts = O.port_pair_terms(inp["hist"], inp["nsamples"], b=b)
    H, W = counts.shape[1:]
    side = 2 * b + 1
    masks = crop_a["port_masks"]
    checked = 0
    for j in (0, 1, 7, 13, 40, 84):
        idx = b * side + b + j
        dy, dx = idx // side - b, idx % side - b
        ys = slice(w, H - w - dy)
        x0, x1 = max(w, w - dx), min(W - w, W - w - dx)
        acc = np.zeros((H - 2 * w - dy, x1 - x0), np.float32)
        cnt = np.zeros(acc.shape, np.int32)
        for oy in range(-w, w + 1):          # same order as the reference's patch iterator: row-major
            for ox in range(-w, w + 1):
                acc = acc + sums[j, w + oy:H - w - dy + oy, x0 + ox:x1 + ox]
                cnt = cnt + counts[j, w + oy:H - w - dy + oy, x0 + ox:x1 + ox]
        with np.errstate(invalid="ignore", divide="ignore"):
            sel = (acc / cnt.astype(np.float32)) <= tau
        got = (masks[ys, x0:x1, idx >> 5] >> np.uint32(idx & 31)) & 1
        assert np.array_equal(sel, got.astype(bool)), "offset %d" % j
        checked += sel.size
    assert checked > 50000


def test_port_matches_reference_golden_at_c1_scale():
    """BASELINE.json configs[0], G(256,256) s=1 m=1: the C restatement against the stored output of the compiled
    reference (tests/golden/make_golden_scale.py); inputs regenerated by the seeded generator."""
    import os
    from bcd_b200 import host
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scale_c1_256.npz"))
    c1 = host.synth(256, 256, spp=64)
    chk = np.array([float(np.asarray(c1[k], np.float64).sum()) for k in ("color", "nsamples", "hist", "cov")])
    assert np.allclose(chk, g["checksum"], rtol=1e-12)
    out = O.port_denoise(**c1, m=1.0, scales=1)
    assert O.rel_linf(out, g["ref_s1_m1"]) <= 1e-5

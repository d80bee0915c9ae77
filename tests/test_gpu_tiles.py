"""GPU: the row-tile path (bcd_b200_tile_* + bcd_b200/tiles.py) reproduces the full-frame result. All tiles run on
one GPU here (loopback exchange); the multi-process NCCL path moves the same strips (tests/test_tiles_cpu.py checks the
exchange logic with gloo)."""
import numpy as np
import pytest

import oracles as O

pytestmark = pytest.mark.gpu


def _full(capi, inp, **kw):
    h = capi.Handle(0)
    h.upload(**inp)
    h.run(**kw)
    out, st = h.download(), h.stats()
    h.close()
    return out, st


def _tiled(capi, tiles, inp, count, streamed=False, **kw):
    import torch
    params = capi.make_params(**kw)
    # all tiles exchange strips on shared streams: one main stream, one marking stream per level
    stream, aux = torch.cuda.Stream(), tiles.make_aux_streams(torch.device("cuda", 0), params.nb_scales)
    with torch.cuda.stream(stream):
        workers = [tiles.make_worker(inp, params, i, count, main_stream=stream, aux_stream=aux, streamed=streamed) for i in range(count)]
        tiles.run_tiles_loopback(workers)
        out = np.concatenate([w.download() for w in workers], axis=0)
    st = {}
    for w in workers:
        for k, v in w.h.stats().items():
            st[k] = st.get(k, 0) + v
        w.h.close()
    return out, st


@pytest.mark.parametrize("count,kw", [
    (2, dict(m=1.0, scales=1)),
    (2, dict(m=0.0, scales=2)),
    (3, dict(m=1.0, scales=3)),
    (4, dict(m=1.0, scales=2)),
])
def test_tiles_equal_full_frame(count, kw):
    from bcd_b200 import capi, host, tiles
    inp = host.synth(120, 208, spp=24)  # W=120, H=208
    ref, st_ref = _full(capi, inp, **kw)
    out, st = _tiled(capi, tiles, inp, count, **kw)
    for k in ("interior", "strong", "active", "full", "fallback", "sum_n_full"):
        assert st[k] == st_ref[k], (k, st[k], st_ref[k])
    assert O.rel_linf(out, ref) <= 2e-6  # only the order of the float atomics differs


def test_tiles_with_streamed_upload_equal_full_frame():
    """Row tiles fed by bcd_b200_upload_async: the banded level-0 selection is clipped to the tile's selection rows and
    the pyramid is built on demand — same counters and image as the full-frame run."""
    from bcd_b200 import capi, host, tiles
    inp = host.synth(120, 208, spp=24)
    kw = dict(m=1.0, scales=3)
    ref, st_ref = _full(capi, inp, **kw)
    out, st = _tiled(capi, tiles, inp, 3, streamed=True, **kw)
    for k in ("interior", "strong", "active", "full", "fallback", "sum_n_full"):
        assert st[k] == st_ref[k], (k, st[k], st_ref[k])
    assert O.rel_linf(out, ref) <= 2e-6


def test_tiles_match_reference_golden(crop_a):
    from conftest import inputs_of
    from bcd_b200 import capi, tiles
    inp = inputs_of(crop_a)
    out, _ = _tiled(capi, tiles, inp, 2, m=1.0, scales=1)
    assert O.rel_linf(out, crop_a["ref_s1_m1"]) <= 1e-4


def test_too_small_tiles_are_rejected():
    from bcd_b200 import capi, tiles
    with pytest.raises(capi.BcdError):
        tiles.plan(64, 96, 0, 8, capi.make_params(scales=3))


# ------------------------------------------------------------------------------------------- peer mode / group API
def _counters_equal(a, b):
    for k in ("interior", "strong", "active", "full", "fallback", "sum_n_full"):
        assert a[k] == b[k], (k, a[k], b[k])


@pytest.mark.parametrize("count,kw", [
    (2, dict(m=1.0, scales=1)),
    (2, dict(m=0.0, scales=2)),
    (3, dict(m=1.0, scales=3)),
    (4, dict(m=1.0, scales=2)),
    (2, dict(m=1.0, scales=2, b=4)),          # marking kernel without a ring: the AS rows travel as a strip
    (2, dict(m=1.0, scales=1, w=2, b=5)),
])
def test_group_tiles_on_one_device_equal_full_frame(count, kw):
    """The group API in peer mode (bcd_b200_group_denoise: one host thread per tile, strips pushed into the neighbour's
    arena and awaited on the device, marking ring chained from tile to tile) with every tile on device 0."""
    from bcd_b200 import capi, host
    inp = host.synth(120, 208, spp=24)
    ref, st_ref = _full(capi, inp, **kw)
    g = capi.Group([0] * count, tiles_per_gpu=1)
    try:
        out = g.denoise(**inp, **kw)
        assert g.tiles() == count
        _counters_equal(g.stats(), st_ref)
        assert O.rel_linf(out, ref) <= 2e-6  # only the order of the float atomics differs
        # a second, different frame through the same group: epochs advance, buffers / links / pinned ranges are reused
        inp2 = host.synth(120, 208, spp=16, seed=7)
        ref2, st_ref2 = _full(capi, inp2, **kw)
        out2 = g.denoise(**inp2, **kw)
        _counters_equal(g.stats(), st_ref2)
        assert O.rel_linf(out2, ref2) <= 2e-6
        # another geometry: the tiles are set up and connected again
        inp3 = host.synth(96, 160, spp=16, seed=11)
        ref3, _ = _full(capi, inp3, **kw)
        assert O.rel_linf(g.denoise(**inp3, **kw), ref3) <= 2e-6
    finally:
        g.close()


@pytest.mark.parametrize("gpus,per,kw", [
    (2, 2, dict(m=1.0, scales=2)),
    (2, 3, dict(m=1.0, scales=1)),
    (3, 2, dict(m=0.0, scales=2)),
])
def test_group_with_several_cyclic_tiles_per_gpu(gpus, per, kw):
    """tiles_per_gpu > 1: tile j runs on GPU j mod G, so neighbouring tiles sit on different GPUs and a GPU holds several
    handles whose marking chains depend on each other through the other GPUs."""
    from bcd_b200 import capi, host
    inp = host.synth(136, 64 * gpus * per + 24, spp=16)
    ref, st_ref = _full(capi, inp, **kw)
    g = capi.Group([0] * gpus, tiles_per_gpu=per)
    try:
        for _ in range(2):
            out = g.denoise(**inp, **kw)
            assert g.tiles() == gpus * per
            _counters_equal(g.stats(), st_ref)
            assert O.rel_linf(out, ref) <= 2e-6
    finally:
        g.close()


@pytest.mark.parametrize("count,kw", [
    (3, dict(m=1.0, scales=3)),           # odd width and height: integer halving of the pyramid, clamped merge at the borders
    (2, dict(m=1.0, scales=2, b=10)),     # wide search window: warp-skewed marking kernel, AS rows travel as a strip
    (2, dict(m=1.0, scales=1, b=3)),      # small window: generic marking kernel
])
def test_group_tiles_odd_sizes_and_other_radii(count, kw):
    from bcd_b200 import capi, host
    inp = host.synth(131, 173, spp=16, seed=23)
    ref, st_ref = _full(capi, inp, **kw)
    g = capi.Group([0] * count, tiles_per_gpu=1)
    try:
        out = g.denoise(**inp, **kw)
        assert g.tiles() == count
        _counters_equal(g.stats(), st_ref)
        assert O.rel_linf(out, ref) <= 2e-6
    finally:
        g.close()


def test_group_reduces_the_tile_count_for_small_frames():
    from bcd_b200 import capi, host
    inp = host.synth(64, 96, spp=16)
    ref, _ = _full(capi, inp, m=1.0, scales=3)
    g = capi.Group([0] * 8)
    try:
        out = g.denoise(**inp, m=1.0, scales=3)
        assert 1 <= g.tiles() < 8
        assert O.rel_linf(out, ref) <= 2e-6
    finally:
        g.close()


def test_group_matches_reference_golden(crop_a):
    from conftest import inputs_of
    from bcd_b200 import capi
    out = capi.denoise_multi(**inputs_of(crop_a), devices=[0, 0], m=1.0, scales=1)
    assert O.rel_linf(out, crop_a["ref_s1_m1"]) <= 1e-4


def test_group_on_several_gpus_equals_one_gpu():
    """Real peer memory: one tile per GPU over NVLink. Skipped on a one-GPU box."""
    from bcd_b200 import capi, host
    n = capi.lib().bcd_b200_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = min(n, 8)
    inp = host.synth(640, 64 * n + 40, spp=16)
    for kw in (dict(m=1.0, scales=3), dict(m=0.0, scales=1), dict(m=1.0, scales=2, b=4)):
        ref, st_ref = _full(capi, inp, **kw)
        g = capi.Group(list(range(n)), tiles_per_gpu=1)
        try:
            for _ in range(3):
                out = g.denoise(**inp, **kw)
                assert g.tiles() == n
                _counters_equal(g.stats(), st_ref)
                assert O.rel_linf(out, ref) <= 2e-6
        finally:
            g.close()


def test_cxx_session_keeps_its_device_session_across_frames():
    """bcd::MultiscaleDenoiser as a renderer holds it: one object, pageable DeepImage storage, denoise() per frame. The
    second frame reuses handle, buffers and page-locked ranges and must give the same pixels as the C ABI."""
    from bcd_b200 import capi, host
    kw = dict(m=1.0, scales=3)
    s = host.CxxSession(scales=3, devices=[0, 0])
    try:
        for seed in (3, 5, 3):
            inp = host.synth(160, 176, spp=16, seed=seed)
            ref, _ = _full(capi, inp, **kw)
            s.load(**inp)
            assert s.denoise(m=1.0)
            assert s.devices_used() == 2
            assert O.rel_linf(s.fetch(), ref) <= 2e-6
    finally:
        s.close()


def test_peer_tiles_across_processes_equal_one_gpu():
    """One process per GPU (torchrun), arenas shared through CUDA IPC — the path bench.py's N > 1 arm takes.
    Skipped on a one-GPU box."""
    import os
    import subprocess
    import sys
    from bcd_b200 import capi
    n = min(capi.lib().bcd_b200_device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29431", os.path.join(root, "tools", "dist_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_PARITY OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])

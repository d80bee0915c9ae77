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

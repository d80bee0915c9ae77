"""TEST INFRASTRUCTURE: ctypes loaders for the two parity oracles.

* ``ref``  – oracle/_ref/libbcd_ref.so: the unmodified reference CPU path (superboubek/bcd src/core/*.cpp) compiled
             in place by oracle/Makefile.  Travels to the GPU box as a prebuilt file; never reads /root/reference
             at run time.
* ``port`` – oracle/liboracle.so: plain-C restatement of the same algorithm (oracle/bcd_oracle.c) with stage-level
             entry points (selection masks, marking, pyramid).

Nothing under bcd_b200/ imports this module; the product path never routes through an oracle.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbcd_ref.so")
PORT_SO = os.path.join(ROOT, "oracle", "liboracle.so")

_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def have_ref():
    return os.path.exists(REF_SO)


def have_port():
    return os.path.exists(PORT_SO)


_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        lib.bcdref_max_threads.restype = C.c_int
        lib.bcdref_denoise.restype = C.c_int
        lib.bcdref_denoise.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int,
                                       C.c_float, C.c_int, C.c_int, C.c_float,
                                       C.c_int, C.c_float, C.c_int, C.c_int, _fp, C.c_int]
        lib.bcdref_accumulate.restype = C.c_int
        lib.bcdref_accumulate.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                          _fp, _fp, _fp, _fp]
        lib.bcdref_spike_removal.restype = C.c_int
        lib.bcdref_spike_removal.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float]
        _ref = lib
    return _ref


def ref_denoise(color, nsamples, hist, cov, tau=1.0, w=1, b=6, eps=1e-8, random_order=False, m=1.0,
                cores=1, scales=1, quiet=True):
    """Reference Denoiser / MultiscaleDenoiser on host arrays (H,W,C) float32 -> (H,W,3)."""
    H, W, _ = color.shape
    B = hist.shape[2]
    out = np.zeros((H, W, 3), np.float32)
    rc = ref_lib().bcdref_denoise(np.ascontiguousarray(color, np.float32), np.ascontiguousarray(nsamples, np.float32),
                                  np.ascontiguousarray(hist, np.float32), np.ascontiguousarray(cov, np.float32),
                                  W, H, B, tau, w, b, eps, int(random_order), float(m), cores, scales, out,
                                  int(quiet))
    if rc != 0:
        raise RuntimeError("reference denoise() returned false")
    return out


def ref_accumulate(samples, nb_bins=20, gamma=2.2, max_value=2.5):
    """samples: (H,W,spp,3) float32 -> dict(color,nsamples,hist,cov) via the reference SamplesAccumulator."""
    H, W, spp, _ = samples.shape
    color = np.zeros((H, W, 3), np.float32)
    ns = np.zeros((H, W, 1), np.float32)
    hist = np.zeros((H, W, 3 * nb_bins), np.float32)
    cov = np.zeros((H, W, 6), np.float32)
    ref_lib().bcdref_accumulate(np.ascontiguousarray(samples, np.float32), W, H, spp, nb_bins, gamma, max_value,
                                color, ns, hist, cov)
    return dict(color=color, nsamples=ns, hist=hist, cov=cov)


def load_test_raw(path):
    """data/raw/test.raw layout (src/raw_converter/main.cpp:65-72,159-171): 5 x int32 header
    {version, W, H, spp, channels} then float RGB samples, pixel-major."""
    with open(path, "rb") as f:
        hdr = np.frombuffer(f.read(20), np.int32)
        _, W, H, spp, ch = [int(v) for v in hdr]
        data = np.frombuffer(f.read(4 * W * H * spp * ch), np.float32)  # (the file carries 4 trailing bytes)
    return data.reshape(H, W, spp, ch)


def rel_linf(a, ref):
    """max|a-ref| / max|ref| over all pixels x channels. NaNs must coincide."""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    nan_a, nan_r = np.isnan(a), np.isnan(ref)
    if (nan_a != nan_r).any():
        return float("inf")
    ok = ~nan_r
    if not ok.any():
        return 0.0
    return float(np.abs(a[ok] - ref[ok]).max() / max(np.abs(ref[ok]).max(), 1e-30))


# ------------------------------------------------------------------------------------------------ port (C restatement)
class _Params(C.Structure):
    _fields_ = [("tau", C.c_float), ("w", C.c_int), ("b", C.c_int), ("eps", C.c_float), ("marking", C.c_int)]


class _Stats(C.Structure):
    _fields_ = [(k, C.c_longlong) for k in ("interior", "active", "full", "fallback", "strong", "sum_n_full")]


_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_port = None


def port_lib():
    global _port
    if _port is None:
        lib = C.CDLL(PORT_SO)
        lib.bcdo_mask_words.restype = C.c_int
        lib.bcdo_mask_words.argtypes = [C.c_int]
        lib.bcdo_pixel_cov.argtypes = [_fp, _fp, C.c_longlong, _fp]
        lib.bcdo_select_all.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _u32p, _i32p]
        lib.bcdo_marking.argtypes = [_u32p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        lib.bcdo_pair_terms_all.argtypes = [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _u8p]
        lib.bcdo_eigh.restype = C.c_int
        lib.bcdo_eigh.argtypes = [_fp, C.c_int, _fp, _fp, _fp]
        lib.bcdo_denoise.restype = C.c_int
        lib.bcdo_denoise.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.POINTER(_Params), _fp,
                                     C.POINTER(_Stats), C.c_void_p, C.c_void_p]
        lib.bcdo_denoise_multiscale.restype = C.c_int
        lib.bcdo_denoise_multiscale.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.POINTER(_Params),
                                                C.c_int, _fp]
        lib.bcdo_downscale_sum.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _fp]
        lib.bcdo_downscale_avg.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _fp]
        lib.bcdo_downscale_cov.argtypes = [_fp, _fp, C.c_int, C.c_int, _fp]
        lib.bcdo_interpolate.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _fp, C.c_int, C.c_int]
        lib.bcdo_merge.argtypes = [_fp, C.c_int, C.c_int, _fp]
        _port = lib
    return _port


def _c(a):
    return np.ascontiguousarray(a, np.float32)


def port_select(hist, nsamples, w=1, b=6, tau=1.0):
    """-> masks (H,W,words) uint32 in window row-major bit order, counts (H,W) int32."""
    H, W, B = hist.shape
    words = port_lib().bcdo_mask_words(b)
    masks = np.zeros((H, W, words), np.uint32)
    counts = np.zeros((H, W), np.int32)
    port_lib().bcdo_select_all(_c(hist), _c(nsamples), W, H, B, w, b, tau, masks, counts)
    return masks, counts


def port_pair_terms(hist, nsamples, b=6):
    """-> sums (J,H,W) float32 (NaN where the offset leaves the image), counts (J,H,W) uint8 (255 there)."""
    H, W, B = hist.shape
    J = ((2 * b + 1) ** 2 + 1) // 2
    sums = np.empty((J, H, W), np.float32)
    counts = np.empty((J, H, W), np.uint8)
    port_lib().bcdo_pair_terms_all(_c(hist), _c(nsamples), W, H, B, b, sums, counts)
    return sums, counts


def port_marking(masks, counts, w=1, b=6):
    H, W = counts.shape
    active = np.zeros((H, W), np.uint8)
    port_lib().bcdo_marking(np.ascontiguousarray(masks), np.ascontiguousarray(counts), W, H, w, b, active)
    return active


def port_eigh(A):
    n = A.shape[0]
    ev = np.zeros(n, np.float32)
    V = np.zeros((n, n), np.float32)  # column-major storage: V.T[r, c]... returned transposed below
    work = np.zeros(n * n + 4 * n, np.float32)
    rc = port_lib().bcdo_eigh(_c(A.T), n, ev, V, work)  # column-major == transpose of row-major
    return rc, ev, V.T.copy()


def port_denoise(color, nsamples, hist, cov, tau=1.0, w=1, b=6, eps=1e-8, m=1.0, scales=1, want_stats=False):
    H, W, _ = color.shape
    B = hist.shape[2]
    out = np.zeros((H, W, 3), np.float32)
    p = _Params(tau, w, b, eps, 1 if m == 1.0 else 0)
    st = _Stats()
    if scales > 1:
        rc = port_lib().bcdo_denoise_multiscale(_c(color), _c(nsamples), _c(hist), _c(cov), W, H, B, C.byref(p),
                                                scales, out)
    else:
        rc = port_lib().bcdo_denoise(_c(color), _c(nsamples), _c(hist), _c(cov), W, H, B, C.byref(p), out,
                                     C.byref(st), None, None)
    if rc != 0:
        raise RuntimeError("port denoise failed")
    if want_stats:
        return out, {k: getattr(st, k) for k, _ in _Stats._fields_}
    return out


def ref_spike_removal(color, nsamples, hist, cov, factor=2.0):
    """Reference SpikeRemovalFilter::filter (in place in the reference) -> filtered copies."""
    H, W, _ = color.shape
    c, n, h, v = [np.ascontiguousarray(a, np.float32).copy() for a in (color, nsamples, hist, cov)]
    ref_lib().bcdref_spike_removal(c, n, h, v, W, H, h.shape[2], factor)
    return dict(color=c, nsamples=n, hist=h, cov=v)


def port_spike_removal(color, nsamples, hist, cov, factor=2.0, int_abs=True):
    H, W, _ = color.shape
    c, n, h, v = [np.ascontiguousarray(a, np.float32).copy() for a in (color, nsamples, hist, cov)]
    src = np.zeros((H, W), np.int32)
    L = port_lib()
    L.bcdo_spike_removal.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _i32p]
    L.bcdo_spike_removal(c, n, h, v, W, H, h.shape[2], factor, int(int_abs), src)
    return dict(color=c, nsamples=n, hist=h, cov=v), src

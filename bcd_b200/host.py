"""ctypes binding of libbcdcore_b200.so (include/bcd_b200_host.h): the C++ drop-in classes bcd::Denoiser /
bcd::MultiscaleDenoiser driven like a host renderer would, the re-implemented bcd::SamplesAccumulator, and the
seeded synthetic-input generator G(W,H,spp,seed,sigma) of SURVEY §8d used by bench.py and the tests."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbcdcore_b200.so")
SYMBOLS = ["bcd_synth_generate", "bcd_host_accumulate", "bcd_host_cxx_denoise", "bcd_host_exr_load", "bcd_host_exr_fetch",
           "bcd_host_exr_write", "bcd_host_exr_write_ex", "bcd_host_exr_last_error", "bcd_host_params_load", "bcd_host_params_write"]
SYMBOLS += ["bcd_host_session_create", "bcd_host_session_destroy", "bcd_host_session_load", "bcd_host_session_denoise",
            "bcd_host_session_fetch", "bcd_host_session_nb_devices_used"]
SYNTH_LIB_PATH = os.path.join(_HERE, "libbcdsynth.so")
DEFAULT_SEED = 20170627
_lib = None
_synth_lib = None


def synth_lib():
    """The generator alone (libbcdsynth.so: no CUDA library behind it) — what bench.py's reference arm loads."""
    global _synth_lib
    if _synth_lib is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise RuntimeError("%s is missing: build it with `make`" % SYNTH_LIB_PATH)
        L = C.CDLL(SYNTH_LIB_PATH)
        L.bcd_synth_generate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_int, C.c_float, C.c_float,
                                         C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _synth_lib = L
    return _synth_lib


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: build it with `make`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        fp = C.c_void_p
        L.bcd_synth_generate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_int, C.c_float, C.c_float,
                                         C.c_int, C.c_int, fp, fp, fp, fp]
        L.bcd_host_accumulate.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, fp, fp, fp, fp]
        L.bcd_host_cxx_denoise.argtypes = [fp, fp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float,
                                           C.c_int, C.c_float, C.c_int, C.c_int, fp, C.POINTER(C.c_int)]
        L.bcd_host_exr_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.bcd_host_exr_fetch.argtypes = [fp]
        L.bcd_host_exr_write.argtypes = [C.c_char_p, C.c_int, fp, C.c_int, C.c_int, C.c_int]
        L.bcd_host_exr_write_ex.argtypes = [C.c_char_p, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.bcd_host_exr_last_error.restype = C.c_char_p
        L.bcd_host_params_load.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        L.bcd_host_params_write.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_char_p, C.c_char_p, C.c_char_p]
        L.bcd_host_session_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]
        L.bcd_host_session_create.restype = C.c_void_p
        L.bcd_host_session_destroy.argtypes = [C.c_void_p]
        L.bcd_host_session_destroy.restype = None
        L.bcd_host_session_load.argtypes = [C.c_void_p, fp, fp, fp, fp, C.c_int, C.c_int, C.c_int]
        L.bcd_host_session_denoise.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float]
        L.bcd_host_session_fetch.argtypes = [C.c_void_p, fp]
        L.bcd_host_session_nb_devices_used.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def synth(W, H, spp=64, seed=DEFAULT_SEED, sigma=0.6, nb_bins=20, gamma=2.2, max_value=2.5, row0=0, rows=None, out=None):
    """G(W,H,spp,seed,sigma) -> dict(color (rows,W,3), nsamples (rows,W,1), hist (rows,W,3*nb_bins), cov (rows,W,6)).
    `out` may hold preallocated arrays (e.g. numpy views of pinned torch tensors)."""
    rows = H - row0 if rows is None else rows
    if out is None:
        out = dict(color=np.empty((rows, W, 3), np.float32), nsamples=np.empty((rows, W, 1), np.float32),
                   hist=np.empty((rows, W, 3 * nb_bins), np.float32), cov=np.empty((rows, W, 6), np.float32))
    synth_lib().bcd_synth_generate(W, H, spp, seed, sigma, nb_bins, gamma, max_value, row0, rows,
                                   out["color"].ctypes.data, out["nsamples"].ctypes.data, out["hist"].ctypes.data,
                                   out["cov"].ctypes.data)
    return out


class CxxSession:
    """A host renderer's view of the drop-in across frames: ONE bcd::Denoiser / bcd::MultiscaleDenoiser object and one
    set of DeepImages in pageable std::vector storage (bcd_host_session_*)."""

    def __init__(self, scales=1, devices=None, pin=True):
        if devices is None:
            ids, n = None, 0
        elif devices == "all":
            ids, n = None, -1
        else:
            ids, n = (C.c_int * len(devices))(*devices), len(devices)
        self._s = lib().bcd_host_session_create(int(scales), ids, n, int(bool(pin)))
        self.shape = None

    def load(self, color, nsamples, hist, cov):
        color = np.ascontiguousarray(color, np.float32)
        H, W, _ = color.shape
        ns = np.ascontiguousarray(nsamples, np.float32)
        hist = np.ascontiguousarray(hist, np.float32)
        cov = np.ascontiguousarray(cov, np.float32)
        lib().bcd_host_session_load(self._s, color.ctypes.data, ns.ctypes.data, hist.ctypes.data, cov.ctypes.data, W, H, hist.shape[2])
        self.shape = (H, W)

    def denoise(self, tau=1.0, w=1, b=6, eps=1e-8, random_order=False, m=1.0):
        return bool(lib().bcd_host_session_denoise(self._s, tau, w, b, eps, int(random_order), m))

    def fetch(self):
        out = np.empty(self.shape + (3,), np.float32)
        lib().bcd_host_session_fetch(self._s, out.ctypes.data)
        return out

    def devices_used(self):
        return int(lib().bcd_host_session_nb_devices_used(self._s))

    def close(self):
        if self._s:
            lib().bcd_host_session_destroy(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def accumulate(samples, nb_bins=20, gamma=2.2, max_value=2.5):
    samples = np.ascontiguousarray(samples, np.float32)
    H, W, spp, _ = samples.shape
    out = dict(color=np.empty((H, W, 3), np.float32), nsamples=np.empty((H, W, 1), np.float32),
               hist=np.empty((H, W, 3 * nb_bins), np.float32), cov=np.empty((H, W, 6), np.float32))
    lib().bcd_host_accumulate(samples.ctypes.data, W, H, spp, nb_bins, gamma, max_value, out["color"].ctypes.data,
                              out["nsamples"].ctypes.data, out["hist"].ctypes.data, out["cov"].ctypes.data)
    return out


def cxx_denoise(color, nsamples, hist, cov, tau=1.0, w=1, b=6, eps=1e-8, random_order=True, m=1.0, scales=1, device=0):
    """Runs bcd::Denoiser / bcd::MultiscaleDenoiser (the C++ drop-in) -> (ok, out, progress_calls)."""
    color = np.ascontiguousarray(color, np.float32)
    H, W, _ = color.shape
    ns = np.ascontiguousarray(nsamples, np.float32)
    hist = np.ascontiguousarray(hist, np.float32)
    cov = np.ascontiguousarray(cov, np.float32)
    out = np.zeros((H, W, 3), np.float32)
    calls = C.c_int(0)
    ok = lib().bcd_host_cxx_denoise(color.ctypes.data, ns.ctypes.data, hist.ctypes.data, cov.ctypes.data, W, H,
                                    hist.shape[2], tau, w, b, eps, int(random_order), m, scales, device,
                                    out.ctypes.data, C.byref(calls))
    return bool(ok), out, calls.value


# ---- EXR files and .bcd.json presets through bcd::ImageIO / bcd::ParametersIO (include/bcd/io/*.h) ----------------------
class ExrError(RuntimeError):
    pass


def exr_load(path, multi_channel):
    """bcd::ImageIO::loadMultiChannelsEXR (Bin_%04d channels) or loadEXR (R, G, B) -> (H, W, depth) float32."""
    w, h, d = C.c_int(), C.c_int(), C.c_int()
    if not lib().bcd_host_exr_load(os.fsencode(path), int(bool(multi_channel)), C.byref(w), C.byref(h), C.byref(d)):
        raise ExrError(lib().bcd_host_exr_last_error().decode("utf-8", "replace"))
    out = np.empty((h.value, w.value, d.value), np.float32)
    lib().bcd_host_exr_fetch(out.ctypes.data)
    return out


def exr_write(path, image, multi_channel):
    """bcd::ImageIO::writeMultiChannelsEXR (FLOAT Bin_%04d, ZIP) or writeEXR (HALF RGBA, ZIP) of an (H, W, depth) array."""
    image = np.ascontiguousarray(image, np.float32)
    H, W, D = image.shape
    if not lib().bcd_host_exr_write(os.fsencode(path), int(bool(multi_channel)), image.ctypes.data, W, H, D):
        raise ExrError(lib().bcd_host_exr_last_error().decode("utf-8", "replace"))


def exr_write_ex(path, image, multi_channel, pixel_type, compression):
    """The codec's writer with an explicit pixel type (1 HALF, 2 FLOAT) and compression (0 NONE, 2 ZIPS, 3 ZIP, 4 PIZ)."""
    image = np.ascontiguousarray(image, np.float32)
    H, W, D = image.shape
    if not lib().bcd_host_exr_write_ex(os.fsencode(path), int(bool(multi_channel)), image.ctypes.data, W, H, D,
                                       int(pixel_type), int(compression)):
        raise ExrError("exr_write_ex failed")


PARAM_KEYS = ("performSpikeRemovalPrefiltering", "spikeRemovalThresholdStDevFactor", "nbOfScales", "histoDistanceThreshold",
              "useCuda", "nbOfCores", "patchRadius", "searchWindowRadius", "randomPixelOrder",
              "markedPixelsSkippingProbability", "minEigenValue")
PARAM_DEFAULTS = (1.0, 1.5, 3.0, 1.0, 1.0, 0.0, 1.0, 6.0, 1.0, 1.0, 1e-8)  # ParametersIO.h / IDenoiser.h defaults


def params_load(path, defaults=PARAM_DEFAULTS):
    """bcd::ParametersIO::load -> dict of the 11 numeric fields + the three input file names, or None on failure."""
    v = (C.c_double * 11)(*defaults)
    bufs = [C.create_string_buffer(4096) for _ in range(3)]
    if not lib().bcd_host_params_load(os.fsencode(path), v, bufs[0], bufs[1], bufs[2], 4096):
        return None
    out = dict(zip(PARAM_KEYS, list(v)))
    out.update(inputColorFile=bufs[0].value.decode(), inputHistoFile=bufs[1].value.decode(), inputCovarFile=bufs[2].value.decode())
    return out


def params_write(path, values, color="", hist="", cov=""):
    v = (C.c_double * 11)(*[float(values[k]) for k in PARAM_KEYS])
    return bool(lib().bcd_host_params_write(os.fsencode(path), v, os.fsencode(color), os.fsencode(hist), os.fsencode(cov)))

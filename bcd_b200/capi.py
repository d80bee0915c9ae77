"""ctypes binding of the C ABI declared in include/bcd_b200.h (libbcd_b200.so: sm_100a kernels + host engine).

This is the only way Python reaches the kernels; there is no Python or CPU fallback. Importing works anywhere (so the
CPU test-suite can check that the library loads and exports every declared symbol); creating a handle without a CUDA
device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbcd_b200.so")

STAGE_NAMES = ["prepass", "pair", "select", "marking", "fallback", "filter_stats", "filter_eig", "filter_apply",
               "finalize", "pyramid", "exchange"]

# every symbol include/bcd_b200.h declares (tests/test_capi_symbols.py cross-checks this list against the header)
SYMBOLS = [
    "bcd_b200_default_params", "bcd_b200_last_error", "bcd_b200_abi_version", "bcd_b200_device_count",
    "bcd_b200_denoise", "bcd_b200_create", "bcd_b200_destroy", "bcd_b200_set_stream", "bcd_b200_set_streams", "bcd_b200_upload",
    "bcd_b200_upload_async", "bcd_b200_host_register", "bcd_b200_host_unregister",
    "bcd_b200_bind_device", "bcd_b200_run", "bcd_b200_download", "bcd_b200_result_device_ptr", "bcd_b200_get_stats",
    "bcd_b200_get_timings", "bcd_b200_mask_words", "bcd_b200_debug_read_selection", "bcd_b200_debug_read_active", "bcd_b200_debug_read_pair_terms",
    "bcd_b200_spike_removal", "bcd_b200_spike_removal_host", "bcd_b200_download_inputs",
    # row-tile API (multi-GPU), wrapped by bcd_b200/tiles.py
    "bcd_b200_set_level_aux_stream", "bcd_b200_tile_mark_enqueue", "bcd_b200_tile_mark_wait",
    "bcd_b200_tile_plan", "bcd_b200_tile_begin", "bcd_b200_tile_select", "bcd_b200_tile_mark", "bcd_b200_tile_filter",
    "bcd_b200_tile_finalize", "bcd_b200_tile_merge", "bcd_b200_tile_xfer_bytes", "bcd_b200_tile_export",
    "bcd_b200_tile_import", "bcd_b200_tile_end", "bcd_b200_tile_download",
    "bcd_b200_accumulate_samples",
    # peer mode of the row tiles + the group API (bcd_b200_multi.cu)
    "bcd_b200_tile_setup", "bcd_b200_tile_connect", "bcd_b200_tile_run",
    "bcd_b200_group_create", "bcd_b200_group_destroy", "bcd_b200_group_set_pinning", "bcd_b200_group_denoise",
    "bcd_b200_group_get_stats", "bcd_b200_group_get_timings", "bcd_b200_group_tiles", "bcd_b200_denoise_multi",
    "bcd_b200_group_set_tiles_per_gpu", "bcd_b200_group_local_tiles", "bcd_b200_group_create_ranked", "bcd_b200_group_setup",
    "bcd_b200_group_connect", "bcd_b200_group_tile_info", "bcd_b200_group_set_tile_buffers", "bcd_b200_group_upload_tile",
    "bcd_b200_group_run", "bcd_b200_group_download_tile",
]


class Images(C.Structure):
    _fields_ = [("color", C.c_void_p), ("nsamples", C.c_void_p), ("hist", C.c_void_p), ("cov", C.c_void_p),
                ("width", C.c_int32), ("height", C.c_int32), ("hist_depth", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("tau", C.c_float), ("patch_radius", C.c_int32), ("search_radius", C.c_int32), ("min_eig", C.c_float),
                ("random_order", C.c_int32), ("skip_prob", C.c_float), ("nb_scales", C.c_int32), ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("interior", "strong", "active", "full", "fallback", "sum_n_full",
                                         "kernel_launches")]


class Timings(C.Structure):
    _fields_ = [("stage_ms", C.c_float * len(STAGE_NAMES)), ("total_ms", C.c_float)]


PROGRESS_FN = C.CFUNCTYPE(None, C.c_float, C.c_void_p)


class BcdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("bcd_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load libbcd_b200.so (built in-tree by `make` / __graft_entry__.build()). Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: build it with `make` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.bcd_b200_last_error.restype = C.c_char_p
        L.bcd_b200_default_params.argtypes = [C.POINTER(Params)]
        L.bcd_b200_default_params.restype = None
        L.bcd_b200_denoise.argtypes = [C.POINTER(Images), C.POINTER(Params), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.bcd_b200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.bcd_b200_destroy.argtypes = [C.c_void_p]
        L.bcd_b200_destroy.restype = None
        L.bcd_b200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.bcd_b200_set_streams.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bcd_b200_upload.argtypes = [C.c_void_p, C.POINTER(Images)]
        L.bcd_b200_upload_async.argtypes = [C.c_void_p, C.POINTER(Images)]
        L.bcd_b200_host_register.argtypes = [C.c_void_p, C.c_size_t]
        L.bcd_b200_host_unregister.argtypes = [C.c_void_p]
        L.bcd_b200_bind_device.argtypes = [C.c_void_p, C.POINTER(Images)]
        L.bcd_b200_run.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p]
        L.bcd_b200_download.argtypes = [C.c_void_p, C.c_void_p]
        L.bcd_b200_result_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.bcd_b200_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.bcd_b200_get_timings.argtypes = [C.c_void_p, C.POINTER(Timings)]
        L.bcd_b200_mask_words.argtypes = [C.c_int]
        L.bcd_b200_debug_read_selection.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bcd_b200_debug_read_active.argtypes = [C.c_void_p, C.c_void_p]
        L.bcd_b200_debug_read_pair_terms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bcd_b200_spike_removal.argtypes = [C.c_void_p, C.c_float, C.c_int]
        L.bcd_b200_download_inputs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bcd_b200_spike_removal_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                                  C.c_float, C.c_int, C.c_int]
        L.bcd_b200_accumulate_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                  C.c_float, C.c_float]
        L.bcd_b200_group_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int]
        L.bcd_b200_group_destroy.argtypes = [C.c_void_p]
        L.bcd_b200_group_destroy.restype = None
        L.bcd_b200_group_set_pinning.argtypes = [C.c_void_p, C.c_int]
        L.bcd_b200_group_denoise.argtypes = [C.c_void_p, C.POINTER(Images), C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p]
        L.bcd_b200_group_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.bcd_b200_group_get_timings.argtypes = [C.c_void_p, C.POINTER(Timings)]
        L.bcd_b200_group_tiles.argtypes = [C.c_void_p]
        L.bcd_b200_group_local_tiles.argtypes = [C.c_void_p]
        L.bcd_b200_group_set_tiles_per_gpu.argtypes = [C.c_void_p, C.c_int]
        L.bcd_b200_group_create_ranked.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int, C.c_int, C.c_int]
        L.bcd_b200_group_setup.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Params), C.c_void_p, C.POINTER(C.c_int32), C.c_int]
        L.bcd_b200_group_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.bcd_b200_group_tile_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.bcd_b200_group_set_tile_buffers.argtypes = [C.c_void_p, C.c_int, C.POINTER(Images), C.c_void_p]
        L.bcd_b200_group_upload_tile.argtypes = [C.c_void_p, C.c_int, C.POINTER(Images)]
        L.bcd_b200_group_run.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int]
        L.bcd_b200_group_download_tile.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.bcd_b200_denoise_multi.argtypes = [C.POINTER(Images), C.POINTER(Params), C.c_void_p, C.POINTER(C.c_int32), C.c_int,
                                             C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise BcdError(rc, lib().bcd_b200_last_error().decode("utf-8", "replace"))


def make_params(tau=1.0, w=1, b=6, eps=1e-8, random_order=False, m=1.0, scales=1):
    return Params(float(tau), int(w), int(b), float(eps), int(bool(random_order)), float(m), int(scales), 0)


def _host_images(color, nsamples, hist, cov):
    color = np.ascontiguousarray(color, np.float32)
    H, W = color.shape[0], color.shape[1]
    nsamples = np.ascontiguousarray(nsamples, np.float32).reshape(H, W)
    hist = np.ascontiguousarray(hist, np.float32)
    cov = np.ascontiguousarray(cov, np.float32)
    if color.shape != (H, W, 3) or hist.shape[:2] != (H, W) or cov.shape != (H, W, 6):
        raise ValueError("expected color (H,W,3), nsamples (H,W[,1]), hist (H,W,B), cov (H,W,6)")
    keep = (color, nsamples, hist, cov)
    img = Images(color.ctypes.data, nsamples.ctypes.data, hist.ctypes.data, cov.ctypes.data, W, H, hist.shape[2])
    return img, keep


def denoise(color, nsamples, hist, cov, params=None, device=0, **kw):
    """One-shot host-in / host-out call: bcd_b200_denoise (what bcd::Denoiser::denoise() forwards to)."""
    p = params if params is not None else make_params(**kw)
    img, keep = _host_images(color, nsamples, hist, cov)
    out = np.empty((img.height, img.width, 3), np.float32)
    _check(lib().bcd_b200_denoise(C.byref(img), C.byref(p), out.ctypes.data, device, None, None))
    return out


def denoise_multi(color, nsamples, hist, cov, devices, params=None, **kw):
    """One-shot multi-GPU call: bcd_b200_denoise_multi (row tiles over peer memory; devices may repeat a GPU)."""
    p = params if params is not None else make_params(**kw)
    img, keep = _host_images(color, nsamples, hist, cov)
    out = np.empty((img.height, img.width, 3), np.float32)
    ids = (C.c_int32 * len(devices))(*devices)
    _check(lib().bcd_b200_denoise_multi(C.byref(img), C.byref(p), out.ctypes.data, ids, len(devices), None, None))
    return out


class Group:
    """N GPUs behind one host-buffer call (bcd_b200_group_*): what bcd::Denoiser / bcd::MultiscaleDenoiser hold."""

    def __init__(self, devices, pin=None, tiles_per_gpu=None, rank=0, world=1):
        self._g = C.c_void_p()
        ids = (C.c_int32 * len(devices))(*devices)
        _check(lib().bcd_b200_group_create_ranked(C.byref(self._g), ids, len(devices), rank, world))
        if pin is not None:
            _check(lib().bcd_b200_group_set_pinning(self._g, int(bool(pin))))
        if tiles_per_gpu is not None:
            _check(lib().bcd_b200_group_set_tiles_per_gpu(self._g, int(tiles_per_gpu)))

    # ---- ranked use (one process per GPU): see include/bcd_b200.h
    TOKEN_BYTES = 96

    def setup(self, W, H, B, params, capacity=64):
        """-> [(tile index, token bytes)] of the local tiles."""
        toks = C.create_string_buffer(self.TOKEN_BYTES * capacity)
        idx = (C.c_int32 * capacity)()
        _check(lib().bcd_b200_group_setup(self._g, W, H, B, C.byref(params), toks, idx, capacity))
        n = self.local_tiles()
        return [(int(idx[i]), toks.raw[i * self.TOKEN_BYTES:(i + 1) * self.TOKEN_BYTES]) for i in range(n)]

    def connect(self, tokens_by_tile=None):
        if tokens_by_tile is None:
            _check(lib().bcd_b200_group_connect(self._g, None, 0))
        else:
            buf = C.create_string_buffer(b"".join(tokens_by_tile), self.TOKEN_BYTES * len(tokens_by_tile))
            _check(lib().bcd_b200_group_connect(self._g, buf, len(tokens_by_tile)))

    def local_tiles(self):
        return int(lib().bcd_b200_group_local_tiles(self._g))

    def tile_info(self, local_tile):
        """-> dict(image_width, image_height, own_row0, own_row1, sub_row0, sub_row1, tile_index, tile_count)."""
        v = (C.c_int32 * 8)()
        _check(lib().bcd_b200_group_tile_info(self._g, local_tile, v))
        return dict(zip(("image_width", "image_height", "own_row0", "own_row1", "sub_row0", "sub_row1", "tile_index", "tile_count"),
                        [int(x) for x in v]))

    def set_tile_buffers(self, local_tile, color_ptr, ns_ptr, hist_ptr, cov_ptr, W, rows, B, out_ptr):
        img = Images(color_ptr, ns_ptr, hist_ptr, cov_ptr, W, rows, B)
        _check(lib().bcd_b200_group_set_tile_buffers(self._g, local_tile, C.byref(img), C.c_void_p(out_ptr)))

    def upload_tile(self, local_tile, color_ptr, ns_ptr, hist_ptr, cov_ptr, W, rows, B):
        img = Images(color_ptr, ns_ptr, hist_ptr, cov_ptr, W, rows, B)
        _check(lib().bcd_b200_group_upload_tile(self._g, local_tile, C.byref(img)))

    def run(self, params=None, from_host=False):
        _check(lib().bcd_b200_group_run(self._g, C.byref(params) if params is not None else None, int(bool(from_host))))

    def download_tile(self, local_tile, out_ptr):
        _check(lib().bcd_b200_group_download_tile(self._g, local_tile, C.c_void_p(out_ptr)))

    def close(self):
        if self._g:
            lib().bcd_b200_group_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def denoise(self, color, nsamples, hist, cov, params=None, out=None, **kw):
        p = params if params is not None else make_params(**kw)
        img, keep = _host_images(color, nsamples, hist, cov)
        if out is None:
            out = np.empty((img.height, img.width, 3), np.float32)
        _check(lib().bcd_b200_group_denoise(self._g, C.byref(img), C.byref(p), out.ctypes.data, None, None))
        return out

    def denoise_ptrs(self, color_ptr, ns_ptr, hist_ptr, cov_ptr, W, H, B, out_ptr, params):
        img = Images(color_ptr, ns_ptr, hist_ptr, cov_ptr, W, H, B)
        _check(lib().bcd_b200_group_denoise(self._g, C.byref(img), C.byref(params), C.c_void_p(out_ptr), None, None))

    def tiles(self):
        return int(lib().bcd_b200_group_tiles(self._g))

    def stats(self):
        s = Stats()
        _check(lib().bcd_b200_group_get_stats(self._g, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in Stats._fields_}

    def timings(self):
        t = Timings()
        _check(lib().bcd_b200_group_get_timings(self._g, C.byref(t)))
        d = {n: float(t.stage_ms[i]) for i, n in enumerate(STAGE_NAMES)}
        d["total"] = float(t.total_ms)
        return d


class Handle:
    """Persistent device buffers + stream (bcd_b200_create ... bcd_b200_destroy)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().bcd_b200_create(C.byref(self._h), device))
        self.device = device
        self.shape = None
        self._keep = None

    def close(self):
        if self._h:
            lib().bcd_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        _check(lib().bcd_b200_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_streams(self, main_ptr, aux_ptr):
        _check(lib().bcd_b200_set_streams(self._h, C.c_void_p(main_ptr), C.c_void_p(aux_ptr)))

    def upload(self, color, nsamples, hist, cov):
        img, keep = _host_images(color, nsamples, hist, cov)
        _check(lib().bcd_b200_upload(self._h, C.byref(img)))
        self.shape = (img.height, img.width, img.hist_depth)

    def upload_ptrs(self, color_ptr, ns_ptr, hist_ptr, cov_ptr, W, H, B, streamed=False):
        """Host pointers (e.g. pinned torch tensors' data_ptr()). streamed=True: bcd_b200_upload_async — the copies
        overlap the next run(); the caller keeps the host buffers alive and unmodified until that run returned."""
        img = Images(color_ptr, ns_ptr, hist_ptr, cov_ptr, W, H, B)
        _check((lib().bcd_b200_upload_async if streamed else lib().bcd_b200_upload)(self._h, C.byref(img)))
        self.shape = (H, W, B)

    def upload_async(self, color, nsamples, hist, cov):
        img, keep = _host_images(color, nsamples, hist, cov)
        self._pending_host = keep  # the arrays must outlive the streamed copy
        _check(lib().bcd_b200_upload_async(self._h, C.byref(img)))
        self.shape = (img.height, img.width, img.hist_depth)

    def bind_device(self, color_ptr, ns_ptr, hist_ptr, cov_ptr, W, H, B):
        """Borrow DEVICE pointers (e.g. torch CUDA tensors' data_ptr()) as the level-0 inputs."""
        img = Images(color_ptr, ns_ptr, hist_ptr, cov_ptr, W, H, B)
        _check(lib().bcd_b200_bind_device(self._h, C.byref(img)))
        self.shape = (H, W, B)

    def run(self, params=None, progress=None, **kw):
        p = params if params is not None else make_params(**kw)
        cb = PROGRESS_FN(lambda v, _u: progress(v)) if progress else None
        _check(lib().bcd_b200_run(self._h, C.byref(p), cb, None))
        self._last_b = p.search_radius

    def download(self, out=None):
        H, W, _ = self.shape
        if out is None:
            out = np.empty((H, W, 3), np.float32)
        _check(lib().bcd_b200_download(self._h, out.ctypes.data))
        return out

    def download_ptr(self, host_ptr):
        _check(lib().bcd_b200_download(self._h, C.c_void_p(host_ptr)))

    def result_device_ptr(self):
        p = C.c_void_p()
        _check(lib().bcd_b200_result_device_ptr(self._h, C.byref(p)))
        return p.value

    def accumulate_samples(self, samples, weights=None, nb_bins=20, gamma=2.2, max_value=2.5, device_ptrs=None):
        """bcd_b200_accumulate_samples: (H, W, spp, 3) host samples (or device_ptrs=(samples_ptr, weights_ptr|0, H, W, spp))
        -> the handle's resident inputs (SamplesAccumulator on the GPU)."""
        if device_ptrs is not None:
            sp, wp, H, W, spp = device_ptrs
            _check(lib().bcd_b200_accumulate_samples(self._h, C.c_void_p(sp), C.c_void_p(wp) if wp else None, 1, W, H, spp, nb_bins, gamma, max_value))
        else:
            samples = np.ascontiguousarray(samples, np.float32)
            H, W, spp, _ = samples.shape
            wts = np.ascontiguousarray(weights, np.float32) if weights is not None else None
            _check(lib().bcd_b200_accumulate_samples(self._h, samples.ctypes.data, wts.ctypes.data if wts is not None else None, 0,
                                                     W, H, spp, nb_bins, gamma, max_value))
        self.shape = (H, W, 3 * nb_bins)

    def spike_removal(self, factor=2.0, int_abs=True):
        """SpikeRemovalFilter on the resident (uploaded) inputs; they stay resident for the next run."""
        _check(lib().bcd_b200_spike_removal(self._h, float(factor), int(bool(int_abs))))

    def download_inputs(self):
        H, W, B = self.shape
        out = dict(color=np.empty((H, W, 3), np.float32), nsamples=np.empty((H, W, 1), np.float32),
                   hist=np.empty((H, W, B), np.float32), cov=np.empty((H, W, 6), np.float32))
        _check(lib().bcd_b200_download_inputs(self._h, out["color"].ctypes.data, out["nsamples"].ctypes.data,
                                              out["hist"].ctypes.data, out["cov"].ctypes.data))
        return out

    def stats(self):
        s = Stats()
        _check(lib().bcd_b200_get_stats(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in Stats._fields_}

    def timings(self):
        t = Timings()
        _check(lib().bcd_b200_get_timings(self._h, C.byref(t)))
        d = {n: float(t.stage_ms[i]) for i, n in enumerate(STAGE_NAMES)}
        d["total"] = float(t.total_ms)
        return d

    def read_selection(self):
        H, W, _ = self.shape
        words = lib().bcd_b200_mask_words(self._last_b)
        masks = np.zeros((H, W, words), np.uint32)
        counts = np.zeros((H, W), np.int32)
        _check(lib().bcd_b200_debug_read_selection(self._h, masks.ctypes.data, counts.ctypes.data))
        return masks, counts

    def read_active(self):
        H, W, _ = self.shape
        active = np.zeros((H, W), np.uint8)
        _check(lib().bcd_b200_debug_read_active(self._h, active.ctypes.data))
        return active

    def read_pair_terms(self, b=6):
        """After a single-scale run: (sums (J,H,W) float32, counts (J,H,W) uint8) of the selection stage."""
        H, W, _ = self.shape
        J = ((2 * b + 1) ** 2 + 1) // 2
        sums, counts = np.empty((J, H, W), np.float32), np.empty((J, H, W), np.uint8)
        _check(lib().bcd_b200_debug_read_pair_terms(self._h, sums.ctypes.data, counts.ctypes.data))
        return sums, counts

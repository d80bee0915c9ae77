#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native BCD hot path (contract: see the task statement, section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--m 1|0] [--size WxH] [--scales S]

N = 1 workload (BASELINE.json `metric` / configs[2]): synthetic G(1920,1080,spp=64,seed=20170627) colour + histogram +
covariance buffers, MultiscaleDenoiser s=3, defaults b=6, w=1, tau=1, m=1 (marking strategy, raster order).
A "step" is one complete denoise of that frame.

  value        : Mpixels/s with the inputs already resident in HBM (bind_device + bcd_b200_run; CUDA events on the
                 launching stream around exactly K steps, nothing else inside the timed region).
  e2e          : the same metric through ONE C-ABI call per frame with host buffers in and out — bcd_b200_group_denoise,
                 the call bcd::Denoiser::denoise() forwards to: upload (pinned host -> device, streamed under the level-0
                 selection) + run + download, all inside the timed region. `ms_per_step_handle_api` is the same frame
                 through bcd_b200_upload_async + bcd_b200_run + bcd_b200_download.
  e2e_plugin   : the same through the C++ drop-in itself — one bcd::MultiscaleDenoiser object, DeepImages in pageable
                 std::vector storage, denoise() per frame; first call and steady state separately.
  roofline     : dominant kernel (k_pair_terms_sparse, stage "pair") — algorithmic bytes / CUDA-event time vs the
                 measured HBM peak (MEASURED_PEAKS.json); the binding roof of this kernel is FP32 issue / XU, not HBM
                 (SURVEY 8d), so `roofline_fp32` reports the bin-term rate against the FP32 pipe as well.
  nonuniform_n : the same workload with 48 samples per pixel (general-n path of the pair-term kernel).
  cpu_baseline : the UNMODIFIED reference CPU path (oracle/_ref, all host threads) on a bounded crop of the same frame.
  parity       : the CUDA path on a window of this very frame against the reference on the same window.

`--impl reference` times only that reference arm (rank 0), one bounded crop per step; nothing of the product is mapped
into that process.
N > 1 (launched by torchrun, one process per GPU): the frame is cut into row tiles held by a ranked bcd_b200_group; the
tiles exchange their strips through peer memory (CUDA IPC over NVLink) — bcd_b200/tiles.py: bench_tiled, DESIGN.md 7.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# every stream of a handle gets its own hardware queue (must be set before the CUDA context exists)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")  # a lazy kernel load waits for an idle device (see bcd_b200_tile_setup)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpixels/s denoised"
UNIT = "Mpix/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--m", type=float, default=1.0, help="markedPixelsSkippingProbability (1 = default marking strategy)")
    ap.add_argument("--size", default=None, help="WxH (default 1920x1080 for every N; the 4K scaling series uses --size 4096x4096)")
    ap.add_argument("--scales", type=int, default=3)
    ap.add_argument("--w", type=int, default=1)
    ap.add_argument("--b", type=int, default=6)
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--tiles-per-gpu", type=int, default=0, help="N > 1: row tiles per GPU, dealt out cyclically (0 = chosen from the frame height)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-plugin", action="store_true", help="skip the e2e_plugin leg (bcd::MultiscaleDenoiser::denoise() on pageable buffers)")
    ap.add_argument("--alt-spp", type=int, default=48, help="second resident measurement with this sample count (not a power of two); 0 = skip")
    ap.add_argument("--parity-crop", type=int, default=256, help="window of the single-thread reference run the frame's parity is checked on (m > 0)")
    ap.add_argument("--cpu-crop", type=int, default=768, help="side of the square crop the CPU reference is timed on (768: about 5 s of CPU work on 16 threads, 25 s on 8)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """SM clock + throttle reasons of one GPU, sampled DURING the timed region: NVML (pynvml) polled every 2 ms from a
    thread — the timed region of the default run is only ~0.2 s, too short for `nvidia-smi -lms` to deliver a sample —
    with the nvidia-smi loop of the profiling recipe as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    _nvml = None

    def __init__(self, index=0, uuid=None):
        self.rows, self.proc, self.index, self.uuid = [], None, index, uuid
        self.samples, self.reasons, self.max_mhz, self.stop_flag, self.t, self.handle = [], set(), None, False, None, None

    @classmethod
    def _init_nvml(cls):
        if cls._nvml is None:
            try:
                import pynvml
                pynvml.nvmlInit()
                cls._nvml = pynvml
            except Exception:
                cls._nvml = False
        return cls._nvml

    def _open_nvml(self):
        nv = self._init_nvml()
        if not nv:
            return False
        try:
            h = None
            if self.uuid:
                for u in (self.uuid, "GPU-" + self.uuid):
                    try:
                        h = nv.nvmlDeviceGetHandleByUUID(u.encode() if isinstance(u, str) else u)
                        break
                    except Exception:
                        h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                idx = self.index
                if vis:
                    ent = [v.strip() for v in vis.split(",") if v.strip()]
                    if self.index < len(ent) and ent[self.index].isdigit():
                        idx = int(ent[self.index])
                h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.handle = h
            return True
        except Exception:
            return False

    def _poll_nvml(self):
        nv, h = self._nvml, self.handle
        bits = []
        for name, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                           ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
            v = getattr(nv, attr, None)
            if v is None:
                v = getattr(nv, attr.replace("ClocksEventReason", "ClocksThrottleReason"), None)
            if v is not None:
                bits.append((name, v))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                if get_reasons:
                    r = get_reasons(h)
                    for name, bit in bits:
                        if r & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self._open_nvml():
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.handle is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            sm = self.samples
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml, 2 ms period, timed region only"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 20"}


def gpu_uuid(local_rank):
    try:
        import torch
        return str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------ CPU reference
def host_threads():
    """Host threads the CPU reference may use: the affinity mask of this process (torchrun exports OMP_NUM_THREADS=1
    for nproc > 1, which an explicit m_nbOfCores overrides — reference Denoiser.cpp:112-115)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def centre_crop(inputs, crop):
    H, W, _ = inputs["color"].shape
    c = min(crop, H, W)
    y0, x0 = (H - c) // 2, (W - c) // 2  # centred: straddles the albedo-texture boundary at x = W/2
    return {k: np.ascontiguousarray(v[y0:y0 + c, x0:x0 + c]) for k, v in inputs.items()}, c


def cpu_reference_sample(inputs, crop, scales, m, w, b, cores=None):
    """Times the unmodified reference (oracle/_ref) on a crop x crop window of the frame, all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracles as O
    sub, c = centre_crop(inputs, crop)
    kind = "reference" if O.have_ref() else "port"
    cores = host_threads() if cores is None else cores
    t = time.perf_counter()
    if kind == "reference":
        out = O.ref_denoise(**sub, cores=cores, m=m, scales=scales, w=w, b=b, random_order=False)
        ncores = cores
    else:
        out = O.port_denoise(**sub, m=m, scales=scales, w=w, b=b)
        ncores = 1
    dt = time.perf_counter() - t
    return {"value": c * c / dt / 1e6, "unit": UNIT, "cores": int(ncores), "kind": kind, "seconds": dt, "crop": c, "out": out,
            "sample": "%dx%d centre crop of the same synthetic frame, s=%d m=%g b=%d w=%d, raster order, %s" % (
                c, c, scales, m, b, w, "OpenMP, %d threads (m_nbOfCores set explicitly)" % ncores if kind == "reference" else "scalar C port")}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------------------------- main
def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = a.gpus
    size = a.size or "1920x1080"
    W, H = [int(v) for v in size.lower().split("x")]
    workload = "G(%d,%d,spp=%d,seed=20170627) synthetic colour+histogram+covariance, MultiscaleDenoiser s=%d, b=%d w=%d tau=1 m=%g" % (
        W, H, a.spp, a.scales, a.b, a.w, a.m)
    config = {"workload": workload, "width": W, "height": H, "scales": a.scales, "search_radius": a.b, "patch_radius": a.w,
              "skip_prob": a.m, "order": "raster (the reference's deterministic order)",
              "l2_policy": "inputs (%.0f MB) exceed the 126 MB L2; no explicit flush" % (W * H * 280 / 1e6),
              "parallelism": "1 GPU" if n_gpus == 1 else "row tiles x%d (one rank per GPU)" % n_gpus}

    # torchrun exports OMP_NUM_THREADS=1 for nproc > 1: the synthetic-input generator (OpenMP) would crawl. Give every
    # rank its share of the host threads before any OpenMP runtime is loaded.
    if world > 1 or os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(max(1, host_threads() // (1 if a.impl == "reference" else world)))

    from bcd_b200 import host

    # ---------------------------------------------------------------------------------------------- reference arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        # nothing of the product path is mapped here: the generator comes from libbcdsynth.so (no CUDA library behind
        # it), the timed call is bcdref_denoise of oracle/_ref/libbcd_ref.so (the reference compiled in place)
        inp = host.synth(W, H, spp=a.spp)
        crop = min(a.cpu_crop, W, H)
        budget_s = 200.0  # the whole --steps K --warmup W run must end within a few minutes
        times, info = [], None
        for i in range(a.warmup + a.steps):
            info = cpu_reference_sample(inp, crop, a.scales, a.m, a.w, a.b)
            if i == 0 and info["seconds"] * (a.warmup + a.steps) > budget_s:
                # first sample too slow for K + W repeats on this host: shrink the window (Mpix/s is size-independent
                # to first order: the clipped-window border is a few percent of these crops)
                scale = (budget_s / (info["seconds"] * (a.warmup + a.steps))) ** 0.5
                crop = max(192, int(crop * scale) // 64 * 64)
            if i >= a.warmup:
                times.append((info["crop"], info["seconds"]))
        pix = sum(c * c for c, _ in times)
        secs = sum(t for _, t in times)
        val = pix / secs / 1e6
        info["value"] = val
        config["workload"] = workload + " — reference arm: each step denoises a %dx%d centre crop of that frame" % (times[-1][0], times[-1][0])
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": 1e3 * secs / len(times), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # --------------------------------------------------------------------------------------------------- our arm
    import torch
    from bcd_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # libraries (NCCL's version banner) may write to fd 1; rank 0's stdout must carry exactly ONE JSON line, so fd 1
        # is pointed at stderr until that line is printed (tiles.bench_tiled restores it)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if world > 1:
        from bcd_b200 import tiles
        return tiles.bench_tiled(a, rank, world, local_rank, W, H, config, workload, dist, saved_stdout)

    # ---- inputs: generated straight into pinned host memory
    B = 60
    pin = dict(color=torch.empty((H, W, 3), dtype=torch.float32).pin_memory(),
               nsamples=torch.empty((H, W, 1), dtype=torch.float32).pin_memory(),
               hist=torch.empty((H, W, B), dtype=torch.float32).pin_memory(),
               cov=torch.empty((H, W, 6), dtype=torch.float32).pin_memory())
    t0 = time.time()
    host.synth(W, H, spp=a.spp, out={k: v.numpy() for k, v in pin.items()})
    gen_s = time.time() - t0
    out_pin = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    dev = {k: v.cuda(non_blocking=True) for k, v in pin.items()}
    torch.cuda.synchronize()

    stream = torch.cuda.Stream()  # explicit stream shared by the torch events and the handle
    torch.cuda.set_stream(stream)
    h = capi.Handle(local_rank)
    h.set_stream(stream.cuda_stream)
    params = capi.make_params(w=a.w, b=a.b, m=a.m, scales=a.scales)
    h2d = sum(v.numel() * 4 for v in pin.values())
    d2h = out_pin.numel() * 4

    def step_resident():
        h.run(params)

    def step_e2e():
        # what bcd_b200_denoise (the call behind bcd::Denoiser::denoise) does per frame: streamed upload — the histogram
        # rows arrive under the level-0 pair-term kernel — run, download
        h.upload_ptrs(pin["color"].data_ptr(), pin["nsamples"].data_ptr(), pin["hist"].data_ptr(), pin["cov"].data_ptr(), W, H, B,
                      streamed=True)
        h.run(params)
        h.download_ptr(out_pin.data_ptr())

    def step_e2e_serial():
        h.upload_ptrs(pin["color"].data_ptr(), pin["nsamples"].data_ptr(), pin["hist"].data_ptr(), pin["cov"].data_ptr(), W, H, B)
        h.run(params)
        h.download_ptr(out_pin.data_ptr())

    def timed(fn, bind):
        if bind:
            h.bind_device(dev["color"].data_ptr(), dev["nsamples"].data_ptr(), dev["hist"].data_ptr(), dev["cov"].data_ptr(), W, H, B)
        for _ in range(a.warmup):
            fn()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank, gpu_uuid(local_rank))
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(a.steps):
            fn()  # nothing but the step itself inside the timed region
        e1.record(stream)
        torch.cuda.synchronize()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        # per-stage device times and the launch count: sampled on extra, untimed steps (resolving the stage events costs
        # host time that does not belong to a step)
        stage_acc, launches, probes = {}, 0, 3
        for _ in range(probes):
            fn()
            for k, v in h.timings().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v * a.steps / probes
            launches = h.stats()["kernel_launches"] * a.steps
        return ms, stage_acc, launches, clocks

    ms_res, stages, launches, clocks = timed(step_resident, True)
    stats = h.stats()
    ms_e2e_handle, stages_e2e, _, clocks_e2e = timed(step_e2e, False)
    ms_e2e_serial, _, _, _ = timed(step_e2e_serial, False)
    handle_img = out_pin.numpy().copy()

    # ---- e2e: ONE C-ABI call per frame with host buffers in and out — bcd_b200_group_denoise, what bcd::Denoiser::denoise()
    #      forwards to. On one GPU the group cuts the frame into a few row tiles whose uploads are chained: the first tile
    #      is selected, marked and filtered while the rows of the next ones are still crossing PCIe. A frame spans the
    #      streams of several tile handles, so the K steps are bracketed by two events recorded on an idle device.
    grp = capi.Group([local_rank])
    ptrs = (pin["color"].data_ptr(), pin["nsamples"].data_ptr(), pin["hist"].data_ptr(), pin["cov"].data_ptr())

    def step_group():
        grp.denoise_ptrs(ptrs[0], ptrs[1], ptrs[2], ptrs[3], W, H, B, out_pin.data_ptr(), params)

    for _ in range(max(a.warmup, 1)):
        step_group()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        step_group()
    torch.cuda.synchronize()
    e1.record(stream)
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    e2e_tiles = grp.tiles()
    grp.close()
    # bytes that really cross PCIe per frame: every tile uploads its band plus its halo rows
    from bcd_b200 import tiles as tiles_mod
    if e2e_tiles > 1:
        rows_up = sum(t.sub_row1 - t.sub_row0 for t in (tiles_mod.plan(W, H, i, e2e_tiles, params) for i in range(e2e_tiles)))
    else:
        rows_up = H
    h2d_e2e = rows_up * W * (3 + 1 + B + 6) * 4
    result_ok = bool(np.isfinite(out_pin.numpy()[2:-2, 2:-2]).all())
    e2e_vs_handle = float(np.abs(out_pin.numpy()[2:-2, 2:-2] - handle_img[2:-2, 2:-2]).max() / np.abs(handle_img[2:-2, 2:-2]).max())

    # ---- the plugin call itself: ONE bcd::MultiscaleDenoiser object and pageable std::vector DeepImages, denoise() per
    #      frame (what a host renderer / the reference's GUI does, src/gui/GuiWindow.cpp:345-376). First call = device
    #      session created, buffers allocated, image storage page-locked; steady state = every call after it.
    e2e_plugin = None
    if not a.no_plugin:
        sess = host.CxxSession(scales=a.scales, devices=[local_rank])
        sess.load(**{k: v.numpy() for k, v in pin.items()})
        t0 = time.perf_counter()
        ok_first = sess.denoise(m=a.m, w=a.w, b=a.b)
        first_ms = (time.perf_counter() - t0) * 1e3
        for _ in range(max(0, a.warmup - 1)):
            sess.denoise(m=a.m, w=a.w, b=a.b)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            ok_first = sess.denoise(m=a.m, w=a.w, b=a.b) and ok_first
        steady_ms = (time.perf_counter() - t0) * 1e3 / a.steps
        plug = sess.fetch()
        sess.close()
        ref_img = out_pin.numpy()
        e2e_plugin = {"value": W * H / (steady_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": steady_ms, "first_call_ms": first_ms,
                      "ok": bool(ok_first), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                      "rel_linf_vs_handle_api": float(np.abs(plug[2:-2, 2:-2] - ref_img[2:-2, 2:-2]).max() / np.abs(ref_img[2:-2, 2:-2]).max()),
                      "api": "bcd::MultiscaleDenoiser::denoise() on one persistent object, DeepImage storage in pageable std::vector "
                             "(page-locked by the object on first sight), host wall clock around denoise()"}

    # ---- the path a renderer with a non-power-of-two / adaptive sample count takes (K1a's uniform-2^k shortcut is off)
    nonuniform = None
    if a.alt_spp > 0 and a.alt_spp != a.spp:
        alt = host.synth(W, H, spp=a.alt_spp)
        h.upload(**alt)
        del alt
        for _ in range(a.warmup):
            h.run(params)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        acc_pair = 0.0
        e0.record(stream)
        for _ in range(a.steps):
            h.run(params)
            acc_pair += h.timings()["pair"]
        e1.record(stream)
        torch.cuda.synchronize()
        ms_alt = e0.elapsed_time(e1) / a.steps
        nonuniform = {"spp": a.alt_spp, "value": W * H / (ms_alt * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_alt,
                      "pair_ms_per_step": acc_pair / a.steps, "work_counters": h.stats(),
                      "what": "same workload with %d samples per pixel: n is not a power of two, so the pair-term kernel evaluates the "
                              "reference's full term (13 packed operations per visited bin instead of 9)" % a.alt_spp}

    value = W * H * a.steps / (ms_res * 1e-3) / 1e6
    e2e_val = W * H * a.steps / (ms_e2e * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (k_pair_terms_tiled: every level launches it once)
    pk, pk_src = peaks()
    npix_all = sum((W >> k) * (H >> k) for k in range(a.scales))
    pair_ms = stages["pair"] / a.steps                      # all levels, per step
    J = ((2 * a.b + 1) ** 2 + 1) // 2
    alg_bytes = npix_all * (B + 1) * 4                      # compulsory: histogram + nSamples read once (244 B/px)
    achieved = alg_bytes / (pair_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "pair_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            # the ncu capture is the level-0 launch; DRAM traffic of this kernel is proportional to the pixel count, so the
            # launch set of one step (all levels) is the capture scaled by npix_all / npix_level0 — same units as `achieved`
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            if traffic is not None and (W, H) == (1920, 1080):
                traffic = traffic * npix_all / float(W * H)
            else:
                traffic = None  # captured at 1080p only
        except Exception:
            traffic = None
    roofline = {"kernel": "k_pair_terms_sparse (stage 'pair', %d launches/step)" % a.scales, "bound": "hbm",
                "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                "peak_source": pk_src + " (burst copy figure; kernel timed alone by CUDA events)", "traffic": traffic,
                "ms_per_launch_set": pair_ms,
                "traffic_source": "profiles/pair_kernel_traffic.json (ncu --set full, level-0 launch) scaled to the launch set of a step",
                "note": "FP32-issue-bound kernel: HBM fraction is low by construction (SURVEY 8d); see roofline_fp32"}
    terms = npix_all * J * B                                # bin terms actually evaluated per step
    fp32_peak = 148 * 128 * 2 * (clocks.get("sm_mhz") or 1900.0) * 1e6 / 1e12  # FMA TFLOP/s at the sampled clock
    roofline_fp32 = {"bin_terms_per_s": terms / (pair_ms * 1e-3), "flop_equiv_per_term": 10,
                     "achieved_tflops_equiv": terms * 10 / (pair_ms * 1e-3) / 1e12, "fp32_peak_tflops": fp32_peak,
                     "frac": terms * 10 / (pair_ms * 1e-3) / 1e12 / fp32_peak}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_res / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_e2e, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / a.steps,
                    "api": "bcd_b200_group_denoise(host buffers in, host buffer out; pinned): one C-ABI call per frame, the call "
                           "bcd::Denoiser::denoise() forwards to; %d row tiles with chained uploads on the one GPU" % e2e_tiles,
                    "tiles": e2e_tiles, "rel_linf_vs_whole_frame": e2e_vs_handle,
                    "h2d_note": "counted from the rows every tile uploads (band + halo): %d rows for a %d-row frame" % (rows_up, H),
                    "ms_per_step_handle_api": ms_e2e_handle / a.steps,
                    "handle_api": "bcd_b200_upload_async + bcd_b200_run + bcd_b200_download on the whole frame (copies overlap the level-0 selection only)",
                    "ms_per_step_serial_upload": ms_e2e_serial / a.steps},
            "gpu_launches": launches, "roofline": roofline, "roofline_fp32": roofline_fp32,
            "stage_ms_per_step": {k: v / a.steps for k, v in stages.items()},
            "stage_ms_per_step_e2e": {k: v / a.steps for k, v in stages_e2e.items()},
            "work_counters": stats, "input_generation_s": gen_s, "output_finite": result_ok}
    if e2e_plugin is not None:
        line["e2e_plugin"] = e2e_plugin
    if nonuniform is not None:
        line["nonuniform_n"] = nonuniform
    if not a.no_cpu_baseline:
        frame = {k: v.numpy() for k, v in pin.items()}
        cb = cpu_reference_sample(frame, a.cpu_crop, a.scales, a.m, a.w, a.b)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        # parity of this very frame: the CUDA path against the reference on the same window. With the marking strategy
        # (m = 1) the reference is only deterministic on ONE thread in raster order (SURVEY finding 2), so that case
        # pays a second, smaller single-thread run; for m = 0 the timed multi-thread run above is the checker.
        if a.m > 0:
            chk = cpu_reference_sample(frame, a.parity_crop, a.scales, a.m, a.w, a.b, cores=1)
        else:
            chk = cb
        sub, c = centre_crop(frame, chk["crop"])
        h.upload(**sub)
        h.run(params)
        ours = h.download(np.empty((c, c, 3), np.float32))
        line["parity"] = {"rel_linf": float(np.abs(ours - chk["out"]).max() / np.abs(chk["out"]).max()), "tolerance": 1e-4,
                          "sample": chk["sample"], "checker": chk["kind"],
                          "what": "bcd_b200_run on that window vs the CPU reference on the same window (max|a-b| / max|ref|)"}
    print(json.dumps(line))
    h.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())

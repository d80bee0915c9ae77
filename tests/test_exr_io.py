"""SURVEY §8 row f2/f4 — EXR entry points, `.bcd.json` presets and the command-line front ends (CPU tests).

The from-scratch scan-line codec (bcd_b200/io/exr_codec.cpp behind bcd::ImageIO) is pinned two ways:
  (1) committed files written by the reference's own code (tests/golden/io, generator tests/golden/make_golden_io.py);
  (2) live, where oracle/_ref/libexr_ref.so exists (the reference's src/io/exr/io_exr.cpp over the vendored OpenEXR 2.2,
      compiled in place by oracle/Makefile): every compression and pixel type the reader accepts, both directions.
Values must match exactly (HALF files: exactly the binary16 rounding of the input).
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from bcd_b200 import host

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden", "io")
EXR_REF = os.path.join(ROOT, "oracle", "_ref", "libexr_ref.so")
BCD_CLI = os.path.join(ROOT, "bin", "bcd_cli")
RAW2BCD = os.path.join(ROOT, "bin", "raw2bcd")

needs_ref = pytest.mark.skipif(not os.path.exists(EXR_REF), reason="oracle/_ref/libexr_ref.so not built (needs /root/reference)")


def ref():
    R = C.CDLL(EXR_REF)
    R.exrref_read_rgb.restype = C.POINTER(C.c_float)
    R.exrref_read_multi.restype = C.POINTER(C.c_float)
    R.exrref_free_rgb.argtypes = [C.c_void_p]
    R.exrref_free_multi.argtypes = [C.c_void_p]
    return R


def ref_read_rgb(path):
    R = ref()
    w, h = C.c_int(), C.c_int()
    p = R.exrref_read_rgb(os.fsencode(path), C.byref(w), C.byref(h))
    a = np.ctypeslib.as_array(p, shape=(3, h.value, w.value)).copy()
    R.exrref_free_rgb(p)
    return a


def ref_read_multi(path):
    R = ref()
    w, h, n = C.c_int(), C.c_int(), C.c_int()
    p = R.exrref_read_multi(os.fsencode(path), C.byref(w), C.byref(h), C.byref(n))
    a = np.ctypeslib.as_array(p, shape=(n.value, h.value, w.value)).copy()
    R.exrref_free_multi(p)
    return a


def same_bits(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def as_half(a):
    with np.errstate(over="ignore"):
        return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


# ------------------------------------------------------------------------------------------ committed fixtures
def test_reads_reference_written_files():
    v = np.load(os.path.join(GOLD, "ref_io_values.npz"))
    color = host.exr_load(os.path.join(GOLD, "ref_color_piz.exr"), False)       # HALF RGBA, PIZ
    assert same_bits(color.transpose(2, 0, 1), as_half(v["color"]))
    bins = host.exr_load(os.path.join(GOLD, "ref_bins_zip.exr"), True)          # FLOAT Bin_%04d, ZIP
    assert same_bits(bins.transpose(2, 0, 1), v["bins"])
    smooth = host.exr_load(os.path.join(GOLD, "ref_smooth_piz.exr"), False)     # FLOAT RGB, PIZ
    assert same_bits(smooth.transpose(2, 0, 1), v["smooth"])


def test_grey_colour_file_collapses_to_one_channel(tmp_path):
    """ImageIO.cpp:41-52: R == G == B everywhere -> depth 1; writeEXR expands one channel to grey RGB (:86-103)."""
    g = np.random.default_rng(2).random((9, 14, 1), dtype=np.float32)
    p = str(tmp_path / "grey.exr")
    host.exr_write(p, g, False)
    back = host.exr_load(p, False)
    assert back.shape == (9, 14, 1) and same_bits(back, as_half(g))


def test_error_paths_return_false_with_a_message(tmp_path):
    with pytest.raises(host.ExrError, match="cannot read"):
        host.exr_load(str(tmp_path / "missing.exr"), False)
    junk = tmp_path / "junk.exr"
    junk.write_bytes(b"not an exr file at all")
    with pytest.raises(host.ExrError, match="not an OpenEXR file"):
        host.exr_load(str(junk), True)
    good = open(os.path.join(GOLD, "ref_bins_zip.exr"), "rb").read()
    for cut in (30, 400, len(good) // 2, len(good) - 7):
        t = tmp_path / ("cut%d.exr" % cut)
        t.write_bytes(good[:cut])
        with pytest.raises(host.ExrError):
            host.exr_load(str(t), True)
    # a colour file has no Bin_0000 channel
    with pytest.raises(host.ExrError, match="Bin_0000"):
        host.exr_load(os.path.join(GOLD, "ref_color_piz.exr"), True)
    # corrupt payload: flip bytes inside the first compressed chunk
    bad = bytearray(good)
    for i in range(len(bad) - 600, len(bad) - 560):
        bad[i] ^= 0x5a
    t = tmp_path / "corrupt.exr"
    t.write_bytes(bytes(bad))
    try:
        got = host.exr_load(str(t), True)   # zlib may still inflate garbage of the right size; it must not crash
        assert got.shape[2] == 7
    except host.ExrError:
        pass


def _exr_header(width, height, channels, compression):
    """Minimal single-part scan-line OpenEXR header (version 2): FLOAT channels, data window at the origin."""
    import struct

    def attr(name, typ, payload):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload

    chl = b"".join(c.encode() + b"\0" + struct.pack("<iBBBBii", 2, 0, 0, 0, 0, 1, 1) for c in channels) + b"\0"
    box = struct.pack("<iiii", 0, 0, width - 1, height - 1)
    hdr = struct.pack("<II", 20000630, 2)
    hdr += attr("channels", "chlist", chl) + attr("compression", "compression", bytes([compression]))
    hdr += attr("dataWindow", "box2i", box) + attr("displayWindow", "box2i", box)
    hdr += attr("lineOrder", "lineOrder", b"\0") + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    hdr += attr("screenWindowCenter", "v2f", struct.pack("<ff", 0, 0)) + attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    return hdr + b"\0"


def test_hostile_headers_cannot_make_the_loaders_allocate_the_declared_window(tmp_path):
    """a half-megabyte file may declare a 2^20-wide window with thousands of channels and point every table
    entry at one small chunk. The loaders must refuse it (false + message) before sizing anything from the header — no
    std::bad_alloc through the host process, no gigabytes zero-filled."""
    import struct
    import resource
    channels = ["Bin_%04d" % i for i in range(2000)]
    for compression, lines in ((3, 16), (0, 1)):  # ZIP (16 lines per chunk), NONE
        height = 64
        hdr = _exr_header(1 << 20, height, channels, compression)
        nchunks = (height + lines - 1) // lines
        table_at = len(hdr)
        chunk_at = table_at + 8 * nchunks
        payload = b"\x78\x9c" + b"\0" * 64
        body = struct.pack("<ii", 0, len(payload)) + payload
        data = hdr + b"".join(struct.pack("<Q", chunk_at) for _ in range(nchunks)) + body
        path = str(tmp_path / ("hostile_%d.exr" % compression))
        open(path, "wb").write(data)
        before = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
        for multi in (True, False):
            with pytest.raises(host.ExrError) as e:
                host.exr_load(path, multi)
            assert str(e.value)
        assert resource.getrusage(resource.RUSAGE_SELF).ru_maxrss - before < 200 * 1024  # KiB: nothing near the declared 500+ GB
    # distinct table entries that all name the same block of lines: duplicates are refused as well
    hdr = _exr_header(8, 32, ["R", "G", "B"], 0)
    row = struct.pack("<ii", 0, 8 * 3 * 4) + b"\0" * (8 * 3 * 4)
    table_at = len(hdr)
    offs = [table_at + 8 * 32 + i * len(row) for i in range(32)]
    data = hdr + b"".join(struct.pack("<Q", o) for o in offs) + row * 32  # every chunk says y = 0
    path = str(tmp_path / "dup.exr")
    open(path, "wb").write(data)
    with pytest.raises(host.ExrError):
        host.exr_load(path, False)


def test_own_round_trip_special_values(tmp_path):
    rng = np.random.default_rng(3)
    a = rng.standard_normal((23, 31, 4)).astype(np.float32) * 50
    a[0, 0] = [np.inf, -np.inf, np.nan, -0.0]
    a[1, 1] = [1e-45, 6e-8, 65504.0, 65520.0]     # float denormal, half denormal, half max, rounds to inf in half
    a[2, 2] = [1e38, -1e38, 2049.0, 2051.0]        # ties of the half rounding
    p = str(tmp_path / "multi.exr")
    host.exr_write(p, a, True)
    assert same_bits(host.exr_load(p, True), a)    # FLOAT channels: bit-exact, NaN payload included
    c = np.ascontiguousarray(a[:, :, :3])
    p = str(tmp_path / "rgb.exr")
    host.exr_write(p, c, False)
    back = host.exr_load(p, False)
    exp = as_half(c)
    assert np.array_equal(np.isnan(back), np.isnan(exp))
    assert same_bits(np.nan_to_num(back, nan=7.0), np.nan_to_num(exp, nan=7.0))


# ------------------------------------------------------------------------- live against the reference's library
@needs_ref
@pytest.mark.parametrize("compression", [0, 1, 2, 3, 4], ids=["NONE", "RLE", "ZIPS", "ZIP", "PIZ"])
@pytest.mark.parametrize("pixel_type", [1, 2, 0], ids=["HALF", "FLOAT", "UINT"])
def test_reader_against_reference_writer(tmp_path, compression, pixel_type):
    R = ref()
    rng = np.random.default_rng(100 * compression + pixel_type)
    cases = []
    for (w, h) in ((157, 75), (1, 1), (3, 40), (64, 5), (33, 33)):
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        noisy = (rng.random((5, h, w), dtype=np.float32) * np.float32(300.0)).astype(np.float32)
        noisy[0, : h // 3] = 0
        smooth = np.stack([np.round(xx / 4) / 8, np.round(yy / 3) * 0.25, np.floor((xx + yy) / 16),
                           np.full((h, w), 64, np.float32), xx * 0 + yy]).astype(np.float32)   # few symbols: 14-bit wavelet
        cases += [(w, h, noisy), (w, h, smooth)]
    k = 0
    for (w, h, a) in cases:
        for rgb, n in ((1, 3), (0, 5)):
            for (mx, my, dec) in ((0, 0, 0), (7, 2, 1)):
                data = np.ascontiguousarray(a[:n])
                if pixel_type == 0:
                    data = np.floor(data * 1000).astype(np.float32)
                p = str(tmp_path / ("f%d.exr" % k))
                k += 1
                assert R.exrref_write_generic(os.fsencode(p), data.ctypes.data_as(C.c_void_p), w, h, n, rgb, pixel_type,
                                              compression, mx, my, dec) == 0
                got = host.exr_load(p, not rgb)
                exp = data
                if pixel_type == 1:
                    exp = as_half(data)
                elif pixel_type == 0:
                    exp = data.astype(np.uint32).astype(np.float32)
                if rgb and got.shape[2] == 1:   # grey detection
                    assert same_bits(exp[0], exp[1]) and same_bits(exp[0], exp[2])
                    exp = exp[:1]
                assert same_bits(got.transpose(2, 0, 1), exp), (w, h, rgb, mx, my, dec)


@needs_ref
def test_reference_wrappers_both_directions(tmp_path):
    """io_exr.cpp's four entry points: what the reference's tools write we read; what we write its readers read."""
    R = ref()
    rng = np.random.default_rng(11)
    h, w = 61, 83
    color = (rng.random((h, w, 3), dtype=np.float32) * 4).astype(np.float32)
    color[5, 5] = [3000.0, 1e-7, 70000.0]
    bins = rng.integers(0, 12, (h, w, 61)).astype(np.float32)
    # reference -> ours
    planar = np.ascontiguousarray(color.transpose(2, 0, 1))
    p = str(tmp_path / "ref_rgb.exr")
    R.exrref_write_rgb(os.fsencode(p), planar.ctypes.data_as(C.c_void_p), w, h)
    assert same_bits(host.exr_load(p, False), as_half(color))
    planar = np.ascontiguousarray(bins.transpose(2, 0, 1))
    p = str(tmp_path / "ref_multi.exr")
    R.exrref_write_multi(os.fsencode(p), planar.ctypes.data_as(C.c_void_p), w, h, 61)
    assert same_bits(host.exr_load(p, True), bins)
    # ours -> reference
    p = str(tmp_path / "our_rgb.exr")
    host.exr_write(p, color, False)
    assert same_bits(ref_read_rgb(p), as_half(color).transpose(2, 0, 1))
    p = str(tmp_path / "our_multi.exr")
    host.exr_write(p, bins, True)
    assert same_bits(ref_read_multi(p), bins.transpose(2, 0, 1))


@needs_ref
@pytest.mark.parametrize("compression", [0, 2, 3, 4], ids=["NONE", "ZIPS", "ZIP", "PIZ"])
@pytest.mark.parametrize("pixel_type", [1, 2], ids=["HALF", "FLOAT"])
def test_writer_against_reference_reader(tmp_path, compression, pixel_type):
    """Every compression the writer offers (PIZ: bitmap LUT + wavelet + Huffman with run-length codes), read back by the
    reference's library (classic and fast Huffman decoders) and by our own reader: bit-identical."""
    rng = np.random.default_rng(7 * compression + pixel_type)
    for (h, w) in ((75, 157), (1, 1), (40, 3), (5, 64), (33, 33), (100, 200)):
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        noisy = (rng.random((h, w, 5), dtype=np.float32) * 300).astype(np.float32)
        smooth = np.stack([np.round(xx / 4) / 8, np.round(yy / 3) * 0.25, np.floor((xx + yy) / 16),
                           np.full((h, w), 64, np.float32), xx * 0 + yy], axis=2).astype(np.float32)
        const = np.full((h, w, 5), 3.5, np.float32)
        for a in (noisy, smooth, const):
            for multi, n in ((True, 5), (False, 3)):
                p = str(tmp_path / "w.exr")
                data = np.ascontiguousarray(a[:, :, :n])
                host.exr_write_ex(p, data, multi, pixel_type, compression)
                exp = data if pixel_type == 2 else as_half(data)
                got_ref = ref_read_multi(p) if multi else ref_read_rgb(p)
                assert same_bits(got_ref, exp.transpose(2, 0, 1)), (h, w, multi)
                mine = host.exr_load(p, multi)
                assert same_bits(mine, exp if mine.shape[2] == n else exp[:, :, :1]), (h, w, multi)


@needs_ref
def test_lossy_codecs_are_rejected_not_misread(tmp_path):
    R = ref()
    a = np.random.default_rng(0).random((3, 40, 40), dtype=np.float32)
    for comp, name in ((5, "PXR24"), (6, "B44"), (8, "DWAA")):
        p = str(tmp_path / ("lossy%d.exr" % comp))
        assert R.exrref_write_generic(os.fsencode(p), a.ctypes.data_as(C.c_void_p), 40, 40, 3, 1, 1, comp, 0, 0, 0) == 0
        with pytest.raises(host.ExrError, match=name):
            host.exr_load(p, False)


def test_reader_survives_mutated_files_under_sanitizers(tmp_path):
    """tests/cxx/exr_fuzz.cpp built with AddressSanitizer + UBSan: thousands of mutated copies of the reference-written
    fixtures (bit flips, truncation, extreme length / coordinate fields, spliced blocks) through exrRead — every one is
    either decoded or refused with a message; no crash, no out-of-bounds access, no unbounded allocation."""
    exe = str(tmp_path / "exr_fuzz")
    build = subprocess.run(["g++", "-std=c++11", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                            "-I" + os.path.join(ROOT, "bcd_b200", "io"), "-o", exe, os.path.join(HERE, "cxx", "exr_fuzz.cpp"),
                            os.path.join(ROOT, "bcd_b200", "io", "exr_codec.cpp"), "-lz"],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("sanitizer runtime not available: " + build.stderr[-200:])
    assert build.returncode == 0, build.stderr[-2000:]
    seeds = [os.path.join(GOLD, f) for f in ("ref_color_piz.exr", "ref_bins_zip.exr", "ref_smooth_piz.exr")]
    r = subprocess.run([exe, "1500", "20170627", str(tmp_path / "scratch.exr")] + seeds, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    accepted, rejected = [int(v) for v in r.stdout.split()[1::2]]
    assert accepted + rejected == 4500 and rejected > 1000 and accepted > 100


# ------------------------------------------------------------------------------------------------ .bcd.json
def test_preset_written_by_the_reference_loads():
    got = host.params_load(os.path.join(GOLD, "ref_preset.bcd.json"))
    exp = dict(performSpikeRemovalPrefiltering=0, spikeRemovalThresholdStDevFactor=2.25, nbOfScales=2,
               histoDistanceThreshold=np.float32(0.8), useCuda=0, nbOfCores=4, patchRadius=2, searchWindowRadius=9,
               randomPixelOrder=0, markedPixelsSkippingProbability=0.25, minEigenValue=np.float32(1e-6))
    for k, v in exp.items():
        assert np.float32(got[k]) == np.float32(v), k
    assert got["inputColorFile"] == os.path.join(GOLD, "frame.exr")
    assert got["inputHistoFile"] == os.path.join(GOLD, "aux/frame_hist.exr")
    assert got["inputCovarFile"] == os.path.join(GOLD, "../b/frame_cov.exr")


def test_preset_round_trip_and_partial_files(tmp_path):
    vals = dict(zip(host.PARAM_KEYS, (0, 2.5, 2, 1.25, 0, 4, 2, 9, 0, 0.5, 1e-6)))
    d = tmp_path / "shots" / "a"
    d.mkdir(parents=True)
    p = str(d / "x.bcd.json")
    assert host.params_write(p, vals, str(d / "f.exr"), str(d / "aux" / "f_hist.exr"), str(tmp_path / "shots" / "b" / "f_cov.exr"))
    text = json.load(open(p))    # valid JSON, the reference's keys (ParametersIO.cpp:132-159), names relative to the folder
    assert sorted(text) == sorted(list(host.PARAM_KEYS) + ["inputColorFile", "inputHistoFile", "inputCovarFile"])
    assert (text["inputColorFile"], text["inputHistoFile"], text["inputCovarFile"]) == ("f.exr", "aux/f_hist.exr", "../b/f_cov.exr")
    assert text["useCuda"] is False and text["patchRadius"] == 2 and text["histoDistanceThreshold"] == 1.25
    back = host.params_load(p)
    for k in host.PARAM_KEYS:
        assert np.float32(back[k]) == np.float32(vals[k]), k
    assert back["inputHistoFile"] == str(d / "aux" / "f_hist.exr")
    # a key that is absent keeps the caller's value (ParametersIO.cpp:56-112); other JSON content is ignored
    q = tmp_path / "partial.bcd.json"
    q.write_text('{ "nbOfScales" : 5,\n "comment": {"nested": [1, 2, {"x": "}"}]}, "minEigenValue": 1e-3, "useCuda": true }')
    got = host.params_load(str(q))
    assert got["nbOfScales"] == 5 and np.float32(got["minEigenValue"]) == np.float32(1e-3) and got["useCuda"] == 1
    assert got["patchRadius"] == 1 and got["searchWindowRadius"] == 6 and got["inputColorFile"] == ""
    # malformed / missing
    q.write_text('{ "nbOfScales" : }')
    assert host.params_load(str(q)) is None
    assert host.params_load(str(tmp_path / "nope.bcd.json")) is None


def test_preset_parser_survives_mutated_text(tmp_path):
    """Random edits of a valid preset (deleted / duplicated / replaced characters, truncation): load() returns a dict or
    None, never hangs or crashes, and whatever it accepts keeps the numeric fields finite-or-default typed."""
    good = open(os.path.join(GOLD, "ref_preset.bcd.json")).read()
    rng = np.random.default_rng(42)
    alphabet = list('{}[]",:\\ \n\ttruefalsn0123456789.eE+-u') + ["\\u12", '\\"', "nbOfScales"]
    p = tmp_path / "m.bcd.json"
    ok = 0
    for _ in range(1500):
        t = list(good)
        for _ in range(int(rng.integers(1, 6))):
            i = int(rng.integers(0, len(t)))
            kind = int(rng.integers(0, 4))
            if kind == 0:
                del t[i]
            elif kind == 1:
                t.insert(i, alphabet[int(rng.integers(0, len(alphabet)))])
            elif kind == 2:
                t[i] = alphabet[int(rng.integers(0, len(alphabet)))]
            else:
                t = t[:i]
            if not t:
                break
        p.write_text("".join(t))
        got = host.params_load(str(p))
        if got is not None:
            ok += 1
            assert set(host.PARAM_KEYS) <= set(got)
    assert 0 < ok < 1500


# --------------------------------------------------------------------------------------------- command lines
def run(cmd):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)


def test_raw2bcd_writes_the_three_inputs(tmp_path):
    """raw dump -> <prefix>.exr / _hist.exr / _cov.exr with the reference's conventions (src/raw_converter/main.cpp:
    36-41,142-184): the statistics equal bcd::SamplesAccumulator's, the sample count rides as the last histogram channel."""
    rng = np.random.default_rng(4)
    h, w, spp = 12, 17, 9
    samples = (rng.random((h, w, spp, 3), dtype=np.float32) * 2).astype(np.float32)
    raw = tmp_path / "in.raw"
    with open(raw, "wb") as f:
        f.write(np.array([1, w, h, spp, 3], np.int32).tobytes())
        f.write(samples.tobytes())
    r = run([RAW2BCD, str(raw), str(tmp_path / "out")])
    assert r.returncode == 0, r.stderr
    st = host.accumulate(samples)
    assert same_bits(host.exr_load(str(tmp_path / "out.exr"), False), as_half(st["color"]))
    hist = host.exr_load(str(tmp_path / "out_hist.exr"), True)
    assert same_bits(hist[:, :, :-1], st["hist"]) and same_bits(hist[:, :, -1:], st["nsamples"])
    assert same_bits(host.exr_load(str(tmp_path / "out_cov.exr"), True), st["cov"])
    assert run([RAW2BCD]).returncode == 1
    assert run([RAW2BCD, str(tmp_path / "none.raw"), str(tmp_path / "o")]).returncode == 1


def test_bcd_cli_argument_handling(tmp_path):
    """Option names, defaults and error behaviour of the reference's bcd-cli (src/cli/main.cpp:99-384): exit code 1
    plus a message, never a crash; without a CUDA device the run fails loudly (there is no CPU path)."""
    r = run([BCD_CLI])
    assert r.returncode == 1 and "Missing required program argument(s): -i -h -c -o" in r.stdout and "Usage:" in r.stdout
    for opt in ("-o", "-i", "-h", "-c", "-a", "-d", "-b", "-w", "-r", "-p", "--p-factor", "-m", "-s", "--ncores", "--use-cuda", "-e"):
        assert ("    " + opt + " ") in r.stdout, opt
    r = run([BCD_CLI, "-i"])
    assert r.returncode == 1 and "after '-i'" in r.stdout
    r = run([BCD_CLI, "-m", "1.5"])
    assert r.returncode == 1 and "expecting float in [0,1] after '-m'" in r.stdout
    r = run([BCD_CLI, "-r", "2"])
    assert r.returncode == 1 and "expecting 0 or 1 after '-r'" in r.stdout
    out = str(tmp_path / "o.exr")
    r = run([BCD_CLI, "-o", out, "-i", str(tmp_path / "missing.exr")])
    assert r.returncode == 1 and "assuming" in r.stdout and "couldn't load input color image file" in r.stdout
    # a complete set of inputs: on a machine without a GPU the denoiser refuses instead of falling back to anything
    rng = np.random.default_rng(5)
    st = host.accumulate((rng.random((24, 24, 8, 3), dtype=np.float32) * 2).astype(np.float32))
    host.exr_write(str(tmp_path / "f.exr"), st["color"], False)
    host.exr_write(str(tmp_path / "f_hist.exr"), np.concatenate([st["hist"], st["nsamples"]], axis=2), True)
    host.exr_write(str(tmp_path / "f_cov.exr"), st["cov"], True)
    preset = tmp_path / "p.bcd.json"
    preset.write_text(json.dumps({"inputColorFile": "f.exr", "inputHistoFile": "f_hist.exr", "inputCovarFile": "f_cov.exr",
                                  "nbOfScales": 1, "performSpikeRemovalPrefiltering": False}))
    r = run([BCD_CLI, "-a", str(preset), "-o", out])
    from bcd_b200 import capi
    if capi.lib().bcd_b200_device_count() < 1:
        assert r.returncode == 1 and "denoising failed" in r.stderr, (r.stdout, r.stderr)
    else:
        assert r.returncode == 0 and host.exr_load(out, False).shape == (24, 24, 3)

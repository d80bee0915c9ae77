"""Generates the committed I/O fixtures under tests/golden/io/ — run HERE (build container), where /root/reference
exists. Everything is produced by the reference's own code:

  ref_color_piz.exr   writeImageEXR     (src/io/exr/io_exr.cpp:126-164: HALF RGBA, PIZ)   <- oracle/_ref/libexr_ref.so
  ref_bins_zip.exr    writeMultiImageEXR (io_exr.cpp:215-241: FLOAT Bin_%04d, ZIP)        <- oracle/_ref/libexr_ref.so
  ref_smooth_piz.exr  OpenEXR API, FLOAT R,G,B, PIZ on a smooth image (both wavelet variants, Huffman runs)
  ref_io_values.npz   the arrays that went in
  ref_preset.bcd.json bcd::ParametersIO::write (src/io/ParametersIO.cpp:116-162 with the vendored nlohmann::json),
                      compiled on the spot from the reference's sources

    python tests/golden/make_golden_io.py
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "io")
REF = "/root/reference"

PRESET_MAIN = r'''
#include "ParametersIO.h"
int main(int argc, char** argv)
{
	bcd::PipelineParameters p;
	p.m_inputFileNames.m_colors = "/data/shots/a/frame.exr";
	p.m_inputFileNames.m_histograms = "/data/shots/a/aux/frame_hist.exr";
	p.m_inputFileNames.m_covariances = "/data/shots/b/frame_cov.exr";
	p.m_prefilteringParameters.m_performSpikeRemoval = false;
	p.m_prefilteringParameters.m_spikeRemovalThresholdStDevFactor = 2.25f;
	p.m_denoiserParameters.m_nbOfScales = 2;
	bcd::DenoiserParameters& d = p.m_denoiserParameters.m_monoscaleParameters;
	d.m_histogramDistanceThreshold = 0.8f;
	d.m_patchRadius = 2;
	d.m_searchWindowRadius = 9;
	d.m_minEigenValue = 1.e-6f;
	d.m_useRandomPixelOrder = false;
	d.m_markedPixelsSkippingProbability = 0.25f;
	d.m_nbOfCores = 4;
	d.m_useCuda = false;
	bcd::ParametersIO::write(p, argv[1]);
	return 0;
}
'''


def main():
    os.makedirs(OUT, exist_ok=True)
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libexr_ref.so"))
    rng = np.random.default_rng(20170627)
    h, w = 40, 52
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    color = np.stack([0.2 + xx / w, 0.5 + 0.4 * np.sin(yy / 5), (xx + yy) / (w + h)]).astype(np.float32)
    color *= (1 + 0.3 * rng.standard_normal(color.shape)).astype(np.float32)
    color[0, 3, 4] = 1234.5          # firefly
    color[:, 10:14, 20:30] = 0.0     # flat block
    R.exrref_write_rgb(os.path.join(OUT, "ref_color_piz.exr").encode(), color.ctypes.data_as(C.c_void_p), w, h)
    bins = rng.integers(0, 9, (7, h, w)).astype(np.float32)
    bins[5] = rng.random((h, w), dtype=np.float32) * 1e-3
    bins[6] = 81.0
    R.exrref_write_multi(os.path.join(OUT, "ref_bins_zip.exr").encode(), bins.ctypes.data_as(C.c_void_p), w, h, 7)
    smooth = np.stack([np.round(xx / 4) / 8, np.round(yy / 3) * 0.25, np.floor((xx + yy) / 16)]).astype(np.float32)
    assert R.exrref_write_generic(os.path.join(OUT, "ref_smooth_piz.exr").encode(), smooth.ctypes.data_as(C.c_void_p),
                                  w, h, 3, 1, 2, 4, 0, 0, 0) == 0
    np.savez_compressed(os.path.join(OUT, "ref_io_values.npz"), color=color, bins=bins, smooth=smooth)

    # scratch inside the repository (nothing outside /root/repo is written); the preset sits in <scratch>/shots/a/ and
    # names its inputs relative to that folder
    with tempfile.TemporaryDirectory(dir=OUT) as tmp:
        src = os.path.join(tmp, "main.cpp")
        open(src, "w").write(PRESET_MAIN.replace("/data/", tmp + "/"))
        exe = os.path.join(tmp, "writer")
        # the reference's write() has no return statement (ParametersIO.cpp:162): GCC ends the function with a trap, which
        # fires after the file has been written and closed, so the exit status is ignored and the file checked instead
        subprocess.run(["g++", "-std=c++11", "-O0", "-w", "-include", "cmath", "-I%s/include/bcd/core" % REF,
                        "-I%s/include/bcd/io" % REF, "-I%s/ext/json/src" % REF, "-o", exe, src,
                        "%s/src/io/ParametersIO.cpp" % REF, "%s/src/core/Utils.cpp" % REF], check=True)
        os.makedirs(os.path.join(tmp, "shots", "a"))
        preset = os.path.join(tmp, "shots", "a", "ref_preset.bcd.json")
        subprocess.run([exe, preset], check=False, stdout=subprocess.DEVNULL)
        import json
        assert len(json.load(open(preset))) == 14
        os.replace(preset, os.path.join(OUT, "ref_preset.bcd.json"))
    print(open(os.path.join(OUT, "ref_preset.bcd.json")).read())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

// TEST INFRASTRUCTURE — not part of the product path.
//
// C shim over the UNMODIFIED reference CPU implementation (superboubek/bcd), compiled in place from
// /root/reference/src/core/*.cpp by oracle/Makefile into oracle/_ref/libbcd_ref.so.
// No reference source is copied into this repository; this file only *calls* the reference's public API
// (bcd::Denoiser / bcd::MultiscaleDenoiser / bcd::SamplesAccumulator, include/bcd/core/IDenoiser.h:73-97,
// SamplesAccumulator.h:48-80) so that tests and bench.py's cpu_baseline / --impl reference arm can run
// the real reference on the same host buffers as the B200 path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library.

#include "Denoiser.h"
#include "MultiscaleDenoiser.h"
#include "IDenoiser.h"
#include "DeepImage.h"
#include "SamplesAccumulator.h"
#include "SpikeRemovalFilter.h"

#include <cstring>
#include <iostream>
#include <sstream>
#include <memory>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace
{
	struct CoutSilencer
	{
		std::streambuf* m_old;
		std::ostringstream m_sink;
		explicit CoutSilencer(bool i_on) : m_old(nullptr) { if(i_on) m_old = std::cout.rdbuf(m_sink.rdbuf()); }
		~CoutSilencer() { if(m_old) std::cout.rdbuf(m_old); }
	};

	void wrap(bcd::Deepimf& o_img, const float* i_p, int i_w, int i_h, int i_d)
	{
		o_img.resize(i_w, i_h, i_d);
		std::memcpy(o_img.getDataPtr(), i_p, sizeof(float) * size_t(i_w) * i_h * i_d);
	}
}

extern "C"
{

int bcdref_max_threads()
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

// Runs the reference Denoiser (nb_scales == 1) or MultiscaleDenoiser (nb_scales > 1) on host buffers.
// Layout: HWC float32; cov order xx,yy,zz,yz,xz,xy. out_rgb is W*H*3.
int bcdref_denoise(
		const float* color, const float* nsamples, const float* hist, const float* cov,
		int width, int height, int hist_depth,
		float tau, int patch_radius, int search_radius, float min_eig,
		int random_order, float skip_prob, int nb_cores, int nb_scales,
		float* out_rgb, int quiet)
{
	CoutSilencer silencer(quiet != 0);
	bcd::Deepimf colorImg, nsImg, histImg, covImg, outImg;
	wrap(colorImg, color, width, height, 3);
	wrap(nsImg, nsamples, width, height, 1);
	wrap(histImg, hist, width, height, hist_depth);
	wrap(covImg, cov, width, height, 6);
	outImg = colorImg; // callers pre-size the output as a copy of the colour image (src/cli/main.cpp:445)

	bcd::DenoiserInputs inputs;
	inputs.m_pColors = &colorImg;
	inputs.m_pNbOfSamples = &nsImg;
	inputs.m_pHistograms = &histImg;
	inputs.m_pSampleCovariances = &covImg;
	bcd::DenoiserOutputs outputs;
	outputs.m_pDenoisedColors = &outImg;
	bcd::DenoiserParameters params;
	params.m_histogramDistanceThreshold = tau;
	params.m_patchRadius = patch_radius;
	params.m_searchWindowRadius = search_radius;
	params.m_minEigenValue = min_eig;
	params.m_useRandomPixelOrder = (random_order != 0);
	params.m_markedPixelsSkippingProbability = skip_prob;
	params.m_nbOfCores = nb_cores;
	params.m_useCuda = false;

	std::unique_ptr<bcd::IDenoiser> uDenoiser;
	if(nb_scales > 1)
		uDenoiser.reset(new bcd::MultiscaleDenoiser(nb_scales));
	else
		uDenoiser.reset(new bcd::Denoiser());
	uDenoiser->setInputs(inputs);
	uDenoiser->setOutputs(outputs);
	uDenoiser->setParameters(params);
	bool ok = uDenoiser->denoise();
	if(!ok)
		return -1;
	std::memcpy(out_rgb, outImg.getDataPtr(), sizeof(float) * size_t(width) * height * 3);
	return 0;
}

// Feeds raw samples (pixel-major: [line][col][sample][rgb]) through the reference SamplesAccumulator with
// weight 1 and default-constructed HistogramParameters overridden by the arguments.
int bcdref_accumulate(
		const float* samples, int width, int height, int spp,
		int nb_bins, float gamma, float max_value,
		float* color, float* nsamples, float* hist, float* cov)
{
	bcd::HistogramParameters hp;
	hp.m_nbOfBins = nb_bins;
	hp.m_gamma = gamma;
	hp.m_maxValue = max_value;
	bcd::SamplesAccumulator acc(width, height, hp);
	const float* p = samples;
	for(int l = 0; l < height; ++l)
		for(int c = 0; c < width; ++c)
			for(int s = 0; s < spp; ++s, p += 3)
				acc.addSample(l, c, p[0], p[1], p[2]);
	bcd::SamplesStatisticsImages stats = acc.extractSamplesStatistics();
	size_t n = size_t(width) * height;
	std::memcpy(color, stats.m_meanImage.getDataPtr(), sizeof(float) * n * 3);
	std::memcpy(nsamples, stats.m_nbOfSamplesImage.getDataPtr(), sizeof(float) * n);
	std::memcpy(hist, stats.m_histoImage.getDataPtr(), sizeof(float) * n * 3 * nb_bins);
	std::memcpy(cov, stats.m_covarImage.getDataPtr(), sizeof(float) * n * 6);
	return 0;
}

// Reference SpikeRemovalFilter (the prefilter immediately before the path), in place.
int bcdref_spike_removal(
		float* color, float* nsamples, float* hist, float* cov,
		int width, int height, int hist_depth, float threshold_std_factor)
{
	bcd::Deepimf colorImg, nsImg, histImg, covImg;
	wrap(colorImg, color, width, height, 3);
	wrap(nsImg, nsamples, width, height, 1);
	wrap(histImg, hist, width, height, hist_depth);
	wrap(covImg, cov, width, height, 6);
	bcd::SpikeRemovalFilter::filter(colorImg, nsImg, histImg, covImg, threshold_std_factor);
	size_t n = size_t(width) * height;
	std::memcpy(color, colorImg.getDataPtr(), sizeof(float) * n * 3);
	std::memcpy(nsamples, nsImg.getDataPtr(), sizeof(float) * n);
	std::memcpy(hist, histImg.getDataPtr(), sizeof(float) * n * hist_depth);
	std::memcpy(cov, covImg.getDataPtr(), sizeof(float) * n * 6);
	return 0;
}

} // extern "C"

// Synthetic benchmark inputs G(W,H,spp,seed,sigma) of SURVEY §8d and the ctypes-facing host helpers
// (include/bcd_b200_host.h). Host code only; nothing here is on the denoising hot path.
#include "bcd_b200_host.h"
#include "SamplesAccumulator.h"
#ifndef BCD_SYNTH_ONLY
#include "Denoiser.h"
#include "MultiscaleDenoiser.h"
#endif

#include <cmath>
#include <cstring>
#include <memory>

namespace
{
	inline uint64_t splitmix64(uint64_t x)
	{
		uint64_t z = x + 0x9E3779B97F4A7C15ull;
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		return z ^ (z >> 31);
	}
	inline double u01(uint64_t k) { return (double(splitmix64(k) >> 11) + 0.5) / 9007199254740992.0; }

	struct Shape { int type; double cx, cy, r, ry, v[3]; };

	void copyOut(const bcd::SamplesStatisticsImages& s, float* color, float* nsamples, float* hist, float* cov)
	{
		std::memcpy(color, s.m_meanImage.getDataPtr(), sizeof(float) * s.m_meanImage.getSize());
		std::memcpy(nsamples, s.m_nbOfSamplesImage.getDataPtr(), sizeof(float) * s.m_nbOfSamplesImage.getSize());
		std::memcpy(hist, s.m_histoImage.getDataPtr(), sizeof(float) * s.m_histoImage.getSize());
		std::memcpy(cov, s.m_covarImage.getDataPtr(), sizeof(float) * s.m_covarImage.getSize());
	}
}

extern "C" int bcd_synth_generate(int W, int H, int spp, uint64_t s, double sigma,
		int nb_bins, float gamma, float max_value, int row0, int rows,
		float* color, float* nsamples, float* hist, float* cov)
{
	const double kTwoPi = 6.283185307179586476925286766559;
	const double g[3] = { 1.0, 0.8, 0.6 };
	Shape shapes[12];
	for(uint64_t i = 0; i < 12; ++i)
	{
		Shape& sh = shapes[i];
		sh.type = int(i & 1);
		sh.cx = u01(7 * s + 11 * i + 1) * W;
		sh.cy = u01(7 * s + 11 * i + 2) * H;
		sh.r = (0.03 + 0.09 * u01(7 * s + 11 * i + 3)) * W;
		sh.ry = sh.r * (0.5 + u01(7 * s + 11 * i + 4));
		for(uint64_t c = 0; c < 3; ++c)
			sh.v[c] = 0.1 + 1.4 * u01(7 * s + 11 * i + 5 + c);
	}
	bcd::HistogramParameters hp;
	hp.m_nbOfBins = nb_bins;
	hp.m_gamma = gamma;
	hp.m_maxValue = max_value;
	bcd::SamplesAccumulator acc(W, rows, hp); // pixels are independent, so rows can be filled concurrently
#pragma omp parallel for schedule(dynamic, 4)
	for(int ly = 0; ly < rows; ++ly)
	{
		const int y = row0 + ly;
		for(int x = 0; x < W; ++x)
		{
			double L[3];
			for(int c = 0; c < 3; ++c)
				L[c] = (0.05 + 0.9 * double(x) / W) * g[c];
			for(int i = 0; i < 12; ++i)
			{
				const Shape& sh = shapes[i];
				const double dx = x - sh.cx, dy = y - sh.cy;
				const bool inside = sh.type == 0 ? (dx * dx + dy * dy <= sh.r * sh.r) : (std::fabs(dx) <= sh.r && std::fabs(dy) <= sh.ry);
				if(inside)
					for(int c = 0; c < 3; ++c)
						L[c] = sh.v[c];
			}
			if(3 * y > 2 * H)
				for(int c = 0; c < 3; ++c)
					L[c] = std::fmax(0.02, L[c] + 0.15 * std::sin(kTwoPi * x / 37.0) * std::sin(kTwoPi * y / 53.0));
			if(2 * x >= W)
			{
				const double albedo = 0.3 + 1.4 * u01(13 * s + uint64_t(y) * uint64_t(W) + uint64_t(x));
				for(int c = 0; c < 3; ++c)
					L[c] *= albedo;
			}
			for(int k = 0; k < spp; ++k)
			{
				const uint64_t base = ((s * 1000003ull + uint64_t(y)) * 1000003ull + uint64_t(x)) * 1000003ull + uint64_t(k);
				const bool firefly = u01(8 * base + 7) < 2e-3;
				float rgb[3];
				for(uint64_t c = 0; c < 3; ++c)
				{
					const double z = std::sqrt(-2.0 * std::log(u01(8 * base + 2 * c))) * std::cos(kTwoPi * u01(8 * base + 2 * c + 1));
					double v = L[c] * std::exp(sigma * z - 0.5 * sigma * sigma);
					if(firefly)
						v *= 20.0;
					rgb[c] = float(v);
				}
				acc.addSample(ly, x, rgb[0], rgb[1], rgb[2]);
			}
		}
	}
	copyOut(acc.extractSamplesStatistics(), color, nsamples, hist, cov);
	return 0;
}

extern "C" int bcd_host_accumulate(const float* samples, int width, int height, int spp, int nb_bins, float gamma, float max_value,
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
			for(int k = 0; k < spp; ++k, p += 3)
				acc.addSample(l, c, p[0], p[1], p[2]);
	copyOut(acc.extractSamplesStatistics(), color, nsamples, hist, cov);
	return 0;
}

#ifndef BCD_SYNTH_ONLY
extern "C" int bcd_host_cxx_denoise(const float* color, const float* nsamples, const float* hist, const float* cov,
		int width, int height, int hist_depth,
		float tau, int patch_radius, int search_radius, float min_eig, int random_order, float skip_prob,
		int nb_scales, int device, float* out_rgb, int* progress_calls)
{
	bcd::Deepimf colorImg(width, height, 3), nsImg(width, height, 1), histImg(width, height, hist_depth), covImg(width, height, 6);
	colorImg.copyDataFrom(color);
	nsImg.copyDataFrom(nsamples);
	histImg.copyDataFrom(hist);
	covImg.copyDataFrom(cov);
	bcd::Deepimf outImg(colorImg); // callers pre-size the output as a copy of the colour image (reference src/cli/main.cpp:445)

	bcd::DenoiserInputs inputs;
	inputs.m_pColors = &colorImg;
	inputs.m_pNbOfSamples = &nsImg;
	inputs.m_pHistograms = &histImg;
	inputs.m_pSampleCovariances = &covImg;
	bcd::DenoiserOutputs outputs;
	outputs.m_pDenoisedColors = &outImg;
	bcd::DenoiserParameters parameters;
	parameters.m_histogramDistanceThreshold = tau;
	parameters.m_patchRadius = patch_radius;
	parameters.m_searchWindowRadius = search_radius;
	parameters.m_minEigenValue = min_eig;
	parameters.m_useRandomPixelOrder = random_order != 0;
	parameters.m_markedPixelsSkippingProbability = skip_prob;

	std::unique_ptr<bcd::Denoiser> uDenoiser(nb_scales > 1 ? new bcd::MultiscaleDenoiser(nb_scales) : new bcd::Denoiser());
	uDenoiser->setDevice(device);
	bcd::IDenoiser& rDenoiser = *uDenoiser; // use it through the interface, like src/cli/main.cpp:457-468
	rDenoiser.setInputs(inputs);
	rDenoiser.setOutputs(outputs);
	rDenoiser.setParameters(parameters);
	int calls = 0;
	rDenoiser.setProgressCallback([&calls](float) { ++calls; });
	const bool ok = rDenoiser.denoise();
	if(progress_calls)
		*progress_calls = calls;
	if(ok)
		outImg.copyDataTo(out_rgb);
	return ok ? 1 : 0;
}

// ---- a host renderer's view of the drop-in, kept across frames (bench.py's e2e_plugin leg, tests): one Denoiser object,
// one set of DeepImages in ordinary pageable std::vector storage, denoise() called frame after frame on them
struct bcd_host_session
{
	std::unique_ptr<bcd::Denoiser> denoiser;
	bcd::Deepimf color, nsamples, hist, cov, out;
};

extern "C" bcd_host_session* bcd_host_session_create(int nb_scales, const int* devices, int nb_devices, int pin_host)
{
	bcd_host_session* s = new bcd_host_session();
	s->denoiser.reset(nb_scales > 1 ? new bcd::MultiscaleDenoiser(nb_scales) : new bcd::Denoiser());
	if(nb_devices != 0 || devices)
		s->denoiser->setDevices(devices, nb_devices);
	s->denoiser->setPinHostBuffers(pin_host != 0);
	return s;
}

extern "C" void bcd_host_session_destroy(bcd_host_session* s) { delete s; }

extern "C" int bcd_host_session_load(bcd_host_session* s, const float* color, const float* nsamples, const float* hist,
		const float* cov, int width, int height, int hist_depth)
{
	if(!s) return 0;
	// resize() keeps the storage when the size is unchanged: the renderer overwrites its images in place
	if(s->color.getWidth() != width || s->color.getHeight() != height || s->hist.getDepth() != hist_depth)
	{
		s->color.resize(width, height, 3);
		s->nsamples.resize(width, height, 1);
		s->hist.resize(width, height, hist_depth);
		s->cov.resize(width, height, 6);
		s->out.resize(width, height, 3);
	}
	s->color.copyDataFrom(color);
	s->nsamples.copyDataFrom(nsamples);
	s->hist.copyDataFrom(hist);
	s->cov.copyDataFrom(cov);
	return 1;
}

extern "C" int bcd_host_session_denoise(bcd_host_session* s, float tau, int patch_radius, int search_radius, float min_eig,
		int random_order, float skip_prob)
{
	if(!s) return 0;
	bcd::DenoiserInputs inputs;
	inputs.m_pColors = &s->color;
	inputs.m_pNbOfSamples = &s->nsamples;
	inputs.m_pHistograms = &s->hist;
	inputs.m_pSampleCovariances = &s->cov;
	bcd::DenoiserOutputs outputs;
	outputs.m_pDenoisedColors = &s->out;
	bcd::DenoiserParameters parameters;
	parameters.m_histogramDistanceThreshold = tau;
	parameters.m_patchRadius = patch_radius;
	parameters.m_searchWindowRadius = search_radius;
	parameters.m_minEigenValue = min_eig;
	parameters.m_useRandomPixelOrder = random_order != 0;
	parameters.m_markedPixelsSkippingProbability = skip_prob;
	bcd::IDenoiser& rDenoiser = *s->denoiser;
	rDenoiser.setInputs(inputs);
	rDenoiser.setOutputs(outputs);
	rDenoiser.setParameters(parameters);
	return rDenoiser.denoise() ? 1 : 0;
}

extern "C" int bcd_host_session_fetch(bcd_host_session* s, float* out_rgb)
{
	if(!s || s->out.isEmpty()) return 0;
	s->out.copyDataTo(out_rgb);
	return 1;
}

extern "C" int bcd_host_session_nb_devices_used(bcd_host_session* s)
{
	return s ? s->denoiser->getParameters().m_nbOfCores : 0;
}
#endif // BCD_SYNTH_ONLY

// Host-renderer style use of the re-authored bcd:: API (compiled by tests/test_host_api.py against include/bcd/core
// with bare header names, like the reference's own apps: src/cli/main.cpp:9-24). Checks the validation behaviour of
// Denoiser::inputsOutputsAreOk (reference Denoiser.cpp:269-346) and, when a GPU is present (argv[1] == "gpu"),
// that denoise() fills the output image.
#include "Denoiser.h"
#include "MultiscaleDenoiser.h"
#include "IDenoiser.h"
#include "DeepImage.h"
#include "SamplesAccumulator.h"
#include "CovarianceMatrix.h"
#include "SpikeRemovalFilter.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>

using namespace bcd;

static int g_failures = 0;
#define CHECK(cond) do { if(!(cond)) { std::printf("CHECK FAILED line %d: %s\n", __LINE__, #cond); ++g_failures; } } while(0)

int main(int argc, char** argv)
{
	const bool gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;
	// --- DeepImage container semantics
	Deepimf img(5, 4, 3);
	img.fill(1.f);
	img.set(2, 3, 1, 7.f);
	CHECK(img.get(2, 3, 1) == 7.f && img.get(PixelPosition(2, 3), 1) == 7.f);
	CHECK(img.getDataPtr()[(2 * 5 + 3) * 3 + 1] == 7.f);
	CHECK(img.getSize() == 60 && img.getWidth() == 5 && img.getHeight() == 4 && img.getDepth() == 3);
	int n = 0;
	for(float* p : img) { n += (p[0] == 1.f); }
	CHECK(n == 20);
	Deepimf moved(std::move(img));
	CHECK(img.isEmpty() && moved.getSize() == 60);
	CHECK(moved.clamp(PixelPosition(-3, 9)) == PixelPosition(0, 4));
	CHECK(int(ESymMatData::e_yz) == 3 && int(ESymMatData::e_xy) == 5);

	// --- accumulate a tiny frame like a renderer would
	const int W = 40, H = 32, spp = 24;
	SamplesAccumulator acc(W, H, HistogramParameters());
	unsigned s = 12345u;
	for(int l = 0; l < H; ++l)
		for(int c = 0; c < W; ++c)
			for(int k = 0; k < spp; ++k)
			{
				s = s * 1664525u + 1013904223u;
				const float u = (s >> 8) * (1.f / 16777216.f);
				const float base = c < W / 2 ? 0.2f : 0.8f;
				acc.addSample(l, c, base * (0.5f + u), base * (0.4f + u), base * (0.3f + u));
			}
	SamplesStatisticsImages stats = acc.extractSamplesStatistics();
	CHECK(stats.m_histoImage.getDepth() == 60 && stats.m_nbOfSamplesImage.get(3, 3, 0) == float(spp));

	// --- validation path (no GPU needed): null / empty / mismatched
	DenoiserInputs inputs;
	DenoiserOutputs outputs;
	Deepimf out(stats.m_meanImage);
	{
		Denoiser d;
		d.setInputs(inputs);
		d.setOutputs(outputs);
		CHECK(!d.inputsOutputsAreOk());
		CHECK(!d.denoise());
	}
	inputs.m_pColors = &stats.m_meanImage;
	inputs.m_pNbOfSamples = &stats.m_nbOfSamplesImage;
	inputs.m_pHistograms = &stats.m_histoImage;
	inputs.m_pSampleCovariances = &stats.m_covarImage;
	outputs.m_pDenoisedColors = &out;
	{
		Deepimf wrong(W + 1, H, 1);
		DenoiserInputs bad = inputs;
		bad.m_pNbOfSamples = &wrong;
		Denoiser d;
		d.setInputs(bad);
		d.setOutputs(outputs);
		CHECK(!d.denoise());
		Deepimf empty;
		bad = inputs;
		bad.m_pHistograms = &empty;
		d.setInputs(bad);
		CHECK(!d.denoise());
	}
	// --- row f1: the prefilter bcd_cli runs before denoising (src/cli/main.cpp:428-434), in place on the four images
	if(gpu)
	{
		SamplesStatisticsImages spiky(stats);
		spiky.m_meanImage.set(10, 10, 0, 40.f); // a firefly
		const float before = spiky.m_meanImage.get(10, 10, 0);
		SpikeRemovalFilter::filter(spiky.m_meanImage, spiky.m_nbOfSamplesImage, spiky.m_histoImage, spiky.m_covarImage, 2.f);
		CHECK(before == 40.f && spiky.m_meanImage.get(10, 10, 0) < 2.f);
		CHECK(spiky.m_meanImage.get(20, 5, 1) == stats.m_meanImage.get(20, 5, 1));
	}
	// --- the real thing through the abstract interface
	DenoiserParameters params;
	CHECK(params.m_searchWindowRadius == 6 && params.m_patchRadius == 1 && params.m_markedPixelsSkippingProbability == 1.f);
	params.m_useRandomPixelOrder = false;
	std::unique_ptr<IDenoiser> uDenoiser(new MultiscaleDenoiser(2));
	uDenoiser->setInputs(inputs);
	uDenoiser->setOutputs(outputs);
	uDenoiser->setParameters(params);
	float last = -1.f;
	bool monotone = true;
	uDenoiser->setProgressCallback([&](float p) { monotone = monotone && p >= last; last = p; });
	const bool ok = uDenoiser->denoise();
	if(gpu)
	{
		CHECK(ok);
		CHECK(monotone && last == 1.f);
		CHECK(out.getWidth() == W && out.getHeight() == H && out.getDepth() == 3);
		bool finite = true;
		double diff = 0.;
		for(int i = 0; i < out.getSize(); ++i)
		{
			finite = finite && std::isfinite(out.get(i));
			diff += std::fabs(out.get(i) - stats.m_meanImage.get(i));
		}
		CHECK(finite);
		CHECK(diff > 0.);
		CHECK(uDenoiser->getParameters().m_nbOfCores == 1);
	}
	else
		CHECK(!ok); // no device: denoise() must fail (never silently fall back to a CPU path)
	// --- bcd_b200 extension: several GPUs behind the same call (here two row tiles on device 0, so that a one-GPU box can
	//     run it), and the object reused frame after frame like the reference's GUI does (src/gui/GuiWindow.cpp:345-376)
	if(gpu)
	{
		Deepimf outSingle(stats.m_meanImage), outMulti(stats.m_meanImage);
		DenoiserOutputs oSingle, oMulti;
		oSingle.m_pDenoisedColors = &outSingle;
		oMulti.m_pDenoisedColors = &outMulti;
		Denoiser single;
		single.setInputs(inputs);
		single.setOutputs(oSingle);
		single.setParameters(params);
		CHECK(single.denoise());
		CHECK(single.getNbOfDevices() == 1 && single.getParameters().m_nbOfCores == 1);
		Denoiser multi;
		const int ids[2] = { 0, 0 };
		multi.setDevices(ids, 2);
		multi.setInputs(inputs);
		multi.setOutputs(oMulti);
		multi.setParameters(params);
		for(int frame = 0; frame < 2; ++frame) // the second call reuses the device session (handles, peer links, pinned ranges)
		{
			outMulti.fill(-1.f);
			CHECK(multi.denoise());
			CHECK(multi.getNbOfDevices() == 2 && multi.getParameters().m_nbOfCores == 2); // two tiles worked on the frame
			double worst = 0., scale = 0.;
			for(int i = 0; i < outSingle.getSize(); ++i)
			{
				worst = std::fmax(worst, std::fabs(double(outMulti.get(i)) - double(outSingle.get(i))));
				scale = std::fmax(scale, std::fabs(double(outSingle.get(i))));
			}
			CHECK(worst <= 1e-5 * scale); // only the order of the float atomics differs between the tilings
		}
		multi.releaseDeviceResources();
		CHECK(multi.denoise()); // and comes back after an explicit release
	}
	std::printf("dropin_check: %d failure(s)\n", g_failures);
	return g_failures ? 1 : 0;
}

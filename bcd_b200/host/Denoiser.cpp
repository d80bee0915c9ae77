// bcd::Denoiser / bcd::MultiscaleDenoiser for the B200 path: validation with the reference's messages, then one call
// into the C ABI (include/bcd_b200.h). Replaces the CPU drivers of the reference:
//   Denoiser::denoise              /root/reference/src/core/Denoiser.cpp:84-212
//   Denoiser::inputsOutputsAreOk   /root/reference/src/core/Denoiser.cpp:238-348
//   MultiscaleDenoiser::denoise    /root/reference/src/core/MultiscaleDenoiser.cpp:31-136
#include "Denoiser.h"
#include "MultiscaleDenoiser.h"
#include "bcd_b200.h"

#include <iostream>

namespace bcd
{

	namespace
	{
		void forwardProgress(float i_progress, void* i_pUser)
		{
			(*static_cast<std::function<void(float)>*>(i_pUser))(i_progress);
		}

		bool sameSize(const Deepimf& a, const Deepimf& b) { return a.getWidth() == b.getWidth() && a.getHeight() == b.getHeight(); }

		void reportSizeMismatch(const char* i_pName, const Deepimf& i_rImage, const Deepimf& i_rColors)
		{
			std::cerr << "Aborting denoising: input " << i_pName << " image is " << i_rImage.getWidth() << "x" << i_rImage.getHeight()
					<< "but input color image is " << i_rColors.getWidth() << "x" << i_rColors.getHeight() << std::endl;
		}
	}

	bool Denoiser::inputsOutputsAreOk()
	{
		struct Named { const Deepimf* p; const char* name; };
		const Named images[4] = {
			{ m_inputs.m_pColors, "color" },
			{ m_inputs.m_pNbOfSamples, "number of samples" },
			{ m_inputs.m_pHistograms, "histogram" },
			{ m_inputs.m_pSampleCovariances, "covariance" } };
		bool ok = true;
		for(const Named& img : images)
			if(!img.p)
			{
				std::cerr << "Aborting denoising: nullptr for input " << img.name << " image" << std::endl;
				ok = false;
			}
		if(!ok)
			return false;
		for(const Named& img : images)
			if(img.p->isEmpty())
			{
				std::cerr << "Aborting denoising: input " << img.name << " image is empty" << std::endl;
				ok = false;
			}
		if(!ok)
			return false;
		for(int i = 1; i < 4; ++i)
			if(!sameSize(*images[i].p, *images[0].p))
			{
				reportSizeMismatch(images[i].name, *images[i].p, *images[0].p);
				ok = false;
			}
		return ok;
	}

	bool Denoiser::runOnDevice(int i_nbOfScales)
	{
		if(!inputsOutputsAreOk())
			return false;
		if(!m_outputs.m_pDenoisedColors)
		{
			std::cerr << "Aborting denoising: nullptr for output image" << std::endl;
			return false;
		}
		const Deepimf& rColors = *m_inputs.m_pColors;
		if(rColors.getDepth() != 3 || m_inputs.m_pNbOfSamples->getDepth() != 1 || m_inputs.m_pSampleCovariances->getDepth() != 6)
		{
			std::cerr << "Aborting denoising: expected depths 3 (color), 1 (number of samples), 6 (covariance)" << std::endl;
			return false;
		}
		bcd_b200_images images;
		images.color = rColors.getDataPtr();
		images.nsamples = m_inputs.m_pNbOfSamples->getDataPtr();
		images.hist = m_inputs.m_pHistograms->getDataPtr();
		images.cov = m_inputs.m_pSampleCovariances->getDataPtr();
		images.width = rColors.getWidth();
		images.height = rColors.getHeight();
		images.hist_depth = m_inputs.m_pHistograms->getDepth();

		bcd_b200_params params;
		bcd_b200_default_params(&params);
		params.tau = m_parameters.m_histogramDistanceThreshold;
		params.patch_radius = m_parameters.m_patchRadius;
		params.search_radius = m_parameters.m_searchWindowRadius;
		params.min_eig = m_parameters.m_minEigenValue;
		params.random_order = m_parameters.m_useRandomPixelOrder ? 1 : 0;
		params.skip_prob = m_parameters.m_markedPixelsSkippingProbability;
		params.nb_scales = i_nbOfScales;

		// the reference overwrites m_nbOfCores with the real team size (Denoiser.cpp:121); one GPU plays that role here
		m_parameters.m_nbOfCores = 1;

		// the output is resized to W x H x 3 and overwritten (Denoiser.cpp:207-208)
		m_outputs.m_pDenoisedColors->resize(images.width, images.height, 3);
		const int rc = bcd_b200_denoise(&images, &params, m_outputs.m_pDenoisedColors->getDataPtr(), m_device,
				&forwardProgress, &m_progressCallback);
		if(rc != BCD_B200_OK)
		{
			std::cerr << "Aborting denoising: " << bcd_b200_last_error() << std::endl;
			return false;
		}
		return true;
	}

	bool Denoiser::denoise()
	{
		return runOnDevice(1);
	}

	bool MultiscaleDenoiser::denoise()
	{
		return runOnDevice(m_nbOfScales);
	}

} // namespace bcd

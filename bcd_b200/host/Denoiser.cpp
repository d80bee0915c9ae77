// bcd::Denoiser / bcd::MultiscaleDenoiser for the B200 path: validation with the reference's messages, then one call
// into the C ABI (include/bcd_b200.h). Replaces the CPU drivers of the reference:
//   Denoiser::denoise              /root/reference/src/core/Denoiser.cpp:84-212
//   Denoiser::inputsOutputsAreOk   /root/reference/src/core/Denoiser.cpp:238-348
//   MultiscaleDenoiser::denoise    /root/reference/src/core/MultiscaleDenoiser.cpp:31-136
#include "Denoiser.h"
#include "MultiscaleDenoiser.h"
#include "bcd_b200.h"

#include <cstdlib>
#include <cstring>
#include <algorithm>
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

	Denoiser::Denoiser() : IDenoiser(), m_devices(1, 0), m_allDevices(false), m_pinHostBuffers(true), m_pSession(nullptr)
	{
		const char* pPin = std::getenv("BCD_B200_PIN"); // "0": never page-lock the caller's image storage
		if(pPin && pPin[0] == '0')
			m_pinHostBuffers = false;
		// BCD_B200_DEVICES: "all", or a comma-separated list of CUDA device ordinals
		const char* pEnv = std::getenv("BCD_B200_DEVICES");
		if(pEnv && *pEnv)
		{
			if(std::strcmp(pEnv, "all") == 0)
			{
				m_devices.clear();
				m_allDevices = true;
			}
			else
			{
				std::vector<int> ids;
				for(const char* p = pEnv; *p;)
				{
					char* pEnd = nullptr;
					const long v = std::strtol(p, &pEnd, 10);
					if(pEnd == p)
						break;
					ids.push_back(int(v));
					p = (*pEnd == ',') ? pEnd + 1 : pEnd;
				}
				if(!ids.empty())
					m_devices = ids;
			}
		}
	}

	Denoiser::~Denoiser()
	{
		releaseDeviceResources();
	}

	void Denoiser::releaseDeviceResources()
	{
		if(m_pSession)
			bcd_b200_group_destroy(static_cast<bcd_b200_group*>(m_pSession));
		m_pSession = nullptr;
	}

	void Denoiser::setDevice(int i_device)
	{
		if(m_devices.size() == 1 && m_devices[0] == i_device && !m_allDevices)
			return;
		releaseDeviceResources();
		m_devices.assign(1, i_device);
		m_allDevices = false;
	}

	void Denoiser::setDevices(const int* i_pDeviceIds, int i_nbOfDevices)
	{
		releaseDeviceResources();
		m_devices.clear();
		m_allDevices = i_nbOfDevices <= 0;
		for(int i = 0; i < i_nbOfDevices; ++i)
			m_devices.push_back(i_pDeviceIds ? i_pDeviceIds[i] : i);
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

		// the device session lives as long as the object: handles, buffers, streams, peer links, page-locked ranges
		if(!m_pSession)
		{
			bcd_b200_group* pGroup = nullptr;
			const int rcCreate = bcd_b200_group_create(&pGroup, m_allDevices ? nullptr : m_devices.data(),
					m_allDevices ? 0 : int(m_devices.size()));
			if(rcCreate != BCD_B200_OK)
			{
				std::cerr << "Aborting denoising: " << bcd_b200_last_error() << std::endl;
				return false;
			}
			m_pSession = pGroup;
		}
		bcd_b200_group* pGroup = static_cast<bcd_b200_group*>(m_pSession);
		bcd_b200_group_set_pinning(pGroup, m_pinHostBuffers ? 1 : 0);

		// the output is resized to W x H x 3 and overwritten (Denoiser.cpp:207-208)
		m_outputs.m_pDenoisedColors->resize(images.width, images.height, 3);
		const int rc = bcd_b200_group_denoise(pGroup, &images, &params, m_outputs.m_pDenoisedColors->getDataPtr(),
				&forwardProgress, &m_progressCallback);
		// the reference overwrites m_nbOfCores with the real team size (Denoiser.cpp:121); the GPUs that worked on the
		// frame play that role here
		m_parameters.m_nbOfCores = std::max(1, bcd_b200_group_tiles(pGroup));
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

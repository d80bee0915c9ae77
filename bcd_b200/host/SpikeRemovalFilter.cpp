// bcd::SpikeRemovalFilter for the B200 path: in-place host API of the reference
// (/root/reference/src/core/SpikeRemovalFilter.cpp:18-75) forwarded to the C ABI (include/bcd_b200.h).
#include "SpikeRemovalFilter.h"
#include "bcd_b200.h"

#include <iostream>

namespace bcd
{

	bool SpikeRemovalFilter::s_useIntegerAbsLikeGccBuild = true;
	int SpikeRemovalFilter::s_device = 0;
	bool SpikeRemovalFilter::s_lastCallSucceeded = true;

	void SpikeRemovalFilter::filter(
			DeepImage<float>& io_rInputColorImage,
			DeepImage<float>& io_rInputNbOfSamplesImage,
			DeepImage<float>& io_rInputHistogramImage,
			DeepImage<float>& io_rInputCovImage,
			float i_thresholdStDevFactor)
	{
		s_lastCallSucceeded = false;
		// the four images describe the same pixels: same size, depths 3 / 1 / any / 6 (the C layer reads W x H x depth
		// floats from each pointer, so a smaller histogram or covariance file must be refused here)
		const int w = io_rInputColorImage.getWidth(), h = io_rInputColorImage.getHeight();
		const DeepImage<float>* others[3] = { &io_rInputNbOfSamplesImage, &io_rInputHistogramImage, &io_rInputCovImage };
		const char* names[3] = { "number of samples", "histogram", "covariance" };
		const int depths[3] = { 1, 0, 6 };
		if(w <= 0 || h <= 0 || io_rInputColorImage.getDepth() != 3)
		{
			std::cerr << "SpikeRemovalFilter: the color image must be a non-empty 3-channel image" << std::endl;
			return;
		}
		for(int i = 0; i < 3; ++i)
			if(others[i]->getWidth() != w || others[i]->getHeight() != h || (depths[i] && others[i]->getDepth() != depths[i]) ||
					others[i]->getDepth() <= 0)
			{
				std::cerr << "SpikeRemovalFilter: input " << names[i] << " image is " << others[i]->getWidth() << "x" << others[i]->getHeight()
						<< "x" << others[i]->getDepth() << " but input color image is " << w << "x" << h << "; images left untouched" << std::endl;
				return;
			}
		const int rc = bcd_b200_spike_removal_host(
				io_rInputColorImage.getDataPtr(), io_rInputNbOfSamplesImage.getDataPtr(),
				io_rInputHistogramImage.getDataPtr(), io_rInputCovImage.getDataPtr(),
				io_rInputColorImage.getWidth(), io_rInputColorImage.getHeight(), io_rInputHistogramImage.getDepth(),
				i_thresholdStDevFactor, s_useIntegerAbsLikeGccBuild ? 1 : 0, s_device);
		if(rc != BCD_B200_OK) // the reference's filter cannot fail; report and leave the images untouched
			std::cerr << "SpikeRemovalFilter: " << bcd_b200_last_error() << std::endl;
		else
			s_lastCallSucceeded = true;
	}

} // namespace bcd

// bcd::SpikeRemovalFilter for the B200 path: in-place host API of the reference
// (/root/reference/src/core/SpikeRemovalFilter.cpp:18-75) forwarded to the C ABI (include/bcd_b200.h).
#include "SpikeRemovalFilter.h"
#include "bcd_b200.h"

#include <iostream>

namespace bcd
{

	bool SpikeRemovalFilter::s_useIntegerAbsLikeGccBuild = true;
	int SpikeRemovalFilter::s_device = 0;

	void SpikeRemovalFilter::filter(
			DeepImage<float>& io_rInputColorImage,
			DeepImage<float>& io_rInputNbOfSamplesImage,
			DeepImage<float>& io_rInputHistogramImage,
			DeepImage<float>& io_rInputCovImage,
			float i_thresholdStDevFactor)
	{
		const int rc = bcd_b200_spike_removal_host(
				io_rInputColorImage.getDataPtr(), io_rInputNbOfSamplesImage.getDataPtr(),
				io_rInputHistogramImage.getDataPtr(), io_rInputCovImage.getDataPtr(),
				io_rInputColorImage.getWidth(), io_rInputColorImage.getHeight(), io_rInputHistogramImage.getDepth(),
				i_thresholdStDevFactor, s_useIntegerAbsLikeGccBuild ? 1 : 0, s_device);
		if(rc != BCD_B200_OK) // the reference's filter cannot fail; report and leave the images untouched
			std::cerr << "SpikeRemovalFilter: " << bcd_b200_last_error() << std::endl;
	}

} // namespace bcd

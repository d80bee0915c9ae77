// bcd_b200 — spike-removal prefilter with the reference's public surface
// (superboubek/bcd include/bcd/core/SpikeRemovalFilter.h:20-30; algorithm src/core/SpikeRemovalFilter.cpp:18-116).
// The four images are filtered in place; the work runs on the GPU behind bcd_b200_spike_removal_host().
#ifndef BCD_B200_SPIKE_REMOVAL_FILTER_H
#define BCD_B200_SPIKE_REMOVAL_FILTER_H

#include "DeepImage.h"

namespace bcd
{

	class SpikeRemovalFilter
	{
	public:
		/// A pixel further than i_thresholdStDevFactor standard deviations from the mean of its 3x3 neighbourhood (any
		/// channel) is replaced, in all four images, by the neighbour with the L1-median colour.
		static void filter(
				DeepImage<float>& io_rInputColorImage,
				DeepImage<float>& io_rInputNbOfSamplesImage,
				DeepImage<float>& io_rInputHistogramImage,
				DeepImage<float>& io_rInputCovImage,
				float i_thresholdStDevFactor = 2.f);

		/// bcd_b200 extension. The reference's source calls an unqualified abs() on floats, which binds to ::abs(int)
		/// under GCC/glibc and to the float overload under MSVC; true (default) reproduces the GCC build.
		static bool s_useIntegerAbsLikeGccBuild;
		static int s_device;
		/// bcd_b200 extension: false when the last filter() call left the images untouched (mismatched sizes, no device);
		/// filter() keeps the reference's void signature.
		static bool s_lastCallSucceeded;
	};

} // namespace bcd

#endif // BCD_B200_SPIKE_REMOVAL_FILTER_H

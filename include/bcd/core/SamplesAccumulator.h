// bcd_b200 — renderer-side accumulation of the per-pixel statistics the denoiser consumes, with the reference's
// public surface (superboubek/bcd include/bcd/core/SamplesAccumulator.h:20-80). Host code (stays on the CPU per the
// north star); re-implemented in bcd_b200/host/SamplesAccumulator.cpp and checked bit-for-bit against the reference
// by tests/test_host_api.py.
#ifndef BCD_B200_SAMPLES_ACCUMULATOR_H
#define BCD_B200_SAMPLES_ACCUMULATOR_H

#include "DeepImage.h"

namespace bcd
{

	struct HistogramParameters
	{
		int m_nbOfBins = 20;     ///< per colour channel
		float m_gamma = 2.2f;    ///< bins are uniform in value^(1/gamma)
		float m_maxValue = 2.5f; ///< value mapped to the last regular bin boundary
	};

	struct SamplesStatisticsImages
	{
		SamplesStatisticsImages() = default;
		SamplesStatisticsImages(int width, int height, int nbOfBins);

		DeepImage<float> m_nbOfSamplesImage;
		DeepImage<float> m_meanImage;
		DeepImage<float> m_covarImage;
		DeepImage<float> m_histoImage;
	};

	class SamplesAccumulator
	{
	public:
		SamplesAccumulator(int width, int height, const HistogramParameters& histo);

		void addSample(int line, int column, float r, float g, float b, float weight = 1.f);

		/// Finalised copy of the statistics (the accumulator stays usable).
		SamplesStatisticsImages getSamplesStatistics() const;
		/// Finalises in place and moves the images out; the accumulator must not be used afterwards.
		SamplesStatisticsImages extractSamplesStatistics();

	private:
		void finalise(SamplesStatisticsImages& stats) const;

	private:
		int m_width;
		int m_height;
		HistogramParameters m_histo;
		SamplesStatisticsImages m_stats;
		DeepImage<float> m_sumW2;
		bool m_isValid;
	};

} // namespace bcd

#endif // BCD_B200_SAMPLES_ACCUMULATOR_H

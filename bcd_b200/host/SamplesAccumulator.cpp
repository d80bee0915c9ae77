// bcd::SamplesAccumulator re-implemented for this repository (host code). Follows the arithmetic
// of the reference so that renderers produce bit-identical denoiser inputs:
//   addSample              /root/reference/src/core/SamplesAccumulator.cpp:44-107  (weighted sums, soft-binned histogram)
//   computeSampleStatistics /root/reference/src/core/SamplesAccumulator.cpp:110-141 (mean, covariance, bias correction)
#include "SamplesAccumulator.h"
#include "CovarianceMatrix.h"

#include <cmath>

namespace bcd
{

	SamplesStatisticsImages::SamplesStatisticsImages(int i_width, int i_height, int i_nbOfBins) :
			m_nbOfSamplesImage(i_width, i_height, 1),
			m_meanImage(i_width, i_height, 3),
			m_covarImage(i_width, i_height, 6),
			m_histoImage(i_width, i_height, 3 * i_nbOfBins)
	{
	}

	SamplesAccumulator::SamplesAccumulator(int i_width, int i_height, const HistogramParameters& i_rHistogramParameters) :
			m_width(i_width), m_height(i_height),
			m_histogramParameters(i_rHistogramParameters),
			m_samplesStatisticsImages(i_width, i_height, i_rHistogramParameters.m_nbOfBins),
			m_squaredWeightSumsImage(i_width, i_height, 1),
			m_isValid(true)
	{
		m_samplesStatisticsImages.m_nbOfSamplesImage.fill(0.f);
		m_samplesStatisticsImages.m_meanImage.fill(0.f);
		m_samplesStatisticsImages.m_covarImage.fill(0.f);
		m_samplesStatisticsImages.m_histoImage.fill(0.f);
		m_squaredWeightSumsImage.fill(0.f);
	}

	void SamplesAccumulator::addSample(int i_line, int i_column, float i_sampleR, float i_sampleG, float i_sampleB, float i_weight)
	{
		const float rgb[3] = { i_sampleR, i_sampleG, i_sampleB };
		const int nbOfBins = m_histogramParameters.m_nbOfBins;
		const float saturation = 2.f; // normalised value mapped to the very last bin

		m_samplesStatisticsImages.m_nbOfSamplesImage.get(i_line, i_column, 0) += i_weight;
		m_squaredWeightSumsImage.get(i_line, i_column, 0) += i_weight * i_weight;

		float* pMean = &m_samplesStatisticsImages.m_meanImage.get(i_line, i_column, 0);
		for(int c = 0; c < 3; ++c)
			pMean[c] += i_weight * rgb[c];

		// second moments in the xx,yy,zz,yz,xz,xy order; products associate left to right like the reference's
		float* pCov = &m_samplesStatisticsImages.m_covarImage.get(i_line, i_column, 0);
		pCov[int(ESymMatData::e_xx)] += i_weight * i_sampleR * i_sampleR;
		pCov[int(ESymMatData::e_yy)] += i_weight * i_sampleG * i_sampleG;
		pCov[int(ESymMatData::e_zz)] += i_weight * i_sampleB * i_sampleB;
		pCov[int(ESymMatData::e_yz)] += i_weight * i_sampleG * i_sampleB;
		pCov[int(ESymMatData::e_xz)] += i_weight * i_sampleR * i_sampleB;
		pCov[int(ESymMatData::e_xy)] += i_weight * i_sampleR * i_sampleG;

		float* pHisto = &m_samplesStatisticsImages.m_histoImage.get(i_line, i_column, 0);
		for(int c = 0; c < 3; ++c)
		{
			float v = rgb[c] > 0 ? rgb[c] : 0;
			if(m_histogramParameters.m_gamma > 1)
				v = std::pow(v, 1.f / m_histogramParameters.m_gamma);
			if(m_histogramParameters.m_maxValue > 0)
				v = v / m_histogramParameters.m_maxValue;
			if(v > saturation)
				v = saturation;

			const float position = v * (nbOfBins - 2);
			int lowBin = int(position);
			float highWeight;
			if(lowBin < nbOfBins - 2)
				highWeight = position - lowBin; // linear split between two regular bins
			else
			{
				lowBin = nbOfBins - 2;            // v >= 1: split between the last two bins
				highWeight = (v - 1.0f) / (saturation - 1.f);
			}
			const float lowWeight = 1.0f - highWeight;
			pHisto[c * nbOfBins + lowBin] += i_weight * lowWeight;
			pHisto[c * nbOfBins + lowBin + 1] += i_weight * highWeight;
		}
	}

	void SamplesAccumulator::finalise(SamplesStatisticsImages& io_rStats) const
	{
		const int covIndex[6][2] = { {0, 0}, {1, 1}, {2, 2}, {1, 2}, {0, 2}, {0, 1} }; // xx,yy,zz,yz,xz,xy
		for(int line = 0; line < m_height; ++line)
			for(int column = 0; column < m_width; ++column)
			{
				const float weightSum = io_rStats.m_nbOfSamplesImage.get(line, column, 0);
				const float squaredWeightSum = m_squaredWeightSumsImage.get(line, column, 0);
				const float invWeightSum = 1.f / weightSum;
				float mean[3];
				for(int c = 0; c < 3; ++c)
				{
					mean[c] = invWeightSum * io_rStats.m_meanImage.get(line, column, c);
					io_rStats.m_meanImage.set(line, column, c, mean[c]);
				}
				const float unbias = 1.f / (1 - squaredWeightSum / (weightSum * weightSum));
				for(int e = 0; e < 6; ++e)
				{
					float cov = io_rStats.m_covarImage.get(line, column, e) * invWeightSum;
					cov -= mean[covIndex[e][0]] * mean[covIndex[e][1]];
					io_rStats.m_covarImage.set(line, column, e, cov * unbias);
				}
			}
	}

	SamplesStatisticsImages SamplesAccumulator::getSamplesStatistics() const
	{
		SamplesStatisticsImages stats(m_samplesStatisticsImages);
		finalise(stats);
		return stats;
	}

	SamplesStatisticsImages SamplesAccumulator::extractSamplesStatistics()
	{
		finalise(m_samplesStatisticsImages);
		m_isValid = false;
		return std::move(m_samplesStatisticsImages);
	}

} // namespace bcd

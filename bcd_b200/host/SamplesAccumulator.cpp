// bcd::SamplesAccumulator re-implemented for this repository (host code). Follows the arithmetic
// of the reference so that renderers produce bit-identical denoiser inputs:
//   addSample              /root/reference/src/core/SamplesAccumulator.cpp:44-107  (weighted sums, soft-binned histogram)
//   computeSampleStatistics /root/reference/src/core/SamplesAccumulator.cpp:110-141 (mean, covariance, bias correction)
#include "SamplesAccumulator.h"
#include "CovarianceMatrix.h"

#include <cmath>

namespace bcd
{

	SamplesStatisticsImages::SamplesStatisticsImages(int width, int height, int nbOfBins) :
			m_nbOfSamplesImage(width, height, 1),
			m_meanImage(width, height, 3),
			m_covarImage(width, height, 6),
			m_histoImage(width, height, 3 * nbOfBins)
	{
	}

	SamplesAccumulator::SamplesAccumulator(int width, int height, const HistogramParameters& histo) :
			m_width(width), m_height(height),
			m_histo(histo),
			m_stats(width, height, histo.m_nbOfBins),
			m_sumW2(width, height, 1),
			m_isValid(true)
	{
		m_stats.m_nbOfSamplesImage.fill(0.f);
		m_stats.m_meanImage.fill(0.f);
		m_stats.m_covarImage.fill(0.f);
		m_stats.m_histoImage.fill(0.f);
		m_sumW2.fill(0.f);
	}

	void SamplesAccumulator::addSample(int line, int column, float r, float g, float b, float weight)
	{
		const float rgb[3] = { r, g, b };
		const int nBins = m_histo.m_nbOfBins;
		const float saturation = 2.f; // normalised value mapped to the very last bin

		m_stats.m_nbOfSamplesImage.get(line, column, 0) += weight;
		m_sumW2.get(line, column, 0) += weight * weight;

		float* pMean = &m_stats.m_meanImage.get(line, column, 0);
		for(int c = 0; c < 3; ++c)
			pMean[c] += weight * rgb[c];

		// second moments in the xx,yy,zz,yz,xz,xy order; products associate left to right like the reference's
		float* pCov = &m_stats.m_covarImage.get(line, column, 0);
		pCov[int(ESymMatData::e_xx)] += weight * r * r;
		pCov[int(ESymMatData::e_yy)] += weight * g * g;
		pCov[int(ESymMatData::e_zz)] += weight * b * b;
		pCov[int(ESymMatData::e_yz)] += weight * g * b;
		pCov[int(ESymMatData::e_xz)] += weight * r * b;
		pCov[int(ESymMatData::e_xy)] += weight * r * g;

		float* pHisto = &m_stats.m_histoImage.get(line, column, 0);
		for(int c = 0; c < 3; ++c)
		{
			float v = rgb[c] > 0 ? rgb[c] : 0;
			if(m_histo.m_gamma > 1)
				v = std::pow(v, 1.f / m_histo.m_gamma);
			if(m_histo.m_maxValue > 0)
				v = v / m_histo.m_maxValue;
			if(v > saturation)
				v = saturation;

			const float position = v * (nBins - 2);
			int lowBin = int(position);
			float highWeight;
			if(lowBin < nBins - 2)
				highWeight = position - lowBin; // linear split between two regular bins
			else
			{
				lowBin = nBins - 2;            // v >= 1: split between the last two bins
				highWeight = (v - 1.0f) / (saturation - 1.f);
			}
			const float lowWeight = 1.0f - highWeight;
			pHisto[c * nBins + lowBin] += weight * lowWeight;
			pHisto[c * nBins + lowBin + 1] += weight * highWeight;
		}
	}

	void SamplesAccumulator::finalise(SamplesStatisticsImages& stats) const
	{
		const int covIndex[6][2] = { {0, 0}, {1, 1}, {2, 2}, {1, 2}, {0, 2}, {0, 1} }; // xx,yy,zz,yz,xz,xy
		for(int line = 0; line < m_height; ++line)
			for(int column = 0; column < m_width; ++column)
			{
				const float weightSum = stats.m_nbOfSamplesImage.get(line, column, 0);
				const float squaredWeightSum = m_sumW2.get(line, column, 0);
				const float invWeightSum = 1.f / weightSum;
				float mean[3];
				for(int c = 0; c < 3; ++c)
				{
					mean[c] = invWeightSum * stats.m_meanImage.get(line, column, c);
					stats.m_meanImage.set(line, column, c, mean[c]);
				}
				const float unbias = 1.f / (1 - squaredWeightSum / (weightSum * weightSum));
				for(int e = 0; e < 6; ++e)
				{
					float cov = stats.m_covarImage.get(line, column, e) * invWeightSum;
					cov -= mean[covIndex[e][0]] * mean[covIndex[e][1]];
					stats.m_covarImage.set(line, column, e, cov * unbias);
				}
			}
	}

	SamplesStatisticsImages SamplesAccumulator::getSamplesStatistics() const
	{
		SamplesStatisticsImages copy(m_stats);
		finalise(copy);
		return copy;
	}

	SamplesStatisticsImages SamplesAccumulator::extractSamplesStatistics()
	{
		finalise(m_stats);
		m_isValid = false;
		return std::move(m_stats);
	}

} // namespace bcd

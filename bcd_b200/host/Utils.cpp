// bcd::Utils, re-implemented (behaviour of /root/reference/src/core/Utils.cpp:21-113).
#include "Utils.h"

#include <algorithm>
#include <cstring>

namespace bcd
{

	std::string Utils::extractFolderPath(const std::string& filePath)
	{
		const size_t slash = filePath.rfind('/');
		return slash == std::string::npos ? std::string() : filePath.substr(0, slash + 1);
	}

	std::string Utils::getRelativePathFromFolder(const std::string& filePath, const std::string& folderPath)
	{
		// longest common prefix that ends at a separator, then one "../" per folder level left below it
		size_t common = 0;
		for(size_t i = 0, n = std::min(filePath.size(), folderPath.size()); i < n && filePath[i] == folderPath[i]; ++i)
			if(filePath[i] == '/')
				common = i + 1;
		std::string rel;
		for(size_t i = common; i < folderPath.size(); ++i)
			if(folderPath[i] == '/')
				rel += "../";
		return rel + filePath.substr(common);
	}

	DeepImage<float> Utils::mergeHistogramAndNbOfSamples(const DeepImage<float>& histogram, const DeepImage<float>& nbOfSamples)
	{
		const int depth = histogram.getDepth();
		DeepImage<float> merged(histogram.getWidth(), histogram.getHeight(), depth + 1);
		const float* src = histogram.getDataPtr();
		const float* count = nbOfSamples.getDataPtr();
		float* dst = merged.getDataPtr();
		const long np = (long)histogram.getWidth() * histogram.getHeight();
#pragma omp parallel for schedule(static)
		for(long p = 0; p < np; ++p)
		{
			float* o = dst + (size_t)p * (depth + 1);
			if(depth) memcpy(o, src + (size_t)p * depth, (size_t)depth * sizeof(float));
			o[depth] = count[p];
		}
		return merged;
	}

	bool Utils::separateNbOfSamplesFromHistogram(DeepImage<float>& histogramOut, DeepImage<float>& nbOfSamplesOut,
			const DeepImage<float>& histogramAndNbOfSamples)
	{
		const int width = histogramAndNbOfSamples.getWidth(), height = histogramAndNbOfSamples.getHeight();
		const int depth = histogramAndNbOfSamples.getDepth() - 1;
		if(depth < 0)
			return false;
		histogramOut.resize(width, height, depth);
		nbOfSamplesOut.resize(width, height, 1);
		const float* src = histogramAndNbOfSamples.getDataPtr();
		float* hist = histogramOut.getDataPtr();
		float* count = nbOfSamplesOut.getDataPtr();
		const long np = (long)width * height;
#pragma omp parallel for schedule(static)
		for(long p = 0; p < np; ++p)
		{
			const float* in = src + (size_t)p * (depth + 1);
			if(depth) memcpy(hist + (size_t)p * depth, in, (size_t)depth * sizeof(float));
			count[p] = in[depth];
		}
		return true;
	}

} // namespace bcd

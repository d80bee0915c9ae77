// bcd::Utils (reference: /root/reference/src/core/Utils.cpp:21-113), re-implemented.
#include "Utils.h"

#include <cstring>

namespace bcd
{

	bool Utils::separateNbOfSamplesFromHistogram(DeepImage<float>& o_rHistoImage, DeepImage<float>& o_rNbOfSamplesImage,
			const DeepImage<float>& i_rHistoAndNbOfSamplesImage)
	{
		const int w = i_rHistoAndNbOfSamplesImage.getWidth(), h = i_rHistoAndNbOfSamplesImage.getHeight();
		const int d = i_rHistoAndNbOfSamplesImage.getDepth() - 1;
		if(d < 0)
			return false;
		o_rHistoImage.resize(w, h, d);
		o_rNbOfSamplesImage.resize(w, h, 1);
		const float* src = i_rHistoAndNbOfSamplesImage.getDataPtr();
		float* hist = o_rHistoImage.getDataPtr();
		float* ns = o_rNbOfSamplesImage.getDataPtr();
		const size_t np = (size_t)w * h;
		for(size_t p = 0; p < np; ++p, src += d + 1)
		{
			if(d) memcpy(hist + p * d, src, (size_t)d * sizeof(float));
			ns[p] = src[d];
		}
		return true;
	}

	DeepImage<float> Utils::mergeHistogramAndNbOfSamples(const DeepImage<float>& i_rHistoImage, const DeepImage<float>& i_rNbOfSamplesImage)
	{
		const int w = i_rHistoImage.getWidth(), h = i_rHistoImage.getHeight(), d = i_rHistoImage.getDepth();
		DeepImage<float> merged(w, h, d + 1);
		const float* hist = i_rHistoImage.getDataPtr();
		const float* ns = i_rNbOfSamplesImage.getDataPtr();
		float* dst = merged.getDataPtr();
		const size_t np = (size_t)w * h;
		for(size_t p = 0; p < np; ++p, dst += d + 1)
		{
			if(d) memcpy(dst, hist + p * d, (size_t)d * sizeof(float));
			dst[d] = ns[p];
		}
		return merged;
	}

	std::string Utils::extractFolderPath(const std::string& i_rFilePath)
	{
		const size_t pos = i_rFilePath.rfind('/');
		return pos == std::string::npos ? std::string() : i_rFilePath.substr(0, pos + 1);
	}

	std::string Utils::getRelativePathFromFolder(const std::string& i_rFileAbsolutePath, const std::string& i_rFolderAbsolutePath)
	{
		// longest common prefix that ends at a separator, then one "../" per remaining folder level
		const size_t n = std::min(i_rFileAbsolutePath.size(), i_rFolderAbsolutePath.size());
		size_t common = 0;
		for(size_t i = 0; i < n && i_rFileAbsolutePath[i] == i_rFolderAbsolutePath[i]; ++i)
			if(i_rFileAbsolutePath[i] == '/')
				common = i + 1;
		std::string rel;
		for(size_t i = common; i < i_rFolderAbsolutePath.size(); ++i)
			if(i_rFolderAbsolutePath[i] == '/')
				rel += "../";
		return rel + i_rFileAbsolutePath.substr(common);
	}

} // namespace bcd

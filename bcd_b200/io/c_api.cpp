// C entry points over bcd::ImageIO / bcd::ParametersIO for the Python tests (include/bcd_b200_host.h).
#include "bcd_b200_host.h"

#include "DeepImage.h"
#include "ImageIO.h"
#include "ParametersIO.h"
#include "exr_codec.h"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace bcd;

namespace
{
	thread_local DeepImage<float> t_image; // image of the last successful bcd_host_exr_load of this thread

	void copyOut(const std::string& s, char* dst, int cap)
	{
		if(!dst || cap <= 0) return;
		const size_t n = s.size() < (size_t)cap - 1 ? s.size() : (size_t)cap - 1;
		memcpy(dst, s.data(), n);
		dst[n] = 0;
	}
}

extern "C" {

int bcd_host_exr_load(const char* path, int multi_channel, int* width, int* height, int* depth)
{
	DeepImage<float> img;
	const bool ok = multi_channel ? ImageIO::loadMultiChannelsEXR(img, path) : ImageIO::loadEXR(img, path);
	if(!ok) return 0;
	t_image = img;
	if(width) *width = img.getWidth();
	if(height) *height = img.getHeight();
	if(depth) *depth = img.getDepth();
	return 1;
}

int bcd_host_exr_fetch(float* out_hwc)
{
	const size_t n = (size_t)t_image.getWidth() * t_image.getHeight() * t_image.getDepth();
	if(!out_hwc || n == 0) return 0;
	memcpy(out_hwc, t_image.getDataPtr(), n * sizeof(float));
	return 1;
}

int bcd_host_exr_write(const char* path, int multi_channel, const float* hwc, int width, int height, int depth)
{
	if(!hwc || width <= 0 || height <= 0 || depth <= 0) return 0;
	DeepImage<float> img(width, height, depth);
	memcpy(img.getDataPtr(), hwc, (size_t)width * height * depth * sizeof(float));
	return (multi_channel ? ImageIO::writeMultiChannelsEXR(img, path) : ImageIO::writeEXR(img, path)) ? 1 : 0;
}

const char* bcd_host_exr_last_error(void) { return ImageIO::lastError().c_str(); }

int bcd_host_exr_write_ex(const char* path, int multi_channel, const float* hwc, int width, int height, int depth,
		int pixel_type, int compression)
{
	if(!hwc || width <= 0 || height <= 0 || depth <= 0 || (!multi_channel && depth != 3)) return 0;
	const size_t np = (size_t)width * height;
	std::vector<float> planes((size_t)depth * np);
	for(size_t i = 0; i < np; ++i)
		for(int z = 0; z < depth; ++z)
			planes[z * np + i] = hwc[i * depth + z];
	std::vector<std::string> names;
	std::vector<const float*> ptrs;
	for(int z = 0; z < depth; ++z)
	{
		char buf[32];
		if(multi_channel) snprintf(buf, sizeof buf, "Bin_%04d", z);
		else snprintf(buf, sizeof buf, "%c", "RGB"[z]);
		names.push_back(buf);
		ptrs.push_back(planes.data() + z * np);
	}
	std::string err;
	return bcdio::exrWrite(path, width, height, names, ptrs, (bcdio::ExrPixelType)pixel_type, (bcdio::ExrCompression)compression, err) ? 1 : 0;
}

int bcd_host_params_load(const char* path, double* v, char* color_file, char* hist_file, char* cov_file, int cap)
{
	PipelineParameters p;
	// the caller's values are the defaults a missing key keeps
	p.m_prefilteringParameters.m_performSpikeRemoval = v[0] != 0.0;
	p.m_prefilteringParameters.m_spikeRemovalThresholdStDevFactor = (float)v[1];
	p.m_denoiserParameters.m_nbOfScales = (int)v[2];
	DenoiserParameters& d = p.m_denoiserParameters.m_monoscaleParameters;
	d.m_histogramDistanceThreshold = (float)v[3];
	d.m_useCuda = v[4] != 0.0;
	d.m_nbOfCores = (int)v[5];
	d.m_patchRadius = (int)v[6];
	d.m_searchWindowRadius = (int)v[7];
	d.m_useRandomPixelOrder = v[8] != 0.0;
	d.m_markedPixelsSkippingProbability = (float)v[9];
	d.m_minEigenValue = (float)v[10];
	if(!ParametersIO::load(p, path)) return 0;
	v[0] = p.m_prefilteringParameters.m_performSpikeRemoval;
	v[1] = p.m_prefilteringParameters.m_spikeRemovalThresholdStDevFactor;
	v[2] = p.m_denoiserParameters.m_nbOfScales;
	v[3] = d.m_histogramDistanceThreshold;
	v[4] = d.m_useCuda;
	v[5] = d.m_nbOfCores;
	v[6] = d.m_patchRadius;
	v[7] = d.m_searchWindowRadius;
	v[8] = d.m_useRandomPixelOrder;
	v[9] = d.m_markedPixelsSkippingProbability;
	v[10] = d.m_minEigenValue;
	copyOut(p.m_inputFileNames.m_colors, color_file, cap);
	copyOut(p.m_inputFileNames.m_histograms, hist_file, cap);
	copyOut(p.m_inputFileNames.m_covariances, cov_file, cap);
	return 1;
}

int bcd_host_params_write(const char* path, const double* v, const char* color_file, const char* hist_file, const char* cov_file)
{
	PipelineParameters p;
	p.m_prefilteringParameters.m_performSpikeRemoval = v[0] != 0.0;
	p.m_prefilteringParameters.m_spikeRemovalThresholdStDevFactor = (float)v[1];
	p.m_denoiserParameters.m_nbOfScales = (int)v[2];
	DenoiserParameters& d = p.m_denoiserParameters.m_monoscaleParameters;
	d.m_histogramDistanceThreshold = (float)v[3];
	d.m_useCuda = v[4] != 0.0;
	d.m_nbOfCores = (int)v[5];
	d.m_patchRadius = (int)v[6];
	d.m_searchWindowRadius = (int)v[7];
	d.m_useRandomPixelOrder = v[8] != 0.0;
	d.m_markedPixelsSkippingProbability = (float)v[9];
	d.m_minEigenValue = (float)v[10];
	p.m_inputFileNames.m_colors = color_file ? color_file : "";
	p.m_inputFileNames.m_histograms = hist_file ? hist_file : "";
	p.m_inputFileNames.m_covariances = cov_file ? cov_file : "";
	return ParametersIO::write(p, path) ? 1 : 0;
}

}

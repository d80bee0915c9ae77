// bcd::ImageIO over the from-scratch EXR codec (reference behaviour: /root/reference/src/io/ImageIO.cpp:22-141,
// /root/reference/src/io/exr/io_exr.cpp:45-373). Unlike the reference's wrappers nothing here calls exit().
#include "ImageIO.h"
#include "DeepImage.h"
#include "exr_codec.h"

#include <cstdio>
#include <iostream>
#include <vector>

namespace bcd
{

	namespace
	{
		thread_local std::string t_lastError;

		bool failWith(const std::string& msg, const char* path)
		{
			t_lastError = msg;
			std::cerr << "error :: '" << path << "': " << msg << std::endl;
			return false;
		}

		std::string binName(int i)
		{
			char buf[32];
			snprintf(buf, sizeof buf, "Bin_%04d", i);
			return buf;
		}
	}

	const std::string& ImageIO::lastError() { return t_lastError; }

	bool ImageIO::loadEXR(DeepImage<float>& o_rImage, const char* i_pFilePath)
	{
		std::cout << "Loading " << i_pFilePath << std::endl;
		bcdio::ExrFile f;
		std::string err;
		if(!bcdio::exrRead(i_pFilePath, f, err))
			return failWith(err, i_pFilePath);
		const size_t np = (size_t)f.width * f.height;
		const std::vector<float> zeros(np, 0.f);
		const float* rgb[3];
		const char* names[3] = {"R", "G", "B"};
		for(int z = 0; z < 3; ++z)
		{
			const int c = f.find(names[z]);
			rgb[z] = c >= 0 ? f.planes[c].data() : zeros.data();
		}
		bool grey = true; // a colour file whose three channels agree everywhere is a 1-channel image (ImageIO.cpp:41-52)
		for(size_t i = 0; i < np && grey; ++i)
			grey = rgb[0][i] == rgb[1][i] && rgb[0][i] == rgb[2][i];
		const int depth = grey ? 1 : 3;
		o_rImage.resize(f.width, f.height, depth);
		float* dst = o_rImage.getDataPtr();
		for(size_t i = 0; i < np; ++i)
			for(int z = 0; z < depth; ++z)
				dst[i * depth + z] = rgb[z][i];
		t_lastError.clear();
		return true;
	}

	bool ImageIO::loadMultiChannelsEXR(DeepImage<float>& o_rImage, const char* i_pFilePath)
	{
		std::cout << "Loading " << i_pFilePath << std::endl;
		bcdio::ExrFile f;
		std::string err;
		if(!bcdio::exrRead(i_pFilePath, f, err))
			return failWith(err, i_pFilePath);
		std::vector<int> chan;
		for(int i = 0;; ++i)
		{
			const int c = f.find(binName(i));
			if(c < 0) break;
			chan.push_back(c);
		}
		if(chan.empty())
			return failWith("no Bin_0000 channel", i_pFilePath);
		const int depth = (int)chan.size();
		const size_t np = (size_t)f.width * f.height;
		o_rImage.resize(f.width, f.height, depth);
		float* dst = o_rImage.getDataPtr();
		for(int z = 0; z < depth; ++z)
		{
			const float* src = f.planes[chan[z]].data();
			for(size_t i = 0; i < np; ++i)
				dst[i * depth + z] = src[i];
		}
		t_lastError.clear();
		return true;
	}

	bool ImageIO::writeEXR(const DeepImage<float>& i_rImage, const char* i_pFilePath)
	{
		const int w = i_rImage.getWidth(), h = i_rImage.getHeight(), d = i_rImage.getDepth();
		if(d != 1 && d < 3)
			return failWith("writeEXR needs 1 or 3 channels", i_pFilePath);
		const size_t np = (size_t)w * h;
		std::vector<float> planes(4 * np);
		const float* src = i_rImage.getDataPtr();
		for(size_t i = 0; i < np; ++i)
		{
			for(int z = 0; z < 3; ++z)
				planes[z * np + i] = src[i * d + (d == 1 ? 0 : z)];
			planes[3 * np + i] = 1.f;
		}
		std::vector<std::string> names = {"R", "G", "B", "A"};
		std::vector<const float*> ptrs = {planes.data(), planes.data() + np, planes.data() + 2 * np, planes.data() + 3 * np};
		std::string err;
		if(!bcdio::exrWrite(i_pFilePath, w, h, names, ptrs, bcdio::EXR_HALF, bcdio::EXR_PIZ, err)) // RgbaOutputFile's default
			return failWith(err, i_pFilePath);
		t_lastError.clear();
		return true;
	}

	bool ImageIO::writeMultiChannelsEXR(const DeepImage<float>& i_rImage, const char* i_pFilePath)
	{
		const int w = i_rImage.getWidth(), h = i_rImage.getHeight(), d = i_rImage.getDepth();
		if(d < 1 || d > 10000)
			return failWith("writeMultiChannelsEXR needs 1..10000 channels", i_pFilePath);
		const size_t np = (size_t)w * h;
		std::vector<float> planes((size_t)d * np);
		const float* src = i_rImage.getDataPtr();
		for(size_t i = 0; i < np; ++i)
			for(int z = 0; z < d; ++z)
				planes[z * np + i] = src[i * d + z];
		std::vector<std::string> names;
		std::vector<const float*> ptrs;
		for(int z = 0; z < d; ++z)
		{
			names.push_back(binName(z));
			ptrs.push_back(planes.data() + z * np);
		}
		std::string err;
		if(!bcdio::exrWrite(i_pFilePath, w, h, names, ptrs, bcdio::EXR_FLOAT, bcdio::EXR_ZIP, err))
			return failWith(err, i_pFilePath);
		t_lastError.clear();
		return true;
	}

} // namespace bcd

// bcd::ImageIO over the from-scratch EXR codec (reference behaviour: /root/reference/src/io/ImageIO.cpp:22-141,
// /root/reference/src/io/exr/io_exr.cpp:45-373). Unlike the reference's wrappers nothing here calls exit().
#include "ImageIO.h"
#include "DeepImage.h"
#include "exr_codec.h"

#include <algorithm>
#include <cstdio>
#include <exception>
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

	static bool loadEXRBody(DeepImage<float>& o_rImage, const char* i_pFilePath);
	static bool loadMultiChannelsEXRBody(DeepImage<float>& o_rImage, const char* i_pFilePath);

	// The loaders size their images from the file's header; whatever goes wrong on the way — including running out of
	// memory on a hostile header — comes back as `false` + lastError(), never as an exception through the host renderer.
	bool ImageIO::loadEXR(DeepImage<float>& o_rImage, const char* i_pFilePath)
	{
		try
		{
			return loadEXRBody(o_rImage, i_pFilePath);
		}
		catch(const std::exception& e)
		{
			o_rImage.clearAndFreeMemory();
			return failWith(std::string("cannot load the image: ") + e.what(), i_pFilePath);
		}
	}

	bool ImageIO::loadMultiChannelsEXR(DeepImage<float>& o_rImage, const char* i_pFilePath)
	{
		try
		{
			return loadMultiChannelsEXRBody(o_rImage, i_pFilePath);
		}
		catch(const std::exception& e)
		{
			o_rImage.clearAndFreeMemory();
			return failWith(std::string("cannot load the image: ") + e.what(), i_pFilePath);
		}
	}

	static bool loadEXRBody(DeepImage<float>& o_rImage, const char* i_pFilePath)
	{
		std::cout << "Loading " << i_pFilePath << std::endl;
		bcdio::ExrReader rd;
		std::string err;
		if(!bcdio::exrOpen(i_pFilePath, rd, err))
			return failWith(err, i_pFilePath);
		const int w = rd.info.width, h = rd.info.height;
		const size_t np = (size_t)w * h;
		// R, G, B straight into interleaved pixels; a channel the file does not have stays 0
		std::vector<float> rgb(np * 3, 0.f);
		const std::vector<int> channelOfSlot = {rd.info.find("R"), rd.info.find("G"), rd.info.find("B")};
		if(!bcdio::exrDecodeInterleaved(rd, channelOfSlot, rgb.data(), err))
			return failWith(err, i_pFilePath);
		bool grey = true; // a colour file whose three channels agree everywhere is a 1-channel image (ImageIO.cpp:41-52)
		for(size_t i = 0; i < np && grey; ++i)
			grey = rgb[3 * i] == rgb[3 * i + 1] && rgb[3 * i] == rgb[3 * i + 2];
		const int depth = grey ? 1 : 3;
		o_rImage.resize(w, h, depth);
		float* dst = o_rImage.getDataPtr();
		if(grey)
			for(size_t i = 0; i < np; ++i)
				dst[i] = rgb[3 * i];
		else
			std::copy(rgb.begin(), rgb.end(), dst);
		t_lastError.clear();
		return true;
	}

	static bool loadMultiChannelsEXRBody(DeepImage<float>& o_rImage, const char* i_pFilePath)
	{
		std::cout << "Loading " << i_pFilePath << std::endl;
		bcdio::ExrReader rd;
		std::string err;
		if(!bcdio::exrOpen(i_pFilePath, rd, err))
			return failWith(err, i_pFilePath);
		std::vector<int> channelOfSlot; // Bin_0000, Bin_0001, ... as long as they are there
		for(int i = 0;; ++i)
		{
			const int c = rd.info.find(binName(i));
			if(c < 0) break;
			channelOfSlot.push_back(c);
		}
		if(channelOfSlot.empty())
			return failWith("no Bin_0000 channel", i_pFilePath);
		o_rImage.resize(rd.info.width, rd.info.height, (int)channelOfSlot.size());
		if(!bcdio::exrDecodeInterleaved(rd, channelOfSlot, o_rImage.getDataPtr(), err))
		{
			o_rImage.clearAndFreeMemory();
			return failWith(err, i_pFilePath);
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
#pragma omp parallel for schedule(static)
		for(long l = 0; l < (long)h; ++l)
			for(size_t i = (size_t)l * w; i < (size_t)(l + 1) * w; ++i)
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

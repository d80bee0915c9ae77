// TEST INFRASTRUCTURE — C entry points over the reference's own EXR layer, compiled in place by oracle/Makefile into
// oracle/_ref/libexr_ref.so: /root/reference/src/io/exr/io_exr.cpp on top of the vendored OpenEXR 2.2 library
// (/root/reference/ext/openexr). Used by tests/test_exr_io.py to pin bcd_b200/io/exr_codec.cpp: files written by one
// side must read back identically through the other. Nothing of the product links against it.
#include "io_exr.h"

#include <ImfChannelList.h>
#include <ImfFrameBuffer.h>
#include <ImfHeader.h>
#include <ImfOutputFile.h>
#include <half.h>

#include <cstdio>
#include <cstdlib>
#include <exception>
#include <string>
#include <vector>

extern "C" {

// the reference's four wrappers, verbatim behaviour (planar float buffers)
float* exrref_read_rgb(const char* path, int* w, int* h) { return readImageEXR(path, w, h); }
void exrref_free_rgb(float* p) { delete[] p; }
float* exrref_read_multi(const char* path, int* w, int* h, int* n) { return readMultiImageEXR(path, w, h, n); }
void exrref_free_multi(float* p) { free(p); }
void exrref_write_rgb(const char* path, const float* planar, int w, int h) { writeImageEXR(path, planar, w, h); }
void exrref_write_multi(const char* path, const float* planar, int w, int h, int n) { writeMultiImageEXR(path, planar, w, h, n); }

// generic writer through the OpenEXR API, for the variants the wrappers never produce: any compression
// (Imf::Compression value), HALF (1) / FLOAT (2) / UINT (0) channels, a data window that does not start at (0, 0),
// decreasing line order. rgb_names != 0: channels R, G, B (nchan = 3); else Bin_%04d. Returns 0, or 1 on an exception.
int exrref_write_generic(const char* path, const float* planar, int w, int h, int nchan, int rgb_names, int pixel_type,
                         int compression, int min_x, int min_y, int decreasing_y)
{
	try
	{
		using namespace Imf;
		using namespace Imath;
		const Box2i dataWindow(V2i(min_x, min_y), V2i(min_x + w - 1, min_y + h - 1));
		const Box2i displayWindow(V2i(0, 0), V2i(min_x + w - 1, min_y + h - 1));
		Header header(displayWindow, dataWindow, 1.f, V2f(0, 0), 1.f, decreasing_y ? DECREASING_Y : INCREASING_Y,
		              (Compression)compression);
		const size_t np = (size_t)w * h;
		std::vector<std::vector<half> > halves;
		std::vector<std::vector<unsigned int> > uints;
		FrameBuffer fb;
		for(int c = 0; c < nchan; ++c)
		{
			char name[16];
			if(rgb_names) snprintf(name, sizeof name, "%c", "RGB"[c]);
			else snprintf(name, sizeof name, "Bin_%04d", c);
			header.channels().insert(name, Channel((PixelType)pixel_type));
			const float* src = planar + (size_t)c * np;
			char* base;
			size_t px;
			if(pixel_type == HALF)
			{
				halves.push_back(std::vector<half>(np));
				for(size_t i = 0; i < np; ++i) halves.back()[i] = half(src[i]);
				base = (char*)halves.back().data();
				px = sizeof(half);
			}
			else if(pixel_type == UINT)
			{
				uints.push_back(std::vector<unsigned int>(np));
				for(size_t i = 0; i < np; ++i) uints.back()[i] = (unsigned int)src[i];
				base = (char*)uints.back().data();
				px = sizeof(unsigned int);
			}
			else
			{
				base = (char*)src;
				px = sizeof(float);
			}
			fb.insert(name, Slice((PixelType)pixel_type, base - ((size_t)min_x + (size_t)min_y * w) * px, px, px * w));
		}
		OutputFile file(path, header);
		file.setFrameBuffer(fb);
		file.writePixels(h);
		return 0;
	}
	catch(const std::exception& e)
	{
		fprintf(stderr, "exrref_write_generic: %s\n", e.what());
		return 1;
	}
}

}

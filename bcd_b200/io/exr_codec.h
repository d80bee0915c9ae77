// Minimal OpenEXR 2.x scanline codec, written from scratch for the file formats the reference's I/O layer produces and
// consumes (superboubek/bcd src/io/exr/io_exr.cpp:45-373 through the vendored OpenEXR 2.2 library):
//   * colour images:         HALF R,G,B,A scan lines, PIZ compressed (RgbaOutputFile defaults, io_exr.cpp:156-164)
//   * histogram / covariance: FLOAT channels Bin_0000 .. Bin_NNNN, ZIP compressed (Header defaults, io_exr.cpp:215-236)
// Reads single-part scan-line files with UINT / HALF / FLOAT channels compressed with NONE, RLE, ZIPS, ZIP or PIZ;
// writes HALF or FLOAT channels with NONE, ZIPS, ZIP or PIZ. Tiled, deep, multi-part files and the lossy codecs (PXR24, B44,
// DWA) are rejected with a message. The format follows the published OpenEXR file layout; parity is pinned against the
// reference's own library built in place (oracle/_ref/libexr_ref.so, tests/test_exr_io.py).
#ifndef BCD_B200_EXR_CODEC_H
#define BCD_B200_EXR_CODEC_H

#include <string>
#include <vector>

namespace bcdio
{

enum ExrPixelType { EXR_UINT = 0, EXR_HALF = 1, EXR_FLOAT = 2 };
enum ExrCompression { EXR_NONE = 0, EXR_RLE = 1, EXR_ZIPS = 2, EXR_ZIP = 3, EXR_PIZ = 4 };

struct ExrFile
{
	int width = 0, height = 0;          // size of the data window
	int minX = 0, minY = 0;             // its origin
	int compression = EXR_NONE;
	std::vector<std::string> names;     // channel names in file (alphabetical) order
	std::vector<int> types;             // ExrPixelType per channel
	std::vector<std::vector<float> > planes; // one width*height plane per channel, converted to float

	int find(const std::string& name) const
	{
		for(size_t i = 0; i < names.size(); ++i)
			if(names[i] == name) return (int)i;
		return -1;
	}
};

// Two-step reading for callers that want the pixels interleaved (HWC) without the intermediate planes: exrOpen parses
// the header and the chunk table (info.planes stays empty), exrDecodeInterleaved decompresses straight into
// hwc[(y * width + x) * depth + z] where depth = channelOfSlot.size() and channelOfSlot[z] is the index of the file
// channel that goes to position z (or -1: the position is left untouched, e.g. pre-filled with zeros).
struct ExrReader
{
	struct Chunk { size_t offset, size; long line; int lines; };
	std::vector<unsigned char> file; // the whole file
	ExrFile info;
	std::vector<Chunk> chunks;
	size_t lineBytes = 0;
};
bool exrOpen(const char* path, ExrReader& rd, std::string& err);
bool exrDecodeInterleaved(const ExrReader& rd, const std::vector<int>& channelOfSlot, float* hwc, std::string& err);

// All of them return false and fill `err` on failure; nothing is thrown and the process is never terminated (the reference's
// wrappers call exit(-1), io_exr.cpp:117-120 — not reproduced).
bool exrRead(const char* path, ExrFile& out, std::string& err);
bool exrWrite(const char* path, int width, int height, const std::vector<std::string>& names,
		const std::vector<const float*>& planes, ExrPixelType type, ExrCompression compression, std::string& err);

// IEEE binary16 <-> binary32, round to nearest even like OpenEXR's half (half.h:440-490)
unsigned short floatToHalf(float f);
float halfToFloat(unsigned short h);

} // namespace bcdio

#endif

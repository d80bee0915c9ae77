// bcd_b200 — EXR entry points with the reference's public surface (superboubek/bcd include/bcd/io/ImageIO.h:20-35;
// behaviour of src/io/ImageIO.cpp:22-141 over src/io/exr/io_exr.cpp). The files are read and written by this
// repository's own scan-line codec (bcd_b200/io/exr_codec.cpp), not by the OpenEXR library.
#ifndef BCD_B200_IMAGE_IO_H
#define BCD_B200_IMAGE_IO_H

#include <string>

namespace bcd
{

	template<typename T> class DeepImage;

	class ImageIO
	{
	private:
		ImageIO() {}

	public:
		/// Channels R, G, B of any supported pixel type -> W x H x 3 image (x 1 when the three are identical everywhere,
		/// like the reference, ImageIO.cpp:41-52). A missing channel reads as 0.
		static bool loadEXR(DeepImage<float>& o_rImage, const char* i_pFilePath);
		/// Channels Bin_0000, Bin_0001, ... (as many as are present consecutively) -> W x H x n image.
		static bool loadMultiChannelsEXR(DeepImage<float>& o_rImage, const char* i_pFilePath);

		/// HALF R, G, B, A (A = 1), PIZ compressed, like the reference's RgbaOutputFile; a 1-channel image is written as grey.
		static bool writeEXR(const DeepImage<float>& i_rImage, const char* i_pFilePath);
		/// FLOAT channels Bin_0000 ... Bin_{depth-1}, ZIP compressed, like the reference.
		static bool writeMultiChannelsEXR(const DeepImage<float>& i_rImage, const char* i_pFilePath);

		/// bcd_b200 extension: message of the last failed call of this thread ("" if none).
		static const std::string& lastError();
	};

} // namespace bcd

#endif // BCD_B200_IMAGE_IO_H

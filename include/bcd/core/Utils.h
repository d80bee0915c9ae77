// bcd_b200 — helper functions of the reference's core library kept for the front ends
// (superboubek/bcd include/bcd/core/Utils.h:20-41; behaviour of src/core/Utils.cpp:21-113), written from scratch.
#ifndef BCD_B200_UTILS_H
#define BCD_B200_UTILS_H

#include "DeepImage.h"

#include <string>

namespace bcd
{

	class Utils
	{
	public:
		/// The histogram files carry the per-pixel sample count as one extra, last channel (src/raw_converter/main.cpp:
		/// 176-184): splits a (W, H, D+1) image into histogram (D) and count (1).
		static bool separateNbOfSamplesFromHistogram(
				DeepImage<float>& o_rHistoImage,
				DeepImage<float>& o_rNbOfSamplesImage,
				const DeepImage<float>& i_rHistoAndNbOfSamplesImage);

		/// Inverse of separateNbOfSamplesFromHistogram.
		static DeepImage<float> mergeHistogramAndNbOfSamples(
				const DeepImage<float>& i_rHistoImage,
				const DeepImage<float>& i_rNbOfSamplesImage);

		/// "a/b/c.exr" -> "a/b/", "c.exr" -> "" ('/' separated).
		static std::string extractFolderPath(const std::string& i_rFilePath);

		/// Path of a file relative to a folder (both given in the same absolute or relative frame; the folder ends with '/').
		static std::string getRelativePathFromFolder(const std::string& i_rFileAbsolutePath, const std::string& i_rFolderAbsolutePath);
	};

} // namespace bcd

#endif // BCD_B200_UTILS_H

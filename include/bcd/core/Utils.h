// bcd_b200 — the four helpers of the reference's core library that its front ends use (API contract: superboubek/bcd
// include/bcd/core/Utils.h:20-41; behaviour: src/core/Utils.cpp:21-113). Written from scratch in bcd_b200/host/Utils.cpp.
#ifndef BCD_B200_UTILS_H
#define BCD_B200_UTILS_H

#include "DeepImage.h"

#include <string>

namespace bcd
{

	class Utils
	{
	public:
		// ---- paths ('/' separated), as the preset files use them
		/// Everything up to and including the last '/': "a/b/c.exr" -> "a/b/", "c.exr" -> "".
		static std::string extractFolderPath(const std::string& filePath);
		/// `filePath` as seen from `folderPath` (which ends with '/'); both in the same frame, absolute or not:
		/// ("/s/b/f.exr", "/s/a/") -> "../b/f.exr".
		static std::string getRelativePathFromFolder(const std::string& filePath, const std::string& folderPath);

		// ---- the histogram files carry the per-pixel sample count as one extra, last channel
		//      (src/raw_converter/main.cpp:176-184, src/cli/main.cpp:185)
		/// (W, H, D) histogram + (W, H, 1) count -> (W, H, D + 1).
		static DeepImage<float> mergeHistogramAndNbOfSamples(const DeepImage<float>& histogram, const DeepImage<float>& nbOfSamples);
		/// (W, H, D + 1) -> (W, H, D) histogram and (W, H, 1) count; false when the input has no channel at all.
		static bool separateNbOfSamplesFromHistogram(DeepImage<float>& histogramOut, DeepImage<float>& nbOfSamplesOut,
				const DeepImage<float>& histogramAndNbOfSamples);
	};

} // namespace bcd

#endif // BCD_B200_UTILS_H

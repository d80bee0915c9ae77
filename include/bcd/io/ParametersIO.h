// bcd_b200 — `.bcd.json` pipeline presets. API contract from superboubek/bcd include/bcd/io/ParametersIO.h:17-79
// (holder and field names, ParametersIO::load / write and their selector argument); file format and behaviour from
// src/io/ParametersIO.cpp:36-162. The reference parses with the vendored nlohmann::json; here a small flat-object JSON
// reader / writer in bcd_b200/io/ParametersIO.cpp does the same job.
#ifndef BCD_B200_PARAMETERS_IO_H
#define BCD_B200_PARAMETERS_IO_H

#include "IDenoiser.h"

#include <string>

namespace bcd
{

	/// Which groups of keys load() / write() touch (all by default).
	struct PipelineParametersSelector
	{
		bool m_inputFileNames = true;
		bool m_prefilteringParameters = true;
		bool m_denoiserParameters = true;
	};

	/// The three EXR inputs of a shot (stored relative to the preset's folder, handed back prefixed with it).
	struct InputFileNames
	{
		std::string m_colors, m_histograms, m_covariances;
	};

	/// Spike-removal prefilter switch and its threshold in standard deviations.
	struct PrefilteringParameters
	{
		bool m_performSpikeRemoval = true;
		float m_spikeRemovalThresholdStDevFactor = 1.5f;
	};

	/// Pyramid depth + the per-scale denoiser tunables.
	struct MultiscaleDenoiserParameters
	{
		int m_nbOfScales = 3;
		DenoiserParameters m_monoscaleParameters;
	};

	/// Everything a preset can hold.
	struct PipelineParameters
	{
		InputFileNames m_inputFileNames;
		PrefilteringParameters m_prefilteringParameters;
		MultiscaleDenoiserParameters m_denoiserParameters;
	};

	class ParametersIO
	{
	public:
		/// "bcd.json" — what the reference's tools append to a preset's name.
		static const std::string& getPipelineParametersFileExtension()
		{
			static const std::string s_extension("bcd.json");
			return s_extension;
		}

		/// A key that is absent leaves the corresponding field of o_rParams untouched. False (message on cerr) for an
		/// unreadable or malformed file.
		static bool load(PipelineParameters& o_rParams, const std::string& i_rFilePath,
				PipelineParametersSelector i_selector = PipelineParametersSelector());

		/// Sorted keys, four-space indentation (the layout nlohmann's dump(4) gives the reference's files).
		static bool write(const PipelineParameters& i_rParams, const std::string& i_rFilePath,
				PipelineParametersSelector i_selector = PipelineParametersSelector());

	private:
		ParametersIO();
	};

} // namespace bcd

#endif // BCD_B200_PARAMETERS_IO_H

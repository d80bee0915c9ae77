// bcd_b200 — `.bcd.json` pipeline presets with the reference's public surface
// (superboubek/bcd include/bcd/io/ParametersIO.h:17-79; behaviour of src/io/ParametersIO.cpp:36-162). The reference
// parses with the vendored nlohmann::json; here a small flat-object JSON reader/writer does the same job.
#ifndef BCD_B200_PARAMETERS_IO_H
#define BCD_B200_PARAMETERS_IO_H

#include "IDenoiser.h"

#include <string>

namespace bcd
{

	struct InputFileNames
	{
		std::string m_colors;
		std::string m_histograms;
		std::string m_covariances;
	};

	struct PrefilteringParameters
	{
		PrefilteringParameters() : m_performSpikeRemoval(true), m_spikeRemovalThresholdStDevFactor(1.5f) {}
		bool m_performSpikeRemoval;
		float m_spikeRemovalThresholdStDevFactor;
	};

	struct MultiscaleDenoiserParameters
	{
		MultiscaleDenoiserParameters() : m_nbOfScales(3), m_monoscaleParameters() {}
		int m_nbOfScales;
		DenoiserParameters m_monoscaleParameters;
	};

	struct PipelineParameters
	{
		InputFileNames m_inputFileNames;
		PrefilteringParameters m_prefilteringParameters;
		MultiscaleDenoiserParameters m_denoiserParameters;
	};

	struct PipelineParametersSelector
	{
		PipelineParametersSelector() : m_inputFileNames(true), m_prefilteringParameters(true), m_denoiserParameters(true) {}
		bool m_inputFileNames;
		bool m_prefilteringParameters;
		bool m_denoiserParameters;
	};

	class ParametersIO
	{
	private:
		ParametersIO() {}

	public:
		static const std::string& getPipelineParametersFileExtension()
		{
			static const std::string extension = "bcd.json";
			return extension;
		}

		/// Keys that are absent leave the corresponding field of o_rParams untouched; file names are stored relative to
		/// the preset's folder and come back prefixed with it. False (with a message on cerr) for an unreadable or
		/// malformed file.
		static bool load(PipelineParameters& o_rParams, const std::string& i_rFilePath,
				PipelineParametersSelector i_selector = PipelineParametersSelector());

		static bool write(const PipelineParameters& i_rParams, const std::string& i_rFilePath,
				PipelineParametersSelector i_selector = PipelineParametersSelector());
	};

} // namespace bcd

#endif // BCD_B200_PARAMETERS_IO_H

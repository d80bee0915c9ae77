// bcd_cli — command-line front end of the Bayesian collaborative denoiser on the B200 path.
// Same options, defaults, file conventions and messages as the reference's bcd-cli
// (superboubek/bcd src/cli/main.cpp:46-59 defaults, :99-121 usage, :124-384 parsing, :422-474 pipeline), re-authored
// over this repository's bcd:: classes. Differences, all documented in INTEGRATION.md:
//   * "-a <file.bcd.json>" is honoured (the reference prints it in its usage text but never parses it): the preset is
//     applied first, explicit options given after it override it;
//   * "--device <n>" selects the GPU; "--ncores" and "--use-cuda" are accepted and ignored (there is no CPU path);
//   * a failed denoise() ends the program with exit code 1 instead of writing the untouched input.
#include "DeepImage.h"
#include "Denoiser.h"
#include "IDenoiser.h"
#include "MultiscaleDenoiser.h"
#include "SpikeRemovalFilter.h"
#include "Utils.h"
#include "ImageIO.h"
#include "ParametersIO.h"

#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

using namespace std;
using namespace bcd;

namespace
{

	const char* g_programPath = "bcd_cli";

	struct Arguments
	{
		string m_output, m_color, m_hist, m_cov, m_preset;
		float m_threshold = 1.f;
		int m_patchRadius = 1, m_searchRadius = 6;
		float m_minEigenValue = 1.e-8f;
		bool m_randomOrder = true;
		bool m_prefilterSpikes = true;
		float m_prefilterFactor = 2.f;
		float m_skipProbability = 1.f;
		int m_nbOfScales = 3;
		int m_nbOfCores = 0;
		bool m_useCuda = true;
		int m_device = 0;
		int m_nbOfGpus = 1;          // --gpus: devices m_device .. m_device + n - 1 (0 = every visible GPU)
		bool m_spikeFloatAbs = false; // --spike-float-abs: the float reading of the reference's abs() in the spike prefilter
	};

	void printUsage()
	{
		const Arguments d;
		cout << "Bayesian Collaborative Denoising" << endl << endl;
		cout << "Usage: " << g_programPath << " <arguments list>" << endl;
		cout << "Only EXR images are supported." << endl << endl;
		cout << "Required arguments list (unless a pipeline file is provided and contains this data):" << endl;
		cout << "    -o <output>          The file path to the output image" << endl;
		cout << "    -i <input>           The file path to the input image" << endl;
		cout << "    -h <hist>            The file path to the input histograms buffer" << endl;
		cout << "    -c <cov>             The file path to the input covariance matrices buffer" << endl;
		cout << "Optional arguments list:" << endl;
		cout << "    -a <file>            The file path to the .bcd.json file containing arguments for the program" << endl;
		cout << "    -d <float>           Histogram patch distance threshold (default: " << d.m_threshold << ")" << endl;
		cout << "    -b <int>             Radius of search windows (default: " << d.m_searchRadius << ")" << endl;
		cout << "    -w <int>             Radius of patches (default: " << d.m_patchRadius << ")" << endl;
		cout << "    -r <0/1>             1 for random pixel order (accepted; the B200 path always realises the raster order) (default: " << (d.m_randomOrder ? 1 : 0) << ")" << endl;
		cout << "    -p <0/1>             1 for a spike removal prefiltering (default: " << (d.m_prefilterSpikes ? 1 : 0) << ")" << endl;
		cout << "    --p-factor <float>   Factor that is multiplied by standard deviation to get the threshold for classifying spikes during prefiltering. Put lower value to remove more spikes (default: " << d.m_prefilterFactor << ")" << endl;
		cout << "    -m <float in [0,1]>  Probability of skipping marked centers of denoised patches. 1 accelerates a lot the computations. 0 helps removing potential grid artifacts (default: " << d.m_skipProbability << ")" << endl;
		cout << "    -s <int>             Number of Scales for Multi-Scaling (default: " << d.m_nbOfScales << ")" << endl;
		cout << "    --ncores <nbOfCores> Accepted for compatibility (the B200 path has no CPU threads to size)" << endl;
		cout << "    --use-cuda <0/1>     Accepted for compatibility (the B200 path is CUDA-only)" << endl;
		cout << "    -e <float>           Minimum eigen value for matrix inversion (default: " << d.m_minEigenValue << ")" << endl;
		cout << "    --device <int>       CUDA device index (default: " << d.m_device << ")" << endl;
		cout << "    --gpus <int>         Number of GPUs the frame is spread over, starting at --device; 0 = all visible (default: " << d.m_nbOfGpus << ")" << endl;
		cout << "    --spike-float-abs <0/1>  1: spike prefilter compares |colour - mean| as a float (MSVC builds of bcd-cli); 0: truncated to an integer like the GCC build (default: 0)" << endl;
	}

	template <typename T> bool parseValue(const char* text, T& out)
	{
		istringstream iss(text);
		iss >> out;
		return !iss.fail();
	}

	bool parse01(const char* option, const char* text, bool& out)
	{
		int v = -1;
		if(!parseValue(text, v) || (v != 0 && v != 1))
		{
			cout << "ERROR in program arguments: expecting 0 or 1 after '" << option << "'" << endl;
			return false;
		}
		out = v == 1;
		return true;
	}

	bool parseArguments(int argc, const char** argv, Arguments& a)
	{
		struct Expect { const char* option; const char* what; };
		static const Expect expectations[] = {
			{"-o", "file path to the output image"}, {"-i", "file path to the input color image"},
			{"-h", "file path to the input histogram image"}, {"-c", "file path to the input covariance matrix image"},
			{"-a", "file path to the .bcd.json pipeline file"}, {"-d", "histogram patch distance threshold"},
			{"-b", "radius of search window"}, {"-w", "radius of patch"}, {"-e", "minimum eigen value"},
			{"-r", "0 or 1"}, {"-p", "0 or 1"}, {"--p-factor", "standard deviation factor for spike prefiltering threshold"},
			{"-m", "float in [0,1]"}, {"-s", "number of scales"}, {"--ncores", "number of cores for OpenMP"},
			{"--use-cuda", "0 or 1"}, {"--device", "CUDA device index"}, {"--gpus", "number of GPUs"},
			{"--spike-float-abs", "0 or 1"}};
		for(int i = 1; i < argc; ++i)
		{
			const string opt = argv[i];
			const Expect* e = nullptr;
			for(const Expect& x : expectations)
				if(opt == x.option) e = &x;
			if(!e)
				continue; // unknown words are ignored, like the reference does
			if(++i == argc)
			{
				cout << "ERROR in program arguments: expecting " << e->what << " after '" << opt << "'" << endl;
				return false;
			}
			const char* v = argv[i];
			if(opt == "-o")
			{
				a.m_output = v;
				if(!ofstream(a.m_output.c_str(), ofstream::out | ofstream::app))
				{
					cout << "ERROR in program arguments: cannot write output file '" << a.m_output << "'" << endl;
					return false;
				}
			}
			else if(opt == "-i") a.m_color = v;
			else if(opt == "-h") a.m_hist = v;
			else if(opt == "-c") a.m_cov = v;
			else if(opt == "-a")
			{
				a.m_preset = v;
				PipelineParameters pp;
				pp.m_prefilteringParameters.m_performSpikeRemoval = a.m_prefilterSpikes;
				pp.m_prefilteringParameters.m_spikeRemovalThresholdStDevFactor = a.m_prefilterFactor;
				pp.m_denoiserParameters.m_nbOfScales = a.m_nbOfScales;
				DenoiserParameters& p = pp.m_denoiserParameters.m_monoscaleParameters;
				p.m_histogramDistanceThreshold = a.m_threshold;
				p.m_patchRadius = a.m_patchRadius;
				p.m_searchWindowRadius = a.m_searchRadius;
				p.m_minEigenValue = a.m_minEigenValue;
				p.m_useRandomPixelOrder = a.m_randomOrder;
				p.m_markedPixelsSkippingProbability = a.m_skipProbability;
				p.m_nbOfCores = a.m_nbOfCores;
				p.m_useCuda = a.m_useCuda;
				pp.m_inputFileNames.m_colors = a.m_color;
				pp.m_inputFileNames.m_histograms = a.m_hist;
				pp.m_inputFileNames.m_covariances = a.m_cov;
				if(!ParametersIO::load(pp, a.m_preset))
				{
					cout << "ERROR in program arguments: couldn't load pipeline file '" << a.m_preset << "'" << endl;
					return false;
				}
				a.m_prefilterSpikes = pp.m_prefilteringParameters.m_performSpikeRemoval;
				a.m_prefilterFactor = pp.m_prefilteringParameters.m_spikeRemovalThresholdStDevFactor;
				a.m_nbOfScales = pp.m_denoiserParameters.m_nbOfScales;
				a.m_threshold = p.m_histogramDistanceThreshold;
				a.m_patchRadius = p.m_patchRadius;
				a.m_searchRadius = p.m_searchWindowRadius;
				a.m_minEigenValue = p.m_minEigenValue;
				a.m_randomOrder = p.m_useRandomPixelOrder;
				a.m_skipProbability = p.m_markedPixelsSkippingProbability;
				a.m_nbOfCores = p.m_nbOfCores;
				a.m_useCuda = p.m_useCuda;
				a.m_color = pp.m_inputFileNames.m_colors;
				a.m_hist = pp.m_inputFileNames.m_histograms;
				a.m_cov = pp.m_inputFileNames.m_covariances;
			}
			else if(opt == "-d") parseValue(v, a.m_threshold);
			else if(opt == "-b") parseValue(v, a.m_searchRadius);
			else if(opt == "-w") parseValue(v, a.m_patchRadius);
			else if(opt == "-e") parseValue(v, a.m_minEigenValue);
			else if(opt == "-r") { if(!parse01("-r", v, a.m_randomOrder)) return false; }
			else if(opt == "-p") { if(!parse01("-p", v, a.m_prefilterSpikes)) return false; }
			else if(opt == "--p-factor") parseValue(v, a.m_prefilterFactor);
			else if(opt == "-m")
			{
				float m = -1.f;
				if(!parseValue(v, m) || m < 0.f || m > 1.f)
				{
					cout << "ERROR in program arguments: expecting float in [0,1] after '-m'" << endl;
					return false;
				}
				a.m_skipProbability = m;
			}
			else if(opt == "-s") parseValue(v, a.m_nbOfScales);
			else if(opt == "--ncores") parseValue(v, a.m_nbOfCores);
			else if(opt == "--use-cuda") { if(!parse01("--use-cuda", v, a.m_useCuda)) return false; }
			else if(opt == "--device") parseValue(v, a.m_device);
			else if(opt == "--gpus") parseValue(v, a.m_nbOfGpus);
			else if(opt == "--spike-float-abs") { if(!parse01("--spike-float-abs", v, a.m_spikeFloatAbs)) return false; }
		}
		// the histogram and covariance files default to <colour>_hist.exr / <colour>_cov.exr (main.cpp:345-371)
		if(!a.m_color.empty() && a.m_color.size() > 4)
		{
			const string stem = a.m_color.substr(0, a.m_color.size() - 4);
			if(a.m_hist.empty())
			{
				a.m_hist = stem + "_hist.exr";
				cout << "Warning: input histogram file not provided by -h argument: assuming '" << a.m_hist << "'" << endl;
			}
			if(a.m_cov.empty())
			{
				a.m_cov = stem + "_cov.exr";
				cout << "Warning: input covariance file not provided by -c argument: assuming '" << a.m_cov << "'" << endl;
			}
		}
		if(a.m_color.empty() || a.m_hist.empty() || a.m_cov.empty() || a.m_output.empty())
		{
			cout << "ERROR: Missing required program argument(s):";
			if(a.m_color.empty()) cout << " -i";
			if(a.m_hist.empty()) cout << " -h";
			if(a.m_cov.empty()) cout << " -c";
			if(a.m_output.empty()) cout << " -o";
			cout << endl << endl;
			printUsage();
			return false;
		}
		return true;
	}

	/// Negative, infinite and NaN output values are written as 0 (main.cpp:389-420).
	void zeroBadValues(DeepImage<float>& io_rImage)
	{
		float* p = io_rImage.getDataPtr();
		const size_t n = (size_t)io_rImage.getWidth() * io_rImage.getHeight() * io_rImage.getDepth();
		for(size_t i = 0; i < n; ++i)
			if(p[i] < 0.f || std::isnan(p[i]) || std::isinf(p[i]))
				p[i] = 0.f;
	}

	int run(int argc, const char** argv)
	{
		Arguments a;
		if(!parseArguments(argc, argv, a))
			return 1;
		DeepImage<float> color, nbOfSamples, histogram, covariance, histAndCount;
		if(!ImageIO::loadEXR(color, a.m_color.c_str()))
		{
			cout << "ERROR in program arguments: couldn't load input color image file '" << a.m_color << "'" << endl;
			return 1;
		}
		if(!ImageIO::loadMultiChannelsEXR(histAndCount, a.m_hist.c_str()))
		{
			cout << "ERROR in program arguments: couldn't load input histogram image file '" << a.m_hist << "'" << endl;
			return 1;
		}
		Utils::separateNbOfSamplesFromHistogram(histogram, nbOfSamples, histAndCount);
		histAndCount.clearAndFreeMemory();
		if(!ImageIO::loadMultiChannelsEXR(covariance, a.m_cov.c_str()))
		{
			cout << "ERROR in program arguments: couldn't load input covariance matrix image file '" << a.m_cov << "'" << endl;
			return 1;
		}
		if(color.getDepth() == 1) // loadEXR collapses grey images; the denoiser works on three channels
		{
			DeepImage<float> rgb(color.getWidth(), color.getHeight(), 3);
			const float* s = color.getDataPtr();
			float* d = rgb.getDataPtr();
			for(size_t i = 0, n = (size_t)color.getWidth() * color.getHeight(); i < n; ++i)
				d[3 * i] = d[3 * i + 1] = d[3 * i + 2] = s[i];
			color = rgb;
		}

		SpikeRemovalFilter::s_device = a.m_device;
		SpikeRemovalFilter::s_useIntegerAbsLikeGccBuild = !a.m_spikeFloatAbs;
		if(a.m_prefilterSpikes)
		{
			SpikeRemovalFilter::filter(color, nbOfSamples, histogram, covariance, a.m_prefilterFactor);
			if(!SpikeRemovalFilter::s_lastCallSucceeded) // mismatched input files or no device: do not denoise unfiltered inputs silently
			{
				std::cerr << "ERROR: spike removal prefiltering failed" << std::endl;
				return 1;
			}
		}

		DenoiserInputs inputs;
		inputs.m_pColors = &color;
		inputs.m_pNbOfSamples = &nbOfSamples;
		inputs.m_pHistograms = &histogram;
		inputs.m_pSampleCovariances = &covariance;
		DeepImage<float> denoised(color);
		DenoiserOutputs outputs;
		outputs.m_pDenoisedColors = &denoised;
		DenoiserParameters parameters;
		parameters.m_histogramDistanceThreshold = a.m_threshold;
		parameters.m_patchRadius = a.m_patchRadius;
		parameters.m_searchWindowRadius = a.m_searchRadius;
		parameters.m_minEigenValue = a.m_minEigenValue;
		parameters.m_useRandomPixelOrder = a.m_randomOrder;
		parameters.m_markedPixelsSkippingProbability = a.m_skipProbability;
		parameters.m_nbOfCores = a.m_nbOfCores;
		parameters.m_useCuda = a.m_useCuda;

		unique_ptr<Denoiser> denoiser(a.m_nbOfScales > 1 ? new MultiscaleDenoiser(a.m_nbOfScales) : new Denoiser());
		if(a.m_nbOfGpus == 1)
			denoiser->setDevice(a.m_device);
		else if(a.m_nbOfGpus <= 0)
			denoiser->setDevices(nullptr, 0); // every visible GPU
		else
		{
			std::vector<int> ids;
			for(int i = 0; i < a.m_nbOfGpus; ++i) ids.push_back(a.m_device + i);
			denoiser->setDevices(ids.data(), int(ids.size()));
		}
		denoiser->setInputs(inputs);
		denoiser->setOutputs(outputs);
		denoiser->setParameters(parameters);
		if(!denoiser->denoise())
		{
			cerr << "Error in program '" << g_programPath << "': denoising failed" << endl;
			return 1;
		}
		zeroBadValues(denoised);
		if(!ImageIO::writeEXR(denoised, a.m_output.c_str()))
			return 1;
		cout << "Written denoised output in file " << a.m_output << endl;
		return 0;
	}

} // namespace

int main(int argc, const char** argv)
{
	const chrono::steady_clock::time_point t0 = chrono::steady_clock::now();
	if(argc > 0) g_programPath = argv[0];
	const int rc = run(argc, argv);
	const double s = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
	cout << "Program total time: " << s << " s" << endl;
	cout << "Exiting program!" << endl;
	return rc;
}

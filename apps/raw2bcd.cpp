// raw2bcd — converts a raw dump of all the samples of a rendering into the three EXR inputs of bcd_cli
// (same file conventions as the reference's converter, superboubek/bcd src/raw_converter/main.cpp:36-41,65-72,142-184):
//   input:  5 x int32 {version, width, height, nbOfSamples, nbOfChannels} then, pixel-major, nbOfSamples x nbOfChannels
//           float32 per pixel (the first three channels are R, G, B)
//   output: <prefix>.exr (mean colour), <prefix>_hist.exr (histogram + sample count as the last channel),
//           <prefix>_cov.exr (covariance)
#include "DeepImage.h"
#include "SamplesAccumulator.h"
#include "Utils.h"
#include "ImageIO.h"

#include <cstdint>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

using namespace std;
using namespace bcd;

int main(int argc, const char** argv)
{
	if(argc != 3)
	{
		cout << "raw2bcd" << endl << endl;
		cout << "Usage: raw2bcd <input> <outputPrefix>" << endl;
		cout << "Converts a raw file with all samples into the inputs for the BayesianCollaborativeDenoiser program" << endl;
		cout << "Required arguments list:" << endl;
		cout << "    <input>           The file path to the input image" << endl;
		cout << "    <outputPrefix>    The file path to the output image, without .exr extension" << endl;
		return 1;
	}
	ifstream in(argv[1], ios::binary);
	if(!in)
	{
		cerr << "ERROR in program arguments: cannot open input file '" << argv[1] << "'" << endl;
		return 1;
	}
	int32_t header[5] = {0, 0, 0, 0, 0};
	in.read(reinterpret_cast<char*>(header), sizeof header);
	const int32_t width = header[1], height = header[2], nbOfSamples = header[3], nbOfChannels = header[4];
	if(!in || width <= 0 || height <= 0 || nbOfSamples <= 0 || nbOfChannels < 3 || nbOfChannels > 16)
	{
		cerr << "ERROR: '" << argv[1] << "' does not start with a valid raw header" << endl;
		return 1;
	}
	cout << "Version: " << header[0] << endl;
	cout << "Resolution: " << width << "x" << height << endl;
	cout << "Nb of samples: " << nbOfSamples << endl;
	cout << "Nb of channels: " << nbOfChannels << endl;
	const string prefix = argv[2];
	const string colorPath = prefix + ".exr", histPath = prefix + "_hist.exr", covPath = prefix + "_cov.exr";
	for(const string* p : {&colorPath, &histPath, &covPath})
		if(!ofstream(p->c_str()))
		{
			cerr << "ERROR in program arguments: cannot write output file '" << *p << "'" << endl;
			return 1;
		}

	SamplesAccumulator accumulator(width, height, HistogramParameters());
	vector<float> pixelSamples((size_t)nbOfSamples * nbOfChannels);
	for(int32_t line = 0; line < height; ++line)
		for(int32_t col = 0; col < width; ++col)
		{
			in.read(reinterpret_cast<char*>(pixelSamples.data()), (streamsize)(pixelSamples.size() * sizeof(float)));
			if(!in)
			{
				cerr << "ERROR: '" << argv[1] << "' ends before pixel (" << line << "," << col << ")" << endl;
				return 1;
			}
			for(int32_t s = 0; s < nbOfSamples; ++s)
			{
				const float* v = pixelSamples.data() + (size_t)s * nbOfChannels;
				accumulator.addSample(line, col, v[0], v[1], v[2]);
			}
		}
	SamplesStatisticsImages stats = accumulator.extractSamplesStatistics();
	const DeepImage<float> histAndCount = Utils::mergeHistogramAndNbOfSamples(stats.m_histoImage, stats.m_nbOfSamplesImage);
	bool ok = ImageIO::writeEXR(stats.m_meanImage, colorPath.c_str());
	ok = ImageIO::writeMultiChannelsEXR(stats.m_covarImage, covPath.c_str()) && ok;
	ok = ImageIO::writeMultiChannelsEXR(histAndCount, histPath.c_str()) && ok;
	cout << "Exiting program!" << endl;
	return ok ? 0 : 1;
}

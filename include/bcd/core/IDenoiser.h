// bcd_b200 — source-compatible re-statement of the reference's denoiser interface
// (superboubek/bcd include/bcd/core/IDenoiser.h:20-97). Field and method names are the API a host renderer or
// bcd_cli compiles against, so they are kept; defaults match IDenoiser.h:23-33. Implementations of denoise() live
// in bcd_b200/host/Denoiser.cpp and forward to the C ABI of include/bcd_b200.h.
#ifndef BCD_B200_I_DENOISER_H
#define BCD_B200_I_DENOISER_H

#include <functional>
#include <memory>

namespace bcd
{

	template <class T> class DeepImage;

	/// Tunables of the collaborative filter.
	class DenoiserParameters
	{
	public:
		DenoiserParameters() :
				m_histogramDistanceThreshold(1.f),
				m_patchRadius(1),
				m_searchWindowRadius(6),
				m_minEigenValue(0.00000001f),
				m_useRandomPixelOrder(true),
				m_markedPixelsSkippingProbability(1.f),
				m_nbOfCores(0),
				m_useCuda(true)
		{
		}

	public:
		float m_histogramDistanceThreshold;      ///< patches closer than this (chi-square histogram distance) are "similar"
		int m_patchRadius;                       ///< patches are (2r+1)^2 pixels
		int m_searchWindowRadius;                ///< similar patches are searched in a (2r+1)^2 window
		float m_minEigenValue;                   ///< floor used when inverting covariance matrices
		bool m_useRandomPixelOrder;              ///< accepted for compatibility; the B200 path always realises the raster order
		float m_markedPixelsSkippingProbability; ///< 1 = marking strategy, 0 = filter every pixel
		int m_nbOfCores;                         ///< CPU threads in the reference; unused here (overwritten with 1 by denoise())
		bool m_useCuda;                          ///< kept for compatibility; the B200 path is CUDA-only
	};

	/// Borrowed input images (the caller keeps ownership).
	class DenoiserInputs
	{
	public:
		DenoiserInputs() : m_pColors(nullptr), m_pNbOfSamples(nullptr), m_pHistograms(nullptr), m_pSampleCovariances(nullptr) {}

	public:
		const DeepImage<float>* m_pColors;            ///< mean colour, depth 3
		const DeepImage<float>* m_pNbOfSamples;       ///< sample count, depth 1
		const DeepImage<float>* m_pHistograms;        ///< per-channel sample histograms, depth 3 * nbOfBins
		const DeepImage<float>* m_pSampleCovariances; ///< sample covariance xx,yy,zz,yz,xz,xy, depth 6
	};

	/// Borrowed output image; denoise() resizes it to W x H x 3 and overwrites it.
	class DenoiserOutputs
	{
	public:
		DenoiserOutputs() : m_pDenoisedColors(nullptr) {}

	public:
		DeepImage<float>* m_pDenoisedColors;
	};

	class IDenoiser
	{
	public:
		IDenoiser() : m_progressCallback([](float) {}) {}
		virtual ~IDenoiser() {}

		virtual bool denoise() = 0;

		const DenoiserInputs& getInputs() const { return m_inputs; }
		void setInputs(const DenoiserInputs& i_rInputs) { m_inputs = i_rInputs; }
		const DenoiserOutputs& getOutputs() const { return m_outputs; }
		void setOutputs(const DenoiserOutputs& i_rOutputs) { m_outputs = i_rOutputs; }
		const DenoiserParameters& getParameters() const { return m_parameters; }
		void setParameters(const DenoiserParameters& i_rParameters) { m_parameters = i_rParameters; }
		void setProgressCallback(std::function<void(float)> i_progressCallback) { m_progressCallback = i_progressCallback; }

	protected:
		DenoiserParameters m_parameters;
		DenoiserInputs m_inputs;
		DenoiserOutputs m_outputs;
		std::function<void(float)> m_progressCallback;
	};

} // namespace bcd

#endif // BCD_B200_I_DENOISER_H

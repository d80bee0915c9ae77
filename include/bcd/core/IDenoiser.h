// bcd_b200 — the denoiser interface a host renderer or bcd_cli compiles against, re-stated for the B200 path.
// API contract taken from superboubek/bcd include/bcd/core/IDenoiser.h:20-97: the names of the three parameter /
// input / output holders, of their public fields (with the defaults of :23-33) and of the IDenoiser methods are kept,
// because that is the drop-in surface; everything else (layout of this file, in-class initialisers, comments) is ours.
// denoise() is implemented in bcd_b200/host/Denoiser.cpp and forwards to the C ABI of include/bcd_b200.h.
#ifndef BCD_B200_I_DENOISER_H
#define BCD_B200_I_DENOISER_H

#include <functional>
#include <memory>

namespace bcd
{

	template <class T> class DeepImage;

	/// What goes in: four images owned by the caller, borrowed for the duration of denoise().
	class DenoiserInputs
	{
	public:
		const DeepImage<float>* m_pColors = nullptr;            ///< mean colour per pixel, depth 3
		const DeepImage<float>* m_pNbOfSamples = nullptr;       ///< sample count per pixel, depth 1
		const DeepImage<float>* m_pHistograms = nullptr;        ///< per-channel sample histograms, depth 3 * nbOfBins
		const DeepImage<float>* m_pSampleCovariances = nullptr; ///< sample covariance xx,yy,zz,yz,xz,xy, depth 6
	};

	/// What comes out: one caller-owned image that denoise() resizes to W x H x 3 and overwrites.
	class DenoiserOutputs
	{
	public:
		DeepImage<float>* m_pDenoisedColors = nullptr;
	};

	/// Tunables of the collaborative filter (defaults of the reference).
	class DenoiserParameters
	{
	public:
		float m_histogramDistanceThreshold = 1.f;      ///< chi-square histogram distance below which two patches are "similar"
		int m_patchRadius = 1;                         ///< a patch is (2r+1)^2 pixels
		int m_searchWindowRadius = 6;                  ///< similar patches are looked for in a (2r+1)^2 window
		float m_minEigenValue = 1.e-8f;                ///< eigenvalue floor of the covariance inversions
		bool m_useRandomPixelOrder = true;             ///< accepted; the B200 path always realises the raster order
		float m_markedPixelsSkippingProbability = 1.f; ///< 1 = marking strategy, 0 = every pixel is filtered
		int m_nbOfCores = 0;                           ///< OpenMP team size of the reference; no meaning here
		bool m_useCuda = true;                         ///< accepted; the B200 path is CUDA-only
	};

	/// Base of bcd::Denoiser / bcd::MultiscaleDenoiser: holders by value, progress callback, one blocking entry point.
	class IDenoiser
	{
	public:
		virtual ~IDenoiser() {}

		/// Blocking; false (with a message on cerr) when the inputs are unusable or the GPU path fails.
		virtual bool denoise() = 0;

		void setInputs(const DenoiserInputs& i_rInputs) { m_inputs = i_rInputs; }
		void setOutputs(const DenoiserOutputs& i_rOutputs) { m_outputs = i_rOutputs; }
		void setParameters(const DenoiserParameters& i_rParameters) { m_parameters = i_rParameters; }
		/// Called from the thread that runs denoise(), with values rising from 0 to 1.
		void setProgressCallback(std::function<void(float)> i_progressCallback) { m_progressCallback = i_progressCallback; }

		const DenoiserInputs& getInputs() const { return m_inputs; }
		const DenoiserOutputs& getOutputs() const { return m_outputs; }
		const DenoiserParameters& getParameters() const { return m_parameters; }

	protected:
		DenoiserInputs m_inputs;
		DenoiserOutputs m_outputs;
		DenoiserParameters m_parameters;
		std::function<void(float)> m_progressCallback = [](float) {};
	};

} // namespace bcd

#endif // BCD_B200_I_DENOISER_H

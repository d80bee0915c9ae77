// bcd_b200 — single-scale denoiser with the reference's public surface (superboubek/bcd include/bcd/core/Denoiser.h:25-44).
// denoise() forwards to bcd_b200_denoise() (include/bcd_b200.h) with nb_scales = 1; the per-thread accumulation
// images and the DenoisingUnit of the reference have no equivalent here (their work happens in the CUDA kernels).
#ifndef BCD_B200_DENOISER_H
#define BCD_B200_DENOISER_H

#include "IDenoiser.h"
#include "DeepImage.h"

namespace bcd
{

	class Denoiser : public IDenoiser
	{
	public:
		Denoiser() : IDenoiser(), m_device(0) {}
		virtual ~Denoiser() {}

		virtual bool denoise();

		/// Same checks and messages as the reference (Denoiser.cpp:269-346): null, empty or size-mismatched images.
		bool inputsOutputsAreOk();

		/// bcd_b200 extension: CUDA device used by denoise() (default 0).
		void setDevice(int i_device) { m_device = i_device; }
		int getDevice() const { return m_device; }

	protected:
		bool runOnDevice(int i_nbOfScales);

	private:
		int m_device;
	};

} // namespace bcd

#endif // BCD_B200_DENOISER_H

// bcd_b200 — multiscale denoiser with the reference's public surface
// (superboubek/bcd include/bcd/core/MultiscaleDenoiser.h:23-30). The pyramid build, the per-scale denoising and the
// merge (MultiscaleDenoiser.cpp:31-136,243-334,453-548) run on the GPU behind bcd_b200_denoise(nb_scales).
#ifndef BCD_B200_MULTISCALE_DENOISER_H
#define BCD_B200_MULTISCALE_DENOISER_H

#include "Denoiser.h"

namespace bcd
{

	class MultiscaleDenoiser : public Denoiser
	{
	public:
		MultiscaleDenoiser(int i_nbOfScales) : Denoiser(), m_nbOfScales(i_nbOfScales) {}
		virtual ~MultiscaleDenoiser() {}

		virtual bool denoise();

	private:
		int m_nbOfScales;
	};

} // namespace bcd

#endif // BCD_B200_MULTISCALE_DENOISER_H

// bcd_b200 — single-scale denoiser with the reference's public surface (superboubek/bcd include/bcd/core/Denoiser.h:25-44).
// denoise() forwards to the C ABI (include/bcd_b200.h, bcd_b200_group_denoise) with nb_scales = 1; the per-thread
// accumulation images and the DenoisingUnit of the reference have no equivalent here (their work happens in the CUDA
// kernels). The object keeps its device session — buffers, streams, peer links between the GPUs, page-locked host
// ranges — from one denoise() to the next, like the reference's GUI calls denoise() again and again on one object
// (src/gui/GuiWindow.cpp:345-376).
#ifndef BCD_B200_DENOISER_H
#define BCD_B200_DENOISER_H

#include "IDenoiser.h"
#include "DeepImage.h"

#include <vector>

namespace bcd
{

	class Denoiser : public IDenoiser
	{
	public:
		Denoiser();
		virtual ~Denoiser();

		virtual bool denoise();

		/// Same checks and messages as the reference (Denoiser.cpp:269-346): null, empty or size-mismatched images.
		bool inputsOutputsAreOk();

		/// bcd_b200 extension: CUDA device used by denoise() (default 0, or the list in the environment variable
		/// BCD_B200_DEVICES = "all" | "0,1,2,...").
		void setDevice(int i_device);
		int getDevice() const { return m_devices.empty() ? 0 : m_devices[0]; }
		/// bcd_b200 extension: several GPUs of one box; the frame is cut into row tiles that exchange their halos over
		/// NVLink. i_pDeviceIds == nullptr: devices 0 .. i_nbOfDevices-1; i_nbOfDevices <= 0: every visible device.
		void setDevices(const int* i_pDeviceIds, int i_nbOfDevices);
		int getNbOfDevices() const { return int(m_devices.size()); }
		/// bcd_b200 extension: page-lock the input / output image storage the first time it is seen (default true; the
		/// upload of a pageable 580 MB frame is what dominates a 1080p denoise() otherwise).
		void setPinHostBuffers(bool i_pin) { m_pinHostBuffers = i_pin; }
		/// bcd_b200 extension: frees the device session now instead of in the destructor.
		void releaseDeviceResources();

	protected:
		bool runOnDevice(int i_nbOfScales);

	private:
		Denoiser(const Denoiser&);            // the device session is not copyable
		Denoiser& operator=(const Denoiser&);

		std::vector<int> m_devices; // empty: every visible device
		bool m_allDevices;
		bool m_pinHostBuffers;
		void* m_pSession;           // bcd_b200_group*
	};

} // namespace bcd

#endif // BCD_B200_DENOISER_H

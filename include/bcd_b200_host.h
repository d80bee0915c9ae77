/*
 * bcd_b200_host.h — small C entry points of libbcdcore_b200.so used by the Python tests and bench.py:
 * the seeded synthetic-input generator G of SURVEY §8d (benchmark workload), the SamplesAccumulator path for raw
 * sample dumps, and a C wrapper around bcd::Denoiser / bcd::MultiscaleDenoiser so that the C++ drop-in classes
 * themselves can be exercised from ctypes.
 */
#ifndef BCD_B200_HOST_H
#define BCD_B200_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* G(W, H, spp, seed, sigma): counter-based synthetic radiance + log-normal samples (+ fireflies), accumulated with
 * bcd::SamplesAccumulator (weight 1; nb_bins, gamma, max_value as in HistogramParameters). Outputs are HWC float32:
 * color W*H*3, nsamples W*H, hist W*H*3*nb_bins, cov W*H*6. Rows [row0, row0+rows) of the W x H image only (so row
 * tiles can be generated independently). Returns 0. */
int bcd_synth_generate(int width, int height, int spp, uint64_t seed, double sigma,
                       int nb_bins, float gamma, float max_value, int row0, int rows,
                       float* color, float* nsamples, float* hist, float* cov);

/* raw samples [line][col][sample][rgb] -> statistics through bcd::SamplesAccumulator */
int bcd_host_accumulate(const float* samples, int width, int height, int spp, int nb_bins, float gamma, float max_value,
                        float* color, float* nsamples, float* hist, float* cov);

/* bcd::Denoiser (nb_scales == 1) / bcd::MultiscaleDenoiser (nb_scales > 1) driven exactly like a host renderer
 * would (setInputs / setOutputs / setParameters / setProgressCallback / denoise()). Returns 1 when denoise()
 * returned true, 0 otherwise; *progress_calls receives the number of progress callbacks. */
int bcd_host_cxx_denoise(const float* color, const float* nsamples, const float* hist, const float* cov,
                         int width, int height, int hist_depth,
                         float tau, int patch_radius, int search_radius, float min_eig, int random_order, float skip_prob,
                         int nb_scales, int device, float* out_rgb, int* progress_calls);

/* The same drop-in seen by a host renderer that keeps its objects: ONE bcd::Denoiser / bcd::MultiscaleDenoiser and one
 * set of DeepImages in pageable std::vector storage, denoise() called frame after frame (reference call site
 * src/gui/GuiWindow.cpp:345-376). devices == NULL && nb_devices == 0: the object's default (device 0 or BCD_B200_DEVICES);
 * nb_devices < 0: every visible GPU. load = the renderer writing a frame into its images (not part of denoise());
 * denoise returns 1 when denoise() returned true; fetch copies the output image. */
typedef struct bcd_host_session bcd_host_session;
bcd_host_session* bcd_host_session_create(int nb_scales, const int* devices, int nb_devices, int pin_host);
void bcd_host_session_destroy(bcd_host_session* s);
int bcd_host_session_load(bcd_host_session* s, const float* color, const float* nsamples, const float* hist, const float* cov,
                          int width, int height, int hist_depth);
int bcd_host_session_denoise(bcd_host_session* s, float tau, int patch_radius, int search_radius, float min_eig,
                             int random_order, float skip_prob);
int bcd_host_session_fetch(bcd_host_session* s, float* out_rgb);
int bcd_host_session_nb_devices_used(bcd_host_session* s); /* m_nbOfCores after denoise(): GPUs that worked on the frame */

/* bcd::ImageIO from C (tests): load returns 1 and the image size (HWC float32; multi_channel = 0: R,G,B file -> depth 3,
 * or 1 for a grey file; 1: Bin_%04d file), fetch copies the pixels of the image loaded last by this thread; write
 * returns 1 on success. bcd_host_exr_last_error: message of the last failure. */
int bcd_host_exr_load(const char* path, int multi_channel, int* width, int* height, int* depth);
int bcd_host_exr_fetch(float* out_hwc);
int bcd_host_exr_write(const char* path, int multi_channel, const float* hwc, int width, int height, int depth);
const char* bcd_host_exr_last_error(void);
/* the codec's writer with an explicit pixel type (1 HALF, 2 FLOAT) and compression (0 NONE, 2 ZIPS, 3 ZIP, 4 PIZ);
 * multi_channel = 0 needs depth 3 (channels R, G, B) */
int bcd_host_exr_write_ex(const char* path, int multi_channel, const float* hwc, int width, int height, int depth,
                          int pixel_type, int compression);

/* bcd::ParametersIO from C (tests). v[11] = {performSpikeRemoval, spikeRemovalThresholdStDevFactor, nbOfScales,
 * histoDistanceThreshold, useCuda, nbOfCores, patchRadius, searchWindowRadius, randomPixelOrder,
 * markedPixelsSkippingProbability, minEigenValue}: on load, in = the values a missing key keeps, out = the result. */
int bcd_host_params_load(const char* path, double* v, char* color_file, char* hist_file, char* cov_file, int cap);
int bcd_host_params_write(const char* path, const double* v, const char* color_file, const char* hist_file, const char* cov_file);

#ifdef __cplusplus
}
#endif
#endif

/* TEST INFRASTRUCTURE — plain-C restatement of the reference hot path (see bcd_oracle.c). */
#ifndef BCD_ORACLE_H
#define BCD_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	float tau;        /* DenoiserParameters::m_histogramDistanceThreshold */
	int   w;          /* m_patchRadius */
	int   b;          /* m_searchWindowRadius */
	float eps;        /* m_minEigenValue */
	int   marking;    /* 1: m_markedPixelsSkippingProbability == 1 (raster order, 1 core); 0: == 0 */
} bcdo_params;

typedef struct {
	long long interior, active, full, fallback, strong, sum_n_full;
} bcdo_stats;

/* A.1 */
void bcdo_pixel_cov(const float* cov, const float* ns, long long npix, float* pixcov);
/* A.2 / A.3 */
float bcdo_pair_term(const float* hx, const float* hy, float nx, float ny, int B, int* cnt);
float bcdo_patch_distance(const float* hist, const float* ns, int W, int H, int B, int w,
		int pl, int pc, int ql, int qc);
/* every pair term of a frame, planes [J][H][W] (see bcd_oracle.c) */
void bcdo_pair_terms_all(const float* hist, const float* ns, int W, int H, int B, int b, float* sums, unsigned char* counts);
/* A.4: masks[(l*W+c)*words + k]: bit index = (dy+b)*(2b+1) + (dx+b), words = ceil((2b+1)^2/32); counts[l*W+c] */
int  bcdo_mask_words(int b);
void bcdo_select_all(const float* hist, const float* ns, int W, int H, int B, int w, int b, float tau,
		uint32_t* masks, int* counts);
/* A.5 closed form: active(p) <=> no q<p (raster), q in S(p), active(q) && |S(q)| >= d+1 */
void bcdo_marking(const uint32_t* masks, const int* counts, int W, int H, int w, int b, uint8_t* active);
/* Eigen::SelfAdjointEigenSolver<MatrixXf>::compute restated; A is n x n column-major (lower triangle read),
   evecs column-major (column k = k-th eigenvector), evals ascending. work: >= n*n + 4*n floats. */
int  bcdo_eigh(const float* A, int n, float* evals, float* evecs, float* work);
/* A.1-A.10 single scale; raster order; out is W*H*3. sum/count (optional, may be NULL) receive the accumulators. */
int  bcdo_denoise(const float* color, const float* ns, const float* hist, const float* cov,
		int W, int H, int B, const bcdo_params* p, float* out, bcdo_stats* stats,
		float* opt_sum, int* opt_count);
/* A.11 */
void bcdo_downscale_sum(const float* in, int W, int H, int D, float* out);
void bcdo_downscale_avg(const float* in, int W, int H, int D, float* out);
void bcdo_downscale_cov(const float* cov, const float* ns, int W, int H, float* out);
void bcdo_interpolate(const float* low, int Wl, int Hl, int D, float* out, int Wu, int Hu);
void bcdo_merge(float* hi, int W, int H, const float* low_out);
int  bcdo_denoise_multiscale(const float* color, const float* ns, const float* hist, const float* cov,
		int W, int H, int B, const bcdo_params* p, int nscales, float* out);

/* row f1: SpikeRemovalFilter (in place); src_index optional */
void bcdo_spike_removal(float* color, float* ns, float* hist, float* cov, int W, int H, int B, float factor, int int_abs,
		int* src_index);

#ifdef __cplusplus
}
#endif
#endif

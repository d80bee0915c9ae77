// Host-side launchers of the bcd_b200 kernels (one per stage of the hot path; see DESIGN.md for the stage map).
#pragma once
#include "common.cuh"

namespace bcdk
{

// accumulators of the data-dependent work counters (device resident)
struct DeviceCounters
{
	unsigned long long interior, strong, active, full, fallback, sum_n_full;
	int nFull, nFallback; // list lengths of the current level
	unsigned int markTimeouts, pad_; // bounded waits of the marking kernel that expired (must stay 0)
};

// Raises cudaFuncAttributeMaxDynamicSharedMemorySize of `kernel` on the current device to at least `bytes` — never lowers
// it, and serialises the callers: several host threads (one per row tile) launch the same kernels with different
// shared-memory sizes, and a plain cudaFuncSetAttribute before every launch lets one thread shrink the limit between
// another thread's attribute call and its launch ("too many resources requested for launch").
cudaError_t allowDynamicSmem(const void* kernel, size_t bytes);
// multiprocessors of the current device (cached; 148 on a B200): persistent grids and residency tests are sized from it
int currentSmCount();
// K0  Denoiser::computePixelCovFromSampleCov
cudaError_t launchPixelCov(const float* cov, const float* ns, size_t npix, float* pixcov, cudaStream_t st, int* launches);
// K1a/b/c
extern bool g_forceDensePair;
extern int g_pairKernel; // 2 = current sparse kernel, 1 = its predecessor (test hook BCD_B200_PAIR=v1)
cudaError_t launchPairTerms(const float* hist, const float* ns, const LevelGeom& g, float* sP, unsigned char* cP,
		cudaStream_t st, int* launches);
cudaError_t launchPatchSelect(const float* sP, const unsigned char* cP, const LevelGeom& g, float tau,
		uint32_t* mplus, cudaStream_t st, int* launches);
cudaError_t launchSymmetrize(const uint32_t* mplus, const LevelGeom& g, uint32_t* smask, int* counts, uint32_t* pre, int prePitch, int fastLayout,
		cudaStream_t st, int* launches, int rowLo = -1, int rowHi = -1); // rows [rowLo, rowHi) only (default: [g.yLo, g.yHi))
// K2  marking strategy in raster order over the owned rows. asBits / actBits: H x Wp bit-rows (common.cuh), zeroed by
// the caller (asBits rows above rowBeg may instead hold the upper row tile's hand-off); progress: >= ceil(rows/32).
extern bool g_forceGenericMarking;
extern int g_markingBandsPerCta;
extern int g_markingTrace;
extern int g_markingKernel; // b = 6: 1 = systolic kernel (default), 0 = shared-memory wavefront (test hook BCD_B200_MARKING=v4)
// K3e for d <= 32: lanes per matrix (16 or 8: several matrices per warp, k_eig_sub.cu; 32: one warp per matrix). Test hook
// BCD_B200_EIG_GROUP.
extern int g_eigGroup;
extern bool g_eigGenericOrder;
cudaError_t launchEigChainSub(float* A, float* Minv, const float* noise, int count, const int* countDev, int first,
		int n, int np, float eps, int doClamp, int group, int smCount, cudaStream_t st);
int markingPreLayout(int b); // what k_symmetrize must write: 0 generic, 1 `pre2` row-major, 2 `pre2` band-major (systolic kernel)
size_t markingPreBytes(int W, int H, int b, int Pw); // capacity of the marking input for any kernel choice
cudaError_t launchMarking(const uint32_t* pre, int prePitch, const LevelGeom& g, float skipProb, uint32_t seed,
		uint32_t* actBits, uint32_t* asBits, int Wp, unsigned int* progress, cudaStream_t st, int* launches,
		unsigned int* timeouts = nullptr, uint32_t epoch = 1, const uint32_t* ringIn = nullptr, uint32_t* ringOut = nullptr);
// (systolic kernel, b = 6) epoch: tag of the hand-off ring words of this launch, different from the previous launches that
// used the same scratch; ringIn / ringOut: inbox fed by the upper row tile / inbox of the lower row tile (markingRingWords)
bool markingUsesRing(const LevelGeom& g); // the systolic kernel serves this geometry (ringIn / ringOut are honoured)
cudaError_t launchMarkingWarmup(cudaStream_t st); // loads the kernels a warm-up frame does not launch
cudaError_t launchAllActive(const LevelGeom& g, uint32_t* actBits, int Wp, cudaStream_t st, int* launches);
// work lists
cudaError_t launchBuildLists(const int* counts, const uint32_t* actBits, int Wp, const LevelGeom& g,
		int* fullList, int* fallbackList, DeviceCounters* ctr, cudaStream_t st, int* launches);
// K3f denoiseOnlyMainPatch
cudaError_t launchFallback(const float* color, const uint32_t* smask, const int* counts, const LevelGeom& g,
		const int* fallbackList, const int* nFallback, float4* acc, cudaStream_t st, int* launches);
// K3 two-step Bayesian filter over a chunk [first, first+count) of the full list
struct FilterBuffers
{
	float* A;      // chunk x d*d   (C - Sigma_n, later C2 + Sigma_n)
	float* Minv;   // chunk x d*d
	float* mean;   // chunk x d
	float* noise;  // chunk x np*6
	float* xScratch; // optional global scratch for X when it does not fit shared memory (gridCap x nsmax*d)
	int gridCap;
	int smCount;     // multiprocessors of the device (persistent grids are sized from it)
};
// count: capacity of the chunk; countDev (optional): device-resident list length that clips it (no host read-back);
// ev (optional): 6 events recorded around the 5 launches (stats | eig | apply 1 | eig | apply 2)
cudaError_t launchFilterChunk(const float* color, const float* pixcov, const uint32_t* smask, const int* counts,
		const LevelGeom& g, float minEig, const int* fullList, int first, int count, const int* countDev, const FilterBuffers& fb,
		float4* acc, cudaStream_t st, int* launches, cudaEvent_t* ev);
// halo strip added into the accumulator (row tiles)
cudaError_t launchAddStrip(float4* acc, const float4* strip, size_t n, cudaStream_t st, int* launches);
// peer-memory strips of the row-tile path (k_peer.cu): push = copy into the receiving GPU's arena + epoch flag there;
// recv = wait for the local flag (bounded), then copy (or add float4) the slot into the image
cudaError_t launchPeerPush(const void* src, void* dstPeer, size_t bytes, uint32_t* flagPeer, uint32_t epoch, unsigned int* ticket,
		cudaStream_t st, int* launches);
cudaError_t launchPeerRecv(void* dst, const void* slot, size_t bytes, bool addFloat4, const uint32_t* flag, uint32_t epoch,
		unsigned int* timeouts, cudaStream_t st, int* launches);
cudaError_t launchPeerWarmup(unsigned int* scratch /* 4 words */, cudaStream_t st);
// K4 finalAggregation
cudaError_t launchFinalize(const float4* acc, size_t npix, float* out, cudaStream_t st, int* launches);
// (row f1) SpikeRemovalFilter: out-of-place; src = W*H ints of scratch
cudaError_t launchSpikeFilter(const float* color, const float* ns, const float* hist, const float* cov, int W, int H, int B,
		float factor, int intAbs, int* src, float* colorOut, float* nsOut, float* histOut, float* covOut, cudaStream_t st, int* launches);
// (row f3, device side) SamplesAccumulator on the GPU: samples [pixel][spp][rgb] (+ optional weights [pixel][spp]) -> statistics
cudaError_t launchAccumulateSamples(const float* samples, const float* weights, size_t npix, int spp, int nbBins, float gamma,
		float maxValue, float* color, float* ns, float* hist, float* cov, cudaStream_t st, int* launches);
// K5 pyramid down, K6 merge
cudaError_t launchDownscale(const float* color, const float* ns, const float* hist, const float* cov, int W, int H, int B,
		float* colorD, float* nsD, float* histD, float* covD, cudaStream_t st, int* launches, int parts = 3,
		int rowBeg = 0, int rows = -1); // rows >= 0: only these rows of the coarser level (parts 1 / 4 / 8)
// parts: 1 = sample counts + histograms, 2 = colour + covariance (reads the FINER level's sample counts), 4 = sample counts
// only, 8 = histograms only
cudaError_t launchMerge(float* hi, int W, int H, const float* lowOut, float* tmpLow, cudaStream_t st, int* launches);

} // namespace bcdk

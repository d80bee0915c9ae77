// Streaming (HBM-bound) stages of the hot path: prepass, work lists, final aggregation, pyramid build and merge.
// Compiled with --fmad=false so that every expression rounds exactly like the reference's scalar C++.
//
// Replaces (reference, paths relative to /root/reference):
//   Denoiser::computePixelCovFromSampleCov       src/core/Denoiser.cpp:357-373            -> k_pixel_cov
//   Denoiser::finalAggregation                   src/core/Denoiser.cpp:434-470            -> k_finalize
//   MultiscaleDenoiser::downscaleSum/Average/SampleCovarianceSum  MultiscaleDenoiser.cpp:243-334 -> k_downscale_*
//   MultiscaleDenoiser::mergeOutputs (downscale, interpolate, lowPass) MultiscaleDenoiser.cpp:453-548 -> k_merge

#include "common.cuh"
#include "kernels.h"
#include <map>
#include <mutex>
#include <utility>

namespace bcdk
{

int currentSmCount()
{
	static int cached[64] = {0};
	int dev = 0;
	if(cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
		return 148;
	if(cached[dev] == 0)
	{
		int sms = 0;
		if(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
			sms = 148;
		cached[dev] = sms;
	}
	return cached[dev];
}

cudaError_t allowDynamicSmem(const void* kernel, size_t bytes)
{
	static std::mutex mu;
	static std::map<std::pair<int, const void*>, size_t> granted; // (device, kernel) -> current limit
	if(bytes <= 48 * 1024)
		return cudaSuccess; // always allowed
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if(e != cudaSuccess) return e;
	std::lock_guard<std::mutex> lk(mu);
	size_t& cur = granted[std::make_pair(dev, kernel)];
	if(bytes <= cur)
		return cudaSuccess;
	e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
	if(e == cudaSuccess) cur = bytes;
	return e;
}

__global__ void k_pixel_cov(const float* __restrict__ cov, const float* __restrict__ ns, size_t npix,
		float* __restrict__ pixcov)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= npix * 6)
		return;
	const size_t p = i / 6;
	const float inv = __fdiv_rn(1.f, ns[p]);
	pixcov[i] = __fmul_rn(cov[i], inv);
}

cudaError_t launchPixelCov(const float* cov, const float* ns, size_t npix, float* pixcov, cudaStream_t st, int* launches)
{
	const size_t n = npix * 6;
	k_pixel_cov<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cov, ns, npix, pixcov);
	*launches += 1;
	return cudaGetLastError();
}

// Work lists of the owned patch centres: full (|S| >= d+1, DenoisingUnit.cpp:182) and fallback.
__global__ void k_build_lists(const int* __restrict__ counts, const uint32_t* __restrict__ actBits, int Wp, LevelGeom g,
		int* __restrict__ fullList, int* __restrict__ fallbackList, DeviceCounters* __restrict__ ctr)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y * blockDim.y + threadIdx.y;
	const bool in = x < g.W && y < g.H;
	const bool interior = in && x >= g.w && x <= g.W - 1 - g.w && y >= g.rowBeg && y < g.rowEnd;
	const size_t p = in ? (size_t)y * g.W + x : 0;
	const int n = interior ? counts[p] : 0;
	const bool strong = interior && n >= g.d + 1;
	const bool act = interior && testRowBit(actBits, Wp, y, x);
	const bool full = act && strong, fb = act && !strong;

	const unsigned lane = (threadIdx.x + blockDim.x * threadIdx.y) & 31;
	const unsigned mFull = __ballot_sync(0xffffffffu, full);
	const unsigned mFb = __ballot_sync(0xffffffffu, fb);
	const unsigned mInt = __ballot_sync(0xffffffffu, interior);
	const unsigned mStrong = __ballot_sync(0xffffffffu, strong);
	const unsigned mAct = __ballot_sync(0xffffffffu, act);
	int nsum = full ? n : 0;
	for(int o = 16; o > 0; o >>= 1)
		nsum += __shfl_xor_sync(0xffffffffu, nsum, o);
	int baseFull = 0, baseFb = 0;
	if(lane == 0)
	{
		if(mFull) baseFull = atomicAdd(&ctr->nFull, __popc(mFull));
		if(mFb) baseFb = atomicAdd(&ctr->nFallback, __popc(mFb));
		if(mInt) atomicAdd(&ctr->interior, (unsigned long long)__popc(mInt));
		if(mStrong) atomicAdd(&ctr->strong, (unsigned long long)__popc(mStrong));
		if(mAct) atomicAdd(&ctr->active, (unsigned long long)__popc(mAct));
		if(mFull) atomicAdd(&ctr->full, (unsigned long long)__popc(mFull));
		if(mFb) atomicAdd(&ctr->fallback, (unsigned long long)__popc(mFb));
		if(nsum) atomicAdd(&ctr->sum_n_full, (unsigned long long)nsum);
	}
	baseFull = __shfl_sync(0xffffffffu, baseFull, 0);
	baseFb = __shfl_sync(0xffffffffu, baseFb, 0);
	const unsigned below = (1u << lane) - 1u;
	if(full)
		fullList[baseFull + __popc(mFull & below)] = (int)p;
	if(fb)
		fallbackList[baseFb + __popc(mFb & below)] = (int)p;
}

cudaError_t launchBuildLists(const int* counts, const uint32_t* actBits, int Wp, const LevelGeom& g,
		int* fullList, int* fallbackList, DeviceCounters* ctr, cudaStream_t st, int* launches)
{
	dim3 block(32, 8), grid(divUp(g.W, 32), divUp(g.H, 8));
	k_build_lists<<<grid, block, 0, st>>>(counts, actBits, Wp, g, fullList, fallbackList, ctr);
	*launches += 1;
	return cudaGetLastError();
}

// row tiles: partial {sum, count} rows accumulated by the neighbouring tile are added to the owner's rows
__global__ void k_add_strip(float4* __restrict__ acc, const float4* __restrict__ strip, size_t n)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n)
		return;
	float4 a = acc[i];
	const float4 s = strip[i];
	a.x += s.x; a.y += s.y; a.z += s.z; a.w += s.w;
	acc[i] = a;
}

cudaError_t launchAddStrip(float4* acc, const float4* strip, size_t n, cudaStream_t st, int* launches)
{
	if(n == 0)
		return cudaSuccess;
	k_add_strip<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(acc, strip, n);
	*launches += 1;
	return cudaGetLastError();
}

// out = (1/count) * sum; count == 0 gives inf * 0 = NaN exactly like Denoiser.cpp:463-466
__global__ void k_finalize(const float4* __restrict__ acc, size_t npix, float* __restrict__ out)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= npix)
		return;
	const float4 a = acc[i];
	const float ci = __fdiv_rn(1.f, a.w);
	out[i * 3 + 0] = __fmul_rn(ci, a.x);
	out[i * 3 + 1] = __fmul_rn(ci, a.y);
	out[i * 3 + 2] = __fmul_rn(ci, a.z);
}

cudaError_t launchFinalize(const float4* acc, size_t npix, float* out, cudaStream_t st, int* launches)
{
	k_finalize<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(acc, npix, out);
	*launches += 1;
	return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------
// (device side) SamplesAccumulator::addSample + computeSampleStatistics —
// src/core/SamplesAccumulator.cpp:44-107,110-141 — for a renderer whose samples already live on the GPU: one thread per
// pixel walks its samples in order (the reference's accumulation order, separate multiply and add: this file is built
// with --fmad=false), so mean, sample count and covariance are the reference's numbers bit for bit; the histogram goes
// through powf (gamma), which differs from glibc's in the last place for some inputs — a 1e-7 shift of weight between
// two neighbouring bins. The per-thread histogram lives in shared memory ([bin][thread]: conflict-free).
// ---------------------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT)
k_accumulate_samples(const float* __restrict__ samples, const float* __restrict__ weights, size_t npix, int spp, int nbBins,
		float gamma, float maxValue, float* __restrict__ color, float* __restrict__ ns, float* __restrict__ hist, float* __restrict__ cov)
{
	extern __shared__ float shHist[]; // [3 * nbBins][NT]
	const int tid = threadIdx.x;
	const size_t p = (size_t)blockIdx.x * NT + tid;
	const int B = 3 * nbBins;
	for(int k = 0; k < B; ++k)
		shHist[k * NT + tid] = 0.f;
	if(p >= npix)
		return;
	const float saturation = 2.f;
	const float invGamma = __fdiv_rn(1.f, gamma);
	float wSum = 0.f, w2Sum = 0.f, sum[3] = {0.f, 0.f, 0.f}, c2[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
	const float* sp = samples + p * (size_t)spp * 3;
	for(int k = 0; k < spp; ++k, sp += 3)
	{
		const float w = weights ? weights[p * (size_t)spp + k] : 1.f;
		const float rgb[3] = {sp[0], sp[1], sp[2]};
		wSum = __fadd_rn(wSum, w);
		w2Sum = __fadd_rn(w2Sum, __fmul_rn(w, w));
		for(int c = 0; c < 3; ++c)
			sum[c] = __fadd_rn(sum[c], __fmul_rn(w, rgb[c]));
		// xx, yy, zz, yz, xz, xy: (w * a) * b like the reference's left-to-right products
		c2[0] = __fadd_rn(c2[0], __fmul_rn(__fmul_rn(w, rgb[0]), rgb[0]));
		c2[1] = __fadd_rn(c2[1], __fmul_rn(__fmul_rn(w, rgb[1]), rgb[1]));
		c2[2] = __fadd_rn(c2[2], __fmul_rn(__fmul_rn(w, rgb[2]), rgb[2]));
		c2[3] = __fadd_rn(c2[3], __fmul_rn(__fmul_rn(w, rgb[1]), rgb[2]));
		c2[4] = __fadd_rn(c2[4], __fmul_rn(__fmul_rn(w, rgb[0]), rgb[2]));
		c2[5] = __fadd_rn(c2[5], __fmul_rn(__fmul_rn(w, rgb[0]), rgb[1]));
		for(int c = 0; c < 3; ++c)
		{
			float v = rgb[c] > 0.f ? rgb[c] : 0.f;
			if(gamma > 1.f) v = powf(v, invGamma);
			if(maxValue > 0.f) v = __fdiv_rn(v, maxValue);
			v = v > saturation ? saturation : v;
			const float position = __fmul_rn(v, (float)(nbBins - 2));
			int lowBin = (int)position;
			float highWeight;
			if(lowBin < nbBins - 2)
				highWeight = __fsub_rn(position, (float)lowBin);
			else
			{
				lowBin = nbBins - 2;
				highWeight = __fdiv_rn(__fsub_rn(v, 1.f), __fsub_rn(saturation, 1.f));
			}
			const float lowWeight = __fsub_rn(1.f, highWeight);
			float* hb = shHist + (c * nbBins + lowBin) * NT + tid;
			hb[0] = __fadd_rn(hb[0], __fmul_rn(w, lowWeight));
			hb[NT] = __fadd_rn(hb[NT], __fmul_rn(w, highWeight));
		}
	}
	// computeSampleStatistics (:110-141)
	const float invW = __fdiv_rn(1.f, wSum);
	float mean[3];
	for(int c = 0; c < 3; ++c)
	{
		mean[c] = __fmul_rn(invW, sum[c]);
		color[p * 3 + c] = mean[c];
	}
	ns[p] = wSum;
	const float unbias = __fdiv_rn(1.f, __fsub_rn(1.f, __fdiv_rn(w2Sum, __fmul_rn(wSum, wSum))));
	const int ia[6] = {0, 1, 2, 1, 0, 0}, ib[6] = {0, 1, 2, 2, 2, 1};
	for(int e = 0; e < 6; ++e)
	{
		float cv = __fmul_rn(c2[e], invW);
		cv = __fsub_rn(cv, __fmul_rn(mean[ia[e]], mean[ib[e]]));
		cov[p * 6 + e] = __fmul_rn(cv, unbias);
	}
	for(int k = 0; k < B; ++k)
		hist[p * B + k] = shHist[k * NT + tid];
}

cudaError_t launchAccumulateSamples(const float* samples, const float* weights, size_t npix, int spp, int nbBins, float gamma,
		float maxValue, float* color, float* ns, float* hist, float* cov, cudaStream_t st, int* launches)
{
	constexpr int NT = 64;
	const size_t smem = (size_t)3 * nbBins * NT * sizeof(float);
	cudaError_t err = allowDynamicSmem((const void*)k_accumulate_samples<NT>, smem);
	if(err != cudaSuccess) return err;
	k_accumulate_samples<NT><<<(unsigned)((npix + NT - 1) / NT), NT, smem, st>>>(samples, weights, npix, spp, nbBins, gamma, maxValue,
			color, ns, hist, cov);
	*launches += 1;
	return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
//  SpikeRemovalFilter::filter — src/core/SpikeRemovalFilter.cpp:18-75 with :77-92 (average and
// standard deviation of the 3x3 neighbourhood, sequential sums) and :94-116 (L1 median of the 9 colours).
// k_spike_source decides, per pixel, which pixel it is copied from (itself, or the median neighbour of its clamped
// 3x3 window when it is an outlier); k_spike_gather then copies colour, nSamples, histogram and covariance.
// intAbs = 1 reproduces the overload the reference's unqualified abs() resolves to under GCC/glibc (::abs(int), i.e.
// the float is truncated first — what the compiled oracle does); 0 is the float overload (MSVC, the intent).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float absMode(float v, int intAbs)
{
	if(intAbs)
	{
		const int t = (int)v; // C conversion: truncation toward zero (cvt.rzi)
		return (float)(t < 0 ? -t : t);
	}
	return fabsf(v);
}

__global__ void k_spike_source(const float* __restrict__ color, int W, int H, float factor, int intAbs, int* __restrict__ src)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y * blockDim.y + threadIdx.y;
	if(x >= W || y >= H)
		return;
	const int cl = y < 1 ? 1 : (y > H - 2 ? H - 2 : y);
	const int cc = x < 1 ? 1 : (x > W - 2 ? W - 2 : x);
	float v[3][9];
	for(int k = 0; k < 9; ++k)
	{
		const float* p = color + ((size_t)(cl - 1 + k / 3) * W + (cc - 1 + k % 3)) * 3;
		v[0][k] = p[0]; v[1][k] = p[1]; v[2][k] = p[2];
	}
	const float* me = color + ((size_t)y * W + x) * 3;
	bool spike = false;
	for(int c = 0; c < 3; ++c)
	{
		float total = 0.f;
		for(int k = 0; k < 9; ++k) total = total + v[c][k];
		const float avg = total / 9;
		total = 0.f;
		for(int k = 0; k < 9; ++k) total = total + (v[c][k] - avg) * (v[c][k] - avg);
		const float sd = sqrtf(total / 8);
		spike = spike || absMode(me[c] - avg, intAbs) > factor * sd;
	}
	int source = y * W + x;
	if(spike)
	{
		int median = 0;
		float best = -1.f;
		for(int cand = 0; cand < 9; ++cand)
		{
			float tot = 0.f;
			for(int k = 0; k < 9; ++k)
			{
				if(intAbs)
				{
					const int a0 = (int)(v[0][k] - v[0][cand]), a1 = (int)(v[1][k] - v[1][cand]), a2 = (int)(v[2][k] - v[2][cand]);
					tot = tot + (float)((a0 < 0 ? -a0 : a0) + (a1 < 0 ? -a1 : a1) + (a2 < 0 ? -a2 : a2));
				}
				else
					tot = tot + (fabsf(v[0][k] - v[0][cand]) + fabsf(v[1][k] - v[1][cand]) + fabsf(v[2][k] - v[2][cand]));
			}
			if(best < 0 || tot < best)
			{
				best = tot;
				median = cand;
			}
		}
		source = (cl - 1 + median / 3) * W + (cc - 1 + median % 3);
	}
	src[(size_t)y * W + x] = source;
}

__global__ void k_spike_gather(const float* __restrict__ in, const int* __restrict__ src, size_t npix, int D, float* __restrict__ out)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= npix * D)
		return;
	const size_t p = i / D;
	const int z = (int)(i - p * D);
	out[i] = in[(size_t)src[p] * D + z];
}

cudaError_t launchSpikeFilter(const float* color, const float* ns, const float* hist, const float* cov, int W, int H, int B,
		float factor, int intAbs, int* src, float* colorOut, float* nsOut, float* histOut, float* covOut, cudaStream_t st, int* launches)
{
	dim3 block(32, 8), grid(divUp(W, 32), divUp(H, 8));
	k_spike_source<<<grid, block, 0, st>>>(color, W, H, factor, intAbs, src);
	const size_t npix = (size_t)W * H;
	auto blocks = [](size_t n) { return (unsigned)((n + 255) / 256); };
	k_spike_gather<<<blocks(npix * 3), 256, 0, st>>>(color, src, npix, 3, colorOut);
	k_spike_gather<<<blocks(npix), 256, 0, st>>>(ns, src, npix, 1, nsOut);
	k_spike_gather<<<blocks(npix * B), 256, 0, st>>>(hist, src, npix, B, histOut);
	k_spike_gather<<<blocks(npix * 6), 256, 0, st>>>(cov, src, npix, 6, covOut);
	*launches += 5;
	return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// pyramid
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int clampi(int v, int n) { return v <= 0 ? 0 : (v >= n ? n - 1 : v); }

// mode 0: sum (histogram, nSamples)  1: 0.25 * sum (colour)
// rows [rowBeg, rowBeg + rows) of the downscaled image
__global__ void k_downscale(const float* __restrict__ in, int W, int H, int D, int mode, float* __restrict__ out,
		int rowBeg, int rows)
{
	const int Wd = W / 2;
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= (size_t)Wd * rows * D)
		return;
	i += (size_t)rowBeg * Wd * D;
	const int z = (int)(i % D);
	const size_t pix = i / D;
	const int c = (int)(pix % Wd), l = (int)(pix / Wd);
	const int l1 = clampi(2 * l + 1, H), c1 = clampi(2 * c + 1, W);
	const float v1 = in[((size_t)(2 * l) * W + 2 * c) * D + z];
	const float v2 = in[((size_t)l1 * W + 2 * c) * D + z];
	const float v3 = in[((size_t)(2 * l) * W + c1) * D + z];
	const float v4 = in[((size_t)l1 * W + c1) * D + z];
	const float s = v1 + v2 + v3 + v4;
	out[i] = mode ? 0.25f * s : s;
}

__global__ void k_downscale_cov(const float* __restrict__ cov, const float* __restrict__ ns, int W, int H,
		float* __restrict__ out)
{
	const int Wd = W / 2, Hd = H / 2;
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= (size_t)Wd * Hd * 6)
		return;
	const int z = (int)(i % 6);
	const size_t pix = i / 6;
	const int c = (int)(pix % Wd), l = (int)(pix / Wd);
	const int l1 = clampi(2 * l + 1, H), c1 = clampi(2 * c + 1, W);
	const size_t i1 = (size_t)(2 * l) * W + 2 * c, i2 = (size_t)l1 * W + 2 * c;
	const size_t i3 = (size_t)(2 * l) * W + c1, i4 = (size_t)l1 * W + c1;
	const float n1 = ns[i1], n2 = ns[i2], n3 = ns[i3], n4 = ns[i4];
	const float nSum = n1 + n2 + n3 + n4;
	const float sq = (1.f / 4.f) * (1.f / 4.f);
	const float w1 = sq * nSum / n1, w2 = sq * nSum / n2, w3 = sq * nSum / n3, w4 = sq * nSum / n4;
	out[i] = w1 * cov[i1 * 6 + z] + w2 * cov[i2 * 6 + z] + w3 * cov[i3 * 6 + z] + w4 * cov[i4 * 6 + z];
}

cudaError_t launchDownscale(const float* color, const float* ns, const float* hist, const float* cov, int W, int H, int B,
		float* colorD, float* nsD, float* histD, float* covD, cudaStream_t st, int* launches, int parts, int rowBeg, int rows)
{
	if(rows < 0) { rowBeg = 0; rows = H / 2; } // the whole level (only part 1 takes a row range)
	const size_t npd = (size_t)(W / 2) * (H / 2), npr = (size_t)(W / 2) * rows;
	if(npd == 0 || npr == 0)
		return cudaSuccess;
	auto blocks = [](size_t n) { return (unsigned)((n + 255) / 256); };
	if(parts & (1 | 4)) // sample counts: read by the selection stage of the coarser level and by the covariance part below
	{
		k_downscale<<<blocks(npr), 256, 0, st>>>(ns, W, H, 1, 0, nsD, rowBeg, rows);
		*launches += 1;
	}
	if(parts & (1 | 8)) // histograms: what the selection stage of the coarser level reads
	{
		k_downscale<<<blocks(npr * B), 256, 0, st>>>(hist, W, H, B, 0, histD, rowBeg, rows);
		*launches += 1;
	}
	if(parts & 2) // what its filters read
	{
		k_downscale<<<blocks(npd * 3), 256, 0, st>>>(color, W, H, 3, 1, colorD, 0, H / 2);
		k_downscale_cov<<<blocks(npd * 6), 256, 0, st>>>(cov, ns, W, H, covD);
		*launches += 2;
	}
	return cudaGetLastError();
}

// interpolate (MultiscaleDenoiser.cpp:473-512) evaluated at one upscaled sample
__device__ __forceinline__ float interpAt(const float* __restrict__ low, int Wl, int Hl, int L, int C, int z)
{
	const float mw = 9.f / 16.f, aw = 3.f / 16.f, dw = 1.f / 16.f;
	const int l = L / 2, c = C / 2;
	const int al = clampi(l + ((L % 2) * 2 - 1), Hl);
	const int ac = clampi(c + ((C % 2) * 2 - 1), Wl);
	const int lc = clampi(l, Hl), cc = clampi(c, Wl);
	const float p1 = low[((size_t)lc * Wl + cc) * 3 + z];
	const float p2 = low[((size_t)lc * Wl + ac) * 3 + z];
	const float p3 = low[((size_t)al * Wl + cc) * 3 + z];
	const float p4 = low[((size_t)al * Wl + ac) * 3 + z];
	return mw * p1 + aw * (p2 + p3) + dw * p4;
}

// hi <- (hi - interpolate(downscale(hi))) + interpolate(lowOut)   (mergeOutputs :453-466)
__global__ void k_merge(float* __restrict__ hi, int W, int H, const float* __restrict__ tmpLow,
		const float* __restrict__ lowOut)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= (size_t)W * H * 3)
		return;
	const int z = (int)(i % 3);
	const size_t pix = i / 3;
	const int C = (int)(pix % W), L = (int)(pix / W);
	const int Wl = W / 2, Hl = H / 2;
	float v = hi[i];
	v = v - interpAt(tmpLow, Wl, Hl, L, C, z);
	v = v + interpAt(lowOut, Wl, Hl, L, C, z);
	hi[i] = v;
}

cudaError_t launchMerge(float* hi, int W, int H, const float* lowOut, float* tmpLow, cudaStream_t st, int* launches)
{
	const size_t npd = (size_t)(W / 2) * (H / 2);
	if(npd == 0)
		return cudaSuccess;
	k_downscale<<<(unsigned)((npd * 3 + 255) / 256), 256, 0, st>>>(hi, W, H, 3, 1, tmpLow, 0, H / 2);
	k_merge<<<(unsigned)(((size_t)W * H * 3 + 255) / 256), 256, 0, st>>>(hi, W, H, tmpLow, lowOut);
	*launches += 2;
	return cudaGetLastError();
}

} // namespace bcdk

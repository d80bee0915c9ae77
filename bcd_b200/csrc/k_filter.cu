// K3 — the collaborative filter applied to every active patch centre of a level.
//
// Replaces (reference, paths relative to /root/reference):
//   DenoisingUnit::denoiseOnlyMainPatch            src/core/DenoisingUnit.cpp:455-481   -> k_fallback
//   DenoisingUnit::computeNoiseCovPatchesMean      :400-419                             -> k_filter_stats
//   pickColorPatchesFromColorImage / empiricalMean / centerPointCloud / empiricalCovarianceMatrix :483-536
//                                                                                       -> k_filter_stats, k_filter_apply<1>
//   substract/addCovMatPatchToMatrix :538-576, clampNegativeEigenValues :606-630, inverseSymmetricMatrix :578-604
//     with Eigen::SelfAdjointEigenSolver<MatrixXf>::compute (ext/eigen/Eigen/src/Eigenvalues/
//     SelfAdjointEigenSolver.h:395-445,486-548,808-864; Tridiagonalization.h:346-378,427-441)
//                                                                                       -> k_eig_chain
//   finalDenoisingMatrixMultiplication :656-670, multiplyCovMatPatchByVector :633-654   -> k_filter_apply<1|2>
//   denoiseSelectedPatchesStep1/Step2 :421-453, aggregateOutputPatches :672-693         -> the 5-kernel chain below
//
// Per "full" pixel (|S(p)| >= d+1) the chain is
//   stats  : X (n x d gather), Sigma_n, mu, C = cov(X) - Sigma_n                       -> A, mean, noise   (global)
//   eig    : Minv = inverse( clamp+(A) + Sigma_n )                 (two eigendecompositions, matrix stays in smem)
//   apply 1: Y_i = X_i - Sigma_n Minv (X_i - mu);  mu2, A = cov(Y) + Sigma_n           -> A, mean
//   eig    : Minv = inverse(A)
//   apply 2: Z_i = X_i - Sigma_n Minv (X_i - mu2); sum[q_i + o] += Z_i, count += 1     (vector float4 atomics)
// The d x d matrices (d = 3(2w+1)^2 = 27 / 75 / 147) live in shared memory; X lives in shared memory when it fits
// and in a per-CTA global scratch slot otherwise. Sums over the similar patches run in the reference's order
// (ascending window index) with separate multiply and add (this file is compiled with --fmad=false).

#include "common.cuh"
#include "kernels.h"
#include <cfloat>
#include <algorithm>

namespace bcdk
{

enum { XX = 0, YY = 1, ZZ = 2, YZ = 3, XZ = 4, XY = 5 }; // reference CovarianceMatrix.h:18-27

// element (a, b) of the symmetric 3x3 block stored as xx,yy,zz,yz,xz,xy
__device__ __forceinline__ int symIndex(int a, int b)
{
	if(a == b) return a;            // xx, yy, zz
	const int s = a + b;            // (1,2)->3 yz ; (0,2)->2 xz ; (0,1)->1 xy
	return s == 3 ? YZ : (s == 2 ? XZ : XY);
}

template <int NT>
__device__ __forceinline__ float blockReduceSum(float v, float* red)
{
	for(int o = 16; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	if(NT == 32)
		return v;
	__syncthreads();
	if((threadIdx.x & 31) == 0)
		red[threadIdx.x >> 5] = v;
	__syncthreads();
	float t = 0.f;
	for(int i = 0; i < NT / 32; ++i)
		t += red[i];
	return t;
}

template <int NT>
__device__ __forceinline__ float blockReduceMax(float v, float* red)
{
	for(int o = 16; o > 0; o >>= 1)
		v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
	if(NT == 32)
		return v;
	__syncthreads();
	if((threadIdx.x & 31) == 0)
		red[threadIdx.x >> 5] = v;
	__syncthreads();
	float t = red[0];
	for(int i = 1; i < NT / 32; ++i)
		t = fmaxf(t, red[i]);
	return t;
}

// Expand the ordered similar set S(p) into the list of the similar patch centres' pixel indices (ascending window
// index = the order selectSimilarPatches pushes the centres). words: smem [Nw], prefix: smem [Nw+1]. Returns n.
template <int NT>
__device__ int expandSimilarList(const uint32_t* __restrict__ smaskP, const LevelGeom& g, int px, int py, uint32_t* words,
		int* prefix, int* qbase)
{
	const int tid = threadIdx.x;
	__syncthreads();
	for(int i = tid; i < g.Nw; i += NT)
		words[i] = smaskP[i];
	__syncthreads();
	if(tid == 0)
	{
		int acc = 0;
		for(int i = 0; i < g.Nw; ++i)
		{
			prefix[i] = acc;
			acc += __popc(words[i]);
		}
		prefix[g.Nw] = acc;
	}
	__syncthreads();
	for(int idx = tid; idx < g.Ns; idx += NT)
	{
		const uint32_t wv = words[idx >> 5];
		const int bit = idx & 31;
		if((wv >> bit) & 1u)
		{
			const int dy = idx / g.side, dx = idx - dy * g.side;
			qbase[prefix[idx >> 5] + __popc(wv & ((1u << bit) - 1u))] = (py + dy - g.b) * g.W + (px + dx - g.b);
		}
	}
	__syncthreads();
	return prefix[g.Nw];
}

// X[i][jj] = colour of patch pixel k = jj/3 (row-major) channel jj%3 of the i-th similar patch (:483-498).
// Thread t owns component jj = t % d for the patches i = t / d, t / d + NT / d, ... (no division in the loop).
template <int NT>
__device__ void gatherPatches(const float* __restrict__ color, const LevelGeom& g, const int* qbase, int n, float* X)
{
	const int d = g.d, pw = 2 * g.w + 1;
	const int groups = NT / d;
	const int jj = threadIdx.x % d, i0 = threadIdx.x / d;
	if(i0 >= groups)
		return;
	const int k = jj / 3, ch = jj - 3 * k;
	// element index in 32-bit unsigned arithmetic (W * H <= 2^30, checkImages: 3 W H < 2^32; a negative offset wraps back)
	const unsigned off = (unsigned)(((k / pw - g.w) * g.W + (k % pw - g.w)) * 3 + ch);
	// the gather is latency-bound (L2 hits behind an index load): four independent loads in flight per thread
	int i = i0;
	for(; i + 3 * groups < n; i += 4 * groups)
	{
		const unsigned e0 = (unsigned)qbase[i] * 3u + off, e1 = (unsigned)qbase[i + groups] * 3u + off;
		const unsigned e2 = (unsigned)qbase[i + 2 * groups] * 3u + off, e3 = (unsigned)qbase[i + 3 * groups] * 3u + off;
		const float v0 = __ldg(color + e0), v1 = __ldg(color + e1), v2 = __ldg(color + e2), v3 = __ldg(color + e3);
		X[i * d + jj] = v0;
		X[(i + groups) * d + jj] = v1;
		X[(i + 2 * groups) * d + jj] = v2;
		X[(i + 3 * groups) * d + jj] = v3;
	}
	for(; i < n; i += groups)
		X[i * d + jj] = __ldg(color + ((unsigned)qbase[i] * 3u + off));
}

// ---------------------------------------------------------------------------------------------------------------
// K3f: fallback — mean of the similar colour patches added to the main patch only. One warp per pixel.
// ---------------------------------------------------------------------------------------------------------------
template <int MAXT>
__global__ void __launch_bounds__(128)
k_fallback(const float* __restrict__ color, const uint32_t* __restrict__ smask, const int* __restrict__ counts,
		LevelGeom g, const int* __restrict__ list, const int* __restrict__ nList, float4* __restrict__ acc)
{
	const int lane = threadIdx.x & 31;
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int nwarps = (gridDim.x * blockDim.x) >> 5;
	const int total = *nList;
	const int d = g.d, pw = 2 * g.w + 1;
	for(int s = warp; s < total; s += nwarps)
	{
		const int p = list[s];
		const int py = p / g.W, px = p - py * g.W;
		float a[MAXT];
		int off[MAXT]; // offset of (patch pixel, channel) relative to the patch centre in the colour image
#pragma unroll
		for(int t = 0; t < MAXT; ++t)
		{
			a[t] = 0.f;
			const int jj = lane + 32 * t;
			const int k = jj / 3, ch = jj - 3 * k;
			off[t] = ((k / pw - g.w) * g.W + (k % pw - g.w)) * 3 + ch;
		}
		for(int wd = 0; wd < g.Nw; ++wd)
		{
			uint32_t bits = smask[(size_t)p * g.Nw + wd];
			while(bits)
			{
				const int bit = __ffs(bits) - 1;
				bits &= bits - 1;
				const int idx = wd * 32 + bit;
				const int qy = py + idx / g.side - g.b, qx = px + idx % g.side - g.b;
				const float* base = color + ((size_t)qy * g.W + qx) * 3;
#pragma unroll
				for(int t = 0; t < MAXT; ++t)
					if(lane + 32 * t < d)
						a[t] = __fadd_rn(a[t], __ldg(base + off[t]));
			}
		}
		const float invn = __fdiv_rn(1.f, (float)counts[p]); // n == 0 -> inf, inf * 0 = NaN like the reference
#pragma unroll
		for(int t = 0; t < MAXT; ++t)
		{
			const int jj = lane + 32 * t;
			if(jj < d)
			{
				const int k = jj / 3, ch = jj - 3 * k;
				const size_t x = (size_t)(py + k / pw - g.w) * g.W + (px + k % pw - g.w);
				float* dst = reinterpret_cast<float*>(acc + x);
				atomicAdd(dst + ch, __fmul_rn(invn, a[t]));
				if(ch == 0)
					atomicAdd(dst + 3, 1.f);
			}
		}
	}
}

cudaError_t launchFallback(const float* color, const uint32_t* smask, const int* counts, const LevelGeom& g,
		const int* fallbackList, const int* nFallback, float4* acc, cudaStream_t st, int* launches)
{
	const int blocks = currentSmCount() * 16;
	*launches += 1;
	if(g.d <= 32)
		k_fallback<1><<<blocks, 128, 0, st>>>(color, smask, counts, g, fallbackList, nFallback, acc);
	else if(g.d <= 96)
		k_fallback<3><<<blocks, 128, 0, st>>>(color, smask, counts, g, fallbackList, nFallback, acc);
	else if(g.d <= 160)
		k_fallback<5><<<blocks, 128, 0, st>>>(color, smask, counts, g, fallbackList, nFallback, acc);
	else
		return cudaErrorInvalidValue;
	return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// shared-memory carve-up common to the stats / apply kernels
// ---------------------------------------------------------------------------------------------------------------
struct FilterSmem
{
	float* X;             // [nmax][d] (or global scratch)
	float* mean;          // [d]
	float* noise;         // [np*6]
	float* Minv;          // [d*d]  (apply kernels)
	float* vec;           // [NG][3][d] float4 per-group vectors (apply kernels), 16-byte aligned
	float* red;           // [32]
	uint32_t* words;      // [Nw]
	int* prefix;          // [Nw+1]
	int* qbase;           // [Ns] pixel index of every similar patch centre
};

static size_t filterSmemBytes(const LevelGeom& g, bool xInSmem, bool withMinv, int ng)
{
	size_t f = 0;
	if(xInSmem) f += (size_t)g.Ns * g.d;
	f += g.d + g.np * 6 + 32;
	if(withMinv) f += (size_t)g.d * g.d + (size_t)ng * 12 * g.d + 4;
	size_t bytes = f * 4 + (size_t)g.Nw * 4 + (size_t)(g.Nw + 1) * 4 + (size_t)g.Ns * 4;
	return (bytes + 15) & ~(size_t)15;
}

__device__ __forceinline__ FilterSmem carveFilterSmem(unsigned char* raw, const LevelGeom& g, float* xGlobal,
		bool withMinv, int ng)
{
	FilterSmem s;
	float* f = reinterpret_cast<float*>(raw);
	if(xGlobal) s.X = xGlobal; else { s.X = f; f += (size_t)g.Ns * g.d; }
	s.mean = f; f += g.d;
	s.noise = f; f += g.np * 6;
	s.red = f; f += 32;
	s.Minv = nullptr; s.vec = nullptr;
	if(withMinv)
	{
		s.Minv = f; f += (size_t)g.d * g.d;
		f = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(f) + 15) & ~(uintptr_t)15); // float4 vectors
		s.vec = f; f += (size_t)ng * 12 * g.d;
	}
	s.words = reinterpret_cast<uint32_t*>(f);
	s.prefix = reinterpret_cast<int*>(s.words + g.Nw);
	s.qbase = s.prefix + g.Nw + 1;
	return s;
}

// mean over i (sequential, :500-509), centring in place (:511-520) and the lower triangle of the empirical
// covariance (:522-536; products summed in patch order, then scaled by 1/(n-1)). sgn*Sigma_n is added on the
// 3x3 diagonal blocks (:538-576) and the symmetric result is written to Aout (column-major d x d).
// DT = compile-time d (27) selects the register-tiled covariance: thread (group, tile) accumulates one 3 x 3 tile of
// the lower triangle (45 tiles for d = 27) over its share of the patches — 6 shared loads feed 9 FMAs instead of 2 per
// FMA — and the NT/45 groups are reduced through the (then dead) patch matrix X. DT = 0: one element per thread.
template <int NT, int DT>
__device__ void meanCenterCov(float* X, int n, const LevelGeom& g, float* mean, const float* noise, float sgn,
		float* __restrict__ Aout, float* __restrict__ meanOut)
{
	const int d = g.d, tid = threadIdx.x;
	const float invn = __fdiv_rn(1.f, (float)n);
	for(int jj = tid; jj < d; jj += NT)
	{
		float a = 0.f;
		for(int i = 0; i < n; ++i)
			a = __fadd_rn(a, X[i * d + jj]);
		a = __fmul_rn(a, invn);
		mean[jj] = a;
		meanOut[jj] = a;
	}
	__syncthreads();
	{
		const int groups = NT / d, jj = tid % d, i0 = tid / d;
		if(i0 < groups)
		{
			const float m = mean[jj];
			for(int i = i0; i < n; i += groups)
				X[i * d + jj] = __fsub_rn(X[i * d + jj], m);
		}
	}
	__syncthreads();
	const float invn1 = __fdiv_rn(1.f, (float)(n - 1));
	if constexpr(DT > 0)
	{
		constexpr int T = DT / 3, LT = T * (T + 1) / 2, NGC = NT / LT;
		static_assert(DT % 3 == 0 && NGC >= 1 && (NGC - 1) * LT * 9 <= 28 * DT, "tile partials must fit the first 28 rows of X");
		const int grp = tid / LT, tile = tid - grp * LT;
		int C = 0, R = tile;
		while(R >= T - C) { R -= T - C; ++C; }
		R += C; // R >= C
		float acc[9];
#pragma unroll
		for(int k = 0; k < 9; ++k) acc[k] = 0.f;
		if(grp < NGC)
		{
			const int i0 = (n * grp) / NGC, i1 = (n * (grp + 1)) / NGC;
			const float* xr = X + (size_t)i0 * DT + 3 * R;
			const float* xc = X + (size_t)i0 * DT + 3 * C;
			for(int i = i0; i < i1; ++i, xr += DT, xc += DT)
			{
				const float r0 = xr[0], r1 = xr[1], r2 = xr[2], c0 = xc[0], c1 = xc[1], c2 = xc[2];
				acc[0] = fmaf(r0, c0, acc[0]); acc[1] = fmaf(r0, c1, acc[1]); acc[2] = fmaf(r0, c2, acc[2]);
				acc[3] = fmaf(r1, c0, acc[3]); acc[4] = fmaf(r1, c1, acc[4]); acc[5] = fmaf(r1, c2, acc[5]);
				acc[6] = fmaf(r2, c0, acc[6]); acc[7] = fmaf(r2, c1, acc[7]); acc[8] = fmaf(r2, c2, acc[8]);
			}
		}
		if(NGC > 1)
		{
			__syncthreads(); // every thread is done reading X: its first rows become the reduction scratch
			if(grp >= 1 && grp < NGC)
#pragma unroll
				for(int k = 0; k < 9; ++k) X[((grp - 1) * LT + tile) * 9 + k] = acc[k];
			__syncthreads();
		}
		if(grp == 0)
		{
			for(int g2 = 1; g2 < NGC; ++g2)
#pragma unroll
				for(int k = 0; k < 9; ++k) acc[k] = __fadd_rn(acc[k], X[((g2 - 1) * LT + tile) * 9 + k]);
#pragma unroll
			for(int a = 0; a < 3; ++a)
#pragma unroll
				for(int b3 = 0; b3 < 3; ++b3)
				{
					const int r = 3 * R + a, c = 3 * C + b3;
					float v = __fmul_rn(acc[a * 3 + b3], invn1);
					if(R == C)
						v = __fadd_rn(v, __fmul_rn(sgn, noise[R * 6 + symIndex(a, b3)]));
					Aout[r + (size_t)c * DT] = v;
					Aout[c + (size_t)r * DT] = v;
				}
		}
		return;
	}
	const int L = d * (d + 1) / 2;
	for(int e = tid; e < L; e += NT)
	{
		int c = 0, r = e;
		while(r >= d - c) { r -= d - c; ++c; }
		r += c; // r >= c
		float a = 0.f;
		for(int i = 0; i < n; ++i)
			a = fmaf(X[i * d + r], X[i * d + c], a); // products summed in patch order (fused: rounding-level difference)
		a = __fmul_rn(a, invn1);
		if(r / 3 == c / 3)
			a = __fadd_rn(a, __fmul_rn(sgn, noise[(r / 3) * 6 + symIndex(r % 3, c % 3)]));
		Aout[r + (size_t)c * d] = a;
		Aout[c + (size_t)r * d] = a;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// K3a: stats
// ---------------------------------------------------------------------------------------------------------------
template <int NT, int DT>
__global__ void __launch_bounds__(NT)
k_filter_stats(const float* __restrict__ color, const float* __restrict__ pixcov, const uint32_t* __restrict__ smask,
		LevelGeom g, const int* __restrict__ list, int first, int count, const int* __restrict__ countDev,
		float* __restrict__ A, float* __restrict__ meanG, float* __restrict__ noiseG, float* xScratch)
{
	extern __shared__ __align__(16) unsigned char smemRaw[];
	const int d = g.d, np = g.np, pw = 2 * g.w + 1, tid = threadIdx.x;
	count = chunkCount(count, countDev, first);
	FilterSmem s = carveFilterSmem(smemRaw, g, xScratch ? xScratch + (size_t)blockIdx.x * g.Ns * d : nullptr, false, 0);
	for(int slot = blockIdx.x; slot < count; slot += gridDim.x)
	{
		const int p = list[first + slot];
		const int py = p / g.W, px = p - py * g.W;
		const int n = expandSimilarList<NT>(smask + (size_t)p * g.Nw, g, px, py, s.words, s.prefix, s.qbase);
		gatherPatches<NT>(color, g, s.qbase, n, s.X);
		// Sigma_n[k] = (1/n) sum_i pixelCov(q_i + o_k)   (:400-419)
		const float invn = __fdiv_rn(1.f, (float)n);
		for(int e = tid; e < np * 6; e += NT)
		{
			const int k = e / 6, el = e - 6 * k;
			const int off6 = ((k / pw - g.w) * g.W + (k % pw - g.w)) * 6 + el;
			float a = 0.f;
			for(int i = 0; i < n; ++i)
				a = __fadd_rn(a, __ldg(pixcov + (size_t)s.qbase[i] * 6 + off6));
			a = __fmul_rn(a, invn);
			s.noise[e] = a;
			noiseG[(size_t)slot * np * 6 + e] = a;
		}
		__syncthreads();
		meanCenterCov<NT, DT>(s.X, n, g, s.mean, s.noise, -1.f, A + (size_t)slot * d * d, meanG + (size_t)slot * d);
		__syncthreads();
	}
}

// ---------------------------------------------------------------------------------------------------------------
// K3b / K3c: apply.  STEP 1: Y in place, then mu2 and A = cov(Y) + Sigma_n.  STEP 2: Z scattered to the accumulator.
// Groups of GS threads handle one similar patch at a time: t = -(Minv (X_i - mean)); out = Sigma_n t + X_i.
// ---------------------------------------------------------------------------------------------------------------
// DT = compile-time d (27): lane r of a group keeps row r of Minv in registers for the whole pixel (the products
// then need one broadcast LDS.128 of the 4 centred patches per 4 FMAs and no Minv traffic); DT = 0: Minv in shared memory.
// Synchronisation among the GS threads that share one xcv4 / tv4 / zv4 scratch: a group of 32 is a warp (the d = 27
// configuration), which lets the four warps of a CTA run through their patches independently instead of in lock-step;
// wider groups span warps and need the CTA barrier.
template <int GS>
__device__ __forceinline__ void groupSync()
{
	if(GS == 32) __syncwarp();
	else __syncthreads();
}

template <int NT, int GS, int STEP, int DT>
// d = 27: 72 registers (7 CTAs of 128 threads per SM; without the bound step 1 takes 96 and only 5 fit)
__global__ void __launch_bounds__(NT, DT > 0 ? 7 : 1)
k_filter_apply(const float* __restrict__ color, const uint32_t* __restrict__ smask, LevelGeom g,
		const int* __restrict__ list, int first, int count, const int* __restrict__ countDev,
		float* __restrict__ A, const float* __restrict__ MinvG, float* __restrict__ meanG,
		const float* __restrict__ noiseG, float4* __restrict__ acc, float* xScratch)
{
	extern __shared__ __align__(16) unsigned char smemRaw[];
	count = chunkCount(count, countDev, first);
	constexpr int NG = NT / GS;
	constexpr int PB = 4; // similar patches per group iteration: one Minv load feeds 4 FMAs (register blocking)
	const int d = g.d, np = g.np, pw = 2 * g.w + 1, tid = threadIdx.x;
	const int gi = tid / GS, gl = tid - gi * GS;
	FilterSmem s = carveFilterSmem(smemRaw, g, xScratch ? xScratch + (size_t)blockIdx.x * g.Ns * d : nullptr, true, NG);
	float4* xcv4 = reinterpret_cast<float4*>(s.vec) + (size_t)gi * 3 * d; // [d] centred patches (4 at a time)
	float4* tv4 = xcv4 + d;                                                 // [d] -(Minv xc)
	float4* zv4 = tv4 + d;                                                  // [d] outputs (step 2)
	for(int slot = blockIdx.x; slot < count; slot += gridDim.x)
	{
		const int p = list[first + slot];
		const int py = p / g.W, px = p - py * g.W;
		const int n = expandSimilarList<NT>(smask + (size_t)p * g.Nw, g, px, py, s.words, s.prefix, s.qbase);
		gatherPatches<NT>(color, g, s.qbase, n, s.X);
		float mrow[DT > 0 ? DT : 1];
		if(DT > 0)
		{
			// Minv is symmetric: row r = column r, contiguous in the column-major matrix
			const float* src = MinvG + (size_t)slot * DT * DT + (size_t)(gl < DT ? gl : 0) * DT;
#pragma unroll
			for(int c = 0; c < DT; ++c) mrow[c] = __ldg(src + c);
		}
		else
			for(int e = tid; e < d * d; e += NT)
				s.Minv[e] = MinvG[(size_t)slot * d * d + e];
		for(int e = tid; e < d; e += NT)
			s.mean[e] = meanG[(size_t)slot * d + e];
		for(int e = tid; e < np * 6; e += NT)
			s.noise[e] = noiseG[(size_t)slot * np * 6 + e];
		__syncthreads();
		for(int i0 = 0; i0 < n; i0 += NG * PB)
		{
			const int ib = i0 + gi * PB;                    // first patch of this group
			const int nv = min(max(n - ib, 0), PB);         // valid patches of this group
			const float* X0 = s.X + (size_t)min(ib, n - 1) * d;
			const float* X1 = s.X + (size_t)min(ib + 1, n - 1) * d;
			const float* X2 = s.X + (size_t)min(ib + 2, n - 1) * d;
			const float* X3 = s.X + (size_t)min(ib + 3, n - 1) * d;
			for(int c = gl; c < d; c += GS)
			{
				const float m = s.mean[c];
				xcv4[c] = make_float4(__fsub_rn(X0[c], m), __fsub_rn(X1[c], m), __fsub_rn(X2[c], m), __fsub_rn(X3[c], m));
			}
			groupSync<GS>();
			if(DT > 0)
			{
				if(gl < DT)
				{
					float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
					for(int c = 0; c < DT; ++c)
					{
						const float m = mrow[c];
						const float4 v = xcv4[c];
						a0 = fmaf(m, v.x, a0); a1 = fmaf(m, v.y, a1); a2 = fmaf(m, v.z, a2); a3 = fmaf(m, v.w, a3);
					}
					tv4[gl] = make_float4(-a0, -a1, -a2, -a3);
				}
			}
			else
				for(int r = gl; r < d; r += GS)
				{
					float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
					for(int c = 0; c < d; ++c)
					{
						const float m = s.Minv[r + c * d];
						const float4 v = xcv4[c];
						a0 = fmaf(m, v.x, a0); a1 = fmaf(m, v.y, a1); a2 = fmaf(m, v.z, a2); a3 = fmaf(m, v.w, a3);
					}
					tv4[r] = make_float4(-a0, -a1, -a2, -a3);
				}
			groupSync<GS>();
			for(int jj = gl; jj < d; jj += GS)
			{
				const int k = jj / 3, a3i = jj - 3 * k;
				const float* cb = s.noise + 6 * k;
				const float c0 = cb[symIndex(a3i, 0)], c1 = cb[symIndex(a3i, 1)], c2 = cb[symIndex(a3i, 2)];
				const float4 t0 = tv4[3 * k], t1 = tv4[3 * k + 1], t2 = tv4[3 * k + 2];
				// multiplyCovMatPatchByVector :633-654 (left-to-right sums), then += noisy patch (:668)
				float4 o;
				o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c0, t0.x), __fmul_rn(c1, t1.x)), __fmul_rn(c2, t2.x)), X0[jj]);
				o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c0, t0.y), __fmul_rn(c1, t1.y)), __fmul_rn(c2, t2.y)), X1[jj]);
				o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c0, t0.z), __fmul_rn(c1, t1.z)), __fmul_rn(c2, t2.z)), X2[jj]);
				o.w = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c0, t0.w), __fmul_rn(c1, t1.w)), __fmul_rn(c2, t2.w)), X3[jj]);
				if(STEP == 1)
				{
					// in place: every group owns its PB rows; rows are distinct when valid
					float* Xw = s.X + (size_t)ib * d + jj;
					if(nv > 0) Xw[0] = o.x;
					if(nv > 1) Xw[d] = o.y;
					if(nv > 2) Xw[2 * d] = o.z;
					if(nv > 3) Xw[3 * d] = o.w;
				}
				else
					zv4[jj] = o;
			}
			if(STEP == 2)
			{
				groupSync<GS>();
				// compile-time patch geometry for the common widths: the two divisions become multiplications
				constexpr int NPC = DT / 3, PWC = DT == 27 ? 3 : DT == 75 ? 5 : DT == 147 ? 7 : 0;
				const int npv = PWC ? NPC : np, pwv = PWC ? PWC : pw;
				for(int e = gl; e < nv * npv; e += GS)
				{
					const int pi = e / npv, k = e - pi * npv;
					const size_t x = (size_t)(s.qbase[ib + pi] + (k / pwv - g.w) * g.W + (k % pwv - g.w));
					const float4 zr = zv4[3 * k], zg = zv4[3 * k + 1], zb = zv4[3 * k + 2];
					const float r = pi == 0 ? zr.x : pi == 1 ? zr.y : pi == 2 ? zr.z : zr.w;
					const float gg = pi == 0 ? zg.x : pi == 1 ? zg.y : pi == 2 ? zg.z : zg.w;
					const float bb = pi == 0 ? zb.x : pi == 1 ? zb.y : pi == 2 ? zb.z : zb.w;
					atomicAdd(acc + x, make_float4(r, gg, bb, 1.f));
				}
			}
			// STEP 1: rows written above are read again (as X0..X3) only by their own group, in the next slot's
			// gather or in meanCenterCov after the barrier below. The next iteration rewrites xcv4 before its first
			// barrier (its readers passed the second barrier), tv4/zv4 only after that barrier.
			// A short last group reads row n-1 through its clamped X1..X3 while the owner of that row may be rewriting it:
			// those lanes' results are discarded (nv), so the groups need no barrier between them inside the loop.
			if(STEP == 1)
				groupSync<GS>();
		}
		__syncthreads();
		if(STEP == 1)
		{
			meanCenterCov<NT, DT>(s.X, n, g, s.mean, s.noise, 1.f, A + (size_t)slot * d * d, meanG + (size_t)slot * d);
			__syncthreads();
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// K3e: symmetric eigen-function chain in shared memory, one matrix per CTA of NT >= d threads.
//   doClamp: M <- V max(0,L) V^T + Sigma_n  (clampNegativeEigenValues + addCovMatPatchToMatrix), then
//   always : out <- V diag(1/max(eps,L)) V^T (inverseSymmetricMatrix).
// The decomposition follows Eigen's SelfAdjointEigenSolver::compute step for step (scaling by max|coeff| of the
// lower triangle, Householder tridiagonalisation in place, Q accumulated from the reflectors, implicit symmetric QR
// with Wilkinson shift and the same deflation test, at most 30 n sweeps). Eigenvalue sorting is skipped because
// V f(L) V^T does not depend on the order.
// ---------------------------------------------------------------------------------------------------------------
#define MAT(r, c) M[(r) + (c) * n]

__device__ __forceinline__ float ldsF(uint32_t addr)
{
	float v;
	asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ void stsF(uint32_t addr, float v)
{
	asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory");
}

// JacobiRotation::makeGivens (Jacobi.h:216-252): c = p/r, s = -q/r with r = sqrt(p^2+q^2) in every branch of the
// reference (the branches only guard against overflow). The matrix is scaled into [-1,1], so p^2+q^2 cannot
// overflow; the common range uses rsqrt + one Newton step (<= 1 ulp) with the exact +-1 / 0 values of the
// reference's q == 0 and p == 0 cases patched in by selects; tiny magnitudes take the reference's formulas.
__device__ __forceinline__ void makeGivens(float p, float q, float& c, float& s)
{
	const float h2 = fmaf(p, p, q * q);
	if(h2 > 1e-30f)
	{
		float ri;
		asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ri) : "f"(h2));
		ri = ri * fmaf(-0.5f * h2, ri * ri, 1.5f);
		c = p * ri;
		s = -q * ri;
		if(q == 0.f) { c = p < 0.f ? -1.f : 1.f; s = 0.f; }
		if(p == 0.f) { c = 0.f; s = q < 0.f ? 1.f : -1.f; }
		return;
	}
	if(q == 0.f) { c = p < 0.f ? -1.f : 1.f; s = 0.f; return; }
	if(p == 0.f) { c = 0.f; s = q < 0.f ? 1.f : -1.f; return; }
	if(fabsf(p) > fabsf(q))
	{
		const float t = __fdiv_rn(q, p);
		float u = __fsqrt_rn(fmaf(t, t, 1.f));
		if(p < 0.f) u = -u;
		c = __fdiv_rn(1.f, u);
		s = -t * c;
	}
	else
	{
		const float t = __fdiv_rn(p, q);
		float u = __fsqrt_rn(fmaf(t, t, 1.f));
		if(q < 0.f) u = -u;
		s = __fdiv_rn(-1.f, u);
		c = -t * s;
	}
}

struct EigSmem
{
	float *M, *O, *diag, *sub, *dW, *sW, *hco, *pv, *red;
};

template <int NT>
__device__ void eigFunction(const EigSmem& S, int n, int mode, float eps)
{
	float* M = S.M;
	const int tid = threadIdx.x;
	// ---- scale the lower triangle into [-1,1], clear the strict upper triangle
	float lmax = 0.f;
	for(int e = tid; e < n * n; e += NT)
	{
		const int r = e % n, c = e / n;
		if(r < c) M[e] = 0.f;
		else lmax = fmaxf(lmax, fabsf(M[e]));
	}
	float scale = blockReduceMax<NT>(lmax, S.red);
	if(scale == 0.f) scale = 1.f;
	for(int e = tid; e < n * n; e += NT)
		if(e % n >= e / n) M[e] = __fdiv_rn(M[e], scale);
	__syncthreads();

	// ---- Householder tridiagonalisation (lower triangle), Tridiagonalization.h:346-378
	for(int i = 0; i < n - 1; ++i)
	{
		const int rs = n - i - 1;
		float* v = &MAT(i + 1, i);
		float part = 0.f;
		for(int k = 1 + tid; k < rs; k += NT)
			part += v[k] * v[k];
		const float tailSq = blockReduceSum<NT>(part, S.red);
		const float c0 = v[0];
		float beta, h;
		if(tailSq <= FLT_MIN)
		{
			h = 0.f;
			beta = c0;
			for(int k = 1 + tid; k < rs; k += NT) v[k] = 0.f;
		}
		else
		{
			beta = __fsqrt_rn(c0 * c0 + tailSq);
			if(c0 >= 0.f) beta = -beta;
			const float den = c0 - beta;
			for(int k = 1 + tid; k < rs; k += NT) v[k] = __fdiv_rn(v[k], den);
			h = __fdiv_rn(beta - c0, beta);
		}
		__syncthreads();
		// p = A_sym (h v), v[0] == 1 implicitly
		for(int r = tid; r < rs; r += NT)
		{
			float a = 0.f;
			for(int c = 0; c < rs; ++c)
			{
				const float m = r >= c ? MAT(i + 1 + r, i + 1 + c) : MAT(i + 1 + c, i + 1 + r);
				a += m * (h * (c == 0 ? 1.f : v[c]));
			}
			S.pv[r] = a;
		}
		__syncthreads();
		part = 0.f;
		for(int k = tid; k < rs; k += NT)
			part += S.pv[k] * (k == 0 ? 1.f : v[k]);
		const float dot = blockReduceSum<NT>(part, S.red);
		const float f = h * -0.5f * dot;
		__syncthreads();
		for(int k = tid; k < rs; k += NT)
			S.pv[k] += f * (k == 0 ? 1.f : v[k]);
		__syncthreads();
		for(int c = tid; c < rs; c += NT)
		{
			const float pc = S.pv[c], vc = c == 0 ? 1.f : v[c];
			for(int r = c; r < rs; ++r)
				MAT(i + 1 + r, i + 1 + c) -= (r == 0 ? 1.f : v[r]) * pc + S.pv[r] * vc;
		}
		__syncthreads();
		if(tid == 0)
		{
			v[0] = beta;
			S.hco[i] = h;
		}
		__syncthreads();
	}
	for(int k = tid; k < n; k += NT)
	{
		S.diag[k] = MAT(k, k);
		if(k < n - 1) S.sub[k] = MAT(k + 1, k);
	}
	__syncthreads();
	// ---- Q = H_0 ... H_{n-2} evaluated in place (HouseholderSequence.h:242-267, Householder.h:110-130)
	for(int k = tid; k < n; k += NT)
		MAT(k, k) = 1.f;
	__syncthreads();
	for(int k = n - 2; k >= 0; --k)
	{
		const int cs = n - k - 1;
		const float tau = S.hco[k];
		if(cs == 1)
		{
			if(tid == 0) MAT(k + 1, k + 1) *= 1.f - tau;
		}
		else if(tau != 0.f)
		{
			const float* ess = &MAT(k + 2, k);
			for(int c = tid; c < cs; c += NT)
			{
				float* col = &MAT(k + 1, k + 1 + c);
				float a = 0.f;
				for(int r = 1; r < cs; ++r)
					a += ess[r - 1] * col[r];
				const float tmp = a + col[0];
				col[0] -= tau * tmp;
				for(int r = 1; r < cs; ++r)
					col[r] -= tau * ess[r - 1] * tmp;
			}
		}
		__syncthreads();
		for(int r = k + 1 + tid; r < n; r += NT)
			MAT(r, k) = 0.f;
		__syncthreads();
	}
	// ---- implicit symmetric QR on the tridiagonal (SelfAdjointEigenSolver.h:486-530, :808-864)
	{
		int end = n - 1, start = 0, iter = 0;
		const float precision = 2.f * FLT_EPSILON;
		while(end > 0)
		{
			for(int i = tid; i < end; i += NT)
			{
				const float a = fabsf(S.sub[i]);
				if(a <= (fabsf(S.diag[i]) + fabsf(S.diag[i + 1])) * precision || a <= FLT_MIN)
					S.sub[i] = 0.f;
			}
			__syncthreads();
			while(end > 0 && S.sub[end - 1] == 0.f) end--;
			if(end <= 0) break;
			if(++iter > 30 * n) break;
			start = end - 1;
			while(start > 0 && S.sub[start - 1] != 0.f) start--;

			// one QR step on [start, end]; every thread tracks the scalars, thread 0 records them in dW/sW,
			// thread r rotates row r of Q (n <= NT)
			const float td = (S.diag[end - 1] - S.diag[end]) * 0.5f;
			const float e = S.sub[end - 1];
			float mu = S.diag[end];
			if(td == 0.f) mu -= fabsf(e);
			else
			{
				const float e2 = e * e;
				const float ax = fabsf(td), ay = fabsf(e);
				const float pmax = fmaxf(ax, ay), qp = __fdiv_rn(fminf(ax, ay), pmax);
				const float hh = pmax * __fsqrt_rn(1.f + qp * qp); // numext::hypot, MathFunctions.h:351-374
				if(e2 == 0.f) mu -= __fdiv_rn(e, td + (td > 0.f ? 1.f : -1.f)) * __fdiv_rn(e, hh);
				else mu -= __fdiv_rn(e2, td + (td > 0.f ? hh : -hh));
			}
			float x = S.diag[start] - mu;
			float z = S.sub[start];
			float dk = S.diag[start], sk = S.sub[start], skm1 = 0.f;
			const bool hasRow = tid < n;
			// raw shared-window addresses: the bulge-chase loop below is the hot spot of the whole filter, and generic
			// pointers make the compiler rebuild the shared base and reload `n` on every rotation
			const uint32_t diagA = (uint32_t)__cvta_generic_to_shared(S.diag), subA = (uint32_t)__cvta_generic_to_shared(S.sub);
			const uint32_t dWA = (uint32_t)__cvta_generic_to_shared(S.dW), sWA = (uint32_t)__cvta_generic_to_shared(S.sW);
			const uint32_t colStep = 4u * (uint32_t)n;
			uint32_t qA = (uint32_t)__cvta_generic_to_shared(M) + 4u * (uint32_t)(hasRow ? tid : 0) + colStep * (uint32_t)start; // &Q(tid, k)
			float qk = hasRow ? ldsF(qA) : 0.f;
			for(int k = start; k < end; ++k)
			{
				float c, s;
				makeGivens(x, z, c, s);
				const float dk1 = ldsF(diagA + 4u * (uint32_t)(k + 1));
				const float sdk = fmaf(s, dk, c * sk);
				const float dkp1 = fmaf(s, sk, c * dk1);
				const float ndk = fmaf(c, fmaf(c, dk, -s * sk), -s * fmaf(c, sk, -s * dk1));
				const float ndk1 = fmaf(s, sdk, c * dkp1);
				const float nsk = fmaf(c, sdk, -s * dkp1);
				if(tid == 0)
				{
					stsF(dWA + 4u * (uint32_t)k, ndk);
					if(k > start) stsF(sWA + 4u * (uint32_t)(k - 1), fmaf(c, skm1, -s * z));
				}
				x = nsk;
				float nsk1 = 0.f;
				if(k < end - 1)
				{
					const float sk1 = ldsF(subA + 4u * (uint32_t)(k + 1));
					z = -s * sk1;
					nsk1 = c * sk1;
				}
				if(hasRow)
				{
					const float y = ldsF(qA + colStep);
					stsF(qA, fmaf(c, qk, -s * y));
					qk = fmaf(s, qk, c * y);
				}
				qA += colStep;
				dk = ndk1;
				skm1 = nsk;
				sk = nsk1;
			}
			if(hasRow) stsF(qA, qk);
			if(tid == 0)
			{
				S.dW[end] = dk;
				S.sW[end - 1] = skm1;
			}
			__syncthreads();
			for(int k = start + tid; k <= end; k += NT)
			{
				S.diag[k] = S.dW[k];
				if(k < end) S.sub[k] = S.sW[k];
			}
			__syncthreads();
		}
	}
	// ---- O = V f(L) V^T  (tmp(k,c) = f_k V(c,k); out(r,c) = sum_k V(r,k) tmp(k,c), :593-603 / :621-629)
	for(int k = tid; k < n; k += NT)
	{
		const float l = S.diag[k] * scale;
		S.pv[k] = mode == 0 ? fmaxf(0.f, l) : __fdiv_rn(1.f, fmaxf(eps, l));
	}
	__syncthreads();
	const int L = n * (n + 1) / 2;
	for(int e = tid; e < L; e += NT)
	{
		int c = 0, r = e;
		while(r >= n - c) { r -= n - c; ++c; }
		r += c;
		float a = 0.f;
		for(int k = 0; k < n; ++k)
			a += MAT(r, k) * (S.pv[k] * MAT(c, k));
		S.O[r + c * n] = a;
		S.O[c + r * n] = a;
	}
	__syncthreads();
}
#undef MAT

// inverseSymmetricMatrix (DenoisingUnit.cpp:578-604) when no eigenvalue needs clamping: V diag(1/l) V^T is then simply
// A^-1. Fast path: Cholesky A = L L^T in S.O, L^-T in the (free) strict upper triangle of S.M, A^-1 = L^-T L^-1 back in
// S.O. Accepted only if every pivot is positive and 1/||A^-1||_inf > 2 eps, which proves lambda_min(A) > eps, i.e. that
// max(eps, lambda) == lambda for every eigenvalue; otherwise the caller runs the reference's eigen path on the intact
// lower triangle of S.M. Returns true (uniformly) on success.
#define MAT(r, c) M[(r) + (c) * n]
#define OUT(r, c) O[(r) + (c) * n]
template <int NT>
__device__ bool choleskyInverse(const EigSmem& S, int n, float eps)
{
	float* M = S.M;
	float* O = S.O;
	const int tid = threadIdx.x;
	for(int e = tid; e < n * n; e += NT)
		O[e] = (e % n >= e / n) ? M[e] : 0.f;
	__syncthreads();
	bool ok = true;
	for(int k = 0; k < n; ++k)
	{
		const float piv = OUT(k, k);
		if(!(piv > 0.f) || !(piv < 3e38f)) { ok = false; break; } // uniform: every thread reads the same value
		const float inv = rsqrtf(piv);
		const float lkkInv = inv * fmaf(-0.5f * piv, inv * inv, 1.5f);
		__syncthreads();
		for(int r = k + tid; r < n; r += NT)
			OUT(r, k) = r == k ? piv * lkkInv : OUT(r, k) * lkkInv;
		__syncthreads();
		for(int c = k + 1 + tid; c < n; c += NT)
		{
			const float lck = OUT(c, k);
			for(int r = c; r < n; ++r)
				OUT(r, c) = fmaf(-OUT(r, k), lck, OUT(r, c));
		}
		__syncthreads();
	}
	if(!ok)
		return false;
	// X = L^-1 by forward substitution, one column per thread; X(i,j) (i > j) stored at M(j,i), X(j,j) in pv[j]
	for(int j = tid; j < n; j += NT)
		S.pv[j] = __fdiv_rn(1.f, OUT(j, j));
	__syncthreads();
	for(int j = tid; j < n; j += NT)
	{
		const float xjj = S.pv[j];
		for(int i = j + 1; i < n; ++i)
		{
			float a = OUT(i, j) * xjj;
			for(int k = j + 1; k < i; ++k)
				a = fmaf(OUT(i, k), MAT(j, k), a);
			MAT(j, i) = -a * S.pv[i];
		}
	}
	__syncthreads();
	// A^-1(r,c) = sum_{k >= r} X(k,r) X(k,c)   (r >= c)
	const int L = n * (n + 1) / 2;
	for(int e = tid; e < L; e += NT)
	{
		int c = 0, r = e;
		while(r >= n - c) { r -= n - c; ++c; }
		r += c;
		float a = S.pv[r] * (r == c ? S.pv[r] : MAT(c, r));
		for(int k = r + 1; k < n; ++k)
			a = fmaf(MAT(r, k), MAT(c, k), a);
		OUT(r, c) = a;
	}
	__syncthreads();
	for(int e = tid; e < n * n; e += NT)
	{
		const int r = e % n, c = e / n;
		if(r < c) O[e] = OUT(c, r);
	}
	__syncthreads();
	float rowMax = 0.f;
	for(int r = tid; r < n; r += NT)
	{
		float a = 0.f;
		for(int c = 0; c < n; ++c)
			a += fabsf(OUT(r, c));
		rowMax = fmaxf(rowMax, a);
	}
	const float norm = blockReduceMax<NT>(rowMax, S.red);
	__syncthreads();
	return norm < 0.5f / eps && norm == norm;
}
#undef MAT
#undef OUT

template <int NT>
__global__ void __launch_bounds__(NT)
k_eig_chain(const float* __restrict__ A, float* __restrict__ MinvG, const float* __restrict__ noiseG, int count,
		const int* __restrict__ countDev, int first, int n, int np, float eps, int doClamp)
{
	extern __shared__ __align__(16) unsigned char smemRaw[];
	count = chunkCount(count, countDev, first);
	float* f = reinterpret_cast<float*>(smemRaw);
	EigSmem S;
	S.M = f; f += n * n;
	S.O = f; f += n * n;
	S.diag = f; f += n;
	S.sub = f; f += n;
	S.dW = f; f += n;
	S.sW = f; f += n;
	S.hco = f; f += n;
	S.pv = f; f += n;
	S.red = f; f += 32;
	float* noise = f;
	const int tid = threadIdx.x;
	for(int slot = blockIdx.x; slot < count; slot += gridDim.x)
	{
		__syncthreads();
		for(int e = tid; e < n * n; e += NT)
			S.M[e] = A[(size_t)slot * n * n + e];
		for(int e = tid; e < np * 6; e += NT)
			noise[e] = noiseG[(size_t)slot * np * 6 + e];
		__syncthreads();
		if(doClamp)
		{
			eigFunction<NT>(S, n, 0, 0.f);
			// M <- O + Sigma_n on the 3x3 diagonal blocks (:538-556)
			for(int e = tid; e < n * n; e += NT)
			{
				const int r = e % n, c = e / n;
				float v = S.O[e];
				if(r / 3 == c / 3)
					v = __fadd_rn(v, noise[(r / 3) * 6 + symIndex(r % 3, c % 3)]);
				S.M[e] = v;
			}
			__syncthreads();
		}
		if(!choleskyInverse<NT>(S, n, eps))
			eigFunction<NT>(S, n, 1, eps);
		for(int e = tid; e < n * n; e += NT)
			MinvG[(size_t)slot * n * n + e] = S.O[e];
	}
}

// ---------------------------------------------------------------------------------------------------------------
// launcher of the chain for one chunk of the full list
// ---------------------------------------------------------------------------------------------------------------
// `count` = capacity of the chunk (slots [first, first + count) of the list); countDev != nullptr: the list length lives
// on the device (the kernels clip the chunk to it and idle CTAs leave at once), so no host read-back sits between the
// marking stage and the filter. Grids are persistent (slot loops), capped at a few CTAs per SM.
// ev: six events recorded around the five launches (resolved by the caller after its final synchronisation).
template <int NT, int GS, int DT>
static cudaError_t runFilterChunk(const float* color, const float* pixcov, const uint32_t* smask, const LevelGeom& g,
		float minEig, const int* fullList, int first, int count, const int* countDev, const FilterBuffers& fb, float4* acc,
		cudaStream_t st, int* launches, cudaEvent_t* ev, int eigNT)
{
	const size_t smemLimit = 200 * 1024;
	bool xInSmem = filterSmemBytes(g, true, true, NT / GS) <= smemLimit;
	if(!xInSmem && !fb.xScratch)
		return cudaErrorInvalidConfiguration;
	float* xs = xInSmem ? nullptr : fb.xScratch;
	const int sms = fb.smCount > 0 ? fb.smCount : currentSmCount();
	const int grid = std::min(count, xInSmem ? sms * 16 : fb.gridCap);
	const size_t smStats = filterSmemBytes(g, xInSmem, false, 0);
	const size_t smApply = filterSmemBytes(g, xInSmem, true, NT / GS);
	const size_t smEig = ((size_t)2 * g.d * g.d + 6 * g.d + 32 + g.np * 6) * 4;
	cudaError_t err;
	auto kStats = k_filter_stats<NT, DT>;
	auto kA1 = k_filter_apply<NT, GS, 1, DT>;
	auto kA2 = k_filter_apply<NT, GS, 2, DT>;
	if((err = allowDynamicSmem((const void*)kStats, smStats)) != cudaSuccess) return err;
	if((err = allowDynamicSmem((const void*)kA1, smApply)) != cudaSuccess) return err;
	if((err = allowDynamicSmem((const void*)kA2, smApply)) != cudaSuccess) return err;

	auto launchEig = [&](int doClamp) -> cudaError_t {
		if(eigNT == 32 && g.d <= 32 && (g_eigGroup == 16 || g_eigGroup == 8))
			return launchEigChainSub(fb.A, fb.Minv, fb.noise, count, countDev, first, g.d, g.np, minEig, doClamp, g_eigGroup, sms, st);
		const int eg = std::min(count, sms * 8);
		if(eigNT == 32)
		{
			allowDynamicSmem((const void*)k_eig_chain<32>, smEig);
			k_eig_chain<32><<<eg, 32, smEig, st>>>(fb.A, fb.Minv, fb.noise, count, countDev, first, g.d, g.np, minEig, doClamp);
		}
		else if(eigNT == 96)
		{
			allowDynamicSmem((const void*)k_eig_chain<96>, smEig);
			k_eig_chain<96><<<eg, 96, smEig, st>>>(fb.A, fb.Minv, fb.noise, count, countDev, first, g.d, g.np, minEig, doClamp);
		}
		else
		{
			allowDynamicSmem((const void*)k_eig_chain<160>, smEig);
			k_eig_chain<160><<<eg, 160, smEig, st>>>(fb.A, fb.Minv, fb.noise, count, countDev, first, g.d, g.np, minEig, doClamp);
		}
		return cudaGetLastError();
	};

	if(ev) cudaEventRecord(ev[0], st);
	kStats<<<grid, NT, smStats, st>>>(color, pixcov, smask, g, fullList, first, count, countDev, fb.A, fb.mean, fb.noise, xs);
	if((err = cudaGetLastError()) != cudaSuccess) return err;
	if(ev) cudaEventRecord(ev[1], st);
	if((err = launchEig(1)) != cudaSuccess) return err;
	if(ev) cudaEventRecord(ev[2], st);
	kA1<<<grid, NT, smApply, st>>>(color, smask, g, fullList, first, count, countDev, fb.A, fb.Minv, fb.mean, fb.noise, acc, xs);
	if((err = cudaGetLastError()) != cudaSuccess) return err;
	if(ev) cudaEventRecord(ev[3], st);
	if((err = launchEig(0)) != cudaSuccess) return err;
	if(ev) cudaEventRecord(ev[4], st);
	kA2<<<grid, NT, smApply, st>>>(color, smask, g, fullList, first, count, countDev, fb.A, fb.Minv, fb.mean, fb.noise, acc, xs);
	if((err = cudaGetLastError()) != cudaSuccess) return err;
	if(ev) cudaEventRecord(ev[5], st);
	*launches += 5;
	return cudaSuccess;
}

cudaError_t launchFilterChunk(const float* color, const float* pixcov, const uint32_t* smask, const int* counts,
		const LevelGeom& g, float minEig, const int* fullList, int first, int count, const int* countDev, const FilterBuffers& fb,
		float4* acc, cudaStream_t st, int* launches, cudaEvent_t* ev)
{
	(void)counts;
	if(count <= 0)
		return cudaSuccess;
	if(g.d == 27) // the default patch radius: compile-time d
		return runFilterChunk<128, 32, 27>(color, pixcov, smask, g, minEig, fullList, first, count, countDev, fb, acc, st, launches, ev, 32);
	if(g.d <= 32)
		return runFilterChunk<128, 32, 0>(color, pixcov, smask, g, minEig, fullList, first, count, countDev, fb, acc, st, launches, ev, 32);
	if(g.d <= 96)
		return runFilterChunk<192, 96, 0>(color, pixcov, smask, g, minEig, fullList, first, count, countDev, fb, acc, st, launches, ev, 96);
	if(g.d <= 160)
		return runFilterChunk<160, 160, 0>(color, pixcov, smask, g, minEig, fullList, first, count, countDev, fb, acc, st, launches, ev, 160);
	return cudaErrorInvalidValue;
}

} // namespace bcdk

// K2 — the "marking strategy" (m = markedPixelsSkippingProbability > 0) resolved in the reference's sequential
// raster order, in parallel.
//
// Replaces (reference, paths relative to /root/reference): the skip test DenoisingUnit.cpp:164-173, the marking
// DenoisingUnit.cpp:690, the fallback threshold :182 and the raster pixel order Denoiser.cpp:136-147 (the only
// deterministic order: useRandomPixelOrder = false, nbOfCores = 1).
//
// Closed form (validated against the literal sequential process by tests/test_oracle.py):
//     active(p)  <=>  not exists q < p (raster), q in S(p), AS(q)        with AS(q) = active(q) and |S(q)| >= d+1
// because only fully filtered pixels mark their similar set and S is symmetric. p depends on the b rows above
// (columns c-b..c+b) and on the b pixels to its left.
//
// Parallelisation: a warp owns a band of 32 consecutive rows, lane l = row r0+l, and walks the columns with a skew
// of SK = b+2 columns per lane (at step t lane l is at column t - SK*l), so every lane finds the AS bits it needs
// produced by the lanes above at least two steps earlier. AS bits live as bit-rows in shared memory (one funnel
// shift extracts the 2b+1 bits of a row above). The per-pixel input (`pre`: the similar patches before p in raster
// order + the strong flag, packed by k_symmetrize) is streamed with cp.async in chunks of 32 steps into a
// double-buffered shared-memory tile (coalesced 128-byte rows instead of 32 scattered per-lane loads).
// Bands are chained through global memory: once per 32 steps a band flushes the completed words of its last b rows
// to `asBits`, fences once, and publishes how many columns its last row has completed; the next band (another CTA,
// all CTAs co-resident: one warp each) polls that counter once per 32 steps.

#include "common.cuh"
#include "kernels.h"
#include <cuda_pipeline.h>

namespace bcdk
{

__device__ __forceinline__ uint32_t hash32(uint32_t a, uint32_t b, uint32_t seed)
{
	uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ seed;
	h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
	return h;
}

// `len` (<= 32) bits starting at bit `off` of the little-endian word array w (nwords words)
__device__ __forceinline__ uint32_t extractBits(const uint32_t* w, int off, int len, int nwords)
{
	const int w0 = off >> 5, sh = off & 31;
	const uint32_t lo = w[w0];
	const uint32_t hi = (w0 + 1 < nwords) ? w[w0 + 1] : 0u;
	const uint32_t v = __funnelshift_r(lo, hi, sh);
	return len >= 32 ? v : (v & ((1u << len) - 1u));
}

// BT = compile-time search radius (6, 10, 14) or 0 for a run-time radius.
template <int BT>
__global__ void __launch_bounds__(32)
k_marking(const uint32_t* __restrict__ pre, LevelGeom g, float skipProb, uint32_t seed,
		uint32_t* __restrict__ actBits, uint32_t* asBits, int Wp, unsigned int* progress)
{
	extern __shared__ uint32_t smem[];
	const int b = BT ? BT : g.b;
	const int side = 2 * b + 1, SK = b + 2;
	const int W = g.W, H = g.H, w = g.w;
	const int Pw = BT ? (BT * (2 * BT + 1) + BT + 1 + 31) / 32 : g.Pw;
	const int preRow = 32 * Pw + 1;               // padded: lane stride must be odd
	uint32_t* asS = smem;                         // [(32 + b)][Wp]
	uint32_t* preS = smem + (size_t)(32 + b) * Wp; // [2][32][preRow]
	const int lane = threadIdx.x;
	const int band = blockIdx.x;
	const int r0 = g.rowBeg + band * 32;
	const int r = r0 + lane;
	const bool rowValid = r < g.rowEnd;
	const uint32_t thr = skipProb >= 1.f ? 0xffffffffu : (uint32_t)(skipProb * 4294967296.f);
	const bool always = skipProb >= 1.f;
	const int centreBit = b * side + b;
	const int lastLane = min(31, g.rowEnd - 1 - r0); // lane of the last valid row of the band

	for(int e = lane; e < (32 + b) * Wp; e += 32)
		asS[e] = 0u;

	// chunk q of row l covers columns [32q - SK*l, 32q - SK*l + 32)
	auto issueRow = [&](int q, int l)
	{
		const int rr = r0 + l;
		const int c0 = 32 * q - SK * l;
		uint32_t* dst = preS + (size_t)(q & 1) * 32 * preRow + (size_t)l * preRow;
		for(int e = lane; e < 32 * Pw; e += 32)
		{
			const int c = c0 + e / Pw;
			const bool ok = rr < g.rowEnd && c >= 0 && c < W;
			const uint32_t* src = ok ? pre + ((size_t)rr * W + c) * Pw + (e % Pw) : pre;
			__pipeline_memcpy_async(dst + e, src, 4, ok ? 0 : 4);
		}
	};
	for(int l = 0; l < 32; ++l)
		issueRow(0, l);
	__pipeline_commit();
	__pipeline_wait_prior(0);
	__syncwarp();

	uint32_t* myRow = asS + (size_t)(lane + b) * Wp;
	uint32_t curWord = 0u, curAct = 0u, ownHist = 0u;
	int copiedWords = 0;  // words [0, copiedWords) of the predecessor rows are in shared memory
	int flushedWords = 0; // words [0, flushedWords) of my row are in asBits
	const int tEnd = W + SK * 31;
	const volatile unsigned int* prevProgress = progress + (band > 0 ? band - 1 : 0);
	const size_t rowBase = (size_t)(rowValid ? r : 0);

	for(int t = 0; t < tEnd; ++t)
	{
		const int s = t & 31, q = t >> 5;
		if(s == 0)
		{
			if(q > 0)
			{
				__pipeline_wait_prior(0); // chunk q of `pre` (issued during chunk q-1) has landed
				__syncwarp();
			}
			// rows of the previous band: every bit up to column t + 31 + b, i.e. words <= ((t+31+b+MARK_PAD)>>5)
			const int needWords = min(Wp, ((t + 31 + b + MARK_PAD) >> 5) + 1);
			const unsigned int needCols = (unsigned int)min(W, (needWords << 5) - MARK_PAD);
			if(band > 0 && lane == 0) // band 0 reads rows the host prepared (zeros, or the upper row tile's hand-off)
				while(*prevProgress < needCols) { __nanosleep(100); }
			__syncwarp();
			if(needWords > copiedWords)
			{
				__threadfence();
				const int nw = needWords - copiedWords;
				for(int e = lane; e < b * nw; e += 32)
				{
					const int k = e / nw, wi = copiedWords + e % nw;
					const int rr = r0 - b + k; // absolute row of the predecessor
					uint32_t v = 0u;
					if(rr >= 0)
						v = *(volatile uint32_t*)(asBits + (size_t)rr * Wp + wi);
					asS[(size_t)k * Wp + wi] = v;
				}
				copiedWords = needWords;
				__syncwarp();
			}
		}
		// prefetch row s of the next chunk (one row per step keeps the copies coalesced and off the critical path)
		if(32 * (q + 1) < tEnd)
			issueRow(q + 1, s);
		if(s == 31)
			__pipeline_commit();

		const int c = t - SK * lane;
		const bool valid = rowValid && c >= 0 && c < W;
		const uint32_t* pw = preS + (size_t)(q & 1) * 32 * preRow + (size_t)lane * preRow + s * Pw;
		uint32_t hit = 0u, strongBit = 0u;
		if(valid)
		{
			const int posLo = c - b + MARK_PAD;
			const int wi = posLo >> 5, sh = posLo & 31;
#pragma unroll
			for(int k = 1; k <= (BT ? BT : 15); ++k)
			{
				if(!BT && k > b) break;
				const uint32_t* rowAbove = asS + (size_t)(lane + b - k) * Wp;
				const uint32_t field = __funnelshift_r(rowAbove[wi], rowAbove[wi + 1], sh);
				hit |= field & extractBits(pw, (b - k) * side, side, Pw);
			}
			hit |= extractBits(pw, b * side, b, Pw) & ownHist;
			strongBit = (pw[centreBit >> 5] >> (centreBit & 31)) & 1u;
		}
		const bool centre = valid && r >= w && r <= H - 1 - w && c >= w && c <= W - 1 - w;
		const bool skipIfMarked = always || hash32((uint32_t)r, (uint32_t)c, seed) < thr;
		const bool act = centre && !(hit != 0u && skipIfMarked);
		const uint32_t as = (act && strongBit) ? 1u : 0u;
		if(valid)
		{
			ownHist = (ownHist >> 1) | (as << (b - 1));
			const int pos = c + MARK_PAD;
			curWord |= as << (pos & 31);
			curAct |= (act ? 1u : 0u) << (pos & 31);
			myRow[pos >> 5] = curWord;
			if((pos & 31) == 31 || c == W - 1)
			{
				actBits[rowBase * Wp + (pos >> 5)] = curAct;
				curWord = 0u;
				curAct = 0u;
			}
		}
		__syncwarp();
		if(s == 31 || t == tEnd - 1)
		{
			// flush the completed words of the hand-off rows, one fence, one progress store per 32 steps
			const int done = rowValid ? min(max(c + 1, 0), W) : 0;                     // columns completed by my row
			const int complete = done >= W ? Wp : ((done + MARK_PAD) >> 5);           // fully written words
			if(rowValid && (lane >= 32 - b || r >= g.rowEnd - b))
				for(int wi = flushedWords; wi < complete; ++wi)
					*(volatile uint32_t*)(asBits + (size_t)r * Wp + wi) = myRow[wi];
			flushedWords = max(flushedWords, complete);
			__syncwarp();
			if(lane == lastLane)
			{
				__threadfence();
				const unsigned int cols = done >= W ? (unsigned int)W : (unsigned int)max(0, (complete << 5) - MARK_PAD);
				*(volatile unsigned int*)(progress + band) = cols;
			}
		}
	}
}

// m = 0: every interior pixel is processed (the mask is never consulted, DenoisingUnit.cpp:165-166)
__global__ void k_all_active_bits(LevelGeom g, uint32_t* __restrict__ actBits, int Wp)
{
	const int wi = blockIdx.x * blockDim.x + threadIdx.x;
	const int y = blockIdx.y;
	if(wi >= Wp)
		return;
	uint32_t v = 0u;
	if(y >= g.w && y <= g.H - 1 - g.w)
		for(int bit = 0; bit < 32; ++bit)
		{
			const int x = wi * 32 + bit - MARK_PAD;
			if(x >= g.w && x <= g.W - 1 - g.w)
				v |= 1u << bit;
		}
	actBits[(size_t)y * Wp + wi] = v;
}

cudaError_t launchAllActive(const LevelGeom& g, uint32_t* actBits, int Wp, cudaStream_t st, int* launches)
{
	dim3 block(64), grid(divUp(Wp, 64), g.H);
	k_all_active_bits<<<grid, block, 0, st>>>(g, actBits, Wp);
	*launches += 1;
	return cudaGetLastError();
}

cudaError_t launchMarking(const uint32_t* pre, const LevelGeom& g, float skipProb, uint32_t seed,
		uint32_t* actBits, uint32_t* asBits, int Wp, unsigned int* progress, cudaStream_t st, int* launches)
{
	const int rows = g.rowEnd - g.rowBeg;
	if(rows <= 0)
		return cudaSuccess;
	const int bands = divUp(rows, 32);
	const size_t smem = ((size_t)(32 + g.b) * Wp + (size_t)2 * 32 * (32 * g.Pw + 1)) * 4;
	cudaError_t err;
	if((err = cudaMemsetAsync(progress, 0, sizeof(unsigned int) * bands, st)) != cudaSuccess) return err;
	*launches += 1;
#define LAUNCH_MARK(BT) \
	do { \
		if((err = cudaFuncSetAttribute(k_marking<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err; \
		k_marking<BT><<<bands, 32, smem, st>>>(pre, g, skipProb, seed, actBits, asBits, Wp, progress); \
	} while(0)
	if(g.b == 6) LAUNCH_MARK(6);
	else if(g.b == 10) LAUNCH_MARK(10);
	else if(g.b == 14) LAUNCH_MARK(14);
	else LAUNCH_MARK(0);
#undef LAUNCH_MARK
	return cudaGetLastError();
}

} // namespace bcdk

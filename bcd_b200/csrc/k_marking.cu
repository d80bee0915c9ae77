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

// Generic kernel (any radius b <= 15): one pixel per lane-step.
template <int BT>
__global__ void __launch_bounds__(32)
k_marking_generic(const uint32_t* __restrict__ pre, int prePitch, LevelGeom g, float skipProb, uint32_t seed,
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
			const uint32_t* src = ok ? pre + ((size_t)rr * prePitch + c) * Pw + (e % Pw) : pre;
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

// ---------------------------------------------------------------------------------------------------------------
// Fast kernel (b = 6, 10, 14): PX = 4 pixels per lane-step. Lane l handles columns C..C+3, C = 4 (t - 3 l): the
// lane above is 12 columns ahead, which covers the b-column look-ahead of the last pixel of the block (b + 4 <= 12
// needs b <= 8; for b = 10, 14 the skew grows to 4 and 5 steps). One funnel shift per row above yields the 2b+4
// bits all four pixels need; the 4 x Pw `pre` words of the block arrive as Pw LDS.128 from a cp.async-fed tile
// (16-byte copies: `pre` rows are pitched to a multiple of 4 pixels). The sequential part is only the same-row
// chain (4 x ~5 dependent ALU ops); everything else is independent work for the single resident warp.
// Total dependent steps ~ (W + 4 SKS H) / 4 instead of W + (b+2) H.
// ---------------------------------------------------------------------------------------------------------------
template <int BT>
__global__ void __launch_bounds__(32)
k_marking_fast(const uint32_t* __restrict__ pre, int prePitch, LevelGeom g, float skipProb, uint32_t seed,
		uint32_t* __restrict__ actBits, uint32_t* asBits, int Wp, unsigned int* progress)
{
	extern __shared__ __align__(16) uint32_t smem[];
	constexpr int b = BT, side = 2 * BT + 1;
	constexpr int PX = 4;
	constexpr int SKS = (BT + PX + PX - 1) / PX + 0; // steps of skew: 4*SKS >= b + 4
	constexpr int SKC = PX * SKS;                    // columns of skew
	constexpr int PWT = (BT * side + BT + 1 + 31) / 32;
	constexpr int CHS = 8;                           // steps per `pre` chunk (32 pixels per row)
	constexpr int RS = CHS * PX * PWT + 4;           // tile row stride in words (== 4 mod 32: LDS.128 conflict-free)
	constexpr int centreBit = BT * side + BT;
	static_assert(2 * BT + PX <= 32, "window must fit one funnel shift");
	const int W = g.W, H = g.H, w = g.w;
	uint32_t* asS = smem;                                                // [(32 + b)][Wp]
	uint32_t* preS = smem + (((size_t)(32 + b) * Wp + 3) & ~(size_t)3);  // [2][32][RS], 16-byte aligned
	const int lane = threadIdx.x;
	const int band = blockIdx.x;
	const int r0 = g.rowBeg + band * 32;
	const int r = r0 + lane;
	const bool rowValid = r < g.rowEnd;
	const uint32_t thr = skipProb >= 1.f ? 0xffffffffu : (uint32_t)(skipProb * 4294967296.f);
	const bool always = skipProb >= 1.f;
	const int lastLane = min(31, g.rowEnd - 1 - r0);
	const int nBlocks = (W + PX - 1) / PX;
	const int tEnd = nBlocks + SKS * 31;
	const int nChunks = (tEnd + CHS - 1) / CHS;

	for(int e = lane; e < (32 + b) * Wp; e += 32)
		asS[e] = 0u;

	// chunk q of row l covers pixels [32 q - SKC l, +32): 8*PWT 16-byte units per row
	auto issueRows = [&](int q, int lFirst, int nRows)
	{
		constexpr int UNITS = CHS * PX * PWT / 4; // per row
		for(int u = lane; u < nRows * UNITS; u += 32)
		{
			const int l = lFirst + u / UNITS, v = u - (u / UNITS) * UNITS;
			const int rr = r0 + l;
			const int c = CHS * PX * q - SKC * l + (4 * v) / PWT;   // pixel of the first word of this unit
			const int wordInRow = 4 * v;                            // word offset inside the 32-pixel segment
			const bool ok = rr < g.rowEnd && c >= 0 && c < prePitch;
			// a unit never straddles the segment: offsets below are in words relative to the segment start
			const int segStart = CHS * PX * q - SKC * l;
			const uint32_t* src = ok ? pre + ((size_t)rr * prePitch + segStart) * PWT + wordInRow : pre;
			uint32_t* dst = preS + (size_t)(q & 1) * 32 * RS + (size_t)l * RS + wordInRow;
			// the segment may start left of the image or end right of the pitch: whole units only when the first
			// and last word are inside
			const int cLast = segStart + (wordInRow + 3) / PWT;
			const bool inside = ok && segStart + wordInRow / PWT >= 0 && cLast < prePitch;
			__pipeline_memcpy_async(dst, inside ? src : pre, 16, inside ? 0 : 16);
		}
	};
	issueRows(0, 0, 32);
	__pipeline_commit();
	__pipeline_wait_prior(0);
	__syncwarp();

	uint32_t* myRow = asS + (size_t)(lane + b) * Wp;
	uint32_t curWord = 0u, curAct = 0u, ownHist = 0u;
	int copiedWords = 0, flushedWords = 0;
	const volatile unsigned int* prevProgress = progress + (band > 0 ? band - 1 : 0);
	const size_t rowBase = (size_t)(rowValid ? r : 0);
	constexpr uint32_t sideMask = (1u << side) - 1u;

	for(int t = 0; t < tEnd; ++t)
	{
		const int s = t % CHS, q = t / CHS;
		if(s == 0)
		{
			if(q > 0)
			{
				__pipeline_wait_prior(0);
				__syncwarp();
			}
			// predecessor rows: lane 0 will reach column 4 (t + CHS) - 1 + b within this chunk
			const int lastCol = PX * (t + CHS) - 1 + b;
			const int needWords = min(Wp, ((lastCol + MARK_PAD) >> 5) + 1);
			const unsigned int needCols = (unsigned int)min(W, (needWords << 5) - MARK_PAD);
			if(band > 0 && lane == 0)
				while(*prevProgress < needCols) { __nanosleep(64); }
			__syncwarp();
			if(needWords > copiedWords)
			{
				__threadfence();
				const int nw = needWords - copiedWords;
				for(int e = lane; e < b * nw; e += 32)
				{
					const int k = e / nw, wi = copiedWords + e % nw;
					const int rr = r0 - b + k;
					uint32_t v = 0u;
					if(rr >= 0)
						v = *(volatile uint32_t*)(asBits + (size_t)rr * Wp + wi);
					asS[(size_t)k * Wp + wi] = v;
				}
				copiedWords = needWords;
				__syncwarp();
			}
		}
		if(q + 1 < nChunks)
			issueRows(q + 1, s * (32 / CHS), 32 / CHS);
		if(s == CHS - 1)
			__pipeline_commit();

		const int C = PX * (t - SKS * lane);
		const bool valid = rowValid && C >= 0 && C < W;
		// 4 pixels x PWT words of `pre`
		uint32_t pw[PX * PWT];
		{
			const uint4* src = reinterpret_cast<const uint4*>(preS + (size_t)(q & 1) * 32 * RS + (size_t)lane * RS + s * PX * PWT);
#pragma unroll
			for(int i = 0; i < PWT; ++i)
			{
				const uint4 v = src[i];
				pw[4 * i + 0] = v.x; pw[4 * i + 1] = v.y; pw[4 * i + 2] = v.z; pw[4 * i + 3] = v.w;
			}
		}
		uint32_t hitAbove[PX];
#pragma unroll
		for(int i = 0; i < PX; ++i) hitAbove[i] = 0u;
		if(valid)
		{
			const int posLo = C - b + MARK_PAD;
			const int wi = posLo >> 5, sh = posLo & 31;
#pragma unroll
			for(int k = 1; k <= BT; ++k)
			{
				const uint32_t* rowAbove = asS + (size_t)(lane + b - k) * Wp;
				const uint32_t win = __funnelshift_r(rowAbove[wi], rowAbove[wi + 1], sh); // columns C-b .. C-b+31
#pragma unroll
				for(int i = 0; i < PX; ++i)
				{
					constexpr int dummy = 0; (void)dummy;
					const int off = (BT - k) * side;
					const int w0 = off >> 5, so = off & 31;
					const uint32_t lo = pw[i * PWT + w0];
					const uint32_t hi = (w0 + 1 < PWT) ? pw[i * PWT + w0 + 1] : 0u;
					const uint32_t sm = __funnelshift_r(lo, hi, so) & sideMask;
					hitAbove[i] |= (win >> i) & sm;
				}
			}
		}
		uint32_t asNibble = 0u, actNibble = 0u;
#pragma unroll
		for(int i = 0; i < PX; ++i)
		{
			const int c = C + i;
			const int off = BT * side;
			const int w0 = off >> 5, so = off & 31;
			const uint32_t lo = pw[i * PWT + w0];
			const uint32_t hi = (w0 + 1 < PWT) ? pw[i * PWT + w0 + 1] : 0u;
			const uint32_t smSame = __funnelshift_r(lo, hi, so) & ((1u << BT) - 1u);
			const uint32_t strongBit = (pw[i * PWT + (centreBit >> 5)] >> (centreBit & 31)) & 1u;
			const bool centre = valid && r >= w && r <= H - 1 - w && c >= w && c <= W - 1 - w;
			const bool skipIfMarked = always || hash32((uint32_t)r, (uint32_t)c, seed) < thr;
			const uint32_t hit = hitAbove[i] | (smSame & ownHist);
			const bool act = centre && !(hit != 0u && skipIfMarked);
			const uint32_t as = (act && strongBit) ? 1u : 0u;
			ownHist = (ownHist >> 1) | (as << (BT - 1));
			asNibble |= as << i;
			actNibble |= (act ? 1u : 0u) << i;
		}
		if(valid)
		{
			const int pos = C + MARK_PAD; // multiple of 4: the block never straddles a word
			curWord |= asNibble << (pos & 31);
			curAct |= actNibble << (pos & 31);
			myRow[pos >> 5] = curWord;
			if((pos & 31) == 28 || C + PX >= W)
			{
				actBits[rowBase * Wp + (pos >> 5)] = curAct;
				curWord = 0u;
				curAct = 0u;
			}
		}
		__syncwarp();
		if(s == CHS - 1 || t == tEnd - 1)
		{
			const int done = rowValid ? min(max(C + PX, 0), W) : 0;
			const int complete = done >= W ? Wp : ((done + MARK_PAD) >> 5);
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

bool g_forceGenericMarking = false; // test hook (BCD_B200_GENERIC_MARKING=1)

cudaError_t launchMarking(const uint32_t* pre, int prePitch, const LevelGeom& g, float skipProb, uint32_t seed,
		uint32_t* actBits, uint32_t* asBits, int Wp, unsigned int* progress, cudaStream_t st, int* launches)
{
	const int rows = g.rowEnd - g.rowBeg;
	if(rows <= 0)
		return cudaSuccess;
	const int bands = divUp(rows, 32);
	cudaError_t err;
	if((err = cudaMemsetAsync(progress, 0, sizeof(unsigned int) * bands, st)) != cudaSuccess) return err;
	*launches += 1;
	const size_t asWords = ((size_t)(32 + g.b) * Wp + 3) & ~(size_t)3;
#define LAUNCH_FAST(BT) \
	do { \
		const size_t smem = (asWords + (size_t)2 * 32 * (8 * 4 * g.Pw + 4)) * 4; \
		if((err = cudaFuncSetAttribute(k_marking_fast<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err; \
		k_marking_fast<BT><<<bands, 32, smem, st>>>(pre, prePitch, g, skipProb, seed, actBits, asBits, Wp, progress); \
	} while(0)
	if(g.b == 6 && !g_forceGenericMarking) LAUNCH_FAST(6);
	else if(g.b == 10 && !g_forceGenericMarking) LAUNCH_FAST(10);
	else if(g.b == 14 && !g_forceGenericMarking) LAUNCH_FAST(14);
	else
	{
		const size_t smem = ((size_t)(32 + g.b) * Wp + (size_t)2 * 32 * (32 * g.Pw + 1)) * 4;
		if((err = cudaFuncSetAttribute(k_marking_generic<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
		k_marking_generic<0><<<bands, 32, smem, st>>>(pre, prePitch, g, skipProb, seed, actBits, asBits, Wp, progress);
	}
#undef LAUNCH_FAST
	return cudaGetLastError();
}

} // namespace bcdk

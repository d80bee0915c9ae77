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
// of SK = b+1 columns per lane (at step t lane l is at column t - SK*l), so every lane finds the AS bits it needs
// already produced by the lanes above it in the previous steps. AS bits live as bit-rows in shared memory (one
// funnel shift extracts the 2b+1 bits of a row above); the pixel's "before" half of S(p) plus its strong flag
// arrive as the packed `pre` words written by k_symmetrize. Bands are chained through global memory: the last b
// rows of a band are flushed word by word to `asBits` and the band publishes how many columns its last row has
// completed; the next band (another CTA, all CTAs co-resident) polls that counter once per 32 columns.

#include "common.cuh"
#include "kernels.h"

namespace bcdk
{

__device__ __forceinline__ uint32_t hash32(uint32_t a, uint32_t b, uint32_t seed)
{
	uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ seed;
	h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
	return h;
}

// `len` (<= 32) bits starting at bit `off` of the little-endian word array w (static indices when off is constant)
template <typename WordsT>
__device__ __forceinline__ uint32_t extractBits(const WordsT& w, int off, int len, int nwords)
{
	const int w0 = off >> 5, sh = off & 31;
	const uint32_t lo = w[w0];
	const uint32_t hi = (w0 + 1 < nwords) ? w[w0 + 1] : 0u;
	const uint32_t v = __funnelshift_r(lo, hi, sh);
	return len >= 32 ? v : (v & ((1u << len) - 1u));
}

constexpr int MARK_PAD = 32; // bit position of column 0 inside a shared/global AS bit-row

// BT = compile-time search radius (6, 10, 14) or 0 for the generic path that re-reads `pre` from L1.
template <int BT>
__global__ void __launch_bounds__(32)
k_marking(const uint32_t* __restrict__ pre, LevelGeom g, float skipProb, uint32_t seed,
		unsigned char* __restrict__ active, uint32_t* asBits, int Wp, unsigned int* progress)
{
	extern __shared__ uint32_t asS[]; // [(32 + b)][Wp]
	const int b = BT ? BT : g.b;
	const int side = 2 * b + 1, SK = b + 1;
	const int W = g.W, H = g.H, w = g.w;
	const int Pw = g.Pw;
	constexpr int PWT = BT ? (BT * (2 * BT + 1) + BT + 1 + 31) / 32 : 1;
	const int lane = threadIdx.x;
	const int band = blockIdx.x;
	const int r0 = g.rowBeg + band * 32;
	const int r = r0 + lane;
	const bool rowValid = r < g.rowEnd;
	const uint32_t thr = skipProb >= 1.f ? 0xffffffffu : (uint32_t)(skipProb * 4294967296.f);
	const bool always = skipProb >= 1.f;

	for(int e = lane; e < (32 + b) * Wp; e += 32)
		asS[e] = 0u;
	__syncwarp();

	uint32_t* myRow = asS + (size_t)(lane + b) * Wp;
	uint32_t curWord = 0u, ownHist = 0u;
	int copiedWords = 0; // words [0, copiedWords) of the predecessor rows are in shared memory
	const int tEnd = W + SK * 31;
	const volatile unsigned int* prevProgress = progress + (band > 0 ? band - 1 : 0); // written by the band above
	const size_t rowBase = (size_t)(rowValid ? r : 0) * W;

	uint32_t pw0[PWT], pw1[PWT]; // software-pipelined `pre` words (distance 1)
#pragma unroll
	for(int i = 0; i < PWT; ++i) { pw0[i] = 0u; pw1[i] = 0u; }
	if(BT)
	{
		const int c1 = 0 - SK * lane; // column at step 0
		if(rowValid && c1 >= 0 && c1 < W)
		{
#pragma unroll
			for(int i = 0; i < PWT; ++i) pw1[i] = pre[(rowBase + c1) * Pw + i];
		}
	}

	for(int t = 0; t < tEnd; ++t)
	{
		if((t & 31) == 0)
		{
			// rows of the previous band: need every bit up to column t + 31 + b, i.e. words <= ((t+31+b+MARK_PAD)>>5)
			const int needWords = min(Wp, ((t + 31 + b + MARK_PAD) >> 5) + 1);
			const unsigned int needCols = (unsigned int)min(W, (needWords << 5) - MARK_PAD);
			if(band > 0 && lane == 0) // band 0 reads rows the host prepared (zeros, or the upper row tile's hand-off)
				while(*prevProgress < needCols) { __nanosleep(64); }
			__syncwarp();
			__threadfence();
			for(int e = lane; e < b * (needWords - copiedWords); e += 32)
			{
				const int k = e / (needWords - copiedWords), wi = copiedWords + e % (needWords - copiedWords);
				const int rr = r0 - b + k; // absolute row of the predecessor
				uint32_t v = 0u;
				if(rr >= 0)
					v = *(volatile uint32_t*)(asBits + (size_t)rr * Wp + wi);
				asS[(size_t)k * Wp + wi] = v;
			}
			copiedWords = needWords;
			__syncwarp();
		}
		const int c = t - SK * lane;
		const bool colValid = c >= 0 && c < W;
		const bool valid = rowValid && colValid;
		// `pre` words of this pixel
		if(BT)
		{
#pragma unroll
			for(int i = 0; i < PWT; ++i) pw0[i] = pw1[i];
			const int cn = c + 1;
			if(rowValid && cn >= 0 && cn < W)
			{
#pragma unroll
				for(int i = 0; i < PWT; ++i) pw1[i] = pre[(rowBase + cn) * Pw + i];
			}
		}
		const uint32_t* pg = pre + (rowBase + (colValid ? c : 0)) * Pw;
		uint32_t hit = 0u;
		const int posLo = c - b + MARK_PAD; // >= 0 whenever c >= -(MARK_PAD - b); clamp below
		if(valid)
		{
			const int wi = posLo >> 5, sh = posLo & 31;
#pragma unroll
			for(int k = 1; k <= (BT ? BT : 32); ++k)
			{
				if(!BT && k > b) break;
				const uint32_t* rowAbove = asS + (size_t)(lane + b - k) * Wp;
				const uint32_t field = __funnelshift_r(rowAbove[wi], rowAbove[wi + 1], sh);
				const uint32_t sm = BT ? extractBits(pw0, (b - k) * side, side, PWT) : extractBits(pg, (b - k) * side, side, Pw);
				hit |= field & sm;
			}
			const uint32_t smSame = BT ? extractBits(pw0, b * side, b, PWT) : extractBits(pg, b * side, b, Pw);
			hit |= smSame & ownHist;
		}
		const int centreBit = b * side + b;
		const uint32_t strongW = BT ? pw0[centreBit >> 5] : (valid ? pg[centreBit >> 5] : 0u);
		const bool strong = valid && ((strongW >> (centreBit & 31)) & 1u);
		const bool centre = valid && r >= w && r <= H - 1 - w && c >= w && c <= W - 1 - w;
		const bool skipIfMarked = always || hash32((uint32_t)r, (uint32_t)c, seed) < thr;
		const bool act = centre && !(hit != 0u && skipIfMarked);
		const uint32_t as = (act && strong) ? 1u : 0u;
		if(valid)
		{
			active[rowBase + c] = act ? 1 : 0;
			ownHist = (ownHist >> 1) | (as << (b - 1));
			const int pos = c + MARK_PAD;
			curWord |= as << (pos & 31);
			myRow[pos >> 5] = curWord;
			if((pos & 31) == 31 || c == W - 1)
			{
				if(lane >= 32 - b || r >= g.rowEnd - b)
					*(volatile uint32_t*)(asBits + (size_t)r * Wp + (pos >> 5)) = curWord;
				curWord = 0u;
				__threadfence();
			}
		}
		__syncwarp();
		// the last valid row of the band publishes its progress (columns completed) once per word
		if(valid && ((c & 31) == 31 || c == W - 1) && (lane == 31 || r == g.rowEnd - 1))
		{
			__threadfence();
			*(volatile unsigned int*)(progress + band) = (unsigned int)(c + 1);
		}
	}
}

cudaError_t launchMarking(const uint32_t* pre, const LevelGeom& g, float skipProb, uint32_t seed,
		unsigned char* active, uint32_t* asBits, int Wp, unsigned int* progress, cudaStream_t st, int* launches)
{
	const int rows = g.rowEnd - g.rowBeg;
	if(rows <= 0)
		return cudaSuccess;
	const int bands = divUp(rows, 32);
	const size_t smem = (size_t)(32 + g.b) * Wp * 4;
	cudaError_t err;
	if((err = cudaMemsetAsync(progress, 0, sizeof(unsigned int) * bands, st)) != cudaSuccess) return err;
	*launches += 1;
#define LAUNCH_MARK(BT) \
	do { \
		if((err = cudaFuncSetAttribute(k_marking<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err; \
		k_marking<BT><<<bands, 32, smem, st>>>(pre, g, skipProb, seed, active, asBits, Wp, progress); \
	} while(0)
	if(g.b == 6) LAUNCH_MARK(6);
	else if(g.b == 10) LAUNCH_MARK(10);
	else if(g.b == 14) LAUNCH_MARK(14);
	else LAUNCH_MARK(0);
#undef LAUNCH_MARK
	return cudaGetLastError();
}

} // namespace bcdk

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
#include <atomic>
#include <cstdio>

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
// A band waits for the progress word of the band above. CUDA does not promise that CTAs start in blockIdx order, so the
// wait is bounded in time like the systolic kernel's: on expiry the band counts a time-out (the run then fails with an
// error instead of hanging the GPU) and goes on with whatever it has.
__device__ __forceinline__ void boundedProgressWait(const volatile unsigned int* prevProgress, unsigned int needCols, unsigned int ns,
		unsigned int* timeouts)
{
	if(*prevProgress >= needCols)
		return;
	const long long t0 = clock64();
	while(*prevProgress < needCols)
	{
		if(clock64() - t0 > (1ll << 33)) // ~4 s
		{
			if(timeouts) atomicAdd(timeouts, 1u);
			return;
		}
		__nanosleep(ns);
	}
}

template <int BT>
__global__ void __launch_bounds__(32)
k_marking_generic(unsigned int* timeouts, const uint32_t* __restrict__ pre, int prePitch, LevelGeom g, float skipProb, uint32_t seed,
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
	const uint32_t thr = (uint32_t)(fminf(skipProb, 0.99999994f) * 4294967296.f);
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
				boundedProgressWait(prevProgress, needCols, 100, timeouts);
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
		const bool skipIfMarked = always || hash32((uint32_t)r, (uint32_t)c, seed) <= thr;
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
// Fast kernel (b = 6, 10, 14): PX = 4 pixels per lane-step. Lane l handles columns C..C+3, C = 4 (t - SKS l): the lane
// above is 4 SKS >= b + 4 columns ahead, which covers the b-column look-ahead of the last pixel of the block.
// The single resident warp is latency-bound (every instruction waits ~4 cycles for its predecessor), so the kernel
// is written for minimum instruction count per step:
//   * the per-pixel input comes in the `pre2` layout written by k_symmetrize (fastMarkWords(b) words per pixel): the
//     (2b+1)-bit mask of each row above, already shifted left by (x & 3) and packed FPW fields per word, so that
//     hit(pixel i) = OR_a (win2[a] & pre2[i][a]) costs one LOP3 per word — win2 packs the funnel-shifted AS windows
//     of FPW rows the same way (one PRMT per pair of rows for b = 6);
//   * one LDS.128 per pixel (b = 6) feeds it, from a cp.async-fed double-buffered tile (16-byte copies);
//   * the only sequential part is the same-row chain (4 x ~6 dependent ALU ops).
// Total dependent steps ~ W/4 + SKS*H.
// ---------------------------------------------------------------------------------------------------------------
template <int BT, bool ALWAYS>
__global__ void __launch_bounds__(32)
k_marking_fast(unsigned int* timeouts, const uint32_t* __restrict__ pre2, int prePitch, LevelGeom g, float skipProb, uint32_t seed,
		uint32_t* __restrict__ actBits, uint32_t* asBits, int Wp, unsigned int* progress)
{
	extern __shared__ __align__(16) uint32_t smem[];
	constexpr int b = BT;
	constexpr int PX = 4;
	constexpr int SKS = (BT + PX + PX - 1) / PX; // steps of skew: 4*SKS >= b + 4
	constexpr int SKC = PX * SKS;                // columns of skew
	constexpr int FW = 2 * BT + 1 + 3;           // bits of a shifted row field
	constexpr int FPW = FW <= 16 ? 2 : 1;        // fields per word
	constexpr int NF = (BT + FPW - 1) / FPW;     // words holding the rows above
	constexpr int PW2 = (NF + 1 + 3) & ~3;       // words per pixel (the word NF holds the same-row bits + strong flag)
	constexpr int CHS = 8;                       // steps per chunk (32 pixels per row)
	constexpr int RS = CHS * PX * PW2 + 4;       // tile row stride in words (== 4 mod 32: LDS.128 conflict-free)
	constexpr int UNITS = CHS * PX * PW2 / 4;    // 16-byte units per tile row
	constexpr int UPL = UNITS * (32 / CHS) / 32; // units per lane per step
	static_assert(2 * BT + PX <= 32, "window must fit one funnel shift");
	static_assert((UNITS * (32 / CHS)) % 32 == 0, "units must split evenly over the lanes");
	const int W = g.W, H = g.H, w = g.w;
	uint32_t* asS = smem;                                                // [(32 + b)][Wp]
	uint32_t* preS = smem + (((size_t)(32 + b) * Wp + 3) & ~(size_t)3);  // [2][32][RS], 16-byte aligned
	const int lane = threadIdx.x;
	const int band = blockIdx.x;
	const int r0 = g.rowBeg + band * 32;
	const int r = r0 + lane;
	const bool rowValid = r < g.rowEnd;
	const bool rowCentre = rowValid && r >= w && r <= H - 1 - w;
	const uint32_t thr = (uint32_t)(fminf(skipProb, 0.99999994f) * 4294967296.f);
	const int lastLane = min(31, g.rowEnd - 1 - r0);
	const int nBlocks = (W + PX - 1) / PX;
	const int tEnd = nBlocks + SKS * 31;
	const int nChunks = (tEnd + CHS - 1) / CHS;

	for(int e = lane; e < (32 + b) * Wp; e += 32)
		asS[e] = 0u;

	// per-lane constants of the prefetch: unit u = lane + 32 i of the 4 rows issued per step
	int uRow[UPL], uWord[UPL];
#pragma unroll
	for(int i = 0; i < UPL; ++i)
	{
		const int u = lane + 32 * i;
		uRow[i] = u / UNITS;
		uWord[i] = 4 * (u % UNITS);
	}
	// rows [lFirst, lFirst + 4) of chunk q: row l covers pixels [32 q - SKC l, +32)
	auto issueRows = [&](int q, int lFirst)
	{
#pragma unroll
		for(int i = 0; i < UPL; ++i)
		{
			const int l = lFirst + uRow[i];
			const int rr = r0 + l;
			const int px = CHS * PX * q - SKC * l + uWord[i] / PW2; // a unit never straddles a pixel: PW2 % 4 == 0
			const bool inside = rr < g.rowEnd && px >= 0 && px < prePitch;
			const uint32_t* src = inside ? pre2 + ((size_t)rr * prePitch + px) * PW2 + (uWord[i] % PW2) : pre2;
			uint32_t* dst = preS + (size_t)(q & 1) * 32 * RS + (size_t)l * RS + uWord[i];
			__pipeline_memcpy_async(dst, src, 16, inside ? 0 : 16);
		}
	};
	for(int l0 = 0; l0 < 32; l0 += 32 / CHS)
		issueRows(0, l0);
	__pipeline_commit();
	__pipeline_wait_prior(0);
	__syncwarp();

	uint32_t* myRow = asS + (size_t)(lane + b) * Wp;
	uint32_t curWord = 0u, curAct = 0u, ownHist = 0u;
	int copiedWords = 0, flushedWords = 0;
	const volatile unsigned int* prevProgress = progress + (band > 0 ? band - 1 : 0);
	const size_t rowBase = (size_t)(rowValid ? r : 0);

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
			// predecessor rows: lane 0 reaches column 4 (t + CHS) - 1 + b within this chunk
			const int lastCol = PX * (t + CHS) - 1 + b;
			const int needWords = min(Wp, ((lastCol + MARK_PAD) >> 5) + 1);
			if(needWords > copiedWords)
			{
				const unsigned int needCols = (unsigned int)min(W, (needWords << 5) - MARK_PAD);
				if(band > 0 && lane == 0)
					boundedProgressWait(prevProgress, needCols, 32, timeouts);
				__syncwarp();
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
			issueRows(q + 1, s * (32 / CHS));
		if(s == CHS - 1)
			__pipeline_commit();

		const int C = PX * (t - SKS * lane);
		const bool valid = rowValid && C >= 0 && C < W;
		const uint4* src = reinterpret_cast<const uint4*>(preS + (size_t)(q & 1) * 32 * RS + (size_t)lane * RS + s * PX * PW2);
		uint32_t win2[NF];
		{
			const int posLo = (valid ? C : 0) - b + MARK_PAD;
			const int wi = posLo >> 5, sh = posLo & 31;
			uint32_t win[BT];
#pragma unroll
			for(int k = 1; k <= BT; ++k)
			{
				const uint32_t* rowAbove = asS + (size_t)(lane + b - k) * Wp;
				win[k - 1] = __funnelshift_r(rowAbove[wi], rowAbove[wi + 1], sh); // columns C-b .. C-b+31
			}
#pragma unroll
			for(int a = 0; a < NF; ++a)
			{
				if(FPW == 2)
					win2[a] = (2 * a + 1 < BT) ? __byte_perm(win[2 * a], win[(2 * a + 1 < BT) ? 2 * a + 1 : 2 * a], 0x5410) : (win[2 * a] & 0xffffu);
				else
					win2[a] = win[a];
			}
		}
		uint32_t asNibble = 0u, actNibble = 0u;
#pragma unroll
		for(int i = 0; i < PX; ++i)
		{
			uint32_t pw[PW2];
#pragma unroll
			for(int v4 = 0; v4 < PW2 / 4; ++v4)
			{
				const uint4 v = src[i * (PW2 / 4) + v4];
				pw[4 * v4 + 0] = v.x; pw[4 * v4 + 1] = v.y; pw[4 * v4 + 2] = v.z; pw[4 * v4 + 3] = v.w;
			}
			uint32_t hit = pw[NF] & ownHist & ((1u << BT) - 1u); // same row: bits 0..b-1 = columns c-b..c-1
#pragma unroll
			for(int a = 0; a < NF; ++a)
				hit |= win2[a] & pw[a];
			const int c = C + i;
			const bool centre = rowCentre && valid && c >= w && c <= W - 1 - w;
			const bool skipIfMarked = ALWAYS || hash32((uint32_t)r, (uint32_t)c, seed) <= thr;
			const bool act = centre && !(hit != 0u && skipIfMarked);
			const uint32_t as = (act && ((pw[NF] >> 16) & 1u)) ? 1u : 0u;
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

// ---------------------------------------------------------------------------------------------------------------
// Systolic kernel (b = 6): the same wavefront (lane l = row r0 + l, 4 pixels per step, lane l at block t - 3 l), but the
// AS bits never go through shared memory inside a band. The 16-column windows of the six rows above travel DOWN the
// warp as a three-word packet, one shuffle per word and step:
//     P_l(t)     = fields of rows l-1 .. l-6 at columns [C-6, C+9],  C = 4 (t - 3 l)        (what lane l consumes at step t)
//     P_l+1(t+1) = { own row at [C-14, C+1] (bits 14..29 of the lane's 32-column history), rows l-1 .. l-5 of P_l(t-2) }
// so a step costs 3 SHFL + 3 PRMT instead of 12 LDS + 6 SHF + 3 PRMT + STS + a warp barrier. The per-pixel input
// (`pre2`, 16 bytes per pixel; bit 17 of word 3 = "is a patch centre", bit 16 = strong) is read straight from global
// memory / L2 eight steps ahead into registers (every lane streams its own row: 64 contiguous bytes per step).
// A CTA holds NW consecutive bands, one warp each: lane 31 of a band hands its packets to lane 0 of the next band
// through a tagged 16-byte slot per step in shared memory (tag = step + 1, written and read as one 128-bit access).
// CTAs — and, over NVLink, the row tiles of neighbouring GPUs — are chained through a ring in global memory: the last
// six lanes of a CTA's last band publish the four AS bits of every block they finish as ONE 32-bit word that carries
// its own epoch tag (bits 8..31), so no fence, no progress counter and no clearing is needed; a feeder warp of the next
// CTA polls the words (volatile loads) and turns them into lane-0 packets. The first band of a launch is fed from the AS
// bit-rows above the launch range (zeros, or the strip received from the upper row tile) unless an inbox ring is given.
// The kernel only resolves the sequential part, AS(p) = active(p) and strong(p): the AS bit-rows are written as whole
// 32-bit words, all lanes at the same step (every 8 steps: no divergence), and k_active_from_as then derives
// active(p) = centre(p) and no AS bit under the earlier part of S(p) for every pixel in parallel.
// Every wait is bounded in time: on a time-out the CTA raises `dead`, counts it and finishes with whatever it has
// (the run then fails with an error; the GPU does not hang).
// ---------------------------------------------------------------------------------------------------------------
constexpr int SYS_PF = 4;                           // steps of read-ahead (= the unroll of the step loop: the loop body must fit
                                                    // the ~6 KB L0 instruction cache — a lone latency-bound warp cannot hide fetch misses)
constexpr long long SYS_TIMEOUT_CYCLES = 1ll << 33; // ~4 s

struct SysArgs
{
	const uint32_t* pre2; int prePitch;
	LevelGeom g; float skipProb; uint32_t seed;
	uint32_t* actBits; uint32_t* asBits; int Wp;
	uint32_t* ring; int ringPitch;   // hand-off ring of this launch: [CTAs][6 rows][ringPitch blocks]
	const uint32_t* inRing;          // non-null: the first band is fed from this ring [6][ringPitch] (written by the upper row tile)
	uint32_t* outRing;               // non-null: the last band of the launch publishes here [6][ringPitch] (the lower row tile's inbox)
	uint32_t tag;                    // 0x80000000 | epoch << 8: valid marker of the words of `ring` written by this launch
	uint32_t tagPeer;                // same for inRing / outRing (an epoch the neighbouring row tiles agree on)
	unsigned int* timeouts;
	int nBlocks, tEnd, bands;
	int trace;                       // BCD_B200_MARK_TRACE=1: every band prints its timeline (cycles since its CTA started)
};

__device__ __forceinline__ uint32_t ldVolatileGlobal(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void stVolatileGlobal(uint32_t* p, uint32_t v)
{
	asm volatile("st.volatile.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ldVolatileShared4(const uint4* p)
{
	uint4 v;
	asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
			: "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
	return v;
}
__device__ __forceinline__ void stVolatileShared4(uint4* p, uint4 v)
{
	asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"((uint32_t)__cvta_generic_to_shared(p)),
			"r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

constexpr int SYS_PAD = 96;                         // slack entries in front of every ring row / packet row: the producers
                                                    // store unconditionally, indices before the first block land here
__host__ __device__ inline int sysPktStride(int tEnd) { return tEnd + SYS_PAD; }

__device__ __forceinline__ uint4 ldVolatileSharedAddr4(uint32_t addr)
{
	uint4 v;
	asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ void stVolatileSharedAddr4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
	asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// cold path of the packet wait (the producer is normally far ahead): bounded spin until the packets of the steps
// t .. t + SYS_PF - 1 (16 bytes apart from addr) carry their tags
__device__ __noinline__ void sysWaitPackets(uint32_t addr, uint32_t t, volatile int* dead, long long tStart, unsigned int* timeouts)
{
	for(;;)
	{
		bool all = true;
#pragma unroll
		for(int s = 0; s < SYS_PF; ++s)
			all = all && ldVolatileSharedAddr4(addr + 16u * s).w == t + (uint32_t)s + 1u;
		if(all || *dead)
			return;
		if(clock64() - tStart > SYS_TIMEOUT_CYCLES) { *dead = 1; atomicAdd(timeouts, 1u); }
		__nanosleep(20);
	}
}
template <typename T> __device__ __forceinline__ T opaque(T v)
{
	asm volatile("" : "+r"(v)); // keeps a loop-invariant value in its register instead of recomputing it every step
	return v;
}

template <int NW, bool ALWAYS>
__global__ void __launch_bounds__(32 * (NW + 1), 1)
k_marking_sys6(const SysArgs a)
{
	constexpr int b = 6, PX = 4, SKS = 3;
	extern __shared__ __align__(16) uint32_t smem[];
	const int tEnd = a.tEnd, nBlocks = a.nBlocks;
	const int pktStride = sysPktStride(tEnd);
	uint4* pktS = reinterpret_cast<uint4*>(smem);                         // [NW + 1][pktStride] lane-0 packets of the CTA's bands
	volatile int* dead = reinterpret_cast<volatile int*>(pktS + (size_t)(NW + 1) * pktStride);
	const LevelGeom& g = a.g;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int cta = blockIdx.x;
	for(int e = threadIdx.x; e < (NW + 1) * pktStride; e += blockDim.x)
		pktS[e] = make_uint4(0u, 0u, 0u, 0u);
	if(threadIdx.x == 0)
		*dead = 0;
	__syncthreads();

	if(warp == NW)
	{
		// ---- feeder of the CTA's first band: packets P_0(t) = fields of the six rows above at columns [4t-6, 4t+9]
		const int band = cta * NW;
		const int r0 = g.rowBeg + band * 32;
		const uint32_t* src = band > 0 ? a.ring + (size_t)(cta - 1) * 6 * a.ringPitch : a.inRing;
		const uint32_t tag = band > 0 ? a.tag : a.tagPeer;
		uint4* dst = pktS + SYS_PAD;
		const long long t0 = clock64();
		for(int T = 0; T < tEnd; T += 32)
		{
			const int t = T + lane;
			if(t >= tEnd)
				break;
			uint32_t f[6] = {0u, 0u, 0u, 0u, 0u, 0u};
			if(t < nBlocks + 2)
			{
				if(src == nullptr)
				{
					const int posLo = PX * t - b + MARK_PAD;
					const int wi = posLo >> 5, sh = posLo & 31;
#pragma unroll
					for(int k = 1; k <= 6; ++k)
					{
						const int rr = r0 - k;
						if(rr >= 0)
							f[k - 1] = __funnelshift_r(a.asBits[(size_t)rr * a.Wp + wi], a.asBits[(size_t)rr * a.Wp + wi + 1], sh) & 0xffffu;
					}
				}
				else
				{
					// blocks t-2 .. t+2 cover columns 4t-8 .. 4t+11; a field is bits 2..17 of the 20-bit concatenation.
					// Row 0 (the band's last row) is written last: wait for its newest word first, then verify all of them.
					const uint32_t* s0 = src + SYS_PAD;
					const int newest = min(t + 2, nBlocks - 1);
					while((ldVolatileGlobal(s0 + newest) & 0xffffff00u) != tag && !*dead)
					{
						if(clock64() - t0 > SYS_TIMEOUT_CYCLES) { *dead = 1; atomicAdd(a.timeouts, 1u); }
						__nanosleep(100);
					}
					uint32_t v[6][5];
					bool all = false;
					while(!all)
					{
						all = true;
#pragma unroll
						for(int k = 0; k < 6; ++k)
#pragma unroll
							for(int j = 0; j < 5; ++j)
							{
								const int blk = t - 2 + j;
								uint32_t x = tag;
								if(blk >= 0 && blk < nBlocks)
									x = ldVolatileGlobal(s0 + (size_t)k * a.ringPitch + blk);
								v[k][j] = x;
								all = all && (x & 0xffffff00u) == tag;
							}
						if(!all)
						{
							if(*dead) break;
							if(clock64() - t0 > SYS_TIMEOUT_CYCLES) { *dead = 1; atomicAdd(a.timeouts, 1u); break; }
							__nanosleep(100);
						}
					}
#pragma unroll
					for(int k = 0; k < 6; ++k)
					{
						const uint32_t cat = (v[k][0] & 15u) | ((v[k][1] & 15u) << 4) | ((v[k][2] & 15u) << 8) | ((v[k][3] & 15u) << 12) | ((v[k][4] & 15u) << 16);
						f[k] = (cat >> 2) & 0xffffu;
					}
				}
			}
			stVolatileShared4(dst + t, make_uint4(f[0] | (f[1] << 16), f[2] | (f[3] << 16), f[4] | (f[5] << 16), (uint32_t)t + 1u));
		}
		return;
	}

	// ---- compute warps: one band each
	const int band = cta * NW + warp;
	if(band >= a.bands)
		return;
	const int r0 = g.rowBeg + band * 32;
	const int lastLane = min(31, g.rowEnd - 1 - r0);
	const int r = r0 + lane;
	const bool rowValid = r < g.rowEnd;
	const uint32_t thr = (uint32_t)(fminf(a.skipProb, 0.99999994f) * 4294967296.f);
	const size_t rowBase = (size_t)(rowValid ? r : 0) * a.Wp;
	const uint32_t pktBase = (uint32_t)__cvta_generic_to_shared(pktS);
	// my lane-0 packets: entry t at inAddr + 16 t
	uint32_t inAddr = opaque(pktBase + 16u * (uint32_t)(warp * pktStride + SYS_PAD));
	// lane 31 hands its packets to the next band of the CTA ("lane 32" is 96 steps behind: its step t - 95 consumes my
	// packet of step t); every other lane — and the CTA's last band — stores into the spare row NW, which nobody reads
	const bool handsOn = lane == 31 && warp + 1 < NW && band + 1 < a.bands;
	uint32_t outAddr = opaque(pktBase + 16u * (uint32_t)((handsOn ? warp + 1 : NW) * pktStride + SYS_PAD - 95));
	// hand-off ring: the CTA's last band feeds the next CTA; the launch's last band feeds the lower row tile (if any).
	// In such a band every lane stores every step: the six hand-off rows into the ring, the others into the spare row.
	const int ringRow = lastLane - lane;
	uint32_t* ringBase = nullptr;
	uint32_t myTag = a.tag;
	if(band == a.bands - 1)
	{
		ringBase = a.outRing;
		myTag = a.tagPeer;
	}
	else if(warp == NW - 1)
		ringBase = a.ring + (size_t)cta * 6 * a.ringPitch;
	const bool warpEmits = ringBase != nullptr; // warp-uniform
	uint32_t* ringPtr = a.ring + (size_t)gridDim.x * 6 * a.ringPitch; // spare row
	if(warpEmits && rowValid && ringRow >= 0 && ringRow < 6)
		ringPtr = ringBase + (size_t)ringRow * a.ringPitch;
	ringPtr += SYS_PAD - SKS * lane; // + t = + SYS_PAD + block index
	// per-lane constant of the word stores: after a step t with (t+1) % 8 == 0 the lane has completed n = 4(t+1) - 12 lane
	// columns; the newest complete 32-bit word starts 32 - (n mod 32) bits into the 64-bit history (n mod 32 == 0: 32)
	const int mW = (-12 * lane) & 31;
	const int shW = mW ? 32 - mW : 32;
	const bool isLane0 = opaque(lane) == 0;

	// per-pixel input, SYS_PF steps ahead, band-major: the lanes of the band read 512 contiguous bytes per load. Steps
	// whose block lies outside the row are invalidated when consumed.
	const int nBlocksU = opaque(nBlocks);
	const uint32_t rowInval = opaque(rowValid ? 0u : 0x30000u);
	const uint4* pbase = reinterpret_cast<const uint4*>(a.pre2) + sysPreIndex(band, 0, 0, lane, tEnd);
	const int lastStep = opaque(tEnd - 1);
	uint4 buf[SYS_PF][PX];
	auto fetch = [&](int slot, int tf)
	{
		const uint4* p4 = pbase + (size_t)min(tf, lastStep) * (PX * 32);
#pragma unroll
		for(int i = 0; i < PX; ++i)
			buf[slot][i] = __ldg(p4 + i * 32);
	};
#pragma unroll
	for(int s = 0; s < SYS_PF; ++s)
		fetch(s, s);

	uint32_t hh = 0u, sLo = 0u;             // AS bits of my row: 64-bit history, bit 31 of hh = the newest column
	uint32_t o0 = 0u, o1 = 0u, o2 = 0u;     // packet for the lane below
	uint32_t d1a = 0u, d1b = 0u, d1c = 0u, d2a = 0u, d2b = 0u, d2c = 0u; // my packets of the last two steps
	const long long tStart = clock64();
	long long tFirst = 0;
	int nWaits = 0;
	int tb = opaque(-SKS * lane);           // block index of the current step

#pragma unroll 1
	for(int tc = 0; tc < tEnd; tc += SYS_PF)
	{
		// lane-0 packets of the SYS_PF steps of this chunk (one tag test for all of them)
		uint4 pk[SYS_PF];
		uint32_t bad = 0u;
#pragma unroll
		for(int s = 0; s < SYS_PF; ++s)
		{
			pk[s] = ldVolatileSharedAddr4(inAddr + 16u * s);
			bad |= pk[s].w ^ (uint32_t)(tc + s + 1);
		}
		if(bad != 0u)
		{
			sysWaitPackets(inAddr, (uint32_t)tc, dead, tStart, a.timeouts);
#pragma unroll
			for(int s = 0; s < SYS_PF; ++s)
				pk[s] = ldVolatileSharedAddr4(inAddr + 16u * s);
			++nWaits;
			if(a.trace && tc == 0) tFirst = clock64();
		}
#pragma unroll
		for(int s = 0; s < SYS_PF; ++s)
		{
			const int t = tc + s;
			uint32_t i0 = __shfl_up_sync(0xffffffffu, o0, 1);
			uint32_t i1 = __shfl_up_sync(0xffffffffu, o1, 1);
			uint32_t i2 = __shfl_up_sync(0xffffffffu, o2, 1);
			if(isLane0) { i0 = pk[s].x; i1 = pk[s].y; i2 = pk[s].z; }

			// outside the row (or a row beyond the launch range): "not a centre, not strong"
			const uint32_t inval = (unsigned)(tb + s) < (unsigned)nBlocksU ? rowInval : 0x30000u;
			const uint32_t hhBefore = hh;
#pragma unroll
			for(int i = 0; i < PX; ++i)
			{
				const uint4 pw = buf[s][i];
				uint32_t above = (i0 & pw.x) | (i1 & pw.y) | (i2 & pw.z);
				uint32_t same = pw.w << 26; // bits 0..5 of word 3 = columns c-6..c-1 -> bits 26..31 of the history
				if(!ALWAYS)
				{
					const uint32_t keep = hash32((uint32_t)r, (uint32_t)(PX * (tb + s) + i), a.seed) <= thr ? 0xffffffffu : 0u;
					above &= keep;
					same &= keep;
				}
				const uint32_t flags = (~pw.w & 0x30000u) | inval; // bit 17: not a patch centre, bit 16: not strong
				const uint32_t hitS = above | flags | (same & hh);  // zero <=> patch centre, not marked, strong: AS(p)
				hh = (hh >> 1) | (hitS == 0u ? 0x80000000u : 0u);
			}
			fetch(s, tc + s + SYS_PF);
			sLo = __funnelshift_r(sLo, hhBefore, 4);
			// packet for the lane below (consumed at its step t + 1, block C - 8): my row at [C-14, C+1] + my packet of step t - 2
			o0 = __byte_perm(hh >> 14, d2a, 0x5410);
			o1 = __byte_perm(d2a, d2b, 0x5432);
			o2 = __byte_perm(d2b, d2c, 0x5432);
			d2a = d1a; d2b = d1b; d2c = d1c;
			d1a = i0; d1b = i1; d1c = i2;
			stVolatileSharedAddr4(outAddr + 16u * s, o0, o1, o2, (uint32_t)(t - 95) + 1u);
			if(warpEmits)
				stVolatileGlobal(ringPtr + s, myTag | (hh >> 28));
		}
		tb += SYS_PF;
		inAddr += 16u * SYS_PF;
		outAddr += 16u * SYS_PF;
		ringPtr += SYS_PF;
		// whole words of the AS bit-row, all lanes at once (every 8 steps = 32 columns)
		const int n = PX * (tc + SYS_PF) - 4 * SKS * lane;
		const int wd = ((n + 32) >> 5) - 1;
		if(((tc + SYS_PF) & 7) == 0 && rowValid && wd >= 1 && wd < a.Wp)
			a.asBits[rowBase + wd] = __funnelshift_rc(sLo, hh, shW);
	}
	if(a.trace && lane == 0)
	{
		unsigned int smid;
		asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
		unsigned long long gt;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
		printf("[mark trace] band %3d cta %3d sm %3u  first packet at %9lld  end at %9lld cycles  (%d steps, %d waits)  globaltimer end %llu ns\n",
				band, cta, smid, tFirst ? tFirst - tStart : 0ll, clock64() - tStart, tEnd, nWaits, gt);
	}
	// the next band's remaining steps (it runs 96 steps behind) lie beyond the right edge of the image: empty packets
	if(handsOn)
		for(int t0 = tEnd - 95; t0 < tEnd; ++t0)
			stVolatileShared4(pktS + (size_t)(warp + 1) * pktStride + SYS_PAD + t0, make_uint4(0u, 0u, 0u, (uint32_t)t0 + 1u));
}

// active(p) from the finished AS bit-rows (second, fully parallel half of the systolic path): a thread owns a block of 4
// pixels, 8 consecutive lanes assemble one 32-bit word of the act bit-row.
template <bool ALWAYS>
__global__ void __launch_bounds__(256)
k_active_from_as(const uint32_t* __restrict__ pre2, int prePitch, LevelGeom g, float skipProb, uint32_t seed,
		const uint32_t* __restrict__ asBits, uint32_t* __restrict__ actBits, int Wp)
{
	constexpr int b = 6, PX = 4;
	const int lane = threadIdx.x;
	const int blk = blockIdx.x * 32 + lane;
	const int r = g.rowBeg + blockIdx.y * blockDim.y + threadIdx.y;
	const int C = PX * blk;
	const bool ok = r < g.rowEnd && C < prePitch;
	const uint32_t thr = (uint32_t)(fminf(skipProb, 0.99999994f) * 4294967296.f);
	uint32_t nib = 0u;
	if(ok)
	{
		const int posLo = C - b + MARK_PAD;
		const int wi = posLo >> 5, sh = posLo & 31;
		uint32_t win[6];
#pragma unroll
		for(int k = 1; k <= 6; ++k)
		{
			const int rr = r - k;
			win[k - 1] = rr >= 0 ? __funnelshift_r(asBits[(size_t)rr * Wp + wi], asBits[(size_t)rr * Wp + wi + 1], sh) : 0u;
		}
		const uint32_t own = __funnelshift_r(asBits[(size_t)r * Wp + wi], asBits[(size_t)r * Wp + wi + 1], sh); // bit j = column C-6+j
		const uint32_t w0 = __byte_perm(win[0], win[1], 0x5410), w1 = __byte_perm(win[2], win[3], 0x5410), w2 = __byte_perm(win[4], win[5], 0x5410);
		const int rr = r - g.rowBeg, ln = rr & 31;
		const uint4* p4 = reinterpret_cast<const uint4*>(pre2) + sysPreIndex(rr >> 5, blk + 3 * ln, 0, ln, sysTEnd(g.W));
#pragma unroll
		for(int i = 0; i < PX; ++i)
		{
			const uint4 pw = __ldg(p4 + i * 32);
			const uint32_t hit = (w0 & pw.x) | (w1 & pw.y) | (w2 & pw.z) | (pw.w & 0x3fu & (own >> i));
			const bool centre = (pw.w >> 17) & 1u;
			const bool skipIfMarked = ALWAYS || hash32((uint32_t)r, (uint32_t)(C + i), seed) <= thr;
			const bool act = centre && !(hit != 0u && skipIfMarked);
			nib |= (act ? 1u : 0u) << i;
		}
	}
	uint32_t v = nib << (4 * (lane & 7));
	v |= __shfl_xor_sync(0xffffffffu, v, 1);
	v |= __shfl_xor_sync(0xffffffffu, v, 2);
	v |= __shfl_xor_sync(0xffffffffu, v, 4);
	if((lane & 7) == 0 && r < g.rowEnd && C < g.W)
		actBits[(size_t)r * Wp + ((C + MARK_PAD) >> 5)] = v;
}

// Row tiles, systolic kernel: the AS bits of the upper tile's last six rows arrived as ring words (4 bits per block of 4
// columns) in this tile's inbox; k_active_from_as wants them as bit-rows above the first owned row.
__global__ void k_ring_to_asbits(const uint32_t* __restrict__ inbox, int ringPitch, int W, uint32_t* __restrict__ asBits, int Wp, int rowBeg)
{
	const int wi = blockIdx.x * blockDim.x + threadIdx.x;
	const int k = blockIdx.y;          // ring row k = image row rowBeg - 1 - k
	const int row = rowBeg - 1 - k;
	if(wi >= Wp || row < 0)
		return;
	const uint32_t* src = inbox + (size_t)k * ringPitch + SYS_PAD;
	uint32_t v = 0u;
	for(int bit = 0; bit < 32; ++bit)
	{
		const int x = wi * 32 + bit - MARK_PAD;
		if(x >= 0 && x < W)
			v |= ((src[x >> 2] >> (x & 3)) & 1u) << bit;
	}
	asBits[(size_t)row * Wp + wi] = v;
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

// epoch of the local hand-off ring: process-wide and monotonic, so a ring word left in (possibly recycled) device memory
// by any earlier launch can never carry the tag of the current one
static uint32_t nextLocalEpoch()
{
	static std::atomic<uint32_t> e{0};
	return ++e;
}

bool g_forceGenericMarking = false; // test hook (BCD_B200_GENERIC_MARKING=1)
int g_markingKernel = 1;            // b = 6: 1 = systolic kernel (default), 0 = shared-memory wavefront (BCD_B200_MARKING=v4)
int g_markingTrace = 0;              // BCD_B200_MARK_TRACE=1
int g_markingBandsPerCta = 4;       // systolic kernel: bands (warps) per CTA, 1 / 2 / 4 / 8 (BCD_B200_MARKING_NW)

// layout of the marking input k_symmetrize must write: 0 generic, 1 `pre2` row-major (k_marking_fast), 2 `pre2` band-major
// (k_marking_sys6)
int markingPreLayout(int b)
{
	if(g_forceGenericMarking || !(b == 6 || b == 10 || b == 14))
		return 0;
	return (b == 6 && g_markingKernel == 1) ? 2 : 1;
}
size_t markingPreBytes(int W, int H, int b, int Pw)
{
	const size_t rowMajor = (size_t)(H > 0 ? H : 1) * prePitchOf(W) * (size_t)(Pw > fastMarkWords(b) ? Pw : fastMarkWords(b)) * 4 + 64;
	const size_t bandMajor = b == 6 ? sysPreBytes(W, H) : 0;
	return rowMajor > bandMajor ? rowMajor : bandMajor;
}

// true when launchMarking runs the systolic kernel for this geometry: row tiles then chain their launches through ring
// words in peer memory (ringIn / ringOut) instead of exchanging AS bit-row strips
static int sysBandsPerCta(const LevelGeom& g)
{
	const int rows = g.rowEnd - g.rowBeg;
	if(rows <= 0 || g.b != 6 || g_forceGenericMarking || g_markingKernel != 1)
		return 0;
	const int bands = divUp(rows, 32);
	const int tEnd = sysTEnd(g.W);
	static const int cand[4] = {8, 4, 2, 1};
	for(int i = 0; i < 4; ++i)
	{
		const int nw = cand[i];
		if(nw > 1 && g_markingBandsPerCta < nw)
			continue;
		const size_t smem = (size_t)(nw + 1) * sysPktStride(tEnd) * 16 + 16;
		if(smem <= 112 * 1024 && divUp(bands, nw) <= 2 * currentSmCount())
			return nw;
	}
	return 0;
}
bool markingUsesRing(const LevelGeom& g) { return sysBandsPerCta(g) > 0; }

// tile_setup: loads the kernels a frame may use although the warm-up pass did not launch them (the ring converter, and the
// systolic kernels of the other skipping mode — the skipping probability may change from frame to frame)
cudaError_t launchMarkingWarmup(cudaStream_t st)
{
	k_ring_to_asbits<<<dim3(1, 6), 64, 0, st>>>(nullptr, 0, 0, nullptr, 0, 0);
	cudaError_t err = cudaGetLastError();
	cudaFuncAttributes fa;
#define TOUCH(NWV) \
	if(err == cudaSuccess) err = cudaFuncGetAttributes(&fa, k_marking_sys6<NWV, true>); \
	if(err == cudaSuccess) err = cudaFuncGetAttributes(&fa, k_marking_sys6<NWV, false>);
	TOUCH(1) TOUCH(2) TOUCH(4) TOUCH(8)
#undef TOUCH
	if(err == cudaSuccess) err = cudaFuncGetAttributes(&fa, k_active_from_as<true>);
	if(err == cudaSuccess) err = cudaFuncGetAttributes(&fa, k_active_from_as<false>);
	return err;
}

cudaError_t launchMarking(const uint32_t* pre, int prePitch, const LevelGeom& g, float skipProb, uint32_t seed,
		uint32_t* actBits, uint32_t* asBits, int Wp, unsigned int* progress, cudaStream_t st, int* launches,
		unsigned int* timeouts, uint32_t epoch, const uint32_t* ringIn, uint32_t* ringOut)
{
	const int rows = g.rowEnd - g.rowBeg;
	if(rows <= 0)
		return cudaSuccess;
	const int bands = divUp(rows, 32);
	cudaError_t err;
	if(g.b == 6 && !g_forceGenericMarking && g_markingKernel == 1)
	{
		SysArgs a;
		a.pre2 = pre; a.prePitch = prePitch; a.g = g; a.skipProb = skipProb; a.seed = seed;
		a.actBits = actBits; a.asBits = asBits; a.Wp = Wp;
		a.ringPitch = markingRingPitch(g.W);
		a.ring = progress + 4; // scratch layout (markingScratchWords): 4 spare words, then the hand-off ring
		a.inRing = ringIn; a.outRing = ringOut;
		a.tag = 0x80000000u | ((nextLocalEpoch() & 0x7fffffu) << 8);
		a.tagPeer = 0x80000000u | ((epoch & 0x7fffffu) << 8);
		a.timeouts = timeouts ? timeouts : progress;
		a.nBlocks = (g.W + 3) / 4;
		a.tEnd = (a.nBlocks + 3 * 31 + 8 + 7) & ~7;
		a.bands = bands;
		a.trace = g_markingTrace;
		const bool always = skipProb >= 1.f;
#define LAUNCH_SYS(NWV) \
		do { \
			const size_t smem = (size_t)((NWV) + 1) * sysPktStride(a.tEnd) * 16 + 16; \
			if(smem <= 112 * 1024 && divUp(bands, NWV) <= 2 * currentSmCount()) { \
				auto kern = always ? k_marking_sys6<NWV, true> : k_marking_sys6<NWV, false>; \
				if((err = allowDynamicSmem((const void*)kern, smem)) != cudaSuccess) return err; \
				kern<<<divUp(bands, NWV), 32 * ((NWV) + 1), smem, st>>>(a); \
				if((err = cudaGetLastError()) != cudaSuccess) return err; \
				if(ringIn) { \
					k_ring_to_asbits<<<dim3(divUp(Wp, 64), 6), 64, 0, st>>>(ringIn, a.ringPitch, g.W, asBits, Wp, g.rowBeg); \
					*launches += 1; \
				} \
				dim3 ablock(32, 8), agrid(divUp(divUp(prePitch, 4), 32), divUp(rows, 8)); \
				if(always) k_active_from_as<true><<<agrid, ablock, 0, st>>>(pre, prePitch, g, skipProb, seed, asBits, actBits, Wp); \
				else k_active_from_as<false><<<agrid, ablock, 0, st>>>(pre, prePitch, g, skipProb, seed, asBits, actBits, Wp); \
				*launches += 2; \
				return cudaGetLastError(); \
			} \
		} while(0)
		if(g_markingBandsPerCta >= 8) LAUNCH_SYS(8);
		if(g_markingBandsPerCta >= 4) LAUNCH_SYS(4);
		if(g_markingBandsPerCta >= 2) LAUNCH_SYS(2);
		LAUNCH_SYS(1);
#undef LAUNCH_SYS
	}
	if((err = cudaMemsetAsync(progress, 0, sizeof(unsigned int) * bands, st)) != cudaSuccess) return err;
	*launches += 1;
	const size_t asWords = ((size_t)(32 + g.b) * Wp + 3) & ~(size_t)3;
#define LAUNCH_FAST(BT) \
	do { \
		const size_t smem = (asWords + (size_t)2 * 32 * (8 * 4 * fastMarkWords(BT) + 4)) * 4; \
		if(skipProb >= 1.f) { \
			if((err = allowDynamicSmem((const void*)k_marking_fast<BT, true>, smem)) != cudaSuccess) return err; \
			k_marking_fast<BT, true><<<bands, 32, smem, st>>>(timeouts, pre, prePitch, g, skipProb, seed, actBits, asBits, Wp, progress); \
		} else { \
			if((err = allowDynamicSmem((const void*)k_marking_fast<BT, false>, smem)) != cudaSuccess) return err; \
			k_marking_fast<BT, false><<<bands, 32, smem, st>>>(timeouts, pre, prePitch, g, skipProb, seed, actBits, asBits, Wp, progress); \
		} \
	} while(0)
	if(g.b == 6 && !g_forceGenericMarking) LAUNCH_FAST(6);
	else if(g.b == 10 && !g_forceGenericMarking) LAUNCH_FAST(10);
	else if(g.b == 14 && !g_forceGenericMarking) LAUNCH_FAST(14);
	else
	{
		const size_t smem = ((size_t)(32 + g.b) * Wp + (size_t)2 * 32 * (32 * g.Pw + 1)) * 4;
		if((err = allowDynamicSmem((const void*)k_marking_generic<0>, smem)) != cudaSuccess) return err;
		k_marking_generic<0><<<bands, 32, smem, st>>>(timeouts, pre, prePitch, g, skipProb, seed, actBits, asBits, Wp, progress);
	}
#undef LAUNCH_FAST
	return cudaGetLastError();
}

} // namespace bcdk

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
k_marking_fast(const uint32_t* __restrict__ pre2, int prePitch, LevelGeom g, float skipProb, uint32_t seed,
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
					while(*prevProgress < needCols) { __nanosleep(32); }
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
// AS bits never go through shared memory. The 16-column windows of the six rows above travel DOWN the warp as a
// three-word packet, one shuffle per word and step:
//     P_l(t)   = fields of rows l-1 .. l-6 at columns [C-6, C+9],  C = 4 (t - 3 l)          (what lane l consumes at step t)
//     P_l+1(t+1) = { own row at [C-14, C+1] (bits 14..29 of the lane's 32-column history), rows l-1 .. l-5 of P_l(t-2) }
// so a step costs 3 SHFL + 3 PRMT instead of 12 LDS + 6 SHF + 3 PRMT + STS + a warp barrier, and the compute warp does
// nothing else: a second warp streams the per-pixel input (`pre2`, cp.async into a double-buffered tile, one named
// barrier per 8 steps), a third one feeds lane 0 — from the AS bit-rows above the launch range (band 0: zeros, or the
// upper row tile's hand-off) or from the previous band's hand-off ring. Bands (one CTA each, all co-resident) are
// chained through that ring in global memory: the last six lanes of a band publish the four AS bits of every block
// they finish as ONE 32-bit word carrying its own valid flag (bit 31), so no fence and no progress counter is needed;
// the next band's feeder polls the words it needs (volatile loads) eight steps at a time.
// Every wait is bounded: a feeder that times out raises `dead` and the band finishes with whatever it has (the parity
// tests then fail; the GPU does not hang).
// ---------------------------------------------------------------------------------------------------------------
constexpr int SYS_CHS = 8;                          // steps per tile chunk (32 pixels per row)
constexpr int SYS_RS = SYS_CHS * 4 * 4 + 4;         // tile row stride in words (== 4 mod 32: LDS.128 conflict-free)
constexpr int SYS_PR = 64;                          // packet ring entries (steps)
constexpr int SYS_SPIN_LIMIT = 1 << 22;

__device__ __forceinline__ void namedBarrier64() { asm volatile("bar.sync 1, 64;" ::: "memory"); }
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

template <bool ALWAYS>
__global__ void __launch_bounds__(96)
k_marking_sys6(const uint32_t* __restrict__ pre2, int prePitch, LevelGeom g, float skipProb, uint32_t seed,
		uint32_t* __restrict__ actBits, uint32_t* asBits, int Wp, uint32_t* ring, int ringPitch,
		unsigned int* timeouts)
{
	constexpr int b = 6, PX = 4, SKS = 3, SKC = PX * SKS, PW2 = 4;
	extern __shared__ __align__(16) uint32_t smem[];
	uint32_t* tileS = smem;                                              // [2][32][SYS_RS]
	uint4* pktS = reinterpret_cast<uint4*>(smem + 2 * 32 * SYS_RS);      // [SYS_PR]
	volatile int* flags = reinterpret_cast<volatile int*>(pktS + SYS_PR); // [0] packets ready, [1] steps consumed, [2] dead
	const int W = g.W, H = g.H, w = g.w;
	const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
	const int band = blockIdx.x;
	const int r0 = g.rowBeg + band * 32;
	const int lastLane = min(31, g.rowEnd - 1 - r0);
	const int nBlocks = (W + PX - 1) / PX;
	const int tEnd = nBlocks + SKS * 31;
	const int nChunks = (tEnd + SYS_CHS - 1) / SYS_CHS;
	if(threadIdx.x < 3)
		flags[threadIdx.x] = 0;
	__syncthreads();

	if(role == 1)
	{
		// ---- tile feeder: chunk q of row l covers pixels [32 q - 12 l, +32); one 16-byte unit (= one pixel) per lane
		for(int q = 0; q < nChunks; ++q)
		{
			uint32_t* stage = tileS + (size_t)(q & 1) * 32 * SYS_RS;
			for(int l = 0; l < 32; ++l)
			{
				const int rr = r0 + l;
				const int px = SYS_CHS * PX * q - SKC * l + lane;
				const bool inside = rr < g.rowEnd && px >= 0 && px < prePitch;
				const uint32_t* src = inside ? pre2 + ((size_t)rr * prePitch + px) * PW2 : pre2;
				__pipeline_memcpy_async(stage + (size_t)l * SYS_RS + lane * PW2, src, 16, inside ? 0 : 16);
			}
			__pipeline_commit();
			__pipeline_wait_prior(0);
			__threadfence_block();
			namedBarrier64(); // chunk q has landed; the compute warp arrives here after finishing chunk q - 1
		}
		return;
	}
	if(role == 2)
	{
		// ---- lane-0 feeder: packets P_0(t), eight steps per batch
		const uint32_t* prevRing = ring + (size_t)(band > 0 ? band - 1 : 0) * 6 * ringPitch;
		for(int T = 0; T < tEnd; T += 8)
		{
			int spins = 0;
			while(flags[1] < T + 8 - SYS_PR && !flags[2])
				if(++spins > SYS_SPIN_LIMIT) { flags[2] = 1; atomicAdd(timeouts, 1u); }
			const int t = T + lane;
			if(lane < 8 && t < tEnd)
			{
				uint32_t f[6] = {0u, 0u, 0u, 0u, 0u, 0u};
				if(t < nBlocks + 2)
				{
					if(band == 0)
					{
						const int posLo = PX * t - b + MARK_PAD;
						const int wi = posLo >> 5, sh = posLo & 31;
#pragma unroll
						for(int k = 1; k <= 6; ++k)
						{
							const int rr = r0 - k;
							if(rr >= 0)
								f[k - 1] = __funnelshift_r(asBits[(size_t)rr * Wp + wi], asBits[(size_t)rr * Wp + wi + 1], sh) & 0xffffu;
						}
					}
					else
					{
						// blocks t-2 .. t+2 cover columns 4t-8 .. 4t+11; the field is bits 2..17 of the 20-bit concatenation
						uint32_t v[6][5];
						bool all = false;
						int sp = 0;
						while(!all)
						{
							all = true;
#pragma unroll
							for(int k = 0; k < 6; ++k)
#pragma unroll
								for(int j = 0; j < 5; ++j)
								{
									const int blk = t - 2 + j;
									uint32_t x = 0x80000000u;
									if(blk >= 0 && blk < nBlocks)
										x = ldVolatileGlobal(prevRing + (size_t)k * ringPitch + blk);
									v[k][j] = x;
									all = all && (x & 0x80000000u) != 0u;
								}
							if(!all)
							{
								if(flags[2]) break;
								if(++sp > SYS_SPIN_LIMIT / 64) { flags[2] = 1; atomicAdd(timeouts, 1u); break; }
								__nanosleep(40);
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
				pktS[t & (SYS_PR - 1)] = make_uint4(f[0] | (f[1] << 16), f[2] | (f[3] << 16), f[4] | (f[5] << 16), 0u);
			}
			__syncwarp();
			__threadfence_block();
			if(lane == 0)
				flags[0] = T + 8;
		}
		return;
	}

	// ---- compute warp
	const int r = r0 + lane;
	const bool rowValid = r < g.rowEnd;
	const bool rowCentre = rowValid && r >= w && r <= H - 1 - w;
	const uint32_t thr = (uint32_t)(fminf(skipProb, 0.99999994f) * 4294967296.f);
	const size_t rowBase = (size_t)(rowValid ? r : 0) * Wp;
	const int ringRow = lastLane - lane; // 0..5 for the six hand-off rows of the band
	uint32_t* myRing = ring + ((size_t)band * 6 + (ringRow >= 0 && ringRow < 6 ? ringRow : 0)) * ringPitch;
	const bool emits = rowValid && ringRow >= 0 && ringRow < 6;
	uint32_t hh = 0u;                 // AS bits of my row: bit 31 = the newest column
	uint32_t o0 = 0u, o1 = 0u, o2 = 0u; // packet for the lane below
	uint32_t d1a = 0u, d1b = 0u, d1c = 0u, d2a = 0u, d2b = 0u, d2c = 0u; // my packets of the last two steps
	uint32_t curAS = 0u, curAct = 0u;

	for(int t = 0; t < tEnd; ++t)
	{
		const int s = t % SYS_CHS, q = t / SYS_CHS;
		if(s == 0)
		{
			namedBarrier64();
			if(lane == 0)
				flags[1] = t;
			int spins = 0;
			while(flags[0] < t + 8 && !flags[2])
				if(++spins > SYS_SPIN_LIMIT) { flags[2] = 1; atomicAdd(timeouts, 1u); }
			__threadfence_block();
		}
		const int C = PX * (t - SKS * lane);
		const bool valid = rowValid && C >= 0 && C < W;
		const uint4* src = reinterpret_cast<const uint4*>(tileS + (size_t)(q & 1) * 32 * SYS_RS + (size_t)lane * SYS_RS + s * PX * PW2);
		uint4 pw[PX];
#pragma unroll
		for(int i = 0; i < PX; ++i)
			pw[i] = src[i];
		const uint4 first = pktS[t & (SYS_PR - 1)];
		uint32_t i0 = __shfl_up_sync(0xffffffffu, o0, 1);
		uint32_t i1 = __shfl_up_sync(0xffffffffu, o1, 1);
		uint32_t i2 = __shfl_up_sync(0xffffffffu, o2, 1);
		if(lane == 0) { i0 = first.x; i1 = first.y; i2 = first.z; }

		uint32_t asNibble = 0u, actNibble = 0u;
#pragma unroll
		for(int i = 0; i < PX; ++i)
		{
			const uint32_t above = (i0 & pw[i].x) | (i1 & pw[i].y) | (i2 & pw[i].z);
			const uint32_t hit = above | ((pw[i].w << 26) & hh); // same row: bits 0..5 of word 3 = columns c-6..c-1
			const int c = C + i;
			const bool centre = rowCentre && valid && c >= w && c <= W - 1 - w;
			const bool skipIfMarked = ALWAYS || hash32((uint32_t)r, (uint32_t)c, seed) <= thr;
			const bool act = centre && !(hit != 0u && skipIfMarked);
			const bool as = act && ((pw[i].w >> 16) & 1u) != 0u;
			hh = (hh >> 1) | (as ? 0x80000000u : 0u);
			asNibble |= (as ? 1u : 0u) << i;
			actNibble |= (act ? 1u : 0u) << i;
		}
		// packet for the lane below (consumed at its step t + 1, block C - 8): my row at [C-14, C+1] + my packet of step t - 2
		o0 = __byte_perm(hh >> 14, d2a, 0x5410);
		o1 = __byte_perm(d2a, d2b, 0x5432);
		o2 = __byte_perm(d2b, d2c, 0x5432);
		d2a = d1a; d2b = d1b; d2c = d1c;
		d1a = i0; d1b = i1; d1c = i2;
		if(valid)
		{
			const int pos = C + MARK_PAD; // multiple of 4: a block never straddles a word
			curAS |= asNibble << (pos & 31);
			curAct |= actNibble << (pos & 31);
			if(emits)
				stVolatileGlobal(myRing + (C >> 2), 0x80000000u | asNibble);
			if((pos & 31) == 28 || C + PX >= W)
			{
				actBits[rowBase + (pos >> 5)] = curAct;
				asBits[rowBase + (pos >> 5)] = curAS;
				curAS = 0u;
				curAct = 0u;
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
int g_markingKernel = 1;            // b = 6: 1 = systolic kernel (default), 0 = shared-memory wavefront (BCD_B200_MARKING=v4)

bool markingUsesFastLayout(int b) { return (b == 6 || b == 10 || b == 14) && !g_forceGenericMarking; }

cudaError_t launchMarking(const uint32_t* pre, int prePitch, const LevelGeom& g, float skipProb, uint32_t seed,
		uint32_t* actBits, uint32_t* asBits, int Wp, unsigned int* progress, cudaStream_t st, int* launches,
		unsigned int* timeouts)
{
	const int rows = g.rowEnd - g.rowBeg;
	if(rows <= 0)
		return cudaSuccess;
	const int bands = divUp(rows, 32);
	cudaError_t err;
	if(g.b == 6 && !g_forceGenericMarking && g_markingKernel == 1 && bands <= 4 * 148)
	{
		// scratch layout (markingScratchWords): [0] time-out counter, [4 ..) hand-off ring [bands][6][ringPitch]
		const int ringPitch = markingRingPitch(g.W);
		if((err = cudaMemsetAsync(progress, 0, sizeof(uint32_t) * (4 + (size_t)bands * 6 * ringPitch), st)) != cudaSuccess) return err;
		*launches += 1;
		const size_t smem = ((size_t)2 * 32 * SYS_RS + 4 * SYS_PR + 4) * 4;
		if(skipProb >= 1.f)
		{
			if((err = cudaFuncSetAttribute(k_marking_sys6<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
			k_marking_sys6<true><<<bands, 96, smem, st>>>(pre, prePitch, g, skipProb, seed, actBits, asBits, Wp, progress + 4, ringPitch, timeouts ? timeouts : progress);
		}
		else
		{
			if((err = cudaFuncSetAttribute(k_marking_sys6<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
			k_marking_sys6<false><<<bands, 96, smem, st>>>(pre, prePitch, g, skipProb, seed, actBits, asBits, Wp, progress + 4, ringPitch, timeouts ? timeouts : progress);
		}
		return cudaGetLastError();
	}
	if((err = cudaMemsetAsync(progress, 0, sizeof(unsigned int) * bands, st)) != cudaSuccess) return err;
	*launches += 1;
	const size_t asWords = ((size_t)(32 + g.b) * Wp + 3) & ~(size_t)3;
#define LAUNCH_FAST(BT) \
	do { \
		const size_t smem = (asWords + (size_t)2 * 32 * (8 * 4 * fastMarkWords(BT) + 4)) * 4; \
		if(skipProb >= 1.f) { \
			if((err = cudaFuncSetAttribute(k_marking_fast<BT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err; \
			k_marking_fast<BT, true><<<bands, 32, smem, st>>>(pre, prePitch, g, skipProb, seed, actBits, asBits, Wp, progress); \
		} else { \
			if((err = cudaFuncSetAttribute(k_marking_fast<BT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err; \
			k_marking_fast<BT, false><<<bands, 32, smem, st>>>(pre, prePitch, g, skipProb, seed, actBits, asBits, Wp, progress); \
		} \
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

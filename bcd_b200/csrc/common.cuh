// Shared device/host helpers for the bcd_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace bcdk
{

// Geometry of one pyramid level as seen by one GPU. (Single GPU: the whole image; row tiles: a horizontal band
// plus its halos — the kernels only ever see a W x H image and a range of owned patch-centre rows.)
struct LevelGeom
{
	int W, H;        // image size at this level
	int B;           // histogram depth (3 * nbOfBins)
	int w, b;        // patch radius, search radius
	int side;        // 2b+1
	int Ns;          // side*side   (search window size)
	int centre;      // (Ns-1)/2    (window index of the centre)
	int J;           // (Ns+1)/2    (number of offsets delta >= 0 in raster order, j=0 is the centre)
	int Jw;          // ceil(J/32)  words of the half mask M+
	int Nw;          // ceil(Ns/32) words of the full mask S
	int Pw;          // ceil((centre+1)/32) words of the marking input (bits before the centre + strong flag)
	int np;          // (2w+1)^2    pixels per patch
	int d;           // 3*np        colour patch dimension
	int rowBeg, rowEnd; // owned patch-centre rows [rowBeg, rowEnd) (subset of the interior rows)
	int yLo, yHi;       // rows on which the selection stage (K1) runs; [0, H) unless this is a row tile
};

inline LevelGeom makeGeom(int W, int H, int B, int w, int b)
{
	LevelGeom g;
	g.W = W; g.H = H; g.B = B; g.w = w; g.b = b;
	g.side = 2 * b + 1;
	g.Ns = g.side * g.side;
	g.centre = (g.Ns - 1) / 2;
	g.J = (g.Ns + 1) / 2;
	g.Jw = (g.J + 31) / 32;
	g.Nw = (g.Ns + 31) / 32;
	g.Pw = (g.centre + 1 + 31) / 32;
	g.np = (2 * w + 1) * (2 * w + 1);
	g.d = 3 * g.np;
	g.rowBeg = w;
	g.rowEnd = H - w;
	g.yLo = 0;
	g.yHi = H;
	return g;
}

// offset j (0..J-1) -> (dy, dx), raster order after the centre: dy in [0,b], dx in [-b,b]
__host__ __device__ inline void offsetOf(int j, int b, int& dy, int& dx)
{
	const int side = 2 * b + 1;
	const int idx = b * side + b + j;
	dy = idx / side - b;
	dx = idx % side - b;
}

__host__ __device__ inline int divUp(int a, int b) { return (a + b - 1) / b; }

// Bit-rows (AS flags and active flags of the marking stage): row y starts at word y*Wp, column c is bit c + MARK_PAD.
constexpr int MARK_PAD = 32;
// (odd word count: lanes of the marking kernel read different rows at nearby word indices -> distinct banks)
__host__ __device__ inline int markingRowWords(int W) { return ((W + 2 * MARK_PAD + 31) / 32 + 2) | 1; }
// words per pixel of the fast marking layout `pre2` (b = 6: 4, b = 10: 12, b = 14: 16); see k_marking_fast
__host__ __device__ inline int fastMarkWords(int b)
{
	const int fw = 2 * b + 4, fpw = fw <= 16 ? 2 : 1, nf = (b + fpw - 1) / fpw;
	return (nf + 1 + 3) & ~3;
}
// scratch of the marking stage (uint32 words): progress counters of the shared-memory wavefront kernel, or the time-out
// counter + the band hand-off ring [bands][6 rows][ringPitch blocks of 4 columns] of the systolic kernel
// (a ring row has 96 slack words in front and 112 behind: the producers store unconditionally for every step)
__host__ __device__ inline int markingRingPitch(int W) { return ((W + 3) / 4 + 96 + 112 + 3) & ~3; }
inline size_t markingScratchWords(int W, int H)
{
	const size_t bands = (size_t)((H > 0 ? H : 1) + 31) / 32 + 2;
	return 4 + (bands * 6 + 1) * (size_t)markingRingPitch(W) + bands; // + 1: the spare row of the systolic kernel
}
// Systolic marking kernel (b = 6): number of steps of a band (lane l handles block t - 3 l at step t; multiple of 8) and
// its band-major `pre2` layout — [band][step t][pixel i of the block][lane] x 16 bytes, so that the 32 lanes (rows) of a
// band read 512 contiguous bytes per load instead of 32 different rows
__host__ __device__ inline int sysTEnd(int W) { return ((W + 3) / 4 + 3 * 31 + 8 + 7) & ~7; }
__host__ __device__ inline size_t sysPreIndex(int band, int t, int i, int lane, int tEnd)
{
	return (((size_t)band * tEnd + t) * 4 + i) * 32 + lane; // in 16-byte units
}
inline size_t sysPreBytes(int W, int rows) { return (size_t)((rows > 0 ? rows : 1) + 31) / 32 * sysTEnd(W) * 128 * 16; }
// `pre` rows are pitched to a multiple of 4 pixels so that 4-pixel groups are 16-byte aligned (cp.async 16)
__host__ __device__ inline int prePitchOf(int W) { return (W + 3) & ~3; }
// slots of a filter chunk [first, first + cap): the host passes the capacity, the device the list length when the host
// never read it back (row-tile peer mode: no host hop between marking and filter)
__device__ __forceinline__ int chunkCount(int cap, const int* countDev, int first)
{
	if(!countDev)
		return cap;
	const int left = *countDev - first;
	return left < cap ? (left > 0 ? left : 0) : cap;
}
__device__ __forceinline__ bool testRowBit(const uint32_t* bits, int Wp, int y, int x)
{
	const int pos = x + MARK_PAD;
	return (bits[(size_t)y * Wp + (pos >> 5)] >> (pos & 31)) & 1u;
}

} // namespace bcdk

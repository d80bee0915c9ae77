// K3e, sub-group variant for d <= 32 (patch radius 1, d = 27): 32/G matrices per warp, G lanes per matrix.
//
// Same algorithm and reference mapping as k_eig_chain in k_filter.cu (clampNegativeEigenValues / inverseSymmetricMatrix,
// /root/reference/src/core/DenoisingUnit.cpp:578-630, with Eigen's SelfAdjointEigenSolver: scaling, Householder
// tridiagonalisation, Q accumulation, implicit symmetric QR with Wilkinson shift, and the Cholesky fast path for the
// clamped inverse). Why a second layout: in the one-warp-per-matrix kernel every lane redundantly tracks the scalar
// bulge-chase recurrence (the hot loop, ~60 % of the instructions) while only one Q row per lane is rotated. Here the
// warp carries 2 (G=16) or 4 (G=8) independent recurrences at once — the scalar work per matrix drops by that factor —
// and each lane rotates ceil(n/G) rows. Matrices in one warp run in lock-step: loops with data-dependent trip counts
// (QR sweeps, deflation) iterate to the warp maximum with per-matrix predicates, so every __syncwarp() is reached by
// all lanes.
//
// Shared memory holds ONE n x n buffer per matrix (3.9 KB at d = 27): V f(L) V^T is written straight to global memory
// (the chain's own scratch, L2-resident) and read back with the noise blocks added; the Cholesky factor, the inverse
// of the factor (strict upper triangle) and the factorised input share that buffer, and A^-1 goes straight to the
// output. A matrix whose Cholesky path is rejected is re-read from global memory for the eigen path (rare).

#include "common.cuh"
#include "kernels.h"
#include <algorithm>
#include <cfloat>

namespace bcdk
{

namespace
{

constexpr unsigned FULL = 0xffffffffu;

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

template <int G>
__device__ __forceinline__ float subSum(float v)
{
#pragma unroll
	for(int o = G / 2; o > 0; o >>= 1)
		v += __shfl_xor_sync(FULL, v, o);
	return v;
}
template <int G>
__device__ __forceinline__ float subMax(float v)
{
#pragma unroll
	for(int o = G / 2; o > 0; o >>= 1)
		v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
	return v;
}

// element (a, b) of the symmetric 3x3 block stored as xx,yy,zz,yz,xz,xy (reference CovarianceMatrix.h:18-27)
__device__ __forceinline__ int symIndex(int a, int b)
{
	if(a == b) return a;
	const int s = a + b;
	return s == 3 ? 3 : (s == 2 ? 4 : 5);
}

// c = p/r, s = -q/r (JacobiRotation::makeGivens, Jacobi.h:216-252), see k_filter.cu
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

struct SubSmem
{
	float *M, *diag, *sub, *dW, *sW, *hco, *pv, *rowAbs;
};

#define MAT(r, c) M[(r) + (c) * n]

// outG (global, n x n) <- V f(L) V^T of the matrix whose lower triangle is in M (mode 0: f = max(0, l); mode 1:
// f = 1/max(eps, l)). `commit` (per matrix): when false the matrix only keeps the warp in step and nothing is written.
// NC > 0: the matrix order is the compile-time constant NC (27 for patch radius 1: index arithmetic folds, the inner
// loops unroll); NC = 0: run-time order nRt.
template <int G, int NC>
__device__ void eigFunctionSub(const SubSmem& S, int nRt, int mode, float eps, bool commit, float* __restrict__ outG)
{
	constexpr int RMAX = ((NC > 0 ? NC : 32) + G - 1) / G;
	const int n = NC > 0 ? NC : nRt;
	float* M = S.M;
	const int lid = threadIdx.x & (G - 1), sgrp = (threadIdx.x & 31) / G;
	// ---- scale the lower triangle into [-1,1], clear the strict upper triangle
	float lmax = 0.f;
	for(int e = lid; e < n * n; e += G)
	{
		const int r = e % n, c = e / n;
		if(r < c) M[e] = 0.f;
		else lmax = fmaxf(lmax, fabsf(M[e]));
	}
	float scale = subMax<G>(lmax);
	if(scale == 0.f) scale = 1.f;
	__syncwarp();
	for(int e = lid; e < n * n; e += G)
		if(e % n >= e / n) M[e] = __fdiv_rn(M[e], scale);
	__syncwarp();

	// ---- Householder tridiagonalisation (Tridiagonalization.h:346-378); uniform trip counts
	for(int i = 0; i < n - 1; ++i)
	{
		const int rs = n - i - 1;
		float* v = &MAT(i + 1, i);
		float part = 0.f;
		for(int k = 1 + lid; k < rs; k += G)
			part += v[k] * v[k];
		const float tailSq = subSum<G>(part);
		const float c0 = v[0];
		const bool degenerate = tailSq <= FLT_MIN;
		float beta = __fsqrt_rn(c0 * c0 + tailSq);
		if(c0 >= 0.f) beta = -beta;
		const float den = c0 - beta;
		float h = __fdiv_rn(beta - c0, beta);
		if(degenerate) { h = 0.f; beta = c0; }
		__syncwarp();
		for(int k = 1 + lid; k < rs; k += G)
			v[k] = degenerate ? 0.f : __fdiv_rn(v[k], den);
		__syncwarp();
		for(int r = lid; r < rs; r += G) // p = A_sym (h v), v[0] == 1 implicitly
		{
			float a = 0.f;
			for(int c = 0; c < rs; ++c)
			{
				const float m = r >= c ? MAT(i + 1 + r, i + 1 + c) : MAT(i + 1 + c, i + 1 + r);
				a += m * (h * (c == 0 ? 1.f : v[c]));
			}
			S.pv[r] = a;
		}
		__syncwarp();
		part = 0.f;
		for(int k = lid; k < rs; k += G)
			part += S.pv[k] * (k == 0 ? 1.f : v[k]);
		const float f = h * -0.5f * subSum<G>(part);
		__syncwarp();
		for(int k = lid; k < rs; k += G)
			S.pv[k] += f * (k == 0 ? 1.f : v[k]);
		__syncwarp();
		for(int c = lid; c < rs; c += G)
		{
			const float pc = S.pv[c], vc = c == 0 ? 1.f : v[c];
			for(int r = c; r < rs; ++r)
				MAT(i + 1 + r, i + 1 + c) -= (r == 0 ? 1.f : v[r]) * pc + S.pv[r] * vc;
		}
		__syncwarp();
		if(lid == 0)
		{
			v[0] = beta;
			S.hco[i] = h;
		}
		__syncwarp();
	}
	for(int k = lid; k < n; k += G)
	{
		S.diag[k] = MAT(k, k);
		if(k < n - 1) S.sub[k] = MAT(k + 1, k);
	}
	__syncwarp();
	// ---- Q = H_0 ... H_{n-2} in place (HouseholderSequence.h:242-267)
	for(int k = lid; k < n; k += G)
		MAT(k, k) = 1.f;
	__syncwarp();
	for(int k = n - 2; k >= 0; --k)
	{
		const int cs = n - k - 1;
		const float tau = S.hco[k];
		if(cs == 1)
		{
			if(lid == 0) MAT(k + 1, k + 1) *= 1.f - tau;
		}
		else if(tau != 0.f)
		{
			const float* ess = &MAT(k + 2, k);
			for(int c = lid; c < cs; c += G)
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
		__syncwarp();
		for(int r = k + 1 + lid; r < n; r += G)
			MAT(r, k) = 0.f;
		__syncwarp();
	}
	// ---- implicit symmetric QR (SelfAdjointEigenSolver.h:486-530, :808-864), lock-step over the matrices of the warp
	{
		int end = n - 1, start = 0, iter = 0;
		bool done = false;
		const float precision = 2.f * FLT_EPSILON;
		const uint32_t diagA = (uint32_t)__cvta_generic_to_shared(S.diag), subA = (uint32_t)__cvta_generic_to_shared(S.sub);
		const uint32_t dWA = (uint32_t)__cvta_generic_to_shared(S.dW), sWA = (uint32_t)__cvta_generic_to_shared(S.sW);
		const uint32_t mA = (uint32_t)__cvta_generic_to_shared(M);
		const uint32_t colStep = 4u * (uint32_t)n;
		while(true)
		{
			// deflation test (SelfAdjointEigenSolver.h:497-503) and, in the same pass, the bit set nz = {i < end :
			// sub[i] != 0} of the matrix (one ballot per round of G entries) — the scans for `end` and `start`
			// (:506-520) then are two bit searches instead of two serial walks through shared memory
			uint32_t nz = 0u;
#pragma unroll
			for(int j = 0; j < RMAX; ++j)
			{
				const int i = lid + j * G;
				bool keep = false;
				if(!done && i < end)
				{
					const float a = fabsf(S.sub[i]);
					if(a <= (fabsf(S.diag[i]) + fabsf(S.diag[i + 1])) * precision || a <= FLT_MIN)
						S.sub[i] = 0.f;
					else
						keep = true; // a > FLT_MIN (or NaN): stays non-zero
				}
				const uint32_t bal = __ballot_sync(FULL, keep);
				nz |= ((bal >> (sgrp * G)) & ((1u << G) - 1u)) << (j * G);
			}
			__syncwarp();
			if(!done)
			{
				if(nz == 0u) { end = 0; done = true; } // every sub[i], i < end, is zero
				else
				{
					end = 32 - __clz(nz); // highest non-zero sub-diagonal entry + 1
					if(++iter > 30 * n) done = true;
					else
					{
						const uint32_t z = ~nz & ((1u << (end - 1)) - 1u); // zero entries below end - 1
						start = z ? 32 - __clz(z) : 0;
					}
				}
			}
			if(__all_sync(FULL, done))
				break;
			const int len = done ? 0 : end - start;
			const int maxLen = __reduce_max_sync(FULL, len);
			float x = 0.f, z = 0.f, dk = 0.f, sk = 0.f, skm1 = 0.f;
			float qk[RMAX];
			uint32_t qA = mA + 4u * (uint32_t)lid + colStep * (uint32_t)start; // &Q(lid, k); rows lid + j*G
			if(!done)
			{
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
				dk = S.diag[start];
				sk = S.sub[start];
				x = dk - mu;
				z = sk;
#pragma unroll
				for(int j = 0; j < RMAX; ++j)
					qk[j] = (lid + j * G < n) ? ldsF(qA + 4u * (uint32_t)(j * G)) : 0.f;
			}
			for(int kk = 0; kk < maxLen; ++kk)
			{
				const int k = start + kk;
				if(!done && k < end)
				{
					float c, s;
					makeGivens(x, z, c, s);
					const float dk1 = ldsF(diagA + 4u * (uint32_t)(k + 1));
					const float sdk = fmaf(s, dk, c * sk);
					const float dkp1 = fmaf(s, sk, c * dk1);
					const float ndk = fmaf(c, fmaf(c, dk, -s * sk), -s * fmaf(c, sk, -s * dk1));
					const float ndk1 = fmaf(s, sdk, c * dkp1);
					const float nsk = fmaf(c, sdk, -s * dkp1);
					if(lid == 0)
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
#pragma unroll
					for(int j = 0; j < RMAX; ++j)
						if(lid + j * G < n)
						{
							const float y = ldsF(qA + colStep + 4u * (uint32_t)(j * G));
							stsF(qA + 4u * (uint32_t)(j * G), fmaf(c, qk[j], -s * y));
							qk[j] = fmaf(s, qk[j], c * y);
						}
					qA += colStep;
					dk = ndk1;
					skm1 = nsk;
					sk = nsk1;
				}
			}
			if(!done)
			{
#pragma unroll
				for(int j = 0; j < RMAX; ++j)
					if(lid + j * G < n)
						stsF(qA + 4u * (uint32_t)(j * G), qk[j]);
				if(lid == 0)
				{
					S.dW[end] = dk;
					S.sW[end - 1] = skm1;
				}
			}
			__syncwarp();
			if(!done)
				for(int k = start + lid; k <= end; k += G)
				{
					S.diag[k] = S.dW[k];
					if(k < end) S.sub[k] = S.sW[k];
				}
			__syncwarp();
		}
	}
	// ---- O = V f(L) V^T
	for(int k = lid; k < n; k += G)
	{
		const float l = S.diag[k] * scale;
		S.pv[k] = mode == 0 ? fmaxf(0.f, l) : __fdiv_rn(1.f, fmaxf(eps, l));
	}
	__syncwarp();
	const int L = n * (n + 1) / 2;
	if(commit)
		for(int e = lid; e < L; e += G)
		{
			int c = 0, r = e;
			while(r >= n - c) { r -= n - c; ++c; }
			r += c;
			float a = 0.f;
			for(int k = 0; k < n; ++k)
				a += MAT(r, k) * (S.pv[k] * MAT(c, k));
			outG[r + c * n] = a;
			outG[c + r * n] = a;
		}
	__syncwarp();
}

// Cholesky fast path of the clamped inverse (see choleskyInverse in k_filter.cu), in place: L overwrites the lower
// triangle of M, X = L^-1 goes to the strict upper triangle (X(i,j), i > j, at M(j,i); X(j,j) in pv[j]) and
// A^-1 = X^T X is written to outG. Returns true per matrix when the factorisation exists and 1/||A^-1||_inf > 2 eps;
// never breaks out early, so all matrices of the warp stay in step (a failed factorisation just computes garbage that
// the caller replaces).
template <int G, int NC>
__device__ bool choleskyInverseSub(const SubSmem& S, int nRt, float eps, float* __restrict__ outG, bool valid)
{
	const int n = NC > 0 ? NC : nRt;
	float* M = S.M;
	const int lid = threadIdx.x & (G - 1);
	bool ok = true;
	for(int k = 0; k < n; ++k)
	{
		const float piv = MAT(k, k);
		if(!(piv > 0.f) || !(piv < 3e38f)) ok = false;
		const float pv = ok ? piv : 1.f;
		float inv;
		asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(pv));
		const float lkkInv = inv * fmaf(-0.5f * pv, inv * inv, 1.5f);
		__syncwarp();
		for(int r = k + lid; r < n; r += G)
			MAT(r, k) = r == k ? pv * lkkInv : MAT(r, k) * lkkInv;
		__syncwarp();
		for(int c = k + 1 + lid; c < n; c += G)
		{
			const float lck = MAT(c, k);
			for(int r = c; r < n; ++r)
				MAT(r, c) = fmaf(-MAT(r, k), lck, MAT(r, c));
		}
		__syncwarp();
	}
	for(int j = lid; j < n; j += G)
	{
		S.pv[j] = __fdiv_rn(1.f, MAT(j, j));
		S.rowAbs[j] = 0.f;
	}
	__syncwarp();
	for(int j = lid; j < n; j += G) // forward substitution, one column of X per lane
	{
		const float xjj = S.pv[j];
		for(int i = j + 1; i < n; ++i)
		{
			float a = MAT(i, j) * xjj;
			for(int k = j + 1; k < i; ++k)
				a = fmaf(MAT(i, k), MAT(j, k), a);
			MAT(j, i) = -a * S.pv[i];
		}
	}
	__syncwarp();
	// A^-1(r,c) = sum_{k >= r} X(k,r) X(k,c)   (r >= c); the row sums of |A^-1| give its infinity norm
	const int L = n * (n + 1) / 2;
	for(int e = lid; e < L; e += G)
	{
		int c = 0, r = e;
		while(r >= n - c) { r -= n - c; ++c; }
		r += c;
		float a = S.pv[r] * (r == c ? S.pv[r] : MAT(c, r));
		for(int k = r + 1; k < n; ++k)
			a = fmaf(MAT(r, k), MAT(c, k), a);
		if(valid)
		{
			outG[r + c * n] = a;
			outG[c + r * n] = a;
		}
		atomicAdd(&S.rowAbs[r], fabsf(a));
		if(r != c) atomicAdd(&S.rowAbs[c], fabsf(a));
	}
	__syncwarp();
	float rowMax = 0.f;
	for(int r = lid; r < n; r += G)
		rowMax = fmaxf(rowMax, S.rowAbs[r]);
	const float norm = subMax<G>(rowMax);
	return ok && norm < 0.5f / eps && norm == norm;
}
#undef MAT

// A (in): the matrices; in a clamp launch A[slot] is overwritten with V max(0,L) V^T (its consumer, the apply kernel,
// rewrites A anyway). MinvG (out): the clamped inverses.
template <int G, int NC>
__global__ void __launch_bounds__(32, G == 16 ? 24 : 14)
k_eig_chain_sub(float* __restrict__ A, float* __restrict__ MinvG, const float* __restrict__ noiseG, int count,
		const int* __restrict__ countDev, int first, int nRt, int np, float eps, int doClamp)
{
	extern __shared__ __align__(16) unsigned char smemRaw[];
	const int n = NC > 0 ? NC : nRt;
	count = chunkCount(count, countDev, first);
	constexpr int NM = 32 / G; // matrices per warp
	const int sub = threadIdx.x / G, lid = threadIdx.x & (G - 1);
	const int perMat = n * n + 7 * n + np * 6;
	float* f = reinterpret_cast<float*>(smemRaw) + (size_t)sub * perMat;
	SubSmem S;
	S.M = f; f += n * n;
	S.diag = f; f += n;
	S.sub = f; f += n;
	S.dW = f; f += n;
	S.sW = f; f += n;
	S.hco = f; f += n;
	S.pv = f; f += n;
	S.rowAbs = f; f += n;
	float* noise = f;
	for(int base = blockIdx.x * NM; base < count; base += gridDim.x * NM)
	{
		const int slot = base + sub;
		const bool valid = slot < count;
		float* Aslot = A + (size_t)(valid ? slot : 0) * n * n;
		float* outSlot = MinvG + (size_t)(valid ? slot : 0) * n * n;
		__syncwarp();
		for(int e = lid; e < n * n; e += G)
			S.M[e] = valid ? Aslot[e] : ((e % n == e / n) ? 1.f : 0.f); // identity keeps idle lanes harmless
		for(int e = lid; e < np * 6; e += G)
			noise[e] = valid ? noiseG[(size_t)slot * np * 6 + e] : 0.f;
		__syncwarp();
		// M <- (global matrix) + Sigma_n on the 3x3 diagonal blocks (DenoisingUnit.cpp:538-556)
		auto reloadWithNoise = [&](bool addNoise)
		{
			for(int e = lid; e < n * n; e += G)
			{
				const int r = e % n, c = e / n;
				float v = valid ? Aslot[e] : (r == c ? 1.f : 0.f);
				if(addNoise && r / 3 == c / 3)
					v = __fadd_rn(v, noise[(r / 3) * 6 + symIndex(r % 3, c % 3)]);
				S.M[e] = v;
			}
			__syncwarp();
		};
		if(doClamp)
		{
			eigFunctionSub<G, NC>(S, n, 0, 0.f, valid, Aslot); // A[slot] <- V max(0,L) V^T
			reloadWithNoise(true);
		}
		const bool okChol = choleskyInverseSub<G, NC>(S, n, eps, outSlot, valid);
		if(!__all_sync(FULL, okChol)) // rare: the failed matrices of the warp take the reference's eigen path
		{
			reloadWithNoise(doClamp != 0);
			eigFunctionSub<G, NC>(S, n, 1, eps, valid && !okChol, outSlot);
		}
	}
}

} // namespace

bool g_eigGenericOrder = false; // BCD_B200_EIG_GENERIC=1: the run-time-order kernel also for d = 27 (A/B hook)
int g_eigGroup = 16; // lanes per matrix for d <= 32 (BCD_B200_EIG_GROUP = 32 selects the one-warp-per-matrix kernel)

cudaError_t launchEigChainSub(float* A, float* Minv, const float* noise, int count, const int* countDev, int first,
		int n, int np, float eps, int doClamp, int group, int smCount, cudaStream_t st)
{
	const int nm = 32 / group;
	const size_t smem = (size_t)nm * (n * n + 7 * n + np * 6) * 4;
	// persistent warps (slot loop): two waves of the resident CTAs at most
	const int resident = (group == 16 ? 24 : 14) * (smCount > 0 ? smCount : currentSmCount());
	const int grid = std::min((count + nm - 1) / nm, 2 * resident);
	cudaError_t err;
#define LAUNCH_SUB(GV, NV) \
	do { \
		if((err = allowDynamicSmem((const void*)k_eig_chain_sub<GV, NV>, smem)) != cudaSuccess) return err; \
		k_eig_chain_sub<GV, NV><<<grid, 32, smem, st>>>(A, Minv, noise, count, countDev, first, n, np, eps, doClamp); \
	} while(0)
	if(group == 16)
	{
		if(n == 27 && !g_eigGenericOrder) LAUNCH_SUB(16, 27);
		else LAUNCH_SUB(16, 0);
	}
	else
	{
		if(n == 27 && !g_eigGenericOrder) LAUNCH_SUB(8, 27);
		else LAUNCH_SUB(8, 0);
	}
#undef LAUNCH_SUB
	return cudaGetLastError();
}

} // namespace bcdk

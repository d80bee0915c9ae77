// Micro-benchmark (sm_100a): issue rate of scalar vs packed FP32 (FFMA / FFMA2), FLO and MUFU.RCP, per SM per clock.
// Used to decide whether the pair-term kernel (K1a) can trade issue slots for packed f32x2 arithmetic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096, UNROLL = 8;

template <int MODE>
__global__ void k(float* out, float seed, long long* cycles)
{
	float a[UNROLL], b[UNROLL];
	unsigned long long p[UNROLL];
	unsigned int u[UNROLL];
	for(int i = 0; i < UNROLL; ++i)
	{
		a[i] = seed + i + threadIdx.x * 1e-3f;
		b[i] = seed * 0.5f + i;
		asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[i]), "f"(b[i]));
		u[i] = __float_as_uint(a[i]) | 1u;
	}
	const float c0 = seed * 1.0001f, c1 = seed * 0.9999f;
	unsigned long long pc;
	asm("mov.b64 %0, {%1,%2};" : "=l"(pc) : "f"(c0), "f"(c1));
	__syncthreads();
	const long long t0 = clock64();
	for(int it = 0; it < ITER; ++it)
	{
#pragma unroll
		for(int i = 0; i < UNROLL; ++i)
		{
			if(MODE == 0) a[i] = fmaf(a[i], c0, c1);
			if(MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pc));
			if(MODE == 2) asm volatile("bfind.u32 %0, %0;" : "+r"(u[i]));
			if(MODE == 3) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
			if(MODE == 4) { a[i] = fmaf(a[i], c0, c1); asm volatile("bfind.u32 %0, %0;" : "+r"(u[i])); } // co-issue?
			if(MODE == 5) { asm volatile("bfind.u32 %0, %0;" : "+r"(u[i])); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
		}
	}
	const long long t1 = clock64();
	float s = 0.f;
	for(int i = 0; i < UNROLL; ++i)
	{
		float x, y;
		asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i]));
		s += a[i] + b[i] + x + y + __uint_as_float(u[i]);
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if(threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instrPerSlot)
{
	const int threads = 1024, blocks = 148;
	float* out; long long* cyc;
	cudaMalloc(&out, threads * blocks * 4);
	cudaMalloc(&cyc, blocks * 8);
	k<MODE><<<blocks, threads>>>(out, 1.0f, cyc);
	k<MODE><<<blocks, threads>>>(out, 1.0f, cyc);
	cudaDeviceSynchronize();
	long long h[148];
	cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
	double avg = 0;
	for(int i = 0; i < blocks; ++i) avg += h[i];
	avg /= blocks;
	const double warpInstr = (double)ITER * UNROLL * instrPerSlot * (threads / 32);
	printf("%-28s %8.0f cycles  %.3f warp-instr/clk/SM  (%.1f lane-ops/clk/SM)\n", name, avg, warpInstr / avg, warpInstr * 32 / avg);
	cudaFree(out); cudaFree(cyc);
}

int main()
{
	run<0>("FFMA", 1);
	run<1>("FFMA2 (f32x2)", 1);
	run<2>("FLO (bfind)", 1);
	run<3>("MUFU.RCP", 1);
	run<4>("FFMA + FLO interleaved", 2);
	run<5>("FLO + MUFU.RCP interleaved", 2);
	return 0;
}

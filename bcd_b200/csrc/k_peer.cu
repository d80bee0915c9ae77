// Peer-memory exchange kernels of the row-tile path (multi-GPU; nothing in the reference to replace: it has
// no multi-GPU path). A strip travels as
//     sender:   k_push   — copies its rows straight into a slot of the RECEIVER's arena (stores over NVLink, posted),
//                          every thread fences system-wide, the last block to finish publishes the frame's epoch in the
//                          receiver's flag word;
//     receiver: k_recv   — one thread per block polls the local flag (bounded in time), then the block copies / adds the
//                          slot into its own image.
// No host thread, no NCCL call and no stream of the other GPU is involved: the only coupling is the flag word.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace bcdk
{

namespace
{

constexpr long long PEER_TIMEOUT_CYCLES = 1ll << 33; // ~4 s at 2 GHz: a lost neighbour fails the run, it never hangs the GPU

__device__ __forceinline__ uint32_t ldAcquireSys(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void stReleaseSys(uint32_t* p, uint32_t v)
{
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(256)
k_push(const T* __restrict__ src, T* __restrict__ dstPeer, size_t n, uint32_t* flagPeer, uint32_t epoch, unsigned int* ticket)
{
	for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		dstPeer[i] = src[i];
	__threadfence_system();
	__syncthreads();
	if(threadIdx.x == 0)
	{
		const unsigned int t = atomicAdd(ticket, 1u);
		if(t == gridDim.x - 1)
		{
			*ticket = 0u; // the next launch on this ticket is ordered behind this one by its stream
			__threadfence_system();
			stReleaseSys(flagPeer, epoch);
		}
	}
}

__device__ __forceinline__ float4 addv(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ uint32_t addv(uint32_t a, uint32_t b) { return a + b; }

template <typename T, bool ADD>
__global__ void __launch_bounds__(256)
k_recv(T* __restrict__ dst, const T* slot, size_t n, const uint32_t* flag, uint32_t epoch, unsigned int* timeouts)
{
	if(threadIdx.x == 0)
	{
		const long long t0 = clock64();
		while((int)(ldAcquireSys(flag) - epoch) < 0)
		{
			if(clock64() - t0 > PEER_TIMEOUT_CYCLES) { atomicAdd(timeouts, 1u); break; }
			__nanosleep(200);
		}
	}
	__syncthreads();
	for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
	{
		const T v = __ldcg(slot + i); // written by the peer: read at the coherence point (L2)
		dst[i] = ADD ? addv(dst[i], v) : v;
	}
}

int gridFor(size_t n) { return (int)std::min<size_t>((n + 255) / 256, (size_t)currentSmCount()); }
// a receiving kernel may sit on the device for milliseconds waiting for its neighbour: few blocks (grid-stride copy of a
// strip of some ten thousand elements), so that the waiting threads do not take the slots of the kernels that do work
int recvGridFor(size_t n) { return (int)std::max<size_t>(1, std::min<size_t>((n + 1023) / 1024, 16)); }

} // namespace

cudaError_t launchPeerPush(const void* src, void* dstPeer, size_t bytes, uint32_t* flagPeer, uint32_t epoch, unsigned int* ticket,
		cudaStream_t st, int* launches)
{
	if(bytes % 16 == 0 && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dstPeer % 16) == 0)
	{
		const size_t n = bytes / 16;
		k_push<float4><<<gridFor(n), 256, 0, st>>>(static_cast<const float4*>(src), static_cast<float4*>(dstPeer), n, flagPeer, epoch, ticket);
	}
	else
	{
		const size_t n = bytes / 4;
		k_push<uint32_t><<<gridFor(n), 256, 0, st>>>(static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dstPeer), n, flagPeer, epoch, ticket);
	}
	*launches += 1;
	return cudaGetLastError();
}

cudaError_t launchPeerRecv(void* dst, const void* slot, size_t bytes, bool addFloat4, const uint32_t* flag, uint32_t epoch,
		unsigned int* timeouts, cudaStream_t st, int* launches)
{
	if(addFloat4)
	{
		const size_t n = bytes / 16;
		k_recv<float4, true><<<recvGridFor(n), 256, 0, st>>>(static_cast<float4*>(dst), static_cast<const float4*>(slot), n, flag, epoch, timeouts);
	}
	else if(bytes % 16 == 0 && ((uintptr_t)dst % 16) == 0 && ((uintptr_t)slot % 16) == 0)
	{
		const size_t n = bytes / 16;
		k_recv<float4, false><<<recvGridFor(n), 256, 0, st>>>(static_cast<float4*>(dst), static_cast<const float4*>(slot), n, flag, epoch, timeouts);
	}
	else
	{
		const size_t n = bytes / 4;
		k_recv<uint32_t, false><<<recvGridFor(n), 256, 0, st>>>(static_cast<uint32_t*>(dst), static_cast<const uint32_t*>(slot), n, flag, epoch, timeouts);
	}
	*launches += 1;
	return cudaGetLastError();
}

// empty launches of every exchange kernel (tile_setup: loads them while no neighbour can be waiting); scratch = one word
cudaError_t launchPeerWarmup(unsigned int* scratch, cudaStream_t st)
{
	uint32_t* w = reinterpret_cast<uint32_t*>(scratch);
	k_push<float4><<<1, 256, 0, st>>>(nullptr, nullptr, 0, w + 1, 0u, scratch);
	k_push<uint32_t><<<1, 256, 0, st>>>(nullptr, nullptr, 0, w + 1, 0u, scratch);
	k_recv<float4, true><<<1, 256, 0, st>>>(nullptr, nullptr, 0, w + 1, 0u, scratch + 2);
	k_recv<float4, false><<<1, 256, 0, st>>>(nullptr, nullptr, 0, w + 1, 0u, scratch + 2);
	k_recv<uint32_t, false><<<1, 256, 0, st>>>(nullptr, nullptr, 0, w + 1, 0u, scratch + 2);
	return cudaGetLastError();
}

} // namespace bcdk

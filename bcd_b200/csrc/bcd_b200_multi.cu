// Multi-GPU side of the C ABI (include/bcd_b200.h): peer mode of the row-tile path and the group API.
//
// The reference has no multi-GPU path; what is split here is its single-image loop nest
//   Denoiser::denoise            /root/reference/src/core/Denoiser.cpp:84-212
//   MultiscaleDenoiser::denoise  /root/reference/src/core/MultiscaleDenoiser.cpp:31-136
// into horizontal bands, one per GPU. Design:
//   * one handle (bcd_b200_capi.cu) per GPU holds the band plus its halo rows and runs the ordinary stage kernels on it;
//   * the strips the bands owe each other — AS bits of the raster-order marking chain, accumulator halos, output rows
//     for the pyramid merge — are pushed by kernels of the SENDING GPU into an arena of the RECEIVING GPU over NVLink
//     and consumed behind an epoch flag (k_peer.cu); the systolic marking kernel chains neighbouring GPUs word by word;
//   * every pyramid level is a pipeline of its own (marking -> filter -> halo exchange -> finalize -> row exchange) on
//     its own stream, the list lengths stay on the device, so the host enqueues a whole frame without waiting once;
//   * a group = the N handles of a box + one host thread each, alive across frames, behind one host-buffer call.
#include "engine.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <chrono>
#include <thread>
#include <unistd.h>

using namespace bcdk;
using namespace bcdeng;

namespace
{

constexpr int ARENA_FLAG_WORDS = 16;
// flag / ticket slots of a level: kind + side (the side of the RECEIVER the strip comes from)
enum { SLOT_ACC = 0, SLOT_OUT_A = 2, SLOT_OUT_B = 4, SLOT_AS = 6, SLOTS_PER_LEVEL = 8 };
constexpr int MAX_LEVELS = 16;
constexpr int TIMEOUT_WORD = MAX_LEVELS * SLOTS_PER_LEVEL; // index of the time-out counter in the ticket buffer

struct Token
{
	cudaIpcMemHandle_t ipc; // 64 bytes
	uint64_t ptr;           // arena address in the exporting process
	uint64_t bytes;
	int64_t pid;
	int32_t device;
	int32_t magic;
};
static_assert(sizeof(Token) <= sizeof(bcd_b200_peer_token), "token layout");
constexpr int32_t TOKEN_MAGIC = 0x42434432; // "BCD2"

size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline int opposite(int side) { return side == BCD_B200_SIDE_UP ? BCD_B200_SIDE_DOWN : BCD_B200_SIDE_UP; }

// arena layout of the current geometry (identical on every tile of the frame: it depends on the level widths only)
size_t layoutArena(bcd_b200_handle* h, const bcd_b200_params& p)
{
	const int s = p.nb_scales, bw = p.search_radius + p.patch_radius;
	h->peer.layout.assign(s, bcd_b200_handle::ArenaLevel());
	size_t off = 0;
	for(int k = 0; k < s; ++k)
	{
		bcd_b200_handle::ArenaLevel& a = h->peer.layout[k];
		const int W = h->levels[k].W;
		a.flags = off; off = alignUp(off + ARENA_FLAG_WORDS * 4, 256);
		a.inbox = off; off = alignUp(off + (size_t)6 * markingRingPitch(W) * 4, 256);
		a.asStrip = off; off = alignUp(off + (size_t)p.search_radius * markingRowWords(W) * 4, 256);
		for(int side = 0; side < 2; ++side) { a.accHalo[side] = off; off = alignUp(off + (size_t)bw * W * sizeof(float4), 256); }
		for(int use = 0; use < 2; ++use)
			for(int side = 0; side < 2; ++side) { a.outRows[use][side] = off; off = alignUp(off + (size_t)2 * W * 3 * sizeof(float), 256); }
	}
	return off;
}

void applyTile(bcd_b200_handle* h, const bcd_b200_params& p, const bcd_b200_tile& t)
{
	h->tiled = t.tile_count > 1;
	h->imageH = t.image_height;
	h->subRow0 = t.sub_row0; h->subRow1 = t.sub_row1;
	h->ownRow0 = t.own_row0; h->ownRow1 = t.own_row1;
	h->tileIndex = t.tile_index; h->tileCount = t.tile_count;
	h->runParams = p;
}

// Arenas of other processes mapped into this one. An IPC handle can be opened once per process and device, but with
// several tiles per GPU two local tiles are neighbours of the same remote tile (tile j + 1 lies between the local tiles j
// and j + 2 when the box is split over two processes): the mapping is shared and reference-counted.
struct IpcMapping { void* ptr; int refs; };
std::mutex g_ipcMutex;
std::map<std::tuple<int, int64_t, uint64_t>, IpcMapping> g_ipcMappings; // (local device, exporting pid, address there)

cudaError_t openArena(int device, const Token& tok, void** out)
{
	std::lock_guard<std::mutex> lk(g_ipcMutex);
	const auto key = std::make_tuple(device, tok.pid, tok.ptr);
	auto it = g_ipcMappings.find(key);
	if(it != g_ipcMappings.end())
	{
		++it->second.refs;
		*out = it->second.ptr;
		return cudaSuccess;
	}
	void* ptr = nullptr;
	cudaError_t e = cudaIpcOpenMemHandle(&ptr, tok.ipc, cudaIpcMemLazyEnablePeerAccess);
	if(e != cudaSuccess) { cudaGetLastError(); return e; }
	g_ipcMappings[key] = IpcMapping{ptr, 1};
	*out = ptr;
	return cudaSuccess;
}

void closeArena(void* ptr)
{
	std::lock_guard<std::mutex> lk(g_ipcMutex);
	for(auto it = g_ipcMappings.begin(); it != g_ipcMappings.end(); ++it)
		if(it->second.ptr == ptr)
		{
			if(--it->second.refs == 0)
			{
				cudaIpcCloseMemHandle(ptr);
				g_ipcMappings.erase(it);
			}
			return;
		}
}

void closeNeighbour(bcd_b200_handle* h, int side)
{
	if(h->peer.nb[side] && h->peer.nbIpc[side])
		closeArena(h->peer.nb[side]);
	h->peer.nb[side] = nullptr;
	h->peer.nbIpc[side] = false;
}

struct Ctx // what the strip helpers need
{
	bcd_b200_handle* h;
	int level;
	bcd_b200_handle::ArenaLevel a;
	uint32_t epoch;
	bool has[2];
	unsigned int* tickets;
	uint32_t* myFlags;
	char* arena;
};

Ctx ctxOf(bcd_b200_handle* h, int k)
{
	Ctx c;
	c.h = h; c.level = k; c.a = h->peer.layout[k]; c.epoch = h->peer.epoch;
	c.has[BCD_B200_SIDE_UP] = !h->peer.solo && h->tileIndex > 0;
	c.has[BCD_B200_SIDE_DOWN] = !h->peer.solo && h->tileIndex < h->tileCount - 1;
	c.tickets = h->peer.tickets.as<unsigned int>();
	c.arena = h->peer.arena.as<char>();
	c.myFlags = reinterpret_cast<uint32_t*>(c.arena + c.a.flags);
	return c;
}

// rows [row, row + rows) of a local image go to the neighbour on `side`, into `slotOff` of its arena; flag slot `slot`
int pushRows(const Ctx& c, int side, const char* base, size_t rowBytes, int row, int rows, size_t slotOff, int slot, cudaStream_t st)
{
	if(!c.has[side]) return 0;
	bcd_b200_handle* h = c.h;
	char* nb = h->peer.nb[side];
	// the receiver sees this strip arrive from its opposite side
	uint32_t* flag = reinterpret_cast<uint32_t*>(nb + c.a.flags) + slot + opposite(side);
	CU(launchPeerPush(base + (size_t)row * rowBytes, nb + slotOff, rowBytes * rows, flag, c.epoch,
			c.tickets + c.level * SLOTS_PER_LEVEL + slot + side, st, &h->launches));
	return 0;
}

// the strip the neighbour on `side` pushed into `slotOff` of my arena lands in rows [row, row + rows) (added when `add`)
int recvRows(const Ctx& c, int side, char* base, size_t rowBytes, int row, int rows, size_t slotOff, int slot, bool add, cudaStream_t st)
{
	if(!c.has[side]) return 0;
	bcd_b200_handle* h = c.h;
	CU(launchPeerRecv(base + (size_t)row * rowBytes, c.arena + slotOff, rowBytes * rows, add, c.myFlags + slot + side, c.epoch,
			c.tickets + TIMEOUT_WORD, st, &h->launches));
	return 0;
}

// owned patch-centre rows of tile `index` at level k, like levelGeom() computes them for the handle's own tile
bool tileLevelGeom(const bcd_b200_handle* h, int index, int k, const bcd_b200_params& p, LevelGeom& g)
{
	bcd_b200_tile t;
	if(bcd_b200_tile_plan(h->W, h->imageH, index, h->tileCount, &p, &t) != 0)
		return false;
	int W = h->W, H = t.sub_row1 - t.sub_row0;
	for(int l = 0; l < k; ++l) { W /= 2; H /= 2; }
	g = makeGeom(W, H, h->B, p.patch_radius, p.search_radius);
	const int own0 = (t.own_row0 - t.sub_row0) >> k;
	const int own1 = t.own_row1 >= h->imageH ? H : (t.own_row1 - t.sub_row0) >> k;
	g.rowBeg = std::max(g.w, own0);
	g.rowEnd = std::min(H - g.w, own1);
	return true;
}

// Can the marking chain cross the boundary below tile `upper` word by word (ring of the systolic kernel)? Both tiles
// must run that kernel at this level, and the last band of the upper tile must hold the six hand-off rows. Both sides
// of the boundary evaluate this identically; otherwise the AS rows travel as a strip after the upper tile's kernel.
bool ringAcross(const bcd_b200_handle* h, int upper, int k, const bcd_b200_params& p)
{
	LevelGeom gu, gl;
	if(!tileLevelGeom(h, upper, k, p, gu) || !tileLevelGeom(h, upper + 1, k, p, gl))
		return false;
	if(!markingUsesRing(gu) || !markingUsesRing(gl))
		return false;
	const int rows = gu.rowEnd - gu.rowBeg;
	const int lastBand = rows - (divUp(rows, 32) - 1) * 32;
	return lastBand >= 6;
}

// marking chain of level k in peer mode, on the level's stream
int markPeer(bcd_b200_handle* h, int k, const bcd_b200_params& p)
{
	LevelBufs& L = h->levels[k];
	const LevelGeom g = levelGeom(h, k, p);
	const Ctx c = ctxOf(h, k);
	cudaStream_t st = auxOf(h, k);
	const int Wp = markingRowWords(g.W), b = p.search_radius;
	int own0, own1;
	tileOwnedRows(h, k, own0, own1);
	if(L.selectPending)
	{
		// patch selection of the level at the head of its own stream (see levelPairAndPatchSelect)
		CU(cudaStreamWaitEvent(st, L.evPair, 0));
		int rc = patchSelectRows(h, k, p, -1, -1, st, L.selectPrivate);
		if(rc) return rc;
		L.selectPending = false;
		CU(cudaEventRecord(L.evSel, st));
	}
	else
		CU(cudaStreamWaitEvent(st, L.evSel, 0));
	STAGE_ON(h, st, BCD_B200_STAGE_MARKING, {
		if(p.skip_prob > 0.f)
		{
			const bool ringUp = c.has[BCD_B200_SIDE_UP] && ringAcross(h, h->tileIndex - 1, k, p);
			const bool ringDown = c.has[BCD_B200_SIDE_DOWN] && ringAcross(h, h->tileIndex, k, p);
			const size_t rowBytes = (size_t)Wp * 4;
			if(c.has[BCD_B200_SIDE_UP] && !ringUp)
			{
				int rc = recvRows(c, BCD_B200_SIDE_UP, (char*)L.asBits.p, rowBytes, own0 - b, b, c.a.asStrip, SLOT_AS, false, st);
				if(rc) return rc;
			}
			const uint32_t* ringIn = ringUp ? reinterpret_cast<const uint32_t*>(c.arena + c.a.inbox) : nullptr;
			uint32_t* ringOut = ringDown ? reinterpret_cast<uint32_t*>(h->peer.nb[BCD_B200_SIDE_DOWN] + c.a.inbox) : nullptr;
			CU(launchMarking(L.pre.as<uint32_t>(), prePitchOf(g.W), g, p.skip_prob, 0x9E3779B9u + (uint32_t)k, L.active.as<uint32_t>(),
					L.asBits.as<uint32_t>(), Wp, L.progress.as<unsigned int>(), st, &h->launches,
					&L.ctr.as<DeviceCounters>()->markTimeouts, c.epoch, ringIn, ringOut));
			if(c.has[BCD_B200_SIDE_DOWN] && !ringDown)
			{
				int rc = pushRows(c, BCD_B200_SIDE_DOWN, (const char*)L.asBits.p, rowBytes, own1 - b, b, c.a.asStrip, SLOT_AS, st);
				if(rc) return rc;
			}
		}
		else
			CU(launchAllActive(g, L.active.as<uint32_t>(), Wp, st, &h->launches));
		CU(launchBuildLists(L.counts.as<int>(), L.active.as<uint32_t>(), Wp, g, L.fullList.as<int>(),
				L.fallbackList.as<int>(), L.ctr.as<DeviceCounters>(), st, &h->launches));
		// work counters of the level: statistics, and (whole-frame runs) the list lengths the host sizes the filter with
		CU(cudaMemcpyAsync(L.hostCtr, L.ctr.p, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, st));
	});
	CU(cudaEventRecord(L.evMark, st));
	return 0;
}

// filter -> accumulator halos -> finalize -> output rows of level k, on the level's stream
int filterPeer(bcd_b200_handle* h, int k, const bcd_b200_params& p)
{
	LevelBufs& L = h->levels[k];
	const LevelGeom g = levelGeom(h, k, p);
	const Ctx c = ctxOf(h, k);
	cudaStream_t st = auxOf(h, k);
	const int bw = p.search_radius + p.patch_radius;
	int own0, own1;
	tileOwnedRows(h, k, own0, own1);
	CU(cudaStreamWaitEvent(st, h->peer.evColor, 0));
	int rc;
	if(h->tiled)
		rc = levelFilterOn(h, k, p, st, true, true); // list lengths stay on the device: no host hop in the chain of the tiles
	else
	{
		// whole frame: nothing depends on another GPU, and the host has queued all the work that does not need the
		// counters — waiting for them here costs no device time and spares the empty chunk launches
		rc = levelMarkWait(h, k);
		if(!rc) rc = levelFilterOn(h, k, p, st, false, true);
	}
	if(rc) return rc;
	const size_t accRow = (size_t)g.W * sizeof(float4), outRow = (size_t)g.W * 3 * sizeof(float);
	STAGE_ON(h, st, BCD_B200_STAGE_EXCHANGE, {
		// partial sums accumulated outside my band go to their owner and are ADDED there
		if((rc = pushRows(c, BCD_B200_SIDE_UP, (const char*)L.acc.p, accRow, own0 - bw, bw, c.a.accHalo[BCD_B200_SIDE_DOWN], SLOT_ACC, st))) return rc;
		if((rc = pushRows(c, BCD_B200_SIDE_DOWN, (const char*)L.acc.p, accRow, own1, bw, c.a.accHalo[BCD_B200_SIDE_UP], SLOT_ACC, st))) return rc;
		if((rc = recvRows(c, BCD_B200_SIDE_UP, (char*)L.acc.p, accRow, own0, bw, c.a.accHalo[BCD_B200_SIDE_UP], SLOT_ACC, true, st))) return rc;
		if((rc = recvRows(c, BCD_B200_SIDE_DOWN, (char*)L.acc.p, accRow, own1 - bw, bw, c.a.accHalo[BCD_B200_SIDE_DOWN], SLOT_ACC, true, st))) return rc;
	});
	STAGE_ON(h, st, BCD_B200_STAGE_FINALIZE, {
		CU(launchFinalize(L.acc.as<float4>(), (size_t)g.W * g.H, L.out.as<float>(), st, &h->launches));
	});
	STAGE_ON(h, st, BCD_B200_STAGE_EXCHANGE, {
		// 2 output rows each way: downscale + interpolate of the merge reach 2 rows into the neighbouring band
		if((rc = pushRows(c, BCD_B200_SIDE_UP, (const char*)L.out.p, outRow, own0, 2, c.a.outRows[0][BCD_B200_SIDE_DOWN], SLOT_OUT_A, st))) return rc;
		if((rc = pushRows(c, BCD_B200_SIDE_DOWN, (const char*)L.out.p, outRow, own1 - 2, 2, c.a.outRows[0][BCD_B200_SIDE_UP], SLOT_OUT_A, st))) return rc;
		if((rc = recvRows(c, BCD_B200_SIDE_UP, (char*)L.out.p, outRow, own0 - 2, 2, c.a.outRows[0][BCD_B200_SIDE_UP], SLOT_OUT_A, false, st))) return rc;
		if((rc = recvRows(c, BCD_B200_SIDE_DOWN, (char*)L.out.p, outRow, own1, 2, c.a.outRows[0][BCD_B200_SIDE_DOWN], SLOT_OUT_A, false, st))) return rc;
	});
	CU(cudaEventRecord(h->peer.evOut[k], st));
	return 0;
}

// merged rows of level k (>= 1) feed the merge of level k - 1 on the neighbouring tiles
int mergedRowsPeer(bcd_b200_handle* h, int k, const bcd_b200_params& p)
{
	LevelBufs& L = h->levels[k];
	const Ctx c = ctxOf(h, k);
	cudaStream_t st = h->stream;
	int own0, own1, rc;
	tileOwnedRows(h, k, own0, own1);
	const size_t outRow = (size_t)L.W * 3 * sizeof(float);
	(void)p;
	STAGE_ON(h, st, BCD_B200_STAGE_EXCHANGE, {
		if((rc = pushRows(c, BCD_B200_SIDE_UP, (const char*)L.out.p, outRow, own0, 2, c.a.outRows[1][BCD_B200_SIDE_DOWN], SLOT_OUT_B, st))) return rc;
		if((rc = pushRows(c, BCD_B200_SIDE_DOWN, (const char*)L.out.p, outRow, own1 - 2, 2, c.a.outRows[1][BCD_B200_SIDE_UP], SLOT_OUT_B, st))) return rc;
		if((rc = recvRows(c, BCD_B200_SIDE_UP, (char*)L.out.p, outRow, own0 - 2, 2, c.a.outRows[1][BCD_B200_SIDE_UP], SLOT_OUT_B, false, st))) return rc;
		if((rc = recvRows(c, BCD_B200_SIDE_DOWN, (char*)L.out.p, outRow, own1, 2, c.a.outRows[1][BCD_B200_SIDE_DOWN], SLOT_OUT_B, false, st))) return rc;
	});
	return 0;
}

// solo = the warm-up pass of bcd_b200_tile_setup: the same launches without any neighbour (see there)
int tileRun(bcd_b200_handle* h, const bcd_b200_params& p, bool solo = false, bcd_b200_progress_fn progress = nullptr, void* user = nullptr)
{
	if(!h->peer.enabled) return fail(BCD_B200_ERR_STATE, "bcd_b200_tile_run called before bcd_b200_tile_setup");
	if(!h->haveInputs) return fail(BCD_B200_ERR_STATE, "bcd_b200_tile_run called before upload/bind_device");
	const bcd_b200_params& q = h->runParams;
	if(p.patch_radius != q.patch_radius || p.search_radius != q.search_radius || p.nb_scales != q.nb_scales)
		return fail(BCD_B200_ERR_STATE, "radii / scales differ from bcd_b200_tile_setup (w=%d b=%d s=%d)", q.patch_radius, q.search_radius, q.nb_scales);
	if(h->H != h->subRow1 - h->subRow0 || h->W != h->peer.geomW)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "resident inputs are %dx%d but the tile was set up for %dx%d", h->W, h->H,
				h->peer.geomW, h->subRow1 - h->subRow0);
	// per-run switches of the handle, cleared again on every way out of this function
	struct RunFlags
	{
		bcd_b200_handle* h;
		~RunFlags() { h->deferSelect = false; h->nsEarly = false; h->peer.solo = false; }
	} runFlags{h};
	h->peer.solo = solo;
	for(int side = 0; side < 2 && !solo; ++side)
	{
		const bool has = side == BCD_B200_SIDE_UP ? h->tileIndex > 0 : h->tileIndex < h->tileCount - 1;
		if(has && !h->peer.nb[side])
			return fail(BCD_B200_ERR_STATE, "tile %d of %d is not connected to its %s neighbour (bcd_b200_tile_connect)", h->tileIndex,
					h->tileCount, side == BCD_B200_SIDE_UP ? "upper" : "lower");
	}
	CU(cudaSetDevice(h->device));
	h->runParams = p;
	memset(&h->stats, 0, sizeof(h->stats));
	memset(&h->timings, 0, sizeof(h->timings));
	h->eventsUsed = 0;
	h->spans.clear();
	h->launches = 0;
	h->ran = false;
	h->nsEarly = false;
	h->deferSelect = g_deferSelect;
	if(!solo) h->peer.epoch += 1;
	cudaStream_t st = h->stream;
	const int s = p.nb_scales;
	int rc;
	CU(cudaEventRecord(h->evTotal[0], st));
	h->pyrHistDone = h->pyrColorDone = 0;
	memset(h->pixcovDone, 0, sizeof(h->pixcovDone));
	// (A/B variant BCD_B200_COLOR_FIRST=1) under a streamed upload that sends covariance + colour first, everything that
	// needs only those — the colour side of the pyramid, the per-pixel covariances — goes to the front of the main stream
	const bool colourEarly = h->up.pending && g_colorFirst;
	if(colourEarly)
	{
		CU(cudaStreamWaitEvent(st, h->up.evColorCov, 0));
		for(int k = 1; k < s; ++k) // sample counts first: the covariance of level k + 1 is weighted with those of level k
			if((rc = pyramidLevel(h, k, 4))) return rc;
		h->nsEarly = true;
		for(int k = 1; k < s; ++k)
			if((rc = pyramidLevel(h, k, 2))) return rc;
		for(int k = 0; k < s; ++k)
			if((rc = levelPixelCov(h, k, p))) return rc;
		CU(cudaEventRecord(h->peer.evColor, st));
	}
	// selection of every level on the main stream, finest first (its marking chain is the longest); each level's chain
	// follows on the level's own stream as soon as its selection is there
	int streamedCoarse = 0; // whole-frame runs under a streamed upload: levels whose selection ran band by band behind level 0
	for(int k = 0; k < s; ++k)
	{
		if(h->tiled || k == 0 || k > streamedCoarse)
		{
			if(h->tiled)
				rc = levelSelect(h, k, p);
			else
			{
				rc = k > 0 ? pyramidLevel(h, k, 1) : 0;
				if(!rc) rc = levelClear(h, k, p);
				if(!rc) rc = levelPairAndPatchSelect(h, k, p, k == 0 ? s - 1 : 0);
				if(k == 0) streamedCoarse = h->up.coarseDone;
			}
			if(rc) return rc;
		}
		else
		{
			// selected on the coarse stream already; the accumulator / bit-rows / counters still have to be cleared before
			// the level's marking stream may start: re-record the selection event behind the clear
			CU(cudaStreamWaitEvent(st, h->levels[k].evSel, 0));
			if((rc = levelClear(h, k, p))) return rc;
			CU(cudaEventRecord(h->levels[k].evSel, st));
		}
		if((rc = markPeer(h, k, p))) return rc;
	}
	// colour / covariance side of the pyramid, per-pixel covariances (unless they went first, above)
	if((rc = finishUpload(h))) return rc;
	if(!colourEarly)
	{
		for(int k = 1; k < s; ++k)
			if((rc = pyramidLevel(h, k, 2))) return rc;
		for(int k = 0; k < s; ++k)
			if((rc = levelPixelCov(h, k, p))) return rc;
		CU(cudaEventRecord(h->peer.evColor, st));
	}
	// the levels' filter pipelines run side by side
	for(int i = 0; i < s; ++i)
	{
		// order of issue (whole frames: also the order in which the host waits for the levels' counters): the level whose
		// marking finishes first. Resident inputs: the finest (its chain was started first). Row tiles: the coarsest (its
		// chain through the ranks is the shortest). Streamed upload with the coarser levels selected behind the bands:
		// the coarser levels are through long before level 0, whose last band only arrives with the end of the transfer.
		const int k = (h->tiled || (colourEarly && streamedCoarse > 0)) ? s - 1 - i : i;
		if((rc = filterPeer(h, k, p))) return rc;
	}
	// merges coarsest to finest on the main stream (MultiscaleDenoiser.cpp:79-134)
	CU(cudaStreamWaitEvent(st, h->peer.evOut[s - 1], 0));
	for(int k = s - 2; k >= 0; --k)
	{
		CU(cudaStreamWaitEvent(st, h->peer.evOut[k], 0));
		if((rc = levelMerge(h, k))) return rc;
		if(k > 0 && (rc = mergedRowsPeer(h, k, p))) return rc;
	}
	CU(cudaMemcpyAsync(h->peer.hostTimeouts, h->peer.tickets.as<unsigned int>() + TIMEOUT_WORD, sizeof(unsigned int),
			cudaMemcpyDeviceToHost, st));
	h->deferSelect = false;
	h->nsEarly = false;
	CU(cudaEventRecord(h->evTotal[1], st));
	if(progress)
	{
		// the host has nothing left to do but wait: report the levels as they finish (weights like
		// MultiscaleDenoiser.cpp:85,103-112: a finer scale counts 4x), on the caller's thread
		const float total = float(((1 << (2 * s)) - 1) / 3);
		float doneWeight = 0.f;
		std::vector<bool> seen(s, false);
		while(cudaStreamQuery(st) == cudaErrorNotReady)
		{
			for(int k = 0; k < s; ++k)
				if(!seen[k] && cudaEventQuery(h->peer.evOut[k]) == cudaSuccess)
				{
					seen[k] = true;
					doneWeight += float(1 << (2 * (s - 1 - k)));
					progress(std::min(0.99f, doneWeight / total), user);
				}
			std::this_thread::sleep_for(std::chrono::microseconds(50));
		}
		cudaGetLastError();
	}
	CU(cudaStreamSynchronize(st));
	unsigned int markTimeouts = 0;
	for(int k = 0; k < s && h->tiled; ++k) // (whole-frame runs took the counters in levelMarkWait)
	{
		const DeviceCounters c = *h->levels[k].hostCtr;
		h->levels[k].counters = c;
		markTimeouts += c.markTimeouts;
		addCounters(h, c);
	}
	if(*h->peer.hostTimeouts || markTimeouts)
	{
		const unsigned int t = *h->peer.hostTimeouts;
		CU(cudaMemsetAsync(h->peer.tickets.as<unsigned int>() + TIMEOUT_WORD, 0, sizeof(unsigned int), st));
		CU(cudaStreamSynchronize(st));
		return fail(BCD_B200_ERR_CUDA, "tile %d of %d: %u strip waits and %u marking hand-offs timed out (neighbouring tile lost or out of step)",
				h->tileIndex, h->tileCount, t, markTimeouts);
	}
	h->stats.kernel_launches = h->launches;
	h->timingsPending = true;
	if(h->trace && (rc = collectTimings(h))) return rc;
	h->ran = true;
	return 0;
}

} // namespace

int peerPrepare(bcd_b200_handle* h, const bcd_b200_params* p, const bcd_b200_tile* t, int hist_depth);

namespace bcdeng
{

void releaseNeighbourArena(bcd_b200_handle* h, int side) { closeNeighbour(h, side); }

// bcd_b200_run in its "levels" schedule: the whole frame as one tile without neighbours
int runLevelPipelines(bcd_b200_handle* h, const bcd_b200_params* p, bcd_b200_progress_fn progress, void* user)
{
	bcd_b200_handle::Peer& P = h->peer;
	const bool same = P.enabled && !h->tiled && h->tileCount == 1 && P.geomW == h->W && P.geomH == h->H && P.geomHist == h->B &&
			P.geomB == p->search_radius && P.geomWr == p->patch_radius && P.geomScales == p->nb_scales;
	if(!same)
	{
		// inputs are resident (or streaming in): only the bookkeeping of the level pipelines is (re)built
		bcd_b200_tile t;
		t.image_width = h->W; t.image_height = h->H;
		t.own_row0 = 0; t.own_row1 = h->H; t.sub_row0 = 0; t.sub_row1 = h->H;
		t.tile_index = 0; t.tile_count = 1;
		const bool had = h->haveInputs, pending = h->up.pending;
		int rc = peerPrepare(h, p, &t, h->B);
		h->haveInputs = had;
		h->up.pending = pending;
		if(rc) return rc;
	}
	if(progress) progress(0.f, user);
	int rc = tileRun(h, *p, true, progress, user);
	h->peer.solo = false;
	if(rc) return rc;
	if(progress) progress(1.f, user);
	return 0;
}

} // namespace bcdeng

// buffers, per-level filter scratch, arena, events of the level pipelines for tile t (tile_count == 1: the whole frame)
int peerPrepare(bcd_b200_handle* h, const bcd_b200_params* p, const bcd_b200_tile* t, int hist_depth)
{
	int rc = checkParams(p);
	if(rc) return rc;
	if(hist_depth <= 0 || hist_depth > 255 || t->image_width <= 0 || t->sub_row1 <= t->sub_row0 || t->tile_count < 1 ||
			t->tile_index < 0 || t->tile_index >= t->tile_count)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "bad tile (%d of %d, rows [%d,%d), %d channels)", t->tile_index, t->tile_count,
				t->sub_row0, t->sub_row1, hist_depth);
	if(p->nb_scales > MAX_LEVELS)
		return fail(BCD_B200_ERR_UNSUPPORTED, "too many scales");
	CU(cudaSetDevice(h->device));
	CU(cudaStreamSynchronize(h->stream));
	const int rows = t->sub_row1 - t->sub_row0;
	if(h->haveInputs && (h->W != t->image_width || h->H != rows || h->B != hist_depth))
		h->haveInputs = false; // resident inputs of another geometry
	h->W = t->image_width; h->H = rows; h->B = hist_depth;
	if(h->levels.empty()) h->levels.resize(1);
	applyTile(h, *p, *t);
	rc = ensureBuffers(h, *p);
	if(rc) return rc;
	// everything a run may need is created now: allocations and stream creation synchronise with running kernels, and
	// during a run a neighbouring tile's kernel may already be waiting for this tile
	if(t->tile_count > 1 && (rc = prepareUpload(h, h->W, h->H, h->B))) return rc;
	if((rc = reserveEvents(h, 1024))) return rc;
	// per-level filter scratch: the levels are filtered side by side
	const LevelGeom g0 = makeGeom(h->W, h->H, h->B, p->patch_radius, p->search_radius);
	const size_t perSlot = (size_t)2 * g0.d * g0.d * 4;
	for(int k = 0; k < p->nb_scales; ++k)
	{
		LevelBufs& L = h->levels[k];
		const LevelGeom g = levelGeom(h, k, *p);
		const size_t bound = std::max<size_t>((size_t)std::max(0, g.rowEnd - g.rowBeg) * std::max(0, g.W - 2 * g.w), 1);
		size_t chunk = ((size_t)1536 << 20 >> std::min(2 * k, 8)) / perSlot;
		chunk = std::min(std::max<size_t>(chunk, 1024), bound);
		L.filterChunk = (int)chunk;
		CU(L.fA.reserve(chunk * g.d * g.d * 4));
		CU(L.fMinv.reserve(chunk * g.d * g.d * 4));
		CU(L.fMean.reserve(chunk * g.d * 4));
		CU(L.fNoise.reserve(chunk * g.np * 6 * 4));
		if(k > 0) // pair-term planes of its own: the level's patch selection runs on its stream next to the pair kernel of level k + 1
		{
			const size_t npl = std::max<size_t>((size_t)g.W * g.H, 1);
			CU(L.sPl.reserve(npl * g.J * sizeof(float)));
			CU(L.cPl.reserve(npl * g.J));
			CU(L.mplusl.reserve(npl * g.Jw * 4));
		}
	}
	bcd_b200_handle::Peer& P = h->peer;
	const size_t bytes = layoutArena(h, *p);
	closeNeighbour(h, BCD_B200_SIDE_UP);
	closeNeighbour(h, BCD_B200_SIDE_DOWN);
	CU(P.arena.reserve(bytes));
	P.arenaBytes = bytes;
	CU(cudaMemset(P.arena.p, 0, P.arena.cap)); // flags and ring tags start from a defined state: epoch 0
	P.epoch = 0;
	CU(P.tickets.reserve((TIMEOUT_WORD + 4) * sizeof(unsigned int)));
	CU(cudaMemset(P.tickets.p, 0, P.tickets.cap));
	if(!P.hostTimeouts) CU(cudaMallocHost(&P.hostTimeouts, sizeof(unsigned int)));
	*P.hostTimeouts = 0;
	if(!P.evColor) CU(cudaEventCreateWithFlags(&P.evColor, cudaEventDisableTiming));
	while((int)P.evOut.size() < p->nb_scales)
	{
		cudaEvent_t e;
		CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		P.evOut.push_back(e);
	}
	P.geomW = h->W; P.geomH = h->H; P.geomHist = h->B; P.geomB = p->search_radius; P.geomWr = p->patch_radius; P.geomScales = p->nb_scales;
	P.enabled = true;
	return 0;
}

extern "C"
{

int bcd_b200_tile_setup(bcd_b200_handle* h, const bcd_b200_params* p, const bcd_b200_tile* t, int hist_depth,
		bcd_b200_peer_token* token_out)
{
	if(!h || !t || !token_out) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle / tile / token");
	int rc = peerPrepare(h, p, t, hist_depth);
	if(rc) return rc;
	bcd_b200_handle::Peer& P = h->peer;
	const size_t bytes = P.arenaBytes;
	// Warm-up: one pass over zeroed inputs without neighbours, plus empty launches of the exchange kernels. CUDA loads a
	// kernel the first time it is used and that load waits for the device to go idle — in a real frame it would wait for
	// this tile's kernels that are waiting for a neighbour whose host thread is stuck in the same way. After this pass
	// every kernel the geometry needs is resident.
	{
		LevelBufs& L0 = h->levels[0];
		const float *sc = L0.color, *sn = L0.ns, *sh = L0.hist, *sv = L0.cov;
		const bool hadInputs = h->haveInputs, hadPending = h->up.pending;
		const size_t np = (size_t)h->W * h->H;
		CU(cudaMemsetAsync(L0.ownColor.p, 0, np * 3 * sizeof(float), h->stream));
		CU(cudaMemsetAsync(L0.ownNs.p, 0, np * sizeof(float), h->stream));
		CU(cudaMemsetAsync(L0.ownHist.p, 0, np * h->B * sizeof(float), h->stream));
		CU(cudaMemsetAsync(L0.ownCov.p, 0, np * 6 * sizeof(float), h->stream));
		L0.color = L0.ownColor.as<float>(); L0.ns = L0.ownNs.as<float>(); L0.hist = L0.ownHist.as<float>(); L0.cov = L0.ownCov.as<float>();
		h->haveInputs = true;
		h->up.pending = false;
		rc = tileRun(h, *p, true);
		if(!rc)
		{
			cudaError_t e = launchPeerWarmup(P.tickets.as<unsigned int>() + TIMEOUT_WORD + 1, h->stream);
			if(e != cudaSuccess) rc = fail(BCD_B200_ERR_CUDA, "warm-up of the exchange kernels failed: %s", cudaGetErrorString(e));
		}
		if(!rc)
		{
			cudaError_t e = launchMarkingWarmup(h->stream);
			if(e != cudaSuccess) rc = fail(BCD_B200_ERR_CUDA, "warm-up of the marking kernels failed: %s", cudaGetErrorString(e));
		}
		L0.color = sc; L0.ns = sn; L0.hist = sh; L0.cov = sv;
		h->haveInputs = hadInputs && sc != L0.ownColor.as<float>(); // uploaded inputs were overwritten, bound ones are intact
		h->up.pending = false;
		(void)hadPending;
		h->peer.solo = false;
		h->ran = false;
		if(rc) return rc;
	}
	CU(cudaDeviceSynchronize());
	Token tok;
	memset(&tok, 0, sizeof(tok));
	if(cudaIpcGetMemHandle(&tok.ipc, P.arena.p) != cudaSuccess)
		cudaGetLastError(); // same-process groups do not need it
	tok.ptr = (uint64_t)(uintptr_t)P.arena.p;
	tok.bytes = bytes;
	tok.pid = (int64_t)getpid();
	tok.device = h->device;
	tok.magic = TOKEN_MAGIC;
	memset(token_out, 0, sizeof(*token_out));
	memcpy(token_out->bytes, &tok, sizeof(tok));
	return 0;
}

int bcd_b200_tile_connect(bcd_b200_handle* h, int side, const bcd_b200_peer_token* neighbour)
{
	if(!h || side < 0 || side > 1) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle / bad side");
	if(!h->peer.enabled) return fail(BCD_B200_ERR_STATE, "bcd_b200_tile_connect called before bcd_b200_tile_setup");
	CU(cudaSetDevice(h->device));
	closeNeighbour(h, side);
	if(!neighbour)
		return 0;
	Token tok;
	memcpy(&tok, neighbour->bytes, sizeof(tok));
	if(tok.magic != TOKEN_MAGIC || tok.bytes != h->peer.arenaBytes)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "neighbour token does not match this tile's geometry (arena %llu vs %llu bytes)",
				(unsigned long long)tok.bytes, (unsigned long long)h->peer.arenaBytes);
	if(tok.pid == (int64_t)getpid())
	{
		if(tok.device != h->device)
		{
			int can = 0;
			CU(cudaDeviceCanAccessPeer(&can, h->device, tok.device));
			if(!can)
				return fail(BCD_B200_ERR_UNSUPPORTED, "device %d cannot access device %d (no peer path)", h->device, tok.device);
			cudaError_t e = cudaDeviceEnablePeerAccess(tok.device, 0);
			if(e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
			else if(e != cudaSuccess) return fail(BCD_B200_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", tok.device, cudaGetErrorString(e));
		}
		h->peer.nb[side] = reinterpret_cast<char*>((uintptr_t)tok.ptr);
		h->peer.nbIpc[side] = false;
	}
	else
	{
		void* ptr = nullptr;
		cudaError_t e = openArena(h->device, tok, &ptr);
		if(e != cudaSuccess)
		{
			return fail(BCD_B200_ERR_CUDA, "cudaIpcOpenMemHandle (arena of device %d, pid %lld) failed: %s", tok.device, (long long)tok.pid,
					cudaGetErrorString(e));
		}
		h->peer.nb[side] = static_cast<char*>(ptr);
		h->peer.nbIpc[side] = true;
	}
	return 0;
}

int bcd_b200_tile_run(bcd_b200_handle* h, const bcd_b200_params* p)
{
	if(!h) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle");
	if(p)
	{
		int rc = checkParams(p);
		if(rc) return rc;
	}
	const bcd_b200_params q = p ? *p : h->runParams;
	return tileRun(h, q);
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// group: the GPUs of one box (or this process' share of them) behind one call
//
// The frame is cut into `tiles` horizontal bands that are dealt out to the GPUs CYCLICALLY: tile j runs on GPU j mod G.
// With several tiles per GPU the raster-order marking chain — the one dependency through all rows of the frame —
// pipelines: while the GPUs select their second tiles the chain already runs through the first G tiles, and the filters
// follow it band by band, instead of the chain starting only when whole eighths of the frame have been selected.
// Every local tile has its own handle (streams, buffers, arena) and its own host thread, alive across frames.
// ---------------------------------------------------------------------------------------------------------------------
struct bcd_b200_group
{
	int rank = 0, world = 1;       // this process' position when the GPUs of the box are spread over several processes
	std::vector<int> dev;          // CUDA devices of this process
	int tilesPerGpu = 0;           // 0: chosen from the frame height
	int hostTiles = 1;             // one GPU, frames from host memory: tiles whose chained uploads pipeline with the kernels
	                               // (BCD_B200_HOST_TILES; measured slower than the whole frame so far, see chooseTiles)
	bool hostPath = false;         // the current setup is for bcd_b200_group_denoise (host buffers)
	bool pin = true;
	std::vector<std::pair<const void*, size_t>> pinned;

	struct Slot // one local tile
	{
		bcd_b200_group* g = nullptr;
		int slot = 0, device = 0;
		bcd_b200_handle* h = nullptr;
		std::thread th;
		bcd_b200_tile plan;
		bcd_b200_peer_token token;
		bcd_b200_images hostIn;    // this tile's rows in host memory (e2e jobs)
		float* hostOut = nullptr;
		int rc = 0;
		std::string err;
	};
	std::vector<Slot*> slots;      // grow-only; the first nLocal are in use for the current frame geometry
	int nLocal = 0;
	int tiles = 0;                 // tiles of the whole frame (all processes)
	std::vector<bcd_b200_peer_token> allTokens; // by tile index
	// worker dispatch
	std::mutex mu;
	std::condition_variable cvGo, cvDone;
	uint64_t gen = 0;
	int pending = 0;
	bool quit = false;
	int job = 0;
	// barrier among the local workers
	std::mutex bmu;
	std::condition_variable bcv;
	int bcount = 0;
	uint64_t bgen = 0;
	// current job
	bcd_b200_params p;
	int W = 0, H = 0, B = 0;
	bool ready = false;            // set up and connected for (W, H, B, radii, scales, tiles)
	int conW = 0, conH = 0, conB = 0, conTiles = 0, conWr = 0, conBr = 0, conS = 0;
	bcd_b200_stats stats;
	bcd_b200_timings timings;
};

namespace
{

enum { JOB_SETUP = 1, JOB_CONNECT = 2, JOB_RUN = 3, JOB_E2E = 4, JOB_RUN_DOWNLOAD = 5 };

typedef bcd_b200_group::Slot Slot;

int gpusTotal(const bcd_b200_group* g) { return g->world * (int)g->dev.size(); }
// tile j runs on global GPU j mod G; global GPU q belongs to process q / nLocalGpus
bool tileIsLocal(const bcd_b200_group* g, int j, int* localGpu)
{
	const int q = j % gpusTotal(g), per = (int)g->dev.size();
	if(q / per != g->rank) return false;
	if(localGpu) *localGpu = q % per;
	return true;
}

void groupBarrier(bcd_b200_group* g, int parties)
{
	std::unique_lock<std::mutex> lk(g->bmu);
	const uint64_t my = g->bgen;
	if(++g->bcount == parties)
	{
		g->bcount = 0;
		++g->bgen;
		g->bcv.notify_all();
	}
	else
		g->bcv.wait(lk, [&] { return g->bgen != my; });
}

void note(Slot& s, int rc)
{
	if(rc && !s.rc) { s.rc = rc; s.err = bcd_b200_last_error(); }
}

void slotJob(Slot& s)
{
	bcd_b200_group* g = s.g;
	const int idx = s.plan.tile_index;
	switch(g->job)
	{
	case JOB_SETUP:
		note(s, bcd_b200_tile_setup(s.h, &g->p, &s.plan, g->B, &s.token));
		break;
	case JOB_CONNECT:
		note(s, bcd_b200_tile_connect(s.h, BCD_B200_SIDE_UP, idx > 0 ? &g->allTokens[idx - 1] : nullptr));
		note(s, bcd_b200_tile_connect(s.h, BCD_B200_SIDE_DOWN, idx < g->tiles - 1 ? &g->allTokens[idx + 1] : nullptr));
		break;
	case JOB_RUN:
		note(s, bcd_b200_tile_run(s.h, &g->p));
		break;
	case JOB_RUN_DOWNLOAD: // the uploads were enqueued (chained) by the dispatching thread
	{
		int rc = bcd_b200_tile_run(s.h, &g->p);
		if(!rc) rc = bcd_b200_tile_download(s.h, s.hostOut);
		note(s, rc);
		break;
	}
	case JOB_E2E:
	{
		// every tile must reach the run (its neighbours wait for its strips), so a local failure is only noted
		int rc = bcd_b200_upload_async(s.h, &s.hostIn);
		if(!rc) rc = bcd_b200_tile_run(s.h, &g->p);
		if(!rc) rc = bcd_b200_tile_download(s.h, s.hostOut);
		note(s, rc);
		break;
	}
	default:
		break;
	}
}

void workerLoop(Slot* s, uint64_t seen)
{
	bcd_b200_group* g = s->g;
	cudaSetDevice(s->device);
	for(;;)
	{
		{
			std::unique_lock<std::mutex> lk(g->mu);
			g->cvGo.wait(lk, [&] { return g->quit || g->gen != seen; });
			if(g->quit) return;
			seen = g->gen;
		}
		if(s->slot < g->nLocal)
			slotJob(*s);
		{
			std::lock_guard<std::mutex> lk(g->mu);
			if(--g->pending == 0)
				g->cvDone.notify_all();
		}
	}
}

// runs `job` on every local tile (one thread each) and returns the first failure
int dispatch(bcd_b200_group* g, int job)
{
	for(int i = 0; i < g->nLocal; ++i) { g->slots[i]->rc = 0; g->slots[i]->err.clear(); }
	if(g->nLocal == 1) // one tile in this process (one process per GPU): no thread hop, the caller drives it
	{
		g->job = job;
		cudaSetDevice(g->slots[0]->device);
		slotJob(*g->slots[0]);
	}
	else
	{
		std::unique_lock<std::mutex> lk(g->mu);
		g->job = job;
		g->pending = (int)g->slots.size();
		++g->gen;
		g->cvGo.notify_all();
		g->cvDone.wait(lk, [&] { return g->pending == 0; });
	}
	for(int i = 0; i < g->nLocal; ++i)
		if(g->slots[i]->rc)
			return fail(g->slots[i]->rc, "tile %d of %d (device %d): %s", g->slots[i]->plan.tile_index, g->tiles, g->slots[i]->device,
					g->slots[i]->err.c_str());
	return 0;
}

int ensureSlots(bcd_b200_group* g, int count)
{
	while((int)g->slots.size() < count)
	{
		Slot* s = new Slot();
		s->g = g;
		s->slot = (int)g->slots.size();
		g->slots.push_back(s);
	}
	return 0;
}

// page-lock a host range the first time it is seen (a renderer reuses its image storage frame after frame)
void pinRange(bcd_b200_group* g, const void* ptr, size_t bytes)
{
	if(!g->pin || !ptr || !bytes) return;
	const char* a = static_cast<const char*>(ptr);
	for(size_t i = 0; i < g->pinned.size();)
	{
		const char* b = static_cast<const char*>(g->pinned[i].first);
		if(b == a && g->pinned[i].second == bytes)
			return;
		if(a < b + g->pinned[i].second && b < a + bytes) // overlaps a stale registration (storage was reallocated)
		{
			if(cudaHostUnregister(const_cast<char*>(b)) != cudaSuccess) cudaGetLastError();
			g->pinned.erase(g->pinned.begin() + i);
		}
		else
			++i;
	}
	if(g->pinned.size() >= 16)
	{
		if(cudaHostUnregister(const_cast<void*>(g->pinned.front().first)) != cudaSuccess) cudaGetLastError();
		g->pinned.erase(g->pinned.begin());
	}
	cudaError_t e = cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterPortable);
	if(e == cudaSuccess)
		g->pinned.push_back({ptr, bytes});
	else
		cudaGetLastError(); // already registered by the caller, or not registrable: the copies still work (pageable)
}

// number of tiles for a frame: tilesPerGpu x GPUs, fewer when the bands would get too thin for the pyramid
int chooseTiles(bcd_b200_group* g, int W, int H, const bcd_b200_params& p, std::vector<bcd_b200_tile>& plan, bool hostPath)
{
	const int G = gpusTotal(g);
	int T = g->tilesPerGpu;
	if(T <= 0)
	{
		// One tile per GPU. Several cyclic tiles per GPU (tile j on GPU j mod G) are meant to let the marking chain — one
		// dependency through all rows — run through the first G tiles while the GPUs still select their second ones. As
		// measured (profiles/r02_scaling.md) that does not pay yet: the tiles of one GPU are driven by independent host
		// threads on streams of equal priority, so their selections share the GPU and finish TOGETHER instead of tile
		// after tile (4096^2 on 4 GPUs: 28.1 ms with 4 tiles per GPU against 28.4 ms with one), and at 1080p the small
		// tiles make a rank launch-bound (8 GPUs: 4.2 / 3.6 ms with 2 / 3 tiles against 3.24 ms). Next step: stream
		// priorities by tile order. bcd_b200_group_set_tiles_per_gpu / BCD_B200_TILES_PER_GPU select it explicitly.
		T = 1;
		// Frames that come from host memory (bcd_b200_group_denoise: what bcd::Denoiser calls) are upload-bound on one
		// GPU: 581 MB at 1080p take 10.5 ms of PCIe 5 against 13 ms of kernels that can only finish once the last row has
		// landed. Cut into a few tiles whose uploads are chained, the first tile selects, marks and filters while the
		// rows of the next ones are still in flight. Measured at 1080p (profiles/r02_host_tiles.md): 21.0 / 19.6 / 20.7 ms
		// with 2 / 3 / 4 tiles against 17.4 ms for the whole frame — the tiles' band kernels get too small (12 bands of
		// 30 rows per tile) and compete with the filters of the tile before — so the default stays 1; BCD_B200_HOST_TILES.
		if(G == 1 && hostPath)
			T = std::max(1, std::min(g->hostTiles, H / 256));
	}
	for(int tiles = G * T; tiles > 1; tiles = tiles > G ? tiles - G : tiles - 1)
	{
		plan.resize(tiles);
		bool ok = true;
		for(int j = 0; j < tiles && ok; ++j)
			ok = bcd_b200_tile_plan(W, H, j, tiles, &p, &plan[j]) == 0;
		if(ok) return tiles;
	}
	plan.resize(1);
	bcd_b200_tile_plan(W, H, 0, 1, &p, &plan[0]);
	return 1;
}

void collectGroupStats(bcd_b200_group* g)
{
	memset(&g->stats, 0, sizeof(g->stats));
	memset(&g->timings, 0, sizeof(g->timings));
	for(int i = 0; i < g->nLocal; ++i)
	{
		const bcd_b200_stats& s = g->slots[i]->h->stats;
		g->stats.interior += s.interior; g->stats.strong += s.strong; g->stats.active += s.active; g->stats.full += s.full;
		g->stats.fallback += s.fallback; g->stats.sum_n_full += s.sum_n_full; g->stats.kernel_launches += s.kernel_launches;
	}
}

// per-stage sums over the local tiles, resolved on demand (the event queries cost ~0.1 ms per tile)
int collectGroupTimings(bcd_b200_group* g)
{
	memset(&g->timings, 0, sizeof(g->timings));
	for(int i = 0; i < g->nLocal; ++i)
	{
		bcd_b200_timings t;
		int rc = bcd_b200_get_timings(g->slots[i]->h, &t);
		if(rc) return rc;
		for(int k = 0; k < BCD_B200_STAGE_COUNT; ++k) g->timings.stage_ms[k] += t.stage_ms[k];
		g->timings.total_ms = std::max(g->timings.total_ms, t.total_ms);
	}
	return 0;
}

} // namespace

extern "C"
{

int bcd_b200_group_create_ranked(bcd_b200_group** out, const int32_t* device_ids, int nb_gpus, int rank, int world)
{
	if(!out) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null group pointer");
	*out = nullptr;
	const int ndev = bcd_b200_device_count();
	if(ndev == 0)
		return fail(BCD_B200_ERR_NO_DEVICE, "no CUDA device available; bcd_b200 has no CPU path");
	if(nb_gpus <= 0) nb_gpus = device_ids ? 1 : ndev; // "all"
	if(nb_gpus > 64 || world < 1 || rank < 0 || rank >= world)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "bad group request (%d GPUs, rank %d of %d)", nb_gpus, rank, world);
	bcd_b200_group* g = new bcd_b200_group();
	g->rank = rank;
	g->world = world;
	memset(&g->stats, 0, sizeof(g->stats));
	memset(&g->timings, 0, sizeof(g->timings));
	{
		const char* e = getenv("BCD_B200_PIN");
		g->pin = !(e && e[0] == '0');
		e = getenv("BCD_B200_TILES_PER_GPU");
		if(e && atoi(e) > 0) g->tilesPerGpu = std::min(atoi(e), 16);
		e = getenv("BCD_B200_HOST_TILES");
		if(e && atoi(e) > 0) g->hostTiles = std::min(atoi(e), 8);
	}
	for(int i = 0; i < nb_gpus; ++i)
	{
		const int d = device_ids ? device_ids[i] : i;
		if(d < 0 || d >= ndev)
		{
			delete g;
			return fail(BCD_B200_ERR_INVALID_ARGUMENT, "device %d out of range (%d devices)", d, ndev);
		}
		g->dev.push_back(d);
	}
	*out = g;
	return 0;
}

int bcd_b200_group_create(bcd_b200_group** out, const int32_t* device_ids, int nb_gpus)
{
	return bcd_b200_group_create_ranked(out, device_ids, nb_gpus, 0, 1);
}

void bcd_b200_group_destroy(bcd_b200_group* g)
{
	if(!g) return;
	{
		std::lock_guard<std::mutex> lk(g->mu);
		g->quit = true;
	}
	g->cvGo.notify_all();
	for(Slot* s : g->slots)
		if(s->th.joinable()) s->th.join();
	// the neighbours' arenas must be unmapped before anyone frees them
	for(Slot* s : g->slots)
		if(s->h && s->h->peer.enabled)
		{
			cudaSetDevice(s->h->device);
			bcd_b200_tile_connect(s->h, BCD_B200_SIDE_UP, nullptr);
			bcd_b200_tile_connect(s->h, BCD_B200_SIDE_DOWN, nullptr);
		}
	for(Slot* s : g->slots)
	{
		bcd_b200_destroy(s->h);
		delete s;
	}
	for(auto& r : g->pinned)
		if(cudaHostUnregister(const_cast<void*>(r.first)) != cudaSuccess) cudaGetLastError();
	delete g;
}

int bcd_b200_group_set_pinning(bcd_b200_group* g, int pin_host)
{
	if(!g) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null group");
	g->pin = pin_host != 0;
	return 0;
}

int bcd_b200_group_set_tiles_per_gpu(bcd_b200_group* g, int tiles_per_gpu)
{
	if(!g) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null group");
	if(tiles_per_gpu < 0 || tiles_per_gpu > 16) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "tiles per GPU must be 0 (automatic) .. 16");
	if(tiles_per_gpu != g->tilesPerGpu) g->ready = false;
	g->tilesPerGpu = tiles_per_gpu;
	return 0;
}

int bcd_b200_group_tiles(bcd_b200_group* g) { return g ? g->tiles : 0; }
int bcd_b200_group_local_tiles(bcd_b200_group* g) { return g ? g->nLocal : 0; }

int bcd_b200_group_setup(bcd_b200_group* g, int width, int height, int hist_depth, const bcd_b200_params* p,
		bcd_b200_peer_token* tokens_out, int32_t* tile_indices_out, int capacity)
{
	if(!g) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null group");
	int rc = checkParams(p);
	if(rc) return rc;
	if(width <= 0 || height <= 0 || hist_depth <= 0)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "bad frame %dx%dx%d", width, height, hist_depth);
	g->ready = false;
	g->p = *p;
	g->W = width; g->H = height; g->B = hist_depth;
	std::vector<bcd_b200_tile> plan;
	g->tiles = chooseTiles(g, width, height, *p, plan, g->hostPath);
	// local tiles, in frame order
	int n = 0;
	for(int j = 0; j < g->tiles; ++j)
	{
		int lg = 0;
		if(!tileIsLocal(g, j, &lg)) continue;
		ensureSlots(g, n + 1);
		Slot& s = *g->slots[n];
		const int device = g->dev[lg];
		if(s.h && s.device != device) // cannot happen: slot i always maps to local GPU i mod (GPUs of this process)
			return fail(BCD_B200_ERR_STATE, "tile slot %d moved from device %d to %d", n, s.device, device);
		s.device = device;
		if(!s.h)
		{
			rc = bcd_b200_create(&s.h, device);
			if(rc) return rc;
		}
		s.plan = plan[j];
		++n;
	}
	g->nLocal = n;
	for(Slot* s : g->slots)
		if(!s->th.joinable())
			s->th = std::thread(workerLoop, s, g->gen); // no job is in flight: gen is stable
	if(g->tiles == 1)
	{
		g->allTokens.clear();
		return 0;
	}
	rc = dispatch(g, JOB_SETUP);
	if(rc) return rc;
	g->allTokens.assign(g->tiles, bcd_b200_peer_token());
	for(int i = 0; i < n; ++i)
		g->allTokens[g->slots[i]->plan.tile_index] = g->slots[i]->token;
	if(tokens_out || tile_indices_out)
	{
		if(capacity < n) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "room for %d tokens, %d local tiles", capacity, n);
		for(int i = 0; i < n; ++i)
		{
			if(tokens_out) tokens_out[i] = g->slots[i]->token;
			if(tile_indices_out) tile_indices_out[i] = g->slots[i]->plan.tile_index;
		}
	}
	return 0;
}

int bcd_b200_group_connect(bcd_b200_group* g, const bcd_b200_peer_token* tokens_by_tile, int nb_tiles)
{
	if(!g) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null group");
	if(g->tiles > 1)
	{
		if(tokens_by_tile)
		{
			if(nb_tiles != g->tiles) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "%d tokens for %d tiles", nb_tiles, g->tiles);
			g->allTokens.assign(tokens_by_tile, tokens_by_tile + nb_tiles);
		}
		int rc = dispatch(g, JOB_CONNECT);
		if(rc) return rc;
	}
	g->ready = true;
	g->conW = g->W; g->conH = g->H; g->conB = g->B; g->conTiles = g->tiles;
	g->conWr = g->p.patch_radius; g->conBr = g->p.search_radius; g->conS = g->p.nb_scales;
	return 0;
}

int bcd_b200_group_tile_info(bcd_b200_group* g, int local_tile, bcd_b200_tile* out)
{
	if(!g || !out || local_tile < 0 || local_tile >= g->nLocal) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "bad local tile %d", local_tile);
	*out = g->slots[local_tile]->plan;
	return 0;
}

int bcd_b200_group_set_tile_buffers(bcd_b200_group* g, int local_tile, const bcd_b200_images* rows_in, float* rows_out)
{
	if(!g || local_tile < 0 || local_tile >= g->nLocal) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "bad local tile %d", local_tile);
	Slot& s = *g->slots[local_tile];
	if(rows_in)
	{
		int rc = checkImages(rows_in);
		if(rc) return rc;
		if(rows_in->width != g->W || rows_in->height != s.plan.sub_row1 - s.plan.sub_row0 || rows_in->hist_depth != g->B)
			return fail(BCD_B200_ERR_INVALID_ARGUMENT, "tile %d holds %dx%d rows, got %dx%d", s.plan.tile_index, g->W,
					s.plan.sub_row1 - s.plan.sub_row0, rows_in->width, rows_in->height);
		s.hostIn = *rows_in;
	}
	s.hostOut = rows_out;
	return 0;
}

int bcd_b200_group_upload_tile(bcd_b200_group* g, int local_tile, const bcd_b200_images* rows_in)
{
	if(!g || local_tile < 0 || local_tile >= g->nLocal) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "bad local tile %d", local_tile);
	return bcd_b200_upload(g->slots[local_tile]->h, rows_in);
}

int bcd_b200_group_run(bcd_b200_group* g, const bcd_b200_params* p, int from_host)
{
	if(!g) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null group");
	if(!g->ready) return fail(BCD_B200_ERR_STATE, "bcd_b200_group_run called before bcd_b200_group_setup / _connect");
	if(p)
	{
		int rc = checkParams(p);
		if(rc) return rc;
		if(p->patch_radius != g->conWr || p->search_radius != g->conBr || p->nb_scales != g->conS)
			return fail(BCD_B200_ERR_STATE, "radii / scales differ from bcd_b200_group_setup");
		g->p = *p;
	}
	int rc;
	if(g->tiles == 1)
	{
		if(g->nLocal < 1) return 0; // the only tile lives in another process
		Slot& s = *g->slots[0];
		CU(cudaSetDevice(s.device));
		rc = 0;
		if(from_host) rc = bcd_b200_upload_async(s.h, &s.hostIn);
		if(!rc) rc = bcd_b200_run(s.h, &g->p, nullptr, nullptr);
		if(!rc && from_host) rc = bcd_b200_download(s.h, s.hostOut);
	}
	else
		rc = dispatch(g, from_host ? JOB_E2E : JOB_RUN);
	if(rc) return rc;
	collectGroupStats(g);
	return 0;
}

int bcd_b200_group_download_tile(bcd_b200_group* g, int local_tile, float* rows_out)
{
	if(!g || local_tile < 0 || local_tile >= g->nLocal) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "bad local tile %d", local_tile);
	Slot& s = *g->slots[local_tile];
	if(g->tiles == 1) return bcd_b200_download(s.h, rows_out);
	return bcd_b200_tile_download(s.h, rows_out);
}

int bcd_b200_group_denoise(bcd_b200_group* g, const bcd_b200_images* in, const bcd_b200_params* p, float* out_rgb,
		bcd_b200_progress_fn progress, void* user)
{
	if(!g) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null group");
	if(g->world != 1) return fail(BCD_B200_ERR_STATE, "bcd_b200_group_denoise needs the whole frame: use the per-tile calls in a ranked group");
	if(!out_rgb) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "Aborting denoising: nullptr for output image");
	int rc = checkImages(in);
	if(rc) return rc;
	rc = checkParams(p);
	if(rc) return rc;
	const size_t np = (size_t)in->width * in->height;
	CU(cudaSetDevice(g->dev[0]));
	pinRange(g, in->color, np * 3 * sizeof(float));
	pinRange(g, in->nsamples, np * sizeof(float));
	pinRange(g, in->hist, np * in->hist_depth * sizeof(float));
	pinRange(g, in->cov, np * 6 * sizeof(float));
	pinRange(g, out_rgb, np * 3 * sizeof(float));
	const bool same = g->ready && g->conW == in->width && g->conH == in->height && g->conB == in->hist_depth &&
			g->conWr == p->patch_radius && g->conBr == p->search_radius && g->conS == p->nb_scales;
	if(!same)
	{
		g->hostPath = true;
		rc = bcd_b200_group_setup(g, in->width, in->height, in->hist_depth, p, nullptr, nullptr, 0);
		g->hostPath = false;
		if(!rc) rc = bcd_b200_group_connect(g, nullptr, 0);
		if(rc) return rc;
	}
	if(g->tiles == 1)
	{
		Slot& s = *g->slots[0];
		CU(cudaSetDevice(s.device));
		rc = bcd_b200_upload_async(s.h, in); // the host buffers outlive the call, so the transfer can stream under the run
		if(!rc) rc = bcd_b200_run(s.h, p, progress, user);
		if(!rc) rc = bcd_b200_download(s.h, out_rgb);
		if(!rc) collectGroupStats(g);
		return rc;
	}
	const size_t W = (size_t)in->width;
	for(int i = 0; i < g->nLocal; ++i)
	{
		Slot& s = *g->slots[i];
		const bcd_b200_tile& t = s.plan;
		s.hostIn = *in;
		s.hostIn.color = in->color + (size_t)t.sub_row0 * W * 3;
		s.hostIn.nsamples = in->nsamples + (size_t)t.sub_row0 * W;
		s.hostIn.hist = in->hist + (size_t)t.sub_row0 * W * in->hist_depth;
		s.hostIn.cov = in->cov + (size_t)t.sub_row0 * W * 6;
		s.hostIn.height = t.sub_row1 - t.sub_row0;
		s.hostOut = out_rgb + (size_t)t.own_row0 * W * 3;
	}
	if(progress) progress(0.f, user);
	// uploads: enqueued here, in tile order; tiles that share a device (and so a PCIe link) are chained, so that the rows
	// of the first tile arrive first and its kernels start while the next tiles are still in flight
	{
		std::map<int, cudaEvent_t> lastOnDevice;
		for(int i = 0; i < g->nLocal; ++i)
		{
			Slot& s = *g->slots[i];
			auto it = lastOnDevice.find(s.device);
			rc = uploadAsyncAfter(s.h, &s.hostIn, it == lastOnDevice.end() ? nullptr : it->second);
			if(rc) return rc;
			lastOnDevice[s.device] = s.h->up.evAll;
		}
	}
	g->p = *p;
	rc = dispatch(g, JOB_RUN_DOWNLOAD);
	if(rc) return rc;
	collectGroupStats(g);
	if(progress) progress(1.f, user);
	return 0;
}

int bcd_b200_group_get_stats(bcd_b200_group* g, bcd_b200_stats* s)
{
	if(!g || !s) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null group / output");
	*s = g->stats;
	return 0;
}

int bcd_b200_group_get_timings(bcd_b200_group* g, bcd_b200_timings* t)
{
	if(!g || !t) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null group / output");
	int rc = collectGroupTimings(g);
	if(rc) return rc;
	*t = g->timings;
	return 0;
}

int bcd_b200_denoise_multi(const bcd_b200_images* in, const bcd_b200_params* p, float* out_rgb, const int32_t* device_ids,
		int nb_gpus, bcd_b200_progress_fn progress, void* user)
{
	if(!out_rgb)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "Aborting denoising: nullptr for output image");
	int rc = checkImages(in);
	if(rc) return rc;
	rc = checkParams(p);
	if(rc) return rc;
	bcd_b200_group* g = nullptr;
	rc = bcd_b200_group_create(&g, device_ids, nb_gpus);
	if(rc) return rc;
	bcd_b200_group_set_pinning(g, 0); // one frame: page-locking the buffers costs more than it saves
	rc = bcd_b200_group_denoise(g, in, p, out_rgb, progress, user);
	std::string msg = rc ? bcd_b200_last_error() : "";
	bcd_b200_group_destroy(g);
	if(rc) return fail(rc, "%s", msg.c_str());
	return 0;
}

} // extern "C"

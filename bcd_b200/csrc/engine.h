// Internal header of the host-side engine (not part of the C ABI): the device buffers, streams and events a
// bcd_b200_handle owns, and the per-level stage functions bcd_b200_capi.cu implements. Shared with bcd_b200_multi.cu,
// which drives the same stages in row-tile mode over peer memory (NVLink) without any host hop between them.
#pragma once
#include "../../include/bcd_b200.h"
#include "kernels.h"

#include <string>
#include <vector>

namespace bcdeng
{
using bcdk::DeviceCounters;
using bcdk::LevelGeom;

// records the thread-local message bcd_b200_last_error() returns; returns `code`
int fail(int code, const char* fmt, ...);

#define CU(call) \
	do { \
		cudaError_t e_ = (call); \
		if(e_ != cudaSuccess) \
			return fail(BCD_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while(0)

// grow-only device buffer
struct DevBuf
{
	void* p = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t bytes)
	{
		if(bytes <= cap)
			return cudaSuccess;
		if(p) cudaFree(p);
		p = nullptr;
		cap = 0;
		cudaError_t e = cudaMalloc(&p, bytes);
		if(e == cudaSuccess) cap = bytes;
		return e;
	}
	void release()
	{
		if(p) cudaFree(p);
		p = nullptr;
		cap = 0;
	}
	template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct LevelBufs
{
	int W = 0, H = 0;
	const float *color = nullptr, *ns = nullptr, *hist = nullptr, *cov = nullptr; // inputs of this level
	DevBuf ownColor, ownNs, ownHist, ownCov;                                         // storage when owned
	DevBuf altColor, altNs, altHist, altCov;                                         // second set (spike prefilter output)
	DevBuf out;                                                                      // W*H*3
	// per-level state of the hot path (kept per level so that the stages of different levels can overlap)
	DevBuf smask, counts, pre, active, asBits, progress, fullList, fallbackList, acc, pixcov, ctr;
	// private pair-term planes / half masks of a coarser level whose selection runs band by band under a streamed upload,
	// interleaved with the level-0 bands that use the shared scratch (allocated on first use)
	DevBuf sPl, cPl, mplusl;
	// filter chain scratch of this level when the levels are filtered side by side (row-tile peer mode)
	DevBuf fA, fMinv, fMean, fNoise;
	int filterChunk = 0;
	DeviceCounters* hostCtr = nullptr; // pinned read-back of ctr
	DeviceCounters counters;           // host copy used to size the filter launches
	cudaEvent_t evSel = nullptr, evMark = nullptr, evPair = nullptr;
	cudaStream_t fbStream = nullptr;                 // side stream of the fallback kernel (runs next to the filter chain)
	cudaEvent_t evFbBegin = nullptr, evFbDone = nullptr;
	bool selectPending = false, selectPrivate = false; // level pipelines: K1b/K1c of this level still to be issued on its stream
};

} // namespace bcdeng

using bcdeng::DevBuf;
using bcdeng::LevelBufs;
using bcdk::LevelGeom;

struct bcd_b200_handle
{
	int device = 0;
	cudaStream_t ownStream = nullptr, stream = nullptr;
	int W = 0, H = 0, B = 0;
	bool haveInputs = false;
	std::vector<LevelBufs> levels; // [0] = full resolution
	// scratch shared by all levels (sized for level 0)
	DevBuf sP, cP, mplus, tmpLow;          // scratch of the selection stage / merge (levels run it one after the other)
	DevBuf fA, fMinv, fMean, fNoise, fX;   // scratch of the filter chain (chunked)
	// marking (K2) runs on a second, high-priority stream so that the latency-bound raster-order chain of level k
	// overlaps the selection of level k-1 and the filters of the coarser levels on the main stream
	cudaStream_t ownAux = nullptr, aux = nullptr;
	std::vector<cudaStream_t> ownAuxLevel; // levels >= 1 get their own marking stream (the chains of different levels are independent)
	std::vector<cudaStream_t> userAuxLevel; // caller-owned marking stream per level (bcd_b200_set_level_aux_stream), or null
	// streamed upload (bcd_b200_upload_async): the histogram arrives in row bands on a copy stream, one event per band;
	// the level-0 pair-term kernel (K1a) is launched band by band behind those events, so the PCIe transfer of the
	// (dominant) histogram overlaps the dominant kernel
	struct Upload
	{
		cudaStream_t copy = nullptr;
		bool pending = false;
		bool histOrdered = false; // the main stream has waited for every histogram chunk of the pending upload
		int bands = 0;
		int rowEnd[16];
		cudaEvent_t ev[16] = {};
		cudaEvent_t evAll = nullptr, evOrder = nullptr;
		cudaEvent_t evColorCov = nullptr; // covariance + colour of the pending upload have landed
		// banded level-0 selection: K1b/K1c of band i-1 run on `sel` next to K1a of band i+1 on the main stream
		cudaStream_t sel = nullptr;
		cudaEvent_t evPair[16] = {};
		cudaEvent_t evSelDone = nullptr;
		int coarseDone = 0; // levels 1..coarseDone had their selection streamed behind the level-0 bands of this run
		cudaStream_t coarse = nullptr; // the stream of that work
	} up;
	int filterChunk = 0;
	LevelGeom g0;      // geometry of level 0 in the last run (debug read-backs)
	bool ran = false;
	bcd_b200_stats stats;
	bcd_b200_timings timings;
	std::vector<cudaEvent_t> events; // grow-only pool; the first eventsUsed are in use by the spans of the current run
	struct Span { int b, e, stage; };
	std::vector<Span> spans;         // timed stage instances of the current run
	int smCount = 148;               // multiprocessors of the device
	bool trace = false;              // BCD_B200_TRACE=1 (read once at bcd_b200_create)
	bool nsEarly = false;            // this run built the sample counts of the pyramid ahead of the histograms (colour-first upload)
	bool deferSelect = false;        // set by the level-pipeline schedule for the duration of a run (see levelPairAndPatchSelect)
	bool timingsPending = false;     // the last run's spans have not been resolved into `timings` yet
	cudaEvent_t evTotal[2] = {nullptr, nullptr};
	size_t eventsUsed = 0;
	int launches = 0;
	// row-tile mode (multi-GPU): the resident inputs are rows [subRow0, subRow1) of an imageH-row frame and this
	// handle produces the output rows [ownRow0, ownRow1) (frame coordinates)
	// row-tile mode builds the pyramid lazily: histogram part of level k before its selection, colour part before its
	// filter (highest level built so far), so that a streamed upload overlaps the level-0 selection here too
	int pyrHistDone = 0, pyrColorDone = 0;
	bool pixcovDone[16] = {};
	bool tiled = false;
	int imageH = 0, subRow0 = 0, subRow1 = 0, ownRow0 = 0, ownRow1 = 0;
	int tileIndex = 0, tileCount = 1;
	// ---- peer mode of the row-tile path (bcd_b200_multi.cu): every strip the neighbouring tiles exchange lands in an
	// "arena" of the receiver (one allocation, so one IPC handle per rank) — pushed there by the sender's kernels over
	// NVLink, followed by an epoch flag — and is consumed by kernels that wait for the flag. No host, no NCCL call sits
	// between the stages.
	struct ArenaLevel // byte offsets inside an arena; identical on every rank (depends on W, b, w and the level only)
	{
		size_t flags = 0;            // ARENA_FLAG_WORDS uint32
		size_t inbox = 0;            // [6][markingRingPitch(W_k)] ring words written by the upper tile's marking kernel
		size_t asStrip = 0;          // [b][Wp] AS bit-rows of the upper tile (marking kernels without a ring)
		size_t accHalo[2] = {0, 0};  // [(b+w)][W_k] float4 partial sums from the neighbour on that side
		size_t outRows[2][2] = {{0, 0}, {0, 0}}; // [use: after finalize / after merge][side]: 2 rows x W_k x 3 floats
	};
	struct Peer
	{
		bool enabled = false;        // tile_run uses the peer path
		bool solo = false;           // warm-up pass of tile_setup: the tile runs without neighbours
		DevBuf arena;
		size_t arenaBytes = 0;
		std::vector<ArenaLevel> layout;
		char* nb[2] = {nullptr, nullptr};  // neighbours' arenas as seen from this device (BCD_B200_SIDE_UP / _DOWN)
		bool nbIpc[2] = {false, false};    // opened with cudaIpcOpenMemHandle (to be closed)
		uint32_t epoch = 0;                // frame counter all tiles of a group agree on (starts at 1)
		DevBuf tickets;                    // "last block signals" counters of the push kernels + the time-out counter
		unsigned int* hostTimeouts = nullptr; // pinned read-back
		cudaEvent_t evColor = nullptr;     // colour / covariance pyramid + per-pixel covariances of all levels ready
		std::vector<cudaEvent_t> evOut;    // level k finalized and its halo rows received
		int geomW = 0, geomH = 0, geomHist = 0, geomB = 0, geomWr = 0, geomScales = 0; // geometry the arena was laid out for
	} peer;
	bcd_b200_params runParams;
};


namespace bcdeng
{

// geometry / rows of level k as seen by the handle (whole frame, or the resident band of a row tile)
bcdk::LevelGeom levelGeom(const bcd_b200_handle* h, int k, const bcd_b200_params& p);
void tileOwnedRows(const bcd_b200_handle* h, int k, int& own0, int& own1);
cudaStream_t auxOf(bcd_b200_handle* h, int k);

extern bool g_schedLevels, g_deferSelect, g_colorFirst;
int patchSelectRows(bcd_b200_handle* h, int k, const bcd_b200_params& p, int r0, int r1, cudaStream_t st, bool privateScratch);
void releaseNeighbourArena(bcd_b200_handle* h, int side); // bcd_b200_multi.cu: unmaps (reference-counted) the neighbour's arena
int checkParams(const bcd_b200_params* p);
int checkImages(const bcd_b200_images* in);
int ensureBuffers(bcd_b200_handle* h, const bcd_b200_params& p);
int finishUpload(bcd_b200_handle* h);
int uploadAsyncAfter(bcd_b200_handle* h, const bcd_b200_images* in, cudaEvent_t after);
int prepareUpload(bcd_b200_handle* h, int width, int height, int histDepth);
int reserveEvents(bcd_b200_handle* h, size_t count);
int pyramidLevel(bcd_b200_handle* h, int k, int parts);
int levelClear(bcd_b200_handle* h, int k, const bcd_b200_params& p);
int levelPixelCov(bcd_b200_handle* h, int k, const bcd_b200_params& p);
int levelSelect(bcd_b200_handle* h, int k, const bcd_b200_params& p);
int levelPairAndPatchSelect(bcd_b200_handle* h, int k, const bcd_b200_params& p, int coarseLevels);
int runLevelPipelines(bcd_b200_handle* h, const bcd_b200_params* p, bcd_b200_progress_fn progress, void* user);
int levelMerge(bcd_b200_handle* h, int k);
int collectTimings(bcd_b200_handle* h);
int stageBegin(bcd_b200_handle* h, int stage, cudaStream_t st);
int stageEnd(bcd_b200_handle* h, cudaStream_t st);
int takeEvent(bcd_b200_handle* h, int* idx);
int levelFilterOn(bcd_b200_handle* h, int k, const bcd_b200_params& p, cudaStream_t st, bool deviceCounts, bool privateScratch);
int levelMarkWait(bcd_b200_handle* h, int k);
void addCounters(bcd_b200_handle* h, const bcdk::DeviceCounters& c);

#define STAGE_ON(h, st_, stage, ...) \
	do { \
		int rc_ = bcdeng::stageBegin(h, stage, st_); \
		if(rc_) return rc_; \
		__VA_ARGS__; \
		rc_ = bcdeng::stageEnd(h, st_); \
		if(rc_) return rc_; \
	} while(0)
#define STAGE(h, stage, ...) STAGE_ON(h, (h)->stream, stage, __VA_ARGS__)

} // namespace bcdeng

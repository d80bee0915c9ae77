// C ABI (include/bcd_b200.h) and the host-side engine that sequences the kernels of the hot path.
//
// Replaces the drivers of the reference (paths relative to /root/reference):
//   Denoiser::denoise            src/core/Denoiser.cpp:84-212        -> Engine::denoiseLevel
//   MultiscaleDenoiser::denoise  src/core/MultiscaleDenoiser.cpp:31-136 -> Engine::run
// Everything data-parallel happens in the kernels (k_*.cu); this file only owns device buffers, streams, events,
// error reporting and the order of launches. There is no CPU fallback: without a CUDA device every entry point fails.

#include "engine.h"

#include <algorithm>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

using namespace bcdk;

namespace
{
thread_local std::string g_lastError;

// A handle keeps several streams busy at once (main, one per pyramid level, copy, banded selection) and, in row-tile
// peer mode, kernels that wait for a neighbour sit at the head of some of them: streams that shared a hardware queue (8
// by default) would serialise behind such a wait. Ask for more queues unless the host application has decided already;
// this only has an effect when the library is loaded before the CUDA context exists.
struct ConnectionsDefault
{
	ConnectionsDefault()
	{
		setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
		// kernels are loaded with the context instead of at their first launch: a lazy load waits for the device to go
		// idle, which a tile whose kernels wait for a neighbour never is (bcd_b200_tile_setup warms up as well)
		setenv("CUDA_MODULE_LOADING", "EAGER", 0);
	}
} g_connectionsDefault;
}

namespace bcdeng
{

int fail(int code, const char* fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_lastError = buf;
	return code;
}

} // namespace bcdeng

using namespace bcdeng;

namespace bcdeng
{

// timed spans: (begin event, end event, stage) triples over a grow-only pool of events; resolved by collectTimings after
// the final synchronisation of the run (no host wait in between)
int takeEvent(bcd_b200_handle* h, int* idx)
{
	if(h->eventsUsed >= h->events.size())
	{
		cudaEvent_t e;
		CU(cudaEventCreate(&e));
		h->events.push_back(e);
	}
	*idx = (int)h->eventsUsed++;
	return 0;
}

int reserveEvents(bcd_b200_handle* h, size_t count)
{
	while(h->events.size() < count)
	{
		cudaEvent_t e;
		CU(cudaEventCreate(&e));
		h->events.push_back(e);
	}
	return 0;
}

int stageBegin(bcd_b200_handle* h, int stage, cudaStream_t st)
{
	int b;
	int rc = takeEvent(h, &b);
	if(rc) return rc;
	CU(cudaEventRecord(h->events[b], st));
	h->spans.push_back({b, -1, stage});
	return 0;
}

int stageEnd(bcd_b200_handle* h, cudaStream_t st)
{
	int e;
	int rc = takeEvent(h, &e);
	if(rc) return rc;
	CU(cudaEventRecord(h->events[e], st));
	h->spans.back().e = e;
	return 0;
}

// Geometry of level k as seen by this handle. Single GPU: the whole level. Row tile: the local sub-image with the
// owned patch-centre rows [rowBeg, rowEnd) and the rows [yLo, yHi) on which the selection stage has to run
// (owned rows +- (b + w): S(p) of an owned p needs the half masks of the b rows above, whose pair terms reach w
// further; the filter gathers colours b + w rows around the owned rows).
LevelGeom levelGeom(const bcd_b200_handle* h, int k, const bcd_b200_params& p)
{
	const LevelBufs& L = h->levels[k];
	LevelGeom g = makeGeom(L.W, L.H, h->B, p.patch_radius, p.search_radius);
	if(h->tiled)
	{
		const int own0 = (h->ownRow0 - h->subRow0) >> k;
		const int own1 = h->ownRow1 >= h->imageH ? L.H : (h->ownRow1 - h->subRow0) >> k;
		g.rowBeg = std::max(g.w, own0);
		g.rowEnd = std::min(L.H - g.w, own1);
		g.yLo = std::max(0, own0 - (g.b + g.w));
		g.yHi = std::min(L.H, own1 + (g.b + g.w));
	}
	return g;
}

void tileOwnedRows(const bcd_b200_handle* h, int k, int& own0, int& own1)
{
	const LevelBufs& L = h->levels[k];
	if(!h->tiled) { own0 = 0; own1 = L.H; return; }
	own0 = (h->ownRow0 - h->subRow0) >> k;
	own1 = h->ownRow1 >= h->imageH ? L.H : (h->ownRow1 - h->subRow0) >> k;
}

// stage 1 (main stream): prepass + distance/selection (K0, K1a-c); clears the accumulators and bit-rows of the level.
// Split in four pieces because the level-0 pieces that only need histogram + sample counts (clear, pair, patch select)
// run under a streamed upload, before the other inputs have arrived.
int levelClear(bcd_b200_handle* h, int k, const bcd_b200_params& p)
{
	LevelBufs& L = h->levels[k];
	const LevelGeom g = levelGeom(h, k, p);
	if(k == 0)
		h->g0 = g;
	cudaStream_t st = h->stream;
	const size_t npix = (size_t)g.W * g.H;
	const int Wp = markingRowWords(g.W);
	STAGE(h, BCD_B200_STAGE_PREPASS, {
		CU(cudaMemsetAsync(L.acc.p, 0, npix * sizeof(float4), st));
		CU(cudaMemsetAsync(L.active.p, 0, (size_t)g.H * Wp * sizeof(uint32_t), st));
		CU(cudaMemsetAsync(L.asBits.p, 0, (size_t)g.H * Wp * sizeof(uint32_t), st));
		CU(cudaMemsetAsync(L.ctr.p, 0, sizeof(DeviceCounters), st));
	});
	return 0;
}

int levelPixelCov(bcd_b200_handle* h, int k, const bcd_b200_params& p)
{
	LevelBufs& L = h->levels[k];
	const LevelGeom g = levelGeom(h, k, p);
	STAGE(h, BCD_B200_STAGE_PREPASS, {
		CU(launchPixelCov(L.cov, L.ns, (size_t)g.W * g.H, L.pixcov.as<float>(), h->stream, &h->launches));
	});
	return 0;
}

// a pending streamed upload becomes an ordinary dependency of the main stream
int finishUpload(bcd_b200_handle* h)
{
	if(h->up.pending)
	{
		CU(cudaStreamWaitEvent(h->stream, h->up.evAll, 0));
		h->up.pending = false;
	}
	return 0;
}

// rows of look-ahead between an upload chunk and the kernel band that consumes it (>= the largest search radius, 15;
// a multiple of the pair kernels' tile heights)
constexpr int UPLOAD_HALO = 16;

// K1b + K1c on the rows [r0, r1) of the level (the whole selection region when r0 < 0)
int patchSelectRows(bcd_b200_handle* h, int k, const bcd_b200_params& p, int r0, int r1, cudaStream_t st, bool privateScratch)
{
	LevelBufs& L = h->levels[k];
	LevelGeom g = levelGeom(h, k, p);
	const int fastLayout = markingPreLayout(g.b);
	LevelGeom gb = g;
	if(r0 >= 0) { gb.yLo = r0; gb.yHi = r1; }
	const float* sP = privateScratch ? L.sPl.as<float>() : h->sP.as<float>();
	const unsigned char* cP = privateScratch ? L.cPl.as<unsigned char>() : h->cP.as<unsigned char>();
	uint32_t* mplus = privateScratch ? L.mplusl.as<uint32_t>() : h->mplus.as<uint32_t>();
	STAGE_ON(h, st, BCD_B200_STAGE_SELECT, {
		CU(launchPatchSelect(sP, cP, gb, p.tau, mplus, st, &h->launches));
		CU(launchSymmetrize(mplus, g, L.smask.as<uint32_t>(), L.counts.as<int>(),
				L.pre.as<uint32_t>(), prePitchOf(g.W), fastLayout, st, &h->launches, gb.yLo, gb.yHi));
	});
	return 0;
}

int clearPre(bcd_b200_handle* h, int k, const bcd_b200_params& p)
{
	LevelBufs& L = h->levels[k];
	const LevelGeom g = levelGeom(h, k, p);
	const int layout = markingPreLayout(g.b);
	const size_t bytes = layout == 2 ? sysPreBytes(g.W, g.rowEnd - g.rowBeg)
	                                 : (size_t)g.H * prePitchOf(g.W) * (layout == 1 ? fastMarkWords(g.b) : g.Pw) * 4;
	CU(cudaMemsetAsync(L.pre.p, 0, bytes, h->stream));
	return 0;
}

// K1a (+ K1b/K1c). With a streamed upload pending (level 0 of a whole-frame run) the kernels are launched once per
// upload band: the pair terms of band [r0, r1) read histogram rows < min(H, r1 + b) (band heights are multiples of the
// tile height), and the patch selection of a band follows as soon as the pair terms of the next band (its w halo rows)
// are there — so only the last band's kernels run after the transfer.
// whole-frame schedule: false = the levels are filtered one after the other on the main stream behind a host read-back
// of their work counters (round 1); true = every level is a pipeline of its own (marking -> filter -> finalize on the
// level's stream, list lengths stay on the device, no host wait inside a frame) — BCD_B200_SCHED=serial|levels
// streamed upload: covariance + colour before the histogram instead of after it (BCD_B200_COLOR_FIRST=1). Measured at
// 1080p: 18.83 ms against 17.41 ms — sent last, the 75 MB of covariance + colour cross PCIe under the last selection band
// and the level-0 marking for free; sent first, they delay the histogram, whose end is what the tail of the frame waits for.
bool g_colorFirst = false;
bool g_fallbackStream = false; // fallback kernel on a side stream next to the filter chain (BCD_B200_FALLBACK_STREAM=1): measured 13.20 vs 13.12 ms, not the default
bool g_schedLevels = true;
// level pipelines: K1b/K1c at the head of the level's stream instead of behind K1a on the main stream (BCD_B200_DEFER_SELECT=0|1)
bool g_deferSelect = false; // measured: 13.29 ms deferred vs 13.19 ms — kept as an A/B hook only
bool g_streamCoarse = true; // BCD_B200_STREAM_COARSE=0: coarser levels start their selection after the upload (A/B hook)

// Under a streamed upload the coarser levels do not have to wait for the frame either: rows < rowEnd >> k of level k
// can be downscaled as soon as the level-0 chunk ending at rowEnd has landed (histogram + sample counts only), and
// their pair terms / selection follow band by band like level 0's — on the main stream, in the gaps the transfer-bound
// level-0 bands leave. State of one coarser level while the bands go through:
struct CoarseStream { int histRows = 0, pairRows = 0, selLo = 0, selHi = 0; };

int coarseBand(bcd_b200_handle* h, const bcd_b200_params& p, int k, int availRows, bool last, CoarseStream& cs)
{
	const LevelBufs& Lp = h->levels[k - 1];
	LevelBufs& L = h->levels[k];
	const LevelGeom g = levelGeom(h, k, p);
	cudaStream_t st = h->up.coarse; // its own stream: small kernels that fill the gaps of the level-0 chain, never delay it
	int rc;
	const int avail = last ? g.H : std::min(g.H, availRows);
	// pair terms of rows [pairRows, pairEnd) read histogram rows < pairEnd + b; bands end on multiples of 8 rows and are
	// at least COARSE_MIN_BAND rows tall (a band of a few rows cannot fill the device)
	constexpr int COARSE_MIN_BAND = 64;
	const int pairEnd = last ? g.H : std::max(cs.pairRows, (avail - UPLOAD_HALO) & ~7);
	if(!last && pairEnd - cs.pairRows < COARSE_MIN_BAND)
		return 0;
	if(avail > cs.histRows)
	{
		STAGE_ON(h, st, BCD_B200_STAGE_PYRAMID, {
			CU(launchDownscale(Lp.color, Lp.ns, Lp.hist, Lp.cov, Lp.W, Lp.H, h->B, L.ownColor.as<float>(), L.ownNs.as<float>(),
					L.ownHist.as<float>(), L.ownCov.as<float>(), st, &h->launches, h->nsEarly ? 8 : 1, cs.histRows, avail - cs.histRows));
		});
		cs.histRows = avail;
	}
	if(pairEnd > cs.pairRows)
	{
		LevelGeom gb = g;
		gb.yLo = cs.pairRows;
		gb.yHi = pairEnd;
		STAGE_ON(h, st, BCD_B200_STAGE_PAIR, {
			CU(launchPairTerms(L.hist, L.ns, gb, L.sPl.as<float>(), L.cPl.as<unsigned char>(), st, &h->launches));
		});
		// the selection of the previous band needed this band's pair terms for its w halo rows
		if(cs.selHi > cs.selLo && (rc = patchSelectRows(h, k, p, cs.selLo, cs.selHi, st, true))) return rc;
		cs.selLo = cs.pairRows;
		cs.selHi = pairEnd;
		cs.pairRows = pairEnd;
	}
	if(last)
	{
		if(cs.selHi > cs.selLo && (rc = patchSelectRows(h, k, p, cs.selLo, cs.selHi, st, true))) return rc;
		cs.selLo = cs.selHi;
		CU(cudaEventRecord(L.evSel, st));
	}
	return 0;
}

int levelPairAndPatchSelect(bcd_b200_handle* h, int k, const bcd_b200_params& p, int coarseLevels)
{
	LevelBufs& L = h->levels[k];
	const LevelGeom g = levelGeom(h, k, p);
	cudaStream_t st = h->stream;
	int rc;
	if(h->up.pending && h->up.bands <= 1)
		if((rc = finishUpload(h))) return rc;
	if((rc = clearPre(h, k, p))) return rc;
	h->up.coarseDone = 0;
	CoarseStream coarse[16];
	if(k == 0 && h->up.pending && coarseLevels > 0 && g_streamCoarse && !h->tiled)
	{
		for(int l = 1; l <= coarseLevels; ++l)
		{
			LevelBufs& Ll = h->levels[l];
			const LevelGeom gl = levelGeom(h, l, p);
			const size_t npl = std::max<size_t>((size_t)gl.W * gl.H, 1);
			CU(Ll.sPl.reserve(npl * gl.J * sizeof(float)));
			CU(Ll.cPl.reserve(npl * gl.J));
			CU(Ll.mplusl.reserve(npl * gl.Jw * 4));
			if((rc = clearPre(h, l, p))) return rc;
		}
		h->up.coarseDone = coarseLevels;
	}
	if(k == 0 && h->up.pending)
	{
		// chunk i of the upload ends UPLOAD_HALO rows below band i of the kernels, so band i waits for chunk i only; the
		// patch selection of band i-1 (needs the pair terms of band i for its w halo rows) runs on a second stream, so
		// the tails of the small band kernels overlap instead of adding up
		bcd_b200_handle::Upload& up = h->up;
		cudaStream_t sel = up.sel;
		CU(cudaEventRecord(up.evOrder, st)); // `pre` cleared, previous users of the scratch done
		CU(cudaStreamWaitEvent(sel, up.evOrder, 0));
		if(up.coarseDone > 0)
			CU(cudaStreamWaitEvent(up.coarse, up.evOrder, 0));
		int r0 = g.yLo, prev0 = 0, prev1 = 0;
		const int nb = up.bands;
		for(int i = 0; i < nb; ++i)
		{
			// band = rows of chunk i minus the look-ahead, clipped to the rows the selection runs on (row tiles)
			const int r1 = std::min(g.yHi, i == nb - 1 ? g.H : up.rowEnd[i] - UPLOAD_HALO);
			CU(cudaStreamWaitEvent(st, up.ev[i], 0));
			if(r1 > r0)
			{
				LevelGeom gb = g;
				gb.yLo = r0;
				gb.yHi = r1;
				STAGE(h, BCD_B200_STAGE_PAIR, {
					CU(launchPairTerms(L.hist, L.ns, gb, h->sP.as<float>(), h->cP.as<unsigned char>(), st, &h->launches));
				});
				CU(cudaEventRecord(up.evPair[i], st));
				CU(cudaStreamWaitEvent(sel, up.evPair[i], 0));
				if(prev1 > prev0 && (rc = patchSelectRows(h, k, p, prev0, prev1, sel, false))) return rc;
				prev0 = r0;
				prev1 = r1;
				r0 = r1;
			}
			// the coarser levels follow on the main stream: level l can use the rows < rowEnd[i] >> l
			// (level l + 1 is built from the rows of level l that have been downscaled so far)
			CU(cudaStreamWaitEvent(up.coarse, up.ev[i], 0));
			for(int l = 1; l <= up.coarseDone; ++l)
			{
				const int avail = l == 1 ? up.rowEnd[i] >> 1 : coarse[l - 1].histRows >> 1;
				if((rc = coarseBand(h, p, l, avail, i == nb - 1, coarse[l]))) return rc;
			}
		}
		if(prev1 > prev0 && (rc = patchSelectRows(h, k, p, prev0, prev1, sel, false))) return rc;
		CU(cudaEventRecord(up.evSelDone, sel));
		CU(cudaStreamWaitEvent(st, up.evSelDone, 0));
		up.histOrdered = true;
	}
	else
	{
		// Level pipelines (h->deferSelect): only the pair terms run here; the patch selection (K1b/K1c: streams the pair
		// planes, memory-bound) moves to the head of the level's own stream, where it overlaps the issue-bound pair kernel
		// of the next level. Every level then needs its own planes: level 0 keeps the shared ones, the coarser levels use
		// their private ones.
		const bool priv = h->deferSelect && k > 0;
		STAGE(h, BCD_B200_STAGE_PAIR, {
			CU(launchPairTerms(L.hist, L.ns, g, priv ? L.sPl.as<float>() : h->sP.as<float>(),
					priv ? L.cPl.as<unsigned char>() : h->cP.as<unsigned char>(), st, &h->launches));
		});
		if(h->deferSelect)
		{
			CU(cudaEventRecord(L.evPair, st));
			L.selectPending = true;
			L.selectPrivate = priv;
			return 0;
		}
		if((rc = patchSelectRows(h, k, p, -1, -1, st, false))) return rc;
	}
	L.selectPending = false;
	CU(cudaEventRecord(L.evSel, st));
	return 0;
}

// level k (>= 1) of the input pyramid from level k-1 (MultiscaleDenoiser.cpp:41-48); parts: 1 = histogram + sample counts
// (all the selection stage needs), 2 = colour + covariance (needed by the filters only)
int pyramidLevel(bcd_b200_handle* h, int k, int parts)
{
	const LevelBufs& Lp = h->levels[k - 1];
	LevelBufs& L = h->levels[k];
	if(parts == 1 && h->nsEarly) parts = 8; // the sample counts of the pyramid were built ahead of the histograms (tileRun)
	STAGE(h, BCD_B200_STAGE_PYRAMID, {
		CU(launchDownscale(Lp.color, Lp.ns, Lp.hist, Lp.cov, Lp.W, Lp.H, h->B, L.ownColor.as<float>(),
				L.ownNs.as<float>(), L.ownHist.as<float>(), L.ownCov.as<float>(), h->stream, &h->launches, parts));
	});
	return 0;
}

int buildPyramid(bcd_b200_handle* h, int s)
{
	for(int k = 1; k < s; ++k)
	{
		int rc = pyramidLevel(h, k, 3);
		if(rc) return rc;
	}
	return 0;
}

// row-tile entry: selection of level k, with the histogram part of the pyramid built on demand
int levelSelect(bcd_b200_handle* h, int k, const bcd_b200_params& p)
{
	int rc = 0;
	if(k > 0 && h->up.pending && !h->up.histOrdered) // the level-0 bands have not been through yet
		rc = finishUpload(h);
	for(int l = h->pyrHistDone + 1; l <= k && !rc; ++l)
	{
		rc = pyramidLevel(h, l, 1);
		h->pyrHistDone = l;
	}
	if(!rc) rc = levelClear(h, k, p);
	if(!rc) rc = levelPairAndPatchSelect(h, k, p, 0);
	return rc;
}

// marking stream of level k: the caller's aux stream when one was set, else one high-priority stream per level
cudaStream_t auxOf(bcd_b200_handle* h, int k)
{
	if(k < (int)h->userAuxLevel.size() && h->userAuxLevel[k])
		return h->userAuxLevel[k];
	if(h->aux != h->ownAux || k == 0 || k > (int)h->ownAuxLevel.size())
		return h->aux;
	return h->ownAuxLevel[k - 1];
}

// stage 2a (aux stream): marking strategy (K2) over the owned rows + work lists; asynchronous
int levelMarkEnqueue(bcd_b200_handle* h, int k, const bcd_b200_params& p)
{
	LevelBufs& L = h->levels[k];
	const LevelGeom g = levelGeom(h, k, p);
	cudaStream_t st = auxOf(h, k);
	const int Wp = markingRowWords(g.W);
	CU(cudaStreamWaitEvent(st, L.evSel, 0));
	STAGE_ON(h, st, BCD_B200_STAGE_MARKING, {
		if(p.skip_prob > 0.f)
			CU(launchMarking(L.pre.as<uint32_t>(), prePitchOf(g.W), g, p.skip_prob, 0x9E3779B9u + (uint32_t)k, L.active.as<uint32_t>(),
					L.asBits.as<uint32_t>(), Wp, L.progress.as<unsigned int>(), st, &h->launches,
					&L.ctr.as<DeviceCounters>()->markTimeouts));
		else
			CU(launchAllActive(g, L.active.as<uint32_t>(), Wp, st, &h->launches));
		CU(launchBuildLists(L.counts.as<int>(), L.active.as<uint32_t>(), Wp, g, L.fullList.as<int>(),
				L.fallbackList.as<int>(), L.ctr.as<DeviceCounters>(), st, &h->launches));
		CU(cudaMemcpyAsync(L.hostCtr, L.ctr.p, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, st));
	});
	CU(cudaEventRecord(L.evMark, st));
	return 0;
}

void addCounters(bcd_b200_handle* h, const DeviceCounters& c)
{
	h->stats.interior += (int64_t)c.interior;
	h->stats.strong += (int64_t)c.strong;
	h->stats.active += (int64_t)c.active;
	h->stats.full += (int64_t)c.full;
	h->stats.fallback += (int64_t)c.fallback;
	h->stats.sum_n_full += (int64_t)c.sum_n_full;
}

// stage 2b (host): wait for the work counters of the level (the only host synchronisation per level)
int levelMarkWait(bcd_b200_handle* h, int k)
{
	LevelBufs& L = h->levels[k];
	CU(cudaEventSynchronize(L.evMark));
	const DeviceCounters c = *L.hostCtr;
	L.counters = c;
	if(c.markTimeouts)
		return fail(BCD_B200_ERR_CUDA, "marking kernel: %u bounded waits expired at level %d (hand-off between row bands lost)", c.markTimeouts, k);
	addCounters(h, c);
	return 0;
}

int levelMark(bcd_b200_handle* h, int k, const bcd_b200_params& p)
{
	int rc = levelMarkEnqueue(h, k, p);
	if(rc) return rc;
	return levelMarkWait(h, k);
}

// stage 3: fallback mean + two-step Bayesian filter, accumulated into acc.
// deviceCounts = false: the host has read the level's work counters (levelMarkWait) and sizes the chunk loop from them;
// true (row tiles in peer mode): the list length stays on the device, no host wait between marking and filter.
// privateScratch: the level has its own chain scratch (levels filtered side by side on their own streams) instead of
// the one shared by all levels of the handle (levels filtered one after the other on the main stream).
int levelFilterOn(bcd_b200_handle* h, int k, const bcd_b200_params& p, cudaStream_t st, bool deviceCounts, bool privateScratch)
{
	LevelBufs& L = h->levels[k];
	const LevelGeom g = levelGeom(h, k, p);
	const DeviceCounters c = L.counters;
	DeviceCounters* dc = L.ctr.as<DeviceCounters>();
	CU(cudaStreamWaitEvent(st, L.evMark, 0));
	// The fallback pixels (mean of the similar patches, DenoisingUnit.cpp:455-481) and the full pixels (two-step filter) are
	// independent work lists that only meet in the accumulator (atomics): with BCD_B200_FALLBACK_STREAM=1 the fallback
	// kernel runs on a side stream next to the filter chain instead of in front of it (A/B hook; no gain measured).
	cudaStream_t fbSt = (g_fallbackStream && L.fbStream) ? L.fbStream : st;
	if(fbSt != st)
	{
		CU(cudaEventRecord(L.evFbBegin, st)); // everything the level's stream has waited for so far (marking, colour)
		CU(cudaStreamWaitEvent(fbSt, L.evFbBegin, 0));
	}
	STAGE_ON(h, fbSt, BCD_B200_STAGE_FALLBACK, {
		if(deviceCounts || c.nFallback > 0)
			CU(launchFallback(L.color, L.smask.as<uint32_t>(), L.counts.as<int>(), g, L.fallbackList.as<int>(),
					&dc->nFallback, L.acc.as<float4>(), fbSt, &h->launches));
	});
	if(fbSt != st)
		CU(cudaEventRecord(L.evFbDone, fbSt));
	FilterBuffers fb;
	fb.A = (privateScratch ? L.fA : h->fA).as<float>();
	fb.Minv = (privateScratch ? L.fMinv : h->fMinv).as<float>();
	fb.mean = (privateScratch ? L.fMean : h->fMean).as<float>();
	fb.noise = (privateScratch ? L.fNoise : h->fNoise).as<float>();
	fb.xScratch = h->fX.as<float>();
	fb.smCount = h->smCount;
	fb.gridCap = h->smCount * 2;
	const int chunk = privateScratch ? L.filterChunk : h->filterChunk;
	// upper bound of the list length the host knows without a read-back: the owned patch centres
	const long long bound = deviceCounts ? (long long)std::max(0, g.rowEnd - g.rowBeg) * std::max(0, g.W - 2 * g.w) : c.nFull;
	for(long long first = 0; first < bound; first += chunk)
	{
		const int count = (int)std::min<long long>(chunk, bound - first);
		int ei[6];
		for(int i = 0; i < 6; ++i)
		{
			int rc = takeEvent(h, &ei[i]);
			if(rc) return rc;
		}
		cudaEvent_t ev[6];
		for(int i = 0; i < 6; ++i) ev[i] = h->events[ei[i]];
		cudaError_t e = launchFilterChunk(L.color, L.pixcov.as<float>(), L.smask.as<uint32_t>(), L.counts.as<int>(), g,
				p.min_eig, L.fullList.as<int>(), (int)first, count, deviceCounts ? &dc->nFull : nullptr, fb, L.acc.as<float4>(), st,
				&h->launches, ev);
		if(e != cudaSuccess)
			return fail(e == cudaErrorInvalidConfiguration ? BCD_B200_ERR_UNSUPPORTED : BCD_B200_ERR_CUDA,
					"filter chunk failed: %s", cudaGetErrorString(e));
		static const int stageOf[5] = {BCD_B200_STAGE_FILTER_STATS, BCD_B200_STAGE_FILTER_EIG, BCD_B200_STAGE_FILTER_APPLY,
				BCD_B200_STAGE_FILTER_EIG, BCD_B200_STAGE_FILTER_APPLY};
		for(int i = 0; i < 5; ++i)
			h->spans.push_back({ei[i], ei[i + 1], stageOf[i]});
	}
	if(fbSt != st)
		CU(cudaStreamWaitEvent(st, L.evFbDone, 0)); // the accumulator is complete behind this point of the level's stream
	return 0;
}

int levelFilter(bcd_b200_handle* h, int k, const bcd_b200_params& p)
{
	return levelFilterOn(h, k, p, h->stream, false, false);
}

// row-tile entry: filters of level k; colour / covariance part of the pyramid and the per-pixel covariances on demand
int levelFilterTile(bcd_b200_handle* h, int k, const bcd_b200_params& p)
{
	int rc = finishUpload(h);
	for(int l = h->pyrColorDone + 1; l <= k && !rc; ++l)
	{
		rc = pyramidLevel(h, l, 2);
		h->pyrColorDone = l;
	}
	if(!rc && !h->pixcovDone[k])
	{
		rc = levelPixelCov(h, k, p);
		h->pixcovDone[k] = true;
	}
	if(!rc) rc = levelFilter(h, k, p);
	return rc;
}

// stage 4 (main stream): out = sum / count
int levelFinalize(bcd_b200_handle* h, int k, const bcd_b200_params& p)
{
	LevelBufs& L = h->levels[k];
	const LevelGeom g = levelGeom(h, k, p);
	STAGE(h, BCD_B200_STAGE_FINALIZE, {
		CU(launchFinalize(L.acc.as<float4>(), (size_t)g.W * g.H, L.out.as<float>(), h->stream, &h->launches));
	});
	return 0;
}

int levelMerge(bcd_b200_handle* h, int k)
{
	STAGE(h, BCD_B200_STAGE_PYRAMID, {
		CU(launchMerge(h->levels[k].out.as<float>(), h->levels[k].W, h->levels[k].H, h->levels[k + 1].out.as<float>(),
				h->tmpLow.as<float>(), h->stream, &h->launches));
	});
	return 0;
}

int collectTimings(bcd_b200_handle* h)
{
	h->timingsPending = false;
	static const char* names[] = {"prepass", "pair", "select", "marking", "fallback", "filter_stats", "filter_eig", "filter_apply",
			"finalize", "pyramid", "exchange"};
	for(const bcd_b200_handle::Span& sp : h->spans)
	{
		if(sp.e < 0) continue;
		float ms = 0.f;
		CU(cudaEventElapsedTime(&ms, h->events[sp.b], h->events[sp.e]));
		h->timings.stage_ms[sp.stage] += ms;
		if(h->trace) // stage timeline (ms since the start of the run) on stderr: which stages overlap
		{
			float t0 = 0.f;
			cudaEventElapsedTime(&t0, h->evTotal[0], h->events[sp.b]);
			fprintf(stderr, "[bcd_b200 trace] dev %d tile %2d %-12s start %8.3f  dur %7.3f\n", h->device, h->tiled ? h->tileIndex : 0,
					sp.stage >= 0 && sp.stage < BCD_B200_STAGE_COUNT ? names[sp.stage] : "?", t0, ms);
		}
	}
	CU(cudaEventElapsedTime(&h->timings.total_ms, h->evTotal[0], h->evTotal[1]));
	if(h->trace)
		fprintf(stderr, "[bcd_b200 trace] dev %d total %8.3f\n", h->device, h->timings.total_ms);
	h->stats.kernel_launches = h->launches;
	return 0;
}

int ensureBuffers(bcd_b200_handle* h, const bcd_b200_params& p)
{
	const int nl = p.nb_scales;
	if((int)h->levels.size() < nl)
		h->levels.resize(nl);
	int W = h->W, H = h->H;
	for(int k = 0; k < nl; ++k)
	{
		LevelBufs& L = h->levels[k];
		L.W = W;
		L.H = H;
		const size_t np = (size_t)W * H;
		CU(L.out.reserve(std::max<size_t>(np * 3 * sizeof(float), 16)));
		if(k > 0)
		{
			CU(L.ownColor.reserve(std::max<size_t>(np * 3 * sizeof(float), 16)));
			CU(L.ownNs.reserve(std::max<size_t>(np * sizeof(float), 16)));
			CU(L.ownHist.reserve(std::max<size_t>(np * h->B * sizeof(float), 16)));
			CU(L.ownCov.reserve(std::max<size_t>(np * 6 * sizeof(float), 16)));
			L.color = L.ownColor.as<float>();
			L.ns = L.ownNs.as<float>();
			L.hist = L.ownHist.as<float>();
			L.cov = L.ownCov.as<float>();
		}
		W /= 2;
		H /= 2;
	}
	const LevelGeom g = makeGeom(h->W, h->H, h->B, p.patch_radius, p.search_radius);
	const size_t npix = (size_t)g.W * g.H;
	CU(h->sP.reserve(npix * g.J * sizeof(float)));
	CU(h->cP.reserve(npix * g.J));
	CU(h->mplus.reserve(npix * g.Jw * 4));
	CU(h->tmpLow.reserve(std::max<size_t>((size_t)(g.W / 2) * (g.H / 2) * 3 * sizeof(float), 16)));
	for(int k = 0; k < nl; ++k)
	{
		LevelBufs& L = h->levels[k];
		const LevelGeom gk = makeGeom(L.W, L.H, h->B, p.patch_radius, p.search_radius);
		const size_t np = std::max<size_t>((size_t)L.W * L.H, 1);
		const size_t rowWords = (size_t)std::max(L.H, 1) * markingRowWords(L.W);
		CU(L.smask.reserve(np * gk.Nw * 4));
		CU(L.pre.reserve(markingPreBytes(L.W, L.H, gk.b, gk.Pw)));
		CU(L.counts.reserve(np * 4));
		CU(L.active.reserve(rowWords * 4));
		CU(L.asBits.reserve(rowWords * 4));
		{
			const void* before = L.progress.p;
			CU(L.progress.reserve(markingScratchWords(L.W, L.H) * 4));
			if(L.progress.p != before) // ring words carry epoch tags: start from a defined state
				CU(cudaMemset(L.progress.p, 0, L.progress.cap));
		}
		CU(L.fullList.reserve(np * 4));
		CU(L.fallbackList.reserve(np * 4));
		CU(L.acc.reserve(np * sizeof(float4)));
		CU(L.pixcov.reserve(np * 6 * sizeof(float)));
		CU(L.ctr.reserve(sizeof(DeviceCounters)));
		if(!L.hostCtr)
			CU(cudaMallocHost(&L.hostCtr, sizeof(DeviceCounters)));
		if(!L.evSel)
			CU(cudaEventCreateWithFlags(&L.evSel, cudaEventDisableTiming));
		if(!L.evMark)
			CU(cudaEventCreateWithFlags(&L.evMark, cudaEventDisableTiming));
		if(!L.evPair)
			CU(cudaEventCreateWithFlags(&L.evPair, cudaEventDisableTiming));
		if(!L.evFbBegin)
			CU(cudaEventCreateWithFlags(&L.evFbBegin, cudaEventDisableTiming));
		if(!L.evFbDone)
			CU(cudaEventCreateWithFlags(&L.evFbDone, cudaEventDisableTiming));
		if(g_fallbackStream && !L.fbStream) // only with the A/B hook: the default path keeps its set of streams
			CU(cudaStreamCreateWithFlags(&L.fbStream, cudaStreamNonBlocking));
	}
	// filter chunk: bounded scratch (two d x d matrices per slot)
	const size_t perSlot = (size_t)2 * g.d * g.d * 4;
	size_t chunk = (size_t)(768u << 20) / perSlot;
	chunk = std::max<size_t>(1024, std::min<size_t>(chunk, 1u << 18));
	chunk = std::min<size_t>(chunk, std::max<size_t>(npix, 1));
	h->filterChunk = (int)chunk;
	CU(h->fA.reserve(chunk * g.d * g.d * 4));
	CU(h->fMinv.reserve(chunk * g.d * g.d * 4));
	CU(h->fMean.reserve(chunk * g.d * 4));
	CU(h->fNoise.reserve(chunk * g.np * 6 * 4));
	// global X scratch only when n_max x d does not fit shared memory next to the d x d inverse
	const size_t xBytes = (size_t)g.Ns * g.d * 4;
	if(xBytes + (size_t)g.d * g.d * 4 + 8 * 1024 > 190 * 1024)
		CU(h->fX.reserve(xBytes * h->smCount * 2));
	while((int)h->ownAuxLevel.size() < nl - 1)
	{
		int lo = 0, hi = 0;
		cudaDeviceGetStreamPriorityRange(&lo, &hi);
		cudaStream_t a = nullptr;
		CU(cudaStreamCreateWithPriority(&a, cudaStreamNonBlocking, hi));
		h->ownAuxLevel.push_back(a);
	}
	for(int i = 0; i < 2; ++i)
		if(!h->evTotal[i]) CU(cudaEventCreate(&h->evTotal[i]));
	return 0;
}

int checkParams(const bcd_b200_params* p)
{
	if(!p)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null parameters");
	if(p->patch_radius < 0 || p->search_radius < 1 || p->nb_scales < 1 || p->nb_scales > 16)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "bad radii / number of scales (w=%d b=%d s=%d)", p->patch_radius,
				p->search_radius, p->nb_scales);
	if(p->search_radius > 15)
		return fail(BCD_B200_ERR_UNSUPPORTED, "search radius %d > 15 is not supported by the marking kernel", p->search_radius);
	const int d = 3 * (2 * p->patch_radius + 1) * (2 * p->patch_radius + 1);
	if(d > 160)
		return fail(BCD_B200_ERR_UNSUPPORTED, "patch radius %d (d=%d) exceeds the shared-memory eigen solver (d <= 160)",
				p->patch_radius, d);
	if(!(p->skip_prob >= 0.f && p->skip_prob <= 1.f))
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "skipping probability %g outside [0,1]", (double)p->skip_prob);
	return 0;
}

int checkImages(const bcd_b200_images* in)
{
	// same conditions Denoiser::inputsOutputsAreOk rejects (reference Denoiser.cpp:269-346)
	if(!in)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "Aborting denoising: nullptr for inputs");
	if(!in->color)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "Aborting denoising: nullptr for input color image");
	if(!in->nsamples)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "Aborting denoising: nullptr for input number of samples image");
	if(!in->hist)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "Aborting denoising: nullptr for input histogram image");
	if(!in->cov)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "Aborting denoising: nullptr for input covariance image");
	if(in->width <= 0 || in->height <= 0 || in->hist_depth <= 0)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "Aborting denoising: input image is empty (%dx%dx%d)", in->width,
				in->height, in->hist_depth);
	if(in->hist_depth > 255)
		return fail(BCD_B200_ERR_UNSUPPORTED, "histogram depth %d > 255 is not supported", in->hist_depth);
	if((size_t)in->width * in->height > (size_t)1 << 30)
		return fail(BCD_B200_ERR_UNSUPPORTED, "image too large");
	return 0;
}

// streams / events of the streamed upload and the storage of the level-0 inputs. Device allocations and stream creation
// synchronise with running kernels, so row tiles do this in bcd_b200_tile_setup — never while a neighbouring tile's
// kernel may already be waiting for this one.
int prepareUpload(bcd_b200_handle* h, int width, int height, int histDepth)
{
	if(h->levels.empty()) h->levels.resize(1);
	LevelBufs& L = h->levels[0];
	bcd_b200_handle::Upload& up = h->up;
	if(!up.copy)
	{
		CU(cudaStreamCreateWithFlags(&up.copy, cudaStreamNonBlocking));
		CU(cudaEventCreateWithFlags(&up.evAll, cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&up.evOrder, cudaEventDisableTiming));
		CU(cudaEventCreateWithFlags(&up.evColorCov, cudaEventDisableTiming));
		{
			// high priority: the short selection kernels of band i-1 take the SMs the pair kernel of band i+1 frees, so
			// the selection never lags more than one band behind
			int lo = 0, hi = 0;
			cudaDeviceGetStreamPriorityRange(&lo, &hi);
			CU(cudaStreamCreateWithPriority(&up.sel, cudaStreamNonBlocking, hi));
		}
		CU(cudaEventCreateWithFlags(&up.evSelDone, cudaEventDisableTiming));
		CU(cudaStreamCreateWithFlags(&up.coarse, cudaStreamNonBlocking));
		for(int i = 0; i < 16; ++i)
		{
			CU(cudaEventCreateWithFlags(&up.ev[i], cudaEventDisableTiming));
			CU(cudaEventCreateWithFlags(&up.evPair[i], cudaEventDisableTiming));
		}
	}
	const size_t np = (size_t)width * height;
	CU(L.ownColor.reserve(np * 3 * sizeof(float)));
	CU(L.ownNs.reserve(np * sizeof(float)));
	CU(L.ownHist.reserve(np * histDepth * sizeof(float)));
	CU(L.ownCov.reserve(np * 6 * sizeof(float)));
	return 0;
}

} // namespace bcdeng

extern "C"
{

const char* bcd_b200_last_error(void) { return g_lastError.c_str(); }
int bcd_b200_abi_version(void) { return BCD_B200_ABI_VERSION; }

int bcd_b200_device_count(void)
{
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess)
	{
		cudaGetLastError();
		return 0;
	}
	return n;
}

void bcd_b200_default_params(bcd_b200_params* p)
{
	if(!p) return;
	p->tau = 1.f;
	p->patch_radius = 1;
	p->search_radius = 6;
	p->min_eig = 0.00000001f;
	p->random_order = 1;
	p->skip_prob = 1.f;
	p->nb_scales = 1;
	p->reserved = 0;
}

int bcd_b200_mask_words(int search_radius)
{
	const int ns = (2 * search_radius + 1) * (2 * search_radius + 1);
	return (ns + 31) / 32;
}

int bcd_b200_create(bcd_b200_handle** out, int device)
{
	if(!out)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle pointer");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if(e != cudaSuccess || n == 0)
	{
		cudaGetLastError();
		return fail(BCD_B200_ERR_NO_DEVICE, "no CUDA device available (%s); bcd_b200 has no CPU path",
				e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
	}
	if(device < 0) device = 0;
	if(device >= n)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "device %d out of range (%d devices)", device, n);
	CU(cudaSetDevice(device));
	bcd_b200_handle* h = new bcd_b200_handle();
	h->device = device;
	{
		const char* e = getenv("BCD_B200_DENSE_PAIR");
		g_forceDensePair = e && e[0] == '1';
		e = getenv("BCD_B200_STREAM_COARSE");
		g_streamCoarse = !(e && e[0] == '0');
		e = getenv("BCD_B200_PAIR");
		g_pairKernel = (e && !strcmp(e, "v1")) ? 1 : 2;
		e = getenv("BCD_B200_GENERIC_MARKING");
		g_forceGenericMarking = e && e[0] == '1';
		e = getenv("BCD_B200_MARKING");
		g_markingKernel = (e && !strcmp(e, "v4")) ? 0 : 1;
		e = getenv("BCD_B200_MARK_TRACE");
		g_markingTrace = (e && e[0] == '1') ? 1 : 0;
		e = getenv("BCD_B200_MARKING_NW");
		g_markingBandsPerCta = (e && atoi(e) >= 1) ? atoi(e) : 4;
		e = getenv("BCD_B200_EIG_GROUP");
		if(e && (atoi(e) == 8 || atoi(e) == 16 || atoi(e) == 32))
			g_eigGroup = atoi(e);
		e = getenv("BCD_B200_FALLBACK_STREAM");
		if(e) g_fallbackStream = e[0] != '0';
		e = getenv("BCD_B200_COLOR_FIRST");
		if(e) g_colorFirst = e[0] != '0';
		e = getenv("BCD_B200_SCHED");
		if(e) g_schedLevels = strcmp(e, "serial") != 0;
		e = getenv("BCD_B200_DEFER_SELECT");
		if(e) g_deferSelect = e[0] != '0';
		e = getenv("BCD_B200_EIG_GENERIC");
		g_eigGenericOrder = e && e[0] == '1';
		e = getenv("BCD_B200_TRACE");
		h->trace = e && e[0] == '1';
		int sms = 0;
		if(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0)
			h->smCount = sms;
	}
	memset(&h->stats, 0, sizeof(h->stats));
	memset(&h->timings, 0, sizeof(h->timings));
	e = cudaStreamCreateWithFlags(&h->ownStream, cudaStreamNonBlocking);
	if(e != cudaSuccess)
	{
		delete h;
		return fail(BCD_B200_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
	}
	h->stream = h->ownStream;
	{
		int lo = 0, hi = 0;
		cudaDeviceGetStreamPriorityRange(&lo, &hi);
		e = cudaStreamCreateWithPriority(&h->ownAux, cudaStreamNonBlocking, hi);
		if(e != cudaSuccess)
		{
			cudaStreamDestroy(h->ownStream);
			delete h;
			return fail(BCD_B200_ERR_CUDA, "cudaStreamCreateWithPriority failed: %s", cudaGetErrorString(e));
		}
		h->aux = h->ownAux;
	}
	*out = h;
	return 0;
}

void bcd_b200_destroy(bcd_b200_handle* h)
{
	if(!h) return;
	cudaSetDevice(h->device);
	if(h->up.copy) cudaStreamSynchronize(h->up.copy);
	cudaStreamSynchronize(h->stream);
	cudaStreamSynchronize(h->aux);
	for(cudaStream_t a : h->ownAuxLevel) { cudaStreamSynchronize(a); cudaStreamDestroy(a); }
	for(int i = 0; i < 16; ++i) if(h->up.ev[i]) cudaEventDestroy(h->up.ev[i]);
	for(int i = 0; i < 16; ++i) if(h->up.evPair[i]) cudaEventDestroy(h->up.evPair[i]);
	if(h->up.evSelDone) cudaEventDestroy(h->up.evSelDone);
	if(h->up.sel) { cudaStreamSynchronize(h->up.sel); cudaStreamDestroy(h->up.sel); }
	if(h->up.coarse) { cudaStreamSynchronize(h->up.coarse); cudaStreamDestroy(h->up.coarse); }
	if(h->up.evAll) cudaEventDestroy(h->up.evAll);
	if(h->up.evOrder) cudaEventDestroy(h->up.evOrder);
	if(h->up.evColorCov) cudaEventDestroy(h->up.evColorCov);
	if(h->up.copy) cudaStreamDestroy(h->up.copy);
	for(auto& L : h->levels)
	{
		DevBuf* lb[] = {&L.ownColor, &L.ownNs, &L.ownHist, &L.ownCov, &L.altColor, &L.altNs, &L.altHist, &L.altCov, &L.out, &L.smask, &L.counts, &L.pre, &L.active, &L.asBits,
				&L.progress, &L.fullList, &L.fallbackList, &L.acc, &L.pixcov, &L.ctr, &L.sPl, &L.cPl, &L.mplusl, &L.fA, &L.fMinv, &L.fMean, &L.fNoise};
		for(DevBuf* b : lb) b->release();
		if(L.hostCtr) cudaFreeHost(L.hostCtr);
		if(L.evSel) cudaEventDestroy(L.evSel);
		if(L.evMark) cudaEventDestroy(L.evMark);
		if(L.evPair) cudaEventDestroy(L.evPair);
		if(L.evFbBegin) cudaEventDestroy(L.evFbBegin);
		if(L.evFbDone) cudaEventDestroy(L.evFbDone);
		if(L.fbStream) { cudaStreamSynchronize(L.fbStream); cudaStreamDestroy(L.fbStream); }
	}
	DevBuf* bufs[] = {&h->sP, &h->cP, &h->mplus, &h->tmpLow, &h->fA, &h->fMinv, &h->fMean, &h->fNoise, &h->fX, &h->peer.tickets};
	for(DevBuf* b : bufs) b->release();
	for(int side = 0; side < 2; ++side)
		releaseNeighbourArena(h, side);
	h->peer.arena.release();
	if(h->peer.hostTimeouts) cudaFreeHost(h->peer.hostTimeouts);
	if(h->peer.evColor) cudaEventDestroy(h->peer.evColor);
	for(cudaEvent_t e : h->peer.evOut) cudaEventDestroy(e);
	for(cudaEvent_t e : h->events) cudaEventDestroy(e);
	for(int i = 0; i < 2; ++i) if(h->evTotal[i]) cudaEventDestroy(h->evTotal[i]);
	if(h->ownAux) cudaStreamDestroy(h->ownAux);
	if(h->ownStream) cudaStreamDestroy(h->ownStream);
	delete h;
}

int bcd_b200_set_stream(bcd_b200_handle* h, void* cuda_stream)
{
	if(!h) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle");
	h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->ownStream;
	return 0;
}

int bcd_b200_set_streams(bcd_b200_handle* h, void* main_stream, void* aux_stream)
{
	if(!h) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle");
	h->stream = main_stream ? reinterpret_cast<cudaStream_t>(main_stream) : h->ownStream;
	h->aux = aux_stream ? reinterpret_cast<cudaStream_t>(aux_stream) : h->ownAux;
	return 0;
}

int bcd_b200_host_register(const void* ptr, size_t bytes)
{
	if(!ptr || !bytes) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null / empty host range");
	cudaError_t e = cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterDefault);
	if(e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
	if(e != cudaSuccess) { cudaGetLastError(); return fail(BCD_B200_ERR_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e)); }
	return 0;
}

int bcd_b200_host_unregister(const void* ptr)
{
	if(!ptr) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null host pointer");
	cudaError_t e = cudaHostUnregister(const_cast<void*>(ptr));
	if(e != cudaSuccess) { cudaGetLastError(); return fail(BCD_B200_ERR_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e)); }
	return 0;
}

int bcd_b200_set_level_aux_stream(bcd_b200_handle* h, int level, void* aux_stream)
{
	if(!h) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle");
	if(level < 0 || level >= 16) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "level %d out of range", level);
	if((int)h->userAuxLevel.size() <= level) h->userAuxLevel.resize(level + 1, nullptr);
	h->userAuxLevel[level] = reinterpret_cast<cudaStream_t>(aux_stream);
	return 0;
}

int bcd_b200_upload_async(bcd_b200_handle* h, const bcd_b200_images* in)
{
	return bcdeng::uploadAsyncAfter(h, in, nullptr);
}

} // extern "C"

// `after` (optional): an event the copies wait for — the group chains the uploads of the tiles that share a device (and
// therefore a PCIe link), so that the first tile's rows arrive first instead of every tile getting a share of the link
int bcdeng::uploadAsyncAfter(bcd_b200_handle* h, const bcd_b200_images* in, cudaEvent_t after)
{
	if(!h) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle");
	int rc = checkImages(in);
	if(rc) return rc;
	CU(cudaSetDevice(h->device));
	rc = prepareUpload(h, in->width, in->height, in->hist_depth);
	if(rc) return rc;
	LevelBufs& L = h->levels[0];
	bcd_b200_handle::Upload& up = h->up;
	const size_t np = (size_t)in->width * in->height;
	// the copies may not overtake work of the main stream that still reads the resident inputs
	CU(cudaEventRecord(up.evOrder, h->stream));
	CU(cudaStreamWaitEvent(up.copy, up.evOrder, 0));
	if(after)
		CU(cudaStreamWaitEvent(up.copy, after, 0));
	// order: what the first kernels need first — sample counts, then the histogram in row bands, then covariance + colour
	// Chunk heights: up to 12 chunks, a small first one (the kernels start early), then heights that shrink slowly
	// (x0.95: on a B200 behind PCIe 5 the banded kernels of a chunk take about as long as its transfer, so the pipeline
	// is balanced and what counts is an early start and a short last band). Chunk ends are multiples of 8 rows, every
	// chunk >= 24 rows.
	const int H = in->height;
	{
		const int nb = std::max(1, std::min(12, H / 24));
		int r = 0, n = 0;
		if(nb >= 4)
		{
			r = std::max(24, std::min(48, H / 16 / 8 * 8));
			up.rowEnd[n++] = r;
		}
		const int rest = nb - n, Hr = H - r;
		double wsum = 0.0, wi = 1.0;
		for(int i = 0; i < rest; ++i, wi *= 0.95) wsum += wi;
		wi = 1.0;
		for(int i = 0; i < rest && r < H; ++i, wi *= 0.95)
		{
			const int rows = std::max(24, (int)(Hr * wi / wsum + 4.0) / 8 * 8);
			int r1 = r + rows;
			if(i == rest - 1 || H - r1 < 24) r1 = H;
			up.rowEnd[n++] = r1;
			r = r1;
		}
		up.bands = n;
	}
	CU(cudaMemcpyAsync(L.ownNs.p, in->nsamples, np * sizeof(float), cudaMemcpyHostToDevice, up.copy));
	if(g_colorFirst)
	{
		// A/B variant: covariance + colour (13 % of the bytes) before the histogram, so that the colour side of the pyramid
		// and the per-pixel covariances run under the transfer (see g_colorFirst for why it is not the default)
		CU(cudaMemcpyAsync(L.ownCov.p, in->cov, np * 6 * sizeof(float), cudaMemcpyHostToDevice, up.copy));
		CU(cudaMemcpyAsync(L.ownColor.p, in->color, np * 3 * sizeof(float), cudaMemcpyHostToDevice, up.copy));
		CU(cudaEventRecord(up.evColorCov, up.copy));
	}
	const size_t rowFloats = (size_t)in->width * in->hist_depth;
	for(int i = 0, r0 = 0; i < up.bands; ++i)
	{
		const int r1 = up.rowEnd[i];
		CU(cudaMemcpyAsync(L.ownHist.as<float>() + (size_t)r0 * rowFloats, in->hist + (size_t)r0 * rowFloats,
				(size_t)(r1 - r0) * rowFloats * sizeof(float), cudaMemcpyHostToDevice, up.copy));
		CU(cudaEventRecord(up.ev[i], up.copy));
		r0 = r1;
	}
	if(!g_colorFirst)
	{
		CU(cudaMemcpyAsync(L.ownCov.p, in->cov, np * 6 * sizeof(float), cudaMemcpyHostToDevice, up.copy));
		CU(cudaMemcpyAsync(L.ownColor.p, in->color, np * 3 * sizeof(float), cudaMemcpyHostToDevice, up.copy));
		CU(cudaEventRecord(up.evColorCov, up.copy));
	}
	CU(cudaEventRecord(up.evAll, up.copy));
	up.pending = true;
	up.histOrdered = false;
	L.color = L.ownColor.as<float>();
	L.ns = L.ownNs.as<float>();
	L.hist = L.ownHist.as<float>();
	L.cov = L.ownCov.as<float>();
	h->W = in->width; h->H = in->height; h->B = in->hist_depth;
	h->haveInputs = true;
	h->ran = false;
	return 0;
}

extern "C"
{

int bcd_b200_upload(bcd_b200_handle* h, const bcd_b200_images* in)
{
	int rc = bcd_b200_upload_async(h, in);
	if(rc) return rc;
	CU(cudaEventSynchronize(h->up.evAll));
	h->up.pending = false;
	return 0;
}

int bcd_b200_bind_device(bcd_b200_handle* h, const bcd_b200_images* in)
{
	if(!h) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle");
	int rc = checkImages(in);
	if(rc) return rc;
	if(h->levels.empty()) h->levels.resize(1);
	if(h->up.pending)
	{
		CU(cudaEventSynchronize(h->up.evAll));
		h->up.pending = false;
	}
	LevelBufs& L = h->levels[0];
	L.color = in->color; L.ns = in->nsamples; L.hist = in->hist; L.cov = in->cov;
	h->W = in->width; h->H = in->height; h->B = in->hist_depth;
	h->haveInputs = true;
	h->ran = false;
	return 0;
}

int bcd_b200_run(bcd_b200_handle* h, const bcd_b200_params* p, bcd_b200_progress_fn progress, void* user)
{
	if(!h) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle");
	if(!h->haveInputs) return fail(BCD_B200_ERR_STATE, "bcd_b200_run called before upload/bind_device");
	int rc = checkParams(p);
	if(rc) return rc;
	CU(cudaSetDevice(h->device));
	if(g_schedLevels)
		return runLevelPipelines(h, p, progress, user);
	h->deferSelect = false;
	h->nsEarly = false;
	rc = ensureBuffers(h, *p);
	if(rc) return rc;
	memset(&h->stats, 0, sizeof(h->stats));
	memset(&h->timings, 0, sizeof(h->timings));
	h->eventsUsed = 0;
	h->spans.clear();
	h->launches = 0;
	cudaStream_t st = h->stream;
	const int s = p->nb_scales;
	CU(cudaEventRecord(h->evTotal[0], st));
	if(progress) progress(0.f, user);
	h->tiled = false;
	h->runParams = *p;
	// Phase 1 — everything that needs only histograms and sample counts, finest level first: under a streamed upload
	// the level-0 kernels start as soon as the first row bands have landed, and the coarser levels' selection runs while
	// covariance and colour are still in flight. Each level's marking chain (K2) follows on its own high-priority stream.
	int streamedCoarse = 0; // levels 1..streamedCoarse: selection already enqueued behind the level-0 bands
	for(int k = 0; k < s; ++k)
	{
		if(k == 0 || k > streamedCoarse)
		{
			if(k > 0 && (rc = pyramidLevel(h, k, 1))) return rc;
			if((rc = levelClear(h, k, *p))) return rc;
			if((rc = levelPairAndPatchSelect(h, k, *p, k == 0 ? s - 1 : 0))) return rc;
			if(k == 0) streamedCoarse = h->up.coarseDone;
		}
		else
		{
			// accumulator / bit-rows / counters are untouched by the selection, but the marking stream must not start
			// before they are cleared: re-record the level's selection event behind the clear
			CU(cudaStreamWaitEvent(st, h->levels[k].evSel, 0)); // selection of level k (coarse stream) done
			if((rc = levelClear(h, k, *p))) return rc;
			CU(cudaEventRecord(h->levels[k].evSel, st));
		}
		if((rc = levelMarkEnqueue(h, k, *p))) return rc;
	}
	// Phase 2 — colour + covariance: rest of the pyramid (MultiscaleDenoiser.cpp:41-48), per-pixel covariances, filters.
	// Level 0 is filtered first (its marking finished under the coarser selections; theirs finish under this filter),
	// then the coarser levels; the merges run coarsest to finest at the end (MultiscaleDenoiser.cpp:79-134).
	if((rc = finishUpload(h))) return rc;
	for(int k = 1; k < s; ++k)
		if((rc = pyramidLevel(h, k, 2))) return rc;
	for(int k = 0; k < s; ++k)
		if((rc = levelPixelCov(h, k, *p))) return rc;
	const float total = float(((1 << (2 * s)) - 1) / 3); // progress weights like MultiscaleDenoiser.cpp:85,103-112
	float doneWeight = 0.f;
	for(int i = 0; i < s; ++i)
	{
		const int k = i == 0 ? 0 : s - i;
		if((rc = levelMarkWait(h, k))) return rc;
		if((rc = levelFilter(h, k, *p))) return rc;
		if((rc = levelFinalize(h, k, *p))) return rc;
		doneWeight += float(1 << (2 * (s - 1 - k)));
		if(progress)
			progress(std::min(1.f, doneWeight / total), user);
	}
	for(int k = s - 2; k >= 0; --k)
		if((rc = levelMerge(h, k))) return rc;
	if(progress)
		progress(1.f, user);
	CU(cudaEventRecord(h->evTotal[1], st));
	CU(cudaStreamSynchronize(st));
	h->stats.kernel_launches = h->launches;
	h->timingsPending = true; // the spans are resolved when somebody asks (bcd_b200_get_timings) or traces
	if(h->trace && (rc = collectTimings(h))) return rc;
	h->ran = true;
	return 0;
}

int bcd_b200_download(bcd_b200_handle* h, float* out_rgb)
{
	if(!h || !out_rgb) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle / output");
	if(!h->ran) return fail(BCD_B200_ERR_STATE, "bcd_b200_download called before a successful run");
	CU(cudaSetDevice(h->device));
	CU(cudaMemcpyAsync(out_rgb, h->levels[0].out.p, (size_t)h->W * h->H * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return 0;
}

int bcd_b200_result_device_ptr(bcd_b200_handle* h, const float** out_dev)
{
	if(!h || !out_dev) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle / output");
	if(!h->ran) return fail(BCD_B200_ERR_STATE, "no result yet");
	*out_dev = h->levels[0].out.as<float>();
	return 0;
}

int bcd_b200_get_stats(bcd_b200_handle* h, bcd_b200_stats* s)
{
	if(!h || !s) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle / output");
	*s = h->stats;
	return 0;
}

int bcd_b200_get_timings(bcd_b200_handle* h, bcd_b200_timings* t)
{
	if(!h || !t) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle / output");
	if(h->timingsPending)
	{
		CU(cudaSetDevice(h->device));
		int rc = collectTimings(h);
		if(rc) return rc;
	}
	*t = h->timings;
	return 0;
}

int bcd_b200_debug_read_selection(bcd_b200_handle* h, uint32_t* masks, int32_t* counts)
{
	if(!h) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle");
	if(!h->ran) return fail(BCD_B200_ERR_STATE, "no run yet");
	CU(cudaSetDevice(h->device));
	const size_t npix = (size_t)h->g0.W * h->g0.H;
	if(masks) CU(cudaMemcpy(masks, h->levels[0].smask.p, npix * h->g0.Nw * 4, cudaMemcpyDeviceToHost));
	if(counts) CU(cudaMemcpy(counts, h->levels[0].counts.p, npix * 4, cudaMemcpyDeviceToHost));
	return 0;
}

int bcd_b200_debug_read_pair_terms(bcd_b200_handle* h, float* sums, uint8_t* counts)
{
	if(!h || !sums || !counts) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null argument");
	if(!h->ran || h->runParams.nb_scales != 1 || h->tiled)
		return fail(BCD_B200_ERR_STATE, "pair terms are only kept after a single-scale whole-frame run");
	CU(cudaSetDevice(h->device));
	const size_t n = (size_t)h->g0.J * h->g0.W * h->g0.H;
	CU(cudaMemcpy(sums, h->sP.p, n * sizeof(float), cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(counts, h->cP.p, n, cudaMemcpyDeviceToHost));
	return 0;
}

int bcd_b200_debug_read_active(bcd_b200_handle* h, uint8_t* active)
{
	if(!h || !active) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle / output");
	if(!h->ran) return fail(BCD_B200_ERR_STATE, "no run yet");
	CU(cudaSetDevice(h->device));
	const int Wp = markingRowWords(h->g0.W);
	std::vector<uint32_t> bits((size_t)h->g0.H * Wp);
	CU(cudaMemcpy(bits.data(), h->levels[0].active.p, bits.size() * 4, cudaMemcpyDeviceToHost));
	for(int y = 0; y < h->g0.H; ++y)
		for(int x = 0; x < h->g0.W; ++x)
		{
			const int pos = x + MARK_PAD;
			active[(size_t)y * h->g0.W + x] = (uint8_t)((bits[(size_t)y * Wp + (pos >> 5)] >> (pos & 31)) & 1u);
		}
	return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// row-tile API
// ---------------------------------------------------------------------------------------------------------------
int bcd_b200_tile_plan(int image_width, int image_height, int tile_index, int tile_count, const bcd_b200_params* p,
		bcd_b200_tile* out)
{
	int rc = checkParams(p);
	if(rc) return rc;
	if(!out || image_width <= 0 || image_height <= 0 || tile_count < 1 || tile_index < 0 || tile_index >= tile_count)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "bad tile request (%d of %d, %dx%d)", tile_index, tile_count, image_width, image_height);
	const int align = 1 << (p->nb_scales - 1);            // band boundaries fall on the same rows at every level
	const int halo = (p->search_radius + p->patch_radius) * align;
	const int units = (image_height + align - 1) / align; // rows in units of `align`
	const int minRows = (p->search_radius + p->patch_radius + 2) << (p->nb_scales - 1);
	// tile boundaries (rows), tile_count + 1 of them
	std::vector<int> bound(tile_count + 1, 0);
	auto equalSplit = [&]() {
		for(int t = 0; t <= tile_count; ++t)
			bound[t] = t == tile_count ? image_height : (int)std::min<long long>((long long)image_height, ((long long)units * t / tile_count) * align);
	};
	equalSplit();
	// (Band heights shaped by a cost model of the marking chain — start_r = max(a R_r, done_{r-1}), done_r = start_r + c R_r,
	// finish_r = done_r + f R_r with the measured a : c : f = 7.2 : 1.07 : 5.6 — were tried and dropped: the optimum of that
	// model is within 1 % of equal bands, because a tile that gets more rows also selects longer and holds the chain up.)
	const int r0 = bound[tile_index], r1 = bound[tile_index + 1];
	// every level must keep at least b+w owned rows (contributions never reach beyond the direct neighbour)
	if(tile_count > 1 && r1 - r0 < minRows)
		return fail(BCD_B200_ERR_UNSUPPORTED, "tile %d has too few rows (%d) for %d scales and b+w=%d", tile_index, r1 - r0,
				p->nb_scales, p->search_radius + p->patch_radius);
	out->image_width = image_width;
	out->image_height = image_height;
	out->own_row0 = r0;
	out->own_row1 = r1;
	out->sub_row0 = std::max(0, r0 - halo);
	out->sub_row1 = std::min(image_height, r1 + halo);
	out->tile_index = tile_index;
	out->tile_count = tile_count;
	return 0;
}

int bcd_b200_tile_begin(bcd_b200_handle* h, const bcd_b200_params* p, const bcd_b200_tile* t)
{
	if(!h || !t) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle / tile");
	if(!h->haveInputs) return fail(BCD_B200_ERR_STATE, "tile_begin called before upload/bind_device");
	int rc = checkParams(p);
	if(rc) return rc;
	if(t->sub_row1 - t->sub_row0 != h->H || t->image_width != h->W)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "resident inputs are %dx%d but the tile needs %dx%d", h->W, h->H,
				t->image_width, t->sub_row1 - t->sub_row0);
	CU(cudaSetDevice(h->device));
	rc = ensureBuffers(h, *p);
	if(rc) return rc;
	memset(&h->stats, 0, sizeof(h->stats));
	memset(&h->timings, 0, sizeof(h->timings));
	h->eventsUsed = 0;
	h->spans.clear();
	h->launches = 0;
	h->tiled = t->tile_count > 1;
	h->imageH = t->image_height;
	h->subRow0 = t->sub_row0; h->subRow1 = t->sub_row1;
	h->ownRow0 = t->own_row0; h->ownRow1 = t->own_row1;
	h->tileIndex = t->tile_index; h->tileCount = t->tile_count;
	h->runParams = *p;
	h->ran = false;
	h->deferSelect = false;
	h->nsEarly = false;
	CU(cudaEventRecord(h->evTotal[0], h->stream));
	h->pyrHistDone = h->pyrColorDone = 0; // the pyramid is built on demand by tile_select / tile_filter
	memset(h->pixcovDone, 0, sizeof(h->pixcovDone));
	return 0;
}

#define TILE_GUARD(h, level) \
	if(!h) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle"); \
	if(level < 0 || level >= h->runParams.nb_scales) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "bad level %d", level); \
	CU(cudaSetDevice(h->device));

int bcd_b200_tile_select(bcd_b200_handle* h, int level) { TILE_GUARD(h, level); return levelSelect(h, level, h->runParams); }
int bcd_b200_tile_mark(bcd_b200_handle* h, int level) { TILE_GUARD(h, level); return levelMark(h, level, h->runParams); }
int bcd_b200_tile_mark_enqueue(bcd_b200_handle* h, int level) { TILE_GUARD(h, level); return levelMarkEnqueue(h, level, h->runParams); }
int bcd_b200_tile_mark_wait(bcd_b200_handle* h, int level) { TILE_GUARD(h, level); return levelMarkWait(h, level); }
int bcd_b200_tile_filter(bcd_b200_handle* h, int level) { TILE_GUARD(h, level); return levelFilterTile(h, level, h->runParams); }
int bcd_b200_tile_finalize(bcd_b200_handle* h, int level) { TILE_GUARD(h, level); return levelFinalize(h, level, h->runParams); }
int bcd_b200_tile_merge(bcd_b200_handle* h, int level)
{
	TILE_GUARD(h, level);
	if(level >= h->runParams.nb_scales - 1) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "the coarsest level is not merged");
	return levelMerge(h, level);
}

namespace
{
// rows (count, bytes per row, base pointer) of the strip; exportSide: rows sent to `side`; import: rows received
struct Strip { char* base; size_t rowBytes; int rows; int exportRow[2]; int importRow[2]; };

int stripOf(bcd_b200_handle* h, int level, int what, Strip& s)
{
	const bcd_b200_params& p = h->runParams;
	const LevelBufs& L = h->levels[level];
	int own0, own1;
	tileOwnedRows(h, level, own0, own1);
	const int bw = p.search_radius + p.patch_radius;
	switch(what)
	{
	case BCD_B200_XFER_ACC_HALO: // partial sums accumulated outside my band go to their owner and are ADDED there
		s.base = (char*)L.acc.p; s.rowBytes = (size_t)L.W * sizeof(float4); s.rows = bw;
		s.exportRow[BCD_B200_SIDE_UP] = own0 - bw; s.exportRow[BCD_B200_SIDE_DOWN] = own1;
		s.importRow[BCD_B200_SIDE_UP] = own0;      s.importRow[BCD_B200_SIDE_DOWN] = own1 - bw;
		break;
	case BCD_B200_XFER_AS_ROWS: // AS bit-rows of my last b owned rows feed the marking of the band below
		s.base = (char*)L.asBits.p; s.rowBytes = (size_t)markingRowWords(L.W) * 4; s.rows = p.search_radius;
		s.exportRow[BCD_B200_SIDE_UP] = own0;                    s.exportRow[BCD_B200_SIDE_DOWN] = own1 - p.search_radius;
		s.importRow[BCD_B200_SIDE_UP] = own0 - p.search_radius;  s.importRow[BCD_B200_SIDE_DOWN] = own1;
		break;
	case BCD_B200_XFER_OUT_ROWS: // 2 output rows each way (downscale + interpolate of the merge reach 2 rows)
		s.base = (char*)L.out.p; s.rowBytes = (size_t)L.W * 3 * sizeof(float); s.rows = 2;
		s.exportRow[BCD_B200_SIDE_UP] = own0;      s.exportRow[BCD_B200_SIDE_DOWN] = own1 - 2;
		s.importRow[BCD_B200_SIDE_UP] = own0 - 2;  s.importRow[BCD_B200_SIDE_DOWN] = own1;
		break;
	default:
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "unknown strip %d", what);
	}
	return 0;
}
}

int bcd_b200_tile_xfer_bytes(bcd_b200_handle* h, int level, int what, size_t* bytes)
{
	TILE_GUARD(h, level);
	if(!bytes) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null output");
	Strip s;
	int rc = stripOf(h, level, what, s);
	if(rc) return rc;
	*bytes = s.rowBytes * s.rows;
	return 0;
}

int bcd_b200_tile_export(bcd_b200_handle* h, int level, int what, int side, void* dev_buf)
{
	TILE_GUARD(h, level);
	if(!dev_buf || side < 0 || side > 1) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "bad export arguments");
	Strip s;
	int rc = stripOf(h, level, what, s);
	if(rc) return rc;
	const int row = s.exportRow[side];
	if(row < 0 || row + s.rows > h->levels[level].H)
		return fail(BCD_B200_ERR_STATE, "export strip rows [%d,%d) outside the resident band (no neighbour on that side?)", row, row + s.rows);
	cudaStream_t xs = what == BCD_B200_XFER_AS_ROWS ? auxOf(h, level) : h->stream; // the marking chain lives on the aux stream
	CU(cudaMemcpyAsync(dev_buf, s.base + (size_t)row * s.rowBytes, s.rowBytes * s.rows, cudaMemcpyDeviceToDevice, xs));
	return 0;
}

int bcd_b200_tile_import(bcd_b200_handle* h, int level, int what, int side, const void* dev_buf)
{
	TILE_GUARD(h, level);
	if(!dev_buf || side < 0 || side > 1) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "bad import arguments");
	Strip s;
	int rc = stripOf(h, level, what, s);
	if(rc) return rc;
	const int row = s.importRow[side];
	if(row < 0 || row + s.rows > h->levels[level].H)
		return fail(BCD_B200_ERR_STATE, "import strip rows [%d,%d) outside the resident band", row, row + s.rows);
	char* dst = s.base + (size_t)row * s.rowBytes;
	if(what == BCD_B200_XFER_ACC_HALO)
		CU(launchAddStrip(reinterpret_cast<float4*>(dst), reinterpret_cast<const float4*>(dev_buf),
				s.rowBytes * s.rows / sizeof(float4), h->stream, &h->launches));
	else if(what == BCD_B200_XFER_AS_ROWS)
	{
		// the bit-rows are cleared by the selection stage on the main stream: order the copy after it
		CU(cudaStreamWaitEvent(auxOf(h, level), h->levels[level].evSel, 0));
		CU(cudaMemcpyAsync(dst, dev_buf, s.rowBytes * s.rows, cudaMemcpyDeviceToDevice, auxOf(h, level)));
	}
	else
		CU(cudaMemcpyAsync(dst, dev_buf, s.rowBytes * s.rows, cudaMemcpyDeviceToDevice, h->stream));
	return 0;
}

int bcd_b200_tile_end(bcd_b200_handle* h)
{
	if(!h) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle");
	CU(cudaSetDevice(h->device));
	CU(cudaEventRecord(h->evTotal[1], h->stream));
	CU(cudaStreamSynchronize(h->stream));
	int rc = collectTimings(h);
	if(rc) return rc;
	h->ran = true;
	return 0;
}

int bcd_b200_tile_download(bcd_b200_handle* h, float* out_rows)
{
	if(!h || !out_rows) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle / output");
	if(!h->ran) return fail(BCD_B200_ERR_STATE, "tile_download called before tile_end");
	CU(cudaSetDevice(h->device));
	int own0, own1;
	tileOwnedRows(h, 0, own0, own1);
	const size_t rowBytes = (size_t)h->W * 3 * sizeof(float);
	CU(cudaMemcpyAsync(out_rows, (char*)h->levels[0].out.p + (size_t)own0 * rowBytes, rowBytes * (own1 - own0), cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return 0;
}

int bcd_b200_spike_removal(bcd_b200_handle* h, float threshold_std_factor, int int_abs)
{
	if(!h) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle");
	if(!h->haveInputs) return fail(BCD_B200_ERR_STATE, "bcd_b200_spike_removal called before upload");
	LevelBufs& L = h->levels[0];
	if(L.color != L.ownColor.as<float>() && L.color != L.altColor.as<float>())
		return fail(BCD_B200_ERR_STATE, "bcd_b200_spike_removal needs uploaded inputs (bound device inputs are never modified)");
	if(h->W < 3 || h->H < 3)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "spike removal needs at least a 3x3 image");
	CU(cudaSetDevice(h->device));
	{ int rcu = finishUpload(h); if(rcu) return rcu; }
	const size_t np = (size_t)h->W * h->H;
	const bool fromOwn = L.color == L.ownColor.as<float>();
	DevBuf& oc = fromOwn ? L.altColor : L.ownColor;
	DevBuf& on = fromOwn ? L.altNs : L.ownNs;
	DevBuf& oh = fromOwn ? L.altHist : L.ownHist;
	DevBuf& ov = fromOwn ? L.altCov : L.ownCov;
	CU(oc.reserve(np * 3 * sizeof(float)));
	CU(on.reserve(np * sizeof(float)));
	CU(oh.reserve(np * h->B * sizeof(float)));
	CU(ov.reserve(np * 6 * sizeof(float)));
	CU(L.fullList.reserve(np * 4)); // scratch: source pixel of every output pixel
	CU(launchSpikeFilter(L.color, L.ns, L.hist, L.cov, h->W, h->H, h->B, threshold_std_factor, int_abs ? 1 : 0,
			L.fullList.as<int>(), oc.as<float>(), on.as<float>(), oh.as<float>(), ov.as<float>(), h->stream, &h->launches));
	CU(cudaStreamSynchronize(h->stream));
	L.color = oc.as<float>(); L.ns = on.as<float>(); L.hist = oh.as<float>(); L.cov = ov.as<float>();
	h->ran = false;
	return 0;
}

int bcd_b200_accumulate_samples(bcd_b200_handle* h, const float* samples, const float* weights, int samples_on_device,
		int32_t width, int32_t height, int32_t spp, int32_t nb_bins, float gamma, float max_value)
{
	if(!h || !samples) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle / samples");
	if(width <= 0 || height <= 0 || spp <= 0 || nb_bins < 2 || 3 * nb_bins > 255)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "bad sample block (%dx%d, %d spp, %d bins)", width, height, spp, nb_bins);
	CU(cudaSetDevice(h->device));
	{ int rcu = finishUpload(h); if(rcu) return rcu; }
	CU(cudaStreamSynchronize(h->stream));
	const int B = 3 * nb_bins;
	int rc = prepareUpload(h, width, height, B); // storage of the level-0 inputs
	if(rc) return rc;
	LevelBufs& L = h->levels[0];
	const size_t np = (size_t)width * height;
	DevBuf stage, stageW; // host samples are copied into a temporary device block
	const float *dS = samples, *dW = weights;
	if(!samples_on_device)
	{
		CU(stage.reserve(np * spp * 3 * sizeof(float)));
		CU(cudaMemcpyAsync(stage.p, samples, np * spp * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
		dS = stage.as<float>();
		if(weights)
		{
			CU(stageW.reserve(np * spp * sizeof(float)));
			CU(cudaMemcpyAsync(stageW.p, weights, np * spp * sizeof(float), cudaMemcpyHostToDevice, h->stream));
			dW = stageW.as<float>();
		}
	}
	cudaError_t e = launchAccumulateSamples(dS, dW, np, spp, nb_bins, gamma, max_value, L.ownColor.as<float>(), L.ownNs.as<float>(),
			L.ownHist.as<float>(), L.ownCov.as<float>(), h->stream, &h->launches);
	cudaError_t e2 = cudaStreamSynchronize(h->stream);
	stage.release();
	stageW.release();
	if(e != cudaSuccess || e2 != cudaSuccess)
		return fail(BCD_B200_ERR_CUDA, "sample accumulation failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
	L.color = L.ownColor.as<float>(); L.ns = L.ownNs.as<float>(); L.hist = L.ownHist.as<float>(); L.cov = L.ownCov.as<float>();
	h->W = width; h->H = height; h->B = B;
	h->haveInputs = true;
	h->ran = false;
	return 0;
}

int bcd_b200_download_inputs(bcd_b200_handle* h, float* color, float* nsamples, float* hist, float* cov)
{
	if(!h) return fail(BCD_B200_ERR_INVALID_ARGUMENT, "null handle");
	if(!h->haveInputs) return fail(BCD_B200_ERR_STATE, "no resident inputs");
	CU(cudaSetDevice(h->device));
	{ int rcu = finishUpload(h); if(rcu) return rcu; }
	const LevelBufs& L = h->levels[0];
	const size_t np = (size_t)h->W * h->H;
	if(color) CU(cudaMemcpyAsync(color, L.color, np * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
	if(nsamples) CU(cudaMemcpyAsync(nsamples, L.ns, np * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
	if(hist) CU(cudaMemcpyAsync(hist, L.hist, np * h->B * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
	if(cov) CU(cudaMemcpyAsync(cov, L.cov, np * 6 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
	CU(cudaStreamSynchronize(h->stream));
	return 0;
}

int bcd_b200_spike_removal_host(float* color, float* nsamples, float* hist, float* cov, int32_t width, int32_t height,
		int32_t hist_depth, float threshold_std_factor, int int_abs, int device)
{
	bcd_b200_images in = {color, nsamples, hist, cov, width, height, hist_depth};
	int rc = checkImages(&in);
	if(rc) return rc;
	bcd_b200_handle* h = nullptr;
	rc = bcd_b200_create(&h, device);
	if(rc) return rc;
	rc = bcd_b200_upload(h, &in);
	if(!rc) rc = bcd_b200_spike_removal(h, threshold_std_factor, int_abs);
	if(!rc) rc = bcd_b200_download_inputs(h, color, nsamples, hist, cov);
	bcd_b200_destroy(h);
	return rc;
}

int bcd_b200_denoise(const bcd_b200_images* in, const bcd_b200_params* p, float* out_rgb, int device,
		bcd_b200_progress_fn progress, void* user)
{
	if(!out_rgb)
		return fail(BCD_B200_ERR_INVALID_ARGUMENT, "Aborting denoising: nullptr for output image");
	int rc = checkImages(in);
	if(rc) return rc;
	rc = checkParams(p);
	if(rc) return rc;
	bcd_b200_handle* h = nullptr;
	rc = bcd_b200_create(&h, device);
	if(rc) return rc;
	rc = bcd_b200_upload_async(h, in); // the host buffers outlive the call, so the transfer can stream under the run
	if(!rc) rc = bcd_b200_run(h, p, progress, user);
	if(!rc) rc = bcd_b200_download(h, out_rgb);
	bcd_b200_destroy(h);
	return rc;
}

} // extern "C"

/*
 * bcd_b200.h — C ABI of the B200-native Bayesian Collaborative Denoiser hot path.
 *
 * This is the drop-in boundary. The reference has no FFI layer: its boundary is the C++ class
 * bcd::IDenoiser (reference include/bcd/core/IDenoiser.h:73-97) with the concrete bcd::Denoiser
 * (include/bcd/core/Denoiser.h:25-44, src/core/Denoiser.cpp:84-212) and bcd::MultiscaleDenoiser
 * (include/bcd/core/MultiscaleDenoiser.h:23-30, src/core/MultiscaleDenoiser.cpp:31-136). The re-authored,
 * source-compatible C++ headers under include/bcd/core/ forward their denoise() to the entry points below, so a
 * host renderer or bcd_cli keeps compiling unchanged and reaches the sm_100a kernels through plain pointers.
 *
 * Conventions
 *   - all images are HWC-interleaved float32, idx = (line*W + col)*depth + k (reference DeepImage.hpp:385-396);
 *     covariance channel order xx,yy,zz,yz,xz,xy (reference CovarianceMatrix.h:18-27);
 *   - host pointers are borrowed for the duration of a call; nothing is freed by the library;
 *   - every function returns 0 on success or a negative bcd_b200_status; bcd_b200_last_error() returns a
 *     thread-local message. The library never calls exit() (the reference's CUDA path does: CudaUtils.h:18-30);
 *   - a handle is not re-entrant; different handles may be used from different threads.
 */
#ifndef BCD_B200_H
#define BCD_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BCD_B200_ABI_VERSION 2

typedef enum bcd_b200_status {
	BCD_B200_OK = 0,
	BCD_B200_ERR_INVALID_ARGUMENT = -1, /* null / empty / size-mismatched inputs (reference Denoiser.cpp:269-346 returns false) */
	BCD_B200_ERR_CUDA = -2,             /* CUDA runtime failure, message in bcd_b200_last_error() */
	BCD_B200_ERR_NO_DEVICE = -3,
	BCD_B200_ERR_UNSUPPORTED = -4,      /* parameter combination the kernels are not built for */
	BCD_B200_ERR_STATE = -5             /* call order violated (e.g. run before upload) */
} bcd_b200_status;

/* Replaces bcd::DenoiserInputs (reference IDenoiser.h:46-59) */
typedef struct bcd_b200_images {
	const float* color;    /* W*H*3 */
	const float* nsamples; /* W*H   */
	const float* hist;     /* W*H*hist_depth */
	const float* cov;      /* W*H*6 (sample covariance) */
	int32_t width;
	int32_t height;
	int32_t hist_depth;    /* 3 * nbOfBins, 60 with the default HistogramParameters */
} bcd_b200_images;

/* Replaces bcd::DenoiserParameters (reference IDenoiser.h:20-44) + MultiscaleDenoiser's nbOfScales */
typedef struct bcd_b200_params {
	float   tau;            /* m_histogramDistanceThreshold (1.0)   */
	int32_t patch_radius;   /* m_patchRadius (1)                    */
	int32_t search_radius;  /* m_searchWindowRadius (6)             */
	float   min_eig;        /* m_minEigenValue (1e-8)               */
	int32_t random_order;   /* m_useRandomPixelOrder: accepted and ignored — the kernels always realise the
	                           sequential raster order (the reference's only deterministic order) */
	float   skip_prob;      /* m_markedPixelsSkippingProbability: 1 = marking strategy (default), 0 = every pixel;
	                           0<m<1 uses a counter-based hash instead of rand() (DenoisingUnit.cpp:168) */
	int32_t nb_scales;      /* 1 = Denoiser, >1 = MultiscaleDenoiser(nb_scales) */
	int32_t reserved;
} bcd_b200_params;

void bcd_b200_default_params(bcd_b200_params* p);

/* Data-dependent work counters of the last run, summed over scales (bench.py prints them). */
typedef struct bcd_b200_stats {
	int64_t interior;    /* patch centres visited */
	int64_t strong;      /* |S(p)| >= d+1 */
	int64_t active;      /* not skipped by the marking strategy */
	int64_t full;        /* two-step Bayesian filter */
	int64_t fallback;    /* denoiseOnlyMainPatch */
	int64_t sum_n_full;  /* sum of |S(p)| over full pixels */
	int64_t kernel_launches;
} bcd_b200_stats;

/* Per-stage device time of the last run in milliseconds (CUDA events on the handle's stream), summed over scales. */
enum { BCD_B200_STAGE_PREPASS = 0, BCD_B200_STAGE_PAIR = 1, BCD_B200_STAGE_SELECT = 2, BCD_B200_STAGE_MARKING = 3,
       BCD_B200_STAGE_FALLBACK = 4, BCD_B200_STAGE_FILTER_STATS = 5, BCD_B200_STAGE_FILTER_EIG = 6,
       BCD_B200_STAGE_FILTER_APPLY = 7, BCD_B200_STAGE_FINALIZE = 8, BCD_B200_STAGE_PYRAMID = 9,
       BCD_B200_STAGE_EXCHANGE = 10, /* row tiles: strips pushed to / awaited from the neighbouring GPUs */
       BCD_B200_STAGE_COUNT = 11 };
typedef struct bcd_b200_timings {
	float stage_ms[BCD_B200_STAGE_COUNT];
	float total_ms;
} bcd_b200_timings;

typedef void (*bcd_b200_progress_fn)(float progress01, void* user);

const char* bcd_b200_last_error(void);
int bcd_b200_abi_version(void);
int bcd_b200_device_count(void);

/* ---- one-shot entry: what bcd::Denoiser::denoise() / bcd::MultiscaleDenoiser::denoise() forward to -------------
 * Replaces reference Denoiser.cpp:84-212 and MultiscaleDenoiser.cpp:31-136. Host buffers in, host buffer out
 * (out_rgb = W*H*3, caller-allocated). Uses device `device` (0 if negative). */
int bcd_b200_denoise(const bcd_b200_images* in, const bcd_b200_params* p, float* out_rgb, int device,
                     bcd_b200_progress_fn progress, void* user);

/* ---- handle API: device buffers (and streams) persist across frames ------------------------------------------- */
typedef struct bcd_b200_handle bcd_b200_handle;

int  bcd_b200_create(bcd_b200_handle** out, int device);
void bcd_b200_destroy(bcd_b200_handle* h);
/* optional: run on a caller-owned stream (cudaStream_t passed as void*); NULL restores the handle's own stream */
int  bcd_b200_set_stream(bcd_b200_handle* h, void* cuda_stream);
/* main + auxiliary stream: the marking stage (and its AS_ROWS strips in row-tile mode) runs on the auxiliary stream so
 * that it overlaps the selection / filters on the main stream; NULL keeps the handle's own high-priority streams (one
 * per pyramid level) */
int  bcd_b200_set_streams(bcd_b200_handle* h, void* main_stream, void* aux_stream);
/* caller-owned marking stream of one pyramid level (NULL: back to the default). Row-tile drivers give every level its
 * own stream so that the marking chains of the levels — each a dependency through all ranks — run side by side */
int  bcd_b200_set_level_aux_stream(bcd_b200_handle* h, int level, void* aux_stream);
/* host -> device copy of the four input images (pinned or pageable host memory) */
int  bcd_b200_upload(bcd_b200_handle* h, const bcd_b200_images* in);
/* streamed variant: returns as soon as the copies are enqueued (sample counts, then the histogram in row bands, then
 * covariance and colour) on the handle's copy stream. The next bcd_b200_run launches the level-0 pair-term kernel band by
 * band behind the arriving histogram rows, so the transfer overlaps the dominant kernel. The host buffers must stay
 * valid and unmodified until that run (or any other call that consumes the inputs) has returned; use pinned memory for a
 * truly asynchronous copy. bcd_b200_denoise uses this path. */
int  bcd_b200_upload_async(bcd_b200_handle* h, const bcd_b200_images* in);
/* page-lock / release a host range the caller owns (thin wrappers over cudaHostRegister / cudaHostUnregister for hosts
 * that do not link the CUDA runtime themselves): registering the image buffers a renderer reuses frame after frame
 * once is what makes the streamed upload asynchronous. Registering an already registered range is not an error. */
int  bcd_b200_host_register(const void* ptr, size_t bytes);
int  bcd_b200_host_unregister(const void* ptr);
/* same, but the pointers in `in` are DEVICE pointers that the handle borrows until the next upload/bind/destroy */
int  bcd_b200_bind_device(bcd_b200_handle* h, const bcd_b200_images* in);
/* whole hot path on the resident inputs; synchronous with respect to the host on return */
int  bcd_b200_run(bcd_b200_handle* h, const bcd_b200_params* p, bcd_b200_progress_fn progress, void* user);
/* device -> host copy of the W*H*3 result */
int  bcd_b200_download(bcd_b200_handle* h, float* out_rgb);
/* device pointer of the W*H*3 result (valid until the next run/destroy) */
int  bcd_b200_result_device_ptr(bcd_b200_handle* h, const float** out_dev);
int  bcd_b200_get_stats(bcd_b200_handle* h, bcd_b200_stats* s);
int  bcd_b200_get_timings(bcd_b200_handle* h, bcd_b200_timings* t);

/* ----  spike-removal prefilter, the step immediately before the path in bcd_cli / the GUI -------------
 * Replaces SpikeRemovalFilter::filter (reference src/core/SpikeRemovalFilter.cpp:18-75): a pixel whose colour is
 * further than factor * stddev from the mean of its clamped 3x3 neighbourhood is replaced — colour, nSamples,
 * histogram and covariance — by the L1-median neighbour. Runs on the handle's resident inputs (they must have been
 * uploaded, not bound) and leaves the filtered images resident as the inputs of the next bcd_b200_run.
 * int_abs: 1 = the overload the reference's unqualified abs() resolves to under GCC/glibc (::abs(int): what the
 * compiled oracle does), 0 = the float overload (MSVC builds of the reference, and the evident intent). */
int  bcd_b200_spike_removal(bcd_b200_handle* h, float threshold_std_factor, int int_abs);
/* host-side convenience: filters the four host images in place (upload, filter, download) */
int  bcd_b200_spike_removal_host(float* color, float* nsamples, float* hist, float* cov, int32_t width, int32_t height,
                                 int32_t hist_depth, float threshold_std_factor, int int_abs, int device);
/* device -> host copy of the resident level-0 inputs (e.g. after bcd_b200_spike_removal) */
int  bcd_b200_download_inputs(bcd_b200_handle* h, float* color, float* nsamples, float* hist, float* cov);

/* ---- (device side) sample accumulation on the GPU ------------------------------------------------------
 * Replaces SamplesAccumulator::addSample + computeSampleStatistics (reference src/core/SamplesAccumulator.cpp:44-107,
 * 110-141) for a renderer whose samples already live in device memory (or in one host block): samples are
 * [height][width][spp][3] floats, weights (may be NULL: all 1) [height][width][spp]. The statistics — mean colour, sample
 * count, 3*nb_bins histogram (gamma / max_value as in HistogramParameters, SamplesAccumulator.h:20-30), bias-corrected
 * covariance — become the handle's resident inputs, ready for bcd_b200_spike_removal / bcd_b200_run; nothing
 * round-trips through the host. samples_on_device: 1 = device pointers, 0 = host pointers (copied first).
 * Mean, count and covariance are the reference's numbers bit for bit; histogram bins may differ in the last place
 * (device powf vs glibc powf in the gamma curve). */
int  bcd_b200_accumulate_samples(bcd_b200_handle* h, const float* samples, const float* weights, int samples_on_device,
                                 int32_t width, int32_t height, int32_t spp, int32_t nb_bins, float gamma, float max_value);

/* ---- row-tile API (multi-GPU: one handle per GPU, one process per GPU) ---------------------------------------------
 * The frame is cut into horizontal bands. Tile t of n produces the output rows [own_row0, own_row1) and keeps the
 * input rows [sub_row0, sub_row1) resident (its band plus (b+w)*2^(s-1) halo rows, so that every pyramid level has
 * its b+w halo rows and the local pyramid equals the global one restricted to the band). The caller uploads/binds
 * exactly those sub_row1 - sub_row0 rows as the handle's inputs, then drives the stages level by level, coarsest
 * first, exchanging the small strips named below with the neighbouring ranks between stages (bcd_b200/tiles.py does
 * this with torch.distributed send/recv):
 *
 *   tile_begin                      pyramid build (local)
 *   for level = 0 .. s-1:           (finest first: its marking chain is the longest)
 *     tile_select(level)            K0 + K1 on the band (+ halo rows)
 *     [m > 0] on the level's marking stream: import AS_ROWS from the upper neighbour -> tile_mark_enqueue(level) ->
 *             export AS_ROWS to the lower neighbour (the raster-order marking is a chain through the ranks)
 *   for level = 0, s-1, .., 1:
 *     tile_mark_wait(level)         host: work counters of the level
 *     tile_filter(level)            fallback + two-step filter into the local accumulator
 *     export ACC_HALO (up, down) -> neighbours -> import (adds)      the only data-path exchange: (b+w) rows x W x 16 B
 *     tile_finalize(level)
 *     export OUT_ROWS (up, down) -> neighbours -> import             2 rows of the level's output
 *   for level = s-2 .. 0:
 *     tile_merge(level), then OUT_ROWS again (the merged rows feed the next finer level)
 *   tile_download                   owned rows of level 0
 * (any order that respects select -> mark -> filter -> finalize per level and finalize(k), merge(k+1) -> merge(k) works)
 *
 * No collective is needed: replaces nothing in the reference (it has no multi-GPU path). */
typedef struct bcd_b200_tile {
	int32_t image_width, image_height; /* full frame */
	int32_t own_row0, own_row1;        /* output rows produced by this tile (frame coordinates) */
	int32_t sub_row0, sub_row1;        /* input rows resident in the handle (frame coordinates) */
	int32_t tile_index, tile_count;
} bcd_b200_tile;

enum { BCD_B200_XFER_ACC_HALO = 0, BCD_B200_XFER_AS_ROWS = 1, BCD_B200_XFER_OUT_ROWS = 2 };
enum { BCD_B200_SIDE_UP = 0, BCD_B200_SIDE_DOWN = 1 };

/* pure host arithmetic (no device needed): rows of tile `tile_index` of `tile_count` for an image_height-row frame */
int  bcd_b200_tile_plan(int image_width, int image_height, int tile_index, int tile_count, const bcd_b200_params* p,
                        bcd_b200_tile* out);
int  bcd_b200_tile_begin(bcd_b200_handle* h, const bcd_b200_params* p, const bcd_b200_tile* t);
int  bcd_b200_tile_select(bcd_b200_handle* h, int level);
int  bcd_b200_tile_mark(bcd_b200_handle* h, int level);          /* = tile_mark_enqueue + tile_mark_wait */
/* asynchronous halves of tile_mark: enqueue K2 + list builder + counter read-back on the level's marking stream;
 * wait for the counters on the host (needed before tile_filter). Between the two the caller may already export the
 * AS_ROWS strip for the lower neighbour (stream-ordered) and go on with other levels. */
int  bcd_b200_tile_mark_enqueue(bcd_b200_handle* h, int level);
int  bcd_b200_tile_mark_wait(bcd_b200_handle* h, int level);
int  bcd_b200_tile_filter(bcd_b200_handle* h, int level);
int  bcd_b200_tile_finalize(bcd_b200_handle* h, int level);
int  bcd_b200_tile_merge(bcd_b200_handle* h, int level);
/* size in bytes of the strip `what` of `level` (identical on both sides of a tile boundary) */
int  bcd_b200_tile_xfer_bytes(bcd_b200_handle* h, int level, int what, size_t* bytes);
/* copy the strip facing neighbour `side` into dev_buf (device memory, async on the handle's stream) */
int  bcd_b200_tile_export(bcd_b200_handle* h, int level, int what, int side, void* dev_buf);
/* take the strip received from neighbour `side`: ACC_HALO is added into the owned rows, AS_ROWS / OUT_ROWS are
 * copied into the halo rows */
int  bcd_b200_tile_import(bcd_b200_handle* h, int level, int what, int side, const void* dev_buf);
int  bcd_b200_tile_end(bcd_b200_handle* h);                 /* synchronises, collects stats/timings */
int  bcd_b200_tile_download(bcd_b200_handle* h, float* out_rows /* (own_row1-own_row0) * W * 3 */);

/* ---- peer mode of the row-tile API: the tiles exchange their strips through peer memory (NVLink) -----------------------
 * The strips of the table above travel without any host or NCCL hop: every tile owns an "arena" (one device allocation)
 * into which its neighbours' kernels push their rows, followed by an epoch flag; the consuming kernels wait for the flag
 * on the device. The systolic marking kernel (b = 6) even chains its launches on neighbouring GPUs word by word — the
 * raster-order dependency crosses the tile boundary at the pace of the wavefront instead of tile by tile.
 *   tile_setup    sizes the buffers for tile t of a W x H frame (hist_depth channels) and returns the token of the arena;
 *   tile_connect  makes the arena of the neighbour on `side` writable from this GPU: same process -> peer access,
 *                 other process -> CUDA IPC (the 96 token bytes travel by any means, e.g. an all_gather at start-up);
 *   tile_run      the whole schedule of the table above on the resident rows of the tile (upload / bind them first), all
 *                 tiles of the frame concurrently (one thread or process each, every tile the same number of runs since
 *                 its setup; all tiles must have finished tile_setup before the first one runs). Returns after this
 *                 tile's rows are done; tile_download fetches them.
 * Every wait is bounded (~4 s): a lost neighbour makes tile_run fail with BCD_B200_ERR_CUDA, it never hangs the GPU. */
typedef struct bcd_b200_peer_token { unsigned char bytes[96]; } bcd_b200_peer_token;
int  bcd_b200_tile_setup(bcd_b200_handle* h, const bcd_b200_params* p, const bcd_b200_tile* t, int hist_depth,
                         bcd_b200_peer_token* token_out);
int  bcd_b200_tile_connect(bcd_b200_handle* h, int side, const bcd_b200_peer_token* neighbour /* NULL: disconnect */);
int  bcd_b200_tile_run(bcd_b200_handle* h, const bcd_b200_params* p /* NULL: the parameters of tile_setup; same radii / scales */);

/* ---- group API: the GPUs of one box behind one call (what bcd::Denoiser / bcd::MultiscaleDenoiser hold) ---------------
 * A group keeps one handle and one host thread per row tile, every device allocation, stream and peer link across
 * frames, and page-locks the caller's buffers the first time it sees them (a renderer hands over the same DeepImage
 * storage frame after frame; reference call sites src/cli/main.cpp:457-468, src/gui/GuiWindow.cpp:345-376).
 * The frame is cut into tiles_per_gpu x GPUs row tiles dealt out CYCLICALLY (tile j on GPU j mod G): with several tiles
 * per GPU the raster-order marking chain pipelines behind the selection instead of starting when whole 1/G-ths of the
 * frame are selected. tiles_per_gpu = 0 picks it from the frame height; the tile count shrinks when
 * the frame is too small. device_ids may repeat a device (tests on a one-GPU box).
 * group_denoise: host buffers in, host buffer out, like bcd_b200_denoise. pin_host: 1 = cudaHostRegister the five host
 * ranges (cached per pointer), 0 = leave them pageable. */
typedef struct bcd_b200_group bcd_b200_group;
int  bcd_b200_group_create(bcd_b200_group** out, const int32_t* device_ids, int nb_gpus);
void bcd_b200_group_destroy(bcd_b200_group* g);
int  bcd_b200_group_set_pinning(bcd_b200_group* g, int pin_host);
int  bcd_b200_group_set_tiles_per_gpu(bcd_b200_group* g, int tiles_per_gpu);
int  bcd_b200_group_denoise(bcd_b200_group* g, const bcd_b200_images* in, const bcd_b200_params* p, float* out_rgb,
                            bcd_b200_progress_fn progress, void* user);
/* counters summed over the (local) tiles; timings: per-stage sums over the tiles, total_ms = the slowest tile */
int  bcd_b200_group_get_stats(bcd_b200_group* g, bcd_b200_stats* s);
int  bcd_b200_group_get_timings(bcd_b200_group* g, bcd_b200_timings* t);
int  bcd_b200_group_tiles(bcd_b200_group* g);       /* tiles the frame is cut into (all processes) */
int  bcd_b200_group_local_tiles(bcd_b200_group* g); /* tiles held by this process */
/* one-shot multi-GPU variant of bcd_b200_denoise (creates and destroys a group) */
int  bcd_b200_denoise_multi(const bcd_b200_images* in, const bcd_b200_params* p, float* out_rgb, const int32_t* device_ids,
                            int nb_gpus, bcd_b200_progress_fn progress, void* user);

/* The same group when the GPUs of the box are spread over several processes (one process per GPU under torchrun /
 * MPI: rank r of `world`, each with nb_gpus local devices). No process sees the whole frame: it asks for its tiles
 * (group_setup, group_tile_info), hands over their rows and drives them; the arenas of the neighbouring tiles are
 * linked through CUDA IPC from the tokens every process contributes (e.g. one all_gather at start-up, which is also the
 * barrier between setup and the first run).
 *   group_setup         plans the tiles of a W x H frame, sets up the local ones (warm-up pass included), returns their
 *                       tokens and frame-wide tile indices (capacity = room in both arrays);
 *   group_connect       tokens_by_tile[j] = token of tile j for every tile of the frame (NULL in a one-process group);
 *   group_set_tile_buffers  host rows of a local tile (sub_row0 .. sub_row1) and where its output rows go: used by
 *                       group_run(from_host = 1) = streamed upload + run + download per tile, all tiles concurrently;
 *   group_upload_tile   synchronous upload of a local tile's rows; group_run(from_host = 0) then runs on the resident rows;
 *   group_download_tile owned output rows of a local tile after a run. */
int  bcd_b200_group_create_ranked(bcd_b200_group** out, const int32_t* device_ids, int nb_gpus, int rank, int world);
int  bcd_b200_group_setup(bcd_b200_group* g, int width, int height, int hist_depth, const bcd_b200_params* p,
                          bcd_b200_peer_token* tokens_out, int32_t* tile_indices_out, int capacity);
int  bcd_b200_group_connect(bcd_b200_group* g, const bcd_b200_peer_token* tokens_by_tile, int nb_tiles);
int  bcd_b200_group_tile_info(bcd_b200_group* g, int local_tile, bcd_b200_tile* out);
int  bcd_b200_group_set_tile_buffers(bcd_b200_group* g, int local_tile, const bcd_b200_images* rows_in, float* rows_out);
int  bcd_b200_group_upload_tile(bcd_b200_group* g, int local_tile, const bcd_b200_images* rows_in);
int  bcd_b200_group_run(bcd_b200_group* g, const bcd_b200_params* p, int from_host);
int  bcd_b200_group_download_tile(bcd_b200_group* g, int local_tile, float* rows_out);

/* ---- stage-level read-backs used by the parity tests (level 0 of the last run) ---------------------------------
 * masks: W*H*words uint32, bit index (dy+b)*(2b+1)+(dx+b) — the ordered similar set S(p) of
 * DenoisingUnit::selectSimilarPatches (reference DenoisingUnit.cpp:196-219); counts: W*H int32;
 * active: W*H uint8 (marking strategy outcome, DenoisingUnit.cpp:164-173,690). */
int  bcd_b200_mask_words(int search_radius);
int  bcd_b200_debug_read_selection(bcd_b200_handle* h, uint32_t* masks, int32_t* counts);
int  bcd_b200_debug_read_active(bcd_b200_handle* h, uint8_t* active);
/* after a single-scale run: the pair terms of the selection stage (DenoisingUnit.cpp:360-386) for every pixel and
 * every offset delta_j at or after the window centre, planes [J][H][W], J = ((2b+1)^2+1)/2; entries whose partner
 * pixel lies outside the image are undefined */
int  bcd_b200_debug_read_pair_terms(bcd_b200_handle* h, float* sums, uint8_t* counts);

#ifdef __cplusplus
}
#endif
#endif /* BCD_B200_H */

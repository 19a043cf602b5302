/*
 * hopperrender_cuda.h — C ABI of libhopperrender_cuda.so, the B200 (sm_100a) replacement
 * for the OpenCL layer under HopperRender's optical-flow-calc interface.
 *
 * Plain C, plain pointers and sizes. No OpenCL, no torch, no C++ types. Every function
 * returns 0 on success and non-zero on failure (the reference's convention:
 * video/filter/HopperRender/opticalFlowCalc.c:11-15 `CHECK_ERROR` -> return 1);
 * hr_last_error() gives the text that the reference would have printed to stderr.
 *
 * Each entry point names the reference interface it replaces. "HR/" abbreviates
 * /root/reference/video/filter/HopperRender/.
 *
 * Threading: like the reference (one in-order queue, HR/opticalFlowCalc.c:389-390, all
 * calls from the thread that runs the filter graph, filters/filter_internal.h:123-124)
 * a context is driven from one thread at a time.
 */
#ifndef HOPPERRENDER_CUDA_H
#define HOPPERRENDER_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HR_ABI_VERSION 1

/* Pixel formats. NV12 is what the reference negotiates (HR/vf_HopperRender.c:668).
 * P010 (16-bit little-endian, 10 significant bits MSB-aligned, video/img_format.h:237)
 * has no reference implementation; its semantics are defined in DESIGN.md §P010. */
#define HR_PIXFMT_NV12 0
#define HR_PIXFMT_P010 1

/* Frame output modes = enum FrameOutput, HR/vf_HopperRender.c:21 */
#define HR_MODE_WARPED_12 0
#define HR_MODE_WARPED_21 1
#define HR_MODE_BLENDED 2
#define HR_MODE_HSV_FLOW 3
#define HR_MODE_GREY_FLOW 4
#define HR_MODE_SIDE_BY_SIDE_1 5
#define HR_MODE_SIDE_BY_SIDE_2 6

/* Limits of this implementation (the reference build uses 5..16, HR/config.h:6-7). */
#define HR_MIN_SEARCH_RADIUS 2
#define HR_MAX_SEARCH_RADIUS 32

typedef struct HrContext HrContext; /* opaque; replaces the cl_* members of struct OpticalFlowCalc, HR/opticalFlowCalc.h:30-64 */

typedef struct HrInfo {
    int abiVersion;
    int device;            /* CUDA device ordinal                                   */
    int frameHeight;       /* HR/opticalFlowCalc.h:14                                */
    int frameWidth;        /* stride in samples, HR/opticalFlowCalc.h:13             */
    int actualWidth;       /* HR/opticalFlowCalc.h:15                                */
    int pixfmt;
    int resScalar;         /* opticalFlowResScalar, HR/opticalFlowCalc.c:331-334     */
    int lowWidth;          /* opticalFlowFrameWidth, HR/opticalFlowCalc.c:335        */
    int lowHeight;         /* opticalFlowFrameHeight, HR/opticalFlowCalc.c:336       */
    int firstWindow;       /* HR/opticalFlowCalc.c:133-143                           */
    int iterations;        /* HR/opticalFlowCalc.c:146-149                           */
    int searchCtas;        /* CTAs of the persistent search kernel                   */
    int smCount;
    size_t frameBytes;     /* 1.5 * H * stride * bytesPerSample                      */
    size_t deviceBytes;    /* total HBM held by the context                          */
} HrInfo;

/* ---- lifetime: replaces initOpticalFlowCalc / freeOFC, HR/opticalFlowCalc.c:323-442, :236-253.
 * device < 0 selects the current CUDA device. No JIT: kernels are embedded for sm_100a. */
int hr_create(HrContext **out, int frameHeight, int frameWidth /*stride, samples*/, int actualWidth, int pixfmt, int device);
int hr_destroy(HrContext *ctx);
int hr_get_info(const HrContext *ctx, HrInfo *info);
const char *hr_last_error(const HrContext *ctx); /* ctx may be NULL: error of the last failed hr_create */

/* All work of a context is enqueued on one stream (the reference's single in-order queue).
 * Pass a cudaStream_t to share the caller's stream (NULL restores the context's own). */
int hr_set_stream(HrContext *ctx, void *cudaStream);
int hr_synchronize(HrContext *ctx);

/* Pipelined mode (no reference counterpart: the reference has one in-order queue and blocks
 * once per source frame and once per output frame, HR/opticalFlowCalc.c:98-100,112-114). When enabled, the
 * enqueue-only calls (hr_update_frame_device with borrow = 1, hr_calc_flow with seconds = NULL, hr_warp into
 * planes set by hr_set_output_device, hr_step_device) place independent work of consecutive frame pairs on
 * internal streams: the packed copy of frame k is built while the search of pair k runs, the warps of one pair
 * run side by side, the search of pair k+1 runs beside the warps of pair k (two flow buffers). Results are
 * bit-identical to the serial mode. Outputs in caller-owned planes are complete after hr_synchronize, or — for
 * work that the caller enqueues on the context's stream — after hr_pipeline_join; the same two calls are what
 * orders a caller's re-use of borrowed input planes after the warps that still read them. Calls that block or touch
 * host memory (hr_update_frame, hr_download, the taps) behave as before, except that two of them start work ahead
 * of the call that will ask for it: hr_update_frame launches the search of the new pair with the knobs of the
 * previous hr_calc_flow, hr_download warps the frame of the next blending scalar of the pacing (last scalar + last
 * increment) into a second internal frame while its copy runs. A following call whose arguments match finds the
 * work under way or done, any other call launches its own — the results are the same bits either way
 * (HR_AHEAD=0 in the environment turns the guessing off). Ignored while bands are configured. */
int hr_set_pipeline(HrContext *ctx, int enable);
int hr_pipeline_join(HrContext *ctx);
/* One source frame of a device-resident stream in one call (offline / batch interpolation):
 * hr_update_frame_device + hr_calc_flow + nWarps x hr_warp, warp i with blendingScalars[i] into the caller-owned
 * planes outY[i], outUV[i]. The first frame of a stream only primes the pair. Enqueue-only. */
int hr_step_device(HrContext *ctx, const void *dYPlane, const void *dUvPlane, int borrow, int searchRadius, int deltaScalar,
                   int neighborBiasScalar, int nWarps, const float *blendingScalars, int frameOutputMode, float blackLevel,
                   float whiteLevel, void *const *outY, void *const *outUV);

/* nSteps consecutive source frames in one call: hr_step_device for dY[i], dUV[i] with nWarps[i] warps each, their
 * blending scalars and output planes taken in order from the flat arrays. Enqueue-only. */
int hr_steps_device(HrContext *ctx, int nSteps, const void *const *dYPlanes, const void *const *dUvPlanes, int borrow, int searchRadius,
                    int deltaScalar, int neighborBiasScalar, const int *nWarps, const float *blendingScalars, int frameOutputMode,
                    float blackLevel, float whiteLevel, void *const *outY, void *const *outUV);

/* ---- updateFrame, HR/opticalFlowCalc.c:96-107: blocking upload of the Y plane
 * (frameHeight*stride samples) and the interleaved UV plane (frameHeight/2*stride samples) from
 * HOST memory into the older frame slot, then swap, so that slot 1 is the newest frame. Also
 * builds the search's phase-planar packed copy of the frame. */
int hr_update_frame(HrContext *ctx, const void *yPlane, const void *uvPlane);

/* Same, for a frame already in device memory (NVDEC / IMGFMT_CUDA input).
 * borrow = 0: the planes are copied device-to-device into the context's slot.
 * borrow = 1: no copy; the caller keeps both planes valid and unchanged until two further
 *             hr_update_frame* calls have been made. Enqueue-only (does not block). */
int hr_update_frame_device(HrContext *ctx, const void *dYPlane, const void *dUvPlane, int borrow);

/* ---- calculateOpticalFlow, HR/opticalFlowCalc.c:126-203 (K1-K4: delta-sum search over
 * `searchRadius` candidate layers, lowest-layer selection, offset update, 8x8 flow blur).
 * searchRadius / deltaScalar / neighborBiasScalar are the struct fields the reference re-binds
 * on every call (HR/opticalFlowCalc.c:130,166-170). If `seconds` is non-NULL the call blocks
 * and stores the device time from the last hr_update_frame* to the end of the blur
 * (= ofcCalcTime, HR/opticalFlowCalc.c:196-201); if NULL it only enqueues. */
int hr_calc_flow(HrContext *ctx, int searchRadius, int deltaScalar, int neighborBiasScalar, double *seconds);

/* ---- warpFrames, HR/opticalFlowCalc.c:205-234 (K5: flip lookup + bidirectional warp + blend +
 * output levels + output modes, luma and chroma in one launch). Fails for t > 1 (:209-212).
 * black/white are outputBlackLevel/outputWhiteLevel (HR/opticalFlowCalc.c:225-226). Enqueue-only. */
int hr_warp(HrContext *ctx, float blendingScalar, int frameOutputMode, float blackLevel, float whiteLevel);

/* The same K5 for several output frames of the current pair at once (what the filter asks for one after the other,
 * HR/vf_HopperRender.c:357-405: 2 or 3 per source frame at 24->60, up to 6 at 24->144): warp i with blendingScalars[i]
 * into the caller-owned DEVICE planes outY[i] / outUV[i], one kernel launch for up to 8 of them. Enqueue-only. */
int hr_warp_batch(HrContext *ctx, int nWarps, const float *blendingScalars, int frameOutputMode, float blackLevel, float whiteLevel,
                  void *const *outY, void *const *outUV);

/* ---- downloadFrame, HR/opticalFlowCalc.c:109-124: blocking copy of the output frame to HOST
 * planes; `seconds` (may be NULL) receives warp-start -> download-end (= warpCalcTime). */
int hr_download(HrContext *ctx, void *yPlane, void *uvPlane, double *seconds);

/* Zero-copy hand-off (frames that arrive as IMGFMT_CUDA, video/img_format.h:274, and leave as such):
 * hr_update_frame_device(borrow = 1) takes the decoder's device planes, hr_set_output_device + hr_warp write into the
 * device image that travels on, hr_finish waits until everything enqueued is complete and reports the time since the
 * start of the most recent hr_warp (the warpCalcTime of a frame that is never downloaded, HR/opticalFlowCalc.c:117-122).
 * hr_debug_host_transfer_bytes: bytes the interface calls have moved across PCIe in this process so far (uploads of
 * hr_update_frame / hr_band_upload, downloads of hr_download / hr_band_download) — a zero-copy integration leaves both
 * counters where they were. */
int hr_finish(HrContext *ctx, double *seconds);
int hr_debug_host_transfer_bytes(unsigned long long *h2d, unsigned long long *d2h);

/* Page-locked host memory for the frames a filter allocates itself (SURVEY.md §8f N1 "pinned staging"; no reference
 * counterpart — the reference's output images come from mp_image_pool's default allocator, HR/vf_HopperRender.c:385,699,
 * i.e. pageable memory the driver has to stage). A filter hands hr_host_alloc / hr_host_free to
 * mp_image_pool_set_allocator (video/mp_image_pool.h:20-22, through mp_image_from_buffer, video/mp_image.h:139-142:
 * patches/0004): hr_download then copies straight into the image, 2.5 of the 3.5 transfers per source frame at 24->60.
 * No context is needed (the memory is portable across devices); hr_host_free may be called from any thread.
 * hr_debug_host_pointer_kind: 0 = pageable host memory (goes through the pinned ring of hr_staging.h), 1 = page-locked
 * host memory, 2 = device or managed memory. */
int hr_host_alloc(void **out, size_t bytes);
int hr_host_free(void *p);
int hr_debug_host_pointer_kind(const void *p);

/* Device-side view of the output frame (zero-copy hand-off to a CUDA VO; N2). Valid until the
 * next hr_warp on this context; ordered on the context's stream. */
int hr_get_output_device(HrContext *ctx, void **dYPlane, void **dUvPlane);
/* Redirect the output of subsequent hr_warp calls into caller-owned device planes
 * (NULL, NULL restores the internal buffer). */
int hr_set_output_device(HrContext *ctx, void *dYPlane, void *dUvPlane);

/* ---- spatial bands (no reference counterpart: the reference drives one device,
 * HR/opticalFlowCalc.c:279-305). N contexts, one per GPU, created for the FULL frame geometry; context `rank` owns the
 * rows [row0[rank], row1[rank]) of every frame — a whole number of lattice tile rows (32 << resScalar frame rows; the
 * last band ends with the frame). Per source frame every GPU
 *   - uploads only its band (hr_band_upload),
 *   - fetches only its HALO from the neighbouring bands' frame slots over NVLink P2P (hr_band_gather): the rows its
 *     search and its warp can reach beyond the band, i.e. the largest accumulated offset of the configured maximum search
 *     radius (+-32 rows at radius 5, -512/+392 at 16, HR/Kernels/calcDeltaSumsKernel.cl:68-72), and packs what it holds,
 *   - searches ONLY ITS OWN lattice tiles (hr_calc_flow): what tiles of different GPUs owe one another — the tile totals
 *     of windows that span tiles, the window-table entries at a band's edge, the blurred flow — is stored straight into
 *     the peers' memory by the search kernel itself (epoch-tagged 64-bit words in an "exchange arena" with the same layout
 *     on every GPU; consumers poll their own memory only), so the flow is bit-identical to the single-GPU one and whole on
 *     every GPU when the launch ends,
 *   - warps and downloads only its band (hr_warp, hr_band_download).
 * No NCCL, no host synchronisation on the data path: hand-offs between GPUs are device-side (mailbox counters written and
 * polled by one-thread kernels in stream order; tagged words inside the search). Every rank issues the same sequence of
 * calls; a driver of several ranks in one process calls hr_band_upload on all of them before the first hr_band_gather, and
 * hr_calc_flow on all of them before it waits for any. The searches of a group wait for one another ON THE DEVICE: the
 * contexts of a group must sit on different GPUs (several waiting kernels on one GPU are not guaranteed to run at the
 * same time); a group of one band is allowed anywhere. The pipelined mode is off while bands are configured. */
#define HR_MAX_BANDS 16
#define HR_IPC_HANDLE_BYTES 64
int hr_band_configure(HrContext *ctx, int rank, int world, const int *row0, const int *row1);
/* largest search radius the group will use (sizes the halo; default 16 = MAX_SEARCH_RADIUS, HR/config.h:7); the same
 * on every rank, before the first frame */
int hr_band_set_max_radius(HrContext *ctx, int searchRadius);
/* what a peer must map: frame slot 0, frame slot 1, exchange arena — as pointers (same process) ... */
int hr_band_local_pointers(HrContext *ctx, void **slot0, void **slot1, void **arena);
/* ... or as three CUDA IPC handles (another process) */
int hr_band_export_ipc(HrContext *ctx, unsigned char *handles /* 3 * HR_IPC_HANDLE_BYTES */);
int hr_band_open_ipc(HrContext *ctx, const unsigned char *handles, void **slot0, void **slot1, void **arena);
/* peerDevice >= 0: same-process peer on that device (peer access is enabled); < 0: pointers from IPC */
int hr_band_connect(HrContext *ctx, int peerRank, int peerDevice, void *slot0, void *slot1, void *arena);
/* updateFrame, banded, in two phases (see above). yBand / uvBand: first luma / chroma row of the band. */
int hr_band_upload(HrContext *ctx, const void *yBand, const void *uvBand, int sourceIsDevice);
int hr_band_gather(HrContext *ctx, int blocking);
/* downloadFrame of the band's rows only */
int hr_band_download(HrContext *ctx, void *yBand, void *uvBand, double *seconds);
/* rows [*lo, *hi) of every frame this rank holds (band + halo), bytes fetched from peers over NVLink so far */
int hr_band_get_halo(const HrContext *ctx, int *lo, int *hi, unsigned long long *p2pBytes);

/* ---- parity taps (tests only; nothing on the playback path calls them). Blocking.
 * raw / blurred: int16 [2][lowHeight][lowWidth], X plane then Y plane = offsetArray /
 * blurredOffsetArray (HR/opticalFlowCalc.c:397-398). Either pointer may be NULL. */
int hr_get_offsets(HrContext *ctx, int16_t *raw, int16_t *blurred);
int hr_set_blurred_offsets(HrContext *ctx, const int16_t *blurred);
/* Optional blur-only entry: north-star's "blur flow" step on caller data (K4 alone). */
int hr_blur_flow(HrContext *ctx, const int16_t *rawHost, int16_t *blurredHost);
/* Winning layer of every lattice point's window after search step `step` (0 .. 2*iterations-1):
 * uint8 [lowHeight][lowWidth] (= lowestLayerArray at the window representatives,
 * HR/Kernels/determineLowestLayerKernel.cl:10-20). Needs hr_set_trace(ctx, 1) before the flow. */
int hr_set_trace(HrContext *ctx, int enable);
int hr_get_step_layers(HrContext *ctx, int step, uint8_t *layers);

/* Parity tap: out[i] = MUFU.RCP((float)i) for 0 <= i < n, the reciprocal that `/` (div.full.f32) of the
 * reference's level mapping (HR/Kernels/warpFrameKernel.cl:1-7) multiplies by on an NVIDIA device.
 * The CPU oracle's NVIDIA-OpenCL arithmetic reads it from tests/golden/mufu_rcp_table.npy. */
int hr_debug_rcp_table(float *out, int n);

/* Developer tap: the INT-pipe roofline denominator of the search. A kernel that issues nothing but
 * independent vabsdiff4...add chains (SASS VABSDIFF4.U8.ACC, the one instruction a candidate evaluation of
 * HR/Kernels/calcDeltaSumsKernel.cl:96-99 costs here) at full occupancy on `device` (< 0: current):
 * evalsPerSecond = thread-level packed SADs per second over the whole GPU (CUDA events), warpInstrPerClkPerSm = the
 * same as warp instructions per SM clock. Either pointer may be NULL. Blocking. */
int hr_debug_int_peak(int device, double *evalsPerSecond, double *warpInstrPerClkPerSm);
/* Developer tap (host only, no context, no GPU): the predictor of the filter's next blending scalar
 * (csrc/hr_pacing_predict.h; the pacing arithmetic of vf_HopperRender.c:371-374,481 seen through warpFrames' float
 * argument) fed with n scalars in call order. predicted[i] / nextFrame[i] = its guess for scalars[i], made after
 * scalars[0 .. i-1], and whether it expected a new source frame first; have[i] = 0: no guess. */
int hr_debug_predict_pacing(const float *scalars, int n, float *predicted, int *nextFrame, int *have);

/* Developer tap: SM-clock stamps taken by thread 0 of every search CTA at fixed points of the launch
 * (step start, layers reduced, window published / complete, level done, search done, blur done).
 * stamps: int64 [min(maxCtas, searchCtas)][HR_TIMELINE_SLOTS]; unused slots are 0. Blocking. */
#define HR_TIMELINE_SLOTS 128
int hr_set_timeline(HrContext *ctx, int enable);
int hr_get_timeline(HrContext *ctx, long long *stamps, int maxCtas);
/* With HR_TIMELINE_HOST=1 in the environment the stamps live in mapped host memory and can be read while a launch is
 * still running (or is not coming back): the same array, copied without any synchronisation. */
int hr_debug_peek_timeline(HrContext *ctx, long long *stamps, int maxCtas);

/* Developer knob: which generation of the search kernel the following hr_calc_flow launches use. 0 (default): chosen
 * per launch — csrc/hr_search3.cuh (four lattice points per thread, three CTAs per SM) while the search of the previous
 * pair is still running (pipelined, device-resident streams: up to three launches share the SMs), csrc/hr_search.cuh when the
 * launch has the GPU to itself (shortest single launch); 3 / 2: csrc/hr_search3.cuh / csrc/hr_search2.cuh (and its
 * TMA-staged variant) for radii 5..16 on lattices of at most one tile per SM outside band groups, csrc/hr_search.cuh
 * for everything else; 1: csrc/hr_search.cuh always. All write the same tables, totals and offsets (same bits). */
int hr_debug_set_search_generation(HrContext *ctx, int generation);
/* ... and the generation the most recent search launch of the context actually ran (0: none yet). */
int hr_debug_last_search_generation(const HrContext *ctx);
/* Generation 2 has a variant that stages the tile's neighbourhood of all 16 phase planes of the packed frame in shared
 * memory by TMA (resolution scalar 2 — 1080p, 720p —, radius 5..8); on by default where it applies. enable = 0: the
 * following launches read their samples from global memory. hr_debug_last_search_staged: what the last launch ran. */
int hr_debug_set_search_staged(HrContext *ctx, int enable);
int hr_debug_last_search_staged(const HrContext *ctx);

/* Record CUDA events around every kernel launch (off by default; adds two event records per launch). */
int hr_set_profiling(HrContext *ctx, int enable);
/* Device time of the most recent search-kernel / warp-kernel / pack-kernel launch, seconds,
 * measured with CUDA events on the context's stream. Blocking. Any pointer may be NULL. */
int hr_get_kernel_times(HrContext *ctx, double *searchSeconds, double *warpSeconds, double *packSeconds);
/* Kernel launches issued by this context so far (bench.py's gpu_launches). */
uint64_t hr_get_launch_count(const HrContext *ctx);

int hr_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HOPPERRENDER_CUDA_H */

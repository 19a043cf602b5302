/*
 * hr_cuda.cu — C-ABI host layer of libhopperrender_cuda.so (include/hopperrender_cuda.h).
 *
 * Replaces the OpenCL host of the reference (video/filter/HopperRender/opticalFlowCalc.c):
 * context/buffers (:323-442), upload + swap (:96-107), the 66-command flow loop (:126-203,
 * here: one cooperative launch), the warp launches (:205-234, here: one launch for luma and
 * chroma), the download (:109-124). There is no CPU fallback: every entry point fails when CUDA
 * fails.
 */
#include "../../include/hopperrender_cuda.h"
#include "hr_common.cuh"
#include "hr_pack.cuh"
#include "hr_search.cuh"
#include "hr_search2.cuh"
#include "hr_search3.cuh"
#include "hr_staging.h"
#include "hr_pacing_predict.h"
#include "hr_warp.cuh"
#include "hr_warp_fast.cuh"

#include <cuda.h> /* CUtensorMap and its encoder, reached through cudaGetDriverEntryPoint: no link against libcuda */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_CALC_RES 270 /* video/filter/HopperRender/config.h:2 */
#define HR_WARP_STREAMS 3
/* Three search lanes (measured against two, same box: 29.3 -> 25.7 us per source frame at 1080p R = 5, 48.7 -> 45.5 at
 * R = 16, 24.5 -> 18.5 at 720p; nothing at 4K / 8K where the warps are the step): three launches of the third search
 * generation share the SMs at three CTAs each (HR3_CTAS_PER_SM); four lanes at four CTAs (64 registers) measured the same
 * at R = 5 and slower from R = 8 on. The rings follow from the lanes: */
#ifndef HR_SEARCH_LANES
#define HR_SEARCH_LANES 3
#endif
#ifndef HR_FLOW_BUFS
#define HR_FLOW_BUFS (2 * HR_SEARCH_LANES)
#endif
/* HR_FLOW_BUFS: blurred-flow ring: one being written per search lane, the rest read by warps in flight; a multiple of
 * the lanes, so that a buffer is rewritten on the lane that wrote it last */
#ifndef HR_PACK_BUFS
#define HR_PACK_BUFS (HR_SEARCH_LANES + 1)
#endif
/* HR_PACK_BUFS: packed copies in rotation: one per search in flight and the one being built — with fewer, the pack of
 * frame k has to wait for a search that still reads the buffer it overwrites and sits on the lanes' critical path */
#define HR_MAX_WARP_EVENTS 8

struct HrContext {
    int H, W, aW, pixfmt, bps;
    int s, lw, lh, first, iters;
    int device, smCount;
    int tilesX, tilesY, numTiles, grid, multiTile;
    int planePitch, planeSize;
    size_t frameSamples, frameBytes, packedBytes, deviceBytes;

    cudaStream_t stream, ownStream;
    uint8_t *frameBuf[2];          /* owned frame slots                                        */
    const void *fy[2], *fuv[2];    /* [0] previous (frame1 / sourceFrame12), [1] newest         */
    int fslot[2];                  /* which owned slot each of them uses (-1: borrowed)         */
    uint32_t *packed[2];           /* packed copies, same order (two of the HR_PACK_BUFS buffers of packedRing) */
    uint32_t *packedRing[HR_PACK_BUFS];
    int packedSpare[HR_PACK_BUFS - 2]; /* ring indices of the buffers neither slot uses, the one to reuse next first */
    uint8_t *outBuf;
    void *outY, *outUV;            /* current output planes (internal or caller's)              */
    int16_t *off, *blur;
    uint32_t *blurXY;              /* blurred flow, x | y << 16 per lattice point (what the warp reads)     */
    unsigned long long *T;
    int tOff[HR_MAX_LEVELS];
    int tWords;
    unsigned long long *partial;   /* cross-tile window sums [bigStep][tile][HR_RMAX], epoch-tagged  */
    int bigOff[2 * HR_MAX_LEVELS];
    int bigWords;
    uint32_t epoch;                /* tag of the last search launch                                */
    uint8_t *trace;
    int traceOn;
    long long *timeline;
    long long *timelineHost; /* HR_TIMELINE_HOST=1: the stamps live in mapped host memory */
    int timelineOn;
    int useFastWarp;
    int haveRcp8;                  /* MUFU.RCP of the 8-bit level denominators, cached per knob setting     */
    float rcp8Den[2], rcp8[2];

    /* spatial bands: this context owns rows [bandRow0[bandRank], bandRow1[bandRank]) */
    int bandWorld, bandRank;
    int bandRow0[HR_MAX_BANDS], bandRow1[HR_MAX_BANDS];
    unsigned char *peerSlot[HR_MAX_BANDS][2];       /* the peers' two frame slots, mapped into this process   */
    unsigned long long *peerMail[HR_MAX_BANDS];     /* the peers' mailboxes: [0] uploaded, [1] consumed        */
    unsigned char *peerArena[HR_MAX_BANDS];         /* the peers' exchange arenas                              */
    unsigned long long *mail;                       /* own mailbox (inside the arena)                          */
    unsigned long long bandFrames;                  /* frames uploaded so far                                  */
    int bandPending;                                /* hr_band_upload done, hr_band_gather outstanding        */
    int banded;                                     /* bands configured (also with a single band)              */
    int bandSearched;                               /* a search of the group was launched since the last upload */
    int bandMaxRadius;                              /* largest search radius the halo is sized for             */
    int haloLo, haloHi;                             /* rows of every frame this GPU holds: band + halo         */
    unsigned long long bandP2pBytes;                /* bytes fetched from peers' frame slots so far             */
    /* exchange arena: everything the searches of the band group store into one another, one allocation with the same
     * layout on every GPU (hr_common.cuh BandLink) */
    unsigned char *arena;
    size_t arenaBytes, aT, aPartial, aOff, aBlur, aBlurXY, aReady, aDone, aExit, aMail;
    unsigned long long *plainT, *plainPartial;      /* what the context used before the bands were configured  */
    int16_t *plainOff, *plainBlur;
    uint32_t *plainBlurXY;

    /* pipelined mode (hr_set_pipeline): independent work of consecutive frame pairs overlaps on internal streams —
     * pack(k) || search(k) (the search reads frame 2 as it arrived), the warps of one pair on HR_WARP_STREAMS
     * streams, search(k+1) || warps(k) (two flow buffers). Dependencies are CUDA events, never host waits. */
    int pipeline;
    cudaStream_t sPack, sSearch[HR_SEARCH_LANES], sWarp[HR_WARP_STREAMS];
    cudaEvent_t evIn;                          /* main stream: the newest frame's planes are complete          */
    cudaEvent_t evLattice;                     /* main stream: the rows of the newest frame the search reads (its lattice rows) are there */
    int latticeFirst;                          /* this frame was uploaded lattice rows first: the search waits for evLattice, not evIn */
    int splitUpload;                           /* developer knob HR_SPLIT_UPLOAD=0: one transfer per frame, always */
    int stageSplit;                            /* developer knob HR_STAGE_SPLIT=0: pageable planes go up in one piece instead of lattice rows first */
    cudaEvent_t evPack[HR_PACK_BUFS];                     /* by packed-buffer identity: its pack kernel is done           */
    cudaEvent_t packRead[HR_PACK_BUFS];                   /* by packed-buffer identity: the search that read it last (not owned) */
    cudaEvent_t evSearch[HR_FLOW_BUFS];        /* by flow buffer: the search that filled it is done            */
    cudaEvent_t evWarp[HR_FLOW_BUFS][HR_MAX_WARP_EVENTS]; /* by flow buffer: the warps reading it              */
    int nWarpEv[HR_FLOW_BUFS], haveSearch[HR_FLOW_BUFS], havePack[HR_PACK_BUFS], packedId[2];
    /* Every record of the events above takes a serial number. pipe_join orders the main stream after all of them; once
     * the host has waited for the main stream behind such a join, everything up to the join's number is over, and the
     * next join leaves those events out (a blocking update used to spend ~15 us of stream waits on work long finished
     * before its upload could start). */
    unsigned long long evSeq, joinedSeq, doneSeq;
    unsigned long long packSeq[HR_PACK_BUFS], searchSeq[HR_FLOW_BUFS], warpSeq[HR_FLOW_BUFS][HR_MAX_WARP_EVENTS];
    int flowCur;                               /* flow buffer of the most recent search                        */
    int16_t *blurB[HR_FLOW_BUFS];
    uint32_t *blurXYB[HR_FLOW_BUFS];
    /* search lanes: consecutive pairs are independent (the offsets start from zero for every pair,
     * opticalFlowCalc.c:153), so the search of pair k+1 is launched on the next lane — its own stream, window
     * tables, tile totals and raw-offset array — and fills the launch gaps and hand-off waits of the searches before it */
    int lane;                                  /* lane of the most recent search                               */
    int16_t *offL[HR_SEARCH_LANES];
    unsigned long long *TL[HR_SEARCH_LANES], *partialL[HR_SEARCH_LANES];
    unsigned warpRR;

    /* Work started ahead of the blocking calls that ask for it (pipelined mode, host interface): the search of a
     * new pair from hr_update_frame with the knobs of the previous hr_calc_flow, and the next warp of a pair from
     * hr_download with the blending scalar the pacing will most likely ask for next. A call whose arguments match
     * finds its work under way or done; one that does not match launches its own. Results are the same bits. */
    struct FlowKnobs {
        int valid, R, dS, nS, frames;
    } lastFlow, specFlow;
    struct WarpAhead {
        int valid, mode, frames;
        float t, black, white;
        uint32_t epoch;
        cudaEvent_t done; /* not owned: the flow-buffer reader event of the launch */
    } specWarp;
    uint8_t *outBuf2;      /* second internal output frame: what is warped ahead goes here                  */
    /* HSV output mode: flow colours per lattice cell, by flow buffer (built lazily, once per flow) */
    uint32_t *colours[HR_FLOW_BUFS];
    cudaEvent_t evColour[HR_FLOW_BUFS];
    unsigned long long flowSerial[HR_FLOW_BUFS], colourSerial[HR_FLOW_BUFS]; /* content stamps: flow buffer / its colour table */
    unsigned long long flowStamp;
    float lastT, lastDelta, lastBlack, lastWhite;
    HrPacingPredictor *pace; /* the filter's pacing arithmetic followed from the blending scalars it asks for (hr_pacing_predict.h) */
    struct FirstAhead {      /* hr_download's guess when the next output belongs to the next source frame: warped from hr_update_frame */
        int valid, frames;
        float t;
    } firstAhead;
    int lastWarpFrames, lastMode, aheadOn;
    int lastSearchGen; /* generation the most recent search launch ran */
    int lastSearchStaged; /* ... and whether it was the TMA-staged variant */
    int stagedOk;      /* tensor maps of the two packed copies exist (resolution scalar 2, driver has the encoder) */
    int stagedOn;      /* developer knob HR_SEARCH_STAGED=0 / hr_debug_set_search_staged */
    HrTensorMap tmapPacked[HR_PACK_BUFS]; /* [physical packed buffer]: 3-D view (words, rows, phase planes) for the staged search */
    int searchGen; /* 0 (default): chosen per launch (launch_flow); 3 / 2: hr_search3.cuh / hr_search2.cuh where they apply (radius 5..16, one tile per CTA, no bands), 1: hr_search.cuh always */

    /* pageable host planes (hr_staging.h): a ring of pinned chunks and the threads that fill / drain it */
    HrCopyCrew *crew;
    HrStagePlan *plan;
    int stageThreads;
    size_t stageChunk;
    uint8_t *stage;
    cudaEvent_t evStage[HR_STAGE_SLOTS]; /* the copy engine is done with the slot */
    int stageBusy[HR_STAGE_SLOTS];
    unsigned stageNext;

    cudaEvent_t evUpdate, evFlowEnd, evWarpStart, evDlEnd;
    cudaEvent_t evK[6]; /* search start/end, warp start/end, pack start/end */
    int profiling;
    int haveSearchT, haveWarpT, havePackT;
    int framesSeen;
    uint64_t launches;
    char err[256];
};

static char g_createErr[256];
/* bytes that crossed the host <-> device boundary on behalf of the six interface calls (not the parity taps):
 * hr_debug_host_transfer_bytes; what a zero-copy integration must leave untouched */
static unsigned long long g_h2dBytes, g_d2hBytes;

static int fail(HrContext *ctx, const char *fmt, ...) {
    char *dst = ctx ? ctx->err : g_createErr;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 256, fmt, ap);
    va_end(ap);
    fprintf(stderr, "[HopperRender/CUDA] %s\n", dst);
    return 1;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(ctx, "CUDA error in %s: %s (%s)", __func__, cudaGetErrorString(e_), #call); \
    } while (0)

static int bind_device(HrContext *ctx) {
    CU(cudaSetDevice(ctx->device));
    return 0;
}

static int sync_all(HrContext *ctx);
static void pipeline_release(HrContext *ctx);
static int pipe_join(HrContext *ctx);
static int pipe_on(const HrContext *ctx);
static void stage_release(HrContext *ctx);
static int warp_ahead_first(HrContext *ctx);
static int launch_flow(HrContext *ctx, int searchRadius, int deltaScalar, int neighborBiasScalar, cudaStream_t *stOut);

extern "C" int hr_abi_version(void) { return HR_ABI_VERSION; }

extern "C" const char *hr_last_error(const HrContext *ctx) { return ctx ? ctx->err : g_createErr; }

extern "C" uint64_t hr_get_launch_count(const HrContext *ctx) { return ctx ? ctx->launches : 0; }

/* Everything hr_set_pipeline allocates (streams, events, ring entries >= 1, the second output frame). Safe on a
 * partly built set: every handle is NULL until created and NULL again afterwards. */
static void pipeline_release(HrContext *ctx) {
    if (ctx->sPack) cudaStreamDestroy(ctx->sPack);
    ctx->sPack = NULL;
    for (int i = 0; i < HR_SEARCH_LANES; ++i) {
        if (ctx->sSearch[i]) cudaStreamDestroy(ctx->sSearch[i]);
        ctx->sSearch[i] = NULL;
    }
    for (int i = 0; i < HR_WARP_STREAMS; ++i) {
        if (ctx->sWarp[i]) cudaStreamDestroy(ctx->sWarp[i]);
        ctx->sWarp[i] = NULL;
    }
    if (ctx->evIn) cudaEventDestroy(ctx->evIn);
    ctx->evIn = NULL;
    if (ctx->evLattice) cudaEventDestroy(ctx->evLattice);
    ctx->evLattice = NULL;
    for (int b = 0; b < HR_PACK_BUFS; ++b) {
        if (ctx->evPack[b]) cudaEventDestroy(ctx->evPack[b]);
        ctx->evPack[b] = NULL;
        ctx->packRead[b] = NULL;
        ctx->havePack[b] = 0;
    }
    for (int b = 0; b < HR_FLOW_BUFS; ++b) {
        if (ctx->evSearch[b]) cudaEventDestroy(ctx->evSearch[b]);
        ctx->evSearch[b] = NULL;
        ctx->haveSearch[b] = 0;
        ctx->nWarpEv[b] = 0;
        for (int i = 0; i < HR_MAX_WARP_EVENTS; ++i) {
            if (ctx->evWarp[b][i]) cudaEventDestroy(ctx->evWarp[b][i]);
            ctx->evWarp[b][i] = NULL;
        }
    }
    for (int b = 1; b < HR_FLOW_BUFS; ++b) {
        cudaFree(ctx->blurB[b]);
        cudaFree(ctx->blurXYB[b]);
        ctx->blurB[b] = NULL;
        ctx->blurXYB[b] = NULL;
    }
    for (int l = 1; l < HR_SEARCH_LANES; ++l) {
        cudaFree(ctx->offL[l]);
        cudaFree(ctx->TL[l]);
        cudaFree(ctx->partialL[l]);
        ctx->offL[l] = NULL;
        ctx->TL[l] = NULL;
        ctx->partialL[l] = NULL;
    }
    cudaFree(ctx->outBuf2);
    ctx->outBuf2 = NULL;
}

static void window_schedule(int lw, int lh, int *first, int *iters) {
    /* opticalFlowCalc.c:133-149 */
    int windowSize = 1;
    int maxDim = lw > lh ? lw : lh;
    if (maxDim && !(maxDim & (maxDim - 1))) {
        windowSize = maxDim;
    } else {
        while (maxDim & (maxDim - 1)) maxDim &= (maxDim - 1);
        windowSize = maxDim << 1;
    }
    windowSize /= 2;
    int n = 0;
    for (int w = windowSize; w > 1; w >>= 1) n++;
    *first = windowSize;
    *iters = n;
}

extern "C" int hr_destroy(HrContext *ctx) {
    if (!ctx) return 1;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaDeviceSynchronize();
    stage_release(ctx);
    delete ctx->pace;
    ctx->pace = NULL;
    /* ctx->blur / blurXY / off / T / partial alias entry [x] of their rings; entry [0] is freed below through them */
    pipeline_release(ctx);
    if (ctx->banded) { /* the flow arrays live in the arena while bands are configured: back to the context's own */
        ctx->T = ctx->plainT;
        ctx->partial = ctx->plainPartial;
        ctx->off = ctx->plainOff;
        ctx->blur = ctx->plainBlur;
        ctx->blurXY = ctx->plainBlurXY;
    }
    /* (a context whose creation failed half-way has no ring yet: keep what was allocated directly) */
    if (ctx->blurB[0]) ctx->blur = ctx->blurB[0];
    if (ctx->blurXYB[0]) ctx->blurXY = ctx->blurXYB[0];
    if (ctx->offL[0]) ctx->off = ctx->offL[0];
    if (ctx->TL[0]) ctx->T = ctx->TL[0];
    if (ctx->partialL[0]) ctx->partial = ctx->partialL[0];
    cudaFree(ctx->frameBuf[0]);
    cudaFree(ctx->frameBuf[1]);
    for (int b = 0; b < HR_PACK_BUFS; ++b) cudaFree(ctx->packedRing[b]);
    cudaFree(ctx->outBuf);
    cudaFree(ctx->off);
    cudaFree(ctx->blur);
    cudaFree(ctx->blurXY);
    cudaFree(ctx->T);
    cudaFree(ctx->partial);
    cudaFree(ctx->trace);
    for (int b = 0; b < HR_FLOW_BUFS; ++b) {
        cudaFree(ctx->colours[b]);
        if (ctx->evColour[b]) cudaEventDestroy(ctx->evColour[b]);
    }
    cudaFree(ctx->arena);
    if (ctx->timelineHost) cudaFreeHost(ctx->timelineHost);
    else cudaFree(ctx->timeline);
    if (ctx->evUpdate) cudaEventDestroy(ctx->evUpdate);
    if (ctx->evFlowEnd) cudaEventDestroy(ctx->evFlowEnd);
    if (ctx->evWarpStart) cudaEventDestroy(ctx->evWarpStart);
    if (ctx->evDlEnd) cudaEventDestroy(ctx->evDlEnd);
    for (int i = 0; i < 6; ++i)
        if (ctx->evK[i]) cudaEventDestroy(ctx->evK[i]);
    if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
    free(ctx);
    return 0;
}

/* 3-D tensor maps (words of a plane row, plane rows, phase planes) over the two packed copies, boxes of the staged
 * search (hr_search2.cuh): 60 words x 50 rows x 1 plane, out-of-range parts filled with zeros. */
static int make_packed_tensor_maps(HrContext *ctx, uint32_t *const physical[HR_PACK_BUFS]) {
    ctx->stagedOk = 0;
    if (ctx->s != HR_ST_S) return 0;
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = NULL;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return 0; /* no encoder: the unstaged kernels serve every launch */
    }
    static_assert(sizeof(CUtensorMap) == sizeof(HrTensorMap), "tensor map size");
    for (int i = 0; i < HR_PACK_BUFS; ++i) {
        const cuuint64_t dims[3] = {(cuuint64_t)ctx->planePitch, (cuuint64_t)ctx->lh, (cuuint64_t)1 << (2 * ctx->s)};
        const cuuint64_t strides[2] = {(cuuint64_t)ctx->planePitch * 4, (cuuint64_t)ctx->planeSize * 4};
        const cuuint32_t box[3] = {HR_ST_PW, HR_ST_PH, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUtensorMap tm;
        const CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, physical[i], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(ctx, "cuTensorMapEncodeTiled failed (%d) for the packed frame", (int)r);
        memcpy(&ctx->tmapPacked[i], &tm, sizeof(tm));
    }
    ctx->stagedOk = 1;
    return 0;
}

static int create_impl(HrContext *ctx) {
    if (bind_device(ctx)) return 1;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    ctx->smCount = prop.multiProcessorCount;
    if (!prop.cooperativeLaunch) return fail(ctx, "device %d does not support cooperative launch", ctx->device);

    /* opticalFlowCalc.c:331-336 */
    ctx->s = 0;
    while ((ctx->H >> ctx->s) > MAX_CALC_RES) ctx->s++;
    ctx->lw = (int)ceil(ctx->W / pow(2, ctx->s));
    ctx->lh = (int)ceil(ctx->H / pow(2, ctx->s));
    if (ctx->s > 4)
        return fail(ctx, "frames of more than %d lines are not supported (resolution scalar %d > 4)", MAX_CALC_RES << 4, ctx->s);
    window_schedule(ctx->lw, ctx->lh, &ctx->first, &ctx->iters);
    if (ctx->iters < 1 || ctx->iters > HR_MAX_LEVELS) return fail(ctx, "unsupported lattice %dx%d", ctx->lw, ctx->lh);

    ctx->tilesX = (ctx->lw + HR_TILE - 1) / HR_TILE;
    ctx->tilesY = (ctx->lh + HR_TILE - 1) / HR_TILE;
    ctx->numTiles = ctx->tilesX * ctx->tilesY;
    int perSm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, flow_search_generic_kernel<true, false>, HR_THREADS, 0));
    if (perSm > 1) perSm = 1; /* one tile per SM: the search is latency-bound, spread it out */
    if (perSm < 1) return fail(ctx, "search kernel does not fit on an SM");
    const int maxResident = perSm * ctx->smCount;
    ctx->grid = ctx->numTiles < maxResident ? ctx->numTiles : maxResident;
    ctx->multiTile = ctx->numTiles > ctx->grid;
    if ((ctx->numTiles + ctx->grid - 1) / ctx->grid > HR_MAX_TILES_PER_CTA)
        return fail(ctx, "lattice %dx%d needs more than %d tiles per CTA", ctx->lw, ctx->lh, HR_MAX_TILES_PER_CTA);

    ctx->planePitch = (ctx->lw + 31) & ~31;
    ctx->planeSize = ctx->planePitch * ctx->lh;
    const size_t planes = (size_t)1 << (2 * ctx->s);
    ctx->packedBytes = planes * (size_t)ctx->planeSize * sizeof(uint32_t);
    if (planes * (size_t)ctx->planeSize > 0x7fffffffULL) return fail(ctx, "frame too large for the packed layout");
    ctx->frameSamples = (size_t)ctx->H * ctx->W + (size_t)(ctx->H / 2) * ctx->W;
    ctx->frameBytes = ctx->frameSamples * ctx->bps;

    /* level tables; 32-word aligned so that levels never share a 128-byte line */
    int words = 0;
    for (int it = 0; it < ctx->iters; ++it) {
        const int ws = ctx->first >> it;
        const int nwx = (ctx->lw + ws - 1) / ws, nwy = (ctx->lh + ws - 1) / ws;
        ctx->tOff[it] = words;
        words += (nwx * nwy + 31) & ~31;
    }
    ctx->tWords = words;
    /* per-tile totals of the search steps whose windows span several tiles */
    int bwords = 0;
    for (int k = 0; k < 2 * HR_MAX_LEVELS; ++k) ctx->bigOff[k] = -1;
    for (int it = 0; it < ctx->iters; ++it) {
        if ((ctx->first >> it) > HR_TILE) {
            for (int axis = 0; axis < 2; ++axis) {
                ctx->bigOff[it * 2 + axis] = bwords;
                bwords += ctx->numTiles * HR_RMAX;
            }
        }
    }
    ctx->bigWords = bwords;

    const size_t ln = (size_t)ctx->lw * ctx->lh;
    CU(cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking));
    ctx->stream = ctx->ownStream;
    CU(cudaMalloc(&ctx->frameBuf[0], ctx->frameBytes));
    CU(cudaMalloc(&ctx->frameBuf[1], ctx->frameBytes));
    CU(cudaMalloc(&ctx->outBuf, ctx->frameBytes));
    for (int b = 0; b < HR_PACK_BUFS; ++b) CU(cudaMalloc(&ctx->packedRing[b], ctx->packedBytes));
    ctx->packed[0] = ctx->packedRing[0];
    ctx->packed[1] = ctx->packedRing[1];
    for (int i = 0; i < HR_PACK_BUFS - 2; ++i) ctx->packedSpare[i] = 2 + i;
    CU(cudaMalloc(&ctx->off, 2 * ln * sizeof(int16_t)));
    CU(cudaMalloc(&ctx->blur, 2 * ln * sizeof(int16_t)));
    CU(cudaMalloc(&ctx->blurXY, ln * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->T, (size_t)(words ? words : 32) * sizeof(unsigned long long)));
    CU(cudaMalloc(&ctx->partial, (size_t)(bwords ? bwords : 32) * sizeof(unsigned long long)));
    CU(cudaMemset(ctx->frameBuf[0], 0, ctx->frameBytes));
    CU(cudaMemset(ctx->frameBuf[1], 0, ctx->frameBytes));
    CU(cudaMemset(ctx->outBuf, 0, ctx->frameBytes));
    for (int b = 0; b < HR_PACK_BUFS; ++b) CU(cudaMemset(ctx->packedRing[b], 0, ctx->packedBytes));
    if (make_packed_tensor_maps(ctx, ctx->packedRing)) return 1;
    CU(cudaMemset(ctx->off, 0, 2 * ln * sizeof(int16_t)));
    CU(cudaMemset(ctx->blur, 0, 2 * ln * sizeof(int16_t)));
    CU(cudaMemset(ctx->blurXY, 0, ln * sizeof(uint32_t)));
    CU(cudaMemset(ctx->T, 0, (size_t)(words ? words : 32) * sizeof(unsigned long long)));
    CU(cudaMemset(ctx->partial, 0, (size_t)(bwords ? bwords : 32) * sizeof(unsigned long long)));
    ctx->epoch = 0;
    ctx->blurB[0] = ctx->blur;
    ctx->blurXYB[0] = ctx->blurXY;
    ctx->offL[0] = ctx->off;
    ctx->TL[0] = ctx->T;
    ctx->partialL[0] = ctx->partial;
    ctx->packedId[0] = 0;
    ctx->packedId[1] = 1;
    ctx->deviceBytes = 3 * ctx->frameBytes + HR_PACK_BUFS * ctx->packedBytes + 4 * ln * sizeof(int16_t) + (size_t)(words + bwords) * 8 + 516;
    for (int i = 0; i < 2; ++i) {
        ctx->fy[i] = ctx->frameBuf[i];
        ctx->fuv[i] = ctx->frameBuf[i] + (size_t)ctx->H * ctx->W * ctx->bps;
        ctx->fslot[i] = i;
    }
    ctx->outY = ctx->outBuf;
    ctx->outUV = ctx->outBuf + (size_t)ctx->H * ctx->W * ctx->bps;
    CU(cudaEventCreate(&ctx->evUpdate));
    CU(cudaEventCreate(&ctx->evFlowEnd));
    CU(cudaEventCreate(&ctx->evWarpStart));
    CU(cudaEventCreate(&ctx->evDlEnd));
    for (int i = 0; i < 6; ++i) CU(cudaEventCreate(&ctx->evK[i]));
    const char *g = getenv("HR_WARP_GENERIC");
    ctx->useFastWarp = !(g && g[0] == '1');
    const char *ah = getenv("HR_AHEAD");
    ctx->aheadOn = !(ah && ah[0] == '0');
    const char *sg = getenv("HR_SEARCH_GEN"); /* developer knob: 1 / 2 / 3 = that generation for every launch it can serve */
    ctx->searchGen = (sg && sg[0] >= '1' && sg[0] <= '3') ? sg[0] - '0' : 0;
    const char *su = getenv("HR_SPLIT_UPLOAD"); /* developer knob: 0 = never upload a frame lattice rows first */
    ctx->splitUpload = !(su && su[0] == '0');
    const char *stsp = getenv("HR_STAGE_SPLIT");
    ctx->stageSplit = !(stsp && stsp[0] == '0');
    const char *stt = getenv("HR_STAGE_THREADS"), *stc = getenv("HR_STAGE_CHUNK_KB"); /* hr_staging.h */
    ctx->stageThreads = stt ? atoi(stt) : 4;
    if (ctx->stageThreads > 16) ctx->stageThreads = 16;
    const int chunkKb = stc ? atoi(stc) : 1024;
    ctx->stageChunk = (size_t)(chunkKb < 64 ? 64 : chunkKb > 16384 ? 16384 : chunkKb) << 10;
    const char *ss = getenv("HR_SEARCH_STAGED"); /* developer knob: 0 = never the TMA-staged variant */
    ctx->stagedOn = !(ss && ss[0] == '0');
    CU(cudaDeviceSynchronize());
    return 0;
}

extern "C" int hr_create(HrContext **out, int frameHeight, int frameWidth, int actualWidth, int pixfmt, int device) {
    if (!out) return fail(NULL, "hr_create: out is NULL");
    *out = NULL;
    if (pixfmt != HR_PIXFMT_NV12 && pixfmt != HR_PIXFMT_P010) return fail(NULL, "hr_create: unknown pixel format %d", pixfmt);
    if (frameHeight < 8 || frameWidth < 8 || actualWidth < 8 || actualWidth > frameWidth || (frameHeight & 1) || (frameWidth & 1))
        return fail(NULL, "hr_create: unsupported geometry h=%d stride=%d w=%d (need even h/stride >= 8, w <= stride)", frameHeight, frameWidth, actualWidth);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(NULL, "hr_create: no CUDA device available (there is no CPU fallback)");
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return fail(NULL, "hr_create: cudaGetDevice failed");
    if (device >= ndev) return fail(NULL, "hr_create: device %d out of range (%d devices)", device, ndev);
    HrContext *ctx = (HrContext *)calloc(1, sizeof(HrContext));
    if (ctx) {
        ctx->pace = new HrPacingPredictor;
        ctx->pace->reset();
    }
    if (!ctx) return fail(NULL, "hr_create: out of host memory");
    ctx->H = frameHeight;
    ctx->W = frameWidth;
    ctx->aW = actualWidth;
    ctx->pixfmt = pixfmt;
    ctx->bps = pixfmt == HR_PIXFMT_P010 ? 2 : 1;
    ctx->device = device;
    if (create_impl(ctx)) {
        snprintf(g_createErr, sizeof(g_createErr), "%s", ctx->err);
        hr_destroy(ctx);
        return 1;
    }
    *out = ctx;
    return 0;
}

extern "C" int hr_get_info(const HrContext *ctx, HrInfo *info) {
    if (!ctx || !info) return 1;
    memset(info, 0, sizeof(*info));
    info->abiVersion = HR_ABI_VERSION;
    info->device = ctx->device;
    info->frameHeight = ctx->H;
    info->frameWidth = ctx->W;
    info->actualWidth = ctx->aW;
    info->pixfmt = ctx->pixfmt;
    info->resScalar = ctx->s;
    info->lowWidth = ctx->lw;
    info->lowHeight = ctx->lh;
    info->firstWindow = ctx->first;
    info->iterations = ctx->iters;
    info->searchCtas = ctx->grid;
    info->smCount = ctx->smCount;
    info->frameBytes = ctx->frameBytes;
    info->deviceBytes = ctx->deviceBytes;
    return 0;
}

extern "C" int hr_set_stream(HrContext *ctx, void *cudaStream) {
    if (!ctx) return 1;
    if (bind_device(ctx)) return 1;
    if (sync_all(ctx)) return 1;
    ctx->stream = cudaStream ? (cudaStream_t)cudaStream : ctx->ownStream;
    return 0;
}

extern "C" int hr_synchronize(HrContext *ctx) {
    if (!ctx) return 1;
    if (bind_device(ctx)) return 1;
    return sync_all(ctx);
}

extern "C" int hr_set_trace(HrContext *ctx, int enable) {
    if (!ctx) return 1;
    if (bind_device(ctx)) return 1;
    if (enable && !ctx->trace) {
        const size_t n = (size_t)2 * ctx->iters * ctx->lw * ctx->lh;
        CU(cudaMalloc(&ctx->trace, n));
        CU(cudaMemset(ctx->trace, 0, n));
    }
    ctx->traceOn = enable ? 1 : 0;
    return 0;
}

__global__ void rcp_table_kernel(float *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = rcp_approx((float)i);
}

/* Parity tap: MUFU.RCP((float)i), 0 <= i < n — the reciprocal inside div.full.f32, which the warp's
 * level mapping uses. tests/golden/make_rcp_table.py stores it for the CPU-side checks. */
extern "C" int hr_debug_rcp_table(float *out, int n) {
    HrContext *ctx = NULL;
    if (!out || n < 1) return 1;
    float *d = NULL;
    CU(cudaMalloc(&d, (size_t)n * sizeof(float)));
    rcp_table_kernel<<<(n + 255) / 256, 256>>>(d, n);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out, d, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(NULL, "CUDA error in hr_debug_rcp_table: %s", cudaGetErrorString(e));
    return 0;
}

/* ---- INT-pipe roofline denominator: issue rate of the packed-byte SAD ----------------------
 * Every candidate evaluation of the search is ONE vabsdiff4.u32.u32.u32.add (SASS: VABSDIFF4.U8.ACC). This kernel
 * issues nothing else: CHAINS independent accumulator chains per thread, fully unrolled, 2048 resident threads per
 * SM, so that the result is the pipe's issue rate and not its latency. */
#define HR_PEAK_CHAINS 8
#define HR_PEAK_UNROLL 8
__global__ void __launch_bounds__(1024, 2) int_peak_kernel(uint32_t *out, int iters, long long *cycles) {
    uint32_t acc[HR_PEAK_CHAINS], a[HR_PEAK_CHAINS];
    const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
#pragma unroll
    for (int j = 0; j < HR_PEAK_CHAINS; ++j) {
        acc[j] = j;
        a[j] = b ^ (0x01020408u * (j + 1));
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < HR_PEAK_UNROLL; ++u)
#pragma unroll
            for (int j = 0; j < HR_PEAK_CHAINS; ++j) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(a[j]), "r"(b));
    }
    const long long t1 = clock64();
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < HR_PEAK_CHAINS; ++j) x ^= acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

/* Developer tap: packed-SAD issue rate of `device` (< 0: current). evalsPerSecond = thread-level VABSDIFF4.U8.ACC per
 * second over the whole GPU by CUDA events (= candidate evaluations per second if the search did nothing but its
 * SADs); warpInstrPerClkPerSm from the SM clock counter of the slowest CTA. Either pointer may be NULL. */
extern "C" int hr_debug_predict_pacing(const float *scalars, int n, float *predicted, int *nextFrame, int *have) {
    if (!scalars || !predicted || !nextFrame || !have || n < 0) return 1;
    HrPacingPredictor p;
    p.reset();
    for (int i = 0; i < n; ++i) {
        bool nf = false;
        float t = 0.0f;
        have[i] = p.predict(&t, &nf) ? 1 : 0;
        predicted[i] = t;
        nextFrame[i] = nf ? 1 : 0;
        p.observe(scalars[i]);
    }
    return 0;
}

extern "C" int hr_debug_int_peak(int device, double *evalsPerSecond, double *warpInstrPerClkPerSm) {
    HrContext *ctx = NULL;
    if (device < 0) CU(cudaGetDevice(&device));
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 2, threads = 1024, iters = 4096;
    uint32_t *out = NULL;
    long long *cyc = NULL, *hcyc = (long long *)malloc(sizeof(long long) * blocks);
    cudaEvent_t e0 = NULL, e1 = NULL;
    cudaError_t e = hcyc ? cudaSuccess : cudaErrorMemoryAllocation;
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaMalloc(&out, (size_t)blocks * threads * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&cyc, (size_t)blocks * sizeof(long long));
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) {
        int_peak_kernel<<<blocks, threads>>>(out, 64, cyc); /* warm-up */
        e = cudaEventRecord(e0, 0);
        int_peak_kernel<<<blocks, threads>>>(out, iters, cyc);
        if (e == cudaSuccess) e = cudaEventRecord(e1, 0);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e == cudaSuccess) e = cudaMemcpy(hcyc, cyc, (size_t)blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    }
    if (e == cudaSuccess) {
        const double perThread = (double)iters * HR_PEAK_UNROLL * HR_PEAK_CHAINS;
        long long worst = 1;
        for (int i = 0; i < blocks; ++i) worst = hcyc[i] > worst ? hcyc[i] : worst;
        if (evalsPerSecond) *evalsPerSecond = perThread * blocks * threads / ((double)ms * 1e-3);
        /* two 1024-thread CTAs per SM = 64 warps share the SM for `worst` cycles */
        if (warpInstrPerClkPerSm) *warpInstrPerClkPerSm = perThread * 64.0 / (double)worst;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(out);
    cudaFree(cyc);
    free(hcyc);
    if (e != cudaSuccess) return fail(NULL, "CUDA error in hr_debug_int_peak: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int hr_set_timeline(HrContext *ctx, int enable) {
    if (!ctx) return 1;
    if (bind_device(ctx)) return 1;
    if (enable && !ctx->timeline) {
        const size_t n = (size_t)ctx->grid * HR_TIMELINE_SLOTS * sizeof(long long);
        const char *hostTl = getenv("HR_TIMELINE_HOST"); /* developer knob: stamps in mapped host memory (hr_debug_peek_timeline) */
        if (hostTl && hostTl[0] == '1') {
            CU(cudaHostAlloc((void **)&ctx->timelineHost, n, cudaHostAllocMapped));
            memset(ctx->timelineHost, 0, n);
            CU(cudaHostGetDevicePointer((void **)&ctx->timeline, ctx->timelineHost, 0));
        } else {
            CU(cudaMalloc(&ctx->timeline, n));
            CU(cudaMemset(ctx->timeline, 0, n));
        }
    }
    ctx->timelineOn = enable ? 1 : 0;
    return 0;
}

extern "C" int hr_debug_set_search_generation(HrContext *ctx, int generation) {
    if (!ctx) return 1;
    if (generation < 0 || generation > 3) return fail(ctx, "hr_debug_set_search_generation: %d is not 0 (chosen per launch), 1, 2 or 3", generation);
    ctx->searchGen = generation;
    return 0;
}

/* the stamps as they stand right now, without waiting for anything (HR_TIMELINE_HOST=1 only) */
extern "C" int hr_debug_peek_timeline(HrContext *ctx, long long *stamps, int maxCtas) {
    if (!ctx) return 1;
    if (!ctx->timelineHost) return fail(ctx, "hr_debug_peek_timeline: needs HR_TIMELINE_HOST=1 and hr_set_timeline");
    const int n = maxCtas < ctx->grid ? maxCtas : ctx->grid;
    memcpy(stamps, ctx->timelineHost, (size_t)n * HR_TIMELINE_SLOTS * sizeof(long long));
    return 0;
}

extern "C" int hr_debug_last_search_generation(const HrContext *ctx) { return ctx ? ctx->lastSearchGen : 0; }
extern "C" int hr_debug_last_search_staged(const HrContext *ctx) { return ctx ? ctx->lastSearchStaged : 0; }
extern "C" int hr_debug_set_search_staged(HrContext *ctx, int enable) {
    if (!ctx) return 1;
    ctx->stagedOn = enable ? 1 : 0;
    return 0;
}

extern "C" int hr_get_timeline(HrContext *ctx, long long *stamps, int maxCtas) {
    if (!ctx || !stamps) return 1;
    if (!ctx->timeline) return fail(ctx, "hr_get_timeline: the timeline was not enabled");
    if (bind_device(ctx)) return 1;
    const int n = maxCtas < ctx->grid ? maxCtas : ctx->grid;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(stamps, ctx->timeline, (size_t)n * HR_TIMELINE_SLOTS * sizeof(long long), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int hr_set_profiling(HrContext *ctx, int enable) {
    if (!ctx) return 1;
    ctx->profiling = enable ? 1 : 0;
    return 0;
}

/* pack the newest frame (slot 1) into its phase-planar copy */
template <typename T>
static void launch_pack_t(HrContext *ctx, cudaStream_t st) {
    const T *y = (const T *)ctx->fy[1], *uv = (const T *)ctx->fuv[1];
    const bool vec = ctx->s <= 4 && ctx->W % 16 == 0 && (((uintptr_t)y | (uintptr_t)uv) % 16 == 0);
    /* bands: only the rows this GPU holds (band + halo, even bounds) */
    const int rowLo = ctx->banded ? ctx->haloLo : 0, rowHi = ctx->banded ? ctx->haloHi : ctx->H;
    if (vec) {
        dim3 block(128), grid((ctx->W / 16 + 127) / 128, (rowHi - rowLo + 1) / 2);
        switch (ctx->s) {
#define HR_PCASE(S_) case S_: pack_frame16_kernel<T, S_><<<grid, block, 0, st>>>(y, uv, ctx->packed[1], ctx->W, rowHi, ctx->planePitch, ctx->planeSize, rowLo >> 1); break;
            HR_PCASE(0) HR_PCASE(1) HR_PCASE(2) HR_PCASE(3) HR_PCASE(4)
#undef HR_PCASE
        }
    } else {
        const int bx = ctx->s <= 3 ? 128 : 64; /* s <= 4 (hr_create): at most 64 x 16 threads */
        dim3 block(bx, 1 << ctx->s);
        dim3 grid((ctx->lw + bx - 1) / bx, rowHi - rowLo);
        pack_frame_kernel<T><<<grid, block, 0, st>>>(y, uv, ctx->packed[1], ctx->W, ctx->H, ctx->s, ctx->lw, ctx->planePitch, ctx->planeSize, rowLo);
    }
}
static int launch_pack(HrContext *ctx) {
    cudaStream_t st = ctx->stream;
    const int id = ctx->packedId[1];
    if (pipe_on(ctx)) {
        /* on its own stream, next to the search of the same pair (which does not read this copy): after the
         * frame has arrived and after the search that read the buffer being overwritten */
        st = ctx->sPack;
        CU(cudaEventRecord(ctx->evIn, ctx->stream));
        CU(cudaStreamWaitEvent(st, ctx->evIn, 0));
        if (ctx->packRead[id]) CU(cudaStreamWaitEvent(st, ctx->packRead[id], 0));
    }
    if (ctx->profiling) CU(cudaEventRecord(ctx->evK[4], st));
    if (ctx->bps == 1) launch_pack_t<uint8_t>(ctx, st);
    else launch_pack_t<uint16_t>(ctx, st);
    CU(cudaGetLastError());
    if (ctx->profiling) {
        CU(cudaEventRecord(ctx->evK[5], st));
        ctx->havePackT = 1;
    }
    if (ctx->sPack) {
        CU(cudaEventRecord(ctx->evPack[id], st));
        ctx->packSeq[id] = ++ctx->evSeq;
        ctx->havePack[id] = 1;
    }
    ctx->launches++;
    return 0;
}

/* opticalFlowCalc.c:102-105: the slot that held the previous-previous frame receives the new one */
static void rotate_slots(HrContext *ctx, int *freeSlot) {
    /* the owned slot not used by the current newest frame */
    int used = ctx->fslot[1];
    *freeSlot = used == 0 ? 1 : 0;
    ctx->fy[0] = ctx->fy[1];
    ctx->fuv[0] = ctx->fuv[1];
    ctx->fslot[0] = ctx->fslot[1];
    /* the copy just built becomes the one the next search reads; the next frame is packed into the spare buffer, and
     * the buffer the last search read becomes the spare */
    const int id = ctx->packedId[0];
    ctx->packedId[0] = ctx->packedId[1];
    ctx->packedId[1] = ctx->packedSpare[0];
    for (int i = 0; i + 1 < HR_PACK_BUFS - 2; ++i) ctx->packedSpare[i] = ctx->packedSpare[i + 1];
    ctx->packedSpare[HR_PACK_BUFS - 3] = id;
    ctx->packed[0] = ctx->packedRing[ctx->packedId[0]];
    ctx->packed[1] = ctx->packedRing[ctx->packedId[1]];
}

/* ---- pipelined mode ----------------------------------------------------------------------------------- */
static int pipe_on(const HrContext *ctx) { return ctx->pipeline && !ctx->banded; }

/* make `st` wait for every warp that reads flow buffer b */
static int wait_warps(HrContext *ctx, cudaStream_t st, int b) {
    for (int i = 0; i < ctx->nWarpEv[b]; ++i) CU(cudaStreamWaitEvent(st, ctx->evWarp[b][i], 0));
    return 0;
}
/* order the main stream after everything in flight on the internal streams (no host wait) */
static int pipe_join(HrContext *ctx) {
    if (!ctx->sPack) return 0;
    const unsigned long long done = ctx->doneSeq;
    for (int b = 0; b < HR_PACK_BUFS; ++b)
        if (ctx->havePack[b] && ctx->packSeq[b] > done) CU(cudaStreamWaitEvent(ctx->stream, ctx->evPack[b], 0));
    for (int b = 0; b < HR_FLOW_BUFS; ++b) {
        if (ctx->haveSearch[b] && ctx->searchSeq[b] > done) CU(cudaStreamWaitEvent(ctx->stream, ctx->evSearch[b], 0));
        for (int i = 0; i < ctx->nWarpEv[b]; ++i)
            if (ctx->warpSeq[b][i] > done) CU(cudaStreamWaitEvent(ctx->stream, ctx->evWarp[b][i], 0));
    }
    ctx->joinedSeq = ctx->evSeq;
    return 0;
}
/* the host has waited for the main stream: what the last join ordered in front of it is over */
static void joined_work_is_done(HrContext *ctx) { ctx->doneSeq = ctx->joinedSeq; }
static int sync_all(HrContext *ctx) {
    if (pipe_join(ctx)) return 1;
    CU(cudaStreamSynchronize(ctx->stream));
    joined_work_is_done(ctx);
    return 0;
}

/* streams, events, the ring of flow buffers, the second search lane, the second output frame */
static int pipeline_alloc(HrContext *ctx) {
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    /* the search is the critical path of a pair: its CTAs are placed first */
    for (int i = 0; i < HR_SEARCH_LANES; ++i) CU(cudaStreamCreateWithPriority(&ctx->sSearch[i], cudaStreamNonBlocking, hi));
    CU(cudaStreamCreateWithPriority(&ctx->sPack, cudaStreamNonBlocking, lo));
    for (int i = 0; i < HR_WARP_STREAMS; ++i) CU(cudaStreamCreateWithPriority(&ctx->sWarp[i], cudaStreamNonBlocking, lo));
    CU(cudaEventCreateWithFlags(&ctx->evIn, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->evLattice, cudaEventDisableTiming));
    for (int b = 0; b < HR_PACK_BUFS; ++b) CU(cudaEventCreateWithFlags(&ctx->evPack[b], cudaEventDisableTiming));
    for (int b = 0; b < HR_FLOW_BUFS; ++b) {
        CU(cudaEventCreateWithFlags(&ctx->evSearch[b], cudaEventDisableTiming));
        for (int i = 0; i < HR_MAX_WARP_EVENTS; ++i) CU(cudaEventCreateWithFlags(&ctx->evWarp[b][i], cudaEventDisableTiming));
    }
    const size_t ln = (size_t)ctx->lw * ctx->lh;
    /* the flow buffer in use becomes entry 0 of the ring, the search scratch in use lane 0 */
    ctx->blurB[0] = ctx->blur;
    ctx->blurXYB[0] = ctx->blurXY;
    ctx->flowCur = 0;
    for (int b = 1; b < HR_FLOW_BUFS; ++b) {
        CU(cudaMalloc(&ctx->blurB[b], 2 * ln * sizeof(int16_t)));
        CU(cudaMalloc(&ctx->blurXYB[b], ln * sizeof(uint32_t)));
        CU(cudaMemset(ctx->blurB[b], 0, 2 * ln * sizeof(int16_t)));
        CU(cudaMemset(ctx->blurXYB[b], 0, ln * sizeof(uint32_t)));
    }
    const size_t tw = (size_t)(ctx->tWords ? ctx->tWords : 32) * sizeof(unsigned long long);
    const size_t bw = (size_t)(ctx->bigWords ? ctx->bigWords : 32) * sizeof(unsigned long long);
    ctx->offL[0] = ctx->off;
    ctx->TL[0] = ctx->T;
    ctx->partialL[0] = ctx->partial;
    ctx->lane = 0;
    for (int l = 1; l < HR_SEARCH_LANES; ++l) {
        CU(cudaMalloc(&ctx->offL[l], 2 * ln * sizeof(int16_t)));
        CU(cudaMalloc(&ctx->TL[l], tw));
        CU(cudaMalloc(&ctx->partialL[l], bw));
        CU(cudaMemset(ctx->offL[l], 0, 2 * ln * sizeof(int16_t)));
        CU(cudaMemset(ctx->TL[l], 0, tw));
        CU(cudaMemset(ctx->partialL[l], 0, bw));
    }
    CU(cudaMalloc(&ctx->outBuf2, ctx->frameBytes));
    CU(cudaMemset(ctx->outBuf2, 0, ctx->frameBytes));
    ctx->deviceBytes += ctx->frameBytes;
    ctx->deviceBytes += (HR_FLOW_BUFS - 1) * (2 * ln * sizeof(int16_t) + ln * sizeof(uint32_t)) + (HR_SEARCH_LANES - 1) * (2 * ln * sizeof(int16_t) + tw + bw);
    CU(cudaDeviceSynchronize());
    return 0;
}

extern "C" int hr_set_pipeline(HrContext *ctx, int enable) {
    if (!ctx) return 1;
    if (bind_device(ctx)) return 1;
    if (sync_all(ctx)) return 1;
    if (enable && !ctx->sPack && pipeline_alloc(ctx)) {
        /* half a pipeline is none: give back what was created, stay in (or return to) the serial mode on ring entry 0 */
        pipeline_release(ctx);
        ctx->blur = ctx->blurB[0] ? ctx->blurB[0] : ctx->blur;
        ctx->blurXY = ctx->blurXYB[0] ? ctx->blurXYB[0] : ctx->blurXY;
        ctx->flowCur = 0;
        ctx->lane = 0;
        ctx->pipeline = 0;
        cudaGetLastError();
        return 1;
    }
    ctx->pipeline = enable ? 1 : 0;
    ctx->specWarp.valid = ctx->specFlow.valid = 0;
    return 0;
}

extern "C" int hr_pipeline_join(HrContext *ctx) {
    if (!ctx) return 1;
    if (bind_device(ctx)) return 1;
    return pipe_join(ctx);
}

/* ---- pageable host planes through the pinned ring (hr_staging.h) ---- */
static int host_is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
/* 1: the ring is there (built on first use), 0: staging is switched off, -1: error */
static int stage_ready(HrContext *ctx) {
    if (ctx->stageThreads < 1) return 0;
    if (ctx->stage) return 1;
    if (cudaHostAlloc((void **)&ctx->stage, HR_STAGE_SLOTS * ctx->stageChunk, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        ctx->stage = NULL;
        ctx->stageThreads = 0; /* no pinned memory to be had: the driver's own path from here on */
        return 0;
    }
    for (int i = 0; i < HR_STAGE_SLOTS; ++i) {
        if (cudaEventCreateWithFlags(&ctx->evStage[i], cudaEventDisableTiming) != cudaSuccess) return fail(ctx, "staging ring: %s", cudaGetErrorString(cudaGetLastError())), -1;
        ctx->stageBusy[i] = 0;
    }
    ctx->crew = new HrCopyCrew(ctx->stageThreads);
    ctx->plan = new HrStagePlan();
    return 1;
}
static void stage_release(HrContext *ctx) {
    delete ctx->crew;
    ctx->crew = NULL;
    delete ctx->plan;
    ctx->plan = NULL;
    for (int i = 0; i < HR_STAGE_SLOTS; ++i) {
        if (ctx->evStage[i]) cudaEventDestroy(ctx->evStage[i]);
        ctx->evStage[i] = NULL;
    }
    if (ctx->stage) cudaFreeHost(ctx->stage);
    ctx->stage = NULL;
}
/* every transfer of the crew ends in finish(), also one that a CUDA call cut short */
struct StageTransfer {
    HrCopyCrew *crew;
    ~StageTransfer() { crew->finish(); }
};
/* host -> device: the crew fills the ring chunk by chunk, the copy engine follows it; everything is enqueued on return */
/* blockBytes != 0: a pitched picture on both sides — bytes / blockBytes blocks of blockBytes every `pitch` bytes, back
 * to back in the ring, one pitched copy per chunk (blockBytes <= the chunk size) */
static int staged_h2d(HrContext *ctx, uint8_t *dDst, const uint8_t *hSrc, size_t bytes, size_t blockBytes = 0, size_t pitch = 0) {
    const size_t ch = ctx->stageChunk;
    if (blockBytes == pitch) blockBytes = pitch = 0; /* blocks that touch are one range */
    if (blockBytes) ctx->plan->build_blocks(bytes / blockBytes, blockBytes, ch, true);
    else ctx->plan->build(bytes, ch, true);
    const size_t nch = ctx->plan->n;
    const unsigned base = ctx->stageNext;
    ctx->stageNext += (unsigned)nch;
    size_t released = 0;
    for (; released < nch && released < HR_STAGE_SLOTS; ++released) {
        const int slot = (base + released) % HR_STAGE_SLOTS;
        if (ctx->stageBusy[slot]) CU(cudaEventSynchronize(ctx->evStage[slot]));
    }
    ctx->crew->begin(true, (uint8_t *)hSrc, ctx->stage, ch, base, ctx->plan, blockBytes, pitch);
    StageTransfer guard{ctx->crew};
    ctx->crew->release(released);
    for (size_t c = 0; c < nch; ++c) {
        while (!ctx->crew->chunk_done(c)) {
            /* a slot goes back to the crew as soon as the copy engine is through with the chunk that was in it */
            if (released < nch && released < c + HR_STAGE_SLOTS && cudaEventQuery(ctx->evStage[(base + released) % HR_STAGE_SLOTS]) == cudaSuccess)
                ctx->crew->release(++released);
            else
                HrCopyCrew::relax();
        }
        cudaGetLastError(); /* cudaErrorNotReady of the queries */
        const int slot = (base + c) % HR_STAGE_SLOTS;
        const size_t o = ctx->plan->off[c], len = ctx->plan->len[c];
        if (!blockBytes) CU(cudaMemcpyAsync(dDst + o, ctx->stage + slot * ch, len, cudaMemcpyHostToDevice, ctx->stream));
        else CU(cudaMemcpy2DAsync(dDst + (o / blockBytes) * pitch, pitch, ctx->stage + slot * ch, blockBytes, blockBytes, len / blockBytes, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaEventRecord(ctx->evStage[slot], ctx->stream));
        ctx->stageBusy[slot] = 1;
        if (released <= c + 1 && released < nch) { /* the crew is about to run out of slots: wait for the oldest chunk in flight */
            CU(cudaEventSynchronize(ctx->evStage[(base + released) % HR_STAGE_SLOTS]));
            ctx->crew->release(++released);
        }
    }
    return 0;
}
/* The same for an upload of several segments in ONE run of the ring (no seam between them: the crew goes on filling slots
 * while the copy engine works off the previous segment). When the last chunk of segment hookAfter has been handed to the
 * copy engine, hook(ctx) runs on this thread — the crew keeps copying into the slots already released. */
static int staged_h2d_segments(HrContext *ctx, const HrStageSeg *segs, int nSegs, int hookAfter, int (*hook)(HrContext *)) {
    const size_t ch = ctx->stageChunk;
    ctx->plan->build_segments(segs, nSegs, ch);
    const size_t nch = ctx->plan->n;
    const unsigned base = ctx->stageNext;
    ctx->stageNext += (unsigned)nch;
    size_t released = 0;
    for (; released < nch && released < HR_STAGE_SLOTS; ++released) {
        const int slot = (base + released) % HR_STAGE_SLOTS;
        if (ctx->stageBusy[slot]) CU(cudaEventSynchronize(ctx->evStage[slot]));
    }
    ctx->crew->begin(true, NULL, ctx->stage, ch, base, ctx->plan);
    StageTransfer guard{ctx->crew};
    ctx->crew->release(released);
    for (size_t c = 0; c < nch; ++c) {
        while (!ctx->crew->chunk_done(c)) {
            if (released < nch && released < c + HR_STAGE_SLOTS && cudaEventQuery(ctx->evStage[(base + released) % HR_STAGE_SLOTS]) == cudaSuccess)
                ctx->crew->release(++released);
            else
                HrCopyCrew::relax();
        }
        cudaGetLastError(); /* cudaErrorNotReady of the queries */
        const int slot = (base + c) % HR_STAGE_SLOTS;
        const HrStageSeg &sg = segs[ctx->plan->seg[c]];
        const size_t so = ctx->plan->segOff[c], len = ctx->plan->len[c];
        if (!sg.blockBytes) CU(cudaMemcpyAsync(sg.dev + so, ctx->stage + slot * ch, len, cudaMemcpyHostToDevice, ctx->stream));
        else CU(cudaMemcpy2DAsync(sg.dev + (so / sg.blockBytes) * sg.pitch, sg.pitch, ctx->stage + slot * ch, sg.blockBytes, sg.blockBytes, len / sg.blockBytes, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaEventRecord(ctx->evStage[slot], ctx->stream));
        ctx->stageBusy[slot] = 1;
        if (hook && ctx->plan->seg[c] == hookAfter && (c + 1 == nch || ctx->plan->seg[c + 1] != hookAfter)) {
            if (hook(ctx)) return 1;
        }
        if (released <= c + 1 && released < nch) { /* the crew is about to run out of slots: wait for the oldest chunk in flight */
            CU(cudaEventSynchronize(ctx->evStage[(base + released) % HR_STAGE_SLOTS]));
            ctx->crew->release(++released);
        }
    }
    return 0;
}
/* device -> host in two halves, so that the caller can put work between them: the first chunks are handed to the copy
 * engine and the crew is told where they will land, ... */
struct StageDown {
    unsigned base;
    size_t queued;
};
static int staged_d2h_enqueue(HrContext *ctx, const uint8_t *dSrc, const StageDown &d, size_t k) {
    const size_t ch = ctx->stageChunk;
    const int slot = (d.base + k) % HR_STAGE_SLOTS;
    const size_t o = ctx->plan->off[k], len = ctx->plan->len[k];
    CU(cudaMemcpyAsync(ctx->stage + slot * ch, dSrc + o, len, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->evStage[slot], ctx->stream));
    ctx->stageBusy[slot] = 1;
    return 0;
}
static int staged_d2h_begin(HrContext *ctx, uint8_t *hDst, const uint8_t *dSrc, size_t bytes, StageDown *d) {
    const size_t ch = ctx->stageChunk;
    ctx->plan->build(bytes, ch, false);
    const size_t nch = ctx->plan->n;
    d->base = ctx->stageNext;
    ctx->stageNext += (unsigned)nch;
    for (d->queued = 0; d->queued < nch && d->queued < HR_STAGE_SLOTS; ++d->queued) {
        const int slot = (d->base + d->queued) % HR_STAGE_SLOTS;
        if (ctx->stageBusy[slot]) CU(cudaEventSynchronize(ctx->evStage[slot])); /* an upload's chunk still in the slot */
        if (staged_d2h_enqueue(ctx, dSrc, *d, d->queued)) return 1;
    }
    ctx->crew->begin(false, hDst, ctx->stage, ch, d->base, ctx->plan);
    return 0;
}
/* ... and drained in order by the crew, every emptied slot going back to the copy engine; complete on return */
static int staged_d2h_end(HrContext *ctx, const uint8_t *dSrc, StageDown *d) {
    StageTransfer guard{ctx->crew};
    const size_t nch = ctx->plan->n;
    for (size_t c = 0; c < nch; ++c) {
        while (d->queued < nch && (d->queued <= c || ctx->crew->chunk_done(d->queued - HR_STAGE_SLOTS))) {
            ctx->crew->wait_chunk(d->queued - HR_STAGE_SLOTS); /* waits only when chunk c itself has no slot yet */
            if (staged_d2h_enqueue(ctx, dSrc, *d, d->queued)) return 1;
            ++d->queued;
        }
        CU(cudaEventSynchronize(ctx->evStage[(d->base + c) % HR_STAGE_SLOTS]));
        ctx->crew->release(c + 1);
    }
    ctx->crew->finish();
    for (size_t c = nch > HR_STAGE_SLOTS ? nch - HR_STAGE_SLOTS : 0; c < nch; ++c) ctx->stageBusy[(d->base + c) % HR_STAGE_SLOTS] = 0;
    return 0;
}

extern "C" int hr_update_frame(HrContext *ctx, const void *yPlane, const void *uvPlane) {
    if (!ctx) return 1;
    if (!yPlane || !uvPlane) return fail(ctx, "hr_update_frame: NULL plane");
    if (bind_device(ctx)) return 1;
    if (pipe_join(ctx)) return 1; /* the slot being overwritten may still be read by warps in flight */
    CU(cudaEventRecord(ctx->evUpdate, ctx->stream));
    int slot;
    rotate_slots(ctx, &slot);
    uint8_t *dst = ctx->frameBuf[slot];
    const size_t ylen = (size_t)ctx->H * ctx->W * ctx->bps, uvlen = (size_t)(ctx->H / 2) * ctx->W * ctx->bps;
    ctx->fy[1] = dst;
    ctx->fuv[1] = dst + ylen;
    ctx->fslot[1] = slot;
    ctx->framesSeen++;
    ctx->specWarp.valid = ctx->specFlow.valid = 0;
    ctx->latticeFirst = 0;
    /* the search the filter asks for next, with the knobs it used last time: under way while this call waits for the
     * upload and the caller gets round to calculateOpticalFlow */
    const int ahead = pipe_on(ctx) && ctx->aheadOn && ctx->lastFlow.valid && ctx->framesSeen >= 2 && !ctx->traceOn && !ctx->timelineOn;
    /* pageable planes (what mpv's pool and the decoder hand the filter) go through the pinned ring */
    int staged = 0;
    if (ctx->stageThreads >= 1 && (host_is_pageable(yPlane) || host_is_pageable(uvPlane))) {
        staged = stage_ready(ctx);
        if (staged < 0) return 1;
    }
    if (staged && ahead && ctx->splitUpload && ctx->stageSplit && ctx->s >= 1 && (size_t)((1 << ctx->s) - 1) * ctx->W * ctx->bps <= ctx->stageChunk) {
        /* lattice rows first, as below for pinned planes: the crew gathers them into the ring, the copy engine scatters
         * them with pitched copies, the search starts behind them and runs while the other rows follow — all of it one
         * run of the ring (staged_h2d_segments; HR_STAGE_SPLIT=0 restores one piece per frame). */
        const size_t rowBytes = (size_t)ctx->W * ctx->bps;
        uint8_t *src[2] = {(uint8_t *)yPlane, (uint8_t *)uvPlane};
        uint8_t *dpl[2] = {dst, dst + ylen};
        const int rows[2] = {ctx->H, ctx->H / 2}, stride[2] = {1 << ctx->s, (1 << ctx->s) / 2 > 1 ? (1 << ctx->s) / 2 : 1};
        HrStageSeg segs[6];
        int n = 0;
        for (int pl = 0; pl < 2; ++pl) { /* first row of every group of `stride` rows */
            const size_t groups = (size_t)(rows[pl] + stride[pl] - 1) / stride[pl];
            if (stride[pl] > 1) segs[n++] = HrStageSeg{src[pl], dpl[pl], groups * rowBytes, rowBytes, stride[pl] * rowBytes};
            else segs[n++] = HrStageSeg{src[pl], dpl[pl], groups * rowBytes, 0, 0};
        }
        const int latticeSegs = n;
        for (int pl = 0; pl < 2; ++pl) { /* the other rows of every full group, then what is left of a partial last group */
            const size_t full = (size_t)rows[pl] / stride[pl], tail = rows[pl] - full * stride[pl];
            if (stride[pl] > 1 && full > 0)
                segs[n++] = HrStageSeg{src[pl] + rowBytes, dpl[pl] + rowBytes, full * (stride[pl] - 1) * rowBytes, stride[pl] > 2 ? (stride[pl] - 1) * rowBytes : rowBytes,
                                       stride[pl] * rowBytes};
            if (tail > 1) {
                const size_t o = (full * stride[pl] + 1) * rowBytes;
                segs[n++] = HrStageSeg{src[pl] + o, dpl[pl] + o, (tail - 1) * rowBytes, 0, 0};
            }
        }
        auto latticeArrived = [](HrContext *c) -> int {
            if (cudaEventRecord(c->evLattice, c->stream) != cudaSuccess) return fail(c, "hr_update_frame: %s", cudaGetErrorString(cudaGetLastError()));
            c->latticeFirst = 1;
            cudaStream_t st;
            if (launch_flow(c, c->lastFlow.R, c->lastFlow.dS, c->lastFlow.nS, &st)) return 1;
            if (cudaEventRecord(c->evFlowEnd, st) != cudaSuccess) return fail(c, "hr_update_frame: %s", cudaGetErrorString(cudaGetLastError()));
            c->specFlow = c->lastFlow;
            c->specFlow.frames = c->framesSeen;
            return 0;
        };
        if (staged_h2d_segments(ctx, segs, n, latticeSegs - 1, latticeArrived)) return 1;
        g_h2dBytes += ylen + uvlen;
        if (launch_pack(ctx)) return 1;
        ctx->latticeFirst = 0;
    } else if (staged) {
        if ((const uint8_t *)uvPlane == (const uint8_t *)yPlane + ylen) {
            if (staged_h2d(ctx, dst, (const uint8_t *)yPlane, ylen + uvlen)) return 1;
        } else {
            if (staged_h2d(ctx, dst, (const uint8_t *)yPlane, ylen)) return 1;
            if (staged_h2d(ctx, dst + ylen, (const uint8_t *)uvPlane, uvlen)) return 1;
        }
        g_h2dBytes += ylen + uvlen;
        if (launch_pack(ctx)) return 1;
        if (ahead) {
            cudaStream_t st;
            if (launch_flow(ctx, ctx->lastFlow.R, ctx->lastFlow.dS, ctx->lastFlow.nS, &st)) return 1;
            CU(cudaEventRecord(ctx->evFlowEnd, st));
            ctx->specFlow = ctx->lastFlow;
            ctx->specFlow.frames = ctx->framesSeen;
        }
    } else if (ahead && ctx->splitUpload && ctx->s >= 1) {
        /* The search reads the newest frame at its lattice points only: every 2^s-th luma row and the chroma rows under
         * them, a third of the bytes at 1080p. Those rows go first (pitched copies), the search starts behind them and
         * runs while the other rows are still crossing PCIe; pack and warp wait for the whole frame. (Measured against
         * "lattice luma rows + the whole chroma plane first", one pitched copy fewer: the search then starts 12 us later
         * and the caller waits for it: update + flow 107 us instead of 98 us at 1080p.) */
        const size_t rowBytes = (size_t)ctx->W * ctx->bps;
        const uint8_t *src[2] = {(const uint8_t *)yPlane, (const uint8_t *)uvPlane};
        uint8_t *dpl[2] = {dst, dst + ylen};
        const int rows[2] = {ctx->H, ctx->H / 2}, stride[2] = {1 << ctx->s, (1 << ctx->s) / 2 > 1 ? (1 << ctx->s) / 2 : 1};
        for (int pl = 0; pl < 2; ++pl) { /* first row of every group of `stride` rows */
            const int groups = (rows[pl] + stride[pl] - 1) / stride[pl];
            CU(cudaMemcpy2DAsync(dpl[pl], stride[pl] * rowBytes, src[pl], stride[pl] * rowBytes, rowBytes, groups, cudaMemcpyHostToDevice, ctx->stream));
        }
        CU(cudaEventRecord(ctx->evLattice, ctx->stream));
        ctx->latticeFirst = 1;
        cudaStream_t st;
        if (launch_flow(ctx, ctx->lastFlow.R, ctx->lastFlow.dS, ctx->lastFlow.nS, &st)) return 1;
        CU(cudaEventRecord(ctx->evFlowEnd, st));
        ctx->specFlow = ctx->lastFlow;
        ctx->specFlow.frames = ctx->framesSeen;
        for (int pl = 0; pl < 2; ++pl) { /* the other rows of every full group, then what is left of a partial last group */
            const int full = rows[pl] / stride[pl], tail = rows[pl] - full * stride[pl];
            if (stride[pl] > 1 && full > 0)
                CU(cudaMemcpy2DAsync(dpl[pl] + rowBytes, stride[pl] * rowBytes, src[pl] + rowBytes, stride[pl] * rowBytes, (stride[pl] - 1) * rowBytes, full,
                                     cudaMemcpyHostToDevice, ctx->stream));
            if (tail > 1) {
                const size_t o = ((size_t)full * stride[pl] + 1) * rowBytes;
                CU(cudaMemcpyAsync(dpl[pl] + o, src[pl] + o, (size_t)(tail - 1) * rowBytes, cudaMemcpyHostToDevice, ctx->stream));
            }
        }
        g_h2dBytes += ylen + uvlen;
        if (launch_pack(ctx)) return 1; /* records evIn behind the last row */
        ctx->latticeFirst = 0;          /* later launches for this frame (other knobs) wait for all of it */
    } else {
        if ((const uint8_t *)uvPlane == (const uint8_t *)yPlane + ylen) {
            /* planes of one allocation, back to back (mpv's image pool lays NV12 out like this): one transfer */
            CU(cudaMemcpyAsync(dst, yPlane, ylen + uvlen, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            CU(cudaMemcpyAsync(dst, yPlane, ylen, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemcpyAsync(dst + ylen, uvPlane, uvlen, cudaMemcpyHostToDevice, ctx->stream));
        }
        g_h2dBytes += ylen + uvlen;
        if (launch_pack(ctx)) return 1;
        if (ahead) {
            cudaStream_t st;
            if (launch_flow(ctx, ctx->lastFlow.R, ctx->lastFlow.dS, ctx->lastFlow.nS, &st)) return 1;
            CU(cudaEventRecord(ctx->evFlowEnd, st));
            ctx->specFlow = ctx->lastFlow;
            ctx->specFlow.frames = ctx->framesSeen;
        }
    }
    if (warp_ahead_first(ctx)) return 1;
    CU(cudaStreamSynchronize(ctx->stream)); /* the reference's writes are blocking (CL_TRUE) */
    joined_work_is_done(ctx);                /* the join at the top of this call */
    if (staged) memset(ctx->stageBusy, 0, sizeof(ctx->stageBusy)); /* the copy engine is through with every slot */
    return 0;
}

extern "C" int hr_update_frame_device(HrContext *ctx, const void *dY, const void *dUV, int borrow) {
    if (!ctx) return 1;
    if (!dY || !dUV) return fail(ctx, "hr_update_frame_device: NULL plane");
    if (bind_device(ctx)) return 1;
    ctx->specWarp.valid = ctx->specFlow.valid = 0;
    if (!borrow && pipe_join(ctx)) return 1; /* the slot being overwritten may still be read by warps in flight */
    if (!pipe_on(ctx)) CU(cudaEventRecord(ctx->evUpdate, ctx->stream));
    int slot;
    rotate_slots(ctx, &slot);
    const size_t ylen = (size_t)ctx->H * ctx->W * ctx->bps, uvlen = (size_t)(ctx->H / 2) * ctx->W * ctx->bps;
    if (borrow) {
        ctx->fy[1] = dY;
        ctx->fuv[1] = dUV;
        ctx->fslot[1] = -1;
    } else {
        uint8_t *dst = ctx->frameBuf[slot];
        CU(cudaMemcpyAsync(dst, dY, ylen, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(dst + ylen, dUV, uvlen, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->fy[1] = dst;
        ctx->fuv[1] = dst + ylen;
        ctx->fslot[1] = slot;
    }
    if (launch_pack(ctx)) return 1;
    ctx->framesSeen++;
    return 0;
}

/* The search radius the filter uses drifts between MIN_SEARCH_RADIUS and MAX_SEARCH_RADIUS
 * (config.h:6-7, vf_HopperRender.c:326-345): those radii get a kernel with the layer loop fully
 * unrolled and the layer shifts as immediates; any other radius runs the generic kernel. */
/* the TMA-staged variant of generation 2 (resolution scalar 2, radius 5..HR_ST_MAX_RADIUS); NULL: none for this radius */
static const void *staged_kernel_for(int R, int timeline) {
    const void *k = NULL;
    if (timeline) k = R == 5 ? (const void *)flow_search2_staged_kernel<5, true> : NULL;
    else {
        switch (R) {
            case 5: k = (const void *)flow_search2_staged_kernel<5>; break;
            case 6: k = (const void *)flow_search2_staged_kernel<6>; break;
            case 7: k = (const void *)flow_search2_staged_kernel<7>; break;
            case 8: k = (const void *)flow_search2_staged_kernel<8>; break;
            default: break;
        }
    }
    if (k) {
        static const void *prepared[16];
        static int nPrepared = 0;
        int seen = 0;
        for (int i = 0; i < nPrepared; ++i) seen |= prepared[i] == k;
        if (!seen) {
            if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, HR_ST_BYTES) != cudaSuccess) {
                cudaGetLastError();
                return NULL;
            }
            if (nPrepared < 16) prepared[nPrepared++] = k;
        }
    }
    return k;
}

static const void *search_kernel_for(int R, int multi, int timeline, int gen, int *used) {
    *used = 1;
    if (multi) return (const void *)flow_search_generic_kernel<true, false>;
    if (gen == 3 && R >= 5 && R <= 16 && !(timeline && R != 5)) {
        *used = 3;
        if (timeline) return (const void *)flow_search3_kernel<5, true>;
        switch (R) {
#define HR_RCASE(r) case r: return (const void *)flow_search3_kernel<r>;
            HR_RCASE(5) HR_RCASE(6) HR_RCASE(7) HR_RCASE(8) HR_RCASE(9) HR_RCASE(10) HR_RCASE(11) HR_RCASE(12)
            HR_RCASE(13) HR_RCASE(14) HR_RCASE(15) HR_RCASE(16)
#undef HR_RCASE
            default: break;
        }
    }
    if (timeline && gen == 2 && (R == 5 || R == 16)) {
        *used = 2;
        return R == 5 ? (const void *)flow_search2_kernel<5, true> : (const void *)flow_search2_kernel<16, true>;
    }
    if (timeline) return R == 5 ? (const void *)flow_search_kernel<5, false, true> : (const void *)flow_search_generic_kernel<false, true>;
    if (gen == 2) {
        switch (R) {
#define HR_RCASE(r) case r: *used = 2; return (const void *)flow_search2_kernel<r>;
            HR_RCASE(5) HR_RCASE(6) HR_RCASE(7) HR_RCASE(8) HR_RCASE(9) HR_RCASE(10) HR_RCASE(11) HR_RCASE(12)
            HR_RCASE(13) HR_RCASE(14) HR_RCASE(15) HR_RCASE(16)
#undef HR_RCASE
            default: break;
        }
    }
    switch (R) {
#define HR_RCASE(r) case r: return (const void *)flow_search_kernel<r, false, false>;
        HR_RCASE(5) HR_RCASE(6) HR_RCASE(7) HR_RCASE(8) HR_RCASE(9) HR_RCASE(10) HR_RCASE(11) HR_RCASE(12)
        HR_RCASE(13) HR_RCASE(14) HR_RCASE(15) HR_RCASE(16)
#undef HR_RCASE
        default: return (const void *)flow_search_generic_kernel<false, false>;
    }
}

/* enqueue K1-K4 for the current frame pair; *stOut = the stream it went to */
static int launch_flow(HrContext *ctx, int searchRadius, int deltaScalar, int neighborBiasScalar, cudaStream_t *stOut) {
    FlowParams P;
    memset(&P, 0, sizeof(P));
    P.p1 = ctx->packed[0];
    P.f2y = ctx->fy[1];
    P.f2uv = ctx->fuv[1];
    P.bps = ctx->bps;
    P.planePitch = ctx->planePitch;
    P.planeSize = ctx->planeSize;
    P.W = ctx->W;
    P.H = ctx->H;
    P.s = ctx->s;
    P.lw = ctx->lw;
    P.lh = ctx->lh;
    P.first = ctx->first;
    P.iters = ctx->iters;
    P.R = searchRadius;
    P.dS = deltaScalar;
    P.nS = neighborBiasScalar;
    for (int z = 0; z < HR_RMAX; ++z) {
        const int rel = z - searchRadius / 2;
        P.cand[z] = z < searchRadius ? rel * abs(rel) : 0;
    }
    P.tilesX = ctx->tilesX;
    P.tilesY = ctx->tilesY;
    P.numTiles = ctx->numTiles;
    memcpy(P.tOff, ctx->tOff, sizeof(P.tOff));
    memcpy(P.bigOff, ctx->bigOff, sizeof(P.bigOff));
    if (++ctx->epoch == 0) ctx->epoch = 1; /* 0 is the tag of never-written words */
    P.epoch = ctx->epoch;
    /* pipelined: into the flow buffer the warps of the previous pair are not reading */
    const int pl = pipe_on(ctx);
    const int fb = pl ? (ctx->flowCur + 1) % HR_FLOW_BUFS : ctx->flowCur;
    /* the next search lane — unless a tap wants to look at this launch's tables afterwards */
    const int lane = (pl && !ctx->traceOn && !ctx->timelineOn) ? (ctx->lane + 1) % HR_SEARCH_LANES : ctx->lane;
    cudaStream_t st = pl ? ctx->sSearch[lane] : ctx->stream;
    P.T = ctx->TL[lane];
    P.partial = ctx->partialL[lane];
    P.off = ctx->offL[lane];
    P.blur = ctx->blurB[fb];
    P.blurXY = ctx->blurXYB[fb];
    P.trace = ctx->traceOn ? ctx->trace : NULL;
    P.timeline = ctx->timelineOn ? ctx->timeline : NULL;
    int grid = ctx->grid;
    if (ctx->banded) {
        /* this GPU's tile rows only; tables, totals and the flow live in the exchange arena (same layout everywhere) */
        const int tileRows = HR_TILE << ctx->s;
        const int r0 = ctx->bandRow0[ctx->bandRank], r1 = ctx->bandRow1[ctx->bandRank];
        BandLink &B = P.band;
        B.world = ctx->bandWorld;
        B.rank = ctx->bandRank;
        B.tileRow0 = r0 / tileRows;
        B.tileRow1 = (r1 + tileRows - 1) / tileRows;
        B.tile0 = B.tileRow0 * ctx->tilesX;
        B.up = ctx->bandRank > 0 ? ctx->bandRank - 1 : -1;
        B.down = ctx->bandRank + 1 < ctx->bandWorld ? ctx->bandRank + 1 : -1;
        for (int g = 0; g < ctx->bandWorld; ++g) {
            if (!ctx->peerArena[g]) return fail(ctx, "hr_calc_flow: band peer %d is not connected", g);
            B.peerDelta[g] = (long long)((intptr_t)ctx->peerArena[g] - (intptr_t)ctx->arena);
        }
        B.ready = (unsigned long long *)(ctx->arena + ctx->aReady);
        B.done = (unsigned long long *)(ctx->arena + ctx->aDone);
        B.exitCount = (unsigned int *)(ctx->arena + ctx->aExit);
        P.T = (unsigned long long *)(ctx->arena + ctx->aT);
        P.partial = (unsigned long long *)(ctx->arena + ctx->aPartial);
        P.off = (int16_t *)(ctx->arena + ctx->aOff);
        P.blur = (int16_t *)(ctx->arena + ctx->aBlur);
        P.blurXY = (uint32_t *)(ctx->arena + ctx->aBlurXY);
        P.trace = NULL;
        P.timeline = NULL;
        ctx->bandSearched = 1;
        grid = (B.tileRow1 - B.tileRow0) * ctx->tilesX;
        if (grid < 1 || grid > ctx->smCount) return fail(ctx, "hr_calc_flow: a band of %d tiles does not fit the GPU", grid);
    }
    void *args[] = {&P};
    if (pl) {
        /* after: the newest frame (evIn, recorded by the update), the packed copy of the previous frame, the warps
         * that read the flow buffer about to be overwritten (those of the pair before the previous one) */
        CU(cudaStreamWaitEvent(st, ctx->latticeFirst ? ctx->evLattice : ctx->evIn, 0));
        if (ctx->havePack[ctx->packedId[0]]) CU(cudaStreamWaitEvent(st, ctx->evPack[ctx->packedId[0]], 0));
        if (wait_warps(ctx, st, fb)) return 1;
        ctx->nWarpEv[fb] = 0;
        /* the flow buffer was last written HR_FLOW_BUFS pairs back: on this lane when the ring is a multiple of the
         * lanes (stream order); in any case after that search, even when no warp ever read its flow */
        if (ctx->haveSearch[fb]) CU(cudaStreamWaitEvent(st, ctx->evSearch[fb], 0));
    } else if (ctx->sPack) {
        /* the pipeline was used earlier: order this launch after what is left of it */
        if (pipe_join(ctx)) return 1;
    }
    if (ctx->profiling) CU(cudaEventRecord(ctx->evK[0], st));
    /* Which generation? Alone, a launch of the first generation is the shortest (1080p: 39 vs 42 us at R = 5, 59 vs 69 us
     * at R = 16); launches of the third share the SMs (three CTAs per SM, one per search lane), and a stream of pairs
     * goes through 35 % faster with it. So: the third generation while the previous pair's search is still under way (device-resident
     * streams enqueue far ahead of the GPU), the first when this launch will have the GPU to itself (the blocking call
     * sequence of the filter). All generations write the same bits (tests/test_gpu_search2.py). */
    int gen = ctx->searchGen;
    if (gen == 0) {
        gen = 1;
        if (pl && ctx->haveSearch[ctx->flowCur] && cudaEventQuery(ctx->evSearch[ctx->flowCur]) == cudaErrorNotReady) gen = 3;
        cudaGetLastError(); /* cudaErrorNotReady is not sticky, but leave nothing behind */
    }
    int genUsed = 1;
    const void *kfn = ctx->banded ? (searchRadius == 5 ? (const void *)flow_search_band_kernel<5> : (const void *)flow_search_band_kernel<0>) : search_kernel_for(searchRadius, ctx->multiTile, ctx->timelineOn, gen, &genUsed);
    const void *staged = (!ctx->banded && !ctx->multiTile && gen == 2 && ctx->stagedOk && ctx->stagedOn) ? staged_kernel_for(searchRadius, ctx->timelineOn) : NULL;
    ctx->lastSearchStaged = staged != NULL;
    if (staged) {
        /* the packed copy this launch reads, by its physical buffer */
        void *sargs[] = {&P, &ctx->tmapPacked[ctx->packedId[0]]};
        ctx->lastSearchGen = 2;
        CU(cudaLaunchCooperativeKernel(staged, dim3(grid), dim3(HR_THREADS), sargs, HR_ST_BYTES, st));
    } else {
        ctx->lastSearchGen = genUsed;
        CU(cudaLaunchCooperativeKernel(kfn, dim3(grid), dim3(genUsed == 3 ? HR3_THREADS : HR_THREADS), args, 0, st));
    }
    if (ctx->profiling) {
        CU(cudaEventRecord(ctx->evK[1], st));
        ctx->haveSearchT = 1;
    }
    ctx->launches++;
    ctx->flowCur = fb;
    ctx->flowSerial[fb] = ++ctx->flowStamp;
    if (!ctx->banded) {
        ctx->blur = ctx->blurB[fb];
        ctx->blurXY = ctx->blurXYB[fb];
        ctx->lane = lane;
        ctx->off = ctx->offL[lane];
        ctx->T = ctx->TL[lane];
        ctx->partial = ctx->partialL[lane];
    }
    if (ctx->sPack) {
        CU(cudaEventRecord(ctx->evSearch[fb], st));
        ctx->searchSeq[fb] = ++ctx->evSeq;
        ctx->haveSearch[fb] = 1;
        ctx->packRead[ctx->packedId[0]] = ctx->evSearch[fb];
    }
    *stOut = st;
    return 0;
}

extern "C" int hr_calc_flow(HrContext *ctx, int searchRadius, int deltaScalar, int neighborBiasScalar, double *seconds) {
    if (!ctx) return 1;
    if (searchRadius < HR_MIN_SEARCH_RADIUS || searchRadius > HR_MAX_SEARCH_RADIUS)
        return fail(ctx, "hr_calc_flow: search radius %d outside [%d, %d]", searchRadius, HR_MIN_SEARCH_RADIUS, HR_MAX_SEARCH_RADIUS);
    if (deltaScalar < 0 || deltaScalar > 31 || neighborBiasScalar < 0 || neighborBiasScalar > 31)
        return fail(ctx, "hr_calc_flow: scalar out of range");
    if (ctx->banded && searchRadius > ctx->bandMaxRadius)
        return fail(ctx, "hr_calc_flow: search radius %d needs a larger halo than the bands were configured for (hr_band_set_max_radius: %d)", searchRadius, ctx->bandMaxRadius);
    if (bind_device(ctx)) return 1;
    const int pl = pipe_on(ctx);
    if (ctx->specFlow.valid && ctx->specFlow.frames == ctx->framesSeen && ctx->specFlow.R == searchRadius && ctx->specFlow.dS == deltaScalar &&
        ctx->specFlow.nS == neighborBiasScalar) {
        /* hr_update_frame launched exactly this search already (same frame pair, same knobs): nothing to enqueue, and
         * the pair's first output, if it was warped ahead behind that search, stands */
        ctx->specFlow.valid = 0;
    } else {
        ctx->specWarp.valid = 0; /* a new flow: whatever was warped ahead is void */
        ctx->specFlow.valid = 0;
        cudaStream_t st;
        if (launch_flow(ctx, searchRadius, deltaScalar, neighborBiasScalar, &st)) return 1;
        if (!pl || seconds) CU(cudaEventRecord(ctx->evFlowEnd, st));
    }
    ctx->lastFlow.valid = 1;
    ctx->lastFlow.R = searchRadius;
    ctx->lastFlow.dS = deltaScalar;
    ctx->lastFlow.nS = neighborBiasScalar;
    if (seconds) {
        CU(cudaEventSynchronize(ctx->evFlowEnd));
        float ms = 0.f;
        if (ctx->framesSeen > 0 && cudaEventElapsedTime(&ms, ctx->evUpdate, ctx->evFlowEnd) != cudaSuccess) {
            cudaGetLastError(); /* pipelined device updates do not stamp their start */
            ms = 0.f;
        }
        *seconds = (double)ms * 1e-3;
    }
    return 0;
}

__global__ void rcp_pair_kernel(float a, float b, float *out) {
    out[0] = rcp_approx(a);
    out[1] = rcp_approx(b);
}
static int device_rcp(HrContext *ctx, const float *den, float *out) {
    float *d = NULL;
    CU(cudaMalloc(&d, 2 * sizeof(float)));
    rcp_pair_kernel<<<1, 1, 0, ctx->stream>>>(den[0], den[1], d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, "CUDA error in device_rcp: %s", cudaGetErrorString(e));
    return 0;
}

/* (v, v) as one 64-bit word of two fp32 */
static F2 pair_f32(float v) {
    uint32_t b;
    memcpy(&b, &v, 4);
    return ((F2)b << 32) | b;
}

template <typename T>
static void launch_wide16(dim3 grid, dim3 block, cudaStream_t st, const WarpParams<T> &P, const WarpFastArgs &A, const WarpBatch &B);
template <>
void launch_wide16<uint8_t>(dim3 grid, dim3 block, cudaStream_t st, const WarpParams<uint8_t> &P, const WarpFastArgs &A, const WarpBatch &B) {
    warp_fast_kernel<uint8_t, 16><<<grid, block, 0, st>>>(P, A, B);
}
template <>
void launch_wide16<uint16_t>(dim3, dim3, cudaStream_t, const WarpParams<uint16_t> &, const WarpFastArgs &, const WarpBatch &) {} /* 32-byte units are not built */

/* flow colours of the HSV mode for flow buffer fb (hr_warp.cuh): built once per flow, on the stream of the first warp
 * that needs them; later warps of the same flow on other streams wait for evColour[fb] */
static int ensure_colours(HrContext *ctx, int fb, cudaStream_t st) {
    const int ln = ctx->lw * ctx->lh;
    if (!ctx->colours[fb]) {
        CU(cudaMalloc(&ctx->colours[fb], (size_t)ln * sizeof(uint32_t)));
        ctx->deviceBytes += (size_t)ln * sizeof(uint32_t);
        ctx->colourSerial[fb] = ~0ull; /* nothing built yet */
    }
    if (!ctx->evColour[fb]) CU(cudaEventCreateWithFlags(&ctx->evColour[fb], cudaEventDisableTiming));
    if (ctx->colourSerial[fb] != ctx->flowSerial[fb]) {
        flow_colour_kernel<<<(ln + 255) / 256, 256, 0, st>>>(ctx->blurXY, ctx->colours[fb], ln, ctx->s <= 2 ? 4 : 1); /* fb is the current flow buffer */
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->evColour[fb], st));
        ctx->colourSerial[fb] = ctx->flowSerial[fb];
        ctx->launches++;
    } else {
        CU(cudaStreamWaitEvent(st, ctx->evColour[fb], 0));
    }
    return 0;
}

/* K5 for n output frames of the current pair (same mode and levels, blend scalars ts[i], planes outY[i] / outUV[i]):
 * one launch per HR_WARP_BATCH outputs when the fast kernel applies, one per output otherwise */
template <typename T>
static int launch_warp(HrContext *ctx, int n, const float *ts, void *const *outY, void *const *outUV, int mode, float black, float white, int warpStream) {
    WarpParams<T> P;
    P.f1y = (const T *)ctx->fy[0];
    P.f1uv = (const T *)ctx->fuv[0];
    P.f2y = (const T *)ctx->fy[1];
    P.f2uv = (const T *)ctx->fuv[1];
    P.outY = (T *)outY[0];
    P.outUV = (T *)outUV[0];
    P.flow = ctx->blur;
    P.flowXY = ctx->blurXY;
    P.lw = ctx->lw;
    P.lh = ctx->lh;
    P.H = ctx->H;
    P.W = ctx->W;
    P.aW = ctx->aW;
    P.s = ctx->s;
    P.mode = mode;
    P.t12 = ts[0];            /* opticalFlowCalc.c:215-216, float */
    P.t21 = 1.0f - ts[0];
    P.black = black;
    P.white = white;
    int r0 = 0, r1 = ctx->H;
    if (ctx->banded) {
        r0 = ctx->bandRow0[ctx->bandRank];
        r1 = ctx->bandRow1[ctx->bandRank];
    }
    WarpFastArgs A;
    memset(&A, 0, sizeof(A));
    A.white = white;
    /* level constants (hr_warp.cuh: 8-bit = the reference's expressions as NVIDIA OpenCL compiles them — also what the
     * HSV mode of P010 uses; 16-bit = DESIGN.md §P010, correctly rounded reciprocals) */
    const bool is16 = sizeof(T) == 2 && mode != HR_MODE_HSV_FLOW;
    volatile float b16 = black / 255.0f, w16 = white / 255.0f;
    b16 = b16 * 65472.0f;
    w16 = w16 * 65472.0f;
    const float outMax = is16 ? 65472.0f : 255.0f, inMax = is16 ? 65535.0f : 255.0f, mid = is16 ? 32768.0f : 128.0f;
    A.sub[0] = is16 ? (float)b16 : black;
    A.sub[1] = mid;
    A.den[0] = is16 ? (float)w16 - (float)b16 : white - black;
    A.den[1] = is16 ? (float)w16 : white;
    int denOk = 1;
    for (int c = 0; c < 2; ++c) denOk = denOk && fabsf(A.den[c]) >= 1.17549435e-38f && fabsf(A.den[c]) <= 8.50705917e37f;
    if (denOk) {
        if (!is16 && (!ctx->haveRcp8 || ctx->rcp8Den[0] != A.den[0] || ctx->rcp8Den[1] != A.den[1])) {
            /* MUFU.RCP of the two denominators, read back once per change of the level knobs */
            if (device_rcp(ctx, A.den, ctx->rcp8)) return 1;
            ctx->rcp8Den[0] = A.den[0];
            ctx->rcp8Den[1] = A.den[1];
            ctx->haveRcp8 = 1;
        }
        for (int c = 0; c < 2; ++c) {
            volatile float rc = is16 ? 1.0f / A.den[c] : ctx->rcp8[c];
            A.rcp[c] = rc;
            A.subIsInt[c] = A.sub[c] >= 0.0f && A.sub[c] < 4194304.0f && floorf(A.sub[c]) == A.sub[c];
            /* the map is monotonic: it stays inside the output range iff its two ends do */
            volatile float lo = (0.0f - A.sub[c]) * rc, hi = (inMax - A.sub[c]) * rc;
            if (c == 0) {
                lo = lo * outMax;
                hi = hi * outMax;
            } else {
                lo = fmaf(lo, outMax, mid);
                hi = fmaf(hi, outMax, mid);
            }
            const float mn = fminf(lo, hi), mx = fmaxf(lo, hi);
            A.clampNeeded[c] = !(mn >= 0.0f && mx < (is16 ? 65536.0f : 256.0f));
            /* the same constants as fp32 pairs (hr_warp_fast.cuh, blend_pair) */
            volatile float msub = HR_MAGIC + A.sub[c];
            A.lc[c].rcp = pair_f32(rc);
            A.lc[c].mul = pair_f32(outMax);
            A.lc[c].add = pair_f32(mid);
            A.lc[c].negSub = pair_f32(-A.sub[c]);
            A.lc[c].negMsub = pair_f32(-msub);
            A.lc[c].lo = 0.0f;
            A.lc[c].hi = outMax;
        }
    }
    A.negM = pair_f32(-HR_MAGIC);
    A.M = pair_f32(HR_MAGIC);
    /* The unit of the fast kernel: as wide as a lattice cell, at most 16 bytes per row; every plane and the row pitch
     * aligned to it, blend scalars in [0,1], level denominators for which div.full.f32 is MUFU.RCP * x (hr_warp.cuh).
     * A frame that misses the wide unit's alignment falls back to the 4-sample unit, then to the per-sample kernel. */
    uintptr_t align = (uintptr_t)P.f1y | (uintptr_t)P.f1uv | (uintptr_t)P.f2y | (uintptr_t)P.f2uv;
    int tsOk = 1;
    for (int i = 0; i < n; ++i) {
        align |= (uintptr_t)outY[i] | (uintptr_t)outUV[i];
        tsOk = tsOk && ts[i] >= 0.0f && ts[i] <= 1.0f;
    }
    int uw = ctx->s == 2 ? 4 : (ctx->s == 3 ? 8 : (sizeof(T) == 2 ? 8 : 16));
    const char *uwEnv = getenv("HR_WARP_UNIT"); /* developer knob: force a narrower unit */
    if (uwEnv && atoi(uwEnv) >= 4 && atoi(uwEnv) < uw) uw = atoi(uwEnv) >= 8 ? 8 : 4;
    for (; uw >= 4; uw >>= 1) {
        const uintptr_t ub = (uintptr_t)uw * sizeof(T);
        if (ctx->W % uw == 0 && ctx->aW >= 2 * uw && align % ub == 0) break;
    }
    /* thread = uw samples x ROWS rows (cells taller than that are covered by several threads: measured, smaller units
     * and more resident warps beat fewer flow look-ups per sample); row groups of the luma plane, then of the chroma
     * plane. Band boundaries are multiples of 2^(s+1) rows, so row groups do not straddle them. */
    const int ROWS = warp_rows((uw >= 4 ? uw : 4) * (int)sizeof(T));
    A.lumaG0 = r0 / ROWS;
    A.lumaGroups = (r1 + ROWS - 1) / ROWS - A.lumaG0;
    A.chromaG0 = (r0 >> 1) / ROWS;
    A.chromaGN = ((r1 >> 1) + ROWS - 1) / ROWS - A.chromaG0;
    const int groups = A.lumaGroups + A.chromaGN;
    const int fast = ctx->useFastWarp && ctx->s >= 2 && uw >= 4 && mode != HR_MODE_SIDE_BY_SIDE_2 && tsOk && denOk && ctx->H >= 2 * (ROWS + 2);
    /* pipelined: warps into caller-owned planes are independent of one another -> round-robin over the warp
     * streams, each after the search that produced the flow; warps into the internal output frame stay on the
     * main stream (the download that follows is ordered there) */
    const int pl = pipe_on(ctx), fb = ctx->flowCur;
    cudaStream_t st = ctx->stream;
    if (pl) {
        if (warpStream == 2) st = ctx->sWarp[0]; /* warped ahead: one stream, so that they never overlap one another */
        else if (warpStream == 1) st = ctx->sWarp[ctx->warpRR++ % HR_WARP_STREAMS];
        if (ctx->haveSearch[fb]) CU(cudaStreamWaitEvent(st, ctx->evSearch[fb], 0));
        else CU(cudaStreamWaitEvent(st, ctx->evIn, 0));
    }
    if (fast && mode == HR_MODE_HSV_FLOW) {
        if (ensure_colours(ctx, fb, st)) return 1;
        A.colours = ctx->colours[fb];
    }
    if (ctx->profiling) CU(cudaEventRecord(ctx->evK[2], st));
    if (fast) {
        for (int i0 = 0; i0 < n; i0 += HR_WARP_BATCH) {
            WarpBatch B;
            memset(&B, 0, sizeof(B));
            B.n = n - i0 < HR_WARP_BATCH ? n - i0 : HR_WARP_BATCH;
            for (int i = 0; i < B.n; ++i) {
                volatile float t12 = ts[i0 + i], t21 = 1.0f - t12, mt = HR_MAGIC * t12;
                B.t12[i] = t12;
                B.t21[i] = t21;
                B.t12x2[i] = pair_f32(t12);
                B.t21x2[i] = pair_f32(t21);
                B.negMt12[i] = pair_f32(-mt);
                B.outY[i] = outY[i0 + i];
                B.outUV[i] = outUV[i0 + i];
            }
            dim3 block(32, 4);
            A.unitsX = (ctx->aW + uw - 1) / uw;
            A.edgeBlocks = (2 * groups * ROWS + 127) / 128; /* the two edge columns as single-row units */
            A.coreBlocksX = A.unitsX > 2 ? (A.unitsX - 2 + 31) / 32 : 0;
            dim3 grid(A.edgeBlocks + A.coreBlocksX * ((groups + 3) / 4), 1, B.n);
            if (uw == 4) warp_fast_kernel<T, 4><<<grid, block, 0, st>>>(P, A, B);
            else if (uw == 8) warp_fast_kernel<T, 8><<<grid, block, 0, st>>>(P, A, B);
            else launch_wide16<T>(grid, block, st, P, A, B);
            ctx->launches++;
        }
    } else {
        /* per-sample kernel, row groups of 4 */
        const int lg0 = r0 / 4, lgn = (r1 + 3) / 4 - lg0, cg0 = (r0 >> 1) / 4, cgn = ((r1 >> 1) + 3) / 4 - cg0;
        dim3 block(32, 8);
        dim3 grid((ctx->aW + 127) / 128, (lgn + cgn + 7) / 8);
        for (int i = 0; i < n; ++i) {
            P.t12 = ts[i];
            P.t21 = 1.0f - ts[i];
            P.outY = (T *)outY[i];
            P.outUV = (T *)outUV[i];
            warp_generic_kernel<T><<<grid, block, 0, st>>>(P, lgn, lg0, cg0, cgn);
            ctx->launches++;
        }
    }
    CU(cudaGetLastError());
    if (ctx->profiling) {
        CU(cudaEventRecord(ctx->evK[3], st));
        ctx->haveWarpT = 1;
    }
    if (pl) {
        if (ctx->nWarpEv[fb] == HR_MAX_WARP_EVENTS) { /* fold the recorded readers into one event */
            if (wait_warps(ctx, st, fb)) return 1;
            ctx->nWarpEv[fb] = 0;
        }
        ctx->warpSeq[fb][ctx->nWarpEv[fb]] = ++ctx->evSeq;
        CU(cudaEventRecord(ctx->evWarp[fb][ctx->nWarpEv[fb]++], st));
    }
    return 0;
}
/* which stream a warp launch goes to in pipelined mode: internal output frame -> main stream (0), caller-owned planes ->
 * round robin (1), the frame warped ahead -> its own stream (2) */
static int warp_stream_kind(const HrContext *ctx) {
    if (ctx->outY == ctx->outBuf2) return 2;
    return ctx->outY != ctx->outBuf ? 1 : 0;
}
static int launch_warp_one(HrContext *ctx, float t, int mode, float black, float white) {
    void *oy = ctx->outY, *ouv = ctx->outUV;
    const int kind = warp_stream_kind(ctx);
    return ctx->bps == 1 ? launch_warp<uint8_t>(ctx, 1, &t, &oy, &ouv, mode, black, white, kind) : launch_warp<uint16_t>(ctx, 1, &t, &oy, &ouv, mode, black, white, kind);
}

extern "C" int hr_warp(HrContext *ctx, float t, int mode, float black, float white) {
    if (!ctx) return 1;
    if (t > 1.0f) { /* opticalFlowCalc.c:209-212 */
        printf("Error: Blending scalar is greater than 1.0\n");
        return fail(ctx, "hr_warp: blending scalar %f is greater than 1.0", (double)t);
    }
    if (mode < 0 || mode > 6) return fail(ctx, "hr_warp: unknown output mode %d", mode);
    if (bind_device(ctx)) return 1;
    if (!ctx->pipeline && ctx->sPack && pipe_join(ctx)) return 1; /* leftovers of an earlier pipelined phase */
    CU(cudaEventRecord(ctx->evWarpStart, ctx->stream));
    const int internal = ctx->outY == ctx->outBuf;
    if (internal) {
        /* history for hr_download's guess of the next blending scalar (vf_HopperRender.c:371-374: t advances by a
         * constant ratio within a source frame) */
        if (ctx->lastWarpFrames == ctx->framesSeen && t > ctx->lastT) ctx->lastDelta = t - ctx->lastT;
        if (!(ctx->lastWarpFrames == ctx->framesSeen && memcmp(&t, &ctx->lastT, sizeof(float)) == 0)) ctx->pace->observe(t); /* (the same frame warped again, another mode: not an output of the pacing) */
        ctx->lastT = t;
        ctx->lastWarpFrames = ctx->framesSeen;
        ctx->lastMode = mode;
        ctx->lastBlack = black;
        ctx->lastWhite = white;
        if (ctx->specWarp.valid) {
            const HrContext::WarpAhead w = ctx->specWarp;
            ctx->specWarp.valid = 0;
            if (w.frames == ctx->framesSeen && w.epoch == ctx->epoch && w.mode == mode && memcmp(&w.t, &t, sizeof(float)) == 0 && w.black == black &&
                w.white == white) {
                /* hr_download warped exactly this frame ahead into the second internal frame: make it the output */
                uint8_t *b = ctx->outBuf;
                ctx->outBuf = ctx->outBuf2;
                ctx->outBuf2 = b;
                ctx->outY = ctx->outBuf;
                ctx->outUV = ctx->outBuf + (size_t)ctx->H * ctx->W * ctx->bps;
                CU(cudaStreamWaitEvent(ctx->stream, w.done, 0));
                return 0;
            }
        }
    }
    return launch_warp_one(ctx, t, mode, black, white);
}

/* hr_download, after its copy has been enqueued: warp the frame the pacing will most likely ask for next */
static int warp_ahead_launch(HrContext *ctx, float tp) {
    const int fb = ctx->flowCur;
    void *keepY = ctx->outY, *keepUV = ctx->outUV;
    ctx->outY = ctx->outBuf2;
    ctx->outUV = ctx->outBuf2 + (size_t)ctx->H * ctx->W * ctx->bps;
    const int rc = launch_warp_one(ctx, tp, ctx->lastMode, ctx->lastBlack, ctx->lastWhite);
    ctx->outY = keepY;
    ctx->outUV = keepUV;
    if (rc) return 1;
    ctx->specWarp.valid = 1;
    ctx->specWarp.t = tp;
    ctx->specWarp.mode = ctx->lastMode;
    ctx->specWarp.black = ctx->lastBlack;
    ctx->specWarp.white = ctx->lastWhite;
    ctx->specWarp.frames = ctx->framesSeen;
    ctx->specWarp.epoch = ctx->epoch;
    ctx->specWarp.done = ctx->evWarp[fb][ctx->nWarpEv[fb] - 1];
    return 0;
}
static int warp_ahead(HrContext *ctx) {
    if (!pipe_on(ctx) || !ctx->aheadOn || ctx->outY != ctx->outBuf || !ctx->outBuf2 || ctx->specWarp.valid) return 0;
    if (ctx->lastWarpFrames != ctx->framesSeen) return 0;
    float tp = 0.0f;
    bool nextFrame = false;
    ctx->firstAhead.valid = 0;
    if (!ctx->pace->predict(&tp, &nextFrame)) return 0;
    if (nextFrame) { /* the next source frame comes first: its update warps this one behind the search it starts */
        ctx->firstAhead.valid = 1;
        ctx->firstAhead.frames = ctx->framesSeen + 1;
        ctx->firstAhead.t = tp;
        return 0;
    }
    return warp_ahead_launch(ctx, tp);
}
/* hr_update_frame, after it has started the new pair's search and the pack: the pair's first output, as hr_download
 * guessed it, behind the search and the last row of the upload */
static int warp_ahead_first(HrContext *ctx) {
    const int want = ctx->firstAhead.valid && ctx->firstAhead.frames == ctx->framesSeen;
    ctx->firstAhead.valid = 0;
    if (!want || !pipe_on(ctx) || !ctx->aheadOn || ctx->outY != ctx->outBuf || !ctx->outBuf2 || ctx->specWarp.valid) return 0;
    if (!ctx->specFlow.valid || ctx->specFlow.frames != ctx->framesSeen) return 0; /* no search was started ahead */
    CU(cudaStreamWaitEvent(ctx->sWarp[0], ctx->evIn, 0)); /* the search may have started behind the lattice rows alone */
    return warp_ahead_launch(ctx, ctx->firstAhead.t);
}

/* One source frame of a device-resident stream in ONE call: update + flow + nWarps warps, each into its own
 * caller-owned planes. Enqueue-only. */
extern "C" int hr_step_device(HrContext *ctx, const void *dY, const void *dUV, int borrow, int searchRadius, int deltaScalar, int neighborBiasScalar,
                              int nWarps, const float *blendingScalars, int frameOutputMode, float blackLevel, float whiteLevel, void *const *outY,
                              void *const *outUV) {
    if (!ctx) return 1;
    if (nWarps < 0 || (nWarps > 0 && (!blendingScalars || !outY || !outUV))) return fail(ctx, "hr_step_device: bad warp list");
    if (hr_update_frame_device(ctx, dY, dUV, borrow)) return 1;
    if (ctx->framesSeen < 2) return 0; /* the first frame of a stream has no partner yet (vf_HopperRender.c:490-495) */
    if (hr_calc_flow(ctx, searchRadius, deltaScalar, neighborBiasScalar, NULL)) return 1;
    return hr_warp_batch(ctx, nWarps, blendingScalars, frameOutputMode, blackLevel, whiteLevel, outY, outUV);
}

/* n output frames of the current pair into caller-owned device planes, one launch (include/hopperrender_cuda.h) */
extern "C" int hr_warp_batch(HrContext *ctx, int nWarps, const float *blendingScalars, int frameOutputMode, float blackLevel, float whiteLevel,
                             void *const *outY, void *const *outUV) {
    if (!ctx) return 1;
    if (nWarps < 0 || (nWarps > 0 && (!blendingScalars || !outY || !outUV))) return fail(ctx, "hr_warp_batch: bad warp list");
    if (frameOutputMode < 0 || frameOutputMode > 6) return fail(ctx, "hr_warp_batch: unknown output mode %d", frameOutputMode);
    for (int i = 0; i < nWarps; ++i) {
        if (blendingScalars[i] > 1.0f) return fail(ctx, "hr_warp_batch: blending scalar %f is greater than 1.0", (double)blendingScalars[i]); /* opticalFlowCalc.c:209-212 */
        if (!outY[i] || !outUV[i]) return fail(ctx, "hr_warp_batch: NULL output plane");
    }
    if (nWarps == 0) return 0;
    if (bind_device(ctx)) return 1;
    if (!ctx->pipeline && ctx->sPack && pipe_join(ctx)) return 1; /* leftovers of an earlier pipelined phase */
    /* developer knob HR_WARP_SPLIT=k: at most k outputs per launch (each launch on the next warp stream) */
    static int split = -1;
    if (split < 0) {
        const char *e = getenv("HR_WARP_SPLIT");
        split = e ? atoi(e) : 0;
    }
    const int per = split > 0 ? split : nWarps;
    for (int i = 0; i < nWarps; i += per) {
        const int n = nWarps - i < per ? nWarps - i : per;
        if (ctx->bps == 1 ? launch_warp<uint8_t>(ctx, n, blendingScalars + i, outY + i, outUV + i, frameOutputMode, blackLevel, whiteLevel, 1)
                          : launch_warp<uint16_t>(ctx, n, blendingScalars + i, outY + i, outUV + i, frameOutputMode, blackLevel, whiteLevel, 1))
            return 1;
    }
    return 0;
}

/* nSteps consecutive source frames in one call (hr_step_device in a loop): warps of step i use
 * blendingScalars[first .. first + nWarps[i]) and the output planes at the same indices, first = sum of nWarps[0..i). */
extern "C" int hr_steps_device(HrContext *ctx, int nSteps, const void *const *dY, const void *const *dUV, int borrow, int searchRadius, int deltaScalar,
                               int neighborBiasScalar, const int *nWarps, const float *blendingScalars, int frameOutputMode, float blackLevel,
                               float whiteLevel, void *const *outY, void *const *outUV) {
    if (!ctx) return 1;
    if (nSteps < 0 || (nSteps > 0 && (!dY || !dUV || !nWarps))) return fail(ctx, "hr_steps_device: bad frame list");
    int first = 0;
    for (int i = 0; i < nSteps; ++i) {
        if (hr_step_device(ctx, dY[i], dUV[i], borrow, searchRadius, deltaScalar, neighborBiasScalar, nWarps[i], blendingScalars ? blendingScalars + first : NULL,
                           frameOutputMode, blackLevel, whiteLevel, outY ? outY + first : NULL, outUV ? outUV + first : NULL))
            return 1;
        first += nWarps[i];
    }
    return 0;
}

extern "C" int hr_download(HrContext *ctx, void *yPlane, void *uvPlane, double *seconds) {
    if (!ctx) return 1;
    if (!yPlane || !uvPlane) return fail(ctx, "hr_download: NULL plane");
    if (bind_device(ctx)) return 1;
    const size_t ylen = (size_t)ctx->H * ctx->W * ctx->bps, uvlen = (size_t)(ctx->H / 2) * ctx->W * ctx->bps;
    int staged = 0;
    if (ctx->stageThreads >= 1 && (host_is_pageable(yPlane) || host_is_pageable(uvPlane))) {
        staged = stage_ready(ctx);
        if (staged < 0) return 1;
    }
    if (staged) {
        /* pageable planes: device -> pinned ring by the copy engine, ring -> the caller's planes by the crew; the
         * next frame is warped ahead while the crew copies */
        const int oneRun = (uint8_t *)uvPlane == (uint8_t *)yPlane + ylen && (uint8_t *)ctx->outUV == (uint8_t *)ctx->outY + ylen;
        StageDown d;
        const size_t run = oneRun ? ylen + uvlen : ylen;
        if (staged_d2h_begin(ctx, (uint8_t *)yPlane, (const uint8_t *)ctx->outY, run, &d)) return 1;
        int failed = 0;
        if (oneRun && ctx->plan->n <= HR_STAGE_SLOTS) failed = warp_ahead(ctx); /* the whole frame is with the copy engine already */
        if (staged_d2h_end(ctx, (const uint8_t *)ctx->outY, &d) || failed) return 1;
        if (!oneRun) {
            if (staged_d2h_begin(ctx, (uint8_t *)uvPlane, (const uint8_t *)ctx->outUV, uvlen, &d)) return 1;
            if (staged_d2h_end(ctx, (const uint8_t *)ctx->outUV, &d)) return 1;
        }
        g_d2hBytes += ylen + uvlen;
        CU(cudaEventRecord(ctx->evDlEnd, ctx->stream));
        if (warp_ahead(ctx)) return 1; /* no-op when it was started above */
        CU(cudaEventSynchronize(ctx->evDlEnd));
        if (seconds) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->evWarpStart, ctx->evDlEnd) != cudaSuccess) {
                cudaGetLastError();
                ms = 0.f;
            }
            *seconds = (double)ms * 1e-3;
        }
        return 0;
    }
    if ((uint8_t *)uvPlane == (uint8_t *)yPlane + ylen && (uint8_t *)ctx->outUV == (uint8_t *)ctx->outY + ylen) {
        CU(cudaMemcpyAsync(yPlane, ctx->outY, ylen + uvlen, cudaMemcpyDeviceToHost, ctx->stream)); /* back-to-back planes: one transfer */
    } else {
        CU(cudaMemcpyAsync(yPlane, ctx->outY, ylen, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(uvPlane, ctx->outUV, uvlen, cudaMemcpyDeviceToHost, ctx->stream));
    }
    g_d2hBytes += ylen + uvlen;
    CU(cudaEventRecord(ctx->evDlEnd, ctx->stream));
    if (warp_ahead(ctx)) return 1;
    CU(cudaEventSynchronize(ctx->evDlEnd));
    if (seconds) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->evWarpStart, ctx->evDlEnd) != cudaSuccess) {
            cudaGetLastError(); /* no warp was recorded yet */
            ms = 0.f;
        }
        *seconds = (double)ms * 1e-3;
    }
    return 0;
}


/* ---------------------------------------------------------------------------------------------------
 * Spatial bands: N contexts, one per GPU, each owning a band of rows of every frame.
 * A band owner uploads, warps and downloads only its rows; the rows of the other bands are pulled from
 * the peers' frame slots over NVLink P2P (peer-mapped pointers, in-process or through CUDA IPC), so
 * that every GPU holds the whole frame pair for the (replicated, bit-identical) flow search.
 * Hand-off between GPUs: every context has a two-counter mailbox in device memory — frames whose band
 * is uploaded, frames whose gather is complete — written by a one-thread kernel on the owner's stream
 * (st.release.sys) and polled by a one-thread kernel on the reader's stream (ld.acquire.sys). No NCCL,
 * no host synchronisation on the data path.
 * ------------------------------------------------------------------------------------------------- */
__global__ void band_signal_kernel(unsigned long long *p, unsigned long long v) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__global__ void band_wait_kernel(const unsigned long long *p, unsigned long long target) {
    unsigned long long v;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if (v < target) __nanosleep(200);
    } while (v < target);
}

/* rows [lo, hi) of every frame that rank r needs besides its own band: what its search (frame-1 samples of its tiles
 * at the largest accumulated offset) and its warp (displaced source rows) can reach, even bounds, inside the frame */
static void band_reach(const HrContext *ctx, int r, int *lo, int *hi) {
    const int R = ctx->bandMaxRadius;
    const int neg = (R / 2) * (R / 2) * ctx->iters, pos = (R - 1 - R / 2) * (R - 1 - R / 2) * ctx->iters; /* calcDeltaSumsKernel.cl:68-72, one shift per level */
    const int reach = (neg > pos ? neg : pos) + 2;
    int a = ctx->bandRow0[r] - reach, b = ctx->bandRow1[r] + reach;
    a = a < 0 ? 0 : a & ~1;
    b = b > ctx->H ? ctx->H : (b + 1) & ~1;
    if (b > ctx->H) b = ctx->H;
    *lo = a;
    *hi = b;
}
static size_t arena_align(size_t v) { return (v + 255) & ~(size_t)255; }

/* The halo of one frame in ONE launch: wait until every owner has uploaded its band of the frame (mailbox counter),
 * copy the halo rows out of the owners' frame slots (peer-mapped 128-bit loads over NVLink), and — the last CTA to
 * finish — tell the peers that this GPU no longer reads their slots of this frame. */
#define HR_HALO_MAX_SEGS (2 * HR_MAX_BANDS)
struct HaloJob {
    int nSeg;
    const uint4 *src[HR_HALO_MAX_SEGS];
    uint4 *dst[HR_HALO_MAX_SEGS];
    unsigned long long units[HR_HALO_MAX_SEGS];           /* 16-byte units of the segment                     */
    const unsigned long long *uploaded[HR_HALO_MAX_SEGS]; /* the owner's "frames uploaded" counter              */
    unsigned long long frame;                             /* wait for uploaded >= frame                         */
    unsigned long long *gathered;                         /* own "frames fetched" counter                       */
    unsigned int *exitCount;
};
__global__ void __launch_bounds__(256) band_halo_kernel(const HaloJob J) {
    if (threadIdx.x < J.nSeg) {
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(J.uploaded[threadIdx.x]) : "memory");
        } while (v < J.frame);
    }
    __syncthreads();
    for (int sgm = 0; sgm < J.nSeg; ++sgm) {
        const uint4 *src = J.src[sgm];
        uint4 *dst = J.dst[sgm];
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < J.units[sgm]; i += (unsigned long long)gridDim.x * blockDim.x)
            dst[i] = src[i];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(J.exitCount, 1u);
        if (prev == gridDim.x - 1) {
            *J.exitCount = 0u;
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(J.gathered), "l"(J.frame) : "memory");
        }
    }
}

extern "C" int hr_band_configure(HrContext *ctx, int rank, int world, const int *row0, const int *row1) {
    if (!ctx || !row0 || !row1) return 1;
    if (world < 1 || world > HR_MAX_BANDS || rank < 0 || rank >= world) return fail(ctx, "hr_band_configure: rank %d / world %d out of range (max %d bands)", rank, world, HR_MAX_BANDS);
    if (bind_device(ctx)) return 1;
    /* a band is a whole number of lattice tile rows (32 lattice rows = 32 << s frame rows): the search is split by tiles */
    const int unit = HR_TILE << ctx->s;
    int expect = 0;
    for (int r = 0; r < world; ++r) {
        if (row0[r] != expect || row1[r] <= row0[r] || (r + 1 < world && row1[r] % unit != 0))
            return fail(ctx, "hr_band_configure: band %d = [%d, %d) must continue the previous band and end on a multiple of %d rows", r, row0[r], row1[r], unit);
        expect = row1[r];
        ctx->bandRow0[r] = row0[r];
        ctx->bandRow1[r] = row1[r];
    }
    if (expect != ctx->H) return fail(ctx, "hr_band_configure: the bands cover %d rows, the frame has %d", expect, ctx->H);
    if ((row1[rank] - row0[rank] + unit - 1) / unit * ctx->tilesX > ctx->smCount)
        return fail(ctx, "hr_band_configure: band %d has more lattice tiles than the GPU has SMs", rank);
    /* a (re)configuration restarts the frame count: every rank of the group reconfigures together, counters and tags
     * start from zero again once nothing of the old stream is in flight */
    if (sync_all(ctx)) return 1;
    if (ctx->pipeline) {
        ctx->pipeline = 0;
        ctx->specWarp.valid = ctx->specFlow.valid = 0;
    }
    if (!ctx->arena) {
        const size_t ln = (size_t)ctx->lw * ctx->lh;
        size_t o = 0;
        ctx->aT = o;        o = arena_align(o + (size_t)(ctx->tWords ? ctx->tWords : 32) * 8);
        ctx->aPartial = o;  o = arena_align(o + (size_t)(ctx->bigWords ? ctx->bigWords : 32) * 8);
        ctx->aOff = o;      o = arena_align(o + 2 * ln * sizeof(int16_t));
        ctx->aBlur = o;     o = arena_align(o + 2 * ln * sizeof(int16_t));
        ctx->aBlurXY = o;   o = arena_align(o + ln * sizeof(uint32_t));
        ctx->aReady = o;    o = arena_align(o + HR_MAX_BANDS * 8);
        ctx->aDone = o;     o = arena_align(o + HR_MAX_BANDS * 8);
        ctx->aExit = o;     o = arena_align(o + 16); /* exit counters: search launch, halo launch */
        ctx->aMail = o;     o = arena_align(o + 256);
        ctx->arenaBytes = o;
        CU(cudaMalloc(&ctx->arena, o));
        ctx->deviceBytes += o;
        ctx->plainT = ctx->T;
        ctx->plainPartial = ctx->partial;
        ctx->plainOff = ctx->off;
        ctx->plainBlur = ctx->blur;
        ctx->plainBlurXY = ctx->blurXY;
    }
    CU(cudaMemset(ctx->arena, 0, ctx->arenaBytes));
    CU(cudaDeviceSynchronize());
    ctx->epoch = 0; /* the group counts its searches together */
    ctx->T = (unsigned long long *)(ctx->arena + ctx->aT);
    ctx->partial = (unsigned long long *)(ctx->arena + ctx->aPartial);
    ctx->off = (int16_t *)(ctx->arena + ctx->aOff);
    ctx->blur = (int16_t *)(ctx->arena + ctx->aBlur);
    ctx->blurXY = (uint32_t *)(ctx->arena + ctx->aBlurXY);
    ctx->mail = (unsigned long long *)(ctx->arena + ctx->aMail);
    ctx->flowCur = 0;
    ctx->flowSerial[0] = ++ctx->flowStamp;
    ctx->banded = 1;
    ctx->bandWorld = world;
    ctx->bandRank = rank;
    ctx->bandFrames = 0;
    ctx->bandPending = 0;
    if (ctx->bandMaxRadius < HR_MIN_SEARCH_RADIUS) ctx->bandMaxRadius = 16; /* MAX_SEARCH_RADIUS, config.h:7 */
    band_reach(ctx, rank, &ctx->haloLo, &ctx->haloHi);
    for (int r = 0; r < HR_MAX_BANDS; ++r) {
        ctx->peerSlot[r][0] = ctx->peerSlot[r][1] = NULL;
        ctx->peerMail[r] = NULL;
        ctx->peerArena[r] = NULL;
    }
    ctx->peerSlot[rank][0] = ctx->frameBuf[0];
    ctx->peerSlot[rank][1] = ctx->frameBuf[1];
    ctx->peerArena[rank] = ctx->arena;
    ctx->peerMail[rank] = ctx->mail;
    return 0;
}

/* Largest search radius the group will use: it sizes the halo (rows of the neighbouring bands every GPU holds).
 * Default 16 = MAX_SEARCH_RADIUS (config.h:7): -512 / +392 rows; radius 5: +-32 rows. Same value on every rank,
 * before the first frame. */
extern "C" int hr_band_set_max_radius(HrContext *ctx, int searchRadius) {
    if (!ctx) return 1;
    if (!ctx->banded) return fail(ctx, "hr_band_set_max_radius: configure the bands first");
    if (searchRadius < HR_MIN_SEARCH_RADIUS || searchRadius > HR_MAX_SEARCH_RADIUS) return fail(ctx, "hr_band_set_max_radius: radius %d out of range", searchRadius);
    if (ctx->bandFrames) return fail(ctx, "hr_band_set_max_radius: the stream has started");
    ctx->bandMaxRadius = searchRadius;
    band_reach(ctx, ctx->bandRank, &ctx->haloLo, &ctx->haloHi);
    return 0;
}

/* The three device allocations a peer needs to see: frame slot 0, frame slot 1, exchange arena. */
extern "C" int hr_band_local_pointers(HrContext *ctx, void **slot0, void **slot1, void **arena) {
    if (!ctx || !ctx->arena) return 1;
    if (slot0) *slot0 = ctx->frameBuf[0];
    if (slot1) *slot1 = ctx->frameBuf[1];
    if (arena) *arena = ctx->arena;
    return 0;
}
extern "C" int hr_band_export_ipc(HrContext *ctx, unsigned char *handles /* 3 x HR_IPC_HANDLE_BYTES */) {
    if (!ctx || !handles) return 1;
    if (!ctx->arena) return fail(ctx, "hr_band_export_ipc: configure the bands first");
    if (bind_device(ctx)) return 1;
    static_assert(sizeof(cudaIpcMemHandle_t) == HR_IPC_HANDLE_BYTES, "IPC handle size");
    void *ptrs[3] = {ctx->frameBuf[0], ctx->frameBuf[1], ctx->arena};
    for (int i = 0; i < 3; ++i) {
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, ptrs[i]));
        memcpy(handles + i * HR_IPC_HANDLE_BYTES, &h, HR_IPC_HANDLE_BYTES);
    }
    return 0;
}
extern "C" int hr_band_open_ipc(HrContext *ctx, const unsigned char *handles, void **slot0, void **slot1, void **arena) {
    if (!ctx || !handles || !slot0 || !slot1 || !arena) return 1;
    if (bind_device(ctx)) return 1;
    void **outs[3] = {slot0, slot1, arena};
    for (int i = 0; i < 3; ++i) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + i * HR_IPC_HANDLE_BYTES, HR_IPC_HANDLE_BYTES);
        CU(cudaIpcOpenMemHandle(outs[i], h, cudaIpcMemLazyEnablePeerAccess));
    }
    return 0;
}
/* pointers must be valid in this process and on this device (same process: enable peer access first —
 * done here when peerDevice >= 0; another process: hr_band_open_ipc) */
extern "C" int hr_band_connect(HrContext *ctx, int peerRank, int peerDevice, void *slot0, void *slot1, void *arena) {
    if (!ctx) return 1;
    if (peerRank < 0 || peerRank >= ctx->bandWorld || peerRank == ctx->bandRank) return fail(ctx, "hr_band_connect: bad peer rank %d", peerRank);
    if (!slot0 || !slot1 || !arena) return fail(ctx, "hr_band_connect: NULL pointer");
    if (bind_device(ctx)) return 1;
    if (peerDevice >= 0 && peerDevice != ctx->device) {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, ctx->device, peerDevice));
        if (!can) return fail(ctx, "hr_band_connect: device %d cannot access device %d", ctx->device, peerDevice);
        cudaError_t e = cudaDeviceEnablePeerAccess(peerDevice, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
    }
    ctx->peerSlot[peerRank][0] = (unsigned char *)slot0;
    ctx->peerSlot[peerRank][1] = (unsigned char *)slot1;
    ctx->peerArena[peerRank] = (unsigned char *)arena;
    ctx->peerMail[peerRank] = (unsigned long long *)((unsigned char *)arena + ctx->aMail);
    return 0;
}

/* does rank a hold rows of rank b's band in its halo (= does a read b's frame slots)? */
static int band_reads(const HrContext *ctx, int a, int b) {
    if (a == b) return 0;
    int lo, hi;
    band_reach(ctx, a, &lo, &hi);
    return lo < ctx->bandRow1[b] && hi > ctx->bandRow0[b];
}

/* Phase 1 of a banded updateFrame: upload (or copy from device memory) the rows of this context's band
 * of the new frame and tell the peers. yBand / uvBand point at the band's first luma row / first chroma
 * row. Enqueue-only for device sources; host sources are copied asynchronously (pinned memory advised). */
extern "C" int hr_band_upload(HrContext *ctx, const void *yBand, const void *uvBand, int sourceIsDevice) {
    if (!ctx) return 1;
    if (!ctx->banded) return fail(ctx, "hr_band_upload: bands are not configured");
    if (ctx->bandPending) return fail(ctx, "hr_band_upload: the previous frame was not gathered (hr_band_gather)");
    if (!yBand || !uvBand) return fail(ctx, "hr_band_upload: NULL plane");
    for (int r = 0; r < ctx->bandWorld; ++r)
        if (!ctx->peerMail[r]) return fail(ctx, "hr_band_upload: peer %d is not connected", r);
    if (bind_device(ctx)) return 1;
    CU(cudaEventRecord(ctx->evUpdate, ctx->stream));
    int slot;
    rotate_slots(ctx, &slot);
    const unsigned long long n = ctx->bandFrames; /* this is frame n; frame n-2 lived in the same slot */
    if (n >= 2 && !ctx->bandSearched) {
        /* The slot about to be overwritten held frame n-2, whose halo rows the neighbours fetched from here. A search of
         * the group since then already proves that they are done (every GPU enters a pair's search after it has fetched
         * that pair's halos, and no search ends before all have entered); without one, ask their counters. */
        for (int r = 0; r < ctx->bandWorld; ++r) {
            if (!band_reads(ctx, r, ctx->bandRank)) continue;
            band_wait_kernel<<<1, 1, 0, ctx->stream>>>(ctx->peerMail[r] + 1, n - 1); /* peer r has fetched its halo of frame n-2 */
            ctx->launches++;
        }
    }
    ctx->bandSearched = 0;
    unsigned char *dst = ctx->frameBuf[slot];
    const size_t rowBytes = (size_t)ctx->W * ctx->bps, ylen = (size_t)ctx->H * rowBytes;
    const int r0 = ctx->bandRow0[ctx->bandRank], r1 = ctx->bandRow1[ctx->bandRank];
    const cudaMemcpyKind kind = sourceIsDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CU(cudaMemcpyAsync(dst + (size_t)r0 * rowBytes, yBand, (size_t)(r1 - r0) * rowBytes, kind, ctx->stream));
    CU(cudaMemcpyAsync(dst + ylen + (size_t)(r0 >> 1) * rowBytes, uvBand, (size_t)((r1 >> 1) - (r0 >> 1)) * rowBytes, kind, ctx->stream));
    if (!sourceIsDevice) g_h2dBytes += (size_t)(r1 - r0) * rowBytes + (size_t)((r1 >> 1) - (r0 >> 1)) * rowBytes;
    if (ctx->bandWorld > 1) {
        band_signal_kernel<<<1, 1, 0, ctx->stream>>>(ctx->mail + 0, n + 1);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    ctx->fy[1] = dst;
    ctx->fuv[1] = dst + ylen;
    ctx->fslot[1] = slot;
    ctx->bandPending = 1;
    return 0;
}

/* Phase 2: fetch the halo — the rows of the neighbouring bands within reach of this band's search and warp — from the
 * owners' frame slots over NVLink (P2P copies, each after the owner's upload has been signalled), tell the peers, pack
 * the rows held here for the search. blocking != 0: waits for the stream like updateFrame does. */
extern "C" int hr_band_gather(HrContext *ctx, int blocking) {
    if (!ctx) return 1;
    if (!ctx->bandPending) return fail(ctx, "hr_band_gather: no upload outstanding");
    if (bind_device(ctx)) return 1;
    const int slot = ctx->fslot[1];
    const unsigned long long n = ctx->bandFrames;
    unsigned char *dst = ctx->frameBuf[slot];
    const size_t rowBytes = (size_t)ctx->W * ctx->bps, ylen = (size_t)ctx->H * rowBytes;
    /* rows, planes and slots aligned to 16 bytes (every real frame): one launch waits, copies and signals */
    const int fused = rowBytes % 16 == 0 && ((uintptr_t)dst % 16) == 0 && ylen % 16 == 0 && ctx->bandWorld > 1;
    if (fused) {
        HaloJob J;
        memset(&J, 0, sizeof(J));
        unsigned long long total = 0;
        for (int k = 1; k < ctx->bandWorld; ++k) {
            const int r = (ctx->bandRank + k) % ctx->bandWorld;
            if (!band_reads(ctx, ctx->bandRank, r)) continue;
            const unsigned char *src = ctx->peerSlot[r][slot];
            const int a = ctx->bandRow0[r] > ctx->haloLo ? ctx->bandRow0[r] : ctx->haloLo;
            const int b = ctx->bandRow1[r] < ctx->haloHi ? ctx->bandRow1[r] : ctx->haloHi;
            const size_t off[2] = {(size_t)a * rowBytes, ylen + (size_t)(a >> 1) * rowBytes};
            const size_t len[2] = {(size_t)(b - a) * rowBytes, (size_t)((b >> 1) - (a >> 1)) * rowBytes};
            for (int pl = 0; pl < 2; ++pl) {
                if (((uintptr_t)(src + off[pl])) % 16) return fail(ctx, "hr_band_gather: peer slot not aligned");
                J.src[J.nSeg] = (const uint4 *)(src + off[pl]);
                J.dst[J.nSeg] = (uint4 *)(dst + off[pl]);
                J.units[J.nSeg] = len[pl] / 16;
                J.uploaded[J.nSeg] = ctx->peerMail[r] + 0;
                total += len[pl];
                ++J.nSeg;
            }
        }
        J.frame = n + 1;
        J.gathered = ctx->mail + 1;
        J.exitCount = (unsigned int *)(ctx->arena + ctx->aExit) + 2;
        if (J.nSeg) {
            int ctas = (int)((total / 16 + 255) / 256);
            if (ctas > 2 * ctx->smCount) ctas = 2 * ctx->smCount;
            if (ctas < 1) ctas = 1;
            band_halo_kernel<<<ctas, 256, 0, ctx->stream>>>(J);
            ctx->launches++;
            ctx->bandP2pBytes += total;
        } else {
            band_signal_kernel<<<1, 1, 0, ctx->stream>>>(ctx->mail + 1, n + 1);
            ctx->launches++;
        }
    }
    for (int k = 1; k < ctx->bandWorld && !fused; ++k) {
        const int r = (ctx->bandRank + k) % ctx->bandWorld;
        if (!band_reads(ctx, ctx->bandRank, r)) continue;
        band_wait_kernel<<<1, 1, 0, ctx->stream>>>(ctx->peerMail[r] + 0, n + 1);
        ctx->launches++;
        const unsigned char *src = ctx->peerSlot[r][slot];
        const int a = ctx->bandRow0[r] > ctx->haloLo ? ctx->bandRow0[r] : ctx->haloLo;
        const int b = ctx->bandRow1[r] < ctx->haloHi ? ctx->bandRow1[r] : ctx->haloHi;
        CU(cudaMemcpyAsync(dst + (size_t)a * rowBytes, src + (size_t)a * rowBytes, (size_t)(b - a) * rowBytes, cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(dst + ylen + (size_t)(a >> 1) * rowBytes, src + ylen + (size_t)(a >> 1) * rowBytes, (size_t)((b >> 1) - (a >> 1)) * rowBytes,
                           cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->bandP2pBytes += (size_t)(b - a) * rowBytes + (size_t)((b >> 1) - (a >> 1)) * rowBytes;
    }
    if (ctx->bandWorld > 1 && !fused) {
        band_signal_kernel<<<1, 1, 0, ctx->stream>>>(ctx->mail + 1, n + 1);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    if (launch_pack(ctx)) return 1;
    ctx->framesSeen++;
    ctx->bandFrames = n + 1;
    ctx->bandPending = 0;
    if (blocking) CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int hr_band_get_halo(const HrContext *ctx, int *lo, int *hi, unsigned long long *p2pBytes) {
    if (!ctx || !ctx->banded) return 1;
    if (lo) *lo = ctx->haloLo;
    if (hi) *hi = ctx->haloHi;
    if (p2pBytes) *p2pBytes = ctx->bandP2pBytes;
    return 0;
}

/* downloadFrame for a band: only this context's rows, to host pointers of the band's first rows */
extern "C" int hr_band_download(HrContext *ctx, void *yBand, void *uvBand, double *seconds) {
    if (!ctx) return 1;
    if (!ctx->banded) return fail(ctx, "hr_band_download: bands are not configured");
    if (!yBand || !uvBand) return fail(ctx, "hr_band_download: NULL plane");
    if (bind_device(ctx)) return 1;
    const size_t rowBytes = (size_t)ctx->W * ctx->bps;
    const int r0 = ctx->bandRow0[ctx->bandRank], r1 = ctx->bandRow1[ctx->bandRank];
    CU(cudaMemcpyAsync(yBand, (const unsigned char *)ctx->outY + (size_t)r0 * rowBytes, (size_t)(r1 - r0) * rowBytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(uvBand, (const unsigned char *)ctx->outUV + (size_t)(r0 >> 1) * rowBytes, (size_t)((r1 >> 1) - (r0 >> 1)) * rowBytes, cudaMemcpyDeviceToHost,
                       ctx->stream));
    g_d2hBytes += (size_t)(r1 - r0) * rowBytes + (size_t)((r1 >> 1) - (r0 >> 1)) * rowBytes;
    CU(cudaEventRecord(ctx->evDlEnd, ctx->stream));
    CU(cudaEventSynchronize(ctx->evDlEnd));
    if (seconds) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->evWarpStart, ctx->evDlEnd) != cudaSuccess) {
            cudaGetLastError();
            ms = 0.f;
        }
        *seconds = (double)ms * 1e-3;
    }
    return 0;
}

extern "C" int hr_debug_host_transfer_bytes(unsigned long long *h2d, unsigned long long *d2h) {
    if (h2d) *h2d = g_h2dBytes;
    if (d2h) *d2h = g_d2hBytes;
    return 0;
}

/* Page-locked host memory for frames the filter allocates itself (its output image pool, patches/0004) */
extern "C" int hr_host_alloc(void **out, size_t bytes) {
    if (!out || bytes == 0) return 1;
    *out = NULL;
    if (cudaHostAlloc(out, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        *out = NULL;
        return 1;
    }
    return 0;
}
extern "C" int hr_host_free(void *p) {
    if (!p) return 0;
    if (cudaFreeHost(p) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    return 0;
}
extern "C" int hr_debug_host_pointer_kind(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return a.type == cudaMemoryTypeUnregistered ? 0 : a.type == cudaMemoryTypeHost ? 1 : 2;
}

/* Everything enqueued so far is complete when this returns (outputs in caller-owned device planes included);
 * seconds: device time from the start of the most recent hr_warp to now (= warpCalcTime of a frame that is not
 * downloaded, opticalFlowCalc.c:117-122). */
extern "C" int hr_finish(HrContext *ctx, double *seconds) {
    if (!ctx) return 1;
    if (bind_device(ctx)) return 1;
    if (pipe_join(ctx)) return 1;
    CU(cudaEventRecord(ctx->evDlEnd, ctx->stream));
    CU(cudaEventSynchronize(ctx->evDlEnd));
    if (seconds) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->evWarpStart, ctx->evDlEnd) != cudaSuccess) {
            cudaGetLastError(); /* no warp was recorded yet */
            ms = 0.f;
        }
        *seconds = (double)ms * 1e-3;
    }
    return 0;
}

extern "C" int hr_get_output_device(HrContext *ctx, void **dY, void **dUV) {
    if (!ctx) return 1;
    if (dY) *dY = ctx->outY;
    if (dUV) *dUV = ctx->outUV;
    return 0;
}

extern "C" int hr_set_output_device(HrContext *ctx, void *dY, void *dUV) {
    if (!ctx) return 1;
    if ((dY == NULL) != (dUV == NULL)) return fail(ctx, "hr_set_output_device: give both planes or neither");
    ctx->specWarp.valid = 0;
    if (dY) {
        ctx->outY = dY;
        ctx->outUV = dUV;
    } else {
        ctx->outY = ctx->outBuf;
        ctx->outUV = ctx->outBuf + (size_t)ctx->H * ctx->W * ctx->bps;
    }
    return 0;
}

extern "C" int hr_get_offsets(HrContext *ctx, int16_t *raw, int16_t *blurred) {
    if (!ctx) return 1;
    if (bind_device(ctx)) return 1;
    const size_t n = 2 * (size_t)ctx->lw * ctx->lh * sizeof(int16_t);
    if (sync_all(ctx)) return 1;
    if (raw) CU(cudaMemcpy(raw, ctx->off, n, cudaMemcpyDeviceToHost));
    if (blurred) CU(cudaMemcpy(blurred, ctx->blur, n, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int hr_set_blurred_offsets(HrContext *ctx, const int16_t *blurred) {
    if (!ctx || !blurred) return 1;
    if (bind_device(ctx)) return 1;
    const size_t n = 2 * (size_t)ctx->lw * ctx->lh * sizeof(int16_t);
    ctx->specWarp.valid = 0;
    if (sync_all(ctx)) return 1;
    CU(cudaMemcpy(ctx->blur, blurred, n, cudaMemcpyHostToDevice));
    ctx->flowSerial[ctx->flowCur] = ++ctx->flowStamp;
    const int ln = ctx->lw * ctx->lh;
    pack_flow_kernel<<<(ln + 255) / 256, 256, 0, ctx->stream>>>(ctx->blur, ctx->blurXY, ln);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int hr_blur_flow(HrContext *ctx, const int16_t *rawHost, int16_t *blurredHost) {
    if (!ctx || !rawHost || !blurredHost) return 1;
    if (bind_device(ctx)) return 1;
    const size_t n = 2 * (size_t)ctx->lw * ctx->lh * sizeof(int16_t);
    int16_t *dIn = NULL, *dOut = NULL;
    CU(cudaMalloc(&dIn, n));
    CU(cudaMalloc(&dOut, n));
    CU(cudaMemcpyAsync(dIn, rawHost, n, cudaMemcpyHostToDevice, ctx->stream));
    dim3 block(32, 8), grid((ctx->lw + 31) / 32, (ctx->lh + 7) / 8, 2);
    blur_flow_kernel<<<grid, block, 0, ctx->stream>>>(dIn, dOut, ctx->lh, ctx->lw);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(blurredHost, dOut, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dIn);
    cudaFree(dOut);
    if (e != cudaSuccess) return fail(ctx, "CUDA error in hr_blur_flow: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int hr_get_step_layers(HrContext *ctx, int step, uint8_t *layers) {
    if (!ctx || !layers) return 1;
    if (!ctx->trace) return fail(ctx, "hr_get_step_layers: tracing was not enabled");
    if (step < 0 || step >= 2 * ctx->iters) return fail(ctx, "hr_get_step_layers: step %d out of range", step);
    if (bind_device(ctx)) return 1;
    const size_t ln = (size_t)ctx->lw * ctx->lh;
    if (sync_all(ctx)) return 1;
    CU(cudaMemcpy(layers, ctx->trace + (size_t)step * ln, ln, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int hr_get_kernel_times(HrContext *ctx, double *searchSeconds, double *warpSeconds, double *packSeconds) {
    if (!ctx) return 1;
    if (bind_device(ctx)) return 1;
    if (sync_all(ctx)) return 1;
    double *outs[3] = {searchSeconds, warpSeconds, packSeconds};
    const int have[3] = {ctx->haveSearchT, ctx->haveWarpT, ctx->havePackT};
    for (int i = 0; i < 3; ++i) {
        if (!outs[i]) continue;
        *outs[i] = 0.0;
        if (!have[i]) continue;
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ctx->evK[2 * i], ctx->evK[2 * i + 1]));
        *outs[i] = (double)ms * 1e-3;
    }
    return 0;
}

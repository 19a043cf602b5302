/*
 * hr_common.cuh — shared declarations of the sm_100a device code of the HopperRender hot path
 * (hr_pack.cuh, hr_search*.cuh, hr_warp.cuh, hr_warp_fast.cuh).
 *
 * Written from scratch for B200; it is not a translation of the reference's OpenCL kernels
 * (video/filter/HopperRender/Kernels/*.cl) but computes the same results:
 *
 *   pack_frame16_kernel    per source frame: NV12/P010 -> phase-planar packed (Y,U,V,0) words, so
 *                          that every delta-sum evaluation is ONE coalesced 32-bit load and ONE
 *                          VABSDIFF4.U8.ACC (replaces the three strided byte gathers of
 *                          calcDeltaSumsKernel.cl:96-98).
 *   flow_search*_kernel    one persistent cooperative launch for all 2*iterations search steps
 *                          (K1 calcDeltaSumsKernel.cl:34-189 + K2 determineLowestLayerKernel.cl:2-22
 *                          + K3 adjustOffsetArrayKernel.cl:2-18) and the 8x8 flow blur
 *                          (K4 blurFlowKernel.cl:15-89); offsets are kept at window granularity.
 *                          Three generations with the same tables and results (hr_search.cuh,
 *                          hr_search2.cuh, hr_search3.cuh), chosen per launch by hr_cuda.cu.
 *   warp_fast_kernel /     K5 warpFrameKernel.cl:114-182: flip lookup, bidirectional warp, blend,
 *   warp_generic_kernel    levels, output modes; luma and chroma (and the 2-6 outputs of a frame
 *                          pair) in one launch, 32 / 64 / 128-bit accesses by resolution scalar.
 *
 * Compiled with -fmad=false: no IMPLICIT contraction. The warp's float arithmetic is the one the
 * NVIDIA OpenCL compiler emits for the unmodified reference kernel (hr_warp.cuh): its contractions
 * and MUFU.RCP divisions are written out explicitly, which is what makes 8-bit output bit-identical
 * to the reference run on the same GPU (tests/test_gpu_vs_reference_opencl.py).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HR_TILE 32                 /* lattice points per tile side (one CTA)                     */
#define HR_MAX_TILES_PER_CTA 4
#define HR_MAX_LEVELS 16
#define HR_ZCHUNK 8                /* candidate layers in flight per thread                       */
#define HR_RMAX 32                 /* HR_MAX_SEARCH_RADIUS                                       */
#define HR_FIRST_NEIGHBOR_ITERATION 4 /* calcDeltaSumsKernel.cl:1 */
#define HR_TIMELINE_SLOTS 128

/* Spatial bands: the lattice tiles are split by tile rows over `world` GPUs, every GPU searches its own
 * tiles only. What tiles hand to one another across GPUs — tile totals of the windows that span tiles, window-table
 * entries at a band's edge, the flow itself — is stored straight into the peers' copies over NVLink: all of it lives in
 * one "exchange arena" per GPU with the same layout everywhere, so the address of a word in peer g's arena is the local
 * address plus peerDelta[g] (0 for the GPU itself). Consumers only ever poll their own memory. */
#define HR_MAX_BANDS 16
struct BandLink {
    int world, rank;                       /* world <= 1: no bands                                                  */
    int tile0;                             /* first lattice tile of this GPU (tiles tile0 .. tile0 + gridDim.x - 1)  */
    int tileRow0, tileRow1;                /* its tile rows                                                          */
    int up, down;                          /* GPUs that own the tile row above / below the band (-1: none)           */
    long long peerDelta[HR_MAX_BANDS];     /* byte distance from a local arena address to the same word at GPU g     */
    unsigned long long *ready, *done;      /* local [HR_MAX_BANDS]: epoch GPU g has entered / finished writing        */
    unsigned int *exitCount;               /* local: CTAs of this launch that have stored all their results           */
};

struct FlowParams {
    const uint32_t *p1;      /* packed previous frame (frame1): all phase planes                 */
    const void *f2y, *f2uv;  /* newest frame (frame2) as it arrived, NV12 / P010: read at the lattice points only, so
                              * that its packed copy (needed as frame1 of the NEXT pair) can be built concurrently */
    int bps;                 /* bytes per sample of f2y / f2uv (1 or 2; the search uses the top 8 bits)  */
    int planePitch;          /* words per packed plane row                                       */
    int planeSize;           /* words per packed plane                                           */
    int W, H, s, lw, lh;
    int first, iters, R, dS, nS;
    int cand[HR_RMAX];       /* signed-square layer shifts, calcDeltaSumsKernel.cl:68-72         */
    int tilesX, tilesY, numTiles;
    unsigned long long *T;   /* per-level window offset tables: epoch << 32 | (x | y << 16)       */
    int tOff[HR_MAX_LEVELS]; /* word offset of level `it` in T                                    */
    unsigned long long *partial; /* cross-tile window sums [bigStep][tile][HR_RMAX]: epoch << 32 | tile total */
    int bigOff[2 * HR_MAX_LEVELS]; /* word offset of search step k in partial (-1: not a big step) */
    uint32_t epoch;          /* tag of this launch (never 0, differs from the previous launch)     */
    int16_t *off;            /* raw offsets  [2][lh][lw]  (offsetArray)                           */
    int16_t *blur;           /* blurred      [2][lh][lw]  (blurredOffsetArray)                    */
    uint32_t *blurXY;        /* the same, one word per lattice point: x | y << 16 (read by the warp)  */
    uint8_t *trace;          /* optional [steps][lh][lw] winning layer per point, or NULL         */
    long long *timeline;     /* optional [ctas][HR_TIMELINE_SLOTS] clock64 stamps of thread 0, or NULL */
    BandLink band;
};

template <typename T>
struct WarpParams {
    const T *f1y, *f1uv;     /* sourceFrame12 = previous frame                                    */
    const T *f2y, *f2uv;     /* sourceFrame21 = newest frame                                      */
    T *outY, *outUV;
    const int16_t *flow;     /* blurred offsets [2][lh][lw]                                       */
    const uint32_t *flowXY;  /* the same as one word per lattice point: x | y << 16                  */
    int lw, lh, H, W, aW, s, mode;
    float t12, t21, black, white;
};

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                                */
/* ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ int hr_min(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int hr_max(int a, int b) { return a > b ? a : b; }

/* calcDeltaSumsKernel.cl:84-93 single reflection, then clamp (the reference leaves |offset| >= dim
 * undefined; this implementation clamps). */
__device__ __forceinline__ int search_mirror(int p, int D) {
    if (p >= D) p = 2 * D - p - 1;
    else if (p < 0) p = -p - 1;
    return hr_min(hr_max(p, 0), D - 1);
}
/* calcDeltaSumsKernel.cl:68-72: signed square of the relative layer */
__device__ __forceinline__ int candidate(int z, int R) {
    const int rel = z - (R >> 1);
    return rel * (rel < 0 ? -rel : rel);
}
__device__ __forceinline__ uint32_t ldcg_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
/* the same word at another GPU (peer-mapped memory, NVLink) */
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
template <typename T>
__device__ __forceinline__ T *at_peer(T *local, long long delta) {
    return reinterpret_cast<T *>(reinterpret_cast<char *>(local) + delta);
}

/* ---- tile-to-tile hand-off without barriers ------------------------------------------------------
 * Every word that one CTA produces for another (window-table entries, per-tile window sums) is a
 * 64-bit word written exactly once per launch: payload in the low half, the launch's epoch tag in the
 * high half, stored with one 64-bit store. A consumer loads the word it needs and re-loads while the
 * tag is stale: the payload and its "ready" flag arrive together, so no fence, no flag array, no
 * atomics and no grid barrier are needed, and a CTA only ever waits for the words it actually reads.
 * All CTAs are co-resident (cooperative launch) and producers never wait on consumers of the same
 * level, so polling cannot deadlock. */
__device__ __forceinline__ void put_tagged(unsigned long long *p, uint32_t epoch, uint32_t payload) {
    st_relaxed_u64(p, ((unsigned long long)epoch << 32) | payload);
}
/* ... and into the arena of every GPU of a band group (the own one included: delta 0) */
__device__ __forceinline__ void put_tagged_all(const BandLink &B, unsigned long long *p, uint32_t epoch, uint32_t payload) {
    const unsigned long long v = ((unsigned long long)epoch << 32) | payload;
    for (int g = 0; g < B.world; ++g) st_relaxed_sys_u64(at_peer(p, B.peerDelta[g]), v);
}
__device__ __forceinline__ uint32_t get_tagged(const unsigned long long *p, uint32_t epoch) {
    unsigned long long v = ld_relaxed_u64(p);
    while ((uint32_t)(v >> 32) != epoch) v = ld_relaxed_u64(p);
    return (uint32_t)v;
}

/*
 * hr_kernels.cuh — sm_100a device code of the HopperRender hot path.
 *
 * Written from scratch for B200; it is not a translation of the reference's OpenCL kernels
 * (video/filter/HopperRender/Kernels/*.cl) but computes the same results:
 *
 *   pack_frame_kernel      per source frame: NV12/P010 -> phase-planar packed (Y,U,V,0) words, so
 *                          that every delta-sum evaluation is ONE coalesced 32-bit load and ONE
 *                          VABSDIFF4.U8.ACC (replaces the three strided byte gathers of
 *                          calcDeltaSumsKernel.cl:96-98).
 *   flow_search_kernel     one persistent cooperative launch for all 2*iterations search steps
 *                          (K1 calcDeltaSumsKernel.cl:34-189 + K2 determineLowestLayerKernel.cl:2-22
 *                          + K3 adjustOffsetArrayKernel.cl:2-18) and the 8x8 flow blur
 *                          (K4 blurFlowKernel.cl:15-89); offsets are kept at window granularity.
 *   warp_blend_kernel      K5 warpFrameKernel.cl:114-182: flip lookup, bidirectional warp, blend,
 *                          levels, output modes; luma and chroma in one launch, 32-bit stores.
 *
 * Compiled with -fmad=false: the warp's float arithmetic must round after every operation (the
 * reference's expressions evaluated in IEEE single precision without contraction), which is
 * what the parity tests check bit for bit.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HR_TILE 32                 /* lattice points per tile side; CTA = 32 x 32 threads        */
#define HR_MAX_TILES_PER_CTA 8
#define HR_MAX_LEVELS 16
#define HR_ZCHUNK 8                /* candidate layers evaluated per register chunk              */
#define HR_RMAX 32                 /* HR_MAX_SEARCH_RADIUS                                       */
#define HR_FIRST_NEIGHBOR_ITERATION 4 /* calcDeltaSumsKernel.cl:1 */

struct FlowParams {
    const uint32_t *p1;      /* packed previous frame (frame1): all phase planes                 */
    const uint32_t *p2;      /* packed newest frame (frame2): only phase plane (0,0) is read      */
    int planePitch;          /* words per packed plane row                                       */
    int planeSize;           /* words per packed plane                                           */
    int W, H, s, lw, lh;
    int first, iters, R, dS, nS;
    int tilesX, numTiles;
    uint32_t *T;             /* per-level window offset tables, int16x2 (x | y << 16)             */
    int tOff[HR_MAX_LEVELS]; /* word offset of level `it` in T                                    */
    uint32_t *bigSums;       /* cross-CTA window sums for windows > tile: [bigStep][win][HR_RMAX] */
    int bigOff[2 * HR_MAX_LEVELS]; /* word offset of search step k in bigSums (-1: not a big step) */
    int bigWords;
    unsigned long long *bar; /* monotonic grid-barrier counter                                   */
    unsigned long long barBase;
    int16_t *off;            /* raw offsets  [2][lh][lw]  (offsetArray)                           */
    int16_t *blur;           /* blurred      [2][lh][lw]  (blurredOffsetArray)                    */
    uint8_t *trace;          /* optional [steps][lh][lw] winning layer per point, or NULL         */
};

template <typename T>
struct WarpParams {
    const T *f1y, *f1uv;     /* sourceFrame12 = previous frame                                    */
    const T *f2y, *f2uv;     /* sourceFrame21 = newest frame                                      */
    T *outY, *outUV;
    const int16_t *flow;     /* blurred offsets [2][lh][lw]                                       */
    const uint8_t *lut;      /* [2][256] 8-bit levels LUT (Y then UV)                             */
    int lw, lh, H, W, aW, s, mode, lutIdentity;
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
__device__ __forceinline__ void red_release_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

/* Grid-wide barrier of the persistent search kernel (all CTAs are co-resident: cooperative
 * launch). The counter only grows; `target` is carried by every CTA. */
__device__ __forceinline__ void grid_barrier(unsigned long long *bar, unsigned long long &target, unsigned nCtas) {
    __syncthreads();
    target += nCtas;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        __threadfence();
        red_release_add_u64(bar, 1ULL);
        while (ld_acquire_u64(bar) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

/* ------------------------------------------------------------------------------------------ */
/* pack: frame -> phase-planar packed words                                                      */
/*   word(px,py)[ly][lx] = Y(x,y) | U(x,y) << 8 | V(x,y) << 16, x = lx<<s | px, y = ly<<s | py,     */
/*   with U,V taken at chroma row y>>1, byte column x&~1 (+1): calcDeltaSumsKernel.cl:96-98.       */
/*   P010: the top 8 bits of every sample.                                       */
/* ------------------------------------------------------------------------------------------ */
template <typename T>
__device__ __forceinline__ uint32_t top8(T v);
template <>
__device__ __forceinline__ uint32_t top8<uint8_t>(uint8_t v) { return v; }
template <>
__device__ __forceinline__ uint32_t top8<uint16_t>(uint16_t v) { return (uint32_t)v >> 8; }

template <typename T>
__global__ void pack_frame_kernel(const T *__restrict__ yPlane, const T *__restrict__ uvPlane, uint32_t *__restrict__ packed,
                                  int W, int H, int s, int lw, int planePitch, int planeSize) {
    const int row = blockIdx.y;
    const int lx = blockIdx.x * blockDim.x + threadIdx.x;
    const int px = threadIdx.y;
    const int x = (lx << s) | px;
    if (lx >= lw || x >= W) return;
    const uint32_t yv = top8<T>(__ldg(yPlane + (size_t)row * W + x));
    const T *uvp = uvPlane + (size_t)(row >> 1) * W + (x & ~1);
    const uint32_t uv = top8<T>(__ldg(uvp));
    const uint32_t vv = top8<T>(__ldg(uvp + 1));
    const int m = (1 << s) - 1;
    const int plane = ((row & m) << s) | px;
    packed[(size_t)plane * planeSize + (size_t)(row >> s) * planePitch + lx] = yv | (uv << 8) | (vv << 16);
}

/* ------------------------------------------------------------------------------------------ */
/* search                                                                                        */
/* ------------------------------------------------------------------------------------------ */
struct SearchShared {
    uint32_t warpSums[HR_TILE][HR_RMAX];         /* [warp][z]   windows >= tile                     */
    union {
        uint32_t segSums[HR_ZCHUNK][HR_TILE][16]; /* [z in chunk][row][segment] windows < tile       */
        struct {                                  /* blur phase                                     */
            int16_t tX[40 * 40], tY[40 * 40];     /* tile + 4-point halo of the raw offsets          */
            int hX[40 * 32], hY[40 * 32];         /* horizontal 8-tap sums                           */
        } blur;
    };
    int16_t curX[256], curY[256];                /* current offsets of the tile's windows            */
    uint16_t cnt[256];                           /* in-lattice points of each window                 */
    uint8_t win[256];                            /* winning layer of each window this step           */
    int tileOffX[HR_MAX_TILES_PER_CTA], tileOffY[HR_MAX_TILES_PER_CTA];
    int uniWinner;
};

/* Window total for layer z (calcDeltaSumsKernel.cl:99-150 summed over the window, mod 2^32):
 * every point of a window shares offset and neighbours, so both biases enter as count * bias. */
__device__ __forceinline__ uint32_t window_total(uint32_t sad, int z, int R, int curAxis, uint32_t count, bool useNb,
                                                 int nb0, int nb1, int nb2, int nb3, int dS, int nS) {
    const int own = (int)(int16_t)(curAxis + candidate(z, R));
    uint32_t bias = (uint32_t)(uint16_t)(own < 0 ? -own : own);
    if (useNb) {
        uint32_t nb = 0;
        int d;
        d = nb0 - own; nb += (uint32_t)(uint16_t)(d < 0 ? -d : d);
        d = nb1 - own; nb += (uint32_t)(uint16_t)(d < 0 ? -d : d);
        d = nb2 - own; nb += (uint32_t)(uint16_t)(d < 0 ? -d : d);
        d = nb3 - own; nb += (uint32_t)(uint16_t)(d < 0 ? -d : d);
        bias += nb << nS;
    }
    return (sad << dS) + count * bias;
}

/* The four neighbour offsets of calcDeltaSumsKernel.cl:112-128 for the window whose lattice
 * origin is (x0,y0): positions +-2*ws clamped to the lattice, read from the previous level. */
__device__ __forceinline__ void load_neighbours(const FlowParams &P, int it, int ws, int axis, int x0, int y0, int &nb0,
                                                int &nb1, int &nb2, int &nb3) {
    const int pws = ws << 1;
    const int pnwx = (P.lw + pws - 1) / pws;
    const uint32_t *Tp = P.T + P.tOff[it - 1];
    const int yd = hr_min(y0 + pws, P.lh - 1) / pws, yu = hr_max(y0 - pws, 0) / pws;
    const int xr = hr_min(x0 + pws, P.lw - 1) / pws, xl = hr_max(x0 - pws, 0) / pws;
    const int xc = x0 / pws, yc = y0 / pws;
    const uint32_t a = ldcg_u32(Tp + yd * pnwx + xc); /* down  */
    const uint32_t b = ldcg_u32(Tp + yc * pnwx + xr); /* right */
    const uint32_t c = ldcg_u32(Tp + yc * pnwx + xl); /* left  */
    const uint32_t d = ldcg_u32(Tp + yu * pnwx + xc); /* up    */
    const int sh = axis ? 16 : 0;
    nb0 = (int)(int16_t)(a >> sh);
    nb1 = (int)(int16_t)(b >> sh);
    nb2 = (int)(int16_t)(c >> sh);
    nb3 = (int)(int16_t)(d >> sh);
}

/* SADs of HR_ZCHUNK layers for one lattice point. axis 0: layers move x, axis 1: layers move y. */
__device__ __forceinline__ void eval_chunk(const FlowParams &P, bool inLat, int cx, int cy, int ox, int oy, int axis, int z0,
                                           uint32_t v2, uint32_t (&sad)[HR_ZCHUNK]) {
    const int s = P.s, m = (1 << s) - 1;
#pragma unroll
    for (int j = 0; j < HR_ZCHUNK; ++j) sad[j] = 0;
    if (!inLat) return;
    if (axis == 0) {
        const int y = search_mirror((cy << s) + oy, P.H);
        const uint32_t *rowp = P.p1 + (size_t)((y & m) << s) * P.planeSize + (size_t)(y >> s) * P.planePitch;
#pragma unroll
        for (int j = 0; j < HR_ZCHUNK; ++j) {
            const int z = z0 + j;
            if (z < P.R) {
                const int x = search_mirror((cx << s) + (int)(int16_t)(ox + candidate(z, P.R)), P.W);
                const uint32_t v1 = __ldg(rowp + (size_t)(x & m) * P.planeSize + (x >> s));
                sad[j] = __vsadu4(v1, v2);
            }
        }
    } else {
        const int x = search_mirror((cx << s) + ox, P.W);
        const uint32_t *colp = P.p1 + (size_t)(x & m) * P.planeSize + (x >> s);
#pragma unroll
        for (int j = 0; j < HR_ZCHUNK; ++j) {
            const int z = z0 + j;
            if (z < P.R) {
                const int y = search_mirror((cy << s) + (int)(int16_t)(oy + candidate(z, P.R)), P.H);
                const uint32_t v1 = __ldg(colp + (size_t)((y & m) << s) * P.planeSize + (size_t)(y >> s) * P.planePitch);
                sad[j] = __vsadu4(v1, v2);
            }
        }
    }
}

__global__ void __launch_bounds__(HR_TILE *HR_TILE, 1) flow_search_kernel(const FlowParams P) {
    __shared__ SearchShared sh;
    const int lane = threadIdx.x, warp = threadIdx.y, tid = warp * HR_TILE + lane;
    const unsigned nCtas = gridDim.x;
    unsigned long long barTarget = P.barBase;
    const int R = P.R;

    if (tid < HR_MAX_TILES_PER_CTA) {
        sh.tileOffX[tid] = 0;
        sh.tileOffY[tid] = 0;
    }
    __syncthreads();

    for (int it = 0; it < P.iters; ++it) {
        const int ws = P.first >> it;
        const int nwx = (P.lw + ws - 1) / ws; /* windows per row at this level */
        uint32_t *Tcur = P.T + P.tOff[it];
        const bool useNb = it >= HR_FIRST_NEIGHBOR_ITERATION;

        if (ws >= HR_TILE) {
            /* ---------------- windows cover one tile or more ---------------------------------- */
            const bool big = ws > HR_TILE;
            for (int axis = 0; axis < 2; ++axis) {
                const int step = it * 2 + axis;
                /* phase A: tile SADs; for big windows accumulate across CTAs */
                int slot = 0;
                for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas, ++slot) {
                    const int tx0 = (tile % P.tilesX) * HR_TILE, ty0 = (tile / P.tilesX) * HR_TILE;
                    const int cx = tx0 + lane, cy = ty0 + warp;
                    const bool inLat = cx < P.lw && cy < P.lh;
                    const int ox = sh.tileOffX[slot], oy = sh.tileOffY[slot];
                    const uint32_t v2 = inLat ? __ldg(P.p2 + (size_t)cy * P.planePitch + cx) : 0u;
                    for (int z0 = 0; z0 < R; z0 += HR_ZCHUNK) {
                        uint32_t sad[HR_ZCHUNK];
                        eval_chunk(P, inLat, cx, cy, ox, oy, axis, z0, v2, sad);
#pragma unroll
                        for (int j = 0; j < HR_ZCHUNK; ++j) {
                            if (z0 + j < R) { /* uniform */
                                const uint32_t r = __reduce_add_sync(0xffffffffu, sad[j]);
                                if (lane == 0) sh.warpSums[warp][z0 + j] = r;
                            }
                        }
                    }
                    __syncthreads();
                    if (warp == 0) {
                        uint32_t tot = 0;
                        if (lane < R) {
#pragma unroll 8
                            for (int w = 0; w < HR_TILE; ++w) tot += sh.warpSums[w][lane];
                        }
                        const int wx = tx0 / ws, wy = ty0 / ws;
                        if (big) {
                            if (lane < R) atomicAdd(P.bigSums + P.bigOff[step] + (size_t)(wy * nwx + wx) * HR_RMAX + lane, tot);
                        } else {
                            sh.warpSums[0][lane] = (lane < R) ? tot : 0u; /* keep for phase B (one tile per pass) */
                        }
                    }
                    __syncthreads();
                    if (!big) {
                        /* phase B inline for tile-sized windows */
                        if (warp == 0) {
                            const int wx = tx0 / ws, wy = ty0 / ws;
                            const int x0 = wx * ws, y0 = wy * ws;
                            const uint32_t count = (uint32_t)(hr_min(x0 + ws, P.lw) - x0) * (uint32_t)(hr_min(y0 + ws, P.lh) - y0);
                            int nb0 = 0, nb1 = 0, nb2 = 0, nb3 = 0;
                            if (useNb) load_neighbours(P, it, ws, axis, x0, y0, nb0, nb1, nb2, nb3);
                            const int cur = axis ? sh.tileOffY[slot] : sh.tileOffX[slot];
                            const uint32_t S = (lane < R) ? window_total(sh.warpSums[0][lane], lane, R, cur, count, useNb, nb0, nb1, nb2, nb3, P.dS, P.nS) : 0xffffffffu;
                            const uint32_t mn = __reduce_min_sync(0xffffffffu, S);
                            const unsigned ballot = __ballot_sync(0xffffffffu, S == mn && lane < R);
                            const int winner = __ffs(ballot) - 1;
                            if (lane == 0) {
                                if (axis) sh.tileOffY[slot] = (int)(int16_t)(cur + candidate(winner, R));
                                else sh.tileOffX[slot] = (int)(int16_t)(cur + candidate(winner, R));
                                sh.uniWinner = winner;
                            }
                        }
                        __syncthreads();
                        if (P.trace && inLat) P.trace[((size_t)step * P.lh + cy) * P.lw + cx] = (uint8_t)sh.uniWinner;
                        if (axis == 1 && tid == 0)
                            Tcur[(ty0 / ws) * nwx + (tx0 / ws)] = (uint32_t)(uint16_t)sh.tileOffX[slot] | ((uint32_t)(uint16_t)sh.tileOffY[slot] << 16);
                        __syncthreads();
                    }
                }
                if (big) {
                    grid_barrier(P.bar, barTarget, nCtas);
                    /* phase B: every CTA derives the winner of the window(s) its tiles belong to */
                    slot = 0;
                    for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas, ++slot) {
                        const int tx0 = (tile % P.tilesX) * HR_TILE, ty0 = (tile / P.tilesX) * HR_TILE;
                        const int cx = tx0 + lane, cy = ty0 + warp;
                        const bool inLat = cx < P.lw && cy < P.lh;
                        if (warp == 0) {
                            const int wx = tx0 / ws, wy = ty0 / ws;
                            const int x0 = wx * ws, y0 = wy * ws;
                            const uint32_t count = (uint32_t)(hr_min(x0 + ws, P.lw) - x0) * (uint32_t)(hr_min(y0 + ws, P.lh) - y0);
                            int nb0 = 0, nb1 = 0, nb2 = 0, nb3 = 0;
                            if (useNb) load_neighbours(P, it, ws, axis, x0, y0, nb0, nb1, nb2, nb3);
                            const int cur = axis ? sh.tileOffY[slot] : sh.tileOffX[slot];
                            const uint32_t sad = (lane < R) ? ldcg_u32(P.bigSums + P.bigOff[step] + (size_t)(wy * nwx + wx) * HR_RMAX + lane) : 0u;
                            const uint32_t S = (lane < R) ? window_total(sad, lane, R, cur, count, useNb, nb0, nb1, nb2, nb3, P.dS, P.nS) : 0xffffffffu;
                            const uint32_t mn = __reduce_min_sync(0xffffffffu, S);
                            const unsigned ballot = __ballot_sync(0xffffffffu, S == mn && lane < R);
                            const int winner = __ffs(ballot) - 1;
                            if (lane == 0) {
                                if (axis) sh.tileOffY[slot] = (int)(int16_t)(cur + candidate(winner, R));
                                else sh.tileOffX[slot] = (int)(int16_t)(cur + candidate(winner, R));
                                sh.uniWinner = winner;
                            }
                        }
                        __syncthreads();
                        if (P.trace && inLat) P.trace[((size_t)step * P.lh + cy) * P.lw + cx] = (uint8_t)sh.uniWinner;
                        if (axis == 1 && tid == 0)
                            Tcur[(ty0 / ws) * nwx + (tx0 / ws)] = (uint32_t)(uint16_t)sh.tileOffX[slot] | ((uint32_t)(uint16_t)sh.tileOffY[slot] << 16);
                        __syncthreads();
                    }
                }
            }
        } else {
            /* ---------------- several windows per tile: both axis steps back to back ---------- */
            const int wpt = HR_TILE / ws;          /* windows per tile side */
            const int nWin = wpt * wpt;
            const int lg = 31 - __clz(ws);
            for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas) {
                const int tx0 = (tile % P.tilesX) * HR_TILE, ty0 = (tile / P.tilesX) * HR_TILE;
                const int cx = tx0 + lane, cy = ty0 + warp;
                const bool inLat = cx < P.lw && cy < P.lh;
                const int myWin = (warp >> lg) * wpt + (lane >> lg);
                const uint32_t v2 = inLat ? __ldg(P.p2 + (size_t)cy * P.planePitch + cx) : 0u;
                /* window set-up: parent offsets, in-lattice count */
                if (tid < nWin) {
                    const int x0 = tx0 + (tid % wpt) * ws, y0 = ty0 + (tid / wpt) * ws;
                    int c = 0;
                    int16_t px = 0, py = 0;
                    if (x0 < P.lw && y0 < P.lh) {
                        c = (hr_min(x0 + ws, P.lw) - x0) * (hr_min(y0 + ws, P.lh) - y0);
                        if (it > 0) {
                            const int pws = ws << 1;
                            const int pnwx = (P.lw + pws - 1) / pws;
                            const uint32_t pv = ldcg_u32(P.T + P.tOff[it - 1] + (y0 / pws) * pnwx + (x0 / pws));
                            px = (int16_t)(pv & 0xffffu);
                            py = (int16_t)(pv >> 16);
                        }
                    }
                    sh.cnt[tid] = (uint16_t)c;
                    sh.curX[tid] = px;
                    sh.curY[tid] = py;
                }
                __syncthreads();
                for (int axis = 0; axis < 2; ++axis) {
                    const int step = it * 2 + axis;
                    const int ox = sh.curX[myWin], oy = sh.curY[myWin];
                    /* per-window running minimum, owned by thread tid < nWin */
                    uint32_t bestS = 0xffffffffu;
                    int bestZ = 0;
                    int nb0 = 0, nb1 = 0, nb2 = 0, nb3 = 0, wcur = 0;
                    uint32_t wcount = 0;
                    if (tid < nWin) {
                        wcount = sh.cnt[tid];
                        wcur = axis ? sh.curY[tid] : sh.curX[tid];
                        if (useNb && wcount) load_neighbours(P, it, ws, axis, tx0 + (tid % wpt) * ws, ty0 + (tid / wpt) * ws, nb0, nb1, nb2, nb3);
                    }
                    for (int z0 = 0; z0 < R; z0 += HR_ZCHUNK) {
                        uint32_t sad[HR_ZCHUNK];
                        eval_chunk(P, inLat, cx, cy, ox, oy, axis, z0, v2, sad);
#pragma unroll
                        for (int j = 0; j < HR_ZCHUNK; ++j) {
                            if (z0 + j < R) {
                                uint32_t v = sad[j];
                                for (int o = ws >> 1; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                                if ((lane & (ws - 1)) == 0) sh.segSums[j][warp][lane >> lg] = v;
                            }
                        }
                        __syncthreads();
                        if (tid < nWin && wcount) {
                            const int wlx = tid % wpt, wly = tid / wpt;
                            for (int j = 0; j < HR_ZCHUNK && z0 + j < R; ++j) {
                                uint32_t sadw = 0;
                                for (int r = 0; r < ws; ++r) sadw += sh.segSums[j][wly * ws + r][wlx];
                                const uint32_t S = window_total(sadw, z0 + j, R, wcur, wcount, useNb, nb0, nb1, nb2, nb3, P.dS, P.nS);
                                if (z0 + j == 0 || S < bestS) { /* first minimum: determineLowestLayerKernel.cl:13-18 */
                                    bestS = S;
                                    bestZ = z0 + j;
                                }
                            }
                        }
                        __syncthreads();
                    }
                    if (tid < nWin && wcount) {
                        sh.win[tid] = (uint8_t)bestZ;
                        const int16_t nv = (int16_t)(wcur + candidate(bestZ, R));
                        if (axis) sh.curY[tid] = nv;
                        else sh.curX[tid] = nv;
                    }
                    __syncthreads();
                    if (P.trace && inLat) P.trace[((size_t)step * P.lh + cy) * P.lw + cx] = sh.win[myWin];
                }
                /* publish this tile's windows for the next level / neighbours / blur */
                if (tid < nWin && sh.cnt[tid]) {
                    const int gx = tx0 / ws + (tid % wpt), gy = ty0 / ws + (tid / wpt);
                    Tcur[gy * nwx + gx] = (uint32_t)(uint16_t)sh.curX[tid] | ((uint32_t)(uint16_t)sh.curY[tid] << 16);
                }
                __syncthreads();
            }
        }
        /* Next level reads other CTAs' windows of this level (neighbour bias) only from iteration 4
         * on, and big-window levels synchronise on their own before they read. */
        const int nws = ws >> 1;
        if (it + 1 < P.iters && (it + 1) >= HR_FIRST_NEIGHBOR_ITERATION && nws <= HR_TILE) grid_barrier(P.bar, barTarget, nCtas);
    }

    /* ------------- expand last level to the raw offset array and blur it (K4) ----------------- */
    grid_barrier(P.bar, barTarget, nCtas);
    {
        const int lws = P.first >> (P.iters - 1); /* = 2 */
        const int lnwx = (P.lw + lws - 1) / lws;
        const uint32_t *Tl = P.T + P.tOff[P.iters - 1];
        int16_t *tX = sh.blur.tX, *tY = sh.blur.tY;
        int *hX = sh.blur.hX, *hY = sh.blur.hY;
        const size_t ln = (size_t)P.lw * P.lh;
        for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas) {
            const int tx0 = (tile % P.tilesX) * HR_TILE, ty0 = (tile / P.tilesX) * HR_TILE;
            for (int i = tid; i < 40 * 40; i += HR_TILE * HR_TILE) {
                const int r = i / 40, c = i % 40;
                int gy = ty0 - 4 + r, gx = tx0 - 4 + c;
                /* blurFlowKernel.cl:5-12 mirror, clamped for lattices smaller than the halo */
                if (gy >= P.lh) gy = 2 * P.lh - gy - 1; else if (gy < 0) gy = -gy - 1;
                if (gx >= P.lw) gx = 2 * P.lw - gx - 1; else if (gx < 0) gx = -gx - 1;
                gy = hr_min(hr_max(gy, 0), P.lh - 1);
                gx = hr_min(hr_max(gx, 0), P.lw - 1);
                const uint32_t v = ldcg_u32(Tl + (gy / lws) * lnwx + (gx / lws));
                tX[i] = (int16_t)(v & 0xffffu);
                tY[i] = (int16_t)(v >> 16);
            }
            __syncthreads();
            for (int i = tid; i < 40 * 32; i += HR_TILE * HR_TILE) {
                const int r = i >> 5, c = i & 31;
                int sx = 0, sy = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    sx += tX[r * 40 + c + k];
                    sy += tY[r * 40 + c + k];
                }
                hX[i] = sx;
                hY[i] = sy;
            }
            __syncthreads();
            const int cx = tx0 + lane, cy = ty0 + warp;
            if (cx < P.lw && cy < P.lh) {
                int sx = 0, sy = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    sx += hX[(warp + k) * 32 + lane];
                    sy += hY[(warp + k) * 32 + lane];
                }
                const size_t idx = (size_t)cy * P.lw + cx;
                P.blur[idx] = (int16_t)(sx / 64);       /* C division truncates toward zero */
                P.blur[ln + idx] = (int16_t)(sy / 64);
                P.off[idx] = tX[(warp + 4) * 40 + lane + 4];
                P.off[ln + idx] = tY[(warp + 4) * 40 + lane + 4];
            }
            __syncthreads();
        }
        /* leave the cross-CTA sums zeroed for the next launch (all consumers passed the barrier) */
        for (int i = blockIdx.x * HR_TILE * HR_TILE + tid; i < P.bigWords; i += nCtas * HR_TILE * HR_TILE) P.bigSums[i] = 0u;
    }
}

/* Stand-alone K4 (parity tap hr_blur_flow): direct 64-tap form of blurFlowKernel.cl:80-88. */
__global__ void blur_flow_kernel(const int16_t *__restrict__ in, int16_t *__restrict__ out, int lh, int lw) {
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = blockIdx.y * blockDim.y + threadIdx.y, gz = blockIdx.z;
    if (gx >= lw || gy >= lh) return;
    const int16_t *src = in + (size_t)gz * lw * lh;
    int sum = 0;
    for (int ky = -4; ky < 4; ++ky)
        for (int kx = -4; kx < 4; ++kx) {
            int y = gy + ky, x = gx + kx;
            if (y >= lh) y = 2 * lh - y - 1; else if (y < 0) y = -y - 1;
            if (x >= lw) x = 2 * lw - x - 1; else if (x < 0) x = -x - 1;
            y = hr_min(hr_max(y, 0), lh - 1);
            x = hr_min(hr_max(x, 0), lw - 1);
            sum += src[(size_t)y * lw + x];
        }
    out[(size_t)gz * lw * lh + (size_t)gy * lw + gx] = (int16_t)(sum / 64);
}

/* ------------------------------------------------------------------------------------------ */
/* warp + flip + blend + levels + output modes (K5)                                              */
/* ------------------------------------------------------------------------------------------ */
/* warpFrameKernel.cl:10-18 */
__device__ __forceinline__ int warp_mirror(int pos, int dim) {
    int res = pos;
    if (pos >= dim - 1) res = pos - ((pos - (dim - 2)) * 2);
    else if (pos < 1) res = -pos + 1;
    return hr_min(hr_max(res, 1), dim - 2);
}
__device__ __forceinline__ unsigned sat_u8(float v) { return __float2uint_rz(fmaxf(fminf(v, 255.0f), 0.0f)); }

/* warpFrameKernel.cl:1-7 */
__device__ __forceinline__ unsigned levels_y8(float v, float black, float white) { return sat_u8((v - black) / (white - black) * 255.0f); }
__device__ __forceinline__ unsigned levels_uv8(float v, float white) { return sat_u8((v - 128.0f) / white * 255.0f + 128.0f); }
/* P010, by construction */
__device__ __forceinline__ unsigned levels_y16(float v, float black, float white) {
    const float b16 = black / 255.0f * 65472.0f, w16 = white / 255.0f * 65472.0f;
    return __float2uint_rz(fmaxf(fminf((v - b16) / (w16 - b16) * 65472.0f, 65472.0f), 0.0f)) & 0xFFC0u;
}
__device__ __forceinline__ unsigned levels_uv16(float v, float white) {
    const float w16 = white / 255.0f * 65472.0f;
    return __float2uint_rz(fmaxf(fminf((v - 32768.0f) / w16 * 65472.0f + 32768.0f, 65472.0f), 0.0f)) & 0xFFC0u;
}

__global__ void levels_lut_kernel(uint8_t *lut, int *identity, float black, float white) {
    const int v = threadIdx.x;
    const unsigned y = levels_y8((float)v, black, white), c = levels_uv8((float)v, white);
    lut[v] = (uint8_t)y;
    lut[256 + v] = (uint8_t)c;
    const int same = __syncthreads_and(y == (unsigned)v && c == (unsigned)v);
    if (v == 0) *identity = same;
}

/* warpFrameKernel.cl:21-111 */
__device__ unsigned visualize_flow(int offsetX, int offsetY, unsigned currPixel, int channel, int resImpact) {
    offsetX = (int)(int16_t)offsetX;
    offsetY = (int)(int16_t)offsetY;
    unsigned r, g, b;
    const int ax = abs(offsetX), ay = abs(offsetY);
    if (ax < 1 && ay < 1) {
        r = g = b = 0;
    } else {
        const float angle_rad = atan2f((float)offsetY, (float)offsetX);
        float angle_deg = angle_rad * (180.0f / 3.14159274101257f);
        if (angle_deg < 0) angle_deg += 360.0f;
        angle_deg = fmodf(angle_deg, 360.0f);
        if (angle_deg < 0) angle_deg += 360.0f;
        const float hue = angle_deg / 360.0f;
        const int h_i = (int)(hue * 6.0f);
        const float f = hue * 6.0f - (float)h_i;
        const float q = 1.0f - f;
        switch (h_i % 6) {
            case 0: r = 255; g = __float2uint_rz(f * 255.0f) & 255u; b = 0; break;
            case 1: r = __float2uint_rz(q * 255.0f) & 255u; g = 255; b = 0; break;
            case 2: r = 0; g = 255; b = __float2uint_rz(f * 255.0f) & 255u; break;
            case 3: r = 0; g = __float2uint_rz(q * 255.0f) & 255u; b = 255; break;
            case 4: r = __float2uint_rz(f * 255.0f) & 255u; g = 0; b = 255; break;
            case 5: r = 255; g = 0; b = __float2uint_rz(q * 255.0f) & 255u; break;
            default: r = g = b = 0; break;
        }
        r = sat_u8((float)r / 255.0f * (float)(ax + ay) * (float)resImpact);
        g = sat_u8((float)g / 255.0f * (float)ay * 2.0f * (float)resImpact);
        b = sat_u8((float)b / 255.0f * (float)(ax + ay) * (float)resImpact);
    }
    if (channel == 0) return ((sat_u8((float)r * 0.299f + (float)g * 0.587f + (float)b * 0.114f) >> 1) + (currPixel >> 1)) & 255u;
    if (channel == 1) return sat_u8((float)r * -0.168736f + (float)g * -0.331264f + (float)b * 0.5f + 128.0f);
    return sat_u8((float)r * 0.5f + (float)g * -0.418688f + (float)b * -0.081312f + 128.0f);
}

template <typename T>
struct SampleTraits;
template <>
struct SampleTraits<uint8_t> {
    static constexpr bool is16 = false;
};
template <>
struct SampleTraits<uint16_t> {
    static constexpr bool is16 = true;
};

/* The flow vectors one output cell needs: o12 at the cell, o21 through the flip indirection
 * (warpFrameKernel.cl:151-156). */
struct CellFlow {
    int x12, y12, x21, y21;
};
template <typename T>
__device__ __forceinline__ CellFlow cell_flow(const WarpParams<T> &P, int adjCx, int adjCy, int cz) {
    const int s = P.s;
    const int scx = cz ? ((adjCx >> s) & ~1) : (adjCx >> s);
    const int scy = cz ? ((adjCy >> s) << 1) : (adjCy >> s);
    const size_t ln = (size_t)P.lw * P.lh;
    CellFlow f;
    f.x12 = __ldg(P.flow + (size_t)scy * P.lw + scx);
    f.y12 = __ldg(P.flow + ln + (size_t)scy * P.lw + scx);
    const int fy = hr_min(hr_max(scy - (f.y12 >> s), 0), P.lh - 1);
    const int fx = hr_min(hr_max(scx - (f.x12 >> s), 0), P.lw - 1);
    f.x21 = __ldg(P.flow + (size_t)fy * P.lw + fx);
    f.y21 = __ldg(P.flow + ln + (size_t)fy * P.lw + fx);
    return f;
}

template <typename T>
__device__ __forceinline__ unsigned finish_blend(const WarpParams<T> &P, unsigned a, unsigned b, int cz, int cx, const CellFlow &f) {
    if (!SampleTraits<T>::is16) {
        unsigned v = __float2uint_rz((float)a * P.t21 + (float)b * P.t12);
        if (P.mode == 3) {
            v = visualize_flow(-f.x12, -f.y12, v & 255u, cz + (cx & (cz ? 1 : 0)), P.s <= 2 ? 4 : 1);
            return cz ? levels_uv8((float)v, P.white) : levels_y8((float)v, P.black, P.white);
        }
        v &= 255u;
        return P.lutIdentity ? v : (unsigned)__ldg(P.lut + (cz ? 256 : 0) + v);
    } else {
        const unsigned v = __float2uint_rz(fminf((float)a * P.t21 + (float)b * P.t12, 65535.0f));
        if (P.mode == 3) {
            const unsigned v8 = visualize_flow(-f.x12, -f.y12, v >> 8, cz + (cx & (cz ? 1 : 0)), P.s <= 2 ? 4 : 1);
            const unsigned l8 = cz ? levels_uv8((float)v8, P.white) : levels_y8((float)v8, P.black, P.white);
            return l8 << 8;
        }
        return cz ? levels_uv16((float)v, P.white) : levels_y16((float)v, P.black, P.white);
    }
}

/* One output sample, every mode: the general path (frame borders, modes 3/4/6, tiny frames). */
template <typename T>
__device__ unsigned warp_sample(const WarpParams<T> &P, int cx, int cy, int cz) {
    const T *s12 = cz ? P.f1uv : P.f1y;
    const T *s21 = cz ? P.f2uv : P.f2y;
    const int dimY = P.H, dimX = P.W, aW = P.aW;
    const int verticalOffset = dimY >> 2;
    int adjCx = cx, adjCy = cy;
    const unsigned neutral = SampleTraits<T>::is16 ? 32768u : 128u;

    if (P.mode == 5 && cx < (aW >> 1)) return s12[(size_t)cy * dimX + cx];
    if (P.mode == 6) {
        const bool inBand = cy >= (verticalOffset >> cz) && cy < ((verticalOffset >> cz) + (dimY >> (1 + cz)));
        if (inBand && cx < (dimX >> 1)) return s12[(size_t)((cy - (verticalOffset >> cz)) << 1) * dimX + (cx << 1) + (cz ? (cx & 1) : 0)];
        if (inBand && cx >= (dimX >> 1) && cx < dimX) {
            adjCx = (cx - (aW >> 1)) << 1;
            adjCy = (cy - (verticalOffset >> cz)) << 1;
        } else {
            return cz ? neutral : 0u;
        }
    }
    const CellFlow f = cell_flow(P, adjCx, adjCy, cz);
    if (P.mode == 4) {
        const unsigned m = (unsigned)(abs(f.x12) + abs(f.y12)) << 2;
        const unsigned v8 = cz ? 128u : (m < 255u ? m : 255u);
        return SampleTraits<T>::is16 ? (v8 << 8) : v8;
    }
    const int dY = cz ? (dimY >> 1) : dimY;
    const float ys = cz ? 0.5f : 1.0f;
    const int nx12 = warp_mirror(adjCx + (int)roundf((float)f.x12 * P.t12), aW);
    const int ny12 = warp_mirror(adjCy + (int)roundf((float)f.y12 * P.t12 * ys), dY);
    const int nx21 = warp_mirror(adjCx - (int)roundf((float)f.x21 * P.t21), aW);
    const int ny21 = warp_mirror(adjCy - (int)roundf((float)f.y21 * P.t21 * ys), dY);
    const size_t i12 = (size_t)ny12 * dimX + (nx12 & (cz ? ~1 : ~0)) + (cx & (cz ? 1 : 0));
    const size_t i21 = (size_t)ny21 * dimX + (nx21 & (cz ? ~1 : ~0)) + (cx & (cz ? 1 : 0));
    if (P.mode == 0) return s12[i12];
    if (P.mode == 1) return s21[i21];
    return finish_blend(P, (unsigned)s12[i12], (unsigned)s21[i21], cz, cx, f);
}

/* Four consecutive samples starting at an arbitrary sample address, from aligned 32-bit loads. */
__device__ __forceinline__ uint32_t load4_u8(const uint8_t *p) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    const unsigned shft = (unsigned)(a & 3) * 8;
    const uint32_t lo = __ldg(q);
    const uint32_t hi = shft ? __ldg(q + 1) : 0u;
    return __funnelshift_r(lo, hi, shft);
}
/* 8 consecutive bytes from an arbitrary byte address (chroma with odd displacement) */
__device__ __forceinline__ uint2 load8_u8(const uint8_t *p) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    const unsigned shft = (unsigned)(a & 3) * 8;
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1);
    const uint32_t w2 = shft ? __ldg(q + 2) : 0u;
    return make_uint2(__funnelshift_r(w0, w1, shft), __funnelshift_r(w1, w2, shft));
}

/* Interior fast path, NV12: the four samples of a 4-aligned quad share one lattice cell
 * (s >= 2), so each source is one translated run. Chroma keeps U/V parity: with an odd
 * displacement d the U bytes come from cx+d-1 and the V bytes from cx+d+1
 * (warpFrameKernel.cl:171 `(newCx & ~1) + (cx & 1)`). Returns false when a border is touched. */
__device__ __forceinline__ bool fetch_quad_u8(const uint8_t *plane, int dimX, int aW, int cx0, int row, int d, int cz, uint32_t &out) {
    if (cx0 + d < 1 || cx0 + 3 + d > aW - 2) return false;
    const uint8_t *base = plane + (size_t)row * dimX;
    if (!cz || !(d & 1)) {
        out = load4_u8(base + cx0 + d);
    } else {
        const uint2 w = load8_u8(base + cx0 + d - 1);
        out = __byte_perm(w.x, w.y, 0x5230);
    }
    return true;
}

template <typename T>
__global__ void __launch_bounds__(256) warp_blend_kernel(const WarpParams<T> P, int useFast) {
    const int cx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int row = blockIdx.y * blockDim.y + threadIdx.y;
    if (cx0 >= P.aW || row >= P.H + (P.H >> 1)) return;
    const int cz = row >= P.H;
    const int cy = cz ? row - P.H : row;
    T *outRow = (cz ? P.outUV : P.outY) + (size_t)cy * P.W;

    if (!SampleTraits<T>::is16 && useFast && P.s >= 2 && cx0 + 3 < P.aW && (P.mode <= 2 || P.mode == 5)) {
        const uint8_t *s12 = (const uint8_t *)(cz ? P.f1uv : P.f1y);
        const uint8_t *s21 = (const uint8_t *)(cz ? P.f2uv : P.f2y);
        const int half = P.aW >> 1;
        if (P.mode == 5 && cx0 + 3 < half) {
            *reinterpret_cast<uint32_t *>(outRow + cx0) = *reinterpret_cast<const uint32_t *>(s12 + (size_t)cy * P.W + cx0);
            return;
        }
        if (!(P.mode == 5 && cx0 < half)) {
            const CellFlow f = cell_flow(P, cx0, cy, cz);
            const int dY = cz ? (P.H >> 1) : P.H;
            const float ys = cz ? 0.5f : 1.0f;
            const int d12 = (int)roundf((float)f.x12 * P.t12), d21 = -(int)roundf((float)f.x21 * P.t21);
            const int ny12 = warp_mirror(cy + (int)roundf((float)f.y12 * P.t12 * ys), dY);
            const int ny21 = warp_mirror(cy - (int)roundf((float)f.y21 * P.t21 * ys), dY);
            uint32_t a = 0, b = 0;
            bool ok = true;
            if (P.mode != 1) ok = fetch_quad_u8(s12, P.W, P.aW, cx0, ny12, d12, cz, a);
            if (ok && P.mode != 0) ok = fetch_quad_u8(s21, P.W, P.aW, cx0, ny21, d21, cz, b);
            if (ok) {
                uint32_t o;
                if (P.mode == 0) o = a;
                else if (P.mode == 1) o = b;
                else {
                    o = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const unsigned av = (a >> (8 * k)) & 255u, bv = (b >> (8 * k)) & 255u;
                        unsigned v = __float2uint_rz((float)av * P.t21 + (float)bv * P.t12) & 255u;
                        if (!P.lutIdentity) v = __ldg(P.lut + (cz ? 256 : 0) + v);
                        o |= v << (8 * k);
                    }
                }
                *reinterpret_cast<uint32_t *>(outRow + cx0) = o;
                return;
            }
        }
    }
    /* general path */
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const int cx = cx0 + k;
        if (cx < P.aW) outRow[cx] = (T)warp_sample(P, cx, cy, cz);
    }
}

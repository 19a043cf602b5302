/*
 * hr_search.cuh — the block-offset search (K1+K2+K3) and the flow blur (K4) as ONE persistent
 * cooperative kernel.
 *
 * Reference semantics: video/filter/HopperRender/Kernels/calcDeltaSumsKernel.cl:34-189,
 * determineLowestLayerKernel.cl:2-22, adjustOffsetArrayKernel.cl:2-18, blurFlowKernel.cl:15-89,
 * driver loop opticalFlowCalc.c:126-203 (66 queue commands -> 1 launch).
 *
 * Work decomposition (B200: 148 SMs x 4 sub-partitions):
 *   - CTA = 4 warps = one 32x32 tile of the lattice; every WARP owns a fixed 16x16 block, every
 *     thread a column of 8 points in it (lane&15 = column, lane>>4 = upper/lower 8 rows). The
 *     480x270 lattice of every 16:9 format gives 135 CTAs <= 148 SMs.
 *   - Windows <= 16 are warp-local: 8 register adds + <= 5 shuffles per layer, no shared memory,
 *     no block barrier. Offsets live in registers across iterations.
 *   - Window = 32: four warp totals meet in shared memory (one block barrier per step).
 *   - Windows >= 64: CTA totals meet in L2 (red.add) + one grid barrier per step.
 *   - Neighbour bias (iteration >= 4) reads the previous level's window table from L2; one grid
 *     barrier per iteration orders it. 11 grid barriers per flow at 480x270.
 *   - Every evaluation is one 32-bit load from the phase-planar packed frame (hr_pack.cuh) and one
 *     VABSDIFF4.U8.ACC; both biases are added once per window as count*bias (mod 2^32 — exact,
 *     because every point of a window shares offset and neighbours).
 */
#pragma once
#include "hr_common.cuh"

#define HR_WARPS 4          /* warps per CTA                                  */
#define HR_BLK 16           /* lattice points per warp-block side             */
#define HR_PPT 8            /* points per thread                              */

struct SearchShared {
    uint32_t warpTot[HR_WARPS][HR_RMAX];          /* per-warp block totals, windows >= 32           */
    int tileOffX[HR_MAX_TILES_PER_CTA], tileOffY[HR_MAX_TILES_PER_CTA];
    int winner;
    union {
        struct {                                  /* blur phase                                     */
            int16_t tX[40 * 40], tY[40 * 40];     /* tile + 4-point halo of the raw offsets          */
            int hX[40 * 32], hY[40 * 32];         /* horizontal 8-tap sums                           */
        } blur;
    };
};

/* Window total for layer z (calcDeltaSumsKernel.cl:99-150 summed over the window, mod 2^32). */
__device__ __forceinline__ uint32_t window_total(uint32_t sad, int c, int curAxis, uint32_t count, bool useNb, int nb0, int nb1,
                                                 int nb2, int nb3, int dS, int nS) {
    const int own = (int)(int16_t)(curAxis + c);
    uint32_t bias = (uint32_t)(uint16_t)abs(own);
    if (useNb) {
        const uint32_t nb = (uint32_t)(uint16_t)abs(nb0 - own) + (uint32_t)(uint16_t)abs(nb1 - own) +
                            (uint32_t)(uint16_t)abs(nb2 - own) + (uint32_t)(uint16_t)abs(nb3 - own);
        bias += nb << nS;
    }
    return (sad << dS) + count * bias;
}

/* The four neighbour offsets of calcDeltaSumsKernel.cl:112-128 for the window whose lattice origin
 * is (x0,y0): positions +-2*ws clamped to the lattice, read from the previous level's table. All
 * points of a window resolve to the same four windows, so one lookup serves the whole window. */
__device__ __forceinline__ void load_neighbours(const FlowParams &P, int it, int ws, int axis, int x0, int y0, int &nb0, int &nb1,
                                                int &nb2, int &nb3) {
    const int pws = ws << 1;
    const int lgp = 31 - __clz(pws);
    const int pnwx = (P.lw + pws - 1) >> lgp;
    const uint32_t *Tp = P.T + P.tOff[it - 1];
    const int yd = hr_min(y0 + pws, P.lh - 1) >> lgp, yu = hr_max(y0 - pws, 0) >> lgp;
    const int xr = hr_min(x0 + pws, P.lw - 1) >> lgp, xl = hr_max(x0 - pws, 0) >> lgp;
    const int xc = x0 >> lgp, yc = y0 >> lgp;
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

/* Per-thread view of its 8 lattice points for one search step. */
struct PointSet {
    int fixedIdx[HR_PPT]; /* word index contributed by the axis that does not move this step */
    int moveBase[HR_PPT]; /* full-resolution coordinate on the searched axis before the layer shift */
    int mulA, mulB;       /* word index of moving coordinate p = (p & m) * mulA + (p >> s) * mulB     */
    int D;                /* frame extent along the searched axis                                    */
    bool interior;        /* warp-uniform: no layer of any point of the warp leaves the frame          */
};

/* |a-b| over the four packed bytes, summed, plus c: one VABSDIFF4.U8.ACC */
__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

/* Set up the PointSet for a step: the fixed axis is mirrored once, the moving axis keeps its base.
 * axis 0: layers move x; axis 1: layers move y. Rows of the thread: yTop .. yTop+7 (clamped). */
template <bool UNIFORM>
__device__ __forceinline__ void make_points(const FlowParams &P, int axis, int cx, int yTop, const int (&offX)[HR_PPT],
                                            const int (&offY)[HR_PPT], int ox, int oy, PointSet &ps) {
    const int s = P.s, m = (1 << s) - 1;
    int lo = 0x7fffffff, hi = -0x7fffffff;
#pragma unroll
    for (int k = 0; k < HR_PPT; ++k) {
        const int cyk = hr_min(yTop + k, P.lh - 1);
        const int oxk = UNIFORM ? ox : offX[k], oyk = UNIFORM ? oy : offY[k];
        if (axis == 0) {
            const int y = search_mirror((cyk << s) + oyk, P.H);
            ps.fixedIdx[k] = ((y & m) << s) * P.planeSize + (y >> s) * P.planePitch;
            ps.moveBase[k] = (cx << s) + oxk;
        } else {
            const int x = search_mirror((cx << s) + oxk, P.W);
            ps.fixedIdx[k] = (x & m) * P.planeSize + (x >> s);
            ps.moveBase[k] = (cyk << s) + oyk;
        }
        lo = hr_min(lo, ps.moveBase[k]);
        hi = hr_max(hi, ps.moveBase[k]);
    }
    ps.mulA = axis ? (P.planeSize << s) : P.planeSize;
    ps.mulB = axis ? P.planePitch : 1;
    ps.D = axis ? P.H : P.W;
    ps.interior = __all_sync(0xffffffffu, lo + P.cand[0] >= 0 && hi + P.cand[P.R - 1] < ps.D);
}

/* Issue the loads of up to HR_ZCHUNK layers x 8 points before anything consumes them: with one
 * warp per SM sub-partition the memory latency is hidden by instruction-level parallelism only. */
__device__ __forceinline__ void load_chunk(const FlowParams &P, const PointSet &ps, int z0, int nz, uint32_t (&v1)[HR_ZCHUNK][HR_PPT]) {
    const int s = P.s, m = (1 << s) - 1;
    if (ps.interior) {
#pragma unroll
        for (int j = 0; j < HR_ZCHUNK; ++j) {
            if (j < nz) {
                const int c = P.cand[z0 + j];
#pragma unroll
                for (int k = 0; k < HR_PPT; ++k) {
                    const int p = ps.moveBase[k] + c;
                    v1[j][k] = __ldg(P.p1 + (ps.fixedIdx[k] + (p & m) * ps.mulA + (p >> s) * ps.mulB));
                }
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < HR_ZCHUNK; ++j) {
            if (j < nz) {
                const int c = P.cand[z0 + j];
#pragma unroll
                for (int k = 0; k < HR_PPT; ++k) {
                    const int p = search_mirror(ps.moveBase[k] + c, ps.D);
                    v1[j][k] = __ldg(P.p1 + (ps.fixedIdx[k] + (p & m) * ps.mulA + (p >> s) * ps.mulB));
                }
            }
        }
    }
}

/* One search step for windows of WS <= 16 lattice points (warp-local).
 * offX/offY: per-point offsets in registers, updated in place. wz[g]: winning layer of the g-th
 * window stacked in the thread's column (for the trace tap). */
template <int WS>
__device__ __forceinline__ void step_small(const FlowParams &P, int it, int axis, int lane, int bx0, int by0, int cx, int yTop,
                                           const uint32_t (&v2)[HR_PPT], unsigned vmask, int (&offX)[HR_PPT], int (&offY)[HR_PPT],
                                           int (&wz)[HR_PPT / (WS < HR_PPT ? WS : HR_PPT)]) {
    constexpr int VG = WS < HR_PPT ? WS : HR_PPT; /* rows of one window held by one thread              */
    constexpr int NG = HR_PPT / VG;                /* windows stacked in the thread's column             */
    constexpr int NF = NG > WS ? NG / WS : 1;      /* windows one lane finalises (2 for WS=2, else 1)    */
    const int lx = lane & 15, half = lane >> 4;
    const int R = P.R;
    const bool useNb = it >= HR_FIRST_NEIGHBOR_ITERATION;
    const int sub = lx & (WS - 1);                 /* lane's column inside its window                    */

    PointSet ps;
    make_points<false>(P, axis, cx, yTop, offX, offY, 0, 0, ps);

    /* the window(s) this lane finalises: slot i handles stacked window g = sub + i*WS (if < NG) */
    int cur[NF], nbA[NF], nbB[NF];                 /* neighbours packed 2 x int16                        */
    uint32_t cnt[NF], bestS[NF];
    int bestZ[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        const int g = sub + i * WS;
        const int x0 = bx0 + (lx & ~(WS - 1));
        const int y0 = by0 + (WS == 16 ? 0 : half * 8 + g * VG);
        cnt[i] = 0;
        nbA[i] = nbB[i] = 0;
        cur[i] = 0;
        bestS[i] = 0xffffffffu;
        bestZ[i] = 0;
        if (g < NG && x0 < P.lw && y0 < P.lh) {
            cnt[i] = (uint32_t)(hr_min(x0 + WS, P.lw) - x0) * (uint32_t)(hr_min(y0 + WS, P.lh) - y0);
            if (useNb) {
                int n0, n1, n2, n3;
                load_neighbours(P, it, WS, axis, x0, y0, n0, n1, n2, n3);
                nbA[i] = (n0 & 0xffff) | (n1 << 16);
                nbB[i] = (n2 & 0xffff) | (n3 << 16);
            }
        }
#pragma unroll
        for (int gg = 0; gg < NG; ++gg)
            if (gg == g) cur[i] = axis ? offY[gg * VG] : offX[gg * VG];
    }

    for (int z0 = 0; z0 < R; z0 += HR_ZCHUNK) {
        const int nz = hr_min(HR_ZCHUNK, R - z0);
        uint32_t v1[HR_ZCHUNK][HR_PPT];
        load_chunk(P, ps, z0, nz, v1);
        uint32_t sg[HR_ZCHUNK][NG];
        /* thread-local column sums */
#pragma unroll
        for (int j = 0; j < HR_ZCHUNK; ++j) {
            if (j < nz) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    uint32_t a = 0;
#pragma unroll
                    for (int r = 0; r < VG; ++r) {
                        const int k = g * VG + r;
                        a = sad4_acc(((vmask >> k) & 1u) ? v1[j][k] : v2[k], v2[k], a);
                    }
                    sg[j][g] = a;
                }
            }
        }
        /* butterflies across the window's columns (independent chains for every layer) */
#pragma unroll
        for (int o = 1; o < WS && o < 16; o <<= 1) {
#pragma unroll
            for (int j = 0; j < HR_ZCHUNK; ++j)
                if (j < nz) {
#pragma unroll
                    for (int g = 0; g < NG; ++g) sg[j][g] += __shfl_xor_sync(0xffffffffu, sg[j][g], o);
                }
        }
        if (WS == 16) {
#pragma unroll
            for (int j = 0; j < HR_ZCHUNK; ++j)
                if (j < nz) sg[j][0] += __shfl_xor_sync(0xffffffffu, sg[j][0], 16);
        }
        /* biases + first-minimum scan (determineLowestLayerKernel.cl:13-18) */
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            if (cnt[i]) {
#pragma unroll
                for (int j = 0; j < HR_ZCHUNK; ++j) {
                    if (j < nz) {
                        uint32_t sad = sg[j][(NG > 1) ? i * WS : 0];
                        if (NG > 1) { /* window sub + i*WS: a register select, sub < WS <= 4 here */
#pragma unroll
                            for (int q = 1; q < WS && i * WS + q < NG; ++q)
                                if (sub == q) sad = sg[j][i * WS + q];
                        }
                        const uint32_t S = window_total(sad, P.cand[z0 + j], cur[i], cnt[i], useNb, (int)(int16_t)nbA[i], nbA[i] >> 16,
                                                        (int)(int16_t)nbB[i], nbB[i] >> 16, P.dS, P.nS);
                        if (z0 + j == 0 || S < bestS[i]) {
                            bestS[i] = S;
                            bestZ[i] = z0 + j;
                        }
                    }
                }
            }
        }
    }
    /* winners back to every lane of the window, then the offset update (adjustOffsetArrayKernel.cl:11-17) */
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const int src = (lane & ~(WS - 1) & (WS == 16 ? 0 : 31)) | (g & (WS - 1));
        wz[g] = __shfl_sync(0xffffffffu, bestZ[g / WS < NF ? g / WS : 0], src);
        const int upd = P.cand[wz[g]];
#pragma unroll
        for (int r = 0; r < VG; ++r) {
            if (axis) offY[g * VG + r] += upd;
            else offX[g * VG + r] += upd;
        }
    }
}

/* Per-warp block totals (windows >= 32): sum of the warp's 256 points for every layer, lane z
 * ends up holding layer z. Offsets are uniform over the tile. */
__device__ __forceinline__ uint32_t block_totals(const FlowParams &P, int axis, int lane, int cx, int yTop, const uint32_t (&v2)[HR_PPT],
                                                 unsigned vmask, int ox, int oy) {
    const int dummy[HR_PPT] = {0, 0, 0, 0, 0, 0, 0, 0};
    PointSet ps;
    make_points<true>(P, axis, cx, yTop, dummy, dummy, ox, oy, ps);
    const int R = P.R;
    uint32_t mineTot = 0;
    for (int z0 = 0; z0 < R; z0 += HR_ZCHUNK) {
        const int nz = hr_min(HR_ZCHUNK, R - z0);
        uint32_t v1[HR_ZCHUNK][HR_PPT];
        load_chunk(P, ps, z0, nz, v1);
#pragma unroll
        for (int j = 0; j < HR_ZCHUNK; ++j) {
            if (j < nz) {
                uint32_t a = 0;
#pragma unroll
                for (int k = 0; k < HR_PPT; ++k) a = sad4_acc(((vmask >> k) & 1u) ? v1[j][k] : v2[k], v2[k], a);
                a = __reduce_add_sync(0xffffffffu, a);
                if (lane == z0 + j) mineTot = a;
            }
        }
    }
    return mineTot;
}

/* Executed by one full warp: lane z holds the window's SAD for layer z; returns the winner. */
__device__ __forceinline__ int finalize_warp(const FlowParams &P, int it, int ws, int axis, int lane, uint32_t sad, int x0, int y0, int cur) {
    const int R = P.R;
    const bool useNb = it >= HR_FIRST_NEIGHBOR_ITERATION;
    const uint32_t count = (uint32_t)(hr_min(x0 + ws, P.lw) - x0) * (uint32_t)(hr_min(y0 + ws, P.lh) - y0);
    int nb0 = 0, nb1 = 0, nb2 = 0, nb3 = 0;
    if (useNb) load_neighbours(P, it, ws, axis, x0, y0, nb0, nb1, nb2, nb3);
    const uint32_t S = (lane < R) ? window_total(sad, P.cand[lane < R ? lane : 0], cur, count, useNb, nb0, nb1, nb2, nb3, P.dS, P.nS) : 0xffffffffu;
    const uint32_t mn = __reduce_min_sync(0xffffffffu, S);
    const unsigned ballot = __ballot_sync(0xffffffffu, S == mn && lane < R);
    return __ffs(ballot) - 1;
}

/* Per-tile thread geometry */
struct TileGeom {
    int tx0, ty0, bx0, by0, cx, yTop;
    unsigned vmask;
};
__device__ __forceinline__ TileGeom tile_geom(const FlowParams &P, int tile, int warp, int lx, int half) {
    TileGeom g;
    g.tx0 = (tile % P.tilesX) * HR_TILE;
    g.ty0 = (tile / P.tilesX) * HR_TILE;
    g.bx0 = g.tx0 + (warp & 1) * HR_BLK;
    g.by0 = g.ty0 + (warp >> 1) * HR_BLK;
    g.cx = hr_min(g.bx0 + lx, P.lw - 1);
    g.yTop = g.by0 + half * 8;
    g.vmask = 0;
#pragma unroll
    for (int k = 0; k < HR_PPT; ++k)
        if (g.bx0 + lx < P.lw && g.yTop + k < P.lh) g.vmask |= 1u << k;
    return g;
}
__device__ __forceinline__ void load_frame2(const FlowParams &P, const TileGeom &g, uint32_t (&v2)[HR_PPT]) {
#pragma unroll
    for (int k = 0; k < HR_PPT; ++k) v2[k] = __ldg(P.p2 + hr_min(g.yTop + k, P.lh - 1) * P.planePitch + g.cx);
}

__global__ void __launch_bounds__(HR_WARPS * 32, 1) flow_search_kernel(const FlowParams P) {
    __shared__ SearchShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lx = lane & 15, half = lane >> 4;
    const unsigned nCtas = gridDim.x;
    unsigned long long barTarget = P.barBase;
    const int R = P.R;
    const bool multi = P.numTiles > (int)nCtas;
    const size_t ln = (size_t)P.lw * P.lh;

    if (tid < HR_MAX_TILES_PER_CTA) {
        sh.tileOffX[tid] = 0;
        sh.tileOffY[tid] = 0;
    }
    __syncthreads();

    /* register state of the tile this CTA owns (re-loaded per tile when a CTA owns several) */
    int offX[HR_PPT], offY[HR_PPT];
    uint32_t v2[HR_PPT];
#pragma unroll
    for (int k = 0; k < HR_PPT; ++k) offX[k] = offY[k] = 0;
    TileGeom tg = tile_geom(P, blockIdx.x, warp, lx, half);
    load_frame2(P, tg, v2);
    bool smallStarted = false;

    for (int it = 0; it < P.iters; ++it) {
        const int ws = P.first >> it;
        const int lgw = 31 - __clz(ws);
        const int nwx = (P.lw + ws - 1) >> lgw;
        uint32_t *Tcur = P.T + P.tOff[it];

        if (ws >= HR_TILE) {
            const bool big = ws > HR_TILE;
            for (int axis = 0; axis < 2; ++axis) {
                const int step = it * 2 + axis;
                int slot = 0;
                for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas, ++slot) {
                    if (multi) {
                        tg = tile_geom(P, tile, warp, lx, half);
                        load_frame2(P, tg, v2);
                    }
                    const int ox = sh.tileOffX[slot], oy = sh.tileOffY[slot];
                    const uint32_t tot = block_totals(P, axis, lane, tg.cx, tg.yTop, v2, tg.vmask, ox, oy);
                    sh.warpTot[warp][lane] = tot;
                    __syncthreads();
                    const int wx = tg.tx0 >> lgw, wy = tg.ty0 >> lgw;
                    if (warp == 0) {
                        const uint32_t t4 = sh.warpTot[0][lane] + sh.warpTot[1][lane] + sh.warpTot[2][lane] + sh.warpTot[3][lane];
                        if (big) {
                            if (lane < R) atomicAdd(P.bigSums + P.bigOff[step] + (wy * nwx + wx) * HR_RMAX + lane, t4);
                        } else {
                            const int cur = axis ? oy : ox;
                            const int winner = finalize_warp(P, it, ws, axis, lane, t4, wx << lgw, wy << lgw, cur);
                            if (lane == 0) {
                                if (axis) sh.tileOffY[slot] = cur + P.cand[winner];
                                else sh.tileOffX[slot] = cur + P.cand[winner];
                                sh.winner = winner;
                            }
                        }
                    }
                    __syncthreads();
                    if (!big) {
                        if (P.trace) {
#pragma unroll
                            for (int k = 0; k < HR_PPT; ++k)
                                if ((tg.vmask >> k) & 1u) P.trace[((size_t)step * P.lh + tg.yTop + k) * P.lw + tg.cx] = (uint8_t)sh.winner;
                        }
                        if (axis == 1 && tid == 0) Tcur[wy * nwx + wx] = (uint32_t)(uint16_t)sh.tileOffX[slot] | ((uint32_t)(uint16_t)sh.tileOffY[slot] << 16);
                        __syncthreads();
                    }
                }
                if (big) {
                    grid_barrier(P.bar, barTarget, nCtas);
                    slot = 0;
                    for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas, ++slot) {
                        if (multi) tg = tile_geom(P, tile, warp, lx, half);
                        const int wx = tg.tx0 >> lgw, wy = tg.ty0 >> lgw;
                        if (warp == 0) {
                            const int cur = axis ? sh.tileOffY[slot] : sh.tileOffX[slot];
                            const uint32_t sad = (lane < R) ? ldcg_u32(P.bigSums + P.bigOff[step] + (wy * nwx + wx) * HR_RMAX + lane) : 0u;
                            const int winner = finalize_warp(P, it, ws, axis, lane, sad, wx << lgw, wy << lgw, cur);
                            if (lane == 0) {
                                if (axis) sh.tileOffY[slot] = cur + P.cand[winner];
                                else sh.tileOffX[slot] = cur + P.cand[winner];
                                sh.winner = winner;
                            }
                        }
                        __syncthreads();
                        if (P.trace) {
#pragma unroll
                            for (int k = 0; k < HR_PPT; ++k)
                                if ((tg.vmask >> k) & 1u) P.trace[((size_t)step * P.lh + tg.yTop + k) * P.lw + tg.cx] = (uint8_t)sh.winner;
                        }
                        if (axis == 1 && tid == 0) Tcur[wy * nwx + wx] = (uint32_t)(uint16_t)sh.tileOffX[slot] | ((uint32_t)(uint16_t)sh.tileOffY[slot] << 16);
                        __syncthreads();
                    }
                }
            }
        } else {
            /* ---------------- warp-local windows ------------------------------------------------- */
            int slot = 0;
            for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas, ++slot) {
                if (multi) {
                    tg = tile_geom(P, tile, warp, lx, half);
                    load_frame2(P, tg, v2);
                }
                if (!smallStarted) {
                    /* first warp-local level: every point inherits the tile's offset */
#pragma unroll
                    for (int k = 0; k < HR_PPT; ++k) {
                        offX[k] = sh.tileOffX[slot];
                        offY[k] = sh.tileOffY[slot];
                    }
                } else if (multi) {
#pragma unroll
                    for (int k = 0; k < HR_PPT; ++k) {
                        const size_t idx = (size_t)hr_min(tg.yTop + k, P.lh - 1) * P.lw + tg.cx;
                        offX[k] = P.off[idx];
                        offY[k] = P.off[ln + idx];
                    }
                }
                for (int axis = 0; axis < 2; ++axis) {
                    int wzk[HR_PPT]; /* winner per point, for the trace only */
                    if (ws == 16) {
                        int wz[1];
                        step_small<16>(P, it, axis, lane, tg.bx0, tg.by0, tg.cx, tg.yTop, v2, tg.vmask, offX, offY, wz);
#pragma unroll
                        for (int k = 0; k < HR_PPT; ++k) wzk[k] = wz[0];
                    } else if (ws == 8) {
                        int wz[1];
                        step_small<8>(P, it, axis, lane, tg.bx0, tg.by0, tg.cx, tg.yTop, v2, tg.vmask, offX, offY, wz);
#pragma unroll
                        for (int k = 0; k < HR_PPT; ++k) wzk[k] = wz[0];
                    } else if (ws == 4) {
                        int wz[2];
                        step_small<4>(P, it, axis, lane, tg.bx0, tg.by0, tg.cx, tg.yTop, v2, tg.vmask, offX, offY, wz);
#pragma unroll
                        for (int k = 0; k < HR_PPT; ++k) wzk[k] = wz[k >> 2];
                    } else {
                        int wz[4];
                        step_small<2>(P, it, axis, lane, tg.bx0, tg.by0, tg.cx, tg.yTop, v2, tg.vmask, offX, offY, wz);
#pragma unroll
                        for (int k = 0; k < HR_PPT; ++k) wzk[k] = wz[k >> 1];
                    }
                    if (P.trace) {
#pragma unroll
                        for (int k = 0; k < HR_PPT; ++k)
                            if ((tg.vmask >> k) & 1u) P.trace[((size_t)(it * 2 + axis) * P.lh + tg.yTop + k) * P.lw + tg.cx] = (uint8_t)wzk[k];
                    }
                }
                /* publish this level's windows (neighbours / next level / blur) */
#pragma unroll
                for (int k = 0; k < HR_PPT; ++k) {
                    const int y = tg.yTop + k;
                    if (((tg.vmask >> k) & 1u) && ((tg.bx0 + lx) & (ws - 1)) == 0 && (y & (ws - 1)) == 0)
                        Tcur[(y >> lgw) * nwx + ((tg.bx0 + lx) >> lgw)] = (uint32_t)(uint16_t)offX[k] | ((uint32_t)(uint16_t)offY[k] << 16);
                }
                if (multi || it == P.iters - 1) {
#pragma unroll
                    for (int k = 0; k < HR_PPT; ++k)
                        if ((tg.vmask >> k) & 1u) {
                            const size_t idx = (size_t)(tg.yTop + k) * P.lw + tg.cx;
                            P.off[idx] = (int16_t)offX[k];
                            P.off[ln + idx] = (int16_t)offY[k];
                        }
                }
            }
            smallStarted = true;
        }
        const int nws = ws >> 1;
        if (it + 1 < P.iters && (it + 1) >= HR_FIRST_NEIGHBOR_ITERATION && nws <= HR_TILE) grid_barrier(P.bar, barTarget, nCtas);
    }

    /* ------------- blur the raw offsets (K4), reading the last level's window table --------------- */
    grid_barrier(P.bar, barTarget, nCtas);
    {
        const int lws = P.first >> (P.iters - 1); /* = 2 */
        const int lgl = 31 - __clz(lws);
        const int lnwx = (P.lw + lws - 1) >> lgl;
        const uint32_t *Tl = P.T + P.tOff[P.iters - 1];
        int16_t *tX = sh.blur.tX, *tY = sh.blur.tY;
        int *hX = sh.blur.hX, *hY = sh.blur.hY;
        constexpr int NT = HR_WARPS * 32;
        for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas) {
            const int tx0 = (tile % P.tilesX) * HR_TILE, ty0 = (tile / P.tilesX) * HR_TILE;
#pragma unroll
            for (int u = 0; u < (40 * 40 + NT - 1) / NT; ++u) {
                const int i = tid + u * NT;
                if (i >= 40 * 40) break;
                const int r = i / 40, c = i - r * 40;
                int gy = ty0 - 4 + r, gx = tx0 - 4 + c;
                /* blurFlowKernel.cl:5-12 mirror, clamped for lattices smaller than the halo */
                if (gy >= P.lh) gy = 2 * P.lh - gy - 1; else if (gy < 0) gy = -gy - 1;
                if (gx >= P.lw) gx = 2 * P.lw - gx - 1; else if (gx < 0) gx = -gx - 1;
                gy = hr_min(hr_max(gy, 0), P.lh - 1);
                gx = hr_min(hr_max(gx, 0), P.lw - 1);
                const uint32_t v = ldcg_u32(Tl + (gy >> lgl) * lnwx + (gx >> lgl));
                tX[i] = (int16_t)(v & 0xffffu);
                tY[i] = (int16_t)(v >> 16);
            }
            __syncthreads();
            for (int i = tid; i < 40 * 32; i += NT) {
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
            for (int i = tid; i < 32 * 32; i += NT) {
                const int r = i >> 5, c = i & 31;
                const int x = tx0 + c, y = ty0 + r;
                if (x < P.lw && y < P.lh) {
                    int sx = 0, sy = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        sx += hX[(r + k) * 32 + c];
                        sy += hY[(r + k) * 32 + c];
                    }
                    const size_t idx = (size_t)y * P.lw + x;
                    P.blur[idx] = (int16_t)(sx / 64); /* C division truncates toward zero */
                    P.blur[ln + idx] = (int16_t)(sy / 64);
                }
            }
            __syncthreads();
        }
        /* leave the cross-CTA sums zeroed for the next launch (all consumers passed the barrier) */
        for (int i = blockIdx.x * NT + tid; i < P.bigWords; i += nCtas * NT) P.bigSums[i] = 0u;
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

/*
 * hr_search.cuh — the block-offset search (K1+K2+K3) and the flow blur (K4) as ONE persistent
 * cooperative kernel.
 *
 * Reference semantics: video/filter/HopperRender/Kernels/calcDeltaSumsKernel.cl:34-189,
 * determineLowestLayerKernel.cl:2-22, adjustOffsetArrayKernel.cl:2-18, blurFlowKernel.cl:15-89,
 * driver loop opticalFlowCalc.c:126-203 (66 queue commands -> 1 launch).
 *
 * Work decomposition (B200: 148 SMs x 4 sub-partitions):
 *   - CTA = 16 warps = one 32x32 tile of the lattice (480x270 -> 135 CTAs <= 148 SMs, one per SM,
 *     4 warps per sub-partition to hide the L2 latency of the sample loads).
 *   - warp = one 8x8 block of the tile; thread = two vertically adjacent points of it
 *     (lane&7 = column, lane>>3 = row pair). Both points always share a window (window >= 2), so a
 *     thread carries ONE offset pair in registers through all 16 steps.
 *   - Windows <= 8 are warp-local: xor-butterflies (window 8: one REDUX), every lane of a window
 *     scores the layers itself. No shared memory, no block barrier.
 *   - Windows 16 / 32: warp REDUX totals meet in shared memory; one warp per window scores the
 *     layers (lane = layer). Two block barriers per step.
 *   - Windows >= 64 span CTAs: tile totals meet in L2 (red.add) + one grid barrier per step.
 *   - Neighbour bias (iteration >= 4) reads the previous level's window table from L2; one grid
 *     barrier per level orders it.
 *   - Every evaluation is one 32-bit load from the phase-planar packed frame (hr_pack.cuh) and one
 *     VABSDIFF4.U8.ACC; both biases are added once per window as count*bias (mod 2^32 — exact,
 *     because every point of a window shares offset and neighbours).
 */
#pragma once
#include "hr_common.cuh"

#define HR_THREADS 512      /* threads per CTA                                 */
#define HR_NWARPS 16        /* warps per CTA = 8x8 blocks per tile             */

struct SearchShared {
    uint32_t warpTot[HR_NWARPS][HR_RMAX];        /* per-warp block totals, windows >= 16            */
    int winner[4];                               /* winning layer of the tile's window(s)           */
    union {
        int4 parked[HR_MAX_TILES_PER_CTA][HR_THREADS]; /* MULTI only: per-thread state of each owned tile */
        struct {                                  /* blur phase                                      */
            int16_t tX[40 * 40], tY[40 * 40];     /* tile + 4-point halo of the raw offsets           */
            int hX[40 * 32], hY[40 * 32];         /* horizontal 8-tap sums                            */
        } blur;
    };
};

/* |a-b| over the four packed bytes, summed, plus c: one VABSDIFF4.U8.ACC */
__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

/* Window total for layer shift c (calcDeltaSumsKernel.cl:99-150 summed over the window, mod 2^32):
 * (SAD << deltaScalar) + count * (|own| + (sum of the four |neighbour - own|) << neighborBiasScalar).
 * The reference takes the magnitudes on 16-bit values; offsets never leave +-(16 levels x 16^2), so
 * plain 32-bit |a-b| (one VABSDIFF each) is the same number. n[]: the four neighbour offsets. */
__device__ __forceinline__ uint32_t window_total(uint32_t sad, int c, int cur, uint32_t count, bool useNb, const int (&n)[4], int dS, int nS) {
    const int own = cur + c;
    uint32_t bias = (uint32_t)abs(own);
    if (useNb) bias += __sad(n[3], own, __sad(n[2], own, __sad(n[1], own, __sad(n[0], own, 0u)))) << nS;
    return (sad << dS) + count * bias;
}

/* The four neighbour windows of calcDeltaSumsKernel.cl:112-128 for the window whose lattice origin
 * is (x0,y0): positions +-2*ws clamped to the lattice, read from the previous level's table (words
 * hold x | y << 16). All points of a window resolve to the same four windows, so one lookup serves
 * the whole window and both axis steps of the level. */
__device__ __forceinline__ void load_neighbours(const FlowParams &P, int it, int ws, int x0, int y0, uint32_t (&nw)[4]) {
    const int pws = ws << 1;
    const int lgp = 31 - __clz(pws);
    const int pnwx = (P.lw + pws - 1) >> lgp;
    const unsigned long long *Tp = P.T + P.tOff[it - 1];
    const int yd = hr_min(y0 + pws, P.lh - 1) >> lgp, yu = hr_max(y0 - pws, 0) >> lgp;
    const int xr = hr_min(x0 + pws, P.lw - 1) >> lgp, xl = hr_max(x0 - pws, 0) >> lgp;
    const int xc = x0 >> lgp, yc = y0 >> lgp;
    const unsigned long long *q[4] = {Tp + yd * pnwx + xc /* down */, Tp + yc * pnwx + xr /* right */, Tp + yc * pnwx + xl /* left */,
                                      Tp + yu * pnwx + xc /* up */};
    unsigned long long v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = ld_relaxed_u64(q[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        while ((uint32_t)(v[i] >> 32) != P.epoch) v[i] = ld_relaxed_u64(q[i]); /* the producing tile has not got there yet */
        nw[i] = (uint32_t)v[i];
    }
}
/* pick one axis out of the four neighbour words */
__device__ __forceinline__ void neighbour_axis(const uint32_t (&nw)[4], int axis, int (&n)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) n[i] = axis ? ((int)nw[i] >> 16) : (int)(int16_t)nw[i];
}

/* Per-thread geometry inside a tile */
struct TileGeom {
    int tx0, ty0;         /* lattice origin of the tile                                    */
    int px, py;           /* lattice position of the thread's upper point (clamped)         */
    uint32_t m0, m1;      /* all-ones if the upper / lower point is inside the lattice      */
};
__device__ __forceinline__ TileGeom tile_geom(const FlowParams &P, int tile, int warp, int lane) {
    TileGeom g;
    g.tx0 = (tile % P.tilesX) * HR_TILE;
    g.ty0 = (tile / P.tilesX) * HR_TILE;
    const int x = g.tx0 + (warp & 3) * 8 + (lane & 7);
    const int y = g.ty0 + (warp >> 2) * 8 + (lane >> 3) * 2;
    g.m0 = (x < P.lw && y < P.lh) ? 0xffffffffu : 0u;
    g.m1 = (x < P.lw && y + 1 < P.lh) ? 0xffffffffu : 0u;
    g.px = x;
    g.py = y;
    return g;
}

/* Layer shift of layer z: calcDeltaSumsKernel.cl:68-72. RT > 0: search radius known at compile time. */
template <int RT>
__device__ __forceinline__ int layer_shift(const FlowParams &P, int z) {
    if (RT > 0) {
        const int rel = z - RT / 2;
        return rel * (rel < 0 ? -rel : rel);
    }
    return P.cand[z];
}

/* SAD of the thread's two points for layers z0 .. z0+HR_ZCHUNK-1 (those below R) of one search step.
 * Packed word of full-resolution sample (x,y): plane ((y&m)<<s | (x&m)), row y>>s, column x>>s.
 * INTERIOR (warp-uniform): no layer of any lane leaves the frame on the searched axis -> no mirror. */
template <int RT, bool INTERIOR>
__device__ __forceinline__ void eval_chunk(const FlowParams &P, int R, int axis, int fa, int fb, int ma, int mb, uint32_t m0, uint32_t m1, uint32_t v2a,
                                           uint32_t v2b, int z0, uint32_t (&acc)[HR_ZCHUNK]) {
    const int s = P.s, m = (1 << s) - 1;
    const int mulA = axis ? (P.planeSize << s) : P.planeSize;
    const int mulB = axis ? P.planePitch : 1;
    const int D = axis ? P.H : P.W;
    uint32_t va[HR_ZCHUNK], vb[HR_ZCHUNK];
#pragma unroll
    for (int j = 0; j < HR_ZCHUNK; ++j) {
        if (z0 + j < R) {
            const int c = layer_shift<RT>(P, z0 + j);
            int pa = ma + c, pb = mb + c;
            if (!INTERIOR) {
                pa = search_mirror(pa, D);
                pb = search_mirror(pb, D);
            }
            va[j] = __ldg(P.p1 + (fa + (pa & m) * mulA + (pa >> s) * mulB));
            vb[j] = __ldg(P.p1 + (fb + (pb & m) * mulA + (pb >> s) * mulB));
        }
    }
#pragma unroll
    for (int j = 0; j < HR_ZCHUNK; ++j)
        if (z0 + j < R) acc[j] = sad4_acc(vb[j] & m1, v2b, sad4_acc(va[j] & m0, v2a, 0u));
}

/* Executed by one full warp: lane z holds the window's SAD for layer z; returns the winner. */
__device__ __forceinline__ int finalize_warp(const FlowParams &P, int R, int it, int ws, int axis, int lane, uint32_t sad, int x0, int y0, int cur) {
    const bool useNb = it >= HR_FIRST_NEIGHBOR_ITERATION;
    const uint32_t count = (uint32_t)(hr_min(x0 + ws, P.lw) - x0) * (uint32_t)(hr_min(y0 + ws, P.lh) - y0);
    int n[4] = {0, 0, 0, 0};
    if (useNb) {
        uint32_t nw[4];
        load_neighbours(P, it, ws, x0, y0, nw);
        neighbour_axis(nw, axis, n);
    }
    const uint32_t S = (lane < R) ? window_total(sad, P.cand[lane < R ? lane : 0], cur, count, useNb, n, P.dS, P.nS) : 0xffffffffu;
    const uint32_t mn = __reduce_min_sync(0xffffffffu, S);
    const unsigned ballot = __ballot_sync(0xffffffffu, S == mn && lane < R);
    return __ffs(ballot) - 1;
}

__device__ __forceinline__ void trace_store(const FlowParams &P, const TileGeom &g, int step, int winner) {
    if (P.trace) {
        if (g.m0) P.trace[((size_t)step * P.lh + g.py) * P.lw + g.px] = (uint8_t)winner;
        if (g.m1) P.trace[((size_t)step * P.lh + g.py + 1) * P.lw + g.px] = (uint8_t)winner;
    }
}

template <int RT, bool MULTI>
__global__ void __launch_bounds__(HR_THREADS, 1) flow_search_kernel(const FlowParams P) {
    __shared__ SearchShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned nCtas = gridDim.x;
    const int R = RT > 0 ? RT : P.R;
    const int s = P.s, m = (1 << s) - 1;
    const size_t ln = (size_t)P.lw * P.lh;
    int stampIdx = 0;
#define HR_STAMP()                                                                                         \
    if (P.timeline && tid == 0 && stampIdx < HR_TIMELINE_SLOTS) P.timeline[blockIdx.x * HR_TIMELINE_SLOTS + stampIdx++] = clock64();
    HR_STAMP();

    /* per-thread state of the owned tile: offset pair and the two frame2 words */
    int ox = 0, oy = 0;
    uint32_t v2a = 0, v2b = 0;
    TileGeom g = tile_geom(P, blockIdx.x, warp, lane);

    auto load_frame2 = [&](const TileGeom &gg) {
        const int cx = hr_min(gg.px, P.lw - 1);
        v2a = __ldg(P.p2 + hr_min(gg.py, P.lh - 1) * P.planePitch + cx) & gg.m0;
        v2b = __ldg(P.p2 + hr_min(gg.py + 1, P.lh - 1) * P.planePitch + cx) & gg.m1;
    };
    if (!MULTI) {
        load_frame2(g);
    } else {
        int slot = 0;
        for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas, ++slot) {
            g = tile_geom(P, tile, warp, lane);
            load_frame2(g);
            sh.parked[slot][tid] = make_int4(0, 0, (int)v2a, (int)v2b);
        }
    }
#define HR_UNPARK(slot_, tile_)                                   \
    if (MULTI) {                                                  \
        g = tile_geom(P, tile_, warp, lane);                      \
        const int4 st_ = sh.parked[slot_][tid];                   \
        ox = st_.x; oy = st_.y; v2a = (uint32_t)st_.z; v2b = (uint32_t)st_.w; \
    }
#define HR_PARK(slot_) \
    if (MULTI) sh.parked[slot_][tid] = make_int4(ox, oy, (int)v2a, (int)v2b);

    for (int it = 0; it < P.iters; ++it) {
        const int ws = P.first >> it;
        const int lgw = 31 - __clz(ws);
        const int nwx = (P.lw + ws - 1) >> lgw;
        unsigned long long *Tcur = P.T + P.tOff[it];
        const bool useNb = it >= HR_FIRST_NEIGHBOR_ITERATION;
        const bool small = ws <= 8;      /* warp-local windows                         */
        const bool big = ws > HR_TILE;   /* windows spanning several tiles (CTAs)      */
        uint32_t nw[4] = {0u, 0u, 0u, 0u};

        for (int axis = 0; axis < 2; ++axis) {
            const int step = it * 2 + axis;
            int slot = 0;
            for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas, ++slot) {
                HR_UNPARK(slot, tile);
                HR_STAMP(); /* step start (after the neighbour-table wait) */
                const int cur = axis ? oy : ox;
                /* the thread's own window (warp-local levels only) */
                const int x0 = g.px & ~(ws - 1), y0 = g.py & ~(ws - 1);
                uint32_t count = 0;
                int nb[4] = {0, 0, 0, 0};
                if (small) {
                    if (x0 < P.lw && y0 < P.lh) {
                        count = (uint32_t)(hr_min(x0 + ws, P.lw) - x0) * (uint32_t)(hr_min(y0 + ws, P.lh) - y0);
                        if (useNb && (axis == 0 || MULTI)) load_neighbours(P, it, ws, x0, y0, nw);
                    }
                    neighbour_axis(nw, axis, nb);
                }
                /* sample addressing of this step: fixed part (the axis that does not move) and the moving
                 * coordinate before the layer shift, for the upper (a) and lower (b) point */
                const int cx = hr_min(g.px, P.lw - 1), cy0 = hr_min(g.py, P.lh - 1), cy1 = hr_min(g.py + 1, P.lh - 1);
                int fa, fb, ma, mb;
                if (axis == 0) {
                    const int ya = search_mirror((cy0 << s) + oy, P.H), yb = search_mirror((cy1 << s) + oy, P.H);
                    fa = ((ya & m) << s) * P.planeSize + (ya >> s) * P.planePitch;
                    fb = ((yb & m) << s) * P.planeSize + (yb >> s) * P.planePitch;
                    ma = mb = (cx << s) + ox;
                } else {
                    const int x = search_mirror((cx << s) + ox, P.W);
                    fa = fb = (x & m) * P.planeSize + (x >> s);
                    ma = (cy0 << s) + oy;
                    mb = (cy1 << s) + oy;
                }
                const int cmin = layer_shift<RT>(P, 0), cmax = layer_shift<RT>(P, R - 1);
                const bool interior = __all_sync(0xffffffffu, hr_min(ma, mb) + cmin >= 0 && hr_max(ma, mb) + cmax < (axis ? P.H : P.W));

                uint32_t bestS = 0xffffffffu, mine = 0u;
                int winner = 0;
                auto do_chunk = [&](int z0) {
                    uint32_t acc[HR_ZCHUNK];
                    if (interior) eval_chunk<RT, true>(P, R, axis, fa, fb, ma, mb, g.m0, g.m1, v2a, v2b, z0, acc);
                    else eval_chunk<RT, false>(P, R, axis, fa, fb, ma, mb, g.m0, g.m1, v2a, v2b, z0, acc);
#pragma unroll
                    for (int j = 0; j < HR_ZCHUNK; ++j) {
                        if (z0 + j < R) {
                            uint32_t a = acc[j];
                            if (ws >= 8) {
                                a = __reduce_add_sync(0xffffffffu, a);
                            } else {
                                a += __shfl_xor_sync(0xffffffffu, a, 1);
                                if (ws == 4) {
                                    a += __shfl_xor_sync(0xffffffffu, a, 2);
                                    a += __shfl_xor_sync(0xffffffffu, a, 8);
                                }
                            }
                            if (small) {
                                /* first-minimum scan, determineLowestLayerKernel.cl:13-18 */
                                const uint32_t S = window_total(a, layer_shift<RT>(P, z0 + j), cur, count, useNb, nb, P.dS, P.nS);
                                if (z0 + j == 0 || S < bestS) {
                                    bestS = S;
                                    winner = z0 + j;
                                }
                            } else if (lane == z0 + j) {
                                mine = a;
                            }
                        }
                    }
                };
                if constexpr (RT > 0) {
#pragma unroll
                    for (int z0 = 0; z0 < RT; z0 += HR_ZCHUNK) do_chunk(z0);
                } else {
#pragma unroll 1
                    for (int z0 = 0; z0 < R; z0 += HR_ZCHUNK) do_chunk(z0);
                }
                HR_STAMP(); /* layers evaluated and reduced */
                if (!small) {
                    sh.warpTot[warp][lane] = mine;
                    __syncthreads();
                    if (big) {
                        if (warp == 0) {
                            uint32_t t = 0;
#pragma unroll
                            for (int w = 0; w < HR_NWARPS; ++w) t += sh.warpTot[w][lane];
                            put_tagged(P.partial + P.bigOff[step] + tile * HR_RMAX + lane, P.epoch, t);
                        }
                        __syncthreads(); /* warpTot is reused below */
                        continue; /* scored below, from the totals of all tiles of the window */
                    }
                    if (ws == HR_TILE) {
                        if (warp == 0) {
                            uint32_t t = 0;
#pragma unroll
                            for (int w = 0; w < HR_NWARPS; ++w) t += sh.warpTot[w][lane];
                            const int wz = finalize_warp(P, R, it, ws, axis, lane, t, g.tx0, g.ty0, cur);
                            if (lane == 0) sh.winner[0] = wz;
                        }
                    } else if (((warp & 1) | ((warp >> 2) & 1)) == 0) { /* window 16: leader warp of each 2x2 warp group */
                        const uint32_t t = sh.warpTot[warp][lane] + sh.warpTot[warp + 1][lane] + sh.warpTot[warp + 4][lane] + sh.warpTot[warp + 5][lane];
                        const int wx0 = g.tx0 + (warp & 2) * 8, wy0 = g.ty0 + (warp >> 3) * 16;
                        int wz = 0;
                        if (wx0 < P.lw && wy0 < P.lh) wz = finalize_warp(P, R, it, ws, axis, lane, t, wx0, wy0, cur);
                        if (lane == 0) sh.winner[(warp >> 3) * 2 + ((warp >> 1) & 1)] = wz;
                    }
                    __syncthreads();
                    winner = sh.winner[ws == HR_TILE ? 0 : (warp >> 3) * 2 + ((warp >> 1) & 1)];
                }
                if (axis) oy += P.cand[winner];
                else ox += P.cand[winner];
                trace_store(P, g, step, winner);
                /* publish this level's windows (neighbours of the next level / blur) */
                if (axis == 1 && g.m0 && g.px == (g.px & ~(ws - 1)) && g.py == (g.py & ~(ws - 1)))
                    put_tagged(Tcur + (g.py >> lgw) * nwx + (g.px >> lgw), P.epoch, (uint32_t)(uint16_t)ox | ((uint32_t)(uint16_t)oy << 16));
                HR_PARK(slot);
            }
            if (big) {
                slot = 0;
                for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas, ++slot) {
                    HR_UNPARK(slot, tile);
                    const int wx = g.tx0 >> lgw, wy = g.ty0 >> lgw;
                    /* the tiles of this window */
                    const int tpw = ws >> 5;
                    const int ax0 = wx * tpw, ay0 = wy * tpw;
                    const int ax1 = hr_min(ax0 + tpw, P.tilesX) - 1, ay1 = hr_min(ay0 + tpw, P.tilesY) - 1;
                    HR_STAMP(); /* tile total published */
                    {
                        /* sum the window's tile totals in a fixed order: warp w takes tiles w, w+16, ... (independent
                         * L2 loads), lane = layer; the 16 warp sums meet in shared memory */
                        const int wT = ax1 - ax0 + 1, nT = wT * (ay1 - ay0 + 1);
                        const unsigned long long *ps = P.partial + P.bigOff[step] + lane;
                        uint32_t t = 0;
#pragma unroll 4
                        for (int i = warp; i < nT; i += HR_NWARPS) t += get_tagged(ps + ((ay0 + i / wT) * P.tilesX + ax0 + i % wT) * HR_RMAX, P.epoch);
                        sh.warpTot[warp][lane] = t;
                    }
                    __syncthreads();
                    if (warp == 0) {
                        const int cur = axis ? oy : ox;
                        uint32_t sad = 0;
#pragma unroll
                        for (int w = 0; w < HR_NWARPS; ++w) sad += sh.warpTot[w][lane];
                        const int wz = finalize_warp(P, R, it, ws, axis, lane, sad, wx << lgw, wy << lgw, cur);
                        if (lane == 0) sh.winner[0] = wz;
                    }
                    __syncthreads();
                    const int winner = sh.winner[0];
                    if (axis) oy += P.cand[winner];
                    else ox += P.cand[winner];
                    trace_store(P, g, step, winner);
                    if (axis == 1 && tid == 0 && g.tx0 == (wx << lgw) && g.ty0 == (wy << lgw))
                        put_tagged(Tcur + wy * nwx + wx, P.epoch, (uint32_t)(uint16_t)ox | ((uint32_t)(uint16_t)oy << 16));
                    HR_PARK(slot);
                    if (MULTI) __syncthreads();
                }
            }
        }
        HR_STAMP(); /* level done */
    }

    HR_STAMP(); /* search done */
    /* raw offsets (offsetArray) */
    {
        int slot = 0;
        for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas, ++slot) {
            HR_UNPARK(slot, tile);
            if (g.m0) {
                const size_t idx = (size_t)g.py * P.lw + g.px;
                P.off[idx] = (int16_t)ox;
                P.off[ln + idx] = (int16_t)oy;
            }
            if (g.m1) {
                const size_t idx = (size_t)(g.py + 1) * P.lw + g.px;
                P.off[idx] = (int16_t)ox;
                P.off[ln + idx] = (int16_t)oy;
            }
        }
    }
#undef HR_UNPARK
#undef HR_PARK

    /* ------------- blur the raw offsets (K4), reading the last level's window table --------------- */
    {
        const int lws = P.first >> (P.iters - 1); /* = 2 */
        const int lgl = 31 - __clz(lws);
        const int lnwx = (P.lw + lws - 1) >> lgl;
        const unsigned long long *Tl = P.T + P.tOff[P.iters - 1];
        int16_t *tX = sh.blur.tX, *tY = sh.blur.tY;
        int *hX = sh.blur.hX, *hY = sh.blur.hY;
        constexpr int NT = HR_THREADS;
        for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas) {
            const int ttx = tile % P.tilesX, tty = tile / P.tilesX;
            const int tx0 = ttx * HR_TILE, ty0 = tty * HR_TILE;
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
                const uint32_t v = get_tagged(Tl + (gy >> lgl) * lnwx + (gx >> lgl), P.epoch);
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
    }
    HR_STAMP(); /* blur done */
#undef HR_STAMP
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

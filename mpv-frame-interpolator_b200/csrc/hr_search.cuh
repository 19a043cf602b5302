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
 *   - Windows >= 64 span CTAs: every tile publishes its totals, every tile of the window sums them in
 *     a fixed order. Neighbour bias (iteration >= 4) and the blur halo read the previous level's
 *     window table of the adjacent tiles. Both hand-offs go through epoch-tagged 64-bit words in L2
 *     (hr_common.cuh): no grid barrier, no atomics, a CTA only waits for the words it reads.
 *   - Every step is compiled for its window class and axis (templates), and for the search radii
 *     the filter uses (5..16) with the layer loop unrolled and the layer shifts as immediates.
 *   - Every evaluation is one 32-bit load from the phase-planar packed frame (hr_pack.cuh) and one
 *     VABSDIFF4.U8.ACC; both biases are added once per window as count*bias (mod 2^32 — exact,
 *     because every point of a window shares offset and neighbours).
 */
#pragma once
#include "hr_common.cuh"
#include <type_traits>

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
__device__ __forceinline__ void neighbour_ptrs(const FlowParams &P, int it, int ws, int x0, int y0, const unsigned long long *(&q)[4]) {
    const int pws = ws << 1;
    const int lgp = 31 - __clz(pws);
    const int pnwx = (P.lw + pws - 1) >> lgp;
    const unsigned long long *Tp = P.T + P.tOff[it - 1];
    const int yd = hr_min(y0 + pws, P.lh - 1) >> lgp, yu = hr_max(y0 - pws, 0) >> lgp;
    const int xr = hr_min(x0 + pws, P.lw - 1) >> lgp, xl = hr_max(x0 - pws, 0) >> lgp;
    const int xc = x0 >> lgp, yc = y0 >> lgp;
    q[0] = Tp + yd * pnwx + xc; /* down */
    q[1] = Tp + yc * pnwx + xr; /* right */
    q[2] = Tp + yc * pnwx + xl; /* left */
    q[3] = Tp + yu * pnwx + xc; /* up */
}
/* first half: issue the four loads (they may come back stale: the producing tile has not got there yet) */
__device__ __forceinline__ void neighbours_issue(const FlowParams &P, int it, int ws, int x0, int y0, unsigned long long (&v)[4]) {
    const unsigned long long *q[4];
    neighbour_ptrs(P, it, ws, x0, y0, q);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = ld_relaxed_u64(q[i]);
}
/* second half: re-load the words whose tag is stale, hand out the payloads */
__device__ __forceinline__ void neighbours_wait(const FlowParams &P, int it, int ws, int x0, int y0, unsigned long long (&v)[4], uint32_t (&nw)[4]) {
    bool stale = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) stale |= (uint32_t)(v[i] >> 32) != P.epoch;
    if (stale) {
        const unsigned long long *q[4];
        neighbour_ptrs(P, it, ws, x0, y0, q);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            while ((uint32_t)(v[i] >> 32) != P.epoch) v[i] = ld_relaxed_u64(q[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) nw[i] = (uint32_t)v[i];
}
__device__ __forceinline__ void load_neighbours(const FlowParams &P, int it, int ws, int x0, int y0, uint32_t (&nw)[4]) {
    unsigned long long v[4];
    neighbours_issue(P, it, ws, x0, y0, v);
    neighbours_wait(P, it, ws, x0, y0, v, nw);
}
/* pick one axis out of the four neighbour words */
__device__ __forceinline__ void neighbour_axis(const uint32_t (&nw)[4], int axis, int (&n)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) n[i] = axis ? ((int)nw[i] >> 16) : (int)(int16_t)nw[i];
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

/* Executed by one full warp: lane z holds the window's SAD for layer z; returns the winner.
 * nw: the window's four neighbour words (loaded on the first axis step of a level, reused on the second). */
template <int AXIS>
__device__ __forceinline__ int finalize_warp(const FlowParams &P, int R, int it, int ws, int lane, uint32_t sad, int x0, int y0, int cur, bool loadNb,
                                             uint32_t (&nw)[4]) {
    const bool useNb = it >= HR_FIRST_NEIGHBOR_ITERATION;
    const uint32_t count = (uint32_t)(hr_min(x0 + ws, P.lw) - x0) * (uint32_t)(hr_min(y0 + ws, P.lh) - y0);
    int n[4] = {0, 0, 0, 0};
    if (useNb) {
        if (loadNb) load_neighbours(P, it, ws, x0, y0, nw);
        neighbour_axis(nw, AXIS, n);
    }
    const uint32_t S = (lane < R) ? window_total(sad, P.cand[lane < R ? lane : 0], cur, count, useNb, n, P.dS, P.nS) : 0xffffffffu;
    const uint32_t mn = __reduce_min_sync(0xffffffffu, S);
    const unsigned ballot = __ballot_sync(0xffffffffu, S == mn && lane < R);
    return __ffs(ballot) - 1;
}

/* Everything a thread carries: its place in the tile and the search state of its two points. */
struct Thr {
    int tile, tx0, ty0;   /* owned tile and its lattice origin                                        */
    int px, py;           /* lattice position of the upper point; the lower one is (px, py+1)           */
    uint32_t m0, m1;      /* all-ones if the upper / lower point is inside the lattice                  */
    int cxs, cy0s, cy1s;  /* full-resolution coordinates of the (clamped) points: lattice << s          */
    int ox, oy;           /* the offset pair both points share                                         */
    uint32_t v2a, v2b;    /* frame2 words of the two points (0 when outside)                            */
    uint32_t nw[4];       /* neighbour words of the current level                                      */
};
__device__ __forceinline__ void thr_place(const FlowParams &P, Thr &t, int tile, int warp, int lane) {
    t.tile = tile;
    t.tx0 = (tile % P.tilesX) * HR_TILE;
    t.ty0 = (tile / P.tilesX) * HR_TILE;
    t.px = t.tx0 + (warp & 3) * 8 + (lane & 7);
    t.py = t.ty0 + (warp >> 2) * 8 + (lane >> 3) * 2;
    t.m0 = (t.px < P.lw && t.py < P.lh) ? 0xffffffffu : 0u;
    t.m1 = (t.px < P.lw && t.py + 1 < P.lh) ? 0xffffffffu : 0u;
    t.cxs = hr_min(t.px, P.lw - 1) << P.s;
    t.cy0s = hr_min(t.py, P.lh - 1) << P.s;
    t.cy1s = hr_min(t.py + 1, P.lh - 1) << P.s;
}

__device__ __forceinline__ void trace_store(const FlowParams &P, const Thr &t, int step, int winner) {
    if (P.trace) {
        if (t.m0) P.trace[((size_t)step * P.lh + t.py) * P.lw + t.px] = (uint8_t)winner;
        if (t.m1) P.trace[((size_t)step * P.lh + t.py + 1) * P.lw + t.px] = (uint8_t)winner;
    }
}

/* One search step of one tile. WS: 2, 4, 8 (warp-local windows), 16, 32 (tile-local) or 64 (= any
 * window larger than a tile; the real size is `ws`). AXIS 0: the layers move x, 1: y.
 * Packed word of full-resolution sample (x,y): plane ((y&m)<<s | (x&m)), row y>>s, column x>>s.
 * For WS = 64 the step only publishes the tile's totals; big_finish() scores them. */
template <int RT, int WS, int AXIS, bool BANDS = false>
__device__ __forceinline__ void search_step(const FlowParams &P, SearchShared &sh, Thr &t, int it, int ws, int lane, int warp, long long *fine = nullptr) {
    /* DBG instantiations only (fine == nullptr folds away elsewhere): clock stamps of thread 0 inside the step */
#define FST(k) if (fine) { asm volatile("" ::: "memory"); fine[k] = clock64(); asm volatile("" ::: "memory"); }
    FST(0)
    constexpr bool small = WS <= 8;
    const int R = RT > 0 ? RT : P.R;
    const int s = P.s, m = (1 << s) - 1;
    const int step = it * 2 + AXIS;
    const bool useNb = it >= HR_FIRST_NEIGHBOR_ITERATION;
    const int cur = AXIS ? t.oy : t.ox;

    /* the thread's own window (warp-local levels): population; its neighbours are fetched further down,
     * after the sample loads have been issued */
    uint32_t count = 0;
    int nb[4] = {0, 0, 0, 0};
    const int x0 = t.px & ~(WS - 1), y0 = t.py & ~(WS - 1);
    const bool ownWindow = small && x0 < P.lw && y0 < P.lh;
    if (ownWindow) count = (uint32_t)(hr_min(x0 + WS, P.lw) - x0) * (uint32_t)(hr_min(y0 + WS, P.lh) - y0);

    /* sample addressing: the fixed part (the axis that does not move) and the moving coordinate before
     * the layer shift, for the upper (a) and lower (b) point */
    int fa, fb, ma, mb;
    if (AXIS == 0) {
        const int ya = search_mirror(t.cy0s + t.oy, P.H), yb = search_mirror(t.cy1s + t.oy, P.H);
        fa = ((ya & m) << s) * P.planeSize + (ya >> s) * P.planePitch;
        fb = ((yb & m) << s) * P.planeSize + (yb >> s) * P.planePitch;
        ma = mb = t.cxs + t.ox;
    } else {
        const int x = search_mirror(t.cxs + t.ox, P.W);
        fa = fb = (x & m) * P.planeSize + (x >> s);
        ma = t.cy0s + t.oy;
        mb = t.cy1s + t.oy;
    }
    const int D = AXIS ? P.H : P.W;
    const int mulA = AXIS ? (P.planeSize << s) : P.planeSize;
    /* warp-uniform: no layer of any lane leaves the frame on the searched axis -> no mirror */
    const bool interior = __all_sync(0xffffffffu, ma + layer_shift<RT>(P, 0) >= 0 && mb + layer_shift<RT>(P, R - 1) < D);

    uint32_t bestS = 0xffffffffu, mine = 0u;
    int winner = 0;
    auto issue = [&](int z0, uint32_t (&va)[HR_ZCHUNK], uint32_t (&vb)[HR_ZCHUNK]) {
#pragma unroll
        for (int j = 0; j < HR_ZCHUNK; ++j) {
            if (z0 + j < R) {
                const int c = layer_shift<RT>(P, z0 + j);
                int pa = ma + c, pb = mb + c;
                if (!interior) {
                    pa = search_mirror(pa, D);
                    if (AXIS) pb = search_mirror(pb, D);
                }
                if (AXIS == 0) {
                    const int xi = (pa & m) * mulA + (pa >> s);
                    va[j] = __ldg(P.p1 + (fa + xi));
                    vb[j] = __ldg(P.p1 + (fb + xi));
                } else {
                    va[j] = __ldg(P.p1 + (fa + (pa & m) * mulA + (pa >> s) * P.planePitch));
                    vb[j] = __ldg(P.p1 + (fb + (pb & m) * mulA + (pb >> s) * P.planePitch));
                }
            }
        }
    };
    auto consume = [&](int z0, const uint32_t (&va)[HR_ZCHUNK], const uint32_t (&vb)[HR_ZCHUNK]) {
#pragma unroll
        for (int j = 0; j < HR_ZCHUNK; ++j) {
            if (z0 + j < R) {
                uint32_t a = sad4_acc(vb[j] & t.m1, t.v2b, sad4_acc(va[j] & t.m0, t.v2a, 0u));
                if (WS >= 8) {
                    a = __reduce_add_sync(0xffffffffu, a);
                } else {
                    a += __shfl_xor_sync(0xffffffffu, a, 1);
                    if (WS == 4) {
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
    /* the neighbour words of the level (first axis step only; the second reuses them). They come from
     * the adjacent tiles and may still be in flight: poll them only after this step's sample loads have
     * been issued. Windows 16 / 32: the warp that will score the window fetches them. */
    bool scorer = false;
    int sx0 = 0, sy0 = 0;
    if (WS == HR_TILE) {
        scorer = warp == 0;
        sx0 = t.tx0;
        sy0 = t.ty0;
    } else if (WS == 16) {
        sx0 = t.tx0 + (warp & 2) * 8;
        sy0 = t.ty0 + (warp >> 3) * 16;
        scorer = ((warp & 1) | ((warp >> 2) & 1)) == 0 && sx0 < P.lw && sy0 < P.lh;
    }
    const bool wantNb = useNb && AXIS == 0 && WS <= HR_TILE && (small ? ownWindow : scorer);
    const int nx0 = small ? x0 : sx0, ny0 = small ? y0 : sy0;
    unsigned long long nbv[4] = {0ull, 0ull, 0ull, 0ull};
    auto prefetch_neighbours = [&]() {
        if (wantNb) neighbours_issue(P, it, WS, nx0, ny0, nbv);
    };
    auto fetch_neighbours = [&]() {
        if (wantNb) neighbours_wait(P, it, WS, nx0, ny0, nbv, t.nw);
        if (small && useNb) neighbour_axis(t.nw, AXIS, nb);
    };
    if constexpr (RT > 0) {
        uint32_t va[HR_ZCHUNK], vb[HR_ZCHUNK];
        prefetch_neighbours();
        issue(0, va, vb);
        FST(1)
        fetch_neighbours();
        FST(2)
        if (fine) { /* make the stamp wait for the samples */
            uint32_t x = 0;
            for (int j = 0; j < HR_ZCHUNK; ++j)
                if (j < R) x ^= va[j] ^ vb[j];
            if (x == 0x12345679u) fine[5] = 0;
            FST(3)
        }
        consume(0, va, vb);
        FST(4)
#pragma unroll
        for (int z0 = HR_ZCHUNK; z0 < RT; z0 += HR_ZCHUNK) {
            uint32_t wa[HR_ZCHUNK], wb[HR_ZCHUNK];
            issue(z0, wa, wb);
            consume(z0, wa, wb);
        }
    } else {
        uint32_t va[HR_ZCHUNK], vb[HR_ZCHUNK];
        prefetch_neighbours();
        issue(0, va, vb);
        fetch_neighbours();
        consume(0, va, vb);
#pragma unroll 1
        for (int z0 = HR_ZCHUNK; z0 < R; z0 += HR_ZCHUNK) {
            issue(z0, va, vb);
            consume(z0, va, vb);
        }
    }

    if (!small) {
        sh.warpTot[warp][lane] = mine;
        __syncthreads();
        if (WS > HR_TILE) {
            if (warp == 0) {
                uint32_t tt = 0;
#pragma unroll
                for (int w = 0; w < HR_NWARPS; ++w) tt += sh.warpTot[w][lane];
                unsigned long long *slot = P.partial + P.bigOff[step] + t.tile * HR_RMAX + lane;
                if (BANDS) put_tagged_all(P.band, slot, P.epoch, tt); /* the window's tiles live on several GPUs */
                else put_tagged(slot, P.epoch, tt);
            }
            __syncthreads(); /* warpTot is reused by big_finish */
            return;
        }
        if (WS == HR_TILE) {
            if (warp == 0) {
                uint32_t tt = 0;
#pragma unroll
                for (int w = 0; w < HR_NWARPS; ++w) tt += sh.warpTot[w][lane];
                const int wz = finalize_warp<AXIS>(P, R, it, WS, lane, tt, t.tx0, t.ty0, cur, false, t.nw);
                if (lane == 0) sh.winner[0] = wz;
            }
        } else if (((warp & 1) | ((warp >> 2) & 1)) == 0) { /* window 16: leader warp of each 2x2 warp group */
            const uint32_t tt = sh.warpTot[warp][lane] + sh.warpTot[warp + 1][lane] + sh.warpTot[warp + 4][lane] + sh.warpTot[warp + 5][lane];
            int wz = 0;
            if (scorer) wz = finalize_warp<AXIS>(P, R, it, WS, lane, tt, sx0, sy0, cur, false, t.nw);
            if (lane == 0) sh.winner[(warp >> 3) * 2 + ((warp >> 1) & 1)] = wz;
        }
        __syncthreads();
        winner = sh.winner[WS == HR_TILE ? 0 : (warp >> 3) * 2 + ((warp >> 1) & 1)];
    }
    FST(5)
    if (AXIS) t.oy += P.cand[winner];
    else t.ox += P.cand[winner];
    trace_store(P, t, step, winner);
    /* publish this level's windows (neighbours of the next level / blur) */
    if (AXIS == 1 && t.m0 && (t.px & (WS - 1)) == 0 && (t.py & (WS - 1)) == 0) {
        const int lgw = 31 - __clz(WS), nwx = (P.lw + WS - 1) >> lgw;
        unsigned long long *slot = P.T + P.tOff[it] + (t.py >> lgw) * nwx + (t.px >> lgw);
        const uint32_t word = (uint32_t)(uint16_t)t.ox | ((uint32_t)(uint16_t)t.oy << 16);
        put_tagged(slot, P.epoch, word);
        if (BANDS) {
            /* the tiles across a band's edge read this level's windows (neighbour bias of the next level, blur halo) */
            const int tileRow = t.ty0 / HR_TILE;
            if (tileRow == P.band.tileRow0 && P.band.up >= 0) st_relaxed_sys_u64(at_peer(slot, P.band.peerDelta[P.band.up]), ((unsigned long long)P.epoch << 32) | word);
            if (tileRow == P.band.tileRow1 - 1 && P.band.down >= 0) st_relaxed_sys_u64(at_peer(slot, P.band.peerDelta[P.band.down]), ((unsigned long long)P.epoch << 32) | word);
        }
    }
}

#undef FST

/* Second half of a step whose windows span several tiles: sum the totals of the window's tiles in a
 * fixed order (warp w takes tiles w, w+16, ... — independent L2 loads, lane = layer) and score. */
template <int RT, int AXIS>
__device__ __forceinline__ void big_finish(const FlowParams &P, SearchShared &sh, Thr &t, int it, int ws, int lane, int warp, bool loadNb) {
    const int R = RT > 0 ? RT : P.R;
    const int step = it * 2 + AXIS;
    const int lgw = 31 - __clz(ws), nwx = (P.lw + ws - 1) >> lgw;
    const int wx = t.tx0 >> lgw, wy = t.ty0 >> lgw;
    const int lgt = lgw - 5, tpw = 1 << lgt;                     /* tiles per window side */
    const int ax0 = wx << lgt, ay0 = wy << lgt;
    const unsigned long long *ps = P.partial + P.bigOff[step] + lane;
    uint32_t tt = 0;
    for (int i0 = warp; i0 < tpw * tpw; i0 += 4 * HR_NWARPS) {
        /* up to four tile totals per pass: issue the loads together, then re-load the ones whose tag is stale */
        const unsigned long long *q[4];
        unsigned long long v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + k * HR_NWARPS;
            const int tx = ax0 + (i & (tpw - 1)), ty = ay0 + (i >> lgt);
            q[k] = (i < tpw * tpw && tx < P.tilesX && ty < P.tilesY) ? ps + (ty * P.tilesX + tx) * HR_RMAX : nullptr;
            v[k] = q[k] ? ld_relaxed_u64(q[k]) : 0ull;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (q[k]) {
                while ((uint32_t)(v[k] >> 32) != P.epoch) v[k] = ld_relaxed_u64(q[k]);
                tt += (uint32_t)v[k];
            }
        }
    }
    sh.warpTot[warp][lane] = tt;
    __syncthreads();
    if (warp == 0) {
        const int cur = AXIS ? t.oy : t.ox;
        uint32_t sad = 0;
#pragma unroll
        for (int w = 0; w < HR_NWARPS; ++w) sad += sh.warpTot[w][lane];
        const int wz = finalize_warp<AXIS>(P, R, it, ws, lane, sad, wx << lgw, wy << lgw, cur, loadNb, t.nw);
        if (lane == 0) sh.winner[0] = wz;
    }
    __syncthreads();
    const int winner = sh.winner[0];
    if (AXIS) t.oy += P.cand[winner];
    else t.ox += P.cand[winner];
    trace_store(P, t, step, winner);
    if (AXIS == 1 && threadIdx.x == 0 && t.tx0 == (wx << lgw) && t.ty0 == (wy << lgw))
        put_tagged(P.T + P.tOff[it] + wy * nwx + wx, P.epoch, (uint32_t)(uint16_t)t.ox | ((uint32_t)(uint16_t)t.oy << 16));
}

/* HR_SEARCH_MAXNREG: registers per thread the search may use. One CTA of 512 threads per SM either way; a lower
 * cap leaves room for the pack and warp CTAs of the neighbouring pairs on the same SM (pipelined mode). */
#ifndef HR_SEARCH_MAXNREG
#define HR_SEARCH_MAXNREG 88 /* measured (tools/diag_pipeline.py): no spills at R = 5, 24 bytes at R = 16, same serial time as 117 */
#endif
#define HR_SEARCH_BOUNDS __maxnreg__(HR_SEARCH_MAXNREG)
template <int RT, bool MULTI, bool DBG, bool BANDS = false>
__device__ __forceinline__ void flow_search_body(const FlowParams &P) {
    static_assert(!(BANDS && MULTI), "a band's tiles always fit the GPU");
    __shared__ SearchShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned nCtas = gridDim.x;
    const size_t ln = (size_t)P.lw * P.lh;
    int stampIdx = 0;
#define HR_STAMP() \
    if (DBG && P.timeline && tid == 0 && stampIdx < HR_TIMELINE_SLOTS) P.timeline[blockIdx.x * HR_TIMELINE_SLOTS + stampIdx++] = clock64();
    HR_STAMP();
    if (DBG && P.timeline && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        P.timeline[blockIdx.x * HR_TIMELINE_SLOTS + HR_TIMELINE_SLOTS - 2] = (long long)gt;
    }

    Thr t;
    t.ox = t.oy = 0;
    t.nw[0] = t.nw[1] = t.nw[2] = t.nw[3] = 0u;
    /* frame2 sample of a lattice point as the packed word Y | U << 8 | V << 16 (calcDeltaSumsKernel.cl:96-98:
     * chroma at row y >> 1, byte column x & ~1 (+1)); P010: the top 8 bits of every sample */
    auto frame2_word = [&](int x, int y) -> uint32_t {
        const size_t iy = (size_t)y * P.W + x, iuv = (size_t)(y >> 1) * P.W + (x & ~1);
        if (P.bps == 1) {
            const uint8_t *fy = (const uint8_t *)P.f2y, *fuv = (const uint8_t *)P.f2uv;
            return (uint32_t)__ldg(fy + iy) | ((uint32_t)__ldg(fuv + iuv) << 8) | ((uint32_t)__ldg(fuv + iuv + 1) << 16);
        }
        const uint16_t *fy = (const uint16_t *)P.f2y, *fuv = (const uint16_t *)P.f2uv;
        return ((uint32_t)__ldg(fy + iy) >> 8) | ((uint32_t)__ldg(fuv + iuv) & 0xff00u) | (((uint32_t)__ldg(fuv + iuv + 1) & 0xff00u) << 8);
    };
    auto load_frame2 = [&]() {
        t.v2a = frame2_word(t.cxs, t.cy0s) & t.m0;
        t.v2b = frame2_word(t.cxs, t.cy1s) & t.m1;
    };
    const int firstTile = BANDS ? P.band.tile0 + (int)blockIdx.x : (int)blockIdx.x;
    if (!MULTI) {
        thr_place(P, t, firstTile, warp, lane);
        load_frame2();
    } else {
        int slot = 0;
        for (int tile = blockIdx.x; tile < P.numTiles; tile += nCtas, ++slot) {
            thr_place(P, t, tile, warp, lane);
            load_frame2();
            sh.parked[slot][tid] = make_int4(0, 0, (int)t.v2a, (int)t.v2b);
        }
    }
    if (BANDS && P.band.world > 1) {
        /* Every GPU of the group announces that it has entered this pair's search — by stream order that also says its
         * previous search and the warps that read the previous flow are done, so its tables and its flow buffer may be
         * written — and nobody stores into a peer before that peer has said so. All GPUs are then inside the same launch,
         * which is what the polling below relies on. */
        if (blockIdx.x == 0 && tid < P.band.world && tid != P.band.rank)
            st_release_sys_u64(at_peer(P.band.ready + P.band.rank, P.band.peerDelta[tid]), P.epoch);
        if (warp == 0 && lane < P.band.world && lane != P.band.rank)
            while ((uint32_t)ld_acquire_sys_u64(P.band.ready + lane) != P.epoch) {}
        __syncthreads();
    }
    /* a CTA that owns several tiles (MULTI) parks the per-thread state of each in shared memory */
    auto for_tiles = [&](auto body) {
        int slot = 0;
        for (int tile = firstTile; tile < (BANDS ? firstTile + 1 : P.numTiles); tile += nCtas, ++slot) {
            if (MULTI) {
                thr_place(P, t, tile, warp, lane);
                const int4 st = sh.parked[slot][tid];
                t.ox = st.x;
                t.oy = st.y;
                t.v2a = (uint32_t)st.z;
                t.v2b = (uint32_t)st.w;
            }
            body();
            if (MULTI) {
                sh.parked[slot][tid] = make_int4(t.ox, t.oy, (int)t.v2a, (int)t.v2b);
                __syncthreads();
            }
        }
    };
    auto level = [&](auto wsTag, int it, int ws) {
        constexpr int WS = decltype(wsTag)::value;
        if constexpr (WS > HR_TILE) {
            for_tiles([&] { search_step<RT, WS, 0, BANDS>(P, sh, t, it, ws, lane, warp); });
            HR_STAMP();
            for_tiles([&] { big_finish<RT, 0>(P, sh, t, it, ws, lane, warp, true); });
            HR_STAMP();
            for_tiles([&] { search_step<RT, WS, 1, BANDS>(P, sh, t, it, ws, lane, warp); });
            HR_STAMP();
            for_tiles([&] { big_finish<RT, 1>(P, sh, t, it, ws, lane, warp, MULTI); });
            HR_STAMP();
        } else {
            for_tiles([&] {
                /* DBG: six stamps inside each step of the three warp-local levels, slots 40.. (tools/diag_fine.py) */
                const int fineLevel = it - (P.iters - 3);
                long long *fine = (DBG && P.timeline && tid == 0 && WS <= 8 && fineLevel >= 0 && fineLevel < 3)
                                      ? P.timeline + blockIdx.x * HR_TIMELINE_SLOTS + 40 + fineLevel * 12 : nullptr;
                search_step<RT, WS, 0, BANDS>(P, sh, t, it, ws, lane, warp, fine);
                HR_STAMP();
                search_step<RT, WS, 1, BANDS>(P, sh, t, it, ws, lane, warp, fine ? fine + 6 : nullptr);
                HR_STAMP();
            });
        }
    };
    for (int it = 0; it < P.iters; ++it) {
        const int ws = P.first >> it;
        if (ws > HR_TILE) level(std::integral_constant<int, 64>(), it, ws);
        else if (ws == 32) level(std::integral_constant<int, 32>(), it, ws);
        else if (ws == 16) level(std::integral_constant<int, 16>(), it, ws);
        else if (ws == 8) level(std::integral_constant<int, 8>(), it, ws);
        else if (ws == 4) level(std::integral_constant<int, 4>(), it, ws);
        else level(std::integral_constant<int, 2>(), it, ws);
    }
    HR_STAMP(); /* search done */

    /* raw offsets (offsetArray) */
    const int copies = BANDS ? P.band.world : 1; /* bands: every GPU gets the whole flow */
    for_tiles([&] {
        for (int g = 0; g < copies; ++g) {
            int16_t *off = BANDS ? at_peer(P.off, P.band.peerDelta[g]) : P.off;
            if (t.m0) {
                const size_t idx = (size_t)t.py * P.lw + t.px;
                off[idx] = (int16_t)t.ox;
                off[ln + idx] = (int16_t)t.oy;
            }
            if (t.m1) {
                const size_t idx = (size_t)(t.py + 1) * P.lw + t.px;
                off[idx] = (int16_t)t.ox;
                off[ln + idx] = (int16_t)t.oy;
            }
        }
    });

    /* ------------- blur the raw offsets (K4), reading the last level's window table --------------- */
    {
        const int lws = P.first >> (P.iters - 1); /* = 2 */
        const int lgl = 31 - __clz(lws);
        const int lnwx = (P.lw + lws - 1) >> lgl;
        const unsigned long long *Tl = P.T + P.tOff[P.iters - 1];
        int16_t *tX = sh.blur.tX, *tY = sh.blur.tY;
        int *hX = sh.blur.hX, *hY = sh.blur.hY;
        constexpr int NT = HR_THREADS;
        for (int tile = firstTile; tile < (BANDS ? firstTile + 1 : P.numTiles); tile += nCtas) {
            const int ttx = tile % P.tilesX, tty = tile / P.tilesX;
            const int tx0 = ttx * HR_TILE, ty0 = tty * HR_TILE;
            {
                constexpr int NU = (40 * 40 + NT - 1) / NT;
                const unsigned long long *q[NU];
                unsigned long long v[NU];
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int i = tid + u * NT;
                    q[u] = nullptr;
                    v[u] = 0ull;
                    if (i < 40 * 40) {
                        const int r = i / 40, c = i - r * 40;
                        int gy = ty0 - 4 + r, gx = tx0 - 4 + c;
                        /* blurFlowKernel.cl:5-12 mirror, clamped for lattices smaller than the halo */
                        if (gy >= P.lh) gy = 2 * P.lh - gy - 1; else if (gy < 0) gy = -gy - 1;
                        if (gx >= P.lw) gx = 2 * P.lw - gx - 1; else if (gx < 0) gx = -gx - 1;
                        gy = hr_min(hr_max(gy, 0), P.lh - 1);
                        gx = hr_min(hr_max(gx, 0), P.lw - 1);
                        q[u] = Tl + (gy >> lgl) * lnwx + (gx >> lgl);
                        v[u] = ld_relaxed_u64(q[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int i = tid + u * NT;
                    if (q[u]) {
                        while ((uint32_t)(v[u] >> 32) != P.epoch) v[u] = ld_relaxed_u64(q[u]); /* a neighbour tile is still searching */
                        tX[i] = (int16_t)(v[u] & 0xffffu);
                        tY[i] = (int16_t)((v[u] >> 16) & 0xffffu);
                    }
                }
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
                    const int bx = sx / 64, by = sy / 64; /* C division truncates toward zero */
                    for (int g = 0; g < copies; ++g) {
                        int16_t *blur = BANDS ? at_peer(P.blur, P.band.peerDelta[g]) : P.blur;
                        uint32_t *blurXY = BANDS ? at_peer(P.blurXY, P.band.peerDelta[g]) : P.blurXY;
                        blur[idx] = (int16_t)bx;
                        blur[ln + idx] = (int16_t)by;
                        blurXY[idx] = (uint32_t)(uint16_t)bx | ((uint32_t)(uint16_t)by << 16);
                    }
                }
            }
            __syncthreads();
        }
    }
    if (BANDS && P.band.world > 1) {
        /* The launch may only end when the whole flow is here: the last CTA of every GPU to have stored its results tells
         * the peers (after a system-wide fence), and one CTA of every GPU waits for all of them. */
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            const unsigned prev = atomicAdd(P.band.exitCount, 1u);
            if (prev == gridDim.x - 1) {
                *P.band.exitCount = 0u; /* for the next launch (stream order) */
                __threadfence_system();
                for (int g = 0; g < P.band.world; ++g)
                    if (g != P.band.rank) st_release_sys_u64(at_peer(P.band.done + P.band.rank, P.band.peerDelta[g]), P.epoch);
            }
        }
        if (blockIdx.x == 0 && warp == 0 && lane < P.band.world && lane != P.band.rank)
            while ((uint32_t)ld_acquire_sys_u64(P.band.done + lane) != P.epoch) {}
    }
    HR_STAMP(); /* blur done */
    if (DBG && P.timeline && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        P.timeline[blockIdx.x * HR_TIMELINE_SLOTS + HR_TIMELINE_SLOTS - 1] = (long long)gt;
    }
#undef HR_STAMP
}
/* radius known at compile time (5..16, what the filter uses): register cap; the generic instantiations (any
 * radius 2..32, several tiles per CTA) keep all the registers one CTA per SM can have */
template <int RT, bool MULTI, bool DBG>
__global__ void HR_SEARCH_BOUNDS flow_search_kernel(const FlowParams P) {
    static_assert(RT > 0 && !MULTI, "capped instantiation");
    flow_search_body<RT, MULTI, DBG>(P);
}
template <bool MULTI, bool DBG>
__global__ void __launch_bounds__(HR_THREADS, 1) flow_search_generic_kernel(const FlowParams P) {
    flow_search_body<0, MULTI, DBG>(P);
}
/* one band of a frame that is split over several GPUs (BandLink, hr_common.cuh): the default radius unrolled (RT = 5),
 * any other radius generic (RT = 0) */
template <int RT>
__global__ void __launch_bounds__(HR_THREADS, 1) flow_search_band_kernel(const FlowParams P) {
    flow_search_body<RT, false, false, true>(P);
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

/*
 * hr_search2.cuh — second generation of the block-offset search (K1+K2+K3, 16 steps) + flow blur (K4): the same
 * decomposition, hand-off words and results as hr_search.cuh, organised around what ncu showed the first one to be
 * bound by (profiles/r02_search_v1_source_counters.txt): not the SADs (2.8 % of the instructions) but everything
 * around them — twelve inlined step variants of 8-10 KB each that run once per launch (127 KB of straight-line code:
 * "no instruction" is the largest stall reason), per-thread constants rebuilt in every step under the register cap,
 * mirror arithmetic on every candidate, redundant scoring and three block barriers per cross-tile step.
 *
 * Reference semantics: video/filter/HopperRender/Kernels/calcDeltaSumsKernel.cl:34-189,
 * determineLowestLayerKernel.cl:2-22, adjustOffsetArrayKernel.cl:2-18, blurFlowKernel.cl:15-89,
 * driver loop opticalFlowCalc.c:126-203.
 *
 * What is different here:
 *   - ONE step body, looped 2 x iterations times with the window size and the axis as run-time values: the code of
 *     a launch is a few KB and stays in the instruction caches from the second step on.
 *   - A thread's place (lattice point, validity, frame-2 words, layer shift of "its" layer) is computed once and kept.
 *   - Sample addressing without mirror arithmetic whenever no layer of any lane of the warp leaves the frame
 *     (warp-uniform test); the lower point is "upper point + one packed row". Only warps at the frame border take
 *     the reflected path. Points outside the lattice issue no loads; warps wholly outside skip the evaluation.
 *   - Windows of 8 and more: lane z scores layer z (one window_total per lane instead of R per lane), the winner
 *     is a REDUX.MIN + ballot. Windows 16 / 32 / cross-tile: the warps' totals meet in shared memory (double
 *     buffered by step parity), every warp of the window scores them itself: ONE block barrier per tile-local
 *     step, TWO per cross-tile step (three / three before), no winner broadcast.
 *   - The neighbour-window words of a level are requested at the top of its X step and looked at only when the
 *     scoring needs them, after the samples, the SADs and the reductions.
 *
 * Used for the radii the filter's auto-adjust visits (5..16), one tile per CTA, no bands, no taps that need the
 * in-kernel timeline; everything else runs hr_search.cuh. Tables, tile totals and outputs are the same words at
 * the same places, so the two generations are interchangeable launch by launch (tests/test_gpu_search2.py).
 */
#pragma once
#include "hr_search.cuh"

struct Search2Shared {
    uint32_t warpTot[2][HR_NWARPS][HR_RMAX]; /* per-warp block totals of a step, by step parity                 */
    uint32_t bigTot[2][HR_NWARPS][HR_RMAX];  /* per-warp sums of the other tiles' totals (cross-tile steps)      */
    struct {                                 /* blur phase                                                       */
        int16_t tX[40 * 40], tY[40 * 40];
        int hX[40 * 32], hY[40 * 32];
    } blur;
};

__device__ __forceinline__ int signed_square(int rel) { return rel * (rel < 0 ? -rel : rel); }

/* ---- staged variant (STAGED): the tile's neighbourhood of every phase plane in shared memory, fetched by TMA ----
 * At resolution scalar 2 (1080p, 720p) the packed frame has 16 phase planes. A tile's 32x32 points, displaced by up
 * to 9 lattice cells (R = 5: 8 levels x 4 samples + the layer shift = 36 samples), read a box of each plane around the
 * tile: 16 boxes of 50 rows x 60 words = 188 KB, fetched once per launch by 16 cp.async.bulk.tensor.3d loads (one
 * elected thread, one mbarrier), out-of-frame parts zero-filled. The box starts 12 cells left of the tile, not 9: the
 * innermost TMA coordinate has to be a multiple of 16 bytes (measured, tools/tma_probe: any other start is an illegal
 * instruction); and rows are 60 words so that the four rows a warp reads (2 * 60 mod 32 = 24) fall on disjoint banks.
 * A step whose samples all lie inside the frame (no reflection) and inside the box (warp-uniform test), once the
 * fetch has landed (mbarrier probed, never waited for: until then the steps read global memory), costs two LDS per
 * layer and thread at a fixed lane offset — no global load, no L1 tag look-up. */
#define HR_ST_S 2
#define HR_ST_X0 12 /* cells staged left of the tile  */
#define HR_ST_X1 16 /* ... right of it                 */
#define HR_ST_Y 9   /* ... above and below             */
#define HR_ST_PW 60
#define HR_ST_PH 50
#define HR_ST_PLANE 3008 /* words per staged plane: 50 * 60 = 3000 rounded up to a multiple of 128 bytes */
#define HR_ST_PLANES 16
#define HR_ST_BYTES (HR_ST_PLANES * HR_ST_PLANE * 4)
#define HR_ST_MAX_RADIUS 8 /* radii whose layer shifts (up to 16 samples) usually stay inside the box */

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
/* one box of a 3-D tensor (words, rows, planes) into shared memory; completion is counted on the mbarrier */
__device__ __forceinline__ void tma_load_3d(void *dst, const void *tmap, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
struct alignas(64) HrTensorMap { unsigned char bytes[128]; }; /* a CUtensorMap, as the driver encoded it */

/* DBG: thread 0 of every CTA stamps the SM clock into P.timeline (hr_set_timeline / hr_get_timeline): slot 0 launch
 * entered, 1 + 4 * step + {0: step entered, 1: sample loads issued, 2: the window's totals are in the lane,
 * 3: winner known}, 100 search done, 101 blur done, 126 / 127 globaltimer at entry / exit; lane 0 of warp w keeps
 * 8 * step + phase in slot 104 + w (with the timeline in mapped host memory, HR_TIMELINE_HOST=1, the host can read
 * where every warp of a launch that does not return is standing: hr_debug_peek_timeline). */
template <int RT, bool DBG, bool STAGED>
__device__ __forceinline__ void flow_search2_body(const FlowParams &P, const HrTensorMap *tmap) {
    static_assert(RT >= 2 && RT <= HR_RMAX, "search radius");
    __shared__ Search2Shared sh;
    extern __shared__ __align__(128) uint32_t stage[]; /* STAGED: [16 planes][50 rows][60 words] */
    __shared__ __align__(8) uint64_t stageBar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#define S2_STAMP(slot) \
    if (DBG && P.timeline && tid == 0) { asm volatile("" ::: "memory"); P.timeline[blockIdx.x * HR_TIMELINE_SLOTS + (slot)] = clock64(); asm volatile("" ::: "memory"); }
#define S2_AT(phase) \
    if (DBG && P.timeline && lane == 0) { *(volatile long long *)(P.timeline + blockIdx.x * HR_TIMELINE_SLOTS + 104 + warp) = 8 * step + (phase); }
    S2_STAMP(0)
    if (DBG && P.timeline && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        P.timeline[blockIdx.x * HR_TIMELINE_SLOTS + HR_TIMELINE_SLOTS - 2] = (long long)gt;
    }
    const int s = P.s, m = (1 << s) - 1, one = 1 << s;
    const int tile = blockIdx.x;
    const int tx0 = (tile % P.tilesX) * HR_TILE, ty0 = (tile / P.tilesX) * HR_TILE;
    if (STAGED) {
        if (tid == 0) mbar_init(&stageBar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&stageBar, HR_ST_PLANES * HR_ST_PH * HR_ST_PW * 4);
#pragma unroll 1
            for (int pl = 0; pl < HR_ST_PLANES; ++pl) tma_load_3d(stage + pl * HR_ST_PLANE, tmap, &stageBar, tx0 - HR_ST_X0, ty0 - HR_ST_Y, pl);
        }
    }
    /* the thread's two lattice points: (px, py) and (px, py + 1) */
    const int px = tx0 + (warp & 3) * 8 + (lane & 7);
    const int py = ty0 + (warp >> 2) * 8 + (lane >> 3) * 2;
    const bool vA = px < P.lw && py < P.lh, vB = px < P.lw && py + 1 < P.lh;
    /* Does the warp's 8x8 block touch the lattice at all? Plain arithmetic on the block origin, NOT a vote over vA: the
     * compiler re-evaluates a loop-invariant vote inside the step loop without a warp barrier in front of it, and lanes
     * that arrive there apart (after polling for different neighbour words) then each see their own subset — the lanes
     * outside the lattice decide "no point is valid", take the other branch, and the warp waits for itself forever
     * (observed: R = 5, the warps that straddle the last lattice row; tools/diag_search2_where.py). */
    const bool warpValid = tx0 + (warp & 3) * 8 < P.lw && ty0 + (warp >> 2) * 8 < P.lh;
    /* clamped full-resolution coordinates (the reflected path; the same numbers as hr_search.cuh) */
    const int cxs = hr_min(px, P.lw - 1) << s, cy0s = hr_min(py, P.lh - 1) << s, cy1s = hr_min(py + 1, P.lh - 1) << s;
    constexpr int CMIN = -(RT / 2) * (RT / 2), CMAX = (RT - 1 - RT / 2) * (RT - 1 - RT / 2);
    const int candLane = signed_square(lane - RT / 2); /* layer shift of layer `lane` (calcDeltaSumsKernel.cl:68-72) */

    /* frame-2 words Y | U << 8 | V << 16 of the two points (calcDeltaSumsKernel.cl:96-98), 0 outside the lattice */
    uint32_t v2a = 0u, v2b = 0u;
    {
        auto frame2_word = [&](int x, int y) -> uint32_t {
            const size_t iy = (size_t)y * P.W + x, iuv = (size_t)(y >> 1) * P.W + (x & ~1);
            if (P.bps == 1) {
                const uint8_t *fy = (const uint8_t *)P.f2y, *fuv = (const uint8_t *)P.f2uv;
                return (uint32_t)__ldg(fy + iy) | ((uint32_t)__ldg(fuv + iuv) << 8) | ((uint32_t)__ldg(fuv + iuv + 1) << 16);
            }
            const uint16_t *fy = (const uint16_t *)P.f2y, *fuv = (const uint16_t *)P.f2uv;
            return ((uint32_t)__ldg(fy + iy) >> 8) | ((uint32_t)__ldg(fuv + iuv) & 0xff00u) | (((uint32_t)__ldg(fuv + iuv + 1) & 0xff00u) << 8);
        };
        if (vA) v2a = frame2_word(cxs, cy0s);
        if (vB) v2b = frame2_word(cxs, cy1s);
    }

    int ox = 0, oy = 0;
    uint32_t nw[4] = {0u, 0u, 0u, 0u};
    const int pitch = P.planePitch, planeSize = P.planeSize;
    const uint32_t *__restrict__ p1 = P.p1;

    /* STAGED: word of this thread's upper point in plane 0 at zero displacement */
    const int stLane = ((warp >> 2) * 8 + (lane >> 3) * 2 + HR_ST_Y) * HR_ST_PW + (warp & 3) * 8 + (lane & 7) + HR_ST_X0;
    bool stageReady = false; /* has this warp seen the fetch complete? */

    const int steps = 2 * P.iters;
#pragma unroll 1
    for (int step = 0; step < steps; ++step) {
        __syncwarp(); /* the lanes of a warp leave a step apart when they polled for different words */
        const int it = step >> 1, axis = step & 1, par = step & 1;
        const int ws = P.first >> it;
        const bool useNb = it >= HR_FIRST_NEIGHBOR_ITERATION;
        /* the window both points belong to (tiles, warps and thread pairs are aligned to every window size) */
        S2_STAMP(1 + 4 * step)
        S2_AT(0)
        const int x0 = px & -ws, y0 = py & -ws;
        const bool winOk = x0 < P.lw && y0 < P.lh;
        const uint32_t count = winOk ? (uint32_t)(hr_min(x0 + ws, P.lw) - x0) * (uint32_t)(hr_min(y0 + ws, P.lh) - y0) : 0u;

        /* neighbour windows of the level: request now, look at them before the scoring */
        const bool wantNb = useNb && axis == 0 && winOk;
        unsigned long long nbv[4] = {0ull, 0ull, 0ull, 0ull};
        if (wantNb) neighbours_issue(P, it, ws, x0, y0, nbv);

        /* ---- samples of frame 1 at the R shifted positions: acc[z] = SAD of the thread's two points ---- */
        uint32_t acc[RT];
        if (warpValid) {
            uint32_t va[RT], vb[RT];
            const int xs = (px << s) + ox, ys = (py << s) + oy;
            const int low = vB ? one : 0;
            bool ok;
            if (axis == 0) ok = xs + CMIN >= 0 && xs + CMAX < P.W && ys >= 0 && ys + low < P.H;
            else ok = xs >= 0 && xs < P.W && ys + CMIN >= 0 && ys + low + CMAX < P.H;
            const int mulA = axis ? (planeSize << s) : planeSize, mulB = axis ? pitch : 1;
            /* bit 0: no layer of this lane leaves the frame; bit 1: ... nor the staged halo */
            unsigned where = ok ? 1u : 0u;
            if (STAGED) {
                if (!stageReady) stageReady = mbar_test(&stageBar, 0);
                const int xlo = (ox + (axis ? 0 : CMIN)) >> HR_ST_S, xhi = (ox + (axis ? 0 : CMAX)) >> HR_ST_S;
                const int ylo = (oy + (axis ? CMIN : 0)) >> HR_ST_S, yhi = (oy + (axis ? CMAX : 0)) >> HR_ST_S;
                if (ok && stageReady && xlo >= -HR_ST_X0 && xhi <= HR_ST_X1 && ylo >= -HR_ST_Y && yhi <= HR_ST_Y) where = 3u;
            }
            if (!vA) where = 3u;
            where = __reduce_and_sync(0xffffffffu, where);
            if (STAGED && (where & 2u)) {
                /* every sample of the warp is in shared memory: plane (phase) and cell displacement are the same
                 * arithmetic as below on the staged box; the lower point is one staged row further */
                constexpr int sm = (1 << HR_ST_S) - 1;
                const int F = axis ? (ox & sm) * HR_ST_PLANE + (ox >> HR_ST_S) : ((oy & sm) << HR_ST_S) * HR_ST_PLANE + (oy >> HR_ST_S) * HR_ST_PW;
                const int Mv = axis ? oy : ox;
                const int sA = axis ? (HR_ST_PLANE << HR_ST_S) : HR_ST_PLANE, sB = axis ? HR_ST_PW : 1;
                const uint32_t *base = stage + stLane + F;
#pragma unroll
                for (int z = 0; z < RT; ++z) {
                    const int p = Mv + signed_square(z - RT / 2);
                    const uint32_t *q = base + ((p & sm) * sA + (p >> HR_ST_S) * sB);
                    va[z] = vA ? q[0] : 0u;
                    vb[z] = vB ? q[HR_ST_PW] : 0u;
                }
            } else if (where & 1u) {
                /* no reflection anywhere in the warp; the lower point is one packed row below the upper one */
                const int F = axis ? (xs & m) * planeSize + (xs >> s) : ((ys & m) << s) * planeSize + (ys >> s) * pitch;
                const int Mv = axis ? ys : xs;
#pragma unroll
                for (int z = 0; z < RT; ++z) {
                    const int p = Mv + signed_square(z - RT / 2);
                    const uint32_t *q = p1 + (F + (p & m) * mulA + (p >> s) * mulB);
                    va[z] = vA ? __ldg(q) : 0u;
                    vb[z] = vB ? __ldg(q + pitch) : 0u;
                }
            } else {
                int fa, fb, ma, mb;
                if (axis == 0) {
                    const int ya = search_mirror(cy0s + oy, P.H), yb = search_mirror(cy1s + oy, P.H);
                    fa = ((ya & m) << s) * planeSize + (ya >> s) * pitch;
                    fb = ((yb & m) << s) * planeSize + (yb >> s) * pitch;
                    ma = mb = cxs + ox;
                } else {
                    const int x = search_mirror(cxs + ox, P.W);
                    fa = fb = (x & m) * planeSize + (x >> s);
                    ma = cy0s + oy;
                    mb = cy1s + oy;
                }
                const int D = axis ? P.H : P.W;
#pragma unroll
                for (int z = 0; z < RT; ++z) {
                    const int c = signed_square(z - RT / 2);
                    const int pa = search_mirror(ma + c, D), pb = search_mirror(mb + c, D);
                    va[z] = vA ? __ldg(p1 + (fa + (pa & m) * mulA + (pa >> s) * mulB)) : 0u;
                    vb[z] = vB ? __ldg(p1 + (fb + (pb & m) * mulA + (pb >> s) * mulB)) : 0u;
                }
            }
            S2_STAMP(2 + 4 * step)
            S2_AT(1)
#pragma unroll
            for (int z = 0; z < RT; ++z) acc[z] = sad4_acc(vb[z], v2b, sad4_acc(va[z], v2a, 0u));
        } else {
#pragma unroll
            for (int z = 0; z < RT; ++z) acc[z] = 0u;
        }

        int winner = 0;
        const int cur = axis ? oy : ox;
        if (ws >= 8) {
            /* the 8x8 block of the warp lies in one window: lane z keeps the block total of layer z */
            uint32_t mine = 0u;
#pragma unroll
            for (int z = 0; z < RT; ++z) {
                const uint32_t r = __reduce_add_sync(0xffffffffu, acc[z]);
                if (lane == z) mine = r;
            }
            S2_AT(2)
            if (ws >= 16) {
                sh.warpTot[par][warp][lane] = mine;
                __syncthreads();
                S2_AT(3)
                if (ws == 16) {
                    const int w0 = warp & 10; /* first warp of the 2x2 warp group */
                    mine = sh.warpTot[par][w0][lane] + sh.warpTot[par][w0 + 1][lane] + sh.warpTot[par][w0 + 4][lane] + sh.warpTot[par][w0 + 5][lane];
                } else if (ws == HR_TILE) {
                    mine = 0u;
#pragma unroll
                    for (int w = 0; w < HR_NWARPS; ++w) mine += sh.warpTot[par][w][lane];
                } else {
                    /* the window spans tiles: publish this tile's totals, add up the totals of all its tiles (warp w
                     * takes tiles w, w + 16, ..; lane = layer), every warp then sums the 16 partial sums itself */
                    if (warp == 0) {
                        uint32_t tt = 0u;
#pragma unroll
                        for (int w = 0; w < HR_NWARPS; ++w) tt += sh.warpTot[par][w][lane];
                        put_tagged(P.partial + P.bigOff[step] + tile * HR_RMAX + lane, P.epoch, tt);
                    }
                    const int lgw = 31 - __clz(ws), lgt = lgw - 5, tpw = 1 << lgt;
                    const int ax0 = (tx0 >> lgw) << lgt, ay0 = (ty0 >> lgw) << lgt;
                    const unsigned long long *ps = P.partial + P.bigOff[step] + lane;
                    uint32_t sum = 0u;
                    for (int i0 = warp; i0 < tpw * tpw; i0 += 4 * HR_NWARPS) {
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
                                sum += (uint32_t)v[k];
                            }
                        }
                    }
                    sh.bigTot[par][warp][lane] = sum;
                    S2_AT(4)
                    __syncthreads();
                    mine = 0u;
#pragma unroll
                    for (int w = 0; w < HR_NWARPS; ++w) mine += sh.bigTot[par][w][lane];
                }
            }
            S2_STAMP(3 + 4 * step)
            S2_AT(5)
            /* score: lane z = layer z; first minimum (determineLowestLayerKernel.cl:13-18) */
            int n[4] = {0, 0, 0, 0};
            if (useNb) {
                if (wantNb) neighbours_wait(P, it, ws, x0, y0, nbv, nw);
                neighbour_axis(nw, axis, n);
            }
            const uint32_t S = lane < RT ? window_total(mine, candLane, cur, count, useNb, n, P.dS, P.nS) : 0xffffffffu;
            const uint32_t mn = __reduce_min_sync(0xffffffffu, S);
            winner = __ffs(__ballot_sync(0xffffffffu, S == mn && lane < RT)) - 1;
        } else {
            /* windows of 4x4 / 2x2 points inside the warp: every lane of a window gets its totals and scores them */
            if (ws == 4) {
#pragma unroll
                for (int z = 0; z < RT; ++z) {
                    uint32_t a = acc[z];
                    a += __shfl_xor_sync(0xffffffffu, a, 1);
                    a += __shfl_xor_sync(0xffffffffu, a, 2);
                    a += __shfl_xor_sync(0xffffffffu, a, 8);
                    acc[z] = a;
                }
            } else {
#pragma unroll
                for (int z = 0; z < RT; ++z) acc[z] += __shfl_xor_sync(0xffffffffu, acc[z], 1);
            }
            S2_STAMP(3 + 4 * step)
            S2_AT(5)
            int n[4] = {0, 0, 0, 0};
            if (useNb) {
                if (wantNb) neighbours_wait(P, it, ws, x0, y0, nbv, nw);
                neighbour_axis(nw, axis, n);
            }
            uint32_t best = 0xffffffffu;
#pragma unroll
            for (int z = 0; z < RT; ++z) {
                const uint32_t S = window_total(acc[z], signed_square(z - RT / 2), cur, count, useNb, n, P.dS, P.nS);
                if (z == 0 || S < best) {
                    best = S;
                    winner = z;
                }
            }
        }

        S2_STAMP(4 + 4 * step)
        S2_AT(6)
        /* adjustOffsetArrayKernel.cl:2-18 */
        const int shift = signed_square(winner - RT / 2);
        if (axis) oy += shift;
        else ox += shift;
        if (P.trace) {
            if (vA) P.trace[((size_t)step * P.lh + py) * P.lw + px] = (uint8_t)winner;
            if (vB) P.trace[((size_t)step * P.lh + py + 1) * P.lw + px] = (uint8_t)winner;
        }
        /* publish this level's window (neighbours of the next level, blur halo) */
        if (axis && vA && px == x0 && py == y0) {
            const int lgw = 31 - __clz(ws), nwx = (P.lw + ws - 1) >> lgw;
            put_tagged(P.T + P.tOff[it] + (y0 >> lgw) * nwx + (x0 >> lgw), P.epoch, (uint32_t)(uint16_t)ox | ((uint32_t)(uint16_t)oy << 16));
        }
    }

    S2_STAMP(100)
    /* raw offsets (offsetArray) */
    const size_t ln = (size_t)P.lw * P.lh;
    if (vA) {
        const size_t idx = (size_t)py * P.lw + px;
        P.off[idx] = (int16_t)ox;
        P.off[ln + idx] = (int16_t)oy;
    }
    if (vB) {
        const size_t idx = (size_t)(py + 1) * P.lw + px;
        P.off[idx] = (int16_t)ox;
        P.off[ln + idx] = (int16_t)oy;
    }

    /* ------------- blur the raw offsets (K4), reading the last level's window table --------------- */
    {
        const int lws = P.first >> (P.iters - 1); /* = 2 */
        const int lgl = 31 - __clz(lws);
        const int lnwx = (P.lw + lws - 1) >> lgl;
        const unsigned long long *Tl = P.T + P.tOff[P.iters - 1];
        int16_t *tX = sh.blur.tX, *tY = sh.blur.tY;
        int *hX = sh.blur.hX, *hY = sh.blur.hY;
        constexpr int NT = HR_THREADS;
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
                P.blur[idx] = (int16_t)bx;
                P.blur[ln + idx] = (int16_t)by;
                P.blurXY[idx] = (uint32_t)(uint16_t)bx | ((uint32_t)(uint16_t)by << 16);
            }
        }
    }
    S2_STAMP(101)
    if (DBG && P.timeline && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        P.timeline[blockIdx.x * HR_TIMELINE_SLOTS + HR_TIMELINE_SLOTS - 1] = (long long)gt;
    }
#undef S2_STAMP
#undef S2_AT
}

template <int RT, bool DBG = false>
__global__ void __launch_bounds__(HR_THREADS, 1) flow_search2_kernel(const __grid_constant__ FlowParams P) {
    flow_search2_body<RT, DBG, false>(P, nullptr);
}
/* resolution scalar 2, radius <= HR_ST_MAX_RADIUS; launched with HR_ST_BYTES of dynamic shared memory */
template <int RT, bool DBG = false>
__global__ void __launch_bounds__(HR_THREADS, 1) flow_search2_staged_kernel(const __grid_constant__ FlowParams P, const __grid_constant__ HrTensorMap tmap) {
    flow_search2_body<RT, DBG, true>(P, &tmap);
}

/*
 * hr_search3.cuh — third generation of the block-offset search (K1+K2+K3, 16 steps) + flow blur (K4): four lattice
 * points per thread.
 *
 * What the first two generations showed (profiles/r02_search_generations.txt): a step is bound by the NUMBER of warp
 * instructions an SM has to issue for it, ~400 per warp and step whatever the organisation, of which the loads and
 * SADs are a tenth; everything else is per THREAD (window origin, population, neighbour windows, range tests, sample
 * addressing, scoring, offset update, publishing). So: fewer threads. A thread owns a 2x2 block of lattice points
 * (the smallest window), a warp a 16x8 block, a CTA of 8 warps the 32x32 tile:
 *   - per-thread work is issued half as often; per layer a thread adds two loads and two SADs (the right-hand points
 *     are the next word of the same packed row, the lower points one packed row down);
 *   - window 2 needs no exchange at all, window 4 two shuffles, window 8 four (a warp holds two 8x8 windows side by
 *     side), windows 16 and up one REDUX per layer (lane z keeps layer z) and the shared-memory exchange of
 *     generation 2 between 2 / 8 warps;
 *   - twenty loads in flight per thread at R = 5, at most 32 (eight layers at a time) for the larger radii;
 *   - 256 threads x <= 80 registers: THREE CTAs fit an SM, so the searches of three consecutive frame pairs (pipelined
 *     mode, three search lanes) run side by side on the same SMs — measured 39.0 -> 31.5 us per source frame at 1080p
 *     R = 5 with two lanes at two CTAs per SM (cap 128), 25.7 us with three at three, where one launch alone is no
 *     faster than the first generation (41.6 vs 39.9 us). At 80 registers radii up to 11 keep everything in registers,
 *     12..16 spill 4 to 40 bytes per thread (ptxas -v) and still gain (R = 16: 48.7 -> 45.5 us).
 * Tables, tile totals, outputs: the same words at the same places as hr_search.cuh — interchangeable launch by launch.
 *
 * Reference semantics: video/filter/HopperRender/Kernels/calcDeltaSumsKernel.cl:34-189,
 * determineLowestLayerKernel.cl:2-22, adjustOffsetArrayKernel.cl:2-18, blurFlowKernel.cl:15-89,
 * driver loop opticalFlowCalc.c:126-203.
 */
#pragma once
#include "hr_search2.cuh"

#define HR3_THREADS 256
#define HR3_NWARPS 8
#ifndef HR3_CTAS_PER_SM
#define HR3_CTAS_PER_SM 3 /* register cap 80, one CTA per search lane of hr_cuda.cu (2: cap 128, the two-lane build) */
#endif

struct Search3Shared {
    uint32_t warpTot[2][HR3_NWARPS][HR_RMAX]; /* per-warp block totals of a step, by step parity               */
    uint32_t bigTot[2][HR3_NWARPS][HR_RMAX];  /* per-warp sums of the other tiles' totals (cross-tile steps)    */
    struct {
        int16_t tX[40 * 40], tY[40 * 40];
        int hX[40 * 32], hY[40 * 32];
    } blur;
};

/* DBG: the stamps of hr_search2.cuh (slot 0, 1 + 4 * step + k, 100, 101, 126, 127). */
/* Several CTAs per SM: in pipelined mode the searches of consecutive frame pairs (the search lanes of hr_cuda.cu)
 * then run side by side on the same SMs, each filling the issue slots the others leave
 * empty while they wait — which the first generation could not do (512 threads x 88 registers: one CTA per SM). */
#ifdef HR3_MAXNREG
#define HR3_BOUNDS __maxnreg__(HR3_MAXNREG)
#else
#define HR3_BOUNDS __launch_bounds__(HR3_THREADS, HR3_CTAS_PER_SM)
#endif
template <int RT, bool DBG = false>
__global__ void HR3_BOUNDS flow_search3_kernel(const __grid_constant__ FlowParams P) {
    static_assert(RT >= 2 && RT <= HR_RMAX, "search radius");
    constexpr int ZC = RT < HR_ZCHUNK ? RT : HR_ZCHUNK; /* layers in flight at once: 4 * ZC loads per thread */
    __shared__ Search3Shared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#define S3_STAMP(slot) \
    if (DBG && P.timeline && tid == 0) { asm volatile("" ::: "memory"); P.timeline[blockIdx.x * HR_TIMELINE_SLOTS + (slot)] = clock64(); asm volatile("" ::: "memory"); }
    S3_STAMP(0)
    if (DBG && P.timeline && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        P.timeline[blockIdx.x * HR_TIMELINE_SLOTS + HR_TIMELINE_SLOTS - 2] = (long long)gt;
    }
    const int s = P.s, m = (1 << s) - 1, one = 1 << s;
    const int tile = blockIdx.x;
    const int tx0 = (tile % P.tilesX) * HR_TILE, ty0 = (tile / P.tilesX) * HR_TILE;
    /* the warp's 16x8 block and the thread's 2x2 points (px, py), (px + 1, py), (px, py + 1), (px + 1, py + 1) */
    const int wbx = tx0 + (warp & 1) * 16, wby = ty0 + (warp >> 1) * 8;
    const int px = wbx + (lane & 7) * 2, py = wby + (lane >> 3) * 2;
    const bool vx0 = px < P.lw, vx1 = px + 1 < P.lw, vy0 = py < P.lh, vy1 = py + 1 < P.lh;
    const bool v00 = vx0 && vy0, v01 = vx1 && vy0, v10 = vx0 && vy1, v11 = vx1 && vy1;
    const bool warpValid = wbx < P.lw && wby < P.lh; /* plain arithmetic, not a vote: see hr_search2.cuh */
    /* clamped full-resolution coordinates (the reflected path; the same numbers as hr_search.cuh) */
    const int cx0s = hr_min(px, P.lw - 1) << s, cx1s = hr_min(px + 1, P.lw - 1) << s;
    const int cy0s = hr_min(py, P.lh - 1) << s, cy1s = hr_min(py + 1, P.lh - 1) << s;
    constexpr int CMIN = -(RT / 2) * (RT / 2), CMAX = (RT - 1 - RT / 2) * (RT - 1 - RT / 2);
    const int candLane = signed_square(lane - RT / 2);

    /* frame-2 words Y | U << 8 | V << 16 of the four points (calcDeltaSumsKernel.cl:96-98), 0 outside the lattice */
    uint32_t w00 = 0u, w01 = 0u, w10 = 0u, w11 = 0u;
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
        if (v00) w00 = frame2_word(cx0s, cy0s);
        if (v01) w01 = frame2_word(cx1s, cy0s);
        if (v10) w10 = frame2_word(cx0s, cy1s);
        if (v11) w11 = frame2_word(cx1s, cy1s);
    }

    int ox = 0, oy = 0;
    uint32_t nw[4] = {0u, 0u, 0u, 0u};
    const int pitch = P.planePitch, planeSize = P.planeSize;
    const uint32_t *__restrict__ p1 = P.p1;

    const int steps = 2 * P.iters;
#pragma unroll 1
    for (int step = 0; step < steps; ++step) {
        __syncwarp();
        const int it = step >> 1, axis = step & 1, par = step & 1;
        const int ws = P.first >> it;
        const bool useNb = it >= HR_FIRST_NEIGHBOR_ITERATION;
        S3_STAMP(1 + 4 * step)
        const int x0 = px & -ws, y0 = py & -ws;
        const bool winOk = x0 < P.lw && y0 < P.lh;
        const uint32_t count = winOk ? (uint32_t)(hr_min(x0 + ws, P.lw) - x0) * (uint32_t)(hr_min(y0 + ws, P.lh) - y0) : 0u;

        const bool wantNb = useNb && axis == 0 && winOk;
        unsigned long long nbv[4] = {0ull, 0ull, 0ull, 0ull};
        if (wantNb) neighbours_issue(P, it, ws, x0, y0, nbv);

        /* ---- acc[z] = SAD of the thread's four points for layer z ---- */
        uint32_t acc[RT];
        if (warpValid) {
            const int xs = (px << s) + ox, ys = (py << s) + oy;
            const int right = vx1 ? one : 0, low = vy1 ? one : 0;
            bool ok;
            if (axis == 0) ok = xs + CMIN >= 0 && xs + right + CMAX < P.W && ys >= 0 && ys + low < P.H;
            else ok = xs >= 0 && xs + right < P.W && ys + CMIN >= 0 && ys + low + CMAX < P.H;
            const int mulA = axis ? (planeSize << s) : planeSize, mulB = axis ? pitch : 1;
            if (__all_sync(0xffffffffu, ok || !v00)) {
                /* no reflection anywhere in the warp: right-hand points one word on, lower points one packed row down */
                const int F = axis ? (xs & m) * planeSize + (xs >> s) : ((ys & m) << s) * planeSize + (ys >> s) * pitch;
                const int Mv = axis ? ys : xs;
#pragma unroll
                for (int z0 = 0; z0 < RT; z0 += ZC) {
                    uint32_t a00[ZC], a01[ZC], a10[ZC], a11[ZC];
#pragma unroll
                    for (int j = 0; j < ZC; ++j) {
                        if (z0 + j < RT) {
                            const int p = Mv + signed_square(z0 + j - RT / 2);
                            const uint32_t *q = p1 + (F + (p & m) * mulA + (p >> s) * mulB);
                            a00[j] = v00 ? __ldg(q) : 0u;
                            a01[j] = v01 ? __ldg(q + 1) : 0u;
                            a10[j] = v10 ? __ldg(q + pitch) : 0u;
                            a11[j] = v11 ? __ldg(q + pitch + 1) : 0u;
                        }
                    }
                    if (z0 == 0) { S3_STAMP(2 + 4 * step) }
#pragma unroll
                    for (int j = 0; j < ZC; ++j)
                        if (z0 + j < RT) acc[z0 + j] = sad4_acc(a11[j], w11, sad4_acc(a10[j], w10, sad4_acc(a01[j], w01, sad4_acc(a00[j], w00, 0u))));
                }
            } else {
                /* reflected path: every point on its own (the numbers of hr_search.cuh) */
                const int D = axis ? P.H : P.W;
                int fx0 = 0, fx1 = 0, fy0 = 0, fy1 = 0; /* the fixed coordinate's share of the word index */
                if (axis == 0) {
                    const int ya = search_mirror(cy0s + oy, P.H), yb = search_mirror(cy1s + oy, P.H);
                    fy0 = ((ya & m) << s) * planeSize + (ya >> s) * pitch;
                    fy1 = ((yb & m) << s) * planeSize + (yb >> s) * pitch;
                } else {
                    const int xa = search_mirror(cx0s + ox, P.W), xb = search_mirror(cx1s + ox, P.W);
                    fx0 = (xa & m) * planeSize + (xa >> s);
                    fx1 = (xb & m) * planeSize + (xb >> s);
                }
                const int m0 = axis ? cy0s + oy : cx0s + ox, m1 = axis ? cy1s + oy : cx1s + ox;
#pragma unroll
                for (int z0 = 0; z0 < RT; z0 += ZC) {
                    uint32_t a00[ZC], a01[ZC], a10[ZC], a11[ZC];
#pragma unroll
                    for (int j = 0; j < ZC; ++j) {
                        if (z0 + j < RT) {
                            const int c = signed_square(z0 + j - RT / 2);
                            const int pa = search_mirror(m0 + c, D), pb = search_mirror(m1 + c, D);
                            const int ia = (pa & m) * mulA + (pa >> s) * mulB, ib = (pb & m) * mulA + (pb >> s) * mulB;
                            if (axis == 0) { /* the moving coordinate is x: ia / ib belong to the left / right points */
                                a00[j] = v00 ? __ldg(p1 + (fy0 + ia)) : 0u;
                                a01[j] = v01 ? __ldg(p1 + (fy0 + ib)) : 0u;
                                a10[j] = v10 ? __ldg(p1 + (fy1 + ia)) : 0u;
                                a11[j] = v11 ? __ldg(p1 + (fy1 + ib)) : 0u;
                            } else {         /* the moving coordinate is y: ia / ib belong to the upper / lower points */
                                a00[j] = v00 ? __ldg(p1 + (fx0 + ia)) : 0u;
                                a01[j] = v01 ? __ldg(p1 + (fx1 + ia)) : 0u;
                                a10[j] = v10 ? __ldg(p1 + (fx0 + ib)) : 0u;
                                a11[j] = v11 ? __ldg(p1 + (fx1 + ib)) : 0u;
                            }
                        }
                    }
                    if (z0 == 0) { S3_STAMP(2 + 4 * step) }
#pragma unroll
                    for (int j = 0; j < ZC; ++j)
                        if (z0 + j < RT) acc[z0 + j] = sad4_acc(a11[j], w11, sad4_acc(a10[j], w10, sad4_acc(a01[j], w01, sad4_acc(a00[j], w00, 0u))));
                }
            }
        } else {
#pragma unroll
            for (int z = 0; z < RT; ++z) acc[z] = 0u;
        }

        int winner = 0;
        const int cur = axis ? oy : ox;
        if (ws >= 16) {
            /* the warp's 16x8 block lies in one window: lane z keeps the block total of layer z */
            uint32_t mine = 0u;
#pragma unroll
            for (int z = 0; z < RT; ++z) {
                const uint32_t r = __reduce_add_sync(0xffffffffu, acc[z]);
                if (lane == z) mine = r;
            }
            sh.warpTot[par][warp][lane] = mine;
            __syncthreads();
            if (ws == 16) {
                mine += sh.warpTot[par][warp ^ 2][lane]; /* the block above / below */
            } else if (ws == HR_TILE) {
                mine = 0u;
#pragma unroll
                for (int w = 0; w < HR3_NWARPS; ++w) mine += sh.warpTot[par][w][lane];
            } else {
                if (warp == 0) {
                    uint32_t tt = 0u;
#pragma unroll
                    for (int w = 0; w < HR3_NWARPS; ++w) tt += sh.warpTot[par][w][lane];
                    put_tagged(P.partial + P.bigOff[step] + tile * HR_RMAX + lane, P.epoch, tt);
                }
                const int lgw = 31 - __clz(ws), lgt = lgw - 5, tpw = 1 << lgt;
                const int ax0 = (tx0 >> lgw) << lgt, ay0 = (ty0 >> lgw) << lgt;
                const unsigned long long *ps = P.partial + P.bigOff[step] + lane;
                uint32_t sum = 0u;
                for (int i0 = warp; i0 < tpw * tpw; i0 += 4 * HR3_NWARPS) {
                    const unsigned long long *q[4];
                    unsigned long long v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = i0 + k * HR3_NWARPS;
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
                __syncthreads();
                mine = 0u;
#pragma unroll
                for (int w = 0; w < HR3_NWARPS; ++w) mine += sh.bigTot[par][w][lane];
            }
            S3_STAMP(3 + 4 * step)
            int n[4] = {0, 0, 0, 0};
            if (useNb) {
                if (wantNb) neighbours_wait(P, it, ws, x0, y0, nbv, nw);
                neighbour_axis(nw, axis, n);
            }
            const uint32_t S = lane < RT ? window_total(mine, candLane, cur, count, useNb, n, P.dS, P.nS) : 0xffffffffu;
            const uint32_t mn = __reduce_min_sync(0xffffffffu, S);
            winner = __ffs(__ballot_sync(0xffffffffu, S == mn && lane < RT)) - 1;
        } else {
            /* windows of 8x8 (4x4 threads: lane bits 0, 1, 3, 4), 4x4 (2x2 threads: bits 0, 3) or 2x2 points (the thread) */
            if (ws == 8) {
#pragma unroll
                for (int z = 0; z < RT; ++z) {
                    uint32_t a = acc[z];
                    a += __shfl_xor_sync(0xffffffffu, a, 1);
                    a += __shfl_xor_sync(0xffffffffu, a, 2);
                    a += __shfl_xor_sync(0xffffffffu, a, 8);
                    a += __shfl_xor_sync(0xffffffffu, a, 16);
                    acc[z] = a;
                }
            } else if (ws == 4) {
#pragma unroll
                for (int z = 0; z < RT; ++z) {
                    uint32_t a = acc[z];
                    a += __shfl_xor_sync(0xffffffffu, a, 1);
                    a += __shfl_xor_sync(0xffffffffu, a, 8);
                    acc[z] = a;
                }
            }
            S3_STAMP(3 + 4 * step)
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
        S3_STAMP(4 + 4 * step)

        /* adjustOffsetArrayKernel.cl:2-18 */
        const int shift = signed_square(winner - RT / 2);
        if (axis) oy += shift;
        else ox += shift;
        if (P.trace) {
            uint8_t *tr = P.trace + (size_t)step * P.lh * P.lw;
            if (v00) tr[(size_t)py * P.lw + px] = (uint8_t)winner;
            if (v01) tr[(size_t)py * P.lw + px + 1] = (uint8_t)winner;
            if (v10) tr[(size_t)(py + 1) * P.lw + px] = (uint8_t)winner;
            if (v11) tr[(size_t)(py + 1) * P.lw + px + 1] = (uint8_t)winner;
        }
        /* publish this level's window (neighbours of the next level, blur halo) */
        if (axis && v00 && px == x0 && py == y0) {
            const int lgw = 31 - __clz(ws), nwx = (P.lw + ws - 1) >> lgw;
            put_tagged(P.T + P.tOff[it] + (y0 >> lgw) * nwx + (x0 >> lgw), P.epoch, (uint32_t)(uint16_t)ox | ((uint32_t)(uint16_t)oy << 16));
        }
    }
    S3_STAMP(100)

    /* raw offsets (offsetArray) */
    const size_t ln = (size_t)P.lw * P.lh;
    {
        const int16_t vx = (int16_t)ox, vy = (int16_t)oy;
        if (v00) { const size_t i = (size_t)py * P.lw + px; P.off[i] = vx; P.off[ln + i] = vy; }
        if (v01) { const size_t i = (size_t)py * P.lw + px + 1; P.off[i] = vx; P.off[ln + i] = vy; }
        if (v10) { const size_t i = (size_t)(py + 1) * P.lw + px; P.off[i] = vx; P.off[ln + i] = vy; }
        if (v11) { const size_t i = (size_t)(py + 1) * P.lw + px + 1; P.off[i] = vx; P.off[ln + i] = vy; }
    }

    /* ------------- blur the raw offsets (K4), reading the last level's window table --------------- */
    {
        const int lws = P.first >> (P.iters - 1); /* = 2 */
        const int lgl = 31 - __clz(lws);
        const int lnwx = (P.lw + lws - 1) >> lgl;
        const unsigned long long *Tl = P.T + P.tOff[P.iters - 1];
        int16_t *tX = sh.blur.tX, *tY = sh.blur.tY;
        int *hX = sh.blur.hX, *hY = sh.blur.hY;
        constexpr int NT = HR3_THREADS;
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
    S3_STAMP(101)
    if (DBG && P.timeline && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        P.timeline[blockIdx.x * HR_TIMELINE_SLOTS + HR_TIMELINE_SLOTS - 1] = (long long)gt;
    }
#undef S3_STAMP
}

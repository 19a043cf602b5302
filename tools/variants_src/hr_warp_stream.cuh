/*
 * hr_warp_stream.cuh — K5 (warpFrameKernel.cl:114-182), BlendedFrame mode, for large frames (4K and up):
 * the same arithmetic and the same per-thread unit as warp_fast_kernel (4 samples x 4 rows inside one lattice
 * cell, hr_warp_fast.cuh), organised as a PERSISTENT, SOFTWARE-PIPELINED kernel.
 *
 * Why: at 4K P010 the one-shot kernel is latency-bound, not bandwidth- or issue-bound (ncu: 11.8 of 18 stall
 * cycles per issue are long-scoreboard; DRAM 24 %, issue 38 %). Every thread runs three dependent memory round
 * trips (flow word -> flipped flow word -> samples) before it can compute, all CTAs of a wave do so in lock
 * step, and the last wave leaves most SMs idle (SM-active 36 k of 53 k cycles). Here
 *   - the grid is sized to the machine once (occupancy x SMs); warp w walks the units w, w + NW, w + 2 NW, ...
 *     of the luma plane, then of the chroma plane (shifted so that the load stays balanced): no waves, no tail;
 *   - a warp has four units in flight, one per stage:
 *         A  unit k     : coordinates, issue the load of the cell's flow word
 *         B  unit k-1   : flip index from the flow word, issue the load of the flipped word
 *         C  unit k-2   : displacements, range tests, issue the 4 x 2 source-run loads into the spare register set
 *         D  unit k-3   : blend + levels + store from the other register set
 *     so the three round trips of one unit overlap the arithmetic of the three units before it.
 * A unit = one warp x (4 samples x 4 rows) = 128 samples x 4 rows of one plane.
 */
#pragma once
#include "hr_warp_fast.cuh"

struct StreamUnit {
    int cx0, cy0;   /* first sample / first row of this thread's block                    */
    uint32_t w12;   /* the cell's flow word (valid one iteration after stage A)            */
    uint32_t w21;   /* the flipped cell's flow word (valid one iteration after stage B)    */
    int flags;      /* 0: nothing to do, 1: block path, 2: per-sample path (partial block) */
};

template <typename T, bool CHROMA, int VAR>
__device__ __forceinline__ void warp_stream_plane(const WarpParams<T> &P, const WarpFastArgs &A, int first, int NW, int lane) {
    constexpr bool is16 = SampleTraits<T>::is16;
    typedef typename RunType<T>::type Run;
    constexpr int cz = CHROMA ? 1 : 0;
    constexpr int ROWS = 4;
    const int planeH = CHROMA ? (P.H >> 1) : P.H;
    const T *s12 = CHROMA ? P.f1uv : P.f1y;
    const T *s21 = CHROMA ? P.f2uv : P.f2y;
    T *outp = CHROMA ? P.outUV : P.outY;
    const int s = P.s;
    const int colBlocks = (P.aW + 127) >> 7;
    const int groups = CHROMA ? A.chromaGN : A.lumaGroups;
    const int g0 = CHROMA ? A.chromaG0 : A.lumaG0;
    const int nUnits = groups * colBlocks;
    if (first >= nUnits) return;
    const int n = (nUnits - first + NW - 1) / NW; /* units of this warp */
    /* unit index -> (row group, column block), advanced incrementally */
    int g = first / colBlocks, cb = first - g * colBlocks;
    const int dq = NW / colBlocks, dr = NW - dq * colBlocks;
    const BlendK K = make_blendk<is16>(P.t12, P.t21, A, cz);

    StreamUnit uB = {0, 0, 0u, 0u, 0}, uC = {0, 0, 0u, 0u, 0}, uD = {0, 0, 0u, 0u, 0};
    bool dInterior = false;

    auto stageA = [&](int k) -> StreamUnit {
        StreamUnit u = {0, 0, 0u, 0u, 0};
        if (k < n) {
            u.cx0 = (cb * 32 + lane) * 4;
            u.cy0 = (g0 + g) * ROWS;
            cb += dr;
            g += dq;
            if (cb >= colBlocks) {
                cb -= colBlocks;
                g += 1;
            }
            if (u.cx0 < P.aW && u.cy0 < planeH) {
                u.flags = (u.cx0 + 3 >= P.aW || u.cy0 + ROWS > planeH) ? 2 : 1;
                int lx = u.cx0 >> s, ly = u.cy0 >> s;
                if (CHROMA) {
                    lx &= ~1;
                    ly <<= 1;
                }
                if (u.flags == 1) u.w12 = __ldg(P.flowXY + ly * P.lw + lx);
            }
        }
        return u;
    };
    auto stageB = [&](StreamUnit &u) {
        if (u.flags == 1) {
            int lx = u.cx0 >> s, ly = u.cy0 >> s;
            if (CHROMA) {
                lx &= ~1;
                ly <<= 1;
            }
            const int x12 = (int)(int16_t)u.w12, y12 = (int)u.w12 >> 16;
            const int fy = hr_min(hr_max(ly - (y12 >> s), 0), P.lh - 1);
            const int fx = hr_min(hr_max(lx - (x12 >> s), 0), P.lw - 1);
            u.w21 = __ldg(P.flowXY + fy * P.lw + fx);
        }
    };
    /* issue the source loads of unit u into (na, nb); returns whether the block lies inside the frame */
    auto stageC = [&](const StreamUnit &u, Run (&na)[ROWS], Run (&nb)[ROWS]) -> bool {
        if (u.flags != 1) return false;
        const int x12 = (int)(int16_t)u.w12, y12 = (int)u.w12 >> 16, x21 = (int)(int16_t)u.w21, y21 = (int)u.w21 >> 16;
        float fe12 = (float)y12 * P.t12, fe21 = (float)y21 * P.t21;
        if (CHROMA) {
            fe12 *= 0.5f;
            fe21 *= 0.5f;
        }
        const int d12 = round_half_away((float)x12 * P.t12), d21 = -round_half_away((float)x21 * P.t21);
        const int e12 = round_half_away(fe12), e21 = -round_half_away(fe21);
        const int a12 = u.cx0 + d12, a21 = u.cx0 + d21, b12 = u.cy0 + e12, b21 = u.cy0 + e21;
        const bool interior = (unsigned)(a12 - 1) <= (unsigned)(P.aW - 6) && (unsigned)(a21 - 1) <= (unsigned)(P.aW - 6) &&
                              (unsigned)(b12 - 1) <= (unsigned)(planeH - ROWS - 2) && (unsigned)(b21 - 1) <= (unsigned)(planeH - ROWS - 2);
        if (interior) {
            RunSrc<T> A12, A21;
            A12.set(s12, b12 * P.W + a12, P.W, CHROMA && (d12 & 1));
            A21.set(s21, b21 * P.W + a21, P.W, CHROMA && (d21 & 1));
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                na[r] = A12.template row<CHROMA>(r);
                nb[r] = A21.template row<CHROMA>(r);
            }
        } else {
            /* a frame border is involved: every sample through the mirror + clamp, chroma through the pair rule */
#pragma unroll 1
            for (int r = 0; r < ROWS; ++r) {
                const Run a = load_run4_border(s12 + (size_t)warp_mirror(b12 + r, planeH) * P.W, u.cx0, d12, P.aW, cz);
                const Run b = load_run4_border(s21 + (size_t)warp_mirror(b21 + r, planeH) * P.W, u.cx0, d21, P.aW, cz);
                /* blend and store right away: the register sets are indexed statically only */
                store_run(outp + (size_t)(u.cy0 + r) * P.W + u.cx0, blend_run<CHROMA, VAR >= 2, VAR != 3>(K, a, b));
            }
        }
        return interior;
    };
    auto stageD = [&](const StreamUnit &u, bool interior, const Run (&ca)[ROWS], const Run (&cbuf)[ROWS]) {
        if (u.flags == 1) {
            if (interior) {
                T *po = outp + u.cy0 * P.W + u.cx0;
#pragma unroll
                for (int r = 0; r < ROWS; ++r) store_run(po + r * P.W, blend_run<CHROMA, VAR >= 2, VAR != 3>(K, ca[r], cbuf[r]));
            }
        } else if (u.flags == 2) {
            warp_thread_slow(P, u.cx0, u.cy0, cz, hr_min(ROWS, planeH - u.cy0));
        }
    };

    Run sa0[ROWS], sb0[ROWS], sa1[ROWS], sb1[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) sa0[r] = sb0[r] = sa1[r] = sb1[r] = Run();
    /* iteration k: A(k), B(k-1), C(k-2) into the spare set, D(k-3) from the current set */
    auto iter = [&](int k, Run (&ca)[ROWS], Run (&cbuf)[ROWS], Run (&na)[ROWS], Run (&nb)[ROWS]) {
        StreamUnit uA = stageA(k);
        stageB(uB);
        const bool cInterior = stageC(uC, na, nb);
        stageD(uD, dInterior, ca, cbuf);
        uD = uC;
        dInterior = cInterior;
        uC = uB;
        uB = uA;
    };
    const int total = n + 3;
    for (int k = 0; k < total; k += 2) {
        iter(k, sa0, sb0, sa1, sb1);
        iter(k + 1, sa1, sb1, sa0, sb0);
    }
}

template <typename T>
__global__ void __launch_bounds__(128) warp_stream_kernel(const __grid_constant__ WarpParams<T> P, const __grid_constant__ WarpFastArgs A) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int NW = (gridDim.x * blockDim.x) >> 5;
    const int colBlocks = (P.aW + 127) >> 7;
    const int nL = A.lumaGroups * colBlocks;
    /* chroma unit c goes to warp (nL + c) mod NW: the chroma plane continues where the luma plane stopped */
    const int firstC = ((w - nL) % NW + NW) % NW;
    auto run = [&](auto chromaTag, int first) {
        constexpr bool CH = decltype(chromaTag)::value;
        const int cz = CH ? 1 : 0;
        const int var = A.subIsInt[cz] ? (A.clampNeeded[cz] ? 2 : 1) : 3;
        if (var == 1) warp_stream_plane<T, CH, 1>(P, A, first, NW, lane);
        else if (var == 2) warp_stream_plane<T, CH, 2>(P, A, first, NW, lane);
        else warp_stream_plane<T, CH, 3>(P, A, first, NW, lane);
    };
    run(std::false_type(), w);
    run(std::true_type(), firstC);
}

/*
 * hr_warp_stage.cuh — K5 (warpFrameKernel.cl:114-182), BlendedFrame mode, for large frames (4K and up): the same
 * arithmetic and the same per-thread unit as warp_fast_kernel (4 samples x 4 rows inside one lattice cell,
 * hr_warp_fast.cuh), as a PERSISTENT kernel whose source samples travel global -> shared memory by asynchronous
 * 16-byte copies (cp.async), two units ahead of the arithmetic.
 *
 * Why: at 4K P010 the one-shot kernel is latency-bound (ncu: 12 of 18 stall cycles per issue are long-scoreboard,
 * DRAM 24 %, issue 43 %): every thread runs three dependent round trips — flow word, flipped flow word, samples —
 * before it can compute, and the registers that hold the samples in flight cap the number of resident warps. Here
 * a warp owns a private, double-buffered staging area in shared memory and keeps four units in flight:
 *     A  unit k     : coordinates, issue the load of the cell's flow word
 *     B  unit k-1   : flip index, issue the load of the flipped word
 *     C  unit k-2   : displacements, the warp's bounding box of both source blocks (shuffles), cp.async of the box
 *                     rows into the spare staging buffer
 *     D  unit k-3   : wait for its copies, read the unaligned 4-sample runs from shared memory, blend, store
 * so the round trips of a unit overlap the arithmetic of the three before it at the cost of a handful of registers.
 * A unit = one warp x (4 samples x 4 rows) = 128 samples x 4 rows of one plane. Units whose box does not fit the
 * staging buffer (displacements that differ by more than 32 samples / 4 rows inside the unit), that touch a frame
 * border or a partial column are computed straight from global memory, exactly like warp_fast_kernel does.
 */
#pragma once
#include "hr_warp_fast.cuh"

#define HR_STAGE_ROWS 8 /* staged source rows per unit: 4 + the spread of the vertical displacement */

template <typename T>
struct StageGeom {
    /* bytes per staged row: 128 samples + 32 samples of horizontal spread + 3 of the odd-chroma pick + 16-byte alignment */
    static constexpr int PITCH = (int)sizeof(T) == 1 ? 192 : 352;
    static constexpr int BUF = PITCH * HR_STAGE_ROWS; /* one source block */
};

__device__ __forceinline__ void cp_async16(void *smemDst, const void *gmemSrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmemSrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

/* four consecutive samples at byte offset `off` of a staged block (any 2-byte / 1-byte alignment) */
template <bool CHROMA>
__device__ __forceinline__ uint32_t staged_run(const uint8_t *buf, int off, bool odd, uint8_t) {
    if (odd) off -= 1;
    const uint32_t *p = reinterpret_cast<const uint32_t *>(buf + (off & ~3));
    const unsigned sh = (unsigned)(off & 3) * 8;
    const uint32_t w0 = p[0], w1 = p[1];
    const uint32_t lo = __funnelshift_r(w0, w1, sh);
    if (!CHROMA) return lo;
    const uint32_t w2 = p[2];
    const uint32_t hi = __funnelshift_r(w1, w2, sh);
    return __byte_perm(lo, hi, odd ? 0x5230u : 0x3210u);
}
template <bool CHROMA>
__device__ __forceinline__ uint2 staged_run(const uint8_t *buf, int off, bool odd, uint16_t) {
    if (odd) off -= 2;
    const uint32_t *p = reinterpret_cast<const uint32_t *>(buf + (off & ~3));
    const unsigned sh = (unsigned)(off & 2) * 8;
    const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
    if (!CHROMA) return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
    const uint32_t w3 = p[3];
    const uint32_t s0 = __funnelshift_r(w0, w1, sh), s1 = __funnelshift_r(w1, w2, sh), s2 = __funnelshift_r(w2, w3, sh);
    return odd ? make_uint2(__byte_perm(s0, s1, 0x7610), __byte_perm(s1, s2, 0x7610)) : make_uint2(s0, s1);
}

struct StageUnit {
    int cx0, cy0;        /* first sample / first row of this thread's block                                      */
    uint32_t w12, w21;   /* flow word of the cell, of the flipped cell                                            */
    int flags;           /* 0: nothing to do, 1: block path, 2: per-sample path (partial block)                   */
    int off12, off21;    /* stage D: byte offset of the thread's first run in the staged block (staged units), or  */
                         /*          sample offset of it in the plane (direct units)                              */
    int how;             /* stage D: 0 nothing, 1 staged, 2 direct from global memory, 3 frame border, 4 per sample */
    bool odd12, odd21;
    int b12, b21, d12, d21; /* direct / border units only: source rows and horizontal displacements                */
};

template <typename T, bool CHROMA, int VAR>
__device__ __forceinline__ void warp_stage_plane(const WarpParams<T> &P, const WarpFastArgs &A, int first, int NW, int lane, uint8_t *stage) {
    constexpr bool is16 = SampleTraits<T>::is16;
    typedef typename RunType<T>::type Run;
    constexpr int cz = CHROMA ? 1 : 0;
    constexpr int ROWS = 4, BPS = (int)sizeof(T), PITCH = StageGeom<T>::PITCH, BUF = StageGeom<T>::BUF;
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
    int g = first / colBlocks, cb = first - g * colBlocks;
    const int dq = NW / colBlocks, dr = NW - dq * colBlocks;
    const BlendK K = make_blendk<is16>(P.t12, P.t21, A, cz);
    const int rowBytes = P.W * BPS;

    auto stageA = [&](int k) -> StageUnit {
        StageUnit u;
        u.cx0 = u.cy0 = 0;
        u.w12 = u.w21 = 0u;
        u.flags = u.how = 0;
        u.off12 = u.off21 = u.b12 = u.b21 = u.d12 = u.d21 = 0;
        u.odd12 = u.odd21 = false;
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
    auto stageB = [&](StageUnit &u) {
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
    /* displacements of the unit, bounding boxes of the warp, asynchronous copies into `buf` (two source blocks) */
    auto stageC = [&](StageUnit &u, uint8_t *buf) {
        bool interior = false;
        int a12 = 0, a21 = 0;
        if (u.flags == 1) {
            const int x12 = (int)(int16_t)u.w12, y12 = (int)u.w12 >> 16, x21 = (int)(int16_t)u.w21, y21 = (int)u.w21 >> 16;
            float fe12 = (float)y12 * P.t12, fe21 = (float)y21 * P.t21;
            if (CHROMA) {
                fe12 *= 0.5f;
                fe21 *= 0.5f;
            }
            u.d12 = round_half_away((float)x12 * P.t12);
            u.d21 = -round_half_away((float)x21 * P.t21);
            a12 = u.cx0 + u.d12;
            a21 = u.cx0 + u.d21;
            u.b12 = u.cy0 + round_half_away(fe12);
            u.b21 = u.cy0 - round_half_away(fe21);
            u.odd12 = CHROMA && (u.d12 & 1);
            u.odd21 = CHROMA && (u.d21 & 1);
            interior = (unsigned)(a12 - 1) <= (unsigned)(P.aW - 6) && (unsigned)(a21 - 1) <= (unsigned)(P.aW - 6) &&
                       (unsigned)(u.b12 - 1) <= (unsigned)(planeH - ROWS - 2) && (unsigned)(u.b21 - 1) <= (unsigned)(planeH - ROWS - 2);
        }
        u.how = u.flags == 0 ? 0 : u.flags == 2 ? 4 : interior ? 2 : 3;
        /* the warp's boxes: byte range [x0, x1) of a row (chroma with an odd displacement reads one sample to the left
         * and two to the right of its run), rows [y0, y1] */
        const unsigned full = 0xffffffffu;
        const bool in = u.how == 2;
        const bool allIn = __all_sync(full, in || u.flags == 0) && __any_sync(full, in);
        if (allIn) {
            const int big = 0x3fffffff;
            int x0a = in ? (a12 - (CHROMA ? 1 : 0)) * BPS : big, x1a = in ? (a12 + 4 + (CHROMA ? 2 : 0)) * BPS : -big;
            int x0b = in ? (a21 - (CHROMA ? 1 : 0)) * BPS : big, x1b = in ? (a21 + 4 + (CHROMA ? 2 : 0)) * BPS : -big;
            int y0a = in ? u.b12 : big, y1a = in ? u.b12 + ROWS - 1 : -big, y0b = in ? u.b21 : big, y1b = in ? u.b21 + ROWS - 1 : -big;
            x0a = __reduce_min_sync(full, x0a) & ~15;
            x0b = __reduce_min_sync(full, x0b) & ~15;
            x1a = (__reduce_max_sync(full, x1a) + 15) & ~15;
            x1b = (__reduce_max_sync(full, x1b) + 15) & ~15;
            y0a = __reduce_min_sync(full, y0a);
            y0b = __reduce_min_sync(full, y0b);
            y1a = __reduce_max_sync(full, y1a);
            y1b = __reduce_max_sync(full, y1b);
            const int wa = x1a - x0a, wb = x1b - x0b, ha = y1a - y0a + 1, hb = y1b - y0b + 1;
            if (wa <= PITCH && wb <= PITCH && ha <= HR_STAGE_ROWS && hb <= HR_STAGE_ROWS && x1a <= rowBytes && x1b <= rowBytes && x0a >= 0 && x0b >= 0) {
                const uint8_t *ga = reinterpret_cast<const uint8_t *>(s12) + (size_t)y0a * rowBytes + x0a;
                const uint8_t *gb = reinterpret_cast<const uint8_t *>(s21) + (size_t)y0b * rowBytes + x0b;
                /* lane l copies 16-byte chunk l of every row it reaches: a row is at most PITCH / 16 = 12 (22) chunks, so
                 * one (two) passes over the lanes cover it; rows are strided over nothing — every lane walks all rows */
                const int ca = wa >> 4, cbk = wb >> 4; /* 16-byte chunks per row */
#pragma unroll 1
                for (int q = lane; q < ca; q += 32) {
                    const uint8_t *g = ga + q * 16;
                    uint8_t *d = buf + q * 16;
                    for (int r = 0; r < ha; ++r) cp_async16(d + r * PITCH, g + (size_t)r * rowBytes);
                }
#pragma unroll 1
                for (int q = lane; q < cbk; q += 32) {
                    const uint8_t *g = gb + q * 16;
                    uint8_t *d = buf + BUF + q * 16;
                    for (int r = 0; r < hb; ++r) cp_async16(d + r * PITCH, g + (size_t)r * rowBytes);
                }
                if (in) {
                    u.how = 1;
                    u.off12 = (u.b12 - y0a) * PITCH + a12 * BPS - x0a;
                    u.off21 = (u.b21 - y0b) * PITCH + a21 * BPS - x0b;
                }
            }
        }
        if (u.how == 2) { /* direct: remember where the runs start in the plane */
            u.off12 = u.b12 * P.W + a12;
            u.off21 = u.b21 * P.W + a21;
        }
        cp_async_commit();
    };
    auto stageD = [&](const StageUnit &u, const uint8_t *buf) {
        cp_async_wait<1>(); /* everything but the copies issued in this iteration */
        __syncwarp();
        if (u.how == 1) {
            T *po = outp + u.cy0 * P.W + u.cx0;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const Run a = staged_run<CHROMA>(buf, u.off12 + r * PITCH, u.odd12, T());
                const Run b = staged_run<CHROMA>(buf + BUF, u.off21 + r * PITCH, u.odd21, T());
                store_run(po + r * P.W, blend_run<CHROMA, VAR >= 2, VAR != 3>(K, a, b));
            }
        } else if (u.how == 2) {
            T *po = outp + u.cy0 * P.W + u.cx0;
            RunSrc<T> A12, A21;
            A12.set(s12, u.off12, P.W, u.odd12);
            A21.set(s21, u.off21, P.W, u.odd21);
            Run ra[ROWS], rb[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                ra[r] = A12.template row<CHROMA>(r);
                rb[r] = A21.template row<CHROMA>(r);
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) store_run(po + r * P.W, blend_run<CHROMA, VAR >= 2, VAR != 3>(K, ra[r], rb[r]));
        } else if (u.how == 3) {
            /* a frame border is involved: every sample through the mirror + clamp, chroma through the pair rule */
#pragma unroll 1
            for (int r = 0; r < ROWS; ++r) {
                const Run a = load_run4_border(s12 + (size_t)warp_mirror(u.b12 + r, planeH) * P.W, u.cx0, u.d12, P.aW, cz);
                const Run b = load_run4_border(s21 + (size_t)warp_mirror(u.b21 + r, planeH) * P.W, u.cx0, u.d21, P.aW, cz);
                store_run(outp + (size_t)(u.cy0 + r) * P.W + u.cx0, blend_run<CHROMA, VAR >= 2, VAR != 3>(K, a, b));
            }
        } else if (u.how == 4) {
            warp_thread_slow(P, u.cx0, u.cy0, cz, hr_min(ROWS, planeH - u.cy0));
        }
        __syncwarp(); /* the buffer is free for the copies of the unit after next */
    };

    StageUnit uB = stageA(n), uC = uB, uD = uB; /* k = n: an empty unit */
    const int total = n + 3;
    for (int k = 0; k < total; ++k) {
        StageUnit uA = stageA(k);
        stageB(uB);
        stageC(uC, stage + (k & 1) * 2 * BUF);
        stageD(uD, stage + ((k + 1) & 1) * 2 * BUF);
        uD = uC;
        uC = uB;
        uB = uA;
    }
    cp_async_wait<0>();
}

template <typename T>
__global__ void __launch_bounds__(128) warp_stage_kernel(const __grid_constant__ WarpParams<T> P, const __grid_constant__ WarpFastArgs A) {
    extern __shared__ __align__(128) uint8_t hr_stage_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int NW = (gridDim.x * blockDim.x) >> 5;
    uint8_t *stage = hr_stage_smem + (size_t)wib * 4 * StageGeom<T>::BUF; /* per warp: 2 buffers x 2 source blocks */
    const int colBlocks = (P.aW + 127) >> 7;
    const int nL = A.lumaGroups * colBlocks;
    /* chroma unit c goes to warp (nL + c) mod NW: the chroma plane continues where the luma plane stopped */
    const int firstC = ((w - nL) % NW + NW) % NW;
    auto run = [&](auto chromaTag, int first) {
        constexpr bool CH = decltype(chromaTag)::value;
        const int cz = CH ? 1 : 0;
        const int var = A.subIsInt[cz] ? (A.clampNeeded[cz] ? 2 : 1) : 3;
        if (var == 1) warp_stage_plane<T, CH, 1>(P, A, first, NW, lane, stage);
        else if (var == 2) warp_stage_plane<T, CH, 2>(P, A, first, NW, lane, stage);
        else warp_stage_plane<T, CH, 3>(P, A, first, NW, lane, stage);
    };
    run(std::false_type(), w);
    run(std::true_type(), firstC);
}

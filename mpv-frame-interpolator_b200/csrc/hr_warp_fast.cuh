/*
 * hr_warp_fast.cuh — K5 (warpFrameKernel.cl:114-182) for the cases a player produces: output modes 0, 1,
 * 2 and 5, resolution scalar >= 2 (every frame taller than 540 lines), planes aligned to 8 bytes, level
 * knobs for which the reference's division is MUFU.RCP * x. Everything else runs warp_generic_kernel
 * (hr_warp.cuh), which computes the same numbers one sample at a time.
 *
 * HBM-bound by nature (2 frames read, 1 written, 518 KB of flow), issue-bound in practice: the work per
 * output sample is two byte->float conversions, the blend, two truncations and the level map. The kernel
 * is organised around the instruction count:
 *   - thread = 4 samples x 4 rows inside ONE lattice cell, so that
 *     the flow vector, the flipped vector (warpFrameKernel.cl:155-156), the four roundings and every
 *     bounds test are done once per thread; the flow comes as one packed (x | y << 16) word per cell
 *     (written by the search kernel's blur tail next to the planar array of the C interface);
 *   - a source run = 4 consecutive samples at an arbitrary displacement = two aligned 32-bit loads and a
 *     funnel shift (chroma with an odd displacement: warpFrameKernel.cl:171 picks U from column c+d-1 and
 *     V from c+d+1: six samples, one byte permute);
 *   - conversions through the 2^23 magic number, with the subtractions folded into the arithmetic where
 *     that is exact:  fl(b * t) == fma(2^23 + b, t, -(2^23 * t))   (one rounding of the exact product),
 *                     v - sub   == (2^23 + v) - (2^23 + sub)       (integers below 2^24),
 *     all of it two samples per instruction on the packed fp32 pipe (FADD2 / FMUL2 / FFMA2);
 *   - 128-thread CTAs (4 row groups x 128 samples), about ten per SM at 1080p, so that the block
 *     scheduler evens out the border CTAs.
 * The float expressions are those of hr_warp.cuh (header there): bit-identical to the reference kernel
 * as the NVIDIA OpenCL compiler builds it.
 */
#pragma once
#include "hr_warp.cuh"
#include <type_traits>

struct WarpFastArgs {
    int lumaGroups, lumaG0, chromaG0, chromaGN; /* row groups of the launch: luma first, then chroma       */
    /* level map, per plane kind [0] luma [1] chroma. 8-bit: sub = black / 128, rcp = MUFU.RCP of den (read back
     * from the device by the host once per knob setting). 16-bit: sub = b16 / 32768, rcp = correctly rounded
     * reciprocal (host). */
    float sub[2], den[2], rcp[2];
    int clampNeeded[2];    /* the map can leave [0, max]: clamp in float before the truncation             */
    int subIsInt[2];       /* sub is an integer below 2^22: (2^23 + v) - (2^23 + sub) is exact              */
};

__device__ __forceinline__ int round_half_away(float x) {
    /* round() of the reference as compiled: trunc(x + copysign(0.5, x)), the add rounded toward zero */
    return __float2int_rz(__fadd_rz(x, copysignf(0.5f, x)));
}

/* ---- source runs ----------------------------------------------------------------------------------- */
template <typename T>
struct RunSrc; /* where a thread's source block starts: aligned word pointer, bit shift, chroma-odd flag */
template <>
struct RunSrc<uint8_t> {
    const uint32_t *q;
    unsigned sh;
    bool more, odd;
    int rowWords;
    __device__ __forceinline__ void set(const uint8_t *plane, int o, int W, bool oddDisp) {
        odd = oddDisp;
        if (odd) o -= 1;
        q = reinterpret_cast<const uint32_t *>(plane) + (o >> 2);
        sh = (unsigned)(o & 3) * 8;
        more = sh != 0 || odd; /* an odd chroma displacement reads six samples */
        rowWords = W >> 2;
    }
    template <bool CHROMA>
    __device__ __forceinline__ uint32_t row(int r) const {
        const uint32_t *p = q + r * rowWords;
        const uint32_t w0 = __ldg(p), w1 = more ? __ldg(p + 1) : 0u;
        const uint32_t lo = __funnelshift_r(w0, w1, sh);
        if (!CHROMA) return lo;
        const uint32_t w2 = (odd && sh == 24) ? __ldg(p + 2) : 0u;
        const uint32_t hi = __funnelshift_r(w1, w2, sh);
        return __byte_perm(lo, hi, odd ? 0x5230u : 0x3210u);
    }
};
template <>
struct RunSrc<uint16_t> {
    const uint32_t *q;
    unsigned sh;
    bool more, odd;
    int rowWords;
    __device__ __forceinline__ void set(const uint16_t *plane, int o, int W, bool oddDisp) {
        odd = oddDisp;
        if (odd) o -= 1;
        q = reinterpret_cast<const uint32_t *>(plane) + (o >> 1);
        sh = (unsigned)(o & 1) * 16;
        more = sh != 0;
        rowWords = W >> 1;
    }
    template <bool CHROMA>
    __device__ __forceinline__ uint2 row(int r) const {
        const uint32_t *p = q + r * rowWords;
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1);
        if (!CHROMA) {
            const uint32_t w2 = more ? __ldg(p + 2) : 0u;
            return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
        }
        const uint32_t w2 = (more || odd) ? __ldg(p + 2) : 0u;
        const uint32_t w3 = (more && odd) ? __ldg(p + 3) : 0u;
        const uint32_t s0 = __funnelshift_r(w0, w1, sh), s1 = __funnelshift_r(w1, w2, sh), s2 = __funnelshift_r(w2, w3, sh);
        /* even displacement: samples 0,1 | 2,3; odd: 0,3 | 2,5 */
        return odd ? make_uint2(__byte_perm(s0, s1, 0x7610), __byte_perm(s1, s2, 0x7610)) : make_uint2(s0, s1);
    }
};

/* ---- blend + levels, two samples per instruction ----------------------------------------------------- */
struct BlendK {
    F2 t12, t21, negM, negMt12, M, rcp, mul, add, negSub, negMsub;
    float lo, hi;
};
template <bool IS16>
__device__ __forceinline__ BlendK make_blendk(float t12, float t21, const WarpFastArgs &A, int cz) {
    BlendK K;
    const float sub = A.sub[cz];
    const float rcp = A.rcp[cz];
    const float mul = IS16 ? 65472.0f : 255.0f, add = IS16 ? 32768.0f : 128.0f;
    K.t12 = f2_make(t12, t12);
    K.t21 = f2_make(t21, t21);
    K.negM = f2_make(-HR_MAGIC, -HR_MAGIC);
    K.negMt12 = f2_make(-(HR_MAGIC * t12), -(HR_MAGIC * t12));
    K.M = f2_make(HR_MAGIC, HR_MAGIC);
    K.rcp = f2_make(rcp, rcp);
    K.mul = f2_make(mul, mul);
    K.add = f2_make(add, add);
    K.negSub = f2_make(-sub, -sub);
    K.negMsub = f2_make(-(HR_MAGIC + sub), -(HR_MAGIC + sub));
    K.lo = 0.0f;
    K.hi = mul;
    return K;
}
/* one pair: A, B = the bit patterns 0x4B000000 | sample of the frame-1 / frame-2 samples; returns the bit
 * patterns 0x4B000000 | output of the pair */
template <bool CHROMA, bool CLAMP, bool SUBINT>
__device__ __forceinline__ F2 blend_pair(const BlendK &K, F2 A, F2 B) {
    const F2 a = f2_add(A, K.negM);                         /* (float)f1                                  */
    const F2 p = f2_fma(B, K.t12, K.negMt12);               /* (float)f2 * t12, one rounding               */
    const F2 bl = f2_fma(a, K.t21, p);                      /* fma(f1, t21, f2 * t12)                      */
    const F2 vM = f2_add_rz(bl, K.M);                       /* 2^23 + trunc(blend)                         */
    const F2 d = SUBINT ? f2_add(vM, K.negMsub) : f2_add(f2_add(vM, K.negM), K.negSub); /* v - sub            */
    F2 x = f2_mul(d, K.rcp);
    x = CHROMA ? f2_fma(x, K.mul, K.add) : f2_mul(x, K.mul);
    if (CLAMP) x = f2_make(fmaxf(fminf(f2_lo(x), K.hi), K.lo), fmaxf(fminf(f2_hi(x), K.hi), K.lo));
    return f2_add_rz(x, K.M);
}
__device__ __forceinline__ uint32_t f2_lo_bits(F2 v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t f2_hi_bits(F2 v) { return (uint32_t)(v >> 32); }

template <bool CHROMA, bool CLAMP, bool SUBINT>
__device__ __forceinline__ uint32_t blend_run(const BlendK &K, uint32_t wa, uint32_t wb) {
    const F2 r0 = blend_pair<CHROMA, CLAMP, SUBINT>(K, f2_bits(__byte_perm(wa, 0x4B000000u, 0x7540u), __byte_perm(wa, 0x4B000000u, 0x7541u)),
                                                    f2_bits(__byte_perm(wb, 0x4B000000u, 0x7540u), __byte_perm(wb, 0x4B000000u, 0x7541u)));
    const F2 r1 = blend_pair<CHROMA, CLAMP, SUBINT>(K, f2_bits(__byte_perm(wa, 0x4B000000u, 0x7542u), __byte_perm(wa, 0x4B000000u, 0x7543u)),
                                                    f2_bits(__byte_perm(wb, 0x4B000000u, 0x7542u), __byte_perm(wb, 0x4B000000u, 0x7543u)));
    return __byte_perm(__byte_perm(f2_lo_bits(r0), f2_hi_bits(r0), 0x0040), __byte_perm(f2_lo_bits(r1), f2_hi_bits(r1), 0x0040), 0x5410);
}
/* P010: the low 16 bits of each result hold trunc(x) < 65536; clamp to 65472, round to the nearest 10-bit
 * code ((v + 32) & 0xFFC0) on both halves at once */
template <bool CHROMA, bool CLAMP, bool SUBINT>
__device__ __forceinline__ uint2 blend_run(const BlendK &K, uint2 wa, uint2 wb) {
    const F2 r0 = blend_pair<CHROMA, CLAMP, SUBINT>(K, f2_bits(__byte_perm(wa.x, 0x4B000000u, 0x7510u), __byte_perm(wa.x, 0x4B000000u, 0x7532u)),
                                                    f2_bits(__byte_perm(wb.x, 0x4B000000u, 0x7510u), __byte_perm(wb.x, 0x4B000000u, 0x7532u)));
    const F2 r1 = blend_pair<CHROMA, CLAMP, SUBINT>(K, f2_bits(__byte_perm(wa.y, 0x4B000000u, 0x7510u), __byte_perm(wa.y, 0x4B000000u, 0x7532u)),
                                                    f2_bits(__byte_perm(wb.y, 0x4B000000u, 0x7510u), __byte_perm(wb.y, 0x4B000000u, 0x7532u)));
    const uint32_t p0 = __vminu2(__byte_perm(f2_lo_bits(r0), f2_hi_bits(r0), 0x5410), 0xFFC0FFC0u);
    const uint32_t p1 = __vminu2(__byte_perm(f2_lo_bits(r1), f2_hi_bits(r1), 0x5410), 0xFFC0FFC0u);
    return make_uint2((p0 + 0x00200020u) & 0xFFC0FFC0u, (p1 + 0x00200020u) & 0xFFC0FFC0u);
}

__device__ __forceinline__ void store_run(uint8_t *p, uint32_t v) { *reinterpret_cast<uint32_t *>(p) = v; }
__device__ __forceinline__ void store_run(uint16_t *p, uint2 v) { *reinterpret_cast<uint2 *>(p) = v; }
__device__ __forceinline__ uint32_t load_own(const uint8_t *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }
__device__ __forceinline__ uint2 load_own(const uint16_t *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }

/* rare threads: a partial column / row group, or the column that straddles the middle in mode 5 */
template <typename T>
__device__ __noinline__ void warp_thread_slow(const WarpParams<T> &P, int cx0, int cy0, int cz, int nrows) {
    T *out = cz ? P.outUV : P.outY;
    for (int r = 0; r < nrows; ++r)
        for (int k = 0; k < 4; ++k)
            if (cx0 + k < P.aW) out[(size_t)(cy0 + r) * P.W + cx0 + k] = (T)warp_sample(P, cx0 + k, cy0 + r, cz);
}

template <typename T, int ROWS, bool CHROMA>
__device__ __forceinline__ void warp_fast_thread(const WarpParams<T> &P, const WarpFastArgs &A, int cx0, int cy0) {
    constexpr bool is16 = SampleTraits<T>::is16;
    typedef typename RunType<T>::type Run;
    constexpr int cz = CHROMA ? 1 : 0;
    const int planeH = CHROMA ? (P.H >> 1) : P.H;
    const T *s12 = CHROMA ? P.f1uv : P.f1y;
    const T *s21 = CHROMA ? P.f2uv : P.f2y;
    T *po = (CHROMA ? P.outUV : P.outY) + cy0 * P.W + cx0;
    const int mode = P.mode;

    if (cx0 + 3 >= P.aW || cy0 + ROWS > planeH) {
        warp_thread_slow(P, cx0, cy0, cz, hr_min(ROWS, planeH - cy0));
        return;
    }
    if (mode == 5) {
        const int half = P.aW >> 1;
        if (cx0 + 3 < half) { /* left half of SideBySide1: frame 1 as it is (warpFrameKernel.cl:131-133) */
#pragma unroll
            for (int r = 0; r < ROWS; ++r) store_run(po + r * P.W, load_own(s12 + (cy0 + r) * P.W + cx0));
            return;
        }
        if (cx0 < half) {
            warp_thread_slow(P, cx0, cy0, cz, ROWS);
            return;
        }
    }

    /* the cell's flow and its flip (warpFrameKernel.cl:151-156) */
    const int s = P.s;
    int lx = cx0 >> s, ly = cy0 >> s;
    if (CHROMA) {
        lx &= ~1;
        ly <<= 1;
    }
    const uint32_t w12 = __ldg(P.flowXY + ly * P.lw + lx);
    const int x12 = (int)(int16_t)w12, y12 = (int)w12 >> 16;
    const int fy = hr_min(hr_max(ly - (y12 >> s), 0), P.lh - 1);
    const int fx = hr_min(hr_max(lx - (x12 >> s), 0), P.lw - 1);
    const uint32_t w21 = __ldg(P.flowXY + fy * P.lw + fx);
    const int x21 = (int)(int16_t)w21, y21 = (int)w21 >> 16;

    /* displacements (warpFrameKernel.cl:165-168) */
    float fe12 = (float)y12 * P.t12, fe21 = (float)y21 * P.t21;
    if (CHROMA) {
        fe12 *= 0.5f;
        fe21 *= 0.5f;
    }
    const int d12 = round_half_away((float)x12 * P.t12), d21 = -round_half_away((float)x21 * P.t21);
    const int e12 = round_half_away(fe12), e21 = -round_half_away(fe21);
    const int a12 = cx0 + d12, a21 = cx0 + d21, b12 = cy0 + e12, b21 = cy0 + e21;
    /* all four source blocks inside [1, aW-2] x [1, planeH-2]: the mirror/clamp of warpFrameKernel.cl:10-18 is
     * the identity, rows and columns are consecutive */
    const bool interior = (unsigned)(a12 - 1) <= (unsigned)(P.aW - 6) && (unsigned)(a21 - 1) <= (unsigned)(P.aW - 6) &&
                          (unsigned)(b12 - 1) <= (unsigned)(planeH - ROWS - 2) && (unsigned)(b21 - 1) <= (unsigned)(planeH - ROWS - 2);
    const int var = mode < 2 ? 0 : (A.subIsInt[cz] ? (A.clampNeeded[cz] ? 2 : 1) : 3);
    const BlendK K = make_blendk<is16>(P.t12, P.t21, A, cz);
    /* one output row from its two source runs. VAR 0: WarpedFrame12 / WarpedFrame21, the sample as it is
     * (warpFrameKernel.cl:170-173); 1..3: blend + levels (no clamp / clamp / clamp and non-integer black) */
    auto emit = [&](auto varTag, int r, Run a, Run b) {
        constexpr int VAR = decltype(varTag)::value;
        if (VAR == 0) store_run(po + r * P.W, mode == 0 ? a : b);
        else store_run(po + r * P.W, blend_run<CHROMA, VAR >= 2, VAR != 3>(K, a, b));
    };
    auto rows = [&](auto varTag) {
        constexpr int VAR = decltype(varTag)::value;
        if (interior) {
            Run ra[ROWS], rb[ROWS];
            RunSrc<T> A12, A21;
            if (VAR != 0 || mode == 0) {
                A12.set(s12, b12 * P.W + a12, P.W, CHROMA && (d12 & 1));
#pragma unroll
                for (int r = 0; r < ROWS; ++r) ra[r] = A12.template row<CHROMA>(r);
            }
            if (VAR != 0 || mode == 1) {
                A21.set(s21, b21 * P.W + a21, P.W, CHROMA && (d21 & 1));
#pragma unroll
                for (int r = 0; r < ROWS; ++r) rb[r] = A21.template row<CHROMA>(r);
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) emit(varTag, r, ra[r], rb[r]);
        } else {
            /* a frame border is involved: every sample through the mirror + clamp, chroma through the pair rule */
#pragma unroll 1
            for (int r = 0; r < ROWS; ++r) {
                Run a = Run(), b = Run();
                if (VAR != 0 || mode == 0) a = load_run4_border(s12 + (size_t)warp_mirror(b12 + r, planeH) * P.W, cx0, d12, P.aW, cz);
                if (VAR != 0 || mode == 1) b = load_run4_border(s21 + (size_t)warp_mirror(b21 + r, planeH) * P.W, cx0, d21, P.aW, cz);
                emit(varTag, r, a, b);
            }
        }
    };
    switch (var) {
        case 0: rows(std::integral_constant<int, 0>()); break;
        case 1: rows(std::integral_constant<int, 1>()); break;
        case 2: rows(std::integral_constant<int, 2>()); break;
        default: rows(std::integral_constant<int, 3>()); break;
    }
}

/* grid: x = 128-sample column blocks, y = groups of 4 row groups; row groups of the luma plane first */
template <typename T, int ROWS>
/* twelve 128-thread CTAs per SM (40 registers): the kernel is latency-bound at 4K and above, resident warps are what
 * hides its three dependent round trips (tools/diag_launch.py, 8K P010: 80 us at 9 CTAs of 8-row units, 74 us at 10,
 * 67 us with 4-row units at 10-12) */
#ifndef HR_WARP_MINBLOCKS
#define HR_WARP_MINBLOCKS 12
#endif
__global__ void __launch_bounds__(128, HR_WARP_MINBLOCKS) warp_fast_kernel(const __grid_constant__ WarpParams<T> P, const __grid_constant__ WarpFastArgs A) {
    const int cx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int rg = blockIdx.y * 4 + threadIdx.y;
    if (cx0 >= P.aW) return;
    if (rg < A.lumaGroups) {
        warp_fast_thread<T, ROWS, false>(P, A, cx0, (A.lumaG0 + rg) * ROWS);
    } else if (rg - A.lumaGroups < A.chromaGN) {
        const int cy0 = (A.chromaG0 + rg - A.lumaGroups) * ROWS;
        if (cy0 < (P.H >> 1)) warp_fast_thread<T, ROWS, true>(P, A, cx0, cy0);
    }
}

/* packed copy of a planar blurred flow (parity tap hr_set_blurred_offsets) */
__global__ void pack_flow_kernel(const int16_t *__restrict__ flow, uint32_t *__restrict__ flowXY, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flowXY[i] = (uint32_t)(uint16_t)flow[i] | ((uint32_t)(uint16_t)flow[n + i] << 16);
}

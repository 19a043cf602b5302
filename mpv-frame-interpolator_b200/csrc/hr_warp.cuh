/*
 * hr_warp.cuh — K5: flip lookup + bidirectional warp + time-weighted blend + output levels + the
 * seven output modes, luma and chroma planes in ONE launch (warpFrameKernel.cl:114-182, two
 * launches in the reference, opticalFlowCalc.c:229-232).
 *
 * This file: the arithmetic of one output sample (warp_sample, every mode and geometry), the generic
 * kernel that applies it sample by sample (modes 3/4/6, frames of 540 lines or less, unaligned planes,
 * degenerate level knobs) and the helpers shared with the fast kernel of hr_warp_fast.cuh.
 *
 * Arithmetic: the float expressions are evaluated exactly as the reference kernel executes them on
 * an NVIDIA OpenCL device (PTX of the unmodified .cl source, tools/dump_ref_ptx.py):
 *     blend   = fma(f1, 1-t, f2 * t)                       (a*s21 + b*s12 contracted once)
 *     luma    = ((v - black) * rcp(white - black)) * 255    ('/' is div.full.f32 = MUFU.RCP + FMUL)
 *     chroma  = fma((v - 128) * rcp(white), 255, 128)
 *     round() = trunc(x + copysign(0.5, x)) with the add rounded toward zero (== roundf)
 * so 8-bit output is bit-identical to the reference run on the same GPU (tests/
 * test_gpu_vs_reference_opencl.py); the IEEE/no-contraction reading of the source differs from it
 * by at most +-1 LSB per operation. The library is compiled with -fmad=false, every fma below is
 * explicit.
 */
#pragma once
#include "hr_common.cuh"

/* warpFrameKernel.cl:10-18 */
__device__ __forceinline__ int warp_mirror(int pos, int dim) {
    int res = pos;
    if (pos >= dim - 1) res = pos - ((pos - (dim - 2)) * 2);
    else if (pos < 1) res = -pos + 1;
    return hr_min(hr_max(res, 1), dim - 2);
}
/* '/' of the reference as compiled for NVIDIA OpenCL devices: div.full.f32 */
__device__ __forceinline__ float div_full(float a, float b) {
    float r;
    asm("div.full.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float rcp_approx(float b) {
    float r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(b));
    return r;
}
/* div.full.f32 is MUFU.RCP(b) * a while |b| is inside this range (outside it pre-scales both) */
__device__ __forceinline__ bool div_full_is_rcp_mul(float b) { return fabsf(b) >= 1.17549435e-38f && fabsf(b) <= 8.50705917e37f; }

__device__ __forceinline__ unsigned sat_u8(float v) { return __float2uint_rz(fmaxf(fminf(v, 255.0f), 0.0f)); }
/* (unsigned char)(float) of OpenCL C as NVIDIA compiles it: cvt.rzi.u16.f32, low byte stored */
__device__ __forceinline__ unsigned cvt_uchar(float v) {
    unsigned short r;
    asm("cvt.rzi.u16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return (unsigned)r & 255u;
}

/* warpFrameKernel.cl:1-7 */
__device__ __forceinline__ unsigned levels_y8(float v, float black, float white) { return sat_u8(div_full(v - black, white - black) * 255.0f); }
__device__ __forceinline__ unsigned levels_uv8(float v, float white) { return sat_u8(__fmaf_rn(div_full(v - 128.0f, white), 255.0f, 128.0f)); }

/* P010 output levels, defined by construction: the 8-bit knobs are mapped onto the
 * MSB-aligned 10-bit range (65472 = 1023 << 6), the division is a multiplication by the correctly
 * rounded reciprocal, and the result is rounded to the nearest 10-bit code (so that the default
 * levels are an exact identity). */
struct Levels16 {
    float b16, rY, rUV;
};
__device__ __forceinline__ Levels16 make_levels16(float black, float white) {
    Levels16 L;
    L.b16 = __fmul_rn(__fdiv_rn(black, 255.0f), 65472.0f);
    const float w16 = __fmul_rn(__fdiv_rn(white, 255.0f), 65472.0f);
    L.rY = __fdiv_rn(1.0f, w16 - L.b16);
    L.rUV = __fdiv_rn(1.0f, w16);
    return L;
}
__device__ __forceinline__ unsigned levels_y16(float v, const Levels16 &L) {
    const float x = fmaxf(fminf((v - L.b16) * L.rY * 65472.0f, 65472.0f), 0.0f);
    return (__float2uint_rz(x) + 32u) & 0xFFC0u;
}
__device__ __forceinline__ unsigned levels_uv16(float v, const Levels16 &L) {
    const float x = fmaxf(fminf(__fmaf_rn((v - 32768.0f) * L.rUV, 65472.0f, 32768.0f), 65472.0f), 0.0f);
    return (__float2uint_rz(x) + 32u) & 0xFFC0u;
}

/* warpFrameKernel.cl:21-111. The hue comes from CUDA's atan2f/fmodf (the OpenCL built-ins are a
 * different polynomial: +-1 on the 8-bit channels, tests allow it); the scaling and the RGB->YUV
 * products follow the reference's contraction pattern. */
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
        const float hue = div_full(angle_deg, 360.0f);
        const int h_i = (int)(hue * 6.0f);
        const float f = __fmaf_rn(hue, 6.0f, -(float)h_i);
        const float q = 1.0f - f;
        switch (h_i % 6) {
            case 0: r = 255; g = cvt_uchar(f * 255.0f); b = 0; break;
            case 1: r = cvt_uchar(q * 255.0f); g = 255; b = 0; break;
            case 2: r = 0; g = 255; b = cvt_uchar(f * 255.0f); break;
            case 3: r = 0; g = cvt_uchar(q * 255.0f); b = 255; break;
            case 4: r = cvt_uchar(f * 255.0f); g = 0; b = 255; break;
            case 5: r = 255; g = 0; b = cvt_uchar(q * 255.0f); break;
            default: r = g = b = 0; break;
        }
        const float gq = div_full((float)g, 255.0f) * (float)ay;
        r = sat_u8(div_full((float)r, 255.0f) * (float)(ax + ay) * (float)resImpact);
        g = sat_u8(__fmaf_rn(div_full((float)g, 255.0f), (float)ay, gq) * (float)resImpact);
        b = sat_u8(div_full((float)b, 255.0f) * (float)(ax + ay) * (float)resImpact);
    }
    const float fr = (float)r, fg = (float)g, fb = (float)b;
    if (channel == 0) return ((sat_u8(__fmaf_rn(fb, 0.114f, __fmaf_rn(fr, 0.299f, fg * 0.587f))) >> 1) + (currPixel >> 1)) & 255u;
    if (channel == 1) return sat_u8(__fmaf_rn(fb, 0.5f, __fmaf_rn(fr, -0.168736f, fg * -0.331264f)) + 128.0f);
    return sat_u8(__fmaf_rn(fb, -0.081312f, __fmaf_rn(fr, 0.5f, fg * -0.418688f)) + 128.0f);
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
    const int ln = P.lw * P.lh;
    const int i12 = scy * P.lw + scx;
    CellFlow f;
    f.x12 = __ldg(P.flow + i12);
    f.y12 = __ldg(P.flow + ln + i12);
    const int fy = hr_min(hr_max(scy - (f.y12 >> s), 0), P.lh - 1);
    const int fx = hr_min(hr_max(scx - (f.x12 >> s), 0), P.lw - 1);
    const int i21 = fy * P.lw + fx;
    f.x21 = __ldg(P.flow + i21);
    f.y21 = __ldg(P.flow + ln + i21);
    return f;
}

/* blend + mode 3 + levels of one sample pair (warpFrameKernel.cl:175-180) */
template <typename T>
__device__ __forceinline__ unsigned finish_blend(const WarpParams<T> &P, unsigned a, unsigned b, int cz, int cx, const CellFlow &f) {
    const float bl = __fmaf_rn((float)a, P.t21, (float)b * P.t12);
    if (!SampleTraits<T>::is16) {
        unsigned v = cvt_uchar(bl);
        if (P.mode == 3) v = visualize_flow(-f.x12, -f.y12, v, cz + (cx & (cz ? 1 : 0)), P.s <= 2 ? 4 : 1);
        return cz ? levels_uv8((float)v, P.white) : levels_y8((float)v, P.black, P.white);
    } else {
        const unsigned v = __float2uint_rz(fmaxf(fminf(bl, 65535.0f), 0.0f));
        if (P.mode == 3) {
            const unsigned v8 = visualize_flow(-f.x12, -f.y12, v >> 8, cz + (cx & (cz ? 1 : 0)), P.s <= 2 ? 4 : 1);
            const unsigned l8 = cz ? levels_uv8((float)v8, P.white) : levels_y8((float)v8, P.black, P.white);
            return l8 << 8;
        }
        const Levels16 L = make_levels16(P.black, P.white);
        return cz ? levels_uv16((float)v, L) : levels_y16((float)v, L);
    }
}

/* One output sample, every mode: frame borders, modes 3/4/6, small frames (warpFrameKernel.cl:119-181). */
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

/* four samples at a frame border: each through the mirror + clamp of warpFrameKernel.cl:10-18 and,
 * for chroma, the pair alignment of :171 */
__device__ __forceinline__ uint32_t load_run4_border(const uint8_t *row, int cx0, int d, int aW, int cz) {
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int nx = warp_mirror(cx0 + k + d, aW);
        v |= (uint32_t)__ldg(row + (cz ? (nx & ~1) + (k & 1) : nx)) << (8 * k);
    }
    return v;
}
__device__ __forceinline__ uint2 load_run4_border(const uint16_t *row, int cx0, int d, int aW, int cz) {
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int nx = warp_mirror(cx0 + k + d, aW);
        v[k] = __ldg(row + (cz ? (nx & ~1) + (k & 1) : nx));
    }
    return make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
}

template <typename T>
struct RunType;
template <>
struct RunType<uint8_t> {
    typedef uint32_t type;
};
template <>
struct RunType<uint16_t> {
    typedef uint2 type;
};

/* two fp32 values in one 64-bit register pair: sm_100 executes add/mul/fma on both per instruction
 * (FADD2 / FMUL2 / FFMA2), each half rounded exactly like the scalar instruction */
typedef unsigned long long F2;
__device__ __forceinline__ F2 f2_make(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ F2 f2_bits(uint32_t lo, uint32_t hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ float f2_lo(F2 v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo;
}
__device__ __forceinline__ float f2_hi(F2 v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return hi;
}
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) {
    F2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ F2 f2_add_rz(F2 a, F2 b) {
    F2 r;
    asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) {
    F2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) {
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

#define HR_MAGIC 8388608.0f /* 2^23: u8 / u16 <-> f32 through the mantissa (exact for 0 <= v < 2^23) */

/* Every output sample through warp_sample(): thread = 4 samples x 4 rows; row groups of the luma plane first
 * (lumaGroups of them from group lumaG0), then chromaGN groups of the chroma plane from chromaG0 (whole frame:
 * 0, all, 0, all; a spatial band: its rows only). */
template <typename T>
__global__ void __launch_bounds__(256) warp_generic_kernel(const WarpParams<T> P, int lumaGroups, int lumaG0, int chromaG0, int chromaGN) {
    constexpr int ROWS = 4;
    const int cx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int rg = blockIdx.y * blockDim.y + threadIdx.y;
    const int cz = rg >= lumaGroups;
    if (cz && rg - lumaGroups >= chromaGN) return;
    const int cy0 = (cz ? chromaG0 + rg - lumaGroups : lumaG0 + rg) * ROWS;
    const int planeH = cz ? (P.H >> 1) : P.H;
    if (cx0 >= P.aW || cy0 >= planeH) return;
    T *out = cz ? P.outUV : P.outY;
    const int nrows = hr_min(ROWS, planeH - cy0);
    for (int r = 0; r < nrows; ++r) {
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const int cx = cx0 + k;
            if (cx < P.aW) out[(size_t)(cy0 + r) * P.W + cx] = (T)warp_sample(P, cx, cy0 + r, cz);
        }
    }
}

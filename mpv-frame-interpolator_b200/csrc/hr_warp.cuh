/*
 * hr_warp.cuh — K5 (output side of the hot path): flow flip, bidirectional warp, time-weighted blend, output levels
 * and the seven output modes; luma and chroma planes in ONE launch (the reference launches its kernel twice,
 * opticalFlowCalc.c:229-232; behaviour defined by warpFrameKernel.cl:1-182).
 *
 * This file holds what every warp kernel shares:
 *   - the float arithmetic of one output sample, evaluated exactly as the reference kernel executes on an NVIDIA
 *     OpenCL device (PTX of the unmodified .cl source, tools/dump_ref_ptx.py):
 *         blend   = fma(f1, 1-t, f2 * t)                       (a*s21 + b*s12 contracted once)
 *         luma    = ((v - black) * rcp(white - black)) * 255    ('/' is div.full.f32 = MUFU.RCP + FMUL)
 *         chroma  = fma((v - 128) * rcp(white), 255, 128)
 *         round() = trunc(x + copysign(0.5, x)) with the add rounded toward zero (== roundf)
 *     so 8-bit output is bit-identical to the reference run on the same GPU (tests/test_gpu_vs_reference_opencl.py).
 *     The library is compiled with -fmad=false; every fma below is written out.
 *   - the flow colours of the HSV output mode as a per-cell table (the flow is constant over a lattice cell, so
 *     the colour is too: flow_colour_kernel computes it once per flow, the warp kernels only add the luma term),
 *   - the per-sample path (sample_value): every mode and geometry, one sample at a time. It serves partial thread
 *     units and seams of the fast kernel (hr_warp_fast.cuh) and — through warp_generic_kernel — frames of 540 lines
 *     or less, unaligned planes, degenerate level knobs and the half-size side-by-side mode.
 */
#pragma once
#include "hr_common.cuh"

/* ---- arithmetic primitives ------------------------------------------------------------------------------ */
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
__device__ __forceinline__ unsigned sat_u8(float v) { return __float2uint_rz(fmaxf(fminf(v, 255.0f), 0.0f)); }
/* (unsigned char)(float) of OpenCL C as NVIDIA compiles it: cvt.rzi.u16.f32, low byte stored */
__device__ __forceinline__ unsigned cvt_uchar(float v) {
    unsigned short r;
    asm("cvt.rzi.u16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return (unsigned)r & 255u;
}
__device__ __forceinline__ int round_half_away(float x) {
    /* round() of the reference as compiled: trunc(x + copysign(0.5, x)), the add rounded toward zero */
    return __float2int_rz(__fadd_rz(x, copysignf(0.5f, x)));
}

/* Reflection of a displaced coordinate back into the picture: about 1 on the low side, about dim - 2 on the high
 * side, then held inside [1, dim - 2] (what warpFrameKernel.cl:10-18 computes; an undisplaced column 0 therefore
 * reads column 1 and column dim - 1 reads dim - 3). */
__device__ __forceinline__ int reflect_inner(int p, int dim) {
    const int hiEdge = dim - 2;
    if (p > hiEdge) p = 2 * hiEdge - p;
    else if (p < 1) p = 1 - p;
    return hr_min(hr_max(p, 1), hiEdge);
}

/* 8-bit output levels (warpFrameKernel.cl:1-7) */
__device__ __forceinline__ unsigned levels_y8(float v, float black, float white) { return sat_u8(div_full(v - black, white - black) * 255.0f); }
__device__ __forceinline__ unsigned levels_uv8(float v, float white) { return sat_u8(__fmaf_rn(div_full(v - 128.0f, white), 255.0f, 128.0f)); }

/* P010 output levels, defined by construction: the 8-bit knobs are mapped onto the MSB-aligned
 * 10-bit range (65472 = 1023 << 6), the division is a multiplication by the correctly rounded reciprocal, and the
 * result is rounded to the nearest 10-bit code (so that the default levels are an exact identity). */
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

/* ---- HSV flow mode: the colour of a flow vector ----------------------------------------------------------
 * A vector's direction picks a fully saturated hue, its length the brightness (warpFrameKernel.cl:21-111 on the
 * negated vector). The result does not depend on the picture except for half of the luma, so it is a property of
 * the lattice cell: one table word per cell,
 *     byte 0 = luma share of the colour, already halved    byte 1 = U    byte 2 = V    (before output levels).
 * The hue comes from CUDA's atan2f / fmodf (OpenCL's built-ins are a different polynomial: a unit step on an 8-bit
 * channel now and then, the tests allow it); every other operation follows the reference's compiled arithmetic. */
__device__ __forceinline__ uint32_t flow_colour(int vx, int vy, int gain) {
    const int mx = abs(vx), my = abs(vy);
    float rgb[3] = {0.0f, 0.0f, 0.0f};
    if (mx >= 1 || my >= 1) {
        float deg = atan2f((float)vy, (float)vx) * (180.0f / 3.14159274101257f);
        if (deg < 0) deg += 360.0f;
        deg = fmodf(deg, 360.0f);
        if (deg < 0) deg += 360.0f;
        const float turn = div_full(deg, 360.0f);            /* hue as a fraction of the circle          */
        const int sixth = (int)(turn * 6.0f);
        const float rise = __fmaf_rn(turn, 6.0f, -(float)sixth), fall = 1.0f - rise;
        const int sector = sixth % 6;
        /* around the circle the channels take turns: one is full, its neighbour ramps up (even sectors) or down
         * (odd sectors), the third is off */
        const int full = ((sector + 1) >> 1) % 3, ramp = (7 - sector) % 3;
        const unsigned slope = cvt_uchar(((sector & 1) ? fall : rise) * 255.0f);
        unsigned ch[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) ch[c] = c == full ? 255u : (c == ramp ? slope : 0u);
        /* brightness: red and blue follow |x| + |y|, green twice |y| (g/255*|y| + g/255*|y| as one fma) */
        const float len = (float)(mx + my), k = (float)gain;
        const float gs = div_full((float)ch[1], 255.0f);
        rgb[0] = (float)sat_u8(div_full((float)ch[0], 255.0f) * len * k);
        rgb[1] = (float)sat_u8(__fmaf_rn(gs, (float)my, gs * (float)my) * k);
        rgb[2] = (float)sat_u8(div_full((float)ch[2], 255.0f) * len * k);
    }
    const unsigned y = sat_u8(__fmaf_rn(rgb[2], 0.114f, __fmaf_rn(rgb[0], 0.299f, rgb[1] * 0.587f)));
    const unsigned u = sat_u8(__fmaf_rn(rgb[2], 0.5f, __fmaf_rn(rgb[0], -0.168736f, rgb[1] * -0.331264f)) + 128.0f);
    const unsigned v = sat_u8(__fmaf_rn(rgb[2], -0.081312f, __fmaf_rn(rgb[0], 0.5f, rgb[1] * -0.418688f)) + 128.0f);
    return (y >> 1) | (u << 8) | (v << 16);
}
/* gain: the reference brightens the colours fourfold at resolution scalars <= 2 (warpFrameKernel.cl:178) */
__device__ __forceinline__ int colour_gain(int s) { return s <= 2 ? 4 : 1; }

/* one table word per lattice cell from the packed blurred flow (x | y << 16); the mode shows the vector negated */
__global__ void flow_colour_kernel(const uint32_t *__restrict__ flowXY, uint32_t *__restrict__ table, int n, int gain) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t w = __ldg(flowXY + i);
    table[i] = flow_colour((int)(int16_t)(-(int)(int16_t)w), (int)(int16_t)(-((int)w >> 16)), gain);
}

/* ---- the vectors of one output position ------------------------------------------------------------------ */
/* Lattice cell of a (possibly remapped) output position: luma (x >> s, y >> s); a chroma sample takes the vector of
 * the even lattice column / the doubled lattice row (warpFrameKernel.cl:151-152). */
__device__ __forceinline__ int2 lattice_cell(int x, int y, int s, bool chroma) {
    return chroma ? make_int2((x >> s) & ~1, (y >> s) << 1) : make_int2(x >> s, y >> s);
}
/* forward vector (frame 1 -> 2) at the cell, backward vector through the "flip": the vector stored where the forward
 * vector came from, cell - (forward >> s), held inside the lattice (warpFrameKernel.cl:153-156) */
struct VectorPair {
    int fx, fy, bx, by;
};
template <typename T>
__device__ __forceinline__ VectorPair vector_pair(const WarpParams<T> &P, int2 cell) {
    const uint32_t fw = __ldg(P.flowXY + cell.y * P.lw + cell.x);
    VectorPair v;
    v.fx = (int)(int16_t)fw;
    v.fy = (int)fw >> 16;
    const int ry = hr_min(hr_max(cell.y - (v.fy >> P.s), 0), P.lh - 1);
    const int rx = hr_min(hr_max(cell.x - (v.fx >> P.s), 0), P.lw - 1);
    const uint32_t bw = __ldg(P.flowXY + ry * P.lw + rx);
    v.bx = (int)(int16_t)bw;
    v.by = (int)bw >> 16;
    return v;
}
/* displacement of the sampling position for blend scalar t: forward into frame 1, backward into frame 2; chroma rows
 * move half as far (warpFrameKernel.cl:165-168; the product with 0.5 is a second rounding, kept) */
struct Shift {
    int x1, y1, x2, y2;
};
__device__ __forceinline__ Shift shift_for(const VectorPair &v, float t12, float t21, bool chroma) {
    float fy = (float)v.fy * t12, by = (float)v.by * t21;
    if (chroma) {
        fy *= 0.5f;
        by *= 0.5f;
    }
    Shift d;
    d.x1 = round_half_away((float)v.fx * t12);
    d.y1 = round_half_away(fy);
    d.x2 = -round_half_away((float)v.bx * t21);
    d.y2 = -round_half_away(by);
    return d;
}

/* ---- one output sample, every mode ------------------------------------------------------------------------ */
/* blend + levels of a sample pair in the 8-bit reading (the HSV mode works at 8 bits for P010 too) */
template <typename T>
__device__ __forceinline__ unsigned blend_finish(const WarpParams<T> &P, float t12, float t21, unsigned a, unsigned b, bool chroma, bool oddColumn, uint32_t colour) {
    const float mix = __fmaf_rn((float)a, t21, (float)b * t12);
    constexpr bool is16 = SampleTraits<T>::is16;
    if (P.mode == HR_MODE_HSV_FLOW) {
        const unsigned pic = is16 ? (__float2uint_rz(fmaxf(fminf(mix, 65535.0f), 0.0f)) >> 8) : cvt_uchar(mix);
        const unsigned v8 = !chroma ? ((colour & 255u) + (pic >> 1)) & 255u : (oddColumn ? (colour >> 16) & 255u : (colour >> 8) & 255u);
        const unsigned l8 = chroma ? levels_uv8((float)v8, P.white) : levels_y8((float)v8, P.black, P.white);
        return is16 ? l8 << 8 : l8;
    }
    if (!is16) {
        const unsigned v = cvt_uchar(mix);
        return chroma ? levels_uv8((float)v, P.white) : levels_y8((float)v, P.black, P.white);
    }
    const unsigned v = __float2uint_rz(fmaxf(fminf(mix, 65535.0f), 0.0f));
    const Levels16 L = make_levels16(P.black, P.white);
    return chroma ? levels_uv16((float)v, L) : levels_y16((float)v, L);
}

/* Value of output sample (cx, cy) of the luma (chroma = false) or interleaved chroma plane. Output modes
 * (enum FrameOutput, vf_HopperRender.c:21; geometry warpFrameKernel.cl:119-148):
 *   0 / 1  the displaced sample of frame 1 / frame 2 as it is      2  blend + levels      3  flow colours over the blend
 *   4      vector length as grey                                    5  left half: frame 1 untouched, right half: mode 2
 *   6      a half-height strip: frame 1 at half size on the left, the mode-2 picture at half size on the right */
template <typename T>
__device__ unsigned sample_value(const WarpParams<T> &P, float t12, float t21, int cx, int cy, bool chroma) {
    const T *from1 = chroma ? P.f1uv : P.f1y;
    const T *from2 = chroma ? P.f2uv : P.f2y;
    const int planeRows = chroma ? (P.H >> 1) : P.H;
    const bool oddColumn = chroma && (cx & 1);
    constexpr unsigned grey = SampleTraits<T>::is16 ? 32768u : 128u;
    int px = cx, py = cy; /* the picture position this output sample shows */

    if (P.mode == HR_MODE_SIDE_BY_SIDE_1 && cx < (P.aW >> 1)) return from1[(size_t)cy * P.W + cx];
    if (P.mode == HR_MODE_SIDE_BY_SIDE_2) {
        const int stripTop = (P.H >> 2) >> (chroma ? 1 : 0);
        const int row = cy - stripTop;
        if (row < 0 || row >= (planeRows >> 1)) return chroma ? grey : 0u; /* bars above and below the strip */
        if (cx < (P.W >> 1)) return from1[(size_t)(2 * row) * P.W + 2 * cx + (oddColumn ? 1 : 0)];
        px = 2 * (cx - (P.aW >> 1));
        py = 2 * row;
    }
    const int2 cell = lattice_cell(px, py, P.s, chroma);
    const VectorPair v = vector_pair(P, cell);
    if (P.mode == HR_MODE_GREY_FLOW) {
        const unsigned len4 = (unsigned)(abs(v.fx) + abs(v.fy)) << 2;
        const unsigned g8 = chroma ? 128u : (len4 < 255u ? len4 : 255u);
        return SampleTraits<T>::is16 ? g8 << 8 : g8;
    }
    const Shift d = shift_for(v, t12, t21, chroma);
    /* chroma keeps U on even and V on odd columns whatever the displacement (warpFrameKernel.cl:171) */
    const int pairMask = chroma ? ~1 : ~0, lane = oddColumn ? 1 : 0;
    const size_t at1 = (size_t)reflect_inner(py + d.y1, planeRows) * P.W + (reflect_inner(px + d.x1, P.aW) & pairMask) + lane;
    const size_t at2 = (size_t)reflect_inner(py + d.y2, planeRows) * P.W + (reflect_inner(px + d.x2, P.aW) & pairMask) + lane;
    if (P.mode == HR_MODE_WARPED_12) return from1[at1];
    if (P.mode == HR_MODE_WARPED_21) return from2[at2];
    const uint32_t colour = P.mode == HR_MODE_HSV_FLOW ? flow_colour(-v.fx, -v.fy, colour_gain(P.s)) : 0u;
    return blend_finish(P, t12, t21, (unsigned)from1[at1], (unsigned)from2[at2], chroma, oddColumn, colour);
}

/* Every output sample through sample_value(): thread = 4 samples x 4 rows; row groups of the luma plane first
 * (lumaGroups of them from group lumaG0), then chromaGN groups of the chroma plane from chromaG0 (whole frame:
 * 0, all, 0, all; a spatial band: its rows only). */
template <typename T>
__global__ void __launch_bounds__(256) warp_generic_kernel(const WarpParams<T> P, int lumaGroups, int lumaG0, int chromaG0, int chromaGN) {
    constexpr int ROWS = 4;
    const int cx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int rg = blockIdx.y * blockDim.y + threadIdx.y;
    const bool chroma = rg >= lumaGroups;
    if (chroma && rg - lumaGroups >= chromaGN) return;
    const int cy0 = (chroma ? chromaG0 + rg - lumaGroups : lumaG0 + rg) * ROWS;
    const int planeH = chroma ? (P.H >> 1) : P.H;
    if (cx0 >= P.aW || cy0 >= planeH) return;
    T *out = chroma ? P.outUV : P.outY;
    const int nrows = hr_min(ROWS, planeH - cy0);
    for (int r = 0; r < nrows; ++r) {
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const int cx = cx0 + k;
            if (cx < P.aW) out[(size_t)(cy0 + r) * P.W + cx] = (T)sample_value(P, P.t12, P.t21, cx, cy0 + r, chroma);
        }
    }
}

/* ---- packed fp32 pairs ------------------------------------------------------------------------------------
 * two fp32 values in one 64-bit register pair: sm_100 executes add/mul/fma on both per instruction
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

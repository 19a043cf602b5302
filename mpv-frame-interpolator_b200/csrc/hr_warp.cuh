/*
 * hr_warp.cuh — K5: flip lookup + bidirectional warp + time-weighted blend + output levels + the
 * seven output modes, luma and chroma planes in ONE launch (warpFrameKernel.cl:114-182, two
 * launches in the reference, opticalFlowCalc.c:229-232).
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
 *
 * Work decomposition (HBM-bound: 2 frames read + 1 frame written per launch):
 *   thread = 4 samples x 4 rows of one plane. For resolution scalars >= 2 (every frame higher than
 *   540 lines) the 4x4 block lies inside one lattice cell, so the two flow vectors (o12 at the cell,
 *   o21 through the flip indirection) and both displacements are computed ONCE per thread; each row
 *   is then two unaligned 4-sample source runs (two aligned 32/64-bit loads + funnel shift each),
 *   a 4-sample blend and one 32-bit (NV12) / 64-bit (P010) store; a warp stores 128 / 256
 *   contiguous bytes per row. u8<->f32 conversions go through the 2^23 magic number (PRMT + FADD,
 *   full-rate pipes) instead of I2F/F2I. Frame borders, modes 3/4/6, resolution scalars < 2 and
 *   out-of-range level denominators take the per-sample path warp_sample(), which computes the same
 *   numbers.
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

/* ---- the block path ------------------------------------------------------------------------------ */

/* four consecutive samples from an arbitrary (sample-aligned) address, built from aligned loads */
__device__ __forceinline__ uint32_t load_run4(const uint8_t *p) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 3) * 8;
    const uint32_t lo = __ldg(q);
    const uint32_t hi = sh ? __ldg(q + 1) : 0u;
    return __funnelshift_r(lo, hi, sh);
}
__device__ __forceinline__ uint2 load_run4(const uint16_t *p) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 2) * 8;
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1);
    const uint32_t w2 = sh ? __ldg(q + 2) : 0u;
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}
/* chroma with an odd displacement d: U samples come from column cx+d-1, V samples from cx+d+1
 * (warpFrameKernel.cl:171 `(newCx & ~1) + (cx & 1)`): six samples starting at cx0+d-1, picked 0,3,2,5 */
__device__ __forceinline__ uint32_t load_run4_uv_odd(const uint8_t *p /* = row + cx0 + d - 1 */) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 3) * 8;
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1);
    const uint32_t w2 = sh ? __ldg(q + 2) : 0u;
    return __byte_perm(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), 0x5230);
}
__device__ __forceinline__ uint2 load_run4_uv_odd(const uint16_t *p) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 2) * 8;
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
    const uint32_t w3 = sh ? __ldg(q + 3) : 0u;
    const uint32_t s0 = __funnelshift_r(w0, w1, sh), s1 = __funnelshift_r(w1, w2, sh), s2 = __funnelshift_r(w2, w3, sh);
    /* samples 0,3 | 2,5 */
    return make_uint2(__byte_perm(s0, s1, 0x7610), __byte_perm(s1, s2, 0x7610));
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

/* u8 / u16 -> f32 and back through the 2^23 magic number (exact for 0 <= v < 2^23) */
#define HR_MAGIC 8388608.0f
__device__ __forceinline__ float byte_to_float(uint32_t w, int k) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + (unsigned)k)) - HR_MAGIC;
}
__device__ __forceinline__ float half_to_float(uint32_t w, int k) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, k ? 0x7532u : 0x7510u)) - HR_MAGIC;
}
/* trunc(x) for 0 <= x < 2^23 as the bits 0x4B000000 | trunc(x) */
__device__ __forceinline__ uint32_t trunc_bits(float x) { return __float_as_uint(__fadd_rz(x, HR_MAGIC)); }

/* blend + levels of four 8-bit sample pairs (one 32-bit word each) -> four output bytes.
 * (uchar)(f1*s21 + f2*s12), then luma ((v - black) * rY) * 255 / chroma fma((v - 128) * rUV, 255, 128):
 * warpFrameKernel.cl:1-7,175-180 as compiled for NVIDIA OpenCL devices (header of this file). */
struct Blend8 {
    F2 t12, t21, negMagic, magic, nsub, rcp, k255, k128;
    bool chroma, clampNeeded;
};
__device__ __forceinline__ Blend8 make_blend8(float t12, float t21, float black, float white, int cz) {
    Blend8 B;
    const float sub = cz ? 128.0f : black, rcp = cz ? rcp_approx(white) : rcp_approx(white - black);
    B.t12 = f2_make(t12, t12);
    B.t21 = f2_make(t21, t21);
    B.negMagic = f2_make(-HR_MAGIC, -HR_MAGIC);
    B.magic = f2_make(HR_MAGIC, HR_MAGIC);
    B.nsub = f2_make(-sub, -sub);
    B.rcp = f2_make(rcp, rcp);
    B.k255 = f2_make(255.0f, 255.0f);
    B.k128 = f2_make(128.0f, 128.0f);
    B.chroma = cz != 0;
    /* the clamp to [0,255] is only needed when the level map can leave that range (it is monotonic) */
    const float xlo = cz ? __fmaf_rn((0.0f - sub) * rcp, 255.0f, 128.0f) : ((0.0f - sub) * rcp) * 255.0f;
    const float xhi = cz ? __fmaf_rn((255.0f - sub) * rcp, 255.0f, 128.0f) : ((255.0f - sub) * rcp) * 255.0f;
    B.clampNeeded = !(fminf(xlo, xhi) >= 0.0f && fmaxf(xlo, xhi) < 256.0f);
    return B;
}
template <bool CLAMP>
__device__ __forceinline__ uint32_t blend8_quad(const Blend8 &B, uint32_t wa, uint32_t wb) {
    uint32_t res[4];
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
        const F2 a = f2_add(f2_bits(__byte_perm(wa, 0x4B000000u, 0x7540u + k), __byte_perm(wa, 0x4B000000u, 0x7541u + k)), B.negMagic);
        const F2 b = f2_add(f2_bits(__byte_perm(wb, 0x4B000000u, 0x7540u + k), __byte_perm(wb, 0x4B000000u, 0x7541u + k)), B.negMagic);
        const F2 v = f2_add(f2_add_rz(f2_fma(a, B.t21, f2_mul(b, B.t12)), B.magic), B.negMagic);
        F2 x = f2_mul(f2_add(v, B.nsub), B.rcp);
        x = B.chroma ? f2_fma(x, B.k255, B.k128) : f2_mul(x, B.k255);
        if (CLAMP) x = f2_make(fmaxf(fminf(f2_lo(x), 255.0f), 0.0f), fmaxf(fminf(f2_hi(x), 255.0f), 0.0f));
        x = f2_add_rz(x, B.magic);
        res[k] = __float_as_uint(f2_lo(x));
        res[k + 1] = __float_as_uint(f2_hi(x));
    }
    return __byte_perm(__byte_perm(res[0], res[1], 0x0040), __byte_perm(res[2], res[3], 0x0040), 0x5410);
}
/* the same for four 16-bit pairs (P010) */
__device__ __forceinline__ uint2 blend16_quad(const F2 &t12, const F2 &t21, const Levels16 &L, int cz, uint2 a, uint2 b) {
    const F2 negMagic = f2_make(-HR_MAGIC, -HR_MAGIC), magic = f2_make(HR_MAGIC, HR_MAGIC);
    uint32_t res[4];
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
        const uint32_t wa = k < 2 ? a.x : a.y, wb = k < 2 ? b.x : b.y;
        const F2 fa = f2_add(f2_bits(__byte_perm(wa, 0x4B000000u, 0x7510u), __byte_perm(wa, 0x4B000000u, 0x7532u)), negMagic);
        const F2 fb = f2_add(f2_bits(__byte_perm(wb, 0x4B000000u, 0x7510u), __byte_perm(wb, 0x4B000000u, 0x7532u)), negMagic);
        F2 bl = f2_fma(fa, t21, f2_mul(fb, t12));
        bl = f2_make(fminf(f2_lo(bl), 65535.0f), fminf(f2_hi(bl), 65535.0f));
        const F2 v = f2_add(f2_add_rz(bl, magic), negMagic);
        res[k] = cz ? levels_uv16(f2_lo(v), L) : levels_y16(f2_lo(v), L);
        res[k + 1] = cz ? levels_uv16(f2_hi(v), L) : levels_y16(f2_hi(v), L);
    }
    return make_uint2(res[0] | (res[1] << 16), res[2] | (res[3] << 16));
}

/* ROWS consecutive source rows of one 4-sample column, for a run that lies inside the frame: the first
 * word and the alignment are computed once, every row is two (chroma, odd displacement: three) aligned
 * 32-bit loads and a funnel shift. o = sample offset of the run's first sample in the plane. */
template <int ROWS>
__device__ __forceinline__ void load_rows_interior(const uint8_t *plane, int o, int W, bool odd, uint32_t (&out)[ROWS]) {
    if (odd) o -= 1; /* six samples from cx0+d-1, picked 0,3,2,5 (see load_run4_uv_odd) */
    const uint32_t *q = reinterpret_cast<const uint32_t *>(plane) + (o >> 2);
    const unsigned sh = (unsigned)(o & 3) * 8;
    const int W4 = W >> 2;
    if (!odd) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) out[r] = __funnelshift_r(__ldg(q + r * W4), __ldg(q + r * W4 + 1), sh);
    } else {
        const bool third = (o & 3) == 3;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const uint32_t w0 = __ldg(q + r * W4), w1 = __ldg(q + r * W4 + 1), w2 = third ? __ldg(q + r * W4 + 2) : 0u;
            out[r] = __byte_perm(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), 0x5230);
        }
    }
}
template <int ROWS>
__device__ __forceinline__ void load_rows_interior(const uint16_t *plane, int o, int W, bool odd, uint2 (&out)[ROWS]) {
    if (odd) o -= 1;
    const uint32_t *q = reinterpret_cast<const uint32_t *>(plane) + (o >> 1);
    const unsigned sh = (unsigned)(o & 1) * 16;
    const int W2 = W >> 1;
    if (!odd) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const uint32_t w0 = __ldg(q + r * W2), w1 = __ldg(q + r * W2 + 1), w2 = sh ? __ldg(q + r * W2 + 2) : 0u;
            out[r] = make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
        }
    } else {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const uint32_t w0 = __ldg(q + r * W2), w1 = __ldg(q + r * W2 + 1), w2 = __ldg(q + r * W2 + 2), w3 = sh ? __ldg(q + r * W2 + 3) : 0u;
            const uint32_t s0 = __funnelshift_r(w0, w1, sh), s1 = __funnelshift_r(w1, w2, sh), s2 = __funnelshift_r(w2, w3, sh);
            out[r] = make_uint2(__byte_perm(s0, s1, 0x7610), __byte_perm(s1, s2, 0x7610));
        }
    }
}

/* ROWS = rows per thread (4, or 8 when the lattice cell is at least 8 rows tall): thread = 4 samples x
 * ROWS rows. lumaGroups = ceil(H / ROWS): row groups of the luma plane come first, then the chroma plane's. */
template <typename T, int ROWS>
__global__ void __launch_bounds__(256, ROWS == 4 ? 6 : 4) warp_blend_kernel(const WarpParams<T> P, int useFast, int lumaGroups, int lumaG0, int chromaG0,
                                                                            int chromaGN) {
    /* the launch covers lumaGroups row groups of the luma plane starting at group lumaG0, then chromaGN groups
     * of the chroma plane starting at chromaG0 (whole frame: 0, all, 0, all; a spatial band: its rows only) */
    constexpr bool is16 = SampleTraits<T>::is16;
    typedef typename RunType<T>::type Run;
    const int cx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int rg = blockIdx.y * blockDim.y + threadIdx.y;
    const int cz = rg >= lumaGroups;
    if (cz && rg - lumaGroups >= chromaGN) return;
    const int cy0 = (cz ? chromaG0 + rg - lumaGroups : lumaG0 + rg) * ROWS;
    const int planeH = cz ? (P.H >> 1) : P.H;
    if (cx0 >= P.aW || cy0 >= planeH) return;
    const T *s12 = cz ? P.f1uv : P.f1y;
    const T *s21 = cz ? P.f2uv : P.f2y;
    T *out = cz ? P.outUV : P.outY;
    const int nrows = hr_min(ROWS, planeH - cy0);

    bool done = false;
    if (useFast && cx0 + 3 < P.aW) {
        const int half = P.aW >> 1;
        if (P.mode == 5 && cx0 + 3 < half) {
            /* left half of SideBySide1: frame1 as it is (warpFrameKernel.cl:131-133) */
            for (int r = 0; r < nrows; ++r) {
                const size_t o = (size_t)(cy0 + r) * P.W + cx0;
                if (is16) *reinterpret_cast<uint2 *>(out + o) = *reinterpret_cast<const uint2 *>(s12 + o);
                else *reinterpret_cast<uint32_t *>(out + o) = *reinterpret_cast<const uint32_t *>(s12 + o);
            }
            done = true;
        } else if (!(P.mode == 5 && cx0 < half)) {
            const CellFlow f = cell_flow(P, cx0, cy0, cz);
            const float ys = cz ? 0.5f : 1.0f;
            const int d12 = (int)roundf((float)f.x12 * P.t12), d21 = -(int)roundf((float)f.x21 * P.t21);
            const int e12 = (int)roundf((float)f.y12 * P.t12 * ys), e21 = -(int)roundf((float)f.y21 * P.t21 * ys);
            /* every source column inside [1, aW-2]: the mirror/clamp of warpFrameKernel.cl:10-18 is the identity
             * and the four samples are one run; otherwise (frame border) they are fetched one by one */
            const bool in12 = cx0 + d12 >= 1 && cx0 + 3 + d12 <= P.aW - 2;
            const bool in21 = cx0 + d21 >= 1 && cx0 + 3 + d21 <= P.aW - 2;
            const bool odd12 = cz && (d12 & 1), odd21 = cz && (d21 & 1);
            /* rows inside [1, planeH-2] for all rows of the block -> consecutive source rows */
            const bool rows12 = cy0 + e12 >= 1 && cy0 + ROWS - 1 + e12 <= planeH - 2;
            const bool rows21 = cy0 + e21 >= 1 && cy0 + ROWS - 1 + e21 <= planeH - 2;
            Run ra[ROWS], rb[ROWS];
            if (in12 && in21 && rows12 && rows21 && P.mode >= 2) {
                /* the common case: both source blocks lie inside the frame */
                load_rows_interior<ROWS>(s12, (cy0 + e12) * P.W + cx0 + d12, P.W, odd12, ra);
                load_rows_interior<ROWS>(s21, (cy0 + e21) * P.W + cx0 + d21, P.W, odd21, rb);
            } else {
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const int cy = hr_min(cy0 + r, planeH - 1);
                    ra[r] = rb[r] = Run();
                    if (P.mode != 1) {
                        const T *row = s12 + (size_t)warp_mirror(cy + e12, planeH) * P.W;
                        if (in12) ra[r] = odd12 ? load_run4_uv_odd(row + cx0 + d12 - 1) : load_run4(row + cx0 + d12);
                        else ra[r] = load_run4_border(row, cx0, d12, P.aW, cz);
                    }
                    if (P.mode != 0) {
                        const T *row = s21 + (size_t)warp_mirror(cy + e21, planeH) * P.W;
                        if (in21) rb[r] = odd21 ? load_run4_uv_odd(row + cx0 + d21 - 1) : load_run4(row + cx0 + d21);
                        else rb[r] = load_run4_border(row, cx0, d21, P.aW, cz);
                    }
                }
            }
            /* blend, levels, store — two samples per instruction (FMUL2 / FFMA2 / FADD2) */
            T *po = out + cy0 * P.W + cx0;
            if constexpr (!is16) {
                if (P.mode < 2) {
#pragma unroll
                    for (int r = 0; r < ROWS; ++r)
                        if (r < nrows) *reinterpret_cast<uint32_t *>(po + r * P.W) = P.mode == 0 ? ra[r] : rb[r];
                } else {
                    const Blend8 B = make_blend8(P.t12, P.t21, P.black, P.white, cz);
                    if (B.clampNeeded) {
#pragma unroll
                        for (int r = 0; r < ROWS; ++r)
                            if (r < nrows) *reinterpret_cast<uint32_t *>(po + r * P.W) = blend8_quad<true>(B, ra[r], rb[r]);
                    } else {
#pragma unroll
                        for (int r = 0; r < ROWS; ++r)
                            if (r < nrows) *reinterpret_cast<uint32_t *>(po + r * P.W) = blend8_quad<false>(B, ra[r], rb[r]);
                    }
                }
            } else {
                const Levels16 L = make_levels16(P.black, P.white);
                const F2 t12 = f2_make(P.t12, P.t12), t21 = f2_make(P.t21, P.t21);
#pragma unroll
                for (int r = 0; r < ROWS; ++r)
                    if (r < nrows) *reinterpret_cast<uint2 *>(po + r * P.W) = P.mode == 0 ? ra[r] : P.mode == 1 ? rb[r] : blend16_quad(t12, t21, L, cz, ra[r], rb[r]);
            }
            done = true;
        }
    }
    if (done) return;
    /* per-sample path */
    for (int r = 0; r < nrows; ++r) {
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const int cx = cx0 + k;
            if (cx < P.aW) out[(size_t)(cy0 + r) * P.W + cx] = (T)warp_sample(P, cx, cy0 + r, cz);
        }
    }
}

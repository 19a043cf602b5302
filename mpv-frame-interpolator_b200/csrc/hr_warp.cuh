#pragma once
#include "hr_common.cuh"

/* ------------------------------------------------------------------------------------------ */
/* warp + flip + blend + levels + output modes (K5)                                              */
/* ------------------------------------------------------------------------------------------ */
/* warpFrameKernel.cl:10-18 */
__device__ __forceinline__ int warp_mirror(int pos, int dim) {
    int res = pos;
    if (pos >= dim - 1) res = pos - ((pos - (dim - 2)) * 2);
    else if (pos < 1) res = -pos + 1;
    return hr_min(hr_max(res, 1), dim - 2);
}
__device__ __forceinline__ unsigned sat_u8(float v) { return __float2uint_rz(fmaxf(fminf(v, 255.0f), 0.0f)); }

/* warpFrameKernel.cl:1-7 */
__device__ __forceinline__ unsigned levels_y8(float v, float black, float white) { return sat_u8((v - black) / (white - black) * 255.0f); }
__device__ __forceinline__ unsigned levels_uv8(float v, float white) { return sat_u8((v - 128.0f) / white * 255.0f + 128.0f); }
/* P010, by construction */
__device__ __forceinline__ unsigned levels_y16(float v, float black, float white) {
    const float b16 = black / 255.0f * 65472.0f, w16 = white / 255.0f * 65472.0f;
    return __float2uint_rz(fmaxf(fminf((v - b16) / (w16 - b16) * 65472.0f, 65472.0f), 0.0f)) & 0xFFC0u;
}
__device__ __forceinline__ unsigned levels_uv16(float v, float white) {
    const float w16 = white / 255.0f * 65472.0f;
    return __float2uint_rz(fmaxf(fminf((v - 32768.0f) / w16 * 65472.0f + 32768.0f, 65472.0f), 0.0f)) & 0xFFC0u;
}

__global__ void levels_lut_kernel(uint8_t *lut, int *identity, float black, float white) {
    const int v = threadIdx.x;
    const unsigned y = levels_y8((float)v, black, white), c = levels_uv8((float)v, white);
    lut[v] = (uint8_t)y;
    lut[256 + v] = (uint8_t)c;
    const int same = __syncthreads_and(y == (unsigned)v && c == (unsigned)v);
    if (v == 0) *identity = same;
}

/* warpFrameKernel.cl:21-111 */
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
        const float hue = angle_deg / 360.0f;
        const int h_i = (int)(hue * 6.0f);
        const float f = hue * 6.0f - (float)h_i;
        const float q = 1.0f - f;
        switch (h_i % 6) {
            case 0: r = 255; g = __float2uint_rz(f * 255.0f) & 255u; b = 0; break;
            case 1: r = __float2uint_rz(q * 255.0f) & 255u; g = 255; b = 0; break;
            case 2: r = 0; g = 255; b = __float2uint_rz(f * 255.0f) & 255u; break;
            case 3: r = 0; g = __float2uint_rz(q * 255.0f) & 255u; b = 255; break;
            case 4: r = __float2uint_rz(f * 255.0f) & 255u; g = 0; b = 255; break;
            case 5: r = 255; g = 0; b = __float2uint_rz(q * 255.0f) & 255u; break;
            default: r = g = b = 0; break;
        }
        r = sat_u8((float)r / 255.0f * (float)(ax + ay) * (float)resImpact);
        g = sat_u8((float)g / 255.0f * (float)ay * 2.0f * (float)resImpact);
        b = sat_u8((float)b / 255.0f * (float)(ax + ay) * (float)resImpact);
    }
    if (channel == 0) return ((sat_u8((float)r * 0.299f + (float)g * 0.587f + (float)b * 0.114f) >> 1) + (currPixel >> 1)) & 255u;
    if (channel == 1) return sat_u8((float)r * -0.168736f + (float)g * -0.331264f + (float)b * 0.5f + 128.0f);
    return sat_u8((float)r * 0.5f + (float)g * -0.418688f + (float)b * -0.081312f + 128.0f);
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
    const size_t ln = (size_t)P.lw * P.lh;
    CellFlow f;
    f.x12 = __ldg(P.flow + (size_t)scy * P.lw + scx);
    f.y12 = __ldg(P.flow + ln + (size_t)scy * P.lw + scx);
    const int fy = hr_min(hr_max(scy - (f.y12 >> s), 0), P.lh - 1);
    const int fx = hr_min(hr_max(scx - (f.x12 >> s), 0), P.lw - 1);
    f.x21 = __ldg(P.flow + (size_t)fy * P.lw + fx);
    f.y21 = __ldg(P.flow + ln + (size_t)fy * P.lw + fx);
    return f;
}

template <typename T>
__device__ __forceinline__ unsigned finish_blend(const WarpParams<T> &P, unsigned a, unsigned b, int cz, int cx, const CellFlow &f) {
    if (!SampleTraits<T>::is16) {
        unsigned v = __float2uint_rz((float)a * P.t21 + (float)b * P.t12);
        if (P.mode == 3) {
            v = visualize_flow(-f.x12, -f.y12, v & 255u, cz + (cx & (cz ? 1 : 0)), P.s <= 2 ? 4 : 1);
            return cz ? levels_uv8((float)v, P.white) : levels_y8((float)v, P.black, P.white);
        }
        v &= 255u;
        return P.lutIdentity ? v : (unsigned)__ldg(P.lut + (cz ? 256 : 0) + v);
    } else {
        const unsigned v = __float2uint_rz(fminf((float)a * P.t21 + (float)b * P.t12, 65535.0f));
        if (P.mode == 3) {
            const unsigned v8 = visualize_flow(-f.x12, -f.y12, v >> 8, cz + (cx & (cz ? 1 : 0)), P.s <= 2 ? 4 : 1);
            const unsigned l8 = cz ? levels_uv8((float)v8, P.white) : levels_y8((float)v8, P.black, P.white);
            return l8 << 8;
        }
        return cz ? levels_uv16((float)v, P.white) : levels_y16((float)v, P.black, P.white);
    }
}

/* One output sample, every mode: the general path (frame borders, modes 3/4/6, tiny frames). */
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

/* Four consecutive samples starting at an arbitrary sample address, from aligned 32-bit loads. */
__device__ __forceinline__ uint32_t load4_u8(const uint8_t *p) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    const unsigned shft = (unsigned)(a & 3) * 8;
    const uint32_t lo = __ldg(q);
    const uint32_t hi = shft ? __ldg(q + 1) : 0u;
    return __funnelshift_r(lo, hi, shft);
}
/* 8 consecutive bytes from an arbitrary byte address (chroma with odd displacement) */
__device__ __forceinline__ uint2 load8_u8(const uint8_t *p) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    const unsigned shft = (unsigned)(a & 3) * 8;
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1);
    const uint32_t w2 = shft ? __ldg(q + 2) : 0u;
    return make_uint2(__funnelshift_r(w0, w1, shft), __funnelshift_r(w1, w2, shft));
}

/* Interior fast path, NV12: the four samples of a 4-aligned quad share one lattice cell
 * (s >= 2), so each source is one translated run. Chroma keeps U/V parity: with an odd
 * displacement d the U bytes come from cx+d-1 and the V bytes from cx+d+1
 * (warpFrameKernel.cl:171 `(newCx & ~1) + (cx & 1)`). Returns false when a border is touched. */
__device__ __forceinline__ bool fetch_quad_u8(const uint8_t *plane, int dimX, int aW, int cx0, int row, int d, int cz, uint32_t &out) {
    if (cx0 + d < 1 || cx0 + 3 + d > aW - 2) return false;
    const uint8_t *base = plane + (size_t)row * dimX;
    if (!cz || !(d & 1)) {
        out = load4_u8(base + cx0 + d);
    } else {
        const uint2 w = load8_u8(base + cx0 + d - 1);
        out = __byte_perm(w.x, w.y, 0x5230);
    }
    return true;
}

template <typename T>
__global__ void __launch_bounds__(256) warp_blend_kernel(const WarpParams<T> P, int useFast) {
    const int cx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int row = blockIdx.y * blockDim.y + threadIdx.y;
    if (cx0 >= P.aW || row >= P.H + (P.H >> 1)) return;
    const int cz = row >= P.H;
    const int cy = cz ? row - P.H : row;
    T *outRow = (cz ? P.outUV : P.outY) + (size_t)cy * P.W;

    if (!SampleTraits<T>::is16 && useFast && P.s >= 2 && cx0 + 3 < P.aW && (P.mode <= 2 || P.mode == 5)) {
        const uint8_t *s12 = (const uint8_t *)(cz ? P.f1uv : P.f1y);
        const uint8_t *s21 = (const uint8_t *)(cz ? P.f2uv : P.f2y);
        const int half = P.aW >> 1;
        if (P.mode == 5 && cx0 + 3 < half) {
            *reinterpret_cast<uint32_t *>(outRow + cx0) = *reinterpret_cast<const uint32_t *>(s12 + (size_t)cy * P.W + cx0);
            return;
        }
        if (!(P.mode == 5 && cx0 < half)) {
            const CellFlow f = cell_flow(P, cx0, cy, cz);
            const int dY = cz ? (P.H >> 1) : P.H;
            const float ys = cz ? 0.5f : 1.0f;
            const int d12 = (int)roundf((float)f.x12 * P.t12), d21 = -(int)roundf((float)f.x21 * P.t21);
            const int ny12 = warp_mirror(cy + (int)roundf((float)f.y12 * P.t12 * ys), dY);
            const int ny21 = warp_mirror(cy - (int)roundf((float)f.y21 * P.t21 * ys), dY);
            uint32_t a = 0, b = 0;
            bool ok = true;
            if (P.mode != 1) ok = fetch_quad_u8(s12, P.W, P.aW, cx0, ny12, d12, cz, a);
            if (ok && P.mode != 0) ok = fetch_quad_u8(s21, P.W, P.aW, cx0, ny21, d21, cz, b);
            if (ok) {
                uint32_t o;
                if (P.mode == 0) o = a;
                else if (P.mode == 1) o = b;
                else {
                    o = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const unsigned av = (a >> (8 * k)) & 255u, bv = (b >> (8 * k)) & 255u;
                        unsigned v = __float2uint_rz((float)av * P.t21 + (float)bv * P.t12) & 255u;
                        if (!P.lutIdentity) v = __ldg(P.lut + (cz ? 256 : 0) + v);
                        o |= v << (8 * k);
                    }
                }
                *reinterpret_cast<uint32_t *>(outRow + cx0) = o;
                return;
            }
        }
    }
    /* general path */
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const int cx = cx0 + k;
        if (cx < P.aW) outRow[cx] = (T)warp_sample(P, cx, cy, cz);
    }
}


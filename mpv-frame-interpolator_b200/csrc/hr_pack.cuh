#pragma once
#include "hr_common.cuh"

/* ------------------------------------------------------------------------------------------ */
/* pack: frame -> phase-planar packed words                                                      */
/*   word(px,py)[ly][lx] = Y(x,y) | U(x,y) << 8 | V(x,y) << 16, x = lx<<s | px, y = ly<<s | py,     */
/*   with U,V taken at chroma row y>>1, byte column x&~1 (+1): calcDeltaSumsKernel.cl:96-98.       */
/*   P010: the top 8 bits of every sample.                                       */
/*                                                                                                */
/* HBM-bound (reads 1.5 B, writes 4 B per pixel). Thread = 16 consecutive pixels of two rows that  */
/* share a chroma row: 128-bit loads, and for every x-phase one vector store of 16 >> s            */
/* consecutive words, so a warp writes 32 * (16 >> s) * 4 contiguous bytes per phase plane.        */
/* ------------------------------------------------------------------------------------------ */
template <typename T>
__device__ __forceinline__ uint32_t top8(T v);
template <>
__device__ __forceinline__ uint32_t top8<uint8_t>(uint8_t v) { return v; }
template <>
__device__ __forceinline__ uint32_t top8<uint16_t>(uint16_t v) { return (uint32_t)v >> 8; }

/* 16 consecutive samples, their top 8 bits, as 16 bytes in a uint4 (16-byte aligned source) */
__device__ __forceinline__ uint4 load16_top8(const uint8_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ uint4 load16_top8(const uint16_t *p) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p)), b = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
    /* high bytes of the eight halves of a, then of b */
    return make_uint4(__byte_perm(a.x, a.y, 0x7531), __byte_perm(a.z, a.w, 0x7531), __byte_perm(b.x, b.y, 0x7531), __byte_perm(b.z, b.w, 0x7531));
}
__device__ __forceinline__ uint32_t byte_of(const uint4 &v, int i) {
    const uint32_t w = i < 4 ? v.x : i < 8 ? v.y : i < 12 ? v.z : v.w;
    return (w >> (8 * (i & 3))) & 255u;
}

template <int N>
__device__ __forceinline__ void store_words(uint32_t *dst, const uint32_t (&w)[16], int first, int stride);
template <>
__device__ __forceinline__ void store_words<16>(uint32_t *dst, const uint32_t (&w)[16], int, int) {
#pragma unroll
    for (int j = 0; j < 4; ++j) reinterpret_cast<uint4 *>(dst)[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}
template <>
__device__ __forceinline__ void store_words<8>(uint32_t *dst, const uint32_t (&w)[16], int first, int) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
        reinterpret_cast<uint4 *>(dst)[j] = make_uint4(w[first + 8 * j], w[first + 8 * j + 2], w[first + 8 * j + 4], w[first + 8 * j + 6]);
}
template <>
__device__ __forceinline__ void store_words<4>(uint32_t *dst, const uint32_t (&w)[16], int first, int) {
    *reinterpret_cast<uint4 *>(dst) = make_uint4(w[first], w[first + 4], w[first + 8], w[first + 12]);
}
template <>
__device__ __forceinline__ void store_words<2>(uint32_t *dst, const uint32_t (&w)[16], int first, int) {
    *reinterpret_cast<uint2 *>(dst) = make_uint2(w[first], w[first + 8]);
}
template <>
__device__ __forceinline__ void store_words<1>(uint32_t *dst, const uint32_t (&w)[16], int first, int) {
    *dst = w[first];
}

/* S = resolution scalar 0..4 (16 >> S words per phase and thread). Needs W % 16 == 0 and 16-byte
 * aligned planes; pack_frame_kernel below handles everything else. */
template <typename T, int S>
__global__ void __launch_bounds__(128) pack_frame16_kernel(const T *__restrict__ yPlane, const T *__restrict__ uvPlane, uint32_t *__restrict__ packed,
                                                            int W, int H, int planePitch, int planeSize, int crow0) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int crow = crow0 + blockIdx.y; /* chroma row = pair of luma rows; crow0: first one of the rows to pack */
    if (x0 >= W) return;
    constexpr int M = (1 << S) - 1, N = 16 >> S;
    const uint4 uv = load16_top8(uvPlane + (size_t)crow * W + x0);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int row = crow * 2 + h;
        if (row >= H) break;
        const uint4 yv = load16_top8(yPlane + (size_t)row * W + x0);
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = byte_of(yv, i) | (byte_of(uv, i & ~1) << 8) | (byte_of(uv, (i & ~1) + 1) << 16);
        uint32_t *base = packed + (size_t)(((row & M) << S)) * planeSize + (size_t)(row >> S) * planePitch + (x0 >> S);
#pragma unroll
        for (int p = 0; p <= M; ++p) store_words<N>(base + (size_t)p * planeSize, w, p, 1 << S);
    }
}

/* any geometry: one word per thread */
template <typename T>
__global__ void pack_frame_kernel(const T *__restrict__ yPlane, const T *__restrict__ uvPlane, uint32_t *__restrict__ packed,
                                  int W, int H, int s, int lw, int planePitch, int planeSize, int row0) {
    const int row = row0 + blockIdx.y;
    const int lx = blockIdx.x * blockDim.x + threadIdx.x;
    const int px = threadIdx.y;
    const int x = (lx << s) | px;
    if (lx >= lw || x >= W) return;
    const uint32_t yv = top8<T>(__ldg(yPlane + (size_t)row * W + x));
    const T *uvp = uvPlane + (size_t)(row >> 1) * W + (x & ~1);
    const uint32_t uv = top8<T>(__ldg(uvp));
    const uint32_t vv = top8<T>(__ldg(uvp + 1));
    const int m = (1 << s) - 1;
    const int plane = ((row & m) << s) | px;
    packed[(size_t)plane * planeSize + (size_t)(row >> s) * planePitch + lx] = yv | (uv << 8) | (vv << 16);
}

#pragma once
#include "hr_common.cuh"

/* ------------------------------------------------------------------------------------------ */
/* pack: frame -> phase-planar packed words                                                      */
/*   word(px,py)[ly][lx] = Y(x,y) | U(x,y) << 8 | V(x,y) << 16, x = lx<<s | px, y = ly<<s | py,     */
/*   with U,V taken at chroma row y>>1, byte column x&~1 (+1): calcDeltaSumsKernel.cl:96-98.       */
/*   P010: the top 8 bits of every sample.                                       */
/* ------------------------------------------------------------------------------------------ */
template <typename T>
__device__ __forceinline__ uint32_t top8(T v);
template <>
__device__ __forceinline__ uint32_t top8<uint8_t>(uint8_t v) { return v; }
template <>
__device__ __forceinline__ uint32_t top8<uint16_t>(uint16_t v) { return (uint32_t)v >> 8; }

template <typename T>
__global__ void pack_frame_kernel(const T *__restrict__ yPlane, const T *__restrict__ uvPlane, uint32_t *__restrict__ packed,
                                  int W, int H, int s, int lw, int planePitch, int planeSize) {
    const int row = blockIdx.y;
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


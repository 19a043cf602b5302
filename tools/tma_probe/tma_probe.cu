// Stand-alone probe of the TMA box the staged search uses: 16 boxes of 52 x 50 words out of a (480 x 270 x 16) tensor,
// one mbarrier, out-of-range coordinates. Prints the first mismatch against a host gather, and the time of the fetch.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe [variant]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)
struct alignas(64) TMap { unsigned char b[128]; };
static int PWh = 52, PHh = 50, PLANEh = 2624;
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int VARIANT>
__global__ void __launch_bounds__(512, 1) probe(const __grid_constant__ TMap tm, const TMap *gtm, uint32_t *out, int c0, int c1, long long *cycles, int PW, int PH, int PLANE) {
    extern __shared__ __align__(128) uint32_t stage[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (tid == 0) {
        if (VARIANT == 2) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar)) : "memory");
        } else if (VARIANT == 3) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(16 * 4096) : "memory");
            for (int pl = 0; pl < 16; ++pl)
                asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(stage + pl * PLANE)), "l"(out + pl * 1024), "r"(4096), "r"(s32(&bar)) : "memory");
        } else {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(16 * PH * PW * 4) : "memory");
        for (int pl = 0; pl < 16; ++pl) {
            if (VARIANT == 5)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(s32(stage + pl * PLANE)), "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1 + pl * 270) : "memory");
            else if (VARIANT == 4)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(s32(stage + pl * PLANE)), "l"(gtm), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(pl) : "memory");
            else if (VARIANT == 0)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(s32(stage + pl * PLANE)), "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(pl) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(s32(stage + pl * PLANE)), "l"(&tm), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(pl) : "memory");
        }
        }
    }
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(&bar)), "r"(0) : "memory");
    const long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    for (int i = tid; i < 16 * PLANE; i += blockDim.x) out[(size_t)blockIdx.x * 16 * PLANE + i] = stage[i];
}
int main(int argc, char **argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    if (argc > 3) { PWh = atoi(argv[2]); PHh = atoi(argv[3]); PLANEh = (PWh * PHh * 4 + 127) / 128 * 32; }
    const int PW = PWh, PH = PHh, PLANE = PLANEh;
    printf("box %d x %d words, plane stride %d words\n", PW, PH, PLANE);
    const int pitch = 480, lh = 270, planes = 16;
    const size_t n = (size_t)pitch * lh * planes;
    std::vector<uint32_t> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (uint32_t)(i * 2654435761u) | 1u;
    uint32_t *d, *out;
    long long *cyc;
    const int grid = 135;
    CK(cudaMalloc(&d, n * 4));
    CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, (size_t)grid * 16 * PLANE * 4));
    CK(cudaMemset(out, 0xff, (size_t)grid * 16 * PLANE * 4));
    CK(cudaMalloc(&cyc, grid * 8));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no encoder\n"); return 1; }
    typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)lh, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * lh * 4};
    const cuuint32_t box[3] = {(cuuint32_t)PW, (cuuint32_t)PH, 1}, es[3] = {1, 1, 1};
    CUtensorMap tm;
    const cuuint64_t dims2[2] = {(cuuint64_t)pitch, (cuuint64_t)lh * planes};
    CUresult r = variant == 5 ? ((Enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims2, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) :
                 ((Enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    TMap t;
    memcpy(&t, &tm, 128);
    const int c0 = argc > 4 ? atoi(argv[4]) : 32 * 3 - 9, c1 = argc > 5 ? atoi(argv[5]) : 32 * 8 - 9; /* a tile of the last tile row: rows 247 .. 296 of 270 */
    auto k = variant == 0 ? probe<0> : variant == 1 ? probe<1> : variant == 2 ? probe<2> : variant == 3 ? probe<3> : variant == 4 ? probe<4> : probe<5>;
    TMap *gtm;
    CK(cudaMalloc(&gtm, 128));
    CK(cudaMemcpy(gtm, &t, 128, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * PLANE * 4));
    for (int rep = 0; rep < 3; ++rep) {
        k<<<grid, 512, 16 * PLANE * 4>>>(t, gtm, out, c0, c1, cyc, PW, PH, PLANE);
        CK(cudaDeviceSynchronize());
    }
    std::vector<uint32_t> o((size_t)16 * PLANE);
    CK(cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<long long> hc(grid);
    CK(cudaMemcpy(hc.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int pl = 0; pl < 16; ++pl)
        for (int rr = 0; rr < PH; ++rr)
            for (int cc = 0; cc < PW; ++cc) {
                const int gy = c1 + rr, gx = c0 + cc;
                const uint32_t exp = (gy >= 0 && gy < lh && gx >= 0 && gx < pitch) ? h[((size_t)pl * lh + gy) * pitch + gx] : 0u;
                const uint32_t got = o[(size_t)pl * PLANE + rr * PW + cc];
                if (exp != got && bad++ < 5) printf("mismatch plane %d row %d col %d: got %08x exp %08x\n", pl, rr, cc, got, exp);
            }
    long long mx = 0, sum = 0;
    for (auto v : hc) { mx = v > mx ? v : mx; sum += v; }
    printf("variant %d: %lld mismatches; fetch of 16 boxes (166400 B) per CTA, 135 CTAs at once: mean %lld max %lld cycles\n", variant, bad, sum / grid, mx);
    return bad != 0;
}

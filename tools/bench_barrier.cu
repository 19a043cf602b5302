// Microbenchmark: cost of one grid-wide barrier for 135 CTAs x 128 threads on B200, three variants,
// plus the cost of a dependent L2 round trip (ld.cg chain) and a shuffle+redux chain, to budget the
// search kernel's critical path. Build: nvcc -O3 -arch=sm_100a -o bench_barrier bench_barrier.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int VARIANT>
__global__ void k_barrier(unsigned long long *bar, int iters, long long *cycles) {
    cg::grid_group grid = cg::this_grid();
    unsigned long long target = 0;
    const unsigned n = gridDim.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (VARIANT == 0) {  // fence + red.release + ld.acquire poll + fence
            __syncthreads();
            target += n;
            if (threadIdx.x == 0) {
                __threadfence();
                asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar), "l"(1ULL) : "memory");
                while (ld_acq(bar) < target) {}
                __threadfence();
            }
            __syncthreads();
        } else if (VARIANT == 1) {  // red.release + relaxed poll, single acquire fence at the end
            __syncthreads();
            target += n;
            if (threadIdx.x == 0) {
                asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar), "l"(1ULL) : "memory");
                while (ld_relaxed(bar) < target) {}
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
            }
            __syncthreads();
        } else {
            grid.sync();
        }
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = (t1 - t0);
}

__global__ void k_chain(const unsigned *idx, int iters, unsigned *out, long long *cycles) {
    unsigned v = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        unsigned w;
        asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(w) : "l"(idx + v));
        v = w;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = v;
    if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = (t1 - t0);
}

int main() {
    unsigned long long *bar;
    long long *cyc, h;
    cudaMalloc(&bar, 128);
    cudaMalloc(&cyc, 8);
    int iters = 200;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int grid : {135, 148}) {
        for (int v = 0; v < 3; ++v) {
            cudaMemset(bar, 0, 128);
            void *args[] = {&bar, &iters, &cyc};
            const void *fn = v == 0 ? (const void *)k_barrier<0> : v == 1 ? (const void *)k_barrier<1> : (const void *)k_barrier<2>;
            cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(128), args, 0, 0);  // warm
            cudaDeviceSynchronize();
            cudaMemset(bar, 0, 128);
            cudaEventRecord(e0);
            cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(128), args, 0, 0);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("grid %d variant %d: %.3f us per barrier (event), %.0f cycles per barrier (clock64) err=%s\n", grid, v, ms * 1e3 / iters, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
        }
    }
    // dependent L2 round trips
    unsigned *idx, *out;
    int n = 1 << 20;
    cudaMalloc(&idx, n * 4);
    cudaMalloc(&out, 148 * 128 * 4);
    unsigned *hidx = (unsigned *)malloc(n * 4);
    for (int i = 0; i < n; ++i) hidx[i] = (unsigned)((i * 7919u + 12345u) % n);
    cudaMemcpy(idx, hidx, n * 4, cudaMemcpyHostToDevice);
    int ci = 2000;
    for (int rep = 0; rep < 2; ++rep) {
        k_chain<<<135, 32>>>(idx, ci, out, cyc);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent ld.cg chain (4 MB table, L2 resident): %.0f cycles per load\n", (double)h / ci);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("clock rate attr %d kHz\n", clk);
    return 0;
}

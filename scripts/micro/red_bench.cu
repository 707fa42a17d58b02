// Microbenchmark: throughput of the scatter-add primitives the histogram stage can be built from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench red_bench.cu && ./red_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef uint32_t u32; typedef uint64_t u64;
__device__ __forceinline__ u32 mix(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__global__ void fill_idx(u32 *idx, u64 n, u64 range, u64 slice, u32 seed)
{   // items grouped by slice: item i goes to slice (i * n_slices / n), random inside the slice
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 n_slices = (range + slice - 1) / slice;
    const u64 s = (u64)((__uint128_t)i * n_slices / n);
    const u64 lo = s * slice, len = min(slice, range - lo);
    idx[i] = (u32)(lo + (((u64)mix((u32)i ^ seed) << 32 | mix((u32)(i >> 3) + seed)) % len));
}
__global__ void red64(const u32 *__restrict__ idx, u64 n, unsigned long long *h)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(h + __ldcs(idx + i), 0x100000001ull);
}
__global__ void red32(const u32 *__restrict__ idx, u64 n, u32 *h)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(h + __ldcs(idx + i), 1u);
}
__global__ void red32x2(const u32 *__restrict__ idx, u64 n, u32 *h)
{   // interleaved {cov, uniq} as two 32-bit REDs
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { const u64 j = 2ull * __ldcs(idx + i); atomicAdd(h + j, 1u); atomicAdd(h + j + 1, 1u); }
}
__global__ void red16pair(const u32 *__restrict__ idx, u64 n, u32 *h)
{   // {cov:16, uniq:16} packed in one u32 word
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(h + __ldcs(idx + i), 0x10001u);
}
// shared-memory histogram: each CTA owns `bins` consecutive bins and a contiguous share of the items
__global__ void smem_hist(const u32 *__restrict__ idx, u64 n, u32 bins, u32 *out)
{
    extern __shared__ u32 sh[];
    for (u32 k = threadIdx.x; k < bins; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    const u64 per = (n + gridDim.x - 1) / gridDim.x, lo = per * blockIdx.x, hi = min(n, lo + per);
    for (u64 i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&sh[__ldcs(idx + i) % bins], 0x10001u);
    __syncthreads();
    for (u32 k = threadIdx.x; k < bins; k += blockDim.x) out[(u64)blockIdx.x * bins + k] = sh[k];
}
template <class F> float timeit(F f, int reps = 3)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
    return best;
}
int main()
{
    const u64 n = 1ull << 28;   // 268M items
    u32 *idx; cudaMalloc(&idx, n * 4);
    void *h; const u64 hist_bytes = 14ull << 30; cudaMalloc(&h, hist_bytes);
    cudaMemset(h, 0, hist_bytes);
    const int grid = 148 * 16, block = 256;
    struct { const char *name; u64 range_bins; u64 slice_bins; } cfgs[] = {
        {"one 4M-bin slice (L2 resident)", 1ull << 22, 1ull << 22},
        {"one 16M-bin slice", 1ull << 24, 1ull << 24},
        {"1.75G bins, grouped by 4M-bin slice", 1750ull << 20, 1ull << 22},
        {"1.75G bins, grouped by 2M-bin slice", 1750ull << 20, 1ull << 21},
        {"1.75G bins, grouped by 8M-bin slice", 1750ull << 20, 1ull << 23},
        {"1.75G bins, ungrouped (direct)", 1750ull << 20, 1750ull << 20},
    };
    for (auto &c : cfgs) {
        fill_idx<<<(unsigned)((n + 255) / 256), 256>>>(idx, n, c.range_bins, c.slice_bins, 12345u);
        cudaDeviceSynchronize();
        float t64 = timeit([&] { red64<<<grid, block>>>(idx, n, (unsigned long long *)h); });
        float t32 = timeit([&] { red32<<<grid, block>>>(idx, n, (u32 *)h); });
        float t32x2 = timeit([&] { red32x2<<<grid, block>>>(idx, n, (u32 *)h); });
        float t16 = timeit([&] { red16pair<<<grid, block>>>(idx, n, (u32 *)h); });
        printf("%-40s  RED.64 %7.3f ms (%6.1f G/s) | RED.32 %7.3f ms (%6.1f G/s) | 2xRED.32 %7.3f ms | RED.32 packed16 %7.3f ms\n", c.name, t64, n / t64 / 1e6,
               t32, n / t32 / 1e6, t32x2, t16);
    }
    // shared-memory histograms: 148*k CTAs, items pre-grouped per CTA
    for (u32 bins : {8192u, 16384u, 32768u, 49152u}) {
        cudaFuncSetAttribute(smem_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        const int g = 148 * 8;
        float t = timeit([&] { smem_hist<<<g, 512, bins * 4>>>(idx, n, bins, (u32 *)h); });
        printf("smem histogram %6u bins/CTA, 512 thr: %7.3f ms (%6.1f G/s)\n", bins, t, n / t / 1e6);
    }
    // streaming copy for reference
    float tc = timeit([&] { cudaMemcpyAsync(h, (char *)h + (4ull << 30), 4ull << 30, cudaMemcpyDeviceToDevice); });
    printf("copy 4 GiB: %.3f ms (%.1f GB/s r+w)\n", tc, 8.0 * 1024 * 1024 * 1024 / tc / 1e6);
    float tm = timeit([&] { cudaMemsetAsync(h, 0, 8ull << 30); });
    printf("memset 8 GiB: %.3f ms (%.1f GB/s)\n", tm, 8.0 * 1024 * 1024 * 1024 / tm / 1e6);
    return 0;
}

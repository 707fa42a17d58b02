// Microbenchmark of the fused zero-fill + accumulate schemes (roles, look-ahead, slice size).
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <algorithm>
typedef uint32_t u32; typedef uint64_t u64;
__device__ __forceinline__ u32 mix(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__global__ void fill_idx(u32 *idx, u64 n, u64 range, u64 slice, u32 seed)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 n_slices = (range + slice - 1) / slice;
    const u64 s = (u64)((__uint128_t)i * n_slices / n);
    const u64 lo = s * slice, len = min(slice, range - lo);
    idx[i] = (u32)(lo + (((u64)mix((u32)i ^ seed) << 32 | mix((u32)(i >> 3) + seed)) % len));
}
struct Flags { u32 zero_done[4096]; u32 acc_done[4096]; };
__device__ __forceinline__ void wait_count(const u32 *flag, u32 need)
{
    u32 spins = 0;
    while (*(volatile const u32 *)flag < need && ++spins < (1u << 22)) __nanosleep(40);
    __threadfence();
}
// roles: n_zero zeroer CTAs run `ahead` slices ahead of the accumulators
template <int PRE>
__global__ void __launch_bounds__(256) fused(Flags *f, const u32 *__restrict__ items, u64 n, unsigned long long *__restrict__ hist, u64 bins,
                                             u32 shift, u32 n_slices, u32 n_zero, u32 ahead, int do_fence)
{
    const u32 tid = threadIdx.x;
    const u32 n_acc = gridDim.x - n_zero;
    if (blockIdx.x < n_zero) {
        for (u32 b = 0; b < n_slices; ++b) {
            if (b >= ahead + 1) { if (tid == 0) wait_count(&f->acc_done[b - ahead - 1], n_acc); __syncthreads(); }
            const u64 lo = (u64)b << shift, hi = min(bins, (u64)(b + 1) << shift);
            const u64 n64 = (hi - lo) >> 6, per = (n64 + n_zero - 1) / n_zero;
            const u64 c0 = min(n64, per * blockIdx.x), c1 = min(n64, c0 + per);
            uint4 *dst = reinterpret_cast<uint4 *>(hist + lo + (c0 << 6));
            const u32 n16 = (u32)((c1 - c0) << 5);
            const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll 8
            for (u32 k = tid; k < n16; k += 256) dst[k] = zero;
            __syncthreads();
            if (tid == 0) { if (do_fence) __threadfence(); atomicAdd(&f->zero_done[b], 1u); }
        }
        return;
    }
    const u32 me = blockIdx.x - n_zero;
    const u64 per_slice = n / n_slices;
    u32 pre[PRE];
    for (u32 b = 0; b < n_slices; ++b) {
        const u64 s0 = per_slice * b, cnt = (b + 1 == n_slices) ? n - s0 : per_slice, per = (cnt + n_acc - 1) / n_acc;
        const u64 first = s0 + min(cnt, per * me), last = s0 + min(cnt, per * (me + 1));
#pragma unroll
        for (int k = 0; k < PRE; ++k) { const u64 j = first + tid + (u64)k * 256; pre[k] = j < last ? __ldcs(items + j) : 0xFFFFFFFFu; }
        if (tid == 0) wait_count(&f->zero_done[b], n_zero);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PRE; ++k) if (pre[k] != 0xFFFFFFFFu) atomicAdd(hist + pre[k], 0x100000001ull);
        for (u64 j = first + tid + (u64)PRE * 256; j < last; j += 256) atomicAdd(hist + __ldcs(items + j), 0x100000001ull);
        __syncthreads();
        if (tid == 0) atomicAdd(&f->acc_done[b], 1u);
    }
}
__global__ void red64(const u32 *__restrict__ idx, u64 n, unsigned long long *h)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(h + __ldcs(idx + i), 0x100000001ull);
}
// zero only, by all CTAs of a persistent grid, slice after slice (no flags): store throughput of the zero pattern
__global__ void __launch_bounds__(256) zero_only(unsigned long long *__restrict__ hist, u64 bins, u32 shift, u32 n_slices)
{
    for (u32 b = 0; b < n_slices; ++b) {
        const u64 lo = (u64)b << shift, hi = min(bins, (u64)(b + 1) << shift);
        const u64 n64 = (hi - lo) >> 6, per = (n64 + gridDim.x - 1) / gridDim.x;
        const u64 c0 = min(n64, per * blockIdx.x), c1 = min(n64, c0 + per);
        uint4 *dst = reinterpret_cast<uint4 *>(hist + lo + (c0 << 6));
        const u32 n16 = (u32)((c1 - c0) << 5);
        const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll 8
        for (u32 k = threadIdx.x; k < n16; k += 256) dst[k] = zero;
    }
}
template <class F> float timeit(F f, int reps = 3)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
    return best;
}
int main(int argc, char **argv)
{
    const u64 n = argc > 1 ? strtoull(argv[1], 0, 10) : 1000000000ull;
    const u64 bins = 1744ull << 20;
    u32 *idx; cudaMalloc(&idx, n * 4);
    unsigned long long *h; cudaMalloc(&h, bins * 8);
    Flags *f; cudaMalloc(&f, sizeof(Flags));
    int per_sm = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused<12>, 256, 0);
    printf("n=%llu items, %llu bins, fused<12> occupancy %d CTAs/SM\n", (unsigned long long)n, (unsigned long long)bins, per_sm);
    for (u32 shift : {22u, 21u, 23u}) {
        const u32 n_slices = (u32)((bins + (1ull << shift) - 1) >> shift);
        fill_idx<<<(unsigned)((n + 255) / 256), 256>>>(idx, n, bins, 1ull << shift, 777u);
        cudaDeviceSynchronize();
        float tm = timeit([&] { cudaMemsetAsync(h, 0, bins * 8); });
        float tr = timeit([&] { red64<<<148 * 16, 256>>>(idx, n, h); });
        float tz = timeit([&] { zero_only<<<148 * 8, 256>>>(h, bins, shift, n_slices); });
        float tz1 = timeit([&] { zero_only<<<148, 256>>>(h, bins, shift, n_slices); });
        printf("slice 2^%u bins (%u slices): memset %.3f ms | grouped RED (fills) %.3f ms | zero_only 1184 CTAs %.3f ms, 148 CTAs %.3f ms\n", shift, n_slices, tm, tr, tz, tz1);
        for (u32 nzf : {1u, 2u, 3u})
            for (u32 ahead : {1u, 2u})
                for (int fence : {1}) {
                    const u32 grid = 148 * std::min(per_sm, 8), n_zero = 148 * nzf;
                    u64 nn = n; u64 bb = bins; u32 sh = shift, ns = n_slices, nz = n_zero, ah = ahead; int df = fence;
                    void *args[] = {&f, &idx, &nn, &h, &bb, &sh, &ns, &nz, &ah, &df};
                    float t = timeit([&] { cudaMemsetAsync(f, 0, sizeof(Flags)); cudaLaunchCooperativeKernel((const void *)fused<12>, dim3(grid), dim3(256), args, 0, 0); });
                    printf("   fused: n_zero %4u ahead %u fence %d : %.3f ms\n", n_zero, ahead, fence, t);
                }
    }
    return 0;
}

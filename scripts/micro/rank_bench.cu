// Microbenchmark: ways to rank 32-bit keys into <= 512 buckets inside a CTA tile (the core of a multisplit).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef uint32_t u32; typedef uint64_t u64;
#define FULL 0xffffffffu
__device__ __forceinline__ u32 mix(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__global__ void fill(u32 *k, u64 n, u32 nb) { u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) k[i] = mix((u32)i) % nb; }
// (a) returning shared atomic per key
__global__ void rank_atoms(const u32 *__restrict__ keys, u64 n, u32 nb, u32 *out)
{
    __shared__ u32 cnt[512];
    for (u32 b = threadIdx.x; b < 512; b += blockDim.x) cnt[b] = 0;
    __syncthreads();
    u32 acc = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += atomicAdd(&cnt[__ldcs(keys + i)], 1u);
    if (acc == 0x12345678u) out[0] = acc;
}
// (b) non-returning shared atomic per key (count only)
__global__ void count_reds(const u32 *__restrict__ keys, u64 n, u32 nb, u32 *out)
{
    __shared__ u32 cnt[512];
    for (u32 b = threadIdx.x; b < 512; b += blockDim.x) cnt[b] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) atomicAdd(&cnt[__ldcs(keys + i)], 1u);
    __syncthreads();
    if (threadIdx.x < 512 && cnt[threadIdx.x] == 0x12345678u) out[0] = 1;
}
// (c) MATCH.ANY + warp-private counter read-modify-write by the group leader
__global__ void rank_match(const u32 *__restrict__ keys, u64 n, u32 nb, u32 *out)
{
    __shared__ u32 cnt[8][512];
    for (u32 b = threadIdx.x; b < 8 * 512; b += blockDim.x) (&cnt[0][0])[b] = 0;
    __syncthreads();
    u32 acc = 0;
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 n_it = (n + stride - 1) / stride;
    for (u64 it = 0; it < n_it; ++it) {
        const u64 i = it * stride + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        const bool act = i < n;
        const u32 am = __ballot_sync(FULL, act);
        if (act) {
            const u32 k = __ldcs(keys + i);
            const u32 peers = __match_any_sync(am, k);
            const int leader = __ffs(peers) - 1;
            u32 old = 0;
            if ((int)lane == leader) { old = cnt[w][k]; cnt[w][k] = old + __popc(peers); }
            old = __shfl_sync(peers, old, leader);
            acc += old + __popc(peers & ((1u << lane) - 1));
        }
        __syncwarp();
    }
    if (acc == 0x12345678u) out[0] = acc;
}
// (d) MATCH only
__global__ void match_only(const u32 *__restrict__ keys, u64 n, u32 nb, u32 *out)
{
    u32 acc = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 n_it = (n + stride - 1) / stride;
    for (u64 it = 0; it < n_it; ++it) {
        const u64 i = it * stride + (u64)blockIdx.x * blockDim.x + threadIdx.x;
        const u32 k = i < n ? __ldcs(keys + i) : 0;
        acc += __match_any_sync(FULL, k);
    }
    if (acc == 0x12345678u) out[0] = acc;
}
// (e) plain streaming read (upper bound)
__global__ void read_only(const u32 *__restrict__ keys, u64 n, u32 nb, u32 *out)
{
    u32 acc = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += __ldcs(keys + i);
    if (acc == 0x12345678u) out[0] = acc;
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
    const u64 n = 1ull << 29;
    u32 *k, *out; cudaMalloc(&k, n * 4); cudaMalloc(&out, 64);
    for (u32 nb : {16u, 416u}) {
        fill<<<(unsigned)((n + 255) / 256), 256>>>(k, n, nb);
        cudaDeviceSynchronize();
        for (int g : {148 * 4, 148 * 8}) {
            float a = timeit([&] { rank_atoms<<<g, 256>>>(k, n, nb, out); });
            float b = timeit([&] { count_reds<<<g, 256>>>(k, n, nb, out); });
            float c = timeit([&] { rank_match<<<g, 256>>>(k, n, nb, out); });
            float d = timeit([&] { match_only<<<g, 256>>>(k, n, nb, out); });
            float e = timeit([&] { read_only<<<g, 256>>>(k, n, nb, out); });
            printf("%3u buckets, grid %4d: ATOMS(return) %.3f ms (%.0f G/s) | RED.shared %.3f ms (%.0f G/s) | MATCH+rmw %.3f ms (%.0f G/s) | MATCH only %.3f ms (%.0f G/s) | read %.3f ms\n",
                   nb, g, a, n / a / 1e6, b, n / b / 1e6, c, n / c / 1e6, d, n / d / 1e6, e);
        }
    }
    return 0;
}

// kernels.cuh - hand-written sm_100a kernels of the SLIMM profiling hot path.
//
// All of this is HBM- / L2-atomic-bound integer work (no dense contraction, so no tensor cores):
// the design rules are coalesced streaming of the record SoA, warp-level run analysis instead of
// per-thread rescans, L2-resident scatter targets, and as few same-address atomics as possible.
//
// Data layout in HBM (DESIGN.md has the long form):
//   records   : struct-of-arrays  read_id[N] u32 | ref_id[N] u32 | begin_pos[N] i32, in file order,
//               non-decreasing in read_id (or, after the device sort of unsorted input, read_id[N] +
//               packed {ref,pos}[N] u64)
//   ref_meta  : uint4[G] = {len, nb = len/w+1, bin offset lo, hi}; offsets padded to 64 bins so every
//               512-byte warp step of the stats kernel belongs to one reference
//   hist      : u64[Bp]  = {lo: cov bin, hi: uniq_cov bin} interleaved: one 64-bit RED per (read, ref) pair
//   items     : u32[N]   bucketed scatter stream: padded bin index | uniq << 31 (0xFFFFFFFF = repeat hit)
//   cov2      : u32[Bp]  uniq_cov2 (only with SLIMM_GPU_KEEP_UNIQ_COV2)
//   stats     : u32[G*4] = {nz, reads_count, uniq nz, uniq_reads_count}
//   assign    : u32[(17+T)*G] = uniq_reads_count2[G] | lca_count[G*8] | child_mark[G*8] | fb_mark[T*G]
//               keyed by (reference, lineage level) instead of taxon id, so no device hash map
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;

#define FULL 0xffffffffu
#define ITEM_SKIP 0xFFFFFFFFu
#define LCA_REPLICAS 16
#define MAX_BUCKETS 1024

struct DevScalars {
    unsigned long long n_reads;   // matches_count   (partial per rank)      } summed across ranks
    unsigned long long n_uniq;    // uniq_matches_count                      } by the caller
    unsigned long long n_uniq2;   // uniq_matches_count2
    unsigned long long n_pairs;   // sum of reads_count
    u32 flags;                    // bit0: read ids not non-decreasing, bit1: ref_id >= G
    u32 n_valid, failed_cov, failed_ucov, failed_minread, ref_count;
    float cut, ucut;
    u32 done_ctr;
    u32 pad;
};

// record accessors: plain SoA, or {read_id[], packed (ref | pos<<32)[]} after the device sort
struct RecSoA {
    const u32 *rid; const u32 *ref; const i32 *pos;
    __device__ __forceinline__ u32 read(u64 i) const { return __ldg(rid + i); }
    __device__ __forceinline__ u32 refid(u64 i) const { return __ldg(ref + i); }
    __device__ __forceinline__ u32 upos(u64 i) const { return (u32)__ldg(pos + i); }
};
struct RecPacked {
    const u32 *rid; const uint2 *rp;
    __device__ __forceinline__ u32 read(u64 i) const { return __ldg(rid + i); }
    __device__ __forceinline__ u32 refid(u64 i) const { return __ldg(&rp[i].x); }
    __device__ __forceinline__ u32 upos(u64 i) const { return __ldg(&rp[i].y); }
};

__device__ __forceinline__ u32 warp_sum(u32 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ u32 warp_or(u32 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(FULL, v, o);
    return v;
}

// one atomic per distinct key per warp; must be reached by all 32 lanes
__device__ __forceinline__ void warp_agg_add(u32 *base, u32 key, bool active)
{
    unsigned act = __ballot_sync(FULL, active);
    if (active) {
        unsigned peers = __match_any_sync(act, key);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(base + key, (u32)__popc(peers));
    }
}

// padded global bin index of a record (reference src/slimm.hpp:200-201): u32 wrap of beginPos + avg/2,
// clamp to the contig length, integer division by the bin width
__device__ __forceinline__ u64 bin_of(const uint4 *__restrict__ meta, u32 g, u32 upos, u32 half_avg, u32 w)
{
    const uint4 m = __ldg(meta + g);   // {len, nb, off_lo, off_hi}
    const u32 center = min(upos + half_avg, m.x);
    return (((u64)m.w << 32) | m.z) + center / w;
}

// ------------------------------------------------------------------------------------------------
// Run analysis for one warp-row of 32 consecutive records.  A read is a run of equal read_id.  For
// its record each lane learns
//   head  : first record of its read,
//   first : first record of its (read, ref) pair in file order - only these contribute
//           (repeat hits are dropped, reference src/read_stat.hpp:125-131),
//   multi : the read names another reference as well (not unique, src/read_stat.hpp:72-75).
// Runs inside the row are resolved with two MATCH.ANY; only the runs touching the row's two edges
// look at neighbouring rows.  bad |= 1 when read ids decrease (input not grouped by read).
// ------------------------------------------------------------------------------------------------
template <class Rec>
__device__ __forceinline__ void analyze_row(const Rec &rec, u64 row0, u64 n, bool active, u32 r, u32 g,
                                            bool &head, bool &first, bool &multi, u32 &bad)
{
    const u32 lane = threadIdx.x & 31;
    const unsigned act = __ballot_sync(FULL, active);
    const u32 n_act = __popc(act);                      // active lanes are 0..n_act-1
    u32 edge_lo = 0, edge_hi = 0;                       // read ids just before / after this row
    bool has_lo = false, has_hi = false;
    if (lane == 0 && row0 > 0 && row0 < n) { edge_lo = rec.read(row0 - 1); has_lo = true; }
    if (lane == 0 && row0 + n_act < n) { edge_hi = rec.read(row0 + n_act); has_hi = true; }
    edge_lo = __shfl_sync(FULL, edge_lo, 0); edge_hi = __shfl_sync(FULL, edge_hi, 0);
    has_lo = __shfl_sync(FULL, (int)has_lo, 0); has_hi = __shfl_sync(FULL, (int)has_hi, 0);
    u32 prev = __shfl_up_sync(FULL, r, 1);
    if (lane == 0) prev = edge_lo;
    head = first = multi = false;
    if (!active) return;
    if ((lane > 0 || has_lo) && prev > r) bad |= 1u;
    const unsigned M = __match_any_sync(act, r);
    const unsigned P = __match_any_sync(act, ((unsigned long long)r << 32) | g);
    head = lane == (u32)(__ffs(M) - 1);
    first = lane == (u32)(__ffs(P) - 1);
    multi = P != M;
    if ((M & 1u) && has_lo && edge_lo == r) {           // my run started in an earlier row
        head = false;
        u64 j = row0;
        while (first && j > 0 && rec.read(j - 1) == r) {
            --j;
            if (rec.refid(j) == g) { first = false; break; }
            multi = true;
        }
    }
    if (first && !multi && (M >> (n_act - 1)) && has_hi && edge_hi == r) {   // my run continues past this row
        u64 j = row0 + n_act;
        while (j < n && rec.read(j) == r) {
            if (rec.refid(j) != g) { multi = true; break; }
            ++j;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1 coverage.  Replaces reference src/slimm.hpp:194-257 + src/read_stat.hpp:116-135.
//   MODE 0 (direct) : one 64-bit RED per contributing record straight into hist (histogram fits in L2)
//   MODE 1 (emit)   : every record emits one 32-bit item into its bin-range bucket (block-level
//                     multisplit through shared memory); k_accumulate applies the buckets one L2-sized
//                     histogram slice after the other, so the random read-modify-writes never reach HBM
// ------------------------------------------------------------------------------------------------
#define COV_ROWS 8   // rows of 256 records per block tile
template <class Rec, int MODE>
__global__ void __launch_bounds__(256)
k_coverage(Rec rec, u64 n, const uint4 *__restrict__ meta, u32 G, u32 half_avg, u32 w,
           unsigned long long *__restrict__ hist, u32 *__restrict__ items, u32 *__restrict__ cursor, u32 shift,
           u32 n_buckets, DevScalars *sc)
{
    __shared__ u32 s_cnt[MODE ? MAX_BUCKETS : 1], s_base[MODE ? MAX_BUCKETS : 1];
    __shared__ u32 s_h, s_u, s_b;
    const u32 tid = threadIdx.x;
    if (MODE) for (u32 b = tid; b < n_buckets; b += 256) s_cnt[b] = 0;
    if (tid == 0) { s_h = 0; s_u = 0; s_b = 0; }
    __syncthreads();
    u32 heads = 0, uniq = 0, bad = 0;
    const u64 tile = 256ull * COV_ROWS;
    for (u64 t0 = (u64)blockIdx.x * tile; t0 < n; t0 += (u64)gridDim.x * tile) {   // block-uniform trip count
        u32 item[COV_ROWS], rank[COV_ROWS];
#pragma unroll
        for (int k = 0; k < COV_ROWS; ++k) {
            const u64 row0 = t0 + (u64)k * 256 + (tid & ~31u);
            const u64 i = row0 + (tid & 31);
            const bool active = i < n;
            u32 r = 0, g = 0;
            if (active) { r = rec.read(i); g = rec.refid(i); }
            bool head, first, multi;
            analyze_row(rec, row0, n, active, r, g, head, first, multi, bad);
            item[k] = ITEM_SKIP; rank[k] = 0xFFFFFFFFu;   // rank: bucket << 16 | slot inside this tile's share
            if (active) {
                if (g >= G) { bad |= 2u; }
                else {
                    if (head) { ++heads; uniq += !multi; }
                    const u64 b = bin_of(meta, g, rec.upos(i), half_avg, w);
                    if (MODE == 0) {
                        if (first) atomicAdd(hist + b, multi ? 1ull : 0x100000001ull);   // cov += 1 [, uniq_cov += 1]
                    } else {
                        const u32 bucket = (u32)(b >> shift);
                        rank[k] = (bucket << 16) | atomicAdd(&s_cnt[bucket], 1u);       // tile holds < 2^16 records
                        if (first) item[k] = (u32)b | (multi ? 0u : 0x80000000u);
                    }
                }
            }
        }
        if (MODE) {
            __syncthreads();
            for (u32 b = tid; b < n_buckets; b += 256) {
                const u32 c = s_cnt[b];
                if (c) { s_base[b] = atomicAdd(cursor + b, c); s_cnt[b] = 0; }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < COV_ROWS; ++k)
                if (rank[k] != 0xFFFFFFFFu) items[s_base[rank[k] >> 16] + (rank[k] & 0xFFFFu)] = item[k];
            // no barrier needed here: s_base is rewritten only after the next tile's first barrier
        }
    }
    heads = warp_sum(heads); uniq = warp_sum(uniq); bad = warp_or(bad);
    if ((tid & 31) == 0) { atomicAdd(&s_h, heads); atomicAdd(&s_u, uniq); if (bad) atomicOr(&s_b, bad); }
    __syncthreads();
    if (tid == 0) {
        if (s_h) atomicAdd(&sc->n_reads, (unsigned long long)s_h);
        if (s_u) atomicAdd(&sc->n_uniq, (unsigned long long)s_u);
        if (s_b) atomicOr(&sc->flags, s_b);
    }
}

// bucket sizes for the multisplit: every record lands in the bucket of its bin
template <class Rec>
__global__ void __launch_bounds__(256)
k_bucket_count(Rec rec, u64 n, const uint4 *__restrict__ meta, u32 G, u32 half_avg, u32 w, u32 shift, u32 n_buckets,
               u32 *__restrict__ bucket_cnt, DevScalars *sc)
{
    __shared__ u32 s_cnt[MAX_BUCKETS];
    for (u32 b = threadIdx.x; b < n_buckets; b += 256) s_cnt[b] = 0;
    __syncthreads();
    u32 bad = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u32 g = rec.refid(i);
        if (g >= G) { bad = 2u; continue; }
        atomicAdd(&s_cnt[(u32)(bin_of(meta, g, rec.upos(i), half_avg, w) >> shift)], 1u);
    }
    __syncthreads();
    for (u32 b = threadIdx.x; b < n_buckets; b += 256)
        if (s_cnt[b]) atomicAdd(bucket_cnt + b, s_cnt[b]);
    if (bad) atomicOr(&sc->flags, bad);
}

// exclusive scan of the bucket sizes -> write cursors (one block; n_buckets <= 1024)
__global__ void __launch_bounds__(1024) k_bucket_scan(const u32 *__restrict__ bucket_cnt, u32 n_buckets, u32 *__restrict__ cursor)
{
    __shared__ u32 s[1024];
    const u32 tid = threadIdx.x;
    const u32 v = tid < n_buckets ? bucket_cnt[tid] : 0;
    s[tid] = v;
    __syncthreads();
    for (u32 d = 1; d < 1024; d <<= 1) {
        const u32 t = tid >= d ? s[tid - d] : 0;
        __syncthreads();
        s[tid] += t;
        __syncthreads();
    }
    if (tid < n_buckets) cursor[tid] = s[tid] - v;
}

// apply the bucketed items in stream order: the blocks in flight work on one or two adjacent
// L2-resident histogram slices
__global__ void __launch_bounds__(256)
k_accumulate(const uint4 *__restrict__ items4, u64 n_items, unsigned long long *__restrict__ hist)
{
    const u64 n4 = n_items >> 2;
    const u64 chunk = 256ull * 4;                       // uint4 per block iteration
    const u64 base = (u64)blockIdx.x * chunk;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const u64 j = base + (u64)k * 256 + threadIdx.x;
        if (j < n4) {
            const uint4 v = __ldcs(items4 + j);
            if (v.x != ITEM_SKIP) atomicAdd(hist + (v.x & 0x7FFFFFFFu), (v.x >> 31) ? 0x100000001ull : 1ull);
            if (v.y != ITEM_SKIP) atomicAdd(hist + (v.y & 0x7FFFFFFFu), (v.y >> 31) ? 0x100000001ull : 1ull);
            if (v.z != ITEM_SKIP) atomicAdd(hist + (v.z & 0x7FFFFFFFu), (v.z >> 31) ? 0x100000001ull : 1ull);
            if (v.w != ITEM_SKIP) atomicAdd(hist + (v.w & 0x7FFFFFFFu), (v.w >> 31) ? 0x100000001ull : 1ull);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n_items & 3)) {   // tail
        const u32 v = reinterpret_cast<const u32 *>(items4)[n4 * 4 + threadIdx.x];
        if (v != ITEM_SKIP) atomicAdd(hist + (v & 0x7FFFFFFFu), (v >> 31) ? 0x100000001ull : 1ull);
    }
}

// ------------------------------------------------------------------------------------------------
// K3: per-reference segmented reduction over the interleaved bins.  Replaces
// bins_coverage::none_zero_bin_count (src/reference_contig.hpp:84-91) for cov and uniq_cov and
// recovers reads_count / uniq_reads_count as the bin sums (each pair adds 1 to exactly one bin).
// A warp step is 32 x 16 B = 64 bins; segments are padded to 64 bins (padding stays zero).
// ------------------------------------------------------------------------------------------------
#define STATS_STEPS_PER_WARP 16
__global__ void __launch_bounds__(256)
k_ref_stats(const uint4 *__restrict__ hist4, u64 n_steps, const u64 *__restrict__ off /*[G+1] padded, in bins*/,
            u32 G, u32 *__restrict__ stats)
{
    const u32 lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 chunk = warp; chunk * STATS_STEPS_PER_WARP < n_steps; chunk += n_warps) {
        const u64 s0 = chunk * STATS_STEPS_PER_WARP;
        const u64 s1 = min(s0 + (u64)STATS_STEPS_PER_WARP, n_steps);
        u32 lo = 0, hi = G;                              // largest g with off[g] <= first bin of the chunk
        const u64 bin0 = s0 * 64;
        while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (__ldg(off + mid) <= bin0) lo = mid; else hi = mid; }
        u32 g = lo;
        u64 g_end = __ldg(off + g + 1);
        u32 nz = 0, sum = 0, unz = 0, usum = 0;
        for (u64 s = s0; s < s1; ++s) {
            if (s * 64 >= g_end) {
                nz = warp_sum(nz); sum = warp_sum(sum); unz = warp_sum(unz); usum = warp_sum(usum);
                if (lane == 0 && (nz | unz)) {
                    atomicAdd(stats + 4 * g + 0, nz); atomicAdd(stats + 4 * g + 1, sum);
                    if (unz) { atomicAdd(stats + 4 * g + 2, unz); atomicAdd(stats + 4 * g + 3, usum); }
                }
                nz = sum = unz = usum = 0;
                while (s * 64 >= g_end) { ++g; g_end = __ldg(off + g + 1); }
            }
            const uint4 v = __ldcs(hist4 + s * 32 + lane);   // {cov0, ucov0, cov1, ucov1}
            nz += (v.x != 0) + (v.z != 0); sum += v.x + v.z;
            unz += (v.y != 0) + (v.w != 0); usum += v.y + v.w;
        }
        nz = warp_sum(nz); sum = warp_sum(sum); unz = warp_sum(unz); usum = warp_sum(usum);
        if (lane == 0 && (nz | unz)) {
            atomicAdd(stats + 4 * g + 0, nz); atomicAdd(stats + 4 * g + 1, sum);
            if (unz) { atomicAdd(stats + 4 * g + 2, unz); atomicAdd(stats + 4 * g + 3, usum); }
        }
    }
}

// nonzero uniq_cov2 bins per reference (raw output only): one warp per reference
__global__ void k_cov2_nz(const u32 *__restrict__ cov2, const u64 *__restrict__ off, const uint4 *__restrict__ meta,
                          u32 G, u32 *__restrict__ out)
{
    const u32 lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= G) return;
    const u64 a = off[warp];
    const u32 nb = meta[warp].y;
    u32 nz = 0;
    for (u32 b = lane; b < nb; b += 32) nz += cov2[a + b] != 0;
    nz = warp_sum(nz);
    if (lane == 0) out[warp] = nz;
}

// uniq_cov2 starts as uniq_cov of the surviving references (a read with one target whose reference
// survives stays unique); k_assign adds the reads that BECAME unique.  One warp step = 64 bins.
__global__ void __launch_bounds__(256)
k_cov2_base(const uint4 *__restrict__ hist4, u64 n_steps, const u64 *__restrict__ off, u32 G,
            const u32 *__restrict__ valid_bits, uint2 *__restrict__ cov2_2)
{
    const u32 lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 s = warp; s < n_steps; s += n_warps) {
        u32 lo = 0, hi = G;
        const u64 bin0 = s * 64;
        while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (__ldg(off + mid) <= bin0) lo = mid; else hi = mid; }
        uint2 o = make_uint2(0, 0);
        if ((__ldg(valid_bits + (lo >> 5)) >> (lo & 31)) & 1u) {
            const uint4 v = __ldg(hist4 + s * 32 + lane);
            o = make_uint2(v.y, v.w);
        }
        cov2_2[s * 32 + lane] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// K4: exact-order quantile cut-offs + valid mask.  Replaces coverage_cut_off /
// uniq_coverage_cut_off (src/slimm.hpp:328-344,672-688), get_quantile_cut_off (src/misc.hpp:197-216)
// and the reference loop of filter_alignments (src/slimm.hpp:354-378).
// grid = 2 CTAs (cov, uniq_cov) x 1024 threads.  The f32 folds are sequential in one thread on
// purpose - the surviving set must be bit-exact and depends on every rounding - but they run out of
// shared memory, staged 8192 values at a time by the whole CTA.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float f32_up(float x) { return __uint_as_float(__float_as_uint(x) + 1u); }   // x > 0
__device__ __forceinline__ float f32_down(float x) { return __uint_as_float(__float_as_uint(x) - 1u); } // x > 0
#define CUT_CHUNK 8192

__global__ void __launch_bounds__(1024)
k_cutoffs(const u32 *__restrict__ stats, const uint4 *__restrict__ meta, u32 G, float q, u32 min_reads,
          float *__restrict__ cp_all /*[2][G]*/, u32 *__restrict__ scratch /*[2][npow2]*/, u32 npow2,
          u32 *__restrict__ valid_bits, unsigned char *__restrict__ valid_bytes, DevScalars *sc)
{
    const u32 which = blockIdx.x;             // 0: cov, 1: uniq_cov
    if (min_reads == 0) {                     // -mr default: 1 + (matches_count-1)/10000 (src/slimm.hpp:458-459)
        const u32 R = (u32)sc->n_reads;
        min_reads = R ? 1u + (R - 1u) / 10000u : 0u;
    }
    const u32 tid = threadIdx.x;
    float *cp = cp_all + (size_t)which * G;
    u32 *v = scratch + (size_t)which * npow2;
    __shared__ u32 s_scan[1024];
    __shared__ float s_buf[CUT_CHUNK];
    __shared__ u32 s_base, s_i;
    __shared__ float s_f;
    __shared__ int s_done;
    __shared__ bool s_last;

    // cov_percent = float(nz) / number_of_bins (src/reference_contig.hpp:148-155)
    for (u32 g = tid; g < G; g += 1024)
        cp[g] = __fdiv_rn((float)stats[4 * g + 2 * which], (float)meta[g].y);
    if (tid == 0) s_base = 0;
    __syncthreads();
    float cut = 0.0f;
    if (q < 1.0f) {
        // ordered compaction of cp[g] over references with unique reads (ascending g)
        for (u32 g0 = 0; g0 < G; g0 += 1024) {
            const u32 g = g0 + tid;
            const u32 keep = (g < G && stats[4 * g + 3] > 0) ? 1u : 0u;
            s_scan[tid] = keep;
            __syncthreads();
            for (u32 d = 1; d < 1024; d <<= 1) {
                u32 t = tid >= d ? s_scan[tid - d] : 0;
                __syncthreads();
                s_scan[tid] += t;
                __syncthreads();
            }
            if (keep) v[s_base + s_scan[tid] - 1] = __float_as_uint(cp[g]);
            __syncthreads();
            if (tid == 1023) s_base += s_scan[1023];
            __syncthreads();
        }
        const u32 n = s_base;
        // total = std::accumulate(v, 0.0f): left fold in reference order
        if (tid == 0) s_f = 0.0f;
        for (u32 c0 = 0; c0 < n; c0 += CUT_CHUNK) {
            const u32 cn = min((u32)CUT_CHUNK, n - c0);
            for (u32 k = tid; k < cn; k += 1024) s_buf[k] = __uint_as_float(v[c0 + k]);
            __syncthreads();
            if (tid == 0) {
                float total = s_f;
                for (u32 k = 0; k < cn; ++k) total = __fadd_rn(total, s_buf[k]);
                s_f = total;
            }
            __syncthreads();
        }
        const float total = s_f;
        // pad to a power of two and sort ascending (values are >= 0: u32 order == f32 order)
        u32 m = 1;
        while (m < n) m <<= 1;
        for (u32 i = n + tid; i < m; i += 1024) v[i] = 0xFFFFFFFFu;
        __syncthreads();
        for (u32 k = 2; k <= m; k <<= 1)
            for (u32 j = k >> 1; j > 0; j >>= 1) {
                for (u32 t = tid; t < m; t += 1024) {
                    u32 p = t ^ j;
                    if (p > t) {
                        u32 a = v[t], b = v[p];
                        bool up = (t & k) == 0;
                        if ((a > b) == up) { v[t] = b; v[p] = a; }
                    }
                }
                __syncthreads();
            }
        // i = n-1; while ((sub/total) < q && i > 0) { sub += v[i]; --i; }  cutoff = v[i]
        if (tid == 0) {
            s_done = (n == 0 || !(total > 0.0f) || !(q > 0.0f));   // 0/0 = NaN: NaN < q is false; q <= 0: never true
            s_i = n ? n - 1 : 0;
            float s = 0.0f;
            if (!s_done) {
                // (sub/total) < q  <=>  sub < s*, s* = smallest f32 with fl(s*/total) >= q
                // (x -> fl(x/total) is monotone), so the loop needs no division
                s = __fmul_rn(q, total);
                if (s <= 0.0f) s = __uint_as_float(1u);
                while (__fdiv_rn(s, total) >= q && s > __uint_as_float(1u)) s = f32_down(s);
                while (__fdiv_rn(s, total) < q) s = f32_up(s);
            }
            s_f = s;
        }
        __syncthreads();
        const float sstar = s_f;
        float sub = 0.0f;                                           // only thread 0's copy matters
        u32 c_hi = n;                                               // values [c_lo, c_hi) staged, walked downwards
        while (!s_done) {
            const u32 c_lo = c_hi > CUT_CHUNK ? c_hi - CUT_CHUNK : 0;
            for (u32 k = tid; k < c_hi - c_lo; k += 1024) s_buf[k] = __uint_as_float(v[c_lo + k]);
            __syncthreads();
            if (tid == 0) {
                u32 i = s_i;
                while (sub < sstar && i > 0 && i >= c_lo) {
                    sub = __fadd_rn(sub, s_buf[i - c_lo]);
                    --i;
                    if (i < c_lo) break;
                }
                s_i = i;
                if (!(sub < sstar) || i == 0 || c_lo == 0) s_done = 1;
            }
            __syncthreads();
            c_hi = c_lo;
        }
        if (n > 0) cut = __uint_as_float(v[s_i]);
    }
    if (tid == 0) {
        if (which == 0) sc->cut = cut; else sc->ucut = cut;
        __threadfence();
        s_last = atomicAdd(&sc->done_ctr, 1u) == 1u;
    }
    __syncthreads();
    if (!s_last) return;
    // last CTA: valid set + -v counters (src/slimm.hpp:354-378)
    __threadfence();
    const float c0 = *(volatile float *)&sc->cut, c1 = *(volatile float *)&sc->ucut;
    const float *cpa = cp_all, *ucpa = cp_all + G;
    u32 nv = 0, fc = 0, fu = 0, fm = 0, rc = 0;
    unsigned long long pairs = 0;
    for (u32 g0 = 0; g0 < G; g0 += 1024) {
        const u32 g = g0 + tid;
        bool ok = false;
        if (g < G) {
            const u32 reads = stats[4 * g + 1];
            if (reads > 0) {
                ++rc; pairs += reads;
                const float a = __ldcg(cpa + g), b = __ldcg(ucpa + g);
                ok = a >= c0 && b >= c1;
                if (ok) ++nv;
                else { fu += b < c1; fm += reads < min_reads; fc += a < c0; }
            }
            valid_bytes[g] = ok;
        }
        const u32 word = __ballot_sync(FULL, ok);
        if ((tid & 31) == 0 && g < G) valid_bits[g >> 5] = word;
    }
    nv = warp_sum(nv); fc = warp_sum(fc); fu = warp_sum(fu); fm = warp_sum(fm); rc = warp_sum(rc);
    pairs = warp_sum64(pairs);
    if ((tid & 31) == 0) {
        atomicAdd(&sc->n_valid, nv); atomicAdd(&sc->failed_cov, fc); atomicAdd(&sc->failed_ucov, fu);
        atomicAdd(&sc->failed_minread, fm); atomicAdd(&sc->ref_count, rc);
        atomicAdd(&sc->n_pairs, pairs);
    }
}

// ------------------------------------------------------------------------------------------------
// K5+K6: reassignment + LCA.  The head record of every read with >= 2 records walks its run
// (single-record reads are unique reads: their contribution to uniq_reads_count2 is
// valid[g] * uniq_reads_count[g], added by k_finish_assign without touching the records again).
// Replaces the read loop of filter_alignments (src/slimm.hpp:380-391, read_stat::update
// src/read_stat.hpp:98-114), slimm::get_lca (src/slimm.hpp:516-531) and phase 1 of
// get_reads_lca_count (:536-557).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_valid(const u32 *__restrict__ vb, u32 g) { return (__ldg(vb + (g >> 5)) >> (g & 31)) & 1u; }

template <class Rec>
__global__ void __launch_bounds__(256)
k_assign(Rec rec, u64 n, const uint4 *__restrict__ meta, const uint4 *__restrict__ lin4, const u32 *__restrict__ top_idx,
         const u32 *__restrict__ vb, u32 G, u32 half_avg, u32 w, u32 *__restrict__ uniq2_extra, u32 *__restrict__ lca_rep,
         u32 *__restrict__ child_mark, u32 *__restrict__ fb_mark, u32 *__restrict__ cov2,
         unsigned char *__restrict__ res_kind, u32 *__restrict__ res_val)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    u32 *lca_cnt = lca_rep + (u64)(blockIdx.x % LCA_REPLICAS) * 8 * G;
    const u32 lane = threadIdx.x & 31;
    for (u64 base = (u64)blockIdx.x * blockDim.x; base < n; base += stride) {   // block-uniform trip count
        const u64 i = base + threadIdx.x;
        const bool active = i < n;
        u32 r = active ? rec.read(i) : 0;
        u32 prev = __shfl_up_sync(FULL, r, 1);
        u32 next = __shfl_down_sync(FULL, r, 1);
        bool is_lca = false;
        u32 key = 0;
        if (active) {
            if (lane == 0) prev = i > 0 ? rec.read(i - 1) : ~r;
            if (lane == 31 || i + 1 >= n) next = i + 1 < n ? rec.read(i + 1) : ~r;
            const bool head = i == 0 || prev != r;
            if (head && next != r) {                   // single-record read
                if (res_kind) {
                    const u32 g = rec.refid(i);
                    if (g < G && is_valid(vb, g)) { res_kind[i] = 1; res_val[i] = g; }
                }
            } else if (head) {
                const u32 ref0 = rec.refid(i);
                u32 g0 = 0xFFFFFFFFu, gmax = 0, eq = 0xFFu;
                u64 lead = i;
                uint4 la = make_uint4(0, 0, 0, 0), lb = la;
                bool multi = false, was_multi = false;
                u64 j = i;
                do {
                    const u32 h = rec.refid(j);
                    was_multi |= h != ref0;
                    if (h < G && is_valid(vb, h)) {
                        if (g0 == 0xFFFFFFFFu) {
                            g0 = gmax = h; lead = j;
                            la = __ldg(lin4 + 2 * (u64)h); lb = __ldg(lin4 + 2 * (u64)h + 1);
                        } else if (h != g0) {
                            multi = true;
                            gmax = max(gmax, h);
                            const uint4 ha = __ldg(lin4 + 2 * (u64)h), hb = __ldg(lin4 + 2 * (u64)h + 1);
                            eq &= (ha.x == la.x) | ((ha.y == la.y) << 1) | ((ha.z == la.z) << 2) | ((ha.w == la.w) << 3) |
                                  ((hb.x == lb.x) << 4) | ((hb.y == lb.y) << 5) | ((hb.z == lb.z) << 6) | ((hb.w == lb.w) << 7);
                        }
                    }
                    ++j;
                } while (j < n && rec.read(j) == r);
                const u64 run_end = j;
                if (g0 != 0xFFFFFFFFu && !multi) {     // sole survivor (:383-390)
                    if (was_multi) {                   // the read BECAME unique through the filter
                        atomicAdd(uniq2_extra + g0, 1u);
                        if (cov2) atomicAdd(cov2 + bin_of(meta, g0, rec.upos(lead), half_avg, w), 1u);
                    }
                    if (res_kind) { res_kind[i] = 1; res_val[i] = g0; }
                } else if (g0 != 0xFFFFFFFFu) {        // level-wise LCA over 8-slot lineages, zeros included
                    const bool fb = eq == 0;           // no level agrees: slot 7 of the largest reference id
                    const u32 level = fb ? 7 : __ffs(eq) - 1, owner = fb ? gmax : g0;
                    is_lca = true;
                    key = owner * 8 + level;
                    const u32 trow = fb ? __ldg(top_idx + gmax) : 0;
                    for (u64 k = lead; k < run_end; ++k) {              // children[lca] U= S (:555)
                        const u32 h = rec.refid(k);
                        if (h < G && is_valid(vb, h)) {
                            u32 *mk = fb ? fb_mark + (u64)trow * G + h : child_mark + (u64)h * 8 + level;
                            if (*mk == 0) *mk = 1;
                        }
                    }
                    if (res_kind) {
                        res_kind[i] = 2;
                        res_val[i] = __ldg(reinterpret_cast<const u32 *>(lin4) + (u64)owner * 8 + level);
                    }
                }
            }
        }
        warp_agg_add(lca_cnt, key, is_lca);
    }
}

// lca_count[s] = sum of the replicas (taken apart only to spread same-address atomics)
__global__ void k_fold_lca(const u32 *__restrict__ lca_rep, u32 n_slots, u32 *__restrict__ lca_cnt)
{
    const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    u32 v = 0;
#pragma unroll
    for (int r = 0; r < LCA_REPLICAS; ++r) v += lca_rep[(u64)r * n_slots + s];
    lca_cnt[s] = v;
}

// uniq_reads_count2[g] = (valid[g] ? uniq_reads_count[g] : 0) + reads that became unique;
// uniq_matches_count2 = sum.  Runs once, after the per-rank partials have been summed.
__global__ void k_finish_assign(u32 *__restrict__ uniq2, const u32 *__restrict__ stats, const u32 *__restrict__ vb, u32 G,
                                DevScalars *sc)
{
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    u32 v = 0;
    if (g < G) {
        v = uniq2[g] + (is_valid(vb, g) ? stats[4 * g + 3] : 0u);
        uniq2[g] = v;
    }
    unsigned long long s = warp_sum64(v);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&sc->n_uniq2, s);
}

// ------------------------------------------------------------------------------------------------
// helpers for the unsorted-input path and bin readout
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_values(const u32 *__restrict__ ref, const i32 *__restrict__ pos, u64 n, uint2 *__restrict__ out)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = make_uint2(ref[i], (u32)pos[i]);
}

__global__ void k_extract_bins(const u32 *__restrict__ src, u64 first, u32 stride_words, u32 word, u32 nb, u32 *__restrict__ out)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) out[b] = src[(first + b) * stride_words + word];
}

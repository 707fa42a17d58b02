// slimm_gpu.cu - sm_100a kernels + C ABI of the SLIMM profiling hot path (see include/slimm_gpu.h).
//
// Data layout in HBM (DESIGN.md has the long form):
//   records   : struct-of-arrays  read_id[N] u32 | ref_id[N] u32 | begin_pos[N] i32, in file order
//               (or, after the device sort of unsorted input, read_id[N] + packed {ref,pos}[N] u64)
//   ref_meta  : uint4[G] = {len, nb = len/w+1, bin offset lo, hi}; offsets are padded to 64 bins so
//               every 512-byte warp step of the stats kernel belongs to one reference
//   hist      : u64[Bp]  = {lo: cov bin, hi: uniq_cov bin} interleaved, one 64-bit RED per pair
//   cov2      : u32[Bp]  uniq_cov2 (only with SLIMM_GPU_KEEP_UNIQ_COV2)
//   stats     : u32[G*4] = {nz, reads_count, uniq nz, uniq_reads_count}
//   assign    : u32[(17+T)*G] = uniq_reads_count2[G] | lca_count[G*8] | child_mark[G*8] | fb_mark[T*G]
//               keyed by (reference, lineage level) instead of taxon id, so no device hash map
//
// Every kernel here is HBM / L2-atomic bound integer work: no tensor cores, no CPU fallback.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/slimm_gpu.h"

typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;

// ------------------------------------------------------------------------------------------------
// device-side scalars
// ------------------------------------------------------------------------------------------------
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
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one atomic per distinct key per warp; must be reached by all 32 lanes
__device__ __forceinline__ void warp_agg_add(u32 *base, u32 key, bool active)
{
    unsigned act = __ballot_sync(0xffffffffu, active);
    if (active) {
        unsigned peers = __match_any_sync(act, key);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(base + key, (u32)__popc(peers));
    }
}

// ------------------------------------------------------------------------------------------------
// K1: coverage.  One thread per record.  Replaces reference src/slimm.hpp:194-257 +
// src/read_stat.hpp:116-135: a record contributes iff it is the first of its (read, ref) pair in
// file order; a read is unique iff all its records name one reference.  Reads are runs of equal
// read_id (input is non-decreasing in read_id, checked here: bit0 of flags).
// ------------------------------------------------------------------------------------------------
template <class Rec>
__global__ void __launch_bounds__(256)
k_coverage(Rec rec, u64 n, const uint4 *__restrict__ meta, u32 G, u32 half_avg, u32 w,
           unsigned long long *__restrict__ hist, DevScalars *sc)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    u32 heads = 0, uniq = 0, bad = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u32 r = rec.read(i), g = rec.refid(i);
        if (g >= G) { bad |= 2u; continue; }
        bool head = true, first = true, multi = false;
        if (i > 0) {
            u32 pr = rec.read(i - 1);
            if (pr > r) bad |= 1u;
            head = pr != r;
        }
        if (!head) {                       // look back over the run for an earlier hit of (read, g)
            u64 j = i;
            while (j > 0 && rec.read(j - 1) == r) {
                --j;
                if (rec.refid(j) == g) { first = false; break; }
                multi = true;
            }
        }
        if (!first) continue;              // repeat hit: dropped (src/read_stat.hpp:125-131)
        if (!multi) {                      // look ahead: does the read name any other reference?
            u64 j = i + 1;
            while (j < n && rec.read(j) == r) {
                if (rec.refid(j) != g) { multi = true; break; }
                ++j;
            }
        }
        if (head) { ++heads; uniq += !multi; }
        const uint4 m = __ldg(meta + g);   // {len, nb, off_lo, off_hi}
        u32 center = rec.upos(i) + half_avg;          // u32 wrap as in src/slimm.hpp:200
        center = min(center, m.x);
        const u64 b = (((u64)m.w << 32) | m.z) + center / w;
        atomicAdd(hist + b, multi ? 1ull : 0x100000001ull);   // cov += 1 [, uniq_cov += 1]
    }
    heads = warp_sum(heads); uniq = warp_sum(uniq);
    bad |= __shfl_xor_sync(0xffffffffu, bad, 16); bad |= __shfl_xor_sync(0xffffffffu, bad, 8);
    bad |= __shfl_xor_sync(0xffffffffu, bad, 4);  bad |= __shfl_xor_sync(0xffffffffu, bad, 2);
    bad |= __shfl_xor_sync(0xffffffffu, bad, 1);
    __shared__ u32 s_h, s_u, s_b;
    if (threadIdx.x == 0) { s_h = 0; s_u = 0; s_b = 0; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_h, heads); atomicAdd(&s_u, uniq); if (bad) atomicOr(&s_b, bad); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_h) atomicAdd(&sc->n_reads, (unsigned long long)s_h);
        if (s_u) atomicAdd(&sc->n_uniq, (unsigned long long)s_u);
        if (s_b) atomicOr(&sc->flags, s_b);
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
        // reference owning bin s0*64: largest g with off[g] <= bin
        u32 lo = 0, hi = G;
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
            const uint4 v = __ldg(hist4 + s * 32 + lane);   // {cov0, ucov0, cov1, ucov1}
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

// ------------------------------------------------------------------------------------------------
// K4: exact-order quantile cut-offs + valid mask.  Replaces coverage_cut_off /
// uniq_coverage_cut_off (src/slimm.hpp:328-344,672-688), get_quantile_cut_off (src/misc.hpp:197-216)
// and the reference loop of filter_alignments (src/slimm.hpp:354-378).
// grid = 2 CTAs (cov, uniq_cov) x 1024 threads.  The f32 folds are sequential in one thread on
// purpose: the surviving set must be bit-exact and depends on every rounding.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float f32_up(float x) { return __uint_as_float(__float_as_uint(x) + 1u); }   // x > 0
__device__ __forceinline__ float f32_down(float x) { return __uint_as_float(__float_as_uint(x) - 1u); } // x > 0

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
    __shared__ u32 s_base, s_n;
    __shared__ float s_cut;
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
        // total = std::accumulate(v, 0.0f): left fold in reference order, one thread
        if (tid == 0) {
            float total = 0.0f;
            for (u32 i = 0; i < n; ++i) total = __fadd_rn(total, __uint_as_float(v[i]));
            s_cut = total;
            s_n = n;
        }
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
        if (tid == 0) {
            const float total = s_cut;
            float c = 0.0f;
            if (n > 0) {
                u32 i = n - 1;
                if (total > 0.0f && q > 0.0f) {      // q <= 0: (sub/total) < q is never true
                    // (sub/total) < q  <=>  sub < s*, s* = smallest f32 with fl(s*/total) >= q
                    // (x -> fl(x/total) is monotone), so the loop needs no division
                    float s = __fmul_rn(q, total);
                    if (s <= 0.0f) s = __uint_as_float(1u);
                    while (__fdiv_rn(s, total) >= q && s > __uint_as_float(1u)) s = f32_down(s);
                    while (__fdiv_rn(s, total) < q) s = f32_up(s);
                    float sub = 0.0f;
                    while (sub < s && i > 0) { sub = __fadd_rn(sub, __uint_as_float(v[i])); --i; }
                }   // total == 0: 0/0 is NaN, NaN < q is false, the reference loop is not entered
                c = __uint_as_float(v[i]);
            }
            s_cut = c;
        }
        __syncthreads();
        cut = s_cut;
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
        const u32 word = __ballot_sync(0xffffffffu, ok);
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
// K5+K6: reassignment + LCA.  One thread per record; the first record of a read whose reference
// survived the filter is the read's leader and walks the run.  Replaces the read loop of
// filter_alignments (src/slimm.hpp:380-391, read_stat::update src/read_stat.hpp:98-114),
// slimm::get_lca (src/slimm.hpp:516-531) and phase 1 of get_reads_lca_count (:536-557).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_valid(const u32 *__restrict__ vb, u32 g) { return (__ldg(vb + (g >> 5)) >> (g & 31)) & 1u; }

template <class Rec>
__global__ void __launch_bounds__(256)
k_assign(Rec rec, u64 n, const uint4 *__restrict__ meta, const uint4 *__restrict__ lin4, const u32 *__restrict__ top_idx,
         const u32 *__restrict__ vb, u32 G, u32 half_avg, u32 w, u32 *__restrict__ uniq2, u32 *__restrict__ lca_cnt,
         u32 *__restrict__ child_mark, u32 *__restrict__ fb_mark, u32 *__restrict__ cov2,
         unsigned char *__restrict__ res_kind, u32 *__restrict__ res_val, DevScalars *sc)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    u32 n_u2 = 0;
    for (u64 base = (u64)blockIdx.x * blockDim.x; base < n; base += stride) {   // block-uniform trip count
        const u64 i = base + threadIdx.x;
        u32 kind = 0, key = 0;
        if (i < n) {
            const u32 g = rec.refid(i);
            bool leader = g < G && is_valid(vb, g);
            const u32 r = leader ? rec.read(i) : 0;
            if (leader) {                  // an earlier surviving record of this read leads instead
                u64 j = i;
                while (j > 0 && rec.read(j - 1) == r) {
                    --j;
                    u32 h = rec.refid(j);
                    if (h < G && is_valid(vb, h)) { leader = false; break; }
                }
            }
            if (leader) {
                const uint4 la = __ldg(lin4 + 2 * (u64)g), lb = __ldg(lin4 + 2 * (u64)g + 1);
                u32 eq = 0xFFu, gmax = g;
                bool multi = false;
                u64 j = i + 1;
                while (j < n && rec.read(j) == r) {
                    const u32 h = rec.refid(j);
                    if (h < G && h != g && is_valid(vb, h)) {
                        multi = true;
                        gmax = max(gmax, h);
                        const uint4 ha = __ldg(lin4 + 2 * (u64)h), hb = __ldg(lin4 + 2 * (u64)h + 1);
                        u32 m = (ha.x == la.x) | ((ha.y == la.y) << 1) | ((ha.z == la.z) << 2) | ((ha.w == la.w) << 3) |
                                ((hb.x == lb.x) << 4) | ((hb.y == lb.y) << 5) | ((hb.z == lb.z) << 6) | ((hb.w == lb.w) << 7);
                        eq &= m;
                    }
                    ++j;
                }
                const u64 run_end = j;
                if (!multi) {              // sole survivor: uniq_reads_count2 / uniq_cov2 (:383-390)
                    kind = 1; key = g; ++n_u2;
                    if (cov2) {
                        const uint4 m = __ldg(meta + g);
                        u32 center = min(rec.upos(i) + half_avg, m.x);
                        atomicAdd(cov2 + (((u64)m.w << 32) | m.z) + center / w, 1u);
                    }
                    if (res_kind) { res_kind[i] = 1; res_val[i] = g; }
                } else {                   // level-wise LCA over 8-slot lineages, zeros included
                    kind = 2;
                    u32 level, owner;
                    const bool fb = eq == 0;
                    if (!fb) { level = __ffs(eq) - 1; owner = g; } else { level = 7; owner = gmax; }
                    key = owner * 8 + level;
                    const u32 trow = fb ? __ldg(top_idx + gmax) : 0;
                    for (u64 k = i; k < run_end; ++k) {         // children[lca] U= S (:555)
                        const u32 h = rec.refid(k);
                        if (h < G && is_valid(vb, h)) {
                            u32 *mk = fb ? fb_mark + (u64)trow * G + h : child_mark + (u64)h * 8 + level;
                            if (*mk == 0) *mk = 1;
                        }
                    }
                    if (res_kind) {
                        const u32 *lp = reinterpret_cast<const u32 *>(lin4);
                        res_kind[i] = 2; res_val[i] = __ldg(lp + (u64)owner * 8 + level);
                    }
                }
            }
        }
        warp_agg_add(uniq2, key, kind == 1);
        warp_agg_add(lca_cnt, key, kind == 2);
    }
    n_u2 = warp_sum(n_u2);
    __shared__ u32 s_u2;
    if (threadIdx.x == 0) s_u2 = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && n_u2) atomicAdd(&s_u2, n_u2);
    __syncthreads();
    if (threadIdx.x == 0 && s_u2) atomicAdd(&sc->n_uniq2, (unsigned long long)s_u2);
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

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
enum { ST_CREATED = 0, ST_COVERAGE = 1, ST_FILTER = 2, ST_ASSIGN = 3 };

struct slimm_gpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    bool own_stream = true;
    cudaEvent_t upload_done = nullptr;
    u32 G = 0, w = 0, avg = 0, flags = 0, n_top = 0, npow2 = 1;
    int sm_count = 148;
    std::vector<u32> h_len, h_lin, h_top_vals, h_top_idx;
    std::vector<u64> h_off;                 // [G+1] padded bin offsets
    u64 Bp = 0, B = 0;
    // device
    uint4 *d_meta = nullptr; u32 *d_lin = nullptr; u32 *d_top_idx = nullptr; u64 *d_off = nullptr;
    unsigned long long *d_hist = nullptr; u32 *d_cov2 = nullptr;
    u32 *d_stats = nullptr; float *d_cp = nullptr; u32 *d_scratch = nullptr;
    u32 *d_valid_bits = nullptr; unsigned char *d_valid_bytes = nullptr;
    u32 *d_assign = nullptr; u64 assign_words = 0;
    DevScalars *d_sc = nullptr;
    u32 *d_tmp_bins = nullptr; u32 tmp_bins_cap = 0;
    // records
    u32 *d_rid = nullptr, *d_ref = nullptr; i32 *d_pos = nullptr;
    u64 n = 0, cap = 0;
    bool external = false;
    // sorted copies
    u32 *d_rid_sorted = nullptr; uint2 *d_rp_sorted = nullptr; bool use_sorted = false; u64 sorted_cap = 0;
    // per-read results
    unsigned char *d_kind = nullptr; u32 *d_val = nullptr; u64 res_cap = 0;
    int stage = ST_CREATED;
    u64 global_hits = 0; bool have_global_hits = false;
    bool was_sorted = true;
    float q = 0.95f; u32 min_reads_opt = 0;
    bool timing = false;
    cudaEvent_t ev[SLIMM_GPU_T_COUNT][2] = {};
    bool ev_used[SLIMM_GPU_T_COUNT] = {};
    u64 launches = 0;
    std::vector<u32> h_assign;              // host copy of the assign block
    bool h_assign_ok = false;
    std::string err;
};

static const char *k_errstr[] = {"ok", "invalid argument", "no CUDA device", "CUDA error", "out of memory",
                                 "stage called out of order", "too many records (> 2^32-1)"};

#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            char b_[512];                                                                                     \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->err = b_;                                                                                    \
            return e_ == cudaErrorMemoryAllocation ? SLIMM_GPU_ENOMEM : SLIMM_GPU_ECUDA;                      \
        }                                                                                                     \
    } while (0)

static int fail(slimm_gpu_ctx *ctx, int code, const char *msg) { if (ctx) ctx->err = msg; return code; }

struct TimeScope {
    slimm_gpu_ctx *c; int id;
    TimeScope(slimm_gpu_ctx *c_, int id_) : c(c_), id(id_) {
        if (c->timing) { cudaEventRecord(c->ev[id][0], c->stream); c->ev_used[id] = true; }
    }
    ~TimeScope() { if (c->timing) cudaEventRecord(c->ev[id][1], c->stream); }
};

static int grid_for(slimm_gpu_ctx *ctx, u64 n, int block, int blocks_per_sm)
{
    u64 want = (n + block - 1) / block;
    u64 cap = (u64)ctx->sm_count * blocks_per_sm;
    return (int)std::max<u64>(1, std::min(want, cap));
}

extern "C" {

const char *slimm_gpu_strerror(int code) { return (code >= 0 && code <= 6) ? k_errstr[code] : "unknown error"; }
const char *slimm_gpu_last_error(const slimm_gpu_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }

int slimm_gpu_device_count(int *n)
{
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (n) *n = (e == cudaSuccess) ? c : 0;
    return (e == cudaSuccess && c > 0) ? SLIMM_GPU_OK : SLIMM_GPU_ENODEVICE;
}

int slimm_gpu_host_alloc(void **p, uint64_t bytes)
{
    return cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? SLIMM_GPU_OK : SLIMM_GPU_ENOMEM;
}
int slimm_gpu_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? SLIMM_GPU_OK : SLIMM_GPU_ECUDA; }

static int layout_bins(slimm_gpu_ctx *ctx)
{
    const u32 G = ctx->G;
    ctx->h_off.assign((size_t)G + 1, 0);
    std::vector<uint4> meta(G);
    u64 B = 0;
    for (u32 g = 0; g < G; ++g) {
        const u32 nb = ctx->h_len[g] / ctx->w + 1u;                 // src/reference_contig.hpp:80
        const u64 off = ctx->h_off[g];
        meta[g] = make_uint4(ctx->h_len[g], nb, (u32)off, (u32)(off >> 32));
        ctx->h_off[g + 1] = off + (((u64)nb + 63) & ~63ull);
        B += nb;
    }
    ctx->B = B;
    const u64 Bp = ctx->h_off[G];
    if (Bp != ctx->Bp || !ctx->d_hist) {
        if (ctx->d_hist) cudaFree(ctx->d_hist);
        if (ctx->d_cov2) cudaFree(ctx->d_cov2);
        ctx->d_hist = nullptr; ctx->d_cov2 = nullptr;
        CU(cudaMalloc(&ctx->d_hist, std::max<u64>(Bp, 64) * 8));
        if (ctx->flags & SLIMM_GPU_KEEP_UNIQ_COV2) CU(cudaMalloc(&ctx->d_cov2, std::max<u64>(Bp, 64) * 4));
        ctx->Bp = Bp;
    }
    CU(cudaMemcpy(ctx->d_meta, meta.data(), (size_t)G * sizeof(uint4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_off, ctx->h_off.data(), ((size_t)G + 1) * 8, cudaMemcpyHostToDevice));
    return SLIMM_GPU_OK;
}

int slimm_gpu_create(const slimm_gpu_config *cfg, slimm_gpu_ctx **out)
{
    if (!cfg || !out || cfg->n_refs == 0 || !cfg->ref_len || !cfg->lineage || cfg->bin_width == 0) return SLIMM_GPU_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) return SLIMM_GPU_ENODEVICE;
    slimm_gpu_ctx *ctx = new slimm_gpu_ctx();
    *out = ctx;   // returned even on failure so the caller can read last_error, then destroy
    ctx->device = cfg->device;
    CU(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    ctx->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ctx->upload_done, cudaEventDisableTiming));
    for (int i = 0; i < SLIMM_GPU_T_COUNT; ++i) { CU(cudaEventCreate(&ctx->ev[i][0])); CU(cudaEventCreate(&ctx->ev[i][1])); }
    const u32 G = ctx->G = cfg->n_refs;
    ctx->w = cfg->bin_width; ctx->avg = cfg->avg_read_length; ctx->flags = cfg->flags;
    ctx->h_len.assign(cfg->ref_len, cfg->ref_len + G);
    ctx->h_lin.assign(cfg->lineage, cfg->lineage + (size_t)G * 8);
    // dense index of the distinct level-7 taxa (rows of the fallback child marks)
    std::map<u32, u32> top;
    for (u32 g = 0; g < G; ++g) top.emplace(ctx->h_lin[(size_t)g * 8 + 7], 0);
    u32 t = 0;
    for (auto &kv : top) { kv.second = t++; ctx->h_top_vals.push_back(kv.first); }
    ctx->n_top = t;
    if ((u64)ctx->n_top * G > (1ull << 28)) return fail(ctx, SLIMM_GPU_EINVAL, "lineage table has too many distinct top-level taxa");
    ctx->h_top_idx.resize(G);
    for (u32 g = 0; g < G; ++g) ctx->h_top_idx[g] = top[ctx->h_lin[(size_t)g * 8 + 7]];
    ctx->npow2 = 1;
    while (ctx->npow2 < G) ctx->npow2 <<= 1;
    ctx->assign_words = (u64)(17 + ctx->n_top) * G;
    CU(cudaMalloc(&ctx->d_meta, (size_t)G * sizeof(uint4)));
    CU(cudaMalloc(&ctx->d_off, ((size_t)G + 1) * 8));
    CU(cudaMalloc(&ctx->d_lin, (size_t)G * 32));
    CU(cudaMalloc(&ctx->d_top_idx, (size_t)G * 4));
    CU(cudaMalloc(&ctx->d_stats, (size_t)G * 16));
    CU(cudaMalloc(&ctx->d_cp, (size_t)G * 8));
    CU(cudaMalloc(&ctx->d_scratch, (size_t)ctx->npow2 * 8));
    CU(cudaMalloc(&ctx->d_valid_bits, ((size_t)G + 31) / 32 * 4 + 4));
    CU(cudaMalloc(&ctx->d_valid_bytes, G));
    CU(cudaMalloc(&ctx->d_assign, ctx->assign_words * 4));
    CU(cudaMalloc(&ctx->d_sc, sizeof(DevScalars)));
    CU(cudaMemcpy(ctx->d_lin, ctx->h_lin.data(), (size_t)G * 32, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_top_idx, ctx->h_top_idx.data(), (size_t)G * 4, cudaMemcpyHostToDevice));
    int rc = layout_bins(ctx);
    if (rc) return rc;
    if (cfg->reserve_records) {
        ctx->cap = cfg->reserve_records;
        CU(cudaMalloc(&ctx->d_rid, ctx->cap * 4)); CU(cudaMalloc(&ctx->d_ref, ctx->cap * 4)); CU(cudaMalloc(&ctx->d_pos, ctx->cap * 4));
    }
    return SLIMM_GPU_OK;
}

static void free_records(slimm_gpu_ctx *ctx)
{
    if (!ctx->external) { cudaFree(ctx->d_rid); cudaFree(ctx->d_ref); cudaFree(ctx->d_pos); }
    ctx->d_rid = ctx->d_ref = nullptr; ctx->d_pos = nullptr; ctx->cap = 0; ctx->n = 0; ctx->external = false;
}

int slimm_gpu_destroy(slimm_gpu_ctx *ctx)
{
    if (!ctx) return SLIMM_GPU_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    free_records(ctx);
    cudaFree(ctx->d_meta); cudaFree(ctx->d_off); cudaFree(ctx->d_lin); cudaFree(ctx->d_top_idx); cudaFree(ctx->d_hist);
    cudaFree(ctx->d_cov2); cudaFree(ctx->d_stats); cudaFree(ctx->d_cp); cudaFree(ctx->d_scratch); cudaFree(ctx->d_valid_bits);
    cudaFree(ctx->d_valid_bytes); cudaFree(ctx->d_assign); cudaFree(ctx->d_sc); cudaFree(ctx->d_tmp_bins);
    cudaFree(ctx->d_rid_sorted); cudaFree(ctx->d_rp_sorted); cudaFree(ctx->d_kind); cudaFree(ctx->d_val);
    for (int i = 0; i < SLIMM_GPU_T_COUNT; ++i) { if (ctx->ev[i][0]) cudaEventDestroy(ctx->ev[i][0]); if (ctx->ev[i][1]) cudaEventDestroy(ctx->ev[i][1]); }
    if (ctx->upload_done) cudaEventDestroy(ctx->upload_done);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
    return SLIMM_GPU_OK;
}

int slimm_gpu_reset(slimm_gpu_ctx *ctx, uint32_t bin_width, uint32_t avg_read_length)
{
    if (!ctx) return SLIMM_GPU_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->external) { ctx->d_rid = ctx->d_ref = nullptr; ctx->d_pos = nullptr; ctx->cap = 0; ctx->external = false; }
    ctx->n = 0; ctx->stage = ST_CREATED; ctx->use_sorted = false; ctx->have_global_hits = false; ctx->h_assign_ok = false;
    if (avg_read_length) ctx->avg = avg_read_length;
    if (bin_width && bin_width != ctx->w) { ctx->w = bin_width; return layout_bins(ctx); }
    return SLIMM_GPU_OK;
}

int slimm_gpu_set_stream(slimm_gpu_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return SLIMM_GPU_EINVAL;
    CU(cudaSetDevice(ctx->device));
    if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; }
    else { CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    return SLIMM_GPU_OK;
}

static int reserve(slimm_gpu_ctx *ctx, u64 need)
{
    if (need <= ctx->cap) return SLIMM_GPU_OK;
    u64 ncap = std::max<u64>(need, ctx->cap + ctx->cap / 2);
    u32 *rid = nullptr, *ref = nullptr; i32 *pos = nullptr;
    CU(cudaMalloc(&rid, ncap * 4)); CU(cudaMalloc(&ref, ncap * 4)); CU(cudaMalloc(&pos, ncap * 4));
    if (ctx->n) {
        CU(cudaStreamSynchronize(ctx->copy_stream));
        CU(cudaMemcpy(rid, ctx->d_rid, ctx->n * 4, cudaMemcpyDeviceToDevice));
        CU(cudaMemcpy(ref, ctx->d_ref, ctx->n * 4, cudaMemcpyDeviceToDevice));
        CU(cudaMemcpy(pos, ctx->d_pos, ctx->n * 4, cudaMemcpyDeviceToDevice));
    }
    cudaFree(ctx->d_rid); cudaFree(ctx->d_ref); cudaFree(ctx->d_pos);
    ctx->d_rid = rid; ctx->d_ref = ref; ctx->d_pos = pos; ctx->cap = ncap;
    return SLIMM_GPU_OK;
}

int slimm_gpu_push(slimm_gpu_ctx *ctx, const uint32_t *read_id, const uint32_t *ref_id, const int32_t *begin_pos, uint64_t n)
{
    if (!ctx || (n && (!read_id || !ref_id || !begin_pos))) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_CREATED) return fail(ctx, SLIMM_GPU_ESTATE, "push after coverage; call slimm_gpu_reset first");
    if (ctx->external) return fail(ctx, SLIMM_GPU_ESTATE, "push after push_device");
    if (ctx->n + n > 0xFFFFFFFFull) return fail(ctx, SLIMM_GPU_ERANGE, "more than 2^32-1 records");
    if (n == 0) return SLIMM_GPU_OK;
    CU(cudaSetDevice(ctx->device));
    int rc = reserve(ctx, ctx->n + n);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->d_rid + ctx->n, read_id, n * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(cudaMemcpyAsync(ctx->d_ref + ctx->n, ref_id, n * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(cudaMemcpyAsync(ctx->d_pos + ctx->n, begin_pos, n * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    ctx->n += n;
    return SLIMM_GPU_OK;
}

int slimm_gpu_push_device(slimm_gpu_ctx *ctx, const uint32_t *d_read_id, const uint32_t *d_ref_id, const int32_t *d_begin_pos, uint64_t n)
{
    if (!ctx || !d_read_id || !d_ref_id || !d_begin_pos) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_CREATED || ctx->n != 0) return fail(ctx, SLIMM_GPU_ESTATE, "push_device needs an empty context");
    if (n > 0xFFFFFFFFull) return fail(ctx, SLIMM_GPU_ERANGE, "more than 2^32-1 records");
    free_records(ctx);
    ctx->d_rid = const_cast<u32 *>(d_read_id); ctx->d_ref = const_cast<u32 *>(d_ref_id); ctx->d_pos = const_cast<i32 *>(d_begin_pos);
    ctx->n = n; ctx->cap = n; ctx->external = true;
    return SLIMM_GPU_OK;
}

int slimm_gpu_sync_uploads(slimm_gpu_ctx *ctx)
{
    if (!ctx) return SLIMM_GPU_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->copy_stream));
    return SLIMM_GPU_OK;
}

static int launch_coverage(slimm_gpu_ctx *ctx)
{
    const int grid = grid_for(ctx, ctx->n, 256, 8);
    const u32 half = ctx->avg / 2u;
    if (ctx->use_sorted) {
        RecPacked rec{ctx->d_rid_sorted, ctx->d_rp_sorted};
        k_coverage<<<grid, 256, 0, ctx->stream>>>(rec, ctx->n, ctx->d_meta, ctx->G, half, ctx->w, ctx->d_hist, ctx->d_sc);
    } else {
        RecSoA rec{ctx->d_rid, ctx->d_ref, ctx->d_pos};
        k_coverage<<<grid, 256, 0, ctx->stream>>>(rec, ctx->n, ctx->d_meta, ctx->G, half, ctx->w, ctx->d_hist, ctx->d_sc);
    }
    ctx->launches++;
    CU(cudaGetLastError());
    return SLIMM_GPU_OK;
}

static int zero_state(slimm_gpu_ctx *ctx)
{
    TimeScope ts(ctx, SLIMM_GPU_T_ZERO);
    CU(cudaMemsetAsync(ctx->d_hist, 0, std::max<u64>(ctx->Bp, 64) * 8, ctx->stream));
    if (ctx->d_cov2) CU(cudaMemsetAsync(ctx->d_cov2, 0, std::max<u64>(ctx->Bp, 64) * 4, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_sc, 0, sizeof(DevScalars), ctx->stream));
    return SLIMM_GPU_OK;
}

// stable device sort by read id for input that is not grouped by read (coordinate-sorted BAMs,
// shuffled files): CUB radix sort (library code, the fallback path only) on {read_id, (ref,pos)}
static int sort_records(slimm_gpu_ctx *ctx)
{
    TimeScope ts(ctx, SLIMM_GPU_T_SORT);
    const u64 n = ctx->n;
    if (ctx->sorted_cap < n) {
        cudaFree(ctx->d_rid_sorted); cudaFree(ctx->d_rp_sorted);
        ctx->d_rid_sorted = nullptr; ctx->d_rp_sorted = nullptr;
        CU(cudaMalloc(&ctx->d_rid_sorted, n * 4)); CU(cudaMalloc(&ctx->d_rp_sorted, n * 8));
        ctx->sorted_cap = n;
    }
    uint2 *packed = nullptr;
    CU(cudaMalloc(&packed, n * 8));
    k_pack_values<<<grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(ctx->d_ref, ctx->d_pos, n, packed);
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx->d_rid, ctx->d_rid_sorted, (const u64 *)packed,
                                    (u64 *)ctx->d_rp_sorted, n, 0, 32, ctx->stream);
    void *tmp = nullptr;
    CU(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, ctx->d_rid, ctx->d_rid_sorted, (const u64 *)packed,
                                                    (u64 *)ctx->d_rp_sorted, n, 0, 32, ctx->stream);
    ctx->launches += 1;   // k_pack_values; the CUB sort kernels are library code and not counted
    cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp); cudaFree(packed);
    if (e != cudaSuccess) { ctx->err = std::string("cub radix sort failed: ") + cudaGetErrorString(e); return SLIMM_GPU_ECUDA; }
    ctx->use_sorted = true;
    return SLIMM_GPU_OK;
}

int slimm_gpu_coverage(slimm_gpu_ctx *ctx)
{
    if (!ctx) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_CREATED) return fail(ctx, SLIMM_GPU_ESTATE, "coverage already ran; call slimm_gpu_reset for a new sample");
    CU(cudaSetDevice(ctx->device));
    // kernels wait for the uploads without blocking the host
    CU(cudaEventRecord(ctx->upload_done, ctx->copy_stream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->upload_done, 0));
    for (int i = 0; i < SLIMM_GPU_T_COUNT; ++i) ctx->ev_used[i] = false;
    ctx->use_sorted = false; ctx->was_sorted = true; ctx->h_assign_ok = false;
    int rc = zero_state(ctx);
    if (rc) return rc;
    if (ctx->n) {
        { TimeScope ts(ctx, SLIMM_GPU_T_COVERAGE); rc = launch_coverage(ctx); }
        if (rc) return rc;
        // optimistic: the kernel assumed non-decreasing read ids and verified it on the fly
        u32 flags = 0;
        CU(cudaMemcpyAsync(&flags, &ctx->d_sc->flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (flags & 2u) return fail(ctx, SLIMM_GPU_EINVAL, "a record references a contig id >= n_refs");
        if (flags & 1u) {
            ctx->was_sorted = false;
            rc = sort_records(ctx); if (rc) return rc;
            rc = zero_state(ctx); if (rc) return rc;
            { TimeScope ts(ctx, SLIMM_GPU_T_COVERAGE); rc = launch_coverage(ctx); }
            if (rc) return rc;
        }
    }
    ctx->stage = ST_COVERAGE;
    return SLIMM_GPU_OK;
}

int slimm_gpu_bins_device(slimm_gpu_ctx *ctx, void **d_ptr, uint64_t *n_u32)
{
    if (!ctx || !d_ptr || !n_u32) return SLIMM_GPU_EINVAL;
    *d_ptr = ctx->d_hist; *n_u32 = ctx->Bp * 2;
    return SLIMM_GPU_OK;
}

int slimm_gpu_counters_device(slimm_gpu_ctx *ctx, void **d_ptr, uint64_t *n_u64)
{
    if (!ctx || !d_ptr || !n_u64) return SLIMM_GPU_EINVAL;
    *d_ptr = &ctx->d_sc->n_reads; *n_u64 = 2;
    return SLIMM_GPU_OK;
}

int slimm_gpu_set_global_hits(slimm_gpu_ctx *ctx, uint64_t hits)
{
    if (!ctx) return SLIMM_GPU_EINVAL;
    if (hits > 0xFFFFFFFFull) return fail(ctx, SLIMM_GPU_ERANGE, "more than 2^32-1 records");
    ctx->global_hits = hits; ctx->have_global_hits = true;
    return SLIMM_GPU_OK;
}

int slimm_gpu_filter(slimm_gpu_ctx *ctx, float cov_cut_off, uint32_t min_reads)
{
    if (!ctx) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_COVERAGE) return fail(ctx, SLIMM_GPU_ESTATE, "filter needs coverage first");
    CU(cudaSetDevice(ctx->device));
    ctx->q = cov_cut_off; ctx->min_reads_opt = min_reads;
    {
        TimeScope ts(ctx, SLIMM_GPU_T_STATS);
        CU(cudaMemsetAsync(ctx->d_stats, 0, (size_t)ctx->G * 16, ctx->stream));
        const u64 n_steps = ctx->Bp / 64;
        const u64 n_chunks = (n_steps + STATS_STEPS_PER_WARP - 1) / STATS_STEPS_PER_WARP;
        const int grid = grid_for(ctx, n_chunks * 32, 256, 8);
        if (n_steps) k_ref_stats<<<grid, 256, 0, ctx->stream>>>((const uint4 *)ctx->d_hist, n_steps, ctx->d_off, ctx->G, ctx->d_stats);
        ctx->launches += 1;
        CU(cudaGetLastError());
    }
    {
        TimeScope ts(ctx, SLIMM_GPU_T_CUTOFF);
        k_cutoffs<<<2, 1024, 0, ctx->stream>>>(ctx->d_stats, ctx->d_meta, ctx->G, cov_cut_off, min_reads, ctx->d_cp, ctx->d_scratch,
                                               ctx->npow2, ctx->d_valid_bits, ctx->d_valid_bytes, ctx->d_sc);
        ctx->launches++;
        CU(cudaGetLastError());
    }
    ctx->stage = ST_FILTER;
    return SLIMM_GPU_OK;
}

int slimm_gpu_assign(slimm_gpu_ctx *ctx)
{
    if (!ctx) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_FILTER) return fail(ctx, SLIMM_GPU_ESTATE, "assign needs filter first");
    CU(cudaSetDevice(ctx->device));
    const u32 G = ctx->G;
    if ((ctx->flags & SLIMM_GPU_READ_RESULTS) && ctx->res_cap < ctx->n) {
        cudaFree(ctx->d_kind); cudaFree(ctx->d_val);
        ctx->d_kind = nullptr; ctx->d_val = nullptr;
        CU(cudaMalloc(&ctx->d_kind, std::max<u64>(ctx->n, 1))); CU(cudaMalloc(&ctx->d_val, std::max<u64>(ctx->n, 1) * 4));
        ctx->res_cap = ctx->n;
    }
    TimeScope ts(ctx, SLIMM_GPU_T_ASSIGN);
    CU(cudaMemsetAsync(ctx->d_assign, 0, ctx->assign_words * 4, ctx->stream));
    if (ctx->d_kind) CU(cudaMemsetAsync(ctx->d_kind, 0, std::max<u64>(ctx->n, 1), ctx->stream));
    if (ctx->n) {
        u32 *uniq2 = ctx->d_assign, *lca = uniq2 + G, *cm = lca + (u64)8 * G, *fb = cm + (u64)8 * G;
        const int grid = grid_for(ctx, ctx->n, 256, 8);
        const u32 half = ctx->avg / 2u;
        if (ctx->use_sorted) {
            RecPacked rec{ctx->d_rid_sorted, ctx->d_rp_sorted};
            k_assign<<<grid, 256, 0, ctx->stream>>>(rec, ctx->n, ctx->d_meta, (const uint4 *)ctx->d_lin, ctx->d_top_idx, ctx->d_valid_bits,
                                                    G, half, ctx->w, uniq2, lca, cm, fb, ctx->d_cov2, ctx->d_kind, ctx->d_val, ctx->d_sc);
        } else {
            RecSoA rec{ctx->d_rid, ctx->d_ref, ctx->d_pos};
            k_assign<<<grid, 256, 0, ctx->stream>>>(rec, ctx->n, ctx->d_meta, (const uint4 *)ctx->d_lin, ctx->d_top_idx, ctx->d_valid_bits,
                                                    G, half, ctx->w, uniq2, lca, cm, fb, ctx->d_cov2, ctx->d_kind, ctx->d_val, ctx->d_sc);
        }
        ctx->launches++;
        CU(cudaGetLastError());
    }
    ctx->stage = ST_ASSIGN;
    ctx->h_assign_ok = false;
    return SLIMM_GPU_OK;
}

int slimm_gpu_assign_device(slimm_gpu_ctx *ctx, void **d_ptr, uint64_t *n_u32)
{
    if (!ctx || !d_ptr || !n_u32) return SLIMM_GPU_EINVAL;
    *d_ptr = ctx->d_assign; *n_u32 = ctx->assign_words;
    return SLIMM_GPU_OK;
}

int slimm_gpu_run(slimm_gpu_ctx *ctx, float cov_cut_off, uint32_t min_reads)
{
    int rc = slimm_gpu_coverage(ctx);
    if (rc) return rc;
    rc = slimm_gpu_filter(ctx, cov_cut_off, min_reads);
    if (rc) return rc;
    return slimm_gpu_assign(ctx);
}

// ---- results -----------------------------------------------------------------------------------
int slimm_gpu_get_summary(slimm_gpu_ctx *ctx, slimm_gpu_summary *out)
{
    if (!ctx || !out) return SLIMM_GPU_EINVAL;
    if (ctx->stage < ST_COVERAGE) return fail(ctx, SLIMM_GPU_ESTATE, "nothing has run yet");
    CU(cudaSetDevice(ctx->device));
    DevScalars s;
    CU(cudaMemcpyAsync(&s, ctx->d_sc, sizeof s, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    memset(out, 0, sizeof *out);
    out->hits_count = (u32)(ctx->have_global_hits ? ctx->global_hits : ctx->n);
    out->matches_count = (u32)s.n_reads; out->uniq_matches_count = (u32)s.n_uniq; out->uniq_matches_count2 = (u32)s.n_uniq2;
    out->reference_count = s.ref_count; out->n_valid = s.n_valid; out->failed_by_cov = s.failed_cov;
    out->failed_by_uniq_cov = s.failed_ucov; out->failed_by_min_read = s.failed_minread;
    out->min_reads = ctx->min_reads_opt ? ctx->min_reads_opt : ((u32)s.n_reads ? 1u + ((u32)s.n_reads - 1u) / 10000u : 0u);
    out->coverage_cut_off = s.cut; out->uniq_coverage_cut_off = s.ucut; out->n_pairs = s.n_pairs; out->n_bins = ctx->B;
    out->input_was_sorted = ctx->was_sorted;
    return SLIMM_GPU_OK;
}

static int fetch_assign(slimm_gpu_ctx *ctx)
{
    if (ctx->h_assign_ok) return SLIMM_GPU_OK;
    if (ctx->stage < ST_ASSIGN) return fail(ctx, SLIMM_GPU_ESTATE, "assign has not run");
    CU(cudaSetDevice(ctx->device));
    ctx->h_assign.resize(ctx->assign_words);
    CU(cudaMemcpyAsync(ctx->h_assign.data(), ctx->d_assign, ctx->assign_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->h_assign_ok = true;
    return SLIMM_GPU_OK;
}

int slimm_gpu_get_ref_stats(slimm_gpu_ctx *ctx, uint32_t *reads_count, uint32_t *uniq_reads_count, uint32_t *uniq_reads_count2,
                            uint32_t *nz_bins, uint32_t *uniq_nz_bins, float *cov_percent, float *uniq_cov_percent, uint8_t *valid)
{
    if (!ctx) return SLIMM_GPU_EINVAL;
    if (ctx->stage < ST_FILTER) return fail(ctx, SLIMM_GPU_ESTATE, "filter has not run");
    CU(cudaSetDevice(ctx->device));
    const u32 G = ctx->G;
    std::vector<u32> st((size_t)G * 4);
    std::vector<float> cp((size_t)G * 2);
    CU(cudaMemcpyAsync(st.data(), ctx->d_stats, (size_t)G * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(cp.data(), ctx->d_cp, (size_t)G * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (valid) CU(cudaMemcpyAsync(valid, ctx->d_valid_bytes, G, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (u32 g = 0; g < G; ++g) {
        if (nz_bins) nz_bins[g] = st[4 * (size_t)g];
        if (reads_count) reads_count[g] = st[4 * (size_t)g + 1];
        if (uniq_nz_bins) uniq_nz_bins[g] = st[4 * (size_t)g + 2];
        if (uniq_reads_count) uniq_reads_count[g] = st[4 * (size_t)g + 3];
        if (cov_percent) cov_percent[g] = cp[g];
        if (uniq_cov_percent) uniq_cov_percent[g] = cp[(size_t)G + g];
    }
    if (uniq_reads_count2) {
        int rc = fetch_assign(ctx);
        if (rc) return rc;
        memcpy(uniq_reads_count2, ctx->h_assign.data(), (size_t)G * 4);
    }
    return SLIMM_GPU_OK;
}

int slimm_gpu_get_lca_counts(slimm_gpu_ctx *ctx, uint32_t *taxon, uint32_t *count, uint64_t cap, uint64_t *n)
{
    if (!ctx || !n) return SLIMM_GPU_EINVAL;
    int rc = fetch_assign(ctx);
    if (rc) return rc;
    const u32 G = ctx->G;
    const u32 *lca = ctx->h_assign.data() + G;
    std::map<u32, u32> acc;                      // taxon -> reads whose LCA it is
    for (u64 s = 0; s < (u64)G * 8; ++s)
        if (lca[s]) acc[ctx->h_lin[s]] += lca[s];
    *n = acc.size();
    u64 i = 0;
    for (auto &kv : acc) {
        if (i >= cap) break;
        if (taxon) taxon[i] = kv.first;
        if (count) count[i] = kv.second;
        ++i;
    }
    return SLIMM_GPU_OK;
}

int slimm_gpu_get_lca_children(slimm_gpu_ctx *ctx, uint32_t *taxon, uint32_t *ref, uint64_t cap, uint64_t *n)
{
    if (!ctx || !n) return SLIMM_GPU_EINVAL;
    int rc = fetch_assign(ctx);
    if (rc) return rc;
    const u32 G = ctx->G;
    const u32 *cm = ctx->h_assign.data() + (u64)9 * G, *fb = ctx->h_assign.data() + (u64)17 * G;
    std::vector<u64> pairs;
    for (u64 s = 0; s < (u64)G * 8; ++s)
        if (cm[s]) pairs.push_back(((u64)ctx->h_lin[s] << 32) | (s >> 3));
    for (u32 t = 0; t < ctx->n_top; ++t)
        for (u32 g = 0; g < G; ++g)
            if (fb[(u64)t * G + g]) pairs.push_back(((u64)ctx->h_top_vals[t] << 32) | g);
    std::sort(pairs.begin(), pairs.end());
    pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
    *n = pairs.size();
    for (u64 i = 0; i < pairs.size() && i < cap; ++i) {
        if (taxon) taxon[i] = (u32)(pairs[i] >> 32);
        if (ref) ref[i] = (u32)pairs[i];
    }
    return SLIMM_GPU_OK;
}

int slimm_gpu_fetch_bins(slimm_gpu_ctx *ctx, int which, uint32_t ref, uint32_t *out, uint32_t cap)
{
    if (!ctx || !out || ref >= ctx->G || which < 0 || which > 2) return SLIMM_GPU_EINVAL;
    if (ctx->stage < ST_COVERAGE) return fail(ctx, SLIMM_GPU_ESTATE, "coverage has not run");
    if (which == 2 && !ctx->d_cov2) return fail(ctx, SLIMM_GPU_EINVAL, "uniq_cov2 needs SLIMM_GPU_KEEP_UNIQ_COV2");
    CU(cudaSetDevice(ctx->device));
    const u32 nb = ctx->h_len[ref] / ctx->w + 1u;
    if (cap < nb) return fail(ctx, SLIMM_GPU_EINVAL, "output buffer smaller than the number of bins");
    if (ctx->tmp_bins_cap < nb) {
        cudaFree(ctx->d_tmp_bins); ctx->d_tmp_bins = nullptr;
        CU(cudaMalloc(&ctx->d_tmp_bins, (size_t)nb * 4));
        ctx->tmp_bins_cap = nb;
    }
    const u32 *src = which == 2 ? ctx->d_cov2 : (const u32 *)ctx->d_hist;
    k_extract_bins<<<(nb + 255) / 256, 256, 0, ctx->stream>>>(src, ctx->h_off[ref], which == 2 ? 1 : 2, which == 1 ? 1 : 0, nb, ctx->d_tmp_bins);
    ctx->launches++;
    CU(cudaMemcpyAsync(out, ctx->d_tmp_bins, (size_t)nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SLIMM_GPU_OK;
}

int slimm_gpu_get_uniq2_nz(slimm_gpu_ctx *ctx, uint32_t *out)
{
    if (!ctx || !out) return SLIMM_GPU_EINVAL;
    if (!ctx->d_cov2) return fail(ctx, SLIMM_GPU_EINVAL, "uniq_cov2 needs SLIMM_GPU_KEEP_UNIQ_COV2");
    if (ctx->stage < ST_ASSIGN) return fail(ctx, SLIMM_GPU_ESTATE, "assign has not run");
    CU(cudaSetDevice(ctx->device));
    u32 *d_out = nullptr;
    CU(cudaMalloc(&d_out, (size_t)ctx->G * 4));
    k_cov2_nz<<<(ctx->G * 32 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_cov2, ctx->d_off, ctx->d_meta, ctx->G, d_out);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(out, d_out, (size_t)ctx->G * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(ctx, SLIMM_GPU_ECUDA, cudaGetErrorString(e));
    return SLIMM_GPU_OK;
}

int slimm_gpu_read_results(slimm_gpu_ctx *ctx, uint32_t *read_id, uint8_t *kind, uint32_t *value, uint64_t cap, uint64_t *n)
{
    if (!ctx || !n) return SLIMM_GPU_EINVAL;
    if (!(ctx->flags & SLIMM_GPU_READ_RESULTS)) return fail(ctx, SLIMM_GPU_EINVAL, "needs SLIMM_GPU_READ_RESULTS");
    if (ctx->stage < ST_ASSIGN) return fail(ctx, SLIMM_GPU_ESTATE, "assign has not run");
    CU(cudaSetDevice(ctx->device));
    const u64 N = ctx->n;
    std::vector<unsigned char> k(N);
    std::vector<u32> v(N), r(N);
    if (N) {
        CU(cudaMemcpyAsync(k.data(), ctx->d_kind, N, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(v.data(), ctx->d_val, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(r.data(), ctx->use_sorted ? ctx->d_rid_sorted : ctx->d_rid, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    u64 m = 0;
    for (u64 i = 0; i < N; ++i)
        if (k[i]) {
            if (m < cap) { if (read_id) read_id[m] = r[i]; if (kind) kind[m] = k[i]; if (value) value[m] = v[i]; }
            ++m;
        }
    *n = m;
    return SLIMM_GPU_OK;
}

// ---- instrumentation ---------------------------------------------------------------------------
int slimm_gpu_enable_timing(slimm_gpu_ctx *ctx, int on) { if (!ctx) return SLIMM_GPU_EINVAL; ctx->timing = on != 0; return SLIMM_GPU_OK; }

int slimm_gpu_get_timings(slimm_gpu_ctx *ctx, float *ms, int n)
{
    if (!ctx || !ms) return SLIMM_GPU_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n && i < SLIMM_GPU_T_COUNT; ++i) {
        ms[i] = 0.0f;
        if (ctx->timing && ctx->ev_used[i]) cudaEventElapsedTime(&ms[i], ctx->ev[i][0], ctx->ev[i][1]);
    }
    return SLIMM_GPU_OK;
}

int slimm_gpu_get_launch_count(slimm_gpu_ctx *ctx, uint64_t *n) { if (!ctx || !n) return SLIMM_GPU_EINVAL; *n = ctx->launches; return SLIMM_GPU_OK; }

}  // extern "C"

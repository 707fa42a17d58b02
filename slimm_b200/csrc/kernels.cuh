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
//   items     : u32[N]   per record, in record order: padded bin index | uniq << 31 (0xFFFFFFFF = repeat hit);
//               grouped : u32[<=N] the same words grouped by histogram slice (k_split)
//   cov2      : u32[Bp]  uniq_cov2 (only with SLIMM_GPU_KEEP_UNIQ_COV2)
//   stats     : u32[G*4] = {nz, reads_count, uniq nz, uniq_reads_count}
//   assign    : u32[(17+T)*G] = uniq_reads_count2[G] | lca_count[G*8] | child_mark[G*8] | fb_mark[T*G]
//               keyed by (reference, lineage level) instead of taxon id, so no device hash map
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cuda/barrier>
#include <stdint.h>

typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;

#define FULL 0xffffffffu
#define ITEM_SKIP 0xFFFFFFFFu

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
    float tot[2];                 // K4: in-order f32 sum of cov_percent / uniq_cov_percent over the references with unique reads
    u32 n_members[2];             // K4: how many references that is (the same number twice)
    u32 tot_ready[2];             // K4: tot[] is final (k_cut_fold runs beside k_cut_sort_cluster, which peeks)
    u32 k_done[2];                // K4: running sums P_0 .. P_{k_done-1} exist (the chain ends early once the stop is certain)
};

// record accessors: plain SoA, or {read_id[], packed (ref | pos<<32)[]} after the device sort
struct RecSoA {
    const u32 *rid; const u32 *ref; const i32 *pos;
    __device__ __forceinline__ u32 read(u32 i) const { return __ldg(rid + i); }
    __device__ __forceinline__ u32 refid(u32 i) const { return __ldg(ref + i); }
    __device__ __forceinline__ u32 upos(u32 i) const { return (u32)__ldg(pos + i); }
    // records i .. i+3 with 128-bit loads (i a multiple of 4, arrays 16-byte aligned)
    __device__ __forceinline__ void load_rg4(u32 i, uint4 &r, uint4 &g) const
    {
        r = __ldg(reinterpret_cast<const uint4 *>(rid + i));
        g = __ldg(reinterpret_cast<const uint4 *>(ref + i));
    }
    __device__ __forceinline__ uint4 load_p4(u32 i) const { return __ldg(reinterpret_cast<const uint4 *>(pos + i)); }
    // lanes 0..2 pull the 128-byte lines that hold record i of the three arrays towards the SM
    __device__ __forceinline__ void prefetch(u32 i, u32 lane) const
    {
        const void *a = lane == 0 ? (const void *)(rid + i) : lane == 1 ? (const void *)(ref + i) : (const void *)(pos + i);
        if (lane < 3) asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
    }
};
struct RecPacked {
    const u32 *rid; const uint2 *rp;
    __device__ __forceinline__ u32 read(u32 i) const { return __ldg(rid + i); }
    __device__ __forceinline__ u32 refid(u32 i) const { return __ldg(&rp[i].x); }
    __device__ __forceinline__ u32 upos(u32 i) const { return __ldg(&rp[i].y); }
    __device__ __forceinline__ void load_rg4(u32 i, uint4 &r, uint4 &g) const
    {
        r = __ldg(reinterpret_cast<const uint4 *>(rid + i));
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(rp + i)), b = __ldg(reinterpret_cast<const uint4 *>(rp + i + 2));
        g = make_uint4(a.x, a.z, b.x, b.z);
    }
    __device__ __forceinline__ uint4 load_p4(u32 i) const
    {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(rp + i)), b = __ldg(reinterpret_cast<const uint4 *>(rp + i + 2));
        return make_uint4(a.y, a.w, b.y, b.w);
    }
    __device__ __forceinline__ void prefetch(u32 i, u32 lane) const
    {
        const void *a = lane == 0 ? (const void *)(rid + i) : lane == 1 ? (const void *)(rp + i) : (const void *)(rp + i + 16);
        if (lane < 3) asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
    }
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

// padded global bin index of a record (reference src/slimm.hpp:200-201): u32 wrap of beginPos + avg/2,
// clamp to the contig length, integer division by the bin width
// exact n / d for a divisor fixed per sample: t = umulhi(mul, n); q = (t + ((n - t) >> s1)) >> s2
// (round-up method; the host derives {mul, s1, s2} from the bin width, slimm_gpu.cu bin_div_for)
struct BinDiv { u32 mul, s1, s2; };
__device__ __forceinline__ u32 fast_div(u32 n, const BinDiv &d)
{
    const u32 t = __umulhi(d.mul, n);
    return (t + ((n - t) >> d.s1)) >> d.s2;
}
__device__ __forceinline__ u64 bin_of_meta(const uint4 &m /* {len, nb, off_lo, off_hi} */, u32 upos, u32 half_avg, const BinDiv &wdiv)
{
    const u32 center = min(upos + half_avg, m.x);
    return (((u64)m.w << 32) | m.z) + fast_div(center, wdiv);
}
__device__ __forceinline__ u64 bin_of(const uint4 *__restrict__ meta, u32 g, u32 upos, u32 half_avg, const BinDiv &wdiv)
{
    return bin_of_meta(__ldg(meta + g), upos, half_avg, wdiv);
}

// ------------------------------------------------------------------------------------------------
// Sliding 32-record window over the read-grouped record stream.  A read is a run of equal read_id.
// A window always STARTS at the head of a run; every run that also ends inside the window is
// "whole" and is resolved with ballots and bit masks alone - no per-thread rescans, no carried
// state.  The window that follows starts at the head of the first run that did not end here, so a
// record is analysed exactly once (as a lane of a whole run).  A run longer than 32 records never
// fits a window: the warp walks it cooperatively (the *_long_run functions).
// Record indices are 32-bit: a context holds at most SLIMM_MAX_RECORDS = 2^32 - 256 records.
// ------------------------------------------------------------------------------------------------
struct Window {
    u32 r, g;        // read id / reference id of this lane's record
    u32 M;           // lanes of my run when it is whole, else 0
    int s, e;        // first / last lane of my run (valid when whole)
    bool whole;      // my run starts and ends inside the window and is owned by this chunk
    bool long_run;   // warp-uniform: the window holds one run only and it does not end here
    u32 next;        // warp-uniform: head of the first run not resolved here
};

#define LANE_LT(lane) ((1u << (lane)) - 1u)
#define LANE_LE(lane) (FULL >> (31 - (lane)))
#define LANE_GE(lane) (FULL << (lane))

// masks from the ballots of run heads (H, bit 0 set) and run ends (E) of a window of n_in records
__device__ __forceinline__ void window_masks(u32 H, u32 E, u32 n_in, u32 p, u32 chunk_end, u32 lane, bool in, Window &w)
{
    w.s = 31 - __clz(H & LANE_LE(lane));
    const u32 Eg = E & LANE_GE(lane);
    w.e = __ffs(Eg) - 1;
    w.whole = in && Eg != 0 && p + (u32)w.s < chunk_end;   // runs headed at or after chunk_end belong to the next chunk
    w.M = w.whole ? (LANE_LE(w.e) & LANE_GE(w.s)) : 0u;
    w.long_run = false;
    if ((E >> (n_in - 1)) & 1u) w.next = p + n_in;
    else {
        const u32 s_last = 31 - __clz(H);
        w.next = p + s_last;
        w.long_run = s_last == 0;
    }
}

// One window's worth of record fields, fetched ahead of their use: k_coverage issues the loads of the window that
// follows (its start is known as soon as the ballots of the current one are in) before it works through the current
// one, so the DRAM latency of the three arrays is hidden behind ~150 instructions of independent work.
struct WinRegs { u32 r, g, upos, nx; };   // nx: read id of the record behind the window (lane 31 only)

template <class Rec>
__device__ __forceinline__ WinRegs fetch_window(const Rec &rec, u32 p, u32 n, u32 lane)
{
    const u32 i = p + lane;
    const bool in = i < n;
    WinRegs x;
    x.r = in ? rec.read(i) : 0u;
    x.g = in ? rec.refid(i) : 0u;
    x.upos = in ? rec.upos(i) : 0u;
    x.nx = (lane == 31 && i + 1 < n) ? rec.read(i + 1) : 0u;
    return x;
}

__device__ __forceinline__ void analyse_window(const WinRegs &x, u32 p, u32 n, u32 chunk_end, u32 lane, Window &w, u32 &bad)
{
    const u32 i = p + lane;
    const bool in = i < n;
    const bool has_nx = i + 1 < n;
    w.r = x.r;
    w.g = x.g;
    u32 nx = __shfl_down_sync(FULL, w.r, 1);
    if (lane == 31) nx = x.nx;
    const u32 pv = __shfl_up_sync(FULL, w.r, 1);
    const bool head = in && (lane == 0 || pv != w.r);
    const bool last = in && (!has_nx || nx != w.r);
    if (has_nx && nx < w.r) bad |= 1u;                    // read ids must be non-decreasing
    const u32 H = __ballot_sync(FULL, head), E = __ballot_sync(FULL, last);
    window_masks(H, E, min(32u, n - p), p, chunk_end, lane, in, w);
}

// first run head at or after c0 (n when there is none); warp-uniform
template <class Rec>
__device__ __forceinline__ u32 find_head(const Rec &rec, u32 c0, u32 n, u32 lane)
{
    if (c0 == 0) return 0;
    for (u32 q = c0; q < n; q += 32) {
        const u32 i = q + lane;
        const bool h = i < n && rec.read(i) != rec.read(i - 1);
        const u32 b = __ballot_sync(FULL, h);
        if (b) return q + (u32)(__ffs(b) - 1);
    }
    return n;
}

#define CHUNK 2048u          // records per warp work unit
#define PREFETCH_AHEAD 160u  // records between a window and the lines prefetched for the windows after it
#define CW_SLOT (CHUNK + 32) // compact words a chunk can emit (its last run may reach 31 records past the chunk)
#define LR_SLOT 64u          // long runs a chunk can own (each is longer than 32 records)
#define RS_SLOT (CHUNK / 2 + 16) // multi-target reads a chunk can own (each has at least two records)
#define CW_HEAD 0x80000000u
#define MAX_BUCKETS 512      // padded bin ids fit 31 bits, slices are >= 2^22 bins

struct CovParams {
    const uint4 *meta; const uint2 *meta2 /* {len, bin offset lo}: half the footprint, for histograms of fewer than 2^31 bins */; u32 G, half_avg; BinDiv wdiv;
    unsigned long long *hist;       // MODE 0
    u32 *items, *bucket_cnt; u32 shift, n_buckets;   // MODE 1
    u32 *cw, *cw_idx; uint2 *chunk_cnt; u32 *lr;     // compact stream of the multi-mapped reads for k_assign
    u32 *rs;                                         // per chunk, one entry per multi-target read: start inside the chunk's compact words | words << 16
    unsigned char *res_kind;        // optional per-read results: marks the head of every single-target read
    DevScalars *sc;
    cudaTextureObject_t meta2_tex;                   // meta2 behind the texture path (k_coverage_tile, TEXG)
};

// contribution of one record: direct RED into the interleaved histogram, or a 32-bit item
template <int MODE>
__device__ __forceinline__ void emit(const CovParams &P, u32 i, u64 b, bool first, bool multi)
{
    if (MODE == 0) {
        if (first) atomicAdd(P.hist + b, multi ? 1ull : 0x100000001ull);   // cov += 1 [, uniq_cov += 1]
    } else {
        __stcs(P.items + i, first ? ((u32)b | (multi ? 0u : 0x80000000u)) : ITEM_SKIP);
    }
}

// ------------------------------------------------------------------------------------------------
// K1 coverage.  Replaces reference src/slimm.hpp:194-257 + src/read_stat.hpp:116-135,72-75.
// Per record: head (first record of its read), first (first record of its (read, ref) pair in file
// order - repeat hits are dropped, src/read_stat.hpp:125-131), multi (the read names another
// reference as well).
//   MODE 0 (direct) : one 64-bit RED per contributing record straight into hist (histogram ~ L2-sized)
//   MODE 1 (items)  : items[i] = padded bin | uniq << 31 (ITEM_SKIP for a repeat hit), in record order,
//                     plus the number of items per histogram slice ("bucket"); k_split then groups the
//                     items by slice and k_accumulate_fused applies them one L2-resident slice after the
//                     other, so the random read-modify-writes never reach HBM
// Both modes also write the COMPACT STREAM k_assign works on: for every read with several targets its
// distinct reference ids in file order (bit 31 marks the first), chunk by chunk; the reads with one target
// (the majority) never come back.  Runs longer than 32 records are listed by their start instead.
// ------------------------------------------------------------------------------------------------
template <class Rec, int MODE>
__device__ __noinline__ u32 coverage_long_run(const Rec &rec, u32 p, u32 n, u32 lane, const CovParams &P, u32 *s_cnt, u32 &uniq,
                                              u32 &bad, u32 lr_base, u32 &n_lr)
{
    const u32 r0 = rec.read(p), gh = rec.refid(p);
    bool multi = false;
    u32 end = p;
    for (u32 q = p;; q += 32) {                                    // pass 1: where the run ends, one reference or several
        const u32 i = q + lane;
        const bool in = i < n && rec.read(i) == r0;
        const u32 inb = __ballot_sync(FULL, in);
        multi |= __any_sync(FULL, in && rec.refid(i) != gh);
        if (inb != FULL) { end = q + (inb == 0 ? 0 : 32 - __clz(inb)); break; }
    }
    if (end < n && rec.read(end) < r0) bad |= 1u;
    if (lane == 0) {
        uniq += !multi;
        if (multi) { if (n_lr < LR_SLOT) P.lr[lr_base + n_lr] = p; ++n_lr; }
        else if (P.res_kind) P.res_kind[p] = 3;
    }
    n_lr = __shfl_sync(FULL, n_lr, 0);
    for (u32 q = p; q < end; q += 32) {                            // pass 2: first-occurrence test against the run so far
        const u32 i = q + lane;
        const bool in = i < end;
        u32 g = 0;
        bool first = false, ok = false;
        u64 b = 0;
        if (in) {
            g = rec.refid(i);
            ok = g < P.G;
            first = multi ? true : i == p;
            if (multi) for (u32 j = p; j < i; ++j) if (rec.refid(j) == g) { first = false; break; }
            if (!ok) bad |= 2u;
            else {
                b = bin_of(P.meta, g, rec.upos(i), P.half_avg, P.wdiv);
                emit<MODE>(P, i, b, first, multi);
            }
        }
        if (MODE == 1) {
            if (in && ok && first) atomicAdd(&s_cnt[(u32)(b >> P.shift)], 1u);
            if (in && !ok) __stcs(P.items + i, ITEM_SKIP);
        }
    }
    return end;
}

// EXTRA: the optional outputs (record index per compact word, per-read result marks) are wanted
template <class Rec, int MODE, bool EXTRA>
__global__ void __launch_bounds__(256, 6)
k_coverage(Rec rec, u32 n, CovParams P)
{
    __shared__ u32 s_cnt[MODE ? MAX_BUCKETS : 1];                  // items per histogram slice (shared-memory REDs)
    __shared__ u32 s_h, s_u, s_b;
    const u32 tid = threadIdx.x, lane = tid & 31;
    if (MODE) for (u32 b = tid; b < MAX_BUCKETS; b += 256) s_cnt[b] = 0;
    if (tid == 0) { s_h = 0; s_u = 0; s_b = 0; }
    __syncthreads();
    u32 heads = 0, uniq = 0, bad = 0;
    const u32 n_chunks = n / CHUNK + (n % CHUNK != 0);
    const u32 wg = (blockIdx.x * blockDim.x + tid) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (u32 c = wg; c < n_chunks; c += nw) {
        const u32 c0 = c * CHUNK, c1 = min(c0 + CHUNK, n);
        u32 p = find_head(rec, c0, n, lane);
        u32 n_cw = 0, n_lr = 0, n_rs = 0;                          // compact words / long runs / multi-target reads of this chunk (warp-uniform)
        u32 *const cw_c = P.cw + (u64)c * CW_SLOT;
        u32 *const cwi_c = EXTRA && P.cw_idx ? P.cw_idx + (u64)c * CW_SLOT : nullptr;
        u32 *const rs_c = P.rs + (u64)c * RS_SLOT;
        WinRegs cur = fetch_window(rec, p, n, lane);
        while (p < c1) {
            Window win;
            analyse_window(cur, p, n, c1, lane, win, bad);
            if (win.long_run) {
                if (lane == 0) ++heads;
                p = coverage_long_run<Rec, MODE>(rec, p, n, lane, P, s_cnt, uniq, bad, c * LR_SLOT, n_lr);
                cur = fetch_window(rec, p, n, lane);
                continue;
            }
            const u32 g = win.g;
            const bool ok = g < P.G;
            const uint4 meta = __ldg(P.meta + (ok ? g : 0u));       // issued now, used after the run analysis
            const u32 upos = cur.upos;
            cur = fetch_window(rec, win.next, n, lane);            // the next window's loads fly during this window's work
            const u32 gh = __shfl_sync(FULL, g, win.whole ? win.s : (int)lane);
            const u32 Wm = __ballot_sync(FULL, win.whole && g != gh);
            const bool multi = (Wm & win.M) != 0;                  // the read names another reference as well
            const bool is_head = win.whole && (int)lane == win.s;
            bool first = is_head;
            if (Wm) {                                              // some read with several references: repeat hits need a look
                const int dist = multi ? (int)lane - win.s : 0;    // records of my read before me
                const int span = (int)__reduce_max_sync(FULL, (u32)dist);
                bool rep = false;
                for (int d = 1; d <= span; d += 2) {               // the same reference earlier in my read?  (two steps per trip)
                    const u32 t1 = __shfl_up_sync(FULL, g, d), t2 = __shfl_up_sync(FULL, g, d + 1);
                    rep |= (d <= dist && t1 == g) | (d < dist && t2 == g);
                }
                first = multi ? !rep : is_head;
                const u32 C = __ballot_sync(FULL, multi && first); // the compact stream keeps the distinct references
                const u32 Hm = __ballot_sync(FULL, multi && is_head);
                if (multi && first) {
                    const u32 local = n_cw + __popc(C & LANE_LT(lane));
                    cw_c[local] = (ok ? g : 0u) | (is_head ? CW_HEAD : 0u);   // an id out of range is reported later; the word stays harmless
                    if (EXTRA && cwi_c) cwi_c[local] = p + lane;
                    if (is_head) rs_c[n_rs + __popc(Hm & LANE_LT(lane))] = local | ((u32)__popc(C & win.M) << 16);
                }
                n_cw += __popc(C);
                n_rs += __popc(Hm);
            }
            if (is_head) {
                ++heads; uniq += !multi;
                if (EXTRA && P.res_kind && !multi) P.res_kind[p + lane] = 3;
            }
            if (win.whole) {
                if (!ok) bad |= 2u;
                const bool put = ok && first;
                const u64 b = bin_of_meta(meta, upos, P.half_avg, P.wdiv);
                if (MODE == 0) { if (put) atomicAdd(P.hist + b, multi ? 1ull : 0x100000001ull); }   // cov += 1 [, uniq_cov += 1]
                else {
                    __stcs(P.items + p + lane, put ? ((u32)b | (multi ? 0u : 0x80000000u)) : ITEM_SKIP);
                    if (put) atomicAdd(&s_cnt[(u32)(b >> P.shift)], 1u);
                }
            }
            p = win.next;
        }
        if (lane == 0) P.chunk_cnt[c] = make_uint2(n_cw, min(n_lr, 0xFFu) | (n_rs << 8));
    }
    heads = warp_sum(heads); uniq = warp_sum(uniq); bad = warp_or(bad);
    if (lane == 0) { atomicAdd(&s_h, heads); atomicAdd(&s_u, uniq); if (bad) atomicOr(&s_b, bad); }
    __syncthreads();
    if (MODE)
        for (u32 b = tid; b < P.n_buckets; b += 256)
            if (s_cnt[b]) atomicAdd(P.bucket_cnt + b, s_cnt[b]);
    if (tid == 0) {
        if (s_h) atomicAdd(&P.sc->n_reads, (unsigned long long)s_h);
        if (s_u) atomicAdd(&P.sc->n_uniq, (unsigned long long)s_u);
        if (s_b) atomicOr(&P.sc->flags, s_b);
    }
}

#include "coverage_tile.cuh"

// Per-slice bookkeeping shared by k_coverage (count), k_bucket_scan, k_split (cursor) and k_accumulate.
struct Sched {
    u32 total_items, pad[3];
    u32 cursor[MAX_BUCKETS];     // k_split's write cursors (start of every slice's items, advanced by the split)
    u32 start[MAX_BUCKETS];      // start of every slice's items
    u32 count[MAX_BUCKETS];      // items per slice (filled by k_coverage)
};

// exclusive scan of the slice sizes -> item starts / write cursors (one block)
__global__ void __launch_bounds__(MAX_BUCKETS) k_bucket_scan(Sched *sd, u32 n_buckets)
{
    __shared__ u32 s[MAX_BUCKETS];
    const u32 tid = threadIdx.x;
    const u32 v = tid < n_buckets ? sd->count[tid] : 0;
    s[tid] = v;
    __syncthreads();
    for (u32 d = 1; d < MAX_BUCKETS; d <<= 1) {
        const u32 t = tid >= d ? s[tid - d] : 0;
        __syncthreads();
        s[tid] += t;
        __syncthreads();
    }
    if (tid < n_buckets) sd->cursor[tid] = sd->start[tid] = s[tid] - v;
    if (tid == MAX_BUCKETS - 1) sd->total_items = s[tid];          // records minus repeat hits
}

// ------------------------------------------------------------------------------------------------
// K1b multisplit: groups the items by histogram slice.  A CTA ranks a tile of 256 x SPLIT_ITEMS items with
// one returning shared-memory atomic per item (measured on B200: they run at streaming-read speed, while
// MATCH.ANY over ~400 distinct slices is 8x slower), orders the tile by slice in shared memory and copies
// every slice's share to its reserved place in the output, so the global stores are runs of consecutive words.
// ------------------------------------------------------------------------------------------------
#define SPLIT_TILE 8192
// PEER: every slice has its own destination (dest[slice]: where THIS rank's items of the slice go inside the receive
// buffer of the rank that owns the slice - local memory or a peer's, mapped over NVLink): the all-to-all of a sharded
// run happens inside the split, tile by tile, as plain stores.
// NT threads x (SPLIT_TILE / NT) items: the same tile with fewer registers per thread buys resident warps (256 x 32: 92 registers,
// 16 warps per SM; 512 x 16: 32 warps per SM).
// a tile's items into registers: 128-bit loads for whole tiles of an aligned array
template <int NT, int ITEMS>
__device__ __forceinline__ void load_tile(const u32 *__restrict__ items, u64 t0, u32 n, bool vec, u32 tid, u32 (&item)[ITEMS])
{
    if (vec && t0 + (u64)NT * ITEMS <= n) {
        const uint4 *p = reinterpret_cast<const uint4 *>(items + t0) + tid;
#pragma unroll
        for (int k = 0; k < ITEMS / 4; ++k) {
            const uint4 v = __ldcs(p + k * NT);
            item[4 * k] = v.x; item[4 * k + 1] = v.y; item[4 * k + 2] = v.z; item[4 * k + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const u64 i = t0 + (u64)k * NT + tid;
            item[k] = i < n ? __ldcs(items + i) : ITEM_SKIP;
        }
    }
}

// Branch-free ranking: ITEM_SKIP (repeat hits, the padding of the last tile) is ranked too, in a slot of its own - slot
// min(0x7FFFFFFF >> shift, MAX_BUCKETS), which lies beyond the slices in use unless the bins reach the last 2^shift below 2^31
// (EXACT then tests for ITEM_SKIP explicitly) - and staged behind the tile's other items, where the copy loop never looks.
template <bool EXACT>
__device__ __forceinline__ u32 split_slot(u32 item, u32 shift)
{
    if (EXACT) return item == ITEM_SKIP ? (u32)MAX_BUCKETS : (item & 0x7FFFFFFFu) >> shift;
    return min((item << 1) >> (shift + 1), (u32)MAX_BUCKETS);
}

template <bool PEER, int NT, bool EXACT>
__global__ void __launch_bounds__(NT, NT >= 1024 ? 1 : 2)
k_split(const u32 *__restrict__ items, u32 n, u32 shift, u32 n_buckets, Sched *sd, u32 *__restrict__ out, u32 *const *__restrict__ dest,
        const u32 *__restrict__ n_ptr /* not null: the number of items lives on the device */)
{
    if (n_ptr) n = *n_ptr;
    constexpr int ITEMS = SPLIT_TILE / NT, NW = NT / 32;
    constexpr int SCAN_T = NT < MAX_BUCKETS ? NT : MAX_BUCKETS;    // threads that take part in the scan of the slice counts
    constexpr int HALVES = MAX_BUCKETS / SCAN_T;
    __shared__ u32 s_cnt[MAX_BUCKETS + 1];     // items of each slice in this tile (+ the slot of ITEM_SKIP), then the slice's tile-local start
    __shared__ u32 s_delta[MAX_BUCKETS];       // global start - tile-local start (mod 2^32)
    __shared__ u32 *s_dst[PEER ? MAX_BUCKETS : 1];   // PEER: destination of the slice's first item of this tile, minus the tile-local start
    __shared__ u32 s_item[SPLIT_TILE];
    __shared__ u32 s_warp_tot[NW];
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool vec = (reinterpret_cast<uintptr_t>(items) & 15) == 0;
    const u32 skip_slot = EXACT ? (u32)MAX_BUCKETS : min(0x7FFFFFFFu >> shift, (u32)MAX_BUCKETS);
    const u64 n_tiles = ((u64)n + SPLIT_TILE - 1) / SPLIT_TILE;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (u32 b = tid; b <= MAX_BUCKETS; b += NT) s_cnt[b] = 0;
        __syncthreads();
        const u64 t0 = tile * SPLIT_TILE;
        u32 item[ITEMS], where[ITEMS];                         // where: slot << 16 | rank inside (tile, slot)
        load_tile<NT, ITEMS>(items, t0, n, vec, tid, item);
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const u32 bk = split_slot<EXACT>(item[k], shift);
            where[k] = (bk << 16) | atomicAdd(&s_cnt[bk], 1u);
        }
        __syncthreads();
        // exclusive scan of the slice counts over the tile (HALVES slices per scanning thread, halves in order)
        u32 tot[HALVES], excl[HALVES], base = 0;
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
            const u32 b = tid + h * SCAN_T;
            tot[h] = (tid < SCAN_T && b < n_buckets) ? s_cnt[b] : 0u;
            u32 x = tot[h];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL, x, o); if ((int)lane >= o) x += y; }
            if (lane == 31) s_warp_tot[wid] = x;
            __syncthreads();
            u32 wbase = 0, all = 0;
#pragma unroll
            for (int k = 0; k < SCAN_T / 32; ++k) { const u32 t = s_warp_tot[k]; if (k < (int)wid) wbase += t; all += t; }
            excl[h] = base + wbase + x - tot[h];
            base += all;
            __syncthreads();
        }
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
            const u32 b = tid + h * SCAN_T;
            if (tid < SCAN_T && b < n_buckets) {
                s_cnt[b] = excl[h];
                if (tot[h]) {
                    const u32 at = atomicAdd(&sd->cursor[b], tot[h]);
                    s_delta[b] = at - excl[h];
                    if (PEER) s_dst[b] = dest[b] + (at - sd->start[b]) - excl[h];
                }
            }
        }
        if (tid == 0) s_cnt[skip_slot] = base;                 // ITEM_SKIP is staged behind the items
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) s_item[s_cnt[where[k] >> 16] + (where[k] & 0xFFFFu)] = item[k];
        __syncthreads();
        const u32 total = base;                                // items of this tile (thread-uniform)
        if (!PEER && total == SPLIT_TILE) {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const u32 j = k * NT + tid, v = s_item[j];
                out[j + s_delta[(v << 1) >> (shift + 1)]] = v;
            }
        } else
            for (u32 j = tid; j < total; j += NT) {            // the slice of an item is a function of the item
                const u32 v = s_item[j], bk = (v << 1) >> (shift + 1);
                if (PEER) s_dst[bk][j] = v; else out[j + s_delta[bk]] = v;
            }
        __syncthreads();
    }
}

// Where THIS rank's items of every slice go (sharded runs, peer-to-peer exchange), computed on the device from the all-gathered
// slice counts, so that no host round trip sits between the coverage stage and the split: the receive buffer of owner q holds
// its slices in ascending order, inside a slice the source ranks in ascending order (slimm_gpu_split_to_peers has the host twin).
// One CTA of MAX_BUCKETS threads.
__global__ void __launch_bounds__(MAX_BUCKETS)
k_peer_dest(const u32 *__restrict__ all_counts /*[n_ranks][n_slices]*/, u32 n_slices, u32 n_ranks, u32 me, u32 *const *__restrict__ peer_recv /*[n_ranks]*/,
            u64 recv_cap, u32 **__restrict__ dest /*[MAX_BUCKETS]*/, u32 *__restrict__ n_recv, u32 *__restrict__ overflow)
{
    __shared__ u32 s[MAX_BUCKETS + 1];
    const u32 tid = threadIdx.x;
    u32 tot = 0, before = 0;                   // items of slice tid from all ranks / from the ranks below me
    if (tid < n_slices)
        for (u32 src = 0; src < n_ranks; ++src) { const u32 c = all_counts[(size_t)src * n_slices + tid]; if (src < me) before += c; tot += c; }
    s[tid + 1] = tot;
    if (tid == 0) s[0] = 0;
    __syncthreads();
    for (u32 d = 1; d < MAX_BUCKETS; d <<= 1) {                    // inclusive scan of s[1..]
        const u32 t = tid + 1 > d ? s[tid + 1 - d] : 0;
        __syncthreads();
        s[tid + 1] += t;
        __syncthreads();
    }
    if (tid < n_slices) {
        u32 q = 0;                                                 // owner of slice tid: slices [ns q / n, ns (q+1) / n)
        while ((u64)n_slices * (q + 1) / n_ranks <= tid) ++q;
        const u32 lo = (u32)((u64)n_slices * q / n_ranks);
        dest[tid] = peer_recv[q] + (s[tid] - s[lo]) + before;
    }
    if (tid < n_ranks) {
        const u32 lo = (u32)((u64)n_slices * tid / n_ranks), hi = (u32)((u64)n_slices * (tid + 1) / n_ranks);
        if ((u64)(s[hi] - s[lo]) > recv_cap) atomicOr(overflow, 1u);
        if (tid == me) *n_recv = s[hi] - s[lo];
    }
}

// The same multisplit with the tile brought in by the bulk-copy engine (cp.async.bulk, completion on an mbarrier) and double
// buffered: while a tile is ranked, ordered and copied out, the next 32 KB are already on their way into shared memory, and no
// register holds an item while it waits for DRAM (16 ranks per thread are all that stays live).  Local splits only.
#define SPLITB_NT 512
#define SPLITB_SMEM (2 * SPLIT_TILE * 4 + SPLIT_TILE * 4)      // two raw tiles + the ordered tile (dynamic shared memory)
__global__ void __launch_bounds__(SPLITB_NT, 2)
k_split_bulk(const u32 *__restrict__ items, u32 n, u32 shift, u32 n_buckets, Sched *sd, u32 *__restrict__ out, const u32 *__restrict__ n_ptr)
{
    using barrier_t = cuda::barrier<cuda::thread_scope_block>;
    constexpr int NT = SPLITB_NT, ITEMS = SPLIT_TILE / NT, NW = NT / 32;
    extern __shared__ __align__(128) u32 sb_dyn[];
    u32 *const raw0 = sb_dyn, *const raw1 = sb_dyn + SPLIT_TILE, *const s_item = sb_dyn + 2 * SPLIT_TILE;
    __shared__ u32 s_cnt[MAX_BUCKETS];
    __shared__ u32 s_delta[MAX_BUCKETS];
    __shared__ u32 s_warp_tot[NW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier_t bar[2];
    if (n_ptr) n = *n_ptr;
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        init(&bar[0], NT); init(&bar[1], NT);
        cuda::device::experimental::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    const u64 n_tiles = ((u64)n + SPLIT_TILE - 1) / SPLIT_TILE, n_full = (u64)n / SPLIT_TILE;   // tiles, whole tiles (bulk copies)
    barrier_t::arrival_token tok0, tok1;
    // every thread arrives on the buffer's barrier when its copy is issued (thread 0 adds the bytes to expect) and waits on it
    // when the tile is needed
    auto issue = [&](u64 tile, u32 *dst, barrier_t &b) -> barrier_t::arrival_token {
        if (tile >= n_full) return b.arrive();                   // the last, partial tile is loaded by the threads themselves
        if (tid == 0) {
            cuda::device::experimental::fence_proxy_async_shared_cta();   // the buffer's earlier readers come first
            cuda::device::memcpy_async_tx(dst, items + tile * SPLIT_TILE, cuda::aligned_size_t<16>(SPLIT_TILE * 4), b);
            return cuda::device::barrier_arrive_tx(b, 1, SPLIT_TILE * 4);
        }
        return b.arrive();
    };
    u64 tile = blockIdx.x;
    if (tile < n_tiles) tok0 = issue(tile, raw0, bar[0]);
    for (u32 it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
        const bool odd = it & 1;
        u32 *const raw = odd ? raw1 : raw0;
        const u64 next = tile + gridDim.x;
        if (next < n_tiles) { if (odd) tok0 = issue(next, raw0, bar[0]); else tok1 = issue(next, raw1, bar[1]); }
        for (u32 b = tid; b < MAX_BUCKETS; b += NT) s_cnt[b] = 0;
        if (odd) bar[1].wait(std::move(tok1)); else bar[0].wait(std::move(tok0));
        const u64 t0 = tile * SPLIT_TILE;
        if (tile >= n_full) {                                    // partial tile
            for (u32 j = tid; j < SPLIT_TILE; j += NT) raw[j] = t0 + j < n ? __ldcs(items + t0 + j) : ITEM_SKIP;
        }
        __syncthreads();
        u32 where[ITEMS];                                        // slice << 16 | rank inside (tile, slice)
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const u32 v = raw[k * NT + tid];
            where[k] = 0xFFFFFFFFu;
            if (v != ITEM_SKIP) {
                const u32 bk = (v & 0x7FFFFFFFu) >> shift;
                where[k] = (bk << 16) | atomicAdd(&s_cnt[bk], 1u);
            }
        }
        __syncthreads();
        // exclusive scan of the slice counts over the tile: one slice per thread
        const u32 tot = tid < n_buckets ? s_cnt[tid] : 0u;
        u32 x = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL, x, o); if ((int)lane >= o) x += y; }
        if (lane == 31) s_warp_tot[wid] = x;
        __syncthreads();
        u32 wbase = 0, total = 0;
#pragma unroll
        for (int k = 0; k < NW; ++k) { const u32 t = s_warp_tot[k]; if (k < (int)wid) wbase += t; total += t; }
        const u32 excl = wbase + x - tot;
        if (tid < n_buckets) {
            s_cnt[tid] = excl;
            if (tot) s_delta[tid] = atomicAdd(&sd->cursor[tid], tot) - excl;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ITEMS; ++k)
            if (where[k] != 0xFFFFFFFFu) s_item[s_cnt[where[k] >> 16] + (where[k] & 0xFFFFu)] = raw[k * NT + tid];
        __syncthreads();
        for (u32 j = tid; j < total; j += NT) {                  // the slice of an item is a function of the item
            const u32 v = s_item[j];
            out[j + s_delta[(v & 0x7FFFFFFFu) >> shift]] = v;
        }
        __syncthreads();
    }
}

// ---- routed exchange (sharded runs, the default over NVLink) ---------------------------------------------------------
// Peer stores of 80-byte runs (k_split<PEER>: a tile's share of ONE slice) use a third of the link.  Routed, a tile is only
// partitioned by OWNER RANK: a rank's share of a tile is one contiguous kilobyte-sized segment, appended to the region the
// receiver keeps for this source - full 128-byte stores.  The receiver then groups what it received by slice with the
// local k_split (its runs are long: only the slices it owns occur), from slice totals it already knows from the all-gathered
// counts.  k_peer_route_plan derives everything from that table on the device; no host round trip.
#define ROUTE_MAX_RANKS 32
struct RoutePlan {
    u32 *dest[ROUTE_MAX_RANKS];      // where THIS rank's items for rank q start inside q's receive buffer
    u32 cursor[ROUTE_MAX_RANKS];     // items appended so far (k_route)
    unsigned char owner[MAX_BUCKETS];
};

__global__ void __launch_bounds__(MAX_BUCKETS)
k_peer_route_plan(const u32 *__restrict__ all_counts /*[n_ranks][n_slices]*/, u32 n_slices, u32 n_ranks, u32 me, u32 *const *__restrict__ peer_recv,
                  u64 recv_cap, RoutePlan *__restrict__ plan, Sched *__restrict__ local /* the receiver's split of what it gets */,
                  u32 *__restrict__ n_recv, u32 *__restrict__ overflow)
{
    __shared__ u32 s[MAX_BUCKETS + 1];
    __shared__ u32 s_from[ROUTE_MAX_RANKS][ROUTE_MAX_RANKS];       // [q][src]: items src sends to q
    const u32 tid = threadIdx.x;
    u32 q_of = 0;                                                  // owner of slice tid: slices [ns q / n, ns (q+1) / n)
    if (tid < n_slices) while ((u64)n_slices * (q_of + 1) / n_ranks <= tid) ++q_of;
    if (tid < n_slices) plan->owner[tid] = (unsigned char)q_of;
    for (u32 i = tid; i < ROUTE_MAX_RANKS * ROUTE_MAX_RANKS; i += MAX_BUCKETS) (&s_from[0][0])[i] = 0;
    __syncthreads();
    u32 mine = 0;                                                  // items of slice tid this rank will hold (0 for slices it does not own)
    if (tid < n_slices)
        for (u32 src = 0; src < n_ranks; ++src) {
            const u32 c = all_counts[(size_t)src * n_slices + tid];
            if (c) atomicAdd(&s_from[q_of][src], c);
            if (q_of == me) mine += c;
        }
    s[tid + 1] = mine;
    if (tid == 0) s[0] = 0;
    __syncthreads();
    for (u32 d = 1; d < MAX_BUCKETS; d <<= 1) {                    // inclusive scan of s[1..]
        const u32 t = tid + 1 > d ? s[tid + 1 - d] : 0;
        __syncthreads();
        s[tid + 1] += t;
        __syncthreads();
    }
    local->cursor[tid] = local->start[tid] = s[tid];
    local->count[tid] = mine;
    if (tid == 0) { local->total_items = s[MAX_BUCKETS]; *n_recv = s[MAX_BUCKETS]; }
    if (tid < n_ranks) {
        u32 before = 0, all = 0;
        for (u32 src = 0; src < n_ranks; ++src) { if (src < me) before += s_from[tid][src]; all += s_from[tid][src]; }
        plan->dest[tid] = peer_recv[tid] + before;
        plan->cursor[tid] = 0;
        if ((u64)all > recv_cap) atomicOr(overflow, 1u);
    }
}

template <int NT>
__global__ void __launch_bounds__(NT, 2)
k_route(const u32 *__restrict__ items, u32 n, u32 shift, u32 n_ranks, RoutePlan *__restrict__ plan)
{
    constexpr int ITEMS = SPLIT_TILE / NT, NW = NT / 32;
    __shared__ u32 s_wcnt[NW][ROUTE_MAX_RANKS];                    // items of a (warp, rank) in this tile, then their tile-local start
    __shared__ u32 s_seg[ROUTE_MAX_RANKS + 1];                     // tile-local start of every rank's segment
    __shared__ u32 s_at[ROUTE_MAX_RANKS];                          // where the segment goes inside the receiver's region for this source
    __shared__ u32 s_warp_tot[NW];
    __shared__ unsigned char s_owner[MAX_BUCKETS];
    __shared__ u32 s_item[SPLIT_TILE];
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (u32 b = tid; b < MAX_BUCKETS; b += NT) s_owner[b] = plan->owner[b];
    const u64 n_tiles = ((u64)n + SPLIT_TILE - 1) / SPLIT_TILE;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (u32 i = tid; i < NW * ROUTE_MAX_RANKS; i += NT) (&s_wcnt[0][0])[i] = 0;
        __syncthreads();
        const u64 t0 = tile * SPLIT_TILE;
        u32 item[ITEMS], where[ITEMS];                             // where: rank << 16 | rank inside (tile, warp, rank)
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const u64 i = t0 + (u64)k * NT + tid;
            item[k] = i < n ? __ldcs(items + i) : ITEM_SKIP;
        }
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            where[k] = 0xFFFFFFFFu;
            if (item[k] != ITEM_SKIP) {
                const u32 q = s_owner[(item[k] & 0x7FFFFFFFu) >> shift];
                where[k] = (q << 16) | atomicAdd(&s_wcnt[wid][q], 1u);   // a warp's own counters: no contention between warps
            }
        }
        __syncthreads();
        {   // exclusive scan over (rank major, warp minor): thread t <-> rank t / NW, warp t % NW
            const u32 q = tid / NW, w = tid % NW;
            const u32 v = q < n_ranks ? s_wcnt[w][q] : 0u;
            u32 x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL, x, o); if ((int)lane >= o) x += y; }
            if (lane == 31) s_warp_tot[wid] = x;
            __syncthreads();
            u32 wbase = 0, all = 0;
#pragma unroll
            for (int k = 0; k < NW; ++k) { const u32 t = s_warp_tot[k]; if (k < (int)wid) wbase += t; all += t; }
            const u32 excl = wbase + x - v;
            if (q < n_ranks) { s_wcnt[w][q] = excl; if (w == 0) s_seg[q] = excl; }
            if (tid == 0) s_seg[n_ranks] = all;
        }
        __syncthreads();
        if (tid < n_ranks) {
            const u32 len = s_seg[tid + 1] - s_seg[tid];
            s_at[tid] = len ? atomicAdd(&plan->cursor[tid], len) : 0u;
        }
#pragma unroll
        for (int k = 0; k < ITEMS; ++k)
            if (where[k] != 0xFFFFFFFFu) s_item[s_wcnt[wid][where[k] >> 16] + (where[k] & 0xFFFFu)] = item[k];
        __syncthreads();
        for (u32 q = 0; q < n_ranks; ++q) {                        // a rank's segment: one contiguous run of full-width stores (peer memory over NVLink)
            const u32 a = s_seg[q], b = s_seg[q + 1];
            u32 *dst = plan->dest[q] + s_at[q];
            for (u32 j = a + tid; j < b; j += NT) dst[j - a] = s_item[j];
        }
        __syncthreads();
    }
}

// ---- block exchange (sharded runs; SLIMM_PEER_ROUTE=2) ------------------------------------------------------------------
// Routed, every item is ranked twice (by owner rank before it travels, by slice after) and the sender's ranking pass runs at
// link speed at best.  Here the sender groups its items by slice exactly as a single GPU does (k_split into local memory, HBM
// speed); a rank owns CONSECUTIVE slices, so its share of the grouped items is one contiguous block, and k_peer_copy streams
// the n blocks to their owners (128-byte-aligned stores over NVLink, all links at once).  The receiver holds the blocks source
// by source, each ordered by slice - what k_fine_count / k_fine_split need (a tile spans one or two coarse slices; the few
// tiles across two blocks go through their overflow paths) - so there is no second coarse pass at all.
struct CopyPlan {
    u32 *dst[ROUTE_MAX_RANKS];       // where THIS rank's block for rank q starts inside q's receive buffer
    u32 src_off[ROUTE_MAX_RANKS];    // where the block starts inside this rank's grouped items
    u32 len[ROUTE_MAX_RANKS];
};

__global__ void __launch_bounds__(MAX_BUCKETS)
k_peer_copy_plan(const u32 *__restrict__ all_counts /*[n_ranks][n_slices]*/, u32 n_slices, u32 n_ranks, u32 me, u32 *const *__restrict__ peer_recv,
                 u64 recv_cap, CopyPlan *__restrict__ plan, u32 *__restrict__ n_recv, u32 *__restrict__ overflow)
{
    __shared__ u32 s_from[ROUTE_MAX_RANKS][ROUTE_MAX_RANKS];       // [q][src]: items src sends to q
    const u32 tid = threadIdx.x;
    u32 q_of = 0;                                                  // owner of slice tid: slices [ns q / n, ns (q+1) / n)
    if (tid < n_slices) while ((u64)n_slices * (q_of + 1) / n_ranks <= tid) ++q_of;
    for (u32 i = tid; i < ROUTE_MAX_RANKS * ROUTE_MAX_RANKS; i += MAX_BUCKETS) (&s_from[0][0])[i] = 0;
    __syncthreads();
    if (tid < n_slices)
        for (u32 src = 0; src < n_ranks; ++src) {
            const u32 c = all_counts[(size_t)src * n_slices + tid];
            if (c) atomicAdd(&s_from[q_of][src], c);
        }
    __syncthreads();
    if (tid < n_ranks) {
        u32 before = 0, all = 0, off = 0;
        for (u32 src = 0; src < n_ranks; ++src) { if (src < me) before += s_from[tid][src]; all += s_from[tid][src]; }
        for (u32 q = 0; q < tid; ++q) off += s_from[q][me];
        plan->dst[tid] = peer_recv[tid] + before;
        plan->src_off[tid] = off;
        plan->len[tid] = s_from[tid][me];
        if ((u64)all > recv_cap) atomicOr(overflow, 1u);
        if (tid == me) *n_recv = (u64)all > recv_cap ? 0u : all;    // (on overflow nothing is copied and nothing is read)
    }
}

#define PCOPY_CHUNK 8192u
__global__ void __launch_bounds__(256)
k_peer_copy(const u32 *__restrict__ grouped, const CopyPlan *__restrict__ plan, u32 n_ranks, u32 me, const u32 *__restrict__ overflow)
{
    if (*overflow) return;                                         // some receive buffer is too small (every rank sees that from the table): nobody
                                                                   // writes, slimm_gpu_get_summary reports SLIMM_GPU_ERANGE on every rank
    const u32 tid = threadIdx.x;
    for (u32 k = 0; k < n_ranks; ++k) {
        const u32 q = (me + 1 + k + blockIdx.x) % n_ranks;         // CTAs start on different links
        u32 *__restrict__ dst = plan->dst[q];
        const u32 *__restrict__ src = grouped + plan->src_off[q];
        const u32 len = plan->len[q];
        // chunks are cut where the DESTINATION is 128-byte aligned: whole lines over the link
        const u32 head = min(len, (32u - (u32)((reinterpret_cast<uintptr_t>(dst) >> 2) & 31u)) & 31u);
        if (blockIdx.x == 0 && tid < head) dst[tid] = src[tid];
        const u32 body = len - head, n_chunks = (body + PCOPY_CHUNK - 1) / PCOPY_CHUNK;
        for (u32 c = blockIdx.x; c < n_chunks; c += gridDim.x) {
            const u32 o = head + c * PCOPY_CHUNK, m = min(PCOPY_CHUNK, len - o);
            u32 v[PCOPY_CHUNK / 256];
#pragma unroll
            for (u32 j = 0; j < PCOPY_CHUNK / 256; ++j) if (j * 256 + tid < m) v[j] = __ldcs(src + o + j * 256 + tid);
#pragma unroll
            for (u32 j = 0; j < PCOPY_CHUNK / 256; ++j) if (j * 256 + tid < m) dst[o + j * 256 + tid] = v[j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2 accumulate: applies the grouped items in stream order, one 64-bit RED each: the CTAs in flight work
// on one or two adjacent histogram slices, so the REDs meet in L2 and a slice's sectors travel HBM -> L2 ->
// HBM once.  (Zero-filling the slices in L2 right before their REDs - one fused persistent kernel with
// zeroer and accumulator CTAs - was measured too, scripts/micro/fused_bench.cu: no faster than this kernel
// behind a memset, and the memset overlaps k_coverage / k_split on a second stream for free.)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_accumulate(const u32 *__restrict__ grouped, const Sched *__restrict__ sd, unsigned long long *__restrict__ hist, u32 per_block,
             u32 n_given /* used when sd == nullptr: items received from other ranks */)
{
    const u32 n_items = sd ? sd->total_items : n_given;
    if (per_block == 0) {                                          // grid-stride
        const u32 stride = gridDim.x * blockDim.x;
        for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items && i + stride > i; i += stride) {
            const u32 v = __ldcs(grouped + i);
            atomicAdd(hist + (v & 0x7FFFFFFFu), (v >> 31) ? 0x100000001ull : 1ull);
        }
        return;
    }
    const u64 first = (u64)blockIdx.x * per_block, last = min((u64)n_items, first + per_block);   // a block owns consecutive items
    for (u64 j0 = first; j0 < last; j0 += 256 * 8) {
        u32 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { const u64 j = j0 + (u64)k * 256 + threadIdx.x; v[k] = j < last ? __ldcs(grouped + j) : ITEM_SKIP; }
#pragma unroll
        for (int k = 0; k < 8; ++k) if (v[k] != ITEM_SKIP) atomicAdd(hist + (v[k] & 0x7FFFFFFFu), (v[k] >> 31) ? 0x100000001ull : 1ull);
    }
}

// ------------------------------------------------------------------------------------------------
// K2 in shared memory.  The slice-grouped items are split once more, into FINE slices of 2^14 bins, whose
// {cov, uniq_cov} counters (128 KB) fit one SM's shared memory: a CTA then owns a fine slice, applies its
// items with shared-memory atomics, reduces the per-reference statistics (K3) while the bins are still on
// chip and writes the slice out once - no zero-fill pass, no read-modify-write through L2, no scan pass.
//   k_fine_count  items per fine slice (tile-private counters in shared memory: a tile of slice-grouped items
//                 spans one or two coarse slices, i.e. at most FINE_REL consecutive fine slices; stragglers
//                 go through global atomics)
//   k_fine_scan   exclusive scan -> start / write cursor of every fine slice
//   k_fine_split  the multisplit of k_split, with the fine slices of the tile's coarse slices as buckets
//   k_fine_accumulate  one CTA per fine slice
// ------------------------------------------------------------------------------------------------
#define FINE_SHIFT 14
#define FINE_BINS (1u << FINE_SHIFT)
#define FINE_REL MAX_BUCKETS      // fine slices covered by a tile's shared-memory tables
#define FINE_ITEMS 16               // items per thread of a k_fine_count / k_fine_split tile (16 measured faster than 32 here, 32 faster in k_split)
#define FINE_TILE (256 * FINE_ITEMS)
#define FINE_REF_BLOCK 1024u       // fine_ref holds the reference of the first bin of every block of this many bins
#define FINE_PRE 12               // items per thread k_fine_accumulate requests before it zero-fills its bins
#ifndef FINE_PACKED_CTAS
#define FINE_PACKED_CTAS 3          // CTAs per SM of the packed k_fine_accumulate (64 KB of shared memory each)
#endif

// first fine slice of the coarse slice the tile's first item belongs to
__device__ __forceinline__ u32 fine_base(u32 first_item, u32 cshift) { return ((first_item & 0x7FFFFFFFu) >> cshift) << (cshift - FINE_SHIFT); }

// index of an item's fine slice relative to the tile's first one, clamped to FINE_REL (the overflow slot: items of a tile that
// spans more than FINE_REL fine slices, and ITEM_SKIP).  Fast form: one shift-and-subtract, one shift, one minimum - the
// difference is taken modulo 2^17, which is exact as long as there are at most FINE_WRAP_SAFE fine slices (a slice BEFORE the
// tile's first one, and ITEM_SKIP, then still land beyond FINE_REL); EXACT: the plain 32-bit difference and an explicit test.
#define FINE_WRAP_SAFE ((1u << 17) - FINE_REL)
template <bool EXACT>
__device__ __forceinline__ u32 fine_rel(u32 item, u32 base, u32 neg_base_shl /* 0 - (base << (FINE_SHIFT + 1)) */)
{
    if (EXACT) return item == ITEM_SKIP ? (u32)FINE_REL : min(((item & 0x7FFFFFFFu) >> FINE_SHIFT) - base, (u32)FINE_REL);
    return min(((item << 1) + neg_base_shl) >> (FINE_SHIFT + 1), (u32)FINE_REL);
}

template <bool EXACT>
__global__ void __launch_bounds__(256)
k_fine_count(const u32 *__restrict__ items, const u32 *__restrict__ n_ptr /* number of items on the device, or null: n_given */, u32 n_given, u32 cshift,
             u32 *__restrict__ fine_cnt)
{
    __shared__ u32 s_cnt[FINE_REL + 1];
    const u32 n = n_ptr ? *n_ptr : n_given;
    const u32 tid = threadIdx.x;
    const bool vec = (reinterpret_cast<uintptr_t>(items) & 15) == 0;
    const u64 n_tiles = ((u64)n + FINE_TILE - 1) / FINE_TILE;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (u32 b = tid; b <= FINE_REL; b += 256) s_cnt[b] = 0;
        __syncthreads();
        const u64 t0 = tile * FINE_TILE;
        const u32 base = fine_base(__ldg(items + t0), cshift), nbs = 0u - (base << (FINE_SHIFT + 1));
        u32 item[FINE_ITEMS];
        load_tile<256, FINE_ITEMS>(items, t0, n, vec, tid, item);
#pragma unroll
        for (int k = 0; k < FINE_ITEMS; ++k) atomicAdd(&s_cnt[fine_rel<EXACT>(item[k], base, nbs)], 1u);
        __syncthreads();
        if (s_cnt[FINE_REL]) {                                   // rare: a tile over more than two coarse slices (or the last, partial tile)
#pragma unroll
            for (int k = 0; k < FINE_ITEMS; ++k)
                if (item[k] != ITEM_SKIP && fine_rel<EXACT>(item[k], base, nbs) == FINE_REL) atomicAdd(fine_cnt + ((item[k] & 0x7FFFFFFFu) >> FINE_SHIFT), 1u);
        }
        for (u32 b = tid; b < FINE_REL; b += 256)
            if (s_cnt[b]) atomicAdd(fine_cnt + base + b, s_cnt[b]);
        __syncthreads();
    }
}

// start[f] = cursor[f] = items before fine slice f; start[n_fine] = all items.  One cluster of FINE_SCAN_CL CTAs: every warp
// owns a contiguous chunk of the counts - chunk totals first, exchanged through distributed shared memory, then the scan with
// the totals of the chunks before as carry; a lane takes 16 consecutive counts per step with 128-bit loads and stores (the
// stores of one SM alone would take longer than everything else: eight SMs share them).
// Also lists the HOT slices (65536 items or more: the ones the packed accumulate leaves to the wide one) in hot[0 .. *n_hot) and the
// VERY hot ones (FINE_VHOT items or more: a cluster of CTAs shares each of them) in vhot[0 .. *n_vhot) instead.
#ifndef FINE_VHOT
#define FINE_VHOT (1u << 19)
#endif
#define FINE_SCAN_CL 8
__global__ void __cluster_dims__(FINE_SCAN_CL, 1, 1) __launch_bounds__(1024)
k_fine_scan(const u32 *__restrict__ cnt, u32 n_fine, u32 *__restrict__ start, u32 *__restrict__ cursor, u32 *__restrict__ hot, u32 *__restrict__ n_hot,
            u32 *__restrict__ vhot, u32 *__restrict__ n_vhot)
{
    namespace cg = cooperative_groups;
    __shared__ u32 s_warp[32];
    cg::cluster_group cl = cg::this_cluster();
    const u32 r = cl.block_rank();
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, gw = r * 32 + wid;
    const u32 per = (((n_fine + 32 * FINE_SCAN_CL - 1) / (32 * FINE_SCAN_CL)) + 511) & ~511u;    // counts per warp, a multiple of 512
    const u32 lo = (u32)min((u64)n_fine, (u64)gw * per), hi = (u32)min((u64)n_fine, (u64)lo + per);
    const bool vec = ((reinterpret_cast<uintptr_t>(cnt) | reinterpret_cast<uintptr_t>(start) | reinterpret_cast<uintptr_t>(cursor)) & 15) == 0;
    u32 sum = 0;
    for (u32 f0 = lo; f0 < hi; f0 += 512) {
        const u32 f = f0 + 16 * lane;
        if (vec && f + 16 <= hi) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { const uint4 q = __ldg(reinterpret_cast<const uint4 *>(cnt + f) + k); sum += q.x + q.y + q.z + q.w; }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) if (f + k < hi) sum += __ldg(cnt + f + k);
        }
    }
    sum = warp_sum(sum);
    if (lane == 0) s_warp[wid] = sum;
    cl.sync();
    u32 run = 0;                                                 // items before my chunk: the warps before mine, in every CTA up to mine
#pragma unroll
    for (u32 q = 0; q < FINE_SCAN_CL; ++q)
        if (q < r || (q == r && lane < wid)) run += *(cl.map_shared_rank(s_warp, q) + lane);
    run = warp_sum(run);
    cl.sync();                                                   // nobody reads my totals any more
    for (u32 f0 = lo; f0 < hi; f0 += 512) {
        const u32 f = f0 + 16 * lane;
        const bool whole = vec && f + 16 <= hi;
        u32 v[16];
        if (whole) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(cnt + f) + k);
                v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = f + k < hi ? __ldg(cnt + f + k) : 0u;
        }
        u32 mine = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) mine += v[k];
        u32 x = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL, x, o); if ((int)lane >= o) x += y; }
        u32 e[16];
        e[0] = run + x - mine;
#pragma unroll
        for (int k = 1; k < 16; ++k) e[k] = e[k - 1] + v[k - 1];
        if (whole) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint4 q = make_uint4(e[4 * k], e[4 * k + 1], e[4 * k + 2], e[4 * k + 3]);
                reinterpret_cast<uint4 *>(start + f)[k] = q;
                reinterpret_cast<uint4 *>(cursor + f)[k] = q;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) if (f + k < hi) { start[f + k] = e[k]; cursor[f + k] = e[k]; }
        }
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (v[k] >= 65536u) {                                // (v is 0 beyond hi)  at most 2^16 hot slices (fewer than 2^32 items)
                if (v[k] >= FINE_VHOT && vhot) vhot[atomicAdd(n_vhot, 1u)] = f + k;
                else hot[atomicAdd(n_hot, 1u)] = f + k;
            }
        run += __shfl_sync(FULL, x, 31);
    }
    if (gw == 32 * FINE_SCAN_CL - 1 && lane == 31) start[n_fine] = run;   // the last warp's carry ends up as the total (its chunk may be empty)
}

// (A branch-free form of this kernel - every item ranked through the tables with an overflow slot, as k_fine_count and k_split
// do - executes 28 % fewer instructions and is 7 % SLOWER, ncu r2m: the kernel is bound by shared-memory wavefronts (atomics,
// scatter into the staging tile, bucket look-ups: l1tex 73-77 %), and the bursts of sixteen atomics per thread fill the MIO queue.)
// NT threads x ITEMS items per tile (256 x 16, or 512 x 16: longer runs per fine slice and half as many tiles)
template <int NT, int ITEMS>
#ifndef FINE_SPLIT_OCC
#define FINE_SPLIT_OCC 4          // CTAs of 256 threads per SM the register allocation of k_fine_split is bounded for
#endif
__global__ void __launch_bounds__(NT, NT >= 512 ? 2 : FINE_SPLIT_OCC)
k_fine_split(const u32 *__restrict__ items, const u32 *__restrict__ n_ptr, u32 n_given, u32 cshift, u32 *__restrict__ cursor,
             u32 *__restrict__ out)
{
    constexpr int TILE = NT * ITEMS, NW = NT / 32;
    constexpr int SCAN_T = NT < FINE_REL ? NT : FINE_REL;      // threads that take part in the scan of the bucket counts
    constexpr int HALVES = FINE_REL / SCAN_T;
    __shared__ u32 s_cnt[FINE_REL];            // items of each fine slice in this tile, then the slice's tile-local start
    __shared__ u32 s_delta[FINE_REL];          // global start - tile-local start (mod 2^32)
    __shared__ u32 s_item[TILE];
    __shared__ u32 s_warp_tot[NW];
    const u32 n = n_ptr ? *n_ptr : n_given;
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const u64 n_tiles = ((u64)n + TILE - 1) / TILE;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (u32 b = tid; b < FINE_REL; b += NT) s_cnt[b] = 0;
        __syncthreads();
        const u64 t0 = tile * TILE;
        const u32 base = fine_base(__ldg(items + t0), cshift);
        u32 item[ITEMS], where[ITEMS];                       // where: bucket << 16 | rank inside (tile, bucket)
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const u64 i = t0 + (u64)k * NT + tid;
            item[k] = i < n ? __ldcs(items + i) : ITEM_SKIP;
        }
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            where[k] = 0xFFFFFFFFu;
            if (item[k] != ITEM_SKIP) {
                const u32 f = (item[k] & 0x7FFFFFFFu) >> FINE_SHIFT;
                if (f - base < FINE_REL) where[k] = ((f - base) << 16) | atomicAdd(&s_cnt[f - base], 1u);
                else out[atomicAdd(cursor + f, 1u)] = item[k];     // a tile that spans more than two coarse slices: one by one
            }
        }
        __syncthreads();
        // exclusive scan of the bucket counts over the tile (HALVES buckets per scanning thread, halves in order)
        u32 tot[HALVES], excl[HALVES], run = 0;
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
            const u32 b = tid + h * SCAN_T;
            tot[h] = tid < SCAN_T ? s_cnt[b] : 0u;
            u32 x = tot[h];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL, x, o); if ((int)lane >= o) x += y; }
            if (lane == 31) s_warp_tot[wid] = x;
            __syncthreads();
            u32 wbase = 0, all = 0;
#pragma unroll
            for (int k = 0; k < SCAN_T / 32; ++k) { const u32 t = s_warp_tot[k]; if (k < (int)wid) wbase += t; all += t; }
            excl[h] = run + wbase + x - tot[h];
            run += all;
            __syncthreads();
        }
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
            const u32 b = tid + h * SCAN_T;
            if (tid < SCAN_T) {
                s_cnt[b] = excl[h];
                if (tot[h]) s_delta[b] = atomicAdd(cursor + base + b, tot[h]) - excl[h];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ITEMS; ++k)
            if (where[k] != 0xFFFFFFFFu) {
                const u32 bk = where[k] >> 16;
                const u32 pos = s_cnt[bk] + (where[k] & 0xFFFFu);
                s_item[pos] = item[k];
            }
        __syncthreads();
        const u32 total = run;                                 // items of this tile that went through the tables
        for (u32 j = tid; j < total; j += NT) {
            const u32 v = s_item[j];
            out[j + s_delta[((v & 0x7FFFFFFFu) >> FINE_SHIFT) - base]] = v;
        }
        __syncthreads();
    }
}

// Persistent CTAs take fine slices from a ticket counter: bins in shared memory, items applied with shared atomics,
// statistics of the references the slice touches (a warp owns consecutive 64-bin steps; a step never straddles two
// references because the bin offsets are padded to 64), and - when the bins are kept - one coalesced write of the slice.
//   PACKED  {cov:16 | uniq_cov:16} in ONE word per bin and one atomic per item.  Exact whenever the slice holds fewer than
//           65536 items (no counter can reach 2^16) - every slice but the hottest few; 64 KB per CTA, two CTAs per SM, so
//           one CTA's fill / scan overlaps the other's item loads.  Launched for the slices with fewer than 65536 items;
//   wide    two u32 per bin (128 KB, one CTA per SM) takes the others (or all of them, SLIMM_GPU_FINE=wide).
template <bool PACKED, int NT, bool COMPACT /* PACKED only: hist4 is the array of {cov:16 | uniq_cov:16} words (half the bytes) */>
__global__ void __launch_bounds__(NT, PACKED ? FINE_PACKED_CTAS : 1)
k_fine_accumulate(const u32 *__restrict__ fine, const u32 *__restrict__ start, u32 f_lo, u32 f_hi, u64 Bp, const u64 *__restrict__ off, u32 G,
                  const u32 *__restrict__ fine_ref /* the reference that holds the first bin of every block of FINE_REF_BLOCK bins */,
                  u32 *__restrict__ stats, uint4 *__restrict__ hist4 /* nullptr: bins are not kept */, u32 *__restrict__ ticket,
                  u32 min_cnt, u32 max_cnt /* this launch takes the slices with min_cnt <= items < max_cnt */,
                  const u32 *__restrict__ hot, const u32 *__restrict__ n_hot /* not null: walk this list of slices instead of all of them */)
{
    static_assert(PACKED || !COMPACT, "compact bins are the packed counters as they are");
    extern __shared__ u32 sh[];                                // PACKED: bins[FINE_BINS]; wide: cov[FINE_BINS] | uniq_cov[FINE_BINS]
    __shared__ u32 s_next;
    constexpr u32 WORDS = PACKED ? FINE_BINS : 2 * FINE_BINS;
    constexpr int SPW = 256 / (NT / 32);                       // 64-bin steps per warp
    constexpr int PRE = PACKED ? 16 : FINE_PRE;                // items per thread requested up front
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const u64 n_steps = Bp >> 6;
    const u32 n_list = hot ? *n_hot : 0u;
    u32 idx = blockIdx.x;                                      // list mode: my position in the list
    for (u32 f = hot ? (idx < n_list ? hot[idx] : f_hi) : f_lo + blockIdx.x; f < f_hi;) {
        if (tid == 0) s_next = (hot ? 0u : f_lo) + gridDim.x + atomicAdd(ticket, 1u);   // the next slice's ticket travels while this one is worked on (list mode: the next list position)
        const u32 lo = __ldg(start + f), hi = __ldg(start + f + 1);
        const u32 cnt = hi - lo;
        const bool mine = cnt >= min_cnt && cnt < max_cnt;
        if (mine && (cnt || hist4)) {                          // an empty slice only matters when somebody reads the bins
            // the first items of every thread are requested first of all: their DRAM latency hides behind the reference lookup
            // and the zero-fill of the bins
            const u32 *my_items = fine + lo + tid;
            u32 pre[PRE];
#pragma unroll
            for (int k = 0; k < PRE; ++k) pre[k] = (u32)k * NT + tid < cnt ? __ldcs(my_items + k * NT) : ITEM_SKIP;
            const u64 bin0 = (u64)f << FINE_SHIFT;
            // my SPW steps; the first reference is looked up now, among the few references the slice touches
            const u64 step0 = (bin0 >> 6) + wid * SPW;
            const u32 my_steps = step0 >= n_steps ? 0u : (u32)min((u64)SPW, n_steps - step0);
            u32 g = 0;
            u64 g_end = 0;
            if (cnt && my_steps) {
                const u64 first_bin = step0 << 6;
                u32 a = __ldg(fine_ref + first_bin / FINE_REF_BLOCK), b = min(__ldg(fine_ref + first_bin / FINE_REF_BLOCK + 1) + 1u, G);   // largest g in [a, b) with off[g] <= my first bin
                while (b - a > 1) { const u32 mid = (a + b) >> 1; if (__ldg(off + mid) <= first_bin) a = mid; else b = mid; }
                g = a;
                g_end = __ldg(off + g + 1);
            }
            for (u32 k = tid; k < WORDS / 4; k += NT) reinterpret_cast<uint4 *>(sh)[k] = make_uint4(0, 0, 0, 0);
            __syncthreads();
            // the items beyond the first PRE per thread arrive TB per thread at a time; in the wide kernel a batch is requested before
            // the one before it is applied: a hot slice streams its items with 2 x TB loads per thread in flight instead of waiting for
            // DRAM once per batch (it was bound by exactly that: 4 loads per thread and round trip = 1.6 TB/s over 148 SMs, launch list r2o)
            constexpr int TB = PACKED ? 4 : 12;
            auto apply = [&](u32 it) {
                if (it != ITEM_SKIP) {
                    const u32 b = it & (FINE_BINS - 1);
                    if (PACKED) atomicAdd(&sh[b], (it >> 31) ? 0x10001u : 1u);
                    else { atomicAdd(&sh[b], 1u); if (it >> 31) atomicAdd(&sh[FINE_BINS + b], 1u); }
                }
            };
            if (PACKED) {                                          // (no registers to spare at three CTAs per SM: one batch at a time)
#pragma unroll
                for (int k = 0; k < PRE; ++k) apply(pre[k]);
                for (u32 i0 = PRE * NT; i0 < cnt; i0 += TB * NT) { // a slice with more items than usual
                    u32 v[TB];
#pragma unroll
                    for (int k = 0; k < TB; ++k) v[k] = i0 + k * NT + tid < cnt ? __ldcs(my_items + i0 + k * NT) : ITEM_SKIP;
#pragma unroll
                    for (int k = 0; k < TB; ++k) apply(v[k]);
                }
            } else {
                u32 v[TB];
#pragma unroll
                for (int k = 0; k < TB; ++k) v[k] = (u32)(PRE + k) * NT + tid < cnt ? __ldcs(my_items + (PRE + k) * NT) : ITEM_SKIP;
#pragma unroll
                for (int k = 0; k < PRE; ++k) apply(pre[k]);
                for (u32 i0 = PRE * NT; i0 < cnt; i0 += TB * NT) {
                    u32 nx[TB];
#pragma unroll
                    for (int k = 0; k < TB; ++k) nx[k] = i0 + (TB + k) * NT + tid < cnt ? __ldcs(my_items + i0 + (TB + k) * NT) : ITEM_SKIP;
#pragma unroll
                    for (int k = 0; k < TB; ++k) apply(v[k]);
#pragma unroll
                    for (int k = 0; k < TB; ++k) v[k] = nx[k];
                }
            }
            __syncthreads();
            if (my_steps) {
                // bins [wid * SPW * 64, (wid + 1) * SPW * 64) of the slice, eight steps at a time; lane l holds bins 2l, 2l+1 of a step
                u32 nz = 0, sum = 0, unz = 0, usum = 0;
                u32 nzp = 0, sump = 0;                         // PACKED: {cov | uniq_cov << 16} partial counts / sums of the fast path below
#pragma unroll 1
                for (int h = 0; h < SPW / 8; ++h) {
                    const u32 wb = (wid * SPW + h * 8) * 64 + 2 * lane;
                    if (PACKED && cnt && (u32)(h * 8 + 8) <= my_steps && ((step0 + h * 8 + 8) << 6) <= g_end) {
                        // eight whole steps inside the current reference (almost always): the statistics are taken from the packed
                        // words as they are - a slice with fewer than 65536 items cannot carry from one half into the other, so ONE
                        // add per word sums cov and uniq_cov together - and the words go out without being unpacked (the packed
                        // pass was bound by instruction issue: 31 thread instructions per bin, ncu r2z)
                        uint2 w[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) w[t] = *reinterpret_cast<const uint2 *>(sh + wb + t * 64);
                        if (hist4) {
#pragma unroll
                            for (int t = 0; t < 8; ++t) {
                                if (COMPACT) __stcs(reinterpret_cast<uint2 *>(hist4) + (step0 + h * 8 + t) * 32 + lane, w[t]);
                                else __stcs(hist4 + (step0 + h * 8 + t) * 32 + lane, make_uint4(w[t].x & 0xFFFFu, w[t].x >> 16, w[t].y & 0xFFFFu, w[t].y >> 16));
                            }
                        }
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            sump += w[t].x + w[t].y;
                            nzp += min(w[t].x & 0xFFFFu, 1u) + min(w[t].y & 0xFFFFu, 1u) + ((min(w[t].x >> 16, 1u) + min(w[t].y >> 16, 1u)) << 16);
                        }
                        continue;
                    }
                    if (PACKED) { nz += nzp & 0xFFFFu; unz += nzp >> 16; sum += sump & 0xFFFFu; usum += sump >> 16; nzp = sump = 0; }
                    uint4 v[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        if (PACKED) {
                            const uint2 w2 = *reinterpret_cast<const uint2 *>(sh + wb + t * 64);
                            if (COMPACT && hist4 && (u32)(h * 8 + t) < my_steps) __stcs(reinterpret_cast<uint2 *>(hist4) + (step0 + h * 8 + t) * 32 + lane, w2);
                            v[t] = make_uint4(w2.x & 0xFFFFu, w2.x >> 16, w2.y & 0xFFFFu, w2.y >> 16);
                        } else {
                            const uint2 c2 = *reinterpret_cast<const uint2 *>(sh + wb + t * 64),
                                        u2 = *reinterpret_cast<const uint2 *>(sh + FINE_BINS + wb + t * 64);
                            v[t] = make_uint4(c2.x, u2.x, c2.y, u2.y);
                        }
                    }
                    if (hist4 && !COMPACT) {
                        uint4 *dst = hist4 + (step0 + h * 8) * 32 + lane;
#pragma unroll
                        for (int t = 0; t < 8; ++t) if ((u32)(h * 8 + t) < my_steps) __stcs(dst + t * 32, v[t]);
                    }
                    if (cnt) {
                        const u64 first_bin = (step0 + h * 8) << 6;
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            if ((u32)(h * 8 + t) < my_steps) {
                                if (first_bin + (u64)t * 64 >= g_end) {    // warp-uniform: the step starts another reference
                                    nz = warp_sum(nz); sum = warp_sum(sum); unz = warp_sum(unz); usum = warp_sum(usum);
                                    if (lane == 0 && (nz | unz)) {
                                        atomicAdd(stats + 4 * g + 0, nz); atomicAdd(stats + 4 * g + 1, sum);
                                        if (unz) { atomicAdd(stats + 4 * g + 2, unz); atomicAdd(stats + 4 * g + 3, usum); }
                                    }
                                    nz = sum = unz = usum = 0;
                                    while (first_bin + (u64)t * 64 >= g_end) { ++g; g_end = __ldg(off + g + 1); }
                                }
                                nz += (v[t].x != 0) + (v[t].z != 0); sum += v[t].x + v[t].z;
                                unz += (v[t].y != 0) + (v[t].w != 0); usum += v[t].y + v[t].w;
                            }
                        }
                    }
                }
                if (cnt) {
                    if (PACKED) { nz += nzp & 0xFFFFu; unz += nzp >> 16; sum += sump & 0xFFFFu; usum += sump >> 16; }
                    nz = warp_sum(nz); sum = warp_sum(sum); unz = warp_sum(unz); usum = warp_sum(usum);
                    if (lane == 0 && (nz | unz)) {
                        atomicAdd(stats + 4 * g + 0, nz); atomicAdd(stats + 4 * g + 1, sum);
                        if (unz) { atomicAdd(stats + 4 * g + 2, unz); atomicAdd(stats + 4 * g + 3, usum); }
                    }
                }
            }
        }
        __syncthreads();                                       // everybody is done with the bins; s_next is visible
        if (hot) { idx = s_next; f = idx < n_list ? hot[idx] : f_hi; }
        else f = s_next;
        __syncthreads();
    }
}

// The VERY hot fine slices (FINE_VHOT items or more; lognormal communities put millions of items on the slices of their top
// genomes) would each keep one CTA busy long after the rest of the GPU has finished.  Here a cluster of FINE_CL CTAs shares a
// slice: every CTA applies its share of the items to a private copy of the bins (wide counters, 128 KB), the copies are summed
// through distributed shared memory - CTA r of the cluster sums bins [r, r+1) * FINE_BINS / FINE_CL of all copies into its own -
// and every CTA reduces the statistics of and writes out its own share of the bins.  The list is walked cluster by cluster.
#define FINE_CL 8
__global__ void __cluster_dims__(FINE_CL, 1, 1) __launch_bounds__(1024, 1)
k_fine_accumulate_cluster(const u32 *__restrict__ fine, const u32 *__restrict__ start, u64 Bp, const u64 *__restrict__ off, u32 G,
                          const u32 *__restrict__ fine_ref, u32 *__restrict__ stats, uint4 *__restrict__ hist4 /* nullptr: bins are not kept */,
                          const u32 *__restrict__ vhot, const u32 *__restrict__ n_vhot)
{
    namespace cg = cooperative_groups;
    extern __shared__ u32 sh[];                                // cov[FINE_BINS] | uniq_cov[FINE_BINS]
    constexpr u32 NT = 1024, SHARE = FINE_BINS / FINE_CL;      // bins a CTA sums, reduces and writes
    constexpr u32 TAKE = NT * 8;                               // items a CTA takes at a time
    static_assert(SHARE == 64 * (NT / 32), "one 64-bin step per warp");
    cg::cluster_group cl = cg::this_cluster();
    const u32 r = cl.block_rank(), cid = blockIdx.x / FINE_CL, n_cl = gridDim.x / FINE_CL;
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const u64 n_steps = Bp >> 6;
    const u32 n_list = *n_vhot;
    for (u32 idx = cid; idx < n_list; idx += n_cl) {
        const u32 f = vhot[idx];
        const u32 lo = __ldg(start + f), cnt = __ldg(start + f + 1) - lo;
        for (u32 k = tid; k < 2 * FINE_BINS / 4; k += NT) reinterpret_cast<uint4 *>(sh)[k] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        // my share of the items, eight per thread at a time; a batch is requested before the one before it is applied
        u32 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = r * TAKE + k * NT + tid < cnt ? __ldcs(fine + lo + r * TAKE + k * NT + tid) : ITEM_SKIP;
        for (u32 i0 = r * TAKE; i0 < cnt; i0 += FINE_CL * TAKE) {
            u32 nx[8];
            const u32 i1 = i0 + FINE_CL * TAKE;
#pragma unroll
            for (int k = 0; k < 8; ++k) nx[k] = i1 + k * NT + tid < cnt ? __ldcs(fine + lo + i1 + k * NT + tid) : ITEM_SKIP;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (v[k] != ITEM_SKIP) {
                    const u32 b = v[k] & (FINE_BINS - 1);
                    atomicAdd(&sh[b], 1u);
                    if (v[k] >> 31) atomicAdd(&sh[FINE_BINS + b], 1u);
                }
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = nx[k];
        }
        cl.sync();                                             // every copy is complete
#pragma unroll
        for (int h = 0; h < 2; ++h) {                          // cov, uniq_cov
#pragma unroll
            for (u32 w0 = 0; w0 < SHARE; w0 += NT) {
                const u32 w = h * FINE_BINS + r * SHARE + w0 + tid;
                u32 a = sh[w];
#pragma unroll
                for (u32 q = 1; q < FINE_CL; ++q) a += *(cl.map_shared_rank(sh, (r + q) % FINE_CL) + w);
                sh[w] = a;
            }
        }
        cl.sync();                                             // nobody reads my copy any more; my share is summed
        const u64 step = (((u64)f << FINE_SHIFT) >> 6) + r * (SHARE / 64) + wid;      // my warp's 64 bins
        if (step < n_steps) {
            const u64 first_bin = step << 6;
            u32 a = __ldg(fine_ref + first_bin / FINE_REF_BLOCK), b = min(__ldg(fine_ref + first_bin / FINE_REF_BLOCK + 1) + 1u, G);    // largest g in [a, b) with off[g] <= my first bin
            while (b - a > 1) { const u32 mid = (a + b) >> 1; if (__ldg(off + mid) <= first_bin) a = mid; else b = mid; }
            const u32 wb = r * SHARE + wid * 64 + 2 * lane;    // lane l holds bins 2l, 2l+1 of the step
            const uint2 c2 = *reinterpret_cast<const uint2 *>(sh + wb), u2 = *reinterpret_cast<const uint2 *>(sh + FINE_BINS + wb);
            if (hist4) __stcs(hist4 + step * 32 + lane, make_uint4(c2.x, u2.x, c2.y, u2.y));
            const u32 nz = warp_sum((c2.x != 0) + (c2.y != 0)), sum = warp_sum(c2.x + c2.y);
            const u32 unz = warp_sum((u2.x != 0) + (u2.y != 0)), usum = warp_sum(u2.x + u2.y);
            if (lane == 0 && (nz | unz)) {
                atomicAdd(stats + 4 * a + 0, nz); atomicAdd(stats + 4 * a + 1, sum);
                if (unz) { atomicAdd(stats + 4 * a + 2, unz); atomicAdd(stats + 4 * a + 3, usum); }
            }
        }
        __syncthreads();                                       // my bins are read before the next slice zero-fills them
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
k_ref_stats(const uint4 *__restrict__ hist4, u64 step_lo, u64 n_steps /* end of the step range */, const u64 *__restrict__ off /*[G+1] padded, in bins*/,
            u32 G, u32 *__restrict__ stats)
{
    const u32 lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 chunk = warp; step_lo + chunk * STATS_STEPS_PER_WARP < n_steps; chunk += n_warps) {
        const u64 s0 = step_lo + chunk * STATS_STEPS_PER_WARP;
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
// compact storage of a fine-slice run: which layout holds this bin (see k_extract_bins_compact)
__device__ __forceinline__ bool fine_slice_is_wide(const u32 *__restrict__ fine_start, u64 bin)
{
    const u32 f = (u32)(bin >> FINE_SHIFT);
    return __ldg(fine_start + f + 1) - __ldg(fine_start + f) >= 65536u;
}
__global__ void __launch_bounds__(256)
k_cov2_base(const uint4 *__restrict__ hist4, u64 n_steps, const u64 *__restrict__ off, u32 G,
            const u32 *__restrict__ valid_bits, uint2 *__restrict__ cov2_2,
            const u32 *__restrict__ hist16 /* not null: compact storage (see k_extract_bins_compact) */, const u32 *__restrict__ fine_start)
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
            if (hist16 && !fine_slice_is_wide(fine_start, bin0)) {
                const uint2 w2 = __ldg(reinterpret_cast<const uint2 *>(hist16 + bin0) + lane);
                o = make_uint2(w2.x >> 16, w2.y >> 16);
            } else {
                const uint4 v = __ldg(hist4 + s * 32 + lane);
                o = make_uint2(v.y, v.w);
            }
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

// The descending walk of get_quantile_cut_off over the staged values buf[k] = v[c_lo + k]:
//   while (sub < sstar && i > 0) { sub += v[i]; --i; }   restricted to indices >= c_lo.
// Eight values are loaded up front and their running sums formed back to back, so the only serial chain is
// the f32 additions themselves (the sums are the same left-to-right sums, rounded once per addition).
__device__ __forceinline__ void walk_chunk(const float *buf, u32 c_lo, u32 &i, float &sub, float sstar)
{
    const u32 stop = max(c_lo, 1u);                                // index 0 is never added
    while (i >= stop) {
        const u32 m = min(i - stop + 1u, 8u);
        float pre[9];
        pre[0] = sub;
#pragma unroll
        for (int k = 0; k < 8; ++k) pre[k + 1] = __fadd_rn(pre[k], k < (int)m ? buf[i - k - c_lo] : 0.0f);
        u32 take = m;                                              // first k with !(pre[k] < sstar): v[i - k] is not added
#pragma unroll
        for (int k = 7; k >= 0; --k) if (k < (int)m && !(pre[k] < sstar)) take = k;
        sub = pre[take];
        i -= take;
        if (take < m) return;
    }
}

// left fold total + buf[0] + buf[1] + ... (one rounding per addition, in order): the values are loaded sixteen at a time so
// that the only serial chain is the additions themselves
__device__ __forceinline__ float fold_in_order(float total, const float *buf, u32 n)
{
    // the next sixteen values are loaded while the current sixteen are added: the additions (4 cycles each, one after the other)
    // are the only thing the thread ever waits for
    float x[16], y[16];
    u32 k = 0;
    if (n >= 16) {
#pragma unroll
        for (int t = 0; t < 16; ++t) x[t] = buf[t];
        for (; k + 32 <= n; k += 16) {
#pragma unroll
            for (int t = 0; t < 16; ++t) y[t] = buf[k + 16 + t];
#pragma unroll
            for (int t = 0; t < 16; ++t) total = __fadd_rn(total, x[t]);
#pragma unroll
            for (int t = 0; t < 16; ++t) x[t] = y[t];
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) total = __fadd_rn(total, x[t]);
        k += 16;
    }
    for (; k < n; ++k) total = __fadd_rn(total, buf[k]);
    return total;
}

// running sums in place: buf[k] = carry + buf[0] + ... + buf[k] (one rounding per addition, in order); returns the last one
__device__ __forceinline__ float running_sums_in_order(float run, float *buf, u32 n)
{
    float x[16], y[16];
    u32 k = 0;
    if (n >= 16) {
#pragma unroll
        for (int t = 0; t < 16; ++t) x[t] = buf[t];
        for (; k + 32 <= n; k += 16) {
#pragma unroll
            for (int t = 0; t < 16; ++t) y[t] = buf[k + 16 + t];
#pragma unroll
            for (int t = 0; t < 16; ++t) { run = __fadd_rn(run, x[t]); buf[k + t] = run; }
#pragma unroll
            for (int t = 0; t < 16; ++t) x[t] = y[t];
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) { run = __fadd_rn(run, x[t]); buf[k + t] = run; }
        k += 16;
    }
    for (; k < n; ++k) { run = __fadd_rn(run, buf[k]); buf[k] = run; }
    return run;
}

// valid set + -v counters (src/slimm.hpp:354-378); one CTA of 1024 threads, after both cut-offs are known
__device__ __forceinline__ void cutoffs_valid_set(const u32 *__restrict__ stats, u32 G, u32 min_reads, const float *__restrict__ cp_all,
                                                  u32 *__restrict__ valid_bits, unsigned char *__restrict__ valid_bytes, DevScalars *sc)
{
    const u32 tid = threadIdx.x;
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
                s_f = fold_in_order(s_f, s_buf, cn);
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
                walk_chunk(s_buf, c_lo, i, sub, sstar);
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
    cutoffs_valid_set(stats, G, min_reads, cp_all, valid_bits, valid_bytes, sc);
}

// ------------------------------------------------------------------------------------------------
// K4 on a thread-block cluster.  Same contract as k_cutoffs, for up to CUT_CL * CUT_SHARE references: the
// ascending sort runs as a bitonic network over a key array DISTRIBUTED over the shared memories of the 8
// CTAs of a cluster - compare-exchange partners further apart than one CTA's share are reached through
// DSMEM - so no step touches L2; the two order-sensitive f32 folds stay sequential in one thread, fed from
// shared-memory chunks staged by the whole CTA.  grid = 2 clusters (cov, uniq_cov) x 8 CTAs x 1024 threads.
// ------------------------------------------------------------------------------------------------
#define CUT_CL 8
#define CUT_SHARE 32768        // keys per CTA (128 KB of dynamic shared memory)

namespace cgx = cooperative_groups;

__device__ __forceinline__ u32 *dsm_key(cgx::cluster_group &cl, u32 *keys, u32 idx, u32 share_log)
{
    return cl.map_shared_rank(keys, idx >> share_log) + (idx & ((1u << share_log) - 1u));
}

// K4 in three kernels, so that the two order-sensitive f32 chains (each one dependent addition per reference: ~100 us at
// 50 000 references, whatever the hardware) run CONCURRENTLY instead of back to back behind the sort:
//   k_cut_fold          (side stream)  total = std::accumulate(v, 0.0f) in ascending reference order
//   k_cut_sort_cluster  (main stream)  members -> distributed shared memory, bitonic sort, sorted keys to global memory, and the
//                                      descending running sums P_k = v[n-1] + v[n-2] + ... (k terms) - the `sub` of
//                                      get_quantile_cut_off after k trips, which does not depend on the total
//   k_cut_finish        (after both)   first k with !(P_k / total < q) by a PARALLEL search (same comparison, same operands,
//                                      monotone in k), cut-off = v[n-1-k]; then the valid set and the -v counters
#define FOLD_CHUNK 4096
#define PREFIX_CHUNK 2048
__global__ void __launch_bounds__(1024)
k_cut_fold(const u32 *__restrict__ stats, const uint4 *__restrict__ meta, u32 G, DevScalars *sc)
{
    __shared__ float s_buf[2][FOLD_CHUNK];
    const u32 which = blockIdx.x, tid = threadIdx.x;
    float total = 0.0f;                                            // thread 0's copy is the fold
    // references without unique reads add +0.0f: exact, and the chain stays free of branches
    auto stage = [&](u32 c0, float *buf, u32 first, u32 step) {
        for (u32 k = first; k < FOLD_CHUNK; k += step) {
            const u32 g = c0 + k;
            float x = 0.0f;
            if (g < G && stats[4 * g + 3] > 0) x = __fdiv_rn((float)stats[4 * g + 2 * which], (float)__ldg(&meta[g].y));
            buf[k] = x;
        }
    };
    stage(0, s_buf[0], tid, 1024);
    __syncthreads();
    u32 cur = 0;
    for (u32 c0 = 0; c0 < G; c0 += FOLD_CHUNK) {                   // the other warps stage the next chunk while thread 0 folds this one
        if (tid >= 32 && c0 + FOLD_CHUNK < G) stage(c0 + FOLD_CHUNK, s_buf[cur ^ 1], tid - 32, 992);
        if (tid == 0) total = fold_in_order(total, s_buf[cur], min((u32)FOLD_CHUNK, G - c0));
        __syncthreads();
        cur ^= 1;
    }
    if (tid == 0) { sc->tot[which] = total; __threadfence(); *(volatile u32 *)&sc->tot_ready[which] = 1u; }
}

__global__ void __cluster_dims__(CUT_CL, 1, 1) __launch_bounds__(1024)
k_cut_sort_cluster(const u32 *__restrict__ stats, const uint4 *__restrict__ meta, u32 G, float q, float *__restrict__ cp_all /*[2][G]*/,
                   u32 *__restrict__ sorted /*[2][G] ascending*/, float *__restrict__ prefix /*[2][G]: P_0 = 0, P_k*/, DevScalars *sc)
{
    extern __shared__ u32 s_keys[];           // this CTA's share of the distributed key array
    cgx::cluster_group cl = cgx::this_cluster();
    const u32 rank = cl.block_rank();
    const u32 which = blockIdx.x / CUT_CL;    // 0: cov, 1: uniq_cov
    const u32 tid = threadIdx.x, lane = tid & 31;
    float *cp = cp_all + (size_t)which * G;
    __shared__ u32 s_n;                       // rank 0's copy counts the members of the whole cluster
    __shared__ float s_p[PREFIX_CHUNK];
    __shared__ float s_carry;
    __shared__ int s_stop;
    // padded size: a power of two that holds every reference (how many have unique reads is only known after the pass below)
    u32 m = CUT_CL * 1024;
    while (m < G) m <<= 1;
    const u32 share_log = 31 - __clz(m / CUT_CL), share = 1u << share_log;
    for (u32 k = tid; k < share; k += 1024) s_keys[k] = 0xFFFFFFFFu;
    if (tid == 0) s_n = 0;
    cl.sync();                                // every share initialised, the counter zeroed
    // my range of references: cov_percent = float(nz) / number_of_bins (src/reference_contig.hpp:148-155) for everybody; the members'
    // keys go to the distributed array, a warp's batch at the place a cluster-wide counter hands out (their order does not matter:
    // they are sorted next)
    {
        u32 *const counter = cl.map_shared_rank(&s_n, 0);
        const u32 per = (G + CUT_CL - 1) / CUT_CL, g_end = min(G, (rank + 1) * per);
        for (u32 g0 = rank * per; g0 < g_end; g0 += 1024) {
            const u32 g = g0 + tid;
            float x = 0.0f;
            bool keep = false;
            if (g < g_end) {
                x = __fdiv_rn((float)stats[4 * g + 2 * which], (float)meta[g].y);
                cp[g] = x;
                keep = stats[4 * g + 3] > 0;
            }
            const u32 bal = __ballot_sync(FULL, keep);
            u32 base = 0;
            if (lane == 0 && bal) base = atomicAdd(counter, (u32)__popc(bal));
            base = __shfl_sync(FULL, base, 0);
            if (keep) *dsm_key(cl, s_keys, base + __popc(bal & LANE_LT(lane)), share_log) = __float_as_uint(x);
        }
    }
    cl.sync();
    const u32 n = *cl.map_shared_rank(&s_n, 0);
    cl.sync();
    // bitonic sort, ascending (values are >= 0: u32 order == f32 order)
    for (u32 k = 2; k <= m; k <<= 1)
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            if (j >= share) {                 // partner lives in another CTA: the lower rank of the pair works
                const u32 prank = rank ^ (j >> share_log);
                if (prank > rank) {
                    u32 *other = cl.map_shared_rank(s_keys, prank);
                    for (u32 o = tid; o < share; o += 1024) {
                        const u32 t = (rank << share_log) | o;
                        const u32 a = s_keys[o], b = other[o];
                        const bool up = (t & k) == 0;
                        if ((a > b) == up) { s_keys[o] = b; other[o] = a; }
                    }
                }
                cl.sync();
            } else {
                for (u32 o = tid; o < share; o += 1024) {
                    const u32 po = o ^ j;
                    if (po > o) {
                        const u32 t = (rank << share_log) | o;
                        const u32 a = s_keys[o], b = s_keys[po];
                        const bool up = (t & k) == 0;
                        if ((a > b) == up) { s_keys[o] = b; s_keys[po] = a; }
                    }
                }
                __syncthreads();
            }
        }
    cl.sync();
    // my share of the sorted keys -> global memory
    for (u32 o = tid; o < share; o += 1024) {
        const u32 t = (rank << share_log) | o;
        if (t < n) sorted[(size_t)which * G + t] = s_keys[o];
    }
    if (rank == 0) {
        // P_0 = 0, P_k = P_{k-1} + v[n-k]: one thread, the values staged PREFIX_CHUNK at a time by the whole CTA.  Once k_cut_fold
        // (running beside this kernel) has published the total, the chain ends with the first chunk whose last sum stops the loop
        // of get_quantile_cut_off: k_cut_finish will not look further
        if (tid == 0) { s_carry = 0.0f; s_stop = 0; sc->n_members[which] = n; }
        float *P = prefix + (size_t)which * G;
        u32 k_done = 0;
        for (u32 k0 = 0; k0 < n; k0 += PREFIX_CHUNK) {             // P_{k0} .. P_{k0 + cn - 1}
            const u32 cn = min((u32)PREFIX_CHUNK, n - k0);
            // s_p[k] = v[n - (k0 + k)] for k0 + k >= 1 (the term that makes P_{k0+k} out of P_{k0+k-1}); P_0 = 0
            for (u32 k = tid; k < cn; k += 1024) {
                const u32 kk = k0 + k;
                s_p[k] = kk == 0 ? 0.0f : __uint_as_float(*dsm_key(cl, s_keys, n - kk, share_log));
            }
            __syncthreads();
            if (tid == 0) {
                const float run = running_sums_in_order(s_carry, s_p, cn);
                s_carry = run;
                if (*(volatile u32 *)&sc->tot_ready[which]) {
                    __threadfence();
                    if (!(__fdiv_rn(run, *(volatile float *)&sc->tot[which]) < q)) s_stop = 1;
                }
            }
            __syncthreads();
            for (u32 k = tid; k < cn; k += 1024) P[k0 + k] = s_p[k];
            k_done = k0 + cn;
            if (s_stop) break;
            __syncthreads();
        }
        if (tid == 0) sc->k_done[which] = k_done;
    }
    cl.sync();                                // nobody leaves while rank 0 may still read its share
}

// Both cut-offs by a parallel search over the running sums (every CTA for itself: two rounds of 1024 probes, P_k / total < q is
// monotone in k), then the valid set and the -v counters of the CTA's 1024 references
__global__ void __launch_bounds__(1024)
k_cut_finish(const u32 *__restrict__ stats, u32 G, float q, u32 min_reads, const float *__restrict__ cp_all, const u32 *__restrict__ sorted,
             const float *__restrict__ prefix, u32 *__restrict__ valid_bits, unsigned char *__restrict__ valid_bytes, DevScalars *sc)
{
    __shared__ u32 s_k[2][2];
    __shared__ float s_cut[2];
    const u32 tid = threadIdx.x;
    if (min_reads == 0) {                     // -mr default: 1 + (matches_count-1)/10000 (src/slimm.hpp:458-459)
        const u32 R = (u32)sc->n_reads;
        min_reads = R ? 1u + (R - 1u) / 10000u : 0u;
    }
    const u32 n = sc->n_members[0];
    const bool active = q < 1.0f && n > 0;
    // i = n-1; while ((sub/total) < q && i > 0) { sub += v[i]; --i; }: after k trips sub == P_k; the loop also ends after n - 1 trips
    if (tid < 4) s_k[tid >> 1][tid & 1] = n ? n - 1 : 0;
    __syncthreads();
    if (active) {
        const u32 stride = (n + 1023) / 1024;
#pragma unroll
        for (int which = 0; which < 2; ++which) {                  // round 1: every stride-th k (the running sums end at k_done: a stop lies before it, or k_done == n)
            const u32 k = tid * stride;
            if (k < sc->k_done[which] && !(__fdiv_rn(prefix[(size_t)which * G + k], sc->tot[which]) < q)) atomicMin(&s_k[which][0], k);
        }
        __syncthreads();
#pragma unroll
        for (int which = 0; which < 2; ++which) {                  // round 2: the stride before the first probe that stopped
            const u32 hi = s_k[which][0];                          // the first probe that stopped, or n - 1: the first stop lies in (hi - stride, hi]
            const u32 lo = hi >= stride ? hi - stride : 0u;
            for (u32 k = lo + tid; k <= hi && k < sc->k_done[which]; k += 1024)
                if (!(__fdiv_rn(prefix[(size_t)which * G + k], sc->tot[which]) < q)) { atomicMin(&s_k[which][1], k); break; }
        }
    }
    __syncthreads();
    if (tid < 2) s_cut[tid] = active ? __uint_as_float(sorted[(size_t)tid * G + n - 1 - s_k[tid][1]]) : 0.0f;
    __syncthreads();
    const float c0 = s_cut[0], c1 = s_cut[1];
    if (blockIdx.x == 0 && tid == 0) { sc->cut = c0; sc->ucut = c1; }
    // valid set + -v counters (src/slimm.hpp:354-378) of my 1024 references
    const float *cpa = cp_all, *ucpa = cp_all + G;
    const u32 g = blockIdx.x * 1024 + tid;
    u32 nv = 0, fc = 0, fu = 0, fm = 0, rc = 0;
    unsigned long long pairs = 0;
    bool ok = false;
    if (g < G) {
        const u32 reads = stats[4 * g + 1];
        if (reads > 0) {
            rc = 1; pairs = reads;
            const float a = cpa[g], b = ucpa[g];
            ok = a >= c0 && b >= c1;
            if (ok) nv = 1;
            else { fu = b < c1; fm = reads < min_reads; fc = a < c0; }
        }
        valid_bytes[g] = ok;
    }
    const u32 word = __ballot_sync(FULL, ok);
    if ((tid & 31) == 0 && g < G) valid_bits[g >> 5] = word;
    nv = warp_sum(nv); fc = warp_sum(fc); fu = warp_sum(fu); fm = warp_sum(fm); rc = warp_sum(rc);
    pairs = warp_sum64(pairs);
    if ((tid & 31) == 0 && (rc | nv)) {
        atomicAdd(&sc->n_valid, nv); atomicAdd(&sc->failed_cov, fc); atomicAdd(&sc->failed_ucov, fu);
        atomicAdd(&sc->failed_minread, fm); atomicAdd(&sc->ref_count, rc);
        atomicAdd(&sc->n_pairs, pairs);
    }
}

// ------------------------------------------------------------------------------------------------
// K5+K6: reassignment + LCA on the compact stream k_coverage left behind: the distinct reference ids of
// every read with several targets, chunk by chunk (reads with one target are settled without touching
// the records again: k_finish_assign adds valid[g] * uniq_reads_count[g]).  The same sliding windows as
// in k_coverage, now with explicit head bits; a read's surviving set S = targets /\ valid is resolved with
// ballots and REDUX restricted to the read's lanes:
//   |S| = 1 -> the read BECAME unique through the filter (src/slimm.hpp:383-390)
//   |S| >= 2 -> level-wise LCA over the 8-slot lineages, zeros included (slimm::get_lca, :516-531): the
//              first level on which all of S agree, else slot 7 of the largest reference id;
//              count[lca] += 1 and children[lca] U= S (phase 1 of get_reads_lca_count, :536-557),
//              both keyed by (reference, level) instead of taxon id.
// Replaces the read loop of filter_alignments (src/slimm.hpp:380-391, read_stat::update
// src/read_stat.hpp:98-114), slimm::get_lca and phase 1 of get_reads_lca_count.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_valid(const u32 *__restrict__ vb, u32 g) { return (__ldg(vb + (g >> 5)) >> (g & 31)) & 1u; }

// bit l set: the lineages of g and g0 differ on level l
__device__ __forceinline__ u32 lineage_diff(const uint4 *__restrict__ lin4, u32 g, const uint4 &a0, const uint4 &b0)
{
    const uint4 a = __ldg(lin4 + 2 * (u64)g), b = __ldg(lin4 + 2 * (u64)g + 1);
    return (u32)(a.x != a0.x) | ((u32)(a.y != a0.y) << 1) | ((u32)(a.z != a0.z) << 2) | ((u32)(a.w != a0.w) << 3) |
           ((u32)(b.x != b0.x) << 4) | ((u32)(b.y != b0.y) << 5) | ((u32)(b.z != b0.z) << 6) | ((u32)(b.w != b0.w) << 7);
}

struct AssignParams {
    const u32 *cw, *cw_idx; const uint2 *chunk_cnt; const u32 *lr;
    const uint4 *meta; const uint4 *lin4; const u32 *top_idx; const u32 *vb; u32 G, half_avg; BinDiv wdiv;
    u32 *uniq2_extra, *lca_cnt, *child_mark, *fb_mark, *cov2;
    unsigned char *res_kind; u32 *res_val;
};

__device__ __forceinline__ void mark_child(const AssignParams &P, u32 h, bool fb, u32 level, u32 owner)
{
    u32 *mk = fb ? P.fb_mark + (u64)__ldg(P.top_idx + owner) * P.G + h : P.child_mark + (u64)h * 8 + level;
    if (*mk == 0) *mk = 1;
}

// count[lca] += 1 through a small direct-mapped cache in shared memory: a taxon that dominates the sample
// costs one global atomic per CTA instead of one per read
#define LCA_CACHE 2048
#define LCA_EMPTY 0xFFFFFFFFu
__device__ __forceinline__ void count_lca(u32 *s_key, u32 *s_val, u32 *__restrict__ lca_cnt, u32 key)
{
    const u32 slot = (key * 2654435761u) >> 21;
    const u32 old = atomicCAS(&s_key[slot], LCA_EMPTY, key);
    if (old == LCA_EMPTY || old == key) atomicAdd(&s_val[slot], 1u);
    else atomicAdd(lca_cnt + key, 1u);
}

// a read of more than 32 records, walked by the whole warp on the original records
template <class Rec>
__device__ __noinline__ void assign_long_run(const Rec &rec, u32 p, u32 n, u32 lane, const AssignParams &P, u32 *s_key, u32 *s_val)
{
    const u32 r0 = rec.read(p);
    bool have = false, multi = false;
    u32 g0 = 0, neq = 0, gmax = 0, lead = p, end = p;
    uint4 a0 = make_uint4(0, 0, 0, 0), b0 = a0;
    for (u32 q = p;; q += 32) {
        const u32 i = q + lane;
        const bool in = i < n && rec.read(i) == r0;
        const u32 inb = __ballot_sync(FULL, in);
        const u32 g = in ? rec.refid(i) : 0u;
        const bool v = in && g < P.G && is_valid(P.vb, g);
        const u32 V = __ballot_sync(FULL, v);
        if (!have && V) {
            const int f = __ffs(V) - 1;
            g0 = __shfl_sync(FULL, g, f);
            lead = q + f;
            a0 = __ldg(P.lin4 + 2 * (u64)g0); b0 = __ldg(P.lin4 + 2 * (u64)g0 + 1);
            have = true;
        }
        if (have) {
            const bool d = v && g != g0;
            const u32 mine = d ? lineage_diff(P.lin4, g, a0, b0) : 0u;
            multi |= __any_sync(FULL, d);
            neq |= __reduce_or_sync(FULL, mine);
            gmax = max(gmax, __reduce_max_sync(FULL, v ? g : 0u));
        }
        if (inb != FULL) { end = q + (inb == 0 ? 0 : 32 - __clz(inb)); break; }
    }
    if (!have) return;
    if (!multi) {                                                  // sole survivor of a read with several targets
        if (lane == 0) {
            atomicAdd(P.uniq2_extra + g0, 1u);
            if (P.cov2) atomicAdd(P.cov2 + bin_of(P.meta, g0, rec.upos(lead), P.half_avg, P.wdiv), 1u);
            if (P.res_kind) { P.res_kind[p] = 1; P.res_val[p] = g0; }
        }
        return;
    }
    const u32 eq = ~neq & 0xFFu;
    const bool fb = eq == 0;                                       // no level agrees: slot 7 of the largest reference id
    const u32 level = fb ? 7u : (u32)(__ffs(eq) - 1), owner = fb ? gmax : g0;
    for (u32 q = p; q < end; q += 32) {                            // children[lca] U= S
        const u32 i = q + lane;
        if (i < end) {
            const u32 h = rec.refid(i);
            if (h < P.G && is_valid(P.vb, h)) mark_child(P, h, fb, level, owner);
        }
    }
    if (lane == 0) {
        count_lca(s_key, s_val, P.lca_cnt, owner * 8 + level);
        if (P.res_kind) { P.res_kind[p] = 2; P.res_val[p] = __ldg(reinterpret_cast<const u32 *>(P.lin4) + (u64)owner * 8 + level); }
    }
}

// ------------------------------------------------------------------------------------------------
// K5+K6, one THREAD per multi-target read.  k_coverage lists, per chunk, where every multi-target read
// starts inside the chunk's compact words (rs); a warp takes a chunk and its lanes take 32 reads at a time.
// A read's few words are fetched four at a time and the lineage rows of all four are requested before any
// of them is needed, so a read costs two or three memory round trips instead of one per word and level.
// L16: lineages as 8 x 16-bit per-level dense taxon indices (one 16-byte row per reference; equality of the
// indices is equality of the taxon ids, zeros included) - used whenever there are fewer than 65536
// references; otherwise the 8 x 32-bit taxon ids themselves (two 16-byte rows).
// Reads longer than 32 records are still walked by the whole warp (assign_long_run).
// ------------------------------------------------------------------------------------------------
#ifndef ASSIGN_READS_OCC
#define ASSIGN_READS_OCC 4          // CTAs per SM the register allocation of k_assign_reads is bounded for
#endif
struct Lin16 { uint4 v; };
struct Lin32 { uint4 a, b; };
__device__ __forceinline__ void lin_load(const uint4 *__restrict__ t, u32 g, Lin16 &o) { o.v = __ldg(t + g); }
__device__ __forceinline__ void lin_load(const uint4 *__restrict__ t, u32 g, Lin32 &o) { o.a = __ldg(t + 2 * (u64)g); o.b = __ldg(t + 2 * (u64)g + 1); }
__device__ __forceinline__ bool lin_row_valid(const Lin16 &o) { return (o.v.w >> 31) != 0; }    // k_lin_valid's bit
__device__ __forceinline__ bool lin_row_valid(const Lin32 &) { return false; }                  // (never used: VROW is a Lin16 matter)
__device__ __forceinline__ void lin_zero(Lin16 &o) { o.v = make_uint4(0, 0, 0, 0); }
__device__ __forceinline__ void lin_zero(Lin32 &o) { o.a = make_uint4(0, 0, 0, 0); o.b = o.a; }
// acc |= x ^ y, slot by slot
__device__ __forceinline__ void lin_acc(Lin16 &acc, const Lin16 &x, const Lin16 &y)
{
    acc.v.x |= x.v.x ^ y.v.x; acc.v.y |= x.v.y ^ y.v.y; acc.v.z |= x.v.z ^ y.v.z; acc.v.w |= x.v.w ^ y.v.w;
}
__device__ __forceinline__ void lin_acc(Lin32 &acc, const Lin32 &x, const Lin32 &y)
{
    acc.a.x |= x.a.x ^ y.a.x; acc.a.y |= x.a.y ^ y.a.y; acc.a.z |= x.a.z ^ y.a.z; acc.a.w |= x.a.w ^ y.a.w;
    acc.b.x |= x.b.x ^ y.b.x; acc.b.y |= x.b.y ^ y.b.y; acc.b.z |= x.b.z ^ y.b.z; acc.b.w |= x.b.w ^ y.b.w;
}
// bit l set: some surviving reference differs from the first one on level l
__device__ __forceinline__ u32 lin_neq(const Lin16 &acc)
{
    const u32 w[4] = {acc.v.x, acc.v.y, acc.v.z, acc.v.w};
    u32 m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) m |= ((u32)((w[k] & 0xFFFFu) != 0) << (2 * k)) | ((u32)((w[k] >> 16) != 0) << (2 * k + 1));
    return m;
}
__device__ __forceinline__ u32 lin_neq(const Lin32 &acc)
{
    return (u32)(acc.a.x != 0) | ((u32)(acc.a.y != 0) << 1) | ((u32)(acc.a.z != 0) << 2) | ((u32)(acc.a.w != 0) << 3) |
           ((u32)(acc.b.x != 0) << 4) | ((u32)(acc.b.y != 0) << 5) | ((u32)(acc.b.z != 0) << 6) | ((u32)(acc.b.w != 0) << 7);
}

// the per-sample copy of the 16-bit lineage rows with the valid bit of the reference folded into bit 15 of slot 7 (the dense
// index of the top level never needs it): one 16-byte gather per compact word instead of a row and a bit from two tables
__global__ void k_lin_valid(const uint4 *__restrict__ lin16, const u32 *__restrict__ vb, u32 G, uint4 *__restrict__ out)
{
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    uint4 r = lin16[g];
    r.w = (r.w & 0x7FFFFFFFu) | (is_valid(vb, g) ? 0x80000000u : 0u);
    out[g] = r;
}

#define ASG_STAGE 512u              // compact words of a warp's 32 reads staged in shared memory (more: read in place)

// EXTRA: uniq_cov2 bins / per-read results are wanted (they need the record index of every compact word)
// VROW:  lin_tab is the per-sample table of k_lin_valid (Lin16 only): the row carries the valid bit
template <class Rec, class Lin, bool EXTRA, bool VROW>
__global__ void __launch_bounds__(256, EXTRA ? 4 : ASSIGN_READS_OCC)
k_assign_reads(const __grid_constant__ Rec rec, u32 n, const __grid_constant__ AssignParams P, const u32 *__restrict__ rs_all, const uint4 *__restrict__ lin_tab)
{
    __shared__ u32 s_key[LCA_CACHE], s_val[LCA_CACHE];
    __shared__ u32 s_stage[8][ASG_STAGE];
    const u32 lane = threadIdx.x & 31;
    u32 *const st = s_stage[threadIdx.x >> 5];
    for (u32 k = threadIdx.x; k < LCA_CACHE; k += blockDim.x) { s_key[k] = LCA_EMPTY; s_val[k] = 0; }
    __syncthreads();
    const u32 n_chunks = n / CHUNK + (n % CHUNK != 0);
    const u32 wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (u32 c = wg; c < n_chunks; c += nw) {
        const uint2 cnt = __ldg(P.chunk_cnt + c);
        const u32 n_lr = cnt.y & 0xFFu, n_rs = cnt.y >> 8;               // long runs, multi-target reads of the chunk
        const u32 *cw = P.cw + (u64)c * CW_SLOT;
        const u32 *rs = rs_all + (u64)c * RS_SLOT;
        for (u32 k0 = 0; k0 < n_rs; k0 += 32) {
            const u32 k = k0 + lane;
            const bool active = k < n_rs;
            const u32 entry = active ? __ldg(rs + k) : 0u;         // start | words << 16 (at most 32 words)
            const u32 start = entry & 0xFFFFu, end = start + (entry >> 16);
            // the words of the warp's 32 reads lie next to each other (gaps where repeat hits were): one coalesced copy into shared
            // memory, then every lane walks its own read there instead of 32 lanes reading 32 places of global memory
            const u32 lo = __shfl_sync(FULL, start, 0), hi = __shfl_sync(FULL, end, (int)min(31u, n_rs - k0 - 1u));
            const bool staged = hi - lo <= ASG_STAGE;              // warp-uniform
            __syncwarp();
            if (staged) for (u32 w = lane; w < hi - lo; w += 32) st[w] = __ldcs(cw + lo + w);
            __syncwarp();
            auto word = [&](u32 j) { return (staged ? st[j - lo] : __ldg(cw + j)) & ~CW_HEAD; };
            if (!active) continue;
            u32 g0 = 0, j0 = start, ns = 0, gmax = 0, vmask = 0;  // vmask: which of the read's (at most 32) words survive
            Lin l0, acc;
            lin_zero(l0); lin_zero(acc);
            for (u32 j = start; j < end; j += 4) {
                u32 g[4]; bool v[4]; Lin l[4];
#pragma unroll
                for (int d = 0; d < 4; ++d) g[d] = j + d < end ? word(j + d) : 0xFFFFFFFFu;
#pragma unroll
                for (int d = 0; d < 4; ++d) {                      // the valid bit and the lineage row travel together
                    v[d] = false;
                    lin_zero(l[d]);
                    if (g[d] != 0xFFFFFFFFu) {
                        lin_load(lin_tab, g[d], l[d]);
                        v[d] = VROW ? lin_row_valid(l[d]) : is_valid(P.vb, g[d]);
                    }
                }
#pragma unroll
                for (int d = 0; d < 4; ++d)
                    if (v[d]) {
                        if (ns == 0) { g0 = g[d]; j0 = j + d; l0 = l[d]; }
                        else lin_acc(acc, l[d], l0);
                        gmax = max(gmax, g[d]);
                        vmask |= 1u << (j + d - start);
                        ++ns;
                    }
            }
            if (ns == 1) {                                         // sole survivor: the read became unique through the filter
                atomicAdd(P.uniq2_extra + g0, 1u);
                if (EXTRA && P.cw_idx) {
                    const u32 lead = P.cw_idx[(u64)c * CW_SLOT + j0];
                    if (P.cov2) atomicAdd(P.cov2 + bin_of(P.meta, g0, rec.upos(lead), P.half_avg, P.wdiv), 1u);
                    if (P.res_kind) { const u32 hd = P.cw_idx[(u64)c * CW_SLOT + start]; P.res_kind[hd] = 1; P.res_val[hd] = g0; }
                }
            } else if (ns >= 2) {
                const u32 eq = ~lin_neq(acc) & 0xFFu;
                const bool fb = eq == 0;                           // no level agrees: slot 7 of the largest reference id
                const u32 level = fb ? 7u : (u32)(__ffs(eq) - 1), owner = fb ? gmax : g0;
                // children[lca] U= S: marks are tested before they are written (the hot ones are set early and then only
                // read, which keeps them shared in L2 instead of bouncing), four tests in flight at a time
                u32 *mk_base = fb ? P.fb_mark + (u64)__ldg(P.top_idx + owner) * P.G : P.child_mark + level;
                const u32 mk_stride = fb ? 1u : 8u;
                while (vmask) {
                    u32 *mk[4]; u32 old[4];
#pragma unroll
                    for (int d = 0; d < 4; ++d) {
                        mk[d] = nullptr;
                        if (vmask) {
                            const u32 b = (u32)__ffs(vmask) - 1u;
                            vmask &= vmask - 1u;
                            mk[d] = mk_base + (u64)word(start + b) * mk_stride;
                        }
                    }
#pragma unroll
                    for (int d = 0; d < 4; ++d) old[d] = mk[d] ? *mk[d] : 1u;
#pragma unroll
                    for (int d = 0; d < 4; ++d) if (old[d] == 0) *mk[d] = 1u;
                }
                count_lca(s_key, s_val, P.lca_cnt, owner * 8 + level);
                if (EXTRA && P.res_kind) {
                    const u32 hd = P.cw_idx[(u64)c * CW_SLOT + start];
                    P.res_kind[hd] = 2;
                    P.res_val[hd] = __ldg(reinterpret_cast<const u32 *>(P.lin4) + (u64)owner * 8 + level);
                }
            }
        }
        for (u32 k = 0; k < n_lr; ++k) assign_long_run(rec, __ldg(P.lr + c * LR_SLOT + k), n, lane, P, s_key, s_val);
    }
    __syncthreads();
    for (u32 k = threadIdx.x; k < LCA_CACHE; k += blockDim.x)
        if (s_key[k] != LCA_EMPTY && s_val[k]) atomicAdd(P.lca_cnt + s_key[k], s_val[k]);
}

// the heads of single-target reads were marked 3 by k_coverage: kind 1 with the reference when it survived
__global__ void k_read_results_unique(unsigned char *__restrict__ res_kind, u32 *__restrict__ res_val, const u32 *__restrict__ ref,
                                      u32 ref_stride, const u32 *__restrict__ vb, u32 n)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || res_kind[i] != 3) return;
    const u32 g = ref[(u64)i * ref_stride];
    if (is_valid(vb, g)) { res_kind[i] = 1; res_val[i] = g; } else res_kind[i] = 0;
}

// ------------------------------------------------------------------------------------------------
// K7: per-rank segmented reduction of read counts and contributing-reference sets, for databases whose
// lineage table is tree-consistent (every taxon sits on one level, carries that rank in db.taxid__name, and
// all references under it share the levels above; no zero slots).  There phases 2 and 3 of
// get_reads_lca_count (src/slimm.hpp:560-610) reduce to sums over the references below a taxon:
//   count[t on level L]    = sum over g under t of ( uniq_reads_count2[g] + sum_{l <= L} lca_count[g][l] )
//   children[t on level L] = { g under t : g carries a child mark on a level <= L, or uniq_reads_count2[g] > 0 }
//                            (+ the fallback marks of level-7 taxa)
// and write_abundance (:733-843) needs count, |children|, sum of their lengths, min and max per taxon of the
// requested rank and its parent rank.  One thread per reference; taxa are addressed by a per-level dense
// index (lvl_idx[L][g]).  Any other database takes the general host path (profile_host.cpp).
// agg layout: [which 0: rank, 1: parent rank][count | kn | klen | kmin | kmax][G]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_rank_reduce(const u32 *__restrict__ assign /* uniq2[G] | lca[G*8] | child_mark[G*8] | fb_mark[n_top*G] */, const uint4 *__restrict__ meta,
              const u32 *__restrict__ lvl_idx /*[8][G]*/, const u32 *__restrict__ top_lvl7 /*[n_top] level-7 index of fallback row*/,
              u32 G, u32 n_top, u32 rk, u32 *__restrict__ agg)
{
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const u32 u2 = assign[g];
    const u32 *lc = assign + G + (u64)g * 8, *cm = assign + (u64)9 * G + (u64)g * 8;   // G need not be a multiple of 4: scalar loads
    u32 c[8], marks = 0;
#pragma unroll
    for (int l = 0; l < 8; ++l) { c[l] = lc[l]; marks |= (u32)(cm[l] != 0) << l; }
    const u32 len = meta[g].x;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        const u32 L = rk + which;
        u32 cnt = u2;
#pragma unroll
        for (int l = 0; l < 8; ++l) if ((u32)l <= L) cnt += c[l];
        const u32 t = __ldg(lvl_idx + (u64)L * G + g);
        u32 *a = agg + (u64)which * 5 * G;
        if (cnt) atomicAdd(a + t, cnt);
        if ((marks & ((2u << L) - 1u)) != 0 || u2 > 0) {
            atomicAdd(a + G + t, 1u);
            atomicAdd(a + 2 * (u64)G + t, len);
            atomicMin(a + 3 * (u64)G + t, g);
            atomicMax(a + 4 * (u64)G + t, g);
        }
    }
    if (rk + 1 == 7)                                               // children of a fallback LCA may lie outside its subtree
        for (u32 r = 0; r < n_top; ++r)
            if (assign[(u64)17 * G + (u64)r * G + g]) {
                u32 *a = agg + (u64)5 * G;
                const u32 t = top_lvl7[r];
                atomicAdd(a + G + t, 1u);
                atomicMin(a + 3 * (u64)G + t, g);
            }
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
// slimm_gpu_push_packed: the wire format of grouped input (one "new read" bit + a 16-bit reference id per record) back to
// the struct of arrays the kernels read.  read_id[i] = id_base + (set bits among records 0..i) - 1.
//   k_unpack_count   set bits per 1024-record tile
//   k_unpack_scan    exclusive scan of the tile counts (one CTA), continued from / written back to the running id counter
//   k_unpack_write   a warp per tile: word j of the tile is broadcast, lane l writes record 32 j + l (coalesced)
// ------------------------------------------------------------------------------------------------
#define UNPACK_TILE 1024u
__global__ void k_unpack_count(const u32 *__restrict__ bits, u64 n, u32 *__restrict__ tile_cnt)
{
    const u32 lane = threadIdx.x & 31;
    const u64 tile = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_tiles = (n + UNPACK_TILE - 1) / UNPACK_TILE;
    if (tile >= n_tiles) return;
    const u64 i0 = tile * UNPACK_TILE + 32ull * lane;            // my word's first record
    u32 w = i0 < n ? bits[i0 >> 5] : 0u;
    if (i0 < n && n - i0 < 32) w &= (1u << (n - i0)) - 1u;
    const u32 c = warp_sum((u32)__popc(w));
    if (lane == 0) tile_cnt[tile] = c;
}
__global__ void __launch_bounds__(1024) k_unpack_scan(u32 *__restrict__ tile_cnt, u64 n_tiles, u32 *__restrict__ id_counter)
{
    __shared__ u32 s_warp[32];
    __shared__ u32 s_run;
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_run = *id_counter;
    __syncthreads();
    for (u64 t0 = 0; t0 < n_tiles; t0 += 1024) {
        const u64 t = t0 + tid;
        const u32 v = t < n_tiles ? tile_cnt[t] : 0u;
        u32 x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL, x, o); if ((int)lane >= o) x += y; }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        u32 wbase = 0, all = 0;
#pragma unroll
        for (int k = 0; k < 32; ++k) { const u32 q = s_warp[k]; if (k < (int)wid) wbase += q; all += q; }
        const u32 run = s_run;
        if (t < n_tiles) tile_cnt[t] = run + wbase + x - v;      // ids handed out before this tile
        __syncthreads();
        if (tid == 0) s_run = run + all;
        __syncthreads();
    }
    if (tid == 0) *id_counter = s_run;
}
__global__ void k_unpack_write(const u32 *__restrict__ bits, const unsigned short *__restrict__ ref16, u64 n, const u32 *__restrict__ tile_base,
                               u32 *__restrict__ rid, u32 *__restrict__ ref)
{
    const u32 lane = threadIdx.x & 31;
    const u64 tile = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_tiles = (n + UNPACK_TILE - 1) / UNPACK_TILE;
    if (tile >= n_tiles) return;
    const u64 t0 = tile * UNPACK_TILE, i0 = t0 + 32ull * lane;
    u32 w = i0 < n ? bits[i0 >> 5] : 0u;
    if (i0 < n && n - i0 < 32) w &= (1u << (n - i0)) - 1u;
    u32 incl = (u32)__popc(w);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o) incl += y; }
    const u32 before = tile_base[tile] + incl - (u32)__popc(w);  // set bits before my word
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
        const u32 wj = __shfl_sync(FULL, w, j), bj = __shfl_sync(FULL, before, j);
        const u64 i = t0 + 32ull * j + lane;
        if (i < n) {
            rid[i] = bj + (u32)__popc(wj & LANE_LE(lane)) - 1u;
            ref[i] = ref16[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// helpers for the unsorted-input path and bin readout
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_values(const u32 *__restrict__ ref, const i32 *__restrict__ pos, u64 n, uint2 *__restrict__ out)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = make_uint2(ref[i], (u32)pos[i]);
}

// the bins of fine-slice runs with compact storage: slices with fewer than 65536 items live in hist16 as {cov:16 | uniq_cov:16},
// the hot ones in the interleaved 64-bit histogram
__global__ void k_extract_bins_compact(const u32 *__restrict__ hist64, const u32 *__restrict__ hist16, const u32 *__restrict__ fine_start, u64 first, u32 word,
                                       u32 nb, u32 *__restrict__ out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const u64 b = first + i;
    if (fine_slice_is_wide(fine_start, b)) out[i] = hist64[b * 2 + word];
    else { const u32 w = hist16[b]; out[i] = word ? w >> 16 : w & 0xFFFFu; }
}

__global__ void k_extract_bins(const u32 *__restrict__ src, u64 first, u32 stride_words, u32 word, u32 nb, u32 *__restrict__ out)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) out[b] = src[(first + b) * stride_words + word];
}

// coverage_tile.cuh - K1 coverage on warp-private tiles (included by kernels.cuh; same contract and outputs as k_coverage).
//
// Replaces reference src/slimm.hpp:194-257 + src/read_stat.hpp:116-135,72-75.
//
// k_coverage (the sliding 32-record window) spends ~260 thread instructions per record on run analysis that only the
// records of multi-record reads need.  Here a warp owns a tile of COVT_T consecutive records and works in two phases:
//
//   1  record-parallel, four consecutive records per lane (128-bit loads of read ids and reference ids): head bits (a
//      record starts a read) go to a bit map in shared memory, reference ids are staged in shared memory.
//   2  read-parallel: the reads of two or more records are listed from the bit map (a fifth of the reads) and ONE LANE
//      walks one read over the staged reference ids: repeat hits (src/read_stat.hpp:125-131 keeps the first record of a
//      pair only), reads whose records all name one reference (unique after all) and the distinct references of the
//      multi-target reads (the compact stream k_assign_reads works on) become three more bit maps.  Reads longer than
//      32 records are walked by the whole warp.
//   3  record-parallel again (128-bit loads of the positions, staged reference ids): bins, the final item of every
//      record in one 128-bit store, slice counts, and the compact words, compacted with one warp scan per step.
// Every array is read once and every output is written once, coalesced; nothing is patched afterwards.
//
// Ownership is positional: a tile writes the items of ITS records only.  A read that straddles a tile border is walked
// by both tiles (32 staged records of halo on either side; reading only), each marking its own records; the tile that
// holds the read's first record counts it and emits its compact words.  No tile ever writes into another tile's range,
// so there is no ordering between warps to get right.
#pragma once

#ifndef COVT_T
#define COVT_T 512u                       // records per tile
#endif
#define COVT_STEPS (COVT_T / 128u)        // steps of the record-parallel passes (4 records per lane)
#define COVT_WORDS (COVT_T / 32u)         // bit-map words of the tile proper
#define COVT_SG (COVT_T + 64u)            // staged reference ids: 32 halo | tile | 32 halo
#define COVT_HB (COVT_WORDS + 4u)         // head bits: halo word | tile | halo word | all-ones sentinel | pad
#define COVT_BM (COVT_WORDS + 2u)         // a per-record bit map of the tile (+ spill words: a read's bits may reach 31 records past the tile)
#define COVT_LIST (COVT_T / 2u + 2u)      // reads of two or more records headed in the tile, + the one reaching in from the left
#define COVT_LONG 36u                     // reads of more than 32 records that touch the tile (at most T/33 + 2)
#define COVT_WARP_WORDS ((COVT_SG + COVT_HB + 3 * COVT_BM + COVT_T / 4 + COVT_LIST / 2 + COVT_LONG / 2 + 3u) & ~3u)   // per warp, rounded up to 16 bytes
#define COVT_THREADS 256
#define COVT_POS_UNKNOWN 0xFFFFu          // long list: the read starts further left than the halo
static_assert(COVT_WORDS <= 32 && COVT_T % 128 == 0, "one bit-map word per lane, whole steps");
static_assert(CHUNK % COVT_T == 0, "a chunk is a whole number of tiles");

// records base .. base+3 of one array; `left` = records from base to the end of the data
template <class F>
__device__ __forceinline__ uint4 covt_load4(F one, u32 base, u32 left)
{
    u32 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (u32)k < left ? one(base + k) : 0u;
    return make_uint4(v[0], v[1], v[2], v[3]);
}

__device__ __forceinline__ u32 covt_hash6(u32 g) { return (g * 0x9E3779B1u) >> 26; }

// bits [at, at + 32) of a bit map (at + 32 may reach one word past `at`'s)
__device__ __forceinline__ u32 covt_bits(const u32 *bm, u32 at) { return __funnelshift_r(bm[at >> 5], bm[(at >> 5) + 1], at & 31); }
// OR `bits` into a bit map starting at bit `at`
__device__ __forceinline__ void covt_or_bits(u32 *bm, u32 at, u32 bits)
{
    const u32 lo = bits << (at & 31), hi = (at & 31) ? bits >> (32 - (at & 31)) : 0u;
    if (lo) atomicOr(bm + (at >> 5), lo);
    if (hi) atomicOr(bm + (at >> 5) + 1, hi);
}

// one more item of a histogram slice: a shared-memory RED on a shared-space address kept in a register
__device__ __forceinline__ void covt_count(u32 s_cnt_addr, u32 slice)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(s_cnt_addr + 4u * slice) : "memory");
}
#define COVT_DUMMY_SLICE MAX_BUCKETS      // counter of the records that contribute nothing (keeps the RED unconditional)

// A read of more than 32 records, walked by the whole warp on the records themselves: s = its first record (anywhere),
// [t0, t1) = the tile.  Marks the tile's repeat hits in rb / the read as unique in ub; the tile that holds the first
// record counts the read and lists it for k_assign_reads (lr).
template <class Rec, bool EXTRA>
__device__ __noinline__ void covt_long_run(const Rec &rec, u32 s, u32 t0, u32 t1, u32 n, u32 lane, const CovParams &P, u32 *rb, u32 *ub, u32 *uniq_io, u32 lr_base,
                                           u32 *n_lr_io)
{
    const u32 r0 = rec.read(s), gh = rec.refid(s);
    bool multi = false;
    u32 end = s;
    for (u32 q = s;; q += 32) {                                    // where the read ends; one reference or several
        const bool in = n - q > lane && rec.read(q + lane) == r0;  // q < n always
        const u32 inb = __ballot_sync(FULL, in);
        multi |= __any_sync(FULL, in && rec.refid(q + lane) != gh);
        if (inb != FULL) { end = q + (inb == 0 ? 0 : 32 - __clz(inb)); break; }
        if (n - q <= 32) { end = n; break; }
    }
    if (s >= t0) {                                                 // the read starts inside this tile
        u32 n_lr = *n_lr_io;
        if (multi) { if (lane == 0 && n_lr < LR_SLOT) P.lr[lr_base + n_lr] = s; *n_lr_io = n_lr + 1; }
        else { *uniq_io += (lane == 0); if (lane == 0) atomicOr(ub + ((s - t0) >> 5), 1u << ((s - t0) & 31)); }
    }
    const u32 a = max(s, t0), b = min(end, t1);
    for (u32 q = a; q < b; q += 32) {
        const u32 i = q + lane;
        if (i < b) {
            bool rep = i != s;                                     // one reference only: every record but the first repeats it
            if (multi) { const u32 g = rec.refid(i); rep = false; for (u32 j = s; j < i; ++j) if (rec.refid(j) == g) { rep = true; break; } }
            if (rep) atomicOr(rb + ((i - t0) >> 5), 1u << ((i - t0) & 31));
        }
    }
}

// ---- pass 1: read ids and reference ids -> head bits + staged reference ids ---------------------------------------
// WHOLE: all COVT_T records of the tile exist
template <class Rec, bool WHOLE>
__device__ __forceinline__ void covt_pass1(const Rec &rec, u32 t0, u32 rem, u32 lane, u32 carry /* read id of record t0 - 1 */, u32 *sg, u32 *hb, u32 &bad)
{
#pragma unroll
    for (u32 s = 0; s < COVT_STEPS; ++s) {
        const u32 o = s * 128u;
        if (!WHOLE && o >= rem) {                                  // the whole step lies behind the data: heads only
            if ((lane & 7) == 0) hb[1 + 4 * s + (lane >> 3)] = FULL;
            continue;
        }
        const u32 base = t0 + o + 4 * lane;
        const bool all = WHOLE || rem - o >= 128u;
        const u32 left = all ? 4u : (rem - o > 4 * lane ? rem - o - 4 * lane : 0u);   // my records that exist
        uint4 r4, g4;
        if (all) rec.load_rg4(base, r4, g4);
        else { r4 = covt_load4([&](u32 i) { return rec.read(i); }, base, left); g4 = covt_load4([&](u32 i) { return rec.refid(i); }, base, left); }
        const u32 r[4] = {r4.x, r4.y, r4.z, r4.w};
        u32 prev = __shfl_up_sync(FULL, r[3], 1);
        if (lane == 0) prev = carry;
        carry = __shfl_sync(FULL, r[3], 31);
        // a record behind the end of the data reads as the head of a read: it ends the last real one
        const bool very_first = t0 + o == 0 && lane == 0;          // record 0 of the sample has no predecessor
        bool h[4], down;
        h[0] = left < 1 || r[0] != prev || very_first;
        down = left >= 1 && r[0] < prev && !very_first;            // read ids must be non-decreasing
#pragma unroll
        for (int k = 1; k < 4; ++k) { h[k] = left <= (u32)k || r[k] != r[k - 1]; down |= left > (u32)k && r[k] < r[k - 1]; }
        if (down) bad |= 1u;
        u32 v = ((u32)h[0] | ((u32)h[1] << 1) | ((u32)h[2] << 2) | ((u32)h[3] << 3)) << (4 * (lane & 7));
        v |= __shfl_xor_sync(FULL, v, 1); v |= __shfl_xor_sync(FULL, v, 2); v |= __shfl_xor_sync(FULL, v, 4);   // eight lanes make one word
        if ((lane & 7) == 0) hb[1 + 4 * s + (lane >> 3)] = v;
        *reinterpret_cast<uint4 *>(sg + 32 + o + 4 * lane) = g4;
    }
}

// ---- pass 3: positions -> bins; final items, slice counts, compact words of the multi-target reads -----------------
struct CovtOut { u32 n_cw, n_rs; };       // compact words / multi-target reads of the chunk so far (warp-uniform)

template <class Rec, int MODE, bool EXTRA, bool TEXG, bool ALL128>
__device__ __forceinline__ void covt_step3(const Rec &rec, const CovParams &P, uint4 p4, u32 base, u32 left /* my records that exist (ALL128: 4) */, u32 lane,
                                           const u32 *sg_step, const u32 *hb, const u32 *rb, const u32 *ub, const u32 *kb, const unsigned char *cnt, u32 at /* tile-relative record of my first */,
                                           u32 s_cnt_addr, u32 *cw_c, u32 *cwi_c, u32 *rs_c, CovtOut &out, u32 &bad)
{
    const uint4 g4 = *reinterpret_cast<const uint4 *>(sg_step + 4 * lane);
    const u32 g[4] = {g4.x, g4.y, g4.z, g4.w}, ps[4] = {p4.x, p4.y, p4.z, p4.w};
    const u32 h5 = covt_bits(hb + 1, at);                          // bit k (0..4): record at + k starts a read
    const u32 singles = h5 & (h5 >> 1);                            // a read of one record: unique
    const u32 rep4 = covt_bits(rb, at), uq4 = covt_bits(ub, at) | singles, keep4 = covt_bits(kb, at) & 15u;
    bool put[4];
    uint2 m2[4];
    uint4 m4[4];
    const bool check = !ALL128 || max(max(g[0], g[1]), max(g[2], g[3])) >= P.G;   // rare: a reference id out of range / the end of the data
#pragma unroll
    for (int k = 0; k < 4; ++k) {                                  // four gathers in flight
        const bool ok = !check || g[k] < P.G;
        put[k] = ok && (!check || (u32)k < left) && !((rep4 >> k) & 1u);
        if (check && (u32)k < left && !ok) bad |= 2u;
        const u32 gg = ok ? g[k] : 0u;
        if (MODE == 0) m4[k] = __ldg(P.meta + gg);
        else if (TEXG) m2[k] = tex1Dfetch<uint2>(P.meta2_tex, (int)gg);   // random gathers: the texture pipe instead of 32 LSU wavefronts
        else m2[k] = __ldg(P.meta2 + gg);
    }
    u32 item[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool uq = (uq4 >> k) & 1u;
        if (MODE == 0) {
            const u64 b = bin_of_meta(m4[k], ps[k], P.half_avg, P.wdiv);
            if (put[k]) atomicAdd(P.hist + b, uq ? 0x100000001ull : 1ull);   // cov += 1 [, uniq_cov += 1]
        } else {                                                   // padded bin ids fit 31 bits on this path
            const u32 b = m2[k].y + fast_div(min(ps[k] + P.half_avg, m2[k].x), P.wdiv);
            item[k] = put[k] ? (b | (uq ? 0x80000000u : 0u)) : ITEM_SKIP;
            covt_count(s_cnt_addr, put[k] ? b >> P.shift : COVT_DUMMY_SLICE);
        }
        if (EXTRA && P.res_kind && uq && put[k]) P.res_kind[base + k] = 3;
    }
    if (MODE == 1) {
        if (ALL128) __stcs(reinterpret_cast<uint4 *>(P.items + base), make_uint4(item[0], item[1], item[2], item[3]));
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) if ((u32)k < left) __stcs(P.items + base + k, item[k]);
        }
    }
    // compact stream: the kept records' reference ids, in record order (bit 31 marks a read's first); one rs entry per read
    const u32 kh4 = keep4 & h5;                                    // kept heads: multi-target reads that start here
    u32 incl = (u32)__popc(keep4) | ((u32)__popc(kh4) << 16);      // words | reads, scanned together
    if (__any_sync(FULL, keep4 != 0)) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o) incl += y; }
        const u32 tot = __shfl_sync(FULL, incl, 31);
        u32 o = out.n_cw + (incl & 0xFFFFu) - (u32)__popc(keep4), q = out.n_rs + (incl >> 16) - (u32)__popc(kh4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((keep4 >> k) & 1u) {
                const bool head = (h5 >> k) & 1u;
                cw_c[o] = (g[k] < P.G ? g[k] : 0u) | (head ? CW_HEAD : 0u);   // an id out of range is reported later; the word stays harmless
                if (EXTRA && cwi_c) cwi_c[o] = base + k;
                if (head) rs_c[q++] = o | ((u32)cnt[at + k] << 16);
                ++o;
            }
        out.n_cw += tot & 0xFFFFu; out.n_rs += tot >> 16;
    }
}

template <class Rec, int MODE, bool EXTRA, bool TEXG, bool WHOLE>
__device__ __forceinline__ void covt_pass3(const Rec &rec, const CovParams &P, u32 t0, u32 rem, u32 lane, const u32 *sg, const u32 *hb, const u32 *rb, const u32 *ub,
                                           const u32 *kb, const unsigned char *cnt, u32 s_cnt_addr, u32 *cw_c, u32 *cwi_c, u32 *rs_c, CovtOut &out, u32 &bad)
{
    const u32 mine0 = rem > 4 * lane ? rem - 4 * lane : 0u;
    uint4 cur = (WHOLE || rem >= 128u) ? rec.load_p4(t0 + 4 * lane) : covt_load4([&](u32 i) { return rec.upos(i); }, t0 + 4 * lane, mine0);
#pragma unroll
    for (u32 s = 0; s < COVT_STEPS; ++s) {
        const u32 o = s * 128u;
        if (!WHOLE && o >= rem) break;
        const u32 left_step = WHOLE ? COVT_T - o : rem - o;        // records from this step on
        uint4 nx = cur;
        if (s + 1 < COVT_STEPS && (WHOLE || left_step > 128u)) {   // the next step's positions fly during this step's work
            const u32 l2 = left_step - 128u, m2 = l2 > 4 * lane ? l2 - 4 * lane : 0u;
            nx = (WHOLE || l2 >= 128u) ? rec.load_p4(t0 + o + 128u + 4 * lane) : covt_load4([&](u32 i) { return rec.upos(i); }, t0 + o + 128u + 4 * lane, m2);
        }
        const u32 mine = left_step > 4 * lane ? min(left_step - 4 * lane, 4u) : 0u;
        if (WHOLE || left_step >= 128u)
            covt_step3<Rec, MODE, EXTRA, TEXG, true>(rec, P, cur, t0 + o + 4 * lane, 4u, lane, sg + 32 + o, hb, rb, ub, kb, cnt, o + 4 * lane, s_cnt_addr, cw_c, cwi_c, rs_c, out, bad);
        else
            covt_step3<Rec, MODE, EXTRA, TEXG, false>(rec, P, cur, t0 + o + 4 * lane, mine, lane, sg + 32 + o, hb, rb, ub, kb, cnt, o + 4 * lane, s_cnt_addr, cw_c, cwi_c, rs_c, out, bad);
        cur = nx;
    }
}

template <class Rec, int MODE, bool EXTRA, bool TEXG>
__global__ void __launch_bounds__(COVT_THREADS, 4)
k_coverage_tile(const __grid_constant__ Rec rec, u32 n, const __grid_constant__ CovParams P)
{
    extern __shared__ __align__(16) u32 covt_smem[];
    __shared__ u32 s_cnt[MODE ? MAX_BUCKETS + 1 : 1];              // items per histogram slice (shared-memory REDs) + the dummy slot
    __shared__ u32 s_h, s_u, s_b;
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    u32 *const sg = covt_smem + wid * COVT_WARP_WORDS;             // reference ids of records tile - 32 .. tile + T + 31
    u32 *const hb = sg + COVT_SG;                                  // head bits of the same records (+ sentinel word)
    u32 *const rb = hb + COVT_HB;                                  // per record of the tile: a repeat hit (contributes nothing)
    u32 *const ub = rb + COVT_BM;                                  // the first record of a read of several records that all name one reference: unique
    u32 *const kb = ub + COVT_BM;                                  // a distinct reference of a multi-target read headed in this tile: goes to the compact stream
    unsigned char *const cnt = reinterpret_cast<unsigned char *>(kb + COVT_BM);   // at a kept head: compact words of its read
    unsigned short *const list = reinterpret_cast<unsigned short *>(kb + COVT_BM + COVT_T / 4);
    unsigned short *const llist = list + COVT_LIST;
    if (MODE) for (u32 b = tid; b < MAX_BUCKETS; b += COVT_THREADS) s_cnt[b] = 0;
    if (tid == 0) { s_h = 0; s_u = 0; s_b = 0; }
    __syncthreads();
    u32 heads = 0, uniq = 0, bad = 0;                              // per lane, summed at the end
    u32 s_cnt_addr = (u32)__cvta_generic_to_shared(s_cnt);
    asm volatile("mov.u32 %0, %0;" : "+r"(s_cnt_addr));            // opaque: keeps the address in a register instead of re-deriving it per RED
    const u32 n_chunks = n / CHUNK + (n % CHUNK != 0);
    const u32 wg = (blockIdx.x * blockDim.x + tid) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (u32 c = wg; c < n_chunks; c += nw) {
        CovtOut out{0u, 0u};
        u32 n_lr = 0;                                              // long reads of this chunk (warp-uniform)
        u32 *const cw_c = P.cw + (u64)c * CW_SLOT;
        u32 *const cwi_c = EXTRA && P.cw_idx ? P.cw_idx + (u64)c * CW_SLOT : nullptr;
        u32 *const rs_c = P.rs + (u64)c * RS_SLOT;
        const u32 c0 = c * CHUNK;
        for (u32 tt = 0; tt < CHUNK && n - c0 > tt; tt += COVT_T) {
            const u32 t0 = c0 + tt, rem = n - t0;                  // rem >= 1 records from t0 on
            const u32 nrec = min(rem, COVT_T), t1 = t0 + nrec;
            __syncwarp();                                          // the previous tile is done with the staging area
            // ---- halos: 32 records on either side, reference ids + head bits; clear the bit maps --------------------
            u32 carry;                                             // read id of the record right before the tile
            {
                const bool vl = t0 + lane >= 32u;                  // record t0 - 32 + lane exists
                const u32 jl = t0 - 32u + lane;
                const u32 rl = vl ? rec.read(jl) : 0u;
                const u32 rlp = (vl && jl > 0) ? rec.read(jl - 1) : ~rl;
                sg[lane] = vl ? rec.refid(jl) : 0u;
                const u32 HBL = __ballot_sync(FULL, vl && rl != rlp);
                carry = __shfl_sync(FULL, rl, 31);                 // unused when t0 == 0
                const bool vr = rem > COVT_T + lane;               // record t0 + T + lane exists
                const u32 jr = t0 + COVT_T + lane;
                const u32 rr = vr ? rec.read(jr) : 0u, rrp = vr ? rec.read(jr - 1) : 1u;
                sg[32 + COVT_T + lane] = vr ? rec.refid(jr) : 0u;
                const u32 HBR = __ballot_sync(FULL, !vr || rr != rrp);
                if (lane == 0) { hb[0] = HBL; hb[COVT_WORDS + 1] = HBR; hb[COVT_WORDS + 2] = FULL; }
                for (u32 k = lane; k < 3 * COVT_BM; k += 32) rb[k] = 0u;   // rb | ub | kb are contiguous
            }
            // ---- pass 1 -------------------------------------------------------------------------------------------
            if (rem >= COVT_T) covt_pass1<Rec, true>(rec, t0, rem, lane, carry, sg, hb, bad);
            else covt_pass1<Rec, false>(rec, t0, rem, lane, carry, sg, hb, bad);
            __syncwarp();
            // ---- pass 2: the reads of two or more records, one lane per read ---------------------------------------
            u32 n_list, n_long = 0;
            {
                const u32 W = lane < COVT_WORDS ? hb[1 + lane] : 0u;
                u32 Wn = __shfl_down_sync(FULL, W, 1);
                if (lane == COVT_WORDS - 1) Wn = hb[COVT_WORDS + 1];
                const u32 N = (W >> 1) | (Wn << 31);               // bit b: the record behind b starts another read
                const u32 V = nrec >= 32 * lane + 32 ? FULL : (nrec > 32 * lane ? (1u << (nrec - 32 * lane)) - 1u : 0u);   // records that exist
                heads += __popc(W & V);
                uniq += __popc(W & N & V);
                u32 MH = W & ~N & V;                               // heads of reads of two or more records
                const u32 c_mh = __popc(MH);
                u32 incl = c_mh;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o) incl += y; }
                u32 off = incl - c_mh;
                n_list = __shfl_sync(FULL, incl, 31);
                while (MH) { const u32 b = (u32)__ffs(MH) - 1u; MH &= MH - 1u; list[off++] = (unsigned short)(32u + 32u * lane + b); }
                const u32 HBL = hb[0];
                if (t0 > 0 && !(hb[1] & 1u)) {                     // record t0 continues a read that started further left
                    if (HBL) { if (lane == 0) list[n_list] = (unsigned short)(31 - __clz(HBL)); ++n_list; }
                    else { if (lane == 0) llist[0] = COVT_POS_UNKNOWN; n_long = 1; }
                }
            }
            __syncwarp();
            u32 tail_keep = 0;                                     // one lane at most: kept records of its read behind the chunk's end
            for (u32 q0 = 0; q0 < n_list; q0 += 32) {
                const u32 q = q0 + lane;
                const bool act = q < n_list;
                u32 p = 32, len = 1;                               // staged position of the read's first record, its records
                bool lng = false;
                if (act) {
                    p = list[q];
                    const u32 win = covt_bits(hb, p + 1);          // head bits of positions p+1 .. p+32
                    lng = win == 0;
                    len = lng ? 1u : (u32)__ffs(win);
                }
                const u32 LB = __ballot_sync(FULL, act && lng);    // more than 32 records: the whole warp, below
                if (LB) { if (act && lng) llist[n_long + __popc(LB & LANE_LT(lane))] = (unsigned short)p; n_long += __popc(LB); }
                if (act && !lng) {
                    u32 rep = 0;                                   // bit i: record i of the read repeats an earlier reference of the read
                    bool multi = false;
                    const u32 g0 = sg[p];
                    // references within +-32 of the first one have a bit of their own in `near` (exact); the others are hashed into
                    // `far` and confirmed by a look at the earlier records
                    u64 near = 1ull << 32, far = 0;
                    for (u32 i = 1; i < len; ++i) {
                        const u32 g = sg[p + i];
                        const u32 d = g - g0 + 32u;
                        bool r;
                        if (d < 64u) { r = (near >> d) & 1ull; near |= 1ull << d; }
                        else {
                            const u64 bit = 1ull << covt_hash6(g);
                            r = false;
                            if (far & bit) for (u32 t = 1; t < i; ++t) r |= sg[p + t] == g;
                            far |= bit;
                        }
                        rep |= (u32)r << i;
                        multi |= g != g0;
                    }
                    const bool owned = p >= 32;                    // the read starts inside this tile: count it, emit its words
                    const u32 lenmask = len >= 32 ? FULL : (1u << len) - 1u, keep = ~rep & lenmask;
                    // the read's records from tile-relative record `at` on (a read reaching in from the left: its records inside the tile)
                    const u32 skip = owned ? 0u : 32u - p, at = owned ? p - 32u : 0u;
                    if (rep >> skip) covt_or_bits(rb, at, rep >> skip);
                    if (owned) {
                        if (multi) {
                            covt_or_bits(kb, at, keep);
                            cnt[at] = (unsigned char)__popc(keep);
                            // a chunk's last tile: what the read keeps behind the chunk is emitted right after pass 3 (the words of a read are contiguous)
                            if (tt + COVT_T == CHUNK && at + len > COVT_T) tail_keep = keep >> (COVT_T - at);
                        } else { ++uniq; atomicOr(ub + (at >> 5), 1u << (at & 31)); }   // all records name one reference: a unique read after all
                    } else if (multi && tt > 0) covt_or_bits(kb, 0, keep >> skip);   // headed in the previous tile of this chunk: its words continue here
                }
            }
            // ---- reads of more than 32 records ---------------------------------------------------------------------
            __syncwarp();
            for (u32 k = 0; k < n_long; ++k) {
                const u32 pp = llist[k];
                u32 s;
                if (pp != COVT_POS_UNKNOWN) s = t0 + pp - 32u;
                else {                                             // records t0 - 32 .. t0 all belong to it: walk further left
                    const u32 r0 = rec.read(t0);
                    u32 e = t0 - 32u;
                    for (;;) {
                        const bool eq = e > lane && rec.read(e - 1 - lane) == r0;
                        const u32 nb = ~__ballot_sync(FULL, eq);   // lanes whose record differs or does not exist
                        if (nb) { e -= (u32)__ffs(nb) - 1u; break; }
                        e -= 32u;
                    }
                    s = e;
                }
                covt_long_run<Rec, EXTRA>(rec, s, t0, t1, n, lane, P, rb, ub, &uniq, c * LR_SLOT, &n_lr);
            }
            __syncwarp();
            // ---- pass 3 -------------------------------------------------------------------------------------------
            if (rem >= COVT_T) covt_pass3<Rec, MODE, EXTRA, TEXG, true>(rec, P, t0, rem, lane, sg, hb, rb, ub, kb, cnt, s_cnt_addr, cw_c, cwi_c, rs_c, out, bad);
            else covt_pass3<Rec, MODE, EXTRA, TEXG, false>(rec, P, t0, rem, lane, sg, hb, rb, ub, kb, cnt, s_cnt_addr, cw_c, cwi_c, rs_c, out, bad);
            const u32 TB = __ballot_sync(FULL, tail_keep != 0);
            if (TB) {                                              // the chunk's last read goes on behind the chunk: its remaining words, from the halo
                const u32 tk = __shfl_sync(FULL, tail_keep, __ffs(TB) - 1);
                if ((tk >> lane) & 1u) {
                    const u32 o = out.n_cw + __popc(tk & LANE_LT(lane));
                    cw_c[o] = sg[32 + COVT_T + lane] < P.G ? sg[32 + COVT_T + lane] : 0u;
                    if (EXTRA && cwi_c) cwi_c[o] = t0 + COVT_T + lane;
                }
                out.n_cw += __popc(tk);
            }
        }
        if (lane == 0) P.chunk_cnt[c] = make_uint2(out.n_cw, min(n_lr, 0xFFu) | (out.n_rs << 8));
    }
    heads = warp_sum(heads); uniq = warp_sum(uniq); bad = warp_or(bad);
    if (lane == 0) { atomicAdd(&s_h, heads); atomicAdd(&s_u, uniq); if (bad) atomicOr(&s_b, bad); }
    __syncthreads();
    if (MODE)
        for (u32 b = tid; b < P.n_buckets; b += COVT_THREADS)
            if (s_cnt[b]) atomicAdd(P.bucket_cnt + b, s_cnt[b]);
    if (tid == 0) {
        if (s_h) atomicAdd(&P.sc->n_reads, (unsigned long long)s_h);
        if (s_u) atomicAdd(&P.sc->n_uniq, (unsigned long long)s_u);
        if (s_b) atomicOr(&P.sc->flags, s_b);
    }
}

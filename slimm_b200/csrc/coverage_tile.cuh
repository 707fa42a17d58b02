// coverage_tile.cuh - K1 coverage on warp-private tiles (included by kernels.cuh; same contract and outputs as k_coverage).
//
// Replaces reference src/slimm.hpp:194-257 + src/read_stat.hpp:116-135,72-75.
//
// k_coverage (the sliding 32-record window) spends ~260 thread instructions per record on run analysis that only the
// records of multi-record reads need.  Here a warp owns a tile of COVT_T consecutive records and works in two phases:
//
//   A  record-parallel, four consecutive records per lane (three 128-bit loads): head bits (a record starts a read) go
//      to a bit map in shared memory, reference ids are staged in shared memory, and EVERY record's contribution is
//      written at once under the assumption that holds for almost all of them - a read of one record is unique, any
//      other record is the first hit of its (read, reference) pair in a read with several targets.
//   B  read-parallel: the reads of two or more records are listed from the bit map (a fifth of the reads) and ONE LANE
//      walks one read over the staged reference ids: repeat hits (src/read_stat.hpp:125-131 keeps the first record of a
//      pair only) and reads whose records all name one reference are the exceptions; they PATCH what phase A wrote
//      (item -> ITEM_SKIP / unique; in direct mode a compensating 64-bit RED: the two packed counters are one integer
//      mod 2^64, so +1 followed by -1 is exact in any order).  The same walk emits the compact stream k_assign_reads
//      works on.  Reads longer than 32 records are walked by the whole warp.
//
// Ownership is positional: a tile writes the items of ITS records only.  A read that straddles a tile border is walked
// by both tiles (32 staged records of halo on either side; reading only), each patching its own records; the tile that
// holds the read's first record counts it and emits its compact words.  No tile ever writes into another tile's range,
// so there is no ordering between warps to get right.
#pragma once

#define COVT_T 1024u                      // records per tile
#define COVT_STEPS (COVT_T / 128u)        // phase A steps (4 records per lane)
#define COVT_WORDS (COVT_T / 32u)         // head-bit words of the tile proper
#define COVT_SG (COVT_T + 64u)            // staged reference ids: 32 halo | tile | 32 halo
#define COVT_HB (COVT_WORDS + 4u)         // head bits: halo word | tile | halo word | all-ones sentinel | pad
#define COVT_LIST (COVT_T / 2u + 2u)      // reads of two or more records headed in the tile, + the one reaching in from the left
#define COVT_LONG 36u                     // reads of more than 32 records that touch the tile (at most T/33 + 2)
#define COVT_WARP_WORDS 1400u             // COVT_SG + COVT_HB + COVT_LIST/2 + COVT_LONG/2, rounded up to 16 bytes
#define COVT_THREADS 256
#define COVT_POS_UNKNOWN 0xFFFFu          // long list: the read starts further left than the halo
static_assert(COVT_WORDS <= 32, "one head-bit word per lane");
static_assert(COVT_SG + COVT_HB + COVT_LIST / 2 + COVT_LONG / 2 <= COVT_WARP_WORDS && COVT_WARP_WORDS % 4 == 0, "per-warp shared memory layout");
static_assert(CHUNK % COVT_T == 0, "a chunk is a whole number of tiles");

struct Quad { uint4 r, g, p; };           // read id, reference id, position of four consecutive records

// records base .. base+3; `left` = records from base to the end of the data (the step is `full` when all 128 exist)
template <class Rec>
__device__ __forceinline__ Quad load_quad(const Rec &rec, u32 base, u32 left, bool full)
{
    Quad q;
    if (full) rec.load4(base, q.r, q.g, q.p);
    else {
        u32 r[4], g[4], p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool in = (u32)k < left;
            r[k] = in ? rec.read(base + k) : 0u; g[k] = in ? rec.refid(base + k) : 0u; p[k] = in ? rec.upos(base + k) : 0u;
        }
        q.r = make_uint4(r[0], r[1], r[2], r[3]); q.g = make_uint4(g[0], g[1], g[2], g[3]); q.p = make_uint4(p[0], p[1], p[2], p[3]);
    }
    return q;
}

__device__ __forceinline__ u32 covt_hash6(u32 g) { return (g * 0x9E3779B1u) >> 26; }

// phase B patch of one record: what phase A wrote for it is taken back (repeat hit) or upgraded to "unique read"
template <class Rec, int MODE>
__device__ __forceinline__ void covt_patch(const Rec &rec, const CovParams &P, u32 *s_cnt, u32 j, u32 g, bool to_unique)
{
    if (g >= P.G) return;                                          // phase A wrote ITEM_SKIP and raised the error flag
    const u64 b = bin_of(P.meta, g, rec.upos(j), P.half_avg, P.wdiv);
    if (MODE == 0) atomicAdd(P.hist + b, to_unique ? 0x100000000ull : 0xFFFFFFFFFFFFFFFFull);   // uniq_cov += 1 | cov -= 1
    else if (to_unique) __stcs(P.items + j, (u32)b | 0x80000000u);
    else { __stcs(P.items + j, ITEM_SKIP); atomicSub(&s_cnt[(u32)(b >> P.shift)], 1u); }
}

// A read of more than 32 records, walked by the whole warp on the records themselves: s = its first record (anywhere),
// [t0, t1) = the tile whose records are patched.  Returns nothing; uniq / n_lr are warp-uniform.
template <class Rec, int MODE, bool EXTRA>
__device__ __noinline__ void covt_long_run(const Rec &rec, u32 s, u32 t0, u32 t1, u32 n, u32 lane, const CovParams &P, u32 *s_cnt, u32 *uniq_io, u32 lr_base, u32 *n_lr_io)
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
    const bool owned = s >= t0;                                    // the read starts inside this tile
    if (owned) {
        u32 n_lr = *n_lr_io;
        if (lane == 0) {
            if (multi) { if (n_lr < LR_SLOT) P.lr[lr_base + n_lr] = s; }
            else if (EXTRA && P.res_kind) P.res_kind[s] = 3;
        }
        if (multi) *n_lr_io = n_lr + 1; else *uniq_io += (lane == 0);
    }
    const u32 a = max(s, t0), b = min(end, t1);
    for (u32 q = a; q < b; q += 32) {
        const u32 i = q + lane;
        if (i < b) {
            const u32 g = rec.refid(i);
            bool rep = i != s;                                     // one reference only: every record but the first repeats it
            if (multi) { rep = false; for (u32 j = s; j < i; ++j) if (rec.refid(j) == g) { rep = true; break; } }
            if (rep) covt_patch<Rec, MODE>(rec, P, s_cnt, i, g, false);
            else if (!multi) covt_patch<Rec, MODE>(rec, P, s_cnt, i, g, true);
        }
    }
}

// one more item of a histogram slice: a shared-memory RED on a shared-space address kept in a register
__device__ __forceinline__ void covt_count(u32 s_cnt_addr, u32 slice)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(s_cnt_addr + 4u * slice) : "memory");
}
#define COVT_DUMMY_SLICE MAX_BUCKETS      // counter of the records that contribute nothing (keeps the RED unconditional)

// the contributions of a lane's four records.  CHECK: some reference id may be out of range / some record may not exist
template <class Rec, int MODE, bool EXTRA, bool CHECK>
__device__ __forceinline__ void covt_emit4(const CovParams &P, const u32 (&g)[4], const u32 (&ps)[4], const bool (&in)[4], u32 singles /* bit k: record k is a read of its own */,
                                           u32 base, bool vec_store, u32 s_cnt_addr, u32 &bad)
{
    bool put[4];
    uint2 m2[4];
    uint4 m4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {                                  // four gathers in flight
        const bool ok = !CHECK || g[k] < P.G;
        put[k] = ok && (!CHECK || in[k]);
        if (CHECK && in[k] && !ok) bad |= 2u;
        const u32 gg = ok ? g[k] : 0u;
        if (MODE == 0) m4[k] = __ldg(P.meta + gg); else m2[k] = __ldg(P.meta2 + gg);
    }
    u32 item[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool single = (singles >> k) & 1u;                   // a read of one record: unique
        if (MODE == 0) {
            const u64 b = bin_of_meta(m4[k], ps[k], P.half_avg, P.wdiv);
            if (put[k]) atomicAdd(P.hist + b, single ? 0x100000001ull : 1ull);
        } else {                                                   // padded bin ids fit 31 bits on this path
            const u32 b = m2[k].y + fast_div(min(ps[k] + P.half_avg, m2[k].x), P.wdiv);
            item[k] = put[k] ? (b | (single ? 0x80000000u : 0u)) : ITEM_SKIP;
            covt_count(s_cnt_addr, put[k] ? b >> P.shift : COVT_DUMMY_SLICE);
        }
        if (EXTRA && P.res_kind && single && put[k]) P.res_kind[base + k] = 3;
    }
    if (MODE == 1) {
        if (vec_store) __stcs(reinterpret_cast<uint4 *>(P.items + base), make_uint4(item[0], item[1], item[2], item[3]));
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (in[k]) __stcs(P.items + base + k, item[k]);
        }
    }
}

// phase A, one step: 4 records per lane starting at record `base` (= tile + 128 s + 4 lane).  ALL128: all 128 records of the
// step exist (every step but the last of the data); otherwise `left` = records from base to the end of the data.
// r_after / has_after: the read id of the record right behind the step (warp-uniform), if there is one.
template <class Rec, int MODE, bool EXTRA, bool ALL128>
__device__ __forceinline__ void covt_step(const CovParams &P, const Quad &q, u32 base, u32 left, u32 lane, u32 &carry, bool first_record, u32 r_after,
                                          bool has_after, u32 *sg_step, u32 *hb_step, u32 s_cnt_addr, u32 &bad)
{
    const u32 r[4] = {q.r.x, q.r.y, q.r.z, q.r.w}, g[4] = {q.g.x, q.g.y, q.g.z, q.g.w}, ps[4] = {q.p.x, q.p.y, q.p.z, q.p.w};
    u32 prev = __shfl_up_sync(FULL, r[3], 1);
    if (lane == 0) prev = carry;
    carry = __shfl_sync(FULL, r[3], 31);
    bool in[4], h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) in[k] = ALL128 || (u32)k < left;
    // a record behind the end of the data reads as the head of a read: it ends the last real one
    const bool very_first = first_record && lane == 0;             // record 0 of the sample has no predecessor
    h[0] = !in[0] || r[0] != prev || very_first;
    bool down = in[0] && r[0] < prev && !very_first;               // read ids must be non-decreasing
#pragma unroll
    for (int k = 1; k < 4; ++k) { h[k] = !in[k] || r[k] != r[k - 1]; down |= in[k] && r[k] < r[k - 1]; }
    if (down) bad |= 1u;
    u32 nib = (u32)h[0] | ((u32)h[1] << 1) | ((u32)h[2] << 2) | ((u32)h[3] << 3);
    u32 hn = __shfl_down_sync(FULL, nib, 1) & 1u;                  // does the record behind my four start a read?
    if (lane == 31) hn = has_after ? (u32)(r_after != r[3]) : 1u;
    const u32 nib5 = nib | (hn << 4);
    const u32 singles = nib5 & (nib5 >> 1);                        // bit k: record k starts a read and so does record k + 1
    {   // head bits in record order: eight lanes make one word
        u32 v = nib << (4 * (lane & 7));
        v |= __shfl_xor_sync(FULL, v, 1); v |= __shfl_xor_sync(FULL, v, 2); v |= __shfl_xor_sync(FULL, v, 4);
        if ((lane & 7) == 0) hb_step[lane >> 3] = v;
    }
    *reinterpret_cast<uint4 *>(sg_step + 4 * lane) = q.g;
    if (ALL128 && max(max(g[0], g[1]), max(g[2], g[3])) < P.G)
        covt_emit4<Rec, MODE, EXTRA, false>(P, g, ps, in, singles, base, true, s_cnt_addr, bad);
    else
        covt_emit4<Rec, MODE, EXTRA, true>(P, g, ps, in, singles, base, ALL128, s_cnt_addr, bad);
}

// phase A over one tile.  WHOLE: all COVT_T records of the tile exist.  r_halo: read id of the record behind the tile
template <class Rec, int MODE, bool EXTRA, bool WHOLE>
__device__ __forceinline__ void covt_phase_a(const Rec &rec, const CovParams &P, u32 t0, u32 rem, u32 lane, u32 carry, u32 r_halo, u32 *sg, u32 *hb,
                                             u32 s_cnt_addr, u32 &bad)
{
    Quad cur = load_quad(rec, t0 + 4 * lane, rem > 4 * lane ? rem - 4 * lane : 0u, WHOLE || rem >= 128u);
    // read id of the first record behind a step (lane 31's last record needs it): fetched by lane 0 one step ahead of its use,
    // so that nothing waits on the quad that was only just requested
    u32 ra = (lane == 0 && (WHOLE || rem > 128u)) ? rec.read(t0 + 128u) : 0u;
#pragma unroll 2
    for (u32 s = 0; s < COVT_STEPS; ++s) {
        const u32 o = s * 128u;
        if (!WHOLE && o >= rem) {                                  // the whole step lies behind the data: heads only
            if ((lane & 7) == 0) hb[1 + 4 * s + (lane >> 3)] = FULL;
            continue;
        }
        const u32 left = WHOLE ? COVT_T - o : rem - o;             // records from this step on (at least: enough to tell a whole step)
        const bool has_after = WHOLE ? (s + 1 < COVT_STEPS || rem > COVT_T) : left > 128u;
        Quad nx = cur;
        u32 ra_next = 0;
        if (s + 1 < COVT_STEPS && has_after) {                     // the next step's loads fly during this step's work
            const u32 l2 = left - 128u;
            nx = load_quad(rec, t0 + o + 128u + 4 * lane, l2 > 4 * lane ? l2 - 4 * lane : 0u, WHOLE || l2 >= 128u);
            if (s + 2 < COVT_STEPS && lane == 0 && (WHOLE || l2 > 128u)) ra_next = rec.read(t0 + o + 256u);
        }
        const u32 r_after = s + 1 < COVT_STEPS ? __shfl_sync(FULL, ra, 0) : r_halo;
        const u32 mine = left > 4 * lane ? left - 4 * lane : 0u;
        if (WHOLE || left >= 128u)
            covt_step<Rec, MODE, EXTRA, true>(P, cur, t0 + o + 4 * lane, mine, lane, carry, t0 + o == 0, r_after, has_after, sg + 32 + o, hb + 1 + 4 * s, s_cnt_addr, bad);
        else
            covt_step<Rec, MODE, EXTRA, false>(P, cur, t0 + o + 4 * lane, mine, lane, carry, t0 + o == 0, r_after, has_after, sg + 32 + o, hb + 1 + 4 * s, s_cnt_addr, bad);
        cur = nx;
        ra = ra_next;
    }
}

template <class Rec, int MODE, bool EXTRA>
__global__ void __launch_bounds__(COVT_THREADS, 4)
k_coverage_tile(const __grid_constant__ Rec rec, u32 n, const __grid_constant__ CovParams P)
{
    extern __shared__ __align__(16) u32 covt_smem[];
    __shared__ u32 s_cnt[MODE ? MAX_BUCKETS + 1 : 1];              // items per histogram slice (shared-memory REDs) + the dummy slot
    __shared__ u32 s_h, s_u, s_b;
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    u32 *const sg = covt_smem + wid * COVT_WARP_WORDS;             // reference ids of records tile - 32 .. tile + T + 31
    u32 *const hb = sg + COVT_SG;                                  // head bits of the same records (+ sentinel word)
    unsigned short *const list = reinterpret_cast<unsigned short *>(hb + COVT_HB);
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
        u32 n_cw = 0, n_lr = 0, n_rs = 0;                          // compact words / long reads / multi-target reads of this chunk (warp-uniform)
        u32 *const cw_c = P.cw + (u64)c * CW_SLOT;
        u32 *const cwi_c = EXTRA && P.cw_idx ? P.cw_idx + (u64)c * CW_SLOT : nullptr;
        u32 *const rs_c = P.rs + (u64)c * RS_SLOT;
        const u32 c0 = c * CHUNK;
        for (u32 tt = 0; tt < CHUNK && n - c0 > tt; tt += COVT_T) {
            const u32 t0 = c0 + tt, rem = n - t0;                  // rem >= 1 records from t0 on
            const u32 nrec = min(rem, COVT_T), t1 = t0 + nrec;
            __syncwarp();                                          // the previous tile's phase B is done with the staging area
            // ---- halos: 32 records on either side, reference ids + head bits -------------------------------------
            u32 carry, r_halo;                                     // read ids of the records right before / right behind the tile
            {
                const bool vl = t0 + lane >= 32u;                  // record t0 - 32 + lane exists
                const u32 jl = t0 - 32u + lane;
                const u32 rl = vl ? rec.read(jl) : 0u;
                const u32 rlp = (vl && jl > 0) ? rec.read(jl - 1) : ~rl;
                sg[lane] = vl ? rec.refid(jl) : 0u;
                const u32 HBL = __ballot_sync(FULL, vl && rl != rlp);
                carry = __shfl_sync(FULL, rl, 31);                 // id of record t0 - 1 (unused when t0 == 0)
                const bool vr = rem > COVT_T + lane;               // record t0 + T + lane exists
                const u32 jr = t0 + COVT_T + lane;
                const u32 rr = vr ? rec.read(jr) : 0u, rrp = vr ? rec.read(jr - 1) : 1u;
                sg[32 + COVT_T + lane] = vr ? rec.refid(jr) : 0u;
                const u32 HBR = __ballot_sync(FULL, !vr || rr != rrp);
                r_halo = __shfl_sync(FULL, rr, 0);
                if (lane == 0) { hb[0] = HBL; hb[COVT_WORDS + 1] = HBR; hb[COVT_WORDS + 2] = FULL; }
            }
            // ---- phase A -----------------------------------------------------------------------------------------
            if (rem >= COVT_T) covt_phase_a<Rec, MODE, EXTRA, true>(rec, P, t0, rem, lane, carry, r_halo, sg, hb, s_cnt_addr, bad);
            else covt_phase_a<Rec, MODE, EXTRA, false>(rec, P, t0, rem, lane, carry, r_halo, sg, hb, s_cnt_addr, bad);
            __syncwarp();
            // ---- phase B: list the reads of two or more records --------------------------------------------------
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
                const u32 cnt = __popc(MH);
                u32 incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o) incl += y; }
                u32 off = incl - cnt;
                n_list = __shfl_sync(FULL, incl, 31);
                while (MH) { const u32 b = (u32)__ffs(MH) - 1u; MH &= MH - 1u; list[off++] = (unsigned short)(32u + 32u * lane + b); }
                const u32 HBL = hb[0];
                if (t0 > 0 && !(hb[1] & 1u)) {                     // record t0 continues a read that started further left
                    if (HBL) { if (lane == 0) list[n_list] = (unsigned short)(31 - __clz(HBL)); ++n_list; }
                    else { if (lane == 0) llist[0] = COVT_POS_UNKNOWN; n_long = 1; }
                }
            }
            __syncwarp();
            for (u32 q0 = 0; q0 < n_list; q0 += 32) {
                const u32 q = q0 + lane;
                const bool act = q < n_list;
                u32 p = 32, len = 1;                               // staged position of the read's first record, its records
                bool lng = false;
                if (act) {
                    p = list[q];
                    const u32 w = (p + 1) >> 5;
                    const u32 win = __funnelshift_r(hb[w], hb[w + 1], (p + 1) & 31);   // head bits of positions p+1 .. p+32
                    lng = win == 0;
                    len = lng ? 1u : (u32)__ffs(win);
                }
                const u32 LB = __ballot_sync(FULL, act && lng);    // more than 32 records: the whole warp, below
                if (LB) { if (act && lng) llist[n_long + __popc(LB & LANE_LT(lane))] = (unsigned short)p; n_long += __popc(LB); }
                const bool go = act && !lng;
                const bool owned = go && p >= 32;                  // the read starts inside this tile: count it, emit its words
                // room for the compact words is reserved BEFORE the walk, one word per record (an upper bound: repeat hits
                // leave gaps; every entry of rs carries its own word count), so the walk can write them as it goes
                u32 incl = owned ? len : 0u;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o) incl += y; }
                const u32 o_start = n_cw + incl - (owned ? len : 0u);
                n_cw += __shfl_sync(FULL, incl, 31);
                u32 rep = 0;                                       // bit i: record i of the read repeats an earlier reference of the read
                bool multi = false;
                u32 o = o_start;
                if (go) {
                    const u32 g0 = sg[p];
                    // references within +-32 of the first one have a bit of their own in `near` (exact); the others are hashed into
                    // `far` and confirmed by a look at the earlier records
                    u64 near = 1ull << 32, far = 0;
                    if (owned) { cw_c[o] = g0 | CW_HEAD; if (EXTRA && cwi_c) cwi_c[o] = t0 + p - 32u; ++o; }
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
                        if (owned && !r) { cw_c[o] = g; if (EXTRA && cwi_c) cwi_c[o] = t0 + p + i - 32u; ++o; }
                    }
                }
                const u32 EB = __ballot_sync(FULL, owned && multi);
                if (owned && multi) rs_c[n_rs + __popc(EB & LANE_LT(lane))] = o_start | ((o - o_start) << 16);
                n_rs += __popc(EB);
                if (go) {
                    const u32 jb = t0 + p - 32u;                   // record index of the read's first record
                    if (owned && !multi) {                         // all records name one reference: a unique read after all
                        ++uniq;
                        covt_patch<Rec, MODE>(rec, P, s_cnt, jb, sg[p], true);
                        if (EXTRA && P.res_kind && sg[p] < P.G) P.res_kind[jb] = 3;
                    }
                    // my records of the read: positions 32 .. 32 + nrec - 1
                    const u32 lo_i = p < 32 ? 32 - p : 0u, hi_i = min(len, 32 + nrec - p);
                    u32 pm = rep & (hi_i >= 32 ? FULL : (1u << hi_i) - 1u) & ~((1u << lo_i) - 1u);
                    while (pm) {
                        const u32 i = (u32)__ffs(pm) - 1u;
                        pm &= pm - 1u;
                        covt_patch<Rec, MODE>(rec, P, s_cnt, jb + i, sg[p + i], false);
                    }
                }
            }
            // ---- reads of more than 32 records -------------------------------------------------------------------
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
                covt_long_run<Rec, MODE, EXTRA>(rec, s, t0, t1, n, lane, P, s_cnt, &uniq, c * LR_SLOT, &n_lr);
            }
        }
        if (lane == 0) P.chunk_cnt[c] = make_uint2(n_cw, min(n_lr, 0xFFu) | (n_rs << 8));
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

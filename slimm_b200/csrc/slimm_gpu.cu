// slimm_gpu.cu - C ABI of the SLIMM profiling hot path (see include/slimm_gpu.h); the kernels live in
// kernels.cuh, the host tail (rank aggregation) in profile_host.cpp.  There is no CPU fallback.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <chrono>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/slimm_gpu.h"
#include "kernels.cuh"
#include "profile_host.h"

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
enum { ST_CREATED = 0, ST_COVERAGE = 1, ST_FILTER = 2, ST_ASSIGN = 3 };

struct slimm_gpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, aux_stream = nullptr;
    cudaEvent_t zero_start = nullptr, zero_done = nullptr, stats_ready = nullptr, fold_done = nullptr;
    u32 *d_cut_sorted = nullptr; float *d_cut_prefix = nullptr;   // K4: sorted cov_percent keys and their descending running sums, [2][G] each
    bool own_stream = true;
    cudaEvent_t upload_done = nullptr;
    u32 G = 0, w = 0, avg = 0, flags = 0, n_top = 0, npow2 = 1;
    int sm_count = 148;
    std::vector<u32> h_len, h_lin, h_top_vals, h_top_idx;
    std::vector<u64> h_off;                 // [G+1] padded bin offsets
    u64 Bp = 0, B = 0;
    // device
    cudaTextureObject_t meta2_tex = 0; int cov_gather = 0;   // 1: meta2 through the texture path in k_coverage_tile (default; SLIMM_COV_GATHER=ldg: plain loads)
    uint4 *d_meta = nullptr; uint2 *d_meta2 = nullptr; u32 *d_lin = nullptr; u32 *d_top_idx = nullptr; u64 *d_off = nullptr;
    unsigned long long *d_hist = nullptr; u32 *d_cov2 = nullptr;
    u32 *d_hist16 = nullptr; u64 hist16_cap = 0;   // compact bins {cov:16 | uniq_cov:16} of the fine slices with fewer than 65536 items (fine-slice runs)
    bool hist_compact = false, compact_bins = true;   // this run's bins are in the compact layout; SLIMM_GPU_COMPACT_BINS=0: always the interleaved 64-bit histogram
    u32 *d_stats = nullptr; float *d_cp = nullptr; u32 *d_scratch = nullptr;
    u32 *d_valid_bits = nullptr; unsigned char *d_valid_bytes = nullptr;
    u32 *d_assign = nullptr; u64 assign_words = 0;
    // bucketed scatter (histogram larger than L2)
    u32 *d_items = nullptr, *d_grouped = nullptr; u64 items_cap = 0; u32 bucket_shift = 22;
    Sched *d_sched = nullptr;
    u32 *d_cw = nullptr, *d_cw_idx = nullptr, *d_lr = nullptr; uint2 *d_chunk_cnt = nullptr; u64 cw_chunks = 0;   // compact stream for k_assign
    u32 *d_rs = nullptr;                    // per chunk: start | words << 16 of every multi-target read inside the chunk's compact words
    uint4 *d_lin16 = nullptr;               // lineages as 8 x 16-bit per-level dense taxon indices (fewer than 65536 references)
    uint4 *d_lin16v = nullptr;              // per sample: the same rows with the valid bit folded in (k_lin_valid); null: top level too wide
    BinDiv wdiv{0, 0, 0};
    int scatter_mode = -1;                  // -1 auto, 0 direct, 1 bucketed
    int cutoff_mode = -1;                   // -1 auto (cluster/DSMEM sort when it fits), 1 global-memory sort
    int cov_variant = 1;                    // 1: warp-private tiles (k_coverage_tile), 0: sliding 32-record windows (k_coverage)
    bool used_bucket = false;
    bool finished = false;                  // k_finish_assign ran
    std::unique_ptr<slimm_host::ProfilePlan> plan;
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
    float tail_host_ms = 0.0f;
    // rank reduction on the device (tree-consistent databases)
    u32 *d_lvl_idx = nullptr, *d_top_lvl7 = nullptr, *d_agg = nullptr; u32 *h_agg = nullptr; DevScalars *h_sc = nullptr;
    u32 shard_rank = 0, shard_n = 1;        // histogram slices sharded over ranks (slimm_gpu_set_shard)
    // peer-to-peer item exchange: every rank's receive buffer is mapped into every other rank (CUDA IPC over NVLink)
    u32 *d_recv = nullptr; u64 recv_cap = 0, n_recv = 0; std::vector<u32 *> peer_recv; bool p2p = false, split_pending = false;
    u32 **d_dest = nullptr; u32 **d_peer_recv = nullptr; u32 *d_n_recv = nullptr;   // d_n_recv[0]: items this rank receives, [1]: a receive buffer would overflow
    // routed exchange: a tile's share of a RANK travels as one segment, the receiver groups by slice (k_route, k_peer_route_plan)
    RoutePlan *d_route = nullptr; Sched *d_sched2 = nullptr; u32 *d_recv2 = nullptr; u64 recv2_cap = 0; bool routed = false; int route_mode = 2;   // 0: k_split<PEER> stores runs per slice, 1: k_route + receiver-side k_split, 2: local k_split + k_peer_copy of the owners' blocks
    CopyPlan *d_copy_plan = nullptr;
    bool n_recv_on_device = false;          // the split was planned on the device (slimm_gpu_split_to_peers_device): n_recv lives there
    // slimm_gpu_push_packed: staging of the wire format + the running read-id counter
    u32 *d_pk_bits = nullptr; unsigned short *d_pk_ref16 = nullptr; u32 *d_pk_tiles = nullptr, *d_pk_counter = nullptr; u64 pk_cap = 0;
    int push_kind = 0;                      // 0: nothing pushed yet, 1: slimm_gpu_push, 2: slimm_gpu_push_packed
    bool verify_pending = false;            // the optimistic checks of the coverage stage (ids non-decreasing, reference ids in range) have not been read back yet
    // fine slices: the histogram is accumulated in shared memory, 2^14 bins per CTA (k_fine_*)
    u32 *d_fine_cnt = nullptr, *d_fine_start = nullptr, *d_fine_cursor = nullptr, *d_fine = nullptr, *d_fine_ref = nullptr, *d_fine_hot = nullptr; u64 fine_slices_cap = 0, fine_cap = 0;
    int acc_mode = 1;                       // 1: fine slices in shared memory, 0: 64-bit REDs into L2-resident slices
    bool fine_packed = true;                // slices with fewer than 65536 items use 16+16-bit counters (64 KB per CTA)
    bool fine_cluster = true;               // slices with FINE_VHOT items or more are shared by a cluster of CTAs (SLIMM_GPU_FINE_CLUSTER=0: one CTA each)
    bool stats_done = false;                // the accumulate stage already reduced the per-reference statistics
    bool shard_acc_done = false;
    int tail_mode = -1;                     // -1 auto (device reduction when the database allows), 1 general host path
    std::vector<u32> h_assign;              // host copy of the assign block
    bool h_assign_ok = false;
    std::string err;
};

#define SLIMM_MAX_RECORDS 0xFFFFFF00ull

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

// {mul, s1, s2} with n / d == (t + ((n - t) >> s1)) >> s2, t = umulhi(mul, n), for every u32 n (round-up method)
static BinDiv bin_div_for(u32 d)
{
    if (d == 1) return BinDiv{0, 0, 0};
    u32 l = 0;
    while ((1ull << l) < d) ++l;                                   // ceil(log2 d)
    const u64 mul = ((1ull << 32) * ((1ull << l) - d)) / d + 1;
    return BinDiv{(u32)mul, 1, l - 1};
}

static int layout_bins(slimm_gpu_ctx *ctx)
{
    ctx->wdiv = bin_div_for(ctx->w);
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
        cudaFree(ctx->d_hist16); ctx->d_hist16 = nullptr; ctx->hist16_cap = 0;
        if (ctx->d_cov2) cudaFree(ctx->d_cov2);
        ctx->d_hist = nullptr; ctx->d_cov2 = nullptr;
        CU(cudaMalloc(&ctx->d_hist, std::max<u64>(Bp, 64) * 8));
        if (ctx->flags & SLIMM_GPU_KEEP_UNIQ_COV2) CU(cudaMalloc(&ctx->d_cov2, std::max<u64>(Bp, 64) * 4));
        ctx->Bp = Bp;
    }
    {   // the reference that holds the first bin of every block of FINE_REF_BLOCK bins: a warp of the accumulate kernels starts its segment
        // walk there (one look-up instead of a binary search over the references of the slice)
        const u64 n_blk = (Bp + FINE_REF_BLOCK - 1) / FINE_REF_BLOCK;
        std::vector<u32> fr(n_blk + 1);
        u32 g = 0;
        for (u64 k = 0; k <= n_blk; ++k) {
            const u64 bin = std::min(k * FINE_REF_BLOCK, Bp ? Bp - 1 : 0);
            while (g + 1 < G && ctx->h_off[g + 1] <= bin) ++g;
            fr[k] = g;
        }
        cudaFree(ctx->d_fine_ref); ctx->d_fine_ref = nullptr;
        CU(cudaMalloc(&ctx->d_fine_ref, (n_blk + 1) * 4));
        CU(cudaMemcpy(ctx->d_fine_ref, fr.data(), (n_blk + 1) * 4, cudaMemcpyHostToDevice));
    }
    CU(cudaMemcpy(ctx->d_meta, meta.data(), (size_t)G * sizeof(uint4), cudaMemcpyHostToDevice));
    {
        std::vector<uint2> meta2(G);
        for (u32 g = 0; g < G; ++g) meta2[g] = make_uint2(meta[g].x, meta[g].z);
        CU(cudaMemcpy(ctx->d_meta2, meta2.data(), (size_t)G * sizeof(uint2), cudaMemcpyHostToDevice));
    }
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
    CU(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&ctx->zero_start, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->zero_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->stats_ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ctx->fold_done, cudaEventDisableTiming));
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
    CU(cudaMalloc(&ctx->d_meta2, (size_t)G * sizeof(uint2)));
    {
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = ctx->d_meta2; rd.res.linear.desc = cudaCreateChannelDesc<uint2>(); rd.res.linear.sizeInBytes = (size_t)G * sizeof(uint2);
        cudaTextureDesc td{};
        td.readMode = cudaReadModeElementType;
        if (G <= (1u << 27)) CU(cudaCreateTextureObject(&ctx->meta2_tex, &rd, &td, nullptr));
    }
    ctx->cov_gather = ctx->meta2_tex ? 1 : 0;
    if (const char *e = getenv("SLIMM_COV_GATHER")) ctx->cov_gather = strcmp(e, "ldg") && ctx->meta2_tex ? 1 : 0;
    CU(cudaMalloc(&ctx->d_off, ((size_t)G + 1) * 8));
    CU(cudaMalloc(&ctx->d_lin, (size_t)G * 32));
    CU(cudaMalloc(&ctx->d_top_idx, (size_t)G * 4));
    CU(cudaMalloc(&ctx->d_stats, (size_t)G * 16));
    CU(cudaMalloc(&ctx->d_cp, (size_t)G * 8));
    CU(cudaMalloc(&ctx->d_scratch, (size_t)ctx->npow2 * 8));
    CU(cudaMalloc(&ctx->d_valid_bits, ((size_t)G + 31) / 32 * 4 + 4));
    CU(cudaMalloc(&ctx->d_valid_bytes, G));
    CU(cudaMalloc(&ctx->d_assign, ctx->assign_words * 4));
    CU(cudaMalloc(&ctx->d_sched, sizeof(Sched)));
    if (const char *e = getenv("SLIMM_GPU_TAIL")) ctx->tail_mode = !strcmp(e, "host") ? 1 : -1;
    if (const char *e = getenv("SLIMM_GPU_CUTOFF")) ctx->cutoff_mode = !strcmp(e, "global") ? 1 : -1;
    CU(cudaFuncSetAttribute(k_cut_sort_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, CUT_SHARE * 4));
    CU(cudaMalloc(&ctx->d_cut_sorted, (size_t)cfg->n_refs * 8));
    CU(cudaMalloc(&ctx->d_cut_prefix, (size_t)cfg->n_refs * 8));
    if (const char *e = getenv("SLIMM_GPU_ACC")) ctx->acc_mode = !strcmp(e, "l2") ? 0 : 1;
    if ((ctx->flags & SLIMM_GPU_SKIP_BINS) && (ctx->flags & SLIMM_GPU_KEEP_UNIQ_COV2)) return fail(ctx, SLIMM_GPU_EINVAL, "SLIMM_GPU_SKIP_BINS and SLIMM_GPU_KEEP_UNIQ_COV2 exclude each other");
    CU(cudaFuncSetAttribute(k_fine_accumulate<false, 1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * FINE_BINS * 4));
    CU(cudaFuncSetAttribute(k_fine_accumulate_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * FINE_BINS * 4));
    CU(cudaFuncSetAttribute(k_fine_accumulate<true, 512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FINE_BINS * 4));
    CU(cudaFuncSetAttribute(k_fine_accumulate<true, 512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FINE_BINS * 4));
    if (const char *e = getenv("SLIMM_GPU_FINE")) ctx->fine_packed = strcmp(e, "wide") != 0;
    if (const char *e = getenv("SLIMM_GPU_FINE_CLUSTER")) ctx->fine_cluster = atoi(e) != 0;
    if (const char *e = getenv("SLIMM_GPU_COMPACT_BINS")) ctx->compact_bins = atoi(e) != 0;
    if (const char *e = getenv("SLIMM_GPU_COV")) ctx->cov_variant = !strcmp(e, "window") ? 0 : 1;
    if (G < 65536) {
        // per level, the dense index of every distinct taxon id (zeros included): equal indices <=> equal ids
        std::vector<unsigned short> l16((size_t)G * 8);
        for (u32 l = 0; l < 8; ++l) {
            std::map<u32, u32> idx;
            for (u32 g = 0; g < G; ++g) idx.emplace(ctx->h_lin[(size_t)g * 8 + l], 0);
            u32 k = 0;
            for (auto &kv : idx) kv.second = k++;
            for (u32 g = 0; g < G; ++g) l16[(size_t)g * 8 + l] = (unsigned short)idx[ctx->h_lin[(size_t)g * 8 + l]];
            if (l == 7 && k <= 32768) CU(cudaMalloc(&ctx->d_lin16v, (size_t)G * 16));   // bit 15 of slot 7 is free for the valid flag
        }
        CU(cudaMalloc(&ctx->d_lin16, (size_t)G * 16));
        CU(cudaMemcpy(ctx->d_lin16, l16.data(), (size_t)G * 16, cudaMemcpyHostToDevice));
    }
    if (const char *e = getenv("SLIMM_GPU_SCATTER")) ctx->scatter_mode = !strcmp(e, "direct") ? 0 : !strcmp(e, "bucket") ? 1 : -1;
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
    if (ctx->meta2_tex) cudaDestroyTextureObject(ctx->meta2_tex);
    cudaFree(ctx->d_meta); cudaFree(ctx->d_meta2); cudaFree(ctx->d_off); cudaFree(ctx->d_lin); cudaFree(ctx->d_top_idx); cudaFree(ctx->d_hist); cudaFree(ctx->d_hist16);
    cudaFree(ctx->d_cov2); cudaFree(ctx->d_stats); cudaFree(ctx->d_cp); cudaFree(ctx->d_scratch); cudaFree(ctx->d_valid_bits);
    cudaFree(ctx->d_valid_bytes); cudaFree(ctx->d_assign); cudaFree(ctx->d_sc); cudaFree(ctx->d_tmp_bins);
    cudaFree(ctx->d_items); cudaFree(ctx->d_grouped); cudaFree(ctx->d_sched); cudaFree(ctx->d_lvl_idx); cudaFree(ctx->d_top_lvl7); cudaFree(ctx->d_agg); if (ctx->h_agg) cudaFreeHost(ctx->h_agg); if (ctx->h_sc) cudaFreeHost(ctx->h_sc); cudaFree(ctx->d_cw); cudaFree(ctx->d_cw_idx); cudaFree(ctx->d_lr); cudaFree(ctx->d_chunk_cnt);
    cudaFree(ctx->d_rs); cudaFree(ctx->d_lin16); cudaFree(ctx->d_lin16v); cudaFree(ctx->d_dest); cudaFree(ctx->d_peer_recv); cudaFree(ctx->d_n_recv);
    cudaFree(ctx->d_fine_cnt); cudaFree(ctx->d_fine_start); cudaFree(ctx->d_fine_cursor); cudaFree(ctx->d_fine); cudaFree(ctx->d_fine_ref); cudaFree(ctx->d_fine_hot); cudaFree(ctx->d_copy_plan);
    for (u32 q = 0; q < ctx->peer_recv.size(); ++q) if (ctx->peer_recv[q] && q != ctx->shard_rank) cudaIpcCloseMemHandle(ctx->peer_recv[q]);
    cudaFree(ctx->d_recv);
    cudaFree(ctx->d_rid_sorted); cudaFree(ctx->d_rp_sorted); cudaFree(ctx->d_kind); cudaFree(ctx->d_val);
    for (int i = 0; i < SLIMM_GPU_T_COUNT; ++i) { if (ctx->ev[i][0]) cudaEventDestroy(ctx->ev[i][0]); if (ctx->ev[i][1]) cudaEventDestroy(ctx->ev[i][1]); }
    if (ctx->upload_done) cudaEventDestroy(ctx->upload_done);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
    if (ctx->zero_start) cudaEventDestroy(ctx->zero_start);
    if (ctx->zero_done) cudaEventDestroy(ctx->zero_done);
    if (ctx->stats_ready) cudaEventDestroy(ctx->stats_ready);
    if (ctx->fold_done) cudaEventDestroy(ctx->fold_done);
    cudaFree(ctx->d_cut_sorted); cudaFree(ctx->d_cut_prefix);
    cudaFree(ctx->d_route); cudaFree(ctx->d_sched2); cudaFree(ctx->d_recv2);
    cudaFree(ctx->d_pk_bits); cudaFree(ctx->d_pk_ref16); cudaFree(ctx->d_pk_tiles); cudaFree(ctx->d_pk_counter);
    delete ctx;
    return SLIMM_GPU_OK;
}

int slimm_gpu_reset(slimm_gpu_ctx *ctx, uint32_t bin_width, uint32_t avg_read_length)
{
    if (!ctx) return SLIMM_GPU_EINVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->external) { ctx->d_rid = ctx->d_ref = nullptr; ctx->d_pos = nullptr; ctx->cap = 0; ctx->external = false; }
    ctx->n = 0; ctx->stage = ST_CREATED; ctx->use_sorted = false; ctx->have_global_hits = false; ctx->h_assign_ok = false; ctx->finished = false;
    ctx->push_kind = 0; ctx->verify_pending = false;
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
    if (ctx->n + n > SLIMM_MAX_RECORDS) return fail(ctx, SLIMM_GPU_ERANGE, "more than 2^32-256 records in one context");
    if (ctx->push_kind == 2) return fail(ctx, SLIMM_GPU_ESTATE, "slimm_gpu_push and slimm_gpu_push_packed cannot be mixed inside a sample");
    if (n == 0) return SLIMM_GPU_OK;
    ctx->push_kind = 1;
    CU(cudaSetDevice(ctx->device));
    int rc = reserve(ctx, ctx->n + n);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->d_rid + ctx->n, read_id, n * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(cudaMemcpyAsync(ctx->d_ref + ctx->n, ref_id, n * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(cudaMemcpyAsync(ctx->d_pos + ctx->n, begin_pos, n * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    ctx->n += n;
    return SLIMM_GPU_OK;
}

int slimm_gpu_push_packed(slimm_gpu_ctx *ctx, const uint32_t *new_read_bits, const uint16_t *ref_id16, const int32_t *begin_pos, uint64_t n)
{
    if (!ctx || (n && (!new_read_bits || !ref_id16 || !begin_pos))) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_CREATED) return fail(ctx, SLIMM_GPU_ESTATE, "push after coverage; call slimm_gpu_reset first");
    if (ctx->external) return fail(ctx, SLIMM_GPU_ESTATE, "push after push_device");
    if (ctx->push_kind == 1) return fail(ctx, SLIMM_GPU_ESTATE, "slimm_gpu_push and slimm_gpu_push_packed cannot be mixed inside a sample");
    if (ctx->G > 65536) return fail(ctx, SLIMM_GPU_EINVAL, "the packed wire format needs fewer than 65 537 contigs");
    if (ctx->n + n > SLIMM_MAX_RECORDS) return fail(ctx, SLIMM_GPU_ERANGE, "more than 2^32-256 records in one context");
    if (n == 0) return SLIMM_GPU_OK;
    CU(cudaSetDevice(ctx->device));
    int rc = reserve(ctx, ctx->n + n);
    if (rc) return rc;
    if (ctx->pk_cap < n) {                    // the copy stream orders a refill behind the kernels that read the staging area
        CU(cudaStreamSynchronize(ctx->copy_stream));
    cudaFree(ctx->d_pk_bits); cudaFree(ctx->d_pk_ref16); cudaFree(ctx->d_pk_tiles);
        ctx->d_pk_bits = nullptr; ctx->d_pk_ref16 = nullptr; ctx->d_pk_tiles = nullptr; ctx->pk_cap = 0;
        CU(cudaMalloc(&ctx->d_pk_bits, (n + 31) / 32 * 4 + 4)); CU(cudaMalloc(&ctx->d_pk_ref16, n * 2 + 2));
        CU(cudaMalloc(&ctx->d_pk_tiles, ((n + UNPACK_TILE - 1) / UNPACK_TILE + 1) * 4));
        ctx->pk_cap = n;
    }
    if (!ctx->d_pk_counter) { CU(cudaMalloc(&ctx->d_pk_counter, 4)); }
    if (ctx->push_kind == 0) CU(cudaMemsetAsync(ctx->d_pk_counter, 0, 4, ctx->copy_stream));   // a new sample: ids start at 0
    ctx->push_kind = 2;
    CU(cudaMemcpyAsync(ctx->d_pos + ctx->n, begin_pos, n * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(cudaMemcpyAsync(ctx->d_pk_ref16, ref_id16, n * 2, cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(cudaMemcpyAsync(ctx->d_pk_bits, new_read_bits, (n + 31) / 32 * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    const u64 n_tiles = (n + UNPACK_TILE - 1) / UNPACK_TILE;
    const unsigned grid = (unsigned)((n_tiles * 32 + 255) / 256);
    k_unpack_count<<<grid, 256, 0, ctx->copy_stream>>>(ctx->d_pk_bits, n, ctx->d_pk_tiles);
    k_unpack_scan<<<1, 1024, 0, ctx->copy_stream>>>(ctx->d_pk_tiles, n_tiles, ctx->d_pk_counter);
    k_unpack_write<<<grid, 256, 0, ctx->copy_stream>>>(ctx->d_pk_bits, ctx->d_pk_ref16, n, ctx->d_pk_tiles, ctx->d_rid + ctx->n, ctx->d_ref + ctx->n);
    ctx->launches += 3;
    CU(cudaGetLastError());
    ctx->n += n;
    return SLIMM_GPU_OK;
}

int slimm_gpu_push_device(slimm_gpu_ctx *ctx, const uint32_t *d_read_id, const uint32_t *d_ref_id, const int32_t *d_begin_pos, uint64_t n)
{
    if (!ctx || !d_read_id || !d_ref_id || !d_begin_pos) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_CREATED || ctx->n != 0) return fail(ctx, SLIMM_GPU_ESTATE, "push_device needs an empty context");
    if (n > SLIMM_MAX_RECORDS) return fail(ctx, SLIMM_GPU_ERANGE, "more than 2^32-256 records in one context");
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

}  // extern "C" (templates need C++ linkage)

// histogram slices of 2^22 bins (32 MB of interleaved u64) stay L2-resident while their items are applied
#define BUCKET_SHIFT 22

// Accumulate + per-reference statistics in shared memory, fine slice by fine slice.  items: grouped by coarse slice
// (their number is sd->total_items on the device, or n_given); n_cap bounds it on the host; out_buf receives the items
// grouped by fine slice; [lo_bin, hi_bin) are the bins this rank owns.
static bool force_exact();
static int fine_accumulate(slimm_gpu_ctx *ctx, const u32 *items, const u32 *n_ptr, u64 n_given, u64 n_cap, u32 *out_buf, u64 lo_bin, u64 hi_bin)
{
    const u64 n_fine = (ctx->Bp + FINE_BINS - 1) >> FINE_SHIFT;
    if (ctx->fine_slices_cap < n_fine) {
        cudaFree(ctx->d_fine_cnt); cudaFree(ctx->d_fine_start); cudaFree(ctx->d_fine_cursor);
        ctx->d_fine_cnt = ctx->d_fine_start = ctx->d_fine_cursor = nullptr; ctx->fine_slices_cap = 0;
        CU(cudaMalloc(&ctx->d_fine_cnt, (n_fine + 8) * 4)); CU(cudaMalloc(&ctx->d_fine_start, (n_fine + 1) * 4));
        CU(cudaMalloc(&ctx->d_fine_cursor, (n_fine + 1) * 4));
        ctx->fine_slices_cap = n_fine;
    }
    if (!ctx->d_fine_hot) CU(cudaMalloc(&ctx->d_fine_hot, 2 * 65536 * 4));   // slices with >= 65536 items: fewer than 2^16 of them; second half: the very hot ones
    TimeScope ts(ctx, SLIMM_GPU_T_ACCUM);
    CU(cudaMemsetAsync(ctx->d_fine_cnt, 0, (n_fine + 8) * 4, ctx->stream));
    const bool exact = n_fine > FINE_WRAP_SAFE || force_exact();               // (bins within 2^23 of 2^31: the kernels take the plain 32-bit slice difference)
    // spare words behind the counts (zeroed with them): ticket of the packed pass, number of hot slices, ticket of the hot pass, number of very hot slices
    u32 *ticket = ctx->d_fine_cnt + n_fine, *n_hot = ticket + 1, *hot_ticket = ticket + 2, *n_vhot = ticket + 3;
    u32 *vhot = ctx->fine_cluster && ctx->fine_packed ? ctx->d_fine_hot + 65536 : nullptr;
    CU(cudaMemsetAsync(ctx->d_stats, 0, (size_t)ctx->G * 16, ctx->stream));
    const u64 n_tiles = (n_cap + FINE_TILE - 1) / FINE_TILE;
    if (n_tiles) {
        const int cgrid = (int)std::max<u64>(1, std::min<u64>(n_tiles, (u64)ctx->sm_count * 8));
        if (exact) k_fine_count<true><<<cgrid, 256, 0, ctx->stream>>>(items, n_ptr, (u32)n_given, ctx->bucket_shift, ctx->d_fine_cnt);
        else k_fine_count<false><<<cgrid, 256, 0, ctx->stream>>>(items, n_ptr, (u32)n_given, ctx->bucket_shift, ctx->d_fine_cnt);
    }
    k_fine_scan<<<FINE_SCAN_CL, 1024, 0, ctx->stream>>>(ctx->d_fine_cnt, (u32)n_fine, ctx->d_fine_start, ctx->d_fine_cursor, ctx->d_fine_hot, n_hot, vhot, n_vhot);
    if (n_tiles) {
        const int fnt = getenv("SLIMM_FINE_SPLIT_NT") ? atoi(getenv("SLIMM_FINE_SPLIT_NT")) : 256;   // tile shape (experiments)
        if (fnt >= 512) {
            const u64 tiles = (n_cap + 512 * FINE_ITEMS - 1) / (512 * FINE_ITEMS);
            const int sgrid = (int)std::max<u64>(1, std::min<u64>(tiles, (u64)ctx->sm_count * 2));
            k_fine_split<512, FINE_ITEMS><<<sgrid, 512, 0, ctx->stream>>>(items, n_ptr, (u32)n_given, ctx->bucket_shift, ctx->d_fine_cursor, out_buf);
        } else {
            const int sgrid = (int)std::max<u64>(1, std::min<u64>(n_tiles, (u64)ctx->sm_count * FINE_SPLIT_OCC));
            k_fine_split<256, FINE_ITEMS><<<sgrid, 256, 0, ctx->stream>>>(items, n_ptr, (u32)n_given, ctx->bucket_shift, ctx->d_fine_cursor, out_buf);
        }
    }
    const u64 f_lo = lo_bin >> FINE_SHIFT, f_hi = (std::min(hi_bin, ctx->Bp) + FINE_BINS - 1) >> FINE_SHIFT;
    if (f_hi > f_lo) {
        uint4 *hist4 = (ctx->flags & SLIMM_GPU_SKIP_BINS) ? nullptr : (uint4 *)ctx->d_hist;
        // compact bins: the slices of the packed pass keep their 16+16-bit words (4 instead of 8 bytes per bin written); the hot slices
        // go to the interleaved 64-bit histogram as before; readers tell the two by the slice's item count (k_extract_bins_compact)
        u32 *hist16 = nullptr;
        ctx->hist_compact = false;
        if (hist4 && ctx->fine_packed && ctx->compact_bins) {
            if (ctx->hist16_cap < ctx->Bp) {
                cudaFree(ctx->d_hist16); ctx->d_hist16 = nullptr; ctx->hist16_cap = 0;
                CU(cudaMalloc(&ctx->d_hist16, std::max<u64>(ctx->Bp, 64) * 4));
                ctx->hist16_cap = ctx->Bp;
            }
            hist16 = ctx->d_hist16;
            ctx->hist_compact = true;
        }
        if (ctx->fine_packed) {
            // slices with fewer than 65536 items (all but the hottest): packed counters, two CTAs per SM; then the rest, wide
            const unsigned grid2 = (unsigned)std::min<u64>(f_hi - f_lo, (u64)ctx->sm_count * FINE_PACKED_CTAS);
            if (hist16)
                k_fine_accumulate<true, 512, true><<<grid2, 512, FINE_BINS * 4, ctx->stream>>>(out_buf, ctx->d_fine_start, (u32)f_lo, (u32)f_hi, ctx->Bp, ctx->d_off,
                                                                                            ctx->G, ctx->d_fine_ref, ctx->d_stats, (uint4 *)hist16, ticket, 0u, 65536u, nullptr, nullptr);
            else
                k_fine_accumulate<true, 512, false><<<grid2, 512, FINE_BINS * 4, ctx->stream>>>(out_buf, ctx->d_fine_start, (u32)f_lo, (u32)f_hi, ctx->Bp, ctx->d_off,
                                                                                             ctx->G, ctx->d_fine_ref, ctx->d_stats, hist4, ticket, 0u, 65536u, nullptr, nullptr);
            // the hot slices k_fine_scan listed: wide counters, one CTA per slice (list positions handed out by a ticket); the very hot
            // ones: a cluster of CTAs per slice
            const unsigned grid1 = (unsigned)std::min<u64>(f_hi - f_lo, (u64)ctx->sm_count);
            k_fine_accumulate<false, 1024, false><<<grid1, 1024, 2 * FINE_BINS * 4, ctx->stream>>>(out_buf, ctx->d_fine_start, (u32)f_lo, (u32)f_hi, ctx->Bp, ctx->d_off,
                                                                                     ctx->G, ctx->d_fine_ref, ctx->d_stats, hist4, hot_ticket, 65536u, 0xFFFFFFFFu, ctx->d_fine_hot, n_hot);
            ctx->launches++;
            if (vhot) {
                const unsigned n_cl = std::max(1u, (unsigned)ctx->sm_count / FINE_CL - 2u);
                k_fine_accumulate_cluster<<<n_cl * FINE_CL, 1024, 2 * FINE_BINS * 4, ctx->stream>>>(out_buf, ctx->d_fine_start, ctx->Bp, ctx->d_off, ctx->G, ctx->d_fine_ref,
                                                                                                 ctx->d_stats, hist4, vhot, n_vhot);
                ctx->launches++;
            }
        } else {
            const unsigned grid = (unsigned)std::min<u64>(f_hi - f_lo, (u64)ctx->sm_count);
            k_fine_accumulate<false, 1024, false><<<grid, 1024, 2 * FINE_BINS * 4, ctx->stream>>>(out_buf, ctx->d_fine_start, (u32)f_lo, (u32)f_hi, ctx->Bp, ctx->d_off,
                                                                                    ctx->G, ctx->d_fine_ref, ctx->d_stats, hist4, ticket, 0u, 0xFFFFFFFFu, nullptr, nullptr);
        }
    }
    ctx->launches += 4;
    CU(cudaGetLastError());
    ctx->stats_done = true;
    return SLIMM_GPU_OK;
}

// ITEM_SKIP is ranked in the slot min(0x7FFFFFFF >> shift, MAX_BUCKETS); when the slices in use reach that slot the kernels test for it explicitly
// (SLIMM_FORCE_EXACT=1 takes the EXACT instantiations regardless: they are otherwise only reached with more than 2.13e9 bins)
static bool force_exact() { const char *e = getenv("SLIMM_FORCE_EXACT"); return e && atoi(e) != 0; }
static bool split_needs_exact(u32 shift, u32 n_buckets) { return force_exact() || std::min<u32>(0x7FFFFFFFu >> shift, MAX_BUCKETS) < n_buckets; }

// K1b: shape of the split CTAs (SLIMM_SPLIT_NT = 256 | 512 | 1024 threads over the same 8192-item tile)
template <bool PEER>
static void launch_k_split(slimm_gpu_ctx *ctx, int sgrid, const u32 *items, u32 n, u32 shift, u32 n_buckets, u32 *out, u32 *const *dest)
{
    const int nt = getenv("SLIMM_SPLIT_NT") ? atoi(getenv("SLIMM_SPLIT_NT")) : 512;
    const int per_sm = nt >= 1024 ? 1 : 2;
    const u64 n_tiles = ((u64)n + SPLIT_TILE - 1) / SPLIT_TILE;
    const int grid = (int)std::max<u64>(1, std::min<u64>(n_tiles, (u64)ctx->sm_count * per_sm));
    (void)sgrid;
    const int bulk = getenv("SLIMM_SPLIT_BULK") ? atoi(getenv("SLIMM_SPLIT_BULK")) : 0;   // tiles through the bulk-copy engine (local splits)
    if (!PEER && bulk && ((uintptr_t)items & 15) == 0 && n_buckets <= SPLITB_NT) {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(k_split_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, SPLITB_SMEM); attr = true; }
        const int bgrid = (int)std::max<u64>(1, std::min<u64>(n_tiles, (u64)ctx->sm_count * 2));
        k_split_bulk<<<bgrid, SPLITB_NT, SPLITB_SMEM, ctx->stream>>>(items, n, shift, n_buckets, ctx->d_sched, out, nullptr);
        return;
    }
    const bool exact = split_needs_exact(shift, n_buckets);
#define SLIMM_SPLIT_LAUNCH(NT_)                                                                                                                   \
    do {                                                                                                                                          \
        if (exact) k_split<PEER, NT_, true><<<grid, NT_, 0, ctx->stream>>>(items, n, shift, n_buckets, ctx->d_sched, out, dest, nullptr);           \
        else k_split<PEER, NT_, false><<<grid, NT_, 0, ctx->stream>>>(items, n, shift, n_buckets, ctx->d_sched, out, dest, nullptr);                \
    } while (0)
    if (nt == 256) SLIMM_SPLIT_LAUNCH(256);
    else if (nt >= 1024) SLIMM_SPLIT_LAUNCH(1024);
    else SLIMM_SPLIT_LAUNCH(512);
#undef SLIMM_SPLIT_LAUNCH
}

static bool aligned16(const RecSoA &r) { return ((((uintptr_t)r.rid) | ((uintptr_t)r.ref) | ((uintptr_t)r.pos)) & 15u) == 0; }
static bool aligned16(const RecPacked &r) { return ((((uintptr_t)r.rid) | ((uintptr_t)r.rp)) & 15u) == 0; }

// K1: the tile kernel (128-bit loads: the arrays must be 16-byte aligned), else the sliding windows
template <class Rec, int MODE>
static void launch_k_coverage(slimm_gpu_ctx *ctx, const Rec &rec, u32 n, const CovParams &P, bool extra, int grid)
{
    if (ctx->cov_variant == 1 && aligned16(rec)) {
        const u64 n_chunks = ((u64)n + CHUNK - 1) / CHUNK;
        const int warps = COVT_THREADS / 32;
        const int tgrid = (int)std::max<u64>(1, std::min<u64>((n_chunks + warps - 1) / warps, (u64)ctx->sm_count * 4));
        const size_t dyn = (size_t)warps * COVT_WARP_WORDS * 4;
        if (extra) k_coverage_tile<Rec, MODE, true, false><<<tgrid, COVT_THREADS, dyn, ctx->stream>>>(rec, n, P);
        else if (ctx->cov_gather == 1 && MODE == 1) k_coverage_tile<Rec, MODE, false, true><<<tgrid, COVT_THREADS, dyn, ctx->stream>>>(rec, n, P);
        else k_coverage_tile<Rec, MODE, false, false><<<tgrid, COVT_THREADS, dyn, ctx->stream>>>(rec, n, P);
    } else if (extra) k_coverage<Rec, MODE, true><<<grid, 256, 0, ctx->stream>>>(rec, n, P);
    else k_coverage<Rec, MODE, false><<<grid, 256, 0, ctx->stream>>>(rec, n, P);
}

template <class Rec>
static int launch_coverage_t(slimm_gpu_ctx *ctx, Rec rec)
{
    const u32 n = (u32)ctx->n;
    const u64 n_chunks = ((u64)n + CHUNK - 1) / CHUNK;
    static const int cov_ctas = getenv("SLIMM_COV_CTAS") ? atoi(getenv("SLIMM_COV_CTAS")) : 6;   // CTAs per SM (experiments)
    const int grid = (int)std::max<u64>(1, std::min<u64>((n_chunks + 7) / 8, (u64)ctx->sm_count * cov_ctas));
    // compact stream of the multi-mapped reads (k_assign's input), one slot per chunk
    const bool want_idx = (ctx->flags & (SLIMM_GPU_KEEP_UNIQ_COV2 | SLIMM_GPU_READ_RESULTS)) != 0;
    if (ctx->cw_chunks < n_chunks) {
        cudaFree(ctx->d_cw); cudaFree(ctx->d_cw_idx); cudaFree(ctx->d_lr); cudaFree(ctx->d_chunk_cnt); cudaFree(ctx->d_rs);
        ctx->d_cw = ctx->d_cw_idx = ctx->d_lr = nullptr; ctx->d_chunk_cnt = nullptr; ctx->d_rs = nullptr; ctx->cw_chunks = 0;
        CU(cudaMalloc(&ctx->d_rs, n_chunks * RS_SLOT * 4));
        CU(cudaMalloc(&ctx->d_cw, n_chunks * CW_SLOT * 4));
        if (want_idx) CU(cudaMalloc(&ctx->d_cw_idx, n_chunks * CW_SLOT * 4));
        CU(cudaMalloc(&ctx->d_lr, n_chunks * LR_SLOT * 4));
        CU(cudaMalloc(&ctx->d_chunk_cnt, n_chunks * sizeof(uint2)));
        ctx->cw_chunks = n_chunks;
    }
    if ((ctx->flags & SLIMM_GPU_READ_RESULTS) && ctx->res_cap < ctx->n) {
        cudaFree(ctx->d_kind); cudaFree(ctx->d_val);
        ctx->d_kind = nullptr; ctx->d_val = nullptr;
        CU(cudaMalloc(&ctx->d_kind, std::max<u64>(ctx->n, 1))); CU(cudaMalloc(&ctx->d_val, std::max<u64>(ctx->n, 1) * 4));
        ctx->res_cap = ctx->n;
    }
    if (ctx->d_kind) CU(cudaMemsetAsync(ctx->d_kind, 0, std::max<u64>(ctx->n, 1), ctx->stream));
    CovParams P{};
    P.meta = ctx->d_meta; P.meta2 = ctx->d_meta2; P.meta2_tex = ctx->meta2_tex; P.G = ctx->G; P.half_avg = ctx->avg / 2u; P.wdiv = ctx->wdiv; P.hist = ctx->d_hist;
    P.cw = ctx->d_cw; P.cw_idx = ctx->d_cw_idx; P.chunk_cnt = ctx->d_chunk_cnt; P.lr = ctx->d_lr; P.rs = ctx->d_rs;
    P.res_kind = (ctx->flags & SLIMM_GPU_READ_RESULTS) ? ctx->d_kind : nullptr; P.sc = ctx->d_sc;
    if (!ctx->used_bucket) {
        TimeScope ts(ctx, SLIMM_GPU_T_COVERAGE);
        launch_k_coverage<Rec, 0>(ctx, rec, n, P, want_idx || P.res_kind, grid);
        ctx->launches++;
        CU(cudaGetLastError());
        return SLIMM_GPU_OK;
    }
    const u32 shift = ctx->bucket_shift;
    const u32 n_buckets = (u32)((ctx->Bp + (1ull << shift) - 1) >> shift);
    if (ctx->items_cap < n) {
        cudaFree(ctx->d_items); cudaFree(ctx->d_grouped); ctx->d_items = nullptr; ctx->d_grouped = nullptr;
        CU(cudaMalloc(&ctx->d_items, (((u64)n + 3) & ~3ull) * 4));
        CU(cudaMalloc(&ctx->d_grouped, (((u64)n + 3) & ~3ull) * 4));
        ctx->items_cap = n;
    }
    {
        TimeScope ts(ctx, SLIMM_GPU_T_COVERAGE);
        CU(cudaMemsetAsync(ctx->d_sched, 0, sizeof(Sched), ctx->stream));
        P.items = ctx->d_items; P.shift = shift; P.n_buckets = n_buckets;
        P.bucket_cnt = reinterpret_cast<u32 *>(reinterpret_cast<char *>(ctx->d_sched) + offsetof(Sched, count));
        launch_k_coverage<Rec, 1>(ctx, rec, n, P, want_idx || P.res_kind, grid);
        ctx->launches++;
    }
    {
        TimeScope ts(ctx, SLIMM_GPU_T_SPLIT);   // the multisplit: slice starts + unit schedule, then group the items by slice
        k_bucket_scan<<<1, MAX_BUCKETS, 0, ctx->stream>>>(ctx->d_sched, n_buckets);
        const u64 n_tiles = ((u64)n + SPLIT_TILE - 1) / SPLIT_TILE;
        const int sgrid = (int)std::max<u64>(1, std::min<u64>(n_tiles, (u64)ctx->sm_count * 4));
        if (ctx->shard_n > 1 && ctx->p2p) { ctx->split_pending = true; ctx->launches += 1; }   // slimm_gpu_split_to_peers splits straight into the owners' buffers
        else {
            launch_k_split<false>(ctx, sgrid, ctx->d_items, n, shift, n_buckets, ctx->d_grouped, nullptr);
            ctx->launches += 2;
        }
    }
    if (ctx->shard_n > 1) { CU(cudaGetLastError()); return SLIMM_GPU_OK; }   // the caller exchanges the items, then slimm_gpu_accumulate_items
    if (ctx->acc_mode == 1) return fine_accumulate(ctx, ctx->d_grouped, &ctx->d_sched->total_items, 0, n, ctx->d_items, 0, ctx->Bp);
    {
        CU(cudaStreamWaitEvent(ctx->stream, ctx->zero_done, 0));   // the histogram was zero-filled on the side stream meanwhile
        TimeScope ts(ctx, SLIMM_GPU_T_ACCUM);
        {
            // Measured on B200 (cfg5): the tighter the window of items in flight, the better the REDs hit L2 - 512 items
            // per block 8.7 ms, 4096 12.0 ms, 65536 30.5 ms, grid-stride 28-34 ms.  SLIMM_ACC_SHAPE overrides (experiments).
            static const char *shape = getenv("SLIMM_ACC_SHAPE");
            const u32 per_block = shape && !strncmp(shape, "stride", 6) ? 0u : shape ? (u32)atoi(shape) : 512u;
            const unsigned g = per_block ? (unsigned)std::max<u64>(1, ((u64)n + per_block - 1) / per_block)
                                         : (unsigned)(ctx->sm_count * (shape && !strcmp(shape, "stride8") ? 8 : 16));
            k_accumulate<<<g, 256, 0, ctx->stream>>>(ctx->d_grouped, ctx->d_sched, ctx->d_hist, per_block, 0);
        }
        ctx->launches++;
    }
    CU(cudaGetLastError());
    return SLIMM_GPU_OK;
}

extern "C" {

static int accumulate_owned(slimm_gpu_ctx *ctx, const u32 *d_items, u64 n_items);
static u32 n_slices_of(const slimm_gpu_ctx *ctx) { return (u32)((ctx->Bp + (1ull << ctx->bucket_shift) - 1) >> ctx->bucket_shift); }

// slices [lo, hi) of the histogram owned by rank r of n: contiguous, as even as possible
static void owned_slices(const slimm_gpu_ctx *ctx, u32 r, u32 *lo, u32 *hi)
{
    const u64 ns = n_slices_of(ctx);
    *lo = (u32)(ns * r / ctx->shard_n); *hi = (u32)(ns * (r + 1) / ctx->shard_n);
}

static void owned_bins(const slimm_gpu_ctx *ctx, u64 *lo_bin, u64 *hi_bin)
{
    u32 lo, hi;
    owned_slices(ctx, ctx->shard_rank, &lo, &hi);
    *lo_bin = std::min<u64>(ctx->Bp, (u64)lo << ctx->bucket_shift);
    *hi_bin = std::min<u64>(ctx->Bp, (u64)hi << ctx->bucket_shift);
}

static void choose_scatter(slimm_gpu_ctx *ctx)
{
    // bucketed scatter when the interleaved histogram is much larger than L2 (and bin ids fit 31 bits)
    const bool big = ctx->Bp * 8 > (96ull << 20) && ctx->n >= (1u << 20);
    ctx->bucket_shift = BUCKET_SHIFT;
    if (const char *e = getenv("SLIMM_BUCKET_SHIFT")) ctx->bucket_shift = (u32)std::max(16, std::min(28, atoi(e)));   // experiments
    while (((ctx->Bp + (1ull << ctx->bucket_shift) - 1) >> ctx->bucket_shift) > MAX_BUCKETS) ++ctx->bucket_shift;
    ctx->used_bucket = ctx->n > 0 && ctx->Bp < 0x7FFFFFFFull && (ctx->scatter_mode == 1 || (ctx->scatter_mode == -1 && big));
    if (ctx->shard_n > 1) ctx->used_bucket = true;   // the items are routed to the ranks that own their slices
}

static int launch_coverage(slimm_gpu_ctx *ctx)
{
    if (ctx->use_sorted) return launch_coverage_t(ctx, RecPacked{ctx->d_rid_sorted, ctx->d_rp_sorted});
    return launch_coverage_t(ctx, RecSoA{ctx->d_rid, ctx->d_ref, ctx->d_pos});
}

static int zero_state(slimm_gpu_ctx *ctx)
{
    choose_scatter(ctx);
    if (!ctx->used_bucket) {
        TimeScope ts(ctx, SLIMM_GPU_T_ZERO);
        CU(cudaMemsetAsync(ctx->d_hist, 0, std::max<u64>(ctx->Bp, 64) * 8, ctx->stream));
    } else if (ctx->acc_mode == 1) {
        // fine slices: every CTA of k_fine_accumulate writes its whole slice - nothing to zero-fill
    } else {   // the big histogram is only needed by k_accumulate: zero-fill it on the side stream, under k_coverage / k_split
        CU(cudaEventRecord(ctx->zero_start, ctx->stream));
        CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->zero_start, 0));
        if (ctx->timing) { cudaEventRecord(ctx->ev[SLIMM_GPU_T_ZERO][0], ctx->aux_stream); ctx->ev_used[SLIMM_GPU_T_ZERO] = true; }
        u64 lo_bin = 0, hi_bin = std::max<u64>(ctx->Bp, 64);
        if (ctx->shard_n > 1) owned_bins(ctx, &lo_bin, &hi_bin);
        if (hi_bin > lo_bin) CU(cudaMemsetAsync(ctx->d_hist + lo_bin, 0, (hi_bin - lo_bin) * 8, ctx->aux_stream));
        if (ctx->timing) cudaEventRecord(ctx->ev[SLIMM_GPU_T_ZERO][1], ctx->aux_stream);
        CU(cudaEventRecord(ctx->zero_done, ctx->aux_stream));
    }
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

// Reads the flags the coverage stage left behind.  Input that was not grouped by read: the records are sorted on the device and
// every stage that already ran is run again (the first pass was wasted, never wrong: nothing of it survives).
static int verify_input(slimm_gpu_ctx *ctx)
{
    if (!ctx->verify_pending) return SLIMM_GPU_OK;
    ctx->verify_pending = false;
    CU(cudaSetDevice(ctx->device));
    for (int pass = 0; pass < 2; ++pass) {
        u32 flags = 0;
        CU(cudaMemcpyAsync(&flags, &ctx->d_sc->flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (flags & 2u) { ctx->stage = ST_CREATED; return fail(ctx, SLIMM_GPU_EINVAL, "a record references a contig id >= n_refs"); }
        if (!(flags & 1u)) return SLIMM_GPU_OK;
        if (pass == 1) return fail(ctx, SLIMM_GPU_ECUDA, "read ids still out of order after the device sort");
        if (ctx->shard_n > 1) { ctx->stage = ST_CREATED; return fail(ctx, SLIMM_GPU_EINVAL, "a sharded run needs every rank's records grouped by read"); }
        const int stage = ctx->stage;
        ctx->was_sorted = false;
        int rc = sort_records(ctx); if (rc) return rc;
        ctx->stats_done = false; ctx->h_assign_ok = false; ctx->finished = false;
        rc = zero_state(ctx); if (rc) return rc;
        rc = launch_coverage(ctx); if (rc) return rc;
        ctx->stage = ST_COVERAGE;
        if (stage >= ST_FILTER) { rc = slimm_gpu_filter(ctx, ctx->q, ctx->min_reads_opt); if (rc) return rc; }
        if (stage >= ST_ASSIGN) { rc = slimm_gpu_assign(ctx); if (rc) return rc; }
    }
    return SLIMM_GPU_OK;
}

int slimm_gpu_coverage(slimm_gpu_ctx *ctx)
{
    if (!ctx) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_CREATED) return fail(ctx, SLIMM_GPU_ESTATE, "coverage already ran; call slimm_gpu_reset for a new sample");
    CU(cudaSetDevice(ctx->device));
    ctx->hist_compact = false;                                  // set again by the fine-slice accumulate when it keeps compact bins
    // kernels wait for the uploads without blocking the host
    CU(cudaEventRecord(ctx->upload_done, ctx->copy_stream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->upload_done, 0));
    for (int i = 0; i < SLIMM_GPU_T_COUNT; ++i) ctx->ev_used[i] = false;
    ctx->use_sorted = false; ctx->was_sorted = true; ctx->h_assign_ok = false; ctx->finished = false; ctx->shard_acc_done = false; ctx->stats_done = false;
    ctx->verify_pending = false; ctx->n_recv_on_device = false;
    int rc = zero_state(ctx);
    if (rc) return rc;
    if (ctx->n) {
        rc = launch_coverage(ctx);
        if (rc) return rc;
        // optimistic: the kernel assumes non-decreasing read ids and reference ids in range and verifies both on the fly; the flags
        // are read back with the first result (verify_input), so the stages queue up behind each other without a host round trip
        ctx->verify_pending = true;
    }
    else if (ctx->shard_n > 1) CU(cudaMemsetAsync(ctx->d_sched, 0, sizeof(Sched), ctx->stream));   // no records on this rank: no items
    ctx->stage = ST_COVERAGE;
    return SLIMM_GPU_OK;
}

int slimm_gpu_bins_device(slimm_gpu_ctx *ctx, void **d_ptr, uint64_t *n_u32)
{
    if (!ctx || !d_ptr || !n_u32) return SLIMM_GPU_EINVAL;
    if (ctx->flags & SLIMM_GPU_SKIP_BINS) return fail(ctx, SLIMM_GPU_EINVAL, "the bins are not kept (SLIMM_GPU_SKIP_BINS)");
    if (ctx->hist_compact) return fail(ctx, SLIMM_GPU_ESTATE, "the bins of this run are kept in the compact per-slice layout: use slimm_gpu_fetch_bins (or SLIMM_GPU_COMPACT_BINS=0)");
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

int slimm_gpu_set_shard(slimm_gpu_ctx *ctx, uint32_t rank, uint32_t n_ranks)
{
    if (!ctx || n_ranks == 0 || rank >= n_ranks) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_CREATED) return fail(ctx, SLIMM_GPU_ESTATE, "set_shard must come before coverage");
    if (n_ranks > 1 && (ctx->flags & SLIMM_GPU_KEEP_UNIQ_COV2)) return fail(ctx, SLIMM_GPU_EINVAL, "uniq_cov2 bins are not available in sharded runs");
    if (n_ranks > 1 && ctx->Bp >= 0x7FFFFFFFull) return fail(ctx, SLIMM_GPU_EINVAL, "sharded runs need fewer than 2^31 padded bins");
    ctx->shard_rank = rank; ctx->shard_n = n_ranks;
    return SLIMM_GPU_OK;
}

int slimm_gpu_get_slice_counts(slimm_gpu_ctx *ctx, uint32_t *counts, uint32_t cap, uint32_t *n_slices)
{
    if (!ctx || !n_slices) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_COVERAGE || !ctx->used_bucket) return fail(ctx, SLIMM_GPU_ESTATE, "slice counts exist after a bucketed coverage stage");
    CU(cudaSetDevice(ctx->device));
    if (counts) { int rc = verify_input(ctx); if (rc) return rc; }
    const u32 ns = n_slices_of(ctx);
    *n_slices = ns;
    if (counts) {
        if (cap < ns) return fail(ctx, SLIMM_GPU_EINVAL, "counts buffer too small");
        CU(cudaMemcpyAsync(counts, reinterpret_cast<char *>(ctx->d_sched) + offsetof(Sched, count), (size_t)ns * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return SLIMM_GPU_OK;
}

int slimm_gpu_slice_counts_device(slimm_gpu_ctx *ctx, void **d_counts, uint32_t *n_slices)
{
    if (!ctx || !d_counts || !n_slices) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_COVERAGE || !ctx->used_bucket) return fail(ctx, SLIMM_GPU_ESTATE, "slice counts exist after a bucketed coverage stage");
    *d_counts = reinterpret_cast<char *>(ctx->d_sched) + offsetof(Sched, count);
    *n_slices = n_slices_of(ctx);
    return SLIMM_GPU_OK;
}

int slimm_gpu_items_device(slimm_gpu_ctx *ctx, void **d_items)
{
    if (!ctx || !d_items) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_COVERAGE || !ctx->used_bucket) return fail(ctx, SLIMM_GPU_ESTATE, "grouped items exist after a bucketed coverage stage");
    *d_items = ctx->d_grouped;
    return SLIMM_GPU_OK;
}

int slimm_gpu_accumulate_items(slimm_gpu_ctx *ctx, const uint32_t *d_items, uint64_t n_items)
{
    if (!ctx || (n_items && !d_items)) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_COVERAGE || ctx->shard_n < 2) return fail(ctx, SLIMM_GPU_ESTATE, "accumulate_items belongs to a sharded run, after coverage");
    if (n_items > SLIMM_MAX_RECORDS) return fail(ctx, SLIMM_GPU_ERANGE, "too many items");
    return accumulate_owned(ctx, d_items, n_items);
}

static int accumulate_owned(slimm_gpu_ctx *ctx, const u32 *d_items, u64 n_items)
{
    CU(cudaSetDevice(ctx->device));
    u64 lo_bin = 0, hi_bin = 0;
    owned_bins(ctx, &lo_bin, &hi_bin);
    if (ctx->acc_mode == 1) {
        if (ctx->fine_cap < n_items) {
            cudaFree(ctx->d_fine); ctx->d_fine = nullptr; ctx->fine_cap = 0;
            CU(cudaMalloc(&ctx->d_fine, std::max<u64>(n_items + n_items / 8, 1024) * 4));
            ctx->fine_cap = n_items + n_items / 8;
        }
        int rc = fine_accumulate(ctx, d_items, ctx->n_recv_on_device && (d_items == ctx->d_recv || d_items == ctx->d_recv2) ? ctx->d_n_recv : nullptr, n_items, n_items,
                                 ctx->d_fine, lo_bin, hi_bin);
        if (rc) return rc;
        ctx->shard_acc_done = true;
        return SLIMM_GPU_OK;
    }
    CU(cudaStreamWaitEvent(ctx->stream, ctx->zero_done, 0));    // the owned bins were zero-filled on the side stream
    if (n_items) {
        TimeScope ts(ctx, SLIMM_GPU_T_ACCUM);
        const u32 per_block = 512;
        k_accumulate<<<(unsigned)((n_items + per_block - 1) / per_block), 256, 0, ctx->stream>>>(d_items, nullptr, ctx->d_hist, per_block, (u32)n_items);
        ctx->launches++;
    }
    {
        TimeScope ts(ctx, SLIMM_GPU_T_STATS);   // partial statistics over the owned bins; the caller sums slimm_gpu_stats_device over ranks
        CU(cudaMemsetAsync(ctx->d_stats, 0, (size_t)ctx->G * 16, ctx->stream));
        const u64 s_lo = lo_bin / 64, s_hi = hi_bin / 64;
        if (s_hi > s_lo) {
            const u64 n_chunks = (s_hi - s_lo + STATS_STEPS_PER_WARP - 1) / STATS_STEPS_PER_WARP;
            k_ref_stats<<<grid_for(ctx, n_chunks * 32, 256, 8), 256, 0, ctx->stream>>>((const uint4 *)ctx->d_hist, s_lo, s_hi, ctx->d_off, ctx->G, ctx->d_stats);
            ctx->launches++;
        }
    }
    CU(cudaGetLastError());
    ctx->shard_acc_done = true;
    return SLIMM_GPU_OK;
}

// ---- peer-to-peer item exchange (one process per GPU on one NVLink/NVSwitch box) ----------------------------------
int slimm_gpu_p2p_reserve(slimm_gpu_ctx *ctx, uint64_t cap_items, void *ipc_handle_64)
{
    if (!ctx || !ipc_handle_64 || cap_items == 0) return SLIMM_GPU_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaSetDevice(ctx->device));
    if (ctx->recv_cap < cap_items) {
        if (ctx->p2p) return fail(ctx, SLIMM_GPU_ESTATE, "the receive buffer cannot grow once peers have mapped it");
        cudaFree(ctx->d_recv); ctx->d_recv = nullptr; ctx->recv_cap = 0;
        CU(cudaMalloc(&ctx->d_recv, cap_items * 4));
        ctx->recv_cap = cap_items;
    }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->d_recv));
    memcpy(ipc_handle_64, &h, 64);
    return SLIMM_GPU_OK;
}

int slimm_gpu_p2p_connect(slimm_gpu_ctx *ctx, const void *ipc_handles, uint32_t n_ranks)
{
    if (!ctx || !ipc_handles) return SLIMM_GPU_EINVAL;
    if (n_ranks != ctx->shard_n || n_ranks < 2) return fail(ctx, SLIMM_GPU_ESTATE, "p2p_connect: call slimm_gpu_set_shard first, with the same number of ranks");
    if (!ctx->d_recv) return fail(ctx, SLIMM_GPU_ESTATE, "p2p_connect: call slimm_gpu_p2p_reserve first");
    CU(cudaSetDevice(ctx->device));
    ctx->peer_recv.assign(n_ranks, nullptr);
    for (u32 q = 0; q < n_ranks; ++q) {
        if (q == ctx->shard_rank) { ctx->peer_recv[q] = ctx->d_recv; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)ipc_handles + (size_t)q * 64, 64);
        void *p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_recv[q] = (u32 *)p;
    }
    if (!ctx->d_dest) CU(cudaMalloc(&ctx->d_dest, MAX_BUCKETS * sizeof(u32 *)));
    if (!ctx->d_n_recv) CU(cudaMalloc(&ctx->d_n_recv, 8));
    if (!ctx->d_route) { CU(cudaMalloc(&ctx->d_route, sizeof(RoutePlan))); CU(cudaMalloc(&ctx->d_sched2, sizeof(Sched))); }
    if (const char *e = getenv("SLIMM_PEER_ROUTE")) ctx->route_mode = std::max(0, std::min(2, atoi(e)));
    if (!ctx->d_copy_plan) CU(cudaMalloc(&ctx->d_copy_plan, sizeof(CopyPlan)));
    cudaFree(ctx->d_peer_recv); ctx->d_peer_recv = nullptr;
    CU(cudaMalloc(&ctx->d_peer_recv, n_ranks * sizeof(u32 *)));
    CU(cudaMemcpy(ctx->d_peer_recv, ctx->peer_recv.data(), n_ranks * sizeof(u32 *), cudaMemcpyHostToDevice));
    ctx->p2p = true;
    return SLIMM_GPU_OK;
}

int slimm_gpu_p2p_disable(slimm_gpu_ctx *ctx)
{
    if (!ctx) return SLIMM_GPU_EINVAL;
    if (ctx->split_pending) return fail(ctx, SLIMM_GPU_ESTATE, "p2p_disable between coverage and split_to_peers");
    ctx->p2p = false;   // the mappings stay open until destroy; the exchange goes back to slimm_gpu_items_device + accumulate_items
    return SLIMM_GPU_OK;
}

int slimm_gpu_split_to_peers(slimm_gpu_ctx *ctx, const uint32_t *all_counts, uint64_t *n_recv)
{
    if (!ctx || !all_counts) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_COVERAGE || !ctx->p2p) return fail(ctx, SLIMM_GPU_ESTATE, "split_to_peers belongs to a peer-to-peer sharded run, after coverage");
    CU(cudaSetDevice(ctx->device));
    const u32 ns = n_slices_of(ctx), nr = ctx->shard_n, me = ctx->shard_rank;
    // receive buffer of owner q: its slices in ascending order, inside a slice the source ranks in ascending order
    std::vector<u32 *> dest(MAX_BUCKETS, nullptr);
    u64 mine = 0;
    for (u32 q = 0; q < nr; ++q) {
        u32 lo, hi;
        owned_slices(ctx, q, &lo, &hi);
        u64 off = 0;
        for (u32 s = lo; s < hi; ++s)
            for (u32 src = 0; src < nr; ++src) {
                if (src == me) dest[s] = ctx->peer_recv[q] + off;
                off += all_counts[(size_t)src * ns + s];
            }
        if (off > ctx->recv_cap) return fail(ctx, SLIMM_GPU_ERANGE, "a rank would receive more items than slimm_gpu_p2p_reserve reserved");
        if (q == me) mine = off;
    }
    ctx->n_recv = mine;
    ctx->n_recv_on_device = false; ctx->routed = false;         // (a context may have used the device-planned exchange for its last sample)
    if (n_recv) *n_recv = mine;
    if (ctx->split_pending) {
        TimeScope ts(ctx, SLIMM_GPU_T_SPLIT);
        CU(cudaMemcpyAsync(ctx->d_dest, dest.data(), MAX_BUCKETS * sizeof(u32 *), cudaMemcpyHostToDevice, ctx->stream));
        const u32 n = (u32)ctx->n;
        const u64 n_tiles = ((u64)n + SPLIT_TILE - 1) / SPLIT_TILE;
        const int sgrid = (int)std::max<u64>(1, std::min<u64>(n_tiles, (u64)ctx->sm_count * 4));
        launch_k_split<true>(ctx, sgrid, ctx->d_items, n, ctx->bucket_shift, ns, nullptr, ctx->d_dest);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(ctx->stream));   // dest is a host vector; and the caller's barrier comes next anyway
        ctx->split_pending = false;
    }
    return SLIMM_GPU_OK;
}

int slimm_gpu_split_to_peers_device(slimm_gpu_ctx *ctx, const uint32_t *d_all_counts)
{
    if (!ctx || !d_all_counts) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_COVERAGE || !ctx->p2p) return fail(ctx, SLIMM_GPU_ESTATE, "split_to_peers belongs to a peer-to-peer sharded run, after coverage");
    CU(cudaSetDevice(ctx->device));
    const u32 ns = n_slices_of(ctx);
    TimeScope ts(ctx, SLIMM_GPU_T_SPLIT);
    CU(cudaMemsetAsync(ctx->d_n_recv, 0, 8, ctx->stream));
    ctx->routed = ctx->route_mode == 1 && ctx->shard_n <= ROUTE_MAX_RANKS && ctx->acc_mode == 1;
    const bool blocks = ctx->route_mode == 2 && ctx->shard_n <= ROUTE_MAX_RANKS && ctx->acc_mode == 1;
    if (blocks) {
        // group by slice locally (as one GPU does), then every owner's block travels as one contiguous copy
        k_peer_copy_plan<<<1, MAX_BUCKETS, 0, ctx->stream>>>(d_all_counts, ns, ctx->shard_n, ctx->shard_rank, ctx->d_peer_recv, ctx->recv_cap, ctx->d_copy_plan,
                                                             ctx->d_n_recv, ctx->d_n_recv + 1);
        ctx->launches++;
        ctx->n_recv_on_device = true;
        ctx->n_recv = ctx->recv_cap;
        if (ctx->split_pending) {
            const u32 n = (u32)ctx->n;
            launch_k_split<false>(ctx, 0, ctx->d_items, n, ctx->bucket_shift, ns, ctx->d_grouped, nullptr);
            const u64 chunks = ((u64)n + PCOPY_CHUNK - 1) / PCOPY_CHUNK + ctx->shard_n;
            const int grid = (int)std::max<u64>(1, std::min<u64>(chunks, (u64)ctx->sm_count * 4));
            k_peer_copy<<<grid, 256, 0, ctx->stream>>>(ctx->d_grouped, ctx->d_copy_plan, ctx->shard_n, ctx->shard_rank, ctx->d_n_recv + 1);
            ctx->launches += 2;
            ctx->split_pending = false;
        }
        CU(cudaGetLastError());
        return SLIMM_GPU_OK;
    }
    if (ctx->routed) {
        if (ctx->recv2_cap < ctx->recv_cap) {
            cudaFree(ctx->d_recv2); ctx->d_recv2 = nullptr; ctx->recv2_cap = 0;
            CU(cudaMalloc(&ctx->d_recv2, ctx->recv_cap * 4));
            ctx->recv2_cap = ctx->recv_cap;
        }
        k_peer_route_plan<<<1, MAX_BUCKETS, 0, ctx->stream>>>(d_all_counts, ns, ctx->shard_n, ctx->shard_rank, ctx->d_peer_recv, ctx->recv_cap, ctx->d_route,
                                                              ctx->d_sched2, ctx->d_n_recv, ctx->d_n_recv + 1);
    } else
        k_peer_dest<<<1, MAX_BUCKETS, 0, ctx->stream>>>(d_all_counts, ns, ctx->shard_n, ctx->shard_rank, ctx->d_peer_recv, ctx->recv_cap, ctx->d_dest, ctx->d_n_recv,
                                                        ctx->d_n_recv + 1);
    ctx->launches++;
    ctx->n_recv_on_device = true;
    ctx->n_recv = ctx->recv_cap;              // an upper bound for the host side (buffers, grids); the kernels read the count on the device
    if (ctx->split_pending) {
        const u32 n = (u32)ctx->n;
        if (ctx->routed) {
            const u64 n_tiles = ((u64)n + SPLIT_TILE - 1) / SPLIT_TILE;
            const int grid = (int)std::max<u64>(1, std::min<u64>(n_tiles, (u64)ctx->sm_count * 2));
            k_route<512><<<grid, 512, 0, ctx->stream>>>(ctx->d_items, n, ctx->bucket_shift, ctx->shard_n, ctx->d_route);
        } else launch_k_split<true>(ctx, 0, ctx->d_items, n, ctx->bucket_shift, ns, nullptr, ctx->d_dest);
        ctx->launches++;
        ctx->split_pending = false;
    }
    CU(cudaGetLastError());
    return SLIMM_GPU_OK;
}

// element-wise sum over the contexts of a device buffer (u32 words), through the host: small buffers only
static int sum_over_contexts(slimm_gpu_ctx **ctxs, u32 n, u32 *(*get)(slimm_gpu_ctx *), u64 words)
{
    std::vector<u32> acc(words, 0), one(words);
    for (u32 r = 0; r < n; ++r) {
        slimm_gpu_ctx *ctx = ctxs[r];
        CU(cudaSetDevice(ctx->device));
        CU(cudaMemcpyAsync(one.data(), get(ctx), words * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        for (u64 i = 0; i < words; ++i) acc[i] += one[i];
    }
    for (u32 r = 0; r < n; ++r) {
        slimm_gpu_ctx *ctx = ctxs[r];
        CU(cudaSetDevice(ctx->device));
        CU(cudaMemcpyAsync(get(ctx), acc.data(), words * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return SLIMM_GPU_OK;
}

int slimm_gpu_run_sharded_local(slimm_gpu_ctx **ctxs, uint32_t n, float cov_cut_off, uint32_t min_reads, uint64_t global_hits)
{
    if (!ctxs || n == 0) return SLIMM_GPU_EINVAL;
    for (u32 r = 0; r < n; ++r)
        if (!ctxs[r] || ctxs[r]->shard_n != n || ctxs[r]->shard_rank != r || ctxs[r]->G != ctxs[0]->G || ctxs[r]->Bp != ctxs[0]->Bp)
            return fail(ctxs[r], SLIMM_GPU_ESTATE, "run_sharded_local: every context needs slimm_gpu_set_shard(r, n) and the same contigs");
    if (n == 1) return slimm_gpu_run(ctxs[0], cov_cut_off, min_reads);
    int rc;
    for (u32 r = 0; r < n; ++r) {                                  // coverage on every device, back to back (asynchronous)
        ctxs[r]->p2p = false;
        if ((rc = slimm_gpu_coverage(ctxs[r]))) return rc;
    }
    const u32 ns = n_slices_of(ctxs[0]);
    std::vector<u32> counts((size_t)n * ns);
    for (u32 r = 0; r < n; ++r) {
        u32 got = 0;
        if ((rc = slimm_gpu_get_slice_counts(ctxs[r], counts.data() + (size_t)r * ns, ns, &got))) return rc;
    }
    // owner q receives, source by source, the block of each source's slice-grouped items that holds q's slices
    for (u32 q = 0; q < n; ++q) {
        slimm_gpu_ctx *ctx = ctxs[q];
        u32 lo, hi;
        owned_slices(ctx, q, &lo, &hi);
        u64 total = 0;
        for (u32 src = 0; src < n; ++src) for (u32 s = lo; s < hi; ++s) total += counts[(size_t)src * ns + s];
        CU(cudaSetDevice(ctx->device));
        if (ctx->recv_cap < total) {
            cudaFree(ctx->d_recv); ctx->d_recv = nullptr; ctx->recv_cap = 0;
            CU(cudaMalloc(&ctx->d_recv, std::max<u64>(total, 1) * 4));
            ctx->recv_cap = total;
        }
        u64 off = 0;
        for (u32 src = 0; src < n; ++src) {
            u64 before = 0, mine = 0;
            for (u32 s = 0; s < lo; ++s) before += counts[(size_t)src * ns + s];
            for (u32 s = lo; s < hi; ++s) mine += counts[(size_t)src * ns + s];
            if (mine) CU(cudaMemcpyPeerAsync(ctx->d_recv + off, ctx->device, ctxs[src]->d_grouped + before, ctxs[src]->device, mine * 4, ctx->stream));
            off += mine;
        }
        ctx->n_recv = total;
    }
    for (u32 q = 0; q < n; ++q) {                                  // the copies read every source's buffer: all of them done before anybody goes on
        slimm_gpu_ctx *ctx = ctxs[q];
        CU(cudaSetDevice(ctx->device));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    for (u32 q = 0; q < n; ++q)
        if ((rc = slimm_gpu_accumulate_items(ctxs[q], ctxs[q]->d_recv, ctxs[q]->n_recv))) return rc;
    if ((rc = sum_over_contexts(ctxs, n, [](slimm_gpu_ctx *c) { return c->d_stats; }, (u64)ctxs[0]->G * 4))) return rc;
    if ((rc = sum_over_contexts(ctxs, n, [](slimm_gpu_ctx *c) { return reinterpret_cast<u32 *>(&c->d_sc->n_reads); }, 4))) return rc;   // n_reads, n_uniq (u64 each; partial sums stay below 2^32)
    for (u32 r = 0; r < n; ++r) {
        if ((rc = slimm_gpu_set_global_hits(ctxs[r], global_hits))) return rc;
        if ((rc = slimm_gpu_filter(ctxs[r], cov_cut_off, min_reads))) return rc;
        if ((rc = slimm_gpu_assign(ctxs[r]))) return rc;
    }
    return sum_over_contexts(ctxs, n, [](slimm_gpu_ctx *c) { return c->d_assign; }, ctxs[0]->assign_words);
}

int slimm_gpu_accumulate_received(slimm_gpu_ctx *ctx)
{
    if (!ctx) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_COVERAGE || !ctx->p2p) return fail(ctx, SLIMM_GPU_ESTATE, "accumulate_received belongs to a peer-to-peer sharded run, after split_to_peers");
    if (ctx->routed && ctx->n_recv_on_device) {
        // what arrived is grouped by source only: group it by slice (the slice totals came with the plan), then go on as usual
        CU(cudaSetDevice(ctx->device));
        const u64 n_tiles = (ctx->recv_cap + SPLIT_TILE - 1) / SPLIT_TILE;
        const int grid = (int)std::max<u64>(1, std::min<u64>(n_tiles, (u64)ctx->sm_count * 2));
        {
            TimeScope ts(ctx, SLIMM_GPU_T_SORT);   // (the sort slot of the timings is free in a sharded run: grouped input is required)
            if (split_needs_exact(ctx->bucket_shift, n_slices_of(ctx)))
                k_split<false, 512, true><<<grid, 512, 0, ctx->stream>>>(ctx->d_recv, 0, ctx->bucket_shift, n_slices_of(ctx), ctx->d_sched2, ctx->d_recv2, nullptr, ctx->d_n_recv);
            else
                k_split<false, 512, false><<<grid, 512, 0, ctx->stream>>>(ctx->d_recv, 0, ctx->bucket_shift, n_slices_of(ctx), ctx->d_sched2, ctx->d_recv2, nullptr, ctx->d_n_recv);
            ctx->launches++;
        }
        CU(cudaGetLastError());
        return accumulate_owned(ctx, ctx->d_recv2, ctx->n_recv);
    }
    return accumulate_owned(ctx, ctx->d_recv, ctx->n_recv);
}

int slimm_gpu_stats_device(slimm_gpu_ctx *ctx, void **d_ptr, uint64_t *n_u32)
{
    if (!ctx || !d_ptr || !n_u32) return SLIMM_GPU_EINVAL;
    *d_ptr = ctx->d_stats; *n_u32 = (u64)ctx->G * 4;
    return SLIMM_GPU_OK;
}

int slimm_gpu_filter(slimm_gpu_ctx *ctx, float cov_cut_off, uint32_t min_reads)
{
    if (!ctx) return SLIMM_GPU_EINVAL;
    if (ctx->stage != ST_COVERAGE) return fail(ctx, SLIMM_GPU_ESTATE, "filter needs coverage first");
    CU(cudaSetDevice(ctx->device));
    ctx->q = cov_cut_off; ctx->min_reads_opt = min_reads;
    if (ctx->shard_n > 1 && !ctx->shard_acc_done) return fail(ctx, SLIMM_GPU_ESTATE, "sharded run: slimm_gpu_accumulate_items must come before filter");
    if (ctx->shard_n == 1 && !ctx->stats_done) {
        TimeScope ts(ctx, SLIMM_GPU_T_STATS);
        CU(cudaMemsetAsync(ctx->d_stats, 0, (size_t)ctx->G * 16, ctx->stream));
        const u64 n_steps = ctx->Bp / 64;
        const u64 n_chunks = (n_steps + STATS_STEPS_PER_WARP - 1) / STATS_STEPS_PER_WARP;
        const int grid = grid_for(ctx, n_chunks * 32, 256, 8);
        if (n_steps) k_ref_stats<<<grid, 256, 0, ctx->stream>>>((const uint4 *)ctx->d_hist, 0, n_steps, ctx->d_off, ctx->G, ctx->d_stats);
        ctx->launches += 1;
        CU(cudaGetLastError());
    }
    {
        TimeScope ts(ctx, SLIMM_GPU_T_CUTOFF);
        if (ctx->G <= (u32)CUT_CL * CUT_SHARE && ctx->cutoff_mode != 1) {
            // the in-order total on the side stream, under the cluster sort + running sums; then the parallel search (kernels.cuh, K4)
            CU(cudaEventRecord(ctx->stats_ready, ctx->stream));
            CU(cudaStreamWaitEvent(ctx->aux_stream, ctx->stats_ready, 0));
            k_cut_fold<<<2, 1024, 0, ctx->aux_stream>>>(ctx->d_stats, ctx->d_meta, ctx->G, ctx->d_sc);
            CU(cudaEventRecord(ctx->fold_done, ctx->aux_stream));
            u32 m = CUT_CL * 1024;
            while (m < ctx->G) m <<= 1;
            const size_t dyn = (size_t)(m / CUT_CL) * 4;
            k_cut_sort_cluster<<<2 * CUT_CL, 1024, dyn, ctx->stream>>>(ctx->d_stats, ctx->d_meta, ctx->G, cov_cut_off, ctx->d_cp, ctx->d_cut_sorted, ctx->d_cut_prefix, ctx->d_sc);
            CU(cudaStreamWaitEvent(ctx->stream, ctx->fold_done, 0));
            k_cut_finish<<<(ctx->G + 1023) / 1024, 1024, 0, ctx->stream>>>(ctx->d_stats, ctx->G, cov_cut_off, min_reads, ctx->d_cp, ctx->d_cut_sorted, ctx->d_cut_prefix,
                                                      ctx->d_valid_bits, ctx->d_valid_bytes, ctx->d_sc);
            ctx->launches += 2;
        } else {
            k_cutoffs<<<2, 1024, 0, ctx->stream>>>(ctx->d_stats, ctx->d_meta, ctx->G, cov_cut_off, min_reads, ctx->d_cp, ctx->d_scratch,
                                                   ctx->npow2, ctx->d_valid_bits, ctx->d_valid_bytes, ctx->d_sc);
        }
        ctx->launches++;
        if (ctx->d_cov2 && ctx->Bp) {   // uniq_cov2 starts as uniq_cov of the surviving references
            const u64 n_steps = ctx->Bp / 64;
            k_cov2_base<<<grid_for(ctx, n_steps * 32, 256, 8), 256, 0, ctx->stream>>>((const uint4 *)ctx->d_hist, n_steps, ctx->d_off, ctx->G,
                                                                                     ctx->d_valid_bits, (uint2 *)ctx->d_cov2,
                                                                                     ctx->hist_compact ? ctx->d_hist16 : nullptr, ctx->d_fine_start);
            ctx->launches++;
        }
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
    TimeScope ts(ctx, SLIMM_GPU_T_ASSIGN);
    CU(cudaMemsetAsync(ctx->d_assign, 0, ctx->assign_words * 4, ctx->stream));
    u32 *uniq2 = ctx->d_assign, *lca = uniq2 + G, *cm = lca + (u64)8 * G, *fb = cm + (u64)8 * G;
    if (ctx->n) {
        const u32 n = (u32)ctx->n;
        const u64 n_chunks = ((u64)n + CHUNK - 1) / CHUNK;
        AssignParams P{};
        P.cw = ctx->d_cw; P.cw_idx = ctx->d_cw_idx; P.chunk_cnt = ctx->d_chunk_cnt; P.lr = ctx->d_lr;
        P.meta = ctx->d_meta; P.lin4 = (const uint4 *)ctx->d_lin; P.top_idx = ctx->d_top_idx; P.vb = ctx->d_valid_bits;
        P.G = G; P.half_avg = ctx->avg / 2u; P.wdiv = ctx->wdiv;
        P.uniq2_extra = uniq2; P.lca_cnt = lca; P.child_mark = cm; P.fb_mark = fb; P.cov2 = ctx->d_cov2;
        P.res_kind = (ctx->flags & SLIMM_GPU_READ_RESULTS) ? ctx->d_kind : nullptr; P.res_val = ctx->d_val;
        static const int asg_ctas = getenv("SLIMM_ASG_CTAS") ? atoi(getenv("SLIMM_ASG_CTAS")) : 8;   // CTAs per SM (experiments)
        const int rgrid = (int)std::max<u64>(1, std::min<u64>((n_chunks + 7) / 8, (u64)ctx->sm_count * asg_ctas));
        const uint4 *lin32 = (const uint4 *)ctx->d_lin;
        const bool extra = ctx->d_cw_idx != nullptr || P.res_kind != nullptr;
        const bool vrow = ctx->d_lin16v != nullptr;
        if (vrow) { k_lin_valid<<<(G + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_lin16, ctx->d_valid_bits, G, ctx->d_lin16v); ctx->launches++; }
        auto launch = [&](auto rec) {
            using R = decltype(rec);
            if (vrow && extra) k_assign_reads<R, Lin16, true, true><<<rgrid, 256, 0, ctx->stream>>>(rec, n, P, ctx->d_rs, ctx->d_lin16v);
            else if (vrow) k_assign_reads<R, Lin16, false, true><<<rgrid, 256, 0, ctx->stream>>>(rec, n, P, ctx->d_rs, ctx->d_lin16v);
            else if (extra && ctx->d_lin16) k_assign_reads<R, Lin16, true, false><<<rgrid, 256, 0, ctx->stream>>>(rec, n, P, ctx->d_rs, ctx->d_lin16);
            else if (extra) k_assign_reads<R, Lin32, true, false><<<rgrid, 256, 0, ctx->stream>>>(rec, n, P, ctx->d_rs, lin32);
            else if (ctx->d_lin16) k_assign_reads<R, Lin16, false, false><<<rgrid, 256, 0, ctx->stream>>>(rec, n, P, ctx->d_rs, ctx->d_lin16);
            else k_assign_reads<R, Lin32, false, false><<<rgrid, 256, 0, ctx->stream>>>(rec, n, P, ctx->d_rs, lin32);
        };
        if (ctx->use_sorted) {
            launch(RecPacked{ctx->d_rid_sorted, ctx->d_rp_sorted});
            if (P.res_kind) k_read_results_unique<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_kind, ctx->d_val, (const u32 *)ctx->d_rp_sorted, 2, ctx->d_valid_bits, n);
        } else {
            launch(RecSoA{ctx->d_rid, ctx->d_ref, ctx->d_pos});
            if (P.res_kind) k_read_results_unique<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_kind, ctx->d_val, ctx->d_ref, 1, ctx->d_valid_bits, n);
        }
        ctx->launches += 1 + (P.res_kind ? 1 : 0);
        CU(cudaGetLastError());
    }
    ctx->stage = ST_ASSIGN;
    ctx->h_assign_ok = false; ctx->finished = false;
    return SLIMM_GPU_OK;
}

int slimm_gpu_assign_device(slimm_gpu_ctx *ctx, void **d_ptr, uint64_t *n_u32)
{
    if (!ctx || !d_ptr || !n_u32) return SLIMM_GPU_EINVAL;
    *d_ptr = ctx->d_assign; *n_u32 = ctx->assign_words;
    return SLIMM_GPU_OK;
}

// adds the unique reads of the surviving references to uniq_reads_count2 and totals uniq_matches_count2;
// lazily, once, after the caller had the chance to sum the per-rank partials
static int finish_assign(slimm_gpu_ctx *ctx)
{
    if (ctx->finished || ctx->stage < ST_ASSIGN) return SLIMM_GPU_OK;
    CU(cudaSetDevice(ctx->device));
    k_finish_assign<<<(ctx->G + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_assign, ctx->d_stats, ctx->d_valid_bits, ctx->G, ctx->d_sc);
    ctx->launches++;
    CU(cudaGetLastError());
    ctx->finished = true;
    return SLIMM_GPU_OK;
}

int slimm_gpu_run(slimm_gpu_ctx *ctx, float cov_cut_off, uint32_t min_reads)
{
    int rc = slimm_gpu_coverage(ctx);
    if (rc) return rc;
    rc = slimm_gpu_filter(ctx, cov_cut_off, min_reads);
    if (rc) return rc;
    rc = slimm_gpu_assign(ctx);
    if (rc) return rc;
    return verify_input(ctx);
}

// ---- results -----------------------------------------------------------------------------------
int slimm_gpu_get_summary(slimm_gpu_ctx *ctx, slimm_gpu_summary *out)
{
    if (!ctx || !out) return SLIMM_GPU_EINVAL;
    if (ctx->stage < ST_COVERAGE) return fail(ctx, SLIMM_GPU_ESTATE, "nothing has run yet");
    CU(cudaSetDevice(ctx->device));
    { int rc = verify_input(ctx); if (rc) return rc; }
    { int rc = finish_assign(ctx); if (rc) return rc; }
    DevScalars s;
    CU(cudaMemcpyAsync(&s, ctx->d_sc, sizeof s, cudaMemcpyDeviceToHost, ctx->stream));
    u32 nrecv[2] = {0, 0};
    if (ctx->n_recv_on_device) CU(cudaMemcpyAsync(nrecv, ctx->d_n_recv, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (nrecv[1]) return fail(ctx, SLIMM_GPU_ERANGE, "a rank received more items than slimm_gpu_p2p_reserve reserved");
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
    { int rc = verify_input(ctx); if (rc) return rc; }
    if (ctx->h_assign_ok) return SLIMM_GPU_OK;
    if (ctx->stage < ST_ASSIGN) return fail(ctx, SLIMM_GPU_ESTATE, "assign has not run");
    CU(cudaSetDevice(ctx->device));
    { int rc = finish_assign(ctx); if (rc) return rc; }
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
    { int rc = verify_input(ctx); if (rc) return rc; }
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
    if ((ctx->flags & SLIMM_GPU_SKIP_BINS) && ctx->used_bucket && ctx->acc_mode == 1) return fail(ctx, SLIMM_GPU_EINVAL, "the bins were not kept (SLIMM_GPU_SKIP_BINS)");
    CU(cudaSetDevice(ctx->device));
    { int rc = verify_input(ctx); if (rc) return rc; }
    if (ctx->shard_n > 1) {   // a sharded run keeps the bins of the slices this rank owns only (the others were never written)
        u64 lo_bin = 0, hi_bin = 0;
        owned_bins(ctx, &lo_bin, &hi_bin);
        const u64 a = ctx->h_off[ref], b = a + ctx->h_len[ref] / ctx->w + 1u;
        if (a < lo_bin || b > hi_bin) return fail(ctx, SLIMM_GPU_ESTATE, "sharded run: the bins of this reference live on the rank that owns its histogram slices");
    }
    const u32 nb = ctx->h_len[ref] / ctx->w + 1u;
    if (cap < nb) return fail(ctx, SLIMM_GPU_EINVAL, "output buffer smaller than the number of bins");
    if (ctx->tmp_bins_cap < nb) {
        cudaFree(ctx->d_tmp_bins); ctx->d_tmp_bins = nullptr;
        CU(cudaMalloc(&ctx->d_tmp_bins, (size_t)nb * 4));
        ctx->tmp_bins_cap = nb;
    }
    const u32 *src = which == 2 ? ctx->d_cov2 : (const u32 *)ctx->d_hist;
    if (which != 2 && ctx->hist_compact)
        k_extract_bins_compact<<<(nb + 255) / 256, 256, 0, ctx->stream>>>((const u32 *)ctx->d_hist, ctx->d_hist16, ctx->d_fine_start, ctx->h_off[ref], which == 1 ? 1 : 0, nb,
                                                                         ctx->d_tmp_bins);
    else
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
    { int rc = verify_input(ctx); if (rc) return rc; }
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
    { int rc = verify_input(ctx); if (rc) return rc; }
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

int slimm_gpu_profile_failed(slimm_gpu_ctx *ctx, uint32_t *n)
{
    if (!ctx || !n) return SLIMM_GPU_EINVAL;
    if (!ctx->plan) return fail(ctx, SLIMM_GPU_ESTATE, "slimm_gpu_set_taxa has not been called");
    *n = ctx->plan->last_failed;
    return SLIMM_GPU_OK;
}

// ---- profile tail fed from the context -----------------------------------------------------------
int slimm_gpu_set_taxa(slimm_gpu_ctx *ctx, uint64_t n_taxa, const uint32_t *taxa_id, const uint8_t *taxa_rank, const uint8_t *taxa_has_name)
{
    if (!ctx || (n_taxa && (!taxa_id || !taxa_rank || !taxa_has_name))) return SLIMM_GPU_EINVAL;
    ctx->plan.reset(new slimm_host::ProfilePlan(ctx->G, ctx->h_len.data(), ctx->h_lin.data(), n_taxa, taxa_id, taxa_rank, taxa_has_name));
    if (ctx->plan->consistent) {             // the rank reduction can run on the device (k_rank_reduce)
        CU(cudaSetDevice(ctx->device));
        const u32 G = ctx->G;
        if (!ctx->d_lvl_idx) {
            CU(cudaMalloc(&ctx->d_lvl_idx, (size_t)8 * G * 4));
            CU(cudaMalloc(&ctx->d_top_lvl7, (size_t)std::max<u32>(ctx->n_top, 1) * 4));
            CU(cudaMalloc(&ctx->d_agg, (size_t)10 * G * 4));
            CU(cudaHostAlloc((void **)&ctx->h_agg, (size_t)10 * G * 4, cudaHostAllocDefault));
            CU(cudaHostAlloc((void **)&ctx->h_sc, sizeof(DevScalars), cudaHostAllocDefault));
        }
        std::vector<u32> top7(std::max<u32>(ctx->n_top, 1), 0);
        for (u32 g = 0; g < G; ++g) top7[ctx->h_top_idx[g]] = ctx->plan->lvl_idx[(size_t)7 * G + g];
        CU(cudaMemcpy(ctx->d_lvl_idx, ctx->plan->lvl_idx.data(), (size_t)8 * G * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(ctx->d_top_lvl7, top7.data(), top7.size() * 4, cudaMemcpyHostToDevice));
    }
    return SLIMM_GPU_OK;
}

int slimm_gpu_profile(slimm_gpu_ctx *ctx, uint32_t rank, float abundance_cut_off, slimm_profile_row *rows, uint64_t cap, uint64_t *n)
{
    if (!ctx || !n) return SLIMM_GPU_EINVAL;
    if (!ctx->plan) return fail(ctx, SLIMM_GPU_ESTATE, "slimm_gpu_set_taxa has not been called");
    if (ctx->stage < ST_ASSIGN) return fail(ctx, SLIMM_GPU_ESTATE, "assign has not run");
    if (rank < 1 || rank > 6) return fail(ctx, SLIMM_GPU_EINVAL, "rank must be 1 (species) .. 6 (phylum)");
    { int rc = verify_input(ctx); if (rc) return rc; }
    if (ctx->plan->consistent && ctx->tail_mode != 1) {
        // K7: per-rank segmented reduction on the device, only the per-taxon aggregates of two ranks come back
        CU(cudaSetDevice(ctx->device));
        { int rc0 = finish_assign(ctx); if (rc0) return rc0; }
        const u32 G = ctx->G;
        CU(cudaMemsetAsync(ctx->d_agg, 0, (size_t)10 * G * 4, ctx->stream));
        CU(cudaMemsetAsync(ctx->d_agg + (size_t)3 * G, 0xFF, (size_t)G * 4, ctx->stream));   // kmin of the rank
        CU(cudaMemsetAsync(ctx->d_agg + (size_t)8 * G, 0xFF, (size_t)G * 4, ctx->stream));   // kmin of the parent rank
        k_rank_reduce<<<(G + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_assign, ctx->d_meta, ctx->d_lvl_idx, ctx->d_top_lvl7, G, ctx->n_top, rank, ctx->d_agg);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(ctx->h_agg, ctx->d_agg, (size_t)10 * G * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(ctx->h_sc, ctx->d_sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<slimm_profile_row> out;
        int rc0 = ctx->plan->finish_from_aggregates(ctx->h_agg, (u32)ctx->h_sc->n_reads, ctx->avg, ctx->h_sc->cut, abundance_cut_off, rank, out);
        if (rc0) return fail(ctx, rc0, "profile aggregation failed (inconsistent stage outputs)");
        *n = out.size();
        for (u64 i = 0; i < out.size() && i < cap; ++i)
            if (rows) rows[i] = out[i];
        ctx->tail_host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return SLIMM_GPU_OK;
    }
    int rc = fetch_assign(ctx);
    if (rc) return rc;
    slimm_gpu_summary sm;
    rc = slimm_gpu_get_summary(ctx, &sm);
    if (rc) return rc;
    const auto t_host0 = std::chrono::steady_clock::now();
    const u32 G = ctx->G;
    const u32 *uniq2 = ctx->h_assign.data(), *lca = uniq2 + G, *cm = lca + (u64)8 * G, *fb = cm + (u64)8 * G;
    slimm_host::ProfilePlan &pl = *ctx->plan;
    pl.begin();
    for (u64 s = 0; s < (u64)G * 8; ++s)
        if (lca[s]) pl.add_direct(pl.slot_t[s], lca[s]);
    // children in ascending reference order per taxon keeps the sets sorted without a sort
    for (u32 g = 0; g < G; ++g) {
        for (u32 l = 0; l < 8; ++l)
            if (cm[(u64)g * 8 + l]) pl.add_child(pl.slot_t[(u64)g * 8 + l], g);
        for (u32 t = 0; t < ctx->n_top; ++t)
            if (fb[(u64)t * G + g]) { int ti = pl.find(ctx->h_top_vals[t]); if (ti >= 0) pl.add_child((u32)ti, g); }
    }
    std::vector<slimm_profile_row> out;
    rc = pl.finish(uniq2, sm.matches_count, ctx->avg, sm.coverage_cut_off, abundance_cut_off, rank, out);
    if (rc) return fail(ctx, rc, "profile aggregation failed (inconsistent stage outputs)");
    *n = out.size();
    for (u64 i = 0; i < out.size() && i < cap; ++i)
        if (rows) rows[i] = out[i];
    ctx->tail_host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
    return SLIMM_GPU_OK;
}

int slimm_gpu_set_scatter_mode(slimm_gpu_ctx *ctx, int mode)
{
    if (!ctx || mode < -1 || mode > 1) return SLIMM_GPU_EINVAL;
    ctx->scatter_mode = mode;
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
        if (i == SLIMM_GPU_T_TAIL_HOST) { ms[i] = ctx->tail_host_ms; continue; }
        if (ctx->timing && ctx->ev_used[i]) cudaEventElapsedTime(&ms[i], ctx->ev[i][0], ctx->ev[i][1]);
    }
    return SLIMM_GPU_OK;
}

int slimm_gpu_get_launch_count(slimm_gpu_ctx *ctx, uint64_t *n) { if (!ctx || !n) return SLIMM_GPU_EINVAL; *n = ctx->launches; return SLIMM_GPU_OK; }

}  // extern "C"

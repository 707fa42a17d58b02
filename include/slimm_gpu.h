/*
 * slimm_gpu.h - C ABI of the B200-native SLIMM profiling hot path.
 *
 * The reference (seqan/slimm 0.3.4) has no plugin / FFI seam: the hot path is five member
 * functions of `class slimm` mutating its members (reference src/slimm.hpp:92-165), called from
 * `slimm::get_profiles()` at src/slimm.hpp:449 (analyze_alignments), :464 (filter_alignments),
 * :485 (get_reads_lca_count) and :489 (write_abundance).  This header DEFINES the seam at those
 * four call sites.  Every entry point names the reference code it replaces.  INTEGRATION.md shows
 * the patch a maintainer of the reference would apply to call it.
 *
 * Conventions: plain C types only; host buffers are caller-owned, device memory is library-owned
 * (unless handed in through slimm_gpu_push_device); every call returns 0 on success or a
 * SLIMM_GPU_E* code (slimm_gpu_strerror / slimm_gpu_last_error give text); no exceptions cross the
 * boundary; one controlling host thread per context.  There is no CPU fallback: without a CUDA
 * device slimm_gpu_create fails with SLIMM_GPU_ENODEVICE.
 *
 * u32 arithmetic wraps exactly as in the reference; f32 results are IEEE binary32 computed with
 * the reference's operation order ("exactness").
 */
#ifndef SLIMM_GPU_H
#define SLIMM_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLIMM_LINEAGE_LENGTH 8 /* reference src/misc.hpp:24-35 taxa_ranks strain..superkingdom */

enum {
    SLIMM_GPU_OK = 0,
    SLIMM_GPU_EINVAL = 1,     /* bad argument / record referencing a contig >= n_refs          */
    SLIMM_GPU_ENODEVICE = 2,  /* no usable CUDA device (there is no CPU fallback)               */
    SLIMM_GPU_ECUDA = 3,      /* a CUDA runtime call failed, see slimm_gpu_last_error           */
    SLIMM_GPU_ENOMEM = 4,     /* device or pinned-host allocation failed                        */
    SLIMM_GPU_ESTATE = 5,     /* stages called out of order                                     */
    SLIMM_GPU_ERANGE = 6      /* more than 2^32-256 records (hits_count is u32 in the reference) */
};

/* flags for slimm_gpu_config.flags */
#define SLIMM_GPU_KEEP_UNIQ_COV2 1u /* also build uniq_cov2 bins (needed by -co / -ro outputs)        */
#define SLIMM_GPU_READ_RESULTS 2u   /* keep per-read reassignment / LCA results for slimm_gpu_read_results */
#define SLIMM_GPU_SKIP_BINS 4u      /* profile-only run (no -co / -ro): when the histogram is accumulated slice by slice in
                                       shared memory, do not write the bins back to HBM; slimm_gpu_fetch_bins and
                                       slimm_gpu_bins_device are then unavailable.  Results are unchanged. */

typedef struct slimm_gpu_ctx slimm_gpu_ctx;

/* Per-sample constants.  Replaces the reference-contig initialisation loop of get_profiles(),
 * reference src/slimm.hpp:420-445, and the reference_contig / bins_coverage constructors,
 * src/reference_contig.hpp:77-82,129-145 (three zeroed histograms of len/w+1 bins per contig). */
typedef struct {
    uint32_t n_refs;            /* contigs in @SQ order                                             */
    const uint32_t *ref_len;    /* [n_refs]   @SQ LN                                                */
    const uint32_t *lineage;    /* [n_refs*8] db.ac__taxid[accession(g)]; zeros for unknown         */
    uint32_t bin_width;         /* -w, or the average read length when 0 (src/slimm.hpp:412-413)    */
    uint32_t avg_read_length;   /* get_avg_read_length(), src/misc.hpp:509-522                      */
    uint64_t reserve_records;   /* initial capacity of the device record store (0: grow on demand)  */
    int32_t device;             /* CUDA device ordinal                                              */
    uint32_t flags;             /* SLIMM_GPU_*                                                      */
} slimm_gpu_config;

/* Scalars of the path; names follow the reference's members (src/slimm.hpp:105-118). */
typedef struct {
    uint32_t hits_count;          /* kept records                                   :212 */
    uint32_t matches_count;       /* distinct reads                                 :257 */
    uint32_t uniq_matches_count;  /* reads with one target (== uniq_hits_count)     :225,236 */
    uint32_t uniq_matches_count2; /* reads with one surviving target                :387 */
    uint32_t reference_count;     /* references with reads                          :264 */
    uint32_t n_valid;             /* |valid_ref_ids|                                :361 */
    uint32_t failed_by_cov;       /* -v counters                                    :365-376 */
    uint32_t failed_by_uniq_cov;
    uint32_t failed_by_min_read;
    uint32_t min_reads;           /* -mr or 1+(R-1)/10000                           :458-459 */
    float coverage_cut_off;       /* coverage_cut_off()                             :328-344 */
    float uniq_coverage_cut_off;  /* uniq_coverage_cut_off()                        :672-688 */
    uint64_t n_pairs;             /* distinct (read, ref) pairs = sum of reads_count */
    uint64_t n_bins;              /* sum over contigs of len/w+1                     */
    uint32_t input_was_sorted;    /* 1: read ids were non-decreasing, no device sort needed */
    uint32_t reserved;
} slimm_gpu_summary;

const char *slimm_gpu_strerror(int code);
const char *slimm_gpu_last_error(const slimm_gpu_ctx *ctx);
int slimm_gpu_device_count(int *n);

int slimm_gpu_create(const slimm_gpu_config *cfg, slimm_gpu_ctx **out);
int slimm_gpu_destroy(slimm_gpu_ctx *ctx);
/* New sample against the same contigs/database (directory mode; reference slimm::reset(),
 * src/slimm.hpp:167-188).  bin_width / avg_read_length may change, 0 keeps the old value. */
int slimm_gpu_reset(slimm_gpu_ctx *ctx, uint32_t bin_width, uint32_t avg_read_length);
/* Run all kernels on this cudaStream_t (e.g. the caller's framework stream).  NULL: a library-owned non-blocking
 * stream; for the legacy default stream pass cudaStreamLegacy ((cudaStream_t)0x1).  Work the caller orders against
 * the stages (NCCL reductions of the *_device() buffers) must be issued on the same stream. */
int slimm_gpu_set_stream(slimm_gpu_ctx *ctx, void *cuda_stream);

/* Pinned host buffers for the decode threads to pack batches into (cudaHostAlloc). */
int slimm_gpu_host_alloc(void **p, uint64_t bytes);
int slimm_gpu_host_free(void *p);

/* Ingest: append a struct-of-arrays batch of KEPT records (flag&4 / rID==-1 records are dropped by
 * the decoder, reference src/slimm.hpp:197) in file order.  read_id is the dense id of
 * qName + ".1"/".2" (src/slimm.hpp:204-208).  Replaces the record loop of analyze_alignments,
 * src/slimm.hpp:194-213.  Host variant: asynchronous H2D on a copy stream when the buffers are
 * pinned; the buffers may be reused after slimm_gpu_sync_uploads.  Device variant: the caller's
 * device arrays are used in place (zero copy; single batch; must outlive the run). */
int slimm_gpu_push(slimm_gpu_ctx *ctx, const uint32_t *read_id, const uint32_t *ref_id,
                   const int32_t *begin_pos, uint64_t n);
int slimm_gpu_push_device(slimm_gpu_ctx *ctx, const uint32_t *d_read_id, const uint32_t *d_ref_id,
                          const int32_t *d_begin_pos, uint64_t n);
/* The same ingest in the 6.125-byte wire format of input GROUPED BY READ (what mappers write; the decoder knows after its own
 * run check): instead of a 32-bit read id one bit per record - bit i%32 of word i/32 set when record i starts a new read - and
 * the reference id as 16 bits (fewer than 65 536 contigs).  The dense read ids are rebuilt on the device (a running count of the
 * set bits, continued across calls), so the PCIe link carries half the bytes of slimm_gpu_push.  The first record of a sample
 * must have its bit set.  Calls may be mixed with slimm_gpu_push only in whole samples (SLIMM_GPU_ESTATE otherwise). */
int slimm_gpu_push_packed(slimm_gpu_ctx *ctx, const uint32_t *new_read_bits, const uint16_t *ref_id16,
                          const int32_t *begin_pos, uint64_t n);
int slimm_gpu_sync_uploads(slimm_gpu_ctx *ctx);

/* Stage 1 - coverage.  Replaces the per-read loop of analyze_alignments (src/slimm.hpp:219-257,
 * read_stat::add_target src/read_stat.hpp:116-135, is_uniq :72-75): first-occurrence dedupe of
 * (read, ref), scatter-add into the cov / uniq_cov bins.  Leaves THIS rank's partial histogram on
 * the device; with several GPUs sum slimm_gpu_bins_device() across ranks before stage 2. */
int slimm_gpu_coverage(slimm_gpu_ctx *ctx);
/* Device pointer to the interleaved {cov, uniq_cov} u32 histogram for an in-place sum over ranks;
 * n_u32 is the number of u32 words.  Only for runs that keep the interleaved layout (small histograms, SLIMM_GPU_ACC=l2,
 * SLIMM_GPU_COMPACT_BINS=0): the fine-slice accumulate keeps most bins as {cov:16 | uniq_cov:16} words and answers
 * SLIMM_GPU_ESTATE here - read bins with slimm_gpu_fetch_bins. */
int slimm_gpu_bins_device(slimm_gpu_ctx *ctx, void **d_ptr, uint64_t *n_u32);
/* Device pointer to the read-level partial counters {matches_count, uniq_matches_count} (2 x u64,
 * summed across ranks together with the bins). */
int slimm_gpu_counters_device(slimm_gpu_ctx *ctx, void **d_ptr, uint64_t *n_u64);
/* Tell the context the global number of kept records when records are sharded over ranks. */
int slimm_gpu_set_global_hits(slimm_gpu_ctx *ctx, uint64_t hits);

/* ---- several GPUs, one process each: histogram slices sharded over ranks ------------------------------------------
 * Reads are sharded by read id (the caller pushes each rank its reads); the bins are sharded by histogram slice
 * (2^22 consecutive padded bins): rank r of n owns slices [S*r/n, S*(r+1)/n).  In a sharded run slimm_gpu_coverage
 * stops after grouping this rank's items by slice; the caller routes every item to the rank that owns its slice
 * (one all-to-all of 4 bytes per record; the per-slice counts give the split sizes), hands the received items to
 * slimm_gpu_accumulate_items, which fills and scans only the owned bins, and sums the partial per-reference
 * statistics (slimm_gpu_stats_device, 16 bytes per reference) and the read counters (slimm_gpu_counters_device)
 * over ranks before slimm_gpu_filter.  Nothing of the size of the histogram ever crosses NVLink.
 * uniq_cov2 bins (SLIMM_GPU_KEEP_UNIQ_COV2) are a single-GPU feature. */
int slimm_gpu_set_shard(slimm_gpu_ctx *ctx, uint32_t rank, uint32_t n_ranks);
/* items per slice of this rank after slimm_gpu_coverage (host copy; counts may be NULL to query n_slices) */
int slimm_gpu_get_slice_counts(slimm_gpu_ctx *ctx, uint32_t *counts, uint32_t cap, uint32_t *n_slices);
/* device pointer to this rank's items grouped by slice (u32 each, slice s starts at the sum of the counts before it) */
int slimm_gpu_items_device(slimm_gpu_ctx *ctx, void **d_items);
/* apply the items received for the owned slices (device pointer) and scan the owned bins */
int slimm_gpu_accumulate_items(slimm_gpu_ctx *ctx, const uint32_t *d_items, uint64_t n_items);
/* device pointer to {nz, reads_count, uniq nz, uniq_reads_count}[n_refs] (u32) for the sum over ranks */
int slimm_gpu_stats_device(slimm_gpu_ctx *ctx, void **d_ptr, uint64_t *n_u32);

/* ---- the same exchange without NCCL: peer-to-peer stores over NVLink, fused into the split ---------------------------
 * On one NVLink/NVSwitch box every rank maps every other rank's receive buffer (CUDA IPC) and the multisplit kernel
 * writes each item straight to the rank that owns its slice, tile by tile - the all-to-all IS the split's store phase.
 *   once:      slimm_gpu_set_shard; slimm_gpu_p2p_reserve (same capacity on every rank; returns the 64-byte IPC handle);
 *              exchange the handles; slimm_gpu_p2p_connect(all handles in rank order)
 *   per sample: slimm_gpu_coverage (stops before the split); all-gather slimm_gpu_get_slice_counts over ranks;
 *              slimm_gpu_split_to_peers(table [n_ranks][n_slices]); a cross-rank barrier ordered on the stream (all
 *              splits complete); slimm_gpu_accumulate_received; then as above (sum the statistics, filter, ...).
 * Three exchanges behind the same calls (SLIMM_PEER_ROUTE, read by slimm_gpu_p2p_connect):
 *   2 (default) blocks: the items are grouped by slice locally, as on one GPU; a rank owns consecutive slices, so its share is one
 *               contiguous block, copied to its receive buffer with 128-byte-aligned stores over NVLink (k_peer_copy).  The
 *               receiver holds the blocks source by source, each ordered by slice, and needs no coarse pass of its own.
 *   1 routed:   a tile's items are ranked by owner rank and travel as one segment per (tile, rank); the receiver groups by slice.
 *   0 split:    the multisplit stores every (tile, slice) run straight into the owner's buffer (runs of ~20 items). */
int slimm_gpu_p2p_reserve(slimm_gpu_ctx *ctx, uint64_t cap_items, void *ipc_handle_64);
int slimm_gpu_p2p_connect(slimm_gpu_ctx *ctx, const void *ipc_handles /* [n_ranks][64] */, uint32_t n_ranks);
int slimm_gpu_split_to_peers(slimm_gpu_ctx *ctx, const uint32_t *all_counts /* [n_ranks][n_slices] */, uint64_t *n_recv);
/* The same exchange planned on the device: d_all_counts is the all-gathered table of slimm_gpu_slice_counts_device (device memory,
 * [n_ranks][n_slices]); no host copy of the counts, no synchronisation - coverage, the collectives and the split queue up on the
 * stream.  A receive buffer that would overflow is reported by slimm_gpu_get_summary (SLIMM_GPU_ERANGE). */
int slimm_gpu_slice_counts_device(slimm_gpu_ctx *ctx, void **d_counts, uint32_t *n_slices);

/* ---- several GPUs driven by ONE host process (the C++ front end: `slimm --gpus N`) -----------------------------------
 * ctxs[0..n): contexts on n different devices, created from the same configuration, slimm_gpu_set_shard(r, n) done, each fed
 * the records of ITS reads (all records of a read in one context, read ids non-decreasing per context).  Runs the whole path:
 * coverage everywhere, the items to the owners of their histogram slices (n x n peer copies, each source's share of an owner
 * is one contiguous block), accumulate, the per-reference statistics summed, filter and assign everywhere, the assign blocks
 * summed.  Afterwards every context answers slimm_gpu_get_summary / get_ref_stats / profile with the global results.  No NCCL,
 * no second process: cudaMemcpyPeerAsync and host-side sums of 16 bytes per reference and the assign block.
 * global_hits: kept records over all contexts.  Replaces, like slimm_gpu_run, reference src/slimm.hpp:449,464,485. */
int slimm_gpu_run_sharded_local(slimm_gpu_ctx **ctxs, uint32_t n, float cov_cut_off, uint32_t min_reads, uint64_t global_hits);
int slimm_gpu_split_to_peers_device(slimm_gpu_ctx *ctx, const uint32_t *d_all_counts);
int slimm_gpu_accumulate_received(slimm_gpu_ctx *ctx);
/* back to the all-to-all exchange (e.g. when another rank could not map the buffers) */
int slimm_gpu_p2p_disable(slimm_gpu_ctx *ctx);

/* Stage 2 - reference filter.  Replaces none_zero_bin_count / cov_percent / uniq_cov_percent
 * (src/reference_contig.hpp:84-91,148-155), coverage_cut_off / uniq_coverage_cut_off
 * (src/slimm.hpp:328-344,672-688) with get_quantile_cut_off (src/misc.hpp:197-216) and the
 * reference loop of filter_alignments (src/slimm.hpp:354-378).  cov_cut_off is -cc, min_reads is
 * -mr (0: default). */
int slimm_gpu_filter(slimm_gpu_ctx *ctx, float cov_cut_off, uint32_t min_reads);

/* Stage 3 - reassignment + LCA.  Replaces the read loop of filter_alignments
 * (src/slimm.hpp:380-391, read_stat::update src/read_stat.hpp:98-114), slimm::get_lca
 * (src/slimm.hpp:516-531) and phase 1 of get_reads_lca_count (:536-557).  Partial per rank: sum
 * slimm_gpu_assign_device() across ranks (u32 words; the child marks are 0/1 counters). */
int slimm_gpu_assign(slimm_gpu_ctx *ctx);
int slimm_gpu_assign_device(slimm_gpu_ctx *ctx, void **d_ptr, uint64_t *n_u32);

/* Convenience: coverage + filter + assign on one GPU. */
int slimm_gpu_run(slimm_gpu_ctx *ctx, float cov_cut_off, uint32_t min_reads);

/* ---- results (device -> caller-owned host buffers; any pointer may be NULL) ------------------- */
int slimm_gpu_get_summary(slimm_gpu_ctx *ctx, slimm_gpu_summary *out);
/* per reference [n_refs]: reference_contig members, src/reference_contig.hpp:100-127 */
int slimm_gpu_get_ref_stats(slimm_gpu_ctx *ctx, uint32_t *reads_count, uint32_t *uniq_reads_count,
                            uint32_t *uniq_reads_count2, uint32_t *nz_bins, uint32_t *uniq_nz_bins,
                            float *cov_percent, float *uniq_cov_percent, uint8_t *valid);
/* taxon_id__read_count after phase 1 as (taxon, count) sorted by taxon; taxon_id__children after
 * phase 1 as distinct (taxon, ref) pairs sorted.  *n receives the number available; at most cap
 * entries are written. */
int slimm_gpu_get_lca_counts(slimm_gpu_ctx *ctx, uint32_t *taxon, uint32_t *count, uint64_t cap, uint64_t *n);
int slimm_gpu_get_lca_children(slimm_gpu_ctx *ctx, uint32_t *taxon, uint32_t *ref, uint64_t cap, uint64_t *n);
/* bins of one reference; which: 0 cov, 1 uniq_cov, 2 uniq_cov2 (needs SLIMM_GPU_KEEP_UNIQ_COV2).
 * Replaces direct reads of bins_height in write_coverage / write_raw_stat (src/slimm.hpp:846-943). */
int slimm_gpu_fetch_bins(slimm_gpu_ctx *ctx, int which, uint32_t ref, uint32_t *out, uint32_t cap);
/* nonzero uniq_cov2 bins per reference (raw output column), needs SLIMM_GPU_KEEP_UNIQ_COV2 */
int slimm_gpu_get_uniq2_nz(slimm_gpu_ctx *ctx, uint32_t *uniq2_nz_bins);
/* per-read outcome (needs SLIMM_GPU_READ_RESULTS): for every read with >= 1 surviving target its id,
 * kind (1: reassigned to a sole survivor, 2: LCA) and value (reference id or LCA taxon). */
int slimm_gpu_read_results(slimm_gpu_ctx *ctx, uint32_t *read_id, uint8_t *kind, uint32_t *value,
                           uint64_t cap, uint64_t *n);

/* ---- instrumentation ------------------------------------------------------------------------- */
enum { SLIMM_GPU_T_SORT = 0, SLIMM_GPU_T_ZERO, SLIMM_GPU_T_SPLIT, SLIMM_GPU_T_COVERAGE, SLIMM_GPU_T_ACCUM,
       SLIMM_GPU_T_STATS, SLIMM_GPU_T_CUTOFF, SLIMM_GPU_T_ASSIGN, SLIMM_GPU_T_TAIL_HOST, SLIMM_GPU_T_COUNT };
int slimm_gpu_enable_timing(slimm_gpu_ctx *ctx, int on);
/* CUDA-event durations (ms) of the last run, per kernel group (SLIMM_GPU_T_ZERO: the histogram memset, which runs on a
 * second stream underneath k_coverage / k_split when the bucketed scatter is used; SLIMM_GPU_T_TAIL_HOST: host wall time of the rank
 * aggregation in slimm_gpu_profile, after the results arrived), and launches issued since create */
int slimm_gpu_get_timings(slimm_gpu_ctx *ctx, float *ms, int n);
int slimm_gpu_get_launch_count(slimm_gpu_ctx *ctx, uint64_t *n);

/* Scatter strategy of the coverage stage: -1 automatic (bucketed multisplit when the histogram is much
 * larger than L2), 0 direct global REDs, 1 bucketed.  Results are identical; tests force both. */
int slimm_gpu_set_scatter_mode(slimm_gpu_ctx *ctx, int mode);

/* ---- host-side tail of the path (no GPU needed) ---------------------------------------------- */
/* Rank aggregation and abundances.  Replaces phases 2 and 3 of get_reads_lca_count
 * (src/slimm.hpp:560-610) and the numeric part of write_abundance (:733-843). */
typedef struct {
    uint32_t taxon;        /* taxa_id column (without the '*')                                  */
    uint32_t kind;         /* 0 plain row, 1 "<parent>*" unclassified row, 2 the final "0*" row  */
    uint32_t read_count;
    uint32_t first_child;  /* smallest contributing reference: its lineage names the row; ~0u for 0* */
    double abundance;      /* f32 value for kinds 0/1, f64 for kind 2 (100.0 - float)            */
} slimm_profile_row;

typedef struct {
    uint32_t n_refs;
    const uint32_t *ref_len;           /* [n_refs]   */
    const uint32_t *lineage;           /* [n_refs*8] */
    uint64_t n_taxa;                   /* db.taxid__name restricted to what the caller has */
    const uint32_t *taxa_id;           /* [n_taxa] */
    const uint8_t *taxa_rank;          /* [n_taxa] 0..8 */
    const uint8_t *taxa_has_name;      /* [n_taxa] name != "" */
    uint64_t n_direct;                 /* slimm_gpu_get_lca_counts */
    const uint32_t *direct_taxon, *direct_count;
    uint64_t n_children;               /* slimm_gpu_get_lca_children */
    const uint32_t *child_taxon, *child_ref;
    const uint32_t *uniq_reads_count2; /* [n_refs] */
    uint32_t matches_count, avg_read_length;
    float coverage_cut_off, abundance_cut_off;
    uint32_t rank;                     /* 1 species .. 6 phylum */
} slimm_profile_input;

/* rows: plain rows (ascending taxon), then "<parent>*" rows (ascending), then "0*".  Returns the
 * number of rows through *n (at most cap written). */
int slimm_profile_rows(const slimm_profile_input *in, slimm_profile_row *rows, uint64_t cap, uint64_t *n);

/* The same tail fed straight from the context (no intermediate taxon lists): slimm_gpu_set_taxa once per
 * database (db.taxid__name: rank 0..8 and whether the name is non-empty), then slimm_gpu_profile after
 * slimm_gpu_assign.  Replaces the calls at reference src/slimm.hpp:485 (phases 2-3) and :489 (numbers). */
int slimm_gpu_set_taxa(slimm_gpu_ctx *ctx, uint64_t n_taxa, const uint32_t *taxa_id, const uint8_t *taxa_rank,
                       const uint8_t *taxa_has_name);
int slimm_gpu_profile(slimm_gpu_ctx *ctx, uint32_t rank, float abundance_cut_off, slimm_profile_row *rows,
                      uint64_t cap, uint64_t *n);
/* taxa of the requested rank the last slimm_gpu_profile call dropped for abundance / coverage / missing name
 * (faild_count of write_abundance, reference src/slimm.hpp:797-801; only printed with -v) */
int slimm_gpu_profile_failed(slimm_gpu_ctx *ctx, uint32_t *n);

/* 1 when the lineage table is tree-consistent (every taxon on one level with that rank in db.taxid__name, never 0,
 * one parent): slimm_gpu_profile then runs the rank reduction on the device (k_rank_reduce); any other database
 * takes the general host path, which replays the reference's set semantics ("profile tail"). */
int slimm_profile_db_is_tree_consistent(uint32_t n_refs, const uint32_t *lineage, uint64_t n_taxa, const uint32_t *taxa_id,
                                        const uint8_t *taxa_rank, const uint8_t *taxa_has_name, int *out);

#ifdef __cplusplus
}
#endif
#endif /* SLIMM_GPU_H */

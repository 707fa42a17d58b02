// profile_host.h - dense-array plan for the host tail of the hot path (see profile_host.cpp).
#ifndef SLIMM_PROFILE_HOST_H
#define SLIMM_PROFILE_HOST_H

#include <cstdint>
#include <vector>

#include "../../include/slimm_gpu.h"

namespace slimm_host {

typedef uint32_t u32;
typedef uint64_t u64;

struct ProfilePlan {
    u32 G;
    std::vector<u32> len, lin;
    std::vector<u32> vals;      // sorted distinct taxon ids of the lineage table (dense index -> taxon)
    std::vector<u32> slot_t;    // [G*8] dense taxon index of every lineage slot
    std::vector<uint8_t> rank, named;
    // per-sample scratch, cleared through the touched list
    std::vector<u32> count, direct, pcnt, scnt;
    std::vector<float> pab, sab;
    std::vector<uint8_t> has_count, dirty, seen, has_p, has_s;
    std::vector<std::vector<u32>> kids;
    std::vector<u32> touched, snapshot;
    std::vector<u32> stamp, kmin, kmax;   // duplicate filter over references; smallest / largest member of a normalized set
    u32 epoch = 0;
    mutable u32 last_failed = 0;        // taxa of the requested rank the last finish* call dropped (write_abundance's faild_count)
    // tree-consistent databases: per-level dense taxon index of every reference and the taxa of each level (ascending)
    bool consistent = false;
    std::vector<u32> lvl_idx;            // [8][G]
    std::vector<u32> lvl_taxa[8];        // dense taxon ids (indices into vals) of the level, ascending

    ProfilePlan(u32 n_refs, const u32 *ref_len, const u32 *lineage, u64 n_taxa, const u32 *taxa_id,
                const uint8_t *taxa_rank, const uint8_t *taxa_has_name);
    int find(u32 taxon) const;
    void begin();
    void add_direct(u32 t, u32 c);      // t: dense index
    void add_child(u32 t, u32 ref);
    int finish(const u32 *uniq_reads_count2, u32 matches_count, u32 avg_read_length, float coverage_cut_off,
               float abundance_cut_off, u32 rank, std::vector<slimm_profile_row> &out);

    // rows from the per-level aggregates of k_rank_reduce (agg: [2][count|kn|klen|kmin|kmax][G]); consistent databases only
    int finish_from_aggregates(const u32 *agg, u32 matches_count, u32 avg_read_length, float coverage_cut_off,
                               float abundance_cut_off, u32 rank, std::vector<slimm_profile_row> &out) const;

private:
    void check_consistency();
    void touch(u32 t);
    void normalize(u32 t);
};

}  // namespace slimm_host
#endif

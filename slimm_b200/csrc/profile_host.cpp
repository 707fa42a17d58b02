// profile_host.cpp - host tail of the hot path: rank aggregation + abundances.
//
// Replaces phases 2 and 3 of slimm::get_reads_lca_count (reference src/slimm.hpp:560-610) and the
// numeric part of slimm::write_abundance (:733-843).  O(G + T) work on the arrays the GPU stages
// produced; the text formatting (lineage strings, TSV) stays with the caller.
//
// Sets of contributing references are kept as sorted vectors that are merged lazily; the reference
// uses std::set<uint32_t> per taxon.  The reference walks its snapshot of direct counts in
// libstdc++ hash order; here the order is ascending (rank, taxon) - identical results whenever the
// lineage table is tree-consistent.
#include <algorithm>
#include <cstdint>
#include <unordered_map>
#include <vector>

#include "../../include/slimm_gpu.h"

namespace {

typedef uint32_t u32;
typedef uint64_t u64;

struct Node {
    u32 count = 0;
    bool has_count = false;
    bool dirty = false;
    std::vector<u32> kids;   // taxon_id__children[t]
    void add_count(u32 c) { count += c; has_count = true; }   // u32 wrap as increment_or_initialize
    void normalize()
    {
        if (!dirty) return;
        std::sort(kids.begin(), kids.end());
        kids.erase(std::unique(kids.begin(), kids.end()), kids.end());
        dirty = false;
    }
};

struct TaxInfo { uint8_t rank = 0; uint8_t has_name = 0; };

}  // namespace

extern "C" int slimm_profile_rows(const slimm_profile_input *in, slimm_profile_row *rows, uint64_t cap, uint64_t *n_out)
{
    if (!in || !n_out || in->rank < 1 || in->rank > 6 || !in->lineage || !in->ref_len) return SLIMM_GPU_EINVAL;
    const u32 G = in->n_refs;
    const u32 *lin = in->lineage;
    std::unordered_map<u32, TaxInfo> info;
    info.reserve(in->n_taxa * 2 + 16);
    for (u64 i = 0; i < in->n_taxa; ++i) {
        TaxInfo ti;
        ti.rank = in->taxa_rank[i];
        ti.has_name = in->taxa_has_name[i];
        info[in->taxa_id[i]] = ti;
    }
    auto rank_of = [&](u32 t) -> u32 { auto it = info.find(t); return it == info.end() ? 0u : it->second.rank; };
    auto has_name = [&](u32 t) -> bool { auto it = info.find(t); return it != info.end() && it->second.has_name; };

    std::unordered_map<u32, Node> nodes;
    nodes.reserve(in->n_direct * 4 + (u64)G + 64);
    // phase 1 results from the GPU (src/slimm.hpp:536-557)
    for (u64 i = 0; i < in->n_direct; ++i) nodes[in->direct_taxon[i]].add_count(in->direct_count[i]);
    for (u64 i = 0; i < in->n_children; ++i) {
        if (in->child_ref[i] >= G) return SLIMM_GPU_EINVAL;
        Node &nd = nodes[in->child_taxon[i]];
        nd.kids.push_back(in->child_ref[i]);
        nd.dirty = true;
    }
    // phase 2 (:560-586): push each direct count and its children up the first child's lineage
    std::vector<std::pair<u32, u32>> snapshot;   // (taxon, count)
    snapshot.reserve(in->n_direct);
    for (u64 i = 0; i < in->n_direct; ++i) snapshot.emplace_back(in->direct_taxon[i], in->direct_count[i]);
    std::sort(snapshot.begin(), snapshot.end(), [&](const std::pair<u32, u32> &a, const std::pair<u32, u32> &b) {
        u32 ra = rank_of(a.first), rb = rank_of(b.first);
        return ra != rb ? ra < rb : a.first < b.first;
    });
    std::vector<u32> kids;
    for (auto &tc : snapshot) {
        Node &nd = nodes[tc.first];
        nd.normalize();
        if (nd.kids.empty()) return SLIMM_GPU_EINVAL;           // .at() would throw in the reference
        kids = nd.kids;                                          // copied before the loop (:575)
        const u32 f = kids[0];
        for (u32 j = rank_of(tc.first) + 1; j < 8; ++j) {
            Node &rc = nodes[lin[(u64)f * 8 + j]];               // may rehash: nd is not used below
            rc.add_count(tc.second);
            rc.kids.insert(rc.kids.end(), kids.begin(), kids.end());
            rc.dirty = true;
        }
    }
    // phase 3 (:589-610): uniquely (re)assigned reads up each reference's own lineage
    for (u32 g = 0; g < G; ++g) {
        const u32 u2 = in->uniq_reads_count2 ? in->uniq_reads_count2[g] : 0;
        if (u2 == 0) continue;
        Node &n0 = nodes[lin[(u64)g * 8]];                       // default-inserted like operator[]
        n0.normalize();
        kids = n0.kids;
        for (u32 j = 1; j < 8; ++j) {
            Node &rc = nodes[lin[(u64)g * 8 + j]];
            rc.add_count(u2);
            rc.kids.push_back(g);
            rc.kids.insert(rc.kids.end(), kids.begin(), kids.end());
            rc.dirty = true;
        }
    }

    // write_abundance (:733-843)
    const u32 rk = in->rank, pr = in->rank + 1;
    const float R = (float)in->matches_count;
    std::vector<u32> taxa;
    taxa.reserve(nodes.size());
    for (auto &kv : nodes)
        if (kv.second.has_count) taxa.push_back(kv.first);
    std::sort(taxa.begin(), taxa.end());
    std::unordered_map<u32, float> pab, sab;
    std::unordered_map<u32, u32> pcnt, scnt;
    for (u32 t : taxa)
        if (rank_of(t) == pr) {
            const Node &nd = nodes[t];
            pab[t] = (float)nd.count / R * 100;                   // float(c)/(matches_count) * 100
            pcnt[t] = nd.count;
        }
    std::vector<slimm_profile_row> out;
    std::vector<u32> parents;
    float sum_ab = 0.0f;
    u32 sum_cnt = 0;
    for (u32 t : taxa) {
        if (rank_of(t) != rk) continue;
        Node &nd = nodes[t];
        nd.normalize();
        if (nd.kids.empty()) return SLIMM_GPU_EINVAL;
        u32 gl = 0;
        for (u32 k : nd.kids) gl += in->ref_len[k];                // u32 wrap (:785)
        gl /= (u32)nd.kids.size();
        const float cov = (float)(u32)(nd.count * in->avg_read_length) / gl;   // u32 product (:792)
        const float ab = (float)nd.count / R * 100;
        const u32 p = lin[(u64)nd.kids.back() * 8 + pr];           // lineage of the last child iterated
        if (sab.find(p) == sab.end()) { sab[p] = ab; scnt[p] = nd.count; parents.push_back(p); }
        else { sab[p] += ab; scnt[p] += nd.count; }
        if (ab < in->abundance_cut_off || cov < in->coverage_cut_off || !has_name(t)) continue;
        slimm_profile_row r;
        r.taxon = t; r.kind = 0; r.read_count = nd.count; r.first_child = nd.kids[0]; r.abundance = ab;
        out.push_back(r);
        sum_ab += ab;
        sum_cnt += nd.count;
    }
    std::sort(parents.begin(), parents.end());
    for (u32 p : parents) {
        auto pa = pab.find(p);
        const float uab = (pa == pab.end() ? 0.0f : pa->second) - sab[p];
        auto pc = pcnt.find(p);
        const u32 ucnt = (pc == pcnt.end() ? 0u : pc->second) - scnt[p];
        if (uab > in->abundance_cut_off && has_name(p)) {
            slimm_profile_row r;
            r.taxon = p; r.kind = 1; r.read_count = ucnt; r.abundance = uab; r.first_child = 0xFFFFFFFFu;
            auto it = nodes.find(p);
            if (p != 0 && it != nodes.end()) { it->second.normalize(); if (!it->second.kids.empty()) r.first_child = it->second.kids[0]; }
            out.push_back(r);
            sum_cnt += ucnt;
            sum_ab += uab;
        }
    }
    slimm_profile_row last;
    last.taxon = 0; last.kind = 2; last.first_child = 0xFFFFFFFFu;
    last.abundance = 100.0 - sum_ab;                               // double minus float (:835)
    last.read_count = in->matches_count - sum_cnt;                 // u32 wrap
    out.push_back(last);
    *n_out = out.size();
    for (u64 i = 0; i < out.size() && i < cap; ++i)
        if (rows) rows[i] = out[i];
    return SLIMM_GPU_OK;
}

// profile_host.cpp - host tail of the hot path: rank aggregation + abundances.
//
// Replaces phases 2 and 3 of slimm::get_reads_lca_count (reference src/slimm.hpp:560-610) and the
// numeric part of slimm::write_abundance (:733-843).  O(G + T) work on what the GPU stages produced;
// text formatting (lineage strings, TSV) stays with the caller.
//
// The reference keeps unordered_map<taxon, count> and unordered_map<taxon, std::set<ref>>.  Here every
// taxon that can occur is a value of the [G,8] lineage table, so a plan built once per database maps
// lineage slots to dense taxon indices and all per-sample state lives in flat arrays; sets of
// contributing references are plain vectors that are de-duplicated lazily with a stamp array (only set
// semantics, the smallest and the largest member are ever read: no sort), and only collected for taxa
// whose set is ever read (taxa with direct counts, the requested rank and its parent rank, strain level).
// The reference walks its snapshot of direct counts in libstdc++ hash order; here the order is
// ascending (rank, taxon) - identical results whenever the lineage table is tree-consistent
// (SURVEY.md appendix A8).
#include "profile_host.h"

#include <algorithm>
#include <cstring>

namespace slimm_host {

ProfilePlan::ProfilePlan(u32 n_refs, const u32 *ref_len, const u32 *lineage, u64 n_taxa, const u32 *taxa_id,
                         const uint8_t *taxa_rank, const uint8_t *taxa_has_name)
    : G(n_refs), len(ref_len, ref_len + n_refs), lin(lineage, lineage + (size_t)n_refs * 8)
{
    vals = lin;
    std::sort(vals.begin(), vals.end());
    vals.erase(std::unique(vals.begin(), vals.end()), vals.end());
    const size_t T = vals.size();
    slot_t.resize(lin.size());
    for (size_t s = 0; s < lin.size(); ++s) slot_t[s] = (u32)(std::lower_bound(vals.begin(), vals.end(), lin[s]) - vals.begin());
    rank.assign(T, 0);   // db.taxid__name[t] default-inserts (strain_lv, "") for unknown taxa
    named.assign(T, 0);
    for (u64 i = 0; i < n_taxa; ++i) {
        auto it = std::lower_bound(vals.begin(), vals.end(), taxa_id[i]);
        if (it != vals.end() && *it == taxa_id[i]) {
            rank[it - vals.begin()] = taxa_rank[i];
            named[it - vals.begin()] = taxa_has_name[i];
        }
    }
    count.assign(T, 0); direct.assign(T, 0); has_count.assign(T, 0); dirty.assign(T, 0); seen.assign(T, 0);
    kids.resize(T);
    pab.assign(T, 0.0f); sab.assign(T, 0.0f); pcnt.assign(T, 0); scnt.assign(T, 0); has_p.assign(T, 0); has_s.assign(T, 0);
    stamp.assign(G, 0); kmin.assign(T, 0xFFFFFFFFu); kmax.assign(T, 0);
    check_consistency();
}

int ProfilePlan::find(u32 taxon) const
{
    auto it = std::lower_bound(vals.begin(), vals.end(), taxon);
    return (it != vals.end() && *it == taxon) ? (int)(it - vals.begin()) : -1;
}

void ProfilePlan::touch(u32 t)
{
    if (!seen[t]) { seen[t] = 1; touched.push_back(t); }
}

void ProfilePlan::normalize(u32 t)
{
    if (!dirty[t]) return;
    std::vector<u32> &k = kids[t];
    if (++epoch == 0) { std::fill(stamp.begin(), stamp.end(), 0u); epoch = 1; }
    size_t m = 0;
    u32 lo = 0xFFFFFFFFu, hi = 0;
    for (u32 r : k)
        if (stamp[r] != epoch) { stamp[r] = epoch; k[m++] = r; lo = std::min(lo, r); hi = std::max(hi, r); }
    k.resize(m);
    kmin[t] = lo; kmax[t] = hi;
    dirty[t] = 0;
}

void ProfilePlan::begin()
{
    for (u32 t : touched) {
        count[t] = 0; direct[t] = 0; has_count[t] = 0; dirty[t] = 0; seen[t] = 0; kids[t].clear();
        pab[t] = sab[t] = 0.0f; pcnt[t] = scnt[t] = 0; has_p[t] = has_s[t] = 0;
    }
    touched.clear();
    snapshot.clear();
}

void ProfilePlan::add_direct(u32 t, u32 c)
{
    touch(t);
    if (!direct[t] && c) snapshot.push_back(t);
    count[t] += c; direct[t] += c; has_count[t] = 1;   // increment_or_initialize (src/misc.hpp:138-147)
}

void ProfilePlan::add_child(u32 t, u32 ref)
{
    touch(t);
    std::vector<u32> &k = kids[t];
    if (k.empty() || k.back() != ref) { k.push_back(ref); dirty[t] = 1; }
}

int ProfilePlan::finish(const u32 *uniq_reads_count2, u32 matches_count, u32 avg_read_length, float coverage_cut_off,
                        float abundance_cut_off, u32 rk, std::vector<slimm_profile_row> &out)
{
    out.clear();
    last_failed = 0;
    if (rk < 1 || rk > 6) return SLIMM_GPU_EINVAL;
    const u32 pr = rk + 1;
    // a taxon's reference set is only ever read if it has a direct count (phase 2 reads it live), is at the
    // requested rank or its parent rank (write_abundance) or at strain level (phase 3 reads children[lineage[0]])
    auto need_kids = [&](u32 t) { return direct[t] != 0 || rank[t] == rk || rank[t] == pr || rank[t] == 0; };
    // phase 2 (:560-586): push each direct count and its children up the first child's lineage
    std::sort(snapshot.begin(), snapshot.end(), [&](u32 a, u32 b) { return rank[a] != rank[b] ? rank[a] < rank[b] : a < b; });
    std::vector<u32> tmp;
    for (u32 t : snapshot) {
        normalize(t);
        if (kids[t].empty()) return SLIMM_GPU_EINVAL;            // .at() would throw in the reference
        tmp = kids[t];                                           // copied before the loop (:575)
        const u32 f = kmin[t], c = direct[t];
        for (u32 j = rank[t] + 1; j < 8; ++j) {
            const u32 rc = slot_t[(size_t)f * 8 + j];
            touch(rc);
            count[rc] += c; has_count[rc] = 1;
            if (need_kids(rc)) {
                kids[rc].insert(kids[rc].end(), tmp.begin(), tmp.end()); dirty[rc] = 1;
                if (kids[rc].size() > 4 * (size_t)G + 64) normalize(rc);
            }
        }
    }
    // phase 3 (:589-610): uniquely (re)assigned reads up each reference's own lineage
    for (u32 g = 0; g < G; ++g) {
        const u32 u2 = uniq_reads_count2 ? uniq_reads_count2[g] : 0;
        if (u2 == 0) continue;
        const u32 t0 = slot_t[(size_t)g * 8];
        normalize(t0);
        const std::vector<u32> k0 = kids[t0];                      // copied once before the loop (:594)
        for (u32 j = 1; j < 8; ++j) {
            const u32 rc = slot_t[(size_t)g * 8 + j];
            touch(rc);
            count[rc] += u2; has_count[rc] = 1;
            if (need_kids(rc)) {
                std::vector<u32> &k = kids[rc];
                k.push_back(g);
                k.insert(k.end(), k0.begin(), k0.end());
                dirty[rc] = 1;
                if (k.size() > 4 * (size_t)G + 64) normalize(rc);
            }
        }
    }
    // write_abundance (:733-843); rows in ascending taxon order
    std::sort(touched.begin(), touched.end());
    const float R = (float)matches_count;
    for (u32 t : touched)
        if (has_count[t] && rank[t] == pr) { pab[t] = (float)count[t] / R * 100; pcnt[t] = count[t]; has_p[t] = 1; }
    std::vector<u32> parents;
    float sum_ab = 0.0f;
    u32 sum_cnt = 0;
    const size_t n_touched = touched.size();   // touch() below may append parents: iterate the fixed prefix
    for (size_t ti = 0; ti < n_touched; ++ti) {
        const u32 t = touched[ti];
        if (!has_count[t] || rank[t] != rk) continue;
        normalize(t);
        const std::vector<u32> &k = kids[t];
        if (k.empty()) return SLIMM_GPU_EINVAL;
        u32 gl = 0;
        for (u32 r : k) gl += len[r];                              // u32 wrap (:785)
        gl /= (u32)k.size();
        const float cov = (float)(u32)(count[t] * avg_read_length) / gl;   // u32 product (:792)
        const float ab = (float)count[t] / R * 100;
        const u32 p = slot_t[(size_t)kmax[t] * 8 + pr];            // lineage of the last child iterated (std::set order)
        if (!has_s[p]) { has_s[p] = 1; sab[p] = ab; scnt[p] = count[t]; parents.push_back(p); touch(p); }
        else { sab[p] += ab; scnt[p] += count[t]; }
        if (ab < abundance_cut_off || cov < coverage_cut_off || !named[t]) { ++last_failed; continue; }
        slimm_profile_row r;
        r.taxon = vals[t]; r.kind = 0; r.read_count = count[t]; r.first_child = kmin[t]; r.abundance = ab;
        out.push_back(r);
        sum_ab += ab;
        sum_cnt += count[t];
    }
    std::sort(parents.begin(), parents.end());
    for (u32 p : parents) {
        const float uab = (has_p[p] ? pab[p] : 0.0f) - sab[p];
        const u32 ucnt = (has_p[p] ? pcnt[p] : 0u) - scnt[p];
        if (uab > abundance_cut_off && named[p]) {
            slimm_profile_row r;
            r.taxon = vals[p]; r.kind = 1; r.read_count = ucnt; r.abundance = uab; r.first_child = 0xFFFFFFFFu;
            if (vals[p] != 0) { normalize(p); if (!kids[p].empty()) r.first_child = kmin[p]; }
            out.push_back(r);
            sum_cnt += ucnt;
            sum_ab += uab;
        }
    }
    slimm_profile_row last;
    last.taxon = 0; last.kind = 2; last.first_child = 0xFFFFFFFFu;
    last.abundance = 100.0 - sum_ab;                               // double minus float (:835)
    last.read_count = matches_count - sum_cnt;                     // u32 wrap
    out.push_back(last);
    return SLIMM_GPU_OK;
}

// A lineage table is tree-consistent when every taxon occupies one level only, db.taxid__name gives it exactly
// that level as its rank, it is never 0, and all references that share it share every level above it.  Then the
// aggregation of slimm::get_reads_lca_count / write_abundance does not depend on visit order or on which child is
// "first", and reduces to sums over the references below a taxon (k_rank_reduce).
void ProfilePlan::check_consistency()
{
    consistent = false;
    const size_t T = vals.size();
    std::vector<u32> level_of(T, 0xFFFFFFFFu), parent_of(T, 0xFFFFFFFFu);
    for (u32 g = 0; g < G; ++g)
        for (u32 l = 0; l < 8; ++l) {
            const u32 t = slot_t[(size_t)g * 8 + l];
            if (vals[t] == 0 || rank[t] != l) return;
            if (level_of[t] == 0xFFFFFFFFu) level_of[t] = l;
            else if (level_of[t] != l) return;
            if (l < 7) {
                const u32 up = slot_t[(size_t)g * 8 + l + 1];
                if (parent_of[t] == 0xFFFFFFFFu) parent_of[t] = up;
                else if (parent_of[t] != up) return;
            }
        }
    lvl_idx.assign((size_t)8 * G, 0);
    std::vector<u32> pos(T, 0);
    for (u32 l = 0; l < 8; ++l) lvl_taxa[l].clear();
    for (size_t t = 0; t < T; ++t)
        if (level_of[t] != 0xFFFFFFFFu) { pos[t] = (u32)lvl_taxa[level_of[t]].size(); lvl_taxa[level_of[t]].push_back((u32)t); }
    for (u32 g = 0; g < G; ++g)
        for (u32 l = 0; l < 8; ++l) lvl_idx[(size_t)l * G + g] = pos[slot_t[(size_t)g * 8 + l]];
    consistent = true;
}

int ProfilePlan::finish_from_aggregates(const u32 *agg, u32 matches_count, u32 avg_read_length, float coverage_cut_off,
                                        float abundance_cut_off, u32 rk, std::vector<slimm_profile_row> &out) const
{
    out.clear();
    last_failed = 0;
    if (!consistent || rk < 1 || rk > 6) return SLIMM_GPU_EINVAL;
    const u32 pr = rk + 1;
    const u32 *cnt_r = agg, *kn_r = agg + G, *klen_r = agg + 2 * (size_t)G, *kmin_r = agg + 3 * (size_t)G, *kmax_r = agg + 4 * (size_t)G;
    const u32 *cnt_p = agg + 5 * (size_t)G, *kn_p = cnt_p + G, *kmin_p = cnt_p + 3 * (size_t)G;
    const float R = (float)matches_count;
    const std::vector<u32> &tr = lvl_taxa[rk], &tp = lvl_taxa[pr];
    std::vector<float> sab_l(tp.size(), 0.0f);
    std::vector<u32> scnt_l(tp.size(), 0);
    std::vector<uint8_t> has_s_l(tp.size(), 0);
    float sum_ab = 0.0f;
    u32 sum_cnt = 0;
    for (size_t j = 0; j < tr.size(); ++j) {                      // write_abundance rank pass (:776-813), ascending taxon
        const u32 c = cnt_r[j];
        if (c == 0) continue;
        if (kn_r[j] == 0) return SLIMM_GPU_EINVAL;
        const u32 t = tr[j];
        const u32 gl = klen_r[j] / kn_r[j];                        // u32 sum (wraps) / count (:785-789)
        const float cov = (float)(u32)(c * avg_read_length) / gl;  // u32 product (:792)
        const float ab = (float)c / R * 100;
        const u32 p = lvl_idx[(size_t)pr * G + kmax_r[j]];         // parent of the last child iterated (std::set order)
        if (!has_s_l[p]) { has_s_l[p] = 1; sab_l[p] = ab; scnt_l[p] = c; }
        else { sab_l[p] += ab; scnt_l[p] += c; }
        if (ab < abundance_cut_off || cov < coverage_cut_off || !named[t]) { ++last_failed; continue; }
        slimm_profile_row r;
        r.taxon = vals[t]; r.kind = 0; r.read_count = c; r.first_child = kmin_r[j]; r.abundance = ab;
        out.push_back(r);
        sum_ab += ab;
        sum_cnt += c;
    }
    for (size_t j = 0; j < tp.size(); ++j) {                      // "<parent>*" rows (:816-831), ascending taxon
        if (!has_s_l[j]) continue;
        const u32 p = tp[j];
        const bool has_p = cnt_p[j] != 0;
        const float uab = (has_p ? (float)cnt_p[j] / R * 100 : 0.0f) - sab_l[j];
        const u32 ucnt = (has_p ? cnt_p[j] : 0u) - scnt_l[j];
        if (uab > abundance_cut_off && named[p]) {
            slimm_profile_row r;
            r.taxon = vals[p]; r.kind = 1; r.read_count = ucnt; r.abundance = uab;
            r.first_child = kn_p[j] ? kmin_p[j] : 0xFFFFFFFFu;
            out.push_back(r);
            sum_cnt += ucnt;
            sum_ab += uab;
        }
    }
    slimm_profile_row last;
    last.taxon = 0; last.kind = 2; last.first_child = 0xFFFFFFFFu;
    last.abundance = 100.0 - sum_ab;                               // double minus float (:835)
    last.read_count = matches_count - sum_cnt;                     // u32 wrap
    out.push_back(last);
    return SLIMM_GPU_OK;
}

}  // namespace slimm_host

extern "C" int slimm_profile_rows(const slimm_profile_input *in, slimm_profile_row *rows, uint64_t cap, uint64_t *n_out)
{
    if (!in || !n_out || in->rank < 1 || in->rank > 6 || !in->lineage || !in->ref_len) return SLIMM_GPU_EINVAL;
    slimm_host::ProfilePlan plan(in->n_refs, in->ref_len, in->lineage, in->n_taxa, in->taxa_id, in->taxa_rank, in->taxa_has_name);
    plan.begin();
    for (uint64_t i = 0; i < in->n_direct; ++i) {
        int t = plan.find(in->direct_taxon[i]);
        if (t < 0) return SLIMM_GPU_EINVAL;
        plan.add_direct((uint32_t)t, in->direct_count[i]);
    }
    for (uint64_t i = 0; i < in->n_children; ++i) {
        int t = plan.find(in->child_taxon[i]);
        if (t < 0 || in->child_ref[i] >= in->n_refs) return SLIMM_GPU_EINVAL;
        plan.add_child((uint32_t)t, in->child_ref[i]);
    }
    std::vector<slimm_profile_row> out;
    int rc = plan.finish(in->uniq_reads_count2, in->matches_count, in->avg_read_length, in->coverage_cut_off,
                         in->abundance_cut_off, in->rank, out);
    if (rc) return rc;
    *n_out = out.size();
    for (uint64_t i = 0; i < out.size() && i < cap; ++i)
        if (rows) rows[i] = out[i];
    return SLIMM_GPU_OK;
}

extern "C" int slimm_profile_db_is_tree_consistent(uint32_t n_refs, const uint32_t *lineage, uint64_t n_taxa, const uint32_t *taxa_id,
                                                   const uint8_t *taxa_rank, const uint8_t *taxa_has_name, int *out)
{
    if (!lineage || !out || (n_taxa && (!taxa_id || !taxa_rank || !taxa_has_name))) return SLIMM_GPU_EINVAL;
    std::vector<uint32_t> len(n_refs, 1);
    slimm_host::ProfilePlan plan(n_refs, len.data(), lineage, n_taxa, taxa_id, taxa_rank, taxa_has_name);
    *out = plan.consistent ? 1 : 0;
    return SLIMM_GPU_OK;
}

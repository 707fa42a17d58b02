/*
 * slimm_oracle.c - CPU restatement of SLIMM's profiling hot path on integer arrays.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under slimm_b200/ links, imports or calls this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and
 * there only as the checker or the CPU baseline.
 *
 * Parity pin: the reference (seqan/slimm 0.3.4) ships no tests or golden vectors for this path
 * ("parity unpinned" by the reference's own suite).  This restatement is pinned instead against
 * outputs of the reference binaries themselves, built from /root/reference by oracle/Makefile
 * into oracle/_ref/ and run by tests/golden/make_golden.py; the resulting _profile.tsv /
 * _raw.tsv / _coverage.tsv / -v stderr files are committed under tests/golden/ and
 * tests/test_oracle_golden.py diffs this code against every one of them.
 *
 * Each stage cites the reference lines it restates (paths relative to the reference root).
 * u32 arithmetic wraps; f32 is IEEE binary32, one rounding per operation (built with
 * -ffp-contract=off, no -ffast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LIN 8 /* LINAGE_LENGTH, src/misc.hpp taxa_ranks strain..superkingdom */

typedef struct {
    /* references (contigs in @SQ order), src/slimm.hpp:420-445 */
    uint32_t n_refs;
    const uint32_t *ref_len;  /* [G]   */
    const uint32_t *lineage;  /* [G*8] db.ac__taxid[acc(g)], zeros when the accession is unknown */
    uint32_t bin_width;       /* -w, or avg_read_length when 0 (src/slimm.hpp:412-413) */
    uint32_t avg_read_length; /* src/misc.hpp:509-522 */
    float cov_cut_off;        /* -cc */
    uint32_t min_reads;       /* -mr; 0 => 1 + (R-1)/10000 (src/slimm.hpp:458-459) */
    /* kept records in file order (unmapped / rID==-1 already dropped, src/slimm.hpp:197) */
    uint64_t n_records;
    const uint32_t *read_id;  /* dense id of qName + (".1"|".2") (src/slimm.hpp:204-208) */
    const uint32_t *ref_id;
    const int32_t *begin_pos; /* POS-1 */
} oracle_input;

typedef struct {
    uint32_t hits, n_reads, n_uniq, n_uniq2;                  /* hits_count, matches_count, uniq_matches_count(2) */
    uint32_t failed_by_cov, failed_by_uniq_cov, failed_by_min_read, n_valid, min_reads;
    float cut, ucut;
    uint64_t n_bins;          /* sum nb[g] */
    uint64_t n_pairs;         /* distinct (read, ref) */
    /* per reference [G] */
    uint32_t *nb, *reads_count, *uniq_reads_count, *uniq_reads_count2, *nz, *unz, *unz2;
    float *cp, *ucp;
    uint8_t *valid;
    uint64_t *bin_off;        /* [G+1] */
    uint32_t *cov, *uniq_cov, *uniq_cov2; /* [n_bins] */
    /* per read, indexed by read id, [max_read_id+1]; reads that never occur have n_targets 0 */
    uint32_t n_read_slots;
    uint32_t *read_n_targets; /* |targets(read)|                     */
    uint32_t *read_n_valid;   /* |S(read)|                            */
    uint32_t *read_assigned;  /* the sole survivor when |S|==1, else 0xFFFFFFFF */
    uint32_t *read_lca;       /* LCA taxon when |S|>=2, else 0xFFFFFFFF */
    /* LCA phase 1 (src/slimm.hpp:536-557): sparse results sorted by taxon (then ref) */
    uint64_t n_direct;
    uint32_t *direct_taxid, *direct_count;
    uint64_t n_child_pairs;
    uint32_t *child_taxid, *child_ref;
} oracle_output;

/* ---- A1: ingest (src/slimm.hpp:194-213, src/read_stat.hpp:116-135) ----------------------- */
static uint32_t bin_of(const oracle_input *in, uint64_t i)
{
    /* src/slimm.hpp:200-201: uint32 arithmetic (beginPos = -1 wraps), clamp to the contig
     * length, integer division by the bin width */
    uint32_t g = in->ref_id[i];
    uint32_t center = (uint32_t)in->begin_pos[i] + in->avg_read_length / 2u;
    if (center > in->ref_len[g]) center = in->ref_len[g];
    return center / in->bin_width;
}

/* ---- A4: quantile cut-off (src/misc.hpp:197-216) ---------------------------------------- */
static int cmp_f32(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

static float quantile_cut_off(float *v, uint32_t n, float q)
{
    if (n == 0) return 0.0f;
    volatile float total = 0.0f; /* std::accumulate(..., (float)0): left fold in index order */
    for (uint32_t i = 0; i < n; ++i) total = total + v[i];
    qsort(v, n, sizeof(float), cmp_f32);
    volatile float sub = 0.0f;
    uint32_t i = n - 1;
    while ((sub / total) < q && i > 0) { /* NaN (total == 0) compares false: loop not entered */
        sub = sub + v[i];
        --i;
    }
    return v[i];
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

void oracle_free(oracle_output *o)
{
    free(o->nb); free(o->reads_count); free(o->uniq_reads_count); free(o->uniq_reads_count2);
    free(o->nz); free(o->unz); free(o->unz2); free(o->cp); free(o->ucp); free(o->valid);
    free(o->bin_off); free(o->cov); free(o->uniq_cov); free(o->uniq_cov2);
    free(o->read_n_targets); free(o->read_n_valid); free(o->read_assigned); free(o->read_lca);
    free(o->direct_taxid); free(o->direct_count); free(o->child_taxid); free(o->child_ref);
    memset(o, 0, sizeof(*o));
}

/* returns 0 on success, 1 on bad input, 2 on allocation failure */
int oracle_run(const oracle_input *in, oracle_output *o)
{
    const uint32_t G = in->n_refs;
    const uint64_t N = in->n_records;
    memset(o, 0, sizeof(*o));
    if (in->bin_width == 0) return 1;
    for (uint64_t i = 0; i < N; ++i)
        if (in->ref_id[i] >= G) return 1;

    /* reference_contig / bins_coverage construction: nb = len / w + 1
     * (src/reference_contig.hpp:77-82,129-145) */
    o->nb = calloc(G ? G : 1, 4);
    o->bin_off = calloc((size_t)G + 1, 8);
    for (uint32_t g = 0; g < G; ++g) {
        o->nb[g] = in->ref_len[g] / in->bin_width + 1u;
        o->bin_off[g + 1] = o->bin_off[g] + o->nb[g];
    }
    o->n_bins = o->bin_off[G];
    size_t nbins = o->n_bins ? o->n_bins : 1;
    o->cov = calloc(nbins, 4); o->uniq_cov = calloc(nbins, 4); o->uniq_cov2 = calloc(nbins, 4);
    o->reads_count = calloc(G ? G : 1, 4); o->uniq_reads_count = calloc(G ? G : 1, 4);
    o->uniq_reads_count2 = calloc(G ? G : 1, 4);
    o->nz = calloc(G ? G : 1, 4); o->unz = calloc(G ? G : 1, 4); o->unz2 = calloc(G ? G : 1, 4);
    o->cp = calloc(G ? G : 1, 4); o->ucp = calloc(G ? G : 1, 4); o->valid = calloc(G ? G : 1, 1);
    if (!o->cov || !o->uniq_cov || !o->uniq_cov2) return 2;

    /* group records by read, keeping file order inside a read: stable counting sort on the id.
     * (The reference keys an unordered_map by the read-name string, src/slimm.hpp:204-211;
     * the dense id stands for that string.) */
    uint32_t max_id = 0;
    for (uint64_t i = 0; i < N; ++i) if (in->read_id[i] > max_id) max_id = in->read_id[i];
    const uint32_t slots = N ? max_id + 1u : 0u;
    o->n_read_slots = slots;
    uint64_t *rstart = calloc((size_t)slots + 2, 8);
    uint64_t *order = malloc((N ? N : 1) * 8);
    if (!rstart || !order) return 2;
    for (uint64_t i = 0; i < N; ++i) rstart[in->read_id[i] + 1]++;
    for (uint32_t r = 0; r < slots; ++r) rstart[r + 1] += rstart[r];
    {
        uint64_t *cur = malloc(((size_t)slots + 1) * 8);
        if (!cur) return 2;
        memcpy(cur, rstart, ((size_t)slots + 1) * 8);
        for (uint64_t i = 0; i < N; ++i) order[cur[in->read_id[i]]++] = i;
        free(cur);
    }
    o->hits = (uint32_t)N; /* ++hits_count per kept record, src/slimm.hpp:212 */

    /* A1 target lists: per read the distinct references in first-appearance order, each with the
     * bin of its FIRST record only - add_target iterates the targets by value, so the push_back
     * for a repeated (read, ref) hit lands in a temporary (src/read_stat.hpp:125-131). */
    uint32_t *tref = malloc((N ? N : 1) * 4), *tbin = malloc((N ? N : 1) * 4);
    uint64_t *tstart = calloc((size_t)slots + 1, 8);
    if (!tref || !tbin || !tstart) return 2;
    uint64_t P = 0;
    for (uint32_t r = 0; r < slots; ++r) {
        tstart[r] = P;
        for (uint64_t k = rstart[r]; k < rstart[r + 1]; ++k) {
            uint64_t i = order[k];
            uint32_t g = in->ref_id[i];
            int seen = 0;
            for (uint64_t t = tstart[r]; t < P; ++t) if (tref[t] == g) { seen = 1; break; }
            if (!seen) { tref[P] = g; tbin[P] = bin_of(in, i); ++P; }
        }
    }
    tstart[slots] = P;
    o->n_pairs = P;

    /* A2 coverage (src/slimm.hpp:219-257; is_uniq src/read_stat.hpp:72-75) */
    o->read_n_targets = calloc(slots ? slots : 1, 4);
    o->read_n_valid = calloc(slots ? slots : 1, 4);
    o->read_assigned = malloc((slots ? slots : 1) * 4);
    o->read_lca = malloc((slots ? slots : 1) * 4);
    for (uint32_t r = 0; r < slots; ++r) {
        uint64_t a = tstart[r], b = tstart[r + 1];
        o->read_n_targets[r] = (uint32_t)(b - a);
        o->read_assigned[r] = 0xFFFFFFFFu;
        o->read_lca[r] = 0xFFFFFFFFu;
        if (b == a) continue;
        o->n_reads++; /* matches_count = reads.size(), :257 */
        if (b - a == 1) {
            uint32_t g = tref[a];
            o->reads_count[g] += 1;           /* pos_count is always 1, see A1 */
            o->cov[o->bin_off[g] + tbin[a]] += 1;
            o->uniq_reads_count[g] += 1;
            o->uniq_cov[o->bin_off[g] + tbin[a]] += 1;
            o->n_uniq++;                      /* uniq_matches_count == uniq_hits_count */
        } else {
            for (uint64_t t = a; t < b; ++t) {
                o->reads_count[tref[t]] += 1;
                o->cov[o->bin_off[tref[t]] + tbin[t]] += 1;
            }
        }
    }

    /* A3 per-reference statistics (src/reference_contig.hpp:84-91,148-155) */
    for (uint32_t g = 0; g < G; ++g) {
        uint32_t nz = 0, unz = 0;
        for (uint64_t b = o->bin_off[g]; b < o->bin_off[g + 1]; ++b) {
            nz += o->cov[b] != 0;
            unz += o->uniq_cov[b] != 0;
        }
        o->nz[g] = nz; o->unz[g] = unz;
        o->cp[g] = (float)nz / (float)o->nb[g];
        o->ucp[g] = (float)unz / (float)o->nb[g];
    }

    /* A4 cut-offs (src/slimm.hpp:328-344,672-688): only references with unique reads take part,
     * in ascending reference order; both stay 0 when -cc >= 1 */
    o->cut = 0.0f; o->ucut = 0.0f;
    if (in->cov_cut_off < 1.0f && N > 0) {
        float *v = malloc((G ? G : 1) * 4);
        uint32_t n = 0;
        for (uint32_t g = 0; g < G; ++g) if (o->uniq_reads_count[g] > 0) v[n++] = o->cp[g];
        o->cut = quantile_cut_off(v, n, in->cov_cut_off);
        n = 0;
        for (uint32_t g = 0; g < G; ++g) if (o->uniq_reads_count[g] > 0) v[n++] = o->ucp[g];
        o->ucut = quantile_cut_off(v, n, in->cov_cut_off);
        free(v);
    }

    /* A5 valid set + -v counters (src/slimm.hpp:354-378; min_reads default :458-459) */
    o->min_reads = in->min_reads ? in->min_reads : (o->n_reads ? 1u + (o->n_reads - 1u) / 10000u : 0u);
    for (uint32_t g = 0; g < G && N > 0; ++g) {
        if (o->reads_count[g] == 0) continue;
        if (o->cp[g] >= o->cut && o->ucp[g] >= o->ucut) {
            o->valid[g] = 1; o->n_valid++;
        } else {
            if (o->ucp[g] < o->ucut) o->failed_by_uniq_cov++;
            if (o->reads_count[g] < o->min_reads) o->failed_by_min_read++;
            if (o->cp[g] < o->cut) o->failed_by_cov++;
        }
    }

    /* A6 reassign (src/slimm.hpp:380-391, src/read_stat.hpp:98-114) and
     * A7 LCA (src/slimm.hpp:516-557): level-wise agreement over the 8-slot lineages of the
     * surviving references; no agreeing level => slot 7 of the largest reference id
     * (std::set iterates ascending, so the last taxa_id assigned is that one). */
    uint64_t *dkeys = malloc((size_t)(slots ? slots : 1) * 8); /* taxon per multi read */
    uint64_t nd = 0;
    uint64_t *ckeys = malloc((size_t)(P ? P : 1) * 8);         /* (taxon << 32 | ref) */
    uint64_t nc = 0;
    if (!dkeys || !ckeys) return 2;
    for (uint32_t r = 0; r < slots && N > 0; ++r) {
        uint64_t a = tstart[r], b = tstart[r + 1];
        uint32_t ns = 0, sole = 0, sole_bin = 0, gmax = 0;
        for (uint64_t t = a; t < b; ++t)
            if (o->valid[tref[t]]) {
                if (ns == 0) { sole = tref[t]; sole_bin = tbin[t]; }
                if (ns == 0 || tref[t] > gmax) gmax = tref[t];
                ++ns;
            }
        o->read_n_valid[r] = ns;
        if (ns == 1) {
            o->uniq_reads_count2[sole] += 1;
            o->n_uniq2++;
            o->uniq_cov2[o->bin_off[sole] + sole_bin] += 1;
            o->read_assigned[r] = sole;
        } else if (ns >= 2) {
            uint32_t lca = in->lineage[(size_t)gmax * LIN + 7];
            for (int l = 0; l < LIN; ++l) {
                int all_same = 1;
                uint32_t v0 = in->lineage[(size_t)sole * LIN + l];
                for (uint64_t t = a; t < b; ++t)
                    if (o->valid[tref[t]] && in->lineage[(size_t)tref[t] * LIN + l] != v0) { all_same = 0; break; }
                if (all_same) { lca = v0; break; }
            }
            o->read_lca[r] = lca;
            dkeys[nd++] = lca;
            for (uint64_t t = a; t < b; ++t)
                if (o->valid[tref[t]]) ckeys[nc++] = ((uint64_t)lca << 32) | tref[t];
        }
    }
    for (uint32_t g = 0; g < G; ++g) {
        uint32_t unz2 = 0;
        for (uint64_t b = o->bin_off[g]; b < o->bin_off[g + 1]; ++b) unz2 += o->uniq_cov2[b] != 0;
        o->unz2[g] = unz2;
    }
    /* taxon_id__read_count (direct part) and taxon_id__children as sorted sparse lists */
    qsort(dkeys, nd, 8, cmp_u64);
    qsort(ckeys, nc, 8, cmp_u64);
    o->direct_taxid = malloc((nd ? nd : 1) * 4); o->direct_count = malloc((nd ? nd : 1) * 4);
    for (uint64_t i = 0; i < nd; ++i) {
        if (o->n_direct && o->direct_taxid[o->n_direct - 1] == (uint32_t)dkeys[i]) {
            o->direct_count[o->n_direct - 1]++;
        } else {
            o->direct_taxid[o->n_direct] = (uint32_t)dkeys[i];
            o->direct_count[o->n_direct] = 1;
            o->n_direct++;
        }
    }
    o->child_taxid = malloc((nc ? nc : 1) * 4); o->child_ref = malloc((nc ? nc : 1) * 4);
    for (uint64_t i = 0; i < nc; ++i) {
        if (i && ckeys[i] == ckeys[i - 1]) continue;
        o->child_taxid[o->n_child_pairs] = (uint32_t)(ckeys[i] >> 32);
        o->child_ref[o->n_child_pairs] = (uint32_t)ckeys[i];
        o->n_child_pairs++;
    }
    free(dkeys); free(ckeys); free(tref); free(tbin); free(tstart); free(rstart); free(order);
    return 0;
}

/* ---- A10 raw-only folds (src/reference_contig.hpp:191-207, src/misc.hpp:285-289) -------- */
/* depth = f32 left fold of float(bin) over the bins divided by the bin count; 0 when no bin is
 * set.  Exposed so tests can compare _raw.tsv depth columns. */
float oracle_cov_depth(const uint32_t *bins, uint32_t nb)
{
    uint32_t nz = 0;
    for (uint32_t i = 0; i < nb; ++i) nz += bins[i] != 0;
    if (nz == 0) return 0.0f;
    volatile float s = 0.0f;
    for (uint32_t i = 0; i < nb; ++i) s = s + (float)bins[i];
    return s / (float)nb; /* vSum / v.size(): size_t -> float conversion of the divisor */
}

/* abundance columns of _raw.tsv (src/slimm.hpp:259-302): count[] is reads_count or
 * uniq_reads_count, denom is hits_count or uniq_hits_count */
void oracle_raw_abundance(const uint32_t *count, const uint32_t *ref_len, uint32_t G, uint32_t denom, float *out)
{
    volatile float total = 0.0f;
    for (uint32_t g = 0; g < G; ++g) {
        if (count[g] > 0) {
            out[g] = (float)(uint32_t)(count[g] * 100u) / (float)denom;
            total = total + out[g] / (float)ref_len[g];
        } else {
            out[g] = 0.0f;
        }
    }
    for (uint32_t g = 0; g < G; ++g)
        if (count[g] > 0) out[g] = (out[g] * 100.0f) / (total * (float)ref_len[g]);
}

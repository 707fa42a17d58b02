// slimm_main.cpp - drop-in `slimm` command line on top of the B200 hot path (libslimm_gpu.so).
//
// Same arguments, database format, output file names and TSV contents as the reference front end
// (reference src/slimm.cpp:60-204, slimm::get_profiles src/slimm.hpp:395-496, the writers :733-943 and
// src/file_helper.hpp).  The per-file driver keeps the reference's order of work; the five hot-path member
// functions are replaced by the C ABI of include/slimm_gpu.h, the SeqAn record loop by the threaded decoder
// in alignment_decoder.hpp.  There is no CPU fallback: without a CUDA device the run fails.
//
// Deliberate differences ("front end"): profile rows are written in ascending taxon order (the
// reference prints them in libstdc++ hash-iteration order); `-r strains` / `-r superkingdom` are rejected (the
// reference indexes past its rank vector there); in directory mode every file gets its own cut-offs (the
// reference keeps the first file's cached cut-offs).
#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/slimm_gpu.h"
#include "alignment_decoder.hpp"
#include "sldb.hpp"

using namespace slimm_fe;

static const char *RANK_NAMES[9] = {"strain", "species", "genus", "family", "order", "class", "phylum", "superkingdom", "intermidiate"};
static const char *RANK_SHORT[9] = {"r", "s", "g", "f", "o", "c", "p", "k", "i"};

struct Options {
    float cov_cut_off = 0.95f, abundance_cut_off = 0.01f;
    uint32_t bin_width = 0, min_reads = 0;
    bool verbose = false, is_directory = false, raw_output = false, coverage_output = false;
    std::string rank = "species", input_path, output_prefix, database_path;
    // extras of this front end (not in the reference)
    std::string dump_records;        // --dump-records FILE: decode only, write the SoA (test hook, needs no GPU)
    bool exact_ids = false;          // --exact-ids: always assign read ids through the name table (skip the grouped-input fast path)
    int threads = 0, device = 0;
    int gpus = 1;                    // --gpus N: reads sharded over N devices of this host (slimm_gpu_run_sharded_local)
};

struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), tl = t0;
    double lap() { auto n = std::chrono::steady_clock::now(); double s = std::chrono::duration<double>(n - tl).count(); tl = n; return s; }
    double elapsed() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

// ---- file names (reference src/file_helper.hpp:48-123) -------------------------------------------
static std::string get_file_name(const std::string &s) { return s.substr(s.find_last_of("/\\") + 1); }
static std::string get_directory(const std::string &s) { return s.substr(0, s.find_last_of("/\\")); }

static std::string get_tsv_file_name(const std::string &output_prefix, const std::string &input_path, const std::string &suffix)
{
    const std::string dir_name = get_directory(output_prefix);
    std::string file_name = get_file_name(output_prefix);
    if (file_name.empty()) {
        file_name = get_file_name(input_path);
        const size_t dot = file_name.find_last_of(".");
        if ((file_name.find(".sam") != std::string::npos && file_name.find(".sam") == dot) ||
            (file_name.find(".bam") != std::string::npos && file_name.find(".bam") == dot))
            file_name.replace(dot, 4, "");
    }
    return dir_name + "/" + file_name + suffix + ".tsv";
}

static std::vector<std::string> get_bam_files_in_directory(const std::string &directory)
{
    std::vector<std::string> out;
    DIR *dir = opendir(directory.c_str());
    if (!dir) return out;
    struct dirent *ent;
    struct stat st;
    while ((ent = readdir(dir)) != NULL) {
        const std::string file_name = ent->d_name, full = directory + "/" + file_name;
        if (file_name[0] == '.' || stat(full.c_str(), &st) == -1 || (st.st_mode & S_IFDIR) != 0) continue;
        if (full.find(".sam") == full.find_last_of(".") || full.find(".bam") == full.find_last_of(".")) out.push_back(full);
    }
    closedir(dir);
    // The reference takes the files in readdir() order, which no two file systems agree on - and its -w / -mr defaults stick from the
    // first file on.  Sorted by name the run is reproducible, and samples mapped against the same references (same @SQ lines) sit
    // next to each other more often than not, so the GPU context is reused
    std::sort(out.begin(), out.end());
    return out;
}

// accession = contig name up to the first whitespace, '.' or '|' (get_accession_id, reference src/misc.hpp:415-422)
static std::string get_accession_id(const std::string &name)
{
    size_t i = 0;
    while (i < name.size() && !isspace((unsigned char)name[i]) && name[i] != '.' && name[i] != '|') ++i;
    return name.substr(0, i);
}

// ---- command line (reference src/slimm.cpp:60-180) -----------------------------------------------
static void print_help()
{
    std::cout << "slimm - Species Level Identification of Microbes from Metagenomes (B200 hot path)\n\n"
                 "SYNOPSIS\n    slimm [OPTIONS] \"DB\" \"IN\"\n\n"
                 "OPTIONS\n"
                 "    -h, --help                    Display this help message.\n"
                 "    --version                     Display version information.\n"
                 "    -o, --output-prefix PREFIX    output path prefix.\n"
                 "    -w, --bin-width INT           Set the width of a single bin in neuclotides. Default: 0 (average read length).\n"
                 "    -mr, --min-reads INT          Minimum number of matching reads to consider a reference present. Default: 0.\n"
                 "    -r, --rank STR                The taxonomic rank of identification. One of species, genus, family, order,\n"
                 "                                  class, phylum. Default: species.\n"
                 "    -cc, --cov-cut-off DOUBLE     the quantile of coverages to use as a cutoff. In range [0.0..1.0]. Default: 0.95.\n"
                 "    -ac, --abundance-cut-off DOUBLE  do not report abundances below this value. In range [0.0..10.0]. Default: 0.01.\n"
                 "    -d, --directory               Input is a directory.\n"
                 "    -ro, --raw-output             Output raw reference statstics\n"
                 "    -co, --coverage-output        Output raw coverage statstics\n"
                 "    -v, --verbose                 Enable verbose output.\n"
                 "    --threads INT                 host decode threads (default: all cores); --device INT  CUDA device (default 0)\n"
                 "    --exact-ids                   always assign read ids through the read-name table (skip the grouped-input fast path)\n"
                 "    --gpus INT                    shard the reads over INT GPUs of this host, starting at --device (profile-only runs)\n";
}

static bool parse_number(const std::string &s, double &v)
{
    char *end = nullptr;
    v = strtod(s.c_str(), &end);
    return end && *end == 0 && !s.empty();
}

// 0: run, 1: error (exit 1), 2: help / version shown (exit 0)
static int parse_command_line(int argc, char **argv, Options &o)
{
    std::vector<std::string> pos;
    bool o_set = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i], val;
        bool has_val = false;
        if (a.size() > 1 && a[0] == '-' && !(a.size() > 1 && (isdigit((unsigned char)a[1]) || a[1] == '.'))) {
            const size_t eq = a.find('=');
            if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; }
            auto value = [&](std::string &out) {
                if (has_val) { out = val; return true; }
                if (i + 1 >= argc) { std::cerr << "slimm: option requires an argument -- " << a << "\n"; return false; }
                out = argv[++i];
                return true;
            };
            std::string s;
            double d;
            if (a == "-h" || a == "--help") { print_help(); return 2; }
            else if (a == "--version") { std::cout << "slimm version: 0.3.4 (B200 hot path)\n"; return 2; }
            else if (a == "-o" || a == "--output-prefix") { if (!value(o.output_prefix)) return 1; o_set = true; }
            else if (a == "-w" || a == "--bin-width") {
                if (!value(s)) return 1;
                if (!parse_number(s, d) || d != std::floor(d)) { std::cerr << "slimm: the given value '" << s << "' cannot be casted to integer\n"; return 1; }
                o.bin_width = (uint32_t)(int64_t)d;
            } else if (a == "-mr" || a == "--min-reads") {
                if (!value(s)) return 1;
                if (!parse_number(s, d) || d != std::floor(d)) { std::cerr << "slimm: the given value '" << s << "' cannot be casted to integer\n"; return 1; }
                o.min_reads = (uint32_t)(int64_t)d;
            } else if (a == "-r" || a == "--rank") {
                if (!value(o.rank)) return 1;
            } else if (a == "-cc" || a == "--cov-cut-off") {
                if (!value(s)) return 1;
                if (!parse_number(s, d)) { std::cerr << "slimm: the given value '" << s << "' cannot be casted to double\n"; return 1; }
                if (d < 0.0 || d > 1.0) { std::cerr << "slimm: the given value '" << s << "' is not in the interval [0.0:1.0]\n"; return 1; }
                o.cov_cut_off = (float)d;
            } else if (a == "-ac" || a == "--abundance-cut-off") {
                if (!value(s)) return 1;
                if (!parse_number(s, d)) { std::cerr << "slimm: the given value '" << s << "' cannot be casted to double\n"; return 1; }
                if (d < 0.0 || d > 10.0) { std::cerr << "slimm: the given value '" << s << "' is not in the interval [0.0:10.0]\n"; return 1; }
                o.abundance_cut_off = (float)d;
            } else if (a == "-d" || a == "--directory") o.is_directory = true;
            else if (a == "-ro" || a == "--raw-output") o.raw_output = true;
            else if (a == "-co" || a == "--coverage-output") o.coverage_output = true;
            else if (a == "-v" || a == "--verbose") o.verbose = true;
            else if (a == "--threads") { if (!value(s)) return 1; o.threads = atoi(s.c_str()); }
            else if (a == "--device") { if (!value(s)) return 1; o.device = atoi(s.c_str()); }
            else if (a == "--gpus") { if (!value(s)) return 1; o.gpus = std::max(1, atoi(s.c_str())); }
            else if (a == "--dump-records") { if (!value(o.dump_records)) return 1; }
            else if (a == "--exact-ids") o.exact_ids = true;
            else { std::cerr << "slimm: illegal option -- " << a << "\n"; return 1; }
        } else pos.push_back(a);
    }
    if (pos.size() < 2) { std::cerr << "slimm: Not enough arguments were provided.\nTry 'slimm --help' for more information.\n"; return 1; }
    if (pos.size() > 2) { std::cerr << "slimm: Too many arguments were provided.\nTry 'slimm --help' for more information.\n"; return 1; }
    o.database_path = pos[0];
    o.input_path = pos[1];
    if (o.database_path.size() < 5 || o.database_path.compare(o.database_path.size() - 5, 5, ".sldb") != 0) {
        std::cerr << "slimm: the given path '" << o.database_path << "' does not have one of the valid file extensions [*.sldb]\n";
        return 1;
    }
    static const char *ok_ranks[] = {"species", "genus", "family", "order", "class", "phylum"};
    bool rank_ok = false;
    for (const char *r : ok_ranks) rank_ok |= o.rank == r;
    if (!rank_ok) {
        std::cerr << "slimm: the given value '" << o.rank << "' is not in the list of allowed values [species, genus, family, order, class, phylum]\n";
        return 1;
    }
    if (!o_set) o.output_prefix = o.input_path;
    return 0;
}

// ---- one input file ------------------------------------------------------------------------------
struct GpuError { int rc; std::string what; };

static void check(int rc, slimm_gpu_ctx *ctx, const char *what)
{
    if (rc != SLIMM_GPU_OK) {
        std::string msg = std::string(what) + ": " + slimm_gpu_strerror(rc);
        const char *d = ctx ? slimm_gpu_last_error(ctx) : "";
        if (d && *d) msg += std::string(" (") + d + ")";
        throw GpuError{rc, msg};
    }
}

static std::string lineage_string(uint32_t rank, const uint32_t *lin, const SlimmDb &db)   // reference src/slimm.hpp:690-708
{
    std::string out;
    for (uint32_t i = 7;; --i) {
        std::string nm = db.name_of(lin[i]);
        if (nm.empty()) nm = std::string("unknown_") + RANK_NAMES[i];
        if (!out.empty()) out += "|";
        out += std::string(RANK_SHORT[i]) + "__" + nm;
        if (i == rank) break;
    }
    return out;
}

struct FileState {
    uint32_t hits_count = 0;
};
static double g_db_seconds = 0;      // loading the .sldb (reported under -v)

// The GPU side of a run: one context per device (--gpus), kept across the files of a directory as long as the @SQ lines (and
// the outputs asked for) stay the same - the bin layout, the lineage tables and the device buffers are reused and only
// slimm_gpu_reset runs between samples (reference slimm::reset(), src/slimm.hpp:167-188, does the same for its members).
struct GpuSet {
    std::vector<slimm_gpu_ctx *> ctx;
    std::vector<std::string> names;
    std::vector<uint32_t> lengths;
    uint32_t flags = 0;
    int device = -1;
    bool matches(const AlignmentHeader &hd, uint32_t f, int dev, int n) const
    {
        return !ctx.empty() && (int)ctx.size() == n && flags == f && device == dev && names == hd.names && lengths == hd.lengths;
    }
    void destroy() { for (slimm_gpu_ctx *c : ctx) slimm_gpu_destroy(c); ctx.clear(); }
};

static bool dump_records_file(const Options &opt, AlignmentDecoder &dec, uint32_t avg, int threads)
{
    // test hook: header, average read length and the kept records as little-endian arrays
    std::vector<uint32_t> rid, ref;
    std::vector<int32_t> pos;
    const size_t cap = 1u << 16;
    std::vector<uint32_t> b_rid(cap), b_ref(cap);
    std::vector<int32_t> b_pos(cap);
    RecordBatch batch{b_rid.data(), b_ref.data(), b_pos.data(), 0, cap};
    DecodeStats st;
    std::string err;
    bool ok = false;
    for (int attempt = (opt.exact_ids || threads < 3) ? 1 : 0; attempt < 2; ++attempt) {   // grouped-input fast path first (it needs parse workers to pay off), the exact name table if a read comes back
        rid.clear(); ref.clear(); pos.clear();
        batch.n = 0;
        ok = dec.decode(threads, batch, [&](RecordBatch b) {
            rid.insert(rid.end(), b.read_id, b.read_id + b.n); ref.insert(ref.end(), b.ref_id, b.ref_id + b.n);
            pos.insert(pos.end(), b.begin_pos, b.begin_pos + b.n);
            b.n = 0;
            return b;
        }, st, err, attempt == 0);
        if (ok || !st.not_grouped) break;
    }
    if (!ok) { std::cerr << "slimm: " << err << "\n"; return false; }
    std::ofstream f(opt.dump_records, std::ios::binary);
    const uint64_t G = dec.header().names.size(), N = rid.size();
    const uint64_t hdr[5] = {G, N, avg, st.records_in_file, st.reads};
    f.write((const char *)hdr, sizeof hdr);
    f.write((const char *)dec.header().lengths.data(), (std::streamsize)(G * 4));
    for (const std::string &s : dec.header().names) { const uint32_t l = (uint32_t)s.size(); f.write((const char *)&l, 4); f.write(s.data(), l); }
    f.write((const char *)rid.data(), (std::streamsize)(N * 4));
    f.write((const char *)ref.data(), (std::streamsize)(N * 4));
    f.write((const char *)pos.data(), (std::streamsize)(N * 4));
    return (bool)f;
}

static void write_profile(const Options &opt, const std::string &input, slimm_gpu_ctx *ctx, const SlimmDb &db,
                          const std::vector<uint32_t> &lineage, uint32_t rank)
{
    const std::string path = get_tsv_file_name(opt.output_prefix, input, "_profile");
    std::ofstream os(path);
    os << "taxa_level\ttaxa_id\tlinage\tabundance\tread_count\n";
    uint64_t n = 0;
    check(slimm_gpu_profile(ctx, rank, opt.abundance_cut_off, nullptr, 0, &n), ctx, "slimm_gpu_profile");
    std::vector<slimm_profile_row> rows(n ? n : 1);
    check(slimm_gpu_profile(ctx, rank, opt.abundance_cut_off, rows.data(), rows.size(), &n), ctx, "slimm_gpu_profile");
    const uint32_t zeros[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t count = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const slimm_profile_row &r = rows[i];
        const uint32_t *lin = r.first_child != 0xFFFFFFFFu ? &lineage[(size_t)r.first_child * 8] : zeros;
        os << RANK_NAMES[rank] << "\t";
        if (r.kind == 0) {
            os << r.taxon << "\t" << lineage_string(rank, lin, db) << "\t" << (float)r.abundance << "\t" << r.read_count << "\n";
            ++count;
        } else if (r.kind == 1) {
            os << r.taxon << "*\t" << lineage_string(rank + 1, lin, db) << "|" << RANK_SHORT[rank] << "__" << db.name_of(r.taxon)
               << "_unclassified\t" << (float)r.abundance << "\t" << r.read_count << "\n";
        } else {
            os << "0*\t" << lineage_string(rank, zeros, db) << "\t" << r.abundance << "\t" << r.read_count << "\n";
        }
    }
    if (opt.verbose) {
        uint32_t failed = 0;
        slimm_gpu_profile_failed(ctx, &failed);
        std::cerr << "\n" << std::setw(4) << count << std::setw(15) << RANK_NAMES[rank] << " (" << failed << " bellow cutoff i.e. "
                  << opt.abundance_cut_off << ")";
    }
}

// mean of the bin heights as the reference folds it (reference_contig::_get_cov_depth, src/reference_contig.hpp:191-207,
// mean() src/misc.hpp:285-289): a sequential f32 sum of float(bin).  While the sum stays below 2^24 every partial sum is
// an exactly representable integer, so the fold equals float(sum of the bins); beyond that the bins are folded one by one.
static float cov_depth(slimm_gpu_ctx *ctx, int which, uint32_t ref, uint32_t nz, uint32_t sum, uint32_t n_bins, std::vector<uint32_t> &scratch)
{
    if (nz == 0) return 0.0f;
    float total;
    if (sum < (1u << 24)) total = (float)sum;
    else {
        scratch.resize(n_bins);
        check(slimm_gpu_fetch_bins(ctx, which, ref, scratch.data(), n_bins), ctx, "slimm_gpu_fetch_bins");
        total = 0.0f;
        for (uint32_t b = 0; b < n_bins; ++b) total += (float)scratch[b];
    }
    return total / (float)(size_t)n_bins;
}

static void write_raw_stat(const Options &opt, const std::string &input, slimm_gpu_ctx *ctx, const SlimmDb &db,
                           const std::vector<std::string> &accession, const std::vector<uint32_t> &taxa_id,
                           const std::vector<uint32_t> &ref_len, uint32_t bin_width, const slimm_gpu_summary &sm)
{
    const uint32_t G = (uint32_t)ref_len.size();
    std::vector<uint32_t> reads(G), ureads(G), ureads2(G), nz(G), unz(G), unz2(G);
    std::vector<float> cp(G), ucp(G);
    check(slimm_gpu_get_ref_stats(ctx, reads.data(), ureads.data(), ureads2.data(), nz.data(), unz.data(), cp.data(), ucp.data(), nullptr), ctx,
          "slimm_gpu_get_ref_stats");
    check(slimm_gpu_get_uniq2_nz(ctx, unz2.data()), ctx, "slimm_gpu_get_uniq2_nz");
    // abundances exactly as the reference folds them (src/slimm.hpp:259-302): u32 products wrap, f32 sums run in
    // ascending reference order
    std::vector<float> ab(G, 0.0f), uab(G, 0.0f);
    const uint32_t hits = sm.hits_count, uhits = sm.uniq_matches_count;
    float total = 0.0f;
    for (uint32_t g = 0; g < G; ++g)
        if (reads[g] > 0) { ab[g] = float(reads[g] * 100u) / hits; total += ab[g] / ref_len[g]; }
    for (uint32_t g = 0; g < G; ++g)
        if (reads[g] > 0) ab[g] = (ab[g] * 100) / (total * ref_len[g]);
    total = 0.0f;
    for (uint32_t g = 0; g < G; ++g)
        if (ureads[g] > 0) { uab[g] = float(ureads[g] * 100u) / uhits; total += uab[g] / ref_len[g]; }
    for (uint32_t g = 0; g < G; ++g)
        if (ureads[g] > 0) uab[g] = (uab[g] * 100) / (total * ref_len[g]);

    std::ofstream os(get_tsv_file_name(opt.output_prefix, input, "_raw"));
    os << "accesion\ttaxaid\tname\treads_count\tabundance\tuniq1_abundance\tuniq2_abundance\tgenome_length\tuniq1_reads_count\t"
          "uniq2_reads_count\tbins_count\tbins_count(>0)\tuniq1_bins_count(>0)\tuniq2_bins_count(>0)\tcoverage_depth\t"
          "uniq1_coverage_depth\tuniq2_coverage_depth\tcoverage(%)\tuniq1_coverage(%)\tuniq2_coverage(%)\n";
    std::vector<uint32_t> scratch;
    for (uint32_t g = 0; g < G; ++g) {
        std::string name = db.name_of(taxa_id[g]);
        if (name.empty()) name = "no_name_found";
        const uint32_t nb = ref_len[g] / bin_width + 1u;
        os << accession[g] << "\t" << taxa_id[g] << "\t" << name << "\t" << reads[g] << "\t" << ab[g] << "\t" << uab[g] << "\t" << 0.0f
           << "\t" << ref_len[g] << "\t" << ureads[g] << "\t" << ureads2[g] << "\t" << nb << "\t" << nz[g] << "\t" << unz[g] << "\t" << unz2[g]
           << "\t" << cov_depth(ctx, 0, g, nz[g], reads[g], nb, scratch) << "\t" << cov_depth(ctx, 1, g, unz[g], ureads[g], nb, scratch) << "\t"
           << cov_depth(ctx, 2, g, unz2[g], ureads2[g], nb, scratch) << "\t" << cp[g] << "\t" << ucp[g] << "\t" << float(unz2[g]) / nb << "\n";
    }
}

static void write_coverage(const Options &opt, const std::string &input, slimm_gpu_ctx *ctx, const SlimmDb &db,
                           const std::vector<std::string> &accession, const std::vector<uint32_t> &lineage,
                           const std::vector<uint32_t> &ref_len, uint32_t bin_width)
{
    const uint32_t G = (uint32_t)ref_len.size();
    std::vector<uint8_t> valid(G);
    check(slimm_gpu_get_ref_stats(ctx, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, valid.data()), ctx, "slimm_gpu_get_ref_stats");
    static const char *suffix[3] = {"_coverage", "_uniq_coverage", "_uniq_coverage2"};
    std::ofstream os[3];
    for (int k = 0; k < 3; ++k) os[k].open(get_tsv_file_name(opt.output_prefix, input, suffix[k]));
    std::vector<uint32_t> bins;
    for (uint32_t g = 0; g < G; ++g) {
        if (!valid[g]) continue;
        const uint32_t nb = ref_len[g] / bin_width + 1u;
        bins.resize(nb);
        for (int k = 0; k < 3; ++k) {
            os[k] << accession[g];
            for (int l = 0; l < 8; ++l) os[k] << "," << db.name_of(lineage[(size_t)g * 8 + l]);
            check(slimm_gpu_fetch_bins(ctx, k, g, bins.data(), nb), ctx, "slimm_gpu_fetch_bins");
            for (uint32_t b = 0; b < nb; ++b) os[k] << "," << bins[b];
            os[k] << "\n";
        }
    }
}

// slimm::get_profiles (reference src/slimm.hpp:395-496) for one file; returns false when the file could not be read
static bool get_profiles(Options &opt, const SlimmDb &db, const std::string &input, uint32_t index, uint32_t n_files, FileState &fs, GpuSet &gs)
{
    Timer watch;
    std::cerr << "\nReading " << index + 1 << " of " << n_files << " files ... (" << get_file_name(input) << ")\n"
              << "=================================================================\n";
    AlignmentDecoder dec;
    std::string err;
    if (!dec.open(input, err)) { std::cerr << err << "\n"; return false; }
    uint32_t avg_read_length = 0;
    if (!dec.avg_read_length(100000, avg_read_length, err)) { std::cerr << "slimm: " << input << ": " << err << "\n"; exit(1); }
    if (opt.bin_width == 0) opt.bin_width = avg_read_length;      // stays set for the following files, as in the reference
    const int threads = opt.threads > 0 ? opt.threads : (int)std::max(1u, std::thread::hardware_concurrency());
    if (!opt.dump_records.empty()) { if (!dump_records_file(opt, dec, avg_read_length, threads)) exit(1); return true; }

    const AlignmentHeader &hd = dec.header();
    const uint32_t G = (uint32_t)hd.names.size();
    if (G == 0) { std::cerr << "slimm: " << input << " has no reference sequences in its header\n"; exit(1); }
    std::cerr << "Intializing coverages for all reference genome ... ";
    std::vector<std::string> accession(G);
    std::vector<uint32_t> taxa_id(G, 0), lineage((size_t)G * 8, 0);
    for (uint32_t g = 0; g < G; ++g) {                            // reference src/slimm.hpp:428-445
        accession[g] = get_accession_id(hd.names[g]);
        auto it = db.ac__taxid.find(accession[g]);
        if (it != db.ac__taxid.end()) { taxa_id[g] = it->second[0]; memcpy(&lineage[(size_t)g * 8], it->second.data(), 32); }
    }
    const uint32_t flags = (opt.raw_output || opt.coverage_output) ? SLIMM_GPU_KEEP_UNIQ_COV2 : SLIMM_GPU_SKIP_BINS;   // profile-only runs never read the bins back
    const int n_gpus = opt.gpus;
    if (n_gpus > 1 && (opt.raw_output || opt.coverage_output)) {
        std::cerr << "\nslimm: --gpus works for profile-only runs (the bins of -ro / -co stay on the GPU that owns their histogram slice)\n";
        exit(1);
    }
    Timer ct;
    const bool reuse = gs.matches(hd, flags, opt.device, n_gpus);
    if (!reuse) {
        gs.destroy();
        slimm_gpu_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.n_refs = G; cfg.ref_len = hd.lengths.data(); cfg.lineage = lineage.data(); cfg.bin_width = opt.bin_width;
        cfg.avg_read_length = avg_read_length; cfg.flags = flags;
        for (int r = 0; r < n_gpus; ++r) {
            cfg.device = opt.device + r;
            slimm_gpu_ctx *c = nullptr;
            const int rc = slimm_gpu_create(&cfg, &c);
            if (rc != SLIMM_GPU_OK) {
                std::cerr << "\nslimm: cannot set up the GPU hot path on device " << cfg.device << ": " << slimm_gpu_strerror(rc);
                if (c && *slimm_gpu_last_error(c)) std::cerr << " (" << slimm_gpu_last_error(c) << ")";
                std::cerr << "\n";
                if (c) slimm_gpu_destroy(c);
                gs.destroy();
                exit(1);
            }
            gs.ctx.push_back(c);
        }
        gs.names = hd.names; gs.lengths = hd.lengths; gs.flags = flags; gs.device = opt.device;
    }
    slimm_gpu_ctx *ctx = gs.ctx[0];                               // after the run every context answers with the global results
    bool ok = true;
    double ctx_sec = 0, gsec = 0, dsec = 0;
    try {
        for (int r = 0; r < n_gpus; ++r) {
            if (reuse) check(slimm_gpu_reset(gs.ctx[r], opt.bin_width, avg_read_length), gs.ctx[r], "slimm_gpu_reset");
            if (n_gpus > 1) check(slimm_gpu_set_shard(gs.ctx[r], (uint32_t)r, (uint32_t)n_gpus), gs.ctx[r], "slimm_gpu_set_shard");
        }
        if (!reuse) {
            std::vector<uint32_t> t_id; std::vector<uint8_t> t_rank, t_named;
            for (const auto &kv : db.taxid__name) { t_id.push_back(kv.first); t_rank.push_back(kv.second.first); t_named.push_back(!kv.second.second.empty()); }
            for (slimm_gpu_ctx *c : gs.ctx) check(slimm_gpu_set_taxa(c, t_id.size(), t_id.data(), t_rank.data(), t_named.data()), c, "slimm_gpu_set_taxa");
        }
        ctx_sec = ct.elapsed();
        std::cerr << "[" << watch.lap() << " secs]" << std::endl;

        std::cerr << "Analysing alignments, reads and references ....... ";
        // decode threads fill pinned struct-of-arrays batches; each full batch is uploaded asynchronously while the next fills.
        // Several GPUs: batch k goes to device k mod N, cut at the last read boundary (the tail moves into the next batch), so
        // every read lives on one device and every device sees non-decreasing read ids
        const size_t cap = 1u << 22;
        const int NB = 3;
        void *pin[NB] = {nullptr, nullptr, nullptr};
        RecordBatch batches[NB];
        for (int b = 0; b < NB; ++b) {
            check(slimm_gpu_host_alloc(&pin[b], cap * 12), ctx, "slimm_gpu_host_alloc");
            batches[b].read_id = (uint32_t *)pin[b]; batches[b].ref_id = batches[b].read_id + cap;
            batches[b].begin_pos = (int32_t *)(batches[b].ref_id + cap); batches[b].cap = cap; batches[b].n = 0;
        }
        int cur = 0;
        uint64_t pushed = 0, turn = 0;
        DecodeStats st;
        Timer dt;
        bool dec_ok = false;
        auto sync_all = [&] { for (slimm_gpu_ctx *c : gs.ctx) check(slimm_gpu_sync_uploads(c), c, "slimm_gpu_sync_uploads"); };
        // Mapper output is grouped by read: ids by counting runs, checked in parallel by the parse workers.  If a read name
        // comes back after its run has ended (coordinate-sorted input), what was pushed is dropped and the file is decoded
        // again through the exact name table.
        for (int attempt = (opt.exact_ids || threads < 3) ? 1 : 0; attempt < 2; ++attempt) {
            cur = 0; pushed = 0; turn = 0;
            batches[0].n = 0;
            dec_ok = dec.decode(threads, batches[0], [&](RecordBatch b) {
                size_t cut = b.n;                                  // records of this batch that go out now
                if (n_gpus > 1 && b.n == b.cap) {                  // a full batch (the last one never is): keep the last read together
                    while (cut > 0 && b.read_id[cut - 1] == b.read_id[b.n - 1]) --cut;
                    if (cut == 0) cut = b.n;                       // one read fills the whole batch: it stays on this device, no turn taken
                }
                slimm_gpu_ctx *to = gs.ctx[turn % (uint64_t)n_gpus];
                check(slimm_gpu_push(to, b.read_id, b.ref_id, b.begin_pos, cut), to, "slimm_gpu_push");
                if (!(n_gpus > 1 && cut == b.n && b.n == b.cap)) ++turn;   // (a batch that is one single read keeps the turn: its next records follow it)
                pushed += cut;
                const int nxt = (cur + 1) % NB;
                if (pushed >= (uint64_t)(NB - 1) * cap) sync_all();   // the buffer about to be refilled is free again
                const size_t tail = b.n - cut;
                if (tail) {
                    memcpy(batches[nxt].read_id, b.read_id + cut, tail * 4); memcpy(batches[nxt].ref_id, b.ref_id + cut, tail * 4);
                    memcpy(batches[nxt].begin_pos, b.begin_pos + cut, tail * 4);
                }
                batches[nxt].n = tail;
                cur = nxt;
                return batches[cur];
            }, st, err, attempt == 0);
            if (dec_ok || !st.not_grouped) break;
            sync_all();
            for (slimm_gpu_ctx *c : gs.ctx) check(slimm_gpu_reset(c, 0, 0), c, "slimm_gpu_reset");
            for (int r = 0; r < n_gpus && n_gpus > 1; ++r) check(slimm_gpu_set_shard(gs.ctx[r], (uint32_t)r, (uint32_t)n_gpus), gs.ctx[r], "slimm_gpu_set_shard");
            if (opt.verbose) std::cerr << "\n  (input is not grouped by read: decoding again with the exact read-name table) ";
        }
        if (!dec_ok) throw GpuError{SLIMM_GPU_EINVAL, input + ": " + err};
        sync_all();
        dsec = dt.elapsed();
        for (int b = 0; b < NB; ++b) slimm_gpu_host_free(pin[b]);
        fs.hits_count = (uint32_t)st.records_kept;
        if (st.records_kept == 0) {
            std::cerr << "[" << watch.lap() << " secs]" << std::endl;
            std::cerr << "[WARNING] No mapped reads found in BAM file!" << std::endl;
            return true;
        }
        Timer gt;
        if (n_gpus > 1) check(slimm_gpu_run_sharded_local(gs.ctx.data(), (uint32_t)n_gpus, opt.cov_cut_off, opt.min_reads, st.records_kept), ctx, "slimm_gpu_run_sharded_local");
        else {
            check(slimm_gpu_coverage(ctx), ctx, "slimm_gpu_coverage");
            check(slimm_gpu_filter(ctx, opt.cov_cut_off, opt.min_reads), ctx, "slimm_gpu_filter");
            check(slimm_gpu_assign(ctx), ctx, "slimm_gpu_assign");
        }
        slimm_gpu_summary sm;
        check(slimm_gpu_get_summary(ctx, &sm), ctx, "slimm_gpu_get_summary");
        gsec = gt.elapsed();
        std::cerr << "[" << watch.lap() << " secs]" << std::endl;
        if (opt.min_reads == 0) opt.min_reads = sm.min_reads;     // reference src/slimm.hpp:458-459 (persists across files)
        if (opt.verbose) {
            std::vector<uint32_t> reads(G);
            check(slimm_gpu_get_ref_stats(ctx, reads.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr), ctx, "slimm_gpu_get_ref_stats");
            uint32_t matched_ref_length = 0;
            for (uint32_t g = 0; g < G; ++g) if (reads[g] > 0) matched_ref_length += hd.lengths[g];
            std::cerr << "  " << sm.hits_count << " records processed." << std::endl;
            std::cerr << "    " << sm.matches_count << " matching reads" << std::endl;
            std::cerr << "    " << sm.uniq_matches_count << " uniquily matching reads" << std::endl;
            std::cerr << "  references with reads = " << sm.reference_count << std::endl;
            std::cerr << "  expected bins coverage = " << float(avg_read_length * sm.matches_count) / matched_ref_length << std::endl;
            std::cerr << "  bins coverage cut-off = " << sm.coverage_cut_off << " (" << opt.cov_cut_off << " quantile)\n";
            std::cerr << "  uniq bins coverage cut-off = " << sm.uniq_coverage_cut_off << " (" << opt.cov_cut_off << " quantile)\n\n";
            std::cerr << "  decode: " << st.records_in_file << " records, " << st.reads << " reads in " << dsec << " s ("
                      << st.records_in_file / dsec / 1e6 << " M records/s, " << threads << " host threads); GPU stages: " << gsec * 1e3 << " ms\n";
            std::cerr << "  phases: database " << g_db_seconds << " s (once), GPU context + tables " << ctx_sec << " s" << (reuse ? " (context reused)" : "")
                      << ", decode + upload " << dsec << " s, GPU stages " << gsec << " s on " << n_gpus << " GPU(s)\n";
        }
        std::cerr << "Filtering unlikely sequences ..................... ";
        std::cerr << "[" << watch.lap() << " secs]" << std::endl;
        if (opt.verbose) {
            std::cerr << "  " << sm.n_valid << " passed the threshould coverage.\n";
            std::cerr << "  " << sm.failed_by_cov << " ref's couldn't pass the coverage threshould.\n";
            std::cerr << "  " << sm.failed_by_uniq_cov << " ref's couldn't pass the uniq coverage threshould.\n";
            std::cerr << "  uniquily matching reads increased from " << sm.uniq_matches_count << " to " << sm.uniq_matches_count2 << "\n\n";
        }
        if (opt.raw_output) {
            std::cerr << "Writing features to a file ....................... ";
            write_raw_stat(opt, input, ctx, db, accession, taxa_id, hd.lengths, opt.bin_width, sm);
            std::cerr << "[" << watch.lap() << " secs]" << std::endl;
        }
        if (opt.coverage_output) {
            std::cerr << "Writing coverage profiles to a file ....................... ";
            write_coverage(opt, input, ctx, db, accession, lineage, hd.lengths, opt.bin_width);
            std::cerr << "[" << watch.lap() << " secs]" << std::endl;
        }
        std::cerr << "Assigning reads to Least Common Ancestor (LCA) ... ";
        std::cerr << "[" << watch.lap() << " secs]" << std::endl;
        std::cerr << "Writing taxnomic profile(s) ...................... ";
        uint32_t rank = 1;
        for (uint32_t r = 1; r <= 6; ++r) if (opt.rank == RANK_NAMES[r]) rank = r;
        write_profile(opt, input, ctx, db, lineage, rank);
        if (opt.verbose) std::cerr << "\n.................................................. ";
        std::cerr << "[" << watch.lap() << " secs]" << std::endl;
        std::cerr << "[Done!] File took " << watch.elapsed() << " secs to process.\n";
    } catch (const GpuError &e) {
        std::cerr << "\nslimm: " << e.what << "\n";
        ok = false;
    }
    if (!ok) { gs.destroy(); exit(1); }
    return true;
}

int main(int argc, char **argv)
{
    Options opt;
    const int pr = parse_command_line(argc, argv, opt);
    if (pr != 0) return pr == 1;
    Timer watch;
    std::vector<std::string> inputs;
    if (opt.is_directory) {
        inputs = get_bam_files_in_directory(opt.input_path);
        if (opt.verbose) std::cerr << inputs.size() << " SAM/BAM Files found under the directory: " << opt.input_path << "!\n";
    } else {
        if (access(opt.input_path.c_str(), 0) == 0) inputs.push_back(opt.input_path);
        else { std::cerr << opt.input_path << " is not a file use -d option for a directory.\n"; return 1; }
    }
    SlimmDb db;
    std::string err;
    if (opt.dump_records.empty() && !load_sldb(opt.database_path, db, err)) { std::cerr << "slimm: " << err << "\n"; return 1; }
    g_db_seconds = watch.elapsed();
    uint32_t total_hits = 0;
    GpuSet gs;
    for (uint32_t n = 0; n < inputs.size(); ++n) {
        FileState fs;
        get_profiles(opt, db, inputs[n], n, (uint32_t)inputs.size(), fs, gs);
        total_hits += fs.hits_count;
    }
    gs.destroy();
    std::cerr << "\n*****************************************************************\n";
    std::cerr << total_hits << " SAM/BAM alignment records are proccessed.\n";
    std::cerr << "Taxonomic profiles are written to: \n   " << get_directory(opt.output_prefix) << "\n";
    std::cerr << "Total time elapsed: " << watch.elapsed() << " secs\n";
    return 0;
}

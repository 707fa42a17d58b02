// Test program (tests/test_frontend_cpu.py::test_batched_parser_equals_line_parser): the chunk parser of the decode pipeline
// (parse_sam_range: eight lines at a time, prefetched contig look-ups, SWAR numbers) against the line parser (parse_sam_line)
// on random pieces of a SAM file, two thirds of them with random bytes overwritten (tabs, newlines, digits, garbage): same
// records, same counters, same error text.  Also parse_u32_swar against parse_u32 on random fields.
#include "alignment_decoder.hpp"
#include <iostream>
#include <fstream>
#include <random>
using namespace slimm_fe;
static void by_line(const char *p, const char *e, const NameIndex &idx, ParsedChunk &out)
{
    while (p < e) { const char *nl = (const char *)memchr(p, '\n', e - p); const char *le = nl ? nl : e; if (!parse_sam_line(p, le, idx, out, nullptr)) return; p = le + 1; }
}
int main(int argc, char **argv)
{
    std::ifstream f(argv[1], std::ios::binary);
    std::string data(40u << 20, 0);
    f.read(&data[0], data.size()); data.resize(f.gcount());
    std::vector<std::string> names; size_t p0 = 0;
    while (p0 < data.size() && data[p0] == '@') { size_t nl = data.find('\n', p0); if (data.compare(p0, 3, "@SQ") == 0) { size_t a = data.find("SN:", p0) + 3, b = data.find('\t', a); names.push_back(data.substr(a, b - a)); } p0 = nl + 1; }
    NameIndex idx; idx.build(names);
    std::mt19937_64 rng(7);
    long bad = 0, errs = 0;
    for (int it = 0; it < 3000; ++it) {
        size_t a = p0 + rng() % (data.size() - p0 - 70000), len = 100 + rng() % 60000;
        a = data.find('\n', a) + 1;
        std::string chunk = data.substr(a, len);
        int muts = it % 3 == 0 ? 0 : (int)(rng() % 6);
        for (int m = 0; m < muts; ++m) { size_t at = rng() % chunk.size(); const char c[] = {'\t', '\n', '9', 'x', '\r', '*', '@', ' ', '0'}; chunk[at] = c[rng() % sizeof c]; }
        ParsedChunk A, B;
        by_line(chunk.data(), chunk.data() + chunk.size(), idx, A);
        parse_sam_range(chunk.data(), chunk.data() + chunk.size(), idx, B);
        bool same = A.n == B.n && A.error == B.error && A.n_records == B.n_records;
        for (size_t i = 0; same && i < A.n; ++i) same = A.hash[i] == B.hash[i] && A.ref[i] == B.ref[i] && A.pos[i] == B.pos[i] && A.key_len[i] == B.key_len[i] && memcmp(A.keys.data() + A.key_off[i], B.keys.data() + B.key_off[i], A.key_len[i]) == 0;
        if (!A.error.empty()) ++errs;
        if (!same) { ++bad; if (bad < 4) std::cout << "MISMATCH it " << it << " n " << A.n << " vs " << B.n << " err '" << A.error << "' vs '" << B.error << "' recs " << A.n_records << " " << B.n_records << "\n"; }
    }
    long bad_num = 0;
    for (int it = 0; it < 2000000; ++it) {
        char buf[64]; memset(buf, 'x', sizeof buf);
        int n = 1 + rng() % 11;
        std::string fld;
        for (int k = 0; k < n; ++k) { int r = rng() % 40; fld.push_back(r < 36 ? char('0' + r % 10) : (r == 36 ? '/' : r == 37 ? ':' : r == 38 ? ' ' : char(rng() % 256))); }
        int off = rng() % 12;
        memcpy(buf + off, fld.data(), n);
        uint32_t x = 1, y = 2;
        const bool rx = parse_u32(buf + off, buf + off + n, x), ry = parse_u32_swar(buf + off, buf + off + n, buf, y);
        if (rx != ry || (rx && x != y)) ++bad_num;
    }
    std::cout << "bad " << bad << " bad_numbers " << bad_num << " (cases with an error: " << errs << ")\n";
    return bad || bad_num ? 1 : 0;
}

// sldb.hpp - reader of the SLIMM database file written by the reference's slimm_build.
//
// The reference stores `slimm_database` (src/misc.hpp:77-100) through cereal's BinaryOutputArchive
// (save_slimm_database, src/misc.hpp:178-185).  That archive has no header or version; on disk it is, little
// endian (cereal types/unordered_map.hpp, types/string.hpp, types/vector.hpp, types/tuple.hpp):
//   u64 n;  n x { u64 len; char accession[len]; u64 k; u32 lineage[k] }          ac__taxid (k == 8)
//   u64 m;  m x { u32 taxid; u32 rank (taxa_ranks, 0..8); u64 len; char name[len] }   taxid__name
#pragma once
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace slimm_fe {

struct SlimmDb {
    std::unordered_map<std::string, std::array<uint32_t, 8>> ac__taxid;
    std::unordered_map<uint32_t, std::pair<uint8_t, std::string>> taxid__name;   // (rank, name)

    const std::string &name_of(uint32_t taxid) const   // db.taxid__name[t] default-constructs missing entries: name ""
    {
        static const std::string none;
        auto it = taxid__name.find(taxid);
        return it == taxid__name.end() ? none : it->second.second;
    }
};

static inline bool load_sldb(const std::string &path, SlimmDb &db, std::string &err)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { err = "Could not open " + path + "!"; return false; }
    std::vector<char> buf;
    char tmp[1 << 16];
    size_t got;
    while ((got = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    fclose(f);
    size_t off = 0;
    auto need = [&](size_t n) { return off + n <= buf.size(); };
    auto u64 = [&]() { uint64_t v; memcpy(&v, buf.data() + off, 8); off += 8; return v; };
    auto u32 = [&]() { uint32_t v; memcpy(&v, buf.data() + off, 4); off += 4; return v; };
    const char *bad = "is not a SLIMM database (truncated or not written by slimm_build)";
    if (!need(8)) { err = path + " " + bad; return false; }
    const uint64_t n = u64();
    for (uint64_t i = 0; i < n; ++i) {
        if (!need(8)) { err = path + " " + bad; return false; }
        const uint64_t len = u64();
        if (!need(len + 8)) { err = path + " " + bad; return false; }
        std::string acc(buf.data() + off, len);
        off += len;
        const uint64_t k = u64();
        if (!need(k * 4)) { err = path + " " + bad; return false; }
        std::array<uint32_t, 8> lin{};
        for (uint64_t j = 0; j < k; ++j) { const uint32_t v = u32(); if (j < 8) lin[j] = v; }
        db.ac__taxid[acc] = lin;
    }
    if (!need(8)) { err = path + " " + bad; return false; }
    const uint64_t m = u64();
    for (uint64_t i = 0; i < m; ++i) {
        if (!need(16)) { err = path + " " + bad; return false; }
        const uint32_t taxid = u32(), rank = u32();
        const uint64_t len = u64();
        if (!need(len)) { err = path + " " + bad; return false; }
        db.taxid__name[taxid] = std::make_pair((uint8_t)rank, std::string(buf.data() + off, len));
        off += len;
    }
    return true;
}

}  // namespace slimm_fe

// alignment_decoder.hpp - SAM / BAM decoding for the SLIMM hot path, in pipelined host threads.
//
// Replaces what the reference does with SeqAn in the record loop of slimm::analyze_alignments (reference
// src/slimm.hpp:194-213: readRecord, the unmapped / rID == -1 skip, read_name + ".1" / ".2") and in
// get_avg_read_length (src/misc.hpp:509-522), and turns the string-keyed `reads` map into dense read ids.
// Output: struct-of-arrays batches {read_id, ref_id, begin_pos} of the KEPT records in file order, which is
// what slimm_gpu_push takes.
//
//   reader (inflate workers for BGZF)  ->  framer (whole lines / whole BAM records)  ->  parse workers
//        ->  consumer on the calling thread (dense read ids by first appearance, fills the batch)
//
// Semantics kept from SeqAn 2.4 (include/seqan/bam_io/read_sam.h:255-396, read_bam.h:194-277):
//   SAM  rID = index of RNAME among the @SQ lines, "*" -> -1; beginPos = (int32)(uint32)POS - 1
//   BAM  rID = refID, beginPos = pos, names and lengths from the binary reference list
//   kept <=> !(flag & 4) && rID != -1;   key = QNAME + (flag & 0x40 ? ".1" : flag & 0x80 ? ".2" : "")
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "bytesource.hpp"
#include "pipeline.hpp"

namespace slimm_fe {

// ---- hashing ------------------------------------------------------------------------------------
static inline uint64_t mum(uint64_t a, uint64_t b)
{
    const __uint128_t r = (__uint128_t)a * b;
    return (uint64_t)r ^ (uint64_t)(r >> 64);
}
static inline uint64_t rd64u(const char *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint64_t rd32u(const char *p) { uint32_t v; memcpy(&v, p, 4); return v; }
// fixed-size loads only (no variable-length memcpy): whole 8-byte words, then the LAST eight bytes once more (they overlap the
// words before when the length is not a multiple of 8); fewer than 8 bytes: two overlapping 4-byte loads or three single bytes
static inline uint64_t hash_bytes(const char *p, size_t n)
{
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xD6E8FEB86659FD93ull);
    if (n >= 8) {
        const char *last = p + n - 8;
        for (; p < last; p += 8) h = mum(h ^ rd64u(p), 0xA0761D6478BD642Full);
        h = mum(h ^ rd64u(last), 0xE7037ED1A0B428DBull);
    } else if (n >= 4) h = mum(h ^ (rd32u(p) | (rd32u(p + n - 4) << 32)), 0xE7037ED1A0B428DBull);
    else if (n) h = mum(h ^ ((uint64_t)(unsigned char)p[0] | ((uint64_t)(unsigned char)p[n >> 1] << 8) | ((uint64_t)(unsigned char)p[n - 1] << 16)), 0xE7037ED1A0B428DBull);
    else h = mum(h, 0xE7037ED1A0B428DBull);
    return h ^ (h >> 29);
}

// ---- header -------------------------------------------------------------------------------------
struct AlignmentHeader {
    std::vector<std::string> names;   // @SQ SN / BAM reference names, in order
    std::vector<uint32_t> lengths;    // @SQ LN / l_ref
};

// open-addressing map contig name -> index (RNAME lookup of the SAM text parser).  A slot is 32 bytes and holds the name
// itself when it is at most 23 bytes long (accessions are), so a look-up touches ONE cache line of a table that does not fit
// the L2 of a core (50 000 contigs: 4 MB) instead of three dependent ones (slot -> std::string -> characters).
class NameIndex {
public:
    void build(const std::vector<std::string> &names)
    {
        names_ = &names;
        size_t cap = 16;
        while (cap < names.size() * 2 + 2) cap <<= 1;
        slot_.assign(cap, Slot{});
        mask_ = cap - 1;
        for (size_t i = 0; i < names.size(); ++i) {
            const uint64_t h = hash_bytes(names[i].data(), names[i].size());
            size_t s = h & mask_;
            bool dup = false;
            while (slot_[s].idx1) {
                if ((*names_)[slot_[s].idx1 - 1] == names[i]) { dup = true; break; }   // duplicate @SQ: SeqAn's name cache resolves to the first
                s = (s + 1) & mask_;
            }
            if (dup) continue;
            Slot &sl = slot_[s];
            sl.idx1 = (uint32_t)i + 1;
            sl.tag = (uint32_t)(h >> 32);
            sl.len = names[i].size() <= INLINE ? (uint8_t)names[i].size() : (uint8_t)0xFF;
            if (names[i].size() <= INLINE) memcpy(sl.name, names[i].data(), names[i].size());
        }
    }
    void prefetch(uint64_t h) const { __builtin_prefetch(&slot_[h & mask_]); }
    int32_t find(const char *p, size_t n) const { return find(hash_bytes(p, n), p, n); }
    int32_t find(uint64_t h, const char *p, size_t n) const      // h = hash_bytes(p, n)
    {
        const uint32_t tag = (uint32_t)(h >> 32);
        size_t s = h & mask_;
        while (slot_[s].idx1) {
            const Slot &sl = slot_[s];
            if (sl.tag == tag) {
                if (sl.len != 0xFF) { if (sl.len == n && memcmp(sl.name, p, n) == 0) return (int32_t)(sl.idx1 - 1); }
                else {
                    const std::string &nm = (*names_)[sl.idx1 - 1];
                    if (nm.size() == n && memcmp(nm.data(), p, n) == 0) return (int32_t)(sl.idx1 - 1);
                }
            }
            s = (s + 1) & mask_;
        }
        return -1;
    }

private:
    static const size_t INLINE = 23;
    struct alignas(32) Slot { uint32_t idx1 = 0, tag = 0; uint8_t len = 0; char name[INLINE] = {}; };
    static_assert(sizeof(Slot) == 32, "one slot, half a cache line");
    const std::vector<std::string> *names_ = nullptr;
    std::vector<Slot> slot_;
    size_t mask_ = 0;
};

// ---- one parsed chunk ---------------------------------------------------------------------------
struct ParsedChunk {
    // arrays sized ahead of the records (ensure()), n of them in use: a record costs six plain stores, not six push_backs
    std::vector<uint64_t> hash;
    std::vector<uint32_t> key_off, key_len, ref;
    std::vector<int32_t> pos;
    std::vector<char> keys;          // read keys back to back (keys_used bytes)
    std::vector<uint8_t> new_run;    // grouped-input fast path: record i >= 1 starts another read than record i - 1
    std::vector<uint32_t> local_id;  // grouped-input fast path: runs started between record 0 and record i (inclusive)
    std::vector<uint64_t> heads, scratch;   // grouped-input fast path: the chunk's run-head hashes (kept here so that a recycled chunk brings its buffers)
    size_t n = 0, keys_used = 0;
    uint64_t n_records = 0;          // records seen, kept or not
    std::string error;
    void clear() { n = 0; keys_used = 0; new_run.clear(); n_records = 0; error.clear(); }
    void ensure(size_t records)
    {
        if (hash.size() < records) { hash.resize(records); key_off.resize(records); key_len.resize(records); ref.resize(records); pos.resize(records); }
        if (keys.size() < records * 24) keys.resize(records * 24);
    }
    bool same_key(size_t i, size_t j) const
    {
        return hash[i] == hash[j] && key_len[i] == key_len[j] && memcmp(keys.data() + key_off[i], keys.data() + key_off[j], key_len[i]) == 0;
    }
    size_t size() const { return n; }
    // readable_end (optional): bytes up to there may be read (the rest of the record): short names are copied as two fixed
    // 16-byte moves instead of a variable-length memcpy call
    void add(const char *name, size_t name_len, uint32_t flag, uint32_t rid, int32_t begin_pos, const char *readable_end = nullptr)
    {
        if (n == hash.size()) ensure(std::max<size_t>(1024, n * 2));
        if (keys_used + name_len + 34 > keys.size()) keys.resize(std::max(keys.size() * 2, keys_used + name_len + 34 + 4096));
        char *k = keys.data() + keys_used;
        if (name_len <= 32 && readable_end && name + 32 <= readable_end) { memcpy(k, name, 16); memcpy(k + 16, name + 16, 16); }
        else memcpy(k, name, name_len);
        size_t len = name_len;
        if (flag & 0x40u) { k[len++] = '.'; k[len++] = '1'; }               // reference src/slimm.hpp:205-208
        else if (flag & 0x80u) { k[len++] = '.'; k[len++] = '2'; }
        key_off[n] = (uint32_t)keys_used; key_len[n] = (uint32_t)len;
        hash[n] = hash_bytes(k, len);
        ref[n] = rid; pos[n] = begin_pos;
        ++n;
        keys_used += len;
    }
};

struct ParseJob {
    std::string head;                 // one record stitched together across a chunk boundary (may be empty)
    std::shared_ptr<Buffer> buf;
    size_t begin = 0, end = 0;        // whole records inside buf
};

// ---- SAM text -----------------------------------------------------------------------------------
static inline bool parse_u32(const char *p, const char *e, uint32_t &out)
{
    if (p == e) return false;
    uint64_t v = 0;
    if (e - p <= 9) {                                            // cannot overflow: one range test per digit is all
        for (; p < e; ++p) {
            const unsigned d = (unsigned)(unsigned char)*p - '0';
            if (d > 9) return false;
            v = v * 10 + d;
        }
    } else
        for (; p < e; ++p) {
            if (*p < '0' || *p > '9') return false;
            v = v * 10 + (uint64_t)(*p - '0');
            if (v > 0xFFFFFFFFull) return false;
        }
    out = (uint32_t)v;
    return true;
}

// The same for a field of at most eight characters when the eight bytes that END at e may be read (lo = start of the record):
// one 8-byte load, the bytes before the field replaced by '0', all digits tested and combined at once.
static inline bool parse_u32_swar(const char *p, const char *e, const char *lo, uint32_t &out)
{
    const size_t n = (size_t)(e - p);
    if (n == 0 || n > 8 || e - 8 < lo) return parse_u32(p, e, out);
    uint64_t v = rd64u(e - 8);                                  // first character of the field in the lowest of its bytes
    if (n < 8) v = (v >> (8 * (8 - n)) << (8 * (8 - n))) | (0x3030303030303030ull >> (8 * n));
    v -= 0x3030303030303030ull;
    if ((v | (v + 0x7676767676767676ull)) & 0x8080808080808080ull) return false;    // a byte above 9 (or below '0': it wrapped)
    v = v * 10 + (v >> 8);
    v = (((v & 0x000000FF000000FFull) * 0x000F424000000064ull) + (((v >> 16) & 0x000000FF000000FFull) * 0x0000271000000001ull)) >> 32;
    out = (uint32_t)v;
    return true;
}

// one alignment line [p, e) without its newline; seq_len (optional) receives the length of SEQ ("*" -> 0)
static inline bool parse_sam_line(const char *p, const char *e, const NameIndex &idx, ParsedChunk &out, uint32_t *seq_len)
{
    if (e > p && e[-1] == '\r') --e;
    if (p == e) return true;                                    // blank line
    if (*p == '@') return true;                                 // header line
    ++out.n_records;
    // the first four tabs: QNAME, FLAG, RNAME and POS end within the first few dozen bytes, one or two 16-byte compares find them
    const char *tabs[4] = {nullptr, nullptr, nullptr, nullptr};
    int nt = 0;
    const char *q = p;
#if defined(__SSE2__)
    const __m128i tab = _mm_set1_epi8('\t');
    while (nt < 4 && q + 16 <= e) {
        unsigned m = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i *>(q)), tab));
        while (m && nt < 4) { tabs[nt++] = q + __builtin_ctz(m); m &= m - 1; }
        q += 16;
    }
#endif
    if (nt < 4) {
        if (nt) q = std::max(q, tabs[nt - 1] + 1);              // (the vector loop may have stopped short of e)
        for (; q < e && nt < 4; ++q) if (*q == '\t') tabs[nt++] = q;
    }
    const char *t1 = tabs[0], *t2 = tabs[1], *t3 = tabs[2], *t4 = tabs[3];
    if (!t3) { out.error = "SAM record with fewer than 4 fields"; return false; }
    if (!t4) t4 = e;
    uint32_t flag = 0, pos1 = 0;
    if (!parse_u32(t1 + 1, t2, flag) || flag > 0xFFFFu) { out.error = "SAM FLAG is not a 16-bit number"; return false; }
    if (!parse_u32(t3 + 1, t4, pos1)) { out.error = "SAM POS is not a number"; return false; }
    if (seq_len) {                                              // SEQ is field 10
        const char *f = t4;
        int field = 4;
        while (f < e && field < 9) { f = (const char *)memchr(f + 1, '\t', e - f - 1); if (!f) { f = e; break; } ++field; }
        *seq_len = 0;
        if (f < e && field == 9) {
            const char *s = f + 1, *se = (const char *)memchr(s, '\t', e - s);
            if (!se) se = e;
            *seq_len = (se - s == 1 && *s == '*') ? 0u : (uint32_t)(se - s);
        }
    }
    const char *rn = t2 + 1;
    const size_t rn_len = (size_t)(t3 - rn);
    if ((flag & 4u) || (rn_len == 1 && *rn == '*')) return true;   // reference src/slimm.hpp:197-198
    const int32_t rid = idx.find(rn, rn_len);
    if (rid < 0) { out.error = "SAM record names a reference that is not in the header: " + std::string(rn, rn_len); return false; }
    out.add(p, (size_t)(t1 - p), flag, (uint32_t)rid, (int32_t)pos1 - 1);
    return true;
}

// The chunk parser of the decode pipeline: the same records as parse_sam_line line by line, but eight lines at a time - their
// fields are located and the RNAME slots of the contig index requested first, the look-ups (one cache line each in a table
// that does not fit a core's L2) and the stores follow when the lines have arrived.
static inline void parse_sam_range(const char *p, const char *e, const NameIndex &idx, ParsedChunk &out)
{
    struct Line { const char *p, *e, *t1, *t2, *t3, *t4; uint64_t h; };
    constexpr int B = 8;
    Line L[B];
#if defined(__SSE2__)
    const __m128i tab = _mm_set1_epi8('\t');
#endif
    while (p < e) {
        int k = 0;
        while (k < B && p < e) {
            const char *nl = (const char *)memchr(p, '\n', e - p);
            const char *le = nl ? nl : e, *next = le + 1;
            if (le > p && le[-1] == '\r') --le;
            if (p == le || *p == '@') { p = next; continue; }     // blank line, header line
            Line &l = L[k];
            l.p = p; l.e = le;
            const char *tabs[4] = {nullptr, nullptr, nullptr, nullptr};
            int nt = 0;
            const char *q = p;
#if defined(__SSE2__)
            while (nt < 4 && q + 16 <= le) {
                unsigned m = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i *>(q)), tab));
                while (m && nt < 4) { tabs[nt++] = q + __builtin_ctz(m); m &= m - 1; }
                q += 16;
            }
#endif
            if (nt < 4) {
                if (nt) q = std::max(q, tabs[nt - 1] + 1);
                for (; q < le && nt < 4; ++q) if (*q == '\t') tabs[nt++] = q;
            }
            l.t1 = tabs[0]; l.t2 = tabs[1]; l.t3 = tabs[2]; l.t4 = tabs[3] ? tabs[3] : le;
            if (l.t3) { l.h = hash_bytes(l.t2 + 1, (size_t)(l.t3 - l.t2 - 1)); idx.prefetch(l.h); }
            ++k;
            p = next;
        }
        for (int i = 0; i < k; ++i) {
            const Line &l = L[i];
            ++out.n_records;
            if (!l.t3) { out.error = "SAM record with fewer than 4 fields"; return; }
            uint32_t flag = 0, pos1 = 0;
            if (!parse_u32(l.t1 + 1, l.t2, flag) || flag > 0xFFFFu) { out.error = "SAM FLAG is not a 16-bit number"; return; }
            if (!parse_u32_swar(l.t3 + 1, l.t4, l.p, pos1)) { out.error = "SAM POS is not a number"; return; }
            const char *rn = l.t2 + 1;
            const size_t rn_len = (size_t)(l.t3 - rn);
            if ((flag & 4u) || (rn_len == 1 && *rn == '*')) continue;   // reference src/slimm.hpp:197-198
            const int32_t rid = idx.find(l.h, rn, rn_len);
            if (rid < 0) { out.error = "SAM record names a reference that is not in the header: " + std::string(rn, rn_len); return; }
            out.add(l.p, (size_t)(l.t1 - l.p), flag, (uint32_t)rid, (int32_t)pos1 - 1, l.e);
        }
    }
}

// ---- BAM ----------------------------------------------------------------------------------------
static inline uint32_t rd32(const char *p) { uint32_t v; memcpy(&v, p, 4); return v; }

// one record [p, p + 4 + block_size)
static inline bool parse_bam_record(const char *p, size_t n, uint32_t n_refs, ParsedChunk &out, uint32_t *seq_len)
{
    ++out.n_records;
    if (n < 36) { out.error = "BAM record shorter than its fixed part"; return false; }
    const int32_t ref_id = (int32_t)rd32(p + 4), pos = (int32_t)rd32(p + 8);
    const uint32_t l_name = (unsigned char)p[12];
    const uint32_t flag = rd32(p + 16) >> 16;                   // n_cigar_op (low half) | flag (high half)
    if (seq_len) *seq_len = rd32(p + 20);
    if (36 + (size_t)l_name > n || l_name == 0) { out.error = "BAM record with a bad read name length"; return false; }
    if ((flag & 4u) || ref_id < 0) return true;
    if ((uint32_t)ref_id >= n_refs) { out.error = "BAM record references a contig beyond the reference list"; return false; }
    out.add(p + 36, l_name - 1, flag, (uint32_t)ref_id, pos);
    return true;
}

static inline void parse_bam_range(const char *p, const char *e, uint32_t n_refs, ParsedChunk &out)
{
    while (p + 4 <= e) {
        const size_t bs = rd32(p);
        if (p + 4 + bs > e) { out.error = "BAM record crosses its chunk"; return; }
        if (!parse_bam_record(p, 4 + bs, n_refs, out, nullptr)) return;
        p += 4 + bs;
    }
}

// ---- dense read ids -----------------------------------------------------------------------------
// key string -> id by first appearance.  Open addressing on the 64-bit key hash; a hit is confirmed against
// the stored key, so two different names never share an id.
class ReadIdTable {
public:
    ReadIdTable() { slots_.assign(1u << 16, Slot{0, 0}); mask_ = slots_.size() - 1; }
    void prefetch(uint64_t h) const { __builtin_prefetch(&slots_[h & mask_]); }
    uint32_t size() const { return (uint32_t)key_off_.size(); }
    bool full() const { return key_off_.size() >= 0xFFFFFFFEull; }
    uint32_t lookup_or_insert(uint64_t h, const char *key, uint32_t len)
    {
        size_t s = h & mask_;
        for (;;) {
            Slot &sl = slots_[s];
            if (sl.id1 == 0) break;
            if (sl.hash == h) {
                const uint32_t id = sl.id1 - 1;
                if (key_len_[id] == len && memcmp(key_at(id), key, len) == 0) return id;
            }
            s = (s + 1) & mask_;
        }
        const uint32_t id = (uint32_t)key_off_.size();
        slots_[s] = Slot{h, id + 1};
        store_key(key, len);
        if ((size_t)(id + 1) * 10 > slots_.size() * 6) grow();
        return id;
    }

private:
    struct Slot { uint64_t hash; uint32_t id1; };
    static const size_t BLOCK = 1u << 24;
    const char *key_at(uint32_t id) const { const uint64_t o = key_off_[id]; return blocks_[o >> 24].get() + (o & (BLOCK - 1)); }
    void store_key(const char *key, uint32_t len)
    {
        if (blocks_.empty() || used_ + len > BLOCK) {
            blocks_.emplace_back(new char[std::max<size_t>(BLOCK, len)]);
            used_ = 0;
        }
        memcpy(blocks_.back().get() + used_, key, len);
        key_off_.push_back(((uint64_t)(blocks_.size() - 1) << 24) | used_);
        key_len_.push_back(len);
        used_ += len;
    }
    void grow()
    {
        std::vector<Slot> old;
        old.swap(slots_);
        slots_.assign(old.size() * 2, Slot{0, 0});
        mask_ = slots_.size() - 1;
        for (const Slot &o : old)
            if (o.id1) {
                size_t s = o.hash & mask_;
                while (slots_[s].id1) s = (s + 1) & mask_;
                slots_[s] = o;
            }
    }
    std::vector<Slot> slots_;
    size_t mask_ = 0;
    std::vector<std::unique_ptr<char[]>> blocks_;
    size_t used_ = 0;
    std::vector<uint64_t> key_off_;
    std::vector<uint32_t> key_len_;
};

// ---- grouped-input fast path ----------------------------------------------------------------------
// Mapper output is grouped by read: every read name forms ONE run of consecutive records.  Then the dense id of a
// record is simply the number of runs before it, and the serial name -> id table is not needed at all - provided no
// name ever starts a second run.  The parse workers check exactly that, in parallel: every run head's 64-bit key hash
// goes into this sharded set; a hash seen twice (a name that comes back, or a 2^-64 collision) clears `grouped` and
// the caller decodes again through the exact table.  Both paths give identical ids whenever the fast one applies.
class RunHeadSet {
public:
    RunHeadSet() : shards_(N_SHARDS) {}
    // expected number of run heads (reads), e.g. from the size of the file: the shard tables start large enough that they are
    // never rehashed (a rehash re-inserts every hash of the shard: cache misses all of them)
    void reserve(uint64_t expected)
    {
        size_t cap = 1u << 10;
        while (cap < expected / N_SHARDS * 2 + 16 && cap < (1u << 26)) cap <<= 1;
        for (Shard &sh : shards_) {
            std::lock_guard<std::mutex> lk(sh.m);
            if (sh.slots.size() < cap && sh.n == 0) { sh.slots.assign(cap, 0); publish(sh); }
        }
    }
    // true when h was already there
    bool insert(uint64_t h)
    {
        if (h == 0) h = EMPTY_ALIAS;
        Shard &sh = shards_[h >> (64 - SHARD_BITS)];
        std::lock_guard<std::mutex> lk(sh.m);
        return insert_locked(sh, h);
    }
    // A chunk's run heads at once: grouped by shard first, so that a shard is locked once per chunk.  The tables together are far
    // larger than any cache (16 bytes per read), so every insert is a miss: the slots of the shards LOOKAHEAD places further on
    // are requested while this shard's share goes in, which overlaps the misses instead of paying them one after the other.
    // true when any of the hashes was already there (or occurs twice here).
    bool insert_all(std::vector<uint64_t> &hs, std::vector<uint64_t> &scratch)
    {
        if (hs.empty()) return false;
        uint32_t count[N_SHARDS + 1] = {0};
        for (uint64_t &h : hs) { if (h == 0) h = EMPTY_ALIAS; ++count[(h >> (64 - SHARD_BITS)) + 1]; }
        for (size_t k = 0; k < N_SHARDS; ++k) count[k + 1] += count[k];
        scratch.resize(hs.size());
        uint32_t at[N_SHARDS];
        memcpy(at, count, sizeof at);
        for (uint64_t h : hs) scratch[at[h >> (64 - SHARD_BITS)]++] = h;
        bool dup = false;
        // every caller starts its round over the shards somewhere else: workers that finish their chunks together would otherwise
        // queue up behind each other's locks shard after shard
        const size_t first = (size_t)rot_.fetch_add(0x9E37u, std::memory_order_relaxed) & (N_SHARDS - 1);
        auto request = [&](size_t k) {                       // (the table may be replaced meanwhile: a prefetch is only a hint)
            const Shard &sh = shards_[k];
            const uint64_t *base = sh.pf_base.load(std::memory_order_relaxed);
            const size_t mask = sh.pf_mask.load(std::memory_order_relaxed);
            if (!base) return;
            for (uint32_t i = count[k]; i < count[k + 1]; ++i) __builtin_prefetch(base + ((scratch[i] * 0xD6E8FEB86659FD93ull >> 20) & mask), 1);
        };
        for (size_t kk = 0; kk < LOOKAHEAD; ++kk) request((first + kk) & (N_SHARDS - 1));
        for (size_t kk = 0; kk < N_SHARDS; ++kk) {
            const size_t k = (first + kk) & (N_SHARDS - 1);
            if (kk + LOOKAHEAD < N_SHARDS) request((first + kk + LOOKAHEAD) & (N_SHARDS - 1));
            if (count[k] == count[k + 1]) continue;
            Shard &sh = shards_[k];
            std::lock_guard<std::mutex> lk(sh.m);
            for (uint32_t i = count[k]; i < count[k + 1]; ++i) dup |= insert_locked(sh, scratch[i]);
        }
        return dup;
    }

private:
    static const int SHARD_BITS = 10;
    static const size_t N_SHARDS = 1u << SHARD_BITS;
    static const size_t LOOKAHEAD = 24;
    static const uint64_t EMPTY_ALIAS = 0x9E3779B97F4A7C15ull;   // 0 marks an empty slot
    struct Shard {
        std::mutex m; std::vector<uint64_t> slots; size_t n = 0;
        std::atomic<const uint64_t *> pf_base{nullptr}; std::atomic<size_t> pf_mask{0};   // where to prefetch (written under the lock)
    };
    static void publish(Shard &sh)
    {
        sh.pf_base.store(sh.slots.data(), std::memory_order_relaxed);
        sh.pf_mask.store(sh.slots.size() - 1, std::memory_order_relaxed);
    }
    static bool insert_locked(Shard &sh, uint64_t h)
    {
        if (sh.slots.empty()) { sh.slots.assign(1u << 10, 0); publish(sh); }
        size_t mask = sh.slots.size() - 1, s = (h * 0xD6E8FEB86659FD93ull >> 20) & mask;
        while (sh.slots[s]) {
            if (sh.slots[s] == h) return true;
            s = (s + 1) & mask;
        }
        sh.slots[s] = h;
        if (++sh.n * 10 > sh.slots.size() * 6) {
            std::vector<uint64_t> old;
            old.swap(sh.slots);
            sh.pf_base.store(nullptr, std::memory_order_relaxed);
            sh.slots.assign(old.size() * 2, 0);
            mask = sh.slots.size() - 1;
            for (uint64_t v : old)
                if (v) {
                    size_t t = (v * 0xD6E8FEB86659FD93ull >> 20) & mask;
                    while (sh.slots[t]) t = (t + 1) & mask;
                    sh.slots[t] = v;
                }
            publish(sh);
        }
        return false;
    }
    std::vector<Shard> shards_;
    std::atomic<uint32_t> rot_{0};
};

// ---- the decoder --------------------------------------------------------------------------------
struct RecordBatch {
    uint32_t *read_id = nullptr, *ref_id = nullptr;
    int32_t *begin_pos = nullptr;
    size_t n = 0, cap = 0;
};

struct DecodeStats {
    uint64_t records_in_file = 0, records_kept = 0, reads = 0;
    double seconds = 0.0;
    bool not_grouped = false;        // grouped-input fast path only: a read name came back after its run had ended - decode again, exact
};

class AlignmentDecoder {
public:
    static const size_t CHUNK = 4u << 20;

    bool open(const std::string &path, std::string &err)
    {
        path_ = path;
        if (!file_.open(path)) { err = "Could not open " + path + "!"; return false; }
        comp_ = detect_compression(file_);
        return read_header(err);
    }
    const AlignmentHeader &header() const { return header_; }
    bool is_bam() const { return is_bam_; }

    // get_avg_read_length (reference src/misc.hpp:509-522): mean SEQ length (integer division) over the first
    // `sample` records that carry a sequence.  Returns false when no record does (the reference divides by zero).
    bool avg_read_length(uint32_t sample, uint32_t &avg, std::string &err)
    {
        uint32_t count = 0, total = 0;
        ParsedChunk scratch;
        auto one = [&](const char *p, size_t n) {
            uint32_t sl = 0;
            const uint64_t before = scratch.n_records;
            const bool ok = is_bam_ ? parse_bam_record(p, n, (uint32_t)header_.names.size(), scratch, &sl)
                                    : parse_sam_line(p, p + n, names_, scratch, &sl);
            if (!ok) {   // an unknown RNAME does not stop the reference's length sampling; anything else is fatal in SeqAn too
                if (scratch.error.compare(0, 16, "SAM record names") != 0) { err = scratch.error; return false; }
                scratch.error.clear();
            }
            if (scratch.n_records != before && sl > 0) { total += sl; ++count; }
            if (scratch.size() > 4096) scratch.clear();
            return true;
        };
        std::unique_ptr<ChunkReader> rd = make_reader(file_, comp_, CHUNK, 1);
        bool ok = frame(*rd, [&](ParseJob &job) {
            if (!job.head.empty() && !one(job.head.data(), job.head.size())) return false;
            const char *p = job.buf ? job.buf->p + job.begin : nullptr, *e = job.buf ? job.buf->p + job.end : nullptr;
            while (p && p < e && count < sample) {
                size_t n;
                if (is_bam_) n = 4 + (size_t)rd32(p);
                else { const char *nl = (const char *)memchr(p, '\n', e - p); n = nl ? (size_t)(nl - p) : (size_t)(e - p); }
                if (!one(p, n)) return false;
                p += n + (is_bam_ ? 0 : 1);
            }
            return count < sample;
        }, err, sample);
        if (!ok && !err.empty()) return false;
        if (count == 0) { err = "no record with a sequence: the average read length is undefined"; return false; }
        avg = total / count;
        return true;
    }

    // Decodes the whole file.  sink(batch) is called on the calling thread with every filled batch (and the last,
    // partial one); it returns the batch to fill next (double buffering is the sink's business) with n = the records it
    // already holds (0, or the tail of a read the sink moved over so that a read never straddles two batches).
    // assume_grouped: dense ids by counting runs of equal names, verified in parallel (RunHeadSet).  When the file
    // turns out not to be grouped by read the call stops early with st.not_grouped set (and returns false with an
    // empty err): the caller discards what the sink received and calls decode again with assume_grouped = false.
    template <class Sink>
    bool decode(int n_threads, RecordBatch first, Sink &&sink, DecodeStats &st, std::string &err, bool assume_grouped = false)
    {
        st = DecodeStats();
        RunHeadSet heads;
        if (assume_grouped && comp_ == Compression::none) heads.reserve(std::min<uint64_t>(file_.n / 64, 64ull << 20));   // a guess from the file size (16 bytes per expected read, at most 1 GB); more reads than that rehash
        std::atomic<bool> came_back(false);
        n_threads = std::max(1, n_threads);
        const int n_parse = std::max(1, comp_ == Compression::bgzf ? n_threads / 2 : n_threads - 1);
        const int n_inflate = std::max(1, n_threads - n_parse);
        const uint32_t n_refs = (uint32_t)header_.names.size();
        OrderedStage<ParseJob, ParsedChunk> parse(n_parse, (size_t)n_parse * 3 + 2, [this, n_refs, assume_grouped, &heads, &came_back](ParseJob &job, ParsedChunk &out) {
            const size_t guess = (job.end - job.begin) / (is_bam_ ? 200 : 250) + 16;
            out.clear();                                      // (a recycled chunk keeps its arrays)
            out.ensure(guess);
            if (!job.head.empty()) {
                if (is_bam_) parse_bam_record(job.head.data(), job.head.size(), n_refs, out, nullptr);
                else parse_sam_line(job.head.data(), job.head.data() + job.head.size(), names_, out, nullptr);
            }
            if (out.error.empty() && job.buf) {
                if (is_bam_) parse_bam_range(job.buf->p + job.begin, job.buf->p + job.end, n_refs, out);
                else parse_sam_range(job.buf->p + job.begin, job.buf->p + job.end, names_, out);
            }
            if (assume_grouped && out.error.empty()) {        // run heads inside the chunk; record 0 is the consumer's business
                const size_t n = out.size();
                out.new_run.assign(n, 0);
                std::vector<uint64_t> &hs = out.heads, &scratch = out.scratch;
                hs.clear();
                hs.reserve(n);
                out.local_id.resize(n);
                uint32_t runs = 0;                              // runs that START inside the chunk behind record 0
                if (n) out.local_id[0] = 0;
                for (size_t i = 1; i < n; ++i) {
                    if (!out.same_key(i, i - 1)) { out.new_run[i] = 1; hs.push_back(out.hash[i]); ++runs; }
                    out.local_id[i] = runs;                     // the record's read id relative to record 0's: the consumer only adds a base
                }
                if (heads.insert_all(hs, scratch)) came_back.store(true, std::memory_order_relaxed);
            }
        });
        std::string frame_err;
        std::atomic<bool> stop(false);                        // the consumer gave up (not grouped, parse error, sink error): stop reading and inflating
        std::thread framer([&] {
            std::unique_ptr<ChunkReader> rd = make_reader(file_, comp_, CHUNK, n_inflate);
            frame(*rd, [&](ParseJob &job) {
                if (stop.load(std::memory_order_relaxed)) return false;
                parse.push(std::move(job));
                return !stop.load(std::memory_order_relaxed);
            }, frame_err, 0);
            parse.close();
        });
        // the framer is joined on every way out of this function, exceptions of the sink included (a joinable std::thread that is
        // destroyed ends the process)
        struct JoinGuard {
            std::thread &t; std::atomic<bool> &stop; decltype(parse) &stage;
            ~JoinGuard() { if (t.joinable()) { stop.store(true); stage.abort(); t.join(); } }
        } guard{framer, stop, parse};
        RecordBatch batch = first;
        ReadIdTable table;
        std::string prev_key;
        uint64_t prev_hash = 0;
        uint32_t prev_id = 0;
        bool have_prev = false;
        ParsedChunk ch;
        bool ok = true;
        uint64_t next_id = 0;                                 // grouped fast path: runs seen so far
        for (;;) {
            parse.recycle(std::move(ch));                     // the chunk consumed last: its arrays serve another job
            if (!parse.pop(ch)) break;
            if (!ch.error.empty()) { err = ch.error; ok = false; break; }
            st.records_in_file += ch.n_records;
            const size_t n = ch.size();
            if (assume_grouped) {
                if (came_back.load(std::memory_order_relaxed)) { st.not_grouped = true; ok = false; break; }
                if (n) {
                    // does the chunk continue the previous chunk's last read?  Everything else was settled by the worker: record i
                    // has the id of record 0 plus local_id[i], so the serial part of the decoder is one add and two copies per record
                    const char *key = ch.keys.data() + ch.key_off[0];
                    const bool fresh = !(have_prev && ch.hash[0] == prev_hash && prev_key.size() == ch.key_len[0] && memcmp(prev_key.data(), key, ch.key_len[0]) == 0);
                    if (fresh && heads.insert(ch.hash[0])) came_back.store(true, std::memory_order_relaxed);
                    if (fresh) ++next_id;
                    const uint64_t id0 = next_id - 1;
                    next_id = id0 + ch.local_id[n - 1] + 1;
                    if (next_id > 0xFFFFFFFEull) { err = "more than 2^32-2 distinct reads"; ok = false; break; }
                    for (size_t i = 0; i < n;) {
                        if (batch.n == batch.cap) batch = sink(batch);   // the sink hands back the batch to fill next (n = records it already holds)
                        const size_t m = std::min(n - i, batch.cap - batch.n);
                        uint32_t *rid = batch.read_id + batch.n;
                        const uint32_t *loc = ch.local_id.data() + i;
                        const uint32_t base = (uint32_t)id0;
                        for (size_t k = 0; k < m; ++k) rid[k] = base + loc[k];
                        memcpy(batch.ref_id + batch.n, ch.ref.data() + i, m * 4);
                        memcpy(batch.begin_pos + batch.n, ch.pos.data() + i, m * 4);
                        batch.n += m;
                        i += m;
                    }
                }
                if (!ok) break;
                if (n) {
                    prev_key.assign(ch.keys.data() + ch.key_off[n - 1], ch.key_len[n - 1]); prev_hash = ch.hash[n - 1]; have_prev = true;
                }
                st.records_kept += n;
                continue;
            }
            for (size_t i = 0; i < n; ++i) {
                if (i + 8 < n) table.prefetch(ch.hash[i + 8]);
                const char *key = ch.keys.data() + ch.key_off[i];
                const uint32_t len = ch.key_len[i];
                uint32_t id;
                if (have_prev && ch.hash[i] == prev_hash && prev_key.size() == len && memcmp(prev_key.data(), key, len) == 0) id = prev_id;
                else {
                    if (table.full()) { err = "more than 2^32-2 distinct reads"; ok = false; break; }
                    id = table.lookup_or_insert(ch.hash[i], key, len);
                    prev_key.assign(key, len); prev_hash = ch.hash[i]; prev_id = id; have_prev = true;
                }
                if (batch.n == batch.cap) batch = sink(batch);   // the sink hands back the batch to fill next (n = records it already holds)
                batch.read_id[batch.n] = id; batch.ref_id[batch.n] = ch.ref[i]; batch.begin_pos[batch.n] = ch.pos[i];
                ++batch.n;
            }
            if (!ok) break;
            st.records_kept += n;
        }
        if (ok && assume_grouped && came_back.load()) { st.not_grouped = true; ok = false; }   // every worker is done by now
        if (!ok) { stop.store(true); parse.abort(); }
        framer.join();
        if (ok && !frame_err.empty()) { err = frame_err; ok = false; }
        if (ok && batch.n) sink(batch);
        st.reads = assume_grouped ? next_id : table.size();
        return ok;
    }

private:
    // Cuts the byte stream into jobs of whole records.  emit returns false to stop early.  limit_hint is unused
    // by the framing itself (the callers stop through emit).
    template <class Emit>
    bool frame(ChunkReader &rd, Emit &&emit, std::string &err, uint32_t /*limit_hint*/)
    {
        std::string carry;
        bool in_header = is_bam_;        // the binary header of a BAM still has to be skipped
        Buffer raw;
        while (rd.next(raw)) {
            if (!raw.error.empty()) { err = raw.error; return false; }
            std::shared_ptr<Buffer> buf = std::make_shared<Buffer>(std::move(raw));
            size_t begin = 0;
            if (in_header) {             // accumulate until the header is complete, then continue behind it
                carry.append(buf->p, buf->n);
                size_t hdr = 0;
                AlignmentHeader tmp;
                if (!parse_bam_header(carry.data(), carry.size(), tmp, hdr)) continue;
                in_header = false;
                std::string rest = carry.substr(hdr);
                carry.clear();
                buf = std::make_shared<Buffer>();
                buf->own.reset(new char[rest.size() ? rest.size() : 1]);
                memcpy(buf->own.get(), rest.data(), rest.size());
                buf->p = buf->own.get(); buf->n = rest.size();
            }
            ParseJob job;
            const char *p = buf->p, *e = buf->p + buf->n;
            if (is_bam_) {
                if (!carry.empty()) {    // finish the record that started in the previous chunk
                    while (carry.size() < 4 && begin < buf->n) carry.push_back(p[begin++]);
                    if (carry.size() < 4) continue;
                    const size_t need = 4 + (size_t)rd32(carry.data());
                    const size_t take = std::min(need - carry.size(), buf->n - begin);
                    carry.append(p + begin, take);
                    begin += take;
                    if (carry.size() < need) continue;
                    job.head.swap(carry);
                    carry.clear();
                }
                size_t end = begin;
                while (end + 4 <= buf->n) {
                    const size_t bs = rd32(p + end);
                    if (end + 4 + bs > buf->n) break;
                    end += 4 + bs;
                }
                carry.assign(p + end, buf->n - end);
                job.buf = buf; job.begin = begin; job.end = end;
            } else {
                const char *first_nl = (const char *)memchr(p, '\n', buf->n);
                if (!first_nl) { carry.append(p, buf->n); continue; }
                if (!carry.empty()) {
                    carry.append(p, first_nl - p);
                    job.head.swap(carry);
                    carry.clear();
                    begin = (size_t)(first_nl - p) + 1;
                }
                const char *last_nl = (const char *)memrchr(p + begin, '\n', buf->n - begin);
                const size_t end = last_nl ? (size_t)(last_nl - p) + 1 : begin;
                carry.assign(p + end, buf->n - end);
                job.buf = buf; job.begin = begin; job.end = end;
            }
            if (!emit(job)) return false;
        }
        if (in_header) { err = "truncated BAM header"; return false; }
        if (!carry.empty()) {
            if (is_bam_) { err = "truncated BAM record at the end of the file"; return false; }
            ParseJob job;                // last line without a newline
            job.head.swap(carry);
            if (!emit(job)) return false;
        }
        return true;
    }

    static bool parse_bam_header(const char *p, size_t n, AlignmentHeader &h, size_t &hdr_len)
    {
        if (n < 12) return false;
        const size_t l_text = rd32(p + 4);
        size_t off = 8 + l_text;
        if (n < off + 4) return false;
        const uint32_t n_ref = rd32(p + off);
        off += 4;
        for (uint32_t i = 0; i < n_ref; ++i) {
            if (n < off + 4) return false;
            const size_t l_name = rd32(p + off);
            if (n < off + 4 + l_name + 4) return false;
            h.names.emplace_back(p + off + 4, l_name ? l_name - 1 : 0);
            h.lengths.push_back(rd32(p + off + 4 + l_name));
            off += 4 + l_name + 4;
        }
        hdr_len = off;
        return true;
    }

    bool read_header(std::string &err)
    {
        std::unique_ptr<ChunkReader> rd = make_reader(file_, comp_, 1u << 20, 1);
        std::string acc;
        Buffer b;
        bool first = true, done = false;
        while (!done && rd->next(b)) {
            if (!b.error.empty()) { err = b.error; return false; }
            acc.append(b.p, b.n);
            if (first && acc.size() >= 4) { is_bam_ = memcmp(acc.data(), "BAM\1", 4) == 0; first = false; }
            if (first) continue;
            if (is_bam_) {
                size_t hdr = 0;
                AlignmentHeader tmp;
                if (parse_bam_header(acc.data(), acc.size(), tmp, hdr)) { header_ = std::move(tmp); done = true; }
            } else {
                // the header ends at the first line that does not start with '@'
                size_t p = 0;
                bool complete = false;
                while (p < acc.size()) {
                    if (acc[p] != '@') { complete = true; break; }
                    const size_t nl = acc.find('\n', p);
                    if (nl == std::string::npos) break;
                    p = nl + 1;
                }
                if (complete) done = true;
            }
        }
        if (first) { is_bam_ = false; }
        if (is_bam_ && !done) { err = path_ + ": truncated BAM header"; return false; }
        if (!is_bam_) {
            size_t p = 0;
            while (p < acc.size() && acc[p] == '@') {
                size_t nl = acc.find('\n', p);
                if (nl == std::string::npos) nl = acc.size();
                size_t le = nl;
                if (le > p && acc[le - 1] == '\r') --le;
                if (le - p >= 3 && acc.compare(p, 3, "@SQ") == 0) {
                    std::string sn;
                    uint32_t ln = 0;
                    size_t f = p;
                    while (f < le) {
                        size_t t = acc.find('\t', f);
                        if (t == std::string::npos || t > le) t = le;
                        if (t - f > 3 && acc.compare(f, 3, "SN:") == 0) sn = acc.substr(f + 3, t - f - 3);
                        else if (t - f > 3 && acc.compare(f, 3, "LN:") == 0) ln = (uint32_t)strtoull(acc.substr(f + 3, t - f - 3).c_str(), nullptr, 10);
                        f = t + 1;
                    }
                    header_.names.push_back(sn);
                    header_.lengths.push_back(ln);
                }
                p = nl + 1;
            }
        }
        names_.build(header_.names);
        return true;
    }

    std::string path_;
    MappedFile file_;
    Compression comp_ = Compression::none;
    bool is_bam_ = false;
    AlignmentHeader header_;
    NameIndex names_;
};

}  // namespace slimm_fe

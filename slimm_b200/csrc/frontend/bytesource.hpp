// bytesource.hpp - the input file as a sequence of decompressed byte chunks, in file order.
//   plain file  : zero-copy slices of the mapping
//   BGZF (BAM, bgzipped SAM): blocks are independent deflate streams with their sizes in the block header and
//                 trailer, so groups of blocks are inflated by worker threads straight into their final place
//   plain gzip  : one sequential zlib stream (possibly several members)
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "pipeline.hpp"

namespace slimm_fe {

struct MappedFile {
    const unsigned char *p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open(const std::string &path)
    {
        fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { ::close(fd); fd = -1; return false; }
        n = (size_t)st.st_size;
        if (n) {
            void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { ::close(fd); fd = -1; return false; }
            p = (const unsigned char *)m;
            madvise(m, n, MADV_SEQUENTIAL);
        }
        return true;
    }
    ~MappedFile()
    {
        if (p) munmap((void *)p, n);
        if (fd >= 0) ::close(fd);
    }
};

struct Buffer {
    const char *p = nullptr;
    size_t n = 0;
    std::unique_ptr<char[]> own;   // empty for slices of the mapping
    std::string error;
};

enum class Compression { none, gzip, bgzf };

static inline uint32_t le16(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static inline uint32_t le32(const unsigned char *p) { return le16(p) | (le16(p + 2) << 16); }

// total size of the BGZF block at p (0: not a BGZF block)
static inline size_t bgzf_block_size(const unsigned char *p, size_t avail)
{
    if (avail < 18 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
    const size_t xlen = le16(p + 10);
    if (avail < 12 + xlen) return 0;
    for (size_t o = 12; o + 4 <= 12 + xlen;) {
        const size_t slen = le16(p + o + 2);
        if (p[o] == 'B' && p[o + 1] == 'C' && slen == 2) return (size_t)le16(p + o + 4) + 1;
        o += 4 + slen;
    }
    return 0;
}

static inline Compression detect_compression(const MappedFile &f)
{
    if (f.n < 2 || f.p[0] != 0x1f || f.p[1] != 0x8b) return Compression::none;
    return bgzf_block_size(f.p, f.n) ? Compression::bgzf : Compression::gzip;
}

class ChunkReader {
public:
    virtual ~ChunkReader() {}
    virtual bool next(Buffer &out) = 0;   // false at the end of the file; out.error set on failure
};

class PlainReader : public ChunkReader {
public:
    PlainReader(const MappedFile &f, size_t chunk) : f_(f), chunk_(chunk) {}
    bool next(Buffer &out) override
    {
        if (off_ >= f_.n) return false;
        out = Buffer();
        out.p = (const char *)f_.p + off_;
        out.n = std::min(chunk_, f_.n - off_);
        off_ += out.n;
        return true;
    }

private:
    const MappedFile &f_;
    size_t chunk_, off_ = 0;
};

class GzipReader : public ChunkReader {
public:
    GzipReader(const MappedFile &f, size_t chunk) : f_(f), chunk_(chunk)
    {
        memset(&z_, 0, sizeof z_);
        ok_ = inflateInit2(&z_, 15 + 32) == Z_OK;
        z_.next_in = (Bytef *)f_.p;
    }
    ~GzipReader() override { if (ok_) inflateEnd(&z_); }
    bool next(Buffer &out) override
    {
        out = Buffer();
        if (!ok_ || done_) return false;
        out.own.reset(new char[chunk_]);
        size_t have = 0;
        while (have < chunk_) {
            if (z_.avail_in == 0) {   // zlib counts in 32 bits: feed the mapping in pieces
                const size_t left = f_.n - in_off_;
                if (left == 0) { done_ = true; break; }
                const size_t take = std::min<size_t>(left, 1u << 30);
                z_.next_in = (Bytef *)f_.p + in_off_;
                z_.avail_in = (uInt)take;
                in_off_ += take;
            }
            z_.next_out = (Bytef *)out.own.get() + have;
            z_.avail_out = (uInt)(chunk_ - have);
            const int rc = inflate(&z_, Z_NO_FLUSH);
            have = chunk_ - z_.avail_out;
            if (rc == Z_STREAM_END) {
                if (z_.avail_in == 0 && in_off_ >= f_.n) { done_ = true; break; }
                inflateReset(&z_);    // next gzip member
            } else if (rc != Z_OK) {
                out.error = std::string("gzip stream is corrupt: ") + (z_.msg ? z_.msg : "inflate failed");
                done_ = true;
                break;
            }
        }
        out.p = out.own.get();
        out.n = have;
        return have > 0 || !out.error.empty();
    }

private:
    const MappedFile &f_;
    size_t chunk_, in_off_ = 0;
    z_stream z_;
    bool ok_ = false, done_ = false;
};

class BgzfReader : public ChunkReader {
    struct Block { size_t in_off, in_len, out_off, out_len; };
    struct Group { std::vector<Block> blocks; size_t out_len = 0; std::string error; };

public:
    BgzfReader(const MappedFile &f, size_t chunk, int n_threads)
        : f_(f), chunk_(chunk), stage_(n_threads, (size_t)n_threads * 2 + 2, [this](Group &g, Buffer &b) { inflate_group(g, b); })
    {
        feeder_ = std::thread([this] { feed(); });
    }
    ~BgzfReader() override
    {
        stage_.abort();
        if (feeder_.joinable()) feeder_.join();
    }
    bool next(Buffer &out) override
    {
        for (;;) {
            out = Buffer();
            if (!stage_.pop(out)) return false;
            if (out.n || !out.error.empty()) return true;   // an empty group (EOF marker block) carries nothing
        }
    }

private:
    void feed()
    {
        size_t off = 0;
        Group g;
        while (off < f_.n) {
            const size_t bs = bgzf_block_size(f_.p + off, f_.n - off);
            if (bs < 26 || off + bs > f_.n) { g.error = "truncated or corrupt BGZF block"; break; }
            const size_t xlen = le16(f_.p + off + 10);
            const size_t isize = le32(f_.p + off + bs - 4);
            // the deflate payload must fit between the extra field and the trailer; a BGZF block inflates to at most 64 KiB
            if (bs < 12 + xlen + 8 || isize > 65536) { g.error = "truncated or corrupt BGZF block"; break; }
            g.blocks.push_back(Block{off + 12 + xlen, bs - 12 - xlen - 8, g.out_len, isize});
            g.out_len += isize;
            off += bs;
            if (g.out_len >= chunk_) { stage_.push(std::move(g)); g = Group(); }
        }
        if (!g.blocks.empty() || !g.error.empty()) stage_.push(std::move(g));
        stage_.close();
    }
    void inflate_group(Group &g, Buffer &b)
    {
        b.error = g.error;
        b.own.reset(new char[g.out_len ? g.out_len : 1]);
        b.p = b.own.get();
        b.n = g.out_len;
        z_stream z;
        memset(&z, 0, sizeof z);
        if (inflateInit2(&z, -15) != Z_OK) { b.error = "zlib initialisation failed"; return; }
        for (const Block &k : g.blocks) {
            if (k.out_len == 0) continue;
            inflateReset(&z);
            z.next_in = (Bytef *)f_.p + k.in_off; z.avail_in = (uInt)k.in_len;
            z.next_out = (Bytef *)b.own.get() + k.out_off; z.avail_out = (uInt)k.out_len;
            const int rc = inflate(&z, Z_FINISH);
            if (rc != Z_STREAM_END || z.avail_out != 0) { b.error = "BGZF block does not inflate to its recorded size"; break; }
        }
        inflateEnd(&z);
    }
    const MappedFile &f_;
    size_t chunk_;
    OrderedStage<Group, Buffer> stage_;
    std::thread feeder_;
};

static inline std::unique_ptr<ChunkReader> make_reader(const MappedFile &f, Compression c, size_t chunk, int n_threads)
{
    if (c == Compression::bgzf) return std::unique_ptr<ChunkReader>(new BgzfReader(f, chunk, n_threads));
    if (c == Compression::gzip) return std::unique_ptr<ChunkReader>(new GzipReader(f, chunk));
    return std::unique_ptr<ChunkReader>(new PlainReader(f, chunk));
}

}  // namespace slimm_fe

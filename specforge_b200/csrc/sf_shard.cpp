// sf_shard.cpp — packed on-disk shards of offline feature records (host code, no CUDA).
//
// The reference stores ONE `torch.save` dict per sample (scripts/prepare_hidden_states.py:446-480, path scheme :578-595)
// and reads it back with `torch.load(..., mmap=True)` (runtime/data_plane/feature_store.py:235-240), then truncates
// (algorithms/eagle3/data.py:10-27) and pads/concatenates on the host (data/utils.py:106-200) before a pageable H2D copy.
// At config 2 a sample is 67 MB and a GPU consumes 33 of them per second; this file is the I/O side of that feed:
//
//   shard file ("SFPK", little endian; written by specforge_b200/shards.py)
//     header   128 B   magic "SFPK", version, n_features, n_records, offsets of the three sections, file size
//     features n_features x 64 B   raw key name, dtype code, element bytes, elements per token
//     index    n_records x 16 B    payload offset (4096-aligned), token count, CRC-32 of the payload
//     payload  per record: feature blocks in table order, [tokens, width] row-major, each block 64-B aligned
//
// sf_shard_read_batch() gathers B records straight into batch-major destination buffers ([B, pad_tokens, width], e.g.
// pinned host tensors), truncating to max_tokens and zero-filling the tail: truncation + collation + the copy into
// pinned memory are one pass of pread() calls issued by a few worker threads, no intermediate tensors.
#include "../../include/specforge_b200.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace sf {
int set_error(int code, const char* fmt, ...);
}

namespace {

constexpr uint32_t kVersion = 1;
constexpr size_t kHeaderBytes = 128, kFeatureBytes = 64, kIndexBytes = 16;

#pragma pack(push, 1)
struct Header {
    char magic[4];
    uint32_t version;
    uint32_t n_features;
    uint32_t flags;
    uint64_t n_records;
    uint64_t feature_off, index_off, data_off, file_bytes;
    uint8_t reserved[kHeaderBytes - 56];
};
struct FeatureRec {
    char name[40];
    uint32_t dtype;        // SF_DT_*
    uint32_t elem_bytes;
    uint64_t width;        // elements per token
    uint64_t reserved;
};
struct IndexRec {
    uint64_t offset;
    uint32_t num_tokens;
    uint32_t crc32;
};
#pragma pack(pop)
static_assert(sizeof(Header) == kHeaderBytes && sizeof(FeatureRec) == kFeatureBytes && sizeof(IndexRec) == kIndexBytes, "layout");

struct Shard {
    int fd = -1;
    std::string path;
    Header h{};
    std::vector<FeatureRec> feat;
    std::vector<IndexRec> index;
    std::vector<uint64_t> row_bytes;   // per feature: width * elem_bytes
    const uint8_t* map = nullptr;      // read-only mapping of the whole file (null: fall back to pread)
};

uint64_t align_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }

// offset of feature f's block inside a record of `tokens` tokens (blocks are 64-B aligned, table order)
uint64_t block_offset(const Shard& s, int f, uint64_t tokens) {
    uint64_t off = 0;
    for (int i = 0; i < f; ++i) off = align_up(off + tokens * s.row_bytes[i], 64);
    return off;
}

bool pread_full(int fd, void* dst, size_t n, uint64_t off) {
    uint8_t* p = static_cast<uint8_t*>(dst);
    while (n) {
        const ssize_t r = pread(fd, p, n, (off_t)off);
        if (r < 0) { if (errno == EINTR) continue; return false; }
        if (r == 0) return false;   // short file
        p += r; off += (uint64_t)r; n -= (size_t)r;
    }
    return true;
}

uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {   // zlib polynomial, bitwise-reflected, table driven
    static uint32_t table[256];
    static std::atomic<bool> ready{false};
    if (!ready.load(std::memory_order_acquire)) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready.store(true, std::memory_order_release);
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return ~crc;
}

}  // namespace

extern "C" int sf_shard_open(const char* path, void** handle) {
    if (!path || !handle) return sf::set_error(-22, "shard: null argument");
    *handle = nullptr;
    Shard* s = new Shard();
    s->path = path;
    s->fd = open(path, O_RDONLY | O_CLOEXEC);
    if (s->fd < 0) { const int e = errno; delete s; return sf::set_error(-e, "shard: cannot open %s: %s", path, strerror(e)); }
    auto fail = [&](int code, const char* what) {
        close(s->fd);
        delete s;
        return sf::set_error(code, "shard %s: %s", path, what);
    };
    struct stat st{};
    if (fstat(s->fd, &st) != 0) return fail(-5, "fstat failed");
    if ((size_t)st.st_size < kHeaderBytes || !pread_full(s->fd, &s->h, kHeaderBytes, 0)) return fail(-22, "truncated header");
    if (memcmp(s->h.magic, "SFPK", 4) != 0) return fail(-22, "bad magic (not an SFPK shard)");
    if (s->h.version != kVersion) return fail(-22, "unsupported shard version");
    if (s->h.file_bytes != (uint64_t)st.st_size) return fail(-22, "file size differs from the header (truncated or appended)");
    if (s->h.n_features == 0 || s->h.n_features > 64) return fail(-22, "bad feature count");
    if (s->h.feature_off + (uint64_t)s->h.n_features * kFeatureBytes > s->h.file_bytes ||
        s->h.index_off + s->h.n_records * kIndexBytes > s->h.file_bytes || s->h.data_off > s->h.file_bytes)
        return fail(-22, "section offsets outside the file");
    s->feat.resize(s->h.n_features);
    if (!pread_full(s->fd, s->feat.data(), s->feat.size() * kFeatureBytes, s->h.feature_off)) return fail(-5, "cannot read feature table");
    s->index.resize(s->h.n_records);
    if (s->h.n_records && !pread_full(s->fd, s->index.data(), s->index.size() * kIndexBytes, s->h.index_off))
        return fail(-5, "cannot read index");
    for (auto& f : s->feat) {
        f.name[sizeof(f.name) - 1] = 0;
        if (f.elem_bytes == 0 || f.elem_bytes > 16 || f.width == 0) return fail(-22, "bad feature descriptor");
        s->row_bytes.push_back(f.width * f.elem_bytes);
    }
    for (const auto& r : s->index) {
        const uint64_t bytes = block_offset(*s, (int)s->feat.size(), r.num_tokens);
        if (r.offset < s->h.data_off || r.offset + bytes > s->h.file_bytes) return fail(-22, "record outside the file");
    }
    // Map the file: a warm page cache is then read at memcpy speed by every worker (pread pays a syscall + a page-cache
    // lookup per 4 KB); cold ranges are requested ahead with MADV_WILLNEED.  SF_SHARD_NO_MMAP=1 forces the pread path.
    const char* no_mmap = getenv("SF_SHARD_NO_MMAP");
    if (!(no_mmap && no_mmap[0] == '1')) {
        void* m = mmap(nullptr, (size_t)s->h.file_bytes, PROT_READ, MAP_SHARED, s->fd, 0);
        if (m != MAP_FAILED) s->map = static_cast<const uint8_t*>(m);
    }
    *handle = s;
    return 0;
}

extern "C" void sf_shard_close(void* handle) {
    Shard* s = static_cast<Shard*>(handle);
    if (!s) return;
    if (s->map) munmap(const_cast<uint8_t*>(s->map), (size_t)s->h.file_bytes);
    if (s->fd >= 0) close(s->fd);
    delete s;
}

extern "C" int64_t sf_shard_num_records(void* handle) { return handle ? (int64_t)static_cast<Shard*>(handle)->h.n_records : -1; }
extern "C" int sf_shard_num_features(void* handle) { return handle ? (int)static_cast<Shard*>(handle)->h.n_features : -1; }

extern "C" int sf_shard_feature_info(void* handle, int f, char* name40, int* dtype, int* elem_bytes, int64_t* width) {
    Shard* s = static_cast<Shard*>(handle);
    if (!s || f < 0 || f >= (int)s->feat.size()) return sf::set_error(-22, "shard: bad feature index %d", f);
    if (name40) memcpy(name40, s->feat[f].name, 40);
    if (dtype) *dtype = (int)s->feat[f].dtype;
    if (elem_bytes) *elem_bytes = (int)s->feat[f].elem_bytes;
    if (width) *width = (int64_t)s->feat[f].width;
    return 0;
}

extern "C" int64_t sf_shard_record_tokens(void* handle, int64_t rec) {
    Shard* s = static_cast<Shard*>(handle);
    if (!s || rec < 0 || (uint64_t)rec >= s->h.n_records) return sf::set_error(-22, "shard: bad record index %lld", (long long)rec);
    return s->index[rec].num_tokens;
}

// CRC-32 of record `rec`'s payload, recomputed from disk; returns 0 when it matches the index, 1 when it does not.
extern "C" int sf_shard_verify_record(void* handle, int64_t rec) {
    Shard* s = static_cast<Shard*>(handle);
    if (!s || rec < 0 || (uint64_t)rec >= s->h.n_records) return sf::set_error(-22, "shard: bad record index %lld", (long long)rec);
    const IndexRec& r = s->index[rec];
    const uint64_t bytes = block_offset(*s, (int)s->feat.size(), r.num_tokens);
    std::vector<uint8_t> buf(1 << 20);
    uint32_t crc = 0;
    for (uint64_t done = 0; done < bytes;) {
        const size_t n = (size_t)std::min<uint64_t>(buf.size(), bytes - done);
        if (!pread_full(s->fd, buf.data(), n, r.offset + done)) return sf::set_error(-5, "shard %s: read failed", s->path.c_str());
        crc = crc32_update(crc, buf.data(), n);
        done += n;
    }
    return crc == r.crc32 ? 0 : 1;
}

extern "C" int sf_shard_read_batch(void* handle, const int64_t* records, int n_rec, int64_t max_tokens, int64_t pad_tokens,
                                   void* const* dst, int n_threads) {
    Shard* s = static_cast<Shard*>(handle);
    if (!s || !records || !dst || n_rec <= 0) return sf::set_error(-22, "shard: bad read_batch argument");
    if (max_tokens <= 0 || pad_tokens < 0) return sf::set_error(-22, "shard: max_tokens must be positive");
    const int nf = (int)s->feat.size();
    struct Job { int f; int b; uint64_t src; uint64_t bytes; uint8_t* dst; uint64_t zero; };
    std::vector<Job> jobs;
    for (int b = 0; b < n_rec; ++b) {
        const int64_t rec = records[b];
        if (rec < 0 || (uint64_t)rec >= s->h.n_records) return sf::set_error(-22, "shard: record %lld out of range", (long long)rec);
        const IndexRec& r = s->index[rec];
        const uint64_t take = std::min<uint64_t>(r.num_tokens, (uint64_t)max_tokens);
        if (take > (uint64_t)pad_tokens) return sf::set_error(-22, "shard: record %lld has %llu tokens after truncation, pad_tokens=%lld",
                                                              (long long)rec, (unsigned long long)take, (long long)pad_tokens);
        for (int f = 0; f < nf; ++f) {
            if (!dst[f]) continue;
            const uint64_t rb = s->row_bytes[f];
            uint8_t* d = static_cast<uint8_t*>(dst[f]) + (uint64_t)b * (uint64_t)pad_tokens * rb;
            // large blocks are split so that a few records still keep every worker busy
            const uint64_t total = take * rb, chunk = 8ull << 20;
            const uint64_t base = r.offset + block_offset(*s, f, r.num_tokens);
            for (uint64_t o = 0; o < total || o == 0; o += chunk) {
                const uint64_t n = total > o ? std::min(chunk, total - o) : 0;
                const bool last = o + n >= total;
                jobs.push_back({f, b, base + o, n, d + o, last ? ((uint64_t)pad_tokens - take) * rb : 0});
                if (last) break;
            }
        }
    }
    if (s->map)   // ask for every range of the batch up front so cold pages stream in while the first ones are copied
        for (const Job& j : jobs)
            if (j.bytes) {
                const uint64_t a = j.src & ~4095ull;
                madvise(const_cast<uint8_t*>(s->map) + a, (size_t)(j.src + j.bytes - a), MADV_WILLNEED);
            }
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    auto work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= jobs.size() || failed.load()) return;
            const Job& j = jobs[i];
            if (j.bytes) {
                if (s->map) memcpy(j.dst, s->map + j.src, (size_t)j.bytes);
                else if (!pread_full(s->fd, j.dst, (size_t)j.bytes, j.src)) { failed.store(1); return; }
            }
            if (j.zero) memset(j.dst + j.bytes, 0, (size_t)j.zero);
        }
    };
    int nt = n_threads <= 0 ? 4 : n_threads;
    if ((size_t)nt > jobs.size()) nt = (int)jobs.size();
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (failed.load()) return sf::set_error(-5, "shard %s: read failed (short file or I/O error)", s->path.c_str());
    return 0;
}

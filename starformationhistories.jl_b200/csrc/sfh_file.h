// sfh_file.h -- the on-disk container for template stacks and fit results.
//
// The reference defines no file format: users `Serialization.serialize` their template matrices by hand
// (examples/fitting1.ipynb cell 96) and rebuild them otherwise.  A 40 GB stack that takes minutes to construct needs to
// be reusable across runs and loadable shard by shard (one process per GPU, each touching only its bin rows), so this
// file fixes one: a little-endian, page-aligned, memory-mappable table of named n-dimensional arrays.
//
//   offset 0      FileHeader   (128 bytes)
//   offset 128    ArrayEntry[narrays]   (128 bytes each)
//   offset header_bytes = round_up(128 + 128*narrays, 4096):  the arrays, each starting on a 4096-byte boundary, in
//                 the order of the table, stored exactly as they sit in host memory (column-major for matrices: the
//                 "models" array IS the stack_models matrix, src/fitting/utilities.jl:12-13).
//
// Every array carries a 64-bit checksum that is a SUM of independently mixed words,
//       sum_i  mix64( w_i  xor  (i+1) * 0x9E3779B97F4A7C15 )      (mod 2^64; w_i = i-th little-endian 64-bit word of the
//                                                                  array, the tail zero-padded; mix64 = splitmix64 finaliser)
// so it can be computed in any order: by several host threads, per bin-row shard, or vectorised in numpy
// (tests/file_ref.py restates reader, writer and checksum independently of this file).
//
// Host-only C++; nothing here touches CUDA.
#ifndef SFH_FILE_H
#define SFH_FILE_H

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace sfh {
namespace file {

constexpr uint64_t kAlign = 4096;
constexpr uint32_t kVersion = 1;
constexpr uint32_t kEndianTag = 0x01020304u;
constexpr char kMagic[8] = {'S', 'F', 'H', 'F', 'I', 'L', 'E', '1'};
constexpr int kMaxDims = 4;
constexpr int kNameLen = 48;

struct FileHeader {  // 128 bytes
    char magic[8];
    uint32_t version, endian;
    uint64_t header_bytes, file_bytes;
    int32_t narrays, kind;
    int64_t attrs[8];
    uint64_t table_checksum;
    uint64_t reserved[2];
};
static_assert(sizeof(FileHeader) == 128, "FileHeader layout");

struct ArrayEntry {  // 128 bytes
    char name[kNameLen];
    int32_t dtype, ndim;
    int64_t dims[kMaxDims];
    uint64_t offset, nbytes, checksum;
    uint64_t reserved[2];
};
static_assert(sizeof(ArrayEntry) == 128, "ArrayEntry layout");

inline uint64_t round_up_u64(uint64_t a, uint64_t b) { return (a + b - 1) / b * b; }

inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// checksum of the words [word0, word0 + ceil(nbytes/8)) of an array whose first byte in this call is `p`
inline uint64_t checksum_range(const unsigned char *p, uint64_t nbytes, uint64_t word0) {
    constexpr uint64_t kGolden = 0x9E3779B97F4A7C15ull;
    const uint64_t nfull = nbytes / 8;
    uint64_t acc = 0, idx = (word0 + 1) * kGolden;
    for (uint64_t i = 0; i < nfull; ++i, idx += kGolden) {
        uint64_t w;
        memcpy(&w, p + 8 * i, 8);
        acc += mix64(w ^ idx);
    }
    if (nbytes % 8) {
        uint64_t w = 0;
        memcpy(&w, p + 8 * nfull, nbytes % 8);
        acc += mix64(w ^ idx);
    }
    return acc;
}

inline uint64_t checksum(const void *data, uint64_t nbytes, int nthreads = 0) {
    const unsigned char *p = (const unsigned char *)data;
    const uint64_t nwords = (nbytes + 7) / 8;
    if (nthreads <= 0) nthreads = (int)std::min<uint64_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
    if (nwords < (1u << 20) || nthreads == 1) return checksum_range(p, nbytes, 0);
    std::vector<uint64_t> part((size_t)nthreads, 0);
    std::vector<std::thread> th;
    const uint64_t per = (nwords + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const uint64_t w0 = per * t, w1 = std::min(nwords, w0 + per);
        if (w0 >= w1) break;
        th.emplace_back([=, &part] {
            const uint64_t b0 = 8 * w0, b1 = std::min(nbytes, 8 * w1);
            part[(size_t)t] = checksum_range(p + b0, b1 - b0, w0);
        });
    }
    for (auto &x : th) x.join();
    uint64_t acc = 0;
    for (uint64_t v : part) acc += v;
    return acc;
}

inline size_t dtype_size(int dtype) { return dtype == 0 ? 4 : (dtype == 3 ? 1 : 8); }  // SFH_F32, SFH_U8, else F64/I64

struct ArraySpec {  // what a writer supplies per array
    std::string name;
    int dtype = 1, ndim = 1;
    int64_t dims[kMaxDims] = {0, 1, 1, 1};
    const void *ptr = nullptr;  // may be nullptr for a section the caller fills in through the mapping (see Writer)
};

inline bool spec_bytes(const ArraySpec &a, uint64_t *nbytes, std::string *err) {
    if (a.name.empty() || a.name.size() >= (size_t)kNameLen) { *err = "array name empty or longer than 47 bytes"; return false; }
    if (a.dtype < 0 || a.dtype > 3) { *err = "bad array dtype"; return false; }
    if (a.ndim < 1 || a.ndim > kMaxDims) { *err = "array ndim must be 1..4"; return false; }
    uint64_t n = 1;
    for (int d = 0; d < a.ndim; ++d) {
        if (a.dims[d] < 0) { *err = "negative array dimension"; return false; }
        if (a.dims[d] && n > (UINT64_MAX / 16) / (uint64_t)a.dims[d]) { *err = "array too large"; return false; }
        n *= (uint64_t)a.dims[d];
    }
    *nbytes = n * dtype_size(a.dtype);
    return true;
}

// Writer: lays the file out, maps it read/write so large sections can be produced in place (the device stack is copied
// straight into the mapping, column block by column block), then checksums, writes the table, syncs and renames.
class Writer {
public:
    ~Writer() { abort(); }
    bool begin(const char *path, int kind, const int64_t attrs[8], const std::vector<ArraySpec> &specs, std::string *err) {
        final_path_ = path;
        tmp_path_ = final_path_ + ".tmp." + std::to_string((long long)getpid());
        entries_.assign(specs.size(), ArrayEntry{});
        memset(&hdr_, 0, sizeof hdr_);
        memcpy(hdr_.magic, kMagic, 8);
        hdr_.version = kVersion; hdr_.endian = kEndianTag; hdr_.narrays = (int32_t)specs.size(); hdr_.kind = kind;
        if (attrs) memcpy(hdr_.attrs, attrs, sizeof hdr_.attrs);
        hdr_.header_bytes = round_up_u64(sizeof(FileHeader) + sizeof(ArrayEntry) * specs.size(), kAlign);
        uint64_t off = hdr_.header_bytes;
        for (size_t i = 0; i < specs.size(); ++i) {
            uint64_t nb = 0;
            if (!spec_bytes(specs[i], &nb, err)) return false;
            for (size_t k = 0; k < i; ++k)
                if (specs[k].name == specs[i].name) { *err = "duplicate array name '" + specs[i].name + "'"; return false; }
            ArrayEntry &e = entries_[i];
            memset(&e, 0, sizeof e);
            memcpy(e.name, specs[i].name.data(), specs[i].name.size());
            e.dtype = specs[i].dtype; e.ndim = specs[i].ndim;
            for (int d = 0; d < kMaxDims; ++d) e.dims[d] = d < specs[i].ndim ? specs[i].dims[d] : 1;
            e.offset = off; e.nbytes = nb;
            off = round_up_u64(off + nb, kAlign);
        }
        hdr_.file_bytes = off;
        fd_ = ::open(tmp_path_.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
        if (fd_ < 0) { *err = "cannot create '" + tmp_path_ + "': " + strerror(errno); return false; }
        if (ftruncate(fd_, (off_t)hdr_.file_bytes) != 0) { *err = std::string("ftruncate: ") + strerror(errno); abort(); return false; }
        map_ = (unsigned char *)mmap(nullptr, (size_t)hdr_.file_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd_, 0);
        if (map_ == (unsigned char *)MAP_FAILED) { map_ = nullptr; *err = std::string("mmap: ") + strerror(errno); abort(); return false; }
        for (size_t i = 0; i < specs.size(); ++i)
            if (specs[i].ptr && entries_[i].nbytes) memcpy(map_ + entries_[i].offset, specs[i].ptr, (size_t)entries_[i].nbytes);
        return true;
    }
    void *section(size_t i) { return map_ + entries_[i].offset; }
    bool commit(std::string *err) {
        for (ArrayEntry &e : entries_) e.checksum = checksum(map_ + e.offset, e.nbytes);
        if (!entries_.empty()) memcpy(map_ + sizeof(FileHeader), entries_.data(), sizeof(ArrayEntry) * entries_.size());
        hdr_.table_checksum = checksum(map_ + sizeof(FileHeader), sizeof(ArrayEntry) * entries_.size(), 1);
        memcpy(map_, &hdr_, sizeof hdr_);
        bool ok = msync(map_, (size_t)hdr_.file_bytes, MS_SYNC) == 0;
        if (!ok) *err = std::string("msync: ") + strerror(errno);
        munmap(map_, (size_t)hdr_.file_bytes); map_ = nullptr;
        if (ok && fsync(fd_) != 0) { ok = false; *err = std::string("fsync: ") + strerror(errno); }
        ::close(fd_); fd_ = -1;
        if (ok && rename(tmp_path_.c_str(), final_path_.c_str()) != 0) { ok = false; *err = std::string("rename: ") + strerror(errno); }
        if (!ok) unlink(tmp_path_.c_str());
        tmp_path_.clear();
        return ok;
    }
    void abort() {
        if (map_) { munmap(map_, (size_t)hdr_.file_bytes); map_ = nullptr; }
        if (fd_ >= 0) { ::close(fd_); fd_ = -1; }
        if (!tmp_path_.empty()) { unlink(tmp_path_.c_str()); tmp_path_.clear(); }
    }

private:
    std::string final_path_, tmp_path_;
    FileHeader hdr_{};
    std::vector<ArrayEntry> entries_;
    unsigned char *map_ = nullptr;
    int fd_ = -1;
};

// Reader: read-only mapping; the header and the table are validated at open, array payloads on request.
class Reader {
public:
    ~Reader() { close(); }
    bool open(const char *path, std::string *err) {
        fd_ = ::open(path, O_RDONLY);
        if (fd_ < 0) { *err = std::string("cannot open '") + path + "': " + strerror(errno); return false; }
        struct stat st;
        if (fstat(fd_, &st) != 0) { *err = std::string("fstat: ") + strerror(errno); close(); return false; }
        size_ = (uint64_t)st.st_size;
        if (size_ < sizeof(FileHeader)) { *err = "file shorter than its header"; close(); return false; }
        map_ = (const unsigned char *)mmap(nullptr, (size_t)size_, PROT_READ, MAP_SHARED, fd_, 0);
        if (map_ == (const unsigned char *)MAP_FAILED) { map_ = nullptr; *err = std::string("mmap: ") + strerror(errno); close(); return false; }
        memcpy(&hdr_, map_, sizeof hdr_);
        if (memcmp(hdr_.magic, kMagic, 8) != 0) { *err = "not an sfhcuda file (bad magic)"; close(); return false; }
        if (hdr_.endian != kEndianTag) { *err = "file written with another byte order"; close(); return false; }
        if (hdr_.version != kVersion) { *err = "unsupported file version " + std::to_string(hdr_.version); close(); return false; }
        const uint64_t table = sizeof(ArrayEntry) * (uint64_t)std::max(hdr_.narrays, 0);
        if (hdr_.narrays < 0 || hdr_.header_bytes % kAlign || sizeof(FileHeader) + table > hdr_.header_bytes ||
            hdr_.header_bytes > size_ || hdr_.file_bytes != size_) {
            *err = "corrupt header (sizes inconsistent with the file: truncated?)"; close(); return false;
        }
        if (checksum(map_ + sizeof(FileHeader), table, 1) != hdr_.table_checksum) { *err = "corrupt array table (checksum mismatch)"; close(); return false; }
        entries_.resize((size_t)hdr_.narrays);
        if (hdr_.narrays) memcpy(entries_.data(), map_ + sizeof(FileHeader), table);
        for (const ArrayEntry &e : entries_) {
            uint64_t n = dtype_size(e.dtype);
            bool ok = e.dtype >= 0 && e.dtype <= 3 && e.ndim >= 1 && e.ndim <= kMaxDims && e.name[kNameLen - 1] == 0;
            // same overflow-checked product as the writer's spec_bytes: dims whose product wraps mod 2^64 to match nbytes would pass
            // every other check here and index far outside the mapping later (the table checksum is not cryptographic)
            for (int d = 0; ok && d < e.ndim; ++d) {
                ok = e.dims[d] >= 0 && !(e.dims[d] && n > (UINT64_MAX / 16) / (uint64_t)e.dims[d]);
                n *= (uint64_t)e.dims[d];
            }
            if (!ok || n != e.nbytes || e.offset % kAlign || e.offset < hdr_.header_bytes || e.offset > size_ || e.nbytes > size_ - e.offset) {
                *err = "corrupt array table entry"; close(); return false;
            }
        }
        return true;
    }
    void close() {
        if (map_) { munmap((void *)map_, (size_t)size_); map_ = nullptr; }
        if (fd_ >= 0) { ::close(fd_); fd_ = -1; }
        entries_.clear();
    }
    const FileHeader &header() const { return hdr_; }
    int count() const { return (int)entries_.size(); }
    const ArrayEntry &entry(int i) const { return entries_[(size_t)i]; }
    const void *data(int i) const { return map_ + entries_[(size_t)i].offset; }
    int find(const char *name) const {
        for (size_t i = 0; i < entries_.size(); ++i)
            if (strncmp(entries_[i].name, name, kNameLen) == 0) return (int)i;
        return -1;
    }
    bool verify(int i) const { return checksum(data(i), entries_[(size_t)i].nbytes) == entries_[(size_t)i].checksum; }

private:
    FileHeader hdr_{};
    std::vector<ArrayEntry> entries_;
    const unsigned char *map_ = nullptr;
    uint64_t size_ = 0;
    int fd_ = -1;
};

}  // namespace file
}  // namespace sfh
#endif  // SFH_FILE_H

#include "gguf_reader.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>

#include "common.h"

namespace p5 {

namespace {

enum : uint32_t { T_U8, T_I8, T_U16, T_I16, T_U32, T_I32, T_F32, T_BOOL, T_STR, T_ARR, T_U64, T_I64, T_F64 };

struct Cursor {
    const uint8_t* p;
    const uint8_t* end;
    const std::string& path;
    template <class T>
    T get() {
        P5_REQUIRE(size_t(end - p) >= sizeof(T), P5_ERR_FORMAT, "%s: truncated GGUF header", path.c_str());
        T v;
        memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    std::string str() {
        uint64_t n = get<uint64_t>();
        P5_REQUIRE(n <= uint64_t(end - p), P5_ERR_FORMAT, "%s: corrupt GGUF string", path.c_str());
        std::string s(reinterpret_cast<const char*>(p), n);
        p += n;
        return s;
    }
};

void read_scalar(Cursor& c, uint32_t t, GgufValue& v) {
    switch (t) {
        case T_U8: v.u = c.get<uint8_t>(); v.f = double(v.u); break;
        case T_I8: { int8_t x = c.get<int8_t>(); v.u = uint64_t(int64_t(x)); v.f = x; break; }
        case T_U16: v.u = c.get<uint16_t>(); v.f = double(v.u); break;
        case T_I16: { int16_t x = c.get<int16_t>(); v.u = uint64_t(int64_t(x)); v.f = x; break; }
        case T_U32: v.u = c.get<uint32_t>(); v.f = double(v.u); break;
        case T_I32: { int32_t x = c.get<int32_t>(); v.u = uint64_t(int64_t(x)); v.f = x; break; }
        case T_F32: v.f = c.get<float>(); v.u = uint64_t(v.f); break;
        case T_BOOL: v.u = c.get<uint8_t>() != 0; v.f = double(v.u); break;
        case T_U64: v.u = c.get<uint64_t>(); v.f = double(v.u); break;
        case T_I64: { int64_t x = c.get<int64_t>(); v.u = uint64_t(x); v.f = double(x); break; }
        case T_F64: v.f = c.get<double>(); v.u = uint64_t(v.f); break;
        default: throw Error(P5_ERR_FORMAT, strf("%s: unknown GGUF metadata type %u", c.path.c_str(), t));
    }
}

void read_value(Cursor& c, uint32_t t, GgufValue& v) {
    v.type = t;
    if (t == T_STR) {
        v.s = c.str();
    } else if (t == T_ARR) {
        uint32_t et = c.get<uint32_t>();
        uint64_t n = c.get<uint64_t>();
        v.arr_len = n;
        P5_REQUIRE(et != T_ARR, P5_ERR_FORMAT, "%s: nested GGUF arrays are not supported", c.path.c_str());
        if (et == T_STR) {
            // every string takes at least its 8-byte length: a count beyond that is a corrupt file, not an allocation
            P5_REQUIRE(n <= uint64_t(c.end - c.p) / 8, P5_ERR_FORMAT, "%s: corrupt GGUF string array (%llu entries)", c.path.c_str(),
                       (unsigned long long)n);
            v.strs.reserve(n);
            for (uint64_t i = 0; i < n; ++i) v.strs.push_back(c.str());
        } else {
            GgufValue tmp;
            for (uint64_t i = 0; i < n; ++i) read_scalar(c, et, tmp);  // skipped: not needed by this model
        }
    } else {
        read_scalar(c, t, v);
    }
}

}  // namespace

GgufFile::GgufFile(const std::string& path) : path_(path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    P5_REQUIRE(fd_ >= 0, P5_ERR_IO, "cannot open %s: %s", path.c_str(), strerror(errno));
    struct stat st;
    if (fstat(fd_, &st) != 0 || st.st_size < 24) {
        ::close(fd_);
        throw Error(P5_ERR_FORMAT, strf("%s: not a GGUF file (too small)", path.c_str()));
    }
    size_ = uint64_t(st.st_size);
    void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (m == MAP_FAILED) {
        ::close(fd_);
        throw Error(P5_ERR_IO, strf("mmap of %s failed: %s", path.c_str(), strerror(errno)));
    }
    map_ = static_cast<const uint8_t*>(m);
    try {
        Cursor c{map_, map_ + size_, path_};
        P5_REQUIRE(memcmp(c.p, "GGUF", 4) == 0, P5_ERR_FORMAT, "%s: bad magic, not a GGUF file", path.c_str());
        c.p += 4;
        uint32_t version = c.get<uint32_t>();
        P5_REQUIRE(version == 2 || version == 3, P5_ERR_FORMAT, "%s: unsupported GGUF version %u", path.c_str(), version);
        uint64_t n_tensors = c.get<uint64_t>();
        uint64_t n_kv = c.get<uint64_t>();
        P5_REQUIRE(n_tensors < (1u << 20) && n_kv < (1u << 20), P5_ERR_FORMAT, "%s: corrupt GGUF counts", path.c_str());
        for (uint64_t i = 0; i < n_kv; ++i) {
            std::string key = c.str();
            uint32_t t = c.get<uint32_t>();
            GgufValue v;
            read_value(c, t, v);
            meta_[key] = std::move(v);
        }
        std::vector<GgufTensor> list;
        for (uint64_t i = 0; i < n_tensors; ++i) {
            GgufTensor t;
            t.name = c.str();
            uint32_t nd = c.get<uint32_t>();
            P5_REQUIRE(nd <= 4, P5_ERR_FORMAT, "%s: tensor %s has %u dims", path.c_str(), t.name.c_str(), nd);
            for (uint32_t d = 0; d < nd; ++d) t.ne.push_back(c.get<uint64_t>());
            t.type = c.get<uint32_t>();
            t.offset = c.get<uint64_t>();
            list.push_back(std::move(t));
        }
        uint64_t align = meta_u64("general.alignment", 32);
        P5_REQUIRE(align >= 1 && align <= 4096, P5_ERR_FORMAT, "%s: bad alignment", path.c_str());
        uint64_t data_start = (uint64_t(c.p - map_) + align - 1) / align * align;
        P5_REQUIRE(data_start <= size_, P5_ERR_FORMAT, "%s: tensor data starts beyond the file", path.c_str());
        for (auto& t : list) {
            // Tensors of other ggml types (decoder tensors, BF16, quantised blocks) may sit in the file unused: they are
            // recorded without data and only asking for one of them is an error (GgufFile::tensor).
            t.supported = t.type == 0 || t.type == 1;
            if (t.supported) {
                // overflow-checked size: a corrupt shape must not wrap around and pass the bounds check
                uint64_t n = 1;
                bool ok = true;
                for (uint64_t d : t.ne) {
                    if (d != 0 && n > (uint64_t(1) << 62) / d) ok = false;
                    n *= d;
                }
                const uint64_t esz = t.type == 1 ? 2 : 4, room = size_ - data_start;
                ok = ok && n <= room / esz && t.offset <= room && n * esz <= room - t.offset;
                P5_REQUIRE(ok, P5_ERR_FORMAT, "%s: tensor %s exceeds the file", path.c_str(), t.name.c_str());
                t.data = map_ + data_start + t.offset;
            }
            tensors_[t.name] = t;
        }
    } catch (...) {
        munmap(const_cast<uint8_t*>(map_), size_);
        ::close(fd_);
        throw;
    }
}

GgufFile::~GgufFile() {
    if (map_) munmap(const_cast<uint8_t*>(map_), size_);
    if (fd_ >= 0) ::close(fd_);
}

const GgufValue& GgufFile::meta(const std::string& k) const {
    auto it = meta_.find(k);
    P5_REQUIRE(it != meta_.end(), P5_ERR_FORMAT, "%s: metadata key %s is missing", path_.c_str(), k.c_str());
    return it->second;
}
uint64_t GgufFile::meta_u64(const std::string& k, uint64_t dflt) const {
    auto it = meta_.find(k);
    return it == meta_.end() ? dflt : it->second.u;
}
double GgufFile::meta_f64(const std::string& k, double dflt) const {
    auto it = meta_.find(k);
    return it == meta_.end() ? dflt : it->second.f;
}
std::string GgufFile::meta_str(const std::string& k, const std::string& dflt) const {
    auto it = meta_.find(k);
    return it == meta_.end() ? dflt : it->second.s;
}
const GgufTensor& GgufFile::tensor(const std::string& name) const {
    auto it = tensors_.find(name);
    P5_REQUIRE(it != tensors_.end(), P5_ERR_FORMAT, "%s: tensor %s is missing", path_.c_str(), name.c_str());
    P5_REQUIRE(it->second.supported, P5_ERR_UNSUPPORTED,
               "%s: tensor %s has ggml type %u; only F32/F16 weights are supported (use prostt5-f16.gguf)", path_.c_str(),
               name.c_str(), it->second.type);
    return it->second;
}

}  // namespace p5

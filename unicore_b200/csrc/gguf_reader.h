// Minimal GGUF (v2/v3) reader: metadata scalars / strings / arrays and F32 / F16 tensors, read from a
// read-only mmap.  The weight-file contract is `<model>/prostt5-f16.gguf` [REF src/modules/createdb.rs:148];
// the container format belongs to ggml (not in the reference tree).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace p5 {

struct GgufTensor {
    std::string name;
    std::vector<uint64_t> ne;  // ggml order: ne[0] is the contiguous dimension
    uint32_t type = 0;         // ggml type: 0 = F32, 1 = F16; anything else is recorded but cannot be read
    uint64_t offset = 0;       // from the start of the data section
    const uint8_t* data = nullptr;
    bool supported = true;     // false: unused tensor of another ggml type (data == nullptr)
    uint64_t n_elements() const {
        uint64_t n = 1;
        for (uint64_t d : ne) n *= d;
        return n;
    }
    uint64_t n_bytes() const { return n_elements() * (type == 1 ? 2 : 4); }
};

struct GgufValue {
    uint32_t type = 0;
    uint64_t u = 0;  // integers / bool
    double f = 0;    // floats
    std::string s;
    std::vector<std::string> strs;  // string arrays (tokenizer.ggml.tokens)
    uint64_t arr_len = 0;
};

class GgufFile {
public:
    explicit GgufFile(const std::string& path);
    ~GgufFile();
    GgufFile(const GgufFile&) = delete;
    GgufFile& operator=(const GgufFile&) = delete;

    bool has_meta(const std::string& k) const { return meta_.count(k) != 0; }
    const GgufValue& meta(const std::string& k) const;
    uint64_t meta_u64(const std::string& k, uint64_t dflt) const;
    double meta_f64(const std::string& k, double dflt) const;
    std::string meta_str(const std::string& k, const std::string& dflt) const;

    bool has_tensor(const std::string& name) const { return tensors_.count(name) != 0; }
    const GgufTensor& tensor(const std::string& name) const;
    const std::map<std::string, GgufTensor>& tensors() const { return tensors_; }
    const std::string& path() const { return path_; }

private:
    std::string path_;
    int fd_ = -1;
    const uint8_t* map_ = nullptr;
    uint64_t size_ = 0;
    std::map<std::string, GgufValue> meta_;
    std::map<std::string, GgufTensor> tensors_;
};

}  // namespace p5

// C entry points over the host helpers so that the CPU tests can compare them, function by function,
// with the Python restatement of the reference's Rust in oracle/host_oracle.py.
#include <cstring>
#include <string>

#include "host.h"
#include "md5.h"

extern "C" {

// writes at most cap-1 bytes + NUL; returns the full length
static size_t put(const std::string& s, char* out, size_t cap) {
    if (cap) {
        const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
        memcpy(out, s.data(), n);
        out[n] = 0;
    }
    return s.size();
}

size_t ubh_md5_hex(const void* data, size_t n, char* out, size_t cap) {
    ub::Md5 m;
    m.update(data, n);
    return put(m.hex(), out, cap);
}
size_t ubh_hashed_name(const void* seq, size_t n, char* out, size_t cap) {
    return put(ub::hashed_name(std::string(static_cast<const char*>(seq), n)), out, cap);
}
size_t ubh_sanitize_header(const void* key, size_t n, char* out, size_t cap) {
    return put(ub::sanitize_header(std::string(static_cast<const char*>(key), n)), out, cap);
}
// records as "header\0sequence\0header\0sequence\0..."; returns the number of records
size_t ubh_read_fasta(const char* path, char* out, size_t cap, size_t* needed) {
    auto recs = ub::read_fasta(path);
    std::string blob;
    for (auto& kv : recs) {
        blob += kv.first; blob.push_back('\0');
        blob += kv.second; blob.push_back('\0');
    }
    *needed = blob.size();
    if (blob.size() <= cap) memcpy(out, blob.data(), blob.size());
    return recs.size();
}
}

#include "host.h"

#include <dirent.h>
#include <signal.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <thread>
#include <unordered_map>

#include "md5.h"
#include "prostt5_b200.h"

namespace ub {

int g_verbosity = 3;

void msg(int level, const std::string& s) {
    if (g_verbosity >= level) {
        fputs(s.c_str(), stdout);
        fputc('\n', stdout);
        fflush(stdout);
    }
}

void die(int code, const std::string& what) {
    const char* label = "Unknown error";
    switch (code) {  // [REF src/envs/error_handler.rs:17-33]
        case ERR_GENERAL: label = "Error: "; break;
        case ERR_FILE_NOT_FOUND: label = "File not found: "; break;
        case ERR_FILE_INVALID: label = "Invalid file given: "; break;
        case ERR_BINARY_NOT_FOUND: label = "Binary not found: "; break;
        case ERR_MODULE_NOT_IMPLEMENTED: label = "Module not implemented: "; break;
        case ERR_ARGPARSE: label = "Argument parsing error: "; break;
        case ERR_OUTPUT_EXISTS: label = "Output file already exists: "; break;
    }
    if (g_verbosity >= 1) fprintf(stderr, "%s%s\n", label, what.c_str());
    exit(code);
}

bool path_exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}
bool is_dir(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
bool is_file(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
std::string base_name(const std::string& p) {
    std::string s = p;
    while (s.size() > 1 && s.back() == '/') s.pop_back();
    const size_t k = s.rfind('/');
    return k == std::string::npos ? s : s.substr(k + 1);
}
std::string parent_dir(const std::string& p) {
    std::string s = p;
    while (s.size() > 1 && s.back() == '/') s.pop_back();
    const size_t k = s.rfind('/');
    if (k == std::string::npos) return "";
    if (k == 0) return "/";
    return s.substr(0, k);
}
std::string file_stem(const std::string& p) {
    const std::string b = base_name(p);
    const size_t k = b.rfind('.');
    return (k == std::string::npos || k == 0) ? b : b.substr(0, k);
}
static std::string extension(const std::string& p) {
    const std::string b = base_name(p);
    const size_t k = b.rfind('.');
    return (k == std::string::npos || k == 0) ? "" : b.substr(k + 1);
}
void mkdir_p(const std::string& p) {
    if (p.empty() || is_dir(p)) return;
    mkdir_p(parent_dir(p));
    if (mkdir(p.c_str(), 0777) != 0 && !is_dir(p)) die(ERR_GENERAL, "Could not create directory " + p);
}

std::vector<std::string> list_files_with_ext(const std::string& dir, const std::string& ext) {
    std::vector<std::string> files;
    DIR* d = opendir(dir.c_str());
    if (!d) die(ERR_GENERAL, "Could not read directory " + dir);
    while (dirent* e = readdir(d)) {
        const std::string full = dir + (dir.back() == '/' ? "" : "/") + e->d_name;
        if (is_file(full) && extension(e->d_name) == ext) files.push_back(full);
    }
    closedir(d);
    std::sort(files.begin(), files.end());
    return files;
}

// Rust's BufRead::lines(): split at '\n', drop one trailing '\r'; lines that are not valid UTF-8 are
// errors and are skipped by the reference's filter_map(|l| l.ok()).
static bool valid_utf8(const std::string& s) {
    size_t i = 0;
    const size_t n = s.size();
    while (i < n) {
        const unsigned char c = s[i];
        size_t len;
        uint32_t cp;
        if (c < 0x80) { ++i; continue; }
        else if ((c & 0xE0) == 0xC0) { len = 2; cp = c & 0x1F; }
        else if ((c & 0xF0) == 0xE0) { len = 3; cp = c & 0x0F; }
        else if ((c & 0xF8) == 0xF0) { len = 4; cp = c & 0x07; }
        else return false;
        if (i + len > n) return false;
        for (size_t k = 1; k < len; ++k) {
            const unsigned char d = s[i + k];
            if ((d & 0xC0) != 0x80) return false;
            cp = (cp << 6) | (d & 0x3F);
        }
        if ((len == 2 && cp < 0x80) || (len == 3 && cp < 0x800) || (len == 4 && (cp < 0x10000 || cp > 0x10FFFF)) ||
            (cp >= 0xD800 && cp <= 0xDFFF))
            return false;
        i += len;
    }
    return true;
}

template <class F>
static void for_each_line(const std::string& path, F&& fn) {
    std::ifstream in(path, std::ios::binary);
    if (!in) die(ERR_GENERAL, "Unable to open file " + path);
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        fn(line);
    }
}

std::vector<std::pair<std::string, std::string>> read_fasta(const std::string& path) {
    std::vector<std::pair<std::string, std::string>> out;
    std::unordered_map<std::string, size_t> pos;
    auto insert = [&](const std::string& h, const std::string& s) {
        auto it = pos.find(h);
        if (it == pos.end()) {
            pos[h] = out.size();
            out.emplace_back(h, s);
        } else {
            out[it->second].second = s;
        }
    };
    std::string header, seq;
    for_each_line(path, [&](const std::string& line) {
        if (!valid_utf8(line)) return;
        if (!line.empty() && line[0] == '>') {
            if (!header.empty()) {
                insert(header, seq);
                seq.clear();
            }
            header = line.substr(1);
        } else {
            seq += line;
        }
    });
    insert(header, seq);
    return out;
}

// Unicode White_Space code points (char::is_whitespace)
static bool is_ws_cp(uint32_t cp) {
    return (cp >= 0x09 && cp <= 0x0D) || cp == 0x20 || cp == 0x85 || cp == 0xA0 || cp == 0x1680 ||
           (cp >= 0x2000 && cp <= 0x200A) || cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
}

std::string sanitize_header(const std::string& key) {
    std::string out;
    size_t i = 0;
    const size_t n = key.size();
    while (i < n) {
        const unsigned char c = key[i];
        size_t len = 1;
        uint32_t cp = c;
        if (c >= 0x80) {
            if ((c & 0xE0) == 0xC0) { len = 2; cp = c & 0x1F; }
            else if ((c & 0xF0) == 0xE0) { len = 3; cp = c & 0x0F; }
            else if ((c & 0xF8) == 0xF0) { len = 4; cp = c & 0x07; }
            if (i + len > n) len = 1;
            for (size_t k = 1; k < len; ++k) cp = (cp << 6) | (static_cast<unsigned char>(key[i + k]) & 0x3F);
        }
        const bool repl = is_ws_cp(cp) || cp == ';' || cp == ':' || cp == ',' || cp == '=' || cp == '/' || cp == '(' || cp == ')';
        if (repl) out.push_back('_');
        else out.append(key, i, len);
        i += len;
    }
    return out;
}

std::string hashed_name(const std::string& seq) { return "unicore_" + Md5::of(seq).substr(0, 10); }

std::vector<Record> collect_records(const std::string& input, const std::string& map_path, long max_len) {
    std::vector<std::string> files;
    if (is_dir(input)) {
        DIR* d = opendir(input.c_str());
        if (!d) die(ERR_GENERAL, "Could not read directory " + input);
        while (dirent* e = readdir(d)) {
            const std::string name = e->d_name;
            const std::string full = input + (input.back() == '/' ? "" : "/") + name;
            const std::string ext = extension(name);
            if (is_file(full) && (ext == "fasta" || ext == "fa")) files.push_back(full);
        }
        closedir(d);
        std::sort(files.begin(), files.end());  // read_dir order is unspecified; sorted here for reproducible output
    } else {
        if (!is_file(input)) die(ERR_GENERAL, "Input is not a directory or a file");
        files.push_back(input);
    }
    std::ofstream map(map_path, std::ios::binary);
    if (!map) die(ERR_GENERAL, "Could not create " + map_path);
    std::vector<Record> recs;
    std::unordered_map<std::string, size_t> seen;
    for (const std::string& f : files) {
        const std::string species = file_stem(f);
        for (auto& kv : read_fasta(f)) {
            const std::string& value = kv.second;
            if (max_len >= 0 && value.size() > size_t(max_len)) continue;
            if (value.size() < 2) {
                msg(3, "Skipping " + kv.first + " as it is too short");
                continue;
            }
            const std::string key = sanitize_header(kv.first);
            const std::string name = hashed_name(value);
            auto it = seen.find(name);
            if (it == seen.end()) {
                seen[name] = recs.size();
                recs.push_back({name, value});
            } else {
                recs[it->second].seq = value;  // HashMap insert: a 40-bit prefix collision keeps the last sequence
            }
            map << name << '\t' << species << '\t' << key << '\n';
        }
    }
    map.flush();
    if (!map) die(ERR_GENERAL, "Could not write " + map_path);
    return recs;
}

void write_fasta(const std::string& path, const std::vector<Record>& recs) {
    std::ofstream out(path, std::ios::binary);
    if (!out) die(ERR_GENERAL, "Could not create " + path);
    for (const Record& r : recs) out << '>' << r.name << '\n' << r.seq << '\n';
    out.flush();
    if (!out) die(ERR_GENERAL, "Could not write " + path);
}

std::vector<Record> read_fasta_records(const std::string& path) {
    std::vector<Record> recs;
    bool have = false;
    Record cur;
    for_each_line(path, [&](const std::string& line) {
        if (!line.empty() && line[0] == '>') {
            if (have) recs.push_back(cur);
            cur = Record{line.substr(1), ""};
            have = true;
        } else if (have) {
            for (char c : line)
                if (!isspace(static_cast<unsigned char>(c))) cur.seq.push_back(c);
        }
    });
    if (have) recs.push_back(cur);
    return recs;
}

static void write_dbtype(const std::string& path, int32_t type) {
    std::ofstream out(path, std::ios::binary);
    out.write(reinterpret_cast<const char*>(&type), 4);  // little-endian host
    out.flush();
    if (!out) die(ERR_GENERAL, "Could not write " + path);
}

template <class F>
static void write_one_db(const std::string& db, size_t n, int32_t dbtype, F&& payload) {
    std::ofstream data(db, std::ios::binary);
    std::ofstream index(db + ".index", std::ios::binary);
    if (!data || !index) die(ERR_GENERAL, "Could not create database " + db);
    uint64_t off = 0;
    for (size_t i = 0; i < n; ++i) {
        const std::string& p = payload(i);
        data.write(p.data(), std::streamsize(p.size()));
        data.write("\n\0", 2);
        const uint64_t len = p.size() + 2;
        index << i << '\t' << off << '\t' << len << '\n';
        off += len;
    }
    data.flush();
    index.flush();
    if (!data || !index) die(ERR_GENERAL, "Could not write database " + db);
    write_dbtype(db + ".dbtype", dbtype);
}

void write_foldseek_db(const std::string& db, const std::vector<Record>& recs, const std::vector<std::string>& ss,
                       const std::string& source_name) {
    if (recs.size() != ss.size()) die(ERR_GENERAL, "internal: 3Di count differs from the record count");
    for (size_t i = 0; i < recs.size(); ++i)
        if (recs[i].seq.size() != ss[i].size()) die(ERR_GENERAL, "internal: 3Di length differs for " + recs[i].name);
    constexpr int32_t kAminoAcids = 0, kGeneric = 12;  // MMseqs2 Parameters::DBTYPE_*
    write_one_db(db, recs.size(), kAminoAcids, [&](size_t i) -> const std::string& { return recs[i].seq; });
    write_one_db(db + "_ss", recs.size(), kAminoAcids, [&](size_t i) -> const std::string& { return ss[i]; });
    write_one_db(db + "_h", recs.size(), kGeneric, [&](size_t i) -> const std::string& { return recs[i].name; });
    std::ofstream lookup(db + ".lookup", std::ios::binary);
    for (size_t i = 0; i < recs.size(); ++i) {
        const std::string& h = recs[i].name;
        size_t e = 0;
        while (e < h.size() && !isspace(static_cast<unsigned char>(h[e]))) ++e;
        lookup << i << '\t' << h.substr(0, e) << '\t' << 0 << '\n';
    }
    std::ofstream source(db + ".source", std::ios::binary);
    source << 0 << '\t' << source_name << '\n';
    lookup.flush();
    source.flush();
    if (!lookup || !source) die(ERR_GENERAL, "Could not write lookup/source of " + db);
}

void write_sequence_db(const std::string& db, const std::vector<Record>& recs, const std::string& source_name) {
    constexpr int32_t kAminoAcids = 0, kGeneric = 12;
    write_one_db(db, recs.size(), kAminoAcids, [&](size_t i) -> const std::string& { return recs[i].seq; });
    write_one_db(db + "_h", recs.size(), kGeneric, [&](size_t i) -> const std::string& { return recs[i].name; });
    std::ofstream lookup(db + ".lookup", std::ios::binary);
    for (size_t i = 0; i < recs.size(); ++i) {
        const std::string& h = recs[i].name;
        size_t e = 0;
        while (e < h.size() && !isspace(static_cast<unsigned char>(h[e]))) ++e;
        lookup << i << '\t' << h.substr(0, e) << '\t' << 0 << '\n';
    }
    std::ofstream source(db + ".source", std::ios::binary);
    source << 0 << '\t' << source_name << '\n';
    lookup.flush();
    source.flush();
    if (!lookup || !source) die(ERR_GENERAL, "Could not write lookup/source of " + db);
}

size_t create_gene_specific_fasta(const std::string& input_db, const std::string& gene_dir,
                                  const std::vector<std::string>& gene_lists, bool with_db) {
    const std::vector<std::string> names = read_db(input_db + "_h"), aa = read_db(input_db), ss = read_db(input_db + "_ss");
    if (names.size() != aa.size() || names.size() != ss.size())
        die(ERR_GENERAL, "Lengths of names, amino acid and 3di sequences in database are not same");
    std::unordered_map<std::string, size_t> idx;
    for (size_t i = 0; i < names.size(); ++i) idx[names[i]] = i;  // HashMap insert: a repeated name keeps the last entry
    size_t cnt = 0;
    for (const std::string& gene_path : gene_lists) {
        const std::string gene = file_stem(gene_path);
        const std::string out_dir = gene_dir + "/" + gene;
        mkdir_p(out_dir);
        std::vector<Record> aa_recs, ss_recs;
        for_each_line(gene_path, [&](const std::string& line) {
            if (!valid_utf8(line)) return;
            std::vector<std::string> parts;  // split_whitespace()
            size_t i = 0;
            while (i < line.size()) {
                while (i < line.size() && isspace(static_cast<unsigned char>(line[i]))) ++i;
                size_t j = i;
                while (j < line.size() && !isspace(static_cast<unsigned char>(line[j]))) ++j;
                if (j > i) parts.push_back(line.substr(i, j - i));
                i = j;
            }
            if (parts.size() != 2) die(ERR_GENERAL, "Invalid line in gene mapping file: " + line);
            auto it = idx.find(parts[0]);
            if (it == idx.end()) die(ERR_GENERAL, "Sequence " + parts[1] + " not found in the database");
            aa_recs.push_back({parts[1], aa[it->second]});
            ss_recs.push_back({parts[1], ss[it->second]});
        });
        write_fasta(out_dir + "/aa.fasta", aa_recs);
        write_fasta(out_dir + "/3di.fasta", ss_recs);
        if (with_db) {
            write_sequence_db(out_dir + "/" + gene + "_db", aa_recs, "aa.fasta");
            write_sequence_db(out_dir + "/" + gene + "_db_ss", ss_recs, "3di.fasta");
        }
        ++cnt;
    }
    return cnt;
}

std::vector<std::string> read_db(const std::string& path) {
    std::vector<std::string> out;
    std::ifstream in(path, std::ios::binary);
    if (!in) die(ERR_GENERAL, "Unable to read db " + path);
    std::string line;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (!line.empty() && line[0] == '\0') line.erase(0, 1);
        if (!line.empty()) out.push_back(line);
    }
    return out;
}

void split_by_lookup(const std::string& lookup_db, std::vector<Record>& recs, std::vector<Record>& found,
                     std::vector<std::string>& found_ss) {
    const std::string ss_path = lookup_db + "_ss";
    if (!is_file(lookup_db) || !is_file(ss_path))
        die(ERR_GENERAL, "Custom lookup database does not exist or improperly formatted.");
    msg(3, "\nLoading the database...");
    const std::vector<std::string> aa = read_db(lookup_db), ss = read_db(ss_path);
    if (aa.size() != ss.size()) die(ERR_GENERAL, "The custom lookup database is not properly formatted.");
    std::unordered_map<std::string, std::string> table;
    for (size_t i = 0; i < aa.size(); ++i) table[aa[i]] = ss[i];  // insert: a repeated sequence keeps the last 3Di
    std::vector<Record> keep;
    std::vector<std::pair<Record, std::string>> hit;
    for (Record& r : recs) {
        auto it = table.find(r.seq);
        if (it != table.end() && it->second.size() == r.seq.size()) hit.emplace_back(std::move(r), it->second);
        else keep.push_back(std::move(r));
    }
    std::sort(hit.begin(), hit.end(), [](const auto& a, const auto& b) { return a.first.name < b.first.name; });
    msg(3, std::to_string(hit.size()) + " sequences found from the lookup database");
    msg(3, std::to_string(keep.size()) + " sequences not found and will be predicted");
    recs = std::move(keep);
    for (auto& h : hit) {
        found.push_back(std::move(h.first));
        found_ss.push_back(std::move(h.second));
    }
}

void split_by_afdb(const std::string& dir, std::vector<Record>& recs, std::vector<Record>& found,
                   std::vector<std::string>& found_ss) {
    std::string md5_path = dir + "/md5";
    if (is_file(dir + "/00.tsv")) md5_path = dir;
    if (!is_file(md5_path + "/00.tsv"))
        die(ERR_FILE_NOT_FOUND, md5_path + "/00.tsv: AFDB lookup tables not found (this build does not download them; fetch the "
                                "256 md5 tables with the reference `unicore createdb --afdb-lookup` where a network exists)");
    std::vector<std::vector<std::pair<size_t, std::string>>> split(256);  // table -> (record index, md5 hex)
    for (size_t i = 0; i < recs.size(); ++i) {
        Md5 m;
        m.update(recs[i].seq.data(), recs[i].seq.size());
        m.update("\n", 1);  // the tables hash the sequence with its line feed
        const std::string hex = m.hex();
        split[std::stoul(hex.substr(0, 2), nullptr, 16)].emplace_back(i, hex);
    }
    std::vector<char> hit(recs.size(), 0);
    std::vector<std::string> hit_ss(recs.size());
    static const char* dg = "0123456789abcdef";
    for (int t = 0; t < 256; ++t) {
        if (split[t].empty()) continue;
        const std::string table = md5_path + "/" + dg[t >> 4] + dg[t & 15] + ".tsv";
        std::unordered_map<std::string, std::string> map;
        for_each_line(table, [&](const std::string& line) {
            const size_t a = line.find('\t');
            if (a == std::string::npos) return;
            const size_t b = line.find('\t', a + 1);
            map[line.substr(0, a)] = line.substr(a + 1, b == std::string::npos ? std::string::npos : b - a - 1);
        });
        for (const auto& e : split[t]) {
            auto it = map.find(e.second);
            if (it != map.end() && it->second.size() == recs[e.first].seq.size()) {
                hit[e.first] = 1;
                hit_ss[e.first] = it->second;
            }
        }
    }
    std::vector<Record> keep;
    std::vector<std::pair<Record, std::string>> conv;
    for (size_t i = 0; i < recs.size(); ++i) {
        if (hit[i]) conv.emplace_back(std::move(recs[i]), std::move(hit_ss[i]));
        else keep.push_back(std::move(recs[i]));
    }
    std::sort(conv.begin(), conv.end(), [](const auto& a, const auto& b) { return a.first.name < b.first.name; });
    msg(3, std::to_string(conv.size()) + " sequences found from the lookup tables");
    msg(3, std::to_string(keep.size()) + " sequences not found and will be predicted");
    recs = std::move(keep);
    for (auto& c : conv) {
        found.push_back(std::move(c.first));
        found_ss.push_back(std::move(c.second));
    }
}

static std::vector<std::string> split_ws(const std::string& line) {
    std::vector<std::string> parts;
    size_t i = 0;
    while (i < line.size()) {
        while (i < line.size() && isspace(static_cast<unsigned char>(line[i]))) ++i;
        size_t j = i;
        while (j < line.size() && !isspace(static_cast<unsigned char>(line[j]))) ++j;
        if (j > i) parts.push_back(line.substr(i, j - i));
        i = j;
    }
    return parts;
}

// Rust's `{}` for f64: shortest decimal that round-trips, no exponent in this range, no trailing ".0"
static std::string rust_f64(double x) {
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), x);
    return std::string(buf, r.ptr);
}

std::pair<size_t, size_t> profile_clusters(const std::string& tsv, const std::string& mapping, const std::string& out_dir,
                                           size_t threshold, bool print_copiness) {
    std::unordered_map<std::string, std::set<std::string>> gene_to_spe;
    std::set<std::string> species;
    for_each_line(mapping, [&](const std::string& line) {
        if (!valid_utf8(line)) return;
        const auto parts = split_ws(line);
        if (parts.size() < 2) die(ERR_GENERAL, "Invalid line in mapping file: " + line);
        gene_to_spe[parts[0]].insert(parts[1]);
        species.insert(parts[1]);
    });
    const size_t species_count = species.size();
    std::ofstream cop;
    if (print_copiness) {
        cop.open(out_dir + "/copiness.tsv", std::ios::binary);
        if (!cop) die(ERR_GENERAL, "Could not create " + out_dir + "/copiness.tsv");
        cop << "Query\tMultipleCopyPercent\tSingleCopyPercent\n";
    }
    std::map<std::string, int> spe_cnt;
    std::map<std::string, std::set<std::string>> gene2spe;
    std::map<std::string, size_t> spe_full_cnt;
    for (const auto& s : species) spe_full_cnt[s] = 0;
    size_t total = 0, core = 0;
    bool have = false;
    std::string cur;
    auto flush_query = [&] {
        ++total;
        size_t single = 0;
        for (const auto& kv : spe_cnt) single += kv.second == 1;
        const double sp = double(single) * 100.0 / double(species_count), mp = double(spe_cnt.size()) * 100.0 / double(species_count);
        if (g_verbosity >= 4) {
            char b[256];
            snprintf(b, sizeof(b), "Gene %s reported %.2f%% single copy and %.2f%% multiple copy", cur.c_str(), sp, mp);
            msg(4, b);
        }
        if (print_copiness) cop << cur << '\t' << rust_f64(mp) << '\t' << rust_f64(sp) << '\n';
        if (single * 100 >= threshold * species_count) {
            std::string name = cur;  // query.split('-').nth(1).unwrap_or(query)
            const size_t a = cur.find('-');
            if (a != std::string::npos) {
                const size_t b = cur.find('-', a + 1);
                name = cur.substr(a + 1, b == std::string::npos ? std::string::npos : b - a - 1);
            }
            std::ofstream out(out_dir + "/" + name + ".txt", std::ios::binary);
            if (!out) die(ERR_GENERAL, "Could not create " + out_dir + "/" + name + ".txt");
            for (const auto& kv : gene2spe)  // species-sorted (the reference iterates a HashMap: order unspecified)
                if (kv.second.size() == 1) out << *kv.second.begin() << '\t' << kv.first << '\n';
            ++core;
            for (const auto& kv : spe_cnt)
                if (kv.second == 1) {
                    auto it = spe_full_cnt.find(kv.first);
                    if (it == spe_full_cnt.end()) die(ERR_GENERAL, "Species " + kv.first + " not found in the mapping file");
                    ++it->second;
                }
        }
    };
    for_each_line(tsv, [&](const std::string& line) {
        if (!valid_utf8(line)) return;
        const auto parts = split_ws(line);
        if (parts.size() < 2) die(ERR_GENERAL, "Invalid line in tsv file: " + line);
        if (!have || parts[0] != cur) {
            if (have) flush_query();
            cur = parts[0];
            have = true;
            spe_cnt.clear();
            gene2spe.clear();
        }
        auto it = gene_to_spe.find(parts[1]);
        if (it != gene_to_spe.end())
            for (const auto& spe : it->second) {
                ++spe_cnt[spe];
                gene2spe[spe].insert(parts[1]);
            }
    });
    if (have) flush_query();
    msg(3, std::to_string(core) + " structural core genes found from " + std::to_string(total) + " candidates");
    const size_t core_threshold = (core + 1) / 2;
    for (const auto& kv : spe_full_cnt)
        if (kv.second < core_threshold && g_verbosity >= 2)
            fprintf(stderr, "Warning: Species %s has only %zu core genes out of %zu core genes\n", kv.first.c_str(), kv.second, core);
    return {core, total};
}

void write_checkpoint(const std::string& path, const std::string& content) {
    std::ofstream out(path, std::ios::binary | std::ios::trunc);
    out << content;
    out.flush();
    if (!out) die(ERR_GENERAL, "Could not write checkpoint " + path);
}
std::string read_checkpoint(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) die(ERR_GENERAL, "Could not read checkpoint " + path);
    return std::string((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
}

// One process per GPU (`--procs N`): the caller has done all the file work; here the process forks N-1 children BEFORE
// any CUDA call (the records are inherited through fork, nothing is parsed twice), stays rank 0 itself, hands the
// NCCL id to the children through pipes, and every rank predicts its count-shard; ONE ncclAllGather inside
// p5_predict_sharded leaves all 3Di strings on every rank.  Rank 0 returns them to the caller (the only DB writer),
// the children exit.  The communicator is set up on a thread of its own while the weights load.
namespace {
struct ChildWatch {  // a rank that dies would leave the others waiting inside NCCL for ever
    std::vector<pid_t> pids;
    std::atomic<bool> done{false};
    std::thread th;
    void start() {
        th = std::thread([this] {
            size_t alive = pids.size();
            while (alive > 0) {
                for (pid_t& pid : pids) {
                    if (pid <= 0) continue;
                    int st = 0;
                    const pid_t r = waitpid(pid, &st, WNOHANG);
                    if (r == pid) {
                        pid = 0;
                        --alive;
                        if (!(WIFEXITED(st) && WEXITSTATUS(st) == 0)) {
                            fprintf(stderr, "Error: a prediction rank failed (%s %d)\n", WIFEXITED(st) ? "exit code" : "signal",
                                    WIFEXITED(st) ? WEXITSTATUS(st) : WTERMSIG(st));
                            for (pid_t q : pids)
                                if (q > 0) kill(q, SIGTERM);
                            _exit(ERR_GENERAL);
                        }
                    }
                }
                if (alive) usleep(done.load() ? 2000 : 50000);
            }
        });
    }
    void join() {
        done = true;
        if (th.joinable()) th.join();
    }
};
}  // namespace

static void apply_options(p5_model* m, const PredictOptions& opt, const std::vector<Record>& recs, bool talk) {
    if (opt.max_batch_tokens > 0 && p5_set_option(m, "max_batch_tokens", opt.max_batch_tokens) != 0)
        die(ERR_GENERAL, p5_last_error());
    if (opt.map_rare_to_x >= 0 && p5_set_option(m, "map_rare_to_x", opt.map_rare_to_x) != 0) die(ERR_GENERAL, p5_last_error());
    if (opt.head_include_eos >= 0 && p5_set_option(m, "head_include_eos", opt.head_include_eos) != 0) die(ERR_GENERAL, p5_last_error());
    if (!talk) return;
    size_t n_long = 0;
    for (const Record& r : recs) n_long += opt.split_len > 0 && r.seq.size() > opt.split_len;
    if (n_long)
        msg(3, std::to_string(n_long) + " sequences are longer than " + std::to_string(opt.split_len) +
                   " residues and are predicted in chunks of that length (Foldseek's --prostt5-split-length default, which the "
                   "reference relies on; --split-len 0 / --prostt5-split-length 0 predicts them in one piece)");
}

std::vector<std::string> predict_3di(const std::string& model_dir, const std::vector<Record>& recs,
                                     const PredictOptions& opt) {
    using clk = std::chrono::steady_clock;
    std::vector<int> devs = opt.devices;
    const int world = opt.procs > 1 ? opt.procs : 1;
    int rank = 0;
    std::vector<int> id_pipes;  // rank 0: write ends towards ranks 1..N-1
    int my_pipe = -1;
    ChildWatch watch;
    if (world > 1) {
        if (devs.empty())
            for (int r = 0; r < world; ++r) devs.push_back(r);
        if (int(devs.size()) != world) die(ERR_ARGPARSE, "--procs N needs exactly N devices in --devices");
        fflush(stdout);
        fflush(stderr);
        for (int r = 1; r < world; ++r) {
            int fd[2];
            if (pipe(fd) != 0) die(ERR_GENERAL, "pipe failed");
            const pid_t pid = fork();
            if (pid < 0) die(ERR_GENERAL, "fork failed");
            if (pid == 0) {  // rank r
                close(fd[1]);
                for (int w : id_pipes) close(w);
                rank = r;
                my_pipe = fd[0];
                id_pipes.clear();
                watch.pids.clear();
                break;
            }
            close(fd[0]);
            id_pipes.push_back(fd[1]);
            watch.pids.push_back(pid);
        }
        if (rank == 0) watch.start();
    }
    auto fail = [&](const std::string& what) {
        if (rank == 0) die(ERR_GENERAL, what);
        fprintf(stderr, "Error (rank %d): %s\n", rank, what.c_str());
        _exit(ERR_GENERAL);
    };

    const auto t0 = clk::now();
    p5_comm* comm = nullptr;
    std::thread comm_thread;
    std::string comm_err;
    double comm_s = 0;
    if (world > 1) {
        uint8_t id[P5_COMM_ID_BYTES];
        if (rank == 0) {
            if (p5_comm_unique_id(id) != 0) fail(std::string("NCCL: ") + p5_last_error());
            for (int w : id_pipes) {
                if (write(w, id, sizeof id) != ssize_t(sizeof id)) fail("could not hand the NCCL id to a rank");
                close(w);
            }
        } else {
            size_t got = 0;
            while (got < sizeof id) {
                const ssize_t n = read(my_pipe, id + got, sizeof id - got);
                if (n <= 0) fail("could not read the NCCL id from rank 0");
                got += size_t(n);
            }
            close(my_pipe);
        }
        const int dev = devs[rank];
        comm_thread = std::thread([&, dev] {  // ~0.4 s of NCCL set-up, hidden behind the weight load
            const auto c0 = clk::now();
            uint8_t idc[P5_COMM_ID_BYTES];
            memcpy(idc, id, sizeof idc);
            if (p5_comm_create(idc, rank, world, dev, &comm) != 0) comm_err = p5_last_error();
            comm_s = std::chrono::duration<double>(clk::now() - c0).count();
        });
        // the thread copies `id` first thing; keep it alive until the join below
        p5_model* m = nullptr;
        const int one = dev;
        const int rc = p5_model_load(model_dir.c_str(), &one, 1, &m);
        const std::string load_err = rc != 0 ? p5_last_error() : "";
        comm_thread.join();
        if (rc != 0) fail("ProstT5 model: " + load_err);
        if (!comm_err.empty()) fail("NCCL communicator: " + comm_err);
        const auto t1 = clk::now();
        apply_options(m, opt, recs, rank == 0);
        std::vector<uint64_t> off(recs.size() + 1, 0);
        for (size_t i = 0; i < recs.size(); ++i) off[i + 1] = off[i] + recs[i].seq.size();
        std::string aa;
        aa.reserve(off.back());
        for (const Record& r : recs) aa += r.seq;
        std::string out(aa.size(), '\0');
        if (p5_predict_sharded(m, comm, reinterpret_cast<const uint8_t*>(aa.data()), off.data(), recs.size(),
                               reinterpret_cast<uint8_t*>(&out[0]), opt.split_len) != 0)
            fail(std::string("ProstT5 prediction failed: ") + p5_last_error());
        const auto t2 = clk::now();
        double st[14] = {0};
        p5_get_stats(m, st, 14);
        p5_model_free(m);
        p5_comm_free(comm);
        if (rank != 0) {
            fflush(stdout);
            fflush(stderr);
            _exit(0);
        }
        watch.join();
        std::vector<std::string> ss(recs.size());
        for (size_t i = 0; i < recs.size(); ++i) ss[i] = out.substr(off[i], off[i + 1] - off[i]);
        const double load_s = std::chrono::duration<double>(t1 - t0).count();
        const double pred_s = std::chrono::duration<double>(t2 - t1).count();
        msg(3, "ProstT5: " + std::to_string(recs.size()) + " sequences, " + std::to_string(off.back()) + " residues on " +
                   std::to_string(world) + " ranks in " + std::to_string(pred_s) + " s (" +
                   std::to_string(off.back() / std::max(pred_s, 1e-9)) + " residues/s incl. the NCCL all-gather; weights + communicator in " +
                   std::to_string(load_s) + " s, communicator alone " + std::to_string(comm_s) + " s)");
        if (!opt.stats_json.empty()) {
            std::ofstream js(opt.stats_json);
            js << "{\"sequences\": " << recs.size() << ", \"residues\": " << off.back() << ", \"ranks\": " << world
               << ", \"predict_seconds\": " << pred_s << ", \"load_seconds\": " << load_s << ", \"comm_init_seconds\": " << comm_s
               << ", \"residues_per_second\": " << off.back() / std::max(pred_s, 1e-9) << ", \"rank0_batches\": " << st[0]
               << ", \"rank0_tokens\": " << st[1] << ", \"rank0_kernel_launches\": " << st[3] << ", \"rank0_device_ms\": " << st[4]
               << "}\n";
        }
        return ss;
    }

    p5_model* m = nullptr;
    if (p5_model_load(model_dir.c_str(), devs.empty() ? nullptr : devs.data(), devs.empty() ? -1 : int(devs.size()), &m) != 0)
        die(ERR_GENERAL, std::string("ProstT5 model: ") + p5_last_error());
    const auto t1 = clk::now();
    apply_options(m, opt, recs, true);
    std::vector<uint64_t> off(recs.size() + 1, 0);
    for (size_t i = 0; i < recs.size(); ++i) off[i + 1] = off[i] + recs[i].seq.size();
    std::string aa;
    aa.reserve(off.back());
    for (const Record& r : recs) aa += r.seq;
    std::string out(aa.size(), '\0');
    if (p5_predict(m, reinterpret_cast<const uint8_t*>(aa.data()), off.data(), recs.size(),
                   reinterpret_cast<uint8_t*>(&out[0]), opt.split_len) != 0)
        die(ERR_GENERAL, std::string("ProstT5 prediction failed: ") + p5_last_error());
    const auto t2 = clk::now();
    double st[14] = {0};
    p5_get_stats(m, st, 14);
    p5_model_free(m);
    std::vector<std::string> ss(recs.size());
    for (size_t i = 0; i < recs.size(); ++i) ss[i] = out.substr(off[i], off[i + 1] - off[i]);
    const double load_s = std::chrono::duration<double>(t1 - t0).count();
    const double pred_s = std::chrono::duration<double>(t2 - t1).count();
    msg(3, "ProstT5: " + std::to_string(recs.size()) + " sequences, " + std::to_string(off.back()) + " residues in " +
               std::to_string(pred_s) + " s (" + std::to_string(off.back() / std::max(pred_s, 1e-9)) +
               " residues/s; weights loaded in " + std::to_string(load_s) + " s)");
    if (!opt.stats_json.empty()) {
        std::ofstream js(opt.stats_json);
        js << "{\"sequences\": " << recs.size() << ", \"residues\": " << off.back() << ", \"predict_seconds\": " << pred_s
           << ", \"load_seconds\": " << load_s << ", \"residues_per_second\": " << off.back() / std::max(pred_s, 1e-9)
           << ", \"batches\": " << st[0] << ", \"tokens\": " << st[1] << ", \"kernel_launches\": " << st[3]
           << ", \"device_ms_max\": " << st[4] << "}\n";
    }
    return ss;
}

}  // namespace ub

// Host side of `unicore createdb` in C++ (the reference's is Rust; no Rust toolchain exists in this
// image): FASTA parsing, record naming, .map file, checkpoint, MMseqs/Foldseek DB writer.  Each function
// cites the reference code whose behaviour it keeps.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace ub {

// error codes of the reference CLI [REF src/envs/error_handler.rs:5-14]
enum : int {
    ERR_GENERAL = 0x01, ERR_FILE_NOT_FOUND = 0x10, ERR_FILE_INVALID = 0x11, ERR_BINARY_NOT_FOUND = 0x20,
    ERR_MODULE_NOT_IMPLEMENTED = 0x30, ERR_ARGPARSE = 0x40, ERR_OUTPUT_EXISTS = 0x50,
};

extern int g_verbosity;  // 0 quiet, 1 +errors, 2 +warnings, 3 +info, 4 +debug [REF src/util/message.rs:4-22]
void msg(int level, const std::string& s);                 // stdout when verbosity >= level
[[noreturn]] void die(int code, const std::string& what);  // "<label>: what" on stderr, exit(code)

// [REF src/seq/fasta_io.rs:6-25] records in first-appearance order; a repeated header keeps its first
// position and takes the LAST sequence (HashMap insert semantics); the trailing record is always
// inserted, so an empty file yields one ("", "") record.
std::vector<std::pair<std::string, std::string>> read_fasta(const std::string& path);

// [REF src/modules/createdb.rs:15-18,101] whitespace (Unicode White_Space) ; : , = / ( ) -> '_'
std::string sanitize_header(const std::string& key);

// [REF src/modules/createdb.rs:104-106]
std::string hashed_name(const std::string& seq);

struct Record {  // one DB entry
    std::string name;  // header line without '>'
    std::string seq;
};

// [REF src/modules/createdb.rs:68-111] input listing, length filters, naming, de-duplication, .map file.
// Returns the unique records in first-appearance order.
std::vector<Record> collect_records(const std::string& input, const std::string& map_path, long max_len);

// [REF src/seq/fasta_io.rs:27-48] ">name\nseq\n"
void write_fasta(const std::string& path, const std::vector<Record>& recs);

// plain FASTA reader for the foldseek-argv shim (every record kept, in file order)
std::vector<Record> read_fasta_records(const std::string& path);

// MMseqs2/Foldseek sequence DB triple <db>, <db>_h, <db>_ss (+ .index, .dbtype, .lookup, .source)
// (SURVEY.md §8a row DBW); entries share one physical order in the three data files, which is what the
// reference's reader relies on [REF src/seq/create_gene_specific_fasta.rs:9-44].
void write_foldseek_db(const std::string& db, const std::vector<Record>& recs, const std::vector<std::string>& ss,
                       const std::string& source_name);

// [REF src/seq/create_gene_specific_fasta.rs:9-25]
std::vector<std::string> read_db(const std::string& path);

// [REF src/seq/afdb_lookup.rs:131-181] run_custom: `lookup_db` and `lookup_db`_ss are read with read_db and
// zipped into an (amino-acid string -> 3Di string) table; records whose sequence is in the table move to
// `found` (+ their 3Di in `found_ss`), sorted by name; `recs` keeps the ones that must be predicted.
void split_by_lookup(const std::string& lookup_db, std::vector<Record>& recs, std::vector<Record>& found,
                     std::vector<std::string>& found_ss);

// Sequence DB without a 3Di companion: <db>, <db>_h (+ .index, .dbtype, .lookup, .source) — what
// `foldseek base:createdb <fasta> <db> --shuffle 0` writes for the per-gene FASTA files of `unicore tree`
// [REF src/modules/tree.rs:78-110].
void write_sequence_db(const std::string& db, const std::vector<Record>& recs, const std::string& source_name);

// [REF src/seq/create_gene_specific_fasta.rs:27-88] for every gene list file (lines "<db name> <label>") writes
// <gene_dir>/<gene>/aa.fasta and 3di.fasta from the DB triple; with_db additionally writes the per-gene
// Foldseek DBs <gene>_db and <gene>_db_ss natively.  Returns the number of genes.
size_t create_gene_specific_fasta(const std::string& input_db, const std::string& gene_dir,
                                  const std::vector<std::string>& gene_lists, bool with_db);

// [REF src/modules/profile.rs:13-147] taxonomic profile of the clusters: `mapping` is createdb's .map (gene ->
// species), `tsv` the cluster/search table (query, target per line, grouped by query).  Writes copiness.tsv
// (optional) and one <gene>.txt per structural core gene (single-copy in >= threshold % of the species).
// Returns (core genes, candidates).
std::pair<size_t, size_t> profile_clusters(const std::string& tsv, const std::string& mapping, const std::string& out_dir,
                                           size_t threshold, bool print_copiness);

// [REF src/seq/afdb_lookup.rs:50-129] run_afdb with the tables already on disk (`<dir>/md5/XX.tsv` or `<dir>/XX.tsv`,
// lines "<md5 hex of sequence + LF>\t<3Di>"; XX = first byte of that md5).  This build has no network code: missing
// tables are an error instead of a 30 GB download.  Same outputs as split_by_lookup.
void split_by_afdb(const std::string& dir, std::vector<Record>& recs, std::vector<Record>& found,
                   std::vector<std::string>& found_ss);

// [REF src/util/checkpoint.rs:2-10]
void write_checkpoint(const std::string& path, const std::string& content);
std::string read_checkpoint(const std::string& path);

bool path_exists(const std::string& p);
bool is_dir(const std::string& p);
bool is_file(const std::string& p);
std::string parent_dir(const std::string& p);  // Rust Path::parent(): "" for a bare file name
std::string file_stem(const std::string& p);
std::string base_name(const std::string& p);
void mkdir_p(const std::string& p);
std::vector<std::string> list_files_with_ext(const std::string& dir, const std::string& ext);  // sorted

// runs the predictor over the records: returns one 3Di string per record (calls the C ABI)
struct PredictOptions {
    std::vector<int> devices;  // empty = every visible device
    int procs = 0;             // > 1: one process per GPU (fork), count-sharding + one NCCL all-gather
    // Foldseek's own default applies in the reference, which passes no split flag [REF src/modules/createdb.rs:158-166]:
    // --prostt5-split-length is believed to default to 1024 (unverified here): longer sequences are
    // predicted in consecutive chunks.  0 = never split (full-length attention).
    uint32_t split_len = 1024;
    int map_rare_to_x = -1;  // -1 = library default (U, Z, O, B -> X)
    int head_include_eos = -1;  // -1 = library default (the </s> row is part of the CNN head's input: SURVEY.md Q3)
    int64_t max_batch_tokens = 0;  // 0 = library default
    std::string stats_json;        // optional path
};
std::vector<std::string> predict_3di(const std::string& model_dir, const std::vector<Record>& recs,
                                     const PredictOptions& opt);

}  // namespace ub

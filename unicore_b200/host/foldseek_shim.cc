// foldseek-b200: answers the argv the reference's createdb sends to its `foldseek` binary
//   createdb <fasta> <outdb> --prostt5-model <dir> --threads N [--gpu 1]     [REF src/modules/createdb.rs:158-166]
//   version                                                                 [REF src/modules/config.rs:49-60]
// so an UNMODIFIED reference `unicore` uses the B200 path after `unicore config --set-foldseek <this>`
// [REF src/modules/config.rs:62-79].  Every other verb is handed to the real foldseek named by
// $UNICORE_B200_REAL_FOLDSEEK (cluster, createtsv, search, ... [REF src/modules/cluster.rs:45-72]).
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "host.h"

using namespace ub;

static int passthrough(char** argv) {
    const char* real = getenv("UNICORE_B200_REAL_FOLDSEEK");
    if (!real || !*real) {
        fprintf(stderr, "foldseek-b200: `%s` is not implemented here and UNICORE_B200_REAL_FOLDSEEK is not set\n", argv[1]);
        return 1;
    }
    argv[0] = const_cast<char*>(real);
    execv(real, argv);
    perror("foldseek-b200: exec of the real foldseek failed");
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 2) { fputs("foldseek-b200 createdb <fasta> <db> --prostt5-model <dir> [--threads N] [--gpu 1]\n", stderr); return 1; }
    const std::string verb = argv[1];
    if (verb == "version") { puts("foldseek-b200 0.1.0"); return 0; }
    bool prostt5 = false;
    for (int i = 2; i < argc; ++i) prostt5 |= !strcmp(argv[i], "--prostt5-model");
    if (verb == "base:createdb" && !getenv("UNICORE_B200_REAL_FOLDSEEK")) {
        // `foldseek base:createdb <fasta> <db> --shuffle 0 -v V` of `unicore tree` [REF src/modules/tree.rs:87-105]:
        // plain sequence DB, answered natively when no real foldseek is configured
        std::vector<std::string> pos;
        for (int i = 2; i < argc; ++i) {
            const std::string a = argv[i];
            if (a == "--shuffle" || a == "-v" || a == "--threads" || a == "--dbtype" || a == "--compressed") ++i;
            else if (!a.empty() && a[0] == '-') { fprintf(stderr, "foldseek-b200: unknown option %s\n", a.c_str()); return 1; }
            else pos.push_back(a);
        }
        if (pos.size() != 2 || !is_file(pos[0])) { fputs("foldseek-b200: base:createdb needs <fasta> <db>\n", stderr); return 1; }
        write_sequence_db(pos[1], read_fasta_records(pos[0]), base_name(pos[0]));
        return 0;
    }
    if (verb != "createdb" || !prostt5) return passthrough(argv);
    std::vector<std::string> pos;
    std::string model;
    PredictOptions popt;
    for (int i = 2; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&]() -> std::string {
            if (i + 1 >= argc) { fprintf(stderr, "foldseek-b200: %s needs a value\n", a.c_str()); exit(1); }
            return argv[++i];
        };
        if (a == "--prostt5-model") model = value();
        else if (a == "--threads" || a == "--gpu" || a == "--shuffle" || a == "--dbtype" || a == "--compressed") value();
        else if (a == "-v") g_verbosity = atoi(value().c_str());
        else if (a == "--prostt5-split-length") popt.split_len = uint32_t(atol(value().c_str()));  // default 1024 (host.h)
        else if (a == "--prostt5-rare-residues") popt.map_rare_to_x = value() == "own" ? 0 : 1;
        else if (a == "--prostt5-head-eos") popt.head_include_eos = atoi(value().c_str()) != 0;
        else if (!a.empty() && a[0] == '-') { fprintf(stderr, "foldseek-b200: unknown option %s\n", a.c_str()); return 1; }
        else pos.push_back(a);
    }
    if (pos.size() != 2) { fputs("foldseek-b200: createdb needs <fasta> <db>\n", stderr); return 1; }
    if (!is_file(pos[0])) { fprintf(stderr, "foldseek-b200: input %s does not exist\n", pos[0].c_str()); return 1; }
    std::vector<Record> recs = read_fasta_records(pos[0]);
    std::vector<std::string> ss = predict_3di(model, recs, popt);
    write_foldseek_db(pos[1], recs, ss, base_name(pos[0]));
    return 0;
}

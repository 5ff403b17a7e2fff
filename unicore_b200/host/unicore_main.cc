// unicore-b200: the `createdb` verb of the reference CLI [REF src/util/arg_parser.rs:183-215;
// src/modules/createdb.rs:20-217] with the ProstT5 step done in-process on B200 GPUs through
// libprostt5_b200.so instead of a spawned `foldseek createdb --prostt5-model`.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "host.h"

using namespace ub;

static const char* kUsage =
    "Create Foldseek database from amino acid sequences\n\n"
    "Usage: unicore-b200 createdb [OPTIONS] <INPUT> <OUTPUT> <MODEL>\n\n"
    "Arguments:\n"
    "  <INPUT>   Input directory with fasta files or a single fasta file\n"
    "  <OUTPUT>  Output foldseek database\n"
    "  <MODEL>   ProstT5 model (directory holding prostt5-f16.gguf)\n\n"
    "Options:\n"
    "  -k, --keep                    Keep intermediate files\n"
    "  -o, --overwrite               Force overwrite output database\n"
    "      --max-len <MAX_LEN>       Set maximum sequence length threshold\n"
    "  -g, --gpu                     Accepted for compatibility (this build always runs on the GPU)\n"
    "      --afdb-lookup <PATH>      Use AFDB lookup tables already on disk (<PATH>/md5/XX.tsv); no download in this build\n"
    "      --custom-lookup <PATH>    Use custom lookup database, accepts any Foldseek database to reference against\n"
    "      --threads <THREADS>       Accepted for compatibility [default: 0]\n"
    "  -v, --verbosity <VERBOSITY>   0: quiet, 1: +errors, 2: +warnings, 3: +info, 4: +debug [default: 3]\n"
    "      --devices <LIST>          Comma-separated CUDA devices (default: all visible)\n"
    "      --procs <N>               One process per GPU on N GPUs: sequences sharded by count, one NCCL all-gather of\n"
    "                                the 3Di strings, this process writes the DB (default: threads in one process)\n"
    "      --split-len <N>           Predict sequences longer than N residues in chunks of N [default: 1024, Foldseek's\n"
    "                                --prostt5-split-length default, which the reference relies on]; 0 = never split\n"
    "      --rare-residues <x|own>   U, Z, O, B tokenise as X (default, ProstT5's preprocessing) or as their own tokens\n"
    "      --head-eos <0|1>          Whether the </s> row is part of the CNN head's input [default: 1]\n"
    "      --max-batch-tokens <N>    Tokens per forward pass\n"
    "      --stats-json <PATH>       Write throughput counters as JSON\n";

static int createdb(int argc, char** argv) {
    std::vector<std::string> pos;
    bool keep = false, overwrite = false;
    long max_len = -1;
    std::string afdb, custom;
    PredictOptions popt;
    for (int i = 0; i < argc; ++i) {
        const std::string a = argv[i];
        auto value = [&](const char* name) -> std::string {
            if (i + 1 >= argc) die(ERR_ARGPARSE, std::string("createdb - ") + name);
            return argv[++i];
        };
        if (a == "-k" || a == "--keep") keep = true;
        else if (a == "-o" || a == "--overwrite") overwrite = true;
        else if (a == "-g" || a == "--gpu") {}
        else if (a == "--max-len") max_len = atol(value("max_len").c_str());
        else if (a == "--afdb-lookup") afdb = value("afdb_lookup");
        else if (a == "--custom-lookup") custom = value("custom_lookup");
        else if (a == "--threads") value("threads");
        else if (a == "-v" || a == "--verbosity") g_verbosity = atoi(value("verbosity").c_str());
        else if (a == "--devices") {
            const std::string v = value("devices");
            size_t s = 0;
            while (s <= v.size()) {
                size_t e = v.find(',', s);
                if (e == std::string::npos) e = v.size();
                if (e > s) popt.devices.push_back(atoi(v.substr(s, e - s).c_str()));
                s = e + 1;
            }
        } else if (a == "--procs") popt.procs = atoi(value("procs").c_str());
        else if (a == "--rare-residues") popt.map_rare_to_x = value("rare_residues") == "own" ? 0 : 1;
        else if (a == "--head-eos") popt.head_include_eos = atoi(value("head_eos").c_str()) != 0;
        else if (a == "--split-len") popt.split_len = uint32_t(atol(value("split_len").c_str()));
        else if (a == "--max-batch-tokens") popt.max_batch_tokens = atol(value("max_batch_tokens").c_str());
        else if (a == "--stats-json") popt.stats_json = value("stats_json");
        else if (a == "-h" || a == "--help") { fputs(kUsage, stdout); return 0; }
        else pos.push_back(a);
    }
    if (pos.size() < 1) die(ERR_ARGPARSE, "createdb - input");
    if (pos.size() < 2) die(ERR_ARGPARSE, "createdb - output");
    if (pos.size() < 3) die(ERR_ARGPARSE, "createdb - model");
    if (pos.size() > 3) die(ERR_ARGPARSE, "createdb - unexpected argument " + pos[3]);
    const std::string input = pos[0], output = pos[1], model = pos[2];
    if (!afdb.empty() && !custom.empty())
        die(ERR_ARGPARSE, "Both afdb_lookup and custom_lookup are specified. Please specify only one.");

    std::string parent = parent_dir(output);
    if (parent.empty()) parent = ".";
    msg(4, "Parent directory: " + parent);
    mkdir_p(parent);

    const std::string chk = parent + "/createdb.chk";
    if (path_exists(chk)) {
        if (read_checkpoint(chk) == "1" && !overwrite) die(ERR_GENERAL, "Database already exists, skipping createdb module");
    } else {
        write_checkpoint(chk, "0");
    }
    std::vector<Record> recs = collect_records(input, output + ".map", max_len);
    // The reference builds this path as curr_dir/parent/combined_aa.fasta, which breaks for absolute
    // outputs [REF src/modules/createdb.rs:115-127]; the intermediate belongs next to the output.
    const std::string combined = parent + "/combined_aa.fasta";

    if (path_exists(model + "/cnn.safetensors") || path_exists(model + "/model/cnn.safetensors"))
        die(ERR_GENERAL, "Old weight files detected from the given path. Please provide different path for the model weights");
    if (!path_exists(model + "/prostt5-f16.gguf"))
        die(ERR_FILE_NOT_FOUND, model + "/prostt5-f16.gguf (this build does not download weights; run `foldseek databases ProstT5 " +
                                    model + " tmp` where a network exists)");
    // --custom-lookup [REF src/seq/afdb_lookup.rs:131-181]: sequences present in the lookup DB take its 3Di,
    // only the rest is predicted; the final DB holds the predicted entries followed by the converted ones
    // (header-sorted), which is what base:createdb + base:concatdbs produce [REF createdb.rs:168-205].
    std::vector<Record> found;
    std::vector<std::string> found_ss;
    if (!custom.empty()) split_by_lookup(custom, recs, found, found_ss);
    if (!afdb.empty()) split_by_afdb(afdb, recs, found, found_ss);
    write_fasta(combined, recs);
    std::vector<std::string> ss;
    if (!recs.empty()) ss = predict_3di(model, recs, popt);
    else msg(3, "Every sequence was found in the lookup database: nothing to predict");
    recs.insert(recs.end(), found.begin(), found.end());
    ss.insert(ss.end(), found_ss.begin(), found_ss.end());
    write_foldseek_db(output, recs, ss, base_name(combined));
    if (!keep) remove(combined.c_str());
    write_checkpoint(chk, "1");
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2 || !strcmp(argv[1], "help") || !strcmp(argv[1], "-h") || !strcmp(argv[1], "--help")) {
        fputs("unicore-b200: B200-native `unicore createdb`\n\nCommands:\n  createdb   Create Foldseek database from amino acid sequences\n"
              "  profile    Taxonomic profiling of the clusters: core structures (reference `unicore profile`)\n"
              "  genefasta  Per-gene aa/3Di FASTA (+ DBs) from a database and a profile directory (first step of `unicore tree`)\n"
              "  version    Print version\n\n", stdout);
        fputs(kUsage, stdout);
        return argc < 2 ? ERR_ARGPARSE : 0;
    }
    const std::string verb = argv[1];
    if (verb == "version") { puts("unicore-b200 0.1.0 (unicore v1.1.1 createdb contract)"); return 0; }
    if (verb == "profile") {
        // [REF src/util/arg_parser.rs:274-293; src/modules/profile.rs:149-171]
        std::vector<std::string> pos;
        size_t threshold = 80;
        bool copiness = true;
        for (int i = 2; i < argc; ++i) {
            const std::string a = argv[i];
            if ((a == "-t" || a == "--threshold") && i + 1 < argc) {
                const long t = atol(argv[++i]);
                if (t < 0 || t > 100) die(ERR_ARGPARSE, "profile - threshold must be in 0..100");
                threshold = size_t(t);
            } else if ((a == "-p" || a == "--print-copiness") && i + 1 < argc) copiness = std::string(argv[++i]) != "false";
            else if (a == "--threads" && i + 1 < argc) ++i;
            else if ((a == "-v" || a == "--verbosity") && i + 1 < argc) g_verbosity = atoi(argv[++i]);
            else pos.push_back(a);
        }
        if (pos.size() < 1) die(ERR_ARGPARSE, "profile - input");
        if (pos.size() < 2) die(ERR_ARGPARSE, "profile - mapping");
        if (pos.size() < 3) die(ERR_ARGPARSE, "profile - output");
        mkdir_p(pos[2]);
        write_checkpoint(pos[2] + "/profile.chk", "0");
        profile_clusters(pos[1], pos[0] + ".map", pos[2], threshold, copiness);
        write_checkpoint(pos[2] + "/profile.chk", "1");
        return 0;
    }
    if (verb == "genefasta") {
        // tree-side consumer of the DB [REF src/modules/tree.rs:57-110]: unicore-b200 genefasta <db> <profile_dir> <out_dir> [--db]
        std::vector<std::string> pos;
        bool with_db = false;
        for (int i = 2; i < argc; ++i) {
            if (!strcmp(argv[i], "--db")) with_db = true;
            else if (!strcmp(argv[i], "-v") && i + 1 < argc) g_verbosity = atoi(argv[++i]);
            else pos.push_back(argv[i]);
        }
        if (pos.size() != 3) die(ERR_ARGPARSE, "genefasta <db> <profile_dir> <out_dir> [--db]");
        std::vector<std::string> lists = list_files_with_ext(pos[1], "txt");
        const size_t n = create_gene_specific_fasta(pos[0], pos[2] + "/fasta", lists, with_db);
        msg(3, "Gene specific fasta files prepared in: " + pos[2] + "/fasta (" + std::to_string(n) + " genes)");
        return 0;
    }
    if (verb == "createdb") {
        if (argc == 2) { fputs(kUsage, stdout); return ERR_ARGPARSE; }
        return createdb(argc - 2, argv + 2);
    }
    die(ERR_MODULE_NOT_IMPLEMENTED, verb + " (this build replaces the createdb hot path only; use the reference unicore for the other modules)");
}

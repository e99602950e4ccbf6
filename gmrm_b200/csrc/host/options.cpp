// Command line of gmrm (behaviour of src/options.cpp): "--flag value" pairs, unknown flag
// fatal, missing value of the last flag fatal, echo under "ardyh command line options:".
#include <sys/stat.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "host.hpp"

namespace host {

namespace {
[[noreturn]] void fatal(const std::string& msg) {
    std::cout << msg << std::endl;
    exit(EXIT_FAILURE);
}
const char* value_of(int argc, char** argv, int& i) {
    if (i == argc - 1)
        fatal(std::string("FATAL  : missing argument for last option \"") + argv[i] + "\". Please check your input and relaunch.");
    return argv[++i];
}
unsigned positive(const char* flag, const char* v, int min, const char* what) {
    if (atoi(v) < min) fatal(std::string("FATAL  : option ") + flag + " has to be a " + what + " integer! (" + v + " was passed)");
    return (unsigned)atoi(v);
}
std::string trim(const std::string& s) {
    const char* ws = " \t\r\n\f\v";
    const size_t b = s.find_first_not_of(ws);
    if (b == std::string::npos) return "";
    return s.substr(b, s.find_last_not_of(ws) - b + 1);
}
std::vector<std::string> split_ws(const std::string& s) {
    std::vector<std::string> out;
    std::istringstream is(s);
    std::string tok;
    while (is >> tok) out.push_back(tok);
    return out;
}

void read_group_mixture_file(Options& o, int rank) {
    std::ifstream f(o.group_mixture_file);
    if (!f.is_open()) {
        printf("FATAL  : can not open the mixture file %s. Use the --group-mixture-file option!\n", o.group_mixture_file.c_str());
        exit(1);
    }
    if (rank == 0) std::cout << "INFO   : Reading group mixtures from [" + o.group_mixture_file + "]." << std::endl;
    std::vector<std::vector<std::string>> rows;
    std::string line;
    int k0 = -1;
    while (getline(f, line)) {
        line = trim(line);
        if (line.empty()) continue;
        std::vector<std::string> t = split_ws(line);
        if (k0 < 0) k0 = (int)t.size();
        if ((int)t.size() != k0) {
            printf("FATAL  : check your mixture file. The same number of mixtures is expected for all groups.\n");
            printf("       : got %d mixtures for group %d, while first group had %d.\n", (int)t.size(), (int)rows.size(), k0);
            exit(1);
        }
        rows.push_back(t);
    }
    o.ngroups = (int)rows.size();
    o.nmixtures = k0 < 0 ? 0 : k0;
    o.cva.assign((size_t)o.ngroups * o.nmixtures, 0.0);
    for (int g = 0; g < o.ngroups; g++)
        for (int k = 0; k < o.nmixtures; k++) {
            const double v = std::stod(rows[g][k]);
            o.cva[(size_t)g * o.nmixtures + k] = v;
            if (k == 0 && v != 0.0) {
                printf("FATAL  : First element of group mixture must be 0.0! Check your input file %s.\n", o.group_mixture_file.c_str());
                exit(1);
            }
            if (k > 0 && v <= o.cva[(size_t)g * o.nmixtures + k - 1]) {
                printf("FATAL  : Mixtures must be given in ascending order! Check your input file %s.\n", o.group_mixture_file.c_str());
                exit(1);
            }
        }
}
}  // namespace

Options parse_options(int argc, char** argv, int rank) {
    Options o;
    std::stringstream ss;
    ss << "\nardyh command line options:\n";
    for (int i = 1; i < argc; ++i) {
        const char* a = argv[i];
        if (!strcmp(a, "--bed-file")) { o.bed_file = value_of(argc, argv, i); ss << "--bed-file " << o.bed_file << "\n"; }
        else if (!strcmp(a, "--dim-file")) { o.dim_file = value_of(argc, argv, i); ss << "--dim-file " << o.dim_file << "\n"; }
        else if (!strcmp(a, "--phen-files")) {
            const std::string list = value_of(argc, argv, i);
            ss << "--phen-files " << list << "\n";
            std::stringstream sl(list);
            std::string fp;
            while (getline(sl, fp, ',')) {
                std::ifstream t(fp);
                if (!t.is_open()) fatal("FATAL: file " + fp + " not found");
                o.phen_files.push_back(fp);
            }
        }
        else if (!strcmp(a, "--group-index-file")) { o.group_index_file = value_of(argc, argv, i); ss << "--group-index-file " << o.group_index_file << "\n"; }
        else if (!strcmp(a, "--group-mixture-file")) { o.group_mixture_file = value_of(argc, argv, i); ss << "--group-mixture-file " << o.group_mixture_file << "\n"; }
        else if (!strcmp(a, "--verbosity")) { o.verbosity = atoi(value_of(argc, argv, i)); ss << "--verbosity " << o.verbosity << "\n"; }
        else if (!strcmp(a, "--shuffle-markers")) { o.shuffle = atoi(value_of(argc, argv, i)) != 0; ss << "--shuffle-markers " << o.shuffle << "\n"; }
        else if (!strcmp(a, "--mimic-hydra")) { o.mimic_hydra = true; ss << "--mimic-hydra 1\n"; }
        else if (!strcmp(a, "--seed")) { o.seed = positive("--seed", value_of(argc, argv, i), 0, "positive"); ss << "--seed " << o.seed << "\n"; }
        else if (!strcmp(a, "--iterations")) { o.iterations = positive("--iterations", value_of(argc, argv, i), 1, "strictly positive"); ss << "--iterations " << o.iterations << "\n"; }
        else if (!strcmp(a, "--trunc-markers")) { o.truncm = positive("--trunc-markers", value_of(argc, argv, i), 1, "strictly positive"); ss << "--trunc-markers " << o.truncm << "\n"; }
        else if (!strcmp(a, "--S")) {
            const std::string list = value_of(argc, argv, i);
            std::stringstream sl(list);
            std::string tok;
            while (getline(sl, tok, ',')) {
                const double v = std::stod(tok);
                if (!(v > 0.0) || (!o.S.empty() && !(v > o.S.back()))) fatal("FATAL  : option --S expects strictly positive, ascending values");
                o.S.push_back(v);
            }
            ss << "--S " << list << "\n";
        }
        else if (!strcmp(a, "--out-dir")) {
            o.out_dir = value_of(argc, argv, i);
            struct stat st;
            if (stat(o.out_dir.c_str(), &st) != 0) mkdir(o.out_dir.c_str(), 0755);
            ss << "--out-dir " << o.out_dir << "\n";
        }
        else if (!strcmp(a, "--output-thin-rate")) { o.thin = positive("--output-thin-rate", value_of(argc, argv, i), 1, "strictly positive"); ss << "--output-thin-rate " << o.thin << "\n"; }
        else if (!strcmp(a, "--predict")) { o.predict = true; ss << "--predict 1\n"; }
        else if (!strcmp(a, "--bim-file")) { o.bim_file = value_of(argc, argv, i); ss << "--bim-file " << o.bim_file << "\n"; }
        else if (!strcmp(a, "--ref-bim-file")) { o.ref_bim_file = value_of(argc, argv, i); ss << "--ref-bim-file " << o.ref_bim_file << "\n"; }
        // ---- supersets of the reference's flag list
        else if (!strcmp(a, "--vranks")) { o.vranks = (int)positive("--vranks", value_of(argc, argv, i), 1, "strictly positive"); ss << "--vranks " << o.vranks << "\n"; }
        else if (!strcmp(a, "--sync-rate")) { o.sync_rate = (int)positive("--sync-rate", value_of(argc, argv, i), 1, "strictly positive"); ss << "--sync-rate " << o.sync_rate << "\n"; }
        else if (!strcmp(a, "--burn-in")) { o.burn_in = positive("--burn-in", value_of(argc, argv, i), 0, "positive"); ss << "--burn-in " << o.burn_in << "\n"; }
        else if (!strcmp(a, "--check-inputs")) { o.check_inputs = true; ss << "--check-inputs 1\n"; }
        else if (!strcmp(a, "--dump-inputs")) { o.dump_inputs = value_of(argc, argv, i); ss << "--dump-inputs " << o.dump_inputs << "\n"; }
        else if (!strcmp(a, "--selftest-predict")) { o.selftest_predict = true; ss << "--selftest-predict 1\n"; }
        else if (!strcmp(a, "--selftest-outputs")) { o.selftest_outputs = true; ss << "--selftest-outputs 1\n"; }
        else if (!strcmp(a, "--replay-file")) { o.replay_file = value_of(argc, argv, i); ss << "--replay-file " << o.replay_file << "\n"; }
        else if (!strcmp(a, "--gpus")) { o.gpus = (int)positive("--gpus", value_of(argc, argv, i), 1, "strictly positive"); ss << "--gpus " << o.gpus << "\n"; }
        else fatal(std::string("FATAL: option \"") + a + "\" unknown");
    }
    o.echo = ss.str();
    if (rank == 0) std::cout << o.echo << std::endl;

    // options.cpp:175-220
    if (o.bed_file.empty()) fatal("FATAL  : no bed file provided! Please use the --bed-file option.");
    if (o.dim_file.empty()) fatal("FATAL  : no dim file provided! Please use the --dim-file option.");
    if (o.phen_files.empty()) fatal("FATAL  : no phen file(s) provided! Please use the --phen-files option.");
    if (!o.predict && (o.group_index_file.empty() != o.group_mixture_file.empty()))
        fatal("FATAL  : you need to activate BOTH --group-index-file and --group-mixture-file");
    if (o.predict) {                                                              // options.cpp:205-214
        if (o.bim_file.empty()) fatal("FATAL  : you need to pass a bim file with --bim-file when activating --predict");
        if (o.ref_bim_file.empty()) fatal("FATAL  : you need to pass a reference bim file with --ref-bim-file when activating --predict");
    }
    if (o.mimic_hydra && o.phen_files.size() > 1) fatal("FATAL  : with --mimic-hydra, only a single phenotype can be processed.");
    if (rank == 0 && o.mimic_hydra)   // the flag only picks the reference's Mersenne-Twister seeds and shuffle stream (bayes.cpp:798-799, phenotype.cpp:316)
        std::cout << "WARNING: --mimic-hydra selects the reference's PRNG seeding; this build draws from counter-based Philox streams "
                     "(or replays logged variates, --replay-file): the flag has no effect here." << std::endl;
    if (rank == 0 && !o.S.empty())    // options.hpp:35: get_s() has no caller in the reference either
        std::cout << "WARNING: --S is parsed and checked but not used (the reference never reads it either)." << std::endl;
    if (o.predict) return o;                                                      // options.hpp:11-12: the mixture file is not read with --predict
    read_group_mixture_file(o, rank);
    return o;
}

}  // namespace host

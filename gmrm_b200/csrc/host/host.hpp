// Host side of gmrm_b200: the reference's process interface in front of the C ABI.
// Same flags, same input formats, same output files and stdout lines as medical-genomics-group/gmrm;
// the numerical work is delegated to libgmrm_b200.so (include/gmrm_b200.h).  Written from the
// behaviour of src/options.cpp, src/dimensions.cpp, src/phenotype.cpp:587-673, src/bayes.cpp:830-853
// and src/xfiles.{hpp,cpp}; no code is shared with them.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace host {

struct Options {
    // reference flags (src/options.cpp:26-156)
    std::string bed_file, dim_file, bim_file, ref_bim_file, group_index_file, group_mixture_file, out_dir;
    std::vector<std::string> phen_files;
    int verbosity = 0;
    bool shuffle = true, mimic_hydra = false, predict = false;
    unsigned seed = 0, iterations = 1, truncm = 0, thin = 1;
    std::vector<double> S;
    // supersets (defaults reproduce the reference: one exchange per marker-step)
    int vranks = 0;          // --vranks: total virtual ranks; 0 = 2048 per GPU
    int gpus = 1;            // --gpus: GPUs of this node sharing the chain (one host thread each)
    int sync_rate = 1;       // --sync-rate
    unsigned burn_in = 0;    // --burn-in: iterations left out of the posterior-mean summary (.mbet)
    std::string replay_file;     // --replay-file: logged variates of a reference run (layout: host.hpp, Replay), fed instead of the Philox streams
    bool check_inputs = false;   // --check-inputs: parse everything, print a summary, no GPU work
    std::string dump_inputs;     // --dump-inputs <dir>: with --check-inputs, write the parsed phenotypes / groups / mixtures as raw binaries (tests)
    bool selftest_predict = false;   // --selftest-predict: with --check-inputs and --predict, read the .bim pair and the .bet histories and print what was found (tests)
    bool selftest_outputs = false;   // --selftest-outputs: with --check-inputs, write a fixed 2-iteration .csv/.bet/.cpn history from 2 "ranks" (tests)
    // derived from the .grm file
    int ngroups = 0, nmixtures = 0;
    std::vector<double> cva;     // [G][K]
    std::string echo;            // the "ardyh command line options" block
};

// Parses argv exactly as the reference does (unknown flag, missing value, bad value: message + exit(1)),
// checks the option set (options.cpp:175-220) and reads the mixture file (options.cpp:222-286).
Options parse_options(int argc, char** argv, int rank);

struct Dims { int N = 0, Mt = 0; };
Dims read_dim_file(const std::string& path, unsigned truncm);          // dimensions.cpp:8-29, dimensions.hpp:13-15

struct Phen {
    std::string path, stem;
    std::vector<double> eps;       // 4*ceil(N/4) slots, centred/scaled, 0 at NA and in the pad
    std::vector<uint8_t> mask4;    // ceil(N/4) bytes
    int nonas = 0, nas = 0;
};
Phen read_phen_file(const std::string& path, int N, int verbosity);    // phenotype.cpp:587-673

std::vector<int32_t> read_group_index_file(const std::string& path, int G, int Mt);   // bayes.cpp:830-853

// reads markers [begin, begin+count) of a PLINK .bed (3 magic bytes skipped, bayes.cpp:882) in chunks
class BedReader {
public:
    BedReader(const std::string& path, int N);
    ~BedReader();
    void read(int marker_begin, int count, uint8_t* dst);
    int mbytes() const { return mbytes_; }
private:
    int fd_ = -1, mbytes_ = 0;
};

// .csv/.bet/.cpn writers with the reference's layouts (xfiles.cpp:17-45, xfiles.hpp:24-37)
class OutFiles {
public:
    OutFiles(const std::string& out_dir, const std::string& stem, bool create);   // rank 0 deletes + creates
    ~OutFiles();
    void write_csv(unsigned it, unsigned nthinned, const double* sigmag, int G, double sigmae, int m0_sum, const double* pi, int K);
    void write_bet(unsigned Mt, unsigned it, unsigned nthinned, int first, int count, const double* betas, bool is_rank0);
    void write_cpn(unsigned Mt, unsigned it, unsigned nthinned, int first, int count, const int32_t* comp, bool is_rank0);
    std::string csv_path, bet_path, cpn_path;
private:
    int csv_ = -1, bet_ = -1, cpn_ = -1;
};

// ---- replay of a reference run's random variates (north-star replay mode).
// File: 8 bytes "GMRMRPL1", int32 R, Mm, T, G, K, iterations; T*G doubles sigmag_init; then per iteration
// perm int32[R*Mm], u f64[Mm*R*T], z f64[Mm*R*T], mu_draw f64[T], sigg_unit f64[T*G], pi_unit f64[T*G*K], sige_unit f64[T]
// -- the arrays of gmrm_replay (include/gmrm_b200.h), NaN where the reference drew nothing.
struct Replay {
    int R = 0, Mm = 0, T = 0, G = 0, K = 0, iterations = 0;
    std::vector<double> sigmag_init;
    struct It { std::vector<int32_t> perm; std::vector<double> u, z, mu_draw, sigg_unit, pi_unit, sige_unit; };
    std::vector<It> its;
};
Replay read_replay_file(const std::string& path);

// Default number of virtual ranks (markers in flight per step) for Mt markers on `gpus` GPUs: 2,048 per GPU, but never more than
// Mt / 64 in total -- every virtual rank keeps at least 64 markers of its own, so a step samples at most 1/64 of the markers
// against the same residuals (with Mt ranks the sampler would degenerate into a fully synchronous sweep).
int default_vranks(int Mt, int gpus);

// ---- association pass, the reference's --predict mode (predict.cpp)
struct BimCross {
    std::vector<std::string> ids;            // ids of --bim-file in row order = global marker order
    std::map<std::string, int> ref_index;    // id -> row in --ref-bim-file
};
BimCross cross_bim_files(const std::string& bim, const std::string& ref_bim, bool verbose);     // bayes.cpp:286-316
std::vector<double> read_bet_mean(const std::string& path, size_t expect_mt, unsigned* niter);  // bayes.cpp:38-78
constexpr int kMlmaLine = 123;                                                                  // LLEN - 1, bayes.cpp:218
std::string mlma_line(const std::string& id, int mglo, int rmglo, double beta, double tdist, double se, double pval);

}  // namespace host

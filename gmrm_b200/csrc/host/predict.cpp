// Host side of the association pass, the reference's --predict mode (Bayes::predict, src/bayes.cpp:14-284;
// Bayes::cross_bim_files, 286-316; file names phenotype.cpp:115-127): reads the .bim pair and the .bet history, and
// writes <stem>.mlma with the reference's fixed-width lines.  The sums run on the GPU (gmrm_predict).
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "host.hpp"

namespace host {

// Rows of a .bim are "chr id genPos physPos allele1 allele2" (bayes.cpp:296-300); reading stops at the first row that
// does not parse, as the reference's stream extraction does.
static std::vector<std::string> read_bim_ids(const std::string& path) {
    std::ifstream in(path);
    if (!in) {
        printf("FATAL  : can not open the file [%s] to read.\n", path.c_str());
        exit(EXIT_FAILURE);
    }
    std::vector<std::string> ids;
    std::string id, a1, a2;
    unsigned chr, pos;
    float gpos;
    while (in >> chr >> id >> gpos >> pos >> a1 >> a2) ids.push_back(id);
    return ids;
}

BimCross cross_bim_files(const std::string& bim, const std::string& ref_bim, bool verbose) {
    BimCross x;
    if (verbose) {
        printf("INFO   : bim file:     %s\n", bim.c_str());
        printf("INFO   : ref bim file: %s\n", ref_bim.c_str());
    }
    x.ids = read_bim_ids(bim);
    if (verbose) printf("INFO   : found %d ids in bim file\n", (int)x.ids.size());
    const std::vector<std::string> ref = read_bim_ids(ref_bim);
    for (size_t i = 0; i < ref.size(); i++) x.ref_index[ref[i]] = (int)i;     // a repeated id keeps its last row (310-312)
    if (verbose) printf("INFO   : found %d ids in reference bim file\n", (int)ref.size());
    return x;
}

// Mean over the recorded iterations of every marker's beta (bayes.cpp:38-78).  The file is uint32 Mt, then per saved
// iteration uint32 it + Mt doubles (xfiles.hpp:24-37).
std::vector<double> read_bet_mean(const std::string& path, size_t expect_mt, unsigned* niter_out) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) {
        printf("FATAL  : could not open the .bet file %s\n", path.c_str());
        exit(EXIT_FAILURE);
    }
    struct stat st;
    fstat(fd, &st);
    uint32_t mt = 0;
    if (pread(fd, &mt, 4, 0) != 4) mt = 0;
    if (mt != expect_mt) {                                                        // bayes.cpp:45-48
        printf("Mismatch between expected and Mtot read from .bet file: %lu vs %d\n", (unsigned long)expect_mt, (int)mt);
        exit(1);
    }
    const size_t rec = 4 + (size_t)mt * 8;
    if ((size_t)st.st_size < 4 || ((size_t)st.st_size - 4) % rec != 0) {          // the reference asserts this (50)
        printf("FATAL  : %s is not a whole number of iterations of %u markers\n", path.c_str(), mt);
        exit(EXIT_FAILURE);
    }
    const unsigned niter = (unsigned)(((size_t)st.st_size - 4) / rec);
    std::vector<double> sum(mt, 0.0), it(mt);
    for (unsigned i = 0; i < niter; i++) {
        const off_t off = 4 + (off_t)rec * i + 4;
        size_t done = 0;
        while (done < (size_t)mt * 8) {
            const ssize_t r = pread(fd, (char*)it.data() + done, (size_t)mt * 8 - done, off + (off_t)done);
            if (r <= 0) {
                printf("FATAL  : short read from %s\n", path.c_str());
                exit(EXIT_FAILURE);
            }
            done += (size_t)r;
        }
        for (uint32_t j = 0; j < mt; j++) sum[j] += it[j];
    }
    close(fd);
    for (auto& v : sum) v /= (double)niter;                                       // 0 iterations: NaN, as in the reference
    if (niter_out) *niter_out = niter;
    return sum;
}

// One line of the .mlma (bayes.cpp:230-236): 123 bytes, or empty if a value does not fit its column
std::string mlma_line(const std::string& id, int mglo, int rmglo, double beta, double tdist, double se, double pval) {
    char buf[512];
    const int n = snprintf(buf, sizeof buf, "%20s %8d %8d %20.15f %20.15f %20.15f %20.15f\n", id.c_str(), mglo, rmglo, beta, tdist, se, pval);
    if (n != kMlmaLine) return std::string();
    return std::string(buf, (size_t)n);
}

}  // namespace host

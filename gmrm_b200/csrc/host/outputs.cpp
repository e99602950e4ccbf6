// Output writers with the reference's layouts: .csv (src/xfiles.cpp:17-45), .bet / .cpn history files
// (write_ofile_h1, src/xfiles.hpp:24-37); files are deleted and re-created at start (src/bayes.cpp:323-324).
#include <fcntl.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "host.hpp"

namespace host {

namespace {
int open_out(const std::string& path, bool create) {
    if (create) unlink(path.c_str());                                            // delete_output_files
    const int fd = open(path.c_str(), create ? (O_CREAT | O_WRONLY | O_EXCL) : O_WRONLY, 0644);   // MPI_MODE_CREATE | WRONLY | EXCL
    if (fd < 0) {
        printf("FATAL  : could not open output file %s\n", path.c_str());
        exit(EXIT_FAILURE);
    }
    return fd;
}
void pwrite_all(int fd, const void* buf, size_t n, off_t off) {
    const char* p = (const char*)buf;
    size_t done = 0;
    while (done < n) {
        const ssize_t w = pwrite(fd, p + done, n - done, off + (off_t)done);
        if (w <= 0) {
            printf("FATAL  : write to an output file failed\n");
            exit(EXIT_FAILURE);
        }
        done += (size_t)w;
    }
}
}  // namespace

OutFiles::OutFiles(const std::string& out_dir, const std::string& stem, bool create) {
    const std::string base = (out_dir.empty() ? std::string() : out_dir + "/") + stem;
    csv_path = base + ".csv"; bet_path = base + ".bet"; cpn_path = base + ".cpn";
    csv_ = open_out(csv_path, create);
    bet_ = open_out(bet_path, create);
    cpn_ = open_out(cpn_path, create);
}
OutFiles::~OutFiles() {
    if (csv_ >= 0) close(csv_);
    if (bet_ >= 0) close(bet_);
    if (cpn_ >= 0) close(cpn_);
}

void OutFiles::write_csv(unsigned it, unsigned nthinned, const double* sigmag, int G, double sigmae, int m0_sum, const double* pi, int K) {
    std::string line;
    char b[64];
    snprintf(b, sizeof b, "%5d, %4d", it, G); line += b;
    double sg_sum = 0.0;
    for (int g = 0; g < G; g++) { snprintf(b, sizeof b, ", %20.15f", sigmag[g]); line += b; sg_sum += sigmag[g]; }
    snprintf(b, sizeof b, ", %20.15f, %20.15f, %7d, %4d, %2d", sigmae, sg_sum / (sigmae + sg_sum), m0_sum, G, K); line += b;
    for (int i = 0; i < G * K; i++) { snprintf(b, sizeof b, ", %20.15f", pi[i]); line += b; }
    line += "\n";
    pwrite_all(csv_, line.data(), line.size(), (off_t)nthinned * (off_t)line.size());    // xfiles.cpp:45: constant line length assumed
}

template <class T>
static void write_h1(int fd, unsigned Mt, unsigned it, unsigned nthinned, int first, int count, const T* data, bool is_rank0) {
    const size_t rec = sizeof(unsigned) + (size_t)Mt * sizeof(T);
    if (is_rank0) {
        if (nthinned == 0) pwrite_all(fd, &Mt, sizeof Mt, 0);
        pwrite_all(fd, &it, sizeof it, (off_t)(sizeof(unsigned) + (size_t)nthinned * rec));
    }
    pwrite_all(fd, data, (size_t)count * sizeof(T), (off_t)(2 * sizeof(unsigned) + (size_t)nthinned * rec + (size_t)first * sizeof(T)));
}
void OutFiles::write_bet(unsigned Mt, unsigned it, unsigned nthinned, int first, int count, const double* betas, bool is_rank0) {
    write_h1(bet_, Mt, it, nthinned, first, count, betas, is_rank0);
}
void OutFiles::write_cpn(unsigned Mt, unsigned it, unsigned nthinned, int first, int count, const int32_t* comp, bool is_rank0) {
    write_h1(cpn_, Mt, it, nthinned, first, count, comp, is_rank0);
}

}  // namespace host

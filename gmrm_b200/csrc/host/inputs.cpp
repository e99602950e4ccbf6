// Input readers of the reference's process interface, restated from their behaviour:
//   .dim   Dimensions::read_dim_file        src/dimensions.cpp:8-29   (+ --trunc-markers, dimensions.hpp:13-15)
//   .phen  Phenotype::read_file             src/phenotype.cpp:587-673
//   .gri   Bayes::read_group_index_file     src/bayes.cpp:830-853
//   .bed   Bayes::load_genotype             src/bayes.cpp:867-900     (3 magic bytes skipped, not validated)
#include <fcntl.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>

#include "host.hpp"

namespace host {

Dims read_dim_file(const std::string& path, unsigned truncm) {
    std::ifstream f(path);
    if (!f.is_open()) {
        std::cout << "FATAL: could not open dim file: " << path << std::endl;
        exit(EXIT_FAILURE);
    }
    std::string line;
    getline(f, line);
    std::istringstream is(line);
    std::vector<std::string> tok;
    std::string t;
    while (is >> t) tok.push_back(t);
    if (tok.size() != 2) {
        std::cout << "FATAL: dim file should contain a single line with 2 integers" << std::endl;
        exit(EXIT_FAILURE);
    }
    Dims d;
    d.N = atoi(tok[0].c_str());
    d.Mt = atoi(tok[1].c_str());
    if (truncm > 0 && (int)truncm < d.Mt) d.Mt = (int)truncm;      // --trunc-markers
    return d;
}

Phen read_phen_file(const std::string& path, int N, int verbosity) {
    Phen p;
    p.path = path;
    const size_t slash = path.find_last_of('/');
    const std::string base = slash == std::string::npos ? path : path.substr(slash + 1);
    const size_t dot = base.find_last_of('.');
    p.stem = (dot == std::string::npos || dot == 0) ? base : base.substr(0, dot);   // fs::path::stem
    std::ifstream f(path);
    if (!f.is_open()) {
        std::cout << "FATAL: could not open phenotype file: " << path << std::endl;
        exit(EXIT_FAILURE);
    }
    const int im4 = (N + 3) / 4;
    p.mask4.clear();
    std::vector<double> data;
    data.reserve(N);
    const double NA = std::numeric_limits<double>::max();
    double sum = 0.0;
    std::string line;
    int line_n = 0;
    while (getline(f, line)) {
        const int m4 = line_n % 4;
        if (m4 == 0) p.mask4.push_back(0x0F);
        std::istringstream is(line);
        std::string fid, iid, val;
        is >> fid >> iid >> val;
        if (val == "NA") {
            p.nas++;
            data.push_back(NA);
            if (verbosity >= 2) std::cout << " ... found NA on line " << line_n << ", m4 = " << m4 << " on byte " << line_n / 4 << std::endl;
            p.mask4[line_n / 4] &= (uint8_t)~(1u << m4);
        } else {
            p.nonas++;
            const double v = atof(val.c_str());
            data.push_back(v);
            sum += v;
        }
        line_n++;
    }
    if (p.nas + p.nonas != N) {
        printf("FATAL  : phenotype file %s has %d rows, the dim file says %d individuals.\n", path.c_str(), p.nas + p.nonas, N);
        exit(EXIT_FAILURE);
    }
    const int m4 = line_n % 4;
    if (m4 != 0) {
        for (int i = m4; i < 4; i++) p.mask4[line_n / 4] &= (uint8_t)~(1u << i);
        std::cout << "Setting last " << 4 - m4 << " bits to NAs" << std::endl;
    }
    p.mask4.resize(im4, 0);
    // centre and scale (phenotype.cpp:647-667); the pad slots stay 0
    p.eps.assign((size_t)im4 * 4, 0.0);
    const double avg = sum / (double)p.nonas;
    double sqn = 0.0;
    for (size_t i = 0; i < data.size(); i++) {
        if (data[i] == NA) {
            p.eps[i] = 0.0;
        } else {
            p.eps[i] = data[i] - avg;
            sqn += p.eps[i] * p.eps[i];
        }
    }
    sqn = sqrt((double)(p.nonas - 1) / sqn);
    for (size_t i = 0; i < data.size(); i++) p.eps[i] *= sqn;
    return p;
}

std::vector<int32_t> read_group_index_file(const std::string& path, int G, int Mt) {
    std::ifstream f(path.c_str());
    if (!f) {
        std::cout << "FATAL  : can not open the group file [" << path << "] to read. Use the --group-index-file option!" << std::endl;
        exit(EXIT_FAILURE);
    }
    std::vector<int32_t> gi;
    std::string label;
    int group;
    while (f >> label >> group) {
        if (group >= G || group < 0) {      // the reference tests `group > G`; G itself would index past the tables
            printf("FATAL  : group index file contains a value that exceeds the number of groups given in group mixture file.\n");
            printf("       : check the consistency between your group index and mixture input files.\n");
            exit(1);
        }
        gi.push_back(group);
    }
    if ((int)gi.size() < Mt) {
        printf("FATAL  : group index file %s lists %d markers, %d are needed.\n", path.c_str(), (int)gi.size(), Mt);
        exit(1);
    }
    gi.resize(Mt);                          // --trunc-markers keeps the first Mt
    return gi;
}

BedReader::BedReader(const std::string& path, int N) : mbytes_((N + 3) / 4) {
    fd_ = open(path.c_str(), O_RDONLY);
    if (fd_ < 0) {
        printf("FATAL  : could not open bed file %s\n", path.c_str());
        exit(EXIT_FAILURE);
    }
}
BedReader::~BedReader() {
    if (fd_ >= 0) close(fd_);
}
void BedReader::read(int marker_begin, int count, uint8_t* dst) {
    size_t want = (size_t)count * mbytes_, done = 0;
    const off_t off = 3 + (off_t)marker_begin * mbytes_;      // bayes.cpp:882
    while (done < want) {
        const ssize_t r = pread(fd_, dst + done, want - done, off + (off_t)done);
        if (r <= 0) {
            printf("FATAL  : short read from the bed file (marker %d + %d, %zu of %zu bytes)\n", marker_begin, count, done, want);
            exit(EXIT_FAILURE);
        }
        done += (size_t)r;
    }
}

Replay read_replay_file(const std::string& path) {
    Replay r;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { printf("FATAL  : can not open the replay file %s\n", path.c_str()); exit(EXIT_FAILURE); }
    auto need = [&](void* dst, size_t bytes) {
        if (fread(dst, 1, bytes, f) != bytes) { printf("FATAL  : replay file %s is truncated\n", path.c_str()); exit(EXIT_FAILURE); }
    };
    char magic[8];
    need(magic, 8);
    if (memcmp(magic, "GMRMRPL1", 8) != 0) { printf("FATAL  : %s is not a replay file\n", path.c_str()); exit(EXIT_FAILURE); }
    int32_t h[6];
    need(h, sizeof h);
    r.R = h[0]; r.Mm = h[1]; r.T = h[2]; r.G = h[3]; r.K = h[4]; r.iterations = h[5];
    if (r.R < 1 || r.Mm < 1 || r.T < 1 || r.G < 1 || r.K < 2 || r.iterations < 0) { printf("FATAL  : bad header in replay file %s\n", path.c_str()); exit(EXIT_FAILURE); }
    const size_t T = r.T, G = r.G, K = r.K, nuz = (size_t)r.Mm * r.R * T;
    r.sigmag_init.resize(T * G);
    need(r.sigmag_init.data(), T * G * 8);
    r.its.resize(r.iterations);
    for (auto& it : r.its) {
        it.perm.resize((size_t)r.R * r.Mm); it.u.resize(nuz); it.z.resize(nuz);
        it.mu_draw.resize(T); it.sigg_unit.resize(T * G); it.pi_unit.resize(T * G * K); it.sige_unit.resize(T);
        need(it.perm.data(), it.perm.size() * 4); need(it.u.data(), nuz * 8); need(it.z.data(), nuz * 8);
        need(it.mu_draw.data(), T * 8); need(it.sigg_unit.data(), T * G * 8); need(it.pi_unit.data(), T * G * K * 8); need(it.sige_unit.data(), T * 8);
    }
    fclose(f);
    return r;
}

int default_vranks(int Mt, int gpus) {
    long long vr = 2048LL * gpus;
    const long long cap = Mt / 64;
    if (vr > cap) vr = cap;
    vr -= vr % gpus;
    if (vr < gpus) vr = gpus;
    return (int)vr;
}

}  // namespace host

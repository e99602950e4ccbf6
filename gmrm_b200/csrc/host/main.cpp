// gmrm_b200_cli -- drop-in for the reference executable (src/main.cpp:8-24): the Gibbs mode (Bayes::process,
// src/bayes.cpp:318-677) and the association pass (--predict, Bayes::predict, src/bayes.cpp:14-284) with the same
// flags, inputs, outputs and stdout lines; the marker sums run on B200s through the C ABI (include/gmrm_b200.h).
// One host thread per GPU drives its shard of the markers.
#include <fcntl.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <iostream>
#include <mutex>
#include <thread>

#include "gmrm_b200.h"
#include "host.hpp"

namespace {

class Barrier {
public:
    explicit Barrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        const int gen = gen_;
        if (++count_ == n_) { count_ = 0; gen_++; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen != gen_; });
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    int n_, count_ = 0, gen_ = 0;
};

void ck(int rc, const char* what) {
    if (rc != 0) {
        printf("FATAL  : %s: %s\n", what, gmrm_last_error());
        exit(EXIT_FAILURE);
    }
}
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Shared {
    host::Options opt;
    host::Dims dims;
    std::vector<host::Phen> phens;
    std::vector<int32_t> group_index;
    uint8_t nccl_id[128];
    void* peer_ptrs[8][6] = {};
    int vranks = 0;
    host::Replay replay;                       // --replay-file (empty: production Philox streams)
    // --predict
    host::BimCross bim;
    std::vector<uint8_t> keep;                 // [Mt] 1 = the marker's id is in the reference .bim
    std::vector<std::vector<double>> bet_mean; // [T][Mt]
};

// this GPU's block of the .bed into the engine (Bayes::load_genotype, bayes.cpp:867-900), in chunks: two pinned buffers take
// turns -- chunk k+1 is read from the file (a few threads, disjoint pread ranges) while chunk k is uploaded and transcoded
void load_block(gmrm_engine* e, const host::Options& o, int N, int S, int M) {
    host::BedReader bed(o.bed_file, N);
    size_t chunk_bytes = 256u << 20;
    if (const char* v = getenv("GMRM_CLI_CHUNK_KB")) chunk_bytes = (size_t)std::max(1, atoi(v)) << 10;   // tests: many small chunks
    const int chunk = std::max(1, (int)(chunk_bytes / (size_t)bed.mbytes()));
    // pinned read buffers: the upload is then plain DMA (gmrm_host_alloc, include/gmrm_b200.h)
    const size_t cap = (size_t)std::min(chunk, std::max(M, 1)) * bed.mbytes();
    uint8_t* buf[2] = {(uint8_t*)gmrm_host_alloc(cap), M > chunk ? (uint8_t*)gmrm_host_alloc(cap) : nullptr};
    if (!buf[0] || (M > chunk && !buf[1])) ck(-4, "gmrm_host_alloc");
    const int nreaders = 4;
    auto read_chunk = [&](int k) {
        const int done = k * chunk, n = std::min(chunk, M - done);
        std::vector<std::thread> th;
        for (int r = 0; r < nreaders; r++) {
            const int lo = (int)((int64_t)n * r / nreaders), hi = (int)((int64_t)n * (r + 1) / nreaders);
            if (hi > lo) th.emplace_back([&, lo, hi] { bed.read(S + done + lo, hi - lo, buf[k & 1] + (size_t)lo * bed.mbytes()); });
        }
        for (auto& t : th) t.join();
    };
    const int nchunks = (M + chunk - 1) / chunk;
    if (nchunks > 0) read_chunk(0);
    for (int k = 0; k < nchunks; k++) {
        const int done = k * chunk, n = std::min(chunk, M - done);
        std::thread next;
        if (k + 1 < nchunks) next = std::thread(read_chunk, k + 1);      // the other buffer: its upload (chunk k-1) has returned
        ck(gmrm_upload_bed(e, buf[k & 1], S + done, n), "gmrm_upload_bed");
        if (next.joinable()) next.join();
    }
    gmrm_host_free(buf[0]);
    if (buf[1]) gmrm_host_free(buf[1]);
    ck(gmrm_finalize_bed(e), "gmrm_finalize_bed");
}

// --predict (Bayes::predict, bayes.cpp:14-284): per trait, genetic values from the mean of the .bet history, then the
// per-marker association statistics, written to <stem>.mlma in global marker order.
void predict_worker(int rank, Shared* sh, Barrier* bar) {
    const host::Options& o = sh->opt;
    const int ngpu = o.gpus, T = (int)sh->phens.size(), N = sh->dims.N, Mt = sh->dims.Mt;
    const double t_start = now();
    gmrm_config cfg{};
    cfg.device = rank; cfg.N = N; cfg.Mt = Mt; cfg.T = T; cfg.G = 1; cfg.K = 2;      // groups and mixtures play no part here
    cfg.world_size = ngpu; cfg.world_rank = rank; cfg.vranks = sh->vranks; cfg.sync_rate = 1; cfg.shuffle = 0; cfg.seed = o.seed;
    gmrm_engine* e = nullptr;
    ck(gmrm_create(&cfg, &e), "gmrm_create");
    if (ngpu > 1) {
        if (rank == 0) ck(gmrm_comm_unique_id(sh->nccl_id), "gmrm_comm_unique_id");
        bar->wait();
        ck(gmrm_comm_init(e, sh->nccl_id), "gmrm_comm_init");
    }
    int32_t S = 0, M = 0;
    ck(gmrm_shard_info(e, &S, &M, nullptr, nullptr, nullptr), "gmrm_shard_info");
    const double t_load = now();
    load_block(e, o, N, S, M);
    if (rank == 0) printf("INFO   : time to load genotype data = %.3f seconds.\n", now() - t_load);
    for (int t = 0; t < T; t++) ck(gmrm_set_phenotype(e, t, sh->phens[t].eps.data(), sh->phens[t].mask4.data(), sh->phens[t].nonas), "gmrm_set_phenotype");
    const double t_stats = now();
    ck(gmrm_compute_marker_stats(e), "gmrm_compute_marker_stats");
    if (rank == 0) printf("INFO   : Time to compute the markers' statistics: %.2f seconds.\n", now() - t_stats);

    int kept_before = 0;                                 // lines the shards before this one write (bayes.cpp:240-248)
    for (int j = 0; j < S; j++) kept_before += sh->keep[j];
    std::vector<double> beta(std::max(M, 1)), tdist(std::max(M, 1)), se(std::max(M, 1)), pval(std::max(M, 1));
    for (int t = 0; t < T; t++) {
        const std::string path = (o.out_dir.empty() ? std::string() : o.out_dir + "/") + sh->phens[t].stem + ".mlma";
        if (rank == 0) {                                 // delete_output_prediction_files + exclusive create (phenotype.cpp:146-170)
            unlink(path.c_str());
            const int fd = open(path.c_str(), O_CREAT | O_WRONLY | O_EXCL, 0644);
            if (fd < 0) { printf("FATAL  : could not open output file %s\n", path.c_str()); exit(EXIT_FAILURE); }
            close(fd);
        }
        bar->wait();
        ck(gmrm_predict(e, t, sh->phens[t].eps.data(), sh->bet_mean[t].data() + S, sh->keep.data() + S, nullptr, beta.data(), tdist.data(),
                        se.data(), pval.data()), "gmrm_predict");
        std::string text;
        text.reserve((size_t)M * host::kMlmaLine);
        for (int j = 0; j < M; j++) {
            const std::string& id = sh->bim.ids[S + j];
            if (!sh->keep[S + j]) {
                printf("WARNING: marker id %s excluded -- no match\n", id.c_str());            // bayes.cpp:227
                continue;
            }
            const std::string line = host::mlma_line(id, S + j, sh->bim.ref_index.at(id), beta[j], tdist[j], se[j], pval[j]);
            if (line.empty()) { printf("FATAL  : the .mlma line of marker %s does not fit the fixed width of %d bytes\n", id.c_str(), host::kMlmaLine); exit(EXIT_FAILURE); }
            text += line;
        }
        const int fd = open(path.c_str(), O_WRONLY);
        if (fd < 0) { printf("FATAL  : could not open output file %s\n", path.c_str()); exit(EXIT_FAILURE); }
        size_t done = 0;
        while (done < text.size()) {
            const ssize_t w = pwrite(fd, text.data() + done, text.size() - done, (off_t)kept_before * host::kMlmaLine + (off_t)done);
            if (w <= 0) { printf("FATAL  : write to %s failed\n", path.c_str()); exit(EXIT_FAILURE); }
            done += (size_t)w;
        }
        close(fd);
        bar->wait();
    }
    if (rank == 0) printf("INFO   : Time to compute the predictions: %.2f seconds.\n", now() - t_start);
    gmrm_destroy(e);
}

void worker(int rank, Shared* sh, Barrier* bar) {
    const host::Options& o = sh->opt;
    const int ngpu = o.gpus, T = (int)sh->phens.size(), G = o.ngroups, K = o.nmixtures, N = sh->dims.N, Mt = sh->dims.Mt;
    gmrm_config cfg{};
    cfg.device = rank; cfg.N = N; cfg.Mt = Mt; cfg.T = T; cfg.G = G; cfg.K = K;
    cfg.world_size = ngpu; cfg.world_rank = rank; cfg.vranks = sh->vranks; cfg.sync_rate = o.sync_rate;
    cfg.shuffle = o.shuffle ? 1 : 0; cfg.seed = o.seed;
    gmrm_engine* e = nullptr;
    ck(gmrm_create(&cfg, &e), "gmrm_create");
    if (ngpu > 1) {
        if (rank == 0) ck(gmrm_comm_unique_id(sh->nccl_id), "gmrm_comm_unique_id");
        bar->wait();
        ck(gmrm_comm_init(e, sh->nccl_id), "gmrm_comm_init");
    }
    int32_t S = 0, M = 0;
    ck(gmrm_shard_info(e, &S, &M, nullptr, nullptr, nullptr), "gmrm_shard_info");

    // ---- genotypes (Bayes::load_genotype, bayes.cpp:867-900): this rank's block, in chunks
    const double t_load = now();
    load_block(e, o, N, S, M);
    if (ngpu > 1 && o.sync_rate == 1) {      // list exchange: the shards read each other's columns over NVLink
        ck(gmrm_comm_local_buffers(e, sh->peer_ptrs[rank]), "gmrm_comm_local_buffers");
        bar->wait();
        for (int r = 0; r < ngpu; r++)
            if (r != rank) ck(gmrm_comm_set_peer_buffers(e, r, r, sh->peer_ptrs[r]), "gmrm_comm_set_peer_buffers");
    }
    if (rank == 0) printf("INFO   : time to load genotype data = %.3f seconds.\n", now() - t_load);
    for (int t = 0; t < T; t++) ck(gmrm_set_phenotype(e, t, sh->phens[t].eps.data(), sh->phens[t].mask4.data(), sh->phens[t].nonas), "gmrm_set_phenotype");
    ck(gmrm_set_groups(e, sh->group_index.data(), o.cva.data()), "gmrm_set_groups");
    const double t_stats = now();
    ck(gmrm_compute_marker_stats(e), "gmrm_compute_marker_stats");
    if (rank == 0) printf("INFO   : Time to compute the markers' statistics: %.2f seconds.\n", now() - t_stats);
    const bool replaying = !sh->replay.its.empty();
    ck(gmrm_init_chain(e, replaying ? sh->replay.sigmag_init.data() : nullptr), "gmrm_init_chain");

    // ---- output files (bayes.cpp:322-324): rank 0 deletes and creates, the others open
    std::vector<host::OutFiles*> outs(T, nullptr);
    if (rank == 0)
        for (int t = 0; t < T; t++) outs[t] = new host::OutFiles(o.out_dir, sh->phens[t].stem, true);
    bar->wait();
    if (rank != 0)
        for (int t = 0; t < T; t++) outs[t] = new host::OutFiles(o.out_dir, sh->phens[t].stem, false);

    std::vector<double> sigmag((size_t)T * G), sigmae(T), pi((size_t)T * G * K), mu(T), betas(M);
    std::vector<int32_t> m0((size_t)T * G), cass((size_t)T * G * K), comp(M);
    // posterior means of beta (<stem>.mbet) only when --burn-in asks for them: no read-back is staged for them otherwise
    const bool want_mean = o.burn_in > 0 && o.burn_in < o.iterations;
    std::vector<std::vector<double>> bmean(T, std::vector<double>(want_mean ? M : 0, 0.0));
    gmrm_state st{sigmag.data(), sigmae.data(), pi.data(), mu.data(), m0.data(), cass.data()};
    // betas / components leave the GPU through the staged path (gmrm_stage_outputs): their copy to pinned host memory
    // runs under the next iteration, and the files of iteration i are written after iteration i+1 has been enqueued
    struct PendingOut { unsigned it = 0; bool save = false, mean = false; std::vector<double> sigmag, pi; std::vector<double> sigmae; std::vector<int32_t> m0; } po;
    auto write_pending = [&]() {
        if (!po.it) return;
        for (int t = 0; t < T; t++) {
            ck(gmrm_fetch_outputs(e, t, betas.data(), comp.data()), "gmrm_fetch_outputs");
            if (po.mean) for (int j = 0; j < M; j++) bmean[t][j] += betas[j];
            if (!po.save) continue;
            const unsigned nthinned = po.it / o.thin - 1;                                          // bayes.cpp:660
            if (rank == 0) {
                int m0_sum = 0;
                for (int g = 0; g < G; g++) m0_sum += po.m0[(size_t)t * G + g];
                outs[t]->write_csv(po.it, nthinned, &po.sigmag[(size_t)t * G], G, po.sigmae[t], m0_sum, &po.pi[(size_t)t * G * K], K);
            }
            outs[t]->write_bet((unsigned)Mt, po.it, nthinned, S, M, betas.data(), rank == 0);
            outs[t]->write_cpn((unsigned)Mt, po.it, nthinned, S, M, comp.data(), rank == 0);
        }
        po.it = 0;
    };
    for (unsigned it = 1; it <= o.iterations; it++) {
        const double ts = now();
        gmrm_replay rp{};
        if (replaying) {
            if (it > sh->replay.its.size()) { printf("FATAL  : the replay file holds %d iterations, iteration %u asked for\n", (int)sh->replay.its.size(), it); exit(EXIT_FAILURE); }
            const host::Replay::It& ri = sh->replay.its[it - 1];
            rp.perm = ri.perm.data(); rp.u = ri.u.data(); rp.z = ri.z.data(); rp.mu_draw = ri.mu_draw.data();
            rp.sigg_unit = ri.sigg_unit.data(); rp.pi_unit = ri.pi_unit.data(); rp.sige_unit = ri.sige_unit.data();
        }
        ck(gmrm_run_iteration_async(e, (int32_t)it, replaying ? &rp : nullptr), "gmrm_run_iteration_async");
        write_pending();                                     // iteration it-1: its files are written while iteration it runs
        ck(gmrm_wait_iteration(e), "gmrm_wait_iteration");
        ck(gmrm_get_state(e, &st), "gmrm_get_state");
        gmrm_timing tm{};
        gmrm_get_timing(e, &tm);
        if (rank % 10 == 0)
            for (int t = 0; t < T; t++) {
                double sg = 0.0;
                for (int g = 0; g < G; g++) sg += sigmag[(size_t)t * G + g];
                printf("RESULT : i:%d r:%d p:%d  sum sigmaG = %20.15f  sigmaE = %20.15f\n", it, rank, t, sg, sigmae[t]);   // bayes.cpp:641
            }
        if (rank == 0) printf("RESULT : It %d  total proc time = %7.3f sec, with sync time = %7.3f\n", it, now() - ts, tm.exchange_ms * 1e-3);   // bayes.cpp:655
        const bool save = it % o.thin == 0, mean = want_mean && it > o.burn_in;
        if (save || mean) {
            ck(gmrm_stage_outputs(e), "gmrm_stage_outputs");
            po.it = it; po.save = save; po.mean = mean;
            po.sigmag = sigmag; po.sigmae = sigmae; po.pi = pi; po.m0 = m0;
        }
    }
    write_pending();
    // ---- superset: posterior means of beta over the iterations after --burn-in, <stem>.mbet (Mt doubles)
    if (want_mean)
        for (int t = 0; t < T; t++) {
            const double inv = 1.0 / (double)(o.iterations - o.burn_in);
            for (auto& v : bmean[t]) v *= inv;
            const std::string path = (o.out_dir.empty() ? std::string() : o.out_dir + "/") + sh->phens[t].stem + ".mbet";
            if (rank == 0) { FILE* f = fopen(path.c_str(), "wb"); if (f) fclose(f); }
            bar->wait();
            FILE* f = fopen(path.c_str(), "r+b");
            if (f) { fseek(f, (long)S * 8, SEEK_SET); fwrite(bmean[t].data(), 8, M, f); fclose(f); }
        }
    bar->wait();
    for (auto* p : outs) delete p;
    gmrm_destroy(e);
}

}  // namespace

int main(int argc, char** argv) {
    Shared sh;
    sh.opt = host::parse_options(argc, argv, 0);
    const host::Options& o = sh.opt;
    sh.dims = host::read_dim_file(o.dim_file, o.truncm);
    printf("INFO   : N = %d individuals, M = %d markers, %d trait(s), %d group(s) x %d mixtures\n", sh.dims.N, sh.dims.Mt, (int)o.phen_files.size(),
           o.ngroups, o.nmixtures);
    for (const auto& f : o.phen_files) sh.phens.push_back(host::read_phen_file(f, sh.dims.N, o.verbosity));
    std::cout.flush();
    if (o.predict) {
        // the reference's ranks are the virtual ranks (include/gmrm_b200.h, gmrm_predict): default one per GPU
        int vr = o.vranks > 0 ? o.vranks : o.gpus;
        if (vr > sh.dims.Mt) vr = sh.dims.Mt;
        vr -= vr % o.gpus;
        if (vr < o.gpus) {
            printf("FATAL  : %d markers cannot be shared by %d GPUs\n", sh.dims.Mt, o.gpus);
            return EXIT_FAILURE;
        }
        sh.vranks = vr;
        sh.bim = host::cross_bim_files(o.bim_file, o.ref_bim_file, true);
        if ((int)sh.bim.ids.size() < sh.dims.Mt) {
            printf("FATAL  : %s lists %d ids for %d markers\n", o.bim_file.c_str(), (int)sh.bim.ids.size(), sh.dims.Mt);
            return EXIT_FAILURE;
        }
        sh.keep.resize(sh.dims.Mt);
        for (int j = 0; j < sh.dims.Mt; j++) sh.keep[j] = sh.bim.ref_index.count(sh.bim.ids[j]) ? 1 : 0;
        for (size_t t = 0; t < sh.phens.size(); t++) {
            unsigned niter = 0;
            const std::string bet = (o.out_dir.empty() ? std::string() : o.out_dir + "/") + sh.phens[t].stem + ".bet";
            sh.bet_mean.push_back(host::read_bet_mean(bet, sh.bim.ref_index.size(), &niter));
            printf("INFO   : Number of recorded iterations in .bet file %d: %u\n", (int)t, niter);                   // bayes.cpp:52-53
            if ((int)sh.bet_mean.back().size() < sh.dims.Mt) {
                printf("FATAL  : %s holds %d markers, the run has %d\n", bet.c_str(), (int)sh.bet_mean.back().size(), sh.dims.Mt);
                return EXIT_FAILURE;
            }
        }
        printf("INFO   : %d GPU(s), %d marker block(s) (ranks of the reference)\n", o.gpus, vr);
        if (o.check_inputs) {
            if (o.selftest_predict) {                    // what the readers found, and one formatted line (tests)
                int kept = 0;
                for (auto k : sh.keep) kept += k;
                printf("SELFTEST: kept %d of %d markers\n", kept, sh.dims.Mt);
                for (size_t t = 0; t < sh.bet_mean.size(); t++) {
                    double s1 = 0.0;
                    for (int j = 0; j < sh.dims.Mt; j++) s1 += sh.bet_mean[t][j] * (j + 1);
                    printf("SELFTEST: trait %d weighted beta mean %.17g\n", (int)t, s1);
                }
                const int j = sh.dims.Mt - 1;
                if (sh.keep[j]) printf("SELFTEST: %s", host::mlma_line(sh.bim.ids[j], j, sh.bim.ref_index.at(sh.bim.ids[j]), 0.25, -1.5, 0.125, 0.0625).c_str());
            }
            printf("INFO   : inputs parsed; --check-inputs given, no GPU work\n");
            return 0;
        }
        fflush(stdout);
        Barrier pbar(o.gpus);
        std::vector<std::thread> pth;
        for (int r = 0; r < o.gpus; r++) pth.emplace_back(predict_worker, r, &sh, &pbar);
        for (auto& t : pth) t.join();
        return 0;
    }
    if (o.group_index_file.empty()) {
        printf("FATAL  : --group-index-file and --group-mixture-file are required\n");
        return EXIT_FAILURE;
    }
    printf("INFO   : Reading groups from %s.\n", o.group_index_file.c_str());
    sh.group_index = host::read_group_index_file(o.group_index_file, o.ngroups, sh.dims.Mt);
    // virtual ranks: the run is the reference under `mpirun -n vranks`
    // default: 2,048 markers in flight per GPU (the benchmarked setting), capped at Mt / 64 in total (host.hpp)
    int vr = o.vranks > 0 ? o.vranks : host::default_vranks(sh.dims.Mt, o.gpus);
    if (!o.replay_file.empty()) {                        // a replayed run IS the reference under `mpirun -n R`: R comes from the log
        sh.replay = host::read_replay_file(o.replay_file);
        const host::Replay& r = sh.replay;
        if (r.T != (int)sh.phens.size() || r.G != o.ngroups || r.K != o.nmixtures || r.Mm != (sh.dims.Mt + r.R - 1) / r.R) {
            printf("FATAL  : replay file %s was logged for T=%d G=%d K=%d R=%d Mm=%d, the run has T=%d G=%d K=%d Mt=%d\n", o.replay_file.c_str(),
                   r.T, r.G, r.K, r.R, r.Mm, (int)sh.phens.size(), o.ngroups, o.nmixtures, sh.dims.Mt);
            return EXIT_FAILURE;
        }
        if (o.vranks > 0 && o.vranks != r.R) printf("WARNING: --vranks %d ignored, the replay file was logged with %d ranks\n", o.vranks, r.R);
        vr = r.R;
        printf("INFO   : replaying the logged variates of %d iterations of a %d-rank reference run\n", r.iterations, r.R);
    }
    if (vr > sh.dims.Mt) vr = sh.dims.Mt;
    if (vr % o.gpus) {
        if (!o.replay_file.empty()) { printf("FATAL  : %d logged ranks cannot be shared evenly by %d GPUs\n", vr, o.gpus); return EXIT_FAILURE; }
        vr -= vr % o.gpus;
    }
    if (vr < o.gpus) {
        printf("FATAL  : %d markers cannot be shared by %d GPUs\n", sh.dims.Mt, o.gpus);
        return EXIT_FAILURE;
    }
    sh.vranks = vr;
    if ((long long)vr * 64 > sh.dims.Mt)                  // few markers per virtual rank: most of the sweep samples against stale residuals
        printf("WARNING: %d virtual ranks for %d markers (%d markers each): every step samples %.1f %% of the markers against the same residuals, "
               "like the reference under `mpirun -n %d`; pass a smaller --vranks (default: at most Mt/64) unless that is intended.\n",
               vr, sh.dims.Mt, sh.dims.Mt / vr, 100.0 * vr / sh.dims.Mt, vr);
    printf("INFO   : %d GPU(s), %d virtual ranks (markers in flight per step), sync rate %d\n", o.gpus, vr, o.sync_rate);
    if (o.check_inputs) {
        for (const auto& p : sh.phens) printf("INFO   : %s: %d observed, %d NA\n", p.path.c_str(), p.nonas, p.nas);
        if (!o.dump_inputs.empty()) {                    // raw dumps for the CPU parity tests of the readers
            auto dump = [&](const std::string& name, const void* p, size_t n) {
                FILE* f = fopen((o.dump_inputs + "/" + name).c_str(), "wb");
                if (f) { fwrite(p, 1, n, f); fclose(f); }
            };
            for (size_t t = 0; t < sh.phens.size(); t++) {
                dump("eps" + std::to_string(t) + ".f64", sh.phens[t].eps.data(), sh.phens[t].eps.size() * 8);
                dump("mask" + std::to_string(t) + ".u8", sh.phens[t].mask4.data(), sh.phens[t].mask4.size());
            }
            dump("groups.i32", sh.group_index.data(), sh.group_index.size() * 4);
            dump("cva.f64", o.cva.data(), o.cva.size() * 8);
        }
        if (o.selftest_outputs) {                        // the writers, exercised exactly as the iteration loop does from 2 ranks
            const int G = o.ngroups, K = o.nmixtures, Mt = sh.dims.Mt, M0 = Mt / 2;
            host::OutFiles r0(o.out_dir, "selftest", true), r1(o.out_dir, "selftest", false);
            for (unsigned it = 1; it <= 4; it++) {
                if (it % 2) continue;                    // thin rate 2: iterations 2 and 4
                const unsigned nthinned = it / 2 - 1;
                std::vector<double> sg(G), pi((size_t)G * K), b(Mt);
                std::vector<int32_t> c(Mt);
                for (int g = 0; g < G; g++) sg[g] = 0.1 * (g + 1) + 0.001 * it;
                for (int i = 0; i < G * K; i++) pi[i] = (i + 1.0) / (G * K * 10.0) + 1e-4 * it;
                for (int j = 0; j < Mt; j++) { b[j] = 1e-3 * j - 0.5 * it; c[j] = (j + it) % K; }
                r0.write_csv(it, nthinned, sg.data(), G, 0.5 + 0.01 * it, 1234 + (int)it, pi.data(), K);
                r0.write_bet((unsigned)Mt, it, nthinned, 0, M0, b.data(), true);
                r1.write_bet((unsigned)Mt, it, nthinned, M0, Mt - M0, b.data() + M0, false);
                r0.write_cpn((unsigned)Mt, it, nthinned, 0, M0, c.data(), true);
                r1.write_cpn((unsigned)Mt, it, nthinned, M0, Mt - M0, c.data() + M0, false);
            }
        }
        printf("INFO   : inputs parsed; --check-inputs given, no GPU work\n");
        return 0;
    }
    fflush(stdout);
    Barrier bar(o.gpus);
    std::vector<std::thread> th;
    for (int r = 0; r < o.gpus; r++) th.emplace_back(worker, r, &sh, &bar);
    for (auto& t : th) t.join();
    return 0;
}

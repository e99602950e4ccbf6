// TEST INFRASTRUCTURE ONLY (oracle/): implementation of the thread-rank MPI stand-in
// declared in oracle/ref_shim/mpi.h, plus the process entry point that runs the
// reference's own main() (compiled with -Dmain=gmrm_main) once per rank-thread.
//
//   GMRM_SHIM_NRANKS   number of ranks (threads) to run, default 1
//   GMRM_RNG_LOG_DIR   if set, every random variate the reference draws is appended to
//                      $GMRM_RNG_LOG_DIR/rank<r>.bin  (record format in shim_hooks.h)
//
// Collectives are two-phase: publish pointers, barrier, copy, barrier.  Reductions
// are summed in rank order so a run is reproducible.
#include "mpi.h"
#include "shim_hooks.h"

#include <barrier>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <memory>
#include <mutex>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

int gmrm_main(int argc, char* argv[]);

namespace {

int g_nranks = 1;
thread_local int t_rank = 0;
std::unique_ptr<std::barrier<>> g_bar;
std::vector<const void*> g_send;
std::vector<int> g_count;
std::vector<FILE*> g_log;

size_t type_size(MPI_Datatype t) {
    switch (t) {
    case MPI_CHAR: case MPI_UNSIGNED_CHAR: case MPI_C_BOOL: return 1;
    case MPI_INT: case MPI_INTEGER: case MPI_UNSIGNED: return 4;
    case MPI_DOUBLE: case MPI_UNSIGNED_LONG_LONG: return 8;
    }
    fprintf(stderr, "mpi_shim: unknown datatype %d\n", t);
    abort();
}

void rank_sync() { g_bar->arrive_and_wait(); }

template <typename T>
void reduce_into(T* out, int count, MPI_Op op) {
    for (int i = 0; i < count; i++) {
        T acc = static_cast<const T*>(g_send[0])[i];
        for (int r = 1; r < g_nranks; r++) {
            T v = static_cast<const T*>(g_send[r])[i];
            if (op == MPI_SUM) acc += v;
            else acc = v > acc ? v : acc;
        }
        out[i] = acc;
    }
}

}  // namespace

extern "C" int gmrm_shim_rank() { return t_rank; }

extern "C" void gmrm_shim_log(const void* bytes, size_t nbytes) {
    FILE* f = g_log.empty() ? nullptr : g_log[t_rank];
    if (f) fwrite(bytes, 1, nbytes, f);
}

int MPI_Init(int*, char***) { return MPI_SUCCESS; }
int MPI_Finalize() { rank_sync(); return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int* r) { *r = t_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int* n) { *n = g_nranks; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm) { rank_sync(); return MPI_SUCCESS; }
double MPI_Wtime() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}
int MPI_Abort(MPI_Comm, int code) { fflush(stdout); _exit(code); }
int MPI_Type_size(MPI_Datatype t, int* s) { *s = (int)type_size(t); return MPI_SUCCESS; }

int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm) {
    g_send[t_rank] = buf;
    rank_sync();
    if (t_rank != root) memcpy(buf, g_send[root], (size_t)count * type_size(t));
    rank_sync();
    return MPI_SUCCESS;
}

int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op op, MPI_Comm) {
    g_send[t_rank] = send;
    rank_sync();
    switch (t) {
    case MPI_DOUBLE: reduce_into(static_cast<double*>(recv), count, op); break;
    case MPI_INT: case MPI_INTEGER: reduce_into(static_cast<int*>(recv), count, op); break;
    case MPI_UNSIGNED: reduce_into(static_cast<unsigned*>(recv), count, op); break;
    case MPI_UNSIGNED_LONG_LONG: reduce_into(static_cast<unsigned long long*>(recv), count, op); break;
    default: fprintf(stderr, "mpi_shim: Allreduce type %d unsupported\n", t); abort();
    }
    rank_sync();
    return MPI_SUCCESS;
}

int MPI_Allgather(const void* send, int scount, MPI_Datatype st, void* recv, int rcount, MPI_Datatype, MPI_Comm) {
    g_send[t_rank] = send;
    rank_sync();
    const size_t nb = (size_t)scount * type_size(st);
    for (int r = 0; r < g_nranks; r++)
        memcpy(static_cast<char*>(recv) + (size_t)r * rcount * type_size(st), g_send[r], nb);
    rank_sync();
    return MPI_SUCCESS;
}

int MPI_Allgatherv(const void* send, int scount, MPI_Datatype st, void* recv, const int* rcounts,
                   const int* displs, MPI_Datatype, MPI_Comm) {
    g_send[t_rank] = send;
    g_count[t_rank] = scount;
    rank_sync();
    const size_t ts = type_size(st);
    for (int r = 0; r < g_nranks; r++) {
        if (g_count[r] != rcounts[r]) { fprintf(stderr, "mpi_shim: Allgatherv count mismatch\n"); abort(); }
        if (rcounts[r] > 0)
            memcpy(static_cast<char*>(recv) + (size_t)displs[r] * ts, g_send[r], (size_t)rcounts[r] * ts);
    }
    rank_sync();
    return MPI_SUCCESS;
}

// ---- MPI-IO on POSIX fds.  O_EXCL is dropped: all rank-threads open the same path.
int MPI_File_open(MPI_Comm, const char* path, int amode, MPI_Info, MPI_File* fh) {
    int flags = 0;
    if (amode & MPI_MODE_RDONLY) flags |= O_RDONLY;
    if (amode & MPI_MODE_WRONLY) flags |= O_WRONLY;
    if (amode & MPI_MODE_CREATE) flags |= O_CREAT;
    int fd = open(path, flags, 0644);
    if (fd < 0) { perror(path); return 1; }
    *fh = fd;
    rank_sync();
    return MPI_SUCCESS;
}
int MPI_File_close(MPI_File* fh) { rank_sync(); close(*fh); return MPI_SUCCESS; }
int MPI_File_delete(const char* path, MPI_Info) {
    rank_sync();
    if (t_rank == 0) unlink(path);
    rank_sync();
    return MPI_SUCCESS;
}
int MPI_File_get_size(MPI_File fh, MPI_Offset* size) {
    struct stat st;
    if (fstat(fh, &st) != 0) return 1;
    *size = st.st_size;
    return MPI_SUCCESS;
}
static int rw_at(MPI_File fh, MPI_Offset off, void* buf, int count, MPI_Datatype t, bool wr) {
    size_t nb = (size_t)count * type_size(t), done = 0;
    while (done < nb) {
        ssize_t n = wr ? pwrite(fh, static_cast<char*>(buf) + done, nb - done, off + done)
                       : pread(fh, static_cast<char*>(buf) + done, nb - done, off + done);
        if (n <= 0) { if (!wr && n == 0) break; perror("mpi_shim io"); return 1; }
        done += (size_t)n;
    }
    return MPI_SUCCESS;
}
int MPI_File_read_at(MPI_File fh, MPI_Offset off, void* buf, int c, MPI_Datatype t, MPI_Status*) { return rw_at(fh, off, buf, c, t, false); }
int MPI_File_read_at_all(MPI_File fh, MPI_Offset off, void* buf, int c, MPI_Datatype t, MPI_Status*) { return rw_at(fh, off, buf, c, t, false); }
int MPI_File_write_at(MPI_File fh, MPI_Offset off, const void* buf, int c, MPI_Datatype t, MPI_Status*) { return rw_at(fh, off, const_cast<void*>(buf), c, t, true); }
int MPI_File_write_at_all(MPI_File fh, MPI_Offset off, const void* buf, int c, MPI_Datatype t, MPI_Status*) { return rw_at(fh, off, const_cast<void*>(buf), c, t, true); }

int main(int argc, char* argv[]) {
    if (const char* e = getenv("GMRM_SHIM_NRANKS")) g_nranks = atoi(e);
    if (g_nranks < 1) g_nranks = 1;
    g_bar = std::make_unique<std::barrier<>>(g_nranks);
    g_send.assign(g_nranks, nullptr);
    g_count.assign(g_nranks, 0);
    if (const char* d = getenv("GMRM_RNG_LOG_DIR")) {
        g_log.assign(g_nranks, nullptr);
        for (int r = 0; r < g_nranks; r++) {
            std::string p = std::string(d) + "/rank" + std::to_string(r) + ".bin";
            g_log[r] = fopen(p.c_str(), "wb");
            if (!g_log[r]) { perror(p.c_str()); return 2; }
        }
    }
    std::vector<std::thread> th;
    for (int r = 0; r < g_nranks; r++)
        th.emplace_back([r, argc, argv] { t_rank = r; gmrm_main(argc, argv); });
    for (auto& t : th) t.join();
    for (FILE* f : g_log) if (f) fclose(f);
    return 0;
}

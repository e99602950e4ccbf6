// C ABI of gmrm_b200 (include/gmrm_b200.h): engine state, HBM buffers, per-iteration driver.
// One engine == one GPU == one contiguous shard of markers.  No CPU fallback anywhere: every
// entry point that computes does so by launching the kernels in kernels.cu.
#include "../../include/gmrm_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "kernels.cuh"

using namespace gmrm;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? GMRM_ENOMEM : GMRM_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                        cudaGetErrorString(e_));                                                   \
    } while (0)

// ---- NCCL, bound at run time so that a single-GPU run needs no NCCL at all -----------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct Nccl {
    void* h = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load() {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names)
            if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!h) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
        Broadcast = (decltype(Broadcast))dlsym(h, "ncclBroadcast");
        AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && CommDestroy && AllReduce && Broadcast && AllGather;
    }
} g_nccl;
constexpr int kNcclInt32 = 2, kNcclFloat64 = 8, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values (nccl.h)

#define NC(call)                                                                                        \
    do {                                                                                                \
        int r_ = (call);                                                                                \
        if (r_ != 0) return fail(GMRM_ENCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call,                  \
                                 g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error");     \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int alloc(size_t count) {
        free();
        n = count;
        if (count == 0) return 0;
        CU(cudaMalloc(&p, count * sizeof(T)));
        return 0;
    }
    int zero(cudaStream_t s) {
        if (n) CU(cudaMemsetAsync(p, 0, n * sizeof(T), s));
        return 0;
    }
    void free() {
        if (p) cudaFree(p);
        p = nullptr; n = 0;
    }
    ~DevBuf() { free(); }
};

void block_of(int Mt, int R, int r, int& S, int& M) {   // Bayes::set_block_of_markers, bayes.cpp:903-925
    const int size = Mt / R, modu = Mt % R;
    M = size + (r < modu ? 1 : 0);
    S = r * size + std::min(r, modu);
}

}  // namespace

struct gmrm_engine {
    gmrm_config cfg{};
    Layout L{};
    cudaStream_t stream = nullptr;
    int Vl = 0, r0 = 0, marker_begin = 0, Mloc = 0, Mm = 0;
    bool bed_final = false, stats_done = false, chain_ready = false, groups_set = false;
    bool buffers_exported = false;   // peers hold pointers / IPC handles of bed, miss_off, miss_idx: they must not be reallocated any more
    std::vector<char> phen_set;
    std::vector<double> h_cva;
    std::vector<int32_t> h_nonas;

    DevBuf<uint8_t> bed, mask4, stage, stage2;     // stage/stage2: double-buffered PLINK staging of gmrm_upload_bed
    // hybrid plan of the step kernel (one trait; GMRM_HYBRID=0 switches it off): second copy of the direct rows, 2-bit fields
    DevBuf<uint8_t> bed2;
    bool hybrid = false;
    int hyb_npass = 0;
    cudaStream_t copy_stream = nullptr;
    // output staging: betas/components of an iteration are snapshotted on the device and copied to
    // pinned host memory on the copy stream while the next iteration runs
    DevBuf<double> out_betas; DevBuf<int32_t> out_comp;
    double* h_out_betas = nullptr; int32_t* h_out_comp = nullptr;
    double* h_out_state = nullptr;   // pinned: [sigmag T*G | sigmae T | pi T*G*K | mu T] doubles, then [m0 T*G | cass T*G*K] int32
    cudaEvent_t ev_staged = nullptr, ev_out = nullptr;
    bool out_pending = false;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    DevBuf<double> eps, mave, msig, betas, cva, cvai, partial, spart, bsq, esq, sigmag, sigmae, pi, mu, mu_old;
    DevBuf<double> delta, delta_tot, gc, bsq_part;
    DevBuf<int32_t> comp, group_loc, mtotgrp, steptab, cass, m0, nonas, err, tmp_cols;
    DevBuf<uint32_t> miss_off, miss_idx, miss_cnt;
    // individuals without a phenotype, per trait (what the marker statistics take out of the whole-column dosage counts)
    std::vector<std::vector<uint32_t>> h_na;
    DevBuf<uint32_t> na_off, na_idx;
    bool na_dirty = true;
    // missing-genotype lists are collected per uploaded chunk (the base-3 bytes do not carry them) and
    // assembled into one CSR by gmrm_finalize_bed
    struct MissChunk { int begin = 0, count = 0; std::vector<uint32_t> cnt; DevBuf<uint32_t>* idx = nullptr; uint64_t total = 0; };
    std::vector<MissChunk> miss_chunks;
    // exchange by published lists (world_size > 1, sync_rate == 1): every GPU applies every GPU's published updates
    // itself, reading the other shards' columns over NVLink peer memory; the lists travel by one small all-gather
    bool list_exchange = false;
    // fused increment exchange (the default at sync_rate 1 on several GPUs; GMRM_EXCHANGE=lists|nccl|delta select the others):
    // every GPU applies its own list, the increments are reduced and the new residuals broadcast inside the step kernel over
    // NVLink peer memory (StepParams::xd_world).  Its receive and landing buffers sit behind this GPU's list block in `plist`,
    // so the six exported buffers stay what they were.
    bool xdelta = false;
    const uint8_t* peer_bed[kMaxGpus] = {};
    const uint32_t* peer_moff[kMaxGpus] = {};
    const uint32_t* peer_midx[kMaxGpus] = {};
    double* peer_plist[kMaxGpus] = {};               // peers' (and own) list buffers [2][world][T][ld]
    double* peer_eps[kMaxGpus] = {};                 // peers' (and own) residual arrays: the row-sharded update stores its rows there
    unsigned long long* peer_rflags[kMaxGpus] = {};  // peers' (and own) row-flag arrays [world][nsm]
    DevBuf<unsigned long long> rflags;
    unsigned long long row_seq = 0;                  // launches with a row-sharded update so far (identical on every GPU)
    bool row_shard = true;                           // GMRM_ROWSHARD=0: every GPU applies every update to every row (round-1 behaviour)
    void* ipc_opened[kMaxGpus][6] = {};
    bool list_p2p = true;            // lists are pushed into the peers' buffers by the sampler kernel (GMRM_EXCHANGE=nccl: all-gather)
    bool merge_pending = false;      // delta exchange: the all-reduced deltas wait to be merged by the next step kernel (fused merge)
    unsigned long long xseq = 0;     // exchange sequence number: one per sampled step, identical on all GPUs
    unsigned long long pend_seq = 0; // sequence number of the step whose lists are pending
    int peers_set = 0;
    DevBuf<double> plist;            // [world or 1][T][publist_doubles(Vl)] compacted published lists; the sampler writes this GPU's block
    int step_tc = 1, step_rpp = 1;   // traits per step launch, rows per pass (step_plan)
    bool force_flush = false;        // GMRM_FORCE_FLUSH=1: update-only launch after every step (timing aid)
    int step_pf = 1;                 // GMRM_STEP_PF=0 turns the L2 prefetch of the streaming loads off
    int pdl = 1;                     // GMRM_PDL=0: plain launches in the marker loop (no programmatic dependent launch)
    int step_warps = 16;             // consumer warps of the step kernel (GMRM_STEP_WARPS=20: measured alternative)
    DevBuf<PubEntry> pub;
    DevBuf<int64_t> npub;
    DevBuf<unsigned long long> prof;   // GMRM_STEP_PROF=1: cycle counters of the step kernel
    // replay staging
    DevBuf<int32_t> rep_perm;
    DevBuf<double> rep_u, rep_z, rep_small;

    ncclComm_t comm = nullptr;

    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> dot_ev;
    // gmrm_run_iteration_async / gmrm_wait_iteration: what the enqueued iteration leaves for the wait
    bool it_pending = false;
    int it_Mm = 0; int64_t it_launches = 0; bool it_multi = false;
    int32_t* h_err = nullptr; int64_t* h_pub = nullptr;     // pinned
    int timing_detail = 0;           // 0: iteration totals only; 1: + the step kernel (2 events per step); 2: every phase (6 events per step)
    gmrm_timing last{};

    ~gmrm_engine() {
        if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
        for (auto& g : ipc_opened) for (void* q : g) if (q) cudaIpcCloseMemHandle(q);
        for (auto& c : miss_chunks) delete c.idx;
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        for (auto& x : ev_h2d) if (x) cudaEventDestroy(x);
        for (auto& x : ev_free) if (x) cudaEventDestroy(x);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (ev_staged) cudaEventDestroy(ev_staged);
        if (ev_out) cudaEventDestroy(ev_out);
        if (h_out_betas) cudaFreeHost(h_out_betas);
        if (h_out_comp) cudaFreeHost(h_out_comp);
        if (h_out_state) cudaFreeHost(h_out_state);
        if (h_err) cudaFreeHost(h_err);
        if (h_pub) cudaFreeHost(h_pub);
        for (auto& e : dot_ev) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};

extern "C" {

const char* gmrm_last_error(void) { return g_err.c_str(); }
const char* gmrm_version(void) { return "gmrm_b200 0.1 (sm_100a)"; }

int gmrm_create(const gmrm_config* c, gmrm_engine** out) {
    if (!c || !out) return fail(GMRM_EINVAL, "null argument");
    *out = nullptr;
    if (c->N < 2 || c->Mt < 1 || c->T < 1 || c->T > 32 || c->G < 1 || c->K < 2 || c->K > kMaxK)
        return fail(GMRM_EINVAL, "bad dimensions N=%d Mt=%d T=%d G=%d K=%d (T<=32, 2<=K<=%d)", c->N, c->Mt, c->T, c->G, c->K, kMaxK);
    if (c->world_size < 1 || c->world_rank < 0 || c->world_rank >= c->world_size)
        return fail(GMRM_EINVAL, "bad world_rank %d / world_size %d", c->world_rank, c->world_size);
    if (c->vranks < c->world_size || c->vranks % c->world_size != 0 || c->vranks > c->Mt)
        return fail(GMRM_EINVAL, "vranks=%d must be a multiple of world_size=%d and <= Mt=%d", c->vranks, c->world_size, c->Mt);
    if (c->sync_rate < 1) return fail(GMRM_EINVAL, "sync_rate must be >= 1");
    if (c->G > 112)   // beta_sq_kernel keeps G x 256 partial sums in shared memory (227 KB per CTA)
        return fail(GMRM_EINVAL, "G=%d marker groups: at most 112 are supported (shared memory of the per-group sum of squares)", c->G);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(GMRM_ENODEVICE, "no CUDA device: gmrm_b200 has no CPU path");
    }
    if (c->device < 0 || c->device >= ndev) return fail(GMRM_EINVAL, "device %d out of range (%d visible)", c->device, ndev);
    CU(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, c->device));
    if (prop.major != 10) return fail(GMRM_ENODEVICE, "device %s is sm_%d%d; this library carries sm_100a code only", prop.name, prop.major, prop.minor);

    auto* e = new gmrm_engine();
    e->cfg = *c;
    const int nsm = c->nsm > 0 ? c->nsm : prop.multiProcessorCount;
    e->L = make_layout(c->N, nsm);
    e->Vl = c->vranks / c->world_size;
    e->r0 = c->world_rank * e->Vl;
    int S, M, Slast, Mlast;
    block_of(c->Mt, c->vranks, e->r0, S, M);
    block_of(c->Mt, c->vranks, e->r0 + e->Vl - 1, Slast, Mlast);
    e->marker_begin = S;
    e->Mloc = Slast + Mlast - S;
    e->Mm = (c->Mt + c->vranks - 1) / c->vranks;
    if (const char* v = getenv("GMRM_STEP_WARPS")) e->step_warps = atoi(v) == 20 ? 20 : 16;
    if (const char* v = getenv("GMRM_STEP_PF")) e->step_pf = atoi(v);
    if (const char* v = getenv("GMRM_PDL")) e->pdl = atoi(v) != 0;
    if (const char* v = getenv("GMRM_FORCE_FLUSH")) e->force_flush = atoi(v) != 0;
    if (getenv("GMRM_STEP_PROF")) {
        if (e->prof.alloc(64) != 0 || e->prof.zero(nullptr) != 0) { delete e; return GMRM_ECUDA; }
    }
    step_plan(e->L, e->Vl, c->T, &e->step_tc, &e->step_rpp);
    if (e->step_tc < 1 || e->step_rpp < 1) {
        const int vl = e->Vl;
        delete e;
        return fail(GMRM_EINVAL, "%d virtual ranks per GPU do not fit the step kernel's shared memory (partials + one table slot)", vl);
    }
    {
        const char* hv = getenv("GMRM_HYBRID");
        e->hybrid = c->T == 1 && GMRM_STEP_DIRECT >= 1 && !(hv && atoi(hv) == 0);
        e->hyb_npass = step_npass(e->L, e->step_rpp);
    }
    e->phen_set.assign(c->T, 0);
    e->h_na.assign(c->T, {});
    e->h_nonas.assign(c->T, 0);

    int rc = 0;
    auto A = [&](int r) { if (rc == 0) rc = r; };
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) { delete e; return fail(GMRM_ECUDA, "stream create failed"); }
    for (auto& evn : e->ev) if (cudaEventCreate(&evn) != cudaSuccess) { delete e; return fail(GMRM_ECUDA, "event create failed"); }
    const Layout& L = e->L;
    const int T = c->T, G = c->G, K = c->K;
    A(e->bed.alloc((size_t)e->Mloc * L.col_stride));
    if (e->hybrid) A(e->bed2.alloc((size_t)e->Mloc * e->hyb_npass * L.nsm * kRowBytes));
    A(e->mask4.alloc((size_t)T * L.col_stride));
    A(e->eps.alloc((size_t)T * L.npad));
    A(e->mave.alloc((size_t)T * e->Mloc)); A(e->msig.alloc((size_t)T * e->Mloc));
    A(e->betas.alloc((size_t)T * e->Mloc)); A(e->comp.alloc((size_t)T * e->Mloc));
    A(e->group_loc.alloc(e->Mloc)); A(e->mtotgrp.alloc(G));
    A(e->cva.alloc((size_t)G * K)); A(e->cvai.alloc((size_t)G * K));
    A(e->steptab.alloc((size_t)e->Mm * e->Vl));
    A(e->partial.alloc((size_t)e->Vl * T * L.nsm));
    A(e->spart.alloc((size_t)T * L.nsm));
    A(e->pub.alloc((size_t)e->Vl * T));
    A(e->cass.alloc((size_t)T * G * K)); A(e->m0.alloc((size_t)T * G));
    A(e->bsq.alloc((size_t)T * G)); A(e->esq.alloc(T)); A(e->bsq_part.alloc((size_t)beta_sq_scratch_doubles(T, G)));
    A(e->sigmag.alloc((size_t)T * G)); A(e->sigmae.alloc(T)); A(e->pi.alloc((size_t)T * G * K));
    A(e->mu.alloc(T)); A(e->mu_old.alloc(T)); A(e->nonas.alloc(T)); A(e->gc.alloc((size_t)T * G * 4 * K));
    A(e->err.alloc(1)); A(e->npub.alloc(1));
    A(e->miss_off.alloc((size_t)e->Mloc + 1)); A(e->miss_idx.alloc(1));
    // Exchange at sync_rate 1 (GMRM_EXCHANGE): "lists" (default) -- every GPU applies every GPU's published updates itself, reading
    // the other shards' columns over NVLink; "delta" -- every GPU applies its own updates and the residual deltas are all-reduced
    // (what sync_rate > 1 always does); "nccl" -- lists, gathered by NCCL instead of pushed by the sampler kernel
    const char* xmode = getenv("GMRM_EXCHANGE");
    const bool x1 = c->world_size > 1 && c->sync_rate == 1;
    e->list_exchange = x1 && xmode && (strcmp(xmode, "lists") == 0 || strcmp(xmode, "nccl") == 0);
    e->xdelta = x1 && !e->list_exchange && !(xmode && strcmp(xmode, "delta") == 0);
    if (c->world_size > kMaxGpus) { delete e; return fail(GMRM_EINVAL, "world_size %d > %d", c->world_size, kMaxGpus); }
    if (c->world_size > 1 && !e->list_exchange && !e->xdelta) { A(e->delta.alloc((size_t)T * L.npad)); A(e->delta_tot.alloc((size_t)T * L.npad)); }
    if (const char* v = getenv("GMRM_EXCHANGE")) e->list_p2p = strcmp(v, "nccl") != 0;
    A(e->plist.alloc((size_t)(e->list_exchange ? 2 * c->world_size : 1) * T * publist_doubles(e->Vl) +   // 2: parity of the exchange sequence
                     (e->xdelta ? 4 + (size_t)2 * (c->world_size + 1) * T * L.npad : 0)));               // receive + landing buffers of the increment exchange
    A(e->rflags.alloc((size_t)kMaxGpus * L.nsm));
    if (const char* v = getenv("GMRM_ROWSHARD")) e->row_shard = atoi(v) != 0;
    if (rc != 0) { delete e; return rc; }
    // everything starts zeroed: genotype tiles (dosage 0), residuals, chain state, missing lists
    for (auto* b : {&e->eps, &e->mave, &e->msig, &e->betas, &e->spart, &e->bsq, &e->esq, &e->sigmag, &e->sigmae, &e->pi, &e->mu,
                    &e->mu_old, &e->partial, &e->delta, &e->delta_tot})
        if (b->zero(e->stream) != 0) { delete e; return GMRM_ECUDA; }
    if (e->plist.zero(e->stream) != 0 || e->rflags.zero(e->stream) != 0) { delete e; return GMRM_ECUDA; }
    if (e->xdelta)   // exchange buffers (behind the list block, see xrecv_offset below): "nothing here yet"
        launch_fill_u64(reinterpret_cast<unsigned long long*>(e->plist.p + (((size_t)T * publist_doubles(e->Vl) + 3) & ~(size_t)3)),
                        (size_t)2 * (c->world_size + 1) * T * L.npad, kXdSentinel, e->stream);
    for (auto* b : {&e->comp, &e->cass, &e->m0, &e->err, &e->steptab, &e->nonas, &e->group_loc, &e->mtotgrp})
        if (b->zero(e->stream) != 0) { delete e; return GMRM_ECUDA; }
    if (e->hybrid && e->bed2.zero(e->stream)) { delete e; return GMRM_ECUDA; }
    if (e->bed.zero(e->stream) || e->mask4.zero(e->stream) || e->miss_off.zero(e->stream) || e->npub.zero(e->stream)) { delete e; return GMRM_ECUDA; }
    if (cudaMemsetAsync(e->pub.p, 0, e->pub.n * sizeof(PubEntry), e->stream) != cudaSuccess || cudaStreamSynchronize(e->stream) != cudaSuccess) {
        delete e;
        return fail(GMRM_ECUDA, "initial memset failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = e;
    return GMRM_OK;
}

void gmrm_destroy(gmrm_engine* e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    cudaStreamSynchronize(e->stream);
    if (e->prof.p) {
        unsigned long long h[64];
        if (cudaMemcpy(h, e->prof.p, sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess && h[7]) {
            fprintf(stderr, "step prof, mean cycles per launch over %llu launches of 2 CTAs:", h[7]);
            for (int i = 8; i < 64; i++) if (h[i]) fprintf(stderr, " [%d] %.0f", i, (double)h[i] / h[7]);
            fprintf(stderr, "\n");
        }
    }
    delete e;
}

int gmrm_shard_info(const gmrm_engine* e, int32_t* marker_begin, int32_t* marker_count, int64_t* column_stride_bytes,
                    int32_t* individuals_per_lane, int32_t* tiles) {
    if (!e) return fail(GMRM_EINVAL, "null engine");
    if (marker_begin) *marker_begin = e->marker_begin;
    if (marker_count) *marker_count = e->Mloc;
    if (column_stride_bytes) *column_stride_bytes = e->L.col_stride;
    if (individuals_per_lane) *individuals_per_lane = 4;   /* one quad per byte */
    if (tiles) *tiles = e->L.nsm;
    return GMRM_OK;
}

// ------------------------------------------------------------------------------------ genotypes
static int ensure_stage(gmrm_engine* e, size_t bytes) {
    bytes = (bytes + 15) & ~(size_t)15;
    if (e->stage.n >= bytes) return 0;
    return e->stage.alloc(bytes);
}

static int chunk_markers(const gmrm_engine* e) {
    size_t budget = 256u << 20;
    if (const char* v = getenv("GMRM_STAGE_MB")) budget = (size_t)std::max(1, atoi(v)) << 20;   // staging chunk size (tests use a small one)
    size_t n = budget / (size_t)e->L.mbytes;
    return (int)std::max<size_t>(1, std::min<size_t>(n, 65535));   // grid.y limit
}

// PLINK bytes of markers [lb, lb+n) (shard-local) sit in e->stage: transcode them into the HBM layout and
// collect their missing-genotype lists.
static int ingest_staged(gmrm_engine* e, const uint8_t* staged, int lb, int n) {
    if (e->miss_cnt.n < (size_t)n + 1) { int rc = e->miss_cnt.alloc((size_t)n + 1); if (rc) return rc; }
    CU(cudaMemsetAsync(e->miss_cnt.p, 0, ((size_t)n + 1) * 4, e->stream));
    launch_transcode(staged, n, e->L, e->bed.p + (size_t)lb * e->L.col_stride, e->miss_cnt.p, e->stream);
    if (e->hybrid) launch_direct_plane(staged, n, e->L, e->hyb_npass, e->bed2.p + (size_t)lb * e->hyb_npass * e->L.nsm * kRowBytes, e->stream);
    CU(cudaGetLastError());
    gmrm_engine::MissChunk ch;
    ch.begin = lb; ch.count = n; ch.cnt.resize(n);
    CU(cudaMemcpyAsync(ch.cnt.data(), e->miss_cnt.p, (size_t)n * 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    std::vector<uint32_t> off((size_t)n + 1, 0);
    uint64_t tot = 0;
    for (int j = 0; j < n; j++) { off[j] = (uint32_t)tot; tot += ch.cnt[j]; }
    if (tot > 0xffffffffull) return fail(GMRM_EINVAL, "more than 2^32 missing genotypes in one chunk");
    off[n] = (uint32_t)tot;
    ch.total = tot;
    // an earlier upload of an overlapping range is superseded
    for (size_t i = 0; i < e->miss_chunks.size();) {
        auto& o = e->miss_chunks[i];
        if (o.begin < lb + n && lb < o.begin + o.count) {
            if (o.begin < lb || o.begin + o.count > lb + n) return fail(GMRM_EINVAL, "re-upload of markers [%d, %d) only partly covers an earlier chunk [%d, %d)", lb, lb + n, o.begin, o.begin + o.count);
            delete o.idx;
            e->miss_chunks.erase(e->miss_chunks.begin() + i);
        } else {
            i++;
        }
    }
    if (tot) {
        ch.idx = new DevBuf<uint32_t>();
        int rc = ch.idx->alloc(tot);
        if (rc) { delete ch.idx; return rc; }
        CU(cudaMemcpyAsync(e->miss_cnt.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice, e->stream));
        launch_fill_missing(staged, n, e->L, e->miss_cnt.p, ch.idx->p, e->stream);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(e->stream));
    }
    e->miss_chunks.push_back(std::move(ch));
    return 0;
}

int gmrm_upload_bed(gmrm_engine* e, const uint8_t* bed, int32_t marker_begin, int32_t marker_count) {
    if (!e || !bed) return fail(GMRM_EINVAL, "null argument");
    if (e->buffers_exported) return fail(GMRM_EINVAL, "genotypes cannot be re-uploaded after the buffers were exported to the other GPUs");
    if (marker_count < 0 || marker_begin < e->marker_begin || marker_begin + marker_count > e->marker_begin + e->Mloc)
        return fail(GMRM_EINVAL, "markers [%d, %d) outside this shard [%d, %d)", marker_begin, marker_begin + marker_count,
                    e->marker_begin, e->marker_begin + e->Mloc);
    CU(cudaSetDevice(e->cfg.device));
    // Two device staging buffers: the H2D copy of chunk k+1 (own stream) overlaps the transcode and the missing-list
    // pass of chunk k.  With a pinned host buffer (gmrm_host_alloc) the copies are true DMA at PCIe speed; a pageable
    // one is staged by the driver and the calls simply serialise.
    const int chunk = chunk_markers(e);
    const size_t stage_bytes = (((size_t)std::min(chunk, std::max(marker_count, 1)) * e->L.mbytes) + 15) & ~(size_t)15;
    int rc = ensure_stage(e, stage_bytes);
    if (rc) return rc;
    if (marker_count > chunk && e->stage2.n < stage_bytes && (rc = e->stage2.alloc(stage_bytes))) return rc;
    if (!e->copy_stream) {
        CU(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) { CU(cudaEventCreateWithFlags(&e->ev_h2d[i], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&e->ev_free[i], cudaEventDisableTiming)); }
    }
    e->bed_final = false;
    uint8_t* stg[2] = {e->stage.p, e->stage2.p};
    const int nchunks = (marker_count + chunk - 1) / chunk;
    auto issue_h2d = [&](int k) -> int {
        const int done = k * chunk, n = std::min(chunk, marker_count - done);
        if (k >= 2) CU(cudaStreamWaitEvent(e->copy_stream, e->ev_free[k & 1], 0));          // the buffer's previous chunk has been ingested
        CU(cudaMemcpyAsync(stg[k & 1], bed + (size_t)done * e->L.mbytes, (size_t)n * e->L.mbytes, cudaMemcpyHostToDevice, e->copy_stream));
        CU(cudaEventRecord(e->ev_h2d[k & 1], e->copy_stream));
        return 0;
    };
    if (nchunks > 0 && (rc = issue_h2d(0))) return rc;
    for (int k = 0; k < nchunks; k++) {
        const int done = k * chunk, n = std::min(chunk, marker_count - done);
        if (k + 1 < nchunks && (rc = issue_h2d(k + 1))) return rc;
        CU(cudaStreamWaitEvent(e->stream, e->ev_h2d[k & 1], 0));
        if ((rc = ingest_staged(e, stg[k & 1], marker_begin - e->marker_begin + done, n))) return rc;   // ends with a stream sync
        CU(cudaEventRecord(e->ev_free[k & 1], e->stream));
    }
    CU(cudaStreamSynchronize(e->copy_stream));
    return GMRM_OK;
}

int gmrm_generate_bed(gmrm_engine* e, uint32_t seed, double maf_lo, double maf_hi, double missing_rate) {
    if (!e) return fail(GMRM_EINVAL, "null engine");
    if (e->buffers_exported) return fail(GMRM_EINVAL, "genotypes cannot be regenerated after the buffers were exported to the other GPUs");
    CU(cudaSetDevice(e->cfg.device));
    const int chunk = chunk_markers(e);
    int rc = ensure_stage(e, (size_t)std::min(chunk, e->Mloc) * e->L.mbytes);
    if (rc) return rc;
    e->bed_final = false;
    for (int done = 0; done < e->Mloc; done += chunk) {
        const int n = std::min(chunk, e->Mloc - done);
        launch_generate_plink(e->stage.p, n, e->marker_begin + done, e->L, seed, maf_lo, maf_hi, missing_rate, e->stream);
        if ((rc = ingest_staged(e, e->stage.p, done, n))) return rc;
    }
    return GMRM_OK;
}

int gmrm_finalize_bed(gmrm_engine* e) {
    if (!e) return fail(GMRM_EINVAL, "null engine");
    if (e->buffers_exported) return fail(GMRM_EINVAL, "the missing-genotype lists cannot be rebuilt after the buffers were exported to the other GPUs");
    CU(cudaSetDevice(e->cfg.device));
    std::vector<uint32_t> cnt(e->Mloc, 0), off((size_t)e->Mloc + 1, 0);
    for (auto& c : e->miss_chunks)
        for (int j = 0; j < c.count; j++) cnt[c.begin + j] = c.cnt[j];
    uint64_t tot = 0;
    for (int j = 0; j < e->Mloc; j++) { off[j] = (uint32_t)tot; tot += cnt[j]; }
    if (tot > 0xffffffffull) return fail(GMRM_EINVAL, "more than 2^32 missing genotypes in one shard");
    off[e->Mloc] = (uint32_t)tot;
    int rc = e->miss_idx.alloc(std::max<uint64_t>(tot, 1));
    if (rc) return rc;
    CU(cudaMemcpyAsync(e->miss_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice, e->stream));
    for (auto& c : e->miss_chunks)
        if (c.total) CU(cudaMemcpyAsync(e->miss_idx.p + off[c.begin], c.idx->p, c.total * 4, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->stage.free(); e->stage2.free();
    e->bed_final = true;
    e->stats_done = false;
    return GMRM_OK;
}

int gmrm_download_bed(gmrm_engine* e, uint8_t* out, int32_t marker_begin, int32_t marker_count) {
    if (!e || !out) return fail(GMRM_EINVAL, "null argument");
    if (marker_count < 0 || marker_begin < e->marker_begin || marker_begin + marker_count > e->marker_begin + e->Mloc)
        return fail(GMRM_EINVAL, "markers outside this shard");
    if (!e->bed_final) { int rcf = gmrm_finalize_bed(e); if (rcf) return rcf; }   // the missing-genotype lists are part of the layout
    CU(cudaSetDevice(e->cfg.device));
    const int chunk = chunk_markers(e);
    DevBuf<uint8_t> tmp;
    int rc = tmp.alloc((((size_t)std::min(chunk, std::max(marker_count, 1)) * e->L.mbytes) + 15) & ~(size_t)15);
    if (rc) return rc;
    for (int done = 0; done < marker_count; done += chunk) {
        const int n = std::min(chunk, marker_count - done);
        const int lb = marker_begin - e->marker_begin + done;
        launch_untranscode(e->bed.p + (size_t)lb * e->L.col_stride, n, e->L, e->miss_off.p + lb, e->miss_idx.p, tmp.p, e->stream);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out + (size_t)done * e->L.mbytes, tmp.p, (size_t)n * e->L.mbytes, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    return GMRM_OK;
}

// ---------------------------------------------------------------------------- phenotypes, groups
int gmrm_set_phenotype(gmrm_engine* e, int32_t t, const double* eps0, const uint8_t* mask4, int32_t nonas) {
    if (!e || !eps0 || !mask4) return fail(GMRM_EINVAL, "null argument");
    if (t < 0 || t >= e->cfg.T) return fail(GMRM_EINVAL, "trait %d out of range", t);
    if (nonas < 2 || nonas > e->cfg.N) return fail(GMRM_EINVAL, "nonas=%d out of range", nonas);
    CU(cudaSetDevice(e->cfg.device));
    const Layout& L = e->L;
    std::vector<double> h((size_t)L.npad, 0.0);
    std::vector<uint8_t> nm((size_t)L.col_stride, 0);
    int seen = 0;
    std::vector<uint32_t> nas;
    for (int i = 0; i < L.N; i++) {
        const bool obs = (mask4[i / 4] >> (i % 4)) & 1;
        h[i] = obs ? eps0[i] : 0.0;
        if (!obs) { nas.push_back((uint32_t)i); continue; }
        seen++;
        nm[i / 4] |= (uint8_t)(1u << (i % 4));
    }
    if (seen != nonas) return fail(GMRM_EINVAL, "mask4 has %d observed individuals but nonas=%d", seen, nonas);
    CU(cudaMemcpyAsync(e->eps.p + (size_t)t * L.npad, h.data(), h.size() * 8, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->mask4.p + (size_t)t * L.col_stride, nm.data(), nm.size(), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->nonas.p + t, &nonas, 4, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->h_nonas[t] = nonas;
    e->h_na[t] = std::move(nas);
    e->na_dirty = true;
    e->phen_set[t] = 1;
    e->stats_done = false;
    e->chain_ready = false;
    return GMRM_OK;
}

int gmrm_set_groups(gmrm_engine* e, const int32_t* group_index, const double* cva) {
    if (!e || !group_index || !cva) return fail(GMRM_EINVAL, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    const int G = e->cfg.G, K = e->cfg.K;
    std::vector<int32_t> mt(G, 0);
    for (int j = 0; j < e->cfg.Mt; j++) {
        if (group_index[j] < 0 || group_index[j] >= G) return fail(GMRM_EINVAL, "marker %d has group %d outside [0, %d)", j, group_index[j], G);
        mt[group_index[j]]++;                                         // bayes.cpp:807-809
    }
    std::vector<double> cvai((size_t)G * K, 0.0);
    for (int g = 0; g < G; g++) {
        if (cva[g * K] != 0.0) return fail(GMRM_EINVAL, "first mixture of group %d must be 0.0", g);     // options.cpp:273-276
        for (int k = 1; k < K; k++) {
            if (cva[g * K + k] <= cva[g * K + k - 1]) return fail(GMRM_EINVAL, "mixtures of group %d not ascending", g);   // 278-281
            cvai[g * K + k] = 1.0 / cva[g * K + k];                    // 282
        }
    }
    e->h_cva.assign(cva, cva + (size_t)G * K);
    CU(cudaMemcpyAsync(e->group_loc.p, group_index + e->marker_begin, (size_t)e->Mloc * 4, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->mtotgrp.p, mt.data(), (size_t)G * 4, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->cva.p, cva, (size_t)G * K * 8, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->cvai.p, cvai.data(), (size_t)G * K * 8, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->groups_set = true;
    e->chain_ready = false;
    return GMRM_OK;
}

// ------------------------------------------------------------------------------ marker statistics
static int upload_na_lists(gmrm_engine* e) {
    if (!e->na_dirty) return 0;
    const int T = e->cfg.T;
    std::vector<uint32_t> off((size_t)T + 1, 0), idx;
    for (int t = 0; t < T; t++) { idx.insert(idx.end(), e->h_na[t].begin(), e->h_na[t].end()); off[t + 1] = (uint32_t)idx.size(); }
    int rc = e->na_off.alloc((size_t)T + 1); if (rc) return rc;
    rc = e->na_idx.alloc(std::max<size_t>(idx.size(), 1)); if (rc) return rc;
    CU(cudaMemcpyAsync(e->na_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice, e->stream));
    if (!idx.empty()) CU(cudaMemcpyAsync(e->na_idx.p, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->na_dirty = false;
    return 0;
}

int gmrm_compute_marker_stats(gmrm_engine* e) {
    if (!e) return fail(GMRM_EINVAL, "null engine");
    if (!e->bed_final) return fail(GMRM_EINVAL, "call gmrm_finalize_bed first");
    for (int t = 0; t < e->cfg.T; t++)
        if (!e->phen_set[t]) return fail(GMRM_EINVAL, "phenotype %d not set", t);
    CU(cudaSetDevice(e->cfg.device));
    { const int rc = upload_na_lists(e); if (rc) return rc; }
    launch_stats(e->bed.p, e->Mloc, e->L, e->mask4.p, e->miss_off.p, e->miss_idx.p, e->nonas.p, e->na_off.p, e->na_idx.p, e->cfg.T, e->mave.p,
                 e->msig.p, e->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    e->stats_done = true;
    return GMRM_OK;
}

int gmrm_get_marker_stats(gmrm_engine* e, int32_t t, double* mave, double* msig) {
    if (!e || t < 0 || t >= e->cfg.T) return fail(GMRM_EINVAL, "bad argument");
    if (!e->stats_done) return fail(GMRM_EINVAL, "marker statistics not computed");
    CU(cudaSetDevice(e->cfg.device));
    if (mave) CU(cudaMemcpy(mave, e->mave.p + (size_t)t * e->Mloc, (size_t)e->Mloc * 8, cudaMemcpyDeviceToHost));
    if (msig) CU(cudaMemcpy(msig, e->msig.p + (size_t)t * e->Mloc, (size_t)e->Mloc * 8, cudaMemcpyDeviceToHost));
    return GMRM_OK;
}

// ------------------------------------------------------------------------------ shared launch glue
// One marker-step (kernels.cu K1) for all traits: pending updates of the previous step (none when pend_step < 0),
// tables, and the dot products of the V columns `cols` (V == 0: update only).  `pend_cols` (non-null) overrides
// the pending list with a single local one of pV entries whose PubEntry block is e->pub (test hook).
struct Pending { bool any = false; bool own_only = false; };
// list buffer layout: [parity of the exchange sequence][GPU][trait][publist_doubles(Vl)] (one block without exchange)
static size_t list_block(const gmrm_engine* e) { return (size_t)e->cfg.T * publist_doubles(e->Vl); }
static double* lists_of(gmrm_engine* e, unsigned long long seq) { return e->plist.p + (e->list_exchange ? (size_t)(seq & 1) * e->cfg.world_size * list_block(e) : 0); }
static double* own_list(gmrm_engine* e, unsigned long long seq) { return lists_of(e, seq) + (e->list_exchange ? (size_t)e->cfg.world_rank * list_block(e) : 0); }
static size_t xrecv_offset(const gmrm_engine* e) { return (list_block(e) + 3) & ~(size_t)3; }   // in doubles; 32-byte aligned
static size_t xland_offset(const gmrm_engine* e) { return xrecv_offset(e) + (size_t)2 * e->cfg.world_size * e->cfg.T * e->L.npad; }

static int launch_step_all(gmrm_engine* e, const int32_t* cols, int V, const Pending& pend, double* partial, int* nlaunch, bool in_loop = false) {
    const int T = e->cfg.T;
    int tc = e->step_tc, rpp = e->step_rpp;
    if (V > e->Vl || V == 0) {                                  // test hook with its own marker count / update-only launch
        step_plan(e->L, V, T, &tc, &rpp);
        if (tc < 1 || rpp < 1) return fail(GMRM_EINVAL, "%d markers per step do not fit the step kernel's shared memory", V);
    }
    for (int t0 = 0; t0 < T; t0 += tc) {
        StepParams p{};
        p.bed = e->bed.p; p.col_stride = e->L.col_stride; p.nrows = e->L.nrows; p.cols = cols; p.V = V;
        p.eps = e->eps.p; p.npad = e->L.npad; p.Ttot = T; p.t0 = t0; p.rows_per_pass = rpp; p.npass = step_npass(e->L, rpp);
        p.partial = partial; p.spart = e->spart.p;
        p.mask4 = e->mask4.p;
        p.pV = e->Vl;
        if (pend.any && e->list_exchange && !pend.own_only) {   // the lists of all GPUs (after the all-gather)
            p.pG = e->cfg.world_size; p.plist = lists_of(e, e->pend_seq);
            for (int g = 0; g < p.pG; g++) { p.pbed[g] = e->peer_bed[g]; p.pmiss_off[g] = e->peer_moff[g]; p.pmiss_idx[g] = e->peer_midx[g]; }
            p.wait_seq = e->pend_seq;
            if (e->row_shard) {                                     // update work and column traffic shared by rows (kernels.cu, a')
                p.rs_world = e->cfg.world_size; p.rs_rank = e->cfg.world_rank; p.rflag_mine = e->rflags.p; p.row_seq = ++e->row_seq;
                for (int g = 0; g < p.pG; g++) { p.peps[g] = e->peer_eps[g]; p.rflag_peer[g] = e->peer_rflags[g]; }
            }
        } else if (pend.any) {                                  // this GPU's own list
            p.pG = 1; p.plist = own_list(e, e->pend_seq); p.wait_seq = e->pend_seq;
            p.pbed[0] = e->bed.p; p.pmiss_off[0] = e->miss_off.p; p.pmiss_idx[0] = e->miss_idx.p;
            if (e->xdelta && !pend.own_only) {                  // ... as increments, reduced and broadcast inside the kernel
                const int W = e->cfg.world_size;
                p.xd_world = W; p.xd_rank = e->cfg.world_rank; p.row_seq = ++e->row_seq;
                for (int g = 0; g < W; g++) { p.xrecv[g] = e->peer_plist[g] + xrecv_offset(e); p.xland[g] = e->peer_plist[g] + xland_offset(e); }
            }
        }
        p.delta = (e->cfg.world_size > 1 && !e->list_exchange && !e->xdelta) ? e->delta.p : nullptr;
        p.merge_tot = (e->merge_pending && in_loop) ? e->delta_tot.p : nullptr;   // only launches of the marker loop take the fused merge
        p.err = e->err.p;
        p.prof = e->prof.p;
        p.pf = e->step_pf;
        p.pdl = e->pdl && in_loop;
        if (e->hybrid && V > 0 && T == 1 && rpp == e->step_rpp && p.npass == e->hyb_npass) {   // the plan the second copy was laid out for
            p.bed2 = e->bed2.p; p.drows = e->hyb_npass * e->L.nsm; p.ndir = 1;
        }
        const int rc = launch_step(e->L, std::min(tc, T - t0), p, e->stream);
        if (rc != 0) return fail(GMRM_ECUDA, "step kernel launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError()));
        if (nlaunch) (*nlaunch)++;
    }
    if (in_loop) e->merge_pending = false;                       // every trait chunk merged its traits' rows
    return 0;
}

static SampleParams sample_params(gmrm_engine* e, const int32_t* cols, int V, const double* partial) {
    SampleParams p{};
    p.V = V; p.T = e->cfg.T; p.G = e->cfg.G; p.K = e->cfg.K; p.N = e->cfg.N; p.nsm = e->L.nsm;
    p.seed = e->cfg.seed; p.r0 = e->r0; p.R = e->cfg.vranks; p.marker_begin = e->marker_begin; p.Mloc = e->Mloc;
    p.cols = cols; p.partial = partial; p.spart = e->spart.p; p.miss_off = e->miss_off.p; p.miss_idx = e->miss_idx.p;
    p.eps = e->eps.p; p.npad = e->L.npad; p.mave = e->mave.p; p.msig = e->msig.p; p.betas = e->betas.p; p.comp = e->comp.p;
    p.group = e->group_loc.p; p.sigmag = e->sigmag.p;
    if (e->step_pf) { p.pf_bed = e->bed.p; p.pf_col_stride = e->L.col_stride; }
    p.gc = e->gc.p; p.nonas = e->nonas.p; p.cass = e->cass.p; p.pub = e->pub.p; p.plist = own_list(e, e->xseq); p.seq = e->xseq; p.err = e->err.p; p.npublished = e->npub.p;
    return p;
}

static int check_step_error(gmrm_engine* e) {
    int32_t herr = 0;
    CU(cudaMemcpyAsync(&herr, e->err.p, 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (herr == 10) {
        CU(cudaMemset(e->err.p, 0, 4));
        return fail(GMRM_ECUDA, "step kernel: dynamic shared memory does not start at or below address %u", kTabBase);
    }
    return 0;
}

int gmrm_dot_products(gmrm_engine* e, const int32_t* local_ids, int32_t n, double* out) {
    if (!e || !local_ids || !out || n < 0) return fail(GMRM_EINVAL, "bad argument");
    if (!e->stats_done) return fail(GMRM_EINVAL, "marker statistics not computed");
    for (int i = 0; i < n; i++)
        if (local_ids[i] < 0 || local_ids[i] >= e->Mloc) return fail(GMRM_EINVAL, "local marker %d out of range", local_ids[i]);
    if (n == 0) return GMRM_OK;
    CU(cudaSetDevice(e->cfg.device));
    const int T = e->cfg.T;
    DevBuf<int32_t> cols; DevBuf<double> partial, res;
    const int chunk = std::min(n, 1024);
    int rc = cols.alloc(n); if (rc) return rc;
    rc = partial.alloc((size_t)chunk * T * e->L.nsm); if (rc) return rc;
    rc = res.alloc((size_t)n * T); if (rc) return rc;
    CU(cudaMemcpyAsync(cols.p, local_ids, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
    for (int done = 0; done < n; done += chunk) {
        const int m = std::min(chunk, n - done);
        rc = launch_step_all(e, cols.p + done, m, Pending{}, partial.p, nullptr); if (rc) return rc;
        SampleParams sp = sample_params(e, cols.p + done, m, partial.p);
        launch_finish_dots(sp, res.p + (size_t)done * T, e->stream);
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(out, res.p, (size_t)n * T * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return check_step_error(e);
}

int gmrm_decode_marker(gmrm_engine* e, int32_t local_id, double* a, double* b) {
    if (!e || local_id < 0 || local_id >= e->Mloc || (!a && !b)) return fail(GMRM_EINVAL, "bad argument");
    if (!e->bed_final) { int rcf = gmrm_finalize_bed(e); if (rcf) return rcf; }
    CU(cudaSetDevice(e->cfg.device));
    DevBuf<double> da, db;
    int rc = da.alloc(e->cfg.N); if (rc) return rc;
    rc = db.alloc(e->cfg.N); if (rc) return rc;
    uint32_t off[2];
    CU(cudaMemcpy(off, e->miss_off.p + local_id, 8, cudaMemcpyDeviceToHost));
    launch_decode_column(e->bed.p + (size_t)local_id * e->L.col_stride, e->L, e->miss_idx.p + off[0], off[1] - off[0], da.p, db.p, e->stream);
    CU(cudaGetLastError());
    if (a) CU(cudaMemcpyAsync(a, da.p, (size_t)e->cfg.N * 8, cudaMemcpyDeviceToHost, e->stream));
    if (b) CU(cudaMemcpyAsync(b, db.p, (size_t)e->cfg.N * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return GMRM_OK;
}

int gmrm_decode_namask(gmrm_engine* e, int32_t t, double* na) {
    if (!e || !na || t < 0 || t >= e->cfg.T) return fail(GMRM_EINVAL, "bad argument");
    CU(cudaSetDevice(e->cfg.device));
    DevBuf<double> d;
    int rc = d.alloc(e->cfg.N); if (rc) return rc;
    launch_decode_namask(e->mask4.p + (size_t)t * e->L.col_stride, e->L, d.p, e->stream);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(na, d.p, (size_t)e->cfg.N * 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return GMRM_OK;
}

int gmrm_apply_update(gmrm_engine* e, int32_t trait, int32_t local_id, double dbeta) {
    if (!e || trait < 0 || trait >= e->cfg.T || local_id < 0 || local_id >= e->Mloc) return fail(GMRM_EINVAL, "bad argument");
    if (!e->stats_done) return fail(GMRM_EINVAL, "marker statistics not computed");
    CU(cudaSetDevice(e->cfg.device));
    const int T = e->cfg.T;
    double av, sg;
    CU(cudaMemcpy(&av, e->mave.p + (size_t)trait * e->Mloc + local_id, 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&sg, e->msig.p + (size_t)trait * e->Mloc + local_id, 8, cudaMemcpyDeviceToHost));
    // a one-item published list for `trait`, empty lists for the others (Phenotype::update_epsilon's dbeta[3], phenotype.cpp:328)
    const size_t ld = publist_doubles(e->Vl);
    std::vector<double> lists((size_t)T * ld, 0.0);
    e->pend_seq = e->xseq;
    for (int t = 0; t < T; t++)                                  // every segment header carries the sequence number the kernel waits for
        for (int sg = 0; sg < publist_segments(e->Vl); sg++) {
            const unsigned long long h = seg_header(t == trait && sg == 0 ? 1 : 0, e->pend_seq);
            memcpy(&lists[(size_t)t * ld + (size_t)sg * kSegDoubles], &h, 8);
        }
    PubItem it{dbeta * sg, av, local_id, 0};
    memcpy(&lists[(size_t)trait * ld + 2], &it, sizeof it);
    CU(cudaMemcpyAsync(own_list(e, e->pend_seq), lists.data(), lists.size() * 8, cudaMemcpyHostToDevice, e->stream));
    Pending one; one.any = true; one.own_only = true;
    int rc = launch_step_all(e, nullptr, 0, one, nullptr, nullptr); if (rc) return rc;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    return check_step_error(e);
}

// ------------------------------------------------------------------------------ association pass
// Bayes::predict (src/bayes.cpp:14-284) for one trait on this GPU's shard.  The reference's ranks are the engine's
// virtual ranks: block r of markers (set_block_of_markers) sees y_k = y - (g - g_r), g_r its own genetic values.
int gmrm_predict(gmrm_engine* e, int32_t t, const double* y, const double* beta_mean, const uint8_t* keep, double* g_out,
                 double* beta, double* tdist, double* se, double* pval) {
    if (!e || !y || !beta_mean) return fail(GMRM_EINVAL, "null argument");
    if (t < 0 || t >= e->cfg.T) return fail(GMRM_EINVAL, "trait %d out of range", t);
    if (!e->stats_done) return fail(GMRM_EINVAL, "marker statistics not computed");
    if (e->cfg.world_size > 1 && !e->comm) return fail(GMRM_EINVAL, "world_size > 1 needs gmrm_comm_init");
    CU(cudaSetDevice(e->cfg.device));
    const Layout& L = e->L;
    const int Mloc = e->Mloc, Vl = e->Vl, R = e->cfg.vranks, nsm = L.nsm;
    cudaStream_t s = e->stream;
    const uint8_t* mask_t = e->mask4.p + (size_t)t * L.col_stride;
    const double* mave_t = e->mave.p + (size_t)t * Mloc;
    const double* msig_t = e->msig.p + (size_t)t * Mloc;
    int rc;

    // the phenotype, 0 at NA individuals and in the padding (phenotype.cpp:647-667)
    std::vector<uint8_t> hmask((size_t)L.col_stride);
    CU(cudaMemcpy(hmask.data(), mask_t, hmask.size(), cudaMemcpyDeviceToHost));
    std::vector<double> hy((size_t)L.npad, 0.0);
    for (int i = 0; i < L.N; i++)
        if ((hmask[i >> 2] >> (i & 3)) & 1) hy[i] = y[i];

    int maxlen = 0;
    for (int v = 0; v < Vl; v++) { int S, M; block_of(e->cfg.Mt, R, e->r0 + v, S, M); maxlen = std::max(maxlen, M); }
    const int vchunk = std::max(1, std::min(maxlen, 1024));            // markers per step-kernel launch
    int tc = 0, rpp = 0;
    step_plan(L, vchunk, 1, &tc, &rpp);
    if (tc < 1 || rpp < 1) return fail(GMRM_EINVAL, "%d markers per launch do not fit the step kernel's shared memory", vchunk);

    DevBuf<double> d_y, d_beta, d_xtx, d_scr, part, g, gk, yk, partial, spart, sumsq, outs;
    DevBuf<uint8_t> d_keep;
    DevBuf<int32_t> cols;
    if ((rc = d_y.alloc((size_t)L.npad)) || (rc = d_beta.alloc(std::max(Mloc, 1))) || (rc = d_xtx.alloc(std::max(Mloc, 1))) ||
        (rc = d_scr.alloc((size_t)2 * std::max(Mloc, 1))) || (rc = part.alloc((size_t)gvalue_chunks(maxlen) * L.npad)) ||
        (rc = g.alloc((size_t)L.npad)) || (rc = gk.alloc((size_t)L.npad)) || (rc = yk.alloc((size_t)L.npad)) ||
        (rc = partial.alloc((size_t)vchunk * nsm)) || (rc = spart.alloc((size_t)nsm)) || (rc = sumsq.alloc(1)) ||
        (rc = outs.alloc((size_t)4 * std::max(Mloc, 1))) || (rc = cols.alloc(std::max(Mloc, 1))))
        return rc;
    if (keep && (rc = d_keep.alloc(std::max(Mloc, 1)))) return rc;
    CU(cudaMemcpyAsync(d_y.p, hy.data(), hy.size() * 8, cudaMemcpyHostToDevice, s));
    if (Mloc > 0) {
        CU(cudaMemcpyAsync(d_beta.p, beta_mean, (size_t)Mloc * 8, cudaMemcpyHostToDevice, s));
        if (keep) CU(cudaMemcpyAsync(d_keep.p, keep, (size_t)Mloc, cudaMemcpyHostToDevice, s));
    }
    const uint8_t* keep_d = keep ? d_keep.p : nullptr;
    launch_iota(cols.p, Mloc, s);
    // sum (a b na)^2 per marker from the dosage counts under the trait's NA mask (bayes.cpp:190-195)
    if ((rc = upload_na_lists(e))) return rc;
    launch_stats(e->bed.p, Mloc, L, mask_t, e->miss_off.p, e->miss_idx.p, e->nonas.p + t, e->na_off.p + t, e->na_idx.p, 1, d_scr.p,
                 d_scr.p + std::max(Mloc, 1), s, d_xtx.p);
    CU(cudaGetLastError());

    // pass 1: g = sum over all blocks of their genetic values (bayes.cpp:87-136)
    if ((rc = g.zero(s))) return rc;
    for (int v = 0; v < Vl; v++) {
        int S, M; block_of(e->cfg.Mt, R, e->r0 + v, S, M);
        const int b0 = S - e->marker_begin;
        launch_gvalues(e->bed.p, L, e->miss_off.p, e->miss_idx.p, b0, b0 + M, mave_t, msig_t, d_beta.p, keep_d, mask_t, part.p, gk.p, g.p, s);
        CU(cudaGetLastError());
    }
    if (e->cfg.world_size > 1) NC(g_nccl.AllReduce(g.p, g.p, (size_t)L.npad, kNcclFloat64, kNcclSum, e->comm, s));
    if (g_out) CU(cudaMemcpyAsync(g_out, g.p, (size_t)L.N * 8, cudaMemcpyDeviceToHost, s));

    // pass 2, block by block: y_k, its variance, and the markers' statistics (bayes.cpp:138-214)
    for (int v = 0; v < Vl; v++) {
        int S, M; block_of(e->cfg.Mt, R, e->r0 + v, S, M);
        if (M == 0) continue;
        const int b0 = S - e->marker_begin;
        launch_gvalues(e->bed.p, L, e->miss_off.p, e->miss_idx.p, b0, b0 + M, mave_t, msig_t, d_beta.p, keep_d, mask_t, part.p, gk.p, nullptr, s);
        launch_predict_residual(d_y.p, g.p, gk.p, L, yk.p, s);
        launch_eps_sumsq(yk.p, L.npad, L.N, 1, sumsq.p, s);
        CU(cudaGetLastError());
        for (int done = 0; done < M; done += vchunk) {
            const int V = std::min(vchunk, M - done);
            StepParams p{};
            p.bed = e->bed.p; p.col_stride = L.col_stride; p.nrows = L.nrows; p.cols = cols.p + b0 + done; p.V = V;
            p.eps = yk.p; p.npad = L.npad; p.Ttot = 1; p.t0 = 0; p.rows_per_pass = rpp; p.npass = step_npass(L, rpp);
            p.partial = partial.p; p.spart = spart.p; p.mask4 = mask_t; p.pV = Vl;
            p.err = e->err.p; p.prof = e->prof.p; p.pf = e->step_pf;
            const int lrc = launch_step(L, 1, p, s);
            if (lrc != 0) return fail(GMRM_ECUDA, "step kernel launch failed (%d): %s", lrc, cudaGetErrorString(cudaGetLastError()));
            launch_predict_finish(cols.p + b0 + done, V, nsm, partial.p, d_xtx.p, sumsq.p, e->h_nonas[t], keep_d,
                                  outs.p, outs.p + Mloc, outs.p + 2 * (size_t)Mloc, outs.p + 3 * (size_t)Mloc, s);
            CU(cudaGetLastError());
        }
    }
    if (Mloc > 0) {
        if (beta) CU(cudaMemcpyAsync(beta, outs.p, (size_t)Mloc * 8, cudaMemcpyDeviceToHost, s));
        if (tdist) CU(cudaMemcpyAsync(tdist, outs.p + Mloc, (size_t)Mloc * 8, cudaMemcpyDeviceToHost, s));
        if (se) CU(cudaMemcpyAsync(se, outs.p + 2 * (size_t)Mloc, (size_t)Mloc * 8, cudaMemcpyDeviceToHost, s));
        if (pval) CU(cudaMemcpyAsync(pval, outs.p + 3 * (size_t)Mloc, (size_t)Mloc * 8, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    return check_step_error(e);
}

// Genetic values alone (the first sum of Bayes::predict, src/bayes.cpp:87-136): g = sum over ALL markers of all shards
// of ((a - mave) msig) beta under the trait's NA mask.  One pass over the shard's genotypes + one all-reduce.
int gmrm_genetic_values(gmrm_engine* e, int32_t t, const double* beta, double* g_out) {
    if (!e || !beta || !g_out) return fail(GMRM_EINVAL, "null argument");
    if (t < 0 || t >= e->cfg.T) return fail(GMRM_EINVAL, "trait %d out of range", t);
    if (!e->stats_done) return fail(GMRM_EINVAL, "marker statistics not computed");
    if (e->cfg.world_size > 1 && !e->comm) return fail(GMRM_EINVAL, "world_size > 1 needs gmrm_comm_init");
    CU(cudaSetDevice(e->cfg.device));
    const Layout& L = e->L;
    const int Mloc = e->Mloc;
    cudaStream_t s = e->stream;
    DevBuf<double> d_beta, part, g, gk;
    int rc;
    const int blk = 65536;                                    // markers per launch_gvalues call (<= 64 chunks each)
    if ((rc = d_beta.alloc(std::max(Mloc, 1))) || (rc = part.alloc((size_t)gvalue_chunks(std::min(Mloc, blk)) * L.npad)) ||
        (rc = g.alloc((size_t)L.npad)) || (rc = gk.alloc((size_t)L.npad)))
        return rc;
    if (Mloc > 0) CU(cudaMemcpyAsync(d_beta.p, beta, (size_t)Mloc * 8, cudaMemcpyHostToDevice, s));
    if ((rc = g.zero(s))) return rc;
    for (int b0 = 0; b0 < Mloc; b0 += blk) {
        launch_gvalues(e->bed.p, L, e->miss_off.p, e->miss_idx.p, b0, std::min(Mloc, b0 + blk), e->mave.p + (size_t)t * Mloc,
                       e->msig.p + (size_t)t * Mloc, d_beta.p, nullptr, e->mask4.p + (size_t)t * L.col_stride, part.p, gk.p, g.p, s);
        CU(cudaGetLastError());
    }
    if (e->cfg.world_size > 1) NC(g_nccl.AllReduce(g.p, g.p, (size_t)L.npad, kNcclFloat64, kNcclSum, e->comm, s));
    CU(cudaMemcpyAsync(g_out, g.p, (size_t)L.N * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return GMRM_OK;
}

// ------------------------------------------------------------------------------------- the chain
int gmrm_init_chain(gmrm_engine* e, const double* sigmag_init) {
    if (!e) return fail(GMRM_EINVAL, "null engine");
    if (!e->stats_done) return fail(GMRM_EINVAL, "marker statistics not computed");
    if (!e->groups_set) return fail(GMRM_EINVAL, "groups not set");
    CU(cudaSetDevice(e->cfg.device));
    const int T = e->cfg.T, G = e->cfg.G, K = e->cfg.K;
    std::vector<int32_t> mt(G);
    CU(cudaMemcpy(mt.data(), e->mtotgrp.p, (size_t)G * 4, cudaMemcpyDeviceToHost));
    // sigmaG ~ Beta(1,1), zero for empty groups (bayes.cpp:326-331); Beta(1,1) == U(0,1)
    std::vector<double> sg((size_t)T * G);
    for (int t = 0; t < T; t++)
        for (int g = 0; g < G; g++) {
            double v = sigmag_init ? sigmag_init[t * G + g] : draw_uniform(e->cfg.seed, STREAM_SIGMAG0, 0, (uint32_t)g, (uint32_t)t);
            if (mt[g] == 0) v = 0.0;
            sg[t * G + g] = v;
        }
    // pi_prior (bayes.hpp:34-47)
    std::vector<double> pi((size_t)T * G * K);
    for (int g = 0; g < G; g++) {
        double sum_cva = 0.0;
        for (int j = 0; j < K - 1; j++) sum_cva += e->h_cva[g * K + j + 1];
        for (int t = 0; t < T; t++) {
            double* row = &pi[((size_t)t * G + g) * K];
            row[0] = 0.5;
            for (int j = 1; j < K; j++) row[j] = row[0] * e->h_cva[g * K + j] / sum_cva;
        }
    }
    CU(cudaMemcpyAsync(e->sigmag.p, sg.data(), sg.size() * 8, cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->pi.p, pi.data(), pi.size() * 8, cudaMemcpyHostToDevice, e->stream));
    if (e->betas.zero(e->stream) || e->comp.zero(e->stream) || e->mu.zero(e->stream) || e->mu_old.zero(e->stream) ||
        e->cass.zero(e->stream) || e->m0.zero(e->stream) || e->err.zero(e->stream))
        return GMRM_ECUDA;
    // sigmaE start: sum eps^2 na / nonas / 2 (phenotype.cpp:448-457); eps is 0 at NA and pad slots
    launch_eps_sumsq(e->eps.p, e->L.npad, e->L.npad, T, e->esq.p, e->stream);
    launch_init_sigmae(e->esq.p, e->nonas.p, T, e->sigmae.p, e->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    e->chain_ready = true;
    return GMRM_OK;
}

static int upload_opt(DevBuf<double>& buf, const double* src, size_t n, cudaStream_t s, const double** dptr) {
    *dptr = nullptr;
    if (!src) return 0;
    if (buf.n < n) { int rc = buf.alloc(n); if (rc) return rc; }
    CU(cudaMemcpyAsync(buf.p, src, n * 8, cudaMemcpyHostToDevice, s));
    *dptr = buf.p;
    return 0;
}

int gmrm_run_iteration(gmrm_engine* e, int32_t it, const gmrm_replay* rp) {
    const int rc = gmrm_run_iteration_async(e, it, rp);
    return rc ? rc : gmrm_wait_iteration(e);
}

int gmrm_run_iteration_async(gmrm_engine* e, int32_t it, const gmrm_replay* rp) {
    if (!e) return fail(GMRM_EINVAL, "null engine");
    if (!e->chain_ready) return fail(GMRM_EINVAL, "call gmrm_init_chain first");
    if (e->it_pending) return fail(GMRM_EINVAL, "an iteration is already enqueued: call gmrm_wait_iteration first");
    if (e->cfg.world_size > 1 && !e->comm) return fail(GMRM_EINVAL, "world_size > 1 needs gmrm_comm_init");
    CU(cudaSetDevice(e->cfg.device));
    const gmrm_config& c = e->cfg;
    const Layout& L = e->L;
    const int T = c.T, G = c.G, K = c.K, R = c.vranks, Mm = e->Mm, Vl = e->Vl;
    cudaStream_t s = e->stream;
    int rc;

    // ---- replay variates for this iteration
    const double *d_u = nullptr, *d_z = nullptr;
    const int32_t* d_perm = nullptr;
    const double *d_mu = nullptr, *d_sigg = nullptr, *d_piu = nullptr, *d_sige = nullptr;
    if (rp) {
        if ((rp->u == nullptr) != (rp->z == nullptr)) return fail(GMRM_EINVAL, "replay u and z must be given together");
        const size_t nuz = (size_t)Mm * R * T;
        if ((rc = upload_opt(e->rep_u, rp->u, nuz, s, &d_u))) return rc;
        if ((rc = upload_opt(e->rep_z, rp->z, nuz, s, &d_z))) return rc;
        if (rp->perm) {
            if (e->rep_perm.n < (size_t)R * Mm && (rc = e->rep_perm.alloc((size_t)R * Mm))) return rc;
            CU(cudaMemcpyAsync(e->rep_perm.p, rp->perm, (size_t)R * Mm * 4, cudaMemcpyHostToDevice, s));
            d_perm = e->rep_perm.p;
        }
        const size_t nsmall = (size_t)T + (size_t)T * G + (size_t)T * G * K + T;
        std::vector<double> small(nsmall, 0.0);
        if (e->rep_small.n < nsmall && (rc = e->rep_small.alloc(nsmall))) return rc;
        double* hp = small.data();
        if (rp->mu_draw) { memcpy(hp, rp->mu_draw, T * 8); d_mu = e->rep_small.p; }
        hp += T;
        if (rp->sigg_unit) { memcpy(hp, rp->sigg_unit, (size_t)T * G * 8); d_sigg = e->rep_small.p + T; }
        hp += (size_t)T * G;
        if (rp->pi_unit) { memcpy(hp, rp->pi_unit, (size_t)T * G * K * 8); d_piu = e->rep_small.p + T + (size_t)T * G; }
        hp += (size_t)T * G * K;
        if (rp->sige_unit) { memcpy(hp, rp->sige_unit, T * 8); d_sige = e->rep_small.p + T + (size_t)T * G + (size_t)T * G * K; }
        CU(cudaMemcpyAsync(e->rep_small.p, small.data(), nsmall * 8, cudaMemcpyHostToDevice, s));
        CU(cudaStreamSynchronize(s));   // host staging vectors go out of scope
    }

    if (e->timing_detail && e->dot_ev.size() < (size_t)6 * Mm) {
        const size_t old = e->dot_ev.size();
        e->dot_ev.resize((size_t)6 * Mm);
        for (size_t i = old; i < e->dot_ev.size(); i++) CU(cudaEventCreate(&e->dot_ev[i]));
    }

    if (e->prof.p)
        if (const char* pv = getenv("GMRM_STEP_PROF"))
            if (atoi(pv) == it) CU(cudaMemsetAsync(e->prof.p, 0, 64 * 8, s));   // count from iteration GMRM_STEP_PROF on
    int64_t launches = 0;
    CU(cudaEventRecord(e->ev[0], s));
    // ---- prologue (bayes.cpp:347-368)
    MuDrawParams mp{};
    mp.T = T; mp.it = it; mp.seed = c.seed; mp.sigmae = e->sigmae.p; mp.nonas = e->nonas.p; mp.mu = e->mu.p; mp.mu_old = e->mu_old.p; mp.rep_mu = d_mu;
    launch_mu_draw(mp, s);
    launch_eps_offset(e->eps.p, e->mask4.p, L, T, e->mu_old.p, e->mu.p, s);
    launch_steptab(e->steptab.p, Mm, Vl, e->r0, R, c.Mt, e->marker_begin, c.shuffle, c.seed, it, d_perm, s);
    launch_group_consts(T, G, K, c.N, e->sigmag.p, e->sigmae.p, e->pi.p, e->cva.p, e->cvai.p, e->nonas.p, e->gc.p, s);
    CU(cudaMemsetAsync(e->cass.p, 0, e->cass.n * 4, s));
    CU(cudaMemsetAsync(e->npub.p, 0, 8, s));
    launches += 4;

    // ---- marker loop (bayes.cpp:375-555).  The residual update of step s is fused into the step kernel of
    // step s+1 (it needs the same rows of eps the table build reads); an update-only launch flushes it after the
    // last step.  Exchange (bayes.cpp:495-553):
    //   sync_rate == 1: the GPUs all-gather their published lists (V x 16 B each) after the sampler; the next step
    //                   kernel applies ALL lists in global virtual-rank order, reading the other shards' columns
    //                   over NVLink peer memory -- every GPU keeps the same residuals, bit for bit, as a single
    //                   GPU with R virtual ranks would
    //   sync_rate  > 1: each GPU applies its own list, accumulates what it changed, and every sync_rate steps the
    //                   residual deltas are all-reduced and merged
    CU(cudaEventRecord(e->ev[1], s));
    const bool multi = c.world_size > 1;
    if (e->list_exchange || e->xdelta) {
        if (e->peers_set != c.world_size - 1) return fail(GMRM_EINVAL, "world_size > 1 with sync_rate 1 needs the peers' buffers (gmrm_comm_import_buffers / gmrm_comm_set_peer_buffers)");
    }
    Pending pend;                             // published updates of the previous step still to be applied?
    for (int st = 0; st < Mm; st++) {
        const int32_t* cols = e->steptab.p + (size_t)st * Vl;
        int nl = 0;
        if (e->timing_detail) CU(cudaEventRecord(e->dot_ev[6 * st], s));
        if ((rc = launch_step_all(e, cols, Vl, pend, e->partial.p, &nl, st > 0))) return rc;   // step 0 follows memsets, not a kernel of the loop
        if (e->timing_detail) CU(cudaEventRecord(e->dot_ev[6 * st + 1], s));
        e->xseq++;                                           // this step's lists: sequence number xseq, buffer parity xseq & 1
        SampleParams sp = sample_params(e, cols, Vl, e->partial.p);
        sp.it = it; sp.step = st; sp.rep_u = d_u; sp.rep_z = d_z; sp.pdl = e->pdl;
        if (e->list_exchange && e->list_p2p) {               // the sampler pushes the list into every peer's buffer and raises our flag there
            sp.world = c.world_size; sp.rank = c.world_rank; sp.seq = e->xseq;
            for (int g = 0; g < c.world_size; g++)
                sp.peer_list[g] = e->peer_plist[g] + ((size_t)(e->xseq & 1) * c.world_size + c.world_rank) * list_block(e);
        }
        launch_sample(sp, s);
        if (e->timing_detail > 1) CU(cudaEventRecord(e->dot_ev[6 * st + 2], s));
        pend.any = true;
        e->pend_seq = e->xseq;
        const bool delta_exchange = multi && !e->list_exchange && !e->xdelta && ((st + 1) % c.sync_rate == 0 || st == Mm - 1);
        if (e->list_exchange && !e->list_p2p) {
            NC(g_nccl.AllGather(own_list(e, e->xseq), lists_of(e, e->xseq), (size_t)T * publist_doubles(Vl), kNcclFloat64, e->comm, s));
            if (e->timing_detail > 1) CU(cudaEventRecord(e->dot_ev[6 * st + 5], s));
            launches += 1;
        }
        if (delta_exchange || st == Mm - 1 || e->force_flush || e->timing_detail > 1) {   // detail 2: the update phase as its own launch, timed alone
            if ((rc = launch_step_all(e, nullptr, 0, pend, nullptr, &nl))) return rc;
            pend.any = false;
        }
        if (e->timing_detail > 1) CU(cudaEventRecord(e->dot_ev[6 * st + 3], s));
        launches += nl + 1;
        if (delta_exchange) {
            NC(g_nccl.AllReduce(e->delta.p, e->delta_tot.p, (size_t)T * L.npad, kNcclFloat64, kNcclSum, e->comm, s));
            if (e->timing_detail > 1) CU(cudaEventRecord(e->dot_ev[6 * st + 5], s));
            if (st == Mm - 1) { launch_eps_merge(e->eps.p, e->delta.p, e->delta_tot.p, L, T, s); launches += 1; }   // the epilogue reads eps
            else e->merge_pending = true;                        // merged by the next step kernel, CTA by CTA (no extra launch)
            launches += 1;
        }
        if (e->timing_detail > 1) CU(cudaEventRecord(e->dot_ev[6 * st + 4], s));
    }
    CU(cudaEventRecord(e->ev[2], s));

    // ---- epilogue (bayes.cpp:562-651)
    launch_beta_sq(e->betas.p, e->group_loc.p, e->Mloc, T, G, e->bsq.p, e->bsq_part.p, s);
    if (multi) {   // Allreduce of beta_sqn and cass (bayes.cpp:575-588)
        NC(g_nccl.AllReduce(e->bsq.p, e->bsq.p, (size_t)T * G, kNcclFloat64, kNcclSum, e->comm, s));
        NC(g_nccl.AllReduce(e->cass.p, e->cass.p, (size_t)T * G * K, kNcclInt32, kNcclSum, e->comm, s));
    }
    launch_eps_sumsq(e->eps.p, L.npad, c.N, T, e->esq.p, s);
    GlobalDrawParams gp{};
    gp.T = T; gp.G = G; gp.K = K; gp.N = c.N; gp.it = it; gp.seed = c.seed; gp.mtotgrp = e->mtotgrp.p; gp.bsq = e->bsq.p;
    gp.cass = e->cass.p; gp.esq = e->esq.p; gp.sigmag = e->sigmag.p; gp.sigmae = e->sigmae.p; gp.pi = e->pi.p; gp.m0 = e->m0.p;
    gp.rep_sigg_unit = d_sigg; gp.rep_pi_unit = d_piu; gp.rep_sige_unit = d_sige; gp.err = e->err.p;
    launch_global_draw(gp, s);
    if (multi) {   // Bcast of shard 0's sigmaG, sigmaE, pi (bayes.cpp:626,638,649)
        NC(g_nccl.Broadcast(e->sigmag.p, e->sigmag.p, (size_t)T * G, kNcclFloat64, 0, e->comm, s));
        NC(g_nccl.Broadcast(e->sigmae.p, e->sigmae.p, (size_t)T, kNcclFloat64, 0, e->comm, s));
        NC(g_nccl.Broadcast(e->pi.p, e->pi.p, (size_t)T * G * K, kNcclFloat64, 0, e->comm, s));
    }
    launches += 3;
    CU(cudaEventRecord(e->ev[3], s));
    CU(cudaGetLastError());

    if (!e->h_err) {
        CU(cudaHostAlloc((void**)&e->h_err, 4, cudaHostAllocDefault));
        CU(cudaHostAlloc((void**)&e->h_pub, 8, cudaHostAllocDefault));
    }
    CU(cudaMemcpyAsync(e->h_err, e->err.p, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(e->h_pub, e->npub.p, 8, cudaMemcpyDeviceToHost, s));
    e->it_pending = true; e->it_Mm = Mm; e->it_launches = launches; e->it_multi = multi;
    return GMRM_OK;
}

int gmrm_wait_iteration(gmrm_engine* e) {
    if (!e) return fail(GMRM_EINVAL, "null engine");
    if (!e->it_pending) return fail(GMRM_EINVAL, "no iteration enqueued");
    CU(cudaSetDevice(e->cfg.device));
    const gmrm_config& c = e->cfg;
    cudaStream_t s = e->stream;
    const int Mm = e->it_Mm;
    const int64_t launches = e->it_launches;
    const bool multi = e->it_multi;
    e->it_pending = false;
    CU(cudaStreamSynchronize(s));
    const int32_t herr = *e->h_err;
    const int64_t hpub = *e->h_pub;
    float ms_loop = 0, ms_all = 0;
    CU(cudaEventElapsedTime(&ms_loop, e->ev[1], e->ev[2]));
    CU(cudaEventElapsedTime(&ms_all, e->ev[0], e->ev[3]));
    e->last.marker_loop_ms = ms_loop; e->last.iteration_ms = ms_all; e->last.launches = launches; e->last.steps = Mm; e->last.published = hpub;
    e->last.dot_kernel_ms = 0.0; e->last.sample_kernel_ms = 0.0; e->last.update_kernel_ms = 0.0; e->last.exchange_ms = 0.0; e->last.allreduce_ms = 0.0;
    if (e->timing_detail)
        for (int st = 0; st < Mm; st++) {
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, e->dot_ev[6 * st], e->dot_ev[6 * st + 1]));
            e->last.dot_kernel_ms += ms;
            if (e->timing_detail < 2) continue;
            CU(cudaEventElapsedTime(&ms, e->dot_ev[6 * st + 1], e->dot_ev[6 * st + 2]));
            e->last.sample_kernel_ms += ms;
            CU(cudaEventElapsedTime(&ms, e->dot_ev[6 * st + 2], e->dot_ev[6 * st + 3]));
            e->last.update_kernel_ms += ms;
            CU(cudaEventElapsedTime(&ms, e->dot_ev[6 * st + 3], e->dot_ev[6 * st + 4]));
            e->last.exchange_ms += ms;
            if (e->list_exchange && !e->list_p2p) {
                CU(cudaEventElapsedTime(&ms, e->dot_ev[6 * st + 2], e->dot_ev[6 * st + 5]));
                e->last.allreduce_ms += ms;
            } else if (multi && !e->list_exchange && !e->xdelta && ((st + 1) % c.sync_rate == 0 || st == Mm - 1)) {
                CU(cudaEventElapsedTime(&ms, e->dot_ev[6 * st + 3], e->dot_ev[6 * st + 5]));
                e->last.allreduce_ms += ms;
            }
        }
    if (herr == 10) {
        CU(cudaMemset(e->err.p, 0, 4));
        return fail(GMRM_ECUDA, "step kernel: dynamic shared memory does not start at or below address %u", kTabBase);
    }
    if (herr != 0) {
        CU(cudaMemset(e->err.p, 0, 4));
        return fail(GMRM_EREPLAY, "replay variates exhausted (code %d): the chain asked for a draw the reference did not make", herr);
    }
    return GMRM_OK;
}

int gmrm_get_state(gmrm_engine* e, gmrm_state* o) {
    if (!e || !o) return fail(GMRM_EINVAL, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    const size_t T = e->cfg.T, G = e->cfg.G, K = e->cfg.K;
    if (o->sigmag) CU(cudaMemcpy(o->sigmag, e->sigmag.p, T * G * 8, cudaMemcpyDeviceToHost));
    if (o->sigmae) CU(cudaMemcpy(o->sigmae, e->sigmae.p, T * 8, cudaMemcpyDeviceToHost));
    if (o->pi) CU(cudaMemcpy(o->pi, e->pi.p, T * G * K * 8, cudaMemcpyDeviceToHost));
    if (o->mu) CU(cudaMemcpy(o->mu, e->mu.p, T * 8, cudaMemcpyDeviceToHost));
    if (o->m0) CU(cudaMemcpy(o->m0, e->m0.p, T * G * 4, cudaMemcpyDeviceToHost));
    if (o->cass) CU(cudaMemcpy(o->cass, e->cass.p, T * G * K * 4, cudaMemcpyDeviceToHost));
    return GMRM_OK;
}

int gmrm_get_betas(gmrm_engine* e, int32_t t, double* betas) {
    if (!e || !betas || t < 0 || t >= e->cfg.T) return fail(GMRM_EINVAL, "bad argument");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaMemcpy(betas, e->betas.p + (size_t)t * e->Mloc, (size_t)e->Mloc * 8, cudaMemcpyDeviceToHost));
    return GMRM_OK;
}
int gmrm_get_components(gmrm_engine* e, int32_t t, int32_t* comp) {
    if (!e || !comp || t < 0 || t >= e->cfg.T) return fail(GMRM_EINVAL, "bad argument");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaMemcpy(comp, e->comp.p + (size_t)t * e->Mloc, (size_t)e->Mloc * 4, cudaMemcpyDeviceToHost));
    return GMRM_OK;
}
int gmrm_get_epsilon(gmrm_engine* e, int32_t t, double* eps) {
    if (!e || !eps || t < 0 || t >= e->cfg.T) return fail(GMRM_EINVAL, "bad argument");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaMemcpy(eps, e->eps.p + (size_t)t * e->L.npad, (size_t)e->cfg.N * 8, cudaMemcpyDeviceToHost));
    return GMRM_OK;
}
// Asynchronous read-back of an iteration's betas / components (what .bet / .cpn record, bayes.cpp:666-667):
// gmrm_stage_outputs snapshots them on the device and starts the device-to-host copy on a second stream;
// gmrm_fetch_outputs waits for that copy and hands the values out.  Calling run_iteration in between overlaps the
// copy with the next iteration.
static size_t state_doubles(const gmrm_engine* e) { const size_t T = e->cfg.T, G = e->cfg.G, K = e->cfg.K; return T * G + T + T * G * K + T; }
static size_t state_ints(const gmrm_engine* e) { const size_t T = e->cfg.T, G = e->cfg.G, K = e->cfg.K; return T * G + T * G * K; }
int gmrm_stage_outputs(gmrm_engine* e) {
    if (!e) return fail(GMRM_EINVAL, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    const size_t n = (size_t)e->cfg.T * e->Mloc;
    if (!e->ev_staged) {
        int rc = e->out_betas.alloc(n); if (rc) return rc;
        rc = e->out_comp.alloc(n); if (rc) return rc;
        CU(cudaHostAlloc((void**)&e->h_out_betas, n * 8, cudaHostAllocDefault));
        CU(cudaHostAlloc((void**)&e->h_out_comp, n * 4, cudaHostAllocDefault));
        CU(cudaHostAlloc((void**)&e->h_out_state, state_doubles(e) * 8 + state_ints(e) * 4, cudaHostAllocDefault));
        CU(cudaEventCreateWithFlags(&e->ev_staged, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming));
        if (!e->copy_stream) CU(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    }
    if (e->out_pending) CU(cudaEventSynchronize(e->ev_out));          // the previous snapshot's copy still owns the buffers
    CU(cudaMemcpyAsync(e->out_betas.p, e->betas.p, n * 8, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaMemcpyAsync(e->out_comp.p, e->comp.p, n * 4, cudaMemcpyDeviceToDevice, e->stream));
    CU(cudaEventRecord(e->ev_staged, e->stream));
    CU(cudaStreamWaitEvent(e->copy_stream, e->ev_staged, 0));
    CU(cudaMemcpyAsync(e->h_out_betas, e->out_betas.p, n * 8, cudaMemcpyDeviceToHost, e->copy_stream));
    CU(cudaMemcpyAsync(e->h_out_comp, e->out_comp.p, n * 4, cudaMemcpyDeviceToHost, e->copy_stream));
    {   // the global parameters of the same iteration (what .csv records): read on the engine's stream, in order with
        // the chain, so that the next iteration cannot overwrite them first -- a few hundred bytes
        const size_t T = e->cfg.T, G = e->cfg.G, K = e->cfg.K;
        double* hd = e->h_out_state;
        int32_t* hi = reinterpret_cast<int32_t*>(hd + state_doubles(e));
        CU(cudaMemcpyAsync(hd, e->sigmag.p, T * G * 8, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(hd + T * G, e->sigmae.p, T * 8, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(hd + T * G + T, e->pi.p, T * G * K * 8, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(hd + T * G + T + T * G * K, e->mu.p, T * 8, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(hi, e->m0.p, T * G * 4, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaMemcpyAsync(hi + T * G, e->cass.p, T * G * K * 4, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaEventRecord(e->ev_staged, e->stream));
        CU(cudaStreamWaitEvent(e->copy_stream, e->ev_staged, 0));
    }
    CU(cudaEventRecord(e->ev_out, e->copy_stream));
    e->out_pending = true;
    return GMRM_OK;
}
int gmrm_fetch_outputs(gmrm_engine* e, int32_t t, double* betas, int32_t* comp) {
    if (!e || t < 0 || t >= e->cfg.T) return fail(GMRM_EINVAL, "bad argument");
    if (!e->out_pending) return fail(GMRM_EINVAL, "no staged outputs: call gmrm_stage_outputs first");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaEventSynchronize(e->ev_out));
    if (betas) memcpy(betas, e->h_out_betas + (size_t)t * e->Mloc, (size_t)e->Mloc * 8);
    if (comp) memcpy(comp, e->h_out_comp + (size_t)t * e->Mloc, (size_t)e->Mloc * 4);
    return GMRM_OK;
}
int gmrm_fetch_state(gmrm_engine* e, gmrm_state* o) {
    if (!e || !o) return fail(GMRM_EINVAL, "null argument");
    if (!e->out_pending) return fail(GMRM_EINVAL, "no staged outputs: call gmrm_stage_outputs first");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaEventSynchronize(e->ev_out));
    const size_t T = e->cfg.T, G = e->cfg.G, K = e->cfg.K;
    const double* hd = e->h_out_state;
    const int32_t* hi = reinterpret_cast<const int32_t*>(hd + state_doubles(e));
    if (o->sigmag) memcpy(o->sigmag, hd, T * G * 8);
    if (o->sigmae) memcpy(o->sigmae, hd + T * G, T * 8);
    if (o->pi) memcpy(o->pi, hd + T * G + T, T * G * K * 8);
    if (o->mu) memcpy(o->mu, hd + T * G + T + T * G * K, T * 8);
    if (o->m0) memcpy(o->m0, hi, T * G * 4);
    if (o->cass) memcpy(o->cass, hi + T * G, T * G * K * 4);
    return GMRM_OK;
}
int gmrm_get_timing(gmrm_engine* e, gmrm_timing* out) {
    if (!e || !out) return fail(GMRM_EINVAL, "null argument");
    *out = e->last;
    return GMRM_OK;
}
int gmrm_set_timing_detail(gmrm_engine* e, int32_t on) {
    if (!e) return fail(GMRM_EINVAL, "null engine");
    e->timing_detail = on < 0 ? 0 : (on > 2 ? 2 : on);
    return GMRM_OK;
}

// ------------------------------------------------------------------------------------ multi-GPU
int gmrm_comm_unique_id(uint8_t id[128]) {
    if (!id) return fail(GMRM_EINVAL, "null argument");
    if (!g_nccl.load()) return fail(GMRM_ENCCL, "libnccl.so.2 not found: %s", dlerror());
    ncclUniqueId u;
    NC(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return GMRM_OK;
}
int gmrm_comm_init(gmrm_engine* e, const uint8_t id[128]) {
    if (!e || !id) return fail(GMRM_EINVAL, "null argument");
    if (!g_nccl.load()) return fail(GMRM_ENCCL, "libnccl.so.2 not found: %s", dlerror());
    CU(cudaSetDevice(e->cfg.device));
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    NC(g_nccl.CommInitRank(&e->comm, e->cfg.world_size, u, e->cfg.world_rank));
    return GMRM_OK;
}

int gmrm_debug_chunk_offset(int32_t rows, int32_t slot, int32_t lane16, int32_t k) {
    if (rows < 1 || rows > kMaxSlots || slot < 0 || slot >= rows || lane16 < 0 || lane16 > 15 || k < 0 || k > 3) return -1;
    return host_chunk_offset(rows, slot, lane16, k);
}

// Test hook, no device needed: the step kernel's launch plan for (N, nsm, V, T) -- traits per launch, rows per pass,
// passes, dynamic shared memory -- and, if `ranges` is given, the rows [start, start+count) every CTA owns in every pass
// (ranges[(pass * nsm + cta) * 2 + {0,1}], room for 64 * nsm * 2 ints).
int gmrm_debug_step_plan(int32_t N, int32_t nsm, int32_t V, int32_t T, int32_t* traits_per_launch, int32_t* rows_per_pass,
                         int32_t* npass, int32_t* smem_bytes, int32_t* nrows, int32_t* ranges) {
    if (N < 1 || nsm < 1 || V < 0 || T < 1) return fail(GMRM_EINVAL, "bad argument");
    const Layout L = make_layout(N, nsm);
    int tc = 0, rpp = 0;
    step_plan(L, V, T, &tc, &rpp);
    if (tc < 1 || rpp < 1) return fail(GMRM_EINVAL, "does not fit");
    const int np = step_npass(L, rpp);
    if (traits_per_launch) *traits_per_launch = tc;
    if (rows_per_pass) *rows_per_pass = rpp;
    if (npass) *npass = np;
    if (smem_bytes) *smem_bytes = step_smem_bytes(L, V, tc, rpp);
    if (nrows) *nrows = L.nrows;
    if (ranges)
        for (int c = 0; c < nsm; c++) {
            int start[64], count[64];
            host_pass_rows(L.nrows, np, nsm, c, start, count);
            for (int q = 0; q < np; q++) { ranges[((size_t)q * nsm + c) * 2] = start[q]; ranges[((size_t)q * nsm + c) * 2 + 1] = count[q]; }
        }
    return GMRM_OK;
}

// Pinned host memory for gmrm_upload_bed sources (true asynchronous DMA) and output buffers.
void* gmrm_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); fail(GMRM_ENOMEM, "cudaHostAlloc of %zu bytes failed", bytes); return nullptr; }
    return p;
}
void gmrm_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// Buffers the other GPUs read in the list exchange: genotypes, missing-list offsets and indices.  Call after
// gmrm_finalize_bed.  Across processes they travel as CUDA IPC handles (3 x 64 bytes); inside one process as plain
// pointers (+ peer access).
int gmrm_comm_export_buffers(gmrm_engine* e, uint8_t handles[384]) {
    if (!e || !handles) return fail(GMRM_EINVAL, "null argument");
    if (!e->bed_final) return fail(GMRM_EINVAL, "call gmrm_finalize_bed first");
    CU(cudaSetDevice(e->cfg.device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, e->bed.p)); memcpy(handles, &h, 64);
    CU(cudaIpcGetMemHandle(&h, e->miss_off.p)); memcpy(handles + 64, &h, 64);
    CU(cudaIpcGetMemHandle(&h, e->miss_idx.p)); memcpy(handles + 128, &h, 64);
    CU(cudaIpcGetMemHandle(&h, e->plist.p)); memcpy(handles + 192, &h, 64);
    CU(cudaIpcGetMemHandle(&h, e->eps.p)); memcpy(handles + 256, &h, 64);
    CU(cudaIpcGetMemHandle(&h, e->rflags.p)); memcpy(handles + 320, &h, 64);
    e->buffers_exported = true;
    return GMRM_OK;
}
static int set_peer(gmrm_engine* e, int rank, void* const p[6]) {
    if (rank < 0 || rank >= e->cfg.world_size || rank == e->cfg.world_rank) return fail(GMRM_EINVAL, "bad peer rank %d", rank);
    if (!e->peer_bed[rank]) e->peers_set++;
    e->peer_bed[rank] = (const uint8_t*)p[0]; e->peer_moff[rank] = (const uint32_t*)p[1]; e->peer_midx[rank] = (const uint32_t*)p[2];
    e->peer_plist[rank] = (double*)p[3];
    e->peer_eps[rank] = (double*)p[4]; e->peer_rflags[rank] = (unsigned long long*)p[5];
    const int me = e->cfg.world_rank;
    e->peer_bed[me] = e->bed.p; e->peer_moff[me] = e->miss_off.p; e->peer_midx[me] = e->miss_idx.p;
    e->peer_plist[me] = e->plist.p;
    e->peer_eps[me] = e->eps.p; e->peer_rflags[me] = e->rflags.p;
    return GMRM_OK;
}
int gmrm_comm_import_buffers(gmrm_engine* e, int32_t rank, const uint8_t handles[384]) {
    if (!e || !handles) return fail(GMRM_EINVAL, "null argument");
    if (!e->bed_final) return fail(GMRM_EINVAL, "call gmrm_finalize_bed first");
    CU(cudaSetDevice(e->cfg.device));
    if (rank < 0 || rank >= e->cfg.world_size || rank == e->cfg.world_rank) return fail(GMRM_EINVAL, "bad peer rank %d", rank);
    void* p[6];
    for (int i = 0; i < 6; i++) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * i, 64);
        CU(cudaIpcOpenMemHandle(&p[i], h, cudaIpcMemLazyEnablePeerAccess));
        e->ipc_opened[rank][i] = p[i];
    }
    return set_peer(e, rank, p);
}
int gmrm_comm_local_buffers(gmrm_engine* e, void* ptrs[6]) {
    if (!e || !ptrs) return fail(GMRM_EINVAL, "null argument");
    if (!e->bed_final) return fail(GMRM_EINVAL, "call gmrm_finalize_bed first");
    ptrs[0] = e->bed.p; ptrs[1] = e->miss_off.p; ptrs[2] = e->miss_idx.p; ptrs[3] = e->plist.p;
    ptrs[4] = e->eps.p; ptrs[5] = e->rflags.p;
    e->buffers_exported = true;
    return GMRM_OK;
}
int gmrm_comm_set_peer_buffers(gmrm_engine* e, int32_t rank, int32_t peer_device, void* const ptrs[6]) {
    if (!e || !ptrs) return fail(GMRM_EINVAL, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    const cudaError_t pe = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return fail(GMRM_ECUDA, "no peer access from device %d to %d: %s", e->cfg.device, peer_device, cudaGetErrorString(pe));
    cudaGetLastError();
    return set_peer(e, rank, ptrs);
}

}  // extern "C"

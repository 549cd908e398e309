// C-ABI implementation (include/sdv.h) — host side of the B200 sliding-window BA/VIO backend.
//
// Replaces, behind the C ABI, the reference's Ceres problem build + solve:
//   AOptimizer::localMapVIOptimization / localMapBA   (cpp/src/optimizers/AOptimizer.cpp:299-446)
//   addResidualsLocalMap / addIMUResiduals / addMarginalizationResiduals
//   (AngularAdjustmentCERESAnalytic.cpp:212-339, AOptimizer.cpp:22-96, …Analytic.cpp:341-486)
// There is NO CPU fallback: without a CUDA device sdv_create fails with SDV_ERR_NO_DEVICE.
#include "../../include/sdv.h"
#include "sdv_kernels.cuh"
#include "sdv_fused.cuh"
#include "sdv_chol.cuh"
#include "sdv_chol_band.cuh"
#include "sdv_preint.cuh"
#include "sdv_marg.cuh"
#include "sdv_peer.cuh"
#include "sdv_struct.cuh"
#include "sdv_viinit.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <thread>
#include <vector>

using namespace sdv;

namespace {

struct Arena { // host-pinned mirror of one contiguous device allocation
    size_t size = 0;
    std::vector<std::pair<size_t, size_t>> parts;
    size_t add(size_t bytes) {
        size_t off = (size + 255) & ~size_t(255);
        size = off + bytes;
        return off;
    }
};

// NCCL through dlopen so that single-GPU use has no NCCL dependency
struct NcclUid { char internal[128]; };
typedef int (*nccl_get_uid_t)(NcclUid *);
typedef int (*nccl_init_rank_t)(void **, int, NcclUid, int);
typedef int (*nccl_allreduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_destroy_t)(void *);
struct NcclApi {
    void *lib = nullptr;
    nccl_get_uid_t get_uid = nullptr;
    nccl_init_rank_t init_rank = nullptr;
    nccl_allreduce_t allreduce = nullptr;
    nccl_destroy_t destroy = nullptr;
    bool load() {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        get_uid = (nccl_get_uid_t)dlsym(lib, "ncclGetUniqueId");
        init_rank = (nccl_init_rank_t)dlsym(lib, "ncclCommInitRank");
        allreduce = (nccl_allreduce_t)dlsym(lib, "ncclAllReduce");
        destroy = (nccl_destroy_t)dlsym(lib, "ncclCommDestroy");
        return get_uid && init_rank && allreduce && destroy;
    }
};
NcclApi g_nccl;
constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0;

} // namespace

// A few persistent host threads per handle: the structure pass of an upload (validation, slot lists) runs on them in parallel
// and one of them packs the bulk arena and issues its H2D copy, so the transfer overlaps the structure pass.
struct HostPool {
    // Jobs are short (tens of microseconds) and arrive in bursts at the start of an upload; a futex wake-up costs as much as a
    // job (more inside a VM).  Hence: the submitting thread SPINS on the group counter, and a worker keeps polling for ~1.5 ms
    // after its last job before it blocks on the condition variable — back-to-back solves never pay a wake-up, an idle handle
    // burns no CPU.
    struct Group { std::atomic<int> pending{0}; }; // jobs submitted together
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv;
    std::vector<std::pair<std::function<void()>, Group *>> jobs; // guarded by m
    size_t next = 0;                                              // guarded by m
    std::atomic<int> avail{0};                                    // queued, not yet taken
    std::atomic<bool> stop{false};
    bool start(int n) {
        try {
            for (int i = (int)th.size(); i < n; i++) th.emplace_back([this] { loop(); });
        } catch (...) {
            return !th.empty();
        }
        return true;
    }
    void loop() {
        auto idle_since = std::chrono::steady_clock::now();
        for (;;) {
            std::pair<std::function<void()>, Group *> job;
            bool have = false;
            if (avail.load(std::memory_order_acquire) > 0 || std::chrono::steady_clock::now() - idle_since > std::chrono::microseconds(1500)) {
                std::unique_lock<std::mutex> lk(m);
                if (avail.load(std::memory_order_acquire) == 0) cv.wait(lk, [this] { return stop.load() || next < jobs.size(); });
                if (stop.load()) return;
                if (next < jobs.size()) {
                    job = std::move(jobs[next++]);
                    avail.fetch_sub(1, std::memory_order_acq_rel);
                    have = true;
                }
            } else {
                if (stop.load(std::memory_order_relaxed)) return;
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            if (have) {
                job.first();
                job.second->pending.fetch_sub(1, std::memory_order_acq_rel);
                idle_since = std::chrono::steady_clock::now();
            }
        }
    }
    void submit(Group &g, std::function<void()> f) {
        g.pending.fetch_add(1, std::memory_order_acq_rel);
        {
            std::lock_guard<std::mutex> lk(m);
            if (next == jobs.size()) { // nothing queued: recycle the job list
                jobs.clear();
                next = 0;
            }
            jobs.emplace_back(std::move(f), &g);
            avail.fetch_add(1, std::memory_order_acq_rel);
        }
        cv.notify_one();
    }
    void wait(Group &g) {
        while (g.pending.load(std::memory_order_acquire) != 0) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m);
            stop.store(true);
        }
        cv.notify_all();
        for (auto &t : th)
            if (t.joinable()) t.join();
    }
};

typedef void (*lin_visual_fn_t)(const sdv::DevProblem *, sdv::LinBuf, sdv::LinBuf, const sdv::LMState *, sdv::Accum *, int);
static lin_visual_fn_t lin_visual_fn(int kind, bool smem, bool early) {
    using namespace sdv;
    if (kind == SDV_FACTOR_ANGULAR) {
        if (smem) return early ? k_lin_visual<0, true, true> : k_lin_visual<0, true, false>;
        return early ? k_lin_visual<0, false, true> : k_lin_visual<0, false, false>;
    }
    if (smem) return early ? k_lin_visual<1, true, true> : k_lin_visual<1, true, false>;
    return early ? k_lin_visual<1, false, true> : k_lin_visual<1, false, false>;
}

struct MargState;

struct sdv_handle {
    sdv_config cfg;
    SolverOpts opt;
    int device = 0, num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;
    // input arena
    unsigned char *d_in = nullptr, *h_in = nullptr;
    size_t in_cap = 0, in_bytes = 0, h2d_last = 0;
    unsigned char *d_in2 = nullptr, *h_in2 = nullptr; // bulk data arena (measurements, observation indices, landmarks), packed by a helper thread
    size_t in2_cap = 0, in2_bytes = 0;
    // scratch
    unsigned char *d_scr = nullptr;
    size_t scr_cap = 0;
    // small pinned readback
    unsigned char *h_rb = nullptr;
    size_t rb_cap = 0;
    HostPool pool;                        // see HostPool
    cudaStream_t copy_stream = nullptr;   // H2D of the bulk arena, issued by a pool thread while the structure pass runs
    cudaEvent_t ev_bulk = nullptr;
    unsigned char *d_flush = nullptr; // 256 MiB L2-flush buffer of sdv_time_kernel's cold variants (allocated on first use)
    unsigned char *d_out = nullptr; // solution blocks
    size_t out_cap = 0;
    DevProblem P;              // host copy of the problem description of the resident window
    DevProblem *d_P = nullptr; // its device copy (head of the input arena): the ONLY window-dependent kernel argument
    LinBuf B[2];
    LMState *d_st = nullptr;
    Accum *d_acc = nullptr;
    double *d_Sb = nullptr, *d_Lo = nullptr, *d_scale_p = nullptr, *d_damp_p = nullptr, *d_graw_p = nullptr, *d_dxp = nullptr,
           *d_scale_l = nullptr, *d_red = nullptr;
    size_t sb_elems = 0;
    bool resident = false;
    int lin_grid = 0, lin_smem = 0, fac_grid = 0;
    lin_visual_fn_t lin_fn = nullptr;
    int chol_cluster = 0, chol_rows_roles = 0, chol_smem_chain = 0; // wide-band fallback k_chol_chain: cluster size (0 = per-panel launches), tile rows per CTA, dynamic smem
    double *d_partial = nullptr, *d_dinv = nullptr, *d_prof = nullptr;
    int64_t launches = 0;
    // whole-solve CUDA graph: prologue -> WHILE(LM iteration) -> epilogue (single-GPU only)
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    cudaStream_t stream2 = nullptr;
    cudaStream_t side = nullptr;           // fork/join branch for the small factor kernels (runs concurrently with the landmark kernels)
    cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
    unsigned long long cond = 0;
    bool graph_ok = false;
    // Everything the captured launches depend on.  The problem description itself is read from device memory (d_P), the
    // scratch arena is laid out from CAPACITIES that only grow, grids are persistent upper bounds: consecutive windows of a
    // running back end (same keyframe count, a few landmarks more or less) replay ONE instantiated graph.
    struct GraphSig {
        const void *d_in, *d_scr, *d_out;
        int F, C, vio, kind, n_pad, band_bw, band_smem, chol_cluster, chol_rows_roles, chol_smem_chain;
        int cap_O, cap_L, cap_P, cap_nm, cap_nfull, cap_l2l;
        int fused_grid, fused_grid_back, fac_grid, cost_grid, p2l_grid, flags, world;
    } graph_sig;
    int cap_O = 0, cap_L = 0, cap_P = 0, cap_nm = 0, cap_nfull = 0, cap_l2l = 0; // scratch-layout capacities (only grow)
    int cost_grid = 0, p2l_grid = 0;
    int64_t graph_launches_fixed = 0, graph_launches_iter = 0, graph_builds = 0;
    // k LM iterations per trip of the WHILE node (a trip costs ~6 us on B200; every kernel returns at once after termination, so a
    // trip that is cut short only launches a few empty kernels): k = the largest divisor <= 4 of the PREVIOUS solve's iteration count —
    // consecutive windows of a running back end take the same number of iterations.  Instantiated graphs that fall out of use are
    // parked (keyed by signature + k, least recently used evicted): a handle that alternates between kinds of solves — the front-end
    // optimizer instance runs landmarkOptimization and the single-frame solves in turn — or between iteration counts replays them.
    int graph_unroll = 1;
    static constexpr int GRAPH_CACHE = 8;
    struct GraphSlot {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t gexec = nullptr;
        unsigned long long cond = 0;
        GraphSig sig;
        int unroll = 0;
        int64_t l_fixed = 0, l_iter = 0;
        uint64_t last_use = 0;
    } gcache[GRAPH_CACHE];
    uint64_t graph_clock = 0;
    unsigned char *h_sol = nullptr;
    size_t sol_cap = 0;
    // comm
    void *comm = nullptr;
    int rank = 0, world = 1;
    std::vector<int> tmp_lmk_ptr, tmp_slot_ptr, tmp_slot_frame, tmp_slot_obs_ptr, tmp_slot_obs; // reused between uploads
    std::vector<char> tmp_same_prev;
    std::vector<int> tmp_tile_ptr;
    double *d_xchg = nullptr;    // multi-GPU exchange buffer (k_band_exchange)
    int last_iters = 3;          // LM iterations of the previous solve: how many the host enqueues before it first looks at the status (N > 1)
    double *d_lmk_aux = nullptr; // [L][LMK_AUX]: V^-1, g_l, D_l of every eliminated landmark (k_lin_schur -> k_backsub_cost)
    int fused_grid = 0, fused_grid_back = 0;
    std::vector<uint32_t> tmp_tile_nz;
    bool attrs_done = false, viinit_attr_done = false;
    const void *lin_fn_cached = nullptr;
    int lin_smem_cached = -1, lin_per_sm = 1;
    int band_bw = -1, band_smem = 0; // 16-column block half-bandwidth of the reduced system (-1 = not computed), k_chol_band shared memory
    int chol_tiles_nz = 0, chol_tiles_all = 0; // structurally non-zero tiles of L / all lower tiles
    LMState h_state;
    Accum h_acc;
    MargState *marg = nullptr; // scratch and result of the last sdv_marginalize (sdv_marg_host.cuh)
    // peer-memory exchange (sdv_peer.cuh): own area, the peers' areas opened through CUDA IPC
    unsigned char *d_peer = nullptr;
    void *peer_base[PEER_MAX_WORLD] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    PeerXchg peer{};
    bool peer_ok = false;
    size_t peer_bytes = 0;
    cudaGraphExec_t sgraph_exec = nullptr; // the eight launches of the device structure pass, replayed per upload
    const void *sgraph_dP = nullptr;
    int sgraph_capO = 0, sgraph_capL = 0;
    bool force_host_slots = false; // upload: the device structure pass (sdv_struct.cuh) does not apply to this window, redo on the host
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return SDV_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

int fail(sdv_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    return code;
}

template <class T> T *at(unsigned char *base, size_t off) { return reinterpret_cast<T *>(base + off); }

int ensure(sdv_handle *h, unsigned char **d, size_t *cap, size_t need, bool pinned_host = false, bool write_combined = false) {
    if (need <= *cap) return SDV_OK;
    size_t ncap = std::max(need, *cap * 2);
    if (pinned_host) {
        if (*d) cudaFreeHost(*d);
        // the input arena is only ever written by the CPU (sequentially) and read by the copy engine: write-combined
        CK(cudaHostAlloc((void **)d, ncap, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    } else {
        if (*d) cudaFree(*d);
        CK(cudaMalloc((void **)d, ncap));
    }
    *cap = ncap;
    return SDV_OK;
}

} // namespace

namespace {
int build_solve_graph(sdv_handle *h);
void destroy_graph(sdv_handle *h);
void free_marg_state(sdv_handle *h);
}

extern "C" {

int sdv_abi_version(void) { return SDV_ABI_VERSION; }

const char *sdv_strerror(int s) {
    switch (s) {
    case SDV_OK: return "ok";
    case SDV_ERR_INVALID_ARGUMENT: return "invalid argument";
    case SDV_ERR_CUDA: return "CUDA error";
    case SDV_ERR_NO_DEVICE: return "no CUDA device (this backend has no CPU path)";
    case SDV_ERR_UNSUPPORTED: return "unsupported input";
    case SDV_ERR_NUMERICAL_FAILURE: return "numerical failure";
    case SDV_ERR_COMM: return "communicator error";
    }
    return "unknown status";
}

const char *sdv_last_error(const sdv_handle *h) { return h ? h->err.c_str() : ""; }

void sdv_default_config(sdv_config *c) {
    std::memset(c, 0, sizeof(*c));
    c->abi_version = SDV_ABI_VERSION;
    c->device = 0;
    c->max_num_iterations = 20;           // AOptimizer.cpp:380
    c->max_consecutive_invalid_steps = 5; // Ceres 2.2 default
    c->jacobi_scaling = 1;                // Ceres 2.2 default
    c->function_tolerance = 1e-3;         // AOptimizer.cpp:384
    c->gradient_tolerance = 1e-10;
    c->parameter_tolerance = 1e-8;
    c->initial_trust_region_radius = 1e4;
    c->max_trust_region_radius = 1e16;
    c->min_trust_region_radius = 1e-32;
    c->min_lm_diagonal = 1e-6;
    c->max_lm_diagonal = 1e32;
    c->min_relative_decrease = 1e-3;
}

int sdv_create(sdv_handle **out, const sdv_config *cfg) {
    if (!out || !cfg || cfg->abi_version != SDV_ABI_VERSION) return SDV_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device >= ndev) return SDV_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return SDV_ERR_NO_DEVICE;
    if (prop.major < 10) return SDV_ERR_NO_DEVICE; // kernels are built for sm_100a only
    sdv_handle *h = new sdv_handle();
    h->cfg = *cfg;
    h->device = cfg->device;
    h->num_sms = prop.multiProcessorCount;
    SolverOpts &o = h->opt;
    o.max_num_iterations = cfg->max_num_iterations;
    o.max_consecutive_invalid_steps = cfg->max_consecutive_invalid_steps;
    o.jacobi_scaling = cfg->jacobi_scaling;
    o.function_tolerance = cfg->function_tolerance;
    o.gradient_tolerance = cfg->gradient_tolerance;
    o.parameter_tolerance = cfg->parameter_tolerance;
    o.initial_radius = cfg->initial_trust_region_radius;
    o.max_radius = cfg->max_trust_region_radius;
    o.min_radius = cfg->min_trust_region_radius;
    o.min_diag = cfg->min_lm_diagonal;
    o.max_diag = cfg->max_lm_diagonal;
    o.min_relative_decrease = cfg->min_relative_decrease;
    if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return SDV_ERR_CUDA;
    }
    for (int i = 0; i < 4; i++) cudaEventCreate(&h->ev[i]);
    cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_bulk, cudaEventDisableTiming);
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&h->ev_fork[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming);
    }
    std::memset(&h->P, 0, sizeof(h->P));
    *out = h;
    return SDV_OK;
}

int sdv_destroy(sdv_handle *h) {
    if (!h) return SDV_ERR_INVALID_ARGUMENT;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->comm && g_nccl.destroy) g_nccl.destroy(h->comm);
    destroy_graph(h);
    if (h->sgraph_exec) cudaGraphExecDestroy(h->sgraph_exec);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->ev_bulk) cudaEventDestroy(h->ev_bulk);
    for (int i = 0; i < 2; i++) {
        if (h->ev_fork[i]) cudaEventDestroy(h->ev_fork[i]);
        if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    }
    if (h->h_sol) cudaFreeHost(h->h_sol);
    if (h->d_in) cudaFree(h->d_in);
    if (h->h_in) cudaFreeHost(h->h_in);
    if (h->d_in2) cudaFree(h->d_in2);
    if (h->h_in2) cudaFreeHost(h->h_in2);
    if (h->d_scr) cudaFree(h->d_scr);
    if (h->h_rb) cudaFreeHost(h->h_rb);
    if (h->d_out) cudaFree(h->d_out);
    if (h->d_flush) cudaFree(h->d_flush);
    free_marg_state(h);
    for (int p = 0; p < PEER_MAX_WORLD; p++)
        if (h->peer_base[p] && p != h->rank) cudaIpcCloseMemHandle(h->peer_base[p]);
    if (h->d_peer) cudaFree(h->d_peer);
    for (int i = 0; i < 4; i++)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    cudaStreamDestroy(h->stream);
    delete h;
    return SDV_OK;
}

// Landmark shard of `rank`: contiguous landmark range [l0, l1) balanced by observation count, and its observation
// range [o0, o1).  Pure host logic (no CUDA call) so that it can be tested without a GPU.
int sdv_shard_range(const int32_t *obs_lmk, int32_t n_obs, int32_t n_lmks, int32_t rank, int32_t world, int32_t *l0, int32_t *l1,
                    int32_t *o0, int32_t *o1) {
    if (n_obs < 0 || n_lmks < 0 || world < 1 || rank < 0 || rank >= world || (n_obs > 0 && !obs_lmk) || !l0 || !l1 || !o0 || !o1)
        return SDV_ERR_INVALID_ARGUMENT;
    std::vector<int> ptr(n_lmks + 1, 0);
    int o = 0;
    for (int l = 0; l < n_lmks; l++) {
        ptr[l] = o;
        while (o < n_obs && obs_lmk[o] == l) o++;
    }
    if (o != n_obs) return SDV_ERR_INVALID_ARGUMENT; // not landmark-major
    ptr[n_lmks] = n_obs;
    auto cut = [&](int r) {
        long long target = (long long)n_obs * r / world;
        int l = (int)(std::lower_bound(ptr.begin(), ptr.end(), (int)target) - ptr.begin());
        return std::min(l, (int)n_lmks);
    };
    *l0 = rank == 0 ? 0 : cut(rank);
    *l1 = rank == world - 1 ? n_lmks : cut(rank + 1);
    *o0 = ptr[*l0];
    *o1 = ptr[*l1];
    return SDV_OK;
}

int sdv_comm_unique_id(void *out) {
    if (!out) return SDV_ERR_INVALID_ARGUMENT;
    if (!g_nccl.load()) return SDV_ERR_COMM;
    NcclUid id;
    if (g_nccl.get_uid(&id) != 0) return SDV_ERR_COMM;
    std::memcpy(out, &id, sizeof(id));
    return SDV_OK;
}

int sdv_comm_init(sdv_handle *h, const void *uid, int32_t rank, int32_t world) {
    if (!h || !uid || world < 1 || rank < 0 || rank >= world) return SDV_ERR_INVALID_ARGUMENT;
    // a graph captured before the communicator existed bakes the old window and contains no NCCL call: never reuse it
    destroy_graph(h);
    h->resident = false;
    if (world == 1) {
        h->rank = 0;
        h->world = 1;
        return SDV_OK;
    }
    if (!g_nccl.load()) return fail(h, SDV_ERR_COMM, "libnccl.so.2 not loadable");
    cudaSetDevice(h->device);
    NcclUid id;
    std::memcpy(&id, uid, sizeof(id));
    if (g_nccl.init_rank(&h->comm, world, id, rank) != 0) return fail(h, SDV_ERR_COMM, "ncclCommInitRank failed");
    h->rank = rank;
    h->world = world;
    h->resident = false;
    return SDV_OK;
}

// Peer-memory exchange (sdv_peer.cuh).  Layout of an exchange area, identical on every rank:
//   [data 2 x world x cap doubles | flags 2 x world x PEER_MAX_CHUNKS u64 | epoch 8 u64]
static size_t peer_layout(int world, size_t cap, size_t *o_flags, size_t *o_epoch) {
    const size_t data = sizeof(double) * 2 * (size_t)world * cap;
    *o_flags = data;
    *o_epoch = data + sizeof(unsigned long long) * 2 * (size_t)world * PEER_MAX_CHUNKS;
    return *o_epoch + 64;
}
static size_t peer_cap_doubles() {
    const char *e = getenv("SDV_PEER_CAP_DOUBLES");
    return e ? (size_t)std::max(1024ll, atoll(e)) : ((size_t)1 << 20); // 8 MB per (parity, source) slot
}

int sdv_comm_peer_handle(sdv_handle *h, void *out_64_bytes) {
    if (!h || !out_64_bytes) return SDV_ERR_INVALID_ARGUMENT;
    if (h->world < 2 || h->world > PEER_MAX_WORLD) return fail(h, SDV_ERR_INVALID_ARGUMENT, "peer exchange needs 2..8 ranks (call sdv_comm_init first)");
    cudaSetDevice(h->device);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!h->d_peer) {
        size_t of, oe;
        h->peer_bytes = peer_layout(h->world, peer_cap_doubles(), &of, &oe);
        CK(cudaMalloc((void **)&h->d_peer, h->peer_bytes));
        CK(cudaMemset(h->d_peer, 0, h->peer_bytes));
    }
    cudaIpcMemHandle_t hd;
    CK(cudaIpcGetMemHandle(&hd, h->d_peer));
    std::memcpy(out_64_bytes, &hd, 64);
    return SDV_OK;
}

int sdv_comm_peer_open(sdv_handle *h, const void *handles) {
    if (!h || !handles) return SDV_ERR_INVALID_ARGUMENT;
    if (!h->d_peer) return fail(h, SDV_ERR_INVALID_ARGUMENT, "sdv_comm_peer_handle first");
    cudaSetDevice(h->device);
    size_t of, oe;
    const size_t cap = peer_cap_doubles();
    peer_layout(h->world, cap, &of, &oe);
    PeerXchg X;
    std::memset(&X, 0, sizeof(X));
    X.rank = h->rank;
    X.world = h->world;
    X.cap = cap;
    for (int p = 0; p < h->world; p++) {
        void *base = h->d_peer;
        if (p != h->rank) {
            cudaIpcMemHandle_t hd;
            std::memcpy(&hd, static_cast<const unsigned char *>(handles) + 64 * (size_t)p, 64);
            cudaError_t e = cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(h, SDV_ERR_COMM, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
            }
        }
        h->peer_base[p] = base;
        X.data[p] = reinterpret_cast<double *>(base);
        X.flags[p] = reinterpret_cast<unsigned long long *>(static_cast<unsigned char *>(base) + of);
    }
    X.epoch = reinterpret_cast<unsigned long long *>(h->d_peer + oe);
    h->peer = X;
    h->peer_ok = getenv("SDV_NO_PEER") == nullptr;
    destroy_graph(h);
    h->resident = false;
    return SDV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// upload: validate, derive the reduced-program structure, pack one pinned arena, one H2D copy
// ---------------------------------------------------------------------------------------------------------------------
static int upload_impl(sdv_handle *h, const sdv_window *w, bool sync);
int sdv_upload_window(sdv_handle *h, const sdv_window *w) { return upload_impl(h, w, true); }

// sync = false (sdv_solve_window): the solve is stream-ordered behind the copies and the set-up kernels, the host does not wait
static int upload_impl(sdv_handle *h, const sdv_window *w, bool sync) {
    if (!h || !w) return SDV_ERR_INVALID_ARGUMENT;
    auto t_entry = std::chrono::steady_clock::now();
    h->resident = false; // an upload that fails midway must not leave a half-updated problem marked resident
    if (w->abi_version != SDV_ABI_VERSION) return fail(h, SDV_ERR_INVALID_ARGUMENT, "abi_version mismatch");
    const int F = w->n_frames, C = w->n_cams, L = w->n_lmks, O = w->n_obs, Pn = w->vio ? w->n_imu : 0;
    if (F <= 0 || C <= 0 || L < 0 || O < 0 || Pn < 0 || w->n_fixed < 0) return fail(h, SDV_ERR_INVALID_ARGUMENT, "negative or empty sizes");
    if (!w->T_f_w || !w->T_s_f || !w->K || (L > 0 && !w->lmk_t)) return fail(h, SDV_ERR_INVALID_ARGUMENT, "null frame/camera/landmark arrays");
    if (O > 0 && (!w->obs_lmk || !w->obs_frame || !w->obs_cam)) return fail(h, SDV_ERR_INVALID_ARGUMENT, "null observation index arrays");
    const int kind = w->factor_kind;
    if (kind != SDV_FACTOR_ANGULAR && kind != SDV_FACTOR_PIXEL) return fail(h, SDV_ERR_INVALID_ARGUMENT, "bad factor_kind");
    if (O > 0 && kind == SDV_FACTOR_ANGULAR && !w->obs_bearing) return fail(h, SDV_ERR_INVALID_ARGUMENT, "obs_bearing is null");
    if (O > 0 && kind == SDV_FACTOR_PIXEL && !w->obs_uv) return fail(h, SDV_ERR_INVALID_ARGUMENT, "obs_uv is null");
    if (w->vio && (!w->v || !w->ba || !w->bg)) return fail(h, SDV_ERR_INVALID_ARGUMENT, "vio window without v/ba/bg");
    if (Pn > 0 && (!w->imu_i || !w->imu_j || !w->imu_dt || !w->imu_dR || !w->imu_dv || !w->imu_dp || !w->imu_cov || !w->imu_J_dR_bg ||
                   !w->imu_J_dv_ba || !w->imu_J_dv_bg || !w->imu_J_dp_ba || !w->imu_J_dp_bg || !w->imu_sigma_ba || !w->imu_sigma_bg))
        return fail(h, SDV_ERR_INVALID_ARGUMENT, "null IMU arrays");
    if (w->has_prior && (!w->T_prior || !w->inf_prior)) return fail(h, SDV_ERR_INVALID_ARGUMENT, "has_prior without T_prior/inf_prior");
    const sdv_sparse_prior *sp = w->sparse_prior;
    if (sp) {
        if (sp->n_p2l < 0 || sp->n_l2l < 0) return fail(h, SDV_ERR_INVALID_ARGUMENT, "malformed sparse prior");
        if (sp->has_imu_prior && (sp->frame < 0 || sp->frame >= F || !w->vio)) return fail(h, SDV_ERR_INVALID_ARGUMENT, "sparse prior frame");
        if (sp->n_p2l > 0 && (!sp->has_imu_prior || !sp->p2l_lmk || !sp->p2l_delta || !sp->p2l_sqrt_inf))
            return fail(h, SDV_ERR_INVALID_ARGUMENT, "PoseToLandmark factors need the kept frame and their arrays");
        for (int k = 0; k < sp->n_p2l; k++)
            if (sp->p2l_lmk[k] < 0 || sp->p2l_lmk[k] >= L) return fail(h, SDV_ERR_INVALID_ARGUMENT, "p2l landmark out of range");
        if (sp->has_lmk_prior && (sp->lmk0 < 0 || sp->lmk0 >= L)) return fail(h, SDV_ERR_INVALID_ARGUMENT, "lmk0 out of range");
        if (sp->n_l2l > 0 && (!sp->l2l_a || !sp->l2l_b || !sp->l2l_delta || !sp->l2l_sqrt_inf))
            return fail(h, SDV_ERR_INVALID_ARGUMENT, "null LandmarkToLandmark arrays");
        for (int k = 0; k < sp->n_l2l; k++)
            if (sp->l2l_a[k] < 0 || sp->l2l_a[k] >= L || sp->l2l_b[k] < 0 || sp->l2l_b[k] >= L || sp->l2l_a[k] == sp->l2l_b[k])
                return fail(h, SDV_ERR_INVALID_ARGUMENT, "l2l landmark out of range");
    }
    cudaSetDevice(h->device);
    // ---- bulk data arena: the arrays that are pure copies (78 % of the bytes at C3) have a layout that depends on the sizes
    //      only, so a helper thread packs them into pinned memory while this thread validates the indices and derives the
    //      structure of the reduced system; both arenas then go to the device with one asynchronous copy each.
    const int mplanes = kind == SDV_FACTOR_ANGULAR ? 3 : 2;
    Arena A2;
    const size_t q_om = A2.add(sizeof(double) * mplanes * std::max(O, 1)), q_ol = A2.add(4 * std::max(O, 1)), q_ofc = A2.add(4 * std::max(O, 1));
    const size_t q_lt = A2.add(sizeof(double) * 3 * std::max(L, 1)), q_ow = w->obs_sigma ? A2.add(sizeof(double) * std::max(O, 1)) : 0;
    {
        int rc2;
        if ((rc2 = ensure(h, &h->h_in2, &h->in2_cap, A2.size, true)) != SDV_OK) return rc2;
        if (A2.size > h->in2_bytes) {
            if (h->d_in2) cudaFree(h->d_in2);
            h->d_in2 = nullptr;
            CK(cudaMalloc((void **)&h->d_in2, A2.size * 2));
            h->in2_bytes = A2.size * 2;
        }
    }
    // 3 packing jobs + 3 structure parts beside this thread (SDV_HOST_PARTS up to 8 was measured on the 16-thread box of the pool:
    // no gain over 4 — more spinning workers cost what the shorter parts save)
    const bool have_pool = h->pool.start(getenv("SDV_HOST_PARTS") ? 12 : 8) && !getenv("SDV_NO_HOST_POOL");
    HostPool::Group g_bulk, g_struct;
    bool bulk_copy_issued = false; // written by the packing job that finishes last, read after pool.wait(g_bulk)
    std::atomic<int> bulk_failed{0};
    constexpr int NPIECE = 5;
    std::atomic<int> bulk_left{NPIECE};
    {
        unsigned char *hb2 = h->h_in2, *db2 = h->d_in2;
        const int dev = h->device;
        cudaStream_t cs = h->copy_stream;
        cudaEvent_t evb = h->ev_bulk;
        // five pieces of similar size (measurements in three parts; observation indices; landmarks + weights).  Every piece issues
        // the H2D copy of ITS OWN byte range on the copy stream as soon as it is packed — packing and transfer overlap each other
        // and the structure pass of the other threads —; the piece that finishes last records the event the solve stream waits for.
        auto pack_bulk = [=, &bulk_copy_issued, &bulk_failed, &bulk_left](int piece, bool issue_copy) {
            const size_t mbytes = sizeof(double) * (size_t)mplanes * O, third = (mbytes / 3) & ~size_t(63);
            const unsigned char *msrc = reinterpret_cast<const unsigned char *>(kind == SDV_FACTOR_ANGULAR ? (const void *)w->obs_bearing : (const void *)w->obs_uv);
            size_t c0 = 0, c1 = 0; // byte range of the arena this piece fills
            if (piece < 3 && O) {
                const size_t m0 = third * piece, m1 = piece == 2 ? mbytes : third * (piece + 1);
                std::memcpy(hb2 + q_om + m0, msrc + m0, m1 - m0);
                c0 = q_om + m0;
                c1 = q_om + m1;
            }
            if (piece == 3 && O) {
                std::memcpy(hb2 + q_ol, w->obs_lmk, 4 * (size_t)O);
                int *fc = reinterpret_cast<int *>(hb2 + q_ofc);
                for (int o = 0; o < O; o++) fc[o] = w->obs_frame[o] * C + w->obs_cam[o]; // validated by the other threads; only used if that passes
                c0 = q_ol;
                c1 = q_ofc + 4 * (size_t)O; // (q_ol and q_ofc are adjacent in the arena)
            }
            if (piece == 4) {
                if (L > 0) std::memcpy(hb2 + q_lt, w->lmk_t, sizeof(double) * 3 * (size_t)L);
                c0 = q_lt;
                c1 = q_lt + sizeof(double) * 3 * (size_t)std::max(L, 1);
                if (O && w->obs_sigma) {
                    double *ow = reinterpret_cast<double *>(hb2 + q_ow);
                    for (int o = 0; o < O; o++) ow[o] = 1.0 / w->obs_sigma[o];
                    c1 = q_ow + sizeof(double) * (size_t)O; // (q_lt and q_ow are adjacent in the arena)
                }
            }
            if (!issue_copy) return;
            bool ok = cudaSetDevice(dev) == cudaSuccess;
            if (ok && c1 > c0) ok = cudaMemcpyAsync(db2 + c0, hb2 + c0, c1 - c0, cudaMemcpyHostToDevice, cs) == cudaSuccess;
            if (!ok) bulk_failed.store(1);
            if (bulk_left.fetch_sub(1) == 1) { // every piece has issued its copy: the event closes them all
                if (cudaEventRecord(evb, cs) != cudaSuccess) bulk_failed.store(1);
                bulk_copy_issued = true;
            }
        };
        for (int piece = 0; piece < NPIECE; piece++) {
            if (have_pool) h->pool.submit(g_bulk, [pack_bulk, piece] { pack_bulk(piece, true); });
            else pack_bulk(piece, false);
        }
    }
    struct PoolGuard { // every exit path waits for the jobs that reference this frame's variables
        HostPool *p;
        HostPool::Group *a, *b;
        ~PoolGuard() {
            if (p) {
                p->wait(*a);
                p->wait(*b);
            }
        }
    } pool_guard{have_pool ? &h->pool : nullptr, &g_bulk, &g_struct};
    std::vector<int> &lmk_ptr = h->tmp_lmk_ptr;
    lmk_ptr.assign((size_t)L + 1, 0);
    std::vector<char> pose_used(F, 0), vb_used(F, 0);
    std::vector<int> &slot_ptr = h->tmp_slot_ptr, &slot_frame = h->tmp_slot_frame, &slot_obs_ptr = h->tmp_slot_obs_ptr, &slot_obs = h->tmp_slot_obs;
    std::vector<char> &same_prev = h->tmp_same_prev;
    int ns = 0, nso = 0, max_slots = 1;
    // ---- parallel structure pass (single GPU, no PoseToLandmark pseudo-observations, enough observations to pay): the
    //      observation list is cut at landmark boundaries into one part per thread; every part validates its observations,
    //      fills its stretch of the CSR pointer and builds the slot lists of its landmarks; a short serial step turns the
    //      per-part slot counts into offsets and the parts copy their lists into place.
    const bool par_struct = have_pool && h->world == 1 && O >= 16384 && !(w->sparse_prior && w->sparse_prior->n_p2l > 0);
    // ---- device structure pass (sdv_struct.cuh): the slot lists and tiles are derived on the GPU from the observation arrays of
    //      the bulk arena; the host only validates, counts slots and records the keyframe span of every landmark (the band of the
    //      reduced system).  Windows with landmarks in the reduced system (dense / sparsified-VO priors) and observation lists
    //      that are not grouped by keyframe keep the host path below.
    bool gpu_struct = par_struct && !h->force_host_slots && !getenv("SDV_HOST_SLOTS") && !w->dense_prior &&
                      !(w->sparse_prior && (w->sparse_prior->has_lmk_prior || w->sparse_prior->n_l2l > 0));
    std::vector<int> span_max; // [F]: largest free keyframe index among the landmarks whose smallest free keyframe is f (-1: none)
    if (gpu_struct) {
        constexpr int GP = 4;
        struct alignas(128) GPart {
            std::vector<char> used;
            std::vector<int> span;
            int nslots = 0, max_slots = 1, bad = 0, ungrouped = 0;
        } gp[GP];
        const int32_t *ol = w->obs_lmk, *of = w->obs_frame, *oc = w->obs_cam;
        int cutp[GP + 1];
        cutp[0] = 0;
        cutp[GP] = O;
        for (int t = 1; t < GP; t++) {
            int c0 = (int)((long long)O * t / GP);
            while (c0 < O && c0 > 0 && ol[c0] == ol[c0 - 1]) c0++;
            cutp[t] = std::max(c0, cutp[t - 1]);
        }
        const int first_fixed = F - w->n_fixed; // frames >= first_fixed are constant (no column)
        auto light_part = [&](int t) {
            GPart pt;
            pt.used.assign(F, 0);
            pt.span.assign(F, -1);
            const int ob = cutp[t], oe = cutp[t + 1];
            int prev = ob > 0 ? ol[ob - 1] : -1;
            unsigned bad = (unsigned)(prev < -1) | (unsigned)(prev >= L);
            int frs[MAX_SLOTS + 1], nfr = 0, cur = -1, fmin = F, fmax = -1;
            auto close = [&] {
                pt.max_slots = std::max(pt.max_slots, nfr);
                if (fmax >= 0) pt.span[fmin] = std::max(pt.span[fmin], fmax);
            };
            for (int o = ob; o < oe && !bad; o++) {
                const int l = ol[o], f = of[o], c = oc[o];
                bad |= (unsigned)(l < 0) | (unsigned)(l >= L) | (unsigned)(f < 0) | (unsigned)(f >= F) | (unsigned)(c < 0) | (unsigned)(c >= C) | (unsigned)(l < prev);
                if (bad) break;
                if (l != prev) {
                    close();
                    prev = l;
                    nfr = 0;
                    cur = -1;
                    fmin = F;
                    fmax = -1;
                }
                if (f != cur) { // a new slot — unless this keyframe appeared before in the landmark's list (host grouping then)
                    for (int q = 0; q < nfr && q < MAX_SLOTS; q++) pt.ungrouped |= frs[q] == f;
                    if (nfr < MAX_SLOTS + 1) frs[nfr] = f;
                    nfr = std::min(nfr + 1, MAX_SLOTS + 1);
                    pt.nslots++;
                    cur = f;
                    pt.used[f] = 1;
                    if (f < first_fixed) {
                        fmin = std::min(fmin, f);
                        fmax = std::max(fmax, f);
                    }
                }
            }
            close();
            pt.bad = bad ? 1 : 0;
            gp[t] = std::move(pt);
        };
        for (int t = 1; t < GP; t++) h->pool.submit(g_struct, [&light_part, t] { light_part(t); });
        light_part(0);
        h->pool.wait(g_struct);
        span_max.assign(F, -1);
        int ungrouped = 0;
        for (int t = 0; t < GP; t++) {
            if (gp[t].bad) return fail(h, SDV_ERR_INVALID_ARGUMENT, "observation index out of range or observations not landmark-major (reference walk order)");
            ungrouped |= gp[t].ungrouped;
            ns += gp[t].nslots;
            max_slots = std::max(max_slots, gp[t].max_slots);
            for (int f = 0; f < F; f++) {
                pose_used[f] |= gp[t].used[f];
                span_max[f] = std::max(span_max[f], gp[t].span[f]);
            }
        }
        if (max_slots > MAX_SLOTS) return fail(h, SDV_ERR_UNSUPPORTED, "a landmark is observed from more than 32 keyframes (kernel limit of this build)");
        nso = O;
        if (ungrouped) { // a keyframe re-appears after another one inside a landmark's list: general grouping, on the host
            h->force_host_slots = true;
            const int rc2 = upload_impl(h, w, sync);
            h->force_host_slots = false;
            return rc2;
        }
    }
    if (par_struct && !gpu_struct) {
        static constexpr int MAXPART = 8;
        static const int NPART = [] { // parts of the structure pass (this thread + pool threads); SDV_HOST_PARTS for A/B measurements
            const char *e = getenv("SDV_HOST_PARTS");
            return std::max(1, std::min(MAXPART, e ? atoi(e) : 4));
        }();
        struct alignas(128) Part { // (one cache line pair each: the threads update their own counters all the time)
            int ob = 0, oe = 0;          // observation range, starting at the first observation of a landmark
            int l_lo = 0, l_hi = -1;     // landmarks whose CSR pointer / slot count this part owns: (last landmark before ob, last landmark in range]
            std::vector<int> sf, sop, cnt; // slot frames, slot observation pointers (global: slot_obs is written in place), slots per landmark
            std::vector<char> used, same;
            int max_slots = 1, bad = 0;
        } part[MAXPART];
        const int32_t *ol = w->obs_lmk, *of = w->obs_frame, *oc = w->obs_cam;
        slot_obs.resize((size_t)O + 1);
        same_prev.resize((size_t)L + 1);
        int cutp[MAXPART + 1];
        cutp[0] = 0;
        cutp[NPART] = O;
        for (int t = 1; t < NPART; t++) {
            int c0 = (int)((long long)O * t / NPART);
            while (c0 < O && c0 > 0 && ol[c0] == ol[c0 - 1]) c0++;
            cutp[t] = std::max(c0, cutp[t - 1]);
        }
        int *slot_obs_g = slot_obs.data(), *lmk_ptr_g = lmk_ptr.data();
        auto run_part = [&, slot_obs_g, lmk_ptr_g](int t) {
            Part pt; // worked on locally (vector bookkeeping included), published once at the end
            struct Publish {
                Part &dst, &src;
                ~Publish() { dst = std::move(src); }
            } publish{part[t], pt};
            pt.ob = cutp[t];
            pt.oe = cutp[t + 1];
            pt.used.assign(F, 0);
            int prev = pt.ob > 0 ? ol[pt.ob - 1] : -1;
            if (prev < -1 || prev >= L) { // the neighbour reports it; keep the writes below in range
                pt.bad = 1;
                return;
            }
            pt.l_lo = prev + 1;
            // pass 1: validation + CSR pointer
            unsigned bad = 0;
            for (int o = pt.ob; o < pt.oe; o++) {
                const int l = ol[o], f = of[o], c = oc[o];
                bad |= (unsigned)(l < 0) | (unsigned)(l >= L) | (unsigned)(f < 0) | (unsigned)(f >= F) | (unsigned)(c < 0) | (unsigned)(c >= C) | (unsigned)(l < prev);
                if (bad) break;
                if (l != prev) {
                    for (int q = prev + 1; q <= l; q++) lmk_ptr_g[q] = o;
                    prev = l;
                }
                pt.used[f] = 1;
            }
            if (bad) {
                pt.bad = 1;
                return;
            }
            pt.l_hi = prev;
            // pass 2: slots of landmarks l_lo .. l_hi (distinct keyframes in first-appearance order)
            const int nl = pt.l_hi - pt.l_lo + 1;
            pt.cnt.assign(std::max(nl, 0), 0);
            pt.same.assign(std::max(nl, 0), 0);
            pt.sf.clear();
            pt.sop.clear();
            pt.sf.reserve((size_t)(pt.oe - pt.ob));
            pt.sop.reserve((size_t)(pt.oe - pt.ob));
            int nso_l = pt.ob, prev_first = 0, prev_m = -1;
            for (int l = pt.l_lo; l <= pt.l_hi; l++) {
                const int a0 = lmk_ptr_g[l], b0 = l < pt.l_hi ? lmk_ptr_g[l + 1] : pt.oe;
                const int first = (int)pt.sf.size();
                bool grouped = true;
                {
                    const int nso_save = nso_l;
                    int cur = -1;
                    for (int o = a0; o < b0 && grouped; o++) {
                        const int f = of[o];
                        if (f != cur) {
                            for (int q = first; q < (int)pt.sf.size(); q++) grouped &= pt.sf[q] != f;
                            pt.sf.push_back(f);
                            pt.sop.push_back(nso_l);
                            cur = f;
                        }
                        slot_obs_g[nso_l++] = o;
                    }
                    if (!grouped) {
                        pt.sf.resize(first);
                        pt.sop.resize(first);
                        nso_l = nso_save;
                    }
                }
                if (!grouped) {
                    for (int o = a0; o < b0; o++) {
                        const int f = of[o];
                        bool seen = false;
                        for (int q = first; q < (int)pt.sf.size(); q++) seen |= pt.sf[q] == f;
                        if (!seen) pt.sf.push_back(f);
                    }
                    for (int q = first; q < (int)pt.sf.size(); q++) {
                        pt.sop.push_back(nso_l);
                        const int f = pt.sf[q];
                        for (int o = a0; o < b0; o++)
                            if (of[o] == f) slot_obs_g[nso_l++] = o;
                    }
                }
                const int m = (int)pt.sf.size() - first;
                pt.cnt[l - pt.l_lo] = m;
                pt.max_slots = std::max(pt.max_slots, m);
                bool same = prev_m == m && l > pt.l_lo;
                for (int q = 0; same && q < m; q++) same = pt.sf[first + q] == pt.sf[prev_first + q];
                pt.same[l - pt.l_lo] = same;
                prev_first = first;
                prev_m = m;
            }
        };
        for (int t = 1; t < NPART; t++) h->pool.submit(g_struct, [&run_part, t] { run_part(t); });
        run_part(0);
        h->pool.wait(g_struct);
        int any_bad = 0, tot_slots = 0, off[MAXPART + 1];
        for (int t = 0; t < NPART; t++) {
            any_bad |= part[t].bad;
            off[t] = tot_slots;
            tot_slots += (int)part[t].sf.size();
            max_slots = std::max(max_slots, part[t].max_slots);
        }
        off[NPART] = tot_slots;
        if (any_bad) return fail(h, SDV_ERR_INVALID_ARGUMENT, "observation index out of range or observations not landmark-major (reference walk order)");
        if (max_slots > MAX_SLOTS) return fail(h, SDV_ERR_UNSUPPORTED, "a landmark is observed from more than 32 keyframes (kernel limit of this build)");
        {
            const int last = part[NPART - 1].l_hi; // (every part passed: l_hi is the last landmark seen so far)
            int lastl = -1;
            for (int t = 0; t < NPART; t++) lastl = std::max(lastl, part[t].l_hi);
            (void)last;
            for (int q = lastl + 1; q <= L; q++) lmk_ptr[q] = O;
        }
        for (int t = 0; t < NPART; t++)
            for (int f = 0; f < F; f++) pose_used[f] |= part[t].used[f];
        ns = tot_slots;
        nso = O;
        slot_ptr.resize((size_t)L + 1);
        slot_frame.resize((size_t)O + 1);
        slot_obs_ptr.resize((size_t)O + 2);
        int *slot_ptr_g = slot_ptr.data(), *slot_frame_g = slot_frame.data(), *slot_obs_ptr_g = slot_obs_ptr.data();
        char *same_g = same_prev.data();
        auto place_part = [&, slot_ptr_g, slot_frame_g, slot_obs_ptr_g, same_g](int t) {
            const Part &pt = part[t];
            int sidx = off[t];
            for (int l = pt.l_lo; l <= pt.l_hi; l++) {
                slot_ptr_g[l] = sidx;
                sidx += pt.cnt[l - pt.l_lo];
                same_g[l] = pt.same[l - pt.l_lo];
            }
            if (!pt.sf.empty()) {
                std::memcpy(slot_frame_g + off[t], pt.sf.data(), 4 * pt.sf.size());
                std::memcpy(slot_obs_ptr_g + off[t], pt.sop.data(), 4 * pt.sop.size());
            }
        };
        for (int t = 1; t < NPART; t++) h->pool.submit(g_struct, [&place_part, t] { place_part(t); });
        place_part(0);
        h->pool.wait(g_struct);
        {
            int lastl = -1;
            for (int t = 0; t < NPART; t++) lastl = std::max(lastl, part[t].l_hi);
            for (int q = lastl + 1; q <= L; q++) {
                slot_ptr[q] = tot_slots;
                if (q < L) same_prev[q] = 0;
            }
        }
        slot_obs_ptr[ns] = nso;
    } else if (!gpu_struct)
    // one pass over the observations: range checks, landmark-major order, CSR pointer per landmark, frames in use
    {
        const int32_t *ol = w->obs_lmk, *of = w->obs_frame, *oc = w->obs_cam;
        int prev = -1;
        unsigned bad = 0;
        for (int o = 0; o < O; o++) {
            const int l = ol[o], f = of[o], c = oc[o];
            bad |= (unsigned)(l < 0) | (unsigned)(l >= L) | (unsigned)(f < 0) | (unsigned)(f >= F) | (unsigned)(c < 0) | (unsigned)(c >= C) | (unsigned)(l < prev);
            if (bad) break;
            if (l != prev) {
                for (int q = prev + 1; q <= l; q++) lmk_ptr[q] = o;
                prev = l;
            }
            pose_used[f] = 1;
        }
        if (bad) return fail(h, SDV_ERR_INVALID_ARGUMENT, "observation index out of range or observations not landmark-major (reference walk order)");
        for (int q = prev + 1; q <= L; q++) lmk_ptr[q] = O;
    }
    for (int p = 0; p < Pn; p++)
        if (w->imu_i[p] < 0 || w->imu_i[p] >= F || w->imu_j[p] < 0 || w->imu_j[p] >= F || w->imu_i[p] == w->imu_j[p])
            return fail(h, SDV_ERR_INVALID_ARGUMENT, "imu pair index out of range");
    const sdv_dense_prior *dp = w->dense_prior;
    if (dp) {
        if (dp->n_full <= 0 || dp->n <= 0 || !dp->J || !dp->r0 || dp->frame >= F || (dp->n_keep > 0 && (!dp->keep_lmk || !dp->keep_col)))
            return fail(h, SDV_ERR_INVALID_ARGUMENT, "malformed dense prior");
        if (dp->frame >= 0 && (dp->frame_col < 0 || dp->frame_col + 15 > dp->n)) return fail(h, SDV_ERR_INVALID_ARGUMENT, "dense prior frame_col");
        for (int k = 0; k < dp->n_keep; k++)
            if (dp->keep_lmk[k] < 0 || dp->keep_lmk[k] >= L || dp->keep_col[k] + 3 > dp->n)
                return fail(h, SDV_ERR_INVALID_ARGUMENT, "dense prior landmark index/column out of range");
    }
    cudaSetDevice(h->device);
    const bool timing = getenv("SDV_TIMING") != nullptr;
    auto tu0 = std::chrono::steady_clock::now();

    // ---- reduced-program structure (what Ceres' preprocessor derives: constant blocks dropped, unused blocks dropped)
    for (int p = 0; p < Pn; p++) {
        pose_used[w->imu_i[p]] = pose_used[w->imu_j[p]] = 1;
        vb_used[w->imu_i[p]] = vb_used[w->imu_j[p]] = 1;
    }
    if (w->has_prior)
        for (int f = 0; f < F; f++)
            if (w->has_prior[f]) pose_used[f] = 1;
    if (dp && dp->frame >= 0) {
        pose_used[dp->frame] = 1;
        if (w->vio) vb_used[dp->frame] = 1;
    }
    if (sp && sp->has_imu_prior) {
        pose_used[sp->frame] = 1;
        vb_used[sp->frame] = 1;
    }
    std::vector<int> pose_col(F, -1), vb_col(F, -1), lmk_col(L, -1);
    int n = 0;
    for (int f = 0; f < F; f++) {
        bool fixed = f > (F - w->n_fixed - 1); // AngularAdjustmentCERESAnalytic.cpp:234, AOptimizer.cpp:47
        if (!fixed && pose_used[f]) {
            pose_col[f] = n;
            n += 6;
        }
        if (!fixed && w->vio && vb_used[f]) {
            vb_col[f] = n;
            n += 9;
        }
    }
    if (dp)
        for (int k = 0; k < dp->n_keep; k++) {
            if (dp->keep_col[k] < 0) continue;
            int l = dp->keep_lmk[k];
            if (lmk_col[l] < 0) {
                lmk_col[l] = n;
                n += 3;
            }
        }
    if (sp) {
        auto make_dense = [&](int l) {
            if (lmk_col[l] < 0) {
                lmk_col[l] = n;
                n += 3;
            }
        };
        if (sp->has_lmk_prior) make_dense(sp->lmk0); // moved to elimination group 2, …Analytic.cpp:437-438
        for (int k = 0; k < sp->n_l2l; k++) {         // chain factors couple landmarks, …Analytic.cpp:448-481
            make_dense(sp->l2l_a[k]);
            make_dense(sp->l2l_b[k]);
        }
    }
    // n == 0: every keyframe constant — landmarkOptimization (AOptimizer.cpp:98-150): the reduced system is one block of padding
    // (identity), every landmark is a 3x3 problem and they all share ONE trust region
    if (n == 0 && (L == 0 || w->landmarks_constant)) return fail(h, SDV_ERR_INVALID_ARGUMENT, "window has no free parameter block");
    if (w->landmarks_constant && (dp || sp)) return fail(h, SDV_ERR_UNSUPPORTED, "landmarks_constant with a marginalisation prior (the single-frame solves carry none)");
    if (w->visual_loss_huber_a < 0.0 || w->max_num_iterations < 0) return fail(h, SDV_ERR_INVALID_ARGUMENT, "negative Huber parameter or iteration cap");
    const int n_pad = std::max(32, (n + 31) / 32 * 32), ld = n_pad;

    auto t_s1 = std::chrono::steady_clock::now();
    // ---- landmark shard of this rank (contiguous, balanced by observation count)
    int l0 = 0, l1 = L, o0 = 0, o1 = O;
    if (h->world > 1) {
        auto cut = [&](int r) {
            long long target = (long long)O * r / h->world;
            int l = (int)(std::lower_bound(lmk_ptr.begin(), lmk_ptr.end(), (int)target) - lmk_ptr.begin());
            return std::min(l, L);
        };
        l0 = h->rank == 0 ? 0 : cut(h->rank);         // identical to sdv_shard_range
        l1 = h->rank == h->world - 1 ? L : cut(h->rank + 1);
        o0 = lmk_ptr[l0];
        o1 = lmk_ptr[l1];
    }
    const int Oloc = o1 - o0;

    // ---- slots: (landmark, distinct keyframe) groups over the real observations and the PoseToLandmark pseudo-observations;
    //      slot_obs holds plane indices local to this rank (real observation o -> o - o0, pseudo-observations after them)
    const int np2l = sp ? sp->n_p2l : 0;
    std::vector<int> p2l_plane(std::max(np2l, 1), -1);
    std::vector<std::vector<int>> p2l_of_lmk;
    int n_pseudo = 0;
    if (np2l) {
        p2l_of_lmk.resize(L);
        for (int k = 0; k < np2l; k++) {
            int l = sp->p2l_lmk[k];
            if (l >= l0 && l < l1) {
                p2l_plane[k] = Oloc + 2 * n_pseudo;
                n_pseudo++;
                p2l_of_lmk[l].push_back(k);
            }
        }
    }
    const int Ocap = Oloc + 2 * n_pseudo;
    int prev_first = 0;
    if (!par_struct) {
        slot_ptr.resize((size_t)L + 1);
        slot_frame.resize((size_t)O + 2 * (size_t)n_pseudo + 1);
        slot_obs_ptr.resize((size_t)O + 2 * (size_t)n_pseudo + 2);
        slot_obs.resize((size_t)O + 2 * (size_t)n_pseudo + 1);
        same_prev.resize((size_t)L + 1);
    }
    if (!par_struct) {
        const int32_t *of = w->obs_frame;
        std::vector<int> ef, ep; // (frame, plane index) of the landmark's entries, only used when pseudo-observations exist
        for (int l = 0; l < L; l++) {
            slot_ptr[l] = ns;
            const int a0 = lmk_ptr[l], b0 = lmk_ptr[l + 1];
            const int first_slot = ns;
            if (!np2l || p2l_of_lmk[l].empty()) {
                // fast path: observations of one keyframe are normally adjacent (stereo pairs), so one forward pass builds the
                // slots; a frame that re-appears after another one falls back to the general grouping below
                bool grouped = true;
                {
                    const int ns_save = ns, nso_save = nso;
                    int cur = -1;
                    for (int o = a0; o < b0 && grouped; o++) {
                        const int f = of[o];
                        if (f != cur) {
                            for (int q = first_slot; q < ns; q++) grouped &= slot_frame[q] != f;
                            slot_frame[ns] = f;
                            slot_obs_ptr[ns] = nso;
                            ns++;
                            cur = f;
                        }
                        slot_obs[nso++] = o - o0;
                    }
                    if (!grouped) {
                        ns = ns_save;
                        nso = nso_save;
                    }
                }
                if (!grouped) {
                    // distinct frames in first-appearance order; per-landmark observation counts are small: linear scans
                    for (int o = a0; o < b0; o++) {
                        const int f = of[o];
                        int sidx = -1;
                        for (int q = first_slot; q < ns; q++)
                            if (slot_frame[q] == f) {
                                sidx = q;
                                break;
                            }
                        if (sidx < 0) slot_frame[ns++] = f;
                    }
                    for (int q = first_slot; q < ns; q++) {
                        slot_obs_ptr[q] = nso;
                        const int f = slot_frame[q];
                        for (int o = a0; o < b0; o++)
                            if (of[o] == f) slot_obs[nso++] = o - o0;
                    }
                }
            } else {
                ef.clear();
                ep.clear();
                for (int o = a0; o < b0; o++) {
                    ef.push_back(of[o]);
                    ep.push_back(o - o0);
                }
                for (int k : p2l_of_lmk[l]) {
                    ef.push_back(sp->frame);
                    ep.push_back(p2l_plane[k]);
                    ef.push_back(sp->frame);
                    ep.push_back(p2l_plane[k] + 1);
                }
                for (size_t e = 0; e < ef.size(); e++) {
                    bool seen = false;
                    for (int q = first_slot; q < ns; q++) seen |= slot_frame[q] == ef[e];
                    if (!seen) slot_frame[ns++] = ef[e];
                }
                for (int q = first_slot; q < ns; q++) {
                    slot_obs_ptr[q] = nso;
                    for (size_t e = 0; e < ef.size(); e++)
                        if (ef[e] == slot_frame[q]) slot_obs[nso++] = ep[e];
                }
            }
            if (ns - first_slot > MAX_SLOTS)
                return fail(h, SDV_ERR_UNSUPPORTED, "a landmark is observed from more than 32 keyframes (kernel limit of this build)");
            // same keyframes, same slot order as the previous landmark? (the structure passes below skip such landmarks)
            {
                const int m = ns - first_slot;
                max_slots = std::max(max_slots, m);
                bool same = l > 0 && m == first_slot - prev_first;
                for (int q = 0; same && q < m; q++) same = slot_frame[first_slot + q] == slot_frame[prev_first + q];
                same_prev[l] = same;
                prev_first = first_slot;
            }
        }
        slot_ptr[L] = ns;
        slot_obs_ptr[ns] = nso;
    }
    const int nslots = ns;
    const int nslotobs = nso;
    auto t_s2 = std::chrono::steady_clock::now();
    // ---- tiles of the fused kernels: consecutive landmarks of this rank, at most FT slots and FT_LMK landmarks each
    std::vector<int> &tile_ptr = h->tmp_tile_ptr;
    tile_ptr.clear();
    int tile_capq = 1;
    {
        // tile capacity: the fused kernels run 2 CTAs per SM; when the whole rank fits in ONE wave of tiles of at most FT slots the
        // tiles are sized for exactly that (a second, nearly empty wave would double the kernel time: everything here is
        // latency-bound), otherwise full tiles
        const int nsl_rank = gpu_struct ? nslots : slot_ptr[l1] - slot_ptr[l0], waves1 = h->num_sms * FUSED_CPS;
        int cap = FT;
        if (const char *e = getenv("SDV_FUSED_TILE_SLOTS")) cap = std::max(1, std::min(FT, atoi(e))); // tests: the result must not depend on the tiling
        else if (nsl_rank <= (long long)waves1 * (FT - max_slots)) cap = std::max(32, (nsl_rank + waves1 - 1) / waves1 + max_slots);
        cap = std::min(cap, FT);
        tile_capq = std::max(1, cap - max_slots + 1); // device tiling: buckets of this many slots (sdv_struct.cuh)
        int l = l0;
        while (!gpu_struct && l < l1) {
            tile_ptr.push_back(l);
            int nsl = 0, nlm = 0;
            while (l < l1 && nlm < FT_LMK && nsl + (slot_ptr[l + 1] - slot_ptr[l]) <= (nlm == 0 ? FT : cap)) {
                nsl += slot_ptr[l + 1] - slot_ptr[l];
                nlm++;
                l++;
            }
        }
        tile_ptr.push_back(l1);
    }
    // (device tiling: an upper bound — one tile per bucket, plus the cuts every FT_LMK landmarks; only used to size grids and buffers)
    const int ntiles = gpu_struct ? (nslots > 0 ? nslots / tile_capq + (l1 - l0) / FT_LMK + 2 : 0) : (int)tile_ptr.size() - 1;

    auto t_s3 = std::chrono::steady_clock::now();
    // ---- dense prior column maps
    std::vector<int> mp_src, mp_dst;
    if (dp) {
        if (dp->frame >= 0) {
            for (int q = 0; q < 15; q++) {
                mp_src.push_back(dp->frame_col + q);
                int dst = -1;
                if (q < 6) dst = pose_col[dp->frame] >= 0 ? pose_col[dp->frame] + q : -1;
                else dst = vb_col[dp->frame] >= 0 ? vb_col[dp->frame] + (q - 6) : -1;
                mp_dst.push_back(dst);
            }
        }
        for (int k = 0; k < dp->n_keep; k++) {
            if (dp->keep_col[k] < 0) continue;
            for (int q = 0; q < 3; q++) {
                mp_src.push_back(dp->keep_col[k] + q);
                mp_dst.push_back(lmk_col[dp->keep_lmk[k]] + q);
            }
        }
    }
    const int nm = (int)mp_src.size();

    // ---- tile-level structure of the reduced system (32-column tiles) and its symbolic Cholesky fill.  A window whose
    //      landmarks are seen from a few neighbouring keyframes gives a block-banded reduced system (plus the rows of the kept
    //      landmarks of a prior); the factorisation kernel skips the tiles that are structurally zero, exactly like the sparse
    //      solver of the reference does (AOptimizer.cpp:383, SUITE_SPARSE).  A dense window costs what it did before.
    const int Tt = n_pad / 32;
    std::vector<uint32_t> &tile_nz = h->tmp_tile_nz; // [(Tt + 1)][4] bit j of row i: tile (i, j) of L may be non-zero
    tile_nz.assign((size_t)(Tt + 1) * 4, 0u);
    struct TMask { // up to 256 column groups
        uint64_t w[4];
        bool operator==(const TMask &o) const { return w[0] == o.w[0] && w[1] == o.w[1] && w[2] == o.w[2] && w[3] == o.w[3]; }
        void operator|=(const TMask &o) {
            w[0] |= o.w[0];
            w[1] |= o.w[1];
            w[2] |= o.w[2];
            w[3] |= o.w[3];
        }
    };
    // low[i]: column groups j (of `gs` columns each, at most 128 groups) coupled with group row i before elimination
    auto build_low = [&](const int gs, std::vector<TMask> &low) {
        const int ng = n_pad / gs;
        auto add_cols = [&](TMask &m, int c0, int ncols) {
            if (c0 < 0) return;
            for (int t = c0 / gs; t <= (c0 + ncols - 1) / gs; t++) m.w[t >> 6] |= 1ull << (t & 63);
        };
        low.assign((size_t)ng + 1, TMask{{0, 0, 0, 0}});
        auto add_clique = [&](const TMask &m) {
            for (int wq = 0; wq < 4; wq++) {
                uint64_t bits = m.w[wq];
                while (bits) {
                    const int i = wq * 64 + __builtin_ctzll(bits);
                    bits &= bits - 1;
                    low[i] |= m;
                }
            }
        };
        if (getenv("SDV_CHOL_DENSE")) {
            TMask all{{0, 0, 0, 0}};
            add_cols(all, 0, n_pad);
            add_clique(all);
            return;
        }
        std::vector<TMask> pose_mask(F, TMask{{0, 0, 0, 0}}), frame_mask(F, TMask{{0, 0, 0, 0}});
        for (int f = 0; f < F; f++) {
            add_cols(pose_mask[f], pose_col[f], 6);
            frame_mask[f] = pose_mask[f];
            add_cols(frame_mask[f], vb_col[f], 9);
            add_clique(frame_mask[f]); // diagonal blocks (pose prior, damping)
        }
        // visual factors: every landmark couples the poses of the keyframes that see it (and its own columns when kept)
        std::vector<TMask> uniq;
        TMask last{{0, 0, 0, 0}};
        for (int l = 0; l < L; l++) {
            if (same_prev[l] && lmk_col[l] < 0 && lmk_col[l - 1] < 0) continue; // same clique as the previous landmark
            TMask m{{0, 0, 0, 0}};
            for (int q = slot_ptr[l]; q < slot_ptr[l + 1]; q++) {
                const TMask &pm = pose_mask[slot_frame[q]];
                m |= pm;
            }
            if (lmk_col[l] >= 0) add_cols(m, lmk_col[l], 3);
            if (m == last) continue;
            last = m;
            bool seen = false;
            for (size_t u = uniq.size(); u-- > 0 && uniq.size() - u <= 32;) seen |= uniq[u] == m;
            if (!seen) uniq.push_back(m);
        }
        for (const TMask &m : uniq) add_clique(m);
        for (int p = 0; p < Pn; p++) { // IMUFactor + IMUBiasFactor couple all 15 parameters of both keyframes
            TMask m = frame_mask[w->imu_i[p]];
            m |= frame_mask[w->imu_j[p]];
            add_clique(m);
        }
        if (dp) { // dense marginalisation prior: one clique over everything it touches
            TMask m{{0, 0, 0, 0}};
            for (int c : mp_dst)
                if (c >= 0) add_cols(m, c, 1);
            add_clique(m);
        }
        if (sp) {
            if (sp->has_imu_prior) add_clique(frame_mask[sp->frame]);
            if (sp->has_lmk_prior) {
                TMask m{{0, 0, 0, 0}};
                add_cols(m, lmk_col[sp->lmk0], 3);
                add_clique(m);
            }
            for (int k = 0; k < sp->n_l2l; k++) {
                TMask m{{0, 0, 0, 0}};
                add_cols(m, lmk_col[sp->l2l_a[k]], 3);
                add_cols(m, lmk_col[sp->l2l_b[k]], 3);
                add_clique(m);
            }
        }
        for (int t = 0; t < ng; t++) low[t].w[t >> 6] |= 1ull << (t & 63); // padding columns: identity diagonal
    };
    auto bit = [&](const TMask &m, int j) { return (m.w[j >> 6] >> (j & 63)) & 1ull; };
    // ---- block half-bandwidth of the reduced system at 16-column granularity.  The envelope of a Cholesky factor is the
    //      envelope of the matrix, so max_i (i - first coupled block of row i) bounds the fill: when that band (plus two
    //      look-ahead block rows) fits in the shared memory of one SM, the whole factorisation runs in ONE CTA (k_chol_band).
    int band_bw = -1;
    if (gpu_struct) {
        // every factor couples a contiguous column range [cmin, cmax]: the half-bandwidth is the widest range in 16-column blocks
        band_bw = 0;
        auto clique = [&](int cmin, int cmax) {
            if (cmin >= 0 && cmax >= cmin) band_bw = std::max(band_bw, cmax / 16 - cmin / 16);
        };
        auto frame_lo = [&](int f) { return pose_col[f] >= 0 ? pose_col[f] : vb_col[f]; };
        auto frame_hi = [&](int f) { return vb_col[f] >= 0 ? vb_col[f] + 8 : (pose_col[f] >= 0 ? pose_col[f] + 5 : -1); };
        for (int f = 0; f < F; f++) {
            clique(frame_lo(f), frame_hi(f));                                   // diagonal blocks (pose prior, damping)
            if (span_max[f] >= 0) clique(pose_col[f], pose_col[span_max[f]] + 5); // landmarks couple the POSES of the keyframes that see them
        }
        for (int p = 0; p < Pn; p++) { // IMUFactor + IMUBiasFactor couple all 15 parameters of both keyframes
            const int a = w->imu_i[p], b = w->imu_j[p];
            int lo = -1, hi = -1;
            for (int f : {a, b}) {
                if (frame_lo(f) >= 0) lo = lo < 0 ? frame_lo(f) : std::min(lo, frame_lo(f));
                hi = std::max(hi, frame_hi(f));
            }
            clique(lo, hi);
        }
        // (the sparsified VIO prior without PoseToLandmark factors is one more diagonal clique of its keyframe)
    } else if (n_pad / 16 <= 256) {
        std::vector<TMask> low16;
        build_low(16, low16);
        const int nb16 = n_pad / 16;
        band_bw = 0;
        for (int i = 0; i < nb16; i++) {
            int first = i;
            for (int wq = 3; wq >= 0; wq--)
                if (low16[i].w[wq]) first = wq * 64 + __builtin_ctzll(low16[i].w[wq]);
            band_bw = std::max(band_bw, i - std::min(first, i));
        }
    }
    h->band_bw = band_bw;
    // (the 32-column tile pattern only serves the wide-band fallback k_chol_chain: skipped whenever k_chol_band will run)
    bool band_applies = false;
    if (band_bw >= 0 && std::max(band_bw, 1) <= BAND_MAX_BW && !getenv("SDV_CHOL_VARIANT") && !getenv("SDV_CHOL_DENSE")) {
        const BandPlan plb = band_plan(n_pad, std::max(band_bw, 1));
        band_applies = sizeof(double) * (size_t)plb.o_end <= 220 * 1024 && (size_t)plb.nb * (std::max(band_bw, 1) + 2) * 256 <= (size_t)(n_pad + 32) * ld;
    }
    if (gpu_struct && !band_applies) { // wide band: the cluster Cholesky wants the tile pattern, which is built from the host's slot lists
        h->force_host_slots = true;
        const int rc2 = upload_impl(h, w, sync);
        h->force_host_slots = false;
        return rc2;
    }
    if (Tt <= 128 && !band_applies) {
        std::vector<TMask> low;
        build_low(32, low);
        // symbolic right-looking elimination on the tile graph: the rows below pivot k become mutually coupled
        for (int k = 0; k < Tt; k++) {
            TMask col{{0, 0, 0, 0}}; // rows i > k with (i, k) non-zero
            for (int i = k + 1; i < Tt; i++)
                if (bit(low[i], k)) col.w[i >> 6] |= 1ull << (i & 63);
            for (int i = k + 1; i < Tt; i++)
                if (bit(col, i)) {
                    low[i] |= col;
                }
        }
        int nnz_tiles = 0;
        for (int i = 0; i < Tt; i++)
            for (int j = 0; j <= i; j++)
                if (bit(low[i], j)) {
                    tile_nz[(size_t)i * 4 + (j >> 5)] |= 1u << (j & 31);
                    nnz_tiles++;
                }
        for (int j = 0; j < Tt; j++) tile_nz[(size_t)Tt * 4 + (j >> 5)] |= 1u << (j & 31); // right-hand-side row: dense
        h->chol_tiles_nz = nnz_tiles;
        h->chol_tiles_all = Tt * (Tt + 1) / 2;
    }

    // ---- input arena
    Arena A;
    auto D = sizeof(double);
    const size_t o_P = A.add(sizeof(DevProblem)); // offset 0: the device copy of the problem description travels with the arena
    size_t o_T = A.add(D * 12 * F), o_v = A.add(D * 3 * F), o_ba = A.add(D * 3 * F), o_bg = A.add(D * 3 * F);
    size_t o_hp = A.add(F), o_Tp = A.add(D * 12 * F), o_ip = A.add(D * 6 * F);
    size_t o_pc = A.add(4 * F), o_vc = A.add(4 * F);
    size_t o_Ts = A.add(D * 12 * C), o_K = A.add(D * 4 * C), o_cw = A.add(D * C);
    size_t o_lc = A.add(4 * std::max(L, 1));
    size_t o_tnz = A.add(4 * tile_nz.size());
    size_t o_tile = 0, o_sp = 0, o_sf = 0, o_sop = 0, o_so = 0;
    if (!gpu_struct) {
        o_tile = A.add(4 * tile_ptr.size());
        o_sp = A.add(4 * (L + 1));
        o_sf = A.add(4 * std::max(nslots, 1));
        o_sop = A.add(4 * (nslots + 1));
        o_so = A.add(4 * std::max(nslotobs, 1));
    }
    size_t o_ii = A.add(4 * std::max(Pn, 1)), o_ij = A.add(4 * std::max(Pn, 1));
    size_t o_idt = A.add(D * std::max(Pn, 1)), o_idR = A.add(D * 9 * std::max(Pn, 1)), o_idv = A.add(D * 3 * std::max(Pn, 1)),
           o_idp = A.add(D * 3 * std::max(Pn, 1)), o_icov = A.add(D * 81 * std::max(Pn, 1));
    size_t o_j1 = A.add(D * 9 * std::max(Pn, 1)), o_j2 = A.add(D * 9 * std::max(Pn, 1)), o_j3 = A.add(D * 9 * std::max(Pn, 1)),
           o_j4 = A.add(D * 9 * std::max(Pn, 1)), o_j5 = A.add(D * 9 * std::max(Pn, 1));
    size_t o_sba = A.add(D * std::max(Pn, 1)), o_sbg = A.add(D * std::max(Pn, 1));
    size_t o_mJ = 0, o_mr = 0, o_ms = 0, o_md = 0;
    if (dp) {
        o_mJ = A.add(D * (size_t)dp->n_full * dp->n);
        o_mr = A.add(D * dp->n_full);
        o_ms = A.add(4 * std::max(nm, 1));
        o_md = A.add(4 * std::max(nm, 1));
    }
    size_t o_spb = 0, o_p2l = 0, o_p2p = 0, o_p2d = 0, o_p2s = 0, o_l2a = 0, o_l2b = 0, o_l2d = 0, o_l2s = 0;
    const int nl2l = sp ? sp->n_l2l : 0;
    if (sp) {
        o_spb = A.add(D * 258);
        o_p2l = A.add(4 * std::max(np2l, 1));
        o_p2p = A.add(4 * std::max(np2l, 1));
        o_p2d = A.add(D * 3 * std::max(np2l, 1));
        o_p2s = A.add(D * 9 * std::max(np2l, 1));
        o_l2a = A.add(4 * std::max(nl2l, 1));
        o_l2b = A.add(4 * std::max(nl2l, 1));
        o_l2d = A.add(D * 3 * std::max(nl2l, 1));
        o_l2s = A.add(D * 9 * std::max(nl2l, 1));
    }
    int rc;
    if ((rc = ensure(h, &h->h_in, &h->in_cap, A.size, true, getenv("SDV_ARENA_WC") != nullptr)) != SDV_OK) return rc; // write-combined measured 15 % slower to fill
    {
        size_t dcap = h->in_bytes;
        unsigned char *dptr = h->d_in;
        if (A.size > dcap) {
            if (dptr) cudaFree(dptr);
            CK(cudaMalloc((void **)&dptr, A.size * 2));
            h->d_in = dptr;
            h->in_bytes = A.size * 2;
        }
    }
    unsigned char *hb = h->h_in;
    auto t_pack0 = std::chrono::steady_clock::now();
    std::memcpy(hb + o_T, w->T_f_w, D * 12 * F);
    if (w->vio) {
        std::memcpy(hb + o_v, w->v, D * 3 * F);
        std::memcpy(hb + o_ba, w->ba, D * 3 * F);
        std::memcpy(hb + o_bg, w->bg, D * 3 * F);
    } else {
        std::memset(hb + o_v, 0, D * 3 * F);
        std::memset(hb + o_ba, 0, D * 3 * F);
        std::memset(hb + o_bg, 0, D * 3 * F);
    }
    if (w->has_prior) {
        std::memcpy(hb + o_hp, w->has_prior, F);
        std::memcpy(hb + o_Tp, w->T_prior, D * 12 * F);
        std::memcpy(hb + o_ip, w->inf_prior, D * 6 * F);
    } else {
        std::memset(hb + o_hp, 0, F);
    }
    std::memcpy(hb + o_pc, pose_col.data(), 4 * F);
    std::memcpy(hb + o_tnz, tile_nz.data(), 4 * tile_nz.size());
    if (!gpu_struct) std::memcpy(hb + o_tile, tile_ptr.data(), 4 * tile_ptr.size());
    std::memcpy(hb + o_vc, vb_col.data(), 4 * F);
    std::memcpy(hb + o_Ts, w->T_s_f, D * 12 * C);
    std::memcpy(hb + o_K, w->K, D * 4 * C);
    for (int c = 0; c < C; c++) {
        double focal = (w->K[4 * c] + w->K[4 * c + 1]) / 2; // Camera.h:46
        double sigma = kind == SDV_FACTOR_ANGULAR ? 1.5 / focal : 1.0; // …Analytic.cpp:283 ; BA…Analytic.h:47
        at<double>(hb, o_cw)[c] = 1.0 / sigma;
    }
    if (L > 0) std::memcpy(hb + o_lc, lmk_col.data(), 4 * L);
    if (!gpu_struct) {
        std::memcpy(hb + o_sp, slot_ptr.data(), 4 * (L + 1));
        if (nslots) std::memcpy(hb + o_sf, slot_frame.data(), 4 * nslots);
        std::memcpy(hb + o_sop, slot_obs_ptr.data(), 4 * (nslots + 1));
        if (nslotobs) std::memcpy(hb + o_so, slot_obs.data(), 4 * (size_t)nslotobs);
    }
    // (observation indices, measurements — array-of-structs as the caller provides them — and landmarks: bulk data arena)
    if (Pn) {
        std::memcpy(hb + o_ii, w->imu_i, 4 * Pn);
        std::memcpy(hb + o_ij, w->imu_j, 4 * Pn);
        std::memcpy(hb + o_idt, w->imu_dt, D * Pn);
        std::memcpy(hb + o_idR, w->imu_dR, D * 9 * Pn);
        std::memcpy(hb + o_idv, w->imu_dv, D * 3 * Pn);
        std::memcpy(hb + o_idp, w->imu_dp, D * 3 * Pn);
        std::memcpy(hb + o_icov, w->imu_cov, D * 81 * Pn);
        std::memcpy(hb + o_j1, w->imu_J_dR_bg, D * 9 * Pn);
        std::memcpy(hb + o_j2, w->imu_J_dv_ba, D * 9 * Pn);
        std::memcpy(hb + o_j3, w->imu_J_dv_bg, D * 9 * Pn);
        std::memcpy(hb + o_j4, w->imu_J_dp_ba, D * 9 * Pn);
        std::memcpy(hb + o_j5, w->imu_J_dp_bg, D * 9 * Pn);
        std::memcpy(hb + o_sba, w->imu_sigma_ba, D * Pn);
        std::memcpy(hb + o_sbg, w->imu_sigma_bg, D * Pn);
    }
    if (dp) {
        std::memcpy(hb + o_mJ, dp->J, D * (size_t)dp->n_full * dp->n);
        std::memcpy(hb + o_mr, dp->r0, D * dp->n_full);
        if (nm) {
            std::memcpy(hb + o_ms, mp_src.data(), 4 * nm);
            std::memcpy(hb + o_md, mp_dst.data(), 4 * nm);
        }
    }
    if (sp) {
        double *blob = at<double>(hb, o_spb);
        std::memset(blob, 0, D * 258);
        if (sp->has_imu_prior) {
            std::memcpy(blob, sp->T_prior, D * 12);
            std::memcpy(blob + 12, sp->v_prior, D * 3);
            std::memcpy(blob + 15, sp->ba_prior, D * 3);
            std::memcpy(blob + 18, sp->bg_prior, D * 3);
            std::memcpy(blob + 21, sp->imu_sqrt_inf, D * 225);
        }
        if (sp->has_lmk_prior) {
            std::memcpy(blob + 246, sp->lmk_prior, D * 3);
            std::memcpy(blob + 249, sp->lmk_sqrt_inf, D * 9);
        }
        if (np2l) {
            std::memcpy(hb + o_p2l, sp->p2l_lmk, 4 * np2l);
            std::memcpy(hb + o_p2p, p2l_plane.data(), 4 * np2l);
            std::memcpy(hb + o_p2d, sp->p2l_delta, D * 3 * np2l);
            std::memcpy(hb + o_p2s, sp->p2l_sqrt_inf, D * 9 * np2l);
        }
        if (nl2l) {
            std::memcpy(hb + o_l2a, sp->l2l_a, 4 * nl2l);
            std::memcpy(hb + o_l2b, sp->l2l_b, 4 * nl2l);
            std::memcpy(hb + o_l2d, sp->l2l_delta, D * 3 * nl2l);
            std::memcpy(hb + o_l2s, sp->l2l_sqrt_inf, D * 9 * nl2l);
        }
    }
    unsigned char *db = h->d_in;

    // ---- scratch arena (device only), laid out from capacities that only grow: the pointers captured in the CUDA graph stay
    //      valid while consecutive windows stay below them
    auto grow = [](int &cap, int need) {
        if (need > cap) cap = (need + need / 4 + 255) / 256 * 256;
    };
    grow(h->cap_O, std::max(Ocap, 1));
    grow(h->cap_L, std::max(L, 1));
    grow(h->cap_P, std::max(Pn, 1));
    grow(h->cap_nm, std::max(nm, 1));
    grow(h->cap_nfull, std::max(dp ? dp->n_full : 1, 1));
    grow(h->cap_l2l, std::max(nl2l, 1));
    const size_t cO = h->cap_O, cL = h->cap_L, cP = h->cap_P, cnm = h->cap_nm;
    Arena S;
    size_t s_st = S.add(sizeof(LMState)), s_acc = S.add(sizeof(Accum));
    size_t s_lin[2][14];
    for (int b = 0; b < 2; b++) {
        s_lin[b][0] = S.add(D * FCT_ROW * F * C);
        s_lin[b][1] = S.add(D * 2 * cO);
        s_lin[b][2] = S.add(D * 12 * cO);
        s_lin[b][3] = S.add(D * 6 * cO);
        s_lin[b][4] = S.add(D * 9 * cP);
        s_lin[b][5] = S.add(D * 216 * cP);
        s_lin[b][6] = S.add(D * 6 * cP);
        s_lin[b][7] = S.add(D * 6 * F);
        s_lin[b][8] = S.add(D * 36 * F);
        s_lin[b][9] = S.add(D * (size_t)h->cap_nfull);
        s_lin[b][10] = S.add(D * n_pad);
        s_lin[b][11] = S.add(D * 3 * cL);
        s_lin[b][12] = S.add(D * (18 + 3 * (size_t)h->cap_l2l));
        s_lin[b][13] = S.add(D * 225);
    }
    const size_t sb_elems = (size_t)(n_pad + 32) * ld;
    size_t s_Sb = S.add(D * sb_elems), s_Lo = S.add(D * sb_elems);
    size_t s_sp = S.add(D * n_pad), s_dp = S.add(D * n_pad), s_gp = S.add(D * n_pad), s_dx = S.add(D * n_pad);
    size_t s_sl = S.add(D * 3 * cL);
    size_t s_inf = S.add(D * 81 * cP);
    size_t s_mH = S.add(D * cnm * cnm), s_mg = S.add(D * cnm);
    size_t s_red = S.add(D * 16);
    size_t s_part = S.add(D * (size_t)(n_pad / 32 + 1) * CC_MAX * 32);
    size_t s_dinv = S.add(D * (n_pad + 32));
    size_t s_prof = S.add(D * 8 * CC_MAX);
    size_t s_aux = S.add(D * LMK_AUX * cL);
    // device structure pass (sdv_struct.cuh): scan scratch and the slot / tile lists themselves, sized from the capacities
    const size_t nblkO = cO / SCAN_BLOCK + 2, nblkL = cL / SCAN_BLOCK + 2;
    size_t s_head = S.add(4 * cO), s_sidx = S.add(4 * cO), s_ssum = S.add(4 * nblkO), s_thead = S.add(4 * cL), s_tidx = S.add(4 * cL), s_tsum = S.add(4 * nblkL),
           s_tot = S.add(64), s_gsp = S.add(4 * (cL + 2)), s_gsf = S.add(4 * (cO + 2)), s_gsop = S.add(4 * (cO + 2)), s_gso = S.add(4 * (cO + 2)), s_gtile = S.add(4 * (cL + 4));
    size_t s_xchg = h->world > 1 ? S.add(D * ((size_t)n_pad * n_pad + 3 * (size_t)n_pad + 8)) : 0; // upper bound (no band); the band case uses the head
    if ((rc = ensure(h, &h->d_scr, &h->scr_cap, S.size)) != SDV_OK) return rc;
    unsigned char *sb = h->d_scr;
    // solution buffer: [solution blocks | LMState | Accum] (k_gather_solution), one device-to-host copy per solve
    const size_t out_bytes = D * ((size_t)15 * F + 3 * cL) + sizeof(LMState) + sizeof(Accum) + 256;
    if ((rc = ensure(h, &h->d_out, &h->out_cap, out_bytes)) != SDV_OK) return rc;
    if ((rc = ensure(h, &h->h_rb, &h->rb_cap, sizeof(LMState) + sizeof(Accum) + 256, true)) != SDV_OK) return rc;
    if ((rc = ensure(h, &h->h_sol, &h->sol_cap, out_bytes, true)) != SDV_OK) return rc;

    DevProblem &P = h->P;
    std::memset(&P, 0, sizeof(P));
    P.F = F; P.C = C; P.L = L; P.O = O; P.P = Pn;
    P.vio = w->vio; P.kind = kind;
    P.n = n; P.n_pad = n_pad; P.ld = ld; P.nslots = nslots;
    P.band_bw = 0;
    P.l0 = l0; P.l1 = l1; P.o0 = o0; P.o1 = o1;
    P.rank = h->rank; P.world = h->world;
    P.T_f_w = at<double>(db, o_T); P.v = at<double>(db, o_v); P.ba = at<double>(db, o_ba); P.bg = at<double>(db, o_bg);
    P.has_prior = w->has_prior ? at<unsigned char>(db, o_hp) : nullptr;
    P.T_prior = at<double>(db, o_Tp); P.inf_prior = at<double>(db, o_ip);
    P.pose_col = at<int>(db, o_pc); P.vb_col = at<int>(db, o_vc);
    P.T_s_f = at<double>(db, o_Ts); P.K = at<double>(db, o_K); P.cam_w = at<double>(db, o_cw);
    P.lmk_t = at<double>(h->d_in2, q_lt); P.lmk_col = at<int>(db, o_lc);
    P.tile_nz = at<uint32_t>(db, o_tnz);
    P.ntiles = ntiles;
    P.st_on = gpu_struct ? 1 : 0;
    P.st_capq = tile_capq;
    P.st_head = at<int>(sb, s_head); P.st_sidx = at<int>(sb, s_sidx); P.st_ssum = at<int>(sb, s_ssum); P.st_thead = at<int>(sb, s_thead);
    P.st_tidx = at<int>(sb, s_tidx); P.st_tsum = at<int>(sb, s_tsum); P.st_tot = at<int>(sb, s_tot);
    if (gpu_struct) {
        P.tile_ptr = at<int>(sb, s_gtile);
        P.slot_ptr = at<int>(sb, s_gsp); P.slot_frame = at<int>(sb, s_gsf); P.slot_obs_ptr = at<int>(sb, s_gsop); P.slot_obs = at<int>(sb, s_gso);
    } else {
        P.tile_ptr = at<int>(db, o_tile);
        P.slot_ptr = at<int>(db, o_sp); P.slot_frame = at<int>(db, o_sf); P.slot_obs_ptr = at<int>(db, o_sop); P.slot_obs = at<int>(db, o_so);
    }
    P.obs_lmk = at<int>(h->d_in2, q_ol); P.obs_fc = at<int>(h->d_in2, q_ofc); P.obs_meas = at<double>(h->d_in2, q_om);
    P.obs_w = w->obs_sigma ? at<double>(h->d_in2, q_ow) : nullptr;
    P.imu_i = at<int>(db, o_ii); P.imu_j = at<int>(db, o_ij); P.imu_dt = at<double>(db, o_idt); P.imu_dR = at<double>(db, o_idR);
    P.imu_dv = at<double>(db, o_idv); P.imu_dp = at<double>(db, o_idp); P.imu_cov = at<double>(db, o_icov);
    P.imu_J_dR_bg = at<double>(db, o_j1); P.imu_J_dv_ba = at<double>(db, o_j2); P.imu_J_dv_bg = at<double>(db, o_j3);
    P.imu_J_dp_ba = at<double>(db, o_j4); P.imu_J_dp_bg = at<double>(db, o_j5);
    P.imu_sigma_ba = at<double>(db, o_sba); P.imu_sigma_bg = at<double>(db, o_sbg);
    P.imu_inf_sqrt = at<double>(sb, s_inf);
    if (dp) {
        P.mp_nfull = dp->n_full; P.mp_n = dp->n; P.mp_nmap = nm;
        P.mp_J = at<double>(db, o_mJ); P.mp_r0 = at<double>(db, o_mr);
        P.mp_src_col = at<int>(db, o_ms); P.mp_dst_col = at<int>(db, o_md);
        P.mp_H = at<double>(sb, s_mH); P.mp_g0 = at<double>(sb, s_mg);
    }
    P.Ocap = Ocap;
    P.lmk_const = w->landmarks_constant ? 1 : 0;
    P.max_iter = w->max_num_iterations;
    P.huber_a = w->visual_loss_huber_a;
    if (sp) {
        P.sp_has_imu = sp->has_imu_prior ? 1 : 0;
        P.sp_frame = sp->frame;
        P.sp_has_lmk = sp->has_lmk_prior ? 1 : 0;
        P.sp_lmk0 = sp->lmk0;
        P.sp_np2l = np2l;
        P.sp_nl2l = nl2l;
        P.sp_blob = at<double>(db, o_spb);
        P.sp_p2l_lmk = at<int>(db, o_p2l); P.sp_p2l_plane = at<int>(db, o_p2p);
        P.sp_p2l_delta = at<double>(db, o_p2d); P.sp_p2l_sqrt = at<double>(db, o_p2s);
        P.sp_l2l_a = at<int>(db, o_l2a); P.sp_l2l_b = at<int>(db, o_l2b);
        P.sp_l2l_delta = at<double>(db, o_l2d); P.sp_l2l_sqrt = at<double>(db, o_l2s);
    }
    for (int b = 0; b < 2; b++) {
        LinBuf &B = h->B[b];
        B.fct = at<double>(sb, s_lin[b][0]); B.r = at<double>(sb, s_lin[b][1]); B.Jp = at<double>(sb, s_lin[b][2]);
        B.Jl = at<double>(sb, s_lin[b][3]); B.imu_r = at<double>(sb, s_lin[b][4]); B.imu_J = at<double>(sb, s_lin[b][5]);
        B.bias_r = at<double>(sb, s_lin[b][6]); B.prior_r = at<double>(sb, s_lin[b][7]); B.prior_J = at<double>(sb, s_lin[b][8]);
        B.mp_r = at<double>(sb, s_lin[b][9]); B.xp = at<double>(sb, s_lin[b][10]); B.xl = at<double>(sb, s_lin[b][11]);
        B.sp_r = at<double>(sb, s_lin[b][12]); B.sp_J = at<double>(sb, s_lin[b][13]);
    }
    h->d_st = at<LMState>(sb, s_st);
    h->d_acc = at<Accum>(sb, s_acc);
    h->d_Sb = at<double>(sb, s_Sb); h->d_Lo = at<double>(sb, s_Lo);
    h->d_scale_p = at<double>(sb, s_sp); h->d_damp_p = at<double>(sb, s_dp); h->d_graw_p = at<double>(sb, s_gp); h->d_dxp = at<double>(sb, s_dx);
    h->d_scale_l = at<double>(sb, s_sl);
    h->d_red = at<double>(sb, s_red);
    h->d_partial = at<double>(sb, s_part);
    h->d_dinv = at<double>(sb, s_dinv);
    h->d_prof = at<double>(sb, s_prof);
    h->d_lmk_aux = at<double>(sb, s_aux);
    h->d_xchg = h->world > 1 ? at<double>(sb, s_xchg) : nullptr;
    h->sb_elems = sb_elems;

    // ---- launch geometry
    size_t fct_bytes = (size_t)F * C * FCT_ROW * D;
    P.fct_in_smem = fct_bytes <= 160 * 1024 ? 1 : 0;
    h->lin_smem = P.fct_in_smem ? (int)fct_bytes : 0;
    h->lin_fn = lin_visual_fn(kind, P.fct_in_smem != 0, getenv("SDV_LIN_LATE") == nullptr);
    // function attributes and occupancy answers do not change between windows of the same shape: asked once per handle
    if (!h->attrs_done) {
        CK(cudaFuncSetAttribute(k_trisolve, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(k_chol_band, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        CK(cudaFuncSetAttribute(k_lin_schur<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM_SCHUR));
        CK(cudaFuncSetAttribute(k_lin_schur<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM_SCHUR));
        h->attrs_done = true;
    }
    if ((const void *)h->lin_fn != h->lin_fn_cached || h->lin_smem != h->lin_smem_cached) {
        CK(cudaFuncSetAttribute(h->lin_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        int q = 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, h->lin_fn, LIN_THREADS, h->lin_smem));
        h->lin_per_sm = std::max(q, 1);
        h->lin_fn_cached = (const void *)h->lin_fn;
        h->lin_smem_cached = h->lin_smem;
    }
    const int per_sm = h->lin_per_sm;
    h->lin_grid = std::max(1, std::min((Oloc + LIN_THREADS - 1) / LIN_THREADS, h->num_sms * per_sm));
    h->fac_grid = std::max(1, (std::max(Pn, 1) + FAC_WARPS - 1) / FAC_WARPS);
    // persistent grids of exactly the resident CTAs (__launch_bounds__ of the two kernels): a larger grid runs its tail on part of the machine
    // (rounded up to a power of two below that, so that windows of similar size share one CUDA graph)
    auto pow2ceil = [](int v) {
        int p2 = 1;
        while (p2 < v) p2 <<= 1;
        return p2;
    };
    h->fused_grid = std::max(1, std::min(pow2ceil(ntiles), h->num_sms * FUSED_CPS));
    h->fused_grid_back = std::max(1, std::min(pow2ceil(ntiles), h->num_sms * FUSED_CPS_BACK));
    h->cost_grid = std::max(1, std::min(pow2ceil((Oloc + 255) / 256), h->num_sms * 4));
    h->p2l_grid = std::max(1, (np2l + 127) / 128);
    // dense Cholesky: one thread-block cluster when the reduced system is small enough, per-panel launches otherwise
    h->chol_cluster = 0;
    h->band_smem = 0;
    {
        // banded single-CTA factorisation (variant 6) whenever the band fits in the shared memory of one SM
        const char *v = getenv("SDV_CHOL_VARIANT");
        const int bw = std::max(h->band_bw, 1);
        if (h->band_bw >= 0 && bw <= BAND_MAX_BW && (!v || atoi(v) == 6) && !getenv("SDV_CHOL_DENSE")) {
            const BandPlan pl = band_plan(n_pad, bw);
            const size_t bytes = sizeof(double) * (size_t)pl.o_end;
            if (bytes <= 220 * 1024 && (size_t)pl.nb * (bw + 2) * 256 <= sb_elems) {
                h->band_smem = (int)bytes;
                P.band_bw = bw;
            }
        }
    }
    if (h->band_smem == 0 && n_pad <= 4096 && !getenv("SDV_NO_CLUSTER")) { // wide band (e.g. a dense prior over kept landmarks): 16-CTA cluster Cholesky
        cudaFuncSetAttribute((const void *)k_chol_chain<3, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaFuncSetAttribute((const void *)k_chol_chain<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        CK(cudaGetLastError());
        const int T = n_pad / 32, cs = 16;
        const int rows = std::max(1, (T + cs - 1) / cs); // shared-memory tiles: look-ahead operand / backward-solve tile inverses
        const int smem = (int)(sizeof(double) * (32 * TSTR + 32 + CHAIN_SMEM_DOUBLES + (size_t)2 * rows * 32 * TSTR + (size_t)rows * 32 + 8 * 32 + CC_MAX * 32 + (size_t)rows * 32));
        if (smem <= 200 * 1024) {
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(cs);
            lc.blockDim = dim3(CCT);
            lc.dynamicSmemBytes = smem;
            cudaLaunchAttribute at1[1];
            at1[0].id = cudaLaunchAttributeClusterDimension;
            at1[0].val.clusterDim.x = cs;
            at1[0].val.clusterDim.y = 1;
            at1[0].val.clusterDim.z = 1;
            lc.attrs = at1;
            lc.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, k_chol_chain<3, false>, &lc) == cudaSuccess && nclusters >= 1) {
                h->chol_cluster = cs;
                h->chol_rows_roles = rows;
                h->chol_smem_chain = smem;
            }
            cudaGetLastError();
        }
    }

    // ---- the problem description is complete: it goes to the device with the arena (ONE copy for everything but the bulk data)
    std::memcpy(hb + o_P, &P, sizeof(DevProblem));
    h->d_P = reinterpret_cast<DevProblem *>(h->d_in + o_P);
    auto t_pack1 = std::chrono::steady_clock::now();
    CK(cudaEventRecord(h->ev[0], h->stream));
    CK(cudaMemcpyAsync(h->d_in, hb, A.size, cudaMemcpyHostToDevice, h->stream));
    if (have_pool) h->pool.wait(g_bulk); // the bulk data arena is packed (and its copy issued on the copy stream)
    if (bulk_failed.load()) return fail(h, SDV_ERR_CUDA, "H2D copy of the bulk arena failed");
    if (bulk_copy_issued) CK(cudaStreamWaitEvent(h->stream, h->ev_bulk, 0));
    else CK(cudaMemcpyAsync(h->d_in2, h->h_in2, A2.size, cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->ev[1], h->stream));
    h->h2d_last = A.size + A2.size;
    if (gpu_struct && Oloc > 0) {
        // slot lists and tiles on the device, from the observation arrays that just arrived (sdv_struct.cuh): eight small launches,
        // captured once per handle (everything they need is read from the device-resident problem description) and replayed
        if (!h->stream2) CK(cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
        if (!h->sgraph_exec || h->sgraph_dP != (const void *)h->d_P || h->sgraph_capO != h->cap_O || h->sgraph_capL != h->cap_L) {
            if (h->sgraph_exec) cudaGraphExecDestroy(h->sgraph_exec);
            h->sgraph_exec = nullptr;
            cudaStream_t s2 = h->stream2;
            const int gO = std::min(h->num_sms * 8, (h->cap_O + 255) / 256), nbO = h->cap_O / SCAN_BLOCK + 1;
            const int gL = std::min(h->num_sms * 8, (h->cap_L + 255) / 256), nbL = h->cap_L / SCAN_BLOCK + 1;
            cudaGraph_t g = nullptr;
            CK(cudaStreamBeginCapture(s2, cudaStreamCaptureModeThreadLocal));
            k_struct_heads<<<gO, 256, 0, s2>>>(h->d_P);
            k_scan_blocks<<<nbO, SCAN_T, 0, s2>>>(h->d_P, 0);
            k_scan_sums<<<1, 1024, 0, s2>>>(h->d_P, 0);
            k_struct_slots<<<gO, 256, 0, s2>>>(h->d_P);
            k_struct_tile_heads<<<gL, 256, 0, s2>>>(h->d_P, FT_LMK);
            k_scan_blocks<<<nbL, SCAN_T, 0, s2>>>(h->d_P, 1);
            k_scan_sums<<<1, 1024, 0, s2>>>(h->d_P, 1);
            k_struct_tiles<<<gL, 256, 0, s2>>>(h->d_P);
            CK(cudaStreamEndCapture(s2, &g));
            CK(cudaGraphInstantiate(&h->sgraph_exec, g, 0));
            cudaGraphDestroy(g);
            h->sgraph_dP = (const void *)h->d_P;
            h->sgraph_capO = h->cap_O;
            h->sgraph_capL = h->cap_L;
        }
        CK(cudaGraphLaunch(h->sgraph_exec, h->stream));
        h->launches += 8;
    }
    if (h->band_smem == 0) CK(cudaMemsetAsync(h->d_Lo, 0, sizeof(double) * sb_elems, h->stream)); // cluster Cholesky: tiles outside the structural pattern are never written
    // ---- one-time device setup for this window
    if (Pn > 0) {
        k_imu_inf_sqrt<<<(Pn + 3) / 4, 128, 0, h->stream>>>(h->d_P);
        h->launches++;
    }
    if (dp && nm > 0) {
        k_prior_setup<<<std::min(64, (nm * nm + 255) / 256 + 1), 256, 0, h->stream>>>(h->d_P);
        h->launches++;
    }
    CK(cudaGetLastError());
    if (sync) CK(cudaStreamSynchronize(h->stream));
    h->resident = true;
    auto t_g0 = std::chrono::steady_clock::now();
    if ((rc = build_solve_graph(h)) != SDV_OK) return rc;
    if (timing) {
        auto t_g1 = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        std::fprintf(stderr, "[sdv upload] reduced system n=%d (%d tiles), factor tiles %d of %d structurally non-zero, half-bandwidth %d blocks of 16 (%s)\n", n, n_pad / 32,
                     h->chol_tiles_nz, h->chol_tiles_all, h->band_bw, h->band_smem > 0 ? "k_chol_band" : "cluster Cholesky");
        std::fprintf(stderr, "[sdv upload] total %.3f ms: validation pass %.3f, structure %.3f (columns %.3f, slots %.3f, tiles %.3f, masks %.3f), pack + scratch layout %.3f, "
                             "h2d + setup kernels + sync %.3f, graph %.3f\n",
                     ms(t_entry, t_g1), ms(t_entry, tu0), ms(tu0, t_pack0), ms(tu0, t_s1), ms(t_s1, t_s2), ms(t_s2, t_s3), ms(t_s3, t_pack0), ms(t_pack0, t_pack1), ms(t_pack1, t_g0),
                     ms(t_g0, t_g1));
    }
    return SDV_OK;
}

} // extern "C"

namespace {

void launch_schur(sdv_handle *h) { // fused visual linearisation + landmark Schur complement + assembly of the visual part of S
    const DevProblem &P = h->P;
    if (P.ntiles <= 0) return;
    cudaStream_t s = h->stream;
    if (P.kind == SDV_FACTOR_ANGULAR) k_lin_schur<0><<<h->fused_grid, FT, FT_SMEM_SCHUR, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, h->opt, h->d_Sb, h->d_scale_l, h->d_lmk_aux);
    else k_lin_schur<1><<<h->fused_grid, FT, FT_SMEM_SCHUR, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, h->opt, h->d_Sb, h->d_scale_l, h->d_lmk_aux);
    h->launches++;
}
// the iteration's back-substitution kernel can also clear the reduced system for the next iteration (no memset node per iteration)
// (EXPERIMENT, SDV_FOLD_S_CLEAR=1: measured SLOWER than the memset node — 0.514 against 0.499 ms per C3 solve, median of 200 —, off)
bool backsub_clears_S(const sdv_handle *h) { return h->P.ntiles > 0 && getenv("SDV_FOLD_S_CLEAR"); }

void launch_backsub(sdv_handle *h, bool clear_S = false) { // landmark back-substitution and the candidate cost of the visual factors in one kernel
    const DevProblem &P = h->P;
    if (P.ntiles <= 0) return;
    cudaStream_t s = h->stream;
    double *Sz = clear_S ? h->d_Sb : nullptr;
    const long long nz = clear_S ? (long long)h->sb_elems : 0;
    if (P.kind == SDV_FACTOR_ANGULAR) k_backsub_cost<0><<<h->fused_grid_back, FT, 0, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, h->d_dxp, h->d_lmk_aux, Sz, nz);
    else k_backsub_cost<1><<<h->fused_grid_back, FT, 0, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, h->d_dxp, h->d_lmk_aux, Sz, nz);
    h->launches++;
}

// fork: work launched on h->side after this call runs concurrently with what follows on h->stream (also under stream capture)
bool fork_side(sdv_handle *h, int slot) {
    if (!h->side) return false;
    if (cudaEventRecord(h->ev_fork[slot], h->stream) != cudaSuccess) return false;
    return cudaStreamWaitEvent(h->side, h->ev_fork[slot], 0) == cudaSuccess;
}
void join_side(sdv_handle *h, int slot) {
    cudaEventRecord(h->ev_join[slot], h->side);
    cudaStreamWaitEvent(h->stream, h->ev_join[slot], 0);
}

bool has_factors(const DevProblem &P) {
    return P.rank == 0 && (P.P > 0 || P.has_prior || P.mp_nfull > 0 || P.sp_has_imu || P.sp_has_lmk || P.sp_nl2l > 0);
}

void launch_lin_factors(sdv_handle *h, int which, cudaStream_t s) {
    const DevProblem &P = h->P;
    if (has_factors(P)) {
        k_lin_factors<<<h->fac_grid, FAC_WARPS * 32, 0, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, which);
        h->launches++;
    }
}

// materialise = true: residual + Jacobian planes of every observation (sdv_eval_visual, the round-1 solve path);
// false: the fused solve path only needs the COST of the visual factors here (iteration 0; the candidate cost comes out of
// k_backsub_cost, the Jacobians are formed inside k_lin_schur)
void launch_lin_visual(sdv_handle *h, int which, bool materialise) {
    const DevProblem &P = h->P;
    if (P.o1 > P.o0 && materialise) {
        h->lin_fn<<<h->lin_grid, LIN_THREADS, h->lin_smem, h->stream>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, which);
        h->launches++;
    } else if (P.o1 > P.o0 && which >= 0) {
        const int grid = h->cost_grid;
        if (P.kind == SDV_FACTOR_ANGULAR) k_visual_cost<0><<<grid, 256, 0, h->stream>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, which);
        else k_visual_cost<1><<<grid, 256, 0, h->stream>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, which);
        h->launches++;
    }
    if (P.sp_np2l > 0) {
        k_lin_p2l<<<h->p2l_grid, 128, 0, h->stream>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, which);
        h->launches++;
    }
}

int launch_linearize(sdv_handle *h, int which, bool materialise) {
    launch_lin_visual(h, which, materialise);
    launch_lin_factors(h, which, h->stream);
    return SDV_OK;
}

// one kernel over NVLink peer memory (sdv_peer.cuh) when the areas are open and the payload fits; false = use NCCL
template <int MODE> bool peer_allreduce(sdv_handle *h, const PeerMap &m, size_t count) {
    if (!h->peer_ok || count > h->peer.cap) return false;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((count + 4095) / 4096, (size_t)std::min(PEER_MAX_CHUNKS, h->num_sms)));
    k_peer_allreduce<MODE><<<grid, PEER_THREADS, 0, h->stream>>>(h->peer, m, (unsigned long long)count);
    h->launches++;
    return true;
}

int allreduce(sdv_handle *h, double *buf, size_t count) {
    if (h->world <= 1) return SDV_OK;
    {
        PeerMap m;
        std::memset(&m, 0, sizeof(m));
        m.buf = buf;
        if (peer_allreduce<0>(h, m, count)) return SDV_OK;
    }
    if (g_nccl.allreduce(buf, buf, count, NCCL_DOUBLE, NCCL_SUM, h->comm, h->stream) != 0) return fail(h, SDV_ERR_COMM, "ncclAllReduce failed");
    return SDV_OK;
}

// Multi-GPU exchange buffer of one LM iteration: the BAND of S (row i: its last W = 16 (bw + 1) columns up to the diagonal — every
// structurally non-zero entry of the stored lower triangle), the three rows [g | diag | raw gradient], and the gradient-violation
// flag of this rank's landmark columns — ONE ncclAllReduce instead of the full (n_pad + 3) x n_pad buffer (0.47 MB instead of
// 4.3 MB at C3, 1.9 MB instead of 72 MB at C5).  dir 0: pack, dir 1: unpack the sums.
__global__ void k_band_exchange(double *Sb, double *pack, Accum *acc, int n_pad, int ld, int W, double grad_tol, int dir) {
    const size_t nband = (size_t)n_pad * W, ntot = nband + 3 * (size_t)n_pad;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    for (size_t e = tid; e < ntot; e += nt) {
        double *p;
        if (e < nband) {
            const int i = (int)(e / W), c = (int)(e - (size_t)i * W), col = i - W + 1 + c;
            if (col < 0) {
                if (dir == 0) pack[e] = 0.0;
                continue;
            }
            p = Sb + (size_t)i * ld + col;
        } else {
            p = Sb + (size_t)n_pad * ld + (e - nband); // rows n_pad .. n_pad + 2 are contiguous (ld == n_pad)
        }
        if (dir == 0) pack[e] = *p;
        else *p = pack[e];
    }
    if (tid == 0) {
        // grad_max over the landmark columns is a MAX across ranks: exchanged as a count of ranks above the tolerance
        if (dir == 0) pack[ntot] = (__longlong_as_double((long long)acc->grad_max_bits) > grad_tol) ? 1.0 : 0.0;
        else acc->grad_max_bits = pack[ntot] > 0.0 ? (unsigned long long)__double_as_longlong(1e300) : 0ull;
    }
}

__global__ void k_pack_scalars(const LMState *st, Accum *acc, double *red, int dir, int which /* buffer: -2 candidate, 0 initial */) {
    int b = which >= 0 ? which : 1 - st->cur;
    if (dir == 0) {
        red[0] = acc->cost[b];
        red[1] = acc->model_gd;
        red[2] = acc->model_dd;
        red[3] = acc->step_norm2;
        red[4] = acc->cand_norm2;
        red[5] = which == 0 ? acc->fixed_cost : 0.0; // accumulated on rank 0 during the first linearisation only: reduced once
        red[6] = acc->schur_fail ? 1.0 : 0.0; // a 3x3 block that failed on ONE shard invalidates the step on every rank
    } else {
        acc->cost[b] = red[0];
        acc->model_gd = red[1];
        acc->model_dd = red[2];
        acc->step_norm2 = red[3];
        acc->cand_norm2 = red[4];
        if (which == 0) acc->fixed_cost = red[5];
        acc->schur_fail = red[6] > 0.0 ? 1 : 0;
    }
}

int reduce_scalars(sdv_handle *h, int which) {
    if (h->world <= 1) return SDV_OK;
    {
        PeerMap m;
        std::memset(&m, 0, sizeof(m));
        m.acc = h->d_acc;
        m.st = h->d_st;
        m.which = which;
        if (peer_allreduce<2>(h, m, 7)) return SDV_OK;
    }
    k_pack_scalars<<<1, 1, 0, h->stream>>>(h->d_st, h->d_acc, h->d_red, 0, which);
    int rc = allreduce(h, h->d_red, 7);
    if (rc != SDV_OK) return rc;
    k_pack_scalars<<<1, 1, 0, h->stream>>>(h->d_st, h->d_acc, h->d_red, 1, which);
    h->launches += 2;
    return SDV_OK;
}

int launch_factor_solve(sdv_handle *h) {
    const DevProblem &P = h->P;
    cudaStream_t s = h->stream;
    const int T = P.n_pad / CH_T;
    if (h->band_smem > 0) {
#if SDV_BAND_BABE
        // Two-way dissection (experiment, see sdv_chol_band.cuh): a 2-CTA cluster when both interiors are a few block rows
        // long, the band storage holds two factors and no kept landmark sits behind the frame blocks (a pure band).
        const int nbg = P.n_pad / BN, bwb = P.band_bw;
        if (nbg >= 4 * bwb + 8 && (size_t)2 * nbg * (bwb + 2) * 256 <= h->sb_elems && !getenv("SDV_BAND_NO_BABE")) {
            cudaLaunchConfig_t lc = {};
            cudaLaunchAttribute at[1];
            lc.gridDim = dim3(2);
            lc.blockDim = dim3(BCT);
            lc.dynamicSmemBytes = h->band_smem;
            lc.stream = s;
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            lc.attrs = at;
            lc.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelEx(&lc, k_chol_band, (const DevProblem *)h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, h->opt, (const double *)h->d_Sb, h->d_Lo, h->d_scale_p, h->d_damp_p,
                                               h->d_graw_p, h->d_dxp, h->d_prof);
            if (e != cudaSuccess) return fail(h, SDV_ERR_CUDA, cudaGetErrorString(e));
            h->launches++;
            return SDV_OK;
        }
#endif
        k_chol_band<<<1, BCT, h->band_smem, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, h->opt, h->d_Sb, h->d_Lo, h->d_scale_p, h->d_damp_p, h->d_graw_p, h->d_dxp, h->d_prof);
        h->launches++;
        return SDV_OK;
    }
    if (h->chol_cluster > 0) {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(h->chol_cluster);
        lc.blockDim = dim3(CCT);
        lc.dynamicSmemBytes = h->chol_smem_chain;
        lc.stream = s;
        cudaLaunchAttribute at1[1];
        at1[0].id = cudaLaunchAttributeClusterDimension;
        at1[0].val.clusterDim.x = h->chol_cluster;
        at1[0].val.clusterDim.y = 1;
        at1[0].val.clusterDim.z = 1;
        lc.attrs = at1;
        lc.numAttrs = 1;
        lc.dynamicSmemBytes = h->chol_smem_chain;
        cudaError_t e = cudaLaunchKernelEx(&lc, k_chol_chain<3, false>, (const DevProblem *)h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, h->d_Sb, h->d_Lo, h->d_dinv, (const double *)h->d_damp_p,
                                           (const double *)h->d_graw_p, h->d_dxp, h->chol_rows_roles, h->d_prof);
        if (e != cudaSuccess) return fail(h, SDV_ERR_CUDA, std::string("k_chol_cluster launch: ") + cudaGetErrorString(e));
        h->launches++;
        return SDV_OK;
    }
    for (int k = 0; k < T; k++) {
        int npanel = T - k + 1, ntrail = 0;
        for (int i = k + 1; i <= T; i++) ntrail += std::min(i, T - 1) - k;
        k_chol_panel<<<npanel + ntrail, CH_THREADS, 0, s>>>(h->d_Sb, h->d_Lo, P.ld, T, k, h->d_st, h->d_acc);
        h->launches++;
    }
    k_trisolve<<<1, TS_THREADS, (int)((P.n_pad + 1024 + 32 * 33) * sizeof(double)), s>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, h->d_Lo, h->d_damp_p,
                                                                                        h->d_graw_p, h->d_dxp);
    h->launches++;
    return SDV_OK;
}

int launch_iteration(sdv_handle *h) {
    const DevProblem &P = h->P;
    cudaStream_t s = h->stream;
    // S starts from zero: cleared by k_reset (first iteration) / by the previous iteration's k_backsub_cost, else by a memset node
    if (!backsub_clears_S(h) && cudaMemsetAsync(h->d_Sb, 0, h->sb_elems * sizeof(double), s) != cudaSuccess) return fail(h, SDV_ERR_CUDA, "memset S");
    {
        // the non-visual factors accumulate into S with atomics as well: run them beside the landmark Schur kernel
        const bool fa = P.rank == 0 && (P.P > 0 || P.has_prior || P.mp_nfull > 0);
        const bool fs = P.rank == 0 && (P.sp_has_imu || P.sp_has_lmk || P.sp_nl2l > 0);
        const bool forked = (fa || fs) && fork_side(h, 0);
        cudaStream_t fstream = forked ? h->side : s;
        launch_schur(h);
        if (fa) {
            k_assemble_factors<<<h->fac_grid, FAC_WARPS * 32, 0, fstream>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_Sb);
            h->launches++;
        }
        if (fs) {
            k_assemble_sparse<<<2, 128, 0, fstream>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_Sb);
            h->launches++;
        }
        if (forked) join_side(h, 0);
    }
    if (h->world > 1) {
        // ONE all-reduce per LM iteration: the band of S + [g | diag | grad] + the gradient flag, packed (k_band_exchange).  A
        // reduced system without a band (dense prior over kept landmarks: wide-band fallback) exchanges all its rows.
        const int W = h->band_smem > 0 ? std::min(P.n_pad, 16 * (P.band_bw + 1)) : P.n_pad;
        const size_t count = (size_t)P.n_pad * W + 3 * (size_t)P.n_pad + 1;
        PeerMap pm;
        std::memset(&pm, 0, sizeof(pm));
        pm.buf = h->d_Sb;
        pm.n_pad = P.n_pad;
        pm.ld = P.ld;
        pm.W = W;
        pm.acc = h->d_acc;
        pm.grad_tol = h->opt.gradient_tolerance;
        if (!peer_allreduce<1>(h, pm, count)) { // NCCL: pack -> all-reduce -> unpack
            const int grid = (int)std::min<size_t>((count + 255) / 256, (size_t)h->num_sms * 8);
            k_band_exchange<<<grid, 256, 0, s>>>(h->d_Sb, h->d_xchg, h->d_acc, P.n_pad, P.ld, W, h->opt.gradient_tolerance, 0);
            int rc = g_nccl.allreduce(h->d_xchg, h->d_xchg, count, NCCL_DOUBLE, NCCL_SUM, h->comm, s) != 0 ? fail(h, SDV_ERR_COMM, "ncclAllReduce failed") : SDV_OK;
            if (rc != SDV_OK) return rc;
            k_band_exchange<<<grid, 256, 0, s>>>(h->d_Sb, h->d_xchg, h->d_acc, P.n_pad, P.ld, W, h->opt.gradient_tolerance, 1);
            h->launches += 2;
        }
    }
    if (h->band_smem == 0) { // k_chol_band prepares the system itself
        k_sysprep<<<1, 1024, 0, s>>>(h->d_P, h->d_st, h->d_acc, h->opt, h->d_Sb, h->d_scale_p, h->d_damp_p, h->d_graw_p);
        h->launches++;
    }
    {
        int rcf = launch_factor_solve(h);
        if (rcf != SDV_OK) return rcf;
    }
    {
        // candidate linearisation: the factor kernel only needs the reduced parameters -> beside back-substitution + visual kernel
        const bool forked = has_factors(P) && fork_side(h, 1);
        launch_lin_factors(h, -2, forked ? h->side : s);
        launch_backsub(h, backsub_clears_S(h));
        launch_lin_visual(h, -2, false); // only the PoseToLandmark pseudo-observations are linearised here (the visual cost comes out of k_backsub_cost)
        if (forked) join_side(h, 1);
    }
    int rc = reduce_scalars(h, -2);
    if (rc != SDV_OK) return rc;
    k_ctrl<<<1, 32, 0, s>>>(h->d_P, h->d_st, h->d_acc, h->opt, h->cond);
    h->launches++;
    return SDV_OK;
}

} // namespace

namespace {

int enqueue_prologue(sdv_handle *h) {
    const DevProblem &P = h->P;
    cudaStream_t s = h->stream;
    // x = 0, solver state and accumulators cleared (and the reduced system, when the iterations do not clear it with a memset node)
    k_reset<<<64, 256, 0, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, backsub_clears_S(h) ? h->d_Sb : nullptr, backsub_clears_S(h) ? (long long)h->sb_elems : 0);
    k_prep_table<<<(P.F * P.C + 127) / 128, 128, 0, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, 0);
    h->launches += 2;
    launch_linearize(h, 0, false);
    int rc = reduce_scalars(h, 0);
    if (rc != SDV_OK) return rc;
    k_ctrl_init<<<1, 1, 0, s>>>(h->d_P, h->d_st, h->d_acc, h->opt);
    h->launches++;
    return SDV_OK;
}

size_t solution_doubles(const DevProblem &P) { return (size_t)15 * P.F + 3 * (size_t)std::max(P.L, 1); }
size_t solution_bytes(const DevProblem &P) { return solution_doubles(P) * sizeof(double) + sizeof(LMState) + sizeof(Accum); }

// solution blocks + solver state, packed on the device as [solution | LMState | Accum] (part of the CUDA graph) ...
int enqueue_epilogue(sdv_handle *h) {
    k_gather_solution<<<64, 256, 0, h->stream>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, reinterpret_cast<double *>(h->d_out));
    h->launches++;
    if (h->world > 1) {
        // every rank only solved its own landmarks and left zeros for the others: a sum over ranks is the gather.  Done ONCE,
        // here, so that sdv_download_delta is a plain host copy (no hidden collective, idempotent).  The capacity of the landmark
        // block is reduced (zeros beyond L), so the count does not tie the CUDA graph to the window.
        return allreduce(h, reinterpret_cast<double *>(h->d_out + sizeof(double) * 15 * (size_t)h->P.F + sizeof(LMState) + sizeof(Accum)), 3 * (size_t)h->cap_L);
    }
    return SDV_OK;
}
// ... and ONE copy to pinned host memory, issued behind the graph: its size depends on the window, the graph does not
int enqueue_readback(sdv_handle *h) {
    CK(cudaMemcpyAsync(h->h_sol, h->d_out, solution_bytes(h->P), cudaMemcpyDeviceToHost, h->stream));
    if (h->world > 1 && h->peer_ok) // the time-out word of the peer exchange rides along (a peer that never arrived: SDV_ERR_COMM)
        CK(cudaMemcpyAsync(h->h_rb + sizeof(LMState) + sizeof(Accum) + 64, h->peer.epoch + 2, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    return SDV_OK;
}

void destroy_active_graph(sdv_handle *h) {
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    if (h->graph) cudaGraphDestroy(h->graph);
    h->gexec = nullptr;
    h->graph = nullptr;
    h->graph_ok = false;
    h->cond = 0;
}
void destroy_graph(sdv_handle *h) { // the active graph and every cached one
    destroy_active_graph(h);
    for (auto &c : h->gcache) {
        if (c.gexec) cudaGraphExecDestroy(c.gexec);
        if (c.graph) cudaGraphDestroy(c.graph);
        c = sdv_handle::GraphSlot();
    }
}
int desired_unroll(const sdv_handle *h) {
    static const int forced = getenv("SDV_GRAPH_UNROLL") ? std::max(1, std::min(4, atoi(getenv("SDV_GRAPH_UNROLL")))) : 0;
    if (forced) return forced;
    const int it = std::max(1, h->last_iters);
    for (int k = 4; k > 1; k--)
        if (it % k == 0) return k;
    return 1;
}

// prologue -> WHILE (status == 0) { one LM iteration } -> epilogue, as ONE graph launch per solve.
int build_solve_graph(sdv_handle *h) {
    // N > 1: the NCCL all-reduces are captured into the WHILE body like any other node (NCCL >= 2.9 supports stream capture);
    // SDV_MGPU_NO_GRAPH=1 or a failed capture / instantiation falls back to the host loop that enqueues iterations ahead
    if ((h->world > 1 && getenv("SDV_MGPU_NO_GRAPH")) || getenv("SDV_NO_GRAPH")) {
        destroy_graph(h); // (a graph of an earlier window must not survive)
        return SDV_OK;
    }
    sdv_handle::GraphSig sig;
    std::memset(&sig, 0, sizeof(sig));
    {
        const DevProblem &P = h->P;
        sig.d_in = h->d_in; sig.d_scr = h->d_scr; sig.d_out = h->d_out;
        sig.F = P.F; sig.C = P.C; sig.vio = P.vio; sig.kind = P.kind; sig.n_pad = P.n_pad; sig.band_bw = P.band_bw; sig.band_smem = h->band_smem;
        sig.chol_cluster = h->chol_cluster; sig.chol_rows_roles = h->chol_rows_roles; sig.chol_smem_chain = h->chol_smem_chain;
        sig.cap_O = h->cap_O; sig.cap_L = h->cap_L; sig.cap_P = h->cap_P; sig.cap_nm = h->cap_nm; sig.cap_nfull = h->cap_nfull; sig.cap_l2l = h->cap_l2l;
        sig.fused_grid = h->fused_grid; sig.fused_grid_back = h->fused_grid_back; sig.fac_grid = h->fac_grid; sig.cost_grid = h->cost_grid;
        sig.p2l_grid = h->p2l_grid; sig.world = h->world;
        // which launches exist at all
        sig.flags = (P.ntiles > 0 ? 1 : 0) | (P.o1 > P.o0 ? 2 : 0) | (has_factors(P) ? 4 : 0) | (P.sp_np2l > 0 ? 8 : 0) |
                    ((P.rank == 0 && (P.P > 0 || P.has_prior || P.mp_nfull > 0)) ? 16 : 0) | ((P.rank == 0 && (P.sp_has_imu || P.sp_has_lmk || P.sp_nl2l > 0)) ? 32 : 0);
    }
    const int unroll = desired_unroll(h);
    if (h->graph_ok && h->graph_unroll == unroll && std::memcmp(&h->graph_sig, &sig, sizeof(sig)) == 0) return SDV_OK; // same launches, same arguments: replay
    if (h->graph_ok) { // park the active graph (least recently used slot) ...
        sdv_handle::GraphSlot *c = &h->gcache[0];
        for (auto &q : h->gcache) {
            if (!q.gexec) {
                c = &q;
                break;
            }
            if (q.last_use < c->last_use) c = &q;
        }
        if (c->gexec) cudaGraphExecDestroy(c->gexec);
        if (c->graph) cudaGraphDestroy(c->graph);
        c->graph = h->graph; c->gexec = h->gexec; c->cond = h->cond; c->sig = h->graph_sig; c->unroll = h->graph_unroll;
        c->l_fixed = h->graph_launches_fixed; c->l_iter = h->graph_launches_iter; c->last_use = ++h->graph_clock;
        h->graph = nullptr; h->gexec = nullptr; h->graph_ok = false; h->cond = 0;
    }
    for (auto &c : h->gcache) // ... and look for a parked one that fits
        if (c.gexec && c.unroll == unroll && std::memcmp(&c.sig, &sig, sizeof(sig)) == 0) {
            h->graph = c.graph; h->gexec = c.gexec; h->cond = c.cond; h->graph_sig = c.sig; h->graph_launches_fixed = c.l_fixed; h->graph_launches_iter = c.l_iter;
            h->graph_unroll = unroll;
            h->graph_ok = true;
            c = sdv_handle::GraphSlot();
            return SDV_OK;
        }
    destroy_active_graph(h);
    if (!h->stream2 && cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking) != cudaSuccess) return SDV_OK;
    cudaStream_t s = h->stream;
    int64_t l0 = h->launches;
    bool ok = true;
    cudaGraph_t g = nullptr;
    if (cudaGraphCreate(&g, 0) != cudaSuccess) return SDV_OK;
    cudaGraphConditionalHandle handle = 0;
    if (cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault) != cudaSuccess) {
        cudaGraphDestroy(g);
        cudaGetLastError();
        return SDV_OK;
    }
    h->cond = (unsigned long long)handle;
    int64_t l_fixed = 0, l_iter = 0;
    do {
        if (cudaStreamBeginCaptureToGraph(s, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { ok = false; break; }
        if (enqueue_prologue(h) != SDV_OK) { ok = false; }
        l_fixed = h->launches - l0;
        // conditional WHILE node after everything captured so far
        cudaStreamCaptureStatus cst;
        const cudaGraphNode_t *deps = nullptr;
        size_t ndeps = 0;
        cudaGraph_t cg = nullptr;
        if (ok && cudaStreamGetCaptureInfo_v2(s, &cst, nullptr, &cg, &deps, &ndeps) != cudaSuccess) ok = false;
        cudaGraphNode_t cnode = nullptr;
        cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
        cp.type = cudaGraphNodeTypeConditional;
        cp.conditional.handle = handle;
        cp.conditional.type = cudaGraphCondTypeWhile;
        cp.conditional.size = 1;
        if (ok && cudaGraphAddNode(&cnode, g, deps, ndeps, &cp) != cudaSuccess) ok = false;
        if (ok) {
            cudaGraph_t body = cp.conditional.phGraph_out[0];
            cudaStream_t keep = h->stream;
            h->stream = h->stream2;
            int64_t l1 = h->launches;
            if (cudaStreamBeginCaptureToGraph(h->stream2, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal) != cudaSuccess) ok = false;
            for (int u = 0; ok && u < unroll; u++) // `unroll` LM iterations per trip of the WHILE node (see sdv_handle::graph_unroll)
                if (launch_iteration(h) != SDV_OK) ok = false;
            cudaGraph_t tmp = nullptr;
            if (cudaStreamEndCapture(h->stream2, &tmp) != cudaSuccess) ok = false;
            l_iter = (h->launches - l1) / unroll;
            h->stream = keep;
        }
        if (ok && cudaStreamUpdateCaptureDependencies(s, &cnode, 1, cudaStreamSetCaptureDependencies) != cudaSuccess) ok = false;
        int64_t l2 = h->launches;
        if (ok && enqueue_epilogue(h) != SDV_OK) ok = false;
        l_fixed += h->launches - l2;
        cudaGraph_t out = nullptr;
        if (cudaStreamEndCapture(s, &out) != cudaSuccess) ok = false;
    } while (0);
    h->launches = l0; // nothing was launched, only captured
    if (ok && cudaGraphInstantiate(&h->gexec, g, 0) != cudaSuccess) ok = false;
    cudaGetLastError();
    if (!ok) {
        if (h->gexec) cudaGraphExecDestroy(h->gexec);
        h->gexec = nullptr;
        cudaGraphDestroy(g);
        h->cond = 0;
        h->graph_ok = false;
        return SDV_OK; // fall back to the host-driven loop
    }
    h->graph = g;
    h->graph_ok = true;
    h->graph_unroll = unroll;
    h->graph_builds++;
    h->graph_sig = sig;
    h->graph_launches_fixed = l_fixed;
    h->graph_launches_iter = l_iter;
    return SDV_OK;
}

} // namespace

extern "C" {

int sdv_solve_resident(sdv_handle *h, sdv_stats *stats) {
    if (!h) return SDV_ERR_INVALID_ARGUMENT;
    if (!h->resident) return fail(h, SDV_ERR_INVALID_ARGUMENT, "no window uploaded");
    cudaSetDevice(h->device);
    const DevProblem &P = h->P;
    cudaStream_t s = h->stream;
    int64_t launches0 = h->launches;
    auto t0 = std::chrono::steady_clock::now();
    CK(cudaEventRecord(h->ev[2], s));
    const size_t nd = solution_doubles(P);
    if (h->graph_ok) {
        CK(cudaGraphLaunch(h->gexec, s));
        CK(cudaEventRecord(h->ev[3], s));
        int rcb = enqueue_readback(h);
        if (rcb != SDV_OK) return rcb;
        CK(cudaStreamSynchronize(s));
    } else {
        int rc = enqueue_prologue(h);
        if (rc != SDV_OK) return rc;
        // No CUDA graph here (NCCL inside a conditional WHILE body is not attempted).  Every kernel returns at once when the solve
        // has terminated and the collectives are issued by all ranks alike, so iterations can be enqueued AHEAD of the status:
        // as many as the previous solve took, then the status is read, then one more at a time.  A steady back end pays one
        // host round trip per solve instead of one per iteration.
        int *h_status = reinterpret_cast<int *>(h->h_rb);
        const int max_it = (P.max_iter > 0 ? P.max_iter : h->opt.max_num_iterations) + 1;
        int it = 0, ahead = std::max(1, std::min(h->last_iters, max_it));
        while (it < max_it) {
            for (int q = 0; q < ahead && it < max_it; q++, it++) {
                rc = launch_iteration(h);
                if (rc != SDV_OK) return rc;
            }
            CK(cudaMemcpyAsync(h_status, &h->d_st->status, sizeof(int), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            if (*h_status != 0) break;
            ahead = 1;
        }
        rc = enqueue_epilogue(h);
        if (rc != SDV_OK) return rc;
        CK(cudaEventRecord(h->ev[3], s));
        rc = enqueue_readback(h);
        if (rc != SDV_OK) return rc;
        CK(cudaStreamSynchronize(s));
    }
    CK(cudaGetLastError());
    (void)nd;
    if (h->world > 1 && h->peer_ok) {
        unsigned long long timed_out = 0;
        std::memcpy(&timed_out, h->h_rb + sizeof(LMState) + sizeof(Accum) + 64, sizeof(timed_out));
        if (timed_out) return fail(h, SDV_ERR_COMM, "peer-memory exchange timed out (a rank did not take part in the solve)");
    }
    std::memcpy(h->h_rb, h->h_sol + sizeof(double) * 15 * (size_t)P.F, sizeof(LMState) + sizeof(Accum));
    std::memcpy(&h->h_state, h->h_rb, sizeof(LMState));
    std::memcpy(&h->h_acc, h->h_rb + sizeof(LMState), sizeof(Accum));
    auto t1 = std::chrono::steady_clock::now();
    h->last_iters = std::max(1, h->h_state.iter);
    if (stats) {
        const LMState &st = h->h_state;
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
        stats->iterations = st.iter;
        stats->termination = st.status > 0 ? st.status - 1 : SDV_TERM_NO_CONVERGENCE;
        stats->num_successful_steps = st.n_ok;
        stats->num_unsuccessful_steps = st.n_bad;
        stats->n_reduced = P.n;
        stats->n_residual_blocks = 0;
        stats->initial_cost = st.initial_cost;
        stats->final_cost = st.x_cost;
        stats->fixed_cost = h->h_acc.fixed_cost;
        stats->final_radius = st.radius;
        for (int i = 0; i < SDV_MAX_TRACE; i++) {
            stats->trace_cost[i] = st.trace_cost[i];
            stats->trace_radius[i] = st.trace_radius[i];
            stats->trace_model_change[i] = st.trace_model[i];
            stats->trace_accepted[i] = st.trace_accepted[i];
        }
        stats->ms_solve_device = ms;
        stats->ms_total_host = std::chrono::duration<double, std::milli>(t1 - t0).count();
        if (h->graph_ok) h->launches += h->graph_launches_fixed + (int64_t)(st.iter + (st.status == 1 + SDV_TERM_GRADIENT_TOLERANCE ? 1 : 0)) * h->graph_launches_iter;
        stats->kernel_launches = h->launches - launches0;
    }
    return h->h_state.status == 1 + SDV_TERM_FAILURE ? SDV_ERR_NUMERICAL_FAILURE : SDV_OK;
}

int sdv_download_delta(sdv_handle *h, sdv_delta *out) {
    if (!h || !out || !out->dpose || !out->dlmk) return SDV_ERR_INVALID_ARGUMENT;
    if (!h->resident) return fail(h, SDV_ERR_INVALID_ARGUMENT, "no window uploaded");
    cudaSetDevice(h->device);
    const DevProblem &P = h->P;
    // the last solve already left the solution (summed over ranks at N > 1) in pinned host memory
    const double *hb = reinterpret_cast<const double *>(h->h_sol);
    std::memcpy(out->dpose, hb, sizeof(double) * 6 * P.F);
    if (out->dv) std::memcpy(out->dv, hb + 6 * P.F, sizeof(double) * 3 * P.F);
    if (out->dba) std::memcpy(out->dba, hb + 9 * P.F, sizeof(double) * 3 * P.F);
    if (out->dbg) std::memcpy(out->dbg, hb + 12 * P.F, sizeof(double) * 3 * P.F);
    if (P.L > 0) std::memcpy(out->dlmk, h->h_sol + sizeof(double) * 15 * (size_t)P.F + sizeof(LMState) + sizeof(Accum), sizeof(double) * 3 * P.L);
    return SDV_OK;
}

int sdv_solve_window(sdv_handle *h, const sdv_window *win, sdv_delta *out, sdv_stats *stats) {
    if (!h || !win || !out) return SDV_ERR_INVALID_ARGUMENT;
    auto t0 = std::chrono::steady_clock::now();
    int rc = upload_impl(h, win, false);
    if (rc != SDV_OK) return rc;
    sdv_stats local;
    sdv_stats *st = stats ? stats : &local;
    std::memset(st, 0, sizeof(*st));
    rc = sdv_solve_resident(h, st);
    if (rc != SDV_OK && rc != SDV_ERR_NUMERICAL_FAILURE) return rc;
    int rc2 = sdv_download_delta(h, out);
    if (rc2 != SDV_OK) return rc2;
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    st->ms_h2d = ms;
    st->h2d_bytes = (int64_t)h->h2d_last;
    st->d2h_bytes = (int64_t)(sizeof(double) * ((size_t)15 * h->P.F + 3 * (size_t)h->P.L) + sizeof(LMState) + sizeof(Accum));
    st->ms_total_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (getenv("SDV_TIMING")) std::fprintf(stderr, "[sdv solve_window] total %.3f ms, device solve %.3f ms, h2d %.3f ms\n", st->ms_total_host, st->ms_solve_device, st->ms_h2d);
    return rc;
}

// set the linearisation point of buffer 0 from user-supplied parameter blocks
static int set_point(sdv_handle *h, const sdv_delta *x) {
    const DevProblem &P = h->P;
    std::vector<double> xp(P.n_pad, 0.0), xl(3 * (size_t)std::max(P.L, 1), 0.0);
    std::vector<int> pose_col(P.F), vb_col(P.F), lmk_col(std::max(P.L, 1));
    CK(cudaMemcpy(pose_col.data(), P.pose_col, 4 * P.F, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(vb_col.data(), P.vb_col, 4 * P.F, cudaMemcpyDeviceToHost));
    if (P.L) CK(cudaMemcpy(lmk_col.data(), P.lmk_col, 4 * P.L, cudaMemcpyDeviceToHost));
    if (x) {
        for (int f = 0; f < P.F; f++) {
            if (pose_col[f] >= 0 && x->dpose)
                for (int k = 0; k < 6; k++) xp[pose_col[f] + k] = x->dpose[6 * f + k];
            if (vb_col[f] >= 0)
                for (int k = 0; k < 3; k++) {
                    if (x->dv) xp[vb_col[f] + k] = x->dv[3 * f + k];
                    if (x->dba) xp[vb_col[f] + 3 + k] = x->dba[3 * f + k];
                    if (x->dbg) xp[vb_col[f] + 6 + k] = x->dbg[3 * f + k];
                }
        }
        if (x->dlmk)
            for (int l = 0; l < P.L; l++)
                for (int k = 0; k < 3; k++) {
                    if (lmk_col[l] >= 0) xp[lmk_col[l] + k] = x->dlmk[3 * l + k];
                    xl[3 * (size_t)l + k] = x->dlmk[3 * l + k]; // kept landmarks are mirrored in xl (read by the visual kernel)
                }
    }
    CK(cudaMemcpyAsync(h->B[0].xp, xp.data(), sizeof(double) * P.n_pad, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->B[0].xl, xl.data(), sizeof(double) * 3 * std::max(P.L, 1), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->d_st, 0, sizeof(LMState), h->stream));
    CK(cudaMemsetAsync(h->d_acc, 0, sizeof(Accum), h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SDV_OK;
}

int sdv_eval_visual(sdv_handle *h, const sdv_delta *x, double *r, double *J_pose, double *J_lmk, double *cost) {
    if (!h) return SDV_ERR_INVALID_ARGUMENT;
    if (!h->resident) return fail(h, SDV_ERR_INVALID_ARGUMENT, "no window uploaded");
    cudaSetDevice(h->device);
    const DevProblem &P = h->P;
    int rc = set_point(h, x);
    if (rc != SDV_OK) return rc;
    k_prep_table<<<(P.F * P.C + 127) / 128, 128, 0, h->stream>>>(h->d_P, h->B[0], h->B[1], h->d_st, 0);
    h->launches++;
    launch_linearize(h, 0, true);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    const int Oloc = P.o1 - P.o0;
    const size_t OC = (size_t)std::max(P.Ocap, 1);
    std::vector<double> buf(20 * OC);
    CK(cudaMemcpy(buf.data(), h->B[0].r, sizeof(double) * 2 * OC, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(buf.data() + 2 * OC, h->B[0].Jp, sizeof(double) * 12 * OC, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(buf.data() + 14 * OC, h->B[0].Jl, sizeof(double) * 6 * OC, cudaMemcpyDeviceToHost));
    for (int ol = 0; ol < Oloc; ol++) {
        size_t o = (size_t)P.o0 + ol;
        if (r)
            for (int k = 0; k < 2; k++) r[2 * o + k] = buf[(size_t)k * OC + ol];
        if (J_pose)
            for (int k = 0; k < 12; k++) J_pose[12 * o + k] = buf[(size_t)(2 + k) * OC + ol];
        if (J_lmk)
            for (int k = 0; k < 6; k++) J_lmk[6 * o + k] = buf[(size_t)(14 + k) * OC + ol];
    }
    if (cost) {
        Accum a;
        CK(cudaMemcpy(&a, h->d_acc, sizeof(Accum), cudaMemcpyDeviceToHost));
        *cost = a.cost[0];
    }
    return SDV_OK;
}

int sdv_eval_imu(sdv_handle *h, const sdv_delta *x, double *r_imu, double *J_imu, double *r_bias) {
    if (!h) return SDV_ERR_INVALID_ARGUMENT;
    if (!h->resident) return fail(h, SDV_ERR_INVALID_ARGUMENT, "no window uploaded");
    cudaSetDevice(h->device);
    const DevProblem &P = h->P;
    if (P.P == 0) return SDV_OK;
    int rc = set_point(h, x);
    if (rc != SDV_OK) return rc;
    k_lin_factors<<<h->fac_grid, FAC_WARPS * 32, 0, h->stream>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, 0);
    h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    if (r_imu) CK(cudaMemcpy(r_imu, h->B[0].imu_r, sizeof(double) * 9 * P.P, cudaMemcpyDeviceToHost));
    if (J_imu) CK(cudaMemcpy(J_imu, h->B[0].imu_J, sizeof(double) * 216 * P.P, cudaMemcpyDeviceToHost));
    if (r_bias) CK(cudaMemcpy(r_bias, h->B[0].bias_r, sizeof(double) * 6 * P.P, cudaMemcpyDeviceToHost));
    return SDV_OK;
}

// IMU::processIMU over every keyframe interval (sdv_preint.cuh), host buffers in and out
int sdv_preintegrate(sdv_handle *h, const sdv_imu_intervals *in, sdv_preint *out) {
    if (!h || !in || !out) return SDV_ERR_INVALID_ARGUMENT;
    const int n = in->n_intervals, S = in->n_samples;
    if (n < 0 || S < 0) return fail(h, SDV_ERR_INVALID_ARGUMENT, "negative sizes");
    if (n == 0) return SDV_OK;
    if (!in->sample_ptr || !in->T_f_w || !in->v || !in->ba || !in->bg || (S > 0 && (!in->acc || !in->gyr || !in->dt)))
        return fail(h, SDV_ERR_INVALID_ARGUMENT, "null pre-integration inputs");
    if (!out->dR || !out->dv || !out->dp || !out->cov || !out->J_dR_bg || !out->J_dv_ba || !out->J_dv_bg || !out->J_dp_ba || !out->J_dp_bg)
        return fail(h, SDV_ERR_INVALID_ARGUMENT, "null pre-integration outputs");
    if (in->sample_ptr[0] != 0 || in->sample_ptr[n] != S) return fail(h, SDV_ERR_INVALID_ARGUMENT, "sample_ptr does not cover the samples");
    for (int k = 0; k < n; k++)
        if (in->sample_ptr[k + 1] < in->sample_ptr[k]) return fail(h, SDV_ERR_INVALID_ARGUMENT, "sample_ptr not monotonic");
    if (!(in->rate_hz > 0)) return fail(h, SDV_ERR_INVALID_ARGUMENT, "rate_hz must be positive");
    cudaSetDevice(h->device);
    cudaStream_t s = h->stream;
    const size_t D = sizeof(double);
    Arena A;
    const size_t i_sp = A.add(4 * (size_t)(n + 1)), i_acc = A.add(D * 3 * std::max(S, 1)), i_gyr = A.add(D * 3 * std::max(S, 1)), i_dt = A.add(D * std::max(S, 1));
    const size_t i_T = A.add(D * 12 * n), i_v = A.add(D * 3 * n), i_ba = A.add(D * 3 * n), i_bg = A.add(D * 3 * n), i_st = A.add(D * 9 * n);
    const size_t in_bytes = A.size;
    const size_t o_dR = A.add(D * 9 * n), o_dv = A.add(D * 3 * n), o_dp = A.add(D * 3 * n), o_cov = A.add(D * 81 * n), o_j1 = A.add(D * 9 * n), o_j2 = A.add(D * 9 * n),
                 o_j3 = A.add(D * 9 * n), o_j4 = A.add(D * 9 * n), o_j5 = A.add(D * 9 * n), o_T = A.add(D * 12 * n), o_v = A.add(D * 3 * n);
    std::vector<unsigned char> hb(A.size);
    std::memcpy(&hb[i_sp], in->sample_ptr, 4 * (size_t)(n + 1));
    if (S > 0) {
        std::memcpy(&hb[i_acc], in->acc, D * 3 * S);
        std::memcpy(&hb[i_gyr], in->gyr, D * 3 * S);
        std::memcpy(&hb[i_dt], in->dt, D * S);
    }
    std::memcpy(&hb[i_T], in->T_f_w, D * 12 * n);
    std::memcpy(&hb[i_v], in->v, D * 3 * n);
    std::memcpy(&hb[i_ba], in->ba, D * 3 * n);
    std::memcpy(&hb[i_bg], in->bg, D * 3 * n);
    if (in->dR_stale) std::memcpy(&hb[i_st], in->dR_stale, D * 9 * n);
    unsigned char *d = nullptr;
    CK(cudaMalloc((void **)&d, A.size));
    struct Free {
        unsigned char *p;
        ~Free() { cudaFree(p); }
    } guard{d};
    CK(cudaMemcpyAsync(d, hb.data(), in_bytes, cudaMemcpyHostToDevice, s));
    PreintArgs a;
    a.n_intervals = n;
    a.sample_ptr = at<int>(d, i_sp);
    a.acc = at<double>(d, i_acc); a.gyr = at<double>(d, i_gyr); a.dt = at<double>(d, i_dt);
    a.T_f_w = at<double>(d, i_T); a.v = at<double>(d, i_v); a.ba = at<double>(d, i_ba); a.bg = at<double>(d, i_bg);
    a.dR_stale = in->dR_stale ? at<double>(d, i_st) : nullptr;
    for (int k = 0; k < 6; k++) a.eta[k] = in->eta[k];
    a.rate_hz = in->rate_hz;
    a.dR = at<double>(d, o_dR); a.dv = at<double>(d, o_dv); a.dp = at<double>(d, o_dp); a.cov = at<double>(d, o_cov);
    a.J_dR_bg = at<double>(d, o_j1); a.J_dv_ba = at<double>(d, o_j2); a.J_dv_bg = at<double>(d, o_j3); a.J_dp_ba = at<double>(d, o_j4); a.J_dp_bg = at<double>(d, o_j5);
    a.T_pred = at<double>(d, o_T); a.v_pred = at<double>(d, o_v);
    k_preintegrate<<<(n + PRE_WARPS - 1) / PRE_WARPS, PRE_WARPS * 32, 0, s>>>(a);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&hb[o_dR], d + o_dR, A.size - o_dR, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::memcpy(out->dR, &hb[o_dR], D * 9 * n);
    std::memcpy(out->dv, &hb[o_dv], D * 3 * n);
    std::memcpy(out->dp, &hb[o_dp], D * 3 * n);
    std::memcpy(out->cov, &hb[o_cov], D * 81 * n);
    std::memcpy(out->J_dR_bg, &hb[o_j1], D * 9 * n);
    std::memcpy(out->J_dv_ba, &hb[o_j2], D * 9 * n);
    std::memcpy(out->J_dv_bg, &hb[o_j3], D * 9 * n);
    std::memcpy(out->J_dp_ba, &hb[o_j4], D * 9 * n);
    std::memcpy(out->J_dp_bg, &hb[o_j5], D * 9 * n);
    if (out->T_pred) std::memcpy(out->T_pred, &hb[o_T], D * 12 * n);
    if (out->v_pred) std::memcpy(out->v_pred, &hb[o_v], D * 3 * n);
    return SDV_OK;
}

// AOptimizer::VIInit's problem build + ceres::Solve (AOptimizer.cpp:448-529) as ONE kernel (sdv_viinit.cuh); host buffers in and out
int sdv_viinit(sdv_handle *h, const sdv_window *w, int32_t optim_scale, sdv_viinit_result *out, sdv_stats *stats) {
    if (!h || !w || !out) return SDV_ERR_INVALID_ARGUMENT;
    if (w->abi_version != SDV_ABI_VERSION) return fail(h, SDV_ERR_INVALID_ARGUMENT, "sdv_window.abi_version mismatch");
    const int F = w->n_frames, Pn = w->n_imu;
    if (F <= 0 || Pn < 0) return fail(h, SDV_ERR_INVALID_ARGUMENT, "VIInit needs at least one frame");
    if (!w->T_f_w || !w->v || !out->dv) return fail(h, SDV_ERR_INVALID_ARGUMENT, "null frame arrays");
    if (Pn > 0 && (!w->imu_i || !w->imu_j || !w->imu_dt || !w->imu_dR || !w->imu_dv || !w->imu_dp || !w->imu_cov))
        return fail(h, SDV_ERR_INVALID_ARGUMENT, "null IMU arrays");
    for (int p = 0; p < Pn; p++)
        if (w->imu_i[p] < 0 || w->imu_i[p] >= F || w->imu_j[p] < 0 || w->imu_j[p] >= F || w->imu_i[p] == w->imu_j[p])
            return fail(h, SDV_ERR_INVALID_ARGUMENT, "IMU pair index out of range");
    auto t0 = std::chrono::steady_clock::now();
    cudaSetDevice(h->device);
    cudaStream_t s = h->stream;
    const size_t D = sizeof(double);
    const int n = 3 * F + 2 + (optim_scale ? 1 : 0), Pc = std::max(Pn, 1);
    Arena A;
    const size_t i_P = A.add(sizeof(DevProblem)), i_T = A.add(D * 12 * F), i_v = A.add(D * 3 * F), i_ii = A.add(4 * (size_t)Pc), i_ij = A.add(4 * (size_t)Pc),
                 i_dt = A.add(D * Pc), i_dR = A.add(D * 9 * Pc), i_dv = A.add(D * 3 * Pc), i_dp = A.add(D * 3 * Pc), i_cov = A.add(D * 81 * Pc);
    const size_t in_bytes = A.size;
    const size_t s_inf = A.add(D * 81 * Pc), s_Jw = A.add(D * 81 * Pc), s_rw = A.add(D * 9 * Pc), s_A = A.add(D * (size_t)n * n), s_vec = A.add(D * 8 * (size_t)n);
    const size_t o_out = A.add(D * (3 * (size_t)F + 3)), o_st = A.add(sizeof(LMState)), o_acc = A.add(sizeof(Accum));
    std::vector<unsigned char> hb(A.size);
    unsigned char *d = nullptr;
    CK(cudaMalloc((void **)&d, A.size));
    struct Free {
        unsigned char *p;
        ~Free() { cudaFree(p); }
    } guard{d};
    DevProblem Pd;
    std::memset(&Pd, 0, sizeof(Pd));
    Pd.P = Pn;
    Pd.imu_cov = at<double>(d, i_cov);
    Pd.imu_inf_sqrt = at<double>(d, s_inf);
    std::memcpy(&hb[i_P], &Pd, sizeof(Pd));
    std::memcpy(&hb[i_T], w->T_f_w, D * 12 * F);
    std::memcpy(&hb[i_v], w->v, D * 3 * F);
    if (Pn > 0) {
        std::memcpy(&hb[i_ii], w->imu_i, 4 * (size_t)Pn);
        std::memcpy(&hb[i_ij], w->imu_j, 4 * (size_t)Pn);
        std::memcpy(&hb[i_dt], w->imu_dt, D * Pn);
        std::memcpy(&hb[i_dR], w->imu_dR, D * 9 * Pn);
        std::memcpy(&hb[i_dv], w->imu_dv, D * 3 * Pn);
        std::memcpy(&hb[i_dp], w->imu_dp, D * 3 * Pn);
        std::memcpy(&hb[i_cov], w->imu_cov, D * 81 * Pn);
    }
    CK(cudaMemcpyAsync(d, hb.data(), in_bytes, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(h->ev[2], s));
    const int64_t launches0 = h->launches;
    if (Pn > 0) {
        k_imu_inf_sqrt<<<(Pn + 3) / 4, 128, 0, s>>>(at<DevProblem>(d, i_P));
        h->launches++;
    }
    VIInitArgs a;
    a.F = F; a.P = Pn; a.n = n; a.optim_scale = optim_scale ? 1 : 0;
    a.max_iter = w->max_num_iterations > 0 ? w->max_num_iterations : 50; // int steps = 50 (AOptimizer.cpp:449, :521)
    if (a.max_iter > SDV_MAX_TRACE - 1) a.max_iter = SDV_MAX_TRACE - 1;
    const size_t a_bytes = D * (size_t)n * n;
    a.a_in_smem = a_bytes <= 200 * 1024 ? 1 : 0;
    a.T_f_w = at<double>(d, i_T); a.v = at<double>(d, i_v);
    a.imu_i = at<int>(d, i_ii); a.imu_j = at<int>(d, i_ij);
    a.imu_dt = at<double>(d, i_dt); a.imu_dR = at<double>(d, i_dR); a.imu_dv = at<double>(d, i_dv); a.imu_dp = at<double>(d, i_dp);
    a.inf_sqrt = at<double>(d, s_inf);
    a.Jw = at<double>(d, s_Jw); a.rw = at<double>(d, s_rw); a.A = at<double>(d, s_A); a.vecs = at<double>(d, s_vec);
    a.out = at<double>(d, o_out); a.st = at<LMState>(d, o_st); a.acc = at<Accum>(d, o_acc);
    a.opt = h->opt;
    if (!h->viinit_attr_done) { // (per handle: the attribute belongs to the device the handle is bound to)
        CK(cudaFuncSetAttribute(k_viinit, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        h->viinit_attr_done = true;
    }
    k_viinit<<<1, VI_THREADS, a.a_in_smem ? a_bytes : 0, s>>>(a);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[3], s));
    CK(cudaMemcpyAsync(&hb[o_out], d + o_out, A.size - o_out, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const double *o = reinterpret_cast<const double *>(&hb[o_out]);
    std::memcpy(out->dv, o, D * 3 * F);
    out->r_wi[0] = o[3 * F];
    out->r_wi[1] = o[3 * F + 1];
    out->lambda = optim_scale ? o[3 * F + 2] : 0.0;
    out->scale = std::exp(out->lambda);                                   // the value VIInit returns (AOptimizer.cpp:580)
    {   // R_w_i = exp_so3(r_wi[0], r_wi[1], 0)  (AOptimizer.cpp:540, geometry.h:131-147)
        const double wx = out->r_wi[0], wy = out->r_wi[1], ang = std::sqrt(wx * wx + wy * wy);
        double K[9] = {0, 0, wy, 0, 0, -wx, -wy, wx, 0}, *R = out->R_w_i;
        if (ang < 1e-9) {
            for (int k = 0; k < 9; k++) R[k] = K[k] + (k % 4 == 0 ? 1.0 : 0.0);
        } else {
            for (int k = 0; k < 9; k++) K[k] /= ang;
            const double sn = std::sin(ang), cs = std::cos(ang);
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) {
                    double k2 = 0;
                    for (int q = 0; q < 3; q++) k2 += K[r * 3 + q] * K[q * 3 + c];
                    R[r * 3 + c] = (r == c ? 1.0 : 0.0) + (1.0 - cs) * k2 + sn * K[r * 3 + c];
                }
        }
    }
    LMState st;
    std::memcpy(&st, &hb[o_st], sizeof(LMState));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
        stats->iterations = st.iter;
        stats->termination = st.status > 0 ? st.status - 1 : SDV_TERM_NO_CONVERGENCE;
        stats->num_successful_steps = st.n_ok;
        stats->num_unsuccessful_steps = st.n_bad;
        stats->n_reduced = n;
        stats->n_residual_blocks = Pn;
        stats->initial_cost = st.initial_cost;
        stats->final_cost = st.x_cost;
        stats->final_radius = st.radius;
        for (int i = 0; i < SDV_MAX_TRACE; i++) {
            stats->trace_cost[i] = st.trace_cost[i];
            stats->trace_radius[i] = st.trace_radius[i];
            stats->trace_model_change[i] = st.trace_model[i];
            stats->trace_accepted[i] = st.trace_accepted[i];
        }
        stats->ms_solve_device = ms;
        stats->ms_total_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        stats->kernel_launches = h->launches - launches0;
        stats->h2d_bytes = (int64_t)in_bytes;
        stats->d2h_bytes = (int64_t)(A.size - o_out);
    }
    return st.status == 1 + SDV_TERM_FAILURE ? SDV_ERR_NUMERICAL_FAILURE : SDV_OK;
}

// which: 0 = materialising visual residual+Jacobian kernel (k_lin_visual, the evaluation entry point), 1 = fused linearisation +
//        Schur + assembly (k_lin_schur), 2 = factorisation + solves of the reduced system, 3 = fused back-substitution + candidate
//        cost (k_backsub_cost); + 10 = the same with a cold L2 (see below)
int sdv_time_kernel(sdv_handle *h, int32_t which, int32_t repeats, double *ms_per_launch) {
    if (!h || !ms_per_launch || repeats <= 0) return SDV_ERR_INVALID_ARGUMENT;
    if (!h->resident) return fail(h, SDV_ERR_INVALID_ARGUMENT, "no window uploaded");
    cudaSetDevice(h->device);
    const DevProblem &P = h->P;
    cudaStream_t s = h->stream;
    // a fresh linearisation at x = 0 so that every kernel has valid inputs
    int rc = set_point(h, nullptr);
    if (rc != SDV_OK) return rc;
    k_prep_table<<<(P.F * P.C + 127) / 128, 128, 0, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, 0);
    CK(cudaMemsetAsync(h->B[1].xp, 0, sizeof(double) * P.n_pad, s));
    CK(cudaMemsetAsync(h->B[1].xl, 0, sizeof(double) * 3 * std::max(P.L, 1), s));
    k_prep_table<<<(P.F * P.C + 127) / 128, 128, 0, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, 1);
    launch_linearize(h, 0, true);
    k_ctrl_init<<<1, 1, 0, s>>>(h->d_P, h->d_st, h->d_acc, h->opt);
    CK(cudaStreamSynchronize(s));
    // which + 10: COLD variant — a 256 MiB write (twice the 126 MB L2) between launches, and for the materialising kernel the
    // two linearisation buffers alternate, so neither inputs nor outputs of the previous launch are L2-resident
    const bool cold = which >= 10;
    which = which % 10;
    constexpr size_t FLUSH_BYTES = (size_t)256 << 20;
    if (cold && !h->d_flush) CK(cudaMalloc((void **)&h->d_flush, FLUSH_BYTES));
    float total = 0;
    for (int it = 0; it < repeats + 3; it++) {
        float ms = 0;
        if (cold) CK(cudaMemsetAsync(h->d_flush, it & 0xff, FLUSH_BYTES, s));
        if (which == 0) {
            CK(cudaEventRecord(h->ev[2], s));
            h->lin_fn<<<h->lin_grid, LIN_THREADS, h->lin_smem, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_acc, cold ? (it & 1) : 0); // (x = 0 in both buffers)
            CK(cudaEventRecord(h->ev[3], s));
        } else if (which == 1) {
            CK(cudaMemsetAsync(h->d_Sb, 0, h->sb_elems * sizeof(double), s));
            if (cold) CK(cudaMemsetAsync(h->d_flush, it & 0xff, FLUSH_BYTES, s));
            CK(cudaEventRecord(h->ev[2], s));
            launch_schur(h);
            CK(cudaEventRecord(h->ev[3], s));
        } else if (which == 2) {
            CK(cudaMemsetAsync(h->d_Sb, 0, h->sb_elems * sizeof(double), s));
            launch_schur(h);
            if (P.rank == 0 && (P.P > 0 || P.has_prior || P.mp_nfull > 0))
                k_assemble_factors<<<h->fac_grid, FAC_WARPS * 32, 0, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_Sb);
            if (h->band_smem == 0) k_sysprep<<<1, 1024, 0, s>>>(h->d_P, h->d_st, h->d_acc, h->opt, h->d_Sb, h->d_scale_p, h->d_damp_p, h->d_graw_p);
            CK(cudaEventRecord(h->ev[2], s));
            {
                int rcf = launch_factor_solve(h);
                if (rcf != SDV_OK) return rcf;
            }
            CK(cudaEventRecord(h->ev[3], s));
        } else if (which == 3) {
            // back-substitution (+ candidate cost in the fused path): needs a valid reduced step, so one Schur / factor / solve first
            if (it == 0) {
                CK(cudaMemsetAsync(h->d_Sb, 0, h->sb_elems * sizeof(double), s));
                launch_schur(h);
                if (P.rank == 0 && (P.P > 0 || P.has_prior || P.mp_nfull > 0))
                    k_assemble_factors<<<h->fac_grid, FAC_WARPS * 32, 0, s>>>(h->d_P, h->B[0], h->B[1], h->d_st, h->d_Sb);
                if (h->band_smem == 0) k_sysprep<<<1, 1024, 0, s>>>(h->d_P, h->d_st, h->d_acc, h->opt, h->d_Sb, h->d_scale_p, h->d_damp_p, h->d_graw_p);
                int rcf = launch_factor_solve(h);
                if (rcf != SDV_OK) return rcf;
            }
            CK(cudaEventRecord(h->ev[2], s));
            launch_backsub(h);
            CK(cudaEventRecord(h->ev[3], s));
        } else {
            return fail(h, SDV_ERR_INVALID_ARGUMENT, "unknown kernel id");
        }
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]));
        if (it >= 3) total += ms;
    }
    *ms_per_launch = total / repeats;
    return SDV_OK;
}

// debugging / test aid: copy an internal device buffer to the host.
//   what: 0 = reduced system buffer Sb ((n_pad+32) x ld), 1 = factor Lo, 2 = dxp (n_pad), 3 = scale_p, 4 = damp_p
int sdv_debug_read(sdv_handle *h, int32_t what, double *out, int64_t count) {
    if (!h || !out || !h->resident) return SDV_ERR_INVALID_ARGUMENT;
    cudaSetDevice(h->device);
    const double *src = nullptr;
    size_t avail = 0;
    switch (what) {
    case 0: src = h->d_Sb; avail = h->sb_elems; break;
    case 1: src = h->d_Lo; avail = h->sb_elems; break;
    case 2: src = h->d_dxp; avail = h->P.n_pad; break;
    case 3: src = h->d_scale_p; avail = h->P.n_pad; break;
    case 4: src = h->d_damp_p; avail = h->P.n_pad; break;
    case 5: src = h->d_prof; avail = 8 * CC_MAX; break;
#ifdef SDV_SCHUR_PROF
    case 6:
        CK(cudaMemcpyFromSymbol(out, g_schur_prof, sizeof(double) * std::min<size_t>(32, (size_t)count)));
        return SDV_OK;
#endif
    default: return SDV_ERR_INVALID_ARGUMENT;
    }
    size_t nn = std::min<size_t>(avail, (size_t)count);
    CK(cudaMemcpy(out, src, nn * sizeof(double), cudaMemcpyDeviceToHost));
    return SDV_OK;
}

// how many times this handle captured + instantiated its whole-solve CUDA graph (tests: consecutive windows of a running
// back end must replay one graph)
int sdv_debug_graph_builds(sdv_handle *h, int64_t *count) {
    if (!h || !count) return SDV_ERR_INVALID_ARGUMENT;
    *count = h->graph_builds;
    return SDV_OK;
}

int sdv_debug_dims(sdv_handle *h, int32_t *n, int32_t *n_pad) {
    if (!h || !h->resident) return SDV_ERR_INVALID_ARGUMENT;
    if (n) *n = h->P.n;
    if (n_pad) *n_pad = h->P.n_pad;
    return SDV_OK;
}

} // extern "C"

#include "sdv_marg_host.cuh"

namespace {
void free_marg_state(sdv_handle *h) {
    if (!h->marg) return;
    if (h->marg->d) cudaFree(h->marg->d);
    if (h->marg->hp) cudaFreeHost(h->marg->hp);
    delete h->marg;
    h->marg = nullptr;
}
} // namespace

// comm.cu -- the exchange steps of ONE chain spread over the GPUs of a box: NCCL over
// NVLink / NVSwitch, one communicator per workspace, one process per GPU.
//
// The path exchanges very little: one result block per matrix of an ESS wave (all-gather, <= 1 KB per rank), the
// prior draws of the target nodes from the rank that holds their Cholesky factor (broadcast, n doubles per node)
// and a status word.  Everything is enqueued on the caller's stream, so the exchange is ordered with the kernels
// that produce and consume the data and costs no extra host synchronisation.
//
// NCCL is resolved at run time with dlopen: a process that already loaded it (torch.distributed, backend "nccl")
// hands us the same library; single-GPU runs never touch it.
#include "comm.cuh"

#include <dlfcn.h>

#include <mutex>

namespace dgpb {

namespace {

// the few NCCL declarations this file needs (ABI-stable since NCCL 2.0; nccl.h is not required to build)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
constexpr int ncclSuccess = 0;
constexpr int ncclInt32 = 2, ncclFloat64 = 8;
constexpr int ncclMax = 2;

struct Nccl {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    const char* (*GetLastError)(ncclComm_t) = nullptr;
};
Nccl g_nccl;
std::mutex g_nccl_mutex;

int load_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.handle) return DGPB_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy torch.distributed already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("NCCL is not available: %s", dlerror());
        return DGPB_CUDA_ERROR;
    }
#define DGPB_NCCL_SYM(field, name)                                                  \
    do {                                                                            \
        *(void**)(&g_nccl.field) = dlsym(h, name);                                  \
        if (!g_nccl.field) {                                                        \
            set_error("NCCL symbol %s is missing", name);                           \
            return DGPB_CUDA_ERROR;                                                 \
        }                                                                           \
    } while (0)
    DGPB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    DGPB_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    DGPB_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    DGPB_NCCL_SYM(AllGather, "ncclAllGather");
    DGPB_NCCL_SYM(Broadcast, "ncclBroadcast");
    DGPB_NCCL_SYM(AllReduce, "ncclAllReduce");
    DGPB_NCCL_SYM(GroupStart, "ncclGroupStart");
    DGPB_NCCL_SYM(GroupEnd, "ncclGroupEnd");
    DGPB_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef DGPB_NCCL_SYM
    *(void**)(&g_nccl.GetLastError) = dlsym(h, "ncclGetLastError");   // optional (NCCL >= 2.13)
    g_nccl.handle = h;
    return DGPB_OK;
}

#define DGPB_NCCL_TRY(expr)                                                                      \
    do {                                                                                         \
        int r__ = (expr);                                                                        \
        if (r__ != ncclSuccess) {                                                                \
            set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(r__)); \
            return DGPB_CUDA_ERROR;                                                              \
        }                                                                                        \
    } while (0)

}  // namespace

int comm_allgather(Workspace* ws, const double* send, double* recv, size_t count, cudaStream_t st) {
    DGPB_REQUIRE(ws->comm.world > 1 && ws->comm.nccl, "no communicator");
    DGPB_NCCL_TRY(g_nccl.AllGather(send, recv, count, ncclFloat64, (ncclComm_t)ws->comm.nccl, st));
    return DGPB_OK;
}

int comm_bcast_rows(Workspace* ws, double* base, int64_t row_elems, const int* root_of_row, int rows, cudaStream_t st) {
    DGPB_REQUIRE(ws->comm.world > 1 && ws->comm.nccl, "no communicator");
    int any = 0;
    for (int k = 0; k < rows; ++k) any += root_of_row[k] >= 0;
    if (!any) return DGPB_OK;
    DGPB_NCCL_TRY(g_nccl.GroupStart());
    for (int k = 0; k < rows; ++k) {
        if (root_of_row[k] < 0) continue;   // replicated row: every rank computed it
        double* p = base + (int64_t)k * row_elems;
        int r = g_nccl.Broadcast(p, p, (size_t)row_elems, ncclFloat64, root_of_row[k], (ncclComm_t)ws->comm.nccl, st);
        if (r != ncclSuccess) {
            g_nccl.GroupEnd();
            set_error("ncclBroadcast failed: %s", g_nccl.GetErrorString(r));
            return DGPB_CUDA_ERROR;
        }
    }
    DGPB_NCCL_TRY(g_nccl.GroupEnd());
    return DGPB_OK;
}

int comm_max_flag(Workspace* ws, int local_flag, int* global_flag, cudaStream_t st) {
    *global_flag = local_flag;
    if (ws->comm.world <= 1) return DGPB_OK;
    void* p;
    DGPB_TRY(ws->reserve(SLOT_COMM_FLAG, 2 * sizeof(int), &p));
    int* d = (int*)p;
    int* h = reinterpret_cast<int*>(ws->pinned + kPinnedFlag);
    h[0] = local_flag;
    DGPB_CUDA_TRY(cudaMemcpyAsync(d, h, sizeof(int), cudaMemcpyHostToDevice, st));
    DGPB_NCCL_TRY(g_nccl.AllReduce(d, d + 1, 1, ncclInt32, ncclMax, (ncclComm_t)ws->comm.nccl, st));
    DGPB_CUDA_TRY(cudaMemcpyAsync(h + 1, d + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    DGPB_CUDA_TRY(cudaStreamSynchronize(st));
    *global_flag = h[1];
    return DGPB_OK;
}

}  // namespace dgpb

using namespace dgpb;

extern "C" {

int dgpb_comm_unique_id(char* id128_host) {
    DGPB_REQUIRE(id128_host != nullptr, "NULL argument");
    DGPB_TRY(load_nccl());
    ncclUniqueId id;
    DGPB_NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128_host, id.internal, 128);
    return DGPB_OK;
}

int dgpb_comm_init(dgpb_ws* ws, int rank, int world, const char* id128_host) {
    DGPB_REQUIRE(ws && id128_host, "NULL argument");
    DGPB_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "bad rank / world size");
    if (ws->comm.nccl) {
        g_nccl.CommDestroy((ncclComm_t)ws->comm.nccl);
        ws->comm.nccl = nullptr;
    }
    ws->comm.rank = 0;
    ws->comm.world = 1;
    ws->owner.clear();
    if (world == 1) return DGPB_OK;
    DGPB_TRY(load_nccl());
    DGPB_CUDA_TRY(cudaSetDevice(ws->device));
    ncclUniqueId id;
    memcpy(id.internal, id128_host, 128);
    ncclComm_t c = nullptr;
    DGPB_NCCL_TRY(g_nccl.CommInitRank(&c, world, id, rank));
    ws->comm.nccl = c;
    ws->comm.rank = rank;
    ws->comm.world = world;
    return DGPB_OK;
}

int dgpb_comm_destroy(dgpb_ws* ws) {
    DGPB_REQUIRE(ws != nullptr, "NULL argument");
    if (ws->comm.nccl) {
        g_nccl.CommDestroy((ncclComm_t)ws->comm.nccl);
        ws->comm.nccl = nullptr;
    }
    ws->comm.rank = 0;
    ws->comm.world = 1;
    ws->owner.clear();
    return DGPB_OK;
}

int dgpb_comm_info(const dgpb_ws* ws, int* rank_host, int* world_host) {
    DGPB_REQUIRE(ws && rank_host && world_host, "NULL argument");
    *rank_host = ws->comm.rank;
    *world_host = ws->comm.world;
    return DGPB_OK;
}

// Which rank evaluates which item of an ESS wave, and the angles of the wave -- the host-side plan that every rank
// computes identically (no GPU needed: the multi-process tests call it on CPU).
//   theta0, tmin, tmax : first angle of the wave and the current bracket (imputation.py:81-82, 111-119)
//   u_host[0..nu_left) : the uniforms that the rejections of this wave would consume, in order
//   first              : 1 when the threshold likelihood rides in the wave as item 0, else 0
//   cap                : candidate angles per rank
// thetas_host[s] (s < *n_cand_host) = angle of candidate s under the assumption that candidates 0..s-1 were
// rejected; rank_host[i] / slot_host[i] for item i = first + s (and item 0 = the threshold when first = 1).
int dgpb_ess_plan_wave(double theta0, double tmin, double tmax, const double* u_host, int nu_left, int first, int cap,
                       int world, int* n_cand_host, double* thetas_host, int* rank_host, int* slot_host) {
    DGPB_REQUIRE(u_host && thetas_host && n_cand_host && cap >= 1 && world >= 1 && nu_left >= 0 &&
                     (first == 0 || first == 1), "bad argument");
    const int S = ess_plan_wave(theta0, tmin, tmax, u_host, nu_left, cap * world, thetas_host);
    for (int i = 0; i < first + S; ++i) {
        if (rank_host) rank_host[i] = i % world;
        if (slot_host) slot_host[i] = i / world;
    }
    *n_cand_host = S;
    return DGPB_OK;
}

}  // extern "C"

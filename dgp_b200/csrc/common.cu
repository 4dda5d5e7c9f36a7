// common.cu -- error state, workspace lifetime, descriptor helpers.
#include "common.cuh"

namespace dgpb {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int make_kernel_dev(const dgpb_node* node, int64_t n, const double* src_override, KernelDev* out) {
    DGPB_REQUIRE(node != nullptr, "node is NULL");
    DGPB_REQUIRE(node->kind == DGPB_SEXP || node->kind == DGPB_MATERN25, "unknown kernel kind");
    const int D = node->n_local + node->n_global;
    DGPB_REQUIRE(node->n_local >= 0 && node->n_global >= 0 && D >= 1 && D <= kMaxDim, "node dimension out of range");
    DGPB_REQUIRE(node->nlen == 1 || node->nlen == D, "len(length) must be 1 or D");
    const double* src = src_override ? src_override : node->src;
    DGPB_REQUIRE(src != nullptr, "node->src is NULL");
    DGPB_REQUIRE(node->n_global == 0 || node->gsrc != nullptr, "node->gsrc is NULL but connect is set");
    KernelDev k;
    k.kind = node->kind;
    k.D = D;
    k.ard = node->nlen != 1;
    k.stride = 1;
    k.nugget = node->nugget;
    for (int d = 0; d < kMaxDim; ++d) {
        k.ptr[d] = nullptr;
        k.len[d] = 1.0;
    }
    for (int d = 0; d < node->n_local; ++d) k.ptr[d] = src + (int64_t)node->input_dim[d] * n;
    for (int d = 0; d < node->n_global; ++d) k.ptr[node->n_local + d] = node->gsrc + (int64_t)node->connect[d] * n;
    for (int d = 0; d < D; ++d) k.len[d] = node->length[k.ard ? d : 0];
    *out = k;
    return DGPB_OK;
}

int make_kernel_dev_rowmajor(const double* X, int64_t D, const double* length_host, int64_t nlen, double nugget,
                             int kind, KernelDev* out) {
    DGPB_REQUIRE(X != nullptr && length_host != nullptr, "NULL pointer");
    DGPB_REQUIRE(kind == DGPB_SEXP || kind == DGPB_MATERN25, "unknown kernel kind");
    DGPB_REQUIRE(D >= 1 && D <= kMaxDim, "dimension out of range");
    DGPB_REQUIRE(nlen == 1 || nlen == D, "len(length) must be 1 or D");
    KernelDev k;
    k.kind = kind;
    k.D = (int)D;
    k.ard = nlen != 1;
    k.stride = D;
    k.nugget = nugget;
    for (int d = 0; d < kMaxDim; ++d) {
        k.ptr[d] = nullptr;
        k.len[d] = 1.0;
    }
    for (int d = 0; d < D; ++d) {
        k.ptr[d] = X + d;
        k.len[d] = length_host[k.ard ? d : 0];
    }
    *out = k;
    return DGPB_OK;
}

}  // namespace dgpb

using namespace dgpb;


extern "C" {

const char* dgpb_last_error(void) { return dgpb::g_err; }
int dgpb_version(void) { return 100; }
int64_t dgpb_sizeof_node(void) { return (int64_t)sizeof(dgpb_node); }
int64_t dgpb_launch_count(void) { return (int64_t)dgpb::g_launches.load(); }

int dgpb_ws_create(dgpb_ws** ws, int device) {
    DGPB_REQUIRE(ws != nullptr, "ws is NULL");
    int count = 0;
    DGPB_CUDA_TRY(cudaGetDeviceCount(&count));
    DGPB_REQUIRE(device >= 0 && device < count, "no such CUDA device");
    DGPB_CUDA_TRY(cudaSetDevice(device));
    dgpb_ws* w = new dgpb_ws();
    w->device = device;
    DGPB_CUDA_TRY(cudaMallocHost((void**)&w->pinned, kPinnedDoubles * sizeof(double)));
    *ws = w;
    return DGPB_OK;
}

int dgpb_ws_destroy(dgpb_ws* ws) {
    if (!ws) return DGPB_OK;
    for (int i = 0; i < SLOT_COUNT; ++i)
        if (ws->buf[i]) cudaFree(ws->buf[i]);
    for (auto& kv : ws->cache)
        if (kv.second.T) cudaFree(kv.second.T);
    for (int c = 0; c < 2; ++c) {
        if (ws->wave_stream[c]) cudaStreamDestroy(ws->wave_stream[c]);
        if (ws->wave_assembled[c]) cudaEventDestroy(ws->wave_assembled[c]);
        if (ws->wave_done[c]) cudaEventDestroy(ws->wave_done[c]);
    }
    if (ws->pinned) cudaFreeHost(ws->pinned);
    delete ws;
    return DGPB_OK;
}

int64_t dgpb_ws_bytes(const dgpb_ws* ws) {
    if (!ws) return 0;
    size_t t = ws->total();
    for (const auto& kv : ws->cache) t += kv.second.cap * sizeof(double);
    return (int64_t)t;
}

int dgpb_cache_clear(dgpb_ws* ws) {
    DGPB_REQUIRE(ws != nullptr, "ws is NULL");
    for (auto& kv : ws->cache) kv.second.valid = false;
    ws->owner.clear();
    return DGPB_OK;
}

}  // extern "C"

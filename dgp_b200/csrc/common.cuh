// common.cuh -- shared declarations for libdgpb.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/dgpb.h"

namespace dgpb {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#define DGPB_CUDA_TRY(expr)                                                                     \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            dgpb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                  \
                            cudaGetErrorString(e__));                                           \
            return DGPB_CUDA_ERROR;                                                             \
        }                                                                                       \
    } while (0)

// after every kernel launch: count it and surface launch-configuration errors
#define DGPB_LAUNCHED()                                                                         \
    do {                                                                                        \
        dgpb::g_launches.fetch_add(1, std::memory_order_relaxed);                               \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess) {                                                               \
            dgpb::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,              \
                            cudaGetErrorString(e__));                                           \
            return DGPB_CUDA_ERROR;                                                             \
        }                                                                                       \
    } while (0)

#define DGPB_TRY(expr)                                                                          \
    do {                                                                                        \
        int s__ = (expr);                                                                       \
        if (s__ != DGPB_OK) return s__;                                                         \
    } while (0)

#define DGPB_REQUIRE(cond, msg)                                                                 \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            dgpb::set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, msg);                \
            return DGPB_BAD_ARG;                                                                \
        }                                                                                       \
    } while (0)

// layout of the small device result buffer (SLOT_OUT), in doubles:
//   [0, 128)    dense batch results, 4 per matrix (logdet, quad, sigma2, -)
//   [256, 512)  Vecchia per-node results, 2 per node (quad, logdet)
//   [512, 1024) gradient results (nllik, sigma2, grad[P])
constexpr int kOutDoubles = 1024;
constexpr int kOutVecchia = 256;
constexpr int kOutGrad = 512;

constexpr double kSqrt5 = 2.2360679774997896964;
constexpr int kMaxDim = DGPB_MAX_DIM;

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t cdiv(int64_t x, int64_t m) { return (x + m - 1) / m; }

// ---------------------------------------------------------------------------------------------
// Workspace: slot-addressed device scratch that only grows.  No allocation on the hot path after
// warm-up (SURVEY.md 8b "ownership").
// ---------------------------------------------------------------------------------------------
enum Slot : int {
    SLOT_T = 0,       // factorisation matrices (batched)
    SLOT_T2,          // second set: the next ESS wave is assembled (and factored) while the current one is factored
    SLOT_DIAG,        // diag(L) per batch entry
    SLOT_DIAG2,       //   ... of the second set
    SLOT_OUT,         // small result scalars
    SLOT_OUT2,        //   ... of the second set
    SLOT_INFO,        // int info flags
    SLOT_INFO2,       //   ... of the second set
    SLOT_PART,        // reduction partials
    SLOT_NU,          // ESS prior draws
    SLOT_PROP,        // ESS proposal layer
    SLOT_GEMM_A,      // prediction: cross-kernel matrix R
    SLOT_GEMM_B,      // prediction: padded R^-1
    SLOT_GEMM_C,      // prediction: R * R^-1
    SLOT_MISC,
    SLOT_MISC2,
    SLOT_VX,          // Vecchia: inputs gathered in Vecchia order
    SLOT_VY,          // Vecchia: outputs gathered in Vecchia order
    SLOT_VL,          // Vecchia: sparse inverse-Cholesky rows
    SLOT_VFLAG,       // Vecchia: ready flags of the sparse solve
    SLOT_KNN_PACK,    // neighbour search: candidates packed as TF32 screen rows (tcgen05 path)
    SLOT_COMM,        // multi-GPU: gathered result blocks of an ESS wave (world x kCommBlock doubles)
    SLOT_COMM_FLAG,   // multi-GPU: status word
    SLOT_COUNT
};

// NVTX range of one sub-system call (tracing); costs nothing without a profiler attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define DGPB_NVTX(name) dgpb::NvtxRange nvtx_range__(name)

// ---------------------------------------------------------------------------------------------
// Multi-GPU (comm.cu): one chain spread over the GPUs of a box.  Every rank holds the same latent layers and
// takes the same decisions; what is partitioned is the WORK -- the candidate angles of an ESS wave -- and the
// Cholesky factors that work leaves behind, which stay on the rank that computed them.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 8;
constexpr int kCommBlock = 128;       // doubles per rank in the wave all-gather (4 per matrix, <= 32 matrices)
// layout of the pinned staging buffer (doubles)
constexpr int kPinnedDoubles = 8192;
constexpr int kPinnedVecchia = 1024;  // Vecchia per-node results, cached-threshold quads
constexpr int kPinnedInfo = 2048;     // int info flags of a batch
constexpr int kPinnedFlag = 3072;     // status word of comm_max_flag
constexpr int kPinnedWave = 4096;     // gathered wave results (kMaxRanks x kCommBlock), one region per wave context
constexpr int kPinnedWaveStride = 1024;

struct Comm {
    void* nccl = nullptr;   // ncclComm_t
    int rank = 0, world = 1;
};
// which rank holds chol(K) stored under a cache key (replicated on every rank: decisions must agree)
constexpr int kOwnerNone = -1, kOwnerAll = -2;
struct FactorOwner {
    int rank = kOwnerNone;
    int64_t n = 0;
    bool has_logdet = false;
    double logdet = 0.0;
    bool w_synced = false;   // see CachedFactor
};

// chol(K) of a GP node kept across ESS block updates of one I-step (hyper-parameters fixed)
struct CachedFactor {
    double* T = nullptr;   // complete lower factor (diagonal blocks restored), Geom(n, false) layout
    size_t cap = 0;        // doubles allocated
    int64_t n = 0;
    bool valid = false;
    double logdet = 0.0;   // log|K| of the stored factor (valid when has_logdet)
    bool has_logdet = false;
    bool w_synced = false; // row npad of T holds L^-1 y for the node's CURRENT output (kept in step by rotate_cached_w)
};

struct Workspace {
    int device = 0;
    std::map<int, CachedFactor> cache;
    Comm comm;
    std::map<int, FactorOwner> owner;   // world > 1 only
    // ESS wave pipeline (ess.cu): two factorisation contexts (T set, side buffers, streams) so that the next wave can
    // be assembled and factored while the current one drains its tail
    cudaStream_t wave_stream[2] = {nullptr, nullptr};
    cudaEvent_t wave_assembled[2] = {nullptr, nullptr}, wave_done[2] = {nullptr, nullptr};
    int wave_cur = 0;                   // the context a new block update starts with (the other may still be busy)
    // proposals per block update of a layer pair (keyed by its first upper node): running mean / spread, used to
    // factor waves ahead only while an acceptance is still unlikely
    struct PropStat { double mean = 0.0, m2 = 0.0; int cnt = 0; };
    std::map<int, PropStat> prop_stats;
    void* buf[SLOT_COUNT] = {};
    size_t cap[SLOT_COUNT] = {};
    double* pinned = nullptr;  // small pinned host staging buffer (kPinnedDoubles)

    int reserve(int slot, size_t bytes, void** out) {
        if (bytes > cap[slot]) {
            if (buf[slot]) DGPB_CUDA_TRY(cudaFree(buf[slot]));
            buf[slot] = nullptr;
            cap[slot] = 0;
            size_t want = bytes + bytes / 8 + 256;
            DGPB_CUDA_TRY(cudaMalloc(&buf[slot], want));
            cap[slot] = want;
        }
        *out = buf[slot];
        return DGPB_OK;
    }
    size_t total() const {
        size_t t = 0;
        for (int i = 0; i < SLOT_COUNT; ++i) t += cap[i];
        return t;
    }
};

// ---------------------------------------------------------------------------------------------
// Device-side description of a node's kernel function, passed to kernels BY VALUE.
// The input of point i is gathered straight from the variable-major sources:
//   x_d(i) = ptr[d][i] * inv_len[d]   -- this is `X/self.length` (kernel_class.py:324) without
// materialising per-node input copies.
// ---------------------------------------------------------------------------------------------
struct KernelDev {
    int kind;
    int D;
    int ard;                     // 1: one length per dimension
    int64_t stride;              // element stride between consecutive points (1 = variable-major)
    const double* ptr[kMaxDim];  // base of dimension d: x_d(i) = ptr[d][i*stride] / len[d]
    double len[kMaxDim];         // length-scale per dimension (shared value replicated)
    double nugget;
    __device__ __forceinline__ double x(int d, int64_t i) const { return ptr[d][i * stride] / len[d]; }
    __device__ __forceinline__ double raw(int d, int64_t i) const { return ptr[d][i * stride]; }
};

// scaled coordinate: the reference divides (X / length), it does not multiply by a reciprocal
__device__ __forceinline__ double scaled(double x, double len) { return x / len; }

// Correlation between two scaled points given as register arrays / pointers with stride.
// sexp   : exp(-sum_d (a_d-b_d)^2)                                  kernel_class.py:326-327
// matern : prod_d(1+sqrt5 r+5/3 r^2) * exp(-sqrt5 sum_d r)          kernel_class.py:343-345
template <typename FA, typename FB>
__device__ __forceinline__ double corr_pair(int kind, int D, FA a, FB b) {
    if (kind == DGPB_SEXP) {
        double dist = 0.0;
        for (int d = 0; d < D; ++d) {
            double df = a(d) - b(d);
            dist = __dadd_rn(dist, __dmul_rn(df, df));
        }
        return exp(-dist);
    } else {
        double coef = 1.0, s = 0.0;
        for (int d = 0; d < D; ++d) {
            double r = fabs(a(d) - b(d));
            coef *= 1.0 + kSqrt5 * r + (5.0 / 3.0) * (r * r);
            s += r;
        }
        return coef * exp(-kSqrt5 * s);
    }
}

// exp(x) for x <= ~0 in 18 instructions (the library routine is ~35): k = round(x log2 e) by the magic-number
// trick, two-step Cody-Waite reduction, degree-11 Taylor polynomial on |r| <= ln2/2 (truncation 6e-15 relative),
// scaling through the exponent field.  Arguments below -700 return 0 (the library would give a subnormal < 1e-304).
__device__ __forceinline__ double exp_nonpos(double x) {
    const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);
    const int k = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, -6.93147180369123816490e-01, x);
    r = fma(kf, -1.90821492927058770002e-10, r);
    double p = 2.50521083854417187751e-08;            // 1/11!
    p = fma(p, r, 2.75573192239858906526e-07);        // 1/10!
    p = fma(p, r, 2.75573192239858906526e-06);        // 1/9!
    p = fma(p, r, 2.48015873015873015873e-05);        // 1/8!
    p = fma(p, r, 1.98412698412698412698e-04);        // 1/7!
    p = fma(p, r, 1.38888888888888888889e-03);        // 1/6!
    p = fma(p, r, 8.33333333333333333333e-03);        // 1/5!
    p = fma(p, r, 4.16666666666666666667e-02);        // 1/4!
    p = fma(p, r, 1.66666666666666666667e-01);        // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double s = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    return x < -700.0 ? 0.0 : s;
}

// linear index t over the lower triangle (row-major: 0 -> (0,0), 1 -> (1,0), 2 -> (1,1), ...) -> (i, j).
// FP32 sqrt + integer correction: the FP64 sqrt competed with the DMMA stream for the FP64 pipe and was 15%
// of the trailing-update kernel's stall samples (profiles/r1_update_v2_ncu.md).
__device__ __forceinline__ void tri_index(int t, int& i, int& j) {
    i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    while (i * (i + 1) / 2 > t) --i;
    j = t - i * (i + 1) / 2;
}
// linear index over the "double-width" staircase: row i owns 2 i + 2 entries -> i (i + 1) precede it
__device__ __forceinline__ void stair_index(int t, int& i, int& j) {
    i = (int)((sqrtf(4.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) <= t) ++i;
    while (i * (i + 1) > t) --i;
    j = t - i * (i + 1);
}

// block-wide deterministic sum (fixed tree order); result valid in thread 0 (and broadcast via smem)
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* sred) {
    const int tid = threadIdx.x;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) sred[tid >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (tid < 32) {
        r = (tid < THREADS / 32) ? sred[tid] : 0.0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;
}

// host helpers implemented in common.cu
// node -> KernelDev; `src_override` (if not NULL) replaces node->src (ESS proposals)
int make_kernel_dev(const dgpb_node* node, int64_t n, const double* src_override, KernelDev* out);
// row-major X (n x D) -> KernelDev
int make_kernel_dev_rowmajor(const double* X, int64_t D, const double* length_host, int64_t nlen,
                             double nugget, int kind, KernelDev* out);

// predict.cu: 1 = link_gp sexp exponents on the FP64 tensor path (default), 0 = vector-pipe pair kernel
int linkgp_set_mma(int on);
// predict.cu: 1 = tabulated Matern-2.5 link_gp kernel (default), 0 = direct closed form per pair
int linkgp_set_matern_tab(int on);

}  // namespace dgpb

// the opaque C handle is the workspace itself
struct dgpb_ws : public dgpb::Workspace {};

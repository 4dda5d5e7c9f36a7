// ess_small.cu -- the whole I-step of a SMALL dense DGP (n <= 64 training points) in ONE kernel launch.
//
// imputer.sample(burnin) (dgpsi/imputation.py:22-119, block updates) is a strictly sequential chain: burnin + 1 sweeps
// over the layer pairs, per pair one prior draw per target node (fmvn, functions.py:113-121), one threshold and a
// data-dependent number of proposals.  At n = 10 (BASELINE config 1, demo/step_fct.ipynb) a proposal is a few
// microseconds of arithmetic, so the general path -- a dozen launches and one host synchronisation per wave --
// is bound by launch latency and loses to the CPU.  Here one CTA keeps the layers, the prior draws and the kernel
// matrix in shared memory and runs propose -> kernel matrix -> Cholesky -> solve -> accept / shrink for every block
// update of every sweep with the caller's pre-drawn normals and uniforms; the host sees the result once.
// Decisions follow the reference's rule with the same uniforms in the same order (threshold, first angle, one per
// rejection); the consumed count is returned so the caller can advance its generator by exactly that much.
#include "dense.cuh"

namespace dgpb {

constexpr int kSmallMaxN = 64;
constexpr int kSmallMaxW = 8;    // nodes per layer
constexpr int kSmallMaxL = 8;    // GP layers

struct SmallNode {
    int kind, n_local, n_global, local_from_layer;   // local inputs: rows of layer l - 1 (1) or of the node's own `src` (0)
    int input_dim[kMaxDim];
    int connect[kMaxDim];
    double len[kMaxDim];      // per input dimension (a shared length-scale is replicated)
    double scale, nugget;
    const double* src;        // first layer: the node's own input, variable-major (n_src x n)
    const double* gsrc;       // global input, variable-major
};

struct SmallArgs {
    int n, L, sweeps, nu;
    int width[kSmallMaxL];
    double* layer[kSmallMaxL];          // layer l: width_l x n (variable-major); the last one holds the training outputs
    SmallNode node[kSmallMaxL][kSmallMaxW];
    const double* z;                    // standard normals, n per prior draw, consumed in order
    const double* u;                    // uniforms, consumed in order
    int* counts;                        // [0] status (0 ok, 1 not PD, 2 out of uniforms), [1] uniforms used,
                                        // [2] proposals evaluated, [3] prior draws made
};

// scaled coordinates of the node's inputs into xs[d][i]; `img` = image of the feeding layer (width x n) in shared memory
__device__ void small_coords(const SmallNode& nd, const double* img, int n, double (*xs)[kSmallMaxN]) {
    const int D = nd.n_local + nd.n_global;
    for (int idx = threadIdx.x; idx < D * n; idx += blockDim.x) {
        const int d = idx / n, i = idx - d * n;
        double v;
        if (d < nd.n_local)
            v = nd.local_from_layer ? img[nd.input_dim[d] * n + i] : nd.src[(int64_t)nd.input_dim[d] * n + i];
        else
            v = nd.gsrc[(int64_t)nd.connect[d - nd.n_local] * n + i];
        xs[d][i] = v / nd.len[d];      // X / length (kernel_class.py:324): a division, like the reference
    }
    __syncthreads();
}

// lower triangle of K = corr(x_i, x_j) + nugget I into sK (row stride kSmallMaxN + 1)
__device__ void small_kmatrix(const SmallNode& nd, int n, double (*xs)[kSmallMaxN], double* sK) {
    const int D = nd.n_local + nd.n_global;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int i = idx / n, j = idx - i * n;
        if (j > i) continue;
        double v;
        if (i == j) {
            v = 1.0 + nd.nugget;
        } else if (nd.kind == DGPB_SEXP) {
            double dist = 0.0;
            for (int d = 0; d < D; ++d) {
                const double df = xs[d][i] - xs[d][j];
                dist = __dadd_rn(dist, __dmul_rn(df, df));
            }
            v = exp_nonpos(-dist);
        } else {
            double coef = 1.0, s = 0.0;
            for (int d = 0; d < D; ++d) {
                const double r = fabs(xs[d][i] - xs[d][j]);
                coef *= 1.0 + kSqrt5 * r + (5.0 / 3.0) * (r * r);
                s += r;
            }
            v = coef * exp_nonpos(-kSqrt5 * s);
        }
        sK[i * (kSmallMaxN + 1) + j] = v;
    }
    __syncthreads();
}

// in-place lower Cholesky of sK; returns false (for every thread) on a non-positive pivot
__device__ bool small_cholesky(int n, double* sK, int* sflag) {
    constexpr int LD = kSmallMaxN + 1;
    for (int j = 0; j < n; ++j) {
        if (threadIdx.x == 0) {
            const double d = sK[j * LD + j];
            if (!(d > 0.0)) *sflag = 1;
            sK[j * LD + j] = sqrt(d);
        }
        __syncthreads();
        if (*sflag) return false;
        const double inv = 1.0 / sK[j * LD + j];
        for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) sK[i * LD + j] *= inv;
        __syncthreads();
        const int m = n - j - 1;
        for (int idx = threadIdx.x; idx < m * m; idx += blockDim.x) {
            const int a = idx / m, b = idx - a * m;
            if (b > a) continue;
            const int i = j + 1 + a, k = j + 1 + b;
            sK[i * LD + k] = fma(-sK[i * LD + j], sK[k * LD + j], sK[i * LD + k]);
        }
        __syncthreads();
    }
    return true;
}

// -0.5 (log|scale K| + y'(scale K)^-1 y) from the factor in sK (kernel_class.py:482-488); warp 0 solves, all return it
__device__ double small_loglik_from_factor(int n, const double* sK, const double* y, double scale, double* sw, double* sres) {
    constexpr int LD = kSmallMaxN + 1;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        double acc0 = lane < n ? y[lane] : 0.0, acc1 = lane + 32 < n ? y[lane + 32] : 0.0;   // rows lane, lane + 32
        double quad = 0.0, logdet = 0.0;
        for (int j = 0; j < n; ++j) {
            const double bj = __shfl_sync(0xffffffffu, j < 32 ? acc0 : acc1, j & 31);
            const double djj = sK[j * LD + j];
            const double wj = bj / djj;
            quad = fma(wj, wj, quad);
            logdet += log(djj);
            if (lane > j && lane < n) acc0 = fma(-sK[lane * LD + j], wj, acc0);
            if (lane + 32 > j && lane + 32 < n) acc1 = fma(-sK[(lane + 32) * LD + j], wj, acc1);
        }
        if (lane == 0) *sres = -0.5 * (2.0 * logdet + (double)n * log(scale) + quad / scale);
    }
    __syncthreads();
    const double r = *sres;
    __syncthreads();
    (void)sw;
    return r;
}

__global__ void __launch_bounds__(256, 1) ess_small_kernel(const SmallArgs* __restrict__ ap) {
    extern __shared__ double sm[];
    const SmallArgs& a = *ap;
    const int n = a.n, tid = threadIdx.x;
    double* sK = sm;                                                   // 64 x 65
    double (*xs)[kSmallMaxN] = reinterpret_cast<double (*)[kSmallMaxN]>(sK + kSmallMaxN * (kSmallMaxN + 1));   // 32 x 64
    double* cur = reinterpret_cast<double*>(xs) + kMaxDim * kSmallMaxN;   // current image of the layer being updated (W x 64)
    double* prp = cur + kSmallMaxW * kSmallMaxN;                       // proposal image
    double* nu = prp + kSmallMaxW * kSmallMaxN;                        // prior draws
    double* sres = nu + kSmallMaxW * kSmallMaxN;                       // scalar result
    __shared__ int sflag;
    __shared__ int s_ui, s_nprop, s_draws, s_status;
    if (tid == 0) {
        sflag = 0;
        s_ui = s_nprop = s_draws = s_status = 0;
    }
    __syncthreads();
    constexpr int LD = kSmallMaxN + 1;
    for (int sweep = 0; sweep < a.sweeps && s_status == 0; ++sweep) {
        for (int l = 0; l + 1 < a.L && s_status == 0; ++l) {
            const int Wt = a.width[l], Wu = a.width[l + 1];
            // ---- current layer image and the prior draws nu_k = chol(scale_k K_k) z_k of its nodes
            for (int idx = tid; idx < Wt * n; idx += blockDim.x) cur[idx] = a.layer[l][idx];
            __syncthreads();
            for (int k = 0; k < Wt; ++k) {
                const SmallNode& nd = a.node[l][k];
                // inputs of a target: the layer below (already in global memory) or the node's own src
                small_coords(nd, l > 0 ? a.layer[l - 1] : nullptr, n, xs);
                small_kmatrix(nd, n, xs, sK);
                if (!small_cholesky(n, sK, &sflag)) {
                    if (tid == 0) s_status = 1;
                    __syncthreads();
                    break;
                }
                const double* zk = a.z + (int64_t)s_draws * n;
                if (tid < n) {
                    double s = 0.0;
                    for (int j = 0; j <= tid; ++j) s = fma(sK[tid * LD + j], zk[j], s);
                    nu[k * n + tid] = sqrt(nd.scale) * s;
                }
                __syncthreads();
                if (tid == 0) ++s_draws;
                __syncthreads();
            }
            if (s_status) break;
            // ---- sum of the upper log-likelihoods for a given image of layer l
            auto upper_sum = [&](const double* img, bool* ok) -> double {
                double total = 0.0;
                *ok = true;
                for (int u = 0; u < Wu; ++u) {
                    const SmallNode& nd = a.node[l + 1][u];
                    small_coords(nd, img, n, xs);
                    small_kmatrix(nd, n, xs, sK);
                    if (!small_cholesky(n, sK, &sflag)) {
                        *ok = false;
                        return 0.0;
                    }
                    total += small_loglik_from_factor(n, sK, a.layer[l + 1] + (int64_t)u * n, nd.scale, nullptr, sres);
                }
                return total;
            };
            bool ok;
            if (s_ui + 2 > a.nu) {
                if (tid == 0) s_status = 2;
                __syncthreads();
                break;
            }
            double log_y = upper_sum(cur, &ok);
            if (!ok) {
                if (tid == 0) s_status = 1;
                __syncthreads();
                break;
            }
            int ui = s_ui;
            log_y += log(a.u[ui++]);                          // imputation.py:79
            double theta = 2.0 * M_PI * a.u[ui++];            // imputation.py:81
            double tmin = theta - 2.0 * M_PI, tmax = theta;
            int np_local = 0;
            while (true) {
                const double c = cos(theta), s = sin(theta);
                for (int idx = tid; idx < Wt * n; idx += blockDim.x)
                    prp[idx] = __dadd_rn(__dmul_rn(cur[idx], c), __dmul_rn(nu[idx], s));   // update_f, functions.py:203-208
                __syncthreads();
                const double ll = upper_sum(prp, &ok);
                ++np_local;
                if (!ok) {
                    if (tid == 0) s_status = 1;
                    break;
                }
                if (ll > log_y) {                              // imputation.py:107-110
                    for (int idx = tid; idx < Wt * n; idx += blockDim.x) a.layer[l][idx] = prp[idx];
                    break;
                }
                if (theta < 0.0) tmin = theta; else tmax = theta;   // imputation.py:115-118
                if (ui >= a.nu) {
                    if (tid == 0) s_status = 2;
                    break;
                }
                theta = tmin + (tmax - tmin) * a.u[ui++];      // imputation.py:119
            }
            __syncthreads();
            if (tid == 0) {
                s_ui = ui;
                s_nprop += np_local;
            }
            __threadfence_block();
            __syncthreads();
        }
    }
    if (tid == 0) {
        a.counts[0] = s_status;
        a.counts[1] = s_ui;
        a.counts[2] = s_nprop;
        a.counts[3] = s_draws;
    }
}

constexpr size_t kSmallSmem = sizeof(double) * ((size_t)kSmallMaxN * (kSmallMaxN + 1) + (size_t)kMaxDim * kSmallMaxN +
                                                3 * (size_t)kSmallMaxW * kSmallMaxN + 8);

}  // namespace dgpb

using namespace dgpb;

// burnin + 1 block-update sweeps over a dense DGP with n <= 64 training points and at most 8 nodes per layer, in ONE
// launch.  nodes_host: the GP nodes of all layers, layer by layer (widths_host[l] of them); layer_ptrs_host[l]: device
// image of layer l (width_l x n, variable-major; every node of layer l + 1 must read its local inputs from it; the last
// layer holds the training outputs).  z: device, one row of n standard normals per prior draw in the reference's order
// (sweep, layer pair, target node); u_host: uniforms in the reference's order.  counts_host = {uniforms consumed,
// proposals evaluated, prior draws made}.  DGPB_BAD_ARG if the uniforms ran out (counts are still filled in).
extern "C" int dgpb_ess_sweeps_small(dgpb_ws* ws, const dgpb_node* nodes_host, const int32_t* widths_host, int n_layers,
                                     double* const* layer_ptrs_host, int64_t n, int sweeps, const double* z,
                                     int64_t z_rows, const double* u_host, int nu, int32_t* counts_host, void* stream) {
    DGPB_NVTX("dgpb:ess_sweeps_small");
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(ws && nodes_host && widths_host && layer_ptrs_host && z && u_host && counts_host, "NULL argument");
    DGPB_REQUIRE(n >= 1 && n <= kSmallMaxN && n_layers >= 2 && n_layers <= kSmallMaxL && sweeps >= 1 && nu >= 2,
                 "sizes outside the small-model kernel");
    static thread_local SmallArgs h;   // ~37 KB: not on the stack
    h.n = (int)n;
    h.L = n_layers;
    h.sweeps = sweeps;
    h.nu = nu;
    int64_t draws = 0;
    const dgpb_node* nd = nodes_host;
    for (int l = 0; l < n_layers; ++l) {
        const int w = widths_host[l];
        DGPB_REQUIRE(w >= 1 && w <= kSmallMaxW, "layer width outside the small-model kernel");
        h.width[l] = w;
        h.layer[l] = layer_ptrs_host[l];
        DGPB_REQUIRE(h.layer[l] != nullptr, "NULL layer image");
        if (l + 1 < n_layers) draws += (int64_t)w * sweeps;
        for (int k = 0; k < w; ++k, ++nd) {
            DGPB_REQUIRE(!nd->vecch, "Vecchia nodes are not handled by the small-model kernel");
            DGPB_REQUIRE(nd->kind == DGPB_SEXP || nd->kind == DGPB_MATERN25, "unknown kernel kind");
            const int D = nd->n_local + nd->n_global;
            DGPB_REQUIRE(D >= 1 && D <= kMaxDim && (nd->nlen == 1 || nd->nlen == D), "node dimension out of range");
            SmallNode& s = h.node[l][k];
            s.kind = nd->kind;
            s.n_local = nd->n_local;
            s.n_global = nd->n_global;
            s.local_from_layer = l > 0 ? 1 : 0;
            DGPB_REQUIRE(l == 0 || nd->src == layer_ptrs_host[l - 1], "a node must read the layer below it");
            DGPB_REQUIRE(nd->src != nullptr && (nd->n_global == 0 || nd->gsrc != nullptr), "NULL node input");
            for (int d = 0; d < kMaxDim; ++d) {
                s.input_dim[d] = d < nd->n_local ? nd->input_dim[d] : 0;
                s.connect[d] = d < nd->n_global ? nd->connect[d] : 0;
                s.len[d] = d < D ? nd->length[nd->nlen == 1 ? 0 : d] : 1.0;
            }
            if (l > 0)
                for (int d = 0; d < nd->n_local; ++d)
                    DGPB_REQUIRE(nd->input_dim[d] >= 0 && nd->input_dim[d] < widths_host[l - 1], "input_dim out of range");
            s.scale = nd->scale;
            s.nugget = nd->nugget;
            s.src = nd->src;
            s.gsrc = nd->gsrc;
        }
    }
    DGPB_REQUIRE(z_rows >= draws, "not enough normal draws");
    DGPB_TRY(join_waves(ws, st));
    void *pargs, *pu;
    DGPB_TRY(ws->reserve(SLOT_MISC, sizeof(SmallArgs) + 64, &pargs));
    DGPB_TRY(ws->reserve(SLOT_MISC2, sizeof(double) * (size_t)nu + 64, &pu));
    h.z = z;
    h.u = (const double*)pu;
    h.counts = reinterpret_cast<int*>((char*)pargs + ((sizeof(SmallArgs) + 15) / 16) * 16);
    static bool cfg = false;
    if (!cfg) {
        DGPB_CUDA_TRY(cudaFuncSetAttribute(ess_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmem));
        cfg = true;
    }
    DGPB_CUDA_TRY(cudaMemcpyAsync(pargs, &h, sizeof(SmallArgs), cudaMemcpyHostToDevice, st));
    DGPB_CUDA_TRY(cudaMemcpyAsync(pu, u_host, sizeof(double) * (size_t)nu, cudaMemcpyHostToDevice, st));
    ess_small_kernel<<<1, 256, kSmallSmem, st>>>((const SmallArgs*)pargs);
    DGPB_LAUNCHED();
    int* hc = reinterpret_cast<int*>(ws->pinned + kPinnedInfo);
    DGPB_CUDA_TRY(cudaMemcpyAsync(hc, h.counts, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
    DGPB_CUDA_TRY(cudaStreamSynchronize(st));
    counts_host[0] = hc[1];
    counts_host[1] = hc[2];
    counts_host[2] = hc[3];
    if (hc[0] == 1) {
        set_error("a covariance matrix of the small-model I-step is not positive definite");
        return DGPB_NOT_PD;
    }
    if (hc[0] == 2) {
        set_error("ESS ran out of uniforms after %d proposals", hc[2]);
        return DGPB_BAD_ARG;
    }
    return DGPB_OK;
}

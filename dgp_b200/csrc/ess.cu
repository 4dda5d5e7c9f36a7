// ess.cu -- elliptical slice sampling of one latent layer (Murray, Adams & MacKay) on the device.
// Restates imputer.one_sample_block / one_sample  (dgpsi/imputation.py:44-119, :166-221):
//   nu_k  = chol(scale_k K_k) z_k                      for every target node k   (fmvn / fmvn_sp)
//   log_y = sum_upper loglik(current) + log(u0)
//   theta = 2 pi u1, bracket [theta-2pi, theta]; proposal f' = f cos(theta) + nu sin(theta);
//   accept iff sum_upper loglik(f') > log_y, else shrink the bracket towards 0 and redraw.
// The latent layer, the prior draws and every proposal stay in HBM; the likelihoods of all upper
// nodes of a proposal are evaluated by ONE batched factorisation (blockIdx.z = node).  The host only
// sees one scalar per proposal (the summed log-likelihood) to take the accept/shrink decision with the
// caller's uniforms, which keeps the decision sequence identical to the reference under injected draws.
#include "comm.cuh"
#include "dense.cuh"
#include "vecchia.cuh"

namespace dgpb {

__global__ void propose_kernel(double* __restrict__ prop, const double* __restrict__ f, const double* __restrict__ nu,
                               double c, double s, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // update_f (functions.py:203-208): f*cos(theta) + nu*sin(theta), two products then one sum
    if (i < n) prop[i] = __dadd_rn(__dmul_rn(f[i], c), __dmul_rn(nu[i], s));
}

// raw node inputs and outputs gathered in Vecchia order: Xo[r][k] = x_k(ord[r]), yo[r] = y[ord[r]]
__global__ void gather_ord_kernel(KernelDev kd, const double* __restrict__ y, const int64_t* __restrict__ ord, int64_t n,
                                  double* __restrict__ Xo, double* __restrict__ yo) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int64_t i = ord[r];
    for (int d = 0; d < kd.D; ++d) Xo[r * kd.D + d] = kd.raw(d, i);
    if (yo) yo[r] = y[i];
}

// out[ord[r]] = in[r]      ( x[rev_ord] of imputation.py:61 )
__global__ void scatter_ord_kernel(const double* __restrict__ in, const int64_t* __restrict__ ord, int64_t n,
                                   double* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) out[ord[r]] = in[r];
}

static int node_lengths(const dgpb_node* nd, double* len, int* D) {
    *D = nd->n_local + nd->n_global;
    for (int d = 0; d < *D; ++d) len[d] = nd->length[nd->nlen == 1 ? 0 : d];
    return DGPB_OK;
}

// sum of the log-likelihoods of `U` nodes whose local inputs are read from `src_override`
// (or from node->src when NULL).  Dense nodes are batched; Vecchia nodes use the block kernel.
struct DenseBatchInfo {
    Batch bt;
    Geom g;
    int B = 0;           // matrices in the (single) dense batch, 0 if none or if it needed several batches
    int map[MAXB];       // batch slot -> node index
};

// ---- which rank holds chol(K) of a cache key ---------------------------------------------------------------------
// One GPU: the workspace cache itself (kOwnerAll).  Several GPUs: the replicated owner map -- every rank takes the
// same decisions from it, only the owning rank touches the factor.
static int factor_owner(Workspace* ws, int key, int64_t n, bool* has_logdet = nullptr, double* logdet = nullptr,
                        bool* w_synced = nullptr) {
    if (key < 0) return kOwnerNone;
    if (ws->comm.world <= 1) {
        auto it = ws->cache.find(key);
        if (it == ws->cache.end() || !it->second.valid || it->second.n != n) return kOwnerNone;
        if (has_logdet) *has_logdet = it->second.has_logdet;
        if (logdet) *logdet = it->second.logdet;
        if (w_synced) *w_synced = it->second.w_synced;
        return kOwnerAll;
    }
    auto it = ws->owner.find(key);
    if (it == ws->owner.end() || it->second.rank == kOwnerNone || it->second.n != n) return kOwnerNone;
    if (has_logdet) *has_logdet = it->second.has_logdet;
    if (logdet) *logdet = it->second.logdet;
    if (w_synced) *w_synced = it->second.w_synced;
    return it->second.rank;
}
static inline bool is_mine(const Workspace* ws, int owner) { return owner == kOwnerAll || owner == ws->comm.rank; }
static void set_owner(Workspace* ws, int key, int rank, int64_t n, const double* logdet) {
    if (key < 0 || ws->comm.world <= 1) return;
    FactorOwner& o = ws->owner[key];
    o.rank = rank;
    o.n = n;
    o.has_logdet = logdet != nullptr;
    o.logdet = logdet ? *logdet : 0.0;
    o.w_synced = logdet != nullptr;   // stored from a wave: the y row of T is L^-1 y for the node's output of that moment
    if (!is_mine(ws, rank)) {   // a factor this rank kept from an earlier update is stale now
        auto it = ws->cache.find(key);
        if (it != ws->cache.end()) it->second.valid = false;
    }
}

// keep chol(K) of batch slot b (just factored, diagonal blocks still in the side buffer) under `key`
static int cache_store(Workspace* ws, int key, const Geom& g, const Batch& bt, int b, cudaStream_t st,
                       const double* logdet = nullptr) {
    CachedFactor& cf = ws->cache[key];
    if (cf.cap < g.elems()) {
        if (cf.T) DGPB_CUDA_TRY(cudaFree(cf.T));
        cf.T = nullptr;
        cf.cap = 0;
        DGPB_CUDA_TRY(cudaMalloc((void**)&cf.T, g.elems() * sizeof(double)));
        cf.cap = g.elems();
    }
    DGPB_CUDA_TRY(cudaMemcpyAsync(cf.T, bt.T[b], g.elems() * sizeof(double), cudaMemcpyDeviceToDevice, st));
    Batch one;
    for (int i = 0; i < MAXB; ++i) one.T[i] = one.diag[i] = nullptr;
    one.T[0] = cf.T;
    one.diag[0] = bt.diag[b];
    one.info = bt.info;
    DGPB_TRY(restore_diag_blocks(g, one, 1, st));
    cf.n = g.n;
    cf.valid = true;
    cf.has_logdet = logdet != nullptr;
    cf.logdet = logdet ? *logdet : 0.0;
    cf.w_synced = logdet != nullptr;
    return DGPB_OK;
}

// An accepted ESS move replaces the output y of target node k by y cos(theta) + nu sin(theta) with nu = sqrt(scale) L z,
// L = chol(K_k) -- so L^-1 y, kept as row npad of the node's cached factor, becomes cos(theta) L^-1 y + sin(theta)
// sqrt(scale) z: n multiply-adds keep it in step with the output, and the next threshold that needs y'K^-1y of this
// node (it is an upper node of the layer pair below) is a dot product instead of a triangular solve (n^2).
__global__ void rotate_w_kernel(double* __restrict__ w, const double* __restrict__ z, double c, double s_sq, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = fma(c, w[i], s_sq * z[i]);
}
static int rotate_cached_w(Workspace* ws, const dgpb_node* targets, int M, const int32_t* keys, int64_t n, const double* z,
                           double theta, cudaStream_t st) {
    if (!keys) return DGPB_OK;
    const Geom g = make_geom(n, false);
    for (int k = 0; k < M; ++k) {
        if (keys[k] < 0 || targets[k].vecch) continue;
        bool synced = false;
        const int own = factor_owner(ws, keys[k], n, nullptr, nullptr, &synced);
        if (own == kOwnerNone || !synced || !is_mine(ws, own)) continue;
        CachedFactor& cf = ws->cache[keys[k]];
        if (!cf.valid || !cf.T || !cf.w_synced) continue;
        rotate_w_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(cf.T + (size_t)g.npad * g.ld, z + (int64_t)k * n, cos(theta),
                                                               sin(theta) * sqrt(targets[k].scale), n);
        DGPB_LAUNCHED();
    }
    return DGPB_OK;
}

// out[b] = sum_i w_b[i]^2 over the y rows of cached factors (fixed-order tree sum); NULL row -> 0 (held elsewhere)
__global__ void __launch_bounds__(256) wnorm_kernel(const double* const* __restrict__ Ts, int64_t row_off, int n,
                                                    double* __restrict__ out) {
    __shared__ double sred[8];
    const double* T = Ts[blockIdx.x];
    double s = 0.0;
    if (T)
        for (int i = threadIdx.x; i < n; i += 256) s = fma(T[row_off + i], T[row_off + i], s);
    s = block_sum<256>(s, sred);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

static int nodes_loglik(Workspace* ws, const dgpb_node* nodes, int U, int64_t n, const double* src_override,
                        double* sum_host, cudaStream_t st, DenseBatchInfo* info_out = nullptr) {
    if (info_out) info_out->B = 0;
    int dense_batches = 0;
    double lls[256];
    DGPB_REQUIRE(U >= 1 && U <= 256, "too many upper nodes");
    int nv = 0;
    // Vecchia nodes first (results land in the Vecchia section of the result buffer)
    void* outv = nullptr;
    DGPB_TRY(ws->reserve(SLOT_OUT, sizeof(double) * kOutDoubles, &outv));
    double* out = (double*)outv;
    for (int u = 0; u < U; ++u) {
        if (!nodes[u].vecch) continue;
        DGPB_REQUIRE(nv < (kOutGrad - kOutVecchia) / 2, "too many Vecchia nodes in one layer");
        const dgpb_node* nd = &nodes[u];
        DGPB_REQUIRE(nd->ord && nd->NNarray, "Vecchia node without ord/NNarray");
        KernelDev kd;
        DGPB_TRY(make_kernel_dev(nd, n, src_override, &kd));
        void *Xo, *yo;
        // per-node gather buffers: the stream orders reuse across nodes
        DGPB_TRY(ws->reserve(SLOT_VX, sizeof(double) * (size_t)n * kd.D, &Xo));
        DGPB_TRY(ws->reserve(SLOT_VY, sizeof(double) * (size_t)n, &yo));
        gather_ord_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(kd, nd->output, nd->ord, n, (double*)Xo, (double*)yo);
        DGPB_LAUNCHED();
        VKern vk;
        double len[kMaxDim];
        int D;
        node_lengths(nd, len, &D);
        DGPB_TRY(make_vkern(nd->kind, D, len, D, &vk));
        DGPB_TRY(vecchia_llik_device(ws, vk, (double*)Xo, (double*)yo, nd->NNarray, n, nd->m + 1, nd->nugget, nullptr,
                                     out + kOutVecchia + 2 * nv, st));
        ++nv;
    }
    // dense nodes in batches
    int u0 = 0;
    while (u0 < U) {
        KernelDev kds[MAXB];
        const double* ys[MAXB];
        int map[MAXB];
        ScaleArgs sa;
        int B = 0;
        while (u0 < U && B < MAXB) {
            if (!nodes[u0].vecch) {
                DGPB_TRY(make_kernel_dev(&nodes[u0], n, src_override, &kds[B]));
                ys[B] = nodes[u0].output;
                sa.scale[B] = nodes[u0].scale;
                sa.est[B] = 0;
                map[B] = u0;
                ++B;
            }
            ++u0;
        }
        if (B == 0) break;
        Batch bt;
        Geom g;
        double* outd;
        DGPB_TRY(loglik_batch_device(ws, kds, ys, sa, B, n, &bt, &g, &outd, st));
        ++dense_batches;
        if (info_out) {
            info_out->bt = bt;
            info_out->g = g;
            info_out->B = dense_batches == 1 ? B : 0;
            for (int b = 0; b < B; ++b) info_out->map[b] = map[b];
        }
        int* info_host = reinterpret_cast<int*>(ws->pinned + kPinnedInfo);
        DGPB_CUDA_TRY(cudaMemcpyAsync(ws->pinned, outd, sizeof(double) * 4 * B, cudaMemcpyDeviceToHost, st));
        DGPB_CUDA_TRY(cudaMemcpyAsync(info_host, bt.info, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
        DGPB_CUDA_TRY(cudaStreamSynchronize(st));
        for (int b = 0; b < B; ++b) {
            if (info_host[b] != 0) {
                set_error("covariance of upper node %d is not positive definite (pivot %d)", map[b], info_host[b]);
                return DGPB_NOT_PD;
            }
            const double sc = nodes[map[b]].scale;
            lls[map[b]] = -0.5 * (ws->pinned[4 * b] + (double)n * log(sc) + ws->pinned[4 * b + 1] / sc);
        }
    }
    if (nv > 0) {
        DGPB_CUDA_TRY(cudaMemcpyAsync(ws->pinned + kPinnedVecchia, out + kOutVecchia, sizeof(double) * 2 * nv,
                                      cudaMemcpyDeviceToHost, st));
        DGPB_CUDA_TRY(cudaStreamSynchronize(st));
        int v = 0;
        for (int u = 0; u < U; ++u) {
            if (!nodes[u].vecch) continue;
            double quad = ws->pinned[kPinnedVecchia + 2 * v], logdet = ws->pinned[kPinnedVecchia + 2 * v + 1];
            if (!(quad == quad) || !(logdet == logdet)) {
                set_error("Vecchia block of upper node %d is not positive definite", u);
                return DGPB_NOT_PD;
            }
            lls[u] = -0.5 * (logdet + quad / nodes[u].scale);  // vecchia.py:179
            ++v;
        }
    }
    double s = 0.0;
    for (int u = 0; u < U; ++u) s += lls[u];  // same left-to-right order as imputation.py:70-78
    *sum_host = s;
    return DGPB_OK;
}

// prior draws nu[k] = chol(scale K) z_k for the target nodes.
// Several GPUs: a dense target's draw is formed on the rank that holds its factor (factors that have to be
// computed are dealt round-robin and stay where they were computed) and the row is broadcast, so every rank ends
// up with the same nu bit for bit.
static int prior_draws(Workspace* ws, const dgpb_node* targets, int M, int64_t n, const double* z, double* nu,
                       const int32_t* keys, cudaStream_t st, int ctx = 0) {
    const int G = ws->comm.world;
    constexpr int kMaxTargets = 256;
    DGPB_REQUIRE(M >= 1 && M <= kMaxTargets, "too many target nodes");
    for (int k = 0; k < M; ++k) {
        const dgpb_node* nd = &targets[k];
        if (!nd->vecch) continue;
        DGPB_REQUIRE(nd->ord && nd->NNarray, "Vecchia node without ord/NNarray");
        KernelDev kd;
        DGPB_TRY(make_kernel_dev(nd, n, nullptr, &kd));
        void *Xo, *tmp;
        DGPB_TRY(ws->reserve(SLOT_VX, sizeof(double) * (size_t)n * kd.D, &Xo));
        DGPB_TRY(ws->reserve(SLOT_VY, sizeof(double) * (size_t)n, &tmp));
        gather_ord_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(kd, nullptr, nd->ord, n, (double*)Xo, nullptr);
        DGPB_LAUNCHED();
        VKern vk;
        double len[kMaxDim];
        int D;
        node_lengths(nd, len, &D);
        DGPB_TRY(make_vkern(nd->kind, D, len, D, &vk));
        DGPB_TRY(vecchia_mvn_draw_device(ws, vk, (double*)Xo, nd->NNarray, n, nd->m + 1, nd->scale, nd->nugget,
                                         z + (int64_t)k * n, (double*)tmp, st));
        scatter_ord_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>((double*)tmp, nd->ord, n, nu + (int64_t)k * n);
        DGPB_LAUNCHED();
    }
    // ---- dense targets: who forms which row (decided identically on every rank)
    int own[kMaxTargets], root[kMaxTargets];
    bool fresh[kMaxTargets];
    int nfresh = 0;
    const Geom g = make_geom(n, false);
    for (int k = 0; k < M; ++k) {
        own[k] = kOwnerNone;
        root[k] = -1;
        fresh[k] = false;
        if (targets[k].vecch) continue;
        own[k] = factor_owner(ws, keys ? keys[k] : -1, n);
        if (own[k] == kOwnerNone) {
            fresh[k] = true;
            own[k] = G > 1 ? nfresh % G : kOwnerAll;
            ++nfresh;
        }
        root[k] = own[k] == kOwnerAll ? -1 : own[k];
    }
    // ---- rows from factors kept from an earlier block update of this I-step (same inputs, same theta)
    for (int k = 0; k < M; ++k) {
        if (targets[k].vecch || fresh[k] || !is_mine(ws, own[k])) continue;
        const CachedFactor& cf = ws->cache[keys[k]];
        DGPB_REQUIRE(cf.valid && cf.T && cf.n == n, "cached factor missing on its owner rank");
        DGPB_TRY(launch_trmv(cf.T, g.ld, g.n, sqrt(targets[k].scale), z + (int64_t)k * n, nu + (int64_t)k * n, st));
    }
    // ---- factors that have to be computed, in batches
    int local_bad = 0, bad_node = -1, bad_pivot = 0;
    int k0 = 0;
    while (k0 < M) {
        KernelDev kds[MAXB];
        int map[MAXB];
        int B = 0;
        while (k0 < M && B < MAXB) {
            if (!targets[k0].vecch && fresh[k0] && is_mine(ws, own[k0])) {
                DGPB_TRY(make_kernel_dev(&targets[k0], n, nullptr, &kds[B]));
                map[B] = k0;
                ++B;
            }
            ++k0;
        }
        if (B == 0) continue;
        Batch bt;
        double* outd;
        // the context that is free: a wave factored ahead of the last acceptance may still run in the other one
        if (ws->wave_done[ctx]) DGPB_CUDA_TRY(cudaStreamWaitEvent(st, ws->wave_done[ctx], 0));
        DGPB_TRY(setup_batch_slot(ws, ctx, g, B, &bt, &outd));
        DGPB_TRY(assemble(g, kds, nullptr, bt, B, st));
        DGPB_TRY(factorize(g, bt, B, st, ctx));
        for (int b = 0; b < B; ++b)
            if (keys && keys[map[b]] >= 0) DGPB_TRY(cache_store(ws, keys[map[b]], g, bt, b, st));
        DGPB_TRY(restore_diag_blocks(g, bt, B, st));
        for (int b = 0; b < B; ++b) {
            DGPB_TRY(launch_trmv(bt.T[b], g.ld, g.n, sqrt(targets[map[b]].scale), z + (int64_t)map[b] * n,
                                 nu + (int64_t)map[b] * n, st));
        }
        int* info_host = reinterpret_cast<int*>(ws->pinned + kPinnedInfo);
        DGPB_CUDA_TRY(cudaMemcpyAsync(info_host, bt.info, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
        DGPB_CUDA_TRY(cudaStreamSynchronize(st));
        for (int b = 0; b < B; ++b)
            if (info_host[b] != 0) {
                if (keys && keys[map[b]] >= 0) ws->cache[keys[map[b]]].valid = false;
                if (!local_bad) {
                    bad_node = map[b];
                    bad_pivot = info_host[b];
                }
                local_bad = 1;
            }
        if (local_bad && G <= 1) break;
    }
    if (nfresh > 0) {
        int bad = local_bad;
        if (G > 1) DGPB_TRY(comm_max_flag(ws, local_bad, &bad, st));   // every rank leaves the update the same way
        if (bad) {
            if (local_bad)
                set_error("prior covariance of target node %d is not positive definite (pivot %d)", bad_node, bad_pivot);
            else
                set_error("prior covariance of a target node is not positive definite (found on another rank)");
            return DGPB_NOT_PD;
        }
        for (int k = 0; k < M; ++k)
            if (fresh[k] && keys && keys[k] >= 0) set_owner(ws, keys[k], own[k], n, nullptr);
    }
    if (G > 1) DGPB_TRY(comm_bcast_rows(ws, nu, n, root, M, st));
    return DGPB_OK;
}

}  // namespace dgpb

using namespace dgpb;

namespace dgpb {

int g_ess_target_b = 8;  // matrices per speculative wave (dgpb_tune "ess_batch"); 0/1 = one proposal at a time
int g_ess_cached_threshold = 1;  // threshold from cached factors by a triangular solve (dgpb_tune "ess_trsv")
int g_ess_rotate_w = 1;          // ... or, when the cached L^-1 y was kept in step with the outputs, by a dot product
int g_ess_prefetch = 1;          // assemble the next wave while the current one is factored (dgpb_tune "ess_prefetch")

// |L^-1 y|^2 for a cached factor L (T layout, diagonal blocks restored): forward substitution by ONE CTA per matrix,
// x kept in shared memory.  Per 64-row block the eight warps form y_i - sum_j L_ij x_j for eight rows each
// (row-contiguous double2 loads, 16 in flight per thread), then warp 0 solves the 64 x 64 diagonal block with the
// solution travelling by shuffles.  n^2 flop per matrix: the threshold of an ESS block update whose upper nodes
// kept their inputs (only their outputs moved) costs this instead of a factorisation (n^3/3).
__global__ void __launch_bounds__(256, 1) trsv_quad_kernel(const double* const* __restrict__ Ts, int64_t ld, int n,
                                                           const double* const* __restrict__ ys, double* __restrict__ out) {
    extern __shared__ double xs[];   // n (padded to 64) + 64 right-hand sides + 64 x 65 diagonal block
    const double* __restrict__ T = Ts[blockIdx.x];
    const double* __restrict__ y = ys[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (T == nullptr) {   // factor held by another rank (multi-GPU): that rank fills this entry
        if (tid == 0) out[blockIdx.x] = 0.0;
        return;
    }
    const int nb = (n + 63) / 64;
    double* rs = xs + (size_t)nb * 64;   // right-hand side of the current block after the GEMV part
    double* sd = rs + 64;                // 64 x 65 diagonal block
    double quad = 0.0;
    for (int b = 0; b < nb; ++b) {
        const int r0 = 64 * b;
        // rows r0 + 8 w + {0..7}: dot products with x[0, r0)
        double acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = 0.0;
        for (int c = 2 * lane; c < r0; c += 64) {
            const double2 xv = *reinterpret_cast<const double2*>(&xs[c]);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int row = r0 + 8 * w + r;
                if (row < n) {
                    const double2 lv = *reinterpret_cast<const double2*>(&T[(int64_t)row * ld + c]);
                    acc[r] = fma(lv.x, xv.x, fma(lv.y, xv.y, acc[r]));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double v = acc[r];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const int row = r0 + 8 * w + r;
            if (lane == 0) rs[8 * w + r] = row < n ? y[row] - v : 0.0;
        }
        // the 64 x 64 diagonal block (identity outside the matrix) staged for the substitution
        for (int idx = tid; idx < 64 * 64; idx += 256) {
            const int i = idx >> 6, j = idx & 63;
            const int gi = r0 + i, gj = r0 + j;
            sd[i * 65 + j] = (gi < n && gj <= gi) ? T[(int64_t)gi * ld + gj] : (i == j ? 1.0 : 0.0);
        }
        __syncthreads();
        if (w == 0) {
            // forward substitution on the diagonal block: lane holds rows lane and lane + 32
            double b0 = rs[lane], b1 = rs[lane + 32];
            const int ra = r0 + lane, rb = r0 + lane + 32;
            for (int j = 0; j < 64; ++j) {
                const double bj = __shfl_sync(0xffffffffu, j < 32 ? b0 : b1, j & 31);
                const double xj = bj / sd[j * 65 + j];
                if (lane == (j & 31)) {
                    if (j < 32) b0 = xj; else b1 = xj;
                }
                if (lane > j) b0 = fma(-sd[lane * 65 + j], xj, b0);
                if (lane + 32 > j) b1 = fma(-sd[(lane + 32) * 65 + j], xj, b1);
            }
            xs[r0 + lane] = ra < n ? b0 : 0.0;
            xs[r0 + lane + 32] = rb < n ? b1 : 0.0;
            quad += (ra < n ? b0 * b0 : 0.0) + (rb < n ? b1 * b1 : 0.0);
        }
        __syncthreads();
    }
    if (w == 0) {
        for (int o = 16; o > 0; o >>= 1) quad += __shfl_xor_sync(0xffffffffu, quad, o);
        if (lane == 0) out[blockIdx.x] = quad;
    }
}

// Sum of the upper nodes' log-likelihoods at the CURRENT state from cached factors (see trsv_quad_kernel).
// Returns DGPB_OK with *used = 1 when every node had a cached factor with its log-determinant, else *used = 0.
// Several GPUs: every rank solves with the factors it holds and the quadratic forms are all-gathered.
static int cached_threshold(Workspace* ws, const dgpb_node* nodes, int U, int64_t n, const int32_t* keys, double* sum_host,
                            int* used, cudaStream_t st) {
    *used = 0;
    if (!keys || U > MAXB) return DGPB_OK;
    const int G = ws->comm.world, me = ws->comm.rank;
    const Geom g = make_geom(n, false);
    const size_t smem = ((size_t)((n + 63) / 64) * 64 + 64 + 64 * 65) * sizeof(double);
    if (smem > 200 * 1024) return DGPB_OK;
    const double* hT[MAXB];
    const double* hy[MAXB];
    double logdets[MAXB];
    int own[MAXB];
    int mine = 0;
    bool all_synced = true;   // every node's cached L^-1 y follows its current output: dot products instead of solves
    for (int u = 0; u < U; ++u) {
        if (nodes[u].vecch || keys[u] < 0) return DGPB_OK;
        bool has_logdet = false, synced = false;
        own[u] = factor_owner(ws, keys[u], n, &has_logdet, &logdets[u], &synced);
        if (own[u] == kOwnerNone || !has_logdet) return DGPB_OK;
        all_synced = all_synced && synced;
        hT[u] = nullptr;
        hy[u] = nodes[u].output;
        if (is_mine(ws, own[u])) {
            const CachedFactor& cf = ws->cache[keys[u]];
            DGPB_REQUIRE(cf.valid && cf.T && cf.n == n, "cached factor missing on its owner rank");
            hT[u] = cf.T;
            ++mine;
        }
    }
    void *ptab, *pout;
    DGPB_TRY(ws->reserve(SLOT_MISC, sizeof(double*) * 2 * MAXB, &ptab));
    DGPB_TRY(ws->reserve(SLOT_OUT, sizeof(double) * kOutDoubles, &pout));
    double* out = (double*)pout + kOutVecchia;   // a region the dense batch results do not use (kCommBlock doubles)
    if (mine > 0) {
        const double** dT = (const double**)ptab;
        const double** dy = dT + MAXB;
        DGPB_CUDA_TRY(cudaMemcpyAsync(dT, hT, sizeof(double*) * U, cudaMemcpyHostToDevice, st));
        DGPB_CUDA_TRY(cudaMemcpyAsync(dy, hy, sizeof(double*) * U, cudaMemcpyHostToDevice, st));
        static bool cfg = false;
        if (!cfg) {
            DGPB_CUDA_TRY(cudaFuncSetAttribute(trsv_quad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            cfg = true;
        }
        if (all_synced && g_ess_rotate_w)
            wnorm_kernel<<<U, 256, 0, st>>>(dT, (int64_t)g.npad * g.ld, g.n, out);
        else
            trsv_quad_kernel<<<U, 256, smem, st>>>(dT, g.ld, g.n, dy, out);   // nodes held elsewhere: NULL factor, skipped
        DGPB_LAUNCHED();
    }
    const double* res = ws->pinned + kPinnedWave;
    if (G > 1) {
        void* pc;
        DGPB_TRY(ws->reserve(SLOT_COMM, sizeof(double) * kCommBlock * kMaxRanks, &pc));
        DGPB_TRY(comm_allgather(ws, out, (double*)pc, kCommBlock, st));
        DGPB_CUDA_TRY(cudaMemcpyAsync(ws->pinned + kPinnedWave, pc, sizeof(double) * kCommBlock * G, cudaMemcpyDeviceToHost, st));
    } else {
        DGPB_CUDA_TRY(cudaMemcpyAsync(ws->pinned + kPinnedWave, out, sizeof(double) * U, cudaMemcpyDeviceToHost, st));
    }
    DGPB_CUDA_TRY(cudaStreamSynchronize(st));
    double s = 0.0;
    for (int u = 0; u < U; ++u) {   // same left-to-right order as imputation.py:70-78
        const double sc = nodes[u].scale;
        const int r = own[u] == kOwnerAll ? me : own[u];
        s += -0.5 * (logdets[u] + (double)n * log(sc) + res[(G > 1 ? r * kCommBlock : 0) + u] / sc);
    }
    *sum_host = s;
    *used = 1;
    return DGPB_OK;
}

// ---- ESS waves ---------------------------------------------------------------------------------------------------
// A wave is a list of ITEMS: optionally the threshold (the upper nodes at the current state, item 0) followed by
// candidate angles.  Item i is evaluated by rank i % W in its local slot i / W (W = ranks sharing the wave): the U
// upper-node matrices of local slot l are matrices l U .. l U + U - 1 of that rank's batched factorisation.
// Step 1: assemble the matrices of this rank's items (srcs[l] = latent layer image read by the upper nodes, NULL =
// the nodes' own `src`) into T set `tslot` on stream `st`.
static int dense_items_assemble(Workspace* ws, int tslot, const dgpb_node* nodes, int U, int64_t n,
                                const double* const* srcs, int nlocal, Batch* bt_out, Geom* g_out, cudaStream_t st) {
    const int B = nlocal * U;
    *g_out = make_geom(n, false);
    if (B == 0) return DGPB_OK;
    DGPB_REQUIRE(B <= MAXB, "wave does not fit one batch");
    KernelDev kds[MAXB];
    const double* ys[MAXB];
    for (int i = 0; i < nlocal; ++i)
        for (int u = 0; u < U; ++u) {
            const int b = i * U + u;
            DGPB_TRY(make_kernel_dev(&nodes[u], n, srcs[i], &kds[b]));
            ys[b] = nodes[u].output;
        }
    double* outd;
    DGPB_TRY(setup_batch_slot(ws, tslot, *g_out, B, bt_out, &outd));
    DGPB_TRY(assemble_matrices(*g_out, kds, ys, *bt_out, B, st));
    return DGPB_OK;
}

// Step 2: factorise this rank's matrices, reduce, exchange the result blocks (4 doubles per matrix: log|K|, y'K^-1y,
// -, info) and launch the device-to-host copy of all of them (no host sync).
static int dense_items_factor(Workspace* ws, int ctx, const dgpb_node* nodes, int U, int nlocal, int W, const Batch& bt,
                              const Geom& g, cudaStream_t st) {
    const int B = nlocal * U;
    void* pO;
    DGPB_TRY(ws->reserve(ctx ? SLOT_OUT2 : SLOT_OUT, sizeof(double) * kOutDoubles, &pO));
    double* outd = (double*)pO;
    if (B > 0) {
        ScaleArgs sa;
        for (int b = 0; b < B; ++b) {
            sa.scale[b] = nodes[b % U].scale;
            sa.est[b] = 0;
        }
        DGPB_TRY(factor_reduce(g, bt, B, sa, outd, st, ctx));
    }
    double* stage = ws->pinned + kPinnedWave + (size_t)ctx * kPinnedWaveStride;
    if (W > 1) {
        void* pc;
        DGPB_TRY(ws->reserve(SLOT_COMM, sizeof(double) * kCommBlock * kMaxRanks, &pc));
        DGPB_TRY(comm_allgather(ws, outd, (double*)pc, kCommBlock, st));
        DGPB_CUDA_TRY(cudaMemcpyAsync(stage, pc, sizeof(double) * kCommBlock * W, cudaMemcpyDeviceToHost, st));
    } else if (B > 0) {
        DGPB_CUDA_TRY(cudaMemcpyAsync(stage, outd, sizeof(double) * 4 * B, cudaMemcpyDeviceToHost, st));
    }
    return DGPB_OK;
}

// Step 3: wait for the results and form the per-item sums.  pd[i] = 0 if one of the item's matrices is not positive
// definite (bad_node[i] = which).  The caller decides what an indefinite item means: the reference only ever
// evaluates items up to the first accepted one.
static int dense_items_fetch(Workspace* ws, int ctx, const dgpb_node* nodes, int U, int64_t n, int nitems, int W,
                             double* sums, int* pd, int* bad_node, double* logdets, cudaStream_t st) {
    DGPB_CUDA_TRY(cudaStreamSynchronize(st));
    const double* res = ws->pinned + kPinnedWave + (size_t)ctx * kPinnedWaveStride;
    for (int i = 0; i < nitems; ++i) {
        const double* blk = res + (size_t)(i % W) * kCommBlock + (size_t)(i / W) * U * 4;
        double s = 0.0;
        pd[i] = 1;
        bad_node[i] = -1;
        for (int u = 0; u < U; ++u) {  // same left-to-right order as imputation.py:70-78
            if (blk[4 * u + 3] != 0.0 && pd[i]) {
                pd[i] = 0;
                bad_node[i] = u;
            }
            const double sc = nodes[u].scale;
            s += -0.5 * (blk[4 * u] + (double)n * log(sc) + blk[4 * u + 1] / sc);
            if (logdets) logdets[i * U + u] = blk[4 * u];
        }
        sums[i] = s;
    }
    return DGPB_OK;
}

}  // namespace dgpb

// make the two wave contexts of the workspace (streams + events) on first use
static int wave_contexts_init(Workspace* ws) {
    if (ws->wave_stream[0]) return DGPB_OK;
    for (int c = 0; c < 2; ++c) {
        DGPB_CUDA_TRY(cudaStreamCreateWithFlags(&ws->wave_stream[c], cudaStreamNonBlocking));
        DGPB_CUDA_TRY(cudaEventCreateWithFlags(&ws->wave_assembled[c], cudaEventDisableTiming));
        DGPB_CUDA_TRY(cudaEventCreateWithFlags(&ws->wave_done[c], cudaEventDisableTiming));
    }
    return DGPB_OK;
}

// every stream-ordered reader of the proposal images / prior draws of an earlier block update has finished
static int join_wave_staging(Workspace* ws, cudaStream_t st) {
    for (int c = 0; c < 2; ++c)
        if (ws->wave_assembled[c]) DGPB_CUDA_TRY(cudaStreamWaitEvent(st, ws->wave_assembled[c], 0));
    return DGPB_OK;
}

namespace dgpb {
int g_ess_wave_total = 8;   // candidates per wave aimed at when several GPUs share it (dgpb_tune "ess_wave_total")
// Factor the next single-candidate wave while the current one drains its tail (dgpb_tune "ess_overlap").  Measured on
// BASELINE config 3 (n = 5000, B200): two waves in flight raise the FLOP rate issued from 59.7 % to 63.4 % of the DGEMM
// peak, and the wave that is in flight at every acceptance (one in ~17 for the 8-node layer pair) costs the same
// 6 % again: 0.323 against 0.321 iterations/s.  Off by default: same speed, less energy.
int g_ess_overlap = 0;
}

// The angles ESS will try are known in advance: a rejection is the only branch of the bracket rule
// (imputation.py:111-119), so theta_{k+1} depends on theta_k and the next uniform, never on a likelihood value.
// Proposals are therefore evaluated in WAVES of consecutive candidate angles and the first accepted one wins;
// candidates after it are discarded.  On one GPU a wave is one batched factorisation (g_ess_target_b / n_uppers
// candidates); on W GPUs every rank factors its share of the wave (W times as many candidates in the time of
// one batch), the per-matrix results are all-gathered (1 KB per rank) and every rank replays the same decisions.
// Accept/shrink decisions, the angles reported and the number of uniforms consumed are exactly those of the
// one-at-a-time loop.  When the threshold likelihood is not cached it rides along in the first wave.
//
// Wave pipeline: two factorisation contexts (T set, side buffers, look-ahead streams).  While wave k is factored in
// one context, wave k + 1 -- whose angles are known under the assumption that all of wave k is rejected -- is
// proposed and assembled in the other; when a wave holds a single candidate (an 8-node upper layer: acceptance
// ends about one wave in sixteen) it is also FACTORED there, so its bulk updates fill the SMs that the serial
// panel chain of wave k's last hyper-blocks leaves idle.  An acceptance leaves that wave running; nothing reads its
// results, and the next block update starts in the context that is free.
extern "C" int dgpb_ess_block_cached(dgpb_ws* ws, const dgpb_node* targets, int n_targets,
                                     const int32_t* target_rows_host, double* layer_out, int64_t layer_width,
                                     const dgpb_node* uppers, int n_uppers, int64_t n, const double* z,
                                     const double* u_host, int nu, int* n_prop_host, double* theta_host,
                                     const int32_t* target_keys_host, const int32_t* upper_keys_host,
                                     double* threshold_io_host, void* stream) {
    DGPB_NVTX("dgpb:ess_block");
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(ws && targets && uppers && layer_out && z && u_host && target_rows_host, "NULL argument");
    DGPB_REQUIRE(n_targets >= 1 && n_uppers >= 1 && n >= 1 && nu >= 3, "bad sizes");
    for (int k = 0; k < n_targets; ++k)
        DGPB_REQUIRE(target_rows_host[k] >= 0 && target_rows_host[k] < layer_width, "target row out of range");
    bool all_dense = true;
    for (int u = 0; u < n_uppers; ++u) all_dense = all_dense && !uppers[u].vecch;
    constexpr int kMaxWave = 8;                     // candidate angles per rank and wave
    constexpr int kMaxCand = kMaxWave * kMaxRanks;  // candidate angles per wave
    int cap = 1;  // candidate angles per rank and wave
    if (all_dense && n_uppers <= MAXB) cap = std::max(1, std::min(kMaxWave, std::min(g_ess_target_b, (int)MAXB) / n_uppers));
    const bool batched = all_dense && n_uppers <= MAXB;
    // ranks that share a wave; Vecchia or oversized upper layers are evaluated by every rank (same numbers everywhere)
    const int W = batched ? ws->comm.world : 1;
    const int me = batched ? ws->comm.rank : 0;
    // several GPUs: W ranks already multiply the candidates of a wave; beyond ~g_ess_wave_total candidates per wave the
    // extra ones are almost always behind the accepted one, so each rank takes fewer and its batch gets smaller (a
    // 2-node upper layer on 8 GPUs: one candidate = 2 matrices per rank, 6 ms instead of 15 ms for 4 candidates)
    if (W > 1) cap = std::max(1, std::min(cap, (g_ess_wave_total + W - 1) / W));
    const int nslots = cap + 1;   // local proposal images per wave (rank 0 may carry the threshold item as slot 0)
    // factor the next wave ahead of the decision: single-candidate waves on one GPU (NCCL calls of two waves in
    // flight would have to be ordered across ranks, and multi-candidate waves accept too often for the bet to pay)
    const bool overlap = batched && g_ess_overlap && g_ess_prefetch && cap == 1 && W == 1;
    // g_ess_overlap = 2: only while the wave factored ahead is very unlikely to be wasted -- ESS needs about as many
    // proposals every time a given layer pair is updated (the bracket halves per rejection), so a wave that would
    // END two standard deviations before the running mean of this pair's proposal count is almost never behind an
    // acceptance
    int overlap_until = 1 << 30;   // candidates [0, overlap_until) may be factored ahead
    Workspace::PropStat* pstat = nullptr;
    if (upper_keys_host && upper_keys_host[0] >= 0) {
        pstat = &ws->prop_stats[upper_keys_host[0]];
        if (g_ess_overlap == 2) {
            overlap_until = 0;
            if (pstat->cnt >= 4) {
                const double sd = sqrt(pstat->m2 / (pstat->cnt - 1));
                overlap_until = (int)floor(pstat->mean - 2.0 * sd - 1.0);
            }
        }
    }

    DGPB_TRY(wave_contexts_init(ws));
    DGPB_TRY(join_wave_staging(ws, st));   // readers of SLOT_PROP / SLOT_NU left over from the previous block update
    if (!batched) DGPB_TRY(join_waves(ws, st));   // the unbatched path uses the T sets outside the pipeline
    int cur = ws->wave_cur;   // the context whose buffers are free now (the other may hold a wave still in flight)
    void *pnu, *pprop;
    const size_t layer_elems = (size_t)layer_width * n;
    DGPB_TRY(ws->reserve(SLOT_NU, sizeof(double) * (size_t)n_targets * n, &pnu));
    DGPB_TRY(ws->reserve(SLOT_PROP, sizeof(double) * layer_elems * nslots * 2, &pprop));
    double* nuv = (double*)pnu;
    double* prop = (double*)pprop;

    DGPB_TRY(prior_draws(ws, targets, n_targets, n, z, nuv, target_keys_host, st, cur));

    double log_y = 0.0;
    const bool have_thr = threshold_io_host && *threshold_io_host == *threshold_io_host;
    bool thr_pending = false;  // threshold to be computed together with the first wave
    int thr_from_cache = 0;
    if (!have_thr && g_ess_cached_threshold)
        DGPB_TRY(cached_threshold(ws, uppers, n_uppers, n, upper_keys_host, &log_y, &thr_from_cache, st));
    if (have_thr) {
        log_y = *threshold_io_host;  // sum of the upper log-likelihoods at the current state is already known
    } else if (thr_from_cache) {
        // the upper nodes kept their inputs since their factors were stored: only their outputs moved
    } else if (batched && (cap + 1) * n_uppers <= MAXB) {
        thr_pending = true;
    } else {
        DGPB_TRY(join_waves(ws, st));
        DGPB_TRY(nodes_loglik(ws, uppers, n_uppers, n, nullptr, &log_y, st));
    }
    int ui = 0;
    const double log_u0 = log(u_host[ui++]);                 // imputation.py:79
    if (!thr_pending) log_y += log_u0;
    double theta = 0.0 + (2.0 * M_PI - 0.0) * u_host[ui++];  // uniform(0, 2pi)          imputation.py:81
    double tmin = theta - 2.0 * M_PI, tmax = theta;

    // rows of the layer that are not being updated are shared by every proposal
    for (int s = 0; s < 2 * nslots; ++s)
        DGPB_CUDA_TRY(cudaMemcpyAsync(prop + s * layer_elems, layer_out, sizeof(double) * layer_elems,
                                      cudaMemcpyDeviceToDevice, st));
    int nprop = 0;
    static thread_local cudaEvent_t pre_go = nullptr;
    if (batched) {
        if (!pre_go) DGPB_CUDA_TRY(cudaEventCreateWithFlags(&pre_go, cudaEventDisableTiming));
        DGPB_TRY(reserve_batches(ws, make_geom(n, false), std::min((int)MAXB, (cap + 1) * n_uppers)));
        // everything queued so far (prior draws, layer copies) is what the wave streams must wait for; nothing
        // the waves read is written again before an acceptance
        DGPB_CUDA_TRY(cudaEventRecord(pre_go, st));
    }
    auto plan_wave = [&](double th0, double lmin, double lmax, int uidx, double* out_thetas) {
        return ess_plan_wave(th0, lmin, lmax, u_host + uidx, nu - uidx, cap * W, out_thetas);
    };
    auto launch_proposal = [&](double th, double* image, cudaStream_t s2) -> int {
        const double c = cos(th), sn = sin(th);
        for (int k = 0; k < n_targets; ++k) {
            const int64_t row = target_rows_host[k];
            propose_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s2>>>(image + row * n, layer_out + row * n,
                                                                  nuv + (int64_t)k * n, c, sn, n);
            DGPB_LAUNCHED();
        }
        return DGPB_OK;
    };
    // one wave in flight in one of the two contexts
    struct Wave {
        bool staged = false, factored = false;
        int S = 0, first = 0, nlocal = 0;
        double thetas[kMaxCand];
        Batch bt;
        Geom g;
    } wv[2];
    // propose + assemble this rank's items of a wave (items [0, first) = threshold, then the candidates) in context c
    auto stage_wave = [&](int c, const double* th, int S, int first) -> int {
        Wave& w = wv[c];
        cudaStream_t cs = ws->wave_stream[c];
        double* pbuf = prop + (size_t)c * nslots * layer_elems;
        DGPB_CUDA_TRY(cudaStreamWaitEvent(cs, pre_go, 0));   // the layer image and the prior draws exist
        const double* srcs[kMaxWave + 1];
        int L = 0;
        for (int i = me; i < first + S; i += W) {   // local slot of item i = i / W = L
            DGPB_REQUIRE(L < nslots, "wave item does not fit the proposal buffer");
            if (i < first) {
                srcs[L] = nullptr;
            } else {
                DGPB_TRY(launch_proposal(th[i - first], pbuf + (size_t)L * layer_elems, cs));
                srcs[L] = pbuf + (size_t)L * layer_elems;
            }
            ++L;
        }
        w.S = S;
        w.first = first;
        w.nlocal = L;
        for (int s = 0; s < S; ++s) w.thetas[s] = th[s];
        DGPB_TRY(dense_items_assemble(ws, c, uppers, n_uppers, n, srcs, L, &w.bt, &w.g, cs));
        DGPB_CUDA_TRY(cudaEventRecord(ws->wave_assembled[c], cs));
        w.staged = true;
        w.factored = false;
        return DGPB_OK;
    };
    auto factor_wave = [&](int c) -> int {
        Wave& w = wv[c];
        DGPB_TRY(dense_items_factor(ws, c, uppers, n_uppers, w.nlocal, W, w.bt, w.g, ws->wave_stream[c]));
        DGPB_CUDA_TRY(cudaEventRecord(ws->wave_done[c], ws->wave_stream[c]));
        w.factored = true;
        return DGPB_OK;
    };
    while (true) {
        // ---- candidate angles of this wave (each one assumes every earlier one was rejected)
        double thetas[kMaxCand];
        const int S = plan_wave(theta, tmin, tmax, ui, thetas);
        double* pcur = prop + (size_t)cur * nslots * layer_elems;
        // ---- likelihoods of the wave
        double sums[kMaxCand + 1];
        double wave_logdets[(kMaxCand + 1) * MAXB];
        int pd[kMaxCand + 1], bad[kMaxCand + 1];
        DenseBatchInfo info;
        int first = 0;  // index of candidate 0 in sums[]
        if (batched) {
            Wave& w = wv[cur];
            bool ready = w.staged && w.S == S && !thr_pending;   // staged ahead of time with exactly these angles?
            for (int s = 0; ready && s < S; ++s) ready = w.thetas[s] == thetas[s];
            if (!ready) DGPB_TRY(stage_wave(cur, thetas, S, thr_pending ? 1 : 0));
            first = w.first;
            if (!w.factored) DGPB_TRY(factor_wave(cur));
            // ---- speculate: the next wave in the other context while this one is being factored
            {
                double lmin = tmin, lmax = tmax;
                for (int s = 0; s < S; ++s)
                    if (thetas[s] < 0.0) lmin = thetas[s]; else lmax = thetas[s];
                const int ui_next = ui + S;   // S rejections consume S uniforms (the last one draws the next first angle)
                wv[cur ^ 1].staged = false;
                if (g_ess_prefetch && ui_next <= nu && ui_next >= 1) {
                    double next_thetas[kMaxCand];
                    const double th_next = lmin + (lmax - lmin) * u_host[ui_next - 1];
                    const int Sn = plan_wave(th_next, lmin, lmax, ui_next, next_thetas);
                    DGPB_TRY(stage_wave(cur ^ 1, next_thetas, Sn, 0));
                    if (overlap && nprop + S + Sn <= overlap_until) DGPB_TRY(factor_wave(cur ^ 1));
                }
            }
            DGPB_TRY(dense_items_fetch(ws, cur, uppers, n_uppers, n, first + S, W, sums, pd, bad, wave_logdets,
                                       ws->wave_stream[cur]));
            if (first == 1) {
                if (!pd[0]) {
                    set_error("covariance of upper node %d is not positive definite", bad[0]);
                    return DGPB_NOT_PD;
                }
                log_y = sums[0] + log_u0;
                thr_pending = false;
            }
        } else {
            DGPB_TRY(launch_proposal(thetas[0], pcur, st));
            DGPB_TRY(nodes_loglik(ws, uppers, n_uppers, n, pcur, &sums[0], st, &info));
            pd[0] = 1;
        }
        // ---- replay the one-at-a-time decisions over the wave
        int accepted = -1;
        for (int s = 0; s < S; ++s) {
            if (theta_host) theta_host[nprop] = thetas[s];
            ++nprop;
            if (!pd[first + s]) {
                if (n_prop_host) *n_prop_host = nprop;
                set_error("covariance of upper node %d is not positive definite (proposal %d)", bad[first + s], nprop);
                return DGPB_NOT_PD;
            }
            if (sums[first + s] > log_y) {  // imputation.py:107-110
                accepted = s;
                break;
            }
            if (thetas[s] < 0.0) tmin = thetas[s]; else tmax = thetas[s];  // imputation.py:115-118
            if (s + 1 < S) ++ui;  // the uniform that produced thetas[s + 1]
        }
        if (accepted >= 0) {
            const int ia = first + accepted;             // item index of the accepted candidate
            const int owner = ia % W, slot = ia / W;     // the rank that evaluated it, its local slot there
            if (batched) {   // this rank's candidates behind the accepted one were issued for nothing
                int behind = 0;
                for (int i = me; i < first + S; i += W) behind += i > ia;
                if (wv[cur ^ 1].factored) behind += wv[cur ^ 1].nlocal;   // ... and so was the wave factored ahead
                profile_wasted(behind * n_uppers, n);
                DGPB_CUDA_TRY(cudaStreamWaitEvent(st, ws->wave_done[cur], 0));   // (already complete: the host waited)
            }
            const double* pa = pcur + (size_t)slot * layer_elems;
            if (owner != me) {   // evaluated elsewhere: the same elementwise proposal, recomputed here
                DGPB_TRY(launch_proposal(thetas[accepted], pcur, st));
                pa = pcur;
            }
            for (int k = 0; k < n_targets; ++k) {
                const int64_t row = target_rows_host[k];
                DGPB_CUDA_TRY(cudaMemcpyAsync(layer_out + row * n, pa + row * n, sizeof(double) * n,
                                              cudaMemcpyDeviceToDevice, st));
            }
            DGPB_TRY(rotate_cached_w(ws, targets, n_targets, target_keys_host, n, z, thetas[accepted], st));
            // the factors of the accepted proposal are the prior factors these nodes need as targets of the
            // next layer pair (they stay on the rank that computed them); the accepted log-likelihood is the
            // next threshold when their outputs are fixed
            if (upper_keys_host) {
                if (batched) {
                    for (int u = 0; u < n_uppers; ++u) {
                        if (upper_keys_host[u] < 0) continue;
                        const double* ld = &wave_logdets[ia * n_uppers + u];
                        if (owner == me)
                            DGPB_TRY(cache_store(ws, upper_keys_host[u], wv[cur].g, wv[cur].bt, slot * n_uppers + u, st, ld));
                        set_owner(ws, upper_keys_host[u], W > 1 ? owner : kOwnerAll, n, ld);
                    }
                } else {
                    for (int b = 0; b < info.B; ++b)
                        if (upper_keys_host[info.map[b]] >= 0) {
                            DGPB_TRY(cache_store(ws, upper_keys_host[info.map[b]], info.g, info.bt, b, st));
                            set_owner(ws, upper_keys_host[info.map[b]], kOwnerAll, n, nullptr);
                        }
                    if (info.B == 0)  // batch not reusable: drop anything stale
                        for (int u = 0; u < n_uppers; ++u)
                            if (upper_keys_host[u] >= 0) {
                                ws->cache[upper_keys_host[u]].valid = false;
                                ws->owner.erase(upper_keys_host[u]);
                            }
                }
            }
            if (threshold_io_host) *threshold_io_host = sums[first + accepted];
            // the next block update starts in this context: the other one may still hold the wave factored ahead
            ws->wave_cur = cur;
            break;
        }
        if (ui >= nu) {
            if (n_prop_host) *n_prop_host = nprop;
            set_error("ESS ran out of uniforms after %d proposals", nprop);
            return DGPB_BAD_ARG;
        }
        theta = tmin + (tmax - tmin) * u_host[ui++];  // imputation.py:119
        if (batched) cur ^= 1;   // the wave staged (and maybe factored) ahead becomes the current one
    }
    if (n_prop_host) *n_prop_host = nprop;
    if (pstat) {   // Welford update of this pair's proposal count
        pstat->cnt += 1;
        const double d = nprop - pstat->mean;
        pstat->mean += d / pstat->cnt;
        pstat->m2 += d * (nprop - pstat->mean);
    }
    return DGPB_OK;
}

// The output of the node stored under `key` was changed by something other than an ESS move of this library (the
// exact Hetero draw): its cached L^-1 y no longer matches, the next threshold solves with the factor again.
extern "C" int dgpb_cache_output_changed(dgpb_ws* ws, int key) {
    DGPB_REQUIRE(ws != nullptr, "ws is NULL");
    auto it = ws->cache.find(key);
    if (it != ws->cache.end()) it->second.w_synced = false;
    auto io = ws->owner.find(key);
    if (io != ws->owner.end()) io->second.w_synced = false;
    return DGPB_OK;
}

extern "C" int dgpb_ess_block(dgpb_ws* ws, const dgpb_node* targets, int n_targets, const int32_t* target_rows_host,
                              double* layer_out, int64_t layer_width, const dgpb_node* uppers, int n_uppers, int64_t n,
                              const double* z, const double* u_host, int nu, int* n_prop_host, double* theta_host,
                              void* stream) {
    return dgpb_ess_block_cached(ws, targets, n_targets, target_rows_host, layer_out, layer_width, uppers, n_uppers, n, z,
                                 u_host, nu, n_prop_host, theta_host, nullptr, nullptr, nullptr, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// Likelihood layers (dgpsi/likelihood_class.py): the upper "nodes" of the last GP layer are likelihood nodes whose
// log-likelihood is a sum over data points of an elementwise function of one or two latent columns.
namespace dgpb {

constexpr int kMaxLik = 8;
constexpr int kLikWave = 8;

struct LikArgs {
    int nl;
    int kind[kMaxLik];
    int n_in[kMaxLik];
    int row[kMaxLik][DGPB_LIK_MAX_IN];
    double param[kMaxLik];
    const double* y[kMaxLik];
    const double* src[kLikWave + 1];   // candidate images of the feeding layer (layer_width x n)
};

// log Phi(x): erfc keeps full relative accuracy in the lower tail until it underflows; below that the
// asymptotic series  -x^2/2 - log(-x sqrt(2 pi)) + log(1 - 1/x^2 + 3/x^4 - 15/x^6 + 105/x^8)
__device__ __forceinline__ double log_ndtr_dev(double x) {
    if (x > -35.0) return log(0.5 * erfc(-x * 0.70710678118654752440));
    const double r = 1.0 / (x * x);
    return -0.5 * x * x - log(-x) - 0.91893853320467274178 + log1p(r * (-1.0 + r * (3.0 + r * (-15.0 + r * 105.0))));
}

// log-likelihood of one data point; f(j) = j-th latent input of the node
template <typename F>
__device__ __forceinline__ double lik_point(int kind, int K, double param, double y, F f) {
    if (kind == DGPB_LIK_POISSON) {        // likelihood_class.py:39-48
        const double f0 = f(0);
        return y * f0 - exp(f0) - lgamma(y + 1.0);
    } else if (kind == DGPB_LIK_HETERO) {  // likelihood_class.py:110-116
        const double f0 = f(0), f1 = f(1);
        const double r2 = (y - f0) * (y - f0);
        return -0.5 * (1.8378770664093453 + f1 + exp(log(r2) - f1));
    } else if (kind == DGPB_LIK_NEGBIN) {  // likelihood_class.py:264-272
        const double f0 = f(0), f1 = f(1);
        const double nn = exp(-f1), a = f0 + f1;
        const double softplus = fmax(a, 0.0) + log1p(exp(-fabs(a)));
        return lgamma(y + nn) - lgamma(nn) - lgamma(y + 1.0) + y * a - (y + nn) * softplus;
    } else if (kind == DGPB_LIK_ZIP || kind == DGPB_LIK_ZINB) {  // likelihood_class.py:497-525, 653-693
        double base, fpi;  // log-density of the count part at y, logit of the zero-inflation probability
        if (kind == DGPB_LIK_ZIP) {
            const double f0 = f(0);
            fpi = f(1);
            base = y == 0.0 ? -exp(f0) : -exp(f0) + y * f0 - lgamma(y + 1.0);
        } else {
            const double f0 = f(0), f1 = f(1);
            fpi = f(2);
            const double nn = exp(-f1), a = f0 + f1;
            const double softplus = fmax(a, 0.0) + log1p(exp(-fabs(a)));
            base = lgamma(y + nn) - lgamma(nn) - lgamma(y + 1.0) + y * a - (y + nn) * softplus;
        }
        const double pi = 1.0 / (1.0 + exp(-fpi));
        const double l1m = log1p(-pi) + base;
        if (y != 0.0) return l1m;
        const double lp = log(pi);
        return fmax(lp, l1m) + log1p(exp(-fabs(lp - l1m)));
    } else if (kind == DGPB_LIK_CAT_LOGIT) {   // likelihood_class.py:339-341
        const double f0 = f(0);
        return y * f0 - (fmax(f0, 0.0) + log1p(exp(-fabs(f0))));
    } else if (kind == DGPB_LIK_CAT_PROBIT) {  // likelihood_class.py:342-343
        const double f0 = f(0);
        return y * log_ndtr_dev(f0) + (1.0 - y) * log_ndtr_dev(-f0);
    } else if (kind == DGPB_LIK_CAT_ROBUSTMAX) {  // likelihood_class.py:345-353; argmax = first maximum
        int best = 0;
        double fb = f(0);
        for (int j = 1; j < K; ++j) {
            const double fj = f(j);
            if (fj > fb) { fb = fj; best = j; }
        }
        return best == (int)y ? log(1.0 - param) : log(param / (double)(K - 1));
    } else {                                      // softmax, likelihood_class.py:354-358
        double mx = f(0);
        for (int j = 1; j < K; ++j) mx = fmax(mx, f(j));
        double se = 0.0;
        for (int j = 0; j < K; ++j) se += exp(f(j) - mx);
        return f((int)y) - (log(se) + mx);
    }
}

// out[s] = sum over likelihood nodes (in order) and data points of the log-likelihood of candidate s
__global__ void __launch_bounds__(512) lik_sum_kernel(LikArgs a, int64_t n, double* __restrict__ out) {
    __shared__ double red[512];
    const double* F = a.src[blockIdx.x];
    double total = 0.0;
    for (int l = 0; l < a.nl; ++l) {
        const double* y = a.y[l];
        const int kind = a.kind[l], K = a.n_in[l];
        const double param = a.param[l];
        double acc = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += 512)
            acc += lik_point(kind, K, param, y[i], [&](int j) { return F[(int64_t)a.row[l][j] * n + i]; });
        red[threadIdx.x] = acc;
        __syncthreads();
        for (int w = 256; w > 0; w >>= 1) {
            if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
            __syncthreads();
        }
        total += red[0];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = total;
}

// validate the descriptors and pack them for the kernel
static int pack_liks(const dgpb_lik* liks, int n_liks, int64_t layer_width, LikArgs* a) {
    DGPB_REQUIRE(n_liks >= 1 && n_liks <= kMaxLik, "bad number of likelihood nodes");
    a->nl = n_liks;
    for (int l = 0; l < n_liks; ++l) {
        const int kind = liks[l].kind;
        DGPB_REQUIRE(kind >= DGPB_LIK_POISSON && kind <= DGPB_LIK_ZINB && liks[l].y, "bad likelihood node");
        const int need = (kind == DGPB_LIK_POISSON || kind == DGPB_LIK_CAT_LOGIT || kind == DGPB_LIK_CAT_PROBIT) ? 1
                         : (kind == DGPB_LIK_HETERO || kind == DGPB_LIK_NEGBIN || kind == DGPB_LIK_ZIP) ? 2
                         : kind == DGPB_LIK_ZINB ? 3 : liks[l].n_in;
        DGPB_REQUIRE(liks[l].n_in == need && need >= 1 && need <= DGPB_LIK_MAX_IN, "bad number of likelihood inputs");
        DGPB_REQUIRE((kind != DGPB_LIK_CAT_SOFTMAX && kind != DGPB_LIK_CAT_ROBUSTMAX) || need >= 2,
                     "a multi-class likelihood needs at least two inputs");
        for (int j = 0; j < need; ++j) {
            DGPB_REQUIRE(liks[l].rows[j] >= 0 && (layer_width <= 0 || liks[l].rows[j] < layer_width),
                         "likelihood input row out of range");
            a->row[l][j] = liks[l].rows[j];
        }
        a->kind[l] = kind;
        a->n_in[l] = need;
        a->param[l] = liks[l].param;
        a->y[l] = liks[l].y;
    }
    return DGPB_OK;
}

}  // namespace dgpb

extern "C" int dgpb_lik_loglik(const dgpb_lik* liks, int n_liks, const double* layer, int64_t n, double* out_host,
                               void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(liks && layer && out_host && n >= 1, "bad argument");
    LikArgs a;
    DGPB_TRY(pack_liks(liks, n_liks, 0, &a));
    a.src[0] = layer;
    double* outd;
    DGPB_CUDA_TRY(cudaMallocAsync((void**)&outd, sizeof(double), st));
    lik_sum_kernel<<<1, 512, 0, st>>>(a, n, outd);
    DGPB_LAUNCHED();
    DGPB_CUDA_TRY(cudaMemcpyAsync(out_host, outd, sizeof(double), cudaMemcpyDeviceToHost, st));
    DGPB_CUDA_TRY(cudaFreeAsync(outd, st));
    DGPB_CUDA_TRY(cudaStreamSynchronize(st));
    return DGPB_OK;
}

// ESS update of target GP nodes whose outputs feed likelihood nodes (imputation.py:44-119 / :166-221 with
// `linked_kernel.type == 'likelihood'`).  Same contract as dgpb_ess_block for draws, angles and uniforms; candidate
// angles are evaluated in waves of up to 8 by one reduction launch, the threshold rides along in the first wave.
extern "C" int dgpb_ess_block_lik(dgpb_ws* ws, const dgpb_node* targets, int n_targets, const int32_t* target_rows_host,
                                  double* layer_out, int64_t layer_width, const dgpb_lik* liks, int n_liks, int64_t n,
                                  const double* z, const double* u_host, int nu, int* n_prop_host, double* theta_host,
                                  const int32_t* target_keys_host, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(ws && targets && liks && layer_out && z && u_host && target_rows_host, "NULL argument");
    DGPB_REQUIRE(n_targets >= 1 && n_liks >= 1 && n_liks <= kMaxLik && n >= 1 && nu >= 3, "bad sizes");
    for (int k = 0; k < n_targets; ++k)
        DGPB_REQUIRE(target_rows_host[k] >= 0 && target_rows_host[k] < layer_width, "target row out of range");
    LikArgs a;
    DGPB_TRY(pack_liks(liks, n_liks, layer_width, &a));
    DGPB_TRY(join_wave_staging(ws, st));   // SLOT_PROP / SLOT_NU may still be read by a wave staged ahead
    void *pnu, *pprop, *pout;
    const size_t layer_elems = (size_t)layer_width * n;
    DGPB_TRY(ws->reserve(SLOT_NU, sizeof(double) * (size_t)n_targets * n, &pnu));
    DGPB_TRY(ws->reserve(SLOT_PROP, sizeof(double) * layer_elems * kLikWave, &pprop));
    DGPB_TRY(ws->reserve(SLOT_OUT, sizeof(double) * kOutDoubles, &pout));
    double* nuv = (double*)pnu;
    double* prop = (double*)pprop;
    double* outd = (double*)pout;
    DGPB_TRY(prior_draws(ws, targets, n_targets, n, z, nuv, target_keys_host, st, ws->wave_cur));
    for (int s = 0; s < kLikWave; ++s)  // rows that are not being updated are shared by every proposal
        DGPB_CUDA_TRY(cudaMemcpyAsync(prop + s * layer_elems, layer_out, sizeof(double) * layer_elems,
                                      cudaMemcpyDeviceToDevice, st));
    int ui = 0;
    const double log_u0 = log(u_host[ui++]);              // imputation.py:79
    double theta = 2.0 * M_PI * u_host[ui++];             // imputation.py:81
    double tmin = theta - 2.0 * M_PI, tmax = theta;
    double log_y = 0.0;
    bool have_thr = false;
    int nprop = 0;
    while (true) {
        double thetas[kLikWave];
        const int S = std::max(1, std::min(kLikWave, 1 + (nu - ui)));
        thetas[0] = theta;
        {
            double lmin = tmin, lmax = tmax;
            for (int s = 1; s < S; ++s) {
                if (thetas[s - 1] < 0.0) lmin = thetas[s - 1]; else lmax = thetas[s - 1];
                thetas[s] = lmin + (lmax - lmin) * u_host[ui + s - 1];
            }
        }
        for (int s = 0; s < S; ++s) {
            const double c = cos(thetas[s]), sn = sin(thetas[s]);
            for (int k = 0; k < n_targets; ++k) {
                const int64_t row = target_rows_host[k];
                propose_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(prop + s * layer_elems + row * n, layer_out + row * n,
                                                                      nuv + (int64_t)k * n, c, sn, n);
                DGPB_LAUNCHED();
            }
        }
        int first = 0;
        if (!have_thr) a.src[first++] = layer_out;
        for (int s = 0; s < S; ++s) a.src[first + s] = prop + s * layer_elems;
        lik_sum_kernel<<<first + S, 512, 0, st>>>(a, n, outd);
        DGPB_LAUNCHED();
        DGPB_CUDA_TRY(cudaMemcpyAsync(ws->pinned, outd, sizeof(double) * (first + S), cudaMemcpyDeviceToHost, st));
        DGPB_CUDA_TRY(cudaStreamSynchronize(st));
        if (!have_thr) {
            log_y = ws->pinned[0] + log_u0;
            have_thr = true;
        }
        int accepted = -1;
        for (int s = 0; s < S; ++s) {
            if (theta_host) theta_host[nprop] = thetas[s];
            ++nprop;
            if (ws->pinned[first + s] > log_y) {  // imputation.py:107-110 (a NaN likelihood is a rejection there too)
                accepted = s;
                break;
            }
            if (thetas[s] < 0.0) tmin = thetas[s]; else tmax = thetas[s];
            if (s + 1 < S) ++ui;
        }
        if (accepted >= 0) {
            const double* pa = prop + accepted * layer_elems;
            for (int k = 0; k < n_targets; ++k) {
                const int64_t row = target_rows_host[k];
                DGPB_CUDA_TRY(cudaMemcpyAsync(layer_out + row * n, pa + row * n, sizeof(double) * n,
                                              cudaMemcpyDeviceToDevice, st));
            }
            DGPB_TRY(rotate_cached_w(ws, targets, n_targets, target_keys_host, n, z, thetas[accepted], st));
            break;
        }
        if (ui >= nu) {
            if (n_prop_host) *n_prop_host = nprop;
            set_error("ESS ran out of uniforms after %d proposals", nprop);
            return DGPB_BAD_ARG;
        }
        theta = tmin + (tmax - tmin) * u_host[ui++];
    }
    if (n_prop_host) *n_prop_host = nprop;
    return DGPB_OK;
}

// knn.cu -- exact FP64 nearest-neighbour search on sm_100a  (nn / get_pred_nn, dgpsi/vecchia.py:20-109).
//
// The search is brute force (10-20 dimensional inputs leave nothing to prune) in two stages:
//   1. SCREEN on the FP64 tensor path.  |q - x|^2 = |q|^2 + |x|^2 - 2 q.x ; the cross term is a DMMA product
//      (mma.m8n8k4: 8 queries x 8 candidates x 4 dimensions per instruction), the norms initialise the
//      accumulator.  Every query keeps the LC = 32 or 64 smallest screened distances in a sorted
//      shared-memory list owned by its warp (insertions become rare after the first few hundred candidates and
//      are done by the whole warp: ballot for the slot, shfl_up for the tail, last entry = the query's threshold).
//   2. RANK exactly.  The LC survivors of a query are re-evaluated with the reference arithmetic -- sum_k
//      (q_k - x_k)^2, separate multiply and add in ascending k, what the CPU libraries compute --
//      and ranked by (distance, index), so the index sets match bit for bit and ties keep the smaller index.
// The screen can only lose a true neighbour if its rounding error exceeds the gap between the exact m-th
// distance and the screened LC-th one; the kernel checks that gap against a rigorous error bound per query and
// flags the query otherwise (exact duplicates / lattice ties), and flagged queries are redone by the scalar
// exact kernel below.  So the output is always that of the exact search.
#include "common.cuh"
#include "vecchia.cuh"

namespace dgpb {

// ------------------------------------------------------------------------------------------------
// scalar exact kernel (fallback, and the path for conditioning sets larger than the screen's lists).
// Each thread owns one query and scans candidate tiles staged in shared memory, keeping its current
// m best in a sorted private list.  `only` (optional): process flagged queries only.
// ------------------------------------------------------------------------------------------------
template <int DMAX, bool ORDERED>
__global__ void __launch_bounds__(128) knn_kernel(const double* __restrict__ q, int64_t M, const double* __restrict__ x,
                                                  int64_t n, int D, int m, int64_t* __restrict__ NN, int ldnn,
                                                  const unsigned char* __restrict__ only) {
    constexpr int TC = 128;
    __shared__ double xc[TC][DMAX];
    const int tid = threadIdx.x;
    const int64_t qi = (int64_t)blockIdx.x * 128 + tid;
    const bool active = qi < M && (only == nullptr || only[qi] != 0);
    if (only != nullptr && !__syncthreads_or(active)) return;
    double qv[DMAX];
#pragma unroll
    for (int k = 0; k < DMAX; ++k) qv[k] = (active && k < D) ? q[qi * D + k] : 0.0;
    double bd[kMaxBlock];
    int bi[kMaxBlock];
    int cnt = 0;
    double thr = INFINITY;
    // ORDERED: candidates are j < i only (vecchia.py:42-51,84-107)
    const int64_t cmax = ORDERED ? min(n, (int64_t)blockIdx.x * 128 + 128) : n;
    for (int64_t c0 = 0; c0 < cmax; c0 += TC) {
        __syncthreads();
        for (int idx = tid; idx < TC * DMAX; idx += 128) {
            int r = idx / DMAX, k = idx % DMAX;
            int64_t j = c0 + r;
            xc[r][k] = (j < n && k < D) ? x[j * D + k] : 0.0;
        }
        __syncthreads();
        if (!active) continue;
        int64_t lim = min((int64_t)TC, (ORDERED ? qi : n) - c0);
        for (int r = 0; r < lim; ++r) {
            double dist = 0.0;
#pragma unroll
            for (int k = 0; k < DMAX; ++k) {
                double df = __dsub_rn(qv[k], xc[r][k]);
                dist = __dadd_rn(dist, __dmul_rn(df, df));
            }
            if (m > 0 && (cnt < m || dist < thr)) {
                int pos = cnt < m ? cnt : m - 1;
                while (pos > 0 && bd[pos - 1] > dist) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = dist;
                bi[pos] = (int)(c0 + r);
                if (cnt < m) ++cnt;
                if (cnt == m) thr = bd[m - 1];
            }
        }
    }
    if (!active) return;
    if (ORDERED) {
        // row = {i} U neighbours, sorted by index descending, -1 padded
        for (int a = 1; a < cnt; ++a) {  // insertion sort of indices, descending
            int v = bi[a], p = a;
            while (p > 0 && bi[p - 1] < v) {
                bi[p] = bi[p - 1];
                --p;
            }
            bi[p] = v;
        }
        NN[qi * ldnn] = qi;
        for (int a = 0; a < ldnn - 1; ++a) NN[qi * ldnn + 1 + a] = a < cnt ? (int64_t)bi[a] : -1;
    } else {
        for (int a = 0; a < m; ++a) NN[qi * ldnn + a] = (int64_t)bi[a];
    }
}

__global__ void knn_all_kernel(int64_t M, int m, int64_t* NN) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * m) return;
    int64_t k = idx / m, c = idx % m;
    NN[idx] = (c + k) % m;  // (arange(m)+arange(k)[:,None]) % m   vecchia.py:23-26
}

template <bool ORDERED>
static int launch_knn(const double* q, int64_t M, const double* x, int64_t n, int D, int m, int64_t* NN, int ldnn,
                      const unsigned char* only, cudaStream_t st) {
    unsigned grid = (unsigned)cdiv(M, 128);
#define KNN_CASE(DM)                                                                            \
    if (D <= DM) {                                                                              \
        knn_kernel<DM, ORDERED><<<grid, 128, 0, st>>>(q, M, x, n, D, m, NN, ldnn, only);        \
        DGPB_LAUNCHED();                                                                        \
        return DGPB_OK;                                                                         \
    }
    KNN_CASE(2) KNN_CASE(4) KNN_CASE(8) KNN_CASE(12) KNN_CASE(16) KNN_CASE(24) KNN_CASE(32)
#undef KNN_CASE
    set_error("kNN: dimension %d > %d", D, kMaxDim);
    return DGPB_BAD_ARG;
}

// ------------------------------------------------------------------------------------------------
// tensor-core screen + exact ranking
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void knn_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void knn_dmma_init(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
                 : "=d"(d0), "=d"(d1)
                 : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
__device__ __forceinline__ void knn_cp8(void* smem_dst, const void* gsrc, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = pred ? 8 : 0;  // src-size 0 -> zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}

// Offer (d, j) to the list (ld, li) of 32 * NL entries kept SORTED ascending, slot s = lane + 32 e; d, j and thr
// are warp-uniform.  One ballot per 32 slots gives the insertion point, the tail moves up one slot through
// shfl_up, and the new last entry is the query's threshold: a chain of four dependent steps (the unsorted
// arg-max / replace / re-max form was ten and bounded the kernel).  Returns the new threshold.
template <int NL>
__device__ __forceinline__ double knn_list_insert(double* ld, int* li, int lane, double d, int j, double thr) {
    if (!(d < thr)) return thr;
    double v[NL];
    int ix[NL];
    int p = 0;   // number of entries <= d: the slot the candidate takes (ties stay behind earlier candidates)
#pragma unroll
    for (int e = 0; e < NL; ++e) {
        v[e] = ld[lane + 32 * e];
        ix[e] = li[lane + 32 * e];
        p += __popc(__ballot_sync(0xffffffffu, v[e] <= d));
    }
    double last = 0.0;
#pragma unroll
    for (int e = 0; e < NL; ++e) {
        const int slot = lane + 32 * e;
        double pv = __shfl_up_sync(0xffffffffu, v[e], 1);
        int pi = __shfl_up_sync(0xffffffffu, ix[e], 1);
        if (e > 0) {   // slot 32 e takes the last entry of the previous group
            const double cv = __shfl_sync(0xffffffffu, v[e - 1], 31);
            const int ci = __shfl_sync(0xffffffffu, ix[e - 1], 31);
            if (lane == 0) {
                pv = cv;
                pi = ci;
            }
        }
        double nv = v[e];
        int ni = ix[e];
        if (slot == p) {
            nv = d;
            ni = j;
        } else if (slot > p) {
            nv = pv;
            ni = pi;
        }
        if (slot >= p) {
            ld[slot] = nv;
            li[slot] = ni;
        }
        if (e == NL - 1) last = __shfl_sync(0xffffffffu, nv, 31);
    }
    __syncwarp();
    return last;
}

constexpr int kKnnTC = 128;   // candidates per staged tile
constexpr int kKnnQ = 128;    // queries per CTA: 8 warps x 16

// row stride (doubles) of a staged candidate: = 4 or 12 mod 16, so the 8 x 4 fragment loads are conflict-free
template <int KS>
struct KnnPad {
    static constexpr int value = ((4 * KS) % 16 == 4 || (4 * KS) % 16 == 12) ? 4 * KS : 4 * KS + 4;
};
template <int KS, int NL>
constexpr size_t knn_smem_bytes() {
    return (size_t)2 * kKnnTC * KnnPad<KS>::value * 8 + 2 * kKnnTC * 8 + (size_t)kKnnQ * 32 * NL * 12;
}

// KS = ceil(D / 4) k-steps, NL = list entries per lane (list capacity 32 NL >= m + 3).
template <int KS, int NL, bool ORDERED>
__global__ void __launch_bounds__(256, KS <= 3 && NL == 1 ? 3 : 2) knn_mma_kernel(const double* __restrict__ q, int64_t M, const double* __restrict__ x,
                                                      int64_t n, int D, int m, int64_t* __restrict__ NN, int ldnn,
                                                      unsigned char* __restrict__ flags, int dbg) {
    constexpr int DP = KnnPad<KS>::value;
    constexpr int LC = 32 * NL;
    extern __shared__ __align__(16) unsigned char knn_smem[];
    double* xs = reinterpret_cast<double*>(knn_smem);                   // [2][TC][DP] raw candidate coordinates
    double* xn = xs + 2 * kKnnTC * DP;                                  // [2][TC]     squared norms
    double* ldist = xn + 2 * kKnnTC;                                    // [Q][LC]     screened distances
    int* lidx = reinterpret_cast<int*>(ldist + (size_t)kKnnQ * LC);     // [Q][LC]     candidate indices
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int64_t q0 = (int64_t)blockIdx.x * kKnnQ;
    // zero both tile buffers once: the pad columns [D, DP) are never written again
    for (int i = tid; i < 2 * kKnnTC * DP; i += 256) xs[i] = 0.0;
    for (int i = tid; i < kKnnQ * LC; i += 256) {
        ldist[i] = INFINITY;
        lidx[i] = -1;
    }
    // A fragments: -2 q for the warp's two query octets
    double aq[2][KS];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int64_t qi = q0 + 16 * w + 8 * s + g;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int k = 4 * ks + t4;
            aq[s][ks] = (qi < M && k < D) ? -2.0 * q[qi * D + k] : 0.0;
        }
    }
    double thr[2] = {INFINITY, INFINITY};
    double xnmax = 0.0;
    // ORDERED: candidates are j < i only (vecchia.py:42-51,84-107)
    const int64_t cmax = ORDERED ? min(n, q0 + kKnnQ) : n;
    const int ntiles = (int)((cmax + kKnnTC - 1) / kKnnTC);
    auto load_tile = [&](int tile) {
        double* dst = xs + (size_t)(tile & 1) * kKnnTC * DP;
        const int64_t c0 = (int64_t)tile * kKnnTC;
        for (int i = tid; i < kKnnTC * D; i += 256) {
            const int r = i / D, k = i - r * D;
            const bool ok = c0 + r < n;
            knn_cp8(&dst[r * DP + k], ok ? &x[(c0 + r) * D + k] : x, ok);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    __syncthreads();
    if (ntiles > 0) load_tile(0);
    for (int tile = 0; tile < ntiles; ++tile) {
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();   // tile has landed for everybody; everybody is done with tile - 1
        const double* xt = xs + (size_t)(tile & 1) * kKnnTC * DP;
        double* xnt = xn + (tile & 1) * kKnnTC;
        if (tid < kKnnTC) {
            double sn = 0.0;
#pragma unroll
            for (int k = 0; k < 4 * KS; ++k) sn += xt[tid * DP + k] * xt[tid * DP + k];
            xnt[tid] = sn;
            xnmax = fmax(xnmax, sn);
        }
        __syncthreads();
        if (tile + 1 < ntiles) load_tile(tile + 1);
        const int64_t c0 = (int64_t)tile * kKnnTC;
        // candidates of this tile a query may take: local index < lim[s]
        int lim[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int64_t qi = q0 + 16 * w + 8 * s + g;
            const int64_t jl = (ORDERED ? min(n, qi) : n) - c0;
            lim[s] = qi < M ? (int)max((int64_t)0, min(jl, (int64_t)kKnnTC)) : 0;
        }
        // two candidate octets x two query octets per step: four independent DMMA chains in flight before the
        // first result is tested
        for (int cg = 0; cg < kKnnTC / 8; cg += 2) {
            double b[2][KS];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) b[h][ks] = xt[(8 * (cg + h) + g) * DP + 4 * ks + t4];
            double2 nn2[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                nn2[h] = *reinterpret_cast<const double2*>(&xnt[8 * (cg + h) + 2 * t4]);
            }
            // screened value = |x|^2 - 2 q.x  (the query's own |q|^2 is a per-list constant: it is left out of
            // the accumulator -- the candidate norms feed the first DMMA as its C operand -- and added back only
            // where a true distance is needed, the loss check)
            double d[2][2][2];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int s = 0; s < 2; ++s) knn_dmma_init(d[h][s][0], d[h][s][1], aq[s][0], b[h][0], nn2[h].x, nn2[h].y);
#pragma unroll
            for (int ks = 1; ks < KS; ++ks)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int s = 0; s < 2; ++s) knn_dmma(d[h][s][0], d[h][s][1], aq[s][ks], b[h][ks]);
            // one vote for the whole step; the per-tile work below only runs when some lane has a hit
            bool anyhit = false;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int jl0 = 8 * (cg + h) + 2 * t4;
                    anyhit = anyhit || (d[h][s][0] < thr[s] && jl0 < lim[s]) || (d[h][s][1] < thr[s] && jl0 + 1 < lim[s]);
                }
            if (dbg && tile > 0) anyhit = false;   // probe only: scan cost without list maintenance (results invalid)
            if (!__any_sync(0xffffffffu, anyhit)) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int jl0 = 8 * (cg + h) + 2 * t4;
                    const bool h0 = d[h][s][0] < thr[s] && jl0 < lim[s];
                    const bool h1 = d[h][s][1] < thr[s] && jl0 + 1 < lim[s];
                    unsigned hm = __ballot_sync(0xffffffffu, h0 || h1);
                    while (hm) {   // rare after warm-up: offer the hits one at a time, the whole warp helping
                        const int src = __ffs(hm) - 1;
                        hm &= hm - 1;
                        const int ql = 16 * w + 8 * s + (src >> 2);
                        const int jj = (int)(c0 + 8 * (cg + h) + 2 * (src & 3));
                        const double e0 = __shfl_sync(0xffffffffu, d[h][s][0], src);
                        const double e1 = __shfl_sync(0xffffffffu, d[h][s][1], src);
                        const int f = __shfl_sync(0xffffffffu, (int)h0 | ((int)h1 << 1), src);
                        double tq = __shfl_sync(0xffffffffu, thr[s], src);
                        if (f & 1) tq = knn_list_insert<NL>(ldist + (size_t)ql * LC, lidx + (size_t)ql * LC, lane, e0, jj, tq);
                        if (f & 2) tq = knn_list_insert<NL>(ldist + (size_t)ql * LC, lidx + (size_t)ql * LC, lane, e1, jj + 1, tq);
                        if (g == (src >> 2)) thr[s] = tq;
                    }
                }
        }
    }
    // largest candidate norm seen (tracked by the threads that computed the norms) -> every thread
    __shared__ double s_xnmax[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xnmax = fmax(xnmax, __shfl_xor_sync(0xffffffffu, xnmax, o));
    if (lane == 0) s_xnmax[w] = xnmax;
    __syncthreads();
    xnmax = s_xnmax[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) xnmax = fmax(xnmax, s_xnmax[i]);
    // ---- exact ranking of the survivors: one query at a time per warp, list entry per lane
    for (int ql = 16 * w; ql < 16 * w + 16; ++ql) {
        const int64_t qi = q0 + ql;
        if (qi >= M) break;
        double dv[NL];
        int jv[NL];
        double smax = 0.0;      // screened LC-th distance (the list's maximum)
        int filled = 0;
#pragma unroll
        for (int e = 0; e < NL; ++e) {
            const int j = lidx[(size_t)ql * LC + lane + 32 * e];
            jv[e] = j;
            smax = fmax(smax, ldist[(size_t)ql * LC + lane + 32 * e]);
            double dist = INFINITY;
            if (j >= 0) {
                dist = 0.0;
                for (int k = 0; k < D; ++k) {
                    const double df = __dsub_rn(q[qi * D + k], x[(int64_t)j * D + k]);
                    dist = __dadd_rn(dist, __dmul_rn(df, df));
                }
                ++filled;
            }
            dv[e] = dist;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            smax = fmax(smax, __shfl_xor_sync(0xffffffffu, smax, o));
            filled += __shfl_xor_sync(0xffffffffu, filled, o);
        }
        // rank by (distance, index)
        int rank[NL];
#pragma unroll
        for (int e = 0; e < NL; ++e) rank[e] = 0;
#pragma unroll
        for (int e2 = 0; e2 < NL; ++e2)
            for (int l2 = 0; l2 < 32; ++l2) {
                const double od = __shfl_sync(0xffffffffu, dv[e2], l2);
                const int oj = __shfl_sync(0xffffffffu, jv[e2], l2);
#pragma unroll
                for (int e = 0; e < NL; ++e)
                    rank[e] += (oj >= 0) && (od < dv[e] || (od == dv[e] && oj < jv[e]));
            }
        const int take = min(m, filled);
        // exact m-th distance and the loss check
        double dm = 0.0;
#pragma unroll
        for (int e = 0; e < NL; ++e)
            if (jv[e] >= 0 && rank[e] < take) dm = fmax(dm, dv[e]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
        if (lane == 0 && filled == LC) {
            // candidates outside the list have screened distance >= smax; |screened - exact| <= eps
            double qq = 0.0;
            for (int k = 0; k < D; ++k) qq += q[qi * D + k] * q[qi * D + k];
            const double eps = 8.0 * (double)(D + 4) * 1.1102230246251565e-16 * (qq + xnmax);
            if (!(smax + qq - eps > dm)) flags[qi] = 1;   // lists hold |x|^2 - 2 q.x
        }
        if (ORDERED) {
            // row = {i} U the m nearest j < i, sorted by index descending, -1 padded
            if (lane == 0) NN[qi * ldnn] = qi;
            for (int a = lane; a < ldnn - 1; a += 32)
                if (a >= take) NN[qi * ldnn + 1 + a] = -1;
#pragma unroll
            for (int e = 0; e < NL; ++e) {
                const bool sel = jv[e] >= 0 && rank[e] < take;
                int r2 = 0;
#pragma unroll
                for (int e2 = 0; e2 < NL; ++e2)
                    for (int l2 = 0; l2 < 32; ++l2) {
                        const int oj = __shfl_sync(0xffffffffu, jv[e2], l2);
                        const int orank = __shfl_sync(0xffffffffu, rank[e2], l2);
                        r2 += (oj >= 0 && orank < take && oj > jv[e]);
                    }
                if (sel) NN[qi * ldnn + 1 + r2] = jv[e];
            }
        } else {
#pragma unroll
            for (int e = 0; e < NL; ++e)
                if (jv[e] >= 0 && rank[e] < take) NN[qi * ldnn + rank[e]] = jv[e];
        }
    }
}

static int g_knn_mma = 1;   // dgpb_tune("knn_mma", 0) forces the scalar exact kernel (tests compare the two)

template <bool ORDERED>
static int knn_search(Workspace* ws, const double* q, int64_t M, const double* x, int64_t n, int D, int m, int64_t* NN,
                      int ldnn, cudaStream_t st) {
    const int KS = (D + 3) / 4;
    const int NL = (m + 3 <= 32) ? 1 : ((m + 3 <= 64) ? 2 : 0);
    if (!g_knn_mma || NL == 0 || m == 0) return launch_knn<ORDERED>(q, M, x, n, D, m, NN, ldnn, nullptr, st);
    void* pf;
    DGPB_TRY(ws->reserve(SLOT_VFLAG, (size_t)M + 16, &pf));
    unsigned char* flags = (unsigned char*)pf;
    DGPB_CUDA_TRY(cudaMemsetAsync(flags, 0, (size_t)M, st));
    const unsigned grid = (unsigned)cdiv(M, kKnnQ);
#define KNN_MMA_CASE(KSV, NLV)                                                                                      \
    if (KS <= KSV && NL == NLV) {                                                                                   \
        static bool cfg = false;                                                                                    \
        if (!cfg) {                                                                                                 \
            DGPB_CUDA_TRY(cudaFuncSetAttribute(knn_mma_kernel<KSV, NLV, ORDERED>,                                   \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,                        \
                                               (int)knn_smem_bytes<KSV, NLV>()));                                   \
            cfg = true;                                                                                             \
        }                                                                                                           \
        knn_mma_kernel<KSV, NLV, ORDERED><<<grid, 256, knn_smem_bytes<KSV, NLV>(), st>>>(q, M, x, n, D, m, NN, ldnn, \
                                                                                        flags, g_knn_mma == 2);                     \
        DGPB_LAUNCHED();                                                                                            \
        return launch_knn<ORDERED>(q, M, x, n, D, m, NN, ldnn, flags, st);                                          \
    }
    KNN_MMA_CASE(1, 1) KNN_MMA_CASE(2, 1) KNN_MMA_CASE(3, 1) KNN_MMA_CASE(5, 1) KNN_MMA_CASE(8, 1)
    KNN_MMA_CASE(1, 2) KNN_MMA_CASE(2, 2) KNN_MMA_CASE(3, 2) KNN_MMA_CASE(5, 2) KNN_MMA_CASE(8, 2)
#undef KNN_MMA_CASE
    return launch_knn<ORDERED>(q, M, x, n, D, m, NN, ldnn, nullptr, st);
}

int knn_set_mma(int on) {
    g_knn_mma = on;   // 0 = scalar kernel, 1 = screen + rank, 2 = probe: screen without list maintenance (invalid results)
    return DGPB_OK;
}

static int knn_tls_ws(dgpb_ws** out) {
    static thread_local dgpb_ws* ws = nullptr;
    if (!ws) {
        int dev = 0;
        DGPB_CUDA_TRY(cudaGetDevice(&dev));
        DGPB_TRY(dgpb_ws_create(&ws, dev));
    }
    *out = ws;
    return DGPB_OK;
}

}  // namespace dgpb

using namespace dgpb;

extern "C" {

int dgpb_knn_ordered(const double* x, int64_t n, int64_t D, int64_t m, int64_t* NN, void* stream) {
    DGPB_REQUIRE(x && NN && n >= 1 && D >= 1 && D <= kMaxDim, "bad argument");
    m = std::min(m, n - 1);
    DGPB_REQUIRE(m >= 0 && m < kMaxBlock, "m out of range");
    dgpb_ws* ws;
    DGPB_TRY(knn_tls_ws(&ws));
    if (m == 0) {
        // single point: NN = [[0]]
        return launch_knn<true>(x, n, x, n, (int)D, 0, NN, 1, nullptr, (cudaStream_t)stream);
    }
    return knn_search<true>(ws, x, n, x, n, (int)D, (int)m, NN, (int)m + 1, (cudaStream_t)stream);
}

int dgpb_knn(const double* query, int64_t M, const double* x, int64_t n, int64_t D, int64_t m, int64_t* NN,
             void* stream) {
    DGPB_REQUIRE(query && x && NN && n >= 1 && D >= 1 && D <= kMaxDim, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    m = std::min(m, n);
    if (M == 0) return DGPB_OK;
    if (m == n) {
        knn_all_kernel<<<(unsigned)cdiv(M * m, 256), 256, 0, st>>>(M, (int)m, NN);
        DGPB_LAUNCHED();
        return DGPB_OK;
    }
    DGPB_REQUIRE(m >= 1 && m <= kMaxBlock, "m out of range (max 64)");
    dgpb_ws* ws;
    DGPB_TRY(knn_tls_ws(&ws));
    return knn_search<false>(ws, query, M, x, n, (int)D, (int)m, NN, (int)m, st);
}

}  // extern "C"

// knn.cu -- exact FP64 nearest-neighbour search on sm_100a  (nn / get_pred_nn, dgpsi/vecchia.py:20-109).
//
// The search is brute force (10-20 dimensional inputs leave nothing to prune) in two stages:
//   1. SCREEN on the tensor path (FP64 DMMA form described here; the default is the split-TF32 form further down,
//      which keeps everything but the arithmetic of the cross term).  |q - x|^2 = |q|^2 + |x|^2 - 2 q.x ; the cross term is a DMMA product
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

// Exact ranking of one query's survivors (list entry per lane): distances re-evaluated with the reference
// arithmetic, ranked by (distance, index), written in the reference's output format.  `bound`: every candidate the
// screen dropped has exact distance >= bound (meaningful when the list is full); if the exact m-th distance does
// not stay below it the query is flagged for the scalar exact kernel.
template <int NL, bool ORDERED>
__device__ __forceinline__ void knn_rank_and_write(const double* __restrict__ q, const double* __restrict__ x, int64_t qi,
                                                   int D, int m, const int* __restrict__ lidx_q, double bound,
                                                   int64_t* __restrict__ NN, int ldnn, unsigned char* __restrict__ flags,
                                                   int lane, int istride = 1) {
    constexpr int LC = 32 * NL;
    double dv[NL];
    int jv[NL];
    int filled = 0;
#pragma unroll
    for (int e = 0; e < NL; ++e) {
        const int j = lidx_q[(lane + 32 * e) * istride];   // istride: survivor slots of one query (1, or slot-major sets)
        jv[e] = j;
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
    for (int o = 16; o > 0; o >>= 1) filled += __shfl_xor_sync(0xffffffffu, filled, o);
    int rank[NL];
#pragma unroll
    for (int e = 0; e < NL; ++e) rank[e] = 0;
#pragma unroll
    for (int e2 = 0; e2 < NL; ++e2)
        for (int l2 = 0; l2 < 32; ++l2) {
            const double od = __shfl_sync(0xffffffffu, dv[e2], l2);
            const int oj = __shfl_sync(0xffffffffu, jv[e2], l2);
#pragma unroll
            for (int e = 0; e < NL; ++e) rank[e] += (oj >= 0) && (od < dv[e] || (od == dv[e] && oj < jv[e]));
        }
    const int take = min(m, filled);
    double dm = 0.0;   // exact m-th distance
#pragma unroll
    for (int e = 0; e < NL; ++e)
        if (jv[e] >= 0 && rank[e] < take) dm = fmax(dm, dv[e]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
    if (lane == 0 && filled == LC && !(bound > dm)) flags[qi] = 1;
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
        double smax = 0.0;      // screened LC-th value (the list's maximum)
#pragma unroll
        for (int e = 0; e < NL; ++e) smax = fmax(smax, ldist[(size_t)ql * LC + lane + 32 * e]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) smax = fmax(smax, __shfl_xor_sync(0xffffffffu, smax, o));
        // candidates outside the list have screened value >= smax; lists hold |x|^2 - 2 q.x; |screened - exact| <= eps
        double qq = 0.0;
        for (int k = 0; k < D; ++k) qq += q[qi * D + k] * q[qi * D + k];
        const double eps = 8.0 * (double)(D + 4) * 1.1102230246251565e-16 * (qq + xnmax);
        knn_rank_and_write<NL, ORDERED>(q, x, qi, D, m, lidx + (size_t)ql * LC, smax + qq - eps, NN, ldnn, flags, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// Screen on the TF32 tensor path with split operands.  The screen only has to be right to within the gap between
// the exact m-th and the screened LC-th distance (a few per cent of the distance itself), so it does not need
// FP64: every coordinate (shifted by the first candidate, so offsets do not eat mantissa) is split into two TF32
// numbers v = hi + lo (relative error 2^-22), the cross term is hi.hi + hi.lo + lo.hi -- three
// mma.m16n8k8.tf32 with FP32 accumulation, products of 10-bit mantissas are exact in FP32 -- and the candidate
// norm initialises the accumulator.  16 queries x 8 candidates x 8 dimensions per instruction at the HMMA rate
// instead of 8 x 8 x 4 at the FP64 rate; the per-query lists, the exact FP64 ranking, the error-bound check
// (bound 2^-17 (|q| + |x|max)^2, four times the analytic estimate) and the scalar fallback are those of the DMMA
// version, so the result is still exactly that of the exact search.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned knn_tf32(float v) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void knn_hmma(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Offer (d, j) to the sorted FP32 list (see knn_list_insert).
template <int NL>
__device__ __forceinline__ float knn_list_insert_f(float* ld, int* li, int lane, float d, int j, float thr) {
    if (!(d < thr)) return thr;
    float v[NL];
    int ix[NL];
    int p = 0;
#pragma unroll
    for (int e = 0; e < NL; ++e) {
        v[e] = ld[lane + 32 * e];
        ix[e] = li[lane + 32 * e];
        p += __popc(__ballot_sync(0xffffffffu, v[e] <= d));
    }
    float last = 0.f;
#pragma unroll
    for (int e = 0; e < NL; ++e) {
        const int slot = lane + 32 * e;
        float pv = __shfl_up_sync(0xffffffffu, v[e], 1);
        int pi = __shfl_up_sync(0xffffffffu, ix[e], 1);
        if (e > 0) {
            const float cv = __shfl_sync(0xffffffffu, v[e - 1], 31);
            const int ci = __shfl_sync(0xffffffffu, ix[e - 1], 31);
            if (lane == 0) {
                pv = cv;
                pi = ci;
            }
        }
        float nv = v[e];
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

template <int KS8, int NL>
constexpr size_t knn_tf32_smem_bytes() {
    return (size_t)2 * kKnnTC * (8 * KS8 + 4) * 8 + 2 * kKnnTC * 4 + (size_t)kKnnQ * 32 * NL * 8 + 64;
}

// KS8 = ceil(D / 8) k-steps of 8 dimensions, NL = list entries per lane.
template <int KS8, int NL, bool ORDERED>
__global__ void __launch_bounds__(256, 2) knn_tf32_kernel(const double* __restrict__ q, int64_t M, const double* __restrict__ x,
                                                          int64_t n, int D, int m, int64_t* __restrict__ NN, int ldnn,
                                                          unsigned char* __restrict__ flags) {
    constexpr int DPF = 8 * KS8 + 4;   // float2 row stride: = 4 mod 16 -> conflict-free fragment loads
    constexpr int LC = 32 * NL;
    extern __shared__ __align__(16) unsigned char knn_smem[];
    // [2][TC][DPF] float2 = {dim t4, dim t4 + 4} of one k-step (dims permuted so a B fragment is ONE 8-byte load),
    // hi plane then lo plane
    float2* xs = reinterpret_cast<float2*>(knn_smem);
    float* xn = reinterpret_cast<float*>(xs + 2 * kKnnTC * DPF);         // [2][TC] squared norms
    float* ldist = xn + 2 * kKnnTC;                                      // [Q][LC] screened values
    int* lidx = reinterpret_cast<int*>(ldist + (size_t)kKnnQ * LC);      // [Q][LC]
    float* s_xnmax = reinterpret_cast<float*>(lidx + (size_t)kKnnQ * LC);  // [8]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int64_t q0 = (int64_t)blockIdx.x * kKnnQ;
    for (int i = tid; i < 2 * kKnnTC * DPF; i += 256) xs[i] = make_float2(0.f, 0.f);   // pad columns stay zero
    for (int i = tid; i < kKnnQ * LC; i += 256) {
        ldist[i] = INFINITY;
        lidx[i] = -1;
    }
    // A fragments: -2 (q - x_0), split; rows g and g + 8 of the warp's 16 queries
    unsigned ah[KS8][4], al[KS8][4];
#pragma unroll
    for (int ks = 0; ks < KS8; ++ks)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int64_t qi = q0 + 16 * w + g + 8 * (r & 1);
            const int k = 8 * ks + t4 + 4 * (r >> 1);
            const float v = (qi < M && k < D) ? (float)(-2.0 * (q[qi * D + k] - x[k])) : 0.f;
            ah[ks][r] = knn_tf32(v);
            al[ks][r] = knn_tf32(v - __uint_as_float(ah[ks][r]));
        }
    float thr[2] = {INFINITY, INFINITY};
    float xnmax = 0.f;
    const int64_t cmax = ORDERED ? min(n, q0 + kKnnQ) : n;
    const int ntiles = (int)((cmax + kKnnTC - 1) / kKnnTC);
    // staging: two threads per candidate, each converts half of the dimensions.  The FP64 loads of the NEXT tile
    // are issued into registers before the current tile is scanned and converted / stored after it, so their
    // latency is covered by the scan.
    constexpr int KH = 4 * KS8;
    const int sc = tid >> 1, sh = tid & 1;
    const int kh = (D + 1) / 2, k0s = sh * kh, k1s = min(D, k0s + kh);
    double stg[KH], shift[KH];
#pragma unroll
    for (int k = 0; k < KH; ++k) shift[k] = (k0s + k < k1s) ? x[k0s + k] : 0.0;
    auto load_tile = [&](int tile) {
        const int64_t j = (int64_t)tile * kKnnTC + sc;
#pragma unroll
        for (int k = 0; k < KH; ++k) stg[k] = (j < n && k0s + k < k1s) ? x[j * D + k0s + k] : shift[k];
    };
    auto store_tile = [&](int tile) {
        float2* dst = xs + (size_t)(tile & 1) * kKnnTC * DPF + sc * DPF;
        double sn = 0.0;
#pragma unroll
        for (int k = 0; k < KH; ++k) {
            const double v = stg[k] - shift[k];
            sn += v * v;
            const float hi = __uint_as_float(knn_tf32((float)v));
            if (k0s + k < k1s) {
                const int kk = k0s + k;   // dimension kk -> k-step kk/8, slot kk%4, half (kk/4)%2
                float* slot = reinterpret_cast<float*>(dst + (kk >> 3) * 8 + (kk & 3)) + ((kk >> 2) & 1);
                slot[0] = hi;                                                   // hi plane
                slot[8] = __uint_as_float(knn_tf32((float)(v - (double)hi)));   // lo plane: 4 float2 further
            }
        }
        sn += __shfl_xor_sync(0xffffffffu, sn, 1);
        if (sh == 0) {
            xn[(tile & 1) * kKnnTC + sc] = (float)sn;
            xnmax = fmaxf(xnmax, (float)sn);
        }
    };
    __syncthreads();
    if (ntiles > 0) {
        load_tile(0);
        store_tile(0);
    }
    for (int tile = 0; tile < ntiles; ++tile) {
        __syncthreads();   // tile staged; everybody done with the buffer that is refilled below
        if (tile + 1 < ntiles) load_tile(tile + 1);
        const float2* xt = xs + (size_t)(tile & 1) * kKnnTC * DPF;
        const float* xnt = xn + (tile & 1) * kKnnTC;
        const int64_t c0 = (int64_t)tile * kKnnTC;
        int lim[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int64_t qi = q0 + 16 * w + 8 * s + g;
            const int64_t jl = (ORDERED ? min(n, qi) : n) - c0;
            lim[s] = qi < M ? (int)max((int64_t)0, min(jl, (int64_t)kKnnTC)) : 0;
        }
        const bool full = __all_sync(0xffffffffu, lim[0] == kKnnTC && lim[1] == kKnnTC);
        for (int cg = 0; cg < kKnnTC / 8; cg += 2) {
            float c[2][4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float2 nn2 = *reinterpret_cast<const float2*>(&xnt[8 * (cg + h) + 2 * t4]);
                c[h][0] = c[h][2] = nn2.x;
                c[h][1] = c[h][3] = nn2.y;
            }
#pragma unroll
            for (int ks = 0; ks < KS8; ++ks)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 bh = xt[(8 * (cg + h) + g) * DPF + 8 * ks + t4];       // {dim t4, dim t4 + 4}, hi
                    const float2 bl = xt[(8 * (cg + h) + g) * DPF + 8 * ks + 4 + t4];   // the same two, lo
                    knn_hmma(c[h], ah[ks], __float_as_uint(bh.x), __float_as_uint(bh.y));
                    knn_hmma(c[h], ah[ks], __float_as_uint(bl.x), __float_as_uint(bl.y));
                    knn_hmma(c[h], al[ks], __float_as_uint(bh.x), __float_as_uint(bh.y));
                }
            bool anyhit = false;
            if (full) {   // every candidate of the tile is admissible for every query of the warp: no index tests
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int s = 0; s < 2; ++s) anyhit = anyhit || fminf(c[h][2 * s], c[h][2 * s + 1]) < thr[s];
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        const int jl0 = 8 * (cg + h) + 2 * t4;
                        anyhit = anyhit || (c[h][2 * s] < thr[s] && jl0 < lim[s]) || (c[h][2 * s + 1] < thr[s] && jl0 + 1 < lim[s]);
                    }
            }
            if (!__any_sync(0xffffffffu, anyhit)) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int jl0 = 8 * (cg + h) + 2 * t4;
                    const bool h0 = c[h][2 * s] < thr[s] && jl0 < lim[s];
                    const bool h1 = c[h][2 * s + 1] < thr[s] && jl0 + 1 < lim[s];
                    unsigned hm = __ballot_sync(0xffffffffu, h0 || h1);
                    while (hm) {
                        const int src = __ffs(hm) - 1;
                        hm &= hm - 1;
                        const int ql = 16 * w + 8 * s + (src >> 2);
                        const int jj = (int)(c0 + 8 * (cg + h) + 2 * (src & 3));
                        const float e0 = __shfl_sync(0xffffffffu, c[h][2 * s], src);
                        const float e1 = __shfl_sync(0xffffffffu, c[h][2 * s + 1], src);
                        const int f = __shfl_sync(0xffffffffu, (int)h0 | ((int)h1 << 1), src);
                        float tq = __shfl_sync(0xffffffffu, thr[s], src);
                        if (f & 1) tq = knn_list_insert_f<NL>(ldist + (size_t)ql * LC, lidx + (size_t)ql * LC, lane, e0, jj, tq);
                        if (f & 2) tq = knn_list_insert_f<NL>(ldist + (size_t)ql * LC, lidx + (size_t)ql * LC, lane, e1, jj + 1, tq);
                        if (g == (src >> 2)) thr[s] = tq;
                    }
                }
        }
        if (tile + 1 < ntiles) store_tile(tile + 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xnmax = fmaxf(xnmax, __shfl_xor_sync(0xffffffffu, xnmax, o));
    if (lane == 0) s_xnmax[w] = xnmax;
    __syncthreads();
    xnmax = s_xnmax[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) xnmax = fmaxf(xnmax, s_xnmax[i]);
    for (int ql = 16 * w; ql < 16 * w + 16; ++ql) {
        const int64_t qi = q0 + ql;
        if (qi >= M) break;
        float smaxf = 0.f;
#pragma unroll
        for (int e = 0; e < NL; ++e) smaxf = fmaxf(smaxf, ldist[(size_t)ql * LC + lane + 32 * e]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) smaxf = fmaxf(smaxf, __shfl_xor_sync(0xffffffffu, smaxf, o));
        // lists hold |x - x0|^2 - 2 (q - x0).(x - x0) in FP32; |screened - exact| <= eps
        double qq = 0.0;
        for (int k = 0; k < D; ++k) {
            const double v = q[qi * D + k] - x[k];
            qq += v * v;
        }
        const double rr = sqrt(qq) + sqrt((double)xnmax);
        const double eps = 7.62939453125e-06 * rr * rr;   // 2^-17 (|q| + |x|max)^2
        knn_rank_and_write<NL, ORDERED>(q, x, qi, D, m, lidx + (size_t)ql * LC, (double)smaxf + qq - eps, NN, ldnn, flags,
                                        lane);
    }
}

}  // namespace dgpb
#include "knn_tc5.cuh"
namespace dgpb {

// dgpb_tune("knn_mma", v): 0 scalar exact kernel, 1 FP64 DMMA screen, 3 warp-level split-TF32 screen (mma.sync),
// 5 split-TF32 screen on tcgen05 (TMA bulk copies, TMEM accumulators; default)
static int g_knn_mma = 5;

// tcgen05 screen: pack the candidates once, then one CTA per 128 queries.  Returns DGPB_OK with *done = 0 when the
// shape does not fit (too many dimensions for the shared-memory stages): the caller takes the warp-level kernel.
template <bool ORDERED>
static int knn_tc5_search(Workspace* ws, const double* q, int64_t M, const double* x, int64_t n, int D, int m, int NL,
                          int64_t* NN, int ldnn, unsigned char* flags, cudaStream_t st, int* done) {
    *done = 0;
    const int K8 = tc5::k8_of(D);
    int stages = tc5::kMaxStages;
    size_t smem = NL == 1 ? tc5::smem_bytes<1>(K8, stages) : tc5::smem_bytes<2>(K8, stages);
    if (smem > 227 * 1024) {
        stages = 2;
        smem = NL == 1 ? tc5::smem_bytes<1>(K8, stages) : tc5::smem_bytes<2>(K8, stages);
    }
    if (K8 > tc5::kMaxK8 || smem > 227 * 1024) return DGPB_OK;
    const int64_t tiles = cdiv(n, tc5::kN);
    void* pp;
    const size_t pack_bytes = (size_t)tiles * 2 * K8 * tc5::kN * 16;
    DGPB_TRY(ws->reserve(SLOT_KNN_PACK, pack_bytes + 256, &pp));
    float* packed = (float*)pp;
    float* xnmax = reinterpret_cast<float*>((char*)pp + pack_bytes);
    DGPB_CUDA_TRY(cudaMemsetAsync(xnmax, 0, sizeof(float), st));
    tc5::knn_tc5_pack_kernel<<<(unsigned)tiles, tc5::kN, 0, st>>>(x, n, D, K8, packed, xnmax);
    DGPB_LAUNCHED();
    const unsigned grid = (unsigned)cdiv(M, tc5::kQ);
#define KNN_TC5_CASE(NLV)                                                                                               \
    if (NL == NLV) {                                                                                                    \
        DGPB_CUDA_TRY(cudaFuncSetAttribute(tc5::knn_tc5_kernel<NLV, ORDERED>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)smem));                                                                 \
        tc5::knn_tc5_kernel<NLV, ORDERED><<<grid, tc5::kThreads, smem, st>>>(q, M, x, n, D, K8, stages, m, packed, xnmax, NN, ldnn, \
                                                                             flags);                                    \
        DGPB_LAUNCHED();                                                                                                \
    }
    KNN_TC5_CASE(1) KNN_TC5_CASE(2)
#undef KNN_TC5_CASE
    *done = 1;
    return DGPB_OK;
}


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
    if (g_knn_mma == 5) {
        int done = 0;
        DGPB_TRY(knn_tc5_search<ORDERED>(ws, q, M, x, n, D, m, NL, NN, ldnn, flags, st, &done));
        if (done) return launch_knn<ORDERED>(q, M, x, n, D, m, NN, ldnn, flags, st);
    }
    if (g_knn_mma == 3 || g_knn_mma == 5) {
        const int KS8 = (D + 7) / 8;
#define KNN_TF32_CASE(KSV, NLV)                                                                                          \
    if (KS8 == KSV && NL == NLV) {                                                                                       \
        static bool cfg = false;                                                                                         \
        if (!cfg) {                                                                                                      \
            DGPB_CUDA_TRY(cudaFuncSetAttribute(knn_tf32_kernel<KSV, NLV, ORDERED>,                                       \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,                             \
                                               (int)knn_tf32_smem_bytes<KSV, NLV>()));                                   \
            cfg = true;                                                                                                  \
        }                                                                                                                \
        knn_tf32_kernel<KSV, NLV, ORDERED><<<grid, 256, knn_tf32_smem_bytes<KSV, NLV>(), st>>>(q, M, x, n, D, m, NN,      \
                                                                                              ldnn, flags);              \
        DGPB_LAUNCHED();                                                                                                 \
        return launch_knn<ORDERED>(q, M, x, n, D, m, NN, ldnn, flags, st);                                               \
    }
        KNN_TF32_CASE(1, 1) KNN_TF32_CASE(2, 1) KNN_TF32_CASE(3, 1) KNN_TF32_CASE(4, 1)
        KNN_TF32_CASE(1, 2) KNN_TF32_CASE(2, 2) KNN_TF32_CASE(3, 2) KNN_TF32_CASE(4, 2)
#undef KNN_TF32_CASE
    }
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
    g_knn_mma = on;   // 0 = scalar kernel, 1 = DMMA screen + rank, 2 = probe: DMMA screen without list maintenance
                      // (invalid results), 3 = split-TF32 screen + rank
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
    DGPB_NVTX("dgpb:knn_ordered");
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
    DGPB_NVTX("dgpb:knn");
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

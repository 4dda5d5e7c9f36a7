// dense.cu -- dense FP64 GP-node linear algebra on sm_100a:
//   * tiled kernel-matrix construction (sexp / Matern-2.5, shared or per-dimension length-scales),
//   * blocked right-looking Cholesky whose panel TRSM and trailing SYRK run on the FP64 tensor path
//     (mma.sync.m8n8k4.f64 -> DMMA), batched over GP nodes with blockIdx.z,
//   * the "sliding window" partial factorisation of [[K],[y'],[I]] that yields log|K|, y'K^-1y,
//     K^-1y and K^-1 in one pass of identical panel/update kernels (n^3 flop total),
//   * the fused gradient contraction sum_ij (K^-1 - aa'/s2)_ij dK_ij/dlog(theta_p) with dK recomputed
//     on the fly.
// Reference semantics: dgpsi/kernel_class.py:304-359 (k_matrix), :403-449 (llik), :481-492
// (log_likelihood_func), :735-748 (compute_stats); dgpsi/functions.py:16-121.
#include "dense.cuh"
#include "vecchia.cuh"

#include <algorithm>
#include <mutex>

namespace dgpb {

// ------------------------------------------------------------------------------------------------
// small PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = pred ? 16 : 0;  // src-size 0 -> 16 bytes of zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ------------------------------------------------------------------------------------------------
// 1. kernel-matrix construction
// ------------------------------------------------------------------------------------------------
// One CTA builds a 64x64 tile (lower-triangular tile pairs only).  Scaled coordinates of the 64 row
// points and 64 column points are staged in shared memory (coalesced loads from the variable-major
// sources); each thread then produces a 4 x 4 register tile -- rows 4ty..4ty+3, columns tx + 16c -- so one
// dimension costs 6 shared-memory loads per 16 entries (the one-entry-per-thread form was LDS-bound at
// 2 loads per entry per dimension) and consecutive threads still store consecutive columns (128-byte lines).
// HBM-write bound: 8 n^2 / 2 bytes (lower) or 8 n^2 (MIRROR).  Arithmetic per entry is unchanged: squared
// differences accumulated with separate multiply and add in ascending d (what scipy's pdist does).
// KIND is a template parameter so the squared-exponential instance does not carry the Matern accumulators: 64
// registers instead of 96, four resident CTAs per SM instead of two -- the kernel is bound by FP64 issue (three
// non-fused operations per dimension and entry, kept for bit-parity with pdist) plus the prologue / store phases
// that only other resident CTAs can cover.
template <bool MIRROR, int KIND>
__global__ void __launch_bounds__(256, KIND == DGPB_SEXP ? 4 : 2) kbuild_kernel(KernelDev kd, double* __restrict__ T, int64_t ld, int n, int npad,
                                                     const double* __restrict__ wdiag) {
    __shared__ __align__(16) double xi[kMaxDim][64];
    __shared__ double xj[kMaxDim][64];
    const int tid = threadIdx.x;
    const int t = blockIdx.x;
    int ti, tj;
    tri_index(t, ti, tj);
    const int D = kd.D;
    for (int idx = tid; idx < D * 64; idx += 256) {
        int d = idx >> 6, l = idx & 63;
        int gi = ti * 64 + l, gj = tj * 64 + l;
        xi[d][l] = gi < n ? kd.x(d, gi) : 0.0;
        xj[d][l] = gj < n ? kd.x(d, gj) : 0.0;
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
    double acc[4][4];   // sexp: squared distance; matern: running product of the polynomial factors
    double sr[4][4];    // matern: sum of r
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            acc[r][c] = KIND == DGPB_SEXP ? 0.0 : 1.0;
            sr[r][c] = 0.0;
        }
    for (int d = 0; d < D; ++d) {
        const double2 a01 = *reinterpret_cast<const double2*>(&xi[d][4 * ty]);
        const double2 a23 = *reinterpret_cast<const double2*>(&xi[d][4 * ty + 2]);
        const double av[4] = {a01.x, a01.y, a23.x, a23.y};
        double bv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) bv[c] = xj[d][tx + 16 * c];
        if (KIND == DGPB_SEXP) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double df = av[r] - bv[c];
                    acc[r][c] = __dadd_rn(acc[r][c], __dmul_rn(df, df));
                }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double rr = fabs(av[r] - bv[c]);
                    acc[r][c] *= 1.0 + kSqrt5 * rr + (5.0 / 3.0) * (rr * rr);
                    sr[r][c] += rr;
                }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int gi = ti * 64 + 4 * ty + r;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int gj = tj * 64 + tx + 16 * c;
            if (gj > gi) continue;
            double v;
            if (gi == gj) {
                v = gi < n ? 1.0 + kd.nugget * (wdiag ? wdiag[gi] : 1.0) : 1.0;
            } else if (gi >= n) {
                v = 0.0;
            } else {
                v = KIND == DGPB_SEXP ? exp_nonpos(-acc[r][c]) : acc[r][c] * exp_nonpos(-kSqrt5 * sr[r][c]);
            }
            if (MIRROR) {
                if (gi < n) {
                    T[(int64_t)gi * ld + gj] = v;
                    T[(int64_t)gj * ld + gi] = v;
                }
            } else {
                T[(int64_t)gi * ld + gj] = v;
            }
        }
    }
}

// dK/dlog(theta_p) slices for the API-parity entry point dgpb_kmatrix (kernel_class.py:328-351).
// One thread per (i,j); not on the training hot path (the gradient kernel never materialises these).
__global__ void dk_kernel(KernelDev kd, double* __restrict__ dK, int n, int P, int nugget_est,
                          const double* __restrict__ wdiag) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * n) return;
    int i = (int)(idx / n), j = (int)(idx % n);
    const int D = kd.D;
    const int nl = kd.ard ? D : 1;
    const int64_t nn = (int64_t)n * n;
    if (i == j) {
        for (int p = 0; p < nl; ++p) dK[p * nn + idx] = 0.0;
        if (nugget_est) dK[nl * nn + idx] = kd.nugget * (wdiag ? wdiag[i] : 1.0);
        return;
    }
    if (nugget_est) dK[nl * nn + idx] = 0.0;
    if (kd.kind == DGPB_SEXP) {
        double dist = 0.0;
        for (int d = 0; d < D; ++d) {
            double df = kd.x(d, i) - kd.x(d, j);
            dist = __dadd_rn(dist, __dmul_rn(df, df));
        }
        double K = exp(-dist);
        if (kd.ard) {
            for (int d = 0; d < D; ++d) {
                double df = kd.x(d, i) - kd.x(d, j);
                dK[d * nn + idx] = 2.0 * (df * df) * K;
            }
        } else {
            dK[idx] = (2.0 * dist) * K;
        }
    } else {
        double coef = 1.0, s = 0.0, csum = 0.0;
        for (int d = 0; d < D; ++d) {
            double r = fabs(kd.x(d, i) - kd.x(d, j));
            double poly = 1.0 + kSqrt5 * r + (5.0 / 3.0) * (r * r);
            coef *= poly;
            s += r;
            csum += (5.0 / 3.0) * (r * r) * (1.0 + kSqrt5 * r) / poly;
        }
        double K = coef * exp(-kSqrt5 * s);
        if (kd.ard) {
            for (int d = 0; d < D; ++d) {
                double r = fabs(kd.x(d, i) - kd.x(d, j));
                double poly = 1.0 + kSqrt5 * r + (5.0 / 3.0) * (r * r);
                dK[d * nn + idx] = ((5.0 / 3.0) * (r * r) * (1.0 + kSqrt5 * r) / poly) * K;
            }
        } else {
            dK[idx] = csum * K;
        }
    }
}

// y row (row npad) and, for the augmented layout, the unit entries of the identity rows.
__global__ void rows_init_kernel(double* __restrict__ T, int64_t ld, int n, int npad, const double* __restrict__ y,
                                 int aug) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < ld) T[(int64_t)npad * ld + j] = (j < n && y) ? y[j] : 0.0;
    if (aug && j < npad) T[(int64_t)(npad + 1 + j) * ld + j] = 1.0;
}

// ------------------------------------------------------------------------------------------------
// 2. blocked Cholesky: panel kernel (POTF2 + triangular inverse + TRSM-as-GEMM on DMMA)
// ------------------------------------------------------------------------------------------------
// Every CTA of the launch factors the same NB x NB diagonal block redundantly in shared memory (it is on
// the critical path either way, and this removes one dependent launch per panel), inverts it, and then
// applies X <- X * inv(L_kk)' to its own 128 rows of the panel with FP64 mma.  CTA 0 publishes L_kk and
// diag(L_kk) to the side buffer (not into T: sibling CTAs are still reading the unfactored block).
}  // namespace dgpb
#include "potf2.cuh"
namespace dgpb {

// (b) panel rows: X <- X * inv(L_kk)'  on the FP64 tensor path, 128-row tiles, grid-stride over tiles.
//     105 KB of shared memory and < 128 registers so a CTA co-resides with a trailing-update CTA.
__global__ void __launch_bounds__(256, 2) trsm_kernel(Batch bt, int64_t ld, int npad, int k0, int row_lo, int row_hi,
                                                      int early) {
    extern __shared__ double smem[];
    double* sD = smem;             // 64 x LDS : inv(L_kk)
    double* sA = sD + 64 * LDS;    // 128 x LDS: panel rows being solved
    const int tid = threadIdx.x;
    double* __restrict__ T = bt.T[blockIdx.z];
    const double* __restrict__ dinv = bt.diag[blockIdx.z] + (size_t)npad * (1 + NB);
    const int ntiles = (row_hi - row_lo + TM - 1) / TM;
    pdl_wait();
    if (early) pdl_trigger();
    auto load_tile = [&](int tile) {
        const int r0 = row_lo + tile * TM;
        for (int c = tid; c < TM * 32; c += 256) {
            int lr = c >> 5, ch = c & 31;
            int gr = r0 + lr;
            bool ok = gr < row_hi;
            cp_async16(&sA[lr * LDS + ch * 2], T + (int64_t)(ok ? gr : k0) * ld + k0 + ch * 2, ok);
        }
        cp_async_commit();
    };
    for (int c = tid; c < 64 * 32; c += 256) {
        int lr = c >> 5, ch = c & 31;
        cp_async16(&sD[lr * LDS + ch * 2], dinv + lr * 64 + ch * 2, true);
    }
    load_tile(blockIdx.x);
    const int w = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int r0 = row_lo + tile * TM;
        cp_async_wait<0>();
        __syncthreads();
        if (tile + (int)gridDim.x >= ntiles) pdl_trigger();   // this CTA's last tile
        double acc[2][8][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int nj = 0; nj < 8; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;
#pragma unroll 4
        for (int kk = 0; kk < 64; kk += 4) {
            double a[2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) a[mi] = sA[(16 * w + 8 * mi + g) * LDS + kk + t4];
#pragma unroll
            for (int nj = 0; nj < 8; ++nj) {
                double b = sD[(8 * nj + g) * LDS + kk + t4];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) dmma884(acc[mi][nj][0], acc[mi][nj][1], a[mi], b);
            }
        }
        __syncthreads();  // every warp is done reading sA before the next tile overwrites it
        if (tile + (int)gridDim.x < ntiles) load_tile(tile + gridDim.x);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
            int gr = r0 + 16 * w + 8 * mi + g;
            if (gr < row_hi) {
#pragma unroll
                for (int nj = 0; nj < 8; ++nj)
                    *reinterpret_cast<double2*>(&T[(int64_t)gr * ld + k0 + 8 * nj + 2 * t4]) =
                        make_double2(acc[mi][nj][0], acc[mi][nj][1]);
            }
        }
    }
}
constexpr size_t kTrsmSmem = (size_t)(64 * LDS + TM * LDS) * sizeof(double);

// ------------------------------------------------------------------------------------------------
// 3. trailing update  C[r,c] -= P_r P_c'  over the lower triangle of rows/cols [lo, row_hi) restricted to
//    columns [lo, col_hi), FP64 mma, K = NB.
// ------------------------------------------------------------------------------------------------
// 128 (rows) x 64 (cols) tile per CTA, 8 warps as 4 x 2, warp tile 32x32 = 4x4 m8n8 accumulators (64
// registers), so TWO CTAs fit per SM (<= 128 registers, 2 x 108 KB shared memory): while one CTA waits for
// its C tile / panel rows or drains its stores, the other keeps the DMMA pipe busy.  The accumulators are
// INITIALISED with the C tile (global loads issued before the panel data is needed) and the A fragments
// are negated, so D = (-A)B' + C is a single DMMA chain per k-step.  `narrow` = only the first 64-column
// block (the next panel's columns: the look-ahead part of the update).
constexpr int UBN = 64;
constexpr int ULD = 36;  // 32-wide K chunks, 36 % 16 == 4
constexpr size_t kUpdateSmem = (size_t)(2 * (TM + UBN) * ULD) * sizeof(double);

// `flags` (probe only, 0 in production): 1 = skip C loads, 2 = skip C stores, 4 = skip the DMMA loop body,
// 8 = skip the panel loads.  16 (production): trigger the dependent launch at once instead of near the end.
__global__ void __launch_bounds__(256, 2) update_kernel(Batch bt, int64_t ld, int k0, int kc, int lo, int row_hi,
                                                        int col_hi, int ncol_tiles, int flags) {
    extern __shared__ double smem[];
    int ti, tj;
    if (ncol_tiles > 0) {  // look-ahead part: the first `ncol_tiles` 64-column blocks of every row tile
        ti = blockIdx.x / ncol_tiles;
        tj = blockIdx.x - ti * ncol_tiles;
    } else {               // row tile ti owns column tiles 0 .. 2 ti + 1
        stair_index(blockIdx.x, ti, tj);
    }
    const int ra = lo + ti * TM, rb = lo + tj * UBN;
    if (rb >= col_hi || ra >= row_hi || rb > ra + TM - 1) return;
    pdl_wait();
    if (flags & 16) pdl_trigger();
    const int tid = threadIdx.x;
    double* __restrict__ T = bt.T[blockIdx.z];
    // per-thread cp.async slots: 8 (A) + 4 (B) 16-byte chunks per 32-wide K chunk
    const int lr0 = tid >> 4, colo = (tid & 15) * 2;
    const int64_t stride16 = 16 * ld;
    const double* gA = T + (int64_t)(ra + lr0) * ld + k0 + colo;
    const double* gB = T + (int64_t)(rb + lr0) * ld + k0 + colo;
    unsigned okA = 0, okB = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) okA |= (unsigned)(ra + lr0 + 16 * r < row_hi) << r;
#pragma unroll
    for (int r = 0; r < 4; ++r) okB |= (unsigned)(rb + lr0 + 16 * r < col_hi) << r;
    const int soff = lr0 * ULD + colo;
    auto load_chunk = [&](int c) {
        double* a = smem + (size_t)(c & 1) * (TM + UBN) * ULD + soff;
        double* b = a + TM * ULD;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            bool ok = (okA >> r) & 1u;
            cp_async16(a + r * 16 * ULD, ok ? gA + r * stride16 + c * 32 : T, ok);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            bool ok = (okB >> r) & 1u;
            cp_async16(b + r * 16 * ULD, ok ? gB + r * stride16 + c * 32 : T, ok);
        }
        cp_async_commit();
    };
    if (!(flags & 8)) {
        load_chunk(0);
        load_chunk(1);
    }

    // L2 prefetch of the C tile that the CTA taking over this SM slot will need (CTAs are dispatched in
    // linear order, 2 per SM): its accumulator loads then hit L2 instead of queueing on HBM behind the
    // store bursts of the current wave.
    {
        const int lin2 = (int)(blockIdx.z * gridDim.x + blockIdx.x) + 296;
        const int z2 = lin2 / (int)gridDim.x, x2 = lin2 - z2 * (int)gridDim.x;
        if (z2 < (int)gridDim.z) {
            int ti2, tj2;
            if (ncol_tiles > 0) {
                ti2 = x2 / ncol_tiles;
                tj2 = x2 - ti2 * ncol_tiles;
            } else {
                stair_index(x2, ti2, tj2);
            }
            const int ra2 = lo + ti2 * TM, rb2 = lo + tj2 * UBN;
            const double* T2 = bt.T[z2];
#pragma unroll
            for (int l = tid; l < TM * 4; l += 256) {
                const int r = ra2 + (l >> 2), cseg = rb2 + (l & 3) * 16;
                if (r < row_hi && cseg < col_hi)
                    asm volatile("prefetch.global.L2 [%0];\n" ::"l"(T2 + (int64_t)r * ld + cseg));
            }
        }
    }

    const int w = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int wm = w >> 1, wn = w & 1;
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        int gr = ra + 32 * wm + 8 * mi + g;
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
            int gc = rb + 32 * wn + 8 * nj + 2 * t4;
            if (gr < row_hi && gc < col_hi && !(flags & 1)) {
                double2 v = *reinterpret_cast<const double2*>(&T[(int64_t)gr * ld + gc]);
                acc[mi][nj][0] = v.x;
                acc[mi][nj][1] = v.y;
            } else {
                acc[mi][nj][0] = acc[mi][nj][1] = 0.0;
            }
        }
    }

    for (int c = 0; c < kc; ++c) {
        if (c + 1 < kc) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        const double* a_s = smem + (size_t)(c & 1) * (TM + UBN) * ULD;
        const double* b_s = a_s + TM * ULD;
        if (c + 2 >= kc) pdl_trigger();
#pragma unroll 2
        for (int kk = 0; kk < ((flags & 4) ? 0 : 32); kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = -a_s[(32 * wm + 8 * mi + g) * ULD + kk + t4];
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) b[nj] = b_s[(32 * wn + 8 * nj + g) * ULD + kk + t4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) dmma884(acc[mi][nj][0], acc[mi][nj][1], a[mi], b[nj]);
        }
        if (c + 2 < kc && !(flags & 8)) {
            __syncthreads();  // all warps finished with this stage before it is refilled
            load_chunk(c + 2);
        }
    }

#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        int gr = ra + 32 * wm + 8 * mi + g;
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
            int gc = rb + 32 * wn + 8 * nj + 2 * t4;
            if (gr < row_hi && gc < col_hi && !(flags & 2))
                *reinterpret_cast<double2*>(&T[(int64_t)gr * ld + gc]) = make_double2(acc[mi][nj][0], acc[mi][nj][1]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 4. reductions / extraction
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reduce_kernel(Batch bt, int64_t ld, int n, int npad, ScaleArgs sa,
                                                     double* __restrict__ out) {
    __shared__ double sred[8];
    const int b = blockIdx.x, tid = threadIdx.x;
    const double* dg = bt.diag[b];
    const double* yrow = bt.T[b] + (int64_t)npad * ld;
    double s1 = 0.0, s2 = 0.0;
    for (int i = tid; i < n; i += 256) s1 += log(fabs(dg[i]));
    for (int i = tid; i < npad; i += 256) s2 += yrow[i] * yrow[i];
    s1 = block_sum<256>(s1, sred);
    s2 = block_sum<256>(s2, sred);
    if (tid == 0) {
        out[b * 4 + 0] = 2.0 * s1;
        out[b * 4 + 1] = s2;
        out[b * 4 + 2] = sa.est[b] ? s2 / (double)n : sa.scale[b];
        out[b * 4 + 3] = (double)bt.info[b];   // 0 or failing column + 1: travels with the results (ESS wave gather)
    }
}

__global__ void restore_diag_kernel(Batch bt, int64_t ld, int npad) {
    const int b = blockIdx.z, kb = blockIdx.x;
    double* T = bt.T[b];
    const double* blk = bt.diag[b] + npad + (size_t)kb * NB * NB;
    for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x) {
        int i = idx >> 6, j = idx & 63;
        if (j <= i) T[(int64_t)(kb * NB + i) * ld + kb * NB + j] = blk[idx];
    }
}

// K^-1 (full symmetric n x n, ld = n) and K^-1 y out of the Schur complement of the augmented layout
__global__ void extract_inverse_kernel(const double* __restrict__ T, int64_t ld, int n, int npad,
                                       double* __restrict__ Rinv, double* __restrict__ Rinv_y) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * n) return;
    int i = (int)(idx / n), j = (int)(idx % n);
    int hi = i > j ? i : j, lo = i > j ? j : i;
    Rinv[idx] = -T[(int64_t)(npad + 1 + hi) * ld + npad + 1 + lo];
    if (j == 0 && Rinv_y) Rinv_y[i] = -T[(int64_t)(npad + 1 + i) * ld + npad];
}

// nu = sqrt(scale) * L z   (fmvn, functions.py:113-121, z injected); one warp per row, fixed order
__global__ void trmv_kernel(const double* __restrict__ T, int64_t ld, int n, double sqrt_scale,
                            const double* __restrict__ z, double* __restrict__ nu) {
    int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    const double* Lr = T + (int64_t)row * ld;
    double s = 0.0;
    for (int j = lane; j <= row; j += 32) s += Lr[j] * z[j];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) nu[row] = sqrt_scale * s;
}

// ------------------------------------------------------------------------------------------------
// 5. fused gradient contraction (kernel_class.py:418-427,435 without the P dense cho_solves)
//    S_p = sum_ij W_ij dK_p,ij,  W = K^-1 - a a'/sigma2;  partials[tile][p] summed in fixed order later.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_kernel(KernelDev kd, const double* __restrict__ T, int64_t ld, int n,
                                                   int npad, int P, int nugget_est, const double* __restrict__ out4,
                                                   double* __restrict__ partials) {
    __shared__ double xi[kMaxDim][64];
    __shared__ double xj[kMaxDim][64];
    __shared__ double ai[64], aj[64];
    __shared__ double sred[8];
    const int tid = threadIdx.x;
    const int t = blockIdx.x;
    int ti, tj;
    tri_index(t, ti, tj);
    const int D = kd.D;
    const double sigma2 = out4[2];
    for (int idx = tid; idx < D * 64; idx += 256) {
        int d = idx >> 6, l = idx & 63;
        int gi = ti * 64 + l, gj = tj * 64 + l;
        xi[d][l] = gi < n ? kd.x(d, gi) : 0.0;
        xj[d][l] = gj < n ? kd.x(d, gj) : 0.0;
    }
    if (tid < 64) {
        int gi = ti * 64 + tid, gj = tj * 64 + tid;
        ai[tid] = gi < n ? -T[(int64_t)(npad + 1 + gi) * ld + npad] : 0.0;
        aj[tid] = gj < n ? -T[(int64_t)(npad + 1 + gj) * ld + npad] : 0.0;
    }
    __syncthreads();
    // per-entry weights W_ij (x2 for strict lower, 0 for masked) and kernel values, kept in registers
    double Wv[16], Kv[16];
    double diagW = 0.0;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        int idx = tid + 256 * e;
        int li = idx >> 6, lj = idx & 63;
        int gi = ti * 64 + li, gj = tj * 64 + lj;
        Wv[e] = 0.0;
        Kv[e] = 0.0;
        if (gi < n && gj <= gi) {
            double kinv = -T[(int64_t)(npad + 1 + gi) * ld + npad + 1 + gj];
            double wgt = kinv - ai[li] * aj[lj] / sigma2;
            if (gi == gj) {
                diagW += wgt;
            } else {
                Wv[e] = 2.0 * wgt;
                if (kd.kind == DGPB_SEXP) {
                    double dist = 0.0;
                    for (int d = 0; d < D; ++d) {
                        double df = xi[d][li] - xj[d][lj];
                        dist = __dadd_rn(dist, __dmul_rn(df, df));
                    }
                    Kv[e] = exp(-dist);
                } else {
                    double coef = 1.0, s = 0.0;
                    for (int d = 0; d < D; ++d) {
                        double r = fabs(xi[d][li] - xj[d][lj]);
                        coef *= 1.0 + kSqrt5 * r + (5.0 / 3.0) * (r * r);
                        s += r;
                    }
                    Kv[e] = coef * exp(-kSqrt5 * s);
                }
            }
        }
    }
    const int nl = kd.ard ? D : 1;
    for (int p = 0; p < nl; ++p) {
        double acc = 0.0;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            int idx = tid + 256 * e;
            int li = idx >> 6, lj = idx & 63;
            double c = 0.0;
            const int d0 = kd.ard ? p : 0, d1 = kd.ard ? p + 1 : D;
            for (int d = d0; d < d1; ++d) {
                double df = xi[d][li] - xj[d][lj];
                if (kd.kind == DGPB_SEXP) {
                    c += 2.0 * (df * df);
                } else {
                    double r = fabs(df);
                    c += (5.0 / 3.0) * (r * r) * (1.0 + kSqrt5 * r) / (1.0 + kSqrt5 * r + (5.0 / 3.0) * (r * r));
                }
            }
            acc += Wv[e] * (c * Kv[e]);
        }
        acc = block_sum<256>(acc, sred);
        if (tid == 0) partials[(int64_t)t * P + p] = acc;
    }
    if (nugget_est) {
        double acc = block_sum<256>(diagW, sred);
        if (tid == 0) partials[(int64_t)t * P + nl] = kd.nugget * acc;
    }
}

// out_final = [nllik, sigma2, grad[0..P)]
__global__ void __launch_bounds__(256) grad_finish_kernel(const double* __restrict__ partials, int ntiles, int P, int n,
                                                          int scale_est, const double* __restrict__ out4,
                                                          double* __restrict__ out_final) {
    __shared__ double sred[8];
    const int p = blockIdx.x, tid = threadIdx.x;
    double s = 0.0;
    for (int t = tid; t < ntiles; t += 256) s += partials[(int64_t)t * P + p];
    s = block_sum<256>(s, sred);
    if (tid == 0) {
        out_final[2 + p] = 0.5 * s;
        if (p == 0) {
            double logdet = out4[0], quad = out4[1], sigma2 = out4[2];
            out_final[0] = scale_est ? 0.5 * (logdet + (double)n * log(sigma2)) : 0.5 * (logdet + quad / sigma2);
            out_final[1] = sigma2;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host-side orchestration
// ------------------------------------------------------------------------------------------------
static int configure_once() {
    static bool done = false;
    if (done) return DGPB_OK;
    DGPB_CUDA_TRY(cudaFuncSetAttribute(trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrsmSmem));
    DGPB_CUDA_TRY(cudaFuncSetAttribute(potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPotf2Smem));
    DGPB_CUDA_TRY(cudaFuncSetAttribute(update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpdateSmem));
    done = true;
    return DGPB_OK;
}

int launch_trmv(const double* T, int64_t ld, int n, double sqrt_scale, const double* z, double* nu, cudaStream_t st) {
    trmv_kernel<<<(unsigned)cdiv((int64_t)n * 32, 256), 256, 0, st>>>(T, ld, n, sqrt_scale, z, nu);
    DGPB_LAUNCHED();
    return DGPB_OK;
}

int setup_batch_slot(Workspace* ws, int tslot, const Geom& g, int B, Batch* bt, double** out_dev) {
    DGPB_REQUIRE(B >= 1 && B <= MAXB, "batch size out of range");
    void *pT, *pD, *pO, *pI;
    DGPB_TRY(ws->reserve(tslot ? SLOT_T2 : SLOT_T, g.elems() * sizeof(double) * B, &pT));
    DGPB_TRY(ws->reserve(tslot ? SLOT_DIAG2 : SLOT_DIAG, diag_elems(g) * sizeof(double) * B, &pD));
    DGPB_TRY(ws->reserve(tslot ? SLOT_OUT2 : SLOT_OUT, sizeof(double) * kOutDoubles, &pO));
    DGPB_TRY(ws->reserve(tslot ? SLOT_INFO2 : SLOT_INFO, sizeof(int) * MAXB, &pI));
    for (int b = 0; b < MAXB; ++b) {
        bt->T[b] = b < B ? (double*)pT + g.elems() * b : nullptr;
        bt->diag[b] = b < B ? (double*)pD + diag_elems(g) * b : nullptr;
    }
    bt->info = (int*)pI;
    *out_dev = (double*)pO;
    return DGPB_OK;
}

int setup_batch(Workspace* ws, const Geom& g, int B, Batch* bt, double** out_dev) {
    return setup_batch_slot(ws, 0, g, B, bt, out_dev);
}

// Grow both T sets and the side buffers to `B` matrices NOW, so that no later setup_batch_slot call of the same
// sequence reallocates a buffer that kernels in flight still use.
int reserve_batches(Workspace* ws, const Geom& g, int B) {
    Batch bt;
    double* out;
    DGPB_TRY(setup_batch_slot(ws, 0, g, B, &bt, &out));
    DGPB_TRY(setup_batch_slot(ws, 1, g, B, &bt, &out));
    return DGPB_OK;
}

int assemble_matrices(const Geom& g, const KernelDev* kds, const double* const* ys, const Batch& bt, int B,
                      cudaStream_t st) {
    const int nt = g.npad / 64;
    for (int b = 0; b < B; ++b) {
        if (g.aug) {
            // identity rows and the Schur-complement block start at zero
            DGPB_CUDA_TRY(cudaMemsetAsync(bt.T[b] + (size_t)(g.npad + 1) * g.ld, 0,
                                          (size_t)(g.R - g.npad - 1) * g.ld * sizeof(double), st));
        }
        if (kds[b].kind == DGPB_SEXP)
            kbuild_kernel<false, DGPB_SEXP><<<nt * (nt + 1) / 2, 256, 0, st>>>(kds[b], bt.T[b], g.ld, g.n, g.npad, nullptr);
        else
            kbuild_kernel<false, DGPB_MATERN25><<<nt * (nt + 1) / 2, 256, 0, st>>>(kds[b], bt.T[b], g.ld, g.n, g.npad, nullptr);
        DGPB_LAUNCHED();
        rows_init_kernel<<<(unsigned)cdiv(g.ld, 256), 256, 0, st>>>(bt.T[b], g.ld, g.n, g.npad, ys ? ys[b] : nullptr,
                                                                   g.aug ? 1 : 0);
        DGPB_LAUNCHED();
    }
    return DGPB_OK;
}

int assemble(const Geom& g, const KernelDev* kds, const double* const* ys, const Batch& bt, int B, cudaStream_t st) {
    DGPB_TRY(assemble_matrices(g, kds, ys, bt, B, st));
    DGPB_CUDA_TRY(cudaMemsetAsync(bt.info, 0, sizeof(int) * MAXB, st));
    return DGPB_OK;
}

// factorise + reduce a batch whose matrices are already assembled (info flags cleared here)
int factor_reduce(const Geom& g, const Batch& bt, int B, const ScaleArgs& sa, double* out, cudaStream_t st, int ctx) {
    DGPB_CUDA_TRY(cudaMemsetAsync(bt.info, 0, sizeof(int) * MAXB, st));
    DGPB_TRY(factorize(g, bt, B, st, ctx));
    DGPB_TRY(reduce_logdet_quad(g, bt, B, sa, out, st));
    return DGPB_OK;
}

// ---- optional launch profiler (bench.py roofline): CUDA-event pairs around every trailing-update launch ----
struct UpdateProfiler {
    bool on = false;
    std::vector<cudaEvent_t> ev;   // pairs
    int used = 0;                  // events in flight
    double ms = 0.0, launches = 0.0, flops = 0.0;
    double total_flops = 0.0;      // n^3/3 (n^3 augmented) per matrix of every factorisation while profiling
    double wasted_flops = 0.0;     // ... of which: speculative ESS candidates behind the accepted one
    cudaStream_t last = nullptr;
    int drain() {
        if (used == 0) return DGPB_OK;
        for (int i = 0; i < used; i += 2) {
            float t = 0.f;
            DGPB_CUDA_TRY(cudaEventSynchronize(ev[i + 1]));  // pairs may sit on different streams (threaded M-step)
            DGPB_CUDA_TRY(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
            ms += t;
            launches += 1.0;
        }
        used = 0;
        return DGPB_OK;
    }
};
static UpdateProfiler g_prof;
static std::mutex g_prof_mutex;  // the M-step issues factorisations from several host threads

// ESS waves evaluate candidate angles ahead of the decision; `matrices` of them (n x n, plain layout) turned out to
// lie behind the accepted one.  They were issued (total_flops) but the reference's schedule never needed them.
void profile_wasted(int matrices, int64_t n) {
    if (!g_prof.on || matrices <= 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    const double nn = (double)n;
    g_prof.wasted_flops += (double)matrices * nn * nn * nn / 3.0;
}

// Two-stream look-ahead: the caller's stream carries the critical path (panel k, then the NARROW update of
// the next panel's 64 columns), a side stream carries the BULK update of everything to the right.  Panel k+1
// therefore overlaps bulk update k.  Dependencies:
//   bulk_k   needs panel_k (event) and bulk_{k-1} (side-stream order);
//   narrow_k needs panel_k (main-stream order) and bulk_{k-1} (event: both write columns [k1, k1+64)).
struct LookAhead {
    cudaStream_t side = nullptr;   // bulk updates, lowest priority
    cudaStream_t crit = nullptr;   // panels, narrow / inner / look-ahead updates: the critical path, highest priority
    std::vector<cudaEvent_t> ev;
    int init(size_t need) {
        if (!side) {
            // The caller's stream (torch's current stream) has priority 0, which is the LOWEST a stream can have, so
            // a "low-priority" side stream next to it is no different: with equal priorities a kernel launched
            // later is only dispatched once the earlier grid has no CTA left to hand out, and every panel would
            // wait behind the whole bulk grid.  The critical path therefore runs on its own highest-priority
            // stream (forked from / joined to the caller's stream): its CTAs take the next SM slot that frees up.
            int lo_pri = 0, hi_pri = 0;
            DGPB_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
            DGPB_CUDA_TRY(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, lo_pri));
            DGPB_CUDA_TRY(cudaStreamCreateWithPriority(&crit, cudaStreamNonBlocking, hi_pri));
        }
        while (ev.size() < need) {
            cudaEvent_t e;
            DGPB_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ev.push_back(e);
        }
        return DGPB_OK;
    }
};
static thread_local LookAhead g_las[2];

int join_waves(Workspace* ws, cudaStream_t st) {
    for (int c = 0; c < 2; ++c)
        if (ws->wave_done[c]) DGPB_CUDA_TRY(cudaStreamWaitEvent(st, ws->wave_done[c], 0));
    return DGPB_OK;
}

// Launch with the programmatic-stream-serialisation attribute (`pdl`): the grid may become resident before the
// preceding kernel of the stream has finished; every kernel launched this way starts with pdl_wait().
static int g_pdl = 1;
static int g_pdl_early_b = 4;   // trigger early while the batch has at most this many matrices
template <typename... KArgs, typename... Args>
static cudaError_t launch_chain(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && g_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static int launch_panel(const Geom& g, const Batch& bt, int B, int k0, int row_hi, cudaStream_t st) {
    DGPB_CUDA_TRY(launch_chain(true, potf2_kernel, dim3((unsigned)B), dim3(256), kPotf2Smem, st, bt, g.ld, g.npad, k0,
                               (long long*)nullptr, (int)(B <= g_pdl_early_b)));
    DGPB_LAUNCHED();
    const int rows = row_hi - (k0 + NB);
    if (rows > 0) {
        const int tiles = (int)cdiv(rows, TM);
        dim3 grid((unsigned)std::min(tiles, std::max(1, 296 / B)), 1, (unsigned)B);
        DGPB_CUDA_TRY(launch_chain(true, trsm_kernel, grid, dim3(256), kTrsmSmem, st, bt, g.ld, g.npad, k0, k0 + NB,
                                   row_hi, (int)(B <= g_pdl_early_b)));
        DGPB_LAUNCHED();
    }
    return DGPB_OK;
}

static int launch_update(const Batch& bt, int B, int64_t ld, int k0, int K, int lo, int row_hi, int col_hi,
                         int ncol_tiles, cudaStream_t st, bool pdl = false) {
    const int rows = row_hi - lo;
    if (rows <= 0 || col_hi <= lo) return DGPB_OK;
    const int ntr = (int)cdiv(rows, TM);
    const unsigned nblk = ncol_tiles > 0 ? (unsigned)(ntr * ncol_tiles) : (unsigned)(ntr * (ntr + 1));
    DGPB_CUDA_TRY(launch_chain(pdl, update_kernel, dim3(nblk, 1, (unsigned)B), dim3(256), kUpdateSmem, st, bt, ld, k0,
                               K / 32, lo, row_hi, col_hi, ncol_tiles, (pdl && B <= g_pdl_early_b) ? 16 : 0));
    DGPB_LAUNCHED();
    return DGPB_OK;
}

// Tunables (dgpb_tune): width of a hyper-block and the smallest remaining window for which one is used.
static int g_hb = 512;
static int g_hb_min_w = 2560;
static int g_hb_graded = 1;
static int g_hb_small_b = 3;    // batches of at most this many matrices use 256-column hyper-blocks: there the
                                // critical path bounds the factorisation, and a K = 512 look-ahead update on it costs
                                // more than the bulk's extra C traffic (n = 5000: B = 1 4.25 -> 3.93 ms, B = 2 5.43 ->
                                // 5.22 ms, augmented 6.82 -> 6.45 ms; B >= 4 is faster with 512)
static int g_crit_stream = 1;   // critical path of the factorisation on its own highest-priority stream

// Right-looking factorisation on two levels.
//   HYPER-BLOCK [h0, h1) of up to `g_hb` columns: factored by super-steps of two 64-column panels
//     panel A -> narrow update of panel B's columns (K = 64) -> panel B -> inner update (K = 128)
//   where the inner update only touches the hyper-block's own columns [lo2, h1) (all rows below), on the
//   caller's stream.  Everything to the right of the hyper-block is updated ONCE per hyper-block with
//   K = h1 - h0 (512): the look-ahead part (the next hyper-block's columns) on the caller's stream, the bulk on a
//   low-priority side stream where it overlaps the next hyper-block's panels and inner updates.
// K = 512 reads and writes a trailing tile once per 4 x 128 columns eliminated (64 flop per byte of C traffic
// instead of 16) and quarters the number of bulk launches.  When the remaining window is small the hyper-block
// degenerates to one super-step (h1 - h0 = 128): there the critical path, not the tensor pipe, bounds the
// step, and a longer inner phase would only lengthen it.
// Dependencies: bulk_h needs the panels of h (event) and bulk_{h-1} (side-stream order); look-ahead_h needs
// bulk_{h-1} (event: both write columns [h1, h1 + next width)).
int factorize(const Geom& g, const Batch& bt, int B, cudaStream_t caller, int ctx) {
    DGPB_NVTX("dgpb:factorize");
    DGPB_TRY(configure_once());
    LookAhead& g_la = g_las[ctx & 1];
    if (g_prof.on) {
        std::lock_guard<std::mutex> lock(g_prof_mutex);
        const double nn = (double)g.n;
        g_prof.total_flops += (double)B * nn * nn * nn * (g.aug ? 1.0 : 1.0 / 3.0);
    }
    DGPB_TRY(g_la.init(2 * (size_t)(g.npad / (2 * NB) + 2) + 6));
    cudaStream_t side = g_la.side;
    cudaStream_t st = g_crit_stream ? g_la.crit : caller;   // the stream of the critical path
    int evi = 0;
    DGPB_CUDA_TRY(cudaEventRecord(g_la.ev[evi], caller));
    DGPB_CUDA_TRY(cudaStreamWaitEvent(side, g_la.ev[evi], 0));
    if (st != caller) DGPB_CUDA_TRY(cudaStreamWaitEvent(st, g_la.ev[evi], 0));
    ++evi;
    auto width_at = [&](int h0) {
        if (h0 >= g.npad) return 0;
        const int window = g.aug ? g.npad + 1 + 2 * NB : g.R - h0;  // rows still active at this column
        int hb = (window >= g_hb_min_w) ? (B <= g_hb_small_b ? std::min(g_hb, 256) : g_hb) : 2 * NB;
        if (g_hb_graded) hb = std::min(hb, std::max(2 * NB, 2 * h0));  // 128, 256, 512: start the side stream early
        return std::min(hb, g.npad - h0);
    };
    cudaEvent_t prev_bulk = nullptr;
    int h0 = 0, hw = width_at(0);
    while (hw > 0) {
        const int h1 = h0 + hw;
        int rh = g.R;
        for (int kA = h0; kA < h1; kA += 2 * NB) {
            const int kB = kA + NB;
            const bool has_b = kB < h1;
            const int rhA = g.aug ? g.npad + 1 + kA + NB : g.R;
            DGPB_TRY(launch_panel(g, bt, B, kA, rhA, st));
            int K = NB, lo2 = kA + NB;
            rh = rhA;
            if (has_b) {
                DGPB_TRY(launch_update(bt, B, g.ld, kA, NB, kB, rhA, std::min(kB + NB, rhA), 1, st, true));
                const int rhB = g.aug ? g.npad + 1 + kB + NB : g.R;
                DGPB_TRY(launch_panel(g, bt, B, kB, rhB, st));
                K = 2 * NB;
                lo2 = kB + NB;
                rh = rhB;
            }
            if (lo2 < h1) DGPB_TRY(launch_update(bt, B, g.ld, kA, K, lo2, rh, h1, (h1 - lo2) / UBN, st, true));
        }
        const int hw_next = width_at(h1);
        if (rh - h1 > 0) {
            const int K = h1 - h0;
            cudaEvent_t ev_panel = g_la.ev[evi++];
            DGPB_CUDA_TRY(cudaEventRecord(ev_panel, st));
            // ---- bulk: columns [h1 + next width, rh) on the side stream
            const int blo = h1 + hw_next;
            const int brows = rh - blo;
            cudaEvent_t this_bulk = nullptr;
            if (brows > 0) {
                DGPB_CUDA_TRY(cudaStreamWaitEvent(side, ev_panel, 0));
                if (g_prof.on) {
                    std::lock_guard<std::mutex> lock(g_prof_mutex);
                    if (g_prof.used + 2 > (int)g_prof.ev.size()) DGPB_TRY(g_prof.drain());
                    const int slot = g_prof.used;
                    DGPB_CUDA_TRY(cudaEventRecord(g_prof.ev[slot], side));
                    DGPB_TRY(launch_update(bt, B, g.ld, h0, K, blo, rh, rh, 0, side));
                    DGPB_CUDA_TRY(cudaEventRecord(g_prof.ev[slot + 1], side));
                    g_prof.used += 2;
                    g_prof.flops += (double)B * 0.5 * (double)brows * (double)(brows + 1) * 2.0 * K;
                } else {
                    DGPB_TRY(launch_update(bt, B, g.ld, h0, K, blo, rh, rh, 0, side));
                }
                this_bulk = g_la.ev[evi++];
                DGPB_CUDA_TRY(cudaEventRecord(this_bulk, side));
            }
            // ---- look-ahead: the next hyper-block's columns [h1, h1 + next width) on the critical path
            if (hw_next > 0) {
                if (prev_bulk) DGPB_CUDA_TRY(cudaStreamWaitEvent(st, prev_bulk, 0));
                DGPB_TRY(launch_update(bt, B, g.ld, h0, K, h1, rh, std::min(h1 + hw_next, rh), hw_next / UBN, st,
                                       true));
            }
            prev_bulk = this_bulk;
        }
        h0 = h1;
        hw = hw_next;
    }
    DGPB_CUDA_TRY(cudaEventRecord(g_la.ev[evi], side));
    DGPB_CUDA_TRY(cudaStreamWaitEvent(caller, g_la.ev[evi], 0));
    ++evi;
    if (st != caller) {
        DGPB_CUDA_TRY(cudaEventRecord(g_la.ev[evi], st));
        DGPB_CUDA_TRY(cudaStreamWaitEvent(caller, g_la.ev[evi], 0));
    }
    return DGPB_OK;
}

int reduce_logdet_quad(const Geom& g, const Batch& bt, int B, const ScaleArgs& sa, double* out, cudaStream_t st) {
    reduce_kernel<<<B, 256, 0, st>>>(bt, g.ld, g.n, g.npad, sa, out);
    DGPB_LAUNCHED();
    return DGPB_OK;
}

int restore_diag_blocks(const Geom& g, const Batch& bt, int B, cudaStream_t st) {
    dim3 grid((unsigned)(g.npad / NB), 1, (unsigned)B);
    restore_diag_kernel<<<grid, 256, 0, st>>>(bt, g.ld, g.npad);
    DGPB_LAUNCHED();
    return DGPB_OK;
}

int loglik_batch_device(Workspace* ws, const KernelDev* kds, const double* const* ys, const ScaleArgs& sa, int B,
                        int64_t n, Batch* bt_out, Geom* g_out, double** out_dev, cudaStream_t st) {
    Geom g = make_geom(n, false);
    Batch bt;
    double* out;
    DGPB_TRY(setup_batch(ws, g, B, &bt, &out));
    DGPB_TRY(assemble(g, kds, ys, bt, B, st));
    DGPB_TRY(factorize(g, bt, B, st));
    DGPB_TRY(reduce_logdet_quad(g, bt, B, sa, out, st));
    if (bt_out) *bt_out = bt;
    if (g_out) *g_out = g;
    *out_dev = out;
    return DGPB_OK;
}

// copy `count` doubles + the info flags back and check positive-definiteness
static int fetch_results(Workspace* ws, const Batch& bt, const double* dev, int count, int B, double* host,
                         cudaStream_t st) {
    int* info_host = reinterpret_cast<int*>(ws->pinned + 2048);
    DGPB_CUDA_TRY(cudaMemcpyAsync(ws->pinned, dev, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
    DGPB_CUDA_TRY(cudaMemcpyAsync(info_host, bt.info, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    DGPB_CUDA_TRY(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b) {
        if (info_host[b] != 0) {
            set_error("matrix %d is not positive definite (pivot %d)", b, info_host[b]);
            return DGPB_NOT_PD;
        }
    }
    for (int i = 0; i < count; ++i) host[i] = ws->pinned[i];
    return DGPB_OK;
}

// host formula of kernel_class.py:482-488 from (logdet K, y'K^-1y): cov = scale*K
static inline double llik_from(double logdetK, double quadK, double scale, int64_t n) {
    return -0.5 * (logdetK + (double)n * log(scale) + quadK / scale);
}

// K_ii += shift_i on an assembled matrix (heteroskedastic noise on top of the node's own nugget)
__global__ void diag_shift_kernel(double* __restrict__ T, int64_t ld, int n, const double* __restrict__ shift) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) T[(size_t)i * ld + i] += shift[i];
}

int grad_pipeline(Workspace* ws, const dgpb_node* node, int64_t n, bool want_grad, double* Rinv, double* Rinv_y,
                  double* out_host, cudaStream_t st, const double* diag_shift = nullptr) {
    DGPB_NVTX("dgpb:grad_pipeline");
    DGPB_TRY(join_waves(ws, st));
    KernelDev kd;
    DGPB_TRY(make_kernel_dev(node, n, nullptr, &kd));
    Geom g = make_geom(n, true);
    Batch bt;
    double* out;
    DGPB_TRY(setup_batch(ws, g, 1, &bt, &out));
    const double* ys[1] = {node->output};
    DGPB_TRY(assemble(g, &kd, ys, bt, 1, st));
    if (diag_shift) {
        diag_shift_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(bt.T[0], g.ld, g.n, diag_shift);
        DGPB_LAUNCHED();
    }
    DGPB_TRY(factorize(g, bt, 1, st));
    ScaleArgs sa;
    sa.scale[0] = node->scale;
    sa.est[0] = node->scale_est;
    DGPB_TRY(reduce_logdet_quad(g, bt, 1, sa, out, st));
    if (Rinv) {
        extract_inverse_kernel<<<(unsigned)cdiv(n * n, 256), 256, 0, st>>>(bt.T[0], g.ld, g.n, g.npad, Rinv, Rinv_y);
        DGPB_LAUNCHED();
    }
    if (want_grad) {
        const int nl = node->nlen;
        const int P = nl + (node->nugget_est ? 1 : 0);
        const int nt = (int)cdiv(n, 64);
        const int ntiles = nt * (nt + 1) / 2;
        void* part;
        DGPB_TRY(ws->reserve(SLOT_PART, sizeof(double) * (size_t)ntiles * P, &part));
        grad_kernel<<<ntiles, 256, 0, st>>>(kd, bt.T[0], g.ld, g.n, g.npad, P, node->nugget_est, out, (double*)part);
        DGPB_LAUNCHED();
        double* fin = out + kOutGrad;
        grad_finish_kernel<<<P, 256, 0, st>>>((double*)part, ntiles, P, g.n, node->scale_est, out, fin);
        DGPB_LAUNCHED();
        DGPB_TRY(fetch_results(ws, bt, fin, P + 2, 1, out_host, st));
    } else {
        double tmp[4];
        DGPB_TRY(fetch_results(ws, bt, out, 4, 1, tmp, st));
    }
    return DGPB_OK;
}

}  // namespace dgpb

using namespace dgpb;

extern "C" {

int dgpb_kmatrix(const double* X, int64_t n, int64_t D, const double* length_host, int64_t nlen, double nugget,
                 const double* wdiag, int kind, int nugget_est, double* K, double* dK, void* stream) {
    DGPB_NVTX("dgpb:kmatrix");
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(n >= 1 && K != nullptr, "n < 1 or K is NULL");
    KernelDev kd;
    DGPB_TRY(make_kernel_dev_rowmajor(X, D, length_host, nlen, nugget, kind, &kd));
    const int nt = (int)cdiv(n, 64);
    if (kd.kind == DGPB_SEXP)
        kbuild_kernel<true, DGPB_SEXP><<<nt * (nt + 1) / 2, 256, 0, st>>>(kd, K, n, (int)n, nt * 64, wdiag);
    else
        kbuild_kernel<true, DGPB_MATERN25><<<nt * (nt + 1) / 2, 256, 0, st>>>(kd, K, n, (int)n, nt * 64, wdiag);
    DGPB_LAUNCHED();
    if (dK) {
        const int P = (int)nlen + (nugget_est ? 1 : 0);
        dk_kernel<<<(unsigned)cdiv(n * n, 256), 256, 0, st>>>(kd, dK, (int)n, P, nugget_est, wdiag);
        DGPB_LAUNCHED();
    }
    return DGPB_OK;
}

int dgpb_loglik_dense(dgpb_ws* ws, const dgpb_node* node, int64_t n, double* out_host, void* stream) {
    DGPB_NVTX("dgpb:loglik_dense");
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(ws && node && out_host && n >= 1, "NULL argument");
    DGPB_TRY(join_waves(ws, st));
    KernelDev kd;
    DGPB_TRY(make_kernel_dev(node, n, nullptr, &kd));
    const double* ys[1] = {node->output};
    ScaleArgs sa;
    sa.scale[0] = node->scale;
    sa.est[0] = 0;
    Batch bt;
    Geom g;
    double* out;
    DGPB_TRY(loglik_batch_device(ws, &kd, ys, sa, 1, n, &bt, &g, &out, st));
    double r[4];
    DGPB_TRY(fetch_results(ws, bt, out, 4, 1, r, st));
    out_host[0] = llik_from(r[0], r[1], node->scale, n);
    return DGPB_OK;
}

int dgpb_nllik_grad_dense(dgpb_ws* ws, const dgpb_node* node, int64_t n, double* out_host, void* stream) {
    DGPB_REQUIRE(ws && node && out_host && n >= 1, "NULL argument");
    return grad_pipeline(ws, node, n, true, nullptr, nullptr, out_host, (cudaStream_t)stream);
}

int dgpb_nllik_grad_dense_batch(dgpb_ws* ws, const dgpb_node* nodes, int B, int64_t n, double* out_host, int ldo,
                                int* status_host, void* stream) {
    DGPB_NVTX("dgpb:nllik_grad_dense_batch");
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(ws && nodes && out_host && status_host && n >= 1, "NULL argument");
    DGPB_REQUIRE(B >= 1 && B <= MAXB, "batch size out of range");
    DGPB_TRY(join_waves(ws, st));
    KernelDev kds[MAXB];
    const double* ys[MAXB];
    ScaleArgs sa;
    int Pmax = 0;
    for (int b = 0; b < B; ++b) {
        DGPB_TRY(make_kernel_dev(&nodes[b], n, nullptr, &kds[b]));
        ys[b] = nodes[b].output;
        sa.scale[b] = nodes[b].scale;
        sa.est[b] = nodes[b].scale_est;
        Pmax = std::max(Pmax, nodes[b].nlen + (nodes[b].nugget_est ? 1 : 0));
    }
    DGPB_REQUIRE(ldo >= Pmax + 2, "ldo too small");
    Geom g = make_geom(n, true);
    Batch bt;
    double* out;
    DGPB_TRY(setup_batch(ws, g, B, &bt, &out));
    DGPB_TRY(assemble(g, kds, ys, bt, B, st));
    DGPB_TRY(factorize(g, bt, B, st));
    DGPB_TRY(reduce_logdet_quad(g, bt, B, sa, out, st));
    const int nt = (int)cdiv(n, 64);
    const int ntiles = nt * (nt + 1) / 2;
    const int LDF = Pmax + 2;
    void *part, *fin;
    DGPB_TRY(ws->reserve(SLOT_PART, sizeof(double) * (size_t)ntiles * Pmax * B, &part));
    DGPB_TRY(ws->reserve(SLOT_MISC2, sizeof(double) * (size_t)LDF * B, &fin));
    DGPB_REQUIRE((size_t)LDF * B <= 2048, "result block too large for the staging buffer");
    for (int b = 0; b < B; ++b) {
        const int P = nodes[b].nlen + (nodes[b].nugget_est ? 1 : 0);
        double* pb = (double*)part + (size_t)ntiles * Pmax * b;
        grad_kernel<<<ntiles, 256, 0, st>>>(kds[b], bt.T[b], g.ld, g.n, g.npad, P, nodes[b].nugget_est, out + 4 * b, pb);
        DGPB_LAUNCHED();
        grad_finish_kernel<<<P, 256, 0, st>>>(pb, ntiles, P, g.n, nodes[b].scale_est, out + 4 * b, (double*)fin + (size_t)LDF * b);
        DGPB_LAUNCHED();
    }
    int* info_host = reinterpret_cast<int*>(ws->pinned + 2048);
    DGPB_CUDA_TRY(cudaMemcpyAsync(ws->pinned, fin, sizeof(double) * (size_t)LDF * B, cudaMemcpyDeviceToHost, st));
    DGPB_CUDA_TRY(cudaMemcpyAsync(info_host, bt.info, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    DGPB_CUDA_TRY(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b) {
        status_host[b] = info_host[b] != 0 ? DGPB_NOT_PD : DGPB_OK;
        for (int i = 0; i < LDF; ++i) out_host[(size_t)b * ldo + i] = ws->pinned[(size_t)LDF * b + i];
    }
    return DGPB_OK;
}

int dgpb_compute_stats(dgpb_ws* ws, const dgpb_node* node, int64_t n, double* Rinv, double* Rinv_y, void* stream) {
    DGPB_REQUIRE(ws && node && Rinv && Rinv_y && n >= 1, "NULL argument");
    return grad_pipeline(ws, node, n, false, Rinv, Rinv_y, nullptr, (cudaStream_t)stream);
}

int dgpb_compute_stats_shifted(dgpb_ws* ws, const dgpb_node* node, int64_t n, const double* diag_shift, double* Rinv,
                               double* Rinv_y, void* stream) {
    DGPB_REQUIRE(ws && node && diag_shift && Rinv && Rinv_y && n >= 1, "NULL argument");
    return grad_pipeline(ws, node, n, false, Rinv, Rinv_y, nullptr, (cudaStream_t)stream, diag_shift);
}

int dgpb_mvn_draw(dgpb_ws* ws, const dgpb_node* node, int64_t n, const double* z, double* nu, void* stream) {
    DGPB_NVTX("dgpb:mvn_draw");
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(ws && node && z && nu && n >= 1, "NULL argument");
    DGPB_TRY(join_waves(ws, st));
    KernelDev kd;
    DGPB_TRY(make_kernel_dev(node, n, nullptr, &kd));
    Geom g = make_geom(n, false);
    Batch bt;
    double* out;
    DGPB_TRY(setup_batch(ws, g, 1, &bt, &out));
    DGPB_TRY(assemble(g, &kd, nullptr, bt, 1, st));
    DGPB_TRY(factorize(g, bt, 1, st));
    DGPB_TRY(restore_diag_blocks(g, bt, 1, st));
    DGPB_TRY(launch_trmv(bt.T[0], g.ld, g.n, sqrt(node->scale), z, nu, st));
    int* info_host = reinterpret_cast<int*>(ws->pinned + 2048);
    DGPB_CUDA_TRY(cudaMemcpyAsync(info_host, bt.info, sizeof(int), cudaMemcpyDeviceToHost, st));
    DGPB_CUDA_TRY(cudaStreamSynchronize(st));
    if (info_host[0] != 0) {
        set_error("covariance is not positive definite (pivot %d)", info_host[0]);
        return DGPB_NOT_PD;
    }
    return DGPB_OK;
}

// Probe: time `reps` bulk trailing-update launches (K = 128, window = n) on B scratch matrices; see `flags` of
// update_kernel.  out_host[0] = average ms per launch, out_host[1] = algorithmic TFLOP/s.
int dgpb_probe_update(dgpb_ws* ws, int64_t n, int B, int flags, int reps, double* out_host) {
    DGPB_REQUIRE(ws && out_host && n >= 256 && B >= 1 && B <= MAXB, "bad argument");
    DGPB_TRY(configure_once());
    DGPB_TRY(join_waves(ws, 0));
    Geom g = make_geom(n, false);
    Batch bt;
    double* out;
    DGPB_TRY(setup_batch(ws, g, B, &bt, &out));
    DGPB_CUDA_TRY(cudaMemset(bt.T[0], 0, g.elems() * sizeof(double) * B));
    const int lo = 512, rows = g.R - lo;
    const int ntr = (int)cdiv(rows, TM);
    cudaEvent_t e0, e1;
    DGPB_CUDA_TRY(cudaEventCreate(&e0));
    DGPB_CUDA_TRY(cudaEventCreate(&e1));
    dim3 grid((unsigned)(ntr * (ntr + 1)), 1, (unsigned)B);
    const int kc = (flags >> 8) ? (flags >> 8) : 4;  // K = 32 kc (panel columns [0, K); the window starts at 512)
    flags &= 255;
    update_kernel<<<grid, 256, kUpdateSmem, 0>>>(bt, g.ld, 0, kc, lo, g.R, g.R, 0, flags);
    DGPB_CUDA_TRY(cudaEventRecord(e0, 0));
    for (int r = 0; r < reps; ++r) update_kernel<<<grid, 256, kUpdateSmem, 0>>>(bt, g.ld, 0, kc, lo, g.R, g.R, 0, flags);
    DGPB_CUDA_TRY(cudaEventRecord(e1, 0));
    DGPB_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    DGPB_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    out_host[0] = ms / reps;
    out_host[1] = (double)B * 0.5 * (double)rows * (double)(rows + 1) * 2.0 * 32.0 * kc / (out_host[0] * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return DGPB_OK;
}

// Probe: time `reps` complete batched log-likelihood pipelines (assemble K from synthetic inputs, factorise,
// reduce) of B matrices.  aug = 1 uses the [[K],[y'],[I]] layout of the gradient path (B must be 1).
// out_host = {ms per pipeline, TFLOP/s counting n^3/3 (plain) or n^3 (aug) per matrix}.
__global__ void probe_fill_kernel(double* __restrict__ x, int64_t len, unsigned seed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    unsigned h = (unsigned)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    x[i] = (double)h * (1.0 / 4294967296.0);
}

int dgpb_probe_factorize(dgpb_ws* ws, int64_t n, int B, int aug, int reps, double* out_host) {
    DGPB_REQUIRE(ws && out_host && n >= 64 && B >= 1 && B <= MAXB && reps >= 1, "bad argument");
    DGPB_REQUIRE(!aug || B == 1, "aug probe is single-matrix");
    DGPB_TRY(join_waves(ws, 0));
    const int D = 8;
    void* px;
    DGPB_TRY(ws->reserve(SLOT_MISC, sizeof(double) * (size_t)(D + 1) * n * B, &px));
    double* X = (double*)px;
    probe_fill_kernel<<<(unsigned)cdiv((int64_t)(D + 1) * n * B, 256), 256>>>(X, (int64_t)(D + 1) * n * B, 12345u);
    KernelDev kds[MAXB];
    const double* ys[MAXB];
    ScaleArgs sa;
    for (int b = 0; b < B; ++b) {
        KernelDev& kd = kds[b];
        kd.kind = DGPB_SEXP;
        kd.D = D;
        kd.ard = 0;
        kd.stride = 1;
        for (int d = 0; d < D; ++d) {
            kd.ptr[d] = X + ((size_t)b * (D + 1) + d) * n;
            kd.len[d] = 0.6;
        }
        kd.nugget = 1e-4;
        ys[b] = X + ((size_t)b * (D + 1) + D) * n;
        sa.scale[b] = 1.0;
        sa.est[b] = 0;
    }
    Geom g = make_geom(n, aug != 0);
    Batch bt;
    double* out;
    DGPB_TRY(setup_batch(ws, g, B, &bt, &out));
    cudaEvent_t e0, e1;
    DGPB_CUDA_TRY(cudaEventCreate(&e0));
    DGPB_CUDA_TRY(cudaEventCreate(&e1));
    cudaStream_t st = 0;
    auto once = [&]() -> int {
        DGPB_TRY(assemble(g, kds, ys, bt, B, st));
        DGPB_TRY(factorize(g, bt, B, st));
        DGPB_TRY(reduce_logdet_quad(g, bt, B, sa, out, st));
        return DGPB_OK;
    };
    DGPB_TRY(once());
    DGPB_CUDA_TRY(cudaEventRecord(e0, st));
    for (int r = 0; r < reps; ++r) DGPB_TRY(once());
    DGPB_CUDA_TRY(cudaEventRecord(e1, st));
    DGPB_CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    DGPB_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    int info[MAXB];
    DGPB_CUDA_TRY(cudaMemcpy(info, bt.info, sizeof(int) * B, cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; ++b) DGPB_REQUIRE(info[b] == 0, "probe matrix not positive definite");
    out_host[0] = ms / reps;
    const double nn = (double)n;
    out_host[1] = (double)B * nn * nn * nn * (aug ? 1.0 : 1.0 / 3.0) / (out_host[0] * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return DGPB_OK;
}

// Development/benchmark tunables: "hb" = hyper-block width (multiple of 128), "hb_min_w" = smallest remaining
// window factored with hyper-blocks, "hb_graded" = 128/256/512 ramp at the start, "ess_batch" = matrices per
// speculative ESS wave.  Returns DGPB_BAD_ARG for an unknown key.
int dgpb_tune(const char* key, int value) {
    DGPB_REQUIRE(key != nullptr, "NULL key");
    const std::string k(key);
    if (k == "hb") {
        DGPB_REQUIRE(value >= 128 && value % 128 == 0 && value <= 2048, "hb must be a multiple of 128 in [128, 2048]");
        g_hb = value;
    } else if (k == "hb_small_b") {
        g_hb_small_b = value;
    } else if (k == "hb_min_w") {
        DGPB_REQUIRE(value >= 0, "hb_min_w must be >= 0");
        g_hb_min_w = value;
    } else if (k == "ess_batch") {
        DGPB_REQUIRE(value >= 0 && value <= MAXB, "ess_batch out of range");
        g_ess_target_b = value;
    } else if (k == "ess_prefetch") {
        g_ess_prefetch = value != 0;
    } else if (k == "ess_overlap") {
        DGPB_REQUIRE(value >= 0 && value <= 2, "ess_overlap is 0, 1 or 2");
        g_ess_overlap = value;
    } else if (k == "ess_wave_total") {
        DGPB_REQUIRE(value >= 1 && value <= 64, "ess_wave_total out of range");
        g_ess_wave_total = value;
    } else if (k == "ess_trsv") {
        g_ess_cached_threshold = value != 0;
    } else if (k == "ess_rotate_w") {
        g_ess_rotate_w = value != 0;
    } else if (k == "linkgp_matern_tab") {
        linkgp_set_matern_tab(value);
    } else if (k == "linkgp_mma") {
        linkgp_set_mma(value);
    } else if (k == "vecchia_small") {
        vecchia_set_small(value);
    } else if (k == "knn_mma") {
        knn_set_mma(value);
    } else if (k == "pdl") {   // 0: plain launches; 1: dependent launches, early trigger for B <= 4; v > 1: for B <= v
        g_pdl = value != 0;
        g_pdl_early_b = value > 1 ? value : (value ? 4 : 0);
    } else if (k == "crit_stream") {
        g_crit_stream = value != 0;
    } else if (k == "hb_graded") {
        g_hb_graded = value != 0;
    } else {
        DGPB_REQUIRE(false, "unknown tunable");
    }
    return DGPB_OK;
}

int dgpb_profile(int on) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (on && g_prof.ev.empty()) {
        g_prof.ev.resize(4096);
        for (auto& e : g_prof.ev) DGPB_CUDA_TRY(cudaEventCreate(&e));
    }
    if (on) {
        g_prof.used = 0;
        g_prof.ms = g_prof.launches = g_prof.flops = g_prof.total_flops = g_prof.wasted_flops = 0.0;
    } else {
        DGPB_TRY(g_prof.drain());
    }
    g_prof.on = on != 0;
    return DGPB_OK;
}

int dgpb_profile_read(double* out_host) {
    DGPB_REQUIRE(out_host != nullptr, "NULL argument");
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    DGPB_TRY(g_prof.drain());
    out_host[0] = g_prof.ms;
    out_host[1] = g_prof.launches;
    out_host[2] = g_prof.flops;
    out_host[3] = g_prof.total_flops;
    out_host[4] = g_prof.wasted_flops;
    return DGPB_OK;
}

int dgpb_potrf(dgpb_ws* ws, double* A, int64_t n, int* info_host, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(ws && A && n >= 1, "NULL argument");
    DGPB_TRY(join_waves(ws, st));
    Geom g = make_geom(n, false);
    Batch bt;
    double* out;
    DGPB_TRY(setup_batch(ws, g, 1, &bt, &out));
    DGPB_CUDA_TRY(cudaMemsetAsync(bt.T[0], 0, g.elems() * sizeof(double), st));
    DGPB_CUDA_TRY(cudaMemcpy2DAsync(bt.T[0], g.ld * sizeof(double), A, n * sizeof(double), n * sizeof(double), n,
                                    cudaMemcpyDeviceToDevice, st));
    if (g.npad > g.n) {
        // identity padding on the diagonal
        std::vector<double> one(1, 1.0);
        for (int i = g.n; i < g.npad; ++i)
            DGPB_CUDA_TRY(cudaMemcpyAsync(bt.T[0] + (size_t)i * g.ld + i, one.data(), sizeof(double),
                                          cudaMemcpyHostToDevice, st));
    }
    DGPB_CUDA_TRY(cudaMemsetAsync(bt.info, 0, sizeof(int) * MAXB, st));
    DGPB_TRY(factorize(g, bt, 1, st));
    DGPB_TRY(restore_diag_blocks(g, bt, 1, st));
    DGPB_CUDA_TRY(cudaMemcpy2DAsync(A, n * sizeof(double), bt.T[0], g.ld * sizeof(double), n * sizeof(double), n,
                                    cudaMemcpyDeviceToDevice, st));
    int* ih = reinterpret_cast<int*>(ws->pinned + 2048);
    DGPB_CUDA_TRY(cudaMemcpyAsync(ih, bt.info, sizeof(int), cudaMemcpyDeviceToHost, st));
    DGPB_CUDA_TRY(cudaStreamSynchronize(st));
    if (info_host) *info_host = ih[0];
    return DGPB_OK;
}

}  // extern "C"

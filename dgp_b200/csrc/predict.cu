// predict.cu -- closed-form GP / linked-GP predictive moments on sm_100a.
//   gp()       dgpsi/functions.py:379-394   -> cross-kernel matrix + FP64-mma GEMM (R * R^-1) + row reductions
//   link_gp()  dgpsi/functions.py:396-430   -> streaming kernels over (test-point tile) x (training-pair tile):
//              IJ_sexp functions.py:432-451, IJ_matern functions.py:453-494, Jd/Jd0 vecchia.py:915-988,
//              trace_sum functions.py:496-506, quad vecchia.py:990-1000.
//   The J matrix (n x n per test point) and the sexp R2sexp/Psexp tables (kernel_class.py:752-764) are never
//   materialised: each J entry is produced in registers and immediately contracted with R^-1 and aa'.
#include "common.cuh"
#include "linkmath.cuh"

namespace dgpb {

__device__ __forceinline__ void dmma884p(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16p(void* smem_dst, const void* gsrc, bool pred) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}

// ------------------------------------------------------------------------------------------------
// general FP64 GEMM  C (M x N) = A (M x K) * B (N x K)'   (K a multiple of 32, lda/ldb even)
// 128x128 tile, 8 warps (2x4), warp tile 64x32, BK = 32, two cp.async stages.
// ------------------------------------------------------------------------------------------------
constexpr int GBK = 32;
constexpr int GLD = 36;  // 36 % 16 == 4: conflict-free fragment loads
constexpr size_t kGemmSmem = (size_t)2 * 2 * 128 * GLD * sizeof(double);

__global__ void __launch_bounds__(256, 1) gemm_nt_kernel(const double* __restrict__ A, int64_t lda,
                                                         const double* __restrict__ B, int64_t ldb,
                                                         double* __restrict__ C, int64_t ldc, int M, int N, int K) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
    auto stageA = [&](int s) { return smem + (size_t)s * 2 * 128 * GLD; };
    auto stageB = [&](int s) { return smem + (size_t)s * 2 * 128 * GLD + 128 * GLD; };
    auto load_stage = [&](int s, int kbase) {
        double* sA = stageA(s);
        double* sB = stageB(s);
        for (int c = tid; c < 128 * 16; c += 256) {
            int lr = c >> 4, co = (c & 15) * 2;
            bool oka = m0 + lr < M, okb = n0 + lr < N;
            cp_async16p(&sA[lr * GLD + co], A + (int64_t)(oka ? m0 + lr : 0) * lda + kbase + co, oka);
            cp_async16p(&sB[lr * GLD + co], B + (int64_t)(okb ? n0 + lr : 0) * ldb + kbase + co, okb);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    const int w = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int wm = w >> 2, wn = w & 3;
    double acc[8][4][2];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;
    const int nk = K / GBK;
    load_stage(0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) {
            load_stage((kt + 1) & 1, (kt + 1) * GBK);
            asm volatile("cp.async.wait_group 1;\n" ::);
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
        const double* sA = stageA(kt & 1);
        const double* sB = stageB(kt & 1);
#pragma unroll 2
        for (int kk = 0; kk < GBK; kk += 4) {
            double a[8], b[4];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) a[mi] = sA[(64 * wm + 8 * mi + g) * GLD + kk + t4];
#pragma unroll
            for (int nj = 0; nj < 4; ++nj) b[nj] = sB[(32 * wn + 8 * nj + g) * GLD + kk + t4];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) dmma884p(acc[mi][nj][0], acc[mi][nj][1], a[mi], b[nj]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        int gr = m0 + 64 * wm + 8 * mi + g;
        if (gr >= M) continue;
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
            int gc = n0 + 32 * wn + 8 * nj + 2 * t4;
            if (gc + 1 < N) {
                *reinterpret_cast<double2*>(&C[(int64_t)gr * ldc + gc]) = make_double2(acc[mi][nj][0], acc[mi][nj][1]);
            } else if (gc < N) {
                C[(int64_t)gr * ldc + gc] = acc[mi][nj][0];
            }
        }
    }
}

int launch_gemm_nt(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int M, int N,
                   int K, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        DGPB_CUDA_TRY(cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
        configured = true;
    }
    DGPB_REQUIRE(K % GBK == 0 && lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0, "gemm: K%32, even leading dims required");
    dim3 grid((unsigned)cdiv(N, 128), (unsigned)cdiv(M, 128));
    gemm_nt_kernel<<<grid, 256, kGemmSmem, st>>>(A, lda, B, ldb, C, ldc, M, N, K);
    DGPB_LAUNCHED();
    return DGPB_OK;
}

// ------------------------------------------------------------------------------------------------
// gp(): cross-kernel matrix  R[t][i] = k(x_t, W_i)   (K_vec_nb, vecchia.py:244-265), zero padded to ldr
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kcross_kernel(KernelDev kw, KernelDev kx, double* __restrict__ R, int64_t ldr,
                                                     int M, int n, int ncols) {
    __shared__ double xt[kMaxDim][64];
    __shared__ double xw[kMaxDim][64];
    const int tid = threadIdx.x;
    const int t0 = blockIdx.y * 64, i0 = blockIdx.x * 64;
    const int D = kw.D;
    for (int idx = tid; idx < D * 64; idx += 256) {
        int d = idx >> 6, l = idx & 63;
        xt[d][l] = t0 + l < M ? kx.x(d, t0 + l) : 0.0;
        xw[d][l] = i0 + l < n ? kw.x(d, i0 + l) : 0.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int e = 0; e < 16; ++e) {
        int idx = tid + 256 * e;
        int lt = idx >> 6, li = idx & 63;
        int gt = t0 + lt, gi = i0 + li;
        if (gt >= M || gi >= ncols) continue;
        double v = 0.0;
        if (gi < n) v = corr_pair(kw.kind, D, [&](int d) { return xw[d][li]; }, [&](int d) { return xt[d][lt]; });
        R[(int64_t)gt * ldr + gi] = v;
    }
}

// m_t = R_t . a ;  v_t = | scale (1 + nugget - R_t . (R R^-1)_t) |   -- one warp per test point
__global__ void gp_finish_kernel(const double* __restrict__ R, const double* __restrict__ TR, int64_t ldr, int M, int n,
                                 const double* __restrict__ alpha, double scale, double nugget,
                                 double* __restrict__ mean, double* __restrict__ var) {
    int t = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    int lane = threadIdx.x & 31;
    if (t >= M) return;
    const double* r = R + (int64_t)t * ldr;
    const double* q = TR + (int64_t)t * ldr;
    double s1 = 0.0, s2 = 0.0;
    for (int i = lane; i < n; i += 32) {
        double ri = r[i];
        s1 += ri * alpha[i];
        s2 += ri * q[i];
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
        mean[t] = s1;
        var[t] = fabs(scale * (1.0 + nugget - s2));
    }
}

// ------------------------------------------------------------------------------------------------
// link_gp(): shared argument block
// ------------------------------------------------------------------------------------------------
struct LinkArgs {
    int kind, n, Dw, Dz, M;
    const double* w1;     // n x Dw
    const double* gw;     // n x Dz or NULL
    const double* Rinv;   // n x n
    const double* alpha;  // n
    const double* m_in;   // M x Dw
    const double* v_in;   // M x Dw
    const double* z;      // M x Dz or NULL
    double len[kMaxDim];  // first Dw local, then Dz global (a shared length is replicated)
    double scale, nugget;
};

constexpr int TT = 8;   // test points per CTA
constexpr int PT = 32;  // pair tile edge

__device__ __forceinline__ void tile_from_linear(int t, int& ti, int& tj) {
    tri_index(t, ti, tj);
}

// global-dimension factor Iz_i = k(gw_i, z_t) on the Dz connected dims (K_vec_nb); returns the factor
__device__ __forceinline__ double global_factor(const LinkArgs& a, int gi, int gt) {
    if (a.Dz == 0) return 1.0;
    if (a.kind == DGPB_SEXP) {
        double dist = 0.0;
        for (int k = 0; k < a.Dz; ++k) {
            double l = a.len[a.Dw + k];
            double df = a.gw[(int64_t)gi * a.Dz + k] / l - a.z[(int64_t)gt * a.Dz + k] / l;
            dist += df * df;
        }
        return exp(-dist);
    }
    double coef = 1.0, s = 0.0;
    for (int k = 0; k < a.Dz; ++k) {
        double l = a.len[a.Dw + k];
        double r = fabs(a.gw[(int64_t)gi * a.Dz + k] / l - a.z[(int64_t)gt * a.Dz + k] / l);
        coef *= 1.0 + kSqrt5 * r + (5.0 / 3.0) * (r * r);
        s += r;
    }
    return coef * exp(-kSqrt5 * s);
}

// ------------------------------------------------------------------------------------------------
// link_gp mean:  m_t = sum_i I_i(t) a_i          (one warp per test point, fixed summation order)
// ------------------------------------------------------------------------------------------------
__global__ void linkgp_mean_kernel(LinkArgs a, double* __restrict__ mean) {
    int t = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    int lane = threadIdx.x & 31;
    if (t >= a.M) return;
    double Ic = 1.0;
    if (a.kind == DGPB_SEXP) {
        for (int k = 0; k < a.Dw; ++k) Ic *= 1.0 + 2.0 * a.v_in[(int64_t)t * a.Dw + k] / (a.len[k] * a.len[k]);
        Ic = 1.0 / sqrt(Ic);
    }
    double s = 0.0;
    for (int i = lane; i < a.n; i += 32) {
        double Ii;
        if (a.kind == DGPB_SEXP) {
            double e = 0.0;
            for (int k = 0; k < a.Dw; ++k) {
                double xz = a.w1[(int64_t)i * a.Dw + k] - a.m_in[(int64_t)t * a.Dw + k];
                e += xz * xz / (2.0 * a.v_in[(int64_t)t * a.Dw + k] + a.len[k] * a.len[k]);
            }
            Ii = Ic * exp(-e);
        } else {
            Ii = 1.0;
            for (int k = 0; k < a.Dw; ++k)
                Ii *= I_matern_dim(a.w1[(int64_t)i * a.Dw + k], a.m_in[(int64_t)t * a.Dw + k],
                                   a.v_in[(int64_t)t * a.Dw + k], a.len[k]);
        }
        Ii *= global_factor(a, i, t);
        s += Ii * a.alpha[i];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) mean[t] = s;
}

// ------------------------------------------------------------------------------------------------
// link_gp second moments, squared-exponential kernel.
// part[(chunk*M + t)*2 + {0,1}] = partial (a'Ja , tr(R^-1 J)) over the pair tiles of this chunk.
// ------------------------------------------------------------------------------------------------
template <int DW, int NP>
__global__ void __launch_bounds__(256) linkgp_sexp_pairs_kernel(LinkArgs a, int PC, double* __restrict__ part) {
    __shared__ double2 sAB[TT][DW];
    __shared__ double sJc[TT];
    __shared__ double sI[PT][DW + 1], sJ[PT][DW + 1];
    __shared__ double gzI[TT][PT], gzJ[TT][PT];
    __shared__ double alI[PT], alJ[PT];
    __shared__ double sred[8];
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * TT;
    for (int idx = tid; idx < TT * DW; idx += 256) {
        int t = idx / DW, k = idx % DW, gt = t0 + t;
        double2 ab = make_double2(0.0, 0.0);
        if (gt < a.M && k < a.Dw) {
            double l = a.len[k];
            double div = 2.0 * a.v_in[(int64_t)gt * a.Dw + k] / (l * l);
            ab.x = 2.0 * a.m_in[(int64_t)gt * a.Dw + k] / l;
            ab.y = 1.0 / (2.0 + 4.0 * div);
        }
        sAB[t][k] = ab;
    }
    if (tid < TT) {
        int gt = t0 + tid;
        double jc = 0.0;
        if (gt < a.M) {
            jc = 1.0;
            for (int k = 0; k < a.Dw; ++k) jc *= 1.0 + 4.0 * a.v_in[(int64_t)gt * a.Dw + k] / (a.len[k] * a.len[k]);
            jc = 1.0 / sqrt(jc);
        }
        sJc[tid] = jc;
    }
    double accq[TT], acct[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) accq[t] = acct[t] = 0.0;
    const int nt = (a.n + PT - 1) / PT;
    const int ntiles = nt * (nt + 1) / 2;
    for (int tl = blockIdx.y; tl < ntiles; tl += PC) {
        int ti, tj;
        tile_from_linear(tl, ti, tj);
        __syncthreads();
        for (int idx = tid; idx < PT * DW; idx += 256) {
            int r = idx / DW, k = idx % DW;
            int gi = ti * PT + r, gj = tj * PT + r;
            sI[r][k] = (gi < a.n && k < a.Dw) ? a.w1[(int64_t)gi * a.Dw + k] / a.len[k] : 0.0;
            sJ[r][k] = (gj < a.n && k < a.Dw) ? a.w1[(int64_t)gj * a.Dw + k] / a.len[k] : 0.0;
        }
        if (tid < PT) {
            int gi = ti * PT + tid, gj = tj * PT + tid;
            alI[tid] = gi < a.n ? a.alpha[gi] : 0.0;
            alJ[tid] = gj < a.n ? a.alpha[gj] : 0.0;
        }
        for (int idx = tid; idx < TT * PT; idx += 256) {
            int t = idx / PT, r = idx % PT, gt = t0 + t;
            int gi = ti * PT + r, gj = tj * PT + r;
            double e1 = 0.0, e2 = 0.0;
            if (a.Dz > 0 && gt < a.M) {
                for (int k = 0; k < a.Dz; ++k) {
                    double l = a.len[a.Dw + k];
                    double zt = a.z[(int64_t)gt * a.Dz + k] / l;
                    if (gi < a.n) {
                        double df = a.gw[(int64_t)gi * a.Dz + k] / l - zt;
                        e1 += df * df;
                    }
                    if (gj < a.n) {
                        double df = a.gw[(int64_t)gj * a.Dz + k] / l - zt;
                        e2 += df * df;
                    }
                }
            }
            gzI[t][r] = e1;
            gzJ[t][r] = e2;
        }
        __syncthreads();
#pragma unroll 1
        for (int e0 = 0; e0 < 4; e0 += NP) {
            double p[NP][DW], e2v[NP], wq[NP], wr[NP];
            int lis[NP], ljs[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                int idx = tid + 256 * (e0 + q);
                int li = idx >> 5, lj = idx & 31;
                int gi = ti * PT + li, gj = tj * PT + lj;
                lis[q] = li;
                ljs[q] = lj;
                bool valid = gi < a.n && gj <= gi;
                double wgt = valid ? (gi == gj ? 1.0 : 2.0) : 0.0;
                double rinv = valid ? a.Rinv[(int64_t)gi * a.n + gj] : 0.0;
                wq[q] = wgt * alI[li] * alJ[lj];
                wr[q] = wgt * rinv;
                double ee = 0.0;
#pragma unroll
                for (int k = 0; k < DW; ++k) {
                    double si = sI[li][k], sj = sJ[lj][k];
                    p[q][k] = si + sj;
                    double df = si - sj;
                    ee += df * df;
                }
                e2v[q] = (gi == gj) ? 0.0 : 0.5 * ee;
            }
#pragma unroll
            for (int t = 0; t < TT; ++t) {
                double E[NP];
#pragma unroll
                for (int q = 0; q < NP; ++q) E[q] = e2v[q] + gzI[t][lis[q]] + gzJ[t][ljs[q]];
#pragma unroll
                for (int k = 0; k < DW; ++k) {
                    double2 ab = sAB[t][k];
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        double u = p[q][k] - ab.x;
                        E[q] += (u * u) * ab.y;
                    }
                }
                double jc = sJc[t];
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    double Jv = jc * exp(-E[q]);
                    accq[t] += wq[q] * Jv;
                    acct[t] += wr[q] * Jv;
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < TT; ++t) {
        double s1 = block_sum<256>(accq[t], sred);
        double s2 = block_sum<256>(acct[t], sred);
        if (tid == 0 && t0 + t < a.M) {
            part[((int64_t)blockIdx.y * a.M + t0 + t) * 2 + 0] = s1;
            part[((int64_t)blockIdx.y * a.M + t0 + t) * 2 + 1] = s2;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// link_gp second moments, squared-exponential kernel, exponents on the FP64 tensor path.
//   J_ij(t) = jc_t exp(-E_t(i,j)),
//   E_t(i,j) = sum_k y_tk (p_k - x'_tk)^2 + 1/2 sum_k (xi_k - xj_k)^2 + sum_k [(ai_k - z_tk)^2 + (aj_k - z_tk)^2]
// with p = xi + xj (scaled local coordinates), a = scaled global coordinates, y = 1/(2 + 8 v/l^2), x' = 2 mu/l.
// Using (xi-xj)^2 = 2 (xi^2 + xj^2) - p^2 the exponent is BILINEAR in a test-point coefficient vector and a pair
// feature vector:
//   E = c_t + 1 * [S_i + S_j] + sum_k (y_tk - 1/2) p_k^2 + sum_k (-2 y_tk x'_tk) p_k + sum_k (-2 z_tk) (ai_k + aj_k),
//   S = sum_k x_k^2 + sum_k a_k^2 per training point,  c_t = sum_k y x'^2 + 2 sum_k z^2,
// i.e. an (8 test points) x (8 pairs) x (F = 1 + 2 Dw + Dz features) product per mma.m8n8k4 chain instead of
// 3 Dw + 3 Dz vector operations per (test point, pair).  What stays on the FP64 vector pipe is the exp and the
// two weighted accumulations.  Every feature is V_i[c] + V_j[c] (optionally squared) for one column c of the
// per-point table V = [x | a | S], so the feature generation is branch-free.
// part[(chunk*M + t)*2 + {0,1}] = partial (a'Ja , tr(R^-1 J)) over the pair tiles of this chunk.
// ------------------------------------------------------------------------------------------------
constexpr int LT = 16;   // test points per CTA (two m8 octets)
constexpr int LPT = 64;  // pair tile edge of the tensor-path kernel (64 octets per warp between barriers)

__device__ __forceinline__ void link_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// KST = number of k-steps when known at compile time (feature map and coefficient fragments live in registers,
// the k loop is unrolled), 0 = runtime loop with the tables in shared memory.
template <int KST>
__global__ void __launch_bounds__(256, 2) linkgp_sexp_mma_kernel(LinkArgs a, int PC, double* __restrict__ part) {
    extern __shared__ double lsm[];
    const int Dw = a.Dw, Dz = a.Dz, DV = Dw + Dz + 1;      // columns of V
    const int F = 1 + 2 * Dw + Dz, KS = (F + 3) / 4, F4 = 4 * KS;
    const int CS = (F4 % 16 == 4 || F4 % 16 == 12) ? F4 : F4 + 4;   // coefficient row stride: conflict-free fragments
    const int VS = DV | 1;                                           // V row stride (odd)
    double* coef = lsm;                    // [LT][CS]
    double* sct = coef + LT * CS;          // [LT]  c_t
    double* sjc = sct + LT;                // [LT]  jc_t
    double* VI = sjc + LT;                 // [PT][VS]
    double* VJ = VI + LPT * VS;             // [PT][VS]
    double* alI = VJ + LPT * VS;            // [PT]
    double* alJ = alI + LPT;                // [PT]
    double* sred = alJ + LPT;               // [8][LT][2]
    int* fcol = reinterpret_cast<int*>(sred + 8 * LT * 2);   // [F4]: column of V, bit 30 = square the sum
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int t0 = blockIdx.x * LT;
    // ---- per-CTA tables: feature map and test-point coefficients
    for (int f = tid; f < F4; f += 256) {
        int c = 0, sq = 0;
        if (f == 0) c = Dw + Dz;
        else if (f <= 2 * Dw) { c = (f - 1) >> 1; sq = ((f - 1) & 1) == 0; }
        else if (f < F) c = f - 1 - Dw;
        fcol[f] = c | (sq << 30);
    }
    for (int idx = tid; idx < LT * CS; idx += 256) coef[idx] = 0.0;
    __syncthreads();
    for (int idx = tid; idx < LT * (Dw + Dz); idx += 256) {
        const int t = idx / (Dw + Dz), k = idx - t * (Dw + Dz), gt = t0 + t;
        if (gt >= a.M) continue;
        if (k < Dw) {
            const double l = a.len[k];
            const double y = 1.0 / (2.0 + 8.0 * a.v_in[(int64_t)gt * Dw + k] / (l * l));
            const double xp = 2.0 * a.m_in[(int64_t)gt * Dw + k] / l;
            coef[t * CS + 1 + 2 * k] = y - 0.5;
            coef[t * CS + 2 + 2 * k] = -2.0 * y * xp;
        } else {
            const int kz = k - Dw;
            coef[t * CS + 1 + 2 * Dw + kz] = -2.0 * (a.z[(int64_t)gt * Dz + kz] / a.len[k]);
        }
    }
    if (tid < LT) {
        const int gt = t0 + tid;
        double c = 0.0, jc = 0.0;
        if (gt < a.M) {
            jc = 1.0;
            for (int k = 0; k < Dw; ++k) {
                const double l = a.len[k], v = a.v_in[(int64_t)gt * Dw + k];
                const double y = 1.0 / (2.0 + 8.0 * v / (l * l));
                const double xp = 2.0 * a.m_in[(int64_t)gt * Dw + k] / l;
                c += y * xp * xp;
                jc *= 1.0 + 4.0 * v / (l * l);
            }
            for (int k = 0; k < Dz; ++k) {
                const double zt = a.z[(int64_t)gt * Dz + k] / a.len[Dw + k];
                c += 2.0 * zt * zt;
            }
            jc = 1.0 / sqrt(jc);
            coef[tid * CS] = 1.0;
        }
        sct[tid] = c;
        sjc[tid] = jc;
    }
    double accq[2] = {0.0, 0.0}, acct[2] = {0.0, 0.0};
    const int nt = (a.n + LPT - 1) / LPT;
    const int ntiles = nt * (nt + 1) / 2;
    // compile-time KS: this lane's feature columns and its coefficient fragments, once per CTA
    constexpr int KR = KST > 0 ? KST : 1;
    int rcol[KR];
    double rco[2][KR];
    double rct[2], rjc[2];
    if (KST > 0) {
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < KR; ++ks) {
            rcol[ks] = fcol[4 * ks + t4];
#pragma unroll
            for (int h = 0; h < 2; ++h) rco[h][ks] = coef[(8 * h + g) * CS + 4 * ks + t4];
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            rct[h] = sct[8 * h + g];
            rjc[h] = sjc[8 * h + g];
        }
    }
    for (int tl = blockIdx.y; tl < ntiles; tl += PC) {
        int ti, tj;
        tile_from_linear(tl, ti, tj);
        __syncthreads();
        // V rows of the two point tiles: [x/l | a/l | S]
        for (int idx = tid; idx < 2 * LPT * (Dw + Dz); idx += 256) {
            const int side = idx / (LPT * (Dw + Dz)), rem = idx - side * LPT * (Dw + Dz);
            const int r = rem / (Dw + Dz), k = rem - r * (Dw + Dz);
            const int gp = (side ? tj : ti) * LPT + r;
            double v = 0.0;
            if (gp < a.n) v = (k < Dw ? a.w1[(int64_t)gp * Dw + k] : a.gw[(int64_t)gp * Dz + (k - Dw)]) / a.len[k];
            (side ? VJ : VI)[r * VS + k] = v;
        }
        if (tid < 2 * LPT) {
            const int side = tid / LPT, r = tid - side * LPT;
            const int gp = (side ? tj : ti) * LPT + r;
            (side ? alJ : alI)[r] = gp < a.n ? a.alpha[gp] : 0.0;
        }
        __syncthreads();
        if (tid < 2 * LPT) {
            const int side = tid / LPT, r = tid - side * LPT;
            double* V = (side ? VJ : VI) + r * VS;
            double s = 0.0;
            for (int k = 0; k < Dw + Dz; ++k) s += V[k] * V[k];
            V[Dw + Dz] = s;
        }
        __syncthreads();
        for (int o = w; o < LPT * LPT / 8; o += 8) {
            // B role: this lane generates the features of pair (li, lj = 8 (o & 3) + g)
            const int li = o >> 3, lj = 8 * (o & 7) + g;
            const int gi = ti * LPT + li, gj = tj * LPT + lj;
            const bool valid = gi < a.n && gj <= gi;
            if (!__any_sync(0xffffffffu, valid)) continue;
            const double wgt = valid ? (gi == gj ? 1.0 : 2.0) : 0.0;
            const double rinv = valid ? a.Rinv[(int64_t)gi * a.n + gj] : 0.0;
            const double wq_b = wgt * alI[li] * alJ[lj];
            const double* vi = VI + li * VS;
            const double* vj = VJ + lj * VS;
            double c[2][2];
            if (KST > 0) {
#pragma unroll
                for (int h = 0; h < 2; ++h) c[h][0] = c[h][1] = rct[h];
#pragma unroll
                for (int ks = 0; ks < KR; ++ks) {
                    const int col = rcol[ks] & 0xffff;
                    double b = vi[col] + vj[col];
                    if (rcol[ks] >> 30) b *= b;
#pragma unroll
                    for (int h = 0; h < 2; ++h) link_dmma(c[h][0], c[h][1], rco[h][ks], b);
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) c[h][0] = c[h][1] = sct[8 * h + g];
                for (int ks = 0; ks < KS; ++ks) {
                    const int fc = fcol[4 * ks + t4];
                    const int col = fc & 0xffff;
                    double b = vi[col] + vj[col];
                    if (fc >> 30) b *= b;
#pragma unroll
                    for (int h = 0; h < 2; ++h) link_dmma(c[h][0], c[h][1], coef[(8 * h + g) * CS + 4 * ks + t4], b);
                }
            }
            const double wr_b = wgt * rinv;
            // C role: test point g (of each octet), pairs 2 t4 and 2 t4 + 1 -> their weights live on lanes 8 t4, 8 t4 + 4
            // (exchanged here, after the tensor work, so the R^-1 load latency is covered by it)
            const double wq0 = __shfl_sync(0xffffffffu, wq_b, 8 * t4), wq1 = __shfl_sync(0xffffffffu, wq_b, 8 * t4 + 4);
            const double wr0 = __shfl_sync(0xffffffffu, wr_b, 8 * t4), wr1 = __shfl_sync(0xffffffffu, wr_b, 8 * t4 + 4);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double jc = KST > 0 ? rjc[h] : sjc[8 * h + g];
                const double J0 = jc * exp_nonpos(-c[h][0]), J1 = jc * exp_nonpos(-c[h][1]);
                accq[h] = fma(wq1, J1, fma(wq0, J0, accq[h]));
                acct[h] = fma(wr1, J1, fma(wr0, J0, acct[h]));
            }
        }
    }
    // lanes t4 = 0..3 of a row hold partial sums of the same test point; then the 8 warps through shared memory
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        accq[h] += __shfl_xor_sync(0xffffffffu, accq[h], 1);
        accq[h] += __shfl_xor_sync(0xffffffffu, accq[h], 2);
        acct[h] += __shfl_xor_sync(0xffffffffu, acct[h], 1);
        acct[h] += __shfl_xor_sync(0xffffffffu, acct[h], 2);
    }
    __syncthreads();
    if (t4 == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            sred[(w * LT + 8 * h + g) * 2 + 0] = accq[h];
            sred[(w * LT + 8 * h + g) * 2 + 1] = acct[h];
        }
    }
    __syncthreads();
    if (tid < LT * 2) {
        const int t = tid >> 1, which = tid & 1;
        double s = 0.0;
        for (int ww = 0; ww < 8; ++ww) s += sred[(ww * LT + t) * 2 + which];   // fixed order
        if (t0 + t < a.M) part[((int64_t)blockIdx.y * a.M + t0 + t) * 2 + which] = s;
    }
}

static size_t linkgp_mma_smem(int Dw, int Dz) {
    const int DV = Dw + Dz + 1, F = 1 + 2 * Dw + Dz, F4 = 4 * ((F + 3) / 4);
    const int CS = (F4 % 16 == 4 || F4 % 16 == 12) ? F4 : F4 + 4, VS = DV | 1;
    return sizeof(double) * (size_t)(LT * CS + 2 * LT + 2 * LPT * VS + 2 * LPT + 8 * LT * 2) + sizeof(int) * F4 + 16;
}

static int g_linkgp_mma = 1;   // dgpb_tune("linkgp_mma", 0): vector-pipe pair kernel
int linkgp_set_mma(int on) {
    g_linkgp_mma = on != 0;
    return DGPB_OK;
}

// ------------------------------------------------------------------------------------------------
// link_gp second moments, Matern-2.5 kernel (direct closed form per pair and dimension)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) linkgp_matern_pairs_kernel(LinkArgs a, int PC, double* __restrict__ part) {
    __shared__ double sM[TT][kMaxDim], sV[TT][kMaxDim];
    __shared__ double sI[PT][kMaxDim + 1], sJ[PT][kMaxDim + 1];
    __shared__ double gzI[TT][PT], gzJ[TT][PT];
    __shared__ double alI[PT], alJ[PT];
    __shared__ double sred[8];
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * TT;
    const int Dw = a.Dw;
    for (int idx = tid; idx < TT * Dw; idx += 256) {
        int t = idx / Dw, k = idx % Dw, gt = t0 + t;
        sM[t][k] = gt < a.M ? a.m_in[(int64_t)gt * Dw + k] : 0.0;
        sV[t][k] = gt < a.M ? a.v_in[(int64_t)gt * Dw + k] : 0.0;
    }
    double accq[TT], acct[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) accq[t] = acct[t] = 0.0;
    const int nt = (a.n + PT - 1) / PT;
    const int ntiles = nt * (nt + 1) / 2;
    for (int tl = blockIdx.y; tl < ntiles; tl += PC) {
        int ti, tj;
        tile_from_linear(tl, ti, tj);
        __syncthreads();
        for (int idx = tid; idx < PT * Dw; idx += 256) {
            int r = idx / Dw, k = idx % Dw;
            int gi = ti * PT + r, gj = tj * PT + r;
            sI[r][k] = gi < a.n ? a.w1[(int64_t)gi * Dw + k] : 0.0;
            sJ[r][k] = gj < a.n ? a.w1[(int64_t)gj * Dw + k] : 0.0;
        }
        if (tid < PT) {
            int gi = ti * PT + tid, gj = tj * PT + tid;
            alI[tid] = gi < a.n ? a.alpha[gi] : 0.0;
            alJ[tid] = gj < a.n ? a.alpha[gj] : 0.0;
        }
        for (int idx = tid; idx < TT * PT; idx += 256) {
            int t = idx / PT, r = idx % PT, gt = t0 + t;
            int gi = ti * PT + r, gj = tj * PT + r;
            gzI[t][r] = (gt < a.M && gi < a.n) ? global_factor(a, gi, gt) : 0.0;
            gzJ[t][r] = (gt < a.M && gj < a.n) ? global_factor(a, gj, gt) : 0.0;
        }
        __syncthreads();
#pragma unroll 1
        for (int e = 0; e < 4; ++e) {
            int idx = tid + 256 * e;
            int li = idx >> 5, lj = idx & 31;
            int gi = ti * PT + li, gj = tj * PT + lj;
            bool valid = gi < a.n && gj <= gi;
            if (!valid) continue;
            double wgt = (gi == gj) ? 1.0 : 2.0;
            double wq = wgt * alI[li] * alJ[lj];
            double wr = wgt * a.Rinv[(int64_t)gi * a.n + gj];
#pragma unroll 1
            for (int t = 0; t < TT; ++t) {
                if (t0 + t >= a.M) break;
                double Jv = 1.0;
                for (int k = 0; k < Dw; ++k) {
                    double zm = sM[t][k], zv = sV[t][k], l = a.len[k];
                    double xi = sI[li][k], xj = sJ[lj][k];
                    if (zv != 0.0) {
                        Jv *= (gi == gj) ? Jd0_dev(xi, zm, zv, l) : Jd_dev(xj, xi, zm, zv, l);
                    } else {
                        Jv *= matern_plain(zm - xi, l) * matern_plain(zm - xj, l);
                    }
                }
                Jv *= gzI[t][li] * gzJ[t][lj];
                // accumulate with a static index (TT is small): select by comparison to stay in registers
#pragma unroll
                for (int tt = 0; tt < TT; ++tt)
                    if (tt == t) {
                        accq[tt] += wq * Jv;
                        acct[tt] += wr * Jv;
                    }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < TT; ++t) {
        double s1 = block_sum<256>(accq[t], sred);
        double s2 = block_sum<256>(acct[t], sred);
        if (tid == 0 && t0 + t < a.M) {
            part[((int64_t)blockIdx.y * a.M + t0 + t) * 2 + 0] = s1;
            part[((int64_t)blockIdx.y * a.M + t0 + t) * 2 + 1] = s2;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// link_gp second moments, Matern-2.5 kernel, TABULATED form.
// Jd(x_i, x_j; mu, v, l) = P1 + P2 + P3 (vecchia.py:915-959) looks like ~400 flop + 3 erf + 5 exp per (pair,
// dimension, test point), but every transcendental in it depends on ONE training point and the test point:
//   P1 = e^{10v/l^2} ep(x1) ep(x2) [ AC(x2) E3A31 + BC(x2) E3A32 ]            (x1 <= x2)
//   P2 = ep(x1) em(x2) [ 1/2 E4A41 (F(x2) - F(x1)) + E4A42 G(x1) - E4A43 G(x2) ]
//   P3 = e^{10v/l^2} em(x1) em(x2) [ AD(x1) E5A51 + BD(x1) E5A52 ]
//   ep(x) = e^{sqrt5 (x - mu)/l}, em = 1/ep, AC = (1 + erf((muC - x)/sqrt(2v)))/2, BC = sqrt(v/2pi) e^{-(x-muC)^2/2v},
//   AD, BD the same around muD = mu + 2 sqrt5 v/l, F = erf((x - mu)/sqrt(2v)), G = sqrt(v/2pi) e^{-(x-mu)^2/2v},
// and E3Axx / E4Axx / E5Axx are polynomials in (x1, x2) whose coefficients E30..E54 do not depend on the test
// point.  So per pair tile and dimension the kernel tabulates the 8 transcendental values per (test point,
// training point) -- 2 x 32 points instead of 1024 pairs -- computes the 12 polynomial coefficients once per
// pair and is left with ~110 multiply-adds per (pair, dimension, test point).  Same closed form, same terms,
// different association (tests: equality with the direct kernel to 1e-10 and with the oracle).
// ------------------------------------------------------------------------------------------------
constexpr int MT = 4;   // test points per CTA of the tabulated kernel

__global__ void __launch_bounds__(256, 1) linkgp_matern_tab_kernel(LinkArgs a, int PC, double* __restrict__ part) {
    extern __shared__ double msm[];
    const int Dw = a.Dw, DS = Dw + 1;
    double* cst = msm;                         // [MT][Dw][16] per (test point, dimension) constants
    double* tab = cst + MT * Dw * 16;          // [2][MT][PT][8]
    double* sI = tab + 2 * MT * PT * 8;        // [PT][DS] raw local coordinates of the row / column points
    double* sJ = sI + PT * DS;
    double* gzI = sJ + PT * DS;                // [MT][PT] global-input factors
    double* gzJ = gzI + MT * PT;
    double* alI = gzJ + MT * PT;               // [PT]
    double* alJ = alI + PT;
    double* sred = alJ + PT;                   // [8]
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * MT;
    for (int idx = tid; idx < MT * Dw; idx += 256) {
        const int t = idx / Dw, k = idx - t * Dw, gt = t0 + t;
        const double l = a.len[k];
        const double zm = gt < a.M ? a.m_in[(int64_t)gt * Dw + k] : 0.0;
        const double zv = gt < a.M ? a.v_in[(int64_t)gt * Dw + k] : 0.0;
        const double muC = zm - 2.0 * kSqrt5 * zv / l, muD = zm + 2.0 * kSqrt5 * zv / l;
        const double c2 = muC * muC, d2 = muD * muD, z2 = zm * zm;
        double* c = cst + idx * 16;
        c[0] = zm;
        c[1] = zv;
        c[2] = muC;
        c[3] = c2 + zv;
        c[4] = c2 * muC + 3.0 * zv * muC;
        c[5] = c2 * c2 + 6.0 * zv * c2 + 3.0 * zv * zv;
        c[6] = muD;
        c[7] = d2 + zv;
        c[8] = d2 * muD + 3.0 * zv * muD;
        c[9] = d2 * d2 + 6.0 * zv * d2 + 3.0 * zv * zv;
        c[10] = z2 + zv;
        c[11] = z2 * zm + 3.0 * zv * zm;
        c[12] = z2 * z2 + 6.0 * zv * z2 + 3.0 * zv * zv;
        c[13] = exp(10.0 * zv / (l * l));
        c[14] = zv == 0.0 ? 1.0 : 0.0;   // deterministic input: J factor = k(mu - x_i) k(mu - x_j)
        c[15] = 0.0;
    }
    double accq[MT], acct[MT];
#pragma unroll
    for (int t = 0; t < MT; ++t) accq[t] = acct[t] = 0.0;
    const int nt = (a.n + PT - 1) / PT;
    const int ntiles = nt * (nt + 1) / 2;
    for (int tl = blockIdx.y; tl < ntiles; tl += PC) {
        int ti, tj;
        tile_from_linear(tl, ti, tj);
        __syncthreads();
        for (int idx = tid; idx < PT * Dw; idx += 256) {
            const int r = idx / Dw, k = idx - r * Dw;
            const int gi = ti * PT + r, gj = tj * PT + r;
            sI[r * DS + k] = gi < a.n ? a.w1[(int64_t)gi * Dw + k] : 0.0;
            sJ[r * DS + k] = gj < a.n ? a.w1[(int64_t)gj * Dw + k] : 0.0;
        }
        if (tid < PT) {
            const int gi = ti * PT + tid, gj = tj * PT + tid;
            alI[tid] = gi < a.n ? a.alpha[gi] : 0.0;
            alJ[tid] = gj < a.n ? a.alpha[gj] : 0.0;
        }
        for (int idx = tid; idx < MT * PT; idx += 256) {
            const int t = idx / PT, r = idx - t * PT, gt = t0 + t;
            const int gi = ti * PT + r, gj = tj * PT + r;
            gzI[idx] = (gt < a.M && gi < a.n) ? global_factor(a, gi, gt) : 0.0;
            gzJ[idx] = (gt < a.M && gj < a.n) ? global_factor(a, gj, gt) : 0.0;
        }
        // this thread's four pairs
        int lis[4], ljs[4];
        bool valid[4], diag[4];
        double wq[4], wr[4], Jv[4][MT];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = tid + 256 * e;
            lis[e] = idx >> 5;
            ljs[e] = idx & 31;
            const int gi = ti * PT + lis[e], gj = tj * PT + ljs[e];
            valid[e] = gi < a.n && gj <= gi;
            diag[e] = gi == gj;
            wr[e] = valid[e] ? (diag[e] ? 1.0 : 2.0) * a.Rinv[(int64_t)gi * a.n + gj] : 0.0;
#pragma unroll
            for (int t = 0; t < MT; ++t) Jv[e][t] = 1.0;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; ++e) wq[e] = valid[e] ? (diag[e] ? 1.0 : 2.0) * alI[lis[e]] * alJ[ljs[e]] : 0.0;
        for (int k = 0; k < Dw; ++k) {
            const double l = a.len[k];
            {   // table entry of (side, test point, training point) = thread
                const int side = tid >> 7, t = (tid >> 5) & 3, r = tid & 31;
                const double* c = cst + (t * Dw + k) * 16;
                const double x = (side ? sJ : sI)[r * DS + k];
                const double zm = c[0], zv = c[1];
                double* T = tab + ((side * MT + t) * PT + r) * 8;
                if (c[14] != 0.0) {
                    T[0] = matern_plain(zm - x, l);
                    T[1] = T[2] = T[3] = T[4] = T[5] = T[6] = T[7] = 0.0;
                } else {
                    const double isv = 1.0 / sqrt(2.0 * zv), gg = sqrt(0.5 * zv / M_PI), h = -0.5 / zv;
                    const double ep = exp(kSqrt5 * (x - zm) / l);
                    const double uC = x - c[2], uD = x - c[6], uZ = x - zm;
                    T[0] = ep;
                    T[1] = 1.0 / ep;
                    T[2] = 0.5 * (1.0 + erf(-uC * isv));
                    T[3] = gg * exp(h * uC * uC);
                    T[4] = 0.5 * (1.0 + erf(uD * isv));
                    T[5] = gg * exp(h * uD * uD);
                    T[6] = erf(uZ * isv);
                    T[7] = gg * exp(h * uZ * uZ);
                }
            }
            __syncthreads();
            const double l2 = l * l, l3 = l2 * l, i9l4 = 1.0 / (9.0 * l2 * l2);
            const double E4 = 25.0 * i9l4;   // E34 = E44 = E54
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (!valid[e]) continue;
                const double xi = sI[lis[e] * DS + k], xj = sJ[ljs[e] * DS + k];
                const bool sw = xi > xj;
                const double x1 = sw ? xj : xi, x2 = sw ? xi : xj;
                const double x1s = x1 * x1, x2s = x2 * x2, x12 = x1 * x2, xs = x1 + x2, xd = x2 - x1;
                const double q25 = 25.0 * x1s * x2s, s12 = x1s + x2s;
                const double E30 = 1.0 + (q25 - 3.0 * kSqrt5 * (3.0 * l3 + 5.0 * l * x12) * xs + 15.0 * l2 * (s12 + 3.0 * x12)) * i9l4;
                const double E31 = (18.0 * kSqrt5 * l3 + 15.0 * kSqrt5 * l * s12 - (75.0 * l2 + 50.0 * x12) * xs + 60.0 * kSqrt5 * l * x12) * i9l4;
                const double E32 = 5.0 * (5.0 * s12 + 15.0 * l2 - 9.0 * kSqrt5 * l * xs + 20.0 * x12) * i9l4;
                const double E33 = 10.0 * (3.0 * kSqrt5 * l - 5.0 * xs) * i9l4;
                const double E40 = 1.0 + (q25 + 3.0 * kSqrt5 * (3.0 * l3 - 5.0 * l * x12) * xd + 15.0 * l2 * (s12 - 3.0 * x12)) * i9l4;
                const double E41 = 5.0 * (3.0 * kSqrt5 * l * (x2s - x1s) + 3.0 * l2 * xs - 10.0 * x12 * xs) * i9l4;
                const double E42 = 5.0 * (5.0 * s12 - 3.0 * l2 - 3.0 * kSqrt5 * l * xd + 20.0 * x12) * i9l4;
                const double E43 = -50.0 * xs * i9l4;
                const double E50 = 1.0 + (q25 + 3.0 * kSqrt5 * (3.0 * l3 + 5.0 * l * x12) * xs + 15.0 * l2 * (s12 + 3.0 * x12)) * i9l4;
                const double E51 = (18.0 * kSqrt5 * l3 + 15.0 * kSqrt5 * l * s12 + (75.0 * l2 + 50.0 * x12) * xs + 60.0 * kSqrt5 * l * x12) * i9l4;
                const double E52 = 5.0 * (5.0 * s12 + 15.0 * l2 + 9.0 * kSqrt5 * l * xs + 20.0 * x12) * i9l4;
                const double E53 = 10.0 * (3.0 * kSqrt5 * l + 5.0 * xs) * i9l4;
                const double x1c = x1s * x1, x2c = x2s * x2;
#pragma unroll
                for (int t = 0; t < MT; ++t) {
                    const double* c = cst + (t * Dw + k) * 16;
                    const double* Ti = tab + ((0 * MT + t) * PT + lis[e]) * 8;
                    const double* Tj = tab + ((1 * MT + t) * PT + ljs[e]) * 8;
                    const double* T1 = sw ? Tj : Ti;   // table row of the smaller coordinate
                    const double* T2 = sw ? Ti : Tj;
                    double Jd;
                    if (c[14] != 0.0) {
                        Jd = Ti[0] * Tj[0];
                    } else {
                        const double zm = c[0], zv = c[1], muC = c[2], muD = c[6];
                        const double2 a01 = *reinterpret_cast<const double2*>(T1), a23 = *reinterpret_cast<const double2*>(T1 + 2);
                        const double2 a45 = *reinterpret_cast<const double2*>(T1 + 4), a67 = *reinterpret_cast<const double2*>(T1 + 6);
                        const double2 b01 = *reinterpret_cast<const double2*>(T2), b23 = *reinterpret_cast<const double2*>(T2 + 2);
                        const double2 b67 = *reinterpret_cast<const double2*>(T2 + 6);
                        const double E3A31 = E30 + muC * E31 + c[3] * E32 + c[4] * E33 + c[5] * E4;
                        const double c2 = muC * muC;
                        const double E3A32 = E31 + (muC + x2) * E32 + (c2 + 2.0 * zv + x2s + muC * x2) * E33 +
                                             (c2 * muC + x2c + x2 * c2 + muC * x2s + 3.0 * zv * x2 + 5.0 * zv * muC) * E4;
                        const double P1 = c[13] * a01.x * b01.x * (b23.x * E3A31 + b23.y * E3A32);
                        const double z2 = zm * zm;
                        const double E4A41 = E40 + zm * E41 + c[10] * E42 + c[11] * E43 + c[12] * E4;
                        const double E4A42 = E41 + (zm + x1) * E42 + (z2 + 2.0 * zv + x1s + zm * x1) * E43 +
                                             (z2 * zm + x1c + x1 * z2 + zm * x1s + 3.0 * zv * x1 + 5.0 * zv * zm) * E4;
                        const double E4A43 = E41 + (zm + x2) * E42 + (z2 + 2.0 * zv + x2s + zm * x2) * E43 +
                                             (z2 * zm + x2c + x2 * z2 + zm * x2s + 3.0 * zv * x2 + 5.0 * zv * zm) * E4;
                        const double P2 = a01.x * b01.y * (0.5 * E4A41 * (b67.x - a67.x) + E4A42 * a67.y - E4A43 * b67.y);
                        const double d2 = muD * muD;
                        const double E5A51 = E50 - muD * E51 + c[7] * E52 - c[8] * E53 + c[9] * E4;
                        const double E5A52 = E51 - (muD + x1) * E52 + (d2 + 2.0 * zv + x1s + muD * x1) * E53 -
                                             (d2 * muD + x1c + x1 * d2 + muD * x1s + 3.0 * zv * x1 + 5.0 * zv * muD) * E4;
                        const double P3 = c[13] * a01.y * b01.y * (a45.x * E5A51 + a45.y * E5A52);
                        Jd = P1 + P2 + P3;
                    }
                    Jv[e][t] *= Jd;
                }
            }
            __syncthreads();   // the tables are rewritten for the next dimension
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (!valid[e]) continue;
#pragma unroll
            for (int t = 0; t < MT; ++t) {
                const double J = Jv[e][t] * gzI[t * PT + lis[e]] * gzJ[t * PT + ljs[e]];
                accq[t] += wq[e] * J;
                acct[t] += wr[e] * J;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < MT; ++t) {
        const double s1 = block_sum<256>(accq[t], sred);
        const double s2 = block_sum<256>(acct[t], sred);
        if (tid == 0 && t0 + t < a.M) {
            part[((int64_t)blockIdx.y * a.M + t0 + t) * 2 + 0] = s1;
            part[((int64_t)blockIdx.y * a.M + t0 + t) * 2 + 1] = s2;
        }
    }
}

static size_t linkgp_matern_tab_smem(int Dw) {
    return sizeof(double) * (size_t)(MT * Dw * 16 + 2 * MT * PT * 8 + 2 * PT * (Dw + 1) + 2 * MT * PT + 2 * PT + 8);
}

static int g_linkgp_matern_tab = 1;   // dgpb_tune("linkgp_matern_tab", 0): direct closed form per pair
int linkgp_set_matern_tab(int on) {
    g_linkgp_matern_tab = on != 0;
    return DGPB_OK;
}

// v_t = | a'Ja - m_t^2 + scale (1 + nugget - tr(R^-1 J)) |     (functions.py:429)
__global__ void linkgp_finish_kernel(const double* __restrict__ part, int PC, int M, const double* __restrict__ mean,
                                     double scale, double nugget, double* __restrict__ var) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M) return;
    double q = 0.0, tr = 0.0;
    for (int c = 0; c < PC; ++c) {
        q += part[((int64_t)c * M + t) * 2 + 0];
        tr += part[((int64_t)c * M + t) * 2 + 1];
    }
    double m = mean[t];
    var[t] = fabs(q - m * m + scale * (1.0 + nugget - tr));
}

__global__ void aggregate_kernel(const double* __restrict__ means, const double* __restrict__ vars, int S, int64_t len,
                                 double* __restrict__ mu, double* __restrict__ sigma2) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    double sm = 0.0, s2 = 0.0;
    for (int s = 0; s < S; ++s) {
        double m = means[(int64_t)s * len + i];
        sm += m;
        s2 += m * m + vars[(int64_t)s * len + i];
    }
    double mbar = sm / (double)S;
    mu[i] = mbar;
    sigma2[i] = s2 / (double)S - mbar * mbar;
}

template <int DW, int NP>
static int launch_sexp_pairs(const LinkArgs& a, dim3 grid, int PC, double* part, cudaStream_t st) {
    linkgp_sexp_pairs_kernel<DW, NP><<<grid, 256, 0, st>>>(a, PC, part);
    DGPB_LAUNCHED();
    return DGPB_OK;
}

}  // namespace dgpb

using namespace dgpb;

extern "C" {

int dgpb_dgemm_nt(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K, void* stream) {
    DGPB_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "NULL or empty operand");
    DGPB_REQUIRE(K % 32 == 0 && N % 2 == 0, "dgpb_dgemm_nt needs K % 32 == 0 and even N");
    return launch_gemm_nt(A, K, B, K, C, N, (int)M, (int)N, (int)K, (cudaStream_t)stream);
}

int dgpb_gp_predict(dgpb_ws* ws, const double* x, int64_t M, const double* W, int64_t n, int64_t D, const double* Rinv,
                    const double* Rinv_y, const double* length_host, int64_t nlen, double scale, double nugget,
                    int kind, double* mean, double* var, void* stream) {
    DGPB_NVTX("dgpb:gp_predict");
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(ws && x && W && Rinv && Rinv_y && mean && var, "NULL argument");
    if (M == 0) return DGPB_OK;
    KernelDev kw, kx;
    DGPB_TRY(make_kernel_dev_rowmajor(W, D, length_host, nlen, nugget, kind, &kw));
    const int64_t Kp = round_up(n, 32);
    const int64_t chunk = 8192;
    void *pR, *pB, *pC;
    DGPB_TRY(ws->reserve(SLOT_GEMM_B, sizeof(double) * (size_t)n * Kp, &pB));
    DGPB_TRY(ws->reserve(SLOT_GEMM_A, sizeof(double) * (size_t)std::min(M, chunk) * Kp, &pR));
    DGPB_TRY(ws->reserve(SLOT_GEMM_C, sizeof(double) * (size_t)std::min(M, chunk) * Kp, &pC));
    // R^-1 with the contraction dimension zero padded to a multiple of 32
    DGPB_CUDA_TRY(cudaMemsetAsync(pB, 0, sizeof(double) * (size_t)n * Kp, st));
    DGPB_CUDA_TRY(cudaMemcpy2DAsync(pB, Kp * sizeof(double), Rinv, n * sizeof(double), n * sizeof(double), n,
                                    cudaMemcpyDeviceToDevice, st));
    for (int64_t m0 = 0; m0 < M; m0 += chunk) {
        const int Mc = (int)std::min(chunk, M - m0);
        DGPB_TRY(make_kernel_dev_rowmajor(x + m0 * D, D, length_host, nlen, nugget, kind, &kx));
        dim3 g1((unsigned)cdiv(Kp, 64), (unsigned)cdiv(Mc, 64));
        kcross_kernel<<<g1, 256, 0, st>>>(kw, kx, (double*)pR, Kp, Mc, (int)n, (int)Kp);
        DGPB_LAUNCHED();
        DGPB_TRY(launch_gemm_nt((double*)pR, Kp, (double*)pB, Kp, (double*)pC, Kp, Mc, (int)n, (int)Kp, st));
        gp_finish_kernel<<<(unsigned)cdiv((int64_t)Mc * 32, 256), 256, 0, st>>>((double*)pR, (double*)pC, Kp, Mc, (int)n,
                                                                               Rinv_y, scale, nugget, mean + m0,
                                                                               var + m0);
        DGPB_LAUNCHED();
    }
    return DGPB_OK;
}

int dgpb_linkgp_predict(dgpb_ws* ws, const double* m_in, const double* v_in, const double* z, int64_t M,
                        const double* w1, const double* gw, int64_t n, int64_t Dw, int64_t Dz, const double* Rinv,
                        const double* Rinv_y, const double* length_host, int64_t nlen, double scale, double nugget,
                        int kind, double* mean, double* var, void* stream) {
    DGPB_NVTX("dgpb:linkgp_predict");
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(ws && m_in && v_in && w1 && Rinv && Rinv_y && mean && var && length_host, "NULL argument");
    DGPB_REQUIRE(Dw >= 1 && Dz >= 0 && Dw + Dz <= kMaxDim, "dimension out of range");
    DGPB_REQUIRE(Dz == 0 || (z && gw), "z/gw required when Dz > 0");
    DGPB_REQUIRE(nlen == 1 || nlen == Dw + Dz, "len(length) must be 1 or Dw+Dz");
    DGPB_REQUIRE(kind == DGPB_SEXP || kind == DGPB_MATERN25, "unknown kernel kind");
    if (M == 0) return DGPB_OK;
    LinkArgs a;
    a.kind = kind;
    a.n = (int)n;
    a.Dw = (int)Dw;
    a.Dz = (int)Dz;
    a.w1 = w1;
    a.gw = gw;
    a.Rinv = Rinv;
    a.alpha = Rinv_y;
    for (int k = 0; k < kMaxDim; ++k) a.len[k] = 1.0;
    for (int k = 0; k < Dw + Dz; ++k) a.len[k] = length_host[nlen == 1 ? 0 : k];
    a.scale = scale;
    a.nugget = nugget;
    const int nt = (int)cdiv(n, PT);
    const int ntiles = nt * (nt + 1) / 2;
    // test points are processed in slabs so the partial buffer stays small
    const int64_t slab = 65536;
    for (int64_t m0 = 0; m0 < M; m0 += slab) {
        const int Mc = (int)std::min(slab, M - m0);
        a.M = Mc;
        a.m_in = m_in + m0 * Dw;
        a.v_in = v_in + m0 * Dw;
        a.z = z ? z + m0 * Dz : nullptr;
        const int ntt = (int)cdiv(Mc, TT);
        int PC = (int)cdiv(148 * 4, ntt);
        PC = std::max(1, std::min(PC, ntiles));
        void* part;
        DGPB_TRY(ws->reserve(SLOT_PART, sizeof(double) * (size_t)PC * Mc * 2, &part));
        linkgp_mean_kernel<<<(unsigned)cdiv((int64_t)Mc * 32, 256), 256, 0, st>>>(a, mean + m0);
        DGPB_LAUNCHED();
        dim3 grid((unsigned)ntt, (unsigned)PC);
        if (kind == DGPB_SEXP && g_linkgp_mma) {
            const int ntt2 = (int)cdiv(Mc, LT);
            const int nt2 = (int)cdiv(n, LPT);
            int PC2 = (int)cdiv(148 * 4, ntt2);
            PC2 = std::max(1, std::min(PC2, nt2 * (nt2 + 1) / 2));
            DGPB_TRY(ws->reserve(SLOT_PART, sizeof(double) * (size_t)PC2 * Mc * 2, &part));
            const size_t smem = linkgp_mma_smem((int)Dw, (int)Dz);
            static size_t configured = 0;
            if (smem > configured) {
                const int cap = (int)linkgp_mma_smem(kMaxDim, kMaxDim);
                DGPB_CUDA_TRY(cudaFuncSetAttribute(linkgp_sexp_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                DGPB_CUDA_TRY(cudaFuncSetAttribute(linkgp_sexp_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                DGPB_CUDA_TRY(cudaFuncSetAttribute(linkgp_sexp_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                DGPB_CUDA_TRY(cudaFuncSetAttribute(linkgp_sexp_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                DGPB_CUDA_TRY(cudaFuncSetAttribute(linkgp_sexp_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                DGPB_CUDA_TRY(cudaFuncSetAttribute(linkgp_sexp_mma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                DGPB_CUDA_TRY(cudaFuncSetAttribute(linkgp_sexp_mma_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                DGPB_CUDA_TRY(cudaFuncSetAttribute(linkgp_sexp_mma_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                DGPB_CUDA_TRY(cudaFuncSetAttribute(linkgp_sexp_mma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
                configured = linkgp_mma_smem(kMaxDim, kMaxDim);
            }
            const int KSv = (1 + 2 * (int)Dw + (int)Dz + 3) / 4;
            const dim3 grid2((unsigned)ntt2, (unsigned)PC2);
#define LINK_MMA_CASE(K)                                                                                   \
    case K:                                                                                                \
        linkgp_sexp_mma_kernel<K><<<grid2, 256, smem, st>>>(a, PC2, (double*)part);                        \
        break;
            switch (KSv <= 8 ? KSv : 0) {
                LINK_MMA_CASE(1) LINK_MMA_CASE(2) LINK_MMA_CASE(3) LINK_MMA_CASE(4) LINK_MMA_CASE(5) LINK_MMA_CASE(6)
                LINK_MMA_CASE(7) LINK_MMA_CASE(8)
                default:
                    linkgp_sexp_mma_kernel<0><<<grid2, 256, smem, st>>>(a, PC2, (double*)part);
            }
#undef LINK_MMA_CASE
            DGPB_LAUNCHED();
            linkgp_finish_kernel<<<(unsigned)cdiv(Mc, 256), 256, 0, st>>>((double*)part, PC2, Mc, mean + m0, scale,
                                                                         nugget, var + m0);
            DGPB_LAUNCHED();
            continue;
        }
        if (kind == DGPB_SEXP) {
            int rc;
            if (Dw <= 1) rc = launch_sexp_pairs<1, 4>(a, grid, PC, (double*)part, st);
            else if (Dw <= 2) rc = launch_sexp_pairs<2, 4>(a, grid, PC, (double*)part, st);
            else if (Dw <= 3) rc = launch_sexp_pairs<3, 4>(a, grid, PC, (double*)part, st);
            else if (Dw <= 4) rc = launch_sexp_pairs<4, 4>(a, grid, PC, (double*)part, st);
            else if (Dw <= 5) rc = launch_sexp_pairs<5, 4>(a, grid, PC, (double*)part, st);
            else if (Dw <= 6) rc = launch_sexp_pairs<6, 4>(a, grid, PC, (double*)part, st);
            else if (Dw <= 8) rc = launch_sexp_pairs<8, 4>(a, grid, PC, (double*)part, st);
            else if (Dw <= 10) rc = launch_sexp_pairs<10, 2>(a, grid, PC, (double*)part, st);
            else if (Dw <= 12) rc = launch_sexp_pairs<12, 2>(a, grid, PC, (double*)part, st);
            else if (Dw <= 16) rc = launch_sexp_pairs<16, 2>(a, grid, PC, (double*)part, st);
            else if (Dw <= 24) rc = launch_sexp_pairs<24, 1>(a, grid, PC, (double*)part, st);
            else rc = launch_sexp_pairs<32, 1>(a, grid, PC, (double*)part, st);
            DGPB_TRY(rc);
        } else if (g_linkgp_matern_tab) {
            const int ntt3 = (int)cdiv(Mc, MT);
            int PC3 = (int)cdiv(148 * 4, ntt3);
            PC3 = std::max(1, std::min(PC3, ntiles));
            DGPB_TRY(ws->reserve(SLOT_PART, sizeof(double) * (size_t)PC3 * Mc * 2, &part));
            static bool cfg = false;
            if (!cfg) {
                DGPB_CUDA_TRY(cudaFuncSetAttribute(linkgp_matern_tab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)linkgp_matern_tab_smem(kMaxDim)));
                cfg = true;
            }
            linkgp_matern_tab_kernel<<<dim3((unsigned)ntt3, (unsigned)PC3), 256, linkgp_matern_tab_smem((int)Dw), st>>>(
                a, PC3, (double*)part);
            DGPB_LAUNCHED();
            linkgp_finish_kernel<<<(unsigned)cdiv(Mc, 256), 256, 0, st>>>((double*)part, PC3, Mc, mean + m0, scale,
                                                                         nugget, var + m0);
            DGPB_LAUNCHED();
            continue;
        } else {
            linkgp_matern_pairs_kernel<<<grid, 256, 0, st>>>(a, PC, (double*)part);
            DGPB_LAUNCHED();
        }
        linkgp_finish_kernel<<<(unsigned)cdiv(Mc, 256), 256, 0, st>>>((double*)part, PC, Mc, mean + m0, scale, nugget,
                                                                     var + m0);
        DGPB_LAUNCHED();
    }
    return DGPB_OK;
}

int dgpb_aggregate(const double* means, const double* vars, int64_t S, int64_t len, double* mu, double* sigma2,
                   void* stream) {
    DGPB_NVTX("dgpb:aggregate");
    DGPB_REQUIRE(means && vars && mu && sigma2 && S >= 1, "NULL argument");
    if (len == 0) return DGPB_OK;
    aggregate_kernel<<<(unsigned)cdiv(len, 256), 256, 0, (cudaStream_t)stream>>>(means, vars, (int)S, len, mu, sigma2);
    DGPB_LAUNCHED();
    return DGPB_OK;
}

}  // extern "C"

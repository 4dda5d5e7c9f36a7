// vecchia.cu -- nearest-neighbour (Vecchia) conditioning on sm_100a.
//   nn / get_pred_nn      dgpsi/vecchia.py:20-109   -> exact FP64 brute-force kNN, bit-exact index output
//   vecchia_llik/_nllik   dgpsi/vecchia.py:164-242  -> one warp per conditioning block (<= 64 points):
//   L_matrix, fmvn_sp     dgpsi/vecchia.py:111-140,409-424     kernel block + Cholesky in shared memory
//   gp_vecch              dgpsi/vecchia.py:635-654
//   link_gp_vecch, IJ_nb  dgpsi/vecchia.py:758-907
// All per-point results are written to arrays and reduced in a fixed order (deterministic sums).
#include "common.cuh"
#include "linkmath.cuh"
#include "vecchia.cuh"

namespace dgpb {


// ------------------------------------------------------------------------------------------------
// per-warp conditioning-block machinery (packed lower-triangular storage: (i,j) -> i(i+1)/2 + j)
// ------------------------------------------------------------------------------------------------
int make_vkern(int kind, int64_t D, const double* length_host, int64_t nlen, VKern* vk) {
    DGPB_REQUIRE(kind == DGPB_SEXP || kind == DGPB_MATERN25, "unknown kernel kind");
    DGPB_REQUIRE(D >= 1 && D <= kMaxDim, "dimension out of range");
    DGPB_REQUIRE(length_host && (nlen == 1 || nlen == D), "len(length) must be 1 or D");
    vk->kind = kind;
    vk->D = (int)D;
    vk->ard = nlen != 1;
    for (int d = 0; d < kMaxDim; ++d) vk->len[d] = d < D ? length_host[nlen == 1 ? 0 : d] : 1.0;
    return DGPB_OK;
}

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// correlation of two points whose SCALED coordinates are rows of xl (b x D)
__device__ __forceinline__ double corr_rows(const VKern& vk, const double* xa, const double* xb) {
    if (vk.kind == DGPB_SEXP) {
        double dist = 0.0;
        for (int k = 0; k < vk.D; ++k) {
            double df = xa[k] - xb[k];
            dist += df * df;
        }
        return exp(-dist);
    }
    double coef = 1.0, s = 0.0;
    for (int k = 0; k < vk.D; ++k) {
        double r = fabs(xa[k] - xb[k]);
        coef *= 1.0 + kSqrt5 * r + (5.0 / 3.0) * (r * r);
        s += r;
    }
    return coef * exp(-kSqrt5 * s);
}

// K (packed lower) of the b gathered points; diagonal = 1 + nug[i]        (K_matrix_nb + add_to_diag_square)
__device__ __forceinline__ void warp_build_K(const VKern& vk, const double* xl, int b, const double* nug, double* A,
                                             int lane) {
    const int np = b * (b + 1) / 2;
    for (int p = lane; p < np; p += 32) {
        int i, j;
        tri_index(p, i, j);
        A[p] = (i == j) ? 1.0 + nug[i] : corr_rows(vk, xl + i * vk.D, xl + j * vk.D);
    }
    __syncwarp();
}

// in-place Cholesky (column Crout form); every lane recomputes the pivot, lanes split the rows
__device__ __forceinline__ void warp_chol(double* A, int b, int lane) {
    for (int j = 0; j < b; ++j) {
        double d = A[tri(j, j)];
        for (int p = 0; p < j; ++p) {
            double l = A[tri(j, p)];
            d -= l * l;
        }
        double sq = sqrt(d);
        double inv = 1.0 / sq;
        for (int i = j + 1 + lane; i < b; i += 32) {
            double s = A[tri(i, j)];
            for (int p = 0; p < j; ++p) s -= A[tri(i, p)] * A[tri(j, p)];
            A[tri(i, j)] = s * inv;
        }
        __syncwarp();
        if (lane == 0) A[tri(j, j)] = sq;
        __syncwarp();
    }
}

// w = L^-1 v in place (column-oriented), lanes over rows
__device__ __forceinline__ void warp_fwd(const double* L, double* v, int b, int lane) {
    for (int j = 0; j < b; ++j) {
        double wj = v[j] / L[tri(j, j)];
        __syncwarp();
        if (lane == 0) v[j] = wj;
        for (int i = j + 1 + lane; i < b; i += 32) v[i] -= L[tri(i, j)] * wj;
        __syncwarp();
    }
}

// v = L^-T v in place
__device__ __forceinline__ void warp_bwd(const double* L, double* v, int b, int lane) {
    for (int j = b - 1; j >= 0; --j) {
        double wj = v[j] / L[tri(j, j)];
        __syncwarp();
        if (lane == 0) v[j] = wj;
        for (int i = lane; i < j; i += 32) v[i] -= L[tri(j, i)] * wj;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// training-side kernel: likelihood terms, gradients, inverse-Cholesky rows
//   mode 0: vecchia_llik  -> vals[i*2 + {0,1}]   = (w_last^2, 2 log L_last)
//   mode 1: vecchia_nllik -> vals[i*(2P+2) + ..] = (w_last^2, 2 log L_last, dquad[P], dlogdet[P])
//   mode 2: L_matrix      -> Lout[i*m1 + c]
// ------------------------------------------------------------------------------------------------
__global__ void vecchia_train_kernel(VKern vk, const double* __restrict__ X, const double* __restrict__ y,
                                     const int64_t* __restrict__ NN, int64_t n, int m1, double nugget,
                                     const double* __restrict__ nugget_diag, int mode, int P, int nugget_est,
                                     double* __restrict__ vals, double* __restrict__ Lout, int per_warp) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * W + w;
    if (i >= n) return;
    double* base = smem + (size_t)w * per_warp;
    double* A = base;                                  // packed K -> L       m1(m1+1)/2
    double* xl = A + m1 * (m1 + 1) / 2;                // scaled coords       m1 * D
    double* yv = xl + m1 * vk.D;                       // y -> w              m1
    double* nug = yv + m1;                             // nugget_i            m1
    double* bv = nug + m1;                             // L^-T e_last         m1
    double* vp = bv + m1;                              // P x m1 (mode 1)
    int* idx = reinterpret_cast<int*>(vp + (mode == 1 ? P * m1 : 0));
    // idx = NN[i][valid][::-1]  (ascending, the point itself last)
    int b = 0;
    for (int c = 0; c < m1; ++c) b += NN[i * m1 + c] >= 0;
    for (int c = lane; c < b; c += 32) idx[c] = (int)NN[i * m1 + (b - 1 - c)];
    __syncwarp();
    for (int p = lane; p < b * vk.D; p += 32) {
        int r = p / vk.D, k = p % vk.D;
        xl[p] = X[(int64_t)idx[r] * vk.D + k] / vk.len[k];
    }
    for (int c = lane; c < b; c += 32) {
        yv[c] = y ? y[idx[c]] : 0.0;
        nug[c] = nugget * (nugget_diag ? nugget_diag[idx[c]] : 1.0);
        bv[c] = (c == b - 1) ? 1.0 : 0.0;
    }
    __syncwarp();
    warp_build_K(vk, xl, b, nug, A, lane);
    warp_chol(A, b, lane);
    if (mode == 2) {
        warp_bwd(A, bv, b, lane);
        for (int c = lane; c < m1; c += 32) Lout[i * m1 + c] = c < b ? bv[b - 1 - c] : 0.0;
        return;
    }
    warp_fwd(A, yv, b, lane);
    const double wl = yv[b - 1], Lll = A[tri(b - 1, b - 1)];
    const int stride = mode == 1 ? 2 * P + 2 : 2;
    if (lane == 0) {
        vals[i * stride + 0] = wl * wl;
        vals[i * stride + 1] = 2.0 * log(fabs(Lll));
    }
    if (mode != 1) return;
    warp_bwd(A, bv, b, lane);
    // vp[p][r] = (dK_p b)_r with the kernel derivatives recomputed from the scaled coordinates
    const int nl = vk.ard ? vk.D : 1;
    for (int r = lane; r < b; r += 32) {
        for (int p = 0; p < nl; ++p) vp[p * m1 + r] = 0.0;
        for (int c = 0; c < b; ++c) {
            if (c == r) continue;
            const double* xa = xl + r * vk.D;
            const double* xb = xl + c * vk.D;
            double Kv = corr_rows(vk, xa, xb) * bv[c];
            double csum = 0.0;
            for (int k = 0; k < vk.D; ++k) {
                double df = xa[k] - xb[k];
                double ck;
                if (vk.kind == DGPB_SEXP) {
                    ck = 2.0 * (df * df);
                } else {
                    double rr = fabs(df);
                    double el1 = 1.0 + kSqrt5 * rr, el2 = (5.0 / 3.0) * (rr * rr);
                    ck = el2 * el1 / (el1 + el2);
                }
                if (vk.ard) vp[k * m1 + r] += ck * Kv; else csum += ck;
            }
            if (!vk.ard) vp[r] += csum * Kv;
        }
        if (nugget_est) vp[nl * m1 + r] = nug[r] * bv[r];
    }
    __syncwarp();
    for (int p = 0; p < P; ++p) {
        double* v = vp + p * m1;
        warp_fwd(A, v, b, lane);
        double s = 0.0;
        for (int c = lane; c < b; c += 32) s += yv[c] * v[c];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            double last = v[b - 1];
            vals[i * stride + 2 + p] = 2.0 * s * wl - last * wl * wl;
            vals[i * stride + 2 + P + p] = last;
        }
    }
}

// column sums of vals (rows x cols) in a fixed order: one CTA per column
__global__ void __launch_bounds__(1024) colsum_kernel(const double* __restrict__ vals, int64_t rows, int cols,
                                                      double* __restrict__ out) {
    __shared__ double sred[32];
    const int c = blockIdx.x, tid = threadIdx.x;
    double s = 0.0;
    for (int64_t r = tid; r < rows; r += 1024) s += vals[r * cols + c];
    s = block_sum<1024>(s, sred);
    if (tid == 0) out[c] = s;
}

// ------------------------------------------------------------------------------------------------
// sparse forward solve  x_i = (z_i - sum_{j>=1} L[i,j] x[NN[i,j]]) / L[i,0]     (forward_solve_sp)
// Dependency-driven: one warp per row, rows claimed in increasing order through a ticket so every
// dependency (always a smaller row index) is owned by a warp that is already running or finished.
// The solution vector itself is the ready flag: it is pre-filled with a sentinel NaN payload that arithmetic
// cannot produce, a waiter polls the 8-byte word with volatile loads (and a short nanosleep back-off) and gets
// readiness and value in one transaction -- no atomics, no fences.  (The first version polled a separate flag
// array with atomicAdd(.., 0): ~300k spinning atomics saturated the L2 atomic units, 8.8 ms per solve at
// n = 100k, two thirds of a Vecchia I-step.)
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long kSpSentinel = 0x7ff8dead0000beefULL;

__global__ void sp_fill_kernel(unsigned long long* __restrict__ x, int64_t n, unsigned int* ticket) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = kSpSentinel;
    if (i == 0) *ticket = 0u;
}

__global__ void __launch_bounds__(256) sp_solve_kernel(const double* __restrict__ L, const int64_t* __restrict__ NN, int64_t n,
                                                       int m1, double inv_sqrt_scale, const double* __restrict__ z,
                                                       double* x, unsigned int* ticket) {
    __shared__ unsigned int s_blk;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    if (threadIdx.x == 0) s_blk = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t i = (int64_t)s_blk * W + w;
    if (i >= n) return;
    const int kmax = (int)min((int64_t)m1, i + 1);
    double s = 0.0;
    for (int j = 1 + lane; j < kmax; j += 32) {
        const int64_t dep = NN[i * m1 + j];
        const double lij = L[i * m1 + j] * inv_sqrt_scale;
        const volatile unsigned long long* px = reinterpret_cast<const volatile unsigned long long*>(x + dep);
        // bounded spin: a dependency that never arrives (corrupt NNarray) poisons the result instead of
        // hanging the device
        unsigned long long bits = *px;
        int spins = 0;
        while (bits == kSpSentinel && ++spins < (1 << 22)) {
            __nanosleep(40);
            bits = *px;
        }
        const double xd = bits == kSpSentinel ? NAN : __longlong_as_double((long long)bits);
        s += lij * xd;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        double r = (z[i] - s) / (L[i * m1] * inv_sqrt_scale);
        unsigned long long rb = (unsigned long long)__double_as_longlong(r);
        if (rb == kSpSentinel) rb = 0x7ff8000000000000ULL;   // cannot happen for computed values; keep waiters safe
        *reinterpret_cast<volatile unsigned long long*>(x + i) = rb;
    }
}

// ------------------------------------------------------------------------------------------------
// prediction: gp_vecch (mode 0) and link_gp_vecch (mode 1), one warp per test point
// ------------------------------------------------------------------------------------------------
struct VPredArgs {
    int mode;
    int64_t M, n;
    int Dw, Dz, mp;
    const double* xq;    // mode 0: M x D test inputs;  mode 1: M x Dw means
    const double* vq;    // mode 1: M x Dw variances
    const double* zq;    // mode 1: M x Dz deterministic global inputs or NULL
    const double* w1;    // mode 0: n x D ; mode 1: n x Dw
    const double* gw;    // mode 1: n x Dz or NULL
    const double* y;     // n
    const int64_t* NN;   // M x mp
    const double* nugget_diag;
    double scale, nugget;
    double* mean;
    double* var;
};

__global__ void vecchia_pred_kernel(VKern vk, VPredArgs a, int per_warp) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t t = (int64_t)blockIdx.x * W + w;
    if (t >= a.M) return;
    const int D = vk.D, mp = a.mp;
    const int bmax = mp + 1;
    double* base = smem + (size_t)w * per_warp;
    double* A = base;                           // packed K -> L -> (mode 1) L^-1
    double* xl = A + bmax * (bmax + 1) / 2;     // scaled coords (bmax x D)
    double* yv = xl + bmax * D;                 // bmax
    double* nug = yv + bmax;                    // bmax
    double* Jm = nug + bmax;                    // mode 1: packed J
    double* Kinv = Jm + (a.mode ? bmax * (bmax + 1) / 2 : 0);   // mode 1: packed K^-1
    double* Iv = Kinv + (a.mode ? bmax * (bmax + 1) / 2 : 0);   // mode 1: I, then alpha
    double* al = Iv + (a.mode ? bmax : 0);
    int* idx = reinterpret_cast<int*>(al + (a.mode ? bmax : 0));
    int nb = 0;
    for (int c = 0; c < mp; ++c) nb += a.NN[t * mp + c] >= 0;
    for (int c = lane; c < nb; c += 32) idx[c] = (int)a.NN[t * mp + c];
    __syncwarp();
    if (a.mode == 0) {
        const int b = nb + 1;
        for (int p = lane; p < b * D; p += 32) {
            int r = p / D, k = p % D;
            double v = r < nb ? a.w1[(int64_t)idx[r] * D + k] : a.xq[t * D + k];
            xl[p] = v / vk.len[k];
        }
        for (int c = lane; c < b; c += 32) {
            yv[c] = c < nb ? a.y[idx[c]] : 0.0;
            nug[c] = c < nb ? a.nugget * (a.nugget_diag ? a.nugget_diag[idx[c]] : 1.0) : a.nugget;
        }
        __syncwarp();
        warp_build_K(vk, xl, b, nug, A, lane);
        warp_chol(A, b, lane);
        warp_fwd(A, yv, nb, lane);   // only the leading nb x nb block is needed for L11^-1 y
        double s = 0.0;
        for (int c = lane; c < nb; c += 32) s += A[tri(nb, c)] * yv[c];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            double Lll = A[tri(nb, nb)];
            a.mean[t] = s;
            a.var[t] = a.scale * Lll * Lll;
        }
        return;
    }
    // ---- linked prediction on the neighbour set (no test point in the block) ----
    const int b = nb, Dw = a.Dw, Dz = a.Dz;
    for (int p = lane; p < b * D; p += 32) {
        int r = p / D, k = p % D;
        double v = k < Dw ? a.w1[(int64_t)idx[r] * Dw + k] : a.gw[(int64_t)idx[r] * Dz + (k - Dw)];
        xl[p] = v / vk.len[k];
    }
    for (int c = lane; c < b; c += 32) {
        yv[c] = a.y[idx[c]];
        nug[c] = a.nugget * (a.nugget_diag ? a.nugget_diag[idx[c]] : 1.0);
    }
    __syncwarp();
    warp_build_K(vk, xl, b, nug, A, lane);
    warp_chol(A, b, lane);
    // Iz factors from the connected global dims (K_vec_nb on the scaled coords)
    const double* zm = a.xq + t * Dw;
    const double* zv = a.vq + t * Dw;
    for (int c = lane; c < b; c += 32) {
        double f = 1.0;
        if (Dz > 0) {
            if (vk.kind == DGPB_SEXP) {
                double dist = 0.0;
                for (int k = 0; k < Dz; ++k) {
                    double df = xl[c * D + Dw + k] - a.zq[t * Dz + k] / vk.len[Dw + k];
                    dist += df * df;
                }
                f = exp(-dist);
            } else {
                double coef = 1.0, s = 0.0;
                for (int k = 0; k < Dz; ++k) {
                    double r = fabs(xl[c * D + Dw + k] - a.zq[t * Dz + k] / vk.len[Dw + k]);
                    coef *= 1.0 + kSqrt5 * r + (5.0 / 3.0) * (r * r);
                    s += r;
                }
                f = coef * exp(-kSqrt5 * s);
            }
        }
        al[c] = f;  // temporarily holds Iz
    }
    __syncwarp();
    // I vector and J matrix (IJ_nb, vecchia.py:838-907)
    double Ic = 1.0, Jc = 1.0;
    if (vk.kind == DGPB_SEXP) {
        for (int k = 0; k < Dw; ++k) {
            double l2 = vk.len[k] * vk.len[k];
            Ic *= 1.0 + 2.0 * zv[k] / l2;
            Jc *= 1.0 + 4.0 * zv[k] / l2;
        }
        Ic = 1.0 / sqrt(Ic);
        Jc = 1.0 / sqrt(Jc);
    }
    for (int c = lane; c < b; c += 32) {
        double v;
        if (vk.kind == DGPB_SEXP) {
            double e = 0.0;
            for (int k = 0; k < Dw; ++k) {
                double xz = a.w1[(int64_t)idx[c] * Dw + k] - zm[k];
                e += xz * xz / (2.0 * zv[k] + vk.len[k] * vk.len[k]);
            }
            v = Ic * exp(-e);
        } else {
            v = 1.0;
            for (int k = 0; k < Dw; ++k) v *= I_matern_dim(a.w1[(int64_t)idx[c] * Dw + k], zm[k], zv[k], vk.len[k]);
        }
        Iv[c] = v * al[c];
    }
    const int np = b * (b + 1) / 2;
    for (int p = lane; p < np; p += 32) {
        int i, j;
        tri_index(p, i, j);
        const double* xi = a.w1 + (int64_t)idx[i] * Dw;
        const double* xj = a.w1 + (int64_t)idx[j] * Dw;
        double v;
        if (vk.kind == DGPB_SEXP) {
            double e = 0.0;
            for (int k = 0; k < Dw; ++k) {
                double l2 = vk.len[k] * vk.len[k];
                double xzi = xi[k] - zm[k], xzj = xj[k] - zm[k];
                if (i == j) {
                    e += 2.0 * xzi * xzi / (4.0 * zv[k] + l2);
                } else {
                    double sp = xzi + xzj, sd = xzi - xzj;
                    e += sp * sp / (8.0 * zv[k] + 2.0 * l2) + sd * sd / (2.0 * l2);
                }
            }
            v = Jc * exp(-e);
        } else {
            v = 1.0;
            for (int k = 0; k < Dw; ++k) {
                if (zv[k] != 0.0) {
                    v *= (i == j) ? Jd0_dev(xi[k], zm[k], zv[k], vk.len[k]) : Jd_dev(xj[k], xi[k], zm[k], zv[k], vk.len[k]);
                } else {
                    v *= matern_plain(zm[k] - xi[k], vk.len[k]) * matern_plain(zm[k] - xj[k], vk.len[k]);
                }
            }
        }
        Jm[p] = v * al[i] * al[j];
    }
    __syncwarp();
    // L^-1 in place: lane c owns column c (forward substitution on e_c)
    for (int c = lane; c < b; c += 32) {
        double dcc = 1.0 / A[tri(c, c)];
        // compute column into Kinv scratch column-wise (use Kinv as temporary full storage of L^-1 packed)
        Kinv[tri(c, c)] = dcc;
        for (int i = c + 1; i < b; ++i) {
            double s = 0.0;
            for (int k = c; k < i; ++k) s += A[tri(i, k)] * Kinv[tri(k, c)];
            Kinv[tri(i, c)] = -s / A[tri(i, i)];
        }
    }
    __syncwarp();
    for (int p = lane; p < np; p += 32) A[p] = Kinv[p];   // A <- L^-1
    __syncwarp();
    // K^-1 = L^-T L^-1 (packed lower):  Kinv_ij = sum_{p >= i} Linv[p][i] Linv[p][j],  i >= j
    for (int p = lane; p < np; p += 32) {
        int i, j;
        tri_index(p, i, j);
        double s = 0.0;
        for (int r = i; r < b; ++r) s += A[tri(r, i)] * A[tri(r, j)];
        Kinv[p] = s;
    }
    __syncwarp();
    // alpha = K^-1 y
    for (int c = lane; c < b; c += 32) {
        double s = 0.0;
        for (int k = 0; k < b; ++k) s += Kinv[k <= c ? tri(c, k) : tri(k, c)] * yv[k];
        al[c] = s;
    }
    __syncwarp();
    double tr = 0.0, qd = 0.0, mt = 0.0;
    for (int p = lane; p < np; p += 32) {
        int i, j;
        tri_index(p, i, j);
        double wgt = (i == j) ? 1.0 : 2.0;
        tr += wgt * Kinv[p] * Jm[p];
        qd += wgt * Jm[p] * al[i] * al[j];
    }
    for (int c = lane; c < b; c += 32) mt += Iv[c] * al[c];
    for (int o = 16; o > 0; o >>= 1) {
        tr += __shfl_xor_sync(0xffffffffu, tr, o);
        qd += __shfl_xor_sync(0xffffffffu, qd, o);
        mt += __shfl_xor_sync(0xffffffffu, mt, o);
    }
    if (lane == 0) {
        a.mean[t] = mt;
        a.var[t] = fabs(qd - mt * mt + a.scale * (1.0 + a.nugget - tr));
    }
}

// ------------------------------------------------------------------------------------------------
// Register-resident variant for conditioning blocks of b <= 32 points (m <= 31: the reference's training
// default m = 25 and BASELINE config 4): lane i owns ROW i of the block's covariance in registers.
//   * row build: lane i evaluates k(x_i, x_k) for every k (coordinates staged in shared memory, row stride odd so
//     the per-lane rows are conflict-free and x_k is a broadcast);
//   * right-looking Cholesky: the pivot and the column entries l_kj travel by shuffles, the rank-1 update is one
//     FMA per (row, column) in registers -- no shared-memory read-modify-write, no barrier per column;
//   * the forward solve L^-1 y is fused into the same column loop (y is an extra right-hand side: w_j is final
//     as soon as column j is).
// PRED = true : gp_vecch (vecchia.py:635-654): block = [neighbours ; test point], y_test = 0, so the last
//               solve entry is -m / L_last and  m = -w_last L_last,  v = scale L_last^2.
// PRED = false: vecchia_llik terms (vecchia.py:164-180): vals[i*2 + {0,1}] = (w_last^2, 2 log L_last).
// ------------------------------------------------------------------------------------------------
struct VSmallArgs {
    int64_t count;        // test points (PRED) or training points
    int cols;             // columns of NN (mp or m1)
    const double* xq;     // PRED: count x D test inputs
    const double* X;      // n x D training inputs (train: Vecchia order)
    const double* y;      // n
    const int64_t* NN;    // count x cols
    const double* nugget_diag;
    double scale, nugget;
    double* out0;         // PRED: mean;  train: vals (count x 2)
    double* out1;         // PRED: var
};

template <bool PRED>
__global__ void __launch_bounds__(256, 3) vecchia_small_kernel(VKern vk, VSmallArgs a) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t t = (int64_t)blockIdx.x * W + w;
    if (t >= a.count) return;
    const int D = vk.D, Dp = D | 1;
    double* xl = smem + (size_t)w * (32 * Dp + 64);
    double* colb = xl + 32 * Dp;   // [2][32]: the current column of L, read back as broadcasts
    // block membership: PRED -> NN[t][0..nb) then the test point; train -> valid NN entries reversed (ascending,
    // the point itself last)
    int nv = 0;
    for (int c = lane; c < a.cols; c += 32) nv += a.NN[t * a.cols + c] >= 0;
    for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
    const int b = PRED ? nv + 1 : nv;
    int64_t src = -1;   // training index of this lane's point (-1: the test point / unused lane)
    if (PRED) {
        if (lane < nv) src = a.NN[t * a.cols + lane];
    } else {
        if (lane < nv) src = a.NN[t * a.cols + (nv - 1 - lane)];
    }
    for (int p0 = 0; p0 < b * D; p0 += 32) {   // uniform trip count: the shuffle needs the whole warp
        const int p = p0 + lane;
        const int r = min(p / D, b - 1), k = p - (p / D) * D;
        const int64_t sr = __shfl_sync(0xffffffffu, src, r);
        if (p < b * D) {
            const double v = sr >= 0 ? a.X[sr * D + k] : a.xq[t * D + k];
            xl[r * Dp + k] = v / vk.len[k];
        }
    }
    double yacc = (src >= 0 && a.y) ? a.y[src] : 0.0;
    const double nug = src >= 0 ? a.nugget * (a.nugget_diag ? a.nugget_diag[src] : 1.0) : a.nugget;
    __syncwarp();
    // ---- row `lane` of K
    double arow[32];
    const double* xi = xl + (lane < b ? lane : 0) * Dp;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        double v = 0.0;
        if (k < b) {   // warp-uniform
            const double* xk = xl + k * Dp;
            if (vk.kind == DGPB_SEXP) {
                double dist = 0.0;
                for (int d = 0; d < D; ++d) {
                    const double df = xi[d] - xk[d];
                    dist += df * df;
                }
                v = exp_nonpos(-dist);
            } else {
                double coef = 1.0, sr = 0.0;
                for (int d = 0; d < D; ++d) {
                    const double r = fabs(xi[d] - xk[d]);
                    coef *= 1.0 + kSqrt5 * r + (5.0 / 3.0) * (r * r);
                    sr += r;
                }
                v = coef * exp_nonpos(-kSqrt5 * sr);
            }
            if (k == lane) v = 1.0 + nug;
        }
        arow[k] = v;
    }
    // ---- Cholesky + fused forward solve
    double lll = 1.0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (j < b) {   // warp-uniform
            const double pj = __shfl_sync(0xffffffffu, arow[j], j);
            const double rs = rsqrt(pj);
            const double lij = arow[j] * rs;          // l_ij for rows i >= j (sqrt(pj) on the diagonal)
            arow[j] = lij;
            double* cb = colb + 32 * (j & 1);
            cb[lane] = lij;                           // column j of L: one store, then broadcast loads
            const double wj = __shfl_sync(0xffffffffu, yacc, j) * rs;
            if (lane > j) yacc = fma(-lij, wj, yacc);
            else if (lane == j) yacc = wj;
            __syncwarp();
            if (j == b - 1) lll = cb[j];
#pragma unroll
            for (int k = j + 1; k < 32; ++k)
                if (k < b) arow[k] = fma(-lij, cb[k], arow[k]);
        }
    }
    const double wl = __shfl_sync(0xffffffffu, yacc, b - 1);
    if (lane == 0) {
        if (PRED) {
            a.out0[t] = -wl * lll;
            a.out1[t] = a.scale * lll * lll;
        } else {
            a.out0[t * 2 + 0] = wl * wl;
            a.out0[t * 2 + 1] = 2.0 * log(fabs(lll));
        }
    }
}

template <bool PRED>
static int small_launch(const VKern& vk, const VSmallArgs& a, cudaStream_t st) {
    const int W = 8;
    const size_t smem = (size_t)W * (32 * (vk.D | 1) + 64) * sizeof(double);
    static bool configured = false;
    if (!configured) {
        DGPB_CUDA_TRY(cudaFuncSetAttribute(vecchia_small_kernel<PRED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)((size_t)W * (32 * (kMaxDim | 1) + 64) * sizeof(double))));
        configured = true;
    }
    vecchia_small_kernel<PRED><<<(unsigned)cdiv(a.count, W), W * 32, smem, st>>>(vk, a);
    DGPB_LAUNCHED();
    return DGPB_OK;
}

// ------------------------------------------------------------------------------------------------
// gp_vecch for SEVERAL squared-exponential nodes that share inputs and neighbours and have ONE length-scale each
// (the first layer of a Vecchia DGP: every node sees the design matrix, vecchia.py:635-654 is called per node by
// kernel_class.py:586-625).  The block's raw squared distances are formed once per test point; each node then
// only exponentiates them with its own length-scale, factors and solves with its own outputs.  Same
// lane-per-row scheme as vecchia_small_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kMultiMax = 32;   // nodes per launch
struct VMultiArgs {
    int64_t M, n;
    int cols, D, B;
    const double* xq;     // M x D
    const double* X;      // n x D
    const double* Y;      // B x n  (row b = outputs of node b)
    const int64_t* NN;    // M x cols
    double inv_l2[kMultiMax], scale[kMultiMax], nugget[kMultiMax];
    double* mean;         // B x M
    double* var;          // B x M
};

__global__ void __launch_bounds__(256, 2) vecchia_multi_kernel(VMultiArgs a) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t t = (int64_t)blockIdx.x * W + w;
    if (t >= a.M) return;
    const int D = a.D, Dp = D | 1;
    double* xl = smem + (size_t)w * (32 * Dp + 32 * 33 + 64);
    double* ds = xl + 32 * Dp;      // raw squared distances of the block, row stride 33
    int nv = 0;
    for (int c = lane; c < a.cols; c += 32) nv += a.NN[t * a.cols + c] >= 0;
    for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
    const int b = nv + 1;
    const int64_t src = lane < nv ? a.NN[t * a.cols + lane] : -1;
    for (int p0 = 0; p0 < b * D; p0 += 32) {
        const int p = p0 + lane;
        const int r = min(p / D, b - 1), k = p - (p / D) * D;
        const int64_t sr = __shfl_sync(0xffffffffu, src, r);
        if (p < b * D) xl[r * Dp + k] = sr >= 0 ? a.X[sr * D + k] : a.xq[t * D + k];
    }
    __syncwarp();
    {
        const double* xi = xl + (lane < b ? lane : 0) * Dp;
        for (int k = 0; k < b; ++k) {
            const double* xk = xl + k * Dp;
            double dist = 0.0;
            for (int d = 0; d < D; ++d) {
                const double df = xi[d] - xk[d];
                dist += df * df;
            }
            ds[lane * 33 + k] = dist;
        }
    }
    __syncwarp();
    double* colb = ds + 32 * 33;   // [2][32]: the current column of L, read back as broadcasts
    for (int nb = 0; nb < a.B; ++nb) {
        const double il2 = a.inv_l2[nb], nug = a.nugget[nb];
        double yacc = src >= 0 ? a.Y[(int64_t)nb * a.n + src] : 0.0;
        double arow[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            double v = 0.0;
            if (k < b) v = (k == lane) ? 1.0 + nug : exp_nonpos(-ds[lane * 33 + k] * il2);
            arow[k] = v;
        }
        double lll = 1.0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (j < b) {
                const double pj = __shfl_sync(0xffffffffu, arow[j], j);
                const double rs = rsqrt(pj);
                const double lij = arow[j] * rs;
                arow[j] = lij;
                double* cb = colb + 32 * (j & 1);
                cb[lane] = lij;                       // column j of L: one store, then broadcast loads
                const double wj = __shfl_sync(0xffffffffu, yacc, j) * rs;
                if (lane > j) yacc = fma(-lij, wj, yacc);
                else if (lane == j) yacc = wj;
                __syncwarp();
                if (j == b - 1) lll = cb[j];
#pragma unroll
                for (int k = j + 1; k < 32; ++k)
                    if (k < b) arow[k] = fma(-lij, cb[k], arow[k]);
            }
        }
        const double wl = __shfl_sync(0xffffffffu, yacc, b - 1);
        if (lane == 0) {
            a.mean[(int64_t)nb * a.M + t] = -wl * lll;
            a.var[(int64_t)nb * a.M + t] = a.scale[nb] * lll * lll;
        }
    }
}

static int g_vecchia_small = 1;   // dgpb_tune("vecchia_small", 0): always use the shared-memory kernels
int vecchia_set_small(int on) {
    g_vecchia_small = on != 0;
    return DGPB_OK;
}

static int train_launch(const VKern& vk, const double* X, const double* y, const int64_t* NN, int64_t n, int64_t m1,
                        double nugget, const double* nugget_diag, int mode, int P, int nugget_est, double* vals,
                        double* Lout, cudaStream_t st) {
    DGPB_REQUIRE(m1 >= 1 && m1 <= kMaxBlock, "conditioning block larger than 64");
    if (mode == 0 && m1 <= 32 && g_vecchia_small) {
        VSmallArgs a;
        a.count = n;
        a.cols = (int)m1;
        a.xq = nullptr;
        a.X = X;
        a.y = y;
        a.NN = NN;
        a.nugget_diag = nugget_diag;
        a.scale = 1.0;
        a.nugget = nugget;
        a.out0 = vals;
        a.out1 = nullptr;
        return small_launch<false>(vk, a, st);
    }
    const int per_warp = (int)(m1 * (m1 + 1) / 2 + m1 * vk.D + 3 * m1 + (mode == 1 ? P * m1 : 0) + (m1 + 1) / 2 + 2);
    int W = 8;
    while (W > 1 && (size_t)W * per_warp * sizeof(double) > 200 * 1024) W >>= 1;
    size_t smem = (size_t)W * per_warp * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
        DGPB_CUDA_TRY(cudaFuncSetAttribute(vecchia_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        configured = 220 * 1024;
    }
    vecchia_train_kernel<<<(unsigned)cdiv(n, W), W * 32, smem, st>>>(vk, X, y, NN, n, (int)m1, nugget, nugget_diag, mode,
                                                                    P, nugget_est, vals, Lout, per_warp);
    DGPB_LAUNCHED();
    return DGPB_OK;
}

static int pred_launch(const VKern& vk, const VPredArgs& a, cudaStream_t st) {
    const int bmax = a.mp + 1;
    DGPB_REQUIRE(bmax <= kMaxBlock, "prediction conditioning set larger than 63");
    if (a.mode == 0 && bmax <= 32 && g_vecchia_small) {
        VSmallArgs sa;
        sa.count = a.M;
        sa.cols = a.mp;
        sa.xq = a.xq;
        sa.X = a.w1;
        sa.y = a.y;
        sa.NN = a.NN;
        sa.nugget_diag = a.nugget_diag;
        sa.scale = a.scale;
        sa.nugget = a.nugget;
        sa.out0 = a.mean;
        sa.out1 = a.var;
        return small_launch<true>(vk, sa, st);
    }
    const int tri_sz = bmax * (bmax + 1) / 2;
    const int per_warp = tri_sz + bmax * vk.D + 2 * bmax + (a.mode ? 2 * tri_sz + 2 * bmax : 0) + (bmax + 1) / 2 + 2;
    int W = 8;
    while (W > 1 && (size_t)W * per_warp * sizeof(double) > 200 * 1024) W >>= 1;
    size_t smem = (size_t)W * per_warp * sizeof(double);
    static bool configured = false;
    if (!configured) {
        DGPB_CUDA_TRY(cudaFuncSetAttribute(vecchia_pred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        configured = true;
    }
    vecchia_pred_kernel<<<(unsigned)cdiv(a.M, W), W * 32, smem, st>>>(vk, a, per_warp);
    DGPB_LAUNCHED();
    return DGPB_OK;
}

// shared by the C entry points and ess.cu
int vecchia_llik_device(Workspace* ws, const VKern& vk, const double* X, const double* y, const int64_t* NN, int64_t n,
                        int64_t m1, double nugget, const double* nugget_diag, double* out2_dev, cudaStream_t st) {
    void* vals;
    DGPB_TRY(ws->reserve(SLOT_MISC, sizeof(double) * (size_t)n * 2, &vals));
    DGPB_TRY(train_launch(vk, X, y, NN, n, m1, nugget, nugget_diag, 0, 0, 0, (double*)vals, nullptr, st));
    colsum_kernel<<<2, 1024, 0, st>>>((double*)vals, n, 2, out2_dev);
    DGPB_LAUNCHED();
    return DGPB_OK;
}

// ------------------------------------------------------------------------------------------------
// Exact conditional draw of the MEAN process under a heteroskedastic Gaussian likelihood with the Vecchia
// approximation (latent Vecchia: imputation.py:141-158, U_matrix / U_matrix_sp vecchia.py:426-445,612-622,
// Hetero.post_het_vecch likelihood_class.py:165-183).  2n variables: y_j (index j < n, observed, noise Gamma_j)
// and f_j (index j + n, latent).  Row i of imp_NN = [i + n, i, neighbours]: f_i conditioned on y_i, on the latent
// values of its neighbours that come EARLIER in the ordering (index + n) and on the observations of those that come
// later.  Per row: K = scale corr(x) (nugget 0) + diag(Gamma on the observed entries + 1e-10), u = chol(K)^-T e_last.
//   U[i][c] (entry of imp_NN[i][c]) is written in imp_NN's own order; the latent entries of row i form row i of the
//   lower-triangular L = U_latent^T (diagonal = c 0), compacted into (depL, depNN, cnt) for the sparse solve;
//   rhs[i] = sd[i] - sum over observed entries U[i][c] y[imp_NN[i][c]]      (one solve gives mu + sample)
// One warp per row.
// ------------------------------------------------------------------------------------------------
__global__ void hetero_u_kernel(VKern vk, const double* __restrict__ X, const int64_t* __restrict__ NN, int64_t n, int m1,
                                double scale, const double* __restrict__ gamma, const double* __restrict__ y,
                                const double* __restrict__ sd, double* __restrict__ Uout, double* __restrict__ depL,
                                int64_t* __restrict__ depNN, int* __restrict__ cnt, double* __restrict__ rhs, int per_warp) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * W + w;
    if (i >= n) return;
    double* base = smem + (size_t)w * per_warp;
    double* A = base;                                  // packed K -> L       m1(m1+1)/2
    double* xl = A + m1 * (m1 + 1) / 2;                // scaled coords       m1 * D
    double* nug = xl + m1 * vk.D;                      // diagonal additions  m1
    double* bv = nug + m1;                             // L^-T e_last         m1
    int* idx = reinterpret_cast<int*>(bv + m1);        // entries, reversed: the latent f_i itself last
    int b = 0;
    for (int c = 0; c < m1; ++c) b += NN[i * m1 + c] >= 0;
    for (int c = lane; c < b; c += 32) idx[c] = (int)NN[i * m1 + (b - 1 - c)];
    __syncwarp();
    for (int p = lane; p < b * vk.D; p += 32) {
        const int r = p / vk.D, k = p % vk.D;
        xl[p] = X[(int64_t)(idx[r] % n) * vk.D + k] / vk.len[k];
    }
    for (int c = lane; c < b; c += 32) {
        const bool latent = idx[c] >= n;
        nug[c] = ((latent ? 0.0 : gamma[idx[c]]) + 1e-10) / scale;    // K / scale = corr + diag(...) / scale
        bv[c] = (c == b - 1) ? 1.0 : 0.0;
    }
    __syncwarp();
    warp_build_K(vk, xl, b, nug, A, lane);
    warp_chol(A, b, lane);
    warp_bwd(A, bv, b, lane);
    const double rs = 1.0 / sqrt(scale);                // chol(scale M) = sqrt(scale) chol(M)
    if (Uout)
        for (int c = lane; c < m1; c += 32) Uout[i * m1 + c] = c < b ? bv[b - 1 - c] * rs : 0.0;
    if (lane == 0) {
        int k = 0;
        double acc = sd[i];
        for (int c = 0; c < b; ++c) {                  // c in imp_NN order: entry c lives at bv[b - 1 - c]
            const int e = idx[b - 1 - c];
            const double u = bv[b - 1 - c] * rs;
            if (e >= n) {
                depL[i * m1 + k] = u;
                depNN[i * m1 + k] = e - n;
                ++k;
            } else {
                acc -= u * y[e];
            }
        }
        cnt[i] = k;
        rhs[i] = acc;
    }
}

// sp_solve_kernel with an explicit number of entries per row (entry 0 = the diagonal)
__global__ void __launch_bounds__(256) sp_solve_cnt_kernel(const double* __restrict__ L, const int64_t* __restrict__ NN,
                                                           const int* __restrict__ cnt, int64_t n, int m1,
                                                           const double* __restrict__ z, double* x, unsigned int* ticket) {
    __shared__ unsigned int s_blk;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    if (threadIdx.x == 0) s_blk = atomicAdd(ticket, 1u);
    __syncthreads();
    const int64_t i = (int64_t)s_blk * W + w;
    if (i >= n) return;
    const int kmax = cnt[i];
    double s = 0.0;
    for (int j = 1 + lane; j < kmax; j += 32) {
        const int64_t dep = NN[i * m1 + j];
        const double lij = L[i * m1 + j];
        double xd = NAN;
        if (dep >= 0 && dep < i) {   // a dependency is always an earlier row; anything else poisons instead of hanging
            const volatile unsigned long long* px = reinterpret_cast<const volatile unsigned long long*>(x + dep);
            unsigned long long bits = *px;
            int spins = 0;
            while (bits == kSpSentinel && ++spins < (1 << 22)) {
                __nanosleep(40);
                bits = *px;
            }
            if (bits != kSpSentinel) xd = __longlong_as_double((long long)bits);
        }
        s += lij * xd;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        double r = (z[i] - s) / L[i * m1];
        unsigned long long rb = (unsigned long long)__double_as_longlong(r);
        if (rb == kSpSentinel) rb = 0x7ff8000000000000ULL;
        *reinterpret_cast<volatile unsigned long long*>(x + i) = rb;
    }
}

int vecchia_mvn_draw_device(Workspace* ws, const VKern& vk, const double* X, const int64_t* NN, int64_t n, int64_t m1,
                            double scale, double nugget, const double* z, double* out, cudaStream_t st) {
    void *Lm, *flags;
    DGPB_TRY(ws->reserve(SLOT_VL, sizeof(double) * (size_t)n * m1, &Lm));
    DGPB_TRY(ws->reserve(SLOT_VFLAG, sizeof(int) * 8, &flags));
    DGPB_TRY(train_launch(vk, X, nullptr, NN, n, m1, nugget, nullptr, 2, 0, 0, nullptr, (double*)Lm, st));
    unsigned int* ticket = (unsigned int*)flags;
    sp_fill_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(reinterpret_cast<unsigned long long*>(out), n, ticket);
    DGPB_LAUNCHED();
    const int W = 8;
    sp_solve_kernel<<<(unsigned)cdiv(n, W), W * 32, 0, st>>>((double*)Lm, NN, n, (int)m1, 1.0 / sqrt(scale), z, out, ticket);
    DGPB_LAUNCHED();
    return DGPB_OK;
}

}  // namespace dgpb

using namespace dgpb;

// a private workspace for the entry points that do not take one (scratch for per-point values)
static thread_local dgpb_ws* tls_ws = nullptr;
static int get_tls_ws(dgpb_ws** out) {
    if (!tls_ws) {
        int dev = 0;
        DGPB_CUDA_TRY(cudaGetDevice(&dev));
        DGPB_TRY(dgpb_ws_create(&tls_ws, dev));
    }
    *out = tls_ws;
    return DGPB_OK;
}

extern "C" {

int dgpb_vecchia_llik(const double* X, const double* y, const int64_t* NN, int64_t n, int64_t D, int64_t m1,
                      const double* length_host, int64_t nlen, double scale, double nugget, const double* nugget_diag,
                      int kind, double* out_host, void* stream) {
    DGPB_NVTX("dgpb:vecchia_llik");
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(X && y && NN && out_host, "NULL argument");
    VKern vk;
    DGPB_TRY(make_vkern(kind, D, length_host, nlen, &vk));
    dgpb_ws* ws;
    DGPB_TRY(get_tls_ws(&ws));
    void* out;
    DGPB_TRY(ws->reserve(SLOT_OUT, sizeof(double) * kOutDoubles, &out));
    DGPB_TRY(vecchia_llik_device(ws, vk, X, y, NN, n, m1, nugget, nugget_diag, (double*)out, st));
    DGPB_CUDA_TRY(cudaMemcpyAsync(ws->pinned, out, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    DGPB_CUDA_TRY(cudaStreamSynchronize(st));
    const double quad = ws->pinned[0], logdet = ws->pinned[1];
    if (!(quad == quad) || !(logdet == logdet)) {
        set_error("Vecchia block is not positive definite");
        return DGPB_NOT_PD;
    }
    out_host[0] = -0.5 * (logdet + quad / scale);  // vecchia.py:179
    return DGPB_OK;
}

int dgpb_vecchia_nllik(const double* X, const double* y, const int64_t* NN, int64_t n, int64_t D, int64_t m1,
                       const double* length_host, int64_t nlen, double scale, double nugget, const double* nugget_diag,
                       int kind, int scale_est, int nugget_est, double* out_host, void* stream) {
    DGPB_NVTX("dgpb:vecchia_nllik");
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(X && y && NN && out_host, "NULL argument");
    VKern vk;
    DGPB_TRY(make_vkern(kind, D, length_host, nlen, &vk));
    const int P = (int)nlen + (nugget_est ? 1 : 0);
    const int cols = 2 * P + 2;
    dgpb_ws* ws;
    DGPB_TRY(get_tls_ws(&ws));
    void *vals, *out;
    DGPB_TRY(ws->reserve(SLOT_MISC, sizeof(double) * (size_t)n * cols, &vals));
    DGPB_TRY(ws->reserve(SLOT_OUT, sizeof(double) * kOutDoubles, &out));
    DGPB_TRY(train_launch(vk, X, y, NN, n, m1, nugget, nugget_diag, 1, P, nugget_est, (double*)vals, nullptr, st));
    colsum_kernel<<<cols, 1024, 0, st>>>((double*)vals, n, cols, (double*)out);
    DGPB_LAUNCHED();
    DGPB_CUDA_TRY(cudaMemcpyAsync(ws->pinned, out, cols * sizeof(double), cudaMemcpyDeviceToHost, st));
    DGPB_CUDA_TRY(cudaStreamSynchronize(st));
    const double quad = ws->pinned[0], logdet = ws->pinned[1];
    if (!(quad == quad) || !(logdet == logdet)) {
        set_error("Vecchia block is not positive definite");
        return DGPB_NOT_PD;
    }
    // vecchia.py:224-238 (origin_n == n)
    double s2 = scale_est ? quad / (double)n : scale;
    out_host[0] = scale_est ? 0.5 * (logdet + (double)n * log(s2)) : 0.5 * (logdet + quad / s2);
    out_host[1] = s2;
    for (int p = 0; p < P; ++p) out_host[2 + p] = 0.5 * (ws->pinned[2 + P + p] - ws->pinned[2 + p] / s2);
    return DGPB_OK;
}

int dgpb_vecchia_Lmatrix(const double* X, const int64_t* NN, int64_t n, int64_t D, int64_t m1, const double* length_host,
                         int64_t nlen, double nugget, int kind, double* L, void* stream) {
    DGPB_NVTX("dgpb:vecchia_Lmatrix");
    DGPB_REQUIRE(X && NN && L, "NULL argument");
    VKern vk;
    DGPB_TRY(make_vkern(kind, D, length_host, nlen, &vk));
    return train_launch(vk, X, nullptr, NN, n, m1, nugget, nullptr, 2, 0, 0, nullptr, L, (cudaStream_t)stream);
}

int dgpb_vecchia_mvn_draw(const double* X, const int64_t* NN, int64_t n, int64_t D, int64_t m1,
                          const double* length_host, int64_t nlen, double scale, double nugget, int kind,
                          const double* z, double* out, void* stream) {
    DGPB_NVTX("dgpb:vecchia_mvn_draw");
    DGPB_REQUIRE(X && NN && z && out, "NULL argument");
    VKern vk;
    DGPB_TRY(make_vkern(kind, D, length_host, nlen, &vk));
    dgpb_ws* ws;
    DGPB_TRY(get_tls_ws(&ws));
    return vecchia_mvn_draw_device(ws, vk, X, NN, n, m1, scale, nugget, z, out, (cudaStream_t)stream);
}

int dgpb_hetero_vecchia_draw(const double* X, const int64_t* imp_NN, int64_t n, int64_t D, int64_t m1,
                             const double* length_host, int64_t nlen, double scale, int kind, const double* gamma,
                             const double* y, const double* sd, double* f_out, double* U_out, void* stream) {
    DGPB_NVTX("dgpb:hetero_vecchia_draw");
    cudaStream_t st = (cudaStream_t)stream;
    DGPB_REQUIRE(X && imp_NN && gamma && y && sd && f_out, "NULL argument");
    DGPB_REQUIRE(n >= 1 && m1 >= 2 && m1 <= 128 && scale > 0.0, "bad sizes");
    VKern vk;
    DGPB_TRY(make_vkern(kind, D, length_host, nlen, &vk));
    dgpb_ws* ws;
    DGPB_TRY(get_tls_ws(&ws));
    void *pL, *pN, *pmisc, *pflag;
    DGPB_TRY(ws->reserve(SLOT_VL, sizeof(double) * (size_t)n * m1, &pL));
    DGPB_TRY(ws->reserve(SLOT_MISC, sizeof(int64_t) * (size_t)n * m1, &pN));
    DGPB_TRY(ws->reserve(SLOT_MISC2, (sizeof(double) + sizeof(int)) * (size_t)n + 64, &pmisc));
    DGPB_TRY(ws->reserve(SLOT_VFLAG, sizeof(int) * 8, &pflag));
    double* rhs = (double*)pmisc;
    int* cnt = reinterpret_cast<int*>(rhs + n);
    const int W = 4;
    const int per_warp = (int)(m1 * (m1 + 1) / 2 + m1 * vk.D + 2 * m1 + (m1 + 1) / 2 + 2);
    const size_t smem = sizeof(double) * (size_t)per_warp * W;
    DGPB_REQUIRE(smem <= 200 * 1024, "conditioning set too large for the block kernel");
    static size_t configured = 0;
    if (smem > configured) {
        DGPB_CUDA_TRY(cudaFuncSetAttribute(hetero_u_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    hetero_u_kernel<<<(unsigned)cdiv(n, W), W * 32, smem, st>>>(vk, X, imp_NN, n, (int)m1, scale, gamma, y, sd, U_out,
                                                               (double*)pL, (int64_t*)pN, cnt, rhs, per_warp);
    DGPB_LAUNCHED();
    unsigned int* ticket = (unsigned int*)pflag;
    sp_fill_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(reinterpret_cast<unsigned long long*>(f_out), n, ticket);
    DGPB_LAUNCHED();
    sp_solve_cnt_kernel<<<(unsigned)cdiv(n, 8), 8 * 32, 0, st>>>((double*)pL, (int64_t*)pN, cnt, n, (int)m1, rhs, f_out,
                                                                ticket);
    DGPB_LAUNCHED();
    return DGPB_OK;
}

int dgpb_gp_vecch(const double* x, int64_t M, const double* w, const double* y, int64_t n, int64_t D, const int64_t* NN,
                  int64_t mp, const double* length_host, int64_t nlen, double scale, double nugget,
                  const double* nugget_diag, int kind, double* mean, double* var, void* stream) {
    DGPB_NVTX("dgpb:gp_vecch");
    DGPB_REQUIRE(x && w && y && NN && mean && var, "NULL argument");
    if (M == 0) return DGPB_OK;
    VKern vk;
    DGPB_TRY(make_vkern(kind, D, length_host, nlen, &vk));
    VPredArgs a{};
    a.mode = 0;
    a.M = M;
    a.n = n;
    a.Dw = (int)D;
    a.Dz = 0;
    a.mp = (int)mp;
    a.xq = x;
    a.w1 = w;
    a.y = y;
    a.NN = NN;
    a.nugget_diag = nugget_diag;
    a.scale = scale;
    a.nugget = nugget;
    a.mean = mean;
    a.var = var;
    return pred_launch(vk, a, (cudaStream_t)stream);
}

int dgpb_gp_vecch_multi(const double* x, int64_t M, const double* w, const double* Y, int64_t n, int64_t D,
                        const int64_t* NN, int64_t mp, int B, const double* length_host, const double* scale_host,
                        const double* nugget_host, double* mean, double* var, void* stream) {
    DGPB_NVTX("dgpb:gp_vecch_multi");
    DGPB_REQUIRE(x && w && Y && NN && mean && var && length_host && scale_host && nugget_host, "NULL argument");
    DGPB_REQUIRE(B >= 1 && B <= kMultiMax, "number of nodes out of range");
    DGPB_REQUIRE(D >= 1 && D <= kMaxDim && mp >= 1 && mp + 1 <= 32, "block too large for the multi-node kernel");
    if (M == 0) return DGPB_OK;
    VMultiArgs a;
    a.M = M;
    a.n = n;
    a.cols = (int)mp;
    a.D = (int)D;
    a.B = B;
    a.xq = x;
    a.X = w;
    a.Y = Y;
    a.NN = NN;
    for (int b = 0; b < B; ++b) {
        a.inv_l2[b] = 1.0 / (length_host[b] * length_host[b]);
        a.scale[b] = scale_host[b];
        a.nugget[b] = nugget_host[b];
    }
    a.mean = mean;
    a.var = var;
    const int W = 8;
    const size_t smem = (size_t)W * (32 * ((int)D | 1) + 32 * 33 + 64) * sizeof(double);
    static bool configured = false;
    if (!configured) {
        DGPB_CUDA_TRY(cudaFuncSetAttribute(vecchia_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)((size_t)W * (32 * (kMaxDim | 1) + 32 * 33 + 64) * sizeof(double))));
        configured = true;
    }
    vecchia_multi_kernel<<<(unsigned)cdiv(M, W), W * 32, smem, (cudaStream_t)stream>>>(a);
    DGPB_LAUNCHED();
    return DGPB_OK;
}

int dgpb_linkgp_vecch(const double* m_in, const double* v_in, const double* z, int64_t M, const double* w1,
                      const double* gw, const double* y, int64_t n, int64_t Dw, int64_t Dz, const int64_t* NN,
                      int64_t mp, const double* length_host, int64_t nlen, double scale, double nugget,
                      const double* nugget_diag, int kind, double* mean, double* var, void* stream) {
    DGPB_NVTX("dgpb:linkgp_vecch");
    DGPB_REQUIRE(m_in && v_in && w1 && y && NN && mean && var, "NULL argument");
    DGPB_REQUIRE(Dz == 0 || (z && gw), "z/gw required when Dz > 0");
    if (M == 0) return DGPB_OK;
    VKern vk;
    DGPB_TRY(make_vkern(kind, Dw + Dz, length_host, nlen, &vk));
    VPredArgs a{};
    a.mode = 1;
    a.M = M;
    a.n = n;
    a.Dw = (int)Dw;
    a.Dz = (int)Dz;
    a.mp = (int)mp;
    a.xq = m_in;
    a.vq = v_in;
    a.zq = z;
    a.w1 = w1;
    a.gw = gw;
    a.y = y;
    a.NN = NN;
    a.nugget_diag = nugget_diag;
    a.scale = scale;
    a.nugget = nugget;
    a.mean = mean;
    a.var = var;
    return pred_launch(vk, a, (cudaStream_t)stream);
}

}  // extern "C"

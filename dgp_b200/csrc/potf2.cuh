// potf2.cuh -- factorisation and inversion of one 64 x 64 diagonal block (the serial step of the blocked
// Cholesky in dense.cu).  Kept in a header so scripts/potf2_probe.cu can time its phases in isolation.
#pragma once
#include "dense.cuh"

namespace dgpb {

#ifndef DGPB_POTF2_VARIANT
#define DGPB_POTF2_VARIANT 0   // timing experiments of scripts/potf2_probe.cu; 0 = production
#endif

constexpr size_t kPotf2Smem = (size_t)(2 * 64 * LDS + 32 * 33 + 16) * sizeof(double);

__device__ __forceinline__ void potf2_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// sD[r0.., c0..] (s x s) = -D[r0.., r0..] * ( L[r0.., c0..] * D[c0.., c0..] ),  D lower-triangular blocks of sD.
// Two products on the FP64 tensor path: 8 x 8 output tiles dealt round-robin to the `nw` warps of the group
// (`w` = this warp's index in it), K steps outside the triangles skipped.  All threads of the CTA must call
// (two __syncthreads inside); sT is this group's scratch (s x 33).
__device__ __forceinline__ void tri_inv_offdiag(const double* sL, double* sD, double* sT, int r0, int c0, int s, int w,
                                                int nw, int lane) {
    const int g = lane >> 2, t4 = lane & 3;
    const int nt = s / 8;
    // T = L21 * D11 : T[a][b] = sum_{k >= b} L[r0+a][c0+k] D[c0+k][c0+b]
    for (int tile = w; tile < nt * nt; tile += nw) {
        const int a0 = (tile / nt) * 8, b0 = (tile % nt) * 8;
        double c0v = 0.0, c1v = 0.0;
        for (int k = b0; k < s; k += 4)
            potf2_dmma(c0v, c1v, sL[(r0 + a0 + g) * LDS + c0 + k + t4], sD[(c0 + k + t4) * LDS + c0 + b0 + g]);
        sT[(a0 + g) * 33 + b0 + 2 * t4] = c0v;
        sT[(a0 + g) * 33 + b0 + 2 * t4 + 1] = c1v;
    }
    __syncthreads();
    // X = -D22 * T : X[a][b] = -sum_{k <= a} D[r0+a][r0+k] T[k][b]
    for (int tile = w; tile < nt * nt; tile += nw) {
        const int a0 = (tile / nt) * 8, b0 = (tile % nt) * 8;
        double c0v = 0.0, c1v = 0.0;
        for (int k = 0; k < a0 + 8; k += 4)
            potf2_dmma(c0v, c1v, -sD[(r0 + a0 + g) * LDS + r0 + k + t4], sT[(k + t4) * 33 + b0 + g]);
        sD[(r0 + a0 + g) * LDS + c0 + b0 + 2 * t4] = c0v;
        sD[(r0 + a0 + g) * LDS + c0 + b0 + 2 * t4 + 1] = c1v;
    }
    __syncthreads();
}

// (a) diagonal block: ONE CTA per matrix.
//   Register layout: the 136 lower-triangular 4 x 4 blocks of the 64 x 64 matrix, one per thread, enumerated
//   from the LAST block row backwards so that the partial fifth warp holds the rows that finish first (threads
//   136..255 only help with loads, the inverse and the write-out).  A step eliminates TWO columns (j, j+1):
//   the owners publish their raw entries as double2 {a_ij, a_i,j+1}, one barrier, then every thread applies the
//   rank-2 update  a_ik -= [c_i.x c_i.y] P^-1 [c_k.x c_k.y]'  (P = the 2 x 2 pivot block) with 8 shared-memory
//   loads (4 rows + 4 columns), 16 multiplies for w_k = P^-1 c_k and 32 FMAs -- no masks: entries of finished
//   rows or columns receive garbage that nothing reads.
//   P^-1 and the scalings of step s+1 are computed by the ONE thread that owns the next pivot's diagonal block,
//   right after its own update of step s and before the barrier (the pivot entries live in its registers), so
//   the other warps never execute the reciprocal-square-root chain; rsqrt(d0) and rsqrt(det P) are independent
//   (d1 = det P / d0).  Measured on B200: DFMA 8 cycles dependent / 2 cycles issue per warp, rsqrt 75 cycles,
//   STS + barrier + LDS 65 cycles -- the step is bounded by that chain, not by throughput.
//   Then inv(L_kk): the four 16 x 16 diagonal blocks by forward substitution with one warp per block (column
//   per lane, solution kept in registers), and two block levels.  Results go to the side buffer.
// `stamps` (probe only, NULL in production): clock64 at the phase boundaries, 8 per CTA.
__global__ void __launch_bounds__(256, 2) potf2_kernel(Batch bt, int64_t ld, int npad, int k0, long long* stamps,
                                                       int early = 0) {
#define DGPB_STAMP(i) do { if (stamps && threadIdx.x == 0) stamps[blockIdx.x * 8 + (i)] = clock64(); } while (0)
#if DGPB_POTF2_VARIANT == 9   // fine-grained stamps inside one loop iteration (thread 0 = block (15,15))
#define DGPB_FSTAMP(i, dep) do { if (stamps && threadIdx.x == 0 && jj == 9) { long long t_; \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) : "d"(dep) : "memory"); stamps[64 + (i)] = t_; } } while (0)
#else
#define DGPB_FSTAMP(i, dep) do { } while (0)
#endif
    DGPB_STAMP(0);
    extern __shared__ double smem[];
    double* sL = smem;
    double* sD = sL + 64 * LDS;
    double* sT = sD + 64 * LDS;
    __shared__ double2 col[2][64];   // raw columns (j, j+1) of the current step; row i at [(i & 3) * 16 + (i >> 2)]:
                                     // lanes holding consecutive blocks read consecutive 16-byte words
    __shared__ double2 piv[2][4];    // {m00, m01}, {m11, rs0}, {rs1, l10}, {L_jj, L_j+1,j+1} of the step
    __shared__ double rdiag[64];     // 1 / L_jj
    const int tid = threadIdx.x;
    const double* __restrict__ T = bt.T[blockIdx.x];
    int bi = 0, bk = 0;
    const bool has_block = tid < 136;
    if (has_block) tri_index(135 - tid, bi, bk);
    // largest row held by this warp (its first thread has the largest block index): warp-uniform liveness test
    int warp_row_hi = -1;
    if ((tid & ~31) < 136) {
        int bi2, bk2;
        tri_index(135 - (tid & ~31), bi2, bk2);
        warp_row_hi = 4 * bi2 + 3;
    }
    pdl_wait();
    if (early) pdl_trigger();
    double a[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = 4 * bi + r, k = 4 * bk + c;
            a[r][c] = (has_block && k <= i) ? T[(int64_t)(k0 + i) * ld + k0 + k] : 0.0;
        }
    for (int idx = tid; idx < 64 * LDS; idx += 256) {
        sL[idx] = 0.0;
        sD[idx] = 0.0;
    }
    // pivot block {a00, a10, a11} of columns (j, j+1) -> piv[buf]
    auto publish_pivot = [&](int buf, double a00, double a10, double a11, int j) {
        const double d0 = a00;
        const double det = fma(a11, d0, -(a10 * a10));       // d0 * d1
        const double rs0 = rsqrt(d0);
        const double rsd = rsqrt(det);
        const double r0 = rs0 * rs0;                          // 1 / d0
        const double s0 = d0 * rs0;                           // sqrt(d0) = L_jj
        const double rs1 = rsd * s0;                          // 1 / sqrt(d1)
        const double r1 = rs1 * rs1;                          // 1 / d1
        const double t = a10 * r0;
        const double m01 = -(t * r1);
        piv[buf][0] = make_double2(fma(-t, m01, r0), m01);
        piv[buf][1] = make_double2(r1, rs0);
        piv[buf][2] = make_double2(rs1, a10 * rs0);
        piv[buf][3] = make_double2(s0, det * r0 * rs1);
        if (!(d0 > 0.0) || !(det > 0.0)) atomicCAS(&bt.info[blockIdx.x], 0, k0 + j + (d0 > 0.0 ? 2 : 1));
    };
    if (has_block && bi == 0 && bk == 0) publish_pivot(0, a[0][0], a[1][0], a[1][1], 0);
    DGPB_STAMP(1);
#pragma unroll 1
    for (int jj = 0; jj < 32; ++jj) {
        const int j = 2 * jj, bj = jj >> 1, buf = jj & 1;
        const bool odd = (jj & 1) != 0;  // columns (2, 3) of the block instead of (0, 1)
        const bool warp_live = warp_row_hi >= j;
        const bool owner = has_block && bk == bj;
        DGPB_FSTAMP(0, a[3][3]);
        if (owner) {
#pragma unroll
            for (int r = 0; r < 4; ++r) col[buf][16 * r + bi] = make_double2(odd ? a[r][2] : a[r][0], odd ? a[r][3] : a[r][1]);
        }
        DGPB_FSTAMP(1, a[3][3]);
        __syncthreads();
        DGPB_FSTAMP(2, a[3][3]);
        if (!warp_live) continue;
        double2 ci[4], ck[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) ci[r] = col[buf][16 * r + bi];
#pragma unroll
        for (int c = 0; c < 4; ++c) ck[c] = col[buf][16 * c + bk];
        const double2 pm = piv[buf][0], pn = piv[buf][1];
        const double m00 = pm.x, m01 = pm.y, m11 = pn.x;
        DGPB_FSTAMP(3, ci[0].x + ci[3].y + ck[0].x + ck[3].y + pm.x + pn.y);
        double2 w[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            w[c].x = fma(m01, ck[c].y, m00 * ck[c].x);
            w[c].y = fma(m11, ck[c].y, m01 * ck[c].x);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) a[r][c] = fma(-ci[r].y, w[c].y, fma(-ci[r].x, w[c].x, a[r][c]));
        DGPB_FSTAMP(4, a[0][0] + a[1][1] + a[2][2] + a[3][3] + a[3][0]);
        // the owner of the next pivot's diagonal block prepares step jj + 1
        if (has_block && bi == bk && bi == ((jj + 1) >> 1) && jj + 1 < 32) {
            if (odd) publish_pivot(buf ^ 1, a[0][0], a[1][0], a[1][1], j + 2);   // next step: first pair of the next block
            else publish_pivot(buf ^ 1, a[2][2], a[3][2], a[3][3], j + 2);       // next step: second pair of this block
        }
        DGPB_FSTAMP(5, a[3][3]);
        if (owner) {
            // final entries of columns j, j+1 of L for this thread's rows
            const double rs0 = pn.y;
            const double2 po = piv[buf][2], ps = piv[buf][3];
            const double rs1 = po.x, l10 = po.y;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = 4 * bi + r;
                const double li0 = ci[r].x * rs0;
                const double li1 = (ci[r].y - li0 * l10) * rs1;
                if (i >= j) sL[i * LDS + j] = (i == j) ? ps.x : li0;
                if (i >= j + 1) sL[i * LDS + j + 1] = (i == j + 1) ? ps.y : li1;
            }
            if (bi == bj) {
                rdiag[j] = rs0;
                rdiag[j + 1] = rs1;
            }
        }
        DGPB_FSTAMP(6, a[3][3]);
    }
    __syncthreads();
    DGPB_STAMP(2);

    // inv(L_kk), 16 x 16 diagonal blocks: warp w < 4 inverts block w.  Lane c (< 16; lanes 16..31 mirror them)
    // solves L x = e_c by forward substitution with x in registers; x_k = 0 for k < c makes the sums mask-free.
    if (tid < 128) {
        const int o = (tid >> 5) * 16, c = tid & 15;
        double x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int k = 0; k + 1 < i; k += 2) {
                s0 = fma(sL[(o + i) * LDS + o + k], x[k], s0);
                s1 = fma(sL[(o + i) * LDS + o + k + 1], x[k + 1], s1);
            }
            if (i & 1) s0 = fma(sL[(o + i) * LDS + o + i - 1], x[i - 1], s0);
            const double ri = rdiag[o + i];
            x[i] = (i == c) ? ri : ((i < c) ? 0.0 : -(s0 + s1) * ri);
        }
        if ((tid & 31) < 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i >= c) sD[(o + i) * LDS + o + c] = x[i];
        }
    }
    __syncthreads();
    DGPB_STAMP(3);
    pdl_trigger();   // the panel-row solve may become resident while the off-diagonal inverse blocks finish
    {
        const int half = tid >> 7;  // warps 0-3: block (16..31, 0..15); warps 4-7: block (48..63, 32..47)
        tri_inv_offdiag(sL, sD, sT + half * 16 * 33, half ? 48 : 16, half ? 32 : 0, 16, (tid >> 5) & 3, 4, tid & 31);
    }
    DGPB_STAMP(4);
    tri_inv_offdiag(sL, sD, sT, 32, 0, 32, tid >> 5, 8, tid & 31);
    DGPB_STAMP(5);

    double* dg = bt.diag[blockIdx.x];
    double* blk = dg + npad + (size_t)(k0 / NB) * NB * NB;
    double* dinv = dg + (size_t)npad * (1 + NB);
    for (int idx = tid; idx < 64 * 32; idx += 256) {   // LDS = 68 and the side buffers keep every row 16-byte aligned
        const int r = idx >> 5, c2 = (idx & 31) * 2;
        reinterpret_cast<double2*>(blk)[idx] = *reinterpret_cast<const double2*>(&sL[r * LDS + c2]);
        reinterpret_cast<double2*>(dinv)[idx] = *reinterpret_cast<const double2*>(&sD[r * LDS + c2]);
    }
    if (tid < 64) dg[k0 + tid] = sL[tid * LDS + tid];
    DGPB_STAMP(6);
}

#undef DGPB_STAMP
#undef DGPB_FSTAMP

}  // namespace dgpb

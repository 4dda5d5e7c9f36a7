// knn_tc5.cuh -- the neighbour-search SCREEN on the 5th-generation tensor cores (tcgen05, sm_100a), included by knn.cu.
//
// What is screened: for query q and candidate x_j, with everything shifted by the first candidate x_0,
//     s(q, j) = |x_j - x_0|^2 - 2 (q - x_0).(x_j - x_0)  ( = |q - x_j|^2 - |q - x_0|^2 )
// in FP32; a per-query sorted list keeps the 32 / 64 smallest, which are then re-ranked with the reference's FP64
// arithmetic (knn_rank_and_write) under an error-bound check -- a query whose screen could have lost a neighbour is
// redone by the exact scalar kernel, so the indices are always those of the exact search.
//
// Why tcgen05: the warp-level TF32 path (knn_tf32_kernel) is bound by instruction issue, not by the tensor pipe
// (profiles/r1s2_knn_tf32_ncu_raw.txt: 165 warp instructions per 16 x 16 step, issue slots 51 % busy, tensor pipe
// 28 %): every mma.sync needs its fragments loaded and its C registers scanned by the warp that issued it.  Here
//   * the candidates are packed ONCE per search into TF32 "screen rows" (knn_tc5_pack_kernel): split operands
//     v = hi + lo as three K-blocks [b_hi | b_lo | b_hi] against query rows [a_hi | a_hi | a_lo] (so one MMA chain
//     forms hi.hi + hi.lo + lo.hi), and |x_j - x_0|^2 split three ways against three columns of ones -- the
//     accumulator IS s(q, j), no epilogue arithmetic;
//   * the packed rows are stored tile by tile in the canonical K-major no-swizzle UMMA layout, so a 128-candidate
//     tile is ONE contiguous block that the TMA engine copies with a single cp.async.bulk (mbarrier complete_tx);
//   * ONE thread issues tcgen05.mma.kind::tf32 (M = 128 queries, N = 128 candidates, K = 8 per instruction) for TWO
//     query groups per candidate tile into double-buffered TMEM accumulators (2 groups x 2 buffers x 128 columns = all
//     512 columns), tcgen05.commit releases the shared-memory stage and hands the accumulators to the epilogue;
//   * eight epilogue warps (two per scheduler) read their 32 TMEM lanes with tcgen05.ld (32 candidates per instruction
//     and thread): one THREAD owns one query, so a candidate costs one compare against the query's threshold, folded
//     into a 32-bit hit mask; the values are parked in shared memory and the rare hits go through the same
//     warp-cooperative sorted-list insertion as before.
#pragma once

namespace tc5 {

constexpr int kQG = 128;         // queries per group = TMEM lanes = UMMA M
constexpr int kGroups = 2;       // query groups per CTA: both multiply the SAME candidate tile (half the TMA traffic)
constexpr int kQ = kQG * kGroups;   // queries per CTA
constexpr int kN = 128;          // candidates per tile = UMMA N
constexpr int kMaxStages = 3;    // shared-memory stages of packed candidate tiles (2 when 3 do not fit)
constexpr int kMaxK8 = 12;       // K' = 8 * K8 <= 96 screen columns (D <= 31)
constexpr int kEpiWarps = 4 * kGroups;   // warp w: query group w / 4, TMEM lane quarter w % 4
constexpr int kThreads = 32 * (kEpiWarps + 2);   // + one TMA warp + one MMA / TMEM-allocation warp

// screen columns: [0, D) a_hi.b_hi, [D, 2D) a_hi.b_lo, [2D, 3D) a_lo.b_hi, [3D, 3D + 3) 1 . |x - x0|^2 (3-way split)
__host__ __device__ inline int k8_of(int D) { return (3 * D + 3 + 7) / 8; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_round(float v) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
// the three TF32 pieces of a double: v ~ p0 + p1 + p2 (33 significant bits)
__device__ __forceinline__ void tf32_split3(double v, float& p0, float& p1, float& p2) {
    p0 = tf32_round((float)v);
    const double r1 = v - (double)p0;
    p1 = tf32_round((float)r1);
    p2 = tf32_round((float)(r1 - (double)p1));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// the same for the single-thread roles (TMA producer, MMA issuer): they share their schedulers with epilogue warps, so a
// failed poll backs off instead of spinning in their issue slots
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, unsigned parity) {
    while (true) {
        unsigned ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        __nanosleep(64);
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA engine: one contiguous block global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// K-major, no swizzle: (8 rows x 16 B) core matrices contiguous, SBO between 8-row groups, LBO between the two
// 16-byte K chunks of one K = 8 (TF32) instruction   (cute/arch/mma_sm100_desc.hpp: SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version of sm_100
    return d;                 // base offset 0, layout type 0 = SWIZZLE_NONE
}
// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = 128   (InstrDescriptor, same header)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kQG >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(kIdesc), "r"((uint32_t)accumulate)
        : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier when they have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- per-query survivor set: the LC smallest screened values as a binary MAX-HEAP owned by ONE thread -----------------
// Layout [slot][query of the CTA]: the 32 threads of a warp touch 32 consecutive words per slot (no bank conflicts).
// The root is the query's threshold; a hit replaces the root and sifts down (<= log2 LC levels, no shuffles, no votes):
// every lane of a warp that holds a hit inserts at the same time.  The exact ranking that follows orders the set.
template <int LC>
__device__ __forceinline__ float heap_replace_root(float* K, int* I, int stride, float key, int idx) {
    int i = 0;
#pragma unroll 1
    while (true) {
        const int l = 2 * i + 1, r = l + 1;
        if (l >= LC) break;
        const float kl = K[l * stride];
        const float kr = r < LC ? K[r * stride] : -INFINITY;
        const bool right = kr > kl;
        const int ch = right ? r : l;
        const float kc = right ? kr : kl;
        if (!(kc > key)) break;
        K[i * stride] = kc;
        I[i * stride] = I[ch * stride];
        i = ch;
    }
    K[i * stride] = key;
    I[i * stride] = idx;
    return K[0];
}

// ---- pack: candidates -> TF32 screen rows, tile by tile in the UMMA layout ------------------------------------------
// packed[tile][chunk c = k / 4][row r < 128][k % 4], K' = 8 K8 columns: a tile is NCH x 2048 contiguous bytes.
__global__ void __launch_bounds__(kN) knn_tc5_pack_kernel(const double* __restrict__ x, int64_t n, int D, int K8,
                                                           float* __restrict__ packed, float* __restrict__ xnmax_out) {
    const int64_t j = (int64_t)blockIdx.x * kN + threadIdx.x;     // one thread per candidate row of the tile
    const int NCH = 2 * K8;
    float* tile = packed + (size_t)blockIdx.x * NCH * kN * 4;
    const int r = threadIdx.x;
    double sn = 0.0;
    for (int k = 0; k < 8 * K8; ++k) {
        float val = 0.f;
        if (j < n) {
            if (k < 3 * D) {
                const int d = k % D, blk = k / D;
                const double v = x[j * D + d] - x[d];
                const float hi = tf32_round((float)v);
                val = (blk == 1) ? tf32_round((float)(v - (double)hi)) : hi;
                if (blk == 0) sn += v * v;
            } else if (k < 3 * D + 3) {
                float p0, p1, p2;
                tf32_split3(sn, p0, p1, p2);
                val = k == 3 * D ? p0 : (k == 3 * D + 1 ? p1 : p2);
            }
        } else if (k == 3 * D) {
            val = 1e30f;    // padding rows can never enter a list
        }
        tile[((size_t)(k >> 2) * kN + r) * 4 + (k & 3)] = val;
    }
    // largest |x - x0|^2 (error bound of the screen)
    __shared__ float smax[kN / 32];
    float mx = j < n ? (float)sn : 0.f;
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < kN / 32; ++i) mx = fmaxf(mx, smax[i]);
        atomicMax(reinterpret_cast<int*>(xnmax_out), __float_as_int(mx));   // non-negative floats order like ints
    }
}

template <int NL>
constexpr size_t smem_bytes(int K8, int kStages) {
    return 1024 /* alignment slack */ + (size_t)2 * K8 * kQ * 16 + (size_t)kStages * 2 * K8 * kN * 16 +
           (size_t)kQ * 32 * NL * 8 + (size_t)kEpiWarps * 32 * 33 * 4 + 256;
}

template <int NL, bool ORDERED>
__global__ void __launch_bounds__(kThreads, 1) knn_tc5_kernel(const double* __restrict__ q, int64_t M,
                                                             const double* __restrict__ x, int64_t n, int D, int K8,
                                                             int kStages, int m, const float* __restrict__ packed,
                                                             const float* __restrict__ xnmax_in, int64_t* __restrict__ NN,
                                                             int ldnn, unsigned char* __restrict__ flags) {
    constexpr int LC = 32 * NL;
    extern __shared__ unsigned char raw_smem[];
    unsigned char* sm = reinterpret_cast<unsigned char*>(((uintptr_t)raw_smem + 1023) & ~(uintptr_t)1023);
    const int NCH = 2 * K8;                                        // 16-byte K chunks
    float* sA = reinterpret_cast<float*>(sm);                      // [group][NCH][128][4]
    float* sB = sA + (size_t)kGroups * NCH * kQG * 4;              // [stages][NCH][128][4]
    float* lkey = sB + (size_t)kStages * NCH * kN * 4;             // [LC][256] heap keys, slot-major
    int* lidx = reinterpret_cast<int*>(lkey + (size_t)kQ * LC);    // [LC][256]
    float* park = reinterpret_cast<float*>(lidx + (size_t)kQ * LC);   // [epilogue warp][32 lanes][33]
    uint64_t* bars = reinterpret_cast<uint64_t*>(park + (size_t)kEpiWarps * 32 * 33);
    uint64_t* full = bars;                  // [kStages] TMA -> MMA
    uint64_t* empty = bars + kMaxStages;    // [kStages] MMA -> TMA
    uint64_t* tfull = bars + 2 * kMaxStages;        // [2] MMA -> epilogue
    uint64_t* tempty = bars + 2 * kMaxStages + 2;   // [2] epilogue -> MMA
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * kQ;
    const int64_t cmax = ORDERED ? min(n, q0 + kQ) : n;
    const int ntiles = (int)((cmax + kN - 1) / kN);
    const unsigned tile_bytes = (unsigned)NCH * kN * 16;

    // ---- one-time setup
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], kEpiWarps);   // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (w == kEpiWarps + 1) {   // TMEM: all 512 columns = 2 groups x 2 buffers of 128 x 128 FP32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_base_slot)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    // query screen rows [a_hi | a_hi | a_lo | 1 1 1 | 0...], a = -2 (q - x0)
    for (int idx = tid; idx < kQ * 8 * K8; idx += kThreads) {
        const int r = idx / (8 * K8), k = idx - r * (8 * K8);
        const int64_t qi = q0 + r;
        float val = 0.f;
        if (qi < M) {
            if (k < 3 * D) {
                const int d = k % D, blk = k / D;
                const double v = -2.0 * (q[qi * D + d] - x[d]);
                const float hi = tf32_round((float)v);
                val = (blk == 2) ? tf32_round((float)(v - (double)hi)) : hi;
            } else if (k < 3 * D + 3) {
                val = 1.f;
            }
        }
        sA[(((size_t)(r / kQG) * NCH + (k >> 2)) * kQG + (r % kQG)) * 4 + (k & 3)] = val;
    }
    for (int i = tid; i < kQ * LC; i += kThreads) {
        lkey[i] = INFINITY;
        lidx[i] = -1;
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes of sA -> visible to the MMA
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_base_slot;

    if (w == kEpiWarps) {
        // ===== TMA producer: one bulk copy per 128-candidate tile
        if (lane == 0) {
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % kStages;
                if (t >= kStages) mbar_wait_relaxed(&empty[s], ((t / kStages) - 1) & 1);
                mbar_expect_tx(&full[s], tile_bytes);
                tma_bulk_load(sB + (size_t)s * NCH * kN * 4, packed + (size_t)t * NCH * kN * 4, tile_bytes, &full[s]);
            }
        }
    } else if (w == kEpiWarps + 1) {
        // ===== MMA issuer: one thread, K8 instructions per tile and query group into the TMEM buffers of the tile's parity
        if (lane == 0) {
            const uint32_t a_base = smem_u32(sA);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % kStages, b = t & 1;
                if (t >= 2) mbar_wait_relaxed(&tempty[b], ((t >> 1) - 1) & 1);   // the epilogue has drained these accumulators
                mbar_wait_relaxed(&full[s], (t / kStages) & 1);                  // the tile has landed
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t b_base = smem_u32(sB + (size_t)s * NCH * kN * 4);
                for (int g = 0; g < kGroups; ++g)
                    for (int k8 = 0; k8 < K8; ++k8) {
                        const uint64_t da = umma_desc(a_base + (uint32_t)((g * NCH + 2 * k8) * kQG * 16), kQG * 16, 128);
                        const uint64_t db = umma_desc(b_base + (uint32_t)(2 * k8) * kN * 16, kN * 16, 128);
                        umma_tf32(tmem_base + (uint32_t)((g * 2 + b) * kN), da, db, k8 > 0);
                    }
                umma_commit(&empty[s]);    // the stage can be refilled once these MMAs have read it
                umma_commit(&tfull[b]);    // ... and the accumulators are complete
            }
        }
    } else {
        // ===== epilogue: warp w = (group g, lane quarter lq) owns TMEM lanes [32 lq, 32 lq + 32) of the group's buffers
        const int g = w >> 2, lq = w & 3;
        const int ql = kQG * g + 32 * lq + lane;
        const int64_t qi = q0 + ql;
        float thr = qi < M ? INFINITY : -INFINITY;
        float* mypark = park + (size_t)w * 32 * 33;
        const int qbase = kQG * g + 32 * lq;     // first query (CTA-local) of this warp
        for (int t = 0; t < ntiles; ++t) {
            const int b = t & 1;
            mbar_wait(&tfull[b], (t >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const int64_t c0 = (int64_t)t * kN;
            const int64_t jl = (ORDERED ? min(n, qi) : n) - c0;
            const int lim = qi < M ? (int)max((int64_t)0, min(jl, (int64_t)kN)) : 0;   // admissible columns of this tile
            for (int ch = 0; ch < kN / 32; ++ch) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(32 * lq) << 16) + (uint32_t)((g * 2 + b) * kN + 32 * ch), v);
                unsigned mask = 0;
#pragma unroll
                for (int c = 0; c < 32; ++c) mask |= (v[c] < thr ? 1u : 0u) << c;
                const int rem = lim - 32 * ch;
                mask &= rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
                if (!__any_sync(0xffffffffu, mask != 0)) continue;
#pragma unroll
                for (int c = 0; c < 32; ++c) mypark[lane * 33 + c] = v[c];   // own row, stride 33: conflict-free
                while (mask) {   // every lane with hits inserts into ITS query's heap, all at the same time
                    const int c = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float e = mypark[lane * 33 + c];
                    if (e < thr) thr = heap_replace_root<LC>(lkey + ql, lidx + ql, kQ, e, (int)(c0 + 32 * ch + c));
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[b]);
        }
        // ---- exact FP64 ranking of the survivors, error-bound check (as the warp-level kernels)
        const float xnmax = *xnmax_in;
        for (int qs = qbase; qs < qbase + 32; ++qs) {
            const int64_t qq_i = q0 + qs;
            if (qq_i >= M) break;
            const float smaxf = lkey[qs];   // root of the heap = largest survivor (+inf while the set is not full)
            double qq = 0.0;
            for (int k = 0; k < D; ++k) {
                const double vv = q[qq_i * D + k] - x[k];
                qq += vv * vv;
            }
            const double rr = sqrt(qq) + sqrt((double)xnmax);
            const double eps = 7.62939453125e-06 * rr * rr;   // 2^-17 (|q| + |x|max)^2
            dgpb::knn_rank_and_write<NL, ORDERED>(q, x, qq_i, D, m, lidx + qs, (double)smaxf + qq - eps, NN, ldnn, flags,
                                                  lane, kQ);
        }
    }
    // ---- teardown
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (w == kEpiWarps + 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
}

}  // namespace tc5

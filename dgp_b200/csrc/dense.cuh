// dense.cuh -- internal interface of the dense FP64 factorisation pipeline (dense.cu).
#pragma once
#include "common.cuh"

namespace dgpb {

constexpr int NB = 64;    // panel width (columns factored per step)
constexpr int TM = 128;   // square tile of the trailing update
constexpr int LDS = 68;   // smem row stride (doubles) of 64-wide K-contiguous tiles; 68 % 16 == 4 makes the
                          // 8x4 FP64 mma fragment loads bank-conflict free
constexpr int MAXB = 32;  // max matrices per batched launch

// Geometry of one factorisation matrix T (R x ld, row-major, lower triangle used):
//   rows/cols [0,npad)          : K (padded with an identity block to a multiple of NB)
//   row/col   npad              : y' (the forward solve L^-1 y falls out of the factorisation)
//   rows/cols npad+1+i, i<npad  : e_i  (only when `aug`): partial Cholesky of the first npad columns leaves
//                                 the Schur complement -[y I]' K^-1 [y I] = -[[quad, a'],[a, K^-1]] there.
struct Geom {
    int n, npad, R;
    int64_t ld;
    bool aug;
    size_t elems() const { return (size_t)R * (size_t)ld; }
};
inline Geom make_geom(int64_t n, bool aug) {
    Geom g;
    g.n = (int)n;
    g.npad = (int)round_up(n, NB);
    g.aug = aug;
    g.R = aug ? 2 * g.npad + 1 : g.npad + 1;
    g.ld = round_up(g.R + 1, 8);
    return g;
}

struct Batch {
    double* T[MAXB];
    double* diag[MAXB];   // per matrix: npad diagonal entries of L, npad/NB diagonal blocks (NB x NB), then
                          // inv(L_kk) of the panel in flight (NB x NB)
    int* info;            // device int[B]: 0 or (failing column + 1)
};

// size in doubles of one `diag` side buffer
inline size_t diag_elems(const Geom& g) { return (size_t)g.npad * (1 + NB) + (size_t)NB * NB; }

// Assemble T for every batch entry: K (lower) from the kernel descriptor, y row, identity rows.
int assemble(const Geom& g, const KernelDev* kds, const double* const* ys, const Batch& bt, int B, cudaStream_t st);
// Sliding-window partial Cholesky of the first npad columns.  `ctx` (0/1) selects one of the calling thread's two
// look-ahead stream sets: two factorisations issued with different contexts run side by side on the GPU.
int factorize(const Geom& g, const Batch& bt, int B, cudaStream_t st, int ctx = 0);
// `st` waits for whatever ESS wave is still in flight in either factorisation context of the workspace (a wave
// assembled and factored ahead of an acceptance): every entry point that uses the T sets outside the wave pipeline
// calls this first.
int join_waves(Workspace* ws, cudaStream_t st);
// out[b*4+0] = 2 sum log diag(L) (first n), out[b*4+1] = |L^-1 y|^2, out[b*4+2] = sigma2 used
// (scale_est ? quad/n : scale_in[b]).
struct ScaleArgs {
    double scale[MAXB];
    int est[MAXB];
};
int reduce_logdet_quad(const Geom& g, const Batch& bt, int B, const ScaleArgs& sa, double* out, cudaStream_t st);
// copy the diagonal blocks kept in the side buffer back into T so that T holds the complete factor L
int restore_diag_blocks(const Geom& g, const Batch& bt, int B, cudaStream_t st);

// nu = sqrt_scale * L z with L the lower factor stored in T
int launch_trmv(const double* T, int64_t ld, int n, double sqrt_scale, const double* z, double* nu, cudaStream_t st);

// allocate/partition workspace buffers for a batch (tslot selects one of the two T sets)
int setup_batch(Workspace* ws, const Geom& g, int B, Batch* bt, double** out_dev);
int setup_batch_slot(Workspace* ws, int tslot, const Geom& g, int B, Batch* bt, double** out_dev);
int reserve_batches(Workspace* ws, const Geom& g, int B);
// assemble without touching the info flags / factorise + reduce an assembled batch (ESS wave pipeline, ess.cu)
int assemble_matrices(const Geom& g, const KernelDev* kds, const double* const* ys, const Batch& bt, int B,
                      cudaStream_t st);
int factor_reduce(const Geom& g, const Batch& bt, int B, const ScaleArgs& sa, double* out, cudaStream_t st, int ctx = 0);

// log-likelihoods of B dense nodes: out_dev[b*4 + {0,1,2}] = logdet K, quad (y'K^-1y), sigma2
int loglik_batch_device(Workspace* ws, const KernelDev* kds, const double* const* ys, const ScaleArgs& sa, int B,
                        int64_t n, Batch* bt_out, Geom* g_out, double** out_dev, cudaStream_t st);

// bench.py accounting: matrices of speculative ESS candidates that lay behind the accepted one
void profile_wasted(int matrices, int64_t n);

// matrices per speculative ESS wave (ess.cu)
extern int g_ess_target_b;
extern int g_ess_cached_threshold;
extern int g_ess_rotate_w;
extern int g_ess_prefetch;
extern int g_ess_overlap;
extern int g_ess_wave_total;

// Programmatic dependent launch (the factorisation's critical path is a chain of ~470 short dependent kernels):
// `pdl_wait` blocks until the preceding grid of the stream has completed and flushed (a no-op for a launch
// without the attribute), `pdl_trigger` lets the next grid of the stream be made resident while this one drains,
// so its launch latency overlaps this grid's tail instead of following it.  Resident CTAs of the next grid hold
// their SM slots while they wait, so WHEN to trigger is a trade: right away (`early`) when the critical path bounds
// the factorisation (few matrices: measured -7 % at B = 1 augmented, -1.5 % at B = 4), near the end of the grid when
// the bulk updates need every slot (B = 8: early costs 1 %).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
}  // namespace dgpb

/*
 * dgpb.h -- C-ABI of libdgpb.so: the B200 (sm_100a) implementation of the dgpsi
 * stochastic-imputation (SI) hot path.
 *
 * The reference (mingdeyu/DGP, `dgpsi` 2.6.0) is pure Python + numba and has NO FFI; the drop-in
 * boundary is therefore the "GP-node seam" of SURVEY.md section 8(b): every numeric method of the
 * reference's `kernel` class (dgpsi/kernel_class.py) and the njit kernels they call
 * (dgpsi/functions.py, dgpsi/vecchia.py, dgpsi/imputation.py).  Each entry point below cites the
 * reference function it replaces.  `dgp_b200/` (Python, mirrors the reference API) binds these
 * symbols through ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - FP64, row-major (C order), contiguous; index arrays are int64 (numpy default, vecchia.py:65);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: DGPB_OK, DGPB_NOT_PD (Cholesky hit a non-positive pivot -> the Python shim raises
 *     numpy.linalg.LinAlgError so dgp.train's restart logic, dgp.py:1402, keeps working),
 *     DGPB_BAD_ARG, DGPB_CUDA_ERROR; dgpb_last_error() returns a message for the calling thread;
 *   - functions that return scalars through `*_host` pointers synchronise `stream` before
 *     returning; all others are asynchronous with respect to the host;
 *   - the library owns only the opaque workspace (scratch + cached factorizations); callers own
 *     every data buffer.
 */
#ifndef DGPB_H
#define DGPB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGPB_OK 0
#define DGPB_NOT_PD 1
#define DGPB_BAD_ARG 2
#define DGPB_CUDA_ERROR 3

#define DGPB_SEXP 0      /* kernel(name='sexp')      kernel_class.py:325-332 */
#define DGPB_MATERN25 1  /* kernel(name='matern2.5') kernel_class.py:333-345 */

/* likelihood nodes of a final layer (likelihood_class.py:8-90, 92-243, 245-292) */
#define DGPB_LIK_POISSON 0
#define DGPB_LIK_HETERO 1
#define DGPB_LIK_NEGBIN 2
/* Categorical (likelihood_class.py:294-360): two classes through a logit / probit link on one latent column,
 * K > 2 classes through softmax / robustmax on K columns */
#define DGPB_LIK_CAT_LOGIT 3
#define DGPB_LIK_CAT_PROBIT 4
#define DGPB_LIK_CAT_SOFTMAX 5
#define DGPB_LIK_CAT_ROBUSTMAX 6
/* zero-inflated counts (likelihood_class.py:470-622, 624-814): ZIP reads log rate and logit zero-inflation, ZINB
 * log mean, log dispersion and logit zero-inflation */
#define DGPB_LIK_ZIP 7
#define DGPB_LIK_ZINB 8
#define DGPB_LIK_MAX_IN 8 /* latent columns one likelihood node can read */

#define DGPB_MAX_DIM 32  /* maximum node input dimension (local + connected global) */

typedef struct dgpb_ws dgpb_ws; /* opaque workspace */

/* One GP node of the hierarchy: the numeric state of the reference's `kernel` object
 * (kernel_class.py:86-144) that the hot path reads.
 *
 * The node input is X = [ src[input_dim[0..n_local), :]^T | gsrc[connect[0..n_global), :]^T ]  (n x D):
 * `src` is the feeding layer stored one variable per ROW (width x n, so a latent column is a
 * contiguous n-vector) and `gsrc` is the global input X^T (d x n).  This is `kernel.input` /
 * `kernel.global_input` (dgp.py:592-624) without materialising per-node copies. */
typedef struct dgpb_node {
    int32_t kind;              /* DGPB_SEXP | DGPB_MATERN25 */
    int32_t n_local;           /* len(input_dim) */
    int32_t n_global;          /* len(connect) or 0 */
    int32_t nlen;              /* len(length): 1 (shared) or n_local+n_global (ARD), kernel_class.py:14-20 */
    int32_t input_dim[DGPB_MAX_DIM];
    int32_t connect[DGPB_MAX_DIM];
    double length[DGPB_MAX_DIM];
    double scale;              /* kernel.scale[0]  */
    double nugget;             /* kernel.nugget[0] */
    int32_t scale_est;         /* kernel.scale_est  */
    int32_t nugget_est;        /* kernel.nugget_est */
    const double* src;         /* device, (src_width x n) */
    const double* gsrc;        /* device, (gsrc_width x n) or NULL */
    double* output;            /* device, n : kernel.output[:,0] */
    /* Vecchia state (kernel_class.py:245-267); NULL when dense */
    const int64_t* ord;        /* device, n  */
    const int64_t* NNarray;    /* device, n x (m+1), index-descending, -1 padded */
    int32_t m;                 /* conditioning-set size */
    int32_t vecch;
} dgpb_node;

/* One likelihood node: which rows of the feeding layer it reads (`input_dim`: one row for Poisson -- the log
 * rate -- and the two-class Categorical; two for Hetero -- mean, log variance -- and NegBin -- log mean, log
 * dispersion; ZIP two, ZINB three; K for the K-class Categorical) and the observed outputs (class labels 0..K-1 as doubles). */
typedef struct dgpb_lik {
    int32_t kind;     /* DGPB_LIK_* */
    int32_t n_in;     /* rows used */
    int32_t rows[DGPB_LIK_MAX_IN];
    const double* y;  /* device, n */
    double param;     /* robustmax_eps (DGPB_LIK_CAT_ROBUSTMAX) */
} dgpb_lik;

const char* dgpb_last_error(void);
int dgpb_version(void);
/* sizeof(dgpb_node) as compiled into the library (binding self-check) */
int64_t dgpb_sizeof_node(void);

int dgpb_ws_create(dgpb_ws** ws, int device);
int dgpb_ws_destroy(dgpb_ws* ws);
/* bytes of device scratch currently held by the workspace */
int64_t dgpb_ws_bytes(const dgpb_ws* ws);

/* ---- 1. kernel matrices -------------------------------------------------------------------- */

/* kernel.k_matrix(fod_eval)  kernel_class.py:304-359; functions.py:16-93 (Matern coefficient loops,
 * fod_exp).  X: n x D.  K: n x n (full, symmetric, diag 1+nugget*wdiag_i).  dK: P x n x n or NULL,
 * P = nlen (+1 when nugget_est: the last slice is nugget*diag(wdiag)).  wdiag NULL = ones. */
int dgpb_kmatrix(const double* X, int64_t n, int64_t D, const double* length_host, int64_t nlen,
                 double nugget, const double* wdiag, int kind, int nugget_est,
                 double* K, double* dK, void* stream);

/* ---- 2. dense likelihood / gradient / statistics ------------------------------------------- */

/* kernel.log_likelihood_func  kernel_class.py:481-488: -0.5*(logdet(scale*K) + y'(scale*K)^-1 y).
 * Blocked FP64 Cholesky (DMMA trailing updates) with y carried as an extra row so the forward
 * solve is fused into the factorisation. out_host[0] = llik. */
int dgpb_loglik_dense(dgpb_ws* ws, const dgpb_node* node, int64_t n, double* out_host, void* stream);

/* kernel.llik(x)  kernel_class.py:403-445 (no-replicate branch), WITHOUT the prior terms (the host
 * adds log_prior / log_prior_fod, kernel_class.py:446-448).  node->length/nugget hold exp(x).
 * out_host = [nllik, scale, grad[0..P)],  P = nlen + nugget_est.  K^-1 comes from one sliding-window
 * partial Cholesky of [[K],[y'],[I]]; the P traces/quadratic forms are one fused pass over K^-1 with
 * dK recomputed on the fly (the P x n x n `fod` tensor of functions.py:36-93 is never built). */
int dgpb_nllik_grad_dense(dgpb_ws* ws, const dgpb_node* node, int64_t n, double* out_host, void* stream);

/* dgpb_nllik_grad_dense for B independent nodes of the same n in ONE batched sliding-window factorisation (the
 * M-step of a DGP: given the imputation the nodes are independent, dgp.py:1391-1398, and their L-BFGS-B runs ask
 * for evaluations at the same time).  Row b of out_host (leading dimension ldo >= max P + 2) = [nllik, scale,
 * grad[0..P_b)]; status_host[b] = DGPB_OK or DGPB_NOT_PD per node.  Results are bit-identical to B separate calls. */
int dgpb_nllik_grad_dense_batch(dgpb_ws* ws, const dgpb_node* nodes, int B, int64_t n, double* out_host, int ldo,
                                int* status_host, void* stream);

/* kernel.compute_stats  kernel_class.py:735-748: Rinv (n x n, full symmetric) and Rinv_y (n). */
int dgpb_compute_stats(dgpb_ws* ws, const dgpb_node* node, int64_t n, double* Rinv, double* Rinv_y,
                       void* stream);
/* The same for R = K + diag(diag_shift) (device, n): the linear solves of Hetero.post_het1
 * (likelihood_class.py:185-210), where the observation variances sit on top of the node's nugget. */
int dgpb_compute_stats_shifted(dgpb_ws* ws, const dgpb_node* node, int64_t n, const double* diag_shift, double* Rinv,
                               double* Rinv_y, void* stream);

/* fmvn(scale*K)  functions.py:113-121 with the standard-normal vector z injected:
 * nu = chol(scale*K) z. */
int dgpb_mvn_draw(dgpb_ws* ws, const dgpb_node* node, int64_t n, const double* z, double* nu, void* stream);

/* ---- 3. elliptical slice sampling ---------------------------------------------------------- */

/* imputer.one_sample_block  imputation.py:44-119 (and one_sample :166-221 when n_targets == 1).
 * targets[k].output is f[:,k]; it is the row `target_rows[k]` of the matrix `layer_out`
 * (width x n) that every upper node reads through src/input_dim.  z: (n_targets x n) standard
 * normals; u_host: nu uniforms consumed in the reference's order (threshold, first angle, one per
 * rejection).  On return the accepted proposal has been written to the target outputs,
 * *n_prop_host = proposals evaluated, theta_host (optional, length >= nu) receives the angles
 * tried.  Returns DGPB_BAD_ARG if the uniforms ran out before acceptance. */
int dgpb_ess_block(dgpb_ws* ws, const dgpb_node* targets, int n_targets, const int32_t* target_rows_host,
                   double* layer_out, int64_t layer_width, const dgpb_node* uppers, int n_uppers,
                   int64_t n, const double* z, const double* u_host, int nu,
                   int* n_prop_host, double* theta_host, void* stream);

/* dgpb_ess_block with factor reuse inside one I-step (the hyper-parameters are fixed between M-steps, so the
 * reference's repeated Cholesky factorisations of unchanged matrices, imputation.py:63,70-78, can be skipped
 * without changing any result):
 *   target_keys_host[k] >= 0 : chol(K) of target k is looked up under this key in the workspace (and stored
 *                              there when it has to be computed) instead of being re-factored for fmvn;
 *   upper_keys_host[u]  >= 0 : on acceptance the factor of upper node u is stored under this key -- it is the
 *                              prior factor that node needs as a target of the next layer pair;
 *   threshold_io_host        : NaN on entry = compute the threshold likelihoods; otherwise the sum of the upper
 *                              log-likelihoods at the current state (valid when neither their inputs nor their
 *                              outputs changed since the value was returned).  On return: the accepted sum.
 * Any of the three may be NULL.  dgpb_cache_clear invalidates every stored factor (call it whenever a
 * hyper-parameter, an ordering or a training input changes). */
int dgpb_ess_block_cached(dgpb_ws* ws, const dgpb_node* targets, int n_targets, const int32_t* target_rows_host,
                          double* layer_out, int64_t layer_width, const dgpb_node* uppers, int n_uppers,
                          int64_t n, const double* z, const double* u_host, int nu,
                          int* n_prop_host, double* theta_host, const int32_t* target_keys_host,
                          const int32_t* upper_keys_host, double* threshold_io_host, void* stream);
int dgpb_cache_clear(dgpb_ws* ws);
/* The output of the node stored under `key` was replaced by the caller (e.g. the exact Hetero draw): the cached
 * L^-1 y of that node is stale; thresholds that need y'K^-1y of it solve with the cached factor again. */
int dgpb_cache_output_changed(dgpb_ws* ws, int key);

/* imputer.sample(burnin) (imputation.py:22-42 with block updates, :44-119) of a SMALL dense DGP -- n <= 64 training
 * points, at most 8 nodes per layer and 8 GP layers -- in ONE kernel launch: one CTA keeps the layers, the prior draws
 * and the kernel matrix in shared memory and runs every block update of `sweeps` sweeps (prior draws, threshold,
 * propose -> kernel matrix -> Cholesky -> solve -> accept / shrink) with the caller's pre-drawn numbers; no host
 * round trip per proposal, per update or per sweep.
 * nodes_host: the GP nodes layer by layer (widths_host[l] per layer); layer_ptrs_host[l]: device image of layer l
 * (width_l x n, one latent column per row; nodes of layer l + 1 read it through src / input_dim; the last layer holds
 * the training outputs).  z: device, z_rows x n standard normals, one row per prior draw in the reference's order
 * (sweep, layer pair, target node); u_host: nu uniforms in the reference's order (threshold, first angle, one per
 * rejection, update after update).  counts_host = {uniforms consumed, proposals evaluated, prior draws made}.
 * Returns DGPB_BAD_ARG when the uniforms ran out (the layers are then partly updated: restore and retry). */
int dgpb_ess_sweeps_small(dgpb_ws* ws, const dgpb_node* nodes_host, const int32_t* widths_host, int n_layers,
                          double* const* layer_ptrs_host, int64_t n, int sweeps, const double* z, int64_t z_rows,
                          const double* u_host, int nu, int32_t* counts_host, void* stream);

/* Likelihood layers.  dgpb_lik_loglik: sum of `llik()` (likelihood_class.py:39-48, 110-116, 264-272) over the
 * likelihood nodes for the latent layer image `layer` (layer_width x n, device); one double to the host.
 * dgpb_ess_block_lik: imputer.one_sample_block / one_sample (imputation.py:44-119, 166-221) for target GP nodes whose
 * outputs feed likelihood nodes only; arguments as dgpb_ess_block_cached with the upper GP nodes replaced by `liks`. */
int dgpb_lik_loglik(const dgpb_lik* liks, int n_liks, const double* layer, int64_t n, double* out_host, void* stream);
int dgpb_ess_block_lik(dgpb_ws* ws, const dgpb_node* targets, int n_targets, const int32_t* target_rows_host,
                       double* layer_out, int64_t layer_width, const dgpb_lik* liks, int n_liks, int64_t n,
                       const double* z, const double* u_host, int nu, int* n_prop_host, double* theta_host,
                       const int32_t* target_keys_host, void* stream);

/* ---- 4. Vecchia ---------------------------------------------------------------------------- */

/* nn()  vecchia.py:42-109: ordered nearest neighbours of the (already ordered, already scaled)
 * points x (n x D): row i = {i} U (m nearest j<i), index-descending, -1 padded. NN: n x (m+1). */
int dgpb_knn_ordered(const double* x, int64_t n, int64_t D, int64_t m, int64_t* NN, void* stream);

/* get_pred_nn()  vecchia.py:20-40: plain kNN of M queries against n points, distance ascending
 * (ties: smaller index first).  NN: M x min(m,n). */
int dgpb_knn(const double* query, int64_t M, const double* x, int64_t n, int64_t D, int64_t m,
             int64_t* NN, void* stream);

/* vecchia_llik  vecchia.py:164-180.  X: n x D and y: n in Vecchia order.  out_host[0] = llik. */
int dgpb_vecchia_llik(const double* X, const double* y, const int64_t* NN, int64_t n, int64_t D, int64_t m1,
                      const double* length_host, int64_t nlen, double scale, double nugget,
                      const double* nugget_diag, int kind, double* out_host, void* stream);

/* vecchia_nllik  vecchia.py:182-242 (origin_n == n branch), without prior terms.
 * out_host = [nllik, scale, grad[0..P)]. */
int dgpb_vecchia_nllik(const double* X, const double* y, const int64_t* NN, int64_t n, int64_t D, int64_t m1,
                       const double* length_host, int64_t nlen, double scale, double nugget,
                       const double* nugget_diag, int kind, int scale_est, int nugget_est,
                       double* out_host, void* stream);

/* L_matrix  vecchia.py:409-424: rows of the sparse inverse Cholesky factor, n x m1. */
int dgpb_vecchia_Lmatrix(const double* X, const int64_t* NN, int64_t n, int64_t D, int64_t m1,
                         const double* length_host, int64_t nlen, double nugget, int kind,
                         double* L, void* stream);

/* fmvn_sp  vecchia.py:133-140 with injected z: x = (L/sqrt(scale))^-1 z by a dependency-driven
 * sparse forward solve (forward_solve_sp, vecchia.py:111-120).  out: n (Vecchia order). */
int dgpb_vecchia_mvn_draw(const double* X, const int64_t* NN, int64_t n, int64_t D, int64_t m1,
                          const double* length_host, int64_t nlen, double scale, double nugget, int kind,
                          const double* z, double* out, void* stream);

/* Exact conditional draw of the mean process under the heteroskedastic Gaussian likelihood with the Vecchia
 * approximation: imputer.one_sample imputation.py:141-158 -> U_matrix_sp (vecchia.py:426-445, 612-622) +
 * Hetero.post_het_vecch (likelihood_class.py:165-183).  X: n x D node inputs and y, gamma = exp(log variance),
 * sd (standard normals, the reference's np.random.randn(n)): n each, ALL in Vecchia order; imp_NN: n x m1 =
 * kernel.imp_NNarray (kernel_class.py:268-273; entries >= n are latent, < n observed, -1 padding); kernel scale, no
 * nugget (the reference passes 0).  f_out: n, the drawn mean process in Vecchia order (the caller applies rev_ord).
 * U_out (optional, n x m1): the entries of U in imp_NN's order, i.e. U_matrix()'s rows reversed. */
int dgpb_hetero_vecchia_draw(const double* X, const int64_t* imp_NN, int64_t n, int64_t D, int64_t m1,
                             const double* length_host, int64_t nlen, double scale, int kind, const double* gamma,
                             const double* y, const double* sd, double* f_out, double* U_out, void* stream);

/* gp_vecch  vecchia.py:635-654.  x: M x D test inputs, w: n x D training inputs, NN: M x mp. */
int dgpb_gp_vecch(const double* x, int64_t M, const double* w, const double* y, int64_t n, int64_t D,
                  const int64_t* NN, int64_t mp, const double* length_host, int64_t nlen, double scale,
                  double nugget, const double* nugget_diag, int kind, double* mean, double* var, void* stream);

/* gp_vecch for B squared-exponential nodes with ONE length-scale each that share the inputs w, the test points and
 * the neighbour array (the first layer of a Vecchia DGP, kernel_class.py:586-625 called per node by
 * emulation.py:780-800): the block's raw squared distances are formed once per test point.  Y: B x n outputs,
 * length/scale/nugget_host: B values each, mean/var: B x M.  Needs mp + 1 <= 32. */
int dgpb_gp_vecch_multi(const double* x, int64_t M, const double* w, const double* Y, int64_t n, int64_t D,
                        const int64_t* NN, int64_t mp, int B, const double* length_host, const double* scale_host,
                        const double* nugget_host, double* mean, double* var, void* stream);

/* link_gp_vecch + IJ_nb  vecchia.py:758-907.  m_in/v_in: M x Dw, z: M x Dz or NULL,
 * w1: n x Dw, gw: n x Dz or NULL. */
int dgpb_linkgp_vecch(const double* m_in, const double* v_in, const double* z, int64_t M,
                      const double* w1, const double* gw, const double* y, int64_t n, int64_t Dw, int64_t Dz,
                      const int64_t* NN, int64_t mp, const double* length_host, int64_t nlen, double scale,
                      double nugget, const double* nugget_diag, int kind, double* mean, double* var,
                      void* stream);

/* ---- 5. closed-form prediction ------------------------------------------------------------- */

/* gp()  functions.py:379-394: m = r'R^-1y, v = |scale(1+nugget - r'R^-1 r)|.
 * x: M x D, W: n x D (local and global columns concatenated). */
int dgpb_gp_predict(dgpb_ws* ws, const double* x, int64_t M, const double* W, int64_t n, int64_t D,
                    const double* Rinv, const double* Rinv_y, const double* length_host, int64_t nlen,
                    double scale, double nugget, int kind, double* mean, double* var, void* stream);

/* link_gp()  functions.py:396-430 with IJ_sexp / IJ_matern (functions.py:432-494), Jd / Jd0
 * (vecchia.py:915-988), trace_sum / quad (functions.py:496-506, vecchia.py:990-1000).
 * m_in/v_in: M x Dw Gaussian inputs, z: M x Dz deterministic global inputs or NULL.
 * The sexp R2sexp/Psexp tables of kernel_class.py:752-764 are recomputed on the fly. */
int dgpb_linkgp_predict(dgpb_ws* ws, const double* m_in, const double* v_in, const double* z, int64_t M,
                        const double* w1, const double* gw, int64_t n, int64_t Dw, int64_t Dz,
                        const double* Rinv, const double* Rinv_y, const double* length_host, int64_t nlen,
                        double scale, double nugget, int kind, double* mean, double* var, void* stream);

/* mixture over S imputations  emulation.py:846-847, linkgp.py:493-494.
 * means/vars: S x len; mu = mean_s m; sigma2 = mean_s(m^2+v) - mu^2. */
int dgpb_aggregate(const double* means, const double* vars, int64_t S, int64_t len, double* mu,
                   double* sigma2, void* stream);

/* ---- 6. one chain on several GPUs ------------------------------------------------ */

/* The reference parallelises the M-step over GP nodes in a process pool (dgp.py:1414-1472, `ptrain`) and nothing
 * else.  Here one process drives one GPU and the processes of a box share ONE chain: every rank holds the same
 * latent layers and takes the same decisions; the candidate angles of an ESS wave (imputation.py:107-119) are
 * dealt over the ranks, each rank factors its share and the per-matrix results are all-gathered over NCCL.
 * dgpb_comm_unique_id: 128-byte NCCL id, created on one rank and handed to the others by the host side
 * (torch.distributed broadcast).  dgpb_comm_init attaches a communicator to the workspace; afterwards
 * dgpb_ess_block_cached / dgpb_ess_block_lik on that workspace are COLLECTIVE calls: every rank must make them
 * with identical arguments (same draws, same uniforms).  world = 1 detaches.  Results are bit-identical to the
 * single-GPU call. */
int dgpb_comm_unique_id(char* id128_host);
int dgpb_comm_init(dgpb_ws* ws, int rank, int world, const char* id128_host);
int dgpb_comm_destroy(dgpb_ws* ws);
int dgpb_comm_info(const dgpb_ws* ws, int* rank_host, int* world_host);
/* Host-only plan of one ESS wave (no GPU needed): candidate angles under the assumption that every earlier one is
 * rejected (imputation.py:111-119; u_host = the uniforms those rejections consume) and the rank / local slot that
 * evaluates each item (item 0 = the threshold when first = 1, then the candidates).  cap = candidates per rank. */
int dgpb_ess_plan_wave(double theta0, double tmin, double tmax, const double* u_host, int nu_left, int first, int cap,
                       int world, int* n_cand_host, double* thetas_host, int* rank_host, int* slot_host);

/* ---- measurement helpers (bench.py / DESIGN.md roofline denominators) ----------------------- */

/* C (M x N) = A (M x K) * B (N x K)^T with the library's own DMMA tile kernel. */
int dgpb_dgemm_nt(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K, void* stream);
/* in-place lower Cholesky of the n x n matrix A (ld = n); info_host = 0 or failing column + 1 */
int dgpb_potrf(dgpb_ws* ws, double* A, int64_t n, int* info_host, void* stream);
/* Launch profiler for the roofline line of bench.py: while on, every trailing-update (SYRK) launch of the
 * blocked factorisation is bracketed by CUDA events on its stream.  dgpb_profile_read fills
 * out_host[0..4] = {total ms, launches timed, algorithmic FLOPs of those launches, FLOPs of every factorisation
 * issued while profiling (n^3/3 per matrix, n^3 for the augmented layout), the part of those FLOPs spent on
 * speculative ESS candidates that lay behind the accepted one (issued, but not needed by the reference's schedule)}. */
int dgpb_profile(int on);
/* micro-probe of the trailing-update kernel (development aid): out_host = {ms per launch, TFLOP/s} */
int dgpb_probe_update(dgpb_ws* ws, int64_t n, int B, int flags, int reps, double* out_host);
int dgpb_profile_read(double* out_host);
/* micro-probe of the whole batched likelihood pipeline (assemble, factorise, reduce) on synthetic inputs:
 * out_host = {ms per pipeline, TFLOP/s counting n^3/3 (aug = 0) or n^3 (aug = 1, B = 1) per matrix} */
int dgpb_probe_factorize(dgpb_ws* ws, int64_t n, int B, int aug, int reps, double* out_host);
/* development/benchmark tunables of the blocked factorisation: "hb" (hyper-block width, multiple of 128),
 * "hb_min_w" (smallest remaining window factored with hyper-blocks), "hb_graded" (0/1: ramp the first hyper-blocks
 * 128, 256, 512), and of the ESS loop: "ess_batch" (matrices per speculative wave; <= 1 = one proposal at a time), "ess_prefetch" (1 = the next wave is proposed and assembled while the current one is
 * factored), "ess_trsv" (1 = threshold of a block
 * update from cached factors by a triangular solve when only the upper nodes' outputs moved),
 * and of the neighbour search: "knn_mma" (3 = split-TF32 tensor-core screen + exact FP64 ranking [default], 1 = FP64 DMMA screen, 0 = scalar
 * exact kernel),
 * and of the Vecchia block kernels: "vecchia_small" (1 = register-resident kernel for blocks of <= 32 points),
 * and of link_gp: "linkgp_mma" (1 = squared-exponential exponents on the FP64 tensor path), "linkgp_matern_tab" (1 = Matern-2.5 J
 * integrals from per-point tables of their transcendental factors) */
int dgpb_tune(const char* key, int value);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t dgpb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DGPB_H */

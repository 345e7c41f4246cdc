/* amb200 — C ABI of the B200-native embedding-set distance path.
 *
 * The reference (SonyCSLParis/audio-metrics, pure Python) has no FFI layer: its
 * boundary for this path is the Python call surface of data.py and the modules under metrics/.
 * Each entry point below is what a binding for one of those functions calls; the
 * reference interface it replaces is cited as  file:line  (relative to the
 * reference's src/audio_metrics/).
 *
 * Conventions
 *  - plain C: pointers, sizes, scalars; no torch / C++ types.
 *  - every call names its CUDA device and stream explicitly; the library never
 *    relies on the caller's current device and never syncs the default stream.
 *  - "device-pointer" functions (amb_*): all array arguments are device memory
 *    on `dev`; work is enqueued on `stream` and the call returns without
 *    synchronising unless stated.  Scratch memory is caller-provided (`ws`,
 *    size from the matching *_ws_bytes query); the library allocates nothing.
 *  - "host-buffer" functions (amb_host_*): array arguments are host memory;
 *    the call stages them through device memory it allocates and frees itself,
 *    runs the same kernels, synchronises and returns results in host memory.
 *    These are what a numpy/ctypes binding of the reference calls directly.
 *  - return 0 on success, a negative AMB_ERR_* otherwise; amb_last_error() gives
 *    the calling thread's last message.  There is no CPU fallback: without a
 *    usable CUDA device every compute call fails with AMB_ERR_CUDA.
 */
#ifndef AMB200_H_
#define AMB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* amb_stream_t; /* a cudaStream_t (0 = legacy default stream) */

enum { AMB_F32 = 0, AMB_F64 = 1, AMB_I32 = 2, AMB_I64 = 3, AMB_U8 = 4 }; /* the integer codes: amb_comm_* only */
enum {
  AMB_OK = 0,
  AMB_ERR_ARG = -1,     /* invalid argument (shape, dtype, k, null pointer) */
  AMB_ERR_CUDA = -2,    /* CUDA runtime / launch failure, or no device */
  AMB_ERR_WS = -3,      /* workspace too small */
  AMB_ERR_NUMERIC = -4  /* iteration failed to converge */
};
enum { AMB_KERNEL_POLY = 0, AMB_KERNEL_RBF = 1 };
/* MMD^2 estimators of metrics/kd.py:38-83 (mmd_est); OR in AMB_MMD_UNIT_DIAGONAL for unit_diagonal=True */
enum { AMB_MMD_UNBIASED = 0, AMB_MMD_BIASED = 1, AMB_MMD_USTAT = 2, AMB_MMD_UNIT_DIAGONAL = 4 };

int amb_version(void);
const char* amb_last_error(void);
/* Number of kernels this library has launched in the calling process (for
 * bench.py's gpu_launches claim). */
long long amb_launch_count(void);

/* Per-launch device timing of the all-pairs tensor-core kernel (CUDA events on the
 * launching stream).  amb_profile_read waits for the recorded launches, clears the
 * log and fills out[4] = { launches, total ms, algorithmic pairs, executed MMA flops }. */
int amb_profile_enable(int on);
int amb_profile_read(double* out);

/* Process-wide diagnostic / tuning options.  They select between implementations that give the
 * same results (the test-suite holds them to each other) and are meant to be set once, not
 * toggled around individual calls.  Each option takes its initial value from an environment
 * variable when the library is first used; after that only amb_set_option changes it — no entry
 * point reads the environment per call.  Unknown names / out-of-range values: AMB_ERR_ARG.
 *   "fad_method"          0 polar iteration on fp64 GEMMs (default) | 1 one-sided block Jacobi   AMB_FAD_METHOD
 *   "fad_factor_eig"      1 = Jacobi eigen-factors instead of pivoted Cholesky                   AMB_FAD_FACTOR=eig
 *   "jacobi_block"        0 auto | 4 | 8 | 16 columns per block of the block-Jacobi kernel        AMB_JACOBI_BS
 *   "jacobi_flat"         1 = round-per-grid-barrier Jacobi kernel                               AMB_JACOBI=flat
 *   "fad_ctas"            0 auto, else the most CTAs the cooperative Frechet kernels may use      AMB_FAD_CTAS
 *   "fad_debug"           1 = print factor ranks / Jacobi sweeps to stderr (synchronises)         AMB_FAD_DEBUG
 *   "engine_passes"       0 auto | 3 = three-MMA split-precision sweep for radii / counts         AMB_PASSES
 *   "engine_cta2"         -1 auto | 0 single-CTA engine | 1 CTA pairs                             AMB_CTA2
 *   "engine_static"       1 = round-robin work items instead of the dynamic hand-out             AMB_SCHED=static
 *   "engine_stages", "engine_grid"   B ring depth / persistent CTAs (0 auto)                      AMB_STAGES, AMB_GRID
 *   "engine_reserve_sms"  SMs the all-pairs sweeps leave free for other streams                   AMB_RESERVE_SMS
 *   "tail_split", "topk_split", "count_split"   column splits of the sweeps (0 auto)              AMB_TAIL_SPLIT, ...
 *   "cov_dfma"            1 = FP64-pipe Gram kernel instead of the int8 tensor-core covariance    AMB_COV=dfma
 *   "debug_single"        engine behind amb_debug_dot_matrix: 0 split | 1 single | 2 CTA pair     AMB_DEBUG_SINGLE */
int amb_set_option(const char* name, int value);
int amb_get_option(const char* name);

/* ------------------------------------------------------------------ statistics
 * AudioMetricsData.add / recompute_stats (data.py:37-58): batch mean and unbiased
 * covariance, kept here as fp64 raw moments so that batches, GPUs and calls add. */

/* sum[d] += column sums of X, gram[d*d] += X^T X (full symmetric matrix), fp64.
 * X is [n, d] row-major with leading dimension ld (elements), dtype f32/f64. */
size_t amb_cov_ws_bytes(long long n, int d);
int amb_cov_accumulate(int dev, amb_stream_t stream, const void* X, int dtype, long long n, int d,
                       long long ld, double* sum, double* gram, void* ws, size_t ws_bytes);
/* The streaming form for the embedding pipeline's small batches (embed.py:226-236: <= 32 rows per
 * call, one boolean mask per item category): ONE kernel launch, no workspace, that adds the raw
 * moments of the rows with mask[row] == mask_value (mask == NULL: all rows) into sum / gram.
 * Deterministic (one writer per element).  Meant for n up to a few thousand rows; larger batches
 * are faster through amb_cov_accumulate. */
int amb_cov_accumulate_masked(int dev, amb_stream_t stream, const void* X, int dtype, long long n, int d,
                              long long ld, const int32_t* mask, int mask_value, double* sum,
                              double* gram);
/* mean = sum/n;  cov = (gram - n mean mean^T)/(n-1), zeros when n == 1
 * (data.py:39-44: torch.mean, torch.cov(correction=1), zeros for n == 1). */
int amb_cov_finalize(int dev, amb_stream_t stream, long long n, int d, const double* sum,
                     const double* gram, double* mean, double* cov);
/* Chan pairwise merge of (n1, mean1, cov1) += (n2, mean2, cov2), in place in the
 * first operand (data.py:77-94 _update_stats).  scratch_mean: [d] fp64 scratch. */
int amb_stats_merge(int dev, amb_stream_t stream, int d, long long n1, double* mean1, double* cov1,
                    long long n2, const double* mean2, const double* cov2, double* scratch_mean);

/* --------------------------------------------------------------- Frechet distance
 * frechet_distance / _frechet_distance (metrics/fad.py:8-31):
 *   |mu_x - mu_y|^2 + tr S_x + tr S_y - 2 sum_i sqrt(lambda_i(S_x S_y)),
 * for `batch` independent pairs.  mu_*: [batch, d], cov_*: [batch, d, d], fp64,
 * out: [batch] fp64.  The eigenvalue sum is evaluated as the nuclear norm of
 * M = F_y^T F_x with S = F F^T (pivoted Cholesky factors): tr(U^T M) with U the polar
 * factor of M from an inverse-free polynomial iteration on fp64 GEMMs (or, option
 * "fad_method" = 1, the sum of the singular values from one-sided Jacobi). */
size_t amb_frechet_ws_bytes(int batch, int d);
int amb_frechet(int dev, amb_stream_t stream, int batch, int d, const double* mu_x,
                const double* cov_x, const double* mu_y, const double* cov_y, double* out, void* ws,
                size_t ws_bytes);

/* ------------------------------------------------------------------ PCA projection
 * IncrementalPCA.partial_fit / transform (projection.py:6-46, used at audio_metrics.py:163-209).
 * The principal axes are the eigenvectors of the d x d scatter matrix (amb_cov_* give it), so
 * the fit is one symmetric eigen-decomposition and the projection one pass over the rows.
 *
 * amb_sym_eig: eigen-decomposition of a symmetric positive semi-definite S [d, d] (fp64) by
 * one-sided Jacobi.  evals[d] descending; evecs[d, d]: row i = unit eigenvector of evals[i], with
 * sklearn's svd_flip(u_based_decision=False) sign (largest-magnitude entry of the row positive).
 * amb_pca_transform: out[n, k] (fp64) = (X - mean) components^T, components [k, d] fp64. */
size_t amb_sym_eig_ws_bytes(int d);
int amb_sym_eig(int dev, amb_stream_t stream, int d, const double* S, double* evals, double* evecs,
                void* ws, size_t ws_bytes);
int amb_pca_transform(int dev, amb_stream_t stream, const void* X, int dtype, long long n, int d,
                      long long ld, const double* mean, const double* components, int k, double* out);

/* --------------------------------------------------------------- packed operands
 * Embeddings are rewritten once per set into the tensor-core operand format
 * (two fp16 planes with one power-of-two scale per 256-row tile, plus per-row
 * squared norm and hi-plane residual); KD and PRDC consume that. */
size_t amb_packed_bytes(long long n, int d);
int amb_pack(int dev, amb_stream_t stream, const void* X, int dtype, long long n, int d,
             long long ld, void* packed);

/* ----------------------------------------------------------------- kernel distance
 * kernel_distance -> kid_features_to_metric (metrics/kd.py:29-35,127-194) with
 * polynomial_kernel (kd.py:112-116), kernel_mmd2 (kd.py:119-124) and the unbiased
 * mmd2 (kd.py:38-83).  F1 [n1,d], F2 [n2,d]; idx [S,2,m] int32 holds, per subset,
 * the m row indices drawn from F1 then the m drawn from F2 (the host draws them
 * with numpy's default_rng exactly as kd.py:176,185-186 does).
 * mmd_est selects the estimator of kd.py:38-83: AMB_MMD_UNBIASED is what kernel_mmd2
 * (kd.py:119-124) hard-codes; "biased" and "u-statistic" are only reachable by calling
 * mmd2() directly in the reference.
 * mmd2_out [S] fp64; stats_out[2] = {mean, population std} (kd.py:189-192). */
size_t amb_kd_ws_bytes(int S, int m, int d);
int amb_kd_subsets(int dev, amb_stream_t stream, const void* F1, long long n1, long long ld1,
                   const void* F2, long long n2, long long ld2, int d, int dtype, const int32_t* idx,
                   int S, int m, int kernel_type, double gamma, double coef0, int degree,
                   double sigma, int mmd_est, double* mmd2_out, double* stats_out, void* ws,
                   size_t ws_bytes);

/* --------------------------------------------------------------------------- PRDC
 * Both PRDC entry points use the tensor-core sweep as a filter with a proven
 * error band and re-decide everything inside the band from the original rows in
 * fp64, so they also take the unpacked matrix (X / R / C, dtype, leading dim).
 *
 * nearest_neighbour_distances (metrics/prdc.py:4-14): radius_i = (k+1)-th
 * smallest Euclidean distance from row i to all rows of the set (self
 * included), returned as the correctly rounded fp32 value of the exact distance.
 * Computes radii for rows [row0, row0+nrows) of the set (row0 % 128 == 0; shards that
 * start on an even tile, row0 % 256 == 0, run on the CTA-pair kernel) against all n rows.
 * radii: [nrows] fp32.  1 <= k <= 29 and k+1 <= n (the reference's kthvalue raises for
 * k+1 > n).  nrows == 0 is a no-op for any row0 (empty row shard).
 * Rows whose answer the tensor-core filter cannot certify (near-ties beyond the kept margin,
 * duplicated or collinear rows) are ALL resolved by an exhaustive exact scan; their number
 * is written to n_exhaustive (device int64, nullable) — a cost indicator, not an error. */
size_t amb_knn_ws_bytes(long long nrows, long long n, int d, int k);
int amb_knn_radii(int dev, amb_stream_t stream, const void* X, int dtype, long long ld,
                  const void* packed, long long n, int d, long long row0, long long nrows, int k,
                  float* radii, long long* n_exhaustive, void* ws, size_t ws_bytes);

/* prdc (metrics/prdc.py:18-50) neighbourhood counts for reference rows
 * [row0, row0+nrows) (row0 % 128 == 0) against all m candidates, strict '<' on
 * fp32 distances as in prdc.py:37-48:
 *   col_count[j]  += #{i in shard : D_ij < r_ref[i]}          (density; >0 => precision)
 *   row_recall[i]  = any_j  D_ij < r_cand[j]                  (i relative to row0)
 *   row_cover[i]   = any_j  D_ij < r_ref[i]   (== min_j D_ij < r_ref[i], prdc.py:48)
 * r_ref: [n_ref] fp32 (indexed by absolute row), r_cand: [m] fp32.
 * col_count [m] int32 is accumulated and must be zeroed by the caller before the
 * first shard; row_recall / row_cover [nrows] are overwritten.  nrows == 0 is a no-op for
 * any row0 (empty row shard).
 *
 * Pairs inside the band go through a refine list that lives in the workspace: its capacity
 * is whatever `ws_bytes` holds beyond the fixed part (amb_prdc_ws_list_cap).  n_uncertain
 * (device int64, nullable) receives the number of such pairs; if it exceeds the capacity
 * the excess pairs were NOT re-decided, the outputs of this call are invalid, and the
 * caller repeats the call (after re-zeroing what it accumulated into col_count) with a
 * workspace of amb_prdc_ws_bytes_cap(n_ref, m, n_uncertain) bytes — the number of pairs in
 * the band does not depend on the capacity — or, if that much memory cannot be had, calls
 * amb_prdc_counts_exact, which needs no list.  amb_host_prdc and the Python binding do
 * exactly that, so no input makes them fail where prdc.py:18-50 returns.
 *   amb_prdc_ws_bytes       workspace with the default capacity amb_prdc_list_cap(n_ref, m)
 *   amb_prdc_ws_bytes_cap   workspace for a given capacity
 *   amb_prdc_ws_list_cap    capacity a workspace of ws_bytes provides */
size_t amb_prdc_ws_bytes(long long n_ref, long long m);
size_t amb_prdc_ws_bytes_cap(long long n_ref, long long m, long long list_cap);
long long amb_prdc_ws_list_cap(long long n_ref, long long m, size_t ws_bytes);
long long amb_prdc_list_cap(long long n_ref, long long m);
int amb_prdc_counts(int dev, amb_stream_t stream, const void* R, long long ldr,
                    const void* packed_ref, long long n_ref, const float* r_ref, const void* C,
                    long long ldc, const void* packed_cand, long long m, const float* r_cand, int d,
                    int dtype, long long row0, long long nrows, int32_t* col_count,
                    uint8_t* row_recall, uint8_t* row_cover, long long* n_uncertain, void* ws,
                    size_t ws_bytes);
/* The same outputs from an exhaustive fp64 evaluation of every (row, candidate) pair on the
 * CUDA cores: no packed operands, no workspace, no list, n_ref * m * d operations.  The last
 * rung of the overflow ladder described above; also a cross-check for the filtered path. */
int amb_prdc_counts_exact(int dev, amb_stream_t stream, const void* R, long long ldr, long long n_ref,
                          const float* r_ref, const void* C, long long ldc, long long m,
                          const float* r_cand, int d, int dtype, long long row0, long long nrows,
                          int32_t* col_count, uint8_t* row_recall, uint8_t* row_cover);
/* totals[4] (int64) += { #cols with count>0, sum of col_count, #rows recalled,
 * #rows covered } — the numerators of precision, density*k, recall, coverage.
 * Null array arguments are skipped. */
int amb_prdc_reduce(int dev, amb_stream_t stream, const int32_t* col_count, long long m,
                    const uint8_t* row_recall, const uint8_t* row_cover, long long nrows,
                    long long* totals);

/* -------------------------------------------------------------- host-buffer calls
 * Same computations with host arrays in and host scalars out (H2D, kernels, D2H
 * and a stream synchronise inside the call). */
int amb_host_stats(int dev, const void* X, int dtype, long long n, int d, double* mean,
                   double* cov);
int amb_host_frechet(int dev, int d, const double* mu_x, const double* cov_x, const double* mu_y,
                     const double* cov_y, double* out);
int amb_host_kd(int dev, const void* F1, long long n1, const void* F2, long long n2, int d,
                int dtype, const int32_t* idx, int S, int m, double gamma, double coef0, int degree,
                double* mmd2_out, double* stats_out);
int amb_host_knn_radii(int dev, const void* X, int dtype, long long n, int d, int k, float* radii);
/* out[4] = precision, recall, density, coverage (prdc.py:50). */
int amb_host_prdc(int dev, const void* ref, long long n, const void* cand, long long m, int d,
                  int dtype, int k, double* out);

/* Everything AudioMetrics.evaluate computes after embedding (audio_metrics.py:254-264) in ONE call on
 * host arrays, sharded over the n_dev devices of devs[] inside this process (one worker thread per
 * device — the reference's multi-GPU mechanism is threads in one process too,
 * util/gpu_parallel.py:20-76): every device uploads the rows once, takes a 256-aligned row shard of
 * both all-pairs sweeps against all columns, and exchanges radii slices (allgather) and
 * per-candidate counts (allreduce) with amb_comm_* on the device (communicators are created on the
 * first call with a device list and kept for the life of the process); statistics, the Frechet
 * distance and the kernel distance run on devs[0] behind its uploads.  No process group; with
 * n_dev == 1 no NCCL either.
 *   want_fad != 0           out[0] = Frechet distance of (candidate, reference)
 *   kd_idx != NULL          out[1], out[2] = kernel_distance_mean / _std; kd_idx [S, 2, msub] int32 as
 *                           amb_kd_subsets (drawn for features_1 = candidate, features_2 = reference)
 *   k > 0                   out[3..6] = precision, recall, density, coverage (prdc.py:50)
 * Entries that were not asked for are NaN. */
int amb_host_evaluate(const int* devs, int n_dev, const void* ref, long long n, const void* cand,
                      long long m, int d, int dtype, int k, const int32_t* kd_idx, int S, int msub,
                      int want_fad, double* out);

/* ------------------------------------------------------------------ collectives
 * The exchanges between row shards (allreduce of fp64 moments, allgather of radii
 * slices, allreduce of per-candidate counts) for a caller that holds device pointers on several
 * GPUs of ONE process — the reference's multi-GPU model (threads of one process,
 * util/gpu_parallel.py:20-76).  amb_comm_init creates one NCCL communicator per listed device
 * (ncclCommInitAll); `rank` is the position in devs[].  Every call is asynchronous on the given
 * stream of that rank's device.  Either one host thread per rank calls its own rank, or one thread
 * issues all ranks between amb_comm_group_begin / _end.  In place: send == recv for allreduce,
 * send == recv + rank * count_per_rank * sizeof(dtype) for allgather.
 * NCCL is bound at run time (libnccl.so.2, shared with a PyTorch already in the process); without
 * it amb_comm_init returns AMB_ERR_CUDA and nothing else in the library is affected.
 * amb_host_evaluate with n_dev > 1 uses exactly these calls. */
typedef struct amb_comm amb_comm_t;
enum { AMB_SUM = 0, AMB_MAX = 1 };
int amb_comm_init(const int* devs, int n_dev, amb_comm_t** comm);
int amb_comm_size(const amb_comm_t* comm);
int amb_comm_device(const amb_comm_t* comm, int rank);
int amb_comm_group_begin(void);
int amb_comm_group_end(void);
int amb_comm_allreduce(amb_comm_t* comm, int rank, const void* send, void* recv, long long count,
                       int dtype, int op, amb_stream_t stream);
int amb_comm_allgather(amb_comm_t* comm, int rank, const void* send, void* recv,
                       long long count_per_rank, int dtype, amb_stream_t stream);
int amb_comm_destroy(amb_comm_t* comm);

/* ------------------------------------------------------------- debug / validation
 * Full dot-product matrix through the tensor-core engine (small sizes only):
 * C[i*ldc + j] = <A_i, B_j> from packed operands; with ldc == 0 only the row
 * sums C[i] = sum_j <A_i, B_j> are written (timing runs).  lbo/sbo override the
 * shared-memory descriptor strides (0 = library default). */
int amb_debug_dot_matrix(int dev, amb_stream_t stream, const void* packed_a, long long na,
                         const void* packed_b, long long nb, int d, float* C, long long ldc,
                         unsigned lbo, unsigned sbo);

#ifdef __cplusplus
}
#endif
#endif /* AMB200_H_ */

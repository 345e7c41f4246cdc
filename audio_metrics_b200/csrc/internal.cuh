// Library-internal declarations shared by the .cu translation units.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/amb200.h"
#include "packed.cuh"

namespace amb {

// Thread-local last-error text behind amb_last_error().
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
int check_cuda(cudaError_t e, const char* what);

// Sets the calling thread's current device for the duration of a call and
// restores the previous one, so the library never depends on (or disturbs)
// the caller's current device.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev);
  ~DeviceGuard();
};

int sm_count(int dev);
// Process-wide tuning / diagnostic options (amb_set_option).  Each is initialised ONCE from its
// environment variable when the library is first used and afterwards only changes through
// amb_set_option; no entry point reads the environment per call.  -1 = "not set" where 0 is a value.
enum Opt {
  kOptJacobiBlock,    // "jacobi_block"        AMB_JACOBI_BS     0 auto | 4 | 8 | 16
  kOptFadCtas,        // "fad_ctas"            AMB_FAD_CTAS      0 auto, else CTA cap of the cooperative FAD kernels
  kOptReserveSms,     // "engine_reserve_sms"  AMB_RESERVE_SMS   SMs the all-pairs sweeps leave free
  kOptFadMethod,      // "fad_method"          AMB_FAD_METHOD    0 polar iteration (GEMMs) | 1 one-sided Jacobi
  kOptPasses,         // "engine_passes"       AMB_PASSES        0 auto | 3 three-MMA split sweep
  kOptCta2,           // "engine_cta2"         AMB_CTA2          -1 auto | 0 single-CTA engine | 1 CTA pairs
  kOptSchedStatic,    // "engine_static"       AMB_SCHED=static  1 round-robin work items
  kOptStages,         // "engine_stages"       AMB_STAGES        0 auto, else B ring depth
  kOptGrid,           // "engine_grid"         AMB_GRID          0 auto, else persistent CTAs
  kOptTailSplit,      // "tail_split"          AMB_TAIL_SPLIT
  kOptTopkSplit,      // "topk_split"          AMB_TOPK_SPLIT
  kOptCountSplit,     // "count_split"         AMB_COUNT_SPLIT
  kOptDebugSingle,    // "debug_single"        AMB_DEBUG_SINGLE  amb_debug_dot_matrix engine: 0 split | 1 single | 2 pair
  kOptCovDfma,        // "cov_dfma"            AMB_COV=dfma      1 FP64-pipe Gram kernel for every input
  kOptFadFactorEig,   // "fad_factor_eig"      AMB_FAD_FACTOR=eig 1 Jacobi eigen-factors instead of pivoted Cholesky
  kOptJacobiFlat,     // "jacobi_flat"         AMB_JACOBI=flat   1 round-per-grid-barrier Jacobi kernel
  kOptFadDebug,       // "fad_debug"           AMB_FAD_DEBUG     1 print ranks / sweeps (synchronises)
  kOptCount
};
int option(Opt o);

// Optional per-launch timing of the pair engine (amb_profile_*): CUDA events on the
// launching stream around the kernel.  No-ops unless enabled.
void* profile_begin(cudaStream_t stream);
void profile_end(void* token, cudaStream_t stream, double alg_pairs, double exec_flops);

// Layout of a packed blob (see packed.cuh): [hi plane][lo plane][inv_scale][norm][rho][row_exp][cmin].
struct PackedLayout {
  long long rows_pad;
  int kpad;
  int kb_count;
  long long plane_halfs;
  size_t off_lo, off_inv, off_norm, off_rho, off_exp, off_cmin, bytes;
};
PackedLayout packed_layout(long long n_rows, int d);

struct PackedPtrs {
  __half* planes;
  long long plane_halfs;
  float* inv_scale;
  float* norm;
  float* rho;   // |x - hi(x)|, rounded up: what the single-pass filter's error band is built from
  int* row_exp; // binary exponent of each row's largest magnitude (scratch of the pack pass)
  float* cmin;  // [rows_pad / 32] smallest squared norm of each 32-row chunk
  long long rows_pad;
  int kb_count;
};
PackedPtrs packed_ptrs(void* blob, long long n_rows, int d);

int launch_pack(cudaStream_t stream, const void* src, int dtype, long long ld, int d,
                long long n_src_rows, const int* gather, long long n_valid, long long row0,
                long long n_rows_out, __half* planes, long long plane_halfs, int kb_count,
                float* inv_scale, float* norm, float* rho, int* row_exp, float* cmin);

// One-sided (Hestenes) Jacobi on the columns of n_mat d x d fp64 matrices stored as rows of Gt
// (fad.cu).  counters: kJacobiCounters zeroed-by-callee ints.  values_only: loose stopping rule that
// is exact for the SUM of the column norms; otherwise columns are orthogonalised to 1e-14.
constexpr int kJacobiCounters = 64;
int launch_jacobi(cudaStream_t st, int dev, double* Gt, int d, int n_mat, int* counters, bool values_only);

// Integer tensor-core Gram / column sums of an fp32 matrix (cov_tc.cu).
size_t cov_tc_ws_bytes(long long n, int d);
int cov_tc_accumulate(cudaStream_t st, int dev, const float* X, long long n, int d, long long ld, double* sum,
                      double* gram, void* ws);

}  // namespace amb

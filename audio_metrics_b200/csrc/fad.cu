// Frechet distance from (mean, covariance) pairs (reference metrics/fad.py:8-31).
//
//   FAD = |mu_x - mu_y|^2 + tr S_x + tr S_y - 2 c,   c = sum_i sqrt(lambda_i(S_x S_y)).
//
// The reference takes complex eigenvalues of the non-symmetric product.  Here c is
// evaluated as the nuclear norm of  M = F_y^T F_x  with  S = F F^T :  the singular
// values of M are exactly sqrt(lambda_i(S_x S_y)), they are real and non-negative
// by construction (matching Re sqrt of a round-off-negative eigenvalue = 0 in
// fad.py:30), and one-sided Jacobi delivers them to high relative accuracy, which
// the trace needs because FAD is a small difference of O(tr S) terms.
//   stage 1  one-sided (Hestenes) Jacobi on S_x and S_y themselves: the rotated
//            columns converge to V Lambda, so F = V Lambda^{1/2} = G Lambda^{-1/2}
//   stage 2  M^T = F_x^T F_y                       (fp64 GEMM, d^3)
//   stage 3  one-sided Jacobi on M; c = sum of column norms
// All in fp64 on the CUDA cores; matrices are d x d (d <= 2048) and live in L2.
// The Jacobi kernel is one cooperative launch per stage: a warp owns one column
// pair per round of a round-robin tournament, a grid barrier separates rounds.
#include <cooperative_groups.h>

#include "internal.cuh"

namespace cg = cooperative_groups;

namespace amb {

constexpr int kJacobiMaxSweeps = 60;

// Round-robin tournament (circle method) on n_even players: round r in
// [0, n_even-1), slot s in [0, n_even/2).  Player n_even-1 is fixed.
__device__ __forceinline__ void tournament_pair(int n_even, int r, int s, int& p, int& q) {
  const int m = n_even - 1;
  if (s == 0) {
    p = m;
    q = r % m;
  } else {
    p = (r + s) % m;
    q = (r - s + m) % m;
  }
}

// Gt: [n_mat][d][d], row j = column j of the matrix being orthogonalised.
// NR > 0: a lane keeps its d/32 <= NR elements of both columns in registers
// between the dot products and the rotation; NR == 0 re-reads them from L2.
// Loads bypass L1 (ld.cg): other SMs rewrite these rows between rounds.
template <int NR>
__global__ void __launch_bounds__(256)
jacobi_kernel(double* __restrict__ Gt, int d, int n_mat, double tol, int* __restrict__ rot_count /* [kJacobiMaxSweeps] zeroed */,
              int* __restrict__ sweeps_done) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const int n_even = d + (d & 1);
  const int slots = n_even / 2;
  const int total = slots * n_mat;
  int sweep = 0;
  for (; sweep < kJacobiMaxSweeps; ++sweep) {
    int my_rot = 0;
    for (int r = 0; r < n_even - 1; ++r) {
      for (int w = warp_global; w < total; w += n_warps) {
        const int mat = w / slots, s = w - mat * slots;
        int p, q;
        tournament_pair(n_even, r, s, p, q);
        if (p >= d || q >= d) continue;   // dummy player of an odd-sized problem
        double* gp = Gt + (static_cast<long long>(mat) * d + p) * d;
        double* gq = Gt + (static_cast<long long>(mat) * d + q) * d;
        double a = 0.0, b = 0.0, c = 0.0;
        double xr[NR > 0 ? NR : 1], yr[NR > 0 ? NR : 1];
        if (NR > 0) {
#pragma unroll
          for (int e = 0; e < NR; ++e) {
            const int k = lane + 32 * e;
            xr[e] = k < d ? __ldcg(gp + k) : 0.0;
            yr[e] = k < d ? __ldcg(gq + k) : 0.0;
            a = fma(xr[e], xr[e], a);
            b = fma(yr[e], yr[e], b);
            c = fma(xr[e], yr[e], c);
          }
        } else {
          for (int k = lane; k < d; k += 32) {
            const double x = __ldcg(gp + k), y = __ldcg(gq + k);
            a = fma(x, x, a);
            b = fma(y, y, b);
            c = fma(x, y, c);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
          c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (a > 0.0 && b > 0.0 && fabs(c) > tol * sqrt(a * b)) {
          const double zeta = (b - a) / (2.0 * c);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / sqrt(1.0 + t * t);
          const double sn = cs * t;
          if (NR > 0) {
#pragma unroll
            for (int e = 0; e < NR; ++e) {
              const int k = lane + 32 * e;
              if (k < d) {
                gp[k] = cs * xr[e] - sn * yr[e];
                gq[k] = sn * xr[e] + cs * yr[e];
              }
            }
          } else {
            for (int k = lane; k < d; k += 32) {
              const double x = __ldcg(gp + k), y = __ldcg(gq + k);
              gp[k] = cs * x - sn * y;
              gq[k] = sn * x + cs * y;
            }
          }
          ++my_rot;
        }
      }
      grid.sync();
    }
    if (lane == 0 && my_rot) atomicAdd(&rot_count[sweep], my_rot);
    grid.sync();
    if (*reinterpret_cast<volatile int*>(&rot_count[sweep]) == 0) { ++sweep; break; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *sweeps_done = sweep;
}

// F^T row j = G^T row j / sqrt(|g_j|)   (eigenvalue lambda_j = |g_j| for PSD input)
__global__ void factor_scale_kernel(double* __restrict__ Gt, int d, int n_mat) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= d * n_mat) return;
  double* g = Gt + static_cast<long long>(w) * d;
  double a = 0.0;
  for (int k = lane; k < d; k += 32) a = fma(g[k], g[k], a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  const double lam = sqrt(a);
  const double sc = lam > 0.0 ? 1.0 / sqrt(lam) : 0.0;
  for (int k = lane; k < d; k += 32) g[k] *= sc;
}

// C[b] = A[b] * B[b]^T  (all row-major d x d, fp64): C[i][j] = <A row i, B row j>.
__global__ void __launch_bounds__(256)
dgemm_nt_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, int d) {
  __shared__ double As[16][65];
  __shared__ double Bs[16][65];
  const long long mo = static_cast<long long>(blockIdx.z) * d * d;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < d; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int rr = e >> 4, kk = e & 15;
      As[kk][rr] = (i0 + rr < d && k0 + kk < d) ? A[mo + static_cast<long long>(i0 + rr) * d + k0 + kk] : 0.0;
      Bs[kk][rr] = (j0 + rr < d && k0 + kk < d) ? B[mo + static_cast<long long>(j0 + rr) * d + k0 + kk] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) { a[v] = As[kk][ty * 4 + v]; b[v] = Bs[kk][tx * 4 + v]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gi = i0 + ty * 4 + i, gj = j0 + tx * 4 + j;
      if (gi < d && gj < d) C[mo + static_cast<long long>(gi) * d + gj] = acc[i][j];
    }
}

// one block per pair: a = |mu_x-mu_y|^2, b = tr S_x + tr S_y, c = sum_j |m_j|
__global__ void __launch_bounds__(256)
fad_combine_kernel(int d, const double* __restrict__ mu_x, const double* __restrict__ cov_x,
                   const double* __restrict__ mu_y, const double* __restrict__ cov_y,
                   const double* __restrict__ Mt, double* __restrict__ out) {
  __shared__ double red[256];
  const int b = blockIdx.x;
  const long long mo = static_cast<long long>(b) * d * d;
  double acc = 0.0;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    const double df = mu_x[static_cast<long long>(b) * d + k] - mu_y[static_cast<long long>(b) * d + k];
    acc += df * df + cov_x[mo + static_cast<long long>(k) * d + k] + cov_y[mo + static_cast<long long>(k) * d + k];
  }
  // column norms of M (rows of Mt): one warp per row, strided
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double csum = 0.0;
  for (int j = warp; j < d; j += 8) {
    const double* g = Mt + mo + static_cast<long long>(j) * d;
    double a = 0.0;
    for (int k = lane; k < d; k += 32) a = fma(g[k], g[k], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    csum += sqrt(a);
  }
  if (lane == 0) acc -= 2.0 * csum;
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[b] = red[0];
}

static int launch_jacobi(cudaStream_t st, int dev, double* Gt, int d, int n_mat, int* counters) {
  const double tol = 1e-14;
  int* rot = counters;
  int* sweeps = counters + kJacobiMaxSweeps;
  int rc = check_cuda(cudaMemsetAsync(counters, 0, (kJacobiMaxSweeps + 1) * sizeof(int), st), "memset");
  if (rc) return rc;
  void* fn = d <= 256 ? reinterpret_cast<void*>(jacobi_kernel<8>)
             : d <= 512 ? reinterpret_cast<void*>(jacobi_kernel<16>)
                        : reinterpret_cast<void*>(jacobi_kernel<0>);
  int per_sm = 0;
  rc = check_cuda(d <= 256   ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_kernel<8>, 256, 0)
                  : d <= 512 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_kernel<16>, 256, 0)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_kernel<0>, 256, 0),
                  "occupancy");
  if (rc) return rc;
  if (per_sm < 1) return set_error(AMB_ERR_CUDA, "jacobi_kernel cannot be resident");
  const int slots = (d + (d & 1)) / 2;
  long long want = (static_cast<long long>(slots) * n_mat + 7) / 8;
  const long long cap = static_cast<long long>(per_sm) * sm_count(dev);
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  void* args[] = {&Gt, &d, &n_mat, const_cast<double*>(&tol), &rot, &sweeps};
  rc = check_cuda(cudaLaunchCooperativeKernel(fn, dim3(static_cast<unsigned>(want)),
                                              dim3(256), args, 0, st),
                  "cudaLaunchCooperativeKernel(jacobi_kernel)");
  if (rc) return rc;
  return check_launch("jacobi_kernel");
}

}  // namespace amb

using namespace amb;

extern "C" {

size_t amb_frechet_ws_bytes(int batch, int d) {
  if (batch <= 0 || d <= 0) return 0;
  return static_cast<size_t>(3) * batch * d * d * 8 + 1024;
}

int amb_frechet(int dev, amb_stream_t stream, int batch, int d, const double* mu_x,
                const double* cov_x, const double* mu_y, const double* cov_y, double* out, void* ws,
                size_t ws_bytes) {
  if (!mu_x || !cov_x || !mu_y || !cov_y || !out || batch <= 0 || d <= 0)
    return set_error(AMB_ERR_ARG, "amb_frechet: bad argument");
  if (d > 2048) return set_error(AMB_ERR_ARG, "amb_frechet: d=%d > 2048 not supported", d);
  const size_t need = amb_frechet_ws_bytes(batch, d);
  if (!ws || ws_bytes < need) return set_error(AMB_ERR_WS, "amb_frechet: workspace %zu < %zu", ws_bytes, need);
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t mat = static_cast<size_t>(d) * d;
  double* G = static_cast<double*>(ws);          // [2*batch] : x factors then y factors
  double* Mt = G + 2 * batch * mat;              // [batch]
  int* counters = reinterpret_cast<int*>(Mt + batch * mat);
  int rc;
  if ((rc = check_cuda(cudaMemcpyAsync(G, cov_x, batch * mat * 8, cudaMemcpyDeviceToDevice, st), "memcpy"))) return rc;
  if ((rc = check_cuda(cudaMemcpyAsync(G + batch * mat, cov_y, batch * mat * 8, cudaMemcpyDeviceToDevice, st), "memcpy"))) return rc;
  if ((rc = launch_jacobi(st, dev, G, d, 2 * batch, counters))) return rc;
  factor_scale_kernel<<<(2 * batch * d * 32 + 255) / 256, 256, 0, st>>>(G, d, 2 * batch);
  if ((rc = check_launch("factor_scale_kernel"))) return rc;
  dim3 ggrid((d + 63) / 64, (d + 63) / 64, batch);
  dgemm_nt_kernel<<<ggrid, 256, 0, st>>>(G, G + batch * mat, Mt, d);   // Mt = Fx^T-rows . Fy^T-rows
  if ((rc = check_launch("dgemm_nt_kernel"))) return rc;
  if ((rc = launch_jacobi(st, dev, Mt, d, batch, counters + 64))) return rc;
  fad_combine_kernel<<<batch, 256, 0, st>>>(d, mu_x, cov_x, mu_y, cov_y, Mt, out);
  return check_launch("fad_combine_kernel");
}

}  // extern "C"

// PCA projection on the device (reference projection.py:6-46, a thin subclass of sklearn's
// IncrementalPCA; call sites audio_metrics.py:163-209).  The reference moves the whole embedding
// set to the host, lets sklearn take an SVD of the centred data, and feeds fp64 projections back
// into every metric.  Here the principal axes come from the symmetric eigen-decomposition of the
// d x d scatter matrix that the covariance kernels already produce (same subspace, same singular
// values: S_i^2 = eigenvalue), and the projection is one HBM-bound pass over the embeddings.
#include "internal.cuh"

namespace amb {

// After one-sided Jacobi on a symmetric PSD matrix the rows of G are lambda_j v_j: eigenvalue =
// row norm, eigenvector = row / norm.  One block: norms, descending rank by counting, normalise,
// sklearn's svd_flip(u_based_decision=False) sign (largest-magnitude entry of each row positive).
__global__ void __launch_bounds__(1024)
sym_eig_finish_kernel(const double* __restrict__ G, int d, double* __restrict__ evals, double* __restrict__ evecs) {
  extern __shared__ double s_norm[];   // [d]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  for (int j = warp; j < d; j += n_warps) {
    const double* g = G + static_cast<long long>(j) * d;
    double a = 0.0;
    for (int k = lane; k < d; k += 32) a = fma(g[k], g[k], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) s_norm[j] = sqrt(a);
  }
  __syncthreads();
  for (int j = warp; j < d; j += n_warps) {
    const double nj = s_norm[j];
    int rank = 0;
    for (int i = lane; i < d; i += 32) {
      const double ni = s_norm[i];
      rank += (ni > nj || (ni == nj && i < j)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    const double* g = G + static_cast<long long>(j) * d;
    // entry of largest magnitude (first one on ties, as numpy's argmax)
    double best = -1.0;
    int best_k = 0;
    for (int k = lane; k < d; k += 32) {
      const double v = fabs(g[k]);
      if (v > best) { best = v; best_k = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
      if (ob > best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
    }
    const double sgn = g[best_k] < 0.0 ? -1.0 : 1.0;
    const double sc = nj > 0.0 ? sgn / nj : 0.0;
    double* out = evecs + static_cast<long long>(rank) * d;
    for (int k = lane; k < d; k += 32) out[k] = g[k] * sc;
    if (lane == 0) evals[rank] = nj;
  }
}

// out[i][c] = sum_k (X[i][k] - mean[k]) comp[c][k], fp64.  One warp per row; the k x d component
// matrix (k <= 64 here: n_pca) is read through L1/L2 by every warp, the embeddings stream once.
template <typename T, int KC>
__global__ void __launch_bounds__(256)
pca_transform_kernel(const T* __restrict__ X, long long n, int d, long long ld, const double* __restrict__ mean,
                     const double* __restrict__ comp, int k, int c0, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  if (row >= n) return;
  const T* x = X + row * ld;
  double acc[KC];
#pragma unroll
  for (int c = 0; c < KC; ++c) acc[c] = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double v = static_cast<double>(x[j]) - mean[j];
#pragma unroll
    for (int c = 0; c < KC; ++c)
      if (c0 + c < k) acc[c] = fma(v, comp[static_cast<long long>(c0 + c) * d + j], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < KC; ++c) {
    double a = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0 && c0 + c < k) out[row * k + c0 + c] = a;
  }
}

}  // namespace amb

using namespace amb;

extern "C" {

size_t amb_sym_eig_ws_bytes(int d) {
  if (d <= 0) return 0;
  return static_cast<size_t>(d) * d * 8 + 1024;
}

int amb_sym_eig(int dev, amb_stream_t stream, int d, const double* S, double* evals, double* evecs, void* ws,
                size_t ws_bytes) {
  if (!S || !evals || !evecs || d <= 0) return set_error(AMB_ERR_ARG, "amb_sym_eig: bad argument");
  if (d > 2048) return set_error(AMB_ERR_ARG, "amb_sym_eig: d=%d > 2048 not supported", d);
  const size_t need = amb_sym_eig_ws_bytes(d);
  if (!ws || ws_bytes < need) return set_error(AMB_ERR_WS, "amb_sym_eig: workspace %zu < %zu", ws_bytes, need);
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* G = static_cast<double*>(ws);
  int* counters = reinterpret_cast<int*>(G + static_cast<size_t>(d) * d);
  int rc;
  if ((rc = check_cuda(cudaMemcpyAsync(G, S, static_cast<size_t>(d) * d * 8, cudaMemcpyDeviceToDevice, st), "memcpy"))) return rc;
  if ((rc = launch_jacobi(st, dev, G, d, 1, counters, false))) return rc;
  sym_eig_finish_kernel<<<1, 1024, static_cast<size_t>(d) * 8, st>>>(G, d, evals, evecs);
  return check_launch("sym_eig_finish_kernel");
}

int amb_pca_transform(int dev, amb_stream_t stream, const void* X, int dtype, long long n, int d, long long ld,
                      const double* mean, const double* components, int k, double* out) {
  if (!X || !mean || !components || !out || n < 0 || d <= 0 || k <= 0 || ld < d)
    return set_error(AMB_ERR_ARG, "amb_pca_transform: bad argument");
  if (dtype != AMB_F32 && dtype != AMB_F64) return set_error(AMB_ERR_ARG, "amb_pca_transform: bad dtype");
  if (n == 0) return AMB_OK;
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((n * 32 + 255) / 256);
  constexpr int KC = 16;   // components per pass over the rows
  for (int c0 = 0; c0 < k; c0 += KC) {
    if (dtype == AMB_F32)
      pca_transform_kernel<float, KC><<<blocks, 256, 0, st>>>(static_cast<const float*>(X), n, d, ld, mean, components,
                                                              k, c0, out);
    else
      pca_transform_kernel<double, KC><<<blocks, 256, 0, st>>>(static_cast<const double*>(X), n, d, ld, mean,
                                                               components, k, c0, out);
    int rc = check_launch("pca_transform_kernel");
    if (rc) return rc;
  }
  return AMB_OK;
}

}  // extern "C"

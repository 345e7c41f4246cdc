// Mean / covariance statistics (reference data.py:37-58, 77-94) as fp64 raw
// moments: column sums and the Gram matrix X^T X, accumulated on the FP64 pipe
// (products of fp32 inputs are exact in fp64, so the only rounding is the fp64
// summation), then finalised to (mean, unbiased covariance) and merged with
// Chan's pairwise update.
#include <cstdlib>

#include "internal.cuh"

namespace amb {

constexpr int kGT = 128;      // Gram tile edge
constexpr int kGK = 16;       // rows per shared-memory stage

template <typename T>
__global__ void __launch_bounds__(256)
gram_tile_kernel(const T* __restrict__ X, long long n, int d, long long ld, int nt, int n_slabs,
                 double* __restrict__ partial /* [n_slabs][d][d], upper tiles only */) {
  __shared__ __align__(16) double As[kGK][kGT];
  __shared__ __align__(16) double Bs[kGK][kGT];
  // tile pair (ti <= tj) from linear index
  int tp = blockIdx.x, ti = 0;
  while (tp >= nt - ti) { tp -= nt - ti; ++ti; }
  const int tj = ti + tp;
  const int slab = blockIdx.y;
  const long long rows_per = (n + n_slabs - 1) / n_slabs;
  const long long r0 = slab * rows_per;
  long long r1 = r0 + rows_per;
  if (r1 > n) r1 = n;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  double acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;

  for (long long r = r0; r < r1; r += kGK) {
    // stage kGK rows x 128 columns of both column blocks (converted to fp64)
    for (int e = tid; e < kGK * kGT; e += 256) {
      const int rr = e / kGT, cc = e % kGT;
      const long long row = r + rr;
      const int ca = ti * kGT + cc, cb = tj * kGT + cc;
      As[rr][cc] = (row < r1 && ca < d) ? static_cast<double>(X[row * ld + ca]) : 0.0;
      Bs[rr][cc] = (row < r1 && cb < d) ? static_cast<double>(X[row * ld + cb]) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGK; ++kk) {
      double a[8], b[8];
#pragma unroll
      for (int v = 0; v < 8; v += 2) {
        const double2 av = *reinterpret_cast<const double2*>(&As[kk][ty * 8 + v]);
        const double2 bv = *reinterpret_cast<const double2*>(&Bs[kk][tx * 8 + v]);
        a[v] = av.x; a[v + 1] = av.y;
        b[v] = bv.x; b[v + 1] = bv.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  double* out = partial + static_cast<long long>(slab) * d * d;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gi = ti * kGT + ty * 8 + i;
    if (gi >= d) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gj = tj * kGT + tx * 8 + j;
      if (gj < d) out[static_cast<long long>(gi) * d + gj] = acc[i][j];
    }
  }
}

// Small batches (the <= 32-row batches of the embedding pipeline, embed.py:226-236, up to a few
// thousand rows): ONE launch that adds the batch's raw moments straight into the running
// accumulators — sum[d] += column sums, gram[d][d] += X^T X (both triangles) — optionally of only
// the rows whose category matches (mask[row] == mask_value: the pipeline's per-category boolean
// mask, applied while loading instead of by an index_select beforehand).  One CTA per 64 x 64 tile
// on or above the diagonal; every output element has exactly one writer, so the accumulation is
// deterministic without atomics.  The reference finalises and Chan-merges a d x d covariance per
// batch (data.py:37-47,77-94); here nothing is finalised until the statistics are read.
constexpr int kST = 64;   // tile edge of the small-batch kernel: 36 CTAs at d = 512 (the batch is tiny, spread the d x d update)
template <typename T>
__global__ void __launch_bounds__(256)
moments_small_kernel(const T* __restrict__ X, long long n, int d, long long ld, const int32_t* __restrict__ mask,
                     int mask_value, int nt, double* __restrict__ sum, double* __restrict__ gram) {
  // the operand stages and, after the k loop, the transposed tile of the mirrored update share one buffer
  __shared__ __align__(16) double buf[kST * (kST + 1)];
  double (*As)[kST] = reinterpret_cast<double (*)[kST]>(buf);
  double (*Bs)[kST] = reinterpret_cast<double (*)[kST]>(buf + kGK * kST);
  double (*Tr)[kST + 1] = reinterpret_cast<double (*)[kST + 1]>(buf);
  int tp = blockIdx.x, ti = 0;
  while (tp >= nt - ti) { tp -= nt - ti; ++ti; }
  const int tj = ti + tp;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  double csum = 0.0;   // threads 0..63 of a diagonal tile: column sum of column ti*64 + tid
  for (long long r = 0; r < n; r += kGK) {
    for (int e = tid; e < kGK * kST; e += 256) {
      const int rr = e / kST, cc = e % kST;
      const long long row = r + rr;
      const bool live = row < n && (mask == nullptr || mask[row] == mask_value);
      const int ca = ti * kST + cc, cb = tj * kST + cc;
      As[rr][cc] = (live && ca < d) ? static_cast<double>(X[row * ld + ca]) : 0.0;
      Bs[rr][cc] = (live && cb < d) ? static_cast<double>(X[row * ld + cb]) : 0.0;
    }
    __syncthreads();
    if (ti == tj && tid < kST) {
#pragma unroll
      for (int kk = 0; kk < kGK; ++kk) csum += As[kk][tid];
    }
#pragma unroll
    for (int kk = 0; kk < kGK; ++kk) {
      const double2 a0 = *reinterpret_cast<const double2*>(&As[kk][ty * 4]);
      const double2 a1 = *reinterpret_cast<const double2*>(&As[kk][ty * 4 + 2]);
      const double2 b0 = *reinterpret_cast<const double2*>(&Bs[kk][tx * 4]);
      const double2 b1 = *reinterpret_cast<const double2*>(&Bs[kk][tx * 4 + 2]);
      const double a[4] = {a0.x, a0.y, a1.x, a1.y};
      const double b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = ti * kST + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = tj * kST + tx * 4 + j;
      if (gi < d && gj < d) gram[static_cast<long long>(gi) * d + gj] += acc[i][j];
    }
  }
  if (ti != tj) {   // mirror of an off-diagonal tile, written row-wise (the k loop ended on a barrier)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Tr[tx * 4 + j][ty * 4 + i] = acc[i][j];
    __syncthreads();
    for (int e = tid; e < kST * kST; e += 256) {
      const int rr = e / kST, cc = e % kST;
      const int gi = tj * kST + rr, gj = ti * kST + cc;
      if (gi < d && gj < d) gram[static_cast<long long>(gi) * d + gj] += Tr[rr][cc];
    }
  }
  if (ti == tj && tid < kST && ti * kST + tid < d) sum[ti * kST + tid] += csum;
}

// gram[i][j] += sum over slabs of the upper-tile partials, mirrored to the lower
// triangle (tile granularity: inside a diagonal tile both halves were computed).
__global__ void gram_reduce_kernel(const double* __restrict__ partial, int n_slabs, int d,
                                   double* __restrict__ gram) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e >= static_cast<long long>(d) * d) return;
  const int i = static_cast<int>(e / d), j = static_cast<int>(e % d);
  const bool upper = (i / kGT) <= (j / kGT);
  const long long src = upper ? e : static_cast<long long>(j) * d + i;
  double s = 0.0;
  for (int sl = 0; sl < n_slabs; ++sl) s += partial[static_cast<long long>(sl) * d * d + src];
  gram[e] += s;
}

template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ X, long long n, int d, long long ld, int n_slabs,
              double* __restrict__ partial /* [n_slabs][d] */) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int slab = blockIdx.y;
  if (col >= d) return;
  const long long rows_per = (n + n_slabs - 1) / n_slabs;
  const long long r0 = slab * rows_per;
  long long r1 = r0 + rows_per;
  if (r1 > n) r1 = n;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  long long r = r0;
  for (; r + 3 < r1; r += 4) {
    s0 += static_cast<double>(X[r * ld + col]);
    s1 += static_cast<double>(X[(r + 1) * ld + col]);
    s2 += static_cast<double>(X[(r + 2) * ld + col]);
    s3 += static_cast<double>(X[(r + 3) * ld + col]);
  }
  for (; r < r1; ++r) s0 += static_cast<double>(X[r * ld + col]);
  partial[static_cast<long long>(slab) * d + col] = (s0 + s1) + (s2 + s3);
}

__global__ void colsum_reduce_kernel(const double* __restrict__ partial, int n_slabs, int d,
                                     double* __restrict__ sum) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= d) return;
  double s = 0.0;
  for (int sl = 0; sl < n_slabs; ++sl) s += partial[static_cast<long long>(sl) * d + col];
  sum[col] += s;
}

// data.py:39-44: mean = sum/n; cov = (G - n mu mu^T)/(n-1); n == 1 -> zeros.
__global__ void cov_finalize_kernel(long long n, int d, const double* __restrict__ sum,
                                    const double* __restrict__ gram, double* __restrict__ mean,
                                    double* __restrict__ cov) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const double nn = static_cast<double>(n);
  if (e < d) mean[e] = sum[e] / nn;
  if (e >= static_cast<long long>(d) * d) return;
  const int i = static_cast<int>(e / d), j = static_cast<int>(e % d);
  if (n <= 1) { cov[e] = 0.0; return; }
  const double mi = sum[i] / nn, mj = sum[j] / nn;
  cov[e] = (gram[e] - nn * mi * mj) / (nn - 1.0);
}

// data.py:77-94 _update_stats, in place in (mean1, cov1).
__global__ void stats_merge_kernel(int d, double n1, double* __restrict__ mean1, double* __restrict__ cov1,
                                   double n2, const double* __restrict__ mean2,
                                   const double* __restrict__ cov2, double* __restrict__ new_mean) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const double n_total = n1 + n2;
  if (e < d) new_mean[e] = (n1 * mean1[e] + n2 * mean2[e]) / n_total;
  if (e >= static_cast<long long>(d) * d) return;
  const int i = static_cast<int>(e / d), j = static_cast<int>(e % d);
  const double di = mean1[i] - mean2[i], dj = mean1[j] - mean2[j];
  const double w_self = (n1 - 1.0) / (n_total - 1.0);
  const double w_other = (n2 - 1.0) / (n_total - 1.0);
  const double w_diff = (n1 * n2 / n_total) / (n_total - 1.0);
  cov1[e] = w_self * cov1[e] + w_other * cov2[e] + w_diff * (di * dj);
}

static int cov_slabs(long long n, int d) {
  const int nt = (d + kGT - 1) / kGT;
  const int tiles = nt * (nt + 1) / 2;
  long long s = (4ll * 148 + tiles - 1) / tiles;
  const long long max_s = (n + 255) / 256;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return static_cast<int>(s);
}

// fp32 inputs with enough rows go through the exact integer tensor-core path (cov_tc.cu);
// small batches (the 32-row streaming adds) and fp64 inputs stay on the FP64 pipe.
// Option cov_dfma (AMB_COV=dfma) forces the FP64-pipe kernel.
constexpr long long kCovTcMinRows = 4096;
static bool use_cov_tc(int dtype, long long n) {
  if (dtype != AMB_F32 || n < kCovTcMinRows) return false;
  return option(kOptCovDfma) == 0;
}

}  // namespace amb

using namespace amb;

extern "C" {

size_t amb_cov_ws_bytes(long long n, int d) {
  if (n <= 0 || d <= 0) return 0;
  const int s = cov_slabs(n, d);
  const size_t dfma = static_cast<size_t>(s) * (static_cast<size_t>(d) * d + d) * 8 + 256;
  const size_t tc = n >= kCovTcMinRows ? cov_tc_ws_bytes(n, d) : 0;
  return dfma > tc ? dfma : tc;
}

int amb_cov_accumulate(int dev, amb_stream_t stream, const void* X, int dtype, long long n, int d,
                       long long ld, double* sum, double* gram, void* ws, size_t ws_bytes) {
  if (!X || !sum || !gram || n <= 0 || d <= 0 || ld < d)
    return set_error(AMB_ERR_ARG, "amb_cov_accumulate: bad argument (n=%lld d=%d ld=%lld)", n, d, ld);
  if (dtype != AMB_F32 && dtype != AMB_F64) return set_error(AMB_ERR_ARG, "amb_cov_accumulate: bad dtype");
  const size_t need = amb_cov_ws_bytes(n, d);
  if (!ws || ws_bytes < need) return set_error(AMB_ERR_WS, "amb_cov_accumulate: workspace %zu < %zu", ws_bytes, need);
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (use_cov_tc(dtype, n)) return cov_tc_accumulate(st, dev, static_cast<const float*>(X), n, d, ld, sum, gram, ws);
  const int slabs = cov_slabs(n, d);
  const int nt = (d + kGT - 1) / kGT;
  double* gpart = static_cast<double*>(ws);
  double* spart = gpart + static_cast<size_t>(slabs) * d * d;
  dim3 grid(nt * (nt + 1) / 2, slabs);
  dim3 sgrid((d + 255) / 256, slabs);
  int rc;
  if (dtype == AMB_F32) {
    gram_tile_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(X), n, d, ld, nt, slabs, gpart);
    if ((rc = check_launch("gram_tile_kernel"))) return rc;
    colsum_kernel<float><<<sgrid, 256, 0, st>>>(static_cast<const float*>(X), n, d, ld, slabs, spart);
  } else {
    gram_tile_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(X), n, d, ld, nt, slabs, gpart);
    if ((rc = check_launch("gram_tile_kernel"))) return rc;
    colsum_kernel<double><<<sgrid, 256, 0, st>>>(static_cast<const double*>(X), n, d, ld, slabs, spart);
  }
  if ((rc = check_launch("colsum_kernel"))) return rc;
  const long long dd = static_cast<long long>(d) * d;
  gram_reduce_kernel<<<static_cast<unsigned>((dd + 255) / 256), 256, 0, st>>>(gpart, slabs, d, gram);
  if ((rc = check_launch("gram_reduce_kernel"))) return rc;
  colsum_reduce_kernel<<<(d + 255) / 256, 256, 0, st>>>(spart, slabs, d, sum);
  return check_launch("colsum_reduce_kernel");
}

int amb_cov_accumulate_masked(int dev, amb_stream_t stream, const void* X, int dtype, long long n, int d,
                              long long ld, const int32_t* mask, int mask_value, double* sum, double* gram) {
  if (!X || !sum || !gram || n <= 0 || d <= 0 || ld < d)
    return set_error(AMB_ERR_ARG, "amb_cov_accumulate_masked: bad argument (n=%lld d=%d ld=%lld)", n, d, ld);
  if (dtype != AMB_F32 && dtype != AMB_F64) return set_error(AMB_ERR_ARG, "amb_cov_accumulate_masked: bad dtype");
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nt = (d + kST - 1) / kST;
  const unsigned grid = static_cast<unsigned>(nt * (nt + 1) / 2);
  if (dtype == AMB_F32)
    moments_small_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(X), n, d, ld, mask, mask_value, nt, sum, gram);
  else
    moments_small_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(X), n, d, ld, mask, mask_value, nt, sum, gram);
  return check_launch("moments_small_kernel");
}

int amb_cov_finalize(int dev, amb_stream_t stream, long long n, int d, const double* sum,
                     const double* gram, double* mean, double* cov) {
  if (!sum || !gram || !mean || !cov || n <= 0 || d <= 0) return set_error(AMB_ERR_ARG, "amb_cov_finalize: bad argument");
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  const long long dd = static_cast<long long>(d) * d;
  cov_finalize_kernel<<<static_cast<unsigned>((dd + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n, d, sum, gram, mean, cov);
  return check_launch("cov_finalize_kernel");
}

int amb_stats_merge(int dev, amb_stream_t stream, int d, long long n1, double* mean1, double* cov1,
                    long long n2, const double* mean2, const double* cov2, double* scratch_mean) {
  if (!mean1 || !cov1 || !mean2 || !cov2 || !scratch_mean || d <= 0 || n1 <= 0 || n2 <= 0)
    return set_error(AMB_ERR_ARG, "amb_stats_merge: bad argument");
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long dd = static_cast<long long>(d) * d;
  stats_merge_kernel<<<static_cast<unsigned>((dd + 255) / 256), 256, 0, st>>>(
      d, static_cast<double>(n1), mean1, cov1, static_cast<double>(n2), mean2, cov2, scratch_mean);
  int rc = check_launch("stats_merge_kernel");
  if (rc) return rc;
  return check_cuda(cudaMemcpyAsync(mean1, scratch_mean, static_cast<size_t>(d) * 8, cudaMemcpyDeviceToDevice, st),
                    "memcpy(mean)");
}

}  // extern "C"

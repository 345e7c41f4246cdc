// Kernel distance (metrics/kd.py) on the pair engine: gather the subset rows into
// packed operands, run the three kernel blocks of every subset as engine
// problems with the polynomial / RBF epilogue, and reduce to MMD^2 per subset.
#include "engine_launch.cuh"

namespace amb {

// gather tables: packed row (q*S + s)*mp + i  <-  idx[s][q][i]  (i < m), else -1
// problem tables (row-block units of 128): 3s+0 = (f1,f1), 3s+1 = (f2,f2), 3s+2 = (f1,f2)
__global__ void kd_tables_kernel(const int32_t* __restrict__ idx, int S, int m, int mp,
                                 int* __restrict__ gather, int* __restrict__ a_rb0, int* __restrict__ b_rb0) {
  const long long total = 2ll * S * mp;
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(t % mp);
    const long long qs = t / mp;
    const int s = static_cast<int>(qs % S), q = static_cast<int>(qs / S);
    gather[t] = i < m ? idx[(static_cast<long long>(s) * 2 + q) * m + i] : -1;
  }
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    const int rb1 = s * (mp / kBlockRows), rb2 = (S + s) * (mp / kBlockRows);
    a_rb0[3 * s + 0] = rb1; b_rb0[3 * s + 0] = rb1;
    a_rb0[3 * s + 1] = rb2; b_rb0[3 * s + 1] = rb2;
    a_rb0[3 * s + 2] = rb1; b_rb0[3 * s + 2] = rb2;
  }
}

// One block: per-subset MMD^2 (kd.py:38-83: the "unbiased" estimator kernel_mmd2 uses, or the
// "biased" / "u-statistic" ones), then mean and population std over subsets (kd.py:189-192).
// Fixed summation order.
__global__ void kd_finalize_kernel(const double* __restrict__ partial, int S, int per_problem, int m, int mmd_est,
                                   double* __restrict__ mmd2_out, double* __restrict__ stats_out) {
  extern __shared__ double s_mmd[];
  const int est = mmd_est & 3;
  const bool unit_diag = (mmd_est & AMB_MMD_UNIT_DIAGONAL) != 0;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    double off[3], dg[3];   // off-diagonal and diagonal sums of K_XX, K_YY, K_XY
    for (int b = 0; b < 3; ++b) {
      double acc = 0.0, dacc = 0.0;
      const double* p = partial + (static_cast<long long>(3 * s + b)) * per_problem * 2;
      for (int e = 0; e < per_problem; ++e) { acc += p[2 * e]; dacc += p[2 * e + 1]; }
      off[b] = acc;
      dg[b] = dacc;
    }
    const double md = static_cast<double>(m);
    // kd.py:50-68: Kt_*_sum = (K.sum(axis=1) - diag).sum(), with diag = 1 under unit_diagonal
    const double kt_xx = unit_diag ? off[0] + dg[0] - md : off[0];
    const double kt_yy = unit_diag ? off[1] + dg[1] - md : off[1];
    const double sd_x = unit_diag ? md : dg[0], sd_y = unit_diag ? md : dg[1];
    const double k_xy = off[2] + dg[2];
    double v;
    if (est == AMB_MMD_BIASED) {                                   // kd.py:70-75
      v = (kt_xx + sd_x) / (md * md) + (kt_yy + sd_y) / (md * md) - 2.0 * k_xy / (md * md);
    } else {
      v = (kt_xx + kt_yy) / (md * (md - 1.0));                     // kd.py:77
      if (est == AMB_MMD_UNBIASED) v -= 2.0 * k_xy / (md * md);    // kd.py:79
      else v -= 2.0 * (k_xy - dg[2]) / (md * (md - 1.0));          // kd.py:81 (K_XY_sum - trace(K_XY))
    }
    s_mmd[s] = v;
    if (mmd2_out) mmd2_out[s] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0 && stats_out) {
    double mean = 0.0;
    for (int s = 0; s < S; ++s) mean += s_mmd[s];
    mean /= S;
    double var = 0.0;
    for (int s = 0; s < S; ++s) var += (s_mmd[s] - mean) * (s_mmd[s] - mean);
    stats_out[0] = mean;
    stats_out[1] = sqrt(var / S);
  }
}

struct KdWs {
  void* packed;
  int* gather;
  int* a_rb0;
  int* b_rb0;
  double* partial;
  size_t bytes;
  int mp;
};
static KdWs kd_ws(void* ws, int S, int m, int d) {
  KdWs w;
  w.mp = static_cast<int>(round_up_ll(m, kRowPad));
  const long long rows = 2ll * S * w.mp;
  uint8_t* b = static_cast<uint8_t*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = b ? b + off : nullptr; off += static_cast<size_t>(round_up_ll(bytes, 256)); return p; };
  w.packed = take(packed_layout(rows, d).bytes);
  w.gather = reinterpret_cast<int*>(take(static_cast<size_t>(rows) * 4));
  w.a_rb0 = reinterpret_cast<int*>(take(static_cast<size_t>(3) * S * 4));
  w.b_rb0 = reinterpret_cast<int*>(take(static_cast<size_t>(3) * S * 4));
  w.partial = reinterpret_cast<double*>(take(static_cast<size_t>(3) * S * (w.mp / kTileM) * kEpiWarps * 2 * 8));
  w.bytes = off;
  return w;
}

}  // namespace amb

using namespace amb;

extern "C" {

size_t amb_kd_ws_bytes(int S, int m, int d) {
  if (S <= 0 || m <= 0 || d <= 0) return 0;
  return kd_ws(nullptr, S, m, d).bytes;
}

int amb_kd_subsets(int dev, amb_stream_t stream, const void* F1, long long n1, long long ld1,
                   const void* F2, long long n2, long long ld2, int d, int dtype, const int32_t* idx,
                   int S, int m, int kernel_type, double gamma, double coef0, int degree,
                   double sigma, int mmd_est, double* mmd2_out, double* stats_out, void* ws, size_t ws_bytes) {
  if (!F1 || !F2 || !idx || n1 <= 0 || n2 <= 0 || d <= 0 || S <= 0 || m <= 0 || ld1 < d || ld2 < d)
    return set_error(AMB_ERR_ARG, "amb_kd_subsets: bad argument");
  if (S > 4096) return set_error(AMB_ERR_ARG, "amb_kd_subsets: at most 4096 subsets");
  if (kernel_type != AMB_KERNEL_POLY && kernel_type != AMB_KERNEL_RBF)
    return set_error(AMB_ERR_ARG, "amb_kd_subsets: unknown kernel_type %d", kernel_type);
  if (kernel_type == AMB_KERNEL_POLY && degree < 0) return set_error(AMB_ERR_ARG, "amb_kd_subsets: degree < 0");
  if (kernel_type == AMB_KERNEL_RBF && !(sigma > 0)) return set_error(AMB_ERR_ARG, "amb_kd_subsets: sigma <= 0");
  if ((mmd_est & ~AMB_MMD_UNIT_DIAGONAL) < 0 || (mmd_est & ~AMB_MMD_UNIT_DIAGONAL) > AMB_MMD_USTAT)
    return set_error(AMB_ERR_ARG, "amb_kd_subsets: unknown mmd_est %d", mmd_est);
  KdWs w = kd_ws(ws, S, m, d);
  if (!ws || ws_bytes < w.bytes) return set_error(AMB_ERR_WS, "amb_kd_subsets: workspace %zu < %zu", ws_bytes, w.bytes);
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long rows = 2ll * S * w.mp;
  PackedPtrs p = packed_ptrs(w.packed, rows, d);
  int rc;
  kd_tables_kernel<<<256, 256, 0, st>>>(idx, S, m, w.mp, w.gather, w.a_rb0, w.b_rb0);
  if ((rc = check_launch("kd_tables_kernel"))) return rc;
  const long long half = static_cast<long long>(S) * w.mp;
  if ((rc = launch_pack(st, F1, dtype, ld1, d, n1, w.gather, half, 0, half, p.planes, p.plane_halfs,
                        p.kb_count, p.inv_scale, p.norm, p.rho, p.row_exp, p.cmin))) return rc;
  if ((rc = launch_pack(st, F2, dtype, ld2, d, n2, w.gather + half, half, half, half, p.planes, p.plane_halfs,
                        p.kb_count, p.inv_scale, p.norm, p.rho, p.row_exp, p.cmin))) return rc;

  EngineGeom g{};
  g.a_planes = p.planes;
  g.b_planes = p.planes;
  g.a_plane_halfs = p.plane_halfs;
  g.b_plane_halfs = p.plane_halfs;
  g.kb_count = p.kb_count;
  g.a_rb0 = w.a_rb0;
  g.b_rb0 = w.b_rb0;
  g.n_problems = 3 * S;
  g.n_rt = w.mp / kTileM;
  g.n_ct = w.mp / kTileN;
  g.n_split = 1;
  g.lbo_bytes = 128;
  g.sbo_bytes = 512;
  KdEpi epi{};
  epi.inv_a = p.inv_scale;
  epi.inv_b = p.inv_scale;
  epi.kernel_type = kernel_type;
  epi.gamma = gamma;
  epi.coef0 = coef0;
  epi.degree = degree;
  epi.rbf_scale = kernel_type == AMB_KERNEL_RBF ? -1.0 / (2.0 * sigma * sigma) : 0.0;
  epi.norm_a = p.norm;
  epi.norm_b = p.norm;
  epi.m_valid = m;
  epi.partial = w.partial;
  if ((rc = launch_engine(st, dev, g, epi, "pair_engine<kd>", 3.0 * S * static_cast<double>(m) * m))) return rc;
  kd_finalize_kernel<<<1, 256, static_cast<size_t>(S) * 8, st>>>(w.partial, S, g.n_rt * kEpiWarps, m, mmd_est, mmd2_out,
                                                                 stats_out);
  return check_launch("kd_finalize_kernel");
}

}  // extern "C"

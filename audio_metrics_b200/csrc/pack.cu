// Pack kernel: fp32/fp64 rows -> scaled fp16 hi/lo planes in the tensor-core
// shared-memory image (see packed.cuh), plus per-row inverse scale and norm.
// HBM-bound pre-pass: reads n*d*sizeof(T) once (rows are read twice, the second
// time from L1/L2), writes 4 bytes per element.
#include "internal.cuh"

namespace amb {

template <typename T>
__device__ __forceinline__ float to_scaled_hi_lo(T x, T scale, __half& hi, __half& lo);

template <>
__device__ __forceinline__ float to_scaled_hi_lo<float>(float x, float scale, __half& hi,
                                                        __half& lo) {
  float v = x * scale;  // exact: power-of-two scale
  hi = __float2half_rn(v);
  float rem = v - __half2float(hi);  // exact in fp32
  lo = __float2half_rn(rem);
  return __half2float(hi) + __half2float(lo);  // may round, only used for the norm via double below
}
template <>
__device__ __forceinline__ float to_scaled_hi_lo<double>(double x, double scale, __half& hi,
                                                         __half& lo) {
  double v = x * scale;
  hi = __double2half(v);
  double rem = v - static_cast<double>(__half2float(hi));
  lo = __double2half(rem);
  return 0.f;
}

// Thread mapping of both kernels: one warp owns one 8-row group (lane = c*8 + r: row r of the
// group, k-chunk c of each 32-wide k block), so every store instruction of the warp writes 512
// contiguous bytes of a plane.
//
// The power-of-two scale is shared by the 256 packed rows of a column tile of the pair engine
// (fp16 keeps 11 significant bits over 30 binades, so rows far below the tile's maximum lose
// nothing): the engine's epilogues then apply ONE scale per tile instead of one per column.
// row_exp_kernel writes each row's binary exponent, pack_rows_kernel takes the tile maximum.
constexpr int kNoExp = -(1 << 30);

template <typename T>
__device__ __forceinline__ const T* source_row(const T* src, long long ld, long long n_src_rows, const int* gather,
                                               long long n_valid, long long rel_row) {
  long long src_row = -1;
  if (rel_row < n_valid) {
    src_row = gather ? static_cast<long long>(gather[rel_row]) : rel_row;
    if (src_row >= n_src_rows) src_row = -1;
  }
  return src_row >= 0 ? src + src_row * ld : nullptr;
}

template <typename T>
__global__ void __launch_bounds__(256)
row_exp_kernel(const T* __restrict__ src, long long ld, int d, long long n_src_rows, const int* __restrict__ gather,
               long long n_valid, long long row0, long long n_rows_out, int kb_count, int* __restrict__ row_exp) {
  const int lane = threadIdx.x & 31;
  const int r = lane & 7, c = lane >> 3;
  const long long group = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (group * 8 >= n_rows_out) return;
  const T* rowp = source_row(src, ld, n_src_rows, gather, n_valid, group * 8 + r);
  T amax = 0;
  if (rowp) {
    for (int kb = 0; kb < kb_count; ++kb) {
      const int k0 = kb * kBlockK + c * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = k0 + e;
        T v = k < d ? rowp[k] : T(0);
        v = v < 0 ? -v : v;
        amax = v > amax ? v : amax;
      }
    }
  }
  {
    T o = __shfl_xor_sync(0xffffffffu, amax, 8);
    amax = o > amax ? o : amax;
    o = __shfl_xor_sync(0xffffffffu, amax, 16);
    amax = o > amax ? o : amax;
  }
  int ex = kNoExp;
  if (amax > 0 && amax < T(3.0e38)) frexp(static_cast<double>(amax), &ex);  // amax = f*2^ex, f in [.5,1)
  if (c == 0) row_exp[row0 + group * 8 + r] = ex;
}

template <typename T>
__global__ void __launch_bounds__(256)
pack_rows_kernel(const T* __restrict__ src, long long ld, int d, long long n_src_rows,
                 const int* __restrict__ gather,  // nullable; <0 => zero row
                 long long n_valid,               // rows >= n_valid (pre-gather index) are padding
                 long long row0, long long n_rows_out, __half* __restrict__ planes,
                 long long plane_halfs, int kb_count, float* __restrict__ inv_scale,
                 float* __restrict__ norm, float* __restrict__ rho, const int* __restrict__ row_exp) {
  const int lane = threadIdx.x & 31;
  const int r = lane & 7, c = lane >> 3;
  const long long group = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (group * 8 >= n_rows_out) return;
  const long long out_row = row0 + group * 8 + r;   // packed row this lane writes
  const T* rowp = source_row(src, ld, n_src_rows, gather, n_valid, group * 8 + r);

  // scale exponent of the 256-row tile this group belongs to
  int ex = kNoExp;
  {
    const long long tile0 = (out_row / kRowPad) * kRowPad;
#pragma unroll
    for (int i = 0; i < kRowPad / 32; ++i) ex = max(ex, row_exp[tile0 + lane + 32 * i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ex = max(ex, __shfl_xor_sync(0xffffffffu, ex, o));
    if (ex == kNoExp) ex = 0;
  }
  int sh = 15 - ex;                       // tile maximum * 2^sh in [2^14, 2^15)
  sh = sh > 120 ? 120 : (sh < -120 ? -120 : sh);
  const T scale = static_cast<T>(ldexp(1.0, sh));
  const double inv = ldexp(1.0, -sh);

  // split, store, norm
  double nrm = 0.0, res = 0.0;
  const long long rb = out_row / kBlockRows;
  const int rin = static_cast<int>(out_row % kBlockRows);
  const long long chunk_row_off = ((static_cast<long long>(rin >> 3) * 4 + c) * 8 + (rin & 7)) * 8;
  for (int kb = 0; kb < kb_count; ++kb) {
    const int k0 = kb * kBlockK + c * 8;
    alignas(16) __half h[8];
    alignas(16) __half l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = k0 + e;
      T v = (rowp && k < d) ? rowp[k] : T(0);
      to_scaled_hi_lo<T>(v, scale, h[e], l[e]);
      const double xt = static_cast<double>(__half2float(h[e])) + static_cast<double>(__half2float(l[e]));
      nrm += xt * xt;
      const double rm = static_cast<double>(v) * static_cast<double>(scale) - static_cast<double>(__half2float(h[e]));
      res += rm * rm;   // exact remainder of the hi plane (fp64 product by a power of two is exact)
    }
    const long long off = (rb * kb_count + kb) * kChunkHalfs + chunk_row_off;
    *reinterpret_cast<uint4*>(planes + off) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(planes + plane_halfs + off) = *reinterpret_cast<const uint4*>(l);
  }
  nrm += __shfl_xor_sync(0xffffffffu, nrm, 8);
  nrm += __shfl_xor_sync(0xffffffffu, nrm, 16);
  res += __shfl_xor_sync(0xffffffffu, res, 8);
  res += __shfl_xor_sync(0xffffffffu, res, 16);
  if (c == 0) {
    inv_scale[out_row] = static_cast<float>(inv);   // the tile's scale, also on padding rows
    if (rowp) {
      norm[out_row] = static_cast<float>(nrm * inv * inv);
      rho[out_row] = __double2float_ru(sqrt(res) * inv * (1.0 + 1e-7));
    } else {
      norm[out_row] = __int_as_float(0x7f800000);  // +inf: padding row
      rho[out_row] = 0.0f;
    }
  }
}

// cmin[c] = smallest squared norm among packed rows [32 c, 32 c + 32): lets the radii / count
// epilogues bound  |y|^2 - 2<x,y>  for a whole 32-column chunk from the raw accumulators alone.
__global__ void chunk_min_norm_kernel(const float* __restrict__ norm, long long row0, long long n_rows_out,
                                      float* __restrict__ cmin) {
  const long long c = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c * 32 >= n_rows_out) return;
  float v = norm[row0 + c * 32 + lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) cmin[row0 / 32 + c] = v;
}

int launch_pack(cudaStream_t stream, const void* src, int dtype, long long ld, int d,
                long long n_src_rows, const int* gather, long long n_valid, long long row0,
                long long n_rows_out, __half* planes, long long plane_halfs, int kb_count,
                float* inv_scale, float* norm, float* rho, int* row_exp, float* cmin) {
  if (n_rows_out <= 0) return 0;
  if (row0 % kRowPad != 0 || n_rows_out % kRowPad != 0)
    return set_error(AMB_ERR_ARG, "pack: row range must be whole 256-row tiles");
  const long long groups = n_rows_out / 8;
  const int threads = 256;
  const long long blocks = (groups * 32 + threads - 1) / threads;
  if (dtype == AMB_F32) {
    row_exp_kernel<float><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        static_cast<const float*>(src), ld, d, n_src_rows, gather, n_valid, row0, n_rows_out, kb_count, row_exp);
    int rc = check_launch("row_exp_kernel");
    if (rc) return rc;
    pack_rows_kernel<float><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        static_cast<const float*>(src), ld, d, n_src_rows, gather, n_valid, row0, n_rows_out,
        planes, plane_halfs, kb_count, inv_scale, norm, rho, row_exp);
  } else if (dtype == AMB_F64) {
    row_exp_kernel<double><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        static_cast<const double*>(src), ld, d, n_src_rows, gather, n_valid, row0, n_rows_out, kb_count, row_exp);
    int rc = check_launch("row_exp_kernel");
    if (rc) return rc;
    pack_rows_kernel<double><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        static_cast<const double*>(src), ld, d, n_src_rows, gather, n_valid, row0, n_rows_out,
        planes, plane_halfs, kb_count, inv_scale, norm, rho, row_exp);
  } else {
    return set_error(AMB_ERR_ARG, "pack: dtype must be AMB_F32 or AMB_F64");
  }
  int rc = check_launch("pack_rows_kernel");
  if (rc) return rc;
  chunk_min_norm_kernel<<<static_cast<unsigned>((n_rows_out + 255) / 256), 256, 0, stream>>>(norm, row0, n_rows_out, cmin);
  return check_launch("chunk_min_norm_kernel");
}

}  // namespace amb

// Tensor-core covariance moments: Gram matrix X^T X and column sums of an fp32 matrix
// through EXACT integer tensor-core arithmetic (tcgen05.mma kind::i8, int32 accumulators
// in tensor memory), for AudioMetricsData.add / recompute_stats (reference data.py:37-58).
//
// Why integers: the fp32-accumulate MMA kinds truncate (profiles/r01_engine_check.log), and a
// Gram matrix over 10^5..10^6 samples followed by the subtraction of n mu mu^T needs every bit.
// So each column (feature) k is put on a fixed-point grid, q_ik = rint(x_ik 2^e_k) with
// |q| <= 2^30 (e_k from the column's absolute maximum), and q is cut into four balanced
// base-256 digits  q = a3 2^24 + a2 2^16 + a1 2^8 + a0,  a in [-128, 127]  (Ozaki-style
// slicing).  Products of digits accumulate exactly in int32; all sixteen digit pairs are
// formed, the pairs of equal weight 2^(8w), w = s + t = 0..6, share one of seven accumulators
// and the epilogue sums  D_w 2^(8w)  in fp64.  The result is the exact Gram matrix of the
// quantised data (grid 2^-30 of each column's maximum: entries within 2^-6 of the maximum
// keep every fp32 bit) up to fp64 rounding, positive semi-definite like the FP64-pipe
// result, and the column sums are exact integers.
//
// Passes over a chunk of <= kChunkRows samples:
//   col_absmax    HBM read of X                         -> per-column exponent
//   slice_pack    HBM read of X, 4 B/element written    -> int8 digit planes in the K-major
//                 core-matrix image the MMA reads (samples are the K dimension, so this pass
//                 is also the transpose) + per-block column sums of q
//   syrk_i8       warp-specialised tcgen05 kernel: bulk-copy producer / MMA issuer / 4
//                 epilogue warps; work item = (128 x 64 tile on or above the diagonal, K split)
//   reduce        sums the K-split partials, rescales by 2^-(e_k+e_l), mirrors, adds to gram
#include "internal.cuh"
#include "tc05.cuh"

namespace amb {

constexpr int kFB = 128;                      // features per block (MMA M)
constexpr int kFN = 64;                       // features per output tile column block (MMA N)
constexpr int kSK = 64;                       // samples per K block: one 8 KiB int8 chunk
constexpr int kSliceChunk = kFB * kSK;        // bytes (128 features x 64 samples)
constexpr int kDigits = 4;
constexpr int kClasses = 2 * kDigits - 1;     // weights 2^(8w), w = 0..6
constexpr int kSyrkStages = 4;
constexpr int kSyrkStageBytes = kDigits * kSliceChunk + kDigits * (kSliceChunk / 2);   // A3..A0 | B3..B0 (half chunks)
constexpr int kSyrkThreads = 192;
constexpr int kQBits = 30;                    // |q| <= 2^30
constexpr int kMaxKbPerItem = 512;            // 512*64 samples * 49152 (largest class per sample) < 2^31
constexpr long long kChunkRows = 262144;      // samples per pass (bounds the digit-plane workspace: 4 B per element);
                                              // a 200k-row set is one pass: one slice, one SYRK, one reduce launch

// ------------------------------------------------------------------ column scale
__global__ void __launch_bounds__(256)
col_absmax_kernel(const float* __restrict__ X, long long r0, long long r1, int d, long long ld,
                  float* __restrict__ maxabs /* [d_pad], zeroed */) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= d) return;
  const long long rows = r1 - r0;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  long long a = r0 + blockIdx.y * per, b = a + per;
  if (b > r1) b = r1;
  float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
  long long r = a;
  for (; r + 3 < b; r += 4) {
    m0 = fmaxf(m0, fabsf(X[r * ld + col]));
    m1 = fmaxf(m1, fabsf(X[(r + 1) * ld + col]));
    m2 = fmaxf(m2, fabsf(X[(r + 2) * ld + col]));
    m3 = fmaxf(m3, fabsf(X[(r + 3) * ld + col]));
  }
  for (; r < b; ++r) m0 = fmaxf(m0, fabsf(X[r * ld + col]));
  const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));   // fmaxf drops NaNs
  atomicMax(reinterpret_cast<int*>(maxabs) + col, __float_as_int(m));   // m >= 0: int order == float order
}

// exps[k] = e_k with max_k 2^e_k in [2^(kQBits-1), 2^kQBits); 0 for an all-zero / non-finite column.
__global__ void col_exp_kernel(const float* __restrict__ maxabs, int d_pad, int* __restrict__ exps) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= d_pad) return;
  const float m = maxabs[k];
  int e = 0;
  if (m > 0.f && m < 3.0e38f) {
    e = kQBits - 1 - ilogbf(m);
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
  }
  exps[k] = e;
}

// ------------------------------------------------------------------ digit planes
// planes[s][fb][kb] is an 8 KiB chunk: 128 features x 64 samples of digit s (0 = lowest), stored as the
// K-major no-swizzle core-matrix image  [f/8 (16)][k/16 (4)][f%8 (8)][k%16 (16)]  bytes
// (LBO = 128 B to the next 16 samples, SBO = 512 B to the next 8 features).
// Block = (feature block, a strided set of K blocks); warp w: sample group c = w&3 of the K
// block, features [32*(w>>2), +32).  colsum_part[blockIdx.x][c][f] receives the block's sum of q.
__global__ void __launch_bounds__(512)
slice_pack_kernel(const float* __restrict__ X, long long r0, long long r1, int d, long long ld,
                  const int* __restrict__ exps, int8_t* __restrict__ planes, long long plane_bytes, int nkb,
                  double* __restrict__ colsum_part /* [gridDim.x][4][d_pad] */, int d_pad) {
  __shared__ float tile[kSK][kFB + 1];
  const int fb = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = warp & 3;
  const int f = (warp >> 2) * 32 + lane;
  const int col = fb * kFB + f;
  const float scale = col < d ? exp2f(static_cast<float>(exps[col])) : 0.f;
  long long s_acc = 0;
  for (int kb = blockIdx.x; kb < nkb; kb += gridDim.x) {
    const long long row_base = r0 + static_cast<long long>(kb) * kSK;
    __syncthreads();
    for (int e = tid; e < kSK * kFB; e += 512) {
      const int k = e >> 7, ff = e & 127;
      const long long row = row_base + k;
      const int cc = fb * kFB + ff;
      tile[k][ff] = (row < r1 && cc < d) ? X[row * ld + cc] : 0.f;
    }
    __syncthreads();
    alignas(16) int8_t dg[kDigits][16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float v = tile[c * 16 + j][f] * scale;        // power-of-two scale: exact
      int q = __float2int_rn(v);                          // NaN -> 0; the clamp covers +-inf
      q = q > (1 << kQBits) ? (1 << kQBits) : (q < -(1 << kQBits) ? -(1 << kQBits) : q);
      s_acc += q;
#pragma unroll
      for (int s = 0; s < kDigits - 1; ++s) {
        const int dd = ((q + 128) & 255) - 128;           // balanced digit in [-128, 127]
        q = (q - dd) >> 8;
        dg[s][j] = static_cast<int8_t>(dd);
      }
      dg[kDigits - 1][j] = static_cast<int8_t>(q);        // |top digit| <= 64
    }
    const long long chunk = (static_cast<long long>(fb) * nkb + kb) * kSliceChunk;
    const int off = ((f >> 3) * 4 + c) * 128 + (f & 7) * 16;
#pragma unroll
    for (int s = 0; s < kDigits; ++s)
      *reinterpret_cast<uint4*>(planes + s * plane_bytes + chunk + off) = *reinterpret_cast<const uint4*>(dg[s]);
  }
  colsum_part[(static_cast<long long>(blockIdx.x) * 4 + c) * d_pad + col] = static_cast<double>(s_acc);
}

// ------------------------------------------------------------------ SYRK on tcgen05 (kind::i8)
__device__ __forceinline__ void mma_i8_ss(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// signed 8-bit A/B (K-major both), int32 D
__host__ __device__ constexpr uint32_t make_idesc_i8(int m, int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

struct SyrkSmem {
  uint64_t full[kSyrkStages];
  uint64_t empty[kSyrkStages];
  uint64_t acc_full;
  uint32_t tmem_base;
  uint32_t pad_;
};
constexpr size_t kSyrkSmemBytes = size_t(kSyrkStages) * kSyrkStageBytes + sizeof(SyrkSmem);

struct SyrkGeom {
  const int8_t* planes;
  long long plane_bytes;
  int nfb, nkb;
  int n_tiles;       // nfb (nfb + 1): tiles (ti, tj) of 128 x 64 with tj >= 2 ti
  int n_ksplit;
  double* partial;   // [n_ksplit][d_pad][d_pad]
  int d_pad;
};

// 16 consecutive fp32/int32 columns of this warp's 32 lanes
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(kSyrkThreads, 1)
syrk_i8_kernel(const SyrkGeom g) {
  extern __shared__ __align__(1024) uint8_t smem_buf[];
  uint8_t* stage_base = smem_buf;
  SyrkSmem* sh = reinterpret_cast<SyrkSmem*>(smem_buf + size_t(kSyrkStages) * kSyrkStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item -> (ti, tj >= 2 ti, K range)
  const int item = blockIdx.x;
  int tp = item % g.n_tiles, ti = 0;
  while (tp >= 2 * (g.nfb - ti)) { tp -= 2 * (g.nfb - ti); ++ti; }
  const int tj = 2 * ti + tp;                   // 64-feature column block
  const int split = item / g.n_tiles;
  const int kb0 = static_cast<int>(static_cast<long long>(g.nkb) * split / g.n_ksplit);
  const int kb1 = static_cast<int>(static_cast<long long>(g.nkb) * (split + 1) / g.n_ksplit);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSyrkStages; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], 1);
    }
    mbar_init(&sh->acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&sh->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ producer
    int s = 0;
    uint32_t ph = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&sh->empty[s], ph ^ 1);
      if (elect_one()) {
        uint8_t* st = stage_base + size_t(s) * kSyrkStageBytes;
        mbar_expect_tx(&sh->full[s], kSyrkStageBytes);
        const long long ca = (static_cast<long long>(ti) * g.nkb + kb) * kSliceChunk;
        // the 64 features of column block tj are one half (row groups 0-7 or 8-15) of a chunk
        const long long cb = (static_cast<long long>(tj >> 1) * g.nkb + kb) * kSliceChunk + (tj & 1) * (kSliceChunk / 2);
#pragma unroll
        for (int dgt = 0; dgt < kDigits; ++dgt) {
          bulk_g2s(st + dgt * kSliceChunk, g.planes + dgt * g.plane_bytes + ca, kSliceChunk, &sh->full[s]);
          bulk_g2s(st + kDigits * kSliceChunk + dgt * (kSliceChunk / 2), g.planes + dgt * g.plane_bytes + cb,
                   kSliceChunk / 2, &sh->full[s]);
        }
      }
      if (++s == kSyrkStages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc_i8(kFB, kFN);
    const uint64_t desc0 = make_kmajor_desc(smem_u32(stage_base), 128, 512);
    int s = 0;
    uint32_t ph = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&sh->full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t st = desc0 + static_cast<uint64_t>(s * (kSyrkStageBytes >> 4));
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {          // two 32-sample MMA steps per chunk: +256 B
          const uint32_t first = (kb != kb0 || ks != 0) ? 1u : 0u;
#pragma unroll
          for (int w = 0; w < kClasses; ++w) {    // accumulator w collects the digit pairs with sa + sb == w
            bool fresh = true;
#pragma unroll
            for (int sa = 0; sa < kDigits; ++sa) {
              const int sb = w - sa;
              if (sb < 0 || sb >= kDigits) continue;
              const uint64_t a = st + static_cast<uint64_t>(sa * (kSliceChunk >> 4) + ks * 16);
              const uint64_t b = st + static_cast<uint64_t>(kDigits * (kSliceChunk >> 4) + sb * (kSliceChunk >> 5) + ks * 16);
              mma_i8_ss(tmem_base + w * kFN, a, b, idesc, fresh ? first : 1u);
              fresh = false;
            }
          }
        }
        tc_commit(&sh->empty[s]);
      }
      if (++s == kSyrkStages) { s = 0; ph ^= 1; }
    }
    if (elect_one()) tc_commit(&sh->acc_full);
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;            // feature within block ti
    mbar_wait(&sh->acc_full, 0);
    tc_fence_after();
    double* out = g.partial + (static_cast<long long>(split) * g.d_pad + (ti * kFB + row)) * g.d_pad + tj * kFN;
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < kFN; c0 += 16) {
      double v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.0;
      if (kb1 > kb0) {
        // low classes first: the running sum grows towards the leading term
#pragma unroll
        for (int w = 0; w < kClasses; ++w) {
          uint32_t dw[16];
          tmem_ld16(t_addr + w * kFN + c0, dw);
          tmem_wait_ld();
          const double wt = static_cast<double>(1ull << (8 * w));
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fma(static_cast<double>(static_cast<int>(dw[j])), wt, v[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) out[c0 + j] = v[j];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// gram[k][l] += 2^-(e_k+e_l) * sum over K splits of the (upper-tile) partials;
// sum[k] += 2^-e_k * sum of the per-block integer column sums.
__global__ void syrk_reduce_kernel(const double* __restrict__ partial, int n_ksplit, int d_pad, int d,
                                   const int* __restrict__ exps, const double* __restrict__ colsum_part,
                                   int n_colsum_parts, double* __restrict__ gram, double* __restrict__ sum) {
  const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (e < d) {
    double s = 0.0;
    for (int p = 0; p < n_colsum_parts; ++p) s += colsum_part[static_cast<long long>(p) * d_pad + e];
    sum[e] += ldexp(s, -exps[e]);
  }
  if (e >= static_cast<long long>(d) * d) return;
  const int k = static_cast<int>(e / d), l = static_cast<int>(e % d);
  const bool upper = (l / kFN) >= 2 * (k / kFB);   // tile (k/128, l/64) was computed
  const long long src = upper ? static_cast<long long>(k) * d_pad + l : static_cast<long long>(l) * d_pad + k;
  double s = 0.0;
  for (int sp = 0; sp < n_ksplit; ++sp) s += partial[static_cast<long long>(sp) * d_pad * d_pad + src];
  gram[e] += ldexp(s, -(exps[k] + exps[l]));
}

// ------------------------------------------------------------------ host side
struct CovTcPlan {
  int d_pad, nfb, n_tiles;
  long long chunk_rows;   // samples per pass
  int nkb;                // K blocks of a full chunk
  int n_ksplit;
  int slice_blocks;       // gridDim.x of slice_pack_kernel
  size_t off_planes, off_partial, off_colsum, off_maxabs, off_exps, bytes;
};

static CovTcPlan cov_tc_plan(long long n, int d) {
  CovTcPlan p;
  p.d_pad = static_cast<int>(round_up_ll(d, kFB));
  p.nfb = p.d_pad / kFB;
  p.n_tiles = p.nfb * (p.nfb + 1);
  p.chunk_rows = n < kChunkRows ? round_up_ll(n, kSK) : kChunkRows;
  p.nkb = static_cast<int>(p.chunk_rows / kSK);
  int ks = 148 / p.n_tiles;                                         // one wave: at most one item per SM
  const int ks_min = (p.nkb + kMaxKbPerItem - 1) / kMaxKbPerItem;   // int32 accumulation bound
  if (ks < ks_min) ks = ks_min;
  if (ks > p.nkb) ks = p.nkb;
  p.n_ksplit = ks < 1 ? 1 : ks;
  p.slice_blocks = p.nkb < 296 ? p.nkb : 296;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += static_cast<size_t>(round_up_ll(static_cast<long long>(b), 256)); return o; };
  p.off_planes = take(static_cast<size_t>(kDigits) * p.nfb * p.nkb * kSliceChunk);
  p.off_partial = take(static_cast<size_t>(p.n_ksplit) * p.d_pad * p.d_pad * 8);
  p.off_colsum = take(static_cast<size_t>(p.slice_blocks) * 4 * p.d_pad * 8);
  p.off_maxabs = take(static_cast<size_t>(p.d_pad) * 4);
  p.off_exps = take(static_cast<size_t>(p.d_pad) * 4);
  p.bytes = off;
  return p;
}

size_t cov_tc_ws_bytes(long long n, int d) { return cov_tc_plan(n, d).bytes; }

// sum[d] += column sums, gram[d*d] += X^T X for fp32 X, through the integer tensor-core path.
int cov_tc_accumulate(cudaStream_t st, int dev, const float* X, long long n, int d, long long ld, double* sum,
                      double* gram, void* ws) {
  const CovTcPlan p = cov_tc_plan(n, d);
  uint8_t* b = static_cast<uint8_t*>(ws);
  int8_t* planes = reinterpret_cast<int8_t*>(b + p.off_planes);
  double* partial = reinterpret_cast<double*>(b + p.off_partial);
  double* colsum = reinterpret_cast<double*>(b + p.off_colsum);
  float* maxabs = reinterpret_cast<float*>(b + p.off_maxabs);
  int* exps = reinterpret_cast<int*>(b + p.off_exps);
  const long long plane_bytes = static_cast<long long>(p.nfb) * p.nkb * kSliceChunk;
  int rc;
  if ((rc = check_cuda(cudaFuncSetAttribute(syrk_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(kSyrkSmemBytes)), "cudaFuncSetAttribute(syrk_i8)"))) return rc;
  // one scale per column for the whole call, so that every chunk sums on the same grid
  if ((rc = check_cuda(cudaMemsetAsync(maxabs, 0, static_cast<size_t>(p.d_pad) * 4, st), "memset"))) return rc;
  {
    const long long slabs = (n + 1023) / 1024;
    dim3 grid((d + 255) / 256, static_cast<unsigned>(slabs < 592 ? slabs : 592));
    col_absmax_kernel<<<grid, 256, 0, st>>>(X, 0, n, d, ld, maxabs);
    if ((rc = check_launch("col_absmax_kernel"))) return rc;
    col_exp_kernel<<<(p.d_pad + 255) / 256, 256, 0, st>>>(maxabs, p.d_pad, exps);
    if ((rc = check_launch("col_exp_kernel"))) return rc;
  }
  (void)dev;
  for (long long r0 = 0; r0 < n; r0 += p.chunk_rows) {
    const long long r1 = r0 + p.chunk_rows < n ? r0 + p.chunk_rows : n;
    const int nkb = static_cast<int>((r1 - r0 + kSK - 1) / kSK);
    int ks = p.n_ksplit < nkb ? p.n_ksplit : nkb;
    const int sblocks = p.slice_blocks < nkb ? p.slice_blocks : nkb;
    // digit planes of this chunk are indexed with the chunk's own nkb
    const long long pb = static_cast<long long>(p.nfb) * nkb * kSliceChunk;
    slice_pack_kernel<<<dim3(sblocks, p.nfb), 512, 0, st>>>(X, r0, r1, d, ld, exps, planes, pb, nkb, colsum, p.d_pad);
    if ((rc = check_launch("slice_pack_kernel"))) return rc;
    SyrkGeom g{planes, pb, p.nfb, nkb, p.n_tiles, ks, partial, p.d_pad};
    syrk_i8_kernel<<<p.n_tiles * ks, kSyrkThreads, kSyrkSmemBytes, st>>>(g);
    if ((rc = check_launch("syrk_i8_kernel"))) return rc;
    const long long dd = static_cast<long long>(d) * d;
    syrk_reduce_kernel<<<static_cast<unsigned>((dd + 255) / 256), 256, 0, st>>>(partial, ks, p.d_pad, d, exps, colsum,
                                                                               sblocks * 4, gram, sum);
    if ((rc = check_launch("syrk_reduce_kernel"))) return rc;
  }
  (void)plane_bytes;
  return AMB_OK;
}

}  // namespace amb

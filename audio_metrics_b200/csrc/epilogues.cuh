// Epilogue functors for the pair engine (see the concept in pair_engine.cuh).
// Thread = one row of the 128 x 256 tile; `acc` holds 32 consecutive columns of
// that row as raw fp32 bits.  acc * inv_scale_a[row] * inv_scale_b[col] is the
// dot product <x_row, y_col>; inv_scale_b is the same for all 256 columns of a tile
// (pack.cu scales per tile), so column vector 0 is read once per chunk.
#pragma once
#include "internal.cuh"
#include "pair_engine.cuh"

namespace amb {

__device__ __forceinline__ float f32(uint32_t bits) { return __uint_as_float(bits); }
constexpr float kInf = __builtin_huge_valf();

// ----------------------------------------------------------------- debug dump
struct DumpEpi {
  static constexpr int kColVecs = 1;
  static constexpr bool kScratch = false;
  static constexpr bool kChunkMin = false;
  const float* inv_a;
  const float* inv_b;
  float* C;
  long long ldc;
  long long na, nb;
  int checksum_only;       // 1: C[a_row] = sum_j dot (timing runs; no N x M write)
  struct Row { float isr; long long a_row; float sum; };
  __device__ const float* colvec_ptr(int) const { return inv_b; }
  __device__ const float* cmin_ptr() const { return nullptr; }
  __device__ const float* cmax_ptr() const { return nullptr; }
  __device__ void row_begin(Row& r, const ItemCoord&, long long a_row, int, float*) const {
    r.isr = inv_a[a_row];
    r.a_row = a_row;
    r.sum = 0.f;
  }
  __device__ void tile_begin(Row&, const float (*)[kTileN]) const {}
  __device__ void tile_end(Row&, float*) const {}
  template <bool>
  __device__ void chunk(Row& r, const uint32_t (&acc)[32], const float (*cv)[kTileN], int c0, int,
                        long long b_row0, float*, float, float) const {
    if (r.a_row >= na) return;
    const float sc = cv[0][c0] * r.isr;
    if (checksum_only) {
#pragma unroll
      for (int j = 0; j < 32; ++j) r.sum += f32(acc[j]) * sc;
      return;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (b_row0 + j < nb) C[r.a_row * ldc + b_row0 + j] = f32(acc[j]) * sc;
    }
  }
  __device__ void row_end(Row& r, const ItemCoord&, int, long long, int, int, int) const {
    if (checksum_only && r.a_row < na) atomicAdd(&C[r.a_row], r.sum);   // two column halves per row
  }
};

// --------------------------------------------------------- kernel distance sums
// Problems come in triples per subset: 3s+0 = K(f1,f1), 3s+1 = K(f2,f2),
// 3s+2 = K(f1,f2)  (kd.py:119-122).  Diagonal and off-diagonal entries are summed
// separately (kd.py:53-68 needs K.sum(axis=1) - diag, the diagonal sums and, for the
// u-statistic, trace(K_XY)).  Each epilogue warp writes two fp64 partials per work item:
// partial[(item*8 + half*4 + quarter)*2 + {0 off-diagonal, 1 diagonal}].
struct KdEpi {
  static constexpr int kColVecs = 1;
  static constexpr bool kScratch = false;
  static constexpr bool kChunkMin = false;
  const float* inv_a;
  const float* inv_b;
  int kernel_type;
  double gamma, coef0;
  int degree;
  double rbf_scale;          // -1 / (2 sigma^2)
  const float* norm_a;       // for rbf
  const float* norm_b;
  int m_valid;               // rows/cols per problem that are real samples
  double* partial;
  struct Row { double sum; double dsum; double gr; float na; int row_in_problem; bool valid; };
  __device__ const float* colvec_ptr(int) const { return inv_b; }
  __device__ const float* cmin_ptr() const { return nullptr; }
  __device__ const float* cmax_ptr() const { return nullptr; }
  __device__ void row_begin(Row& r, const ItemCoord& c, long long a_row, int, float*) const {
    r.row_in_problem = c.rt * kTileM + static_cast<int>(a_row % kTileM);
    r.valid = r.row_in_problem < m_valid;
    r.sum = 0.0;
    r.dsum = 0.0;
    const double isr = static_cast<double>(inv_a[a_row]);
    r.gr = (kernel_type == AMB_KERNEL_POLY ? gamma : 1.0) * isr;
    r.na = norm_a ? norm_a[a_row] : 0.f;
  }
  __device__ __forceinline__ double kval(const Row& r, float acc, float isc, long long b_row) const {
    const double q = static_cast<double>(acc * isc);   // power-of-two scaling: exact
    if (kernel_type == AMB_KERNEL_POLY) {
      const double u = fma(q, r.gr, coef0);
      double p = u;
      for (int e = 1; e < degree; ++e) p *= u;
      return degree == 0 ? 1.0 : p;
    } else {
      // rbf_kernel (kd.py:86-109): exp(-|x-y|^2 / (2 sigma^2))
      double d2 = static_cast<double>(r.na) + static_cast<double>(norm_b[b_row]) - 2.0 * q * r.gr;
      d2 = d2 < 0 ? 0 : d2;
      return exp(d2 * rbf_scale);
    }
  }
  __device__ void tile_begin(Row&, const float (*)[kTileN]) const {}
  __device__ void tile_end(Row&, float*) const {}
  template <bool>
  __device__ void chunk(Row& r, const uint32_t (&acc)[32], const float (*cv)[kTileN], int c0,
                        int col0, long long b_row0, float*, float, float) const {
    if (!r.valid) return;
    const bool edge = (col0 + 32 > m_valid) || (col0 <= r.row_in_problem && r.row_in_problem < col0 + 32);
    if (!edge && kernel_type == AMB_KERNEL_POLY && degree == 3) {
      double s0 = 0, s1 = 0;
      const double g2 = r.gr * static_cast<double>(cv[0][c0]);   // powers of two: exact
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const double u0 = fma(static_cast<double>(f32(acc[j])), g2, coef0);
        const double u1 = fma(static_cast<double>(f32(acc[j + 1])), g2, coef0);
        s0 = fma(u0 * u0, u0, s0);
        s1 = fma(u1 * u1, u1, s1);
      }
      r.sum += s0 + s1;
    } else {
      double s = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = col0 + j;
        if (col >= m_valid) continue;
        const double kv = kval(r, f32(acc[j]), cv[0][c0 + j], b_row0 + j);
        if (col == r.row_in_problem) r.dsum += kv;
        else s += kv;
      }
      r.sum += s;
    }
  }
  __device__ void row_end(Row& r, const ItemCoord&, int item, long long, int quarter, int lane, int half) const {
    double s = r.valid ? r.sum : 0.0;
    double ds = r.valid ? r.dsum : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ds += __shfl_xor_sync(0xffffffffu, ds, o);
    }
    if (lane == 0) {
      const long long o = (static_cast<long long>(item) * kEpiWarps + half * 4 + quarter) * 2;
      partial[o] = s;
      partial[o + 1] = ds;
    }
  }
};

// ------------------------------------------------------------ error band
// The tensor-core dot product q~ differs from the exact <x, y> by at most
//   kBandDot * |x| |y|
// (operand split 2^-22 twice, dropped lo*lo 2^-22, and 96 accumulate steps that
// each truncate to fp32: measured bias -2.7e-6 on all-positive data, bound
// 96 * 2^-23 = 1.1e-5, doubled for the MMA-internal summation), and the fp32
// epilogue arithmetic adds at most kBandAbs * (|x|^2 + |y|^2).  Every comparison
// the tensor-core pass makes is therefore only trusted outside this band;
// pairs inside it are re-evaluated exactly (fp64) by the refine kernels.
constexpr float kBandDot = 2.4e-5f;
constexpr float kBandAbs = 1.0e-6f;
__host__ __device__ inline float band_key(float nrm_x, float nrm_y_max) {
  // bound on |t~ - t| for t = |y|^2 - 2<x,y>
  return 2.0f * kBandDot * sqrtf(nrm_x) * sqrtf(nrm_y_max) + kBandAbs * (nrm_x + nrm_y_max);
}
// Single-pass engine (hi planes only, pair_engine1_kernel): with x = hx + ex, y = hy + ey,
//   <x,y> - <hx,hy> = <ex,hy> + <hx,ey> + <ex,ey>,   |.| <= (rho_x |y| + |x| rho_y)(1 + 2^-10)
// where rho = |e| is the MEASURED residual norm of the row (pack.cu; <= 2^-11 |x|, typically
// 0.4 of that), by Cauchy-Schwarz; the 32-step truncating accumulation adds at most
// 2 * 32 * 2^-23 |x||y| (same model as above).  rho_y and |y| enter as set-wide maxima.
constexpr float kBandAcc1 = 8.0e-6f;
__host__ __device__ inline float band_key1(float nrm_x, float rho_x, float nrm_y_max, float rho_y_max) {
  const float ax = sqrtf(nrm_x), ay = sqrtf(nrm_y_max);
  return 2.0f * ((rho_x * ay + ax * rho_y_max) * 1.002f + kBandAcc1 * ax * ay) + kBandAbs * (nrm_x + nrm_y_max);
}

// -------------------------------------------------- per-row (k+1)-smallest lists
// Ranking key for row i over columns j:  t_ij = |y_j|^2 - 2 <x_i, y_j>
// (d_ij^2 = |x_i|^2 + t_ij).  Each thread keeps its row's K smallest approximate
// keys, with their column indices, sorted in registers across the whole column
// sweep; one list per (split, column half, row) leaves the kernel.  K exceeds k+1 by a margin
// so that the refine kernel can certify that the exact (k+1)-th neighbour is
// among the kept candidates.
template <int K>
struct TopkEpi {
  static constexpr int kColVecs = 2;   // 0: inv_scale_b, 1: norm_b (+inf on padding)
  static constexpr bool kScratch = true;
  static constexpr bool kChunkMin = true;   // the CTA-pair engine stages cmin_b next to the column vectors
  const float* inv_a;
  const float* inv_b;
  const float* norm_b;
  const float* cmin_b;     // per-chunk minima of norm_b (nullptr: no prefilter)
  float* keys;             // [2 * n_split][list_rows][K]   (list = 2 * split + half)
  int* cols;               // [2 * n_split][list_rows][K]   (-1 = empty)
  long long list_rows;     // rows covered by this launch (multiple of 128)
  long long a_row_base;    // packed row of list row 0
  // mine / peer: where the two threads of a row (one per column half) publish their current
  // K-th smallest key.  A key that is not below the OTHER half's K-th smallest cannot be among
  // the row's K smallest overall, so each half filters with the minimum of the two — the two
  // lists together then take about as many insertions as one list over all columns would.
  // The refine kernel's certificate is unaffected: whatever a list rejected was >= some list's
  // K-th key at that time >= that list's final K-th key >= the K-th smallest of the union.
  struct Row { float m2isr, sc; float v[K]; int c[K]; float* mine; const volatile float* peer; int qn; };
  static constexpr int kQueue = 4;   // (key, column) pairs a thread can park: 8 of its kScratchFloats
  __device__ const float* colvec_ptr(int v) const { return v == 0 ? inv_b : norm_b; }
  __device__ const float* cmin_ptr() const { return cmin_b; }
  __device__ const float* cmax_ptr() const { return nullptr; }
  __device__ void row_begin(Row& r, const ItemCoord&, long long a_row, int half, float* xchg) const {
    r.m2isr = -2.0f * inv_a[a_row];
#pragma unroll
    for (int i = 0; i < K; ++i) { r.v[i] = kInf; r.c[i] = -1; }
    r.mine = xchg + half;
    r.peer = xchg + (half ^ 1);
    *r.mine = kInf;
    r.qn = 0;
  }
  // Sorted insert with no dependent chain: slot i of the new list is old slot i-1 if the key goes
  // in front of it, the key itself if it lands here, old slot i otherwise — every slot from the OLD
  // list and the key alone (ties: behind equal keys already present).  A key that is not below
  // the last slot (or +inf from an idle lane) changes nothing.
  __device__ __forceinline__ void insert(Row& r, float key, int col) const {
    bool p[K];
#pragma unroll
    for (int i = 0; i < K; ++i) p[i] = key < r.v[i];
#pragma unroll
    for (int i = K - 1; i > 0; --i) {
      r.v[i] = p[i - 1] ? r.v[i - 1] : (p[i] ? key : r.v[i]);
      r.c[i] = p[i - 1] ? r.c[i - 1] : (p[i] ? col : r.c[i]);
    }
    r.v[0] = p[0] ? key : r.v[0];
    r.c[0] = p[0] ? col : r.c[0];
  }
  // -2 / (scale_a scale_b) < 0: one value per column tile (a packed 256-row tile has one scale;
  // powers of two, so the product is exact)
  __device__ void tile_begin(Row& r, const float (*cv)[kTileN]) const { r.sc = cv[0][0] * r.m2isr; }
  template <bool kStaged>
  __device__ void chunk(Row& r, const uint32_t (&acc)[32], const float (*cv)[kTileN], int c0, int,
                        long long b_row0, float* scratch, float cmin, float) const {
    const float sc = r.sc;
    const float thr = fminf(r.v[K - 1], *r.peer);
    // Cheapest test first, on the raw accumulators, per 8-column group (kernels that stage the chunk
    // minima): a candidate needs key_j = fma(acc_j, sc, |y_j|^2) < thr.  With m = max acc_j over the
    // group and |y_j|^2 >= cmin over the chunk, acc_j sc + |y_j|^2 >= m sc + cmin as real numbers
    // (sc < 0), and rounding is monotone, so key_j >= fma(m, sc, cmin): a group whose bound is not
    // below thr holds nothing — an exact test, no slack.  Padding columns (|y|^2 = +inf) and an empty
    // list (thr = +inf) come out right by themselves.
    unsigned gm = 0xfu;                     // bit h: group h may hold a candidate of this row
    if (kStaged) {
      gm = 0;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float ma = fmaxf(fmaxf(f32(acc[8 * h + 0]), f32(acc[8 * h + 1])), fmaxf(f32(acc[8 * h + 2]), f32(acc[8 * h + 3])));
        const float mb = fmaxf(fmaxf(f32(acc[8 * h + 4]), f32(acc[8 * h + 5])), fmaxf(f32(acc[8 * h + 6]), f32(acc[8 * h + 7])));
        gm |= (fmaf(fmaxf(ma, mb), sc, cmin) < thr) ? (1u << h) : 0u;
      }
      if (__builtin_expect(!__any_sync(0xffffffffu, gm != 0), 1)) return;
    }
    // Some row of the warp may take new candidates: per group that can hold one, each thread evaluates
    // its 8 keys and the bit mask of those below its threshold.  Candidates are only PARKED here — (key, column) pairs in the thread's shared-memory slots — and
    // inserted by drain() once the engine has handed the accumulator back to the MMA warp
    // (tile_end).  Measured: the scan alone never delays the next tile, the inserts did — a warp
    // that met candidates in several chunks of one tile held the accumulator (and through the pair
    // barrier both CTAs' tensor pipes) beyond the tile period, while on average the epilogue warps
    // idle 40 % of the time.  The groups are walked by a run-time loop (the warp-uniform switch
    // moves the group's eight accumulators into fixed registers) so that this code exists once,
    // not four times: the whole chunk loop stays inside the scheduler's instruction cache.
    unsigned todo = __reduce_or_sync(0xffffffffu, gm);
#pragma unroll 1
    while (todo) {
      const int h = __ffs(todo) - 1;
      todo &= todo - 1;
      float a[8];
      switch (h) {
        case 0:
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = f32(acc[j]);
          break;
        case 1:
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = f32(acc[8 + j]);
          break;
        case 2:
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = f32(acc[16 + j]);
          break;
        default:
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = f32(acc[24 + j]);
          break;
      }
      const float4 na = *reinterpret_cast<const float4*>(&cv[1][c0 + 8 * h]);
      const float4 nb = *reinterpret_cast<const float4*>(&cv[1][c0 + 8 * h + 4]);
      float t[8];
      t[0] = fmaf(a[0], sc, na.x);
      t[1] = fmaf(a[1], sc, na.y);
      t[2] = fmaf(a[2], sc, na.z);
      t[3] = fmaf(a[3], sc, na.w);
      t[4] = fmaf(a[4], sc, nb.x);
      t[5] = fmaf(a[5], sc, nb.y);
      t[6] = fmaf(a[6], sc, nb.z);
      t[7] = fmaf(a[7], sc, nb.w);
      const int col0 = static_cast<int>(b_row0) + 8 * h;
      unsigned pend = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) pend |= (t[j] < thr) ? (1u << j) : 0u;
      // Park them.  If some thread has more than fit (the first tiles of an item, when the list is
      // still filling), everybody inserts what is parked first and the loop goes round again.
      bool again;
#pragma unroll 1
      do {
        again = __any_sync(0xffffffffu, r.qn + __popc(pend) > kQueue);
        if (again) drain(r, scratch);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (((pend >> j) & 1u) && r.qn < kQueue) {
            scratch[2 * r.qn] = t[j];
            scratch[2 * r.qn + 1] = __int_as_float(col0 + j);
            ++r.qn;
            pend &= ~(1u << j);
          }
        }
      } while (again && __any_sync(0xffffffffu, pend != 0));
    }
  }
  // Insert everything parked by chunk(); the warp loops while any lane has an entry left.
  __device__ __forceinline__ void drain(Row& r, const float* scratch) const {
    if (!__any_sync(0xffffffffu, r.qn != 0)) return;
#pragma unroll 1
    do {
      float key = kInf;
      int col = -1;
      if (r.qn) {
        --r.qn;
        key = scratch[2 * r.qn];
        col = __float_as_int(scratch[2 * r.qn + 1]);
      }
      insert(r, key, col);
    } while (__any_sync(0xffffffffu, r.qn != 0));
    *r.mine = r.v[K - 1];
  }
  __device__ void tile_end(Row& r, float* scratch) const { drain(r, scratch); }
  __device__ void row_end(Row& r, const ItemCoord& c, int, long long a_row, int, int, int half) const {
    const long long o = (static_cast<long long>(2 * c.split + half) * list_rows + (a_row - a_row_base)) * K;
#pragma unroll
    for (int i = 0; i < K; ++i) { keys[o + i] = r.v[i]; cols[o + i] = r.c[i]; }
  }
};

// ------------------------------------------------------- neighbourhood counts
// Row i = reference sample, column j = candidate (prdc.py:34-48).
//   in_ref (i,j):  d_ij^2 < r_ref_i^2   <=>  t_ij = |y_j|^2 - 2<x,y>  <  r_ref_i^2 - |x_i|^2  =: A_i
//   in_cand(i,j):  d_ij^2 < r_cand_j^2  <=>  u_ij = |x_i|^2 - 2<x,y>  <  r_cand_j^2 - |y_j|^2 =: B_j
// Thresholds come in pairs (lo, hi) = (A - band, A + band): below lo the pair is a
// certain hit, in [lo, hi) it is appended to the uncertain list and decided later
// in fp64.  lo = hi = -inf on padding rows / columns, |y_j|^2 = +inf on padding.
struct PairEntry { uint32_t i; uint32_t j_kind; };   // j_kind: bit 31 = 1 -> in_cand test

struct CountEpi {
  static constexpr int kColVecs = 4;   // 0: inv_scale_b, 1: norm_b, 2: B_hi, 3: B_lo
  static constexpr bool kScratch = false;
  static constexpr bool kChunkMin = true;   // cmin_b and cmax_bhi are staged by the CTA-pair engine
  const float* inv_a;
  const float* norm_a;
  const float* a_lo;       // indexed by packed A row
  const float* a_hi;
  const float* inv_b;
  const float* norm_b;
  const float* b_lo;       // indexed by packed B row
  const float* b_hi;
  const float* cmin_b;     // per-chunk minima of norm_b (nullptr: no prefilter)
  const float* cmax_bhi;   // per-chunk maxima of b_hi (nullptr: no prefilter)
  int32_t* col_count;      // [m], atomically incremented
  uint8_t* row_recall;     // [rows of this launch], relative to a_row_base
  uint8_t* row_cover;
  long long a_row_base;
  long long a_row_end;     // packed A rows >= this are padding
  PairEntry* list;         // uncertain pairs
  unsigned long long* list_count;
  unsigned long long list_cap;
  struct Row { float m2isr, sc, nx, Alo, Ahi; bool rec, cov; uint32_t i; };
  __device__ const float* colvec_ptr(int v) const {
    return v == 0 ? inv_b : (v == 1 ? norm_b : (v == 2 ? b_hi : b_lo));
  }
  __device__ const float* cmin_ptr() const { return cmin_b; }
  __device__ const float* cmax_ptr() const { return cmax_bhi; }
  __device__ void row_begin(Row& r, const ItemCoord&, long long a_row, int, float*) const {
    r.m2isr = -2.0f * inv_a[a_row];
    r.nx = norm_a[a_row];
    r.Alo = a_lo[a_row];
    r.Ahi = a_hi[a_row];
    r.rec = false;
    r.cov = false;
    r.i = static_cast<uint32_t>(a_row);
  }
  __device__ __forceinline__ void push(uint32_t i, uint32_t j_kind) const {
    const unsigned long long pos = atomicAdd(list_count, 1ull);
    if (pos < list_cap) list[pos] = PairEntry{i, j_kind};
  }
  __device__ void tile_begin(Row& r, const float (*cv)[kTileN]) const { r.sc = cv[0][0] * r.m2isr; }
  __device__ void tile_end(Row&, float*) const {}
  template <bool kStaged>
  __device__ void chunk(Row& r, const uint32_t (&acc)[32], const float (*cv)[kTileN], int c0, int,
                        long long b_row0, float*, float cmin, float cmax) const {
    bool any_ref = false, any_cand = false;
    const float sc = r.sc;                  // -2 / (scale_a scale_b): one per tile (powers of two, exact)
    // Both tests first on the chunk as a whole, from the largest raw accumulator (exact bounds):
    //   in_ref  needs |y_j|^2 + sc acc_j < A_hi,   and |y_j|^2 >= cmin over the chunk;
    //   in_cand needs |x_i|^2 + sc acc_j < B_hi_j, and B_hi_j <= cmax over the chunk.
    // Kernels that do not stage the chunk arrays (cmin = -inf) test every column as before.
    if (kStaged) {
      float m4[4];                           // four independent chains, then a tree
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        m4[h] = f32(acc[8 * h]);
#pragma unroll
        for (int j = 1; j < 8; ++j) m4[h] = fmaxf(m4[h], f32(acc[8 * h + j]));
      }
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      any_ref = fmaf(mx, sc, cmin) < r.Ahi;    // every t_j >= fma(mx, sc, cmin): see TopkEpi::chunk
      any_cand = fmaf(mx, sc, r.nx) < cmax;    // every u_j >= fma(mx, sc, |x|^2), every B_hi_j <= cmax
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        any_ref |= (fmaf(f32(acc[j]), sc, cv[1][c0 + j]) < r.Ahi);
        any_cand |= (fmaf(f32(acc[j]), sc, r.nx) < cv[2][c0 + j]);
      }
    }
    // Hits are rare (about k per row over the whole sweep).  A thread that has one
    // rebuilds its comparisons as bit masks and walks the set bits on its own:
    // certain hits are counted, hits inside the band go to the refine list.
    if (any_ref) {
      unsigned m_hi = 0, m_lo = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float t = fmaf(f32(acc[j]), sc, cv[1][c0 + j]);
        m_hi |= (t < r.Ahi) ? (1u << j) : 0u;
        m_lo |= (t < r.Alo) ? (1u << j) : 0u;
      }
      r.cov |= (m_lo != 0);
      unsigned unc = m_hi & ~m_lo;
      while (m_lo) {
        const int j = __ffs(m_lo) - 1;
        m_lo &= m_lo - 1;
        atomicAdd(col_count + b_row0 + j, 1);
      }
      while (unc) {
        const int j = __ffs(unc) - 1;
        unc &= unc - 1;
        push(r.i, static_cast<uint32_t>(b_row0 + j));
      }
    }
    if (any_cand) {
      unsigned m_hi = 0, m_lo = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float u = fmaf(f32(acc[j]), sc, r.nx);
        m_hi |= (u < cv[2][c0 + j]) ? (1u << j) : 0u;
        m_lo |= (u < cv[3][c0 + j]) ? (1u << j) : 0u;
      }
      r.rec |= (m_lo != 0);
      unsigned unc = m_hi & ~m_lo;
      while (unc) {
        const int j = __ffs(unc) - 1;
        unc &= unc - 1;
        push(r.i, static_cast<uint32_t>(b_row0 + j) | 0x80000000u);
      }
    }
  }
  __device__ void row_end(Row& r, const ItemCoord&, int, long long a_row, int, int, int) const {
    if (a_row < a_row_end) {
      if (r.rec) row_recall[a_row - a_row_base] = 1;
      if (r.cov) row_cover[a_row - a_row_base] = 1;
    }
  }
};

}  // namespace amb

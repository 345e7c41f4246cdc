// PRDC on the pair engine: k-NN radii (prdc.py:4-14) and neighbourhood counts
// (prdc.py:18-50).  The tensor-core sweep is a filter with a proven error band
// (epilogues.cuh); every decision that falls inside the band is re-made from the
// original rows in fp64 by the refine kernels below, so radii are the correctly
// rounded exact distances and counts are the exact-arithmetic counts.
#include <cstdlib>
#include <vector>

#include "engine_launch.cuh"

namespace amb {

// ------------------------------------------------------------------ helpers
// out_max[0] = max |x|^2, out_max[1] = max rho over the real rows of a packed set.
__global__ void max_norm_kernel(const float* __restrict__ norm, const float* __restrict__ rho, long long n,
                                float* out_max) {
  float m = 0.f, mr = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = norm[i];
    if (v < kInf) { m = fmaxf(m, v); mr = fmaxf(mr, rho[i]); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o));
  }
  if ((threadIdx.x & 31) == 0) {   // values are >= 0: integer order == float order
    atomicMax(reinterpret_cast<int*>(out_max), __float_as_int(m));
    atomicMax(reinterpret_cast<int*>(out_max) + 1, __float_as_int(mr));
  }
}

// Error band of the filter pass for row x against a set with maxima mx = {max |y|^2, max rho_y}.
__device__ __forceinline__ float filter_band(bool single_pass, float nx, float rho_x, const float* mx) {
  return single_pass ? band_key1(nx, rho_x, mx[0], mx[1]) : band_key(nx, mx[0]);
}

template <typename T>
__device__ __forceinline__ double exact_sqdist_warp(const T* __restrict__ x, const T* __restrict__ y, int d,
                                                    int lane) {
  double s = 0.0;
  for (int k = lane; k < d; k += 32) {
    const double df = static_cast<double>(x[k]) - static_cast<double>(y[k]);
    s = fma(df, df, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// Distance exactly as a correctly rounded fp32 evaluation would give it; ties in
// exact arithmetic stay ties, which keeps the strict '<' of prdc.py:37-48.
__device__ __forceinline__ float dist_f32(double d2) { return static_cast<float>(sqrt(d2)); }

// ------------------------------------------------------------- radii: refine
// One warp per row: merge the per-split candidate lists, keep the Kt best
// approximate candidates, evaluate their exact squared distances, take the
// (k+1)-th smallest, and certify it against the band.  Uncertified rows go to
// the brute-force list.
template <typename T>
__global__ void __launch_bounds__(256)
knn_refine_kernel(const T* __restrict__ X, long long ld, int d, long long n, long long row0,
                  long long nrows, int k, int Kt, int n_split, long long list_rows,
                  const float* __restrict__ keys, const int* __restrict__ cols,
                  const float* __restrict__ norm, const float* __restrict__ rho, bool single_pass,
                  const float* __restrict__ max_norm, float* __restrict__ radii, int* __restrict__ unresolved,
                  int* __restrict__ n_unresolved) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  if (row >= nrows) return;
  const long long i = row0 + row;
  const T* xi = X + i * ld;
  // --- select the Kt smallest approximate candidates over all splits (lane c keeps the c-th)
  const int n_cand = n_split * Kt;
  float sel_key = kInf;
  int sel_col = -1;
  float last_key = -kInf;
  int last_col = -1;
  for (int r = 0; r < Kt; ++r) {
    float best = kInf;
    int best_col = 0x7fffffff;
    for (int c = lane; c < n_cand; c += 32) {
      const int sp = c / Kt, e = c - sp * Kt;
      const long long o = (static_cast<long long>(sp) * list_rows + row) * Kt + e;
      const float kv = keys[o];
      const int cc = cols[o];
      if (cc < 0) continue;
      const bool after = (kv > last_key) || (kv == last_key && cc > last_col);
      if (after && (kv < best || (kv == best && cc < best_col))) { best = kv; best_col = cc; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oc = __shfl_xor_sync(0xffffffffu, best_col, o);
      if (ob < best || (ob == best && oc < best_col)) { best = ob; best_col = oc; }
    }
    if (best_col == 0x7fffffff) break;   // fewer than Kt candidates exist
    last_key = best;
    last_col = best_col;
    if (lane == r) { sel_key = best; sel_col = best_col; }
  }
  // --- exact squared distances of the selected candidates (lane c <- candidate c).  Only
  // candidates that can be among the k+1 nearest are evaluated: with s_k the (k+1)-th smallest
  // approximate key, k+1 candidates have exact key <= s_k + band, so a candidate whose
  // approximate key exceeds s_k + 2 band (exact key > s_k + band) cannot be one of them.
  const float nx_f = norm[i];
  const float band_f = filter_band(single_pass, nx_f, rho[i], max_norm);
  const float s_k = __shfl_sync(0xffffffffu, sel_key, k);      // +inf when fewer than k+1 candidates exist
  const float cut = s_k + 2.0f * band_f + 1e-6f * (fabsf(s_k) + nx_f);
  // The other side: two candidates whose approximate keys differ by more than 2 band are in that order
  // exactly.  With j0 the highest position <= k that follows such a gap, the j0 candidates in front of
  // it are certainly the j0 nearest — their exact values are not needed, only that they rank first —
  // and the (k+1)-th nearest is the (k+1-j0)-th among the rest.  With a band of 1e-3 of the typical
  // spacing that leaves one or two rows to gather per row instead of k+1.
  const float prev_key = __shfl_up_sync(0xffffffffu, sel_key, 1);
  const bool gap = lane >= 1 && lane <= k &&
                   (sel_key - prev_key) > 2.0f * band_f + 1e-6f * (fabsf(sel_key) + fabsf(prev_key) + nx_f);
  const unsigned gaps = __ballot_sync(0xffffffffu, gap);
  const int j0 = gaps ? 31 - __clz(gaps) : 0;
  double my_d2 = __longlong_as_double(0x7ff0000000000000ll);
  int n_sel = 0;
  for (int c = 0; c < Kt; ++c) {
    const int col = __shfl_sync(0xffffffffu, sel_col, c);
    if (col < 0) break;
    const float key_c = __shfl_sync(0xffffffffu, sel_key, c);
    n_sel = c + 1;
    if (c < j0) {                                               // certainly among the j0 nearest: ranks first
      if (lane == c) my_d2 = -__longlong_as_double(0x7ff0000000000000ll);
      continue;
    }
    if (key_c > cut) continue;                                  // stays +inf: never among the k+1 smallest
    const double d2 = exact_sqdist_warp(xi, X + static_cast<long long>(col) * ld, d, lane);
    if (lane == c) my_d2 = d2;
  }
  // --- (k+1)-th smallest exact value among them: rank by (value, lane)
  int rank = 0;
  for (int c = 0; c < n_sel; ++c) {
    const double o = __shfl_sync(0xffffffffu, my_d2, c);
    rank += (o < my_d2 || (o == my_d2 && c < lane)) ? 1 : 0;
  }
  const unsigned who = __ballot_sync(0xffffffffu, lane < n_sel && rank == k);
  double r2 = 0.0;
  bool ok = who != 0;
  if (ok) r2 = __shfl_sync(0xffffffffu, my_d2, __ffs(who) - 1);
  // --- certificate: every column that was not kept has approximate key >= t_last (the
  // largest kept key), hence exact d^2 >= |x_i|^2 + t_last - band.  Complete lists
  // (fewer than Kt candidates overall) are trivially certified.
  const float t_last = __shfl_sync(0xffffffffu, sel_key, Kt - 1);
  const int last_sel = __shfl_sync(0xffffffffu, sel_col, Kt - 1);
  if (ok && last_sel >= 0) {
    const float nx = nx_f;
    const float band = band_f;
    const double bound = static_cast<double>(nx) + static_cast<double>(t_last) - static_cast<double>(band);
    if (!(r2 < bound)) ok = false;
  }
  if (lane == 0) {
    radii[row] = who != 0 ? dist_f32(r2) : kInf;
    if (!ok) {
      const int p = atomicAdd(n_unresolved, 1);
      unresolved[p] = static_cast<int>(row);   // capacity = nrows
    }
  }
}

// Brute force for rows the certificate rejected (near-ties beyond the kept
// margin, heavy duplication, collinear sets whose distances are tiny differences of
// large norms): exact distances to all n rows, (k+1)-th smallest.  One CTA per
// unresolved row, persistent over the WHOLE list — every row the certificate
// rejected is resolved here, however many there are (a set of identical rows costs
// n^2 d fp64 operations: slow, never wrong).
template <typename T>
__global__ void __launch_bounds__(256)
knn_bruteforce_kernel(const T* __restrict__ X, long long ld, int d, long long n, long long row0, int k,
                      const int* __restrict__ unresolved, const int* __restrict__ n_unresolved,
                      float* __restrict__ radii) {
  __shared__ double s_best[8][32];   // per warp: k+1 smallest (k+1 <= 32)
  __shared__ double s_merge[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double INF = __longlong_as_double(0x7ff0000000000000ll);
  const int count = *n_unresolved;
  for (int u = blockIdx.x; u < count; u += gridDim.x) {
    const long long row = unresolved[u];
    const T* xi = X + (row0 + row) * ld;
    // each warp scans columns warp, warp+8, ...; lane l of the warp holds the l-th smallest so far
    double mine = INF;
    for (long long j = warp; j < n; j += 8) {
      const double d2 = exact_sqdist_warp(xi, X + j * ld, d, lane);
      // insert d2 into the sorted-by-lane list (ascending with lane), keep k+1 entries
      const double kth = __shfl_sync(0xffffffffu, mine, k);
      if (d2 < kth) {
        const unsigned below = __ballot_sync(0xffffffffu, mine <= d2);   // lanes whose value stays
        const int pos = __popc(below);                                   // insertion lane
        const double up = __shfl_up_sync(0xffffffffu, mine, 1);
        if (lane == pos) mine = d2;
        else if (lane > pos) mine = up;
      }
    }
    s_best[warp][lane] = lane <= k ? mine : INF;
    __syncthreads();
    if (warp == 0) {
      // merge 8 sorted lists: lane l picks the overall rank-l element by counting
      double v[8];
#pragma unroll
      for (int w = 0; w < 8; ++w) v[w] = s_best[w][lane];
      // rank of each of my 8 values among all 256 values
      for (int w = 0; w < 8; ++w) {
        int rank = 0;
        for (int w2 = 0; w2 < 8; ++w2)
          for (int l2 = 0; l2 < 32; ++l2) {
            const double o = s_best[w2][l2];
            rank += (o < v[w] || (o == v[w] && (w2 * 32 + l2) < (w * 32 + lane))) ? 1 : 0;
          }
        if (rank == k && v[w] < INF) s_merge[0] = v[w];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) radii[row] = dist_f32(s_merge[0]);
    __syncthreads();
  }
}

// out[c] = max of v[32 c .. 32 c + 31]
__global__ void chunk_max_kernel(const float* __restrict__ v, long long n, float* __restrict__ out) {
  const long long c = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c * 32 >= n) return;
  float x = v[c * 32 + lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  if (lane == 0) out[c] = x;
}

// ---------------------------------------------------------- counts: thresholds
__global__ void prdc_thresholds_kernel(const float* __restrict__ norm, const float* __restrict__ rho,
                                       bool single_pass, const float* __restrict__ radii,
                                       long long n_valid, long long rows_pad,
                                       const float* __restrict__ other_max_norm, float* __restrict__ lo,
                                       float* __restrict__ hi) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= rows_pad) return;
  if (i >= n_valid) { lo[i] = -kInf; hi[i] = -kInf; return; }
  const float nx = norm[i];
  const float r = radii[i];
  // strict '<' (prdc.py:37-48): nothing is closer than a zero radius.  Rows with k+1 exact duplicates
  // have r = 0 and every one of their duplicate pairs would otherwise sit inside the band.
  if (!(r > 0.f)) { lo[i] = -kInf; hi[i] = -kInf; return; }
  const float r2 = r * r;
  // band on the key plus slack for the fp32 rounding of r^2, |x|^2 and of the
  // final fp32 distance the refine kernel compares (3e-7 ~ 2.5 ulp)
  const float band = filter_band(single_pass, nx, rho[i], other_max_norm) + 3e-7f * (r2 + nx);
  const float a = r2 - nx;
  lo[i] = a - band;
  hi[i] = a + band;
}

// -------------------------------------------------------------- counts: refine
template <typename T>
__global__ void __launch_bounds__(256)
prdc_exact_pairs_kernel(const T* __restrict__ R, long long ldr, const T* __restrict__ C, long long ldc, int d,
                        const float* __restrict__ r_ref, const float* __restrict__ r_cand,
                        const PairEntry* __restrict__ list, const unsigned long long* __restrict__ list_count,
                        unsigned long long list_cap, long long row0, int32_t* __restrict__ col_count,
                        uint8_t* __restrict__ row_recall, uint8_t* __restrict__ row_cover) {
  const int lane = threadIdx.x & 31;
  unsigned long long count = *list_count;
  if (count > list_cap) count = list_cap;
  const unsigned long long warps = (static_cast<unsigned long long>(gridDim.x) * blockDim.x) >> 5;
  for (unsigned long long e = (blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x) >> 5;
       e < count; e += warps) {
    const PairEntry p = list[e];
    const long long i = p.i;
    const long long j = p.j_kind & 0x7fffffffu;
    const bool cand_test = (p.j_kind >> 31) != 0;
    const double d2 = exact_sqdist_warp(R + i * ldr, C + j * ldc, d, lane);
    if (lane == 0) {
      const float D = dist_f32(d2);
      if (cand_test) {
        if (D < r_cand[j]) row_recall[i - row0] = 1;
      } else if (D < r_ref[i]) {
        atomicAdd(col_count + j, 1);
        row_cover[i - row0] = 1;
      }
    }
  }
}

// Exhaustive exact counts (no tensor-core filter, no list): the last rung of the overflow ladder —
// when even a refine list sized to the reported number of near-tie pairs does not fit in memory.
// Work item = (reference row, chunk of 1024 candidates), one warp each; the distance is formed by
// the same warp routine as the refine kernels, so a pair decided here or there gets the same answer.
template <typename T>
__global__ void __launch_bounds__(256)
prdc_bruteforce_counts_kernel(const T* __restrict__ R, long long ldr, const T* __restrict__ C, long long ldc, int d,
                              long long m, const float* __restrict__ r_ref, const float* __restrict__ r_cand,
                              long long row0, long long nrows, int32_t* __restrict__ col_count,
                              uint8_t* __restrict__ row_recall, uint8_t* __restrict__ row_cover) {
  const int lane = threadIdx.x & 31;
  const long long chunks = (m + 1023) / 1024;
  const long long items = nrows * chunks;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long it = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5; it < items; it += warps) {
    const long long row = it / chunks, j0 = (it - row * chunks) * 1024;
    const long long i = row0 + row;
    const float ri = r_ref[i];
    const long long j1 = j0 + 1024 < m ? j0 + 1024 : m;
    bool rec = false, cov = false;
    for (long long j = j0; j < j1; ++j) {
      const double d2 = exact_sqdist_warp(R + i * ldr, C + j * ldc, d, lane);
      const float D = dist_f32(d2);
      if (D < ri) {
        cov = true;
        if (lane == 0) atomicAdd(col_count + j, 1);
      }
      rec |= D < r_cand[j];
    }
    if (lane == 0) {
      if (rec) row_recall[row] = 1;
      if (cov) row_cover[row] = 1;
    }
  }
}

__global__ void prdc_reduce_kernel(const int32_t* __restrict__ col_count, long long m,
                                   const uint8_t* __restrict__ row_recall, const uint8_t* __restrict__ row_cover,
                                   long long nrows, unsigned long long* __restrict__ totals) {
  unsigned long long hit = 0, sum = 0, rec = 0, cov = 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long t0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (col_count)
    for (long long j = t0; j < m; j += stride) {
      const int c = col_count[j];
      hit += c > 0;
      sum += static_cast<unsigned long long>(c);
    }
  for (long long i = t0; i < nrows; i += stride) {
    if (row_recall) rec += row_recall[i] != 0;
    if (row_cover) cov += row_cover[i] != 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    hit += __shfl_xor_sync(0xffffffffu, hit, o);
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    rec += __shfl_xor_sync(0xffffffffu, rec, o);
    cov += __shfl_xor_sync(0xffffffffu, cov, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (hit) atomicAdd(totals + 0, hit);
    if (sum) atomicAdd(totals + 1, sum);
    if (rec) atomicAdd(totals + 2, rec);
    if (cov) atomicAdd(totals + 3, cov);
  }
}

// ------------------------------------------------------------------ host side
// Candidates kept per row: twice the k+1 that are needed.  The margin is what lets the
// refine kernel certify its answer: the (k+1)-th exact distance must sit below the Kt-th
// approximate key by more than the error band (rows that fail go to the brute-force
// kernel).  The sorted-insert cost of the epilogue grows like Kt^2, so Kt is not padded
// to a power of two.
static int pick_kt(int k) {
  if (k + 1 > 30) return 0;
  const int need = 2 * (k + 1);
  for (int kt : {8, 12, 16, 24, 32})
    if (need <= kt) return kt;
  return 32;   // 16 <= k <= 29: at least k + 3 candidates
}

// Column splits per row tile.  Items are numbered split-major and dealt round-robin
// to the persistent CTAs (pair_engine.cuh), so the whole grid sweeps one split
// at a time.
//   * Many row tiles (>= 2 per SM): one split.  Every CTA then walks all column
//     tiles at the MMA-bound pace, the grid stays in lockstep and each B tile is
//     fetched from HBM about once per wave (measured at 200k x 200k x 512:
//     5-8 GB of DRAM reads per launch, 98-99 % L2 hit rate, tensor pipe 91 %).
//     `l2_bytes` > 0 instead sizes splits so the B operand of a split is about
//     that many bytes (used by the count pass, which measures best that way).
//   * Few row tiles: split columns so every SM has work, choosing the split
//     count with the smallest round-robin makespan in column tiles.
constexpr long long kCountSplitBytes = 16ll << 20;
// `split_cost`: relative cost added per extra split.  The count epilogue is stateless (only the fixed
// per-item cost of one tile below); the radii epilogue restarts its candidate lists in every split,
// and the refine kernel merges 2 K candidates per split and row — measured at 200k columns, d = 512
// (profiles/r02_shard_sweeps.txt): about 7 % of the sweep per extra split.
static int pick_split(int dev, long long n_rt, long long n_ct, int kb_count, long long l2_bytes, double split_cost = 0.0) {
  const int sms = sm_count(dev);
  int s_min = 1;
  if (l2_bytes > 0) {
    const long long tile_bytes = 2ll * kTileN * kb_count * kBlockK * 2;
    long long tiles_per = l2_bytes / tile_bytes;
    if (tiles_per < 1) tiles_per = 1;
    const long long s = (n_ct + tiles_per - 1) / tiles_per;
    s_min = static_cast<int>(s > 64 ? 64 : s);
  }
  if (n_rt * s_min >= 2ll * sms) return s_min;
  // Makespan of the round-robin deal in column tiles, in closed form (this runs on the host inside
  // every call: a simulation of the deal cost milliseconds at 8 GPUs, where it was the critical path).
  // Split sp holds n_ct (sp + 1) / s - n_ct sp / s tiles, i.e. floor or ceil of n_ct / s; the busiest CTA
  // gets ceil(items / grid) items, each at most ceil(n_ct / s) tiles plus a fixed per-item cost of one.
  const int s_max = static_cast<int>(n_ct < 64 ? n_ct : 64);
  int best_s = s_min;
  long long best_span = -1;
  for (int s = s_min; s <= s_max; ++s) {
    const long long items = n_rt * s;
    const long long grid = items < sms ? items : sms;
    const long long rounds = (items + grid - 1) / grid;
    const long long tiles = (n_ct + s - 1) / s;
    // the last round is partial: CTAs that take part in it carry `rounds` items, the rest one fewer
    const long long span = static_cast<long long>(rounds * (tiles + 1) * (1.0 + split_cost * (s - 1)) * 16.0);
    if (best_span < 0 || span < best_span) { best_span = span; best_s = s; }
  }
  return best_s;
}

// The CTA-pair engine hands out work items dynamically, so the tail of a sweep is at most one
// item long: the count sweep is cut into four column splits to shorten it (measured 29.7 -> 28.0 ms
// at 200k x 200k).  The radii sweep is not: every split restarts its candidate lists, and the
// extra insertions cost more than the tail (30.9 -> 32.1 ms with two splits).
constexpr int kTailSplitTopk = 1;
constexpr int kTailSplitCount = 4;
static int tail_split(int mode, int n_split, int n_ct, int want) {
  if (mode != 2 || n_split != 1 || n_ct < 64 * want) return n_split;
  if (const int v = option(kOptTailSplit)) return v >= 1 && v <= want ? v : n_split;
  return want;
}

struct KnnWs {
  float* keys;
  int* cols;
  int* unresolved;
  int* n_unresolved;
  float* max_norm;
  size_t bytes;
};
static KnnWs knn_ws(void* ws, long long nrows, int Kt, int n_split) {
  const long long list_rows = round_up_ll(nrows, 2 * kTileM);   // the two-CTA engine works on row-tile pairs
  uint8_t* b = static_cast<uint8_t*>(ws);
  size_t off = 0;
  KnnWs w;
  auto take = [&](size_t bytes) { uint8_t* p = b ? b + off : nullptr; off += static_cast<size_t>(round_up_ll(bytes, 256)); return p; };
  // one candidate list per (column split, column half of the tile) and row
  w.keys = reinterpret_cast<float*>(take(static_cast<size_t>(2 * n_split) * list_rows * Kt * 4));
  w.cols = reinterpret_cast<int*>(take(static_cast<size_t>(2 * n_split) * list_rows * Kt * 4));
  w.unresolved = reinterpret_cast<int*>(take(static_cast<size_t>(nrows > 0 ? nrows : 1) * 4));
  w.n_unresolved = reinterpret_cast<int*>(take(256));
  w.max_norm = reinterpret_cast<float*>(take(256));
  w.bytes = off;
  return w;
}

// Engine choice for the filter sweeps: 3 = three-MMA split precision (any d), 1 = single MMA per
// product with a resident A panel (d <= 512), 2 = the same on CTA pairs (cta_group::2).
static int engine_mode(int kb_count, long long row0) {
  if (kb_count > kMaxResidentKb) return 3;
  if (option(kOptPasses) == 3) return 3;
  const bool two = option(kOptCta2) != 0;   // engine_cta2 = 0: single-CTA kernel
  // the pair kernel takes A row tiles two at a time: the shard must start on an even tile
  return (two && (row0 / kTileM) % 2 == 0) ? 2 : 1;
}

// `counter`: a zeroed device int for the CTA-pair kernel's dynamic item hand-out
// (option engine_static keeps the round-robin assignment).
template <class Epi>
static int run_engine(int mode, cudaStream_t st, int dev, EngineGeom g, const Epi& epi, const char* what,
                      double alg_pairs, int* counter) {
  if (mode == 3) return launch_engine(st, dev, g, epi, what, alg_pairs);
  if (mode == 1) return launch_engine1(st, dev, g, epi, what, alg_pairs);
  g.n_rt = (g.n_rt + 1) / 2;   // row-tile pairs
  g.work_counter = option(kOptSchedStatic) ? nullptr : counter;
  return launch_engine2(st, dev, g, epi, what, alg_pairs);
}

template <int K>
static int run_topk(cudaStream_t st, int dev, EngineGeom& g, const PackedPtrs& p, const KnnWs& w,
                    long long list_rows, long long a_row_base, double alg_pairs, int mode) {
  TopkEpi<K> epi{p.inv_scale, p.inv_scale, p.norm, p.cmin, w.keys, w.cols, list_rows, a_row_base};
  return run_engine(mode, st, dev, g, epi, "pair_engine<topk>", alg_pairs, w.n_unresolved + 16);   // zeroed with n_unresolved
}

}  // namespace amb

using namespace amb;

extern "C" {

size_t amb_knn_ws_bytes(long long nrows, long long n, int d, int k) {
  const int Kt = pick_kt(k);
  if (!Kt || nrows < 0 || n <= 0 || d <= 0) return 0;
  // upper bound of pick_split() without knowing the device: one split once there
  // are clearly >= 2 row tiles per SM (up to 160 SMs), else the cap of 64
  const long long n_rt = (nrows + kTileM - 1) / kTileM;
  const int bound = n_rt >= 2ll * 160 ? kTailSplitTopk : 64;
  return knn_ws(nullptr, nrows, Kt, bound).bytes;
}

int amb_knn_radii(int dev, amb_stream_t stream, const void* X, int dtype, long long ld, const void* packed,
                  long long n, int d, long long row0, long long nrows, int k, float* radii,
                  long long* n_exhaustive, void* ws, size_t ws_bytes) {
  if (!X || !packed || !radii || n <= 0 || d <= 0 || nrows < 0 || row0 < 0 || row0 + nrows > n || ld < d)
    return set_error(AMB_ERR_ARG, "amb_knn_radii: bad argument");
  if (k < 1 || k + 1 > n)
    return set_error(AMB_ERR_ARG, "amb_knn_radii: k=%d needs at least k+1 rows (n=%lld); the reference's "
                                  "kthvalue raises here too", k, n);
  if (nrows == 0) {   // an empty row shard (row0 == n is usually unaligned) is a valid no-op
    if (n_exhaustive) {
      DeviceGuard g0(dev);
      if (!g0.ok) return AMB_ERR_CUDA;
      return check_cuda(cudaMemsetAsync(n_exhaustive, 0, 8, static_cast<cudaStream_t>(stream)), "memset");
    }
    return AMB_OK;
  }
  if (row0 % kTileM != 0) return set_error(AMB_ERR_ARG, "amb_knn_radii: row0 must be a multiple of 128");
  if (n >= (1ll << 31)) return set_error(AMB_ERR_ARG, "amb_knn_radii: n must be < 2^31");
  const int Kt = pick_kt(k);
  if (!Kt) return set_error(AMB_ERR_ARG, "amb_knn_radii: k=%d not supported (k <= 29)", k);
  if (dtype != AMB_F32 && dtype != AMB_F64) return set_error(AMB_ERR_ARG, "amb_knn_radii: bad dtype");
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PackedPtrs p = packed_ptrs(const_cast<void*>(packed), n, d);
  const long long n_rt = (nrows + kTileM - 1) / kTileM;
  const long long n_ct = p.rows_pad / kTileN;
  int n_split = pick_split(dev, n_rt, n_ct, p.kb_count, 0, 0.07);
  if (const int v = option(kOptTopkSplit)) {   // tuning knob
    if (v >= 1 && v <= 64 && v <= n_ct) n_split = v;
  }
  n_split = tail_split(engine_mode(p.kb_count, row0), n_split, static_cast<int>(n_ct), kTailSplitTopk);
  KnnWs w = knn_ws(ws, nrows, Kt, n_split);
  if (!ws || ws_bytes < w.bytes) return set_error(AMB_ERR_WS, "amb_knn_radii: workspace %zu < %zu", ws_bytes, w.bytes);
  const long long list_rows = round_up_ll(nrows, 2 * kTileM);
  int rc = check_cuda(cudaMemsetAsync(w.n_unresolved, 0, 512, st), "memset");   // n_unresolved + max_norm
  if (rc) return rc;
  max_norm_kernel<<<64, 256, 0, st>>>(p.norm, p.rho, p.rows_pad, w.max_norm);
  if ((rc = check_launch("max_norm_kernel"))) return rc;
  const int mode = engine_mode(p.kb_count, row0);
  const bool sp = mode != 3;

  EngineGeom g{};
  g.a_planes = p.planes;
  g.b_planes = p.planes;
  g.a_plane_halfs = p.plane_halfs;
  g.b_plane_halfs = p.plane_halfs;
  g.kb_count = p.kb_count;
  g.a_rb_base = static_cast<int>(row0 / kTileM);
  g.n_problems = 1;
  g.n_rt = static_cast<int>(n_rt);
  g.n_ct = static_cast<int>(n_ct);
  g.n_split = n_split;
  g.lbo_bytes = 128;
  g.sbo_bytes = 512;
  const double alg_pairs = static_cast<double>(nrows) * n;
  if (Kt == 8) rc = run_topk<8>(st, dev, g, p, w, list_rows, row0, alg_pairs, mode);
  else if (Kt == 12) rc = run_topk<12>(st, dev, g, p, w, list_rows, row0, alg_pairs, mode);
  else if (Kt == 16) rc = run_topk<16>(st, dev, g, p, w, list_rows, row0, alg_pairs, mode);
  else if (Kt == 24) rc = run_topk<24>(st, dev, g, p, w, list_rows, row0, alg_pairs, mode);
  else rc = run_topk<32>(st, dev, g, p, w, list_rows, row0, alg_pairs, mode);
  if (rc) return rc;

  const unsigned blocks = static_cast<unsigned>((nrows * 32 + 255) / 256);
  const int bf_blocks = 4 * sm_count(dev);
  if (dtype == AMB_F32) {
    const float* Xf = static_cast<const float*>(X);
    knn_refine_kernel<float><<<blocks, 256, 0, st>>>(Xf, ld, d, n, row0, nrows, k, Kt, 2 * n_split, list_rows, w.keys,
                                                     w.cols, p.norm, p.rho, sp, w.max_norm, radii, w.unresolved,
                                                     w.n_unresolved);
    if ((rc = check_launch("knn_refine_kernel"))) return rc;
    knn_bruteforce_kernel<float><<<bf_blocks, 256, 0, st>>>(Xf, ld, d, n, row0, k, w.unresolved, w.n_unresolved, radii);
  } else {
    const double* Xd = static_cast<const double*>(X);
    knn_refine_kernel<double><<<blocks, 256, 0, st>>>(Xd, ld, d, n, row0, nrows, k, Kt, 2 * n_split, list_rows, w.keys,
                                                      w.cols, p.norm, p.rho, sp, w.max_norm, radii, w.unresolved,
                                                      w.n_unresolved);
    if ((rc = check_launch("knn_refine_kernel"))) return rc;
    knn_bruteforce_kernel<double><<<bf_blocks, 256, 0, st>>>(Xd, ld, d, n, row0, k, w.unresolved, w.n_unresolved, radii);
  }
  if ((rc = check_launch("knn_bruteforce_kernel"))) return rc;
  if (n_exhaustive) {   // int32 counter -> caller's int64
    if ((rc = check_cuda(cudaMemsetAsync(n_exhaustive, 0, 8, st), "memset"))) return rc;
    rc = check_cuda(cudaMemcpyAsync(n_exhaustive, w.n_unresolved, 4, cudaMemcpyDeviceToDevice, st), "memcpy");
  }
  return rc;
}

long long amb_prdc_list_cap(long long n_ref, long long m) {
  long long cap = 16 * (n_ref + m);
  return cap < (1ll << 20) ? (1ll << 20) : cap;
}

// bytes of the workspace in front of the refine list: thresholds, chunk maxima, 512-byte header
static size_t prdc_ws_fixed(long long n_ref, long long m) {
  const long long rp = round_up_ll(n_ref, kRowPad), cp = round_up_ll(m, kRowPad);
  return static_cast<size_t>(round_up_ll((2 * rp + 2 * cp + cp / 32) * 4, 256) + 512);
}

size_t amb_prdc_ws_bytes_cap(long long n_ref, long long m, long long list_cap) {
  if (n_ref <= 0 || m <= 0 || list_cap < 1) return 0;
  return prdc_ws_fixed(n_ref, m) + static_cast<size_t>(list_cap) * sizeof(PairEntry);
}

size_t amb_prdc_ws_bytes(long long n_ref, long long m) {
  return amb_prdc_ws_bytes_cap(n_ref, m, amb_prdc_list_cap(n_ref, m));
}

long long amb_prdc_ws_list_cap(long long n_ref, long long m, size_t ws_bytes) {
  if (n_ref <= 0 || m <= 0) return 0;
  const size_t fixed = prdc_ws_fixed(n_ref, m);
  return ws_bytes > fixed ? static_cast<long long>((ws_bytes - fixed) / sizeof(PairEntry)) : 0;
}

int amb_prdc_counts(int dev, amb_stream_t stream, const void* R, long long ldr, const void* packed_ref,
                    long long n_ref, const float* r_ref, const void* C, long long ldc,
                    const void* packed_cand, long long m, const float* r_cand, int d, int dtype,
                    long long row0, long long nrows, int32_t* col_count, uint8_t* row_recall,
                    uint8_t* row_cover, long long* n_uncertain, void* ws, size_t ws_bytes) {
  if (!R || !C || !packed_ref || !packed_cand || !r_ref || !r_cand || !col_count || !row_recall || !row_cover ||
      n_ref <= 0 || m <= 0 || d <= 0 || row0 < 0 || nrows < 0 || row0 + nrows > n_ref || ldr < d || ldc < d)
    return set_error(AMB_ERR_ARG, "amb_prdc_counts: bad argument");
  if (n_ref >= (1ll << 31) || m >= (1ll << 31)) return set_error(AMB_ERR_ARG, "amb_prdc_counts: sizes must be < 2^31");
  if (dtype != AMB_F32 && dtype != AMB_F64) return set_error(AMB_ERR_ARG, "amb_prdc_counts: bad dtype");
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nrows == 0) {   // an empty row shard (row0 == n_ref is usually unaligned) is a valid no-op
    return n_uncertain ? check_cuda(cudaMemsetAsync(n_uncertain, 0, 8, st), "memset") : AMB_OK;
  }
  if (row0 % kTileM != 0) return set_error(AMB_ERR_ARG, "amb_prdc_counts: row0 must be a multiple of 128");
  // the refine list takes whatever the workspace holds beyond the fixed part
  const long long cap = amb_prdc_ws_list_cap(n_ref, m, ws_bytes);
  if (!ws || cap < 1)
    return set_error(AMB_ERR_WS, "amb_prdc_counts: workspace %zu < %zu", ws_bytes, amb_prdc_ws_bytes_cap(n_ref, m, 1));
  PackedPtrs pr = packed_ptrs(const_cast<void*>(packed_ref), n_ref, d);
  PackedPtrs pc = packed_ptrs(const_cast<void*>(packed_cand), m, d);
  // workspace carve-up
  uint8_t* b = static_cast<uint8_t*>(ws);
  float* a_lo = reinterpret_cast<float*>(b);
  float* a_hi = a_lo + pr.rows_pad;
  float* b_lo = a_hi + pr.rows_pad;
  float* b_hi = b_lo + pc.rows_pad;
  float* cmax_bhi = b_hi + pc.rows_pad;   // [cols_pad / 32] chunk maxima of b_hi
  uint8_t* q = b + round_up_ll((2 * pr.rows_pad + 2 * pc.rows_pad + pc.rows_pad / 32) * 4, 256);
  unsigned long long* list_count = reinterpret_cast<unsigned long long*>(q);
  float* max_ref = reinterpret_cast<float*>(q + 64);
  float* max_cand = reinterpret_cast<float*>(q + 128);
  PairEntry* list = reinterpret_cast<PairEntry*>(q + 512);

  int rc;
  if ((rc = check_cuda(cudaMemsetAsync(q, 0, 512, st), "memset"))) return rc;
  if ((rc = check_cuda(cudaMemsetAsync(row_recall, 0, nrows, st), "memset"))) return rc;
  if ((rc = check_cuda(cudaMemsetAsync(row_cover, 0, nrows, st), "memset"))) return rc;
  const int mode = engine_mode(pr.kb_count, row0);
  const bool sp = mode != 3;
  max_norm_kernel<<<64, 256, 0, st>>>(pr.norm, pr.rho, pr.rows_pad, max_ref);
  if ((rc = check_launch("max_norm_kernel"))) return rc;
  max_norm_kernel<<<64, 256, 0, st>>>(pc.norm, pc.rho, pc.rows_pad, max_cand);
  if ((rc = check_launch("max_norm_kernel"))) return rc;
  prdc_thresholds_kernel<<<static_cast<unsigned>((pr.rows_pad + 255) / 256), 256, 0, st>>>(
      pr.norm, pr.rho, sp, r_ref, n_ref, pr.rows_pad, max_cand, a_lo, a_hi);
  if ((rc = check_launch("prdc_thresholds_kernel"))) return rc;
  prdc_thresholds_kernel<<<static_cast<unsigned>((pc.rows_pad + 255) / 256), 256, 0, st>>>(
      pc.norm, pc.rho, sp, r_cand, m, pc.rows_pad, max_ref, b_lo, b_hi);
  if ((rc = check_launch("prdc_thresholds_kernel"))) return rc;
  chunk_max_kernel<<<static_cast<unsigned>((pc.rows_pad + 255) / 256), 256, 0, st>>>(b_hi, pc.rows_pad, cmax_bhi);
  if ((rc = check_launch("chunk_max_kernel"))) return rc;

  EngineGeom g{};
  g.a_planes = pr.planes;
  g.b_planes = pc.planes;
  g.a_plane_halfs = pr.plane_halfs;
  g.b_plane_halfs = pc.plane_halfs;
  g.kb_count = pr.kb_count;
  g.a_rb_base = static_cast<int>(row0 / kTileM);
  g.n_problems = 1;
  g.n_rt = static_cast<int>((nrows + kTileM - 1) / kTileM);
  g.n_ct = static_cast<int>(pc.rows_pad / kTileN);
  // the single-pass kernel keeps the A panel in shared memory: no L2 pressure from A, one split
  g.n_split = pick_split(dev, g.n_rt, g.n_ct, pc.kb_count, sp ? 0 : kCountSplitBytes);
  g.n_split = tail_split(mode, g.n_split, g.n_ct, kTailSplitCount);
  if (const int v = option(kOptCountSplit)) {   // tuning knob
    if (v >= 1 && v <= 64 && v <= g.n_ct) g.n_split = v;
  }
  g.lbo_bytes = 128;
  g.sbo_bytes = 512;
  CountEpi epi{pr.inv_scale, pr.norm, a_lo, a_hi, pc.inv_scale, pc.norm, b_lo, b_hi, pc.cmin, cmax_bhi, col_count, row_recall,
               row_cover, row0, row0 + nrows, list, list_count, static_cast<unsigned long long>(cap)};
  if ((rc = run_engine(mode, st, dev, g, epi, "pair_engine<count>", static_cast<double>(nrows) * m,
                       reinterpret_cast<int*>(q + 192)))) return rc;   // inside the zeroed 512-byte header

  const int blocks = 8 * sm_count(dev);
  if (dtype == AMB_F32)
    prdc_exact_pairs_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(R), ldr,
                                                           static_cast<const float*>(C), ldc, d, r_ref, r_cand, list,
                                                           list_count, cap, row0, col_count, row_recall, row_cover);
  else
    prdc_exact_pairs_kernel<double><<<blocks, 256, 0, st>>>(static_cast<const double*>(R), ldr,
                                                            static_cast<const double*>(C), ldc, d, r_ref, r_cand, list,
                                                            list_count, cap, row0, col_count, row_recall, row_cover);
  if ((rc = check_launch("prdc_exact_pairs_kernel"))) return rc;
  if (n_uncertain)
    rc = check_cuda(cudaMemcpyAsync(n_uncertain, list_count, 8, cudaMemcpyDeviceToDevice, st), "memcpy");
  return rc;
}

int amb_prdc_counts_exact(int dev, amb_stream_t stream, const void* R, long long ldr, long long n_ref,
                          const float* r_ref, const void* C, long long ldc, long long m, const float* r_cand,
                          int d, int dtype, long long row0, long long nrows, int32_t* col_count,
                          uint8_t* row_recall, uint8_t* row_cover) {
  if (!R || !C || !r_ref || !r_cand || !col_count || !row_recall || !row_cover || n_ref <= 0 || m <= 0 || d <= 0 ||
      row0 < 0 || nrows < 0 || row0 + nrows > n_ref || ldr < d || ldc < d)
    return set_error(AMB_ERR_ARG, "amb_prdc_counts_exact: bad argument");
  if (dtype != AMB_F32 && dtype != AMB_F64) return set_error(AMB_ERR_ARG, "amb_prdc_counts_exact: bad dtype");
  if (nrows == 0) return AMB_OK;
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if ((rc = check_cuda(cudaMemsetAsync(row_recall, 0, nrows, st), "memset"))) return rc;
  if ((rc = check_cuda(cudaMemsetAsync(row_cover, 0, nrows, st), "memset"))) return rc;
  const int blocks = 8 * sm_count(dev);
  if (dtype == AMB_F32)
    prdc_bruteforce_counts_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(R), ldr,
                                                                 static_cast<const float*>(C), ldc, d, m, r_ref, r_cand,
                                                                 row0, nrows, col_count, row_recall, row_cover);
  else
    prdc_bruteforce_counts_kernel<double><<<blocks, 256, 0, st>>>(static_cast<const double*>(R), ldr,
                                                                  static_cast<const double*>(C), ldc, d, m, r_ref, r_cand,
                                                                  row0, nrows, col_count, row_recall, row_cover);
  return check_launch("prdc_bruteforce_counts_kernel");
}

int amb_prdc_reduce(int dev, amb_stream_t stream, const int32_t* col_count, long long m,
                    const uint8_t* row_recall, const uint8_t* row_cover, long long nrows,
                    long long* totals) {
  if (!totals || m < 0 || nrows < 0) return set_error(AMB_ERR_ARG, "amb_prdc_reduce: bad argument");
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  prdc_reduce_kernel<<<2 * sm_count(dev), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      col_count, m, row_recall, row_cover, nrows, reinterpret_cast<unsigned long long*>(totals));
  return check_launch("prdc_reduce_kernel");
}

}  // extern "C"

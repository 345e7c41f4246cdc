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
#include <cstdlib>
#include <mutex>

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

// Block one-sided Jacobi for d <= 512: the d columns are cut into blocks of BS; a CTA
// takes one block pair (I, J) into shared memory (2*BS rows of d doubles), one warp per
// column pair, and runs BS rounds that rotate every column of I against every column of J
// between two __syncthreads — a grid barrier is only needed per BLOCK round, nb-1 times
// per sweep instead of d-1 times.  Phase 0 of a sweep rotates the pairs inside each
// block (round robin inside I and inside J), phases 1..nb-1 the cross pairs of a
// round-robin tournament over blocks: every column pair is visited exactly once per
// sweep, which is a cyclic Jacobi ordering.  Same rotations, same convergence test as
// jacobi_kernel; only the order of the pairs differs.
template <int BS, int NR>
__global__ void __launch_bounds__(32 * BS)
jacobi_block_kernel(double* __restrict__ Gt, int d, int n_mat, double tol, int stop_rot,
                    int* __restrict__ rot_count, int* __restrict__ sweeps_done) {
  extern __shared__ double slab[];                 // [2*BS][d]
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;               // 0 .. BS-1
  int nb = (d + BS - 1) / BS;
  nb += nb & 1;                                    // even number of blocks (the last may be virtual)
  const int pairs_per_mat = nb / 2;
  const int total = pairs_per_mat * n_mat;
  const double tol2 = tol * tol;
  int sweep = 0;
  for (; sweep < kJacobiMaxSweeps; ++sweep) {
    int my_rot = 0;
    for (int phase = 0; phase < nb; ++phase) {
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int mat = w / pairs_per_mat, s = w - mat * pairs_per_mat;
        int bi, bj;
        if (phase == 0) { bi = 2 * s; bj = 2 * s + 1; }
        else tournament_pair(nb, phase - 1, s, bi, bj);
        double* base = Gt + static_cast<long long>(mat) * d * d;
        // ---- load the two blocks (rows >= d are virtual: zeros, never stored)
        for (int k = threadIdx.x; k < d; k += blockDim.x) {
#pragma unroll 8
          for (int r = 0; r < 2 * BS; ++r) {
            const int row = (r < BS ? bi * BS + r : bj * BS + (r - BS));
            slab[r * d + k] = row < d ? __ldcg(base + static_cast<long long>(row) * d + k) : 0.0;
          }
        }
        __syncthreads();
        const int inner = phase == 0 ? BS - 1 : BS;
        for (int ir = 0; ir < inner; ++ir) {
          int p, q;   // slab rows of this warp's pair
          if (phase == 0) {
            // two independent round robins, one inside each block: warps [0,BS/2) on I, the rest on J
            const int blk = warp / (BS / 2), sl = warp % (BS / 2);
            tournament_pair(BS, ir, sl, p, q);
            p += blk * BS;
            q += blk * BS;
          } else {
            p = warp;
            q = BS + ((warp + ir) % BS);
          }
          double* gp = slab + p * d;
          double* gq = slab + q * d;
          double a = 0.0, b = 0.0, c = 0.0;
          double xr[NR], yr[NR];
#pragma unroll
          for (int e = 0; e < NR; ++e) {
            const int k = lane + 32 * e;
            xr[e] = k < d ? gp[k] : 0.0;
            yr[e] = k < d ? gq[k] : 0.0;
            a = fma(xr[e], xr[e], a);
            b = fma(yr[e], yr[e], b);
            c = fma(xr[e], yr[e], c);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
          }
          // rotate when |c| > tol sqrt(a b); with zeta = (b-a)/(2c):
          //   t = sgn(zeta)/(|zeta| + sqrt(1+zeta^2)) = sgn(b-a) 2c / (|b-a| + sqrt((b-a)^2 + 4c^2))
          // (one sqrt, one divide, one rsqrt instead of three of each)
          const double ab = a * b;
          if (ab > 0.0 && c * c > tol2 * ab) {
            const double dl = b - a;
            const double h = sqrt(fma(dl, dl, 4.0 * c * c));
            double t = (2.0 * c) / (fabs(dl) + h);
            t = dl < 0.0 ? -t : t;
            const double cs = rsqrt(fma(t, t, 1.0));
            const double sn = cs * t;
#pragma unroll
            for (int e = 0; e < NR; ++e) {
              const int k = lane + 32 * e;
              if (k < d) {
                gp[k] = cs * xr[e] - sn * yr[e];
                gq[k] = sn * xr[e] + cs * yr[e];
              }
            }
            ++my_rot;
          }
          __syncthreads();
        }
        // ---- store
        for (int k = threadIdx.x; k < d; k += blockDim.x) {
#pragma unroll 8
          for (int r = 0; r < 2 * BS; ++r) {
            const int row = (r < BS ? bi * BS + r : bj * BS + (r - BS));
            if (row < d) base[static_cast<long long>(row) * d + k] = slab[r * d + k];
          }
        }
        __syncthreads();
      }
      grid.sync();
    }
    if (lane == 0 && my_rot) atomicAdd(&rot_count[sweep], my_rot);
    grid.sync();
    // Converged when a sweep rotated (almost) nothing: every pair it visited was below tol or
    // was made orthogonal by its rotation, and the few rotations of such a sweep are by angles of
    // the order of tol, so they re-mix the other columns only to second order.
    if (*reinterpret_cast<volatile int*>(&rot_count[sweep]) <= stop_rot) { ++sweep; break; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *sweeps_done = sweep;
}

// Pivoted (diagonal pivoting) Cholesky  S = L L^T  of symmetric positive SEMI-definite
// matrices, blocked right-looking, without physical row/column swaps: step j takes the
// not-yet-used index p with the largest residual diagonal, forms
//   L[:, j] = (A[:, p] - Lpanel[:, :jj] Lpanel[p, :jj]^T) / sqrt(diag_p)
// and stops at the numerical rank (diag_p <= d * eps * max diag, as LAPACK dpstrf), so a
// rank-deficient covariance (N < d, rank-1 embedders) gives a factor with exactly `rank`
// non-zero columns and no square roots of round-off.  L need not be triangular for what
// follows — any factor with S = L L^T has the singular values the trace needs.
// A: [n_mat][d][d] (overwritten by the trailing updates);  Lt: [n_mat][d][d], ZEROED by the
// caller, row j = column j of L.  Grid = n_mat * C CTAs (cooperative): CTA 0 of each group
// factors the PB-column panel out of shared memory, then all C CTAs of the group apply
// A -= Lp Lp^T to their row slice.
__global__ void __launch_bounds__(512)
pchol_kernel(double* __restrict__ A, double* __restrict__ Lt, int d, int n_mat, int C, int PB,
             int* __restrict__ panel_n /* [n_mat] */, int* __restrict__ rank_out /* [n_mat] */) {
  extern __shared__ double sm[];
  double* Lp = sm;                      // [PB][d]
  double* dg = sm + static_cast<size_t>(PB) * d;   // [d] residual diagonal (-inf: used)
  __shared__ double red_v[16];
  __shared__ unsigned long long red_key[32];   // [step parity][warp]
  __shared__ double red_val[32];
  cg::grid_group grid = cg::this_grid();
  const int mat = blockIdx.x / C, c = blockIdx.x % C;
  const bool leader = c == 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double NEG = -__longlong_as_double(0x7ff0000000000000ll);
  double* Am = A + static_cast<long long>(mat) * d * d;
  double* Ltm = Lt + static_cast<long long>(mat) * d * d;
  double tol = 0.0;
  if (leader) {
    double mx = 0.0;
    for (int i = tid; i < d; i += blockDim.x) {
      const double v = Am[static_cast<long long>(i) * d + i];
      dg[i] = v;
      mx = fmax(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red_v[warp] = mx;
    __syncthreads();
    mx = 0.0;
    for (int w = 0; w < 16; ++w) mx = fmax(mx, red_v[w]);
    tol = static_cast<double>(d) * 2.220446049250313e-16 * mx;
    __syncthreads();
  }
  const int n_panels = (d + PB - 1) / PB;
  int j0 = 0;
  bool finished = false;
  for (int pan = 0; pan < n_panels; ++pan) {
    if (leader) {
      int npan = 0;
      if (!finished) {
        // ONE block barrier per pivot step (three before, at 2.5 us per step the dominant cost of the
        // factorisation).  The arg max travels as a 64-bit key — the bits of the residual diagonal with
        // the index in the 11 lowest mantissa bits, so the choice of pivot may be off by 2^-42 — reduced
        // by warp shuffles; the lane that held a warp's best also publishes its exact value, and every
        // thread scans the 16 warp results itself.  The result arrays are double-buffered by step
        // parity, so the writes of step jj + 1 cannot overtake a slow reader of step jj; the column and
        // diagonal writes of step jj are ordered before their readers in step jj + 1 by the same barrier.
        for (int jj = 0; jj < PB && j0 + jj < d; ++jj) {
          unsigned long long key = 0ull;
          double val = NEG;
          for (int i = tid; i < d; i += blockDim.x) {
            const double v = dg[i];
            const unsigned long long kb =
                v > 0.0 ? ((static_cast<unsigned long long>(__double_as_longlong(v)) & ~0x7ffull) | static_cast<unsigned>(i)) : 0ull;
            if (kb > key) { key = kb; val = v; }
          }
          const unsigned long long mine = key;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
            key = ok > key ? ok : key;
          }
          unsigned long long* red_k = red_key + (jj & 1) * 16;
          double* red_d = red_val + (jj & 1) * 16;
          if (lane == 0) red_k[warp] = key;
          if (mine == key && key != 0ull) red_d[warp] = val;     // unique: the index is part of the key
          __syncthreads();
          int best_w = 0;
          key = red_k[0];
#pragma unroll
          for (int w = 1; w < 16; ++w)
            if (red_k[w] > key) { key = red_k[w]; best_w = w; }
          const int pv = static_cast<int>(key & 0x7ffull);
          const double piv = key ? red_d[best_w] : NEG;
          if (!(piv > tol)) { finished = true; break; }   // numerical rank reached (uniform)
          const double root = sqrt(piv);
          const double rs = 1.0 / root;
          for (int i = tid; i < d; i += blockDim.x) {
            const double a = __ldcg(Am + static_cast<long long>(pv) * d + i);   // in flight during the dot product
            double s0 = 0.0, s1 = 0.0;
            int k = 0;
            for (; k + 1 < jj; k += 2) {
              s0 = fma(Lp[k * d + i], Lp[k * d + pv], s0);
              s1 = fma(Lp[(k + 1) * d + i], Lp[(k + 1) * d + pv], s1);
            }
            if (k < jj) s0 = fma(Lp[k * d + i], Lp[k * d + pv], s0);
            const double v = a - (s0 + s1);
            const double di = dg[i];                    // own entries only: no other thread reads dg[i]
            double col = di == NEG ? 0.0 : v * rs;      // rows already used: exactly zero
            if (i == pv) { col = root; dg[i] = NEG; }
            else if (di != NEG) dg[i] = di - col * col;
            Lp[jj * d + i] = col;
          }
          ++npan;
        }
        __syncthreads();
        for (int e = tid; e < npan * d; e += blockDim.x)
          Ltm[static_cast<long long>(j0) * d + e] = Lp[e];
      }
      if (tid == 0) panel_n[mat] = npan;
    }
    grid.sync();
    const int npan = *reinterpret_cast<volatile int*>(panel_n + mat);
    if (npan > 0 && pan + 1 < n_panels) {
      if (!leader) {
        for (int e = tid; e < npan * d; e += blockDim.x)
          Lp[e] = __ldcg(Ltm + static_cast<long long>(j0) * d + e);
      }
      __syncthreads();
      const int rows_per = (d + C - 1) / C;
      const int r0 = c * rows_per;
      const int r1 = min(d, r0 + rows_per);
      for (int k = tid; k < d; k += blockDim.x) {
        for (int i = r0; i < r1; i += 4) {
          double acc[4] = {0.0, 0.0, 0.0, 0.0};
          for (int jj = 0; jj < npan; ++jj) {
            const double l = Lp[jj * d + k];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (i + u < r1) acc[u] = fma(Lp[jj * d + i + u], l, acc[u]);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (i + u < r1) Am[static_cast<long long>(i + u) * d + k] -= acc[u];
        }
      }
    }
    j0 += npan;
    if (pan + 1 < n_panels) grid.sync();
  }
  if (leader && tid == 0) rank_out[mat] = j0;
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

// ---------------------------------------------------------------------------
// Nuclear norm by polar iteration (the default path).
//
// With the polar decomposition M = U H (U orthogonal on the range of M, H symmetric PSD),
// c = sum of singular values = tr H = tr(U^T M) = <U, M>_F.  U is the limit of
//     X_0 = M / |M|_F,     X_{k+1} = X_k (a_k I + b_k A_k + c_k A_k^2),   A_k = X_k^T X_k,
// which acts on every singular value s of X_k as the odd quintic p_k(s) = s (a + b s^2 + c s^4)
// and leaves the singular vectors alone.  The p_k (tools/polar_schedule.py) are chosen greedily
// so that [l_k, 1 + 1/64] is mapped into [l_{k+1}, 1] with the largest l_{k+1}; from l_0 = 1e-12
// nineteen of them reach 0.52, four classical quintic Newton-Schulz steps then reach 1 - 7e-16.
// Error of the result: a singular value s_i < 1e-12 |M|_F is not fully pushed to 1 and its
// contribution s_i f_i is short by at most s_i, so c is short by at most d 1e-12 |M|_F <= 5e-10 c
// at d = 512; rotational errors of U only enter to second order because tr(U^T M) is maximal at
// the polar factor; measured against an fp64 SVD: 1e-13 relative on full-rank, rank-deficient
// (N < d), rank-1 and column-scaled (condition 4e13) inputs.  Nothing is inverted, so singular M
// (zero singular values stay zero) needs no special care — the reason Newton-Schulz on the
// covariances themselves was rejected does not apply to the factored form.
// Cost: 23 x 3 fp64 GEMMs of d^3 (0.13 GFMA each at d = 512) on the FP64 pipe, all SMs: no
// sequential d-step chain like the Jacobi sweeps (12 sweeps x 64 grid-wide block rounds).
constexpr int kPolarSteps = 23;
__constant__ double kPolarCoef[kPolarSteps][3] = {
    {4.1916565787041771, -12.066380161025515, 8.68377021423718},
    {4.1916566436514273, -12.066380317469966, 8.6837703048631969},
    {4.1916566436514655, -12.066380317470065, 8.6837703048632555},
    {4.1916566436516245, -12.06638031747045, 8.6837703048634811},
    {4.1916566436522835, -12.066380317472056, 8.6837703048644173},
    {4.1916566386695626, -12.066380276778837, 8.6837702700967636},
    {4.191656622769198, -12.066380146906818, 8.683770159134296},
    {4.1916565820088225, -12.066379813873592, 8.6837698745787613},
    {4.1916562767511243, -12.066377320678535, 8.6837677444089749},
    {4.1916551057316553, -12.066367755953339, 8.6837595723219021},
    {4.1916500759715642, -12.066327371902723, 8.6837251486053741},
    {4.1916296475970114, -12.06615981794816, 8.6835819104574838},
    {4.1915433819363512, -12.065455223128238, 8.6829799069705818},
    {4.1911819164247035, -12.062503019398758, 8.6804575710918428},
    {4.1896658996547913, -12.050131190543864, 8.6698883400076188},
    {4.1833118675549503, -11.998362821360594, 8.6256725368931271},
    {4.1566694961648514, -11.782935523996235, 8.4418630726327937},
    {4.0452998092620023, -10.910521325149281, 7.7007586089770745},
    {3.6064439676221012, -7.8935092572868069, 5.1883297840173643},
    {15.0 / 8, -10.0 / 8, 3.0 / 8},
    {15.0 / 8, -10.0 / 8, 3.0 / 8},
    {15.0 / 8, -10.0 / 8, 3.0 / 8},
    {15.0 / 8, -10.0 / 8, 3.0 / 8},
};

constexpr int kPgM = 32, kPgN = 64, kPgK = 16, kPgThreads = 256, kPgStages = 4;
constexpr int kPolarPad = 64;   // iteration matrices are dp x dp, dp = d rounded up to 64, zero padded
constexpr int kPgAStride = kPgM + 2;    // doubles per k row of the A tile, TN form ([k][i]); even: 16 B aligned rows
constexpr int kPgA2Stride = kPgK + 2;   // doubles per i row of the A tile, NN form ([i][k])
constexpr int kPgBStride = kPgN + 2;
constexpr int kPgATile = (kPgK * kPgAStride > kPgM * kPgA2Stride ? kPgK * kPgAStride : kPgM * kPgA2Stride);
constexpr int kPgStageDoubles = kPgATile + kPgK * kPgBStride;
constexpr size_t kPgSmemBytes = size_t(kPgStages) * kPgStageDoubles * sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One CTA per matrix: alpha = |Mt|_F, X = Mt / alpha written into the zero-padded dp x dp buffer.
__global__ void __launch_bounds__(1024)
polar_init_kernel(const double* __restrict__ Mt, int d, int dp, double* __restrict__ X) {
  __shared__ double red[32];
  __shared__ double s_inv;
  const long long mo = static_cast<long long>(blockIdx.x) * d * d;
  const long long xo = static_cast<long long>(blockIdx.x) * dp * dp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  double s = 0.0;
  for (long long e = threadIdx.x; e < static_cast<long long>(d) * d; e += blockDim.x) {
    const double v = Mt[mo + e];
    s = fma(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < n_warps; ++w) t += red[w];
    s_inv = t > 0.0 ? 1.0 / sqrt(t) : 0.0;
  }
  __syncthreads();
  const double inv = s_inv;
  for (int i = warp; i < dp; i += n_warps) {     // a warp per row: no integer divisions
    double* x = X + xo + static_cast<long long>(i) * dp;
    const double* m = Mt + mo + static_cast<long long>(i) * d;
    for (int j = lane; j < dp; j += 32) x[j] = (i < d && j < d) ? m[j] * inv : 0.0;
  }
}

// fp64 GEMM tile kernel of the polar iteration on padded dp x dp row-major matrices (batch in z):
//   MODE 0:  C = X^T X                            C[i][j] = sum_k X[k][i] X[k][j]
//   MODE 1:  C = ca I + cb X + cc X^T X           (X = A symmetric: the quintic's matrix Q from A and A^2,
//                                                  formed in the epilogue so that no GEMM loads need arithmetic)
//   MODE 2:  C = X Q                              C[i][j] = sum_k X[i][k] Q[k][j]
// 32 x 64 output tile per CTA, k blocks of 16 through a four-stage cp.async ring (every operand tile
// is a plain copy of 16-byte pieces, so the L2 latency of a stage hides behind three stages of FMAs);
// 256 threads = two groups of four warps that split each k block between them (even / odd k) and
// add their 4 x 4 accumulators at the end: two warps per scheduler instead of one keep the FP64 pipe
// fed across shared-memory and barrier latencies.  A thread's columns are {2 tx, 2 tx + 1, 32 + 2 tx,
// 33 + 2 tx}: its two 16-byte reads of a B row are conflict free.  At d = 512: 128 CTAs, one wave.
template <int MODE>
__global__ void __launch_bounds__(kPgThreads, 1)
polar_gemm_kernel(const double* __restrict__ X, const double* __restrict__ Q, double* __restrict__ C, int dp, int step) {
  extern __shared__ __align__(16) double pg_smem[];
  const long long mo = static_cast<long long>(blockIdx.z) * dp * dp;
  X += mo;
  C += mo;
  if (MODE == 2) Q += mo;
  const int i0 = blockIdx.y * kPgM, j0 = blockIdx.x * kPgN;
  const int tid = threadIdx.x, grp = tid >> 7, t = tid & 127, ty = t >> 4, tx = t & 15;
  const double* Bsrc = MODE == 2 ? Q : X;
  // this thread's 16-byte pieces of a stage: one of the A tile, two of the B tile
  const int a_row = MODE == 2 ? (tid >> 3) : (tid >> 4);          // NN: 32 rows x 8 pieces; TN: 16 rows x 16 pieces
  const int a_col = MODE == 2 ? (tid & 7) * 2 : (tid & 15) * 2;   // in doubles
  const int b_row = tid >> 4, b_col = (tid & 15) * 4;             // 16 rows x 32 pieces, two adjacent per thread
  auto issue = [&](int kb) {
    double* st = pg_smem + static_cast<size_t>(kb % kPgStages) * kPgStageDoubles;
    const int k0 = kb * kPgK;
    if (MODE == 2) cp_async16(st + a_row * kPgA2Stride + a_col, X + static_cast<long long>(i0 + a_row) * dp + k0 + a_col);
    else cp_async16(st + a_row * kPgAStride + a_col, X + static_cast<long long>(k0 + a_row) * dp + i0 + a_col);
    double* bs = st + kPgATile + b_row * kPgBStride + b_col;
    const double* bg = Bsrc + static_cast<long long>(k0 + b_row) * dp + j0 + b_col;
    cp_async16(bs, bg);
    cp_async16(bs + 2, bg + 2);
  };
  double acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
  const int n_kb = dp / kPgK;
#pragma unroll
  for (int s = 0; s < kPgStages - 1; ++s) {
    if (s < n_kb) issue(s);
    cp_async_commit();
  }
  for (int kb = 0; kb < n_kb; ++kb) {
    cp_async_wait<kPgStages - 2>();     // this thread's pieces of stage kb have landed ...
    __syncthreads();                    // ... everyone's have, and everyone is done with stage kb - 1
    if (kb + kPgStages - 1 < n_kb) issue(kb + kPgStages - 1);   // refill the buffer stage kb - 1 used
    cp_async_commit();
    const double* st = pg_smem + static_cast<size_t>(kb % kPgStages) * kPgStageDoubles;
    const double* bs = st + kPgATile;
    // this group's eight k steps of the block, four at a time: all operand loads of a half are issued
    // before its 64 FMAs, so the shared-memory latency is paid once per half instead of once per step
    // (with two warps per scheduler the loads of one step could not hide behind the FMAs of another)
#pragma unroll
    for (int hq = 0; hq < 2; ++hq) {
      double a[4][4], b[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kk = 2 * (4 * hq + q) + grp;
        if (MODE == 2) {
#pragma unroll
          for (int r = 0; r < 4; ++r) a[q][r] = st[(ty * 4 + r) * kPgA2Stride + kk];
        } else {
          const double2 a0 = *reinterpret_cast<const double2*>(st + kk * kPgAStride + ty * 4);
          const double2 a1 = *reinterpret_cast<const double2*>(st + kk * kPgAStride + ty * 4 + 2);
          a[q][0] = a0.x; a[q][1] = a0.y; a[q][2] = a1.x; a[q][3] = a1.y;
        }
        const double2 b0 = *reinterpret_cast<const double2*>(bs + kk * kPgBStride + 2 * tx);
        const double2 b1 = *reinterpret_cast<const double2*>(bs + kk * kPgBStride + 32 + 2 * tx);
        b[q][0] = b0.x; b[q][1] = b0.y; b[q][2] = b1.x; b[q][3] = b1.y;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[q][r], b[q][c], acc[r][c]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // add the two k groups (through the now idle ring) and write the tile
  double* red = pg_smem;   // [128][16]
  if (grp == 1) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) red[(r * 4 + c) * 128 + t] = acc[r][c];
  }
  __syncthreads();
  if (grp == 0) {
    const double ca = MODE == 1 ? kPolarCoef[step][0] : 0.0;
    const double cb = MODE == 1 ? kPolarCoef[step][1] : 0.0;
    const double cc = MODE == 1 ? kPolarCoef[step][2] : 1.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int gi = i0 + ty * 4 + r;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int gj = j0 + 32 * h + 2 * tx;
        double v0 = acc[r][2 * h] + red[(r * 4 + 2 * h) * 128 + t];
        double v1 = acc[r][2 * h + 1] + red[(r * 4 + 2 * h + 1) * 128 + t];
        if (MODE == 1) {
          const double2 x = *reinterpret_cast<const double2*>(X + static_cast<long long>(gi) * dp + gj);
          v0 = fma(cc, v0, cb * x.x) + (gi == gj ? ca : 0.0);
          v1 = fma(cc, v1, cb * x.y) + (gi == gj + 1 ? ca : 0.0);
        }
        *reinterpret_cast<double2*>(C + static_cast<long long>(gi) * dp + gj) = make_double2(v0, v1);
      }
    }
  }
}

// U (dp-strided) <- polar factor of Mt (d x d, batch matrices).  bufs: 4 * batch * dp * dp doubles.
// Returns the buffer that holds U in *u_out.
static int launch_polar(cudaStream_t st, const double* Mt, int d, int batch, double* bufs, const double** u_out) {
  const int dp = static_cast<int>(round_up_ll(d, kPolarPad));
  const size_t mat = static_cast<size_t>(dp) * dp * batch;
  double* X = bufs;
  double* Xn = bufs + mat;
  double* A = bufs + 2 * mat;
  double* Qm = bufs + 3 * mat;
  int rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  auto set_attr = [] {
    cudaError_t e = cudaFuncSetAttribute(polar_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kPgSmemBytes));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(polar_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kPgSmemBytes));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(polar_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kPgSmemBytes));
    return e;
  };
  std::call_once(once, [&] { attr_err = set_attr(); });
  if (attr_err == cudaSuccess) attr_err = set_attr();   // the attribute is per device
  if ((rc = check_cuda(attr_err, "cudaFuncSetAttribute(polar_gemm)"))) return rc;
  polar_init_kernel<<<batch, 1024, 0, st>>>(Mt, d, dp, X);
  if ((rc = check_launch("polar_init_kernel"))) return rc;
  const dim3 grid(dp / kPgN, dp / kPgM, batch);
  for (int k = 0; k < kPolarSteps; ++k) {
    polar_gemm_kernel<0><<<grid, kPgThreads, kPgSmemBytes, st>>>(X, nullptr, A, dp, k);     // A = X^T X
    if ((rc = check_launch("polar_gemm_kernel<0>"))) return rc;
    polar_gemm_kernel<1><<<grid, kPgThreads, kPgSmemBytes, st>>>(A, nullptr, Qm, dp, k);    // Q = a I + b A + c A^2
    if ((rc = check_launch("polar_gemm_kernel<1>"))) return rc;
    polar_gemm_kernel<2><<<grid, kPgThreads, kPgSmemBytes, st>>>(X, Qm, Xn, dp, k);         // X <- X Q
    if ((rc = check_launch("polar_gemm_kernel<2>"))) return rc;
    double* tmp = X; X = Xn; Xn = tmp;
  }
  *u_out = X;
  return AMB_OK;
}

// one block per pair: a = |mu_x-mu_y|^2, b = tr S_x + tr S_y, c = <U, M>_F  (U: dp-strided polar factor)
__global__ void __launch_bounds__(1024)
fad_combine_polar_kernel(int d, int dp, const double* __restrict__ mu_x, const double* __restrict__ cov_x,
                         const double* __restrict__ mu_y, const double* __restrict__ cov_y,
                         const double* __restrict__ Mt, const double* __restrict__ U, double* __restrict__ out) {
  __shared__ double red[32];
  const int b = blockIdx.x;
  const long long mo = static_cast<long long>(b) * d * d;
  const long long uo = static_cast<long long>(b) * dp * dp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  double acc = 0.0;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    const double df = mu_x[static_cast<long long>(b) * d + k] - mu_y[static_cast<long long>(b) * d + k];
    acc += df * df + cov_x[mo + static_cast<long long>(k) * d + k] + cov_y[mo + static_cast<long long>(k) * d + k];
  }
  double c = 0.0;   // a warp per row: coalesced, no integer divisions, fixed order
  for (int i = warp; i < d; i += n_warps) {
    const double* u = U + uo + static_cast<long long>(i) * dp;
    const double* m = Mt + mo + static_cast<long long>(i) * d;
    for (int j = lane; j < d; j += 32) c = fma(u[j], m[j], c);
  }
  acc -= 2.0 * c;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < n_warps; ++w) t += red[w];
    out[b] = t;
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

template <int BS, int NR>
static int launch_jacobi_block(cudaStream_t st, int dev, double* Gt, int d, int n_mat, double tol, int stop_rot,
                               int* rot, int* sweeps) {
  const size_t smem = static_cast<size_t>(2) * BS * d * sizeof(double);
  auto* fn = jacobi_block_kernel<BS, NR>;
  int rc = check_cuda(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)),
                      "cudaFuncSetAttribute(jacobi_block)");
  if (rc) return rc;
  int per_sm = 0;
  if ((rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 32 * BS, smem), "occupancy"))) return rc;
  if (per_sm < 1) return set_error(AMB_ERR_CUDA, "jacobi_block_kernel cannot be resident");
  int nb = (d + BS - 1) / BS;
  nb += nb & 1;
  long long want = static_cast<long long>(nb / 2) * n_mat;
  const long long cap = static_cast<long long>(per_sm) * sm_count(dev);
  if (want > cap) want = cap;
  if (option(kOptFadCtas) > 0 && want > option(kOptFadCtas)) want = option(kOptFadCtas);   // CTAs loop over the block pairs
  void* args[] = {&Gt, &d, &n_mat, &tol, &stop_rot, &rot, &sweeps};
  rc = check_cuda(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(fn), dim3(static_cast<unsigned>(want)),
                                              dim3(32 * BS), args, smem, st),
                  "cudaLaunchCooperativeKernel(jacobi_block_kernel)");
  if (rc) return rc;
  return check_launch("jacobi_block_kernel");
}

// `values_only`: the caller needs the singular values (column norms) but not the rotated
// columns themselves.  The sum of the column norms is stationary at convergence — its error is
// second order in the remaining cosines (for two columns (|a|+|b|)^2 - (s1+s2)^2 = 2|a||b|(1 - sqrt(1-c^2)))
// — so cosines below 1e-8 leave it exact to fp64 round-off, and the sweep that only confirms
// convergence can be dropped.  Eigen-FACTORS (the AMB_FAD_FACTOR=eig path) are first order in the
// cosines and keep the 1e-14 test and the confirming sweep.
int launch_jacobi(cudaStream_t st, int dev, double* Gt, int d, int n_mat, int* counters, bool values_only) {
  double tol = values_only ? 1e-8 : 1e-14;
  int stop_rot = values_only ? d / 8 : 0;
  int* rot = counters;
  int* sweeps = counters + kJacobiMaxSweeps;
  int rc = check_cuda(cudaMemsetAsync(counters, 0, (kJacobiMaxSweeps + 1) * sizeof(int), st), "memset");
  if (rc) return rc;
  const bool flat = option(kOptJacobiFlat) != 0;   // forces the round-per-grid-barrier kernel
  if (d <= 512 && !flat) {
    // block size: 16 columns per block unless that leaves most SMs idle (few matrices), then 8
    // (measured at d=512, one matrix: 15.9 / 13.1 / 15.7 ms for 16 / 8 / 4): more, smaller CTAs in
    // flight at the price of more grid barriers per sweep
    int bs = 16;
    const int sms = sm_count(dev);
    while (bs > 8 && static_cast<long long>(n_mat) * ((d + 2 * bs - 1) / (2 * bs)) * 2 <= sms) bs >>= 1;
    if (option(kOptFadCtas) > 0) {   // stay within the CTA budget: larger blocks = fewer CTAs
      bs = 16;
      while (bs > 4 && static_cast<long long>(n_mat) * ((d + bs - 1) / bs) <= option(kOptFadCtas)) bs >>= 1;
    }
    if (option(kOptJacobiBlock)) bs = option(kOptJacobiBlock);
#define AMB_JB(BS)                                                                                   \
    (d <= 128 ? launch_jacobi_block<BS, 4>(st, dev, Gt, d, n_mat, tol, stop_rot, rot, sweeps)                    \
     : d <= 256 ? launch_jacobi_block<BS, 8>(st, dev, Gt, d, n_mat, tol, stop_rot, rot, sweeps)                  \
                : launch_jacobi_block<BS, 16>(st, dev, Gt, d, n_mat, tol, stop_rot, rot, sweeps))
    if (bs == 16) return AMB_JB(16);
    if (bs == 8) return AMB_JB(8);
    return AMB_JB(4);
#undef AMB_JB
  }
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
  void* args[] = {&Gt, &d, &n_mat, &tol, &rot, &sweeps};
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
  const size_t dp = static_cast<size_t>(round_up_ll(d, kPolarPad));
  // covariance copies (2), factors (2), M (1) | counters | four padded iteration matrices of the polar path
  return static_cast<size_t>(5) * batch * d * d * 8 + 4096 + static_cast<size_t>(batch) * 16 + 256 +
         static_cast<size_t>(4) * batch * dp * dp * 8;
}

int amb_frechet(int dev, amb_stream_t stream, int batch, int d, const double* mu_x,
                const double* cov_x, const double* mu_y, const double* cov_y, double* out, void* ws,
                size_t ws_bytes) {
  if (!mu_x || !cov_x || !mu_y || !cov_y || !out || batch <= 0 || d <= 0)
    return set_error(AMB_ERR_ARG, "amb_frechet: bad argument");
  if (d > 2048) return set_error(AMB_ERR_ARG, "amb_frechet: d=%d > 2048 not supported", d);
  if (batch > 128) return set_error(AMB_ERR_ARG, "amb_frechet: batch=%d > 128 (split the call)", batch);
  const size_t need = amb_frechet_ws_bytes(batch, d);
  if (!ws || ws_bytes < need) return set_error(AMB_ERR_WS, "amb_frechet: workspace %zu < %zu", ws_bytes, need);
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t mat = static_cast<size_t>(d) * d;
  double* G = static_cast<double*>(ws);          // [2*batch] : covariance copies, x then y
  double* Lt = G + 2 * batch * mat;              // [2*batch] : factors (row j = column j), x then y
  double* Mt = Lt + 2 * batch * mat;             // [batch]
  int* counters = reinterpret_cast<int*>(Mt + batch * mat);     // 1024 ints
  int* panel_n = counters + 512;                                // [2*batch] + rank [2*batch]
  double* polar_bufs = reinterpret_cast<double*>(
      static_cast<uint8_t*>(ws) + round_up_ll(static_cast<long long>(5 * batch * mat * 8 + 4096 + batch * 16), 256));
  int rc;
  if ((rc = check_cuda(cudaMemsetAsync(counters, 0, 4096, st), "memset"))) return rc;
  if ((rc = check_cuda(cudaMemcpyAsync(G, cov_x, batch * mat * 8, cudaMemcpyDeviceToDevice, st), "memcpy"))) return rc;
  if ((rc = check_cuda(cudaMemcpyAsync(G + batch * mat, cov_y, batch * mat * 8, cudaMemcpyDeviceToDevice, st), "memcpy"))) return rc;
  double* F = Lt;
  if (option(kOptFadFactorEig)) {   // eigen-factors by Jacobi (the slower first implementation)
    if ((rc = launch_jacobi(st, dev, G, d, 2 * batch, counters, false))) return rc;
    factor_scale_kernel<<<(2 * batch * d * 32 + 255) / 256, 256, 0, st>>>(G, d, 2 * batch);
    if ((rc = check_launch("factor_scale_kernel"))) return rc;
    F = G;
  } else {
    if ((rc = check_cuda(cudaMemsetAsync(Lt, 0, 2 * batch * mat * 8, st), "memset"))) return rc;
    const int PB = d <= 512 ? 32 : (d <= 1024 ? 16 : 8);
    const size_t smem = (static_cast<size_t>(PB) * d + d) * sizeof(double);
    if ((rc = check_cuda(cudaFuncSetAttribute(pchol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(smem)), "cudaFuncSetAttribute(pchol)"))) return rc;
    const int sms = sm_count(dev);
    // matrices per cooperative launch: all CTAs must be co-resident (one per SM at this smem size)
    for (int m0 = 0; m0 < 2 * batch; m0 += sms) {
      int n_mat = 2 * batch - m0 < sms ? 2 * batch - m0 : sms;
      int C = sms / n_mat;
      if (C > 16) C = 16;
      if (option(kOptFadCtas) > 0 && C * n_mat > option(kOptFadCtas)) C = option(kOptFadCtas) / n_mat > 0 ? option(kOptFadCtas) / n_mat : 1;
      double* Ap = G + static_cast<size_t>(m0) * mat;
      double* Lp = Lt + static_cast<size_t>(m0) * mat;
      int* pn = panel_n + m0;
      int* rk = panel_n + 2 * batch + m0;
      int dd = d, pb = PB;
      void* args[] = {&Ap, &Lp, &dd, &n_mat, &C, &pb, &pn, &rk};
      rc = check_cuda(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(pchol_kernel), dim3(n_mat * C), dim3(512),
                                                  args, smem, st), "cudaLaunchCooperativeKernel(pchol_kernel)");
      if (rc) return rc;
      if ((rc = check_launch("pchol_kernel"))) return rc;
    }
  }
  dim3 ggrid((d + 63) / 64, (d + 63) / 64, batch);
  dgemm_nt_kernel<<<ggrid, 256, 0, st>>>(F, F + batch * mat, Mt, d);   // Mt[i][j] = <F_x col i, F_y col j>
  if ((rc = check_launch("dgemm_nt_kernel"))) return rc;
  if (option(kOptFadMethod) == 1) {   // one-sided Jacobi: singular values as column norms (cross-check path)
    if ((rc = launch_jacobi(st, dev, Mt, d, batch, counters + 64, true))) return rc;
    fad_combine_kernel<<<batch, 256, 0, st>>>(d, mu_x, cov_x, mu_y, cov_y, Mt, out);
    if ((rc = check_launch("fad_combine_kernel"))) return rc;
  } else {                            // polar iteration: c = <U, M>
    const double* U = nullptr;
    if ((rc = launch_polar(st, Mt, d, batch, polar_bufs, &U))) return rc;
    fad_combine_polar_kernel<<<batch, 1024, 0, st>>>(d, static_cast<int>(round_up_ll(d, kPolarPad)), mu_x, cov_x, mu_y,
                                                     cov_y, Mt, U, out);
    if ((rc = check_launch("fad_combine_polar_kernel"))) return rc;
  }
  if (option(kOptFadDebug)) {   // ranks, sweeps / rotations per sweep of the Jacobi stages (synchronises)
    int h[1024] = {0};
    if (cudaStreamSynchronize(st) == cudaSuccess &&
        cudaMemcpy(h, counters, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
      fprintf(stderr, "[amb] factor ranks:");
      for (int i = 0; i < 2 * batch && i < 16; ++i) fprintf(stderr, " %d", h[512 + 2 * batch + i]);
      fprintf(stderr, "\n");
      for (int stage = 0; stage < 2; ++stage) {
        const int* c = h + stage * 64;
        fprintf(stderr, "[amb] jacobi stage %d: %d sweeps, rotations:", stage ? 3 : 1, c[kJacobiMaxSweeps]);
        for (int i = 0; i < c[kJacobiMaxSweeps] && i < kJacobiMaxSweeps; ++i) fprintf(stderr, " %d", c[i]);
        fprintf(stderr, "\n");
      }
    }
  }
  return AMB_OK;
}

}  // extern "C"

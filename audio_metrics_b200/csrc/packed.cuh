// Packed operand format shared by the pack kernel and the tcgen05 pair engine.
//
// An embedding matrix X [n, d] (fp32 or fp64, row-major) is rewritten once per
// set into two fp16 planes  hi, lo  with  x * 2^e_row  ~=  hi + lo  (22+ bits of
// the scaled value; e_row is a power of two SHARED BY THE 256 ROWS OF A TILE that puts
// the tile's largest magnitude in [2^14, 2^15) so that neither plane overflows fp16;
// fp16's 30 binades keep rows far below the tile maximum at full precision).
// The planes are stored in HBM already in the shared-memory image the tensor
// core reads (K-major, no swizzle, 8-row x 16-byte core matrices), so that one
// pipeline stage is a handful of contiguous 8 KiB bulk copies and needs no
// tensor map:
//
//   plane[p]  (p = 0 hi, 1 lo)           plane stride = rows_pad * kpad halfs
//     chunk (rb, kb)  : 128 rows x 32 k   8 KiB, at ((rb * KB) + kb) * 4096 halfs
//       [n = row/8 (16)][c = k/8 (4)][r = row%8 (8)][e = k%8 (8)]  halfs
//
// so inside a chunk the descriptor strides are LBO = 128 B (next 8 k) and
// SBO = 512 B (next 8 rows), and two consecutive row blocks of the same kb...
// are NOT adjacent (kb varies fastest), which is why the B operand of a 256-wide
// tile is fetched as two chunk copies placed back to back in shared memory.
//
// rows_pad is a multiple of 256 (one B tile), kpad a multiple of 32; padding is
// zero.  Side arrays, indexed by packed row:  inv_scale[] = 2^-e_row (float),
// norm[] = |hi+lo|^2 * inv_scale^2 (float; +inf marks a padding row).
#pragma once
#include <cstdint>
#include <cuda_fp16.h>

namespace amb {

constexpr int kBlockRows = 128;   // rows per chunk (= MMA M)
constexpr int kBlockK = 32;       // k elements per chunk
constexpr int kChunkHalfs = kBlockRows * kBlockK;      // 4096
constexpr int kChunkBytes = kChunkHalfs * 2;           // 8192
constexpr int kRowPad = 256;      // rows_pad granularity (= tile N)

struct PackedView {
  const __half* planes;     // hi plane; lo plane at + plane_halfs
  long long plane_halfs;    // rows_pad * kpad
  const float* inv_scale;   // [rows_pad]
  const float* norm;        // [rows_pad]
  int rows_pad;
  int kb_count;             // kpad / 32
};

__host__ __device__ inline long long round_up_ll(long long v, long long m) {
  return (v + m - 1) / m * m;
}

}  // namespace amb

// The all-pairs engine: a persistent, warp-specialised tcgen05 kernel that forms
// 128 x 256 tiles of  X * Y^T  in tensor memory and hands every finished tile to
// an epilogue functor while the next tile is being multiplied.  No N x M matrix
// is ever written: the functors reduce tiles to kernel sums (KD), per-row
// (k+1)-smallest lists (radii) or neighbourhood counts (PRDC).
//
// Precision: each operand is the fp16 pair (hi, lo) of packed.cuh and every
// k-step issues three MMAs  hi*hi + hi*lo + lo*hi  into one fp32 accumulator,
// i.e. a dot product of the 22-bit operands (the dropped lo*lo term is 2^-22
// relative).  This is what makes the tensor-core distances usable for exact
// neighbourhood counts and for the 1e-4 KD tolerance; a single fp16 pass is not.
//
// CTA = 320 threads:  warp 0 bulk-copy producer | warp 1 MMA issuer (+ TMEM
// owner) | warps 2..9 epilogue.  A warp can only read the TMEM lane quarter
// 32*(warp%4), so two warps share each quarter and split the tile's columns:
// thread = (tile row, column half).  Two epilogue warps per SM sub-partition
// hide each other's TMEM-load / shared-memory latencies.
// Pipelines: smem ring full/empty (producer <-> MMA), TMEM double buffer
// full/empty (MMA <-> epilogue).  Work item = (problem, row tile, column split);
// a CTA walks items blockIdx.x, +gridDim.x, ... and, inside an item, all column
// tiles of its split, so per-row state lives in registers across the sweep.
#pragma once
#include "packed.cuh"
#include "tc05.cuh"

namespace amb {

constexpr int kTileM = 128;
constexpr int kTileN = 256;
constexpr int kStages = 4;
constexpr int kStageBytes = 6 * kChunkBytes;          // A hi,lo + B hi(2),lo(2) = 48 KiB
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kEngineThreads = 64 + kEpiThreads;
constexpr int kScratchFloats = 9;                     // 8 parked keys, odd stride: conflict-free
constexpr uint32_t kTmemCols = 512;                    // two 256-column accumulators

struct EngineGeom {
  // operands
  const __half* a_planes;
  const __half* b_planes;
  long long a_plane_halfs;
  long long b_plane_halfs;
  int kb_count;              // k blocks of 32
  // problems: problem p uses A row blocks [a_rb0[p], +n_rt) and B row blocks
  // [b_rb0[p], +2*n_ct); nullptr tables mean a single problem starting at 0.
  const int* a_rb0;
  const int* b_rb0;
  int a_rb_base;             // added to every A row block index (row shards)
  int b_rb_base;
  int n_problems;
  int n_rt;                  // row tiles (128) per problem
  int n_ct;                  // column tiles (256) per problem
  int n_split;               // column splits per (problem, row tile)
  // debug knobs (descriptor strides), normally 128 / 512
  uint32_t lbo_bytes;
  uint32_t sbo_bytes;
  int n_stages;              // single-pass kernel: B ring depth chosen by the host (<= kMaxStages)
  int* work_counter;         // CTA-pair kernel: zeroed device counter for dynamic item hand-out (nullptr: static)
};

constexpr int kMaxStages = 8;
template <int NCV>
struct EngineSmemT {
  alignas(128) float colvec[2][NCV][kTileN];   // bulk-copy destinations: keep 16 B aligned
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t cv_full[2];
  uint64_t a_full;      // single-pass kernel: resident A panel landed
  uint64_t a_empty;     //                     ... and is no longer read by any MMA
  uint32_t tmem_base;
  uint32_t pad_;
};

// Epilogues that need it (Epi::kScratch) get kScratchFloats floats of shared memory per
// epilogue thread, placed after EngineSmem, followed by two floats per tile row through
// which the two threads of a row (column halves) can exchange a value.
constexpr size_t kScratchBytes = size_t(kEpiThreads) * kScratchFloats * sizeof(float) + size_t(kTileM) * 2 * sizeof(float);
template <class Epi>
using EngineSmemOf = EngineSmemT<Epi::kColVecs>;
template <class Epi>
constexpr size_t engine_smem_bytes() {
  return size_t(kStages) * kStageBytes + sizeof(EngineSmemOf<Epi>) + (Epi::kScratch ? kScratchBytes : 0);
}

struct ItemCoord {
  int problem, rt, split, ct_begin, ct_end;
};

// Items are numbered split-major: all (problem, row tile) pairs of column split 0
// first, then split 1, ...  The persistent CTAs deal items round-robin, so at any
// moment the whole grid sweeps the SAME column split: that split's B operand
// (sized by the host to fit L2 next to the in-flight A panels) is fetched from
// HBM once instead of once per row tile.
__device__ __forceinline__ ItemCoord decode_item(const EngineGeom& g, int item) {
  ItemCoord c;
  const int per_split = g.n_problems * g.n_rt;
  c.split = item / per_split;
  const int t = item - c.split * per_split;
  c.rt = t % g.n_rt;
  c.problem = t / g.n_rt;
  c.ct_begin = static_cast<int>(static_cast<long long>(g.n_ct) * c.split / g.n_split);
  c.ct_end = static_cast<int>(static_cast<long long>(g.n_ct) * (c.split + 1) / g.n_split);
  return c;
}

// Epilogue concept:
//   struct Epi {
//     static constexpr int kColVecs;                 // column vectors staged per tile
//     static constexpr bool kScratch;                // wants the per-row scratch area
//     struct Row;                                    // per-thread (= per tile row) state
//     const float* colvec_ptr(int v) const;          // global array, indexed by packed B row
//     void row_begin(Row&, const ItemCoord&, long long a_row /*packed A row*/, int half,
//                    float* xchg /*2 floats shared by the two threads of this tile row, or nullptr;
//                                  both threads pass a 64-thread barrier after row_begin*/) const;
//     const float* cmin_ptr() const;                 // per-32-row chunk minima of |y|^2 (packed.cuh), or nullptr
//     const float* cmax_ptr() const;                 // a second per-32-row chunk array (maxima of a threshold), or nullptr
//     void chunk(Row&, const uint32_t (&acc)[32], const float (*cv)[kTileN], int col_in_tile,
//                int col_in_problem, long long b_row0 /*packed B row of chunk column 0*/,
//                float* scratch /*kScratchFloats floats private to this thread, or nullptr*/,
//                float cmin /*min |y|^2 over the chunk's 32 columns if the kernel stages it, else -inf*/,
//                float cmax /*the chunk's value of cmax_ptr() if staged, else +inf*/) const;
//     void row_end(Row&, const ItemCoord&, int item, long long a_row, int quarter, int lane,
//                  int half /*which 128 columns of every tile this thread swept*/) const;
//   };
// A thread sees columns [128*half, 128*half + 128) of every column tile of the item.

// Epilogue role, shared by both kernels: warps 2..9.
template <class Epi>
__device__ __forceinline__ void epilogue_role(const EngineGeom& g, const Epi& epi, EngineSmemOf<Epi>* sh,
                                              uint32_t tmem_base, int warp, int lane) {
  using EngineSmem = EngineSmemOf<Epi>;
  float* scratch = Epi::kScratch
                       ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sh) + sizeof(EngineSmem)) +
                             (threadIdx.x - 64) * kScratchFloats
                       : nullptr;
  const int n_items = g.n_problems * g.n_rt * g.n_split;
  const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32)
  const int half = (warp - 2) >> 2;             // columns [128*half, +128) of each tile
  const int row_in_tile = quarter * 32 + lane;
  float* xchg = Epi::kScratch ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sh) + sizeof(EngineSmem)) +
                                    kEpiThreads * kScratchFloats + row_in_tile * 2
                              : nullptr;
  int acc = 0;
  uint32_t acc_ph = 0;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const ItemCoord c = decode_item(g, item);
    const long long a_row = ((g.a_rb0 ? g.a_rb0[c.problem] : 0) + g.a_rb_base + c.rt) * static_cast<long long>(kTileM) + row_in_tile;
    const long long b_row_base = static_cast<long long>((g.b_rb0 ? g.b_rb0[c.problem] : 0) + g.b_rb_base) * kBlockRows;
    typename Epi::Row row;
    epi.row_begin(row, c, a_row, half, xchg);
    if (Epi::kScratch) named_bar_sync(1 + quarter, 64);   // both column halves of the quarter are in this item
    for (int ct = c.ct_begin; ct < c.ct_end; ++ct) {
      const long long b_row0 = b_row_base + static_cast<long long>(ct) * kTileN;
      mbar_wait(&sh->cv_full[acc], acc_ph);     // column vectors landed (producer bulk copy)
      mbar_wait(&sh->tmem_full[acc], acc_ph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + static_cast<uint32_t>(acc) * kTileN +
                              (static_cast<uint32_t>(quarter * 32) << 16);
      epi.tile_begin(row, sh->colvec[acc]);
#pragma unroll 1
      for (int c0 = half * (kTileN / 2); c0 < (half + 1) * (kTileN / 2); c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_addr + c0, r);
        tmem_wait_ld();
        epi.template chunk<false>(row, r, sh->colvec[acc], c0, ct * kTileN + c0, b_row0 + c0, scratch, -__builtin_huge_valf(), __builtin_huge_valf());
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->tmem_empty[acc]);
      epi.tile_end(row, scratch);
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1;
    }
    epi.row_end(row, c, item, a_row, quarter, lane, half);
  }
}

template <class Epi>
__global__ void __launch_bounds__(kEngineThreads, 1)
pair_engine_kernel(const EngineGeom g, const Epi epi) {
  // no pointer laundering here: everything derived from smem_buf stays in the
  // shared address space for the compiler (LDS/STS instead of generic LD/ST)
  using EngineSmem = EngineSmemOf<Epi>;
  extern __shared__ __align__(1024) uint8_t smem_buf[];
  uint8_t* stage_base = smem_buf;
  EngineSmem* sh = reinterpret_cast<EngineSmem*>(smem_buf + size_t(kStages) * kStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_items = g.n_problems * g.n_rt * g.n_split;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sh->tmem_full[a], 1);
      mbar_init(&sh->tmem_empty[a], kEpiWarps);
      mbar_init(&sh->cv_full[a], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&sh->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ producer
    // (whole warp runs the loops; one elected lane issues — see elect_one())
    // the A row panel is re-read for every column tile of the item: ask L2 to keep it
    const uint64_t keep = l2_policy_evict_last();
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const ItemCoord c = decode_item(g, item);
      const long long a_rb = (g.a_rb0 ? g.a_rb0[c.problem] : 0) + g.a_rb_base + c.rt;
      const long long b_rb_base = (g.b_rb0 ? g.b_rb0[c.problem] : 0) + g.b_rb_base;
      const __half* a_src = g.a_planes + a_rb * g.kb_count * kChunkHalfs;
      for (int ct = c.ct_begin; ct < c.ct_end; ++ct) {
        const __half* b_src = g.b_planes + (b_rb_base + 2ll * ct) * g.kb_count * kChunkHalfs;
        const long long b_next = static_cast<long long>(g.kb_count) * kChunkHalfs;
        // this tile's column vectors ride the same engine: 1 KiB bulk copies into
        // the buffer the epilogue released two tiles ago
        mbar_wait(&sh->tmem_empty[acc], acc_ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&sh->cv_full[acc], Epi::kColVecs * kTileN * 4);
#pragma unroll
          for (int v = 0; v < Epi::kColVecs; ++v)
            bulk_g2s(sh->colvec[acc][v], epi.colvec_ptr(v) + (b_rb_base + 2ll * ct) * kBlockRows, kTileN * 4,
                     &sh->cv_full[acc]);
        }
        acc ^= 1;
        if (acc == 0) acc_ph ^= 1;
        for (int kb = 0; kb < g.kb_count; ++kb) {
          mbar_wait(&sh->empty[s], ph ^ 1);
          if (elect_one()) {
            uint8_t* st = stage_base + size_t(s) * kStageBytes;
            mbar_expect_tx(&sh->full[s], kStageBytes);
            const long long ko = static_cast<long long>(kb) * kChunkHalfs;
            bulk_g2s_hint(st + 0 * kChunkBytes, a_src + ko, kChunkBytes, &sh->full[s], keep);
            bulk_g2s_hint(st + 1 * kChunkBytes, a_src + g.a_plane_halfs + ko, kChunkBytes, &sh->full[s], keep);
            bulk_g2s(st + 2 * kChunkBytes, b_src + ko, kChunkBytes, &sh->full[s]);
            bulk_g2s(st + 3 * kChunkBytes, b_src + b_next + ko, kChunkBytes, &sh->full[s]);
            bulk_g2s(st + 4 * kChunkBytes, b_src + g.b_plane_halfs + ko, kChunkBytes, &sh->full[s]);
            bulk_g2s(st + 5 * kChunkBytes, b_src + g.b_plane_halfs + b_next + ko, kChunkBytes, &sh->full[s]);
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(kTileM, kTileN);
    const uint64_t desc0 = make_kmajor_desc(smem_u32(stage_base), g.lbo_bytes, g.sbo_bytes);
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const ItemCoord c = decode_item(g, item);
      for (int ct = c.ct_begin; ct < c.ct_end; ++ct) {
        mbar_wait(&sh->tmem_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * kTileN;
        for (int kb = 0; kb < g.kb_count; ++kb) {
          mbar_wait(&sh->full[s], ph);
          tc_fence_after();
          if (elect_one()) {
            // descriptor address field counts 16-byte units: stage s at +s*kStageBytes/16
            const uint64_t st = desc0 + static_cast<uint64_t>(s * (kStageBytes >> 4));
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t a_hi = st + (0 * kChunkBytes + ks * 256) / 16;   // two 8-wide k chunks of 128 B
              const uint64_t a_lo = st + (1 * kChunkBytes + ks * 256) / 16;
              const uint64_t b_hi = st + (2 * kChunkBytes + ks * 256) / 16;
              const uint64_t b_lo = st + (4 * kChunkBytes + ks * 256) / 16;
              mma_f16_ss(d_tmem, a_hi, b_lo, idesc, (kb | ks) != 0 ? 1u : 0u);
              mma_f16_ss(d_tmem, a_lo, b_hi, idesc, 1u);
              mma_f16_ss(d_tmem, a_hi, b_hi, idesc, 1u);
            }
            tc_commit(&sh->empty[s]);
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) tc_commit(&sh->tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_ph ^= 1;
      }
    }
    __syncwarp();
  } else {
    epilogue_role(g, epi, sh, tmem_base, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ---------------------------------------------------------------------------
// Single-pass variant (radii and counts): one fp16 MMA per product on the hi
// plane only.  The result is only a filter — every decision inside its (wider)
// error band is re-made exactly by the refine kernels — so the tensor cores do a
// third of the work.  With a third of the math per byte the operand traffic would
// out-run L2 (A + B re-read per tile = 94 B/clk/SM), so the A row panel of the
// item (128 rows x kpad fp16 <= 128 KiB) stays RESIDENT in shared memory for the
// whole column sweep and only B streams through an n_stages-deep ring of 16 KiB
// stages.  Requires kb_count <= 16 (d <= 512); wider inputs use the 3-pass kernel.
constexpr int kStage1Bytes = 2 * kChunkBytes;            // B hi, two row blocks
constexpr int kMaxResidentKb = 16;

template <class Epi>
__global__ void __launch_bounds__(kEngineThreads, 1)
pair_engine1_kernel(const EngineGeom g, const Epi epi) {
  using EngineSmem = EngineSmemOf<Epi>;
  extern __shared__ __align__(1024) uint8_t smem_buf[];
  uint8_t* a_panel = smem_buf;                                           // kb_count chunks
  uint8_t* stage_base = smem_buf + size_t(g.kb_count) * kChunkBytes;
  EngineSmem* sh = reinterpret_cast<EngineSmem*>(stage_base + size_t(g.n_stages) * kStage1Bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_items = g.n_problems * g.n_rt * g.n_split;
  const int n_stages = g.n_stages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sh->tmem_full[a], 1);
      mbar_init(&sh->tmem_empty[a], kEpiWarps);
      mbar_init(&sh->cv_full[a], 1);
    }
    mbar_init(&sh->a_full, 1);
    mbar_init(&sh->a_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&sh->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ producer
    int s = 0;
    uint32_t ph = 0, a_ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const ItemCoord c = decode_item(g, item);
      const long long a_rb = (g.a_rb0 ? g.a_rb0[c.problem] : 0) + g.a_rb_base + c.rt;
      const long long b_rb_base = (g.b_rb0 ? g.b_rb0[c.problem] : 0) + g.b_rb_base;
      const __half* a_src = g.a_planes + a_rb * g.kb_count * kChunkHalfs;
      // resident A panel: one contiguous kb_count * 8 KiB run of the hi plane
      mbar_wait(&sh->a_empty, a_ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&sh->a_full, static_cast<uint32_t>(g.kb_count) * kChunkBytes);
        for (int kb = 0; kb < g.kb_count; ++kb)
          bulk_g2s(a_panel + size_t(kb) * kChunkBytes, a_src + static_cast<long long>(kb) * kChunkHalfs, kChunkBytes,
                   &sh->a_full);
      }
      a_ph ^= 1;
      for (int ct = c.ct_begin; ct < c.ct_end; ++ct) {
        const __half* b_src = g.b_planes + (b_rb_base + 2ll * ct) * g.kb_count * kChunkHalfs;
        const long long b_next = static_cast<long long>(g.kb_count) * kChunkHalfs;
        mbar_wait(&sh->tmem_empty[acc], acc_ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&sh->cv_full[acc], Epi::kColVecs * kTileN * 4);
#pragma unroll
          for (int v = 0; v < Epi::kColVecs; ++v)
            bulk_g2s(sh->colvec[acc][v], epi.colvec_ptr(v) + (b_rb_base + 2ll * ct) * kBlockRows, kTileN * 4,
                     &sh->cv_full[acc]);
        }
        acc ^= 1;
        if (acc == 0) acc_ph ^= 1;
        for (int kb = 0; kb < g.kb_count; ++kb) {
          mbar_wait(&sh->empty[s], ph ^ 1);
          if (elect_one()) {
            uint8_t* st = stage_base + size_t(s) * kStage1Bytes;
            mbar_expect_tx(&sh->full[s], kStage1Bytes);
            const long long ko = static_cast<long long>(kb) * kChunkHalfs;
            bulk_g2s(st, b_src + ko, kChunkBytes, &sh->full[s]);
            bulk_g2s(st + kChunkBytes, b_src + b_next + ko, kChunkBytes, &sh->full[s]);
          }
          if (++s == n_stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc_f16(kTileM, kTileN);
    const uint64_t a_desc0 = make_kmajor_desc(smem_u32(a_panel), g.lbo_bytes, g.sbo_bytes);
    const uint64_t b_desc0 = make_kmajor_desc(smem_u32(stage_base), g.lbo_bytes, g.sbo_bytes);
    int s = 0;
    uint32_t ph = 0, a_ph = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const ItemCoord c = decode_item(g, item);
      mbar_wait(&sh->a_full, a_ph);
      a_ph ^= 1;
      for (int ct = c.ct_begin; ct < c.ct_end; ++ct) {
        mbar_wait(&sh->tmem_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * kTileN;
        for (int kb = 0; kb < g.kb_count; ++kb) {
          mbar_wait(&sh->full[s], ph);
          tc_fence_after();
          if (elect_one()) {
            // descriptor address field counts 16-byte units
            const uint64_t a_d = a_desc0 + static_cast<uint64_t>(kb * (kChunkBytes >> 4));
            const uint64_t b_d = b_desc0 + static_cast<uint64_t>(s * (kStage1Bytes >> 4));
            mma_f16_ss(d_tmem, a_d, b_d, idesc, kb != 0 ? 1u : 0u);
            mma_f16_ss(d_tmem, a_d + 16, b_d + 16, idesc, 1u);     // second 16-wide k step: +256 B
            tc_commit(&sh->empty[s]);
          }
          if (++s == n_stages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) tc_commit(&sh->tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_ph ^= 1;
      }
      if (elect_one()) tc_commit(&sh->a_empty);   // all MMAs that read this A panel are done
    }
    __syncwarp();
  } else {
    epilogue_role(g, epi, sh, tmem_base, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace amb

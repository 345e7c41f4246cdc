// Two-CTA variant of the single-pass pair engine (radii and counts, d <= 512).
//
// A cluster of two CTAs (one SM pair) works on 256 A rows x 256 B columns per tile with
// tcgen05.mma.cta_group::2 (M256 N256 K16, issued by the leader CTA only):
//   * each CTA keeps ITS 128-row A panel resident in shared memory (as pair_engine1_kernel),
//   * each CTA streams only ITS half of the B tile (128 of the 256 columns): 8 KiB per k
//     block instead of 16 KiB, so a 16 KiB ring slot holds K = 64 — four MMAs (512 tensor-pipe
//     cycles) per producer / issuer round trip instead of two, the same ring memory covers
//     twice the time, and the L2 -> SM operand traffic per flop is halved,
//   * each CTA's tensor memory receives the accumulator rows of its own A panel against all
//     256 columns, so the epilogue (one thread = tile row x column half) is unchanged.
//
// Hand-shakes (L = leader CTA, rank 0; P = peer CTA, rank 1):
//   full[s]            local   TMA bytes of this CTA's slot s landed; in L the same barrier also counts
//                              the remote arrive of P's relay warp ("P.full[s] completed"), so the
//                              MMA issuer makes one wait per slot for both halves of the B tile
//   empty[s]           local   L's MMA commit, multicast to both CTAs: slot s may be refilled
//   a_full/peer_a_full, a_empty   the same three for the resident A panel
//   tmem_full[acc]     local   L's MMA commit, multicast: accumulator acc is complete
//   cv_full/cv_empty[q] local  the per-tile column vectors live in their own four-slot ring, so the
//                              producer never waits for the epilogue before prefetching the next
//                              tile's B slots (with the vectors in the accumulator's double buffer
//                              it did, and every tile started on a cold ring)
//   pair_tmem_empty[acc] in L  all 16 epilogue warps of the pair drained acc (gates L's next MMAs)
//
// Work items are handed out DYNAMICALLY (g.work_counter != nullptr): the leader's producer warp
// takes the next item with one atomicAdd, publishes it in a two-slot mailbox in both CTAs
// (sched_item / sched_full) and every other warp of the pair reads it there and acknowledges on
// the leader's sched_empty.  A CTA pair that gets its SMs late (another kernel — the N-independent
// FAD kernels on a second stream — was holding them) simply takes fewer items, and the last wave
// balances itself.
#pragma once
#include "pair_engine.cuh"

namespace amb {

constexpr int kStage2Bytes = 2 * kChunkBytes;            // this CTA's B half, two k blocks (K = 64)
constexpr int kCvSlots = 4;
constexpr int kSchedSlots = 2;
constexpr int kSchedConsumers = 2 * (kEngineThreads / 32) - 1;   // every warp of the pair but the scheduler

template <int NCV>
struct EngineSmem2T {
  alignas(128) float colvec[kCvSlots][NCV][kTileN];
  alignas(32) float cvmin[kCvSlots][kTileN / 32];   // min |y|^2 of each 32-column chunk of the tile
  alignas(32) float cvmax[kCvSlots][kTileN / 32];   // the epilogue's second per-chunk array
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t pair_tmem_empty[2];
  uint64_t cv_full[kCvSlots];
  uint64_t cv_empty[kCvSlots];
  uint64_t a_full;
  uint64_t peer_a_full;
  uint64_t a_empty;
  uint64_t sched_full[kSchedSlots];    // mailbox slot written (both CTAs)
  uint64_t sched_empty[kSchedSlots];   // in L: all other warps of the pair have read the slot
  int sched_item[kSchedSlots];
  uint32_t tmem_base;
  uint32_t pad_;
};
template <class Epi>
using EngineSmem2Of = EngineSmem2T<Epi::kColVecs>;

// ----------------------------------------------------------------- cluster PTX
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA's layout) in CTA `rank`
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// Default (.release.cta) semantics on purpose, as CUTLASS's ClusterBarrier::arrive: what these
// barriers order is shared memory filled by the bulk-copy engine and read by the tensor core,
// not generic-proxy data, and an explicit .release.cluster compiles to MEMBAR.ALL.GPU + error
// barriers on every arrive (measured: it serialised the whole ring, 1.7 k cycles per slot).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a local barrier whose arrivals come from the other CTA: the plain .acquire.cta wait
// (an .acquire.cluster wait is followed by CCTL.IVALL, an L1 invalidate, on every probe)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
// Cluster-scope release / acquire pair for the item mailbox (generic-proxy data crosses CTAs
// here, once per work item, so the full fences are the right tool and their cost is irrelevant).
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins == (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void st_cluster_s32(uint32_t cluster_addr, int v) {
  asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}

// Next work item of this CTA pair.  `it` counts the items this warp has asked for.
//   scheduler = warp 0 of the leader CTA; every other warp of the pair is a consumer.
template <class SM>
__device__ __forceinline__ int next_item(SM* sh, const EngineGeom& g, int& it, int cluster_id, int n_clusters,
                                         bool scheduler, int lane) {
  if (g.work_counter == nullptr) return cluster_id + (it++) * n_clusters;   // static round robin
  const int slot = it & (kSchedSlots - 1);
  const uint32_t par = static_cast<uint32_t>(it / kSchedSlots) & 1u;
  ++it;
  int item = 0;
  if (scheduler) {
    mbar_wait(&sh->sched_empty[slot], par ^ 1u);
    if (lane == 0) {
      item = atomicAdd(g.work_counter, 1);
      sh->sched_item[slot] = item;
      st_cluster_s32(map_to_cta(&sh->sched_item[slot], 1), item);
      mbar_arrive(&sh->sched_full[slot]);
      mbar_arrive_release_cluster(map_to_cta(&sh->sched_full[slot], 1));
    }
    item = __shfl_sync(0xffffffffu, item, 0);
  } else {
    mbar_wait_acquire_cluster(&sh->sched_full[slot], par);
    item = *reinterpret_cast<volatile int*>(&sh->sched_item[slot]);
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(map_to_cta(&sh->sched_empty[slot], 0));
  }
  return item;
}

__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// completion of all MMAs issued so far by this thread arrives (once) on `bar` in BOTH CTAs
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void mma2_f16_ss(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Geometry: as EngineGeom, with n_rt counting row-tile PAIRS (256 A rows); CTA `rank` of the
// cluster owns A row tile 2*pair + rank and B row block 2*ct + rank of every column tile.
template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kEngineThreads, 1)
pair_engine2_kernel(const EngineGeom g, const Epi epi) {
  using EngineSmem = EngineSmem2Of<Epi>;
  extern __shared__ __align__(1024) uint8_t smem_buf[];
  uint8_t* a_panel = smem_buf;                                           // kb_count chunks
  uint8_t* stage_base = smem_buf + size_t(g.kb_count) * kChunkBytes;
  EngineSmem* sh = reinterpret_cast<EngineSmem*>(stage_base + size_t(g.n_stages) * kStage2Bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int n_items = g.n_problems * g.n_rt * g.n_split;
  const int n_stages = g.n_stages;
  const int kb_count = g.kb_count;
  const int n_kp = (kb_count + 1) >> 1;          // K = 64 steps per column tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&sh->full[s], leader ? 2 : 1);   // own producer (+ bytes), and in L the peer's relay
      mbar_init(&sh->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sh->tmem_full[a], 1);
      mbar_init(&sh->pair_tmem_empty[a], 2 * kEpiWarps);
    }
    for (int q = 0; q < kCvSlots; ++q) {
      mbar_init(&sh->cv_full[q], 1);
      mbar_init(&sh->cv_empty[q], kEpiWarps);
    }
    mbar_init(&sh->a_full, 1);
    mbar_init(&sh->peer_a_full, 1);
    mbar_init(&sh->a_empty, 1);
    for (int q = 0; q < kSchedSlots; ++q) {
      mbar_init(&sh->sched_full[q], 1);
      mbar_init(&sh->sched_empty[q], kSchedConsumers);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(&sh->tmem_base, kTmemCols);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ producer (both CTAs)
    int s = 0;
    uint32_t ph = 0, a_ph = 0;
    uint32_t tile = 0;         // running column-tile count: column-vector slot tile % kCvSlots
    int it = 0;
    for (;;) {
      const int item = next_item(sh, g, it, cluster_id, n_clusters, leader, lane);
      if (item >= n_items) break;
      const ItemCoord c = decode_item(g, item);
      const long long a_rb = (g.a_rb0 ? g.a_rb0[c.problem] : 0) + g.a_rb_base + 2ll * c.rt + rank;
      const long long b_rb_base = (g.b_rb0 ? g.b_rb0[c.problem] : 0) + g.b_rb_base;
      const __half* a_src = g.a_planes + a_rb * kb_count * kChunkHalfs;
      mbar_wait(&sh->a_empty, a_ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&sh->a_full, static_cast<uint32_t>(kb_count) * kChunkBytes);
        for (int kb = 0; kb < kb_count; ++kb)
          bulk_g2s(a_panel + size_t(kb) * kChunkBytes, a_src + static_cast<long long>(kb) * kChunkHalfs, kChunkBytes,
                   &sh->a_full);
      }
      a_ph ^= 1;
      for (int ct = c.ct_begin; ct < c.ct_end; ++ct) {
        // this CTA's half of the B tile: row block 2 ct + rank, all k blocks contiguous
        const __half* b_src = g.b_planes + (b_rb_base + 2ll * ct + rank) * kb_count * kChunkHalfs;
        const int q = tile & (kCvSlots - 1);
        mbar_wait(&sh->cv_empty[q], ((tile / kCvSlots) & 1u) ^ 1u);   // the epilogue of four tiles ago
        if (elect_one()) {
          const float* cm = epi.cmin_ptr();
          const float* cx = epi.cmax_ptr();
          mbar_expect_tx(&sh->cv_full[q], Epi::kColVecs * kTileN * 4 + (cm ? kTileN / 32 * 4 : 0) + (cx ? kTileN / 32 * 4 : 0));
#pragma unroll
          for (int v = 0; v < Epi::kColVecs; ++v)
            bulk_g2s(sh->colvec[q][v], epi.colvec_ptr(v) + (b_rb_base + 2ll * ct) * kBlockRows, kTileN * 4,
                     &sh->cv_full[q]);
          if (cm)
            bulk_g2s(sh->cvmin[q], cm + (b_rb_base + 2ll * ct) * kBlockRows / 32, kTileN / 32 * 4, &sh->cv_full[q]);
          if (cx)
            bulk_g2s(sh->cvmax[q], cx + (b_rb_base + 2ll * ct) * kBlockRows / 32, kTileN / 32 * 4, &sh->cv_full[q]);
        }
        ++tile;
        for (int j = 0; j < n_kp; ++j) {
          mbar_wait(&sh->empty[s], ph ^ 1);
          if (elect_one()) {
            const uint32_t bytes = (kb_count - 2 * j >= 2) ? 2u * kChunkBytes : 1u * kChunkBytes;
            mbar_expect_tx(&sh->full[s], bytes);
            bulk_g2s(stage_base + size_t(s) * kStage2Bytes, b_src + static_cast<long long>(2 * j) * kChunkHalfs, bytes,
                     &sh->full[s]);
          }
          if (++s == n_stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // -------------------------------------------------------- MMA issuer (leader CTA)
      constexpr uint32_t idesc = make_idesc_f16(2 * kTileM, kTileN);
      const uint64_t a_desc0 = make_kmajor_desc(smem_u32(a_panel), g.lbo_bytes, g.sbo_bytes);
      const uint64_t b_desc0 = make_kmajor_desc(smem_u32(stage_base), g.lbo_bytes, g.sbo_bytes);
      int s = 0;
      uint32_t ph = 0, a_ph = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      int it = 0;
      for (;;) {
        const int item = next_item(sh, g, it, cluster_id, n_clusters, false, lane);
        if (item >= n_items) break;
        const ItemCoord c = decode_item(g, item);
        mbar_wait(&sh->a_full, a_ph);
        mbar_wait_cluster(&sh->peer_a_full, a_ph);
        a_ph ^= 1;
        for (int ct = c.ct_begin; ct < c.ct_end; ++ct) {
          mbar_wait_cluster(&sh->pair_tmem_empty[acc], acc_ph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * kTileN;
          for (int j = 0; j < n_kp; ++j) {
            mbar_wait(&sh->full[s], ph);        // both CTAs' halves of the slot have landed
            tc_fence_after();
            if (elect_one()) {
              const uint64_t a_d = a_desc0 + static_cast<uint64_t>(2 * j * (kChunkBytes >> 4));
              const uint64_t b_d = b_desc0 + static_cast<uint64_t>(s * (kStage2Bytes >> 4));
              const int chunks = (kb_count - 2 * j >= 2) ? 2 : 1;
              mma2_f16_ss(d_tmem, a_d, b_d, idesc, j != 0 ? 1u : 0u);
              mma2_f16_ss(d_tmem, a_d + 16, b_d + 16, idesc, 1u);                 // +256 B: second 16-wide k step
              if (chunks == 2) {
                mma2_f16_ss(d_tmem, a_d + (kChunkBytes >> 4), b_d + (kChunkBytes >> 4), idesc, 1u);
                mma2_f16_ss(d_tmem, a_d + (kChunkBytes >> 4) + 16, b_d + (kChunkBytes >> 4) + 16, idesc, 1u);
              }
              tc_commit2(&sh->empty[s]);
            }
            if (++s == n_stages) { s = 0; ph ^= 1; }
          }
          if (elect_one()) tc_commit2(&sh->tmem_full[acc]);
          acc ^= 1;
          if (acc == 0) acc_ph ^= 1;
        }
        if (elect_one()) tc_commit2(&sh->a_empty);
      }
      __syncwarp();
    } else {
      // -------------------------------------------------------- relay (peer CTA): tell the leader what landed here
      const uint32_t l_a_full = map_to_cta(&sh->peer_a_full, 0);
      int s = 0;
      uint32_t ph = 0, a_ph = 0;
      int it = 0;
      for (;;) {
        const int item = next_item(sh, g, it, cluster_id, n_clusters, false, lane);
        if (item >= n_items) break;
        const ItemCoord c = decode_item(g, item);
        mbar_wait(&sh->a_full, a_ph);
        a_ph ^= 1;
        if (elect_one()) mbar_arrive_cluster(l_a_full);
        for (int ct = c.ct_begin; ct < c.ct_end; ++ct) {
          for (int j = 0; j < n_kp; ++j) {
            mbar_wait(&sh->full[s], ph);
            if (elect_one()) mbar_arrive_cluster(map_to_cta(&sh->full[s], 0));
            if (++s == n_stages) { s = 0; ph ^= 1; }
          }
        }
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------- epilogue (both CTAs)
    float* scratch = Epi::kScratch
                         ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sh) + sizeof(EngineSmem)) +
                               (threadIdx.x - 64) * kScratchFloats
                         : nullptr;
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    float* xchg = Epi::kScratch ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sh) + sizeof(EngineSmem)) +
                                      kEpiThreads * kScratchFloats + row_in_tile * 2
                                : nullptr;
    const uint32_t l_pair_empty0 = map_to_cta(&sh->pair_tmem_empty[0], 0);
    const uint32_t l_pair_empty1 = map_to_cta(&sh->pair_tmem_empty[1], 0);
    int acc = 0;
    uint32_t acc_ph = 0;
    uint32_t tile = 0;
    int it = 0;
    for (;;) {
      const int item = next_item(sh, g, it, cluster_id, n_clusters, false, lane);
      if (item >= n_items) break;
      const ItemCoord c = decode_item(g, item);
      const long long a_row = ((g.a_rb0 ? g.a_rb0[c.problem] : 0) + g.a_rb_base + 2ll * c.rt + rank) *
                                  static_cast<long long>(kTileM) + row_in_tile;
      const long long b_row_base = static_cast<long long>((g.b_rb0 ? g.b_rb0[c.problem] : 0) + g.b_rb_base) * kBlockRows;
      typename Epi::Row row;
      epi.row_begin(row, c, a_row, half, xchg);
      if (Epi::kScratch) named_bar_sync(1 + quarter, 64);
      for (int ct = c.ct_begin; ct < c.ct_end; ++ct) {
        const long long b_row0 = b_row_base + static_cast<long long>(ct) * kTileN;
        const int q = tile & (kCvSlots - 1);
        mbar_wait(&sh->cv_full[q], (tile / kCvSlots) & 1u);
        mbar_wait_cluster(&sh->tmem_full[acc], acc_ph);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + static_cast<uint32_t>(acc) * kTileN +
                                (static_cast<uint32_t>(quarter * 32) << 16);
        epi.tile_begin(row, sh->colvec[q]);
#pragma unroll 1
        for (int c0 = half * (kTileN / 2); c0 < (half + 1) * (kTileN / 2); c0 += 32) {
          uint32_t r[32];
          tmem_ld32(t_addr + c0, r);
          // the chunk's staged scalars: loaded while the TMEM load is in flight
          const float cmin = Epi::kChunkMin ? sh->cvmin[q][c0 >> 5] : -__builtin_huge_valf();
          const float cmax = Epi::kChunkMin ? sh->cvmax[q][c0 >> 5] : __builtin_huge_valf();
          tmem_wait_ld();
          epi.template chunk<Epi::kChunkMin>(row, r, sh->colvec[q], c0, ct * kTileN + c0, b_row0 + c0, scratch, cmin, cmax);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&sh->cv_empty[q]);
          mbar_arrive_cluster(acc == 0 ? l_pair_empty0 : l_pair_empty1);
        }
        epi.tile_end(row, scratch);   // work that does not need the accumulator: after it is handed back
        ++tile;
        acc ^= 1;
        if (acc == 0) acc_ph ^= 1;
      }
      epi.row_end(row, c, item, a_row, quarter, lane, half);
    }
  }

  // nobody leaves (or frees tensor memory) while the other CTA may still signal or read here
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

}  // namespace amb

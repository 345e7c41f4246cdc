// Host-side launcher for the pair engine; include in the .cu that instantiates
// a given epilogue.
#pragma once
#include <cstdlib>
#include <mutex>

#include "epilogues.cuh"
#include "pair_engine2.cuh"
#include "internal.cuh"

namespace amb {

template <class Epi>
int launch_engine(cudaStream_t stream, int dev, const EngineGeom& g, const Epi& epi, const char* what,
                  double alg_pairs = 0.0) {
  static std::once_flag once;  // per Epi instantiation
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(pair_engine_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(engine_smem_bytes<Epi>()));
  });
  // the attribute is per device; set it again cheaply when several devices are in use
  if (attr_err == cudaSuccess)
    attr_err = cudaFuncSetAttribute(pair_engine_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(engine_smem_bytes<Epi>()));
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "cudaFuncSetAttribute(pair_engine)");
  const long long items = static_cast<long long>(g.n_problems) * g.n_rt * g.n_split;
  if (items <= 0) return AMB_OK;
  const int sms = sm_count(dev);
  const unsigned grid = static_cast<unsigned>(items < sms ? items : sms);
  // executed MMA work: full tiles, three fp16 MMAs per k step
  const double exec_flops = static_cast<double>(g.n_problems) * g.n_rt * g.n_ct * (2.0 * kTileM * kTileN) *
                            (g.kb_count * static_cast<double>(kBlockK)) * 3.0;
  void* tok = profile_begin(stream);
  pair_engine_kernel<Epi><<<grid, kEngineThreads, engine_smem_bytes<Epi>(), stream>>>(g, epi);
  profile_end(tok, stream, alg_pairs, exec_flops);
  return check_launch(what);
}

// Single-pass variant (pair_engine1_kernel): resident A panel + B ring sized to what is
// left of the 227 KiB of shared memory.  Only for kb_count <= kMaxResidentKb.
constexpr size_t kMaxDynSmem = 232448;
template <class Epi>
int launch_engine1(cudaStream_t stream, int dev, EngineGeom g, const Epi& epi, const char* what,
                   double alg_pairs = 0.0) {
  if (g.kb_count > kMaxResidentKb) return set_error(AMB_ERR_ARG, "%s: kb_count %d too large for the resident panel", what, g.kb_count);
  const size_t fixed = size_t(g.kb_count) * kChunkBytes + sizeof(EngineSmemOf<Epi>) + (Epi::kScratch ? kScratchBytes : 0);
  int n_stages = static_cast<int>((kMaxDynSmem - fixed) / kStage1Bytes);
  if (n_stages > kMaxStages) n_stages = kMaxStages;
  if (const int v = option(kOptStages)) {   // tuning knob: shallower B ring
    if (v >= 2 && v < n_stages) n_stages = v;
  }
  g.n_stages = n_stages;
  const size_t smem = fixed + size_t(n_stages) * kStage1Bytes;
  cudaError_t attr_err = cudaFuncSetAttribute(pair_engine1_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(kMaxDynSmem));
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "cudaFuncSetAttribute(pair_engine1)");
  const long long items = static_cast<long long>(g.n_problems) * g.n_rt * g.n_split;
  if (items <= 0) return AMB_OK;
  int sms = sm_count(dev);
  if (const int v = option(kOptGrid)) {   // experiment knob: fewer persistent CTAs than SMs
    if (v >= 1 && v < sms) sms = v;
  }
  const unsigned grid = static_cast<unsigned>(items < sms ? items : sms);
  const double exec_flops = static_cast<double>(g.n_problems) * g.n_rt * g.n_ct * (2.0 * kTileM * kTileN) *
                            (g.kb_count * static_cast<double>(kBlockK));
  void* tok = profile_begin(stream);
  pair_engine1_kernel<Epi><<<grid, kEngineThreads, smem, stream>>>(g, epi);
  profile_end(tok, stream, alg_pairs, exec_flops);
  return check_launch(what);
}

// Two-CTA variant (pair_engine2_kernel): g.n_rt counts row-tile PAIRS.
template <class Epi>
int launch_engine2(cudaStream_t stream, int dev, EngineGeom g, const Epi& epi, const char* what,
                   double alg_pairs = 0.0) {
  if (g.kb_count > kMaxResidentKb) return set_error(AMB_ERR_ARG, "%s: kb_count %d too large for the resident panel", what, g.kb_count);
  const size_t fixed = size_t(g.kb_count) * kChunkBytes + sizeof(EngineSmem2Of<Epi>) + (Epi::kScratch ? kScratchBytes : 0);
  int n_stages = static_cast<int>((kMaxDynSmem - fixed) / kStage2Bytes);
  if (n_stages > kMaxStages) n_stages = kMaxStages;
  if (const int v = option(kOptStages)) {
    if (v >= 2 && v < n_stages) n_stages = v;
  }
  if (n_stages < 2) return set_error(AMB_ERR_ARG, "%s: shared memory too small for the B ring", what);
  g.n_stages = n_stages;
  const size_t smem = fixed + size_t(n_stages) * kStage2Bytes;
  cudaError_t attr_err = cudaFuncSetAttribute(pair_engine2_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(kMaxDynSmem));
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "cudaFuncSetAttribute(pair_engine2)");
  const long long items = static_cast<long long>(g.n_problems) * g.n_rt * g.n_split;
  if (items <= 0) return AMB_OK;
  int pairs = (sm_count(dev) - option(kOptReserveSms)) / 2;   // SMs left to another stream's kernels
  if (pairs < 1) pairs = 1;
  if (const int v = option(kOptGrid) / 2) {
    if (v >= 1 && v < pairs) pairs = v;
  }
  const unsigned grid = 2u * static_cast<unsigned>(items < pairs ? items : pairs);
  const double exec_flops = static_cast<double>(g.n_problems) * g.n_rt * g.n_ct * (2.0 * 2 * kTileM * kTileN) *
                            (g.kb_count * static_cast<double>(kBlockK));
  void* tok = profile_begin(stream);
  pair_engine2_kernel<Epi><<<grid, kEngineThreads, smem, stream>>>(g, epi);   // __cluster_dims__(2,1,1)
  profile_end(tok, stream, alg_pairs, exec_flops);
  return check_launch(what);
}

}  // namespace amb

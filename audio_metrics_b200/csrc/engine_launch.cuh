// Host-side launcher for the pair engine; include in the .cu that instantiates
// a given epilogue.
#pragma once
#include <mutex>

#include "epilogues.cuh"
#include "internal.cuh"

namespace amb {

template <class Epi>
int launch_engine(cudaStream_t stream, int dev, const EngineGeom& g, const Epi& epi, const char* what,
                  double alg_pairs = 0.0) {
  static std::once_flag once;  // per Epi instantiation
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(pair_engine_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(kEngineSmemBytes));
  });
  // the attribute is per device; set it again cheaply when several devices are in use
  if (attr_err == cudaSuccess)
    attr_err = cudaFuncSetAttribute(pair_engine_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(kEngineSmemBytes));
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "cudaFuncSetAttribute(pair_engine)");
  const long long items = static_cast<long long>(g.n_problems) * g.n_rt * g.n_split;
  if (items <= 0) return AMB_OK;
  const int sms = sm_count(dev);
  const unsigned grid = static_cast<unsigned>(items < sms ? items : sms);
  // executed MMA work: full tiles, three fp16 MMAs per k step
  const double exec_flops = static_cast<double>(g.n_problems) * g.n_rt * g.n_ct * (2.0 * kTileM * kTileN) *
                            (g.kb_count * static_cast<double>(kBlockK)) * 3.0;
  void* tok = profile_begin(stream);
  pair_engine_kernel<Epi><<<grid, kEngineThreads, kEngineSmemBytes, stream>>>(g, epi);
  profile_end(tok, stream, alg_pairs, exec_flops);
  return check_launch(what);
}

}  // namespace amb

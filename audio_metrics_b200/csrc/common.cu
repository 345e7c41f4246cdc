// Error plumbing, device guard, packed-blob layout, pack entry point and the
// debug dot-matrix entry point.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "engine_launch.cuh"

namespace amb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return AMB_OK;
  return set_error(AMB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), what);
}

DeviceGuard::DeviceGuard(int dev) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error(AMB_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    (void)cudaGetLastError();
    ok = false;
    return;
  }
  if (dev < 0 || dev >= n) {
    set_error(AMB_ERR_ARG, "device index %d out of range [0,%d)", dev, n);
    ok = false;
    return;
  }
  if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  if (prev != dev && cudaSetDevice(dev) != cudaSuccess) {
    set_error(AMB_ERR_CUDA, "cudaSetDevice(%d) failed", dev);
    ok = false;
  }
}
DeviceGuard::~DeviceGuard() {
  if (ok && prev >= 0) {
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
}

int sm_count(int dev) {
  static std::mutex mu;
  static int cache[64] = {0};
  std::lock_guard<std::mutex> lk(mu);
  if (dev >= 0 && dev < 64 && cache[dev]) return cache[dev];
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 0 && dev < 64) cache[dev] = n;
  return n;
}

struct OptDef { const char* name; const char* env; int dflt; int lo, hi; };
static const OptDef kOptDefs[kOptCount] = {
    {"jacobi_block", "AMB_JACOBI_BS", 0, 0, 16},
    {"fad_ctas", "AMB_FAD_CTAS", 0, 0, 4096},
    {"engine_reserve_sms", "AMB_RESERVE_SMS", 0, 0, 128},
    {"fad_method", "AMB_FAD_METHOD", 0, 0, 1},
    {"engine_passes", "AMB_PASSES", 0, 0, 3},
    {"engine_cta2", "AMB_CTA2", -1, -1, 1},
    {"engine_static", "AMB_SCHED", 0, 0, 1},
    {"engine_stages", "AMB_STAGES", 0, 0, 8},
    {"engine_grid", "AMB_GRID", 0, 0, 4096},
    {"tail_split", "AMB_TAIL_SPLIT", 0, 0, 64},
    {"topk_split", "AMB_TOPK_SPLIT", 0, 0, 64},
    {"count_split", "AMB_COUNT_SPLIT", 0, 0, 64},
    {"debug_single", "AMB_DEBUG_SINGLE", 0, 0, 2},
    {"cov_dfma", "AMB_COV", 0, 0, 1},
    {"fad_factor_eig", "AMB_FAD_FACTOR", 0, 0, 1},
    {"jacobi_flat", "AMB_JACOBI", 0, 0, 1},
    {"fad_debug", "AMB_FAD_DEBUG", 0, 0, 1},
};
static std::atomic<int> g_opts[kOptCount];
static std::once_flag g_opts_once;
static void init_options() {
  for (int i = 0; i < kOptCount; ++i) {
    int v = kOptDefs[i].dflt;
    if (const char* e = getenv(kOptDefs[i].env)) {
      // numeric values as given; a word ("static", "dfma", "eig", "flat") switches the option on
      if ((e[0] >= '0' && e[0] <= '9') || e[0] == '-') v = atoi(e);
      else if (e[0]) v = 1;
      if (v < kOptDefs[i].lo || v > kOptDefs[i].hi) v = kOptDefs[i].dflt;
    }
    g_opts[i].store(v, std::memory_order_relaxed);
  }
}
int option(Opt o) {
  std::call_once(g_opts_once, init_options);
  return g_opts[o].load(std::memory_order_relaxed);
}

struct ProfRec { cudaEvent_t e0, e1; double pairs, flops; };
static std::mutex g_prof_mu;
static std::atomic<int> g_prof_on{0};
static std::vector<ProfRec> g_prof;

void* profile_begin(cudaStream_t stream) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return nullptr;
  ProfRec* r = new ProfRec{};
  if (cudaEventCreate(&r->e0) != cudaSuccess || cudaEventCreate(&r->e1) != cudaSuccess) { delete r; return nullptr; }
  cudaEventRecord(r->e0, stream);
  return r;
}
void profile_end(void* token, cudaStream_t stream, double alg_pairs, double exec_flops) {
  if (!token) return;
  ProfRec* r = static_cast<ProfRec*>(token);
  cudaEventRecord(r->e1, stream);
  r->pairs = alg_pairs;
  r->flops = exec_flops;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(*r);
  delete r;
}

PackedLayout packed_layout(long long n_rows, int d) {
  PackedLayout L;
  L.rows_pad = round_up_ll(n_rows > 0 ? n_rows : 1, kRowPad);
  L.kpad = static_cast<int>(round_up_ll(d, kBlockK));
  L.kb_count = L.kpad / kBlockK;
  L.plane_halfs = L.rows_pad * L.kpad;
  L.off_lo = static_cast<size_t>(L.plane_halfs) * 2;
  L.off_inv = L.off_lo * 2;
  L.off_norm = L.off_inv + static_cast<size_t>(L.rows_pad) * 4;
  L.off_rho = L.off_norm + static_cast<size_t>(L.rows_pad) * 4;
  L.off_exp = L.off_rho + static_cast<size_t>(L.rows_pad) * 4;
  L.off_cmin = L.off_exp + static_cast<size_t>(L.rows_pad) * 4;
  L.bytes = L.off_cmin + static_cast<size_t>(L.rows_pad / 32) * 4;
  L.bytes = static_cast<size_t>(round_up_ll(static_cast<long long>(L.bytes), 256));
  return L;
}

PackedPtrs packed_ptrs(void* blob, long long n_rows, int d) {
  PackedLayout L = packed_layout(n_rows, d);
  PackedPtrs p;
  uint8_t* b = static_cast<uint8_t*>(blob);
  p.planes = reinterpret_cast<__half*>(b);
  p.plane_halfs = L.plane_halfs;
  p.inv_scale = reinterpret_cast<float*>(b + L.off_inv);
  p.norm = reinterpret_cast<float*>(b + L.off_norm);
  p.rho = reinterpret_cast<float*>(b + L.off_rho);
  p.row_exp = reinterpret_cast<int*>(b + L.off_exp);
  p.cmin = reinterpret_cast<float*>(b + L.off_cmin);
  p.rows_pad = L.rows_pad;
  p.kb_count = L.kb_count;
  return p;
}

}  // namespace amb

using namespace amb;

extern "C" {

int amb_version(void) { return 100; }
const char* amb_last_error(void) { return g_err; }
long long amb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int amb_set_option(const char* name, int value) {
  std::call_once(g_opts_once, init_options);
  for (int i = 0; name && i < kOptCount; ++i) {
    if (strcmp(name, kOptDefs[i].name) != 0) continue;
    if (value < kOptDefs[i].lo || value > kOptDefs[i].hi)
      return set_error(AMB_ERR_ARG, "amb_set_option: %s must be in [%d, %d]", name, kOptDefs[i].lo, kOptDefs[i].hi);
    if (i == kOptJacobiBlock && value != 0 && value != 4 && value != 8 && value != 16)
      return set_error(AMB_ERR_ARG, "amb_set_option: jacobi_block must be 0, 4, 8 or 16");
    g_opts[i].store(value, std::memory_order_relaxed);
    return AMB_OK;
  }
  return set_error(AMB_ERR_ARG, "amb_set_option: unknown option '%s'", name ? name : "(null)");
}

int amb_get_option(const char* name) {
  std::call_once(g_opts_once, init_options);
  for (int i = 0; name && i < kOptCount; ++i)
    if (strcmp(name, kOptDefs[i].name) == 0) return g_opts[i].load(std::memory_order_relaxed);
  return set_error(AMB_ERR_ARG, "amb_get_option: unknown option '%s'", name ? name : "(null)");
}

int amb_profile_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
  return AMB_OK;
}

// out[0] = engine launches, out[1] = total device ms, out[2] = algorithmic pairs,
// out[3] = executed MMA flops.  Waits for the recorded events; clears the log.
int amb_profile_read(double* out) {
  if (!out) return set_error(AMB_ERR_ARG, "amb_profile_read: null");
  std::vector<ProfRec> recs;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    recs.swap(g_prof);
  }
  out[0] = out[1] = out[2] = out[3] = 0.0;
  for (ProfRec& r : recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      out[0] += 1.0;
      out[1] += ms;
      out[2] += r.pairs;
      out[3] += r.flops;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  return AMB_OK;
}

size_t amb_packed_bytes(long long n, int d) {
  if (n < 0 || d <= 0) return 0;
  return packed_layout(n, d).bytes;
}

int amb_pack(int dev, amb_stream_t stream, const void* X, int dtype, long long n, int d,
             long long ld, void* packed) {
  if (!packed || (n > 0 && !X) || n < 0 || d <= 0 || ld < d)
    return set_error(AMB_ERR_ARG, "amb_pack: bad argument (n=%lld d=%d ld=%lld)", n, d, ld);
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  PackedPtrs p = packed_ptrs(packed, n, d);
  return launch_pack(static_cast<cudaStream_t>(stream), X, dtype, ld, d, n, nullptr, n, 0,
                     p.rows_pad, p.planes, p.plane_halfs, p.kb_count, p.inv_scale, p.norm, p.rho, p.row_exp, p.cmin);
}

int amb_debug_dot_matrix(int dev, amb_stream_t stream, const void* packed_a, long long na,
                         const void* packed_b, long long nb, int d, float* C, long long ldc,
                         unsigned lbo, unsigned sbo) {
  if (!packed_a || !packed_b || !C || na <= 0 || nb <= 0 || d <= 0 || ldc < 0)
    return set_error(AMB_ERR_ARG, "amb_debug_dot_matrix: bad argument");
  DeviceGuard guard(dev);
  if (!guard.ok) return AMB_ERR_CUDA;
  PackedPtrs a = packed_ptrs(const_cast<void*>(packed_a), na, d);
  PackedPtrs b = packed_ptrs(const_cast<void*>(packed_b), nb, d);
  EngineGeom g{};
  g.a_planes = a.planes;
  g.b_planes = b.planes;
  g.a_plane_halfs = a.plane_halfs;
  g.b_plane_halfs = b.plane_halfs;
  g.kb_count = a.kb_count;
  g.a_rb0 = nullptr;
  g.b_rb0 = nullptr;
  g.n_problems = 1;
  g.n_rt = static_cast<int>(a.rows_pad / kTileM);
  g.n_ct = static_cast<int>(b.rows_pad / kTileN);
  g.n_split = 1;
  g.lbo_bytes = lbo ? lbo : 128;
  g.sbo_bytes = sbo ? sbo : 512;
  DumpEpi epi{a.inv_scale, b.inv_scale, C, ldc, na, nb, ldc == 0 ? 1 : 0};
  // option debug_single = 1: hi planes only through the single-pass kernel (11-bit operands)
  const int single = option(kOptDebugSingle);
  if (single == 2 && g.kb_count <= kMaxResidentKb) {   // ... on CTA pairs (cta_group::2)
    g.n_rt /= 2;                                             // rows_pad is a multiple of 256
    return launch_engine2(static_cast<cudaStream_t>(stream), dev, g, epi, "pair_engine2<dump>",
                          static_cast<double>(na) * nb);
  }
  if (single == 1 && g.kb_count <= kMaxResidentKb)
    return launch_engine1(static_cast<cudaStream_t>(stream), dev, g, epi, "pair_engine1<dump>",
                          static_cast<double>(na) * nb);
  return launch_engine(static_cast<cudaStream_t>(stream), dev, g, epi, "pair_engine<dump>",
                       static_cast<double>(na) * nb);
}

}  // extern "C"
